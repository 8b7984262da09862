"""econpdes_b200: B200-native finite-scheme hot path of EconPDEs.jl behind its own API."""
from . import models  # noqa: F401
