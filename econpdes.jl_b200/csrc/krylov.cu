// Placeholder until the multigrid-preconditioned Krylov solver lands.
#include "engine.cuh"

namespace econpdes {
namespace {
struct NotBuilt final : LinearSolver {
  int solve(const double*, const double*, double*, const econpdes_options&, int*, double*) override {
    throw std::runtime_error("2-D/3-D linear solver not built yet");
  }
};
}  // namespace
std::unique_ptr<LinearSolver> make_krylov_solver(Ctx&) { return std::make_unique<NotBuilt>(); }
}  // namespace econpdes
