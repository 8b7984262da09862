#pragma once
#include "kernels.cuh"
