// Shared helpers for libgfr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gfr_b200.h"

#define GFR_RETURN_IF_NULL(p) do { if ((p) == nullptr) return GFR_E_NULL; } while (0)

// Kernel launches report configuration errors immediately; execution errors surface at the caller's next sync.
static inline int gfr_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? GFR_OK : (int)e;
}

__host__ __device__ static inline int gfr_ceil_div(int a, int b) { return (a + b - 1) / b; }
