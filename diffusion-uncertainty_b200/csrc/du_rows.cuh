// du_rows.cuh — the generic "rows view" streaming driver shared by the elementwise kernels (du_step.cu, du_widen.cu).
#pragma once
#include "du_common.cuh"

namespace du {

// Generic driver: functor f.template run<VEC>(b, i) handles VEC elements of row b starting at i.
template <bool VECTOR, typename F>
__global__ void __launch_bounds__(256) rows_kernel(int64_t B, int64_t n, const __grid_constant__ F f) {
  constexpr int VEC = VECTOR ? 4 : 1;
  const int64_t groups = (n + VEC - 1) / VEC;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y)
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x)
      f.template run<VEC>(b, g * VEC);
}

template <typename F>
static int launch_rows(int64_t B, int64_t n, bool vec, const F& f, cudaStream_t st) {
  if (B == 0 || n == 0) return DU_OK;
  RowGrid g = row_grid(B, vec ? n / 4 : n, 256);
  if (vec) rows_kernel<true, F><<<g.grid, g.block, 0, st>>>(B, n, f);
  else rows_kernel<false, F><<<g.grid, g.block, 0, st>>>(B, n, f);
  DU_LAUNCH_CHECK("rows_kernel");
  return DU_OK;
}

template <int VEC>
__device__ __forceinline__ void loadv(const void* base, int64_t idx, int dt, float (&v)[VEC]) {
  if constexpr (VEC == 4) load4(base, idx, dt, v);
  else v[0] = load1(base, idx, dt);
}
template <int VEC>
__device__ __forceinline__ void storev(void* base, int64_t idx, int dt, const float (&v)[VEC]) {
  if constexpr (VEC == 4) store4(base, idx, dt, v);
  else store1(base, idx, dt, v[0]);
}

}  // namespace du
