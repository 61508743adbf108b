// du_select.cu — F2a: per-row linear-interpolated quantile, bit-identical to torch.quantile(..., dim=1).
//
// Exact order statistics by MSB-first radix select on order-preserving uint32 keys.  This file holds
// the reference-grade exact path (one CTA per row, four 8-bit passes over the L2-resident row);
// the fused step (du_fused.cu) selects inside distributed shared memory instead.
#include "du_common.cuh"

namespace du {

struct RowSelectResult {
  uint32_t key_lo, key_hi;
  int has_nan;
};

// iterate over a row (global memory) with 128-bit loads when possible
template <typename F>
__device__ __forceinline__ void for_each_value(const float* __restrict__ r, int64_t n, bool vec, F&& f) {
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t g = threadIdx.x; g < n4; g += blockDim.x) {
      uint4 q = ldg_stream_128(r + 4 * g);
      f(__uint_as_float(q.x)); f(__uint_as_float(q.y)); f(__uint_as_float(q.z)); f(__uint_as_float(q.w));
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) f(__ldg(r + i));
  }
}

// Locate the bin of `hist[256]` that contains rank k (0-based, among the counted elements).
// Executed by warp 0; results in sh[0] = bin, sh[1] = count below the bin, sh[2] = count in the bin.
__device__ __forceinline__ void locate_bin_warp0(const uint32_t* hist, uint32_t k, uint32_t* sh) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  uint32_t c[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[8 * lane + j]; tot += c[j]; }
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  uint32_t excl = incl - tot;
  if (k >= excl && k < incl) {
    uint32_t cum = excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (k < cum + c[j]) { sh[0] = 8 * lane + j; sh[1] = cum; sh[2] = c[j]; break; }
      cum += c[j];
    }
  }
}

// One CTA selects the lo-th and hi-th (hi == lo or lo + 1) smallest keys of a row in global memory.
__device__ void select_row_exact(const float* __restrict__ r, int64_t n, bool vec, uint32_t lo, uint32_t hi,
                                 uint32_t* hist /*256*/, uint32_t* sh /*8*/, RowSelectResult& out) {
  uint32_t prefix = 0, pmask = 0, k = lo, below = 0, bincount = 0;
  if (threadIdx.x == 0) { sh[3] = 0; sh[4] = 0xffffffffu; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int j = threadIdx.x; j < 256; j += blockDim.x) hist[j] = 0;
    __syncthreads();
    uint32_t nan_seen = 0;
    for_each_value(r, n, vec, [&](float f) {
      if (pass == 0 && f != f) nan_seen = 1;
      uint32_t key = float_to_key(f);
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    });
    if (pass == 0 && __any_sync(0xffffffffu, nan_seen) && (threadIdx.x & 31) == 0) sh[3] = 1;
    __syncthreads();
    locate_bin_warp0(hist, k, sh);
    __syncthreads();
    prefix |= sh[0] << shift;
    pmask |= 255u << shift;
    below += sh[1];
    k -= sh[1];
    bincount = sh[2];
    __syncthreads();
  }
  out.key_lo = prefix;
  out.has_nan = (int)sh[3];
  const uint32_t count_le = below + bincount;  // number of keys <= key_lo
  if (hi < count_le) {
    out.key_hi = prefix;
  } else {
    // hi-th statistic is the smallest key above key_lo
    uint32_t best = 0xffffffffu;
    for_each_value(r, n, vec, [&](float f) {
      uint32_t key = float_to_key(f);
      if (key > prefix && key < best) best = key;
    });
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0) atomicMin(&sh[4], best);
    __syncthreads();
    out.key_hi = sh[4];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) quantile_rows_kernel(const float* __restrict__ u, int64_t B, int64_t n, int64_t stride,
                                                             uint32_t lo, uint32_t hi, float w, int lerp_fma,
                                                             float* __restrict__ thr_out, int32_t* __restrict__ rank_out,
                                                             float* __restrict__ val_out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sh[8];
  for (int64_t row = blockIdx.x; row < B; row += gridDim.x) {
    const float* r = u + row * stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(r) & 15) == 0) && ((n & 3) == 0);
    RowSelectResult res;
    select_row_exact(r, n, vec, lo, hi, hist, sh, res);
    if (threadIdx.x == 0) {
      float a = key_to_float(res.key_lo), b = key_to_float(res.key_hi);
      float t = lerp_torch(a, b, w, lerp_fma);
      if (res.has_nan) { t = __int_as_float(0x7fc00000); a = t; b = t; }
      thr_out[row] = t;
      if (rank_out) { rank_out[2 * row] = (int32_t)lo; rank_out[2 * row + 1] = (int32_t)hi; }
      if (val_out) { val_out[2 * row] = a; val_out[2 * row + 1] = b; }
    }
    __syncthreads();
  }
}

}  // namespace du

using namespace du;

extern "C" size_t du_quantile_scratch_bytes(int64_t B, int64_t n) {
  (void)n;
  return (size_t)(B > 0 ? B : 1) * 64;  // reserved: per-row counters of the multi-CTA path
}

extern "C" int du_quantile_threshold(const float* u, int64_t B, int64_t n, int64_t stride, float q, int lerp_fma,
                                     float* thr_out, int32_t* rank_out, float* val_out, void* scratch, size_t scratch_bytes,
                                     du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  (void)scratch; (void)scratch_bytes;
  if (B < 0) return set_error(DU_ERR_BAD_ARG, "du_quantile_threshold: negative batch");
  if (!(q >= 0.0f && q <= 1.0f)) return set_error(DU_ERR_BAD_ARG, "quantile() q values must be in the range [0, 1]");
  if (n <= 0) return set_error(DU_ERR_BAD_ARG, "quantile() input tensor must be non-empty");
  if (n > (int64_t)1 << 24) return set_error(DU_ERR_TOO_LARGE, "quantile() input tensor is too large");
  if (B == 0) return DU_OK;
  if (!u || !thr_out) return set_error(DU_ERR_BAD_ARG, "du_quantile_threshold: null pointer");
  if (!aligned(u, 4)) return set_error(DU_ERR_ALIGN, "du_quantile_threshold: misaligned input");
  // torch: ranks = q (fp32 tensor) * (n - 1)  ->  fp32 product; lo = floor, hi = ceil, weight = rank - lo
  volatile float rank = q * (float)(n - 1);
  float fl = floorf(rank), ce = ceilf(rank);
  uint32_t lo = (uint32_t)fl, hi = (uint32_t)ce;
  float w = rank - fl;
  unsigned grid = (unsigned)(B < 4096 ? B : 4096);
  quantile_rows_kernel<<<grid, 1024, 0, (cudaStream_t)stream>>>(u, B, n, stride, lo, hi, w, lerp_fma, thr_out, rank_out, val_out);
  DU_LAUNCH_CHECK("quantile_rows_kernel");
  return DU_OK;
}
