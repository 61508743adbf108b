// du_select.cu — F2a: per-row linear-interpolated quantile, bit-identical to torch.quantile(..., dim=1).
//
// Exact order statistics by MSB-first radix select on order-preserving uint32 keys.  This file holds
// the unfused exact path, one CTA per row: a two-pass select with compaction (default) and the four-pass 8-bit select it falls
// back to under heavy ties; the fused step (du_fused.cu) selects inside distributed shared memory instead.
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

// ---------------------------------------------------------------------------------------------------------------------
// Two-pass select (the default): pass 1 histograms the top 11 key bits of the row (2048 bins) and locates the bin of rank lo;
// pass 2 compacts that bin's keys (typically 1-3 % of the row) into shared memory and remembers the smallest key above the bin;
// two more histogram levels (11 + 10 bits) over the LIST finish the select, warp 0 ranks the handful of keys that share 22 bits.
// The row is read twice (from L2 when it was just written) instead of four times, and everything after pass 2 works on ~1000 keys.
// Bins that do not fit the list (heavy ties, rows beyond ~300 k elements with dense bins) take the four-pass path above.
constexpr int Q_BINS = 2048;
constexpr int Q_LIST_CAP = 8192;
constexpr int Q_TINY_CAP = 1024;
constexpr int Q_THREADS = 1024;

// block-wide: bin of `hist[Q_BINS]` holding rank k; sh[0] = bin, sh[1] = count below, sh[2] = count in the bin, sh[5] = next
// non-empty bin above it (Q_BINS if none).  wsum: 32 words of scratch.
__device__ __forceinline__ void locate_block(const uint32_t* hist, uint32_t k, uint32_t* sh, uint32_t* wsum) {
  constexpr int PER = Q_BINS / Q_THREADS;   // 2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t c[PER], tot = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) { c[j] = hist[tid * PER + j]; tot += c[j]; }
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[warp] = incl;
  if (tid == 0) sh[5] = Q_BINS;
  __syncthreads();
  uint32_t wtot = wsum[lane], wincl = wtot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
    if (lane >= o) wincl += t;
  }
  const uint32_t excl = __shfl_sync(0xffffffffu, wincl - wtot, warp) + incl - tot;
  if (k >= excl && k < excl + tot) {
    uint32_t cum = excl;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (k < cum + c[j]) { sh[0] = tid * PER + j; sh[1] = cum; sh[2] = c[j]; break; }
      cum += c[j];
    }
  }
  __syncthreads();
  const uint32_t sel = sh[0];
  uint32_t nb = Q_BINS;
#pragma unroll
  for (int j = PER - 1; j >= 0; --j)
    if (c[j] != 0 && (uint32_t)(tid * PER + j) > sel) nb = tid * PER + j;
  nb = __reduce_min_sync(0xffffffffu, nb);
  if (lane == 0 && nb < (uint32_t)Q_BINS) atomicMin(&sh[5], nb);
  __syncthreads();
}

__global__ void __launch_bounds__(Q_THREADS) quantile_rows_kernel(const float* __restrict__ u, int64_t B, int64_t n, int64_t stride,
                                                                  uint32_t lo, uint32_t hi, float w, int lerp_fma,
                                                                  float* __restrict__ thr_out, int32_t* __restrict__ rank_out,
                                                                  float* __restrict__ val_out) {
  __shared__ uint32_t hist[Q_BINS];
  __shared__ uint32_t list[Q_LIST_CAP];
  __shared__ uint32_t tiny[Q_TINY_CAP];
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t sh[16];   // [0..2],[5] locate; [3] nan; [4] smallest key above the selected bin; [6] list count; [7] tiny count; [8],[9] result; [10] smallest key of the list's next level-1 bin
  const int tid = threadIdx.x, lane = tid & 31;
  const bool need_next = hi > lo;
  for (int64_t row = blockIdx.x; row < B; row += gridDim.x) {
    const float* r = u + row * stride;
    const bool vec = ((reinterpret_cast<uintptr_t>(r) & 15) == 0) && ((n & 3) == 0);
    for (int j = tid; j < Q_BINS; j += Q_THREADS) hist[j] = 0;
    if (tid < 16) sh[tid] = (tid == 4 || tid == 10) ? 0xffffffffu : 0u;
    __syncthreads();
    // ---- pass 1: top 11 bits
    uint32_t nan_seen = 0;
    for_each_value(r, n, vec, [&](float f) {
      nan_seen |= (f != f);
      atomicAdd(&hist[float_to_key(f) >> 21], 1u);
    });
    if (__any_sync(0xffffffffu, nan_seen) && lane == 0) sh[3] = 1;
    __syncthreads();
    locate_block(hist, lo, sh, wsum);
    const uint32_t d0 = sh[0], below0 = sh[1], cnt0 = sh[2];
    const int has_nan = (int)sh[3];
    uint32_t key_lo, key_hi;
    __syncthreads();
    if (cnt0 <= (uint32_t)Q_LIST_CAP) {
      // ---- pass 2: compact the bin, smallest key above it
      uint32_t best = 0xffffffffu;
      for_each_value(r, n, vec, [&](float f) {
        const uint32_t key = float_to_key(f);
        const uint32_t bin = key >> 21;
        if (bin == d0) list[atomicAdd(&sh[6], 1u)] = key;
        else if (bin > d0) best = min(best, key);
      });
      best = __reduce_min_sync(0xffffffffu, best);
      if (lane == 0 && best != 0xffffffffu) atomicMin(&sh[4], best);
      for (int j = tid; j < Q_BINS; j += Q_THREADS) hist[j] = 0;
      __syncthreads();
      const uint32_t above_bin = sh[4];
      // ---- level 1 over the list: bits 20..10
      for (uint32_t j = tid; j < cnt0; j += Q_THREADS) atomicAdd(&hist[(list[j] >> 10) & (Q_BINS - 1)], 1u);
      __syncthreads();
      const uint32_t k1 = lo - below0;
      locate_block(hist, k1, sh, wsum);
      const uint32_t d1 = sh[0], below1 = sh[1], cnt1 = sh[2], next1 = sh[5];
      const uint32_t k2 = k1 - below1;
      const uint32_t prefix22 = (d0 << 11) | d1;
      __syncthreads();
      if (cnt1 <= (uint32_t)Q_TINY_CAP) {
        const bool want_next_bin = need_next && (k2 + 1 >= cnt1) && next1 < (uint32_t)Q_BINS;
        const uint32_t next22 = (d0 << 11) | next1;
        uint32_t nbest = 0xffffffffu;
        for (uint32_t j = tid; j < cnt0; j += Q_THREADS) {
          const uint32_t key = list[j];
          if ((key >> 10) == prefix22) tiny[atomicAdd(&sh[7], 1u)] = key;
          else if (want_next_bin && (key >> 10) == next22) nbest = min(nbest, key);
        }
        if (want_next_bin) {
          nbest = __reduce_min_sync(0xffffffffu, nbest);
          if (lane == 0 && nbest != 0xffffffffu) atomicMin(&sh[10], nbest);
        }
        __syncthreads();
        if (tid < 32) {
          const uint32_t m = sh[7];
          uint32_t klo, khi;
          // successor outside the 22-bit group: smallest key of the next non-empty level-1 bin of the list, else of the row above bin d0
          uint32_t outside = above_bin;
          if (want_next_bin) outside = sh[10];
          if (m <= 32u) {
            const uint32_t mine = (lane < (int)m) ? tiny[lane] : 0xffffffffu;
            uint32_t rank = 0;
            for (uint32_t j = 0; j < m; ++j) {
              const uint32_t kj = __shfl_sync(0xffffffffu, mine, (int)j);
              rank += (kj < mine || (kj == mine && j < (uint32_t)lane)) ? 1u : 0u;
            }
            const uint32_t is_lo = __ballot_sync(0xffffffffu, lane < (int)m && rank == k2);
            const uint32_t is_hi = __ballot_sync(0xffffffffu, lane < (int)m && rank == k2 + 1u);
            klo = __shfl_sync(0xffffffffu, mine, __ffs((int)is_lo) - 1);
            khi = klo;
            if (need_next) khi = is_hi ? __shfl_sync(0xffffffffu, mine, __ffs((int)is_hi) - 1) : outside;
          } else {
            uint32_t ans = 0;
#pragma unroll 1
            for (int bit = 9; bit >= 0; --bit) {
              const uint32_t trial = (prefix22 << 10) | ans | (1u << bit);
              uint32_t c = 0;
              for (uint32_t j = lane; j < m; j += 32) c += (tiny[j] < trial);
              c = __reduce_add_sync(0xffffffffu, c);
              if (c <= k2) ans |= (1u << bit);
            }
            klo = (prefix22 << 10) | ans;
            khi = klo;
            if (need_next) {
              uint32_t le = 0, above = 0xffffffffu;
              for (uint32_t j = lane; j < m; j += 32) {
                const uint32_t key = tiny[j];
                le += (key <= klo);
                if (key > klo) above = min(above, key);
              }
              le = __reduce_add_sync(0xffffffffu, le);
              above = __reduce_min_sync(0xffffffffu, above);
              if (k2 + 1 < le) khi = klo;
              else if (above != 0xffffffffu) khi = above;
              else khi = outside;
            }
          }
          if (lane == 0) { sh[8] = klo; sh[9] = khi; }
        }
        __syncthreads();
        key_lo = sh[8]; key_hi = sh[9];
      } else {
        RowSelectResult res;
        select_row_exact(r, n, vec, lo, hi, hist, sh, res);   // heavy ties inside 22 bits: the four-pass path
        key_lo = res.key_lo; key_hi = res.key_hi;
      }
    } else {
      RowSelectResult res;
      select_row_exact(r, n, vec, lo, hi, hist, sh, res);
      key_lo = res.key_lo; key_hi = res.key_hi;
    }
    if (tid == 0) {
      float a = key_to_float(key_lo), b = key_to_float(key_hi);
      float t = lerp_torch(a, b, w, lerp_fma);
      if (has_nan) { t = __int_as_float(0x7fc00000); a = t; b = t; }
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
  quantile_rows_kernel<<<grid, Q_THREADS, 0, (cudaStream_t)stream>>>(u, B, n, stride, lo, hi, w, lerp_fma, thr_out, rank_out, val_out);
  DU_LAUNCH_CHECK("quantile_rows_kernel");
  return DU_OK;
}
