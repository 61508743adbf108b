// du_moments.cu — F1: single-pass reduction over the M axis.
//
// One thread owns VEC consecutive elements of one image and streams the M score tensors through
// registers: every input byte is read exactly once with 128-bit, L1-bypassing loads, nothing is
// stacked or staged.  Variance modes use the shifted-data form about the first sample
// (d_m = x_m - x_0, exact or near-exact in fp32), so the accumulated quantities live on the scale
// of the spread rather than the mean and the fp32 result stays within a few ulp of the
// fp64-accumulated torch CPU value even when variance << mean^2.
#include "du_common.cuh"

namespace du {

struct MomentsParams {
  const void* scores[DU_MAX_M];
  int M;
  int mode;
  int64_t score_stride;
  const void* center;
  int64_t center_stride;
  int center_dtype;
  int unc_dtype;
  int64_t B, n;
  void* unc;
  int64_t unc_stride;
  float* mean;
  int64_t mean_stride;
  float inv_cnt, inv_cm1;  // 1/count and 1/(count-1) (count == 1 -> inf: 0 * inf = NaN like torch.var)
};

// Accumulator for one element.
struct Acc {
  float k;   // shift (first sample) for variance modes, centre for DU_MOM_CENTERED, 0 for RAW
  float s1;  // sum of d
  float s2;  // sum of d^2
};

// reciprocal-multiply instead of IEEE division (<= 1 ulp apart; the kernel is issue-bound otherwise), the same
// expressions as the fused step so that both paths produce the same bits
__device__ __forceinline__ float finish(const Acc& a, int mode, bool centered, float inv_cnt, float inv_cm1, float* mean_out) {
  if (mean_out) *mean_out = fmaf(a.s1, inv_cnt, a.k);
  if (mode == DU_MOM_CENTERED || mode == DU_MOM_RAW) return fmaf(a.s2, inv_cnt, 0.0f);
  if (mode == DU_MOM_PARTIAL_M2 && centered) return a.s2;
  const float m2 = m2_from_sums(a.s1, a.s2, inv_cnt);  // sum of squared deviations about the mean (shifted-data form)
  if (mode == DU_MOM_PARTIAL_M2) return m2;
  const float var = fmaf(m2, inv_cm1, 0.0f);
  return mode == DU_MOM_STD_UNBIASED ? sqrtf(var) : var;
}

// kernel: grid.x covers a row in groups of VEC elements, grid.y strides over images.
// VECTOR = false is the scalar fallback for unaligned / ragged views.
template <typename T, bool VECTOR, int MT>
__global__ void __launch_bounds__(256) moments_kernel(const __grid_constant__ MomentsParams p) {
  using SV = Vec16<T>;
  constexpr int VEC = VECTOR ? SV::VEC : 1;
  const int64_t groups = (p.n + VEC - 1) / VEC;
  const int mode = p.mode;
  const bool centered = (mode == DU_MOM_CENTERED) || (mode == DU_MOM_PARTIAL_M2 && p.center != nullptr);
  const bool shifted = !(centered || mode == DU_MOM_RAW);
  const bool extra = (mode == DU_MOM_VAR_WITH_CENTER);

  for (int64_t b = blockIdx.y; b < p.B; b += gridDim.y) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = g * VEC;
      Acc acc[VEC];
      if constexpr (VECTOR) {
        // every 16-byte load of this group (centre + up to DU_LOAD_BATCH scores) is issued before the first use
        float c[VEC], k[VEC], s1[VEC], s2[VEC];
        const bool same_dt = (p.center != nullptr) && (p.center_dtype == SV::DT);
        uint4 raw_c = make_uint4(0u, 0u, 0u, 0u);
        if (same_dt) raw_c = SV::load(p.center, b * p.center_stride + i);
        const int centre_mode = (p.center == nullptr) ? 0 : (centered ? 1 : (extra ? 2 : 0));
        if (centre_mode && !same_dt) {  // mixed dtypes (e.g. fp32 eps with fp16 scores): scalar loads of the centre
#pragma unroll
          for (int e = 0; e < VEC; ++e) c[e] = load1(p.center, b * p.center_stride + i + e, p.center_dtype);
        }
        if constexpr (MT > 0) accumulate_scores_ct<T, MT>(p.scores, b * p.score_stride, (uint32_t)i * (uint32_t)sizeof(T), raw_c, centre_mode, !same_dt, shifted, c, k, s1, s2);
        else accumulate_scores<T>(p.scores, p.M, b * p.score_stride + i, raw_c, centre_mode, !same_dt, shifted, c, k, s1, s2);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { acc[e].k = k[e]; acc[e].s1 = s1[e]; acc[e].s2 = s2[e]; }
      } else {
        float c0 = 0.0f;
        if (p.center != nullptr) c0 = load1(p.center, b * p.center_stride + i, p.center_dtype);
        acc[0].k = centered ? c0 : 0.0f; acc[0].s1 = 0.0f; acc[0].s2 = 0.0f;
        const int64_t off = b * p.score_stride + i;
        for (int m = 0; m < p.M; ++m) {
          const float x = load1(p.scores[m], off, SV::DT);
          if (shifted && m == 0) acc[0].k = x;
          const float d = x - acc[0].k;
          acc[0].s1 += d;
          acc[0].s2 = fmaf(d, d, acc[0].s2);
        }
        if (extra) {
          const float d = c0 - acc[0].k;
          acc[0].s1 += d;
          acc[0].s2 = fmaf(d, d, acc[0].s2);
        }
      }
      float u[VEC], mu[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) u[e] = finish(acc[e], mode, centered, p.inv_cnt, p.inv_cm1, p.mean ? &mu[e] : nullptr);

      if constexpr (VECTOR) {
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {
          float t4[4] = {u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]};
          store4(p.unc, b * p.unc_stride + i + 4 * h, p.unc_dtype, t4);
          if (p.mean) {
            float m4[4] = {mu[4 * h], mu[4 * h + 1], mu[4 * h + 2], mu[4 * h + 3]};
            store4(p.mean, b * p.mean_stride + i + 4 * h, DU_F32, m4);
          }
        }
      } else {
        store1(p.unc, b * p.unc_stride + i, p.unc_dtype, u[0]);
        if (p.mean) p.mean[b * p.mean_stride + i] = mu[0];
      }
    }
  }
}

template <typename T>
static int launch_moments(const MomentsParams& p, bool vec, cudaStream_t st) {
  constexpr int VEC = Vec16<T>::VEC;
  const int threads = 256;
  int64_t groups = vec ? (p.n / VEC) : p.n;
  RowGrid g = row_grid(p.B, groups, threads);
  if (!vec) moments_kernel<T, false, 0><<<g.grid, g.block, 0, st>>>(p);
  else switch (p.M) {  // compile-time M for the BASELINE sample counts, batched runtime-M loop otherwise
    case 4: moments_kernel<T, true, 4><<<g.grid, g.block, 0, st>>>(p); break;
    case 5: moments_kernel<T, true, 5><<<g.grid, g.block, 0, st>>>(p); break;
    case 8: moments_kernel<T, true, 8><<<g.grid, g.block, 0, st>>>(p); break;
    case 16: moments_kernel<T, true, 16><<<g.grid, g.block, 0, st>>>(p); break;
    default: moments_kernel<T, true, 0><<<g.grid, g.block, 0, st>>>(p); break;
  }
  DU_LAUNCH_CHECK("moments_kernel");
  return DU_OK;
}

// ---- Chan merge of per-rank partials -----------------------------------------------------------------
struct MergeParams {
  const float* means[DU_MAX_M];
  const float* m2s[DU_MAX_M];
  int counts[DU_MAX_M];
  int R, mode;
  int64_t N;
  float* unc;
  float* mean;
};

__global__ void __launch_bounds__(256) moments_merge_kernel(const __grid_constant__ MergeParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += (int64_t)gridDim.x * blockDim.x) {
    if (p.mode == DU_MOM_CENTERED) {
      float s = 0.0f; int tot = 0;
      for (int r = 0; r < p.R; ++r) { s += p.m2s[r][i]; tot += p.counts[r]; }
      p.unc[i] = s / (float)tot;
      continue;
    }
    // pairwise Chan update in rank order: M2 += M2_r + delta^2 * n_a n_r / (n_a + n_r)
    float mean = p.means[0][i], m2 = p.m2s[0][i];
    float na = (float)p.counts[0];
    for (int r = 1; r < p.R; ++r) {
      float nr = (float)p.counts[r];
      if (nr == 0.0f) continue;
      float delta = p.means[r][i] - mean;
      float tot = na + nr;
      m2 = m2 + p.m2s[r][i] + delta * delta * (na * nr / tot);
      mean = mean + delta * (nr / tot);
      na = tot;
    }
    float var = m2 / (na - 1.0f);
    p.unc[i] = (p.mode == DU_MOM_STD_UNBIASED) ? sqrtf(var) : var;
    if (p.mean) p.mean[i] = mean;
  }
}

}  // namespace du

using namespace du;

extern "C" int du_moments(const void* const* scores, int M, int64_t score_stride, int score_dtype,
                          const void* center, int64_t center_stride, int center_dtype, int mode,
                          int64_t B, int64_t n, void* unc_out, int64_t unc_stride, int unc_dtype,
                          float* mean_out, int64_t mean_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!scores || M < 1 || M > DU_MAX_M) return set_error(DU_ERR_BAD_ARG, "du_moments: M=%d must be in [1,%d]", M, DU_MAX_M);
  if (B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_moments: negative size");
  if (B == 0 || n == 0) return DU_OK;
  if (!unc_out) return set_error(DU_ERR_BAD_ARG, "du_moments: null output");
  if (mode < DU_MOM_VAR_UNBIASED || mode > DU_MOM_PARTIAL_M2) return set_error(DU_ERR_BAD_ARG, "du_moments: bad mode %d", mode);
  if (!dtype_ok(score_dtype) || !dtype_ok(unc_dtype) || (center && !dtype_ok(center_dtype)))
    return set_error(DU_ERR_DTYPE, "du_moments: unsupported dtype");
  if ((mode == DU_MOM_CENTERED || mode == DU_MOM_VAR_WITH_CENTER) && !center)
    return set_error(DU_ERR_BAD_ARG, "du_moments: mode %d needs a centre tensor", mode);
  if (mode == DU_MOM_VAR_UNBIASED || mode == DU_MOM_RAW || mode == DU_MOM_STD_UNBIASED) center = nullptr;

  MomentsParams p{};
  const int es = dtype_size(score_dtype);
  const int vec = (score_dtype == DU_F32) ? 4 : 8;
  bool vec_ok = (n % vec == 0) && (score_stride % vec == 0);
  for (int m = 0; m < M; ++m) {
    if (!scores[m]) return set_error(DU_ERR_BAD_ARG, "du_moments: scores[%d] is null", m);
    if (!aligned(scores[m], es)) return set_error(DU_ERR_ALIGN, "du_moments: scores[%d] misaligned", m);
    p.scores[m] = scores[m];
    vec_ok = vec_ok && aligned(scores[m], 16);
  }
  if (center) vec_ok = vec_ok && aligned(center, (size_t)vec * dtype_size(center_dtype)) && (center_stride % vec == 0);
  vec_ok = vec_ok && aligned(unc_out, 4 * (size_t)dtype_size(unc_dtype)) && (unc_stride % 4 == 0);
  if (mean_out) vec_ok = vec_ok && aligned(mean_out, 16) && (mean_stride % 4 == 0);
  p.M = M; p.mode = mode; p.score_stride = score_stride;
  p.center = center; p.center_stride = center_stride; p.center_dtype = center_dtype;
  p.unc = unc_out; p.unc_stride = unc_stride; p.unc_dtype = unc_dtype;
  p.mean = mean_out; p.mean_stride = mean_stride;
  p.B = B; p.n = n;
  const int count = M + (mode == DU_MOM_VAR_WITH_CENTER ? 1 : 0);
  p.inv_cnt = 1.0f / (float)count;
  p.inv_cm1 = 1.0f / (float)(count - 1);
  cudaStream_t st = (cudaStream_t)stream;
  switch (score_dtype) {
    case DU_F32: return launch_moments<float>(p, vec_ok, st);
    case DU_F16: return launch_moments<__half>(p, vec_ok, st);
    default: return launch_moments<__nv_bfloat16>(p, vec_ok, st);
  }
}

extern "C" int du_moments_merge(const float* const* means, const float* const* m2s, const int* counts, int R,
                                int mode, int64_t N, float* unc_out, float* mean_out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (R < 1 || R > DU_MAX_M || !m2s || !counts || !unc_out || N < 0)
    return set_error(DU_ERR_BAD_ARG, "du_moments_merge: bad arguments (R=%d)", R);
  if (mode != DU_MOM_CENTERED && mode != DU_MOM_VAR_UNBIASED && mode != DU_MOM_STD_UNBIASED)
    return set_error(DU_ERR_BAD_ARG, "du_moments_merge: bad mode %d", mode);
  if (mode != DU_MOM_CENTERED && !means) return set_error(DU_ERR_BAD_ARG, "du_moments_merge: means required");
  if (N == 0) return DU_OK;
  MergeParams p{};
  for (int r = 0; r < R; ++r) {
    p.means[r] = means ? means[r] : nullptr;
    p.m2s[r] = m2s[r];
    p.counts[r] = counts[r];
  }
  p.R = R; p.mode = mode; p.N = N; p.unc = unc_out; p.mean = mean_out;
  int threads = 256;
  int64_t blocks = (N + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  moments_merge_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(p);
  DU_LAUNCH_CHECK("moments_merge_kernel");
  return DU_OK;
}
