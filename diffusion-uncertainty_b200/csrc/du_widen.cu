// du_widen.cu — the rows either side of the core path:
//   N4  flip-based uncertainty: H-flip of the model input, (eps - flip(eps_hat))^2 with optional channel amax
//       (SU/scheduling_ddim_flip.py:486-493, SU/scheduling_ddim_flip_threshold.py:497-506)
//   F6  backward of the M-axis reduction, for the schedulers that differentiate the map through the score model
//       (SU/scheduling_ddim_uncertainty_grad.py:536-538, ..._mc_dropout_gradient.py:499-503, ..._model_gradient_guided.py:546-548,
//       PU/pipeline_sampler_class_conditional_uncertainty_guided_gradient.py:190-194)
//   N2  per-timestep / per-pixel threshold fitting: k-th smallest over the sample axis of the accumulated maps
//       (scripts/compute_threshold_pixel_wise.py:90-100, 143-152)
//   N3  per-image reductions of the accumulated map (scripts/uncertainty_benchmark_imagenet.py:314, scripts/compute_ause.py:128)
// All HBM-bound streaming kernels: 128-bit accesses where the views allow, every input read once.
#include "du_rows.cuh"

namespace du {

// ---- N4: flips ----------------------------------------------------------------------------------------------------------
// one thread = 4 consecutive elements of one image row (W % 4 == 0 on the vector path, so a group never straddles rows)
struct FlipF {
  const void* x; int64_t x_stride; int x_dtype;
  int64_t H, W;
  void* out; int64_t out_stride; int out_dtype;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    const int64_t w = i % W, hc = i / W, h = hc % H, c = hc / H;
    float v[VEC];
    loadv<VEC>(x, b * x_stride + (c * H + (H - 1 - h)) * W + w, x_dtype, v);
    storev<VEC>(out, b * out_stride + i, out_dtype, v);
  }
};

// u = (eps - flipH(f))^2 per element
struct FlipSqDiffF {
  const void* eps; int64_t eps_stride; int eps_dtype;
  const void* f; int64_t f_stride; int f_dtype;
  int64_t H, W;
  float* out; int64_t out_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    const int64_t w = i % W, hc = i / W, h = hc % H, c = hc / H;
    float e[VEC], g[VEC], o[VEC];
    loadv<VEC>(eps, b * eps_stride + i, eps_dtype, e);
    loadv<VEC>(f, b * f_stride + (c * H + (H - 1 - h)) * W + w, f_dtype, g);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { const float d = __fsub_rn(e[k], g[k]); o[k] = __fmul_rn(d, d); }
    storev<VEC>(out, b * out_stride + i, DU_F32, o);
  }
};

// u[b,0,h,w] = max_c (eps - flipH(f))^2, NaN-propagating like torch.amax; rows of the output hold H*W elements
struct FlipSqDiffAmaxF {
  const void* eps; int64_t eps_stride; int eps_dtype;
  const void* f; int64_t f_stride; int f_dtype;
  int64_t C, H, W;
  float* out; int64_t out_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {   // i indexes the [H,W] plane
    const int64_t w = i % W, h = i / W;
    float m[VEC];
    bool nan[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { m[k] = -INFINITY; nan[k] = false; }
    for (int64_t c = 0; c < C; ++c) {
      float e[VEC], g[VEC];
      loadv<VEC>(eps, b * eps_stride + (c * H + h) * W + w, eps_dtype, e);
      loadv<VEC>(f, b * f_stride + (c * H + (H - 1 - h)) * W + w, f_dtype, g);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float d = __fsub_rn(e[k], g[k]);
        const float u = __fmul_rn(d, d);
        nan[k] |= (u != u);
        m[k] = fmaxf(m[k], u);
      }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) if (nan[k]) m[k] = __int_as_float(0x7fc00000);
    storev<VEC>(out, b * out_stride + i, DU_F32, m);
  }
};

// ---- F6: backward of the M-axis reduction --------------------------------------------------------------------------------
// var (unbiased, over M samples):      d u / d s_m = 2 (s_m - mean) / (M - 1)
// centered second moment about c:      d u / d s_m = 2 (s_m - c) / M ,   d u / d c = -2 (mean - c)   [= -sum_m d u / d s_m]
// var over M samples + c (M+1 values): d u / d s_m = 2 (s_m - mean') / M,  d u / d c = 2 (c - mean') / M
// grad_m = g_u * du/ds_m.  Two passes over the M tensors per element (mean, then gradients); the second pass hits L2.
struct MomentsBwdF {
  const void* scores[DU_MAX_M]; int M; int64_t score_stride; int score_dtype;
  const void* center; int64_t center_stride; int center_dtype;
  const void* gu; int64_t gu_stride; int gu_dtype;
  int mode;
  void* grads[DU_MAX_M]; int64_t grad_stride; int grad_dtype;
  void* grad_center; int64_t gc_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float g[VEC], c[VEC], sum[VEC];
    loadv<VEC>(gu, b * gu_stride + i, gu_dtype, g);
    const bool has_c = (mode != DU_MOM_VAR_UNBIASED && mode != DU_MOM_STD_UNBIASED);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { c[k] = 0.0f; sum[k] = 0.0f; }
    if (has_c) loadv<VEC>(center, b * center_stride + i, center_dtype, c);
    for (int m = 0; m < M; ++m) {
      float s[VEC];
      loadv<VEC>(scores[m], b * score_stride + i, score_dtype, s);
#pragma unroll
      for (int k = 0; k < VEC; ++k) sum[k] += s[k];
    }
    float ref[VEC], coef;   // du/ds_m = coef * (s_m - ref)
    float cstd[VEC];        // STD_UNBIASED: per-element factor 1 / ((M-1) std)
    if (mode == DU_MOM_VAR_UNBIASED || mode == DU_MOM_STD_UNBIASED) {
      coef = 2.0f / (float)(M - 1);
#pragma unroll
      for (int k = 0; k < VEC; ++k) ref[k] = sum[k] / (float)M;
      if (mode == DU_MOM_STD_UNBIASED) {   // d std / d s_m = (s_m - mean) / ((M-1) std)   (generate_samples.py:941-943)
        float ss[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) ss[k] = 0.0f;
        for (int m = 0; m < M; ++m) {
          float s[VEC];
          loadv<VEC>(scores[m], b * score_stride + i, score_dtype, s);
#pragma unroll
          for (int k = 0; k < VEC; ++k) { const float d = s[k] - ref[k]; ss[k] = fmaf(d, d, ss[k]); }
        }
        coef = 1.0f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) cstd[k] = 1.0f / ((float)(M - 1) * sqrtf(ss[k] / (float)(M - 1)));
      }
    } else if (mode == DU_MOM_CENTERED) {
      coef = 2.0f / (float)M;
#pragma unroll
      for (int k = 0; k < VEC; ++k) ref[k] = c[k];
    } else {
      coef = 2.0f / (float)M;
#pragma unroll
      for (int k = 0; k < VEC; ++k) ref[k] = (sum[k] + c[k]) / (float)(M + 1);
    }
    for (int m = 0; m < M; ++m) {
      float s[VEC], o[VEC];
      loadv<VEC>(scores[m], b * score_stride + i, score_dtype, s);
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = g[k] * ((mode == DU_MOM_STD_UNBIASED ? cstd[k] : coef) * (s[k] - ref[k]));
      if (grads[m]) storev<VEC>(grads[m], b * grad_stride + i, grad_dtype, o);
    }
    if (grad_center && has_c) {
      float o[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        if (mode == DU_MOM_CENTERED) o[k] = g[k] * (-coef * (sum[k] - (float)M * c[k]));
        else o[k] = g[k] * (coef * (c[k] - ref[k]));
      }
      storev<VEC>(grad_center, b * gc_stride + i, grad_dtype, o);
    }
  }
};

// ---- N2: k-th smallest over the sample axis, one column per thread (coalesced rows), bitwise search on ordered keys -------
// x: [N rows, n columns], row stride in elements.  32 passes over the column, each counting keys below a trial value; the
// working set of a CTA (256 columns x N rows) is re-read from L2.  NaN sorts last (torch.argsort).  Exact for any N.
template <typename T>
__global__ void __launch_bounds__(256) column_kth_kernel(const T* __restrict__ x, int64_t N, int64_t n, int64_t row_stride,
                                                         int64_t k, T* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  auto key_of = [](T v) -> uint32_t {
    const float f = (float)v;
    if (f != f) return 0xffffffffu;          // NaN last
    return float_to_key(f);
  };
  uint32_t ans = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t trial = ans | (1u << bit);
    int64_t below = 0;
    for (int64_t r = 0; r < N; ++r) below += (key_of(x[r * row_stride + j]) < trial) ? 1 : 0;
    if (below <= k) ans = trial;
  }
  // ans is the key of the k-th smallest value; return the stored element with that key (keeps -0.0 / NaN payloads as stored)
  T res = x[j];
  for (int64_t r = 0; r < N; ++r) {
    const T v = x[r * row_stride + j];
    if (key_of(v) == ans) { res = v; break; }
  }
  out[j] = res;
}

// ---- N3: per-row sums (fp32 accumulation in a fixed order: deterministic) ------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS) row_sum_kernel(const void* x, int64_t x_stride, int dt, int64_t n, bool vec, float* out) {
  const int64_t b = blockIdx.x;
  float acc = 0.0f;
  if (vec) {
    // four 16-byte loads in flight per thread (one CTA per row: the kernel lives on memory-level parallelism); the summation
    // order is fixed by (thread, trip), so the result is deterministic
    const int64_t n4 = n / 4;
    for (int64_t g0 = threadIdx.x; g0 < n4; g0 += 4 * THREADS) {
      float v[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t g = g0 + (int64_t)j * THREADS;
        if (g < n4) load4(x, b * x_stride + 4 * g, dt, v[j]);
        else { v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0.0f; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += THREADS) acc += load1(x, b * x_stride + i, dt);
  }
  __shared__ float part[THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < THREADS / 32) ? part[threadIdx.x] : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[b] = v;
  }
}

// sum over the T slots of the [B, T, n] accumulation buffer -> [B, n]
struct SlotSumF {
  const void* x; int64_t x_stride; int64_t slot_stride; int x_dtype; int T;
  float* out; int64_t out_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
    for (int t = 0; t < T; ++t) {
      float v[VEC];
      loadv<VEC>(x, b * x_stride + t * slot_stride + i, x_dtype, v);
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] += v[k];
    }
    storev<VEC>(out, b * out_stride + i, DU_F32, acc);
  }
};

// DPM-Solver multistep update (first / second order): out = (a*s + b*m0) + c*(k*(m0 - m1)), one rounding per operation in
// the order the reference's Python expression evaluates (subtractions folded into the signs of b and c: x - y == x + (-y)
// bit for bit).  m1 == nullptr: first order.
struct DpmUpdateF {
  const void* s; int64_t s_stride; int s_dtype;
  const void* m0; int64_t m0_stride; int m0_dtype;
  const void* m1; int64_t m1_stride; int m1_dtype;
  float a, b_, c, k;
  void* out; int64_t out_stride; int out_dtype;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float sv[VEC], v0[VEC], v1[VEC], o[VEC];
    loadv<VEC>(s, b * s_stride + i, s_dtype, sv);
    loadv<VEC>(m0, b * m0_stride + i, m0_dtype, v0);
    if (m1) loadv<VEC>(m1, b * m1_stride + i, m1_dtype, v1);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = __fadd_rn(__fmul_rn(a, sv[e]), __fmul_rn(b_, v0[e]));
      if (m1) t = __fadd_rn(t, __fmul_rn(c, __fmul_rn(k, __fsub_rn(v0[e], v1[e]))));
      o[e] = t;
    }
    storev<VEC>(out, b * out_stride + i, out_dtype, o);
  }
};

static bool view_ok(const void* p, int dt) { return p != nullptr && dtype_ok(dt); }

}  // namespace du

using namespace du;

extern "C" int du_flip_h(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t C, int64_t H, int64_t W,
                         void* out, int64_t out_stride, int out_dtype, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!view_ok(x, x_dtype) || !view_ok(out, out_dtype) || B < 0 || C < 0 || H < 0 || W < 0) return set_error(DU_ERR_BAD_ARG, "du_flip_h: bad arguments");
  FlipF f{x, x_stride, x_dtype, H, W, out, out_stride, out_dtype};
  const int64_t n = C * H * W;
  const bool vec = (W % 4 == 0) && vec4_ok(x, x_stride, x_dtype) && vec4_ok(out, out_stride, out_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_flip_sqdiff(const void* eps, int64_t eps_stride, int eps_dtype, const void* flipped, int64_t f_stride, int f_dtype,
                              int64_t B, int64_t C, int64_t H, int64_t W, int channel_amax, float* out, int64_t out_stride,
                              du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!view_ok(eps, eps_dtype) || !view_ok(flipped, f_dtype) || !out || B < 0 || C < 0 || H < 0 || W < 0)
    return set_error(DU_ERR_BAD_ARG, "du_flip_sqdiff: bad arguments");
  const bool vec = (W % 4 == 0) && vec4_ok(eps, eps_stride, eps_dtype) && vec4_ok(flipped, f_stride, f_dtype) && vec4_ok(out, out_stride, DU_F32);
  if (channel_amax) {
    FlipSqDiffAmaxF f{eps, eps_stride, eps_dtype, flipped, f_stride, f_dtype, C, H, W, out, out_stride};
    return launch_rows(B, H * W, vec, f, (cudaStream_t)stream);
  }
  FlipSqDiffF f{eps, eps_stride, eps_dtype, flipped, f_stride, f_dtype, H, W, out, out_stride};
  return launch_rows(B, C * H * W, vec, f, (cudaStream_t)stream);
}

extern "C" int du_moments_backward(const void* const* scores, int M, int64_t score_stride, int score_dtype, const void* center,
                                   int64_t center_stride, int center_dtype, int mode, const void* grad_u, int64_t gu_stride,
                                   int gu_dtype, int64_t B, int64_t n, void* const* grad_scores, int64_t grad_stride,
                                   int grad_dtype, void* grad_center, int64_t gc_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!scores || !grad_scores || M < 1 || M > DU_MAX_M) return set_error(DU_ERR_BAD_ARG, "du_moments_backward: M must be in [1,%d]", DU_MAX_M);
  if (mode != DU_MOM_VAR_UNBIASED && mode != DU_MOM_CENTERED && mode != DU_MOM_VAR_WITH_CENTER && mode != DU_MOM_STD_UNBIASED)
    return set_error(DU_ERR_BAD_ARG, "du_moments_backward: mode %d has no backward", mode);
  const bool no_centre = (mode == DU_MOM_VAR_UNBIASED || mode == DU_MOM_STD_UNBIASED);
  if (!dtype_ok(score_dtype) || !dtype_ok(grad_dtype) || !view_ok(grad_u, gu_dtype)) return set_error(DU_ERR_DTYPE, "du_moments_backward: bad dtype / null grad_u");
  if (!no_centre && !view_ok(center, center_dtype)) return set_error(DU_ERR_BAD_ARG, "du_moments_backward: mode needs a centre");
  if (B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_moments_backward: bad sizes");
  MomentsBwdF f{};
  bool vec = (n % 4 == 0) && vec4_ok(grad_u, gu_stride, gu_dtype) && (no_centre || vec4_ok(center, center_stride, center_dtype)) &&
             vec4_ok(grad_center, gc_stride, grad_dtype);
  for (int m = 0; m < M; ++m) {
    if (!scores[m]) return set_error(DU_ERR_BAD_ARG, "du_moments_backward: scores[%d] is null", m);
    f.scores[m] = scores[m];
    f.grads[m] = grad_scores[m];
    vec = vec && vec4_ok(scores[m], score_stride, score_dtype) && vec4_ok(grad_scores[m], grad_stride, grad_dtype);
  }
  f.M = M; f.score_stride = score_stride; f.score_dtype = score_dtype;
  f.center = center; f.center_stride = center_stride; f.center_dtype = center_dtype;
  f.gu = grad_u; f.gu_stride = gu_stride; f.gu_dtype = gu_dtype;
  f.mode = mode; f.grad_stride = grad_stride; f.grad_dtype = grad_dtype;
  f.grad_center = grad_center; f.gc_stride = gc_stride;
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_column_kth(const void* x, int dtype, int64_t N, int64_t n, int64_t row_stride, int64_t k, void* out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!x || !out || N < 1 || n < 0 || k < 0 || k >= N) return set_error(DU_ERR_BAD_ARG, "du_column_kth: need N >= 1 rows and 0 <= k < N");
  if (n == 0) return DU_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DU_F32) column_kth_kernel<float><<<grid, 256, 0, st>>>((const float*)x, N, n, row_stride, k, (float*)out);
  else if (dtype == DU_F16) column_kth_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, N, n, row_stride, k, (__half*)out);
  else if (dtype == DU_BF16) column_kth_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, N, n, row_stride, k, (__nv_bfloat16*)out);
  else return set_error(DU_ERR_DTYPE, "du_column_kth: unsupported dtype");
  DU_LAUNCH_CHECK("column_kth_kernel");
  return DU_OK;
}

extern "C" int du_row_sum(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, float* out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!view_ok(x, x_dtype) || !out || B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_row_sum: bad arguments");
  if (B == 0) return DU_OK;
  const bool vec = (n % 4 == 0) && vec4_ok(x, x_stride, x_dtype);
  row_sum_kernel<512><<<(unsigned)B, 512, 0, (cudaStream_t)stream>>>(x, x_stride, x_dtype, n, vec, out);
  DU_LAUNCH_CHECK("row_sum_kernel");
  return DU_OK;
}

extern "C" int du_slot_sum(const void* x, int64_t x_stride, int64_t slot_stride, int x_dtype, int64_t B, int T, int64_t n,
                           float* out, int64_t out_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!view_ok(x, x_dtype) || !out || B < 0 || n < 0 || T < 0) return set_error(DU_ERR_BAD_ARG, "du_slot_sum: bad arguments");
  SlotSumF f{x, x_stride, slot_stride, x_dtype, T, out, out_stride};
  const bool vec = (n % 4 == 0) && vec4_ok(x, x_stride, x_dtype) && (slot_stride % 4 == 0) && vec4_ok(out, out_stride, DU_F32);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_dpm_solver_update(const void* sample, int64_t s_stride, int s_dtype, const void* m0, int64_t m0_stride, int m0_dtype,
                                    const void* m1, int64_t m1_stride, int m1_dtype, float a, float b, float c, float k,
                                    int64_t B, int64_t n, void* out, int64_t out_stride, int out_dtype, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!view_ok(sample, s_dtype) || !view_ok(m0, m0_dtype) || (m1 && !dtype_ok(m1_dtype)) || !view_ok(out, out_dtype) || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_dpm_solver_update: bad arguments");
  DpmUpdateF f{sample, s_stride, s_dtype, m0, m0_stride, m0_dtype, m1, m1_stride, m1_dtype, a, b, c, k, out, out_stride, out_dtype};
  const bool vec = (n % 4 == 0) && vec4_ok(sample, s_stride, s_dtype) && vec4_ok(m0, m0_stride, m0_dtype) &&
                   (!m1 || vec4_ok(m1, m1_stride, m1_dtype)) && vec4_ok(out, out_stride, out_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}
