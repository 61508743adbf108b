// du_step.cu — elementwise kernels of the path: F2 masks, F2c z-norm, F3 DDIM, F4/F5/F6 guided blends
// fused with F3, F7 perturbation, F8 slot copy, batch-axis sum.  All are pure streaming kernels: one
// thread owns 4 consecutive elements of one image (128-bit accesses), reads every input once and
// writes every output once; arithmetic repeats the reference's operations one IEEE rounding at a time
// (no FMA contraction) so fp32 results are bit-identical to the eager torch expressions.
#include <cooperative_groups.h>

#include "du_rows.cuh"

namespace cg = cooperative_groups;

namespace du {

// ---- F2a/F2b masks ----------------------------------------------------------------------------------
struct ThrMaskF {
  const void* u; int64_t u_stride; int u_dtype;
  const float* thr;        // per-row thresholds or null
  const void* thr_map; int thr_dtype;  // one row broadcast over b, or null
  int higher;
  float* mask; int64_t mask_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float x[VEC], t[VEC], m[VEC];
    loadv<VEC>(u, b * u_stride + i, u_dtype, x);
    if (thr) {
      float tb = __ldg(thr + b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) t[e] = tb;
    } else {
      loadv<VEC>(thr_map, i, thr_dtype, t);
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) m[e] = (higher ? (x[e] > t[e]) : (x[e] < t[e])) ? 1.0f : 0.0f;
    storev<VEC>(mask, b * mask_stride + i, DU_F32, m);
  }
};

// ---- F2c z-norm weights -------------------------------------------------------------------------------
__device__ __forceinline__ float zn_weight(float z, int mode, float thr) {
  if (mode == DU_ZN_BELOW) return z < thr ? 1.0f : 0.0f;
  if (mode == DU_ZN_ABOVE) return z > thr ? 1.0f : 0.0f;
  // multiscale: m2*0.8 + m1*0.9 + m0 with the reference's rounding order
  float m2 = (z < -2.0f && z > -3.0f) ? 1.0f : 0.0f;
  float m1 = (z < -1.0f && z > -2.0f) ? 1.0f : 0.0f;
  float m0 = (z >= -1.0f) ? 1.0f : 0.0f;
  return __fadd_rn(__fadd_rn(__fmul_rn(m2, 0.8f), __fmul_rn(m1, 0.9f)), m0);
}

struct ZnormF {
  const void* u; int64_t u_stride; int u_dtype;
  const float* stats; int normalize; int mode; float thr;
  float* z; int64_t z_stride;
  float* w; int64_t w_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float x[VEC], zz[VEC], ww[VEC];
    loadv<VEC>(u, b * u_stride + i, u_dtype, x);
    float mean = 0.0f, sd = 1.0f;
    if (normalize) { mean = __ldg(stats); sd = __ldg(stats + 1); }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      zz[e] = normalize ? __fdiv_rn(__fsub_rn(x[e], mean), sd) : x[e];
      ww[e] = zn_weight(zz[e], mode, thr);
    }
    if (z) storev<VEC>(z, b * z_stride + i, DU_F32, zz);
    if (w) storev<VEC>(w, b * w_stride + i, DU_F32, ww);
  }
};

// ---- F3 DDIM -----------------------------------------------------------------------------------------
struct DdimF {
  const void* mo; int64_t mo_stride; int mo_dtype;
  const void* x; int64_t x_stride; int x_dtype;
  const void* noise; int64_t noise_stride; int noise_dtype;
  du_ddim_coeffs c;
  void* prev; int64_t prev_stride; int prev_dtype;
  void* x0; int64_t x0_stride; int x0_dtype;
  void* eps; int64_t eps_stride; int eps_dtype;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float a[VEC], s[VEC], nz[VEC], pv[VEC], x0v[VEC], ev[VEC];
    loadv<VEC>(mo, b * mo_stride + i, mo_dtype, a);
    loadv<VEC>(x, b * x_stride + i, x_dtype, s);
    if (c.add_noise) loadv<VEC>(noise, b * noise_stride + i, noise_dtype, nz);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      DdimOut o = ddim_update(a[e], s[e], c.add_noise ? nz[e] : 0.0f, c);
      pv[e] = o.prev; x0v[e] = o.x0; ev[e] = o.eps;
    }
    if (prev) storev<VEC>(prev, b * prev_stride + i, prev_dtype, pv);
    if (x0) storev<VEC>(x0, b * x0_stride + i, x0_dtype, x0v);
    if (eps) storev<VEC>(eps, b * eps_stride + i, eps_dtype, ev);
  }
};

// ---- F2 + F4/F5/F6 + F3 ------------------------------------------------------------------------------
__device__ __forceinline__ float posterior_score(float u, float S, float post_M, float inv_alpha_hat) {
  // uncertainty_guidance.py:115-119, one rounding per op
  float inv_var = __fdiv_rn(1.0f, u);
  float trace = __fadd_rn(__fmul_rn(post_M, inv_var), inv_alpha_hat);
  float prec = __fdiv_rn(1.0f, trace);
  return __fmul_rn(prec, __fmul_rn(inv_var, S));
}

__device__ __forceinline__ float guided_eps(int guidance, float eps, float m, float u, float aux, float lam,
                                            float post_M, float inv_alpha_hat) {
  switch (guidance) {
    case DU_GUIDE_POSTERIOR: {
      float post = posterior_score(u, aux, post_M, inv_alpha_hat);
      return __fadd_rn(__fmul_rn(eps, __fsub_rn(1.0f, m)), __fmul_rn(m, post));
    }
    case DU_GUIDE_GRAD_BLEND: {
      float post = __fadd_rn(eps, __fmul_rn(lam, aux));
      return __fadd_rn(__fmul_rn(eps, __fsub_rn(1.0f, m)), __fmul_rn(post, m));
    }
    case DU_GUIDE_GRAD_ADD:
      return __fadd_rn(eps, __fmul_rn(__fmul_rn(lam, aux), m));
    case DU_GUIDE_WEIGHTS:
      return __fmul_rn(eps, m);
    case DU_GUIDE_LINCOMB:   // a*eps + lam*g (a travels in post_M): SU/scheduling_ddim_mc_dropout_gradient.py:514
      return __fadd_rn(__fmul_rn(post_M, eps), __fmul_rn(lam, aux));
    case DU_GUIDE_MUL_BLEND:   // eps (1-m) + eps m g: generate_samples.py:953 (legacy percentile loop)
      return __fadd_rn(__fmul_rn(eps, __fsub_rn(1.0f, m)), __fmul_rn(__fmul_rn(eps, m), aux));
    case DU_GUIDE_SIGN_ADD: {   // eps + u * sign(n) * m: PU/..._guided_second_order.py:249 (torch.sign = (0 < x) - (x < 0): NaN -> 0)
      const float sg = (aux > 0.0f) ? 1.0f : ((aux < 0.0f) ? -1.0f : 0.0f);
      return __fadd_rn(eps, __fmul_rn(__fmul_rn(u, sg), m));
    }
    default:
      return eps;
  }
}

struct GuidedF {
  du_guided_params p;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float e0[VEC], s[VEC], uu[VEC], m[VEC], ax[VEC], eg[VEC], pv[VEC], x0v[VEC];
    loadv<VEC>(p.eps, b * p.eps_stride + i, p.eps_dtype, e0);
    if (!p.skip_ddim) loadv<VEC>(p.sample, b * p.sample_stride + i, p.sample_dtype, s);
    const bool need_u = (p.u != nullptr);
    if (need_u) loadv<VEC>(p.u, b * p.u_stride + i, DU_F32, uu);
    if (p.mask) {
      // mask_period: a [B,1,H,W] mask broadcast over the channels of a [B,C,H,W] row (flip_threshold, SU/scheduling_ddim_flip_threshold.py:541)
      loadv<VEC>(p.mask, b * p.mask_stride + (p.mask_period > 0 ? i % p.mask_period : i), DU_F32, m);
    } else if (p.thr) {
      float tb = __ldg(p.thr + b);
#pragma unroll
      for (int e = 0; e < VEC; ++e) m[e] = (p.higher ? (uu[e] > tb) : (uu[e] < tb)) ? 1.0f : 0.0f;
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) m[e] = 1.0f;
    }
    if (p.aux) loadv<VEC>(p.aux, (p.aux_broadcast ? 0 : b * p.aux_stride) + i, p.aux_dtype, ax);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      eg[e] = guided_eps(p.guidance, e0[e], m[e], need_u ? uu[e] : 0.0f, p.aux ? ax[e] : e0[e], p.lam, p.post_M,
                         p.inv_alpha_hat);
      if (!p.skip_ddim) {
        if (p.guidance == DU_GUIDE_WEIGHTS || p.x0_unguided) {
          // F4: x0 from the UNMASKED model output, direction from the masked one, noise not re-added
          du_ddim_coeffs c = p.ddim;
          c.add_noise = 0;
          c.use_clipped_model_output = 0;
          DdimOut o = ddim_update(e0[e], s[e], 0.0f, c);
          float eps2 = eg[e];
          if (p.ddim.use_clipped_model_output)
            eps2 = __fdiv_rn(__fsub_rn(s[e], __fmul_rn(c.sqrt_alpha_t, o.x0)), c.sqrt_beta_t);
          eg[e] = eps2;
          x0v[e] = o.x0;
          pv[e] = __fadd_rn(__fmul_rn(c.sqrt_alpha_prev, o.x0), __fmul_rn(c.dir_coef, eps2));
        } else {
          du_ddim_coeffs c = p.ddim;
          c.add_noise = 0;  // guided re-steps are deterministic in every reference caller
          DdimOut o = ddim_update(eg[e], s[e], 0.0f, c);
          pv[e] = o.prev; x0v[e] = o.x0;
        }
      }
    }
    if (p.prev_out) storev<VEC>(p.prev_out, b * p.prev_stride + i, p.prev_dtype, pv);
    if (p.x0_out) storev<VEC>(p.x0_out, b * p.x0_stride + i, p.x0_dtype, x0v);
    if (p.eps_out) storev<VEC>(p.eps_out, b * p.eps_out_stride + i, p.eps_out_dtype, eg);
    if (p.mask_out) storev<VEC>(p.mask_out, b * p.mask_out_stride + i, DU_F32, m);
  }
};

// ---- F7 perturbation, F8 slot copy ---------------------------------------------------------------------
struct PerturbF {
  const void* x; int64_t x_stride; int x_dtype;
  const void* nz; int64_t nz_stride; int nz_dtype;   // null: out = a * x
  float a, b_;
  const float* a_rows; const float* b_rows;          // per-row scalars (du_perturb_rows), else null
  void* out; int64_t out_stride; int out_dtype;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float xv[VEC], nv[VEC], o[VEC];
    loadv<VEC>(x, b * x_stride + i, x_dtype, xv);
    const float aa = a_rows ? __ldg(a_rows + b) : a, bb = b_rows ? __ldg(b_rows + b) : b_;
    if (nz) {
      loadv<VEC>(nz, b * nz_stride + i, nz_dtype, nv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = __fadd_rn(__fmul_rn(aa, xv[e]), __fmul_rn(bb, nv[e]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) o[e] = __fmul_rn(aa, xv[e]);
    }
    storev<VEC>(out, b * out_stride + i, out_dtype, o);
  }
};

// ---- EMA of the map with bias correction and square root (PU/..._guided_second_order.py:212-218) ------------------------
struct EmaF {
  const float* mom; const void* u; int u_dtype; float beta, one_minus_beta, denom;
  float* mom_out; float* corrected; float* root;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float uv[VEC], mv[VEC], cv[VEC], rv[VEC];
    loadv<VEC>(u, i, u_dtype, uv);
    if (mom) {
      loadv<VEC>(mom, i, DU_F32, mv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) mv[e] = __fadd_rn(__fmul_rn(beta, mv[e]), __fmul_rn(one_minus_beta, uv[e]));
    } else {
#pragma unroll
      for (int e = 0; e < VEC; ++e) mv[e] = uv[e];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) { cv[e] = __fdiv_rn(mv[e], denom); rv[e] = __fsqrt_rn(cv[e]); }
    storev<VEC>(mom_out, i, DU_F32, mv);
    if (corrected) storev<VEC>(corrected, i, DU_F32, cv);
    if (root) storev<VEC>(root, i, DU_F32, rv);
    (void)b;
  }
};

struct CopyF {
  const void* src; int64_t src_stride; int src_dtype;
  void* dst; int64_t dst_stride; int dst_dtype;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float v[VEC];
    loadv<VEC>(src, b * src_stride + i, src_dtype, v);
    storev<VEC>(dst, b * dst_stride + i, dst_dtype, v);
  }
};

// ---- uint8 image epilogue: (x/2 + 0.5).clamp(0,1) * 255 -> round (half to even) -> uint8 ----------------------------
struct ImageU8F {
  const void* x; int64_t x_stride; int x_dtype;
  uint8_t* out; int64_t out_stride;
  template <int VEC> __device__ __forceinline__ void run(int64_t b, int64_t i) const {
    float v[VEC];
    loadv<VEC>(x, b * x_stride + i, x_dtype, v);
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float t = __fadd_rn(__fmul_rn(v[e], 0.5f), 0.5f);
      t = fminf(fmaxf(t, 0.0f), 1.0f);                 // NaN -> 0 here; torch's NaN -> uint8 cast is undefined anyway
      const uint32_t q = (uint32_t)__float2int_rn(__fmul_rn(t, 255.0f));
      if constexpr (VEC == 4) packed |= q << (8 * e);
      else out[b * out_stride + i] = (uint8_t)q;
    }
    if constexpr (VEC == 4) *reinterpret_cast<uint32_t*>(out + b * out_stride + i) = packed;
  }
};

// ---- batch-axis sum (fp64 accumulation, deterministic) --------------------------------------------------
__global__ void __launch_bounds__(256) batch_sum_kernel(const void* x, int64_t stride, int dt, int64_t B, int64_t n, float* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    int64_t b = 0;
    for (; b + 8 <= B; b += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = load1(x, (b + j) * stride + i, dt);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += (double)v[j];
    }
    for (; b < B; ++b) acc += (double)load1(x, b * stride + i, dt);
    out[i] = (float)acc;
  }
}

// Vector form: one CTA owns 32 x 16-byte column groups (512 contiguous bytes per image); its BS_SPLIT warps each sum a
// contiguous range of images (8 loads of 16 B in flight per thread), the fp64 partials meet in shared memory and are added
// in warp order, so the result does not depend on scheduling.
constexpr int BS_SPLIT = 8;

template <typename T>
__global__ void __launch_bounds__(32 * BS_SPLIT) batch_sum_rows_kernel(const T* __restrict__ x, int64_t stride, int64_t B, int64_t n,
                                                                       float* __restrict__ out, uint64_t pol) {
  using V = Vec16<T>;
  constexpr int VEC = V::VEC;
  __shared__ double part[BS_SPLIT][32][VEC];
  // a fused step launched as this kernel's programmatic dependent (du_fused_params::S_overlap) may start its S-independent
  // pilot now; it orders its first read of `out` with griddepcontrol.wait.  No effect for ordinary successors.
  asm volatile("griddepcontrol.launch_dependents;");
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int64_t per = (B + BS_SPLIT - 1) / BS_SPLIT;
  const int64_t b0 = r * per, b1 = (b0 + per < B) ? b0 + per : B;
  const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * VEC;
  double acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.0;
  if (i < n) {
    const T* col = x + i;
    int64_t b = b0;
    for (; b + 8 <= b1; b += 8) {
      uint4 raw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) raw[j] = ldg_stream_128_pol(col + (b + j) * stride, pol);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v[VEC];
        V::unpack(raw[j], v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += (double)v[e];
      }
    }
    for (; b < b1; ++b) {
      float v[VEC];
      V::unpack(ldg_stream_128_pol(col + b * stride, pol), v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += (double)v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) part[r][lane][e] = acc[e];
  __syncthreads();
  if (r == 0 && i < n) {
    float t[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < BS_SPLIT; ++k) a += part[k][lane][e];
      t[e] = (float)a;
    }
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) *reinterpret_cast<float4*>(out + i + 4 * h) = make_float4(t[4 * h], t[4 * h + 1], t[4 * h + 2], t[4 * h + 3]);
  }
}

// ---- z-norm statistics: fp64 (count, sum, sum of squares about a pivot) per block, fixed-order final merge -
struct ZStat { double s, ss; };

__global__ void __launch_bounds__(256) znorm_partial_kernel(const void* u, int64_t stride, int dt, int64_t B, int64_t n,
                                                            double* partials /*[grid][2]*/, float pivot_hint) {
  // pivot = first element of the tensor: keeps the squared sums small.  Rows are walked as rows (no division per element); the
  // per-thread sums run in fp32 over at most 64 values at a time before they are folded into the fp64 accumulators, so the
  // loop costs two fp32 instructions per element instead of two fp64 ones (the partition into chunks is fixed: deterministic).
  const float pivot = load1(u, 0, dt);
  double s = 0.0, ss = 0.0;
  const bool vec = (n % 4 == 0) && (stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(u) & (dt == DU_F32 ? 15 : 7)) == 0);
  const int64_t groups = vec ? n / 4 : n;
  // grid (x over the groups of a row, y over rows); four loads in flight per thread and trip
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; g0 < groups; g0 += (int64_t)gridDim.x * blockDim.x * 4) {
      float fs = 0.0f, fss = 0.0f;
      if (vec) {
        float v[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t g = g0 + (int64_t)j * blockDim.x;
          if (g < groups) load4(u, b * stride + 4 * g, dt, v[j]);
          else { v[j][0] = v[j][1] = v[j][2] = v[j][3] = pivot; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float d = v[j][e] - pivot; fs += d; fss = fmaf(d, d, fss); }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t g = g0 + (int64_t)j * blockDim.x;
          if (g < groups) { const float d = load1(u, b * stride + g, dt) - pivot; fs += d; fss = fmaf(d, d, fss); }
        }
      }
      s += (double)fs;
      ss += (double)fss;
    }
  }
  __shared__ double sh_s[8], sh_ss[8];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    ss += __shfl_down_sync(0xffffffffu, ss, o);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh_s[w] = s; sh_ss[w] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) { a += sh_s[j]; c += sh_ss[j]; }
    const int64_t blk = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    partials[2 * blk] = a;
    partials[2 * blk + 1] = c;
  }
  (void)pivot_hint;
}

__global__ void __launch_bounds__(256) znorm_final_kernel(const void* u, int dt, const double* partials, int nparts, int64_t total, float* stats) {
  // one CTA: thread i folds partials i, i + 256, ... and a fixed shuffle / shared-memory tree folds the threads — the order
  // depends on nothing but nparts, so the statistics are deterministic (a single thread walking ~1000 partials was latency-bound:
  // 30 us for a 25 MB map)
  double s = 0.0, ss = 0.0;
  for (int j = threadIdx.x; j < nparts; j += 256) { s += partials[2 * j]; ss += partials[2 * j + 1]; }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    ss += __shfl_down_sync(0xffffffffu, ss, o);
  }
  __shared__ double sh_s[8], sh_ss[8];
  if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_ss[threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  s = 0.0; ss = 0.0;
  for (int j = 0; j < 8; ++j) { s += sh_s[j]; ss += sh_ss[j]; }
  const double pivot = (double)load1(u, 0, dt);
  double cnt = (double)total;
  double mean_d = s / cnt;
  double m2 = ss - s * mean_d;
  if (m2 < 0.0) m2 = 0.0;
  stats[0] = (float)(pivot + mean_d);
  stats[1] = (float)sqrt(m2 / (cnt - 1.0));
  stats[2] = (float)cnt;
  stats[3] = (float)m2;
}

__global__ void znorm_combine_kernel(const float* in, int R, float* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  for (int r = 0; r < R; ++r) {
    double nr = in[4 * r + 2], mr = in[4 * r + 0], qr = in[4 * r + 3];
    if (nr == 0.0) continue;
    double tot = n + nr, delta = mr - mean;
    m2 = m2 + qr + delta * delta * (n * nr / tot);
    mean = mean + delta * (nr / tot);
    n = tot;
  }
  out[0] = (float)mean;
  out[1] = (float)sqrt(m2 / (n - 1.0));
  out[2] = (float)n;
  out[3] = (float)m2;
}

static bool check_view(const void* p, int dt) { return p && dtype_ok(dt) && aligned(p, dtype_size(dt)); }

}  // namespace du

using namespace du;

extern "C" int du_threshold_mask(const void* u, int64_t u_stride, int u_dtype, const float* thr, int higher,
                                 int64_t B, int64_t n, float* mask_out, int64_t mask_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(u, u_dtype) || !thr || !mask_out || B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_threshold_mask: bad arguments");
  ThrMaskF f{u, u_stride, u_dtype, thr, nullptr, DU_F32, higher, mask_out, mask_stride};
  bool vec = (n % 4 == 0) && vec4_ok(u, u_stride, u_dtype) && vec4_ok(mask_out, mask_stride, DU_F32);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_tensor_threshold_mask(const void* u, int64_t u_stride, int u_dtype, const void* thr_map, int thr_dtype,
                                        int higher, int64_t B, int64_t n, float* mask_out, int64_t mask_stride,
                                        du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(u, u_dtype) || !check_view(thr_map, thr_dtype) || !mask_out || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_tensor_threshold_mask: bad arguments");
  ThrMaskF f{u, u_stride, u_dtype, nullptr, thr_map, thr_dtype, higher, mask_out, mask_stride};
  bool vec = (n % 4 == 0) && vec4_ok(u, u_stride, u_dtype) && vec4_ok(thr_map, 0, thr_dtype) && vec4_ok(mask_out, mask_stride, DU_F32);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

// grid of the partial-sum kernel: x over a row's groups (1024 per CTA and trip), y over rows, at most ~8 CTAs per SM in all
static dim3 znorm_grid(int64_t B, int64_t n) {
  int64_t gy = B < 1184 ? (B < 1 ? 1 : B) : 1184;
  int64_t gx = (n / 4 + 1023) / 1024;
  if (gx < 1) gx = 1;
  const int64_t cap = (1184 + gy - 1) / gy;
  if (gx > cap) gx = cap;
  return dim3((unsigned)gx, (unsigned)gy);
}
static int znorm_blocks(int64_t B, int64_t n) { const dim3 g = znorm_grid(B, n); return (int)(g.x * g.y); }

extern "C" size_t du_znorm_scratch_bytes(int64_t B, int64_t n) { return (size_t)znorm_blocks(B, n) * 2 * sizeof(double); }

extern "C" int du_znorm_stats(const void* u, int64_t u_stride, int u_dtype, int64_t B, int64_t n, float* stats_out,
                              void* scratch, size_t scratch_bytes, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(u, u_dtype) || !stats_out || !scratch || B <= 0 || n <= 0) return set_error(DU_ERR_BAD_ARG, "du_znorm_stats: bad arguments");
  if (scratch_bytes < du_znorm_scratch_bytes(B, n) || !aligned(scratch, 8)) return set_error(DU_ERR_SCRATCH, "du_znorm_stats: scratch too small or misaligned");
  const int blocks = znorm_blocks(B, n);
  cudaStream_t st = (cudaStream_t)stream;
  znorm_partial_kernel<<<znorm_grid(B, n), 256, 0, st>>>(u, u_stride, u_dtype, B, n, (double*)scratch, 0.0f);
  DU_LAUNCH_CHECK("znorm_partial_kernel");
  znorm_final_kernel<<<1, 256, 0, st>>>(u, u_dtype, (const double*)scratch, blocks, B * n, stats_out);
  DU_LAUNCH_CHECK("znorm_final_kernel");
  return DU_OK;
}

extern "C" int du_znorm_stats_combine(const float* stats_in, int R, float* stats_out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!stats_in || !stats_out || R < 1) return set_error(DU_ERR_BAD_ARG, "du_znorm_stats_combine: bad arguments");
  znorm_combine_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(stats_in, R, stats_out);
  DU_LAUNCH_CHECK("znorm_combine_kernel");
  return DU_OK;
}

extern "C" int du_znorm_weights(const void* u, int64_t u_stride, int u_dtype, const float* stats, int normalize, int mode,
                                float thr, int64_t B, int64_t n, float* z_out, int64_t z_stride, float* w_out,
                                int64_t w_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(u, u_dtype) || (normalize && !stats) || (!z_out && !w_out) || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_znorm_weights: bad arguments");
  if (mode < DU_ZN_BELOW || mode > DU_ZN_MULTISCALE) return set_error(DU_ERR_BAD_ARG, "du_znorm_weights: bad mode %d", mode);
  ZnormF f{u, u_stride, u_dtype, stats, normalize, mode, thr, z_out, z_stride, w_out, w_stride};
  bool vec = (n % 4 == 0) && vec4_ok(u, u_stride, u_dtype) && vec4_ok(z_out, z_stride, DU_F32) && vec4_ok(w_out, w_stride, DU_F32);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

static int check_coeffs(const du_ddim_coeffs* c, const char* who) {
  if (!c) return set_error(DU_ERR_BAD_ARG, "%s: null coefficients", who);
  if (c->prediction_type < DU_PRED_EPSILON || c->prediction_type > DU_PRED_V)
    return set_error(DU_ERR_BAD_ARG, "%s: prediction_type %d must be one of epsilon(0), sample(1), v_prediction(2)", who, c->prediction_type);
  return DU_OK;
}

extern "C" int du_ddim_step(const void* model_output, int64_t mo_stride, int mo_dtype, const void* sample, int64_t s_stride,
                            int s_dtype, const void* noise, int64_t noise_stride, int noise_dtype, const du_ddim_coeffs* c,
                            int64_t B, int64_t n, void* prev_out, int64_t prev_stride, int prev_dtype, void* x0_out,
                            int64_t x0_stride, int x0_dtype, void* eps_out, int64_t eps_stride, int eps_dtype,
                            du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  int rc = check_coeffs(c, "du_ddim_step");
  if (rc) return rc;
  if (!check_view(model_output, mo_dtype) || !check_view(sample, s_dtype) || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_ddim_step: bad inputs");
  if (!prev_out && !x0_out && !eps_out) return set_error(DU_ERR_BAD_ARG, "du_ddim_step: no output requested");
  if (c->add_noise && !check_view(noise, noise_dtype)) return set_error(DU_ERR_BAD_ARG, "du_ddim_step: eta > 0 needs a noise tensor");
  if ((prev_out && !dtype_ok(prev_dtype)) || (x0_out && !dtype_ok(x0_dtype)) || (eps_out && !dtype_ok(eps_dtype)))
    return set_error(DU_ERR_DTYPE, "du_ddim_step: unsupported output dtype");
  DdimF f{model_output, mo_stride, mo_dtype, sample, s_stride, s_dtype, c->add_noise ? noise : nullptr, noise_stride, noise_dtype,
          *c, prev_out, prev_stride, prev_dtype, x0_out, x0_stride, x0_dtype, eps_out, eps_stride, eps_dtype};
  bool vec = (n % 4 == 0) && vec4_ok(model_output, mo_stride, mo_dtype) && vec4_ok(sample, s_stride, s_dtype) &&
             vec4_ok(f.noise, noise_stride, noise_dtype) && vec4_ok(prev_out, prev_stride, prev_dtype) &&
             vec4_ok(x0_out, x0_stride, x0_dtype) && vec4_ok(eps_out, eps_stride, eps_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_guided_step(const du_guided_params* p, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!p) return set_error(DU_ERR_BAD_ARG, "du_guided_step: null params");
  if (p->guidance < DU_GUIDE_NONE || p->guidance > DU_GUIDE_MUL_BLEND) return set_error(DU_ERR_BAD_ARG, "du_guided_step: bad guidance %d", p->guidance);
  if (!check_view(p->eps, p->eps_dtype) || p->B < 0 || p->n < 0) return set_error(DU_ERR_BAD_ARG, "du_guided_step: bad eps view");
  if (!p->skip_ddim) {
    int rc = check_coeffs(&p->ddim, "du_guided_step");
    if (rc) return rc;
    if (!check_view(p->sample, p->sample_dtype)) return set_error(DU_ERR_BAD_ARG, "du_guided_step: bad sample view");
    if (p->guidance == DU_GUIDE_WEIGHTS && p->ddim.prediction_type != DU_PRED_EPSILON)
      return set_error(DU_ERR_BAD_ARG, "du_guided_step: masked re-step is implemented only for prediction type epsilon");
  }
  if (p->guidance != DU_GUIDE_NONE && p->guidance != DU_GUIDE_GRAD_ADD && p->guidance != DU_GUIDE_LINCOMB && !p->mask && !p->thr)
    return set_error(DU_ERR_BAD_ARG, "du_guided_step: guidance needs thr or mask");
  if (p->mask_period < 0 || (p->mask_period > 0 && (!p->mask || p->n % p->mask_period != 0)))
    return set_error(DU_ERR_BAD_ARG, "du_guided_step: mask_period must divide the row length");
  if ((p->thr || p->guidance == DU_GUIDE_POSTERIOR || p->guidance == DU_GUIDE_SIGN_ADD) && !p->u) return set_error(DU_ERR_BAD_ARG, "du_guided_step: u required");
  if ((p->guidance == DU_GUIDE_GRAD_BLEND || p->guidance == DU_GUIDE_GRAD_ADD || p->guidance == DU_GUIDE_LINCOMB || p->guidance == DU_GUIDE_SIGN_ADD ||
       p->guidance == DU_GUIDE_MUL_BLEND) && !check_view(p->aux, p->aux_dtype))
    return set_error(DU_ERR_BAD_ARG, "du_guided_step: gradient tensor required");
  if (p->aux && !dtype_ok(p->aux_dtype)) return set_error(DU_ERR_DTYPE, "du_guided_step: bad aux dtype");
  if (!p->prev_out && !p->x0_out && !p->eps_out && !p->mask_out) return set_error(DU_ERR_BAD_ARG, "du_guided_step: no output requested");
  if (p->skip_ddim && (p->prev_out || p->x0_out)) return set_error(DU_ERR_BAD_ARG, "du_guided_step: skip_ddim with prev/x0 outputs");
  GuidedF f{*p};
  int64_t n = p->n;
  bool vec = (n % 4 == 0) && (p->mask_period % 4 == 0) && vec4_ok(p->eps, p->eps_stride, p->eps_dtype) &&
             (p->skip_ddim || vec4_ok(p->sample, p->sample_stride, p->sample_dtype)) && vec4_ok(p->u, p->u_stride, DU_F32) &&
             vec4_ok(p->mask, p->mask_stride, DU_F32) && vec4_ok(p->aux, p->aux_broadcast ? 0 : p->aux_stride, p->aux_dtype) &&
             vec4_ok(p->prev_out, p->prev_stride, p->prev_dtype) && vec4_ok(p->x0_out, p->x0_stride, p->x0_dtype) &&
             vec4_ok(p->eps_out, p->eps_out_stride, p->eps_out_dtype) && vec4_ok(p->mask_out, p->mask_out_stride, DU_F32);
  return launch_rows(p->B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_batch_sum(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, float* out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(x, x_dtype) || !out || B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_batch_sum: bad arguments");
  if (n == 0) return DU_OK;
  const int vec = (x_dtype == DU_F32) ? 4 : 8;
  if (B >= 2 * BS_SPLIT && n % vec == 0 && aligned(x, 16) && x_stride % vec == 0 && aligned(out, 16)) {
    const unsigned grid = (unsigned)((n / vec + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    // x (the step's eps) is read again by the kernel that consumes the sum: ask L2 to keep it (DU_L2_HINTS=0: normal priority)
    const char* e_h = getenv("DU_L2_HINTS");
    const uint64_t pol = (e_h && atoi(e_h) == 0) ? kL2EvictNormal : kL2EvictLast;
    if (x_dtype == DU_F32) batch_sum_rows_kernel<float><<<grid, 32 * BS_SPLIT, 0, st>>>((const float*)x, x_stride, B, n, out, pol);
    else if (x_dtype == DU_F16) batch_sum_rows_kernel<__half><<<grid, 32 * BS_SPLIT, 0, st>>>((const __half*)x, x_stride, B, n, out, pol);
    else batch_sum_rows_kernel<__nv_bfloat16><<<grid, 32 * BS_SPLIT, 0, st>>>((const __nv_bfloat16*)x, x_stride, B, n, out, pol);
    DU_LAUNCH_CHECK("batch_sum_rows_kernel");
    return DU_OK;
  }
  int64_t blocks = (n + 255) / 256;
  batch_sum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, x_stride, x_dtype, B, n, out);
  DU_LAUNCH_CHECK("batch_sum_kernel");
  return DU_OK;
}

extern "C" int du_perturb(const void* x, int64_t x_stride, int x_dtype, const void* noise, int64_t noise_stride, int noise_dtype,
                          float a, float b, int64_t B, int64_t n, void* out, int64_t out_stride, int out_dtype,
                          du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(x, x_dtype) || (noise && !check_view(noise, noise_dtype)) || !check_view(out, out_dtype) || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_perturb: bad arguments");
  PerturbF f{x, x_stride, x_dtype, noise, noise_stride, noise_dtype, a, b, nullptr, nullptr, out, out_stride, out_dtype};
  bool vec = (n % 4 == 0) && vec4_ok(x, x_stride, x_dtype) && vec4_ok(noise, noise_stride, noise_dtype) && vec4_ok(out, out_stride, out_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_perturb_rows(const void* x, int64_t x_stride, int x_dtype, const void* noise, int64_t noise_stride, int noise_dtype,
                               const float* a_rows, const float* b_rows, int64_t B, int64_t n, void* out, int64_t out_stride,
                               int out_dtype, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(x, x_dtype) || !check_view(noise, noise_dtype) || !check_view(out, out_dtype) || !a_rows || !b_rows || B < 0 || n < 0)
    return set_error(DU_ERR_BAD_ARG, "du_perturb_rows: bad arguments");
  PerturbF f{x, x_stride, x_dtype, noise, noise_stride, noise_dtype, 0.0f, 0.0f, a_rows, b_rows, out, out_stride, out_dtype};
  bool vec = (n % 4 == 0) && vec4_ok(x, x_stride, x_dtype) && vec4_ok(noise, noise_stride, noise_dtype) && vec4_ok(out, out_stride, out_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_ema_update(const float* momentum, const void* u, int u_dtype, float beta, float one_minus_beta, float denom, int64_t N,
                             float* momentum_out, float* corrected_out, float* sqrt_out, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(u, u_dtype) || !momentum_out || N < 0) return set_error(DU_ERR_BAD_ARG, "du_ema_update: bad arguments");
  EmaF f{momentum, u, u_dtype, beta, one_minus_beta, denom, momentum_out, corrected_out, sqrt_out};
  bool vec = (N % 4 == 0) && vec4_ok(u, 0, u_dtype) && vec4_ok(momentum, 0, DU_F32) && vec4_ok(momentum_out, 0, DU_F32) &&
             vec4_ok(corrected_out, 0, DU_F32) && vec4_ok(sqrt_out, 0, DU_F32);
  return launch_rows(1, N, vec, f, (cudaStream_t)stream);
}

extern "C" int du_image_uint8(const void* x, int64_t x_stride, int x_dtype, int64_t B, int64_t n, uint8_t* out,
                              int64_t out_stride, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(x, x_dtype) || !out || B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_image_uint8: bad arguments");
  ImageU8F f{x, x_stride, x_dtype, out, out_stride};
  bool vec = (n % 4 == 0) && vec4_ok(x, x_stride, x_dtype) && aligned(out, 4) && (out_stride % 4 == 0);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}

extern "C" int du_accumulate_slot(const void* src, int64_t src_stride, int src_dtype, int64_t B, int64_t n, void* dst,
                                  int64_t dst_stride, int dst_dtype, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!check_view(src, src_dtype) || !check_view(dst, dst_dtype) || B < 0 || n < 0) return set_error(DU_ERR_BAD_ARG, "du_accumulate_slot: bad arguments");
  CopyF f{src, src_stride, src_dtype, dst, dst_stride, dst_dtype};
  bool vec = (n % 4 == 0) && vec4_ok(src, src_stride, src_dtype) && vec4_ok(dst, dst_stride, dst_dtype);
  return launch_rows(B, n, vec, f, (cudaStream_t)stream);
}
