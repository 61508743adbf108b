// du_fused.cu — the fused uncertainty step: F1 (moments over M) -> F2a (per-image quantile + mask) -> F5
// (posterior score) -> F3 (DDIM x_{t-1}) (+F8: the map goes straight to its accumulation slot), ONE launch.
//
// One thread-block CLUSTER per image.  Each CTA of the cluster owns a contiguous slice of the image:
//   phase A  streams the M score tensors + eps of its slice through registers (128-bit L1-bypassing loads,
//            every HBM byte read once), writes the map to global memory AND keeps it (and eps) in shared
//            memory, and histograms the top key byte on the fly;
//   phase B  MSB-first radix select (4 x 8 bit) of the lo-th / hi-th order statistic over the
//            DISTRIBUTED shared-memory copy of the map: per-CTA histograms are merged into CTA 0's shared
//            memory with DSMEM atomics, two cluster barriers per level, no global-memory traffic at all;
//   phase C  threshold = torch's lerp of the two statistics; mask, posterior blend and DDIM update on the
//            slice, reading u/eps from shared memory and only `sample` from HBM; x_{t-1} written once.
// HBM traffic = (M+1)*s_in + 4 (map) + s_x (sample) + s_x (x_{t-1}) bytes per element: the algorithmic minimum.
#include <cooperative_groups.h>

#include "du_common.cuh"

namespace cg = cooperative_groups;

namespace du {

constexpr int FUSED_LEVELS = 4;  // 4 x 8-bit digits

struct FusedKParams {
  du_fused_params p;
  int64_t L;        // elements per CTA slice (n / cluster size)
  uint32_t lo, hi;  // ranks
  float w;          // lerp weight
};

// warp 0 locates rank k in a 256-bin histogram (generic pointer: may be DSMEM); result -> res[0..2]
__device__ __forceinline__ void locate_bin_256(const uint32_t* hist, uint32_t k, uint32_t* res) {
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
      if (k < cum + c[j]) { res[0] = 8 * lane + j; res[1] = cum; res[2] = c[j]; break; }
      cum += c[j];
    }
  }
}

template <typename T, bool KEEP_EPS, int THREADS>
__global__ void __launch_bounds__(THREADS, 2) fused_step_kernel(const __grid_constant__ FusedKParams kp) {
  using FV = Vec16<T>;
  constexpr int VEC = FV::VEC;
  const du_fused_params& p = kp.p;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned crank = cluster.block_rank();
  const unsigned csize = cluster.num_blocks();
  const int64_t b = blockIdx.y;
  const int64_t L = kp.L;
  const int64_t base = (int64_t)crank * L;  // slice start within the image
  const int tid = threadIdx.x, T_ = blockDim.x;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* u_s = reinterpret_cast<float*>(smem_raw);
  float* eps_s = u_s + L;
  uint32_t* hist = reinterpret_cast<uint32_t*>(KEEP_EPS ? (eps_s + L) : eps_s);  // [256] local
  uint32_t* merged = hist + 256;                                                 // [FUSED_LEVELS][256], used on CTA 0
  uint32_t* misc = merged + FUSED_LEVELS * 256;                                  // [16]
  // misc: 0..2 locate result, 3 nan flag (local), 4 min key (CTA0: cluster-wide), 5 nan flag (CTA0: cluster-wide)

  for (int j = tid; j < 256 * (1 + FUSED_LEVELS); j += T_) hist[j] = 0;
  if (tid < 16) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;
  __syncthreads();
  if (csize > 1) cluster.sync();  // every CTA's shared memory is initialised before any DSMEM access

  uint32_t* merged0 = (csize > 1) ? cluster.map_shared_rank(merged, 0) : merged;
  uint32_t* misc0 = (csize > 1) ? cluster.map_shared_rank(misc, 0) : misc;

  // ---------------------------------------------------------------- phase A: moments + level-0 histogram
  const int mode = p.moments_mode;
  const bool centered = (mode == DU_MOM_CENTERED);
  const bool extra = (mode == DU_MOM_VAR_WITH_CENTER);
  const int count = p.M + (extra ? 1 : 0);
  const float cnt = (float)count;
  uint32_t nan_seen = 0;
  const int centre_mode = centered ? 1 : (extra ? 2 : 0);
  for (int64_t g = tid; g < L / VEC; g += T_) {
    const int64_t i = base + g * VEC;
    float c[VEC], k[VEC], s1[VEC], s2[VEC];
    uint4 raw_c = make_uint4(0u, 0u, 0u, 0u);
    if (KEEP_EPS || centre_mode) raw_c = FV::load(p.eps, b * p.eps_stride + i);
    accumulate_scores<T>(p.scores, p.M, b * p.score_stride + i, raw_c, centre_mode, false, !centered, c, k, s1, s2);
    if (KEEP_EPS && centre_mode == 0) FV::unpack(raw_c, c);
    float u[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (centered) {
        u[e] = s2[e] / cnt;
      } else {
        float m2 = s2[e] - s1[e] * s1[e] / cnt;
        m2 = (m2 < 0.0f) ? 0.0f : m2;
        u[e] = m2 / (float)(count - 1);
      }
      nan_seen |= (u[e] != u[e]);
      atomicAdd(&hist[float_to_key(u[e]) >> 24], 1u);
    }
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
      float4 u4 = make_float4(u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]);
      *reinterpret_cast<float4*>(u_s + g * VEC + 4 * h) = u4;
      *reinterpret_cast<float4*>(p.unc_out + b * p.unc_stride + i + 4 * h) = u4;
      if (KEEP_EPS) *reinterpret_cast<float4*>(eps_s + g * VEC + 4 * h) = make_float4(c[4 * h], c[4 * h + 1], c[4 * h + 2], c[4 * h + 3]);
    }
  }
  if (__any_sync(0xffffffffu, nan_seen) && (tid & 31) == 0) atomicOr(&misc0[5], 1u);

  // ---------------------------------------------------------------- phase B: radix select over DSMEM
  uint32_t prefix = 0, pmask = 0, k_rank = kp.lo, below = 0, bincount = 0;
  for (int level = 0; level < FUSED_LEVELS; ++level) {
    const int shift = 24 - 8 * level;
    if (level > 0) {
      for (int64_t g = tid; g < L / 4; g += T_) {
        float4 v = *reinterpret_cast<const float4*>(u_s + 4 * g);
        uint32_t k0 = float_to_key(v.x), k1 = float_to_key(v.y), k2 = float_to_key(v.z), k3 = float_to_key(v.w);
        if ((k0 & pmask) == prefix) atomicAdd(&hist[(k0 >> shift) & 255u], 1u);
        if ((k1 & pmask) == prefix) atomicAdd(&hist[(k1 >> shift) & 255u], 1u);
        if ((k2 & pmask) == prefix) atomicAdd(&hist[(k2 >> shift) & 255u], 1u);
        if ((k3 & pmask) == prefix) atomicAdd(&hist[(k3 >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    const uint32_t* src = hist;
    if (csize > 1) {
      if (tid < 256) {
        uint32_t v = hist[tid];
        if (v) atomicAdd(&merged0[level * 256 + tid], v);
      }
      cluster.sync();
      src = merged0 + level * 256;
    }
    if (tid < 32) locate_bin_256(src, k_rank, misc);
    __syncthreads();
    prefix |= misc[0] << shift;
    pmask |= 255u << shift;
    below += misc[1];
    k_rank -= misc[1];
    bincount = misc[2];
    __syncthreads();
    if (level + 1 < FUSED_LEVELS) {
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
    }
  }
  const uint32_t key_lo = prefix;
  uint32_t key_hi = key_lo;
  if (kp.hi >= below + bincount) {  // cluster-uniform: the hi-th statistic is the smallest key above key_lo
    uint32_t best = 0xffffffffu;
    for (int64_t g = tid; g < L / 4; g += T_) {
      float4 v = *reinterpret_cast<const float4*>(u_s + 4 * g);
      uint32_t k0 = float_to_key(v.x), k1 = float_to_key(v.y), k2 = float_to_key(v.z), k3 = float_to_key(v.w);
      if (k0 > key_lo && k0 < best) best = k0;
      if (k1 > key_lo && k1 < best) best = k1;
      if (k2 > key_lo && k2 < best) best = k2;
      if (k3 > key_lo && k3 < best) best = k3;
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&misc0[4], best);
  }
  if (csize > 1) cluster.sync(); else __syncthreads();
  if (kp.hi >= below + bincount) key_hi = misc0[4];
  const bool has_nan = misc0[5] != 0;
  float thr = lerp_torch(key_to_float(key_lo), key_to_float(key_hi), kp.w, p.lerp_fma);
  if (has_nan) thr = __int_as_float(0x7fc00000);
  if (crank == 0 && tid == 0 && p.thr_out) p.thr_out[b] = thr;

  // ---------------------------------------------------------------- phase C: mask + posterior + DDIM
  du_ddim_coeffs dc = p.ddim;
  dc.add_noise = 0;
  const bool higher = p.higher != 0;
  constexpr int UNR = 2;  // groups per thread per trip: all global loads of a trip are issued before the arithmetic
  for (int64_t g0 = tid; g0 < L / 4; g0 += (int64_t)UNR * T_) {
    float s[UNR][4], S[UNR][4], e0[UNR][4];
#pragma unroll
    for (int r = 0; r < UNR; ++r) {
      const int64_t g = g0 + (int64_t)r * T_;
      if (g < L / 4) {
        const int64_t i = base + 4 * g;
        load4(p.sample, b * p.sample_stride + i, p.sample_dtype, s[r]);
        if (!KEEP_EPS) load4(p.eps, b * p.eps_stride + i, p.score_dtype, e0[r]);
        if (p.S) {
          float4 s4 = __ldg(reinterpret_cast<const float4*>(p.S + (p.S_broadcast ? 0 : b * p.S_stride) + i));
          S[r][0] = s4.x; S[r][1] = s4.y; S[r][2] = s4.z; S[r][3] = s4.w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < UNR; ++r) {
      const int64_t g = g0 + (int64_t)r * T_;
      if (g >= L / 4) break;
      const int64_t i = base + 4 * g;
      float4 u4 = *reinterpret_cast<const float4*>(u_s + 4 * g);
      float uu[4] = {u4.x, u4.y, u4.z, u4.w};
      if (KEEP_EPS) {
        float4 e4 = *reinterpret_cast<const float4*>(eps_s + 4 * g);
        e0[r][0] = e4.x; e0[r][1] = e4.y; e0[r][2] = e4.z; e0[r][3] = e4.w;
      }
      if (!p.S) {
#pragma unroll
        for (int e = 0; e < 4; ++e) S[r][e] = e0[r][e];
      }
      float pv[4], x0v[4], eg[4], mk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mk[e] = (higher ? (uu[e] > thr) : (uu[e] < thr)) ? 1.0f : 0.0f;
        float inv_var = __fdiv_rn(1.0f, uu[e]);
        float trace = __fadd_rn(__fmul_rn(p.post_M, inv_var), p.inv_alpha_hat);
        float prec = __fdiv_rn(1.0f, trace);
        float post = __fmul_rn(prec, __fmul_rn(inv_var, S[r][e]));
        eg[e] = __fadd_rn(__fmul_rn(e0[r][e], __fsub_rn(1.0f, mk[e])), __fmul_rn(mk[e], post));
        DdimOut o = ddim_update(eg[e], s[r][e], 0.0f, dc);
        pv[e] = o.prev; x0v[e] = o.x0;
      }
      store4(p.prev_out, b * p.prev_stride + i, p.prev_dtype, pv);
      if (p.x0_out) store4(p.x0_out, b * p.x0_stride + i, p.prev_dtype, x0v);
      if (p.eps_out) store4(p.eps_out, b * p.eps_out_stride + i, DU_F32, eg);
      if (p.mask_out) store4(p.mask_out, b * p.mask_out_stride + i, DU_F32, mk);
    }
  }
  if (csize > 1) cluster.sync();  // keep CTA 0's shared memory alive until every peer has read it
}

static size_t fused_smem_bytes(int64_t L, bool keep_eps) {
  return (size_t)L * 4 * (keep_eps ? 2 : 1) + (size_t)(256 * (1 + FUSED_LEVELS) + 16) * 4;
}

struct FusedPlan { int cluster; bool keep_eps; int threads; size_t smem; };

static bool fused_plan(int64_t n, int vec, FusedPlan* out) {
  // override for tuning: DU_FUSED_CLUSTER=<1|2|4|8>, DU_FUSED_KEEP_EPS=<0|1>, DU_FUSED_THREADS=<n>
  const char* e_c = getenv("DU_FUSED_CLUSTER");
  const char* e_k = getenv("DU_FUSED_KEEP_EPS");
  const char* e_t = getenv("DU_FUSED_THREADS");
  const size_t kMax = 227 * 1024, kHalf = 113 * 1024;
  int best_c = 0; bool best_keep = false;
  for (int pass = 0; pass < 2 && !best_c; ++pass) {      // pass 0: two CTAs per SM; pass 1: anything that fits
    for (int keep = 1; keep >= 0 && !best_c; --keep) {
      for (int c = 1; c <= 8; c <<= 1) {
        if (e_c && atoi(e_c) != c) continue;
        if (e_k && atoi(e_k) != keep) continue;
        if (n % ((int64_t)c * vec) != 0) continue;
        size_t s = fused_smem_bytes(n / c, keep != 0);
        if (s <= (pass == 0 ? kHalf : kMax)) { best_c = c; best_keep = keep != 0; break; }
      }
    }
  }
  if (!best_c) return false;
  int64_t L = n / best_c;
  int64_t groups = L / vec;
  int threads = (groups >= 1024) ? 512 : 256;  // the histogram merge uses threads 0..255
  if (e_t) threads = (atoi(e_t) == 512) ? 512 : 256;
  out->cluster = best_c; out->keep_eps = best_keep; out->threads = threads; out->smem = fused_smem_bytes(L, best_keep);
  return true;
}

template <typename T, bool KEEP, int THREADS>
static int launch_fused_t(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  auto kern = fused_step_kernel<T, KEEP, THREADS>;
  DU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)plan.cluster, (unsigned)kp.p.B, 1);
  cfg.blockDim = dim3((unsigned)plan.threads);
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DU_CUDA(cudaLaunchKernelEx(&cfg, kern, kp));
  return DU_OK;
}

template <typename T, bool KEEP>
static int launch_fused(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  if (plan.threads == 512) return launch_fused_t<T, KEEP, 512>(kp, plan, st);
  return launch_fused_t<T, KEEP, 256>(kp, plan, st);
}

}  // namespace du

using namespace du;

extern "C" int du_fused_supported(int64_t n, int score_dtype) {
  if (!dtype_ok(score_dtype) || n <= 0 || n > ((int64_t)1 << 24)) return 0;
  FusedPlan plan;
  return fused_plan(n, score_dtype == DU_F32 ? 4 : 8, &plan) ? plan.cluster : 0;
}

extern "C" int du_fused_uncertainty_step(const du_fused_params* p, du_stream_t stream) {
  if (!p) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: null params");
  if (p->M < 1 || p->M > DU_MAX_M) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: M=%d must be in [1,%d]", p->M, DU_MAX_M);
  if (!dtype_ok(p->score_dtype) || !dtype_ok(p->sample_dtype) || !dtype_ok(p->prev_dtype))
    return set_error(DU_ERR_DTYPE, "du_fused_uncertainty_step: unsupported dtype");
  if (p->moments_mode != DU_MOM_VAR_UNBIASED && p->moments_mode != DU_MOM_CENTERED && p->moments_mode != DU_MOM_VAR_WITH_CENTER)
    return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: moments mode %d not supported", p->moments_mode);
  if (p->B < 0 || p->n <= 0) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: bad sizes");
  if (!(p->q >= 0.0f && p->q <= 1.0f)) return set_error(DU_ERR_BAD_ARG, "quantile() q values must be in the range [0, 1]");
  if (p->n > ((int64_t)1 << 24)) return set_error(DU_ERR_TOO_LARGE, "quantile() input tensor is too large");
  if (p->B == 0) return DU_OK;
  if (p->B > 65535) return set_error(DU_ERR_TOO_LARGE, "du_fused_uncertainty_step: batch > 65535, split the call");
  if (!p->eps || !p->sample || !p->unc_out || !p->prev_out) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: null tensor");
  const int vec = p->score_dtype == DU_F32 ? 4 : 8;
  FusedPlan plan;
  if (!fused_plan(p->n, vec, &plan))
    return set_error(DU_ERR_TOO_LARGE, "du_fused_uncertainty_step: rows of %lld elements do not fit cluster shared memory; use the unfused calls", (long long)p->n);
  // 128-bit access requirements
  bool ok = aligned(p->eps, 16) && (p->eps_stride % vec == 0) && (p->score_stride % vec == 0) &&
            vec4_ok(p->sample, p->sample_stride, p->sample_dtype) && aligned(p->unc_out, 16) && (p->unc_stride % 4 == 0) &&
            vec4_ok(p->prev_out, p->prev_stride, p->prev_dtype) && vec4_ok(p->x0_out, p->x0_stride, p->prev_dtype) &&
            vec4_ok(p->eps_out, p->eps_out_stride, DU_F32) && vec4_ok(p->mask_out, p->mask_out_stride, DU_F32) &&
            (!p->S || (aligned(p->S, 16) && (p->S_broadcast || p->S_stride % 4 == 0)));
  for (int m = 0; m < p->M; ++m) {
    if (!p->scores[m]) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: scores[%d] is null", m);
    ok = ok && aligned(p->scores[m], 16);
  }
  if (!ok) return set_error(DU_ERR_ALIGN, "du_fused_uncertainty_step: tensors must be 16-byte aligned with vector-multiple strides; use the unfused calls");
  FusedKParams kp;
  kp.p = *p;
  kp.L = p->n / plan.cluster;
  volatile float rank = p->q * (float)(p->n - 1);
  float fl = floorf(rank);
  kp.lo = (uint32_t)fl;
  kp.hi = (uint32_t)ceilf(rank);
  kp.w = rank - fl;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->score_dtype) {
    case DU_F32: return plan.keep_eps ? launch_fused<float, true>(kp, plan, st) : launch_fused<float, false>(kp, plan, st);
    case DU_F16: return plan.keep_eps ? launch_fused<__half, true>(kp, plan, st) : launch_fused<__half, false>(kp, plan, st);
    default: return plan.keep_eps ? launch_fused<__nv_bfloat16, true>(kp, plan, st) : launch_fused<__nv_bfloat16, false>(kp, plan, st);
  }
}
