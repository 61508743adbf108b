// du_fused.cu — the fused uncertainty step: F1 (moments over M) -> F2a (per-image quantile + mask) -> F5
// (posterior score) -> F3 (DDIM x_{t-1}) (+F8: the map goes straight to its accumulation slot), ONE launch.
//
// One thread-block CLUSTER per image, every cluster of the batch resident at once (one wave).  Each CTA owns a
// contiguous slice of its image:
//   phase A  streams the M score tensors (+ eps) of the slice through registers (128-bit L1-bypassing loads, M is a
//            template parameter so all M+1 loads of a group are in flight together), writes the map to its
//            accumulation slot AND keeps it in shared memory, and histograms the top 11 value bits on the fly.
//            The map is a variance / second moment, i.e. >= +0, so its IEEE bit pattern IS its order-preserving key.
//   phase B  exact radix select (11 + 10 + 10 bits) of the lo-th / hi-th order statistic over the DISTRIBUTED
//            shared-memory copy of the map: each CTA histograms its slice locally, one cluster barrier per level,
//            every CTA then sums the peers' histograms through DSMEM and locates the rank itself — no global-memory
//            traffic.  While the select runs, the slice of `sample` and `eps` that phase C needs is pulled into L2
//            with cp.async.bulk.prefetch, so HBM stays busy.
//   phase C  threshold = torch's lerp of the two statistics; mask, posterior blend and DDIM update on the slice,
//            u from shared memory, sample/eps from L2; x_{t-1} written once.
// HBM traffic = (M+1)*s_in + 4 (map) + s_x (sample) + s_x (x_{t-1}) bytes per element: the algorithmic minimum
// (eps is read twice, the second time from L2).
//
// Arithmetic: fp32.  Thresholds and masks are exact functions of the map (bit-identical to torch.quantile /
// compare on the same map).  The blend and the DDIM update use reciprocal-multiply instead of IEEE division
// (<= 2 ulp per operation, far inside the 1e-5 relative bar of BASELINE.json); the unfused kernels in du_step.cu keep
// one IEEE rounding per reference operation.
#include "du_fused.cuh"

namespace du {

template <typename T, int MT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fused_step_kernel(const __grid_constant__ FusedKParams kp) {
  using FV = Vec16<T>;
  constexpr int VEC = FV::VEC;
  const du_fused_params& p = kp.p;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned csize = cluster.num_blocks();
  const unsigned crank = (csize > 1) ? cluster.block_rank() : 0u;
  const int64_t b = blockIdx.y;
  const int64_t L = kp.L;
  const int64_t base = (int64_t)crank * L;  // slice start within the image
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* u_s = reinterpret_cast<float*>(smem_raw);
  uint32_t* h0 = reinterpret_cast<uint32_t*>(u_s + L);
  uint32_t* h1 = h0 + H0_WORDS;             // work area: candidate list + level-1 histogram (or levels 1 / 2 on the general path)
  uint32_t* misc = h1 + WORK_WORDS;

  const bool use_tmem = kp.tmem_cols != 0;
  // A kernel that contains tcgen05.alloc keeps a second CTA off the SM until the first one has given up its allocation
  // permit, so warp 0 relinquishes it on every path (measured: without it the non-stash launches ran one CTA per SM).
  if (tid < 32) {
    if (use_tmem) tmem_alloc(&misc[42], kp.tmem_cols);
    else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int j = tid; j < HIST_WORDS; j += THREADS) h0[j] = 0;
  if (tid < MISC_WORDS && tid != 42) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;
  if (use_tmem) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  uint32_t tbase = 0;
  if (use_tmem) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tbase = misc[42];
  }

  // Two CTAs share an SM and would otherwise run the three phases in lockstep: HBM idles while both select and the
  // SM's L2 port is contended while both update.  The second CTA of every SM (linear index >= late_from; clusters are
  // never split) starts late, so its streaming phase covers its neighbour's select + update.
  if (kp.late_ns != 0 && (blockIdx.y * gridDim.x) >= kp.late_from) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    do {
      __nanosleep(500);
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    } while (t - t0 < kp.late_ns);
  }
  stamp(kp, 0);
  // ---------------------------------------------------------------- phase A: moments + level-0 histogram
  const int mode = p.moments_mode;
  const int centre_mode = (mode == DU_MOM_CENTERED) ? 1 : ((mode == DU_MOM_VAR_WITH_CENTER) ? 2 : 0);
  const int64_t srow = b * p.score_stride + base, erow = b * p.eps_stride + base;
  float* urow = p.unc_out + b * p.unc_stride + base;
  uint32_t nan_seen = 0;
  const int ngroups = (int)(L / VEC);
  int trip_a = 0;
  for (int g = tid; g < ngroups; g += THREADS, ++trip_a) {
    float u[VEC];
    uint4 raw_c;
    moments_group<T, MT>(p, srow, erow, (uint32_t)(g * VEC), centre_mode, kp.inv_cnt, kp.inv_cm1, u, raw_c);
    if (use_tmem) tmem_st4(tmem_slot(tbase, kp.tmem_cpg, trip_a), raw_c);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      nan_seen |= (u[e] != u[e]);
      const uint32_t bin = __float_as_uint(u[e]) >> (H1_BITS + H2_BITS);   // sign bit is 0: 12 bits
      atomicAdd(&h0[bin >> 1], (bin & 1u) ? 0x10000u : 1u);
    }
#pragma unroll
    for (int h = 0; h < VEC / 4; ++h) {
      float4 u4 = make_float4(u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]);
      *reinterpret_cast<float4*>(u_s + g * VEC + 4 * h) = u4;
      *reinterpret_cast<float4*>(urow + g * VEC + 4 * h) = u4;
    }
  }
  if (__any_sync(0xffffffffu, nan_seen) && (tid & 31) == 0) misc[3] = 1u;
  if (use_tmem) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  if (tid == 0 && !p.skip_ddim) {  // phase C inputs -> L2 while the select runs
    prefetch_l2_bulk(reinterpret_cast<const char*>(p.sample) + (b * p.sample_stride + base) * (p.sample_dtype == DU_F32 ? 4 : 2),
                     (uint32_t)(L * (p.sample_dtype == DU_F32 ? 4 : 2)));
    if (!use_tmem) prefetch_l2_bulk(reinterpret_cast<const char*>(p.eps) + erow * (int64_t)sizeof(T), (uint32_t)(L * sizeof(T)));
  }

  stamp(kp, 1);
  // ---------------------------------------------------------------- phase B: exact per-image threshold
  const float thr = select_threshold<THREADS>(cluster, csize, crank, kp, u_s, h0, h1, misc);
  stamp(kp, 5);
  if (crank == 0 && tid == 0 && p.thr_out) p.thr_out[b] = thr;

  // ---------------------------------------------------------------- phase C: mask + posterior + DDIM
  const bool fast_c = p.ddim.prediction_type == DU_PRED_EPSILON && !p.ddim.use_clipped_model_output &&
                      p.sample_dtype == DU_F32 && p.prev_dtype == DU_F32 && !p.skip_ddim;
  if constexpr (sizeof(T) == 4) {
    if (fast_c && use_tmem) guided_update_slice<T, THREADS, true, true>(kp, u_s, thr, b, base, tbase);
    else if (fast_c) guided_update_slice<T, THREADS, true, false>(kp, u_s, thr, b, base, 0u);
    else guided_update_slice<T, THREADS, false, false>(kp, u_s, thr, b, base, 0u);
  } else {
    if (fast_c) guided_update_slice<T, THREADS, true, false>(kp, u_s, thr, b, base, 0u);
    else guided_update_slice<T, THREADS, false, false>(kp, u_s, thr, b, base, 0u);
  }
  if (use_tmem) {  // every warp is done with its tensor-memory columns
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) tmem_dealloc(tbase, kp.tmem_cols);
  }
  stamp(kp, 6);
  // peers may still be reading this CTA's histograms: do not exit before they are past their last DSMEM read
  if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

static size_t fused_smem_bytes(int64_t L) { return (size_t)L * 4 + (size_t)(HIST_WORDS + MISC_WORDS) * 4; }

static bool fused_plan(int64_t n, int vec, int64_t B, FusedPlan* out) {
  // override for tuning: DU_FUSED_CLUSTER=<1|2|4|8>, DU_FUSED_THREADS=<256|512|1024>
  const char* e_c = getenv("DU_FUSED_CLUSTER");
  const char* e_t = getenv("DU_FUSED_THREADS");
  if (e_c && atoi(e_c) <= 0) e_c = nullptr;  // empty / 0 = not set
  if (e_t && atoi(e_t) <= 0) e_t = nullptr;
  const size_t kMax = 227 * 1024, kHalf = 113 * 1024;
  int best_c = 0;
  for (int pass = 0; pass < 2 && !best_c; ++pass) {  // pass 0: two CTAs per SM; pass 1: anything that fits
    for (int c = 1; c <= MAX_CLUSTER; c <<= 1) {
      if (e_c && atoi(e_c) != c) continue;
      if (n % ((int64_t)c * vec) != 0) continue;
      if (fused_smem_bytes(n / c) <= (pass == 0 ? kHalf : kMax)) { best_c = c; break; }
    }
  }
  if (!best_c) return false;
  // a handful of images (SD latents, B = 1): spread each image over more SMs as long as the slices stay >= 2048 elements
  // (measured: 21 -> 14 us for the 1 x 4x64x64, M = 16 step)
  while (!e_c && B * best_c < 8 && best_c < 4 && n % ((int64_t)2 * best_c * vec) == 0 && n / (2 * best_c) >= 2048) best_c *= 2;
  const int64_t L = n / best_c;
  const size_t smem = fused_smem_bytes(L);
  const int64_t groups = L / vec;
  int threads = (groups >= 2048) ? 512 : 256;
  if (e_t) { int t = atoi(e_t); threads = (t == 1024 || t == 512) ? t : 256; }
  if (threads == 1024 && smem > kHalf) { /* one CTA per SM anyway */ }
  out->cluster = best_c; out->threads = threads; out->smem = smem;
  out->minb = (threads == 1024) ? 1 : 2;
  return true;
}

template <typename T, int MT, int THREADS, int MINB>
static int launch_fused_t(const FusedKParams& kp_in, const FusedPlan& plan, cudaStream_t st) {
  auto kern = fused_step_kernel<T, MT, THREADS, MINB>;
  // per instantiation and device: raise the dynamic shared-memory limit once, not on every launch
  static size_t smem_set[64] = {0};
  static int occ_cache[64] = {0};
  static size_t occ_smem[64] = {0};
  int dev = 0;
  DU_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || plan.smem > smem_set[dev]) {
    DU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    if (dev >= 0 && dev < 64) smem_set[dev] = plan.smem;
  }
  FusedKParams kp = kp_in;
  if (kp.tmem_cols != 0) {
    // The eps stash asks for tmem_cols tensor-memory columns per CTA.  tcgen05.alloc BLOCKS while the SM's 512 columns are
    // taken, and a blocked CTA whose cluster peers wait at a barrier could deadlock against another cluster, so the stash
    // is only used when every CTA that can be co-resident on an SM (the occupancy the runtime reports for this
    // instantiation and shared-memory size) gets its columns at once.
    int occ = 0;
    if (dev >= 0 && dev < 64 && occ_cache[dev] != 0 && occ_smem[dev] == plan.smem) occ = occ_cache[dev];
    else {
      DU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, plan.smem));
      if (dev >= 0 && dev < 64) { occ_cache[dev] = occ; occ_smem[dev] = plan.smem; }
    }
    if (occ < 1 || kp.tmem_cols * (uint32_t)occ > 512u) { kp.tmem_cols = 0; kp.tmem_cpg = 0; }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)plan.cluster, (unsigned)kp.p.B, 1);
  cfg.blockDim = dim3((unsigned)THREADS);
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

template <typename T, int MT>
static int launch_fused_m(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  if (plan.threads == 1024) return launch_fused_t<T, MT, 1024, 1>(kp, plan, st);
  if (plan.threads == 512) return launch_fused_t<T, MT, 512, 2>(kp, plan, st);
  return launch_fused_t<T, MT, 256, 2>(kp, plan, st);
}

template <typename T>
static int launch_fused(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  switch (kp.p.M) {  // the BASELINE configurations get all their loads in flight at once; any other M runs the batched loop
    case 4: return launch_fused_m<T, 4>(kp, plan, st);
    case 5: return launch_fused_m<T, 5>(kp, plan, st);
    case 8: return launch_fused_m<T, 8>(kp, plan, st);
    case 16: return launch_fused_m<T, 16>(kp, plan, st);
    default: return launch_fused_m<T, 0>(kp, plan, st);
  }
}

}  // namespace du

using namespace du;

static thread_local int g_last_fused_kernel = 0;
extern "C" int du_fused_last_kernel(void) { return g_last_fused_kernel; }

extern "C" int du_fused_supported(int64_t n, int score_dtype) {
  if (!dtype_ok(score_dtype) || n <= 0 || n > ((int64_t)1 << 24)) return 0;
  FusedPlan plan;
  return fused_plan(n, score_dtype == DU_F32 ? 4 : 8, 1 << 20, &plan) ? plan.cluster : 0;
}

extern "C" int du_fused_uncertainty_step(const du_fused_params* p, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
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
  if (!p->eps || !p->unc_out) return set_error(DU_ERR_BAD_ARG, "du_fused_uncertainty_step: null tensor");
  if (p->skip_ddim ? !p->eps_out : (!p->sample || !p->prev_out))
    return set_error(DU_ERR_BAD_ARG, p->skip_ddim ? "du_fused_uncertainty_step: skip_ddim needs eps_out" : "du_fused_uncertainty_step: null tensor");
  const int vec = p->score_dtype == DU_F32 ? 4 : 8;
  FusedPlan plan;
  if (!fused_plan(p->n, vec, p->B, &plan))
    return set_error(DU_ERR_TOO_LARGE, "du_fused_uncertainty_step: rows of %lld elements do not fit cluster shared memory; use the unfused calls", (long long)p->n);
  // 128-bit access requirements
  bool ok = aligned(p->eps, 16) && (p->eps_stride % vec == 0) && (p->score_stride % vec == 0) &&
            (p->skip_ddim || (aligned(p->sample, 16) && (p->sample_stride % (p->sample_dtype == DU_F32 ? 4 : 8) == 0))) &&
            aligned(p->unc_out, 16) && (p->unc_stride % 4 == 0) &&
            (p->skip_ddim || (vec4_ok(p->prev_out, p->prev_stride, p->prev_dtype) && vec4_ok(p->x0_out, p->x0_stride, p->prev_dtype))) &&
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
  const int count = p->M + (p->moments_mode == DU_MOM_VAR_WITH_CENTER ? 1 : 0);
  kp.inv_cnt = 1.0f / (float)count;
  kp.inv_cm1 = 1.0f / (float)(count - 1);  // count == 1 -> inf; 0 * inf = NaN like torch.var
  kp.inv_sqrt_alpha_t = 1.0f / p->ddim.sqrt_alpha_t;
  kp.timeline = nullptr;
  kp.late_from = 0; kp.late_ns = 0;
  kp.tmem_cols = 0; kp.tmem_cpg = 0;
  {
    // eps stash in tensor memory: fp32 scores whose eps is read in phase A anyway, fp32 sample/outputs (the fast update),
    // whole warps only; whether the columns of all co-resident CTAs fit the SM's 512 is decided per instantiation in
    // launch_fused_t
    const char* e_tm = getenv("DU_FUSED_TMEM");
    const int64_t ng = kp.L / 4;
    const bool fast_c = p->ddim.prediction_type == DU_PRED_EPSILON && !p->ddim.use_clipped_model_output &&
                        p->sample_dtype == DU_F32 && p->prev_dtype == DU_F32 && !p->skip_ddim;
    if (!(e_tm && atoi(e_tm) == 0) && p->score_dtype == DU_F32 && p->moments_mode != DU_MOM_VAR_UNBIASED && fast_c && ng % 32 == 0) {
      const uint32_t trips = (uint32_t)((ng + plan.threads - 1) / plan.threads);
      const uint32_t need = trips * 4u * (uint32_t)(plan.threads / 128);
      uint32_t cols = 32;
      while (cols < need) cols <<= 1;
      if (need <= 512) { kp.tmem_cols = cols; kp.tmem_cpg = trips * 4u; }
    }
  }
  if (const char* e_s = getenv("DU_FUSED_SKEW_US")) {  // experiment knob, off by default (measured: no gain)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const double skew_us = atof(e_s);
    const int64_t ctas = p->B * plan.cluster;
    if (skew_us > 0.0 && ctas > sms) {
      kp.late_from = (uint32_t)((sms / plan.cluster) * plan.cluster);
      kp.late_ns = (uint32_t)(skew_us * 1000.0);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const char* tl_path = getenv("DU_FUSED_TIMELINE");  // debug only: synchronises and writes the per-CTA phase stamps
  const size_t tl_words = (size_t)p->B * plan.cluster * TL_SLOTS;
  if (tl_path) {
    DU_CUDA(cudaMalloc(&kp.timeline, tl_words * 8));
    DU_CUDA(cudaMemsetAsync(kp.timeline, 0, tl_words * 8, st));
  }
  // the predictive single-pass kernel where the configuration is eligible (du_fused_pred.cu), else the three-phase kernel
  int rc = launch_fused_pred(kp, plan, st);
  if (rc == 1) {
    rc = DU_OK;
    g_last_fused_kernel = 2;
  } else if (rc == 0) {
    g_last_fused_kernel = 1;
    switch (p->score_dtype) {
      case DU_F32: rc = launch_fused<float>(kp, plan, st); break;
      case DU_F16: rc = launch_fused<__half>(kp, plan, st); break;
      default: rc = launch_fused<__nv_bfloat16>(kp, plan, st); break;
    }
  }
  if (tl_path) {
    unsigned long long* h = (unsigned long long*)malloc(tl_words * 8);
    cudaMemcpyAsync(h, kp.timeline, tl_words * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    FILE* f = fopen(tl_path, "w");
    if (f) {
      for (size_t c = 0; c < tl_words / TL_SLOTS; ++c) {
        for (int k = 0; k < TL_SLOTS; ++k) fprintf(f, "%llu%c", h[c * TL_SLOTS + k], k == TL_SLOTS - 1 ? '\n' : ' ');
      }
      fclose(f);
    }
    free(h);
    cudaFree(kp.timeline);
  }
  return rc;
}
