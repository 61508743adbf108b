// du_fused_pred.cu — the predictive single-pass fused uncertainty step (see the block comment below).
#include <cmath>
#include <type_traits>

#include "du_fused.cuh"

namespace du {

// =====================================================================================================================
// The PREDICTIVE single-pass step.  The three-phase kernel above leaves HBM idle while an image's threshold is selected
// and runs the update as a second, issue-bound pass; both sit on the critical path of every CTA (they all stream in
// lock-step).  Here the update is folded into the streaming pass:
//   pilot     trip 0 of every thread covers a systematic 1/trips row sample of the slice (a warp owns `trips` consecutive
//             rows and reads one per trip).  Its map values are histogrammed; after ONE cluster barrier every CTA locates
//             pilot ranks rank_p -/+ Delta (Delta = a few standard deviations of the sample quantile) in the summed pilot
//             histograms: the level-0 bins [LB, UB) that will contain the image's two order statistics.
//   stream    every other trip loads the M scores, eps, the sample (and S), computes the map value, histograms it, and
//             applies mask -> posterior -> DDIM immediately: below the band the mask is certainly 0 (1 for `lower`), above
//             it certainly 1 (0).  The few per cent inside the band are appended to a candidate list (key, element) and
//             written provisionally.  The pilot trip is then replayed from L2 through the same code.
//   finish    exact select: level-0 bin of rank lo from the FULL histogram; it must lie inside the band (verified, not
//             assumed), the keys of that bin come from the candidate list, list_select() finishes.  The candidates are
//             then patched with the exact compare.
// If the verification fails (band missed, NaN, ties overflowing the lists) the image falls back to the exact select over
// the map in L2 and a full update pass: always the exact result, only slower.
// =====================================================================================================================
// Capacity of the key list of the selected level-0 bin.  That bin holds n * (probability mass of one bin) keys — about 920 of
// the 49152 of an ImageNet-128 image at q = 0.9 (measured: 915..1025 over 1024 images).  With LIST_CAP = 1024 one image in a
// thousand took the general path and held the whole launch back by 20 us (seen as one slow rank at N = 4); 2048 entries leave
// more than 30 standard deviations.  The three-phase kernel keeps 1024 (its shared memory is sized by the map).
constexpr int PRED_LIST_CAP = 2048;
constexpr int PRED_WORK_WORDS = PRED_LIST_CAP + H1_BINS;
static_assert(PRED_LIST_CAP <= H0_WORDS, "list_select parks the tiny list in the level-0 histogram's words");
static_assert(PRED_WORK_WORDS >= WORK_WORDS && PRED_WORK_WORDS >= H0_WORDS, "the general path and the pilot histogram reuse the work area");

struct PredKParams {
  FusedKParams k;
  uint32_t trips;       // row trips per thread: ceil(groups per slice / THREADS)
  uint32_t r_lo, r_hi;  // pilot ranks bounding the band (cluster-wide pilot histogram)
  uint32_t open_low, open_high;  // band is open at that end (rank window touched the ends of the pilot sample)
  uint32_t cand_max;    // capacity of the candidate list (entries)
  uint32_t prefetch_rows;  // rows per warp pulled into L2 while the pilot runs
  uint64_t pol_stream;     // L2 eviction policy of the once-read streams (scores, sample, eps on its last read)
};

// misc words used only here: [44] candidate count, [45] list overflow, [46] LB bin, [47] UB bin
__device__ __forceinline__ void guided_elem(const du_ddim_coeffs& dc, float post_M, float inv_ah, float inv_sa, float u, float e0,
                                            float s, float S, float mk, float& eg, float& x0, float& pv) {
  // same expressions as guided_update_slice<FAST>
  const float inv_var = rcp_fast(u);
  const float prec = rcp_fast(fmaf(post_M, inv_var, inv_ah));
  const float post = prec * (inv_var * S);
  eg = fmaf(mk, post, e0 * (1.0f - mk));
  x0 = (s - dc.sqrt_beta_t * eg) * inv_sa;
  if (dc.clip_sample) x0 = clamp_nan(x0, -dc.clip_range, dc.clip_range);
  pv = fmaf(dc.sqrt_alpha_prev, x0, dc.dir_coef * eg);
}

// OUTS: some of the optional outputs (x0, guided eps, mask) are requested; the common launch writes x_{t-1} only and keeps
// none of those values alive in the (register-bound) streaming loop.
// NARROW: 16-bit scores are read as 8-byte vectors (4 elements per thread and trip, like fp32) instead of 16-byte ones: the
// loop then has the register budget and the trip count of the fp32 instance (8 elements per thread spilled and left only 6
// trips, below the point where the single pass pays).
// SPEC: 1 = the reference's percentile-guided step as its callers run it — variance over the M scores AND the centre
// (DU_MOM_VAR_WITH_CENTER), posterior sum source S given: the two facts are compile-time constants of the streaming loop (no
// per-trip selects on the moments mode, no S-or-eps select); 0 = any mode, S optional.
template <typename T, int MT, int THREADS, int MINB, bool OUTS, bool NARROW, int SPEC>
__global__ void __launch_bounds__(THREADS, MINB) fused_pred_kernel(const __grid_constant__ PredKParams pk) {
  using FV = typename std::conditional<NARROW, Vec8<T>, Vec16<T>>::type;
  constexpr int VEC = FV::VEC;
  constexpr int LOW = H1_BITS + H2_BITS;
  const FusedKParams& kp = pk.k;
  const du_fused_params& p = kp.p;
  cg::cluster_group cluster = cg::this_cluster();
  // the grid is (cluster size, B) with cluster dimensions (cluster size, 1, 1): the rank in the cluster IS blockIdx.x.  Taking it
  // from there (instead of %cluster_ctarank, which arrives in a vector register) keeps every row base address CTA-uniform.
  const unsigned csize = gridDim.x;
  const unsigned crank = blockIdx.x;
  const int64_t b = blockIdx.y;
  const int64_t L = kp.L;
  const int64_t base = (int64_t)crank * L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto sync_all = [&]() { if (csize > 1) cluster_barrier(); else __syncthreads(); };
  auto peer = [&](uint32_t* ptr, unsigned r) -> uint32_t* { return (csize > 1) ? cluster.map_shared_rank(ptr, r) : ptr; };

  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* h0 = reinterpret_cast<uint32_t*>(smem_raw);   // full level-0 histogram (packed 16-bit)
  uint32_t* work = h0 + H0_WORDS;                         // key list + level-1 histogram (or the fallback's levels 1 / 2)
  uint32_t* misc = work + PRED_WORK_WORDS;
  // The pilot histogram (packed 16-bit) shares the work area: peers read it through DSMEM only during their band search,
  // which every CTA acknowledges with a cluster-barrier arrival; the matching wait sits after the streaming pass, and only
  // then is the area cleared for the select.
  uint32_t* hp = work;
  // candidates, structure of arrays: key, element index within the slice, and the inputs of the update (eps, sample, S) so
  // that the patch pass touches no global memory for its reads
  // candidates: one 16-byte record {key, element index, eps bits, sample bits} each — a single predicated st.shared.v4 in the
  // streaming loop, and the patch pass touches no global memory for its reads
  uint4* cand = reinterpret_cast<uint4*>(misc + MISC_WORDS);
  const uint32_t cand_addr = (uint32_t)__cvta_generic_to_shared(cand);
  const uint32_t cand_cnt_addr = (uint32_t)__cvta_generic_to_shared(&misc[44]);

  for (int j = tid; j < H0_WORDS + PRED_WORK_WORDS; j += THREADS) h0[j] = 0;
  if (tid < MISC_WORDS) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;
  __syncthreads();
  stamp(kp, 0);

  const du_ddim_coeffs dc = p.ddim;
  const bool higher = p.higher != 0;
  const float post_M = p.post_M, inv_ah = p.inv_alpha_hat, inv_sa = kp.inv_sqrt_alpha_t;
  const int mode = p.moments_mode;
  const int centre_mode = SPEC ? 2 : ((mode == DU_MOM_CENTERED) ? 1 : ((mode == DU_MOM_VAR_WITH_CENTER) ? 2 : 0));
  const int64_t srow = b * p.score_stride + base, erow = b * p.eps_stride + base, xrow = b * p.sample_stride + base;
  const T* eps_row = reinterpret_cast<const T*>(p.eps) + erow;
  const float* xs = reinterpret_cast<const float*>(p.sample) + xrow;
  float* urow = p.unc_out + b * p.unc_stride + base;
  float* prow = reinterpret_cast<float*>(p.prev_out) + b * p.prev_stride + base;
  const float* Srow = (SPEC || p.S) ? (p.S + (p.S_broadcast ? 0 : b * p.S_stride) + base) : nullptr;
  const bool has_S = SPEC ? true : (Srow != nullptr);
  // mask by band as ONE unsigned compare: `higher`: key >= UB  <=>  key - UB < 2^32 - UB;  `lower`: key < LB  <=>  key - 0 < LB
  uint32_t one_lo = 0, one_w = 0;
  const int ngroups = (int)(L / VEC);
  const int trips = (int)pk.trips;
  uint32_t nan_seen = 0;
  uint32_t LB = 0, UB = 0;   // band in key space, known after the pilot
  const uint64_t pol = pk.pol_stream;

  // one group of VEC elements: map value (from the scores, or replayed from the slot), histogram, mask by band, update
  auto do_group = [&](int g, bool valid, bool from_scores, bool pilot_only) {
    float u[VEC], e0[VEC], s[VEC], Sv[VEC];
    const uint32_t g_elems = (uint32_t)g * VEC;
    if (valid) {
      const uint32_t byte_off = g_elems * (uint32_t)sizeof(T);
      // eps was just read by du_batch_sum (evict-last): it comes from L2.  The pilot reads its rows once more later (normal
      // priority); every other read is the last one and, like the scores and the sample, asks to be evicted first.
      const uint4 raw_e = FV::ld(reinterpret_cast<const char*>(eps_row) + byte_off, pilot_only ? kL2EvictNormal : pol);
      uint4 raw_s[VEC / 4];
      float4 raw_S[VEC / 4];
      if (!pilot_only) {
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) raw_s[h] = ldg_stream_128_pol(xs + g_elems + 4 * h, pol);
        if (has_S) {
#pragma unroll
          for (int h = 0; h < VEC / 4; ++h) raw_S[h] = ld_coherent_f4(Srow + g_elems + 4 * h);
        }
      }
      if (from_scores) {
        float c[VEC], k[VEC], s1[VEC], s2[VEC];
        if constexpr (MT > 1 && SPEC == 1) accumulate_scores_ct<T, MT, true, FV, true>(p.scores, srow, byte_off, raw_e, 2, false, true, c, k, s1, s2, pol);
        else if constexpr (MT > 0) accumulate_scores_ct<T, MT, true, FV>(p.scores, srow, byte_off, raw_e, centre_mode, false, centre_mode != 1, c, k, s1, s2, pol);
        else accumulate_scores<T>(p.scores, p.M, srow + g_elems, raw_e, centre_mode, false, centre_mode != 1, c, k, s1, s2);
#pragma unroll
        for (int e = 0; e < VEC; ++e) u[e] = map_value(centre_mode, s1[e], s2[e], kp.inv_cnt, kp.inv_cm1);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          nan_seen |= (u[e] != u[e]);
          const uint32_t bin = __float_as_uint(u[e]) >> LOW;
          const uint32_t inc = (bin & 1u) ? 0x10000u : 1u;
          atomicAdd(&h0[bin >> 1], inc);
          if (pilot_only) atomicAdd(&hp[bin >> 1], inc);
        }
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h)
          *reinterpret_cast<float4*>(urow + g_elems + 4 * h) = make_float4(u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]);
      } else {
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {   // replay: this thread wrote these values itself
          const float4 u4 = *reinterpret_cast<const float4*>(urow + g_elems + 4 * h);
          u[4 * h] = u4.x; u[4 * h + 1] = u4.y; u[4 * h + 2] = u4.z; u[4 * h + 3] = u4.w;
        }
      }
      if (!pilot_only) {
        FV::unpack(raw_e, e0);
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {
          s[4 * h] = __uint_as_float(raw_s[h].x); s[4 * h + 1] = __uint_as_float(raw_s[h].y);
          s[4 * h + 2] = __uint_as_float(raw_s[h].z); s[4 * h + 3] = __uint_as_float(raw_s[h].w);
          if (has_S) { Sv[4 * h] = raw_S[h].x; Sv[4 * h + 1] = raw_S[h].y; Sv[4 * h + 2] = raw_S[h].z; Sv[4 * h + 3] = raw_S[h].w; }
          else { Sv[4 * h] = e0[4 * h]; Sv[4 * h + 1] = e0[4 * h + 1]; Sv[4 * h + 2] = e0[4 * h + 2]; Sv[4 * h + 3] = e0[4 * h + 3]; }
        }
      }
    }
    if (pilot_only) return;
    // ---- band test and candidate entries first (while eps / sample are freshly unpacked: the streaming loop lives at the
    // register limit, and every value kept alive across the update costs loads in flight), then mask, update, stores
    uint32_t inband = 0;
    if (valid) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) inband |= (((__float_as_uint(u[e]) - LB) < (UB - LB)) ? 1u : 0u) << e;
    }
    {
      // The streaming loop is issue-bound, so the append has no branches: one predicated shared-memory atomic per lane that has
      // candidates (about a quarter of the lanes), then one predicated 16-byte store per candidate.  Order does not matter.
      const uint32_t cnt = __popc(inband);
      uint32_t slot = 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p atom.shared.add.u32 %0, [%1], %3;\n\t}"
                   : "+r"(slot) : "r"(cand_cnt_addr), "r"(inband), "r"(cnt) : "memory");
      const bool room = slot + cnt <= pk.cand_max;
      if (inband != 0u && !room) misc[45] = 1u;
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const uint32_t take = (room && (inband & (1u << e))) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.u32 [%0], {%1, %2, %3, %4};\n\t}"
                     ::"r"(cand_addr + slot * 16u), "r"(__float_as_uint(u[e])), "r"(g_elems + e), "r"(__float_as_uint(e0[e])),
                       "r"(__float_as_uint(s[e])), "r"(take) : "memory");
        slot += take;
      }
    }
    if (valid) {
      float pv[VEC], x0v[VEC], eg[VEC], mk[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        mk[e] = ((__float_as_uint(u[e]) - one_lo) < one_w) ? 1.0f : 0.0f;
        guided_elem(dc, post_M, inv_ah, inv_sa, u[e], e0[e], s[e], Sv[e], mk[e], eg[e], x0v[e], pv[e]);
      }
#pragma unroll
      for (int h = 0; h < VEC / 4; ++h) {
        const int64_t o = g_elems + 4 * h;
        *reinterpret_cast<float4*>(prow + o) = make_float4(pv[4 * h], pv[4 * h + 1], pv[4 * h + 2], pv[4 * h + 3]);
        if constexpr (OUTS) {
          if (p.x0_out) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.x0_out) + b * p.x0_stride + base + o) =
              make_float4(x0v[4 * h], x0v[4 * h + 1], x0v[4 * h + 2], x0v[4 * h + 3]);
          if (p.eps_out) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.eps_out) + b * p.eps_out_stride + base + o) =
              make_float4(eg[4 * h], eg[4 * h + 1], eg[4 * h + 2], eg[4 * h + 3]);
          if (p.mask_out) *reinterpret_cast<float4*>(p.mask_out + b * p.mask_out_stride + base + o) =
              make_float4(mk[4 * h], mk[4 * h + 1], mk[4 * h + 2], mk[4 * h + 3]);
        }
      }
    }
  };
  auto group_of = [&](int j) { return ((warp * trips + j) << 5) | lane; };   // a warp owns `trips` consecutive rows of 32 groups

  // ---------------------------------------------------------------- pilot: trip 0, map values only
  // The pilot (first loads of a cold launch, one cluster barrier, two rank searches) leaves HBM idle for a few microseconds:
  // one lane per warp pulls the warp's next rows of every input into L2 meanwhile (a warp's rows are contiguous).
  if (lane == 0 && pk.prefetch_rows > 0) {
    const int row1 = warp * trips + 1;
    const int rows = min((int)pk.prefetch_rows, min(trips - 1, ngroups / 32 - row1));
    if (rows > 0) {
      const size_t off = (size_t)row1 * 32 * VEC;   // elements
      const uint32_t bytes_t = (uint32_t)rows * 32u * VEC * (uint32_t)sizeof(T), bytes_x = (uint32_t)rows * 32u * VEC * 4u;
      for (int m = 0; m < p.M; ++m) prefetch_l2_bulk(reinterpret_cast<const T*>(p.scores[m]) + srow + off, bytes_t);
      prefetch_l2_bulk(eps_row + off, bytes_t);
      prefetch_l2_bulk(xs + off, bytes_x);
    }
  }
  {
    const int g = group_of(0);
    do_group(g, g < ngroups, true, true);
  }
  sync_all();   // pilot histograms of every CTA are complete
  locate_rank<H0_BINS, THREADS, false, true, true>(cluster, csize, hp, pk.r_lo, misc, pk.r_hi);
  const uint32_t lb_bin = pk.open_low ? 0u : misc[0];
  const uint32_t ub_bin = pk.open_high ? (uint32_t)H0_BINS : misc[1] + 1u;
  if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // done with the peers' pilot histograms
  __syncthreads();
  LB = lb_bin << LOW;
  UB = ub_bin << LOW;   // 4096 << 19 = 2^31: above every finite key and every NaN pattern with sign 0
  one_lo = higher ? UB : 0u;
  one_w = higher ? (0u - UB) : LB;   // (UB = 0 cannot occur: ub_bin >= 1)
  stamp(kp, 1);

  // ---------------------------------------------------------------- stream: every other trip, then the pilot replayed
  // Launched as the programmatic dependent of the du_batch_sum that produces S (S_overlap): everything above ran while that
  // kernel was still summing; its result is complete and visible from here on.  A no-op for an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int j = 1; j < trips; ++j) {
    const int g = group_of(j);
    do_group(g, g < ngroups, true, false);
  }
  {
    const int g = group_of(0);
    do_group(g, g < ngroups, false, false);
  }
  if (__any_sync(0xffffffffu, nan_seen) && lane == 0) misc[3] = 1u;
  stamp(kp, 2);
  if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // nobody reads this CTA's pilot histogram any more
  for (int j = tid; j < PRED_WORK_WORDS; j += THREADS) work[j] = 0;

  // ---------------------------------------------------------------- finish: verify the band, exact select, patch
  sync_all();   // full histograms, candidate lists and flags of every CTA are complete
  bool bad = false;
  for (unsigned r = 0; r < csize; ++r) {
    const uint32_t* pm = peer(misc, r);
    bad |= (pm[3] != 0u) | (pm[45] != 0u);
  }
  uint32_t d0 = 0, below0 = 0;
  if (!bad) {
    locate_rank<H0_BINS, THREADS, false, true>(cluster, csize, h0, kp.lo, misc);
    d0 = misc[0]; below0 = misc[1];
    const uint32_t cnt0 = misc[2];
    __syncthreads();
    bad = !(d0 >= lb_bin && d0 < ub_bin) || cnt0 > (uint32_t)PRED_LIST_CAP;
  }
  stamp(kp, 3);
  float thr;
  bool redo = false;
  if (bad) {
    // exact select over the map in its slot (L2) and a full update pass (`bad` is identical in every CTA of the cluster)
    thr = select_threshold<THREADS>(cluster, csize, crank, kp, urow, h0, work, misc);
    redo = true;
  } else {
    uint32_t* list = work;
    uint32_t* h1 = work + PRED_LIST_CAP;
    const uint32_t want = d0 << LOW;
    const uint32_t ncand = misc[44];
    for (uint32_t i = tid; i < ncand; i += THREADS) {
      const uint32_t key = cand[i].x;
      if ((key >> LOW) == d0) {
        list[atomicAdd(&misc[6], 1u)] = key;
        atomicAdd(&h1[(key >> H2_BITS) & (H1_BINS - 1)], 1u);
      }
    }
    sync_all();   // key lists and level-1 histograms complete
    stamp(kp, 4);
    bool has_nan = false;
    const bool need_next = kp.hi > kp.lo;
    list_select<THREADS>(cluster, csize, crank, kp, h0, list, h1, misc, want, below0, has_nan);
    const uint32_t key_lo = misc[41];
    uint32_t key_hi = misc[43];
    if (need_next && key_hi == 0xffffffffu) {
      // the successor lives in a higher level-0 bin: smallest candidate key above key_lo, cluster-wide ...
      if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
      uint32_t best = 0xffffffffu;
      for (uint32_t i = tid; i < ncand; i += THREADS) { const uint32_t key = cand[i].x; if (key > key_lo) best = min(best, key); }
      best = __reduce_min_sync(0xffffffffu, best);
      if (lane == 0 && best != 0xffffffffu) atomicMin(&misc[4], best);
      sync_all();
      for (unsigned r = 0; r < csize; ++r) key_hi = min(key_hi, peer(misc, r)[4]);
      if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
      if (key_hi == 0xffffffffu) {
        // ... or above the band: one pass over the map in its slot
        if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        const int ng4 = (int)(L / 4);
        for (int g = tid; g < ng4; g += THREADS) {
          const uint4 v = *reinterpret_cast<const uint4*>(urow + 4 * g);
          const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) best = min(best, (kk[e] > key_lo) ? kk[e] : 0xffffffffu);
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (lane == 0 && best != 0xffffffffu) atomicMin(&misc[4], best);
        sync_all();
        for (unsigned r = 0; r < csize; ++r) key_hi = min(key_hi, peer(misc, r)[4]);
        if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        redo = true;   // elements at key_hi were decided without the exact compare
      }
    }
    thr = lerp_torch(__uint_as_float(key_lo), __uint_as_float(key_hi), kp.w, p.lerp_fma);
    stamp(kp, 5);
    if (!redo) {
      // ---- patch the candidates with the exact compare (all inputs come from the shared-memory stash)
      for (uint32_t i = tid; i < ncand; i += THREADS) {
        const uint4 rec = cand[i];
        const float u = __uint_as_float(rec.x);
        const int64_t o = rec.y;
        const float mk = (higher ? (u > thr) : (u < thr)) ? 1.0f : 0.0f;
        float eg, x0, pv;
        const float e0 = __uint_as_float(rec.z);
        const float Sv = has_S ? ld_coherent_f1(Srow + o) : e0;   // the S row is shared by every image of the batch: L2 / L1 resident
        guided_elem(dc, post_M, inv_ah, inv_sa, u, e0, __uint_as_float(rec.w), Sv, mk, eg, x0, pv);
        prow[o] = pv;
        if constexpr (OUTS) {
          if (p.x0_out) reinterpret_cast<float*>(p.x0_out)[b * p.x0_stride + base + o] = x0;
          if (p.eps_out) reinterpret_cast<float*>(p.eps_out)[b * p.eps_out_stride + base + o] = eg;
          if (p.mask_out) p.mask_out[b * p.mask_out_stride + base + o] = mk;
        }
      }
    }
  }
  if (redo) guided_update_slice<T, THREADS, true, false>(kp, urow, thr, b, base, 0u);
  if (crank == 0 && tid == 0 && p.thr_out) p.thr_out[b] = thr;
  stamp(kp, 6);
  if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}


static constexpr size_t kPredFixedBytes = (size_t)(H0_WORDS + PRED_WORK_WORDS + MISC_WORDS) * 4;

template <typename T, int MT>
constexpr bool pred_narrow() { return sizeof(T) == 2 && MT > 0; }

template <typename T, int MT, int THREADS, int MINB, bool OUTS, int SPEC>
static int launch_pred_ts(const PredKParams& pk, const FusedPlan& plan, size_t smem, cudaStream_t st) {
  auto kern = fused_pred_kernel<T, MT, THREADS, MINB, OUTS, pred_narrow<T, MT>(), SPEC>;
  static size_t smem_set[64] = {0};
  int dev = 0;
  DU_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || smem > smem_set[dev]) {
    DU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) smem_set[dev] = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)plan.cluster, (unsigned)pk.k.p.B, 1);
  cfg.blockDim = dim3((unsigned)THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pk.k.p.S && pk.k.p.S_overlap) {   // programmatic dependent of the preceding du_batch_sum (DU_FUSED_PDL=0 disables)
    const char* e_pdl = getenv("DU_FUSED_PDL");
    if (!(e_pdl && atoi(e_pdl) == 0)) {
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 2;
    }
  }
  DU_CUDA(cudaLaunchKernelEx(&cfg, kern, pk));
  return 1;
}

template <typename T, int MT, int THREADS, int MINB, bool OUTS>
static int launch_pred_t(const PredKParams& pk, const FusedPlan& plan, size_t smem, cudaStream_t st) {
  // the specialised streaming loop for the step the reference's callers run (the bench / pipeline configuration); the optional
  // outputs are a test / debugging feature and keep the generic loop
  if constexpr (!OUTS && MT > 0) {
    if (pk.k.p.moments_mode == DU_MOM_VAR_WITH_CENTER && pk.k.p.S != nullptr)
      return launch_pred_ts<T, MT, THREADS, MINB, OUTS, 1>(pk, plan, smem, st);
  }
  return launch_pred_ts<T, MT, THREADS, MINB, OUTS, 0>(pk, plan, smem, st);
}

template <typename T, int MT, bool OUTS>
static int launch_pred_m(const PredKParams& pk, const FusedPlan& plan, int threads, size_t smem, cudaStream_t st) {
#ifdef DU_PRED_DEV   // development builds: one thread configuration (csrc/build.sh -DDU_PRED_DEV), a third of the compile time
  if (threads != 512) return 0;
  return launch_pred_t<T, MT, 512, 2, OUTS>(pk, plan, smem, st);
#else
  switch (threads) {
#ifdef DU_PRED_TUNING   // 1024-thread and 384 / 768-thread CTAs (80 registers): measured slower, kept for sweeps only (csrc/build.sh -DDU_PRED_TUNING)
    case 1024: return launch_pred_t<T, MT, 1024, 1, OUTS>(pk, plan, smem, st);
    case 768: return launch_pred_t<T, MT, 768, 1, OUTS>(pk, plan, smem, st);
    case 384: return launch_pred_t<T, MT, 384, 2, OUTS>(pk, plan, smem, st);
#endif
    case 512: return launch_pred_t<T, MT, 512, 2, OUTS>(pk, plan, smem, st);
    default: return 0;
  }
#endif
}

template <typename T>
static int launch_pred(const PredKParams& pk, const FusedPlan& plan, int threads, size_t smem, cudaStream_t st) {
  const bool outs = pk.k.p.x0_out || pk.k.p.eps_out || pk.k.p.mask_out;
  if (outs) {   // the optional outputs are a test / debugging feature: one instantiation per M class is enough
    return pk.k.p.M == 5 ? launch_pred_m<T, 5, true>(pk, plan, threads, smem, st) : launch_pred_m<T, 0, true>(pk, plan, threads, smem, st);
  }
  switch (pk.k.p.M) {
    case 4: return launch_pred_m<T, 4, false>(pk, plan, threads, smem, st);
    case 5: return launch_pred_m<T, 5, false>(pk, plan, threads, smem, st);
    case 8: return launch_pred_m<T, 8, false>(pk, plan, threads, smem, st);
    case 16: return launch_pred_m<T, 16, false>(pk, plan, threads, smem, st);
    default: return launch_pred_m<T, 0, false>(pk, plan, threads, smem, st);
  }
}

int launch_fused_pred(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  const du_fused_params& p = kp.p;
  // DU_FUSED_PRED=0 forces the three-phase kernel; DU_FUSED_BAND_SIGMA=<x> sets the half-width of the rank window in
  // standard deviations of the pilot quantile (default 6)
  const char* e_p = getenv("DU_FUSED_PRED");
  if (e_p && atoi(e_p) == 0) return 0;
  const bool fast_c = p.ddim.prediction_type == DU_PRED_EPSILON && !p.ddim.use_clipped_model_output &&
                      p.sample_dtype == DU_F32 && p.prev_dtype == DU_F32 && !p.skip_ddim;
  if (!fast_c || (plan.threads != 512 && plan.threads != 1024)) return 0;
  // elements per thread and trip: 4, except 16-bit scores with a runtime-M instance (16-byte vectors of 8)
  const bool outs_req = p.x0_out || p.eps_out || p.mask_out;
  const bool mt_known = outs_req ? (p.M == 5) : (p.M == 4 || p.M == 5 || p.M == 8 || p.M == 16);   // mirrors launch_pred
  const int vec = (p.score_dtype == DU_F32 || mt_known) ? 4 : 8;
  const int64_t L = kp.L, ngroups = L / vec;
  if (L >= 65536) return 0;   // packed 16-bit level-0 counters
  // Threads per CTA: the plan's (512 x 2 CTAs or 1024 x 1 per SM, 64 registers).  384 / 768 threads with 80 registers keep
  // all M + 3 loads of a trip in one batch but measured slower (54.1 vs 49.3 us on the ImageNet-128 step): the loop is
  // closer to issue-bound than to latency-bound.  DU_FUSED_PRED_THREADS=<384|512|768|1024> overrides.
  int threads = plan.threads;
  if (const char* e_t = getenv("DU_FUSED_PRED_THREADS")) {
    const int t = atoi(e_t);
    if (t == 384 || t == 512 || t == 768 || t == 1024) threads = t;
  }
  const int64_t trips = (ngroups + threads - 1) / threads;
  // The pilot (~4 us) and the latency-bound finish (~8 us) are fixed costs: below ~8 trips per thread the three-phase kernel
  // is faster (ImageNet-64, b128: 17.4 us against 18.8 us), above it the single pass wins (ImageNet-128: 49.2 against 54.3).
  const char* e_m = getenv("DU_FUSED_PRED_MIN_TRIPS");
  const int64_t min_trips = (e_m && atoi(e_m) >= 4) ? atoi(e_m) : 8;
  if (trips < min_trips) return 0;
  // pilot sample: trip 0 of every warp = row (warp * trips) of 32 groups
  int64_t pilot_groups = 0;
  for (int w = 0; w < threads / 32; ++w) {
    const int64_t left = ngroups - (int64_t)w * trips * 32;
    pilot_groups += left <= 0 ? 0 : (left < 32 ? left : 32);
  }
  const double n_p = (double)pilot_groups * vec * plan.cluster;
  if (n_p < 256) return 0;
  const char* e_s = getenv("DU_FUSED_BAND_SIGMA");
  const double nsig = (e_s && atof(e_s) > 0.0) ? atof(e_s) : 6.0;
  const double q = (double)p.q;
  const double rank_p = q * (n_p - 1.0);
  const double delta = nsig * std::sqrt(n_p * q * (1.0 - q)) + 4.0;
  PredKParams pk;
  pk.k = kp;
  pk.k.tmem_cols = 0; pk.k.tmem_cpg = 0;
  pk.trips = (uint32_t)trips;
  pk.open_low = (rank_p - delta <= 0.0) ? 1u : 0u;
  pk.open_high = (rank_p + delta >= n_p - 1.0) ? 1u : 0u;
  pk.r_lo = pk.open_low ? 0u : (uint32_t)std::floor(rank_p - delta);
  pk.r_hi = pk.open_high ? (uint32_t)(n_p - 1.0) : (uint32_t)std::ceil(rank_p + delta);
  // candidate list: the band holds about 2 * delta / n_p of the elements plus two level-0 bins; room for 1.6x that, and
  // no more: global loads in flight are buffered in L1, which shares the SM's 256 KB with shared memory, so a large
  // shared-memory footprint throttles the streaming pass (measured: 2 x 116 KB per SM cost 10 us against 2 x 78 KB).
  // DU_FUSED_SMEM_KB caps the per-CTA footprint (default 64).
  const char* e_k = getenv("DU_FUSED_SMEM_KB");
  const size_t cap_bytes = (size_t)((e_k && atoi(e_k) > 0) ? atoi(e_k) : 64) * 1024;
  const size_t kMax = 227 * 1024, kHalf = 113 * 1024;
  const bool one_cta = threads >= 768;
  size_t limit = one_cta ? kMax : kHalf;
  if (cap_bytes * (one_cta ? 2 : 1) < limit) limit = cap_bytes * (one_cta ? 2 : 1);
  const double frac = 2.0 * delta / n_p + 0.02;
  int64_t want = (int64_t)(1.6 * frac * (double)L) + 128;
  if (want > L) want = L;
  int64_t cap = (int64_t)((limit - kPredFixedBytes) / 16);
  if (want < cap) cap = want;
  cap &= ~(int64_t)3;
  if (cap < 256) return 0;
  pk.cand_max = (uint32_t)cap;
  {
    const char* e_r = getenv("DU_FUSED_PREFETCH_ROWS");
    pk.prefetch_rows = e_r ? (uint32_t)atoi(e_r) : 0u;   // measured: the extra traffic delays the pilot more than it saves
  }
  {
    // DU_L2_HINTS=0: every load at normal priority (A/B of the eviction hints)
    const char* e_h = getenv("DU_L2_HINTS");
    pk.pol_stream = (e_h && atoi(e_h) == 0) ? kL2EvictNormal : kL2EvictFirst;
  }
  const size_t smem = kPredFixedBytes + (size_t)cap * 16;
  switch (p.score_dtype) {
    case DU_F32: return launch_pred<float>(pk, plan, threads, smem, st);
    case DU_F16: return launch_pred<__half>(pk, plan, threads, smem, st);
    default: return launch_pred<__nv_bfloat16>(pk, plan, threads, smem, st);
  }
}

}  // namespace du
