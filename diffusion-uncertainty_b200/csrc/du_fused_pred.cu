// du_fused_pred.cu — the predictive single-pass fused uncertainty step (see the block comment below).
#include <cmath>
#include <type_traits>

#include "du_fused.cuh"

namespace du {

// =====================================================================================================================
// The PREDICTIVE single-pass step.  The three-phase kernel leaves HBM idle while an image's threshold is selected and runs the
// update as a second pass; both sit on the critical path of every CTA (they all stream in lock-step).  Here the update is
// folded into the streaming pass:
//   pilot     trip 0 of every thread covers a systematic 1/trips row sample of the slice (a warp owns `trips` consecutive
//             rows and reads one per trip).  Its map values are histogrammed (4096 level-0 bins); after ONE cluster barrier every
//             CTA locates pilot ranks rank_p -/+ Delta (Delta = a few standard deviations of the sample quantile) in the summed
//             pilot histograms: the key range [LB, UB) that will contain the image's two order statistics.
//   stream    every other trip loads the M scores, eps, the sample (and S), computes the map value and applies mask ->
//             posterior -> DDIM immediately: below the band the mask is certainly 0 (1 for `lower`), above it certainly 1 (0).
//             Keys below the band are only COUNTED (a register); the few per cent inside it are histogrammed at FINE resolution
//             (1024 bins across the band), appended to a candidate list (key, element, eps, sample) and written provisionally.
//             The pilot trip is then replayed from L2 through the same code.
//   finish    rank lo - (keys below the band) must fall inside the band (verified, never assumed); ONE search over the summed
//             fine histograms gives the bin of ~3 keys that holds it, every CTA picks those keys out of the cluster's candidate
//             lists (DSMEM reads, no copy) and warp 0 ranks them.  The candidates are then patched with the exact compare.
//             Two cluster barriers in all (pilot, end of stream) plus one split arrive / wait before exit.
// If the verification fails (band missed, NaN, candidate / tie overflow) the image falls back to the exact select over the map
// in L2 (level-0 histogram rebuilt there) and a full update pass: always the exact result, only slower.
// =====================================================================================================================
constexpr int FINE_BINS = H1_BINS;          // 1024 bins across the band (the search reuses locate_rank<H1_BINS>)
constexpr int TINY_CAP = H0_WORDS;          // keys of the selected fine bin, parked in the (then free) pilot-histogram words
constexpr int PRED_WORK_WORDS = (FINE_BINS > WORK_WORDS) ? FINE_BINS : WORK_WORDS;   // fine histogram; the fallback's work area
constexpr int ROWCTR_WORDS = 32;            // one claim counter per warp (<= 32 warps per CTA)

struct PredKParams {
  FusedKParams k;
  uint32_t trips;       // row trips per thread: ceil(groups per slice / THREADS)
  uint32_t r_lo, r_hi;  // pilot ranks bounding the band (cluster-wide pilot histogram)
  uint32_t open_low, open_high;  // band is open at that end (rank window touched the ends of the pilot sample)
  uint32_t cand_max;    // capacity of the candidate list (entries)
  uint32_t steal;       // warps that ran out of rows claim rows of the other blocks of the cluster (DU_FUSED_STEAL=0 disables)
  uint64_t pol_stream;     // L2 eviction policy of the once-read streams (scores, sample, eps on its last read)
};

// misc words used only here: [44] candidate count, [45] list overflow, [46] keys below the band
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

// The exact select inside the band.  Precondition: every CTA of the cluster has its fine histogram `hb`, its candidate records
// (count in misc[44]) and its flags complete, and a cluster barrier has been passed.  `r` = rank of key_lo among the in-band keys
// of the cluster.  Every CTA locates the fine bin in the summed histograms, copies the keys of that bin out of the cluster's
// candidate lists (DSMEM reads) and warp 0 ranks them.  Results: misc[41] = key_lo, misc[43] = key_hi (0xffffffff: the successor
// is not inside the band); returns false (cluster-uniform) when the bin holds more keys than `tiny` takes (heavy ties).
// Leaves ONE cluster-barrier arrival pending (issued after the last DSMEM read).
template <int THREADS>
__device__ __forceinline__ bool band_select(cg::cluster_group& cluster, unsigned csize, unsigned crank, const FusedKParams& kp,
                                            uint32_t* tiny, const uint4* cand, uint32_t* hb, uint32_t* misc, uint32_t LB, uint32_t sh,
                                            uint32_t r) {
  const int tid = threadIdx.x, lane = tid & 31;
  locate_rank<FINE_BINS, THREADS, true, false>(cluster, csize, hb, r, misc);
  stamp(kp, 7);
  const uint32_t f1 = misc[0], below1 = misc[1], cnt1 = misc[2], next1 = misc[5];
  const uint32_t k2 = r - below1;                        // rank of key_lo among the cnt1 keys of the fine bin
  const bool need_next = kp.hi > kp.lo;                  // the upper statistic is the next key in sorted order
  const bool fits = cnt1 <= (uint32_t)TINY_CAP;
  if (fits) {
    const bool want_next_bin = need_next && (k2 + 1 >= cnt1) && next1 < (uint32_t)FINE_BINS;
    uint32_t best = 0xffffffffu;
    for (unsigned i = 0; i < csize; ++i) {
      const unsigned rr = (crank + i) % csize;
      const uint32_t* pm = (csize > 1) ? cluster.map_shared_rank(misc, rr) : misc;
      const uint4* pc = (csize > 1) ? cluster.map_shared_rank(const_cast<uint4*>(cand), rr) : cand;
      const uint32_t cr = pm[44];
      // keys first (4 independent shared-memory / DSMEM loads in flight per thread), then the filter: the loop is latency-bound
      for (uint32_t j0 = tid; j0 < cr; j0 += 4 * THREADS) {
        uint32_t key[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t j = j0 + q4 * THREADS;
          key[q4] = (j < cr) ? pc[j].x : 0xffffffffu;     // (0xffffffff - LB) >> sh is beyond every fine bin
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t fb = (key[q4] - LB) >> sh;
          if (fb == f1) tiny[atomicAdd(&misc[40], 1u)] = key[q4];
          else if (want_next_bin && fb == next1) best = min(best, key[q4]);
        }
      }
    }
    if (want_next_bin) {
      best = __reduce_min_sync(0xffffffffu, best);
      if (lane == 0 && best != 0xffffffffu) atomicMin(&misc[4], best);
    }
  }
  if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");  // peers' lists, histograms and flags are read
  stamp(kp, 8);
  __syncthreads();
  if (fits && tid < 32) {
    // warp 0: k2-th smallest (and its successor) of the cnt1 keys in `tiny` — normally a handful: one key per lane, rank = number
    // of smaller keys (ties broken by position) from m independent shuffle rounds
    const uint32_t m = misc[40];   // == cnt1
    const uint32_t bin_lo = LB + (f1 << sh);
    uint32_t key_lo, key_hi;
    if (m <= 32u) {
      const uint32_t mine = (lane < (int)m) ? tiny[lane] : 0xffffffffu;
      uint32_t rank = 0;
      for (uint32_t j = 0; j < m; ++j) {
        const uint32_t kj = __shfl_sync(0xffffffffu, mine, (int)j);
        rank += (kj < mine || (kj == mine && j < (uint32_t)lane)) ? 1u : 0u;
      }
      const uint32_t is_lo = __ballot_sync(0xffffffffu, lane < (int)m && rank == k2);
      const uint32_t is_hi = __ballot_sync(0xffffffffu, lane < (int)m && rank == k2 + 1u);
      key_lo = __shfl_sync(0xffffffffu, mine, __ffs((int)is_lo) - 1);
      key_hi = key_lo;
      if (need_next) key_hi = is_hi ? __shfl_sync(0xffffffffu, mine, __ffs((int)is_hi) - 1) : misc[4];
    } else {
      // many keys in one fine bin (ties): bitwise search over the `sh` low bits of key - bin_lo
      uint32_t ans = 0;
#pragma unroll 1
      for (int bit = (int)sh - 1; bit >= 0; --bit) {
        const uint32_t trial = ans | (1u << bit);
        uint32_t c = 0;
        for (uint32_t j = lane; j < m; j += 32) c += ((tiny[j] - bin_lo) < trial);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c <= k2) ans |= (1u << bit);
      }
      key_lo = bin_lo + ans;
      key_hi = key_lo;
      if (need_next) {
        // successor: a tie, else the smallest larger key of the bin, else the smallest key of the next non-empty fine bin
        uint32_t le = 0, above = 0xffffffffu;
        for (uint32_t j = lane; j < m; j += 32) {
          const uint32_t key = tiny[j];
          le += (key <= key_lo);
          if (key > key_lo) above = min(above, key);
        }
        le = __reduce_add_sync(0xffffffffu, le);
        above = __reduce_min_sync(0xffffffffu, above);
        if (k2 + 1 < le) key_hi = key_lo;
        else if (above != 0xffffffffu) key_hi = above;
        else key_hi = misc[4];   // 0xffffffff: the successor lies above the band
      }
    }
    if (lane == 0) { misc[41] = key_lo; misc[43] = key_hi; }
  }
  stamp(kp, 10);
  __syncthreads();
  return fits;
}

// OUTS: some of the optional outputs (x0, guided eps, mask) are requested; the common launch writes x_{t-1} only and keeps
// none of those values alive in the streaming loop.
// NARROW: 16-bit scores are read as 8-byte vectors (4 elements per thread and trip, like fp32) instead of 16-byte ones: the
// loop then has the register budget and the trip count of the fp32 instance.
// SPEC: 1 = the reference's percentile-guided step as its callers run it — variance over the M scores AND the centre
// (DU_MOM_VAR_WITH_CENTER), posterior sum source S given: both are compile-time constants of the streaming loop; 0 = any mode.
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
  // [hp: level-0 histogram of the pilot sample, packed 16-bit; peers read it through DSMEM during their band search only; later the
  //  tiny key list of the finish (or the level-0 histogram of the fallback)] [work: the fine band histogram (or the fallback's
  //  levels 1 / 2)] [misc] [candidate records]
  uint32_t* hp = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* work = hp + H0_WORDS;
  uint32_t* hb = work;
  uint32_t* misc = work + PRED_WORK_WORDS;
  // candidates: one 16-byte record {key, element index, eps bits, sample bits} each — a single predicated st.shared.v4 in the
  // streaming loop, and the patch pass touches no global memory for its reads
  // rowctr[w]: the next row trip of warp w's block of rows nobody has claimed yet (work stealing, see the streaming loop)
  uint32_t* rowctr = misc + MISC_WORDS;
  uint4* cand = reinterpret_cast<uint4*>(rowctr + ROWCTR_WORDS);
  const uint32_t cand_addr = (uint32_t)__cvta_generic_to_shared(cand);
  const uint32_t cand_cnt_addr = (uint32_t)__cvta_generic_to_shared(&misc[44]);
  const uint32_t hb_addr = (uint32_t)__cvta_generic_to_shared(hb);

  for (int j = tid; j < H0_WORDS + PRED_WORK_WORDS; j += THREADS) hp[j] = 0;
  if (tid < MISC_WORDS) misc[tid] = (tid == 4) ? 0xffffffffu : 0u;
  if (tid < ROWCTR_WORDS) rowctr[tid] = 1u;      // trip 0 of every block is the pilot row
  __syncthreads();
  stamp(kp, 0);

  const du_ddim_coeffs dc = p.ddim;
  const bool higher = p.higher != 0;
  const float post_M = p.post_M, inv_ah = p.inv_alpha_hat, inv_sa = kp.inv_sqrt_alpha_t;
  const int mode = p.moments_mode;
  const int centre_mode = SPEC ? 2 : ((mode == DU_MOM_CENTERED) ? 1 : ((mode == DU_MOM_VAR_WITH_CENTER) ? 2 : 0));
  // all row pointers address the IMAGE (not this CTA's slice): a warp may process rows of any CTA of its cluster (work stealing)
  const int64_t srow = b * p.score_stride, erow = b * p.eps_stride, xrow = b * p.sample_stride;
  const T* eps_row = reinterpret_cast<const T*>(p.eps) + erow;
  const float* xs = reinterpret_cast<const float*>(p.sample) + xrow;
  float* uimg = p.unc_out + b * p.unc_stride;
  float* urow = uimg + base;                      // this CTA's slice of the map (the rare paths work per slice)
  float* prow = reinterpret_cast<float*>(p.prev_out) + b * p.prev_stride;
  const float* Srow = (SPEC || p.S) ? (p.S + (p.S_broadcast ? 0 : b * p.S_stride)) : nullptr;
  const bool has_S = SPEC ? true : (Srow != nullptr);
  const int ngroups = (int)(L / VEC);
  const int trips = (int)pk.trips;
  uint32_t nan_seen = 0;
  uint32_t below = 0;        // keys under the band seen by this thread
  uint32_t LB = 0, W = 0, sh = 0;   // band [LB, LB + W) in key space and the shift of the fine bins, known after the pilot
  // mask by band as ONE unsigned compare: `higher`: key >= UB  <=>  key - UB < 2^32 - UB;  `lower`: key < LB  <=>  key - 0 < LB
  uint32_t one_lo = 0, one_w = 0;
  const uint64_t pol = pk.pol_stream;

  // one group of VEC elements: map value (from the scores, or replayed from the slot), band bookkeeping, mask by band, update
  auto do_group = [&](uint32_t g_elems /* first element of the group, in the image */, bool valid, bool from_scores, bool pilot_only) {
    float u[VEC], e0[VEC], s[VEC], Sv[VEC];
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
        for (int e = 0; e < VEC; ++e) nan_seen |= (u[e] != u[e]);
        if (pilot_only) {
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const uint32_t bin = __float_as_uint(u[e]) >> LOW;
            atomicAdd(&hp[bin >> 1], (bin & 1u) ? 0x10000u : 1u);
          }
        }
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h)
          *reinterpret_cast<float4*>(uimg + g_elems + 4 * h) = make_float4(u[4 * h], u[4 * h + 1], u[4 * h + 2], u[4 * h + 3]);
      } else {
#pragma unroll
        for (int h = 0; h < VEC / 4; ++h) {   // replay: this thread wrote these values itself
          const float4 u4 = *reinterpret_cast<const float4*>(uimg + g_elems + 4 * h);
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
    // ---- band bookkeeping, all predicated (the loop has no data-dependent branch): keys under the band are counted, keys inside
    // it go to the fine histogram and to the candidate list.  Keys are < 2^31 and so is LB: key - LB has its top bit set exactly
    // when key < LB.
    uint32_t inband = 0;
    if (valid) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const uint32_t d = __float_as_uint(u[e]) - LB;
        below += d >> 31;
        const uint32_t in = (d < W) ? 1u : 0u;
        inband |= in << e;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
                     ::"r"(hb_addr + ((d >> sh) << 2)), "r"(1u), "r"(in) : "memory");
      }
    }
    // Every in-band element is written provisionally with mask 0 (it lies below UB for `higher`, at or above LB for `lower`); the
    // patch pass only has to overwrite those whose exact compare says 1.  Without the optional outputs the record therefore carries
    // x_(t-1) FOR MASK 1, computed here where all operands are in registers (6 more instructions per element; the streaming loop
    // is memory-bound), and the patch is a compare and a store.  With optional outputs it carries eps and the sample instead.
    float pv1[VEC];
    if constexpr (!OUTS) {
      if (valid) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          float eg1, x01;
          guided_elem(dc, post_M, inv_ah, inv_sa, u[e], e0[e], s[e], Sv[e], 1.0f, eg1, x01, pv1[e]);
        }
      }
    }
    {
      // one predicated shared-memory atomic per lane that has candidates (about a quarter of the lanes), then one predicated 16-byte
      // store per candidate.  The order of the list does not matter.
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
                     ::"r"(cand_addr + slot * 16u), "r"(__float_as_uint(u[e])), "r"(g_elems + e),
                       "r"(__float_as_uint(OUTS ? e0[e] : pv1[e])), "r"(__float_as_uint(s[e])), "r"(take) : "memory");
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
          if (p.x0_out) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.x0_out) + b * p.x0_stride + o) =
              make_float4(x0v[4 * h], x0v[4 * h + 1], x0v[4 * h + 2], x0v[4 * h + 3]);
          if (p.eps_out) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.eps_out) + b * p.eps_out_stride + o) =
              make_float4(eg[4 * h], eg[4 * h + 1], eg[4 * h + 2], eg[4 * h + 3]);
          if (p.mask_out) *reinterpret_cast<float4*>(p.mask_out + b * p.mask_out_stride + o) =
              make_float4(mk[4 * h], mk[4 * h + 1], mk[4 * h + 2], mk[4 * h + 3]);
        }
      }
    }
  };
  // A warp owns a BLOCK of `trips` consecutive rows of 32 groups in its CTA's slice; block (r, w) = warp w of CTA r.
  constexpr int NWARPS = THREADS / 32;
  auto row_group = [&](unsigned r, int w, int j, bool& valid) -> uint32_t {
    const int g = ((w * trips + j) << 5) | lane;              // group index within the slice of CTA r
    valid = g < ngroups;
    return (uint32_t)r * (uint32_t)L + (uint32_t)g * VEC;     // first element, in the image
  };

  // ---------------------------------------------------------------- pilot: trip 0, map values only
  {
    bool v;
    const uint32_t ge = row_group(crank, warp, 0, v);
    do_group(ge, v, true, true);
  }
  sync_all();   // pilot histograms of every CTA are complete
  locate_rank<H0_BINS, THREADS, false, true, true>(cluster, csize, hp, pk.r_lo, misc, pk.r_hi);
  const uint32_t lb_bin = pk.open_low ? 0u : misc[0];
  const uint32_t ub_bin = pk.open_high ? (uint32_t)H0_BINS : misc[1] + 1u;
  __syncthreads();
  LB = lb_bin << LOW;
  W = (ub_bin << LOW) - LB;   // UB = 4096 << 19 = 2^31 when open at the top: above every finite key and every NaN pattern
  sh = ((W - 1u) >> 10) ? (32u - (uint32_t)__clz((int)((W - 1u) >> 10))) : 0u;   // smallest shift with (W - 1) >> sh < 1024
  one_lo = higher ? (LB + W) : 0u;
  one_w = higher ? (0u - (LB + W)) : LB;   // (LB + W = 0 cannot occur: ub_bin >= 1)
  stamp(kp, 1);

  // ---------------------------------------------------------------- stream: every other trip, then the pilot replayed
  // Launched as the programmatic dependent of the du_batch_sum that produces S (S_overlap): everything above ran while that
  // kernel was still summing; its result is complete and visible from here on.  A no-op for an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (kp.late_ns != 0) {
    // test knob (DU_FUSED_JITTER_NS / DU_FUSED_JITTER_SEED): every CTA starts its streaming pass after a pseudo-random delay, so
    // that rows get stolen in ever different patterns; the outputs must not depend on it (tests/test_fused_gpu.py stress test)
    uint32_t h = (uint32_t)(blockIdx.y * gridDim.x + blockIdx.x) * 2654435761u ^ kp.late_from;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const unsigned long long wait_ns = h % kp.late_ns;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    do {
      __nanosleep(200);
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    } while (t - t0 < wait_ns);
  }
  // Row trips are CLAIMED, one at a time, from the block's counter (a shared-memory atomic by lane 0): first the warp's own block,
  // then — HBM arbitration is not fair across SMs, the two CTAs of a cluster finish their own rows microseconds apart, and the
  // finish needs both — the unclaimed rows of the other blocks of the cluster, the other CTAs' first (DSMEM atomics).  Every
  // per-element structure (fine histogram, candidate records, counts) is addressed by image element, so a row may be processed
  // by any CTA of the cluster; histograms and counts are summed over the cluster anyway.
  auto claim = [&](uint32_t* ctr) -> int {
    uint32_t j = 0;
    if (lane == 0) j = atomicAdd(ctr, 1u);
    return (int)__shfl_sync(0xffffffffu, j, 0);
  };
  for (int j = claim(&rowctr[warp]); j < trips;) {
    const int j_next = claim(&rowctr[warp]);    // (claimed one trip ahead: the atomic's latency hides behind this trip's loads)
    bool v;
    const uint32_t ge = row_group(crank, warp, j, v);
    do_group(ge, v, true, false);
    j = j_next;
  }
  {
    bool v;
    const uint32_t ge = row_group(crank, warp, 0, v);
    do_group(ge, v, false, false);
  }
  if (pk.steal) {
    for (unsigned i = 1; i <= csize; ++i) {            // the other CTAs first, this CTA's other warps last
      const unsigned r = (crank + i) % csize;
      uint32_t* vc = peer(rowctr, r);
      for (;;) {
        // one pass over the victim CTA's counters: lane l looks at block (warp + l) % NWARPS
        const int vw = (warp + lane) % NWARPS;
        const uint32_t seen = (lane < NWARPS) ? vc[vw] : 0xffffffffu;
        const uint32_t open = __ballot_sync(0xffffffffu, seen < (uint32_t)trips);
        if (open == 0u) break;
        const int pick = (warp + (__ffs((int)open) - 1)) % NWARPS;
        const int j = claim(vc + pick);
        if (j >= trips) continue;                      // somebody else took the block's last row meanwhile: look again
        bool v;
        const uint32_t ge = row_group(r, pick, j, v);
        do_group(ge, v, true, false);
      }
    }
  }
  if (__any_sync(0xffffffffu, nan_seen) && lane == 0) misc[3] = 1u;
  below = __reduce_add_sync(0xffffffffu, below);
  if (lane == 0) atomicAdd(&misc[46], below);
  stamp(kp, 2);

  // ---------------------------------------------------------------- finish: verify the band, exact select, patch
  sync_all();   // fine histograms, counts, candidate lists and flags of every CTA are complete; nobody reads a pilot histogram any more
  bool bad = false;
  uint32_t below_tot = 0, in_tot = 0;
  for (unsigned r = 0; r < csize; ++r) {
    const uint32_t* pm = peer(misc, r);
    bad |= (pm[3] != 0u) | (pm[45] != 0u);
    below_tot += pm[46];
    in_tot += pm[44];
  }
  // rank lo must fall inside the band (this is the verification of the prediction)
  bad |= !(kp.lo >= below_tot && kp.lo - below_tot < in_tot);
  stamp(kp, 3);
  float thr = 0.0f;
  bool redo = false;
  uint32_t ncand = 0;
  if (!bad) {
    ncand = misc[44];
    const bool need_next = kp.hi > kp.lo;
    bad = !band_select<THREADS>(cluster, csize, crank, kp, hp, cand, hb, misc, LB, sh, kp.lo - below_tot);
    if (bad) {
      if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // (band_select left an arrival pending)
    } else {
      const uint32_t key_lo = misc[41];
      uint32_t key_hi = misc[43];
      if (need_next && key_hi == 0xffffffffu) {
        // the successor lies above the band: smallest key above key_lo, one pass over the map in its slot, cluster-wide.  Rows of
        // this slice may have been written by another CTA of the cluster (work stealing): fence - barrier - fence orders them.
        if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        __threadfence();
        sync_all();
        __threadfence();
        uint32_t best = 0xffffffffu;
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
      thr = lerp_torch(__uint_as_float(key_lo), __uint_as_float(key_hi), kp.w, p.lerp_fma);
    }
  }
  stamp(kp, 5);
  if (bad) {
    // exact select over the map in its slot (L2) and a full update pass (`bad` is identical in every CTA of the cluster).  The
    // level-0 histogram the select starts from is rebuilt here; peers are past their DSMEM reads of this CTA's areas (barrier),
    // and the map / provisional x_(t-1) values another CTA wrote into this slice (work stealing) are ordered by fence - barrier - fence.
    __threadfence();
    sync_all();
    __threadfence();
    for (int j = tid; j < H0_WORDS + PRED_WORK_WORDS; j += THREADS) hp[j] = 0;
    if (tid == 0) { misc[4] = 0xffffffffu; misc[5] = 0u; misc[6] = 0u; misc[40] = 0u; }
    __syncthreads();
    const int ng4 = (int)(L / 4);
    for (int g = tid; g < ng4; g += THREADS) {
      const uint4 v = *reinterpret_cast<const uint4*>(urow + 4 * g);
      const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t bin = kk[e] >> LOW;
        atomicAdd(&hp[bin >> 1], (bin & 1u) ? 0x10000u : 1u);
      }
    }
    thr = select_threshold<THREADS>(cluster, csize, crank, kp, urow, hp, work, misc);
    redo = true;
  }
  if (!redo) {
    // ---- patch the candidates with the exact compare (all inputs come from the shared-memory records)
    if constexpr (!OUTS) {
      for (uint32_t i = tid; i < ncand; i += THREADS) {
        const uint4 rec = cand[i];
        const float u = __uint_as_float(rec.x);
        if (higher ? (u > thr) : (u < thr)) prow[rec.y] = __uint_as_float(rec.z);
      }
    } else {
      for (uint32_t i0 = tid; i0 < ncand; i0 += 4 * THREADS) {
        uint4 rec[4];
        float Sv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {   // records and S values first: four independent L2 / shared-memory reads in flight per thread
          const uint32_t i = i0 + q4 * THREADS;
          rec[q4] = (i < ncand) ? cand[i] : make_uint4(0u, 0u, 0u, 0u);
          Sv[q4] = (has_S && i < ncand) ? ld_coherent_f1(Srow + rec[q4].y) : __uint_as_float(rec[q4].z);   // (the S row is shared by the batch: L2 / L1 resident)
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (i0 + q4 * THREADS < ncand) {
            const float u = __uint_as_float(rec[q4].x);
            const int64_t o = rec[q4].y;
            const float mk = (higher ? (u > thr) : (u < thr)) ? 1.0f : 0.0f;
            float eg, x0, pv;
            guided_elem(dc, post_M, inv_ah, inv_sa, u, __uint_as_float(rec[q4].z), __uint_as_float(rec[q4].w), Sv[q4], mk, eg, x0, pv);
            prow[o] = pv;
            if (p.x0_out) reinterpret_cast<float*>(p.x0_out)[b * p.x0_stride + o] = x0;
            if (p.eps_out) reinterpret_cast<float*>(p.eps_out)[b * p.eps_out_stride + o] = eg;
            if (p.mask_out) p.mask_out[b * p.mask_out_stride + o] = mk;
          }
        }
      }
    }
  } else {
    guided_update_slice<T, THREADS, true, false>(kp, urow, thr, b, base, 0u);
  }
  if (crank == 0 && tid == 0 && p.thr_out) p.thr_out[b] = thr;
  stamp(kp, 6);
  if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}


static constexpr size_t kPredFixedBytes = (size_t)(H0_WORDS + PRED_WORK_WORDS + MISC_WORDS + ROWCTR_WORDS) * 4;

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
    case 1024: return launch_pred_t<T, MT, 1024, 1, OUTS>(pk, plan, smem, st);
#ifdef DU_PRED_TUNING   // 384 / 768-thread CTAs (80 registers): measured slower, kept for sweeps only (csrc/build.sh -DDU_PRED_TUNING)
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

// One attempt with a given launch shape: returns 1 if launched, 0 if the shape is not eligible, < 0 on error.
static int try_pred(const FusedKParams& kp_in, int cluster, int threads, int64_t min_trips, cudaStream_t st) {
  const du_fused_params& p = kp_in.p;
  const bool outs_req = p.x0_out || p.eps_out || p.mask_out;
  const bool mt_known = outs_req ? (p.M == 5) : (p.M == 4 || p.M == 5 || p.M == 8 || p.M == 16);   // mirrors launch_pred
  // elements per thread and trip: 4, except 16-bit scores with a runtime-M instance (16-byte vectors of 8)
  const int vec = (p.score_dtype == DU_F32 || mt_known) ? 4 : 8;
  if (p.n % ((int64_t)cluster * vec) != 0) return 0;
  const int64_t L = p.n / cluster, ngroups = L / vec;
  if (L >= 65536) return 0;   // packed 16-bit level-0 counters (pilot histogram, fallback)
  const int64_t trips = (ngroups + threads - 1) / threads;
  if (trips < min_trips) return 0;
  // pilot sample: trip 0 of every warp = row (warp * trips) of 32 groups
  int64_t pilot_groups = 0;
  for (int w = 0; w < threads / 32; ++w) {
    const int64_t left = ngroups - (int64_t)w * trips * 32;
    pilot_groups += left <= 0 ? 0 : (left < 32 ? left : 32);
  }
  const double n_p = (double)pilot_groups * vec * cluster;
  if (n_p < 256) return 0;
  // DU_FUSED_BAND_SIGMA=<x>: half-width of the rank window in standard deviations of the pilot quantile.  Default 5: the band is
  // missed by one image in ~2 million (it then takes the exact fallback, +20 us for that launch); 6 costs 0.4 us per launch, 4 gains 0.4.
  const char* e_s = getenv("DU_FUSED_BAND_SIGMA");
  const double nsig = (e_s && atof(e_s) > 0.0) ? atof(e_s) : 5.0;
  const double q = (double)p.q;
  const double rank_p = q * (n_p - 1.0);
  const double delta = nsig * std::sqrt(n_p * q * (1.0 - q)) + 4.0;
  PredKParams pk;
  pk.k = kp_in;
  pk.k.L = L;
  pk.k.tmem_cols = 0; pk.k.tmem_cpg = 0;
  pk.k.late_ns = 0; pk.k.late_from = 0;
  if (const char* e_j = getenv("DU_FUSED_JITTER_NS")) {
    pk.k.late_ns = (uint32_t)atoi(e_j);
    const char* e_js = getenv("DU_FUSED_JITTER_SEED");
    pk.k.late_from = e_js ? (uint32_t)atoi(e_js) * 747796405u + 1u : 1u;
  }
  pk.trips = (uint32_t)trips;
  pk.open_low = (rank_p - delta <= 0.0) ? 1u : 0u;
  pk.open_high = (rank_p + delta >= n_p - 1.0) ? 1u : 0u;
  pk.r_lo = pk.open_low ? 0u : (uint32_t)std::floor(rank_p - delta);
  pk.r_hi = pk.open_high ? (uint32_t)(n_p - 1.0) : (uint32_t)std::ceil(rank_p + delta);
  // candidate list: the band holds about 2 * delta / n_p of the elements plus two level-0 bins; room for 1.6x that, and
  // no more: global loads in flight are buffered in L1, which shares the SM's 256 KB with shared memory, so a large
  // shared-memory footprint throttles the streaming pass (measured: 2 x 116 KB per SM cost 10 us against 2 x 78 KB, and 2 x 48 KB
  // — a list that overflows — 15 us).  DU_FUSED_SMEM_KB caps the per-CTA footprint (default 64, twice that for one CTA per SM).
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
    const char* e_r = getenv("DU_FUSED_STEAL");
    pk.steal = (e_r && atoi(e_r) == 0) ? 0u : 1u;
  }
  {
    // DU_L2_HINTS=0: every load at normal priority (A/B of the eviction hints)
    const char* e_h = getenv("DU_L2_HINTS");
    pk.pol_stream = (e_h && atoi(e_h) == 0) ? kL2EvictNormal : kL2EvictFirst;
  }
  FusedPlan plan{};
  plan.cluster = cluster; plan.threads = threads; plan.minb = one_cta ? 1 : 2;
  const size_t smem = kPredFixedBytes + (size_t)cap * 16;
  switch (p.score_dtype) {
    case DU_F32: return launch_pred<float>(pk, plan, threads, smem, st);
    case DU_F16: return launch_pred<__half>(pk, plan, threads, smem, st);
    default: return launch_pred<__nv_bfloat16>(pk, plan, threads, smem, st);
  }
}

int launch_fused_pred(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st) {
  const du_fused_params& p = kp.p;
  // DU_FUSED_PRED=0 forces the three-phase kernel
  const char* e_p = getenv("DU_FUSED_PRED");
  if (e_p && atoi(e_p) == 0) return 0;
  const bool fast_c = p.ddim.prediction_type == DU_PRED_EPSILON && !p.ddim.use_clipped_model_output &&
                      p.sample_dtype == DU_F32 && p.prev_dtype == DU_F32 && !p.skip_ddim;
  if (!fast_c) return 0;
  // The pilot (~3.5 us) and the latency-bound finish (~6 us) are fixed costs: below ~6 trips per thread the three-phase kernel is
  // faster (ImageNet-64 b128, 6 trips: 17.9 us against 18.7 us; U-ViT latents, 4 trips at most: the three-phase kernel)
  const char* e_m = getenv("DU_FUSED_PRED_MIN_TRIPS");
  const int64_t min_trips = (e_m && atoi(e_m) >= 4) ? atoi(e_m) : 6;
  // Launch shape.  The kernel keeps no copy of the map in shared memory, so it does not need the three-phase kernel's cluster
  // plan: ONE CTA of 1024 threads per image (no cluster barrier, no DSMEM, no wait for a peer CTA on a busier SM) is the
  // fastest shape measured at batch 128 (ImageNet-128 step: 44.6 us against 47.4 us for clusters of two 512-thread CTAs).
  // DU_FUSED_CLUSTER / DU_FUSED_THREADS / DU_FUSED_PRED_THREADS pin the shape (tests, sweeps).
  const char* e_c = getenv("DU_FUSED_CLUSTER");
  const char* e_t = getenv("DU_FUSED_THREADS");
  const char* e_pt = getenv("DU_FUSED_PRED_THREADS");
  const bool pinned = (e_c && atoi(e_c) > 0) || (e_t && atoi(e_t) > 0) || (e_pt && atoi(e_pt) > 0);
  if (pinned) {
    int threads = plan.threads;
    if (e_pt) {
      const int t = atoi(e_pt);
      if (t == 384 || t == 512 || t == 768 || t == 1024) threads = t;
    }
    if (threads != 384 && threads != 512 && threads != 768 && threads != 1024) return 0;
    return try_pred(kp, plan.cluster, threads, min_trips, st);
  }
  // Few images (a batch sharded over the GPUs of a box: 64 / 32 / 16 images per GPU): one CTA per image would leave most of the
  // 148 SMs idle, so the image is spread over a cluster of 2 / 4 / 8 CTAs of 512 threads — the smallest cluster that yields at
  // least ~100 CTAs while every thread keeps min_trips row trips; if none does, the largest eligible cluster.
  int tried = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int c = (pass == 0 ? 1 : 8); pass == 0 ? c <= 8 : c >= 1; c = (pass == 0 ? c * 2 : c / 2)) {
      if (pass == 0 && p.B * c < 100) continue;
      // (clusters of 8 with 3 trips per thread were measured too: 27.5 us against 19.6 us for 4 x 6 trips at 16 images — the
      // 8-way DSMEM histogram sums cost more than the shorter streaming pass saves)
      const int rc = try_pred(kp, c, c == 1 ? 1024 : 512, min_trips, st);
      ++tried;
      if (rc != 0) return rc;
    }
  }
  if (plan.threads == 512 || plan.threads == 1024) {
    const int rc = try_pred(kp, plan.cluster, plan.threads, min_trips, st);
    if (rc != 0) return rc;
  }
  (void)tried;
  return 0;
}

}  // namespace du
