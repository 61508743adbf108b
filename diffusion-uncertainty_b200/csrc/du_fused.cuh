// du_fused.cuh — device code shared by the two fused-step kernels (du_fused.cu: three-phase kernel and the C-ABI entry;
// du_fused_pred.cu: predictive single-pass kernel): histogram layout, rank location, the exact list select, the
// shared-memory / L2 threshold select and the guided update of a slice.
#pragma once
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "du_common.cuh"

namespace cg = cooperative_groups;

namespace du {

constexpr int H0_BITS = 12, H1_BITS = 10, H2_BITS = 9;  // 31 value bits (sign is always 0)
constexpr int H0_BINS = 1 << H0_BITS, H1_BINS = 1 << H1_BITS, H2_BINS = 1 << H2_BITS;
// Level 0 is counted per CTA in 16-bit halves of 32-bit words (a CTA slice that fits shared memory has < 65536 elements),
// so 4096 bins take the 8 KB that 2048 32-bit bins would.
constexpr int H0_WORDS = H0_BINS / 2;
// work area behind the level-0 histogram: [0, LIST_CAP) candidate keys of the selected level-0 bin, [LIST_CAP, +H1_BINS) their
// level-1 histogram.  The general path (heavy ties) reuses the area as level-1 / level-2 histograms of the whole slice.
constexpr int LIST_CAP = 1024;   // (the predictive kernel keeps a larger list: PRED_LIST_CAP in du_fused_pred.cu)
constexpr int WORK_WORDS = LIST_CAP + H1_BINS;
constexpr int HIST_WORDS = H0_WORDS + WORK_WORDS;
// misc words: [0..2] locate result, [3] nan flag, [4] min larger key, [5] next bin, [6] candidate count, [7] threshold,
// [8..39] warp sums, [40] tiny-list count, [42] tensor-memory base
constexpr int MISC_WORDS = 48;
constexpr int MAX_CLUSTER = 8;

struct FusedKParams {
  du_fused_params p;
  int64_t L;        // elements per CTA slice (n / cluster size)
  uint32_t lo, hi;  // ranks
  float w;          // lerp weight
  float inv_cnt, inv_cm1, inv_sqrt_alpha_t;
  uint32_t tmem_cols;   // tensor-memory columns this CTA allocates for its eps slice (0 = eps is re-read through L2)
  uint32_t tmem_cpg;    // columns per group of 4 warps (= trips * 4)
  uint32_t late_from;   // CTAs whose linear index is >= late_from start their streaming phase late_ns later (0 = off)
  uint32_t late_ns;
  unsigned long long* timeline;  // debug (DU_FUSED_TIMELINE): [CTA][8] globaltimer stamps at the phase boundaries, else null
};

constexpr int TL_SLOTS = 16;  // debug timeline: globaltimer stamps per CTA
__device__ __forceinline__ void stamp(const FusedKParams& kp, int slot) {
  if (kp.timeline && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    kp.timeline[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TL_SLOTS + slot] = t;
    if (slot == 0) {   // last slot: the SM this CTA runs on (which CTAs share an SM explains the spread of the streaming times)
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      kp.timeline[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * TL_SLOTS + TL_SLOTS - 1] = sm;
    }
  }
}

// ---- phase A: the map value of one group of VEC elements --------------------------------------------------------------
// centre_mode: 0 variance over the M scores, 1 second moment about the centre, 2 variance over scores + centre.
// MT > 0: M known at compile time (all loads of the group in flight together); MT == 0: batched runtime-M loop.
template <typename T, int MT>
__device__ __forceinline__ void moments_group(const du_fused_params& p, int64_t srow, int64_t erow, uint32_t g_elems, int centre_mode,
                                              float inv_cnt, float inv_cm1, float (&u)[Vec16<T>::VEC], uint4& raw_c) {
  using V = Vec16<T>;
  constexpr int VEC = V::VEC;
  float c[VEC], k[VEC], s1[VEC], s2[VEC];
  raw_c = make_uint4(0u, 0u, 0u, 0u);
  const uint32_t byte_off = g_elems * (uint32_t)sizeof(T);
  if (centre_mode) raw_c = ldg_stream_128_at(reinterpret_cast<const T*>(p.eps) + erow, byte_off);
  // the scores are read exactly once: L2 evict-first, so that they do not displace eps (kept by du_batch_sum) and the outputs
  if constexpr (MT > 0) accumulate_scores_ct<T, MT, true>(p.scores, srow, byte_off, raw_c, centre_mode, false, centre_mode != 1, c, k, s1, s2, kL2EvictFirst);
  else accumulate_scores<T>(p.scores, p.M, srow + g_elems, raw_c, centre_mode, false, centre_mode != 1, c, k, s1, s2);
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    // + 0.0f: never -0.0, so the bit pattern of the map orders like its value
    u[e] = (centre_mode == 1) ? fmaf(s2[e], inv_cnt, 0.0f) : fmaf(m2_from_sums(s1[e], s2[e], inv_cnt), inv_cm1, 0.0f);
  }
}

// ---- phase B: block-wide search of rank k in the histogram summed over the cluster's CTAs ----------------------------
// hist_local points at this CTA's histogram for the level; peers are reached through DSMEM.  Result in misc[0..2]
// (bin, count below the bin, count in the bin); for LAST also misc[5] = next non-empty bin above (or NBINS).
// PACKED16: two 16-bit counters per word (bin 2i in the low half).
// TWO: `kb` is a second rank (kb >= k); only the two BINS are reported: misc[0] = bin of k, misc[1] = bin of kb.
template <int NBINS, int THREADS, bool LAST, bool PACKED16, bool TWO = false>
__device__ __forceinline__ void locate_rank(cg::cluster_group& cluster, unsigned csize, uint32_t* hist_local, uint32_t k,
                                            uint32_t* misc, uint32_t kb = 0) {
  // bins per thread; packed histograms need an even number.  THREADS need not divide NBINS (384 / 768-thread CTAs): the
  // last active thread's range is cut at NBINS (EXACT == false adds the guards).
  constexpr int PER0 = (NBINS + THREADS - 1) / THREADS;
  constexpr int PER = PACKED16 ? ((PER0 + 1) / 2) * 2 : PER0;
  constexpr bool EXACT = (NBINS % PER == 0) && (NBINS / PER <= THREADS) && ((NBINS / PER) * PER == NBINS);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t c[PER], tot = 0;
  const int first = tid * PER;
#pragma unroll
  for (int j = 0; j < PER; ++j) c[j] = 0;
  if (first < NBINS) {
    for (unsigned r = 0; r < csize; ++r) {
      const uint32_t* h = (csize > 1) ? cluster.map_shared_rank(hist_local, r) : hist_local;
      if constexpr (PACKED16) {
        constexpr int W = PER / 2;
        if constexpr (W % 4 == 0 && EXACT) {
#pragma unroll
          for (int j = 0; j < W; j += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(h + first / 2 + j);
            c[2 * j] += q.x & 0xffffu; c[2 * j + 1] += q.x >> 16; c[2 * j + 2] += q.y & 0xffffu; c[2 * j + 3] += q.y >> 16;
            c[2 * j + 4] += q.z & 0xffffu; c[2 * j + 5] += q.z >> 16; c[2 * j + 6] += q.w & 0xffffu; c[2 * j + 7] += q.w >> 16;
          }
        } else if constexpr (W % 2 == 0 && EXACT) {
#pragma unroll
          for (int j = 0; j < W; j += 2) {
            const uint2 q = *reinterpret_cast<const uint2*>(h + first / 2 + j);
            c[2 * j] += q.x & 0xffffu; c[2 * j + 1] += q.x >> 16; c[2 * j + 2] += q.y & 0xffffu; c[2 * j + 3] += q.y >> 16;
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) {
            if (EXACT || first / 2 + j < NBINS / 2) { const uint32_t q = h[first / 2 + j]; c[2 * j] += q & 0xffffu; c[2 * j + 1] += q >> 16; }
          }
        }
      } else if constexpr (PER % 4 == 0 && EXACT) {   // one 16-byte (DSMEM) load per 4 bins
#pragma unroll
        for (int j = 0; j < PER; j += 4) {
          const uint4 q = *reinterpret_cast<const uint4*>(h + first + j);
          c[j] += q.x; c[j + 1] += q.y; c[j + 2] += q.z; c[j + 3] += q.w;
        }
      } else if constexpr (PER % 2 == 0 && EXACT) {
#pragma unroll
        for (int j = 0; j < PER; j += 2) {
          const uint2 q = *reinterpret_cast<const uint2*>(h + first + j);
          c[j] += q.x; c[j + 1] += q.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < PER; ++j)
          if (EXACT || first + j < NBINS) c[j] += h[first + j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) tot += c[j];
  uint32_t incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  uint32_t* wsum = misc + 8;
  if (lane == 31) wsum[warp] = incl;
  if (LAST && tid == 0) misc[5] = NBINS;
  __syncthreads();
  // exclusive prefix over the warp totals: lane l reads warp l's total, one shuffle scan per warp
  uint32_t wtot = (lane < THREADS / 32) ? wsum[lane] : 0u, wincl = wtot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
    if (lane >= o) wincl += t;
  }
  const uint32_t wprefix = __shfl_sync(0xffffffffu, wincl - wtot, warp);
  const uint32_t excl = wprefix + incl - tot;
  if (k >= excl && k < excl + tot) {
    uint32_t cum = excl;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (k < cum + c[j]) { misc[0] = first + j; if (!TWO) { misc[1] = cum; misc[2] = c[j]; } break; }
      cum += c[j];
    }
  }
  if (TWO && kb >= excl && kb < excl + tot) {
    uint32_t cum = excl;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (kb < cum + c[j]) { misc[1] = first + j; break; }
      cum += c[j];
    }
  }
  __syncthreads();
  if (LAST) {
    const uint32_t sel = misc[0];
    uint32_t nb = NBINS;
#pragma unroll
    for (int j = PER - 1; j >= 0; --j)
      if (c[j] != 0 && (uint32_t)(first + j) > sel) nb = first + j;
    nb = __reduce_min_sync(0xffffffffu, nb);
    if (lane == 0 && nb < (uint32_t)NBINS) atomicMin(&misc[5], nb);
    __syncthreads();
  }
}

__device__ __forceinline__ float rcp_fast(float x) {  // MUFU.RCP: <= 1 ulp, rcp(0) = inf, rcp(inf) = 0
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) {  // torch.clamp: NaN in -> NaN out
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
}
// hist[bin]++ iff a == b, as ONE predicated instruction (the compiler turns `if (..) atomicAdd` into a divergent branch
// per element, which made the select passes issue-bound)
__device__ __forceinline__ void hist_inc_if_eq(uint32_t hist_smem_addr, uint32_t bin, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, %3;\n\t@p red.shared.add.u32 [%0], %1;\n\t}"
               ::"r"(hist_smem_addr + bin * 4u), "r"(1u), "r"(a), "r"(b) : "memory");
}

// cluster-wide barrier with release/acquire on shared::cluster (what the DSMEM histogram exchange needs)
// Only shared-memory histograms cross CTAs here.  They are complete in the owning SM's shared memory once the CTA
// barrier in front has been passed, so the cluster barrier itself is relaxed: the release form costs a MEMBAR.ALL.GPU
// (it waits for every outstanding global store of phase A).
// (-DDU_CLUSTER_BARRIER_RELEASE builds the formally release / acquire form for A/B runs and tool checks.)
__device__ __forceinline__ void cluster_barrier() {
  __syncthreads();
#ifdef DU_CLUSTER_BARRIER_RELEASE
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}

// ---- tensor memory (TMEM, 256 KB per SM, otherwise idle in this tensor-core-free kernel) as a per-thread stash: phase A
// parks the raw 16-byte eps vector of every group there (tcgen05.st), phase C takes it back (tcgen05.ld) instead of
// re-reading eps through L2.  A warp reaches the 32 lanes (warp % 4) * 32.. of the columns it addresses; the four
// warp groups of a CTA use disjoint column ranges.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 tmem_ld4(uint32_t taddr) {
  uint4 v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return v;
}
__device__ __forceinline__ uint32_t tmem_slot(uint32_t tbase, uint32_t cpg, int trip) {
  const uint32_t warp = threadIdx.x >> 5;
  return tbase + (((warp & 3u) * 32u) << 16) + (warp >> 2) * cpg + (uint32_t)trip * 4u;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- the exact select on the candidate lists.  Precondition: every CTA of the cluster has compacted the keys of level-0
// bin d0 (`want` = d0 << 19) into `list` (count in misc[6]) with their level-1 histogram in `h1`, and a cluster barrier has
// been passed.  Every CTA locates the level-1 bin in the summed histograms, picks the handful of keys that share the 22-bit
// prefix out of the cluster's lists (DSMEM reads, no copy) and warp 0 finishes on those few keys.  Results: misc[41] =
// key_lo, misc[43] = key_hi (0xffffffff: the successor is not in level-0 bin d0).  Leaves ONE cluster-barrier arrival
// pending (issued after the last DSMEM read).  `tiny` must be free scratch of at least LIST_CAP words (the level-0
// histogram, which nobody reads any more).
template <int THREADS>
__device__ __forceinline__ void list_select(cg::cluster_group& cluster, unsigned csize, unsigned crank, const FusedKParams& kp,
                                            uint32_t* tiny, uint32_t* list, uint32_t* h1, uint32_t* misc, uint32_t want,
                                            uint32_t below0, bool& has_nan) {
  const int tid = threadIdx.x, lane = tid & 31;
  auto peer = [&](uint32_t* ptr, unsigned r) -> uint32_t* { return (csize > 1) ? cluster.map_shared_rank(ptr, r) : ptr; };
  {
    locate_rank<H1_BINS, THREADS, true, false>(cluster, csize, h1, kp.lo - below0, misc);
    stamp(kp, 7);
    const uint32_t d1 = misc[0], below1 = misc[1], cnt1 = misc[2], next1 = misc[5];
    const uint32_t prefix = want | (d1 << H2_BITS);            // 22 known bits of key_lo
    const uint32_t k2 = kp.lo - below0 - below1;               // rank of key_lo among the cnt1 keys with that prefix
    const bool need_next = kp.hi > kp.lo;                      // the upper statistic is the next key in sorted order
    // ---- pick the keys with the prefix (and, if the successor lives in the next non-empty level-1 bin, the smallest key
    // there) out of the cluster's lists.  h0 is free now: the tiny list goes to its first words.
    has_nan = false;
    {
      const uint32_t next_prefix = want | (next1 << H2_BITS);
      const bool want_next_bin = need_next && (k2 + 1 >= cnt1) && next1 < (uint32_t)H1_BINS;
      uint32_t best = 0xffffffffu;
      for (unsigned i = 0; i < csize; ++i) {
        const unsigned r = (crank + i) % csize;
        const uint32_t* pm = peer(misc, r);
        const uint32_t* pl = peer(list, r);
        const uint32_t cr = pm[6];
        has_nan |= (pm[3] != 0);
        for (uint32_t j = tid; j < cr; j += THREADS) {
          const uint32_t key = pl[j];
          const uint32_t hi22 = key & ~(uint32_t)(H2_BINS - 1);
          if (hi22 == prefix) tiny[atomicAdd(&misc[40], 1u)] = key;
          else if (want_next_bin && hi22 == next_prefix) best = min(best, key);
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
    stamp(kp, 9);
    if (tid < 32) {
      // warp 0: k2-th smallest (and its successor) of the cnt1 keys in `tiny`.  They share 22 bits and are normally a
      // handful: one key per lane, rank = number of smaller keys (ties broken by position) from m independent
      // shuffle / ballot rounds.  A serial single-warp search costs ~5 cycles per dependent instruction, this does not.
      const uint32_t m = misc[40];   // == cnt1
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
        // many keys with one 22-bit prefix (ties): bitwise search over the low 9 bits
        uint32_t ans = 0;
#pragma unroll 1
        for (int b = H2_BITS - 1; b >= 0; --b) {
          const uint32_t trial = prefix | ans | (1u << b);
          uint32_t c = 0;
          for (uint32_t j = lane; j < m; j += 32) c += (tiny[j] < trial);
          c = __reduce_add_sync(0xffffffffu, c);
          if (c <= k2) ans |= (1u << b);
        }
        key_lo = prefix | ans;
        key_hi = key_lo;
        if (need_next) {
          // successor: a tie, else the smallest larger key with the same prefix, else the smallest key of the next level-1 bin
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
          else key_hi = misc[4];   // 0xffffffff: the successor lives in a higher level-0 bin
        }
      }
      if (lane == 0) { misc[41] = key_lo; misc[43] = key_hi; }
    }
    stamp(kp, 10);
    __syncthreads();
  }
}

// ---- phase B: the per-image threshold ------------------------------------------------------------------------------------
// Level 0 (top 12 value bits) was histogrammed on the fly in phase A.  Fast path: ONE pass over the shared-memory map
// compacts the keys of the selected level-0 bin (typically ~1 % of the image) into a per-CTA candidate list and histograms
// their next 10 bits; after one cluster barrier every CTA locates the level-1 bin in the summed histograms, picks the
// handful of candidates that share the 22-bit prefix out of the cluster's lists (DSMEM reads, no copy) and warp 0 finishes
// the exact select on those few keys.  Every CTA computes the same threshold from the same data: no publish step.
// Heavy ties (more than LIST_CAP elements in the level-0 bin) take the general path: two more full histogram passes with
// one cluster barrier each.
template <int THREADS>
__device__ __forceinline__ float select_threshold(cg::cluster_group& cluster, unsigned csize, unsigned crank, const FusedKParams& kp,
                                                  const float* u_s, uint32_t* h0, uint32_t* work, uint32_t* misc) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int ng4 = (int)(kp.L / 4);
  constexpr int LOW = H1_BITS + H2_BITS;
  auto sync_all = [&]() { if (csize > 1) cluster_barrier(); else __syncthreads(); };
  auto peer = [&](uint32_t* ptr, unsigned r) -> uint32_t* { return (csize > 1) ? cluster.map_shared_rank(ptr, r) : ptr; };

  sync_all();  // level-0 histograms of every CTA are complete
  stamp(kp, 2);
  locate_rank<H0_BINS, THREADS, false, true>(cluster, csize, h0, kp.lo, misc);
  const uint32_t d0 = misc[0], below0 = misc[1], cnt0 = misc[2];
  __syncthreads();
  stamp(kp, 3);
  const uint32_t want = d0 << LOW, msk0 = (uint32_t)(H0_BINS - 1) << LOW;
  float thr;

  if (cnt0 <= (uint32_t)LIST_CAP) {
    uint32_t* list = work;
    uint32_t* h1 = work + LIST_CAP;   // zeroed at kernel start
    {
      // ---- compaction.  The pass is issue-bound (every element examined), so the common case is kept to ~3 instructions
      // per element: a trip only records ONE bit, "some of my 4 elements are in the selected bin".  The few flagged trips
      // (about one per thread) are re-read afterwards; list slots come from one shared-memory atomic per warp (the order of
      // the list does not matter).
      const int trips = (ng4 + THREADS - 1) / THREADS;
      const uint32_t* ub = reinterpret_cast<const uint32_t*>(u_s);
      for (int it0 = 0; it0 < trips; it0 += 32) {
        uint32_t flagged = 0;
        const int nj = min(32, trips - it0);
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
          const int g = (it0 + j) * THREADS + tid;
          if (g < ng4) {
            const uint4 v = *reinterpret_cast<const uint4*>(ub + 4 * g);
            const uint32_t t0 = (v.x ^ want) & msk0, t1 = (v.y ^ want) & msk0, t2 = (v.z ^ want) & msk0, t3 = (v.w ^ want) & msk0;
            flagged |= (min(min(t0, t1), min(t2, t3)) == 0u ? 1u : 0u) << j;
          }
        }
        if (!__any_sync(0xffffffffu, flagged != 0u)) continue;
        uint32_t cnt = 0;   // exact count of this thread's candidates (flagged trips only)
        for (uint32_t f = flagged; f; f &= f - 1) {
          const int g = (it0 + __ffs((int)f) - 1) * THREADS + tid;
          const uint4 v = *reinterpret_cast<const uint4*>(ub + 4 * g);
          cnt += ((v.x & msk0) == want) + ((v.y & msk0) == want) + ((v.z & msk0) == want) + ((v.w & msk0) == want);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const uint32_t wtotal = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t wbase = 0;
        if (lane == 0) wbase = atomicAdd(&misc[6], wtotal);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        uint32_t slot = wbase + incl - cnt;   // < cnt0 <= LIST_CAP: the bin holds cnt0 elements cluster-wide
        for (uint32_t f = flagged; f; f &= f - 1) {
          const int g = (it0 + __ffs((int)f) - 1) * THREADS + tid;
          const uint4 v = *reinterpret_cast<const uint4*>(ub + 4 * g);
          const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if ((kk[e] & msk0) == want) {
              list[slot++] = kk[e];
              atomicAdd(&h1[(kk[e] >> H2_BITS) & (H1_BINS - 1)], 1u);
            }
          }
        }
      }
    }
    sync_all();  // candidate lists and their level-1 histograms are complete
    stamp(kp, 4);
    bool has_nan = false;
    const bool need_next = kp.hi > kp.lo;
    list_select<THREADS>(cluster, csize, crank, kp, h0, list, h1, misc, want, below0, has_nan);
    const uint32_t key_lo = misc[41];
    uint32_t key_hi = misc[43];
    if (need_next && key_hi == 0xffffffffu && !has_nan) {
      // rare: key_lo is the largest key of its level-0 bin -> one pass for the smallest key above key_lo, cluster-wide
      // (the condition is identical in every CTA of the cluster: all of them read the same lists)
      if (csize > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
      uint32_t best = 0xffffffffu;
      for (int g = tid; g < ng4; g += THREADS) {
        const uint4 v = *reinterpret_cast<const uint4*>(u_s + 4 * g);
        const uint32_t kk[4] = {v.x & 0x7fffffffu, v.y & 0x7fffffffu, v.z & 0x7fffffffu, v.w & 0x7fffffffu};
#pragma unroll
        for (int e = 0; e < 4; ++e) best = min(best, (kk[e] > key_lo) ? kk[e] : 0xffffffffu);
      }
      best = __reduce_min_sync(0xffffffffu, best);
      if (lane == 0 && best != 0xffffffffu) atomicMin(&misc[4], best);
      sync_all();
      for (unsigned r = 0; r < csize; ++r) key_hi = min(key_hi, peer(misc, r)[4]);
      if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    }
    thr = lerp_torch(__uint_as_float(key_lo), __uint_as_float(key_hi), kp.w, kp.p.lerp_fma);
    if (has_nan) thr = __int_as_float(0x7fc00000);
    return thr;   // one cluster-barrier arrival is pending; the kernel waits for it before it exits
  } else {
    // ---- general path: two more histogram levels over the whole slice
    uint32_t* h1 = work;
    uint32_t* h2 = work + H1_BINS;
    uint32_t k_rank = kp.lo - below0, below = below0;
    {
      const uint32_t h1a = (uint32_t)__cvta_generic_to_shared(h1);
#pragma unroll 4
      for (int g = tid; g < ng4; g += THREADS) {
        const uint4 v = *reinterpret_cast<const uint4*>(u_s + 4 * g);
        hist_inc_if_eq(h1a, (v.x >> H2_BITS) & (H1_BINS - 1), v.x & msk0, want);
        hist_inc_if_eq(h1a, (v.y >> H2_BITS) & (H1_BINS - 1), v.y & msk0, want);
        hist_inc_if_eq(h1a, (v.z >> H2_BITS) & (H1_BINS - 1), v.z & msk0, want);
        hist_inc_if_eq(h1a, (v.w >> H2_BITS) & (H1_BINS - 1), v.w & msk0, want);
      }
    }
    sync_all();
    locate_rank<H1_BINS, THREADS, false, false>(cluster, csize, h1, k_rank, misc);
    const uint32_t d1 = misc[0];
    k_rank -= misc[1];
    below += misc[1];
    __syncthreads();
    const uint32_t prefix21 = want | (d1 << H2_BITS);
    {
      const uint32_t msk = 0x7fffffffu & ~(uint32_t)(H2_BINS - 1), top = prefix21 | (H2_BINS - 1);
      const uint32_t h2a = (uint32_t)__cvta_generic_to_shared(h2);
      uint32_t best = 0xffffffffu;
#pragma unroll 4
      for (int g = tid; g < ng4; g += THREADS) {
        const uint4 v = *reinterpret_cast<const uint4*>(u_s + 4 * g);
        const uint32_t kk[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t key = kk[e] & 0x7fffffffu;
          hist_inc_if_eq(h2a, key & (H2_BINS - 1), key & msk, prefix21);
          best = min(best, (key > top) ? key : 0xffffffffu);
        }
      }
      best = __reduce_min_sync(0xffffffffu, best);
      if ((tid & 31) == 0 && best != 0xffffffffu) atomicMin(&misc[4], best);
    }
    sync_all();
    locate_rank<H2_BINS, THREADS, true, false>(cluster, csize, h2, k_rank, misc);
    const uint32_t key_lo = prefix21 | misc[0];
    below += misc[1];
    const uint32_t bincount = misc[2];
    uint32_t key_hi = key_lo;
    if (kp.hi >= below + bincount) {
      if (misc[5] < (uint32_t)H2_BINS) {
        key_hi = prefix21 | misc[5];
      } else {
        uint32_t best = 0xffffffffu;
        for (unsigned r = 0; r < csize; ++r) best = min(best, peer(misc, r)[4]);
        key_hi = best;
      }
    }
    bool has_nan = false;
    for (unsigned r = 0; r < csize; ++r) has_nan |= (peer(misc, r)[3] != 0);
    thr = lerp_torch(__uint_as_float(key_lo), __uint_as_float(key_hi), kp.w, kp.p.lerp_fma);
    if (has_nan) thr = __int_as_float(0x7fc00000);
  }
  if (csize > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");  // last DSMEM access is behind us
  return thr;
}


// ---- phase C: threshold mask + posterior blend + DDIM update of one slice -----------------------------------------------
// FAST: epsilon prediction, fp32 sample and outputs (every BASELINE configuration); the generic instantiation covers
// the other prediction types and 16-bit samples.
template <typename T, int THREADS, bool FAST, bool TMEM>
__device__ __forceinline__ void guided_update_slice(const FusedKParams& kp, const float* u_s, float thr, int64_t b, int64_t base,
                                                    uint32_t tbase) {
  using FV = Vec16<T>;
  const du_fused_params& p = kp.p;
  const du_ddim_coeffs dc = p.ddim;
  const bool higher = p.higher != 0;
  const float post_M = p.post_M, inv_ah = p.inv_alpha_hat, inv_sa = kp.inv_sqrt_alpha_t;
  const float inv_sb = FAST ? 0.0f : 1.0f / dc.sqrt_beta_t;
  const int tid = threadIdx.x;
  const int ng4 = (int)(kp.L / 4);
  const int64_t xrow = b * p.sample_stride + base, erow = b * p.eps_stride + base;
  const float* Srow = p.S ? (p.S + (p.S_broadcast ? 0 : b * p.S_stride) + base) : nullptr;
  const int pt = FAST ? DU_PRED_EPSILON : dc.prediction_type;
  const bool reclip = FAST ? false : (dc.use_clipped_model_output != 0);
  const bool skip = FAST ? false : (p.skip_ddim != 0);   // guided score only (no sample, no x_{t-1}); generic instantiation
  int trip = 0;
#pragma unroll 4
  for (int g = tid; g < ng4; g += THREADS, ++trip) {
    float s[4], e0[4], S[4];
    if (skip) {
      s[0] = s[1] = s[2] = s[3] = 0.0f;
    } else if (FAST) {
      const uint4 r = ldg_stream_128_pol(reinterpret_cast<const float*>(p.sample) + xrow + 4 * g, kL2EvictFirst);
      s[0] = __uint_as_float(r.x); s[1] = __uint_as_float(r.y); s[2] = __uint_as_float(r.z); s[3] = __uint_as_float(r.w);
    } else {
      load4(p.sample, xrow + 4 * g, p.sample_dtype, s);
    }
    if (TMEM) {
      const uint4 r = tmem_ld4(tmem_slot(tbase, kp.tmem_cpg, trip));
      e0[0] = __uint_as_float(r.x); e0[1] = __uint_as_float(r.y); e0[2] = __uint_as_float(r.z); e0[3] = __uint_as_float(r.w);
    } else {
      load4(p.eps, erow + 4 * g, FV::DT, e0);
    }
    if (Srow) {
      const float4 s4 = ld_coherent_f4(Srow + 4 * g);
      S[0] = s4.x; S[1] = s4.y; S[2] = s4.z; S[3] = s4.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) S[e] = e0[e];
    }
    const float4 u4 = *reinterpret_cast<const float4*>(u_s + 4 * g);
    const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
    float pv[4], x0v[4], eg[4], mk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mk[e] = (higher ? (uu[e] > thr) : (uu[e] < thr)) ? 1.0f : 0.0f;
      const float inv_var = rcp_fast(uu[e]);                        // 1/u (u = 0 -> inf, as the reference)
      const float prec = rcp_fast(fmaf(post_M, inv_var, inv_ah));   // 1/(M/u + 1/abar)
      const float post = prec * (inv_var * S[e]);
      eg[e] = fmaf(mk[e], post, e0[e] * (1.0f - mk[e]));            // eps(1-m) + m*post (NaN/inf of `post` propagate)
      float x0, en;
      if (pt == DU_PRED_EPSILON) { x0 = (s[e] - dc.sqrt_beta_t * eg[e]) * inv_sa; en = eg[e]; }
      else if (pt == DU_PRED_SAMPLE) { x0 = eg[e]; en = (s[e] - dc.sqrt_alpha_t * x0) * inv_sb; }
      else { x0 = dc.sqrt_alpha_t * s[e] - dc.sqrt_beta_t * eg[e]; en = dc.sqrt_alpha_t * eg[e] + dc.sqrt_beta_t * s[e]; }
      if (dc.clip_sample) x0 = clamp_nan(x0, -dc.clip_range, dc.clip_range);
      if (reclip) en = (s[e] - dc.sqrt_alpha_t * x0) * inv_sb;
      pv[e] = fmaf(dc.sqrt_alpha_prev, x0, dc.dir_coef * en);
      x0v[e] = x0;
    }
    if (FAST) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.prev_out) + b * p.prev_stride + base + 4 * g) = make_float4(pv[0], pv[1], pv[2], pv[3]);
    else if (!skip) store4(p.prev_out, b * p.prev_stride + base + 4 * g, p.prev_dtype, pv);
    if (p.x0_out && !skip) store4(p.x0_out, b * p.x0_stride + base + 4 * g, p.prev_dtype, x0v);
    if (p.eps_out) store4(p.eps_out, b * p.eps_out_stride + base + 4 * g, DU_F32, eg);
    if (p.mask_out) store4(p.mask_out, b * p.mask_out_stride + base + 4 * g, DU_F32, mk);
  }
}


struct FusedPlan { int cluster; int threads; int minb; size_t smem; };

// du_fused_pred.cu: launches the predictive kernel when the configuration is eligible; returns 1 if it did, 0 if the
// caller should launch the three-phase kernel, < 0 on error.
int launch_fused_pred(const FusedKParams& kp, const FusedPlan& plan, cudaStream_t st);

}  // namespace du
