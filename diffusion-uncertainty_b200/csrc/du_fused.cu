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
constexpr int LIST_CAP = 1024;
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

__device__ __forceinline__ void stamp(const FusedKParams& kp, int slot) {
  if (kp.timeline && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    kp.timeline[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + slot] = t;
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
  if constexpr (MT > 0) accumulate_scores_ct<T, MT>(p.scores, srow, byte_off, raw_c, centre_mode, false, centre_mode != 1, c, k, s1, s2);
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
template <int NBINS, int THREADS, bool LAST, bool PACKED16>
__device__ __forceinline__ void locate_rank(cg::cluster_group& cluster, unsigned csize, uint32_t* hist_local, uint32_t k,
                                            uint32_t* misc) {
  constexpr int PER = (NBINS + THREADS - 1) / THREADS;
  static_assert(!PACKED16 || PER % 2 == 0, "packed histograms need an even number of bins per thread");
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
        if constexpr (W % 4 == 0) {
#pragma unroll
          for (int j = 0; j < W; j += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(h + first / 2 + j);
            c[2 * j] += q.x & 0xffffu; c[2 * j + 1] += q.x >> 16; c[2 * j + 2] += q.y & 0xffffu; c[2 * j + 3] += q.y >> 16;
            c[2 * j + 4] += q.z & 0xffffu; c[2 * j + 5] += q.z >> 16; c[2 * j + 6] += q.w & 0xffffu; c[2 * j + 7] += q.w >> 16;
          }
        } else if constexpr (W % 2 == 0) {
#pragma unroll
          for (int j = 0; j < W; j += 2) {
            const uint2 q = *reinterpret_cast<const uint2*>(h + first / 2 + j);
            c[2 * j] += q.x & 0xffffu; c[2 * j + 1] += q.x >> 16; c[2 * j + 2] += q.y & 0xffffu; c[2 * j + 3] += q.y >> 16;
          }
        } else {
#pragma unroll
          for (int j = 0; j < W; ++j) { const uint32_t q = h[first / 2 + j]; c[2 * j] += q & 0xffffu; c[2 * j + 1] += q >> 16; }
        }
      } else if constexpr (PER % 4 == 0) {   // one 16-byte (DSMEM) load per 4 bins
#pragma unroll
        for (int j = 0; j < PER; j += 4) {
          const uint4 q = *reinterpret_cast<const uint4*>(h + first + j);
          c[j] += q.x; c[j + 1] += q.y; c[j + 2] += q.z; c[j + 3] += q.w;
        }
      } else if constexpr (PER % 2 == 0) {
#pragma unroll
        for (int j = 0; j < PER; j += 2) {
          const uint2 q = *reinterpret_cast<const uint2*>(h + first + j);
          c[j] += q.x; c[j + 1] += q.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < PER; ++j) c[j] += h[first + j];
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
      if (k < cum + c[j]) { misc[0] = first + j; misc[1] = cum; misc[2] = c[j]; break; }
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
__device__ __forceinline__ void cluster_barrier() {
  __syncthreads();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
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
    locate_rank<H1_BINS, THREADS, true, false>(cluster, csize, h1, kp.lo - below0, misc);
    const uint32_t d1 = misc[0], below1 = misc[1], cnt1 = misc[2], next1 = misc[5];
    const uint32_t prefix = want | (d1 << H2_BITS);            // 22 known bits of key_lo
    const uint32_t k2 = kp.lo - below0 - below1;               // rank of key_lo among the cnt1 keys with that prefix
    const bool need_next = kp.hi > kp.lo;                      // the upper statistic is the next key in sorted order
    // ---- pick the keys with the prefix (and, if the successor lives in the next non-empty level-1 bin, the smallest key
    // there) out of the cluster's lists.  h0 is free now: the tiny list goes to its first words.
    uint32_t* tiny = h0;
    bool has_nan = false;
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
    __syncthreads();
    if (tid < 32) {
      // warp 0: k2-th smallest of the cnt1 keys in `tiny`.  They share 22 bits: bitwise search over the low 9.
      const uint32_t m = misc[40];   // == cnt1
      uint32_t ans = 0;
#pragma unroll 1
      for (int b = H2_BITS - 1; b >= 0; --b) {
        const uint32_t trial = prefix | ans | (1u << b);
        uint32_t c = 0;
        for (uint32_t j = lane; j < m; j += 32) c += (tiny[j] < trial);
        c = __reduce_add_sync(0xffffffffu, c);
        if (c <= k2) ans |= (1u << b);
      }
      const uint32_t key_lo = prefix | ans;
      uint32_t key_hi = key_lo;
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
        else key_hi = misc[4];   // 0xffffffff: the successor lives in a higher level-0 bin (rare path below)
      }
      if (lane == 0) { misc[41] = key_lo; misc[43] = key_hi; }
    }
    __syncthreads();
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
  int trip = 0;
#pragma unroll 4
  for (int g = tid; g < ng4; g += THREADS, ++trip) {
    float s[4], e0[4], S[4];
    if (FAST) {
      const uint4 r = ldg_stream_128(reinterpret_cast<const float*>(p.sample) + xrow + 4 * g);
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
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(Srow + 4 * g));
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
    else store4(p.prev_out, b * p.prev_stride + base + 4 * g, p.prev_dtype, pv);
    if (p.x0_out) store4(p.x0_out, b * p.x0_stride + base + 4 * g, p.prev_dtype, x0v);
    if (p.eps_out) store4(p.eps_out, b * p.eps_out_stride + base + 4 * g, DU_F32, eg);
    if (p.mask_out) store4(p.mask_out, b * p.mask_out_stride + base + 4 * g, DU_F32, mk);
  }
}

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
  if (tid == 0) {  // phase C inputs -> L2 while the select runs
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
                      p.sample_dtype == DU_F32 && p.prev_dtype == DU_F32;
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

struct FusedPlan { int cluster; int threads; int minb; size_t smem; };

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

extern "C" int du_fused_supported(int64_t n, int score_dtype) {
  if (!dtype_ok(score_dtype) || n <= 0 || n > ((int64_t)1 << 24)) return 0;
  FusedPlan plan;
  return fused_plan(n, score_dtype == DU_F32 ? 4 : 8, 1 << 20, &plan) ? plan.cluster : 0;
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
  if (!fused_plan(p->n, vec, p->B, &plan))
    return set_error(DU_ERR_TOO_LARGE, "du_fused_uncertainty_step: rows of %lld elements do not fit cluster shared memory; use the unfused calls", (long long)p->n);
  // 128-bit access requirements
  bool ok = aligned(p->eps, 16) && (p->eps_stride % vec == 0) && (p->score_stride % vec == 0) &&
            aligned(p->sample, 16) && (p->sample_stride % (p->sample_dtype == DU_F32 ? 4 : 8) == 0) &&
            aligned(p->unc_out, 16) && (p->unc_stride % 4 == 0) &&
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
                        p->sample_dtype == DU_F32 && p->prev_dtype == DU_F32;
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
  const size_t tl_words = (size_t)p->B * plan.cluster * 8;
  if (tl_path) {
    DU_CUDA(cudaMalloc(&kp.timeline, tl_words * 8));
    DU_CUDA(cudaMemsetAsync(kp.timeline, 0, tl_words * 8, st));
  }
  int rc;
  switch (p->score_dtype) {
    case DU_F32: rc = launch_fused<float>(kp, plan, st); break;
    case DU_F16: rc = launch_fused<__half>(kp, plan, st); break;
    default: rc = launch_fused<__nv_bfloat16>(kp, plan, st); break;
  }
  if (tl_path) {
    unsigned long long* h = (unsigned long long*)malloc(tl_words * 8);
    cudaMemcpyAsync(h, kp.timeline, tl_words * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    FILE* f = fopen(tl_path, "w");
    if (f) {
      for (size_t c = 0; c < tl_words / 8; ++c) {
        for (int k = 0; k < 8; ++k) fprintf(f, "%llu%c", h[c * 8 + k], k == 7 ? '\n' : ' ');
      }
      fclose(f);
    }
    free(h);
    cudaFree(kp.timeline);
  }
  return rc;
}
