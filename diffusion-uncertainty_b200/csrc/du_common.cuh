// du_common.cuh — shared device helpers for the sm_100a uncertainty-path kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

#include "../../include/du_b200.h"

namespace du {

// ---- error plumbing (du_abi.cu) ----------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#define DU_CUDA(call)                                   \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return ::du::check_cuda(_e, #call); \
  } while (0)
#define DU_LAUNCH_CHECK(name) DU_CUDA(cudaPeekAtLastError())

// ---- device selection (du_abi.cu) -----------------------------------------------------------------
// du_set_device() records the device the calling THREAD wants its next calls to run on; it does not touch the CUDA
// current device.  Every launching entry point holds a DeviceGuard: if the thread's current device differs from the
// wanted one it switches for the duration of the call and restores the caller's device on return, so the library never
// changes what torch.cuda.current_device() reports and never launches on a stale device (the selection is thread-local
// on both sides of the ABI).
int wanted_device();
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  DeviceGuard() {
    const int want = wanted_device();
    if (want < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    if (prev != want && cudaSetDevice(want) == cudaSuccess) switched = true;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

inline bool dtype_ok(int dt) { return dt == DU_F32 || dt == DU_F16 || dt == DU_BF16; }
inline int dtype_size(int dt) { return dt == DU_F32 ? 4 : 2; }
inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- scalar / vector loads with runtime dtype (warp-uniform switch) --------------------------------
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b16) { return __uint_as_float(b16 << 16); }
__device__ __forceinline__ float f16_bits_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }

__device__ __forceinline__ float load1(const void* base, int64_t idx, int dt) {
  if (dt == DU_F32) return __ldg(reinterpret_cast<const float*>(base) + idx);
  uint16_t h = __ldg(reinterpret_cast<const uint16_t*>(base) + idx);
  return dt == DU_F16 ? f16_bits_to_float(h) : bf16_bits_to_float(h);
}

__device__ __forceinline__ void store1(void* base, int64_t idx, int dt, float v) {
  if (dt == DU_F32) reinterpret_cast<float*>(base)[idx] = v;
  else if (dt == DU_F16) reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
}

// streaming 128-bit load: read-only path, do not allocate in L1.  volatile so that the compiler never
// speculates a predicated load (a null score pointer past M would fault); callers therefore issue a whole batch
// of these in program order BEFORE the first use (memory-level parallelism is what the reduction kernels live on).
__device__ __forceinline__ uint4 ldg_stream_128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream_64(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

// The same load with an L2 eviction-priority hint.  The 64-bit policy words are the encodings `createpolicy.fractional.L2::
// evict_{normal,first,last}` (fraction 1.0) produces — the constants CUTLASS passes as TMA cache hints.  Used to keep a tensor
// that two consecutive kernels read (eps: du_batch_sum, then the fused step) resident in the 126 MB L2 while the once-read
// score / sample streams pass through it.
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ uint4 ldg_stream_128_pol(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
  return r;
}

// Plain (coherent) global loads for data another kernel may still have been writing while THIS kernel was already resident
// (the batch-sum row S of a step launched as a programmatic dependent): the non-coherent path (__ldg / ld.global.nc) is
// only defined for memory that is read-only for the kernel's whole lifetime.
__device__ __forceinline__ float4 ld_coherent_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_coherent_f1(const float* p) {
  float r;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// 16-byte vector of T (4 fp32 or 8 fp16/bf16): raw load now, unpack to fp32 at the point of use
// uniform row pointer + 32-bit per-thread byte offset: two integer instructions per load instead of a 64-bit multiply-add
__device__ __forceinline__ uint4 ldg_stream_128_at(const void* row, uint32_t byte_off) {
  return ldg_stream_128(reinterpret_cast<const char*>(row) + byte_off);
}

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int VEC = 4;
  __device__ static __forceinline__ uint4 ld(const void* p, uint64_t pol) { return ldg_stream_128_pol(p, pol); }
  static constexpr int DT = DU_F32;
  __device__ static __forceinline__ uint4 load(const void* base, int64_t idx) {
    return ldg_stream_128(reinterpret_cast<const float*>(base) + idx);
  }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
};
template <> struct Vec16<__half> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ uint4 ld(const void* p, uint64_t pol) { return ldg_stream_128_pol(p, pol); }
  static constexpr int DT = DU_F16;
  __device__ static __forceinline__ uint4 load(const void* base, int64_t idx) {
    return ldg_stream_128(reinterpret_cast<const uint16_t*>(base) + idx);
  }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __half2float(__ushort_as_half((unsigned short)(w[i] & 0xffff)));
      v[2 * i + 1] = __half2float(__ushort_as_half((unsigned short)(w[i] >> 16)));
    }
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ uint4 ld(const void* p, uint64_t pol) { return ldg_stream_128_pol(p, pol); }
  static constexpr int DT = DU_BF16;
  __device__ static __forceinline__ uint4 load(const void* base, int64_t idx) {
    return ldg_stream_128(reinterpret_cast<const uint16_t*>(base) + idx);
  }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};

// 8-byte vector of a 16-bit type (4 elements): the same interface, carried in the low half of a uint4.  Used where 8
// elements per thread cost more registers than the streaming loop has (the predictive fused step with fp16 / bf16 scores).
__device__ __forceinline__ uint4 ldg_stream_64_pol(const void* p, uint64_t pol) {
  uint4 r;
  r.z = 0u; r.w = 0u;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(pol));
  return r;
}
template <typename T> struct Vec8;
template <> struct Vec8<__half> {
  static constexpr int VEC = 4;
  static constexpr int DT = DU_F16;
  __device__ static __forceinline__ uint4 ld(const void* p, uint64_t pol) { return ldg_stream_64_pol(p, pol); }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __half2float(__ushort_as_half((unsigned short)(r.x & 0xffff))); v[1] = __half2float(__ushort_as_half((unsigned short)(r.x >> 16)));
    v[2] = __half2float(__ushort_as_half((unsigned short)(r.y & 0xffff))); v[3] = __half2float(__ushort_as_half((unsigned short)(r.y >> 16)));
  }
};
template <> struct Vec8<__nv_bfloat16> {
  static constexpr int VEC = 4;
  static constexpr int DT = DU_BF16;
  __device__ static __forceinline__ uint4 ld(const void* p, uint64_t pol) { return ldg_stream_64_pol(p, pol); }
  __device__ static __forceinline__ void unpack(const uint4& r, float (&v)[4]) {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
  }
};
template <> struct Vec8<float> : Vec16<float> {};   // (never narrower than 16 bytes for fp32)

// Shifted-data accumulation of one batch of score vectors (SURVEY.md §7 "Variance numerics"): d = x - k with
// k the first sample (variance modes) or the centre; s1 = sum d, s2 = sum d^2.
constexpr int DU_LOAD_BATCH = 8;  // score vectors in flight per thread (16 B each)

// centre_mode: 0 = no centre, 1 = centre is the shift k (DU_MOM_CENTERED), 2 = centre is one more sample
// (DU_MOM_VAR_WITH_CENTER).  raw_c is the already-issued 16-byte load of the centre vector; it is unpacked only
// after the first batch of score loads has been issued, so all of them are in flight together (c_ready: the caller
// already filled c[], e.g. from a centre tensor of another dtype).
template <typename T>
__device__ __forceinline__ void accumulate_scores(const void* const* scores, int M, int64_t off, const uint4& raw_c,
                                                  int centre_mode, bool c_ready, bool shift_first, float (&c)[Vec16<T>::VEC],
                                                  float (&k)[Vec16<T>::VEC], float (&s1)[Vec16<T>::VEC],
                                                  float (&s2)[Vec16<T>::VEC]) {
  using V = Vec16<T>;
  constexpr int VEC = V::VEC;
  for (int m0 = 0; m0 < M; m0 += DU_LOAD_BATCH) {
    uint4 raw[DU_LOAD_BATCH];
#pragma unroll
    for (int j = 0; j < DU_LOAD_BATCH; ++j)
      if (m0 + j < M) raw[j] = V::load(scores[m0 + j], off);
    if (m0 == 0) {
      if (centre_mode && !c_ready) V::unpack(raw_c, c);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { k[e] = (centre_mode == 1) ? c[e] : 0.0f; s1[e] = 0.0f; s2[e] = 0.0f; }
    }
#pragma unroll
    for (int j = 0; j < DU_LOAD_BATCH; ++j) {
      if (m0 + j < M) {
        float x[VEC];
        V::unpack(raw[j], x);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          if (shift_first && m0 + j == 0) k[e] = x[e];
          const float d = x[e] - k[e];
          s1[e] += d;
          s2[e] = fmaf(d, d, s2[e]);
        }
      }
    }
  }
  if (centre_mode == 2) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float d = c[e] - k[e];
      s1[e] += d;
      s2[e] = fmaf(d, d, s2[e]);
    }
  }
}

// Same contract as accumulate_scores with M known at compile time: no predication, every load of a batch (<= 8 score
// vectors + the centre) is issued before the first use.
// HINT: the score loads carry the L2 eviction policy `pol` (see ldg_stream_128_pol)
// SKIP0: the caller guarantees shift_first (the shift IS sample 0): its deviation is x0 - x0, i.e. +0 — or NaN for a non-finite
// x0, in which case every other deviation is NaN or infinite as well and the result is NaN either way — so the term is dropped.
template <typename T, int MT, bool HINT = false, typename V = Vec16<T>, bool SKIP0 = false>
__device__ __forceinline__ void accumulate_scores_ct(const void* const* scores, int64_t row_off, uint32_t byte_off, const uint4& raw_c,
                                                     int centre_mode, bool c_ready, bool shift_first,
                                                     float (&c)[V::VEC], float (&k)[V::VEC],
                                                     float (&s1)[V::VEC], float (&s2)[V::VEC], uint64_t pol = 0) {
  constexpr int VEC = V::VEC;
  constexpr int BATCH = (MT <= 8) ? MT : 8;
  auto load = [&](int m) -> uint4 {
    const char* ptr = reinterpret_cast<const char*>(reinterpret_cast<const T*>(scores[m]) + row_off) + byte_off;
    if constexpr (HINT) return V::ld(ptr, pol);
    else return ldg_stream_128(ptr);
  };
  uint4 raw[BATCH];
#pragma unroll
  for (int m = 0; m < BATCH; ++m) raw[m] = load(m);
  if (centre_mode && !c_ready) V::unpack(raw_c, c);
  if (centre_mode == 1) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) k[e] = c[e];
  } else if (shift_first) {
    V::unpack(raw[0], k);
  } else {
#pragma unroll
    for (int e = 0; e < VEC; ++e) k[e] = 0.0f;
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e) { s1[e] = 0.0f; s2[e] = 0.0f; }
#pragma unroll
  for (int m0 = 0; m0 < MT; m0 += BATCH) {
    if (m0 > 0) {
#pragma unroll
      for (int j = 0; j < BATCH; ++j)
        if (m0 + j < MT) raw[j] = load(m0 + j);
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j) {
      if (m0 + j < MT && !(SKIP0 && m0 + j == 0)) {
        float x[VEC];
        V::unpack(raw[j], x);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float d = x[e] - k[e];
          if (SKIP0 && m0 + j == 1) { s1[e] = d; s2[e] = d * d; }     // (0 + d and fma(d, d, 0): the same values)
          else { s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
        }
      }
    }
  }
  if (centre_mode == 2) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { const float d = c[e] - k[e]; s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
  }
}

// sum of squared deviations about the mean from the shifted sums (NaN is kept; tiny negative rounding residue -> 0)
__device__ __forceinline__ float m2_from_sums(float s1, float s2, float inv_cnt) {
  const float m2 = fmaf(-s1 * inv_cnt, s1, s2);
  return (m2 < 0.0f) ? 0.0f : m2;
}
// The map value of the fused kernels from the shifted sums, never -0.0 (its bit pattern must order like its value):
// centre_mode 1: s2 / count; else max(m2, +0) / (count - 1) with NaN kept (max.NaN) — the same values as
// fmaf(m2_from_sums(..), inv_cm1, 0.0f), two instructions shorter per element.
__device__ __forceinline__ float map_value(int centre_mode, float s1, float s2, float inv_cnt, float inv_cm1) {
  if (centre_mode == 1) return fmaf(s2, inv_cnt, 0.0f);
  float m2 = fmaf(-s1 * inv_cnt, s1, s2);
  asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(m2));
  return m2 * inv_cm1;
}

// 4 consecutive elements starting at element index idx (idx % 4 == 0, pointer suitably aligned)
__device__ __forceinline__ void load4(const void* base, int64_t idx, int dt, float (&v)[4]) {
  if (dt == DU_F32) {
    uint4 r = ldg_stream_128(reinterpret_cast<const float*>(base) + idx);
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  } else {
    uint2 r = ldg_stream_64(reinterpret_cast<const uint16_t*>(base) + idx);
    if (dt == DU_F16) {
      v[0] = f16_bits_to_float(r.x & 0xffff); v[1] = f16_bits_to_float(r.x >> 16);
      v[2] = f16_bits_to_float(r.y & 0xffff); v[3] = f16_bits_to_float(r.y >> 16);
    } else {
      v[0] = bf16_bits_to_float(r.x & 0xffff); v[1] = bf16_bits_to_float(r.x >> 16);
      v[2] = bf16_bits_to_float(r.y & 0xffff); v[3] = bf16_bits_to_float(r.y >> 16);
    }
  }
}

__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void store4(void* base, int64_t idx, int dt, const float (&v)[4]) {
  if (dt == DU_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (dt == DU_F16) {
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = make_uint2(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]));
  } else {
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(base) + idx) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

// vector-path eligibility: every (pointer, stride) pair must keep 4-element groups aligned
inline bool vec4_ok(const void* p, int64_t stride, int dt) {
  if (p == nullptr) return true;
  return aligned(p, 4 * (size_t)dtype_size(dt)) && (stride % 4 == 0);
}

// ---- order-preserving float <-> uint32 key ----------------------------------------------------------
// -0.0 is canonicalised to +0.0 (torch's sort treats them as equal); NaNs are handled by the caller.
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b << 1) == 0) b = 0;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

// torch.lerp's two-branch formula (aten/native/Lerp.h).  fma = 0: one rounding per op (torch's CPU kernel without FMA dispatch);
// fma = 1: contracted like nvcc compiles torch's CUDA kernel (and like torch's vectorised CPU kernels on FMA hardware).
__device__ __forceinline__ float lerp_torch(float a, float b, float w, int fma) {
  float diff = __fsub_rn(b, a);
  if (fabsf(w) < 0.5f) return fma ? __fmaf_rn(w, diff, a) : __fadd_rn(a, __fmul_rn(w, diff));
  float omw = __fsub_rn(1.0f, w);
  return fma ? __fmaf_rn(-diff, omw, b) : __fsub_rn(b, __fmul_rn(diff, omw));
}

// ---- DDIM arithmetic: one IEEE rounding per reference operation (no FMA contraction) ----------------
struct DdimOut { float prev, x0, eps; };

__device__ __forceinline__ DdimOut ddim_update(float mo, float x, float noise, const du_ddim_coeffs& c) {
  float x0, eps;
  if (c.prediction_type == DU_PRED_EPSILON) {
    x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.sqrt_beta_t, mo)), c.sqrt_alpha_t);
    eps = mo;
  } else if (c.prediction_type == DU_PRED_SAMPLE) {
    x0 = mo;
    eps = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.sqrt_alpha_t, x0)), c.sqrt_beta_t);
  } else {
    x0 = __fsub_rn(__fmul_rn(c.sqrt_alpha_t, x), __fmul_rn(c.sqrt_beta_t, mo));
    eps = __fadd_rn(__fmul_rn(c.sqrt_alpha_t, mo), __fmul_rn(c.sqrt_beta_t, x));
  }
  if (c.clip_sample) {  // torch.clamp propagates NaN; fminf/fmaxf would swallow it
    float cl = fminf(fmaxf(x0, -c.clip_range), c.clip_range);
    x0 = (x0 != x0) ? x0 : cl;
  }
  if (c.use_clipped_model_output) eps = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.sqrt_alpha_t, x0)), c.sqrt_beta_t);
  float prev = __fadd_rn(__fmul_rn(c.sqrt_alpha_prev, x0), __fmul_rn(c.dir_coef, eps));
  if (c.add_noise) prev = __fadd_rn(prev, __fmul_rn(c.sigma, noise));
  DdimOut o; o.prev = prev; o.x0 = x0; o.eps = eps;
  return o;
}

// ---- launch geometry --------------------------------------------------------------------------------
struct RowGrid { dim3 grid; dim3 block; };
// The elementwise kernels loop grid-stride in both dimensions, so the grid is capped at a few resident waves' worth of CTAs
// (148 SMs x DU_ROWS_CTAS_PER_SM, default 16; 0 = one CTA per 256 groups as before): thousands of CTAs that each live for a
// microsecond spend their time being launched.
inline int rows_cta_cap() {
  static int cap = -1;
  if (cap < 0) {
    const char* e = getenv("DU_ROWS_CTAS_PER_SM");
    cap = 148 * ((e && atoi(e) >= 0) ? atoi(e) : 16);
  }
  return cap;
}
inline RowGrid row_grid(int64_t B, int64_t n_items_per_row, int threads) {
  RowGrid g;
  g.block = dim3(threads);
  int64_t gx = (n_items_per_row + threads - 1) / threads;
  if (gx < 1) gx = 1;
  int64_t gy = B < 65535 ? B : 65535;
  const int64_t cap = rows_cta_cap();
  if (cap > 0 && gx * gy > cap) {
    if (gy >= cap) { gx = 1; gy = cap; }
    else { gx = (cap + gy - 1) / gy; }
  }
  g.grid = dim3((unsigned)gx, (unsigned)gy);
  return g;
}

}  // namespace du
