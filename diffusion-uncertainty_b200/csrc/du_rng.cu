// du_rng.cu — F7 + SURVEY.md §8f N1: the perturbation builders with the noise drawn IN the kernel.
//
// The reference draws `torch.randn_like(x)` and then combines it with a state tensor in 3 more eager kernels
// (SU/scheduling_ddim_uncertainty_zigzag_centered.py:529-546, uncertainty_guidance.py:86-88): 4 B/element written by the
// generator kernel, read again by the combination.  Here the normal variate is produced in registers and consumed at once:
// x is read, out is written, nothing else touches HBM (8 B/element instead of 16; 12 when the caller also wants the noise).
//
// Bit parity of the noise with torch's CUDA generator is part of the contract (the noise stream is part of the sampling
// trajectory).  torch's `normal_` on a CUDA tensor (ATen/native/cuda/DistributionTemplates.h, `normal_and_transform` ->
// `distribution_nullary_kernel` -> `distribution_elementwise_grid_stride_kernel`, unroll 4) maps variates to elements as:
//   G = min(SMs * (maxThreadsPerSM / 256), ceil(N / 256)) blocks of 256 virtual threads, virtual thread v = block*256 + lane
//   has Philox4x32-10 subsequence v of (seed, offset); its k-th `curand_normal4` call yields the elements
//   li = v + 256*G*(4k + ii), ii = 0..3.
// The kernel below runs exactly that geometry (one real thread per virtual thread), using curand's own device functions, so
// every element gets the same four Philox words and the same Box-Muller arithmetic as torch's kernel.  After the call the
// generator must be advanced by du_randn_offset_increment(N) (what torch's calc_execution_policy adds).
#include <curand_kernel.h>

#include "du_common.cuh"

namespace du {
namespace {

constexpr int kRngBlock = 256;     // torch: block_size_bound
constexpr int kRngUnroll = 4;      // sizeof(float4) / sizeof(float)

struct RngGeometry { unsigned grid; uint64_t increment; };

int rng_geometry(int64_t N, RngGeometry* g) {
  int dev = 0, sms = 0, threads_per_sm = 0;
  DU_CUDA(cudaGetDevice(&dev));
  DU_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DU_CUDA(cudaDeviceGetAttribute(&threads_per_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
  const uint64_t numel = (uint64_t)N;
  uint64_t grid = (numel + kRngBlock - 1) / kRngBlock;
  const uint64_t cap = (uint64_t)sms * (uint64_t)(threads_per_sm / kRngBlock);
  if (grid > cap) grid = cap;
  g->grid = (unsigned)grid;
  // torch: ((numel - 1) / (block * grid * unroll) + 1) * max_generator_offsets_per_curand_call(4); philox_cuda_state rounds
  // the increment up to a multiple of 4 (it already is one)
  g->increment = ((numel - 1) / ((uint64_t)kRngBlock * grid * kRngUnroll) + 1) * 4;
  return DU_OK;
}

// state: nullable device pointer to {seed, offset} (graph-replayable draws: the pair lives in device memory and is advanced
// by du_rng_advance inside the same graph); otherwise the host values are used.
template <bool HAS_X>
__global__ void __launch_bounds__(kRngBlock, 4)
perturb_randn_kernel(const void* __restrict__ x, int x_dtype, int64_t N, uint64_t seed, uint64_t offset,
                     const uint64_t* __restrict__ state, float a, float b, void* __restrict__ out, int out_dtype,
                     void* __restrict__ noise_out, int noise_dtype) {
  if (state) { seed = state[0]; offset += state[1]; }   // device-resident state: the host `offset` is an extra offset on top of it
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)blockDim.x * gridDim.x;
  curandStatePhilox4_32_10_t st;
  curand_init(seed, (unsigned long long)v, offset, &st);
  // Two of torch's loop iterations per trip (KU = 2): the 8 loads of both are issued before the first Philox round, which doubles
  // the bytes in flight per thread — the kernel is bound by memory-level parallelism at 4 CTAs per SM (59 registers), not by the
  // Philox / Box-Muller arithmetic.  The element <-> variate mapping is untouched: iteration k of virtual thread v still gets
  // the k-th curand_normal4 of subsequence v.
  constexpr int KU = 2;
  for (int64_t base = v; base < N; base += stride * kRngUnroll * KU) {
    float xv[KU][kRngUnroll];
    if (HAS_X) {
#pragma unroll
      for (int k = 0; k < KU; ++k) {
#pragma unroll
        for (int ii = 0; ii < kRngUnroll; ++ii) {
          const int64_t li = base + stride * (kRngUnroll * k + ii);
          xv[k][ii] = li < N ? load1(x, li, x_dtype) : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      if (base + stride * kRngUnroll * k >= N) break;
      const float4 r4 = curand_normal4(&st);
      const float r[kRngUnroll] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int ii = 0; ii < kRngUnroll; ++ii) {
        const int64_t li = base + stride * (kRngUnroll * k + ii);
        if (li >= N) continue;
        // torch: static_cast<scalar_t>(rand * std + mean) with std = 1, mean = 0, then the tensor's dtype
        float nz = __fmaf_rn(r[ii], 1.0f, 0.0f);
        if (noise_dtype == DU_F16) nz = __half2float(__float2half_rn(nz));
        else if (noise_dtype == DU_BF16) nz = __bfloat162float(__float2bfloat16_rn(nz));
        if (noise_out) store1(noise_out, li, noise_dtype, nz);
        if (HAS_X) store1(out, li, out_dtype, __fadd_rn(__fmul_rn(a, xv[k][ii]), __fmul_rn(b, nz)));   // as PerturbF (du_step.cu)
      }
    }
  }
}

__global__ void rng_advance_kernel(uint64_t* state, uint64_t increment) {
  if (threadIdx.x == 0 && blockIdx.x == 0) state[1] += increment;
}

}  // namespace
}  // namespace du

using namespace du;

extern "C" int du_randn_offset_increment(int64_t N, uint64_t* increment_out) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (N < 0 || !increment_out) return set_error(DU_ERR_BAD_ARG, "du_randn_offset_increment: bad arguments");
  if (N == 0) { *increment_out = 0; return DU_OK; }
  RngGeometry g;
  int rc = rng_geometry(N, &g);
  if (rc != DU_OK) return rc;
  *increment_out = g.increment;
  return DU_OK;
}

extern "C" int du_perturb_randn(const void* x, int x_dtype, int64_t N, uint64_t seed, uint64_t offset,
                                const uint64_t* device_state, float a, float b, void* out, int out_dtype,
                                void* noise_out, int noise_dtype, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (N < 0 || !dtype_ok(noise_dtype) || (x && (!dtype_ok(x_dtype) || !out || !dtype_ok(out_dtype))) || (!x && !noise_out))
    return set_error(DU_ERR_BAD_ARG, "du_perturb_randn: bad arguments");
  if (N > 0x7fffffffLL)   // torch splits such tensors into 32-bit-indexable pieces with one generator call each
    return set_error(DU_ERR_TOO_LARGE, "du_perturb_randn: more than 2^31-1 elements");
  if (N == 0) return DU_OK;
  RngGeometry g;
  int rc = rng_geometry(N, &g);
  if (rc != DU_OK) return rc;
  if (x)
    perturb_randn_kernel<true><<<g.grid, kRngBlock, 0, (cudaStream_t)stream>>>(x, x_dtype, N, seed, offset, device_state, a, b,
                                                                               out, out_dtype, noise_out, noise_dtype);
  else
    perturb_randn_kernel<false><<<g.grid, kRngBlock, 0, (cudaStream_t)stream>>>(nullptr, DU_F32, N, seed, offset, device_state,
                                                                                a, b, nullptr, DU_F32, noise_out, noise_dtype);
  DU_LAUNCH_CHECK("perturb_randn_kernel");
  return DU_OK;
}

extern "C" int du_rng_advance(uint64_t* device_state, uint64_t increment, du_stream_t stream) {
  du::DeviceGuard _dg;   // the device Python selected for this thread (du_set_device), restored on return
  if (!device_state) return set_error(DU_ERR_BAD_ARG, "du_rng_advance: null state");
  rng_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(device_state, increment);
  DU_LAUNCH_CHECK("rng_advance_kernel");
  return DU_OK;
}
