// membench.cu — load-pattern micro-benchmark for the moments reduction (M score streams + eps -> one map).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/membench tools/membench.cu
// Run on the B200 box: ./tools/membench  (prints GB/s of algorithmic bytes per variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 5;
struct Ptrs { const float* s[M]; const float* eps; float* out; };

__device__ __forceinline__ float4 ldnc(const float* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldplain(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float4 var6(const float4 c, const float4 (&x)[M]) {
  float4 r;
  float* rr = reinterpret_cast<float*>(&r);
  const float* cc = reinterpret_cast<const float*>(&c);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float k = reinterpret_cast<const float*>(&x[0])[e], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int m = 0; m < M; ++m) { float d = reinterpret_cast<const float*>(&x[m])[e] - k; s1 += d; s2 = fmaf(d, d, s2); }
    float d = cc[e] - k; s1 += d; s2 = fmaf(d, d, s2);
    rr[e] = (s2 - s1 * s1 / 6.f) / 5.f;
  }
  return r;
}

// V1: grid-stride, U groups per trip, all loads of a trip issued first
template <int U, bool NC>
__global__ void __launch_bounds__(256) k_gridstride(Ptrs p, int64_t ngroups) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < ngroups; g0 += stride * U) {
    float4 c[U], x[U][M];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t g = g0 + u * stride;
      if (g < ngroups) {
        c[u] = NC ? ldnc(p.eps + 4 * g) : ldplain(p.eps + 4 * g);
#pragma unroll
        for (int m = 0; m < M; ++m) x[u][m] = NC ? ldnc(p.s[m] + 4 * g) : ldplain(p.s[m] + 4 * g);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int64_t g = g0 + u * stride;
      if (g < ngroups) *reinterpret_cast<float4*>(p.out + 4 * g) = var6(c[u], x[u]);
    }
  }
}

// V2: TMA bulk (cp.async.bulk) multi-stage pipeline; warp 0 lane 0 produces, all warps consume
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TILE /*floats per tensor per stage*/, int STAGES, int THREADS>
__global__ void __launch_bounds__(THREADS) k_tma(Ptrs p, int64_t ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* buf = reinterpret_cast<float*>(smem);                       // [STAGES][M+1][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(buf + (size_t)STAGES * (M + 1) * TILE);
  uint64_t* empty = full + STAGES;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // number of tiles this CTA owns
  int64_t my = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) ++my;
  constexpr uint32_t TB = TILE * 4;
  auto issue = [&](int64_t k) {  // k-th tile of this CTA
    const int s = (int)(k % STAGES);
    const int64_t t = blockIdx.x + k * gridDim.x;
    float* dst = buf + (size_t)s * (M + 1) * TILE;
    mbar_expect_tx(&full[s], TB * (M + 1));
    bulk_g2s(dst, p.eps + t * TILE, TB, &full[s]);
#pragma unroll
    for (int m = 0; m < M; ++m) bulk_g2s(dst + (m + 1) * TILE, p.s[m] + t * TILE, TB, &full[s]);
  };
  if (tid == 0) for (int64_t k = 0; k < STAGES && k < my; ++k) issue(k);
  for (int64_t k = 0; k < my; ++k) {
    const int s = (int)(k % STAGES);
    const uint32_t par = (uint32_t)((k / STAGES) & 1);
    mbar_wait(&full[s], par);
    const float* src = buf + (size_t)s * (M + 1) * TILE;
    const int64_t t = blockIdx.x + k * gridDim.x;
    for (int g = tid; g < TILE / 4; g += THREADS) {
      float4 c = *reinterpret_cast<const float4*>(src + 4 * g);
      float4 x[M];
#pragma unroll
      for (int m = 0; m < M; ++m) x[m] = *reinterpret_cast<const float4*>(src + (m + 1) * TILE + 4 * g);
      *reinterpret_cast<float4*>(p.out + t * TILE + 4 * g) = var6(c, x);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
    if (tid == 0 && k + STAGES < my) {
      mbar_wait(&empty[s], par);   // all warps done with stage s
      issue(k + STAGES);
    }
  }
}

// pure copy reference: read 6 streams, write 1 (same bytes as the reduction)
int main() {
  const int64_t N = 128LL * 3 * 128 * 128;
  Ptrs p;
  std::vector<float*> all;
  for (int m = 0; m < M + 2; ++m) { float* d; CK(cudaMalloc(&d, N * 4)); CK(cudaMemset(d, 0, N * 4)); all.push_back(d); }
  for (int m = 0; m < M; ++m) p.s[m] = all[m];
  p.eps = all[M]; p.out = all[M + 1];
  const double bytes = (double)N * 4 * (M + 2);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9);
  };
  const int64_t ng = N / 4;
  char name[128];
  for (int ctas_per_sm : {2, 4, 8, 16, 0}) {
    int grid = ctas_per_sm ? 148 * ctas_per_sm : (int)((ng + 255) / 256);
    snprintf(name, 128, "gridstride U=1 nc grid=%d", grid); timeit(name, [&] { k_gridstride<1, true><<<grid, 256>>>(p, ng); });
    snprintf(name, 128, "gridstride U=2 nc grid=%d", grid); timeit(name, [&] { k_gridstride<2, true><<<grid, 256>>>(p, ng); });
    snprintf(name, 128, "gridstride U=4 nc grid=%d", grid); timeit(name, [&] { k_gridstride<4, true><<<grid, 256>>>(p, ng); });
    snprintf(name, 128, "gridstride U=2 plain grid=%d", grid); timeit(name, [&] { k_gridstride<2, false><<<grid, 256>>>(p, ng); });
  }
#define TMA_CASE(TILE, STAGES, THREADS, CPS) { \
    auto kern = k_tma<TILE, STAGES, THREADS>; \
    size_t sm = (size_t)STAGES * (M + 1) * TILE * 4 + 2 * STAGES * 8; \
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    int grid = 148 * CPS; int64_t nt = N / TILE; \
    snprintf(name, 128, "tma tile=%d stages=%d thr=%d cta/sm=%d smem=%zuK", TILE, STAGES, THREADS, CPS, sm / 1024); \
    timeit(name, [&] { kern<<<grid, THREADS, sm>>>(p, nt); }); }
  TMA_CASE(1024, 2, 256, 2)
  TMA_CASE(1024, 3, 256, 2)
  TMA_CASE(1024, 4, 256, 2)
  TMA_CASE(1024, 4, 256, 1)
  TMA_CASE(2048, 2, 256, 2)
  TMA_CASE(2048, 4, 256, 1)
  TMA_CASE(2048, 3, 512, 1)
  TMA_CASE(512, 4, 128, 4)
  TMA_CASE(512, 6, 256, 2)
  TMA_CASE(1024, 2, 256, 4)
  TMA_CASE(1024, 2, 128, 4)
  // memcpy reference (D2D copy of 100 MB: read+write)
  { float* a = all[0]; float* b = all[1];
    timeit("cudaMemcpyAsync D2D 2x25MB (bytes col n/a)", [&] { cudaMemcpyAsync(b, a, N * 4, cudaMemcpyDeviceToDevice); }); }
  return 0;
}
