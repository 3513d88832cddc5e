// How fast can SMs READ pinned host memory?  The zero-copy host step reads 786 KB of actions per step
// (65,536 envs x 3 x f32) with ordinary per-thread loads and gets ~20 GB/s; the copy engine gets ~50 GB/s.
// Variants: (a) per-thread 4-byte loads of an [N][3] row (what k_step does), (b) one 16-byte load per thread,
// (c) cp.async.bulk of the block's contiguous chunk into shared memory (block = 64 / 256 envs), (d) DMA copy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_scalar(const float* __restrict__ a, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = a[3 * i] + a[3 * i + 1] + a[3 * i + 2];
}
__global__ void k_vec4(const float4* __restrict__ a, float* __restrict__ out, int n4) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = a[i];
  out[i] = v.x + v.y + v.z + v.w;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k_bulk(const float* __restrict__ a, float* __restrict__ out, int n) {
  extern __shared__ __align__(128) float sm[];
  __shared__ uint64_t bar;
  const int i0 = blockIdx.x * blockDim.x;
  const uint32_t bytes = blockDim.x * 12u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm)),
                 "l"(a + 3 * (size_t)i0), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
  const int t = threadIdx.x;
  if (i0 + t < n) out[i0 + t] = sm[3 * t] + sm[3 * t + 1] + sm[3 * t + 2];
}

int main() {
  const int n = 65536;
  const size_t bytes = (size_t)n * 12;
  float *h, *d_out, *d_in;
  cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
  for (size_t k = 0; k < (size_t)n * 3; ++k) h[k] = (float)(k % 7);
  cudaMalloc(&d_out, n * 4); cudaMalloc(&d_in, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn) {
    for (int k = 0; k < 5; ++k) fn();
    cudaDeviceSynchronize(); cudaEventRecord(e0);
    for (int k = 0; k < 50; ++k) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-34s %7.2f us  %6.1f GB/s  (%s)\n", name, ms / 50 * 1e3, bytes / (ms / 50 * 1e-3) * 1e-9, cudaGetErrorString(cudaGetLastError()));
  };
  for (int blk : {64, 256}) {
    char nm[64];
    snprintf(nm, 64, "scalar loads, block %d", blk); timeit(nm, [&] { k_scalar<<<n / blk, blk>>>(h, d_out, n); });
    snprintf(nm, 64, "float4 loads, block %d", blk); timeit(nm, [&] { k_vec4<<<(n * 3 / 4) / blk, blk>>>((const float4*)h, d_out, n * 3 / 4); });
    snprintf(nm, 64, "cp.async.bulk per block, block %d", blk); timeit(nm, [&] { k_bulk<<<n / blk, blk, blk * 12>>>(h, d_out, n); });
  }
  timeit("cudaMemcpyAsync H2D (copy engine)", [&] { cudaMemcpyAsync(d_in, h, bytes, cudaMemcpyHostToDevice, 0); });
  return 0;
}
