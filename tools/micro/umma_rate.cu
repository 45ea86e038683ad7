// Micro-benchmark: sustained rate of back-to-back tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, operands in
// shared memory (K-major SWIZZLE_128B, zeros), one CTA per SM.  Answers: how far below the tensor peak do the N = 64 / N = 96 MMAs of
// the policy's convolutions run?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
template <int N, int SAME_ACC>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncwarp();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tbase;
  if (threadIdx.x == 0) {
    const uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 128 * 128));
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t d = tmem + (SAME_ACC ? 0 : ((i & 1) * 256));
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc(128, N)), "r"(1) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int N, int SAME>
void run(int sms) {
  const int iters = 4096;
  unsigned long long* d; cudaMalloc(&d, sms * 8);
  const size_t smem = 128 * 128 + 256 * 128 + 1024;
  cudaFuncSetAttribute(k_rate<N, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rate<N, SAME><<<sms, 128, smem>>>(16, d);
  k_rate<N, SAME><<<sms, 128, smem>>>(iters, d);
  cudaDeviceSynchronize();
  unsigned long long h[256]; cudaMemcpy(h, d, sms * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; i++) c += h[i]; c /= sms;
  const double per = c / (iters * 4.0), ideal = 128.0 * N * 16 * 2 / 8192.0;
  printf("M=128 N=%3d K=16 %s: %.1f cycles per MMA (ideal at 8192 flop/clk/SM: %.0f) -> %.0f %% of the per-SM tensor peak, err=%s\n", N, SAME ? "one accumulator " : "two accumulators",
         per, ideal, 100.0 * ideal / per, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<32, 1>(sms); run<64, 1>(sms); run<96, 1>(sms); run<128, 1>(sms); run<256, 1>(sms);
  run<64, 0>(sms); run<96, 0>(sms); run<128, 0>(sms); run<256, 0>(sms);
  return 0;
}
