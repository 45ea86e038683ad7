// Micro-benchmark: what a persistent CTA per SM can stream from HBM / L2 into shared memory with cp.async.bulk (1-D TMA), as a function of
// the chunk size and the number of chunks in flight.  One producer thread per CTA, mbarrier per ring slot, nobody consumes the data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_stream bulk_stream.cu && ./bulk_stream
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) k_stream(const uint8_t* src, size_t total_chunks, int chunk, int depth, size_t stride_chunks) {
  extern __shared__ __align__(1024) uint8_t buf[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;");
    size_t k = 0;
    for (size_t c = blockIdx.x; c < total_chunks; c += gridDim.x, k++) {
      const int s = (int)(k % depth);
      if (k >= (size_t)depth) {  // wait for the copy that used this slot
        uint32_t done = 0; const uint32_t par = (uint32_t)((k / depth - 1) & 1);
        while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(buf + (size_t)s * chunk)), "l"(src + c * stride_chunks * chunk), "r"(chunk), "r"(smem_u32(&bar[s])) : "memory");
    }
    for (size_t j = (k > (size_t)depth ? k - depth : 0); j < k; j++) {  // drain
      const int s = (int)(j % depth); uint32_t done = 0; const uint32_t par = (uint32_t)((j / depth) & 1);
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(par) : "memory");
    }
  }
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t bytes = 1ull << 30;  // 1 GiB source: nothing of it is in L2 when a run starts (L2 = 126 MB)
  uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pass = 0; pass < 2; pass++) {  // pass 0: HBM (cold, 512 MB streamed); pass 1: L2-resident (64 MB streamed twice, second run timed)
    for (int chunk : {4096, 16384, 32768}) for (int depth : {1, 2, 4, 8, 12}) {
      if ((size_t)chunk * depth > 196 * 1024) continue;
      const size_t vol = pass == 0 ? (512ull << 20) : (64ull << 20), n = vol / chunk;
      if (pass == 1) k_stream<<<sms, 32, chunk * depth>>>(src, n, chunk, depth, 1);
      cudaEventRecord(e0);
      k_stream<<<sms, 32, chunk * depth>>>(src + (pass == 0 ? 0 : 0), n, chunk, depth, 1);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("%s chunk %5d B x %2d in flight per SM (%3d KB): %7.1f GB/s total, %5.1f GB/s per SM, %s\n", pass == 0 ? "HBM" : "L2 ", chunk, depth, chunk * depth / 1024,
             vol / (ms * 1e-3) / 1e9, vol / (ms * 1e-3) / 1e9 / sms, cudaGetErrorString(cudaGetLastError()));
      if (pass == 0) cudaMemset(src, 1, 256 << 20);  // disturb L2 between HBM runs
    }
  }
  return 0;
}
