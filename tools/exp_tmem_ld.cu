// Microbenchmark: TMEM -> register read bandwidth per SM for several tcgen05.ld shapes and warp counts.
// The scoring epilogue must read EVERY fp32 score out of TMEM (U x I x 4 bytes per sweep), so this rate is a hard
// ceiling of the single-product scoring kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o build/exp_tmem_ld tools/exp_tmem_ld.cu ; run on a B200: ./build/exp_tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SHAPE>
__device__ __forceinline__ uint32_t ld_chunk(uint32_t taddr) {
  uint32_t v[32];
  if (SHAPE == 0) {  // 32x32b.x32: 32 lanes x 32 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  } else if (SHAPE == 1) {  // 16x256b.x8: 16 lanes x 64 columns -> 32 registers (half the lanes of the quarter)
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  } else if (SHAPE == 2) {  // 16x128b.x16
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  } else {  // 32x32b.x32 with .pack::16b (two 16-bit values per register: what fp16 accumulators would allow)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t a = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a ^= v[i];
  return a;
}

// every warp reads `iters` chunks of 32 registers per thread from its lane quarter, cycling over the 512 columns
template <int SHAPE>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  const int cols_per_chunk = SHAPE == 1 ? 64 : (SHAPE == 2 ? 64 : 32);  // 16-lane shapes cover 16 lanes x 64 columns
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * cols_per_chunk + (warp >> 2) * 128) & 511) & ~(uint32_t)(cols_per_chunk - 1);
    acc ^= ld_chunk<SHAPE>(base + col);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

template <int SHAPE>
void run(const char* name, int warps, int iters) {
  unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 4);
  tmem_ld_kernel<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  tmem_ld_kernel<SHAPE><<<148, warps * 32>>>(iters, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double bytes = (double)warps * iters * 32 * 32 * 4;  // registers delivered per CTA
  printf("%-22s warps=%2d: %8.0f cycles, %6.1f B/clk/SM delivered (%s)\n", name, warps, avg, bytes / avg, cudaGetErrorString(e));
  cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
  const int iters = 4096;
  for (int w : {4, 8, 16}) {
    run<0>("32x32b.x32", w, iters);
    run<1>("16x256b.x8", w, iters);
    run<2>("16x128b.x16", w, iters);
  }
  return 0;
}
