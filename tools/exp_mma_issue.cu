// Microbenchmark: what does ISSUING tcgen05.mma cost the issuing warp, and how fast does a chain of MMAs into ONE
// accumulator run compared with independent chains?  (Design input for the issuer of score_topk.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o build/exp_mma_issue tools/exp_mma_issue.cu
#include <cstdio>
#include "../gnn-recsys_b200/csrc/tc5.cuh"
namespace gr { void set_error(const std::string&) {} void count_launch() {} }
using namespace gr::tc5;

// warps [0, n_issuers) each issue `n_mma` MMAs (M = 128, N = 128, K = 16, cta_group::1) into `n_acc` accumulators in
// round-robin (their own set), then commit and wait. cycles[0] = issue loop of warp 0, cycles[1] = until completion.
__global__ void __launch_bounds__(128, 1) issue_kernel(int n_issuers, int n_acc, int n_mma, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 2 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
  const uint32_t idesc = make_idesc_mn(0u, 128, 128);
  const uint64_t da = make_desc_sw128(smem_u32(base)), db = make_desc_sw128(smem_u32(base + 16384));
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp < n_issuers) {
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t d = tmem + (uint32_t)(((warp * n_acc + (i + u) % n_acc) % 4) * 128);
          tc_mma_f16(d, da + (uint64_t)(2 * (u & 3)), db + (uint64_t)(2 * (u & 3)), idesc, 1u);
        }
      }
      __syncwarp();
    }
    t1 = clock64();
    if (elect_one()) tc_commit(bar + warp);
    __syncwarp();
    mbar_wait(bar + warp, 0);
    t2 = clock64();
  }
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int n_mma = 1024;
  for (int issuers : {1, 2, 4})
    for (int acc : {1, 2, 4}) {
      if (issuers * acc > 4 && acc > 1) continue;
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        issue_kernel<<<1, 128, 40000>>>(issuers, acc, n_mma, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep == 1)
          printf("issuers=%d accumulators/issuer=%d: issue loop %6.1f cyc/MMA, until done %6.1f cyc/MMA/issuer (%s)  [64 = pipe rate]\n",
                 issuers, acc, (double)h[0] / n_mma, (double)h[1] / n_mma, cudaGetErrorString(e));
      }
    }
  return 0;
}
