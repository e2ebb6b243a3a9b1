// tcgen05.mma throughput microbenchmark (sm_100a): cycles per MMA for A-in-smem (SS) vs A-in-TMEM
// (TS), bf16 vs tf32, several N.  One CTA per SM, one issuing thread, R MMAs back to back then a
// commit + mbarrier wait; operands are zeros (timing only).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_microbench mma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../human-3d-reconstruction_b200/csrc/ptx.cuh"
using namespace smplb200;

template <int N, bool TF32, bool TS, int GROUPS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 32 + N * 32) / 4 * 8; i += 128) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  constexpr uint32_t idesc = ptx::make_idesc(TF32 ? ptx::kFmtTF32 : ptx::kFmtBF16, 128, N);
  if (warp == 1 && lane == 0) {
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + 128 * 32 * 8;
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int g = 0; g < GROUPS; ++g) {
        const uint64_t bd = ptx::make_smem_desc(b_addr + (g & 7) * 2 * N * 16, N * 16, 128);
        if (TS) {
          if (TF32) ptx::mma_tf32_ts(tm, tm + 256 + (g & 7) * 8, bd, idesc, g > 0);
          else ptx::mma_bf16_ts(tm, tm + 256 + (g & 7) * 8, bd, idesc, g > 0);
        } else {
          const uint64_t ad = ptx::make_smem_desc(a_addr + (g & 7) * 2 * 2048, 2048, 128);
          if (TF32) ptx::mma_tf32(tm, ad, bd, idesc, g > 0);
          else ptx::mma_bf16(tm, ad, bd, idesc, g > 0);
        }
      }
      ptx::tc_commit(&bar);
      ptx::mbar_wait(&bar, phase); phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N, bool TF32, bool TS, int GROUPS>
void run(const char* name, long long* d_out) {
  const int reps = 200;
  const int smem = (128 * 32 + N * 32) * 8 + 1024;
  cudaFuncSetAttribute(k<N, TF32, TS, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TF32, TS, GROUPS><<<148, 128, smem>>>(d_out, 10);
  k<N, TF32, TS, GROUPS><<<148, 128, smem>>>(d_out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d groups=%2d : %7.1f clk/MMA  (floor N/2 = %d)  %s\n", name, N, GROUPS,
         (double)cyc / (reps * GROUPS), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  run<64, false, false, 14>("bf16 SS", d_out);  run<128, false, false, 14>("bf16 SS", d_out);  run<256, false, false, 14>("bf16 SS", d_out);
  run<64, false, true, 14>("bf16 TS", d_out);   run<128, false, true, 14>("bf16 TS", d_out);   run<256, false, true, 14>("bf16 TS", d_out);
  run<96, true, true, 9>("tf32 TS", d_out);     run<128, true, true, 28>("tf32 TS", d_out);    run<192, true, true, 9>("tf32 TS", d_out);
  run<128, true, false, 28>("tf32 SS", d_out);
  run<128, false, true, 42>("bf16 TS", d_out);  run<64, false, true, 42>("bf16 TS", d_out);   run<32, false, true, 42>("bf16 TS", d_out);
  return 0;
}
