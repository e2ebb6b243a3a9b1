// tcgen05.mma cadence with WARP-UNIFORM issue (the form the product kernels use: the whole warp runs the
// loop, one elected lane issues, descriptors live in uniform registers), kind::f16, M = 128, for the
// operand forms and N the fused blendshape+skinning kernel needs: SS (A and B from shared memory) at
// N = 64 / 128 and TS (A in TMEM) at N = 48 / 64 / 96.  One CTA per SM; G MMAs back to back, commit, wait.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mma_issue_microbench mma_issue_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../human-3d-reconstruction_b200/csrc/ptx.cuh"
using namespace smplb200;

template <int N, bool TS, int G>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 32 * 8 + N * 32 * 8) / 4; i += 128) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  constexpr uint32_t idesc = ptx::make_idesc(0u /* f16 */, 128, N);
  if (warp == 1) {
    const uint32_t a_addr = ptx::smem_u32(smem), b_addr = a_addr + 128 * 32 * 8;
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int r = 0; r < reps; ++r) {
      if (ptx::elect_one()) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint64_t bd = ptx::make_smem_desc(b_addr + (g & 7) * 2 * N * 16, N * 16, 128);
          if (TS) ptx::mma_bf16_ts(tm + (g & 1) * 128, tm + 256 + (g & 7) * 8, bd, idesc, g > 1);
          else ptx::mma_bf16(tm + (g & 1) * 128, ptx::make_smem_desc(a_addr + (g & 7) * 2 * 2048, 2048, 128), bd, idesc, g > 1);
        }
        ptx::tc_commit(&bar);
      }
      __syncwarp();
      ptx::mbar_wait(&bar, phase); phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N, bool TS, int G>
void run(long long* d_out) {
  const int reps = 300;
  const int smem = (128 * 32 + N * 32) * 8 + 1024;
  cudaFuncSetAttribute(k<N, TS, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<N, TS, G><<<148, 128, smem>>>(d_out, 10);
  k<N, TS, G><<<148, 128, smem>>>(d_out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  printf("f16 %s N=%3d  %2d MMAs per commit : %7.1f clk/MMA  (math floor N/2 = %d)  %s\n", TS ? "TS" : "SS", N, G,
         (double)cyc / (reps * G), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  run<48, true, 6>(d_out);  run<48, true, 24>(d_out);  run<64, true, 24>(d_out);  run<96, true, 6>(d_out);  run<96, true, 24>(d_out);
  run<128, true, 24>(d_out);
  run<48, false, 24>(d_out); run<64, false, 16>(d_out); run<64, false, 48>(d_out); run<96, false, 48>(d_out); run<128, false, 48>(d_out);
  run<16, true, 48>(d_out); run<32, true, 48>(d_out);
  return 0;
}
