// tcgen05.ld throughput microbenchmark (sm_100a): how many bytes per clock can the warps of one SM read
// out of tensor memory?  One CTA per SM, W warps (warp w reads its own lane quarter w % 4), each issuing
// R rounds of tcgen05.ld.32x32b.xN + wait::ld.  Reported: clk per round and TMEM->RF bytes/clk/SM.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_microbench tmem_ld_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../human-3d-reconstruction_b200/csrc/ptx.cuh"
using namespace smplb200;

template <int X, int BATCH>   // X = columns per instruction (16 or 32), BATCH = loads in flight before the wait
__global__ void __launch_bounds__(1024, 1) k(long long* out, int reps, int* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc(&slot, 512);
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const uint32_t col = (uint32_t)(((r * BATCH + b) * X + (warp >> 2) * 64) & 255);
      if (X == 16) { uint32_t v[16]; ptx::tmem_ld16(tm + col, v); acc ^= v[0] ^ v[15]; }
      else { uint32_t v[32]; ptx::tmem_ld32(tm + col, v); acc ^= v[0] ^ v[31]; }
    }
    ptx::tmem_ld_wait();
  }
  __syncthreads();
  long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(slot, 512);
}

template <int X, int BATCH>
void run(int warps, long long* d_out, int* d_sink) {
  const int reps = 2000;
  k<X, BATCH><<<148, warps * 32>>>(d_out, 50, d_sink);
  k<X, BATCH><<<148, warps * 32>>>(d_out, reps, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
  const double bytes = (double)reps * BATCH * X * 4 * 32 * warps;
  printf("ld.32x32b.x%-2d batch=%d warps=%2d : %7.1f clk/round  %7.1f B/clk/SM  %s\n", X, BATCH, warps,
         (double)cyc / reps, bytes / (double)cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out; int* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4);
  for (int w : {1, 4, 8, 16, 32}) { run<16, 1>(w, d_out, d_sink); run<16, 3>(w, d_out, d_sink); run<32, 1>(w, d_out, d_sink); run<32, 4>(w, d_out, d_sink); }
  return 0;
}
