// Multi-GPU exchange of the small per-body outputs (joints 288 B + kp2d 192 B per body).
//
// Bodies are sharded over the ranks with no data-path collective; the ONLY optional exchange is
// making every rank's joints/kp2d rows visible on every other rank.  Instead of an NCCL launch at the
// tail of each step, a rank PUSHES its rows: plain peer stores over NVLink / NVSwitch into the same
// row offset of a buffer that every rank has mapped (CUDA IPC / symmetric memory, set up once by the
// host side), then raises its flag in every peer's flag array.  `k_wait_rows` is the consumer side:
// it returns once all ranks' flags have reached the epoch.  Both run on a side stream that waits only
// on the "joints ready" event the forward records right after k2, so the exchange of step t overlaps
// the blendshape / skinning kernels of step t and never sits on the compute stream.
//
// Row layout of the gathered buffer: [n_total][120] fp32 = joints (72) | kp2d (48).
#pragma once
#include "common.cuh"

namespace smplb200 {

constexpr int kXchgMaxRanks = 16;
constexpr int kXchgRow = kJ * 3 + kJ * 2;      // 120 floats = 30 float4
constexpr int kXchgThreads = 128;

struct XchgPeers {
  float* buf[kXchgMaxRanks];                   // peer-mapped gathered buffers (entry `rank` = local)
  uint32_t* flags[kXchgMaxRanks];              // peer-mapped flag arrays, uint32[world] each
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Every rank: rows [row0, row0+n) of every peer's buffer <- local joints | kp2d; the last CTA to finish
// publishes `epoch` in flags[peer][rank] for every peer.  `done` is a local CTA counter (zero between launches).
__global__ void __launch_bounds__(kXchgThreads)
k_push_rows(XchgPeers peers, int world, int rank, const float* __restrict__ joints,
            const float* __restrict__ kp2d, long long n, long long row0, uint32_t epoch,
            unsigned int* done) {
  const long long total = n * (kXchgRow / 4);              // float4 items per destination
  const float4* j4 = reinterpret_cast<const float4*>(joints);
  const float4* k4 = reinterpret_cast<const float4*>(kp2d);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (kXchgRow / 4);
    const int q = (int)(i - b * (kXchgRow / 4));
    float4 v;
    if (q < kJ * 3 / 4) v = __ldg(j4 + b * (kJ * 3 / 4) + q);
    else if (kp2d) v = __ldg(k4 + b * (kJ * 2 / 4) + (q - kJ * 3 / 4));
    else v = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long dst = (row0 + b) * (kXchgRow / 4) + q;
#pragma unroll 1
    for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(peers.buf[p])[dst] = v;
  }
  // publish: all of this CTA's peer stores are ordered before its count; the last CTA's flag stores
  // are ordered after every count it observed (fence + release at system scope).
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    last = (prev + 1u == gridDim.x);
    if (last) *done = 0u;                      // every CTA has counted: ready for the next launch
  }
  __syncthreads();
  if (last && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
  }
}

// Consumer side: returns when flags[r] >= epoch for every rank r (wrap-safe compare).  Bounded: a
// peer that never arrives faults this launch after ~4 s instead of hanging the GPU.
__global__ void k_wait_rows(const uint32_t* __restrict__ my_flags, int world, uint32_t epoch) {
  if ((int)threadIdx.x >= world) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}

// Measurement aid (bench.py, SURVEY.md §8d "TF32 and FP32-FMA peaks measured by our bench in the same
// run"): 8 independent FMA chains per thread, `iters` rounds: 16 * iters flop per thread.
__global__ void __launch_bounds__(256)
k_probe_fma(float* __restrict__ sink, int iters, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a + (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678f) sink[0] = s;            // never true: keeps the chains alive
}

}  // namespace smplb200
