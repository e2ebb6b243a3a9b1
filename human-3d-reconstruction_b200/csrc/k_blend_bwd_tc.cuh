// kb1 (large-batch path): blendshape backward  g_coef[b,k] = sum_col g_vposed[b,col] * basis[k,col]
// on tcgen05 / TMEM.  The reduction runs over the 20,736 planar columns, the output is tiny
// ([n, 217]), so the kernel is a stream of g_vposed (83 KB/body, read exactly once) against the
// L2-resident basis:
//
//   D[M = 128 bodies, N = 224 k] += A[128 bodies, 32 cols] * B[224 k, 32 cols]^T      per K-step
//
//   * A (the gradient rows) goes through REGISTERS INTO TENSOR MEMORY: a loader thread owns one
//     body row, reads one full 128-byte line of it per K-step (perfect sector use although the
//     rows are 83 KB apart), rounds to tf32 (and, for 3xTF32, splits off the low part) and writes
//     its TMEM lane with tcgen05.st.  No layout transform in shared memory, no tensor map.
//   * B (the basis) is pre-tiled at model create into K-major no-swizzle images, one contiguous
//     28,672-byte tile [8 chunks][224 rows][4 tf32] per K-step, landed by ONE bulk-TMA copy.
//   * the accumulator D (224 fp32 columns) stays in TMEM for the CTA's whole column slice; the
//     epilogue writes one [128, 224] partial per (slice, body block); kb2 adds the slices in order.
//
// Precision: 1xTF32 for SMPLB200_PREC_TF32 / _BF16; split-bf16 (A_hi B_hi + A_lo B_hi + A_hi B_lo,
// kind::f16, K = 16 per MMA: half the MMAs and half the basis bytes of 3xTF32, ~2^-16 relative) for
// SMPLB200_PREC_BF16X3 / AUTO; 3xTF32 (~2^-21) for SMPLB200_PREC_FP32 at >= 256 bodies.
//
// Warp roles (448 threads): warp 0 = bulk-TMA producer of B, warp 1 = MMA issuer (warp-uniform,
// one elected lane), warps 2..13 = three groups of four A-loader warps (TMEM lane quarter = warp % 4;
// group g takes K-steps i = g mod 3, each thread keeps the NEXT K-step's line in flight while it
// converts the current one); warps 2..5 then run the epilogue.
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kBwdTcGroups = 3;              // loader groups (4 warps each)
constexpr int kBwdTcThreads = (2 + 4 * kBwdTcGroups) * 32;   // 448
constexpr int kBwdTcBodies = 128;
constexpr int kBwdTcStepCols = 32;          // planar columns per K-step (4 MMAs of K = 8)
constexpr uint32_t kBwdTcTile = 8u * kCoefK * 16u;   // one basis K-step tile: 28,672 B

enum : int { kBwdTf32 = 0, kBwdTf32x3 = 1, kBwdBf16x3 = 2 };

template <int MODE>
struct BlendBwdTcCfg {
  static constexpr bool kBf16 = MODE == kBwdBf16x3;
  static constexpr bool kSplit = MODE != kBwdTf32;                  // hi + lo parts of both operands
  // one part of one K-step of B: tf32 [8 chunks][224][4] = 28,672 B, bf16 [4 chunks][224][8] = 14,336 B
  static constexpr uint32_t kBPart = kBf16 ? kBwdTcTile / 2 : kBwdTcTile;
  static constexpr uint32_t kBStage = kBPart * (kSplit ? 2u : 1u);
  static constexpr int kStagesB = MODE == kBwdTf32x3 ? 3 : 6;
  static constexpr int kAPartCols = kBf16 ? 16 : 32;                 // TMEM columns of one part of a K-step
  static constexpr int kAStageCols = kAPartCols * (kSplit ? 2 : 1);
  static constexpr int kStagesA = MODE == kBwdTf32x3 ? 4 : 8;
  static constexpr int kMmaPerPart = kBf16 ? 2 : 4;                  // K = 16 bf16 / 8 tf32 per MMA, 32 columns
  static constexpr int kACol0 = kCoefK;                  // A stages follow the 224 accumulator columns
  static constexpr uint32_t kBarOffset = kStagesB * kBStage;
  static constexpr uint32_t kSmemBytes = kBarOffset + 512;
  static constexpr uint32_t kLbo = kCoefK * 16, kSbo = 128;
  static constexpr uint32_t kIdesc = ptx::make_idesc(kBf16 ? ptx::kFmtBF16 : ptx::kFmtTF32, 128, kCoefK);
  static_assert(kACol0 + kStagesA * kAStageCols <= 512, "TMEM budget");
};

template <int MODE>
__global__ void __launch_bounds__(kBwdTcThreads, 1)
k_blend_bwd_tc(const uint8_t* __restrict__ bimg_hi, const uint8_t* __restrict__ bimg_lo,
               const float* __restrict__ g_vposed, long long n, int NC, int slices,
               float* __restrict__ part /* [slices][n][224] */) {
  using C = BlendBwdTcCfg<MODE>;
  constexpr bool X3 = C::kSplit;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOffset);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + C::kStagesB;
  uint64_t* a_full = b_empty + C::kStagesB;
  uint64_t* a_empty = a_full + C::kStagesA;
  uint64_t* d_full = a_empty + C::kStagesA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x;
  const long long b0 = (long long)blockIdx.y * kBwdTcBodies;
  const int nks_total = NC / kBwdTcStepCols;
  const int ks0 = (int)((long long)slice * nks_total / slices);
  const int nks = (int)((long long)(slice + 1) * nks_total / slices) - ks0;   // >= 1 (slices <= nks_total)

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::kStagesB; ++s) { ptx::mbar_init(b_full + s, 1); ptx::mbar_init(b_empty + s, 1); }
    for (int s = 0; s < C::kStagesA; ++s) { ptx::mbar_init(a_full + s, 4); ptx::mbar_init(a_empty + s, 1); }
    ptx::mbar_init(d_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== bulk-TMA producer: one basis tile (hi [+ lo]) per K-step =====
    if (lane == 0) {
      for (int i = 0; i < nks; ++i) {
        const int s = i % C::kStagesB;
        ptx::mbar_wait(b_empty + s, ((i / C::kStagesB) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(b_full + s, C::kBStage);
        uint8_t* dst = sB + (size_t)s * C::kBStage;
        const size_t src = (size_t)(ks0 + i) * C::kBPart;
        ptx::bulk_g2s(dst, bimg_hi + src, C::kBPart, b_full + s);
        if (X3) ptx::bulk_g2s(dst + C::kBPart, bimg_lo + src, C::kBPart, b_full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow, one elected lane issues) =====
    for (int i = 0; i < nks; ++i) {
      const int sb = i % C::kStagesB, sa = i % C::kStagesA;
      ptx::mbar_wait(b_full + sb, (i / C::kStagesB) & 1);
      ptx::mbar_wait(a_full + sa, (i / C::kStagesA) & 1);
      ptx::tc_fence_after();
      const uint32_t b_addr = ptx::smem_u32(sB + (size_t)sb * C::kBStage);
      const uint32_t a_addr = tmem_base + C::kACol0 + sa * C::kAStageCols;
      if (ptx::elect_one()) {
        constexpr int kGroups = X3 ? 3 : 1;          // (A_hi,B_hi) [, (A_lo,B_hi), (A_hi,B_lo)]
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const uint32_t ap = a_addr + (g == 1 ? C::kAPartCols : 0);
          const uint32_t bp = b_addr + (g == 2 ? C::kBPart : 0);
#pragma unroll
          for (int kk = 0; kk < C::kMmaPerPart; ++kk) {
            const uint64_t bd = ptx::make_smem_desc(bp + kk * 2 * C::kLbo, C::kLbo, C::kSbo);
            const uint32_t acc = (uint32_t)((i | g | kk) != 0);
            if (C::kBf16) ptx::mma_bf16_ts(tmem_base, ap + kk * 8, bd, C::kIdesc, acc);
            else ptx::mma_tf32_ts(tmem_base, ap + kk * 8, bd, C::kIdesc, acc);
          }
        }
        ptx::tc_commit(b_empty + sb);
        ptx::tc_commit(a_empty + sa);
        if (i == nks - 1) ptx::tc_commit(d_full);
      }
      __syncwarp();
    }
  } else {
    // ===== A loaders: global row -> registers -> (tf32 split) -> TMEM lane =====
    const int lw = warp - 2, q = warp & 3, grp = lw >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const long long row = b0 + q * 32 + lane;
    const bool valid = row < n;
    const float4* src = reinterpret_cast<const float4*>(
        g_vposed + (size_t)(valid ? row : 0) * NC + (size_t)ks0 * kBwdTcStepCols);
    float4 cur[8], nxt[8];
    auto load = [&](float4 (&dst)[8], int i) {
#pragma unroll
      for (int v = 0; v < 8; ++v)
        dst[v] = valid ? __ldg(src + (size_t)i * 8 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (grp < nks) load(cur, grp);
    for (int i = grp; i < nks; i += kBwdTcGroups) {
      if (i + kBwdTcGroups < nks) load(nxt, i + kBwdTcGroups);
      const int sa = i % C::kStagesA;
      ptx::mbar_wait(a_empty + sa, ((i / C::kStagesA) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t acol = tmem_base + lane_addr + C::kACol0 + sa * C::kAStageCols;
      if (C::kBf16) {
        // 32 columns -> 16 words of packed bf16 pairs per part (element 2w in the low half)
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 x = cur[v];
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint16_t h0 = f32_to_bf16_rn(xs[2 * e]), h1 = f32_to_bf16_rn(xs[2 * e + 1]);
            hi[2 * v + e] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lo[2 * v + e] = (uint32_t)f32_to_bf16_rn(xs[2 * e] - bf16_to_f32(h0)) |
                            ((uint32_t)f32_to_bf16_rn(xs[2 * e + 1] - bf16_to_f32(h1)) << 16);
          }
        }
        ptx::tmem_st16(acol, hi);
        ptx::tmem_st16(acol + 16, lo);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 x = cur[4 * h + v];
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[4 * v + e] = f32_to_tf32_rn(xs[e]);
              if (X3) lo[4 * v + e] = f32_to_tf32_rn(xs[e] - __uint_as_float(hi[4 * v + e]));
            }
          }
          ptx::tmem_st16(acol + 16 * h, hi);
          if (X3) ptx::tmem_st16(acol + 32 + 16 * h, lo);
        }
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_full + sa);
      __syncwarp();
#pragma unroll
      for (int v = 0; v < 8; ++v) cur[v] = nxt[v];
    }
    if (lw < 4) {
      // ===== epilogue: accumulator row (this body) -> partial[slice][row][0..223] =====
      ptx::mbar_wait(d_full, 0);
      ptx::tc_fence_after();
      float4* dst = reinterpret_cast<float4*>(part + ((size_t)slice * n + (size_t)(valid ? row : 0)) * kCoefK);
#pragma unroll 1
      for (int c = 0; c < kCoefK / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + lane_addr + c * 32, r);
        ptx::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v = 0; v < 8; ++v)
            dst[c * 8 + v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                         __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

template <int MODE>
inline cudaError_t blend_bwd_tc_set_smem() {
  return cudaFuncSetAttribute(k_blend_bwd_tc<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)BlendBwdTcCfg<MODE>::kSmemBytes);
}

inline cudaError_t launch_blend_bwd_tc(const DeviceModel& m, int mode, const float* g_vposed, long long n,
                                       int slices, float* part, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  dim3 grid((unsigned)slices, (unsigned)((n + kBwdTcBodies - 1) / kBwdTcBodies));
  const uint8_t* hi = reinterpret_cast<const uint8_t*>(m.bwd_basis_tf32_hi);
  const uint8_t* lo = reinterpret_cast<const uint8_t*>(m.bwd_basis_tf32_lo);
  const uint8_t* bhi = reinterpret_cast<const uint8_t*>(m.bwd_basis_bf16_hi);
  const uint8_t* blo = reinterpret_cast<const uint8_t*>(m.bwd_basis_bf16_lo);
  if (mode == kBwdBf16x3)
    k_blend_bwd_tc<kBwdBf16x3><<<grid, kBwdTcThreads, BlendBwdTcCfg<kBwdBf16x3>::kSmemBytes, s>>>(bhi, blo, g_vposed, n, m.NC, slices, part);
  else if (mode == kBwdTf32x3)
    k_blend_bwd_tc<kBwdTf32x3><<<grid, kBwdTcThreads, BlendBwdTcCfg<kBwdTf32x3>::kSmemBytes, s>>>(hi, lo, g_vposed, n, m.NC, slices, part);
  else
    k_blend_bwd_tc<kBwdTf32><<<grid, kBwdTcThreads, BlendBwdTcCfg<kBwdTf32>::kSmemBytes, s>>>(hi, lo, g_vposed, n, m.NC, slices, part);
  return cudaGetLastError();
}

}  // namespace smplb200
