// k1 (large-batch path): shape + pose blendshapes as ONE dense contraction on tcgen05 / TMEM.
//
//   vposed^T[col, b] = sum_{k<224} basis^T[col, k] * coef[b, k]          (col = planar column)
//
// with coef = [betas | pose_feature | 1 | 0...] so the v_template add is a row of the contraction
// (K = 10 + 207 + 1 padded to 224 = 14 bf16 / 28 tf32 MMA k-steps).
//
// Orientation: the BASIS tile is the MMA "A" operand (M = 128 planar columns = 128 TMEM lanes)
// and a block of 32 bodies is the "B" operand (N = 32 TMEM columns).  Reasons:
//   * the basis tile (57 KB bf16 / 115 KB split-bf16 or tf32) is loaded ONCE per CTA and stays
//     resident in shared memory while body blocks stream through a TMA/mbarrier ring;
//   * an epilogue thread owns one planar column for 32 bodies, so each of its stores is a fully
//     coalesced 128-byte row segment of the planar vposed[b, plane, v] layout -- no shared-memory
//     transpose and no TMA-store alignment constraints in the epilogue.
//
// Warp roles (192 threads): warp 0 = bulk-TMA producer, warp 1 = single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  Three mbarrier pipelines: smem
// full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), and one barrier for the basis.
//
// Precisions (operands; accumulation is always fp32 in TMEM):
//   BF16    1 MMA group   hi*hi
//   BF16X3  3 MMA groups  hi*hi + hi*lo + lo*hi   (~16 mantissa bits on each operand)
//   TF32    1 MMA group   kind::tf32
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kTcThreads = 192;
constexpr int kTcAccBufs = 4;                  // TMEM accumulator ring (4 x 32 columns)
constexpr int kTcTmemCols = kTcAccBufs * kCoefBlock;  // 128

template <uint32_t PREC>
struct BlendTcCfg {
  static constexpr bool kTf32 = PREC == SMPLB200_PREC_TF32;
  static constexpr int kElem = kTf32 ? 4 : 2;
  static constexpr int kParts = PREC == SMPLB200_PREC_BF16X3 ? 2 : 1;   // hi (+ lo) images
  static constexpr int kStages = PREC == SMPLB200_PREC_BF16 ? 4 : 3;
  static constexpr int kChunkElems = 16 / kElem;                         // K elements per 16 B
  static constexpr int kChunks = kCoefK / kChunkElems;                   // 28 or 56
  static constexpr int kKSteps = kChunks / 2;                            // 14 or 28
  static constexpr uint32_t kABytesPart = kCoefK * 128 * kElem;          // one basis image
  static constexpr uint32_t kBBytesPart = kCoefK * kCoefBlock * kElem;   // one coef image
  static constexpr uint32_t kABytes = kABytesPart * kParts;
  static constexpr uint32_t kBStage = kBBytesPart * kParts;
  static constexpr uint32_t kBarOffset = kABytes + kStages * kBStage;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256;
  static constexpr uint32_t kLboA = 128 * 16, kLboB = kCoefBlock * 16, kSbo = 128;
  static constexpr uint32_t kIdesc =
      ptx::make_idesc(kTf32 ? ptx::kFmtTF32 : ptx::kFmtBF16, 128, kCoefBlock);
};

template <uint32_t PREC>
__global__ void __launch_bounds__(kTcThreads, 1)
k_blend_tc(const uint8_t* __restrict__ basis_hi, const uint8_t* __restrict__ basis_lo,
           const uint8_t* __restrict__ coef_hi, const uint8_t* __restrict__ coef_lo,
           long long n, int nblocks, int blocks_per_cta, int NC, float* __restrict__ vposed) {
  using C = BlendTcCfg<PREC>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOffset);
  uint64_t* bar_a = bars;                        // basis tile landed
  uint64_t* bar_full = bars + 1;                 // [kStages] coef block landed
  uint64_t* bar_empty = bar_full + C::kStages;   // [kStages] MMAs reading the stage retired
  uint64_t* bar_tfull = bar_empty + C::kStages;  // [kTcAccBufs] accumulator ready
  uint64_t* bar_tempty = bar_tfull + kTcAccBufs; // [kTcAccBufs] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + kTcAccBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int blk_begin = blockIdx.y * blocks_per_cta;
  const int blk_end = min(nblocks, blk_begin + blocks_per_cta);
  const int nblk = blk_end - blk_begin;

  if (warp == 0 && lane == 0) {
    ptx::mbar_init(bar_a, 1);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(bar_full + s, 1); ptx::mbar_init(bar_empty + s, 1); }
    for (int a = 0; a < kTcAccBufs; ++a) { ptx::mbar_init(bar_tfull + a, 1); ptx::mbar_init(bar_tempty + a, 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, kTcTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== bulk-TMA producer =====
    if (lane == 0 && nblk > 0) {
      ptx::mbar_arrive_expect_tx(bar_a, C::kABytes);
      ptx::bulk_g2s_split(sA, basis_hi + (size_t)tile * C::kABytesPart, C::kABytesPart, bar_a);
      if (C::kParts == 2)
        ptx::bulk_g2s_split(sA + C::kABytesPart, basis_lo + (size_t)tile * C::kABytesPart,
                            C::kABytesPart, bar_a);
      for (int i = 0; i < nblk; ++i) {
        const int s = i % C::kStages;
        ptx::mbar_wait(bar_empty + s, ((i / C::kStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(bar_full + s, C::kBStage);
        uint8_t* dst = sB + (size_t)s * C::kBStage;
        const size_t src = (size_t)(blk_begin + i) * C::kBBytesPart;
        ptx::bulk_g2s(dst, coef_hi + src, C::kBBytesPart, bar_full + s);
        if (C::kParts == 2) ptx::bulk_g2s(dst + C::kBBytesPart, coef_lo + src, C::kBBytesPart, bar_full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && nblk > 0) {
      ptx::mbar_wait(bar_a, 0);
      const uint32_t a_addr = ptx::smem_u32(sA);
      for (int i = 0; i < nblk; ++i) {
        const int s = i % C::kStages, a = i % kTcAccBufs;
        ptx::mbar_wait(bar_tempty + a, ((i / kTcAccBufs) & 1) ^ 1);
        ptx::mbar_wait(bar_full + s, (i / C::kStages) & 1);
        ptx::tc_fence_after();
        const uint32_t b_addr = ptx::smem_u32(sB + (size_t)s * C::kBStage);
        const uint32_t d_tmem = tmem_base + a * kCoefBlock;
        uint32_t acc = 0;
        // MMA groups: (A part, B part) = (hi,hi) [, (hi,lo), (lo,hi)]
        constexpr int kGroups = C::kParts == 2 ? 3 : 1;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          const uint32_t ap = a_addr + (g == 2 ? C::kABytesPart : 0);
          const uint32_t bp = b_addr + (g == 1 ? C::kBBytesPart : 0);
#pragma unroll
          for (int ks = 0; ks < C::kKSteps; ++ks) {
            const uint64_t ad = ptx::make_smem_desc(ap + ks * 2 * C::kLboA, C::kLboA, C::kSbo);
            const uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * C::kLboB, C::kLboB, C::kSbo);
            if (C::kTf32) ptx::mma_tf32(d_tmem, ad, bd, C::kIdesc, acc);
            else ptx::mma_bf16(d_tmem, ad, bd, C::kIdesc, acc);
            acc = 1;
          }
        }
        ptx::tc_commit(bar_empty + s);   // stage reusable once these MMAs retire
        ptx::tc_commit(bar_tfull + a);   // accumulator ready for the epilogue
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> coalesced planar stores =====
    const int q = warp & 3;                               // TMEM lane quarter of this warp
    const int col = tile * 128 + q * 32 + lane;           // planar column owned by this thread
    for (int i = 0; i < nblk; ++i) {
      const int a = i % kTcAccBufs;
      ptx::mbar_wait(bar_tfull + a, (i / kTcAccBufs) & 1);
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * kCoefBlock, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + a);
      __syncwarp();
      const long long b0 = (long long)(blk_begin + i) * kCoefBlock;
      float* dst = vposed + (size_t)b0 * NC + col;
      const int nb = (int)min((long long)kCoefBlock, n - b0);
#pragma unroll
      for (int j = 0; j < kCoefBlock; ++j)
        if (j < nb) dst[(size_t)j * NC] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTcTmemCols);
}

// fp32 coef [n,224] -> operand images (stand-alone k1 entry point only; the fused forward has
// k2 write the images directly).
__global__ void __launch_bounds__(256)
k_pack_coef(const float* __restrict__ coef, long long n, uint16_t* __restrict__ hi,
            uint16_t* __restrict__ lo, uint32_t* __restrict__ tf) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * kCoefK) return;
  const long long b = idx / kCoefK;
  const int k = int(idx - b * kCoefK);
  const float v = coef[idx];
  const long long blk = b / kCoefBlock;
  const int row = int(b % kCoefBlock);
  if (hi) {
    const uint16_t h = f32_to_bf16_rn(v);
    const size_t off = (size_t)blk * (kCoefK * kCoefBlock) + (size_t)(k >> 3) * (kCoefBlock * 8) + row * 8 + (k & 7);
    hi[off] = h;
    if (lo) lo[off] = f32_to_bf16_rn(__fsub_rn(v, bf16_to_f32(h)));
  }
  if (tf)
    tf[(size_t)blk * (kCoefK * kCoefBlock) + (size_t)(k >> 2) * (kCoefBlock * 4) + row * 4 + (k & 3)] =
        f32_to_tf32_rn(v);
}

template <uint32_t PREC>
inline cudaError_t blend_tc_set_smem() {
  return cudaFuncSetAttribute(k_blend_tc<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)BlendTcCfg<PREC>::kSmemBytes);
}

template <uint32_t PREC>
inline void blend_tc_launch(const DeviceModel& m, int num_sms, const void* chi, const void* clo,
                            long long n, float* vposed, cudaStream_t s) {
  using C = BlendTcCfg<PREC>;
  const int ntile = m.NC / 128;
  const int nblocks = (int)((n + kCoefBlock - 1) / kCoefBlock);
  // amortise the resident basis tile over >= 8 body blocks while keeping >= ~4 CTAs per SM queued
  int bpc = 8;
  while (bpc < nblocks && (long long)ntile * ((nblocks + bpc - 1) / bpc) > 8LL * num_sms) bpc *= 2;
  if (bpc > nblocks) bpc = nblocks;
  const dim3 grid((unsigned)ntile, (unsigned)((nblocks + bpc - 1) / bpc));
  const uint8_t* bh = reinterpret_cast<const uint8_t*>(C::kTf32 ? (const void*)m.basis_tf32 : (const void*)m.basis_bf16_hi);
  const uint8_t* bl = reinterpret_cast<const uint8_t*>(m.basis_bf16_lo);
  k_blend_tc<PREC><<<grid, kTcThreads, C::kSmemBytes, s>>>(
      bh, bl, static_cast<const uint8_t*>(chi), static_cast<const uint8_t*>(clo), n, nblocks, bpc,
      m.NC, vposed);
}

inline cudaError_t launch_blend_tc(const DeviceModel& m, int num_sms, uint32_t prec,
                                   const uint16_t* chi, const uint16_t* clo, const uint32_t* ctf,
                                   long long n, float* vposed, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  switch (prec) {
    case SMPLB200_PREC_BF16: blend_tc_launch<SMPLB200_PREC_BF16>(m, num_sms, chi, nullptr, n, vposed, s); break;
    case SMPLB200_PREC_BF16X3: blend_tc_launch<SMPLB200_PREC_BF16X3>(m, num_sms, chi, clo, n, vposed, s); break;
    case SMPLB200_PREC_TF32: blend_tc_launch<SMPLB200_PREC_TF32>(m, num_sms, ctf, nullptr, n, vposed, s); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace smplb200
