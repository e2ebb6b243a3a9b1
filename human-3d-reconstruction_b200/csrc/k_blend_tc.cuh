// k1 (large-batch path): shape + pose blendshapes as ONE dense contraction on tcgen05 / TMEM.
//
//   vposed^T[col, b] = sum_{k<224} basis^T[col, k] * coef[b, k]          (col = planar column)
//
// with coef = [betas | pose_feature | 1 1 1 | 0...] so the v_template add is part of the
// contraction (K = 10 + 207 + 3 padded to 224 = 14 bf16 / 28 tf32 MMA k-steps).
//
// Orientation and operand placement
//   * the BASIS tile is the MMA "A" operand: M = 128 planar columns = the 128 TMEM lanes.  It is
//     kept RESIDENT IN TENSOR MEMORY (tcgen05.mma with A in TMEM): with A in shared memory every
//     MMA re-reads 128 x 32 B of A, which at N <= 64 exceeds the 128 B/clk shared-memory port
//     (round-1a ncu: tc pipe 65% busy for 21% math).  From TMEM the only shared-memory traffic
//     per MMA is the B operand (64 B/clk).
//   * a block of 128 bodies is the "B" operand (N = 128: 64 clk of tensor time per MMA, enough
//     to hide the ~40 clk the single issuing thread needs per tcgen05.mma), streamed one K half
//     per stage through a bulk-TMA/mbarrier ring from the K-major operand images k2 writes.
//   * an epilogue thread owns one planar column for 128 bodies, so each store instruction of a
//     warp is one fully coalesced 128-byte segment of the planar vposed[b, plane, v] layout.
//
// Scheduling: persistent CTA PAIRS (2-CTA clusters, one CTA per SM).  A pair works on two adjacent
// basis tiles and the same body blocks in lockstep; each CTA fetches half of every coef stage and
// TMA-MULTICASTS it to both (tcgen05.commit multicasts the "stage consumed" arrival back), which
// halves the operand re-read traffic through L2 (round-1 ncu: bf16x3 needed ~10 TB/s of L2->SM
// traffic, above the ~8 TB/s the chip sustained).  The (tile pair, body block) units are split
// into equal contiguous ranges, pair-major: no wave quantisation, few basis-tile switches.
//
// Warp roles (192 threads): warp 0 = bulk-TMA producer, warp 1 = single-thread MMA issuer,
// warps 2..5 = epilogue + basis loader (TMEM lane quarter = warp % 4).  All mbarrier waits use
// try_wait with a suspend-time hint (hardware sleep, no issue-slot polling).
//
// Precisions (operands; accumulation is always fp32 in TMEM):
//   BF16    1 MMA group   hi*hi
//   BF16X3  3 MMA groups  hi*hi + hi*lo + lo*hi   (~16 mantissa bits on each operand)
//   F16X3   the same three groups with fp16 operands (~22 bits each): fp32-class, measured < 1e-6 m
//   TF32    1 MMA group   kind::tf32
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kTcThreads = 192;
constexpr int kTcAccBufs = 2;                           // TMEM accumulator ring (2 x 128 columns)
constexpr int kTcAccCols = kTcAccBufs * kCoefBlock;     // 256
constexpr int kTcTmemCols = 512;

template <uint32_t PREC>
struct BlendTcCfg {
  static constexpr bool kTf32 = PREC == SMPLB200_PREC_TF32;
  static constexpr int kElem = kTf32 ? 4 : 2;
  static constexpr bool kF16 = PREC == SMPLB200_PREC_F16X3;
  static constexpr int kParts = (PREC == SMPLB200_PREC_BF16X3 || kF16) ? 2 : 1;   // hi (+ lo) operands
  static constexpr int kStages = 3;                                      // ring of K-half stages
  static constexpr int kKSteps = kCoefK * kElem / 32;                    // 14 (bf16) or 28 (tf32)
  static constexpr int kKHalf = kKSteps / 2;                             // MMA k-steps per stage
  static constexpr int kAColsPart = kCoefK * kElem / 4;                  // TMEM columns: 112 / 224
  static constexpr int kAWords = kAColsPart;                             // 32-bit words per basis row
  static constexpr uint32_t kBBytesPart = kCoefK * kCoefBlock * kElem;   // one coef image (block)
  static constexpr uint32_t kBHalf = kBBytesPart / 2;                    // its first / second K half
  static constexpr uint32_t kBStage = kBHalf * kParts;
  static constexpr uint32_t kBarOffset = kStages * kBStage;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256;
  static constexpr uint32_t kLboB = kCoefBlock * 16, kSbo = 128;
  static constexpr uint32_t kIdesc =
      ptx::make_idesc(kTf32 ? ptx::kFmtTF32 : (kF16 ? ptx::kFmtF16 : ptx::kFmtBF16), 128, kCoefBlock);
  static_assert(kTcAccCols + kAColsPart * kParts <= kTcTmemCols, "TMEM budget");
  static_assert(kKSteps % 2 == 0, "K halves");
};

template <uint32_t PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
k_blend_tc(const uint32_t* __restrict__ basis_hi, const uint32_t* __restrict__ basis_lo,
           const uint8_t* __restrict__ coef_hi, const uint8_t* __restrict__ coef_lo,
           long long n, int nblocks, long long total_units, int NC, float* __restrict__ vposed) {
  using C = BlendTcCfg<PREC>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOffset);
  uint64_t* bar_a = bars;                        // basis tile resident in TMEM (4 warp arrivals)
  uint64_t* bar_full = bars + 1;                 // [kStages] coef block landed
  uint64_t* bar_empty = bar_full + C::kStages;   // [kStages] MMAs reading the stage retired
  uint64_t* bar_tfull = bar_empty + C::kStages;  // [kTcAccBufs] accumulator ready
  uint64_t* bar_tempty = bar_tfull + kTcAccBufs; // [kTcAccBufs] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + kTcAccBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A cluster is a PAIR of CTAs working on two adjacent basis tiles and the same body blocks in
  // lockstep; `total_units` counts (tile pair, body block) units and every cluster takes an equal
  // contiguous share of the pair-major list.  Each CTA loads half of every coef stage and
  // multicasts it to both, so the operand re-read traffic through L2 is halved.
  const uint32_t crank = ptx::cluster_ctarank();
  const long long cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const long long u0 = total_units * cid / nclusters;
  const long long u1 = total_units * (cid + 1) / nclusters;
  const int nunits = (int)(u1 - u0);
  const int ntile = NC / 128;

  constexpr int kWarpTma = 0, kWarpMma = 1;
  if (warp == kWarpTma && lane == 0) {
    ptx::mbar_init(bar_a, 4);
    for (int s = 0; s < C::kStages; ++s) { ptx::mbar_init(bar_full + s, 1); ptx::mbar_init(bar_empty + s, 2); }
    for (int a = 0; a < kTcAccBufs; ++a) { ptx::mbar_init(bar_tfull + a, 1); ptx::mbar_init(bar_tempty + a, 4); }
    ptx::fence_barrier_init();
  }
  if (warp == kWarpMma) ptx::tmem_alloc(tmem_slot, kTcTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrival
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + kTcAccCols;   // A operand columns follow the accumulators

  if (warp == kWarpTma) {
    // ===== bulk-TMA producer: coef images of the body blocks, one K half per stage.  This CTA
    // fetches HALF of the stage and multicasts it to both CTAs of the pair. =====
    if (lane == 0) {
      constexpr uint32_t kMy = C::kBStage / 2;          // bytes this CTA fetches per stage
      for (int i = 0; i < 2 * nunits; ++i) {
        const int s = i % C::kStages;
        const int blk = (int)((u0 + (i >> 1)) % nblocks);
        ptx::mbar_wait(bar_empty + s, ((i / C::kStages) & 1) ^ 1);   // both CTAs' MMAs retired
        ptx::mbar_arrive_expect_tx(bar_full + s, C::kBStage);
        uint8_t* dst = sB + (size_t)s * C::kBStage;
        const size_t src = (size_t)blk * C::kBBytesPart + (size_t)(i & 1) * C::kBHalf;
        if (C::kParts == 2) {     // stage = [hi half-K image | lo half-K image]: rank 0 -> hi, rank 1 -> lo
          ptx::bulk_g2s_multicast(dst + crank * C::kBHalf, (crank ? coef_lo : coef_hi) + src, C::kBHalf,
                                  bar_full + s, (uint16_t)3);
        } else {                  // single image: each rank fetches half of its bytes
          ptx::bulk_g2s_multicast(dst + crank * kMy, coef_hi + src + crank * kMy, kMy, bar_full + s,
                                  (uint16_t)3);
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer: the whole warp runs the loop convergently and ONE elected lane issues.
    // Keeping control flow warp-uniform lets the descriptors live in uniform registers; issued from
    // a divergent `lane == 0` branch every tcgen05.mma cost ~15 SASS instructions (ELECT + 3x
    // R2UR.BROADCAST + predicate shuffling, ~50 clk), i.e. ~2100 of the 2688 clk a unit's MMAs take. =====
    {
      long long cur_tile = -1;
      uint32_t a_phase = 0;
      for (int i = 0; i < nunits; ++i) {
        const int a = i % kTcAccBufs;
        const long long tile = 2 * ((u0 + i) / nblocks) + crank;
        if (tile != cur_tile) {            // wait until the epilogue warps have (re)loaded A
          ptx::mbar_wait(bar_a, a_phase);
          a_phase ^= 1;
          cur_tile = tile;
        }
        ptx::mbar_wait(bar_tempty + a, ((i / kTcAccBufs) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + a * kCoefBlock;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const int st = 2 * i + kh, s = st % C::kStages;
          ptx::mbar_wait(bar_full + s, (st / C::kStages) & 1);
          ptx::tc_fence_after();
          const uint32_t b_addr = ptx::smem_u32(sB + (size_t)s * C::kBStage);
          constexpr int kGroups = C::kParts == 2 ? 3 : 1;   // (hi,hi) [, (hi,lo), (lo,hi)]
          if (ptx::elect_one()) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
              const uint32_t ap = tmem_a + (g == 2 ? C::kAColsPart : 0) + kh * C::kKHalf * 8;
              const uint32_t bp = b_addr + (g == 1 ? C::kBHalf : 0);
#pragma unroll
              for (int ks = 0; ks < C::kKHalf; ++ks) {
                const uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * C::kLboB, C::kLboB, C::kSbo);
                const uint32_t acc = (kh | g | ks) != 0;
                if (C::kTf32) ptx::mma_tf32_ts(d_tmem, ap + ks * 8, bd, C::kIdesc, acc);
                else ptx::mma_bf16_ts(d_tmem, ap + ks * 8, bd, C::kIdesc, acc);
              }
            }
            ptx::tc_commit_multicast(bar_empty + s, (uint16_t)3);   // tell BOTH producers: stage consumed here
            if (kh == 1) ptx::tc_commit(bar_tfull + a);             // accumulator ready for the epilogue
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue + basis loader =====
    const int q = warp & 3;                               // TMEM lane quarter of this warp
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long long cur_tile = -1;
    for (int i = 0; i < nunits; ++i) {
      const int a = i % kTcAccBufs;
      const long long tile = 2 * ((u0 + i) / nblocks) + crank;
      const bool live = tile < ntile;                     // odd tile count: the pair's second tile is a dummy
      const int blk = (int)((u0 + i) % nblocks);
      if (tile != cur_tile) {
        // Every earlier unit's accumulator was waited on below, so all MMAs that read the old
        // basis tile have retired: overwrite the A operand columns with the new tile's rows.
        cur_tile = tile;
        const size_t row = (size_t)(live ? tile : 0) * 128 + q * 32 + lane;
#pragma unroll
        for (int part = 0; part < C::kParts; ++part) {
          const uint4* src = reinterpret_cast<const uint4*>((part ? basis_lo : basis_hi) + row * C::kAWords);
#pragma unroll 7
          for (int c = 0; c < C::kAWords / 16; ++c) {
            uint32_t w[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 x = __ldg(src + c * 4 + v);
              w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
            ptx::tmem_st16(tmem_a + lane_addr + part * C::kAColsPart + c * 16, w);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_a);
        __syncwarp();
      }
      ptx::mbar_wait(bar_tfull + a, (i / kTcAccBufs) & 1);
      ptx::tc_fence_after();
      const long long b0 = (long long)blk * kCoefBlock;
      const int col = (int)tile * 128 + q * 32 + lane;    // planar column owned by this thread
      const int nb = (int)min((long long)kCoefBlock, n - b0);
      // whole accumulator -> registers, then hand the TMEM buffer straight back to the MMA warp
      uint32_t r[kCoefBlock];
#pragma unroll
      for (int c = 0; c < kCoefBlock / 32; ++c)
        ptx::tmem_ld32(tmem_base + lane_addr + a * kCoefBlock + c * 32,
                       *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_tempty + a);
      __syncwarp();
      // one coalesced 128-byte store per body row; four independent pointer chains, no predicates
      // on the full-block fast path (the naive indexed form cost ~18 SASS instructions per store)
      const size_t ld = (size_t)NC;
      float* p0 = vposed + (size_t)b0 * ld + col;
      if (!live) {
      } else if (nb == kCoefBlock) {
        float* p1 = p0 + ld; float* p2 = p1 + ld; float* p3 = p2 + ld;
        const size_t ld4 = 4 * ld;
#pragma unroll
        for (int j = 0; j < kCoefBlock; j += 4) {
          *p0 = __uint_as_float(r[j]);     p0 += ld4;
          *p1 = __uint_as_float(r[j + 1]); p1 += ld4;
          *p2 = __uint_as_float(r[j + 2]); p2 += ld4;
          *p3 = __uint_as_float(r[j + 3]); p3 += ld4;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kCoefBlock; ++j) {
          if (j < nb) *p0 = __uint_as_float(r[j]);
          p0 += ld;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();          // no CTA exits while its partner may still multicast into it
  if (warp == kWarpMma) ptx::tmem_dealloc(tmem_base, kTcTmemCols);
}

// fp32 coef [n,224] -> operand images (stand-alone k1 entry point only; the fused forward has
// k2 write the images directly).
__global__ void __launch_bounds__(256)
k_pack_coef(const float* __restrict__ coef, long long n, uint16_t* __restrict__ hi,
            uint16_t* __restrict__ lo, uint32_t* __restrict__ tf, int f16) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * kCoefK) return;
  const long long b = idx / kCoefK;
  const int k = int(idx - b * kCoefK);
  const float v = coef[idx];
  const long long blk = b / kCoefBlock;
  const int row = int(b % kCoefBlock);
  if (hi) {
    const size_t off = (size_t)blk * (kCoefK * kCoefBlock) + (size_t)(k >> 3) * (kCoefBlock * 8) + row * 8 + (k & 7);
    if (f16) {
      const __half h = __float2half_rn(v);
      hi[off] = __half_as_ushort(h);
      if (lo) lo[off] = __half_as_ushort(__float2half_rn(__fsub_rn(v, __half2float(h))));
    } else {
      const uint16_t h = f32_to_bf16_rn(v);
      hi[off] = h;
      if (lo) lo[off] = f32_to_bf16_rn(__fsub_rn(v, bf16_to_f32(h)));
    }
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
  const int npair = (ntile + 1) / 2;
  const int nblocks = (int)((n + kCoefBlock - 1) / kCoefBlock);
  const long long total = (long long)npair * nblocks;                 // (tile pair, body block) units
  const unsigned grid = 2u * (unsigned)std::min<long long>(num_sms / 2, total);   // CTA pairs
  const uint32_t* bh = C::kTf32 ? m.basis_rows_tf32 : (C::kF16 ? m.basis_rows_f16_hi : m.basis_rows_bf16_hi);
  k_blend_tc<PREC><<<grid, kTcThreads, C::kSmemBytes, s>>>(
      bh, C::kF16 ? m.basis_rows_f16_lo : m.basis_rows_bf16_lo, static_cast<const uint8_t*>(chi), static_cast<const uint8_t*>(clo), n,
      nblocks, total, m.NC, vposed);
}

inline cudaError_t launch_blend_tc(const DeviceModel& m, int num_sms, uint32_t prec,
                                   const uint16_t* chi, const uint16_t* clo, const uint32_t* ctf,
                                   long long n, float* vposed, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  switch (prec) {
    case SMPLB200_PREC_BF16: blend_tc_launch<SMPLB200_PREC_BF16>(m, num_sms, chi, nullptr, n, vposed, s); break;
    case SMPLB200_PREC_BF16X3: blend_tc_launch<SMPLB200_PREC_BF16X3>(m, num_sms, chi, clo, n, vposed, s); break;
    case SMPLB200_PREC_TF32: blend_tc_launch<SMPLB200_PREC_TF32>(m, num_sms, ctf, nullptr, n, vposed, s); break;
    case SMPLB200_PREC_F16X3: blend_tc_launch<SMPLB200_PREC_F16X3>(m, num_sms, chi, clo, n, vposed, s); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace smplb200
