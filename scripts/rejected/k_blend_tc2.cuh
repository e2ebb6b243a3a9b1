// k1, 2-SM variant: the blendshape contraction with tcgen05 cta_group::2 MMAs (M = 256).
//
// Same math, operands and epilogue as k_blend_tc.cuh; what changes is WHO holds the B operand.
// Round-1 ablations showed the step is bound by the SM<->L2 crossbar: every SM had to ingest the
// whole coef stage (448..896 B per body) for its 128 output columns -- ~46 B/clk per SM for bf16x3
// before a single MMA or store -- and TMA multicast does not change what each SM ingests.  With a
// CTA pair issuing ONE MMA of M = 256 (tile 2p in CTA 0's TMEM lanes, tile 2p+1 in CTA 1's), each
// CTA holds only HALF of the B rows (64 of the block's 128 bodies) in its shared memory and the
// tensor cores read both halves, so the operand bytes each SM pulls through the crossbar halve.
//
// Protocol (cluster of 2, one CTA per SM, persistent over pair-major (tile pair, body block) units):
//   * both CTAs: warp 0 = bulk-TMA producer of ITS half of each K-half stage (one 1 KB chunk row
//     per lane), warps 2..5 = epilogue + loader of ITS basis tile into TMEM;
//   * leader (rank 0) warp 1 = MMA issuer: waits its own `full`, the peer's relayed `pfull`, the
//     accumulator-drained barrier (8 arrivals: 4 local + 4 remote epilogue warps), issues
//     tcgen05.mma.cta_group::2 and commits with multicast so BOTH CTAs see `empty` and `tfull`;
//   * peer (rank 1) warp 1 = relay: waits its local `full[s]` and arrives on the leader's `pfull[s]`.
#pragma once
#include "k_blend_tc.cuh"

namespace smplb200 {

template <uint32_t PREC>
struct BlendTc2Cfg {
  using C1 = BlendTcCfg<PREC>;
  static constexpr int kHalfRows = kCoefBlock / 2;                       // 64 bodies per CTA
  static constexpr int kChunksHalf = C1::kKHalf * 2;                     // 16-byte K chunks per K-half stage
  static constexpr uint32_t kChunkBytes = kHalfRows * 16;                // 1 KB: this CTA's rows of one chunk
  static constexpr uint32_t kStagePart = kChunksHalf * kChunkBytes;      // one part (hi or lo)
  static constexpr uint32_t kBStage = kStagePart * C1::kParts;
  static constexpr int kStages = 4;
  static constexpr uint32_t kBarOffset = kStages * kBStage;
  static constexpr uint32_t kSmemBytes = kBarOffset + 256;
  static constexpr uint32_t kLboB = kHalfRows * 16, kSbo = 128;
  static constexpr uint32_t kIdesc =
      ptx::make_idesc(C1::kTf32 ? ptx::kFmtTF32 : ptx::kFmtBF16, 256, kCoefBlock);
  static_assert(kChunksHalf * C1::kParts <= 32, "one chunk row per producer lane");
};

template <uint32_t PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
k_blend_tc2(const uint32_t* __restrict__ basis_hi, const uint32_t* __restrict__ basis_lo,
            const uint8_t* __restrict__ coef_hi, const uint8_t* __restrict__ coef_lo,
            long long n, int nblocks, long long total_units, int NC, float* __restrict__ vposed) {
  using C1 = BlendTcCfg<PREC>;
  using C = BlendTc2Cfg<PREC>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOffset);
  uint64_t* bar_a = bars;                        // leader: both basis tiles resident (8 arrivals)
  uint64_t* bar_full = bars + 1;                 // [stages] this CTA's half stage landed
  uint64_t* bar_pfull = bar_full + C::kStages;   // [stages] leader: the peer's half stage landed
  uint64_t* bar_empty = bar_pfull + C::kStages;  // [stages] MMAs reading the stage retired (multicast commit)
  uint64_t* bar_tfull = bar_empty + C::kStages;  // [acc] accumulator ready (multicast commit)
  uint64_t* bar_tempty = bar_tfull + kTcAccBufs; // [acc] leader: drained in BOTH CTAs (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + kTcAccBufs);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = ptx::cluster_ctarank();
  const bool leader = crank == 0;
  const long long cid = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const long long u0 = total_units * cid / nclusters;
  const long long u1 = total_units * (cid + 1) / nclusters;
  const int nunits = (int)(u1 - u0);
  const int ntile = NC / 128;
  constexpr int kWarpTma = 0, kWarpMma = 1;

  if (warp == kWarpTma && lane == 0) {
    ptx::mbar_init(bar_a, 8);
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(bar_full + s, 1); ptx::mbar_init(bar_pfull + s, 1); ptx::mbar_init(bar_empty + s, 1);
    }
    for (int a = 0; a < kTcAccBufs; ++a) { ptx::mbar_init(bar_tfull + a, 1); ptx::mbar_init(bar_tempty + a, 8); }
    ptx::fence_barrier_init();
  }
  if (warp == kWarpMma) ptx::tmem_alloc_2sm(tmem_slot, kTcTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();          // both CTAs' barriers and TMEM are set up before any remote arrival
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + kTcAccCols;

  if (warp == kWarpTma) {
    // ===== producer: this CTA's 64 rows of every 16-byte K chunk of the stage, one chunk per lane =====
    const int part = lane / C::kChunksHalf, ch = lane % C::kChunksHalf;
    const bool mine = lane < C::kChunksHalf * C1::kParts;
    for (int i = 0; i < 2 * nunits; ++i) {
      const int s = i % C::kStages;
      const int blk = (int)((u0 + (i >> 1)) % nblocks);
      ptx::mbar_wait(bar_empty + s, ((i / C::kStages) & 1) ^ 1);
      if (lane == 0) ptx::mbar_arrive_expect_tx(bar_full + s, C::kBStage);
      __syncwarp();
      if (mine) {
        // global image: [chunk][128 rows][16 B]; K half (i&1) starts at chunk (i&1)*kChunksHalf
        const uint8_t* img = (part ? coef_lo : coef_hi) + (size_t)blk * C1::kBBytesPart;
        const uint8_t* src = img + (size_t)((i & 1) * C::kChunksHalf + ch) * (kCoefBlock * 16) + crank * C::kChunkBytes;
        uint8_t* dst = sB + (size_t)s * C::kBStage + (size_t)part * C::kStagePart + (size_t)ch * C::kChunkBytes;
        ptx::bulk_g2s(dst, src, C::kChunkBytes, bar_full + s);
      }
      __syncwarp();
    }
  } else if (warp == kWarpMma) {
    if (lane == 0 && leader) {
      // ===== MMA issuer (leader CTA, one thread) =====
      long long cur_pair = -1;
      uint32_t a_phase = 0;
      for (int i = 0; i < nunits; ++i) {
        const int a = i % kTcAccBufs;
        const long long pair = (u0 + i) / nblocks;
        if (pair != cur_pair) { ptx::mbar_wait_cluster(bar_a, a_phase); a_phase ^= 1; cur_pair = pair; }
        ptx::mbar_wait_cluster(bar_tempty + a, ((i / kTcAccBufs) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + a * kCoefBlock;
        uint32_t acc = 0;
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {
          const int st = 2 * i + kh, s = st % C::kStages;
          ptx::mbar_wait(bar_full + s, (st / C::kStages) & 1);
          ptx::mbar_wait_cluster(bar_pfull + s, (st / C::kStages) & 1);
          ptx::tc_fence_after();
          const uint32_t b_addr = ptx::smem_u32(sB + (size_t)s * C::kBStage);
          constexpr int kGroups = C1::kParts == 2 ? 3 : 1;   // (hi,hi) [, (hi,lo), (lo,hi)]
#pragma unroll
          for (int g = 0; g < kGroups; ++g) {
            const uint32_t ap = tmem_a + (g == 2 ? C1::kAColsPart : 0) + kh * C1::kKHalf * 8;
            const uint32_t bp = b_addr + (g == 1 ? C::kStagePart : 0);
#pragma unroll
            for (int ks = 0; ks < C1::kKHalf; ++ks) {
              const uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * C::kLboB, C::kLboB, C::kSbo);
              if (C1::kTf32) ptx::mma_tf32_ts_2sm(d_tmem, ap + ks * 8, bd, C::kIdesc, acc);
              else ptx::mma_bf16_ts_2sm(d_tmem, ap + ks * 8, bd, C::kIdesc, acc);
              acc = 1;
            }
          }
          ptx::tc_commit_2sm(bar_empty + s, (uint16_t)3);   // both producers: stage consumed
        }
        ptx::tc_commit_2sm(bar_tfull + a, (uint16_t)3);     // both epilogues: accumulator ready
      }
    } else if (lane == 0) {
      // ===== relay (peer CTA): tell the leader when this CTA's half of a stage has landed =====
      for (int st = 0; st < 2 * nunits; ++st) {
        const int s = st % C::kStages;
        ptx::mbar_wait(bar_full + s, (st / C::kStages) & 1);
        ptx::mbar_arrive_remote(bar_pfull + s, 0);
      }
    }
  } else {
    // ===== epilogue + basis loader (both CTAs, own tile) =====
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long long cur_tile = -1;
    for (int i = 0; i < nunits; ++i) {
      const int a = i % kTcAccBufs;
      const long long tile = 2 * ((u0 + i) / nblocks) + crank;
      const bool live = tile < ntile;
      const int blk = (int)((u0 + i) % nblocks);
      if (tile != cur_tile) {
        cur_tile = tile;
        const size_t row = (size_t)(live ? tile : 0) * 128 + q * 32 + lane;
#pragma unroll
        for (int part = 0; part < C1::kParts; ++part) {
          const uint4* src = reinterpret_cast<const uint4*>((part ? basis_lo : basis_hi) + row * C1::kAWords);
#pragma unroll 7
          for (int c = 0; c < C1::kAWords / 16; ++c) {
            uint32_t w[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 x = __ldg(src + c * 4 + v);
              w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
            ptx::tmem_st16(tmem_a + lane_addr + part * C1::kAColsPart + c * 16, w);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) ptx::mbar_arrive(bar_a); else ptx::mbar_arrive_remote(bar_a, 0); }
        __syncwarp();
      }
      ptx::mbar_wait(bar_tfull + a, (i / kTcAccBufs) & 1);
      ptx::tc_fence_after();
      const long long b0 = (long long)blk * kCoefBlock;
      const int col = (int)tile * 128 + q * 32 + lane;
      const int nb = (int)min((long long)kCoefBlock, n - b0);
      uint32_t r[kCoefBlock];
#pragma unroll
      for (int c = 0; c < kCoefBlock / 32; ++c)
        ptx::tmem_ld32(tmem_base + lane_addr + a * kCoefBlock + c * 32,
                       *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) ptx::mbar_arrive(bar_tempty + a); else ptx::mbar_arrive_remote(bar_tempty + a, 0); }
      __syncwarp();
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
  ptx::cluster_sync_all();
  if (warp == kWarpMma) ptx::tmem_dealloc_2sm(tmem_base, kTcTmemCols);
}

template <uint32_t PREC>
inline cudaError_t blend_tc2_set_smem() {
  return cudaFuncSetAttribute(k_blend_tc2<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)BlendTc2Cfg<PREC>::kSmemBytes);
}

template <uint32_t PREC>
inline void blend_tc2_launch(const DeviceModel& m, int num_sms, const void* chi, const void* clo,
                             long long n, float* vposed, cudaStream_t s) {
  using C1 = BlendTcCfg<PREC>;
  using C = BlendTc2Cfg<PREC>;
  const int ntile = m.NC / 128;
  const int npair = (ntile + 1) / 2;
  const int nblocks = (int)((n + kCoefBlock - 1) / kCoefBlock);
  const long long total = (long long)npair * nblocks;
  const unsigned grid = 2u * (unsigned)std::min<long long>(num_sms / 2, total);
  const uint32_t* bh = C1::kTf32 ? m.basis_rows_tf32 : m.basis_rows_bf16_hi;
  k_blend_tc2<PREC><<<grid, kTcThreads, C::kSmemBytes, s>>>(
      bh, m.basis_rows_bf16_lo, static_cast<const uint8_t*>(chi), static_cast<const uint8_t*>(clo), n,
      nblocks, total, m.NC, vposed);
}

inline cudaError_t launch_blend_tc2(const DeviceModel& m, int num_sms, uint32_t prec,
                                    const uint16_t* chi, const uint16_t* clo, const uint32_t* ctf,
                                    long long n, float* vposed, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  switch (prec) {
    case SMPLB200_PREC_BF16: blend_tc2_launch<SMPLB200_PREC_BF16>(m, num_sms, chi, nullptr, n, vposed, s); break;
    case SMPLB200_PREC_BF16X3: blend_tc2_launch<SMPLB200_PREC_BF16X3>(m, num_sms, chi, clo, n, vposed, s); break;
    case SMPLB200_PREC_TF32: blend_tc2_launch<SMPLB200_PREC_TF32>(m, num_sms, ctf, nullptr, n, vposed, s); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace smplb200
