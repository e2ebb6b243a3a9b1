// k3 (+k4) tensor-core path: linear blend skinning with the 24-joint blend on tcgen05 / TMEM.
//
//   T[v, (b,e)] = sum_j W[v,j] * A[b,j,e]          e = 12 entries of the 3x4 joint transform
//   verts[b,v,:] = T[v,(b,:)] (3x4) * [vposed[b,v,:]; 1]
//
// Why tensor cores for a "memory-bound" op: dense 24-joint fp32 blending is FMA-bound on the CUDA
// cores (4.1 Mflop per body needs ~97 TFLOP/s at 60% of HBM), and even
// the <=4-nnz form is bound by the shared-memory crossbar (48 gathered floats per vertex-body).
// As an MMA the blend is M = 128 vertices (TMEM lanes), N = 12 * 8 bodies, K = 24 joints, and
// costs ~0.42 clk per vertex-body per SM for ANY weight matrix -- no sparsity assumption.
//
// fp32 fidelity comes from the 3xTF32 split (both operands split exactly into tf32 hi + lo):
//   W*A ~= W_hi*A_hi + W_hi*A_lo + W_lo*A_hi      (dropped lo*lo term ~2^-22 relative)
// tf32 x tf32 products are exact in the fp32 accumulator, so the blend error is ~1e-7 relative,
// the same order as an fp32 FMA chain.  W' = [W_hi | W_lo] rows of a 128-vertex tile are RESIDENT
// IN TENSOR MEMORY as the MMA A operand (no shared-memory re-reads); A' = [A_hi | A_lo] per
// 8-body block (written by k2) is the B operand.  Both A' and the block's planar vposed rows
// stream through bulk-TMA / mbarrier rings with their own producer warps (A': 4 stages from L2;
// vposed: 4 stages from HBM, the 24 rows of a block issued by 24 lanes at once).  The kernel is
// bound by SM<->L2 traffic (operand re-reads + vposed + vertices, ~8 TB/s), so
// each CTA blends TWO adjacent vertex tiles per A' stage (2 x W' in TMEM, 4 accumulators): the A'
// re-read traffic halves and every vposed TMA row is 1 KB.  The grid is tile-pair-fastest: the
// CTAs resident at one time read and write adjacent row chunks of the same bodies (DRAM pages).
//
// Epilogue thread = vertex: reads its 12 blended entries per body from TMEM, the planar vposed
// coordinates (from the shared-memory ring, conflict-free), applies the 3x4 transform with FMAs and
// stores its own x, y, z (direct 12-byte-stride stores, see the epilogue).  k4 (weak-perspective projection,
// SURVEY.md A.8) rides in the epilogue of the CTAs that own vertex tile 0.
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kLbsTiles = 2;                             // vertex tiles per CTA sharing one B stage
constexpr int kLbsWarpTmaV = 0, kLbsWarpMma = 1, kLbsWarpTmaB = 2;   // warp 3 unused; epilogue = warps 4..19
constexpr int kLbsEpiWarp0 = 4;
constexpr int kLbsEpiWarps = 8 * kLbsTiles;              // per tile: two per TMEM lane quarter, 4 bodies each
constexpr int kLbsTcThreads = (kLbsEpiWarp0 + kLbsEpiWarps) * 32;    // 640
constexpr int kLbsBStages = 4;                           // A' images (L2-resident)
constexpr int kLbsVStages = 4;                           // vposed rows come from HBM: deep ring
constexpr int kLbsTcAcc = 2;
constexpr int kLbsN = kLbsBlock * 12;                    // 96
constexpr int kLbsTmemCols = 512;                        // 2 tiles x 2 x 96 accumulators + 2 x 48 of W'
constexpr int kLbsAccCols = kLbsTiles * kLbsTcAcc * kLbsN;   // 384
constexpr uint32_t kLbsBStage = kLbsK * kLbsN * 4;       // 18,432  tf32 hi|lo image of A, one block
constexpr uint32_t kLbsVRow = kLbsTiles * 128 * 4;       // one (body, plane) row of the tile pair: 1 KB
constexpr uint32_t kLbsVStage = kLbsBlock * 3 * kLbsVRow;  // 24,576
constexpr uint32_t kLbsVOff = kLbsBStages * kLbsBStage;
constexpr uint32_t kLbsOutOff = kLbsVOff + kLbsVStages * kLbsVStage;
constexpr uint32_t kLbsBarOff = kLbsOutOff;
constexpr uint32_t kLbsSmemBytes = kLbsBarOff + 256;     // ~160 KB, one CTA per SM
constexpr uint32_t kLbsIdesc = ptx::make_idesc(ptx::kFmtTF32, 128, kLbsN);

__global__ void __launch_bounds__(kLbsTcThreads, 1)
k_lbs_tc(const uint32_t* __restrict__ w_rows, const uint8_t* __restrict__ a_img,
         const float* __restrict__ vposed, long long n, int nblocks, int blocks_per_cta,
         int V, int VP, float* __restrict__ verts, const float* __restrict__ joints_in,
         const float* __restrict__ cam, float* __restrict__ kp2d) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sV = smem + kLbsVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kLbsBarOff);
  uint64_t* bar_w = bars;                            // W' rows resident in TMEM (8 warp arrivals)
  uint64_t* bar_bfull = bars + 1;                    // [B stages] A' image landed
  uint64_t* bar_bempty = bar_bfull + kLbsBStages;    // [B stages] MMAs reading it retired
  uint64_t* bar_vfull = bar_bempty + kLbsBStages;    // [V stages] vposed rows landed
  uint64_t* bar_vempty = bar_vfull + kLbsVStages;    // [V stages] epilogue warps done with them
  uint64_t* bar_tfull = bar_vempty + kLbsVStages;    // [acc] accumulators (both tiles) ready
  uint64_t* bar_tempty = bar_tfull + kLbsTcAcc;      // [acc] accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + kLbsTcAcc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vtiles = VP / 128;
  const int tile0 = blockIdx.x * kLbsTiles;          // tile-pair-fastest grid: co-resident CTAs touch
  const int ntile = min(kLbsTiles, vtiles - tile0);  // adjacent row chunks of the same bodies
  const int blk_begin = blockIdx.y * blocks_per_cta;
  const int blk_end = min(nblocks, blk_begin + blocks_per_cta);
  const int nblk = blk_end - blk_begin;

  if (warp == kLbsWarpTmaV && lane == 0) {
    ptx::mbar_init(bar_w, 4 * kLbsTiles);
    for (int s = 0; s < kLbsBStages; ++s) { ptx::mbar_init(bar_bfull + s, 1); ptx::mbar_init(bar_bempty + s, 1); }
    for (int s = 0; s < kLbsVStages; ++s) { ptx::mbar_init(bar_vfull + s, 1); ptx::mbar_init(bar_vempty + s, kLbsEpiWarps); }
    for (int a = 0; a < kLbsTcAcc; ++a) { ptx::mbar_init(bar_tfull + a, 1); ptx::mbar_init(bar_tempty + a, kLbsEpiWarps); }
    ptx::fence_barrier_init();
  }
  if (warp == kLbsWarpMma) ptx::tmem_alloc(tmem_slot, kLbsTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_w = tmem_base + kLbsAccCols;

  if (warp == kLbsWarpTmaV) {
    // ===== bulk-TMA producer 1: vposed rows of the tile pair (1 KB per body and plane), from HBM.
    // The block's 24 rows are issued by 24 LANES in one warp-wide instruction: a single thread
    // needs ~100 clk per cp.async.bulk (address + uniform-register shuffling), i.e. ~2400 clk per
    // block, which was the whole "loads-only" time of the round-1 ablation.
    {
      const uint32_t row_bytes = (uint32_t)ntile * 128 * 4;
      for (int i = 0; i < nblk; ++i) {
        const int s = i % kLbsVStages;
        const long long b0 = (long long)(blk_begin + i) * kLbsBlock;
        const int nb = (int)min((long long)kLbsBlock, n - b0);
        const int nrows = nb * 3;
        ptx::mbar_wait(bar_vempty + s, ((i / kLbsVStages) & 1) ^ 1);
        if (lane == 0) ptx::mbar_arrive_expect_tx(bar_vfull + s, (uint32_t)nrows * row_bytes);
        __syncwarp();
        const float* src = vposed + (size_t)b0 * 3 * VP + (size_t)tile0 * 128;
        uint8_t* dst = sV + (size_t)s * kLbsVStage;
        if (lane < nrows)
          ptx::bulk_g2s(dst + (size_t)lane * kLbsVRow, src + (size_t)lane * VP, row_bytes, bar_vfull + s);
        __syncwarp();
      }
    }
  } else if (warp == kLbsWarpTmaB) {
    // ===== bulk-TMA producer 2: tf32 hi|lo image of the block's joint transforms (L2-resident) =====
    if (lane == 0) {
      for (int i = 0; i < nblk; ++i) {
        const int s = i % kLbsBStages;
        ptx::mbar_wait(bar_bempty + s, ((i / kLbsBStages) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(bar_bfull + s, kLbsBStage);
        ptx::bulk_g2s(sB + (size_t)s * kLbsBStage, a_img + (size_t)(blk_begin + i) * kLbsBStage, kLbsBStage,
                      bar_bfull + s);
      }
    }
  } else if (warp == kLbsWarpMma) {
    // ===== MMA issuer: one B stage feeds the blend MMAs of BOTH vertex tiles.  The whole warp runs
    // the loop convergently and ONE elected lane issues (descriptors stay in uniform registers; from
    // a divergent `lane == 0` branch every tcgen05.mma costs ~50 clk of issue, 18 per block). =====
    if (nblk > 0) {
      constexpr uint32_t kLboB = kLbsN * 16, kSbo = 128;
      constexpr uint32_t kHalfB = 6 * kLboB;     // byte offset of the A_lo K half in the image
      ptx::mbar_wait(bar_w, 0);
      for (int i = 0; i < nblk; ++i) {
        const int s = i % kLbsBStages, a = i % kLbsTcAcc;
        ptx::mbar_wait(bar_tempty + a, ((i / kLbsTcAcc) & 1) ^ 1);
        ptx::mbar_wait(bar_bfull + s, (i / kLbsBStages) & 1);
        ptx::tc_fence_after();
        const uint32_t b_addr = ptx::smem_u32(sB + (size_t)s * kLbsBStage);
        if (ptx::elect_one()) {
#pragma unroll
          for (int t = 0; t < kLbsTiles; ++t) {
            if (t < ntile) {
              const uint32_t d_tmem = tmem_base + (t * kLbsTcAcc + a) * kLbsN;
#pragma unroll
              for (int g = 0; g < 3; ++g) {  // (W_hi,A_hi), (W_hi,A_lo), (W_lo,A_hi)
                const uint32_t wp = tmem_w + t * kLbsK + (g == 2 ? 24 : 0);
                const uint32_t bp = b_addr + (g == 1 ? kHalfB : 0);
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {  // 24 joints = 3 tf32 k-steps of 8
                  const uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * kLboB, kLboB, kSbo);
                  ptx::mma_tf32_ts(d_tmem, wp + ks * 8, bd, kLbsIdesc, (uint32_t)((g | ks) != 0));
                }
              }
            }
          }
          ptx::tc_commit(bar_bempty + s);
          ptx::tc_commit(bar_tfull + a);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kLbsEpiWarp0) {
    // ===== epilogue (16 warps: tile t, TMEM lane quarter q, bodies 4h .. 4h+3 of each block) =====
    const int ew = warp - kLbsEpiWarp0;
    const int q = warp & 3;
    const int t = (ew >> 2) & 1;
    const int h = ew >> 3;
    const bool live = t < ntile;                       // second tile may not exist (odd tile count)
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int v_local = t * 128 + q * 32 + lane;       // column inside the staged tile-pair row
    const int warp_v0 = (tile0 + t) * 128 + q * 32;
    const int nf = max(0, min(32, V - warp_v0)) * 3;   // floats this warp may store per body
    if (nblk > 0 && h == 0) {
      // W' rows of this vertex tile -> TMEM (A operand of every blend MMA of this tile)
      if (live) {
        const uint4* src = reinterpret_cast<const uint4*>(w_rows + ((size_t)(tile0 + t) * 128 + q * 32 + lane) * kLbsK);
#pragma unroll
        for (int c = 0; c < kLbsK / 16; ++c) {
          uint32_t w[16];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const uint4 x = __ldg(src + c * 4 + v);
            w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
          }
          ptx::tmem_st16(tmem_w + lane_addr + t * kLbsK + c * 16, w);
        }
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_w);
      __syncwarp();
    }
    for (int i = 0; i < nblk; ++i) {
      const int s = i % kLbsVStages, a = i % kLbsTcAcc;
      const long long b0 = (long long)(blk_begin + i) * kLbsBlock;
      const int nb = (int)min((long long)kLbsBlock, n - b0);
      const float* sv = reinterpret_cast<const float*>(sV + (size_t)s * kLbsVStage) + v_local;
      const size_t body_stride = (size_t)V * 3;
      ptx::mbar_wait(bar_vfull + s, (i / kLbsVStages) & 1);
      ptx::mbar_wait(bar_tfull + a, (i / kLbsTcAcc) & 1);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + lane_addr + (t * kLbsTcAcc + a) * kLbsN + h * 48;
      {   // this warp's 4 bodies = 48 TMEM columns
        uint32_t r0[16], r1[16], r2[16];
        ptx::tmem_ld16(t_addr, r0);
        ptx::tmem_ld16(t_addr + 16, r1);
        ptx::tmem_ld16(t_addr + 32, r2);
        float px[4], py[4], pz[4];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const float* p = sv + (h * 4 + bb) * 3 * (kLbsTiles * 128);
          px[bb] = p[0]; py[bb] = p[kLbsTiles * 128]; pz[bb] = p[2 * kLbsTiles * 128];
        }
        ptx::tmem_ld_wait();
        // both the accumulator columns and the vposed stage are now in registers: release them
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(bar_tempty + a); ptx::mbar_arrive(bar_vempty + s); }
        __syncwarp();
        float T[48];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          T[e] = __uint_as_float(r0[e]); T[16 + e] = __uint_as_float(r1[e]); T[32 + e] = __uint_as_float(r2[e]);
        }
        // all FMAs, then DIRECT stores: the lane writes its own x, y, z (12-byte stride across the warp; the three
        // store instructions of a body together cover the warp's 384 contiguous bytes, so every 32-byte sector is
        // completed in L2 at once).  Round 1 interleaved xyz through a shared-memory tile for 128-byte stores;
        // A/B-timed in the fused kernel the direct form is 5 % faster (no STS/LDS/warp barriers in the chain).
        if (live && lane < nf / 3) {
          float* d = verts + ((size_t)(b0 + 4 * h) * V + warp_v0 + lane) * 3;
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            if (h * 4 + bb < nb) {
              const float* tt = T + bb * 12;
              const float x = px[bb], y = py[bb], z = pz[bb];
              d[0] = fmaf(tt[2], z, fmaf(tt[1], y, fmaf(tt[0], x, tt[3])));
              d[1] = fmaf(tt[6], z, fmaf(tt[5], y, fmaf(tt[4], x, tt[7])));
              d[2] = fmaf(tt[10], z, fmaf(tt[9], y, fmaf(tt[8], x, tt[11])));
            }
            d += body_stride;
          }
        }
      }
    }
    // k4: weak-perspective projection of this CTA's bodies (CTAs of the first tile pair only)
    if (tile0 == 0 && kp2d != nullptr) {
      const long long bb0 = (long long)blk_begin * kLbsBlock;
      const long long bb1 = min(n, (long long)blk_end * kLbsBlock);
      for (long long i = bb0 * (kJ * 2) + ((int)threadIdx.x - kLbsEpiWarp0 * 32); i < bb1 * (kJ * 2);
           i += kLbsEpiWarps * 32) {
        const long long b = i / (kJ * 2);
        const int r = int(i - b * (kJ * 2)), j = r >> 1, c = r & 1;
        const float sc = __ldg(cam + b * 3), tt = __ldg(cam + b * 3 + 1 + c);
        kp2d[i] = __fmul_rn(sc, __fadd_rn(__ldg(joints_in + (b * kJ + j) * 3 + c), tt));
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kLbsWarpMma) ptx::tmem_dealloc(tmem_base, kLbsTmemCols);
}

// fp32 A [n,24,12] -> tf32 hi|lo operand image (stand-alone k3 entry point only).
__global__ void __launch_bounds__(256)
k_pack_a(const float* __restrict__ A, long long n, uint32_t* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * (kJ * 12)) return;
  const long long b = idx / (kJ * 12);
  const int r = int(idx - b * (kJ * 12)), jj = r / 12, e = r % 12;
  const float v = A[idx];
  const uint32_t hi = f32_to_tf32_rn(v);
  const uint32_t lo = f32_to_tf32_rn(__fsub_rn(v, __uint_as_float(hi)));
  const long long blk = b / kLbsBlock;
  const int row = int(b % kLbsBlock) * 12 + e;
  uint32_t* im = img + blk * (long long)(kLbsK * kLbsN);
  im[(size_t)(jj >> 2) * (kLbsN * 4) + row * 4 + (jj & 3)] = hi;
  im[(size_t)((24 + jj) >> 2) * (kLbsN * 4) + row * 4 + ((24 + jj) & 3)] = lo;
}

// Grid shape: x = vertex tile pair (fastest), y = body-block group.  The group count is chosen so
// the CTA count lands just under a whole number of waves (one CTA per SM).
inline int lbs_tc_blocks_per_cta(int cta_x, int nblocks, int num_sms) {
  const double slots = (double)num_sms;
  int best_bpc = nblocks;
  double best_eff = -1.0;
  for (int bpc = nblocks; bpc >= 1; --bpc) {
    if (bpc < 16 && bpc < nblocks) break;                 // keep the W' load / prologue amortised
    const int groups = (nblocks + bpc - 1) / bpc;
    const double waves = cta_x * (double)groups / slots;
    const double eff = waves / std::ceil(waves);
    if (eff > best_eff + 0.02) { best_eff = eff; best_bpc = bpc; }
  }
  return best_bpc;
}

inline cudaError_t launch_lbs_tc(const DeviceModel& m, int num_sms, const float* vposed,
                                 const uint32_t* a_img, long long n, float* verts,
                                 const float* joints_in, const float* cam, float* kp2d,
                                 cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int vtiles = m.VP / 128;
  const int cta_x = (vtiles + kLbsTiles - 1) / kLbsTiles;
  const int nblocks = (int)((n + kLbsBlock - 1) / kLbsBlock);
  const int bpc = lbs_tc_blocks_per_cta(cta_x, nblocks, num_sms);
  const dim3 grid((unsigned)cta_x, (unsigned)((nblocks + bpc - 1) / bpc));
  k_lbs_tc<<<grid, kLbsTcThreads, kLbsSmemBytes, s>>>(
      m.w_tf32, reinterpret_cast<const uint8_t*>(a_img), vposed, n, nblocks, bpc, m.V, m.VP, verts,
      joints_in, cam, kp2d);
  return cudaGetLastError();
}

}  // namespace smplb200
