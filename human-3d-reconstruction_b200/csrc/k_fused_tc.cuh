// k1 + k3 FUSED: blendshapes and linear blend skinning of a
// (128-vertex tile, 64-body block) unit in one CTA, both contractions on tcgen05, accumulators in
// tensor memory -- the v_posed intermediate (82,680 B per body written by k1 and read back by k3,
// 2/3 of the unfused step's DRAM traffic) never exists.
//
//   D_p[v, b]     = sum_k basis_p[v, k] * coef[b, k]            p = x, y, z planes; K = 224
//   T[v, (b, e)]  = sum_j W[v, j] * A[b, j, e]                  e = 12 entries of the 3x4 transform
//   verts[b, v, r] = T[v,(b,r,0..2)] . (D_x, D_y, D_z)[v, b] + T[v,(b,r,3)]
//
// Both accumulators have the tile's 128 VERTICES on the TMEM lanes, so the epilogue thread that owns
// vertex v reads its blended transform and its posed rest position for body b from its own lane.
//
// Operand placement (what fits one SM; see DESIGN.md §3d for the arithmetic)
//   * the tile's blendshape basis is the A operand of the D MMAs and stays RESIDENT IN SHARED MEMORY
//     for all of the CTA's units: 3 planes x [224 K][128 rows] fp16 = 168 KB (+ 12 KB: the low halves
//     of the 16 shape/template rows).  It cannot live in TMEM (3 x 112 columns) next to the
//     accumulators, so the D MMAs are SS-form with N = 64: their cost is the shared-memory read of A
//     (4 KB per MMA, ~52 clk measured; scripts/mma_issue_microbench.cu), not tensor math.
//   * coef blocks (B operand, 30 KB per unit) stream through a 3-stage bulk-TMA ring in K chunks; every
//     chunk feeds the MMAs of all three planes before it is released.
//   * skinning weights W' = [W_hi | W_lo] of the tile live in TMEM (32 columns) as the A operand of the
//     blend MMAs (TS form, N = 48 = 4 bodies x 12 entries, ~32 clk each); the joint transforms A' stream
//     through two 2-stage rings (one per blend issuer), one 5.3 KB image per 4-body sub-block.
//   * TMEM: D double-buffered (2 x 3 x 64 columns), T double-buffered (2 x 48), W' (32) = 512 columns.
//
// Precision (SMPLB200_PREC_F16): fp16 operands (11 significant bits), fp32 accumulation.
//   * pose rows (207 of the 224 K): ONE MMA, operands rounded to fp16: measured max vertex error 1.9e-5 m
//     against float64 (stated bound 5e-5 m; TF32 operands give 1.9e-4, bf16 1.4e-3).
//   * shape rows and the template (the O(1 m) and O(0.06 m) terms): exact 3-term split
//     (hi*hi + lo*hi + hi*lo; the template as three fp16 pieces with coefficient 1).
//   * the skinning blend: 3-term fp16 split of BOTH operands (W_hi*A_hi + W_hi*A_lo + W_lo*A_hi, products
//     exact in the fp32 accumulator): ~2e-7, the same fp32-class fidelity as k_lbs_tc's 3xTF32.
//
// The epilogue writes vertices straight from registers (lane = vertex, 12-byte stride), see the store comment.
//
// Schedule: persistent CTAs (one per SM) own equal contiguous ranges of the tile-major unit list.
// Warp roles (672 threads): warp 0 = bulk-TMA producer (basis tile, coef chunks), warps 1 and 20 = blend-MMA
// issuers of T buffer 0 / 1, warp 2 = bulk-TMA producer of the A' images, warp 3 = D-MMA issuer (all issuers
// warp-uniform, one elected lane), warps 4..19 = epilogue: TMEM lane quarter q = warp % 4, slot e takes the
// sub-blocks s = e, e + 2, ... (T buffer e), body half h takes two of the sub-block's four bodies.  The D issuer
// runs up to one unit ahead (D is double-buffered), so the tensor pipe interleaves the next unit's blendshape
// MMAs with this unit's blend MMAs.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kFzBodies = 64;                       // bodies per unit (N of the D MMAs)
constexpr long long kFzMaxBodiesPerLaunch = 8192;   // operand images of one launch stay L2-resident (15 MB)
constexpr int kFzSub = 4;                           // bodies per blend sub-block
constexpr int kFzSubs = kFzBodies / kFzSub;         // 16
constexpr int kFzNT = kFzSub * 12;                  // 48: N of the blend MMAs
constexpr int kFzShapeK = 16;                       // K rows 0..15: betas | template pieces | 0
constexpr int kFzThreads = 672;                     // 21 warps, see "Warp roles"
constexpr int kFzEpiWarp0 = 4, kFzEpiWarps = 16, kFzWarpT1 = 20;  // warp 20: blend issuer of slot 1
constexpr uint32_t kFzPlaneHi = kCoefK * 128 * 2;               // 57,344
constexpr uint32_t kFzPlaneLo = kFzShapeK * 128 * 2;            // 4,096
constexpr uint32_t kFzBasisBytes = 3 * (kFzPlaneHi + kFzPlaneLo);   // 184,320 per vertex tile
constexpr uint32_t kFzCoefLo = kFzShapeK * kFzBodies * 2;       // 2,048
constexpr uint32_t kFzCoefBlock = kFzCoefLo + kCoefK * kFzBodies * 2;   // 30,720 per 64-body block
constexpr int kFzDLead = 6;                                     // D groups run this many sub-blocks ahead of the pacing
constexpr int kFzChunks = 7;                                    // K chunks per unit: 2 k-steps each
constexpr uint32_t kFzCoefStage = kFzCoefLo + 2 * 2048;         // 6,144 (chunk 0 carries the lo rows too)
constexpr int kFzCoefStages = 3;
constexpr uint32_t kFzAImage = 7 * kFzNT * 16;                  // 5,376: [A_hi 3 chunks | A_lo 3 chunks | 0]
constexpr int kFzAStages = 2;                                   // PER SLOT: each blend issuer has its own ring
constexpr uint32_t kFzOffCoef = kFzBasisBytes;
constexpr uint32_t kFzOffA = kFzOffCoef + kFzCoefStages * kFzCoefStage;
constexpr uint32_t kFzOffBar = kFzOffA + 2 * kFzAStages * kFzAImage;
constexpr uint32_t kFzSmemBytes = kFzOffBar + 256;              // 224,512 <= 232,448
constexpr uint32_t kFzTmemD = 0, kFzTmemT = 2 * 3 * kFzBodies, kFzTmemW = kFzTmemT + 2 * kFzNT;   // 0 | 384 | 480
constexpr uint32_t kFzIdescD = ptx::make_idesc(ptx::kFmtF16, 128, kFzBodies);
constexpr uint32_t kFzIdescT = ptx::make_idesc(ptx::kFmtF16, 128, kFzNT);
static_assert(kFzSmemBytes <= 232448, "shared memory budget");
static_assert(kFzTmemW + 32 == 512, "TMEM budget");

__device__ __forceinline__ uint16_t f32_to_f16_rn(float x) { return __half_as_ushort(__float2half_rn(x)); }
__device__ __forceinline__ float f16_to_f32(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// position of old-order coefficient k (betas | pose_feature | 1 1 1 | 0) in the fused kernel's K order
// (betas | 1 1 1 | 0.. up to 16 | pose_feature | 0): returns the OLD index feeding new row `nk`, or -1 (zero)
__device__ __forceinline__ int fz_old_index(int nk, int NB) {
  if (nk < NB) return nk;
  if (nk < NB + 3) return NB + kP + (nk - NB);
  if (nk < kFzShapeK) return -1;
  if (nk < kFzShapeK + kP) return NB + (nk - kFzShapeK);
  return -1;
}

#ifdef SMPLB200_FZ_TIMING
__device__ long long g_fz_time[148 * 32];
// accumulate in registers (acc0..acc2 of the role), written once when the role's loop ends
#define FZ_T0() const long long _t0 = clock64()
#define FZ_ACC(var) do { var += clock64() - _t0; } while (0)
#define FZ_DECL long long fz_a0 = 0, fz_a1 = 0, fz_a2 = 0
#define FZ_OUT(slot, var) do { if (lane == 0) g_fz_time[blockIdx.x * 32 + (slot)] += var; } while (0)
#else
#define FZ_T0() do { } while (0)
#define FZ_ACC(var) do { } while (0)
#define FZ_DECL do { } while (0)
#define FZ_OUT(slot, var) do { } while (0)
#endif

__global__ void __launch_bounds__(kFzThreads, 1)
k_fused_tc(const uint8_t* __restrict__ basis_tiles, const uint32_t* __restrict__ w_rows,
           const uint8_t* __restrict__ coef_img, const uint8_t* __restrict__ a_img,
           long long n, int nblk, long long total_units, int V, float* __restrict__ verts) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sBasis = smem;
  uint8_t* sCoef = smem + kFzOffCoef;
  uint8_t* sA = smem + kFzOffA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFzOffBar);
  uint64_t* bar_bfull = bars;                          // basis tile landed (tx)
  uint64_t* bar_bfree = bars + 1;                      // every MMA reading the old basis tile retired
  uint64_t* bar_w = bars + 2;                          // W' rows of the tile are in TMEM (4 warp arrivals)
  uint64_t* bar_cfull = bars + 3;                      // [3] coef chunk landed
  uint64_t* bar_cempty = bar_cfull + kFzCoefStages;    // [3] MMAs reading it retired
  uint64_t* bar_afull = bar_cempty + kFzCoefStages;    // [slot][2] A' image landed
  uint64_t* bar_aempty = bar_afull + 2 * kFzAStages;   // [slot][2]
  uint64_t* bar_dfull = bar_aempty + 2 * kFzAStages;   // [2] D accumulators of a unit complete
  uint64_t* bar_dempty = bar_dfull + 2;                // [2] drained by the 16 epilogue warps
  uint64_t* bar_tfull = bar_dempty + 2;                // [2] blend accumulator complete
  uint64_t* bar_tempty = bar_tfull + 2;                // [2] read by the 8 warps of its slot
  uint64_t* bar_wfree = bar_tempty + 2;                // slot-1 epilogue warps are done with the old tile's W' (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_wfree + 1);
  volatile uint32_t* pace = tmem_slot + 1;             // blend groups issued so far (paces the D issuer, see below)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SMPLB200_FZ_TIMING
  const long long t_kernel0 = clock64();
#endif
  const long long u0 = total_units * blockIdx.x / gridDim.x;
  const long long u1 = total_units * (blockIdx.x + 1) / gridDim.x;
  const int nunits = (int)(u1 - u0);

  if (warp == 0 && lane == 0) {
    *pace = 0u;
    ptx::mbar_init(bar_bfull, 1); ptx::mbar_init(bar_bfree, 1); ptx::mbar_init(bar_w, 4); ptx::mbar_init(bar_wfree, kFzEpiWarps / 2);
    for (int s = 0; s < kFzCoefStages; ++s) { ptx::mbar_init(bar_cfull + s, 1); ptx::mbar_init(bar_cempty + s, 1); }
    for (int s = 0; s < 2 * kFzAStages; ++s) { ptx::mbar_init(bar_afull + s, 1); ptx::mbar_init(bar_aempty + s, 1); }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(bar_dfull + a, 1); ptx::mbar_init(bar_dempty + a, kFzEpiWarps);
      ptx::mbar_init(bar_tfull + a, 1); ptx::mbar_init(bar_tempty + a, kFzEpiWarps / 2);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer 1: the tile's basis image (on tile switches) and the coef chunks of every unit =====
    if (lane == 0) {
      long long cur_tile = -1;
      int ntile_sw = 0, it = 0;
      for (int i = 0; i < nunits; ++i) {
        const long long u = u0 + i, tile = u / nblk;
        const int blk = (int)(u % nblk);
        if (tile != cur_tile) {
          if (ntile_sw > 0) ptx::mbar_wait_nohint(bar_bfree, (ntile_sw - 1) & 1);   // old tile's D MMAs retired
          ptx::mbar_arrive_expect_tx(bar_bfull, kFzBasisBytes);
          ptx::bulk_g2s_split(sBasis, basis_tiles + (size_t)tile * kFzBasisBytes, kFzBasisBytes, bar_bfull);
          cur_tile = tile;
          ++ntile_sw;
        }
        const uint8_t* src = coef_img + (size_t)blk * kFzCoefBlock;
        for (int c = 0; c < kFzChunks; ++c, ++it) {
          SMPLB200_PROGRESS((i << 8) | c);
          const int s = it % kFzCoefStages;
          ptx::mbar_wait_nohint(bar_cempty + s, ((it / kFzCoefStages) & 1) ^ 1);
          const uint32_t bytes = c == 0 ? kFzCoefStage : 4096u;
          const size_t off = c == 0 ? 0 : (size_t)kFzCoefLo + 4096u * c;
          ptx::mbar_arrive_expect_tx(bar_cfull + s, bytes);
          ptx::bulk_g2s(sCoef + (size_t)s * kFzCoefStage, src + off, bytes, bar_cfull + s);
        }
      }
    }
  } else if (warp == 2) {
    // ===== producer 2: fp16 hi|lo images of the joint transforms, one per 4-body sub-block.  The images
    // alternate between the two blend issuers, and EACH ISSUER HAS ITS OWN 2-STAGE RING: a ring shared by
    // two consumers breaks the parity protocol (a waiter must have seen the barrier's previous phase, which
    // belonged to the other consumer). =====
    FZ_DECL;
    if (lane == 0) {
      int it = 0;
      for (int i = 0; i < nunits; ++i) {
        const int blk = (int)((u0 + i) % nblk);
        const uint8_t* src = a_img + (size_t)blk * kFzSubs * kFzAImage;
        for (int sb = 0; sb < kFzSubs; ++sb, ++it) {
          SMPLB200_PROGRESS((i << 8) | sb);
          const int k = it >> 1;                           // this slot's image counter
          const int s = (it & 1) * kFzAStages + (k & 1);   // [slot][stage]
          { FZ_T0(); ptx::mbar_wait_nohint(bar_aempty + s, ((k >> 1) & 1) ^ 1); FZ_ACC(fz_a0); }
          ptx::mbar_arrive_expect_tx(bar_afull + s, kFzAImage);
          ptx::bulk_g2s(sA + (size_t)s * kFzAImage, src + (size_t)sb * kFzAImage, kFzAImage, bar_afull + s);
        }
      }
      FZ_OUT(0, fz_a0);
    }
  } else if (warp == 3) {
    // ===== D-MMA issuer (blendshapes): whole warp convergent, one elected lane issues.  Its own warp, so
    // the scalar bookkeeping of the D chunks never sits in the blend issuer's loop (a single issuer doing both
    // was the kernel's critical path: ~1300 clk of index math, waits and descriptor building per sub-block).
    // All counters are incremental: no division / modulo in the loop. =====
    FZ_DECL;
    const uint32_t basis_addr = ptx::smem_u32(sBasis);
    const uint32_t coef_addr = ptx::smem_u32(sCoef);
    constexpr uint32_t kLboA = 128 * 16, kLboB = kFzBodies * 16, kSbo = 128;
    // descriptor of address 0 for each operand family; the 14-bit address field (addr >> 4) is added per MMA
    const uint64_t descA0 = ptx::make_smem_desc(0, kLboA, kSbo), descB0 = ptx::make_smem_desc(0, kLboB, kSbo);
    auto dA = [&](uint32_t addr) { return descA0 | (uint64_t)((addr >> 4) & 0x3fffu); };
    auto dB = [&](uint32_t addr) { return descB0 | (uint64_t)((addr >> 4) & 0x3fffu); };
    int blk = (int)(u0 % nblk);                 // body block of the current unit (one division, outside the loop)
    bool new_tile = true;                       // the first unit always loads a basis tile
    uint32_t b_phase = 0;                       // parity of the next bar_bfull wait
    int cs = 0; uint32_t c_phase = 0;           // coef ring stage / parity
    for (int i = 0; i < nunits; ++i) {
      const int a = i & 1;
      const bool last_of_tile = (i + 1 == nunits) || (blk + 1 == nblk);
      if (new_tile) { ptx::mbar_wait_nohint(bar_bfull, b_phase); b_phase ^= 1; }
      { FZ_T0(); ptx::mbar_wait_nohint(bar_dempty + a, ((i >> 1) & 1) ^ 1); FZ_ACC(fz_a1); }     // epilogue drained this buffer (unit i-2)
      const uint32_t d_tmem0 = tmem_base + kFzTmemD + a * 3 * kFzBodies;
      // The unit's 54 MMAs go out chunk by chunk (12 MMAs for chunk 0, 6 for chunks 1..6).  The tensor pipe
      // executes in issue order, so a whole unit of D MMAs issued at once (2.8 kclk) would stall every blend
      // MMA -- and with it the epilogue -- behind it.  A chunk is therefore held back until the blend issuers
      // have reached the matching sub-block of unit i-1, less a lead (`pace`, a monotonic counter in shared
      // memory): the two MMA streams interleave, and the unit's accumulators are complete a few sub-blocks
      // BEFORE the epilogue asks for them.  The issue code is straight-line per chunk: this warp's own
      // instruction latency (branches, descriptor arithmetic) was what made D late.
      const uint32_t pace_base = (uint32_t)(i - 1) * kFzSubs;
#pragma unroll 1
      for (int c = 0; c < kFzChunks; ++c) {
        { FZ_T0(); ptx::mbar_wait_nohint(bar_cfull + cs, c_phase); FZ_ACC(fz_a0); }
        if (i > 0) {      // pacing, once per chunk: chunk c goes out when sub-block 2c - kFzDLead of unit i-1 has
          const int g0 = c == 0 ? 0 : 2 * c + 2;
          const uint32_t need = pace_base + (uint32_t)(g0 > kFzDLead ? g0 - kFzDLead : 0) + 1u;
          FZ_T0(); while ((int32_t)(*pace - need) < 0) __nanosleep(64); FZ_ACC(fz_a2);
        }
        ptx::tc_fence_after();
        const uint32_t c_addr = coef_addr + cs * kFzCoefStage;
        if (ptx::elect_one()) {
          // straight-line issue: every descriptor is (a base computed once per chunk) | (a compile-time offset)
          const uint64_t a_hi = dA(basis_addr), b_0 = dB(c_addr);
          constexpr uint64_t kPlaneA = kFzPlaneHi >> 4, kStepA = (2 * kLboA) >> 4, kLoA = (3 * kFzPlaneHi) >> 4,
                             kPlaneLoA = kFzPlaneLo >> 4, kLoB = kFzCoefLo >> 4, kStepB = 2048 >> 4;
          if (c == 0) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {     // shape rows: hi*hi, hi(basis)*lo(coef), lo(basis)*hi(coef); first pose k-step
              const uint32_t d_tmem = d_tmem0 + p * kFzBodies;
              ptx::mma_bf16(d_tmem, a_hi + p * kPlaneA, b_0 + kLoB, kFzIdescD, 0u);
              ptx::mma_bf16(d_tmem, a_hi + p * kPlaneA, b_0, kFzIdescD, 1u);
              ptx::mma_bf16(d_tmem, a_hi + kLoA + p * kPlaneLoA, b_0 + kLoB, kFzIdescD, 1u);
              ptx::mma_bf16(d_tmem, a_hi + p * kPlaneA + kStepA, b_0 + kLoB + kStepB, kFzIdescD, 1u);
            }
          } else {
            const uint64_t a_c = a_hi + (uint64_t)(2 * c) * kStepA;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
              for (int p = 0; p < 3; ++p)
                ptx::mma_bf16(d_tmem0 + p * kFzBodies, a_c + ks * kStepA + p * kPlaneA, b_0 + ks * kStepB, kFzIdescD, 1u);
          }
          ptx::tc_commit(bar_cempty + cs);
          if (c == kFzChunks - 1) {
            ptx::tc_commit(bar_dfull + a);
            if (last_of_tile) ptx::tc_commit(bar_bfree);     // the producer may overwrite the basis tile
          }
        }
        __syncwarp();
        if (++cs == kFzCoefStages) { cs = 0; c_phase ^= 1; }
      }
      new_tile = (blk + 1 == nblk);
      blk = new_tile ? 0 : blk + 1;
    }
    FZ_OUT(16, fz_a0); FZ_OUT(17, fz_a1); FZ_OUT(18, fz_a2);
  } else if (warp == 1 || warp == kFzWarpT1) {
    // ===== blend-MMA issuers (skinning): one 6-MMA group per 4-body sub-block.  One issuer warp PER T buffer
    // (slot e: sub-blocks e, e+2, ...): a slot's chain "blend MMAs -> epilogue reads T -> buffer free -> next
    // blend" then never waits behind the other slot's bookkeeping in a shared instruction stream. =====
    FZ_DECL;
    const int e = warp == 1 ? 0 : 1;
    const uint32_t tmem_w = tmem_base + kFzTmemW;
    const uint32_t a_addr0 = ptx::smem_u32(sA);
    const uint32_t t_tmem = tmem_base + kFzTmemT + e * kFzNT;
    constexpr uint32_t kLbo = kFzNT * 16, kSbo = 128;
    const uint64_t desc0 = ptx::make_smem_desc(0, kLbo, kSbo);
    auto dS = [&](uint32_t addr) { return desc0 | (uint64_t)((addr >> 4) & 0x3fffu); };
    int blk = (int)(u0 % nblk);
    bool new_tile = true;
    uint32_t w_phase = 0;
    // This loop is ON the accumulator hand-off chain (epilogue frees T -> this warp issues -> MMAs -> epilogue),
    // so it is written to be short: the ring has two stages and a slot takes eight sub-blocks per unit, so the
    // stage is (iteration & 1) -- the loop is unrolled by two and all eight shared-memory descriptors (2 stages x
    // 4 K steps), the barrier addresses and the pace word's address are loop-invariant registers.
    uint64_t dsc[kFzAStages][4];
#pragma unroll
    for (int st = 0; st < kFzAStages; ++st) {
      const uint32_t a_addr = a_addr0 + (e * kFzAStages + st) * kFzAImage;
      dsc[st][0] = dS(a_addr);              // A_hi joints 0..15
      dsc[st][1] = dS(a_addr + 2 * kLbo);   // A_hi 16..23 | (A_lo 0..7 x 0)
      dsc[st][2] = dS(a_addr + 3 * kLbo);   // A_lo joints 0..15
      dsc[st][3] = dS(a_addr + 5 * kLbo);   // A_lo 16..23 | zeros
    }
    const uint32_t pace_addr = ptx::smem_u32(const_cast<uint32_t*>(pace));
    uint32_t a_phase = 0;                       // parity of this slot's A' ring (flips every two sub-blocks)
    uint32_t te_phase = 1;                      // parity of the bar_tempty wait: the buffer starts free
    for (int i = 0; i < nunits; ++i) {
      if (new_tile) { ptx::mbar_wait_nohint(bar_w, w_phase); w_phase ^= 1; }     // the tile's W' rows are in TMEM
#pragma unroll 1
      for (int sb = e; sb < kFzSubs; sb += 4) {
#pragma unroll
        for (int st = 0; st < kFzAStages; ++st) {
          SMPLB200_PROGRESS((i << 8) | (sb + 2 * st));
          // operand first (it has usually landed long ago), then the accumulator: when the epilogue frees the
          // T buffer nothing but the issue itself stands between that arrival and the next blend
          { FZ_T0(); ptx::mbar_wait_nohint(bar_afull + e * kFzAStages + st, a_phase); FZ_ACC(fz_a1); }
          { FZ_T0(); ptx::mbar_wait_nohint(bar_tempty + e, te_phase); FZ_ACC(fz_a0); }
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            ptx::mma_bf16_ts(t_tmem, tmem_w, dsc[st][0], kFzIdescT, 0u);        // W_hi * A_hi
            ptx::mma_bf16_ts(t_tmem, tmem_w + 8, dsc[st][1], kFzIdescT, 1u);
            ptx::mma_bf16_ts(t_tmem, tmem_w, dsc[st][2], kFzIdescT, 1u);        // W_hi * A_lo
            ptx::mma_bf16_ts(t_tmem, tmem_w + 8, dsc[st][3], kFzIdescT, 1u);
            ptx::mma_bf16_ts(t_tmem, tmem_w + 16, dsc[st][0], kFzIdescT, 1u);   // W_lo * A_hi
            ptx::mma_bf16_ts(t_tmem, tmem_w + 24, dsc[st][1], kFzIdescT, 1u);
            ptx::tc_commit(bar_aempty + e * kFzAStages + st);
            ptx::tc_commit(bar_tfull + e);
            if (e == 1)       // lets the D issuer release its next chunks
              asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(pace_addr),
                           "r"((uint32_t)i * kFzSubs + (uint32_t)(sb + 2 * st) + 1u) : "memory");
          }
          __syncwarp();
          te_phase ^= 1;
        }
        a_phase ^= 1;
      }
      new_tile = (blk + 1 == nblk);
      blk = new_tile ? 0 : blk + 1;
    }
    FZ_OUT(2 + e * 4, fz_a0); FZ_OUT(3 + e * 4, fz_a1);
  } else if (warp >= kFzEpiWarp0 && warp < kFzEpiWarp0 + kFzEpiWarps) {
    // ===== epilogue: slot e (T buffer), lane quarter q, body half h (bodies 2h, 2h+1 of each sub-block).
    // Thread = vertex: it reads its blended 3x4 transform (T) and its posed rest position (D) for two bodies from
    // its own TMEM lane, applies the transform and stores the vertex. =====
    FZ_DECL;
    const int ew = warp - kFzEpiWarp0;
    const int q = warp & 3, e = (ew >> 2) & 1, h = ew >> 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long long cur_tile = -1;
    int t_cnt = 0;                    // sub-blocks this warp has consumed (phase of its T buffer)
    uint32_t wf_phase = 0;            // parity of the next bar_wfree wait (slot 0 only)
    const size_t body_stride = (size_t)V * 3;
    for (int i = 0; i < nunits; ++i) {
      const long long u = u0 + i, tile = u / nblk;
      const int blk = (int)(u % nblk), a = i & 1;
      if (tile != cur_tile) {
        // New tile: its W' rows replace the old ones in TMEM.  Every blend MMA of the old tile has retired
        // once BOTH slots have consumed the last blend results of the previous unit: slot 0's warps have
        // (program order); slot 1's warps say so on bar_wfree.
        if (e == 1 || h == 1) {
          if (e == 1 && i > 0) {  // this slot has consumed its last blend result of the old tile (program order)
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar_wfree);
            __syncwarp();
          }
        } else {
          if (i > 0) { ptx::mbar_wait_nohint(bar_wfree, wf_phase); wf_phase ^= 1; }
          ptx::tc_fence_after();
          const uint4* src = reinterpret_cast<const uint4*>(w_rows + ((size_t)tile * 128 + q * 32 + lane) * 32);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t w[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint4 x = __ldg(src + c * 4 + v);
              w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
            }
            ptx::tmem_st16(tmem_base + kFzTmemW + lane_addr + c * 16, w);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_w);
          __syncwarp();
        }
        cur_tile = tile;
      }
      const int warp_v0 = (int)tile * 128 + q * 32;
      const int nv = max(0, min(32, V - warp_v0));        // vertices of this warp that exist (last tile is ragged)
      { FZ_T0(); ptx::mbar_wait_nohint(bar_dfull + a, (i >> 1) & 1); FZ_ACC(fz_a0); }
      const uint32_t d_addr0 = tmem_base + kFzTmemD + lane_addr + a * 3 * kFzBodies + 2 * h;
      uint32_t dx[2], dy[2], dz[2];
      for (int sb = e; sb < kFzSubs; sb += 2) {
        const long long b0 = (long long)blk * kFzBodies + sb * kFzSub;
        SMPLB200_PROGRESS((i << 8) | sb);
        { FZ_T0(); ptx::mbar_wait_nohint(bar_tfull + e, t_cnt & 1); FZ_ACC(fz_a1); }
        ++t_cnt;
        ptx::tc_fence_after();
        uint32_t r0[16], r1[8];
        const uint32_t t_addr = tmem_base + kFzTmemT + lane_addr + e * kFzNT + h * 24;
        ptx::tmem_ld16(t_addr, r0);
        ptx::tmem_ld8(t_addr + 16, r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar_tempty + e);      // T buffer free: the next blend of this slot may start
        __syncwarp();
        ptx::tmem_ld2(d_addr0 + sb * kFzSub, dx);
        ptx::tmem_ld2(d_addr0 + sb * kFzSub + kFzBodies, dy);
        ptx::tmem_ld2(d_addr0 + sb * kFzSub + 2 * kFzBodies, dz);
        ptx::tmem_ld_wait();
        float T[24];
#pragma unroll
        for (int k = 0; k < 16; ++k) T[k] = __uint_as_float(r0[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) T[16 + k] = __uint_as_float(r1[k]);
        float res[6];
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const float* tt = T + bb * 12;
          const float x = __uint_as_float(dx[bb]), y = __uint_as_float(dy[bb]), z = __uint_as_float(dz[bb]);
          res[3 * bb] = fmaf(tt[2], z, fmaf(tt[1], y, fmaf(tt[0], x, tt[3])));
          res[3 * bb + 1] = fmaf(tt[6], z, fmaf(tt[5], y, fmaf(tt[4], x, tt[7])));
          res[3 * bb + 2] = fmaf(tt[10], z, fmaf(tt[9], y, fmaf(tt[8], x, tt[11])));
        }
        if (sb + 2 >= kFzSubs) {                                // this warp has read all it needs of the unit's D
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(bar_dempty + a);
          __syncwarp();
        }
        const long long bh = b0 + 2 * h;                       // first of this warp's two bodies
        // DIRECT stores: the lane writes its own x, y, z (12-byte stride across the warp).  The three store
        // instructions of a body together cover the warp's 384 contiguous bytes, so every 32-byte sector is
        // completed in L2 within a few cycles (ncu: DRAM bytes unchanged).  A/B on one box against the
        // alternatives -- xyz interleave through a shared-memory row + coalesced 128-byte stores (round 1's k3
        // epilogue): 143.8 us; interleave by register shuffles: 139.4 us; this: 136.7 us.
        if (lane < nv) {
          float* d = verts + ((size_t)bh * V + warp_v0 + lane) * 3;
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            if (bh + bb < n) { d[0] = res[3 * bb]; d[1] = res[3 * bb + 1]; d[2] = res[3 * bb + 2]; }
            d += body_stride;
          }
        }
      }
    }

    if (q == 0 && h == 0) { FZ_OUT(10 + 2 * e, fz_a0); FZ_OUT(11 + 2 * e, fz_a1); }
  }
  ptx::tc_fence_before();
  __syncthreads();
#ifdef SMPLB200_FZ_TIMING
  if (threadIdx.x == 0) g_fz_time[blockIdx.x * 32 + 31] += clock64() - t_kernel0;
#endif
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// fp32 coef [n,224] (old K order) and A [n,24,12] -> the fused kernel's operand images (stand-alone entry
// point only; inside smplb200_forward k2 writes the images directly).
__global__ void __launch_bounds__(256)
k_pack_fz(const float* __restrict__ coef, const float* __restrict__ A, long long n, int NB,
          uint8_t* __restrict__ coef_img, uint8_t* __restrict__ a_img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncoef = n * kCoefK, na = n * (kJ * 12);
  if (idx < ncoef) {
    const long long b = idx / kCoefK;
    const int nk = int(idx - b * kCoefK);
    const int ok = fz_old_index(nk, NB);
    const float v = ok < 0 ? 0.f : coef[b * kCoefK + ok];
    const uint16_t hi = f32_to_f16_rn(v);
    uint8_t* img = coef_img + (size_t)(b / kFzBodies) * kFzCoefBlock;
    const int row = int(b % kFzBodies);
    reinterpret_cast<uint16_t*>(img + kFzCoefLo)[(size_t)(nk >> 3) * (kFzBodies * 8) + row * 8 + (nk & 7)] = hi;
    if (nk < kFzShapeK)
      reinterpret_cast<uint16_t*>(img)[(size_t)(nk >> 3) * (kFzBodies * 8) + row * 8 + (nk & 7)] =
          f32_to_f16_rn(__fsub_rn(v, f16_to_f32(hi)));
  } else if (idx < ncoef + na) {
    const long long i = idx - ncoef;
    const long long b = i / (kJ * 12);
    const int r = int(i - b * (kJ * 12)), jj = r / 12, e = r % 12;
    const float v = A[i];
    const uint16_t hi = f32_to_f16_rn(v);
    const uint16_t lo = f32_to_f16_rn(__fsub_rn(v, f16_to_f32(hi)));
    uint16_t* img = reinterpret_cast<uint16_t*>(a_img + (size_t)(b / kFzSub) * kFzAImage);
    const int row = int(b % kFzSub) * 12 + e;
    img[(size_t)(jj >> 3) * (kFzNT * 8) + row * 8 + (jj & 7)] = hi;
    img[(size_t)(3 + (jj >> 3)) * (kFzNT * 8) + row * 8 + (jj & 7)] = lo;
    if (jj < 8) img[(size_t)6 * (kFzNT * 8) + row * 8 + jj] = 0;      // pad chunk (multiplied by zero weights)
  }
}

inline size_t fz_coef_image_bytes(long long n) {
  return (size_t)((std::max<long long>(n, 1) + kFzBodies - 1) / kFzBodies) * kFzCoefBlock;
}
inline size_t fz_a_image_bytes(long long n) {
  return (size_t)((std::max<long long>(n, 1) + kFzBodies - 1) / kFzBodies) * kFzSubs * kFzAImage;
}

inline cudaError_t fused_tc_set_smem() {
  return cudaFuncSetAttribute(k_fused_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFzSmemBytes);
}

inline cudaError_t launch_fused_tc(const DeviceModel& m, int num_sms, const uint8_t* coef_img,
                                   const uint8_t* a_img, long long n, float* verts, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int ntile = m.VP / 128;
  const int nblk = (int)((n + kFzBodies - 1) / kFzBodies);
  const long long total = (long long)ntile * nblk;
  const unsigned grid = (unsigned)std::min<long long>(num_sms, total);
  k_fused_tc<<<grid, kFzThreads, kFzSmemBytes, s>>>(m.fz_basis, m.fz_w, coef_img, a_img, n, nblk, total, m.V, verts);
  return cudaGetLastError();
}

}  // namespace smplb200
