// REJECTED EXPERIMENT (round 1) -- not compiled into libsmpl_b200.so; kept as the record of a measured dead end.
// Correct (38/38 backward parity tests passed with it enabled) but slower than the gather kernel it was to replace:
//   4096 bodies: k_lbs_bwd_vp 306 us + k_lbs_bwd_tc 570 us = 876 us   vs   k_lbs_bwd (gather) 603 us
//   1024 bodies:               84 us +              195 us = 279 us   vs                      ~150 us
// The 12 TMEM-lane rows of a body each re-load the same g / vposed values (4x / 3x redundant, ~170 L1
// wavefronts per warp per K-step), so the loaders, not the MMAs, bound it (~1670 clk per K-step vs ~384 of MMA).
// See DESIGN.md section 7.
// kb3 (large-batch path): skinning backward in two streaming kernels.
//
//   k_lbs_bwd_vp   g_vposed_v = T_R(v)^T g_v            thread = vertex, 16 bodies per CTA
//   k_lbs_bwd_tc   g_A[b]     = P_b^T W                  on tcgen05 / TMEM
//
// with P_b[v, e=(r,c)] = g_v[r] * [vposed_v, 1][c]  (12 products per body-vertex) and W the dense
// [V, 24] skinning weights.  The per-joint gather of the small-batch kernel (k_lbs_bwd) is bound by
// bank-conflicted shared-memory gathers (~24 B x 27.5 K entries per body); as a dense contraction
// over the vertices it is a tensor-core job whose operand is produced on the fly:
//
//   D[M = 128 rows (10 bodies x 12 e), N = 32 (24 joints + pad)] += A[128, 32 v] * B[32, 32 v]^T
//
//   * A: a loader thread owns one row (body, r, c); per K-step of 32 vertices it reads 32 strided
//     g values and one 128-byte line of a vposed plane, multiplies, splits the product hi | lo
//     (3xTF32: the joint gradients keep fp32-class accuracy) and writes its TMEM lane.
//   * B: W^T, pre-tiled at model create ([8 chunks][32 joint rows][4 v] per K-step, tf32 hi | lo),
//     one 8 KB bulk-TMA copy per K-step.
//   * D stays in TMEM for the whole vertex range; the epilogue scatters it to g_A[b, j, e].
//
// Warp roles (320 threads): warp 0 = bulk-TMA producer, warp 1 = MMA issuer, warps 2..9 = two
// groups of four loader warps (TMEM lane quarter = warp % 4), group g takes K-steps i = g mod 2 and
// keeps the next one's loads in flight; warps 2..5 run the epilogue.
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

// ---------------------------------------------------------------------------------------------
constexpr int kBwdVpBodies = 16;

__global__ void __launch_bounds__(kVertTile)
k_lbs_bwd_vp(DeviceModel m, const float* __restrict__ A, const float* __restrict__ g_verts, long long n,
             float* __restrict__ g_vposed) {
  __shared__ __align__(16) float s_A[kBwdVpBodies][kJ * 12];
  const int tid = threadIdx.x;
  const int v = blockIdx.x * kVertTile + tid;
  const long long b0 = (long long)blockIdx.y * kBwdVpBodies;
  const int nb = (int)min((long long)kBwdVpBodies, n - b0);
  for (int i = tid; i < nb * kJ * 12; i += kVertTile) (&s_A[0][0])[i] = __ldg(A + (size_t)b0 * (kJ * 12) + i);
  __syncthreads();
  const int V = m.V, VP = m.VP;
  const bool live = v < V;
  float ws[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t jj = 0;
  const bool ell = m.max_nnz <= 4;
  if (live && ell) {
    const float4 w4 = __ldg(m.ell_w + v);
    ws[0] = w4.x; ws[1] = w4.y; ws[2] = w4.z; ws[3] = w4.w;
    jj = __ldg(m.ell_j + v);
  }
  for (int bi = 0; bi < nb; ++bi) {
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    if (live) {
      const float* gp = g_verts + ((size_t)(b0 + bi) * V + v) * 3;
      const float g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
      float T[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) T[e] = 0.f;
      if (ell) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float* Aj = s_A[bi] + ((jj >> (8 * s)) & 0xffu) * 12;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(ws[s], Aj[4 * r + c], T[3 * r + c]);
        }
      } else {
        const float* wr = m.dense_w + (size_t)v * kJ;
        for (int j = 0; j < kJ; ++j) {
          const float w = __ldg(wr + j);
          if (w == 0.f) continue;
          const float* Aj = s_A[bi] + j * 12;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(w, Aj[4 * r + c], T[3 * r + c]);
        }
      }
      o0 = fmaf(T[6], g2, fmaf(T[3], g1, T[0] * g0));
      o1 = fmaf(T[7], g2, fmaf(T[4], g1, T[1] * g0));
      o2 = fmaf(T[8], g2, fmaf(T[5], g1, T[2] * g0));
    }
    float* dst = g_vposed + (size_t)(b0 + bi) * 3 * VP + v;
    dst[0] = o0; dst[VP] = o1; dst[2 * (size_t)VP] = o2;
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int kBwdLbsThreads = 320;
constexpr int kBwdLbsGroups = 2;
constexpr int kBwdLbsBodies = 10;                 // 10 x 12 = 120 of the 128 MMA rows
constexpr int kBwdLbsN = 32;                      // 24 joints padded to the MMA N granule
constexpr uint32_t kBwdLbsTile = 8u * kBwdLbsN * 16u;    // one W^T K-step tile (one part): 4,096 B
constexpr uint32_t kBwdLbsStage = 2u * kBwdLbsTile;      // hi | lo
constexpr int kBwdLbsStagesB = 8;
constexpr int kBwdLbsStagesA = 6;                 // 64 TMEM columns each (P_hi | P_lo), after D's 32
constexpr int kBwdLbsACol0 = 32;
constexpr uint32_t kBwdLbsBarOffset = kBwdLbsStagesB * kBwdLbsStage;
constexpr uint32_t kBwdLbsSmemBytes = kBwdLbsBarOffset + 512;
constexpr uint32_t kBwdLbsLbo = kBwdLbsN * 16, kBwdLbsSbo = 128;
constexpr uint32_t kBwdLbsIdesc = ptx::make_idesc(ptx::kFmtTF32, 128, kBwdLbsN);
static_assert(kBwdLbsACol0 + kBwdLbsStagesA * 64 <= 512, "TMEM budget");

__global__ void __launch_bounds__(kBwdLbsThreads, 1)
k_lbs_bwd_tc(const uint8_t* __restrict__ wimg /* [VP/32][hi|lo][8][32][4] tf32 */,
             const float* __restrict__ vposed, const float* __restrict__ g_verts, long long n, int V, int VP,
             float* __restrict__ g_A) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwdLbsBarOffset);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + kBwdLbsStagesB;
  uint64_t* a_full = b_empty + kBwdLbsStagesB;
  uint64_t* a_empty = a_full + kBwdLbsStagesA;
  uint64_t* d_full = a_empty + kBwdLbsStagesA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b0 = (long long)blockIdx.x * kBwdLbsBodies;
  const int nks = VP / 32;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kBwdLbsStagesB; ++s) { ptx::mbar_init(b_full + s, 1); ptx::mbar_init(b_empty + s, 1); }
    for (int s = 0; s < kBwdLbsStagesA; ++s) { ptx::mbar_init(a_full + s, 4); ptx::mbar_init(a_empty + s, 1); }
    ptx::mbar_init(d_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nks; ++i) {
        const int s = i % kBwdLbsStagesB;
        ptx::mbar_wait(b_empty + s, ((i / kBwdLbsStagesB) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(b_full + s, kBwdLbsStage);
        ptx::bulk_g2s(sB + (size_t)s * kBwdLbsStage, wimg + (size_t)i * kBwdLbsStage, kBwdLbsStage, b_full + s);
      }
    }
  } else if (warp == 1) {
    for (int i = 0; i < nks; ++i) {
      const int sb = i % kBwdLbsStagesB, sa = i % kBwdLbsStagesA;
      ptx::mbar_wait(b_full + sb, (i / kBwdLbsStagesB) & 1);
      ptx::mbar_wait(a_full + sa, (i / kBwdLbsStagesA) & 1);
      ptx::tc_fence_after();
      const uint32_t b_addr = ptx::smem_u32(sB + (size_t)sb * kBwdLbsStage);
      const uint32_t a_addr = tmem_base + kBwdLbsACol0 + sa * 64;
      if (ptx::elect_one()) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {       // (P_hi, W_hi), (P_lo, W_hi), (P_hi, W_lo)
          const uint32_t ap = a_addr + (g == 1 ? 32 : 0);
          const uint32_t bp = b_addr + (g == 2 ? kBwdLbsTile : 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bd = ptx::make_smem_desc(bp + kk * 2 * kBwdLbsLbo, kBwdLbsLbo, kBwdLbsSbo);
            ptx::mma_tf32_ts(tmem_base, ap + kk * 8, bd, kBwdLbsIdesc, (uint32_t)((i | g | kk) != 0));
          }
        }
        ptx::tc_commit(b_empty + sb);
        ptx::tc_commit(a_empty + sa);
        if (i == nks - 1) ptx::tc_commit(d_full);
      }
      __syncwarp();
    }
  } else {
    // ===== loaders: row = (body, r, c) -> products g_v[r] * [vposed_v, 1][c] for 32 vertices =====
    const int lw = warp - 2, q = warp & 3, grp = lw >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    const int bl = row / 12, e = row - 12 * bl, r = e >> 2, c = e & 3;
    const long long b = b0 + bl;
    const bool valid = row < kBwdLbsBodies * 12 && b < n;
    const float* gsrc = g_verts + (size_t)(valid ? b : 0) * V * 3 + r;
    const float4* psrc = reinterpret_cast<const float4*>(vposed + ((size_t)(valid ? b : 0) * 3 + (c < 3 ? c : 0)) * VP);
    float gc[32], gn[32];
    float4 pc[8], pn[8];
    auto load = [&](float (&g)[32], float4 (&p)[8], int i) {
      const int v0 = i * 32;
#pragma unroll
      for (int u = 0; u < 32; ++u) g[u] = (valid && v0 + u < V) ? __ldg(gsrc + (size_t)(v0 + u) * 3) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        p[u] = (valid && c < 3) ? __ldg(psrc + (size_t)i * 8 + u) : make_float4(1.f, 1.f, 1.f, 1.f);
    };
    if (grp < nks) load(gc, pc, grp);
    for (int i = grp; i < nks; i += kBwdLbsGroups) {
      if (i + kBwdLbsGroups < nks) load(gn, pn, i + kBwdLbsGroups);
      const int sa = i % kBwdLbsStagesA;
      ptx::mbar_wait(a_empty + sa, ((i / kBwdLbsStagesA) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t acol = tmem_base + lane_addr + kBwdLbsACol0 + sa * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 p = pc[4 * h + u];
          const float ps[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float x = gc[16 * h + 4 * u + t] * ps[t];
            hi[4 * u + t] = f32_to_tf32_rn(x);
            lo[4 * u + t] = f32_to_tf32_rn(x - __uint_as_float(hi[4 * u + t]));
          }
        }
        ptx::tmem_st16(acol + 16 * h, hi);
        ptx::tmem_st16(acol + 32 + 16 * h, lo);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_full + sa);
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 32; ++u) gc[u] = gn[u];
#pragma unroll
      for (int u = 0; u < 8; ++u) pc[u] = pn[u];
    }
    if (lw < 4) {
      // ===== epilogue: D[row = (body, e), col = joint] -> g_A[body, joint, e] =====
      ptx::mbar_wait(d_full, 0);
      ptx::tc_fence_after();
      uint32_t d[32];
      ptx::tmem_ld32(tmem_base + lane_addr, d);
      ptx::tmem_ld_wait();
      if (valid) {
        float* dst = g_A + (size_t)b * (kJ * 12) + e;
#pragma unroll
        for (int j = 0; j < kJ; ++j) dst[j * 12] = __uint_as_float(d[j]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

inline cudaError_t launch_lbs_bwd_tc(const DeviceModel& m, const float* vposed, const float* A,
                                     const float* g_verts, long long n, float* g_vposed, float* g_A,
                                     cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  dim3 g1((unsigned)(m.VP / kVertTile), (unsigned)((n + kBwdVpBodies - 1) / kBwdVpBodies));
  k_lbs_bwd_vp<<<g1, kVertTile, 0, s>>>(m, A, g_verts, n, g_vposed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  k_lbs_bwd_tc<<<(unsigned)((n + kBwdLbsBodies - 1) / kBwdLbsBodies), kBwdLbsThreads, kBwdLbsSmemBytes, s>>>(
      reinterpret_cast<const uint8_t*>(m.bwd_w_tf32), vposed, g_verts, n, m.V, m.VP, g_A);
  return cudaGetLastError();
}

}  // namespace smplb200
