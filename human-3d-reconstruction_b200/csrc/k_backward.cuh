// Backward pass of the SMPL layer ("what comes next" after the forward):
//   (g_vertices, g_joints, g_kp2d)  ->  (g_betas, g_pose, g_cam)
// so the layer can sit inside the reference trainer's loss (reference src/lib/trains/trainer.py:31-37
// computes loss = model_with_loss(batch) and calls loss.backward() at :102-104).
//
// Reverse of the forward kernels.  A and vposed are either recomputed (k2 + k1 run again) or, when the
// caller kept the forward's workspace, read from it:
//
//   kb3  skinning:        g_vposed = T_R^T g_v  (thread per vertex),  g_A[j] = sum_v w_vj g_v (x) [vposed_v, 1]
//        k_lbs_bwd_split  the default: both halves as decoupled warp groups of a persistent CTA;
//        k_lbs_bwd<false> lock-step variant without shared-memory staging, for meshes that do not fit
//   kb1  blendshapes:     g_coef = g_vposed . basis^T, split over column slices:
//        k_blend_bwd_tc   (k_blend_bwd_tc.cuh) tcgen05, the default;
//        k_blend_bwd_fma  (here) CUDA cores, only for an explicit precision='fp32' below 256 bodies
//   kb2  k_chain_bwd      chain + regressor + Rodrigues + projection, one warp per body
//
// The derivation is stated step by step in float64 numpy in oracle/smpl_backward_np.py and pinned
// against torch autograd of the oracle by tests/test_oracle_backward.py; this file follows it.
// All sums run in a fixed order (no atomics): gradients are bitwise reproducible run to run.
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

// ---------------------------------------------------------------------------------------------
// kb3: skinning backward, one CTA per body (persistent over bodies).
//   STAGED: one thread bulk-TMA-stages the body's upstream gradient g_v [V,3] and its vposed planes
//   [3,VP] into shared memory (166 KB at V = 6890).  Phase 1 (thread = vertex) runs while the copies
//   land (it reads g_v from global, coalesced); phase 2 (warp = joint, lanes over that joint's
//   skinned vertices) gathers from the staged copy.  Unstaged variant for meshes that do not fit.
// ---------------------------------------------------------------------------------------------
constexpr int kLbsBwdThreads = 768;   // 24 warps: in phase 2 warp j owns joint j

struct LbsBwdArgs {
  const float* vposed;     // [n,3,VP] planar (recomputed)
  const float* A;          // [n,24,12]      (recomputed)
  const float* g_verts;    // [n,V,3] or null
  const float* g_joints;   // [n,24,3] or null  (read only when joints are regressed from vertices)
  const float* g_kp2d;     // [n,24,2] or null  (likewise)
  const float* cam;        // [n,3] or null
  float* g_vposed;         // [n,3,VP] planar out (padding columns written as 0)
  float* g_A;              // [n,24,12] out
  int regressed;
};

inline size_t lbs_bwd_smem_bytes(int V, int VP) {
  // vposed planes | g_v slab (+ up to 15 B of alignment slack on either side of the bulk copy)
  return ((size_t)3 * VP + (size_t)((3 * V + 3) / 4 * 4) + 8) * sizeof(float);
}

template <bool STAGED>
__global__ void __launch_bounds__(kLbsBwdThreads, 1)
k_lbs_bwd(DeviceModel m, LbsBwdArgs a, long long n) {
  extern __shared__ __align__(128) float smem_bw[];
  // The body's 24 transforms stored TRANSPOSED, [entry e][joint j]: the lanes of a warp gather by
  // their own joint index, and with j fastest distinct joints are distinct banks and equal joints
  // broadcast, so the 9 LDS.32 per slot are conflict free (row-major [j][12] read with LDS.128 cost
  // ~10 wavefronts per load for random joints, round-1 ncu).
  __shared__ float s_At[12 * kJ];
  __shared__ float s_gj[kJ * 3];
  __shared__ __align__(8) uint64_t s_bar;
  float* s_vp = smem_bw;                 // [3][VP]
  float* s_gbuf = smem_bw + 3 * m.VP;    // 16-byte aligned landing zone of the g_v slab
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = m.V, VP = m.VP;
  if (STAGED) {
    if (tid == 0) { ptx::mbar_init(&s_bar, 1); ptx::fence_barrier_init(); }
    __syncthreads();
  }
  uint32_t phase = 0;

  for (long long b = blockIdx.x; b < n; b += gridDim.x) {
    __syncthreads();                     // the previous body's phase 2 is done with shared memory
    const float* gv = a.g_verts ? a.g_verts + (size_t)b * V * 3 : nullptr;
    const float* vp = a.vposed + (size_t)b * 3 * VP;
    // the body's g_v slab starts at a 4-byte aligned address: copy the enclosing 16-byte granules
    const uint32_t shift = gv ? (uint32_t)(reinterpret_cast<uintptr_t>(gv) & 15u) : 0u;
    const float* s_g = s_gbuf + shift / 4;
    if (STAGED && tid == 0) {
      // one thread stages the whole body with bulk-TMA copies: no load instructions, no registers
      const uint32_t vp_bytes = 3u * (uint32_t)VP * 4u;
      const uint32_t g_bytes = gv ? ((shift + (uint32_t)V * 12u + 15u) & ~15u) : 0u;
      ptx::fence_proxy_async();          // the previous body's generic-proxy accesses come first
      ptx::mbar_arrive_expect_tx(&s_bar, vp_bytes + g_bytes);
      ptx::bulk_g2s_split(s_vp, vp, vp_bytes, &s_bar);
      if (gv) ptx::bulk_g2s_split(s_gbuf, reinterpret_cast<const uint8_t*>(gv) - shift, g_bytes, &s_bar);
    }
    if (tid < kJ * 12) {
      const int j = tid / 12, e = tid - 12 * j;
      s_At[e * kJ + j] = __ldg(a.A + (size_t)b * (kJ * 12) + tid);
    }
    if (tid >= 640 && tid < 640 + kJ * 3) {
      // effective joint gradient that flows into the VERTICES (regressed joints only):
      // g_joints + s * g_kp2d on x,y
      const int i = tid - 640, j = i / 3, c = i - 3 * j;
      float v = 0.f;
      if (a.regressed) {
        if (a.g_joints) v = __ldg(a.g_joints + (size_t)b * (kJ * 3) + i);
        if (a.g_kp2d && c < 2) v = fmaf(__ldg(a.cam + (size_t)b * 3), __ldg(a.g_kp2d + (size_t)b * (kJ * 2) + 2 * j + c), v);
      }
      s_gj[i] = v;
    }
    if (STAGED && !gv)
      for (int i = tid; i < 3 * V; i += kLbsBwdThreads) s_gbuf[i] = 0.f;
    __syncthreads();
    // regressed joints: g_v += J_regressor[v,:] . g_joint
    auto add_regressed = [&](int v, float g[3]) {
      const float* jr = m.dense_jreg + (size_t)v * kJ;
      for (int j = 0; j < kJ; ++j) {
        const float r = __ldg(jr + j);
        if (r != 0.f) {
          g[0] = fmaf(r, s_gj[3 * j], g[0]); g[1] = fmaf(r, s_gj[3 * j + 1], g[1]);
          g[2] = fmaf(r, s_gj[3 * j + 2], g[2]);
        }
      }
    };
    // Without the regressed-joint term phase 1 needs nothing from shared memory but A: it reads
    // g_v straight from global (coalesced) WHILE the bulk copies land, and the wait moves to phase 2.
    const bool early = STAGED && !a.regressed;
    if (STAGED && !early) {
      ptx::mbar_wait(&s_bar, phase);
      phase ^= 1;
      float* sg = s_gbuf + shift / 4;
      for (int v = tid; v < V; v += kLbsBwdThreads) {
        float g[3] = {sg[3 * v], sg[3 * v + 1], sg[3 * v + 2]};
        add_regressed(v, g);
        sg[3 * v] = g[0]; sg[3 * v + 1] = g[1]; sg[3 * v + 2] = g[2];
      }
      __syncthreads();
    }
    auto g_of = [&](int v, float g[3]) {
      if (STAGED) { g[0] = s_g[3 * v]; g[1] = s_g[3 * v + 1]; g[2] = s_g[3 * v + 2]; }
      else {
        g[0] = gv ? __ldg(gv + 3 * v) : 0.f; g[1] = gv ? __ldg(gv + 3 * v + 1) : 0.f;
        g[2] = gv ? __ldg(gv + 3 * v + 2) : 0.f;
        if (a.regressed) add_regressed(v, g);
      }
    };
    auto vp_of = [&](int c, int v) -> float { return STAGED ? s_vp[c * VP + v] : __ldg(vp + (size_t)c * VP + v); };

    // ---- phase 1: g_vposed_v = T_R(v)^T g_v with T_R = sum_j w_vj Rw_j
#pragma unroll 3
    for (int v = tid; v < VP; v += kLbsBwdThreads) {
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      if (v < V) {
        float g[3];
        if (early) {
          g[0] = gv ? __ldg(gv + 3 * v) : 0.f; g[1] = gv ? __ldg(gv + 3 * v + 1) : 0.f;
          g[2] = gv ? __ldg(gv + 3 * v + 2) : 0.f;
        } else {
          g_of(v, g);
        }
        float T[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) T[e] = 0.f;
        if (m.max_nnz <= 4) {
          const float4 w4 = __ldg(m.ell_w + v);
          const uint32_t jj = __ldg(m.ell_j + v);
          const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            const float* Aj = s_At + ((jj >> (8 * s)) & 0xffu);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(ws[s], Aj[(4 * r + c) * kJ], T[3 * r + c]);
          }
        } else {
          const float* wr = m.dense_w + (size_t)v * kJ;
          for (int j = 0; j < kJ; ++j) {
            const float w = __ldg(wr + j);
            if (w == 0.f) continue;
            const float* Aj = s_At + j;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(w, Aj[(4 * r + c) * kJ], T[3 * r + c]);
          }
        }
        o0 = fmaf(T[6], g[2], fmaf(T[3], g[1], T[0] * g[0]));
        o1 = fmaf(T[7], g[2], fmaf(T[4], g[1], T[1] * g[0]));
        o2 = fmaf(T[8], g[2], fmaf(T[5], g[1], T[2] * g[0]));
      }
      float* dst = a.g_vposed + (size_t)b * 3 * VP + v;
      dst[0] = o0; dst[VP] = o1; dst[2 * (size_t)VP] = o2;
    }

    if (early) {
      ptx::mbar_wait(&s_bar, phase);
      phase ^= 1;
    }
    // ---- phase 2: g_A[j] = sum_{v in skin(j)} w_vj * g_v (x) [vposed_v, 1]   (warp j)
    // Lane l takes entries beg + l, + 32, ...; four entries are fetched per trip so the index /
    // weight loads (L2) and the shared-memory gathers of four entries are in flight together.
    {
      const int beg = __ldg(m.wcsr_ptr + warp), end = __ldg(m.wcsr_ptr + warp + 1);
      float acc[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) acc[e] = 0.f;
      for (int i0 = beg + lane; i0 < end; i0 += 4 * 32) {
        int vi[4]; float wi[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 32 * u;
          const bool ok = i < end;
          vi[u] = ok ? __ldg(m.wcsr_idx + i) : 0;
          wi[u] = ok ? __ldg(m.wcsr_val + i) : 0.f;
        }
        float gg[4][3], pp[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          g_of(vi[u], gg[u]);
          pp[u][0] = vp_of(0, vi[u]); pp[u][1] = vp_of(1, vi[u]); pp[u][2] = vp_of(2, vi[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {     // entries in ascending order: same sum order as one-at-a-time
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float wg = wi[u] * gg[u][r];
            acc[4 * r] = fmaf(wg, pp[u][0], acc[4 * r]);
            acc[4 * r + 1] = fmaf(wg, pp[u][1], acc[4 * r + 1]);
            acc[4 * r + 2] = fmaf(wg, pp[u][2], acc[4 * r + 2]);
            acc[4 * r + 3] += wg;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
      }
      if (lane < 12) {
        float v = acc[0];
#pragma unroll
        for (int e = 1; e < 12; ++e) v = lane == e ? acc[e] : v;
        a.g_A[(size_t)b * (kJ * 12) + warp * 12 + lane] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kb3, default (the body fits shared memory): the two halves of the skinning backward run as two
// DECOUPLED warp groups of one persistent CTA, each looping over the CTA's bodies at its own pace:
//   group P1 (warps 0..11)   g_vposed = T_R^T g_v, thread = vertex, reads g_v straight from global;
//                            needs no staging at all, only the 24 transforms of the body;
//   group P2 (warps 12..23)  g_A[j] = sum_v w_vj g_v (x) [vposed_v, 1] from the bulk-TMA-staged copy
//                            of g_v and vposed (warp w reduces joints w and w + 12).
// The groups never synchronise with each other (named barriers 1 and 2), so P1's global-load
// latency overlaps P2's shared-memory gathers and the staging wait.  In the first version all 24
// warps ran stage -> phase 1 -> phase 2 in lock step (ncu: long_scoreboard + barrier stalls on top,
// LSU pipe 63 %, issue 43 %).
// ---------------------------------------------------------------------------------------------
constexpr int kLbsBwdGroupThreads = kLbsBwdThreads / 2;   // 384

__global__ void __launch_bounds__(kLbsBwdThreads, 1)
k_lbs_bwd_split(DeviceModel m, LbsBwdArgs a, long long n) {
  extern __shared__ __align__(128) float smem_bw[];
  __shared__ float s_At[12 * kJ];          // P1: transforms, transposed [entry][joint] (conflict-free gathers)
  __shared__ float s_gj1[kJ * 3], s_gj2[kJ * 3];   // regressed joints: gradient flowing into the vertices
  __shared__ __align__(8) uint64_t s_bar;
  float* s_vp = smem_bw;                   // [3][VP]
  float* s_gbuf = smem_bw + 3 * m.VP;      // 16-byte aligned landing zone of the g_v slab
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = m.V, VP = m.VP;
  if (tid == 0) { ptx::mbar_init(&s_bar, 1); ptx::fence_barrier_init(); }
  __syncthreads();

  // effective joint gradient that flows into the VERTICES (regressed joints only): g_joints + s * g_kp2d
  auto joint_term = [&](long long b, int i) -> float {
    const int j = i / 3, c = i - 3 * j;
    float v = 0.f;
    if (a.g_joints) v = __ldg(a.g_joints + (size_t)b * (kJ * 3) + i);
    if (a.g_kp2d && c < 2) v = fmaf(__ldg(a.cam + (size_t)b * 3), __ldg(a.g_kp2d + (size_t)b * (kJ * 2) + 2 * j + c), v);
    return v;
  };
  auto add_regressed = [&](const float* gj, int v, float g[3]) {      // g_v += J_regressor[v,:] . g_joint
    const float* jr = m.dense_jreg + (size_t)v * kJ;
    for (int j = 0; j < kJ; ++j) {
      const float r = __ldg(jr + j);
      if (r != 0.f) {
        g[0] = fmaf(r, gj[3 * j], g[0]); g[1] = fmaf(r, gj[3 * j + 1], g[1]); g[2] = fmaf(r, gj[3 * j + 2], g[2]);
      }
    }
  };

  if (warp < kLbsBwdGroupThreads / 32) {
    // ===================== group P1 =====================
    for (long long b = blockIdx.x; b < n; b += gridDim.x) {
      asm volatile("bar.sync 1, %0;" ::"n"(kLbsBwdGroupThreads) : "memory");     // previous body's table reads done
      if (tid < kJ * 12) {
        const int j = tid / 12, e = tid - 12 * j;
        s_At[e * kJ + j] = __ldg(a.A + (size_t)b * (kJ * 12) + tid);
      } else if (a.regressed && tid < kJ * 12 + kJ * 3) {
        s_gj1[tid - kJ * 12] = joint_term(b, tid - kJ * 12);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kLbsBwdGroupThreads) : "memory");
      const float* gv = a.g_verts ? a.g_verts + (size_t)b * V * 3 : nullptr;
#pragma unroll 3
      for (int v = tid; v < VP; v += kLbsBwdGroupThreads) {
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (v < V) {
          float g[3];
          g[0] = gv ? __ldg(gv + 3 * v) : 0.f; g[1] = gv ? __ldg(gv + 3 * v + 1) : 0.f;
          g[2] = gv ? __ldg(gv + 3 * v + 2) : 0.f;
          if (a.regressed) add_regressed(s_gj1, v, g);
          float T[9];
#pragma unroll
          for (int e = 0; e < 9; ++e) T[e] = 0.f;
          if (m.max_nnz <= 4) {
            const float4 w4 = __ldg(m.ell_w + v);
            const uint32_t jj = __ldg(m.ell_j + v);
            const float ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const float* Aj = s_At + ((jj >> (8 * s)) & 0xffu);
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(ws[s], Aj[(4 * r + c) * kJ], T[3 * r + c]);
            }
          } else {
            const float* wr = m.dense_w + (size_t)v * kJ;
            for (int j = 0; j < kJ; ++j) {
              const float w = __ldg(wr + j);
              if (w == 0.f) continue;
              const float* Aj = s_At + j;
#pragma unroll
              for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) T[3 * r + c] = fmaf(w, Aj[(4 * r + c) * kJ], T[3 * r + c]);
            }
          }
          o0 = fmaf(T[6], g[2], fmaf(T[3], g[1], T[0] * g[0]));
          o1 = fmaf(T[7], g[2], fmaf(T[4], g[1], T[1] * g[0]));
          o2 = fmaf(T[8], g[2], fmaf(T[5], g[1], T[2] * g[0]));
        }
        float* dst = a.g_vposed + (size_t)b * 3 * VP + v;
        dst[0] = o0; dst[VP] = o1; dst[2 * (size_t)VP] = o2;
      }
    }
  } else {
    // ===================== group P2 =====================
    const int t2 = tid - kLbsBwdGroupThreads, w2 = warp - kLbsBwdGroupThreads / 32;
    uint32_t phase = 0;
    for (long long b = blockIdx.x; b < n; b += gridDim.x) {
      asm volatile("bar.sync 2, %0;" ::"n"(kLbsBwdGroupThreads) : "memory");     // previous body's gathers done
      const float* gv = a.g_verts ? a.g_verts + (size_t)b * V * 3 : nullptr;
      // the body's g_v slab starts at a 4-byte aligned address: copy the enclosing 16-byte granules
      const uint32_t shift = gv ? (uint32_t)(reinterpret_cast<uintptr_t>(gv) & 15u) : 0u;
      float* s_g = s_gbuf + shift / 4;
      if (t2 == 0) {      // one thread stages the whole body with bulk-TMA copies
        const uint32_t vp_bytes = 3u * (uint32_t)VP * 4u;
        const uint32_t g_bytes = gv ? ((shift + (uint32_t)V * 12u + 15u) & ~15u) : 0u;
        ptx::fence_proxy_async();          // the previous body's generic-proxy accesses come first
        ptx::mbar_arrive_expect_tx(&s_bar, vp_bytes + g_bytes);
        ptx::bulk_g2s_split(s_vp, a.vposed + (size_t)b * 3 * VP, vp_bytes, &s_bar);
        if (gv) ptx::bulk_g2s_split(s_gbuf, reinterpret_cast<const uint8_t*>(gv) - shift, g_bytes, &s_bar);
      }
      if (a.regressed && t2 >= 32 && t2 < 32 + kJ * 3) s_gj2[t2 - 32] = joint_term(b, t2 - 32);
      if (!gv)
        for (int i = t2; i < 3 * V; i += kLbsBwdGroupThreads) s_gbuf[i] = 0.f;
      ptx::mbar_wait(&s_bar, phase);
      phase ^= 1;
      if (a.regressed || !gv) {
        asm volatile("bar.sync 2, %0;" ::"n"(kLbsBwdGroupThreads) : "memory");   // s_gj2 / zero fill visible
        if (a.regressed) {
          for (int v = t2; v < V; v += kLbsBwdGroupThreads) {
            float g[3] = {s_g[3 * v], s_g[3 * v + 1], s_g[3 * v + 2]};
            add_regressed(s_gj2, v, g);
            s_g[3 * v] = g[0]; s_g[3 * v + 1] = g[1]; s_g[3 * v + 2] = g[2];
          }
          asm volatile("bar.sync 2, %0;" ::"n"(kLbsBwdGroupThreads) : "memory");
        }
      }
      // g_A[j] = sum_{v in skin(j)} w_vj * g_v (x) [vposed_v, 1]: warp w2 takes joints w2 and w2 + 12
      for (int j = w2; j < kJ; j += kLbsBwdGroupThreads / 32) {
        const int beg = __ldg(m.wcsr_ptr + j), end = __ldg(m.wcsr_ptr + j + 1);
        float acc[12];
#pragma unroll
        for (int e = 0; e < 12; ++e) acc[e] = 0.f;
        for (int i0 = beg + lane; i0 < end; i0 += 4 * 32) {
          int vi[4]; float wi[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
            const bool ok = i < end;
            vi[u] = ok ? __ldg(m.wcsr_idx + i) : 0;
            wi[u] = ok ? __ldg(m.wcsr_val + i) : 0.f;
          }
          float gg[4][3], pp[4][3];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            gg[u][0] = s_g[3 * vi[u]]; gg[u][1] = s_g[3 * vi[u] + 1]; gg[u][2] = s_g[3 * vi[u] + 2];
            pp[u][0] = s_vp[vi[u]]; pp[u][1] = s_vp[VP + vi[u]]; pp[u][2] = s_vp[2 * VP + vi[u]];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {     // entries in ascending order: same sum order as one-at-a-time
#pragma unroll
            for (int r = 0; r < 3; ++r) {
              const float wg = wi[u] * gg[u][r];
              acc[4 * r] = fmaf(wg, pp[u][0], acc[4 * r]);
              acc[4 * r + 1] = fmaf(wg, pp[u][1], acc[4 * r + 1]);
              acc[4 * r + 2] = fmaf(wg, pp[u][2], acc[4 * r + 2]);
              acc[4 * r + 3] += wg;
            }
          }
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        }
        if (lane < 12) {
          float v = acc[0];
#pragma unroll
          for (int e = 1; e < 12; ++e) v = lane == e ? acc[e] : v;
          a.g_A[(size_t)b * (kJ * 12) + j * 12 + lane] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// kb1: g_coef[b,k] = sum_col g_vposed[b,col] * basis[k,col]  (k < NB + 207; the template row has a
// constant coefficient).  CUDA-core kernel: CTA = 16 bodies x all 224 k x one slice of the planar
// columns; the slices' partial sums are added, in slice order, by kb2.
// ---------------------------------------------------------------------------------------------
constexpr int kBbBodies = 16, kBbCols = 64, kBbThreads = 256, kBbKPerThread = 7;
constexpr int kBbGStride = kBbCols + 4;   // +4 floats: the 8 body rows of a quarter-warp hit 8 bank quads
constexpr size_t kBbSmemBytes = ((size_t)kCoefK * kBbCols + (size_t)kBbBodies * kBbGStride) * sizeof(float);
constexpr int kBbMaxSlices = 81;

__global__ void __launch_bounds__(kBbThreads, 2)
k_blend_bwd_fma(DeviceModel m, const float* __restrict__ g_vposed, long long n, int slices,
                float* __restrict__ part /* [slices][n][224] */) {
  extern __shared__ __align__(16) float smem_bb[];
  float* s_b = smem_bb;                           // [224][64]
  float* s_g = smem_bb + kCoefK * kBbCols;        // [16][68]
  const int tid = threadIdx.x, bq = tid & 7, kq = tid >> 3;
  const int slice = blockIdx.x;
  const long long b0 = (long long)blockIdx.y * kBbBodies;
  const int nb = (int)min((long long)kBbBodies, n - b0);
  const int nchunks = m.NC / kBbCols;             // NC = 3 * VP, VP % 128 == 0
  const int c_beg = (int)((long long)slice * nchunks / slices), c_end = (int)((long long)(slice + 1) * nchunks / slices);
  const int krows = m.NB + kP;                    // rows that carry a gradient
  float acc0[kBbKPerThread], acc1[kBbKPerThread];
#pragma unroll
  for (int i = 0; i < kBbKPerThread; ++i) acc0[i] = acc1[i] = 0.f;

  for (int ch = c_beg; ch < c_end; ++ch) {
    const int col0 = ch * kBbCols;
    for (int idx = tid; idx < kCoefK * (kBbCols / 4); idx += kBbThreads) {
      const int k = idx / (kBbCols / 4), q = idx % (kBbCols / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < krows) v = __ldg(reinterpret_cast<const float4*>(m.basis + (size_t)k * m.NC + col0) + q);
      reinterpret_cast<float4*>(s_b + k * kBbCols)[q] = v;
    }
    {
      const int bi = tid / (kBbCols / 4), q = tid % (kBbCols / 4);   // 16 x 16 float4 == 256 threads
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bi < nb) v = __ldg(reinterpret_cast<const float4*>(g_vposed + (size_t)(b0 + bi) * m.NC + col0) + q);
      reinterpret_cast<float4*>(s_g + bi * kBbGStride)[q] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int q = 0; q < kBbCols / 4; ++q) {
      const float4 g0 = reinterpret_cast<const float4*>(s_g + bq * kBbGStride)[q];
      const float4 g1 = reinterpret_cast<const float4*>(s_g + (bq + 8) * kBbGStride)[q];
#pragma unroll
      for (int i = 0; i < kBbKPerThread; ++i) {
        const float4 bv = reinterpret_cast<const float4*>(s_b + (kq * kBbKPerThread + i) * kBbCols)[q];
        acc0[i] = fmaf(g0.w, bv.w, fmaf(g0.z, bv.z, fmaf(g0.y, bv.y, fmaf(g0.x, bv.x, acc0[i]))));
        acc1[i] = fmaf(g1.w, bv.w, fmaf(g1.z, bv.z, fmaf(g1.y, bv.y, fmaf(g1.x, bv.x, acc1[i]))));
      }
    }
    __syncthreads();
  }
  float* dst = part + ((size_t)slice * n + b0) * kCoefK + kq * kBbKPerThread;
  if (bq < nb) {
#pragma unroll
    for (int i = 0; i < kBbKPerThread; ++i) dst[(size_t)bq * kCoefK + i] = acc0[i];
  }
  if (bq + 8 < nb) {
#pragma unroll
    for (int i = 0; i < kBbKPerThread; ++i) dst[(size_t)(bq + 8) * kCoefK + i] = acc1[i];
  }
}

// ---------------------------------------------------------------------------------------------
// kb2: kinematic chain, folded regressor, Rodrigues and projection backward; one warp per body,
// lane j = joint j.  The forward chain is recomputed in registers, then gradients flow from the
// leaves to the root level by level: children post their contribution to shared memory, parents
// add their children's in ascending joint order.
// ---------------------------------------------------------------------------------------------
struct ChainBwdArgs {
  const float* betas;        // [n,NB]
  const float* pose;         // [n,72]
  const float* cam;          // [n,3] or null
  const float* g_A;          // [n,24,12] or null (no vertex path)
  const float* g_coef_part;  // [slices][n][224] or null
  int slices;
  const float* g_joints;     // [n,24,3] or null
  const float* g_kp2d;       // [n,24,2] or null
  const float* joints_fwd;   // [n,24,3] joints of the forward pass (regressed joints + g_kp2d only)
  float* g_betas;            // [n,NB]
  float* g_pose;             // [n,72]
  float* g_cam;              // [n,3] or null
  int rotate_base;
  int regressed;
};

// dL/dtheta from dL/dR for the HMR-idiom Rodrigues (oracle/smpl_backward_np.py: rodrigues_backward)
__device__ __forceinline__ void rodrigues_bwd(float tx, float ty, float tz, const float g[9], float out[3]) {
  const float eps = 1e-8f;
  const float ex = tx + eps, ey = ty + eps, ez = tz + eps;
  const float a = sqrtf(ex * ex + ey * ey + ez * ez);
  const float inv_a = 1.f / a;
  const float nx = tx * inv_a, ny = ty * inv_a, nz = tz * inv_a;
  float s, c;
  sincosf(0.5f * a, &s, &c);
  float w = c, x = s * nx, y = s * ny, z = s * nz;
  const float inv_q = rsqrtf(w * w + x * x + y * y + z * z);
  w *= inv_q; x *= inv_q; y *= inv_q; z *= inv_q;
  float gw = 2.f * w * (g[0] + g[4] + g[8]) + 2.f * (-z * g[1] + y * g[2] + z * g[3] - x * g[5] - y * g[6] + x * g[7]);
  float gx = 2.f * x * (g[0] - g[4] - g[8]) + 2.f * (y * g[1] + z * g[2] + y * g[3] - w * g[5] + z * g[6] + w * g[7]);
  float gy = 2.f * y * (-g[0] + g[4] - g[8]) + 2.f * (x * g[1] + w * g[2] + x * g[3] + z * g[5] - w * g[6] + z * g[7]);
  float gz = 2.f * z * (-g[0] - g[4] + g[8]) + 2.f * (-w * g[1] + x * g[2] + w * g[3] + y * g[5] + x * g[6] + y * g[7]);
  const float d = gw * w + gx * x + gy * y + gz * z;      // through q / |q|
  gw -= d * w; gx -= d * x; gy -= d * y; gz -= d * z;
  const float gn = gx * nx + gy * ny + gz * nz;
  const float k1 = s * inv_a, k0 = 0.5f * c * gn - 0.5f * s * gw - k1 * gn;
  out[0] = fmaf(k0, nx, k1 * gx);
  out[1] = fmaf(k0, ny, k1 * gy);
  out[2] = fmaf(k0, nz, k1 * gz);
}

__global__ void __launch_bounds__(kChainWarps * 32)
k_chain_bwd(DeviceModel m, ChainBwdArgs a, long long n) {
  __shared__ float s_c[kChainWarps][kJ][16];   // child -> parent: gRw(9) | gtw(3) | gJr(3)
  __shared__ float s_gc[kChainWarps][kCoefK];  // g_coef of this body: the column slices added in order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kChainWarps + warp;
  if (b >= n) return;   // whole warp exits together
  const unsigned full = 0xffffffffu;
  const int j = lane < kJ ? lane : 0;
  const bool active = lane < kJ;
  const int NB = m.NB;

  // g_coef = sum over the column slices of kb1's partials (coalesced, slice order, 7 sums per lane)
  if (a.g_coef_part) {
    float acc[kCoefK / 32];
#pragma unroll
    for (int t = 0; t < kCoefK / 32; ++t) acc[t] = 0.f;
#pragma unroll 4
    for (int s = 0; s < a.slices; ++s) {
      const float* p = a.g_coef_part + ((size_t)s * n + b) * kCoefK + lane;
#pragma unroll
      for (int t = 0; t < kCoefK / 32; ++t) acc[t] += __ldg(p + 32 * t);
    }
#pragma unroll
    for (int t = 0; t < kCoefK / 32; ++t) s_gc[warp][lane + 32 * t] = acc[t];
    __syncwarp();
  }

  // ---- forward recompute (k_pose_chain)
  float th0 = 0.f, th1 = 0.f, th2 = 0.f;
  if (active) {
    const float* p = a.pose + b * (3 * kJ) + 3 * j;
    th0 = __ldg(p); th1 = __ldg(p + 1); th2 = __ldg(p + 2);
  }
  float Rl[9];
  rodrigues_hmr(th0, th1, th2, Rl);
  const float my_beta = lane < NB ? __ldg(a.betas + b * NB + lane) : 0.f;
  float J0 = __ldg(m.j_template + 3 * j), J1 = __ldg(m.j_template + 3 * j + 1), J2 = __ldg(m.j_template + 3 * j + 2);
  for (int k = 0; k < NB; ++k) {
    const float bk = __shfl_sync(full, my_beta, k);
    const float* js = m.j_shapedirs + k * (3 * kJ) + 3 * j;
    J0 = fmaf(bk, __ldg(js), J0); J1 = fmaf(bk, __ldg(js + 1), J1); J2 = fmaf(bk, __ldg(js + 2), J2);
  }
  const int parent = active ? m.parents[j] : 0;
  const int pj = parent < 0 ? 0 : parent;
  const int depth = active ? m.depth[j] : -1;
  const float pJ0 = __shfl_sync(full, J0, pj), pJ1 = __shfl_sync(full, J1, pj), pJ2 = __shfl_sync(full, J2, pj);
  float d0 = J0, d1 = J1, d2 = J2;
  if (depth > 0) { d0 = J0 - pJ0; d1 = J1 - pJ1; d2 = J2 - pJ2; }
  float G[12];
  {
    float R[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = Rl[e];
    if (depth == 0 && a.rotate_base) { R[1] = -R[1]; R[2] = -R[2]; R[4] = -R[4]; R[5] = -R[5]; R[7] = -R[7]; R[8] = -R[8]; }
#pragma unroll
    for (int r = 0; r < 3; ++r) { G[4 * r] = R[3 * r]; G[4 * r + 1] = R[3 * r + 1]; G[4 * r + 2] = R[3 * r + 2]; }
    G[3] = d0; G[7] = d1; G[11] = d2;
    for (int level = 1; level <= m.max_depth; ++level) {
      float P[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) P[e] = __shfl_sync(full, G[e], pj);
      if (depth == level) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float a0 = P[4 * r], a1 = P[4 * r + 1], a2 = P[4 * r + 2], a3 = P[4 * r + 3];
#pragma unroll
          for (int c = 0; c < 3; ++c) G[4 * r + c] = fmaf(a2, R[6 + c], fmaf(a1, R[3 + c], a0 * R[c]));
          G[4 * r + 3] = fmaf(a2, d2, fmaf(a1, d1, a0 * d0)) + a3;
        }
      }
    }
  }
  float Pw[9];   // parent's world rotation
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Pw[3 * r + c] = __shfl_sync(full, G[4 * r + c], pj);
  unsigned childmask = 0;
  for (int jj = 0; jj < kJ; ++jj) {
    const unsigned mm = __ballot_sync(full, active && depth > 0 && parent == jj);
    if (lane == jj) childmask = mm;
  }

  // ---- seeds: projection, joints, skinning
  float gj0 = 0.f, gj1 = 0.f, gj2 = 0.f;
  if (active && a.g_joints && !a.regressed) {
    const float* p = a.g_joints + (b * kJ + j) * 3;
    gj0 = __ldg(p); gj1 = __ldg(p + 1); gj2 = __ldg(p + 2);
  }
  if (a.g_kp2d && a.cam) {
    const float s = __ldg(a.cam + b * 3), tx = __ldg(a.cam + b * 3 + 1), ty = __ldg(a.cam + b * 3 + 2);
    float gk0 = 0.f, gk1 = 0.f, jx = 0.f, jy = 0.f;
    if (active) {
      gk0 = __ldg(a.g_kp2d + (b * kJ + j) * 2); gk1 = __ldg(a.g_kp2d + (b * kJ + j) * 2 + 1);
      if (a.regressed) { jx = __ldg(a.joints_fwd + (b * kJ + j) * 3); jy = __ldg(a.joints_fwd + (b * kJ + j) * 3 + 1); }
      else { jx = G[3]; jy = G[7]; }
    }
    float c0 = active ? fmaf(gk1, jy + ty, gk0 * (jx + tx)) : 0.f, c1 = gk0, c2 = gk1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      c0 += __shfl_xor_sync(full, c0, off); c1 += __shfl_xor_sync(full, c1, off); c2 += __shfl_xor_sync(full, c2, off);
    }
    if (lane == 0 && a.g_cam) { a.g_cam[b * 3] = c0; a.g_cam[b * 3 + 1] = s * c1; a.g_cam[b * 3 + 2] = s * c2; }
    if (!a.regressed) { gj0 = fmaf(s, gk0, gj0); gj1 = fmaf(s, gk1, gj1); }
  } else if (lane < 3 && a.g_cam) {
    a.g_cam[b * 3 + lane] = 0.f;
  }
  float gA[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) gA[e] = 0.f;
  if (active && a.g_A) {
    const float4* p = reinterpret_cast<const float4*>(a.g_A + (b * kJ + j) * 12);
    const float4 r0 = __ldg(p), r1 = __ldg(p + 1), r2 = __ldg(p + 2);
    gA[0] = r0.x; gA[1] = r0.y; gA[2] = r0.z; gA[3] = r0.w;
    gA[4] = r1.x; gA[5] = r1.y; gA[6] = r1.z; gA[7] = r1.w;
    gA[8] = r2.x; gA[9] = r2.y; gA[10] = r2.z; gA[11] = r2.w;
  }
  // A_j = [Rw_j | tw_j - Rw_j Jr_j],  joints_j = tw_j
  float gtw[3] = {gj0 + gA[3], gj1 + gA[7], gj2 + gA[11]};
  float gRw[9], gJr[3], gR[9];
  const float Jr[3] = {J0, J1, J2};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) gRw[3 * r + c] = fmaf(-gA[4 * r + 3], Jr[c], gA[4 * r + c]);
#pragma unroll
  for (int c = 0; c < 3; ++c) gJr[c] = -(G[c] * gA[3] + G[4 + c] * gA[7] + G[8 + c] * gA[11]);
#pragma unroll
  for (int e = 0; e < 9; ++e) gR[e] = 0.f;
  if (active && j >= 1 && a.g_coef_part) {
#pragma unroll
    for (int e = 0; e < 9; ++e) gR[e] = s_gc[warp][NB + 9 * (j - 1) + e];
  }

  // ---- chain, leaves to root
  const float dv[3] = {d0, d1, d2};
  float* sc = &s_c[warp][0][0];
  for (int level = m.max_depth; level >= 1; --level) {
    if (depth == level) {
      float* o = sc + j * 16;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c)    // gR += Pw^T gRw
          gR[3 * i + c] += Pw[i] * gRw[c] + Pw[3 + i] * gRw[3 + c] + Pw[6 + i] * gRw[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < 3; ++i)    // to parent: gRw Rl^T + gtw (x) d
          o[3 * r + i] = fmaf(gtw[r], dv[i], gRw[3 * r] * Rl[3 * i] + gRw[3 * r + 1] * Rl[3 * i + 1] + gRw[3 * r + 2] * Rl[3 * i + 2]);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float gd = Pw[i] * gtw[0] + Pw[3 + i] * gtw[1] + Pw[6 + i] * gtw[2];   // Pw^T gtw
        gJr[i] += gd;
        o[12 + i] = -gd;
        o[9 + i] = gtw[i];
      }
    }
    __syncwarp();
    if (active && depth == level - 1) {
      unsigned cm = childmask;
      while (cm) {
        const int c = __ffs(cm) - 1;
        cm &= cm - 1;
        const float* o = sc + c * 16;
#pragma unroll
        for (int e = 0; e < 9; ++e) gRw[e] += o[e];
#pragma unroll
        for (int e = 0; e < 3; ++e) { gtw[e] += o[9 + e]; gJr[e] += o[12 + e]; }
      }
    }
    __syncwarp();
  }
  if (depth == 0) {   // root: Rw_0 = R_0 (* diag(1,-1,-1)),  tw_0 = Jr_0
    const float f = a.rotate_base ? -1.f : 1.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { gR[3 * r] += gRw[3 * r]; gR[3 * r + 1] += f * gRw[3 * r + 1]; gR[3 * r + 2] += f * gRw[3 * r + 2]; }
#pragma unroll
    for (int e = 0; e < 3; ++e) gJr[e] += gtw[e];
  }

  // ---- folded regressor: g_betas[k] = g_coef[k] + sum_j gJr_j . J_shapedirs[k,j]
  float mine = 0.f;
  for (int k = 0; k < NB; ++k) {
    float v = 0.f;
    if (active) {
      const float* js = m.j_shapedirs + k * (3 * kJ) + 3 * j;
      v = fmaf(gJr[2], __ldg(js + 2), fmaf(gJr[1], __ldg(js + 1), gJr[0] * __ldg(js)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(full, v, off);
    if (lane == k) mine = v;
  }
  if (lane < NB) {
    if (a.g_coef_part) mine += s_gc[warp][lane];
    a.g_betas[b * NB + lane] = mine;
  }
  // ---- Rodrigues
  if (active) {
    float gt[3];
    rodrigues_bwd(th0, th1, th2, gR, gt);
    float* o = a.g_pose + b * (3 * kJ) + 3 * j;
    o[0] = gt[0]; o[1] = gt[1]; o[2] = gt[2];
  }
}

}  // namespace smplb200
