// k3 (+k4) CUDA-core path: linear blend skinning, thread = vertex, bodies looped.
//
//   verts[b,v,:] = ( sum_j w[v,j] * A[b,j] ) * [vposed[b,v,:]; 1]        (3x4 form)
//
// The per-vertex weights stay in registers for the whole CTA lifetime (ELL: <=4 (joint, weight)
// pairs when the model is SMPL-sparse -- skipped terms are exact zeros; or all
// 24 for arbitrary models).  Each body's 24 joint transforms (1152 B) are staged in shared
// memory.  For the ELL path they are stored TRANSPOSED, [entry e][joint j]: the lanes of a warp
// gather by their own joint index, and with j as the fastest index distinct joints are distinct
// banks and equal joints broadcast -- 12 conflict-free LDS.32 per slot.  (Row-major [j][12] read
// with LDS.128 cost ~10 wavefronts per load for random joints: 120 vs 48 per warp and body.) vposed is planar so the three coordinate loads are fully
// coalesced; the xyz-interleaved output goes through a per-warp shared-memory transpose so each
// warp emits three fully coalesced 128-byte stores covering 384 contiguous bytes (a body row is
// only 8-byte aligned, 6890*3*4 = 82,680 B, so 16-byte vectors or TMA stores do not apply).
//
// k4: the weak-perspective projection kp2d = s * (joints_xy + t) rides in the
// epilogue of the CTAs that own vertex tile 0.
#pragma once
#include "common.cuh"

namespace smplb200 {

constexpr int kLbsThreads = kVertTile;  // 128
constexpr int kLbsStage = 8;            // bodies staged per shared-memory refill

template <bool DENSE>
__global__ void __launch_bounds__(kLbsThreads)
k_lbs_fma(DeviceModel m, const float* __restrict__ vposed, const float* __restrict__ A,
          long long n, int bodies_per_cta, float* __restrict__ verts,
          const float* __restrict__ joints_in, const float* __restrict__ cam,
          float* __restrict__ kp2d) {
  __shared__ __align__(16) float s_A[kLbsStage][kJ * 12];   // DENSE: [j][12]; ELL: [e][24] (transposed)
  __shared__ __align__(16) float s_out[kLbsThreads / 32][96];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v = blockIdx.x * kVertTile + tid;
  const long long b_begin = (long long)blockIdx.y * bodies_per_cta;
  const long long b_end = min(n, b_begin + bodies_per_cta);
  const int V = m.V, VP = m.VP;
  const int warp_v0 = blockIdx.x * kVertTile + warp * 32;
  const int warp_valid = max(0, min(32, V - warp_v0));  // vertices of this warp that exist

  float w[DENSE ? kJ : 4];
  uint32_t jj = 0;
  if (DENSE) {
#pragma unroll
    for (int j = 0; j < kJ; ++j) w[j] = m.dense_w[(size_t)v * kJ + j];
  } else {
    const float4 w4 = m.ell_w[v];
    w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
    jj = m.ell_j[v];
  }

  for (long long bs = b_begin; bs < b_end; bs += kLbsStage) {
    const int nb = (int)min((long long)kLbsStage, b_end - bs);
    __syncthreads();
    if (DENSE) {
      const float4* src = reinterpret_cast<const float4*>(A + bs * (kJ * 12));
      float4* dst = reinterpret_cast<float4*>(&s_A[0][0]);
      for (int i = tid; i < nb * (kJ * 3); i += kLbsThreads) dst[i] = __ldg(src + i);
    } else {
      const float* src = A + bs * (kJ * 12);
      for (int i = tid; i < nb * (kJ * 12); i += kLbsThreads) {
        const int bi = i / (kJ * 12), r = i - bi * (kJ * 12), j = r / 12, e = r - 12 * j;
        s_A[bi][e * kJ + j] = __ldg(src + i);
      }
    }
    __syncthreads();
    for (int bi = 0; bi < nb; ++bi) {
      const long long b = bs + bi;
      const float* vp = vposed + (size_t)b * 3 * VP + v;
      const float x = __ldg(vp), y = __ldg(vp + VP), z = __ldg(vp + 2 * VP);
      float T[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] = 0.f;
      if (DENSE) {
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
          const float4* a = reinterpret_cast<const float4*>(&s_A[bi][j * 12]);
          const float4 r0 = a[0], r1 = a[1], r2 = a[2];
          T[0] = fmaf(w[j], r0.x, T[0]); T[1] = fmaf(w[j], r0.y, T[1]);
          T[2] = fmaf(w[j], r0.z, T[2]); T[3] = fmaf(w[j], r0.w, T[3]);
          T[4] = fmaf(w[j], r1.x, T[4]); T[5] = fmaf(w[j], r1.y, T[5]);
          T[6] = fmaf(w[j], r1.z, T[6]); T[7] = fmaf(w[j], r1.w, T[7]);
          T[8] = fmaf(w[j], r2.x, T[8]); T[9] = fmaf(w[j], r2.y, T[9]);
          T[10] = fmaf(w[j], r2.z, T[10]); T[11] = fmaf(w[j], r2.w, T[11]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float* a = &s_A[bi][(jj >> (8 * s)) & 0xff];
#pragma unroll
          for (int e = 0; e < 12; ++e) T[e] = fmaf(w[s], a[e * kJ], T[e]);
        }
      }
      const float ox = fmaf(T[2], z, fmaf(T[1], y, fmaf(T[0], x, T[3])));
      const float oy = fmaf(T[6], z, fmaf(T[5], y, fmaf(T[4], x, T[7])));
      const float oz = fmaf(T[10], z, fmaf(T[9], y, fmaf(T[8], x, T[11])));
      // per-warp transpose to xyz-interleaved, then coalesced stores
      float* so = s_out[warp];
      __syncwarp();
      so[3 * lane] = ox; so[3 * lane + 1] = oy; so[3 * lane + 2] = oz;
      __syncwarp();
      {
        float* dst = verts + ((size_t)b * V + warp_v0) * 3;
        const int nf = warp_valid * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (lane + 32 * k < nf) dst[lane + 32 * k] = so[lane + 32 * k];
      }
    }
  }

  // k4 epilogue: weak-perspective projection of this CTA's bodies
  if (blockIdx.x == 0 && kp2d != nullptr) {
    for (long long i = b_begin * (kJ * 2) + tid; i < b_end * (kJ * 2); i += kLbsThreads) {
      const long long b = i / (kJ * 2);
      const int r = int(i - b * (kJ * 2)), j = r >> 1, c = r & 1;
      const float s = __ldg(cam + b * 3), t = __ldg(cam + b * 3 + 1 + c);
      kp2d[i] = __fmul_rn(s, __fadd_rn(__ldg(joints_in + (b * kJ + j) * 3 + c), t));
    }
  }
}

// joints regressed from skinned vertices (HMR-style, SURVEY.md A.6 "regressed"): one warp per
// (body, joint), CSR over the joint's non-zero regressor entries, fixed-order shuffle reduction.
__global__ void __launch_bounds__(256)
k_regress_joints(DeviceModel m, const float* __restrict__ verts, long long n,
                 float* __restrict__ joints, const float* __restrict__ cam,
                 float* __restrict__ kp2d) {
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n * kJ) return;
  const long long b = wid / kJ;
  const int j = int(wid - b * kJ);
  const int p0 = m.jreg_ptr[j], p1 = m.jreg_ptr[j + 1];
  const float* vb = verts + (size_t)b * m.V * 3;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int p = p0 + lane; p < p1; p += 32) {
    const int vi = m.jreg_idx[p];
    const float r = m.jreg_val[p];
    s0 = fmaf(r, vb[3 * vi], s0); s1 = fmaf(r, vb[3 * vi + 1], s1); s2 = fmaf(r, vb[3 * vi + 2], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) {
    float* jo = joints + (b * kJ + j) * 3;
    jo[0] = s0; jo[1] = s1; jo[2] = s2;
    if (kp2d != nullptr) {
      const float s = cam[b * 3];
      kp2d[(b * kJ + j) * 2] = __fmul_rn(s, __fadd_rn(s0, cam[b * 3 + 1]));
      kp2d[(b * kJ + j) * 2 + 1] = __fmul_rn(s, __fadd_rn(s1, cam[b * 3 + 2]));
    }
  }
}

}  // namespace smplb200
