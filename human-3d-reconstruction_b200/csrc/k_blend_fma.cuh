// k1 (small-batch path): shape + pose blendshapes as a register-tiled fp32 FMA contraction.
//
//   vposed[b, col] = ( sum_{k<207} pose_feature[b,k] * posedirs[k,col] )
//                  + ( ( sum_{k<NB} betas[b,k] * shapedirs[k,col] ) + v_template[col] )
//
// with the same association as the eager layer (shape sum, + template, then + pose sum;
// SURVEY.md A.2/A.5) and every sum taken k-ascending into a single fp32 accumulator
//  so results are reproducible and independent of the batch split.
//
// Thread tile: 4 adjacent planar columns x BB bodies.  Per k: one float4 of the basis
// (coalesced, read once per CTA) and BB/4 broadcast LDS.128 of coefficients feed 4*BB FMAs, the
// 16 FMA : 1 LDS.128 ratio that keeps the FMA pipe, not the shared-memory crossbar, the limiter.
#pragma once
#include "common.cuh"

namespace smplb200 {

constexpr int kFmaThreads = 64;                       // 64 threads x 4 columns = 256 columns / CTA
constexpr int kFmaColsPerCta = kFmaThreads * 4;

template <int BB>
__global__ void __launch_bounds__(kFmaThreads)
k_blend_fma(DeviceModel m, const float* __restrict__ coef, long long n, float* __restrict__ vposed) {
  // coefficients of this CTA's BB bodies, transposed to [k][body] for broadcast LDS.128
  __shared__ __align__(16) float s_c[kCoefK][BB + 4];  // +4: transpose-store conflicts 32-way -> 4-way
  const long long b0 = (long long)blockIdx.y * BB;
  const int nb = (int)min((long long)BB, n - b0);
  for (int idx = threadIdx.x; idx < kCoefK * BB; idx += kFmaThreads) {
    const int bi = idx / kCoefK, k = idx % kCoefK;  // coalesced over k
    s_c[k][bi] = bi < nb ? __ldg(coef + (b0 + bi) * kCoefK + k) : 0.f;
  }
  __syncthreads();
  const int col = blockIdx.x * kFmaColsPerCta + threadIdx.x * 4;
  if (col >= m.NC) return;
  const int NB = m.NB;
  const float* bp = m.basis + col;
  const size_t ld = (size_t)m.NC;

  float acc[BB][4];
#pragma unroll
  for (int i = 0; i < BB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  // pose blend: rows NB .. NB+206
#pragma unroll 2
  for (int k = 0; k < kP; ++k) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + (size_t)(NB + k) * ld));
#pragma unroll
    for (int i4 = 0; i4 < BB / 4; ++i4) {
      const float4 c = *reinterpret_cast<const float4*>(&s_c[NB + k][4 * i4]);
      const float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[4 * i4 + u][0] = fmaf(cc[u], bv.x, acc[4 * i4 + u][0]);
        acc[4 * i4 + u][1] = fmaf(cc[u], bv.y, acc[4 * i4 + u][1]);
        acc[4 * i4 + u][2] = fmaf(cc[u], bv.z, acc[4 * i4 + u][2]);
        acc[4 * i4 + u][3] = fmaf(cc[u], bv.w, acc[4 * i4 + u][3]);
      }
    }
  }
  // shape blend + template, then the final add, body by body
  const float4 vt = __ldg(reinterpret_cast<const float4*>(bp + (size_t)(NB + kP) * ld));
  const int plane = col / m.VP;           // a float4 never straddles planes (VP % 128 == 0)
  const int v = col - plane * m.VP;
#pragma unroll
  for (int i = 0; i < BB; ++i) {
    if (i >= nb) break;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int k = 0; k < NB; ++k) {
      const float4 sv = __ldg(reinterpret_cast<const float4*>(bp + (size_t)k * ld));
      const float c = s_c[k][i];
      s0 = fmaf(c, sv.x, s0); s1 = fmaf(c, sv.y, s1); s2 = fmaf(c, sv.z, s2); s3 = fmaf(c, sv.w, s3);
    }
    float4 o;
    o.x = __fadd_rn(acc[i][0], __fadd_rn(s0, vt.x));
    o.y = __fadd_rn(acc[i][1], __fadd_rn(s1, vt.y));
    o.z = __fadd_rn(acc[i][2], __fadd_rn(s2, vt.z));
    o.w = __fadd_rn(acc[i][3], __fadd_rn(s3, vt.w));
    *reinterpret_cast<float4*>(vposed + ((b0 + i) * 3 + plane) * (size_t)m.VP + v) = o;
  }
}

}  // namespace smplb200
