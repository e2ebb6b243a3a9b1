// k1 (small-batch path): shape + pose blendshapes as a register-tiled fp32 FMA contraction.
//
//   vposed[b, col] = ( sum_{k<207} pose_feature[b,k] * posedirs[k,col] )
//                  + ( ( sum_{k<NB} betas[b,k] * shapedirs[k,col] ) + v_template[col] )
//
// with the same association as the eager layer (shape sum, + template, then + pose sum;
// SURVEY.md A.2/A.5).  The 207 pose terms are split into KS contiguous K slices, one per group of
// 64 threads; every slice is summed k-ascending into a single fp32 accumulator and the slices are
// combined in slice order through shared memory -- a FIXED order, so results are reproducible and
// independent of how the batch is split.  The K split exists because at N <= 64
// the op is latency-bound: one serial 207-step loop per thread with ~4 warps per SM left the FMA
// pipes idle (round-1 launch list: 66 us cold at N = 1); KS = 4 gives 4x the warps and the loads
// are software-pipelined 4 deep.  Every batch size uses the same KS, so the summation order -- and
// therefore every output bit -- is independent of the batch size and of how a batch is sharded.
//
// Thread tile: 4 adjacent planar columns x BB bodies.  Per k: one float4 of the basis (coalesced)
// and BB/4 broadcast LDS.128 of coefficients feed 4*BB FMAs -- the 16 FMA : 1 LDS.128 ratio that
// keeps the FMA pipe, not the shared-memory crossbar, the limiter.
#pragma once
#include "common.cuh"

namespace smplb200 {

constexpr int kFmaColThreads = 64;                    // 64 threads x 4 columns = 256 columns / CTA
constexpr int kFmaColsPerCta = kFmaColThreads * 4;

template <int BB, int KS>
struct BlendFmaCfg {
  static constexpr int kGroups = KS + 1;              // KS pose K-slices + one shape/template group
  static constexpr int kThreads = kFmaColThreads * kGroups;
  static constexpr int kCoefStride = BB + 4;          // +4: transpose-store conflicts 32-way -> 4-way
  static constexpr size_t kCoefBytes = (size_t)kCoefK * kCoefStride * sizeof(float);
  static constexpr size_t kRedBytes = (size_t)KS * BB * 4 * kFmaColThreads * sizeof(float);
  static constexpr size_t kSmemBytes = kCoefBytes + kRedBytes;
};

template <int BB, int KS>
__global__ void __launch_bounds__(kFmaColThreads * (KS + 1), 2)
k_blend_fma(DeviceModel m, const float* __restrict__ coef, long long n, float* __restrict__ vposed) {
  using C = BlendFmaCfg<BB, KS>;
  extern __shared__ __align__(16) float smem_f[];
  float (*s_c)[C::kCoefStride] = reinterpret_cast<float (*)[C::kCoefStride]>(smem_f);  // [k][body]
  float* s_red = smem_f + kCoefK * C::kCoefStride;                                     // [g-1][i][t]
  const int t = threadIdx.x % kFmaColThreads, g = threadIdx.x / kFmaColThreads;
  const long long b0 = (long long)blockIdx.y * BB;
  const int nb = (int)min((long long)BB, n - b0);
  for (int idx = threadIdx.x; idx < kCoefK * BB; idx += C::kThreads) {
    const int bi = idx / kCoefK, k = idx % kCoefK;  // coalesced over k
    s_c[k][bi] = bi < nb ? __ldg(coef + (b0 + bi) * kCoefK + k) : 0.f;
  }
  __syncthreads();
  const int col = blockIdx.x * kFmaColsPerCta + t * 4;
  const bool in_range = col < m.NC;
  const int NB = m.NB;
  const float* bp = m.basis + (in_range ? col : 0);
  const size_t ld = (size_t)m.NC;

  float acc[BB][4];
#pragma unroll
  for (int i = 0; i < BB; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  // group g < KS: pose rows NB + [k0, k1);  group KS: the NB shape rows (then + template)
  const int k0 = g < KS ? NB + g * kP / KS : 0;
  const int k1 = g < KS ? NB + (g + 1) * kP / KS : NB;
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bp + (size_t)k * ld));
#pragma unroll
    for (int i4 = 0; i4 < BB / 4; ++i4) {
      const float4 c = *reinterpret_cast<const float4*>(&s_c[k][4 * i4]);
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
  if (g == KS) {   // v_shaped = shape sum + template
    const float4 vt = __ldg(reinterpret_cast<const float4*>(bp + (size_t)(NB + kP) * ld));
#pragma unroll
    for (int i = 0; i < BB; ++i) {
      acc[i][0] = __fadd_rn(acc[i][0], vt.x); acc[i][1] = __fadd_rn(acc[i][1], vt.y);
      acc[i][2] = __fadd_rn(acc[i][2], vt.z); acc[i][3] = __fadd_rn(acc[i][3], vt.w);
    }
  }
  if (g > 0) {
    float* r = s_red + (size_t)(g - 1) * (BB * 4 * kFmaColThreads) + t;
#pragma unroll
    for (int i = 0; i < BB; ++i)
#pragma unroll
      for (int u = 0; u < 4; ++u) r[(i * 4 + u) * kFmaColThreads] = acc[i][u];
  }
  __syncthreads();
  if (g > 0 || !in_range) return;
  // pose slices in slice order, then v_posed = pose sum + v_shaped  (fixed association)
#pragma unroll
  for (int s = 1; s <= KS; ++s) {
    const float* r = s_red + (size_t)(s - 1) * (BB * 4 * kFmaColThreads) + t;
#pragma unroll
    for (int i = 0; i < BB; ++i)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[i][u] = __fadd_rn(acc[i][u], r[(i * 4 + u) * kFmaColThreads]);
  }
  const int plane = col / m.VP;           // a float4 never straddles planes (VP % 128 == 0)
  const int v = col - plane * m.VP;
#pragma unroll
  for (int i = 0; i < BB; ++i) {
    if (i >= nb) break;
    *reinterpret_cast<float4*>(vposed + ((b0 + i) * 3 + plane) * (size_t)m.VP + v) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

template <int BB, int KS>
inline cudaError_t blend_fma_set_smem() {
  return cudaFuncSetAttribute(k_blend_fma<BB, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)BlendFmaCfg<BB, KS>::kSmemBytes);
}

template <int BB, int KS>
inline void blend_fma_launch(const DeviceModel& m, const float* coef, long long n, float* vposed,
                             cudaStream_t s) {
  using C = BlendFmaCfg<BB, KS>;
  const unsigned gx = (unsigned)((m.NC + kFmaColsPerCta - 1) / kFmaColsPerCta);
  k_blend_fma<BB, KS><<<dim3(gx, (unsigned)((n + BB - 1) / BB)), C::kThreads, C::kSmemBytes, s>>>(m, coef, n, vposed);
}

}  // namespace smplb200
