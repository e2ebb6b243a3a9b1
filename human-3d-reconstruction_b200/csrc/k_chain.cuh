// k2: folded joint regression + Rodrigues + 24-joint kinematic chain, one warp per body.
//
// Lane j (< 24) owns joint j.  The chain is composed level-synchronously down the kintree: at
// level L every lane fetches its parent's world transform with 12 warp shuffles and the lanes
// whose depth == L compose  G_j = G_parent * [R_j | J_j - J_parent].  Joint regression is folded
// at model-create time into  J = J_template + betas * J_shapedirs  (1,440 flop/body
// instead of 992,160), so this kernel reads 340 B/body and never touches the vertex arrays.
//
// Rodrigues mirrors the eager idiom op for op with explicit round-to-nearest
// mul/add so nvcc cannot contract what the eager layer rounds separately.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace smplb200 {

constexpr int kChainWarps = 8;

struct ChainOut {
  float* coef;            // [n, 224] fp32 or null
  float* A;               // [n, 24, 12] or null
  float* joints;          // [n, 24, 3] or null
  uint16_t* coef_bf16_hi; // tensor-core operand images (null unless a tcgen05 path follows)
  uint16_t* coef_bf16_lo;
  int coef_is_f16;        // the two 16-bit images hold fp16 (SMPLB200_PREC_F16X3) instead of bf16
  uint32_t* coef_tf32;
  uint32_t* a_tf32;       // [n/8 blocks][12 chunks][96 rows][4] tf32 hi|lo image of A (LBS blend)
  uint8_t* fz_coef;       // fused kernel: fp16 coef images per 64-body block (k_fused_tc.cuh), or null
  uint8_t* fz_a;          // fused kernel: fp16 hi|lo images of A per 4-body sub-block, or null
  const float* cam;       // [n, 3] (s, tx, ty) or null
  float* kp2d;            // [n, 24, 2] weak-perspective projection of the KINEMATIC joints (k4), or null
};

__device__ __forceinline__ void rodrigues_hmr(float tx, float ty, float tz, float R[9]) {
  const float eps = 1e-8f;
  float ex = __fadd_rn(tx, eps), ey = __fadd_rn(ty, eps), ez = __fadd_rn(tz, eps);
  float n2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
  float angle = __fsqrt_rn(n2);
  float ax = __fdiv_rn(tx, angle), ay = __fdiv_rn(ty, angle), az = __fdiv_rn(tz, angle);
  float half = __fmul_rn(angle, 0.5f);
  float c = cosf(half), s = sinf(half);
  float qw = c, qx = __fmul_rn(s, ax), qy = __fmul_rn(s, ay), qz = __fmul_rn(s, az);
  float qn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qw, qw), __fmul_rn(qx, qx)),
                                            __fmul_rn(qy, qy)), __fmul_rn(qz, qz)));
  float w = __fdiv_rn(qw, qn), x = __fdiv_rn(qx, qn), y = __fdiv_rn(qy, qn), z = __fdiv_rn(qz, qn);
  float w2 = __fmul_rn(w, w), x2 = __fmul_rn(x, x), y2 = __fmul_rn(y, y), z2 = __fmul_rn(z, z);
  float wx = __fmul_rn(w, x), wy = __fmul_rn(w, y), wz = __fmul_rn(w, z);
  float xy = __fmul_rn(x, y), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z);
  R[0] = __fsub_rn(__fsub_rn(__fadd_rn(w2, x2), y2), z2);
  R[1] = __fsub_rn(__fmul_rn(2.f, xy), __fmul_rn(2.f, wz));
  R[2] = __fadd_rn(__fmul_rn(2.f, wy), __fmul_rn(2.f, xz));
  R[3] = __fadd_rn(__fmul_rn(2.f, wz), __fmul_rn(2.f, xy));
  R[4] = __fsub_rn(__fadd_rn(__fsub_rn(w2, x2), y2), z2);
  R[5] = __fsub_rn(__fmul_rn(2.f, yz), __fmul_rn(2.f, wx));
  R[6] = __fsub_rn(__fmul_rn(2.f, xz), __fmul_rn(2.f, wy));
  R[7] = __fadd_rn(__fmul_rn(2.f, wx), __fmul_rn(2.f, yz));
  R[8] = __fadd_rn(__fsub_rn(__fsub_rn(w2, x2), y2), z2);
}

// Canonical K-major no-swizzle operand image offsets (in elements of the operand type):
// an image is [chunks][rows][E] with E elements per 16-byte chunk; see k_blend_tc.cuh.
__device__ __forceinline__ uint32_t f32_to_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint16_t f32_to_bf16_rn(float x) {
  uint16_t r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float bf16_to_f32(uint16_t h) { return __uint_as_float(uint32_t(h) << 16); }
// two floats -> packed bf16x2 (lo element in the low half), one instruction
__device__ __forceinline__ uint32_t f32x2_to_bf16x2_rn(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

constexpr int kCoefBlock = 128;   // bodies per tensor-core coefficient image (MMA N)
constexpr int kLbsBlock = 8;     // bodies per LBS blend image (MMA N = 12 * 8 = 96)
constexpr int kLbsK = 48;        // blend contraction: [A_hi | A_lo] over 24 joints (tf32 split)

__global__ void __launch_bounds__(kChainWarps * 32)
k_pose_chain(DeviceModel m, const float* __restrict__ betas, const float* __restrict__ pose,
             long long n, ChainOut out, int rotate_base) {
  __shared__ __align__(16) float s_coef[kChainWarps][kCoefK];
  __shared__ __align__(16) float s_A[kChainWarps][kJ * 12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = (long long)blockIdx.x * kChainWarps + warp;
  if (b >= n) return;  // whole warp exits together; no block-level sync below
  const int j = lane < kJ ? lane : 0;
  const bool active = lane < kJ;
  const int NB = m.NB;

  // ---- inputs
  float th0 = 0.f, th1 = 0.f, th2 = 0.f;
  if (active) {
    const float* p = pose + b * (3 * kJ) + 3 * j;
    th0 = __ldg(p); th1 = __ldg(p + 1); th2 = __ldg(p + 2);
  }
  float R[9];
  rodrigues_hmr(th0, th1, th2, R);

  // ---- folded joint regression: Jrest = J_template + betas . J_shapedirs
  // (all loads are issued before the first FMA: the loop is fully unrolled over kMaxBetas so the
  //  ~30 L2 round trips overlap instead of chaining)
  const float* bb = betas + b * NB;
  const float my_beta = lane < NB ? __ldg(bb + lane) : 0.f;
  float js0[kMaxBetas], js1[kMaxBetas], js2[kMaxBetas];
#pragma unroll
  for (int k = 0; k < kMaxBetas; ++k) {
    if (k < NB) {
      const float* js = m.j_shapedirs + k * (3 * kJ) + 3 * j;
      js0[k] = __ldg(js); js1[k] = __ldg(js + 1); js2[k] = __ldg(js + 2);
    } else {
      js0[k] = js1[k] = js2[k] = 0.f;
    }
  }
  float J0 = __ldg(m.j_template + 3 * j), J1 = __ldg(m.j_template + 3 * j + 1), J2 = __ldg(m.j_template + 3 * j + 2);
#pragma unroll
  for (int k = 0; k < kMaxBetas; ++k) {
    const float bk = __shfl_sync(0xffffffffu, my_beta, k);
    J0 = fmaf(bk, js0[k], J0); J1 = fmaf(bk, js1[k], J1); J2 = fmaf(bk, js2[k], J2);
  }

  // ---- blendshape coefficients: betas | (R - I) of joints 1..23 | 1 1 1 | zeros
  float* sc = s_coef[warp];
  for (int k = lane; k < kCoefK; k += 32) sc[k] = 0.f;
  __syncwarp();
  if (lane < NB) sc[lane] = my_beta;
  if (lane < 3) sc[NB + kP + lane] = 1.0f;   // v_template rows (split in 3 exact pieces for tcgen05)
  if (active && j >= 1) {
    float* pf = sc + NB + 9 * (j - 1);
    pf[0] = __fsub_rn(R[0], 1.f); pf[1] = R[1]; pf[2] = R[2];
    pf[3] = R[3]; pf[4] = __fsub_rn(R[4], 1.f); pf[5] = R[5];
    pf[6] = R[6]; pf[7] = R[7]; pf[8] = __fsub_rn(R[8], 1.f);
  }

  // ---- kinematic chain
  const int parent = active ? m.parents[j] : 0;
  const int pj = parent < 0 ? 0 : parent;
  const int depth = active ? m.depth[j] : -1;
  float pJ0 = __shfl_sync(0xffffffffu, J0, pj), pJ1 = __shfl_sync(0xffffffffu, J1, pj),
        pJ2 = __shfl_sync(0xffffffffu, J2, pj);
  float tl0 = J0, tl1 = J1, tl2 = J2;
  if (depth > 0) { tl0 = __fsub_rn(J0, pJ0); tl1 = __fsub_rn(J1, pJ1); tl2 = __fsub_rn(J2, pJ2); }
  float G[12];  // [Rw | t] row-major 3x4
  if (depth == 0 && rotate_base) {  // HMR: root * diag(1,-1,-1)
    R[1] = -R[1]; R[2] = -R[2]; R[4] = -R[4]; R[5] = -R[5]; R[7] = -R[7]; R[8] = -R[8];
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    G[4 * r + 0] = R[3 * r + 0]; G[4 * r + 1] = R[3 * r + 1]; G[4 * r + 2] = R[3 * r + 2];
  }
  G[3] = tl0; G[7] = tl1; G[11] = tl2;
  for (int level = 1; level <= m.max_depth; ++level) {
    float P[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) P[e] = __shfl_sync(0xffffffffu, G[e], pj);
    if (depth == level) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float a0 = P[4 * r], a1 = P[4 * r + 1], a2 = P[4 * r + 2], a3 = P[4 * r + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          G[4 * r + c] = fmaf(a2, R[6 + c], fmaf(a1, R[3 + c], __fmul_rn(a0, R[c])));
        G[4 * r + 3] = __fadd_rn(fmaf(a2, tl2, fmaf(a1, tl1, __fmul_rn(a0, tl0))), a3);
      }
    }
  }

  // ---- outputs
  if (active && out.joints) {
    float* jo = out.joints + (b * kJ + j) * 3;
    jo[0] = G[3]; jo[1] = G[7]; jo[2] = G[11];
  }
  // k4 for kinematic joints: the joints are final HERE, ~10 us into the step, so the
  // projection (same two rounded ops as the skinning-epilogue version) and everything downstream of
  // joints/kp2d -- the multi-GPU exchange -- need not wait for the skinning kernel.
  if (active && out.kp2d) {
    const float sc = __ldg(out.cam + b * 3), tx = __ldg(out.cam + b * 3 + 1), ty = __ldg(out.cam + b * 3 + 2);
    float2* kp = reinterpret_cast<float2*>(out.kp2d + (b * kJ + j) * 2);
    *kp = make_float2(__fmul_rn(sc, __fadd_rn(G[3], tx)), __fmul_rn(sc, __fadd_rn(G[7], ty)));
  }
  float* sa = s_A[warp];
  if (active) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float* a = sa + j * 12 + 4 * r;
      a[0] = G[4 * r]; a[1] = G[4 * r + 1]; a[2] = G[4 * r + 2];
      // rest-pose removal: t - Rw * Jrest
      float rj = fmaf(G[4 * r + 2], J2, fmaf(G[4 * r + 1], J1, __fmul_rn(G[4 * r], J0)));
      a[3] = __fsub_rn(G[4 * r + 3], rj);
    }
  }
  __syncwarp();
  if (out.coef) {
    float* dst = out.coef + b * kCoefK;
    for (int k = lane; k < kCoefK; k += 32) dst[k] = sc[k];
  }
  if (out.A) {
    float4* dst = reinterpret_cast<float4*>(out.A + b * (kJ * 12));
    const float4* src = reinterpret_cast<const float4*>(sa);
    for (int k = lane; k < kJ * 3; k += 32) dst[k] = src[k];
  }
  // tensor-core operand images: row = body within its block, K-chunk-major (see k_blend_tc.cuh).
  // Every store is one whole 16-byte chunk (8 bf16 / 4 tf32 of consecutive k for this body); the
  // eight bodies of a CTA fill adjacent chunks, so L2 sees full 128-byte lines.
  if (out.coef_bf16_hi) {
    const long long blk = b / kCoefBlock; const int row = int(b % kCoefBlock);
    uint4* hi = reinterpret_cast<uint4*>(out.coef_bf16_hi + blk * (long long)(kCoefK * kCoefBlock));
    uint4* lo = out.coef_bf16_lo
        ? reinterpret_cast<uint4*>(out.coef_bf16_lo + blk * (long long)(kCoefK * kCoefBlock)) : nullptr;
    if (lane < kCoefK / 8) {
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v0 = sc[8 * lane + 2 * u], v1 = sc[8 * lane + 2 * u + 1];
        if (out.coef_is_f16) {
          const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          wh[u] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
          wl[u] = (uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v0, __half2float(h0)))) |
                  ((uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v1, __half2float(h1)))) << 16);
        } else {
          const uint16_t h0 = f32_to_bf16_rn(v0), h1 = f32_to_bf16_rn(v1);
          wh[u] = (uint32_t)h0 | ((uint32_t)h1 << 16);
          wl[u] = (uint32_t)f32_to_bf16_rn(__fsub_rn(v0, bf16_to_f32(h0))) |
                  ((uint32_t)f32_to_bf16_rn(__fsub_rn(v1, bf16_to_f32(h1))) << 16);
        }
      }
      const size_t off = (size_t)lane * kCoefBlock + row;    // in 16-byte chunks
      hi[off] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      if (lo) lo[off] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
  if (out.coef_tf32) {
    const long long blk = b / kCoefBlock; const int row = int(b % kCoefBlock);
    uint4* im = reinterpret_cast<uint4*>(out.coef_tf32 + blk * (long long)(kCoefK * kCoefBlock));
    for (int c = lane; c < kCoefK / 4; c += 32)
      im[(size_t)c * kCoefBlock + row] =
          make_uint4(f32_to_tf32_rn(sc[4 * c]), f32_to_tf32_rn(sc[4 * c + 1]),
                     f32_to_tf32_rn(sc[4 * c + 2]), f32_to_tf32_rn(sc[4 * c + 3]));
  }
  if (out.a_tf32) {
    // B operand of the blend MMA: rows n = (body_in_block*12 + e), K = 48 = [A_hi | A_lo];
    // chunk c (0..11) holds joints 4(c%6)..4(c%6)+3 of part c/6 for every row.
    const long long blk = b / kLbsBlock; const int bi = int(b % kLbsBlock);
    constexpr int rows = kLbsBlock * 12;
    uint4* im = reinterpret_cast<uint4*>(out.a_tf32 + blk * (long long)(kLbsK * rows));
    for (int item = lane; item < 12 * 12; item += 32) {
      const int c = item / 12, e = item % 12;
      const int j0 = 4 * (c % 6);
      uint32_t w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v = sa[(j0 + u) * 12 + e];
        const uint32_t hi = f32_to_tf32_rn(v);
        w[u] = c < 6 ? hi : f32_to_tf32_rn(__fsub_rn(v, __uint_as_float(hi)));
      }
      im[(size_t)c * rows + bi * 12 + e] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  // ---- operand images of the fused blendshapes+skinning kernel (fp16; layouts in k_fused_tc.cuh) ----
  if (out.fz_coef) {
    // K order: betas | 1 1 1 (template pieces) | 0.. (16 rows) | pose_feature (207) | 0.  One 16-byte
    // chunk = 8 consecutive k of this body; hi chunks 0..27 (+ lo chunks 0..1 for the 16 shape rows).
    uint8_t* img = out.fz_coef + (size_t)(b / 64) * (2048 + kCoefK * 64 * 2);
    const int row = int(b % 64);
    auto value = [&](int nk) -> float {
      if (nk < NB) return sc[nk];
      if (nk < NB + 3) return 1.0f;
      if (nk < 16) return 0.f;
      if (nk < 16 + kP) return sc[NB + (nk - 16)];
      return 0.f;
    };
    if (lane < kCoefK / 8) {
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v0 = value(8 * lane + 2 * u), v1 = value(8 * lane + 2 * u + 1);
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        wh[u] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        wl[u] = (uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v0, __half2float(h0)))) |
                ((uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v1, __half2float(h1)))) << 16);
      }
      reinterpret_cast<uint4*>(img + 2048)[(size_t)lane * 64 + row] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      if (lane < 2) reinterpret_cast<uint4*>(img)[(size_t)lane * 64 + row] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
  if (out.fz_a) {
    // per 4-body sub-block: [chunk 0..2: A_hi joints 8c..8c+7][chunk 3..5: A_lo][chunk 6: zeros], rows = (body, entry)
    uint4* img = reinterpret_cast<uint4*>(out.fz_a + (size_t)(b / 4) * (7 * 48 * 16));
    const int r0 = int(b % 4) * 12;
    for (int item = lane; item < 3 * 12; item += 32) {
      const int c = item / 12, e = item % 12;
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float v0 = sa[(8 * c + 2 * u) * 12 + e], v1 = sa[(8 * c + 2 * u + 1) * 12 + e];
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        wh[u] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        wl[u] = (uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v0, __half2float(h0)))) |
                ((uint32_t)__half_as_ushort(__float2half_rn(__fsub_rn(v1, __half2float(h1)))) << 16);
      }
      img[(size_t)c * 48 + r0 + e] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      img[(size_t)(3 + c) * 48 + r0 + e] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
    if (lane < 12) img[(size_t)6 * 48 + r0 + lane] = make_uint4(0u, 0u, 0u, 0u);
  }
}

}  // namespace smplb200
