// Shared definitions for the SMPL sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smplb200 {

constexpr int kJ = 24;             // joints (fixed by the kintree warp layout: lane j = joint j)
constexpr int kP = 9 * (kJ - 1);   // 207 pose-feature terms
constexpr int kCoefK = 224;        // padded contraction length (betas | pose_feature | 1 | 0...)
constexpr int kMaxBetas = 14;       // NB + 207 pose terms + 3 template rows <= 224
constexpr int kVertTile = 128;     // vertices per tile == TMEM lanes == LBS CTA width

// Device-resident packed model (all pointers are device pointers unless noted).
struct DeviceModel {
  int V, VP, NB, KB;               // verts, verts padded to 128, betas, basis rows (NB + 207 + 1)
  int NC;                          // planar columns = 3 * VP
  int max_nnz;                     // max skinning weights per vertex
  int max_depth;                   // kinematic tree depth
  int jreg_nnz;                    // total non-zeros of the joint regressor
  const float* basis;              // [KB, NC] planar, row NB+207 = v_template
  const float* j_template;         // [J*3]   folded J_regressor^T v_template
  const float* j_shapedirs;        // [NB, J*3] folded J_regressor^T shapedirs
  const int* parents;              // [J]
  const int* depth;                // [J]
  const float4* ell_w;             // [VP] up to 4 weights per vertex (zero padded)
  const uint32_t* ell_j;           // [VP] 4 packed uint8 joint indices
  const float* dense_w;            // [VP, J]
  const int* jreg_ptr;             // [J+1] CSR over joints
  const int* jreg_idx;             // [jreg_nnz] vertex index
  const float* jreg_val;           // [jreg_nnz]
  // tensor-core operands (see k_blend_tc.cuh / k_lbs_tc.cuh)
  const uint32_t* basis_rows_bf16_hi;  // [NC][112] basis^T rows, 2 bf16 per word (TMEM A operand)
  const uint32_t* basis_rows_bf16_lo;  // same, low part of the 2-term bf16 split
  const uint32_t* basis_rows_f16_hi;   // [NC][112] the same rows in fp16 (template split in three fp16 pieces)
  const uint32_t* basis_rows_f16_lo;
  const uint32_t* basis_rows_tf32;     // [NC][224] tf32
  const uint32_t* w_tf32;              // [VP][48] skinning-weight rows, tf32 W_hi(24) | W_lo(24)
  // backward pass (k_backward.cuh)
  const int* wcsr_ptr;                 // [J+1] skinning weights as CSR over JOINTS (transpose of ELL)
  const int* wcsr_idx;                 // [rounds*32] vertex index; lane l of a round has v % 32 == l (bank-conflict free)
  const float* wcsr_val;               // [rounds*32] weight (0 for fillers)
  const float* dense_jreg;             // [VP, J] joint regressor rows (regressed-joint gradient)
  const uint32_t* bwd_basis_tf32_hi;   // [NC/32][8][224][4] K-major tiles of basis[k, col] (k_blend_bwd_tc)
  const uint32_t* bwd_basis_tf32_lo;   // same, low part of the 2-term tf32 split
  const uint16_t* bwd_basis_bf16_hi;   // [NC/32][4][224][8] the same tiles in bf16 (hi | lo)
  const uint16_t* bwd_basis_bf16_lo;
  // fused blendshapes + skinning kernel (k_fused_tc.cuh); null when NB > 13 (shape + template rows must fit one K step)
  const uint8_t* fz_basis;             // [VP/128 tiles][hi x|y|z: 28 chunks x 128 rows x 8 fp16][lo x|y|z: 2 chunks ...]
  const uint32_t* fz_w;                // [VP][32 words] fp16 W_hi(24) 0(8) | W_lo(24) 0(8)
};

}  // namespace smplb200
