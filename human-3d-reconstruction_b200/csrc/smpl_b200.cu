// libsmpl_b200.so -- C ABI (include/smpl_b200.h) over the hand-written sm_100a SMPL kernels.
//
// Host side: model packing at create time, workspace carving, kernel selection and launch.
// No per-call allocation, no host synchronisation, no mutable globals (only a thread_local
// "last CUDA error" cell), no CPU compute fallback.
#include "../../include/smpl_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <utility>
#include <vector>

#include "common.cuh"
#include "k_chain.cuh"
#include "k_blend_fma.cuh"
#include "k_lbs_fma.cuh"
#include "k_blend_tc.cuh"
#include "k_lbs_tc.cuh"
#include "k_fused_tc.cuh"
#include "k_decode.cuh"
#include "k_backward.cuh"
#include "k_blend_bwd_tc.cuh"
#include "k_dcn.cuh"
#include "k_exchange.cuh"

using namespace smplb200;

struct SmplB200Model {
  DeviceModel d;
  int device = 0;
  int num_sms = 0;
  int chunk = 0;        // bodies per k1->k3 pass (tensor-core paths); 0 = whole batch in one pass
  bool lbs_bwd_staged = false;   // backward skinning kernel stages a body's g_v + vposed in smem
  void* blob = nullptr;
  size_t blob_bytes = 0;
};

namespace {

thread_local int tl_last_cuda_error = 0;

constexpr int kDefaultChunk = 0;
constexpr long long kMaxGridYBodies16 = 65535LL * 16;   // FMA kernels: 16 bodies per grid.y slot

inline int cuda_fail(cudaError_t e) {
  tl_last_cuda_error = (int)e;
  return SMPLB200_ERR_CUDA;
}
#define CU_TRY(expr)                                   \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) return cuda_fail(_e);       \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline bool prec_16bit(uint32_t p) { return p == SMPLB200_PREC_BF16 || p == SMPLB200_PREC_BF16X3 || p == SMPLB200_PREC_F16X3; }
inline bool prec_split16(uint32_t p) { return p == SMPLB200_PREC_BF16X3 || p == SMPLB200_PREC_F16X3; }

// RAII device switch: the library always runs on the model's device and restores the caller's.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// ---- host-side bf16 / tf32 rounding (round-to-nearest-even), used only at model create ----
inline uint16_t host_bf16(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float host_bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
// float -> IEEE half, round to nearest even (subnormals and overflow handled); and back
inline uint16_t host_f16(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  const uint32_t absu = u & 0x7fffffffu;
  if (absu >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((absu > 0x7f800000u) ? 0x200u : 0u));
  if (absu >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);              // rounds to >= 65520: infinity
  if (absu < 0x33000001u) return (uint16_t)sign;                           // < 2^-25 (or == 2^-25: ties to even 0)
  const int exp = (int)(absu >> 23) - 127;
  uint32_t mant = (absu & 0x7fffffu) | 0x800000u;                          // 24-bit significand
  int shift = exp >= -14 ? 13 : (13 + (-14 - exp));                        // bits dropped
  const uint32_t half_bit = 1u << (shift - 1);
  const uint32_t rem = mant & ((1u << shift) - 1u);
  uint32_t q = mant >> shift;
  if (rem > half_bit || (rem == half_bit && (q & 1u))) ++q;
  uint32_t h;
  if (exp >= -14) h = ((uint32_t)(exp + 15) << 10) + (q - 0x400u);         // q in [0x400, 0x800]; carry bumps the exponent
  else h = q;                                                              // subnormal (q may reach 0x400 = smallest normal)
  return (uint16_t)(sign | h);
}
inline float host_f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  float mag;
  if (e == 0) mag = std::ldexp((float)m, -24);
  else if (e == 31) mag = m ? NAN : INFINITY;
  else mag = std::ldexp((float)(m | 0x400u), (int)e - 25);
  uint32_t u;
  std::memcpy(&u, &mag, 4);
  u |= sign;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline uint32_t host_tf32(float x) {  // keep 10 mantissa bits, low 13 bits zero
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return u & 0xffffe000u;
  u += 0xfffu + ((u >> 13) & 1u);
  return u & 0xffffe000u;
}
inline float bits_to_f32(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

struct BlobBuilder {
  std::vector<uint8_t> bytes;
  size_t add(const void* src, size_t n) {
    size_t off = align_up(bytes.size(), 256);
    bytes.resize(off + n);
    if (src) std::memcpy(bytes.data() + off, src, n);
    return off;
  }
};

// ---- workspace carving -----------------------------------------------------------------
struct Workspace {
  size_t coef = 0, A = 0, vposed = 0, joints = 0;
  size_t coef_hi = 0, coef_lo = 0, coef_tf32 = 0, a_tf32 = 0;
  size_t fz_coef = 0, fz_a = 0;      // fused kernel operand images
  size_t total = 0;
};

struct Plan {
  uint32_t prec;      // resolved SMPLB200_PREC_*
  uint32_t lbs;       // resolved SMPLB200_LBS_*
  bool regressed;
  bool rotate_base;
};

bool resolve_plan(const SmplB200Model* m, long long n, uint32_t flags, Plan* p) {
  uint32_t prec = flags & SMPLB200_PREC_MASK;
  if (prec > SMPLB200_PREC_F16X3) return false;
  if (prec == SMPLB200_PREC_F16 && !m->d.fz_basis) return false;      // model has too many betas for the fused kernel
  if (prec == SMPLB200_PREC_AUTO)
    prec = n >= SMPLB200_TC_MIN_BATCH ? SMPLB200_PREC_F16X3 : SMPLB200_PREC_FP32;
  uint32_t lbs = flags & SMPLB200_LBS_MASK;
  if (lbs == SMPLB200_LBS_AUTO)
    lbs = n >= SMPLB200_TC_LBS_MIN_BATCH ? SMPLB200_LBS_TC : SMPLB200_LBS_FMA;
  if (lbs == SMPLB200_LBS_FMA && m->d.max_nnz > 4) lbs = SMPLB200_LBS_DENSE;
  if (prec == SMPLB200_PREC_F16) {      // one fused kernel does blendshapes AND skinning on tcgen05
    if ((flags & SMPLB200_LBS_MASK) != SMPLB200_LBS_AUTO && (flags & SMPLB200_LBS_MASK) != SMPLB200_LBS_TC) return false;
    lbs = SMPLB200_LBS_TC;
  }
  if (flags & ~(SMPLB200_PREC_MASK | SMPLB200_JOINTS_REGRESSED | SMPLB200_ROTATE_BASE |
                SMPLB200_LBS_MASK))
    return false;
  p->prec = prec;
  p->lbs = lbs;
  p->regressed = (flags & SMPLB200_JOINTS_REGRESSED) != 0;
  p->rotate_base = (flags & SMPLB200_ROTATE_BASE) != 0;
  return true;
}

Workspace carve(const SmplB200Model* m, long long n, const Plan& p) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nn = (size_t)std::max<long long>(n, 1);
  if (p.prec == SMPLB200_PREC_F16) {     // fused: no coef / A / vposed / tf32 images, only the two fp16 operand images
    w.joints = take(nn * kJ * 3 * sizeof(float));
    w.fz_coef = take(fz_coef_image_bytes(n));
    w.fz_a = take(fz_a_image_bytes(n));
    w.total = off;
    return w;
  }
  w.coef = take(nn * kCoefK * sizeof(float));
  w.A = take(nn * kJ * 12 * sizeof(float));
  const bool chunked = m->chunk > 0 && p.prec != SMPLB200_PREC_FP32 && p.lbs == SMPLB200_LBS_TC &&
                       nn > (size_t)m->chunk;
  w.vposed = take((chunked ? (size_t)m->chunk : nn) * 3 * (size_t)m->d.VP * sizeof(float));
  w.joints = take(nn * kJ * 3 * sizeof(float));
  const size_t coef_blocks = (nn + kCoefBlock - 1) / kCoefBlock;
  const size_t lbs_blocks = (nn + kLbsBlock - 1) / kLbsBlock;
  if (prec_16bit(p.prec)) w.coef_hi = take(coef_blocks * kCoefBlock * kCoefK * 2);
  if (prec_split16(p.prec)) w.coef_lo = take(coef_blocks * kCoefBlock * kCoefK * 2);
  if (p.prec == SMPLB200_PREC_TF32) w.coef_tf32 = take(coef_blocks * kCoefBlock * kCoefK * 4);
  if (p.lbs == SMPLB200_LBS_TC) w.a_tf32 = take(lbs_blocks * kLbsBlock * 12 * kLbsK * 4);
  w.total = off;
  return w;
}

// ---- launches ----------------------------------------------------------------------------
int launch_chain(const SmplB200Model* m, const float* betas, const float* pose, long long n,
                 const ChainOut& out, bool rotate_base, cudaStream_t s) {
  if (n == 0) return SMPLB200_OK;
  const unsigned grid = (unsigned)((n + kChainWarps - 1) / kChainWarps);
  k_pose_chain<<<grid, kChainWarps * 32, 0, s>>>(m->d, betas, pose, n, out, rotate_base ? 1 : 0);
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

int launch_blend_fma(const SmplB200Model* m, const float* coef, long long n, float* vposed,
                     cudaStream_t s) {
  if (n == 0) return SMPLB200_OK;
  // same 4-way K split for every batch size (bitwise shard invariance); body tile 8 or 16
  if (n <= 8) {
    blend_fma_launch<8, 4>(m->d, coef, n, vposed, s);
  } else {
    // bodies ride on grid.y (<= 65535 groups of 16): very large batches go in several launches
    for (long long b0 = 0; b0 < n; b0 += kMaxGridYBodies16) {
      const long long nb = std::min<long long>(kMaxGridYBodies16, n - b0);
      blend_fma_launch<16, 4>(m->d, coef + (size_t)b0 * kCoefK, nb, vposed + (size_t)b0 * 3 * m->d.VP, s);
    }
  }
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

int launch_lbs_fma(const SmplB200Model* m, bool dense, const float* vposed, const float* A,
                   long long n, float* verts, const float* joints_in, const float* cam,
                   float* kp2d, cudaStream_t s) {
  if (n == 0) return SMPLB200_OK;
  const int bodies_per_cta = 16;
  // bodies ride on grid.y (<= 65535 groups of 16): very large batches go in several launches
  for (long long b0 = 0; b0 < n; b0 += kMaxGridYBodies16) {
    const long long nb = std::min<long long>(kMaxGridYBodies16, n - b0);
    dim3 grid((unsigned)(m->d.VP / kVertTile), (unsigned)((nb + bodies_per_cta - 1) / bodies_per_cta));
    const float* vp = vposed + (size_t)b0 * 3 * m->d.VP;
    const float* Ab = A + (size_t)b0 * kJ * 12;
    float* vo = verts + (size_t)b0 * m->d.V * 3;
    const float* ji = joints_in ? joints_in + (size_t)b0 * kJ * 3 : nullptr;
    const float* cm = cam ? cam + (size_t)b0 * 3 : nullptr;
    float* kp = kp2d ? kp2d + (size_t)b0 * kJ * 2 : nullptr;
    if (dense)
      k_lbs_fma<true><<<grid, kLbsThreads, 0, s>>>(m->d, vp, Ab, nb, bodies_per_cta, vo, ji, cm, kp);
    else
      k_lbs_fma<false><<<grid, kLbsThreads, 0, s>>>(m->d, vp, Ab, nb, bodies_per_cta, vo, ji, cm, kp);
  }
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

int launch_regress(const SmplB200Model* m, const float* verts, long long n, float* joints,
                   const float* cam, float* kp2d, cudaStream_t s) {
  if (n == 0) return SMPLB200_OK;
  const long long warps = n * kJ;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  k_regress_joints<<<grid, 256, 0, s>>>(m->d, verts, n, joints, cam, kp2d);
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// opt-in dynamic shared memory for the tcgen05 kernels (per device, idempotent)
cudaError_t configure_tc_kernels() {
  cudaError_t e;
  if ((e = blend_fma_set_smem<8, 4>()) != cudaSuccess) return e;
  if ((e = blend_fma_set_smem<16, 4>()) != cudaSuccess) return e;
  if ((e = blend_tc_set_smem<SMPLB200_PREC_BF16>()) != cudaSuccess) return e;
  if ((e = blend_tc_set_smem<SMPLB200_PREC_BF16X3>()) != cudaSuccess) return e;
  if ((e = blend_tc_set_smem<SMPLB200_PREC_TF32>()) != cudaSuccess) return e;
  if ((e = blend_tc_set_smem<SMPLB200_PREC_F16X3>()) != cudaSuccess) return e;
  if ((e = blend_bwd_tc_set_smem<kBwdTf32>()) != cudaSuccess) return e;
  if ((e = blend_bwd_tc_set_smem<kBwdTf32x3>()) != cudaSuccess) return e;
  if ((e = blend_bwd_tc_set_smem<kBwdBf16x3>()) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_blend_bwd_fma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)kBbSmemBytes)) != cudaSuccess) return e;
  if ((e = fused_tc_set_smem()) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_lbs_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)kLbsSmemBytes);
}

cudaError_t launch_blend_tc_any(const SmplB200Model* m, uint32_t prec, const uint16_t* chi,
                                const uint16_t* clo, const uint32_t* ctf, long long n, float* vposed,
                                cudaStream_t s) {
  return launch_blend_tc(m->d, m->num_sms, prec, chi, clo, ctf, n, vposed, s);
}

size_t coef_image_bytes(long long n, uint32_t prec) {
  const size_t blocks = ((size_t)std::max<long long>(n, 1) + kCoefBlock - 1) / kCoefBlock;
  const size_t one = blocks * kCoefBlock * kCoefK;
  if (prec == SMPLB200_PREC_BF16) return align_up(one * 2, 256);
  if (prec_split16(prec)) return 2 * align_up(one * 2, 256);
  if (prec == SMPLB200_PREC_TF32) return align_up(one * 4, 256);
  return 0;
}
size_t a_image_bytes(long long n) {
  const size_t blocks = ((size_t)std::max<long long>(n, 1) + kLbsBlock - 1) / kLbsBlock;
  return align_up(blocks * kLbsBlock * 12 * kLbsK * 4, 256);
}

}  // namespace

// =============================================================================================
extern "C" {

int smplb200_version(void) { return SMPLB200_VERSION; }
int smplb200_last_cuda_error(void) { return tl_last_cuda_error; }

const char* smplb200_strerror(int status) {
  switch (status) {
    case SMPLB200_OK: return "ok";
    case SMPLB200_ERR_INVALID_ARG: return "invalid argument";
    case SMPLB200_ERR_UNSUPPORTED: return "unsupported model shape or flag combination";
    case SMPLB200_ERR_WORKSPACE: return "workspace missing, too small or misaligned";
    case SMPLB200_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
    case SMPLB200_ERR_CUDA: return "CUDA runtime or launch error (see smplb200_last_cuda_error)";
    case SMPLB200_ERR_NO_DEVICE: return "no usable CUDA device (need compute capability 10.x)";
    case SMPLB200_ERR_ALLOC: return "allocation failed";
    default: return "unknown status";
  }
}

int smplb200_model_create(const SmplB200ModelDesc* desc, SmplB200Model** out_model) {
  if (!desc || !out_model) return SMPLB200_ERR_INVALID_ARG;
  *out_model = nullptr;
  if (desc->struct_size != sizeof(SmplB200ModelDesc)) return SMPLB200_ERR_INVALID_ARG;
  if (!desc->v_template || !desc->shapedirs || !desc->posedirs || !desc->j_regressor ||
      !desc->weights || !desc->parents)
    return SMPLB200_ERR_INVALID_ARG;
  if (desc->num_joints != kJ) return SMPLB200_ERR_UNSUPPORTED;
  if (desc->num_betas < 1 || desc->num_betas > kMaxBetas) return SMPLB200_ERR_UNSUPPORTED;
  if (desc->num_verts < 1 || desc->num_verts > (1 << 24)) return SMPLB200_ERR_UNSUPPORTED;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return SMPLB200_ERR_NO_DEVICE;
  }
  if (desc->device < 0 || desc->device >= ndev) return SMPLB200_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10) return SMPLB200_ERR_NO_DEVICE;  // sm_100a cubin only

  const int V = desc->num_verts, NB = desc->num_betas;
  const int VP = (int)align_up((size_t)V, kVertTile), NC = 3 * VP, KB = NB + kP + 1;

  // ---- kinematic tree
  std::vector<int> parents(kJ), depth(kJ, 0);
  int max_depth = 0;
  for (int j = 0; j < kJ; ++j) {
    int p = desc->parents[j];
    if (j == 0) {
      parents[j] = -1;  // root, whatever sentinel the file used (-1 or 0xFFFFFFFF)
    } else {
      if (p < 0 || p >= j) return SMPLB200_ERR_UNSUPPORTED;  // parents must precede children
      parents[j] = p;
      depth[j] = depth[p] + 1;
      max_depth = std::max(max_depth, depth[j]);
    }
  }

  BlobBuilder bb;
  try {
    // ---- planar fp32 basis [KB, NC]: shapedirs | posedirs | v_template
    std::vector<float> basis((size_t)KB * NC, 0.f);
    auto src_row = [&](int k) -> const float* {
      if (k < NB) return desc->shapedirs + (size_t)k * 3 * V;
      if (k < NB + kP) return desc->posedirs + (size_t)(k - NB) * 3 * V;
      return desc->v_template;
    };
    for (int k = 0; k < KB; ++k) {
      const float* s = src_row(k);
      float* d = basis.data() + (size_t)k * NC;
      for (int v = 0; v < V; ++v)
        for (int c = 0; c < 3; ++c) d[c * VP + v] = s[3 * v + c];
    }
    // ---- folded joint regressor (fp64 accumulation on the host)
    std::vector<float> jt(kJ * 3), jsd((size_t)NB * kJ * 3);
    {
      std::vector<double> acc((size_t)(NB + 1) * kJ * 3, 0.0);
      for (int v = 0; v < V; ++v) {
        const float* r = desc->j_regressor + (size_t)v * kJ;
        for (int j = 0; j < kJ; ++j) {
          const double rj = r[j];
          if (rj == 0.0) continue;
          for (int c = 0; c < 3; ++c) {
            acc[(size_t)NB * kJ * 3 + j * 3 + c] += rj * desc->v_template[3 * v + c];
            for (int k = 0; k < NB; ++k)
              acc[(size_t)k * kJ * 3 + j * 3 + c] += rj * desc->shapedirs[(size_t)k * 3 * V + 3 * v + c];
          }
        }
      }
      for (int i = 0; i < kJ * 3; ++i) jt[i] = (float)acc[(size_t)NB * kJ * 3 + i];
      for (size_t i = 0; i < jsd.size(); ++i) jsd[i] = (float)acc[i];
    }
    // ---- skinning weights: ELL (<=4) + dense, padded to VP rows
    std::vector<float> dense_w((size_t)VP * kJ, 0.f);
    std::vector<float> ell_w((size_t)VP * 4, 0.f);
    std::vector<uint32_t> ell_j((size_t)VP, 0u);
    int max_nnz = 0;
    for (int v = 0; v < V; ++v) {
      int cnt = 0;
      uint32_t packed = 0;
      for (int j = 0; j < kJ; ++j) {
        const float wv = desc->weights[(size_t)v * kJ + j];
        dense_w[(size_t)v * kJ + j] = wv;
        if (wv != 0.f) {
          if (cnt < 4) {
            ell_w[(size_t)v * 4 + cnt] = wv;
            packed |= (uint32_t)j << (8 * cnt);
          }
          ++cnt;
        }
      }
      ell_j[v] = packed;
      max_nnz = std::max(max_nnz, cnt);
    }
    // ---- joint regressor CSR over joints (for SMPLB200_JOINTS_REGRESSED)
    std::vector<int> jptr(kJ + 1, 0), jidx;
    std::vector<float> jval;
    for (int j = 0; j < kJ; ++j) {
      for (int v = 0; v < V; ++v) {
        const float r = desc->j_regressor[(size_t)v * kJ + j];
        if (r != 0.f) { jidx.push_back(v); jval.push_back(r); }
      }
      jptr[j + 1] = (int)jidx.size();
    }
    if (jidx.empty()) { jidx.push_back(0); jval.push_back(0.f); }
    // ---- backward pass: skinning weights as CSR over joints, dense joint-regressor rows
    // Entries of a joint are laid out in ROUNDS of 32 (one per lane of the warp that reduces the
    // joint); within a round lane l holds a vertex with v % 32 == l (or a zero-weight filler), so
    // the 32 shared-memory gathers of a round hit 32 distinct banks (bank = v % 32 for the vposed
    // planes, 3v + r % 32 for g_v).  Unordered lists cost ~3.5 wavefronts per gather (round-1 ncu).
    std::vector<int> wptr(kJ + 1, 0), widx;
    std::vector<float> wval;
    for (int j = 0; j < kJ; ++j) {
      std::vector<std::vector<std::pair<int, float>>> bucket(32);
      for (int v = 0; v < V; ++v) {
        const float wv = desc->weights[(size_t)v * kJ + j];
        if (wv != 0.f) bucket[v & 31].push_back({v, wv});
      }
      size_t rounds = 0;
      for (auto& bk : bucket) rounds = std::max(rounds, bk.size());
      for (size_t r = 0; r < rounds; ++r)
        for (int l = 0; l < 32; ++l) {
          if (r < bucket[l].size()) { widx.push_back(bucket[l][r].first); wval.push_back(bucket[l][r].second); }
          else { widx.push_back(std::min(l, V - 1)); wval.push_back(0.f); }
        }
      wptr[j + 1] = (int)widx.size();
    }
    if (widx.empty()) { widx.push_back(0); wval.push_back(0.f); }
    // basis[k, col] as the B operand of the tensor-core blendshape backward: per K-step of 32
    // planar columns one tile [8 chunks][224 rows k][4 cols], tf32 hi | lo (rows >= NB+207 are 0)
    std::vector<uint32_t> gbh((size_t)NC * kCoefK, 0), gbl((size_t)NC * kCoefK, 0);
    std::vector<uint16_t> gbbh((size_t)NC * kCoefK, 0), gbbl((size_t)NC * kCoefK, 0);   // bf16: [ks][4][224][8]
    for (int ks = 0; ks < NC / 32; ++ks)
      for (int c = 0; c < 4; ++c)
        for (int r = 0; r < NB + kP; ++r)
          for (int e = 0; e < 8; ++e) {
            const float x = basis[(size_t)r * NC + ks * 32 + c * 8 + e];
            const uint16_t hi = host_bf16(x);
            const size_t idx = (((size_t)ks * 4 + c) * kCoefK + r) * 8 + e;
            gbbh[idx] = hi;
            gbbl[idx] = host_bf16(x - host_bf16_to_f32(hi));
          }
    for (int ks = 0; ks < NC / 32; ++ks)
      for (int c = 0; c < 8; ++c)
        for (int r = 0; r < NB + kP; ++r)
          for (int e = 0; e < 4; ++e) {
            const float x = basis[(size_t)r * NC + ks * 32 + c * 4 + e];
            const uint32_t hi = host_tf32(x);
            const size_t idx = (((size_t)ks * 8 + c) * kCoefK + r) * 4 + e;
            gbh[idx] = hi;
            gbl[idx] = host_tf32(x - bits_to_f32(hi));
          }
    std::vector<float> dense_jreg((size_t)VP * kJ, 0.f);
    for (int v = 0; v < V; ++v)
      for (int j = 0; j < kJ; ++j) dense_jreg[(size_t)v * kJ + j] = desc->j_regressor[(size_t)v * kJ + j];

    // ---- tensor-core A operand of k1: basis^T rows [planar column][K], resident in TMEM
    std::vector<uint16_t> bhi((size_t)NC * kCoefK, 0), blo((size_t)NC * kCoefK, 0);
    std::vector<uint32_t> btf((size_t)NC * kCoefK, 0);
    // The v_template row is O(1 m) while blendshape terms are O(1 mm): it is split EXACTLY over
    // three of the spare K rows (coefficient 1.0 each): bf16 hi+mid+lo / tf32 hi+lo(+rest).
    auto tmpl_piece = [&](float x, int piece, bool tf) -> float {
      float rem = x;
      for (int p = 0; p <= piece; ++p) {
        const float h = tf ? bits_to_f32(host_tf32(rem)) : host_bf16_to_f32(host_bf16(rem));
        if (p == piece) return h;
        rem -= h;
      }
      return 0.f;
    };
    for (int col = 0; col < NC; ++col)
      for (int k = 0; k < kCoefK; ++k) {
        const int tp = k - (NB + kP);  // 0,1,2 -> template pieces
        const float raw = k < NB + kP ? basis[(size_t)k * NC + col] : 0.f;
        const float tv = (tp >= 0 && tp < 3) ? basis[(size_t)(NB + kP) * NC + col] : 0.f;
        const float x = (tp >= 0 && tp < 3) ? tmpl_piece(tv, tp, false) : raw;
        const float xt = (tp >= 0 && tp < 3) ? tmpl_piece(tv, tp, true) : raw;
        const uint16_t h = host_bf16(x);
        bhi[(size_t)col * kCoefK + k] = h;
        blo[(size_t)col * kCoefK + k] = host_bf16(x - host_bf16_to_f32(h));
        btf[(size_t)col * kCoefK + k] = host_tf32(xt);
      }
    // the same rows in fp16 (SMPLB200_PREC_F16X3): hi + residual, template as three fp16 pieces
    std::vector<uint16_t> fhi((size_t)NC * kCoefK, 0), flo((size_t)NC * kCoefK, 0);
    for (int col = 0; col < NC; ++col)
      for (int k = 0; k < kCoefK; ++k) {
        const int tp = k - (NB + kP);
        float x = k < NB + kP ? basis[(size_t)k * NC + col] : 0.f;
        if (tp >= 0 && tp < 3) {
          float rem = basis[(size_t)(NB + kP) * NC + col];
          for (int q = 0; q <= tp; ++q) { x = host_f16_to_f32(host_f16(rem)); rem -= x; }
        }
        const uint16_t h = host_f16(x);
        fhi[(size_t)col * kCoefK + k] = h;
        flo[(size_t)col * kCoefK + k] = host_f16(x - host_f16_to_f32(h));
      }
    // skinning weights as the TMEM A operand of the LBS blend: rows [VP][48] = W_hi(24) | W_lo(24)
    std::vector<uint32_t> wtf((size_t)VP * kLbsK, 0);
    for (int v = 0; v < VP; ++v)
      for (int j = 0; j < kJ; ++j) {
        const float x = dense_w[(size_t)v * kJ + j];
        const uint32_t hi = host_tf32(x);
        const uint32_t lo = host_tf32(x - bits_to_f32(hi));
        wtf[(size_t)v * kLbsK + j] = hi;
        wtf[(size_t)v * kLbsK + 24 + j] = lo;
      }

    // ---- fused blendshapes+skinning kernel (k_fused_tc.cuh): per vertex tile the exact shared-memory image
    // of the fp16 basis, K order = betas | template pieces (hi, mid, lo; coefficient 1) | 0 | pose rows | 0
    std::vector<uint16_t> fzb;
    std::vector<uint32_t> fzw;
    const bool fz_ok = NB + 3 <= kFzShapeK;
    if (fz_ok) {
      const int ntile = VP / 128;
      fzb.assign((size_t)ntile * (kFzBasisBytes / 2), 0);
      auto tmpl16 = [&](float x, int piece) -> float {
        float rem = x;
        for (int p = 0; p <= piece; ++p) {
          const float h = host_f16_to_f32(host_f16(rem));
          if (p == piece) return h;
          rem -= h;
        }
        return 0.f;
      };
      for (int t = 0; t < ntile; ++t) {
        uint16_t* tile = fzb.data() + (size_t)t * (kFzBasisBytes / 2);
        for (int pl = 0; pl < 3; ++pl) {
          uint16_t* hi = tile + (size_t)pl * (kFzPlaneHi / 2);
          uint16_t* lo = tile + (size_t)3 * (kFzPlaneHi / 2) + (size_t)pl * (kFzPlaneLo / 2);
          for (int r = 0; r < 128; ++r) {
            const size_t col = (size_t)pl * VP + (size_t)t * 128 + r;
            for (int nk = 0; nk < kCoefK; ++nk) {
              float x = 0.f, xl = 0.f;
              if (nk < NB) {
                x = basis[(size_t)nk * NC + col];
                const float h = host_f16_to_f32(host_f16(x));
                xl = x - h;
              } else if (nk < NB + 3) {
                x = tmpl16(basis[(size_t)(NB + kP) * NC + col], nk - NB);
              } else if (nk >= kFzShapeK && nk < kFzShapeK + kP) {
                x = basis[(size_t)(NB + nk - kFzShapeK) * NC + col];
              }
              hi[(size_t)(nk >> 3) * (128 * 8) + r * 8 + (nk & 7)] = host_f16(x);
              if (nk < kFzShapeK) lo[(size_t)(nk >> 3) * (128 * 8) + r * 8 + (nk & 7)] = host_f16(xl);
            }
          }
        }
      }
      fzw.assign((size_t)VP * 32, 0);
      for (int v = 0; v < VP; ++v)
        for (int j = 0; j < kJ; ++j) {
          const float x = dense_w[(size_t)v * kJ + j];
          const uint16_t h = host_f16(x);
          const uint16_t l = host_f16(x - host_f16_to_f32(h));
          uint16_t* row = reinterpret_cast<uint16_t*>(fzw.data() + (size_t)v * 32);
          row[j] = h;
          row[32 + j] = l;
        }
    }

    std::unique_ptr<SmplB200Model> mp(new (std::nothrow) SmplB200Model());   // freed on every early return / throw
    if (!mp) return SMPLB200_ERR_ALLOC;
    SmplB200Model* m = mp.get();
    const size_t o_basis = bb.add(basis.data(), basis.size() * 4);
    const size_t o_jt = bb.add(jt.data(), jt.size() * 4);
    const size_t o_jsd = bb.add(jsd.data(), jsd.size() * 4);
    const size_t o_par = bb.add(parents.data(), kJ * 4);
    const size_t o_dep = bb.add(depth.data(), kJ * 4);
    const size_t o_ew = bb.add(ell_w.data(), ell_w.size() * 4);
    const size_t o_ej = bb.add(ell_j.data(), ell_j.size() * 4);
    const size_t o_dw = bb.add(dense_w.data(), dense_w.size() * 4);
    const size_t o_jp = bb.add(jptr.data(), jptr.size() * 4);
    const size_t o_ji = bb.add(jidx.data(), jidx.size() * 4);
    const size_t o_jv = bb.add(jval.data(), jval.size() * 4);
    const size_t o_bhi = bb.add(bhi.data(), bhi.size() * 2);
    const size_t o_blo = bb.add(blo.data(), blo.size() * 2);
    const size_t o_btf = bb.add(btf.data(), btf.size() * 4);
    const size_t o_fhi = bb.add(fhi.data(), fhi.size() * 2);
    const size_t o_flo = bb.add(flo.data(), flo.size() * 2);
    const size_t o_wtf = bb.add(wtf.data(), wtf.size() * 4);
    const size_t o_wp = bb.add(wptr.data(), wptr.size() * 4);
    const size_t o_wi = bb.add(widx.data(), widx.size() * 4);
    const size_t o_wv = bb.add(wval.data(), wval.size() * 4);
    const size_t o_djr = bb.add(dense_jreg.data(), dense_jreg.size() * 4);
    const size_t o_gbh = bb.add(gbh.data(), gbh.size() * 4);
    const size_t o_gbl = bb.add(gbl.data(), gbl.size() * 4);
    const size_t o_gbbh = bb.add(gbbh.data(), gbbh.size() * 2);
    const size_t o_gbbl = bb.add(gbbl.data(), gbbl.size() * 2);
    const size_t o_fzb = fz_ok ? bb.add(fzb.data(), fzb.size() * 2) : 0;
    const size_t o_fzw = fz_ok ? bb.add(fzw.data(), fzw.size() * 4) : 0;

    DeviceGuard guard(desc->device);
    if (guard.err != cudaSuccess) return cuda_fail(guard.err);
    cudaError_t e = cudaMalloc(&m->blob, bb.bytes.size());
    if (e != cudaSuccess) { m->blob = nullptr; tl_last_cuda_error = (int)e; cudaGetLastError(); return SMPLB200_ERR_ALLOC; }
    e = cudaMemcpy(m->blob, bb.bytes.data(), bb.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(m->blob); return cuda_fail(e); }
    // opt-in shared memory sizes for the tensor-core kernels
    e = configure_tc_kernels();
    if (e != cudaSuccess) { cudaFree(m->blob); return cuda_fail(e); }

    {  // backward skinning kernel: shared-memory staging when one body's g_v + vposed fit.  The
       // attribute belongs to the FUNCTION on the device, not to this model: always opt in to the
       // device maximum so a second, smaller model can never lower the limit under a larger one.
      const size_t need = lbs_bwd_smem_bytes(V, VP);
      cudaFuncAttributes fa;
      e = cudaFuncGetAttributes(&fa, k_lbs_bwd_split);
      if (e != cudaSuccess) { cudaFree(m->blob); return cuda_fail(e); }
      const size_t cap = (size_t)prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
      if (need <= cap) {
        e = cudaFuncSetAttribute(k_lbs_bwd_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
        if (e != cudaSuccess) { cudaFree(m->blob); return cuda_fail(e); }
        m->lbs_bwd_staged = true;
      }
    }
    m->blob_bytes = bb.bytes.size();
    m->device = desc->device;
    m->num_sms = prop.multiProcessorCount;
    uint8_t* base = static_cast<uint8_t*>(m->blob);
    DeviceModel& d = m->d;
    d.V = V; d.VP = VP; d.NB = NB; d.KB = KB; d.NC = NC;
    d.max_nnz = max_nnz; d.max_depth = max_depth; d.jreg_nnz = (int)jval.size();
    {  // bodies per k1->k3 pass: keeps the vposed intermediate L2-resident (multiple of 128)
      const char* t = std::getenv("SMPLB200_CHUNK");
      int c = t ? std::atoi(t) : kDefaultChunk;
      m->chunk = c <= 0 ? 0 : (c + kCoefBlock - 1) / kCoefBlock * kCoefBlock;
    }
    d.basis = reinterpret_cast<const float*>(base + o_basis);
    d.j_template = reinterpret_cast<const float*>(base + o_jt);
    d.j_shapedirs = reinterpret_cast<const float*>(base + o_jsd);
    d.parents = reinterpret_cast<const int*>(base + o_par);
    d.depth = reinterpret_cast<const int*>(base + o_dep);
    d.ell_w = reinterpret_cast<const float4*>(base + o_ew);
    d.ell_j = reinterpret_cast<const uint32_t*>(base + o_ej);
    d.dense_w = reinterpret_cast<const float*>(base + o_dw);
    d.jreg_ptr = reinterpret_cast<const int*>(base + o_jp);
    d.jreg_idx = reinterpret_cast<const int*>(base + o_ji);
    d.jreg_val = reinterpret_cast<const float*>(base + o_jv);
    d.basis_rows_bf16_hi = reinterpret_cast<const uint32_t*>(base + o_bhi);
    d.basis_rows_bf16_lo = reinterpret_cast<const uint32_t*>(base + o_blo);
    d.basis_rows_tf32 = reinterpret_cast<const uint32_t*>(base + o_btf);
    d.basis_rows_f16_hi = reinterpret_cast<const uint32_t*>(base + o_fhi);
    d.basis_rows_f16_lo = reinterpret_cast<const uint32_t*>(base + o_flo);
    d.w_tf32 = reinterpret_cast<const uint32_t*>(base + o_wtf);
    d.wcsr_ptr = reinterpret_cast<const int*>(base + o_wp);
    d.wcsr_idx = reinterpret_cast<const int*>(base + o_wi);
    d.wcsr_val = reinterpret_cast<const float*>(base + o_wv);
    d.dense_jreg = reinterpret_cast<const float*>(base + o_djr);
    d.bwd_basis_tf32_hi = reinterpret_cast<const uint32_t*>(base + o_gbh);
    d.bwd_basis_tf32_lo = reinterpret_cast<const uint32_t*>(base + o_gbl);
    d.bwd_basis_bf16_hi = reinterpret_cast<const uint16_t*>(base + o_gbbh);
    d.bwd_basis_bf16_lo = reinterpret_cast<const uint16_t*>(base + o_gbbl);
    d.fz_basis = fz_ok ? base + o_fzb : nullptr;
    d.fz_w = fz_ok ? reinterpret_cast<const uint32_t*>(base + o_fzw) : nullptr;
    *out_model = mp.release();
    return SMPLB200_OK;
  } catch (const std::bad_alloc&) {
    return SMPLB200_ERR_ALLOC;
  }
}

void smplb200_model_destroy(SmplB200Model* model) {
  if (!model) return;
  {
    DeviceGuard guard(model->device);
    if (model->blob) cudaFree(model->blob);
  }
  delete model;
}

int32_t smplb200_model_num_verts(const SmplB200Model* m) { return m ? m->d.V : 0; }
int32_t smplb200_model_num_joints(const SmplB200Model* m) { return m ? kJ : 0; }
int32_t smplb200_model_num_betas(const SmplB200Model* m) { return m ? m->d.NB : 0; }
int32_t smplb200_model_device(const SmplB200Model* m) { return m ? m->device : -1; }
int32_t smplb200_model_max_weight_nnz(const SmplB200Model* m) { return m ? m->d.max_nnz : 0; }
size_t smplb200_model_device_bytes(const SmplB200Model* m) { return m ? m->blob_bytes : 0; }
int64_t smplb200_padded_verts(const SmplB200Model* m) { return m ? m->d.VP : 0; }

size_t smplb200_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return 0;
  return carve(model, n, p).total;
}

int smplb200_forward_launch_count(const SmplB200Model* model, int64_t n, uint32_t flags,
                                  int with_projection) {
  Plan p;
  if (!model || n <= 0 || !resolve_plan(model, n, flags, &p)) return 0;
  (void)with_projection;
  if (p.prec == SMPLB200_PREC_F16)       // k2, fused (one launch per 8192 bodies) [, regression]
    return 1 + (int)((n + kFzMaxBodiesPerLaunch - 1) / kFzMaxBodiesPerLaunch) + (p.regressed ? 1 : 0);
  const bool chunked = model->chunk > 0 && p.prec != SMPLB200_PREC_FP32 && p.lbs == SMPLB200_LBS_TC &&
                       n > model->chunk;
  const int passes = chunked ? (int)((n + model->chunk - 1) / model->chunk) : 1;
  return 1 + 2 * passes + (p.regressed ? 1 : 0);
}

int smplb200_pose_chain(const SmplB200Model* model, const float* betas, const float* pose,
                        int64_t n, float* coef, float* A, float* joints, uint32_t flags,
                        void* stream) {
  if (!model || n < 0 || (n > 0 && (!betas || !pose))) return SMPLB200_ERR_INVALID_ARG;
  if (A && !aligned16(A)) return SMPLB200_ERR_ALIGNMENT;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  ChainOut out{};
  out.coef = coef; out.A = A; out.joints = joints;
  return launch_chain(model, betas, pose, n, out, (flags & SMPLB200_ROTATE_BASE) != 0,
                      static_cast<cudaStream_t>(stream));
}

size_t smplb200_blendshapes_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return 0;
  return coef_image_bytes(n, p.prec);
}

int smplb200_blendshapes(const SmplB200Model* model, const float* coef, int64_t n, float* vposed,
                         void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!model || n < 0 || (n > 0 && (!coef || !vposed))) return SMPLB200_ERR_INVALID_ARG;
  if (!aligned16(vposed) || !aligned16(coef)) return SMPLB200_ERR_ALIGNMENT;
  Plan p;
  if (!resolve_plan(model, n, flags, &p)) return SMPLB200_ERR_INVALID_ARG;
  if (n == 0) return SMPLB200_OK;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.prec == SMPLB200_PREC_FP32) return launch_blend_fma(model, coef, n, vposed, s);
  // stand-alone tensor-core entry: operand images are built from the fp32 coefficients first
  const size_t need = coef_image_bytes(n, p.prec);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return SMPLB200_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint16_t* hi = nullptr; uint16_t* lo = nullptr; uint32_t* tf = nullptr;
  if (p.prec == SMPLB200_PREC_TF32) tf = reinterpret_cast<uint32_t*>(ws);
  else hi = reinterpret_cast<uint16_t*>(ws);
  if (prec_split16(p.prec)) lo = reinterpret_cast<uint16_t*>(ws + need / 2);
  const long long total = n * kCoefK;
  k_pack_coef<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(coef, n, hi, lo, tf, p.prec == SMPLB200_PREC_F16X3 ? 1 : 0);
  CU_TRY(cudaGetLastError());
  CU_TRY(launch_blend_tc_any(model, p.prec, hi, lo, tf, n, vposed, s));
  return SMPLB200_OK;
}

size_t smplb200_lbs_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return 0;
  return p.lbs == SMPLB200_LBS_TC ? a_image_bytes(n) : 0;
}

int smplb200_lbs(const SmplB200Model* model, const float* vposed, const float* A, int64_t n,
                 float* vertices, const float* joints_in, const float* cam, float* kp2d,
                 void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!model || n < 0 || (n > 0 && (!vposed || !A || !vertices))) return SMPLB200_ERR_INVALID_ARG;
  if ((kp2d != nullptr) && (!cam || !joints_in)) return SMPLB200_ERR_INVALID_ARG;
  if (!aligned16(A) || !aligned16(vposed)) return SMPLB200_ERR_ALIGNMENT;
  Plan p;
  if (!resolve_plan(model, n, flags, &p)) return SMPLB200_ERR_INVALID_ARG;
  if (n == 0) return SMPLB200_OK;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (p.lbs == SMPLB200_LBS_TC) {
    const size_t need = a_image_bytes(n);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u))
      return SMPLB200_ERR_WORKSPACE;
    uint32_t* img = static_cast<uint32_t*>(workspace);
    const long long total = n * (kJ * 12);
    k_pack_a<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(A, n, img);
    CU_TRY(cudaGetLastError());
    CU_TRY(launch_lbs_tc(model->d, model->num_sms, vposed, img, n, vertices, joints_in, cam, kp2d, s));
    return SMPLB200_OK;
  }
  return launch_lbs_fma(model, p.lbs == SMPLB200_LBS_DENSE, vposed, A, n, vertices, joints_in, cam,
                        kp2d, s);
}

size_t smplb200_blend_skin_workspace_bytes(const SmplB200Model* model, int64_t n) {
  if (!model || n < 0 || !model->d.fz_basis) return 0;
  return align_up(fz_coef_image_bytes(n), 256) + align_up(fz_a_image_bytes(n), 256);
}

int smplb200_blend_skin(const SmplB200Model* model, const float* coef, const float* A, int64_t n,
                        float* vertices, void* workspace, size_t workspace_bytes, void* stream) {
  const bool repack = coef != nullptr || A != nullptr;     // both NULL: reuse the images a previous call left in `workspace`
  if (!model || n < 0 || (n > 0 && !vertices) || (repack && (!coef || !A))) return SMPLB200_ERR_INVALID_ARG;
  if (!model->d.fz_basis) return SMPLB200_ERR_UNSUPPORTED;
  if (n == 0) return SMPLB200_OK;
  const size_t need = smplb200_blend_skin_workspace_bytes(model, n);
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return SMPLB200_ERR_WORKSPACE;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ci = static_cast<uint8_t*>(workspace);
  uint8_t* ai = ci + align_up(fz_coef_image_bytes(n), 256);
  if (repack) {
    const long long total = n * (long long)(kCoefK + kJ * 12);
    k_pack_fz<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(coef, A, n, model->d.NB, ci, ai);
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(launch_fused_tc(model->d, model->num_sms, ci, ai, n, vertices, s));
  return SMPLB200_OK;
}

int smplb200_regress_joints(const SmplB200Model* model, const float* vertices, int64_t n,
                            float* joints, const float* cam, float* kp2d, void* stream) {
  if (!model || n < 0 || (n > 0 && (!vertices || !joints))) return SMPLB200_ERR_INVALID_ARG;
  if (kp2d && !cam) return SMPLB200_ERR_INVALID_ARG;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  return launch_regress(model, vertices, n, joints, cam, kp2d, static_cast<cudaStream_t>(stream));
}

int smplb200_forward(const SmplB200Model* model, const float* betas, const float* pose,
                     const float* cam, int64_t n, float* vertices, float* joints, float* kp2d,
                     void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  return smplb200_forward_opts(model, betas, pose, cam, n, vertices, joints, kp2d, workspace, workspace_bytes,
                               flags, stream, nullptr);
}

int smplb200_forward_opts(const SmplB200Model* model, const float* betas, const float* pose,
                          const float* cam, int64_t n, float* vertices, float* joints, float* kp2d,
                          void* workspace, size_t workspace_bytes, uint32_t flags, void* stream,
                          const SmplB200ForwardOpts* opts) {
  if (!model || n < 0) return SMPLB200_ERR_INVALID_ARG;
  if (opts && opts->struct_size != sizeof(SmplB200ForwardOpts)) return SMPLB200_ERR_INVALID_ARG;
  cudaEvent_t ev_joints = opts ? static_cast<cudaEvent_t>(opts->joints_ready_event) : nullptr;
  if (n == 0) {
    if (ev_joints) {
      DeviceGuard g0(model->device);
      if (g0.err != cudaSuccess) return cuda_fail(g0.err);
      CU_TRY(cudaEventRecord(ev_joints, static_cast<cudaStream_t>(stream)));
    }
    return SMPLB200_OK;
  }
  if (!betas || !pose || !vertices) return SMPLB200_ERR_INVALID_ARG;
  if ((kp2d != nullptr) != (cam != nullptr)) return SMPLB200_ERR_INVALID_ARG;
  Plan p;
  if (!resolve_plan(model, n, flags, &p)) return SMPLB200_ERR_INVALID_ARG;
  const Workspace w = carve(model, n, p);
  if (!workspace || workspace_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return SMPLB200_ERR_WORKSPACE;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* coef = reinterpret_cast<float*>(ws + w.coef);
  float* A = reinterpret_cast<float*>(ws + w.A);
  float* vposed = reinterpret_cast<float*>(ws + w.vposed);
  float* jbuf = joints ? joints : reinterpret_cast<float*>(ws + w.joints);

  // k2
  ChainOut out{};
  out.A = A;
  out.joints = p.regressed ? nullptr : jbuf;
  if (p.prec == SMPLB200_PREC_F16) {
    // fused path: k2 (operand images, joints, kp2d) -> ONE kernel for blendshapes + skinning
    out.A = nullptr;
    out.fz_coef = ws + w.fz_coef;
    out.fz_a = ws + w.fz_a;
    if (kp2d != nullptr && !p.regressed) { out.cam = cam; out.kp2d = kp2d; }
    int st = launch_chain(model, betas, pose, n, out, p.rotate_base, s);
    if (st) return st;
    if (ev_joints && !p.regressed) CU_TRY(cudaEventRecord(ev_joints, s));
    // One launch per <= kFzMaxBodiesPerLaunch bodies: every vertex tile re-reads the launch's coef and A'
    // images (119 KB per 64 bodies), which must stay L2-resident under the stream of vertex writes -- at
    // 65,536 bodies in one launch they did not (7.6 GB of operand re-reads from HBM, 24 M instead of 29 M bodies/s).
    for (long long c0 = 0; c0 < n; c0 += kFzMaxBodiesPerLaunch) {
      const long long nc = std::min<long long>(kFzMaxBodiesPerLaunch, n - c0);
      CU_TRY(launch_fused_tc(model->d, model->num_sms, out.fz_coef + (size_t)(c0 / kFzBodies) * kFzCoefBlock,
                             out.fz_a + (size_t)(c0 / kFzBodies) * kFzSubs * kFzAImage, nc,
                             vertices + (size_t)c0 * model->d.V * 3, s));
    }
    if (p.regressed && (joints || kp2d)) {
      st = launch_regress(model, vertices, n, jbuf, cam, kp2d, s);
      if (st) return st;
      if (ev_joints) CU_TRY(cudaEventRecord(ev_joints, s));
    }
    return SMPLB200_OK;
  }
  if (p.prec == SMPLB200_PREC_FP32) out.coef = coef;
  if (prec_16bit(p.prec)) out.coef_bf16_hi = reinterpret_cast<uint16_t*>(ws + w.coef_hi);
  if (prec_split16(p.prec)) out.coef_bf16_lo = reinterpret_cast<uint16_t*>(ws + w.coef_lo);
  out.coef_is_f16 = p.prec == SMPLB200_PREC_F16X3 ? 1 : 0;
  if (p.prec == SMPLB200_PREC_TF32) out.coef_tf32 = reinterpret_cast<uint32_t*>(ws + w.coef_tf32);
  if (p.lbs == SMPLB200_LBS_TC) out.a_tf32 = reinterpret_cast<uint32_t*>(ws + w.a_tf32);
  // k4 for kinematic joints rides in k2 (the joints are final there); the skinning epilogue keeps its
  // projection for the stand-alone smplb200_lbs entry point
  if (kp2d != nullptr && !p.regressed) { out.cam = cam; out.kp2d = kp2d; }
  int st = launch_chain(model, betas, pose, n, out, p.rotate_base, s);
  if (st) return st;
  if (ev_joints && !p.regressed) CU_TRY(cudaEventRecord(ev_joints, s));   // joints + kp2d are final

  const bool proj_in_lbs = false;
  const bool chunked = model->chunk > 0 && p.prec != SMPLB200_PREC_FP32 && p.lbs == SMPLB200_LBS_TC &&
                       n > model->chunk;
  if (chunked) {
    // k1 -> k3 per chunk of bodies: the planar vposed scratch (chunk * 83 KB) is written by k1 and
    // consumed by k3 while still L2-resident, so it costs L2 bandwidth but (mostly) no HBM traffic.
    const long long C = model->chunk;   // multiple of kCoefBlock (128) and kLbsBlock (8)
    const int V = model->d.V;
    for (long long c0 = 0; c0 < n; c0 += C) {
      const long long nc = std::min<long long>(C, n - c0);
      const size_t cb = (size_t)(c0 / kCoefBlock) * kCoefBlock * kCoefK;   // elements into coef images
      CU_TRY(launch_blend_tc_any(model, p.prec,
                             out.coef_bf16_hi ? out.coef_bf16_hi + cb : nullptr,
                             out.coef_bf16_lo ? out.coef_bf16_lo + cb : nullptr,
                             out.coef_tf32 ? out.coef_tf32 + cb : nullptr, nc, vposed, s));
      const size_t ab = (size_t)(c0 / kLbsBlock) * kLbsBlock * 12 * kLbsK;  // elements into A' images
      CU_TRY(launch_lbs_tc(model->d, model->num_sms, vposed, out.a_tf32 + ab, nc,
                           vertices + (size_t)c0 * V * 3,
                           proj_in_lbs ? jbuf + (size_t)c0 * kJ * 3 : nullptr,
                           proj_in_lbs ? cam + (size_t)c0 * 3 : nullptr,
                           proj_in_lbs ? kp2d + (size_t)c0 * kJ * 2 : nullptr, s));
    }
  } else {
  // k1
  if (p.prec == SMPLB200_PREC_FP32)
    st = launch_blend_fma(model, coef, n, vposed, s);
  else
    CU_TRY(launch_blend_tc_any(model, p.prec, out.coef_bf16_hi, out.coef_bf16_lo, out.coef_tf32, n,
                               vposed, s));
  if (st) return st;

  // k3 (+k4 when joints are kinematic)
  if (p.lbs == SMPLB200_LBS_TC)
    CU_TRY(launch_lbs_tc(model->d, model->num_sms, vposed, out.a_tf32, n, vertices,
                         proj_in_lbs ? jbuf : nullptr, proj_in_lbs ? cam : nullptr,
                         proj_in_lbs ? kp2d : nullptr, s));
  else
    st = launch_lbs_fma(model, p.lbs == SMPLB200_LBS_DENSE, vposed, A, n, vertices,
                        proj_in_lbs ? jbuf : nullptr, proj_in_lbs ? cam : nullptr,
                        proj_in_lbs ? kp2d : nullptr, s);
  if (st) return st;
  }

  if (p.regressed && (joints || kp2d)) st = launch_regress(model, vertices, n, jbuf, cam, kp2d, s);
  if (st) return st;
  if (ev_joints && p.regressed) CU_TRY(cudaEventRecord(ev_joints, s));
  return st;
}

}  // extern "C"

// ---- backward ------------------------------------------------------------------------------
namespace {
struct BwdWorkspace {
  Workspace fwd;                 // recomputed forward intermediates (coef / operand images, A, vposed)
  size_t g_vposed = 0, g_A = 0, part = 0;
  int slices = 1;
  size_t total = 0;
};

// kb1 runs on tcgen05 except for an explicit SMPLB200_PREC_FP32 below the tensor-core batch size
// (CUDA-core FMA kernel).  Operands: split bf16 under AUTO / BF16X3, 3xTF32 for explicit FP32 at
// large batches, plain TF32 for the reduced-precision modes.
constexpr long long kBwdFp32TcMinBatch = 256;   // explicit FP32: CUDA-core FMA kernel below, 3xTF32 from here
inline bool bwd_blend_tc(uint32_t flags, long long n) {
  return (flags & SMPLB200_PREC_MASK) != SMPLB200_PREC_FP32 || n >= kBwdFp32TcMinBatch;
}
// The fused forward (SMPLB200_PREC_F16) leaves no vposed behind: its backward recomputes A and vposed with
// the unfused split-bf16 kernels (gradients do not depend on which forward precision produced the outputs).
inline void bwd_plan(Plan* p, uint32_t* flags) {
  if (p->prec == SMPLB200_PREC_F16) {
    p->prec = SMPLB200_PREC_F16X3;
    *flags = (*flags & ~SMPLB200_PREC_MASK) | SMPLB200_PREC_F16X3;
  }
}
inline int bwd_blend_mode(uint32_t flags, const Plan& p) {
  if ((flags & SMPLB200_PREC_MASK) == SMPLB200_PREC_FP32) return kBwdTf32x3;
  if (p.prec == SMPLB200_PREC_FP32 || prec_split16(p.prec)) return kBwdBf16x3;   // AUTO, BF16X3, F16X3
  return kBwdTf32;
}

BwdWorkspace carve_bwd(const SmplB200Model* m, long long n, const Plan& p, bool vertex_path, bool tc) {
  BwdWorkspace w;
  if (!vertex_path) { w.total = 256; return w; }
  Plan pf = p;
  pf.lbs = SMPLB200_LBS_FMA;      // no skinning in the recompute: no A' image needed
  w.fwd = carve(m, n, pf);
  size_t off = w.fwd.total;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t nn = (size_t)std::max<long long>(n, 1);
  // column slices of the blendshape backward: enough CTAs for ~2 per SM, at most 81
  const long long body_tiles = (long long)(tc ? (nn + kBwdTcBodies - 1) / kBwdTcBodies : (nn + kBbBodies - 1) / kBbBodies);
  const int nunits = tc ? m->d.NC / kBwdTcStepCols : m->d.NC / kBbCols;
  long long s = tc ? (2LL * m->num_sms) / body_tiles : (2LL * m->num_sms + body_tiles - 1) / body_tiles;
  w.slices = (int)std::max<long long>(1, std::min<long long>(s, std::min(kBbMaxSlices, nunits)));
  w.g_vposed = take(nn * 3 * (size_t)m->d.VP * sizeof(float));
  w.g_A = take(nn * kJ * 12 * sizeof(float));
  w.part = take((size_t)w.slices * nn * kCoefK * sizeof(float));
  w.total = off;
  return w;
}
}  // namespace

extern "C" {

size_t smplb200_backward_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags,
                                         int vertex_path) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return 0;
  bwd_plan(&p, &flags);
  return carve_bwd(model, n, p, vertex_path != 0, bwd_blend_tc(flags, n)).total;
}

int smplb200_backward_launch_count(const SmplB200Model* model, int64_t n, uint32_t flags, int vertex_path,
                                   int reuse_forward_workspace) {
  Plan p;
  if (!model || n <= 0 || !resolve_plan(model, n, flags, &p)) return 0;
  if (!vertex_path) return 1;                                  // kb2 alone
  if (p.prec == SMPLB200_PREC_F16) return 5;                   // the fused forward keeps no vposed: always recomputed
  return (reuse_forward_workspace && model->chunk == 0) ? 3 : 5;   // [k2, k1,] kb3, kb1, kb2
}

int smplb200_backward(const SmplB200Model* model, const float* betas, const float* pose,
                      const float* cam, int64_t n, const float* joints_fwd,
                      const float* g_vertices, const float* g_joints, const float* g_kp2d,
                      float* g_betas, float* g_pose, float* g_cam,
                      const void* forward_workspace, size_t forward_workspace_bytes,
                      void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (!model || n < 0) return SMPLB200_ERR_INVALID_ARG;
  if (n == 0) return SMPLB200_OK;
  if (!betas || !pose || !g_betas || !g_pose) return SMPLB200_ERR_INVALID_ARG;
  if ((g_kp2d || g_cam) && !cam) return SMPLB200_ERR_INVALID_ARG;
  Plan p;
  if (!resolve_plan(model, n, flags, &p)) return SMPLB200_ERR_INVALID_ARG;
  if (p.prec == SMPLB200_PREC_F16) { forward_workspace = nullptr; forward_workspace_bytes = 0; }
  bwd_plan(&p, &flags);
  if (p.regressed && g_kp2d && !joints_fwd) return SMPLB200_ERR_INVALID_ARG;
  const bool vertex_path = g_vertices != nullptr || (p.regressed && (g_joints || g_kp2d));
  const bool tc = bwd_blend_tc(flags, n);
  const BwdWorkspace w = carve_bwd(model, n, p, vertex_path, tc);
  if (vertex_path &&
      (!workspace || workspace_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & 255u)))
    return SMPLB200_ERR_WORKSPACE;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);

  ChainBwdArgs cb{};
  cb.betas = betas; cb.pose = pose; cb.cam = cam;
  cb.g_joints = g_joints; cb.g_kp2d = g_kp2d; cb.joints_fwd = joints_fwd;
  cb.g_betas = g_betas; cb.g_pose = g_pose; cb.g_cam = g_cam;
  cb.rotate_base = p.rotate_base ? 1 : 0;
  cb.regressed = p.regressed ? 1 : 0;
  cb.slices = w.slices;

  if (vertex_path) {
    float* coef = reinterpret_cast<float*>(ws + w.fwd.coef);
    float* A = reinterpret_cast<float*>(ws + w.fwd.A);
    float* vposed = reinterpret_cast<float*>(ws + w.fwd.vposed);
    float* g_vposed = reinterpret_cast<float*>(ws + w.g_vposed);
    float* g_A = reinterpret_cast<float*>(ws + w.g_A);
    float* part = reinterpret_cast<float*>(ws + w.part);
    int st = SMPLB200_OK;
    const Workspace fw = carve(model, n, p);      // layout of the forward call's own workspace
    if (forward_workspace && model->chunk == 0 && forward_workspace_bytes >= fw.total &&
        (reinterpret_cast<uintptr_t>(forward_workspace) & 255u) == 0) {
      // reuse the intermediates the forward left behind
      const uint8_t* f = static_cast<const uint8_t*>(forward_workspace);
      A = const_cast<float*>(reinterpret_cast<const float*>(f + fw.A));
      vposed = const_cast<float*>(reinterpret_cast<const float*>(f + fw.vposed));
    } else {
    // recompute k2 + k1 (same kernels and precision as the forward)
    ChainOut out{};
    out.A = A;
    if (p.prec == SMPLB200_PREC_FP32) out.coef = coef;
    if (prec_16bit(p.prec)) out.coef_bf16_hi = reinterpret_cast<uint16_t*>(ws + w.fwd.coef_hi);
    if (prec_split16(p.prec)) out.coef_bf16_lo = reinterpret_cast<uint16_t*>(ws + w.fwd.coef_lo);
    out.coef_is_f16 = p.prec == SMPLB200_PREC_F16X3 ? 1 : 0;
    if (p.prec == SMPLB200_PREC_TF32) out.coef_tf32 = reinterpret_cast<uint32_t*>(ws + w.fwd.coef_tf32);
    st = launch_chain(model, betas, pose, n, out, p.rotate_base, s);
    if (st) return st;
    if (p.prec == SMPLB200_PREC_FP32) st = launch_blend_fma(model, coef, n, vposed, s);
    else CU_TRY(launch_blend_tc_any(model, p.prec, out.coef_bf16_hi, out.coef_bf16_lo, out.coef_tf32, n, vposed, s));
    if (st) return st;
    }
    // kb3
    LbsBwdArgs la{};
    la.vposed = vposed; la.A = A; la.g_verts = g_vertices;
    la.g_joints = g_joints; la.g_kp2d = g_kp2d; la.cam = cam;
    la.g_vposed = g_vposed; la.g_A = g_A; la.regressed = p.regressed ? 1 : 0;
    const unsigned grid = (unsigned)std::min<long long>(n, 4LL * model->num_sms);
    if (model->lbs_bwd_staged)
      k_lbs_bwd_split<<<std::min<unsigned>(grid, (unsigned)model->num_sms), kLbsBwdThreads,
                        lbs_bwd_smem_bytes(model->d.V, model->d.VP), s>>>(model->d, la, n);
    else
      k_lbs_bwd<false><<<grid, kLbsBwdThreads, 0, s>>>(model->d, la, n);
    CU_TRY(cudaGetLastError());
    // kb1
    if (!tc) {
      dim3 g1((unsigned)w.slices, (unsigned)((n + kBbBodies - 1) / kBbBodies));
      k_blend_bwd_fma<<<g1, kBbThreads, kBbSmemBytes, s>>>(model->d, g_vposed, n, w.slices, part);
      CU_TRY(cudaGetLastError());
    } else {   // tcgen05: 3xTF32 for the fp32-class modes, 1xTF32 for the reduced-precision ones
      CU_TRY(launch_blend_bwd_tc(model->d, bwd_blend_mode(flags, p), g_vposed, n, w.slices, part, s));
    }
    cb.g_A = g_A;
    cb.g_coef_part = part;
  }
  const unsigned grid = (unsigned)((n + kChainWarps - 1) / kChainWarps);
  k_chain_bwd<<<grid, kChainWarps * 32, 0, s>>>(model->d, cb, n);
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

}  // extern "C"

extern "C" {

int smplb200_decode_gather(int32_t device, const float* heat, int32_t batch, int32_t num_classes,
                           int32_t height, int32_t width, const float* const* heads,
                           const int32_t* head_channels, int32_t num_heads, int32_t k,
                           float* scores, int64_t* inds, int32_t* clses, float* ys, float* xs,
                           float* const* gathered, void* stream) {
  if (batch < 0 || num_classes < 1 || height < 1 || width < 1 || num_heads < 0 ||
      num_heads > kDecMaxHeads || k < 1 || k > kDecMaxK)
    return SMPLB200_ERR_INVALID_ARG;
  if ((long long)num_classes * height * width < k || (long long)num_classes * height * width > (1LL << 30))
    return SMPLB200_ERR_INVALID_ARG;
  if (batch == 0) return SMPLB200_OK;
  if (!heat || !scores || !inds || !clses || !ys || !xs) return SMPLB200_ERR_INVALID_ARG;
  if (num_heads > 0 && (!heads || !head_channels || !gathered)) return SMPLB200_ERR_INVALID_ARG;
  DecodeHeads dh{};
  dh.n = num_heads;
  for (int h = 0; h < num_heads; ++h) {
    if (!heads[h] || !gathered[h] || head_channels[h] < 1) return SMPLB200_ERR_INVALID_ARG;
    dh.src[h] = heads[h]; dh.dst[h] = gathered[h]; dh.ch[h] = head_channels[h];
  }
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  const size_t key_bytes = (size_t)num_classes * height * width * sizeof(uint32_t);
  if (key_bytes <= 200 * 1024) {
    CU_TRY(cudaFuncSetAttribute(k_decode_gather<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                200 * 1024));   // the largest staged size: never lowered by a concurrent caller
    k_decode_gather<true><<<(unsigned)batch, kDecThreads, key_bytes, static_cast<cudaStream_t>(stream)>>>(
        heat, num_classes, height, width, k, dh, scores, reinterpret_cast<long long*>(inds), clses, ys, xs);
  } else {
    k_decode_gather<false><<<(unsigned)batch, kDecThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        heat, num_classes, height, width, k, dh, scores, reinterpret_cast<long long*>(inds), clses, ys, xs);
  }
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

size_t smplb200_dcn_v2_workspace_bytes(int32_t batch, int32_t channels_in, int32_t height, int32_t width,
                                       int32_t channels_out, uint32_t flags) {
  if (channels_in < 32 || channels_in % 32 || channels_out < 16 || channels_out % 16 || channels_out > kDcnMaxCo ||
      batch < 0 || height < 1 || width < 1 || (flags & ~SMPLB200_DCN_INPUT_NHWC))
    return 0;
  size_t bytes = align_up(dcn_weight_image_bytes(channels_in, channels_out), 256);
  if (!(flags & SMPLB200_DCN_INPUT_NHWC))
    bytes += align_up((size_t)std::max(batch, 1) * channels_in * height * width * sizeof(float), 256);
  return bytes;
}

int smplb200_dcn_v2_forward(int32_t device, const float* input, const float* weight, const float* bias,
                            const float* offset, const float* mask, int32_t batch, int32_t channels_in,
                            int32_t height, int32_t width, int32_t channels_out,
                            int32_t kernel_h, int32_t kernel_w, int32_t stride_h, int32_t stride_w,
                            int32_t pad_h, int32_t pad_w, int32_t dilation_h, int32_t dilation_w,
                            int32_t deformable_group, float* output,
                            void* workspace, size_t workspace_bytes, uint32_t flags, void* stream) {
  if (batch < 0 || channels_in < 1 || channels_out < 1 || height < 1 || width < 1 || stride_h < 1 ||
      stride_w < 1 || pad_h < 0 || pad_w < 0 || dilation_h < 1 || dilation_w < 1 || kernel_h < 1 || kernel_w < 1 ||
      (flags & ~SMPLB200_DCN_INPUT_NHWC))
    return SMPLB200_ERR_INVALID_ARG;
  if (kernel_h != 3 || kernel_w != 3 || deformable_group != 1) return SMPLB200_ERR_UNSUPPORTED;
  const size_t need = smplb200_dcn_v2_workspace_bytes(batch, channels_in, height, width, channels_out, flags);
  if (need == 0) return SMPLB200_ERR_UNSUPPORTED;
  DcnShape sh{};
  sh.B = batch; sh.Ci = channels_in; sh.H = height; sh.W = width; sh.Co = channels_out;
  sh.sh = stride_h; sh.sw = stride_w; sh.ph = pad_h; sh.pw = pad_w; sh.dh = dilation_h; sh.dw = dilation_w;
  sh.Ho = (height + 2 * pad_h - (dilation_h * (kernel_h - 1) + 1)) / stride_h + 1;
  sh.Wo = (width + 2 * pad_w - (dilation_w * (kernel_w - 1) + 1)) / stride_w + 1;
  if (sh.Ho < 1 || sh.Wo < 1) return SMPLB200_ERR_INVALID_ARG;
  if ((long long)batch * height * width >= (1LL << 31)) return SMPLB200_ERR_UNSUPPORTED;   // pixel offsets are int32
  if (batch > 65535 && !(flags & SMPLB200_DCN_INPUT_NHWC)) return SMPLB200_ERR_UNSUPPORTED;   // layout copy: batch on grid.z
  if (batch == 0) return SMPLB200_OK;
  if (!input || !weight || !offset || !mask || !output) return SMPLB200_ERR_INVALID_ARG;
  if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return SMPLB200_ERR_WORKSPACE;
  if (!aligned16(input)) return SMPLB200_ERR_ALIGNMENT;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const long long welems = (long long)channels_out * channels_in * kDcnTaps;
  k_dcn_pack_w<<<(unsigned)((welems + 255) / 256), 256, 0, s>>>(weight, channels_out, channels_in,
                                                              reinterpret_cast<uint16_t*>(ws));
  CU_TRY(cudaGetLastError());
  const float* nhwc = input;
  if (!(flags & SMPLB200_DCN_INPUT_NHWC)) {
    float* xt = reinterpret_cast<float*>(ws + align_up(dcn_weight_image_bytes(channels_in, channels_out), 256));
    const long long HW = (long long)height * width;
    dim3 g((unsigned)((HW + 31) / 32), (unsigned)(channels_in / 32), (unsigned)batch);
    k_nchw_to_nhwc<<<g, 256, 0, s>>>(input, channels_in, HW, xt);
    CU_TRY(cudaGetLastError());
    nhwc = xt;
  }
  const size_t smem = dcn_smem_bytes(channels_out);
  // always opt in to the largest size any supported shape needs: concurrent callers with different
  // Co must not lower each other's limit between the attribute call and the launch
  CU_TRY(cudaFuncSetAttribute(k_dcn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)std::max(dcn_smem_bytes(kDcnMaxCo), dcn_smem_bytes(128))));
  sh.tiles_x = (sh.Wo + kDcnTileW - 1) / kDcnTileW;
  sh.tiles_y = (sh.Ho + kDcnTileH - 1) / kDcnTileH;
  const long long ctas = (long long)batch * sh.tiles_x * sh.tiles_y;
  if (ctas >= (1LL << 31)) return SMPLB200_ERR_UNSUPPORTED;
  const uint32_t idesc = ptx::make_idesc(ptx::kFmtBF16, 128, (uint32_t)channels_out);
  int num_sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  const unsigned grid = (unsigned)std::min<long long>(ctas, num_sms);      // persistent: one CTA per SM
  k_dcn_fwd<<<grid, kDcnThreads, smem, s>>>(
      nhwc, offset, mask, ws, bias, sh, idesc, (int)ctas, dcn_stages_b(channels_out), output);
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

size_t smplb200_host_staging_bytes(const SmplB200Model* model, int64_t n, uint32_t flags) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return 0;
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  size_t off = 0;
  auto take = [&](size_t b) { off = align_up(off + b, 256); };
  take(nn * model->d.NB * 4); take(nn * 3 * kJ * 4); take(nn * 3 * 4);
  take(nn * (size_t)model->d.V * 3 * 4); take(nn * kJ * 3 * 4); take(nn * kJ * 2 * 4);
  return off + carve(model, n, p).total;
}

int smplb200_forward_host(const SmplB200Model* model, const float* betas_host,
                          const float* pose_host, const float* cam_host, int64_t n,
                          float* vertices_host, float* joints_host, float* kp2d_host,
                          void* staging, size_t staging_bytes, uint32_t flags, void* stream) {
  return smplb200_forward_host_opts(model, betas_host, pose_host, cam_host, n, vertices_host, joints_host,
                                    kp2d_host, staging, staging_bytes, flags, stream, nullptr);
}

int smplb200_host_staging_layout(const SmplB200Model* model, int64_t n, uint32_t flags,
                                 size_t* joints_offset, size_t* kp2d_offset) {
  Plan p;
  if (!model || n < 0 || !resolve_plan(model, n, flags, &p)) return SMPLB200_ERR_INVALID_ARG;
  const size_t nn = (size_t)std::max<int64_t>(n, 1);
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  take(nn * model->d.NB * 4); take(nn * 3 * kJ * 4); take(nn * 3 * 4);
  take(nn * (size_t)model->d.V * 3 * 4);
  const size_t oj = take(nn * kJ * 3 * 4), ok = take(nn * kJ * 2 * 4);
  if (joints_offset) *joints_offset = oj;
  if (kp2d_offset) *kp2d_offset = ok;
  return SMPLB200_OK;
}

int smplb200_forward_host_opts(const SmplB200Model* model, const float* betas_host,
                               const float* pose_host, const float* cam_host, int64_t n,
                               float* vertices_host, float* joints_host, float* kp2d_host,
                               void* staging, size_t staging_bytes, uint32_t flags, void* stream,
                               const SmplB200ForwardOpts* opts) {
  if (!model || n < 0) return SMPLB200_ERR_INVALID_ARG;
  if (n == 0) return smplb200_forward_opts(model, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                                           nullptr, 0, flags, stream, opts);
  if (!betas_host || !pose_host) return SMPLB200_ERR_INVALID_ARG;
  if (kp2d_host && !cam_host) return SMPLB200_ERR_INVALID_ARG;
  const size_t need = smplb200_host_staging_bytes(model, n, flags);
  if (need == 0) return SMPLB200_ERR_INVALID_ARG;
  if (!staging || staging_bytes < need || (reinterpret_cast<uintptr_t>(staging) & 255u))
    return SMPLB200_ERR_WORKSPACE;
  DeviceGuard guard(model->device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nn = (size_t)n;
  uint8_t* base = static_cast<uint8_t*>(staging);
  size_t off = 0;
  auto take = [&](size_t b) { uint8_t* p = base + off; off = align_up(off + b, 256); return p; };
  float* d_betas = reinterpret_cast<float*>(take(nn * model->d.NB * 4));
  float* d_pose = reinterpret_cast<float*>(take(nn * 3 * kJ * 4));
  float* d_cam = reinterpret_cast<float*>(take(nn * 3 * 4));
  float* d_verts = reinterpret_cast<float*>(take(nn * (size_t)model->d.V * 3 * 4));
  float* d_joints = reinterpret_cast<float*>(take(nn * kJ * 3 * 4));
  float* d_kp = reinterpret_cast<float*>(take(nn * kJ * 2 * 4));
  void* ws = base + off;
  CU_TRY(cudaMemcpyAsync(d_betas, betas_host, nn * model->d.NB * 4, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(d_pose, pose_host, nn * 3 * kJ * 4, cudaMemcpyHostToDevice, s));
  if (cam_host) CU_TRY(cudaMemcpyAsync(d_cam, cam_host, nn * 3 * 4, cudaMemcpyHostToDevice, s));
  const bool proj = cam_host != nullptr;
  int st = smplb200_forward_opts(model, d_betas, d_pose, proj ? d_cam : nullptr, n, d_verts, d_joints,
                                 proj ? d_kp : nullptr, ws, staging_bytes - off, flags, stream, opts);
  if (st) return st;
  if (vertices_host)
    CU_TRY(cudaMemcpyAsync(vertices_host, d_verts, nn * (size_t)model->d.V * 3 * 4, cudaMemcpyDeviceToHost, s));
  if (joints_host)
    CU_TRY(cudaMemcpyAsync(joints_host, d_joints, nn * kJ * 3 * 4, cudaMemcpyDeviceToHost, s));
  if (kp2d_host && proj)
    CU_TRY(cudaMemcpyAsync(kp2d_host, d_kp, nn * kJ * 2 * 4, cudaMemcpyDeviceToHost, s));
  return SMPLB200_OK;
}

// ---- multi-GPU exchange of joints | kp2d rows over peer-mapped memory (k_exchange.cuh) ----------
int smplb200_push_rows(int32_t device, const float* joints, const float* kp2d, int64_t n, int64_t row_offset,
                       void* const* peer_buffers, void* const* peer_flags, int32_t world, int32_t rank,
                       uint32_t epoch, void* counter, void* stream) {
  if (world < 1 || world > kXchgMaxRanks || rank < 0 || rank >= world || n < 0 || row_offset < 0 ||
      !peer_buffers || !peer_flags || !counter)
    return SMPLB200_ERR_INVALID_ARG;
  if (n > 0 && !joints) return SMPLB200_ERR_INVALID_ARG;
  if (!aligned16(joints) || !aligned16(kp2d)) return SMPLB200_ERR_ALIGNMENT;
  XchgPeers peers{};
  for (int r = 0; r < world; ++r) {
    if (!peer_buffers[r] || !peer_flags[r] || !aligned16(peer_buffers[r])) return SMPLB200_ERR_INVALID_ARG;
    peers.buf[r] = static_cast<float*>(peer_buffers[r]);
    peers.flags[r] = static_cast<uint32_t*>(peer_flags[r]);
  }
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  // Spread thin over ALL SMs: the compute kernels this overlaps are persistent with a static work split, so
  // what matters is the largest disturbance on any one SM, not the total.
  int num_sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  const long long items = (long long)n * (kXchgRow / 4);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((items + kXchgThreads - 1) / kXchgThreads, num_sms));
  k_push_rows<<<grid, kXchgThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      peers, world, rank, joints, kp2d, n, row_offset, epoch, static_cast<unsigned int*>(counter));
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

// The same exchange with NO kernel at all: copy-engine (DMA) peer copies + stream memory operations.  Peer STORES
// from SMs compete with the compute kernels' own global stores for each SM's store path (measured: the kernel push
// adds exactly the NVLink transfer time to the step, ~1.2 us per MB pushed); DMA copies do not touch the SMs.
int smplb200_exchange_rows_dma(int32_t device, const float* joints, const float* kp2d, int64_t n, int64_t row_offset,
                               int64_t rows_total, void* const* peer_slots, void* const* peer_flags,
                               int32_t world, int32_t rank, uint32_t epoch, void* stream) {
  if (world < 1 || world > kXchgMaxRanks || rank < 0 || rank >= world || n < 0 || row_offset < 0 ||
      rows_total < row_offset + n || !peer_slots || !peer_flags)
    return SMPLB200_ERR_INVALID_ARG;
  if (n > 0 && !joints) return SMPLB200_ERR_INVALID_ARG;
  for (int r = 0; r < world; ++r)
    if (!peer_slots[r] || !peer_flags[r]) return SMPLB200_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  // driver entry points for the stream memory operations, resolved through the runtime (no link-time libcuda)
  typedef int (*WriteFn)(void*, unsigned long long, uint32_t, unsigned int);
  typedef int (*WaitFn)(void*, unsigned long long, uint32_t, unsigned int);
  static WriteFn write32 = nullptr;          // function pointers are process-wide constants once resolved
  static WaitFn wait32 = nullptr;
  if (!write32 || !wait32) {
    void *w = nullptr, *q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CU_TRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &w, cudaEnableDefault, &qr));
    CU_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &q, cudaEnableDefault, &qr));
    if (!w || !q) return SMPLB200_ERR_UNSUPPORTED;
    write32 = reinterpret_cast<WriteFn>(w);
    wait32 = reinterpret_cast<WaitFn>(q);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t jrow = kJ * 3 * sizeof(float), krow = kJ * 2 * sizeof(float);
  for (int i = 0; i < world; ++i) {
    const int r = (rank + 1 + i) % world;          // every rank starts with a different peer
    uint8_t* slot = static_cast<uint8_t*>(peer_slots[r]);
    if (n > 0) {
      CU_TRY(cudaMemcpyAsync(slot + (size_t)row_offset * jrow, joints, (size_t)n * jrow, cudaMemcpyDeviceToDevice, s));
      if (kp2d)
        CU_TRY(cudaMemcpyAsync(slot + (size_t)rows_total * jrow + (size_t)row_offset * krow, kp2d, (size_t)n * krow,
                               cudaMemcpyDeviceToDevice, s));
    }
  }
  for (int i = 0; i < world; ++i) {                // in stream order: after every copy above has completed
    const int r = (rank + 1 + i) % world;
    const unsigned long long addr = (unsigned long long)(uintptr_t)peer_flags[r] + (unsigned long long)rank * 4ull;
    if (write32(s, addr, epoch, 0u /* CU_STREAM_WRITE_VALUE_DEFAULT */) != 0) return SMPLB200_ERR_CUDA;
  }
  for (int r = 0; r < world; ++r) {                // consumer side: all ranks' rows of this epoch have landed here
    const unsigned long long addr = (unsigned long long)(uintptr_t)peer_flags[rank] + (unsigned long long)r * 4ull;
    if (wait32(s, addr, epoch, 1u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) return SMPLB200_ERR_CUDA;
  }
  return SMPLB200_OK;
}

int smplb200_wait_rows(int32_t device, const void* my_flags, int32_t world, uint32_t epoch, void* stream) {
  if (world < 1 || world > kXchgMaxRanks || !my_flags) return SMPLB200_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  k_wait_rows<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(my_flags), world, epoch);
  CU_TRY(cudaGetLastError());
  return SMPLB200_OK;
}

int smplb200_probe_fp32_fma(int32_t device, int32_t iters, void* scratch, double* flop, void* stream) {
  if (iters < 1 || !scratch) return SMPLB200_ERR_INVALID_ARG;
  DeviceGuard guard(device);
  if (guard.err != cudaSuccess) return cuda_fail(guard.err);
  int num_sms = 0;
  CU_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  const unsigned grid = (unsigned)num_sms * 8u;
  k_probe_fma<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(scratch), iters, 0.999f, 1e-3f);
  CU_TRY(cudaGetLastError());
  if (flop) *flop = (double)grid * 256.0 * 16.0 * (double)iters;
  return SMPLB200_OK;
}

#ifdef SMPLB200_FZ_TIMING
int smplb200_debug_fz_timing(long long* out /* [148*32] */) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, smplb200::g_fz_time, sizeof(long long) * 148 * 32);
  static long long zero[148 * 32];
  cudaMemcpyToSymbol(smplb200::g_fz_time, zero, sizeof(zero));
  return e == cudaSuccess ? 0 : 5;
}
#endif
#ifdef SMPLB200_DEBUG_WAIT
int smplb200_debug_progress_buffer(unsigned int** host_ptr) {
  unsigned int* h = nullptr;
  if (cudaHostAlloc(&h, 148 * 16 * 4 * 3, cudaHostAllocMapped) != cudaSuccess) return 7;
  std::memset(h, 0xff, 148 * 16 * 4);
  std::memset(h + 148 * 16, 0, 148 * 16 * 4 * 2);
  unsigned int* d = nullptr;
  if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return 5;
  volatile unsigned int* dv = d;
  if (cudaMemcpyToSymbol(smplb200::ptx::g_prog, &dv, sizeof(dv)) != cudaSuccess) return 5;
  *host_ptr = h;
  return 0;
}
int smplb200_debug_wait_dump(unsigned int* out /* [244] */) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out, smplb200::ptx::g_wait_dbg, sizeof(unsigned int) * 244);
  unsigned int zero[244] = {0};
  cudaMemcpyToSymbol(smplb200::ptx::g_wait_dbg, zero, sizeof(zero));
  return e == cudaSuccess ? 0 : 5;
}
#endif

}  // extern "C"
