// DCNv2 forward (modulated deformable 3x3 convolution) as ONE fused implicit GEMM on tcgen05 / TMEM.
//
// Replaces the reference's only native op for the forward direction:
//   modulated_deformable_im2col_cuda  reference src/lib/models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu:26-54,137-190
//   + per-image THCudaBlas_SgemmBatched (bias via a ones-vector GEMM, then weight x columns)
//                                     reference src/lib/models/DCNv2/src/cuda/dcn_v2_cuda.cu:43-173
// The reference materialises the column buffer [B, Ci*9, Ho*Wo] in global memory (37.7 MB per image
// at Ci = 256, 64x64) and reads it back in the GEMM.  Here the columns never exist in memory:
//
//   D[M = 128 output pixels (TMEM lanes), N = Co] += A[128 pixels, 32 k] * B[Co, 32 k]^T    per K-step
//
//   * A (the deformable samples) is produced by loader warps straight INTO TENSOR MEMORY.  The input
//     is read CHANNELS-LAST (a transposed copy made by k_nchw_to_nhwc, or the caller's own NHWC
//     tensor): a bilinear neighbour of a pixel is then one contiguous 128-byte line per 32 channels,
//     and 8 lanes fetch it with ONE 16-byte load each, so a warp instruction touches 4 full lines
//     instead of ~20 scattered 32-byte sectors (first version, NCHW gathers: l1tex 65-78 % busy at
//     ~20 sectors per load, profiles/r01_dcn_ncu.json).  Per K-step (one tap, 32 channels) a warp
//     makes 8 rounds of 4 pixels x 8 lanes, blends the 4 neighbours with the tap's weights, applies
//     the mask, transposes the 32 x 32 values through a private shared-memory tile so that lane =
//     pixel, splits them into bf16 hi + lo and writes its TMEM lanes with tcgen05.st.
//   * per pixel and tap the sampling geometry (4 pixel offsets, 4 bilinear weights with the border
//     rules and the modulation mask folded in) is computed ONCE into shared memory ([tap][pixel][8],
//     read back as two LDS.128) and reused for all Ci channels.
//   * B (the weights) is re-tiled on the device before the launch (k_dcn_pack_w) into K-major bf16
//     hi | lo tiles with K ordered tap-major (k' = tap * Ci + ci), one bulk-TMA copy per K-step.
//   * split-bf16 (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM): ~2^-16 relative, i.e. fp32-class
//     results from the bf16 tensor pipe.
//   * epilogue: thread = pixel reads its Co accumulators, adds the bias, and each warp store covers 32
//     adjacent pixels of one output channel (128 bytes, coalesced) in the NCHW output.
//
// Warp roles (576 threads, persistent over tiles): warp 0 = bulk-TMA producer of the weight tiles,
// warp 1 = MMA issuer (warp-uniform, one elected lane), warps 2..13 = three groups of four loader
// warps (TMEM lane quarter = warp % 4; group g takes K-steps i = g mod 3), warps 14..17 = auxiliary
// group: sampling geometry of the next tile (double-buffered table) and epilogue of the current one
// (double-buffered accumulator when Co <= 128).
#pragma once
#include "common.cuh"
#include "k_chain.cuh"
#include "ptx.cuh"

namespace smplb200 {

constexpr int kDcnLoadGroups = 3;                            // loader groups of four warps: warps 2..13
constexpr int kDcnAuxWarp0 = 2 + 4 * kDcnLoadGroups;         // warps 14..17: tap tables + epilogue
constexpr int kDcnThreads = (kDcnAuxWarp0 + 4) * 32;         // 576
constexpr int kDcnTaps = 9;
constexpr int kDcnTapVals = 8;                               // 4 offsets, 4 (bilinear weight x mask): 2 float4 per (tap, pixel)
constexpr int kDcnMaxStagesB = 3;                            // weight-tile ring (2 stages when Co > 128)
constexpr int kDcnStagesA = 8;                               // 32 TMEM columns each (16 hi | 16 lo)
constexpr int kDcnACol0 = 256;
constexpr int kDcnMaxCo = 256;
constexpr uint32_t kDcnTapBytes = kDcnTaps * kDcnTapVals * 128 * 4;   // 36,864

struct DcnShape {
  int B, Ci, H, W, Co, Ho, Wo;
  int sh, sw, ph, pw, dh, dw;
  int tiles_x, tiles_y;           // output tiles of kDcnTileW x kDcnTileH pixels per image
};

// A CTA's 128 output pixels form a 16 x 8 BLOCK, not a 128-pixel row segment: the input region its
// 9 taps x 4 neighbours sample is then ~(16+8) x (8+10) pixels (~110 KB at Ci = 64) and stays in L1,
// where a row segment samples ~(128+8) x 11 pixels (~380 KB) and re-fetched every line from L2
// (first versions: 4.8 GB of L2->SM traffic for the 64->64 @128x128 layer, the whole run time).
constexpr int kDcnTileW = 16, kDcnTileH = 8;

inline uint32_t dcn_stage_bytes(int Co) { return 128u * (uint32_t)Co; }     // [hi|lo][4 chunks][Co][8 bf16]
constexpr int kDcnTileStride = 36;                           // floats per pixel row of a warp's transpose tile
constexpr uint32_t kDcnTileBytes = 32 * kDcnTileStride * 4;   // 4,608 per loader warp
inline int dcn_stages_b(int Co) { return Co > 128 ? 2 : kDcnMaxStagesB; }
inline size_t dcn_smem_bytes(int Co) {
  return (size_t)dcn_stages_b(Co) * dcn_stage_bytes(Co) + 2 * (size_t)kDcnTapBytes +
         (size_t)4 * kDcnLoadGroups * kDcnTileBytes + 512;
}
inline size_t dcn_weight_image_bytes(int Ci, int Co) { return (size_t)kDcnTaps * (Ci / 32) * dcn_stage_bytes(Co); }

// input [B, C, HW] -> [B, HW, C] (channels-last copy read by the loaders)
__global__ void __launch_bounds__(256)
k_nchw_to_nhwc(const float* __restrict__ in, int C, long long HW, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* src = in + (size_t)blockIdx.z * C * HW;
  float* dst = out + (size_t)blockIdx.z * C * HW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k;
    const long long p = p0 + tx;
    tile[ty + 8 * k][tx] = (c < C && p < HW) ? __ldg(src + (size_t)c * HW + p) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long p = p0 + ty + 8 * k;
    const int c = c0 + tx;
    if (c < C && p < HW) dst[(size_t)p * C + c] = tile[tx][ty + 8 * k];
  }
}

// weight [Co, Ci, 3, 3] fp32 -> per K-step (tap, 32 channels) tile [hi|lo][4 chunks][Co rows][8 bf16]
__global__ void __launch_bounds__(256)
k_dcn_pack_w(const float* __restrict__ w, int Co, int Ci, uint16_t* __restrict__ img) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Co * Ci * kDcnTaps) return;
  const int tap = (int)(idx % kDcnTaps);
  const int ci = (int)((idx / kDcnTaps) % Ci);
  const int co = (int)(idx / ((long long)kDcnTaps * Ci));
  const float x = w[idx];
  const uint16_t hi = f32_to_bf16_rn(x);
  const uint16_t lo = f32_to_bf16_rn(x - bf16_to_f32(hi));
  const size_t ks = (size_t)tap * (Ci / 32) + ci / 32;
  const int c = (ci % 32) / 8, e = ci % 8;
  const size_t part = (size_t)4 * Co * 8;
  const size_t base = ks * 2 * part + ((size_t)c * Co + co) * 8 + e;
  img[base] = hi;
  img[base + part] = lo;
}

// Persistent, warp-specialised: a CTA loops over output tiles (tile = blockIdx.x + n * gridDim.x);
// the sampling geometry of tile n+1 and the epilogue of tile n are done by a dedicated auxiliary
// warp group while the loader groups and the tensor pipe are already on tile n+1, so the per-tile
// prologue / epilogue of the first (one CTA per tile) version -- ~20 % of its warp time sat at the
// final barrier (ncu) -- overlaps the main loop, and TMEM / barriers are set up once per SM.
__global__ void __launch_bounds__(kDcnThreads, 1)
k_dcn_fwd(const float* __restrict__ input /* NHWC */, const float* __restrict__ offset,
          const float* __restrict__ mask, const uint8_t* __restrict__ wimg, const float* __restrict__ bias,
          DcnShape s, uint32_t idesc, int ntiles, int nstB, float* __restrict__ output) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t stage_bytes = 128u * (uint32_t)s.Co;
  uint8_t* sB = smem;
  float* sTap = reinterpret_cast<float*>(smem + (size_t)nstB * stage_bytes);            // [2][tap][128 px][8]
  float* sTile = reinterpret_cast<float*>(smem + (size_t)nstB * stage_bytes + 2 * kDcnTapBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nstB * stage_bytes + 2 * kDcnTapBytes +
                                               4 * kDcnLoadGroups * kDcnTileBytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + kDcnMaxStagesB;
  uint64_t* a_full = b_empty + kDcnMaxStagesB;
  uint64_t* a_empty = a_full + kDcnStagesA;
  uint64_t* tap_full = a_empty + kDcnStagesA;     // [2] aux -> loaders
  uint64_t* tap_empty = tap_full + 2;             // [2] loaders -> aux
  uint64_t* d_full = tap_empty + 2;               // [2] MMA -> aux
  uint64_t* d_empty = d_full + 2;                 // [2] aux -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int steps_per_tap = s.Ci / 32;
  const int nks = kDcnTaps * steps_per_tap;
  const long long HoWo = (long long)s.Ho * s.Wo;
  const bool dbl = s.Co <= 128;                    // two accumulators fit next to the 8 A stages
  const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < kDcnMaxStagesB; ++i) { ptx::mbar_init(b_full + i, 1); ptx::mbar_init(b_empty + i, 1); }
    for (int i = 0; i < kDcnStagesA; ++i) { ptx::mbar_init(a_full + i, 4); ptx::mbar_init(a_empty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(tap_full + i, 4); ptx::mbar_init(tap_empty + i, 4 * kDcnLoadGroups);
      ptx::mbar_init(d_full + i, 1); ptx::mbar_init(d_empty + i, 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== bulk-TMA producer: one weight tile (hi | lo) per K-step, re-streamed (from L2) per tile =====
    if (lane == 0) {
      long long gi = 0;
      for (int n = 0; n < my_tiles; ++n)
        for (int i = 0; i < nks; ++i, ++gi) {
          const int st = (int)(gi % nstB);
          ptx::mbar_wait(b_empty + st, (uint32_t)((gi / nstB) & 1) ^ 1u);
          ptx::mbar_arrive_expect_tx(b_full + st, stage_bytes);
          ptx::bulk_g2s(sB + (size_t)st * stage_bytes, wimg + (size_t)i * stage_bytes, stage_bytes, b_full + st);
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t lbo = (uint32_t)s.Co * 16u, part = 4u * lbo;
    long long gi = 0;
    for (int n = 0; n < my_tiles; ++n) {
      const int db = dbl ? (n & 1) : 0;
      const int use = dbl ? (n >> 1) : n;                      // how often this accumulator was used before
      ptx::mbar_wait(d_empty + db, (uint32_t)(use & 1) ^ 1u);
      const uint32_t d_addr = tmem_base + db * 128;
      for (int i = 0; i < nks; ++i, ++gi) {
        const int sb = (int)(gi % nstB), sa = (int)(gi % kDcnStagesA);
        ptx::mbar_wait(b_full + sb, (uint32_t)((gi / nstB) & 1));
        ptx::mbar_wait(a_full + sa, (uint32_t)((gi / kDcnStagesA) & 1));
        ptx::tc_fence_after();
        const uint32_t b_addr = ptx::smem_u32(sB + (size_t)sb * stage_bytes);
        const uint32_t a_addr = tmem_base + kDcnACol0 + sa * 32;
        if (ptx::elect_one()) {
#pragma unroll
          for (int g = 0; g < 3; ++g) {          // (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo)
            const uint32_t ap = a_addr + (g == 1 ? 16 : 0);
            const uint32_t bp = b_addr + (g == 2 ? part : 0);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t bd = ptx::make_smem_desc(bp + kk * 2 * lbo, lbo, 128);
              ptx::mma_bf16_ts(d_addr, ap + kk * 8, bd, idesc, (uint32_t)((i | g | kk) != 0));
            }
          }
          ptx::tc_commit(b_empty + sb);
          ptx::tc_commit(a_empty + sa);
          if (i == nks - 1) ptx::tc_commit(d_full + db);
        }
        __syncwarp();
      }
    }
  } else if (warp < kDcnAuxWarp0) {
    // ===== loaders: 8 lanes fetch one neighbour line, transpose so that lane = pixel, write TMEM =====
    const int lw = warp - 2, q = warp & 3, grp = lw >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* tile = sTile + (size_t)lw * (32 * kDcnTileStride);
    const int sub = lane & 7, pq = lane >> 3;             // 8 lanes per pixel, 4 pixels per round
    long long gbase = 0;
    for (int n = 0; n < my_tiles; ++n, gbase += nks) {
      const int tb = n & 1;
      ptx::mbar_wait(tap_full + tb, (uint32_t)((n >> 1) & 1));
      const float* tap0 = sTap + (size_t)tb * (kDcnTapBytes / 4);
      for (int i = grp; i < nks; i += kDcnLoadGroups) {
        const long long gi = gbase + i;
        const int t = i / steps_per_tap, cb = (i - t * steps_per_tap) * 32;
        const float4* tvb = reinterpret_cast<const float4*>(tap0 + ((size_t)t * 128 + q * 32) * kDcnTapVals);
        const float* cbase = input + cb + 4 * sub;
        // 8 rounds (pixel pw = 4*rd + pq, channels cb + 4*sub .. +3) in two batches of four: all 16
        // line loads of a batch are issued before the first blend, so four rounds of latency overlap
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 v[4][4];
          float wv[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4* tv = tvb + (4 * (4 * half + u) + pq) * 2;       // this round's pixel: 2 x LDS.128
            const float4 to = tv[0], tw = tv[1];
            const int o0 = __float_as_int(to.x), o1 = __float_as_int(to.y), o2 = __float_as_int(to.z),
                      o3 = __float_as_int(to.w);
            wv[u][0] = tw.x; wv[u][1] = tw.y; wv[u][2] = tw.z; wv[u][3] = tw.w;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][0] = o0 >= 0 ? __ldg(reinterpret_cast<const float4*>(cbase + (size_t)o0 * s.Ci)) : z;
            v[u][1] = o1 >= 0 ? __ldg(reinterpret_cast<const float4*>(cbase + (size_t)o1 * s.Ci)) : z;
            v[u][2] = o2 >= 0 ? __ldg(reinterpret_cast<const float4*>(cbase + (size_t)o2 * s.Ci)) : z;
            v[u][3] = o3 >= 0 ? __ldg(reinterpret_cast<const float4*>(cbase + (size_t)o3 * s.Ci)) : z;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int pw = 4 * (4 * half + u) + pq;
            // sum_i (w_i * mask) * v_i : the modulation mask is folded into the four bilinear weights
            float4 r4;
            r4.x = wv[u][0] * v[u][0].x + wv[u][1] * v[u][1].x + wv[u][2] * v[u][2].x + wv[u][3] * v[u][3].x;
            r4.y = wv[u][0] * v[u][0].y + wv[u][1] * v[u][1].y + wv[u][2] * v[u][2].y + wv[u][3] * v[u][3].y;
            r4.z = wv[u][0] * v[u][0].z + wv[u][1] * v[u][1].z + wv[u][2] * v[u][2].z + wv[u][3] * v[u][3].z;
            r4.w = wv[u][0] * v[u][0].w + wv[u][1] * v[u][1].w + wv[u][2] * v[u][2].w + wv[u][3] * v[u][3].w;
            *reinterpret_cast<float4*>(tile + pw * kDcnTileStride + 4 * sub) = r4;
          }
        }
        __syncwarp();
        float val[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 x = *reinterpret_cast<const float4*>(tile + lane * kDcnTileStride + 4 * k);
          val[4 * k] = x.x; val[4 * k + 1] = x.y; val[4 * k + 2] = x.z; val[4 * k + 3] = x.w;
        }
        __syncwarp();
        const int sa = (int)(gi % kDcnStagesA);
        ptx::mbar_wait(a_empty + sa, (uint32_t)((gi / kDcnStagesA) & 1) ^ 1u);
        ptx::tc_fence_after();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {      // one packed cvt per pair; residuals from the packed hi word
          hi[u] = f32x2_to_bf16x2_rn(val[2 * u], val[2 * u + 1]);
          lo[u] = f32x2_to_bf16x2_rn(val[2 * u] - __uint_as_float(hi[u] << 16),
                                     val[2 * u + 1] - __uint_as_float(hi[u] & 0xffff0000u));
        }
        const uint32_t acol = tmem_base + lane_addr + kDcnACol0 + sa * 32;
        ptx::tmem_st16(acol, hi);
        ptx::tmem_st16(acol + 16, lo);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_full + sa);
        __syncwarp();
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tap_empty + tb);      // this warp is done with the tile's tap table
    }
  } else {
    // ===== auxiliary group: sampling geometry of the NEXT tile, epilogue of the CURRENT one =====
    const int q = warp & 3;
    const int px = q * 32 + lane;                         // pixel within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int tiles_per_img = s.tiles_x * s.tiles_y;
    auto pixel_of = [&](int tile, long long& b, int& ho, int& wo) -> bool {
      b = tile / tiles_per_img;
      const int tix = tile - (int)b * tiles_per_img;
      ho = (tix / s.tiles_x) * kDcnTileH + px / kDcnTileW;
      wo = (tix % s.tiles_x) * kDcnTileW + px % kDcnTileW;
      return ho < s.Ho && wo < s.Wo;
    };
    auto compute_taps = [&](int tile, float* tab) {
      long long b; int ho, wo;
      const bool valid = pixel_of(tile, b, ho, wo);
      const int r = valid ? ho * s.Wo + wo : 0;
#pragma unroll 3
      for (int t = 0; t < kDcnTaps; ++t) {
        float4* dst = reinterpret_cast<float4*>(tab + ((size_t)t * 128 + px) * kDcnTapVals);
        int o[4] = {-1, -1, -1, -1};        // -1: neighbour outside the map (contributes exactly 0)
        float wgt[4] = {0.f, 0.f, 0.f, 0.f};
        float mk = 0.f;
        if (valid) {
          const int i = t / 3, j = t - 3 * i;
          const float* offp = offset + ((size_t)b * 2 * kDcnTaps + 2 * t) * HoWo + r;
          const float off_h = __ldg(offp), off_w = __ldg(offp + HoWo);
          const float h_im = (float)(ho * s.sh - s.ph + i * s.dh) + off_h;
          const float w_im = (float)(wo * s.sw - s.pw + j * s.dw) + off_w;
          if (h_im > -1.f && w_im > -1.f && h_im < (float)s.H && w_im < (float)s.W) {
            mk = __ldg(mask + ((size_t)b * kDcnTaps + t) * HoWo + r);
            const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
            const int h_high = h_low + 1, w_high = w_low + 1;
            const float lh = h_im - (float)h_low, lwd = w_im - (float)w_low;
            const float hh = 1.f - lh, hw = 1.f - lwd;
            const bool t0 = h_low >= 0, t1 = h_high <= s.H - 1, l0 = w_low >= 0, l1 = w_high <= s.W - 1;
            const int img0 = (int)b * s.H * s.W;        // NHWC pixel index of the image's first pixel
            if (t0 && l0) { o[0] = img0 + h_low * s.W + w_low; wgt[0] = hh * hw; }
            if (t0 && l1) { o[1] = img0 + h_low * s.W + w_high; wgt[1] = hh * lwd; }
            if (t1 && l0) { o[2] = img0 + h_high * s.W + w_low; wgt[2] = lh * hw; }
            if (t1 && l1) { o[3] = img0 + h_high * s.W + w_high; wgt[3] = lh * lwd; }
          }
        }
        dst[0] = make_float4(__int_as_float(o[0]), __int_as_float(o[1]), __int_as_float(o[2]), __int_as_float(o[3]));
        dst[1] = make_float4(wgt[0] * mk, wgt[1] * mk, wgt[2] * mk, wgt[3] * mk);
      }
    };
    if (my_tiles > 0) {
      compute_taps((int)blockIdx.x, sTap);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tap_full + 0);
    }
    for (int n = 0; n < my_tiles; ++n) {
      const int tile = (int)blockIdx.x + n * (int)gridDim.x;
      if (n + 1 < my_tiles) {
        const int nb = (n + 1) & 1;
        ptx::mbar_wait(tap_empty + nb, (uint32_t)(((n + 1) >> 1) & 1) ^ 1u);    // loaders left that buffer
        compute_taps(tile + (int)gridDim.x, sTap + (size_t)nb * (kDcnTapBytes / 4));
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tap_full + nb);
      }
      // epilogue: D[pixel, co] + bias -> output[b, co, ho, wo]
      const int db = dbl ? (n & 1) : 0;
      const int use = dbl ? (n >> 1) : n;
      long long b; int ho, wo;
      const bool valid = pixel_of(tile, b, ho, wo);
      ptx::mbar_wait(d_full + db, (uint32_t)(use & 1));
      ptx::tc_fence_after();
      float* dst = output + (size_t)b * s.Co * HoWo + (valid ? ho * s.Wo + wo : 0);
#pragma unroll 1
      for (int c0 = 0; c0 < s.Co; c0 += 16) {
        uint32_t d[16];
        ptx::tmem_ld16(tmem_base + lane_addr + db * 128 + c0, d);
        ptx::tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 16; ++u)
            dst[(size_t)(c0 + u) * HoWo] = __uint_as_float(d[u]) + (bias ? __ldg(bias + c0 + u) : 0.f);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(d_empty + db);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace smplb200
