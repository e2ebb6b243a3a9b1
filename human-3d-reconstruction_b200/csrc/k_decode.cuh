// Fused decode -> gather: the producer of the SMPL layer's per-person parameter vectors.
//
// Replaces, in ONE launch and without the reference's NHWC permute-copy of every head
// (reference src/lib/models/utils.py:23-27 materialises feat.permute(0,2,3,1).contiguous() per
// gather), the chain the reference runs before the SMPL layer:
//     _nms   3x3 max-pool equality mask           reference src/lib/models/decode.py:6-13
//     _topk  per-class top-K, then top-K of C*K   reference src/lib/models/decode.py:26-41
//     _transpose_and_gather_feat(head, inds)      reference src/lib/models/utils.py:12-27
//
// One CTA per image.  The two-stage top-K of the reference is a global top-K by score, so the
// kernel selects the K largest NMS-ed values with an MSB-first radix select (3 passes of a
// 2048-bin shared-memory histogram over order-preserving uint32 keys), collects them, rank-sorts
// the K survivors by (score descending, flat index ascending) and gathers the head channels
// straight from NCHW.  Scores are the heat values themselves (heat * keep), so every output is
// bit-exact against the reference for inputs without score ties (torch.topk leaves the order of
// equal scores implementation-defined; here equal scores are taken lowest index first).
#pragma once
#include "common.cuh"

namespace smplb200 {

constexpr int kDecThreads = 512;
constexpr int kDecMaxK = 256;
constexpr int kDecMaxHeads = 8;

struct DecodeHeads {
  const float* src[kDecMaxHeads];   // [B, ch, H, W]
  float* dst[kDecMaxHeads];         // [B, K, ch]
  int ch[kDecMaxHeads];
  int n;
};

// NMS-ed value of flat element e = (c, y, x) of one image: heat if it equals its 3x3 max
// (padding = -inf, i.e. neighbours outside the map are ignored), else heat * 0.
__device__ __forceinline__ float nms_value(const float* __restrict__ img, int e, int H, int W) {
  const int hw = H * W;
  const int c = e / hw, r = e - c * hw, y = r / W, x = r - y * W;
  const float* p = img + (size_t)c * hw;
  const float v = __ldg(p + r);
  float mx = v;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W) continue;
      mx = fmaxf(mx, __ldg(p + yy * W + xx));
    }
  }
  return v == mx ? v : __fmul_rn(v, 0.f);
}
__device__ __forceinline__ uint32_t sortable_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// STAGED: the image's order-preserving keys are computed once (one 3x3 NMS pass) into dynamic
// shared memory and every later pass reads them from there; used whenever C*H*W*4 B fits
// (<= 200 KB, e.g. 128x128 x up to 3 classes).  Otherwise the keys are recomputed per pass.
template <bool STAGED>
__global__ void __launch_bounds__(kDecThreads)
k_decode_gather(const float* __restrict__ heat, int C, int H, int W, int K, DecodeHeads heads,
                float* __restrict__ scores, long long* __restrict__ inds, int* __restrict__ clses,
                float* __restrict__ ys, float* __restrict__ xs) {
  __shared__ uint32_t s_hist[2048];
  __shared__ uint32_t s_key[kDecMaxK];
  __shared__ int s_idx[kDecMaxK];
  __shared__ int s_rank_idx[kDecMaxK];
  __shared__ uint32_t s_prefix, s_need, s_count, s_eq_base, s_eq_total;
  __shared__ uint32_t s_warp_sum[kDecThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hw = H * W, M = C * hw;
  const float* img = heat + (size_t)b * M;
  extern __shared__ uint32_t s_keys[];
  if (STAGED) {
    for (int e = tid; e < M; e += kDecThreads) s_keys[e] = sortable_key(nms_value(img, e, H, W));
    __syncthreads();
  }
  auto key_of = [&](int e) -> uint32_t {
    return STAGED ? s_keys[e] : sortable_key(nms_value(img, e, H, W));
  };

  // ---- radix select: key T of the K-th largest element and how many == T to take ---------------
  if (tid == 0) { s_prefix = 0; s_need = (uint32_t)K; }
  __syncthreads();
  const int shifts[3] = {21, 10, 0};
  const uint32_t masks[3] = {0x7ffu, 0x7ffu, 0x3ffu};
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = tid; i < 2048; i += kDecThreads) s_hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t hi_mask = pass == 0 ? 0u : (pass == 1 ? 0xffe00000u : 0xfffffc00u);
    // NMS leaves ~8/9 of the map at exactly +0: those all hit one bin, so they are counted with a
    // warp ballot and ONE atomic per warp instead of 32 serialised ones.
    for (int e0 = 0; e0 < M; e0 += kDecThreads) {
      const int e = e0 + tid;
      const uint32_t key = e < M ? key_of(e) : 0u;
      const bool in = e < M && (key & hi_mask) == prefix;
      const bool zero = in && key == 0x80000000u;
      const uint32_t zb = __ballot_sync(0xffffffffu, zero);
      if (zero) {
        if (lane == (__ffs(zb) - 1)) atomicAdd(&s_hist[(key >> shifts[pass]) & masks[pass]], (uint32_t)__popc(zb));
      } else if (in) {
        atomicAdd(&s_hist[(key >> shifts[pass]) & masks[pass]], 1u);
      }
    }
    __syncthreads();
    {
      // Find the highest bin d whose suffix count (bins >= d) covers `need`: a block-wide scan from
      // the top, four bins per thread.  (A single thread walking up to 2047 bins of dependent
      // shared-memory loads cost ~60 K clk per pass -- most of the kernel's first version.)
      const uint32_t need = s_need;
      const int top = 2047 - 4 * tid;                       // this thread: bins top .. top-3
      const uint32_t h0 = s_hist[top], h1 = s_hist[top - 1], h2 = s_hist[top - 2], h3 = s_hist[top - 3];
      const uint32_t v = h0 + h1 + h2 + h3;
      uint32_t inc = v;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += nb;
      }
      __syncthreads();                                      // everyone has read s_need / s_hist
      if (lane == 31) s_warp_sum[warp] = inc;
      __syncthreads();
      uint32_t warp_excl = 0;
      for (int w = 0; w < warp; ++w) warp_excl += s_warp_sum[w];
      const uint32_t above = inc - v + warp_excl;           // elements in bins strictly above `top`
      if (above < need && above + v >= need) {              // the cut falls inside this thread's bins
        uint32_t a = above;
        int d = top;
        if (a + h0 < need) { a += h0; d = top - 1;
          if (a + h1 < need) { a += h1; d = top - 2;
            if (a + h2 < need) { a += h2; d = top - 3; } } }
        s_need = need - a;
        s_prefix = prefix | ((uint32_t)d << shifts[pass]);
        s_eq_total = d == top ? h0 : (d == top - 1 ? h1 : (d == top - 2 ? h2 : h3));   // after the last pass: #(key == T)
      }
    }
    __syncthreads();
  }
  const uint32_t T = s_prefix;         // key of the K-th largest NMS-ed value
  const uint32_t need_eq = s_need;     // how many elements with key == T belong to the top K

  // ---- collect: everything above T (any order), then == T in flat-index order ------------------
  if (tid == 0) { s_count = 0; s_eq_base = 0; }
  __syncthreads();
  if (s_eq_total == need_eq) {
    // no tie straddles the cut (the usual case for real scores): every element with key >= T is in
    // the top K, the order of collection is irrelevant (a rank sort follows) -> no scans, no barriers
    for (int e = tid; e < M; e += kDecThreads) {
      const uint32_t key = key_of(e);
      if (key >= T) {
        const uint32_t slot = atomicAdd(&s_count, 1u);
        s_key[slot] = key; s_idx[slot] = e;
      }
    }
    __syncthreads();
  } else
  for (int e0 = 0; e0 < M; e0 += kDecThreads) {
    const int e = e0 + tid;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (e < M) {
      key = key_of(e);
      gt = key > T; eq = key == T;
    }
    if (gt) {
      const uint32_t slot = atomicAdd(&s_count, 1u);
      s_key[slot] = key; s_idx[slot] = e;
    }
    // ordered selection among the ties: block-wide exclusive scan of the `eq` flags
    const uint32_t ball = __ballot_sync(0xffffffffu, eq);
    const uint32_t in_warp = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) s_warp_sum[warp] = __popc(ball);
    __syncthreads();
    uint32_t before = s_eq_base;
    for (int w = 0; w < warp; ++w) before += s_warp_sum[w];
    if (eq && before + in_warp < need_eq) {
      const uint32_t slot = atomicAdd(&s_count, 1u);
      s_key[slot] = key; s_idx[slot] = e;
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < kDecThreads / 32; ++w) tot += s_warp_sum[w];
      s_eq_base += tot;
    }
    __syncthreads();
  }

  // ---- rank sort of the K survivors: score descending, flat index ascending ---------------------
  if (tid < K) {
    const uint32_t ki = s_key[tid];
    const int ei = s_idx[tid];
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const uint32_t kj = s_key[j];
      rank += (kj > ki) || (kj == ki && s_idx[j] < ei);
    }
    s_rank_idx[rank] = ei;
    const int cls = ei / hw, r = ei - cls * hw;
    const size_t o = (size_t)b * K + rank;
    scores[o] = nms_value(img, ei, H, W);
    inds[o] = r;
    clses[o] = cls;
    ys[o] = (float)(r / W);
    xs[o] = (float)(r % W);
  }
  __syncthreads();

  // ---- gather every head channel at the K peaks, straight from NCHW -----------------------------
  for (int h = 0; h < heads.n; ++h) {
    const int ch = heads.ch[h];
    const float* src = heads.src[h] + (size_t)b * ch * hw;
    float* dst = heads.dst[h] + (size_t)b * K * ch;
    for (int i = tid; i < K * ch; i += kDecThreads) {
      const int k = i / ch, c = i - k * ch;
      const int e = s_rank_idx[k];
      dst[i] = __ldg(src + (size_t)c * hw + (e % hw));
    }
  }
}

}  // namespace smplb200
