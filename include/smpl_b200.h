/*
 * smpl_b200.h -- C ABI of the B200-native SMPL body-model forward pass (libsmpl_b200.so).
 *
 * Drop-in boundary for the PyTorch SMPL layer that BASELINE.json's north_star names
 * (`forward(betas, pose, ...) -> (vertices, joints)`).  The mounted reference snapshot has no
 * SMPL layer, so the boundary follows the one native-op convention the reference
 * does have -- the DCNv2 extension:
 *   - `extern "C"` launchers taking raw `const float*` + sizes + `cudaStream_t`
 *       reference src/lib/models/DCNv2/src/cuda/dcn_v2_im2col_cuda.h:67-99
 *   - a thin tensor shim over them        reference src/lib/models/DCNv2/src/dcn_v2.h:9-39,
 *                                                   src/lib/models/DCNv2/src/vision.cpp:4-9
 *   - an nn.Module on top                 reference src/lib/models/DCNv2/dcn_v2.py:57-128
 * and deliberately departs from it where SURVEY.md §8(b) says so: the callee never allocates per
 * call (caller owns outputs and workspace), never prints, never throws, keeps no mutable globals
 * (contrast reference src/lib/models/DCNv2/src/cuda/dcn_v2_cuda.cu:12) and reports errors as an
 * `int` status.
 *
 * Conventions
 *   - all tensors are fp32, C-contiguous; "device" pointers live on the model's CUDA device;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no host sync;
 *   - thread-safe and re-entrant: one model handle per device, calls may come from any thread.
 *   - N = bodies, V = vertices (6890), J = joints (24), NB = betas (<=14), P = 9*(J-1) = 207.
 *
 * There is NO CPU fallback: every entry point below launches hand-written sm_100a kernels.
 */
#ifndef SMPL_B200_H
#define SMPL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMPLB200_VERSION 120 /* 0.2.0: + forward options (joints-ready event), peer-store row exchange */

/* ---- status codes ---------------------------------------------------------------------- */
enum {
  SMPLB200_OK = 0,
  SMPLB200_ERR_INVALID_ARG = 1,   /* null pointer, bad size, unknown flag               */
  SMPLB200_ERR_UNSUPPORTED = 2,   /* model shape outside what the kernels are built for  */
  SMPLB200_ERR_WORKSPACE = 3,     /* workspace null/too small/misaligned                 */
  SMPLB200_ERR_ALIGNMENT = 4,     /* an output pointer is not 16-byte aligned            */
  SMPLB200_ERR_CUDA = 5,          /* a CUDA runtime call or kernel launch failed         */
  SMPLB200_ERR_NO_DEVICE = 6,     /* no sm_100 device / device index out of range        */
  SMPLB200_ERR_ALLOC = 7          /* host or device allocation failed at model create    */
};

/* ---- forward flags --------------------------------------------------------------------- */
/* blendshape operand precision (accumulate and output are always fp32)                      */
#define SMPLB200_PREC_AUTO   0u  /* fastest path within 1e-5 m: FP32 FMA below SMPLB200_TC_MIN_BATCH bodies, F16X3 from there */
#define SMPLB200_PREC_FP32   1u  /* vectorised FMA kernel, k-ascending single accumulator     */
#define SMPLB200_PREC_BF16   2u  /* tcgen05 kind::f16, bf16 operands                          */
#define SMPLB200_PREC_TF32   3u  /* tcgen05 kind::tf32                                        */
#define SMPLB200_PREC_BF16X3 4u  /* tcgen05, 3-term split bf16 (hi*hi + hi*lo + lo*hi)        */
#define SMPLB200_PREC_F16    5u  /* FUSED blendshapes + skinning kernel (no v_posed round trip through HBM), fp16
                                    operands: pose rows one MMA, shape rows + template exact 3-term split, skinning
                                    blend 3-term split.  Measured max vertex error 1.9e-5 m (stated bound 5e-5 m;
                                    TF32 operands: 1.9e-4).  Needs NB <= 13; the LBS flag must be AUTO or TC.     */
#define SMPLB200_PREC_F16X3  6u  /* tcgen05, 3-term split FP16 (hi*hi + hi*lo + lo*hi; 11+11 significant bits per
                                    operand): fp32-class blendshapes, measured < 1e-6 m; what AUTO resolves to        */
#define SMPLB200_PREC_MASK   0x7u
/* joint output: kinematic J_posed (default) or HMR-style regression from skinned vertices   */
#define SMPLB200_JOINTS_KINEMATIC 0u
#define SMPLB200_JOINTS_REGRESSED (1u << 3)
/* HMR's root pre-rotation by diag(1,-1,-1); default off                      */
#define SMPLB200_ROTATE_BASE      (1u << 4)
/* skinning path                                                                              */
#define SMPLB200_LBS_AUTO   0u          /* tcgen05 blend from SMPLB200_TC_LBS_MIN_BATCH bodies, FMA below */
#define SMPLB200_LBS_FMA    (1u << 5)   /* CUDA-core kernel (ELL-sparse weights if <=4 nnz)  */
#define SMPLB200_LBS_TC     (2u << 5)   /* 3xTF32 tcgen05 blend of the 24 joint transforms   */
#define SMPLB200_LBS_DENSE  (3u << 5)   /* CUDA-core kernel, all 24 weights (debug/fallback) */
#define SMPLB200_LBS_MASK   (3u << 5)

/* Measured crossovers on B200 (profiles/r01_latency_crossover.txt, CUDA-graph replay): the tcgen05
 * blendshape kernel beats the FMA one from 32 bodies, the tcgen05 skinning blend from 384.        */
#define SMPLB200_TC_MIN_BATCH 32
#define SMPLB200_TC_LBS_MIN_BATCH 384

/* ---- model ----------------------------------------------------------------------------- */
typedef struct SmplB200Model SmplB200Model;

/* Host-side description of the model tensors in the layout the eager layer keeps them
 * (SURVEY.md App. A.1).  All pointers are HOST pointers; they are read during create only.   */
typedef struct SmplB200ModelDesc {
  uint32_t struct_size;      /* = sizeof(SmplB200ModelDesc)                                  */
  int32_t device;            /* CUDA device ordinal the handle lives on                      */
  int32_t num_verts;         /* V  (>= 1)                                                    */
  int32_t num_joints;        /* J  (must be 24)                                              */
  int32_t num_betas;         /* NB (1..14)                                                   */
  const float* v_template;   /* [V,3]                                                        */
  const float* shapedirs;    /* [NB, 3V]   column = 3*v + c                                  */
  const float* posedirs;     /* [9*(J-1), 3V]                                                */
  const float* j_regressor;  /* [V, J]     (the model's [J,V] transposed, HMR idiom)         */
  const float* weights;      /* [V, J]     skinning weights                                  */
  const int32_t* parents;    /* [J]        kintree_table[0]; root = -1 (or 0xFFFFFFFF)       */
} SmplB200ModelDesc;

/* Packs the model for the device (folded joint regressor, planar K-padded blendshape basis,
 * pre-tiled bf16/tf32 tensor-core operand images, ELL skinning weights) and uploads it.
 * Synchronous; the only entry point that allocates device memory.                            */
int smplb200_model_create(const SmplB200ModelDesc* desc, SmplB200Model** out_model);
void smplb200_model_destroy(SmplB200Model* model);

/* Introspection (pure host, no CUDA calls).                                                  */
int32_t smplb200_model_num_verts(const SmplB200Model* model);
int32_t smplb200_model_num_joints(const SmplB200Model* model);
int32_t smplb200_model_num_betas(const SmplB200Model* model);
int32_t smplb200_model_device(const SmplB200Model* model);
int32_t smplb200_model_max_weight_nnz(const SmplB200Model* model); /* max non-zeros per vertex */
size_t smplb200_model_device_bytes(const SmplB200Model* model);

/* ---- the forward pass ------------------------------------------------------------------ */
/* Bytes of device scratch `smplb200_forward` needs for `n` bodies with `flags`.              */
size_t smplb200_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags);

/* vertices[N,V,3], joints[N,J,3], kp2d[N,J,2] <- betas[N,NB], pose[N,3J], cam[N,3].
 * DEVICE pointers.  `cam`/`kp2d` may both be NULL (no projection); `joints` may be NULL.
 * `workspace` must be 256-byte aligned and >= smplb200_workspace_bytes(...).
 * Replaces: the forward() of the eager SMPL nn.Module.            */
int smplb200_forward(const SmplB200Model* model,
                     const float* betas, const float* pose, const float* cam, int64_t n,
                     float* vertices, float* joints, float* kp2d,
                     void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Options for the `_opts` variants below (NULL = none).
 *   joints_ready_event  a cudaEvent_t (as void*), recorded on `stream` as soon as `joints` (and `kp2d`)
 *                       are final: right after k2 for kinematic joints (~10 us into the step, the
 *                       projection rides in k2), after the regression kernel for regressed joints.
 *                       Lets the caller start the multi-GPU exchange of the small outputs on a side
 *                       stream while the blendshape / skinning kernels still run (SURVEY.md §5, §8e;
 *                       reference analogue: the DataParallel gather at src/lib/trains/trainer.py:176).   */
typedef struct SmplB200ForwardOpts {
  uint32_t struct_size;        /* = sizeof(SmplB200ForwardOpts) */
  void* joints_ready_event;    /* cudaEvent_t or NULL            */
} SmplB200ForwardOpts;
int smplb200_forward_opts(const SmplB200Model* model,
                          const float* betas, const float* pose, const float* cam, int64_t n,
                          float* vertices, float* joints, float* kp2d,
                          void* workspace, size_t workspace_bytes, uint32_t flags, void* stream,
                          const SmplB200ForwardOpts* opts);

/* Same pass with HOST buffers (pinned recommended): copies betas/pose/cam host->device,
 * runs the forward, copies the requested outputs device->host, all ordered on `stream`.
 * `vertices_host`/`joints_host`/`kp2d_host` may each be NULL to skip that copy.
 * `staging` is device scratch of >= smplb200_host_staging_bytes(model, n, flags) bytes
 * (inputs + outputs + workspace).  Returns without synchronising.                            */
size_t smplb200_host_staging_bytes(const SmplB200Model* model, int64_t n, uint32_t flags);
int smplb200_forward_host(const SmplB200Model* model,
                          const float* betas_host, const float* pose_host, const float* cam_host,
                          int64_t n, float* vertices_host, float* joints_host, float* kp2d_host,
                          void* staging, size_t staging_bytes, uint32_t flags, void* stream);

int smplb200_forward_host_opts(const SmplB200Model* model,
                               const float* betas_host, const float* pose_host, const float* cam_host,
                               int64_t n, float* vertices_host, float* joints_host, float* kp2d_host,
                               void* staging, size_t staging_bytes, uint32_t flags, void* stream,
                               const SmplB200ForwardOpts* opts);
/* Byte offsets, inside `staging`, of the DEVICE copies of joints[N,J,3] and kp2d[N,J,2] the host entry
 * produces (valid once `joints_ready_event` has fired): what a multi-GPU caller pushes to its peers.  */
int smplb200_host_staging_layout(const SmplB200Model* model, int64_t n, uint32_t flags,
                                 size_t* joints_offset, size_t* kp2d_offset);

/* ---- multi-GPU: exchange of the small per-body outputs over peer-mapped memory ----
 * Bodies are sharded over ranks with no data-path collective.  The one optional exchange -- every
 * rank's joints (288 B/body) and kp2d (192 B/body) visible on every rank, what the reference gets from
 * nn.DataParallel's gather (src/lib/trains/trainer.py:176) -- is done with plain peer stores over
 * NVLink / NVSwitch instead of an NCCL launch: `smplb200_push_rows` writes this rank's rows
 * [row_offset, row_offset + n) of the gathered buffer [n_total][120] fp32 (joints 72 | kp2d 48) into
 * EVERY rank's copy of it and then publishes `epoch` in flags[peer][rank]; `smplb200_wait_rows`
 * returns (on the stream) once all `world` flags of this rank have reached `epoch`.
 *   peer_buffers / peer_flags  HOST arrays of `world` DEVICE pointers, peer-mapped by the caller (CUDA
 *                              IPC or symmetric memory); entry `rank` is the local buffer / flag array
 *                              (uint32[world], zero-initialised); world <= 16
 *   counter                    a zero-initialised local uint32 in device memory (scratch of the call)
 * Epochs must increase from call to call (wrap-around is handled).  A peer that never arrives faults
 * the waiting launch after a few seconds instead of hanging the device.                               */
int smplb200_push_rows(int32_t device, const float* joints, const float* kp2d, int64_t n, int64_t row_offset,
                       void* const* peer_buffers, void* const* peer_flags, int32_t world, int32_t rank,
                       uint32_t epoch, void* counter, void* stream);
int smplb200_wait_rows(int32_t device, const void* my_flags, int32_t world, uint32_t epoch, void* stream);
/* The same exchange with no kernel at all: copy-engine peer copies of the two contiguous row blocks, then stream
 * memory operations (cuStreamWriteValue32 publishes `epoch` in every peer's flags[rank]; cuStreamWaitValue32 holds
 * `stream` until all `world` local flags have reached it).  Slot layout here: joints block [rows_total][72] fp32
 * followed by the kp2d block [rows_total][48].  Peer stores issued by SMs share each SM's store path with the
 * compute kernels; DMA copies do not.                                                                     */
int smplb200_exchange_rows_dma(int32_t device, const float* joints, const float* kp2d, int64_t n, int64_t row_offset,
                               int64_t rows_total, void* const* peer_slots, void* const* peer_flags,
                               int32_t world, int32_t rank, uint32_t epoch, void* stream);

/* Measurement aid for bench.py (the fp32-FMA peak is measured in the same run): launches
 * 8 CTAs per SM of 256 threads running 8 independent FMA chains for `iters` rounds; `*flop` receives the
 * flop count of the launch.  `scratch`: >= 4 bytes of device memory.                                  */
int smplb200_probe_fp32_fma(int32_t device, int32_t iters, void* scratch, double* flop, void* stream);

/* ---- per-kernel entry points (unit parity + ncu) ---------------------------------------- */
/* Internal intermediate layouts (owned by this library, stable within a version):
 *   coef   [N, 224] fp32   k<NB: betas; NB..NB+206: pose_feature (R[1:]-I); NB+207..209: 1.0
 *                          (v_template, split exactly over 3 rows for the tcgen05 paths); rest 0
 *   A      [N, J, 12] fp32 rows 0..2 of the rest-pose-removed joint transforms, row-major 3x4
 *   vposed [N, 3, VP] fp32 planar (x-plane, y-plane, z-plane), VP = V rounded up to 128
 */
int64_t smplb200_padded_verts(const SmplB200Model* model);          /* VP                    */
#define SMPLB200_COEF_K 224

/* k2: folded joint regression + Rodrigues + kinematic chain (one warp per body).
 * Writes coef, A, and optionally joints (kinematic J_posed).  Any output may be NULL.
 * (Inside smplb200_forward the same kernel also writes kp2d for kinematic joints.)          */
int smplb200_pose_chain(const SmplB200Model* model, const float* betas, const float* pose,
                        int64_t n, float* coef, float* A, float* joints,
                        uint32_t flags, void* stream);

/* k1: shape + pose blendshapes, coef[N,224] x basis -> vposed (planar).  `flags` selects the
 * FMA or a tcgen05 path via SMPLB200_PREC_*.  The tcgen05 paths need device scratch for the
 * bf16/tf32 operand images (`smplb200_blendshapes_workspace_bytes`, 256-byte aligned); the FMA
 * path needs none (workspace may be NULL).                                                   */
size_t smplb200_blendshapes_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags);
int smplb200_blendshapes(const SmplB200Model* model, const float* coef, int64_t n,
                         float* vposed, void* workspace, size_t workspace_bytes,
                         uint32_t flags, void* stream);

/* k3 (+k4): linear blend skinning with the weak-perspective projection in its epilogue.
 * `joints_in`/`cam`/`kp2d` may be NULL together (no projection).  SMPLB200_LBS_TC needs
 * `smplb200_lbs_workspace_bytes` of scratch for the tf32 hi|lo image of A.                   */
size_t smplb200_lbs_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags);
int smplb200_lbs(const SmplB200Model* model, const float* vposed, const float* A, int64_t n,
                 float* vertices, const float* joints_in, const float* cam, float* kp2d,
                 void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* k1 + k3 fused (SMPLB200_PREC_F16; csrc/k_fused_tc.cuh): coef[N,224] and A[N,J,12] (as written by
 * smplb200_pose_chain) -> vertices[N,V,3] in ONE tcgen05 kernel, accumulators in tensor memory, no v_posed
 * intermediate.  `workspace`: smplb200_blend_skin_workspace_bytes(model, n) of 256-byte aligned scratch for
 * the fp16 operand images (inside smplb200_forward k2 writes those images directly).  `coef` and `A` may
 * BOTH be NULL: the images a previous call left in `workspace` (same n) are reused and only the fused
 * kernel is launched (kernel-only timing).                                                                 */
size_t smplb200_blend_skin_workspace_bytes(const SmplB200Model* model, int64_t n);
int smplb200_blend_skin(const SmplB200Model* model, const float* coef, const float* A, int64_t n,
                        float* vertices, void* workspace, size_t workspace_bytes, void* stream);

/* joints regressed from skinned vertices (SMPLB200_JOINTS_REGRESSED) + optional projection. */
int smplb200_regress_joints(const SmplB200Model* model, const float* vertices, int64_t n,
                            float* joints, const float* cam, float* kp2d, void* stream);

/* ---- the backward pass (what the reference trainer needs next) -------------- */
/* (g_betas[N,NB], g_pose[N,3J], g_cam[N,3]) <- upstream gradients of the forward outputs
 * (g_vertices[N,V,3], g_joints[N,J,3], g_kp2d[N,J,2]; each may be NULL = zero), for the same
 * betas/pose/cam and `flags` as the forward call.  Replaces what autograd derives for the eager
 * layer inside the reference's training step (reference src/lib/trains/trainer.py:31-37 builds the
 * loss, :102-104 call loss.backward()).  Intermediates are recomputed, nothing is saved by the
 * forward.  `joints_fwd` (the forward's joints output) is read only with SMPLB200_JOINTS_REGRESSED
 * and g_kp2d.  `g_cam` may be NULL; it requires `cam`.  The vertex path (g_vertices given, or
 * regressed joints with g_joints/g_kp2d) needs `smplb200_backward_workspace_bytes(..., 1)` of
 * 256-byte aligned scratch; without it (`vertex_path` = 0) the workspace may be NULL.
 * `forward_workspace` (optional, may be NULL): the workspace of the smplb200_forward call being
 * differentiated (same n and flags, untouched since) -- its A and vposed intermediates are then
 * reused instead of recomputed.  Ignored when the forward ran in chunks (SMPLB200_CHUNK).
 * Gradients are summed in a fixed order (no atomics): bitwise reproducible.
 * `g_vertices` is staged with 16-byte bulk copies: a body's [V,3] slab is read from the enclosing
 * 16-byte granules, so up to 12 bytes on either side of it (inside the same allocation for any
 * cudaMalloc / caching-allocator buffer, whose sizes are multiples of 256 bytes) are touched.      */
size_t smplb200_backward_workspace_bytes(const SmplB200Model* model, int64_t n, uint32_t flags,
                                         int vertex_path);
int smplb200_backward(const SmplB200Model* model, const float* betas, const float* pose,
                      const float* cam, int64_t n, const float* joints_fwd,
                      const float* g_vertices, const float* g_joints, const float* g_kp2d,
                      float* g_betas, float* g_pose, float* g_cam,
                      const void* forward_workspace, size_t forward_workspace_bytes,
                      void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);
/* Kernel launches one backward call issues (bench accounting).                                 */
int smplb200_backward_launch_count(const SmplB200Model* model, int64_t n, uint32_t flags,
                                   int vertex_path, int reuse_forward_workspace);

/* ---- the producer of the per-person vectors ------------------------ */
/* Fused NMS + top-K + head gather, one launch:
 *   heat [B,C,H,W] (sigmoid-ed centre heat map), heads[h] [B,ch_h,H,W] (NCHW, e.g. pose72/shape10/cam3)
 *   -> scores[B,K], inds[B,K] (int64 index into H*W), clses[B,K] (int32), ys[B,K], xs[B,K],
 *      gathered[h] [B,K,ch_h]
 * Replaces: _nms + _topk (reference src/lib/models/decode.py:6-13, 26-41) followed by
 * _transpose_and_gather_feat per head (reference src/lib/models/utils.py:12-27), as called at
 * reference src/lib/models/decode.py:80-88.  `heads`, `head_channels`, `gathered` are HOST arrays of
 * `num_heads` (<= 8) entries holding DEVICE pointers; 1 <= k <= 256; bit-exact for tie-free scores. */
int smplb200_decode_gather(int32_t device, const float* heat, int32_t batch, int32_t num_classes,
                           int32_t height, int32_t width, const float* const* heads,
                           const int32_t* head_channels, int32_t num_heads, int32_t k,
                           float* scores, int64_t* inds, int32_t* clses, float* ys, float* xs,
                           float* const* gathered, void* stream);

/* ---- DCNv2 forward: the reference's one native op ------------------- */
/* output[B,Co,Ho,Wo] = modulated deformable convolution of input[B,Ci,H,W] with weight[Co,Ci,kh,kw],
 * offset[B,2*dg*kh*kw,Ho,Wo] (channel 2t = dh, 2t+1 = dw of tap t = i*kw + j), mask[B,dg*kh*kw,Ho,Wo]
 * and bias[Co] (may be NULL).  Same arguments, in the same order, as the reference binding
 * `dcn_v2_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h,
 * pad_w, dilation_h, dilation_w, deformable_group)` (reference src/lib/models/DCNv2/src/dcn_v2.h:9-39,
 * src/vision.cpp:4-9), which replaces modulated_deformable_im2col_cuda + two batched SGEMMs
 * (reference src/cuda/dcn_v2_cuda.cu:43-173).  One fused implicit-GEMM kernel, no column buffer.
 * Built for what the reference network uses (reference src/lib/models/model.py:355: 3x3, one
 * deformable group): kernel 3x3, deformable_group 1, Ci a multiple of 32, Co a multiple of 16 and
 * <= 256; any stride / padding / dilation.  Anything else returns SMPLB200_ERR_UNSUPPORTED.
 * `flags`: 0, or SMPLB200_DCN_INPUT_NHWC when `input` is already channels-last ([B,H,W,Ci], e.g. a
 * torch.channels_last tensor): the kernel samples a channels-last view (one contiguous 128-byte line
 * per neighbour and 32 channels) and otherwise makes that copy itself.
 * `workspace`: smplb200_dcn_v2_workspace_bytes(...) of 256-byte aligned device scratch (the re-tiled
 * bf16 weight image, rebuilt every call because weights are trainable, + the channels-last copy).   */
#define SMPLB200_DCN_INPUT_NHWC 1u
size_t smplb200_dcn_v2_workspace_bytes(int32_t batch, int32_t channels_in, int32_t height, int32_t width,
                                       int32_t channels_out, uint32_t flags);
int smplb200_dcn_v2_forward(int32_t device, const float* input, const float* weight, const float* bias,
                            const float* offset, const float* mask, int32_t batch, int32_t channels_in,
                            int32_t height, int32_t width, int32_t channels_out,
                            int32_t kernel_h, int32_t kernel_w, int32_t stride_h, int32_t stride_w,
                            int32_t pad_h, int32_t pad_w, int32_t dilation_h, int32_t dilation_w,
                            int32_t deformable_group, float* output,
                            void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* ---- misc ------------------------------------------------------------------------------- */
const char* smplb200_strerror(int status);
int smplb200_version(void);
/* Last CUDA error code seen by this thread inside the library (cudaError_t), 0 if none.      */
int smplb200_last_cuda_error(void);
/* Number of kernel launches the given forward configuration issues (for bench accounting).  */
int smplb200_forward_launch_count(const SmplB200Model* model, int64_t n, uint32_t flags,
                                  int with_projection);

#ifdef __cplusplus
}
#endif
#endif /* SMPL_B200_H */
