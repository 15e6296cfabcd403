/*
 * segdistill.h - C ABI of libsegdistill_sm100.so (B200 / sm_100a only).
 *
 * Drop-in boundary for SegDistill's dense distillation-loss hot path.  The
 * reference has NO native interface for this path (0 native source files,
 * setup.py:125 ext_modules=[]): its loss modules call ATen directly.  Each entry
 * point below therefore replaces a *chain of ATen calls* in the reference and
 * cites it (paths relative to the reference root):
 *
 *   sd_kl_rows_fwd_bwd    mmseg/models/distillation/losses.py:35-42 (channel gather),
 *                         :50-58 (group reshape / -1e9 pad), :108-112 (softmax-KL, alpha)
 *                         - CDLoss :130-143, CGDLoss :145-158, CGDLossWS :160-173,
 *                         plain KLDLoss (softmax over the last dim) :9-113
 *   sd_kl_pixels_fwd_bwd  losses.py:47-49 (permute to NHWC rows) + :108-112 - PDLoss :115-128;
 *                         with at_weight != 0 also ATLoss :175-197 (channel-mean MSE :190
 *                         + per-pixel KL :192-195)
 *   sd_kl_rows_up_fwd_bwd the same behind KLDLoss.resize (losses.py:25-33,101-102; ops/wrappers.py:8-29):
 *                         bilinear up-sampling fused into the loss and its backward
 *   sd_mse_fwd_bwd        losses.py:178,190 / :202,235 (nn.MSELoss), :812-830 (feature MSE)
 *   sd_ifvd_sim_fwd_bwd   losses.py:218-235 - IFVDLoss: per-class centres (the loop over C classes :226-230),
 *                         nn.CosineSimilarity to the centre of the pixel's class, 10*nn.MSELoss
 *   sd_kl_rows_multi_fwd_bwd  the same for two `distillation` entries on one pair (opts.py:100-103)
 *   sd_scale_grad         the autograd multiply by grad_output that torch would run in
 *                         backward (loss enters the total as a plain sum,
 *                         mmseg/models/segmentors/SD_structure.py:121-122; 512 under fp16
 *                         loss scaling)
 *   sd_cgd_corr_fwd_bwd   NOT in the reference: builder-defined per-group
 *                         Gram-matrix loss; tcgen05 tensor-core GEMM.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named *_host.
 *   - the library never allocates, frees or synchronises: outputs and workspace are
 *     caller-owned; all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - feature maps are contiguous NCHW: element (b,c,p) at ((b*C + c)*HW + p).
 *   - return value: 0 = ok, negative = SD_ERR_* argument error, positive = cudaError_t.
 *     No C++ exception crosses the boundary.  sd_strerror() names any of them.
 *   - `workspace` must be zero-filled once when it is allocated.  Its first bytes are a counter
 *     arena common to every entry point (the kernels leave it zeroed again), so ONE workspace
 *     of the largest size requested may serve any sequence of calls and shapes on a stream;
 *     it must not be shared by concurrent streams.
 *   - fp32 accumulation regardless of `dtype`; dS has the dtype of S.
 *   - deterministic: no floating-point atomics; reductions run in a fixed order for a
 *     given shape and device.
 */
#ifndef SEGDISTILL_H_
#define SEGDISTILL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_ABI_VERSION 1

#if defined(__GNUC__)
#define SD_API __attribute__((visibility("default")))
#else
#define SD_API
#endif

/* dtype */
#define SD_F32  0
#define SD_BF16 1

/* algo: AUTO picks a TMA-staged kernel when the layout allows it (16-byte aligned rows): the
 * register-resident single pass for rows of up to 16384 elements (several whole rows per CTA pass when they
 * are short powers of two); for longer rows and for two fused
 * losses the grid-resident single pass (rows of up to 64 units spread over every SM, parked in tensor
 * memory), else the cluster-resident single pass (rows that fit 8 CTAs), else the
 * streaming two-phase kernel; the plain multi-pass kernel otherwise.  TMA = AUTO without the generic
 * fallback; STREAM / CLUSTER / GRID force that kernel (rows only; CLUSTER and GRID answer
 * SD_ERR_UNSUPPORTED when the rows do not fit). */
#define SD_ALGO_AUTO    0
#define SD_ALGO_GENERIC 1
#define SD_ALGO_TMA     2
#define SD_ALGO_STREAM  3
#define SD_ALGO_CLUSTER 4
#define SD_ALGO_ROWS1   5   /* one row per CTA pass even where several short rows could be packed (tests) */
#define SD_ALGO_GRID    6   /* grid-resident single pass: long rows spread over all SMs (cooperative launch) */
#define SD_ALGO_WARP    7   /* sd_kl_pixels_fwd_bwd: one warp per pixel column (kl_pixels_warp_kernel; AUTO takes it for bf16) */

/* argument errors */
#define SD_OK               0
#define SD_ERR_NULL        -1   /* a required pointer is NULL */
#define SD_ERR_SHAPE       -2   /* non-positive or overflowing shape */
#define SD_ERR_DTYPE       -3
#define SD_ERR_ALIGN       -4   /* a base pointer is not 16-byte aligned */
#define SD_ERR_WORKSPACE   -5   /* workspace too small */
#define SD_ERR_UNSUPPORTED -6   /* forced algo cannot run this layout */
#define SD_ERR_DEVICE      -7   /* not an sm_100 device */
#define SD_ERR_VALUE       -8   /* tau <= 0, group < 1, ... */

SD_API int         sd_abi_version(void);
SD_API const char* sd_strerror(int rc);
/* SD_OK when the current device can run the kernels (compute capability 10.x). */
SD_API int         sd_device_check(void);

/* ------------------------------------------------------------------ rows (CD / CGD) */
SD_API size_t sd_kl_rows_workspace_bytes(int B, int C, int HW, int group);

/*
 * Softmax-KL over rows made of `group` consecutive channels (after the optional
 * channel gather `chan_perm`), forward and backward fused.
 *   rows R = B*ceil(C/group); a ragged last group counts as a row whose missing
 *   channels contribute nothing (the reference pads them with -1e9).
 *   row_kl[r] = sum_i p_i (log p_i - log q_i),  p = softmax(T_row/tau), q = softmax(S_row/tau)
 *   *loss     = alpha/R * sum_r row_kl[r]
 *   dS        = grad_scale*alpha/(R*tau) * (q - p), written in the ORIGINAL channel order
 *   mse_weight != 0 additionally adds the feature MSE on the same pair:
 *   *mse_loss = mse_weight*mean((S-T)^2), dS += grad_scale*2*mse_weight*(S-T)/numel
 * chan_perm: int32[C] or NULL; row_kl: float[R] or NULL; mse_loss: NULL iff mse_weight == 0.
 */
SD_API int sd_kl_rows_fwd_bwd(const void* S, const void* T, void* dS,
                       float* row_kl, float* loss,
                       const int32_t* chan_perm,
                       int B, int C, int HW, int group, int dtype,
                       float tau, float alpha, float grad_scale,
                       float mse_weight, float* mse_loss,
                       void* workspace, size_t workspace_bytes,
                       int algo, void* stream);

/*
 * Up to two softmax-KL losses over the SAME (S, T) pair in one pass (dispatcher-level batching of
 * mmseg/models/distillation/opts.py:100-103 when two `distillation` entries hook the same tensors,
 * e.g. CD + CGD on the logits): S and T are read once and the SUM of the gradients is written once.
 *   groups/taus/alphas: host arrays [n_losses]; every row of the loss with the larger group must be
 *   a union of whole rows of the other (groups[1] % groups[0] == 0, or groups[1] >= C).
 *   losses[k]: device float[1]; row_kls: NULL or host array of device float[R_k] (entries may be NULL).
 *   dS = grad_scale * sum_k g_k * alpha_k/(R_k*tau_k) * (q_k - p_k), with g_k = *grad_outputs[k]
 *   (device scalars) or 1 when grad_outputs (or the entry) is NULL.
 *   run_if: NULL, or a device word: the launch is a no-op when it reads 0 (conditional backward
 *   re-run after sd_scale_grad2 found non-uniform upstream gradients).
 * algo: SD_ALGO_AUTO (grid- or cluster-resident kernel when the rows fit, else the streaming kernel), or
 * SD_ALGO_GRID / SD_ALGO_CLUSTER / SD_ALGO_STREAM to force one.  TMA paths only: SD_ERR_UNSUPPORTED when the layout
 * cannot take it (call the single-loss entry per loss).
 */
SD_API int sd_kl_rows_multi_fwd_bwd(const void* S, const void* T, void* dS, int n_losses, const int* groups,
                             const float* taus, const float* alphas, float* const* losses,
                             float* const* row_kls, const float* const* grad_outputs,
                             const unsigned* run_if, int B, int C, int HW, int dtype, float grad_scale,
                             void* workspace, size_t workspace_bytes, int algo, void* stream);

/*
 * The channel-mode losses of SEVERAL (student, teacher) pairs in ONE launch: one dispatcher step over a `distillation`
 * list with more than one layer (mmseg/models/distillation/opts.py:87-112; the reference runs its whole op chain once per
 * entry).  Pair k: maps S[k], T[k] of shape (B[k], C[k], HW[k]), rows of groups[k] channels, temperature taus[k], weight
 * alphas[k]; *losses[k] and dS[k] exactly as sd_kl_rows_fwd_bwd computes them for that pair alone (same dtype for all
 * pairs; n_pairs <= 8).  grad_output: NULL or one device scalar multiplied into every gradient.  The units of all
 * pairs form one work list for a persistent cooperative grid: the grid-resident single pass (units parked in tensor
 * memory) when every map has HW % 128 == 0 and no row spans more than 64 units of 16384 elements, else the two-phase
 * streaming kernel (rows of any length); ragged groups either way.  SD_ERR_UNSUPPORTED when a pair's rows are not
 * 16-byte aligned (use the per-pair entry).
 */
SD_API size_t sd_kl_rows_group_workspace_bytes(int n_pairs, const int* B, const int* C, const int* HW, const int* groups,
                                        int dtype);
SD_API int sd_kl_rows_group_fwd_bwd(int n_pairs, const void* const* S, const void* const* T, void* const* dS,
                             float* const* losses, const int* B, const int* C, const int* HW, const int* groups,
                             const float* taus, const float* alphas, int dtype, float grad_scale,
                             const float* grad_output, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ pixels (PD / AT) */
SD_API size_t sd_kl_pixels_workspace_bytes(int B, int C, int HW);

/*
 * Softmax-KL over the channel axis of every pixel (rows R = B*HW, length C, stride HW).
 *   row_kl[b*HW + p], *loss = alpha/R * sum, dS = grad_scale*alpha/(R*tau)*(q - p).
 *   at_weight != 0 adds ATLoss' attention term: with m = mean over channels,
 *   *at_loss = at_weight*mean_{b,p}((m_S - m_T)^2),
 *   dS += grad_scale*2*at_weight*(m_S - m_T)/(C*B*HW).
 */
SD_API int sd_kl_pixels_fwd_bwd(const void* S, const void* T, void* dS,
                         float* row_kl, float* loss,
                         int B, int C, int HW, int dtype,
                         float tau, float alpha, float grad_scale,
                         float at_weight, float* at_loss,
                         void* workspace, size_t workspace_bytes,
                         int algo, void* stream);

/* ------------------------------------------------------------------ rows on up-sampled maps */
SD_API size_t sd_kl_rows_up_workspace_bytes(int B, int C, int Hl, int Wl, int group);
/*
 * The channel-mode loss of sd_kl_rows_fwd_bwd on maps that the reference first resizes to the label
 * size (KLDLoss.resize, mmseg/models/distillation/losses.py:25-33,101-102: F.interpolate(mode='bilinear',
 * align_corners=False) through mmseg/ops/wrappers.py:8-29) - without materialising the resized maps:
 * S, T and dS are the LOW-resolution [B, C, Hl, Wl] maps, rows are `group` channels x (scale*Hl) x (scale*Wl)
 * up-sampled values, dS is the gradient with respect to the low-resolution S (the backward of the
 * interpolation included).  scale in {2, 4, 8}; anything else: SD_ERR_UNSUPPORTED (resize on the host, then
 * sd_kl_rows_fwd_bwd).  row_kl: [B*ceil(C/group)] or null; chan_perm as in sd_kl_rows_fwd_bwd.
 */
SD_API int sd_kl_rows_up_fwd_bwd(const void* S, const void* T, void* dS, float* row_kl, float* loss,
                          const int32_t* chan_perm, int B, int C, int Hl, int Wl, int scale, int group, int dtype,
                          float tau, float alpha, float grad_scale,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * The pixel-mode loss of sd_kl_pixels_fwd_bwd (PDLoss, losses.py:115-128) behind the same resize: softmax over
 * the C channels of every UP-SAMPLED pixel, rows R = B*(scale*Hl)*(scale*Wl); S, T, dS at low resolution.
 * scale in {2, 4, 8}; anything else: SD_ERR_UNSUPPORTED.
 */
SD_API size_t sd_kl_pixels_up_workspace_bytes(int B, int C, int Hl, int Wl, int scale);
SD_API int sd_kl_pixels_up_fwd_bwd(const void* S, const void* T, void* dS, float* loss,
                            int B, int C, int Hl, int Wl, int scale, int dtype,
                            float tau, float alpha, float grad_scale,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ feature MSE */
SD_API size_t sd_mse_workspace_bytes(int64_t numel);
/* *loss = weight*mean((S-T)^2); dS = grad_scale*2*weight*(S-T)/numel. */
SD_API int sd_mse_fwd_bwd(const void* S, const void* T, void* dS, float* loss,
                   int64_t numel, int dtype, float weight, float grad_scale,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ student head: resize + cross-entropy + accuracy */
/*
 * BaseDecodeHead.losses, mmseg/models/decode_heads/decode_head.py:217-237, with the default CrossEntropyLoss
 * (mmseg/models/losses/cross_entropy_loss.py:9-32, :138-198; weight_reduce_loss, losses/utils.py:25-56) and
 * accuracy (mmseg/models/losses/accuracy.py:4-46, top-1), plus autograd's backward, in one pass:
 *   x = bilinear_resize(logits, scale, align_corners=False)            (never materialised; scale 1, 2, 4 or 8)
 *   nll(b,y,x) = -log_softmax_c(x)[label] * class_weight[label] * pixel_weight     (0 where label == ignore_index)
 *   *loss = loss_weight * sum(nll) / denominator
 *           denominator: number of label pixels B*Hs*Ws for reduction='mean' (ignored pixels count, as in the
 *           reference's loss.mean()), avg_factor when given, 1 for reduction='sum'
 *   *acc  = 100 * #{argmax_c x == label} / (B*Hs*Ws)                  (may be NULL)
 *   dlogits = grad_scale * d loss / d logits, low resolution, dtype of logits.
 * logits (B, C, Hl, Wl) contiguous; label (B, scale*Hl, scale*Wl) int64; class_weight float[C] or NULL; pixel_weight
 * float (B, scale*Hl, scale*Wl) or NULL.  A label outside [0, C) that is not ignore_index contributes nothing (the
 * reference's F.cross_entropy asserts on the device).  Deterministic.
 */
SD_API size_t sd_ce_up_workspace_bytes(int B, int C, int Hl, int Wl, int scale);
SD_API int sd_ce_up_fwd_bwd(const void* logits, const int64_t* label, void* dlogits, float* loss, float* acc,
                     const float* class_weight, const float* pixel_weight, int B, int C, int Hl, int Wl, int scale,
                     int dtype, int64_t ignore_index, float loss_weight, double denominator, float grad_scale,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ IFVDLoss similarity term */
/*
 * mmseg/models/distillation/losses.py:218-235 (IFVDLoss without its per-pixel KL, which is sd_kl_pixels_fwd_bwd):
 *   centre_X[:, k] = sum_{p: cls[p] == k} X[:, p] / (n_k + 1e-6)   per sample, X in {S, T}
 *   sim_X(p) = cosine_similarity(X[:, p], centre_X[:, cls[p]])      (ATen semantics, eps = 1e-8)
 *   *loss = weight * mean_{b,p} (sim_S - sim_T)^2;  dS = grad_scale * d loss / d S, including the path through the
 *   class centres.  cls: int32 (B, HW), class index in [0, C) or C for pixels without a class (they keep their own
 *   feature as centre: sim = 1, zero gradient).  Deterministic (no float atomics).
 *   accumulate != 0: dS += the gradient (dS already holds the gradient of the per-pixel KL term of the same loss,
 *   written by sd_kl_pixels_fwd_bwd with the same grad_scale).
 */
SD_API size_t sd_ifvd_sim_workspace_bytes(int B, int C, int HW);
/* largest channel (= class) count the class-sum kernels take: their per-warp bins (C + 1 classes x 32 channel lanes)
 * must fit one CTA's shared memory.  Larger C -> SD_ERR_UNSUPPORTED (the reference's Python loop, :222-230, takes any). */
SD_API int sd_ifvd_max_channels(void);
SD_API int sd_ifvd_sim_fwd_bwd(const void* S, const void* T, const int32_t* cls, void* dS, float* loss,
                        int B, int C, int HW, int dtype, float weight, float grad_scale, int accumulate,
                        void* workspace, size_t workspace_bytes, void* stream);
/*
 * losses.py:218-224: cls[b, y, x] = the label at the nearest-resized position (nn.Upsample(size=(h, w),
 * mode='nearest') of target (B, 1, Ht, Wt), int64) if it is one of the class indices 0..C-1, else C.
 */
SD_API int sd_ifvd_class_map(const int64_t* target, int32_t* cls, int B, int Ht, int Wt, int h, int w, int C,
                      void* stream);

/* ------------------------------------------------------------------ backward helper */
/* dS *= *grad_output (a device scalar); exits without touching dS when it equals 1. */
SD_API int sd_scale_grad(void* dS, int64_t numel, int dtype, const float* grad_output, void* stream);
/* The same launch also appends the step's n loss scalars `values` (device) to the log ring - what sd_log_push does
 * with a launch of its own: ring[(*cursor mod slots)][0..n) = values, *cursor += 1.  The append does not depend on
 * grad_output.  Replaces the per-variable all_reduce + .item() of SD_structure.py:137-142 together with sd_log_push
 * (segdistill_b200/dist.py: DeferredLogs). */
SD_API int sd_scale_grad_log(void* dS, int64_t numel, int dtype, const float* grad_output, const float* values, int n,
                             float* ring, unsigned* cursor, int slots, void* stream);
/* The same for the n_tensors <= 8 gradients of a grouped launch (sd_kl_rows_group_fwd_bwd; opts.py:87-112 sums the entries'
 * losses, so every layer's gradient meets its own upstream factor in backward): dS[k] *= *grad_outputs[k], one launch. */
SD_API int sd_scale_grad_group(int n_tensors, void* const* dS, const int64_t* numel, int dtype,
                               const float* const* grad_outputs, void* stream);

/* Two upstream gradients for one fused two-loss dS: if *grad_output0 == *grad_output1, dS *= that value
 * and *nonuniform_flag = 0; else dS is left alone and *nonuniform_flag = 1 (see run_if above). */
SD_API int sd_scale_grad2(void* dS, int64_t numel, int dtype, const float* grad_output0, const float* grad_output1,
                   unsigned* nonuniform_flag, void* stream);

/* ------------------------------------------------------------------ loss scalars for logging */
/*
 * mmseg/models/segmentors/SD_structure.py:137-142 all-reduces every log variable on every iteration (and blocks on
 * .item()), although the logger reads them every 50 (local_configs/_base_/default_runtime.py:2-7).  Here the step's
 * n scalars (device floats, e.g. the `loss` outputs of the calls above) are appended to a device-resident ring:
 *   ring[(*cursor mod slots) * n + i] = values[i];  *cursor += 1
 * - one 32-thread launch that a CUDA graph can capture (the slot is picked on the device) - and the host all-reduces
 * the whole ring once per `slots` steps (segdistill_b200/dist.py: DeferredLogs).  ring: slots * n floats.
 */
SD_API int sd_log_push(const float* values, int n, float* ring, unsigned* cursor, int slots, void* stream);

/* ------------------------------------------------------------------ CGD correlation (extension) */
SD_API size_t sd_cgd_corr_workspace_bytes(int B, int C, int HW, int group);
/*
 * Builder-defined (no reference counterpart): per (sample, channel group) Gram matrix
 * G = X X^T / HW with X in R^{group x HW};  *loss = alpha*mean((G_S - G_T)^2) over
 * B*ceil(C/group)*group^2 entries;  dS = grad_scale*4*alpha/(N*HW) * (G_S - G_T) X_S.
 * tcgen05 (TMEM accumulators) fed by TMA; bf16 inputs use kind::f16, fp32 inputs kind::tf32.
 */
SD_API int sd_cgd_corr_fwd_bwd(const void* S, const void* T, void* dS, float* loss,
                        int B, int C, int HW, int group, int dtype,
                        float alpha, float grad_scale,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ introspection */
/* Number of kernel launches the library has enqueued since load (bench.py's gpu_launches). */
SD_API uint64_t sd_launch_count(void);
/* Name of the kernel variant the last call on this thread dispatched to ("" if none). */
SD_API const char* sd_last_kernel(void);

#ifdef __cplusplus
}
#endif
#endif /* SEGDISTILL_H_ */
