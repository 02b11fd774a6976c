/* ecgb200.h -- C ABI of libecgb200.so: sm_100a kernels for the PTB-XL 1D-CNN ECG
 * train / infer / Grad-CAM step.
 *
 * The reference (cyu0330/ptbxl-multimodal) has no FFI of its own: every FLOP of
 * its hot path is a PyTorch ATen op reached from src/models/ecg_cnn.py,
 * src/models/ecg_multimodal.py, src/training/loop*.py and
 * src/interpretability/grad_cam_1d.py.  Each entry point below names the
 * reference call site (path:line under /root/reference) whose ATen op it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; nothing is
 *     allocated, freed or synchronised inside the library;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is enqueued on it, so calls are CUDA-graph capturable;
 *   - return value 0 = ok, <0 = ECGB200_E* argument error, >0 = cudaError_t of
 *     the launch; nothing throws across the boundary;
 *   - fp32 tensors use the reference's layouts: activations (B, C, L)
 *     contiguous, Conv1d weight (Co, Ci, 15), Linear weight (out, in);
 *   - `ws` arguments are caller-provided scratch; the matching *_ws_bytes()
 *     query returns the size needed for a given shape.
 */
#ifndef ECGB200_H_
#define ECGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECGB200_KSIZE 15          /* Conv1d kernel_size, ecg_cnn.py:10 */
#define ECGB200_EINVAL (-1)       /* bad shape / null pointer            */
#define ECGB200_EUNSUPPORTED (-2) /* shape outside the kernels' envelope */

/* library / build identification; also the cheapest "is it loaded" probe */
int ecgb200_version(void);
/* compiled-for architecture as an integer (1000 = sm_100a) */
int ecgb200_arch(void);

/* ---------------------------------------------------------------- Conv1d --
 * Weight re-layout done once per optimizer step.
 *   w      (Co, Ci, 15)  reference layout (nn.Conv1d.weight, ecg_cnn.py:13)
 *   w_fwd  (Ci, 15, Co)  operand of ecgb200_conv1d_fwd_f32
 *   w_dgr  (Co, 15, Ci)  tap-flipped transpose: operand of the same kernel when
 *                        it computes the input gradient (dgrad).  May be NULL. */
int ecgb200_conv1d_prep_weights_f32(const float* w, float* w_fwd, float* w_dgr,
                                    int Co, int Ci, void* stream);

/* y[b,o,t] = bias[o] + sum_{c,k} w[o,c,k] * x[b,c,t+k-7]   (zero padded)
 * replaces aten::convolution called from nn.Conv1d, src/models/ecg_cnn.py:13.
 *   x (B,Ci,L)  wt (Ci,15,Co) from prep_weights  bias (Co) or NULL  y (B,Co,L)
 *   stat_part: NULL, or float[2*Co*ntiles] receiving per-(channel, tile)
 *              {sum, centred M2} of y for the train-mode BatchNorm that follows
 *              (ntiles = ecgb200_conv1d_stat_tiles(B, L)).
 * Any Co works; Co % 4 == 0 takes the vectorised weight path. */
int ecgb200_conv1d_fwd_f32(const float* x, const float* wt, const float* bias, float* y,
                           float* stat_part, int B, int Ci, int Co, int L, void* stream);
int ecgb200_conv1d_stat_tiles(int B, int L);

/* dW[o,c,k] = sum_{b,t} dy[b,o,t] * x[b,c,t+k-7];  db[o] = sum_{b,t} dy[b,o,t]
 * replaces the wgrad / bias-grad half of aten::convolution_backward (autograd of
 * ecg_cnn.py:13 reached from loss.backward(), src/training/loop.py:33).
 *   ws: ecgb200_conv1d_wgrad_ws_bytes(B,Ci,Co,L) bytes of scratch (split-K partials,
 *       reduced in a fixed order => run-to-run deterministic). */
int ecgb200_conv1d_wgrad_f32(const float* dy, const float* x, float* dw, float* db, void* ws,
                             int B, int Ci, int Co, int L, void* stream);
size_t ecgb200_conv1d_wgrad_ws_bytes(int B, int Ci, int Co, int L);

/* ------------------------------------------------ bf16 tensor-core path (tcgen05) --
 * Activations in this path are "blocked channels-last" bf16:  A[b][c/8][t][c%8], C % 8 == 0
 * (16 bytes = 8 channels of one time step), accumulation is fp32 in TMEM.
 * pack:   x fp32 (B,Ci,T) -> xb bf16 [B][Cp/8][T][8], Cp = Ci rounded up to 16, pad channels 0
 *         (the (B,12,T) input of ECGCNN.forward, ecg_cnn.py:52).
 * unpack: blocked bf16 -> fp32 (B,C,L)  (hooks / parity checks). */
int ecgb200_pack_input_bf16(const float* x, void* xb, int B, int Ci, int T, void* stream);
int ecgb200_unpack_act_bf16(const void* xb, float* x, int B, int C, int L, void* stream);
/* w fp32 (Co,Ci,15) -> wf bf16 [15][Cip/8][Co][8] (forward operand) and, unless NULL,
 * wd bf16 [15][Co/8][Cip][8] (tap-flipped transpose: dgrad operand); Cip = Ci rounded up to 16. */
int ecgb200_conv1d_prep_weights_bf16(const float* w, void* wf, void* wd, int Co, int Ci, void* stream);
/* Conv1d(k=15,pad=7) as implicit GEMM on tcgen05: yb = conv(xb, wprep) + bias.
 * Same call computes dgrad with (dy, wd, NULL).  Ci % 16 == 0 (Ci % 64 == 0 above 64), Co % 32 == 0, both <= 256.
 * replaces aten::convolution (ecg_cnn.py:13) in bf16 mode. */
int ecgb200_conv1d_fwd_bf16(const void* xb, const void* wprep, const float* bias, void* yb,
                            int B, int Ci, int Co, int L, void* stream);

/* Same conv, persistent one-CTA-per-SM kernel, that also emits the train-mode BatchNorm statistics of
 * its (bf16-rounded) outputs: stat_part float[parts][2][Co] = per-CTA {sum, sum of squares},
 * parts = ecgb200_conv1d_stat_parts_bf16(B,Ci,Co,L) (<= number of SMs).  stat_part may be NULL.
 * ecgb200_conv1d_fwd_bf16 is this call with stat_part == NULL. */
int ecgb200_conv1d_fwd_stats_bf16(const void* xb, const void* wprep, const float* bias, void* yb,
                                  float* stat_part, int B, int Ci, int Co, int L, void* stream);
int ecgb200_conv1d_stat_parts_bf16(int B, int Ci, int Co, int L);
/* Train-mode BatchNorm1d + ReLU + MaxPool1d(2) (+GAP) in ONE pass over yb: finalises the statistics from
 * the conv partials (mean, biased var -> bn_state {mean,rstd,scale,shift}; running stats with momentum and
 * the unbiased variance; *num_batches_tracked += 1), then applies them.  ecg_cnn.py:14-16,46,62.
 * nrep = number of replicas whose B*L samples stat_part covers: 1, or the world size after
 * ecgb200_dp_bn_sync_f32 (SyncBN: statistics over the global batch, as the single-device reference). */
int ecgb200_bn_relu_pool_fwd_train_bf16(const void* yb, const float* stat_part, int nparts, const float* gamma,
                                        const float* beta, float* running_mean, float* running_var,
                                        int64_t* num_batches_tracked, float* bn_state, void* pb, float* gap,
                                        int B, int C, int L, float momentum, float eps, int nrep, void* stream);

/* Last block (gap != NULL; its pooled output only feeds AdaptiveAvgPool1d, ecg_cnn.py:46,62): the same call, which also leaves
 * the routing summary route[2][B][C] fp32 -- per (window, channel) the number of pool pairs with a ReLU-positive maximum and the
 * sum of the raw conv outputs at those (first-index-wins) positions.  Every pooled position of (b, c) receives the same gradient
 * dgap[b][c] / (L/2), so the two batch reductions of the BatchNorm backward are sums of dgap * route over B x C:
 * ecgb200_bn_relu_pool_bwd_route_bf16 then needs ONE pass over y (to write dy) instead of two.  Same results as
 * ecgb200_bn_relu_pool_bwd_bf16 up to fp32 summation order. */
int ecgb200_bn_relu_pool_fwd_train_route_bf16(const void* yb, const float* stat_part, int nparts, const float* gamma,
                                              const float* beta, float* running_mean, float* running_var,
                                              int64_t* num_batches_tracked, float* bn_state, void* pb, float* gap,
                                              float* route, int B, int C, int L, float momentum, float eps, int nrep,
                                              void* stream);
int ecgb200_bn_relu_pool_bwd_route_bf16(const void* yb, const float* bn_state, const float* dgap, const float* route,
                                        void* dyb, float* dgamma, float* dbeta, float* db_part, int B, int C, int L,
                                        int train, void* stream);

/* dW (Co,Ci,15) fp32 and db (Co) fp32 from blocked-bf16 dy [B][Co/8][L][8] and x [B][Cip/8][L][8]
 * (Cip = Ci rounded up to 16) on tcgen05, accumulators resident in TMEM across the whole batch
 * share of a CTA; split-K partials in ws (ecgb200_conv1d_wgrad_bf16_ws_bytes) are reduced in a
 * fixed order.  db = sum over db_part[Co][ndb] (per-sample sums of dy written by
 * ecgb200_bn_relu_pool_bwd_bf16) or 0 when db_part == NULL.  Co % 8 == 0, Co <= 256. */
int ecgb200_conv1d_wgrad_bf16(const void* dyb, const void* xb, float* dw, float* db, const float* db_part,
                              int ndb, void* ws, int B, int Ci, int Co, int L, void* stream);
size_t ecgb200_conv1d_wgrad_bf16_ws_bytes(int B, int Ci, int Co, int L);
/* Blocked-bf16 versions of the BatchNorm + ReLU + MaxPool (+GAP) block (ecg_cnn.py:14-16,46,62);
 * statistics, bn_state, gap, dgap, dgamma, dbeta stay fp32.  ws: ecgb200_bn_bwd_ws_bytes(B,C). */
int ecgb200_bn_train_stats_bf16(const void* yb, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, int64_t* num_batches_tracked, float* bn_state, void* ws,
                                int B, int C, int L, float momentum, float eps, void* stream);
int ecgb200_bn_relu_pool_fwd_bf16(const void* yb, const float* bn_state, void* pb, float* gap,
                                  int B, int C, int L, void* stream);
int ecgb200_bn_relu_pool_bwd_bf16(const void* yb, const float* bn_state, const void* dpb, const float* dgap,
                                  void* dyb, float* dgamma, float* dbeta, float* db_part, void* ws,
                                  int B, int C, int L, int train, void* stream);
/* The two passes of that backward as separate calls (SyncBN under data parallel): pass 1 leaves this replica's
 * partials part[ecgb200_bn_nsplit(B,C)][2][C] = {sum g, sum g*a}; after ecgb200_dp_bn_sync_f32 pass 2 takes the
 * exchanged list (one pair per replica): batch means over all nrep replicas' B*L samples, dgamma / dbeta from pair
 * local_idx (this replica's own) only -- the gradient exchange sums them.  local_idx < 0: affine gradients from the
 * merged sums (single device).  Autograd of torch.nn.SyncBatchNorm semantics for ecg_cnn.py:14-16. */
int ecgb200_bn_relu_pool_bwd_reduce_bf16(const void* yb, const float* bn_state, const void* dpb,
                                         const float* dgap, float* part, int B, int C, int L, void* stream);
int ecgb200_bn_relu_pool_bwd_apply_bf16(const void* yb, const float* bn_state, const void* dpb,
                                        const float* dgap, const float* part, int nparts, int local_idx,
                                        int nrep, void* dyb, float* dgamma, float* dbeta, float* db_part,
                                        int B, int C, int L, int train, void* stream);
/* number of per-channel partials the bf16 BN kernels produce (second dim of db_part) */
int ecgb200_bn_nsplit(int B, int C);

/* ------------------------------------------- BatchNorm1d + ReLU + MaxPool1d --
 * Train-mode statistics from the conv epilogue partials (or from y itself when
 * stat_part == NULL): mean, biased var -> rstd; scale = gamma*rstd,
 * shift = beta - mean*scale; running stats updated with momentum 0.1 and the
 * UNBIASED variance; *num_batches_tracked += 1.   nn.BatchNorm1d, ecg_cnn.py:14.
 *   bn_state: float[4*Co] = {mean, rstd, scale, shift} (saved for backward).
 *   ws: scratch of ecgb200_bn_stats_ws_bytes(B, Co, L) bytes (used when stat_part==NULL). */
int ecgb200_bn_train_stats_f32(const float* y, const float* stat_part, const float* gamma,
                               const float* beta, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, float* bn_state, void* ws,
                               int B, int Co, int L, float momentum, float eps, void* stream);
size_t ecgb200_bn_stats_ws_bytes(int B, int Co, int L);
/* Eval mode: bn_state from the running statistics (no update). */
int ecgb200_bn_eval_state_f32(const float* gamma, const float* beta, const float* running_mean,
                              const float* running_var, float* bn_state, int Co, float eps,
                              void* stream);

/* p[b,c,j] = max_{i in {2j,2j+1}} relu(y[b,c,i]*scale[c] + shift[c]),  Lp = floor(L/2)
 * (BatchNorm apply + ReLU(inplace) + MaxPool1d(2), ecg_cnn.py:14-16).
 * If p == NULL only the global average pool is produced:
 * gap[b,c] = mean_j p[b,c,j]   (AdaptiveAvgPool1d(1)+squeeze, ecg_cnn.py:46,62). */
int ecgb200_bn_relu_pool_fwd_f32(const float* y, const float* bn_state, float* p, float* gap,
                                 int B, int Co, int L, void* stream);

/* Backward of the block above through train-mode (batch-statistics) BatchNorm.
 *   dp (B,Co,Lp) gradient w.r.t. the pooled output, or NULL with dgap (B,Co) given
 *   (then dp[b,c,j] = dgap[b,c]/Lp).  Produces dy (B,Co,L), dgamma, dbeta.
 * With train == 0 the BN is affine with fixed statistics (eval mode; Grad-CAM):
 *   dy = scale * g.   ws: ecgb200_bn_bwd_ws_bytes(B,Co) bytes. */
int ecgb200_bn_relu_pool_bwd_f32(const float* y, const float* bn_state, const float* gamma,
                                 const float* dp, const float* dgap, float* dy, float* dgamma,
                                 float* dbeta, void* ws, int B, int Co, int L, int train,
                                 void* stream);
size_t ecgb200_bn_bwd_ws_bytes(int B, int Co);

/* ----------------------------------------------------------------- Linear --
 * y = act(x W^T + b);  x (M,K), W (N,K), b (N) or NULL, y (M,N); act: 0 none, 1 relu.
 * nn.Linear at ecg_cnn.py:47,50 and ecg_multimodal.py:52-54,85,86. */
int ecgb200_linear_fwd_f32(const float* x, const float* w, const float* b, float* y,
                           int M, int K, int N, int act, void* stream);
/* dx = dy W (M,K) [NULL to skip]; dw = dy^T x (N,K); db = colsum(dy) (N) [NULL to skip].
 * If relu_out != NULL, dy is first masked by relu_out > 0 IN PLACE (ReLU(inplace) bwd). */
int ecgb200_linear_bwd_f32(const float* x, const float* w, float* dy, const float* relu_out,
                           float* dx, float* dw, float* db, int M, int K, int N, void* stream);

/* FiLM: zc = (1 + tanh(film[:, :F])) * z + film[:, F:]   (ecg_multimodal.py:92-96) */
int ecgb200_film_fwd_f32(const float* z, const float* film, float* zc, int B, int F, void* stream);
/* dz = dzc*(1+tanh g);  dfilm[:, :F] = dzc*z*(1-tanh^2 g);  dfilm[:, F:] = dzc */
int ecgb200_film_bwd_f32(const float* z, const float* film, const float* dzc, float* dz,
                         float* dfilm, int B, int F, void* stream);

/* loss = mean_{b,c} [max(x,0) - x*y + log1p(exp(-|x|))]; dlogits = (sigmoid(x)-y)*gscale/(n)
 * F.binary_cross_entropy_with_logits, src/training/loop.py:32; loop_demo.py:10,33.
 * prob (sigmoid, loop.py:63) and dlogits may be NULL.  n = B*C elements. */
int ecgb200_bce_logits_f32(const float* logits, const float* target, float* loss,
                           float* dlogits, float* prob, int n, float gscale, void* stream);

/* Evaluation epilogue on the device (src/training/loop.py:63, scripts/06_ecg_baseline_test.py:127,
 * src/training/metrics.py:37): prob = sigmoid(logits), pred = prob >= threshold (on the fp32 probability, as the
 * reference does), counts[c] += {tp, fp, fn, tn} (int32[C][4], caller zeroes it at the start of an epoch).
 * target / prob / pred / counts may each be NULL.  logits, target, prob: (rows, C) fp32; pred: (rows, C) uint8. */
int ecgb200_eval_counts_f32(const float* logits, const float* target, float* prob, unsigned char* pred, int* counts,
                            int rows, int C, float threshold, void* stream);

/* ------------------------------------------------------------------ AdamW --
 * torch.optim.AdamW.step (scripts/03_train_ecg_baseline.py:133, loop.py:34):
 *   p *= 1-lr*wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *   p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps);   g is first scaled by gscale.
 * One launch over the `ntensors` tensors whose device pointers are listed in the HOST arrays
 * p/g/m/v/numel (copied into the kernel's parameter block).  Hyper-parameters and the step
 * counter live on the DEVICE so that the call can be captured in a CUDA graph and replayed:
 *   hyper    float[6] = {lr, beta1, beta2, eps, weight_decay, gscale}
 *   step_ctr int[1]   = steps taken so far; t = step_ctr+1 is used, then step_ctr += 1. */
int ecgb200_adamw_f32(int ntensors, float* const* p, const float* const* g, float* const* m,
                      float* const* v, const int64_t* numel, const float* hyper, int* step_ctr,
                      void* stream);

/* ------------------------------------------------ fused per-step kernels (TrainStep engine) --
 * Step prologue in one launch: pack x (B,Ci0,T) fp32 -> xb blocked bf16 (ecgb200_pack_input_bf16), re-lay the
 * nlayers <= 4 conv weights (ecgb200_conv1d_prep_weights_bf16), transpose the proj weight wp (F,Cin) ->
 * wpT (Cin,F) [wp may be NULL], and *step_ctr += 1 [may be NULL].  w/wf/wd/co/ci are HOST arrays.
 * x may be NULL (weights only): the engine packs the input + block-1 weights on the critical path and
 * re-lays the other weights beside the first conv. */
int ecgb200_step_prep_bf16(const float* x, void* xb, int B, int Ci0, int T, int nlayers,
                           const float* const* w, void* const* wf, void* const* wd, const int* co,
                           const int* ci, const float* wp, float* wpT, int F, int Cin, int* step_ctr,
                           void* stream);
/* ECGCNN head, forward + loss + input gradients in one launch (ecg_cnn.py:63-64, loop.py:32-33):
 *   z = gap Wp^T + bp; logits = z Wh^T + bh; BCE terms; dlogits = (sigmoid-y)*gscale/(B*NL);
 *   dz = dlogits Wh; dgap = dz Wp.   loss_part[ecgb200_head_loss_parts(B)] = partial sums of the BCE terms.
 * Cin, F <= 256, NL <= 8. */
int ecgb200_head_fwd_bwd_f32(const float* gap, const float* wpT, const float* wp, const float* bp,
                             const float* wh, const float* bh, const float* target, float* z, float* logits,
                             float* dlogits, float* dz, float* dgap, float* loss_part, int B, int Cin, int F,
                             int NL, float gscale, void* stream);
int ecgb200_head_loss_parts(int B);
/* ... and its weight gradients + the scalar loss (only the optimizer needs them; runs beside the conv
 * backward): dWp = dz^T gap, dbp, dWh = dlogits^T z, dbh, loss = mean BCE. */
int ecgb200_head_wgrad_f32(const float* gap, const float* z, const float* dz, const float* dlogits,
                           const float* loss_part, float* dwp, float* dbp, float* dwh, float* dbh, float* loss,
                           int B, int Cin, int F, int NL, void* stream);
/* ECGMultimodal head (FiLM, ecg_multimodal.py:88-99), forward + loss + input gradients in one launch:
 *   h1 = relu(W0 d + b0); h2 = relu(W2 h1 + b2); film = Wf h2 + bf; z = Wp gap + bp;
 *   zc = (1 + tanh film[:F]) z + film[F:]; logits = Wh zc + bh; BCE; and the chain rule back to dgap, dh1.
 * Saves for the weight gradients: z, h1, h2, film, zc, dlogits, dz, dfilm, dh2, dh1.  D <= 8, H <= 64, F, Cin <= 256. */
int ecgb200_mm_head_fwd_bwd_f32(const float* gap, const float* demo, const float* wpT, const float* wp,
                                const float* bp, const float* w0, const float* b0, const float* w2, const float* b2,
                                const float* wf, const float* bf, const float* wh, const float* bh,
                                const float* target, float* z, float* h1, float* h2, float* film, float* zc,
                                float* logits, float* dlogits, float* dz, float* dfilm, float* dh2, float* dh1,
                                float* dgap, float* loss_part, int B, int Cin, int F, int D, int H, int NL,
                                float gscale, void* stream);
/* Weight / bias gradients of up to 6 small Linear layers in one launch: dW_p = A_p^T X_p (N_p x K_p), db_p = colsum
 * (A_p) [db[p] may be NULL]; a/x/dw/db/n/k are HOST arrays.  Also loss = sum(loss_part)/(B*NL) unless NULL. */
int ecgb200_head_wgrad_multi_f32(int nprob, const float* const* a, const float* const* x, float* const* dw,
                                 float* const* db, const int* n, const int* k, const float* loss_part, float* loss,
                                 int B, int NL, void* stream);
/* AdamW (as ecgb200_adamw_f32) over one flat, 16-byte aligned array; t = *step_now, the 1-based step index
 * already incremented by ecgb200_step_prep_bf16. */
int ecgb200_adamw_flat_f32(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                           const int* step_now, void* stream);

/* ------------------------------------------------ data parallel: gradient exchange fused with AdamW --
 * One launch per rank over NVLink peer memory: barrier -> this rank's 1/world shard of the flat parameter space
 * gets g = sum_r G_r (peer loads, fixed order) * hyper[5] -> AdamW (as ecgb200_adamw_flat_f32, moments sharded) ->
 * the new parameters are stored into EVERY rank's buffer (peer stores) -> barrier.  Replaces NCCL all-reduce +
 * a replicated optimizer step.  p / g / flags are HOST arrays of `world` (<= 8) peer-mapped device pointers;
 * flags: ecgb200_dp_flag_words(world) zero-initialised uint32 per rank; n % (4*world) == 0; *step_now = 1-based
 * optimizer step (bias correction), identical on all ranks; all ranks must make the same sequence of calls. */
int ecgb200_dp_adamw_fused_f32(float* const* p, const float* const* g, unsigned int* const* flags, float* m,
                               float* v, int64_t n, int rank, int world, const float* hyper, const int* step_now,
                               void* stream);
int ecgb200_dp_flag_words(int world);
/* The same for ONE BUCKET [off, off + n) of the flat space (off % 4 == 0, n % (4*world) == 0; rank r owns the r-th
 * 1/world of the bucket), so that the block-4 parameters -- 65 % of the bytes, final as soon as wgrad_4 is -- are
 * exchanged and stepped while blocks 3..1 still run backward.  One flag pad per bucket. */
int ecgb200_dp_adamw_fused_range_f32(float* const* p, const float* const* g, unsigned int* const* flags, float* m,
                                     float* v, int64_t off, int64_t n, int rank, int world, const float* hyper,
                                     const int* step_now, void* stream);
/* One-hop ("LL") form of the bucket exchange: every 4-byte payload travels with its flag in one 8-byte word
 * {value, epoch}, so there are no barriers and no fences -- every rank pushes its gradients into the shard owners'
 * inboxes, the owners sum in rank order, apply AdamW and push the new parameters into every rank's inbox, every rank
 * unpacks its inbox into its replica.  Same arithmetic and summation order, i.e. the same bits, as
 * ecgb200_dp_adamw_fused_range_f32, at about half its latency (one NVLink hop per phase instead of fence + flag + poll).
 * p, g, m, v: this rank's flat buffers (need not be peer mapped); inbox: HOST array of `world` peer-mapped pointers to
 * every rank's inbox of ecgb200_dp_ll_inbox_words(n) zero-initialised 8-byte words, one inbox per bucket; ctr: 2 zeroed
 * uint32 of this rank per bucket (epoch, blocks finished). */
int ecgb200_dp_adamw_ll_f32(float* p, const float* g, float* m, float* v, void* const* inbox, unsigned int* ctr,
                            int64_t off, int64_t n, int rank, int world, const float* hyper, const int* step_now,
                            void* stream);
size_t ecgb200_dp_ll_inbox_words(int64_t n);
/* SyncBN exchange in the same one-hop form: inbox = HOST array of peer-mapped pointers to every rank's
 * [2][world][512] 8-byte words for THIS (block, direction) slot; ctr: 1 zeroed uint32 of this rank per slot. */
int ecgb200_dp_bn_sync_ll_f32(const float* local_part, int nparts, int C, void* const* inbox, unsigned int* ctr,
                              float* out, int rank, int world, void* stream);
/* SyncBN statistics exchange over NVLink peer memory (one tiny launch per BatchNorm pass): reduces this replica's
 * local_part[nparts][2][C] to one pair, publishes it in slots[rank] (>= 2*C floats of peer-mapped memory), passes a
 * cross-rank barrier and gathers all replicas' pairs in rank order into out[world][2][C] (device memory of this
 * rank), which the BatchNorm kernels merge like per-CTA partials (nparts = nrep = world).  A slot must not be reused
 * before a later cross-rank barrier (the optimizer exchange); flags: one pad shared by all BatchNorm exchanges.
 * Gives the reference's single-device batch statistics (ecg_cnn.py:14) under data parallel.  C <= 256. */
int ecgb200_dp_bn_sync_f32(const float* local_part, int nparts, int C, float* const* slots,
                           unsigned int* const* flags, float* out, int rank, int world, void* stream);

/* Programmatic dependent launch for the kernels of the bf16 step's critical path (conv, BatchNorm forward): when on,
 * they are launched with cudaLaunchAttributeProgrammaticStreamSerialization and overlap their prologue with the
 * previous kernel's tail (each waits for the previous kernel's completion before touching global memory).
 * Returns the previous setting.  Off by default. */
int ecgb200_set_pdl(int on);

/* Debug only: when buf != NULL, CTA 0 of the bf16 conv kernel writes clock64() stamps of its pipeline
 * events into buf[0..63] (device memory).  NULL switches tracing off (the default). */
int ecgb200_debug_set_trace(long long* buf);
/* A/B switch for the CTA-pair (tcgen05 cta_group::2, M = 256 per instruction) kernels of the wide layers -- blocks 3 and 4 of
 * src/models/ecg_cnn.py:29-33.  Bit 0: forward and dgrad (>= 64 input channels, weights larger than 64 KB; each SM stages half of
 * the weight slab; bit-identical to the one-SM kernel).  Bit 1: wgrad (128 / 256 output channels; taps as M, each SM stages half
 * of the dY tile).  Bit 2: also for small problems where the pair form does not pay (one tile per SM; few items per split) --
 * for tests.  Default 3; 0 selects the one-SM kernels.  Do not flip it between ecgb200_conv1d_stat_parts_bf16 /
 * ecgb200_conv1d_wgrad_bf16_ws_bytes and the launches those calls size. */
int ecgb200_debug_set_conv_pair(int mask);
/* Debug only: when buf != NULL every CTA of the tcgen05 conv / wgrad kernels writes %globaltimer (ns) at its first
 * and after its last instruction to buf[2 * linear block id + {0, 1}] (launch ramp, spread and tail of a grid). */
int ecgb200_debug_set_cta_span(unsigned long long* buf);
/* Debug only: pinned (device-visible) host buffer of >= 4 words; a barrier wait that exceeds its deadlock
 * limit records {0xDEAD, block<<32|thread, smem address<<32|parity} there before the kernel traps. */
int ecgb200_debug_set_diag(unsigned long long* pinned_host);
/* Limits of the bounded spin waits, in milliseconds, 0 = wait for ever: `mbarrier_ms` for the producer / issuer /
 * epilogue barriers inside the tcgen05 kernels (default 30 s), `peer_ms` for the cross-rank flag waits of
 * ecgb200_dp_adamw_fused_f32 / ecgb200_dp_bn_sync_f32 (default 10 min: ranks may drift apart by a validation pass or
 * a checkpoint write).  A wait that exceeds its limit traps the kernel instead of hanging the GPU. */
int ecgb200_set_spin_timeout_ms(unsigned int mbarrier_ms, unsigned int peer_ms);
/* Debug only: one-thread kernel that writes %globaltimer (ns) to buf[idx] in stream order -- placed between the
 * nodes of a captured step it gives the schedule the graph really runs (TrainStep.trace_schedule). */
int ecgb200_debug_stamp(unsigned long long* buf, int idx, void* stream);

/* --------------------------------------------------------------- Grad-CAM --
 * All-class batched Grad-CAM from the raw 4th-conv output A (B,C,L') in eval mode,
 * closed form of GradCAM1D.generate_cam (src/interpretability/grad_cam_1d.py:53-103):
 *   w[n,k,ch] = v[n or 0,k,ch] * s[ch] * count[n,ch] / (Lp*L'),  cam = relu(sum_ch w A)
 *   variant 1: (cam-min)/max at L' then linear upsample to T (grad_cam_1d.py:92-101)
 *   variant 2: upsample then (cam-min)/(max+eps)  (scripts/00_demo_inference.py:39-61,
 *              12_grad_cam_ecg_demo.py:44-75, 13_grad_cam_af.py:51-76)
 *   A (B,C,L'), bn_state of block 4 (eval), v (K,C) shared (v_per_sample=0) or (B,K,C)
 *   cam_lo (B,K,L') may be NULL; cam_hi (B,K,T) may be NULL (T==0); argmax (B,K) int32 of
 *   the final map (first maximal index), may be NULL.
 * With bn_state == NULL, v holds the final channel weights mean_t dY/dA themselves
 * (grad_cam_1d.py:85) as captured by backward hooks, and no mask count is applied. */
int ecgb200_gradcam_f32(const float* A, const float* bn_state, const float* v, int v_per_sample,
                        float* cam_lo, float* cam_hi, int32_t* argmax, int B, int C, int Lq,
                        int K, int T, int variant, float eps, void* stream);

/* out[r] = mean_t x[r,t]   (weights = dYdA.mean(dim=2), grad_cam_1d.py:85) */
int ecgb200_row_mean_f32(const float* x, float* out, int rows, int L, void* stream);

/* ------------------------------------------------------- input pipeline (N1) --
 * per-lead z-score (x-mean)/(std+1e-6), population std, over time
 * (src/datasets/ptbxl.py:122-127).  x, out (B, C, T). */
int ecgb200_zscore_f32(const float* x, float* out, int rows, int T, void* stream);

/* WFDB format-16 decode + transpose + per-lead z-score from the raw .dat frames (N2):
 * out[b,l,t] = z-score over t of (dat[b,t,l] - baseline[l]) / gain[l]   (digital -32768 = NaN, as wfdb.rdsamp);
 * replaces wfdb.rdsamp + np.asarray(float32) + .T + _normalize (src/datasets/ptbxl.py:25-29,36-50,122-127).
 * dat (B,T,n_leads) int16 little-endian frames; normalize = 0 returns the physical signal.  n_leads <= 16. */
int ecgb200_wfdb16_zscore_f32(const void* dat, const float* gain, const int* baseline, float* out, int B,
                              int n_leads, int T, int normalize, void* stream);
/* The same decode + z-score written straight to the first conv's input: blocked channels-last bf16
 * xb [B][Cp/8][T][8], Cp = n_leads rounded up to 16 (padding leads zero) -- N1 fused into the input pack; the fp32
 * (B, leads, T) tensor never exists.  T * n_leads must be even. */
int ecgb200_wfdb16_zscore_pack_bf16(const void* dat, const float* gain, const int* baseline, void* xb, int B,
                                    int n_leads, int T, void* stream);

/* ------------------------------------------------- bf16 inference engine (eval) --
 * model.eval() forward of the reference (src/training/loop.py:52-65, loop_demo.py:59-75,
 * scripts/06_ecg_baseline_test.py:94-106) with nothing but the pooled activations leaving the SM.
 *
 * Eval-mode BatchNorm1d (ecg_cnn.py:14) and the conv bias folded to per-channel fp32
 *   scale = gamma / sqrt(running_var + eps),  shift = (conv_bias - running_mean) * scale + beta. */
int ecgb200_bn_fold_f32(const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, const float* conv_bias, float* scale, float* shift,
                        int C, float eps, void* stream);
/* One ConvBlock in eval mode (ecg_cnn.py:18-20): pb = maxpool2(relu(scale * conv(xb, wprep) + shift)) in the
 * epilogue of the tcgen05 conv kernel; xb / wprep as ecgb200_conv1d_fwd_bf16, pb [B][Co/8][L/2][8] bf16.
 * gap_part != NULL (last block): pb is not written (may be NULL); gap_part[B * ceil(L/128)][4][Co] receives the
 * per-(128-step tile, lane quarter) time sums of the pooled rows for AdaptiveAvgPool1d(1) (ecg_cnn.py:61-62). */
int ecgb200_conv1d_bn_relu_pool_infer_bf16(const void* xb, const void* wprep, const float* scale,
                                           const float* shift, void* pb, float* gap_part, int B, int Ci,
                                           int Co, int L, void* stream);
/* Split-precision ("fp32 on the tensor cores") forms of the same block: every fp32 value x travels as two bf16 planes
 * hi = bf16(x), lo = bf16(x - hi); the products hi*hi + lo*hi + hi*lo are computed as THREE TIMES THE INPUT CHANNELS of the
 * same tcgen05 implicit GEMM (activations [x_hi | x_lo | x_hi] against weights [w_hi | w_hi | w_lo], fp32 accumulation in
 * TMEM), and the epilogue splits the pooled fp32 result again for the next block.  Only lo*lo (<= 2^-18 relative) is
 * dropped: logits agree with the fp32 reference path to ~1e-5 (tests/test_gpu_infer.py) at ~1/3 of the bf16 engine's rate.
 * ecgb200_split_channels(C) = channel count of a split tensor (3 planes of C rounded up to 16, zero-padded to 64 / a
 * multiple of 64).  Replaces Conv1d + BatchNorm1d(eval) + ReLU + MaxPool1d (ecg_cnn.py:13-16) for fp32-exact callers. */
int ecgb200_split_channels(int Ci);
int ecgb200_pack_input_split_bf16(const float* x, void* xb, int B, int Ci, int T, void* stream);
int ecgb200_conv1d_prep_weights_split_bf16(const float* w, void* wf, int Co, int Ci, void* stream);
int ecgb200_conv1d_bn_relu_pool_infer_split_bf16(const void* xb, const void* wprep, const float* scale,
                                                 const float* shift, void* pb, float* gap_part, int B, int Ci,
                                                 int Co, int L, void* stream);
/* y (B, Co, L) fp32 = conv(xb) + bias from split-plane inputs: the RAW 4th-conv output the reference's Grad-CAM forward
 * hook captures (src/interpretability/grad_cam_1d.py:36-43), at fp32 accuracy on the tensor cores. */
int ecgb200_conv1d_fwd_split_f32(const void* xb, const void* wprep, const float* bias, float* y, int B, int Ci, int Co,
                                 int L, void* stream);
/* Fused inference head: gap = inv_lp * sum of a window's `nparts` partials; z = proj(gap) (wpT = proj.weight
 * transposed, (C4, F)); with demo != NULL the DemoEncoder -> film_gen -> FiLM chain of
 * src/models/ecg_multimodal.py:44-59,88-99 (w1 (H,D0); w2 and wf TRANSPOSED: (H_in,H_out) and (H,2F)); logits = head(.) (ecg_cnn.py:63-64);
 * prob = sigmoid(logits) (loop.py:63).  z (B,F) (un-modulated features) and prob may be NULL.
 * C4, F <= 256, H <= 64. */
int ecgb200_infer_head_f32(const float* gap_part, int nparts, float inv_lp, const float* wpT,
                           const float* bp, const float* demo, const float* w1, const float* b1,
                           const float* w2, const float* b2, const float* wf, const float* bf,
                           const float* wh, const float* bh, float* z, float* logits, float* prob,
                           int B, int C4, int F, int D0, int H, int NL, void* stream);
/* out (cols, rows) = in (rows, cols) transposed (proj.weight -> wpT, once per weight refresh). */
int ecgb200_transpose_f32(const float* in, float* out, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECGB200_H_ */
