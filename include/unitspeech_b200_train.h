/* C ABI of the speaker-adaptation (fine-tune) step of libunitspeech_b200.so -- operator level.
 *
 * Stands in for what PyTorch autograd + torch.optim.Adam do underneath the reference's fine-tune loop
 *   decoder.zero_grad(); loss = decoder.fine_tune(...); loss.backward();
 *   clip_grad_norm_(decoder.parameters(), max_norm=1); optimizer.step()          finetune.py:131-165
 * for the objective UnitSpeech.loss_t (unitspeech/unitspeech.py:393-405) through GradLogPEstimator2d
 * (unitspeech/unitspeech.py:124-201).  The reference has no FFI here (the boundary is the Python class); the Python host
 * unitspeech_b200/training.py drives these entries in the order of the reference's forward graph and its reverse.
 *
 * Conventions (in addition to unitspeech_b200.h): every pointer is a DEVICE pointer; "h16" tensors are NHWC fp16
 * ((N, H, W, C), pixel p = y*W + x), everything else fp32; parameters and parameter gradients use the reference's
 * state_dict layouts, except conv weights (training layout, see usb_t_pack_conv); gradients ACCUMULATE into their outputs (+=); activation gradients and parameter gradients
 * carry the loss scale S (usb_t_loss_grad multiplies by it, usb_t_adam divides by it).  Every call only enqueues
 * work on `stream`: no synchronisation, no allocation.  Conv kinds: 0 = 3x3/s1, 1 = 3x3/s2, 2 = 1x1,
 * 3 = ConvTranspose 4x4/s2, and the data-gradient forms 4 = transposed 3x3/s2, 5 = 4x4/s2 over a ConvTranspose
 * output gradient.
 */
#ifndef UNITSPEECH_B200_TRAIN_H
#define UNITSPEECH_B200_TRAIN_H

#include "unitspeech_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fp32 master conv weight -> fp16 GEMM operands.  The master copy is kept in the TRAINING LAYOUT = the forward operand
 * layout (3x3: [Cout][9][Cin]; 1x1: [Cout][Cin]; ConvTranspose 4x4/s2: [4 phases][Cout][4 taps][Cin]); the Python host
 * converts from / to the reference's (Cout, Cin, k, k) / (Cin, Cout, 4, 4) at load_state_dict / state_dict.
 * fwd (kinds 0-3): plain cast; dgrad: the operand of the data-gradient convolution (kind 0 -> run as kind 0 with
 * flipped taps, 1 -> kind 4, 2 -> kind 2 transposed, 3 -> kind 5).  [ci0, ci1) selects a slice of the input channels
 * (skip-concat halves, unitspeech.py:192).  Either output may be NULL. */
int usb_t_pack_conv(usb_handle* h, int32_t kind, const float* w, int32_t Cout, int32_t Cin, int32_t ci0, int32_t ci1,
                    void* fwd, void* dgrad, uint64_t stream);

/* Between usb_t_pack_begin and usb_t_pack_flush, usb_t_pack_conv only records its dgrad part; flush runs all recorded
 * convs in ONE launch (the table is uploaded on first use -- a synchronising step -- and reused afterwards, so a captured
 * CUDA graph replays it). */
int usb_t_pack_begin(usb_handle* h, uint64_t stream);
int usb_t_pack_flush(usb_handle* h, uint64_t stream);

/* dst (fp16) = src (fp32), n a multiple of 4: refreshes the fp16 mirror of the flat master buffer -- with the conv
 * weights kept in the forward operand layout this IS the forward pack of every conv */
int usb_t_cast(usb_handle* h, const float* src, void* dst, int64_t n, uint64_t stream);

/* Conv2d / ConvTranspose2d and their data gradients on the tcgen05 implicit-GEMM kernels   unitspeech.py:49,21,30,66,83-84
 * in0/in1: h16 (N, H, W, C*tot) of which the first C0/C1 channels are contracted (in1: second K source or NULL);
 * w: packed fp16 weights (wZ matrices; b_batch_mode 0 shared, 1 per phase, 2 per sample); bias fp32 (Cout) or NULL;
 * mask fp32 (N, Wout) or NULL; res: h16 residual shaped like out (out = conv*res_scale + res; res_scale dev scalar or
 * NULL = 1); stats: int64 (N, groups, 2) GroupNorm partial sums accumulated in fixed point, or NULL. */
int usb_t_conv(usb_handle* h, int32_t kind, const void* in0, int32_t C0tot, int32_t C0, const void* in1, int32_t C1tot,
               int32_t C1, int32_t N, int32_t H, int32_t W, const void* w, int32_t wZ, int32_t b_batch_mode, int32_t Cout,
               const float* bias, const float* mask, const void* res, const float* res_scale, int64_t* stats, int32_t groups,
               void* out, uint64_t stream);

/* torch.stack([mu, x], 1) -> Block1 conv (2 -> C) + res_conv of downs.0.0                     unitspeech.py:170,49,66
 * x, mu: (N, H, W) fp32; rows: int32 (N) = 0..N-1; w3: (9, 2, C) tap-major fp32; w1: (2, C); raw/res: h16 */
int usb_t_first_conv(usb_handle* h, const float* x, const float* mu, const int32_t* rows, const float* mask, const float* w3,
                     const float* b3, const float* w1, const float* b1, void* raw, void* res, int64_t* stats, int32_t N,
                     int32_t H, int32_t W, int32_t C, uint64_t stream);

/* out = (Mish(GroupNorm(raw)) + addvec[n*addvec_stride + c] + res) * mask                     unitspeech.py:50-55,72-75 */
int usb_t_gn_apply(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta,
                   const float* addvec, int64_t addvec_stride, const void* res, const float* mask, void* out, int32_t N,
                   int32_t H, int32_t W, int32_t C, uint64_t stream);

/* SinusoidalPosEmb -> time MLP -> cat speaker -> Mish -> the stacked ResnetBlock.mlp Linears   unitspeech.py:109-121,165-168,61
 * u: (N, dim+S) = Mish(cat(time_mlp(t), spk)); e: (N, J) */
int usb_t_embed(usb_handle* h, const float* t, const float* spk, const float* freqs, const float* w0, const float* b0,
                const float* w2, const float* b2, const float* wcat, const float* bcat, float* u, float* e, int32_t N,
                int32_t J, uint64_t stream);

/* LinearAttention context (unitspeech.py:86-93) folded with to_out (:94-95): weff (N, C, heads*32) fp16;
 * ctx_out (N, heads, 32, 32) and stat_out (N, heads, 2, 32) = softmax max / normaliser are kept for the backward pass. */
int64_t usb_t_attn_scratch_bytes(int32_t N, int32_t heads, int32_t P);
int usb_t_attn_context(usb_handle* h, const void* qkv, int32_t ld, int32_t koff, int32_t voff, const float* wo, float* scratch,
                       void* weff, float* ctx_out, float* stat_out, int32_t N, int32_t P, int32_t C, int32_t heads,
                       uint64_t stream);

/* final_block GroupNorm + Mish -> final_conv 1x1 -> mask                                      unitspeech.py:198-201 */
int usb_t_final(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta, const float* wf,
                const float* bf, const float* mask, float* score, int32_t N, int32_t H, int32_t W, int32_t C, uint64_t stream);

/* loss = sum((score*sqrt(1-exp(-cum_noise)) + z*mask)^2) / (sum(mask)*n_feats)                 unitspeech.py:402-404
 * partial: 512 doubles of scratch; usb_t_loss_grad: dscore = S * dloss/dscore, msum: one float of scratch */
int usb_t_loss(usb_handle* h, const float* score, const float* zm, const float* mask, const float* t, double* partial,
               float* loss, int32_t B, int32_t T, uint64_t stream);
int usb_t_loss_grad(usb_handle* h, const float* score, const float* zm, const float* mask, const float* t, float loss_scale,
                    float* msum, float* dscore, int32_t B, int32_t T, uint64_t stream);

/* backward of y = (Mish(GroupNorm(raw)) + emb + res) * mask: d_raw (h16), dbias (conv bias), dgamma, dbeta, and the
 * embedding gradient d_emb[n*emb_stride + c] = sum_p dy.  dy = dy0, h16 (dy1 is reserved, pass NULL) -- or, for the final block, the outer
 * product dys[n][p] * wvec[c] (then d_wvec[c] += sum dys * y).  scratch: (3*N*C + N*16) floats. */
int usb_t_gn_bwd(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta, const void* dy0,
                 const void* dy1, const float* dys, const float* wvec, const float* mask, float* scratch, void* d_raw,
                 float* dbias, float* dgamma, float* dbeta, float* d_emb, int64_t emb_stride, float* d_wvec, int32_t N,
                 int32_t H, int32_t W, int32_t C, uint64_t stream);

/* out[n*out_stride_n + c] += sum_p t[n][p][c] (bias gradients);  out = a + b (+ c) on h16 tensors of n elements */
int usb_t_colsum(usb_handle* h, const void* t, int32_t ld, int32_t N, int32_t P, int32_t C, float* out, int64_t out_stride_n,
                 uint64_t stream);
int usb_t_add(usb_handle* h, const void* a, const void* b, const void* c, void* out, int64_t n, uint64_t stream);

/* weight gradient of a conv of kind 0-3 in the training layout (see usb_t_pack_conv).  dy: h16 output gradient (row stride ldy),
 * x: h16 layer input (N, H, W, ldx) whose channels [0, Cs) are the [ci0, ci0+Cs) slice of the Cin_total input channels.
 * per_sample bit 0 (kind 2): dW is (N, Cout, Cs), one matrix per sample; bit 1: the destination slice is known to hold
 * zeros (fresh zero_grad), which lets a launch with one CTA per tile store its result instead of adding it. */
int usb_t_wgrad(usb_handle* h, int32_t kind, const void* dy, int32_t ldy, const void* x, int32_t ldx, int32_t N, int32_t H,
                int32_t W, int32_t Cout, int32_t Cs, int32_t ci0, int32_t Cin_total, float* dW, int32_t per_sample,
                uint64_t stream);
int usb_t_first_conv_wgrad(usb_handle* h, const void* d_raw, const void* d_res0, const void* d_res1, const float* x,
                           const float* mu, const float* mask, float* dW3, float* dW1, int32_t N, int32_t H, int32_t W,
                           int32_t C, uint64_t stream);

/* Residual(Rezero(LinearAttention)) backward (unitspeech.py:36-43,78-106): from G[n] = sum_p d_out q^T (usb_t_wgrad
 * per-sample) and cs = per-sample column sums of d_out: gradients of to_out / g, dctx, and weffT = fp16(g * Weff^T),
 * the per-sample weight of the dq conv; then dk / dv from the saved context and softmax statistics. */
int usb_t_attn_bwd_small(usb_handle* h, const float* G, const float* cs, const float* wo, const float* bo, const float* g,
                         const float* ctx, float* dwo, float* dbo, float* dg, float* dctx, void* weffT, int32_t N, int32_t C,
                         int32_t heads, uint64_t stream);
int usb_t_attn_bwd_dkv(usb_handle* h, const void* qkv, int32_t ld, int32_t koff, int32_t voff, const float* ms,
                       const float* ctx, const float* dctx, void* dkv, int32_t N, int32_t P, int32_t heads, uint64_t stream);

/* backward of usb_t_embed; du: (N, dim+S) scratch */
int usb_t_embed_bwd(usb_handle* h, const float* t, const float* spk, const float* freqs, const float* w0, const float* b0,
                    const float* w2, const float* b2, const float* wcat, const float* u, const float* dE, float* du,
                    float* dw0, float* db0, float* dw2, float* db2, float* dwcat, float* dbcat, int32_t N, int32_t J,
                    uint64_t stream);

/* out[0] += sum a[i] * b[i] (b NULL: sum a) */
int usb_t_dot(usb_handle* h, const float* a, const float* b, int64_t n, float* out, uint64_t stream);

/* clip_grad_norm_(max_norm) + torch.optim.Adam step over flat fp32 buffers               finetune.py:81,163-165
 * sumsq[0] += sum g^2; usb_t_adam: g' = g * inv_scale * min(1, max_norm / (||g|| * inv_scale + 1e-6)); a non-finite
 * norm skips the step and increments *skipped.  step counts from 1; with step_dev != NULL the step number is read from
 * the device counter (*step_dev + 1) and the counter is advanced after the update, so a captured CUDA graph of a whole
 * training step can be replayed. */
int usb_t_sumsq(usb_handle* h, const float* g, int64_t n, double* out, uint64_t stream);
int usb_t_adam(usb_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
               float eps, int32_t step, int32_t* step_dev, const double* sumsq, float inv_scale, float max_norm,
               int32_t* skipped, uint64_t stream);

#ifdef __cplusplus
}
#endif
#endif
