// Launchers of the streaming (HBM-bound) kernels of the reverse-diffusion path.  All tensors are device pointers;
// activations are NHWC fp16 ([row n][pixel p = y*W + x][channel]), statistics are [n][group][sum, sumsq] in int64 fixed point.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "conv_igemm.h"

namespace usb {

// ---- input conv (K = 18) + 1x1 res_conv of downs.0.0 on the 2-channel stack [mu, x]  (unitspeech.py:170,49,66)
struct FirstConvParams {
    const float* x;           // [B][H][W] current x_t
    const float* cond;        // [B][H][W] conditioning mel
    const float* text_uncon;  // [H] (used by rows with mu_row < 0) or null
    const int* x_row;         // [N] utterance index of x for row n
    const int* mu_row;        // [N] utterance index of cond for row n, or -1 -> text_uncon
    const float* mask;        // [N][W]
    const float* w3;          // [9][2][C] fp32 (tap-major repack of block1 conv weight)
    const float* b3;          // [C]
    const float* w1;          // [2][C] res_conv weight repack
    const float* b1;          // [C]
    __half* raw;              // [N][P][C] conv + bias (GroupNorm input)
    __half* res;              // [N][P][C] res_conv + bias
    long long* stats;         // [N][groups][2] fixed point (see conv_igemm.h)
    int N, H, W, C, groups;
};
int launch_first_conv(const FirstConvParams& p, cudaStream_t s);

// ---- out = (Mish(GroupNorm(raw)) + addvec[n][c] + res) * mask   (unitspeech.py:50-55,72-75)
struct GnApplyParams {
    const __half* raw;     // [N][P][C]
    const long long* stats;  // [N][groups][2] fixed point
    const float* gamma;    // [C]
    const float* beta;     // [C]
    const float* addvec;   // [N][addvec_stride] (+ offset already applied) or null
    long long addvec_stride;
    const __half* res;     // [N][P][C] or null
    const float* mask;     // [N][W]
    __half* out;           // [N][P][C]
    int N, P, W, C, groups;
    float eps;
    int dbg;               // experiments: 1 = skip Mish, 2 = skip the store
    unsigned long long* sat = nullptr;   // fp16 saturation report (see ConvParams::sat) or null
};
int launch_gn_apply(const GnApplyParams& p, int num_sms, cudaStream_t s);

// ---- final_block GN+Mish -> final_conv 1x1 -> CFG combine -> posterior update  (unitspeech.py:198-201,322-324,366-370)
struct FinalParams {
    const __half* raw;     // [nb*B][P][C] final_block conv output
    const long long* stats;  // [nb*B][groups][2] fixed point
    const float* gamma;    // [C]
    const float* beta;     // [C]
    const float* wf;       // [C] final_conv weight
    const float* bf;       // [1] final_conv bias (device)
    const float* mask;     // [B][W] (rows 0..B-1 of the level-0 mask)
    int B, nb, P, W, C, groups;
    float eps;
    float a0, a1;          // score = s_full + a0*(s_full - s_0) + a1*(s_full - s_1)   (full = last branch)
    // sampler (null xt -> estimator mode: only `score` rows are written)
    float* xt;             // [B][P] in/out, updated in place
    const float* noise;    // [B][P] or null (treated as 0)
    float c_x, c_s, sigma;
    float* score;          // estimator mode: [nb*B][P] per-row outputs; sampler mode: optional [B][P] combined score
    // last sampler step: also write the updated x_t to the caller's buffer, optionally de-normalised to log-mel
    // (inference.py:140: (y + 1) / 2 * (mel_max - mel_min) + mel_min with per-bin mel_min/mel_max)
    float* out;            // [B][P] or null
    const float* mel_min;  // [H] device or null
    const float* mel_max;  // [H]
    // graph-replayed sampler step: the per-step scalars come from a device table indexed by a device step counter, so ONE
    // captured launch sequence serves every step (null step_ctr: the host-provided scalars above are used)
    const int* step_ctr;          // current step i
    const float* step_tab;        // [n][4] = c_x, c_s, sigma, (1 if last step else 0); a0 / a1 stay kernel arguments
    long long noise_step_stride;  // elements between the noise slices of consecutive steps (noise then is the base)
};
int launch_final(const FinalParams& p, int num_sms, cudaStream_t s);

// ---- time / speaker embedding  (unitspeech.py:109-121,133-134,165-168,61)
struct EmbedParams {
    const float* t;        // [N]
    const float* spk;      // [N][S]
    const float* freqs;    // [dim/2] sinusoidal frequencies (host-computed with the reference's torch ops)
    const float* w0;       // [4*dim][dim]
    const float* b0;       // [4*dim]
    const float* w2;       // [dim][4*dim]
    const float* b2;       // [dim]
    const float* wcat;     // [J][dim+S] all ResnetBlock.mlp Linear weights stacked
    const float* bcat;     // [J]
    float* u;              // [N][dim+S] scratch: Mish(cat(time_mlp(t), spk))
    float* e;              // [N][J] output
    int N, dim, S, J;
    float pe_scale;
};
int launch_embed(const EmbedParams& p, cudaStream_t s);   // t or spk may be null (that half of u is 0); bcat may be null
// step_ctr != null: t_part is the base of the [n][J] per-step table and row *step_ctr is used
// also zeroes `zero_count` 64-bit words at `zero` (the statistics slots of the evaluation that follows: no memset node
// between the kernels of a step)
int launch_emb_combine(const float* t_part, const float* s_part, float* e, int N, int J, const int* step_ctr,
                       unsigned long long* zero, long long zero_count, cudaStream_t s);
// *ctr += 1 (first node of the captured sampler step)
int launch_step_advance(int* ctr, cudaStream_t s);

// ---- LinearAttention context: softmax over positions of k, ctx = k_sm^T v, folded with to_out  (unitspeech.py:86-96)
struct AttnParams {
    const __half* qkv;     // [N][P][ld] fp16; k channels at koff + head*dh, v channels at voff + head*dh
    int ld, koff, voff;
    const float* wo;       // [C][hidden] to_out weight
    float* part;           // scratch of attn_scratch_bytes(): [N][heads][chunks][dh*dh + 2*dh] partials + merged ctx
    __half* weff;          // plain mode: [N][C][hidden] folded per-sample weight (K-major B operand of the to_out GEMM)
    // fused-q mode (wq != null): the whole block collapses to one per-sample 1x1 conv on x,
    //   attn(x)*g + x = (g * Weff[n] * Wq + I) x + g*b_o;  weff then is [N][C][C] and bprime [C]
    const float* wq;       // [hidden][C] query rows of to_qkv, or null
    const float* g;        // Rezero scalar (device)
    const float* bo;       // [C] to_out bias
    float* bprime;         // [C] g * b_o
    // training path: keep the merged context and the softmax statistics for the backward pass (null otherwise)
    float* ctx_out;        // [N][heads][32][32] merged, normalised context (default: behind the partials in `part`)
    float* stat_out;       // [N][heads][2][32]: per channel max (natural-log domain) and normaliser
    int N, P, C, heads, chunk;   // dh = 32, hidden = heads*32
};
int attn_chunks(int P, int chunk);
size_t attn_scratch_bytes(int N, int heads, int P, int chunk);   // partials + merged context
int launch_attn_context(const AttnParams& p, cudaStream_t s);

// ---- small utilities
int launch_downsample_mask(const float* src, float* dst, int N, int Wsrc, int Wdst, cudaStream_t s);
int launch_gather_rows(const float* src, const int* idx, float* dst, int N, int len, cudaStream_t s);

// ---- training objective, forward value (unitspeech/unitspeech.py:376-405); zm / mu may be null
int launch_forward_diffusion(const float* x0, const float* z, const float* cond, const float* mask, const float* t,
                             float beta_min, float beta_max, float* xt, float* zm, float* mu, int B, int F, int T,
                             cudaStream_t s);
// partial: 512 doubles of scratch; loss: device scalar
int launch_diffusion_loss(const float* score, const float* zm, const float* mask, const float* t, float beta_min,
                          float beta_max, double* partial, float* loss, int B, int F, int T, cudaStream_t s);

}  // namespace usb
