// Launchers of the backward / optimizer kernels of the speaker-adaptation (fine-tune) step:
// UnitSpeech.fine_tune -> compute_loss -> loss_t (unitspeech/unitspeech.py:393-411,452-492) followed by
// loss.backward(), clip_grad_norm_(max_norm=1) and Adam (finetune.py:81,159-165).
//
// Conventions: activations and activation gradients are NHWC fp16 ([row n][pixel p = y*W + x][channel]);
// activation gradients carry the loss scale S (the objective is multiplied by S before differentiation so the
// fp16 gradients stay in range); parameter gradients are fp32, in the training layout of launch_pack_conv (conv weights)
// or the reference's layout (everything else), also scaled by S -- the optimizer kernel divides by S.  All parameter-gradient outputs ACCUMULATE (+=, atomics).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace usb {

// ---- GroupNorm + Mish backward (Block, unitspeech.py:46-55; ResnetBlock :70-75)
//   forward: y = (Mish(GN(raw)) + emb[n][c] [+ res]) * m.   dy = (dy0 [+ dy1]) * m  or  dy = dys[n][p] * wvec[c] * m.
//   reduce: sums[0][n][c] = sum_p dg, sums[1][n][c] = sum_p dg * xhat, dg = dy * Mish'(GN(raw)),
//           sums[2][n][c] = sum_p dy (embedding gradient)   or, in the scalar-dy mode, sum_p dys*m * Mish(GN(raw))*m
//   apply:  d_raw = rstd * (gamma * dg - (S1 + xhat * S2) / (cpg * P));  dbias[c] += sum_{n,p} d_raw
struct GnBwdParams {
    const __half* raw;        // [N][P][C] conv output (GroupNorm input)
    const long long* stats;   // [N][groups][2] fixed point (conv_igemm.h)
    const float* gamma;       // [C]
    const float* beta;        // [C]
    const __half* dy0;        // [N][P][C] or null (scalar-dy mode)
    const __half* dy1;        // must be null (reserved: a second contribution is summed by the caller with add_h)
    const float* dys;         // scalar-dy mode: [N][P] fp32
    const float* wvec;        // scalar-dy mode: [C]
    const float* mask;        // [N][W]
    float* sums;              // [3][N][C] fp32 (zeroed by the caller before `reduce`)
    __half* d_raw;            // [N][P][C] (apply)
    float* dbias;             // [C] += (apply), or null
    // emitted by `apply` from the sums: dgamma[c] += sum_n sums1, dbeta[c] += sum_n sums0,
    // d_emb[n*emb_stride + c] = sums2[n][c], d_wvec[c] += sum_n sums2[n][c]  (any may be null)
    float *dgamma, *dbeta, *d_emb, *d_wvec;
    long long emb_stride;
    int N, P, W, C, groups;
    float eps;
};
int launch_gn_bwd_reduce(const GnBwdParams& p, int num_sms, cudaStream_t s);
int launch_gn_bwd_apply(const GnBwdParams& p, int num_sms, cudaStream_t s);

// ---- out[c] (+ n*out_stride_n) += sum_p t[n][p][c]  (bias gradients; per-sample sums when out_stride_n != 0)
int launch_colsum(const __half* t, int ld, int N, int P, int C, float* out, long long out_stride_n, int num_sms,
                  cudaStream_t s);
// ---- out = a + b (+ c), fp16, n8 = element count / 8
int launch_add_h(const __half* a, const __half* b, const __half* c, __half* out, long long n, cudaStream_t s);

// ---- weight gradient of a convolution: for tap t
//   dW[co*s_co + ci*s_ci + tap_off[t] + n*s_n] += sum_{(n, y, x) in the iteration image}
//       A[n][y*a_mul + ady[t]][x*a_mul + adx[t]][co] * B[n][y*b_mul + bdy[t]][x*b_mul + bdx[t]][ci]
// (out-of-range pixels contribute zero = conv padding).  A = output gradient, B = layer input, both NHWC fp16 with
// row strides lda / ldb.  s_n != 0: one gradient matrix per sample (LinearAttention per-sample weights).
constexpr int kWgradMaxTaps = 16;
struct WgradParams {
    const __half* A;   // [N][Ha][Wa][lda]
    const __half* B;   // [N][Hb][Wb][ldb]
    int lda, ldb, Ha, Wa, Hb, Wb;
    int N, Hi, Wi;     // iteration image (per sample)
    int a_mul, b_mul;
    int taps;
    int8_t ady[kWgradMaxTaps], adx[kWgradMaxTaps], bdy[kWgradMaxTaps], bdx[kWgradMaxTaps];
    int Cout, Cin;     // channels taken from A / B
    float* dW;
    long long s_co, s_ci, s_n;
    int tap_off[kWgradMaxTaps];
    int ksplit;        // pixel-range splits per sample (filled by the launcher)
};
int launch_wgrad(WgradParams& p, int num_sms, cudaStream_t s);

// ---- input conv of downs.0.0 (2 -> C, 3x3) and its 1x1 res_conv: weight gradients only (unitspeech.py:170,49,66)
//   dW3[co][ci][kh][kw] += sum d_raw[n][p][co] * in[ci][n][y+kh-1][x+kw-1],  dW1[co][ci] += sum d_res[n][p][co] * in[ci][n][p]
//   in[0] = mu * m, in[1] = x * m
int launch_first_conv_wgrad(const __half* d_raw, const __half* d_res0, const __half* d_res1, const float* x, const float* mu,
                            const float* mask, float* dW3, float* dW1, int N, int H, int W, int C, int num_sms,
                            cudaStream_t s);

// ---- LinearAttention backward (unitspeech.py:78-96,36-43,99-106)
struct AttnBwdParams {
    const float* G;        // [N][C][hidden] = sum_p d_out[n][p][c] * q[n][p][j]
    const float* cs;       // [N][C] = sum_p d_out[n][p][c]
    const float* wo;       // [C][hidden]
    const float* bo;       // [C]
    const float* g;        // Rezero scalar
    const float* ctx;      // [N][heads][32][32]
    float* dwo;            // [C][hidden] +=
    float* dbo;            // [C] +=
    float* dg;             // [1] +=
    float* dctx;           // [N][heads][32][32] (zeroed by the caller), +=
    __half* weffT;         // [N][hidden][C] = fp16(g * Weff[n][c][j]) : per-sample weight of the dq 1x1 conv
    int N, C, heads;
};
int launch_attn_bwd_small(const AttnBwdParams& p, cudaStream_t s);
// dkv[n][p][h*32+d] = dk, dkv[n][p][hidden + h*32+e] = dv from k, v (qkv[.., koff..], qkv[.., voff..]), softmax
// statistics ms[n][h][2][32] (max, sum), ctx and dctx
int launch_attn_bwd_dkv(const __half* qkv, int ld, int koff, int voff, const float* ms, const float* ctx, const float* dctx,
                        __half* dkv, int N, int P, int heads, int num_sms, cudaStream_t s);

// ---- time / speaker embedding backward (unitspeech.py:109-121,133-134,165-168,61)
struct EmbedBwdParams {
    const float* t;        // [N]
    const float* spk;      // [N][S]
    const float* freqs;    // [dim/2]
    const float* w0;       // [4*dim][dim]
    const float* b0;
    const float* w2;       // [dim][4*dim]
    const float* b2;
    const float* wcat;     // [J][dim+S]
    const float* u;        // [N][dim+S] forward value Mish(cat(time_mlp(t), spk))
    const float* dE;       // [N][J]
    float* du;             // [N][dim+S] scratch (zeroed by the launcher)
    float *dw0, *db0, *dw2, *db2, *dwcat, *dbcat;   // +=
    int N, dim, S, J;
    float pe_scale;
};
int launch_embed_bwd(const EmbedBwdParams& p, cudaStream_t s);

// ---- objective gradient: dscore[n][p] = S * 2 (score * sd_n + zm) * sd_n / (sum(mask) * F), sd_n = sqrt(1 - exp(-cum_noise(t_n)))
//      (unitspeech.py:402-404); msum: device scalar scratch
int launch_loss_grad(const float* score, const float* zm, const float* mask, const float* t, float beta_min, float beta_max,
                     float loss_scale, float* msum, float* dscore, int B, int F, int T, cudaStream_t s);
// out[0] += sum_i a[i] * (b ? b[i] : 1)   (final_conv bias gradient)
int launch_dot(const float* a, const float* b, long long n, float* out, cudaStream_t s);

// ---- fp32 master weights -> fp16 GEMM operands.  The master copy (and its gradient / Adam state) is kept in the
// TRAINING LAYOUT = the forward operand layout: kinds 0,1 (3x3): [Cout][9][Cin]; 2 (1x1): [Cout][Cin];
// 3 (ConvTranspose 4x4/s2): [4 phases][Cout][4 taps][Cin] (phase / tap order of pack_conv_host in engine.cu).
//   fwd:   plain fp32 -> fp16 cast (needs the whole input-channel range)
//   dgrad: 0: [Cs][9 (flipped)][Cout]; 1: [4 phases][Cs][4 (zero padded)][Cout]; 2: [Cs][Cout]; 3: [Cs][16][Cout]
// ci0/ci1: the [ci0, ci1) slice of the input channels (skip-concat halves), Cs = ci1 - ci0.  Either output may be null.
// dst = fp16(src), n a multiple of 4 (the flat master buffer -> its fp16 mirror in one launch)
int launch_cast_h(const float* src, __half* dst, long long n, cudaStream_t s);
struct PackDgradParams {
    const float* src;
    __half* dst;
    int Cout, Cs, ci0;
    long long s_src_co, s_dst_ci;
    int n;                       // destination tap slots
    long long src_off[16];       // < 0: the slot is zero (padding tap of the transposed 3x3/s2 conv)
    long long dst_off[16];
};
// Recorder that turns the per-conv data-gradient packs of one weight refresh into ONE launch: between begin and flush,
// launch_pack_conv only records its dgrad part; flush uploads the table when it changed (first use) and launches.
struct PackBatch;
PackBatch* pack_batch_create();
void pack_batch_destroy(PackBatch* b);
void pack_batch_begin(PackBatch* b);
int pack_batch_flush(PackBatch* b, cudaStream_t s);
int launch_pack_conv(int kind, const float* w, int Cout, int Cin, int ci0, int ci1, __half* fwd, __half* dgrad, cudaStream_t s,
                     PackBatch* batch = nullptr);

// ---- clip_grad_norm_ + Adam (finetune.py:81,163-165; torch.optim.Adam defaults betas (0.9, 0.999), eps 1e-8)
int launch_sumsq(const float* g, long long n, double* out, cudaStream_t s);   // out[0] += sum g^2
struct AdamParams {
    float* p; const float* g; float* m; float* v;
    long long n;
    float lr, beta1, beta2, eps;
    float bc1, bc2;           // 1 - beta^step (host-computed; ignored when step_dev is set)
    int* step_dev;            // device step counter or null: the step number is *step_dev + 1, and the counter is
                              // incremented after a non-skipped update (so a captured CUDA graph can be replayed)
    const double* sumsq;      // sum of squares of the scaled gradients
    float inv_scale;          // 1 / loss scale
    float max_norm;           // <= 0: no clipping
    int* skipped;             // += 1 when the gradients are not finite (the step is skipped)
};
int launch_adam(const AdamParams& p, int num_sms, cudaStream_t s);

}  // namespace usb

// ---- weight gradient on the 5th-generation tensor cores (wgrad_tc.cu)
//   D[co][ci] (TMEM, fp32) += sum over a chunk of BH x BW pixels of dY[p][co] * X[p + tap][ci]: both operands are
//   pixel-major tiles ([pixel][64 channels], 128-byte rows, SWIZZLE_128B) exactly as TMA writes them, consumed as
//   MN-major UMMA operands (the GEMM K axis is the pixel axis).  One CTA = one (tap, 128 x n_tile) tile of dW over a
//   range of pixel chunks; partial sums are added to dW with fp32 reductions.
#include <cuda.h>
#include "conv_igemm.h"
namespace usb {
struct WgradTcParams {
    int N, BH, BW, tiles_y, tiles_x;   // iteration image tiling (per sample); kp = BH * BW pixels per chunk (multiple of 16)
    int taps;
    ConvTap atap[kWgradMaxTaps], btap[kWgradMaxTaps];   // per-tap load offsets of dY (A) and X (B)
    int Cout, Cin, n_tile, tiles_m, tiles_n;
    float* dW;
    long long s_co, s_ci, s_n;
    int tap_off[kWgradMaxTaps];
    int ksplit, stages;
    int overwrite;     // dW is known to be zero: a launch with one CTA per tile stores instead of adding
};
int launch_wgrad_tc(WgradTcParams& p, const CUtensorMap& map_a, const CUtensorMap& map_b, int num_sms, cudaStream_t s);
}  // namespace usb
