/* C ABI of libunitspeech_b200.so -- the B200 (sm_100a) reverse-diffusion mel decoder of UnitSpeech.
 *
 * The reference has no FFI: its boundary is the Python class unitspeech.unitspeech.UnitSpeech
 * (unitspeech/unitspeech.py:220).  This header is the boundary we introduce underneath an identically named
 * Python class (unitspeech_b200/decoder.py); each entry cites the reference interface it stands in for.
 *
 * Conventions
 *  - plain C types only; every tensor is a raw pointer to contiguous fp32 in the reference's own layout
 *    ((B, n_feats, T) mels, (B, T) masks, (B, spk_emb_dim) speaker embeddings);
 *  - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are CPU memory;
 *  - the caller owns all tensors; the handle owns weights and workspace;
 *  - every function returns 0 on success, non-zero on failure; usb_last_error() gives the message of the last
 *    failure on the calling thread.  Nothing throws across the ABI and there is no CPU fallback;
 *  - a handle is bound to one device and is not thread-safe; all work is enqueued on the given stream
 *    (pass the integer value of a cudaStream_t, 0 = legacy default stream); no hidden synchronisation except
 *    where stated (the *_host entries, the first call with a new (Be, T) shape or a larger step count, which (re)allocates
 *    workspace, usb_saturation_count, usb_set_output_denorm, and calls made while profiling is on).
 *  - batch semantics: every utterance is sampled exactly like a batch-1 call of the reference (the reference's
 *    own B>1 path is broken, unitspeech/unitspeech.py:338-347,301-305).
 */
#ifndef UNITSPEECH_B200_H
#define UNITSPEECH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct usb_handle usb_handle;

/* Constructor arguments of UnitSpeech / GradLogPEstimator2d (unitspeech/unitspeech.py:221-233,125-162). */
typedef struct usb_config {
    int32_t n_feats;        /* mel bins (80) */
    int32_t dim;            /* base channels (128); must be 64, 128 or 256 */
    int32_t n_mults;        /* len(dim_mults), 2..4 */
    int32_t dim_mults[8];   /* (1, 2, 4, 8) */
    int32_t groups;         /* GroupNorm groups (8) */
    int32_t spk_emb_dim;    /* 256 */
    float pe_scale;         /* 1000 */
    float beta_min;         /* 0.05 (kept for completeness; the schedule is computed by the host) */
    float beta_max;         /* 20 */
    int32_t device;         /* CUDA device ordinal */
} usb_config;

const char* usb_last_error(void);
int usb_version(void);

/* UnitSpeech(...).to(device)                                                  unitspeech/unitspeech.py:221 */
int usb_create(const usb_config* cfg, usb_handle** out);
void usb_destroy(usb_handle* h);

/* load_state_dict: one call per state_dict entry, reference key names and shapes (SURVEY Appendix B.3;
 * inference.py:66-73).  `data` is host fp32.  The extra key "__posemb_freqs" ((dim/2,)) carries the sinusoidal
 * frequency table computed by the host with the reference's torch ops (unitspeech/unitspeech.py:116-118). */
int usb_load_param(usb_handle* h, const char* key, const float* host_data, const int64_t* shape, int32_t ndim);
/* repack to kernel layouts (fp16 K-major GEMM operands) and upload; call after the last usb_load_param */
int usb_finalize_params(usb_handle* h);

/* GradLogPEstimator2d.forward(x, mask, mu, t, spk_emb)                         unitspeech/unitspeech.py:164-201
 * x, mu: (Be, n_feats, T) dev; mask: (Be, T) dev; t: (Be,) dev; spk: (Be, spk_emb_dim) dev; out: (Be, n_feats, T) dev */
int usb_estimator_forward(usb_handle* h, const float* x, const float* mu, const float* mask, const float* t,
                          const float* spk, float* out, int32_t Be, int32_t T, uint64_t stream);

/* UnitSpeech.forward_diffusion(x0, mask, t)                                   unitspeech/unitspeech.py:376-384
 * x0, z: (B, n_feats, T) dev; mask: (B, T) dev; t: (B,) dev.  z is the N(0,1) draw the reference makes at :381, supplied
 * by the host.  xt_out = (x0*exp(-cn/2) + z*sqrt(1-exp(-cn)))*mask; zmask_out = z*mask (may be NULL). */
int usb_forward_diffusion(usb_handle* h, const float* x0, const float* mask, const float* t, const float* z, float* xt_out,
                          float* zmask_out, int32_t B, int32_t T, uint64_t stream);
/* UnitSpeech.loss_t(x0, mask, cond, t, spk_emb) -- the fine-tuning objective    unitspeech/unitspeech.py:393-405
 * Forward value (one fused call); gradients and the optimizer step are the operator-level entries of unitspeech_b200_train.h.
 * loss_out: dev float[1]; xt_out: (B, n_feats, T) dev or NULL. */
int usb_loss_t(usb_handle* h, const float* x0, const float* cond, const float* mask, const float* t, const float* spk,
               const float* z, float* loss_out, float* xt_out, int32_t B, int32_t T, uint64_t stream);

/* UnitSpeech.reverse_diffusion(z, mask, cond, spk_emb, n_timesteps, tg, sg)     unitspeech/unitspeech.py:334-374
 * z, cond: (B, n_feats, T) dev; mask: (B, T) dev; spk: (B, spk_emb_dim) dev;
 * noise: (n, B, n_feats, T) dev, the per-step randn draws (:367), or NULL for all-zero noise;
 * coef: host (n, 3) = [c_x, c_s, sigma] per step i (closed form of :273-296 at table index n-1-i);
 * t_steps: host (n,) = t_i (:361);  text_uncon (n_feats,) and spk_uncon_normed (spk_emb_dim,) come from the
 * loaded parameters (spk_uncon is normalised by its own L2 norm, :358).
 * out: (B, n_feats, T) dev; trace: optional dev (n, B, n_feats, T) receiving x_t after every step (drift reports). */
int usb_reverse_diffusion(usb_handle* h, const float* z, const float* cond, const float* mask, const float* spk,
                          const float* noise, const float* coef_host, const float* t_steps_host, int32_t n_steps,
                          float text_scale, float spk_scale, float* out, float* trace, int32_t B, int32_t T,
                          uint64_t stream);

/* Same call with HOST buffers (pinned or pageable): uploads inputs, runs, downloads `out`, synchronises the stream.
 * This is the end-to-end entry a caller holding CPU tensors uses (inference.py:128-140 with .cpu() tensors). */
int usb_reverse_diffusion_host(usb_handle* h, const float* z, const float* cond, const float* mask,
                               const float* spk, const float* noise, const float* coef_host,
                               const float* t_steps_host, int32_t n_steps, float text_scale, float spk_scale,
                               float* out, int32_t B, int32_t T, uint64_t stream);

/* Front-end glue of UnitSpeech.execute_text_to_speech on the device            unitspeech/unitspeech.py:421-441
 * (durations -> y_lengths, y_mask, generate_path, cond_y = path^T cond_x) for a caller-fixed frame capacity T, so the
 * reference's host round trip int(y_lengths.max()) (:428) disappears.  All pointers dev.  w_ceil: (B, Tx) = ceil(w) *
 * length_scale, already multiplied by x_mask; x_mask: (B, Tx); cond_x: (B, n_feats, Tx).  Outputs: y_lengths (B) int64 =
 * min(clamp_min(sum, 1), T) -- an utterance whose durations exceed the capacity T is truncated at T (the caller can detect
 * it as sum(w_ceil) > T without a host sync); y_mask (B, T); attn (B, Tx, T) the 0/1 alignment path; cond_y (B, n_feats, T). */
int usb_align_expand(usb_handle* h, const float* w_ceil, const float* x_mask, const float* cond_x, int32_t B, int32_t Tx,
                     int32_t n_feats, int32_t T, int64_t* y_lengths, float* y_mask, float* attn, float* cond_y,
                     uint64_t stream);

/* The mel de-normalisation contract of the callers fused into the sampler            inference.py:140
 * mel = (y + 1) / 2 * (mel_max - mel_min) + mel_min with per-bin mel_min / mel_max (host fp32 (n_feats,), as stored in the
 * decoder checkpoint).  While set, `out` of usb_reverse_diffusion[_host] is the de-normalised log-mel, written by the
 * last sampler step (the `trace` stays in normalised space).  Pass NULL, NULL to switch it off.  Synchronises (small
 * host->device copy). */
int usb_set_output_denorm(usb_handle* h, const float* mel_min_host, const float* mel_max_host);

/* fp16 range report.  Activations are stored as fp16 with a saturating pack (values beyond +-65504 are
 * clamped, never inf).  Every conv / GroupNorm-apply epilogue thread that clamped (or exactly reached the limit) bumps a
 * device counter; this reads it (synchronising with the handle's outstanding work) and optionally resets it.  A
 * non-zero count means the fp16 operand format was out of range for this checkpoint / input. */
int usb_saturation_count(usb_handle* h, int64_t* count_out, int32_t reset);

/* Launch-bound (small) workloads: instead of ~120 host launches per diffusion step, the sampler replays ONE captured CUDA
 * graph of a step (device step counter + per-step scalar table; cond / noise / out are staged at fixed device
 * addresses).  mode -1 = auto (on when (CFG branches x utterances) x frames <= 12288, e.g. the reference callers'
 * one-utterance calls, inference.py:128), 0 = off, 1 = on.  Same kernels and arguments as the eager loop: results are
 * bit-identical.  Not used while profiling or when a trace is requested.  usb_graph_steps = steps replayed so far. */
int usb_set_graph_mode(usb_handle* h, int32_t mode);
int64_t usb_graph_steps(usb_handle* h);
/* Small calls also leave most SMs idle in the level-2/3 convolutions (12-60 output tiles for 148 SMs).  In split-K mode
 * those launches cut a tile's K range into up to 8 work items whose fp32 partial tiles are summed in split order by the
 * CTA that finishes last (deterministic).  The split factor depends on the layer geometry only, never on the batch, so
 * an utterance is bit-identical alone or inside a batch AS LONG AS both calls run in the same mode; results differ in
 * the last fp32 bits between the modes.  mode -1 = auto ((CFG branches x utterances) x frames <= 3072), 0 = off, 1 = on. */
int usb_set_splitk_mode(usb_handle* h, int32_t mode);

/* bytes of device workspace the handle holds for (Be, T); 0 if that shape has not been planned yet */
int64_t usb_workspace_bytes(usb_handle* h);
/* number of kernels launched by the handle since creation (bench.py's gpu_launches) */
int64_t usb_launch_count(usb_handle* h);

/* Per-kernel-class device timing for roofline reports.  While on, every launch of usb_reverse_diffusion is bracketed
 * by CUDA events on the caller's stream and the call ends with a stream synchronisation (so never leave it on in a
 * timed run).  usb_get_profile returns accumulated milliseconds, algorithmic work and launch counts for 4 classes:
 * 0 = tcgen05 implicit-GEMM convs (work = FLOPs), 1 = GroupNorm-apply/Mish (bytes), 2 = attention context (bytes),
 * 3 = embedding / input conv / final fused step (bytes).  usb_set_profiling also clears the accumulators. */
int usb_set_profiling(usb_handle* h, int32_t on);
int usb_get_profile(usb_handle* h, double* ms4, double* work4, int64_t* launches4);

/* ---- operator-level entries (parity tests of single kernels against torch fp32) ---------------------------- */
/* NHWC fp16 convolution on the tcgen05 implicit-GEMM kernel.
 * kind: 0 = 3x3 stride 1 pad 1, 1 = 3x3 stride 2 pad 1, 2 = 1x1, 3 = ConvTranspose 4x4 stride 2 pad 1.
 * in0: (N, H, W, C0) fp16 dev; in1: optional second K source (N, H, W, C1) (kind 0/2 only);
 * weight: host fp32 in the reference's layout ((Cout, C0+C1, k, k); ConvTranspose (Cin, Cout, 4, 4));
 * bias: host fp32 (Cout) or NULL; mask: dev fp32 (N, Wout) or NULL; residual: dev fp16 shaped like out or NULL;
 * res_scale: residual blend out = conv*res_scale + residual; stats: dev int64 (N, groups, 2) accumulated fixed-point
 * GroupNorm partials (sum * 2^24, sumsq * 2^18 of conv+bias; integer atomics make them order-independent) or NULL;
 * out: (N, Hout, Wout, Cout) fp16 dev. */
int usb_op_conv(usb_handle* h, int32_t kind, const void* in0, const void* in1, int32_t N, int32_t H, int32_t W,
                int32_t C0, int32_t C1, int32_t Cout, const float* weight_host, const float* bias_host,
                const float* mask, const void* residual, float res_scale, int64_t* stats, int32_t groups, void* out,
                uint64_t stream);
/* kernel-tuning aid: average device milliseconds of one conv launch on internally allocated buffers */
int usb_dbg_conv_time(usb_handle* h, int32_t kind, int32_t N, int32_t H, int32_t W, int32_t C0, int32_t C1,
                      int32_t Cout, int32_t with_stats, int32_t iters, float* ms_out);
/* out = (Mish(GroupNorm(raw)) + addvec[n][c] + res) * mask on NHWC fp16; gamma/beta/addvec dev fp32 */
int usb_op_gn_apply(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta,
                    const float* addvec, const void* res, const float* mask, void* out, int32_t N, int32_t H,
                    int32_t W, int32_t C, int32_t groups, uint64_t stream);
/* LinearAttention context folded with to_out: qkv (N, P, 384) fp16 dev, wo dev fp32 (C, 128) -> weff (N, C, 128) fp16 */
int usb_op_attn_context(usb_handle* h, const void* qkv, const float* wo, void* weff, int32_t N, int32_t P, int32_t C,
                        int32_t heads, uint64_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Vocoder stage: the BigVGAN generator the reference runs after the decoder.
 * Stands in for unitspeech.vocoder.models.BigVGAN (unitspeech/vocoder/models.py:121-191) as built by
 * unitspeech/util.py:174-181 (get_vocoder: load generator weights, remove_weight_norm, eval).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct usb_vocoder usb_vocoder;

/* The fields of the reference's vocoder config.json that the generator reads (models.py:125-167). */
typedef struct usb_vocoder_config {
    int32_t num_mels;                    /* 80 */
    int32_t n_upsamples;                 /* len(upsample_rates) */
    int32_t upsample_rates[8];           /* each 1..4 */
    int32_t upsample_kernel_sizes[8];    /* must equal 2 * rate */
    int32_t upsample_initial_channel;    /* 1536 */
    int32_t resblock_type;               /* 1 = AMPBlock1, 2 = AMPBlock2 */
    int32_t n_resblock_kernels;          /* len(resblock_kernel_sizes) */
    int32_t resblock_kernel_sizes[4];    /* odd, <= 15 */
    int32_t n_dilations;                 /* dilations per resblock (same count for every kernel size) */
    int32_t resblock_dilations[4][4];
    int32_t activation;                  /* 0 = "snake", 1 = "snakebeta" */
    int32_t snake_logscale;              /* 0 / 1 */
    int32_t device;
} usb_vocoder_config;

/* BigVGAN(h)                                                                   unitspeech/vocoder/models.py:123-167 */
int usb_vocoder_create(const usb_vocoder_config* cfg, usb_vocoder** out);
void usb_vocoder_destroy(usb_vocoder* h);
/* generator.load_state_dict(...) after remove_weight_norm(): keys conv_pre.weight, ups.i.0.weight,
 * resblocks.j.convs1.l.weight, resblocks.j.activations.l.act.alpha, ... (host fp32, reference shapes).  A caller
 * holding weight-norm tensors folds them first (weight = g * v / ||v||; unitspeech_b200/vocoder.py does). */
int usb_vocoder_load_param(usb_vocoder* h, const char* key, const float* host_data, const int64_t* shape, int32_t ndim);
int usb_vocoder_finalize_params(usb_vocoder* h);
/* BigVGAN.forward(mel)                                                         unitspeech/vocoder/models.py:169-191
 * mel: (B, num_mels, T) dev fp32; out: (B, T * prod(upsample_rates)) dev fp32 (the reference's (B, 1, samples)). */
int usb_vocoder_forward(usb_vocoder* h, const float* mel, int32_t B, int32_t T, float* out, uint64_t stream);
/* same with host buffers: upload, run, download, synchronise */
int usb_vocoder_forward_host(usb_vocoder* h, const float* mel_host, int32_t B, int32_t T, float* out_host);
/* The vocoder then takes the decoder's NORMALISED mel and applies inference.py:140 while packing its input
 * (mel_min / mel_max host fp32 (num_mels,)); NULL, NULL switches it off. */
int usb_vocoder_set_input_denorm(usb_vocoder* h, const float* mel_min_host, const float* mel_max_host);
long long usb_vocoder_launch_count(const usb_vocoder* h);
size_t usb_vocoder_workspace_bytes(const usb_vocoder* h);
/* tensor-core FLOPs (padded channel counts) of one forward at the current (B, T) plan */
double usb_vocoder_flops_per_call(const usb_vocoder* h);
/* per-class CUDA-event timing of the following forward calls (adds a stream sync; not for timed runs):
 * class 0 = tcgen05 convs (work = FLOPs on padded channels), 1 = fused Snake activation (work = fp16 bytes read +
 * written), 2 = resblock averaging (bytes) */
int usb_vocoder_set_profiling(usb_vocoder* h, int32_t on);
int usb_vocoder_get_profile(usb_vocoder* h, double* ms3, double* work3, long long* launches3);
/* Activation1d (alias_free_torch/act.py:23-28) on NLC fp16: x, out (N, L, C) dev, C % 64 == 0; alpha (C) dev =
 * snake frequency, invbeta (C) dev = 1 / (beta + 1e-9), both already exp'd when the parameters are log-scale;
 * channels [c_real, C) are layout padding and are written as zeros. */
int usb_op_snake_act(const void* x, const float* alpha, const float* invbeta, int32_t N, int32_t L, int32_t C,
                     int32_t c_real, void* out, uint64_t stream);
/* One Conv1d layer of the generator (vocoder/models.py:46-58: kernel k odd, dilation d, "same" padding
 * (k*d - d) / 2, xutils.py:get_padding) on NLC fp16 device tensors: x (N, L, c_in), w (c_out, k * c_in) with
 * w[co][t * c_in + ci] = weight[co][ci][t], bias (c_out) fp32 or null, res (N, L, c_out) or null (out = conv + res; res
 * may be `out`), out (N, L, c_out).  Channel counts are the padded ones (multiples of 64); input channels
 * [c_in_real, c_in) must be zero.  The layer picks its kernel exactly as the vocoder does (swapped-operand kernel, or the
 * 1-D halo kernel). */
int usb_op_conv1d(const void* x, const void* w, const float* bias, const void* res, void* out, int32_t N, int32_t L,
                  int32_t c_in, int32_t c_in_real, int32_t c_out, int32_t k, int32_t dilation, uint64_t stream);
/* kaiser_sinc_filter1d(0.25, 0.3, 12) as the library computes it (alias_free_torch/filter.py:28-57) */
int usb_vocoder_filter(float* out12);

#ifdef __cplusplus
}
#endif
#endif
