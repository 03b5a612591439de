// BigVGAN generator (mel -> waveform), the optional second stage of the reference pipeline
// (unitspeech/vocoder/models.py:121-191; called from unitspeech/util.py:174-181 get_vocoder and the scripts).
//
// Data layout: every activation is NLC fp16 ([utterance][sample][channel], channels padded with zeros to a multiple
// of 64) so that all Conv1d / ConvTranspose1d layers run on the tcgen05 implicit-GEMM kernel of conv_igemm.cu with
// H = 1 (dilated taps are x-offsets of the TMA box; the transposed convs are `stride` phase convs with two taps
// each whose outputs interleave as channel blocks of a [N][L][stride*C] view of the upsampled tensor).
// The anti-aliased activation (alias_free_torch/act.py:23-28: replicate-pad, x2 kaiser-sinc upsample, Snake /
// SnakeBeta, low-pass, stride-2 downsample) is ONE kernel: the 2x-rate signal only ever exists in shared memory.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <type_traits>
#include <cuda_runtime.h>

#include "../../include/unitspeech_b200.h"
#include "conv_igemm.h"

namespace usb {

#define VOC_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +       \
                             std::to_string(__LINE__) + ")");                                                \
    } while (0)
#define VOC_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
constexpr int kActTL = 56;                    // output samples per block
constexpr int kActXRows = kActTL + 12;        // input rows t0-6 .. t0+TL+5
constexpr int kActSRows = 2 * kActTL + 10;    // 2x-rate rows 2*t0-5 .. 2*t0+2*TL+4
constexpr int kActSItems = (kActSRows + 3) / 4;   // phase-2 work items of four 2x-rate rows (31: one round of 32 row slots)
constexpr int kActSmemBytes = (kActXRows + 4 * kActSItems) * 64 * 4;   // 48 KB at vb = 8

static inline int act_vb(int creal) {
    const int nvec = (creal + 7) / 8;
    if (nvec % 8 == 0) return 8;
    for (int vb = 8; vb >= 5; --vb)
        if (nvec % vb == 0) return vb;
    return nvec < 8 ? nvec : 8;
}
static inline size_t act_smem_bytes(int vb) { return static_cast<size_t>(kActXRows + 4 * kActSItems) * vb * 8 * 4; }

struct ActParams {
    const __half* x;       // [N][L][C]
    __half* out;           // [N][L][C]
    const float* alpha;    // [C] frequency a (already exp'd for log-scale parameters); 0 in padding channels
    const float* invbeta;  // [C] 1 / (b + 1e-9); 0 in padding channels
    int L, C;              // C = padded channel count (multiple of 64)
    int Creal;             // channels that carry data; [Creal, C) is written as zeros
    int vb;                // 8-channel vectors per block (blockDim = 32 * vb)
    float filt[12];        // kaiser_sinc_filter1d(0.25, 0.3, 12) -- shared by the up- and the down-sampler
};

// shared-memory element (row, vector cv of 8 channels, j in 0..7): two halves per row (channels 0-3 and 4-7 of every
// vector) so that the threads of a row read/write contiguous float4s (act_sidx, defined inside the kernel)
// Activation1d.forward (alias_free_torch/act.py:23-28) fused:
//   u[2m]   = 2 * sum_{q=-3..2} f[5-2q] * x[clamp(m+q)]        (UpSample1d, resample.py:26-33)
//   u[2m+1] = 2 * sum_{q=-2..3} f[6-2q] * x[clamp(m+q)]
//   s[i]    = u[i] + invbeta * sin(alpha * u[i])^2             (activations.py:47-59,107-120)
//   out[t]  = sum_{k=0..11} f[k] * s[clamp(2t+k-5)]            (LowPassFilter1d stride 2, filter.py:84-95)
__global__ void __launch_bounds__(256, 3) snake_act_kernel(const ActParams p) {
    extern __shared__ float act_smem[];
    float* xs = act_smem;
    const int tid = threadIdx.x;
    // a block covers `vb` 8-channel vectors (blockDim = 32 * vb, shared-memory rows of vb * 8 floats): vb = 8 for the
    // wide stages, 6 / 3 for the 96-, 48- and 24-channel stages, so every thread has work; only real channels are
    // computed, the layout-padding channels [Creal, C) are written as zeros by the last slab
    const int vb = p.vb;
    const int nvec = (p.Creal + 7) >> 3;
    const int ncv = min(vb, nvec - static_cast<int>(blockIdx.y) * vb);
    const int t0 = blockIdx.x * kActTL;
    const long long nbase = static_cast<long long>(blockIdx.z) * p.L;
    if (blockIdx.y == gridDim.y - 1) {
        const int npad = (p.C >> 3) - nvec;
        for (int idx = tid; idx < kActTL * npad; idx += blockDim.x) {
            const int t = t0 + idx / npad;
            if (t < p.L)
                *reinterpret_cast<uint4*>(p.out + (nbase + t) * p.C + (nvec + idx % npad) * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    const int R = blockDim.x / ncv;          // rows processed in parallel (32 when ncv == vb)
    const int cv = tid % ncv, r = tid / ncv;
    const bool active = r < R;
    const int c0 = (blockIdx.y * vb + cv) * 8;
    float* ss = act_smem + kActXRows * vb * 8;  // 4 * kActSItems rows (the last two are computed but never read)
    const int pitch = vb * 8, hoff = vb * 4;
#define act_sidx(row, cv_, half) ((row) * pitch + (half) * hoff + (cv_) * 4)

    if (active)
        for (int row = r; row < kActXRows; row += R) {
            int t = t0 - 6 + row;
            t = t < 0 ? 0 : (t > p.L - 1 ? p.L - 1 : t);
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p.x + (nbase + t) * p.C + c0));
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
            const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]), c = __half22float2(h2[2]), d = __half22float2(h2[3]);
            *reinterpret_cast<float4*>(xs + act_sidx(row, cv, 0)) = make_float4(a.x, a.y, b.x, b.y);
            *reinterpret_cast<float4*>(xs + act_sidx(row, cv, 1)) = make_float4(c.x, c.y, d.x, d.y);
        }
    float al[8], ib[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        al[j] = __ldg(p.alpha + c0 + j);
        ib[j] = __ldg(p.invbeta + c0 + j);
    }
    __syncthreads();

    // phase 2: the 2x-rate signal after the Snake non-linearity, four rows per item.  Rows i0 (odd), i0+1, i0+2, i0+3
    // read the seven input rows m0-2 .. m0+4 (m0 = i0 >> 1): each shared-memory row is loaded once per item.
    if (active)
        for (int it = r; it < kActSItems; it += R) {
            const int ii0 = 4 * it;
            const int i0 = 2 * t0 - 5 + ii0;
            float u[4][8];
            if (i0 >= 0 && i0 + 3 <= 2 * p.L - 1) {
                const int rowb = (i0 >> 1) - 2 - (t0 - 6);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) u[q][j] = 0.f;
#pragma unroll
                for (int xr = 0; xr < 7; ++xr) {
                    const float4 lo = *reinterpret_cast<const float4*>(xs + act_sidx(rowb + xr, cv, 0));
                    const float4 hi = *reinterpret_cast<const float4*>(xs + act_sidx(rowb + xr, cv, 1));
                    const float xv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    // row q uses x[first_q + j] with tap f[10 - 2j] (odd rows 0, 2) or f[11 - 2j] (even rows 1, 3);
                    // first_q - (m0 - 2) = 0, 0, 1, 1
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = xr - (q >> 1);
                        if (j >= 0 && j < 6) {
                            const float w = 2.f * ((q & 1) ? p.filt[11 - 2 * j] : p.filt[10 - 2 * j]);
#pragma unroll
                            for (int c = 0; c < 8; ++c) u[q][c] = fmaf(w, xv[c], u[q][c]);
                        }
                    }
                }
            } else {
                // block edges of the signal: rows are clamped individually (replicate padding of the low-pass filter)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int i = i0 + q;
                    i = i < 0 ? 0 : (i > 2 * p.L - 1 ? 2 * p.L - 1 : i);
                    const int m = i >> 1, odd = i & 1;
                    const int row0 = m - 3 + odd - (t0 - 6);
#pragma unroll
                    for (int c = 0; c < 8; ++c) u[q][c] = 0.f;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        const float w = 2.f * (odd ? p.filt[10 - 2 * j] : p.filt[11 - 2 * j]);
                        const float4 lo = *reinterpret_cast<const float4*>(xs + act_sidx(row0 + j, cv, 0));
                        const float4 hi = *reinterpret_cast<const float4*>(xs + act_sidx(row0 + j, cv, 1));
                        u[q][0] = fmaf(w, lo.x, u[q][0]); u[q][1] = fmaf(w, lo.y, u[q][1]);
                        u[q][2] = fmaf(w, lo.z, u[q][2]); u[q][3] = fmaf(w, lo.w, u[q][3]);
                        u[q][4] = fmaf(w, hi.x, u[q][4]); u[q][5] = fmaf(w, hi.y, u[q][5]);
                        u[q][6] = fmaf(w, hi.z, u[q][6]); u[q][7] = fmaf(w, hi.w, u[q][7]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // MUFU.SIN reduces its argument itself (x / 2pi in fp32): absolute error ~|x| * 2^-24, far below
                    // the fp16 storage of the result for the |alpha * u| < 1e3 seen here
                    const float sn = __sinf(al[j] * u[q][j]);
                    u[q][j] = fmaf(ib[j] * sn, sn, u[q][j]);
                }
                *reinterpret_cast<float4*>(ss + act_sidx(ii0 + q, cv, 0)) = make_float4(u[q][0], u[q][1], u[q][2], u[q][3]);
                *reinterpret_cast<float4*>(ss + act_sidx(ii0 + q, cv, 1)) = make_float4(u[q][4], u[q][5], u[q][6], u[q][7]);
            }
        }
    __syncthreads();

    // phase 3: low-pass + decimation, two outputs per item: 14 shared rows feed 2 x 12 taps
    if (active)
        for (int it = r; it < kActTL / 2; it += R) {
            const int tt = 2 * it;
            const int t = t0 + tt;
            if (t >= p.L) break;
            float o[2][8];
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < 8; ++c) o[q][c] = 0.f;
#pragma unroll
            for (int rr = 0; rr < 14; ++rr) {
                const float4 lo = *reinterpret_cast<const float4*>(ss + act_sidx(2 * tt + rr, cv, 0));
                const float4 hi = *reinterpret_cast<const float4*>(ss + act_sidx(2 * tt + rr, cv, 1));
                const float sv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = rr - 2 * q;
                    if (k >= 0 && k < 12) {
                        const float w = p.filt[k];
#pragma unroll
                        for (int c = 0; c < 8; ++c) o[q][c] = fmaf(w, sv[c], o[q][c]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (t + q >= p.L) break;
                uint4 pk;
                __half2* h2 = reinterpret_cast<__half2*>(&pk);
                h2[0] = __floats2half2_rn(o[q][0], o[q][1]); h2[1] = __floats2half2_rn(o[q][2], o[q][3]);
                h2[2] = __floats2half2_rn(o[q][4], o[q][5]); h2[3] = __floats2half2_rn(o[q][6], o[q][7]);
                *reinterpret_cast<uint4*>(p.out + (nbase + t + q) * p.C + c0) = pk;
            }
        }
}

#undef act_sidx

// =====================================================================================================================
// Activation1d, second version: register-resident 2x-rate signal.
// The first kernel keeps the up-sampled, Snake'd signal in shared memory (25 128-bit shared-memory operations per 8
// output values: it is shared-memory-bandwidth bound at 1/6 of the HBM roofline, ncu profiles/r1_snake_act_*).  Here a thread
// owns 4 channels and a RUN of kAct2TO consecutive outputs and walks it sample by sample: step t computes the two
// 2x-rate samples s[2t+5], s[2t+6] (the last two that output t needs; both are filters over the same six inputs
// x[t .. t+5]) and scatters them into the six partial outputs t .. t+5 they contribute to (12 taps: out[t'] +=
// f[10-2d] s[2t+5] + f[11-2d] s[2t+6], d = t' - t); output t is then complete.  The 2x-rate signal never leaves registers;
// shared memory only holds the fp16 input tile (one 8-byte read per step).  A run re-computes the 10 samples before its
// first output (5 warm-up steps): 30 steps for 25 outputs (24 for 19 in the first version).
// =====================================================================================================================
constexpr int kAct2TO = 25;                 // outputs per run (19: 24 steps per 19 outputs; 25: 30 per 25, two 51 KB buffers, still two blocks per SM)
constexpr int kAct2Steps = kAct2TO + 5;     // + warm-up; a multiple of 6 (the step loop is unrolled by the 6 partial outputs)
static_assert(kAct2Steps % 6 == 0, "the step loop is unrolled in groups of 6");

static inline int act2_ncol(int creal) {    // 4-channel columns per block: the largest even divisor <= 16 of creal / 4
    const int cols = (creal + 3) / 4;
    for (int n = 16; n >= 2; n -= 2)
        if (cols % n == 0) return n;
    return 2;
}
static inline size_t act2_smem_bytes(int ncol) { return static_cast<size_t>((256 / ncol) * kAct2TO + 10) * ncol * 8; }

// packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2: two IEEE fp32 operations per issue slot): a thread's 4 channels are two
// packed pairs, which halves the issue slots of the kernel's ~100 FMAs per output step (ncu: issue- and FMA-pipe bound)
typedef unsigned long long vf2;
__device__ __forceinline__ vf2 vpk2(float lo, float hi) {
    vf2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void vupk2(vf2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ vf2 vfma2(vf2 a, vf2 b, vf2 c) {
    vf2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ vf2 vmul2(vf2 a, vf2 b) {
    vf2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// one 2x-rate sample for 4 channels (two packed pairs) from six consecutive input rows; ODD: i = 2m+1 uses f[10-2j], even
// uses f[11-2j] (f2 = the taps times the up-sampling ratio, replicated into both halves)
// The 12-tap filter is symmetric, f[k] == f[11-k]: g[m] = f[2m] (m = 0..5) holds every tap, f[2m+1] = g[5-m]
// (24 instead of 48 registers for the two packed tap tables).
__host__ __device__ constexpr int act2_tap(int k) { return (k & 1) ? (11 - k) >> 1 : k >> 1; }
template <bool ODD>
__device__ __forceinline__ void act2_srow(const vf2 (&x)[6][2], const vf2 (&g2)[6], const vf2 (&al)[2], const vf2 (&ib)[2],
                                          vf2 (&s)[2]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        vf2 u = vmul2(g2[act2_tap(ODD ? 10 : 11)], x[0][c]);
#pragma unroll
        for (int j = 1; j < 6; ++j) u = vfma2(g2[act2_tap(ODD ? 10 - 2 * j : 11 - 2 * j)], x[j][c], u);
        // MUFU.SIN reduces its argument itself; absolute error ~|x| * 2^-24, far below the fp16 storage of the result
        float a0, a1;
        vupk2(vmul2(al[c], u), a0, a1);
        const vf2 sn = vpk2(__sinf(a0), __sinf(a1));
        s[c] = vfma2(vmul2(ib[c], sn), sn, u);
    }
}

__device__ __forceinline__ void act2_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// A block walks `tiles_per_block` consecutive time tiles with two input buffers: the next tile's inputs arrive by cp.async
// while the current tile is computed (ncu of the one-tile-per-block version: warps stalled on the tile load, 58 % issue-active)
__global__ void __launch_bounds__(256, 2) snake_act2_kernel(const ActParams p, int ncol, int tiles_per_block) {
    extern __shared__ __align__(16) unsigned char act2_smem[];     // 2 x fp16 input tile [rows][ncol * 4 channels]
    const int tid = threadIdx.x;
    const int R = blockDim.x / ncol;                 // runs per block
    const int TLB = R * kAct2TO;                     // outputs per tile
    const int c_blk = blockIdx.y * ncol * 4;         // first channel of this block
    const long long nbase = static_cast<long long>(blockIdx.z) * p.L;
    const int pitch = ncol * 8;                      // bytes per tile row
    const int rows = TLB + 10;                       // inputs t0-5 .. t0+TLB+4
    const int buf_bytes = (rows * pitch + 15) & ~15;
    const int n_tiles_x = (p.L + TLB - 1) / TLB;
    const int tile0 = blockIdx.x * tiles_per_block;
    const int n_my = min(tiles_per_block, n_tiles_x - tile0);
    const uint32_t smem0 = static_cast<uint32_t>(__cvta_generic_to_shared(act2_smem));
    // tile load, replicate padding at the signal ends (UpSample1d pads by replication, resample.py:26-29)
    auto issue_load = [&](int tile, int buf) {
        const int t0 = tile * TLB;
        const int cpr = ncol >> 1;                   // 16-byte chunks per row
        for (int idx = tid; idx < rows * cpr; idx += blockDim.x) {
            const int row = idx / cpr, ch = idx - row * cpr;
            int t = t0 - 5 + row;
            t = t < 0 ? 0 : (t > p.L - 1 ? p.L - 1 : t);
            act2_cp_async16(smem0 + buf * buf_bytes + row * pitch + ch * 16, p.x + (nbase + t) * p.C + c_blk + ch * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int col = tid % ncol, run = tid / ncol;
    const int c0 = c_blk + col * 4;
    vf2 al[2], ib[2], f2[6], f1[6];      // taps f[2m], m = 0..5 (act2_tap)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        al[j] = vpk2(__ldg(p.alpha + c0 + 2 * j), __ldg(p.alpha + c0 + 2 * j + 1));
        ib[j] = vpk2(__ldg(p.invbeta + c0 + 2 * j), __ldg(p.invbeta + c0 + 2 * j + 1));
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        f1[k] = vpk2(p.filt[2 * k], p.filt[2 * k]);
        f2[k] = vpk2(2.f * p.filt[2 * k], 2.f * p.filt[2 * k]);      // UpSample1d multiplies by the ratio (resample.py:31)
    }
    const int imax = 2 * p.L - 1;
    if (n_my > 0) issue_load(tile0, 0);
    for (int kt = 0; kt < n_my; ++kt) {
        const int buf = kt & 1;
        const int t0 = (tile0 + kt) * TLB;
        if (kt + 1 < n_my) {
            issue_load(tile0 + kt + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (blockIdx.y == 0) {                           // layout padding channels [Creal, C) are written as zeros
            const int npad = (p.C - p.Creal) >> 3;
            const int pad0 = p.Creal >> 3;
            for (int idx = tid; idx < TLB * npad; idx += blockDim.x) {
                const int t = t0 + idx / npad;
                if (t < p.L)
                    *reinterpret_cast<uint4*>(p.out + (nbase + t) * p.C + (pad0 + idx % npad) * 8) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        const int tr = t0 + run * kAct2TO;               // first output of this run
        if (tr < p.L) {
            const unsigned char* xcol = act2_smem + buf * buf_bytes + col * 8;
            auto load_row = [&](int t, vf2 (&x)[2]) {        // input row t (tile-relative), 4 channels as two packed pairs
                const uint2 raw = *reinterpret_cast<const uint2*>(xcol + (t - (t0 - 5)) * pitch);
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                x[0] = vpk2(a.x, a.y);
                x[1] = vpk2(b.x, b.y);
            };
            // a sample whose index falls outside [0, 2L-1] is the replicate padding of the low-pass filter (filter.py:90):
            // it equals s[0] / s[2L-1], recomputed from the inputs around that end
            auto edge_sample = [&](int i, vf2 (&sv)[2]) {
                const int ic = i < 0 ? 0 : imax;
                const int m = ic >> 1;
                vf2 x[6][2];
                if (ic & 1) {
        #pragma unroll
                    for (int j = 0; j < 6; ++j) load_row(m - 2 + j, x[j]);
                    act2_srow<true>(x, f2, al, ib, sv);
                } else {
        #pragma unroll
                    for (int j = 0; j < 6; ++j) load_row(m - 3 + j, x[j]);
                    act2_srow<false>(x, f2, al, ib, sv);
                }
            };
            vf2 acc[6][2];
            const vf2 zero2 = vpk2(0.f, 0.f);
        #pragma unroll
            for (int d = 0; d < 6; ++d) acc[d][0] = acc[d][1] = zero2;
            for (int g = 0; g < kAct2Steps / 6; ++g) {
                const int tg = tr - 5 + g * 6;               // first step of this group
                if (tg >= p.L) break;                        // every remaining output lies beyond the signal
                vf2 xw[11][2];                               // inputs tg .. tg+10 serve the six steps of the group
        #pragma unroll
                for (int j = 0; j < 11; ++j) load_row(tg + j, xw[j]);
                // interior group (all but the first and last of a signal): every 2x-rate sample s[2tg+5 .. 2tg+16] exists and
                // every output tg .. tg+5 lies inside the signal, so the per-step range tests and their branches disappear
                // (ncu: integer compares, branches and reconvergence points were ~15 % of the executed instructions)
                const bool interior = 2 * tg + 5 >= 0 && 2 * tg + 16 <= imax && tg + 5 < p.L;
                __half* outg = p.out + (nbase + tg) * p.C + c0;      // output row tg of this thread's channels
                auto six_steps = [&](auto fast_c) {
                    constexpr bool FAST = decltype(fast_c)::value;
        #pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int t = tg + k;                    // this step completes output t (slot k of acc)
                        vf2 so[2], se[2];                        // s[2t+5] (odd index), s[2t+6] (even index)
                        vf2 x6[6][2];
        #pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            x6[j][0] = xw[k + j][0];
                            x6[j][1] = xw[k + j][1];
                        }
                        if (FAST) {
                            act2_srow<true>(x6, f2, al, ib, so);
                            act2_srow<false>(x6, f2, al, ib, se);
                        } else {
                            const int io = 2 * t + 5, ie = 2 * t + 6;
                            if (io >= 0 && io <= imax) act2_srow<true>(x6, f2, al, ib, so);
                            else edge_sample(io, so);
                            if (ie >= 0 && ie <= imax) act2_srow<false>(x6, f2, al, ib, se);
                            else edge_sample(ie, se);
                        }
                        // scatter into the partial outputs t .. t+5: slot (k + d) % 6 holds output t + d
        #pragma unroll
                        for (int d = 0; d < 6; ++d)
        #pragma unroll
                            for (int c = 0; c < 2; ++c)
                                acc[(k + d) % 6][c] = vfma2(f1[act2_tap(10 - 2 * d)], so[c],
                                                            vfma2(f1[act2_tap(11 - 2 * d)], se[c], acc[(k + d) % 6][c]));
                        // output t is complete (all 12 taps added over steps t-5 .. t); emit it unless it is a warm-up step
                        // (t >= tr  <=>  6g + k >= 5)
                        if ((g > 0 || k == 5) && (FAST || t < p.L)) {
                            float o0, o1, o2, o3;
                            vupk2(acc[k][0], o0, o1);
                            vupk2(acc[k][1], o2, o3);
                            uint2 pk;
                            *reinterpret_cast<__half2*>(&pk.x) = __floats2half2_rn(o0, o1);
                            *reinterpret_cast<__half2*>(&pk.y) = __floats2half2_rn(o2, o3);
                            *reinterpret_cast<uint2*>(outg + static_cast<long long>(k) * p.C) = pk;
                        }
                        acc[k][0] = acc[k][1] = zero2;           // the slot now collects output t + 6
                    }
                };
                if (interior) six_steps(std::true_type{});
                else six_steps(std::false_type{});
            }
        }
        __syncthreads();     // every thread is done with this buffer before the load issued next iteration overwrites it
    }
}

// USB_SNAKE_V1=1 selects the first (shared-memory) kernel for A/B measurements
static void launch_snake_act(const ActParams& a, int N, cudaStream_t s) {
    static const bool v1 = getenv("USB_SNAKE_V1") != nullptr;
    if (v1 || (a.Creal & 7)) {
        const int nvec = (a.Creal + 7) / 8;
        const dim3 grid((a.L + kActTL - 1) / kActTL, (nvec + a.vb - 1) / a.vb, N);
        snake_act_kernel<<<grid, 32 * a.vb, act_smem_bytes(a.vb), s>>>(a);
        return;
    }
    const int ncol = act2_ncol(a.Creal);
    const int R = 256 / ncol;
    const int tiles_x = (a.L + R * kAct2TO - 1) / (R * kAct2TO);
    // several tiles per block (double-buffered inputs) as long as the grid keeps >= ~4 blocks per SM
    int tpb = 4;
    while (tpb > 1 && static_cast<long long>((tiles_x + tpb - 1) / tpb) * ((a.Creal / 4) / ncol) * N < 4 * 148) tpb >>= 1;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(snake_act2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        attr = true;
    }
    const dim3 grid((tiles_x + tpb - 1) / tpb, (a.Creal / 4) / ncol, N);
    snake_act2_kernel<<<grid, R * ncol, 2 * ((act2_smem_bytes(ncol) + 15) & ~size_t(15)), s>>>(a, ncol, tpb);
}

// mel (B, M, T) fp32 -> [B][T][Cp] fp16, zero padded channels.  mel_min != null: `mel` is the decoder's normalised
// output and is de-normalised on the way in, (y + 1) / 2 * (mel_max - mel_min) + mel_min (inference.py:140)
__global__ void mel_pack_kernel(const float* __restrict__ mel, __half* __restrict__ out, int B, int M, int T, int Cp,
                                const float* __restrict__ mel_min, const float* __restrict__ mel_max) {
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int groups = Cp / 8;
    if (idx >= static_cast<long long>(B) * groups * T) return;
    const int t = static_cast<int>(idx % T);
    const int cg = static_cast<int>((idx / T) % groups);
    const int b = static_cast<int>(idx / (static_cast<long long>(T) * groups));
    uint4 pk;
    __half* h = reinterpret_cast<__half*>(&pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        float v = c < M ? __ldg(mel + (static_cast<long long>(b) * M + c) * T + t) : 0.f;
        if (mel_min && c < M) {
            const float lo = __ldg(mel_min + c), hi = __ldg(mel_max + c);
            v = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(v, 1.f), 2.f), __fsub_rn(hi, lo)), lo);
        }
        h[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(out + (static_cast<long long>(b) * T + t) * Cp + cg * 8) = pk;
}

// Mean of up to three resblock outputs of one stage in ONE pass (models.py:177-184): out = ((r0 + r1) + r2) * scale with the
// same fp32 additions, in the same order, as the accumulate-as-you-go kernel below -- 8 instead of 24 bytes per element
// and one launch instead of three.  r1 / r2 may be null.
__global__ void stage_mean_kernel(const __half* __restrict__ r0, const __half* __restrict__ r1, const __half* __restrict__ r2,
                                  __half* __restrict__ out, long long n8, float scale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    float v[8];
    {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(r0) + i);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            v[2 * j] = f.x;
            v[2 * j + 1] = f.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const __half* r = k == 0 ? r1 : r2;
        if (r == nullptr) continue;
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(r) + i);
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            v[2 * j] += f.x;
            v[2 * j + 1] += f.y;
        }
    }
    uint4 pk;
    __half2* o2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(v[2 * j] * scale, v[2 * j + 1] * scale);
    reinterpret_cast<uint4*>(out)[i] = pk;
}

// (more than three resblocks per stage) sum of the resblock outputs: mode 0: acc = r; 1: acc += r; 2: out = (acc + r)*scale;
// 3: out = r * scale (single resblock)
__global__ void stage_accum_kernel(const __half* __restrict__ r, float* __restrict__ acc, __half* __restrict__ out,
                                   long long n8, int mode, float scale) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(r) + i);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
    }
    float4* a4 = reinterpret_cast<float4*>(acc) + 2 * i;
    if (mode == 1 || mode == 2) {
        const float4 a = a4[0], b = a4[1];
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (mode <= 1) {
        a4[0] = make_float4(v[0], v[1], v[2], v[3]);
        a4[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 pk;
        __half2* o2 = reinterpret_cast<__half2*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(v[2 * j] * scale, v[2 * j + 1] * scale);
        reinterpret_cast<uint4*>(out)[i] = pk;
    }
}

// conv_post (Conv1d(ch, 1, 7, padding 3)) + tanh (models.py:188-189): one thread per output sample
// w: [7][Cin8] fp32 (Cin8 = real channels rounded up to 8, zero padded)
__global__ void __launch_bounds__(256) conv_post_kernel(const __half* __restrict__ a, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out, int N,
                                                        int L, int C, int Cin8) {
    extern __shared__ float post_w[];
    for (int i = threadIdx.x; i < 7 * Cin8; i += blockDim.x) post_w[i] = __ldg(w + i);
    __syncthreads();
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<long long>(N) * L) return;
    const int t = static_cast<int>(idx % L);
    const long long nbase = idx - t;
    float acc = __ldg(bias);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int tt = t + k - 3;
        if (tt < 0 || tt >= L) continue;
        const __half* row = a + (nbase + tt) * C;
        for (int c = 0; c < Cin8; c += 8) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(row + c));
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
            const float* wk = post_w + k * Cin8 + c;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                acc = fmaf(f.x, wk[2 * j], acc);
                acc = fmaf(f.y, wk[2 * j + 1], acc);
            }
        }
    }
    out[idx] = tanhf(acc);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static inline int pad64(int c) { return (c + 63) / 64 * 64; }

// kaiser_sinc_filter1d(cutoff 0.25, half_width 0.3, kernel 12) -- alias_free_torch/filter.py:28-57, in double precision
static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 64; ++k) {
        term *= (x / (2.0 * k)) * (x / (2.0 * k));
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}
static void kaiser_sinc_12(float* out) {
    const int K = 12, half_size = 6;
    const double cutoff = 0.25, half_width = 0.3, pi = 3.14159265358979323846;
    const double delta_f = 4 * half_width;
    const double A = 2.285 * (half_size - 1) * pi * delta_f + 7.95;
    const double beta = A > 50.0 ? 0.1102 * (A - 8.7) : (A >= 21.0 ? 0.5842 * pow(A - 21, 0.4) + 0.07886 * (A - 21.0) : 0.0);
    double f[12], sum = 0;
    for (int n = 0; n < K; ++n) {
        const double r = 2.0 * n / (K - 1) - 1.0;                       // torch.kaiser_window(periodic=False)
        const double win = bessel_i0(beta * sqrt(fmax(0.0, 1.0 - r * r))) / bessel_i0(beta);
        const double time = (n - half_size) + 0.5;
        const double xx = 2 * cutoff * time;
        const double sinc = xx == 0.0 ? 1.0 : sin(pi * xx) / (pi * xx);
        f[n] = 2 * cutoff * win * sinc;
        sum += f[n];
    }
    // the filter is symmetric (window and sinc are even about the centre); the two halves are forced to be bit-identical
    // so that the kernels can keep six taps instead of twelve in registers
    for (int n = 0; n < K / 2; ++n) out[n] = out[K - 1 - n] = static_cast<float>(0.5 * (f[n] + f[K - 1 - n]) / sum);
}

struct VocParam {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct VocConvW {
    __half* w = nullptr;
    float* bias = nullptr;
    int Cin = 0, Cout = 0;     // padded
    int cin_real = 0;          // input channels that carry data
    int taps = 0, phases = 1;
    int8_t dx[kConvMaxTaps] = {0};
};

struct VocActW {
    float* alpha = nullptr;
    float* invbeta = nullptr;
    int C = 0, Creal = 0;
};

struct VocOp {
    enum Type { PACK, CONV, ACT, ACCUM, POST } type;
    ConvOp conv;
    ActParams act;
    int N = 0;
    // ACCUM
    const __half* r = nullptr;
    const __half* r1 = nullptr;      // mode 4 (stage_mean_kernel): second / third resblock output
    const __half* r2 = nullptr;
    float* acc = nullptr;
    __half* out16 = nullptr;
    long long n8 = 0;
    int mode = 0;
    float scale = 1.f;
    double flops = 0;
};

}  // namespace usb

using namespace usb;

struct usb_vocoder {
    usb_vocoder_config cfg;
    int num_sms = 148;
    bool finalized = false;
    std::map<std::string, VocParam> host;
    std::vector<void*> dev_allocs;
    int n_stage = 0, nk = 0, nd = 0;
    int ch[9] = {0};            // real channels: ch[0] = initial, ch[i+1] after stage i
    VocConvW conv_pre;
    std::vector<VocConvW> ups;
    std::vector<VocConvW> rb_convs;   // [stage][kernel][layer][0|1]  (AMPBlock2: one conv per layer)
    std::vector<VocActW> rb_acts;     // same order as the reference's activations ModuleList per resblock
    VocActW act_post;
    float *post_w = nullptr, *post_b = nullptr;
    int post_cin8 = 0;
    float filt[12];
    // plan
    int plan_B = 0, plan_T = 0;
    void* arena = nullptr;
    size_t arena_bytes = 0;
    __half* melp = nullptr;
    __half* final_act = nullptr;
    int final_L = 0, final_C = 0;
    std::vector<VocOp> ops;
    long long launches = 0;
    double flops_per_call = 0;
    float* mel_range = nullptr;   // device [2][num_mels] (mel_min, mel_max) when the input is the normalised decoder output
    bool denorm = false;
    // optional per-class timing of the next forward calls: 0 conv (tensor), 1 snake activation (HBM), 2 other
    bool profiling = false;
    double prof_ms[3] = {0, 0, 0};
    double prof_work[3] = {0, 0, 0};    // conv: padded FLOPs; activation / other: algorithmic bytes (fp16 read + write)
    long long prof_launches[3] = {0, 0, 0};
};

namespace usb {

template <typename T>
static int voc_upload(usb_vocoder* h, const std::vector<T>& host, T** dev) {
    void* p = nullptr;
    VOC_CUDA(cudaMalloc(&p, host.size() * sizeof(T)));
    h->dev_allocs.push_back(p);
    VOC_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<T*>(p);
    return 0;
}

static int voc_get(usb_vocoder* h, const std::string& key, const std::vector<int64_t>& shape, const VocParam** out) {
    auto it = h->host.find(key);
    if (it == h->host.end()) return set_error("missing vocoder parameter: " + key);
    if (it->second.shape != shape) {
        std::string s = "vocoder parameter " + key + " has shape (";
        for (int64_t d : it->second.shape) s += std::to_string(d) + ",";
        s += ") expected (";
        for (int64_t d : shape) s += std::to_string(d) + ",";
        return set_error(s + ")");
    }
    *out = &it->second;
    return 0;
}

static int voc_load_bias(usb_vocoder* h, const std::string& key, int C, int Cp, float** dev) {
    const VocParam* b;
    VOC_TRY(voc_get(h, key, {C}, &b));
    std::vector<float> v(Cp, 0.f);
    for (int i = 0; i < C; ++i) v[i] = b->data[i];
    return voc_upload(h, v, dev);
}

// Conv1d weight (Cout, Cin, k), dilation d, "same" padding (xutils.get_padding) -> [Cout_p][k*Cin_p] fp16
static int voc_load_conv(usb_vocoder* h, const std::string& prefix, int Cout, int Cin, int k, int dil, VocConvW& w) {
    const VocParam* p;
    VOC_TRY(voc_get(h, prefix + ".weight", {Cout, Cin, k}, &p));
    if (k > kConvMaxTaps || (k & 1) == 0) return set_error("unsupported Conv1d kernel size in " + prefix);
    const int pad = (k * dil - dil) / 2;
    if (pad > 127) return set_error("dilation too large in " + prefix);
    w.Cin = pad64(Cin);
    w.cin_real = Cin;
    w.Cout = pad64(Cout);
    w.taps = k;
    w.phases = 1;
    for (int t = 0; t < k; ++t) w.dx[t] = static_cast<int8_t>(t * dil - pad);
    std::vector<__half> packed(static_cast<size_t>(w.Cout) * k * w.Cin, __float2half_rn(0.f));
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int t = 0; t < k; ++t)
                packed[(static_cast<size_t>(co) * k + t) * w.Cin + ci] =
                    __float2half_rn(p->data[(static_cast<size_t>(co) * Cin + ci) * k + t]);
    VOC_TRY(voc_upload(h, packed, &w.w));
    return voc_load_bias(h, prefix + ".bias", Cout, w.Cout, &w.bias);
}

// ConvTranspose1d weight (Cin, Cout, ku), stride u, padding (ku-u)/2, ku == 2u (models.py:141-146):
//   out[u*x + ph] = sum_a sum_ci in[x + q - a][ci] * w[ci][co][k0 + a*u],  k0 = (ph+pad) % u, q = (ph+pad) / u
// -> [u][Cout_p][2*Cin_p] fp16
static int voc_load_convT(usb_vocoder* h, const std::string& prefix, int Cin, int Cout, int ku, int u, VocConvW& w) {
    const VocParam* p;
    VOC_TRY(voc_get(h, prefix + ".weight", {Cin, Cout, ku}, &p));
    if (ku != 2 * u || u < 1 || u > 4) return set_error("ConvTranspose1d needs kernel == 2*stride and stride <= 4: " + prefix);
    const int pad = (ku - u) / 2;
    w.Cin = pad64(Cin);
    w.Cout = pad64(Cout);
    w.taps = 2;
    w.phases = u;
    std::vector<__half> packed(static_cast<size_t>(u) * w.Cout * 2 * w.Cin, __float2half_rn(0.f));
    for (int ph = 0; ph < u; ++ph) {
        const int k0 = (ph + pad) % u, q = (ph + pad) / u;
        for (int a = 0; a < 2; ++a) {
            w.dx[ph * 2 + a] = static_cast<int8_t>(q - a);
            for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci)
                    packed[((static_cast<size_t>(ph) * w.Cout + co) * 2 + a) * w.Cin + ci] =
                        __float2half_rn(p->data[(static_cast<size_t>(ci) * Cout + co) * ku + k0 + a * u]);
        }
    }
    VOC_TRY(voc_upload(h, packed, &w.w));
    return voc_load_bias(h, prefix + ".bias", Cout, w.Cout, &w.bias);
}

static int voc_load_act(usb_vocoder* h, const std::string& prefix, int C, VocActW& a) {
    const VocParam *pa, *pb;
    VOC_TRY(voc_get(h, prefix + ".act.alpha", {C}, &pa));
    const bool is_beta = h->cfg.activation == 1;
    if (is_beta) VOC_TRY(voc_get(h, prefix + ".act.beta", {C}, &pb));
    else pb = pa;
    a.C = pad64(C);
    a.Creal = C;
    std::vector<float> al(a.C, 0.f), ib(a.C, 0.f);
    for (int c = 0; c < C; ++c) {
        // activations.py:54-59,113-120: exp() in log-scale mode, then 1 / (beta + 1e-9); same fp32 operations
        const float av = h->cfg.snake_logscale ? expf(pa->data[c]) : pa->data[c];
        const float bv = h->cfg.snake_logscale ? expf(pb->data[c]) : pb->data[c];
        al[c] = av;
        ib[c] = 1.0f / (bv + 1e-9f);
    }
    VOC_TRY(voc_upload(h, al, &a.alpha));
    return voc_upload(h, ib, &a.invbeta);
}

static int voc_finalize(usb_vocoder* h) {
    if (h->finalized) return set_error("vocoder parameters already finalized");
    const usb_vocoder_config& c = h->cfg;
    VOC_CUDA(cudaSetDevice(c.device));
    h->n_stage = c.n_upsamples;
    h->nk = c.n_resblock_kernels;
    h->nd = c.n_dilations;
    h->ch[0] = c.upsample_initial_channel;
    for (int i = 0; i < h->n_stage; ++i) h->ch[i + 1] = c.upsample_initial_channel >> (i + 1);
    VOC_TRY(voc_load_conv(h, "conv_pre", h->ch[0], c.num_mels, 7, 1, h->conv_pre));
    h->ups.resize(h->n_stage);
    const int convs_per_layer = c.resblock_type == 1 ? 2 : 1;
    for (int i = 0; i < h->n_stage; ++i) {
        VOC_TRY(voc_load_convT(h, "ups." + std::to_string(i) + ".0", h->ch[i], h->ch[i + 1], c.upsample_kernel_sizes[i],
                               c.upsample_rates[i], h->ups[i]));
        for (int j = 0; j < h->nk; ++j) {
            const std::string pre = "resblocks." + std::to_string(i * h->nk + j);
            const int k = c.resblock_kernel_sizes[j];
            for (int l = 0; l < h->nd; ++l) {
                const int d = c.resblock_dilations[j][l];
                if (convs_per_layer == 2) {
                    VocConvW w1, w2;
                    VOC_TRY(voc_load_conv(h, pre + ".convs1." + std::to_string(l), h->ch[i + 1], h->ch[i + 1], k, d, w1));
                    VOC_TRY(voc_load_conv(h, pre + ".convs2." + std::to_string(l), h->ch[i + 1], h->ch[i + 1], k, 1, w2));
                    h->rb_convs.push_back(w1);
                    h->rb_convs.push_back(w2);
                } else {
                    VocConvW w1;
                    VOC_TRY(voc_load_conv(h, pre + ".convs." + std::to_string(l), h->ch[i + 1], h->ch[i + 1], k, d, w1));
                    h->rb_convs.push_back(w1);
                }
            }
            for (int l = 0; l < h->nd * convs_per_layer; ++l) {
                VocActW a;
                VOC_TRY(voc_load_act(h, pre + ".activations." + std::to_string(l), h->ch[i + 1], a));
                h->rb_acts.push_back(a);
            }
        }
    }
    const int cl = h->ch[h->n_stage];
    VOC_TRY(voc_load_act(h, "activation_post", cl, h->act_post));
    const VocParam *pw, *pb;
    VOC_TRY(voc_get(h, "conv_post.weight", {1, cl, 7}, &pw));
    VOC_TRY(voc_get(h, "conv_post.bias", {1}, &pb));
    h->post_cin8 = (cl + 7) / 8 * 8;
    std::vector<float> w(7 * h->post_cin8, 0.f);
    for (int ci = 0; ci < cl; ++ci)
        for (int k = 0; k < 7; ++k) w[k * h->post_cin8 + ci] = pw->data[ci * 7 + k];
    VOC_TRY(voc_upload(h, w, &h->post_w));
    VOC_TRY(voc_upload(h, pb->data, &h->post_b));
    kaiser_sinc_12(h->filt);
    VOC_CUDA(cudaFuncSetAttribute(snake_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kActSmemBytes));
    h->host.clear();
    h->finalized = true;
    return 0;
}

static double conv1d_flops(const VocConvW& w, int N, int L) {
    return 2.0 * N * L * w.phases * w.taps * static_cast<double>(w.Cin) * w.Cout;
}

static int voc_push_conv(usb_vocoder* h, const VocConvW& w, const __half* in, int N, int L, const __half* res, __half* out) {
    VocOp op;
    op.type = VocOp::CONV;
    VOC_TRY(build_conv1d(op.conv, w.dx, w.taps, w.phases, in, w.Cin, N, L, w.w, w.Cout, w.bias, res, out, w.cin_real));
    op.flops = conv1d_flops(w, N, L);
    h->flops_per_call += op.flops;
    h->ops.push_back(op);
    return 0;
}

static void voc_push_act(usb_vocoder* h, const VocActW& a, const __half* x, __half* out, int N, int L) {
    VocOp op;
    op.type = VocOp::ACT;
    op.N = N;
    op.act.x = x;
    op.act.out = out;
    op.act.alpha = a.alpha;
    op.act.invbeta = a.invbeta;
    op.act.L = L;
    op.act.C = a.C;
    op.act.Creal = a.Creal;
    op.act.vb = act_vb(a.Creal);
    memcpy(op.act.filt, h->filt, sizeof h->filt);
    h->ops.push_back(op);
}

static int voc_build_plan(usb_vocoder* h, int B, int T) {
    if (h->plan_B == B && h->plan_T == T) return 0;
    const usb_vocoder_config& c = h->cfg;
    VOC_CUDA(cudaDeviceSynchronize());
    if (h->arena) cudaFree(h->arena);
    h->arena = nullptr;
    h->ops.clear();
    h->flops_per_call = 0;
    h->plan_B = h->plan_T = 0;
    // buffer sizes: the largest [B][L][Cp] activation of any stage
    size_t max_elems = static_cast<size_t>(B) * T * pad64(h->ch[0]);
    {
        long long L = T;
        for (int i = 0; i < h->n_stage; ++i) {
            L *= c.upsample_rates[i];
            const size_t e = static_cast<size_t>(B) * L * pad64(h->ch[i + 1]);
            if (e > max_elems) max_elems = e;
        }
        if (L * static_cast<long long>(B) > 2000000000LL) return set_error("vocoder batch too long (B*T*hop must stay below 2^31)");
    }
    const int melC = h->conv_pre.Cin;
    const size_t mel_bytes = (static_cast<size_t>(B) * T * melC * 2 + 255) / 256 * 256;
    const size_t buf_bytes = (max_elems * 2 + 255) / 256 * 256;
    // nk <= 3: every resblock of a stage keeps its own fp16 output (R, R1, R2) and one pass averages them; otherwise an
    // fp32 accumulator (2 buffers' worth) is updated after every resblock
    const bool one_pass_mean = h->nk <= 3;
    h->arena_bytes = mel_bytes + 5 * buf_bytes + 2 * buf_bytes;
    VOC_CUDA(cudaMalloc(&h->arena, h->arena_bytes));
    char* base = static_cast<char*>(h->arena);
    h->melp = reinterpret_cast<__half*>(base);
    __half* P = reinterpret_cast<__half*>(base + mel_bytes);
    __half* X = reinterpret_cast<__half*>(base + mel_bytes + buf_bytes);
    __half* R = reinterpret_cast<__half*>(base + mel_bytes + 2 * buf_bytes);
    __half* A = reinterpret_cast<__half*>(base + mel_bytes + 3 * buf_bytes);
    __half* Cc = reinterpret_cast<__half*>(base + mel_bytes + 4 * buf_bytes);
    float* acc = reinterpret_cast<float*>(base + mel_bytes + 5 * buf_bytes);
    __half* Rj[3] = {R, reinterpret_cast<__half*>(base + mel_bytes + 5 * buf_bytes),
                     reinterpret_cast<__half*>(base + mel_bytes + 6 * buf_bytes)};      // (share the accumulator's space)

    int L = T;
    VOC_TRY(voc_push_conv(h, h->conv_pre, h->melp, B, L, nullptr, P));           // models.py:171
    const int cpl = c.resblock_type == 1 ? 2 : 1;
    size_t conv_i = 0, act_i = 0;
    for (int i = 0; i < h->n_stage; ++i) {
        VOC_TRY(voc_push_conv(h, h->ups[i], P, B, L, nullptr, X));                 // models.py:175-176
        L *= c.upsample_rates[i];
        const int Cp = pad64(h->ch[i + 1]);
        for (int j = 0; j < h->nk; ++j) {
            __half* Ro = one_pass_mean ? Rj[j] : R;      // this resblock's running / final output
            for (int l = 0; l < h->nd; ++l) {
                const __half* src = l == 0 ? X : Ro;
                if (cpl == 2) {   // AMPBlock1.forward, models.py:60-69
                    voc_push_act(h, h->rb_acts[act_i + 2 * l], src, A, B, L);
                    VOC_TRY(voc_push_conv(h, h->rb_convs[conv_i + 2 * l], A, B, L, nullptr, Cc));
                    voc_push_act(h, h->rb_acts[act_i + 2 * l + 1], Cc, A, B, L);
                    VOC_TRY(voc_push_conv(h, h->rb_convs[conv_i + 2 * l + 1], A, B, L, src, Ro));
                } else {          // AMPBlock2.forward, models.py:105-112
                    voc_push_act(h, h->rb_acts[act_i + l], src, A, B, L);
                    VOC_TRY(voc_push_conv(h, h->rb_convs[conv_i + l], A, B, L, src, Ro));
                }
            }
            conv_i += static_cast<size_t>(h->nd) * cpl;
            act_i += static_cast<size_t>(h->nd) * cpl;
            if (one_pass_mean && j + 1 < h->nk) continue;      // averaged together after the last resblock
            VocOp op;
            op.type = VocOp::ACCUM;
            op.r = R;
            op.acc = acc;
            op.out16 = P;
            op.n8 = static_cast<long long>(B) * L * Cp / 8;
            op.scale = 1.0f / h->nk;
            if (one_pass_mean) {
                op.mode = 4;
                op.r1 = h->nk > 1 ? Rj[1] : nullptr;
                op.r2 = h->nk > 2 ? Rj[2] : nullptr;
            } else {
                op.mode = j == 0 ? 0 : (j == h->nk - 1 ? 2 : 1);
            }
            h->ops.push_back(op);
        }
    }
    voc_push_act(h, h->act_post, P, A, B, L);                                     // models.py:187
    h->final_act = A;
    h->final_L = L;
    h->final_C = pad64(h->ch[h->n_stage]);
    h->plan_B = B;
    h->plan_T = T;
    return 0;
}

static int voc_forward(usb_vocoder* h, const float* mel, int B, int T, float* out, cudaStream_t s) {
    if (!h->finalized) return set_error("usb_vocoder_finalize_params has not been called");
    if (B < 1 || T < 1) return set_error("vocoder needs B >= 1 and T >= 1");
    VOC_CUDA(cudaSetDevice(h->cfg.device));
    VOC_TRY(voc_build_plan(h, B, T));
    {
        const int Cp = h->conv_pre.Cin;
        const long long n = static_cast<long long>(B) * (Cp / 8) * T;
        mel_pack_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
            mel, h->melp, B, h->cfg.num_mels, T, Cp, h->denorm ? h->mel_range : nullptr,
            h->denorm ? h->mel_range + h->cfg.num_mels : nullptr);
        h->launches++;
    }
    std::vector<cudaEvent_t> evs;
    if (h->profiling) {
        evs.resize(h->ops.size() + 1);
        for (cudaEvent_t& e : evs) VOC_CUDA(cudaEventCreate(&e));
        VOC_CUDA(cudaEventRecord(evs[0], s));
    }
    size_t op_i = 0;
    for (const VocOp& op : h->ops) {
        switch (op.type) {
            case VocOp::CONV: {
                const int e = launch_conv_igemm(op.conv.p, op.conv.a0, op.conv.a1, op.conv.b, op.conv.o, h->num_sms, s);
                if (e != 0) return set_error(std::string("vocoder conv launch: ") + cudaGetErrorString(static_cast<cudaError_t>(e)));
                break;
            }
            case VocOp::ACT: {
                launch_snake_act(op.act, op.N, s);
                break;
            }
            case VocOp::ACCUM:
                if (op.mode == 4)
                    stage_mean_kernel<<<static_cast<unsigned>((op.n8 + 255) / 256), 256, 0, s>>>(op.r, op.r1, op.r2, op.out16,
                                                                                               op.n8, op.scale);
                else
                    stage_accum_kernel<<<static_cast<unsigned>((op.n8 + 255) / 256), 256, 0, s>>>(op.r, op.acc, op.out16, op.n8,
                                                                                                op.mode, op.scale);
                break;
            default: break;
        }
        h->launches++;
        ++op_i;
        if (h->profiling) VOC_CUDA(cudaEventRecord(evs[op_i], s));
    }
    if (h->profiling) {
        VOC_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < h->ops.size(); ++i) {
            const VocOp& op = h->ops[i];
            float ms = 0.f;
            VOC_CUDA(cudaEventElapsedTime(&ms, evs[i], evs[i + 1]));
            const int cls = op.type == VocOp::CONV ? 0 : (op.type == VocOp::ACT ? 1 : 2);
            h->prof_ms[cls] += ms;
            h->prof_launches[cls]++;
            if (cls == 0) h->prof_work[0] += op.flops;
            else if (cls == 1) h->prof_work[1] += 4.0 * op.N * op.act.L * op.act.Creal;
            else h->prof_work[2] += static_cast<double>(op.n8) * 8 * (op.mode == 0 ? 6 : (op.mode == 1 ? 10 : (op.mode == 2 ? 8 : (op.mode == 4 ? 2 + 2 * (1 + (op.r1 != nullptr) + (op.r2 != nullptr)) : 4))));
        }
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
    }
    {
        const long long n = static_cast<long long>(B) * h->final_L;
        conv_post_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 7 * h->post_cin8 * sizeof(float), s>>>(
            h->final_act, h->post_w, h->post_b, out, B, h->final_L, h->final_C, h->post_cin8);
        h->launches++;
    }
    VOC_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace usb

// ===================================================================================================================
// C ABI
// ===================================================================================================================
extern "C" {

int usb_vocoder_create(const usb_vocoder_config* cfg, usb_vocoder** out) {
    if (!cfg || !out) return set_error("null argument");
    if (cfg->n_upsamples < 1 || cfg->n_upsamples > 8) return set_error("vocoder needs 1..8 upsample stages");
    if (cfg->n_resblock_kernels < 1 || cfg->n_resblock_kernels > 4) return set_error("vocoder needs 1..4 resblock kernel sizes");
    if (cfg->n_dilations < 1 || cfg->n_dilations > 4) return set_error("vocoder needs 1..4 dilations per resblock");
    if (cfg->resblock_type != 1 && cfg->resblock_type != 2) return set_error("resblock must be 1 or 2");
    if (cfg->activation != 0 && cfg->activation != 1) return set_error("activation must be 0 (snake) or 1 (snakebeta)");
    if (cfg->num_mels < 1 || cfg->upsample_initial_channel < (1 << cfg->n_upsamples) ||
        cfg->upsample_initial_channel % (1 << cfg->n_upsamples))
        return set_error("upsample_initial_channel must be divisible by 2^len(upsample_rates)");
    if ((cfg->upsample_initial_channel >> cfg->n_upsamples) > 64) return set_error("last stage wider than 64 channels is not supported");
    int ndev = 0;
    VOC_CUDA(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return set_error("no such CUDA device");
    VOC_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    VOC_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return set_error(std::string("this library only runs on sm_100 (B200); found ") + prop.name);
    usb_vocoder* h = new usb_vocoder();
    h->cfg = *cfg;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return 0;
}

void usb_vocoder_destroy(usb_vocoder* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    if (h->arena) cudaFree(h->arena);
    for (void* p : h->dev_allocs) cudaFree(p);
    delete h;
}

int usb_vocoder_load_param(usb_vocoder* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
    if (!h || !key || !data || (ndim > 0 && !shape)) return set_error("null argument");
    if (h->finalized) return set_error("vocoder parameters already finalized");
    VocParam p;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        p.shape.push_back(shape[i]);
        n *= static_cast<size_t>(shape[i]);
    }
    p.data.assign(data, data + n);
    h->host[key] = std::move(p);
    return 0;
}

int usb_vocoder_finalize_params(usb_vocoder* h) {
    if (!h) return set_error("null argument");
    return voc_finalize(h);
}

int usb_vocoder_forward(usb_vocoder* h, const float* mel, int32_t B, int32_t T, float* out, uint64_t stream) {
    if (!h || !mel || !out) return set_error("null argument");
    return voc_forward(h, mel, B, T, out, reinterpret_cast<cudaStream_t>(stream));
}

int usb_vocoder_forward_host(usb_vocoder* h, const float* mel_host, int32_t B, int32_t T, float* out_host) {
    if (!h || !mel_host || !out_host) return set_error("null argument");
    VOC_CUDA(cudaSetDevice(h->cfg.device));
    long long hop = 1;
    for (int i = 0; i < h->cfg.n_upsamples; ++i) hop *= h->cfg.upsample_rates[i];
    const size_t in_b = static_cast<size_t>(B) * h->cfg.num_mels * T * sizeof(float);
    const size_t out_b = static_cast<size_t>(B) * T * hop * sizeof(float);
    float *d_in = nullptr, *d_out = nullptr;
    VOC_CUDA(cudaMalloc(&d_in, in_b));
    if (cudaMalloc(&d_out, out_b) != cudaSuccess) {
        cudaFree(d_in);
        return set_error("cudaMalloc failed for the vocoder output");
    }
    int rc = 0;
    if (cudaMemcpy(d_in, mel_host, in_b, cudaMemcpyHostToDevice) != cudaSuccess) rc = set_error("H2D copy failed");
    if (!rc) rc = voc_forward(h, d_in, B, T, d_out, nullptr);
    if (!rc && cudaMemcpy(out_host, d_out, out_b, cudaMemcpyDeviceToHost) != cudaSuccess) rc = set_error("D2H copy failed");
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

long long usb_vocoder_launch_count(const usb_vocoder* h) { return h ? h->launches : 0; }
size_t usb_vocoder_workspace_bytes(const usb_vocoder* h) { return h ? h->arena_bytes : 0; }
double usb_vocoder_flops_per_call(const usb_vocoder* h) { return h ? h->flops_per_call : 0; }

int usb_vocoder_set_input_denorm(usb_vocoder* h, const float* mel_min_host, const float* mel_max_host) {
    if (!h) return set_error("null argument");
    if ((mel_min_host == nullptr) != (mel_max_host == nullptr)) return set_error("mel_min and mel_max must be given together");
    VOC_CUDA(cudaSetDevice(h->cfg.device));
    h->denorm = mel_min_host != nullptr;
    if (!h->denorm) return 0;
    const size_t n = sizeof(float) * h->cfg.num_mels;
    if (!h->mel_range) {
        void* p = nullptr;
        VOC_CUDA(cudaMalloc(&p, 2 * n));
        h->dev_allocs.push_back(p);
        h->mel_range = static_cast<float*>(p);
    }
    VOC_CUDA(cudaMemcpy(h->mel_range, mel_min_host, n, cudaMemcpyHostToDevice));
    VOC_CUDA(cudaMemcpy(h->mel_range + h->cfg.num_mels, mel_max_host, n, cudaMemcpyHostToDevice));
    return 0;
}

int usb_vocoder_set_profiling(usb_vocoder* h, int32_t on) {
    if (!h) return set_error("null argument");
    h->profiling = on != 0;
    for (int i = 0; i < 3; ++i) {
        h->prof_ms[i] = 0;
        h->prof_work[i] = 0;
        h->prof_launches[i] = 0;
    }
    return 0;
}

int usb_vocoder_get_profile(usb_vocoder* h, double* ms3, double* work3, long long* launches3) {
    if (!h || !ms3 || !work3 || !launches3) return set_error("null argument");
    for (int i = 0; i < 3; ++i) {
        ms3[i] = h->prof_ms[i];
        work3[i] = h->prof_work[i];
        launches3[i] = h->prof_launches[i];
    }
    return 0;
}

int usb_op_snake_act(const void* x, const float* alpha, const float* invbeta, int32_t N, int32_t L, int32_t C,
                     int32_t c_real, void* out, uint64_t stream) {
    if (!x || !alpha || !invbeta || !out) return set_error("null argument");
    if (C % 64 || N < 1 || L < 1 || c_real < 1 || c_real > C) return set_error("snake activation needs C % 64 == 0 and 1 <= c_real <= C");
    ActParams p;
    p.x = static_cast<const __half*>(x);
    p.out = static_cast<__half*>(out);
    p.alpha = alpha;
    p.invbeta = invbeta;
    p.L = L;
    p.C = C;
    p.Creal = c_real;
    p.vb = act_vb(c_real);
    kaiser_sinc_12(p.filt);
    VOC_CUDA(cudaFuncSetAttribute(snake_act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kActSmemBytes));
    launch_snake_act(p, N, reinterpret_cast<cudaStream_t>(stream));
    VOC_CUDA(cudaGetLastError());
    return 0;
}

int usb_op_conv1d(const void* x, const void* w, const float* bias, const void* res, void* out, int32_t N, int32_t L,
                  int32_t c_in, int32_t c_in_real, int32_t c_out, int32_t k, int32_t dilation, uint64_t stream) {
    if (!x || !w || !out) return set_error("null argument");
    if (c_in % 64 || c_out % 64 || N < 1 || L < 1 || c_in_real < 1 || c_in_real > c_in)
        return set_error("conv1d needs channel counts that are multiples of 64 and 1 <= c_in_real <= c_in");
    if (k < 1 || k > kConvMaxTaps || (k & 1) == 0 || dilation < 1) return set_error("conv1d needs an odd kernel size <= 16");
    const int pad = (k * dilation - dilation) / 2;
    if (pad > 127) return set_error("dilation too large");
    int8_t dx[kConvMaxTaps] = {0};
    for (int t = 0; t < k; ++t) dx[t] = static_cast<int8_t>(t * dilation - pad);
    int dev = 0, sms = 0;
    VOC_CUDA(cudaGetDevice(&dev));
    VOC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ConvOp op;
    VOC_TRY(build_conv1d(op, dx, k, 1, static_cast<const __half*>(x), c_in, N, L, static_cast<const __half*>(w), c_out, bias,
                         static_cast<const __half*>(res), static_cast<__half*>(out), c_in_real));
    const int rc = launch_conv_igemm(op.p, op.a0, op.a1, op.b, op.o, sms, reinterpret_cast<cudaStream_t>(stream));
    if (rc != 0) return set_error(std::string("conv1d launch failed: ") + cudaGetErrorString(static_cast<cudaError_t>(rc)));
    VOC_CUDA(cudaGetLastError());
    return 0;
}

int usb_vocoder_filter(float* out12) {
    if (!out12) return set_error("null argument");
    kaiser_sinc_12(out12);
    return 0;
}

}  // extern "C"
