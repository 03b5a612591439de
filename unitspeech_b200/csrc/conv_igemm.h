// Parameters of the generalised implicit-GEMM convolution kernel (conv_igemm.cu).
//
// One kernel covers every dense contraction of the U-Net except the 2-channel input conv:
//   3x3 stride-1 convs of Block (unitspeech/unitspeech.py:46-55), with an optional second K source for the
//   skip concat (:192), 3x3 stride-2 Downsample (:27-33), the 1x1 res_conv/to_qkv/to_out convs (:66,83-84)
//   (to_out with a per-sample folded weight), and ConvTranspose2d 4x4/s2 (:18-24) as four 2x2 phase convs.
//
// GEMM view: M = 128 pixels of one sample (a BH x BW patch of the "tile space"), N = BN output channels,
// K = taps x input channels, walked in 64-channel steps.  Activations are NHWC fp16; the A operand of a K step
// is one TMA box {64 ch, BW, 1, BH, 1} of a 5-D view (c', x, p, y, n) of the input tensor, shifted by the tap
// offset (out-of-bounds rows/columns are zero-filled by TMA = conv padding).  For stride-2 convs the view
// splits H and W by parity (p = row parity, column parity folded into c') so the box stays dense.
#pragma once
#include <cstdint>
#include <string>
#include <cuda.h>
#include <cuda_fp16.h>

namespace usb {

constexpr int kConvThreads = 384;      // warp0 TMA producer, warp1 MMA issuer, warp2 TMEM alloc, warps4-11 epilogue
constexpr int kConvEpilogueThreads = 256;
constexpr int kConvBM = 128;           // pixels per tile (UMMA M)
constexpr int kConvBK = 64;            // channels per K step (128 B of fp16 = one swizzle row)
constexpr int kConvMaxTaps = 16;       // phases x taps
constexpr int kConvSmemBytes = 225 * 1024;
// GroupNorm partial sums are accumulated with integer atomics in fixed point: integer addition is
// associative, so the statistics (and with them every fp16 rounding downstream) are bit-reproducible from
// run to run and independent of the batch an utterance is in.
constexpr float kStatSumScale = 16777216.f;   // 2^24
constexpr float kStatSqScale = 262144.f;      // 2^18

struct ConvTap {
    int16_t c;    // offset added to the channel coordinate (column-parity * C for stride-2 views)
    int8_t dx;    // offset added to the x coordinate
    int8_t dy;    // offset added to the y coordinate
    int8_t p;     // coordinate in the parity dimension
    int8_t pad[3];
};

struct ConvParams {
    // tile space (per sample): Hm x Wm pixels in BH x BW patches, BH*BW == 128, BH divides Hm
    int N, Hm, Wm, BH, BW, tiles_y, tiles_x;
    int Cout, BN, n_tiles_n;
    int phases, taps;          // taps per phase; tap table indexed [phase*taps + t]
    int chunks0, chunks1;      // 64-channel K chunks per tap from source 0 / source 1 (0 = no second source)
    int b_batch_mode;          // B z-coordinate: 0 -> 0, 1 -> phase, 2 -> sample
    int stages;                // smem pipeline depth
    int swap_ab;               // 1: swapped-operand kernel: channels on the MMA M axis, two 128-pixel patches on N
    int patches_per_phase;     // N * tiles_y * tiles_x
    int dbg_flags;             // experiments only: 1 = skip B loads, 2 = skip epilogue math/stores, 4 = skip A loads
    // halo-reuse kernel (3x3 stride-1, Hm % 8 == 0, Cout % 128 == 0): tiles of 8 rows x 32 columns, the activation
    // tile is loaded once per 64-channel chunk with its 1-pixel halo and the nine taps are descriptor offsets into it
    int halo;                  // 1: use conv_igemm_halo_kernel (tiles_y = Hm/8, tiles_x = ceil(Wm/32))
    int halo_hy;               // rows per column of the shared-memory halo tile: 10 (dense) or 16 (power-of-two stride)
    int halo_boff;             // 1: put (start_address >> 7) & 7 into the descriptor's base-offset field
    // 1-D halo kernel (conv1d_halo_kernel; Hm == 1, one phase, one source): per 64-channel chunk the activation tile is loaded
    // ONCE as a box of h1d_rows = 128 + (dx_max - dx_min) (rounded up to 8) positions starting at x0 + h1d_dx0, and the
    // taps are descriptor offsets of (dx - h1d_dx0) rows into it
    int h1d;                   // 1: use conv1d_halo_kernel (map_a1 = the haloed activation box)
    int h1d_rows, h1d_dx0;
    int h1d_na;                // activation buffers in the ring (2..4)
    int h1d_wres;              // 1: all taps x chunks weight tiles stay resident in shared memory (loaded once per CTA)
    int h1d_klast;             // K = 16 slices (1..4) of the LAST 64-channel chunk that hold real channels; the rest is padding
    ConvTap tap[kConvMaxTaps];
    // epilogue: v = acc + bias[c]; stats (sum, sumsq per (n, group)) on v; v = v*res_scale + res; v *= mask
    const float* bias;         // [Cout] or null
    long long* stats;          // [N][groups][2] fixed-point (sum * 2^24, sumsq * 2^18) or null
    int groups;                // GroupNorm groups (stats only)
    const __half* res;         // residual, same geometry as out, or null
    const float* res_scale;    // device scalar (Rezero g) or null (=1)
    const float* mask;         // [N][mask_stride] indexed by the OUTPUT x coordinate, or null
    int mask_stride;
    // the output tile itself is written by TMA (map_out: same 5-D view family as the inputs; for the transposed
    // conv the parity view of the 2H x 2W output, channel coordinate += ox_off*out_c_phase_mul, p = oy_off);
    // the strides below address the residual, which shares the output geometry
    long long o_sn, o_sy, o_sx;
    int out_c_phase_mul;
    int oy_mul, ox_mul;        // yo = y*oy_mul + oy_off[phase], xo = x*ox_mul + ox_off[phase]
    int8_t oy_off[4], ox_off[4];
    // fp16 saturation report: incremented once per epilogue thread and tile in which a value reached the
    // fp16 limit (|v| >= 65504 before the satfinite pack); null = not reported
    unsigned long long* sat;
    // split-K (swapped-operand kernel, small latency-bound calls only): a tile's K range is cut into `ksplit` work items
    // run by different CTAs; each writes its fp32 partial tile to `kpart`, takes a ticket, and the CTA that arrives last
    // sums the partials IN SPLIT ORDER (so the result does not depend on which CTA that is) and runs the epilogue.
    // ksplit is a function of the layer geometry only, never of the batch (bitwise batch invariance).
    int ksplit;                // 0 / 1 = off
    float* kpart;              // [tiles * ksplit][2 patches][128 px][128 ch] fp32
    int* ktick;                // [tiles], zero between launches (the reducing CTA resets its tile's ticket)
};

struct ConvOp {
    ConvParams p;
    CUtensorMap a0, a1, b, o;
};

// engine.cu: records the thread's last error message (usb_last_error) and returns 1
int set_error(const std::string& m);
}  // namespace usb
struct usb_handle;
namespace usb {
// engine.cu: CUDA device ordinal the handle is bound to
int handle_device(const usb_handle* h);
// engine.cu: parameter block + tensor maps of a 1-D (transposed) convolution over NLC fp16 tensors
// (cin_real: input channels that carry data, <= Cin; the K slices of pure padding are skipped)
int build_conv1d(ConvOp& op, const int8_t* dx, int taps, int phases, const __half* in, int Cin, int N, int L,
                 const __half* w, int Cout, const float* bias, const __half* res, __half* out, int cin_real = 0);

// launches on `stream`; returns cudaError_t as int
int launch_conv_igemm(const ConvParams& p, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                      const CUtensorMap& out, int num_sms, cudaStream_t stream);

}  // namespace usb
