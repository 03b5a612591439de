// Host runtime + C ABI of libunitspeech_b200.so: parameter repacking, workspace/plan construction (buffers, TMA
// tensor maps, kernel parameter blocks) and the launch sequences of the score estimator and the sampler loop.
// Reference semantics followed here: GradLogPEstimator2d.forward (unitspeech/unitspeech.py:164-201),
// classifier_free_guidance (:298-331), reverse_diffusion (:334-374).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/unitspeech_b200.h"
#include "../../include/unitspeech_b200_train.h"
#include "conv_igemm.h"
#include "kernels.h"
#include "pdl.h"
#include "train.h"

namespace usb {

static thread_local std::string g_err;
static int fail(const std::string& m) {
    g_err = m;
    return 1;
}
int set_error(const std::string& m) { return fail(m); }
static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }
struct PdlScope {      // programmatic dependent launch for the launches issued inside the scope (pdl.h)
    bool prev;
    explicit PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
    ~PdlScope() { g_pdl = prev; }
};
#define USB_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +            \
                        std::to_string(__LINE__) + ")");                                                     \
    } while (0)
#define USB_LAUNCH(h, expr)                                                                                  \
    do {                                                                                                     \
        int _e = (expr);                                                                                     \
        (h)->launches++;                                                                                     \
        if (_e != 0)                                                                                         \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(static_cast<cudaError_t>(_e)) +       \
                        " (" __FILE__ ":" + std::to_string(__LINE__) + ")");                                 \
    } while (0)
#define USB_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != 0) return _r;    \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// TMA tensor-map encoding through the driver entry point (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode_fn() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    USB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || fn == nullptr) return fail("cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
                 box[1], box[2]);
        return fail(buf);
    }
    return 0;
}

// 5-D activation view (c', x, p, y, n) of an NHWC fp16 tensor [N][H][W][Ctot], channels [0, Cview) visible.
// stride2: parity split for stride-2 convs: c' = (x&1)*Ctot + c, x' = x/2, p = y&1, y' = y/2.
static int make_act_map(CUtensorMap* m, const __half* base, int N, int H, int W, int Ctot, int Cview, bool stride2,
                        int BH, int BW, int box_c = 64) {
    cuuint64_t dims[5], strides[4];
    const cuuint64_t e = 2;
    if (!stride2) {
        dims[0] = Cview; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
        strides[0] = Ctot * e; strides[1] = (cuuint64_t)W * Ctot * e; strides[2] = (cuuint64_t)W * Ctot * e;
        strides[3] = (cuuint64_t)H * W * Ctot * e;
    } else {
        dims[0] = 2 * Ctot; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
        strides[0] = 2 * Ctot * e; strides[1] = (cuuint64_t)W * Ctot * e; strides[2] = (cuuint64_t)2 * W * Ctot * e;
        strides[3] = (cuuint64_t)H * W * Ctot * e;
    }
    cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)BW, 1, (cuuint32_t)BH, 1};
    return encode_map(m, base, 5, dims, strides, box,
                      box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// 4-D activation view (c, y, x, n) of the same tensor: rows before columns, so a box lands in shared memory column
// by column (conv_igemm_halo_kernel)
static int make_act_map_yx(CUtensorMap* m, const __half* base, int N, int H, int W, int Ctot, int Cview, int box_c,
                           int box_y, int box_x) {
    const cuuint64_t e = 2;
    cuuint64_t dims[4] = {(cuuint64_t)Cview, (cuuint64_t)H, (cuuint64_t)W, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)W * Ctot * e, (cuuint64_t)Ctot * e, (cuuint64_t)H * W * Ctot * e};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_y, (cuuint32_t)box_x, 1};
    return encode_map(m, base, 4, dims, strides, box,
                      box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// weights [Z][Cout][K] fp16, K contiguous
static int make_w_map(CUtensorMap* m, const __half* base, int Z, int Cout, int K, int BN) {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Cout, (cuuint64_t)Z};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)Cout * K * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)BN, 1};
    return encode_map(m, base, 3, dims, strides, box);
}

// ---------------------------------------------------------------------------------------------------------------
// weights
// ---------------------------------------------------------------------------------------------------------------
// K3S2D / KT4D are the data-gradient convolutions of the fine-tune step: the transposed 3x3/s2 conv as four phase convs of
// 2x2 (zero-padded) taps, and the 4x4/s2 conv over the output gradient of the ConvTranspose (train.h: launch_pack_conv)
enum ConvKind { K3S1 = 0, K3S2 = 1, K1 = 2, KT4 = 3, K3S2D = 4, KT4D = 5 };
static inline bool kind_up(int kind) { return kind == KT4 || kind == K3S2D; }      // output 2H x 2W, four phases
static inline bool kind_down(int kind) { return kind == K3S2 || kind == KT4D; }    // parity-split input, output H/2 x W/2
constexpr int kAttnChunk = 2048;   // positions per attention partial block (upper bound)
// auto graph mode: (CFG branches x utterances) x frames at or below this replays a captured sampler step (covers one
// utterance of any practical length with CFG, or 16 utterances x 256 frames); larger workloads are GPU-bound already
constexpr long long kGraphAutoRowsFrames = 12288;
// auto split-K mode: one utterance with text + speaker guidance up to 1024 frames (or the same rows x frames otherwise).
// Narrower than the graph threshold because split-K changes the fp32 summation order: calls on either side of the
// threshold agree to ~1e-4, not bitwise (the Python layer decides per JOB, so micro-batches and shards of one job agree)
constexpr long long kSplitKAutoRowsFrames = 3072;
// positions per partial block: ~16 chunks per sample so small levels still fill the GPU, multiples of 64
static inline int attn_chunk_for(int P) {
    int c = ((P / 16 + 63) / 64) * 64;
    return c < 256 ? 256 : (c > kAttnChunk ? kAttnChunk : c);
}

struct ConvW {
    __half* w = nullptr;   // [Z][Cout][K]
    float* bias = nullptr; // [Cout]
    int Cout = 0, Cin = 0, K = 0, Z = 1, kind = K3S1;
};

static inline int taps_of(int kind) {
    return kind == K3S1 || kind == K3S2 ? 9 : (kind == K1 ? 1 : (kind == KT4D ? 16 : 4));
}

// reference layout -> [Z][Cout][tap*Cin + ci] fp16
static void pack_conv_host(int kind, const float* w, int Cout, int Cin, std::vector<__half>& out) {
    if (kind == K3S1 || kind == K3S2) {
        out.resize((size_t)Cout * 9 * Cin);
        for (int co = 0; co < Cout; ++co)
            for (int ci = 0; ci < Cin; ++ci)
                for (int t = 0; t < 9; ++t)
                    out[((size_t)co * 9 + t) * Cin + ci] = __float2half_rn(w[((size_t)co * Cin + ci) * 9 + t]);
    } else if (kind == K1) {
        out.resize((size_t)Cout * Cin);
        for (size_t i = 0; i < out.size(); ++i) out[i] = __float2half_rn(w[i]);
    } else {  // ConvTranspose2d weight (Cin, Cout, 4, 4): out[2a+ph, 2b+pw] phases, 2x2 taps each
        out.resize((size_t)4 * Cout * 4 * Cin);
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                for (int co = 0; co < Cout; ++co)
                    for (int a = 0; a < 2; ++a)
                        for (int b = 0; b < 2; ++b) {
                            // ho = 2*hi - 1 + kh: phase 0 uses kh = 1 (hi = a'), 3 (hi = a'-1); phase 1 uses kh = 0 (hi = a'+1), 2
                            const int kh = ph == 0 ? (a == 0 ? 1 : 3) : (a == 0 ? 0 : 2);
                            const int kw = pw == 0 ? (b == 0 ? 1 : 3) : (b == 0 ? 0 : 2);
                            for (int ci = 0; ci < Cin; ++ci)
                                out[(((size_t)(ph * 2 + pw) * Cout + co) * 4 + (a * 2 + b)) * Cin + ci] =
                                    __float2half_rn(w[(((size_t)ci * Cout + co) * 4 + kh) * 4 + kw]);
                        }
    }
}

static void fill_taps(int kind, int Ctot0, ConvParams& p) {
    p.taps = taps_of(kind);
    p.phases = kind_up(kind) ? 4 : 1;
    memset(p.tap, 0, sizeof p.tap);
    if (kind == K3S1) {
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                ConvTap& t = p.tap[kh * 3 + kw];
                t.dy = (int8_t)(kh - 1);
                t.dx = (int8_t)(kw - 1);
            }
    } else if (kind == K3S2) {
        // input row 2*yo + kh - 1: kh=0 -> odd row yo-1; kh=1 -> even row yo; kh=2 -> odd row yo (same for columns)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                ConvTap& t = p.tap[kh * 3 + kw];
                t.dy = (int8_t)(kh == 0 ? -1 : 0);
                t.p = (int8_t)(kh == 1 ? 0 : 1);
                t.dx = (int8_t)(kw == 0 ? -1 : 0);
                t.c = (int16_t)(kw == 1 ? 0 : Ctot0);
            }
    } else if (kind == KT4) {
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        ConvTap& t = p.tap[(ph * 2 + pw) * 4 + a * 2 + b];
                        t.dy = (int8_t)(ph == 0 ? (a == 0 ? 0 : -1) : (a == 0 ? 1 : 0));
                        t.dx = (int8_t)(pw == 0 ? (b == 0 ? 0 : -1) : (b == 0 ? 1 : 0));
                    }
    } else if (kind == K3S2D) {
        // d_in[2a'+ph] = sum over kh with 2*yo + kh - 1 = 2a'+ph: ph 0 -> kh 1 (yo = a'); ph 1 -> kh 0 (yo = a'+1), kh 2 (yo = a')
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 2; ++b) {
                        ConvTap& t = p.tap[(ph * 2 + pw) * 4 + a * 2 + b];
                        t.dy = (int8_t)(ph == 1 && a == 0 ? 1 : 0);
                        t.dx = (int8_t)(pw == 1 && b == 0 ? 1 : 0);
                    }
    } else if (kind == KT4D) {
        // input row 2*yo + kh - 1 of the 2H-row gradient: kh 0 -> odd row yo-1; 1 -> even row yo; 2 -> odd row yo; 3 -> even row yo+1
        for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
                ConvTap& t = p.tap[kh * 4 + kw];
                t.dy = (int8_t)(kh == 0 ? -1 : (kh == 3 ? 1 : 0));
                t.p = (int8_t)((kh & 1) ? 0 : 1);
                t.dx = (int8_t)(kw == 0 ? -1 : (kw == 3 ? 1 : 0));
                t.c = (int16_t)((kw & 1) ? 0 : Ctot0);
            }
    }
}

struct ConvEpilogue {
    const float* bias = nullptr;
    long long* stats = nullptr;
    int groups = 8;
    const __half* res = nullptr;
    const float* res_scale = nullptr;
    const float* mask = nullptr;  // [N][Wout]
    unsigned long long* sat = nullptr;
    // split-K (latency mode, see ConvParams::ksplit): workspace of `kpart_items` partial tiles + ticket array
    bool splitk = false;
    float* kpart = nullptr;
    int* ktick = nullptr;
    long long kpart_items = 0;
    int num_sms = 148;
    int split_rows = 3;           // rows the split factor is derived for (3: sampler rule below; 0: the launch's own N)
};

// Split factor of a swapped-kernel conv in latency mode: a function of the per-sample geometry only -- the tile count is
// taken AS IF the call had 3 rows (one utterance with text + speaker guidance) whatever the batch is, so an utterance is
// reduced in the same order alone or inside a batch.  A split costs ~8 us (fp32 partial tile through L2, ticket, ordered
// reduction), i.e. about 24 K steps of MMA time: every split keeps at least that much work, at most 8 splits.
static int choose_ksplit(const ConvParams& p, int num_sms, int rows) {
    const long long patches3 = (long long)(rows > 0 ? rows : p.N) * p.tiles_y * p.tiles_x;
    const long long tiles3 = (long long)p.phases * ((patches3 + 1) / 2) * p.n_tiles_n;
    const int ksteps = p.taps * (p.chunks0 + p.chunks1);
    int ks = (int)(num_sms / (tiles3 > 0 ? tiles3 : 1));
    if (ks > 8) ks = 8;
    if (ks > ksteps / 24) ks = ksteps / 24;
    return ks < 1 ? 1 : ks;
}

// Builds the parameter block + tensor maps of one convolution launch.
// in0/in1: NHWC fp16 [N][H][W][Ctot*] of which the first C* channels are contracted; out: [N][Hout][Wout][Cout].
static int build_conv(ConvOp& op, int kind, const __half* in0, int C0tot, int C0, const __half* in1, int C1tot, int C1,
                      int N, int H, int W, const __half* wptr, int wZ, int b_batch_mode, int Cout,
                      const ConvEpilogue& ep, __half* out) {
    USB_TRY(load_encode_fn());
    if (C0 % 64 || C1 % 64 || Cout % 64) return fail("conv channels must be multiples of 64");
    if (kind_down(kind) && (H % 2 || W % 2 || in1 != nullptr || C0 != C0tot)) return fail("bad stride-2 conv geometry");
    if (kind_down(kind) && 2 * C0tot > 32767) return fail("stride-2 conv too wide");
    ConvParams& p = op.p;
    memset(&p, 0, sizeof p);
    fill_taps(kind, C0tot, p);
    p.N = N;
    p.Hm = kind_down(kind) ? H / 2 : H;
    p.Wm = kind_down(kind) ? W / 2 : W;
    int BH = 1;
    while (BH < 16 && p.Hm % (BH * 2) == 0) BH *= 2;
    // swapped-operand kernel (see conv_igemm.cu) for every 3x3 / strided / transposed conv whose Cout is a multiple of 128
    static const bool allow_swap = getenv("USB_NO_SWAP_AB") == nullptr;
    static const bool swap_k1 = getenv("USB_SWAP_K1") != nullptr;
    p.swap_ab = (allow_swap && Cout % 128 == 0 && (kind != K1 || swap_k1) && ep.res == nullptr && b_batch_mode != 2) ? 1 : 0;
    // The swapped-operand kernel masks rows as well as columns, so its patch height need not divide the image height:
    // pick the 128-pixel patch shape that wastes the fewest pixels (e.g. a 10 x 22 level-3 image: 4 x 32 patches cover
    // it at 57 % instead of 34 % for 2 x 64); keep the dividing shape unless the gain is above 5 %.
    const bool halo_candidate = kind == K3S1 && H % 8 == 0 && b_batch_mode == 0 && getenv("USB_NO_HALO") == nullptr;
    if (p.swap_ab && !halo_candidate) {
        auto eff = [&](int bh) {
            const int bw = 128 / bh;
            return (double)p.Hm * p.Wm / ((double)((p.Hm + bh - 1) / bh) * bh * ((p.Wm + bw - 1) / bw) * bw);
        };
        double best = eff(BH) * 1.05;
        for (int bh = 1; bh <= 16; bh *= 2)
            if (eff(bh) > best) { best = eff(bh); BH = bh; }
    }
    p.BH = BH;
    p.BW = 128 / BH;
    p.tiles_y = (p.Hm + BH - 1) / BH;
    p.tiles_x = (p.Wm + p.BW - 1) / p.BW;
    p.patches_per_phase = N * p.tiles_y * p.tiles_x;
    p.Cout = Cout;
    p.BN = p.swap_ab ? 128 : (Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64));
    p.n_tiles_n = Cout / p.BN;
    p.chunks0 = C0 / 64;
    p.chunks1 = in1 ? C1 / 64 : 0;
    p.b_batch_mode = b_batch_mode;
    // pipeline stages + epilogue staging (swapped: 2 x 16 KB, plain: 8 warps x 2 x 4 KB) must fit 225 KB
    p.stages = p.swap_ab ? 4 : (p.BN == 256 ? 3 : (p.BN == 128 ? 5 : 6));
    if (const char* e = getenv("USB_DBG_STAGES")) p.stages = atoi(e);
    if (const char* e = getenv("USB_DBG_FLAGS")) p.dbg_flags = atoi(e);
    if (const char* e = getenv("USB_DBG_BH")) {
        const int bh = atoi(e);
        if (bh > 0 && p.Hm % bh == 0 && 128 % bh == 0) {
            p.BH = bh; p.BW = 128 / bh; p.tiles_y = p.Hm / bh; p.tiles_x = (p.Wm + p.BW - 1) / p.BW;
        }
    }
    p.bias = ep.bias;
    p.stats = ep.stats;
    p.groups = ep.groups;
    if (ep.stats) {
        const int cpg = Cout / ep.groups;
        if (Cout % ep.groups || !(cpg == 8 || cpg == 16 || cpg % 32 == 0)) return fail("unsupported GroupNorm width");
    }
    p.res = ep.res;
    p.res_scale = ep.res_scale;
    p.mask = ep.mask;
    p.sat = ep.sat;
    const int Hout = kind_down(kind) ? H / 2 : (kind_up(kind) ? 2 * H : H);
    const int Wout = kind_down(kind) ? W / 2 : (kind_up(kind) ? 2 * W : W);
    p.mask_stride = Wout;
    p.out_c_phase_mul = kind_up(kind) ? Cout : 0;
    p.o_sx = Cout;
    p.o_sy = (long long)Wout * Cout;
    p.o_sn = (long long)Hout * Wout * Cout;
    p.oy_mul = p.ox_mul = kind_up(kind) ? 2 : 1;
    if (kind_up(kind))
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                p.oy_off[ph * 2 + pw] = (int8_t)ph;
                p.ox_off[ph * 2 + pw] = (int8_t)pw;
            }
    const int K = p.taps * (C0 + (in1 ? C1 : 0));
    // halo-reuse kernel for the 3x3 stride-1 convs of the levels whose height is a multiple of 8 (levels 0 and 1)
    // (USB_NO_HALO=1 falls back to the per-tap loads of the swapped kernel; USB_HALO_HY / USB_HALO_BOFF are the
    // experiment knobs that established the layout: dense 10-row columns, descriptor base-offset field left 0)
    static const bool halo_mode = getenv("USB_NO_HALO") == nullptr;
    if (halo_mode && p.swap_ab && kind == K3S1 && H % 8 == 0 && b_batch_mode == 0) {
        p.halo = 1;
        p.halo_hy = getenv("USB_HALO_HY") ? atoi(getenv("USB_HALO_HY")) : 10;
        if (p.halo_hy != 10 && p.halo_hy != 16) return fail("USB_HALO_HY must be 10 or 16");
        p.halo_boff = getenv("USB_HALO_BOFF") ? atoi(getenv("USB_HALO_BOFF")) : 0;
        p.stages = p.halo_hy == 10 ? 7 : 4;
        if (const char* e = getenv("USB_DBG_STAGES")) p.stages = atoi(e);
        p.tiles_y = H / 8;
        p.tiles_x = (W + 31) / 32;
        USB_TRY(make_act_map_yx(&op.a0, in0, N, H, W, C0tot, C0, 64, p.halo_hy, 34));
        if (in1) USB_TRY(make_act_map_yx(&op.a1, in1, N, H, W, C1tot, C1, 64, p.halo_hy, 34));
        else op.a1 = op.a0;
        USB_TRY(make_w_map(&op.b, wptr, wZ, Cout, K, 128));
        return make_act_map_yx(&op.o, out, N, H, W, Cout, Cout, 32, 8, 4);
    }
    if (p.swap_ab && ep.splitk) {
        // the decision depends on the geometry only (choose_ksplit); a caller with a fixed-size workspace (the fine-tune
        // entry) passes its capacity, the sampler plan sizes its workspace afterwards (conv_split_items)
        const int ks = choose_ksplit(p, ep.num_sms, ep.split_rows);
        const long long items = (long long)p.phases * ((p.patches_per_phase + 1) / 2) * p.n_tiles_n * ks;
        if (ks > 1 && (ep.kpart == nullptr || items <= ep.kpart_items)) {
            p.ksplit = ks;
            p.kpart = ep.kpart;
            p.ktick = ep.ktick;
        }
    }
    USB_TRY(make_act_map(&op.a0, in0, N, H, W, C0tot, C0, kind_down(kind), p.BH, p.BW));
    if (in1) USB_TRY(make_act_map(&op.a1, in1, N, H, W, C1tot, C1, false, p.BH, p.BW));
    else op.a1 = op.a0;
    USB_TRY(make_w_map(&op.b, wptr, wZ, Cout, K, p.BN));
    // output tile store: plain view, or the parity view of the upsampled tensor for the transposed conv;
    // the swapped kernel stores 64-pixel sub-blocks (64 / BW image rows) per TMA
    if (p.swap_ab)   // per-warp stores: 64 pixels x 32 channels (64-byte rows, SWIZZLE_64B)
        USB_TRY(make_act_map(&op.o, out, N, Hout, Wout, Cout, Cout, kind_up(kind), p.BW >= 64 ? 1 : 64 / p.BW,
                             p.BW >= 64 ? 64 : p.BW, 32));
    else   // per-warp stores: 32 pixels (one TMEM lane quarter) x 64 channels
        USB_TRY(make_act_map(&op.o, out, N, Hout, Wout, Cout, Cout, kind_up(kind), p.BW >= 32 ? 1 : 32 / p.BW,
                             p.BW >= 32 ? 32 : p.BW));
    return 0;
}

// 1-D convolution / transposed convolution (vocoder.cu) on the same kernel: NLC fp16 tensors are NHWC with H = 1.
// taps x-offsets are explicit (dilated kernels); phases > 1 = ConvTranspose1d with stride `phases`, whose output
// [N][phases*L][Cout] is addressed as [N][L][phases*Cout] (phase ph lands in channels [ph*Cout, (ph+1)*Cout)).
int build_conv1d(ConvOp& op, const int8_t* dx, int taps, int phases, const __half* in, int Cin, int N, int L,
                 const __half* w, int Cout, const float* bias, const __half* res, __half* out, int cin_real) {
    USB_TRY(load_encode_fn());
    if (Cin % 64 || Cout % 64) return fail("conv channels must be multiples of 64");
    if (phases < 1 || phases > 4 || taps < 1 || phases * taps > kConvMaxTaps) return fail("unsupported 1-D conv tap table");
    if (res != nullptr && phases != 1) return fail("residual epilogue needs a plain conv");
    ConvParams& p = op.p;
    memset(&p, 0, sizeof p);
    p.taps = taps;
    p.phases = phases;
    for (int i = 0; i < phases * taps; ++i) p.tap[i].dx = dx[i];
    p.N = N;
    p.Hm = 1;
    p.Wm = L;
    static const bool allow_swap = getenv("USB_NO_SWAP_AB") == nullptr;
    // USB_H1D: 0 = never use the 1-D halo kernel, 2 = also where the swapped-operand kernel applies (Cout <= 256)
    static const int h1d_mode = getenv("USB_H1D") ? atoi(getenv("USB_H1D")) : 1;
    const bool swap_ok = allow_swap && Cout % 128 == 0 && res == nullptr;
    const bool h1d_ok = h1d_mode != 0 && phases == 1 && taps >= 2;
    p.h1d = (h1d_ok && (!swap_ok || (h1d_mode == 2 && Cout <= 256))) ? 1 : 0;
    p.swap_ab = (swap_ok && !p.h1d) ? 1 : 0;
    p.BH = 1;
    p.BW = 128;
    p.tiles_y = 1;
    p.tiles_x = (L + 127) / 128;
    p.patches_per_phase = N * p.tiles_x;
    p.Cout = Cout;
    if (p.h1d) p.BN = Cout % 256 == 0 ? 256 : (Cout % 192 == 0 ? 192 : (Cout % 128 == 0 ? 128 : 64));
    else p.BN = p.swap_ab ? 128 : (Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64));
    p.n_tiles_n = Cout / p.BN;
    p.chunks0 = Cin / 64;
    p.chunks1 = 0;
    p.b_batch_mode = phases > 1 ? 1 : 0;
    p.stages = p.swap_ab ? 4 : (p.BN == 256 ? 3 : (p.BN == 128 ? 5 : 6));
    p.bias = bias;
    p.groups = 8;
    p.res = res;
    p.mask_stride = L;
    p.out_c_phase_mul = phases > 1 ? Cout : 0;
    p.o_sx = Cout;
    p.o_sy = (long long)L * Cout;
    p.o_sn = (long long)L * Cout;
    p.oy_mul = p.ox_mul = 1;
    for (int ph = 0; ph < phases; ++ph) p.ox_off[ph] = (int8_t)(phases > 1 ? ph : 0);
    USB_TRY(make_act_map(&op.a0, in, N, 1, L, Cin, Cin, false, 1, 128));
    op.a1 = op.a0;
    if (p.h1d) {
        int dmin = dx[0], dmax = dx[0];
        for (int t = 1; t < taps; ++t) {
            dmin = dx[t] < dmin ? dx[t] : dmin;
            dmax = dx[t] > dmax ? dx[t] : dmax;
        }
        p.h1d_dx0 = dmin;
        p.h1d_rows = (128 + (dmax - dmin) + 7) / 8 * 8;
        if (p.h1d_rows > 256) return fail("1-D conv taps reach further than 128 positions");
        const long long budget = kConvSmemBytes - 1024 - 8 * 8192;
        const long long a_stride = ((long long)p.h1d_rows * 128 + 1023) / 1024 * 1024;
        const long long b_bytes = (long long)p.BN * 128;
        const long long w_all = (long long)taps * p.chunks0 * b_bytes;
        if (p.n_tiles_n == 1 && w_all + 2 * a_stride <= budget && getenv("USB_H1D_NO_WRES") == nullptr) {
            p.h1d_wres = 1;
            long long na = (budget - w_all) / a_stride;
            p.h1d_na = (int)(na > 4 ? 4 : na);
            p.stages = 1;
        } else {
            p.h1d_na = p.BN >= 192 ? 2 : 3;
            long long st = (budget - p.h1d_na * a_stride) / b_bytes;
            if (st < 2) return fail("1-D halo conv does not fit in shared memory");
            p.stages = (int)(st > 8 ? 8 : st);
        }
        const int last_real = (cin_real > 0 ? cin_real : Cin) - (p.chunks0 - 1) * 64;
        p.h1d_klast = last_real <= 0 ? 1 : (last_real >= 64 ? 4 : (last_real + 15) / 16);
        USB_TRY(make_act_map(&op.a1, in, N, 1, L, Cin, Cin, false, 1, p.h1d_rows));
    }
    if (const char* e = getenv("USB_DBG_FLAGS")) p.dbg_flags = atoi(e);
    USB_TRY(make_w_map(&op.b, w, phases, Cout, taps * Cin, p.BN));
    if (p.swap_ab) USB_TRY(make_act_map(&op.o, out, N, 1, L, phases * Cout, phases * Cout, false, 1, 64, 32));
    else USB_TRY(make_act_map(&op.o, out, N, 1, L, phases * Cout, phases * Cout, false, 1, 32));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------
struct HostParam {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

struct GnW {
    float* gamma = nullptr;
    float* beta = nullptr;
};
struct ResnetW {
    std::string prefix;
    int Cin = 0, Cout = 0, emb_off = 0;
    bool has_res = false;
    ConvW c1, c2, res;
    GnW g1, g2;
};
struct AttnW {
    std::string prefix;
    int C = 0;
    bool fused_q = false;  // C <= 256: q is never materialised (see AttnParams)
    ConvW kv;              // fused_q: rows [hidden, 3*hidden) of to_qkv
    float* wq = nullptr;   // fused_q: rows [0, hidden) of to_qkv, fp32 [hidden][C]
    float* bprime = nullptr;
    ConvW qkv;
    float* wo = nullptr;  // [C][128]
    float* bo = nullptr;  // [C]
    float* g = nullptr;   // [1]
};

struct Op {
    enum Kind { FIRST, CONV, GN, ATTN } kind;
    int idx;
};

struct Plan {
    int Be = 0, T = 0;
    bool splitk = false;          // latency mode: swapped-kernel convs with few tiles split their K range (ConvParams::ksplit)
    void* kpart = nullptr;        // split-K partial tiles + tickets (own allocation: sized after the launches are planned)
    void* arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<Op> ops;
    std::vector<ConvOp> convs;
    std::vector<GnApplyParams> gns;
    std::vector<AttnParams> attns;
    FirstConvParams first;
    // dynamic inputs
    float* mask[8] = {nullptr};   // per level [Be][W_l]
    int* x_row = nullptr;
    int* mu_row = nullptr;
    float* spk_rows = nullptr;    // [Be][S]
    float* t_rows = nullptr;      // [Be]
    float* u = nullptr;           // [Be][dim+S]
    float* E = nullptr;           // [Be][J]
    float* Spart = nullptr;       // [Be][J] speaker part of the embedding Linears (sampler path)
    long long* stats = nullptr;   // [slots][Be][groups][2] fixed point
    size_t stats_bytes = 0;
    __half* final_raw = nullptr;
    long long* final_stats = nullptr;
    float* xt = nullptr;          // [Be][P0] sampler state (only first B rows used)
    // captured sampler step (one estimator evaluation + fused final update) per (CFG branch count, denorm) -- see
    // reverse_diffusion; tied to this plan's buffers, destroyed with it
    std::map<int, cudaGraphExec_t> step_graphs;
};

}  // namespace usb

using namespace usb;

struct usb_handle {
    usb_config cfg;
    int num_sms = 148;
    int L = 0;
    int C[8] = {0};
    int hidden = 128, heads = 4;
    bool finalized = false;
    std::map<std::string, HostParam> host;
    std::vector<void*> dev_allocs;
    // device weights
    float *first_w3 = nullptr, *first_b3 = nullptr, *first_w1 = nullptr, *first_b1 = nullptr;
    std::vector<ResnetW> resnets;   // order: downs.k.0, downs.k.1 ..., mid_block1, mid_block2, ups.k.0, ups.k.1 ...
    std::vector<AttnW> attns;       // order: downs.k.2 ..., mid_attn, ups.k.2 ...
    std::vector<ConvW> down_convs, up_convs;
    ConvW final_block;
    GnW final_gn;
    float *final_w = nullptr, *final_b = nullptr;
    float *text_uncon = nullptr, *spk_uncon_normed = nullptr;
    float *freqs = nullptr, *mlp_w0 = nullptr, *mlp_b0 = nullptr, *mlp_w2 = nullptr, *mlp_b2 = nullptr;
    float *wcat = nullptr, *bcat = nullptr;
    int J = 0;
    Plan plan;
    float* tpart_buf = nullptr;   // sampler: [n_steps] t, [n_steps][dim+S] u, [n_steps][J] time part
    size_t tpart_cap = 0;
    float* stage_buf = nullptr;   // usb_reverse_diffusion_host: device copies of the host inputs / output
    size_t stage_cap = 0;
    PackBatch* pack = nullptr;    // fine-tune step: recorder of the per-conv data-gradient weight packs (train.h)
    float* t_kpart = nullptr;     // fine-tune step: split-K workspace of usb_t_conv (2 items per SM) + tickets behind it
    int* t_ktick = nullptr;
    float* loss_buf = nullptr;    // loss_t: xt, z*mask, cond*mask, score ([B][n_feats][T] each) + 512 doubles of partials
    size_t loss_cap = 0;
    long long launches = 0;
    // graph-replayed sampler (small workloads are launch-bound: ~120 short kernels per step): staging copies of the
    // caller's cond / noise / out at fixed addresses, the per-step scalar table and the device step counter
    int graph_mode = -1;                 // -1 auto (rows x frames <= kGraphAutoRowsFrames), 0 off, 1 on
    int splitk_mode = -1;                // same encoding: split-K of the few-tile convs (plan property)
    cudaStream_t cap_stream = nullptr;
    float* g_stage = nullptr;            // [B*P] cond, [B*P] out, [n*B*P] noise
    size_t g_cap = 0;
    float* step_tab = nullptr;           // [n][4]
    int step_tab_cap = 0;
    int* step_ctr = nullptr;
    long long graph_steps = 0;           // sampler steps executed as graph replays (usb_graph_steps)
    long long graph_launches_per_step = 0;
    float graph_a0 = 0.f, graph_a1 = 0.f;
    unsigned long long* sat = nullptr;   // device counter of fp16 saturation events (usb_saturation_count)
    float* mel_range = nullptr;          // device [2][n_feats] (mel_min, mel_max) when output de-normalisation is on
    bool denorm = false;
    // optional per-kernel-class timing (bench.py roofline): events around every launch of a profiled call
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Rec { int cls; size_t e0, e1; double work; };
    std::vector<Rec> recs;
    double prof_ms[4] = {0, 0, 0, 0};     // class 0 conv (tensor), 1 gn_apply, 2 attention, 3 other
    double prof_work[4] = {0, 0, 0, 0};   // conv: FLOPs; gn/attn/other: algorithmic bytes
    long long prof_launches[4] = {0, 0, 0, 0};
};

namespace usb {

int handle_device(const usb_handle* h) { return h->cfg.device; }

template <typename T>
static int upload(usb_handle* h, const T* host, size_t count, T** dev) {
    void* p = nullptr;
    USB_CUDA(cudaMalloc(&p, count * sizeof(T)));
    h->dev_allocs.push_back(p);
    USB_CUDA(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<T*>(p);
    return 0;
}

static int get_param(usb_handle* h, const std::string& key, const HostParam** out, size_t expect_count) {
    auto it = h->host.find(key);
    if (it == h->host.end()) return fail("missing parameter: " + key);
    if (expect_count && it->second.data.size() != expect_count)
        return fail("parameter " + key + " has " + std::to_string(it->second.data.size()) + " elements, expected " +
                    std::to_string(expect_count));
    *out = &it->second;
    return 0;
}

static int upload_param(usb_handle* h, const std::string& key, size_t expect_count, float** dev) {
    const HostParam* p;
    USB_TRY(get_param(h, key, &p, expect_count));
    return upload(h, p->data.data(), p->data.size(), dev);
}

static int load_conv(usb_handle* h, const std::string& prefix, int kind, int Cout, int Cin, bool has_bias, ConvW& w) {
    const HostParam* p;
    const int taps = kind == KT4 ? 16 : taps_of(kind);
    USB_TRY(get_param(h, prefix + ".weight", &p, (size_t)Cout * Cin * taps));
    std::vector<__half> packed;
    pack_conv_host(kind, p->data.data(), Cout, Cin, packed);
    USB_TRY(upload(h, packed.data(), packed.size(), &w.w));
    if (has_bias) USB_TRY(upload_param(h, prefix + ".bias", Cout, &w.bias));
    w.Cout = Cout;
    w.Cin = Cin;
    w.kind = kind;
    w.Z = kind == KT4 ? 4 : 1;
    w.K = taps_of(kind) * Cin;
    return 0;
}

static int load_gn(usb_handle* h, const std::string& prefix, int C, GnW& g) {
    USB_TRY(upload_param(h, prefix + ".weight", C, &g.gamma));
    USB_TRY(upload_param(h, prefix + ".bias", C, &g.beta));
    return 0;
}

static int load_resnet(usb_handle* h, const std::string& prefix, int Cin, int Cout, bool first, ResnetW& r) {
    r.prefix = prefix;
    r.Cin = Cin;
    r.Cout = Cout;
    r.has_res = Cin != Cout;
    if (!first) {
        USB_TRY(load_conv(h, prefix + ".block1.block.0", K3S1, Cout, Cin, true, r.c1));
        if (r.has_res) USB_TRY(load_conv(h, prefix + ".res_conv", K1, Cout, Cin, true, r.res));
    }
    USB_TRY(load_conv(h, prefix + ".block2.block.0", K3S1, Cout, Cout, true, r.c2));
    USB_TRY(load_gn(h, prefix + ".block1.block.1", Cout, r.g1));
    USB_TRY(load_gn(h, prefix + ".block2.block.1", Cout, r.g2));
    return 0;
}

static int load_attn(usb_handle* h, const std::string& prefix, int C, AttnW& a) {
    a.prefix = prefix;
    a.C = C;
    static const bool no_fuse = getenv("USB_NO_FUSED_Q") != nullptr;
    a.fused_q = C <= 256 && !no_fuse;
    if (a.fused_q) {
        const HostParam* p;
        const int hid = h->hidden;
        USB_TRY(get_param(h, prefix + ".fn.fn.to_qkv.weight", &p, (size_t)3 * hid * C));
        std::vector<__half> packed;
        pack_conv_host(K1, p->data.data() + (size_t)hid * C, 2 * hid, C, packed);   // k and v rows
        USB_TRY(upload(h, packed.data(), packed.size(), &a.kv.w));
        a.kv.Cout = 2 * hid; a.kv.Cin = C; a.kv.kind = K1; a.kv.Z = 1; a.kv.K = C;
        USB_TRY(upload(h, p->data.data(), (size_t)hid * C, &a.wq));                  // q rows, fp32
        std::vector<float> zeros(C, 0.f);
        USB_TRY(upload(h, zeros.data(), zeros.size(), &a.bprime));
    } else {
        USB_TRY(load_conv(h, prefix + ".fn.fn.to_qkv", K1, 3 * h->hidden, C, false, a.qkv));
    }
    USB_TRY(upload_param(h, prefix + ".fn.fn.to_out.weight", (size_t)C * h->hidden, &a.wo));
    USB_TRY(upload_param(h, prefix + ".fn.fn.to_out.bias", C, &a.bo));
    USB_TRY(upload_param(h, prefix + ".fn.g", 1, &a.g));
    return 0;
}

static int finalize_params(usb_handle* h) {
    const usb_config& c = h->cfg;
    const int L = h->L, dim = c.dim, S = c.spk_emb_dim;
    const std::string e = "estimator.";
    // ---- first conv (2 -> C0) and its res_conv, fp32 tap-major
    {
        const int C0 = h->C[0];
        const HostParam *w3, *w1;
        USB_TRY(get_param(h, e + "downs.0.0.block1.block.0.weight", &w3, (size_t)C0 * 2 * 9));
        USB_TRY(get_param(h, e + "downs.0.0.res_conv.weight", &w1, (size_t)C0 * 2));
        std::vector<float> a((size_t)18 * C0), b((size_t)2 * C0);
        for (int co = 0; co < C0; ++co)
            for (int ci = 0; ci < 2; ++ci) {
                for (int t = 0; t < 9; ++t) a[((size_t)t * 2 + ci) * C0 + co] = w3->data[((size_t)co * 2 + ci) * 9 + t];
                b[(size_t)ci * C0 + co] = w1->data[(size_t)co * 2 + ci];
            }
        USB_TRY(upload(h, a.data(), a.size(), &h->first_w3));
        USB_TRY(upload(h, b.data(), b.size(), &h->first_w1));
        USB_TRY(upload_param(h, e + "downs.0.0.block1.block.0.bias", C0, &h->first_b3));
        USB_TRY(upload_param(h, e + "downs.0.0.res_conv.bias", C0, &h->first_b1));
    }
    // ---- resnets / attention / resampling convs, in forward order
    h->resnets.clear();
    h->attns.clear();
    auto add_resnet = [&](const std::string& pre, int Cin, int Cout, bool first) -> int {
        h->resnets.emplace_back();
        return load_resnet(h, pre, Cin, Cout, first, h->resnets.back());
    };
    auto add_attn = [&](const std::string& pre, int C) -> int {
        h->attns.emplace_back();
        return load_attn(h, pre, C, h->attns.back());
    };
    for (int k = 0; k < L; ++k) {
        const int Cin = k == 0 ? 2 : h->C[k - 1], Cout = h->C[k];
        const std::string pre = e + "downs." + std::to_string(k);
        USB_TRY(add_resnet(pre + ".0", Cin, Cout, k == 0));
        USB_TRY(add_resnet(pre + ".1", Cout, Cout, false));
        USB_TRY(add_attn(pre + ".2", Cout));
        if (k < L - 1) {
            h->down_convs.emplace_back();
            USB_TRY(load_conv(h, pre + ".3.conv", K3S2, Cout, Cout, true, h->down_convs.back()));
        }
    }
    const int mid = h->C[L - 1];
    USB_TRY(add_resnet(e + "mid_block1", mid, mid, false));
    USB_TRY(add_attn(e + "mid_attn", mid));
    USB_TRY(add_resnet(e + "mid_block2", mid, mid, false));
    for (int k = 0; k < L - 1; ++k) {
        const int j = L - 1 - k;  // level of this up stage
        const int Cj = h->C[j], Cn = h->C[j - 1];
        const std::string pre = e + "ups." + std::to_string(k);
        USB_TRY(add_resnet(pre + ".0", 2 * Cj, Cn, false));
        USB_TRY(add_resnet(pre + ".1", Cn, Cn, false));
        USB_TRY(add_attn(pre + ".2", Cn));
        h->up_convs.emplace_back();
        USB_TRY(load_conv(h, pre + ".3.conv", KT4, Cn, Cn, true, h->up_convs.back()));
    }
    USB_TRY(load_conv(h, e + "final_block.block.0", K3S1, dim, dim, true, h->final_block));
    USB_TRY(load_gn(h, e + "final_block.block.1", dim, h->final_gn));
    USB_TRY(upload_param(h, e + "final_conv.weight", dim, &h->final_w));
    USB_TRY(upload_param(h, e + "final_conv.bias", 1, &h->final_b));
    // ---- embeddings
    USB_TRY(upload_param(h, "__posemb_freqs", dim / 2, &h->freqs));
    USB_TRY(upload_param(h, e + "mlp.0.weight", (size_t)4 * dim * dim, &h->mlp_w0));
    USB_TRY(upload_param(h, e + "mlp.0.bias", 4 * dim, &h->mlp_b0));
    USB_TRY(upload_param(h, e + "mlp.2.weight", (size_t)4 * dim * dim, &h->mlp_w2));
    USB_TRY(upload_param(h, e + "mlp.2.bias", dim, &h->mlp_b2));
    {
        int J = 0;
        for (auto& r : h->resnets) {
            r.emb_off = J;
            J += r.Cout;
        }
        h->J = J;
        const int K = dim + S;
        std::vector<float> wc((size_t)J * K), bc(J);
        for (auto& r : h->resnets) {
            const HostParam *w, *b;
            USB_TRY(get_param(h, r.prefix + ".mlp.1.weight", &w, (size_t)r.Cout * K));
            USB_TRY(get_param(h, r.prefix + ".mlp.1.bias", &b, r.Cout));
            memcpy(&wc[(size_t)r.emb_off * K], w->data.data(), w->data.size() * sizeof(float));
            memcpy(&bc[r.emb_off], b->data.data(), b->data.size() * sizeof(float));
        }
        USB_TRY(upload(h, wc.data(), wc.size(), &h->wcat));
        USB_TRY(upload(h, bc.data(), bc.size(), &h->bcat));
    }
    // ---- CFG unconditionals: text_uncon (1, n_feats, 1); spk_uncon / ||spk_uncon|| (unitspeech.py:355,358)
    USB_TRY(upload_param(h, "text_uncon", c.n_feats, &h->text_uncon));
    {
        const HostParam* su;
        USB_TRY(get_param(h, "spk_uncon", &su, S));
        // torch .norm(): fp32 sqrt of the fp32 sum of squares
        float ssq = 0.f;
        for (float v : su->data) ssq += v * v;
        const float nrm = std::sqrt(ssq);
        std::vector<float> sn(S);
        for (int i = 0; i < S; ++i) sn[i] = su->data[i] / nrm;
        USB_TRY(upload(h, sn.data(), sn.size(), &h->spk_uncon_normed));
    }
    h->finalized = true;
    h->host.clear();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// plan: buffers + launch parameter blocks for one (Be, T)
// ---------------------------------------------------------------------------------------------------------------
struct Bump {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off += (bytes + 1023) & ~size_t(1023);
        return o;
    }
};

static void drop_step_graphs(Plan& pl) {
    for (auto& kv : pl.step_graphs) cudaGraphExecDestroy(kv.second);
    pl.step_graphs.clear();
}

static void free_plan(Plan& pl) {
    drop_step_graphs(pl);
    if (pl.kpart) cudaFree(pl.kpart);
    if (pl.arena) cudaFree(pl.arena);
    pl = Plan();
}

static int build_plan(usb_handle* h, int Be, int T) {
    Plan& pl = h->plan;
    const bool splitk = h->splitk_mode == 1 || (h->splitk_mode < 0 && (long long)Be * T <= kSplitKAutoRowsFrames);
    if (pl.Be == Be && pl.T == T && pl.arena && pl.splitk == splitk) return 0;
    cudaDeviceSynchronize();
    free_plan(pl);
    pl.splitk = splitk;
    const usb_config& c = h->cfg;
    const int L = h->L, S = c.spk_emb_dim, dim = c.dim, G = c.groups, hid = h->hidden;
    if (T <= 0 || T % (1 << (L - 1))) return fail("T must be a positive multiple of 2^(len(dim_mults)-1)");
    if (Be <= 0) return fail("empty batch");
    int H[8], W[8];
    size_t PC[8];
    for (int l = 0; l < L; ++l) {
        H[l] = c.n_feats >> l;
        W[l] = T >> l;
        PC[l] = (size_t)Be * H[l] * W[l];
    }
    const int n_slots = 2 * (int)h->resnets.size() + 1;
    // ---- carve the arena (two passes: size, then pointers)
    size_t o_raw[8], o_h1[8], o_r[8], o_y0[8], o_y1[8], o_qkv[8], o_skip[8], o_xin[8], o_upx[8], o_weff[8], o_mask[8];
    Bump b;
    size_t max_part = 0;
    for (int l = 0; l < L; ++l) {
        const size_t act = PC[l] * h->C[l] * sizeof(__half);
        o_raw[l] = b.take(act); o_h1[l] = b.take(act); o_r[l] = b.take(act);
        o_y0[l] = b.take(act); o_y1[l] = b.take(act); o_skip[l] = b.take(act);
        o_qkv[l] = b.take(PC[l] * 3 * hid * sizeof(__half));
        o_xin[l] = l > 0 ? b.take(PC[l] * h->C[l - 1] * sizeof(__half)) : 0;
        o_upx[l] = l < L - 1 ? b.take(PC[l] * h->C[l] * sizeof(__half)) : 0;
        o_weff[l] = b.take((size_t)Be * h->C[l] * (h->C[l] <= 256 ? 256 : hid) * sizeof(__half));
        o_mask[l] = b.take((size_t)Be * W[l] * sizeof(float));
        const size_t part = attn_scratch_bytes(Be, h->heads, H[l] * W[l], attn_chunk_for(H[l] * W[l]));
        if (part > max_part) max_part = part;
    }
    const size_t o_part = b.take(max_part);
    const size_t o_xrow = b.take(Be * sizeof(int)), o_murow = b.take(Be * sizeof(int));
    const size_t o_spk = b.take((size_t)Be * S * sizeof(float)), o_t = b.take(Be * sizeof(float));
    const size_t o_u = b.take((size_t)Be * (dim + S) * sizeof(float)), o_E = b.take((size_t)Be * h->J * sizeof(float));
    const size_t o_Sp = b.take((size_t)Be * h->J * sizeof(float));
    pl.stats_bytes = (size_t)n_slots * Be * G * 2 * sizeof(long long);
    const size_t o_stats = b.take(pl.stats_bytes);
    const size_t o_xt = b.take((size_t)Be * c.n_feats * T * sizeof(float));
    pl.arena_bytes = b.off;
    {
        cudaError_t e = cudaMalloc(&pl.arena, pl.arena_bytes);
        if (e != cudaSuccess) {
            pl.arena = nullptr;
            return fail("workspace allocation of " + std::to_string(pl.arena_bytes >> 20) + " MiB failed: " +
                        cudaGetErrorString(e));
        }
    }
    char* A = static_cast<char*>(pl.arena);
    auto HP = [&](size_t o) { return reinterpret_cast<__half*>(A + o); };
    for (int l = 0; l < L; ++l) pl.mask[l] = reinterpret_cast<float*>(A + o_mask[l]);
    pl.x_row = reinterpret_cast<int*>(A + o_xrow);
    pl.mu_row = reinterpret_cast<int*>(A + o_murow);
    pl.spk_rows = reinterpret_cast<float*>(A + o_spk);
    pl.t_rows = reinterpret_cast<float*>(A + o_t);
    pl.u = reinterpret_cast<float*>(A + o_u);
    pl.E = reinterpret_cast<float*>(A + o_E);
    pl.Spart = reinterpret_cast<float*>(A + o_Sp);
    pl.stats = reinterpret_cast<long long*>(A + o_stats);
    pl.xt = reinterpret_cast<float*>(A + o_xt);
    float* part = reinterpret_cast<float*>(A + o_part);
    pl.Be = Be;
    pl.T = T;

    int slot = 0;
    auto next_stats = [&]() { return pl.stats + (size_t)(slot++) * Be * G * 2; };
    auto push_conv = [&](int kind, const __half* in0, int C0tot, int C0, const __half* in1, int C1tot, int C1, int l,
                         const ConvW* w, const __half* wptr, int wZ, int bmode, int Cout, const ConvEpilogue& ep,
                         __half* out) -> int {
        pl.convs.emplace_back();
        ConvEpilogue eps = ep;
        eps.sat = h->sat;
        eps.splitk = splitk; eps.num_sms = h->num_sms;
        USB_TRY(build_conv(pl.convs.back(), kind, in0, C0tot, C0, in1, C1tot, C1, Be, H[l], W[l], w ? w->w : wptr,
                           w ? w->Z : wZ, bmode, Cout, eps, out));
        pl.ops.push_back({Op::CONV, (int)pl.convs.size() - 1});
        return 0;
    };
    auto push_gn = [&](const __half* raw, const long long* stats, const GnW& g, const float* addvec, const __half* res,
                       int l, int Cc, __half* out) {
        GnApplyParams p;
        p.raw = raw; p.stats = stats; p.gamma = g.gamma; p.beta = g.beta; p.addvec = addvec; p.addvec_stride = h->J;
        p.res = res; p.mask = pl.mask[l]; p.out = out; p.N = Be; p.P = H[l] * W[l]; p.W = W[l]; p.C = Cc; p.groups = G;
        p.eps = 1e-5f;
        p.dbg = 0;
        p.sat = h->sat;
        pl.gns.push_back(p);
        pl.ops.push_back({Op::GN, (int)pl.gns.size() - 1});
    };
    // ResnetBlock (unitspeech.py:70-75); in1 = second K source of the skip concat (:192)
    auto push_resnet = [&](const ResnetW& r, int l, const __half* in0, int C0, const __half* in1, int C1,
                           __half* out) -> int {
        __half *raw = HP(o_raw[l]), *h1 = HP(o_h1[l]), *rb = HP(o_r[l]);
        ConvEpilogue ep;
        ep.groups = G;
        long long* s1 = next_stats();
        ep.bias = r.c1.bias; ep.stats = s1;
        USB_TRY(push_conv(K3S1, in0, C0, C0, in1, C1, C1, l, &r.c1, nullptr, 0, 0, r.Cout, ep, raw));
        push_gn(raw, s1, r.g1, pl.E + r.emb_off, nullptr, l, r.Cout, h1);
        long long* s2 = next_stats();
        ep.bias = r.c2.bias; ep.stats = s2;
        USB_TRY(push_conv(K3S1, h1, r.Cout, r.Cout, nullptr, 0, 0, l, &r.c2, nullptr, 0, 0, r.Cout, ep, raw));
        const __half* resid = in0;
        if (r.has_res) {
            ConvEpilogue er;
            er.bias = r.res.bias;
            USB_TRY(push_conv(K1, in0, C0, C0, in1, C1, C1, l, &r.res, nullptr, 0, 0, r.Cout, er, rb));
            resid = rb;
        } else if (in1) {
            return fail("identity residual with a concatenated input");
        }
        push_gn(raw, s2, r.g2, nullptr, resid, l, r.Cout, out);
        return 0;
    };
    // Residual(Rezero(LinearAttention)) (unitspeech.py:36-43,78-106), output stored masked
    auto push_attn = [&](const AttnW& a, int l, const __half* x, __half* out) -> int {
        __half *qkv = HP(o_qkv[l]), *weff = HP(o_weff[l]);
        AttnParams ap;
        memset(&ap, 0, sizeof ap);
        ap.wo = a.wo; ap.part = part; ap.weff = weff; ap.N = Be; ap.P = H[l] * W[l]; ap.C = a.C;
        ap.heads = h->heads; ap.chunk = attn_chunk_for(H[l] * W[l]); ap.g = a.g; ap.bo = a.bo; ap.qkv = qkv;
        ConvEpilogue e1;
        if (a.fused_q) {
            // kv = W_kv x; context from (k, v); the block's output is one per-sample 1x1 conv on x (no q, no residual read)
            USB_TRY(push_conv(K1, x, a.C, a.C, nullptr, 0, 0, l, &a.kv, nullptr, 0, 0, 2 * hid, e1, qkv));
            ap.ld = 2 * hid; ap.koff = 0; ap.voff = hid; ap.wq = a.wq; ap.bprime = a.bprime;
            pl.attns.push_back(ap);
            pl.ops.push_back({Op::ATTN, (int)pl.attns.size() - 1});
            ConvEpilogue e2;
            e2.bias = a.bprime; e2.mask = pl.mask[l];
            USB_TRY(push_conv(K1, x, a.C, a.C, nullptr, 0, 0, l, nullptr, weff, Be, 2, a.C, e2, out));
            return 0;
        }
        USB_TRY(push_conv(K1, x, a.C, a.C, nullptr, 0, 0, l, &a.qkv, nullptr, 0, 0, 3 * hid, e1, qkv));
        ap.ld = 3 * hid; ap.koff = hid; ap.voff = 2 * hid;
        pl.attns.push_back(ap);
        pl.ops.push_back({Op::ATTN, (int)pl.attns.size() - 1});
        ConvEpilogue e2;
        e2.bias = a.bo; e2.res = x; e2.res_scale = a.g; e2.mask = pl.mask[l];
        USB_TRY(push_conv(K1, qkv, 3 * hid, hid, nullptr, 0, 0, l, nullptr, weff, Be, 2, a.C, e2, out));
        return 0;
    };

    size_t ri = 0, ai = 0;
    // ---- down path
    const __half* x = nullptr;
    for (int k = 0; k < L; ++k) {
        __half *y0 = HP(o_y0[k]), *y1 = HP(o_y1[k]), *skip = HP(o_skip[k]);
        const ResnetW& r0 = h->resnets[ri++];
        if (k == 0) {
            __half *raw = HP(o_raw[0]), *h1 = HP(o_h1[0]), *rb = HP(o_r[0]);
            long long* s1 = next_stats();
            FirstConvParams& f = pl.first;
            memset(&f, 0, sizeof f);
            f.x_row = pl.x_row; f.mu_row = pl.mu_row; f.mask = pl.mask[0];
            f.w3 = h->first_w3; f.b3 = h->first_b3; f.w1 = h->first_w1; f.b1 = h->first_b1;
            f.raw = raw; f.res = rb; f.stats = s1; f.N = Be; f.H = H[0]; f.W = W[0]; f.C = h->C[0]; f.groups = G;
            pl.ops.push_back({Op::FIRST, 0});
            push_gn(raw, s1, r0.g1, pl.E + r0.emb_off, nullptr, 0, r0.Cout, h1);
            long long* s2 = next_stats();
            ConvEpilogue ep;
            ep.groups = G; ep.bias = r0.c2.bias; ep.stats = s2;
            USB_TRY(push_conv(K3S1, h1, r0.Cout, r0.Cout, nullptr, 0, 0, 0, &r0.c2, nullptr, 0, 0, r0.Cout, ep, raw));
            push_gn(raw, s2, r0.g2, nullptr, rb, 0, r0.Cout, y0);
        } else {
            USB_TRY(push_resnet(r0, k, x, h->C[k - 1], nullptr, 0, y0));
        }
        USB_TRY(push_resnet(h->resnets[ri++], k, y0, h->C[k], nullptr, 0, y1));
        USB_TRY(push_attn(h->attns[ai++], k, y1, skip));
        if (k < L - 1) {
            ConvEpilogue ed;
            ed.bias = h->down_convs[k].bias; ed.mask = pl.mask[k + 1];
            __half* xin = HP(o_xin[k + 1]);
            USB_TRY(push_conv(K3S2, skip, h->C[k], h->C[k], nullptr, 0, 0, k, &h->down_convs[k], nullptr, 0, 0, h->C[k],
                              ed, xin));
            x = xin;
        }
    }
    // ---- middle
    const int D = L - 1;
    USB_TRY(push_resnet(h->resnets[ri++], D, HP(o_skip[D]), h->C[D], nullptr, 0, HP(o_y0[D])));
    USB_TRY(push_attn(h->attns[ai++], D, HP(o_y0[D]), HP(o_y1[D])));
    USB_TRY(push_resnet(h->resnets[ri++], D, HP(o_y1[D]), h->C[D], nullptr, 0, HP(o_y0[D])));
    // ---- up path
    const __half* cur = HP(o_y0[D]);
    for (int k = 0; k < L - 1; ++k) {
        const int j = D - k;
        const int Cj = h->C[j], Cn = h->C[j - 1];
        USB_TRY(push_resnet(h->resnets[ri++], j, cur, Cj, HP(o_skip[j]), Cj, HP(o_y1[j])));
        USB_TRY(push_resnet(h->resnets[ri++], j, HP(o_y1[j]), Cn, nullptr, 0, HP(o_y0[j])));
        USB_TRY(push_attn(h->attns[ai++], j, HP(o_y0[j]), HP(o_h1[j])));
        ConvEpilogue eu;
        eu.bias = h->up_convs[k].bias; eu.mask = pl.mask[j - 1];
        USB_TRY(push_conv(KT4, HP(o_h1[j]), Cn, Cn, nullptr, 0, 0, j, &h->up_convs[k], nullptr, 0, 1, Cn, eu,
                          HP(o_upx[j - 1])));
        cur = HP(o_upx[j - 1]);
    }
    // ---- final block conv (GroupNorm/Mish/1x1 are fused into the final kernel)
    {
        ConvEpilogue ef;
        ef.groups = G; ef.bias = h->final_block.bias;
        pl.final_stats = next_stats();
        ef.stats = pl.final_stats;
        pl.final_raw = HP(o_raw[0]);
        USB_TRY(push_conv(K3S1, cur, h->C[0], h->C[0], nullptr, 0, 0, 0, &h->final_block, nullptr, 0, 0, dim, ef,
                          pl.final_raw));
    }
    if (slot != n_slots) return fail("internal: stats slot count mismatch");
    // split-K workspace: sized for the launch with the most work items, shared by all (launches are stream-ordered)
    long long max_items = 0, max_tiles = 0;
    for (const ConvOp& co : pl.convs)
        if (co.p.ksplit > 1) {
            const long long tiles = (long long)co.p.phases * ((co.p.patches_per_phase + 1) / 2) * co.p.n_tiles_n;
            max_tiles = tiles > max_tiles ? tiles : max_tiles;
            max_items = tiles * co.p.ksplit > max_items ? tiles * co.p.ksplit : max_items;
        }
    if (max_items > 0) {
        const size_t part_bytes = (size_t)max_items * 2 * 128 * 128 * sizeof(float);
        cudaError_t e = cudaMalloc(&pl.kpart, part_bytes + (size_t)max_tiles * sizeof(int));
        if (e != cudaSuccess) {
            pl.kpart = nullptr;
            return fail("split-K workspace allocation of " + std::to_string(part_bytes >> 20) + " MiB failed: " + cudaGetErrorString(e));
        }
        int* ktick = reinterpret_cast<int*>(static_cast<char*>(pl.kpart) + part_bytes);
        USB_CUDA(cudaMemset(ktick, 0, (size_t)max_tiles * sizeof(int)));
        for (ConvOp& co : pl.convs)
            if (co.p.ksplit > 1) {
                co.p.kpart = static_cast<float*>(pl.kpart);
                co.p.ktick = ktick;
            }
        pl.arena_bytes += part_bytes;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// execution
// ---------------------------------------------------------------------------------------------------------------
// One launch builds everything a call needs per estimator row (no pageable-host uploads, hence no stream sync):
// the CFG row tables (unitspeech.py:301-317: row k*B+b reads x_t of utterance b and cond of utterance b, or text_uncon when
// k == k_tu), the speaker rows (spk_uncon/||spk_uncon|| when k == k_su, :358), and the frame masks of every level
// (mask[..., ::2] per down-sampling, :182).  spk_rows == null: speaker rows are the caller's (estimator entry).
struct PrepParams {
    const float* mask;      // [B][T]
    const float* spk;       // [B][S] or null
    const float* spk_uncon; // [S]
    int* x_row;
    int* mu_row;
    float* spk_rows;        // [Be][S] or null
    float* mask_l[8];       // per level [Be][T >> l]
    int B, nb, k_tu, k_su, T, S, L;
};
__global__ void __launch_bounds__(256) prep_rows_kernel(const PrepParams p) {
    const int Be = p.B * p.nb;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < Be) {
        const int k = static_cast<int>(i) / p.B, b = static_cast<int>(i) % p.B;
        p.x_row[i] = b;
        p.mu_row[i] = k == p.k_tu ? -1 : b;
    }
    if (p.spk_rows && i < static_cast<long long>(Be) * p.S) {
        const int r = static_cast<int>(i / p.S), c = static_cast<int>(i % p.S);
        const int k = r / p.B, b = r % p.B;
        p.spk_rows[i] = k == p.k_su ? __ldg(p.spk_uncon + c) : __ldg(p.spk + static_cast<long long>(b) * p.S + c);
    }
    if (i < static_cast<long long>(Be) * p.T) {
        const int r = static_cast<int>(i / p.T), w = static_cast<int>(i % p.T);
        const float m = __ldg(p.mask + static_cast<long long>(r % p.B) * p.T + w);
        for (int l = 0; l < p.L; ++l)
            if ((w & ((1 << l) - 1)) == 0) p.mask_l[l][static_cast<long long>(r) * (p.T >> l) + (w >> l)] = m;
    }
}
__global__ void mul_mask_kernel(const float* z, const float* mask, float* out, int B, int P, int W) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * P) return;
    const int b = (int)(i / P), w = (int)((i % P) % W);
    out[i] = z[i] * mask[(long long)b * W + w];
}

static size_t prof_event(usb_handle* h, cudaStream_t s) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev_pool.push_back(e);
    }
    cudaEventRecord(h->ev_pool[h->ev_used], s);
    return h->ev_used++;
}
struct ProfScope {
    usb_handle* h;
    cudaStream_t s;
    int cls;
    double work;
    size_t e0 = 0;
    ProfScope(usb_handle* h_, cudaStream_t s_, int cls_, double work_) : h(h_), s(s_), cls(cls_), work(work_) {
        if (h->profiling) e0 = prof_event(h, s);
    }
    ~ProfScope() {
        if (h->profiling) h->recs.push_back({cls, e0, prof_event(h, s), work});
    }
};
static double conv_flops(const ConvParams& p) {
    const double px = (double)p.phases * p.N * p.Hm * p.Wm;
    return 2.0 * px * p.Cout * (double)p.taps * (p.chunks0 + p.chunks1) * 64.0;
}
static int prof_collect(usb_handle* h, cudaStream_t s) {
    if (!h->profiling) return 0;
    USB_CUDA(cudaStreamSynchronize(s));
    for (const auto& r : h->recs) {
        float ms = 0.f;
        USB_CUDA(cudaEventElapsedTime(&ms, h->ev_pool[r.e0], h->ev_pool[r.e1]));
        h->prof_ms[r.cls] += ms;
        h->prof_work[r.cls] += r.work;
        h->prof_launches[r.cls]++;
    }
    h->recs.clear();
    h->ev_used = 0;
    return 0;
}

struct EstInputs {
    const float* x;           // [*][H][W]
    const float* cond;        // [*][H][W]
    const float* text_uncon;  // [H] or null
    const float* t_rows;      // [Be]
    const float* spk_rows;    // [Be][S]
};

// row tables, speaker rows and the masks of all levels for this call (one launch, see prep_rows_kernel)
static int prepare_rows(usb_handle* h, const float* mask, const float* spk, int B, int nb, int k_tu, int k_su,
                        bool fill_spk, cudaStream_t s) {
    Plan& pl = h->plan;
    PrepParams pp;
    memset(&pp, 0, sizeof pp);
    pp.mask = mask; pp.spk = spk; pp.spk_uncon = h->spk_uncon_normed; pp.x_row = pl.x_row; pp.mu_row = pl.mu_row;
    pp.spk_rows = fill_spk ? pl.spk_rows : nullptr;
    for (int l = 0; l < h->L; ++l) pp.mask_l[l] = pl.mask[l];
    pp.B = B; pp.nb = nb; pp.k_tu = k_tu; pp.k_su = k_su; pp.T = pl.T; pp.S = h->cfg.spk_emb_dim; pp.L = h->L;
    const long long Be = (long long)B * nb;
    const long long total = Be * (pl.T > pp.S ? pl.T : pp.S);
    prep_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pp);
    USB_LAUNCH(h, (int)cudaGetLastError());
    return 0;
}

// everything up to and including the final_block conv
static EmbedParams embed_params(usb_handle* h) {
    const usb_config& c = h->cfg;
    EmbedParams ep;
    memset(&ep, 0, sizeof ep);
    ep.freqs = h->freqs; ep.w0 = h->mlp_w0; ep.b0 = h->mlp_b0; ep.w2 = h->mlp_w2; ep.b2 = h->mlp_b2;
    ep.wcat = h->wcat; ep.bcat = h->bcat; ep.dim = c.dim; ep.S = c.spk_emb_dim; ep.J = h->J; ep.pe_scale = c.pe_scale;
    return ep;
}

static int run_estimator(usb_handle* h, const EstInputs& in, cudaStream_t s, const float* t_part = nullptr,
                         const int* step_ctr = nullptr) {
    Plan& pl = h->plan;
    const usb_config& c = h->cfg;
    if (t_part) {
        // sampler path: E = (time part of this step, computed once per call) + (speaker part, once per call);
        // with step_ctr, t_part is the base of the per-step table and the device counter selects the row
        // (the same launch zeroes the GroupNorm statistics slots: no memset node between the kernels of a step)
        ProfScope ps(h, s, 3, 0.0);
        USB_LAUNCH(h, launch_emb_combine(t_part, pl.Spart, pl.E, pl.Be, h->J, step_ctr,
                                         reinterpret_cast<unsigned long long*>(pl.stats), (long long)(pl.stats_bytes / 8), s));
    }
    EmbedParams ep;
    ep.t = in.t_rows; ep.spk = in.spk_rows; ep.freqs = h->freqs; ep.w0 = h->mlp_w0; ep.b0 = h->mlp_b0;
    ep.w2 = h->mlp_w2; ep.b2 = h->mlp_b2; ep.wcat = h->wcat; ep.bcat = h->bcat; ep.u = pl.u; ep.e = pl.E;
    ep.N = pl.Be; ep.dim = c.dim; ep.S = c.spk_emb_dim; ep.J = h->J; ep.pe_scale = c.pe_scale;
    if (!t_part) {
        ProfScope ps(h, s, 3, 0.0);
        USB_LAUNCH(h, launch_embed(ep, s));
        h->launches++;  // launch_embed issues two kernels
    }
    if (!t_part) USB_CUDA(cudaMemsetAsync(pl.stats, 0, pl.stats_bytes, s));
    for (const Op& op : pl.ops) {
        switch (op.kind) {
            case Op::FIRST: {
                FirstConvParams f = pl.first;
                f.x = in.x; f.cond = in.cond; f.text_uncon = in.text_uncon;
                // reads 2 fp32 planes, writes raw + res fp16
                ProfScope ps(h, s, 3, (double)f.N * f.H * f.W * (8.0 + 4.0 * f.C));
                USB_LAUNCH(h, launch_first_conv(f, s));
                break;
            }
            case Op::CONV: {
                const ConvOp& co = pl.convs[op.idx];
                ProfScope ps(h, s, 0, conv_flops(co.p));
                USB_LAUNCH(h, launch_conv_igemm(co.p, co.a0, co.a1, co.b, co.o, h->num_sms, s));
                break;
            }
            case Op::GN: {
                const GnApplyParams& g = pl.gns[op.idx];
                // algorithmic bytes: read raw, (read res), write out, fp16
                ProfScope ps(h, s, 1, (double)g.N * g.P * g.C * 2.0 * (g.res ? 3.0 : 2.0));
                USB_LAUNCH(h, launch_gn_apply(g, h->num_sms, s));
                break;
            }
            case Op::ATTN: {
                const AttnParams& a = pl.attns[op.idx];
                // algorithmic bytes: k and v read once (fp16, 2*hidden channels)
                ProfScope ps(h, s, 2, (double)a.N * a.P * 2.0 * a.heads * 32 * 2.0);
                USB_LAUNCH(h, launch_attn_context(a, s));
                h->launches += 2;  // three kernels
                break;
            }
        }
    }
    return 0;
}

static FinalParams final_params(usb_handle* h, int B, int nb) {
    Plan& pl = h->plan;
    const usb_config& c = h->cfg;
    FinalParams f;
    memset(&f, 0, sizeof f);
    f.raw = pl.final_raw; f.stats = pl.final_stats; f.gamma = h->final_gn.gamma; f.beta = h->final_gn.beta;
    f.wf = h->final_w; f.bf = h->final_b; f.mask = pl.mask[0]; f.B = B; f.nb = nb; f.P = c.n_feats * pl.T;
    f.W = pl.T; f.C = c.dim; f.groups = c.groups; f.eps = 1e-5f;
    return f;
}

static int check_ready(usb_handle* h) {
    if (!h) return fail("null handle");
    if (!h->finalized) return fail("parameters not finalized (call usb_finalize_params)");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    return 0;
}

static int estimator_forward(usb_handle* h, const float* x, const float* mu, const float* mask, const float* t,
                             const float* spk, float* out, int Be, int T, cudaStream_t s) {
    USB_TRY(check_ready(h));
    USB_TRY(build_plan(h, Be, T));
    USB_TRY(prepare_rows(h, mask, nullptr, Be, 1, -1, -1, false, s));
    EstInputs in{x, mu, nullptr, t, spk};
    USB_TRY(run_estimator(h, in, s));
    FinalParams f = final_params(h, Be, 1);
    f.score = out;
    USB_LAUNCH(h, launch_final(f, h->num_sms, s));
    return 0;
}

// UnitSpeech.forward_diffusion / loss_t (unitspeech/unitspeech.py:376-405), forward value only
static int loss_t(usb_handle* h, const float* x0, const float* cond, const float* mask, const float* t, const float* spk,
                  const float* z, float* loss_out, float* xt_out, int B, int T, cudaStream_t s) {
    USB_TRY(check_ready(h));
    const int F = h->cfg.n_feats;
    const size_t n = static_cast<size_t>(B) * F * T;
    const size_t need = 4 * n * sizeof(float) + 512 * sizeof(double);
    if (need > h->loss_cap) {
        USB_CUDA(cudaDeviceSynchronize());
        if (h->loss_buf) cudaFree(h->loss_buf);
        h->loss_buf = nullptr;
        h->loss_cap = 0;
        USB_CUDA(cudaMalloc(&h->loss_buf, need));
        h->loss_cap = need;
    }
    double* partial = reinterpret_cast<double*>(h->loss_buf);       // first: keeps the doubles 8-byte aligned
    float* xt = h->loss_buf + 1024;
    float *zm = xt + n, *mu = zm + n, *score = mu + n;
    USB_LAUNCH(h, launch_forward_diffusion(x0, z, cond, mask, t, h->cfg.beta_min, h->cfg.beta_max, xt, zm, mu, B, F, T, s));
    USB_TRY(estimator_forward(h, xt, mu, mask, t, spk, score, B, T, s));
    h->launches++;
    USB_LAUNCH(h, launch_diffusion_loss(score, zm, mask, t, h->cfg.beta_min, h->cfg.beta_max, partial, loss_out, B, F, T, s));
    if (xt_out) USB_CUDA(cudaMemcpyAsync(xt_out, xt, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

static int reverse_diffusion(usb_handle* h, const float* z, const float* cond, const float* mask, const float* spk,
                             const float* noise, const float* coef, const float* t_steps, int n_steps, float tg,
                             float sg, float* out, float* trace, int B, int T, cudaStream_t s) {
    USB_TRY(check_ready(h));
    if (n_steps < 2) return fail("n_timesteps must be >= 2 (the reference crashes for 1)");
    if (B <= 0) return fail("empty batch");
    const usb_config& c = h->cfg;
    const bool use_t = tg > 0.f, use_s = sg > 0.f;
    const int nb = 1 + (use_t ? 1 : 0) + (use_s ? 1 : 0);
    const int Be = nb * B;
    USB_TRY(build_plan(h, Be, T));
    Plan& pl = h->plan;
    const int S = c.spk_emb_dim, P = c.n_feats * T;
    // branch order of classifier_free_guidance (unitspeech.py:301-317): [text-uncond] [spk-uncond] full
    int kb = 0, k_tu = -1, k_su = -1;
    if (use_t) k_tu = kb++;
    if (use_s) k_su = kb++;
    USB_TRY(prepare_rows(h, mask, spk, B, nb, k_tu, k_su, true, s));
    {
        const long long tot = (long long)B * P;
        mul_mask_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(z, mask, pl.xt, B, P, T);
        USB_LAUNCH(h, (int)cudaGetLastError());
    }
    // The embedding Linears are linear in [Mish(time_mlp(t)), Mish(spk)]: the time part depends only on the step and
    // the speaker part only on the row, so both are computed once per call (unitspeech.py:165-168,61,72).
    const int Kemb = c.dim + S;
    {
        const size_t need = (size_t)n_steps * (1 + Kemb + h->J);
        if (need > h->tpart_cap) {
            USB_CUDA(cudaStreamSynchronize(s));
            drop_step_graphs(pl);      // the captured step reads the per-step embedding table at its old address
            if (h->tpart_buf) cudaFree(h->tpart_buf);
            h->tpart_buf = nullptr; h->tpart_cap = 0;
            USB_CUDA(cudaMalloc(&h->tpart_buf, need * sizeof(float)));
            h->tpart_cap = need;
        }
        float* d_t = h->tpart_buf;
        float* d_ut = d_t + n_steps;
        float* d_tp = d_ut + (size_t)n_steps * Kemb;
        USB_CUDA(cudaMemcpyAsync(d_t, t_steps, n_steps * sizeof(float), cudaMemcpyHostToDevice, s));
        EmbedParams et = embed_params(h);
        et.t = d_t; et.spk = nullptr; et.u = d_ut; et.e = d_tp; et.N = n_steps;             // time part (+ bias)
        USB_LAUNCH(h, launch_embed(et, s));
        EmbedParams es = embed_params(h);
        es.t = nullptr; es.spk = pl.spk_rows; es.u = pl.u; es.e = pl.Spart; es.N = Be; es.bcat = nullptr;   // speaker part
        USB_LAUNCH(h, launch_embed(es, s));
        h->launches += 2;
    }
    const float* d_tpart = h->tpart_buf + n_steps + (size_t)n_steps * Kemb;
    const float a0 = nb == 3 ? tg : (use_t ? tg : sg);
    auto sampler_final = [&](int i, const float* noise_i, cudaStream_t st) -> int {
        FinalParams f = final_params(h, B, nb);
        f.a0 = a0;
        f.a1 = sg;
        f.xt = pl.xt;
        f.noise = noise_i;
        f.c_x = coef[i * 3 + 0]; f.c_s = coef[i * 3 + 1]; f.sigma = coef[i * 3 + 2];
        if (i == n_steps - 1) {
            // the reference returns xt * mask (:373; xt is already masked by the update): the last step writes the caller's
            // buffer directly, de-normalised to log-mel when usb_set_output_denorm is active (inference.py:140)
            f.out = out;
            if (h->denorm) { f.mel_min = h->mel_range; f.mel_max = h->mel_range + c.n_feats; }
        }
        // reads the final_block conv output of every CFG branch (fp16) + x_t and noise, writes x_t
        ProfScope ps(h, st, 3, (double)Be * P * c.dim * 2.0 + (double)B * P * 12.0);
        USB_LAUNCH(h, launch_final(f, h->num_sms, st));
        return 0;
    };
    // ---- small workloads: replay ONE captured step (device step counter + per-step scalar table) instead of issuing
    // ~120 launches per step from the host.  Same kernels, same arguments: results are bit-identical to the eager loop.
    const bool want_graph = !h->profiling && !trace && n_steps > 2 &&
                            (h->graph_mode == 1 || (h->graph_mode < 0 && (long long)Be * T <= kGraphAutoRowsFrames));
    int first_graph_step = n_steps;
    if (want_graph) {
        const size_t BP = (size_t)B * P;
        const size_t need = BP * (2 + (noise ? (size_t)n_steps : 0));
        if (need > h->g_cap || n_steps > h->step_tab_cap) {
            USB_CUDA(cudaStreamSynchronize(s));
            drop_step_graphs(pl);
            if (need > h->g_cap) {
                if (h->g_stage) cudaFree(h->g_stage);
                h->g_stage = nullptr; h->g_cap = 0;
                USB_CUDA(cudaMalloc(&h->g_stage, need * sizeof(float)));
                h->g_cap = need;
            }
            if (n_steps > h->step_tab_cap) {
                if (h->step_tab) cudaFree(h->step_tab);
                h->step_tab = nullptr; h->step_tab_cap = 0;
                USB_CUDA(cudaMalloc(&h->step_tab, (size_t)n_steps * 4 * sizeof(float)));
                h->step_tab_cap = n_steps;
            }
        }
        if (!h->step_ctr) {
            USB_CUDA(cudaMalloc(&h->step_ctr, sizeof(int)));
            USB_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        }
        first_graph_step = 1;      // step 0 always runs eagerly: it also performs every lazy one-time kernel-attribute set-up
    }
    // USB_PDL=1: latency mode also launches the step's kernels as programmatic dependents (pdl.h): the next kernel's blocks
    // are scheduled and run their prologue while the current one finishes.  Measured on B200 inside the replayed graph it
    // gains nothing (105.1 vs 103.3 ms per 1 x 256 pass): graph replay has already removed the launch gaps and every kernel
    // needs the predecessor's output from its first load.  Off by default, kept for A/B runs.
    static const bool use_pdl = getenv("USB_PDL") != nullptr;
    PdlScope pdl_scope(want_graph && use_pdl);
    EstInputs in{pl.xt, cond, h->text_uncon, pl.t_rows, pl.spk_rows};
    for (int i = 0; i < n_steps && i < first_graph_step; ++i) {
        USB_TRY(run_estimator(h, in, s, d_tpart + (size_t)i * h->J));
        USB_TRY(sampler_final(i, noise ? noise + (size_t)i * B * P : nullptr, s));
        if (trace)
            USB_CUDA(cudaMemcpyAsync(trace + (size_t)i * B * P, pl.xt, (size_t)B * P * sizeof(float),
                                     cudaMemcpyDeviceToDevice, s));
    }
    if (want_graph) {
        const size_t BP = (size_t)B * P;
        float *g_cond = h->g_stage, *g_out = g_cond + BP, *g_noise = g_out + BP;
        std::vector<float> tab((size_t)n_steps * 4);
        for (int i = 0; i < n_steps; ++i) {
            tab[4 * i] = coef[i * 3]; tab[4 * i + 1] = coef[i * 3 + 1]; tab[4 * i + 2] = coef[i * 3 + 2];
            tab[4 * i + 3] = i == n_steps - 1 ? 1.f : 0.f;
        }
        // (pageable source: the runtime stages the bytes before the call returns, so `tab` may go out of scope)
        USB_CUDA(cudaMemcpyAsync(h->step_tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        USB_CUDA(cudaMemcpyAsync(g_cond, cond, BP * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (noise)
            USB_CUDA(cudaMemcpyAsync(g_noise, noise, (size_t)n_steps * BP * sizeof(float), cudaMemcpyDeviceToDevice, s));
        USB_CUDA(cudaMemsetAsync(h->step_ctr, 0, sizeof(int), s));      // step 0 is done; the graph's first node makes it 1
        // key: everything baked into the captured arguments besides the plan's own buffers
        const int key = nb | (h->denorm ? 8 : 0) | (noise ? 16 : 0) | (use_t ? 32 : 0);
        auto it = pl.step_graphs.find(key);
        if (it != pl.step_graphs.end() && (h->graph_a0 != a0 || h->graph_a1 != sg)) {   // guidance scales are kernel arguments
            USB_CUDA(cudaStreamSynchronize(s));     // replays of the old graph may still be in flight on this stream
            drop_step_graphs(pl);
            it = pl.step_graphs.end();
        }
        if (it == pl.step_graphs.end()) {
            cudaStream_t cs = h->cap_stream;
            const long long launches_before = h->launches;
            USB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
            int rc = 0;
            {
                int e = launch_step_advance(h->step_ctr, cs);
                h->launches++;
                if (e != 0) rc = fail("step_advance launch failed");
            }
            EstInputs gin{pl.xt, g_cond, h->text_uncon, pl.t_rows, pl.spk_rows};
            if (!rc) rc = run_estimator(h, gin, cs, d_tpart, h->step_ctr);
            if (!rc) {
                FinalParams f = final_params(h, B, nb);
                f.a0 = a0; f.a1 = sg; f.xt = pl.xt;
                f.noise = noise ? g_noise : nullptr;
                f.noise_step_stride = (long long)BP;
                f.step_ctr = h->step_ctr; f.step_tab = h->step_tab;
                f.out = g_out;
                if (h->denorm) { f.mel_min = h->mel_range; f.mel_max = h->mel_range + c.n_feats; }
                int e = launch_final(f, h->num_sms, cs);
                h->launches++;
                if (e != 0) rc = fail("final launch failed during capture");
            }
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamEndCapture(cs, &graph);
            h->graph_launches_per_step = h->launches - launches_before;
            h->launches = launches_before;
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (ce != cudaSuccess) return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
            pl.step_graphs[key] = exec;
            h->graph_a0 = a0; h->graph_a1 = sg;
            it = pl.step_graphs.find(key);
        }
        for (int i = first_graph_step; i < n_steps; ++i) USB_CUDA(cudaGraphLaunch(it->second, s));
        h->launches += h->graph_launches_per_step * (n_steps - first_graph_step);
        h->graph_steps += n_steps - first_graph_step;
        USB_CUDA(cudaMemcpyAsync(out, g_out, BP * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    USB_TRY(prof_collect(h, s));
    return 0;
}

}  // namespace usb

// ===================================================================================================================
// C ABI
// ===================================================================================================================
extern "C" {

const char* usb_last_error(void) { return g_err.c_str(); }
int usb_version(void) { return 1; }

int usb_create(const usb_config* cfg, usb_handle** out) {
    if (!cfg || !out) return fail("null argument");
    if (!(cfg->dim == 64 || cfg->dim == 128 || cfg->dim == 256)) return fail("dim must be 64, 128 or 256");
    if (cfg->n_mults < 2 || cfg->n_mults > 4) return fail("len(dim_mults) must be 2..4");
    if (cfg->groups != 8) return fail("groups must be 8");
    if (cfg->n_feats % (1 << (cfg->n_mults - 1))) return fail("n_feats must be divisible by 2^(len(dim_mults)-1)");
    if (cfg->spk_emb_dim <= 0) return fail("spk_emb_dim must be positive");
    int ndev = 0;
    USB_CUDA(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail("no such CUDA device");
    USB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    USB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(std::string("this library only runs on sm_100 (B200); found ") + prop.name);
    usb_handle* h = new usb_handle();
    h->cfg = *cfg;
    h->num_sms = prop.multiProcessorCount;
    h->L = cfg->n_mults;
    for (int i = 0; i < h->L; ++i) {
        h->C[i] = cfg->dim * cfg->dim_mults[i];
        if (h->C[i] > 2048) {
            delete h;
            return fail("channel count above 2048 is not supported");
        }
    }
    if (load_encode_fn()) {
        delete h;
        return 1;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess || cudaMemset(p, 0, sizeof(unsigned long long)) != cudaSuccess) {
        delete h;
        return fail("cudaMalloc of the saturation counter failed");
    }
    h->sat = static_cast<unsigned long long*>(p);
    h->dev_allocs.push_back(p);
    if (cudaMalloc(&p, 2 * sizeof(float) * cfg->n_feats) != cudaSuccess) {
        usb_destroy(h);
        return fail("cudaMalloc of the mel range failed");
    }
    h->mel_range = static_cast<float*>(p);
    h->dev_allocs.push_back(p);
    *out = h;
    return 0;
}

void usb_destroy(usb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    free_plan(h->plan);
    for (void* p : h->dev_allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->tpart_buf) cudaFree(h->tpart_buf);
    if (h->g_stage) cudaFree(h->g_stage);
    if (h->step_tab) cudaFree(h->step_tab);
    if (h->step_ctr) cudaFree(h->step_ctr);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->loss_buf) cudaFree(h->loss_buf);
    pack_batch_destroy(h->pack);
    if (h->stage_buf) cudaFree(h->stage_buf);
    delete h;
}

int usb_load_param(usb_handle* h, const char* key, const float* data, const int64_t* shape, int32_t ndim) {
    if (!h || !key || !data) return fail("null argument");
    if (h->finalized) return fail("parameters already finalized");
    HostParam p;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        p.shape.push_back(shape[i]);
        n *= (size_t)shape[i];
    }
    p.data.assign(data, data + n);
    h->host[key] = std::move(p);
    return 0;
}

int usb_finalize_params(usb_handle* h) {
    if (!h) return fail("null handle");
    if (h->finalized) return fail("parameters already finalized");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    return finalize_params(h);
}

int usb_estimator_forward(usb_handle* h, const float* x, const float* mu, const float* mask, const float* t,
                          const float* spk, float* out, int32_t Be, int32_t T, uint64_t stream) {
    return estimator_forward(h, x, mu, mask, t, spk, out, Be, T, reinterpret_cast<cudaStream_t>(stream));
}

int usb_forward_diffusion(usb_handle* h, const float* x0, const float* mask, const float* t, const float* z, float* xt_out,
                          float* zmask_out, int32_t B, int32_t T, uint64_t stream) {
    if (!h || !x0 || !mask || !t || !z || !xt_out) return fail("null argument");
    if (B < 1 || T < 1) return fail("B and T must be positive");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    USB_LAUNCH(h, launch_forward_diffusion(x0, z, nullptr, mask, t, h->cfg.beta_min, h->cfg.beta_max, xt_out, zmask_out,
                                           nullptr, B, h->cfg.n_feats, T, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}

int usb_loss_t(usb_handle* h, const float* x0, const float* cond, const float* mask, const float* t, const float* spk,
               const float* z, float* loss_out, float* xt_out, int32_t B, int32_t T, uint64_t stream) {
    if (!h || !x0 || !cond || !mask || !t || !spk || !z || !loss_out) return fail("null argument");
    return loss_t(h, x0, cond, mask, t, spk, z, loss_out, xt_out, B, T, reinterpret_cast<cudaStream_t>(stream));
}

int usb_reverse_diffusion(usb_handle* h, const float* z, const float* cond, const float* mask, const float* spk,
                          const float* noise, const float* coef_host, const float* t_steps_host, int32_t n_steps,
                          float text_scale, float spk_scale, float* out, float* trace, int32_t B, int32_t T,
                          uint64_t stream) {
    return reverse_diffusion(h, z, cond, mask, spk, noise, coef_host, t_steps_host, n_steps, text_scale, spk_scale, out,
                             trace, B, T, reinterpret_cast<cudaStream_t>(stream));
}

int usb_reverse_diffusion_host(usb_handle* h, const float* z, const float* cond, const float* mask, const float* spk,
                               const float* noise, const float* coef_host, const float* t_steps_host, int32_t n_steps,
                               float text_scale, float spk_scale, float* out, int32_t B, int32_t T, uint64_t stream) {
    USB_TRY(check_ready(h));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t P = (size_t)h->cfg.n_feats * T, S = h->cfg.spk_emb_dim;
    const size_t n_z = (size_t)B * P, n_mask = (size_t)B * T, n_spk = (size_t)B * S;
    const size_t n_noise = noise ? (size_t)n_steps * B * P : 0;
    // device staging for the host entry, kept on the handle and grown on demand (no cudaMalloc / cudaFree per call)
    const size_t need = (3 * n_z + n_mask + n_spk + n_noise) * sizeof(float);
    if (need > h->stage_cap) {
        USB_CUDA(cudaStreamSynchronize(s));
        if (h->stage_buf) cudaFree(h->stage_buf);
        h->stage_buf = nullptr;
        h->stage_cap = 0;
        USB_CUDA(cudaMalloc(&h->stage_buf, need));
        h->stage_cap = need;
    }
    float* d = h->stage_buf;
    float *dz = d, *dc = dz + n_z, *dout = dc + n_z, *dm = dout + n_z, *ds = dm + n_mask, *dn = ds + n_spk;
    int rc = 0;
    auto H2D = [&](float* dst, const float* src, size_t n) {
        return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyHostToDevice, s);
    };
    cudaError_t e = H2D(dz, z, n_z);
    if (e == cudaSuccess) e = H2D(dc, cond, n_z);
    if (e == cudaSuccess) e = H2D(dm, mask, n_mask);
    if (e == cudaSuccess) e = H2D(ds, spk, n_spk);
    if (e == cudaSuccess && noise) e = H2D(dn, noise, n_noise);
    if (e != cudaSuccess) rc = fail(std::string("host->device copy: ") + cudaGetErrorString(e));
    if (!rc)
        rc = reverse_diffusion(h, dz, dc, dm, ds, noise ? dn : nullptr, coef_host, t_steps_host, n_steps, text_scale,
                               spk_scale, dout, nullptr, B, T, s);
    if (!rc) {
        e = cudaMemcpyAsync(out, dout, n_z * sizeof(float), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(std::string("device->host copy: ") + cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(s);
    }
    return rc;
}

int usb_set_profiling(usb_handle* h, int32_t on) {
    if (!h) return fail("null handle");
    h->profiling = on != 0;
    for (int i = 0; i < 4; ++i) {
        h->prof_ms[i] = 0;
        h->prof_work[i] = 0;
        h->prof_launches[i] = 0;
    }
    return 0;
}

int usb_get_profile(usb_handle* h, double* ms4, double* work4, int64_t* launches4) {
    if (!h || !ms4 || !work4 || !launches4) return fail("null argument");
    for (int i = 0; i < 4; ++i) {
        ms4[i] = h->prof_ms[i];
        work4[i] = h->prof_work[i];
        launches4[i] = h->prof_launches[i];
    }
    return 0;
}

int usb_saturation_count(usb_handle* h, int64_t* count_out, int32_t reset) {
    if (!h || !count_out) return fail("null argument");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    unsigned long long v = 0;
    USB_CUDA(cudaMemcpy(&v, h->sat, sizeof v, cudaMemcpyDeviceToHost));   // synchronises with the work that reports into it
    if (reset) USB_CUDA(cudaMemset(h->sat, 0, sizeof v));
    *count_out = static_cast<int64_t>(v);
    return 0;
}

int usb_set_output_denorm(usb_handle* h, const float* mel_min_host, const float* mel_max_host) {
    if (!h) return fail("null handle");
    if ((mel_min_host == nullptr) != (mel_max_host == nullptr)) return fail("mel_min and mel_max must be given together");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    h->denorm = mel_min_host != nullptr;
    if (h->denorm) {
        const size_t n = sizeof(float) * h->cfg.n_feats;
        USB_CUDA(cudaMemcpy(h->mel_range, mel_min_host, n, cudaMemcpyHostToDevice));
        USB_CUDA(cudaMemcpy(h->mel_range + h->cfg.n_feats, mel_max_host, n, cudaMemcpyHostToDevice));
    }
    return 0;
}

int usb_set_graph_mode(usb_handle* h, int32_t mode) {
    if (!h) return fail("null handle");
    if (mode < -1 || mode > 1) return fail("graph mode must be -1 (auto), 0 (off) or 1 (on)");
    h->graph_mode = mode;
    return 0;
}
int64_t usb_graph_steps(usb_handle* h) { return h ? h->graph_steps : 0; }
int usb_set_splitk_mode(usb_handle* h, int32_t mode) {
    if (!h) return fail("null handle");
    if (mode < -1 || mode > 1) return fail("split-K mode must be -1 (auto), 0 (off) or 1 (on)");
    h->splitk_mode = mode;      // takes effect when the next call plans its workspace
    return 0;
}

int64_t usb_workspace_bytes(usb_handle* h) { return h ? (int64_t)h->plan.arena_bytes : 0; }
int64_t usb_launch_count(usb_handle* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------- operator-level
int usb_op_conv(usb_handle* h, int32_t kind, const void* in0, const void* in1, int32_t N, int32_t H, int32_t W,
                int32_t C0, int32_t C1, int32_t Cout, const float* weight_host, const float* bias_host,
                const float* mask, const void* residual, float res_scale, int64_t* stats, int32_t groups, void* out,
                uint64_t stream) {
    if (!h) return fail("null handle");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (kind < 0 || kind > 3) return fail("bad conv kind");
    const int Cin = C0 + (in1 ? C1 : 0);
    std::vector<__half> packed;
    pack_conv_host(kind, weight_host, Cout, Cin, packed);
    __half* dw = nullptr;
    float* db = nullptr;
    float* dscale = nullptr;
    USB_CUDA(cudaMalloc(&dw, packed.size() * sizeof(__half)));
    USB_CUDA(cudaMemcpy(dw, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice));
    if (bias_host) {
        USB_CUDA(cudaMalloc(&db, Cout * sizeof(float)));
        USB_CUDA(cudaMemcpy(db, bias_host, Cout * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (residual) {
        USB_CUDA(cudaMalloc(&dscale, sizeof(float)));
        USB_CUDA(cudaMemcpy(dscale, &res_scale, sizeof(float), cudaMemcpyHostToDevice));
    }
    ConvEpilogue ep;
    ep.bias = db; ep.stats = reinterpret_cast<long long*>(stats); ep.groups = groups > 0 ? groups : 8; ep.res = static_cast<const __half*>(residual);
    ep.res_scale = dscale; ep.mask = mask;
    ConvOp op;
    int rc = build_conv(op, kind, static_cast<const __half*>(in0), C0, C0, static_cast<const __half*>(in1), C1, C1, N, H,
                        W, dw, kind == KT4 ? 4 : 1, kind == KT4 ? 1 : 0, Cout, ep, static_cast<__half*>(out));
    if (!rc) {
        int e = launch_conv_igemm(op.p, op.a0, op.a1, op.b, op.o, h->num_sms, s);
        h->launches++;
        if (e) rc = fail(std::string("conv launch: ") + cudaGetErrorString((cudaError_t)e));
    }
    cudaError_t se = cudaStreamSynchronize(s);
    if (!rc && se != cudaSuccess) rc = fail(std::string("conv kernel: ") + cudaGetErrorString(se));
    cudaFree(dw);
    if (db) cudaFree(db);
    if (dscale) cudaFree(dscale);
    return rc;
}

// timing experiment: conv on internally allocated buffers, returns average milliseconds over `iters` launches
int usb_dbg_conv_time(usb_handle* h, int32_t kind, int32_t N, int32_t H, int32_t W, int32_t C0, int32_t C1,
                      int32_t Cout, int32_t with_stats, int32_t iters, float* ms_out) {
    if (!h) return fail("null handle");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    const int Cin = C0 + C1;
    const int taps = taps_of(kind), Z = kind == KT4 ? 4 : 1;
    const int Hout = kind == K3S2 ? H / 2 : (kind == KT4 ? 2 * H : H), Wout = kind == K3S2 ? W / 2 : (kind == KT4 ? 2 * W : W);
    __half *a0 = nullptr, *a1 = nullptr, *w = nullptr, *out = nullptr;
    float* bias = nullptr;
    long long* st = nullptr;
    const size_t n0 = (size_t)N * H * W * C0, n1 = (size_t)N * H * W * (C1 ? C1 : 1), nw = (size_t)Z * Cout * taps * Cin;
    const size_t no = (size_t)N * Hout * Wout * Cout;
    USB_CUDA(cudaMalloc(&a0, n0 * 2)); USB_CUDA(cudaMalloc(&a1, n1 * 2)); USB_CUDA(cudaMalloc(&w, nw * 2));
    USB_CUDA(cudaMalloc(&out, no * 2)); USB_CUDA(cudaMalloc(&bias, Cout * 4)); USB_CUDA(cudaMalloc(&st, (size_t)N * 16 * 8));
    USB_CUDA(cudaMemset(a0, 0x3c, n0 * 2)); USB_CUDA(cudaMemset(a1, 0x3c, n1 * 2)); USB_CUDA(cudaMemset(w, 0x1c, nw * 2));
    USB_CUDA(cudaMemset(bias, 0, Cout * 4)); USB_CUDA(cudaMemset(st, 0, (size_t)N * 16 * 8));
    ConvEpilogue ep;
    ep.bias = bias; ep.stats = with_stats ? st : nullptr; ep.groups = 8;
    ConvOp op;
    int rc = build_conv(op, kind, a0, C0, C0, C1 ? a1 : nullptr, C1, C1, N, H, W, w, Z, kind == KT4 ? 1 : 0, Cout, ep, out);
    if (!rc) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 2 && !rc; ++i) rc = launch_conv_igemm(op.p, op.a0, op.a1, op.b, op.o, h->num_sms, 0);
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters && !rc; ++i) rc = launch_conv_igemm(op.p, op.a0, op.a1, op.b, op.o, h->num_sms, 0);
        cudaEventRecord(e1, 0);
        cudaError_t se = cudaDeviceSynchronize();
        if (rc) rc = fail(std::string("conv launch: ") + cudaGetErrorString((cudaError_t)rc));
        else if (se != cudaSuccess) rc = fail(std::string("conv kernel: ") + cudaGetErrorString(se));
        else {
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            *ms_out = ms / iters;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(a0); cudaFree(a1); cudaFree(w); cudaFree(out); cudaFree(bias); cudaFree(st);
    return rc;
}

int usb_op_gn_apply(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta,
                    const float* addvec, const void* res, const float* mask, void* out, int32_t N, int32_t H, int32_t W,
                    int32_t C, int32_t groups, uint64_t stream) {
    if (!h) return fail("null handle");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    GnApplyParams p;
    p.raw = static_cast<const __half*>(raw); p.stats = reinterpret_cast<const long long*>(stats); p.gamma = gamma; p.beta = beta; p.addvec = addvec;
    p.addvec_stride = C; p.res = static_cast<const __half*>(res); p.mask = mask; p.out = static_cast<__half*>(out);
    p.N = N; p.P = H * W; p.W = W; p.C = C; p.groups = groups; p.eps = 1e-5f; p.dbg = 0;
    USB_LAUNCH(h, launch_gn_apply(p, h->num_sms, reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}

int usb_op_attn_context(usb_handle* h, const void* qkv, const float* wo, void* weff, int32_t N, int32_t P, int32_t C,
                        int32_t heads, uint64_t stream) {
    if (!h) return fail("null handle");
    USB_CUDA(cudaSetDevice(h->cfg.device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    AttnParams ap;
    memset(&ap, 0, sizeof ap);
    ap.qkv = static_cast<const __half*>(qkv); ap.wo = wo; ap.weff = static_cast<__half*>(weff); ap.N = N; ap.P = P;
    ap.C = C; ap.heads = heads; ap.chunk = attn_chunk_for(P); ap.ld = 3 * heads * 32; ap.koff = heads * 32; ap.voff = 2 * heads * 32;
    const size_t part_bytes = attn_scratch_bytes(N, heads, P, attn_chunk_for(P));
    float* part = nullptr;
    USB_CUDA(cudaMalloc(&part, part_bytes));
    ap.part = part;
    int e = launch_attn_context(ap, s);
    h->launches += 3;
    cudaError_t se = cudaStreamSynchronize(s);
    cudaFree(part);
    if (e) return fail(std::string("attention launch: ") + cudaGetErrorString((cudaError_t)e));
    if (se != cudaSuccess) return fail(std::string("attention kernel: ") + cudaGetErrorString(se));
    return 0;
}

// ===================================================================================================================
// fine-tune step, operator level (include/unitspeech_b200_train.h).  Everything is enqueued on `stream`; no hidden
// synchronisation, no allocation.  The Python host (unitspeech_b200/training.py) owns all buffers.
// ===================================================================================================================
#define USB_T_BEGIN()                                         \
    if (!h) return fail("null handle");                       \
    USB_CUDA(cudaSetDevice(h->cfg.device));                   \
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream)

int usb_t_pack_conv(usb_handle* h, int32_t kind, const float* w, int32_t Cout, int32_t Cin, int32_t ci0, int32_t ci1,
                    void* fwd, void* dgrad, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_pack_conv(kind, w, Cout, Cin, ci0, ci1, static_cast<__half*>(fwd), static_cast<__half*>(dgrad), s, h->pack));
    return 0;
}

int usb_t_pack_begin(usb_handle* h, uint64_t stream) {
    (void)stream;
    if (!h) return fail("null handle");
    if (!h->pack) h->pack = pack_batch_create();
    pack_batch_begin(h->pack);
    return 0;
}

int usb_t_pack_flush(usb_handle* h, uint64_t stream) {
    USB_T_BEGIN();
    if (!h->pack) return fail("usb_t_pack_flush without usb_t_pack_begin");
    USB_LAUNCH(h, pack_batch_flush(h->pack, s));
    return 0;
}

int usb_t_cast(usb_handle* h, const float* src, void* dst, int64_t n, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_cast_h(src, static_cast<__half*>(dst), n, s));
    return 0;
}

int usb_t_conv(usb_handle* h, int32_t kind, const void* in0, int32_t C0tot, int32_t C0, const void* in1, int32_t C1tot,
               int32_t C1, int32_t N, int32_t H, int32_t W, const void* w, int32_t wZ, int32_t b_batch_mode, int32_t Cout,
               const float* bias, const float* mask, const void* res, const float* res_scale, int64_t* stats, int32_t groups,
               void* out, uint64_t stream) {
    USB_T_BEGIN();
    if (kind < 0 || kind > 5) return fail("bad conv kind");
    ConvEpilogue ep;
    ep.bias = bias; ep.stats = reinterpret_cast<long long*>(stats); ep.groups = groups > 0 ? groups : 8;
    ep.res = static_cast<const __half*>(res); ep.res_scale = res_scale; ep.mask = mask;
    // the fine-tune step runs on 8 short crops: the level-2/3 convolutions (forward and data gradient) have 20-100 output
    // tiles for 148 SMs, so they split K (no batch-invariance contract here: the factor follows the launch's own tile count).
    // All usb_t_conv launches are issued on one stream, so one workspace serves them.  USB_FT_NO_SPLITK=1 disables it.
    static const bool no_split = getenv("USB_FT_NO_SPLITK") != nullptr;
    if (!no_split) {
        const long long items = 2LL * h->num_sms;
        if (!h->t_kpart) {
            void* p = nullptr;
            USB_CUDA(cudaMalloc(&p, (size_t)items * 2 * 128 * 128 * sizeof(float) + (size_t)items * sizeof(int)));
            h->dev_allocs.push_back(p);
            h->t_kpart = static_cast<float*>(p);
            h->t_ktick = reinterpret_cast<int*>(h->t_kpart + (size_t)items * 2 * 128 * 128);
            USB_CUDA(cudaMemset(h->t_ktick, 0, (size_t)items * sizeof(int)));
        }
        ep.splitk = true; ep.kpart = h->t_kpart; ep.ktick = h->t_ktick; ep.kpart_items = items; ep.num_sms = h->num_sms;
        ep.split_rows = 0;
    }
    ConvOp op;
    USB_TRY(build_conv(op, kind, static_cast<const __half*>(in0), C0tot, C0, static_cast<const __half*>(in1), C1tot, C1, N, H, W,
                       static_cast<const __half*>(w), wZ, b_batch_mode, Cout, ep, static_cast<__half*>(out)));
    USB_LAUNCH(h, launch_conv_igemm(op.p, op.a0, op.a1, op.b, op.o, h->num_sms, s));
    return 0;
}

int usb_t_first_conv(usb_handle* h, const float* x, const float* mu, const int32_t* rows, const float* mask, const float* w3,
                     const float* b3, const float* w1, const float* b1, void* raw, void* res, int64_t* stats, int32_t N,
                     int32_t H, int32_t W, int32_t C, uint64_t stream) {
    USB_T_BEGIN();
    FirstConvParams f;
    memset(&f, 0, sizeof f);
    f.x = x; f.cond = mu; f.x_row = rows; f.mu_row = rows; f.mask = mask; f.w3 = w3; f.b3 = b3; f.w1 = w1; f.b1 = b1;
    f.raw = static_cast<__half*>(raw); f.res = static_cast<__half*>(res); f.stats = reinterpret_cast<long long*>(stats);
    f.N = N; f.H = H; f.W = W; f.C = C; f.groups = h->cfg.groups;
    USB_LAUNCH(h, launch_first_conv(f, s));
    return 0;
}

int usb_t_gn_apply(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta,
                   const float* addvec, int64_t addvec_stride, const void* res, const float* mask, void* out, int32_t N,
                   int32_t H, int32_t W, int32_t C, uint64_t stream) {
    USB_T_BEGIN();
    GnApplyParams p;
    p.raw = static_cast<const __half*>(raw); p.stats = reinterpret_cast<const long long*>(stats); p.gamma = gamma; p.beta = beta;
    p.addvec = addvec; p.addvec_stride = addvec_stride; p.res = static_cast<const __half*>(res); p.mask = mask;
    p.out = static_cast<__half*>(out); p.N = N; p.P = H * W; p.W = W; p.C = C; p.groups = h->cfg.groups; p.eps = 1e-5f; p.dbg = 0;
    USB_LAUNCH(h, launch_gn_apply(p, h->num_sms, s));
    return 0;
}

int usb_t_embed(usb_handle* h, const float* t, const float* spk, const float* freqs, const float* w0, const float* b0,
                const float* w2, const float* b2, const float* wcat, const float* bcat, float* u, float* e, int32_t N,
                int32_t J, uint64_t stream) {
    USB_T_BEGIN();
    EmbedParams ep;
    memset(&ep, 0, sizeof ep);
    ep.t = t; ep.spk = spk; ep.freqs = freqs; ep.w0 = w0; ep.b0 = b0; ep.w2 = w2; ep.b2 = b2; ep.wcat = wcat; ep.bcat = bcat;
    ep.u = u; ep.e = e; ep.N = N; ep.dim = h->cfg.dim; ep.S = h->cfg.spk_emb_dim; ep.J = J; ep.pe_scale = h->cfg.pe_scale;
    USB_LAUNCH(h, launch_embed(ep, s));
    h->launches++;
    return 0;
}

int64_t usb_t_attn_scratch_bytes(int32_t N, int32_t heads, int32_t P) {
    return static_cast<int64_t>(attn_scratch_bytes(N, heads, P, attn_chunk_for(P)));
}

int usb_t_attn_context(usb_handle* h, const void* qkv, int32_t ld, int32_t koff, int32_t voff, const float* wo, float* scratch,
                       void* weff, float* ctx_out, float* stat_out, int32_t N, int32_t P, int32_t C, int32_t heads,
                       uint64_t stream) {
    USB_T_BEGIN();
    AttnParams ap;
    memset(&ap, 0, sizeof ap);
    ap.qkv = static_cast<const __half*>(qkv); ap.ld = ld; ap.koff = koff; ap.voff = voff; ap.wo = wo; ap.part = scratch;
    ap.weff = static_cast<__half*>(weff); ap.ctx_out = ctx_out; ap.stat_out = stat_out; ap.N = N; ap.P = P; ap.C = C;
    ap.heads = heads; ap.chunk = attn_chunk_for(P);
    USB_LAUNCH(h, launch_attn_context(ap, s));
    h->launches += 2;
    return 0;
}

int usb_t_final(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta, const float* wf,
                const float* bf, const float* mask, float* score, int32_t N, int32_t H, int32_t W, int32_t C, uint64_t stream) {
    USB_T_BEGIN();
    FinalParams f;
    memset(&f, 0, sizeof f);
    f.raw = static_cast<const __half*>(raw); f.stats = reinterpret_cast<const long long*>(stats); f.gamma = gamma; f.beta = beta;
    f.wf = wf; f.bf = bf; f.mask = mask; f.B = N; f.nb = 1; f.P = H * W; f.W = W; f.C = C; f.groups = h->cfg.groups; f.eps = 1e-5f;
    f.score = score;
    USB_LAUNCH(h, launch_final(f, h->num_sms, s));
    return 0;
}

int usb_t_loss(usb_handle* h, const float* score, const float* zm, const float* mask, const float* t, double* partial,
               float* loss, int32_t B, int32_t T, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_diffusion_loss(score, zm, mask, t, h->cfg.beta_min, h->cfg.beta_max, partial, loss, B, h->cfg.n_feats, T, s));
    h->launches++;
    return 0;
}

int usb_t_loss_grad(usb_handle* h, const float* score, const float* zm, const float* mask, const float* t, float loss_scale,
                    float* msum, float* dscore, int32_t B, int32_t T, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_loss_grad(score, zm, mask, t, h->cfg.beta_min, h->cfg.beta_max, loss_scale, msum, dscore, B,
                                   h->cfg.n_feats, T, s));
    h->launches++;
    return 0;
}

int usb_t_gn_bwd(usb_handle* h, const void* raw, const int64_t* stats, const float* gamma, const float* beta, const void* dy0,
                 const void* dy1, const float* dys, const float* wvec, const float* mask, float* scratch, void* d_raw,
                 float* dbias, float* dgamma, float* dbeta, float* d_emb, int64_t emb_stride, float* d_wvec, int32_t N,
                 int32_t H, int32_t W, int32_t C, uint64_t stream) {
    USB_T_BEGIN();
    const int G = h->cfg.groups;
    GnBwdParams p;
    memset(&p, 0, sizeof p);
    p.raw = static_cast<const __half*>(raw); p.stats = reinterpret_cast<const long long*>(stats); p.gamma = gamma; p.beta = beta;
    p.dy0 = static_cast<const __half*>(dy0); p.dy1 = static_cast<const __half*>(dy1); p.dys = dys; p.wvec = wvec; p.mask = mask;
    p.sums = scratch;                                   // [3][N][C]
    p.dgamma = dgamma; p.dbeta = dbeta; p.d_emb = d_emb; p.emb_stride = emb_stride; p.d_wvec = d_wvec;
    p.d_raw = static_cast<__half*>(d_raw); p.dbias = dbias; p.N = N; p.P = H * W; p.W = W; p.C = C; p.groups = G; p.eps = 1e-5f;
    if ((dy0 == nullptr) == (dys == nullptr)) return fail("gn_bwd needs exactly one of dy0 / dys");
    USB_CUDA(cudaMemsetAsync(scratch, 0, static_cast<size_t>(3) * N * C * sizeof(float), s));
    USB_LAUNCH(h, launch_gn_bwd_reduce(p, h->num_sms, s));
    USB_LAUNCH(h, launch_gn_bwd_apply(p, h->num_sms, s));
    return 0;
}

int usb_t_colsum(usb_handle* h, const void* t, int32_t ld, int32_t N, int32_t P, int32_t C, float* out, int64_t out_stride_n,
                 uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_colsum(static_cast<const __half*>(t), ld, N, P, C, out, out_stride_n, h->num_sms, s));
    return 0;
}

int usb_t_add(usb_handle* h, const void* a, const void* b, const void* c, void* out, int64_t n, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_add_h(static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<const __half*>(c),
                               static_cast<__half*>(out), n, s));
    return 0;
}

// weight gradient of a convolution of the given kind, written in the reference's parameter layout.
// dy: output gradient (N, Hout, Wout, ldy) of which Cout channels are used; x: layer input (N, H, W, ldx) of which
// Cs = channels [0, Cs) are used and land in the [ci0, ci0 + Cs) slice of the Cin_total input channels of dW.
// per_sample != 0 (kind 2 only): dW is (N, Cout, Cs), one matrix per sample.
int usb_t_wgrad(usb_handle* h, int32_t kind, const void* dy, int32_t ldy, const void* x, int32_t ldx, int32_t N, int32_t H,
                int32_t W, int32_t Cout, int32_t Cs, int32_t ci0, int32_t Cin_total, float* dW, int32_t per_sample,
                uint64_t stream) {
    USB_T_BEGIN();
    if (kind < 0 || kind > 3) return fail("bad conv kind");
    const int dst_is_zero = (per_sample >> 1) & 1;   // flag bit 1: the destination slice holds zeros (fresh zero_grad)
    per_sample &= 1;
    if (per_sample && kind != K1) return fail("per-sample weight gradients are 1x1 only");
    // output addressing in the training layout (train.h: launch_pack_conv), input channels contiguous
    long long s_co, s_ci = 1, s_n = 0;
    int taps, tap_off[kWgradMaxTaps] = {0};
    float* dWo = dW + ci0;
    if (kind == KT4) {   // [phase][Cout][a*2+b][Cin]; tap (kh, kw): kh 1,3 -> phase row 0 (a = 0,1), kh 0,2 -> phase row 1
        taps = 16; s_co = 4LL * Cin_total;
        for (int kh = 0; kh < 4; ++kh)
            for (int kw = 0; kw < 4; ++kw) {
                const int ph = (kh & 1) ? 0 : 1, a = kh >> 1, pw = (kw & 1) ? 0 : 1, b = kw >> 1;
                tap_off[kh * 4 + kw] = static_cast<int>((static_cast<long long>(ph * 2 + pw) * Cout * 4 + (a * 2 + b)) * Cin_total);
            }
    } else if (kind == K1) {
        taps = 1; s_co = per_sample ? Cs : Cin_total;
        s_n = per_sample ? static_cast<long long>(Cout) * Cs : 0;
        if (per_sample) dWo = dW;
    } else {             // [Cout][9][Cin]
        taps = 9; s_co = 9LL * Cin_total;
        for (int t = 0; t < 9; ++t) tap_off[t] = t * Cin_total;
    }
    static const bool use_mma_sync = getenv("USB_WGRAD_MMA") != nullptr;
    if (!use_mma_sync && Cout % 64 == 0 && Cs % 64 == 0 && ldy % 8 == 0 && ldx % 8 == 0) {
        // tcgen05 path: iteration image = the smaller of the two images; the other side is addressed through the
        // parity-split view with the forward conv's tap table
        WgradTcParams q;
        memset(&q, 0, sizeof q);
        const int Hi = kind == K3S2 ? H / 2 : H, Wi = kind == K3S2 ? W / 2 : W;
        int bestBH = 0, bestBW = 0;
        double best = -1.0;
        for (int bh = 1; bh <= Hi && bh <= 128; ++bh) {
            if (Hi % bh) continue;
            for (int bw = 1; bw <= 128; bw *= 2) {
                const int kp = bh * bw;
                if (kp % 16 || kp > 128 || kp < 32) continue;
                const double eff = static_cast<double>(Wi) / (((Wi + bw - 1) / bw) * bw) + 1e-4 * kp;
                if (eff > best) { best = eff; bestBH = bh; bestBW = bw; }
            }
        }
        if (bestBH == 0) return fail("wgrad: no pixel tiling for this image");
        q.N = N; q.BH = bestBH; q.BW = bestBW; q.tiles_y = Hi / bestBH; q.tiles_x = (Wi + bestBW - 1) / bestBW;
        q.taps = taps;
        for (int t = 0; t < taps; ++t) {
            q.tap_off[t] = tap_off[t];
            if (kind == K3S1) {
                q.btap[t].dy = (int8_t)(t / 3 - 1); q.btap[t].dx = (int8_t)(t % 3 - 1);
            } else if (kind == K3S2) {
                const int kh = t / 3, kw = t % 3;
                q.btap[t].dy = (int8_t)(kh == 0 ? -1 : 0); q.btap[t].p = (int8_t)(kh == 1 ? 0 : 1);
                q.btap[t].dx = (int8_t)(kw == 0 ? -1 : 0); q.btap[t].c = (int16_t)(kw == 1 ? 0 : ldx);
            } else if (kind == KT4) {
                const int kh = t / 4, kw = t % 4;
                q.atap[t].dy = (int8_t)(kh == 0 ? -1 : (kh == 3 ? 1 : 0)); q.atap[t].p = (int8_t)((kh & 1) ? 0 : 1);
                q.atap[t].dx = (int8_t)(kw == 0 ? -1 : (kw == 3 ? 1 : 0)); q.atap[t].c = (int16_t)((kw & 1) ? 0 : ldy);
            }
        }
        if ((kind == K3S2 && (ldx != Cs || 2 * ldx > 32767)) || (kind == KT4 && (ldy != Cout || 2 * ldy > 32767)))
            return fail("wgrad: strided convs need dense tensors");
        q.Cout = Cout; q.Cin = Cs; q.n_tile = Cs <= 128 ? 128 : 256;
        q.tiles_m = (Cout + 127) / 128; q.tiles_n = (Cs + q.n_tile - 1) / q.n_tile;
        q.dW = dWo; q.s_co = s_co; q.s_ci = s_ci; q.s_n = s_n; q.overwrite = dst_is_zero;
        CUtensorMap ma, mb;
        USB_TRY(load_encode_fn());
        const int Hy = kind == K3S2 ? H / 2 : (kind == KT4 ? 2 * H : H), Wy = kind == K3S2 ? W / 2 : (kind == KT4 ? 2 * W : W);
        USB_TRY(make_act_map(&ma, static_cast<const __half*>(dy), N, Hy, Wy, ldy, Cout, kind == KT4, q.BH, q.BW));
        USB_TRY(make_act_map(&mb, static_cast<const __half*>(x), N, H, W, ldx, Cs, kind == K3S2, q.BH, q.BW));
        USB_LAUNCH(h, launch_wgrad_tc(q, ma, mb, h->num_sms, s));
        return 0;
    }
    WgradParams p;
    memset(&p, 0, sizeof p);
    p.A = static_cast<const __half*>(dy); p.B = static_cast<const __half*>(x); p.lda = ldy; p.ldb = ldx;
    p.N = N; p.Cout = Cout; p.Cin = Cs; p.a_mul = p.b_mul = 1;
    p.Hb = H; p.Wb = W;
    if (kind == K3S1 || kind == K1) {
        p.Ha = H; p.Wa = W; p.Hi = H; p.Wi = W;
    } else if (kind == K3S2) {
        p.Ha = H / 2; p.Wa = W / 2; p.Hi = H / 2; p.Wi = W / 2; p.b_mul = 2;
    } else {
        p.Ha = 2 * H; p.Wa = 2 * W; p.Hi = H; p.Wi = W; p.a_mul = 2;
    }
    p.taps = taps; p.s_co = s_co; p.s_ci = s_ci; p.s_n = s_n; p.dW = dWo;
    for (int t = 0; t < taps; ++t) {
        p.tap_off[t] = tap_off[t];
        if (kind == KT4) { p.ady[t] = (int8_t)(t / 4 - 1); p.adx[t] = (int8_t)(t % 4 - 1); }
        else if (kind != K1) { p.bdy[t] = (int8_t)(t / 3 - 1); p.bdx[t] = (int8_t)(t % 3 - 1); }
    }
    USB_LAUNCH(h, launch_wgrad(p, h->num_sms, s));
    return 0;
}

int usb_t_first_conv_wgrad(usb_handle* h, const void* d_raw, const void* d_res0, const void* d_res1, const float* x,
                           const float* mu, const float* mask, float* dW3, float* dW1, int32_t N, int32_t H, int32_t W,
                           int32_t C, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_first_conv_wgrad(static_cast<const __half*>(d_raw), static_cast<const __half*>(d_res0),
                                          static_cast<const __half*>(d_res1), x, mu, mask, dW3, dW1, N, H, W, C, h->num_sms, s));
    return 0;
}

int usb_t_attn_bwd_small(usb_handle* h, const float* G, const float* cs, const float* wo, const float* bo, const float* g,
                         const float* ctx, float* dwo, float* dbo, float* dg, float* dctx, void* weffT, int32_t N, int32_t C,
                         int32_t heads, uint64_t stream) {
    USB_T_BEGIN();
    AttnBwdParams p;
    memset(&p, 0, sizeof p);
    p.G = G; p.cs = cs; p.wo = wo; p.bo = bo; p.g = g; p.ctx = ctx; p.dwo = dwo; p.dbo = dbo; p.dg = dg; p.dctx = dctx;
    p.weffT = static_cast<__half*>(weffT); p.N = N; p.C = C; p.heads = heads;
    USB_CUDA(cudaMemsetAsync(dctx, 0, static_cast<size_t>(N) * heads * 32 * 32 * sizeof(float), s));
    USB_LAUNCH(h, launch_attn_bwd_small(p, s));
    return 0;
}

int usb_t_attn_bwd_dkv(usb_handle* h, const void* qkv, int32_t ld, int32_t koff, int32_t voff, const float* ms,
                       const float* ctx, const float* dctx, void* dkv, int32_t N, int32_t P, int32_t heads, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_attn_bwd_dkv(static_cast<const __half*>(qkv), ld, koff, voff, ms, ctx, dctx, static_cast<__half*>(dkv),
                                      N, P, heads, h->num_sms, s));
    return 0;
}

int usb_t_embed_bwd(usb_handle* h, const float* t, const float* spk, const float* freqs, const float* w0, const float* b0,
                    const float* w2, const float* b2, const float* wcat, const float* u, const float* dE, float* du,
                    float* dw0, float* db0, float* dw2, float* db2, float* dwcat, float* dbcat, int32_t N, int32_t J,
                    uint64_t stream) {
    USB_T_BEGIN();
    EmbedBwdParams p;
    memset(&p, 0, sizeof p);
    p.t = t; p.spk = spk; p.freqs = freqs; p.w0 = w0; p.b0 = b0; p.w2 = w2; p.b2 = b2; p.wcat = wcat; p.u = u; p.dE = dE;
    p.du = du; p.dw0 = dw0; p.db0 = db0; p.dw2 = dw2; p.db2 = db2; p.dwcat = dwcat; p.dbcat = dbcat; p.N = N;
    p.dim = h->cfg.dim; p.S = h->cfg.spk_emb_dim; p.J = J; p.pe_scale = h->cfg.pe_scale;
    USB_LAUNCH(h, launch_embed_bwd(p, s));
    h->launches += 2;
    return 0;
}

int usb_t_dot(usb_handle* h, const float* a, const float* b, int64_t n, float* out, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_dot(a, b, n, out, s));
    return 0;
}

int usb_t_sumsq(usb_handle* h, const float* g, int64_t n, double* out, uint64_t stream) {
    USB_T_BEGIN();
    USB_LAUNCH(h, launch_sumsq(g, n, out, s));
    return 0;
}

int usb_t_adam(usb_handle* h, float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
               float eps, int32_t step, int32_t* step_dev, const double* sumsq, float inv_scale, float max_norm,
               int32_t* skipped, uint64_t stream) {
    USB_T_BEGIN();
    if (step < 1 && !step_dev) return fail("Adam step counts from 1");
    AdamParams a;
    a.p = p; a.g = g; a.m = m; a.v = v; a.n = n; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    a.bc1 = static_cast<float>(1.0 - std::pow(static_cast<double>(beta1), step > 0 ? step : 1));
    a.bc2 = static_cast<float>(1.0 - std::pow(static_cast<double>(beta2), step > 0 ? step : 1));
    a.step_dev = step_dev;
    a.sumsq = sumsq; a.inv_scale = inv_scale; a.max_norm = max_norm; a.skipped = skipped;
    USB_LAUNCH(h, launch_adam(a, h->num_sms, s));
    return 0;
}

}  // extern "C"
