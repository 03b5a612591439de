// Streaming kernels of the score estimator: input conv, fused GroupNorm-apply + Mish + embedding/residual + mask,
// the fused "final block -> 1x1 -> guidance combine -> posterior update" step, and the time/speaker embedding.
// All are HBM-bound: 16-byte vector accesses, one channel octet per thread, consecutive threads on consecutive
// addresses, grids sized in multiples of the SM count.
#include <cstdlib>

#include "kernels.h"
#include "pdl.h"

namespace usb {

namespace {

// single-instruction special functions (MUFU.EX2 / MUFU.RCP, flush-to-zero): the non-ftz intrinsics expand into
// denormal-scaling sequences that made the streaming kernels issue-bound
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;

// Packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2, two IEEE fp32 operations per issue slot; each lane rounds
// exactly like the scalar instruction).  The streaming kernels below are issue-bound, not bandwidth-bound, so halving the
// slots of their FMA work is what moves them towards the HBM roofline.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// x * tanh(softplus(x)) = x * w / (w + 2) with w = e^x (e^x + 2); exact for x > 20 (ratio rounds to 1), which is
// the reference's softplus threshold (unitspeech.py:13-15).
__device__ __forceinline__ float mish_fast(float x) {
    const float e = ex2_ftz(fminf(x, 20.f) * kLog2e);
    const float w = e * (e + 2.f);
    return x * w * rcp_ftz(w + 2.f);
}
__device__ __forceinline__ float mish_precise(float x) {
    const float sp = x > 20.f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h[i]);
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {   // (lo = a, hi = b), +-inf clamped to +-65504
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = pack2(f[0], f[1]);
    o.y = pack2(f[2], f[3]);
    o.z = pack2(f[4], f[5]);
    o.w = pack2(f[6], f[7]);
    return o;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// mean / rstd of GroupNorm group g of row n from the fixed-point (sum, sumsq); biased variance, as torch GroupNorm
__device__ __forceinline__ void group_moments(const long long* stats, int n, int groups, int g, double count,
                                              float eps, float& mean, float& rstd) {
    const double s = static_cast<double>(stats[(static_cast<long long>(n) * groups + g) * 2]) /
                     static_cast<double>(kStatSumScale);
    const double ss = static_cast<double>(stats[(static_cast<long long>(n) * groups + g) * 2 + 1]) /
                      static_cast<double>(kStatSqScale);
    const double mu = s / count;
    double var = ss / count - mu * mu;
    var = var < 0.0 ? 0.0 : var;
    mean = static_cast<float>(mu);
    rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

}  // namespace

// =====================================================================================================================
// input conv
// =====================================================================================================================
// One block = one image row segment of up to 64 pixels: the masked 3 x 66 x 2 input halo is staged in shared memory
// once; each warp then walks 8 pixels, every lane producing CL consecutive output channels (weights in registers),
// so a warp writes one pixel's contiguous channel vector per store instruction.
template <int CL>
__global__ void __launch_bounds__(256) first_conv_kernel(const FirstConvParams p, int tiles_per_block) {
    constexpr int TW = 64;
    __shared__ float tile[3][TW + 2][2];
    __shared__ unsigned long long gsum[16];
    pdl_wait_and_trigger();
    const int n = blockIdx.y;
    const int W = p.W, H = p.H, P = H * W, C = p.C;
    const int tiles_x = (W + TW - 1) / TW;
    const int n_tiles = tiles_x * H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xr = p.x_row[n], mr = p.mu_row[n];
    const float* xs = p.x + static_cast<long long>(xr) * P;
    const float* ms = mr >= 0 ? p.cond + static_cast<long long>(mr) * P : nullptr;
    const float* mk = p.mask + static_cast<long long>(n) * W;
    if (threadIdx.x < 16) gsum[threadIdx.x] = 0ull;
    const int passes = C / (32 * CL);
    const int t_begin = blockIdx.x * tiles_per_block, t_end = min(n_tiles, t_begin + tiles_per_block);
    for (int ps = 0; ps < passes; ++ps) {
        const int c0 = ps * 32 * CL + lane * CL;
        float w3[18][CL], w1[2][CL], b3[CL], b1[CL];
#pragma unroll
        for (int k = 0; k < 18; ++k)
#pragma unroll
            for (int i = 0; i < CL; ++i) w3[k][i] = __ldg(p.w3 + k * C + c0 + i);
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int i = 0; i < CL; ++i) w1[k][i] = __ldg(p.w1 + k * C + c0 + i);
#pragma unroll
        for (int i = 0; i < CL; ++i) {
            b3[i] = __ldg(p.b3 + c0 + i);
            b1[i] = __ldg(p.b1 + c0 + i);
        }
        // GroupNorm partials: one fp32 partial per (warp, row tile, group) -- a partition that does not depend on
        // the launch geometry -- converted to fixed point and accumulated as integers (bitwise batch invariance)
        const int cpg = C / p.groups;
        const int lpg = cpg / CL;  // lanes per group (power of two, <= 32); lanes of one group are adjacent
        long long is = 0, iss = 0;
        for (int t = t_begin; t < t_end; ++t) {
            float s = 0.f, ss = 0.f;
            const int y = t / tiles_x;
            const int x0 = (t - y * tiles_x) * TW;
            __syncthreads();   // previous tile fully consumed
            for (int i = threadIdx.x; i < 3 * (TW + 2); i += blockDim.x) {
                const int r = i / (TW + 2), c = i - r * (TW + 2);
                const int yy = y + r - 1, xx = x0 + c - 1;
                float vm = 0.f, vx = 0.f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float m = __ldg(mk + xx);
                    vm = (ms ? __ldg(ms + yy * W + xx) : __ldg(p.text_uncon + yy)) * m;   // channel 0 = mu
                    vx = __ldg(xs + yy * W + xx) * m;                                     // channel 1 = x_t
                }
                tile[r][c][0] = vm;
                tile[r][c][1] = vx;
            }
            __syncthreads();
            for (int j = 0; j < TW / 8; ++j) {
                const int px = warp * (TW / 8) + j;
                const int x = x0 + px;
                if (x >= W) break;
                float acc[CL], rr[CL];
#pragma unroll
                for (int i = 0; i < CL; ++i) {
                    acc[i] = b3[i];
                    rr[i] = b1[i];
                }
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float2 v = *reinterpret_cast<const float2*>(&tile[dy][px + dx][0]);
#pragma unroll
                        for (int i = 0; i < CL; ++i) {
                            acc[i] = fmaf(v.x, w3[(dy * 3 + dx) * 2][i], acc[i]);
                            acc[i] = fmaf(v.y, w3[(dy * 3 + dx) * 2 + 1][i], acc[i]);
                        }
                    }
                {
                    const float2 v = *reinterpret_cast<const float2*>(&tile[1][px + 1][0]);
#pragma unroll
                    for (int i = 0; i < CL; ++i) rr[i] = fmaf(v.y, w1[1][i], fmaf(v.x, w1[0][i], rr[i]));
                }
#pragma unroll
                for (int i = 0; i < CL; ++i) {
                    s += acc[i];
                    ss = fmaf(acc[i], acc[i], ss);
                }
                const long long o = (static_cast<long long>(n) * P + static_cast<long long>(y) * W + x) * C + c0;
                if (CL == 4) {
                    uint2 a, r;
                    a.x = pack2(acc[0], acc[1]); a.y = pack2(acc[2], acc[3]);
                    r.x = pack2(rr[0], rr[1]);   r.y = pack2(rr[2], rr[3]);
                    *reinterpret_cast<uint2*>(p.raw + o) = a;
                    *reinterpret_cast<uint2*>(p.res + o) = r;
                } else {
                    *reinterpret_cast<uint32_t*>(p.raw + o) = pack2(acc[0], acc[1]);
                    *reinterpret_cast<uint32_t*>(p.res + o) = pack2(rr[0], rr[1]);
                }
            }
            for (int o = lpg >> 1; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            is += __float2ll_rn(s * kStatSumScale);
            iss += __float2ll_rn(ss * kStatSqScale);
        }
        if ((lane & (lpg - 1)) == 0) {
            const int g = c0 / cpg;
            atomicAdd(&gsum[g * 2], static_cast<unsigned long long>(is));
            atomicAdd(&gsum[g * 2 + 1], static_cast<unsigned long long>(iss));
        }
    }
    __syncthreads();
    if (threadIdx.x < p.groups * 2)
        atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + static_cast<long long>(n) * p.groups * 2 + threadIdx.x,
                  gsum[threadIdx.x]);
}

int launch_first_conv(const FirstConvParams& p, cudaStream_t s) {
    if (p.groups != 8 || !(p.C == 64 || p.C == 128 || p.C == 256)) return (int)cudaErrorInvalidValue;
    const int n_tiles = ((p.W + 63) / 64) * p.H;
    int tpb = 16;   // row tiles per block: amortises the per-lane weight loads; keep >= ~8 blocks per SM
    while (tpb > 1 && (long long)((n_tiles + tpb - 1) / tpb) * p.N < 148 * 8) tpb >>= 1;
    dim3 grid((n_tiles + tpb - 1) / tpb, p.N);
    // (a packed fp32x2 version of this kernel was measured and dropped: its 44 packed weights need 163 registers, and it is
    // slower than this one both at one block per SM and at two with spills -- profiles/README.md)
    if (p.C == 64) return (int)launch_k(first_conv_kernel<2>, grid, dim3(256), 0, s, p, tpb);
    return (int)launch_k(first_conv_kernel<4>, grid, dim3(256), 0, s, p, tpb);
}

// =====================================================================================================================
// GroupNorm apply + Mish (+ embedding vector) (+ residual) * mask
// =====================================================================================================================
// out = (Mish(a*v + b) + add + res) * m for 8 channels.  Mish(x) = x * tanh(softplus(x)) = x - 2x/d with d = e^x (e^x + 2) + 2
// (exact for x > 20, the reference's softplus threshold, where 2/d rounds to 0).  The kernel is issue-bound, not
// HBM-bound, unless the instruction count per element stays ~14 (ncu, profiles/): so
//   * d is built negated and halved, dh = -d/2 = e (-e/2 - 1) - 1, and ONE reciprocal serves two channels:
//     -2/d0 = dh1 * rcp(dh0 * dh1) (|dh| <= ~1.2e17 each: the product stays finite) -- no separate scale by -2;
//   * the exponent argument is min(x, 20) * log2(e) from the normalised value itself (no second affine pair in registers);
//   * the mask multiply is skipped for unmasked pixels (m == 1, warp-uniform almost everywhere).
// a2 = a log2(e), b2 = b log2(e): the exponent argument t = log2(e) x comes straight out of one FMA on the raw value,
// and the normalised value x = t ln(2) is never formed: with dq = -((1 + e)^2 + 1) / (2 ln 2) built from e by two FMAs,
// 1/dq = -2 ln(2) / d, so Mish(x) + add = x - 2x/d + add = t / dq + (a v + b + add).  One reciprocal serves a channel
// pair (1/dq0 = dq1 / (dq0 dq1)); the clamp of t keeps dq0 dq1 finite.
template <bool HAS_RES>
__device__ __forceinline__ uint4 gn_mish8(const uint4& rv, const uint4& rr, const f32x2 (&a)[4], const f32x2 (&a2)[4],
                                          const f32x2 (&b2)[4], const f32x2 (&ba)[4], f32x2 m2) {
    const __half2* hv = reinterpret_cast<const __half2*>(&rv);
    const __half2* hr = reinterpret_cast<const __half2*>(&rr);
    constexpr float kC = 0.72134752044448170368f;      // 1 / (2 ln 2)
    const f32x2 mc = pk2(-kC, -kC), m2c = pk2(-2.f * kC, -2.f * kC);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {            // channel pair (2i, 2i+1)
        const float2 vf = __half22float2(hv[i]);
        const f32x2 v = pk2(vf.x, vf.y);
        float t0, t1;
        const f32x2 t = fma2(v, a2[i], b2[i]);
        upk2(t, t0, t1);
        const float e0 = ex2_ftz(fminf(t0, 20.f * kLog2e)), e1 = ex2_ftz(fminf(t1, 20.f * kLog2e));
        const f32x2 e = pk2(e0, e1);
        float d0, d1;
        upk2(fma2(e, fma2(e, mc, m2c), m2c), d0, d1);            // dq = -(e^2 + 2e + 2) / (2 ln 2)
        const float rn = rcp_ftz(d0 * d1);
        const f32x2 q = pk2(d1 * rn, d0 * rn);                    // (1/dq0, 1/dq1)
        f32x2 y = fma2(t, q, fma2(v, a[i], ba[i]));               // x - 2x/d + (x + add)  (ba = b + add)
        if (HAS_RES) {
            const float2 rf = __half22float2(hr[i]);
            y = add2(y, pk2(rf.x, rf.y));
        }
        y = mul2(y, m2);      // unconditional: a test for m == 1 costs a predicate and register copies per pair
        float y0, y1;
        upk2(y, y0, y1);
        ow[i] = pack2(y0, y1);
    }
    return o;
}

// lanes_mod = (pixels a thread advances per load) mod W, so the column of the next pixel is one add and one conditional
// subtract
// U = pixels (16-byte loads per input tensor) in flight per thread, MINB = blocks per SM the register budget is cut for
template <bool HAS_RES, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) gn_apply_kernel(const GnApplyParams p, int pix_per_block, int TP, int lanes_mod) {
    pdl_wait_and_trigger();
    const int n = blockIdx.y;
    const int tq = threadIdx.x % TP;
    const int pl = threadIdx.x / TP;
    const int lanes = blockDim.x / TP;
    const int C = p.C, cpg = C / p.groups;
    __shared__ float s_mean[8], s_rstd[8];
    if (threadIdx.x < p.groups) {   // the double-precision moments once per (sample, group), not per thread
        const double count = static_cast<double>(p.P) * cpg;
        group_moments(p.stats, n, p.groups, threadIdx.x, count, p.eps, s_mean[threadIdx.x], s_rstd[threadIdx.x]);
    }
    __syncthreads();
    f32x2 a[4], a2[4], b2[4], ba[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float af[2], bf[2], baf[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = tq * 8 + 2 * i + j;
            const int g = c / cpg;
            af[j] = s_rstd[g] * __ldg(p.gamma + c);
            bf[j] = __ldg(p.beta + c) - s_mean[g] * af[j];
            baf[j] = bf[j] + (p.addvec ? __ldg(p.addvec + static_cast<long long>(n) * p.addvec_stride + c) : 0.f);
        }
        a[i] = pk2(af[0], af[1]);
        a2[i] = pk2(af[0] * kLog2e, af[1] * kLog2e);
        b2[i] = pk2(bf[0] * kLog2e, bf[1] * kLog2e);
        ba[i] = pk2(baf[0], baf[1]);
    }
    const int mrow = n * p.W;      // 32-bit index of the sample's mask row (N * W < 2^31 always)
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(p.P, p_begin + pix_per_block);
    const long long base = static_cast<long long>(n) * p.P * C + tq * 8;
    int xcol = (p_begin + pl) % p.W;   // column of the thread's next pixel, advanced incrementally
    __half2 amax2 = __float2half2_rn(0.f);   // largest |value| stored by this thread (saturation report)
    // element offset of the thread's current pixel, advanced by adds (no 64-bit multiply per 16-byte vector)
    const long long ustride = static_cast<long long>(lanes) * C;
    long long off0 = base + static_cast<long long>(p_begin + pl) * C;
    for (int pix0 = p_begin + pl; pix0 < p_end; pix0 += lanes * U, off0 += ustride * U) {
        uint4 rv[U], rr[U];
        float m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pix = pix0 + u * lanes;
            if (pix < p_end) {
                const long long o = off0 + u * ustride;
                rv[u] = ldg_stream(reinterpret_cast<const uint4*>(p.raw + o));
                if (HAS_RES) rr[u] = ldg_stream(reinterpret_cast<const uint4*>(p.res + o));
                m[u] = __ldg(p.mask + (mrow + xcol));
            }
            xcol += lanes_mod;
            if (xcol >= p.W) xcol -= p.W;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pix = pix0 + u * lanes;
            if (pix < p_end) {
                const uint4 o4 = gn_mish8<HAS_RES>(rv[u], rr[u], a, a2, b2, ba, pk2(m[u], m[u]));
                const __half2* oh = reinterpret_cast<const __half2*>(&o4);
#pragma unroll
                for (int i = 0; i < 4; ++i) amax2 = __hmax2(amax2, __habs2(oh[i]));
                if (!(p.dbg & 2)) *reinterpret_cast<uint4*>(p.out + off0 + u * ustride) = o4;
            }
        }
    }
    if (p.sat && fmaxf(__low2float(amax2), __high2float(amax2)) >= 65504.f) atomicAdd(p.sat, 1ull);
}

int launch_gn_apply(const GnApplyParams& p, int num_sms, cudaStream_t s) {
    const int TP = p.C / 8;
    if (p.C % 8 || TP > 256 || (p.C / p.groups) < 1 || p.C % p.groups) return (int)cudaErrorInvalidValue;
    const int lanes = TP >= 256 ? 1 : 256 / TP;
    const int threads = TP * lanes;
    // 64 pixels per thread for big tensors; shrink (down to 16 per thread: the per-thread prologue loads 24 affine /
    // embedding values) only while the grid would leave SMs idle
    // (small tensors -- a few samples at the low-resolution levels -- are latency-bound: let them spread further)
    static const int env_ppt = getenv("USB_GN_PPT") ? atoi(getenv("USB_GN_PPT")) : 0;
    const int min_ppt = env_ppt > 0 ? env_ppt : ((long long)p.N * p.P * p.C * 2 < (16ll << 20) ? 8 : 32);
    static const int min_bps = getenv("USB_GN_BPS") ? atoi(getenv("USB_GN_BPS")) : 4;
    int ppb = lanes * 64;
    while (ppb > lanes * min_ppt && (long long)((p.P + ppb - 1) / ppb) * p.N < (long long)num_sms * min_bps) ppb >>= 1;
    dim3 grid((p.P + ppb - 1) / ppb, p.N);
    GnApplyParams q = p;
    static const char* dbg_env = getenv("USB_DBG_GN");
    q.dbg = dbg_env ? atoi(dbg_env) : 0;
    if (p.res) return (int)launch_k(gn_apply_kernel<true, 4, 2>, grid, dim3(threads), 0, s, q, ppb, TP, lanes % p.W);
    // ncu (profiles/): with packed fp32x2 math the no-residual form is neither issue- nor DRAM-bound (61 % / 63 %) but
    // waits on its loads; USB_GN_U8=1 trades occupancy (3 -> 2 blocks per SM) for twice the loads in flight per thread
    static const bool u8 = getenv("USB_GN_U8") != nullptr;
    if (u8) return (int)launch_k(gn_apply_kernel<false, 8, 2>, grid, dim3(threads), 0, s, q, ppb, TP, lanes % p.W);
    return (int)launch_k(gn_apply_kernel<false, 4, 3>, grid, dim3(threads), 0, s, q, ppb, TP, lanes % p.W);
}

// =====================================================================================================================
// final block apply -> 1x1 conv -> classifier-free-guidance combine -> posterior update
// =====================================================================================================================
__global__ void __launch_bounds__(256, 2) final_kernel(const FinalParams p, int pix_per_block, int TP) {
    pdl_wait_and_trigger();
    const int b = blockIdx.y;
    const int tq = threadIdx.x % TP;
    const int pl = threadIdx.x / TP;
    const int lanes = blockDim.x / TP;
    const int C = p.C, cpg = C / p.groups;
    f32x2 a[3][4], sh[3][4], wf[4];      // per channel pair (packed fp32x2, see fma2)
    __shared__ float s_mean[3][8], s_rstd[3][8];
    if (threadIdx.x < p.nb * p.groups) {
        const int k = threadIdx.x / p.groups, g = threadIdx.x % p.groups;
        const double count = static_cast<double>(p.P) * cpg;
        group_moments(p.stats, k * p.B + b, p.groups, g, count, p.eps, s_mean[k][g], s_rstd[k][g]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < p.nb) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float af[2], sf[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = tq * 8 + 2 * i + j;
                    af[j] = s_rstd[k][c / cpg] * __ldg(p.gamma + c);
                    sf[j] = __ldg(p.beta + c) - s_mean[k][c / cpg] * af[j];
                }
                a[k][i] = pk2(af[0], af[1]);
                sh[k][i] = pk2(sf[0], sf[1]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) wf[i] = pk2(__ldg(p.wf + tq * 8 + 2 * i), __ldg(p.wf + tq * 8 + 2 * i + 1));
    const float bf = __ldg(p.bf);
    const f32x2 klog = pk2(kLog2e, kLog2e), mhalf = pk2(-0.5f, -0.5f), mone = pk2(-1.f, -1.f);
    float c_x = p.c_x, c_s = p.c_s, sigma = p.sigma;
    const float* noise = p.noise;
    float* outp = p.out;
    if (p.step_ctr) {
        const int i = *p.step_ctr;
        c_x = __ldg(p.step_tab + 4 * i);
        c_s = __ldg(p.step_tab + 4 * i + 1);
        sigma = __ldg(p.step_tab + 4 * i + 2);
        if (__ldg(p.step_tab + 4 * i + 3) == 0.f) outp = nullptr;      // only the last step writes the caller-visible output
        if (noise) noise += static_cast<long long>(i) * p.noise_step_stride;
    }
    const float* mk = p.mask + static_cast<long long>(b) * p.W;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(p.P, p_begin + pix_per_block);
    // every lane of a pixel group must run the same number of iterations (shuffles below); U pixels x nb branches of
    // 16-byte loads are issued before any of them is consumed (the kernel is a pure HBM stream of `raw`)
    constexpr int U = 2;   // with two 256-thread blocks per SM: 2 x 256 x U x nb 16-byte loads in flight per SM
    const int iters = (p_end - p_begin + lanes * U - 1) / (lanes * U);
    for (int it = 0; it < iters; ++it) {
        uint4 rv[U][3];
        float m[U], xt_in[U], nz_in[U];
        int pixs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pixs[u] = p_begin + (it * U + u) * lanes + pl;
            const bool ok = pixs[u] < p_end;
            m[u] = ok ? __ldg(mk + pixs[u] % p.W) : 0.f;
            // the state and noise of the pixel are fetched with the activations, not after the reduction (the lane that
            // applies the update would otherwise stall the whole warp on two dependent global loads per pixel)
            xt_in[u] = nz_in[u] = 0.f;
            if (ok && tq == 0 && p.xt != nullptr) {
                const long long q = static_cast<long long>(b) * p.P + pixs[u];
                xt_in[u] = p.xt[q];
                if (noise) nz_in[u] = noise[q];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                rv[u][k] = make_uint4(0u, 0u, 0u, 0u);
                if (k < p.nb && ok)
                    rv[u][k] = ldg_stream(reinterpret_cast<const uint4*>(
                        p.raw + (static_cast<long long>(k * p.B + b) * p.P + pixs[u]) * C + tq * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pix = pixs[u];
            const bool ok = pix < p_end;
            float sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (k < p.nb) {
                    float dot = 0.f;
                    const __half2* hv = reinterpret_cast<const __half2*>(&rv[u][k]);
                    // Mish(x) = x - 2x/d, d = e^x (e^x + 2) + 2; one reciprocal serves two channels (MUFU is the scarce pipe)
                    f32x2 dot2 = pk2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // packed fp32x2 (see gn_mish8): x = a v + b; dh = -d/2; Mish(x) = x + x * (-2/d)
                        const float2 vf = __half22float2(hv[i]);
                        const f32x2 x = fma2(pk2(vf.x, vf.y), a[k][i], sh[k][i]);
                        float t0, t1;
                        upk2(mul2(x, klog), t0, t1);
                        const f32x2 e = pk2(ex2_ftz(fminf(t0, 20.f * kLog2e)), ex2_ftz(fminf(t1, 20.f * kLog2e)));
                        float d0, d1;
                        upk2(fma2(e, fma2(e, mhalf, mone), mone), d0, d1);
                        const float rn = rcp_ftz(d0 * d1);             // |dh| <= ~1.2e17 each: the product stays finite
                        const f32x2 y = fma2(x, pk2(d1 * rn, d0 * rn), x);
                        dot2 = fma2(wf[i], y, dot2);
                    }
                    {
                        float s0, s1;
                        upk2(dot2, s0, s1);
                        dot = (s0 + s1) * m[u];     // the mask is per pixel: applied once to the 8-channel partial sum
                    }
                    for (int o = TP >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                    sc[k] = (dot + bf) * m[u];
                }
            }
            if (ok && tq == 0) {
                const long long q = static_cast<long long>(b) * p.P + pix;
                if (p.xt == nullptr) {
                    for (int k = 0; k < p.nb; ++k) p.score[(static_cast<long long>(k * p.B + b)) * p.P + pix] = sc[k];
                } else {
                    const float sf = sc[p.nb - 1];
                    float score = sf;
                    if (p.nb >= 2) score = score + p.a0 * (sf - sc[0]);
                    if (p.nb >= 3) score = score + p.a1 * (sf - sc[1]);
                    const float xn = (c_x * xt_in[u] + c_s * score + sigma * nz_in[u]) * m[u];
                    p.xt[q] = xn;
                    if (p.score) p.score[q] = score;
                    if (outp) {
                        float o = xn;
                        if (p.mel_min) {
                            const int yb = pix / p.W;
                            const float lo = __ldg(p.mel_min + yb), hi = __ldg(p.mel_max + yb);
                            // the reference's four separate fp32 ops (no contraction into an fma): bit-identical to the torch expression
                            o = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(xn, 1.f), 2.f), __fsub_rn(hi, lo)), lo);
                        }
                        outp[q] = o;
                    }
                }
            }
        }
    }
}

int launch_final(const FinalParams& p, int num_sms, cudaStream_t s) {
    const int TP = p.C / 8;
    if (p.C % 8 || TP > 32 || (TP & (TP - 1)) || p.nb < 1 || p.nb > 3) return (int)cudaErrorInvalidValue;
    const int lanes = 256 / TP;
    int ppb = lanes * 32;
    while (ppb > lanes * 2 && (long long)((p.P + ppb - 1) / ppb) * p.B < (long long)num_sms * 8) ppb >>= 1;
    dim3 grid((p.P + ppb - 1) / ppb, p.B);
    return (int)launch_k(final_kernel, grid, dim3(256), 0, s, p, ppb, TP);
}

// =====================================================================================================================
// time / speaker embedding
// =====================================================================================================================
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one block per row: sinusoidal embedding -> Linear -> Mish -> Linear -> cat speaker -> Mish.
// p.t == null: the time half of u is 0; p.spk == null: the speaker half is 0 (Mish(0) = 0), which lets the caller
// split the stacked Linear into a per-step part and a per-row part.
__global__ void __launch_bounds__(256) time_mlp_kernel(const EmbedParams p) {
    extern __shared__ float sm[];  // e[dim], h[4*dim], tm[dim]
    float* e = sm;
    float* h = sm + p.dim;
    float* tm = h + 4 * p.dim;
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = p.dim + p.S;
    if (p.t) {
        const int half = p.dim / 2;
        const float ts = p.pe_scale * p.t[n];
        for (int j = threadIdx.x; j < half; j += blockDim.x) {
            const float arg = ts * p.freqs[j];
            e[j] = sinf(arg);
            e[j + half] = cosf(arg);
        }
        __syncthreads();
        // four output rows per warp iteration: their weight loads are in flight together (the kernel is pure latency;
        // dim is a multiple of 32, so 4*dim and dim are multiples of 32).  Per row the summation order is unchanged.
        for (int i0 = warp * 4; i0 < 4 * p.dim; i0 += 32) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = lane; j < p.dim; j += 32) {
                const float ej = e[j];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] += p.w0[static_cast<long long>(i0 + r) * p.dim + j] * ej;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float a = warp_sum_f(acc[r]);
                if (lane == 0) h[i0 + r] = mish_precise(a + p.b0[i0 + r]);
            }
        }
        __syncthreads();
        for (int i0 = warp * 4; i0 < p.dim; i0 += 32) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = lane; j < 4 * p.dim; j += 32) {
                const float hj = h[j];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] += p.w2[static_cast<long long>(i0 + r) * 4 * p.dim + j] * hj;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float a = warp_sum_f(acc[r]);
                if (lane == 0) tm[i0 + r] = a + p.b2[i0 + r];
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        float v;
        if (i < p.dim) v = p.t ? mish_precise(tm[i]) : 0.f;
        else v = p.spk ? mish_precise(p.spk[static_cast<long long>(n) * p.S + (i - p.dim)]) : 0.f;
        p.u[static_cast<long long>(n) * K + i] = v;
    }
}

// warp per output column j of the stacked ResnetBlock.mlp Linears, looping over rows
__global__ void __launch_bounds__(256) emb_linear_kernel(const EmbedParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= p.J) return;
    const int K = p.dim + p.S;
    const float* w = p.wcat + static_cast<long long>(j) * K;
    const float bj = p.bcat ? p.bcat[j] : 0.f;
    for (int n = 0; n < p.N; ++n) {
        const float* u = p.u + static_cast<long long>(n) * K;
        float acc = 0.f;
        for (int i = lane; i < K; i += 32) acc += __ldg(w + i) * u[i];
        acc = warp_sum_f(acc);
        if (lane == 0) p.e[static_cast<long long>(n) * p.J + j] = acc + bj;
    }
}

int launch_embed(const EmbedParams& p, cudaStream_t s) {
    time_mlp_kernel<<<p.N, 256, 6 * p.dim * sizeof(float), s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    emb_linear_kernel<<<(p.J + 7) / 8, 256, 0, s>>>(p);
    return (int)cudaGetLastError();
}

// =====================================================================================================================
// utilities
// =====================================================================================================================
__global__ void downsample_mask_kernel(const float* src, float* dst, int N, int Wsrc, int Wdst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Wdst) return;
    const int n = i / Wdst, w = i - n * Wdst;
    dst[i] = src[static_cast<long long>(n) * Wsrc + 2 * w];
}
int launch_downsample_mask(const float* src, float* dst, int N, int Wsrc, int Wdst, cudaStream_t s) {
    const int total = N * Wdst;
    downsample_mask_kernel<<<(total + 255) / 256, 256, 0, s>>>(src, dst, N, Wsrc, Wdst);
    return (int)cudaGetLastError();
}

// E[n][j] = T[j] + S[n][j]: per-step time part + per-row speaker part of the stacked ResnetBlock.mlp Linears
__global__ void emb_combine_kernel(const float* t_part, const float* s_part, float* e, int N, int J, const int* step_ctr,
                                   unsigned long long* zero, long long zero_count) {
    pdl_wait_and_trigger();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < zero_count) zero[i] = 0ull;      // the GroupNorm statistics slots of the evaluation that follows
    if (i >= static_cast<long long>(N) * J) return;
    if (step_ctr) t_part += static_cast<long long>(*step_ctr) * J;
    e[i] = t_part[i % J] + s_part[i];
}
int launch_emb_combine(const float* t_part, const float* s_part, float* e, int N, int J, const int* step_ctr,
                       unsigned long long* zero, long long zero_count, cudaStream_t s) {
    const long long nj = static_cast<long long>(N) * J;
    const long long total = nj > zero_count ? nj : zero_count;
    return (int)launch_k(emb_combine_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, t_part, s_part, e,
                         N, J, step_ctr, zero, zero_count);
}
__global__ void step_advance_kernel(int* ctr) {
    pdl_wait_and_trigger();
    *ctr += 1;
}
int launch_step_advance(int* ctr, cudaStream_t s) { return (int)launch_k(step_advance_kernel, dim3(1), dim3(1), 0, s, ctr); }

__global__ void gather_rows_kernel(const float* src, const int* idx, float* dst, int N, int len) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(N) * len) return;
    const int n = static_cast<int>(i / len), k = static_cast<int>(i - static_cast<long long>(n) * len);
    dst[i] = src[static_cast<long long>(idx[n]) * len + k];
}
int launch_gather_rows(const float* src, const int* idx, float* dst, int N, int len, cudaStream_t s) {
    const long long total = static_cast<long long>(N) * len;
    gather_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, idx, dst, N, len);
    return (int)cudaGetLastError();
}

// =====================================================================================================================
// training objective, forward value (UnitSpeech.forward_diffusion / loss_t, unitspeech/unitspeech.py:376-405)
// =====================================================================================================================
// xt = (x0 * exp(-cn/2) + z * sqrt(1 - exp(-cn))) * mask, zm = z * mask, mu = cond * mask,
// cn = beta_min * t + 0.5 * (beta_max - beta_min) * t^2 (get_noise, cumulative)
__global__ void forward_diffusion_kernel(const float* __restrict__ x0, const float* __restrict__ z,
                                         const float* __restrict__ cond, const float* __restrict__ mask,
                                         const float* __restrict__ t, float beta_min, float beta_max,
                                         float* __restrict__ xt, float* __restrict__ zm, float* __restrict__ mu, int B,
                                         int F, int T) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * F * T) return;
    const int b = static_cast<int>(i / (static_cast<long long>(F) * T));
    const int w = static_cast<int>(i % T);
    const float tb = __ldg(t + b);
    const float cn = beta_min * tb + 0.5f * (beta_max - beta_min) * (tb * tb);
    const float m = __ldg(mask + static_cast<long long>(b) * T + w);
    const float zz = z[i];
    xt[i] = (x0[i] * expf(-0.5f * cn) + zz * sqrtf(1.0f - expf(-cn))) * m;
    if (zm) zm[i] = zz * m;
    if (mu) mu[i] = cond[i] * m;
}
int launch_forward_diffusion(const float* x0, const float* z, const float* cond, const float* mask, const float* t,
                             float beta_min, float beta_max, float* xt, float* zm, float* mu, int B, int F, int T,
                             cudaStream_t s) {
    const long long total = static_cast<long long>(B) * F * T;
    forward_diffusion_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(x0, z, cond, mask, t, beta_min,
                                                                                       beta_max, xt, zm, mu, B, F, T);
    return (int)cudaGetLastError();
}

// loss = sum((score * sqrt(1 - exp(-cn)) + zm)^2) / (sum(mask) * F): fixed-shape two-stage reduction in double
// (block partials in a fixed order, then one block), so the value does not depend on scheduling.
constexpr int kLossBlocks = 256;
__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ score, const float* __restrict__ zm,
                                                           const float* __restrict__ mask, const float* __restrict__ t,
                                                           float beta_min, float beta_max, double* __restrict__ partial,
                                                           int B, int F, int T) {
    __shared__ double red[2][256];
    const long long total = static_cast<long long>(B) * F * T;
    const long long mtotal = static_cast<long long>(B) * T;
    double acc = 0.0, macc = 0.0;
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < total; i += 256LL * kLossBlocks) {
        const int b = static_cast<int>(i / (static_cast<long long>(F) * T));
        const float tb = __ldg(t + b);
        const float cn = beta_min * tb + 0.5f * (beta_max - beta_min) * (tb * tb);
        const float v = score[i] * sqrtf(1.0f - expf(-cn)) + zm[i];
        acc += static_cast<double>(v * v);
    }
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < mtotal; i += 256LL * kLossBlocks)
        macc += static_cast<double>(mask[i]);
    red[0][threadIdx.x] = acc;
    red[1][threadIdx.x] = macc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = red[0][0];
        partial[2 * blockIdx.x + 1] = red[1][0];
    }
}
__global__ void __launch_bounds__(256) loss_final_kernel(const double* __restrict__ partial, float* __restrict__ loss, int F) {
    __shared__ double red[2][256];
    red[0][threadIdx.x] = partial[2 * threadIdx.x];
    red[1][threadIdx.x] = partial[2 * threadIdx.x + 1];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            red[0][threadIdx.x] += red[0][threadIdx.x + s];
            red[1][threadIdx.x] += red[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = static_cast<float>(red[0][0] / (red[1][0] * F));
}
int launch_diffusion_loss(const float* score, const float* zm, const float* mask, const float* t, float beta_min,
                          float beta_max, double* partial, float* loss, int B, int F, int T, cudaStream_t s) {
    static_assert(kLossBlocks == 256, "loss_final_kernel reduces exactly 256 partials");
    loss_partial_kernel<<<kLossBlocks, 256, 0, s>>>(score, zm, mask, t, beta_min, beta_max, partial, B, F, T);
    loss_final_kernel<<<1, 256, 0, s>>>(partial, loss, F);
    return (int)cudaGetLastError();
}

}  // namespace usb
