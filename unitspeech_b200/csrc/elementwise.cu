// Streaming kernels of the score estimator: input conv, fused GroupNorm-apply + Mish + embedding/residual + mask,
// the fused "final block -> 1x1 -> guidance combine -> posterior update" step, and the time/speaker embedding.
// All are HBM-bound: 16-byte vector accesses, one channel octet per thread, consecutive threads on consecutive
// addresses, grids sized in multiples of the SM count.
#include "kernels.h"

namespace usb {

namespace {

// x * tanh(softplus(x)) = x * w / (w + 2) with w = e^x (e^x + 2); exact for x > 20 (ratio rounds to 1), which is
// the reference's softplus threshold (unitspeech.py:13-15).
__device__ __forceinline__ float mish_fast(float x) {
    const float e = __expf(fminf(x, 20.f));
    const float w = e * (e + 2.f);
    return x * __fdividef(w, w + 2.f);
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
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
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
__global__ void __launch_bounds__(256) first_conv_kernel(const FirstConvParams p, int pix_per_block) {
    extern __shared__ float sw[];  // [18][C] 3x3 weights, then [2][C] 1x1 weights
    __shared__ unsigned long long gsum[16];
    const int C = p.C;
    const int n = blockIdx.y;
    const int TP = C >> 3;
    const int tq = threadIdx.x % TP;
    const int pl = threadIdx.x / TP;
    const int lanes = blockDim.x / TP;
    for (int i = threadIdx.x; i < 18 * C; i += blockDim.x) sw[i] = p.w3[i];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[18 * C + i] = p.w1[i];
    if (threadIdx.x < 16) gsum[threadIdx.x] = 0ull;
    __syncthreads();

    float b3[8], b1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        b3[i] = p.b3[tq * 8 + i];
        b1[i] = p.b1[tq * 8 + i];
    }
    const int xr = p.x_row[n];
    const int mr = p.mu_row[n];
    const int H = p.H, W = p.W, P = H * W;
    const float* xs = p.x + static_cast<long long>(xr) * P;
    const float* ms = mr >= 0 ? p.cond + static_cast<long long>(mr) * P : nullptr;
    const float* mk = p.mask + static_cast<long long>(n) * W;

    float s = 0.f, ss = 0.f;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(P, p_begin + pix_per_block);
    for (int pix = p_begin + pl; pix < p_end; pix += lanes) {
        const int y = pix / W, x = pix - y * W;
        float in[18];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = y + dy - 1, xx = x + dx - 1;
                float vm = 0.f, vx = 0.f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const float m = __ldg(mk + xx);
                    vm = (ms ? __ldg(ms + yy * W + xx) : __ldg(p.text_uncon + yy)) * m;
                    vx = __ldg(xs + yy * W + xx) * m;
                }
                in[(dy * 3 + dx) * 2 + 0] = vm;  // channel 0 = mu, channel 1 = x (torch.stack([mu, x], 1))
                in[(dy * 3 + dx) * 2 + 1] = vx;
            }
        }
        float acc[8], rr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] = b3[i];
            rr[i] = b1[i];
        }
#pragma unroll
        for (int k = 0; k < 18; ++k) {
            const float4 w0 = *reinterpret_cast<const float4*>(sw + k * C + tq * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + k * C + tq * 8 + 4);
            acc[0] += in[k] * w0.x; acc[1] += in[k] * w0.y; acc[2] += in[k] * w0.z; acc[3] += in[k] * w0.w;
            acc[4] += in[k] * w1.x; acc[5] += in[k] * w1.y; acc[6] += in[k] * w1.z; acc[7] += in[k] * w1.w;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float v = in[8 + k];  // centre tap
            const float4 w0 = *reinterpret_cast<const float4*>(sw + (18 + k) * C + tq * 8);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + (18 + k) * C + tq * 8 + 4);
            rr[0] += v * w0.x; rr[1] += v * w0.y; rr[2] += v * w0.z; rr[3] += v * w0.w;
            rr[4] += v * w1.x; rr[5] += v * w1.y; rr[6] += v * w1.z; rr[7] += v * w1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += acc[i];
            ss += acc[i] * acc[i];
        }
        const long long o = (static_cast<long long>(n) * P + pix) * C + tq * 8;
        *reinterpret_cast<uint4*>(p.raw + o) = pack8(acc);
        *reinterpret_cast<uint4*>(p.res + o) = pack8(rr);
    }
    const int cpg = C / p.groups;
    const int g = (tq * 8) / cpg;
    atomicAdd(&gsum[g * 2], static_cast<unsigned long long>(__float2ll_rn(s * kStatSumScale)));
    atomicAdd(&gsum[g * 2 + 1], static_cast<unsigned long long>(__float2ll_rn(ss * kStatSqScale)));
    __syncthreads();
    if (threadIdx.x < p.groups * 2)
        atomicAdd(reinterpret_cast<unsigned long long*>(p.stats) + static_cast<long long>(n) * p.groups * 2 + threadIdx.x,
                  gsum[threadIdx.x]);
}

int launch_first_conv(const FirstConvParams& p, cudaStream_t s) {
    const int TP = p.C / 8;
    if (p.C % 8 || TP > 256 || 256 % TP || p.groups > 8 || (p.C / p.groups) % 8) return (int)cudaErrorInvalidValue;
    const int P = p.H * p.W;
    int ppb = 1024;
    while (ppb > 64 && (long long)((P + ppb - 1) / ppb) * p.N < 148 * 4) ppb >>= 1;
    dim3 grid((P + ppb - 1) / ppb, p.N);
    first_conv_kernel<<<grid, 256, 20 * p.C * sizeof(float), s>>>(p, ppb);
    return (int)cudaGetLastError();
}

// =====================================================================================================================
// GroupNorm apply + Mish (+ embedding vector) (+ residual) * mask
// =====================================================================================================================
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApplyParams p, int pix_per_block, int TP) {
    const int n = blockIdx.y;
    const int tq = threadIdx.x % TP;
    const int pl = threadIdx.x / TP;
    const int lanes = blockDim.x / TP;
    const int C = p.C, cpg = C / p.groups;
    float a[8], b[8], add[8];
    {
        const double count = static_cast<double>(p.P) * cpg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = tq * 8 + i;
            float mean, rstd;
            group_moments(p.stats, n, p.groups, c / cpg, count, p.eps, mean, rstd);
            a[i] = rstd * __ldg(p.gamma + c);
            b[i] = __ldg(p.beta + c) - mean * a[i];
            add[i] = p.addvec ? __ldg(p.addvec + static_cast<long long>(n) * p.addvec_stride + c) : 0.f;
        }
    }
    const float* mk = p.mask + static_cast<long long>(n) * p.W;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(p.P, p_begin + pix_per_block);
    const long long base = static_cast<long long>(n) * p.P * C + tq * 8;
    for (int pix = p_begin + pl; pix < p_end; pix += lanes) {
        const long long o = base + static_cast<long long>(pix) * C;
        const float m = __ldg(mk + pix % p.W);
        float v[8];
        unpack8(ldg_stream(reinterpret_cast<const uint4*>(p.raw + o)), v);
        float r[8];
        if (p.res) unpack8(ldg_stream(reinterpret_cast<const uint4*>(p.res + o)), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float y = mish_fast(v[i] * a[i] + b[i]) + add[i];
            if (p.res) y += r[i];
            v[i] = y * m;
        }
        *reinterpret_cast<uint4*>(p.out + o) = pack8(v);
    }
}

int launch_gn_apply(const GnApplyParams& p, int num_sms, cudaStream_t s) {
    const int TP = p.C / 8;
    if (p.C % 8 || TP > 256 || (p.C / p.groups) < 1 || p.C % p.groups) return (int)cudaErrorInvalidValue;
    const int lanes = TP >= 256 ? 1 : 256 / TP;
    const int threads = TP * lanes;
    // aim for ~8 blocks per SM over the whole tensor
    int ppb = lanes * 64;
    while (ppb > lanes * 4 && (long long)((p.P + ppb - 1) / ppb) * p.N < (long long)num_sms * 8) ppb >>= 1;
    dim3 grid((p.P + ppb - 1) / ppb, p.N);
    gn_apply_kernel<<<grid, threads, 0, s>>>(p, ppb, TP);
    return (int)cudaGetLastError();
}

// =====================================================================================================================
// final block apply -> 1x1 conv -> classifier-free-guidance combine -> posterior update
// =====================================================================================================================
__global__ void __launch_bounds__(256) final_kernel(const FinalParams p, int pix_per_block, int TP) {
    const int b = blockIdx.y;
    const int tq = threadIdx.x % TP;
    const int pl = threadIdx.x / TP;
    const int lanes = blockDim.x / TP;
    const int C = p.C, cpg = C / p.groups;
    float a[3][8], sh[3][8], wf[8];
    const double count = static_cast<double>(p.P) * cpg;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (k < p.nb) {
            const int n = k * p.B + b;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = tq * 8 + i;
                float mean, rstd;
                group_moments(p.stats, n, p.groups, c / cpg, count, p.eps, mean, rstd);
                a[k][i] = rstd * __ldg(p.gamma + c);
                sh[k][i] = __ldg(p.beta + c) - mean * a[k][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) wf[i] = __ldg(p.wf + tq * 8 + i);
    const float bf = __ldg(p.bf);
    const float* mk = p.mask + static_cast<long long>(b) * p.W;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(p.P, p_begin + pix_per_block);
    // every lane of a pixel group must run the same number of iterations (shuffles below)
    const int iters = (p_end - p_begin + lanes - 1) / lanes;
    for (int it = 0; it < iters; ++it) {
        const int pix = p_begin + it * lanes + pl;
        const bool ok = pix < p_end;
        const float m = ok ? __ldg(mk + pix % p.W) : 0.f;
        float sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < p.nb) {
                float dot = 0.f;
                if (ok) {
                    const long long o = (static_cast<long long>(k * p.B + b) * p.P + pix) * C + tq * 8;
                    float v[8];
                    unpack8(ldg_stream(reinterpret_cast<const uint4*>(p.raw + o)), v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dot += wf[i] * (mish_fast(v[i] * a[k][i] + sh[k][i]) * m);
                }
                for (int o = TP >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                sc[k] = (dot + bf) * m;
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
                const float nz = p.noise ? p.noise[q] : 0.f;
                p.xt[q] = (p.c_x * p.xt[q] + p.c_s * score + p.sigma * nz) * m;
                if (p.score) p.score[q] = score;
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
    final_kernel<<<grid, 256, 0, s>>>(p, ppb, TP);
    return (int)cudaGetLastError();
}

// =====================================================================================================================
// time / speaker embedding
// =====================================================================================================================
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one block per row: sinusoidal embedding -> Linear -> Mish -> Linear -> cat speaker -> Mish
__global__ void __launch_bounds__(256) time_mlp_kernel(const EmbedParams p) {
    extern __shared__ float sm[];  // e[dim], h[4*dim], tm[dim]
    float* e = sm;
    float* h = sm + p.dim;
    float* tm = h + 4 * p.dim;
    const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = p.dim / 2;
    const float ts = p.pe_scale * p.t[n];
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float arg = ts * p.freqs[j];
        e[j] = sinf(arg);
        e[j + half] = cosf(arg);
    }
    __syncthreads();
    for (int i = warp; i < 4 * p.dim; i += 8) {
        float acc = 0.f;
        for (int j = lane; j < p.dim; j += 32) acc += p.w0[static_cast<long long>(i) * p.dim + j] * e[j];
        acc = warp_sum_f(acc);
        if (lane == 0) h[i] = mish_precise(acc + p.b0[i]);
    }
    __syncthreads();
    for (int i = warp; i < p.dim; i += 8) {
        float acc = 0.f;
        for (int j = lane; j < 4 * p.dim; j += 32) acc += p.w2[static_cast<long long>(i) * 4 * p.dim + j] * h[j];
        acc = warp_sum_f(acc);
        if (lane == 0) tm[i] = acc + p.b2[i];
    }
    __syncthreads();
    const int K = p.dim + p.S;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const float v = i < p.dim ? tm[i] : p.spk[static_cast<long long>(n) * p.S + (i - p.dim)];
        p.u[static_cast<long long>(n) * K + i] = mish_precise(v);
    }
}

// warp per output column j of the stacked ResnetBlock.mlp Linears, looping over rows
__global__ void __launch_bounds__(256) emb_linear_kernel(const EmbedParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * 8 + warp;
    if (j >= p.J) return;
    const int K = p.dim + p.S;
    const float* w = p.wcat + static_cast<long long>(j) * K;
    const float bj = p.bcat[j];
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

}  // namespace usb
