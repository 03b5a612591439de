// LinearAttention context (unitspeech/unitspeech.py:86-96), restructured as a streaming reduction:
//   ctx[h][d][e] = sum_p softmax_p(k[h][d][:])[p] * v[h][e][p]          (softmax over ALL positions, no mask)
//   attn(x)[co][p] = sum_{h,d} Weff[co][h*32+d] * q[h][d][p] + b_o[co],  Weff[co][h*32+d] = sum_e Wo[co][h*32+e] ctx[h][d][e]
// Pass 1 (attn_partial_kernel): each block reduces one (sample, head, chunk of positions) with a chunk-local max
//   (online-softmax partial: m[d], s[d] = sum exp(k-m), c[d][e] = sum exp(k-m) v).  The 32x32 outer-product sums run
//   on the warp-level tensor cores (mma.sync m16n8k16, fp16 operands = the stored fp16 v and the fp16-rounded exp,
//   fp32 accumulate); the kernel is bound by reading k and v once from HBM.
// Pass 2 (attn_merge_kernel): merges the chunk partials exactly (rescale by exp(m_chunk - m_global)) and normalises.
// Pass 3 (attn_fold_kernel): folds to_out into the per-sample fp16 weight consumed by the tcgen05 GEMM kernel.
// All reductions run in a fixed order: results are bit-reproducible.
#include "kernels.h"
#include "pdl.h"

namespace usb {

namespace {
__device__ __forceinline__ float exp2f_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kDh = 32;                             // dim_head
constexpr int kPartStride = kDh * kDh + 2 * kDh;    // ctx + m + s
constexpr int kFoldRows = 32;                       // output channels per fold block (8 when the launch would have few blocks)
constexpr int kCtxLd = kDh + 4;                     // 36: float4-aligned, conflict-free row stride
constexpr int kRowHalfs = 40;                       // 32 halfs + 8 pad: 80-byte rows make ldmatrix conflict-free

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
}  // namespace

int attn_chunks(int P, int chunk) { return (P + chunk - 1) / chunk; }

// grid (chunks, N), one warp per head (<= 8 heads).  Each warp streams its head's 64-byte k and v slices with
// cp.async into a double-buffered 32-position staging tile (the next tile is in flight while the current one is
// reduced), keeps a running max per channel (online softmax) and accumulates ctx on mma.sync.
constexpr int kTileP = 32;   // positions per warp tile

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256) attn_partial_kernel(const AttnParams p, int nchunks) {
    // per warp: Kraw[2][32][40], V[2][32][40], P[32][40] halfs = 5 x 2560 B
    extern __shared__ __align__(16) unsigned char sbuf[];
    __shared__ float wscale_s[8][kDh];
    pdl_wait_and_trigger();
    const int chunk_id = blockIdx.x, n = blockIdx.y;
    const int ld = p.ld;
    const int p0 = chunk_id * p.chunk;
    const int p1 = min(p.P, p0 + p.chunk);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int head = warp;
    const __half* base = p.qkv + static_cast<long long>(n) * p.P * ld;
    const int koff = p.koff + head * kDh;
    const int voff = p.voff + head * kDh;
    constexpr int kTileBytes = kTileP * kRowHalfs * 2;   // 2560
    const uint32_t wbase = static_cast<uint32_t>(__cvta_generic_to_shared(sbuf)) + warp * 5 * kTileBytes;
    const uint32_t kraw0 = wbase, vt0 = wbase + 2 * kTileBytes, pt = wbase + 4 * kTileBytes;
    unsigned char* wptr = sbuf + warp * 5 * kTileBytes;

    const int oct = lane & 3;          // this lane handles channels oct*8 .. oct*8+7 of positions (lane>>2) + 8*it
    const int g = lane >> 2;
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    float mrun[8], ssum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mrun[i] = -INFINITY;
        ssum[i] = 0.f;
    }
    const int lm_r = lane & 7, lm_j = lane >> 3;

    auto issue = [&](int t0, int buf) {
#pragma unroll
        for (int it = 0; it < kTileP / 8; ++it) {
            const int pp = g + it * 8;
            const int pos = t0 + pp;
            const bool ok = pos < p1;
            const __half* row = base + static_cast<long long>(ok ? pos : p0) * ld;
            const uint32_t off = buf * kTileBytes + (pp * kRowHalfs + oct * 8) * 2;
            cp_async16(kraw0 + off, row + koff + oct * 8, ok);
            cp_async16(vt0 + off, row + voff + oct * 8, ok);
        }
        cp_async_commit();
    };

    int buf = 0;
    issue(p0, 0);
    for (int t0 = p0; t0 < p1; t0 += kTileP, buf ^= 1) {
        if (t0 + kTileP < p1) {
            issue(t0 + kTileP, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        // tile max per channel: own 4 positions, then the 8 lanes sharing the channel octet
        const unsigned char* kt = wptr + buf * kTileBytes;
        uint4 kraw[kTileP / 8];
        float tmax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tmax[i] = -INFINITY;
#pragma unroll
        for (int it = 0; it < kTileP / 8; ++it) {
            const int pp = g + it * 8;
            kraw[it] = *reinterpret_cast<const uint4*>(kt + (pp * kRowHalfs + oct * 8) * 2);
            if (t0 + pp < p1) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&kraw[it]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    tmax[2 * i] = fmaxf(tmax[2 * i], f.x);
                    tmax[2 * i + 1] = fmaxf(tmax[2 * i + 1], f.y);
                }
            }
        }
        float scale[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float m = tmax[i];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            const float mnew = fmaxf(mrun[i], m * kLog2e);
            scale[i] = exp2f_ftz(mrun[i] - mnew);       // 0 on the first tile (mrun = -inf), 1 when unchanged
            mrun[i] = mnew;
            ssum[i] *= scale[i];
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) wscale_s[warp][lane * 8 + i] = scale[i];
        }
        // P = exp2(k*log2e - mrun) in fp16 (rows beyond the chunk are zero)
#pragma unroll
        for (int it = 0; it < kTileP / 8; ++it) {
            const int pp = g + it * 8;
            uint4 pk = make_uint4(0, 0, 0, 0);
            if (t0 + pp < p1) {
                const __half2* h2 = reinterpret_cast<const __half2*>(&kraw[it]);
                __half2 o2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    o2[i] = __floats2half2_rn(exp2f_ftz(fmaf(f.x, kLog2e, -mrun[2 * i])),
                                              exp2f_ftz(fmaf(f.y, kLog2e, -mrun[2 * i + 1])));
                    const float2 r = __half22float2(o2[i]);   // the normaliser sums the values the MMA sees
                    ssum[2 * i] += r.x;
                    ssum[2 * i + 1] += r.y;
                }
                pk = *reinterpret_cast<const uint4*>(o2);
            }
            *reinterpret_cast<uint4*>(wptr + 4 * kTileBytes + (pp * kRowHalfs + oct * 8) * 2) = pk;
        }
        __syncwarp();
        // rescale the accumulator rows this thread owns: d = mt*16 + g and + 8
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const float sc0 = wscale_s[warp][mt * 16 + g], sc1 = wscale_s[warp][mt * 16 + g + 8];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                acc[mt][nt][0] *= sc0; acc[mt][nt][1] *= sc0;
                acc[mt][nt][2] *= sc1; acc[mt][nt][3] *= sc1;
            }
        }
        const uint32_t vs_base = vt0 + buf * kTileBytes;
#pragma unroll
        for (int ks = 0; ks < kTileP / 16; ++ks) {
            const int pk0 = ks * 16;
            uint32_t a[2][4], b[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int prow = pk0 + (lm_j >> 1) * 8 + lm_r, dcol = mt * 16 + (lm_j & 1) * 8;
                ldmatrix_x4_trans(pt + (prow * kRowHalfs + dcol) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
            }
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                const int prow = pk0 + (lm_j & 1) * 8 + lm_r, ecol = np * 16 + (lm_j >> 1) * 8;
                ldmatrix_x4_trans(vs_base + (prow * kRowHalfs + ecol) * 2, b[2 * np][0], b[2 * np][1], b[2 * np + 1][0],
                                  b[2 * np + 1][1]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_16816(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b[nt][0], b[nt][1]);
        }
        __syncwarp();   // the tile buffers are rewritten by the next iterations' cp.async / P stores
    }

    // ---- each warp owns one head: write its partial (ctx, max in the natural-log domain, normaliser)
    float* out = p.part + ((static_cast<long long>(n) * p.heads + head) * nchunks + chunk_id) * kPartStride;
    {
        const int t = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                *reinterpret_cast<float2*>(out + (mt * 16 + g) * kDh + nt * 8 + 2 * t) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(out + (mt * 16 + g + 8) * kDh + nt * 8 + 2 * t) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 4);
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 8);
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 16);
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                out[kDh * kDh + lane * 8 + i] = mrun[i] * 0.6931471805599453f;
                out[kDh * kDh + kDh + lane * 8 + i] = ssum[i];
            }
        }
    }
}

// grid (heads, N), 256 threads: ctx[n][h][d][e] = sum_c part_c[d][e] exp(m_c[d] - M[d]) / sum_c s_c[d] exp(m_c[d] - M[d])
__global__ void __launch_bounds__(256) attn_merge_kernel(const AttnParams p, int nchunks, float* ctx_out) {
    __shared__ float M_s[kDh], S_s[kDh];
    pdl_wait_and_trigger();
    const int h = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const float* ph = p.part + (static_cast<long long>(n) * p.heads + h) * nchunks * kPartStride;
    if (tid < kDh) {
        float M = -INFINITY;
#pragma unroll 8
        for (int c = 0; c < nchunks; ++c) M = fmaxf(M, ph[c * kPartStride + kDh * kDh + tid]);
        float S = 0.f;
#pragma unroll 8
        for (int c = 0; c < nchunks; ++c)
            S += ph[c * kPartStride + kDh * kDh + kDh + tid] *
                 exp2f_ftz((ph[c * kPartStride + kDh * kDh + tid] - M) * kLog2e);
        M_s[tid] = M;
        S_s[tid] = S;
        if (p.stat_out) {
            float* so = p.stat_out + (static_cast<long long>(n) * p.heads + h) * 2 * kDh;
            so[tid] = M;
            so[kDh + tid] = S;
        }
    }
    __syncthreads();
    float* o = ctx_out + (static_cast<long long>(n) * p.heads + h) * kDh * kDh;
    for (int i = tid; i < kDh * kDh; i += 256) {
        const int d = i >> 5;
        const float M = M_s[d];
        float a = 0.f;
#pragma unroll 8
        for (int c = 0; c < nchunks; ++c)
            a += ph[c * kPartStride + i] * exp2f_ftz((ph[c * kPartStride + kDh * kDh + d] - M) * kLog2e);
        o[i] = a / S_s[d];
    }
}

// grid (ceil(C/32), N): Weff[n][co][h*32+d] = sum_e Wo[co][h*32+e] * ctx[n][h][d][e] for 32 output channels.
// Plain mode writes Weff (fp16).  Fused-q mode goes on to W'[n][co][c] = g * sum_d' Weff[co][d'] Wq[d'][c] + (co == c),
// the per-sample 1x1 weight of the whole Residual(Rezero(LinearAttention)) block, and b' = g * b_o.
template <int kFoldRows>
__global__ void __launch_bounds__(256) attn_fold_kernel(const AttnParams p, const float* ctx_in) {
    extern __shared__ __align__(16) float sm[];   // ctx [heads*32][36], weff tile [32][hidden + 4], Wo tile [32][hidden]
    pdl_wait_and_trigger();
    const int n = blockIdx.y, tid = threadIdx.x;
    const int heads = p.heads, hidden = heads * kDh;
    const int wld = hidden + 4;
    float* ctx = sm;
    float* wt = sm + heads * kDh * kCtxLd;
    float* wos = wt + kFoldRows * wld;
    const float* cn = ctx_in + static_cast<long long>(n) * heads * kDh * kDh;
    for (int i = tid; i < heads * kDh * kDh; i += 256) {
        const int hd = i >> 5, e = i & 31;
        ctx[hd * kCtxLd + e] = cn[i];
    }
    const int co0 = blockIdx.x * kFoldRows;
    for (int i = tid; i < kFoldRows * hidden; i += 256) {
        const int co = co0 + i / hidden;
        wos[i] = co < p.C ? __ldg(p.wo + static_cast<long long>(co) * hidden + i % hidden) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < kFoldRows * hidden; i += 256) {
        const int r = i / hidden, co = co0 + r;
        const int k = i % hidden, h = k / kDh, d = k % kDh;
        const float4* w = reinterpret_cast<const float4*>(wos + r * hidden + h * kDh);
        const float4* c = reinterpret_cast<const float4*>(ctx + (h * kDh + d) * kCtxLd);
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < kDh / 4; ++e) {
            const float4 wv = w[e], cv = c[e];
            a = fmaf(wv.x, cv.x, a); a = fmaf(wv.y, cv.y, a); a = fmaf(wv.z, cv.z, a); a = fmaf(wv.w, cv.w, a);
        }
        if (co < p.C && !p.wq) p.weff[(static_cast<long long>(n) * p.C + co) * hidden + k] = __float2half_rn(a);
        wt[r * wld + k] = a;
    }
    if (!p.wq) return;
    __syncthreads();
    // W'[co0 + r][c] for the block's 32 rows: thread = column c, rpt rows each, Wq streamed once per thread
    const int groups = 256 / p.C;                 // C in {64, 128, 256}
    const int rpt = kFoldRows / groups;           // rows per thread: kFoldRows / 4, / 2 or / 1
    const int c = tid % p.C, r0 = (tid / p.C) * rpt;
    const float g = __ldg(p.g);
    float acc[kFoldRows];
#pragma unroll
    for (int r = 0; r < kFoldRows; ++r) acc[r] = 0.f;
    for (int d0 = 0; d0 < hidden; d0 += 8) {
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = __ldg(p.wq + static_cast<long long>(d0 + i) * p.C + c);
#pragma unroll
        for (int r = 0; r < kFoldRows; ++r) {
            if (r < rpt) {
                const float4* wr = reinterpret_cast<const float4*>(wt + (r0 + r) * wld + d0);
                const float4 a0 = wr[0], a1 = wr[1];
                acc[r] = fmaf(a0.x, w[0], acc[r]); acc[r] = fmaf(a0.y, w[1], acc[r]);
                acc[r] = fmaf(a0.z, w[2], acc[r]); acc[r] = fmaf(a0.w, w[3], acc[r]);
                acc[r] = fmaf(a1.x, w[4], acc[r]); acc[r] = fmaf(a1.y, w[5], acc[r]);
                acc[r] = fmaf(a1.z, w[6], acc[r]); acc[r] = fmaf(a1.w, w[7], acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kFoldRows; ++r) {
        if (r < rpt) {
            const int co = co0 + r0 + r;
            if (co < p.C)
                p.weff[(static_cast<long long>(n) * p.C + co) * p.C + c] =
                    __float2half_rn(acc[r] * g + (co == c ? 1.f : 0.f));
        }
    }
    if (n == 0)
        for (int r = tid; r < kFoldRows && co0 + r < p.C; r += 256) p.bprime[co0 + r] = g * __ldg(p.bo + co0 + r);
}

int launch_attn_context(const AttnParams& p, cudaStream_t s) {
    if (p.heads > 8 || p.heads < 1) return (int)cudaErrorInvalidValue;
    const int nchunks = attn_chunks(p.P, p.chunk);
    static bool part_attr = false;
    if (!part_attr) {   // heads x 12.5 KB of staging
        cudaError_t ea = cudaFuncSetAttribute(attn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (ea != cudaSuccess) return (int)ea;
        part_attr = true;
    }
    dim3 g1(nchunks, p.N);
    cudaError_t e = launch_k(attn_partial_kernel, g1, dim3(32 * p.heads), p.heads * 5 * kTileP * kRowHalfs * 2, s, p, nchunks);
    if (e != cudaSuccess) return (int)e;
    // merged context lives behind the partials in the same scratch buffer
    float* ctx = p.ctx_out ? p.ctx_out : p.part + static_cast<long long>(p.N) * p.heads * nchunks * kPartStride;
    dim3 g2(p.heads, p.N);
    e = launch_k(attn_merge_kernel, g2, dim3(256), 0, s, p, nchunks, ctx);
    if (e != cudaSuccess) return (int)e;
    static bool fold_attr = false;
    if (!fold_attr) {   // ctx + Weff tile + Wo tile = 48.6 KB of dynamic shared memory
        e = cudaFuncSetAttribute(attn_fold_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return (int)e;
        fold_attr = true;
    }
    // few samples: 8 output channels per block instead of 32, four times the blocks (a function of the launch shape only;
    // every output element is computed by the same instruction sequence either way, so results do not change)
    if (static_cast<long long>((p.C + 31) / 32) * p.N < 2 * 148) {
        dim3 g3((p.C + 7) / 8, p.N);
        return (int)launch_k(attn_fold_kernel<8>, g3, dim3(256), (p.heads * kDh * kCtxLd + 8 * (2 * p.heads * kDh + 4)) * sizeof(float),
                             s, p, static_cast<const float*>(ctx));
    } else {
        dim3 g3((p.C + 31) / 32, p.N);
        return (int)launch_k(attn_fold_kernel<32>, g3, dim3(256), (p.heads * kDh * kCtxLd + 32 * (2 * p.heads * kDh + 4)) * sizeof(float),
                             s, p, static_cast<const float*>(ctx));
    }
}

size_t attn_scratch_bytes(int N, int heads, int P, int chunk) {
    return (static_cast<size_t>(N) * heads * attn_chunks(P, chunk) * kPartStride + static_cast<size_t>(N) * heads * kDh * kDh) *
           sizeof(float);
}

}  // namespace usb
