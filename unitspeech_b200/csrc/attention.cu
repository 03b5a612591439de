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
constexpr int kWarpTile = 64;                       // positions staged per warp iteration
constexpr int kFoldRows = 32;                       // output channels per fold block
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

// grid (chunks, heads, N), 128 threads = 4 warps; warp w takes positions p0 + w*64 + 256*i
__global__ void __launch_bounds__(128) attn_partial_kernel(const AttnParams p, int nchunks) {
    // per-warp staging tiles Ps (exp(k - m)) and Vs; the same memory holds the cross-warp reduction afterwards
    __shared__ __align__(16) unsigned char sbuf[2 * 4 * kWarpTile * kRowHalfs * 2];
    __half (*Ps)[kWarpTile][kRowHalfs] = reinterpret_cast<__half (*)[kWarpTile][kRowHalfs]>(sbuf);
    __half (*Vs)[kWarpTile][kRowHalfs] =
        reinterpret_cast<__half (*)[kWarpTile][kRowHalfs]>(sbuf + 4 * kWarpTile * kRowHalfs * 2);
    float (*csum)[kDh][kDh + 1] = reinterpret_cast<float (*)[kDh][kDh + 1]>(sbuf);   // [4][32][33] after the loop
    __shared__ float red[4][kDh];
    __shared__ float mmax[kDh];

    const int chunk_id = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
    const int ld = p.ld;
    const int p0 = chunk_id * p.chunk;
    const int p1 = min(p.P, p0 + p.chunk);
    const __half* base = p.qkv + static_cast<long long>(n) * p.P * ld;
    const int koff = p.koff + head * kDh;
    const int voff = p.voff + head * kDh;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- chunk-local max of k per d: thread = (position slice, 8 channels)
    {
        const int oct = tid & 3, sl = tid >> 2;   // 32 position slices
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
        for (int pos = p0 + sl; pos < p1; pos += 32) {
            const uint4 kv = *reinterpret_cast<const uint4*>(base + static_cast<long long>(pos) * ld + koff + oct * 8);
            const __half2* h2 = reinterpret_cast<const __half2*>(&kv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(h2[i]);
                m[2 * i] = fmaxf(m[2 * i], f.x);
                m[2 * i + 1] = fmaxf(m[2 * i + 1], f.y);
            }
        }
        // lanes with equal (lane & 3) hold the same channels: xor 4, 8, 16 within the warp, then across the 4 warps
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 4));
            m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 8));
            m[i] = fmaxf(m[i], __shfl_xor_sync(0xffffffffu, m[i], 16));
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = m[i];
        }
        __syncthreads();
        if (tid < kDh) mmax[tid] = fmaxf(fmaxf(red[0][tid], red[1][tid]), fmaxf(red[2][tid], red[3][tid]));
        __syncthreads();
    }

    // ---- accumulate ctx[d][e] on tensor cores: A = P^T (d x pos), B = V (pos x e)
    float acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
    const int oct = lane & 3;          // this lane stages channels oct*8 .. oct*8+7 of every position it touches
    float mloc[8], ssum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mloc[i] = mmax[oct * 8 + i] * kLog2e;
        ssum[i] = 0.f;
    }
    const uint32_t ps_base = static_cast<uint32_t>(__cvta_generic_to_shared(&Ps[warp][0][0]));
    const uint32_t vs_base = static_cast<uint32_t>(__cvta_generic_to_shared(&Vs[warp][0][0]));
    // ldmatrix source rows: lane supplies row (lane & 7) of 8x8 matrix (lane >> 3)
    const int lm_r = lane & 7, lm_j = lane >> 3;

    for (int t0 = p0 + warp * kWarpTile; t0 < p1; t0 += 4 * kWarpTile) {
        // stage 64 positions: item = lane + 32*it -> position item/4, channel octet item%4 (= lane & 3)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int pp = (lane >> 2) + it * 8;
            const int pos = t0 + pp;
            uint4 pk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
            if (pos < p1) {
                const __half* row = base + static_cast<long long>(pos) * ld;
                const uint4 kv = *reinterpret_cast<const uint4*>(row + koff + oct * 8);
                vv = *reinterpret_cast<const uint4*>(row + voff + oct * 8);
                const __half2* h2 = reinterpret_cast<const __half2*>(&kv);
                __half2 o2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(h2[i]);
                    o2[i] = __floats2half2_rn(exp2f_ftz(fmaf(f.x, kLog2e, -mloc[2 * i])),
                                              exp2f_ftz(fmaf(f.y, kLog2e, -mloc[2 * i + 1])));
                    const float2 r = __half22float2(o2[i]);   // the normaliser sums the values the MMA sees
                    ssum[2 * i] += r.x;
                    ssum[2 * i + 1] += r.y;
                }
                pk = *reinterpret_cast<const uint4*>(o2);
            }
            *reinterpret_cast<uint4*>(&Ps[warp][pp][oct * 8]) = pk;
            *reinterpret_cast<uint4*>(&Vs[warp][pp][oct * 8]) = vv;
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < kWarpTile / 16; ++ks) {
            const int pk0 = ks * 16;
            uint32_t a[2][4], b[4][2];
            // A fragments (16 d x 16 pos) for d0 = 0 and 16: matrices (pos half, d half) read transposed
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int prow = pk0 + (lm_j >> 1) * 8 + lm_r, dcol = mt * 16 + (lm_j & 1) * 8;
                ldmatrix_x4_trans(ps_base + (prow * kRowHalfs + dcol) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
            }
            // B fragments (16 pos x 8 e) for e0 = 0, 8, 16, 24: two n-tiles per ldmatrix.x4
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
        __syncwarp();
    }

    // ---- fixed-order reduction over the 4 warps (csum aliases the staging tiles: wait until every warp is done)
    __syncthreads();
    {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                csum[warp][mt * 16 + g][nt * 8 + 2 * t] = acc[mt][nt][0];
                csum[warp][mt * 16 + g][nt * 8 + 2 * t + 1] = acc[mt][nt][1];
                csum[warp][mt * 16 + g + 8][nt * 8 + 2 * t] = acc[mt][nt][2];
                csum[warp][mt * 16 + g + 8][nt * 8 + 2 * t + 1] = acc[mt][nt][3];
            }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 4);
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 8);
            ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 16);
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = ssum[i];
        }
    }
    __syncthreads();
    float* out = p.part + ((static_cast<long long>(n) * p.heads + head) * nchunks + chunk_id) * kPartStride;
    for (int i = tid; i < kDh * kDh; i += 128) {
        const int d = i >> 5, e = i & 31;
        out[i] = (csum[0][d][e] + csum[1][d][e]) + (csum[2][d][e] + csum[3][d][e]);
    }
    if (tid < kDh) {
        out[kDh * kDh + tid] = mmax[tid];
        out[kDh * kDh + kDh + tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    }
}

// grid (heads, N), 256 threads: ctx[n][h][d][e] = sum_c part_c[d][e] exp(m_c[d] - M[d]) / sum_c s_c[d] exp(m_c[d] - M[d])
__global__ void __launch_bounds__(256) attn_merge_kernel(const AttnParams p, int nchunks, float* ctx_out) {
    __shared__ float M_s[kDh], S_s[kDh];
    const int h = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
    const float* ph = p.part + (static_cast<long long>(n) * p.heads + h) * nchunks * kPartStride;
    if (tid < kDh) {
        float M = -INFINITY;
        for (int c = 0; c < nchunks; ++c) M = fmaxf(M, ph[c * kPartStride + kDh * kDh + tid]);
        float S = 0.f;
        for (int c = 0; c < nchunks; ++c)
            S += ph[c * kPartStride + kDh * kDh + kDh + tid] *
                 exp2f_ftz((ph[c * kPartStride + kDh * kDh + tid] - M) * kLog2e);
        M_s[tid] = M;
        S_s[tid] = S;
    }
    __syncthreads();
    float* o = ctx_out + (static_cast<long long>(n) * p.heads + h) * kDh * kDh;
    for (int i = tid; i < kDh * kDh; i += 256) {
        const int d = i >> 5;
        const float M = M_s[d];
        float a = 0.f;
        for (int c = 0; c < nchunks; ++c)
            a += ph[c * kPartStride + i] * exp2f_ftz((ph[c * kPartStride + kDh * kDh + d] - M) * kLog2e);
        o[i] = a / S_s[d];
    }
}

// grid (ceil(C/32), N): Weff[n][co][h*32+d] = sum_e Wo[co][h*32+e] * ctx[n][h][d][e] for 32 output channels.
// Plain mode writes Weff (fp16).  Fused-q mode goes on to W'[n][co][c] = g * sum_d' Weff[co][d'] Wq[d'][c] + (co == c),
// the per-sample 1x1 weight of the whole Residual(Rezero(LinearAttention)) block, and b' = g * b_o.
__global__ void __launch_bounds__(256) attn_fold_kernel(const AttnParams p, const float* ctx_in) {
    extern __shared__ float sm[];   // ctx [heads][32][33], then weff tile [32][hidden + 1]
    const int n = blockIdx.y, tid = threadIdx.x;
    const int heads = p.heads, hidden = heads * kDh;
    float* ctx = sm;
    float* wt = sm + heads * kDh * (kDh + 1);
    const float* cn = ctx_in + static_cast<long long>(n) * heads * kDh * kDh;
    for (int i = tid; i < heads * kDh * kDh; i += 256) {
        const int hd = i >> 5, e = i & 31;
        ctx[hd * (kDh + 1) + e] = cn[i];
    }
    __syncthreads();
    const int co0 = blockIdx.x * kFoldRows;
    for (int i = tid; i < kFoldRows * hidden; i += 256) {
        const int r = i / hidden, co = co0 + r;
        const int k = i % hidden, h = k / kDh, d = k % kDh;
        float a = 0.f;
        if (co < p.C) {
            const float* w = p.wo + static_cast<long long>(co) * hidden + h * kDh;
            const float* c = ctx + (h * kDh + d) * (kDh + 1);
#pragma unroll 8
            for (int e = 0; e < kDh; ++e) a += __ldg(w + e) * c[e];
            if (!p.wq) p.weff[(static_cast<long long>(n) * p.C + co) * hidden + k] = __float2half_rn(a);
        }
        wt[r * (hidden + 1) + k] = a;
    }
    if (!p.wq) return;
    __syncthreads();
    // W'[co0 + r][c] for the block's 32 rows: thread = column c, RPT rows each, Wq streamed once per thread
    const int groups = 256 / p.C;                 // C in {64, 128, 256}
    const int rpt = kFoldRows / groups;           // rows per thread: 8, 16 or 32
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
                const float* wr = wt + (r0 + r) * (hidden + 1) + d0;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[r] = fmaf(wr[i], w[i], acc[r]);
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
    dim3 g1(nchunks, p.heads, p.N);
    attn_partial_kernel<<<g1, 128, 0, s>>>(p, nchunks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // merged context lives behind the partials in the same scratch buffer
    float* ctx = p.part + static_cast<long long>(p.N) * p.heads * nchunks * kPartStride;
    dim3 g2(p.heads, p.N);
    attn_merge_kernel<<<g2, 256, 0, s>>>(p, nchunks, ctx);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    dim3 g3((p.C + kFoldRows - 1) / kFoldRows, p.N);
    attn_fold_kernel<<<g3, 256, (p.heads * kDh * (kDh + 1) + kFoldRows * (p.heads * kDh + 1)) * sizeof(float), s>>>(p, ctx);
    return (int)cudaGetLastError();
}

size_t attn_scratch_bytes(int N, int heads, int P, int chunk) {
    return (static_cast<size_t>(N) * heads * attn_chunks(P, chunk) * kPartStride + static_cast<size_t>(N) * heads * kDh * kDh) *
           sizeof(float);
}

}  // namespace usb
