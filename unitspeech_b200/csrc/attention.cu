// LinearAttention context (unitspeech/unitspeech.py:86-96), restructured as a streaming reduction:
//   ctx[h][d][e] = sum_p softmax_p(k[h][d][:])[p] * v[h][e][p]          (softmax over ALL positions, no mask)
//   attn(x)[co][p] = sum_{h,d} Weff[co][h*32+d] * q[h][d][p] + b_o[co],  Weff[co][h*32+d] = sum_e Wo[co][h*32+e] ctx[h][d][e]
// Pass 1 (attn_partial_kernel): each block reduces one (sample, head, chunk of positions) with a chunk-local max
// (online-softmax partial: m[d], s[d] = sum exp(k-m), c[d][e] = sum exp(k-m) v).
// Pass 2 (attn_fold_kernel): merges the chunk partials exactly (rescale by exp(m_chunk - m_global)), normalises, and
// folds to_out into the per-sample fp16 weight consumed by the tcgen05 GEMM kernel.
#include "kernels.h"

namespace usb {

namespace {
__device__ __forceinline__ float exp2f_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr int kDh = 32;            // dim_head
constexpr int kTileP = 128;        // positions staged in shared memory per step
constexpr int kPartStride = kDh * kDh + 2 * kDh;  // ctx + m + s
}  // namespace

int attn_chunks(int P, int chunk) { return (P + chunk - 1) / chunk; }

__global__ void __launch_bounds__(256) attn_partial_kernel(const AttnParams p, int nchunks) {
    __shared__ __align__(16) float stage[2 * kTileP * kDh];   // Ks | Vs, reused for the slice reduction
    __shared__ float red[8][kDh];
    __shared__ float mmax[kDh];
    __shared__ float sum_s[4][kDh];
    float (*Ks)[kDh] = reinterpret_cast<float (*)[kDh]>(stage);                 // exp(k - m)
    float (*Vs)[kDh] = reinterpret_cast<float (*)[kDh]>(stage + kTileP * kDh);
    float (*acc_s)[kDh][kDh + 1] = reinterpret_cast<float (*)[kDh][kDh + 1]>(stage);  // [4][32][33] after the loop

    const int chunk_id = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
    const int hidden = p.heads * kDh;
    const int ld = 3 * hidden;
    const int p0 = chunk_id * p.chunk;
    const int p1 = min(p.P, p0 + p.chunk);
    const __half* base = p.qkv + static_cast<long long>(n) * p.P * ld;
    const int koff = hidden + head * kDh;
    const int voff = 2 * hidden + head * kDh;
    const int tid = threadIdx.x;

    // ---- chunk-local max of k per d
    {
        const int d = tid & 31, sl = tid >> 5;
        float m = -INFINITY;
        for (int pos = p0 + sl; pos < p1; pos += 8)
            m = fmaxf(m, __half2float(base[static_cast<long long>(pos) * ld + koff + d]));
        red[sl][d] = m;
        __syncthreads();
        if (tid < kDh) {
            float mm = red[0][tid];
#pragma unroll
            for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i][tid]);
            mmax[tid] = mm;
        }
        __syncthreads();
    }

    // ---- accumulate: 4 position slices x 64 threads, each thread 2 d x 8 e
    const int slice = tid >> 6;
    const int t64 = tid & 63;
    const int dp = t64 >> 2;   // d pair 0..15
    const int eo = t64 & 3;    // e octet 0..3
    float acc[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float s0 = 0.f, s1 = 0.f;

    for (int t0 = p0; t0 < p1; t0 += kTileP) {
        // stage: 2 threads per position, 16 channels of k and v each
        {
            const int pp = tid >> 1, hf = tid & 1;
            const int pos = t0 + pp;
            float kf[16], vf[16];
            if (pos < p1) {
                const __half* kp = base + static_cast<long long>(pos) * ld + koff + hf * 16;
                const __half* vp = base + static_cast<long long>(pos) * ld + voff + hf * 16;
                const uint4 k0 = *reinterpret_cast<const uint4*>(kp);
                const uint4 k1 = *reinterpret_cast<const uint4*>(kp + 8);
                const uint4 v0 = *reinterpret_cast<const uint4*>(vp);
                const uint4 v1 = *reinterpret_cast<const uint4*>(vp + 8);
                const __half2* hk0 = reinterpret_cast<const __half2*>(&k0);
                const __half2* hk1 = reinterpret_cast<const __half2*>(&k1);
                const __half2* hv0 = reinterpret_cast<const __half2*>(&v0);
                const __half2* hv1 = reinterpret_cast<const __half2*>(&v1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 a = __half22float2(hk0[i]); kf[2 * i] = a.x; kf[2 * i + 1] = a.y;
                    a = __half22float2(hk1[i]); kf[8 + 2 * i] = a.x; kf[8 + 2 * i + 1] = a.y;
                    a = __half22float2(hv0[i]); vf[2 * i] = a.x; vf[2 * i + 1] = a.y;
                    a = __half22float2(hv1[i]); vf[8 + 2 * i] = a.x; vf[8 + 2 * i + 1] = a.y;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    Ks[pp][hf * 16 + i] = exp2f_ftz((kf[i] - mmax[hf * 16 + i]) * 1.4426950408889634f);
                    Vs[pp][hf * 16 + i] = vf[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    Ks[pp][hf * 16 + i] = 0.f;
                    Vs[pp][hf * 16 + i] = 0.f;
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < kTileP / 4; ++i) {
            const int pp = slice * (kTileP / 4) + i;
            const float2 w = *reinterpret_cast<const float2*>(&Ks[pp][dp * 2]);
            const float4 va = *reinterpret_cast<const float4*>(&Vs[pp][eo * 8]);
            const float4 vb = *reinterpret_cast<const float4*>(&Vs[pp][eo * 8 + 4]);
            acc[0][0] += w.x * va.x; acc[0][1] += w.x * va.y; acc[0][2] += w.x * va.z; acc[0][3] += w.x * va.w;
            acc[0][4] += w.x * vb.x; acc[0][5] += w.x * vb.y; acc[0][6] += w.x * vb.z; acc[0][7] += w.x * vb.w;
            acc[1][0] += w.y * va.x; acc[1][1] += w.y * va.y; acc[1][2] += w.y * va.z; acc[1][3] += w.y * va.w;
            acc[1][4] += w.y * vb.x; acc[1][5] += w.y * vb.y; acc[1][6] += w.y * vb.z; acc[1][7] += w.y * vb.w;
            s0 += w.x;
            s1 += w.y;
        }
        __syncthreads();
    }

    // ---- reduce the 4 slices and write the partial
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc_s[slice][dp * 2 + i][eo * 8 + j] = acc[i][j];
    if (eo == 0) {
        sum_s[slice][dp * 2] = s0;
        sum_s[slice][dp * 2 + 1] = s1;
    }
    __syncthreads();
    float* out = p.part + ((static_cast<long long>(n) * p.heads + head) * nchunks + chunk_id) * kPartStride;
    for (int i = tid; i < kDh * kDh; i += 256) {
        const int d = i >> 5, e = i & 31;
        out[i] = acc_s[0][d][e] + acc_s[1][d][e] + acc_s[2][d][e] + acc_s[3][d][e];
    }
    if (tid < kDh) {
        out[kDh * kDh + tid] = mmax[tid];
        out[kDh * kDh + kDh + tid] = sum_s[0][tid] + sum_s[1][tid] + sum_s[2][tid] + sum_s[3][tid];
    }
}

// grid (ceil(C/64), N): merge chunk partials -> ctx (all heads) in smem, then Weff rows for 64 output channels
__global__ void __launch_bounds__(256) attn_fold_kernel(const AttnParams p, int nchunks) {
    extern __shared__ float ctx[];  // [heads][32][33]
    __shared__ float M_s[8][kDh], S_s[8][kDh];
    const int n = blockIdx.y;
    const int tid = threadIdx.x;
    const int heads = p.heads;
    const int hidden = heads * kDh;
    const float* part_n = p.part + static_cast<long long>(n) * heads * nchunks * kPartStride;

    // global max and normaliser per (head, d)
    for (int i = tid; i < heads * kDh; i += 256) {
        const int h = i / kDh, d = i % kDh;
        const float* ph = part_n + static_cast<long long>(h) * nchunks * kPartStride;
        float M = -INFINITY;
        for (int c = 0; c < nchunks; ++c) M = fmaxf(M, ph[c * kPartStride + kDh * kDh + d]);
        float S = 0.f;
        for (int c = 0; c < nchunks; ++c)
            S += ph[c * kPartStride + kDh * kDh + kDh + d] * __expf(ph[c * kPartStride + kDh * kDh + d] - M);
        M_s[h][d] = M;
        S_s[h][d] = S;
    }
    __syncthreads();
    for (int i = tid; i < heads * kDh * kDh; i += 256) {
        const int h = i / (kDh * kDh), r = i % (kDh * kDh), d = r >> 5, e = r & 31;
        const float* ph = part_n + static_cast<long long>(h) * nchunks * kPartStride;
        const float M = M_s[h][d];
        float a = 0.f;
        for (int c = 0; c < nchunks; ++c)
            a += ph[c * kPartStride + r] * __expf(ph[c * kPartStride + kDh * kDh + d] - M);
        ctx[(h * kDh + d) * (kDh + 1) + e] = a / S_s[h][d];
    }
    __syncthreads();
    // Weff[n][co][h*32+d] = sum_e Wo[co][h*32+e] * ctx[h][d][e]
    const int co0 = blockIdx.x * 64;
    for (int i = tid; i < 64 * hidden; i += 256) {
        const int co = co0 + i / hidden;
        if (co >= p.C) break;
        const int k = i % hidden, h = k / kDh, d = k % kDh;
        const float* w = p.wo + static_cast<long long>(co) * hidden + h * kDh;
        const float* c = ctx + (h * kDh + d) * (kDh + 1);
        float a = 0.f;
#pragma unroll 8
        for (int e = 0; e < kDh; ++e) a += __ldg(w + e) * c[e];
        p.weff[(static_cast<long long>(n) * p.C + co) * hidden + k] = __float2half_rn(a);
    }
}

int launch_attn_context(const AttnParams& p, cudaStream_t s) {
    if (p.heads > 8 || p.heads < 1) return (int)cudaErrorInvalidValue;
    const int nchunks = attn_chunks(p.P, p.chunk);
    dim3 g1(nchunks, p.heads, p.N);
    attn_partial_kernel<<<g1, 256, 0, s>>>(p, nchunks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    dim3 g2((p.C + 63) / 64, p.N);
    attn_fold_kernel<<<g2, 256, p.heads * kDh * (kDh + 1) * sizeof(float), s>>>(p, nchunks);
    return (int)cudaGetLastError();
}

}  // namespace usb
