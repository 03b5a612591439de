// Generalised implicit-GEMM convolution for sm_100a: TMA -> swizzled smem -> tcgen05.mma (fp16 operands, fp32
// accumulators in TMEM) -> fused epilogue (bias, GroupNorm partial statistics, Rezero residual, mask, fp16 store).
// See conv_igemm.h for the GEMM view.  Persistent: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...
// Warp roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31, one output pixel per thread).
// Two 256-column accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_igemm.h"
#include "ptx.cuh"

namespace usb {

namespace {

struct TileCoord {
    int ph, n, y0, x0, nt;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int idx) {
    TileCoord t;
    t.nt = idx % p.n_tiles_n; idx /= p.n_tiles_n;
    int tx = idx % p.tiles_x; idx /= p.tiles_x;
    int ty = idx % p.tiles_y; idx /= p.tiles_y;
    t.n = idx % p.N;
    t.ph = idx / p.N;
    t.y0 = ty * p.BH;
    t.x0 = tx * p.BW;
    return t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void flush_stats(long long* dst, float s, float ss, int lane) {
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(dst),
                  static_cast<unsigned long long>(__float2ll_rn(s * kStatSumScale)));
        atomicAdd(reinterpret_cast<unsigned long long*>(dst + 1),
                  static_cast<unsigned long long>(__float2ll_rn(ss * kStatSqScale)));
    }
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const ConvParams p, const __grid_constant__ CUtensorMap map_a0,
                  const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b,
                  int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // 1024-byte aligned tile region (SWIZZLE_128B atoms are 1024 B)
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
    const uint32_t stage_bytes = 16384u + b_bytes;
    const int stages = p.stages;
    const int ksteps = p.taps * (p.chunks0 + p.chunks1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(&tmem_base_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(p, tile);
                const int bz = p.b_batch_mode == 0 ? 0 : (p.b_batch_mode == 1 ? tc.ph : tc.n);
                int kstep = 0;
                for (int t = 0; t < p.taps; ++t) {
                    const ConvTap tap = p.tap[tc.ph * p.taps + t];
                    for (int src = 0; src < 2; ++src) {
                        const int chunks = src ? p.chunks1 : p.chunks0;
                        const CUtensorMap* ma = src ? &map_a1 : &map_a0;
                        for (int cc = 0; cc < chunks; ++cc, ++kstep) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u, 100 + stage);
                            const uint32_t sa = tiles_base + stage * stage_bytes;
                            mbar_arrive_expect_tx(&full_bar[stage], 16384u + b_bytes);
                            tma_load_5d(reinterpret_cast<void*>(__cvta_shared_to_generic(sa)), ma, &full_bar[stage],
                                        tap.c + cc * kConvBK, tc.x0 + tap.dx, tap.p, tc.y0 + tap.dy, tc.n);
                            tma_load_3d(reinterpret_cast<void*>(__cvta_shared_to_generic(sa + 16384u)), &map_b,
                                        &full_bar[stage], kstep * kConvBK, tc.nt * p.BN, bz);
                            if (++stage == stages) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(p.BN));
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1u, 200 + as);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * 256u;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full_bar[stage], phase, 300 + stage);
                    tc_fence_after();
                    const uint32_t sa = tiles_base + stage * stage_bytes;
                    const uint32_t sb = sa + 16384u;
#pragma unroll
                    for (int k = 0; k < kConvBK / 16; ++k) {
                        const uint64_t da = umma_desc_sw128(sa + k * 32);
                        const uint64_t db = umma_desc_sw128(sb + k * 32);
                        tc_mma_f16(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (ks == ksteps - 1) tc_commit(&tmem_full_bar[as]);
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue
        const int ew = warp & 3;
        const int row = ew * 32 + lane;
        const int bw_shift = 31 - __clz(p.BW);
        const int ty = row >> bw_shift;
        const int tx = row & (p.BW - 1);
        const int cpg = p.stats ? p.Cout / p.groups : 0;
        const float rs = p.res_scale ? __ldg(p.res_scale) : 1.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const TileCoord tc = decode_tile(p, tile);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int y = tc.y0 + ty, x = tc.x0 + tx;
            const bool valid = x < p.Wm;
            const int yo = y * p.oy_mul + p.oy_off[tc.ph];
            const int xo = x * p.ox_mul + p.ox_off[tc.ph];
            const long long off = tc.n * p.o_sn + yo * p.o_sy + xo * p.o_sx + tc.nt * p.BN;
            const float m = (p.mask && valid) ? __ldg(p.mask + static_cast<long long>(tc.n) * p.mask_stride + xo) : 1.f;
            long long* stats_n = p.stats ? p.stats + static_cast<long long>(tc.n) * p.groups * 2 : nullptr;

            mbar_wait(&tmem_full_bar[as], aphase, 400 + as);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as) * 256u;
            const int nchunks = p.BN >> 5;
            float gs = 0.f, gss = 0.f;  // running sums of the current GroupNorm group (cpg >= 32)
            for (int j = 0; j < nchunks; ++j) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + j * 32, v);
                tmem_ld_wait();
                if (j == nchunks - 1) {
                    // all of this thread's accumulator reads are done: hand the stage back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(&tmem_empty_bar[as]);
                }
                float f[32];
                const int c0 = tc.nt * p.BN + j * 32;
                if (p.bias) {
                    const float4* bp = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b4 = __ldg(bp + q);
                        f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) + b4.x;
                        f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) + b4.y;
                        f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) + b4.z;
                        f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) + b4.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) f[q] = __uint_as_float(v[q]);
                }
                if (stats_n) {
                    const float vm = valid ? 1.f : 0.f;
                    if (cpg >= 32) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const float a = f[q] * vm;
                            gs += a;
                            gss += a * a;
                        }
                        if (((c0 + 32) % cpg) == 0 || j == nchunks - 1) {
                            flush_stats(stats_n + (c0 / cpg) * 2, gs, gss, lane);
                            gs = gss = 0.f;
                        }
                    } else if (cpg == 16) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float s = 0.f, ss = 0.f;
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float a = f[16 * h + q] * vm;
                                s += a;
                                ss += a * a;
                            }
                            flush_stats(stats_n + ((c0 + 16 * h) / 16) * 2, s, ss, lane);
                        }
                    } else {  // cpg == 8
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            float s = 0.f, ss = 0.f;
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float a = f[8 * h + q] * vm;
                                s += a;
                                ss += a * a;
                            }
                            flush_stats(stats_n + ((c0 + 8 * h) / 8) * 2, s, ss, lane);
                        }
                    }
                }
                if (valid) {
                    if (p.res) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.res + off + j * 32);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 r4 = __ldg(rp + q);
                            const __half2* h2 = reinterpret_cast<const __half2*>(&r4);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 r2 = __half22float2(h2[e]);
                                f[8 * q + 2 * e] = f[8 * q + 2 * e] * rs + r2.x;
                                f[8 * q + 2 * e + 1] = f[8 * q + 2 * e + 1] * rs + r2.y;
                            }
                        }
                    }
                    uint4* op = reinterpret_cast<uint4*>(p.out + off + j * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 o;
                        o.x = pack_half2(f[8 * q + 0] * m, f[8 * q + 1] * m);
                        o.y = pack_half2(f[8 * q + 2] * m, f[8 * q + 3] * m);
                        o.z = pack_half2(f[8 * q + 4] * m, f[8 * q + 5] * m);
                        o.w = pack_half2(f[8 * q + 6] * m, f[8 * q + 7] * m);
                        op[q] = o;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int launch_conv_igemm(const ConvParams& p, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                      int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr_set = true;
    }
    const long long total = static_cast<long long>(p.phases) * p.N * p.tiles_y * p.tiles_x * p.n_tiles_n;
    if (total <= 0 || total > 0x7fffffffLL) return static_cast<int>(cudaErrorInvalidValue);
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    const size_t smem = 1024 + static_cast<size_t>(p.stages) * (16384 + static_cast<size_t>(p.BN) * 128);
    if (smem > static_cast<size_t>(kConvSmemBytes)) return static_cast<int>(cudaErrorInvalidValue);
    conv_igemm_kernel<<<grid, kConvThreads, smem, stream>>>(p, a0, a1, b, static_cast<int>(total));
    return static_cast<int>(cudaGetLastError());
}

}  // namespace usb
