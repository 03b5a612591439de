// Generalised implicit-GEMM convolution for sm_100a: TMA -> swizzled smem -> tcgen05.mma (fp16 operands, fp32
// accumulators in TMEM) -> fused epilogue (bias, GroupNorm partial statistics, Rezero residual, mask) -> fp16 tile
// staged in swizzled smem -> TMA store.  See conv_igemm.h for the GEMM view.
//
// Persistent: one CTA per SM walks tiles blockIdx.x, +gridDim.x, ...   384 threads:
//   warp 0      TMA producer   (whole warp runs the loop, one elected lane issues: keeps ptxas on the uniform path)
//   warp 1      MMA issuer     (same pattern; 4 x tcgen05.mma per 64-channel K step, tcgen05.commit frees the slot)
//   warp 2      TMEM allocator (512 columns = two 256-column accumulator stages)
//   warps 4-11  epilogue, two groups of four warps; group g owns output columns [g*BN/2, (g+1)*BN/2) in 64-column
//               slabs; warp w reads TMEM lanes 32*(w%4)..+31 (one output pixel per thread).
// The epilogue of tile i overlaps the MMAs of tile i+1 (double-buffered accumulators).
#include <type_traits>

#include "conv_igemm.h"
#include "pdl.h"
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

// Sums NV per-lane values across the warp with NV-1 + (5 - log2 NV) shuffles (recursive halving): afterwards lane L
// with (L & (32/NV - 1)) == 0 holds in v[0] the warp total of item L / (32/NV).  Fixed order -> deterministic.
template <int NV>
__device__ __forceinline__ void warp_reduce_items(float (&v)[NV], int lane) {
    int n = NV;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        if (n > 1) {
            n >>= 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < NV / 2; ++i) {
                if (i < n) {
                    const float send = upper ? v[i] : v[i + n];
                    const float keep = upper ? v[i + n] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
}

// One 64-column slab of one output pixel row: bias, statistics, residual, mask, fp16 pack, swizzled smem store.
// G = GroupNorm groups intersecting the slab (64/cpg, or 1 when a group spans >= 64 channels).
template <int G>
__device__ __forceinline__ void epilogue_slab(const ConvParams& p, const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                              int c_glob, bool valid, float m, float rs, bool has_res,
                                              const uint4 (&rr)[8], uint32_t stage_row, int row, long long* stats_n,
                                              int cpg, int lane, float& amax) {
    constexpr int CPG8 = 8 / G;  // 8-column chunks per group inside the slab (G=8 -> 1, 4 -> 2, 2 -> 4, 1 -> 8)
    float acc[2 * G];
#pragma unroll
    for (int i = 0; i < 2 * G; ++i) acc[i] = 0.f;
    const float vm = valid ? 1.f : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(q < 4 ? v0[q * 8 + e] : v1[(q - 4) * 8 + e]);
        if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c_glob + q * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c_glob + q * 8 + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        if (stats_n) {
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                s += f[e];
                ss = fmaf(f[e], f[e], ss);
            }
            acc[q / CPG8] += s * vm;
            acc[G + q / CPG8] += ss * vm;
        }
        if (has_res) {      // rr: the residual row, loaded before the accumulator was waited for (zeros outside the image)
            const uint4 r4 = rr[q];
            const __half2* h2 = reinterpret_cast<const __half2*>(&r4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 r2 = __half22float2(h2[e]);
                f[2 * e] = fmaf(f[2 * e], rs, r2.x);
                f[2 * e + 1] = fmaf(f[2 * e + 1], rs, r2.y);
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            f[e] *= m;
            amax = fmaxf(amax, fabsf(f[e]));
        }
        // 16-byte chunk q of row `row`, 128-byte swizzle (chunk index XOR row%8) = the layout the TMA store expects
        sts128(stage_row + (static_cast<uint32_t>(q ^ (row & 7)) << 4), pack_f16x2_sat(f[0], f[1]),
               pack_f16x2_sat(f[2], f[3]), pack_f16x2_sat(f[4], f[5]), pack_f16x2_sat(f[6], f[7]));
    }
    if (stats_n) {
        warp_reduce_items<2 * G>(acc, lane);
        constexpr int LPI = 32 / (2 * G);  // lanes per item
        if ((lane & (LPI - 1)) == 0) {
            const int item = lane / LPI;          // 0..G-1 sums, G..2G-1 sums of squares
            const int gi = item % G;
            const bool sq = item >= G;
            const int group = c_glob / cpg + gi;  // G > 1: cpg = 64/G divides 64; G == 1: slab inside one group
            const float scale = sq ? kStatSqScale : kStatSumScale;
            atomicAdd(reinterpret_cast<unsigned long long*>(stats_n + group * 2 + (sq ? 1 : 0)),
                      static_cast<unsigned long long>(__float2ll_rn(acc[0] * scale)));
        }
    }
}

// Epilogue of the pixels-on-M kernels (warps 4-11): TMEM -> registers -> bias / statistics / residual / mask -> fp16 ->
// per-warp swizzled staging -> TMA store.  `staging_base`: 8 warps x 8 KB of shared memory.
__device__ __forceinline__ void plain_epilogue(const ConvParams& p, const CUtensorMap& map_out, uint32_t staging_base,
                                               uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int total_tiles,
                                               int warp, int lane) {
    const int ew = warp & 3;                 // TMEM lane quarter
    const int grp = (warp - 4) >> 2;         // column half
    const int row = ew * 32 + lane;
    const int bw_shift = 31 - __clz(p.BW);
    const int ty = row >> bw_shift;
    const int tx = row & (p.BW - 1);
    const int cpg = p.stats ? p.Cout / p.groups : 64;
    const float rs = p.res_scale ? __ldg(p.res_scale) : 1.f;
    // 64-column slabs of the tile: group 0 takes the first half (rounded up), group 1 the rest (BN == 64: none)
    const int total_slabs = p.BN >> 6;
    const int slab_begin = grp == 0 ? 0 : (total_slabs + 1) >> 1;
    const int slabs_split = grp == 0 ? (total_slabs + 1) >> 1 : total_slabs - slab_begin;
    // BN == 64 (one slab per tile): the two groups take alternate tiles instead, so that eight warps, not four, share the
    // epilogue of the narrow layers (their tiles are short: the epilogue, not the tensor pipe, paces them)
    const bool alternate = total_slabs == 1;
    // per-warp staging (32 pixel rows x 128 B) and per-warp TMA stores: no cross-warp barrier in the epilogue
    // (two 4 KB buffers per warp, alternating, so a store only waits for the one issued two slabs earlier)
    const uint32_t warp_buf0 = staging_base + static_cast<uint32_t>(grp * 4 + ew) * 8192u;
    uint32_t buf_sel = 0;
    float amax = 0.f;                             // largest |value| packed by this thread (saturation report)
    const int q0 = ew * 32;                       // first tile pixel of this warp
    const int wty0 = q0 >> bw_shift, wtx0 = q0 & (p.BW - 1);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const TileCoord tc = decode_tile(p, tile);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int y = tc.y0 + ty, x = tc.x0 + tx;
        const bool valid = x < p.Wm;
        const int yo = y * p.oy_mul + p.oy_off[tc.ph];
        const int xo = x * p.ox_mul + p.ox_off[tc.ph];
        const float m = (p.mask && valid) ? __ldg(p.mask + static_cast<long long>(tc.n) * p.mask_stride + xo) : 1.f;
        long long* stats_n = p.stats ? p.stats + static_cast<long long>(tc.n) * p.groups * 2 : nullptr;
        const __half* res_px = (p.res && valid) ? p.res + (tc.n * p.o_sn + yo * p.o_sy + xo * p.o_sx) : nullptr;
        const int slabs = alternate ? ((it & 1) == grp ? 1 : 0) : slabs_split;
        const int sl_begin = alternate ? 0 : slab_begin;
        // the residual does not depend on the accumulator: its loads are in flight while this thread waits for the MMAs
        uint4 rr[8];
        auto load_res = [&](int sl) {
            const __half* rrow = res_px ? res_px + tc.nt * p.BN + (sl_begin + sl) * 64 : nullptr;
#pragma unroll
            for (int q = 0; q < 8; ++q) rr[q] = rrow ? __ldg(reinterpret_cast<const uint4*>(rrow) + q) : make_uint4(0u, 0u, 0u, 0u);
        };
        if (p.res && slabs > 0) load_res(0);

        mbar_wait_a(tfull0 + as * 8, aphase, 400 + as);
        tc_fence_after();
        if (slabs == 0 || (p.dbg_flags & 2)) {
            tc_fence_before();
            mbar_arrive_a(tempty0 + as * 8);
            continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as) * 256u;
        for (int sl = 0; sl < slabs; ++sl) {
            const int col0 = (sl_begin + sl) * 64;
            const int c_glob = tc.nt * p.BN + col0;
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(taddr + col0, v0);
            tmem_ld_32x32(taddr + col0 + 32, v1);
            tmem_ld_wait();
            if (sl == slabs - 1) {
                // this thread's accumulator reads are complete: hand the TMEM stage back to the MMA warp
                tc_fence_before();
                mbar_arrive_a(tempty0 + as * 8);
            }
            // this staging buffer is free once the store issued from it two slabs ago has been read
            const uint32_t warp_buf = warp_buf0 + buf_sel * 4096u;
            const uint32_t stage_row = warp_buf + static_cast<uint32_t>(lane) * 128u;
            buf_sel ^= 1u;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            const bool has_res = p.res != nullptr;
            if (cpg >= 64) epilogue_slab<1>(p, v0, v1, c_glob, valid, m, rs, has_res, rr, stage_row, lane, stats_n, cpg, lane, amax);
            else if (cpg == 32) epilogue_slab<2>(p, v0, v1, c_glob, valid, m, rs, has_res, rr, stage_row, lane, stats_n, cpg, lane, amax);
            else if (cpg == 16) epilogue_slab<4>(p, v0, v1, c_glob, valid, m, rs, has_res, rr, stage_row, lane, stats_n, cpg, lane, amax);
            else epilogue_slab<8>(p, v0, v1, c_glob, valid, m, rs, has_res, rr, stage_row, lane, stats_n, cpg, lane, amax);
            if (has_res && sl + 1 < slabs) load_res(sl + 1);      // next slab's residual: in flight during the store below
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
            __syncwarp();
            if (lane == 0) {
                tma_store_5d_a(&map_out, warp_buf, c_glob + p.ox_off[tc.ph] * p.out_c_phase_mul, tc.x0 + wtx0,
                               p.oy_off[tc.ph], tc.y0 + wty0, tc.n);
                tma_store_commit();
            }
        }
    }
    if (p.sat && amax >= 65504.f) atomicAdd(p.sat, 1ull);
    if (lane == 0) tma_store_wait_all<0>();
}

}  // namespace

__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const ConvParams p, const __grid_constant__ CUtensorMap map_a0,
                  const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_out, int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(8) uint64_t scratch_bar;   // experiments only

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // 1024-byte aligned tile region (SWIZZLE_128B atoms are 1024 B)
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;
    const uint32_t stage_bytes = 16384u + b_bytes;
    const int stages = p.stages;
    const int ksteps = p.taps * (p.chunks0 + p.chunks1);
    const uint32_t full0 = smem_u32_pinned(&full_bar[0]), empty0 = smem_u32_pinned(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32_pinned(&tmem_full_bar[0]), tempty0 = smem_u32_pinned(&tmem_empty_bar[0]);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kConvEpilogueThreads);
        }
        mbar_init(&scratch_bar, 1u << 20);
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
    pdl_wait_and_trigger();   // everything above is local set-up; below this line the predecessor's output is read

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer
        int stage = 0;
        uint32_t phase = 0;
        const bool ld_a = !(p.dbg_flags & 4), ld_b = !(p.dbg_flags & 1);
        const uint32_t tx_bytes = (ld_a ? 16384u : 0u) + (ld_b ? b_bytes : 0u);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const TileCoord tc = decode_tile(p, tile);
            const int bz = p.b_batch_mode == 0 ? 0 : (p.b_batch_mode == 1 ? tc.ph : tc.n);
            const int bn0 = tc.nt * p.BN;
            int kcoord = 0;
            for (int t = 0; t < p.taps; ++t) {
                const ConvTap tap = p.tap[tc.ph * p.taps + t];
                const int cx = tc.x0 + tap.dx, cy = tc.y0 + tap.dy;
                for (int src = 0; src < 2; ++src) {
                    const int chunks = src ? p.chunks1 : p.chunks0;
                    const CUtensorMap* ma = src ? &map_a1 : &map_a0;
                    for (int cc = 0; cc < chunks; ++cc, kcoord += kConvBK) {
                        mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 100 + stage);
                        if (elect_one()) {
                            const uint32_t sa = tiles_base + stage * stage_bytes;
                            const uint32_t fb = full0 + stage * 8;
                            mbar_arrive_expect_tx_a(fb, tx_bytes);
                            if (ld_a) tma_load_5d_a(sa, ma, fb, tap.c + cc * kConvBK, cx, tap.p, cy, tc.n);
                            if (ld_b) tma_load_3d_a(sa + 16384u, &map_b, fb, kcoord, bn0, bz);
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer
        const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(p.BN));
        // descriptor = constant high word | (smem byte address >> 4) in the low word
        const uint32_t desc_hi = static_cast<uint32_t>(umma_desc_sw128(0) >> 32);
        const uint32_t a_lo0 = (tiles_base & 0x3FFFFu) >> 4;
        const uint32_t stage_lo = stage_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait_a(tempty0 + as * 8, aphase ^ 1u, 200 + as);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * 256u;
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait_a(full0 + stage * 8, phase, 300 + stage);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + stage * stage_lo;
                    const uint32_t b_lo = a_lo + (16384u >> 4);
#pragma unroll
                    for (int k = 0; k < kConvBK / 16; ++k) {
                        const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k);
                        const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k);
                        tc_mma_f16(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
                        if (p.dbg_flags & 8) tc_mma_f16(tmem_d, da, db, idesc, 1u);  // experiment: double tensor work
                    }
                    if (p.dbg_flags & 16) tc_commit_a(smem_u32(&scratch_bar));   // experiment: cost of a commit
                    if (p.dbg_flags & 32) { tc_commit_a(smem_u32(&scratch_bar)); tc_commit_a(smem_u32(&scratch_bar)); }
                    tc_commit_a(empty0 + stage * 8);  // frees the smem slot when these MMAs retire
                    if (ks == ksteps - 1) tc_commit_a(tfull0 + as * 8);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        plain_epilogue(p, map_out, tiles_base + stages * stage_bytes, tmem_base, tfull0, tempty0, total_tiles, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// 1-D halo variant of the kernel above (the vocoder's Conv1d layers, alias "same"-padded dilated convs of
// unitspeech/vocoder/models.py:46-74): with one TMA load per (tap, chunk) every activation byte crosses L2 -> shared
// memory `taps` times (3 / 7 / 11), and at 24-192 channels that traffic, not the tensor pipe, bounds the layer.  Here the
// tile's 128 positions are loaded ONCE per 64-channel chunk together with the (k-1)*dilation positions the taps reach
// (box of p.h1d_rows rows starting at x0 + h1d_dx0; out-of-range positions are zero-filled = the conv's zero padding) and
// tap t is the UMMA descriptor whose start address is shifted by (dx_t - h1d_dx0) rows of 128 bytes.  The 128-byte
// swizzle is a function of the absolute shared-memory address, so a start address that is not a multiple of 1024 reads
// the rows the TMA wrote (same property conv_igemm_halo_kernel relies on).  Weight tiles are either streamed through
// a ring (one per (chunk, tap)) or, when taps x chunks tiles fit (the 24- / 48-channel stages), loaded once per CTA and
// kept resident.  K = 16 slices that hold only channel padding (h1d_klast) are not issued.
// K order: chunk-major, taps inside (the packed weight's K coordinate is (tap * chunks + chunk) * 64).
// =====================================================================================================================
template <bool WRES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv1d_halo_kernel(const ConvParams p, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const __grid_constant__ CUtensorMap map_out, int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];      // weight ring
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t afull_bar[4];     // activation ring
    __shared__ __align__(8) uint64_t aempty_bar[4];
    __shared__ __align__(8) uint64_t wres_bar;         // resident weights loaded
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_bytes = static_cast<uint32_t>(p.h1d_rows) * 128u;            // one haloed activation chunk
    const uint32_t a_stride = (a_bytes + 1023u) & ~1023u;
    const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * 128u;                  // one weight tile
    const int na = p.h1d_na;
    const int chunks = p.chunks0;
    const int wtiles = WRES ? p.taps * chunks : p.stages;                  // weight tiles held in shared memory
    const uint32_t w_base = tiles_base + static_cast<uint32_t>(na) * a_stride;
    const uint32_t staging_base = w_base + static_cast<uint32_t>(wtiles) * b_bytes;
    const uint32_t full0 = smem_u32_pinned(&full_bar[0]), empty0 = smem_u32_pinned(&empty_bar[0]);
    const uint32_t afull0 = smem_u32_pinned(&afull_bar[0]), aempty0 = smem_u32_pinned(&aempty_bar[0]);
    const uint32_t tfull0 = smem_u32_pinned(&tmem_full_bar[0]), tempty0 = smem_u32_pinned(&tmem_empty_bar[0]);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 3 && lane == 0) tma_prefetch_desc(&map_a);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 8; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&afull_bar[i], 1);
            mbar_init(&aempty_bar[i], 1);
        }
        mbar_init(&wres_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kConvEpilogueThreads);
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
    pdl_wait_and_trigger();

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer
        if (WRES) {      // weights: every (tap, chunk) tile once, in packed-K order
            if (elect_one()) {
                const uint32_t wb = smem_u32(&wres_bar);
                mbar_arrive_expect_tx_a(wb, static_cast<uint32_t>(wtiles) * b_bytes);
                for (int i = 0; i < wtiles; ++i) tma_load_3d_a(w_base + i * b_bytes, &map_b, wb, i * kConvBK, 0, 0);
            }
            __syncwarp();
        }
        if (!WRES) {     // weights streamed: one tile per (chunk, tap), in the order the MMA warp consumes them
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int bn0 = (tile % p.n_tiles_n) * p.BN;
                for (int cc = 0; cc < chunks; ++cc) {
                    for (int t = 0; t < p.taps; ++t) {
                        mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 100 + stage);
                        if (elect_one()) {
                            const uint32_t fb = full0 + stage * 8;
                            if (p.dbg_flags & 1) {      // experiment: no weight traffic (results are wrong)
                                mbar_arrive_expect_tx_a(fb, 0u);
                            } else {
                                mbar_arrive_expect_tx_a(fb, b_bytes);
                                tma_load_3d_a(w_base + stage * b_bytes, &map_b, fb, (t * chunks + cc) * kConvBK, bn0, 0);
                            }
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ---------------------------------------------------- activation producer (own warp: never waits behind a weight slot)
        int ab = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const TileCoord tc = decode_tile(p, tile);
            for (int cc = 0; cc < chunks; ++cc) {
                mbar_wait_a(aempty0 + ab * 8, aph ^ 1u, 500 + ab);
                if (elect_one()) {
                    const uint32_t fb = afull0 + ab * 8;
                    mbar_arrive_expect_tx_a(fb, a_bytes);
                    tma_load_5d_a(tiles_base + ab * a_stride, &map_a, fb, cc * kConvBK, tc.x0 + p.h1d_dx0, 0, 0, tc.n);
                }
                __syncwarp();
                if (++ab == na) { ab = 0; aph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer
        // At 24-48 channels a tap is two or three N = 64 instructions (~32 tensor cycles each): the issue loop itself is
        // the critical path (ncu: a generic loop spent ~650 cycles per tap on ~95 scalar instructions, tensor pipe 10 %
        // active).  So the tap loop is fully unrolled with the per-tap row offsets in registers, the K = 16 slice count is
        // a compile-time constant of the instantiated chunk body, and with resident weights one elected lane issues a whole
        // chunk without reconverging.
        const uint32_t idesc = umma_idesc_f16(static_cast<uint32_t>(p.BN));
        const uint64_t dhi = static_cast<uint64_t>(static_cast<uint32_t>(umma_desc_sw128(0) >> 32)) << 32;
        const int taps = p.taps, klast = p.h1d_klast, stages = p.stages;
        uint32_t toff[kConvMaxTaps];      // (dx_t - dx0) rows of 128 B, in 16-byte descriptor units
#pragma unroll
        for (int t = 0; t < kConvMaxTaps; ++t)
            toff[t] = t < taps ? static_cast<uint32_t>(p.tap[t].dx - p.h1d_dx0) * 8u : 0u;
        const uint32_t a_lo0 = (tiles_base & 0x3FFFFu) >> 4, a_step = a_stride >> 4;
        const uint32_t w_lo0 = (w_base & 0x3FFFFu) >> 4, b_step = b_bytes >> 4;
        const uint32_t w_tap_step = static_cast<uint32_t>(chunks) * b_step;      // resident tiles are stored tap-major
        int stage = 0, ab = 0;
        uint32_t phase = 0, aph = 0;
        int it = 0;
        if (WRES) {
            mbar_wait_a(smem_u32(&wres_bar), 0u, 600);
            tc_fence_after();
        }
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait_a(tempty0 + as * 8, aphase ^ 1u, 200 + as);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * 256u;
            for (int cc = 0; cc < chunks; ++cc) {
                mbar_wait_a(afull0 + ab * 8, aph, 700 + ab);
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(ab) * a_step;
                const bool last_cc = cc == chunks - 1;
                const uint32_t acc0 = cc != 0 ? 1u : 0u;      // the tile's first MMA overwrites the accumulator
                auto chunk = [&](auto nkc) {
                    constexpr int NK = decltype(nkc)::value;
                    if constexpr (WRES) {
                        if (elect_one()) {
                            uint32_t w_lo = w_lo0 + static_cast<uint32_t>(cc) * b_step;
#pragma unroll
                            for (int t = 0; t < kConvMaxTaps; ++t) {
                                if (t < taps) {
#pragma unroll
                                    for (int k = 0; k < NK; ++k)
                                        tc_mma_f16(tmem_d, dhi | (a_lo + toff[t] + 2u * k), dhi | (w_lo + 2u * k), idesc,
                                                   (t | k) != 0 ? 1u : acc0);
                                    w_lo += w_tap_step;
                                }
                            }
                            tc_commit_a(aempty0 + ab * 8);          // the chunk's MMAs have read the activation buffer
                            if (last_cc) tc_commit_a(tfull0 + as * 8);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int t = 0; t < kConvMaxTaps; ++t) {
                            if (t < taps) {
                                mbar_wait_a(full0 + stage * 8, phase, 300 + stage);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint32_t w_lo = w_lo0 + static_cast<uint32_t>(stage) * b_step;
#pragma unroll
                                    for (int k = 0; k < NK; ++k)
                                        tc_mma_f16(tmem_d, dhi | (a_lo + toff[t] + 2u * k), dhi | (w_lo + 2u * k), idesc,
                                                   (t | k) != 0 ? 1u : acc0);
                                    tc_commit_a(empty0 + stage * 8);
                                    if (t == taps - 1) {
                                        tc_commit_a(aempty0 + ab * 8);
                                        if (last_cc) tc_commit_a(tfull0 + as * 8);
                                    }
                                }
                                __syncwarp();
                                if (++stage == stages) { stage = 0; phase ^= 1u; }
                            }
                        }
                    }
                };
                const int nk = last_cc ? klast : kConvBK / 16;
                if (nk == 4) chunk(std::integral_constant<int, 4>{});
                else if (nk == 3) chunk(std::integral_constant<int, 3>{});
                else if (nk == 2) chunk(std::integral_constant<int, 2>{});
                else chunk(std::integral_constant<int, 1>{});
                if (++ab == na) { ab = 0; aph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        plain_epilogue(p, map_out, staging_base, tmem_base, tfull0, tempty0, total_tiles, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// Swapped-operand variant (Cout % 128 == 0, no residual): 128 output channels are the MMA M dimension (A = weight
// tile) and 256 pixels the N dimension (B = two independent 128-pixel activation patches).  On B200 this issues the
// same M128 x N256 x K16 instructions as the kernel above but measures 20-45 % faster, and it gives Cout = 128 layers
// the N = 256 instruction shape they cannot reach otherwise (a N = 128 instruction runs at half rate).
// The accumulator is transposed in TMEM (lane = channel, column = pixel): bias and GroupNorm partials are per-thread
// scalars, and the tile is transposed through the swizzled staging buffer on its way to the TMA store.
// Epilogue group g owns patch g of the pair: its own coordinates, staging buffer and named barrier.
// =====================================================================================================================
namespace {
struct PatchCoord {
    int ph, n, y0, x0;
    bool valid;
};
// patch `lp` (0 .. patches_per_phase) of phase ph: x fastest, then y, then sample
__device__ __forceinline__ PatchCoord decode_patch(const ConvParams& p, int ph, int lp) {
    PatchCoord c;
    c.ph = ph;
    c.valid = lp < p.patches_per_phase;
    const int tx = lp % p.tiles_x; lp /= p.tiles_x;
    const int ty = lp % p.tiles_y;
    c.n = lp / p.tiles_y;
    c.y0 = ty * p.BH;
    c.x0 = tx * p.BW;
    return c;
}
}  // namespace

// MC = true: launched as clusters of two CTAs that work on the SAME pair of pixel patches and two adjacent 128-channel
// tiles.  Each CTA loads its own weight tile; CTA r loads patch r of the pair ONCE and TMA multicasts it into both CTAs'
// shared memory, so the activation traffic from L2 per K step drops from 2 x 32 KB to 32 KB per cluster (48 -> 32 KB per
// CTA).  A stage is refilled only after BOTH CTAs' MMAs have consumed it: the empty barrier counts two arrivals and every
// MMA warp commits to the barrier of both CTAs (tcgen05.commit ... multicast::cluster).  Both CTAs walk identical item
// sequences (same pair, same taps, same split), so their pipelines stay in lockstep by construction.
template <bool MC>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_swapped_kernel(const ConvParams p, const __grid_constant__ CUtensorMap map_a0,
                          const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b,
                          const __grid_constant__ CUtensorMap map_out, int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float mask_s[2][128];   // per epilogue group: output mask of the patch's columns (0 outside the image)
    __shared__ int mask_ne_s[2][2];    // [group][tile parity]: some mask value of the patch differs from 1 (else the multiply is skipped)
    __shared__ int split_last_s;       // split-K: 1 when this CTA took the last ticket of its tile

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t kWBytes = 16384u, kPBytes = 16384u, kStageBytes = kWBytes + 2 * kPBytes;
    const int stages = p.stages;
    const int ctot = p.chunks0 + p.chunks1;
    const int ksteps_all = p.taps * ctot;
    const int ksplit = p.ksplit > 1 ? p.ksplit : 1;      // work item = (tile, K range); total_tiles counts items
    const int pairs_per_phase = (p.patches_per_phase + 1) >> 1;
    // work walked by this CTA: items blockIdx.x, +gridDim.x, ... of `total_tiles`; in cluster mode the walker is the
    // CLUSTER (index blockIdx.x / 2 of gridDim.x / 2) over `total_tiles` cluster items = (patch pair, channel-tile pair,
    // split), and CTA rank r takes channel tile 2q + r of the pair
    const int crank = MC ? static_cast<int>(cluster_ctarank()) : 0;
    const int walker = MC ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int walkers = MC ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int ntn = MC ? p.n_tiles_n >> 1 : p.n_tiles_n;      // channel tiles (MC: tile pairs) per patch pair
    const int tiles_per_phase = pairs_per_phase * ntn;
    const uint32_t full0 = smem_u32_pinned(&full_bar[0]), empty0 = smem_u32_pinned(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32_pinned(&tmem_full_bar[0]), tempty0 = smem_u32_pinned(&tmem_empty_bar[0]);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], MC ? 2 : 1);      // MC: the MMA warps of both CTAs release a stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kConvEpilogueThreads);
        }
        mask_ne_s[0][0] = mask_ne_s[0][1] = mask_ne_s[1][0] = mask_ne_s[1][1] = 0;
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
    if (MC) cluster_sync_all();   // the peer's barriers exist before anything is multicast to them
    pdl_wait_and_trigger();   // everything above is local set-up; below this line the predecessor's output is read

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int witem = walker; witem < total_tiles; witem += walkers) {
            const int wtile = witem / ksplit, sp = witem - wtile * ksplit;
            const int ph = wtile / tiles_per_phase;
            const int rem = wtile - ph * tiles_per_phase;
            const int nt = MC ? 2 * (rem % ntn) + crank : rem % ntn;
            const int pair = rem / ntn;
            const PatchCoord c0 = decode_patch(p, ph, 2 * pair), c1 = decode_patch(p, ph, 2 * pair + 1);
            const int bz = p.b_batch_mode == 1 ? ph : 0;
            const uint32_t tx_bytes = kWBytes + kPBytes + (c1.valid ? kPBytes : 0u);
            // K step ks = (tap t, source, 64-channel chunk cc) in the order the weights are packed: kcoord = ks * 64
            const int k0 = sp * ksteps_all / ksplit, k1 = (sp + 1) * ksteps_all / ksplit;
            int t = k0 / ctot, r = k0 - t * ctot;
            for (int ks = k0; ks < k1; ++ks) {
                const ConvTap tap = p.tap[ph * p.taps + t];
                const bool src1 = r >= p.chunks0;
                const CUtensorMap* ma = src1 ? &map_a1 : &map_a0;
                const int cc = src1 ? r - p.chunks0 : r;
                mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 100 + stage);
                if (elect_one()) {
                    const uint32_t sw = tiles_base + stage * kStageBytes;
                    const uint32_t fb = full0 + stage * 8;
                    const int cch = tap.c + cc * kConvBK;
                    mbar_arrive_expect_tx_a(fb, tx_bytes);      // (MC: includes the patch the peer multicasts to this CTA)
                    tma_load_3d_a(sw, &map_b, fb, ks * kConvBK, nt * 128, bz);
                    if (!MC) {
                        tma_load_5d_a(sw + kWBytes, ma, fb, cch, c0.x0 + tap.dx, tap.p, c0.y0 + tap.dy, c0.n);
                        if (c1.valid)
                            tma_load_5d_a(sw + kWBytes + kPBytes, ma, fb, cch, c1.x0 + tap.dx, tap.p, c1.y0 + tap.dy, c1.n);
                    } else if (crank == 0) {
                        tma_load_5d_mc_a(sw + kWBytes, ma, fb, cch, c0.x0 + tap.dx, tap.p, c0.y0 + tap.dy, c0.n, 3);
                    } else if (c1.valid) {
                        tma_load_5d_mc_a(sw + kWBytes + kPBytes, ma, fb, cch, c1.x0 + tap.dx, tap.p, c1.y0 + tap.dy, c1.n, 3);
                    }
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
                if (++r == ctot) { r = 0; ++t; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer: D[ch][px] += W[ch][k] * X[px][k]^T
        const uint32_t idesc = umma_idesc_f16(256u);
        const uint32_t desc_hi = static_cast<uint32_t>(umma_desc_sw128(0) >> 32);
        const uint32_t w_lo0 = (tiles_base & 0x3FFFFu) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int witem = walker; witem < total_tiles; witem += walkers, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int sp = witem % ksplit;
            const int ksteps = (sp + 1) * ksteps_all / ksplit - sp * ksteps_all / ksplit;
            mbar_wait_a(tempty0 + as * 8, aphase ^ 1u, 200 + as);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * 256u;
            for (int ks = 0; ks < ksteps; ++ks) {
                mbar_wait_a(full0 + stage * 8, phase, 300 + stage);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t w_lo = w_lo0 + stage * (kStageBytes >> 4);
                    const uint32_t x_lo = w_lo + (kWBytes >> 4);
#pragma unroll
                    for (int k = 0; k < kConvBK / 16; ++k) {
                        const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (w_lo + 2u * k);
                        const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (x_lo + 2u * k);
                        tc_mma_f16(tmem_d, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    if (MC) tc_commit_mc_a(empty0 + stage * 8, 3);      // frees the slot in both CTAs of the cluster
                    else tc_commit_a(empty0 + stage * 8);
                    if (ks == ksteps - 1) tc_commit_a(tfull0 + as * 8);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue: thread = output channel, columns = pixels
        const int ew = warp & 3;
        const int grp = (warp - 4) >> 2;          // which patch of the pair
        const int c = ew * 32 + lane;             // output channel within the 128-channel tile
        const int bw_shift = 31 - __clz(p.BW);
        const int cpg = p.stats ? p.Cout / p.groups : 16;
        // per-warp staging: 64 pixel rows x 32 channels (64 B rows, 64-byte swizzle) = 4 KB, stored by the warp's own
        // TMA (box {32 ch, 64 px}), so the epilogue needs no cross-warp barrier
        const uint32_t warp_buf = tiles_base + stages * kStageBytes + static_cast<uint32_t>(grp * 4 + ew) * 4096u;
        // element (row j, channel lane): byte = j*64 + ((lane>>3) ^ ((j>>1)&3))*16 + (lane&7)*2; the XOR term only depends
        // on (j>>1)&3, so four per-thread bases + a compile-time row offset address every element
        uint32_t sbase[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            sbase[q] = warp_buf + ((static_cast<uint32_t>(lane >> 3) ^ static_cast<uint32_t>(q)) << 4) +
                       static_cast<uint32_t>(lane & 7) * 2u;
        float amax = 0.f;                         // largest |value| packed by this thread (saturation report)
        int it = 0;
        for (int witem = walker; witem < total_tiles; witem += walkers, ++it) {
            const int wtile = witem / ksplit;
            const int ph = wtile / tiles_per_phase;
            const int rem = wtile - ph * tiles_per_phase;
            const int nt = MC ? 2 * (rem % ntn) + crank : rem % ntn;
            const PatchCoord pc = decode_patch(p, ph, 2 * (rem / ntn) + grp);
            // per-CTA output tile / work item ids (split-K workspace and tickets): unique across the cluster's two CTAs
            const int tile = MC ? wtile * 2 + crank : wtile;
            const int item = tile * ksplit + (witem - wtile * ksplit);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait_a(tfull0 + as * 8, aphase, 400 + as);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as) * 256u;
            if (ksplit > 1) {
                // ---- split-K: park this item's fp32 partial tile ([px][ch]: a warp stores 128 contiguous bytes per pixel),
                // take a ticket; only the CTA holding the last ticket goes on to reduce + epilogue
                if (pc.valid) {
                    float* part = p.kpart + ((static_cast<long long>(item) * 2 + grp) * 128) * 128 + c;
                    for (int hc = 0; hc < 2; ++hc) {
                        uint32_t v0[32], v1[32];
                        tmem_ld_32x32(taddr + grp * 128 + hc * 64, v0);
                        tmem_ld_32x32(taddr + grp * 128 + hc * 64 + 32, v1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 64; ++j)
                            __stcg(part + (hc * 64 + j) * 128, __uint_as_float(j < 32 ? v0[j] : v1[j - 32]));
                    }
                }
                tc_fence_before();
                mbar_arrive_a(tempty0 + as * 8);
                __threadfence();
                named_bar_sync(3, kConvEpilogueThreads);
                if (threadIdx.x == 4 * 32) split_last_s = atomicAdd(p.ktick + tile, 1) == ksplit - 1 ? 1 : 0;
                named_bar_sync(3, kConvEpilogueThreads);
                const bool last = split_last_s != 0;
                named_bar_sync(3, kConvEpilogueThreads);      // split_last_s may be rewritten by the next item
                if (!last) continue;
                __threadfence();
                if (threadIdx.x == 4 * 32) p.ktick[tile] = 0;  // ready for the next launch
                if (!pc.valid) continue;
            } else if (!pc.valid || (p.dbg_flags & 2)) {   // odd patch count: the second half of the last pair holds no image
                tc_fence_before();
                mbar_arrive_a(tempty0 + as * 8);
                continue;
            }
            const int wlim = p.Wm - pc.x0;        // pixels with tx >= wlim lie outside the image
            const int hlim = p.Hm - pc.y0;        // rows >= hlim too (patch heights need not divide the image height)
            const int cg = nt * 128 + c;          // global output channel
            const float bias = p.bias ? __ldg(p.bias + cg) : 0.f;
            long long* stats_n = p.stats ? p.stats + static_cast<long long>(pc.n) * p.groups * 2 : nullptr;
            const float* mrow = p.mask ? p.mask + static_cast<long long>(pc.n) * p.mask_stride + p.ox_off[ph] : nullptr;
            float s = 0.f, ss = 0.f;
            const bool full = wlim >= p.BW && hlim >= p.BH;   // every pixel of the patch lies inside the image
            bool domask = false;
            if (mrow) {
                // the mask is 1 almost everywhere (it only zeroes the frames past an utterance's length): the per-element
                // multiply -- and its shared-memory load -- is skipped for patches whose columns are all unmasked
                const int et = lane + ew * 32;    // any 128 threads of the group cover BW <= 128 columns
                if (et < p.BW) {
                    const float mv = et < wlim ? __ldg(mrow + (pc.x0 + et) * p.ox_mul) : 0.f;
                    mask_s[grp][et] = mv;
                    if (mv != 1.f && et < wlim) mask_ne_s[grp][it & 1] = 1;
                }
                named_bar_sync(1 + grp, 128);
                domask = mask_ne_s[grp][it & 1] != 0;
                if (et == 0) mask_ne_s[grp][(it + 1) & 1] = 0;    // nobody touches the other parity's flag during this tile
            }
            for (int hc = 0; hc < 2; ++hc) {
                const int pb = hc * 64;           // first pixel of this 64-pixel sub-block inside the patch
                uint32_t v0[32], v1[32];
                if (ksplit > 1) {
                    // sum of the partial tiles in split order 0 .. ksplit-1 (fixed order: bit-reproducible); 64 independent
                    // L2 loads are in flight per thread and split (a per-element dependent chain would serialise latencies)
                    const float* part = p.kpart + ((static_cast<long long>(tile) * ksplit * 2 + grp) * 128 + pb) * 128 + c;
                    float a[64];
#pragma unroll
                    for (int j = 0; j < 64; ++j) a[j] = __ldcg(part + j * 128);
                    for (int sp = 1; sp < ksplit; ++sp) {
                        const float* ps = part + static_cast<long long>(sp) * 2 * 128 * 128;
                        float b[64];
#pragma unroll
                        for (int j = 0; j < 64; ++j) b[j] = __ldcg(ps + j * 128);
#pragma unroll
                        for (int j = 0; j < 64; ++j) a[j] += b[j];
                    }
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        if (j < 32) v0[j] = __float_as_uint(a[j]);
                        else v1[j - 32] = __float_as_uint(a[j]);
                    }
                } else {
                    tmem_ld_32x32(taddr + grp * 128 + pb, v0);
                    tmem_ld_32x32(taddr + grp * 128 + pb + 32, v1);
                    tmem_ld_wait();
                    if (hc == 1) {
                        tc_fence_before();
                        mbar_arrive_a(tempty0 + as * 8);
                    }
                }
                if (lane == 0) tma_store_wait_read<0>();   // the warp's staging buffer has been read by its last store
                __syncwarp();
                // two pixels per iteration: one packed convert, both halves stored (the statistics keep the pixel order);
                // instantiated with and without the mask multiply (a predicate would still cost the issue slots)
                auto emit_half = [&](auto masked) {
#pragma unroll
                    for (int j = 0; j < 64; j += 2) {
                        float f0 = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]) + bias;
                        float f1 = __uint_as_float(j < 32 ? v0[j + 1] : v1[j - 31]) + bias;
                        if (stats_n) {
                            if (full || (((pb + j) & (p.BW - 1)) < wlim && ((pb + j) >> bw_shift) < hlim)) {
                                s += f0;
                                ss = fmaf(f0, f0, ss);
                            }
                            if (full || (((pb + j + 1) & (p.BW - 1)) < wlim && ((pb + j + 1) >> bw_shift) < hlim)) {
                                s += f1;
                                ss = fmaf(f1, f1, ss);
                            }
                        }
                        if (decltype(masked)::value) {
                            f0 *= mask_s[grp][(pb + j) & (p.BW - 1)];
                            f1 *= mask_s[grp][(pb + j + 1) & (p.BW - 1)];
                        }
                        amax = fmaxf(amax, fmaxf(fabsf(f0), fabsf(f1)));
                        sts_f16_pair(sbase[(j >> 1) & 3] + j * 64, sbase[(j >> 1) & 3] + (j + 1) * 64, pack_f16x2_sat(f0, f1));
                    }
                };
                if (domask) emit_half(std::true_type{});
                else emit_half(std::false_type{});
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const int yb = pc.y0 + (pb >> bw_shift);
                    const int xb = pc.x0 + (pb & (p.BW - 1));
                    const int cph = p.ox_off[ph] * p.out_c_phase_mul + nt * 128 + ew * 32;
                    tma_store_5d_a(&map_out, warp_buf, cph, xb, p.oy_off[ph], yb, pc.n);
                    tma_store_commit();
                }
            }
            if (mrow) named_bar_sync(1 + grp, 128);   // mask_s is rewritten by the next tile
            if (stats_n) {
                // channels of one GroupNorm group are adjacent lanes; a warp holds 32 channels = 32/cpg groups
                // (or part of one group when cpg >= 32)
                const int span = cpg < 32 ? cpg : 32;
                for (int o = span >> 1; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    ss += __shfl_xor_sync(0xffffffffu, ss, o);
                }
                if ((lane & (span - 1)) == 0) {
                    long long* dst = stats_n + (cg / cpg) * 2;
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst),
                              static_cast<unsigned long long>(__float2ll_rn(s * kStatSumScale)));
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst + 1),
                              static_cast<unsigned long long>(__float2ll_rn(ss * kStatSqScale)));
                }
            }
        }
        if (p.sat && amax >= 65504.f) atomicAdd(p.sat, 1ull);
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
    if (MC) cluster_sync_all();   // the peer's last commits arrive on this CTA's barriers: do not exit before it is done
}

// =====================================================================================================================
// Halo-reuse variant of the swapped-operand kernel for 3x3 stride-1 convs (Hm % 8 == 0, Cout % 128 == 0).
// A tile is 8 image rows x 32 columns (256 pixels on the MMA N axis, column index = x*8 + y).  For every 64-channel
// K chunk the activation tile is loaded ONCE with its one-pixel halo -- a TMA box {64 ch, HY rows, 34 columns} of a
// (c, y, x, n) view, so shared-memory rows are ordered column by column, HY rows per column -- and the nine taps are
// nine UMMA descriptors into that buffer: start row dx*HY + dy, 8-row groups HY rows apart (one group = the 8 rows
// of one image column).  Activation traffic from L2 drops from 9 x 32 KB to 43.5 KB (HY = 10) per chunk; only the
// 16 KB weight tiles still stream per tap through the stage ring.
// =====================================================================================================================
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_halo_kernel(const ConvParams p, const __grid_constant__ CUtensorMap map_a0,
                       const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];        // weight stages
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t hfull_bar[2];       // halo buffers
    __shared__ __align__(8) uint64_t hempty_bar[2];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float mask_s[2][16];    // per epilogue group: output mask of its 16 columns (0 outside the image)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr uint32_t kWBytes = 16384u;
    const uint32_t hy = static_cast<uint32_t>(p.halo_hy);
    const uint32_t halo_tx = 34u * hy * 128u;                      // bytes one halo load delivers
    const uint32_t halo_bytes = (halo_tx + 1023u) & ~1023u;        // buffer pitch (1024-aligned for the swizzle)
    const uint32_t w_base = tiles_base + 2u * halo_bytes;
    const int stages = p.stages;
    const int chunks = p.chunks0 + p.chunks1;
    const uint32_t full0 = smem_u32_pinned(&full_bar[0]), empty0 = smem_u32_pinned(&empty_bar[0]);
    const uint32_t hfull0 = smem_u32_pinned(&hfull_bar[0]), hempty0 = smem_u32_pinned(&hempty_bar[0]);
    const uint32_t tfull0 = smem_u32_pinned(&tmem_full_bar[0]), tempty0 = smem_u32_pinned(&tmem_empty_bar[0]);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
        tma_prefetch_desc(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hfull_bar[i], 1);
            mbar_init(&hempty_bar[i], 1);
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kConvEpilogueThreads);
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
    pdl_wait_and_trigger();   // everything above is local set-up; below this line the predecessor's output is read

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer
        int stage = 0, hb = 0;
        uint32_t phase = 0, hphase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles_n;
            int r = tile / p.n_tiles_n;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y;
            const int n = r / p.tiles_y;
            for (int cc = 0; cc < chunks; ++cc) {
                const bool src1 = cc >= p.chunks0;
                mbar_wait_a(hempty0 + hb * 8, hphase ^ 1u, 120 + hb);
                if (elect_one()) {
                    const uint32_t fb = hfull0 + hb * 8;
                    mbar_arrive_expect_tx_a(fb, halo_tx);
                    tma_load_4d_a(tiles_base + hb * halo_bytes, src1 ? &map_a1 : &map_a0, fb,
                                  (src1 ? cc - p.chunks0 : cc) * kConvBK, ty * 8 - 1, tx * 32 - 1, n);
                }
                __syncwarp();
                if (++hb == 2) { hb = 0; hphase ^= 1u; }
                for (int t = 0; t < 9; ++t) {
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 100 + stage);
                    if (elect_one()) {
                        const uint32_t fb = full0 + stage * 8;
                        mbar_arrive_expect_tx_a(fb, kWBytes);
                        tma_load_3d_a(w_base + stage * kWBytes, &map_b, fb, (t * chunks + cc) * kConvBK, nt * 128, 0);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer: D[ch][px] += W[ch][k] * X[px][k]^T
        const uint32_t idesc = umma_idesc_f16(256u);
        const uint32_t desc_hi_w = static_cast<uint32_t>(umma_desc_sw128(0) >> 32);
        int stage = 0, hb = 0;
        uint32_t phase = 0, hphase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait_a(tempty0 + as * 8, aphase ^ 1u, 200 + as);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * 256u;
            for (int cc = 0; cc < chunks; ++cc) {
                mbar_wait_a(hfull0 + hb * 8, hphase, 320 + hb);
                const uint32_t hbase = tiles_base + hb * halo_bytes;
                for (int t = 0; t < 9; ++t) {
                    mbar_wait_a(full0 + stage * 8, phase, 300 + stage);
                    tc_fence_after();
                    if (elect_one()) {
                        const int dy = t / 3, dx = t - dy * 3;      // tap (kh, kw): input pixel (y + kh - 1, x + kw - 1)
                        const uint32_t xaddr = hbase + (static_cast<uint32_t>(dx) * hy + dy) * 128u;
                        const uint64_t db0 = umma_desc_sw128_ex(xaddr, hy * 128u, p.halo_boff ? (xaddr >> 7) : 0u);
                        const uint32_t w_lo = ((w_base + stage * kWBytes) & 0x3FFFFu) >> 4;
#pragma unroll
                        for (int k = 0; k < kConvBK / 16; ++k) {
                            const uint64_t da = (static_cast<uint64_t>(desc_hi_w) << 32) | (w_lo + 2u * k);
                            tc_mma_f16(tmem_d, da, db0 + 2u * k, idesc, (cc | t | k) != 0 ? 1u : 0u);
                        }
                        tc_commit_a(empty0 + stage * 8);
                        if (t == 8) {
                            tc_commit_a(hempty0 + hb * 8);
                            if (cc == chunks - 1) tc_commit_a(tfull0 + as * 8);
                        }
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
                if (++hb == 2) { hb = 0; hphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue: thread = output channel, columns = pixels
        const int ew = warp & 3;
        const int grp = (warp - 4) >> 2;          // columns [128*grp, 128*grp + 128) = image columns 16*grp .. 16*grp+15
        const int c = ew * 32 + lane;
        const int cpg = p.stats ? p.Cout / p.groups : 16;
        // per-warp staging: 32 pixel rows x 32 channels (64 B rows, 64-byte swizzle) = 2 KB, stored by the warp's own TMA
        // (box {32 ch, 8 rows, 4 columns})
        const uint32_t warp_buf = w_base + stages * kWBytes + static_cast<uint32_t>(grp * 4 + ew) * 2048u;
        uint32_t sbase[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            sbase[q] = warp_buf + ((static_cast<uint32_t>(lane >> 3) ^ static_cast<uint32_t>(q)) << 4) +
                       static_cast<uint32_t>(lane & 7) * 2u;
        float amax = 0.f;                         // largest |value| packed by this thread (saturation report)
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int nt = tile % p.n_tiles_n;
            int r = tile / p.n_tiles_n;
            const int tx = r % p.tiles_x; r /= p.tiles_x;
            const int ty = r % p.tiles_y;
            const int n = r / p.tiles_y;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int x0 = tx * 32 + grp * 16;    // first image column of this group
            mbar_wait_a(tfull0 + as * 8, aphase, 400 + as);
            tc_fence_after();
            const int wlim = p.Wm - x0;           // columns with index >= wlim lie outside the image
            if (wlim <= 0 || (p.dbg_flags & 2)) {
                tc_fence_before();
                mbar_arrive_a(tempty0 + as * 8);
                continue;
            }
            const int cg = nt * 128 + c;
            const float bias = p.bias ? __ldg(p.bias + cg) : 0.f;
            long long* stats_n = p.stats ? p.stats + static_cast<long long>(n) * p.groups * 2 : nullptr;
            const float* mrow = p.mask ? p.mask + static_cast<long long>(n) * p.mask_stride : nullptr;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as) * 256u;
            float s = 0.f, ss = 0.f;
            const bool full = wlim >= 16;
            if (mrow) {
                const int et = lane + ew * 32;
                if (et < 16) mask_s[grp][et] = et < wlim ? __ldg(mrow + x0 + et) : 0.f;
                named_bar_sync(1 + grp, 128);
            }
            for (int hc = 0; hc < 4; ++hc) {
                // 32 accumulator columns = image columns 4*hc .. 4*hc+3 of the group, 8 rows each
                if (hc * 4 >= wlim) {             // the rest of the group lies right of the image
                    if (hc == 3) {
                        tc_fence_before();
                        mbar_arrive_a(tempty0 + as * 8);
                    }
                    continue;
                }
                uint32_t v0[32];
                tmem_ld_32x32(taddr + grp * 128 + hc * 32, v0);
                tmem_ld_wait();
                if (hc == 3) {
                    tc_fence_before();
                    mbar_arrive_a(tempty0 + as * 8);
                }
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                // One element at a time, each store a volatile asm WITH the "memory" clobber: this kernel runs at the shared-
                // memory bandwidth limit (UMMA operand reads ~96 B/clk + TMA fills ~40 B/clk of the 128 B/clk), and the
                // faster paired-store loop of the swapped kernel (sts_f16_pair) issues its stores in bursts that take
                // cycles from the tensor pipe: measured 82.4 % instead of 87.8 % tensor-pipe active on the level-0 layers.
                // The clobber spreads the stores over the tile's MMA time, which is what this kernel wants.
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float f = __uint_as_float(v0[j]) + bias;
                    const int xl = hc * 4 + (j >> 3);
                    if (stats_n) {
                        if (full || xl < wlim) {
                            s += f;
                            ss = fmaf(f, f, ss);
                        }
                    }
                    if (mrow) f *= mask_s[grp][xl];
                    amax = fmaxf(amax, fabsf(f));
                    unsigned short hbits;
                    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hbits) : "f"(f));
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(sbase[(j >> 1) & 3] + j * 64), "h"(hbits) : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d_a(&map_out, warp_buf, nt * 128 + ew * 32, ty * 8, x0 + hc * 4, n);
                    tma_store_commit();
                }
            }
            if (mrow) named_bar_sync(1 + grp, 128);
            if (stats_n) {
                const int span = cpg < 32 ? cpg : 32;
                for (int o = span >> 1; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o);
                    ss += __shfl_xor_sync(0xffffffffu, ss, o);
                }
                if ((lane & (span - 1)) == 0) {
                    long long* dst = stats_n + (cg / cpg) * 2;
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst),
                              static_cast<unsigned long long>(__float2ll_rn(s * kStatSumScale)));
                    atomicAdd(reinterpret_cast<unsigned long long*>(dst + 1),
                              static_cast<unsigned long long>(__float2ll_rn(ss * kStatSqScale)));
                }
            }
        }
        if (p.sat && amax >= 65504.f) atomicAdd(p.sat, 1ull);
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int launch_conv_igemm(const ConvParams& p, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                      const CUtensorMap& out, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(conv_igemm_swapped_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(conv_igemm_swapped_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(conv_igemm_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(conv1d_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaFuncSetAttribute(conv1d_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBytes);
        if (e != cudaSuccess) return static_cast<int>(e);
        attr_set = true;
    }
    long long total = static_cast<long long>(p.phases) * p.N * p.tiles_y * p.tiles_x * p.n_tiles_n;
    if (p.halo) {
        if (total <= 0 || total > 0x7fffffffLL) return static_cast<int>(cudaErrorInvalidValue);
        const size_t halo_bytes = (static_cast<size_t>(34) * p.halo_hy * 128 + 1023) / 1024 * 1024;
        const size_t smem = 1024 + 2 * halo_bytes + static_cast<size_t>(p.stages) * 16384 + 16384;
        if (smem > static_cast<size_t>(kConvSmemBytes) || p.stages > 8) return static_cast<int>(cudaErrorInvalidValue);
        const int grid = static_cast<int>(total < num_sms ? total : num_sms);
        return static_cast<int>(launch_k(conv_igemm_halo_kernel, dim3(grid), dim3(kConvThreads), smem, stream, p, a0, a1, b, out,
                                         static_cast<int>(total)));
    }
    if (p.swap_ab) total = static_cast<long long>(p.phases) * ((p.patches_per_phase + 1) / 2) * p.n_tiles_n * (p.ksplit > 1 ? p.ksplit : 1);
    if (total <= 0 || total > 0x7fffffffLL) return static_cast<int>(cudaErrorInvalidValue);
    const int grid = static_cast<int>(total < num_sms ? total : num_sms);
    if (p.swap_ab) {
        const size_t smem = 1024 + static_cast<size_t>(p.stages) * 49152 + 2 * 16384;
        if (smem > static_cast<size_t>(kConvSmemBytes)) return static_cast<int>(cudaErrorInvalidValue);
        // cluster mode (see the kernel; USB_MC=1): an even number of 128-channel tiles, at least two clusters' worth of work.
        // OFF by default: measured on B200 it is neutral in throughput mode (8 979 vs 8 966 frames/s on the 32 x 1000 shard,
        // conv class 2 719 vs 2 721 ms per pass) and 2 % slower for one-utterance calls (84.9 vs 83.2 ms): these launches are
        // not bound by L2 -> shared-memory traffic (tensor pipe 86-89 % active at Be = 96; per-CTA TMA latency x 4 stages at
        // Be = 3), so loading the activation patch once per cluster changes nothing.  tests/test_gpu_ops.py runs it.
        static const bool use_mc = getenv("USB_MC") != nullptr;
        if (use_mc && p.n_tiles_n % 2 == 0 && total >= 4) {
            const long long citems = total / 2;
            // clusters that can be resident at once (GPC boundaries may leave a few SMs unpaired): queried once per
            // shared-memory size, so that the persistent walk never waits for a cluster slot
            static int max_clusters[2] = {0, 0};
            int& mc = max_clusters[p.ksplit > 1 ? 1 : 0];
            if (mc == 0) {
                cudaLaunchConfig_t qc = {};
                qc.gridDim = dim3(num_sms & ~1);
                qc.blockDim = dim3(kConvThreads);
                qc.dynamicSmemBytes = smem;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 2;
                qa[0].val.clusterDim.y = 1;
                qa[0].val.clusterDim.z = 1;
                qc.attrs = qa;
                qc.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, conv_igemm_swapped_kernel<true>, &qc) != cudaSuccess || n < 1) {
                    cudaGetLastError();
                    n = num_sms / 2;
                }
                mc = n < num_sms / 2 ? n : num_sms / 2;
            }
            int clusters = mc;
            if (citems < clusters) clusters = static_cast<int>(citems);
            return static_cast<int>(launch_k_cluster(conv_igemm_swapped_kernel<true>, dim3(2 * clusters), dim3(kConvThreads), smem,
                                                     stream, 2, p, a0, a1, b, out, static_cast<int>(citems)));
        }
        return static_cast<int>(launch_k(conv_igemm_swapped_kernel<false>, dim3(grid), dim3(kConvThreads), smem, stream, p, a0, a1, b,
                                         out, static_cast<int>(total)));
    }
    if (p.h1d) {
        const size_t a_stride = (static_cast<size_t>(p.h1d_rows) * 128 + 1023) / 1024 * 1024;
        const size_t wtiles = p.h1d_wres ? static_cast<size_t>(p.taps) * p.chunks0 : static_cast<size_t>(p.stages);
        const size_t smem = 1024 + p.h1d_na * a_stride + wtiles * p.BN * 128 + 8 * 8192;
        if (smem > static_cast<size_t>(kConvSmemBytes) || p.h1d_na < 1 || p.h1d_na > 4 || p.stages > 8 || p.h1d_klast < 1 || p.h1d_klast > 4 ||
            (p.h1d_wres && (p.n_tiles_n != 1 || wtiles * p.BN * 128 >= (1u << 20))))
            return static_cast<int>(cudaErrorInvalidValue);
        if (p.h1d_wres)
            return static_cast<int>(launch_k(conv1d_halo_kernel<true>, dim3(grid), dim3(kConvThreads), smem, stream, p, a1, b, out,
                                             static_cast<int>(total)));
        return static_cast<int>(launch_k(conv1d_halo_kernel<false>, dim3(grid), dim3(kConvThreads), smem, stream, p, a1, b, out,
                                         static_cast<int>(total)));
    }
    const size_t smem = 1024 + static_cast<size_t>(p.stages) * (16384 + static_cast<size_t>(p.BN) * 128) + 8 * 8192;
    if (smem > static_cast<size_t>(kConvSmemBytes)) return static_cast<int>(cudaErrorInvalidValue);
    return static_cast<int>(launch_k(conv_igemm_kernel, dim3(grid), dim3(kConvThreads), smem, stream, p, a0, a1, b, out,
                                     static_cast<int>(total)));
}

}  // namespace usb
