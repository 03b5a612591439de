// Convolution weight gradient on tcgen05 (see train.h).  Reference op: the weight gradients autograd computes for
// Conv2d / ConvTranspose2d in loss.backward() (finetune.py:163) of the U-Net convs (unitspeech/unitspeech.py:21,30,49,66,83).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warp 2 = TMEM allocator,
// warps 4-7 = epilogue (TMEM -> fp32 stores / reductions into dW).  Persistent: one CTA per SM walks the work items,
// with two TMEM accumulator stages so that an item's epilogue overlaps the next item's main loop.
// Shared-memory stage: (2 + n_tile/64) slabs of 16 KB; a slab is [kp <= 128 pixels][64 channels] fp16 with 128-byte rows
// in the 128-byte swizzle.  UMMA reads it MN-major: 8-pixel groups 1024 B apart (SBO), 64-channel slabs 16 KB apart
// (LBO); one instruction covers 16 pixels (K = 16), so the descriptor start address advances by 2048 B per K step.
#include "ptx.cuh"
#include "train.h"

namespace usb {

namespace {
constexpr uint32_t kSlabBytes = 16384u;

// MN-major operand, SWIZZLE_128B: canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) in fp16 elements
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(kSlabBytes >> 4) << 16;    // leading byte offset: next 64-channel slab
    d |= static_cast<uint64_t>(1024 >> 4) << 32;          // stride byte offset: next 8-pixel group
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// fp16 operands, fp32 accumulate, A and B MN-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_f16_mn(uint32_t n) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
}  // namespace

// One work item = one (tap, 128 x n_tile output tile, pixel-chunk range).  The kernel is persistent: CTA b walks items
// b, b + gridDim.x, ...; two TMEM accumulator stages let the epilogue of one item overlap the main loop of the next.
struct WgradItem {
    int tap, m0, n0, seg, c_begin, c_end;
};
__device__ __forceinline__ WgradItem wgrad_item(const WgradTcParams& p, int item) {
    WgradItem w;
    const int tiles = p.taps * p.tiles_m * p.tiles_n;
    int bx = item % tiles;
    const int by = item / tiles;
    const int tn = bx % p.tiles_n; bx /= p.tiles_n;
    const int tm = bx % p.tiles_m;
    w.tap = bx / p.tiles_m;
    w.m0 = tm * 128;
    w.n0 = tn * p.n_tile;
    w.seg = by / p.ksplit;
    const int part = by % p.ksplit;
    const int chunks_per_sample = p.tiles_y * p.tiles_x;
    const int seg_chunks = p.s_n != 0 ? chunks_per_sample : p.N * chunks_per_sample;
    const int per = (seg_chunks + p.ksplit - 1) / p.ksplit;
    w.c_begin = w.seg * seg_chunks + part * per;
    w.c_end = min(w.c_begin + per, (w.seg + 1) * seg_chunks);
    return w;
}

__global__ void __launch_bounds__(256, 1)
wgrad_tc_kernel(const WgradTcParams p, const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                int total_items) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[4];
    __shared__ __align__(8) uint64_t empty_bar[4];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = p.n_tile >> 6;                       // B slabs
    const uint32_t stage_bytes = (2u + nb) * kSlabBytes;
    const uint32_t tiles_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int stages = p.stages;
    const int kp = p.BH * p.BW;
    const uint32_t full0 = smem_u32_pinned(&full_bar[0]), empty0 = smem_u32_pinned(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32_pinned(&tmem_full_bar[0]), tempty0 = smem_u32_pinned(&tmem_empty_bar[0]);
    const uint32_t tmem_cols = 2u * static_cast<uint32_t>(p.n_tile);   // 256 or 512 (powers of two)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
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
        tmem_alloc(&tmem_base_smem, tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ---------------------------------------------------- TMA producer
        const uint32_t tx_bytes = (2u + nb) * static_cast<uint32_t>(kp) * 128u;
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const WgradItem w = wgrad_item(p, item);
            const ConvTap ta = p.atap[w.tap], tb = p.btap[w.tap];
            for (int c = w.c_begin; c < w.c_end; ++c) {
                int r = c;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int ty = r % p.tiles_y;
                const int n = r / p.tiles_y;
                const int y0 = ty * p.BH, x0 = tx * p.BW;
                mbar_wait_a(empty0 + stage * 8, phase ^ 1u, 500 + stage);
                if (elect_one()) {
                    const uint32_t sa = tiles_base + stage * stage_bytes;
                    const uint32_t fb = full0 + stage * 8;
                    mbar_arrive_expect_tx_a(fb, tx_bytes);
                    for (int s = 0; s < 2; ++s)
                        tma_load_5d_a(sa + s * kSlabBytes, &map_a, fb, w.m0 + s * 64 + ta.c, x0 + ta.dx, ta.p, y0 + ta.dy, n);
                    for (int s = 0; s < nb; ++s)
                        tma_load_5d_a(sa + (2 + s) * kSlabBytes, &map_b, fb, w.n0 + s * 64 + tb.c, x0 + tb.dx, tb.p, y0 + tb.dy, n);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------- MMA issuer
        const uint32_t idesc = umma_idesc_f16_mn(static_cast<uint32_t>(p.n_tile));
        const int ksteps = kp >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;   // non-empty items seen so far: selects the accumulator stage
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const WgradItem w = wgrad_item(p, item);
            if (w.c_begin >= w.c_end) continue;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ++it;
            mbar_wait_a(tempty0 + as * 8, aphase ^ 1u, 650 + as);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as) * static_cast<uint32_t>(p.n_tile);
            for (int c = w.c_begin; c < w.c_end; ++c) {
                mbar_wait_a(full0 + stage * 8, phase, 600 + stage);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = tiles_base + stage * stage_bytes;
                    const uint32_t sb = sa + 2u * kSlabBytes;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t da = umma_desc_mn_sw128(sa + k * 2048u);
                        const uint64_t db = umma_desc_mn_sw128(sb + k * 2048u);
                        tc_mma_f16(tmem_d, da, db, idesc, (c != w.c_begin || k != 0) ? 1u : 0u);
                    }
                    tc_commit_a(empty0 + stage * 8);
                    if (c == w.c_end - 1) tc_commit_a(tfull0 + as * 8);
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------- epilogue: thread = output channel row of the tile
        const int ew = warp & 3;
        int it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
            const WgradItem w = wgrad_item(p, item);
            if (w.c_begin >= w.c_end) continue;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ++it;
            mbar_wait_a(tfull0 + as * 8, aphase, 700 + as);
            tc_fence_after();
            const int co = w.m0 + ew * 32 + lane;
            const int n0 = w.n0;
            float* out = p.dW + p.tap_off[w.tap] + (p.s_n != 0 ? w.seg * p.s_n : 0) + co * p.s_co;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as) * static_cast<uint32_t>(p.n_tile);
            for (int cb = 0; cb < p.n_tile; cb += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + cb, v);
                tmem_ld_wait();
                if (cb + 32 >= p.n_tile) {   // the accumulator stage has been read out: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive_a(tempty0 + as * 8);
                }
                if (co < p.Cout) {
                    if (p.s_ci == 1 && n0 + cb + 32 <= p.Cin && p.ksplit == 1 && p.overwrite) {
                        // the only CTA of this tile and a destination known to be zero: plain 16-byte stores (fp32
                        // reductions resolve in L2 at a fraction of the store rate; a level-3 layer writes 38 MB of gradient)
                        float4* o = reinterpret_cast<float4*>(out + n0 + cb);
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            o[j >> 2] = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                    __uint_as_float(v[j + 3]));
                    } else if (p.s_ci == 1 && n0 + cb + 32 <= p.Cin) {   // contiguous input channels: 16-byte vector reductions
                        float* o = out + n0 + cb;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(v[j])),
                                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                         "f"(__uint_as_float(v[j + 3]))
                                         : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int ci = n0 + cb + j;
                            if (ci < p.Cin) atomicAdd(out + ci * p.s_ci, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

int launch_wgrad_tc(WgradTcParams& p, const CUtensorMap& map_a, const CUtensorMap& map_b, int num_sms, cudaStream_t s) {
    const int kp = p.BH * p.BW;
    if (kp % 16 || kp < 16 || kp > 128 || (p.n_tile != 128 && p.n_tile != 256) || p.taps < 1 || p.taps > kWgradMaxTaps)
        return (int)cudaErrorInvalidValue;
    const int nb = p.n_tile / 64;
    const size_t stage_bytes = static_cast<size_t>(2 + nb) * kSlabBytes;
    p.stages = static_cast<int>((200 * 1024) / stage_bytes);
    if (p.stages > 4) p.stages = 4;
    const size_t smem = p.stages * stage_bytes + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    const long long tiles = static_cast<long long>(p.taps) * p.tiles_m * p.tiles_n;
    const int nseg = p.s_n != 0 ? p.N : 1;
    const long long seg_chunks = static_cast<long long>(p.tiles_y) * p.tiles_x * (p.s_n != 0 ? 1 : p.N);
    // split the pixel range only while the tiles alone leave SMs idle (a split multiplies the reduction traffic), and keep
    // the item count within one wave of the persistent grid
    long long ks = num_sms / (tiles * nseg);
    const long long max_ks = (seg_chunks + 3) / 4;   // at least four chunks per split
    if (ks > max_ks) ks = max_ks;
    if (ks < 1) ks = 1;
    p.ksplit = static_cast<int>(ks);
    const long long items = tiles * nseg * ks;
    const unsigned grid = static_cast<unsigned>(items < num_sms ? items : num_sms);
    wgrad_tc_kernel<<<grid, 256, smem, s>>>(p, map_a, map_b, static_cast<int>(items));
    return (int)cudaGetLastError();
}

}  // namespace usb
