// Device-side front-end glue of UnitSpeech.execute_text_to_speech (unitspeech/unitspeech.py:421-441): durations ->
// frame counts, frame mask, monotonic alignment path (generate_path, unitspeech/util.py:27-41) and the aligned
// conditioning cond_y = path^T . cond_x -- without the reference's host round trip int(y_lengths.max()) (:428).
// The caller fixes the frame capacity T (a multiple of 2^num_downsamplings); utterances shorter than T are masked.
#include <string>

#include <cuda_runtime.h>

#include "../../include/unitspeech_b200.h"
#include "conv_igemm.h"   // usb::set_error

namespace {

// grid B, 256 threads, dynamic smem: (Tx + 1) doubles (exclusive/inclusive duration sums)
__global__ void __launch_bounds__(256) align_expand_kernel(const float* __restrict__ w, const float* __restrict__ x_mask,
                                                           const float* __restrict__ cond_x, int Tx, int F, int T,
                                                           long long* __restrict__ y_lengths, float* __restrict__ y_mask,
                                                           float* __restrict__ attn, float* __restrict__ cond_y) {
    extern __shared__ double cum[];   // cum[i] = sum_{k<=i} w[k]; durations are integers after ceil (:424), sums are exact
    __shared__ long long ylen_s;
    const int b = blockIdx.x;
    const float* wb = w + static_cast<long long>(b) * Tx;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < Tx; ++i) {
            s += static_cast<double>(wb[i]);
            cum[i] = s;
        }
        long long yl = static_cast<long long>(static_cast<float>(s));       // .long() of the fp32 sum (:427)
        if (s < 1.0) yl = 1;                                                 // clamp_min(., 1)
        if (yl > T) yl = T;   // frame capacity: mask / path / cond_y end at T, so the reported length does too
        ylen_s = yl;
        y_lengths[b] = yl;
    }
    __syncthreads();
    const long long ylen = ylen_s;
    float* attn_b = attn + static_cast<long long>(b) * Tx * T;
    for (long long i = threadIdx.x; i < static_cast<long long>(Tx) * T; i += blockDim.x) attn_b[i] = 0.f;
    __syncthreads();
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        const bool valid = j < ylen;
        y_mask[static_cast<long long>(b) * T + j] = valid ? 1.f : 0.f;
        // token of frame j: the first i with cum[i] > j  (path[i][j] = [j < cum[i]] - [j < cum[i-1]])
        int lo = 0, hi = Tx;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cum[mid] > static_cast<double>(j)) hi = mid;
            else lo = mid + 1;
        }
        const int tok = lo;
        const bool on = valid && tok < Tx && x_mask[static_cast<long long>(b) * Tx + tok] != 0.f;
        if (on) attn_b[static_cast<long long>(tok) * T + j] = 1.f;
        for (int f = 0; f < F; ++f)
            cond_y[(static_cast<long long>(b) * F + f) * T + j] = on ? cond_x[(static_cast<long long>(b) * F + f) * Tx + tok] : 0.f;
    }
}

}  // namespace

extern "C" int usb_align_expand(usb_handle* h, const float* w_ceil, const float* x_mask, const float* cond_x, int32_t B,
                                int32_t Tx, int32_t n_feats, int32_t T, int64_t* y_lengths, float* y_mask, float* attn,
                                float* cond_y, uint64_t stream) {
    if (!h || !w_ceil || !x_mask || !cond_x || !y_lengths || !y_mask || !attn || !cond_y) return usb::set_error("null argument");
    if (B < 1 || Tx < 1 || T < 1 || n_feats < 1) return usb::set_error("bad front-end geometry");
    const size_t smem = static_cast<size_t>(Tx) * sizeof(double);
    if (smem > 200 * 1024) return usb::set_error("too many tokens for the alignment kernel");
    {
        cudaError_t e = cudaSetDevice(usb::handle_device(h));
        if (e != cudaSuccess) return usb::set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    if (smem > 48 * 1024) cudaFuncSetAttribute(align_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    align_expand_kernel<<<B, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        w_ceil, x_mask, cond_x, Tx, n_feats, T, reinterpret_cast<long long*>(y_lengths), y_mask, attn, cond_y);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return usb::set_error(std::string("align_expand launch: ") + cudaGetErrorString(e));
    return 0;
}
