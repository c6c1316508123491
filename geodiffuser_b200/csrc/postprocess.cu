// geodiffuser_b200/csrc/postprocess.cu
//
// the per-edit tail after the DDIM loop -- masked histogram matching of the edited image against the (warped) input
// image (image_processing.py:24-77 `_match_cumulative_cdf` / `masked_histogram_matching`, called at editor.py:680, 683, 690).  The
// reference does it in numpy on the host (bincount, cumsum, np.interp, fancy-index lookup); here three small launches keep the image on
// the device:
//   1. hist_kernel   masked 256-bin histograms of source and template, per channel (integer atomics: order-independent, exact)
//   2. lut_kernel    cumulative counts -> quantiles -> np.interp(src_quantiles, tmpl_quantiles, 0..255) in IEEE double with explicit
//                    mul / add / div (no FMA contraction), following numpy's arr_interp: bit-identical look-up table
//   3. apply_kernel  out[p, c] = lut[c][source[p, c]]  (float64, like the reference's result)
// HBM-bound byte work (3 B in, 24 B out per pixel); the image warp that precedes it is the splat of geometry.cu.
#include "common.cuh"

namespace gd {

// counts (C, 2, 256) int32: [c][0] = source under mask_source, [c][1] = template under mask
__global__ void hist_kernel(const unsigned char* __restrict__ src, const unsigned char* __restrict__ tmpl, const float* __restrict__ mask,
                            const float* __restrict__ mask_src, long npix, int C, int* __restrict__ counts) {
    extern __shared__ int sh[];   // C * 2 * 256
    for (int i = threadIdx.x; i < C * 512; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
        const bool ms = mask_src[p] > 0.5f, mt = mask[p] > 0.5f;
        for (int c = 0; c < C; ++c) {
            if (ms) atomicAdd(&sh[(c * 2) * 256 + src[p * C + c]], 1);
            if (mt) atomicAdd(&sh[(c * 2 + 1) * 256 + tmpl[p * C + c]], 1);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 512; i += blockDim.x)
        if (sh[i]) atomicAdd(&counts[i], sh[i]);
}

// one block of 256 threads per channel; lut (C, 256) double
__global__ void lut_kernel(const int* __restrict__ counts, double* __restrict__ lut) {
    __shared__ double qs[256], qt[256];
    __shared__ long tot[2];
    const int c = blockIdx.x, t = threadIdx.x;
    if (t < 2) {
        const int* h = counts + (c * 2 + t) * 256;
        long run = 0;
        double* q = t == 0 ? qs : qt;
        for (int i = 0; i < 256; ++i) { run += h[i]; q[i] = (double)run; }       // np.cumsum (int64), exact in double
        tot[t] = run;
    }
    __syncthreads();
    // quantiles: cumsum / size (one IEEE division each)
    const double s_q = __ddiv_rn(qs[t], (double)tot[0]);
    const double t_q = __ddiv_rn(qt[t], (double)tot[1]);
    __syncthreads();
    qs[t] = s_q; qt[t] = t_q;
    __syncthreads();
    // np.interp(x = qs[t], xp = qt, fp = 0..255)  (numpy/core/src/multiarray/compiled_base.c arr_interp)
    const double x = qs[t];
    double r;
    if (x > qt[255]) r = 255.0;
    else if (x < qt[0]) r = 0.0;
    else {
        int lo = 0, hi = 256;                 // j = (number of xp[i] <= x) - 1
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (qt[mid] <= x) lo = mid + 1; else hi = mid; }
        const int j = lo - 1;
        if (j == 255) r = 255.0;
        else if (qt[j] == x) r = (double)j;
        else {
            const double slope = __ddiv_rn(__dsub_rn((double)(j + 1), (double)j), __dsub_rn(qt[j + 1], qt[j]));
            r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, qt[j])), (double)j);
        }
    }
    lut[c * 256 + t] = r;
}

__global__ void lut_apply_kernel(const unsigned char* __restrict__ src, const double* __restrict__ lut, long n, int C, double* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = lut[(int)(i % C) * 256 + src[i]];
}

}  // namespace gd

using namespace gd;

extern "C" {

// image_processing.py:24-77.  source, template (npix, C) uint8 (HWC images); mask (template side), mask_source (npix) float, > 0.5 selects;
// counts (C, 2, 256) int32 and lut (C, 256) double are caller-provided scratch (counts need not be cleared); out (npix, C) double.
int gd_masked_histogram_match(const unsigned char* source, const unsigned char* tmpl, const float* mask, const float* mask_source, long npix,
                              int C, int* counts, double* lut, double* out, void* stream) {
    GD_CHECK_ARG(source && tmpl && mask && mask_source && counts && lut && out && npix > 0 && C > 0 && C <= 8);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)C * 512 * sizeof(int), st);
    if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    int blocks = ceil_div(npix, 256 * 8);
    if (blocks > 296) blocks = 296;
    hist_kernel<<<blocks, 256, (size_t)C * 512 * sizeof(int), st>>>(source, tmpl, mask, mask_source, npix, C, counts);
    GD_CHECK_LAUNCH();
    lut_kernel<<<C, 256, 0, st>>>(counts, lut);
    GD_CHECK_LAUNCH();
    lut_apply_kernel<<<ceil_div(npix * C, 256), 256, 0, st>>>(source, lut, npix * C, C, out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
