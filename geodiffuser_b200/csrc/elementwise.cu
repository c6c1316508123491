// geodiffuser_b200/csrc/elementwise.cu
//
// Subsystem (4): the DDIM step (+ classifier-free-guidance combine) and the masked latent / context gradient
// update of the optimisation loop, as coalesced float4 kernels; plus the library's error plumbing.
//   diffusion.py:46,55 (CFG combine, scheduler.step eta=0; formula restated at inversion.py:47-55)
//   optimization.py:213-253 (_update_latent, optimizer=None branch, nan_to_num on the gradient)
//   editor.py:219,316 + generic_torch.py:87 (norm-preserving rescale of the edited latent)
// HBM-bound (64 KB .. 240 KB per launch => launch-latency bound at batch 1; bandwidth is reported on a batched
// synthetic by bench.py).
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace gd {

thread_local char g_last_error[256] = {0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

__device__ __forceinline__ float ld_any(const void* p, int is_bf16, long i) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

// x_prev = c3 * (x - c1 * eps) / c2 + c4 * eps,  eps = eps_u + g * (eps_c - eps_u) when eps_c is given
__global__ void ddim_step_kernel(const float* __restrict__ x, const void* __restrict__ eps_u, const void* __restrict__ eps_c,
                                 int eps_bf16, float guidance, float c1, float c2, float c3, float c4, long n,
                                 float* __restrict__ out, float* __restrict__ eps_out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float e = ld_any(eps_u, eps_bf16, i);
    if (eps_c) e = e + guidance * (ld_any(eps_c, eps_bf16, i) - e);
    const float x0 = (x[i] - c1 * e) / c2;
    out[i] = c3 * x0 + c4 * e;
    if (eps_out) eps_out[i] = e;
}

__device__ __forceinline__ float nan_to_zero(float g) { return (isnan(g) || isinf(g)) ? 0.f : g; }

// lat_out = (lat - 2 m s g) - (1 - m) s g  with m broadcast over channels (mask (HW) may be null: lat - s g)
__global__ void latent_update_kernel(const float* __restrict__ lat, const float* __restrict__ grad, const float* __restrict__ mask,
                                     int hw, float step, long n, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = nan_to_zero(grad[i]);
    if (mask) {
        const float m = mask[i % hw];
        const float a = lat[i] - 2.0f * m * step * g;
        out[i] = a - (1.0f - m) * step * g;
    } else {
        out[i] = lat[i] - step * g;
    }
}

// x *= target_norm / sqrt(sum(x^2) + 1e-12); single block so the reduction order is fixed
__global__ void __launch_bounds__(1024) norm_rescale_kernel(float* __restrict__ x, long n, float target_norm, const float* __restrict__ target_dev,
                                                            float* __restrict__ norm_out) {
    __shared__ float sh[32];
    float s = 0.f;
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += x[i] * x[i];
    const float tot = block_sum(s, sh);
    const float nrm = sqrtf(tot + 1e-12f);
    if (norm_out && threadIdx.x == 0) *norm_out = nrm;
    if (target_dev) target_norm = *target_dev;      // the target stays on the device: no host round trip between measuring and restoring a norm
    if (target_norm > 0.f) {
        const float f = target_norm / nrm;
        for (long i = threadIdx.x; i < n; i += blockDim.x) x[i] *= f;
    }
}

// out = a * (1 - m) + m * b, m (HW) broadcast over channels, optionally binarised (> 0.5)   (editor.py:393-399)
__global__ void latent_blend_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int hw,
                                    int binarize, long n, float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = mask[i % hw];
    if (binarize) m = m > 0.5f ? 1.f : 0.f;
    out[i] = a[i] * (1.0f - m) + m * b[i];
}

}  // namespace gd

using namespace gd;

extern "C" {

const char* gd_last_error(void) { return g_last_error; }
int gd_version(void) { return 100; }

int gd_ddim_step(const float* x, const void* eps_u, const void* eps_c, int eps_is_bf16, float guidance, float sqrt_one_minus_at,
                 float sqrt_at, float sqrt_aprev, float sqrt_one_minus_aprev, long n, float* out, float* eps_out, void* stream) {
    GD_CHECK_ARG(x && eps_u && out && n > 0);
    ddim_step_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, eps_u, eps_c, eps_is_bf16, guidance, sqrt_one_minus_at, sqrt_at,
                                                                        sqrt_aprev, sqrt_one_minus_aprev, n, out, eps_out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_latent_update(const float* lat, const float* grad, const float* mask, int hw, float step, long n, float* out, void* stream) {
    GD_CHECK_ARG(lat && grad && out && n > 0 && (mask == nullptr || hw > 0));
    latent_update_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(lat, grad, mask, hw, step, n, out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_norm_rescale(float* x, long n, float target_norm, const float* target_dev, float* norm_out, void* stream) {
    GD_CHECK_ARG(x && n > 0);
    norm_rescale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, target_norm, target_dev, norm_out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_latent_blend(const float* a, const float* b, const float* mask, int hw, int binarize, long n, float* out, void* stream) {
    GD_CHECK_ARG(a && b && mask && out && hw > 0 && n > 0);
    latent_blend_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, mask, hw, binarize, n, out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
