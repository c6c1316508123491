// geodiffuser_b200/csrc/body_norm.cu
//
// GroupNorm (+ optional SiLU) on channels-last bf16 activations, forward and input-gradient, for the UNet body that CALLS the path
// (61 GroupNorms per UNet evaluation).  torch's CUDA group_norm has no NHWC kernel: it copies the activation to NCHW, runs
// RowwiseMoments + ComputeFusedParams + an elementwise apply, a separate SiLU, and cuDNN then copies back to NHWC for the next
// convolution -- 6 launches and ~5 passes over the tensor per norm, ~27 % of the device time of a gradient-free UNet pass on a B200
// (profiles/r01c_phase_kernels.md).  Here: two launches, two reads (the second one from L2) and one write, HBM/L2-bound.
//
//   pass 1  gn_partial_kernel : per (batch, row chunk): per-group partial sums -> partial (B, n_chunks, G, 2)
//   pass 2  gn_apply_kernel   : prologue adds the partial sums of its batch entry in a fixed order (deterministic; no atomics),
//                               then y = silu?(x_hat * gamma + beta)            (forward)
//                               or  dx = rstd * (t - mean_g(t) - x_hat * mean_g(t * x_hat)), t = dz * gamma   (backward)
// Layout: x (B, HW, C) bf16 with C % 8 == 0 (16-byte vectors of 8 channels); thread -> fixed 8-channel slot, rows strided.
// Statistics: fp32 sum / sum of squares per thread over <= ~30 values, combined in fp32 in a fixed tree; variance clamped at 0.
#include "common.cuh"

namespace gd {

typedef __nv_bfloat16 bf16;

struct GnParams {
    const bf16* x;        // (B, HW, C)
    const bf16* dy;       // backward only
    const void* gamma;    // (C) bf16 or fp32
    const void* beta;     // (C)
    int w_bf16;
    float* partial;       // (B, n_chunks, G, 2)
    float* stats;         // (B, G, 2) mean, rstd : written by the forward, read by the backward
    bf16* out;            // y or dx (B, HW, C)
    int B, HW, C, G, Cg, n_chunks, rows_per_chunk, nslot, rpi;
    float eps;
    int silu;
};

__device__ __forceinline__ float ldw(const void* p, int is_bf16, int i) {
    return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

__device__ __forceinline__ float silu_grad(float z) {
    const float s = 1.0f / (1.0f + __expf(-z));
    return s * (1.0f + z * (1.0f - s));
}

template <int BWD>
__global__ void __launch_bounds__(320) gn_partial_kernel(const GnParams p) {
    extern __shared__ float sh[];   // (rpi, C, 2)
    const int t = threadIdx.x, b = blockIdx.y, chunk = blockIdx.x;
    const int C = p.C, nslot = p.nslot, rpi = p.rpi, Cg = p.Cg;
    const bool active = t < nslot * rpi;
    const int slot = t % nslot, r = t / nslot;
    const int c0 = slot * 8;
    float a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
    if (active) {
        const int row0 = chunk * p.rows_per_chunk, row1 = min(p.HW, row0 + p.rows_per_chunk);
        float mean[8], rstd[8], ga[8], be[8];
        if (BWD) {
            int g = c0 / Cg;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                while (c0 + j >= (g + 1) * Cg) ++g;
                mean[j] = p.stats[((long)b * p.G + g) * 2];
                rstd[j] = p.stats[((long)b * p.G + g) * 2 + 1];
                ga[j] = ldw(p.gamma, p.w_bf16, c0 + j);
                be[j] = ldw(p.beta, p.w_bf16, c0 + j);
            }
        }
        for (int row = row0 + r; row < row1; row += rpi) {
            const long off = ((long)b * p.HW + row) * C + c0;
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(p.x + off), f);
            if (!BWD) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { a1[j] += f[j]; a2[j] += f[j] * f[j]; }
            } else {
                float d[8];
                unpack8(*reinterpret_cast<const uint4*>(p.dy + off), d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (f[j] - mean[j]) * rstd[j];
                    float dz = d[j];
                    if (p.silu) dz *= silu_grad(xh * ga[j] + be[j]);
                    const float tt = dz * ga[j];
                    a1[j] += tt;
                    a2[j] += tt * xh;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sh[((long)r * C + c0 + j) * 2] = a1[j];
            sh[((long)r * C + c0 + j) * 2 + 1] = a2[j];
        }
    }
    __syncthreads();
    // 8 threads per group, fixed assignment and fixed shuffle tree: the same bits on every run
    if (t < p.G * 8) {
        const unsigned lanes = __activemask();
        const int g = t >> 3, j = t & 7;
        float s1 = 0.f, s2 = 0.f;
        for (int rr = 0; rr < rpi; ++rr)
            for (int c = g * Cg + j; c < (g + 1) * Cg; c += 8) { s1 += sh[((long)rr * C + c) * 2]; s2 += sh[((long)rr * C + c) * 2 + 1]; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { s1 += __shfl_xor_sync(lanes, s1, o); s2 += __shfl_xor_sync(lanes, s2, o); }
        if (j == 0) {
            float* dst = p.partial + (((long)b * p.n_chunks + chunk) * p.G + g) * 2;
            dst[0] = s1; dst[1] = s2;
        }
    }
}

template <int BWD>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnParams p) {
    __shared__ float st[32][4];     // forward: mean, rstd ; backward: mean, rstd, m1, m2
    const int t = threadIdx.x, b = blockIdx.y;
    const int C = p.C, Cg = p.Cg, nslot = p.nslot;
    if (t < p.G * 8) {
        const unsigned lanes = __activemask();
        const int g = t >> 3, j = t & 7;
        float s1 = 0.f, s2 = 0.f;
        for (int k = j; k < p.n_chunks; k += 8) {
            const float* src = p.partial + (((long)b * p.n_chunks + k) * p.G + g) * 2;
            s1 += src[0]; s2 += src[1];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { s1 += __shfl_xor_sync(lanes, s1, o); s2 += __shfl_xor_sync(lanes, s2, o); }
        if (j == 0) {
            const float inv_n = 1.0f / ((float)p.HW * (float)Cg);
            if (!BWD) {
                const float mean = s1 * inv_n;
                const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
                const float rstd = rsqrtf(var + p.eps);
                st[g][0] = mean; st[g][1] = rstd;
                if (blockIdx.x == 0 && p.stats) { p.stats[((long)b * p.G + g) * 2] = mean; p.stats[((long)b * p.G + g) * 2 + 1] = rstd; }
            } else {
                st[g][0] = p.stats[((long)b * p.G + g) * 2]; st[g][1] = p.stats[((long)b * p.G + g) * 2 + 1];
                st[g][2] = s1 * inv_n; st[g][3] = s2 * inv_n;
            }
        }
    }
    __syncthreads();
    const long nvec = (long)p.HW * nslot;
    for (long i = (long)blockIdx.x * blockDim.x + t; i < nvec; i += (long)gridDim.x * blockDim.x) {
        const int slot = (int)(i % nslot);
        const int c0 = slot * 8;
        const long off = (long)b * p.HW * C + i * 8;
        float f[8], d[8];
        unpack8(*reinterpret_cast<const uint4*>(p.x + off), f);
        if (BWD) unpack8(*reinterpret_cast<const uint4*>(p.dy + off), d);
        int g = c0 / Cg;
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            while (c0 + j >= (g + 1) * Cg) ++g;
            const float ga = ldw(p.gamma, p.w_bf16, c0 + j), be = ldw(p.beta, p.w_bf16, c0 + j);
            const float xh = (f[j] - st[g][0]) * st[g][1];
            if (!BWD) {
                float z = xh * ga + be;
                if (p.silu) z = z / (1.0f + __expf(-z));
                r[j] = z;
            } else {
                float dz = d[j];
                if (p.silu) dz *= silu_grad(xh * ga + be);
                r[j] = st[g][1] * (dz * ga - st[g][2] - xh * st[g][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
        *reinterpret_cast<uint4*>(p.out + off) = o;
    }
}

static int gn_plan(GnParams& p, long ws_floats) {
    p.Cg = p.C / p.G;
    p.nslot = p.C / 8;
    p.rpi = 320 / p.nslot;
    if (p.rpi < 1) return set_error(GD_ERR_UNSUPPORTED, "group norm: C = %d > 2560", p.C);
    if (p.rpi > p.HW) p.rpi = p.HW;
    int n_chunks = (2 * 148 + p.B - 1) / p.B;                 // ~2 CTAs per SM over the batch
    const int max_chunks = (p.HW + p.rpi - 1) / p.rpi;         // at least one row per row-lane
    if (n_chunks > max_chunks) n_chunks = max_chunks;
    p.rows_per_chunk = (p.HW + n_chunks - 1) / n_chunks;
    p.n_chunks = (p.HW + p.rows_per_chunk - 1) / p.rows_per_chunk;
    if ((long)p.B * p.n_chunks * p.G * 2 > ws_floats)
        return set_error(GD_ERR_INVALID, "group norm: workspace of %ld floats < %ld", ws_floats, (long)p.B * p.n_chunks * p.G * 2);
    return GD_OK;
}

template <int BWD> static int gn_run(GnParams& p, long ws_floats, cudaStream_t st) {
    int rc = gn_plan(p, ws_floats);
    if (rc != GD_OK) return rc;
    const int threads = (p.nslot * p.rpi + 31) / 32 * 32;
    const int t1 = threads < 256 ? 256 : threads;              // the group reduction wants G * 8 <= 256 threads
    const size_t smem = (size_t)p.rpi * p.C * 2 * sizeof(float);
    gn_partial_kernel<BWD><<<dim3(p.n_chunks, p.B), t1, smem, st>>>(p);
    GD_CHECK_LAUNCH();
    const long nvec = (long)p.HW * p.nslot;
    int blocks = (int)((nvec + 256 * 4 - 1) / (256 * 4));      // ~4 vectors of 8 channels per thread
    const int cap = (4 * 148 + p.B - 1) / p.B;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    gn_apply_kernel<BWD><<<dim3(blocks, p.B), 256, 0, st>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" {

// y = silu?(group_norm(x)) for x (B, HW, C) bf16 channels-last; gamma / beta (C) bf16 (w_is_bf16 = 1) or fp32.  stats (B, G, 2) receives
// (mean, rstd) for the backward (may be NULL).  workspace: >= gd_group_norm_nhwc_workspace(B, HW, C, G) floats.
int gd_group_norm_nhwc_fwd(const void* x, const void* gamma, const void* beta, int w_is_bf16, int B, int HW, int C, int G, float eps, int silu,
                           float* workspace, long workspace_floats, float* stats, void* y, void* stream) {
    GD_CHECK_ARG(x && gamma && beta && workspace && y && B > 0 && HW > 0 && C > 0 && G > 0 && G <= 32);
    if (C % 8 != 0 || C % G != 0) return set_error(GD_ERR_UNSUPPORTED, "group norm: C = %d must be a multiple of 8 and of G = %d", C, G);
    GnParams p;
    p.x = (const bf16*)x; p.dy = nullptr; p.gamma = gamma; p.beta = beta; p.w_bf16 = w_is_bf16; p.partial = workspace; p.stats = stats;
    p.out = (bf16*)y; p.B = B; p.HW = HW; p.C = C; p.G = G; p.eps = eps; p.silu = silu;
    return gn_run<0>(p, workspace_floats, (cudaStream_t)stream);
}

// dx of the above given dy (B, HW, C) bf16 and the forward's stats; gradients of gamma / beta are not produced (the body's weights
// are frozen in the edit loop: optimization.py:213-219 differentiates w.r.t. the latent and the context only).
int gd_group_norm_nhwc_bwd(const void* x, const void* dy, const void* gamma, const void* beta, int w_is_bf16, const float* stats, int B, int HW,
                           int C, int G, int silu, float* workspace, long workspace_floats, void* dx, void* stream) {
    GD_CHECK_ARG(x && dy && gamma && beta && stats && workspace && dx && B > 0 && HW > 0 && C > 0 && G > 0 && G <= 32);
    if (C % 8 != 0 || C % G != 0) return set_error(GD_ERR_UNSUPPORTED, "group norm: C = %d must be a multiple of 8 and of G = %d", C, G);
    GnParams p;
    p.x = (const bf16*)x; p.dy = (const bf16*)dy; p.gamma = gamma; p.beta = beta; p.w_bf16 = w_is_bf16; p.partial = workspace;
    p.stats = const_cast<float*>(stats); p.out = (bf16*)dx; p.B = B; p.HW = HW; p.C = C; p.G = G; p.eps = 0.f; p.silu = silu;
    return gn_run<1>(p, workspace_floats, (cudaStream_t)stream);
}

// floats of workspace the two entry points above need (an upper bound that does not depend on the plan's details)
int gd_group_norm_nhwc_workspace(int B, int HW, int C, int G) {
    (void)HW; (void)C;
    return (2 * 148 + B) * G * 2;
}

}  // extern "C"
