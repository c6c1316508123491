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
    const bf16* pre_bias; // optional (B, C): the normalised tensor is x + pre_bias[b, c] (conv bias + time-embedding shift folded in)
    const void* gamma;    // (C) bf16 or fp32
    const void* beta;     // (C)
    int w_bf16;
    float* partial;       // (B, n_chunks, G, 2)
    float* stats;         // (B, G, 2) mean, rstd : written by the forward, read by the backward
    float* red;           // (B, G, 2) backward only: mean_g(t), mean_g(t * x_hat)
    unsigned* counter;    // (B) zero on entry, zero again on exit: the last chunk of a batch entry finishes that entry's reduction
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
        float mean[8], rstd[8], ga[8], be[8], pb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pb[j] = 0.f;
        if (p.pre_bias) unpack8(*reinterpret_cast<const uint4*>(p.pre_bias + (long)b * C + c0), pb);
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
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += pb[j];
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
            __threadfence();
        }
    }
    // the chunk that finishes last adds the partial sums of its batch entry -- in chunk order, whichever chunk it is, so the result does not
    // depend on the arrival order -- and publishes the group statistics; the apply kernel then starts without a reduction prologue
    __shared__ int is_last;
    __syncthreads();
    if (t == 0) is_last = (atomicAdd(p.counter + b, 1u) == (unsigned)(p.n_chunks - 1));
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (t < p.G * 8) {
        const unsigned lanes = __activemask();
        const int g = t >> 3, j = t & 7;
        float u1[4] = {0.f, 0.f, 0.f, 0.f}, u2[4] = {0.f, 0.f, 0.f, 0.f};
        const float* src = p.partial + ((long)b * p.n_chunks * p.G + g) * 2;
        int k = j;
        for (; k + 24 < p.n_chunks; k += 32) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 v = __ldcg(reinterpret_cast<const float2*>(src + (long)(k + 8 * u) * p.G * 2));
                u1[u] += v.x; u2[u] += v.y;
            }
        }
        for (int u = 0; k < p.n_chunks; k += 8, ++u) {
            const float2 v = __ldcg(reinterpret_cast<const float2*>(src + (long)k * p.G * 2));
            u1[u & 3] += v.x; u2[u & 3] += v.y;
        }
        float s1 = (u1[0] + u1[1]) + (u1[2] + u1[3]), s2 = (u2[0] + u2[1]) + (u2[2] + u2[3]);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { s1 += __shfl_xor_sync(lanes, s1, o); s2 += __shfl_xor_sync(lanes, s2, o); }
        if (j == 0) {
            const float inv_n = 1.0f / ((float)p.HW * (float)Cg);
            if (!BWD) {
                const float mean = s1 * inv_n;
                const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
                p.stats[((long)b * p.G + g) * 2] = mean;
                p.stats[((long)b * p.G + g) * 2 + 1] = rsqrtf(var + p.eps);
            } else {
                p.red[((long)b * p.G + g) * 2] = s1 * inv_n;
                p.red[((long)b * p.G + g) * 2 + 1] = s2 * inv_n;
            }
        }
    }
    if (t == 0) p.counter[b] = 0u;
}

template <int BWD>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnParams p) {
    __shared__ float st[32][4];     // forward: mean, rstd ; backward: mean, rstd, m1, m2
    const int t = threadIdx.x, b = blockIdx.y;
    const int C = p.C, Cg = p.Cg, nslot = p.nslot;
    if (t < p.G) {
        st[t][0] = p.stats[((long)b * p.G + t) * 2]; st[t][1] = p.stats[((long)b * p.G + t) * 2 + 1];
        if (BWD) { st[t][2] = p.red[((long)b * p.G + t) * 2]; st[t][3] = p.red[((long)b * p.G + t) * 2 + 1]; }
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
        if (p.pre_bias) {
            float pb[8];
            unpack8(*reinterpret_cast<const uint4*>(p.pre_bias + (long)b * C + c0), pb);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += pb[j];
        }
        int g = c0 / Cg;
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
        float r[8], gav[8], bev[8];
        if (p.w_bf16) {
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.gamma) + c0), gav);
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.beta) + c0), bev);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { gav[j] = reinterpret_cast<const float*>(p.gamma)[c0 + j]; bev[j] = reinterpret_cast<const float*>(p.beta)[c0 + j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            while (c0 + j >= (g + 1) * Cg) ++g;
            const float ga = gav[j], be = bev[j];
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

// ---- forward in ONE launch: thread-block clusters + distributed shared memory ----------------------------------------------------
// The two-kernel scheme above pays two launch latencies plus a fence / atomic / re-read chain per norm (~7 us floor each, measured), which
// dominates at the sizes of the edit loop (0.3 .. 16 MB per tensor).  Here one cluster of 8 CTAs owns a (batch entry, slab of `gs` groups):
// CTA r stages rows [r*rpc, (r+1)*rpc) x the slab's W channels in shared memory (ONE global read), reduces its per-group partial sums, the
// cluster exchanges them through DSMEM (rank order: deterministic), and each CTA normalises its staged tile straight out of shared memory.
constexpr int GNC_CLUSTER = 8;
constexpr int GNC_THREADS = 512;

struct GnClusterParams {
    const bf16* x; const bf16* pre_bias; const void* gamma; const void* beta; int w_bf16;
    float* stats; bf16* out;
    int B, HW, C, G, Cg, gs, W, nv, rpi, rpc, n_slabs;
    float eps; int silu;
};

__global__ void __cluster_dims__(GNC_CLUSTER, 1, 1) __launch_bounds__(GNC_THREADS)
gn_cluster_fwd_kernel(const GnClusterParams p) {
    extern __shared__ __align__(16) unsigned char gsm[];
    __shared__ float part[32][2];      // this CTA's partial (sum, sumsq) per group of the slab
    __shared__ float st[32][2];        // mean, rstd per group of the slab
    const int t = threadIdx.x;
    unsigned rank, cid;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(cid));
    const int b = cid / p.n_slabs, slab = cid % p.n_slabs;
    const int W = p.W, nv = p.nv, rpi = p.rpi, Cg = p.Cg, C = p.C;
    const int c_base = slab * W;
    const int r0 = rank * p.rpc, r1 = min(p.HW, r0 + p.rpc);
    const int nact = nv * rpi;                                                   // active threads: (row lane, 8-channel slot)
    bf16* tile = reinterpret_cast<bf16*>(gsm);                                   // [rpc][W]
    float* red = reinterpret_cast<float*>(gsm + (size_t)p.rpc * W * 2);          // [16][nact]: k-th accumulator of thread t at red[k*nact + t]
    float* chs = red + 16 * nact;                                                // [W][2]: per-channel sums of this CTA
    const bool active = t < nact;
    const int v = t % nv, rl = t / nv, c0 = v * 8;
    float pb[8], a1[8], a2[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { pb[j] = 0.f; a1[j] = 0.f; a2[j] = 0.f; ga[j] = 1.f; be[j] = 0.f; }
    if (active) {
        if (p.pre_bias) unpack8(*reinterpret_cast<const uint4*>(p.pre_bias + (long)b * C + c_base + c0), pb);
        // affine parameters of this thread's 8 channels: requested now, consumed after the reduction (their latency hides under phase 1)
        if (p.w_bf16) {
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.gamma) + c_base + c0), ga);
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.beta) + c_base + c0), be);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ga[j] = reinterpret_cast<const float*>(p.gamma)[c_base + c0 + j]; be[j] = reinterpret_cast<const float*>(p.beta)[c_base + c0 + j]; }
        }
#pragma unroll 4
        for (int row = r0 + rl; row < r1; row += rpi) {
            const uint4 raw = *reinterpret_cast<const uint4*>(p.x + ((long)b * p.HW + row) * C + c_base + c0);
            *reinterpret_cast<uint4*>(tile + (long)(row - r0) * W + c0) = raw;
            float f[8];
            unpack8(raw, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float xv = f[j] + pb[j]; a1[j] += xv; a2[j] += xv * xv; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[j * nact + t] = a1[j]; red[(8 + j) * nact + t] = a2[j]; }      // conflict-free: t is the fast index
    }
    __syncthreads();
    // per-channel sums over the row lanes (2W threads, fixed order), then per-group sums (8 threads per group, fixed tree)
    for (int i = t; i < 2 * W; i += blockDim.x) {
        const int which = i / W, c = i - which * W;
        const float* src = red + ((which * 8 + (c & 7)) * nact) + (c >> 3);
        float s = 0.f;
        for (int rr = 0; rr < rpi; ++rr) s += src[rr * nv];
        chs[c * 2 + which] = s;
    }
    __syncthreads();
    if (t < p.gs * 8) {
        const unsigned lanes = __activemask();
        const int g = t >> 3, j = t & 7;
        float s1 = 0.f, s2 = 0.f;
        for (int c = g * Cg + j; c < (g + 1) * Cg; c += 8) { s1 += chs[c * 2]; s2 += chs[c * 2 + 1]; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { s1 += __shfl_xor_sync(lanes, s1, o); s2 += __shfl_xor_sync(lanes, s2, o); }
        if (j == 0) { part[g][0] = s1; part[g][1] = s2; }
    }
    // cluster barrier #1: every CTA's `part` is complete and visible cluster-wide
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (t < p.gs) {
        float s1 = 0.f, s2 = 0.f, u1[GNC_CLUSTER], u2[GNC_CLUSTER];
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(&part[t][0]);
#pragma unroll
        for (int r = 0; r < GNC_CLUSTER; ++r) {        // all 16 remote loads in flight at once (~215 clk each), then summed in rank order
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(u1[r]) : "r"(remote));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(u2[r]) : "r"(remote + 4));
        }
#pragma unroll
        for (int r = 0; r < GNC_CLUSTER; ++r) { s1 += u1[r]; s2 += u2[r]; }   // rank order: the same sum in every CTA, on every run
        const float inv_n = 1.0f / ((float)p.HW * (float)Cg);
        const float mean = s1 * inv_n;
        const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.eps);
        st[t][0] = mean; st[t][1] = rstd;
        if (rank == 0 && p.stats) {
            const int g = slab * p.gs + t;
            p.stats[((long)b * p.G + g) * 2] = mean; p.stats[((long)b * p.G + g) * 2 + 1] = rstd;
        }
    }
    // cluster barrier #2: nobody leaves (and frees its shared memory) while a sibling may still be reading it
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (!active) return;
    {
        int g = c0 / Cg;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            while (c0 + j >= (g + 1) * Cg) ++g;
            // fold everything into y = x * a + c:  a = rstd * gamma,  c = (pre_bias - mean) * a + beta
            const float a = st[g][1] * ga[j];
            be[j] = (pb[j] - st[g][0]) * a + be[j];
            ga[j] = a;
        }
    }
#pragma unroll 4
    for (int row = r0 + rl; row < r1; row += rpi) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(tile + (long)(row - r0) * W + c0), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float z = f[j] * ga[j] + be[j];
            if (p.silu) z = __fdividef(z, 1.0f + __expf(-z));
            f[j] = z;
        }
        uint4 o;
        __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p.out + ((long)b * p.HW + row) * C + c_base + c0) = o;
    }
}

// plan of the cluster kernel; false if the shape does not fit (caller falls back to the two-kernel scheme)
static bool gn_cluster_plan(GnClusterParams& p, size_t* smem) {
    p.Cg = p.C / p.G;
    int gs = 0;
    for (int cand = 1; cand <= p.G; cand *= 2) {
        if (p.G % cand) break;
        const int W = cand * p.Cg;
        if ((W % 8) == 0 && W >= 40) { gs = cand; break; }
    }
    if (!gs || gs > 32) return false;
    p.gs = gs; p.W = gs * p.Cg; p.nv = p.W / 8;
    if (p.nv > GNC_THREADS) return false;
    p.rpi = GNC_THREADS / p.nv;
    p.rpc = (p.HW + GNC_CLUSTER - 1) / GNC_CLUSTER;
    if (p.rpi > (p.rpc + 1) / 2) p.rpi = (p.rpc + 1) / 2;      // at least two rows per row lane, so few-row tiles do not launch idle warps
    if (p.rpi < 1) p.rpi = 1;
    p.n_slabs = p.G / gs;
    *smem = (size_t)p.rpc * p.W * 2 + ((size_t)16 * p.nv * p.rpi + 2 * p.W) * sizeof(float);
    return *smem <= 192 * 1024 && (long)p.B * p.n_slabs * GNC_CLUSTER <= 65535L * 8;
}

// ---- GEGLU: out = a * gelu(g) for proj = [a | g] (rows of 2F, exact erf GELU as F.gelu), and its gradient --------------------------
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ uint4 pack8(const float* r) {
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
    return o;
}

__global__ void geglu_fwd_kernel(const bf16* __restrict__ proj, long rows, int F, bf16* __restrict__ out) {
    const int vpr = F >> 3;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * vpr) return;
    const long r = i / vpr;
    const int c = (int)(i - r * vpr) << 3;
    float a[8], g[8];
    unpack8(*reinterpret_cast<const uint4*>(proj + r * 2 * F + c), a);
    unpack8(*reinterpret_cast<const uint4*>(proj + r * 2 * F + F + c), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_f(g[j]);
    *reinterpret_cast<uint4*>(out + r * F + c) = pack8(a);
}

__global__ void geglu_bwd_kernel(const bf16* __restrict__ proj, const bf16* __restrict__ dy, long rows, int F, bf16* __restrict__ dproj) {
    const int vpr = F >> 3;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * vpr) return;
    const long r = i / vpr;
    const int c = (int)(i - r * vpr) << 3;
    float a[8], g[8], d[8], da[8], dg[8];
    unpack8(*reinterpret_cast<const uint4*>(proj + r * 2 * F + c), a);
    unpack8(*reinterpret_cast<const uint4*>(proj + r * 2 * F + F + c), g);
    unpack8(*reinterpret_cast<const uint4*>(dy + r * F + c), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) { da[j] = d[j] * gelu_f(g[j]); dg[j] = d[j] * a[j] * gelu_grad_f(g[j]); }
    *reinterpret_cast<uint4*>(dproj + r * 2 * F + c) = pack8(da);
    *reinterpret_cast<uint4*>(dproj + r * 2 * F + F + c) = pack8(dg);
}

// out = a + b + bias[c] over (rows, C) bf16 (a residual add with the producing convolution's bias folded in)
__global__ void add_bias_residual_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, const bf16* __restrict__ bias, long nvec, int C,
                                         bf16* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    const int c = (int)((i << 3) % C);
    float x[8], y[8], z[8];
    unpack8(*reinterpret_cast<const uint4*>(a + (i << 3)), x);
    unpack8(*reinterpret_cast<const uint4*>(b + (i << 3)), y);
    unpack8(*reinterpret_cast<const uint4*>(bias + c), z);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (x[j] + y[j]) + z[j];
    *reinterpret_cast<uint4*>(out + (i << 3)) = pack8(x);
}

// ---- LayerNorm over the channel dimension of (rows, C) bf16 token matrices (48 per UNet evaluation) -------------------------------------
// One warp per row, the row held in registers (<= 5 vectors of 8 channels per lane: C <= 1280), mean then centred variance (two passes over
// registers, no E[x^2] - mean^2 cancellation), fp32 mean / rstd saved for the backward (which stays torch's native_layer_norm_backward).
template <int VPL>
__global__ void __launch_bounds__(256) layer_norm_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                                             long rows, int C, float eps, bf16* __restrict__ y, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out) {
    const int lane = threadIdx.x & 31;
    const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const int nvec = C >> 3;
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
            unpack8(*reinterpret_cast<const uint4*>(x + row * C + v * 8), f[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[i][j];
        }
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[i][j] - mean; q += d * d; }
        }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)C + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nvec) {
            float ga[8], be[8];
            unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), ga);
            unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), be);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[i][j] = (f[i][j] - mean) * rstd * ga[j] + be[j];
            *reinterpret_cast<uint4*>(y + row * C + v * 8) = pack8(f[i]);
        }
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
    const long need = (long)p.B * p.n_chunks * p.G * 2 + (long)p.B * p.G * 2;
    if (need > ws_floats) return set_error(GD_ERR_INVALID, "group norm: workspace of %ld floats < %ld", ws_floats, need);
    p.red = p.partial + (long)p.B * p.n_chunks * p.G * 2;
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

static int g_gn_cluster = 1;    // gd_group_norm_config: 1 = one-launch cluster kernel where the shape fits (default), 0 = always two launches

}  // namespace gd

using namespace gd;

extern "C" {

// test / tuning knob (process-wide): selects the forward scheme of gd_group_norm_nhwc_fwd
int gd_group_norm_config(int use_cluster_kernel) {
    g_gn_cluster = use_cluster_kernel ? 1 : 0;
    return GD_OK;
}

// y = silu?(group_norm(x + pre_bias[b, c])) for x (B, HW, C) bf16 channels-last (pre_bias (B, C) bf16 or NULL); gamma / beta (C) bf16 (w_is_bf16 = 1) or fp32.  stats (B, G, 2) receives
// (mean, rstd) (also the backward's input).  workspace: >= gd_group_norm_nhwc_workspace(B, HW, C, G) floats.  counters: >= B unsigned ints
// that are ZERO on entry (they are zero again on exit: allocate and clear once, reuse for every call on the same stream).
int gd_group_norm_nhwc_fwd(const void* x, const void* pre_bias, const void* gamma, const void* beta, int w_is_bf16, int B, int HW, int C, int G, float eps, int silu,
                           float* workspace, long workspace_floats, unsigned* counters, float* stats, void* y, void* stream) {
    GD_CHECK_ARG(x && gamma && beta && workspace && counters && stats && y && B > 0 && HW > 0 && C > 0 && G > 0 && G <= 32);
    if (C % 8 != 0 || C % G != 0) return set_error(GD_ERR_UNSUPPORTED, "group norm: C = %d must be a multiple of 8 and of G = %d", C, G);
    GnParams p;
    p.x = (const bf16*)x; p.pre_bias = (const bf16*)pre_bias; p.dy = nullptr; p.gamma = gamma; p.beta = beta; p.w_bf16 = w_is_bf16; p.partial = workspace; p.stats = stats;
    p.counter = counters; p.out = (bf16*)y; p.B = B; p.HW = HW; p.C = C; p.G = G; p.eps = eps; p.silu = silu;
    GnClusterParams c;
    c.x = p.x; c.pre_bias = p.pre_bias; c.gamma = gamma; c.beta = beta; c.w_bf16 = w_is_bf16; c.stats = stats; c.out = p.out;
    c.B = B; c.HW = HW; c.C = C; c.G = G; c.eps = eps; c.silu = silu;
    size_t smem = 0;
    if (g_gn_cluster && gn_cluster_plan(c, &smem)) {
        static size_t configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(gn_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(192 * 1024));
            if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            configured = 192 * 1024;
        }
        const int threads = (c.nv * c.rpi + 31) / 32 * 32 < 64 ? 64 : (c.nv * c.rpi + 31) / 32 * 32;   // >= gs * 8 (<= 256 only if gs <= 8) -- see below
        gn_cluster_fwd_kernel<<<B * c.n_slabs * GNC_CLUSTER, threads < c.gs * 8 ? (c.gs * 8 + 31) / 32 * 32 : threads, smem, (cudaStream_t)stream>>>(c);
        GD_CHECK_LAUNCH();
        return GD_OK;
    }
    return gn_run<0>(p, workspace_floats, (cudaStream_t)stream);
}

// dx of the above given dy (B, HW, C) bf16 and the forward's stats; gradients of gamma / beta are not produced (the body's weights
// are frozen in the edit loop: optimization.py:213-219 differentiates w.r.t. the latent and the context only).
int gd_group_norm_nhwc_bwd(const void* x, const void* pre_bias, const void* dy, const void* gamma, const void* beta, int w_is_bf16, const float* stats, int B, int HW,
                           int C, int G, int silu, float* workspace, long workspace_floats, unsigned* counters, void* dx, void* stream) {
    GD_CHECK_ARG(x && dy && gamma && beta && stats && workspace && counters && dx && B > 0 && HW > 0 && C > 0 && G > 0 && G <= 32);
    if (C % 8 != 0 || C % G != 0) return set_error(GD_ERR_UNSUPPORTED, "group norm: C = %d must be a multiple of 8 and of G = %d", C, G);
    GnParams p;
    p.x = (const bf16*)x; p.pre_bias = (const bf16*)pre_bias; p.dy = (const bf16*)dy; p.gamma = gamma; p.beta = beta; p.w_bf16 = w_is_bf16; p.partial = workspace;
    p.stats = const_cast<float*>(stats); p.counter = counters; p.out = (bf16*)dx; p.B = B; p.HW = HW; p.C = C; p.G = G; p.eps = 0.f; p.silu = silu;
    return gn_run<1>(p, workspace_floats, (cudaStream_t)stream);
}

// floats of workspace the two entry points above need (an upper bound that does not depend on the plan's details)
int gd_group_norm_nhwc_workspace(int B, int HW, int C, int G) {
    (void)HW; (void)C;
    return (2 * 148 + B) * G * 2 + B * G * 2;
}

// GEGLU of the feed-forward blocks: out (rows, F) = proj[:, :F] * gelu(proj[:, F:]) for proj (rows, 2F) bf16, F % 8 == 0 (torch: chunk + gelu +
// mul on strided halves = 2 non-vectorised kernels), and dproj given dy.
int gd_geglu_fwd(const void* proj, long rows, int F, void* out, void* stream) {
    GD_CHECK_ARG(proj && out && rows > 0 && F > 0 && (F % 8) == 0);
    geglu_fwd_kernel<<<ceil_div(rows * (F / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)proj, rows, F, (bf16*)out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}
int gd_geglu_bwd(const void* proj, const void* dy, long rows, int F, void* dproj, void* stream) {
    GD_CHECK_ARG(proj && dy && dproj && rows > 0 && F > 0 && (F % 8) == 0);
    geglu_bwd_kernel<<<ceil_div(rows * (F / 8), 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)proj, (const bf16*)dy, rows, F, (bf16*)dproj);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// out = a + b + bias[c] for a, b, out (rows, C) bf16 channels-last, bias (C) bf16, C % 8 == 0
int gd_add_bias_residual(const void* a, const void* b, const void* bias, long rows, int C, void* out, void* stream) {
    GD_CHECK_ARG(a && b && bias && out && rows > 0 && C > 0 && (C % 8) == 0);
    const long nvec = rows * (C / 8);
    add_bias_residual_kernel<<<ceil_div(nvec, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, (const bf16*)bias, nvec, C, (bf16*)out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// y = layer_norm(x) over the last dimension of x (rows, C) bf16, gamma / beta (C) bf16, C % 8 == 0, C <= 1280; mean, rstd (rows) fp32 out
// (the operands of aten::native_layer_norm_backward).
int gd_layer_norm_fwd(const void* x, const void* gamma, const void* beta, long rows, int C, float eps, void* y, float* mean, float* rstd,
                      void* stream) {
    GD_CHECK_ARG(x && gamma && beta && y && mean && rstd && rows > 0 && C > 0);
    if ((C % 8) != 0 || C > 1280) return set_error(GD_ERR_UNSUPPORTED, "layer norm: C = %d must be a multiple of 8 and <= 1280", C);
    const int blocks = ceil_div(rows * 32, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int vpl = (C / 8 + 31) / 32;
#define GD_LN(V) layer_norm_fwd_kernel<V><<<blocks, 256, 0, st>>>((const bf16*)x, (const bf16*)gamma, (const bf16*)beta, rows, C, eps, (bf16*)y, mean, rstd)
    if (vpl <= 2) GD_LN(2);
    else if (vpl <= 3) GD_LN(3);
    else GD_LN(5);
#undef GD_LN
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
