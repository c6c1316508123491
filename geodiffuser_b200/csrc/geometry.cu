// geodiffuser_b200/csrc/geometry.cu
//
// Subsystem (1) of the hot path: depth unprojection -> rigid transform -> reprojection into the
// correspondence field, per-resolution resize, mask algebra, forward-splat index (the bit-exact
// integer artefact) and the alpha-composite gather that consumes it.
//
// Reference behaviour being replaced (paths relative to /root/reference/GeoDiffuser/utils/):
//   pixel2cam                      warp_utils.py:738-747
//   object centroid                warp_utils.py:426-427
//   cam2pixel_vanilla              warp_utils.py:599-643
//   T.Resize(BILINEAR, aa=False)   generic_torch.py:156-207
//   process_and_cache_masks        attention_processors.py:319-373
//   rasterize_points (pytorch3d)   warp_utils.py:111-113
//   alpha_composite (pytorch3d)    warp_utils.py:131-176
//   get_mesh / splatter_mesh       warp_utils.py:364-399, 235-298
//   torch_erode / torch_dilate     generic_torch.py:210-235
//
// Everything here is integer / index work preceded by a short fp32 pipeline whose IEEE operation
// order is part of the contract (bit-exact against oracle/geom_cpu.c + oracle/pt3d_cpu.c), so the
// file is compiled with -fmad=false and every fused multiply-add is spelled __fmaf_rn explicitly.
// These kernels are HBM / latency bound (KB..MB per launch): coalesced loads, no tensor cores.
#include "common.cuh"

namespace gd {

struct Mat3 { float m[9]; };
struct Mat34 { float m[12]; };

__device__ __forceinline__ float dot3_fma(const float* a, float b0, float b1, float b2) {
    float acc = __fmul_rn(a[0], b0);
    acc = __fmaf_rn(a[1], b1, acc);
    acc = __fmaf_rn(a[2], b2, acc);
    return acc;
}

// cam (3,H,W) = (Kinv @ [u,v,1]) * depth
__global__ void pixel2cam_kernel(const float* __restrict__ depth, int H, int W, Mat3 Kinv, float* __restrict__ cam) {
    const long hw = (long)H * W;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const int v = (int)(p / W), u = (int)(p % W);
    const float d = depth[p];
#pragma unroll
    for (int r = 0; r < 3; ++r) cam[r * hw + p] = __fmul_rn(dot3_fma(Kinv.m + 3 * r, (float)u, (float)v, 1.0f), d);
}

// canonical centroid: double accumulation left-to-right inside a row (one thread per row), then the
// row partials top-to-bottom (thread 0).  out4 = {mean_x, mean_y, mean_z, count}
// (197 us at 512^2: every thread walks its own row, stride W between neighbouring lanes.  Two single-block forms that stage the rows through shared
//  memory -- 32 rows at a time, and all rows in 16-column strips -- kept the order but measured 426 and 253 us: one SM cannot pull 4 MB faster, and a
//  multi-block form needs a partials workspace the entry point does not have.  Once per edit, 0.04 % of its time: left as it is.)
__global__ void centroid_kernel(const float* __restrict__ cam, const float* __restrict__ mask, int H, int W,
                                float* __restrict__ out4) {
    extern __shared__ double part[];  // H * 4
    const long hw = (long)H * W;
    const int v = threadIdx.x;
    if (v < H) {
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, c = 0.0;
        for (int u = 0; u < W; ++u) {
            const long p = (long)v * W + u;
            if (mask[p] >= 0.5f) {
                r0 += (double)cam[p];
                r1 += (double)cam[hw + p];
                r2 += (double)cam[2 * hw + p];
                c += 1.0;
            }
        }
        part[4 * v + 0] = r0; part[4 * v + 1] = r1; part[4 * v + 2] = r2; part[4 * v + 3] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, c = 0.0;
        for (int r = 0; r < H; ++r) { t0 += part[4 * r]; t1 += part[4 * r + 1]; t2 += part[4 * r + 2]; c += part[4 * r + 3]; }
        out4[0] = (float)(t0 / c); out4[1] = (float)(t1 / c); out4[2] = (float)(t2 / c); out4[3] = (float)c;
    }
}

// coords (H,W,3) = (x_norm, y_norm, Z)
__global__ void project_kernel(const float* __restrict__ cam, int H, int W, Mat34 Rt, Mat3 K, float* __restrict__ coords) {
    const long hw = (long)H * W;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const float c0 = cam[p], c1 = cam[hw + p], c2 = cam[2 * hw + p];
    float q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float rr[3] = {Rt.m[4 * r], Rt.m[4 * r + 1], Rt.m[4 * r + 2]};
        q[r] = __fadd_rn(dot3_fma(rr, c0, c1, c2), Rt.m[4 * r + 3]);
    }
    const float X = dot3_fma(K.m, q[0], q[1], q[2]);
    const float Y = dot3_fma(K.m + 3, q[0], q[1], q[2]);
    float Z = dot3_fma(K.m + 6, q[0], q[1], q[2]);
    if (Z < 1e-3f) Z = 1e-3f;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    coords[3 * p + 0] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(X, Z)), wm1), 1.0f);
    coords[3 * p + 1] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fdiv_rn(Y, Z)), hm1), 1.0f);
    coords[3 * p + 2] = Z;
}

// ---- bilinear (align_corners=False, antialias=False) ------------------------------------------
struct BilinearTap { int y0, y1, x0, x1; float w00, w01, w10, w11; };

__device__ __forceinline__ BilinearTap bilinear_tap(int oy, int ox, int Hin, int Win, int Hout, int Wout) {
    BilinearTap t;
    const float sh = __fdiv_rn((float)Hin, (float)Hout), sw = __fdiv_rn((float)Win, (float)Wout);
    float fy = __fsub_rn(__fmul_rn(sh, __fadd_rn((float)oy, 0.5f)), 0.5f);
    if (fy < 0.0f) fy = 0.0f;
    float fx = __fsub_rn(__fmul_rn(sw, __fadd_rn((float)ox, 0.5f)), 0.5f);
    if (fx < 0.0f) fx = 0.0f;
    t.y0 = (int)fy; t.y1 = t.y0 + ((t.y0 < Hin - 1) ? 1 : 0);
    t.x0 = (int)fx; t.x1 = t.x0 + ((t.x0 < Win - 1) ? 1 : 0);
    const float ly1 = __fsub_rn(fy, (float)t.y0), ly0 = __fsub_rn(1.0f, ly1);
    const float lx1 = __fsub_rn(fx, (float)t.x0), lx0 = __fsub_rn(1.0f, lx1);
    t.w00 = __fmul_rn(ly0, lx0); t.w01 = __fmul_rn(ly0, lx1); t.w10 = __fmul_rn(ly1, lx0); t.w11 = __fmul_rn(ly1, lx1);
    return t;
}
__device__ __forceinline__ float bilinear_apply(const BilinearTap& t, float p00, float p01, float p10, float p11) {
    float acc = __fadd_rn(__fmul_rn(t.w00, p00), __fmul_rn(t.w01, p01));
    acc = __fadd_rn(acc, __fmul_rn(t.w10, p10));
    acc = __fadd_rn(acc, __fmul_rn(t.w11, p11));
    return acc;
}

// generic strided resize: src[c*sc + y*sy + x*sx] -> dst[c*dc + y*dy + x*dx]
__global__ void resize_bilinear_kernel(const float* __restrict__ src, int C, int Hin, int Win, long sc, long sy, long sx,
                                       float* __restrict__ dst, int Hout, int Wout, long dc, long dy, long dx, int c_fastest) {
    const long total = (long)C * Hout * Wout;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c, oy, ox;
    if (c_fastest) { c = (int)(i % C); const long r = i / C; ox = (int)(r % Wout); oy = (int)(r / Wout); }
    else { ox = (int)(i % Wout); const long r = i / Wout; oy = (int)(r % Hout); c = (int)(r / Hout); }
    const BilinearTap t = bilinear_tap(oy, ox, Hin, Win, Hout, Wout);
    const float* s = src + (long)c * sc;
    dst[(long)c * dc + (long)oy * dy + (long)ox * dx] =
        bilinear_apply(t, s[t.y0 * sy + t.x0 * sx], s[t.y0 * sy + t.x1 * sx], s[t.y1 * sy + t.x0 * sx], s[t.y1 * sy + t.x1 * sx]);
}

__device__ __forceinline__ float bin05(float v) { return v > 0.5f ? 1.0f : 0.0f; }

// out (6,S,S): mask_new_warped, mask_warp, amodal_mask, mask_intersection, mask_1_empty, mask_wo_edit
__global__ void masks_build_kernel(const float* __restrict__ image_mask, const float* __restrict__ warped,
                                   const float* __restrict__ amodal, int Hin, int S, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * S) return;
    const int oy = i / S, ox = i % S;
    const BilinearTap t = bilinear_tap(oy, ox, Hin, Hin, S, S);
    const long a = (long)t.y0 * Hin + t.x0, b = (long)t.y0 * Hin + t.x1, c = (long)t.y1 * Hin + t.x0, d = (long)t.y1 * Hin + t.x1;
    const float m_src = bilinear_apply(t, bin05(image_mask[a]), bin05(image_mask[b]), bin05(image_mask[c]), bin05(image_mask[d]));
    const float m_warp = warped ? bilinear_apply(t, warped[a], warped[b], warped[c], warped[d]) : 0.0f;
    const float am = amodal ? bilinear_apply(t, amodal[a], amodal[b], amodal[c], amodal[d]) : 0.0f;
    const float m_amodal = bin05(__fsub_rn(am, m_warp));
    const float m_inter = bin05(__fmul_rn(__fadd_rn(m_warp, m_amodal), m_src));
    const float m_inp = bin05(__fsub_rn(m_src, m_inter));
    const float m_bg = bin05(__fsub_rn(1.0f, __fadd_rn(m_inp, m_warp)));
    const int n = S * S;
    out[i] = m_warp; out[n + i] = m_src; out[2 * n + i] = m_amodal; out[3 * n + i] = m_inter; out[4 * n + i] = m_inp;
    out[5 * n + i] = m_bg;
}

// ---- forward-splat index ------------------------------------------------------------------------
__device__ __forceinline__ float pix_to_ndc(int i, int S) {
    // pytorch3d PixToNonSquareNdc: -offset + (range * i + offset) / S with range = 2, offset = 1
    return __fadd_rn(-1.0f, __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, (float)i), 1.0f), (float)S));
}

constexpr int SPLAT_TILE = 16;
constexpr int SPLAT_CHUNK = 1024;
constexpr int SPLAT_KMAX = 16;

// One CTA per 16x16 output tile; the point list is streamed in chunks, culled against the tile's
// NDC bounding box into shared memory, and each pixel keeps its K nearest in (z, packed index)
// lexicographic order.  Candidate arrival order is irrelevant because the comparator is total.
__global__ void __launch_bounds__(SPLAT_TILE * SPLAT_TILE)
splat_index_kernel(const float* __restrict__ coords, int P, int S, float radius, int K, int* __restrict__ idx,
                   float* __restrict__ zbuf, float* __restrict__ dist2) {
    __shared__ float4 cand[SPLAT_CHUNK];
    __shared__ int ncand;
    const int b = blockIdx.z;
    const int tx = threadIdx.x % SPLAT_TILE, ty = threadIdx.x / SPLAT_TILE;
    const int xi = blockIdx.x * SPLAT_TILE + tx, yi = blockIdx.y * SPLAT_TILE + ty;
    const bool live = xi < S && yi < S;
    const float r2 = __fmul_rn(radius, radius);
    const float cxp = pix_to_ndc(S - 1 - min(xi, S - 1), S), cyp = pix_to_ndc(S - 1 - min(yi, S - 1), S);
    // tile bounds (pixel centres decrease with the index); padded conservatively
    const int x_lo = blockIdx.x * SPLAT_TILE, x_hi = min(x_lo + SPLAT_TILE, S) - 1;
    const int y_lo = blockIdx.y * SPLAT_TILE, y_hi = min(y_lo + SPLAT_TILE, S) - 1;
    const float pad = radius * 1.001f + 1e-6f;
    const float bx0 = pix_to_ndc(S - 1 - x_hi, S) - pad, bx1 = pix_to_ndc(S - 1 - x_lo, S) + pad;
    const float by0 = pix_to_ndc(S - 1 - y_hi, S) - pad, by1 = pix_to_ndc(S - 1 - y_lo, S) + pad;

    float kz[SPLAT_KMAX], kd[SPLAT_KMAX];
    int ki[SPLAT_KMAX];
    int n = 0;
    const float* pts = coords + (long)b * P * 3;
    if (threadIdx.x == 0) ncand = 0;
    __syncthreads();
    for (int base = 0; base < P; base += SPLAT_CHUNK) {
#pragma unroll
        for (int j = 0; j < SPLAT_CHUNK / (SPLAT_TILE * SPLAT_TILE); ++j) {
            const int p = base + j * (SPLAT_TILE * SPLAT_TILE) + threadIdx.x;
            if (p < P) {
                const float px = -pts[3 * p], py = -pts[3 * p + 1], pz = pts[3 * p + 2];   // warp_utils.py:90-91
                if (pz >= 0.0f && px >= bx0 && px <= bx1 && py >= by0 && py <= by1) {
                    const int s = atomicAdd(&ncand, 1);
                    cand[s] = make_float4(px, py, pz, __int_as_float(p));
                }
            }
        }
        __syncthreads();
        const int nc = ncand;
        if (live) {
            for (int c = 0; c < nc; ++c) {
                const float4 q = cand[c];
                const float dx = __fsub_rn(q.x, cxp), dy = __fsub_rn(q.y, cyp);
                const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                if (!(d2 < r2)) continue;
                const int gp = b * P + __float_as_int(q.w);
                int pos = n;
                while (pos > 0 && (kz[pos - 1] > q.z || (kz[pos - 1] == q.z && ki[pos - 1] > gp))) --pos;
                if (pos >= K) continue;
                const int last = (n < K) ? n : K - 1;
                for (int t = last; t > pos; --t) { kz[t] = kz[t - 1]; ki[t] = ki[t - 1]; kd[t] = kd[t - 1]; }
                kz[pos] = q.z; ki[pos] = gp; kd[pos] = d2;
                if (n < K) ++n;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) ncand = 0;
        __syncthreads();
    }
    if (live) {
        const long o = (((long)b * S + yi) * S + xi) * K;
        for (int k = 0; k < K; ++k) {
            const bool f = k < n;
            idx[o + k] = f ? ki[k] : -1;
            if (zbuf) zbuf[o + k] = f ? kz[k] : -1.0f;
            dist2[o + k] = f ? kd[k] : -1.0f;
        }
    }
}

// ---- alpha-composite gather ---------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// out[b,p,c] = post( half( sum_k cum * alpha_k * src[b, idx_k, c] ) ), optionally blended with src through
// `blend` (P): src*(1-m) + m*warped  (attention_processors.py:544).  idx/dist2: (Bi, P, K), Bi in {1, B}.
template <typename TIn, typename TOut>
__global__ void splat_composite_kernel(const TIn* __restrict__ src, long sb, long sp, long sc, const int* __restrict__ idx,
                                       const float* __restrict__ dist2, int Bi, int B, int P, int C, int K, float r2, float tau,
                                       const float* __restrict__ blend, int post, TOut* __restrict__ out, long ob, long op,
                                       long oc, int c_fastest) {
    const long total = (long)B * P * C;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int b, p, c;
    if (c_fastest) { c = (int)(i % C); const long r = i / C; p = (int)(r % P); b = (int)(r / P); }
    else { p = (int)(i % P); const long r = i / P; c = (int)(r % C); b = (int)(r / C); }
    const int bi = (Bi == 1) ? 0 : b;
    const int* pi = idx + ((long)bi * P + p) * K;
    const float* pd = dist2 + ((long)bi * P + p) * K;
    const TIn* s = src + (long)b * sb + (long)c * sc;
    float cum = 1.0f, acc = 0.0f;
    for (int k = 0; k < K; ++k) {
        const int n = pi[k];
        if (n < 0) continue;
        float a = __fsub_rn(1.0f, __fsqrt_rn(fminf(fmaxf(__fdiv_rn(pd[k], r2), 1e-3f), 1.0f)));  // warp_utils.py:131-140
        if (tau != 1.0f) a = powf(a, tau);
        const float f = ldf<TIn>(s + (long)(n - bi * P) * sp);
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(cum, a), f));
        cum = __fmul_rn(cum, __fsub_rn(1.0f, a));
    }
    float w = __half2float(__float2half_rn(acc));   // .to(torch.half)  warp_utils.py:176
    if (blend) {
        const float m = blend[p];
        const float q = ldf<TIn>(s + (long)p * sp);
        w = __fadd_rn(__fmul_rn(q, __fsub_rn(1.0f, m)), __fmul_rn(m, w));
    }
    if (post == 1) w = bin05(w);
    stf<TOut>(out + (long)b * ob + (long)p * op + (long)c * oc, w);
}

// Row form of the composite for channel-fastest sources sharing one index (the per-layer query warp: src (B=heads, P, C=head_dim), Bi = 1).
// A block takes PPB output pixels.  Phase 1: the K blend weights cum_k * alpha_k of each pixel -- they depend only on the pixel -- are formed
// once (same IEEE operations in the same order as splat_composite_kernel, so results are bit-identical) instead of once per (head, channel).
// Phase 2: one thread per (pixel, head, 8-channel vector): K independent 16-byte gathers, explicit mul / add (no contraction).
template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float* f);
template <> __device__ __forceinline__ void st8<float>(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = v;
}

constexpr int ROWS_MAXPPB = 8, ROWS_MAXK = 32;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(128) splat_composite_rows_kernel(const TIn* __restrict__ src, long sb, long sp, const int* __restrict__ idx,
                                                                   const float* __restrict__ dist2, int B, int P, int C, int K, float r2,
                                                                   float tau, const float* __restrict__ blend, int post, TOut* __restrict__ out,
                                                                   long ob, long op, int ppb) {
    __shared__ float s_w[ROWS_MAXPPB][ROWS_MAXK];
    __shared__ int s_n[ROWS_MAXPPB][ROWS_MAXK];
    __shared__ int s_cnt[ROWS_MAXPPB];
    const int p0 = blockIdx.x * ppb;
    // phase 1a: alpha of every (pixel, slot)
    for (int i = threadIdx.x; i < ppb * K; i += blockDim.x) {
        const int pl = i / K, k = i - pl * K, p = p0 + pl;
        int n = -1;
        float a = 0.0f;
        // a pixel whose blend mask is exactly 0 keeps its source value (q (1 - 0) + 0 * splat): its <= K gathers are skipped.  Nine tenths of the
        // pixels of a query warp lie outside the warped object mask, and the serial gather chain of the others set the launch time (12 us).
        if (p < P && !(blend != nullptr && post == 0 && blend[p] == 0.0f)) {
            n = idx[(long)p * K + k];
            if (n >= 0) {
                a = __fsub_rn(1.0f, __fsqrt_rn(fminf(fmaxf(__fdiv_rn(dist2[(long)p * K + k], r2), 1e-3f), 1.0f)));   // warp_utils.py:131-140
                if (tau != 1.0f) a = powf(a, tau);
            }
        }
        s_w[pl][k] = a;
        s_n[pl][k] = n;
    }
    __syncthreads();
    // phase 1b: front-to-back weights, empty slots (n < 0) skipped as in the reference loop -> valid slots compacted to the front
    if (threadIdx.x < ppb) {
        const int pl = threadIdx.x;
        float cum = 1.0f;
        int cnt = 0;
        for (int k = 0; k < K; ++k) {
            const int n = s_n[pl][k];
            if (n < 0) continue;
            const float a = s_w[pl][k];
            s_w[pl][cnt] = __fmul_rn(cum, a);
            s_n[pl][cnt] = n;
            cum = __fmul_rn(cum, __fsub_rn(1.0f, a));
            ++cnt;
        }
        s_cnt[pl] = cnt;
    }
    __syncthreads();
    // phase 2
    const int vpr = C >> 3;                 // 8-channel vectors per (pixel, head) row
    const int per_pixel = B * vpr;
    for (int t = threadIdx.x; t < ppb * per_pixel; t += blockDim.x) {
        const int pl = t / per_pixel, r = t - pl * per_pixel, p = p0 + pl;
        if (p >= P) continue;
        const int b = r / vpr, c = (r - b * vpr) << 3;
        const TIn* s = src + (long)b * sb + c;
        const int cnt = s_cnt[pl];
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        for (int k = 0; k < cnt; ++k) {
            float f[8];
            ld8<TIn>(s + (long)s_n[pl][k] * sp, f);
            const float wk = s_w[pl][k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(wk, f[j]));
        }
        float q[8];
        const float m = blend ? blend[p] : 0.0f;
        if (blend) ld8<TIn>(s + (long)p * sp, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float w = __half2float(__float2half_rn(acc[j]));   // .to(torch.half)  warp_utils.py:176
            if (blend) w = __fadd_rn(__fmul_rn(q[j], __fsub_rn(1.0f, m)), __fmul_rn(m, w));
            if (post == 1) w = bin05(w);
            acc[j] = w;
        }
        st8<TOut>(out + (long)b * ob + (long)p * op + c, acc);
    }
}

// ---- amodal mesh coverage -----------------------------------------------------------------------
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float vx = __fsub_rn(bx, ax), vy = __fsub_rn(by, ay);
    const float l2 = __fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy));
    if (l2 <= 1e-8f) { const float dx = __fsub_rn(px, bx), dy = __fsub_rn(py, by); return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)); }
    float t = __fdiv_rn(__fadd_rn(__fmul_rn(vx, __fsub_rn(px, ax)), __fmul_rn(vy, __fsub_rn(py, ay))), l2);
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f;
    const float qx = __fadd_rn(ax, __fmul_rn(t, vx)), qy = __fadd_rn(ay, __fmul_rn(t, vy));
    const float dx = __fsub_rn(px, qx), dy = __fsub_rn(py, qy);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__device__ void raster_tri(const float* a, const float* b, const float* c, int S, float blur, float* out) {
    if (a[2] < 1e-8f || b[2] < 1e-8f || c[2] < 1e-8f) return;
    const float xmin = fminf(a[0], fminf(b[0], c[0])), xmax = fmaxf(a[0], fmaxf(b[0], c[0]));
    const float ymin = fminf(a[1], fminf(b[1], c[1])), ymax = fmaxf(a[1], fmaxf(b[1], c[1]));
    if (!(xmin == xmin) || !(ymin == ymin) || !(xmax == xmax) || !(ymax == ymax)) return;
    const float fS = (float)S;
    const float fi0 = (float)(S - 1) - ((xmax + 1.0f) * fS - 1.0f) * 0.5f, fi1 = (float)(S - 1) - ((xmin + 1.0f) * fS - 1.0f) * 0.5f;
    const float fj0 = (float)(S - 1) - ((ymax + 1.0f) * fS - 1.0f) * 0.5f, fj1 = (float)(S - 1) - ((ymin + 1.0f) * fS - 1.0f) * 0.5f;
    if (fi1 < -2 || fi0 > S + 1 || fj1 < -2 || fj0 > S + 1) return;
    int x0 = (int)floorf(fmaxf(fi0, -4.f)) - 2, x1 = (int)ceilf(fminf(fi1, fS + 4.f)) + 2;
    int y0 = (int)floorf(fmaxf(fj0, -4.f)) - 2, y1 = (int)ceilf(fminf(fj1, fS + 4.f)) + 2;
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, S - 1); y1 = min(y1, S - 1);
    const float area = edge_fn(c[0], c[1], a[0], a[1], b[0], b[1]);
    if (fabsf(area) <= 1e-8f) return;
    for (int yi = y0; yi <= y1; ++yi)
        for (int xi = x0; xi <= x1; ++xi) {
            const float px = pix_to_ndc(S - 1 - xi, S), py = pix_to_ndc(S - 1 - yi, S);
            const float w0 = __fdiv_rn(edge_fn(px, py, b[0], b[1], c[0], c[1]), area);
            const float w1 = __fdiv_rn(edge_fn(px, py, c[0], c[1], a[0], a[1]), area);
            const float w2 = __fdiv_rn(edge_fn(px, py, a[0], a[1], b[0], b[1]), area);
            bool hit = (w0 > 0.0f && w1 > 0.0f && w2 > 0.0f);
            if (!hit) {
                const float d = fminf(seg_dist2(px, py, a[0], a[1], b[0], b[1]),
                                      fminf(seg_dist2(px, py, b[0], b[1], c[0], c[1]), seg_dist2(px, py, c[0], c[1], a[0], a[1])));
                hit = d < blur;
            }
            if (hit) out[(long)yi * S + xi] = 1.0f;
        }
}

// one thread per 2x2 pixel quad fully inside the mask: triangles (tl,tr,bl) and (bl,tr,br)
__global__ void mesh_mask_kernel(const float* __restrict__ coords, const float* __restrict__ mask, int H, int W, float blur,
                                 float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)(H - 1) * (W - 1)) return;
    const int v = (int)(i / (W - 1)), u = (int)(i % (W - 1));
    const long tl = (long)v * W + u, tr = tl + 1, bl = tl + W, br = bl + 1;
    if (!(mask[tl] >= 0.5f && mask[tr] >= 0.5f && mask[bl] >= 0.5f && mask[br] >= 0.5f)) return;
    float vt[4][3];
    const long ids[4] = {tl, tr, bl, br};
#pragma unroll
    for (int k = 0; k < 4; ++k) { vt[k][0] = -coords[3 * ids[k]]; vt[k][1] = -coords[3 * ids[k] + 1]; vt[k][2] = coords[3 * ids[k] + 2]; }
    raster_tri(vt[0], vt[1], vt[2], H, blur, out);
    raster_tri(vt[2], vt[1], vt[3], H, blur, out);
}

// mode 0: erode (window sum == k*k), mode 1: dilate (window sum >= 1); zero padding
__global__ void morph_kernel(const float* __restrict__ src, int B, int H, int W, int k, int mode, float* __restrict__ dst) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const float* s = src + (i / ((long)H * W)) * (long)H * W;
    const int r = k / 2;
    float sum = 0.0f;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) sum += s[(long)yy * W + xx];
        }
    dst[i] = mode == 0 ? ((sum == (float)(k * k)) ? 1.0f : 0.0f) : ((sum >= 1.0f) ? 1.0f : 0.0f);
}

}  // namespace gd

// ================================================================================================
// C ABI
// ================================================================================================
using namespace gd;

extern "C" {

int gd_corr_pixel2cam(const float* depth, const float* mask, int H, int W, const float* Kinv9_host, float* cam,
                      float* centroid4, void* stream) {
    GD_CHECK_ARG(depth && mask && Kinv9_host && cam && centroid4 && H > 0 && W > 0 && H <= 1024);
    cudaStream_t st = (cudaStream_t)stream;
    Mat3 Ki;
    for (int i = 0; i < 9; ++i) Ki.m[i] = Kinv9_host[i];
    const long hw = (long)H * W;
    pixel2cam_kernel<<<ceil_div(hw, 256), 256, 0, st>>>(depth, H, W, Ki, cam);
    GD_CHECK_LAUNCH();
    const int threads = ((H + 31) / 32) * 32;
    centroid_kernel<<<1, threads, (size_t)H * 4 * sizeof(double), st>>>(cam, mask, H, W, centroid4);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_corr_project(const float* cam, int H, int W, const float* Rt12_host, const float* K9_host, float* coords, void* stream) {
    GD_CHECK_ARG(cam && Rt12_host && K9_host && coords && H > 0 && W > 0);
    Mat34 Rt; Mat3 K;
    for (int i = 0; i < 12; ++i) Rt.m[i] = Rt12_host[i];
    for (int i = 0; i < 9; ++i) K.m[i] = K9_host[i];
    const long hw = (long)H * W;
    project_kernel<<<ceil_div(hw, 256), 256, 0, (cudaStream_t)stream>>>(cam, H, W, Rt, K, coords);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_resize_bilinear(const float* src, int C, int Hin, int Win, int channels_last, float* dst, int Hout, int Wout, void* stream) {
    GD_CHECK_ARG(src && dst && C > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0);
    const long total = (long)C * Hout * Wout;
    if (channels_last)
        resize_bilinear_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, C, Hin, Win, 1, (long)Win * C, C, dst, Hout,
                                                                                   Wout, 1, (long)Wout * C, C, 1);
    else
        resize_bilinear_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(src, C, Hin, Win, (long)Hin * Win, Win, 1, dst,
                                                                                   Hout, Wout, (long)Hout * Wout, Wout, 1, 0);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_masks_build(const float* image_mask, const float* mask_new_warped, const float* amodal_mask, int Hin, int S, float* out6,
                   void* stream) {
    GD_CHECK_ARG(image_mask && out6 && Hin > 0 && S > 0);
    masks_build_kernel<<<ceil_div((long)S * S, 256), 256, 0, (cudaStream_t)stream>>>(image_mask, mask_new_warped, amodal_mask, Hin, S, out6);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_splat_index(const float* coords, int B, int S, float radius_ndc, int K, int* idx, float* zbuf, float* dist2, void* stream) {
    GD_CHECK_ARG(coords && idx && dist2 && B > 0 && S > 0 && K > 0 && K <= SPLAT_KMAX && radius_ndc > 0.0f);
    GD_CHECK_ARG((long)B * S * S < (1L << 31));
    dim3 grid(ceil_div(S, SPLAT_TILE), ceil_div(S, SPLAT_TILE), B);
    splat_index_kernel<<<grid, SPLAT_TILE * SPLAT_TILE, 0, (cudaStream_t)stream>>>(coords, S * S, S, radius_ndc, K, idx, zbuf, dist2);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// dtype enums: 0 = fp32, 1 = bf16.  layout: 0 = channel-first (B,C,P), 1 = channel-last (B,P,C)
int gd_splat_composite(const void* src, int src_dtype, int layout, const int* idx, const float* dist2, int Bi, int B, int P, int C,
                       int K, float r2, float tau, const float* blend_mask, int post, void* out, int out_dtype, void* stream) {
    GD_CHECK_ARG(src && idx && dist2 && out && (Bi == 1 || Bi == B) && B > 0 && P > 0 && C > 0 && K > 0);
    GD_CHECK_ARG((src_dtype == 0 || src_dtype == 1) && (out_dtype == 0 || out_dtype == 1) && (layout == 0 || layout == 1));
    const long total = (long)B * P * C;
    const long sb = (long)P * C, sp = layout ? C : 1, sc = layout ? 1 : P;
    cudaStream_t st = (cudaStream_t)stream;
    if (layout == 1 && Bi == 1 && K <= ROWS_MAXK && (C % 8) == 0) {   // per-layer query warp: blend weights formed once per pixel
        const int ppb = P >= 2048 ? 8 : 2;
        const int gr = ceil_div(P, ppb);
#define GD_LAUNCH_ROWS(TI, TO) \
    splat_composite_rows_kernel<TI, TO><<<gr, 128, 0, st>>>((const TI*)src, sb, sp, idx, dist2, B, P, C, K, r2, tau, blend_mask, post, (TO*)out, sb, sp, ppb)
        if (src_dtype == 0 && out_dtype == 0) GD_LAUNCH_ROWS(float, float);
        else if (src_dtype == 0 && out_dtype == 1) GD_LAUNCH_ROWS(float, __nv_bfloat16);
        else if (src_dtype == 1 && out_dtype == 0) GD_LAUNCH_ROWS(__nv_bfloat16, float);
        else GD_LAUNCH_ROWS(__nv_bfloat16, __nv_bfloat16);
#undef GD_LAUNCH_ROWS
        GD_CHECK_LAUNCH();
        return GD_OK;
    }
    const int g = ceil_div(total, 256);
#define GD_LAUNCH_COMPOSITE(TI, TO)                                                                                         \
    splat_composite_kernel<TI, TO><<<g, 256, 0, st>>>((const TI*)src, sb, sp, sc, idx, dist2, Bi, B, P, C, K, r2, tau, blend_mask, \
                                                      post, (TO*)out, sb, sp, sc, layout)
    if (src_dtype == 0 && out_dtype == 0) GD_LAUNCH_COMPOSITE(float, float);
    else if (src_dtype == 0 && out_dtype == 1) GD_LAUNCH_COMPOSITE(float, __nv_bfloat16);
    else if (src_dtype == 1 && out_dtype == 0) GD_LAUNCH_COMPOSITE(__nv_bfloat16, float);
    else GD_LAUNCH_COMPOSITE(__nv_bfloat16, __nv_bfloat16);
#undef GD_LAUNCH_COMPOSITE
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// The per-layer query warp on one (B = heads, P, C = head_dim) slab with explicit element strides: src / out addressed as
// base + b * head_stride + p * row_stride + c (src_strides / out_strides = {row, head}, HOST arrays; NULL = contiguous (B, P, C)).  One index
// (P, K) shared by all heads.  Same arithmetic as gd_splat_composite (layout 1), bit for bit.
int gd_splat_composite_rows(const void* src, int src_dtype, const long* src_strides, const int* idx, const float* dist2, int B, int P, int C,
                            int K, float r2, float tau, const float* blend_mask, int post, void* out, int out_dtype, const long* out_strides,
                            void* stream) {
    GD_CHECK_ARG(src && idx && dist2 && out && B > 0 && P > 0 && C > 0 && (C % 8) == 0 && K > 0 && K <= ROWS_MAXK);
    GD_CHECK_ARG((src_dtype == 0 || src_dtype == 1) && (out_dtype == 0 || out_dtype == 1));
    const long sp = src_strides ? src_strides[0] : C, sb = src_strides ? src_strides[1] : (long)P * C;
    const long op = out_strides ? out_strides[0] : C, ob = out_strides ? out_strides[1] : (long)P * C;
    GD_CHECK_ARG((sp % 8) == 0 && (sb % 8) == 0 && (op % 8) == 0 && (ob % 8) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int ppb = P >= 2048 ? 8 : 2;
    const int gr = ceil_div(P, ppb);
#define GD_LAUNCH_ROWS(TI, TO) \
    splat_composite_rows_kernel<TI, TO><<<gr, 128, 0, st>>>((const TI*)src, sb, sp, idx, dist2, B, P, C, K, r2, tau, blend_mask, post, (TO*)out, ob, op, ppb)
    if (src_dtype == 0 && out_dtype == 0) GD_LAUNCH_ROWS(float, float);
    else if (src_dtype == 0 && out_dtype == 1) GD_LAUNCH_ROWS(float, __nv_bfloat16);
    else if (src_dtype == 1 && out_dtype == 0) GD_LAUNCH_ROWS(__nv_bfloat16, float);
    else GD_LAUNCH_ROWS(__nv_bfloat16, __nv_bfloat16);
#undef GD_LAUNCH_ROWS
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_mesh_mask(const float* coords, const float* mask, int H, int W, float blur, float* out, void* stream) {
    GD_CHECK_ARG(coords && mask && out && H == W && H > 1);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)H * W * sizeof(float), st);
    if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    mesh_mask_kernel<<<ceil_div((long)(H - 1) * (W - 1), 128), 128, 0, st>>>(coords, mask, H, W, blur, out);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_morph(const float* src, int B, int H, int W, int kernel, int mode, float* dst, void* stream) {
    GD_CHECK_ARG(src && dst && B > 0 && H > 0 && W > 0 && kernel > 0 && (kernel & 1) && (mode == 0 || mode == 1));
    morph_kernel<<<ceil_div((long)B * H * W, 256), 256, 0, (cudaStream_t)stream>>>(src, B, H, W, kernel, mode, dst);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
