// geodiffuser_b200/csrc/losses.cu
//
// Attention-map loss terms of the latent optimisation, forward AND gradient in the same pass (every term is
// piecewise linear in replace_out, so d(loss)/d(replace_out) is emitted while the sums are formed):
//   background_preservation_loss   attention_processors.py:231-246   ("sim")
//   object_placement_loss_geodiff  attention_processors.py:283-287   ("movement")
//   amodal_loss_geodiff            attention_processors.py:289-305 + attention_sharing.py:68-105 + generic_torch.py:145-154
//   get_smoothness_loss            loss.py:22-41
//   removal_loss_geodiff           attention_processors.py:248-280   (consumes corr_gemm.cu's masked arg-max partials)
// plus the output blend of attention_processors.py:617-624 / 922-925.  HBM-bound elementwise work: coalesced along
// the feature dimension, deterministic two-stage reductions (no float atomics), no tensor cores.
#include "common.cuh"

namespace gd {

__device__ __forceinline__ float sgnf(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }
__device__ __forceinline__ float grid_coord(int i, int S) { return (2.0f * (float)i + 1.0f) / (float)S - 1.0f; }  // affine_grid, align_corners=False
__device__ __forceinline__ float grid_dist(int a, int b, int S) {
    const float dx = grid_coord(a % S, S) - grid_coord(b % S, S), dy = grid_coord(a / S, S) - grid_coord(b / S, S);
    return sqrtf(dx * dx + dy * dy + 1e-12f);   // generic_torch.py:132-140
}

struct L1Params {
    const float* e; const float* r; const float* t;   // (H, N, d); t (amodal target) may be null
    const float* m_bg; const float* m_edit; const float* m_am; const float* w_am;   // (N)
    float c_sim, c_mov, c_amo, c_smh, c_smw;   // weight / denominator of each term (0 disables it)
    int H, S, d;
    float* grad;      // (H, N, d)  d(weighted loss)/d r
    float* partials;  // (gridDim.x, 5) unweighted sums: sim, movement, amodal, smooth_h, smooth_w
};

__global__ void __launch_bounds__(256) l1_losses_kernel(const L1Params p) {
    __shared__ float sh[32];
    const int N = p.S * p.S, d = p.d, S = p.S;
    const long total = (long)p.H * N * d;
    float s_sim = 0.f, s_mov = 0.f, s_amo = 0.f, s_h = 0.f, s_w = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int row = (int)((i / d) % N);
        const int y = row / S, x = row % S;
        const float r = p.r[i], e = p.e[i];
        const float ad = fabsf(e - r), sg = sgnf(e - r);
        const float mb = p.m_bg[row], me = p.m_edit ? p.m_edit[row] : 0.f;
        s_sim += ad * mb;
        s_mov += ad * me;
        float g = -sg * (p.c_sim * mb + p.c_mov * me);
        if (p.t) {
            const float wa = p.w_am[row] * p.m_am[row];
            const float dt = p.t[i] - r;
            s_amo += fabsf(dt) * wa;
            g -= sgnf(dt) * p.c_amo * wa;
        }
        const long sy = (long)S * d;
        if (y > 0) { const float dv = r - p.r[i - sy]; g += p.c_smh * sgnf(dv); }
        if (y < S - 1) { const float dv = p.r[i + sy] - r; s_h += fabsf(dv); g -= p.c_smh * sgnf(dv); }
        if (x > 0) { const float dv = r - p.r[i - d]; g += p.c_smw * sgnf(dv); }
        if (x < S - 1) { const float dv = p.r[i + d] - r; s_w += fabsf(dv); g -= p.c_smw * sgnf(dv); }
        p.grad[i] = g;
    }
    float v[5] = {s_sim, s_mov, s_amo, s_h, s_w};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float t = block_sum(v[k], sh);
        if (threadIdx.x == 0) p.partials[blockIdx.x * 5 + k] = t;
    }
}

// removal loss, stage 2: reduce the per-tile masked arg-max partials and emit, per (h, m):
//   term  = w * (-log(p_bg+1e-4) + log(p_in+1e-4))                (unnormalised loss contribution)
//   g     = {g_bg, g_in} = d(weighted loss)/d corr at the two arg-max positions, j = {j_bg, j_in}
//   dex   = sum_k A_e * dL/dA_e = g_bg * p_bg + g_in * p_in       (the softmax-backward row correction)
__global__ void removal_finalize_kernel(const float4* __restrict__ partial, int n_tiles, int H, int M, int S,
                                        const int* __restrict__ rows, const float* __restrict__ mask_in,
                                        const float* __restrict__ mask_bg, float coef, const float* __restrict__ w_dev,
                                        float* __restrict__ term, float2* __restrict__ g, int2* __restrict__ j, float* __restrict__ dex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * M) return;
    if (w_dev) coef *= *w_dev;      // removal weight kept in device memory (adaptive schedule, CUDA-graph replay)
    const int h = i / M, m = i % M;
    float bi = -1.f, bb = -1.f; int ii = 0, ib = 0;
    for (int t = 0; t < n_tiles; ++t) {
        const float4 q = partial[((long)h * n_tiles + t) * M + m];
        if (q.x > bi) { bi = q.x; ii = __float_as_int(q.y); }
        if (q.z > bb) { bb = q.z; ib = __float_as_int(q.w); }
    }
    const float w = expf(-grid_dist(rows[m], ib, S));
    term[i] = w * (-logf(bb + 1e-4f) + logf(bi + 1e-4f));
    const float gb = -coef * w / (bb + 1e-4f) * mask_bg[ib];
    const float gi = coef * w / (bi + 1e-4f) * mask_in[ii];
    g[i] = make_float2(gb, gi);
    j[i] = make_int2(ib, ii);
    // p = corr * mask at the arg-max, so corr there = p when the mask is 1 (else the gradient is 0)
    dex[i] = gb * bb + gi * bi;
}

// extra[h, m, k] = g_bg * A_b[h, j_bg, k] + g_in * A_b[h, j_in, k]      (dL/dA_e rows, fp32)
// (paired: the two rows of (h, m) are rows m and M + m of a_b -- gd_attn_probs_rows2's layout -- instead of rows j[hm])
__global__ void removal_extra_kernel(const __nv_bfloat16* __restrict__ a_b, long ab_hs, int ld, int H, int M, int Nk,
                                     const float2* __restrict__ g, const int2* __restrict__ j, int paired, float* __restrict__ extra) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * M * ld) return;
    const int k = (int)(i % ld);
    const long hm = i / ld;
    const int h = (int)(hm / M);
    float v = 0.f;
    if (k < Nk) {
        const float2 gg = g[hm]; const int2 jj = paired ? make_int2((int)(hm % M), M + (int)(hm % M)) : j[hm];
        const __nv_bfloat16* ab = a_b + (long)h * ab_hs;
        v = gg.x * __bfloat162float(ab[(long)jj.x * ld + k]) + gg.y * __bfloat162float(ab[(long)jj.y * ld + k]);
    }
    extra[i] = v;
}

// The same rows, key-major: extraT[h, k, m] = g_bg * P2[h, m, k] + g_in * P2[h, M + m, k]   (H, Nk, Mp) fp32, Mp = M rounded up to 4.
// The tcgen05 backward's threads own one query row each and walk the keys: with the inpaint rows contiguous along m, the 32 lanes of a warp
// (consecutive query rows -> consecutive slots m) read consecutive floats for a given key -- one or two 128-byte lines per load instead of 32.
// 32 x 32 tiles through shared memory so that both the P2 reads (along k) and the extraT writes (along m) are coalesced.
__global__ void __launch_bounds__(256) removal_extra_t_kernel(const __nv_bfloat16* __restrict__ p2, int ld, int H, int M, int Mp, int Nk,
                                                              const float2* __restrict__ g, float* __restrict__ extra_t) {
    __shared__ float tile[32][33];
    const int h = blockIdx.z, m0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const __nv_bfloat16* base = p2 + (long)h * 2 * M * ld;
    for (int r = ty; r < 32; r += 8) {
        const int m = m0 + r, k = k0 + tx;
        float v = 0.f;
        if (m < M && k < Nk) {
            const float2 gg = g[(long)h * M + m];
            v = gg.x * __bfloat162float(base[(long)m * ld + k]) + gg.y * __bfloat162float(base[(long)(M + m) * ld + k]);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, m = m0 + tx;
        if (k < Nk && m < Mp) extra_t[((long)h * Nk + k) * Mp + m] = tile[tx][r];
    }
}

// terms[0..5] = sim, movement, removal, smoothness, amodal, total(weighted); single block, fixed summation order.
// terms_accum (optional, 6 floats) += terms  -- the controller's running per-step log / loss.
struct LossReduceParams {
    const float* partials; int n_part;     // (n_part, 5)
    const float* rem_terms; int n_rem;     // (H*M)
    float inv_sim, inv_mov, inv_amo, inv_smh, inv_smw, inv_rem;   // 1 / denominators
    float w_sim, w_mov, w_amo, w_sm, w_rem;
    const float* w_rem_dev;                // if set, replaces w_rem
    float amodal_gate;                     // 0 when N <= 32^2 (attention_processors.py:596-597)
    float* terms; float* terms_accum;
};
__global__ void __launch_bounds__(256) loss_reduce_kernel(const LossReduceParams p) {
    __shared__ float sh[32];
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < p.n_part; i += blockDim.x)
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] += p.partials[i * 5 + k];
    for (int i = threadIdx.x; i < p.n_rem; i += blockDim.x) v[5] += p.rem_terms[i];
    float s[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) s[k] = block_sum(v[k], sh);
    if (threadIdx.x == 0) {
        const float sim = s[0] * p.inv_sim, mov = s[1] * p.inv_mov, amo = s[2] * p.inv_amo * p.amodal_gate;
        const float smo = s[3] * p.inv_smh + s[4] * p.inv_smw, rem = s[5] * p.inv_rem;
        const float w_rem = p.w_rem_dev ? *p.w_rem_dev : p.w_rem;
        const float tot = p.w_sim * sim + p.w_mov * mov + w_rem * rem + p.w_sm * smo + p.w_amo * amo;
        const float out[6] = {sim, mov, rem, smo, amo, tot};
        for (int k = 0; k < 6; ++k) { p.terms[k] = out[k]; if (p.terms_accum) p.terms_accum[k] += out[k]; }
    }
}

// ---- amodal interpolation (depends only on the mask and the grid: built once per resolution per edit) ----------
// For every pixel the 4 largest inverse grid distances to pixels of the foreground (mask > 0.5), attention_sharing.py:79-83;
// ties resolved by ascending pixel index.  w[p] = exp(-(1 / max inv) / 5)  (:103)
// One warp per pixel p (round 1: one thread per pixel scanning all N candidates, 1.08 ms at S = 64 on 32 SMs): lane l keeps the top 4 of the
// candidates q = l, l + 32, ... (ascending q, strict comparison: equal values keep the lower index first), then four rounds of a warp arg-max over
// the lanes' current heads under the same total order (value descending, index ascending) merge them -- the result is the serial scan's, bit for bit.
__global__ void __launch_bounds__(256) amodal_knn_kernel(const float* __restrict__ m_edit, int S, int* __restrict__ idx4, float* __restrict__ val4,
                                                         float* __restrict__ w) {
    const int N = S * S;
    const int p = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (p >= N) return;
    float bv[4] = {-1.f, -1.f, -1.f, -1.f}; int bi[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int q = lane; q < N; q += 32) {
        const float fg = (m_edit[q] > 0.5f) ? 1.f : 0.f;
        const float dist = grid_dist(p, q, S) * 512.f / 2.0f + 100000.f * (1.0f - fg);
        const float inv = 1.0f / (dist + 1e-4f);
        if (inv > bv[3]) {
            int pos = 3;
            while (pos > 0 && inv > bv[pos - 1]) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
            bv[pos] = inv; bi[pos] = q;
        }
    }
    float head_v = bv[0]; int head_i = bi[0];
    int taken = 0;
    float first = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = head_v; int i = head_i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (i == head_i && v == head_v) {            // this lane's head won (indices are unique across lanes): advance it
            ++taken;
            head_v = taken == 1 ? bv[1] : taken == 2 ? bv[2] : taken == 3 ? bv[3] : -2.f;
            head_i = taken == 1 ? bi[1] : taken == 2 ? bi[2] : taken == 3 ? bi[3] : 0x7fffffff;
        }
        if (k == 0) first = v;
        if (lane == 0) { idx4[p * 4 + k] = i; val4[p * 4 + k] = v; }
    }
    if (lane == 0) w[p] = expf(-(1.0f / first) / 5.0f);
}

// u[h,p,c] = fg[p] ? e[h,p,c] : sum_k val[p,k] e[h, idx[p,k], c] / (sum_k val[p,k] + 1e-12)
__global__ void amodal_interp_kernel(const float* __restrict__ e, const float* __restrict__ m_edit, const int* __restrict__ idx4,
                                     const float* __restrict__ val4, int H, int N, int d, float* __restrict__ u) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * N * d) return;
    const int c = (int)(i % d), p = (int)((i / d) % N);
    const long hb = (i / ((long)N * d)) * (long)N * d;
    if (m_edit[p] > 0.5f) { u[i] = e[i]; return; }
    float acc = 0.f, vs = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float v = val4[p * 4 + k]; acc += e[hb + (long)idx4[p * 4 + k] * d + c] * v; vs += v; }
    u[i] = acc / (vs + 1e-12f);
}

// 5x5 gaussian (generic_torch.py:13-85, sigma = 4/6), zero padding, over the S x S grid of every (h, c) plane
struct Gauss5 { float k[25]; };
__global__ void smooth5_kernel(const float* __restrict__ u, Gauss5 gk, int H, int S, int d, float* __restrict__ t) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int N = S * S;
    if (i >= (long)H * N * d) return;
    const int p = (int)((i / d) % N), y = p / S, x = p % S;
    float acc = 0.f;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) {
            const int yy = y + dy, xx = x + dx;
            if (yy >= 0 && yy < S && xx >= 0 && xx < S) acc += gk.k[(dy + 2) * 5 + dx + 2] * u[i + ((long)(dy * S + dx)) * d];
        }
    t[i] = acc;
}

// out[h,p,c] = a[h,p,c] * ma[p] + b[h,p,c] * mb[p]   (attention_processors.py:617-624, 922-925)
template <typename TOut>
__global__ void blend_rows_kernel(const float* __restrict__ a, const float* __restrict__ ma, const float* __restrict__ b,
                                  const float* __restrict__ mb, int N, int d, long total, TOut* __restrict__ out, long o_rs, long o_hs) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % d);
    const int p = (int)((i / d) % N);
    const long h = i / ((long)d * N);
    float v = b[i] * (mb ? mb[p] : 1.0f);
    if (a) v = a[i] * ma[p] + v;
    const long o = h * o_hs + (long)p * o_rs + c;
    if (sizeof(TOut) == 2) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(out)[o] = v;
}

}  // namespace gd

using namespace gd;

extern "C" {

int gd_attn_l1_losses(const float* e, const float* r, const float* t, const float* m_bg, const float* m_edit, const float* m_am,
                      const float* w_am, float c_sim, float c_mov, float c_amo, float c_smh, float c_smw, int H, int S, int d,
                      float* grad, float* partials, int n_partials, void* stream) {
    GD_CHECK_ARG(e && r && m_bg && grad && partials && H > 0 && S > 1 && d > 0 && n_partials > 0);
    GD_CHECK_ARG(t == nullptr || (m_am && w_am));
    L1Params p = {e, r, t, m_bg, m_edit, m_am, w_am, c_sim, c_mov, c_amo, c_smh, c_smw, H, S, d, grad, partials};
    l1_losses_kernel<<<n_partials, 256, 0, (cudaStream_t)stream>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_removal_finalize(const float* partial, int n_tiles, int H, int M, int S, const int* rows, const float* mask_in,
                        const float* mask_bg, float coef, const float* w_dev, const void* a_b, int Nb, int Nk, int ld, float* term, float* g2,
                        int* j2, float* delta_extra, float* extra, void* stream) {
    GD_CHECK_ARG(partial && rows && mask_in && mask_bg && term && g2 && j2 && delta_extra && H > 0 && M > 0 && n_tiles > 0);
    GD_CHECK_ARG(a_b == nullptr || extra != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    removal_finalize_kernel<<<ceil_div((long)H * M, 128), 128, 0, st>>>((const float4*)partial, n_tiles, H, M, S, rows, mask_in, mask_bg,
                                                                         coef, w_dev, term, (float2*)g2, (int2*)j2, delta_extra);
    GD_CHECK_LAUNCH();
    if (a_b) {      // base map materialised (cross layers / small levels): gather its two rows per (h, m); otherwise gd_removal_extra_rows follows
        removal_extra_kernel<<<ceil_div((long)H * M * ld, 256), 256, 0, st>>>((const __nv_bfloat16*)a_b, (long)Nb * ld, ld, H, M, Nk,
                                                                               (const float2*)g2, (const int2*)j2, 0, extra);
        GD_CHECK_LAUNCH();
    }
    return GD_OK;
}

// extra[h, m, :] = g_bg * P2[h, m, :] + g_in * P2[h, M + m, :] with P2 from gd_attn_probs_rows2 (the two base-map rows recomputed, not gathered)
int gd_removal_extra_rows(const void* p2, const float* g2, int H, int M, int Nk, int ld, float* extra, int key_major, void* stream) {
    GD_CHECK_ARG(p2 && g2 && extra && H > 0 && M > 0 && Nk > 0 && ld >= Nk);
    if (key_major) {        // extra (H, Nk, Mp), Mp = M rounded up to 4: the layout gd_attn_bwd_sm100 reads with extra_key_major = 1
        const int Mp = (M + 3) / 4 * 4;
        dim3 grid(ceil_div(Nk, 32), ceil_div(Mp, 32), H);
        removal_extra_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)p2, ld, H, M, Mp, Nk, (const float2*)g2, extra);
        GD_CHECK_LAUNCH();
        return GD_OK;
    }
    removal_extra_kernel<<<ceil_div((long)H * M * ld, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)p2, (long)2 * M * ld, ld, H, M, Nk,
                                                                                          (const float2*)g2, nullptr, 1, extra);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// inv[6] = 1/denominator of {sim, movement, amodal, smooth_h, smooth_w, removal}; w[5] = weights {sim, movement, amodal, smoothness, removal}
int gd_loss_reduce(const float* partials, int n_part, const float* rem_terms, int n_rem, const float* inv6_host, const float* w5_host,
                   const float* w_rem_dev, float amodal_gate, float* terms6, float* terms_accum6, void* stream) {
    GD_CHECK_ARG(inv6_host && w5_host && terms6 && (partials || n_part == 0) && (rem_terms || n_rem == 0));
    LossReduceParams p = {partials, n_part, rem_terms, n_rem, inv6_host[0], inv6_host[1], inv6_host[2], inv6_host[3], inv6_host[4],
                          inv6_host[5], w5_host[0], w5_host[1], w5_host[2], w5_host[3], w5_host[4], w_rem_dev, amodal_gate, terms6, terms_accum6};
    loss_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_amodal_knn(const float* m_edit, int S, int* idx4, float* val4, float* w, void* stream) {
    GD_CHECK_ARG(m_edit && idx4 && val4 && w && S > 1);
    amodal_knn_kernel<<<ceil_div((long)S * S * 32, 256), 256, 0, (cudaStream_t)stream>>>(m_edit, S, idx4, val4, w);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_amodal_target(const float* e, const float* m_edit, const int* idx4, const float* val4, const float* gauss25_host, int H, int S,
                     int d, float* scratch, float* target, void* stream) {
    GD_CHECK_ARG(e && m_edit && idx4 && val4 && gauss25_host && scratch && target && H > 0 && S > 1 && d > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const long total = (long)H * S * S * d;
    amodal_interp_kernel<<<ceil_div(total, 256), 256, 0, st>>>(e, m_edit, idx4, val4, H, S * S, d, scratch);
    GD_CHECK_LAUNCH();
    Gauss5 gk;
    for (int i = 0; i < 25; ++i) gk.k[i] = gauss25_host[i];
    smooth5_kernel<<<ceil_div(total, 256), 256, 0, st>>>(scratch, gk, H, S, d, target);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_blend_rows(const float* a, const float* ma, const float* b, const float* mb, int H, int N, int d, void* out, int out_is_bf16,
                  const long* out_strides, void* stream) {
    GD_CHECK_ARG(b && out && H > 0 && N > 0 && d > 0 && (a == nullptr || ma != nullptr));
    const long total = (long)H * N * d;
    const long o_rs = out_strides ? out_strides[0] : d, o_hs = out_strides ? out_strides[1] : (long)N * d;
    if (out_is_bf16) blend_rows_kernel<__nv_bfloat16><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(a, ma, b, mb, N, d, total, (__nv_bfloat16*)out, o_rs, o_hs);
    else blend_rows_kernel<float><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(a, ma, b, mb, N, d, total, (float*)out, o_rs, o_hs);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
