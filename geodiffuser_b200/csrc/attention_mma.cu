// geodiffuser_b200/csrc/attention_mma.cu
//
// Shared-attention forward / backward for the RAGGED and SMALL shapes of the path (cross-attention with
// 77 text keys, the 16^2 / 8^2 UNet levels, head_dim 160) plus the backward of every shape.  Flash-style:
// scores never touch HBM; softmax statistics are one (max, sum) pair per row kept in registers.
// The large self-attention forward (N >= 1024, head_dim 40 / 80) is served by attention_sm100.cu
// (tcgen05 + TMEM + TMA); this file uses warp-level mma.sync (HMMA) tiles, which is the right tool for
// 77-key / 64-row problems that cannot fill a 128xN UMMA tile.
//
// Reference behaviour replaced: attention_sharing.py:30-47 (`compute_attention`: baddbmm + softmax) and the
// torch.bmm(P, V) at its call sites attention_processors.py:427-433, 548-557, 643-647; the backward replaces
// what torch autograd derives for those ops (dQ for self layers -- K/V are detached, attention_sharing.py:242 --
// and dQ + dK for cross layers whose keys come from the edit sample, attention_processors.py:432).
#include "mma_util.cuh"

namespace gd {

constexpr int ATT_MAXG = 8;
constexpr int ATT_THREADS = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnFwdParams {
    const bf16* q[ATT_MAXG];
    const bf16* k[ATT_MAXG];
    const bf16* v[ATT_MAXG];
    float* o[ATT_MAXG];        // (H, N, d) fp32 contiguous, or NULL
    float* lse[ATT_MAXG];
    void* os[ATT_MAXG];        // strided output (row stride os_rs, head stride os_hs; bf16 or fp32), or NULL
    int G, H, N, Nk, d;
    float scale;
    long q_rs, q_hs, kv_rs, kv_hs, os_rs, os_hs;   // element strides of one (H, N, d) slab: token row, head
    int os_bf16;
};

// grid (ceil(N/64), H, G); 4 warps x 16 query rows; key tiles of 64, double-buffered with cp.async
template <int DPAD>
__global__ void __launch_bounds__(ATT_THREADS) flash_fwd_mma_kernel(const AttnFwdParams p) {
    constexpr int LD = DPAD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
    bf16* Ks = Qs + 64 * LD;      // 2 stages
    bf16* Vs = Ks + 2 * 64 * LD;  // 2 stages
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
    const int N = p.N, Nk = p.Nk, d = p.d;
    const bf16* Qg = p.q[g] + (long)h * p.q_hs + (long)q0 * p.q_rs;
    const bf16* Kg = p.k[g] + (long)h * p.kv_hs;
    const bf16* Vg = p.v[g] + (long)h * p.kv_hs;

    // the d..DPAD pad columns are never written by the tile loads but are read by the k-steps: zero them once (head_dim 40 only; the tile
    // loads touch other addresses, and the first __syncthreads of the key loop orders these stores before any fragment read)
    if (d < DPAD)
        for (int i = tid; i < 5 * 64 * ((DPAD - d) / 8); i += ATT_THREADS) {
            const int r = i / ((DPAD - d) / 8), c = d + (i % ((DPAD - d) / 8)) * 8;
            *reinterpret_cast<uint4*>(Qs + r * LD + c) = make_uint4(0, 0, 0, 0);
        }

    const int nT = (Nk + 63) / 64;
    load_tile_async<ATT_THREADS>(Qs, LD, Qg, p.q_rs, d, min(64, N - q0), tid);
    load_tile_async<ATT_THREADS>(Ks, LD, Kg, p.kv_rs, d, min(64, Nk), tid);
    load_tile_async<ATT_THREADS>(Vs, LD, Vg, p.kv_rs, d, min(64, Nk), tid);
    cp_async_commit();

    float o_acc[DPAD / 8][4];
#pragma unroll
    for (int i = 0; i < DPAD / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const float scale2 = p.scale * LOG2E;
    const int row0 = warp * 16;

    for (int jt = 0; jt < nT; ++jt) {
        const int st = jt & 1;
        if (jt + 1 < nT) {
            const int k1 = (jt + 1) * 64;
            load_tile_async<ATT_THREADS>(Ks + (st ^ 1) * 64 * LD, LD, Kg + (long)k1 * p.kv_rs, p.kv_rs, d, min(64, Nk - k1), tid);
            load_tile_async<ATT_THREADS>(Vs + (st ^ 1) * 64 * LD, LD, Vg + (long)k1 * p.kv_rs, p.kv_rs, d, min(64, Nk - k1), tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const bf16* Kt = Ks + st * 64 * LD;
        const bf16* Vt = Vs + st * 64 * LD;

        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < DPAD / 16; ++ks) {
            uint32_t a[4];
            load_a_frag(a, Qs, LD, row0, ks * 16, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                uint32_t b0, b1;
                load_b_frag_nt(b0, b1, Kt, LD, nb * 8, ks * 16, lane);
                mma_bf16_16816(s[nb], a, b0, b1);
            }
        }
        // scale into log2 domain, mask keys beyond Nk, running max
        const int kbase = jt * 64 + (lane & 3) * 2;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int key = kbase + nb * 8 + (e & 1);
                const float v = (key < Nk) ? s[nb][e] * scale2 : -INFINITY;
                s[nb][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            alpha[r] = exp2f(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pa[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = exp2f(s[nb][0] - m_run[0]), p1 = exp2f(s[nb][1] - m_run[0]);
            const float p2 = exp2f(s[nb][2] - m_run[1]), p3 = exp2f(s[nb][3] - m_run[1]);
            rs[0] += p0 + p1; rs[1] += p2 + p3;
            pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
            pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
        l_run[0] = l_run[0] * alpha[0] + rs[0];
        l_run[1] = l_run[1] * alpha[1] + rs[1];
#pragma unroll
        for (int i = 0; i < DPAD / 8; ++i) { o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0]; o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1]; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int n2 = 0; n2 < DPAD / 16; ++n2) {
                uint32_t b[4];
                load_b_frag_nn_x2(b, Vt, LD, kk * 16, n2 * 16, lane);
                mma_bf16_16816(o_acc[2 * n2], pa[kk], b[0], b[1]);
                mma_bf16_16816(o_acc[2 * n2 + 1], pa[kk], b[2], b[3]);
            }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    float* Og = p.o[g] ? p.o[g] + ((long)h * N + q0) * d : nullptr;
    float* Lg = p.lse[g] + (long)h * N + q0;
    unsigned char* Sg = p.os[g] ? reinterpret_cast<unsigned char*>(p.os[g]) + ((long)h * p.os_hs + (long)q0 * p.os_rs) * (p.os_bf16 ? 2 : 4) : nullptr;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = row0 + (lane >> 2) + r * 8;
        if (q0 + row >= N) continue;
        const float inv = 1.0f / l_run[r];
#pragma unroll
        for (int i = 0; i < DPAD / 8; ++i) {
            const int col = i * 8 + (lane & 3) * 2;
            if (col >= d) continue;
            const float a = o_acc[i][2 * r] * inv, b = o_acc[i][2 * r + 1] * inv;
            if (Og) *reinterpret_cast<float2*>(Og + (long)row * d + col) = make_float2(a, b);
            if (Sg) {
                if (p.os_bf16) *reinterpret_cast<uint32_t*>(Sg + ((long)row * p.os_rs + col) * 2) = pack_bf16(a, b);
                else *reinterpret_cast<float2*>(Sg + ((long)row * p.os_rs + col) * 4) = make_float2(a, b);
            }
        }
        if ((lane & 3) == 0) Lg[row] = (m_run[r] + log2f(l_run[r])) * LN2;
    }
}

// ------------------------------------------------------------------------------------------------
// backward.  One kernel, two roles:
//   MODE 0 (dQ): outer rows = queries (X1 = Q, X2 = dO), inner tiles = keys (Y1 = K, Y2 = V); stats indexed by outer row
//   MODE 1 (dK): outer rows = keys    (X1 = K, X2 = V),  inner tiles = queries (Y1 = Q, Y2 = dO); stats indexed by inner col
//   S' = X1 Y1^T ; P' = exp2(S' * scale2 - lse2) ; dP' = X2 Y2^T (+ extra) ; dS' = P' o (dP' - delta) ; OUT += dS' Y1 ; OUT *= scale
// `extra` (H, M, ex_ld) fp32 is dL/dP for the query rows listed in rowmap (query -> slot or -1): the removal-loss term that
// acts on the attention map itself (attention_processors.py:248-280).
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
    const bf16* x1; const bf16* x2; const bf16* y1; const bf16* y2;
    const float* lse; const float* delta;   // (H, Nq) natural-log lse, delta
    const float* extra; const float* extra_scale; const int* rowmap; int ex_ld; int M;
    void* out;                              // MODE 0 / unsplit MODE 1: strided (out_rs, out_hs), fp32 or bf16; split MODE 1: fp32 partials
    int H, n_outer, n_inner, d;
    float scale;
    long x1_rs, x1_hs, x2_rs, x2_hs, y1_rs, y1_hs, y2_rs, y2_hs, out_rs, out_hs;   // element strides (token row, head) per operand
    int out_bf16;
    int chunk;                              // MODE 1 only: inner (query) tiles per blockIdx.z; out is then (gridDim.z, H, n_outer, d) partial sums
};

template <int DPAD, int MODE>
__global__ void __launch_bounds__(ATT_THREADS) flash_bwd_mma_kernel(const AttnBwdParams p) {
    constexpr int LD = DPAD + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* X1 = reinterpret_cast<bf16*>(smem_raw);
    bf16* X2 = X1 + 64 * LD;
    bf16* Y1 = X2 + 64 * LD;      // 2 stages
    bf16* Y2 = Y1 + 2 * 64 * LD;  // 2 stages
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.y, o0 = blockIdx.x * 64;
    const int no = p.n_outer, ni = p.n_inner, d = p.d;
    const int Nq = (MODE == 0) ? no : ni;
    const bf16* x1g = p.x1 + (long)h * p.x1_hs + (long)o0 * p.x1_rs;
    const bf16* x2g = p.x2 + (long)h * p.x2_hs + (long)o0 * p.x2_rs;
    const bf16* y1g = p.y1 + (long)h * p.y1_hs;
    const bf16* y2g = p.y2 + (long)h * p.y2_hs;
    const float* lse = p.lse + (long)h * Nq;
    const float* delta = p.delta + (long)h * Nq;

    if (d < DPAD)       // pad columns d..DPAD (see the forward kernel)
        for (int i = tid; i < 6 * 64 * ((DPAD - d) / 8); i += ATT_THREADS) {
            const int r = i / ((DPAD - d) / 8), c = d + (i % ((DPAD - d) / 8)) * 8;
            *reinterpret_cast<uint4*>(X1 + r * LD + c) = make_uint4(0, 0, 0, 0);
        }
    // dK walks the queries: with Nk = 77 there are only 2 outer tiles per head, so the query range is split over blockIdx.z
    // (16 CTAs walking 4096 queries serially took 370 us per launch) and the partial sums are added in a fixed order afterwards
    const int nT_all = (ni + 63) / 64;
    const int jt0 = (MODE == 1 && p.chunk > 0) ? blockIdx.z * p.chunk : 0;
    const int nT = (MODE == 1 && p.chunk > 0) ? min(nT_all, jt0 + p.chunk) : nT_all;
    load_tile_async<ATT_THREADS>(X1, LD, x1g, p.x1_rs, d, min(64, no - o0), tid);
    load_tile_async<ATT_THREADS>(X2, LD, x2g, p.x2_rs, d, min(64, no - o0), tid);
    load_tile_async<ATT_THREADS>(Y1, LD, y1g + (long)jt0 * 64 * p.y1_rs, p.y1_rs, d, min(64, ni - jt0 * 64), tid);
    load_tile_async<ATT_THREADS>(Y2, LD, y2g + (long)jt0 * 64 * p.y2_rs, p.y2_rs, d, min(64, ni - jt0 * 64), tid);
    cp_async_commit();

    float acc[DPAD / 8][4];
#pragma unroll
    for (int i = 0; i < DPAD / 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const float scale2 = p.scale * LOG2E;
    const int row0 = warp * 16;
    const int orow[2] = {o0 + row0 + (lane >> 2), o0 + row0 + (lane >> 2) + 8};
    float lse_o[2] = {0.f, 0.f}, del_o[2] = {0.f, 0.f};
    int slot_o[2] = {-1, -1};
    const float ex_scale = (p.extra && p.extra_scale) ? *p.extra_scale : 1.0f;
    if (MODE == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (orow[r] < no) {
                lse_o[r] = lse[orow[r]] * LOG2E; del_o[r] = delta[orow[r]];
                if (p.rowmap) slot_o[r] = p.rowmap[orow[r]];
            }
    }

    for (int jt = jt0; jt < nT; ++jt) {
        const int st = (jt - jt0) & 1;
        if (jt + 1 < nT) {
            const int i1 = (jt + 1) * 64;
            load_tile_async<ATT_THREADS>(Y1 + (st ^ 1) * 64 * LD, LD, y1g + (long)i1 * p.y1_rs, p.y1_rs, d, min(64, ni - i1), tid);
            load_tile_async<ATT_THREADS>(Y2 + (st ^ 1) * 64 * LD, LD, y2g + (long)i1 * p.y2_rs, p.y2_rs, d, min(64, ni - i1), tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const bf16* Y1t = Y1 + st * 64 * LD;
        const bf16* Y2t = Y2 + st * 64 * LD;

        float s[8][4], dp[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < DPAD / 16; ++ks) {
            uint32_t a1[4], a2[4];
            load_a_frag(a1, X1, LD, row0, ks * 16, lane);
            load_a_frag(a2, X2, LD, row0, ks * 16, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                uint32_t b0, b1;
                load_b_frag_nt(b0, b1, Y1t, LD, nb * 8, ks * 16, lane);
                mma_bf16_16816(s[nb], a1, b0, b1);
                load_b_frag_nt(b0, b1, Y2t, LD, nb * 8, ks * 16, lane);
                mma_bf16_16816(dp[nb], a2, b0, b1);
            }
        }
        uint32_t da[4][4];
        const int ibase = jt * 64 + (lane & 3) * 2;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            float ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int icol = ibase + nb * 8 + (e & 1);
                const bool ok = (icol < ni) && (orow[r] < no);
                float l2, dl; int slot; int key;
                if (MODE == 0) { l2 = lse_o[r]; dl = del_o[r]; slot = slot_o[r]; key = icol; }
                else {
                    l2 = ok ? lse[icol] * LOG2E : 0.f; dl = ok ? delta[icol] : 0.f;
                    slot = (ok && p.rowmap) ? p.rowmap[icol] : -1; key = orow[r];
                }
                float pv = ok ? exp2f(s[nb][e] * scale2 - l2) : 0.f;
                float dpv = dp[nb][e];
                if (slot >= 0 && ok) dpv += ex_scale * p.extra[((long)h * p.M + slot) * p.ex_ld + key];
                ds[e] = pv * (dpv - dl);
            }
            da[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
            da[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int n2 = 0; n2 < DPAD / 16; ++n2) {
                uint32_t b[4];
                load_b_frag_nn_x2(b, Y1t, LD, kk * 16, n2 * 16, lane);
                mma_bf16_16816(acc[2 * n2], da[kk], b[0], b[1]);
                mma_bf16_16816(acc[2 * n2 + 1], da[kk], b[2], b[3]);
            }
        __syncthreads();
    }
    const bool split = (MODE == 1 && p.chunk > 0);     // partial sums: fp32 (Z, H, n_outer, d) contiguous
    const long ors = split ? d : p.out_rs;
    unsigned char* og = reinterpret_cast<unsigned char*>(p.out) +
                        (split ? ((long)blockIdx.z * p.H + h) * no * d * 4 : (long)h * p.out_hs * (p.out_bf16 ? 2 : 4));
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (orow[r] >= no) continue;
#pragma unroll
        for (int i = 0; i < DPAD / 8; ++i) {
            const int col = i * 8 + (lane & 3) * 2;
            if (col >= d) continue;
            const float a = acc[i][2 * r] * p.scale, b = acc[i][2 * r + 1] * p.scale;
            if (!split && p.out_bf16) *reinterpret_cast<uint32_t*>(og + ((long)orow[r] * ors + col) * 2) = pack_bf16(a, b);
            else *reinterpret_cast<float2*>(og + ((long)orow[r] * ors + col) * 4) = make_float2(a, b);
        }
    }
}

// dO (bf16) and delta for the backward:  dO = g_out * coef[row] + g_loss * loss_scale ;
//   delta[row] = sum_c dO*O  (+ delta_extra[slot(row)]).  One warp per row.
__global__ void attn_bwd_prep_kernel(const void* __restrict__ g_out, int g_out_bf16, long g_rs, long g_hs, const float* __restrict__ coef,
                                     const float* __restrict__ g_loss, const float* __restrict__ loss_scale,
                                     const float* __restrict__ o, const float* __restrict__ delta_extra,
                                     const int* __restrict__ rowmap, int M, int H, int N, int d, bf16* __restrict__ d_o,
                                     float* __restrict__ delta) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= H * N) return;
    const int row = w % N, h = w / N;
    const float cf = coef ? coef[row] : 1.0f;
    const float ls = (g_loss && loss_scale) ? *loss_scale : 1.0f;
    float part = 0.f;
    for (int c = lane; c < d; c += 32) {
        const long i = (long)w * d + c;
        float g = 0.f;
        if (g_out) {
            const long gi = (long)h * g_hs + (long)row * g_rs + c;
            g = (g_out_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(g_out)[gi]) : reinterpret_cast<const float*>(g_out)[gi]) * cf;
        }
        if (g_loss) g += g_loss[i] * ls;
        const bf16 gb = __float2bfloat16_rn(g);
        d_o[i] = gb;
        part += __bfloat162float(gb) * o[i];
    }
    part = warp_sum(part);
    if (lane == 0) {
        if (delta_extra && rowmap) { const int s = rowmap[row]; if (s >= 0) part += delta_extra[(long)h * M + s] * ls; }
        delta[w] = part;
    }
}

// fp32 -> bf16 with optional row gather:  dst[h, m, :] = src[h, rows[m], :]
__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// out[i] = sum_z part[z, i] in ascending z (fixed order: deterministic)
// (part is (Z, H, rows, d) contiguous; out is strided (row stride rs, head stride hs), fp32 or bf16; d % 4 == 0)
__global__ void sum_partials_kernel(const float* __restrict__ part, int Z, long n, int rows, int d, void* __restrict__ out, long rs, long hs,
                                    int out_bf16) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float4 a = *reinterpret_cast<const float4*>(part + i);
    for (int z = 1; z < Z; ++z) {
        const float4 b = *reinterpret_cast<const float4*>(part + (long)z * n + i);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const int c = (int)(i % d);
    const long r = (i / d) % rows, h = i / ((long)d * rows);
    const long o = h * hs + r * rs + c;
    if (out_bf16) {
        uint2 v; v.x = pack_bf16(a.x, a.y); v.y = pack_bf16(a.z, a.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(out) + o) = v;
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = a;
    }
}

template <int DPAD> static int launch_fwd(const AttnFwdParams& p, cudaStream_t st) {
    const size_t smem = (size_t)5 * 64 * (DPAD + 8) * sizeof(bf16);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(flash_fwd_mma_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(ceil_div(p.N, 64), p.H, p.G);
    flash_fwd_mma_kernel<DPAD><<<grid, ATT_THREADS, smem, st>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

template <int DPAD, int MODE> static int launch_bwd(const AttnBwdParams& p, cudaStream_t st, int Z = 1) {
    const size_t smem = (size_t)6 * 64 * (DPAD + 8) * sizeof(bf16);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(flash_bwd_mma_kernel<DPAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(ceil_div(p.n_outer, 64), p.H, Z);
    flash_bwd_mma_kernel<DPAD, MODE><<<grid, ATT_THREADS, smem, st>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int attn_fwd_mma(const AttnFwdParams& p, cudaStream_t st) {
    if (p.d <= 48) return launch_fwd<48>(p, st);
    if (p.d <= 80) return launch_fwd<80>(p, st);
    if (p.d <= 160) return launch_fwd<160>(p, st);
    return set_error(GD_ERR_UNSUPPORTED, "head_dim %d > 160", p.d);
}

template <int MODE> static int bwd_dispatch(const AttnBwdParams& p, int d, cudaStream_t st, int Z) {
    if (d <= 48) return launch_bwd<48, MODE>(p, st, Z);
    if (d <= 80) return launch_bwd<80, MODE>(p, st, Z);
    if (d <= 160) return launch_bwd<160, MODE>(p, st, Z);
    return set_error(GD_ERR_UNSUPPORTED, "head_dim %d > 160", d);
}

}  // namespace gd

using namespace gd;

extern "C" {

// Strides.  Every q / k / v / output operand below is one (H, N, d) "slab" addressed as base + h * head_stride + n * row_stride + c
// (element strides).  `strides` is a HOST array of 6 longs {q_row, q_head, kv_row, kv_head, os_row, os_head}; NULL means contiguous
// (H, N, d) slabs (row = d, head = N * d) -- the reference's head_to_batch_dim layout (attention_sharing.py:210-242).  The projection
// layout (B, N, H*d) that to_q / to_k / to_v produce is {H*d, d}: the kernels read it in place, no head permute copy.
static void fill_strides(const long* s, int N, int Nk, int d, long* q_rs, long* q_hs, long* kv_rs, long* kv_hs, long* o_rs, long* o_hs) {
    *q_rs = s ? s[0] : d; *q_hs = s ? s[1] : (long)N * d;
    *kv_rs = s ? s[2] : d; *kv_hs = s ? s[3] : (long)Nk * d;
    *o_rs = s ? s[4] : d; *o_hs = s ? s[5] : (long)N * d;
}
static bool strides_ok(const long* s) {
    if (!s) return true;
    for (int i = 0; i < 6; ++i) if (s[i] <= 0 || (s[i] % 8) != 0) return false;   // 16-byte rows for cp.async / TMA
    return true;
}

// Forward over G query streams, each an (H, N, d) bf16 slab against its own K/V (H, Nk, d) slabs.  Per stream: o[g] (H,N,d) fp32 contiguous
// and / or os[g] strided in the layout of q (bf16 if os_is_bf16 else fp32) -- at least one of the two; lse[g] (H,N) fp32.
// Pointer arrays are HOST arrays of device pointers; os_host may be NULL (no strided outputs).
int gd_attn_fwd_generic(const void* const* q, const void* const* k, const void* const* v, void* const* o, void* const* lse, void* const* os,
                        int G, int H, int N, int Nk, int d, float scale, const long* strides, int os_is_bf16, void* stream) {
    GD_CHECK_ARG(q && k && v && o && lse && G > 0 && G <= ATT_MAXG && H > 0 && N > 0 && Nk > 0 && d > 0 && (d % 8) == 0 && strides_ok(strides));
    AttnFwdParams p;
    for (int g = 0; g < G; ++g) {
        GD_CHECK_ARG(q[g] && k[g] && v[g] && lse[g] && (o[g] || (os && os[g])));
        p.q[g] = (const bf16*)q[g]; p.k[g] = (const bf16*)k[g]; p.v[g] = (const bf16*)v[g];
        p.o[g] = (float*)o[g]; p.lse[g] = (float*)lse[g]; p.os[g] = os ? os[g] : nullptr;
    }
    p.G = G; p.H = H; p.N = N; p.Nk = Nk; p.d = d; p.scale = scale; p.os_bf16 = os_is_bf16;
    fill_strides(strides, N, Nk, d, &p.q_rs, &p.q_hs, &p.kv_rs, &p.kv_hs, &p.os_rs, &p.os_hs);
    return attn_fwd_mma(p, (cudaStream_t)stream);
}

// dO / delta preparation (see attn_bwd_prep_kernel).  g_out is strided (g_strides = {row, head}, NULL = contiguous (H,N,d)).
int gd_attn_bwd_prep(const void* g_out, int g_out_is_bf16, const long* g_strides, const float* coef, const float* g_loss, const float* loss_scale,
                     const float* o, const float* delta_extra, const int* rowmap, int M, int H, int N, int d, void* d_o_bf16,
                     float* delta, void* stream) {
    GD_CHECK_ARG(o && d_o_bf16 && delta && (g_out || g_loss) && H > 0 && N > 0 && d > 0);
    const long warps = (long)H * N;
    const long g_rs = g_strides ? g_strides[0] : d, g_hs = g_strides ? g_strides[1] : (long)N * d;
    attn_bwd_prep_kernel<<<ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(g_out, g_out_is_bf16, g_rs, g_hs, coef, g_loss, loss_scale, o,
                                                                                     delta_extra, rowmap, M, H, N, d, (bf16*)d_o_bf16, delta);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

static int bwd_fill(AttnBwdParams& p, int mode, const void* q, const void* k, const void* v, const void* d_o, int N, int Nk, int d,
                    const long* strides) {
    long q_rs, q_hs, kv_rs, kv_hs, o_rs, o_hs;
    fill_strides(strides, N, Nk, d, &q_rs, &q_hs, &kv_rs, &kv_hs, &o_rs, &o_hs);
    if (!strides && mode == 1) o_hs = (long)Nk * d;
    p.out_rs = o_rs; p.out_hs = o_hs;
    if (mode == 0) {
        p.x1 = (const bf16*)q; p.x1_rs = q_rs; p.x1_hs = q_hs;
        p.x2 = (const bf16*)d_o; p.x2_rs = d; p.x2_hs = (long)N * d;
        p.y1 = (const bf16*)k; p.y1_rs = kv_rs; p.y1_hs = kv_hs;
        p.y2 = (const bf16*)v; p.y2_rs = kv_rs; p.y2_hs = kv_hs;
        p.n_outer = N; p.n_inner = Nk;
    } else {
        p.x1 = (const bf16*)k; p.x1_rs = kv_rs; p.x1_hs = kv_hs;
        p.x2 = (const bf16*)v; p.x2_rs = kv_rs; p.x2_hs = kv_hs;
        p.y1 = (const bf16*)q; p.y1_rs = q_rs; p.y1_hs = q_hs;
        p.y2 = (const bf16*)d_o; p.y2_rs = d; p.y2_hs = (long)N * d;
        p.n_outer = Nk; p.n_inner = N;
    }
    return GD_OK;
}

// mode 0: out = dQ, strided like q; mode 1: out = dK, strided like k.  q (H,N,d), k, v (H,Nk,d) bf16 slabs (strides as above; entries 4,5 =
// row / head stride of `out`); d_o (H,N,d) bf16 contiguous; lse, delta (H,N) fp32; out fp32 or bf16 (out_is_bf16).
int gd_attn_bwd(int mode, const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* out, int H, int N, int Nk,
                int d, float scale, const long* strides, int out_is_bf16, void* stream) {
    GD_CHECK_ARG(q && k && v && d_o && lse && delta && out && H > 0 && N > 0 && Nk > 0 && d > 0 && (d % 8) == 0 && strides_ok(strides));
    GD_CHECK_ARG(mode == 0 || mode == 1);
    GD_CHECK_ARG((extra == nullptr) == (rowmap == nullptr));
    AttnBwdParams p;
    p.lse = lse; p.delta = delta; p.extra = extra; p.extra_scale = extra_scale; p.rowmap = rowmap; p.ex_ld = ex_ld; p.M = M; p.out = out;
    p.H = H; p.d = d; p.scale = scale; p.chunk = 0; p.out_bf16 = out_is_bf16;
    bwd_fill(p, mode, q, k, v, d_o, N, Nk, d, strides);
    cudaStream_t st = (cudaStream_t)stream;
    return mode == 0 ? bwd_dispatch<0>(p, d, st, 1) : bwd_dispatch<1>(p, d, st, 1);
}

// dK as gd_attn_bwd mode 1, with the query range split over the grid: workspace (>= splits * H * Nk * d floats, d % 4 == 0) receives
// the per-split partial sums, which are then added in ascending order (deterministic).  splits <= 1 or no workspace: same as mode 1.
int gd_attn_bwd_dk_split(const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                         const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* dk, float* workspace,
                         int splits, int H, int N, int Nk, int d, float scale, const long* strides, int out_is_bf16, void* stream) {
    const int nT = (N + 63) / 64;
    if (splits > nT) splits = nT;
    if (!workspace || splits <= 1)
        return gd_attn_bwd(1, q, k, v, d_o, lse, delta, extra, extra_scale, rowmap, ex_ld, M, dk, H, N, Nk, d, scale, strides, out_is_bf16, stream);
    GD_CHECK_ARG(q && k && v && d_o && lse && delta && dk && H > 0 && N > 0 && Nk > 0 && d > 0 && (d % 8) == 0 && strides_ok(strides));
    GD_CHECK_ARG((extra == nullptr) == (rowmap == nullptr));
    AttnBwdParams p;
    p.lse = lse; p.delta = delta; p.extra = extra; p.extra_scale = extra_scale; p.rowmap = rowmap; p.ex_ld = ex_ld; p.M = M; p.out = workspace;
    p.H = H; p.d = d; p.scale = scale; p.out_bf16 = 0;
    p.chunk = (nT + splits - 1) / splits;
    const int Z = (nT + p.chunk - 1) / p.chunk;
    bwd_fill(p, 1, q, k, v, d_o, N, Nk, d, strides);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = bwd_dispatch<1>(p, d, st, Z);
    if (rc != GD_OK) return rc;
    const long n = (long)H * Nk * d;
    sum_partials_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(workspace, Z, n, Nk, d, dk, p.out_rs, p.out_hs, out_is_bf16);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

int gd_cast_f32_to_bf16(const float* src, void* dst, long n, void* stream) {
    GD_CHECK_ARG(src && dst && n > 0);
    cast_bf16_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // extern "C"
