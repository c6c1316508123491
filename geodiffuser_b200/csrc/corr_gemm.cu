// geodiffuser_b200/csrc/corr_gemm.cu
//
// The two dense contractions behind the removal loss (attention_processors.py:248-280), as one batched
// "NT" GEMM  C[h,m,n] = sum_k A[h,m,k] * B[h,n,k]  (bf16 in, fp32 accumulate) with fused epilogues so that
// neither the score matrix nor the (H, M, N) correlation ever reaches HBM in fp32:
//   EPI_PROBS   : A = Q rows (optionally gathered), B = K  ->  P = exp(scale*C - lse[row]) stored bf16
//                 (materialises the base attention map A_b and the inpaint rows of the edit map A_e)
//   EPI_CORRMAX : A = A_e rows, B = A_b               ->  per row masked max / arg-max over n for the
//                 inpaint-column mask and the background-column mask, one partial per 64-wide n tile
#include "mma_util.cuh"

namespace gd {

constexpr int GE_BM = 64, GE_BN = 64, GE_BK = 64, GE_LD = GE_BK + 8, GE_THREADS = 128;
constexpr float GE_LOG2E = 1.4426950408889634f;

struct GemmParams {
    const bf16* a; const bf16* b;
    long a_hs, b_hs;            // head strides (elements)
    int lda, ldb;               // row strides (elements, multiple of 8)
    const int* a_rows;          // optional gather of A rows (M entries)
    const int* a_rows2; int M2; // optional per-head gather: A row of output row m is a_rows2[(h * M2 + m % M2) * 2 + m / M2]  (M = 2 * M2)
    int H, M, N, K;
    // EPI_PROBS
    const float* lse; int lse_hs; float scale; bf16* p_out; long p_hs; int ldp;
    // EPI_CORRMAX
    const float* mask_in; const float* mask_bg; float4* partial;   // (H, n_tiles, M) {max_in, arg_in, max_bg, arg_bg}
};

template <int EPI>
__global__ void __launch_bounds__(GE_THREADS) gemm_nt_kernel(const GemmParams p) {
    __shared__ __align__(16) bf16 As[2][GE_BM * GE_LD];
    __shared__ __align__(16) bf16 Bs[2][GE_BN * GE_LD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.z, m0 = blockIdx.y * GE_BM, n0 = blockIdx.x * GE_BN;
    const bf16* Ag = p.a + (long)h * p.a_hs;
    const bf16* Bg = p.b + (long)h * p.b_hs;
    const int nK = (p.K + GE_BK - 1) / GE_BK;

    auto load_stage = [&](int st, int k0) {
        for (int c = tid; c < GE_BM * (GE_BK / 8); c += GE_THREADS) {
            const int r = c / (GE_BK / 8), kk = (c % (GE_BK / 8)) * 8;
            const int m = m0 + r;
            const bool ok = (m < p.M) && (k0 + kk < p.K);
            const long row = ok ? (p.a_rows2 ? p.a_rows2[((long)h * p.M2 + (m % p.M2)) * 2 + (m / p.M2)] : p.a_rows ? p.a_rows[m] : m) : 0;
            cp_async16(&As[st][r * GE_LD + kk], Ag + row * p.lda + (ok ? k0 + kk : 0), ok);
        }
        for (int c = tid; c < GE_BN * (GE_BK / 8); c += GE_THREADS) {
            const int r = c / (GE_BK / 8), kk = (c % (GE_BK / 8)) * 8;
            const int n = n0 + r;
            const bool ok = (n < p.N) && (k0 + kk < p.K);
            cp_async16(&Bs[st][r * GE_LD + kk], Bg + (long)(ok ? n : 0) * p.ldb + (ok ? k0 + kk : 0), ok);
        }
        cp_async_commit();
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    load_stage(0, 0);
    for (int kt = 0; kt < nK; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < nK) { load_stage(st ^ 1, (kt + 1) * GE_BK); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GE_BK / 16; ++ks) {
            uint32_t a[4];
            load_a_frag(a, As[st], GE_LD, warp * 16, ks * 16, lane);
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                uint32_t b0, b1;
                load_b_frag_nt(b0, b1, Bs[st], GE_LD, nb * 8, ks * 16, lane);
                mma_bf16_16816(acc[nb], a, b0, b1);
            }
        }
        __syncthreads();
    }

    const int rloc = warp * 16 + (lane >> 2);
    if (EPI == 0) {
        const float scale2 = p.scale * GE_LOG2E;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int m = m0 + rloc + r * 8;
            if (m >= p.M) continue;
            const int qrow = p.a_rows2 ? p.a_rows2[((long)h * p.M2 + (m % p.M2)) * 2 + (m / p.M2)] : p.a_rows ? p.a_rows[m] : m;
            const float l2 = p.lse[(long)h * p.lse_hs + qrow] * GE_LOG2E;
            bf16* out = p.p_out + (long)h * p.p_hs + (long)m * p.ldp;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int n = n0 + nb * 8 + (lane & 3) * 2;
                if (n >= p.ldp) continue;
                const float v0 = (n < p.N) ? exp2f(acc[nb][2 * r] * scale2 - l2) : 0.f;
                const float v1 = (n + 1 < p.N) ? exp2f(acc[nb][2 * r + 1] * scale2 - l2) : 0.f;
                *reinterpret_cast<uint32_t*>(out + n) = pack_bf16(v0, v1);
            }
        }
    } else {
        float bi[2] = {-1.f, -1.f}, bb[2] = {-1.f, -1.f};
        int ii[2] = {-1, -1}, ib[2] = {-1, -1};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1, n = n0 + nb * 8 + (lane & 3) * 2 + (e & 1);
                if (n >= p.N) continue;
                const float vi = acc[nb][e] * p.mask_in[n], vb = acc[nb][e] * p.mask_bg[n];
                if (vi > bi[r]) { bi[r] = vi; ii[r] = n; }
                if (vb > bb[r]) { bb[r] = vb; ib[r] = n; }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float vi = __shfl_xor_sync(0xffffffffu, bi[r], o); const int xi = __shfl_xor_sync(0xffffffffu, ii[r], o);
                if (vi > bi[r] || (vi == bi[r] && xi >= 0 && (ii[r] < 0 || xi < ii[r]))) { bi[r] = vi; ii[r] = xi; }
                const float vb = __shfl_xor_sync(0xffffffffu, bb[r], o); const int xb = __shfl_xor_sync(0xffffffffu, ib[r], o);
                if (vb > bb[r] || (vb == bb[r] && xb >= 0 && (ib[r] < 0 || xb < ib[r]))) { bb[r] = vb; ib[r] = xb; }
            }
            const int m = m0 + rloc + r * 8;
            if (m < p.M && (lane & 3) == 0)
                p.partial[((long)h * gridDim.x + blockIdx.x) * p.M + m] = make_float4(bi[r], __int_as_float(ii[r]), bb[r], __int_as_float(ib[r]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Removal-loss gradient of the edit queries, as its own small contraction (round 2b).
//
// The loss reaches dS through dL/dA_e[h, rows[m], k] = extra[h, m, k] = g_bg A_b[j_bg, k] + g_in A_b[j_in, k] (two base-map rows per inpaint row):
//   dS = P o (dP - delta)  +  P o extra          ->   dQ = scale dS K = (flash-backward term) + scale (A_e[rows] o extra) K.
// The second term touches only the M inpaint rows (M ~ 2-10 % of N) and A_e[rows] is already materialised for the correlation, so it is the
// GEMM  W (M x Nk) @ K (Nk x d)  with W = A_e[rows] o extra: 2 H M Nk d FLOP (0.2 GF at M = 76) instead of 32 extra loads per (row, step) inside the
// tcgen05 backward, where the few CTAs that own inpaint rows set the makespan of a one-wave grid (96.8 / 84.5 us with M = 410 / 76 rows
// against 59.7 us without).  delta keeps its removal part (delta_extra, gd_attn_bwd_prep), so the two terms stay independent.
__global__ void __launch_bounds__(256) removal_weighted_rows_kernel(const bf16* __restrict__ a_e, const bf16* __restrict__ p2, const float2* __restrict__ g,
                                                                     int H, int M, int Nk, int ld, bf16* __restrict__ w) {
    const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 2;      // two keys per thread
    if (i >= (long)H * M * ld) return;
    const int k = (int)(i % ld);
    const long hm = i / ld;
    const int h = (int)(hm / M), m = (int)(hm % M);
    float v0 = 0.f, v1 = 0.f;
    if (k < Nk) {
        const float2 gg = g[hm];
        const bf16* pa = p2 + ((long)h * 2 * M + m) * ld + k;
        const bf16* pb = pa + (long)M * ld;
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a_e + i));
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pa));
        const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pb));
        v0 = a.x * (gg.x * x.x + gg.y * y.x);
        v1 = (k + 1 < Nk) ? a.y * (gg.x * x.y + gg.y * y.y) : 0.f;
    }
    *reinterpret_cast<__nv_bfloat162*>(w + i) = __floats2bfloat162_rn(v0, v1);
}

// dq[h, rows[m], :] += (*gscale) * scale * sum_k W[h, m, k] K[h, k, :].  One CTA per (16 inpaint rows, head); its four warps split the keys and
// meet in shared memory; mma.sync m16n8k16 (A = W chunk via ldmatrix, B = K chunk via ldmatrix.trans).  The arithmetic is negligible (12 MMAs per
// 32 keys): the kernel is a latency problem, so every warp keeps RQ_STAGES - 1 chunks of 64 keys (K rows + W rows, cp.async) in flight.
constexpr int RQ_STAGES = 4, RQ_CK = 64, RQ_LDW = RQ_CK + 8;
template <int D> struct RqCfg {
    static constexpr int DP = (D + 15) / 16 * 16;          // 48 / 80 accumulator columns
    static constexpr int LDK = DP + 8;                     // shared row stride of a K chunk (elements): 16-byte aligned rows, conflict-free ldmatrix
    static constexpr int STAGE = (RQ_CK * LDK + 16 * RQ_LDW) * 2;      // bytes: K chunk + W chunk
    static constexpr int SMEM = 4 * RQ_STAGES * STAGE;
};
template <int D>
__global__ void __launch_bounds__(128) removal_dq_rows_kernel(const bf16* __restrict__ w, int ld, const bf16* __restrict__ kmat, long k_rs, long k_hs,
                                                               const int* __restrict__ rows, const float* __restrict__ gscale, float scale, void* dq,
                                                               long dq_rs, long dq_hs, int dq_bf16, int M, int Nk) {
    typedef RqCfg<D> C;
    constexpr int DP = C::DP, LDK = C::LDK, NT = DP / 8, CK = RQ_CK, NST = RQ_STAGES;
    extern __shared__ __align__(16) unsigned char rq_smem[];
    float (*red)[16][DP + 1] = reinterpret_cast<float (*)[16][DP + 1]>(rq_smem);       // reuses the rings once every warp is through its keys
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = blockIdx.y, m0 = blockIdx.x * 16;
    const int gq = lane >> 2, tq = lane & 3;
    const bf16* Kg = kmat + (long)h * k_hs;
    const bf16* Wg = w + (long)h * M * ld;
    unsigned char* ring = rq_smem + warp * NST * C::STAGE;
    auto Kst = [&](int s) { return reinterpret_cast<bf16*>(ring + s * C::STAGE); };
    auto Wst = [&](int s) { return reinterpret_cast<bf16*>(ring + s * C::STAGE) + CK * LDK; };
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
    // columns D .. LDK-1 of the K chunks feed accumulators that are never stored: zero them once so that they stay finite
    for (int c = lane; c < NST * CK * (LDK - D); c += 32) {
        const int s = c / (CK * (LDK - D)), r = (c / (LDK - D)) % CK;
        Kst(s)[r * LDK + D + c % (LDK - D)] = __float2bfloat16(0.f);
    }
    const int per = Nk / 4, kbeg = warp * per, nch = per / CK;       // Nk % 256 == 0
    constexpr int CPR = D / 8;                      // 16-byte pieces per K row
    auto issue = [&](int c) {
        if (c < nch) {
            const int s = c % NST, k0 = kbeg + c * CK;
            bf16* kd = Kst(s);
            bf16* wd = Wst(s);
            for (int i = lane; i < CK * CPR; i += 32) {
                const int r = i / CPR, cc = (i % CPR) * 8;
                cp_async16(kd + r * LDK + cc, Kg + (long)(k0 + r) * k_rs + cc, true);
            }
            for (int i = lane; i < 16 * (CK / 8); i += 32) {
                const int r = i / (CK / 8), cc = (i % (CK / 8)) * 8;
                const bool ok = m0 + r < M;
                cp_async16(wd + r * RQ_LDW + cc, Wg + (long)(ok ? m0 + r : 0) * ld + k0 + cc, ok);
            }
        }
        cp_async_commit();                          // (an empty group past the end keeps the wait count uniform)
    };
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) issue(c);
    for (int c = 0; c < nch; ++c) {
        issue(c + NST - 1);
        cp_async_wait<NST - 1>();
        __syncwarp();
        const bf16* kc = Kst(c % NST);
        const bf16* wc = Wst(c % NST);
#pragma unroll
        for (int ks = 0; ks < CK; ks += 16) {
            uint32_t a[4];
            load_a_frag(a, wc, RQ_LDW, 0, ks, lane);
#pragma unroll
            for (int n = 0; n < NT; n += 2) {
                uint32_t b[4];
                load_b_frag_nn_x2(b, kc, LDK, ks, n * 8, lane);
                mma_bf16_16816(acc[n], a, b[0], b[1]);
                mma_bf16_16816(acc[n + 1], a, b[2], b[3]);
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        red[warp][gq][n * 8 + tq * 2] = acc[n][0];
        red[warp][gq][n * 8 + tq * 2 + 1] = acc[n][1];
        red[warp][gq + 8][n * 8 + tq * 2] = acc[n][2];
        red[warp][gq + 8][n * 8 + tq * 2 + 1] = acc[n][3];
    }
    __syncthreads();
    const float f = scale * (gscale ? *gscale : 1.0f);
    for (int e = tid; e < 16 * D; e += 128) {
        const int r = e / D, c = e % D, m = m0 + r;
        if (m >= M) continue;
        const float v = (red[0][r][c] + red[1][r][c] + red[2][r][c] + red[3][r][c]) * f;     // fixed order: deterministic
        const long off = (long)h * dq_hs + (long)rows[m] * dq_rs + c;
        if (dq_bf16) {
            bf16* o = reinterpret_cast<bf16*>(dq) + off;
            *o = __float2bfloat16(__bfloat162float(*o) + v);
        } else {
            reinterpret_cast<float*>(dq)[off] += v;
        }
    }
}

template <int D> static int launch_removal_dq_rows(dim3 grid, cudaStream_t st, const bf16* w, int ld, const bf16* k, long k_rs, long k_hs, const int* rows,
                                                   const float* gscale, float scale, void* dq, long dq_rs, long dq_hs, int dq_bf16, int M, int Nk) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(removal_dq_rows_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, RqCfg<D>::SMEM);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    removal_dq_rows_kernel<D><<<grid, 128, RqCfg<D>::SMEM, st>>>(w, ld, k, k_rs, k_hs, rows, gscale, scale, dq, dq_rs, dq_hs, dq_bf16, M, Nk);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" {

// P[h, m, :] = softmax row of q[h, rows[m] (or m), :] against k[h] given its natural-log lse; bf16 out, row stride ldp
// (multiple of 8, >= Nk; columns Nk..ldp zero-filled).  q (H,N,d), k (H,Nk,d) bf16 slabs; qk_strides = {q_row, q_head, k_row, k_head}
// element strides (HOST array; NULL = contiguous); lse (H,N).
int gd_attn_probs(const void* q, const void* k, const float* lse, const int* rows, int M, int H, int N, int Nk, int d, float scale,
                  void* p_out, int ldp, const long* qk_strides, void* stream) {
    GD_CHECK_ARG(q && k && lse && p_out && H > 0 && N > 0 && Nk > 0 && M > 0 && d > 0 && (d % 8) == 0 && (ldp % 8) == 0 && ldp >= Nk);
    GemmParams p = {};
    p.a = (const bf16*)q; p.b = (const bf16*)k; p.a_rows = rows;
    p.lda = qk_strides ? (int)qk_strides[0] : d; p.a_hs = qk_strides ? qk_strides[1] : (long)N * d;
    p.ldb = qk_strides ? (int)qk_strides[2] : d; p.b_hs = qk_strides ? qk_strides[3] : (long)Nk * d;
    GD_CHECK_ARG((p.lda % 8) == 0 && (p.ldb % 8) == 0 && (p.a_hs % 8) == 0 && (p.b_hs % 8) == 0);
    p.H = H; p.M = M; p.N = Nk; p.K = d; p.lse = lse; p.lse_hs = N; p.scale = scale; p.p_out = (bf16*)p_out; p.p_hs = (long)M * ldp; p.ldp = ldp;
    dim3 grid(ceil_div(ldp, GE_BN), ceil_div(M, GE_BM), H);
    gemm_nt_kernel<0><<<grid, GE_THREADS, 0, (cudaStream_t)stream>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// The two base-map rows per (head, inpaint row) through which the removal loss back-propagates (attention_processors.py:256-266: the arg-max
// positions j_bg, j_in of the masked correlation): P2[h, t * M + m, :] = softmax row of q[h, j2[(h * M + m) * 2 + t], :], bf16, row stride ldp.
// Replaces the gather from a materialised base map.  j2: (H * M, 2) int32 as written by gd_removal_finalize.
int gd_attn_probs_rows2(const void* q, const void* k, const float* lse, const int* j2, int M, int H, int N, int Nk, int d, float scale,
                        void* p_out, int ldp, const long* qk_strides, void* stream) {
    GD_CHECK_ARG(q && k && lse && j2 && p_out && H > 0 && N > 0 && Nk > 0 && M > 0 && d > 0 && (d % 8) == 0 && (ldp % 8) == 0 && ldp >= Nk);
    GemmParams p = {};
    p.a = (const bf16*)q; p.b = (const bf16*)k; p.a_rows2 = j2; p.M2 = M;
    p.lda = qk_strides ? (int)qk_strides[0] : d; p.a_hs = qk_strides ? qk_strides[1] : (long)N * d;
    p.ldb = qk_strides ? (int)qk_strides[2] : d; p.b_hs = qk_strides ? qk_strides[3] : (long)Nk * d;
    GD_CHECK_ARG((p.lda % 8) == 0 && (p.ldb % 8) == 0 && (p.a_hs % 8) == 0 && (p.b_hs % 8) == 0);
    p.H = H; p.M = 2 * M; p.N = Nk; p.K = d; p.lse = lse; p.lse_hs = N; p.scale = scale; p.p_out = (bf16*)p_out; p.p_hs = (long)2 * M * ldp; p.ldp = ldp;
    dim3 grid(ceil_div(ldp, GE_BN), ceil_div(2 * M, GE_BM), H);
    gemm_nt_kernel<0><<<grid, GE_THREADS, 0, (cudaStream_t)stream>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// partial[h, t, m] = masked (max, argmax) over the t-th 64-wide tile of n of  corr[h,m,n] = sum_k a_e[h,m,k] * a_b[h,n,k].
// a_e (H, M, ld) bf16, a_b (H, Nb, ld) bf16 with the same K = Nk (ld multiple of 8, pad columns zero).
int gd_corr_max_partial(const void* a_e, const void* a_b, int H, int M, int Nb, int Nk, int ld, const float* mask_in, const float* mask_bg,
                        float* partial, void* stream) {
    GD_CHECK_ARG(a_e && a_b && mask_in && mask_bg && partial && H > 0 && M > 0 && Nb > 0 && Nk > 0 && (ld % 8) == 0 && ld >= Nk);
    GemmParams p = {};
    p.a = (const bf16*)a_e; p.b = (const bf16*)a_b; p.a_hs = (long)M * ld; p.b_hs = (long)Nb * ld; p.lda = ld; p.ldb = ld;
    p.H = H; p.M = M; p.N = Nb; p.K = Nk; p.mask_in = mask_in; p.mask_bg = mask_bg; p.partial = (float4*)partial;
    dim3 grid(ceil_div(Nb, GE_BN), ceil_div(M, GE_BM), H);
    gemm_nt_kernel<1><<<grid, GE_THREADS, 0, (cudaStream_t)stream>>>(p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// W[h, m, k] = A_e[h, m, k] * (g2[h,m].x * P2[h, m, k] + g2[h,m].y * P2[h, M + m, k])  (bf16, (H, M, ld), zero past Nk): the removal term's dS rows
int gd_removal_weighted_rows(const void* a_e, const void* p2, const float* g2, int H, int M, int Nk, int ld, void* w_bf16, void* stream) {
    GD_CHECK_ARG(a_e && p2 && g2 && w_bf16 && H > 0 && M > 0 && Nk > 0 && ld >= Nk && ld % 2 == 0);
    const long n = ((long)H * M * ld + 1) / 2;
    removal_weighted_rows_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)a_e, (const bf16*)p2, (const float2*)g2, H, M, Nk, ld, (bf16*)w_bf16);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

// dq[h, rows[m], :] += (*gscale) * scale * sum_k W[h, m, k] K[h, k, :]   (the removal-loss part of dQ; see removal_dq_rows_kernel).
// k: slab with strides_host = {kv_row, kv_head, dq_row, dq_head} (NULL: contiguous (H, Nk, d) / (H, N, d)); dq fp32 or bf16, accumulated in place.
int gd_removal_dq_rows(const void* w_bf16, const void* k, const int* rows, const float* gscale, void* dq, int H, int M, int N, int Nk, int d, float scale,
                       int ld, const long* strides, int dq_is_bf16, void* stream) {
    GD_CHECK_ARG(w_bf16 && k && rows && dq && H > 0 && M > 0 && N > 0 && ld >= Nk && ld % 8 == 0);
    if (!((d == 40 || d == 80) && Nk % 256 == 0))
        return set_error(GD_ERR_UNSUPPORTED, "gd_removal_dq_rows serves d in {40, 80}, Nk %% 256 == 0; got d=%d Nk=%d", d, Nk);
    const long k_rs = strides ? strides[0] : d, k_hs = strides ? strides[1] : (long)Nk * d;
    const long dq_rs = strides ? strides[2] : d, dq_hs = strides ? strides[3] : (long)N * d;
    if ((k_rs % 8) != 0 || (k_hs % 8) != 0 || (reinterpret_cast<uintptr_t>(k) & 15) != 0)
        return set_error(GD_ERR_INVALID, "gd_removal_dq_rows: k base and strides must be 16-byte aligned");
    dim3 grid(ceil_div(M, 16), H);
    cudaStream_t st = (cudaStream_t)stream;
    return d == 40 ? launch_removal_dq_rows<40>(grid, st, (const bf16*)w_bf16, ld, (const bf16*)k, k_rs, k_hs, rows, gscale, scale, dq, dq_rs, dq_hs, dq_is_bf16, M, Nk)
                   : launch_removal_dq_rows<80>(grid, st, (const bf16*)w_bf16, ld, (const bf16*)k, k_rs, k_hs, rows, gscale, scale, dq, dq_rs, dq_hs, dq_is_bf16, M, Nk);
}

}  // extern "C"
