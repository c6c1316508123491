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

}  // extern "C"
