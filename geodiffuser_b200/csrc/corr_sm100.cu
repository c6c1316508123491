// geodiffuser_b200/csrc/corr_sm100.cu
//
// Row A7, the removal loss's correlation of attention maps (attention_processors.py:248-280)
//
//     corr[h, m, n] = sum_k A_e[h, rows[m], k] * A_b[h, n, k]          (M inpaint rows x N base rows x N keys: 2*H*M*N^2 FLOP)
//     p_in[h, m] = max_n corr * M_inpaint[n],   p_bg[h, m], j_bg[h, m] = max / argmax_n corr * M_bg[n]
//
// as ONE tcgen05 kernel that never materialises the base map A_b (round 1 wrote it to HBM in bf16 -- 268 MB per 64^2 layer and pass -- and
// read it back through an mma.sync GEMM).  A CTA owns (head, 128 base rows n, a chunk of <= 256 inpaint rows m) and walks the keys in
// 64-wide steps:
//
//     S    = Q_b[n tile] K_b[k tile]^T                      tcgen05.mma SS (Q, K from shared memory via TMA), fp32 in TMEM
//     P_b  = exp2(S * scale * log2e - lse_b[n] * log2e)     eight softmax warps, two threads per base row n = TMEM lane (32 keys each); bf16 back into TMEM
//     D   += P_b A_e[m chunk, k tile]^T                     tcgen05.mma TS: A = P_b from TMEM, B = the A_e rows (K-major, via TMA)
//
// i.e. the base-map tile is recomputed on the fly from q_b, k_b and the stored log-sum-exp, lives only in TMEM, and the accumulator
// D (128 n x MC m, fp32, <= 256 TMEM columns) holds corr^T for the whole key range.  Epilogue: D is read back 32 columns at a time,
// transposed through a warp-private shared-memory tile, and every thread scans one column m over the warp's 32 base rows for the two
// masked (max, first argmax) pairs: partial[h, 4 * n_tile + lane quarter, m] -- the layout gd_removal_finalize already reduces.
// A_e[rows] itself (H x M x N bf16, ~27 MB at the 64^2 level) is produced by gd_attn_probs as before.
//
// TMEM: S [0,64) | P double buffer [64,96) [96,128) | D [128, 128 + MC).  Roofline: tensor pipe (dense BF16); per key step the
// tensor work is 96 + 4 * MC/2 clk against 512 clk of MUFU for the 8192 exponentials, so it is MUFU-bound below MC ~ 200.
#include "sm100_util.cuh"

namespace gd {

constexpr int CORR_THREADS = 320;     // warps 0-7 softmax (lane quarter x key half), 8 = TMA, 9 = MMA issuer
constexpr int CORR_BM = 128;      // base rows per CTA
constexpr int CORR_BK = 64;       // keys per step

struct CorrMaps { CUtensorMap q, k, a; };
struct CorrParams {
    const float* lse;             // (H, N) natural log
    const float* mask_in; const float* mask_bg;   // (N)
    float4* partial;              // (H, 4 * N/128, M)
    int H, N, M, MC, n_stage, tmem_cols;
    float scale2;
};

template <int D>
__global__ void __launch_bounds__(CORR_THREADS, 2)
removal_corr_sm100_kernel(const __grid_constant__ CorrMaps maps, const CorrParams p) {
    constexpr int KB = (D + 63) / 64;
    constexpr int KSTEPS = (D + 15) / 16;
    constexpr int QTILE_BYTES = 128 * 128;                 // [128 rows][64 bf16] swizzled block of Q_b
    constexpr int KTILE_BYTES = CORR_BK * 128;             // [64 keys][64 bf16] block of K_b
    constexpr int Q_BYTES = KB * QTILE_BYTES, K_BYTES = KB * KTILE_BYTES;
    constexpr uint32_t COL_S = 0, COL_P = 64, COL_D = 128;
    constexpr int MAX_STAGE = 4;

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    const int MC = p.MC;
    const int A_BYTES = MC * 128;                          // [MC rows][64 bf16] block of A_e
    const int STAGE_BYTES = (K_BYTES + A_BYTES + 1023) & ~1023;
    unsigned char* sQ = smem;
    unsigned char* sStage = sQ + Q_BYTES;                  // n_stage x { K tile, A_e tile }
    unsigned char* sT = sStage + p.n_stage * STAGE_BYTES;  // epilogue transpose tiles: 8 warps x [32][33] floats
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT + 8 * 32 * 33 * 4);
    uint64_t* q_full = bars + 0;
    uint64_t* s_full = bars + 1;
    uint64_t* s_free = bars + 2;
    uint64_t* p_full = bars + 3;      // [2]
    uint64_t* p_free = bars + 5;      // [2]  D(j) has consumed P buffer j & 1
    uint64_t* st_full = bars + 7;     // [MAX_STAGE]
    uint64_t* st_empty = bars + 7 + MAX_STAGE;   // [MAX_STAGE]
    uint64_t* d_done = bars + 7 + 2 * MAX_STAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * MAX_STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.z, n0 = blockIdx.x * CORR_BM, m0 = blockIdx.y * MC;
    const int N = p.N;
    const int nT = N / CORR_BK;
    const int NS = p.n_stage;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(d_done, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(p_full + s, 8); mbar_init(p_free + s, 1); }
        for (int s = 0; s < MAX_STAGE; ++s) { mbar_init(st_full + s, 1); mbar_init(st_empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 8 && lane == 0) { tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.a); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 8) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int b = 0; b < KB; ++b) tma_load_3d(sQ + b * QTILE_BYTES, &maps.q, q_full, b * 64, n0, h);
            for (int j = 0; j < nT; ++j) {
                const int s = j % NS;
                const uint32_t ph = (j / NS) & 1;
                mbar_wait_relaxed(st_empty + s, ph ^ 1);
                mbar_expect_tx(st_full + s, K_BYTES + A_BYTES);
                unsigned char* dst = sStage + s * STAGE_BYTES;
#pragma unroll
                for (int b = 0; b < KB; ++b) tma_load_3d(dst + b * KTILE_BYTES, &maps.k, st_full + s, b * 64, j * CORR_BK, h);
                tma_load_3d(dst + K_BYTES, &maps.a, st_full + s, j * CORR_BK, m0, h);     // rows past M are zero-filled
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================
        #ifdef GD_MMA_LANE0
        if (lane == 0) {
#else
        {   // all 32 lanes walk the loop; one elected lane issues (sm100_util.cuh: umma_*_w)
#endif
            constexpr uint32_t IDESC_S = make_idesc(CORR_BM, CORR_BK, 0, 0);
            const uint32_t idesc_d = make_idesc(CORR_BM, MC, 0, 0);      // D[n, m] += P[n, k] A_e[m, k]: both operands K-major
            const uint32_t aQ = smem_addr(sQ);
            auto issue_s = [&](int j) {
                const int s = j % NS;
                mbar_wait_mma(st_full + s, (j / NS) & 1);
                if (j > 0) mbar_wait_mma(s_free, (j - 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_addr(sStage + s * STAGE_BYTES);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                    umma_ss_w(tmem + COL_S, make_desc(aQ + (ks >> 2) * QTILE_BYTES + (ks & 3) * 32, 16, 1024),
                            make_desc(aK + (ks >> 2) * KTILE_BYTES + (ks & 3) * 32, 16, 1024), IDESC_S, ks > 0);
                tc_commit_w(s_full);
            };
            mbar_wait_mma(q_full, 0);
            issue_s(0);
            for (int j = 0; j < nT; ++j) {
                if (j + 1 < nT) issue_s(j + 1);
                const int s = j % NS, b = j & 1;
                mbar_wait_mma(p_full + b, (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aA = smem_addr(sStage + s * STAGE_BYTES + K_BYTES);
#pragma unroll
                for (int kk = 0; kk < CORR_BK / 16; ++kk)
                    umma_ts_w(tmem + COL_D, tmem + COL_P + b * 32 + kk * 8, make_desc(aA + kk * 32, 16, 1024), idesc_d, (j > 0 || kk > 0));
                tc_commit_w(p_free + b);
                tc_commit_w(st_empty + s);      // S(j) (issued earlier) and D(j) are both complete when this fires: K and A_e of the stage are free
            }
            tc_commit_w(d_done);
        }
    } else {
        // ================= softmax warps 0-7: (lane quarter, key half); two threads share a base row n = TMEM lane =================
        const int quarter = warp & 3, half = warp >> 2;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const int n = n0 + quarter * 32 + lane;
        const float lse2 = p.lse[(long)h * N + n] * 1.4426950408889634f;
        const u64 sc2 = pk2(p.scale2, p.scale2), nl2 = pk2(-lse2, -lse2);
        uint32_t s_ready = 0;                                     // probe of s_full(j), issued during step j-1 (see attention_sm100.cu)
        for (int j = 0; j < nT; ++j) {
            if (!s_ready) mbar_wait(s_full, j & 1);
            tc_fence_after();
            uint32_t sr[32];
            tmem_ld32(tmem + lane_off + COL_S + half * 32, sr);
            const int b = j & 1;
            const uint32_t pf_ready = j >= 2 ? mbar_test(p_free + b, ((j - 2) >> 1) & 1) : 0u;
            tmem_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(s_free);
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float x0, x1;
                upk2(fma2(pk2u(sr[2 * c], sr[2 * c + 1]), sc2, nl2), x0, x1);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(ex2(x0), ex2(x1));
                pk[c] = *reinterpret_cast<uint32_t*>(&b2);
            }
            if (j >= 2) {
                if (!pf_ready) mbar_wait(p_free + b, ((j - 2) >> 1) & 1);     // D(j-2) has consumed this P buffer
                tc_fence_after();
            }
            s_ready = (j + 1 < nT) ? mbar_test(s_full, (j + 1) & 1) : 0u;
            tmem_st16(tmem + lane_off + COL_P + b * 32 + half * 16, pk);
            tmem_wait_st();
            tc_fence_before();
            if (lane == 0) mbar_arrive(p_full + b);
        }
        // ---- epilogue: masked (max, first argmax) over this lane quarter's 32 base rows for every inpaint row m of the chunk;
        //      the two warps of a quarter take alternate 32-column blocks of D ----
        mbar_wait(d_done, 0);
        tc_fence_after();
        float* tile = reinterpret_cast<float*>(sT) + warp * 32 * 33;
        const float mi = p.mask_in[n], mb = p.mask_bg[n];
        const int n_part = 4 * (N / CORR_BM);
        float4* out = p.partial + ((long)h * n_part + (4 * blockIdx.x + quarter)) * p.M;
        const int nbase = n0 + quarter * 32;
        for (int c0 = half * 32; c0 < MC; c0 += 64) {
            uint32_t v[32];
            if (MC - c0 >= 32) {
                tmem_ld32(tmem + lane_off + COL_D + c0, v);
            } else {                                            // MC is a multiple of 16: last half block
                tmem_ld16(tmem + lane_off + COL_D + c0, v);
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0u;
            }
            tmem_wait_ld();
            // two passes through the transpose tile: corr * M_inpaint, then corr * M_bg (generic float masks, as the reference multiplies)
            float best[2]; int arg[2];
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const float mk = pass == 0 ? mi : mb;
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 32; ++i) tile[lane * 33 + i] = __uint_as_float(v[i]) * mk;
                __syncwarp();
                float bv = -1.f; int bi = -1;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    const float x = tile[r * 33 + lane];
                    if (x > bv) { bv = x; bi = nbase + r; }
                }
                best[pass] = bv; arg[pass] = bi;
            }
            const int m = m0 + c0 + lane;
            if (c0 + lane < MC && m < p.M) out[m] = make_float4(best[0], __int_as_float(arg[0]), best[1], __int_as_float(arg[1]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

template <int D> static int launch_corr(const CorrMaps& maps, const CorrParams& p, int n_chunks, cudaStream_t st) {
    constexpr int KB = (D + 63) / 64;
    const int stage = (KB * CORR_BK * 128 + p.MC * 128 + 1023) & ~1023;
    const size_t smem = (size_t)KB * 128 * 128 + (size_t)p.n_stage * stage + 8 * 32 * 33 * 4 + 256 + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(removal_corr_sm100_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
        configured = smem;
    }
    dim3 grid(p.N / CORR_BM, n_chunks, p.H);
    removal_corr_sm100_kernel<D><<<grid, CORR_THREADS, smem, st>>>(maps, p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

}  // namespace gd

using namespace gd;

// partial[h, t, m] = masked (max, first argmax) of corr[h, m, n] = sum_k a_e[h, m, k] * A_b[h, n, k] over the t-th block of 32 base rows n,
// with A_b = softmax(scale q_b k_b^T) recomputed tile by tile from q_b, k_b and lse_b (never stored).  q_b, k_b: (H, N, d) bf16 slabs with
// qk_strides_host = {q_row, q_head, k_row, k_head} (NULL: contiguous); lse_b (H, N); a_e (H, M, ld) bf16 contiguous, ld >= N, ld % 8 == 0;
// partial (H, N / 32, M, 4).  Serves the self-attention levels: N % 128 == 0, d in {40, 80}.
extern "C" int gd_removal_corr_sm100(const void* q_b, const void* k_b, const float* lse_b, const void* a_e, int H, int M, int N, int d, float scale,
                                     int ld, const long* qk_strides, const float* mask_in, const float* mask_bg, float* partial, void* stream) {
    GD_CHECK_ARG(q_b && k_b && lse_b && a_e && mask_in && mask_bg && partial && H > 0 && M > 0);
    if (!(N % 128 == 0 && (d == 40 || d == 80)))
        return set_error(GD_ERR_UNSUPPORTED, "gd_removal_corr_sm100 serves N %% 128 == 0, d in {40, 80}; got N=%d d=%d", N, d);
    GD_CHECK_ARG(ld >= N && ld % 8 == 0);
    const long q_rs = qk_strides ? qk_strides[0] : d, q_hs = qk_strides ? qk_strides[1] : (long)N * d;
    const long k_rs = qk_strides ? qk_strides[2] : d, k_hs = qk_strides ? qk_strides[3] : (long)N * d;
    // inpaint rows in chunks of <= 256 (one UMMA N extent), balanced and rounded up to the MMA's 16
    const int n_chunks = (M + 255) / 256;
    int MC = ((M + n_chunks - 1) / n_chunks + 15) / 16 * 16;
    if (MC < 16) MC = 16;
    CorrMaps maps;
    int rc;
    if ((rc = make_map(&maps.q, q_b, N, H, d, q_rs, q_hs, CORR_BM)) != GD_OK) return rc;
    if ((rc = make_map(&maps.k, k_b, N, H, d, k_rs, k_hs, CORR_BK)) != GD_OK) return rc;
    if ((rc = make_map(&maps.a, a_e, M, H, ld, ld, (long)M * ld, MC)) != GD_OK) return rc;
    CorrParams p;
    p.lse = lse_b; p.mask_in = mask_in; p.mask_bg = mask_bg; p.partial = (float4*)partial;
    p.H = H; p.N = N; p.M = M; p.MC = MC; p.scale2 = scale * 1.4426950408889634f;
    const int kb = (d + 63) / 64;
    const int stage = (kb * CORR_BK * 128 + MC * 128 + 1023) & ~1023;
    // S 64 + P 2 x 32 + D MC columns: chunks of up to 128 inpaint rows fit 256 TMEM columns, i.e. TWO CTAs per SM (the strictly sequential
    // S -> exp -> P -> D chain of one CTA leaves both pipes idle most of the time; with the bench edit's M = 76 the 256 CTAs of a 64^2 layer then
    // also run as one wave instead of two).  Shared memory is budgeted accordingly.
    const int fixed = kb * 128 * 128 + 8 * 32 * 33 * 4 + 256 + 1024;      // Q tile, epilogue transpose tiles, barriers, alignment
    bool two = MC <= 128 && (112 * 1024 - fixed) / stage >= 2;
    p.tmem_cols = two ? 256 : 512;
    p.n_stage = two ? (112 * 1024 - fixed) / stage : (int)((184 * 1024 - kb * 128 * 128) / stage);
    if (p.n_stage > 4) p.n_stage = 4;
    if (p.n_stage < 2) return set_error(GD_ERR_UNSUPPORTED, "gd_removal_corr_sm100: stage of %d bytes does not fit twice", stage);
    if (d == 40) return launch_corr<40>(maps, p, n_chunks, (cudaStream_t)stream);
    return launch_corr<80>(maps, p, n_chunks, (cudaStream_t)stream);
}
