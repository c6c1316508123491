// geodiffuser_b200/csrc/attention_sm100.cu
//
// Subsystem (2): fused flash-style shared-attention FORWARD for sm_100a, for the self-attention levels that dominate the
// path (N = Nk in {4096, 1024, 9216, 2304}, head_dim 40 / 80).  Replaces attention_sharing.py:30-47 (`compute_attention`:
// baddbmm -> fp32 softmax over a materialised (H, N, N) map) + torch.bmm(P, V) at attention_processors.py:548-557, 643-647.
//
// One CTA = one (stream g, head h, 128-query tile); 192 threads in three roles:
//   warps 0-3  softmax / correction / epilogue: thread t owns query row t == TMEM lane t
//   warp  4    TMA producer: Q once, then K / V tiles of 128 keys through a 2-stage mbarrier ring (cp.async.bulk.tensor, SWIZZLE_128B)
//   warp  5    MMA issuer (one elected lane): S = Q K^T  (tcgen05.mma, A/B from shared memory, fp32 accumulator in TMEM),
//                                              O += P V   (A = bf16 P read from TMEM, B = V from shared memory, MN-major)
// TMEM columns: [0,128) S  |  [128,192) P (bf16 pairs)  |  [192, 192+DV) O.   S is consumed into registers and released before the
// exponentials run, so QK^T of tile j+1 overlaps softmax of tile j; O is rescaled in TMEM only when the running max grows by more
// than 2^8 (lazy rescaling), which is exact after the final normalisation.  head_dim 40: two CTAs per SM (80 KB smem, 256 TMEM
// columns each) so one CTA's MMAs fill the other's softmax phase.
//
// Roofline: tensor pipe (dense BF16) -- but at head_dim 40 a 128x128 tile costs 16384 exp2 (MUFU, 16/clk/SM = 1024 clk) against
// ~384 clk of UMMA, so this shape is bounded by the special-function unit, not the tensor cores.
#include "sm100_util.cuh"

namespace gd {

constexpr int SM100_MAXG = 8;
constexpr int SM100_THREADS = 192;
constexpr int BM = 128, BN = 128;

struct Sm100Maps {
    CUtensorMap q[SM100_MAXG];
    CUtensorMap k[SM100_MAXG];
    CUtensorMap v[SM100_MAXG];
};
struct Sm100Params {
    float* o[SM100_MAXG];      // (H, N, d) fp32 contiguous, or NULL
    float* lse[SM100_MAXG];
    void* os[SM100_MAXG];      // strided output (row stride os_rs, head stride os_hs, elements; bf16 or fp32), or NULL
    long os_rs, os_hs;
    int os_bf16;
    int H, N, d;
    float scale2;  // scale * log2(e)
#ifdef GD_TRACE
    long long* trace;   // scripts/fwd_trace.cu: clock64 stamps of one CTA's warps, [role][step][8]
    int trace_x;
#endif
};
#ifdef GD_TRACE
#define GD_TR(role, j, slot) do { if (p.trace && blockIdx.x == p.trace_x && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0) \
    p.trace[((role) * 256 + (j)) * 8 + (slot)] = clock64(); } while (0)
#else
#define GD_TR(role, j, slot) do { } while (0)
#endif

// D = head_dim (40 or 80).  KB = number of 64-wide (128-byte) column blocks per operand row; KSTEPS = ceil(D/16); DV = O columns.
// BNK = keys per step (128 or 64).  Packed fp32x2 softmax arithmetic, NP of every 8 score pairs on the FMA-pipe polynomial, the rest on the MUFU.
//
// BNK = 128: TMEM [0,128) S | [128,192) P | [192,192+DV) O in one allocation of 256 (head_dim 40: two CTAs per SM) or 512 columns (head_dim 80).
// BNK =  64: S 64 + P 32 + O DV columns.  head_dim 40: 144 columns, taken as TWO allocations (128: S | O, and 32: P) so that THREE CTAs fit the
//            512 columns of an SM (allocations are powers of two; 3 x 160 = 480) -- twelve softmax warps per SM instead of eight, three per
//            scheduler: the softmax chain of a tile (wait, tcgen05.ld, row max, exponentials, tcgen05.st, handshake) keeps a warp off the MUFU
//            for more than half of its time, and a third resident chain is what fills that pipe.  A CTA that blocks in tcgen05.alloc waits for
//            a CTA that never waits for it, so the split allocation cannot deadlock; shared memory (64 KB per CTA) keeps a fourth CTA out.
//            head_dim 80: 176 columns -> one allocation of 256, two CTAs per SM (one at BNK = 128).
template <int D, int BNK> struct FwdCfg {
    static constexpr int KSTEPS = (D + 15) / 16;
    static constexpr int DV = KSTEPS * 16;
    static constexpr bool SPLIT_ALLOC = (BNK == 64 && DV <= 64);
    static constexpr int CTAS = (BNK == 64) ? (DV <= 64 ? 3 : 2) : (192 + DV <= 256 ? 2 : 1);
    static constexpr int NSTAGE = SPLIT_ALLOC ? 3 : 2;
    static constexpr int TMEM_COLS = SPLIT_ALLOC ? 128 : (BNK == 64 ? 256 : (192 + DV <= 256 ? 256 : 512));
    static constexpr uint32_t COL_S = 0;
    static constexpr uint32_t COL_O = SPLIT_ALLOC ? 64 : (BNK == 64 ? 96 : 192);
    static constexpr uint32_t COL_P = SPLIT_ALLOC ? 0 : (BNK == 64 ? 64 : 128);      // SPLIT_ALLOC: relative to the second allocation
    static constexpr bool ONES = DV > D;             // head_dim 40: column 40 of the zero-filled V padding is set to 1, so that O[:, 40] = row sum of P
    static constexpr int KB = (D + 63) / 64;
    static constexpr size_t SMEM = (size_t)KB * 128 * 128 + (size_t)NSTAGE * 2 * KB * BNK * 128 + 256 + 1024;
};

template <int D, int BNK, int NP>
__global__ void __launch_bounds__(SM100_THREADS, (FwdCfg<D, BNK>::CTAS))
attn_fwd_sm100_kernel(const __grid_constant__ Sm100Maps maps, const Sm100Params p) {
    typedef FwdCfg<D, BNK> C;
    constexpr int KB = C::KB;
    constexpr int KSTEPS = C::KSTEPS;
    constexpr int DV = C::DV;                       // 48 / 80: UMMA N (multiple of 16 at M = 128)
    constexpr int QTILE_BYTES = 128 * 128;          // one [128 rows][64 bf16] swizzled block of Q
    constexpr int KTILE_BYTES = BNK * 128;          // one [BNK rows][64 bf16] swizzled block of K / V
    constexpr int Q_BYTES = KB * QTILE_BYTES, K_BYTES = KB * KTILE_BYTES;
    constexpr int NSTAGE = C::NSTAGE;
#ifdef GD_FWD_NOPROBE
    constexpr bool PROBE = false;
#else
    constexpr bool PROBE = true;
#endif

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + Q_BYTES;
    unsigned char* sV = sK + NSTAGE * K_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSTAGE * K_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* s_full = bars + 1;
    uint64_t* s_free = bars + 2;
    uint64_t* p_full = bars + 3;
    uint64_t* pv_done = bars + 4;
    uint64_t* k_full = bars + 5;                  // [NSTAGE]
    uint64_t* k_empty = k_full + NSTAGE;
    uint64_t* v_full = k_empty + NSTAGE;
    uint64_t* v_empty = v_full + NSTAGE;
    uint64_t* v_tma = v_empty + NSTAGE;           // [NSTAGE]  (ONES: the V tile has landed; v_full follows once its ones column is written)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(v_tma + NSTAGE);   // [2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
    const int N = p.N;
    const int nT = N / BNK;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); mbar_init(v_tma + s, 1); }
        mbar_init(s_full, 1); mbar_init(s_free, 4); mbar_init(p_full, 4); mbar_init(pv_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(C::TMEM_COLS) : "memory");
        if constexpr (C::SPLIT_ALLOC)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot + 1)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 4 && lane == 0) { tma_prefetch_desc(&maps.q[g]); tma_prefetch_desc(&maps.k[g]); tma_prefetch_desc(&maps.v[g]); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot[0], 0);
    const uint32_t tS = tmem + C::COL_S, tO = tmem + C::COL_O, tP = (C::SPLIT_ALLOC ? __shfl_sync(0xffffffffu, tmem_slot[1], 0) : tmem) + C::COL_P;

    if (warp == 4) {
        // ================= TMA producer (runs up to NSTAGE stages ahead: its waits are not latency critical) =================
        // ONES (head_dim 40): TMA zero-fills columns 40..63 of every V row; the producer warp then writes 1.0 into column 40 of the landed tile, so
        // that the P V product accumulates the row sum of P in O[:, 40] -- the softmax warps do not add up their exponentials (64 FADD2 per row
        // and tile less on the warps that bound the kernel), and numerator and denominator see the same bf16-rounded P.
        auto patch_v = [&](int j) {
            const int s = j % NSTAGE;
            mbar_wait_relaxed(v_tma + s, (j / NSTAGE) & 1);
            unsigned char* vt = sV + s * K_BYTES + (D / 64) * KTILE_BYTES;
            constexpr int CH = ((D % 64) * 2) / 16, OFF = ((D % 64) * 2) % 16;       // 16-byte chunk of column D inside its 128-byte row, SWIZZLE_128B
#pragma unroll
            for (int r = lane; r < BNK; r += 32)
                *reinterpret_cast<unsigned short*>(vt + r * 128 + ((CH ^ (r & 7)) * 16) + OFF) = 0x3F80;     // bf16 1.0
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(v_full + s);
        };
        if (lane == 0) {
            mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
            for (int b = 0; b < KB; ++b) tma_load_3d(sQ + b * QTILE_BYTES, &maps.q[g], q_full, b * 64, q0, h);
        }
        for (int j = 0; j < nT; ++j) {
            if (lane == 0) {
                const int s = j % NSTAGE;
                const uint32_t ph = (j / NSTAGE) & 1;
                mbar_wait_relaxed(k_empty + s, ph ^ 1);
                GD_TR(5, j, 0);
                mbar_expect_tx(k_full + s, K_BYTES);
#pragma unroll
                for (int b = 0; b < KB; ++b) tma_load_3d(sK + s * K_BYTES + b * KTILE_BYTES, &maps.k[g], k_full + s, b * 64, j * BNK, h);
                mbar_wait_relaxed(v_empty + s, ph ^ 1);
                GD_TR(5, j, 1);
                uint64_t* vb = C::ONES ? v_tma + s : v_full + s;
                mbar_expect_tx(vb, K_BYTES);
#pragma unroll
                for (int b = 0; b < KB; ++b) tma_load_3d(sV + s * K_BYTES + b * KTILE_BYTES, &maps.v[g], vb, b * 64, j * BNK, h);
            }
            __syncwarp();
            if (C::ONES && j > 0) patch_v(j - 1);     // V(j-1) landed a step ago; P V(j-1) is issued at the end of softmax(j-1): far off the critical path
        }
        if (C::ONES) patch_v(nT - 1);
    } else if (warp == 5) {
        // ================= MMA issuer =================
#ifdef GD_MMA_LANE0
        if (lane == 0) {
#else
        {   // all 32 lanes walk the loop; one elected lane issues (umma_*_w, tc_commit_w)
#endif
            constexpr uint32_t IDESC_QK = make_idesc(BM, BNK, 0, 0);
            constexpr uint32_t IDESC_PV = make_idesc(BM, DV, 0, 1);
            const uint32_t aQ = smem_addr(sQ);
            auto issue_qk = [&](int j) {
                const int s = j % NSTAGE;
                GD_TR(4, j, 0);
                mbar_wait_mma(k_full + s, (j / NSTAGE) & 1);
                GD_TR(4, j, 1);
                if (j > 0) mbar_wait_mma(s_free, (j - 1) & 1);
                GD_TR(4, j, 2);
                tc_fence_after();
                const uint32_t aK = smem_addr(sK + s * K_BYTES);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)     // 64-wide block, then 32 B per 16 elements
                    umma_ss_w(tS, make_desc(aQ + (ks >> 2) * QTILE_BYTES + (ks & 3) * 32, 16, 1024),
                            make_desc(aK + (ks >> 2) * KTILE_BYTES + (ks & 3) * 32, 16, 1024), IDESC_QK, ks > 0);
                tc_commit_w(s_full);        // S(j) complete -> softmax
                tc_commit_w(k_empty + s);   // K stage reusable
            };
            mbar_wait_mma(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < nT; ++j) {
                if (j + 1 < nT) issue_qk(j + 1);
                const int s = j % NSTAGE;
                GD_TR(4, j, 3);
                mbar_wait_mma(v_full + s, (j / NSTAGE) & 1);
                GD_TR(4, j, 4);
                mbar_wait_mma(p_full, j & 1);
                GD_TR(4, j, 5);
                tc_fence_after();
                const uint32_t aV = smem_addr(sV + s * K_BYTES);
#pragma unroll
                for (int kk = 0; kk < BNK / 16; ++kk) {
                    // V tile rows = keys (the MMA K dimension), MN-major: 16 keys = 16 rows x 128 B; LBO = next 64-wide column block
                    umma_ts_w(tO, tP + kk * 8, make_desc(aV + kk * 2048, KTILE_BYTES, 1024), IDESC_PV, (j > 0 || kk > 0));
                }
                tc_commit_w(pv_done);
                tc_commit_w(v_empty + s);
                GD_TR(4, j, 6);
            }
        }
    } else {
        // ================= softmax / correction / epilogue (warps 0-3): thread t owns query row t == TMEM lane t =================
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const float scale2 = p.scale2;
        float m_run = -INFINITY, l_run = 0.f;
        uint32_t s_ready = 0;                        // probe of s_full(j), issued during step j-1
        for (int j = 0; j < nT; ++j) {
            GD_TR(warp, j, 0);
            if (!s_ready) mbar_wait(s_full, j & 1);
            GD_TR(warp, j, 1);
            tc_fence_after();
            uint32_t sr[BNK];
#pragma unroll
            for (int c = 0; c < BNK / 32; ++c) tmem_ld32(tS + lane_off + c * 32, sr + c * 32);
            // pv_done(j-1) is needed after the first chunk of exponentials: probe it now, so that the ~120 clk even a satisfied mbarrier wait
            // costs (scripts/fwd_trace.cu) run under the row max and that chunk instead of stalling the warp there
            const uint32_t pv_ready = (PROBE && j > 0) ? mbar_test(pv_done, (j - 1) & 1) : 0u;
            tmem_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(s_free);      // S(j) is in registers: QK^T(j+1) may overwrite it
            GD_TR(warp, j, 2);
            // row max: four independent FMNMX3 chains
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < BNK; c += 8) {
                mx0 = max3(mx0, __uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
                mx1 = max3(mx1, __uint_as_float(sr[c + 2]), __uint_as_float(sr[c + 3]));
                mx2 = max3(mx2, __uint_as_float(sr[c + 4]), __uint_as_float(sr[c + 5]));
                mx3 = max3(mx3, __uint_as_float(sr[c + 6]), __uint_as_float(sr[c + 7]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const float m_cand = fmaxf(m_run, mx * scale2);
            const bool grow = (m_cand - m_run) > 8.0f;           // lazy rescale: tolerate up to 2^8 headroom
            const float m_new = grow ? m_cand : m_run;
            const float alpha = grow ? ex2(m_run - m_new) : 1.0f;
            m_run = m_new;
            u64 rsA = 0ull, rsB = 0ull;                           // packed row-sum accumulators (two chains); unused with the ones column
            const u64 sc2 = pk2(scale2, scale2), nm2 = pk2(-m_new, -m_new);
            s_ready = 0;
#pragma unroll
            for (int cc = 0; cc < BNK / 32; ++cc) {
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int e = cc * 32 + 2 * c;
                    float p0, p1;
                    const u64 x2 = fma2(pk2u(sr[e], sr[e + 1]), sc2, nm2);
                    u64 p2;
                    if (pair_is_poly<NP>(c)) {
                        p2 = ex2_poly2(x2);
                        upk2(p2, p0, p1);
                    } else {
                        float x0, x1;
                        upk2(x2, x0, x1);
                        p0 = ex2(x0);
                        p1 = ex2(x1);
                        p2 = pk2(p0, p1);
                    }
                    if constexpr (!C::ONES) { if (c & 1) rsB = add2(rsB, p2); else rsA = add2(rsA, p2); }
                    __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
                    pk[c] = *reinterpret_cast<uint32_t*>(&b2);
                }
                if (cc == 0 && j > 0) {
                    // P(j) may overwrite P(j-1), and O may be corrected, only once PV(j-1) has completed.  Waiting here -- one chunk of
                    // exponentials after the row max -- instead of in front of the exponentials keeps the MMA round trip
                    // (p_full -> PV issue -> commit) off the softmax warps' critical path.
                    GD_TR(warp, j, 3);
                    if (!pv_ready) mbar_wait(pv_done, (j - 1) & 1);
                    GD_TR(warp, j, 4);
                    tc_fence_after();
                    if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
                        for (int c = 0; c < DV / 16; ++c) {
                            uint32_t orr[16];
                            tmem_ld16(tO + lane_off + c * 16, orr);
                            tmem_wait_ld();
#pragma unroll
                            for (int e = 0; e < 16; ++e) orr[e] = __float_as_uint(__uint_as_float(orr[e]) * alpha);
                            tmem_st16(tO + lane_off + c * 16, orr);
                        }
                    }
                }
                if (PROBE && cc == BNK / 32 - 1 && j + 1 < nT) s_ready = mbar_test(s_full, (j + 1) & 1);   // consumed at the top of step j+1
                tmem_st16(tP + lane_off + cc * 16, pk);
            }
            if constexpr (!C::ONES) {
                float a0, a1;
                upk2(add2(rsA, rsB), a0, a1);
                l_run = l_run * alpha + (a0 + a1);
            }
            GD_TR(warp, j, 5);
            tmem_wait_st();
            tc_fence_before();
            if (lane == 0) mbar_arrive(p_full);
            GD_TR(warp, j, 6);
        }
        // epilogue: O / l -> global fp32, lse
        mbar_wait(pv_done, (nT - 1) & 1);
        tc_fence_after();
        const int row = q0 + warp * 32 + lane;
        if constexpr (C::ONES) {                     // the row sum sits in O[:, D] (it took part in every lazy rescale of O)
            uint32_t orr[16];
            tmem_ld16(tO + lane_off + (D / 16) * 16, orr);
            tmem_wait_ld();
            l_run = __uint_as_float(orr[D % 16]);
        }
        const float inv = 1.0f / l_run;
        float* og = p.o[g] ? p.o[g] + ((long)h * N + row) * D : nullptr;
        unsigned char* sg = p.os[g] ? reinterpret_cast<unsigned char*>(p.os[g]) + ((long)h * p.os_hs + (long)row * p.os_rs) * (p.os_bf16 ? 2 : 4) : nullptr;
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
            uint32_t orr[16];
            tmem_ld16(tO + lane_off + c * 16, orr);
            tmem_wait_ld();
            float f[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(orr[e]) * inv;
            if (og) {
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    if (c * 16 + e < D) *reinterpret_cast<float4*>(og + c * 16 + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
            }
            if (sg) {
                if (p.os_bf16) {
#pragma unroll
                    for (int e = 0; e < 16; e += 8)
                        if (c * 16 + e < D) {
                            uint4 v;
                            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[e], f[e + 1]), b1 = __floats2bfloat162_rn(f[e + 2], f[e + 3]);
                            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[e + 4], f[e + 5]), b3 = __floats2bfloat162_rn(f[e + 6], f[e + 7]);
                            v.x = *reinterpret_cast<uint32_t*>(&b0); v.y = *reinterpret_cast<uint32_t*>(&b1);
                            v.z = *reinterpret_cast<uint32_t*>(&b2); v.w = *reinterpret_cast<uint32_t*>(&b3);
                            *reinterpret_cast<uint4*>(sg + (c * 16 + e) * 2) = v;
                        }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        if (c * 16 + e < D) *reinterpret_cast<float4*>(sg + (c * 16 + e) * 4) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
                }
            }
        }
        p.lse[g][(long)h * N + row] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C::TMEM_COLS) : "memory");
        if constexpr (C::SPLIT_ALLOC)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot[1]), "r"(32) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Subsystem (3): the matching BACKWARD for the self-attention levels, dQ only -- on this path every K / V is the detached base sample's
// (attention_sharing.py:242), so dK / dV do not exist for self layers.  Replaces autograd through compute_attention + bmm
// (materialised (H, N, N) maps kept for backward in the reference) and, through `extra`, the removal-loss term's dense `dcorr . A_b`.
//
//   S = Q K^T            (tcgen05 SS)        P  = exp2(S * scale2 - lse2)                      (MUFU / FMA-pipe polynomial)
//   dP = dO V^T          (tcgen05 SS)        dS = P o (dP (+ extra) - delta)   -> bf16 in TMEM
//   dQ += dS K           (tcgen05, A = dS from TMEM, B = the K tile already in shared memory read MN-major)     dQ *= scale at the end
//
// One CTA per (head, 128-query tile), one CTA per SM (TMEM: S 128 + dP 128 + dS 64 + dQ <= 80 columns).  Ten warps: 0-7 elementwise (warp w
// owns TMEM lanes 32*(w%4).., key columns 64*(w/4)..: the backward needs no row reductions, so two warps share a row freely), 8 = TMA,
// 9 = MMA issuer.  S and dP are released as soon as they sit in registers, so the two score GEMMs of tile j+1 run under tile j's exponentials.
constexpr int SM100_BWD_THREADS = 320;

struct Sm100BwdMaps { CUtensorMap q, k, v, d_o; };
struct Sm100BwdParams {
    const float* lse; const float* delta;
    const float* extra; const float* extra_scale; const int* rowmap; int ex_ld, M, ex_t;   // ex_t: extra is key-major (H, N, ex_ld)
    void* dq;                  // strided (dq_rs, dq_hs elements), fp32 or bf16
    long dq_rs, dq_hs;
    int dq_bf16;
    int H, N;
    float scale, scale2;
};

// ---------------------------------------------------------------------------------------------------------------------------
// The kernel: one CTA per (head, 128-query tile), ten warps (0-7 elementwise: lane quarter x key half, 8 = TMA, 9 = MMA issuer), TWO CTAs per SM.
// 64-key steps: TMEM per CTA = S 64 + dP 64 + dS 32 + dQ <= 80 = 240 -> 256 columns; shared memory at head_dim 40 = Q 16 K + dO 16 K +
// K ring 4 x 8 K + V ring 2 x 8 K = 80 KB.  S and dP are released as soon as they sit in registers, so the two score GEMMs of step j+1 run
// under step j's exponentials; a K tile is needed at both ends of its step (scores, then dQ), hence its deeper ring.  (Round 1 ran 128-key
// steps with one CTA per SM: same speed without removal rows -- 75 us at H=8, N=4096 -- but 256 CTAs on 148 SMs ran as two waves and the
// CTAs that own removal-loss rows set the makespan: 118 us against 85 us here; history in git.)  Per-score arithmetic is packed
// (FFMA2 / FADD2 / FMUL2): 3 issue slots per score.
template <int D, int NP>
__global__ void __launch_bounds__(SM100_BWD_THREADS, 2)
attn_bwd64_sm100_kernel(const __grid_constant__ Sm100BwdMaps maps, const Sm100BwdParams p) {
    constexpr int KB = (D + 63) / 64;
    constexpr int KSTEPS = (D + 15) / 16;
    constexpr int DV = KSTEPS * 16;
    constexpr int BNK = 64;                              // keys per step
    constexpr int QTILE_BYTES = 128 * 128;               // one [128 rows][64 bf16] swizzled block of Q / dO
    constexpr int KTILE_BYTES = BNK * 128;               // one [64 rows][64 bf16] swizzled block of K / V
    constexpr int Q_BYTES = KB * QTILE_BYTES, K_BYTES = KB * KTILE_BYTES;
    constexpr uint32_t COL_S = 0, COL_DP = 64, COL_DS = 128, COL_DQ = 160;
    constexpr int TMEM_COLS = 256;
    constexpr int NSTAGE = 2;                            // V ring
    constexpr int KSTAGE = 4;                            // K ring (a K tile is read at both ends of its step)
    constexpr int NEW = 8;                               // elementwise warps: (lane quarter, key half)
    constexpr int NC = BNK / 2;                          // 32 key columns per elementwise thread

    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = smem;
    unsigned char* sDO = sQ + Q_BYTES;
    unsigned char* sK = sDO + Q_BYTES;
    unsigned char* sV = sK + KSTAGE * K_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NSTAGE * K_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* v_full = bars + 1;    // [2]
    uint64_t* v_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* s_free = bars + 6;
    uint64_t* ds_full = bars + 7;
    uint64_t* dq_done = bars + 8;
    uint64_t* k_full = bars + 9;              // [KSTAGE]
    uint64_t* k_empty = bars + 9 + KSTAGE;    // [KSTAGE]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9 + 2 * KSTAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y, q0 = blockIdx.x * BM;
    const int N = p.N;
    const int nT = N / BNK;

    if (threadIdx.x == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(v_full + s, 1); mbar_init(v_empty + s, 1); }
        for (int s = 0; s < KSTAGE; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
        mbar_init(s_full, 1); mbar_init(s_free, NEW); mbar_init(ds_full, NEW); mbar_init(dq_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NEW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == NEW && lane == 0) { tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.d_o); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == NEW) {
        // ================= TMA producer =================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * Q_BYTES);
#pragma unroll
            for (int b = 0; b < KB; ++b) {
                tma_load_3d(sQ + b * QTILE_BYTES, &maps.q, q_full, b * 64, q0, h);
                tma_load_3d(sDO + b * QTILE_BYTES, &maps.d_o, q_full, b * 64, q0, h);
            }
            for (int j = 0; j < nT; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                const int ks = j % KSTAGE;
                const uint32_t kph = (j / KSTAGE) & 1;
                mbar_wait_relaxed(k_empty + ks, kph ^ 1);
                mbar_expect_tx(k_full + ks, K_BYTES);
#pragma unroll
                for (int b = 0; b < KB; ++b) tma_load_3d(sK + ks * K_BYTES + b * KTILE_BYTES, &maps.k, k_full + ks, b * 64, j * BNK, h);
                mbar_wait_relaxed(v_empty + s, ph ^ 1);
                mbar_expect_tx(v_full + s, K_BYTES);
#pragma unroll
                for (int b = 0; b < KB; ++b) tma_load_3d(sV + s * K_BYTES + b * KTILE_BYTES, &maps.v, v_full + s, b * 64, j * BNK, h);
            }
        }
    } else if (warp == NEW + 1) {
        // ================= MMA issuer =================
#ifdef GD_MMA_LANE0
        if (lane == 0) {
#else
        {   // all 32 lanes walk the loop; one elected lane issues (sm100_util.cuh: umma_*_w)
#endif
            constexpr uint32_t IDESC_SS = make_idesc(BM, BNK, 0, 0);
            constexpr uint32_t IDESC_DQ = make_idesc(BM, DV, 0, 1);
            const uint32_t aQ = smem_addr(sQ), aDO = smem_addr(sDO);
            auto issue_scores = [&](int j) {
                const int s = j & 1;
                const int kst = j % KSTAGE;
                mbar_wait_mma(k_full + kst, (j / KSTAGE) & 1);
                mbar_wait_mma(v_full + s, (j >> 1) & 1);
                if (j > 0) mbar_wait_mma(s_free, (j - 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_addr(sK + kst * K_BYTES), aV = smem_addr(sV + s * K_BYTES);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                    umma_ss_w(tmem + COL_S, make_desc(aQ + (ks >> 2) * QTILE_BYTES + (ks & 3) * 32, 16, 1024),
                            make_desc(aK + (ks >> 2) * KTILE_BYTES + (ks & 3) * 32, 16, 1024), IDESC_SS, ks > 0);
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                    umma_ss_w(tmem + COL_DP, make_desc(aDO + (ks >> 2) * QTILE_BYTES + (ks & 3) * 32, 16, 1024),
                            make_desc(aV + (ks >> 2) * KTILE_BYTES + (ks & 3) * 32, 16, 1024), IDESC_SS, ks > 0);
                tc_commit_w(s_full);
                tc_commit_w(v_empty + s);
            };
            mbar_wait_mma(q_full, 0);
            issue_scores(0);
            for (int j = 0; j < nT; ++j) {
                if (j + 1 < nT) issue_scores(j + 1);
                const int s = j % KSTAGE;
                mbar_wait_mma(ds_full, j & 1);
                tc_fence_after();
                const uint32_t aK = smem_addr(sK + s * K_BYTES);
#pragma unroll
                for (int kk = 0; kk < BNK / 16; ++kk)
                    umma_ts_w(tmem + COL_DQ, tmem + COL_DS + kk * 8, make_desc(aK + kk * 2048, KTILE_BYTES, 1024), IDESC_DQ, (j > 0 || kk > 0));
                tc_commit_w(dq_done);
                tc_commit_w(k_empty + s);
            }
        }
    } else {
        // ================= elementwise warps 0-7 =================
        const int quarter = warp & 3, half = warp >> 2;
        const int row = q0 + quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const float lse2 = p.lse[(long)h * N + row] * 1.4426950408889634f;
        const float delta = p.delta[(long)h * N + row];
        const u64 sc2 = pk2(p.scale2, p.scale2), nl2 = pk2(-lse2, -lse2), nd2 = pk2(-delta, -delta);
        const int slot = p.rowmap ? p.rowmap[row] : -1;
        const float ex_scale = (p.extra && p.extra_scale) ? *p.extra_scale : 1.0f;
        // this thread's removal-loss row: row-major extra (H, M, ex_ld): 32 consecutive floats per step; key-major (H, N, ex_ld): one float per
        // key, ex_ld apart, consecutive slots (= consecutive lanes) adjacent in memory
        const long ex_ks = p.ex_t ? p.ex_ld : 1;
        const float* exrow = (slot >= 0) ? (p.ex_t ? p.extra + ((long)h * N + half * NC) * p.ex_ld + slot
                                                   : p.extra + ((long)h * p.M + slot) * p.ex_ld + half * NC) : nullptr;
        const bool any_ex = __any_sync(0xffffffffu, exrow != nullptr);   // warp-uniform: does this warp own removal-loss rows at all
        // one pair of scores -> one packed bf16 pair of dS = P o (dP - delta)
        auto ds_pair = [&](uint32_t s0, uint32_t s1, uint32_t g0, uint32_t g1, int c) -> uint32_t {
            const u64 x2 = fma2(pk2u(s0, s1), sc2, nl2);
            u64 p2;
            if (pair_is_poly<NP>(c)) {
                p2 = ex2_poly2(x2);
            } else {
                float x0, x1;
                upk2(x2, x0, x1);
                p2 = pk2(ex2(x0), ex2(x1));
            }
            float d0, d1;
            upk2(mul2(p2, add2(pk2u(g0, g1), nd2)), d0, d1);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
            return *reinterpret_cast<uint32_t*>(&b2);
        };
        if (!any_ex) {
            uint32_t s_ready = 0;                                 // probe of s_full(j), issued during step j-1 (see the forward kernel)
            for (int j = 0; j < nT; ++j) {
                if (!s_ready) mbar_wait(s_full, j & 1);
                tc_fence_after();
                uint32_t sr[NC], dp[NC];
                tmem_ld32(tmem + lane_off + COL_S + half * NC, sr);
                tmem_ld32(tmem + lane_off + COL_DP + half * NC, dp);
                const uint32_t dq_ready = j > 0 ? mbar_test(dq_done, (j - 1) & 1) : 0u;     // needed only after the step's arithmetic
                tmem_wait_ld();
                tc_fence_before();
                if (lane == 0) mbar_arrive(s_free);
                uint32_t pk[NC / 2];
#pragma unroll
                for (int c = 0; c < NC / 2; ++c) pk[c] = ds_pair(sr[2 * c], sr[2 * c + 1], dp[2 * c], dp[2 * c + 1], c);
                if (j > 0) {
                    if (!dq_ready) mbar_wait(dq_done, (j - 1) & 1);              // dS(j-1) has been consumed by its dQ product
                    tc_fence_after();
                }
                s_ready = (j + 1 < nT) ? mbar_test(s_full, (j + 1) & 1) : 0u;
                tmem_st16(tmem + lane_off + COL_DS + half * (NC / 2), pk);
                tmem_wait_st();
                tc_fence_before();
                if (lane == 0) mbar_arrive(ds_full);
            }
        } else {
            // Warps that own removal-loss rows: dL/dP of the row (`extra`) joins dP.  Its 32 floats per step come from L2 (~700 clk): they are
            // requested one step AHEAD, in two groups of 16 that are re-issued as soon as the group has been folded into dP, and the step itself
            // is walked in two 16-column halves so that the prefetch registers fit under the two-CTAs-per-SM register budget.  (Loading them
            // inside the arithmetic made the few CTAs that own inpaint rows set the kernel's makespan: 118 us against 75 us without rows.)
            float ev[NC];
            auto fetch = [&](int j, int g) {
                if (exrow && j < nT) {
                    const float* e0 = exrow + ((long)j * BNK + g * 16) * ex_ks;
#pragma unroll
                    for (int i = 0; i < 16; ++i) ev[g * 16 + i] = __ldg(e0 + i * ex_ks);
                }
            };
            fetch(0, 0);
            fetch(0, 1);
            for (int j = 0; j < nT; ++j) {
                mbar_wait(s_full, j & 1);
                tc_fence_after();
                uint32_t pk[NC / 2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t sr[16], dp[16];
                    tmem_ld16(tmem + lane_off + COL_S + half * NC + g * 16, sr);
                    tmem_ld16(tmem + lane_off + COL_DP + half * NC + g * 16, dp);
                    tmem_wait_ld();
                    if (g == 1) {
                        tc_fence_before();
                        if (lane == 0) mbar_arrive(s_free);
                    }
                    if (exrow) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dp[i] = __float_as_uint(fmaf(ex_scale, ev[g * 16 + i], __uint_as_float(dp[i])));
                    }
                    fetch(j + 1, g);
#pragma unroll
                    for (int c = 0; c < 8; ++c) pk[g * 8 + c] = ds_pair(sr[2 * c], sr[2 * c + 1], dp[2 * c], dp[2 * c + 1], g * 8 + c);
                }
                if (j > 0) {
                    mbar_wait(dq_done, (j - 1) & 1);
                    tc_fence_after();
                }
                tmem_st16(tmem + lane_off + COL_DS + half * (NC / 2), pk);
                tmem_wait_st();
                tc_fence_before();
                if (lane == 0) mbar_arrive(ds_full);
            }
        }
        // epilogue: dQ * scale -> global; the two warps of a quarter split the 16-column chunks
        mbar_wait(dq_done, (nT - 1) & 1);
        tc_fence_after();
        unsigned char* og = reinterpret_cast<unsigned char*>(p.dq) + ((long)h * p.dq_hs + (long)row * p.dq_rs) * (p.dq_bf16 ? 2 : 4);
        const float sc = p.scale;
#pragma unroll
        for (int c = 0; c < DV / 16; ++c) {
            if ((c & 1) != half) continue;
            uint32_t orr[16];
            tmem_ld16(tmem + lane_off + COL_DQ + c * 16, orr);
            tmem_wait_ld();
            float f[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(orr[e]) * sc;
            if (p.dq_bf16) {
#pragma unroll
                for (int e = 0; e < 16; e += 8)
                    if (c * 16 + e < D) {
                        uint4 v;
                        __nv_bfloat162 b0 = __floats2bfloat162_rn(f[e], f[e + 1]), b1 = __floats2bfloat162_rn(f[e + 2], f[e + 3]);
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[e + 4], f[e + 5]), b3 = __floats2bfloat162_rn(f[e + 6], f[e + 7]);
                        v.x = *reinterpret_cast<uint32_t*>(&b0); v.y = *reinterpret_cast<uint32_t*>(&b1);
                        v.z = *reinterpret_cast<uint32_t*>(&b2); v.w = *reinterpret_cast<uint32_t*>(&b3);
                        *reinterpret_cast<uint4*>(og + (c * 16 + e) * 2) = v;
                    }
            } else {
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                    if (c * 16 + e < D) *reinterpret_cast<float4*>(og + (c * 16 + e) * 4) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------------
static int g_bwd_np = 1;
static int g_np = 2;     // tuning knob (gd_attn_sm100_config): g_np of every 8 score pairs on the FMA-pipe polynomial
static int g_bnk = 0;    // keys per step of the forward: 0 = per head_dim (128 at head_dim 40: two CTAs per SM; 64 at head_dim 80: two CTAs per SM
                         // instead of one, 19.7 against 24.6 us at G=3, N=1024), or 64 / 128 forced (A/B measurements)

template <int D, int BNK, int NP> static int launch_sm100(const Sm100Maps& maps, const Sm100Params& p, int G, cudaStream_t st) {
    const size_t smem = FwdCfg<D, BNK>::SMEM;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_sm100_kernel<D, BNK, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(p.N / BM, p.H, G);
    attn_fwd_sm100_kernel<D, BNK, NP><<<grid, SM100_THREADS, smem, st>>>(maps, p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

template <int D, int BNK> static int dispatch_sm100(const Sm100Maps& maps, const Sm100Params& p, int G, cudaStream_t st) {
    switch (g_np) {
        case 0: return launch_sm100<D, BNK, 0>(maps, p, G, st);
        case 1: return launch_sm100<D, BNK, 1>(maps, p, G, st);
        case 2: return launch_sm100<D, BNK, 2>(maps, p, G, st);
        case 3: return launch_sm100<D, BNK, 3>(maps, p, G, st);
        case 4: return launch_sm100<D, BNK, 4>(maps, p, G, st);
    }
    return set_error(GD_ERR_UNSUPPORTED, "gd_attn_sm100_config: no kernel instance for np=%d", g_np);
}

template <int D, int NP> static int launch_bwd64_sm100(const Sm100BwdMaps& maps, const Sm100BwdParams& p, cudaStream_t st) {
    constexpr int KB = (D + 63) / 64;
    const size_t smem = (size_t)2 * KB * 128 * 128 + (size_t)6 * KB * 64 * 128 + 256 + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_bwd64_sm100_kernel<D, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(GD_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(p.N / BM, p.H, 1);
    attn_bwd64_sm100_kernel<D, NP><<<grid, SM100_BWD_THREADS, smem, st>>>(maps, p);
    GD_CHECK_LAUNCH();
    return GD_OK;
}

template <int D> static int dispatch_bwd64_sm100(const Sm100BwdMaps& maps, const Sm100BwdParams& p, cudaStream_t st) {
    switch (g_bwd_np) {
        case 0: return launch_bwd64_sm100<D, 0>(maps, p, st);
        case 1: return launch_bwd64_sm100<D, 1>(maps, p, st);
        case 2: return launch_bwd64_sm100<D, 2>(maps, p, st);
        case 3: return launch_bwd64_sm100<D, 3>(maps, p, st);
        case 4: return launch_bwd64_sm100<D, 4>(maps, p, st);
    }
    return set_error(GD_ERR_UNSUPPORTED, "gd_attn_sm100_config: no backward instance for np=%d", g_bwd_np);
}

}  // namespace gd

using namespace gd;

extern "C" int gd_attn_fwd_sm100(const void* const* q, const void* const* k, const void* const* v, void* const* o, void* const* lse,
                                 void* const* os, int G, int H, int N, int Nk, int d, float scale, const long* strides, int os_is_bf16,
                                 void* stream) {
    GD_CHECK_ARG(q && k && v && o && lse && G > 0 && G <= SM100_MAXG && H > 0);
    if (!(N == Nk && N % 128 == 0 && (d == 40 || d == 80)))
        return set_error(GD_ERR_UNSUPPORTED, "gd_attn_fwd_sm100 serves N == Nk, N %% 128 == 0, d in {40, 80}; got N=%d Nk=%d d=%d", N, Nk, d);
    const long q_rs = strides ? strides[0] : d, q_hs = strides ? strides[1] : (long)N * d;
    const long kv_rs = strides ? strides[2] : d, kv_hs = strides ? strides[3] : (long)N * d;
    Sm100Maps maps;
    Sm100Params p;
    const int bnk = g_bnk ? g_bnk : (d == 40 ? 128 : 64);
    for (int g = 0; g < G; ++g) {
        GD_CHECK_ARG(q[g] && k[g] && v[g] && lse[g] && (o[g] || (os && os[g])));
        int rc;
        if ((rc = make_map(&maps.q[g], q[g], N, H, d, q_rs, q_hs, BM)) != GD_OK) return rc;
        if ((rc = make_map(&maps.k[g], k[g], N, H, d, kv_rs, kv_hs, bnk)) != GD_OK) return rc;
        if ((rc = make_map(&maps.v[g], v[g], N, H, d, kv_rs, kv_hs, bnk)) != GD_OK) return rc;
        p.o[g] = (float*)o[g];
        p.lse[g] = (float*)lse[g];
        p.os[g] = os ? os[g] : nullptr;
    }
    p.os_rs = strides ? strides[4] : d; p.os_hs = strides ? strides[5] : (long)N * d; p.os_bf16 = os_is_bf16;
    if (os && ((p.os_rs % 8) != 0 || (p.os_hs % 8) != 0)) return set_error(GD_ERR_INVALID, "gd_attn_fwd_sm100: output strides must be multiples of 8");
    p.H = H; p.N = N; p.d = d; p.scale2 = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    if (bnk == 64) return d == 40 ? dispatch_sm100<40, 64>(maps, p, G, st) : dispatch_sm100<80, 64>(maps, p, G, st);
    return d == 40 ? dispatch_sm100<40, 128>(maps, p, G, st) : dispatch_sm100<80, 128>(maps, p, G, st);
}

// Tuning knobs of the tcgen05 kernels (process-wide; not part of the reference surface).
//   key 0  fwd: `value` in 0..4 of every 8 score pairs on the FMA-pipe polynomial (default 2)
//   key 2  fwd: keys per step: 0 (default: 128 at head_dim 40, 64 at head_dim 80), or 64 / 128 forced
//   key 3  bwd: value in 0..4 of every 8 score pairs on the polynomial (default 1)
extern "C" int gd_attn_sm100_config(int key, int value) {
    switch (key) {
        case 0: if (value < 0 || value > 4) break; g_np = value; return GD_OK;
        case 2: if (value != 0 && value != 64 && value != 128) break; g_bnk = value; return GD_OK;
        case 3: if (value < 0 || value > 4) break; g_bwd_np = value; return GD_OK;
    }
    return set_error(GD_ERR_INVALID, "gd_attn_sm100_config(key=%d, value=%d): see include/geodiffuser_b200.h", key, value);
}

// dQ of softmax(scale q k^T) v for the self-attention levels (N == Nk, N % 128 == 0, d in {40, 80}); same operands as gd_attn_bwd mode 0.
extern "C" int gd_attn_bwd_sm100(const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                                 const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* dq, int H, int N,
                                 int d, float scale, const long* strides, int dq_is_bf16, int extra_key_major, void* stream) {
    GD_CHECK_ARG(q && k && v && d_o && lse && delta && dq && H > 0);
    GD_CHECK_ARG((extra == nullptr) == (rowmap == nullptr));
    if (!(N % 128 == 0 && (d == 40 || d == 80)))
        return set_error(GD_ERR_UNSUPPORTED, "gd_attn_bwd_sm100 serves N %% 128 == 0, d in {40, 80}; got N=%d d=%d", N, d);
    if (extra && !extra_key_major && (ex_ld % 4 != 0 || ex_ld < N))
        return set_error(GD_ERR_INVALID, "gd_attn_bwd_sm100: extra row stride %d must be >= N and a multiple of 4", ex_ld);
    if (extra && extra_key_major && ex_ld < M) return set_error(GD_ERR_INVALID, "gd_attn_bwd_sm100: key-major extra needs ex_ld >= M (%d < %d)", ex_ld, M);
    const long q_rs = strides ? strides[0] : d, q_hs = strides ? strides[1] : (long)N * d;
    const long kv_rs = strides ? strides[2] : d, kv_hs = strides ? strides[3] : (long)N * d;
    Sm100BwdMaps maps;
    int rc;
    if ((rc = make_map(&maps.q, q, N, H, d, q_rs, q_hs, BM)) != GD_OK) return rc;
    if ((rc = make_map(&maps.d_o, d_o, N, H, d, d, (long)N * d, BM)) != GD_OK) return rc;
    if ((rc = make_map(&maps.k, k, N, H, d, kv_rs, kv_hs, 64)) != GD_OK) return rc;      // 64-key steps = TMA box rows of K / V
    if ((rc = make_map(&maps.v, v, N, H, d, kv_rs, kv_hs, 64)) != GD_OK) return rc;
    Sm100BwdParams p;
    p.lse = lse; p.delta = delta; p.extra = extra; p.extra_scale = extra_scale; p.rowmap = rowmap; p.ex_ld = ex_ld; p.M = M; p.ex_t = extra_key_major; p.dq = dq;
    p.dq_rs = strides ? strides[4] : d; p.dq_hs = strides ? strides[5] : (long)N * d; p.dq_bf16 = dq_is_bf16;
    if ((p.dq_rs % 8) != 0 || (p.dq_hs % 8) != 0) return set_error(GD_ERR_INVALID, "gd_attn_bwd_sm100: dq strides must be multiples of 8");
    p.H = H; p.N = N; p.scale = scale; p.scale2 = scale * 1.4426950408889634f;
    cudaStream_t st = (cudaStream_t)stream;
    return d == 40 ? dispatch_bwd64_sm100<40>(maps, p, st) : dispatch_bwd64_sm100<80>(maps, p, st);
}
