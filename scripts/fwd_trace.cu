// scripts/fwd_trace.cu -- timeline of ONE CTA of the tcgen05 forward (clock64 stamps of its softmax warps, MMA issuer and TMA producer)
// while the full grid runs: where a 128-query x BNK-key step spends its cycles.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DGD_TRACE -o scripts/_bin/fwd_trace scripts/fwd_trace.cu -lcuda
//   scripts/_bin/fwd_trace [bnk=64|128] [np=2] [G=3] [trace_x=5]
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../geodiffuser_b200/csrc/attention_sm100.cu"

namespace gd {
thread_local char g_last_error[256];
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    fprintf(stderr, "error %d: %s\n", code, g_last_error);
    return code;
}
}  // namespace gd

template <int BNK, int NP> static int run(int G, int H, int N, int trace_x) {
    constexpr int D = 40;
    const size_t n = (size_t)H * N * D;
    std::vector<__nv_bfloat16> hq(n);
    srand(1234);
    auto fill = [&](void* dptr) {
        for (size_t i = 0; i < n; ++i) {
            float u = 0.f;
            for (int t = 0; t < 4; ++t) u += (float)rand() / RAND_MAX - 0.5f;
            hq[i] = __float2bfloat16(u * 1.5f * 1.7f);
        }
        cudaMemcpy(dptr, hq.data(), n * 2, cudaMemcpyHostToDevice);
    };
    void *q[8], *k, *v, *o[8], *lse[8];
    cudaMalloc(&k, n * 2); cudaMalloc(&v, n * 2);
    fill(k); fill(v);
    Sm100Maps maps;
    Sm100Params p;
    for (int g = 0; g < G; ++g) {
        cudaMalloc(&q[g], n * 2); fill(q[g]);
        cudaMalloc(&o[g], n * 4); cudaMalloc(&lse[g], (size_t)H * N * 4);
        if (make_map(&maps.q[g], q[g], N, H, D, D, (long)N * D, BM)) return 1;
        if (make_map(&maps.k[g], k, N, H, D, D, (long)N * D, BNK)) return 1;
        if (make_map(&maps.v[g], v, N, H, D, D, (long)N * D, BNK)) return 1;
        p.o[g] = (float*)o[g]; p.lse[g] = (float*)lse[g]; p.os[g] = nullptr;
    }
    p.os_rs = D; p.os_hs = (long)N * D; p.os_bf16 = 0; p.H = H; p.N = N; p.d = D; p.scale2 = 1.4426950408889634f / sqrtf((float)D);
    const int NT = N / BNK;
    long long* trace;
    const size_t tn = 6 * 256 * 8;
    cudaMalloc(&trace, tn * 8);
    cudaMemset(trace, 0, tn * 8);
    p.trace = nullptr; p.trace_x = trace_x;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) if (launch_sm100<D, BNK, NP>(maps, p, G, 0)) return 1;
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) launch_sm100<D, BNK, NP>(maps, p, G, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("BNK=%d NP=%d G=%d H=%d N=%d: %.1f us per launch (%.1f TFLOP/s)\n", BNK, NP, G, H, N, ms / 20 * 1e3, 4.0 * G * H * N * (double)N * D / (ms / 20) / 1e9);
    p.trace = trace;
    launch_sm100<D, BNK, NP>(maps, p, G, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> t(tn);
    cudaMemcpy(t.data(), trace, tn * 8, cudaMemcpyDeviceToHost);
    auto T = [&](int role, int j, int slot) { return t[((size_t)role * 256 + j) * 8 + slot]; };
    const int j0 = 4, j1 = NT - 2;
    double sm[4][8] = {}, mm[8] = {};
    for (int j = j0; j < j1; ++j) {
        for (int w = 0; w < 4; ++w) {
            sm[w][0] += T(w, j, 1) - T(w, j, 0);     // s_full wait
            sm[w][1] += T(w, j, 2) - T(w, j, 1);     // tcgen05.ld
            sm[w][2] += T(w, j, 3) - T(w, j, 2);     // row max + first chunk
            sm[w][3] += T(w, j, 4) - T(w, j, 3);     // pv_done wait
            sm[w][4] += T(w, j, 5) - T(w, j, 4);     // remaining chunks
            sm[w][5] += T(w, j, 6) - T(w, j, 5);     // wait::st + arrive
            sm[w][6] += T(w, j + 1, 0) - T(w, j, 0); // whole step
            sm[w][7] += T(w, j, 1) - T(4, j, 2);     // QK(j) issue -> softmax has S(j)
        }
        mm[0] += T(4, j + 1, 1) - T(4, j + 1, 0);    // k_full wait (for QK(j+1))
        mm[1] += T(4, j + 1, 2) - T(4, j + 1, 1);    // s_free wait
        mm[2] += T(4, j, 3) - T(4, j + 1, 2);        // QK issue
        mm[3] += T(4, j, 4) - T(4, j, 3);            // v_full wait
        mm[4] += T(4, j, 5) - T(4, j, 4);            // p_full wait
        mm[5] += T(4, j, 6) - T(4, j, 5);            // PV issue
        mm[6] += T(4, j, 5) - T(0, j, 6);            // softmax warp 0 arrive p_full -> MMA thread past the wait
        mm[7] += T(0, j + 1, 4) - T(4, j, 6);        // PV(j) issued -> softmax warp 0 past pv_done(j) wait (incl. its own first chunk)
    }
    const double nn = j1 - j0;
    printf("softmax warps, cycles per step (mean over steps %d..%d):\n  warp  s_full-wait  tcgen05.ld  max+chunk0  pv_done-wait  chunks1..  st+arrive   step   QK-issue->S-seen\n", j0, j1 - 1);
    for (int w = 0; w < 4; ++w)
        printf("  %d     %8.0f   %8.0f    %8.0f    %8.0f   %8.0f   %8.0f  %8.0f  %8.0f\n", w, sm[w][0] / nn, sm[w][1] / nn, sm[w][2] / nn, sm[w][3] / nn, sm[w][4] / nn,
               sm[w][5] / nn, sm[w][6] / nn, sm[w][7] / nn);
    printf("MMA issuer: k_full-wait %.0f  s_free-wait %.0f  QK-issue %.0f  v_full-wait %.0f  p_full-wait %.0f  PV-issue %.0f | p_full arrive->seen %.0f  PV issued->pv_done seen by warp 0 %.0f\n",
           mm[0] / nn, mm[1] / nn, mm[2] / nn, mm[3] / nn, mm[4] / nn, mm[5] / nn, mm[6] / nn, mm[7] / nn);
    const long long b = T(0, 8, 0);
    printf("raw timeline of steps 8..10 (cycles relative to warp 0's step-8 start):\n");
    for (int j = 8; j < 11; ++j) {
        for (int w = 0; w < 4; w += 3) {
            printf("  step %2d softmax warp %d:", j, w);
            for (int s = 0; s < 7; ++s) printf(" %7lld", T(w, j, s) - b);
            printf("\n");
        }
        printf("  step %2d MMA (QK(j+1): top, k_full, s_free | PV(j): top, v_full, p_full, issued):", j);
        for (int s = 0; s < 3; ++s) printf(" %7lld", T(4, j + 1, s) - b);
        printf(" |");
        for (int s = 3; s < 7; ++s) printf(" %7lld", T(4, j, s) - b);
        printf("\n  step %2d TMA (k_empty, v_empty passed): %7lld %7lld\n", j, T(5, j, 0) - b, T(5, j, 1) - b);
    }
    return 0;
}

int main(int argc, char** argv) {
    const int bnk = argc > 1 ? atoi(argv[1]) : 64, np = argc > 2 ? atoi(argv[2]) : 2, G = argc > 3 ? atoi(argv[3]) : 3, tx = argc > 4 ? atoi(argv[4]) : 5;
    if (np != 2) { fprintf(stderr, "np=2 only\n"); return 1; }
    return bnk == 64 ? run<64, 2>(G, 8, 4096, tx) : run<128, 2>(G, 8, 4096, tx);
}
