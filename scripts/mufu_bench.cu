// micro-benchmarks behind the softmax design of csrc/attention_sm100.cu: per-SM throughput of MUFU.EX2 as a function of resident warps,
// of the FMA-pipe polynomial, of tcgen05.ld (TMEM -> registers), and of the two running side by side.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly(float x) {
    x = fmaxf(x, -126.f); float t = x + 12582912.f; float f = x - (t - 12582912.f);
    float p = fmaf(f, 0.0551716648f, 0.2426111251f); p = fmaf(p, f, 0.6932609677f); p = fmaf(p, f, 0.9999280572f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// MODE 0: MUFU only, 1: poly only, 2: softmax-like mix per element (FFMA + MUFU + FADD, cvt pack every 2)
template <int MODE> __global__ void k_math(float* out, int iters, float seed, long long* clk) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.f;
    float acc = 0.f; unsigned pk = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]) - 1.5f;
            if (MODE == 1) a[i] = poly(a[i]) - 1.5f;
            if (MODE == 2) { float p = ex2(fmaf(a[i], seed, -1.0f)); acc += p; a[i] = p - 1.5f; }
        }
    }
    long long t1 = clock64();
    float s = acc + __uint_as_float(pk); for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
// warps [0, nld) stream tcgen05.ld.x32 over their TMEM lane quarter; warps [nld, nld+nmath) run the MUFU loop
__global__ void k_tmem(float* out, int iters, int nld, int nmath, long long* clk) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    float s = 0.f;
    long long t0 = clock64();
    if (warp < nld) {
        uint32_t r[32]; uint32_t acc = 0;
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(base + ((it * 4 + c) & 15) * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0] ^ r[31];
            }
        }
        s = __uint_as_float(acc);
    } else if (warp < nld + nmath) {
        float a[16];
        for (int i = 0; i < 16; ++i) a[i] = (threadIdx.x + i) * 1e-3f - 1.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]) - 1.5f;
        }
        for (int i = 0; i < 16; ++i) s += a[i];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) clk[warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    long long* clk; cudaMallocManaged(&clk, 64 * 8);
    const int iters = 2048;
    const char* names[3] = {"MUFU.EX2 only", "poly3 (FMA/ALU pipes)", "FFMA+MUFU+FADD"};
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k_math<0><<<148, threads>>>(out, iters, 1.f, clk);
                if (mode == 1) k_math<1><<<148, threads>>>(out, iters, 1.f, clk);
                if (mode == 2) k_math<2><<<148, threads>>>(out, iters, 1.f, clk);
                cudaDeviceSynchronize();
            }
            printf("%-24s warps/SM %2d: %.2f elem/clk/SM\n", names[mode], threads / 32, (double)threads * 16 * iters / (double)clk[0]);
        }
    for (int nld : {1, 4, 8, 16})
        for (int nmath : {0, 4, 8}) {
            const int threads = 32 * (nld + nmath);
            for (int rep = 0; rep < 2; ++rep) { k_tmem<<<148, threads>>>(out, iters, nld, nmath, clk); cudaDeviceSynchronize(); }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            double ld_bw = (double)nld * iters * 4 * 4096 / (double)clk[0];
            printf("tcgen05.ld.x32 warps %2d + MUFU warps %d: LDTM %.1f B/clk/SM (ld warp0 %lld clk)", nld, nmath, ld_bw, clk[0]);
            if (nmath) printf(", MUFU %.2f elem/clk/SM (math warp %lld clk)", (double)nmath * 32 * 16 * iters / (double)clk[nld], clk[nld]);
            printf("\n");
        }
    return 0;
}
