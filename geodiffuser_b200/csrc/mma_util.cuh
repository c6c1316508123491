// geodiffuser_b200/csrc/mma_util.cuh -- warp-level mma.sync / ldmatrix / cp.async helpers shared by the
// ragged-shape attention kernels (attention_mma.cu) and the NT-GEMM epilogue kernels (corr_gemm.cu).
#pragma once
#include "common.cuh"

namespace gd {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte async copy global -> shared; when !pred the destination is zero-filled
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// A fragment (16 rows x 16 k) of a row-major [rows][ld] bf16 tile at (row0, k0)
__device__ __forceinline__ void load_a_frag(uint32_t* a, const bf16* tile, int ld, int row0, int k0, int lane) {
    ldmatrix_x4(a, tile + (row0 + (lane & 15)) * ld + k0 + (lane >> 4) * 8);
}
// B fragment (16 k x 8 n) where B[k][n] = tile[n][k]  (tile row-major [n][ld]: "NT" operand)
__device__ __forceinline__ void load_b_frag_nt(uint32_t& b0, uint32_t& b1, const bf16* tile, int ld, int n0, int k0, int lane) {
    const bf16* p = tile + (n0 + (lane >> 2)) * ld + k0 + (lane & 3) * 2;
    b0 = *reinterpret_cast<const uint32_t*>(p);
    b1 = *reinterpret_cast<const uint32_t*>(p + 8);
}
// two B fragments (16 k x 16 n) where B[k][n] = tile[k][n]  (tile row-major [k][ld]): r[0..1] -> n0..n0+7, r[2..3] -> n0+8..
__device__ __forceinline__ void load_b_frag_nn_x2(uint32_t* r, const bf16* tile, int ld, int k0, int n0, int lane) {
    ldmatrix_x4_trans(r, tile + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + (lane >> 4) * 8);
}

// copy a [rows x d] bf16 tile (global row stride `gld` elements) into shared [64][ld]; rows >= valid_rows zero-filled
template <int THREADS>
__device__ __forceinline__ void load_tile_async(bf16* dst, int ld, const bf16* src, long gld, int d, int valid_rows, int tid) {
    const int cpr = d >> 3;  // 16B chunks per row
    for (int c = tid; c < 64 * cpr; c += THREADS) {
        const int r = c / cpr, k = (c % cpr) * 8;
        const bool ok = r < valid_rows;
        cp_async16(dst + r * ld + k, src + (long)(ok ? r : 0) * gld + k, ok);
    }
}

}  // namespace gd
