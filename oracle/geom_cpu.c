/*
 * oracle/geom_cpu.c -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * CPU restatement, with the IEEE operation order made explicit, of the float32 pipeline that
 * precedes the integer splat index in GeoDiffuser:
 *   pixel2cam                 GeoDiffuser/utils/warp_utils.py:738-747
 *   object centroid           GeoDiffuser/utils/warp_utils.py:426-427
 *   cam2pixel_vanilla         GeoDiffuser/utils/warp_utils.py:599-643
 *   bilinear down-resize      GeoDiffuser/utils/generic_torch.py:156-207 (torchvision T.Resize,
 *                             antialias=False == F.interpolate(bilinear, align_corners=False))
 *
 * Every 3x3 @ 3xN product follows what torch's CPU sgemm does on this image (verified in
 * oracle/make_golden.py against the reference run here):  acc = a0*b0; acc = fma(a1,b1,acc);
 * acc = fma(a2,b2,acc).  Everything else is a single correctly-rounded IEEE op per Python op.
 *
 * The centroid is the one reduction on the path.  torch's fp32 `mean` has an unspecified
 * (vectorised, cascaded) summation order, so the canonical order defined HERE -- and followed
 * bit-for-bit by the CUDA kernel -- is: double-precision accumulation, left-to-right inside an
 * image row, then the row partials top-to-bottom; mean = (float)(sum / count).
 *
 * Build: gcc -O2 -ffp-contract=off (fmaf() calls are the only fused operations).
 */
#include <math.h>
#include <stdint.h>

static inline float dot3_fma(const float* a, float b0, float b1, float b2) {
    float acc = a[0] * b0;
    acc = fmaf(a[1], b1, acc);
    acc = fmaf(a[2], b2, acc);
    return acc;
}

/* cam: (3, H, W) = (Kinv @ [u, v, 1]) * depth */
void geo_pixel2cam(const float* depth, int H, int W, const float* Kinv, float* cam) {
    const long hw = (long)H * W;
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            const long p = (long)v * W + u;
            const float d = depth[p];
            for (int r = 0; r < 3; ++r) cam[r * hw + p] = dot3_fma(Kinv + 3 * r, (float)u, (float)v, 1.0f) * d;
        }
}

/* canonical centroid of cam over mask >= 0.5; returns count, writes float mean[3] */
long geo_centroid(const float* cam, const float* mask, int H, int W, float* mean) {
    const long hw = (long)H * W;
    double tot[3] = {0.0, 0.0, 0.0};
    long cnt = 0;
    for (int v = 0; v < H; ++v) {
        double row[3] = {0.0, 0.0, 0.0};
        for (int u = 0; u < W; ++u) {
            const long p = (long)v * W + u;
            if (mask[p] >= 0.5f) {
                row[0] += (double)cam[p]; row[1] += (double)cam[hw + p]; row[2] += (double)cam[2 * hw + p];
                ++cnt;
            }
        }
        tot[0] += row[0]; tot[1] += row[1]; tot[2] += row[2];
    }
    for (int r = 0; r < 3; ++r) mean[r] = (float)(tot[r] / (double)cnt);
    return cnt;
}

/* coords: (H, W, 3) = (x_norm, y_norm, Z).  Rt: 3x4 row-major [R | t].  K: 3x3 */
void geo_project(const float* cam, int H, int W, const float* Rt, const float* K, float* coords) {
    const long hw = (long)H * W;
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    for (long p = 0; p < hw; ++p) {
        const float c0 = cam[p], c1 = cam[hw + p], c2 = cam[2 * hw + p];
        float q[3];
        for (int r = 0; r < 3; ++r) {
            const float rr[3] = {Rt[4 * r], Rt[4 * r + 1], Rt[4 * r + 2]};
            q[r] = dot3_fma(rr, c0, c1, c2) + Rt[4 * r + 3];
        }
        const float X = dot3_fma(K, q[0], q[1], q[2]);
        const float Y = dot3_fma(K + 3, q[0], q[1], q[2]);
        float Z = dot3_fma(K + 6, q[0], q[1], q[2]);
        if (Z < 1e-3f) Z = 1e-3f;                           /* .clamp(min=1e-3); NaN propagates */
        coords[3 * p + 0] = (2.0f * (X / Z)) / wm1 - 1.0f;
        coords[3 * p + 1] = (2.0f * (Y / Z)) / hm1 - 1.0f;
        coords[3 * p + 2] = Z;
    }
}

/* planes: (C, Hin, Win) -> (C, Hout, Wout); torch upsample_bilinear2d, align_corners=False */
void geo_resize_bilinear(const float* src, int C, int Hin, int Win, int Hout, int Wout, float* dst) {
    const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
    for (int c = 0; c < C; ++c)
        for (int oy = 0; oy < Hout; ++oy) {
            float fy = sh * ((float)oy + 0.5f) - 0.5f; if (fy < 0.0f) fy = 0.0f;
            int y0 = (int)fy; int y1 = y0 + ((y0 < Hin - 1) ? 1 : 0);
            const float ly1 = fy - (float)y0, ly0 = 1.0f - ly1;
            for (int ox = 0; ox < Wout; ++ox) {
                float fx = sw * ((float)ox + 0.5f) - 0.5f; if (fx < 0.0f) fx = 0.0f;
                int x0 = (int)fx; int x1 = x0 + ((x0 < Win - 1) ? 1 : 0);
                const float lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
                const float* s = src + (long)c * Hin * Win;
                /* torch's CPU kernel (verified here for both NCHW and channels-last inputs):
                 * ((w00*p00 + w01*p01) + w10*p10) + w11*p11 with w_ab = ly_a * lx_b; exact for the
                 * power-of-two ratios the path uses (all four weights are 0.25). */
                const float w00 = ly0 * lx0, w01 = ly0 * lx1, w10 = ly1 * lx0, w11 = ly1 * lx1;
                float acc = w00 * s[(long)y0 * Win + x0] + w01 * s[(long)y0 * Win + x1];
                acc = acc + w10 * s[(long)y1 * Win + x0];
                acc = acc + w11 * s[(long)y1 * Win + x1];
                dst[((long)c * Hout + oy) * Wout + ox] = acc;
            }
        }
}
