/*
 * oracle/pt3d_cpu.c -- TEST INFRASTRUCTURE ONLY (never linked into the product path).
 *
 * CPU restatement of the two pytorch3d operators GeoDiffuser's splat warp calls
 * (reference call sites: GeoDiffuser/utils/warp_utils.py:111-113 `rasterize_points`,
 * :156-160 `compositing.alpha_composite`).  pytorch3d itself is NOT vendored under
 * /root/reference (pinned at git 89653419d0973396f3eff1a381ba09a07fffc2ed in
 * GeoDiffuser/envs/requirements.txt:103) and is not installable here, so this file
 * restates the published algorithm of its *naive CPU* rasteriser:
 *
 *   - output pixel (yi, xi) has NDC centre
 *        xf = -1 + (2*(W-1-xi) + 1) / W ,  yf = -1 + (2*(H-1-yi) + 1) / H
 *     (PixToNonSquareNdc: `-offset + (range*i + offset)/S`, +X left / +Y up);
 *   - a point contributes iff  z >= 0  and  dx*dx + dy*dy < r*r  (strict, fp32, no FMA);
 *   - per pixel keep the K smallest in (z, packed-index) lexicographic order
 *     (the CPU version pops a max-heap of (z, idx, dist2) tuples), written ascending;
 *   - empty slots: idx = -1, zbuf = -1, dist2 = -1;
 *   - idx indexes the PACKED point list (b*P + p).
 *
 *   alpha_composite: out[b,c,y,x] = sum_k cum * a_k * feat[c, idx_k], cum *= (1-a_k),
 *   slots with idx < 0 skipped.
 *
 * PARITY UNPINNED at this boundary: the reference holds no golden vector for it and
 * pytorch3d cannot be executed here; ties on z (every 2-D edit uses constant depth) are
 * resolved by ascending packed index, which is the naive/CPU path's behaviour.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.sh).
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>

static inline float pix_to_ndc(int i, int S) {
    const float range = 2.0f;
    const float offset = range / 2.0f;
    return -offset + (range * (float)i + offset) / (float)S;
}

/* pts: (B, P, 3) fp32, ALREADY in pytorch3d convention (x,y negated by the caller,
 * warp_utils.py:90-91).  Outputs (B, S, S, K). */
void pt3d_rasterize_points(const float* pts, int B, int P, int S, float radius, int K,
                           int32_t* idx, float* zbuf, float* dist2) {
    const float r2 = radius * radius;
    const long npix = (long)B * S * S;
    for (long i = 0; i < npix * K; ++i) { idx[i] = -1; zbuf[i] = -1.0f; dist2[i] = -1.0f; }
    int* cnt = (int*)calloc((size_t)npix, sizeof(int));
    float* cx = (float*)malloc(sizeof(float) * (size_t)S);
    for (int i = 0; i < S; ++i) cx[i] = pix_to_ndc(S - 1 - i, S);   /* centre of column/row i */

    /* conservative pixel footprint: centre spacing is 2/S */
    const float pad = radius * (float)S * 0.5f + 2.0f;
    for (int b = 0; b < B; ++b) {
        for (int p = 0; p < P; ++p) {
            const float* q = pts + ((long)b * P + p) * 3;
            const float px = q[0], py = q[1], pz = q[2];
            if (!(pz >= 0.0f)) continue;          /* `if (pz < 0) continue;` -- NaN never passes dist test either */
            if (!(px == px) || !(py == py)) continue;
            /* column whose centre is nearest: cx[i] = -1 + (2*(S-1-i)+1)/S  =>  i ~ S-1 - ((px+1)*S-1)/2 */
            float fi = (float)(S - 1) - ((px + 1.0f) * (float)S - 1.0f) * 0.5f;
            float fj = (float)(S - 1) - ((py + 1.0f) * (float)S - 1.0f) * 0.5f;
            if (fi < -pad - 1 || fi > S + pad || fj < -pad - 1 || fj > S + pad) continue;
            int x0 = (int)floorf(fi - pad), x1 = (int)ceilf(fi + pad);
            int y0 = (int)floorf(fj - pad), y1 = (int)ceilf(fj + pad);
            if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0;
            if (x1 > S - 1) x1 = S - 1; if (y1 > S - 1) y1 = S - 1;
            const int32_t gp = (int32_t)((long)b * P + p);
            for (int yi = y0; yi <= y1; ++yi) {
                const float dy = py - cx[yi];
                const float dy2 = dy * dy;
                for (int xi = x0; xi <= x1; ++xi) {
                    const float dx = px - cx[xi];
                    const float d2 = dx * dx + dy2;
                    if (!(d2 < r2)) continue;
                    const long pix = ((long)b * S + yi) * S + xi;
                    int32_t* pi = idx + pix * K; float* pz_ = zbuf + pix * K; float* pd = dist2 + pix * K;
                    int n = cnt[pix];
                    /* points arrive in ascending packed index, so ties on z keep arrival order */
                    int pos = n;
                    while (pos > 0 && pz_[pos - 1] > pz) --pos;
                    if (pos >= K) continue;
                    int last = (n < K) ? n : K - 1;
                    for (int t = last; t > pos; --t) { pi[t] = pi[t - 1]; pz_[t] = pz_[t - 1]; pd[t] = pd[t - 1]; }
                    pi[pos] = gp; pz_[pos] = pz; pd[pos] = d2;
                    if (n < K) cnt[pix] = n + 1;
                }
            }
        }
    }
    free(cnt); free(cx);
}

/* idx: (B, K, S, S) int64-free int32, alpha: (B, K, S, S), feat: (C, B*P) packed.  out: (B, C, S, S) */
void pt3d_alpha_composite(const int32_t* idx, const float* alpha, const float* feat,
                          int B, int K, int S, int C, long Ptot, float* out) {
    const long hw = (long)S * S;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (long px = 0; px < hw; ++px) {
                float cum = 1.0f, acc = 0.0f;
                for (int k = 0; k < K; ++k) {
                    const int32_t n = idx[((long)b * K + k) * hw + px];
                    if (n < 0) continue;
                    const float a = alpha[((long)b * K + k) * hw + px];
                    acc += cum * a * feat[(long)c * Ptot + n];
                    cum = cum * (1.0f - a);
                }
                out[((long)b * C + c) * hw + px] = acc;
            }
}

/* Mesh coverage (restates the *observable* result of warp_utils.py:235-298 `splatter_mesh`
 * on an all-ones vertex texture: 1 where any front-facing-or-not triangle with all z > 0
 * covers the pixel centre, else 0).  verts: (V,3) in pytorch3d convention; faces: (F,3). */
static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}
static inline float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float vx = bx - ax, vy = by - ay;
    const float l2 = vx * vx + vy * vy;
    if (l2 <= 1e-8f) { const float dx = px - bx, dy = py - by; return dx * dx + dy * dy; }
    float t = (vx * (px - ax) + vy * (py - ay)) / l2;
    if (t < 0.0f) t = 0.0f; if (t > 1.0f) t = 1.0f;
    const float qx = ax + t * vx, qy = ay + t * vy;
    const float dx = px - qx, dy = py - qy;
    return dx * dx + dy * dy;
}
void pt3d_mesh_coverage(const float* verts, long V, const int32_t* faces, long F, int S, float blur, float* out) {
    (void)V;
    for (long i = 0; i < (long)S * S; ++i) out[i] = 0.0f;
    float* cx = (float*)malloc(sizeof(float) * (size_t)S);
    for (int i = 0; i < S; ++i) cx[i] = pix_to_ndc(S - 1 - i, S);
    for (long f = 0; f < F; ++f) {
        const float* a = verts + 3L * faces[3 * f + 0];
        const float* b = verts + 3L * faces[3 * f + 1];
        const float* c = verts + 3L * faces[3 * f + 2];
        if (a[2] < 1e-8f || b[2] < 1e-8f || c[2] < 1e-8f) continue;        /* behind-camera cull */
        float xmin = fminf(a[0], fminf(b[0], c[0])), xmax = fmaxf(a[0], fmaxf(b[0], c[0]));
        float ymin = fminf(a[1], fminf(b[1], c[1])), ymax = fmaxf(a[1], fmaxf(b[1], c[1]));
        if (!(xmin == xmin) || !(ymin == ymin)) continue;
        /* columns: cx decreasing in i.  i = S-1 - ((x+1)*S-1)/2 */
        float fi0 = (float)(S - 1) - ((xmax + 1.0f) * (float)S - 1.0f) * 0.5f;
        float fi1 = (float)(S - 1) - ((xmin + 1.0f) * (float)S - 1.0f) * 0.5f;
        float fj0 = (float)(S - 1) - ((ymax + 1.0f) * (float)S - 1.0f) * 0.5f;
        float fj1 = (float)(S - 1) - ((ymin + 1.0f) * (float)S - 1.0f) * 0.5f;
        if (fi1 < -2 || fi0 > S + 1 || fj1 < -2 || fj0 > S + 1) continue;
        int x0 = (int)floorf(fi0) - 1, x1 = (int)ceilf(fi1) + 1, y0 = (int)floorf(fj0) - 1, y1 = (int)ceilf(fj1) + 1;
        if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (x1 > S - 1) x1 = S - 1; if (y1 > S - 1) y1 = S - 1;
        const float area = edge_fn(c[0], c[1], a[0], a[1], b[0], b[1]);
        if (fabsf(area) <= 1e-8f) continue;                                  /* kEpsilon zero-area cull */
        for (int yi = y0; yi <= y1; ++yi)
            for (int xi = x0; xi <= x1; ++xi) {
                const float px = cx[xi], py = cx[yi];
                const float w0 = edge_fn(px, py, b[0], b[1], c[0], c[1]) / area;
                const float w1 = edge_fn(px, py, c[0], c[1], a[0], a[1]) / area;
                const float w2 = edge_fn(px, py, a[0], a[1], b[0], b[1]) / area;
                int hit = (w0 > 0.0f && w1 > 0.0f && w2 > 0.0f);
                if (!hit) {                       /* `if (!inside && dist >= blur_radius) return;` */
                    float d = fminf(seg_dist2(px, py, a[0], a[1], b[0], b[1]),
                              fminf(seg_dist2(px, py, b[0], b[1], c[0], c[1]), seg_dist2(px, py, c[0], c[1], a[0], a[1])));
                    hit = d < blur;
                }
                if (hit) out[(long)yi * S + xi] = 1.0f;
            }
    }
    free(cx);
}
