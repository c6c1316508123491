/*
 * geodiffuser_b200.h -- C ABI of libgeodiffuser_b200.so: the B200 (sm_100a) implementation of GeoDiffuser's geometry-warped
 * shared-attention hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in `_host`;
 * `stream` is a cudaStream_t; every function returns 0 on success or one of GD_ERR_* (message via gd_last_error()).
 * No function allocates: outputs and workspaces are passed in.  Launches are asynchronous on `stream`.
 *
 * Each entry point cites the reference code it replaces (paths relative to /root/reference/GeoDiffuser/utils/).
 * The reference is pure Python, so the "FFI" a maintainer binds is ctypes: see INTEGRATION.md.
 */
#ifndef GEODIFFUSER_B200_H
#define GEODIFFUSER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GD_OK 0
#define GD_ERR_INVALID 1
#define GD_ERR_CUDA 2
#define GD_ERR_UNSUPPORTED 3

const char* gd_last_error(void);
int gd_version(void);

/* ---- (1) correspondence field, masks, splat index ------------------------------------------------------------------ */

/* warp_utils.py:738-747 pixel2cam + :426-427 masked centroid.  depth, mask (H,W); Kinv9_host 3x3 row-major;
 * cam (3,H,W) out; centroid4 = {cx, cy, cz, count} out.  Canonical reduction order: double, row-major (see oracle/geom_cpu.c). */
int gd_corr_pixel2cam(const float* depth, const float* mask, int H, int W, const float* Kinv9_host, float* cam, float* centroid4,
                      void* stream);

/* warp_utils.py:599-643 cam2pixel_vanilla.  Rt12_host = top 3 rows of C^-1 T C; coords (H,W,3) = (x_norm, y_norm, Z) out. */
int gd_corr_project(const float* cam, int H, int W, const float* Rt12_host, const float* K9_host, float* coords, void* stream);

/* generic_torch.py:156-207 T.Resize(BILINEAR, antialias=False).  channels_last: src (Hin,Win,C) else (C,Hin,Win). */
int gd_resize_bilinear(const float* src, int C, int Hin, int Win, int channels_last, float* dst, int Hout, int Wout, void* stream);

/* attention_processors.py:338-360 mask algebra at resolution S from (Hin,Hin) planes; mask_new_warped / amodal_mask may be NULL
 * (zeros: the Remover, :856-866).  out6 (6,S,S): mask_new_warped, mask_warp, amodal_mask, mask_intersection, mask_1_empty, mask_wo_edit */
int gd_masks_build(const float* image_mask, const float* mask_new_warped, const float* amodal_mask, int Hin, int S, float* out6,
                   void* stream);

/* warp_utils.py:80-113 (pytorch3d rasterize_points).  coords (B,S,S,3); idx (B,S,S,K) int32 packed b*S*S+p, -1 = empty;
 * zbuf (may be NULL), dist2 (B,S,S,K).  Order: z ascending, ties by packed index.  K <= 16. */
int gd_splat_index(const float* coords, int B, int S, float radius_ndc, int K, int* idx, float* zbuf, float* dist2, void* stream);

/* warp_utils.py:131-176 (alpha weights + pytorch3d alpha_composite + .half()).  dtype: 0 fp32, 1 bf16.
 * layout: 0 (B,C,P) channel-first, 1 (B,P,C) channel-last.  idx/dist2 (Bi,P,K), Bi in {1,B}.  blend_mask (P) or NULL:
 * out = src*(1-m) + m*warped (attention_processors.py:544).  post: 0 none, 1 binarize(>0.5) (generic_torch.py:122). */
int gd_splat_composite(const void* src, int src_dtype, int layout, const int* idx, const float* dist2, int Bi, int B, int P, int C,
                       int K, float r2, float tau, const float* blend_mask, int post, void* out, int out_dtype, void* stream);

/* The per-layer query warp (attention_processors.py:363,424,544: the same index for all heads) on one (B = heads, P, C = head_dim) slab
 * with explicit element strides: element (b, p, c) at base + b*head + p*row + c; *_strides_host = {row, head} (NULL: contiguous (B,P,C)).
 * One warp per output pixel; same arithmetic as gd_splat_composite, bit for bit. */
int gd_splat_composite_rows(const void* src, int src_dtype, const long* src_strides_host, const int* idx, const float* dist2, int B, int P,
                            int C, int K, float r2, float tau, const float* blend_mask, int post, void* out, int out_dtype,
                            const long* out_strides_host, void* stream);

/* warp_utils.py:364-399 get_mesh + :235-298 splatter_mesh (pytorch3d rasterize_meshes): coverage of the object's depth mesh. */
int gd_mesh_mask(const float* coords, const float* mask, int H, int W, float blur, float* out, void* stream);

/* generic_torch.py:210-235 torch_erode (mode 0) / torch_dilate (mode 1), k x k window, zero padding.  src/dst (B,H,W). */
int gd_morph(const float* src, int B, int H, int W, int kernel, int mode, float* dst, void* stream);

/* ---- (2) shared attention forward ---------------------------------------------------------------------------------- */

/* attention_sharing.py:30-47 compute_attention + torch.bmm(P,V) (attention_processors.py:427-433, 548-557, 643-647).
 * G query streams q[g], each an (H,N,d) bf16 SLAB, against k[g], v[g] (H,Nk,d) bf16 slabs.  A slab is addressed as
 * base + h*head_stride + n*row_stride + c (element strides, multiples of 8): strides_host = {q_row, q_head, kv_row, kv_head, os_row, os_head};
 * NULL = contiguous (H,N,d), the reference's head_to_batch_dim layout (attention_sharing.py:210-242).  The projection layout (N, H*d) that
 * to_q / to_k / to_v produce is {H*d, d}: it is read in place, so neither head_to_batch_dim nor batch_to_head_dim copies anything.
 * Outputs per stream: o[g] (H,N,d) fp32 contiguous (feeds the loss terms) and / or os[g], a strided slab (bf16 if os_is_bf16 else fp32) --
 * at least one of the two; lse[g] (H,N) fp32 (natural log).  The pointer arrays are HOST arrays of G device pointers (os_host may be NULL);
 * G <= 8; d % 8 == 0, d <= 160.  Any N, Nk. */
int gd_attn_fwd_generic(const void* const* q_host, const void* const* k_host, const void* const* v_host, void* const* o_host,
                        void* const* lse_host, void* const* os_host, int G, int H, int N, int Nk, int d, float scale,
                        const long* strides_host, int os_is_bf16, void* stream);

/* Same contract, tcgen05 / TMEM / TMA kernel for the large self-attention levels: N == Nk, N % 128 == 0, d in {40, 80}
 * (operands are 3-D TMA tensors (d, N, H) with the strides above). */
int gd_attn_fwd_sm100(const void* const* q_host, const void* const* k_host, const void* const* v_host, void* const* o_host,
                      void* const* lse_host, void* const* os_host, int G, int H, int N, int Nk, int d, float scale,
                      const long* strides_host, int os_is_bf16, void* stream);

/* Tuning knobs of the tcgen05 kernels (process-wide, not part of the reference surface).  key 0: forward, `value` in 0..4 of every 8
 * score pairs of the online softmax evaluated by a degree-3 polynomial on the FMA pipe instead of the MUFU (packed fp32x2 arithmetic;
 * default 2); key 2: forward, keys per step: 0 (default: 128 at head_dim 40, 64 at head_dim 80), or 64 / 128 forced;
 * key 3: polynomial share of the backward kernel (0..4 of 8 pairs, default 1). */
int gd_attn_sm100_config(int key, int value);

/* ---- (3) backward, fused with the attention-map losses --------------------------------------------------------------- */

/* dO = g_out * coef[row] + g_loss * (*loss_scale) (bf16 out), delta[h,row] = sum_c dO*O (+ delta_extra[h, rowmap[row]] * *loss_scale).
 * g_out: (H,N,d) slab fp32/bf16, g_strides_host = {row, head} (NULL: contiguous), or NULL; coef (N) or NULL; g_loss (H,N,d) fp32 or NULL;
 * loss_scale: device scalar or NULL (=1). */
int gd_attn_bwd_prep(const void* g_out, int g_out_is_bf16, const long* g_strides_host, const float* coef, const float* g_loss,
                     const float* loss_scale, const float* o, const float* delta_extra, const int* rowmap, int M, int H, int N, int d,
                     void* d_o_bf16, float* delta, void* stream);

/* What torch autograd derives for softmax(scale q k^T) v (attention_sharing.py:35-45).  mode 0: out = dQ, a slab like q; mode 1: out = dK,
 * a slab like k.  q, k, v slabs with strides_host = {q_row, q_head, kv_row, kv_head, out_row, out_head} (NULL: contiguous); d_o (H,N,d) bf16
 * contiguous; out fp32 or bf16 (out_is_bf16).  extra (H,M,ex_ld) fp32 = dL/dP rows for the queries with rowmap[row] >= 0 (removal loss),
 * scaled by *extra_scale. */
int gd_attn_bwd(int mode, const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* out, int H, int N, int Nk,
                int d, float scale, const long* strides_host, int out_is_bf16, void* stream);

/* dK exactly as gd_attn_bwd mode 1, but with the query range split `splits` ways across the grid (cross layers: Nk = 77 gives only two
 * 64-key tiles per head, so one CTA per tile would walk all N queries serially).  workspace: >= splits * H * Nk * d floats; the partial
 * sums are added in ascending split order, so the result is deterministic.  splits <= 1 or workspace == NULL falls back to mode 1. */
int gd_attn_bwd_dk_split(const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                         const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* dk, float* workspace,
                         int splits, int H, int N, int Nk, int d, float scale, const long* strides_host, int out_is_bf16, void* stream);

/* Same operands and result as gd_attn_bwd mode 0 (dQ), tcgen05 / TMEM / TMA kernel for the self-attention levels:
 * N == Nk, N % 128 == 0, d in {40, 80}.  extra_key_major = 0: extra (H, M, ex_ld) as for gd_attn_bwd (ex_ld % 4 == 0);
 * extra_key_major = 1: extra (H, N, ex_ld) with ex_ld = M rounded up to 4 (gd_removal_extra_rows key_major = 1). */
int gd_attn_bwd_sm100(const void* q, const void* k, const void* v, const void* d_o, const float* lse, const float* delta,
                      const float* extra, const float* extra_scale, const int* rowmap, int ex_ld, int M, void* dq, int H, int N,
                      int d, float scale, const long* strides_host, int dq_is_bf16, int extra_key_major, void* stream);

/* fp32 -> bf16 */
int gd_cast_f32_to_bf16(const float* src, void* dst, long n, void* stream);

/* attention_processors.py:250-252: the base attention map A_b = softmax(scale q_b k_b^T) and the inpaint rows of the edit map, materialised
 * in bf16 from the stored lse: P[h, m, :] for query rows[m] (or m if rows == NULL); row stride ldp (multiple of 8, >= Nk, pad columns zero).
 * q, k slabs with qk_strides_host = {q_row, q_head, k_row, k_head} (NULL: contiguous). */
int gd_attn_probs(const void* q, const void* k, const float* lse, const int* rows, int M, int H, int N, int Nk, int d, float scale,
                  void* p_out, int ldp, const long* qk_strides_host, void* stream);

/* attention_processors.py:250-258: corr = A_e[rows] A_b^T, masked max / arg-max per 64-wide tile of columns.
 * partial (H, ceil(Nb/64), M, 4) = {max_in, argmax_in (int bits), max_bg, argmax_bg (int bits)}. */
int gd_corr_max_partial(const void* a_e_bf16, const void* a_b_bf16, int H, int M, int Nb, int Nk, int ld, const float* mask_in,
                        const float* mask_bg, float* partial, void* stream);

/* The same correlation + masked arg-max for the self-attention levels (N % 128 == 0, d in {40, 80}) as one tcgen05 / TMEM / TMA kernel that
 * recomputes the base map A_b = softmax(scale q_b k_b^T) tile by tile from q_b, k_b (slabs, qk_strides_host as gd_attn_probs) and lse_b (H,N)
 * instead of reading a materialised copy.  a_e (H, M, ld) bf16 from gd_attn_probs; partial (H, N/32, M, 4), one entry per block of 32 base rows. */
int gd_removal_corr_sm100(const void* q_b, const void* k_b, const float* lse_b, const void* a_e_bf16, int H, int M, int N, int d, float scale,
                          int ld, const long* qk_strides_host, const float* mask_in, const float* mask_bg, float* partial, void* stream);

/* attention_processors.py:256-266, backward side: the two base-map rows per (head, inpaint row) at the arg-max positions j2 (H*M, 2) written by
 * gd_removal_finalize, recomputed as softmax rows: p2 (H, 2M, ldp) bf16, rows [0,M) = j_bg, [M,2M) = j_in. */
int gd_attn_probs_rows2(const void* q, const void* k, const float* lse, const int* j2, int M, int H, int N, int Nk, int d, float scale,
                        void* p2_out, int ldp, const long* qk_strides_host, void* stream);

/* extra[h, m, :] = g2[h*M+m].x * p2[h, m, :] + g2[h*M+m].y * p2[h, M+m, :]: dL/dA_e rows for gd_attn_bwd*.  key_major = 0: (H, M, ld) fp32;
 * key_major = 1: transposed, (H, Nk, Mp) fp32 with Mp = M rounded up to 4 -- the layout gd_attn_bwd_sm100 reads coalesced. */
int gd_removal_extra_rows(const void* p2_bf16, const float* g2, int H, int M, int Nk, int ld, float* extra, int key_major, void* stream);

/* The removal term of dQ as its own contraction (self-attention levels): dS gets P o extra on the M inpaint rows only, and A_e[rows] is already
 * materialised, so  dq[h, rows[m], :] += (*gscale) * scale * sum_k W[h,m,k] K[h,k,:]  with  W = A_e[rows] o (g_bg P2[m] + g_in P2[M+m])  (bf16 (H,M,ld)).
 * Replaces the dense `dcorr . A_b` autograd forms at attention_processors.py:262-280; gd_attn_bwd_sm100 then runs without `extra`.
 * strides_host = {kv_row, kv_head, dq_row, dq_head} (NULL: contiguous); d in {40, 80}, Nk % 256 == 0; dq (fp32 / bf16) is accumulated in place. */
int gd_removal_weighted_rows(const void* a_e_bf16, const void* p2_bf16, const float* g2, int H, int M, int Nk, int ld, void* w_bf16, void* stream);
int gd_removal_dq_rows(const void* w_bf16, const void* k, const int* rows, const float* gscale, void* dq, int H, int M, int N, int Nk, int d,
                       float scale, int ld, const long* strides_host, int dq_is_bf16, void* stream);

/* attention_processors.py:231-246 (sim), 283-287 (movement), 289-305 (amodal, target t), loss.py:22-41 (smoothness): unweighted partial
 * sums (n_partials,5) and grad = d(weighted loss)/d replace_out.  c_* = weight / denominator of each term. */
int gd_attn_l1_losses(const float* e, const float* r, const float* t, const float* m_bg, const float* m_edit, const float* m_am,
                      const float* w_am, float c_sim, float c_mov, float c_amo, float c_smh, float c_smw, int H, int S, int d,
                      float* grad, float* partials, int n_partials, void* stream);

/* attention_processors.py:259-266: distance weight, log terms, and the two non-zero gradient entries per row; extra (H,M,ld) and
 * delta_extra (H,M) feed gd_attn_bwd_prep / gd_attn_bwd.  coef = removal weight / (sum(mask_inpaint) * H + 1e-8); if w_dev (device scalar) is given, coef is
 * multiplied by *w_dev on the device: the adaptive schedule (optimization.py:7-105) changes that weight between passes of a captured graph.
 * a_b_bf16 == NULL: `extra` is not written here (gd_attn_probs_rows2 + gd_removal_extra_rows produce it without a materialised base map). */
int gd_removal_finalize(const float* partial, int n_tiles, int H, int M, int S, const int* rows, const float* mask_in,
                        const float* mask_bg, float coef, const float* w_dev, const void* a_b_bf16, int Nb, int Nk, int ld, float* term,
                        float* g2, int* j2, float* delta_extra, float* extra, void* stream);

/* terms6 = {sim, movement, removal, smoothness, amodal, weighted total}; terms_accum6 (or NULL) += terms6.
 * inv6_host = 1/denominators {sim, movement, amodal, smooth_h, smooth_w, removal}; w5_host = weights {sim, movement, amodal, smoothness, removal};
 * w_rem_dev (device scalar or NULL) overrides the removal weight */
int gd_loss_reduce(const float* partials, int n_part, const float* rem_terms, int n_rem, const float* inv6_host, const float* w5_host,
                   const float* w_rem_dev, float amodal_gate, float* terms6, float* terms_accum6, void* stream);

/* attention_sharing.py:68-105 interpolate_from_mask: nearest-4 foreground pixels + weights (mask/grid only: once per resolution). */
int gd_amodal_knn(const float* m_edit, int S, int* idx4, float* val4, float* w, void* stream);

/* attention_processors.py:291-293 + generic_torch.py:145-154: interpolated, foreground-overwritten, 5x5-gaussian-smoothed target. */
int gd_amodal_target(const float* e, const float* m_edit, const int* idx4, const float* val4, const float* gauss25_host, int H, int S,
                     int d, float* scratch, float* target, void* stream);

/* attention_processors.py:617-624, 922-925: out = a * ma[row] + b * mb[row] (a may be NULL; mb NULL = 1); a, b (H,N,d) fp32 contiguous,
 * out a slab with out_strides_host = {row, head} (NULL: contiguous). */
int gd_blend_rows(const float* a, const float* ma, const float* b, const float* mb, int H, int N, int d, void* out, int out_is_bf16,
                  const long* out_strides_host, void* stream);

/* ---- (4) DDIM step and latent update ------------------------------------------------------------------------------- */

/* diffusion.py:46,55: eps = eps_u + g (eps_c - eps_u) if eps_c; x_prev = sqrt(a_prev) (x - sqrt(1-a_t) eps)/sqrt(a_t) + sqrt(1-a_prev) eps */
int gd_ddim_step(const float* x, const void* eps_u, const void* eps_c, int eps_is_bf16, float guidance, float sqrt_one_minus_at,
                 float sqrt_at, float sqrt_aprev, float sqrt_one_minus_aprev, long n, float* out, float* eps_out, void* stream);

/* optimization.py:213-245: nan_to_num(grad); out = (lat - 2 m s g) - (1 - m) s g, m (hw) broadcast over channels (NULL: lat - s g). */
int gd_latent_update(const float* lat, const float* grad, const float* mask, int hw, float step, long n, float* out, void* stream);

/* editor.py:219,316 + generic_torch.py:87: norm_out = sqrt(sum x^2 + 1e-12); if target > 0: x *= target / norm, where target is
 * *target_dev (device scalar) if given, else target_norm. */
int gd_norm_rescale(float* x, long n, float target_norm, const float* target_dev, float* norm_out, void* stream);

/* editor.py:393-399: out = a (1 - m) + m b, m optionally binarised (> 0.5). */
int gd_latent_blend(const float* a, const float* b, const float* mask, int hw, int binarize, long n, float* out, void* stream);

/* ---- (5) post-processing ----------------------------------------------------------------------------------- */

/* image_processing.py:24-77 masked_histogram_matching (editor.py:680,683,690): per channel, the source's values are remapped so that its
 * CDF under mask_source matches the template's CDF under mask.  source, tmpl (npix, C) uint8; masks (npix) float (> 0.5 selects);
 * counts (C,2,256) int32 and lut (C,256) double: scratch; out (npix, C) double -- the reference's np.interp look-up, bit for bit. */
int gd_masked_histogram_match(const unsigned char* source, const unsigned char* tmpl, const float* mask, const float* mask_source, long npix,
                              int C, int* counts, double* lut, double* out, void* stream);

/* ---- caller-side fused op (NOT part of the reference surface; SURVEY 8 row A14 leaves the UNet body to stock torch) -------------
 * GroupNorm (+ SiLU) on channels-last bf16 activations: torch's CUDA group_norm round-trips through NCHW (2 layout copies + 4 kernels
 * per norm, 61 norms per UNet evaluation), which hides the path behind the body.  x, y, dy, dx (B, HW, C) bf16, C % 8 == 0, G <= 32;
 * gamma / beta (C) bf16 or fp32; stats (B, G, 2) = (mean, rstd) saved for the backward; workspace >= gd_group_norm_nhwc_workspace()
 * floats; counters: >= B unsigned ints, zero on entry and zero again on exit (clear once, reuse on the same stream).  The backward returns dx only (the body's weights are frozen in the edit loop, optimization.py:213-219). */
int gd_group_norm_nhwc_fwd(const void* x, const void* pre_bias, const void* gamma, const void* beta, int w_is_bf16, int B, int HW, int C, int G, float eps, int silu,
                           float* workspace, long workspace_floats, unsigned* counters, float* stats, void* y, void* stream);
int gd_group_norm_nhwc_bwd(const void* x, const void* pre_bias, const void* dy, const void* gamma, const void* beta, int w_is_bf16, const float* stats, int B, int HW,
                           int C, int G, int silu, float* workspace, long workspace_floats, unsigned* counters, void* dx, void* stream);
int gd_group_norm_nhwc_workspace(int B, int HW, int C, int G);
/* forward scheme (process-wide knob): 1 = one launch, a cluster of 8 CTAs per (batch entry, slab of groups) staging its rows in shared
 * memory and exchanging partial sums through distributed shared memory (default where the shape fits); 0 = always the two-launch scheme. */
int gd_group_norm_config(int use_cluster_kernel);
/* (pre_bias (B, C) bf16 or NULL: the tensor that is normalised is x + pre_bias[b, c] -- the producing convolution's bias and the
 * time-embedding shift of a ResNet block folded into the norm, instead of two broadcast adds.) */

/* GEGLU of the feed-forward blocks: out (rows, F) = proj[:, :F] * gelu(proj[:, F:]) (exact erf), proj (rows, 2F) bf16, F % 8 == 0; and
 * dproj (rows, 2F) given dy (rows, F). */
int gd_geglu_fwd(const void* proj, long rows, int F, void* out, void* stream);
int gd_geglu_bwd(const void* proj, const void* dy, long rows, int F, void* dproj, void* stream);

/* LayerNorm over the last dimension of x (rows, C) bf16 (one warp per row, row in registers, centred variance): gamma / beta (C) bf16,
 * C % 8 == 0, C <= 1280; mean, rstd (rows) fp32 are the statistics aten::native_layer_norm_backward takes. */
int gd_layer_norm_fwd(const void* x, const void* gamma, const void* beta, long rows, int C, float eps, void* y, float* mean, float* rstd,
                      void* stream);

/* out = a + b + bias[c] for (rows, C) bf16 channels-last tensors: a residual add with the producing convolution's bias folded in. */
int gd_add_bias_residual(const void* a, const void* b, const void* bias, long rows, int C, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
