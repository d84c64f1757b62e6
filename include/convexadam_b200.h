/*
 * convexadam_b200.h -- C ABI of the B200-native convexAdam hot path (libconvexadam_b200.so).
 *
 * Each entry point replaces one torch-op composite of the reference (paths relative to the upstream repo
 * multimodallearning/convexAdam, src/convexAdam/):
 *
 *   ca_mindssc_pooled      MINDSSC + F.avg_pool3d            convex_adam_utils.py:24-68, convex_adam_MIND.py:55-59,118-119,149-150
 *   ca_avgpool3d           F.avg_pool3d                      convex_adam_MIND.py:118-119,149-150,166,191; convex_adam_nnUNet.py:66-67
 *   ca_onehot_pooled       weighted one-hot + avg_pool3d     convex_adam_nnUNet.py:19-38,66-67,98-99
 *   ca_correlate           correlate                         convex_adam_utils.py:72-89
 *   ca_coupled_convex      coupled_convex                    convex_adam_utils.py:93-109
 *   ca_correlate_fast + ca_coupled_convex_certified   the same two, fused use  convex_adam_MIND.py:124-138
 *   ca_inverse_consistency inverse_consistency               convex_adam_utils.py:114-129
 *   ca_field_affine        (x/scale).flip(1), x*scale*gs     convex_adam_MIND.py:134,139,141
 *   ca_trilinear           F.interpolate(mode='trilinear')   convex_adam_MIND.py:141,153,182
 *   ca_adam_run            the Adam instance-optimisation    convex_adam_MIND.py:147-181
 *   ca_adam_run_pools, ca_adam_run_gauss   ... with the search's smoothers   self_configuring/adam_run_withconfig_shiftSpline.py:140-141,217
 *   ca_pack_field          dtype rounding + np.stack(...,3)  convex_adam_MIND.py:198-202
 *   ca_apply_convex        map_coordinates warp              apply_convex.py:13-24
 *   ca_field_mean          np.mean(field[mask], axis=0)      convex_adam_translation.py:88-96
 *   ca_mask_erode, ca_nearest_inside   mask erosion + EDT indices (use_mask)   convex_adam_MIND.py:36-54
 *
 * Conventions: every pointer is a DEVICE pointer to a contiguous array owned by the caller;
 * no function allocates, none synchronises, all work is enqueued on `stream` (a cudaStream_t); the functions are
 * re-entrant per stream.  Return 0 on success, a negative CA_ERR_* otherwise (text via ca_last_error(), per thread).
 * Workspaces: ask ca_*_workspace_bytes first; sizes are upper bounds, any 256-byte aligned device buffer will do.
 * Volumes are (H,W,D) C-order with D fastest; multi-channel arrays are planar (C,H,W,D) unless stated.
 */
#ifndef CONVEXADAM_B200_H
#define CONVEXADAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CA_ABI_VERSION 1

#define CA_OK 0
#define CA_ERR_BAD_ARGUMENT (-1)
#define CA_ERR_UNSUPPORTED (-2)
#define CA_ERR_LAUNCH (-3)
#define CA_ERR_WORKSPACE (-4)

typedef void* ca_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define CA_API __attribute__((visibility("default")))
#else
#define CA_API
#endif

CA_API int ca_abi_version(void);
CA_API const char* ca_last_error(void);
/* 1 when the library was compiled for the GPU (sm_100a); the test-only CPU emulation build returns 0. */
CA_API int ca_is_cuda_build(void);
/* Number of kernels this library has launched in this process so far (bench.py reports the per-step count). */
CA_API unsigned long long ca_launch_count(void);
/* Measurement aid (bench.py): ca_kernel_timing(mask) makes the library bracket the launches of the kernels whose slot
 * bit is set in mask with CUDA
 * events on the launch stream (no synchronisation); after the caller has synchronised, ca_kernel_times sums them per
 * slot (0 raw SSD, 1 box filters, 2 block minima + argmin, 3 coupled-convex selection pass, 4 exact re-decision,
 * 5 Adam gradient, 6 Adam B^3) and returns the number of recorded launches.  ca_kernel_timing(0) switches it off. */
CA_API void ca_kernel_timing(int slot_mask);
CA_API int ca_kernel_times(double* ms, int* count, int nslots);

/* ---- MIND-SSC descriptor fused with block-mean pooling (utils.py:24-68 + MIND.py:118-119) -------------------
 * img (H,W,D) f32.  Up to two pooled outputs (pool factor gs_a / gs_b, 0 = absent; 1 = full resolution, i.e. the
 * plain MINDSSC operator).  out_x: planar (12, H/gs, W/gs, D/gs) f32.  Channel order = the reference's (after the
 * permutation at utils.py:66).  Two passes over the image because of the global mean at utils.py:61; the host
 * sync of the reference (.item()) is replaced by a device scalar. */
CA_API size_t ca_mindssc_workspace_bytes(int H, int W, int D);
CA_API int ca_mindssc_pooled(const float* img, int H, int W, int D, int radius, int dilation, int gs_a, float* out_a,
                      int gs_b, float* out_b, void* workspace, size_t workspace_bytes, ca_stream_t stream);

/* ---- F.avg_pool3d(x, k, stride, padding), count_include_pad=True (torch defaults) -------------------------
 * in (nc, T, H, W) -> out (nc, oT, oH, oW), o = (i + 2*pad - k)/stride + 1.  Exact torch-CPU summation order. */
CA_API int ca_avgpool3d(const float* in, int64_t nc, int T, int H, int W, int k, int stride, int pad, float* out,
                 ca_stream_t stream);

/* ---- nnUNet features (nnUNet.py:19-38) fused with avg_pool3d(gs) ------------------------------------------
 * labels (H,W,D) f32 holding integer class ids; lut (n_lut) int32 maps class id -> channel (or -1 = absent);
 * weight (C) f32 is the value a voxel of that class contributes (the reference's 10*weight, already rounded the
 * way the caller's dtype policy wants).  out planar (C, H/gs, W/gs, D/gs). */
CA_API int ca_onehot_pooled(const float* labels, int H, int W, int D, const int32_t* lut, int n_lut, const float* weight,
                     int C, int gs, float* out, ca_stream_t stream);

/* ---- correlate (utils.py:72-89) -----------------------------------------------------------------------------
 * fix, mov: planar (C,h,w,d) f32 coarse features.  ssd: (L,h,w,d) f32, L = (2*disp_hw+1)^3, label
 * l = kD*S*S + kW*S + kH (reference layout).  argmin: (h,w,d) int64, first minimum; may be NULL.
 * Bit-exact restatement of the torch-CPU arithmetic (sequential channel sum, two zero-padded 27-tap sums, /27).
 * Workspace: ca_correlate_workspace_bytes is the preferred size (a few large launches); any size down to
 * ca_correlate_min_workspace_bytes works (labels are then produced in smaller chunks) with identical results. */
CA_API size_t ca_correlate_workspace_bytes(int C, int h, int w, int d, int disp_hw);
CA_API size_t ca_correlate_min_workspace_bytes(int C, int h, int w, int d, int disp_hw);
CA_API int ca_correlate(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, float* ssd,
                 int64_t* argmin, void* workspace, size_t workspace_bytes, ca_stream_t stream);
/* The same for the H planes [x_begin, x_end) only: ssd (L, x_end-x_begin, w, d), argmin (x_end-x_begin, w, d).  The
 * features are still the whole (C,h,w,d) volumes (a slab needs 2 + 2*disp_hw planes of context).  Used when one
 * oversized pair is split into H slabs over several GPUs; values are those of the full call. */
CA_API size_t ca_correlate_slab_workspace_bytes(int C, int h, int w, int d, int disp_hw, int x_begin, int x_end);
CA_API int ca_correlate_slab(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int x_begin,
                      int x_end, float* ssd, int64_t* argmin, void* workspace, size_t workspace_bytes,
                      ca_stream_t stream);

/* ---- coupled_convex (utils.py:93-109) -----------------------------------------------------------------------
 * ssd (L,h,w,d); argmin (h,w,d) int64, or NULL: the plain argmin of utils.py:87 is then taken here (saves one pass
 * over the volume when ca_correlate was called with argmin == NULL); mesh (3,L) f32 = disp_mesh_t of MIND.py:127
 * (pass the values torch's affine_grid produces -- they are not integral).  disp_soft: (3,h,w,d) f32 out,
 * coarse-voxel units.  sel_out: optional (6,h,w,d) int64, the argmin of each of the six passes (for parity tests).
 * One streaming pass builds per-voxel minima over blocks of 2*disp_hw+1 consecutive labels; the six passes then read
 * those (1/(2*disp_hw+1) of the volume) plus only the blocks whose lower bound does not exceed the cost of the
 * previous selection -- an exact pruning, the selections are bit-identical to the full passes (coupled.cu). */
CA_API size_t ca_coupled_convex_workspace_bytes(int h, int w, int d, int disp_hw);
CA_API int ca_coupled_convex(const float* ssd, const int64_t* argmin, const float* mesh, int h, int w, int d, int disp_hw,
                      float* disp_soft, int64_t* sel_out, void* workspace, size_t workspace_bytes,
                      ca_stream_t stream);

/* ---- correlate + coupled_convex, "certified" mode: the fused path convex_adam_pt uses (MIND.py:124-138) ----------
 * Inside convex_adam_pt the cost volume never escapes: it is only ever consumed through argmin (utils.py:87) and the six
 * selections of coupled_convex (utils.py:104-105).  ca_correlate_fast builds it with separable running sums and FMAs
 * (about 38 instead of 94 fp32 operations per entry); every entry is within a relative 98 * 2^-24 of the reference's
 * value (both are sums of the same non-negative terms; at most 68 roundings there, 30 here).
 * ca_coupled_convex_certified then takes argmin and the six selections from it and proves each one: a voxel whose
 * runner-up cost lies outside the band [best, best * (1 + 2^-16)] cannot have another selection in the reference's
 * arithmetic; the few voxels inside the band are re-decided from exact costs, recomputed in the reference's operation
 * order from the feature volumes (fix, mov: the same (C,h,w,d) arrays given to ca_correlate_fast).  The selections and
 * disp_soft are therefore bit-identical to ca_correlate + ca_coupled_convex; only the volume is approximate.
 * sel_out: optional (7,h,w,d) int64 -- plane 0 the argmin of utils.py:87, planes 1..6 the six passes.
 * flagged_out: optional 7 int32 on the device, the number of voxels that needed the exact re-decision in each.
 * Implementation of ca_correlate_fast: per chunk of (kH,kW) pairs a raw-SSD kernel staged by tensor-map TMA straight from
 * the caller's unpadded feature volumes (out-of-bounds zero fill = the reference's F.pad, utils.py:76) and a separable
 * box-filter kernel; the environment variable CA_CORR_FUSED=1 selects the one-kernel form of the stage
 * (csrc/correlate_fused.cu: a third of the DRAM traffic, same bound on every entry, same certified selections). */
CA_API int ca_correlate_fast(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, float* ssd,
                      void* workspace, size_t workspace_bytes, ca_stream_t stream);
CA_API size_t ca_coupled_convex_certified_workspace_bytes(int h, int w, int d, int disp_hw);
CA_API int ca_coupled_convex_certified(const float* ssd, const float* fix, const float* mov, int C, const float* mesh, int h,
                                int w, int d, int disp_hw, float* disp_soft, int64_t* sel_out, int32_t* flagged_out,
                                void* workspace, size_t workspace_bytes, ca_stream_t stream);

/* The certified mode in pieces, for the slab-parallel path: ca_correlate_slab_fast is ca_correlate_fast for the H planes
 * [x_begin, x_end); ca_cc_blockmin_certified builds bmin (S*S + 2*S, n: block minima, slab minima, slab runner-ups) and the
 * plain argmin of the slab, certified exact; ca_cc_select_certified is one certified selection pass (ca_cc_select with
 * pruning).  fix, mov: the whole (C,h,w,d) feature volumes; x_begin: the slab's first plane; n = slab voxels. */
CA_API int ca_correlate_slab_fast(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int x_begin,
                           int x_end, float* ssd, void* workspace, size_t workspace_bytes, ca_stream_t stream);
CA_API size_t ca_cc_certified_workspace_bytes(int64_t n);
CA_API int ca_cc_blockmin_certified(const float* ssd, int64_t n, int disp_hw, const float* mesh, float* bmin, int64_t* argmin,
                             const float* fix, const float* mov, int C, int h, int w, int d, int x_begin, void* workspace,
                             size_t workspace_bytes, ca_stream_t stream);
CA_API int ca_cc_select_certified(const float* ssd, int64_t n, int disp_hw, const float* mesh, const float* soft, float coeff,
                           const float* bmin, const int64_t* prev, int64_t* sel, const float* fix, const float* mov, int C, int h,
                           int w, int d, int x_begin, void* workspace, size_t workspace_bytes, ca_stream_t stream);

/* coupled_convex in pieces (the loop of utils.py:98-107 run by the host, for H slabs with a halo exchange in between):
 * ca_cc_blockmin: bmin (S*S + S, n): per-voxel minima over the S*S blocks of S consecutive labels, followed by the minima
 *               over the S slabs of S blocks; and (argmin != NULL) the plain argmin of utils.py:87 in the same pass.
 * ca_cc_select: sel(x) = argmin_l ( ssd_l(x) + coeff * || mesh_l - soft(x) ||^2 ) over n voxels (utils.py:104-105);
 *               mesh == soft == NULL gives the plain argmin of utils.py:87.  ssd (L,n), soft (3,n), sel (n) int64.
 *               bmin + prev (the previous selection) enable the exact pruning; both NULL = full pass.
 * ca_cc_smooth: soft = box3(mesh[:, sel]) (utils.py:96,107) for the planes [1, P-1) of a (P,w,d) window of labels whose
 *               first/last plane come from the neighbouring slab; a negative label = outside the volume. soft (3,P-2,w,d). */
CA_API size_t ca_cc_select_workspace_bytes(int64_t n);
CA_API int ca_cc_blockmin(const float* ssd, int64_t n, int disp_hw, float* bmin, int64_t* argmin, void* workspace,
                   size_t workspace_bytes, ca_stream_t stream);
CA_API int ca_cc_select(const float* ssd, int64_t n, int disp_hw, const float* mesh, const float* soft, float coeff,
                 const float* bmin, const int64_t* prev, int64_t* sel, void* workspace, size_t workspace_bytes,
                 ca_stream_t stream);
CA_API int ca_cc_smooth(const int64_t* sel_window, const float* mesh, int P, int w, int d, int disp_hw, float* soft,
                 ca_stream_t stream);

/* ---- inverse_consistency (utils.py:114-129) -------------------------------------------------------------------
 * f1, f2: (3,h,w,d) f32, normalised units, channel 0 = x (last array axis) as in the reference call at
 * MIND.py:139.  gx (d), gy (w), gz (h): the identity-grid axes of F.affine_grid(eye, (1,1,h,w,d)).
 * o1, o2: results after `iters` fixed-point iterations.  workspace: 2 * 3*h*w*d floats. */
CA_API size_t ca_inverse_consistency_workspace_bytes(int h, int w, int d);
CA_API int ca_inverse_consistency(const float* f1, const float* f2, int h, int w, int d, int iters, const float* gx,
                           const float* gy, const float* gz, float* o1, float* o2, void* workspace,
                           size_t workspace_bytes, ca_stream_t stream);

/* ---- per-channel affine on a 3-channel field, optional channel flip (MIND.py:134,139,141) ---------------------
 * out[c] = (in[flip ? 2-c : c] / div[c']) * mul1[c'] * mul2   with c' = the source channel, applied in that order,
 * each op rounded (div/mul by 1 are exact no-ops). */
CA_API int ca_field_affine(const float* in, float* out, int64_t n_per_channel, int flip, const float div[3],
                    const float mul1[3], float mul2, ca_stream_t stream);

/* ---- F.interpolate(mode='trilinear', align_corners=False) (MIND.py:141,153,182) ------------------------------
 * in (C,ih,iw,id) -> out (C,oh,ow,od);  out = interp(in) * post_mul  (post_mul = 1 for the plain operator). */
CA_API int ca_trilinear(const float* in, int C, int ih, int iw, int id, float* out, int oh, int ow, int od, float post_mul,
                 ca_stream_t stream);

/* ---- Adam instance optimisation (MIND.py:147-181) ---------------------------------------------------------------
 * fix, mov: planar (C,ha,wa,da) f32 features on the Adam grid.  p, m, v: (3,ha,wa,da) f32 state, updated in
 * place (p = net[0].weight, m/v = exp_avg/exp_avg_sq; zero m,v and first_step=1 for a fresh run).  Runs steps
 * first_step .. first_step+niter-1 (1-based Adam step numbers; first_step>1 resumes / teacher-forces).
 * u_out (3,ha,wa,da): B^3(p) of the LAST forward, i.e. what the reference returns (the final update is applied
 * to p but does not show in u_out, MIND.py:181).  g_out: optional, dL/dp of the last iteration.
 * g0x (da), g0y (wa), g0z (ha): identity-grid axes of affine_grid (MIND.py:160). lr=1, betas=(0.9,0.999), eps=1e-8. */
CA_API size_t ca_adam_workspace_bytes(int C, int ha, int wa, int da);
CA_API int ca_adam_run(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y, const float* g0z,
                float* u_out, float* g_out, void* workspace, size_t workspace_bytes, ca_stream_t stream);

/* The same loop with the smoother of the self-configuring search in place of B^3 (self_configuring/
 * adam_run_withconfig_shiftSpline.py:217, `avgs[avg_n](net[0].weight)` with avgs = kovesi_spline(sigma, 4),
 * convexAdam_hyper_util.py:475-488): the sequence AvgPool3d(widths[i], stride 1, padding (widths[i]-1)/2), i = 0..nwidths-1
 * (zero padding, divisor w^3; widths odd; `widths` is a HOST array).  Forward u = P_n ... P_1 p, backward the same pools in
 * reverse order.  Everything else as ca_adam_run. */
CA_API int ca_adam_run_pools(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                      float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y, const float* g0z,
                      const int* widths, int nwidths, float* u_out, float* g_out, void* workspace, size_t workspace_bytes,
                      ca_stream_t stream);

/* ... and with the search's GaussianSmoothing(sigma) (convexAdam_hyper_util.py:454-473 over filter1D, :425-437): along each axis in
 * turn, replicate padding by nweights/2 and a valid cross-correlation with the nweights normalised weights (a HOST array,
 * computed by the caller as the reference computes them); backward = the transposed filters in reverse axis order. */
CA_API int ca_adam_run_gauss(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                      float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y, const float* g0z,
                      const float* weights, int nweights, float* u_out, float* g_out, void* workspace, size_t workspace_bytes,
                      ca_stream_t stream);

/* The same on a window of the Adam grid: p, m, v, u_out, g_out hold the H planes [lo, hi) only, (3, hi-lo, wa, da);
 * fix, mov and the grid axes stay whole-grid.  Window edges that are not faces of the grid are treated like faces, so
 * after one iteration only the planes at least 7 away from such an edge are exact (3 + 1 + 3: B^3, the regulariser's
 * neighbours, B^3): the caller owns the interior and refreshes the margin of p from its neighbours every iteration.
 * features_ready != 0: the channel-last copies of fix/mov made by an earlier call are still in `workspace`. */
CA_API size_t ca_adam_window_workspace_bytes(int C, int ha, int wa, int da, int lo, int hi);
CA_API int ca_adam_run_window(const float* fix, const float* mov, int C, int ha, int wa, int da, int lo, int hi,
                       float lambda_weight, float* p, float* m, float* v, int first_step, int niter, const float* g0x,
                       const float* g0y, const float* g0z, float* u_out, float* g_out, void* workspace,
                       size_t workspace_bytes, int features_ready, ca_stream_t stream);

/* ---- output packing (MIND.py:198-202) --------------------------------------------------------------------------
 * field (3,H,W,D) f32 -> out (H,W,D,3) f64; round_half != 0 rounds every value through IEEE half first
 * (the reference's `.to(dtype)` when dtype=float16). */
CA_API int ca_pack_field(const float* field, int H, int W, int D, int round_half, double* out, ca_stream_t stream);

/* ---- apply_convex (apply_convex.py:13-24): the warp either side of the path --------------------------------------
 * warped(x) = moving interpolated linearly (float64 arithmetic) at x + disp(x); 0 where any coordinate leaves
 * [0, size-1] (scipy.ndimage.map_coordinates, order=1, mode='constant').  disp (H,W,D,3) f64 -- the array
 * convex_adam_pt returns; moving, warped (H,W,D) f32 or f64 (moving_is_f64). */
CA_API int ca_apply_convex(const double* disp, const void* moving, int moving_is_f64, int H, int W, int D, void* warped,
                    ca_stream_t stream);

/* ---- mean displacement (convex_adam_translation.py:88-96) -----------------------------------------------------------
 * out4 = (mean of channel 0, 1, 2 over the voxels with mask != 0 -- all voxels when mask is NULL --, voxel count) in
 * float64; field planar (3,n) f32 on the device, mask optional (n) uint8; round_half as in ca_pack_field.  Lets a
 * translation-only caller skip the transfer of the dense field. */
CA_API size_t ca_field_mean_workspace_bytes(void);
CA_API int ca_field_mean(const float* field, const unsigned char* mask, int64_t n, int round_half, double* out4, void* workspace,
                  size_t workspace_bytes, ca_stream_t stream);

/* ---- use_mask pre-processing of extract_features (convex_adam_MIND.py:36-54) ------------------------------------------------
 * ca_mask_erode: out = (avg3(mask) > 0.9) as 0/1 floats, avg3 = ReplicationPad3d(1) + AvgPool3d(3, stride 1) (MIND.py:38-42).
 * ca_nearest_inside: idx (3,h,w,d) int32 = for every voxel the coordinates of the nearest voxel with inside != 0 --
 * scipy.ndimage.distance_transform_edt(inside == 0, return_indices=True)[1] (MIND.py:43), including scipy's choice among
 * equidistant voxels (the reference calls scipy on the host here; this is the same feature transform on the device). */
CA_API int ca_mask_erode(const float* mask, int H, int W, int D, float* out, ca_stream_t stream);
CA_API size_t ca_nearest_inside_workspace_bytes(int h, int w, int d);
CA_API int ca_nearest_inside(const float* inside, int h, int w, int d, int32_t* idx, void* workspace, size_t workspace_bytes,
                      ca_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CONVEXADAM_B200_H */
