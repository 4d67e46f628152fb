/* libcoalign_b200 - C ABI of the B200-native CoAlign hot path.
 *
 * This is the drop-in boundary below the reference's Python model
 * (/root/reference/opencood/models/point_pillar_baseline_multiscale.py:93-135): every entry point
 * replaces one stage of that forward (citations on each function).  Conventions:
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - returns 0 on success, a positive cudaError_t, or a negative CB_ERR_* argument/driver error;
 *   - never allocates and never synchronises: the caller owns inputs, outputs and workspaces,
 *     all work is enqueued on `stream` (CUDA-graph capturable);
 *   - no global state except the lazily resolved cuTensorMapEncodeTiled entry point and
 *     per-kernel shared-memory attributes (idempotent, thread-safe).
 *
 * Activation layouts (bf16, channels innermost):
 *   PF  "padded flat": [n][H+2][W+2][C]; the 1-pixel halo is zero and is never written.
 *       A 3x3/s1 convolution is then a GEMM over the flattened padded pixel index with one
 *       constant row shift per filter tap.
 *   PS  "phase split": tensor (n,H,W,C) stored as 4 parity planes [(h&1)*2+(w&1)][n][Ho+2][Wo+2][C],
 *       Ho=ceil(H/2), Wo=ceil(W/2) (each plane PF-padded).  Input format of the stride-2 convs:
 *       every tap of a 3x3/s2 (or 1x1/s2) convolution becomes (plane, constant row shift).
 *   "precise" mode keeps a second bf16 plane (lo = x - bf16(x)) at element offset `lo_off` behind
 *   the hi plane and evaluates hi*hi + lo*hi + hi*lo (fp32-class accuracy on bf16 tensor cores).
 */
#ifndef COALIGN_B200_H
#define COALIGN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_MAX_KSTEPS 168
#define CB_MAX_AGENTS 64
#define CB_MAX_HEADS 4    /* cls, reg, dir and the stage-1 detector's uncertainty head (point_pillar_uncertainty.py:34) */

/* --------------------------------------------------------------------------------------------
 * cb_version / cb_device_check
 * ------------------------------------------------------------------------------------------ */
int cb_version(void);
/* 0 when the current device is sm_100 (B200); negative otherwise.  The Python side raises - there is
 * no CPU or generic-GPU fallback. */
int cb_device_check(void);

/* --------------------------------------------------------------------------------------------
 * A1/A2  voxelisation - replaces spconv VoxelGeneratorV2.generate / Point2VoxelCPU3d.point_to_voxel
 * as called from /root/reference/opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:62-85
 * and the collate at :145-174 (agent index prepended to coords).
 *
 * points      : concatenated clouds, float32 [sum_P][4]; agent a owns rows [pt_offset[a], pt_offset[a+1])
 * pt_offset   : HOST int32 [n_agents+1]
 * range/vsize : HOST float32 [6]/[3]; grid: HOST int32 [3] = (nx,ny,nz)
 * outputs (device), bit-exact with the serial generator:
 *   voxels     float32 [cap][max_pts][4] zero padded     (cap >= min(sum_P, n_agents*max_voxels))
 *   coords     int32   [cap][4] = [agent, z, y, x]
 *   num_points int32   [cap]
 *   n_voxels   int32   [n_agents+1]: per-agent counts, last = total
 * workspace    : cb_voxelize_workspace_bytes(...) bytes, device
 * ------------------------------------------------------------------------------------------ */
size_t cb_voxelize_workspace_bytes(int n_agents, int sum_points, const int32_t* grid, int max_voxels);
int cb_voxelize(const float* points, const int32_t* pt_offset, int n_agents,
                const float* range, const float* vsize, const int32_t* grid,
                int max_pts, int max_voxels,
                float* voxels, int32_t* coords, int32_t* num_points, int32_t* n_voxels,
                void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * A3+A4+A5  PillarVFE + PFNLayer + PointPillarScatter fused
 *   /root/reference/opencood/models/sub_modules/pillar_vfe.py:31-53,105-155
 *   /root/reference/opencood/models/sub_modules/point_pillar_scatter.py:15-72
 * Reads reference-format voxel tensors, writes the BEV canvas in PS layout (the canvas is only ever
 * consumed by stride-2 convolutions).  The canvas must have been zeroed by the caller.
 *   w        float32 [64][10]  pfn_layers.0.linear.weight
 *   scale/shift float32 [64]   folded eval BatchNorm1d (eps 1e-3)
 *   vsize/center_off HOST float32 [3]: voxel size and (voxel/2 + range_min) rounded from double
 *                              exactly like PillarVFE.__init__ (pillar_vfe.py:84-89)
 *   canvas_agents: agent capacity the PS canvas was allocated for (plane stride), >= n_agents
 *   dirty_rows/dirty_count: optional DEVICE outputs for cb_canvas_clear (may be NULL)
 *   n_voxels_dev: optional DEVICE int32* holding the number of valid rows (<= n_rows); NULL = n_rows
 * ------------------------------------------------------------------------------------------ */
int cb_pfn_scatter(const float* voxels, const int32_t* coords, const int32_t* num_points,
                   int n_rows, const int32_t* n_voxels_dev, int max_pts,
                   const float* w, const float* scale, const float* shift,
                   const float* vsize, const float* center_off,
                   int n_agents, int canvas_agents, int ny, int nx,
                   void* canvas_ps, int64_t lo_off,
                   int64_t* dirty_rows, int32_t* dirty_count, void* stream);

/* Same stages straight from raw points (A1..A5 fused; the (M,32,4) voxel tensor is never
 * materialised).  Uses the same workspace as cb_voxelize (size from cb_voxelize_workspace_bytes).
 * The canvas does not depend on the order of the voxels, so this path does not run the ordered scans of
 * cb_voxelize unless a cap is hit (agents with more points than max_voxels; cells with more than max_pts
 * points): results are identical to cb_voxelize + cb_pfn_scatter bit for bit.
 * Thread safety: the derived PFN coefficients are staged in ONE constant-memory table per process, written on
 * `stream` right before the kernel that reads it: calls with DIFFERENT PFN weights must not be in flight on
 * different streams at the same time (same weights, or one stream: no restriction). */
int cb_points_to_canvas(const float* points, const int32_t* pt_offset, int n_agents,
                        const float* range, const float* vsize, const int32_t* grid,
                        int max_pts, int max_voxels,
                        const float* w, const float* scale, const float* shift,
                        const float* center_off, int canvas_agents, void* canvas_ps, int64_t lo_off,
                        int64_t* dirty_rows, int32_t* dirty_count,
                        void* workspace, size_t workspace_bytes, void* stream);

/* cb_points_to_canvas with the per-agent point offsets in DEVICE memory (int32 [n_agents+1], off[0] = 0, non-decreasing,
 * off[n_agents] <= point_capacity, no agent with more than agent_capacity points): grid sizes and the launch sequence
 * depend only on the two capacities, so a CUDA graph captured once serves frames whose clouds have different sizes
 * (real LiDAR sweeps do; sp_voxel_preprocessor.py:62-85 is called per cloud).  Results are identical to
 * cb_points_to_canvas on the same offsets.  The max_voxels path runs whenever agent_capacity > max_voxels. */
int cb_points_to_canvas_dev(const float* points, const int32_t* pt_offset_dev, int n_agents, int point_capacity,
                            int agent_capacity, const float* range, const float* vsize, const int32_t* grid,
                            int max_pts, int max_voxels,
                            const float* w, const float* scale, const float* shift,
                            const float* center_off, int canvas_agents, void* canvas_ps, int64_t lo_off,
                            int64_t* dirty_rows, int32_t* dirty_count,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Stream-ordered upload of up to CB_MAX_AGENTS+1 int32 values (they travel as kernel arguments: no pinned staging buffer
 * whose lifetime the caller would have to manage).  Used for the offsets above and the scene prefix sums. */
int cb_upload_i32(const int32_t* host_vals, int n, int32_t* dst_dev, void* stream);

/* Sparse canvas reset: zero the cells listed in dirty_rows[0..*dirty_count) (DEVICE arrays written by the two
 * entry points above when their dirty_rows/dirty_count arguments are non-NULL: one canvas row per pillar, -1 =
 * skipped).  Replaces the full-canvas memset between frames (the canvas is ~80 % zeros). */
int cb_canvas_clear(void* canvas_ps, int64_t lo_off, const int64_t* dirty_rows, const int32_t* dirty_count,
                    int capacity, void* stream);

/* --------------------------------------------------------------------------------------------
 * A6..A9  convolutions as implicit GEMMs on tcgen05 tensor cores (TMA-fed, TMEM accumulators)
 *   BasicBlock / ResNetModified   /root/reference/opencood/models/sub_modules/resblock.py:53-69,212-221
 *   deblocks (ConvTranspose k==s) /root/reference/opencood/models/sub_modules/base_bev_backbone_resnet.py:52-65,121-138
 *   shrink header                 /root/reference/opencood/models/sub_modules/downsample_conv.py:18-24
 *   heads                         /root/reference/opencood/models/point_pillar_baseline_multiscale.py:55-63,126-133
 *
 * One GEMM:  D[q][n] = sum_steps sum_{kk<64} A_sel[q + row_off][col + kk] * W[n][w_k + kk]
 * over the flattened padded output-pixel index q (rows_total = n_img*Hp*Wp), followed by the epilogue
 *   v = D + bias[n % cout_mod] (+ residual[q][n]) ; relu ; store.
 * BatchNorm scale is folded into W by the host; `bias` carries the BN shift or the conv bias.
 * ------------------------------------------------------------------------------------------ */
typedef struct cb_kstep {
    int32_t row_off;  /* row shift (tap) + plane offset, in rows of the A tensor */
    int32_t w_k;      /* K coordinate of the 64-wide weight block */
    uint16_t col;     /* first input channel of the 64-wide block */
    uint16_t a_sel;   /* which A tensor (0/1) */
} cb_kstep;

enum { CB_OUT_PF = 0, CB_OUT_PS = 1, CB_OUT_UPSAMPLE = 2, CB_OUT_HEADS = 3 };

typedef struct cb_conv_desc {
    /* A operands: bf16 [a_rows][a_pitch] (a_rows counts hi+lo planes in precise mode) */
    const void* a_ptr[2];
    int64_t a_rows[2];
    int32_t a_pitch[2];
    /* weights: bf16 [w_rows][w_k_total], K-major */
    const void* w_ptr;
    int32_t w_rows;
    int32_t w_k_total;
    /* GEMM row space */
    int32_t n_img, Hp, Wp;     /* rows_total = n_img*Hp*Wp; interior = 1..Hp-2 x 1..Wp-2 */
    int32_t n_total;           /* GEMM N (multiple of block_n) */
    int32_t block_n;           /* 32, 64, 128 or 256 */
    /* epilogue */
    const float* bias;         /* [cout_mod] */
    int32_t cout_mod;
    int32_t relu;
    const void* residual;      /* bf16 PF, same row space, or NULL */
    int32_t res_pitch;
    int64_t res_lo_off;        /* element offset of the lo plane (precise) */
    void* out;
    int32_t out_pitch;
    int32_t out_ch_off;
    int64_t out_lo_off;        /* 0 = bf16 mode; else precise: element offset of the lo plane */
    int32_t out_mode;          /* CB_OUT_* */
    int32_t up_k;              /* CB_OUT_UPSAMPLE: kernel==stride */
    int32_t out_Hp, out_Wp;    /* padded dims of the destination (PS plane dims / upsample target) */
    int64_t out_plane_rows;    /* CB_OUT_PS: rows per parity plane */
    /* CB_OUT_HEADS: fp32 NCHW outputs, channel segments [c0, c0+cn) */
    float* head_out[CB_MAX_HEADS];
    int32_t head_c0[CB_MAX_HEADS];
    int32_t head_cn[CB_MAX_HEADS];
    int32_t n_heads;
    /* K loop */
    int32_t n_ksteps;
    cb_kstep ksteps[CB_MAX_KSTEPS];
} cb_conv_desc;

/* tcgen05 path (the product).  max_ctas <= 0: one persistent CTA per SM. */
int cb_conv_gemm(const cb_conv_desc* desc, int max_ctas, void* stream);
/* Same GEMM on CTA pairs (tcgen05 cta_group::2, 256 x block_n tiles, cluster of 2): halves the weight-tile
 * shared-memory traffic per MAC.  max_clusters <= 0: one cluster per SM pair.  block_n 32 falls back to cb_conv_gemm. */
int cb_conv_gemm_pair(const cb_conv_desc* desc, int max_clusters, void* stream);
/* Channel-major variant for n_total == cout_mod == 128, bf16 PF/PS outputs: the weight tile is the UMMA A operand
 * (M = 128 output channels) and 256 pixels the N side, so one instruction does 128 x 256 x 16 MACs and the activation
 * rows are read from shared memory once per 256-wide tile.  Same descriptor; block_n is ignored.  Returns
 * CB_ERR_ARG for descriptors outside that envelope (callers fall back to cb_conv_gemm). */
int cb_conv_gemm_t(const cb_conv_desc* desc, int max_ctas, void* stream);
/* Halo / resident-weight variant for n_total == cout_mod == 64 layers with at most 10 K-steps, bf16 PF/PS outputs:
 * K-steps whose row shifts are consecutive (the three taps of a 3x3 filter row in the flattened PF row space) share ONE
 * 130-row TMA box - the tap shift is applied to the UMMA descriptor start address - and the weight matrix stays resident
 * in shared memory for the whole launch.  Cuts the TMA shared-memory writes of the N = 64 layers ~4x (they are
 * shared-memory-bandwidth bound).  Returns CB_ERR_ARG outside that envelope (callers use cb_conv_gemm). */
int cb_conv_gemm_halo(const cb_conv_desc* desc, int max_ctas, void* stream);
/* cb_conv_gemm_t with halo boxes: the 256-pixel activation tile of the three taps of a filter row is loaded once
 * (258 rows) and the tap shift is applied to the UMMA B descriptor.  Same envelope as cb_conv_gemm_t. */
int cb_conv_gemm_t_halo(const cb_conv_desc* desc, int max_ctas, void* stream);
/* Plain SIMT fp32-accumulate evaluation of the same descriptor: a validation kernel for the
 * tensor-core path (tests only; never used by the model). */
int cb_conv_gemm_simt(const cb_conv_desc* desc, void* stream);

/* --------------------------------------------------------------------------------------------
 * A10..A13  pose normalisation + affine feature warp + per-pixel ego-row attention, one kernel
 *   normalize_pairwise_tfm  /root/reference/opencood/utils/transformation_utils.py:69-91
 *   warp_affine_simple      /root/reference/opencood/models/sub_modules/torch_transformation_utils.py:322-331
 *   AttFusion.forward       /root/reference/opencood/models/fuse_modules/fusion_in_one.py:96-136
 *   ScaledDotProductAttention /root/reference/opencood/models/fuse_modules/att_fuse.py:43-47
 *   regroup                 /root/reference/opencood/models/fuse_modules/fusion_in_one.py:21-24
 *
 * feat      : bf16, (sum_agents, H, W, C) in PF (in_ps=0) or PS (in_ps=1) layout; `sum_agents` is the
 *             agent capacity of the buffer (PS plane stride)
 * affine    : DEVICE float64 [n_scenes][max_cav][2][3] = normalised ego-row matrices (A_n = affine[b][0][n])
 *             as produced by cb_normalize_affine
 * agent_off : DEVICE int32 [n_scenes+1], prefix sums of record_len
 * out       : bf16 PF (n_scenes, H, W, C)
 * method    : 0 = attention ("att"), 1 = element-wise max ("max", MaxFusion fusion_in_one.py:83-86)
 * ------------------------------------------------------------------------------------------ */
int cb_normalize_affine(const double* pairwise_t_matrix, int n_scenes, int max_cav,
                        int H, int W, double discrete_ratio, double* affine_out, void* stream);
int cb_warp_att_fuse(const void* feat, int in_ps, int64_t in_lo_off, int sum_agents,
                     const double* affine, const int32_t* agent_off, int n_scenes, int max_cav,
                     int H, int W, int C, int method,
                     void* out_pf, int64_t out_lo_off, void* stream);

/* --------------------------------------------------------------------------------------------
 * Detection post-processing (the step after the forward), intermediate fusion:
 *   VoxelPostprocessor.post_process / delta_to_boxes3d
 *       /root/reference/opencood/data_utils/post_processor/voxel_postprocessor.py:243-449
 *   boxes_to_corners_3d, project_box3d, remove_large_pred_bbx, remove_bbx_abnormal_z, nms_rotated,
 *   mask_boxes_outside_range_numpy   /root/reference/opencood/utils/box_utils.py:152-204,278-316,384-421,693-738,840-890
 *   limit_period, rotate_points_along_z, compute_iou (shapely polygon IoU)
 *       /root/reference/opencood/utils/common_utils.py:70-79,105-127,196-218
 * Each of the n_scenes head-output sets is one independent problem (the reference asserts batch 1 per call).
 *   cls/reg/dir_preds : DEVICE float32 NCHW (n, A, H, W), (n, 7A, H, W), (n, num_bins*A, H, W); dir_preds may be NULL
 *   anchors           : DEVICE float32 [H][W][A][7] (generate_anchor_box, cast to float like delta_to_boxes3d does)
 *   tfm               : DEVICE float32 [n][4][4] cav -> ego (`transformation_matrix`)
 *   gt_range          : HOST float64 [6]
 *   order_hwl         : 1 for params['order'] == 'hwl' (boxes [x,y,z,h,w,l,yaw]), 0 otherwise
 *   top_k             : boxes kept for the NMS (reference: 1000), <= 1024
 * outputs (device): out_boxes [n][top_k][8][3] corners, out_scores [n][top_k], both in pick order (score descending);
 *   out_count [n][2] = {boxes returned, anchors above score_threshold (0 -> the reference returns (None, None))}.
 * Ties between equal scores are broken by the lower anchor index (numpy's argsort order is unspecified there).
 * ------------------------------------------------------------------------------------------ */
size_t cb_postprocess_workspace_bytes(int n_scenes, int H, int W, int anchor_num);
int cb_postprocess(const float* cls_preds, const float* reg_preds, const float* dir_preds,
                   int n_scenes, int H, int W, int anchor_num, int num_bins,
                   const float* anchors, const float* tfm,
                   float score_threshold, float dir_offset, float nms_thresh, const double* gt_range,
                   int order_hwl, int top_k,
                   float* out_boxes, float* out_scores, int32_t* out_count,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Stage-1 variant: UncertaintyVoxelPostprocessor.post_process_stage1
 *   /root/reference/opencood/data_utils/post_processor/uncertainty_voxel_postprocessor.py:31-118
 * Per agent: sigmoid + threshold, delta_to_boxes3d, direction-bin fix, corners in the agent's OWN frame (no projection),
 * rotated NMS over the top_k candidates; no size / z filter and no range mask.  Next to the corners it returns the
 * 7-parameter boxes and the flat anchor index (h*W + w)*A + a of every kept box, so that the caller gathers
 * unc_preds[n, a*uncertainty_dim + k, h, w] without another pass (the reference masks unc_preds the same way, :42-56).
 *   out_corners [n][top_k][8][3], out_boxes7 [n][top_k][7], out_index [n][top_k], out_scores [n][top_k], all in pick
 *   order; out_count [n][2] = {boxes kept, anchors above score_threshold}. */
int cb_postprocess_stage1(const float* cls_preds, const float* reg_preds, const float* dir_preds,
                          int n_agents, int H, int W, int anchor_num, int num_bins,
                          const float* anchors, float score_threshold, float dir_offset, float nms_thresh,
                          int order_hwl, int top_k,
                          float* out_corners, float* out_boxes7, int32_t* out_index, float* out_scores,
                          int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * Training step, first piece: PointPillarLoss and its gradients w.r.t. the head outputs
 *   /root/reference/opencood/loss/point_pillar_loss.py:36-116,119-158,201-245
 * sigmoid focal classification loss + smooth-L1 regression loss on the sin-difference encoding + direction-bin cross
 * entropy, normalised like the reference (per-sample positive count, / batch size, loss weights).
 *   cls/reg/dir_preds : DEVICE float32 NCHW (n, A, H, W), (n, 7A, H, W), (n, num_bins*A, H, W); dir_preds may be NULL
 *   pos_equal_one, neg_equal_one : DEVICE (n, H, W, A); targets (n, H, W, 7A); float64 when labels_f64 != 0 (what the
 *                       reference's collate produces), else float32
 *   anchor_yaw_rad    : HOST float64 [A] (dir.args.anchor_yaw in radians), read when dir_preds != NULL
 *   out_losses        : DEVICE float32 [4] = {total_loss, reg_loss, cls_loss, dir_loss} (the reference's loss_dict)
 *   grad_cls/reg/dir  : optional DEVICE float32 outputs, d(total_loss)/d(preds), same NCHW shapes as the inputs
 * Two-stage reductions in a fixed order (bit-reproducible); three launches on the caller's stream.
 * ------------------------------------------------------------------------------------------ */
size_t cb_pointpillar_loss_workspace_bytes(int n, int H, int W, int anchor_num);
int cb_pointpillar_loss(const float* cls_preds, const float* reg_preds, const float* dir_preds,
                        const void* pos_equal_one, const void* neg_equal_one, const void* targets, int labels_f64,
                        int n, int H, int W, int anchor_num, int num_bins,
                        float pos_cls_weight, float alpha, float gamma, float cls_weight,
                        float sigma, float reg_weight, float dir_weight, float dir_offset,
                        const double* anchor_yaw_rad,
                        float* out_losses, float* grad_cls, float* grad_reg, float* grad_dir,
                        void* workspace, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * layout helpers (tests, debugging, interop): dense NCHW float32 <-> PF / PS bf16
 * ------------------------------------------------------------------------------------------ */
int cb_nchw_to_layout(const float* src, int n, int c, int h, int w, int to_ps,
                      void* dst, int64_t lo_off, void* stream);
/* Exact PS -> PF copy of (n, H, W, C) bf16 maps (hi and, when the offsets are non-zero, lo planes).  n_cap = agent
 * capacity of the PS buffer (plane stride).  Used by the single-agent PointPillar path
 * (/root/reference/opencood/models/point_pillar.py:52-84: no fusion stage between encoder level and deblock). */
int cb_ps_to_pf(const void* src_ps, int64_t src_lo_off, int n_cap, int n, int h, int w, int c,
                void* dst_pf, int64_t dst_lo_off, void* stream);
int cb_layout_to_nchw(const void* src, int64_t lo_off, int from_ps, int n, int c, int h, int w,
                      int pitch, int ch_off, float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COALIGN_B200_H */
