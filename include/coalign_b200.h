/* libcoalign_b200 - C ABI of the B200-native CoAlign hot path.
 *
 * This is the drop-in boundary below the reference's Python model
 * (/root/reference/opencood/models/point_pillar_baseline_multiscale.py:93-135): every entry point
 * replaces one stage of that forward (citations on each function).  Conventions:
 *   - plain C: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - returns 0 on success, a positive cudaError_t, or a negative CB_ERR_* argument/driver error;
 *   - never allocates and never synchronises: the caller owns inputs, outputs and workspaces,
 *     all work is enqueued on `stream` (CUDA-graph capturable);
 *   - no hidden global state: the lazily resolved cuTensorMapEncodeTiled entry point, per-kernel shared-memory
 *     attributes (idempotent, thread-safe) and the explicit option table below (cb_set_option; no environment variables
 *     are read anywhere in the library).
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

#include <stddef.h>
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
/* Process-wide kernel-selection options (development / validation switches; defaults = the measured-best product path).
 * Explicit calls replace the environment variables of round 1.  Set them before launching; they are read at launch time
 * (atomic loads), so changing one between launches is well defined, changing one DURING stream capture is not. */
enum {
    CB_OPT_NO_PDL = 0,        /* 1: launch without programmatic dependent launch                         (default 0) */
    CB_OPT_EPI_DIRECT = 1,    /* conv epilogue: 1 = two 256-bit stores per thread where alignment allows (default 1) */
    CB_OPT_TMA_STORE = 2,     /* conv epilogue: 1 = TMA store for PF outputs (measured slower on residual layers; default 0) */
    CB_OPT_HALO_BO = 3,       /* halo kernels: UMMA descriptor base-offset mode experiment               (default 0) */
    CB_OPT_FUSE_VERSION = 4,  /* fusion kernel: 9 = tiled v9 (default), 8 = v8, 1 = first version          */
    CB_OPT_FUSE_BLEND_FP32 = 5, /* 1: fp32 tap blend in bf16 mode instead of packed bf16 FMAs             (default 0) */
    CB_OPT_FUSE_OCC3 = 6,     /* 1: 3 CTAs/SM variant of v9 instead of 4                                  (default 0) */
    CB_OPT_CONV_DEBUG = 7,    /* conv kernel experiment flags (profiles/r1_conv_analysis.md)              (default 0) */
    CB_OPT_COUNT = 8
};
int cb_set_option(int option, int value);     /* returns the previous value, or CB error (< 0) for an unknown option */
int cb_get_option(int option);
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
    /* halo width of the GEMM row space (0 = 1): interior = in_pad..Hp-1-in_pad x in_pad..Wp-1-in_pad.  2 for the 7x7/s2 stem
     * of the camera BEV encoder (lss_submodule.py:371-372), whose taps reach two pixels into the parity planes; a CB_OUT_PF
     * destination with other padded dims (out_Hp/out_Wp != Hp/Wp) is then addressed by pixel, not by GEMM row. */
    int32_t in_pad;
} cb_conv_desc;

/* tcgen05 path (the product).  max_ctas <= 0: one persistent CTA per SM. */
int cb_conv_gemm(const cb_conv_desc* desc, int max_ctas, void* stream);
/* Same GEMM on CTA pairs (tcgen05 cta_group::2, 256 x block_n tiles, cluster of 2): halves the weight-tile
 * shared-memory traffic per MAC.  max_clusters <= 0: one cluster per SM pair.  block_n 32 falls back to cb_conv_gemm. */
int cb_conv_gemm_pair(const cb_conv_desc* desc, int max_clusters, void* stream);
/* Channel-major variant for n_total == cout_mod == 128 or 256, bf16 PF/PS outputs: the weight tile is the UMMA A operand
 * (M = 128 output channels) and 256 pixels the N side, so one instruction does 128 x 256 x 16 MACs and the activation
 * rows are read from shared memory once per 256-wide tile.  256 output channels run as two 128-channel work items per
 * pixel tile on neighbouring CTAs (the second read of the activation box hits L2).  Same descriptor; block_n is ignored.  Returns
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
 * Training step, device side: autograd of the forward above, train-mode BatchNorm, Adam.
 * Reference: the loop body /root/reference/opencood/tools/train.py:105-125 (model.train(); forward; criterion;
 * backward; optimizer.step) with torch.optim.Adam from train_utils.setup_optimizer (train_utils.py:196-206) and
 * nn.BatchNorm2d / BatchNorm1d in training mode (batch statistics, running-stat update with the layer's momentum:
 * resblock.py:38-39, base_bev_backbone_resnet.py:62-63, pillar_vfe.py:25,41-44).
 *
 * Division of labour:
 *   - forward convolutions run through cb_conv_gemm* with RAW (unfolded) weights, bias 0, no ReLU -> z (bf16 PF);
 *   - cb_bn_stats / cb_bn_finalize / cb_bn_apply turn z into y = relu(bn(z) [+ bn_b(z_b)] [+ residual]) in the layout
 *     the next consumer wants (PF, PS, or pixel-shuffled into the concat buffer);
 *   - backward: cb_bn_bwd_reduce + cb_bn_bwd_apply give dz from dy (ReLU mask from y, batch-norm adjoint), the input
 *     gradient is the FORWARD kernel on dz with transposed/flipped packed weights (cb_conv_gemm*), the weight gradient is
 *     cb_wgrad (MN-major tcgen05 GEMM over the pixel index, split-K);
 *   - cb_warp_att_fuse_bwd is the adjoint of cb_warp_att_fuse (soft-max backward over the agents + scatter-add of the
 *     four bilinear taps), cb_pfn_* the train-mode PFN (statistics over all M*32 slots, arg-max routing);
 *   - cb_adam_step is torch.optim.Adam (L2 weight decay folded into the gradient) on flat fp32 buffers.
 * All of them: raw device pointers, caller-owned workspaces, enqueue-only on `stream`, CUDA-graph capturable.
 * ------------------------------------------------------------------------------------------ */

/* ---- weight gradient: dW[m][..] += sum_q dZ[q + a_row_off][m0 + m] * X[q + row_off][col + c]  (see csrc/wgrad.cu) */
#define CB_WGRAD_MAX_BOXES 4
#define CB_WGRAD_MAX_UNITS 28
typedef struct cb_wgrad_box {
    int32_t row_off;      /* row shift (tap) + plane offset of the X operand, rows of the X tensor */
    int32_t out_ld;       /* floats between consecutive m (dZ channels) in dw */
    int64_t out_off;      /* float offset in dw of (m = 0, c = 0) of this 64-channel block */
    uint16_t col;         /* first X channel of the block */
    uint16_t x_sel;       /* which X tensor (0/1) */
    uint32_t pad_;
} cb_wgrad_box;
typedef struct cb_wgrad_unit {
    int32_t m0;           /* first dZ channel (128 per unit; channels past dz_pitch read as zero) */
    int32_t a_row_off;    /* row shift of the dZ operand */
    int32_t m_valid;      /* rows of the 128 that are stored (<= 128) */
    int32_t n_boxes;      /* 1..CB_WGRAD_MAX_BOXES 64-channel X blocks (GEMM N = 64 * n_boxes) */
    cb_wgrad_box box[CB_WGRAD_MAX_BOXES];
} cb_wgrad_unit;
typedef struct cb_wgrad_desc {
    const void* dz_ptr;        /* bf16 [rows_total][dz_pitch] PF row space of the conv output; halo rows zero */
    const void* dz_lo_ptr;     /* precise mode: lo plane of dZ (same shape) or NULL */
    int32_t dz_pitch;
    int32_t k_splits;          /* <= 0: chosen so that units * splits fills the SMs once */
    int64_t rows_total;        /* n_img * Hp * Wp */
    const void* x_ptr[2];      /* bf16 [x_rows][x_pitch] (PF, PS planes, ...); rows outside read as zero */
    int64_t x_rows[2];
    int32_t x_pitch[2];
    int32_t x_lo_rows[2];      /* precise mode: row offset of the lo plane inside x (counted in x_rows) */
    float* dw;                 /* fp32, accumulated into (zero it first) */
    int32_t n_units;
    int32_t pad_;
    cb_wgrad_unit units[CB_WGRAD_MAX_UNITS];
} cb_wgrad_desc;
int cb_wgrad(const cb_wgrad_desc* desc, int max_ctas, void* stream);
/* SIMT evaluation of the same descriptor (validation of the tensor-core kernel; tests only). */
int cb_wgrad_simt(const cb_wgrad_desc* desc, void* stream);

/* ---- train-mode BatchNorm around the conv GEMMs.
 * Row mapping ("y side") shared by cb_bn_apply / cb_bn_bwd_*: z is always PF [n_img][Hp][Wp][c_total] in the GEMM's row
 * space; the activation y (and its gradient dy) lives in
 *   CB_OUT_PF       the same rows, channels [y_ch_off, y_ch_off + c_total)
 *   CB_OUT_PS       4 parity planes (y_Hp, y_Wp, y_plane_rows as in cb_conv_desc)
 *   CB_OUT_UPSAMPLE pixel-shuffled: column block ab = col / c_mod goes to pixel (k*h + a, k*w + b), channel col % c_mod
 * Channel statistics are indexed by col % c_mod (c_total = k*k*c_mod for the transposed convolutions). */
typedef struct cb_map {
    int32_t n_img, Hp, Wp;       /* z row space: rows_total = n_img*Hp*Wp, interior 1..Hp-2 x 1..Wp-2 */
    int32_t c_total, c_mod;
    int32_t y_mode;              /* CB_OUT_PF / CB_OUT_PS / CB_OUT_UPSAMPLE */
    int32_t y_pitch, y_ch_off, up_k, y_Hp, y_Wp;
    int32_t z_at_y;              /* backward kernels only: 1 = the saved forward z is stored at the y position (pixel-shuffled
                                    PF tensor of c_mod channels, pitch z_pitch) instead of in the GEMM's row space */
    int32_t z_pitch;
    int64_t y_plane_rows;
} cb_map;

/* sums[0..c_mod) += sum z, sums[c_mod..2c_mod) += sum z*z over the interior pixels (fp64). */
int cb_bn_stats(const void* z, int64_t z_lo_off, const cb_map* map, double* sums, void* stream);
/* Batch statistics -> per-channel affine (scale = gamma*inv_std, shift = beta - mean*scale), saved mean / inv_std for the
 * backward, running-stat update: running = (1-momentum)*running + momentum*(mean | unbiased var).  count = elements per
 * channel.  gamma == NULL: identity (scale 1, shift = bias or 0). */
int cb_bn_finalize(const double* sums, int c, double count, float eps, float momentum,
                   const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float* scale, float* shift, float* mean, float* inv_std, void* stream);
/* cb_bn_stats + cb_bn_finalize in ONE launch: the CTA that draws the last ticket from `counter` (device int32, zero on
 * entry, reset to zero by the kernel) closes the statistics.  c = map->c_mod; gamma must be non-NULL. */
int cb_bn_stats_finalize(const void* z, int64_t z_lo_off, const cb_map* map, double* sums, int32_t* counter, double count,
                         float eps, float momentum, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float* scale, float* shift, float* mean, float* inv_std, void* stream);
/* y = [relu]( z*scale + shift [+ z_b*scale_b + shift_b] [+ residual] ) written through `map` (bf16, + lo plane when
 * y_lo_off != 0).  residual: bf16 PF in z's row space, pitch res_pitch. */
int cb_bn_apply(const void* z, int64_t z_lo_off, const float* scale, const float* shift,
                const void* z_b, int64_t z_b_lo_off, const float* scale_b, const float* shift_b,
                const void* residual, int32_t res_pitch, int64_t res_lo_off, int relu,
                const cb_map* map, void* y, int64_t y_lo_off, void* stream);
/* dyr = dy * (y > 0 if relu);  sums[0..c) += sum dyr, sums[c..2c) += sum dyr * x_hat, x_hat = (z - mean) * inv_std
 * (mean == NULL: no BatchNorm, only the first half = bias gradient).  mask_scale / mask_shift non-NULL (plain
 * conv + BN + ReLU, i.e. y = relu(z*scale + shift) with nothing added): the ReLU mask is recomputed from z instead of
 * reading y - one tensor read less in each of the two backward passes. */
int cb_bn_bwd_reduce(const void* dy, int64_t dy_lo_off, const void* y, int64_t y_lo_off, int relu,
                     const void* z, int64_t z_lo_off, const float* mean, const float* inv_std,
                     const float* mask_scale, const float* mask_shift,
                     const cb_map* map, double* sums, void* stream);
/* dz = gamma*inv_std * (dyr - sums[c]/count - x_hat * sums[c_mod + c]/count)  (no BatchNorm: dz = dyr), bf16 PF in z's
 * row space, interior rows only (halo rows stay zero).  Also writes d_gamma / d_beta (fp32 [c_mod]) when non-NULL, and
 * dsum_pf = dyr as bf16 PF (pitch c_total) when non-NULL (the identity branch of a residual block). */
int cb_bn_bwd_apply(const void* dy, int64_t dy_lo_off, const void* y, int64_t y_lo_off, int relu,
                    const void* z, int64_t z_lo_off, const float* mean, const float* inv_std, const float* gamma,
                    const float* mask_scale, const float* mask_shift,
                    const double* sums, double count, const cb_map* map,
                    void* dz, int64_t dz_lo_off, void* dsum_pf, int64_t dsum_lo_off,
                    float* d_gamma, float* d_beta, void* stream);

/* ---- heads: d(loss)/d(cls, reg, dir) fp32 NCHW (from cb_pointpillar_loss) -> bf16 PF [n][Hp][Wp][64] (zero padded
 * columns) for the dgrad / wgrad GEMMs, and the bias gradients d_bias[c] = sum over pixels (fp32, [sum of head_cn]). */
int cb_heads_grad_pack(const float* const* grads, const int32_t* head_cn, int n_heads, int n, int H, int W,
                       void* g_pf, int64_t g_lo_off, float* d_bias, void* stream);

/* ---- adjoint of cb_warp_att_fuse: d_feat (fp32, dense [sum_agents][H][W][C], ACCUMULATED with vector reductions; zero it
 * first) from d_fused (bf16 PF [n_scenes][H+2][W+2][C]) - soft-max backward over the agents (att_fuse.py:44-46), gradient of
 * the ego query, scatter-add through the four bilinear taps (torch_transformation_utils.py:322-331); method 1 = MaxFusion
 * (gradient to the first arg-max agent, fusion_in_one.py:83-86). */
int cb_warp_att_fuse_bwd(const void* feat, int in_ps, int64_t in_lo_off, int sum_agents,
                         const double* affine, const int32_t* agent_off, int n_scenes, int max_cav,
                         int H, int W, int C, int method,
                         const void* d_fused_pf, int64_t d_fused_lo_off, float* d_feat, void* stream);
/* out(layout) = bf16( acc_fp32 [n][H][W][C] + addend(layout, optional) ): closes the gradient of a level map (fusion
 * branch + the next level's first block) in the layout its producer's backward reads (PF or PS, + lo plane). */
int cb_grad_combine(const float* acc, const void* addend, int64_t addend_lo_off, int to_ps, int n_cap, int n, int H, int W,
                    int C, void* out, int64_t out_lo_off, void* stream);

/* ---- PFN in training mode, reference-format voxel tensors (what train.py's dataloader delivers).
 * cb_pfn_train_stats: sums[0..10) = sum of the 10 augmented features over all valid points, sums[10..110) = their 10x10
 * Gram matrix (fp64, accumulated).  With them mean / E[x^2] of linear(f) over ALL M*max_pts slots (padded slots are zero
 * rows, pillar_vfe.py:41-44) follow in closed form: cb_pfn_train_finalize writes scale / shift (+ mean, inv_std, running
 * stats with momentum 0.01), after which cb_pfn_scatter runs the forward unchanged.
 * cb_pfn_bwd: arg-max routing of d_canvas (PS layout) to the slot that won the max, ReLU mask, and the reductions
 * bsum[0..64) = d_beta, [64..128) = d_gamma, [128..768) = sum dy * f (fp64); cb_pfn_bwd_finalize turns them into the
 * gradients of linear.weight (64x10), norm.weight, norm.bias. */
int cb_pfn_train_stats(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows_cap,
                       const int32_t* n_voxels_dev, int max_pts, const float* vsize, const float* center_off,
                       double* sums, void* stream);
int cb_pfn_train_finalize(const double* sums, const int32_t* n_voxels_dev, int n_rows_cap, int max_pts,
                          const float* w, const float* gamma, const float* beta, float eps, float momentum,
                          float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* inv_std,
                          void* stream);
int cb_pfn_bwd(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows_cap,
               const int32_t* n_voxels_dev, int max_pts, const float* w, const float* scale, const float* shift,
               const float* mean, const float* inv_std, const float* vsize, const float* center_off,
               const void* d_canvas_ps, int64_t d_canvas_lo_off, int canvas_agents, int ny, int nx,
               double* bsum, void* stream);
int cb_pfn_bwd_finalize(const double* stats_sums, const double* bsum, const int32_t* n_voxels_dev, int n_rows_cap,
                        int max_pts, const float* w, const float* gamma, const float* mean, const float* inv_std,
                        float* d_w, float* d_gamma, float* d_beta, void* stream);

/* ---- layout permutations of parameters / gradients: dst[(r1,r0)][(k1,k0)] = src[r1*s_r1 + r0*s_r0 + k1*s_k1 + k0*s_k0].
 * cb_pack_weight: fp32 -> bf16 [R1*R0][dst_ld] at column k_off (hi; + lo part at column k_off + lo_col_off when non-zero).
 * cb_permute_f32: fp32 -> fp32 [R1*R0][K1*K0] (gradient back to the parameter's own layout), scaled by `alpha`. */
int cb_pack_weight(const float* src, int R1, int R0, int K1, int K0, int64_t s_r1, int64_t s_r0, int64_t s_k1, int64_t s_k0,
                   void* dst, int dst_ld, int k_off, int lo_col_off, void* stream);
int cb_permute_f32(const float* src, int R1, int R0, int K1, int K0, int64_t s_r1, int64_t s_r0, int64_t s_k1, int64_t s_k0,
                   float alpha, float* dst, void* stream);
/* All cb_pack_weight jobs of a model in ONE launch (a training step re-packs ~90 weight matrices; as separate launches they
 * cost more than the 26 M elements they move).  jobs: DEVICE array; `first` = running total of rows*K of the jobs before. */
typedef struct cb_pack_job {
    const float* src;
    void* dst;                   /* bf16, already offset to the job's first row */
    int64_t s_r1, s_r0, s_k1, s_k0, first;
    int32_t R0, K0, rows, K, dst_ld, k_off, lo_col_off, pad_;   /* lo_col_off < 0: dst is fp32 (a cb_permute_f32 job, alpha 1) */
} cb_pack_job;
int cb_pack_weights_batch(const cb_pack_job* jobs_dev, int n_jobs, int64_t total_elems, void* stream);

/* ---- torch.optim.Adam step on flat fp32 buffers (train_utils.py:196-206: lr, eps, weight_decay from the yaml):
 *   g += wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * `step` is the 1-based step count in DEVICE memory (int32, incremented by the kernel's caller via cb_adam_step's
 * `inc_step`), so that a captured graph can be replayed every iteration; grad_scale multiplies g first (1/world). */
int cb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float grad_scale, int32_t* step_dev, int inc_step, void* stream);

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
/* dense NCHW float32 (n,c,h,w) -> PS layout with a `pad`-pixel halo per parity plane (plane dims ceil(h/2)+2*pad,
 * ceil(w/2)+2*pad; n_cap = plane stride in images): the input layout of the camera BEV encoder's 7x7/s2 stem. */
int cb_nchw_to_ps_pad(const float* src, int n, int c, int h, int w, int pad, int n_cap, void* dst, int64_t lo_off,
                      void* stream);
/* fp32 channels-last (n, h, w, c) -> the same PS / pad layout (the BEV accumulator of cb_lift_splat feeds the stem). */
int cb_nhwc_to_ps_pad(const float* src, int n, int c, int h, int w, int pad, int n_cap, void* dst, int64_t lo_off,
                      void* stream);
/* Lift + splat of the camera model: depth soft-max (x) image features, summed into the BEV voxels of their frustum points.
 *   CamEncode.get_depth_dist + the outer product of CamEncode.forward   lss_submodule.py:60-61,134-136
 *   LiftSplatShoot.get_geometry / voxel_pooling                         lift_splat_shoot.py:80-105,115-169
 * depth_logit DEVICE f32 [B*N][D][fH][fW]; feat DEVICE f32 [B*N][C][fH][fW]; cam_mats DEVICE f32 [B*N][24] = inverse(post_rots)
 * (9, row major), post_trans (3), rots * inverse(intrins) (9), trans (3); xs [fW], ys [fH], ds [D] DEVICE f32 = the frustum's
 * image-plane coordinates and depth bins (create_frustum, :64-78); dx, bx HOST f32 [3], nx HOST i32 [3] from gen_dx_bx.
 * acc: DEVICE f32 [B][ny][nx][nz*C] channels-last, ACCUMULATED (zero it first); equals voxel_pooling's `final` with the z
 * planes concatenated along the channels (:163-165), i.e. final[b, z*C + c, y, x] = acc[b][y][x][z*C + c].
 * The lifted (B*N, C, D, fH, fW) tensor, the sort by voxel rank and the cumsum trick are never materialised. */
int cb_lift_splat(const float* depth_logit, const float* feat, const float* cam_mats, const float* xs, const float* ys,
                  const float* ds, int B, int N, int D, int fH, int fW, int C, const float* dx, const float* bx,
                  const int32_t* nx, float* acc, void* stream);
/* Bilinear up-sampling by `scale` (1 = plain copy, 2 = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
 * lss_submodule.py:23-24,36-37) of a PF map (n, h, w, c) into channels [dst_ch_off, dst_ch_off + c) of a PF map
 * (n, scale*h, scale*w, dst_pitch): Up.forward's upsample + torch.cat without materialising either. */
int cb_upsample_concat(const void* src_pf, int64_t src_lo_off, int n, int h, int w, int c, int scale,
                       void* dst_pf, int64_t dst_lo_off, int dst_pitch, int dst_ch_off, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COALIGN_B200_H */
