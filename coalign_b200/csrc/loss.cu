// PointPillarLoss on the GPU with the gradients w.r.t. the head outputs (first piece of SURVEY 8f row 2, the training
// step): focal classification loss, sin-difference smooth-L1 regression loss and the direction-bin cross entropy of
//   /root/reference/opencood/loss/point_pillar_loss.py:36-116 (forward), :119-131 (add_sin_difference),
//   :133-158 (get_direction_target), :201-245 (softmax_cross_entropy_with_logits, weighted_smooth_l1_loss,
//   sigmoid_focal_loss)
// The reference evaluates it with ~40 elementwise torch kernels and autograd; here one pass over the anchors produces the
// three loss sums and d(total)/d(cls_preds, reg_preds, dir_preds) in the NCHW layout of the head outputs.
// Arithmetic is float64 internally (the reference's label tensors are float64, which promotes its regression branch to
// float64; the float32 branches differ from this by float32 rounding only).  Reductions are two-stage in a fixed order:
// results are bit-reproducible from run to run.
#include <math.h>
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

struct LossGeom {
    int n, HW, A, num_bins, has_dir, labels_f64;
    double pos_cls_weight, alpha, gamma, cls_w, sigma, reg_w, dir_w, dir_offset;
    double anchor_yaw[8];
};

__device__ __forceinline__ double ld_label(const void* p, size_t i, int f64) {
    return f64 ? reinterpret_cast<const double*>(p)[i] : (double)reinterpret_cast<const float*>(p)[i];
}

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += s_red[w];                      // fixed order
    return t;
}

// positives per sample: pos_normalizer (point_pillar_loss.py:54).  grid = (chunks, samples); the counts are integers, so
// the fp64 atomics are exact in any order (pos_norm zeroed by the caller).
constexpr int POS_CHUNKS = 32;
__global__ void __launch_bounds__(256) loss_pos_count_kernel(const void* __restrict__ pos, const LossGeom g,
                                                             double* __restrict__ pos_norm) {
    __shared__ double s_red[8];
    const int b = blockIdx.y;
    const size_t per = (size_t)g.HW * g.A;
    double c = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (size_t)gridDim.x * blockDim.x)
        c += ld_label(pos, b * per + i, g.labels_f64) > 0.0 ? 1.0 : 0.0;
    c = block_sum(c, s_red);
    if (threadIdx.x == 0 && c != 0.0) atomicAdd(pos_norm + b, c);
}

// one thread per (sample, pixel, anchor); partial sums {cls, reg, dir} per CTA
__global__ void __launch_bounds__(256) loss_main_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                        const float* __restrict__ dir, const void* __restrict__ pos,
                                                        const void* __restrict__ neg, const void* __restrict__ tgt,
                                                        const LossGeom g, const double* __restrict__ pos_norm,
                                                        float* __restrict__ g_cls, float* __restrict__ g_reg,
                                                        float* __restrict__ g_dir, double* __restrict__ partial) {
    __shared__ double s_red[8];
    const size_t per = (size_t)g.HW * g.A;
    const size_t total = (size_t)g.n * per;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double l_cls = 0.0, l_reg = 0.0, l_dir = 0.0;
    if (e < total) {
        const int b = (int)(e / per);
        const size_t r = e - (size_t)b * per;
        const int hw = (int)(r / g.A), a = (int)(r - (size_t)hw * g.A);
        const double inv_n = 1.0 / (double)g.n;
        const double pn = fmax(pos_norm[b], 1.0);
        const double t = ld_label(pos, e, g.labels_f64);                       // cls_labls (0 / 1)
        const bool positive = t > 0.0;
        const bool negative = ld_label(neg, e, g.labels_f64) > 0.0;
        // ---- classification: sigmoid focal loss (:68-74, :230-245)
        {
            const size_t ci = ((size_t)b * g.A + a) * g.HW + hw;
            const double x = (double)cls[ci];
            const double w = ((positive ? g.pos_cls_weight : 0.0) + (negative ? 1.0 : 0.0)) / pn;
            const double ce = fmax(x, 0.0) - x * t + log1p(exp(-fabs(x)));
            const double p = 1.0 / (1.0 + exp(-x));
            const double pt = t * p + (1.0 - t) * (1.0 - p);
            const double om = 1.0 - pt;
            const double mod = pow(om, g.gamma);
            const double aw = t * g.alpha + (1.0 - t) * (1.0 - g.alpha);
            l_cls = mod * aw * ce * w;
            if (g_cls) {
                const double dpt = (2.0 * t - 1.0) * p * (1.0 - p);
                const double dmod = om > 0.0 ? -g.gamma * pow(om, g.gamma - 1.0) * dpt : 0.0;
                g_cls[ci] = (float)(w * aw * (dmod * ce + mod * (p - t)) * g.cls_w * inv_n);
            }
        }
        // ---- regression: smooth L1 on the sin-difference encoding (:76-82, :119-131, :219-227)
        const double rw = (positive ? 1.0 : 0.0) / pn;
        double tg[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) tg[k] = ld_label(tgt, e * 7 + k, g.labels_f64);
        {
            const double s2 = g.sigma * g.sigma, thr = 1.0 / s2;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const size_t ri = ((size_t)b * g.A * 7 + a * 7 + k) * g.HW + hw;
                const double x = (double)reg[ri];
                double diff, ddiff = 1.0;
                if (k == 6) {                                                  // sin(a) cos(b) - cos(a) sin(b), both depend on x
                    diff = sin(x) * cos(tg[6]) - cos(x) * sin(tg[6]);
                    ddiff = cos(x) * cos(tg[6]) + sin(x) * sin(tg[6]);
                } else {
                    diff = x - tg[k];
                }
                const double ad = fabs(diff);
                const bool small = ad <= thr;
                l_reg += (small ? 0.5 * (ad * g.sigma) * (ad * g.sigma) : ad - 0.5 / s2) * rw;
                if (g_reg) {
                    const double dl = small ? s2 * diff : (diff > 0.0 ? 1.0 : (diff < 0.0 ? -1.0 : 0.0));
                    g_reg[ri] = (float)(dl * ddiff * rw * g.reg_w * inv_n);
                }
            }
        }
        // ---- direction bins: cross entropy against the bin of (target yaw residual + anchor yaw) (:86-94, :133-158)
        if (g.has_dir) {
            const double two_pi = 2.0 * 3.141592653589793;
            const double rot_gt = tg[6] + g.anchor_yaw[a];
            const double v = rot_gt - g.dir_offset;
            const double off = v - floor(v / two_pi + 0.0) * two_pi;          // limit_period(v, 0, 2 pi)
            int bin = (int)floor(off / (two_pi / (double)g.num_bins));
            bin = bin < 0 ? 0 : (bin > g.num_bins - 1 ? g.num_bins - 1 : bin);
            double mx = -INFINITY;
            for (int k = 0; k < g.num_bins; ++k)
                mx = fmax(mx, (double)dir[((size_t)b * g.A * g.num_bins + a * g.num_bins + k) * g.HW + hw]);
            double se = 0.0;
            for (int k = 0; k < g.num_bins; ++k)
                se += exp((double)dir[((size_t)b * g.A * g.num_bins + a * g.num_bins + k) * g.HW + hw] - mx);
            const double lse = mx + log(se);
            const double xt = (double)dir[((size_t)b * g.A * g.num_bins + a * g.num_bins + bin) * g.HW + hw];
            l_dir = (lse - xt) * rw;
            if (g_dir) {
                for (int k = 0; k < g.num_bins; ++k) {
                    const size_t di = ((size_t)b * g.A * g.num_bins + a * g.num_bins + k) * g.HW + hw;
                    const double sm = exp((double)dir[di] - lse);
                    g_dir[di] = (float)((sm - (k == bin ? 1.0 : 0.0)) * rw * g.dir_w * inv_n);
                }
            }
        }
    }
    l_cls = block_sum(l_cls, s_red);
    l_reg = block_sum(l_reg, s_red);
    l_dir = block_sum(l_dir, s_red);
    if (threadIdx.x == 0) {
        partial[(size_t)blockIdx.x * 3 + 0] = l_cls;
        partial[(size_t)blockIdx.x * 3 + 1] = l_reg;
        partial[(size_t)blockIdx.x * 3 + 2] = l_dir;
    }
}

__global__ void __launch_bounds__(256) loss_final_kernel(const double* __restrict__ partial, int n_blocks, const LossGeom g,
                                                         float* __restrict__ out) {
    __shared__ double s_red[8];
    double s[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) {
        s[0] += partial[(size_t)i * 3 + 0]; s[1] += partial[(size_t)i * 3 + 1]; s[2] += partial[(size_t)i * 3 + 2];
    }
    const double c = block_sum(s[0], s_red), r = block_sum(s[1], s_red), d = block_sum(s[2], s_red);
    if (threadIdx.x == 0) {
        const double inv_n = 1.0 / (double)g.n;
        const double cl = c * g.cls_w * inv_n, rl = r * g.reg_w * inv_n, dl = g.has_dir ? d * g.dir_w * inv_n : 0.0;
        out[0] = (float)(cl + rl + dl);                                        // total_loss
        out[1] = (float)rl;                                                    // reg_loss
        out[2] = (float)cl;                                                    // cls_loss
        out[3] = (float)dl;                                                    // dir_loss
    }
}

}  // namespace cb

extern "C" size_t cb_pointpillar_loss_workspace_bytes(int n, int H, int W, int anchor_num) {
    if (n < 1 || H < 1 || W < 1 || anchor_num < 1) return 0;
    const size_t total = (size_t)n * H * W * anchor_num;
    return 256 + (size_t)n * sizeof(double) + ((total + 255) / 256) * 3 * sizeof(double);
}

extern "C" int cb_pointpillar_loss(const float* cls_preds, const float* reg_preds, const float* dir_preds,
                                   const void* pos_equal_one, const void* neg_equal_one, const void* targets,
                                   int labels_f64, int n, int H, int W, int anchor_num, int num_bins,
                                   float pos_cls_weight, float alpha, float gamma, float cls_weight, float sigma,
                                   float reg_weight, float dir_weight, float dir_offset, const double* anchor_yaw_rad,
                                   float* out_losses, float* grad_cls, float* grad_reg, float* grad_dir, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    using namespace cb;
    if (!cls_preds || !reg_preds || !pos_equal_one || !neg_equal_one || !targets || !out_losses || !workspace)
        return CB_ERR_ARG;
    if (n < 1 || H < 1 || W < 1 || anchor_num < 1 || anchor_num > 8) return CB_ERR_ARG;
    if (dir_preds && (num_bins < 1 || !anchor_yaw_rad)) return CB_ERR_ARG;
    if (grad_dir && !dir_preds) return CB_ERR_ARG;
    if (workspace_bytes < cb_pointpillar_loss_workspace_bytes(n, H, W, anchor_num) || ((uintptr_t)workspace & 7)) return CB_ERR_ARG;
    LossGeom g;
    g.n = n; g.HW = H * W; g.A = anchor_num; g.num_bins = num_bins; g.has_dir = dir_preds ? 1 : 0; g.labels_f64 = labels_f64 ? 1 : 0;
    g.pos_cls_weight = pos_cls_weight; g.alpha = alpha; g.gamma = gamma; g.cls_w = cls_weight; g.sigma = sigma;
    g.reg_w = reg_weight; g.dir_w = dir_weight; g.dir_offset = dir_offset;
    for (int i = 0; i < 8; ++i) g.anchor_yaw[i] = (dir_preds && i < anchor_num) ? anchor_yaw_rad[i] : 0.0;
    double* pos_norm = (double*)workspace;
    double* partial = (double*)((char*)workspace + (((size_t)n * sizeof(double) + 255) / 256) * 256);
    const size_t total = (size_t)n * H * W * anchor_num;
    const int n_blocks = (int)((total + 255) / 256);
    if ((size_t)((char*)(partial + (size_t)n_blocks * 3) - (char*)workspace) > workspace_bytes) return CB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t me = cudaMemsetAsync(pos_norm, 0, sizeof(double) * (size_t)n, st);
    if (me != cudaSuccess) return (int)me;
    loss_pos_count_kernel<<<dim3(POS_CHUNKS, n), 256, 0, st>>>(pos_equal_one, g, pos_norm);
    CB_CHECK_LAUNCH();
    loss_main_kernel<<<n_blocks, 256, 0, st>>>(cls_preds, reg_preds, dir_preds, pos_equal_one, neg_equal_one, targets, g,
                                               pos_norm, grad_cls, grad_reg, grad_dir, partial);
    CB_CHECK_LAUNCH();
    loss_final_kernel<<<1, 256, 0, st>>>(partial, n_blocks, g, out_losses);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
