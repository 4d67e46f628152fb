// Detection post-processing on the GPU (the step after the CoAlign forward): anchor-box decoding,
// direction fix, corner projection, size / z filters, top-k by score, rotated (polygon) NMS and the range mask, without a
// host round trip.  Replaces, for intermediate fusion (data_dict = {'ego': ...}),
//   VoxelPostprocessor.post_process / delta_to_boxes3d  /root/reference/opencood/data_utils/post_processor/voxel_postprocessor.py:243-449
//   boxes_to_corners_3d / project_box3d / remove_large_pred_bbx / remove_bbx_abnormal_z / nms_rotated /
//   mask_boxes_outside_range_numpy                      /root/reference/opencood/utils/box_utils.py:152-204,278-316,384-421,693-738,840-890
//   limit_period / rotate_points_along_z / compute_iou  /root/reference/opencood/utils/common_utils.py:70-79,105-127,196-218
// (the reference does sigmoid/threshold on the device, then a Python loop over shapely polygons on the host).
// Five small kernels per call, all on the caller's stream:
//   1 decode + threshold + filters -> 64-bit sort keys (score bits | inverted anchor index) of the surviving candidates
//   2 per scene: radix-select the top_k keys, bitonic sort in shared memory (score descending, anchor index ascending)
//   3 corners of the selected boxes (recomputed from the anchor index; candidates never store 96-byte corner records)
//   4 pairwise polygon IoU > threshold -> suppression bit matrix (float64 Sutherland-Hodgman clipping)
//   5 greedy scan over the bit matrix (one warp, matrix in shared memory), range mask, ordered output
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

constexpr int PP_MAX_TOPK = 1024;
constexpr int PP_WORDS = PP_MAX_TOPK / 64;

struct PostGeom {
    int H, W, A, num_bins, has_dir, order_hwl, top_k, cap;
    int stage1;                    // UncertaintyVoxelPostprocessor.post_process_stage1: no projection, no filters, no range mask
    float score_thr, dir_offset, period, two_pi, nms_thr;
    double range[6];
};

struct PostWs {
    unsigned long long* keys;      // [n_scenes][cap]
    int* counters;                 // [n_scenes][4]: above threshold, candidates (after filters), sorted, -
    int* sorted_idx;               // [n_scenes][PP_MAX_TOPK]
    float* sorted_score;           // [n_scenes][PP_MAX_TOPK]
    float* boxes;                  // [n_scenes][PP_MAX_TOPK][24]
    unsigned char* in_range;       // [n_scenes][PP_MAX_TOPK]
    float* box7;                   // [n_scenes][PP_MAX_TOPK][7]  decoded boxes of the selected anchors (stage-1 output)
    unsigned long long* mask;      // [n_scenes][PP_MAX_TOPK][PP_WORDS]
};

// Box of anchor `idx` (flat (h, w, a) order of the reference's permute(0,2,3,1).reshape): score, projected corners,
// filter verdicts.  float32 arithmetic in the reference's operation order, no FMA contraction where it has separate ops.
__device__ __forceinline__ bool decode_box(const PostGeom& g, const float* __restrict__ cls, const float* __restrict__ reg,
                                           const float* __restrict__ dir, const float* __restrict__ anchors,
                                           const float* __restrict__ T, int idx, float& score, float (&c)[8][3],
                                           bool& keep, bool& in_range, float* __restrict__ box7 = nullptr) {
    const int a = idx % g.A;
    const int hw = idx / g.A;
    const int HW = g.H * g.W;
    const float logit = cls[a * HW + hw];
    score = 1.0f / (1.0f + expf(-logit));                                  // torch.sigmoid
    if (!(score > g.score_thr)) return false;
    float d[7], an[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        d[k] = reg[(a * 7 + k) * HW + hw];
        an[k] = anchors[(size_t)idx * 7 + k];
    }
    // delta_to_boxes3d (voxel_postprocessor.py:430-447)
    const float ad = __fsqrt_rn(__fadd_rn(__fmul_rn(an[4], an[4]), __fmul_rn(an[5], an[5])));
    float b[7];
    b[0] = __fadd_rn(__fmul_rn(d[0], ad), an[0]);
    b[1] = __fadd_rn(__fmul_rn(d[1], ad), an[1]);
    b[2] = __fadd_rn(__fmul_rn(d[2], an[3]), an[2]);
    b[3] = __fmul_rn(expf(d[3]), an[3]);
    b[4] = __fmul_rn(expf(d[4]), an[4]);
    b[5] = __fmul_rn(expf(d[5]), an[5]);
    b[6] = __fadd_rn(d[6], an[6]);
    if (g.has_dir) {                                                       // voxel_postprocessor.py:325-339
        int label = 0;
        float best = dir[(a * g.num_bins) * HW + hw];
        for (int k = 1; k < g.num_bins; ++k) {
            const float v = dir[(a * g.num_bins + k) * HW + hw];
            if (v > best) { best = v; label = k; }
        }
        const float v = __fsub_rn(b[6], g.dir_offset);
        const float rot = __fsub_rn(v, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(v, g.period), 0.0f)), g.period));
        const float y2 = __fadd_rn(__fadd_rn(rot, g.dir_offset), __fmul_rn(g.period, (float)label));
        b[6] = __fsub_rn(y2, __fmul_rn(floorf(__fadd_rn(__fdiv_rn(y2, g.two_pi), 0.5f)), g.two_pi));
    }
    if (box7) {
#pragma unroll
        for (int k = 0; k < 7; ++k) box7[k] = b[k];
    }
    // boxes_to_corners_3d (box_utils.py:186-204): 'hwl' boxes are [x,y,z,h,w,l,yaw]
    const float L = g.order_hwl ? b[5] : b[3], Wd = b[4], Hh = g.order_hwl ? b[3] : b[5];
    const float cs = cosf(b[6]), sn = sinf(b[6]);
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, zmin = INFINITY, zmax = -INFINITY;
    in_range = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float tx = (k == 0 || k == 1 || k == 4 || k == 5) ? 0.5f : -0.5f;
        const float ty = (k == 1 || k == 2 || k == 5 || k == 6) ? 0.5f : -0.5f;
        const float tz = k >= 4 ? 0.5f : -0.5f;
        const float lx = __fmul_rn(L, tx), ly = __fmul_rn(Wd, ty), lz = __fmul_rn(Hh, tz);
        // rotate_points_along_z: [x y z] @ [[c, s, 0], [-s, c, 0], [0, 0, 1]]
        const float rx = __fadd_rn(fmaf(ly, -sn, __fmul_rn(lx, cs)), 0.0f);
        const float ry = __fadd_rn(fmaf(ly, cs, __fmul_rn(lx, sn)), 0.0f);
        const float px = __fadd_rn(rx, b[0]), py = __fadd_rn(ry, b[1]), pz = __fadd_rn(lz, b[2]);
        // project_box3d: T @ [x y z 1]^T  (stage 1 keeps the corners in the agent's own frame,
        // uncertainty_voxel_postprocessor.py:80-82)
        float X = px, Y = py, Z = pz;
        if (!g.stage1) {
            X = fmaf(T[3], 1.0f, fmaf(T[2], pz, fmaf(T[1], py, __fmul_rn(T[0], px))));
            Y = fmaf(T[7], 1.0f, fmaf(T[6], pz, fmaf(T[5], py, __fmul_rn(T[4], px))));
            Z = fmaf(T[11], 1.0f, fmaf(T[10], pz, fmaf(T[9], py, __fmul_rn(T[8], px))));
        }
        c[k][0] = X; c[k][1] = Y; c[k][2] = Z;
        xmin = fminf(xmin, X); xmax = fmaxf(xmax, X);
        ymin = fminf(ymin, Y); ymax = fmaxf(ymax, Y);
        zmin = fminf(zmin, Z); zmax = fmaxf(zmax, Z);
        in_range = in_range && (double)X >= g.range[0] && (double)Y >= g.range[1] && (double)Z >= g.range[2] &&
                   (double)X <= g.range[3] && (double)Y <= g.range[4] && (double)Z <= g.range[5];
    }
    // remove_large_pred_bbx (box_utils.py:855-869; its z_len is the y extent, tested for truthiness only) and
    // remove_bbx_abnormal_z (:886-888)
    const float x_len = __fsub_rn(xmax, xmin), y_len = __fsub_rn(ymax, ymin);
    keep = (x_len <= 6.0f) && (y_len <= 6.0f) && (y_len != 0.0f) && (zmin >= -3.0f) && (zmax <= 1.0f);
    if (g.stage1) { keep = true; in_range = true; }                       // stage 1 applies neither filter nor range mask
    return true;
}

__global__ void __launch_bounds__(256) pp_decode_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                        const float* __restrict__ dir, const float* __restrict__ anchors,
                                                        const float* __restrict__ tfm, const PostGeom g, PostWs ws) {
    const int b = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t HW = (size_t)g.H * g.W;
    bool above = false, keep = false, inr = false;
    float score = 0.f;
    if (idx < g.cap) {
        float c[8][3];
        above = decode_box(g, cls + (size_t)b * g.A * HW, reg + (size_t)b * g.A * 7 * HW,
                           g.has_dir ? dir + (size_t)b * g.A * g.num_bins * HW : nullptr, anchors,
                           g.stage1 ? nullptr : tfm + b * 16, idx, score, c, keep, inr);
        keep = above && keep;
    }
    // warp-aggregated counters
    const unsigned lane = threadIdx.x & 31;
    const unsigned m_above = __ballot_sync(0xffffffffu, above), m_keep = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0) {
        if (m_above) atomicAdd(&ws.counters[b * 4 + 0], __popc(m_above));
        if (m_keep) base = atomicAdd(&ws.counters[b * 4 + 1], __popc(m_keep));
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const int slot = base + __popc(m_keep & ((1u << lane) - 1u));
        // larger key = better: score bits (positive floats order like integers), then LOWER anchor index
        ws.keys[(size_t)b * g.cap + slot] = ((unsigned long long)__float_as_uint(score) << 32) |
                                            (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
    }
}

// One CTA per scene: exact top-k of the candidate keys (MSB-first radix select, 8 bits per pass, early exit once the
// boundary bin is taken whole), then an in-shared-memory bitonic sort (descending).
__global__ void __launch_bounds__(1024) pp_select_sort_kernel(const PostGeom g, PostWs ws) {
    __shared__ unsigned long long skeys[PP_MAX_TOPK];
    __shared__ int hist[256];
    __shared__ int s_sel_bin, s_remaining, s_done, s_count;
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned long long* keys = ws.keys + (size_t)b * g.cap;
    int n = ws.counters[b * 4 + 1];
    n = n < g.cap ? n : g.cap;
    const int K = n < g.top_k ? n : g.top_k;
    unsigned long long prefix = 0;            // selected high bytes so far
    int shift = 64;                           // keys with (key >> shift) > prefix are in; == prefix undecided
    if (tid == 0) { s_remaining = K; s_done = (n <= g.top_k) ? 1 : 0; s_count = 0; }
    __syncthreads();
    if (!s_done) {
        for (int pass = 7; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const int sh = pass * 8;
            for (int i = tid; i < n; i += blockDim.x) {
                const unsigned long long k = keys[i];
                if (shift == 64 || (k >> shift) == prefix) atomicAdd(&hist[(int)((k >> sh) & 255ull)], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int rem = s_remaining, cum = 0, d = 255;
                for (; d >= 0; --d) {
                    if (cum + hist[d] >= rem) break;
                    cum += hist[d];
                }
                s_sel_bin = d;
                s_remaining = rem - cum;                      // still to take from bin d
                s_done = (hist[d] == rem - cum) ? 1 : 0;      // bin d is taken whole: boundary found
            }
            __syncthreads();
            prefix = (prefix << 8) | (unsigned long long)s_sel_bin;
            shift = sh;
            if (s_done) break;
            __syncthreads();
        }
    }
    // selected: n <= top_k -> everything; else (key >> shift) >= prefix (exactly K keys: keys are unique)
    for (int i = tid; i < PP_MAX_TOPK; i += blockDim.x) skeys[i] = 0ull;
    __syncthreads();
    const bool all = n <= g.top_k;
    for (int i = tid; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        if (all || (k >> shift) >= prefix) {
            const int pos = atomicAdd(&s_count, 1);
            if (pos < PP_MAX_TOPK) skeys[pos] = k;
        }
    }
    __syncthreads();
    // bitonic sort, descending, PP_MAX_TOPK elements, one element per thread
    for (int k = 2; k <= PP_MAX_TOPK; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int ixj = tid ^ j;
            if (ixj > tid) {
                const unsigned long long x = skeys[tid], y = skeys[ixj];
                const bool desc = (tid & k) == 0;
                if (desc ? (x < y) : (x > y)) { skeys[tid] = y; skeys[ixj] = x; }
            }
            __syncthreads();
        }
    }
    if (tid < K) {
        const unsigned long long k = skeys[tid];
        ws.sorted_idx[b * PP_MAX_TOPK + tid] = (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull));
        ws.sorted_score[b * PP_MAX_TOPK + tid] = __uint_as_float((unsigned)(k >> 32));
    }
    if (tid == 0) ws.counters[b * 4 + 2] = K;
}

__global__ void __launch_bounds__(128) pp_corners_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                         const float* __restrict__ dir, const float* __restrict__ anchors,
                                                         const float* __restrict__ tfm, const PostGeom g, PostWs ws) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ws.counters[b * 4 + 2]) return;
    const size_t HW = (size_t)g.H * g.W;
    float c[8][3], score;
    bool keep, inr;
    decode_box(g, cls + (size_t)b * g.A * HW, reg + (size_t)b * g.A * 7 * HW,
               g.has_dir ? dir + (size_t)b * g.A * g.num_bins * HW : nullptr, anchors, g.stage1 ? nullptr : tfm + b * 16,
               ws.sorted_idx[b * PP_MAX_TOPK + i], score, c, keep, inr, ws.box7 + ((size_t)b * PP_MAX_TOPK + i) * 7);
    float* o = ws.boxes + ((size_t)b * PP_MAX_TOPK + i) * 24;
#pragma unroll
    for (int k = 0; k < 8; ++k) { o[3 * k] = c[k][0]; o[3 * k + 1] = c[k][1]; o[3 * k + 2] = c[k][2]; }
    ws.in_range[b * PP_MAX_TOPK + i] = inr ? 1 : 0;
}

// ---- polygon IoU (float64): quadrilateral a clipped by the four edges of b (Sutherland-Hodgman), shoelace area
__device__ __forceinline__ double quad_area_signed(const double (&p)[4][2]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 1) & 3;
        s += p[i][0] * p[j][1] - p[j][0] * p[i][1];
    }
    return 0.5 * s;
}

__device__ double quad_intersection_area(const double (&a)[4][2], const double (&b)[4][2]) {
    double cur[10][2], nxt[10][2];
    int n = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) { cur[i][0] = a[i][0]; cur[i][1] = a[i][1]; }
    for (int e = 0; e < 4 && n > 0; ++e) {
        const double x1 = b[e][0], y1 = b[e][1];
        const double ex = b[(e + 1) & 3][0] - x1, ey = b[(e + 1) & 3][1] - y1;
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const int j = (i + 1 == n) ? 0 : i + 1;
            const double px = cur[i][0], py = cur[i][1], qx = cur[j][0], qy = cur[j][1];
            const double dp = ex * (py - y1) - ey * (px - x1);      // >= 0: inside (left of the edge, b counter-clockwise)
            const double dq = ex * (qy - y1) - ey * (qx - x1);
            if (dp >= 0) { nxt[m][0] = px; nxt[m][1] = py; ++m; }
            if ((dp >= 0) != (dq >= 0)) {
                const double t = dp / (dp - dq);
                nxt[m][0] = px + t * (qx - px); nxt[m][1] = py + t * (qy - py); ++m;
            }
        }
        n = m;
        for (int i = 0; i < n; ++i) { cur[i][0] = nxt[i][0]; cur[i][1] = nxt[i][1]; }
    }
    if (n < 3) return 0.0;
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s += cur[i][0] * cur[j][1] - cur[j][0] * cur[i][1];
    }
    return fabs(0.5 * s);
}

__device__ __forceinline__ void load_quad_ccw(const float* box, double (&p)[4][2]) {
    double q[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i][0] = (double)box[3 * i]; q[i][1] = (double)box[3 * i + 1]; }   // convert_format: corners 0..3, (x, y)
    const bool ccw = quad_area_signed(q) >= 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = ccw ? i : 3 - i;
        p[i][0] = q[k][0]; p[i][1] = q[k][1];
    }
}

// grid (PP_WORDS col blocks, PP_WORDS row blocks, scenes), 64 threads: thread t = row rb*64+t against the 64 boxes of
// column block cb.  Only j > i matters for the greedy scan.
__global__ void __launch_bounds__(64) pp_iou_mask_kernel(const PostGeom g, PostWs ws) {
    const int b = blockIdx.z, cb = blockIdx.x, rb = blockIdx.y;
    const int K = ws.counters[b * 4 + 2];
    if (rb * 64 >= K) return;
    const int i = rb * 64 + threadIdx.x;
    __shared__ double sq[64][4][2];
    __shared__ double sarea[64];
    const float* boxes = ws.boxes + (size_t)b * PP_MAX_TOPK * 24;
    unsigned long long bits = 0ull;
    if (cb >= rb && cb * 64 < K) {
        const int j0 = cb * 64 + threadIdx.x;
        if (j0 < K) {
            double p[4][2];
            load_quad_ccw(boxes + (size_t)j0 * 24, p);
#pragma unroll
            for (int v = 0; v < 4; ++v) { sq[threadIdx.x][v][0] = p[v][0]; sq[threadIdx.x][v][1] = p[v][1]; }
            sarea[threadIdx.x] = fabs(quad_area_signed(p));
        }
        __syncthreads();
        if (i < K) {
            double a[4][2];
            load_quad_ccw(boxes + (size_t)i * 24, a);
            const double area_a = fabs(quad_area_signed(a));
            const int nj = min(64, K - cb * 64);
            for (int t = 0; t < nj; ++t) {
                const int j = cb * 64 + t;
                if (j <= i) continue;
                double q[4][2];
#pragma unroll
                for (int v = 0; v < 4; ++v) { q[v][0] = sq[t][v][0]; q[v][1] = sq[t][v][1]; }
                const double inter = quad_intersection_area(a, q);
                const float iou = (float)(inter / (area_a + sarea[t] - inter));     // compute_iou: np.float32
                if (iou > g.nms_thr) bits |= 1ull << t;
            }
        }
    }
    if (i < K) {
        unsigned long long* row = ws.mask + ((size_t)b * PP_MAX_TOPK + i) * PP_WORDS;
        row[cb] = bits;
        if (cb == 0)
            for (int w = gridDim.x; w < PP_WORDS; ++w) row[w] = 0ull;          // top_k < 1024: words past the grid
    }
}

// One CTA per scene: greedy NMS over the score-sorted boxes (warp 0, suppression matrix in shared memory), then the
// range mask (mask_boxes_outside_range_numpy) and the ordered output.
__global__ void __launch_bounds__(1024) pp_nms_scan_kernel(const PostGeom g, PostWs ws, float* __restrict__ out_boxes,
                                                           float* __restrict__ out_scores, int* __restrict__ out_count,
                                                           float* __restrict__ out_box7, int* __restrict__ out_index) {
    extern __shared__ unsigned long long smask[];              // [K][PP_WORDS]
    __shared__ short kept[PP_MAX_TOPK];
    __shared__ short outpos[PP_MAX_TOPK];
    __shared__ int s_nk, s_nout;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int K = ws.counters[b * 4 + 2];
    const unsigned long long* gm = ws.mask + (size_t)b * PP_MAX_TOPK * PP_WORDS;
    for (int i = tid; i < K * PP_WORDS; i += blockDim.x) smask[i] = gm[i];
    __syncthreads();
    if (tid < 32) {
        unsigned long long removed = 0ull;                     // lane w < PP_WORDS: suppression bits of boxes [64w, 64w+64)
        int nk = 0;
        for (int i = 0; i < K; ++i) {
            const unsigned long long r = __shfl_sync(0xffffffffu, removed, i >> 6);
            if (!((r >> (i & 63)) & 1ull)) {
                if (tid == 0) kept[nk] = (short)i;
                ++nk;
                if (tid < PP_WORDS) removed |= smask[i * PP_WORDS + tid];
            }
        }
        // range mask with order-preserving compaction
        int nout = 0;
        for (int base = 0; base < nk; base += 32) {
            const int t = base + tid;
            __syncwarp();
            const bool ok = t < nk && ws.in_range[b * PP_MAX_TOPK + kept[t]] != 0;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (t < nk) outpos[t] = ok ? (short)(nout + __popc(m & ((1u << tid) - 1u))) : (short)-1;
            nout += __popc(m);
        }
        if (tid == 0) { s_nk = nk; s_nout = nout; }
    }
    __syncthreads();
    const int nk = s_nk;
    const float* boxes = ws.boxes + (size_t)b * PP_MAX_TOPK * 24;
    for (int e = tid; e < nk * 24; e += blockDim.x) {
        const int t = e / 24, f = e - t * 24;
        const int pos = outpos[t];
        if (pos >= 0) out_boxes[((size_t)b * g.top_k + pos) * 24 + f] = boxes[(size_t)kept[t] * 24 + f];
    }
    for (int t = tid; t < nk; t += blockDim.x) {
        const int pos = outpos[t];
        if (pos >= 0) {
            out_scores[(size_t)b * g.top_k + pos] = ws.sorted_score[b * PP_MAX_TOPK + kept[t]];
            if (out_index) out_index[(size_t)b * g.top_k + pos] = ws.sorted_idx[b * PP_MAX_TOPK + kept[t]];
        }
    }
    if (out_box7) {
        for (int e = tid; e < nk * 7; e += blockDim.x) {
            const int t = e / 7, f = e - t * 7;
            const int pos = outpos[t];
            if (pos >= 0) out_box7[((size_t)b * g.top_k + pos) * 7 + f] = ws.box7[((size_t)b * PP_MAX_TOPK + kept[t]) * 7 + f];
        }
    }
    if (tid == 0) {
        out_count[b * 2 + 0] = s_nout;
        out_count[b * 2 + 1] = ws.counters[b * 4 + 0];
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t carve(int n_scenes, int cap, char* base, PostWs* ws) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cnt = take((size_t)n_scenes * 4 * sizeof(int));
    const size_t o_keys = take((size_t)n_scenes * cap * sizeof(unsigned long long));
    const size_t o_idx = take((size_t)n_scenes * PP_MAX_TOPK * sizeof(int));
    const size_t o_sc = take((size_t)n_scenes * PP_MAX_TOPK * sizeof(float));
    const size_t o_box = take((size_t)n_scenes * PP_MAX_TOPK * 24 * sizeof(float));
    const size_t o_inr = take((size_t)n_scenes * PP_MAX_TOPK);
    const size_t o_b7 = take((size_t)n_scenes * PP_MAX_TOPK * 7 * sizeof(float));
    const size_t o_mask = take((size_t)n_scenes * PP_MAX_TOPK * PP_WORDS * sizeof(unsigned long long));
    if (ws) {
        ws->counters = (int*)(base + o_cnt);
        ws->keys = (unsigned long long*)(base + o_keys);
        ws->sorted_idx = (int*)(base + o_idx);
        ws->sorted_score = (float*)(base + o_sc);
        ws->boxes = (float*)(base + o_box);
        ws->in_range = (unsigned char*)(base + o_inr);
        ws->box7 = (float*)(base + o_b7);
        ws->mask = (unsigned long long*)(base + o_mask);
    }
    return off;
}

}  // namespace cb

extern "C" size_t cb_postprocess_workspace_bytes(int n_scenes, int H, int W, int anchor_num) {
    if (n_scenes < 1 || H < 1 || W < 1 || anchor_num < 1) return 0;
    return cb::carve(n_scenes, H * W * anchor_num, nullptr, nullptr);
}

namespace cb {
static int postprocess_impl(const float* cls_preds, const float* reg_preds, const float* dir_preds, int n_scenes, int H,
                            int W, int anchor_num, int num_bins, const float* anchors, const float* tfm,
                            float score_threshold, float dir_offset, float nms_thresh, const double* gt_range,
                            int order_hwl, int top_k, float* out_boxes, float* out_scores, int32_t* out_count,
                            void* workspace, size_t workspace_bytes, void* stream, int stage1, float* out_box7,
                            int32_t* out_index) {
    if (!cls_preds || !reg_preds || !anchors || !out_boxes || !out_scores || !out_count || !workspace) return CB_ERR_ARG;
    if (!stage1 && (!tfm || !gt_range)) return CB_ERR_ARG;
    if (n_scenes < 1 || H < 1 || W < 1 || anchor_num < 1 || top_k < 1 || top_k > PP_MAX_TOPK) return CB_ERR_ARG;
    if (dir_preds && num_bins < 1) return CB_ERR_ARG;
    if ((long)H * W * anchor_num >= (1L << 31)) return CB_ERR_ARG;
    PostGeom g;
    g.H = H; g.W = W; g.A = anchor_num; g.num_bins = num_bins; g.has_dir = dir_preds ? 1 : 0; g.order_hwl = order_hwl ? 1 : 0;
    g.top_k = top_k; g.cap = H * W * anchor_num;
    g.score_thr = score_threshold; g.dir_offset = dir_offset; g.nms_thr = nms_thresh;
    g.period = (float)(2.0 * 3.141592653589793 / (double)(num_bins > 0 ? num_bins : 1));
    g.two_pi = (float)(2.0 * 3.141592653589793);
    g.stage1 = stage1 ? 1 : 0;
    for (int i = 0; i < 6; ++i) g.range[i] = gt_range ? gt_range[i] : 0.0;
    PostWs ws;
    if (carve(n_scenes, g.cap, (char*)workspace, &ws) > workspace_bytes) return CB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(ws.counters, 0, (size_t)n_scenes * 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    pp_decode_kernel<<<dim3((g.cap + 255) / 256, n_scenes), 256, 0, st>>>(cls_preds, reg_preds, dir_preds, anchors, tfm, g, ws);
    CB_CHECK_LAUNCH();
    pp_select_sort_kernel<<<n_scenes, 1024, 0, st>>>(g, ws);
    CB_CHECK_LAUNCH();
    pp_corners_kernel<<<dim3((top_k + 127) / 128, n_scenes), 128, 0, st>>>(cls_preds, reg_preds, dir_preds, anchors, tfm, g, ws);
    CB_CHECK_LAUNCH();
    const int blocks = (top_k + 63) / 64;
    pp_iou_mask_kernel<<<dim3(blocks, blocks, n_scenes), 64, 0, st>>>(g, ws);
    CB_CHECK_LAUNCH();
    static cudaError_t attr_err = cudaFuncSetAttribute(pp_nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       PP_MAX_TOPK * PP_WORDS * (int)sizeof(unsigned long long));
    if (attr_err != cudaSuccess) return (int)attr_err;
    pp_nms_scan_kernel<<<n_scenes, 1024, (size_t)top_k * PP_WORDS * sizeof(unsigned long long), st>>>(
        g, ws, out_boxes, out_scores, out_count, out_box7, out_index);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
}  // namespace cb

extern "C" int cb_postprocess(const float* cls_preds, const float* reg_preds, const float* dir_preds, int n_scenes, int H,
                              int W, int anchor_num, int num_bins, const float* anchors, const float* tfm,
                              float score_threshold, float dir_offset, float nms_thresh, const double* gt_range,
                              int order_hwl, int top_k, float* out_boxes, float* out_scores, int32_t* out_count,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return cb::postprocess_impl(cls_preds, reg_preds, dir_preds, n_scenes, H, W, anchor_num, num_bins, anchors, tfm,
                                score_threshold, dir_offset, nms_thresh, gt_range, order_hwl, top_k, out_boxes, out_scores,
                                out_count, workspace, workspace_bytes, stream, 0, nullptr, nullptr);
}

extern "C" int cb_postprocess_stage1(const float* cls_preds, const float* reg_preds, const float* dir_preds, int n_agents,
                                     int H, int W, int anchor_num, int num_bins, const float* anchors,
                                     float score_threshold, float dir_offset, float nms_thresh, int order_hwl, int top_k,
                                     float* out_corners, float* out_boxes7, int32_t* out_index, float* out_scores,
                                     int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
    if (!out_boxes7 || !out_index) return CB_ERR_ARG;
    return cb::postprocess_impl(cls_preds, reg_preds, dir_preds, n_agents, H, W, anchor_num, num_bins, anchors, nullptr,
                                score_threshold, dir_offset, nms_thresh, nullptr, order_hwl, top_k, out_corners, out_scores,
                                out_count, workspace, workspace_bytes, stream, 1, out_boxes7, out_index);
}
