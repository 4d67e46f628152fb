// Fused multi-agent feature warp + per-pixel ego-row attention (A10..A13) for sm_100a.
// One pass over the N agent maps: for every output pixel a warp bilinearly samples each agent's map at
// the affine-transformed location (F.affine_grid + F.grid_sample semantics: bilinear, zero padding,
// align_corners=False), forms the ego-row scaled-dot-product scores, soft-maxes over agents and writes
// the weighted sum.  HBM-bound: channels are innermost so every tap is a coalesced 128-byte read and the
// warped copies are never materialised.  Reference lines: include/coalign_b200.h.
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

constexpr int FUSE_MAX_AGENTS = 8;
constexpr int FUSE_MAX_CHUNKS = 4;     // C <= 256, 64 channels (2 per lane) per chunk

struct FuseGeom {
    int H, W, C, in_ps;
    int Hp, Wp;            // PF dims of input (in_ps=0) / of each parity plane (in_ps=1)
    long plane_rows;       // in_ps=1: rows per parity plane (= sum_agents*Hp*Wp)
    double inv_W, inv_H;   // 1/W, 1/H for the base grid
    float inv_sqrt_c;      // att_fuse.py:44 score scale
};

__device__ __forceinline__ long in_row(const FuseGeom& g, int agent, int y, int x) {
    if (!g.in_ps) return ((long)agent * g.Hp + y + 1) * g.Wp + x + 1;
    const int ph = (y & 1) * 2 + (x & 1);
    return (long)ph * g.plane_rows + ((long)agent * g.Hp + (y >> 1) + 1) * g.Wp + (x >> 1) + 1;
}

__global__ void normalize_affine_kernel(const double* __restrict__ pw, int n_scenes, int L, int H, int W, double ratio,
                                        double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scenes * L) return;
    const int b = i / L, j = i - b * L;
    const double* T = pw + (((long)b * L + 0) * L + j) * 16;           // pairwise[b][0][j] (ego row)
    double* A = out + (long)i * 6;
    // transformation_utils.py:84-89 (same operation order, float64)
    A[0] = T[0];
    A[1] = T[1] * H / W;
    A[2] = T[3] / (1 * ratio * W) * 2;
    A[3] = T[4] * W / H;
    A[4] = T[5];
    A[5] = T[7] / (1 * ratio * H) * 2;
}

template <int CHUNKS>
__global__ void __launch_bounds__(256) warp_att_fuse_kernel(const __nv_bfloat16* __restrict__ feat, long in_lo_off,
                                                            const double* __restrict__ affine,
                                                            const int* __restrict__ agent_off, int n_scenes, int L,
                                                            const FuseGeom g, int method,
                                                            __nv_bfloat16* __restrict__ out, long out_lo_off) {
    const int lane = threadIdx.x & 31;
    const long gw = (long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long)gridDim.x * 8;
    const long total = (long)n_scenes * g.H * g.W;
    const float inv_sqrt_c = (float)(1.0 / sqrt((double)g.C));
    for (long pix = gw; pix < total; pix += nw) {
        const int b = (int)(pix / (g.H * g.W));
        const int rem = (int)(pix - (long)b * g.H * g.W);
        const int h = rem / g.W, w = rem - h * g.W;
        const int a0 = agent_off[b];
        int n = agent_off[b + 1] - a0;
        n = n < FUSE_MAX_AGENTS ? n : FUSE_MAX_AGENTS;
        const double xs = (2.0 * w + 1.0) / g.W - 1.0;                  // affine_grid base grid, align_corners=False
        const double ys = (2.0 * h + 1.0) / g.H - 1.0;
        float x[FUSE_MAX_AGENTS][CHUNKS][2];
        float score[FUSE_MAX_AGENTS];
#pragma unroll
        for (int j = 0; j < FUSE_MAX_AGENTS; ++j) {
            if (j < n) {
                const double* A = affine + ((long)b * L + j) * 6;
                const float gx = (float)(A[0] * xs + A[1] * ys + A[2]);  // grid computed in f64, cast to f32
                const float gy = (float)(A[3] * xs + A[4] * ys + A[5]);
                const float ix = ((gx + 1.f) * g.W - 1.f) * 0.5f;        // grid_sample unnormalise
                const float iy = ((gy + 1.f) * g.H - 1.f) * 0.5f;
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
                // clamp before the int conversion so far-away (or non-finite) coordinates stay out of range
                const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)g.W + 1.f);
                const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)g.H + 1.f);
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) { x[j][c][0] = 0.f; x[j][c][1] = 0.f; }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int xx = x0 + (t & 1), yy = y0 + (t >> 1);
                    const float wt = ((t & 1) ? wx1 : wx0) * ((t >> 1) ? wy1 : wy0);
                    if (xx >= 0 && xx < g.W && yy >= 0 && yy < g.H) {
                        const long row = in_row(g, a0 + j, yy, xx);
                        const uint32_t* src = reinterpret_cast<const uint32_t*>(feat + row * g.C) + lane;
#pragma unroll
                        for (int c = 0; c < CHUNKS; ++c) {
                            const uint32_t u = __ldg(src + 32 * c);
                            float v0 = bf16_lo(u), v1 = bf16_hi(u);
                            if (in_lo_off != 0) {
                                const uint32_t ul = __ldg(reinterpret_cast<const uint32_t*>(feat + in_lo_off + row * g.C) + lane + 32 * c);
                                v0 += bf16_lo(ul); v1 += bf16_hi(ul);
                            }
                            x[j][c][0] = fmaf(wt, v0, x[j][c][0]);
                            x[j][c][1] = fmaf(wt, v1, x[j][c][1]);
                        }
                    }
                }
            }
        }
        float o[CHUNKS][2];
        if (method == 1) {                                               // MaxFusion
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) { o[c][0] = x[0][c][0]; o[c][1] = x[0][c][1]; }
#pragma unroll
            for (int j = 1; j < FUSE_MAX_AGENTS; ++j)
                if (j < n) {
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) { o[c][0] = fmaxf(o[c][0], x[j][c][0]); o[c][1] = fmaxf(o[c][1], x[j][c][1]); }
                }
        } else {
            float smax = -INFINITY;
#pragma unroll
            for (int j = 0; j < FUSE_MAX_AGENTS; ++j) {
                if (j < n) {
                    float d = 0.f;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) d += x[0][c][0] * x[j][c][0] + x[0][c][1] * x[j][c][1];
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
                    score[j] = d * inv_sqrt_c;                            // att_fuse.py:44
                    smax = fmaxf(smax, score[j]);
                }
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < FUSE_MAX_AGENTS; ++j)
                if (j < n) { score[j] = expf(score[j] - smax); den += score[j]; }
            const float inv = 1.f / den;
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) { o[c][0] = 0.f; o[c][1] = 0.f; }
#pragma unroll
            for (int j = 0; j < FUSE_MAX_AGENTS; ++j)
                if (j < n) {
                    const float wj = score[j] * inv;
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) { o[c][0] = fmaf(wj, x[j][c][0], o[c][0]); o[c][1] = fmaf(wj, x[j][c][1], o[c][1]); }
                }
        }
        const long orow = ((long)b * (g.H + 2) + h + 1) * (g.W + 2) + w + 1;   // PF output
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + orow * g.C) + lane;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const uint32_t hi = pack_bf16(o[c][0], o[c][1]);
            dst[32 * c] = hi;
            if (out_lo_off != 0)
                (reinterpret_cast<uint32_t*>(out + out_lo_off + orow * g.C) + lane)[32 * c] =
                    pack_bf16(o[c][0] - bf16_lo(hi), o[c][1] - bf16_hi(hi));
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// Fast path (C = 64 / 128 / 256): LPP = C/8 lanes per pixel, 8 channels (one 16-byte load) per lane and tap,
// so a warp keeps 32/LPP pixels x 4 taps x 16 B in flight per agent and every tap is a full 128..512 B line.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}

// 32-bit row index of pixel (agent,y,x) inside the PF / PS buffer (rows < 2^31 by construction)
__device__ __forceinline__ int in_row_i(const FuseGeom& g, int plane_rows, int agent, int y, int x) {
    if (!g.in_ps) return (agent * g.Hp + y + 1) * g.Wp + x + 1;
    const int ph = (y & 1) * 2 + (x & 1);
    return ph * plane_rows + (agent * g.Hp + (y >> 1) + 1) * g.Wp + (x >> 1) + 1;
}

template <int LPP, int MAXN, bool HAS_LO>
__global__ void __launch_bounds__(256, (MAXN <= 5 ? 3 : 2)) warp_att_fuse_v8_kernel(const __nv_bfloat16* __restrict__ feat, long in_lo_off,
                                                               const double* __restrict__ affine,
                                                               const int* __restrict__ agent_off, int n_scenes, int L,
                                                               const FuseGeom g, int method,
                                                               __nv_bfloat16* __restrict__ out, long out_lo_off) {
    constexpr int PPW = 32 / LPP;                          // pixels per warp
    static_assert(LPP >= MAXN, "one lane of the pixel group per agent");
    // tap records: per (warp, pixel of the warp, agent) 4 row indices (pre-multiplied by LPP) + 4 bilinear weights,
    // computed once by lane `agent` of the pixel group and broadcast to the group's other lanes through shared memory
    __shared__ __align__(16) uint4 s_tap[8][PPW][MAXN][2];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, sub = lane % LPP, pin = lane / LPP, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + wid, nw = gridDim.x * 8;
    const int HW = g.H * g.W;
    const int total = n_scenes * HW;
    const int plane_rows = (int)g.plane_rows;
    const float inv_sqrt_c = g.inv_sqrt_c;
    const uint4* feat4 = reinterpret_cast<const uint4*>(feat);
    const uint4* featl4 = reinterpret_cast<const uint4*>(feat + in_lo_off);
    for (int pix0 = gw * PPW; pix0 < total; pix0 += nw * PPW) {
        const int pix = pix0 + pin;
        const bool live = pix < total;
        const int pc = live ? pix : total - 1;              // idle lanes mirror a valid pixel, stores masked
        const int b = pc / HW;
        const int rem = pc - b * HW;
        const int h = rem / g.W, w = rem - h * g.W;
        const int a0 = agent_off[b];
        int n = agent_off[b + 1] - a0;
        n = n < MAXN ? n : MAXN;
        if (sub < n) {
            const int j = sub;
            const double xs = fma((double)(2 * w + 1), g.inv_W, -1.0);   // affine_grid base grid, align_corners=False
            const double ys = fma((double)(2 * h + 1), g.inv_H, -1.0);
            const double* A = affine + (b * L + j) * 6;
            const float gx = (float)(A[0] * xs + A[1] * ys + A[2]);   // grid in f64, cast to f32 (reference `.to(src)`)
            const float gy = (float)(A[3] * xs + A[4] * ys + A[5]);
            const float ix = ((gx + 1.f) * g.W - 1.f) * 0.5f;         // grid_sample unnormalise
            const float iy = ((gy + 1.f) * g.H - 1.f) * 0.5f;
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float wx1 = ix - fx0, wy1 = iy - fy0;
            // clamp before the int conversion so far-away (or non-finite) coordinates stay out of range
            const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)g.W + 1.f);
            const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)g.H + 1.f);
            // branch-free taps: out-of-range taps get weight 0 and a clamped (valid) address
            const float wxa = (x0 >= 0 && x0 < g.W) ? 1.f - wx1 : 0.f, wxb = (x0 + 1 >= 0 && x0 + 1 < g.W) ? wx1 : 0.f;
            const float wya = (y0 >= 0 && y0 < g.H) ? 1.f - wy1 : 0.f, wyb = (y0 + 1 >= 0 && y0 + 1 < g.H) ? wy1 : 0.f;
            const int xa = min(max(x0, 0), g.W - 1), xb = min(max(x0 + 1, 0), g.W - 1);
            const int ya = min(max(y0, 0), g.H - 1), yb = min(max(y0 + 1, 0), g.H - 1);
            const int ag = a0 + j;
            uint4 rr, ww;
            rr.x = in_row_i(g, plane_rows, ag, ya, xa) * LPP; rr.y = in_row_i(g, plane_rows, ag, ya, xb) * LPP;
            rr.z = in_row_i(g, plane_rows, ag, yb, xa) * LPP; rr.w = in_row_i(g, plane_rows, ag, yb, xb) * LPP;
            ww.x = __float_as_uint(wya * wxa); ww.y = __float_as_uint(wya * wxb);
            ww.z = __float_as_uint(wyb * wxa); ww.w = __float_as_uint(wyb * wxb);
            s_tap[wid][pin][j][0] = rr;
            s_tap[wid][pin][j][1] = ww;
        }
        __syncwarp();
        float x[MAXN][8];
#pragma unroll
        for (int j = 0; j < MAXN; ++j) {
#pragma unroll
            for (int c = 0; c < 8; ++c) x[j][c] = 0.f;
            if (j < n) {
                const uint4 rq = s_tap[wid][pin][j][0];
                const uint4 wq = s_tap[wid][pin][j][1];
                const float wt[4] = {__uint_as_float(wq.x), __uint_as_float(wq.y), __uint_as_float(wq.z), __uint_as_float(wq.w)};
                const unsigned rr[4] = {rq.x, rq.y, rq.z, rq.w};
                uint4 u[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) u[t] = __ldg(feat4 + (rr[t] + sub));   // row pitch C = 8*LPP
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    float v[8];
                    unpack8(u[t], v);
                    if (HAS_LO) {
                        float vl[8];
                        unpack8(__ldg(featl4 + (rr[t] + sub)), vl);
#pragma unroll
                        for (int c = 0; c < 8; ++c) v[c] += vl[c];
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[j][c] = fmaf(wt[t], v[c], x[j][c]);
                }
            }
        }
        __syncwarp();                                        // tap records consumed; next iteration may overwrite
        float o[8];
        if (method == 1) {                                   // MaxFusion (branch-free over absent agents)
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = x[0][c];
#pragma unroll
            for (int j = 1; j < MAXN; ++j) {
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = j < n ? fmaxf(o[c], x[j][c]) : o[c];
            }
        } else {
            // absent agents (j >= n) carry x = 0: their score is forced to -inf, so the warp stays converged and the
            // group reductions can use full-mask shuffles
            float score[MAXN];
            float smax = -INFINITY;
#pragma unroll
            for (int j = 0; j < MAXN; ++j) {
                float d = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) d = fmaf(x[0][c], x[j][c], d);
#pragma unroll
                for (int sft = LPP / 2; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);
                score[j] = j < n ? d * inv_sqrt_c : -INFINITY;
                smax = fmaxf(smax, score[j]);
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < MAXN; ++j) { score[j] = HAS_LO ? expf(score[j] - smax) : __expf(score[j] - smax); den += score[j]; }
            const float inv = 1.f / den;
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = 0.f;
#pragma unroll
            for (int j = 0; j < MAXN; ++j) {
                const float wj = score[j] * inv;
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = fmaf(wj, x[j][c], o[c]);
            }
        }
        if (live) {
            const int orow = (b * (g.H + 2) + h + 1) * (g.W + 2) + w + 1;   // PF output
            uint4 hi;
            hi.x = pack_bf16(o[0], o[1]); hi.y = pack_bf16(o[2], o[3]);
            hi.z = pack_bf16(o[4], o[5]); hi.w = pack_bf16(o[6], o[7]);
            reinterpret_cast<uint4*>(out)[(size_t)orow * LPP + sub] = hi;
            if (HAS_LO) {
                uint4 lo;
                lo.x = pack_bf16(o[0] - bf16_lo(hi.x), o[1] - bf16_hi(hi.x));
                lo.y = pack_bf16(o[2] - bf16_lo(hi.y), o[3] - bf16_hi(hi.y));
                lo.z = pack_bf16(o[4] - bf16_lo(hi.z), o[5] - bf16_hi(hi.z));
                lo.w = pack_bf16(o[6] - bf16_lo(hi.w), o[7] - bf16_hi(hi.w));
                reinterpret_cast<uint4*>(out + out_lo_off)[(size_t)orow * LPP + sub] = lo;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// v9: 16 channels (one 256-bit load) per lane and tap, LPP = C/16 lanes per pixel, so the per-pixel scalar work
// (tap records, soft-max, shuffles) is replicated over half as many lanes as in v8; agents are folded in one by one
// with an online soft-max (running max / denominator / weighted sum), so only the ego vector, the current agent's
// vector and the accumulator are live (3 x 16 registers instead of MAXN x 8 + 8); bilinear blending and the
// accumulator updates use packed FFMA2 (fma.rn.f32x2).  A CTA covers 8 rows x PPW columns and walks a contiguous
// range of such tiles along the row, so the four bilinear taps of neighbouring pixels hit L1.
// ---------------------------------------------------------------------------------------------------------
struct u8x { uint32_t w[8]; };
__device__ __forceinline__ u8x ldg256(const void* p) {
    u8x r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

template <int LPP, int MAXN, bool HAS_LO, bool BF16_BLEND, int OCC = 3>
__global__ void __launch_bounds__(256, OCC) warp_att_fuse_v9_kernel(const __nv_bfloat16* __restrict__ feat, long in_lo_off,
                                                                  const double* __restrict__ affine,
                                                                  const int* __restrict__ agent_off, int n_scenes, int L,
                                                                  const FuseGeom g, int method,
                                                                  __nv_bfloat16* __restrict__ out, long out_lo_off,
                                                                  int tiles_x, int tiles_per_cta) {
    constexpr int PPW = 32 / LPP;                          // pixels per warp (consecutive columns of one row)
    __shared__ __align__(16) uint4 s_tap[8][PPW][MAXN][2];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, sub = lane % LPP, pin = lane / LPP, wid = threadIdx.x >> 5;
    const int bands = (g.H + 7) >> 3;
    const int total_tiles = n_scenes * bands * tiles_x;
    const int plane_rows = (int)g.plane_rows;
    const float inv_sqrt_c = g.inv_sqrt_c;
    const char* featb = reinterpret_cast<const char*>(feat);
    const int t_begin = blockIdx.x * tiles_per_cta;
    const int t_end = min(total_tiles, t_begin + tiles_per_cta);
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int tx = tile % tiles_x;
        const int bb = tile / tiles_x;
        const int band = bb % bands, b = bb / bands;
        const int h = band * 8 + wid;
        const int w = tx * PPW + pin;
        const bool row_ok = h < g.H;                         // warp-uniform
        const bool live = row_ok && w < g.W;
        const int hc = row_ok ? h : g.H - 1, wc = w < g.W ? w : g.W - 1;
        const int a0 = agent_off[b];
        int n = agent_off[b + 1] - a0;
        n = n < MAXN ? n : MAXN;
        if (row_ok) {
            for (int j = sub; j < n; j += LPP) {
                const double xs = fma((double)(2 * wc + 1), g.inv_W, -1.0);   // affine_grid base grid, align_corners=False
                const double ys = fma((double)(2 * hc + 1), g.inv_H, -1.0);
                const double* A = affine + (b * L + j) * 6;
                const float gx = (float)(A[0] * xs + A[1] * ys + A[2]);   // grid in f64, cast to f32 (reference `.to(src)`)
                const float gy = (float)(A[3] * xs + A[4] * ys + A[5]);
                const float ix = ((gx + 1.f) * g.W - 1.f) * 0.5f;         // grid_sample unnormalise
                const float iy = ((gy + 1.f) * g.H - 1.f) * 0.5f;
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const float wx1 = ix - fx0, wy1 = iy - fy0;
                const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)g.W + 1.f);
                const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)g.H + 1.f);
                const float wxa = (x0 >= 0 && x0 < g.W) ? 1.f - wx1 : 0.f, wxb = (x0 + 1 >= 0 && x0 + 1 < g.W) ? wx1 : 0.f;
                const float wya = (y0 >= 0 && y0 < g.H) ? 1.f - wy1 : 0.f, wyb = (y0 + 1 >= 0 && y0 + 1 < g.H) ? wy1 : 0.f;
                const int xa = min(max(x0, 0), g.W - 1), xb = min(max(x0 + 1, 0), g.W - 1);
                const int ya = min(max(y0, 0), g.H - 1), yb = min(max(y0 + 1, 0), g.H - 1);
                const int ag = a0 + j;
                uint4 rr, ww;                                             // byte offsets of the 4 tap rows (row pitch 2*C)
                rr.x = (unsigned)in_row_i(g, plane_rows, ag, ya, xa) * (unsigned)(LPP * 32);
                rr.y = (unsigned)in_row_i(g, plane_rows, ag, ya, xb) * (unsigned)(LPP * 32);
                rr.z = (unsigned)in_row_i(g, plane_rows, ag, yb, xa) * (unsigned)(LPP * 32);
                rr.w = (unsigned)in_row_i(g, plane_rows, ag, yb, xb) * (unsigned)(LPP * 32);
                ww.x = __float_as_uint(wya * wxa); ww.y = __float_as_uint(wya * wxb);
                ww.z = __float_as_uint(wyb * wxa); ww.w = __float_as_uint(wyb * wxb);
                s_tap[wid][pin][j][0] = rr;
                s_tap[wid][pin][j][1] = ww;
            }
        }
        __syncwarp();
        if (row_ok) {
            // x <- this lane's 16 channels of agent j's map, bilinearly sampled at this pixel
            auto gather = [&](int j, f2 (&x)[8]) {
                const uint4 rq = s_tap[wid][pin][j][0];
                const uint4 wq = s_tap[wid][pin][j][1];
                const float wt[4] = {__uint_as_float(wq.x), __uint_as_float(wq.y), __uint_as_float(wq.z), __uint_as_float(wq.w)};
                const unsigned rr[4] = {rq.x, rq.y, rq.z, rq.w};
                // Taps whose weight is zero for every pixel of the warp are neither loaded nor blended (0 * v adds
                // nothing): the agent does not see this part of the map (all four zero), or its transform is a pure
                // pixel-aligned shift such as the ego's identity (only tap 0 non-zero).  Mixed cases take all four.
                const bool only0 = !__any_sync(0xffffffffu, (wq.y | wq.z | wq.w) != 0u);
                const bool none = only0 && !__any_sync(0xffffffffu, wq.x != 0u);
                if (none) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) { x[c].x = 0.f; x[c].y = 0.f; }
                } else if (BF16_BLEND) {
                    // bf16 production mode: the four taps are blended with packed bf16 FMAs (HFMA2.BF16, 16 channels in
                    // 8 instructions per tap instead of 24); the blended vector carries one more bf16 rounding than the
                    // fp32 blend, the same size as the rounding of any activation store of this mode.  Scores, soft-max
                    // and the weighted sum stay fp32.
                    uint32_t xb[8];
                    if (only0) {
                        const u8x u0 = ldg256(featb + (size_t)rr[0] + sub * 32);
                            const __nv_bfloat162 w0 = __float2bfloat162_rn(wt[0]);
#pragma unroll
                        for (int c = 0; c < 8; ++c) xb[c] = bf2_u32(__hmul2(u32_bf2(u0.w[c]), w0));
                    } else {
                        __nv_bfloat162 wb[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) wb[t] = __float2bfloat162_rn(wt[t]);
                        if (OCC >= 4) {                              // 64-register variant: two taps in flight at a time
                            u8x ua = ldg256(featb + (size_t)rr[0] + sub * 32), ub = ldg256(featb + (size_t)rr[1] + sub * 32);
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                xb[c] = bf2_u32(__hfma2(u32_bf2(ub.w[c]), wb[1], __hmul2(u32_bf2(ua.w[c]), wb[0])));
                            ua = ldg256(featb + (size_t)rr[2] + sub * 32); ub = ldg256(featb + (size_t)rr[3] + sub * 32);
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                xb[c] = bf2_u32(__hfma2(u32_bf2(ub.w[c]), wb[3], __hfma2(u32_bf2(ua.w[c]), wb[2], u32_bf2(xb[c]))));
                        } else {
                            u8x u[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) u[t] = ldg256(featb + (size_t)rr[t] + sub * 32);
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                __nv_bfloat162 acc = __hmul2(u32_bf2(u[0].w[c]), wb[0]);
                                acc = __hfma2(u32_bf2(u[1].w[c]), wb[1], acc);
                                acc = __hfma2(u32_bf2(u[2].w[c]), wb[2], acc);
                                acc = __hfma2(u32_bf2(u[3].w[c]), wb[3], acc);
                                xb[c] = bf2_u32(acc);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) { x[c].x = bf16_lo(xb[c]); x[c].y = bf16_hi(xb[c]); }
                } else {
                    auto tap = [&](const u8x& ut, unsigned row_off, int c) {
                        f2 v = {bf16_lo(ut.w[c]), bf16_hi(ut.w[c])};
                        (void)row_off;
                        return v;
                    };
                    if (only0) {
                        const u8x u0 = ldg256(featb + (size_t)rr[0] + sub * 32);
                            u8x l0 = u0;
                        if (HAS_LO) l0 = ldg256(featb + 2 * in_lo_off + (size_t)rr[0] + sub * 32);
                        const f2 w2 = {wt[0], wt[0]};
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            f2 v = tap(u0, rr[0], c);
                            if (HAS_LO) { v.x += bf16_lo(l0.w[c]); v.y += bf16_hi(l0.w[c]); }
                            x[c] = fmul2(w2, v);
                        }
                    } else {
                        u8x u[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) u[t] = ldg256(featb + (size_t)rr[t] + sub * 32);
    #pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            u8x lt = u[t];
                            if (HAS_LO) lt = ldg256(featb + 2 * in_lo_off + (size_t)rr[t] + sub * 32);
                            const f2 w2 = {wt[t], wt[t]};
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                f2 v = tap(u[t], rr[t], c);
                                if (HAS_LO) { v.x += bf16_lo(lt.w[c]); v.y += bf16_hi(lt.w[c]); }
                                x[c] = t == 0 ? fmul2(w2, v) : fma2(w2, v, x[c]);
                            }
                        }
                    }
                }
            };
            // agents are folded in one by one, the ego first: online soft-max over the ego-row scores <x0, xj>/sqrt(C)
            f2 x0[8], o[8];
            gather(0, x0);
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = x0[c];
            float m_run = 0.f, l_run = 1.f;
            if (method != 1) {
                f2 d2 = fmul2(x0[0], x0[0]);
#pragma unroll
                for (int c = 1; c < 8; ++c) d2 = fma2(x0[c], x0[c], d2);
                float d = d2.x + d2.y;
#pragma unroll
                for (int sft = LPP / 2; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);
                m_run = d * inv_sqrt_c;                                    // att_fuse.py:44
            }
            for (int j = 1; j < n; ++j) {
                f2 x[8];
                gather(j, x);
                if (method == 1) {                           // MaxFusion
#pragma unroll
                    for (int c = 0; c < 8; ++c) { o[c].x = fmaxf(o[c].x, x[c].x); o[c].y = fmaxf(o[c].y, x[c].y); }
                } else {
                    f2 d2 = fmul2(x0[0], x[0]);
#pragma unroll
                    for (int c = 1; c < 8; ++c) d2 = fma2(x0[c], x[c], d2);
                    float d = d2.x + d2.y;
#pragma unroll
                    for (int sft = LPP / 2; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);
                    const float sc = d * inv_sqrt_c;
                    const float m_new = fmaxf(m_run, sc);
                    const float alpha = HAS_LO ? expf(m_run - m_new) : __expf(m_run - m_new);
                    const float pj = HAS_LO ? expf(sc - m_new) : __expf(sc - m_new);
                    l_run = fmaf(l_run, alpha, pj);
                    const f2 a2 = {alpha, alpha}, p2 = {pj, pj};
#pragma unroll
                    for (int c = 0; c < 8; ++c) o[c] = fma2(o[c], a2, fmul2(p2, x[c]));
                    m_run = m_new;
                }
            }
            if (method != 1) {
                const float inv = 1.f / l_run;
                const f2 i2 = {inv, inv};
#pragma unroll
                for (int c = 0; c < 8; ++c) o[c] = fmul2(o[c], i2);
            }
            if (live) {
                const int orow = (b * (g.H + 2) + h + 1) * (g.W + 2) + w + 1;   // PF output
                uint32_t hi[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) hi[c] = pack_bf16(o[c].x, o[c].y);
                stg256(reinterpret_cast<char*>(out) + ((size_t)orow * LPP + sub) * 32, hi);
                if (HAS_LO) {
                    uint32_t lo[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) lo[c] = pack_bf16(o[c].x - bf16_lo(hi[c]), o[c].y - bf16_hi(hi[c]));
                    stg256(reinterpret_cast<char*>(out + out_lo_off) + ((size_t)orow * LPP + sub) * 32, lo);
                }
            }
        }
        __syncwarp();                                        // tap records consumed; next tile may overwrite
    }
}

}  // namespace cb

extern "C" int cb_normalize_affine(const double* pairwise_t_matrix, int n_scenes, int max_cav, int H, int W,
                                   double discrete_ratio, double* affine_out, void* stream) {
    if (n_scenes < 1 || max_cav < 1) return CB_ERR_ARG;
    const int n = n_scenes * max_cav;
    cb::normalize_affine_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pairwise_t_matrix, n_scenes, max_cav,
                                                                                  H, W, discrete_ratio, affine_out);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_warp_att_fuse(const void* feat, int in_ps, int64_t in_lo_off, int sum_agents, const double* affine,
                                const int32_t* agent_off, int n_scenes, int max_cav, int H, int W, int C, int method,
                                void* out_pf, int64_t out_lo_off, void* stream) {
    using namespace cb;
    if (C % 64 != 0 || C < 64 || C > 64 * FUSE_MAX_CHUNKS || n_scenes < 1 || max_cav < 1 || sum_agents < 1)
        return CB_ERR_ARG;
    if (method != 0 && method != 1) return CB_ERR_ARG;
    if (!feat || !out_pf || !affine || !agent_off) return CB_ERR_ARG;
    FuseGeom g;
    g.H = H; g.W = W; g.C = C; g.in_ps = in_ps;
    g.Hp = in_ps ? (H + 1) / 2 + 2 : H + 2;
    g.Wp = in_ps ? (W + 1) / 2 + 2 : W + 2;
    g.plane_rows = (long)sum_agents * g.Hp * g.Wp;
    g.inv_W = 1.0 / W; g.inv_H = 1.0 / H;
    g.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    const long total = (long)n_scenes * H * W;
    long blocks = (total + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* f = (const __nv_bfloat16*)feat;
    __nv_bfloat16* o = (__nv_bfloat16*)out_pf;
    if ((C == 64 || C == 128 || C == 256) && opt_get(CB_OPT_FUSE_VERSION) >= 8) {
        const int ppw = 256 / C;                                    // pixels per warp
        long nb = (total + 8L * ppw - 1) / (8L * ppw);
        if (nb > 148 * 24) nb = 148 * 24;
        if ((long)sum_agents * g.Hp * g.Wp * (in_ps ? 4 : 1) * (C / 8) >= (1L << 31) || total >= (1L << 31)) return CB_ERR_ARG;
        if ((in_lo_off != 0) != (out_lo_off != 0)) return CB_ERR_ARG;   // hi/lo planes: both or neither
#define CB_FUSE_LAUNCH(LPP_, MAXN_) do { if (in_lo_off != 0) \
            launch_pdl(warp_att_fuse_v8_kernel<LPP_, MAXN_, true>, dim3((unsigned)nb), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off); \
        else launch_pdl(warp_att_fuse_v8_kernel<LPP_, MAXN_, false>, dim3((unsigned)nb), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off); } while (0)
        const int fuse_ver = opt_get(CB_OPT_FUSE_VERSION);
        // bf16 maps (no lo plane): blend the bilinear taps with packed bf16 FMAs; CB_OPT_FUSE_BLEND_FP32 keeps the fp32 blend
        const bool bf16_blend = opt_get(CB_OPT_FUSE_BLEND_FP32) == 0;
        // 64 registers / 4 CTAs per SM (two taps in flight at a time) measured 4 % faster than 80 / 3 (CB_OPT_FUSE_OCC3)
        const bool occ4 = opt_get(CB_OPT_FUSE_OCC3) == 0;
        const bool small = max_cav <= 5;
        // v9 indexes tap rows by 32-bit BYTE offsets and needs 32-byte aligned buffers
        const bool v9_ok = fuse_ver >= 9 && (long)sum_agents * g.Hp * g.Wp * (in_ps ? 4 : 1) * (long)(2 * C) < (1L << 32) &&
                           ((uintptr_t)feat % 32) == 0 && ((uintptr_t)out_pf % 32) == 0 && (in_lo_off % 16) == 0 &&
                           (out_lo_off % 16) == 0;
        if (v9_ok) {
            const int lpp = C / 16, ppw9 = 32 / lpp;
            const int tiles_x = (W + ppw9 - 1) / ppw9;
            const int total_tiles = n_scenes * ((H + 7) / 8) * tiles_x;
            int grid = 148 * (occ4 && bf16_blend && in_lo_off == 0 ? 4 : 3) * 4;   // ~4 contiguous tile ranges per resident CTA slot
            if (grid > total_tiles) grid = total_tiles;
            const int tiles_per_cta = (total_tiles + grid - 1) / grid;
            grid = (total_tiles + tiles_per_cta - 1) / tiles_per_cta;
#define CB_FUSE9_LAUNCH(LPP_, MAXN_) do { if (in_lo_off != 0) \
            launch_pdl(warp_att_fuse_v9_kernel<LPP_, MAXN_, true, false>, dim3((unsigned)grid), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off, tiles_x, tiles_per_cta); \
        else if (bf16_blend && occ4) launch_pdl(warp_att_fuse_v9_kernel<LPP_, MAXN_, false, true, 4>, dim3((unsigned)grid), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off, tiles_x, tiles_per_cta); \
        else if (bf16_blend) launch_pdl(warp_att_fuse_v9_kernel<LPP_, MAXN_, false, true>, dim3((unsigned)grid), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off, tiles_x, tiles_per_cta); \
        else launch_pdl(warp_att_fuse_v9_kernel<LPP_, MAXN_, false, false>, dim3((unsigned)grid), dim3(256), 0, st, \
                       f, (long)in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, (long)out_lo_off, tiles_x, tiles_per_cta); } while (0)
            switch (C) {
                case 64: if (small) CB_FUSE9_LAUNCH(4, 5); else CB_FUSE9_LAUNCH(4, FUSE_MAX_AGENTS); break;
                case 128: if (small) CB_FUSE9_LAUNCH(8, 5); else CB_FUSE9_LAUNCH(8, FUSE_MAX_AGENTS); break;
                default: if (small) CB_FUSE9_LAUNCH(16, 5); else CB_FUSE9_LAUNCH(16, FUSE_MAX_AGENTS); break;
            }
#undef CB_FUSE9_LAUNCH
            CB_CHECK_LAUNCH();
            return CB_OK;
        }
        switch (C) {
            case 64: if (small) CB_FUSE_LAUNCH(8, 5); else CB_FUSE_LAUNCH(8, FUSE_MAX_AGENTS); break;
            case 128: if (small) CB_FUSE_LAUNCH(16, 5); else CB_FUSE_LAUNCH(16, FUSE_MAX_AGENTS); break;
            default: if (small) CB_FUSE_LAUNCH(32, 5); else CB_FUSE_LAUNCH(32, FUSE_MAX_AGENTS); break;
        }
#undef CB_FUSE_LAUNCH
        CB_CHECK_LAUNCH();
        return CB_OK;
    }
    switch (C / 64) {
        case 1: warp_att_fuse_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(f, in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, out_lo_off); break;
        case 2: warp_att_fuse_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(f, in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, out_lo_off); break;
        case 3: warp_att_fuse_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(f, in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, out_lo_off); break;
        case 4: warp_att_fuse_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(f, in_lo_off, affine, agent_off, n_scenes, max_cav, g, method, o, out_lo_off); break;
        default: return CB_ERR_ARG;
    }
    CB_CHECK_LAUNCH();
    return CB_OK;
}
