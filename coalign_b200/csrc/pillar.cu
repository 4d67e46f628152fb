// Pillar front-end on sm_100a: deterministic voxelisation (A1/A2), PillarVFE+PFN (A3/A4) and
// PointPillarScatter (A5), either staged through reference-format voxel tensors or fused
// points -> canvas.  HBM-bound integer/byte work: coalesced 16-byte point loads, 8 lanes per pillar
// (16-byte stores, 128 B per canvas cell).  See include/coalign_b200.h for the reference lines.
//
// Determinism: spconv's generator is serial (voxel id = order of first appearance, a voxel keeps its
// first `max_pts` points, new voxels are refused after `max_voxels`).  We reproduce it bit-exactly with
//   K1   cell id per point, atomicMin(first point index per cell), per-cell count
//   K2a  per 1024-point chunk: number of "leader" points (first[cell]==i) and sum of their cell counts
//   K2b  per agent: exclusive scan of the chunk totals
//   K2c  per chunk: ordered in-block scan + chunk prefix -> voxel ids, CSR offsets
//   K3   CSR fill (arbitrary order inside a cell)
//   K3b  thread per point: rank inside its cell (points with a smaller index), regroup into voxel-slot order
//   K4   8 lanes per voxel: emit reference-format tensors.
// The hot path (cb_points_to_canvas) does not need the voxel ORDER and uses the v2 pipeline further down.
#include <limits.h>
#include <mutex>
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

constexpr int CHUNK = 1024;       // points per scan chunk (256 threads x 4)
constexpr int NCLS = 5;           // pillar classes by point count: 1 | 2 | 3-4 | 5-8 | 9+

struct AgentOffsets { int n_agents; int off[CB_MAX_AGENTS + 1]; };
// View of the per-agent point offsets: either the by-value kernel parameter (host-array entry points) or a device array
// (cb_points_to_canvas_dev: offsets change every frame without changing the kernel arguments of a captured CUDA graph).
struct AoView { int n_agents; const int* off; };
__device__ __forceinline__ AoView ao_view(const AgentOffsets& v, const int* off_dev) {
    AoView a; a.n_agents = v.n_agents; a.off = off_dev ? off_dev : v.off; return a;
}

struct VoxWs {            // workspace carve-up (device pointers)
    int* first;           // [n_agents][ncell]   min point index per cell (0x7f7f7f7f = empty)
    int* count;           // [n_agents][ncell]
    int* cursor;          // [n_agents][vcap]
    int* cell2vox;        // [n_agents][ncell]
    int* vox_off;         // [n_agents][vcap+1]
    int4* vox_meta;       // [n_agents][vcap]   {CSR begin, point count, x | y<<12 | z<<24, first point index}
    int* cellid;          // [sum_P]
    int* list;            // [sum_P]
    int* nvox;            // [n_agents+1]
    int2* chunk_tot;      // [n_agents][max_chunks]  (leaders, points) per chunk, then exclusive prefixes
    float4* vp;           // [sum_P]  points regrouped per voxel in slot order (first max_pts of each voxel)
    int ncell, vcap, max_chunks;
};

struct Geom { float r0, r1, r2, v0, v1, v2; int gx, gy, gz; };

// PS canvas addressing (include/coalign_b200.h): 4 parity planes of PF-padded half-resolution maps
struct CanvasGeom {
    int ny, nx, Hq, Wq;   // Hq = ceil(ny/2)+2, Wq = ceil(nx/2)+2  (padded plane dims)
    long plane_rows;      // canvas_agents*Hq*Wq
};
__device__ __forceinline__ long canvas_row(const CanvasGeom& c, int a, int y, int x) {
    const int ph = (y & 1) * 2 + (x & 1);                                      // rows < 2^31 (checked on the host)
    return (long)(ph * (int)c.plane_rows + (a * c.Hq + (y >> 1) + 1) * c.Wq + (x >> 1) + 1);
}

__device__ __forceinline__ int find_agent(const AgentOffsets& ao, int i) {
    int a = 0;
    while (a + 1 < ao.n_agents && i >= ao.off[a + 1]) ++a;
    return a;
}

// K1 ------------------------------------------------------------------------------------------
__global__ void vox_assign_kernel(const float4* __restrict__ pts, const __grid_constant__ AgentOffsets ao,
                                  const Geom g, const VoxWs ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ao.off[ao.n_agents]) return;
    const int a = find_agent(ao, i);
    const float4 p = __ldg(pts + i);
    // float32 arithmetic exactly as the serial generator: floor((p - min) / vs); IEEE division, no FMA
    const float fx = floorf(__fdiv_rn(__fsub_rn(p.x, g.r0), g.v0));
    const float fy = floorf(__fdiv_rn(__fsub_rn(p.y, g.r1), g.v1));
    const float fz = floorf(__fdiv_rn(__fsub_rn(p.z, g.r2), g.v2));
    int cell = -1;
    if (fx >= 0.f && fx < (float)g.gx && fy >= 0.f && fy < (float)g.gy && fz >= 0.f && fz < (float)g.gz) {
        cell = ((int)fz * g.gy + (int)fy) * g.gx + (int)fx;
        const long base = (long)a * ws.ncell + cell;
        atomicMin(ws.first + base, i - ao.off[a]);
        atomicAdd(ws.count + base, 1);
    }
    ws.cellid[i] = cell;
}

// K2 ------------------------------------------------------------------------------------------
// Thread t of a chunk CTA owns points [4t, 4t+4) of the chunk: leader flags and their cell counts.
__device__ __forceinline__ void chunk_flags(const AgentOffsets& ao, const VoxWs& ws, int a, int chunk, int (&cell)[4],
                                            int (&lead)[4], int (&cnt)[4]) {
    const int p0 = ao.off[a], np = ao.off[a + 1] - p0;
    const int* first = ws.first + (long)a * ws.ncell;
    const int* count = ws.count + (long)a * ws.ncell;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = chunk * CHUNK + threadIdx.x * 4 + j;
        cell[j] = -1; lead[j] = 0; cnt[j] = 0;
        if (i < np) cell[j] = ws.cellid[p0 + i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = chunk * CHUNK + threadIdx.x * 4 + j;
        if (cell[j] >= 0 && first[cell[j]] == i) { lead[j] = 1; cnt[j] = count[cell[j]]; }
    }
}

// block scan of (v,c) over 256 threads: exclusive prefix of this thread and block totals
__device__ __forceinline__ void block_scan2(int v, int c, int& ex_v, int& ex_c, int& tot_v, int& tot_c) {
    __shared__ int s_v[8], s_c[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int iv = v, ic = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int tv = __shfl_up_sync(0xffffffffu, iv, d), tc = __shfl_up_sync(0xffffffffu, ic, d);
        if (lane >= d) { iv += tv; ic += tc; }
    }
    if (lane == 31) { s_v[warp] = iv; s_c[warp] = ic; }
    __syncthreads();
    int pv = 0, pc = 0, tv = 0, tc = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) { pv += s_v[w]; pc += s_c[w]; }
        tv += s_v[w]; tc += s_c[w];
    }
    ex_v = pv + iv - v; ex_c = pc + ic - c;
    tot_v = tv; tot_c = tc;
    __syncthreads();
}

__global__ void __launch_bounds__(256) vox_chunk_count_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws) {
    const int a = blockIdx.y, chunk = blockIdx.x;
    int cell[4], lead[4], cnt[4];
    chunk_flags(ao, ws, a, chunk, cell, lead, cnt);            // chunks past the agent's end yield zeros
    int ev, ec, tv, tc;
    block_scan2(lead[0] + lead[1] + lead[2] + lead[3], cnt[0] + cnt[1] + cnt[2] + cnt[3], ev, ec, tv, tc);
    if (threadIdx.x == 0) ws.chunk_tot[(long)a * ws.max_chunks + chunk] = make_int2(tv, tc);
}

// one warp per agent: chunk totals -> exclusive prefixes; voxel count
__global__ void vox_chunk_scan_kernel(const VoxWs ws, int n_chunks, int max_voxels) {
    const int a = blockIdx.x, lane = threadIdx.x;
    int2* ct = ws.chunk_tot + (long)a * ws.max_chunks;
    int run_v = 0, run_c = 0;
    for (int base = 0; base < n_chunks; base += 32) {
        const int i = base + lane;
        const int2 t = i < n_chunks ? ct[i] : make_int2(0, 0);
        int iv = t.x, ic = t.y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int tv = __shfl_up_sync(0xffffffffu, iv, d), tc = __shfl_up_sync(0xffffffffu, ic, d);
            if (lane >= d) { iv += tv; ic += tc; }
        }
        if (i < n_chunks) ct[i] = make_int2(run_v + iv - t.x, run_c + ic - t.y);
        run_v += __shfl_sync(0xffffffffu, iv, 31);
        run_c += __shfl_sync(0xffffffffu, ic, 31);
    }
    if (lane == 0) ws.nvox[a] = run_v < max_voxels ? run_v : max_voxels;
}

__global__ void vox_total_kernel(const VoxWs ws, int n_agents, int* n_voxels_out, int* dirty_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int t = 0;
        for (int a = 0; a < n_agents; ++a) {
            if (n_voxels_out) n_voxels_out[a] = ws.nvox[a];
            t += ws.nvox[a];
        }
        ws.nvox[n_agents] = t;
        if (n_voxels_out) n_voxels_out[n_agents] = t;
        if (dirty_count) *dirty_count = t;
    }
}

__global__ void __launch_bounds__(256) vox_chunk_assign_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                               const Geom g, int max_voxels) {
    const int a = blockIdx.y, chunk = blockIdx.x;
    int cell[4], lead[4], cnt[4];
    chunk_flags(ao, ws, a, chunk, cell, lead, cnt);
    int ev, ec, tv, tc;
    block_scan2(lead[0] + lead[1] + lead[2] + lead[3], cnt[0] + cnt[1] + cnt[2] + cnt[3], ev, ec, tv, tc);
    const int2 pre = ws.chunk_tot[(long)a * ws.max_chunks + chunk];
    int pv = pre.x + ev, pc = pre.y + ec;
    int* cell2vox = ws.cell2vox + (long)a * ws.ncell;
    int* vox_off = ws.vox_off + (long)a * (ws.vcap + 1);
    int4* vox_meta = ws.vox_meta + (long)a * ws.vcap;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (lead[j]) {
            if (pv < max_voxels) {
                cell2vox[cell[j]] = pv;
                vox_off[pv] = pc;                           // CSR begin (used by the fill kernel)
                const int cx = cell[j] % g.gx, cyz = cell[j] / g.gx;
                const int cy = cyz % g.gy, cz = cyz / g.gy;
                vox_meta[pv] = make_int4(pc, cnt[j], cx | (cy << 12) | (cz << 24), chunk * CHUNK + (int)threadIdx.x * 4 + j);
            } else {
                cell2vox[cell[j]] = -1;                    // refused: max_voxels reached
            }
            pv += 1; pc += cnt[j];
        }
    }
}

// K3 ------------------------------------------------------------------------------------------
__global__ void vox_fill_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ao.off[ao.n_agents]) return;
    const int cell = ws.cellid[i];
    if (cell < 0) return;
    const int a = find_agent(ao, i);
    const int v = ws.cell2vox[(long)a * ws.ncell + cell];
    if (v < 0) return;
    const int pos = atomicAdd(ws.cursor + (long)a * ws.vcap + v, 1);
    ws.list[ao.off[a] + ws.vox_off[(long)a * (ws.vcap + 1) + v] + pos] = i - ao.off[a];
}

// K3b: thread per point: rank of the point inside its voxel (= number of same-cell points with a smaller index) and
// regrouping into voxel-slot order.  Massively parallel, two dependent loads for 1-point voxels.
__global__ void vox_rank_gather_kernel(const float4* __restrict__ pts, const __grid_constant__ AgentOffsets ao,
                                       const VoxWs ws, int max_pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ao.off[ao.n_agents]) return;
    const int cell = ws.cellid[i];
    if (cell < 0) return;
    const int a = find_agent(ao, i);
    const int v = ws.cell2vox[(long)a * ws.ncell + cell];
    if (v < 0) return;
    const int4 m = ws.vox_meta[(long)a * ws.vcap + v];           // {begin, count, xyz, first}
    const int li = i - ao.off[a];
    int rank = 0;
    if (m.y > 1) {
        const int* lst = ws.list + ao.off[a] + m.x;
        for (int k = 0; k < m.y; ++k) rank += (lst[k] < li) ? 1 : 0;
    }
    if (rank < max_pts) ws.vp[ao.off[a] + m.x + rank] = __ldg(pts + i);
}

// K4 helpers ----------------------------------------------------------------------------------
// PFN in "centred" form.  With q = p - centre (the reference's f_center, pillar_vfe.py:127-131) and m' = mean(q) =
// mean(p) - centre, the 10 input features of pillar_vfe.py:118-135 are  [q + centre, i, q - m', q], so
//   bn(linear(f))_c = A0_c qx + A1_c qy + A2_c qz + A3_c i  +  [ sh_c + C_c . centre + M_c . m' ]
// with A0 = sc (w0 + w4 + w7), A1 = sc (w1 + w5 + w8), A2 = sc (w2 + w6 + w9), A3 = sc w3, C = sc (w0, w1, w2),
// M = -sc (w4, w5, w6).  The bracket is a per-pillar constant, so a point costs 4 FMAs per channel instead of 11 (and
// they are issued two channels at a time as FFMA2).  All large-magnitude terms (centre up to 140 m) are separated from
// the small ones (|q| <= voxel/2), so the result carries the same ~1e-7 relative rounding as the reference's fp32.
struct PfnParams {
    const float* w; const float* scale; const float* shift;     // [64][10], [64], [64]
    float vx, vy, vz, offx, offy, offz;                         // voxel size, voxel/2 + range_min
};
constexpr int PFN_NCOEF = 11;                                   // A0..A3, C0..C2, M0..M2, shift
__device__ __forceinline__ float pfn_coef(int j, const float* wc, float sc, float sh) {
    const double s = (double)sc;
    switch (j) {
        case 0: return (float)(s * ((double)wc[0] + (double)wc[4] + (double)wc[7]));
        case 1: return (float)(s * ((double)wc[1] + (double)wc[5] + (double)wc[8]));
        case 2: return (float)(s * ((double)wc[2] + (double)wc[6] + (double)wc[9]));
        case 3: return (float)(s * (double)wc[3]);
        case 4: return (float)(s * (double)wc[0]);
        case 5: return (float)(s * (double)wc[1]);
        case 6: return (float)(s * (double)wc[2]);
        case 7: return (float)(-s * (double)wc[4]);
        case 8: return (float)(-s * (double)wc[5]);
        case 9: return (float)(-s * (double)wc[6]);
        default: return sh;
    }
}
// coefficient table [PFN_NCOEF][64] (pair-interleaved when read as float2) for the constant bank
__global__ void pfn_coef_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                const float* __restrict__ shift, float* __restrict__ out) {
    const int c = threadIdx.x;
    if (c >= 64) return;
    float wc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) wc[j] = w[c * 10 + j];
#pragma unroll
    for (int j = 0; j < PFN_NCOEF; ++j) out[j * 64 + c] = pfn_coef(j, wc, scale[c], shift[c]);
}
// One table per SLOT: concurrently active engines (different workspaces / streams / captured graphs) each get their own
// copy, so one engine's upload cannot change the coefficients under another engine's running kernel (pfn_slot_for below).
constexpr int PFN_SLOTS = 4;
__constant__ float2 c_pfn_k[PFN_SLOTS][PFN_NCOEF][32];

__device__ __forceinline__ void pillar_centre(const PfnParams& pp, int cz, int cy, int cx, float& ctrx, float& ctry,
                                              float& ctrz) {
    ctrx = __fadd_rn(__fmul_rn((float)cx, pp.vx), pp.offx);     // pillar_vfe.py:127-131, fp32, unfused
    ctry = __fadd_rn(__fmul_rn((float)cy, pp.vy), pp.offy);
    ctrz = __fadd_rn(__fmul_rn((float)cz, pp.vz), pp.offz);
}

// PFN of one pillar for NP channel pairs.  K(j, c) = coefficient j of local pair c (f2); pt(k) = slot k (1 <= k < n).
// Zero-padded slots (n < max_pts) contribute relu(shift) to the max (their 10 features are all zero).
template <int NP, class CoefFn, class PointFn>
__device__ __forceinline__ void pfn_eval(CoefFn K, PointFn pt, const float4 p0, int n, int max_pts, float ctrx,
                                         float ctry, float ctrz, f2 (&best)[NP]) {
    // mean of the centred coordinates, accumulated in fp64 (order-independent to 2^-53: the slot order of the fused
    // path is the arrival order) and rounded to fp32 once
    double sx = (double)__fsub_rn(p0.x, ctrx), sy = (double)__fsub_rn(p0.y, ctry), sz = (double)__fsub_rn(p0.z, ctrz);
    for (int k = 1; k < n; ++k) {
        const float4 p = pt(k);
        sx += (double)__fsub_rn(p.x, ctrx); sy += (double)__fsub_rn(p.y, ctry); sz += (double)__fsub_rn(p.z, ctrz);
    }
    const double inv_n = 1.0 / (double)n;
    const float mx = (float)(sx * inv_n), my = (float)(sy * inv_n), mz = (float)(sz * inv_n);
    f2 B[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const f2 sh = K(10, c);
        f2 b = fma2(K(4, c), f2{ctrx, ctrx}, sh);
        b = fma2(K(5, c), f2{ctry, ctry}, b);
        b = fma2(K(6, c), f2{ctrz, ctrz}, b);
        b = fma2(K(7, c), f2{mx, mx}, b);
        b = fma2(K(8, c), f2{my, my}, b);
        B[c] = fma2(K(9, c), f2{mz, mz}, b);
        best[c].x = n < max_pts ? fmaxf(sh.x, 0.f) : 0.f;
        best[c].y = n < max_pts ? fmaxf(sh.y, 0.f) : 0.f;
    }
    for (int k = 0; k < n; ++k) {
        const float4 p = k == 0 ? p0 : pt(k);
        const float qx = __fsub_rn(p.x, ctrx), qy = __fsub_rn(p.y, ctry), qz = __fsub_rn(p.z, ctrz);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            f2 y = fma2(K(0, c), f2{qx, qx}, B[c]);
            y = fma2(K(1, c), f2{qy, qy}, y);
            y = fma2(K(2, c), f2{qz, qz}, y);
            y = fma2(K(3, c), f2{p.w, p.w}, y);
            best[c].x = fmaxf(best[c].x, y.x);                  // ReLU folded into the max (best >= 0)
            best[c].y = fmaxf(best[c].y, y.y);
        }
    }
}

// bf16 (hi [+ lo]) store of NP channel pairs = NP/4 16-byte vectors at dst (channel offset already applied)
template <int NP>
__device__ __forceinline__ void pfn_store(const f2 (&best)[NP], __nv_bfloat16* dst, long lo_off) {
    uint32_t hi[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) hi[c] = pack_bf16(best[c].x, best[c].y);
#pragma unroll
    for (int q = 0; q < NP / 4; ++q)
        reinterpret_cast<uint4*>(dst)[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    if (lo_off != 0) {
#pragma unroll
        for (int q = 0; q < NP / 4; ++q) {
            uint32_t lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * q + e;
                lo[e] = pack_bf16(best[c].x - bf16_lo(hi[c]), best[c].y - bf16_hi(hi[c]));
            }
            reinterpret_cast<uint4*>(dst + lo_off)[q] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// Staged path: 8 lanes ("group") cooperate on one voxel, lane sub = lane & 7 owns channels 8*sub .. 8*sub+7; the
// coefficient table sits in shared memory (s_k[j][channel], filled by fill_pfn_table at kernel start).
__device__ __forceinline__ void fill_pfn_table(float (*s_k)[64], const PfnParams& pp) {
    for (int c = threadIdx.x; c < 64; c += blockDim.x) {
        float wc[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) wc[j] = __ldg(pp.w + c * 10 + j);
        const float sc = __ldg(pp.scale + c), sh = __ldg(pp.shift + c);
#pragma unroll
        for (int j = 0; j < PFN_NCOEF; ++j) s_k[j][c] = pfn_coef(j, wc, sc, sh);
    }
    __syncthreads();
}
template <class PointFn>
__device__ __forceinline__ void pfn_group_store(PointFn pt, int n, int max_pts, int a, int cz, int cy, int cx,
                                                const PfnParams& pp, const float (*s_k)[64], const CanvasGeom& cg,
                                                __nv_bfloat16* canvas, long lo_off, int sub, long* dirty_slot) {
    float ctrx, ctry, ctrz;
    pillar_centre(pp, cz, cy, cx, ctrx, ctry, ctrz);
    f2 best[4];
    pfn_eval<4>([&](int j, int c) { const float2 v = *reinterpret_cast<const float2*>(&s_k[j][8 * sub + 2 * c]); return f2{v.x, v.y}; },
                pt, pt(0), n, max_pts, ctrx, ctry, ctrz, best);
    const long row = canvas_row(cg, a, cy, cx);
    pfn_store<4>(best, canvas + row * 64 + sub * 8, lo_off);
    if (dirty_slot && sub == 0) *dirty_slot = row;
}

// Fused path: ONE THREAD per pillar and channel group of 2*NP channels (group index GRP); the coefficients live in
// constant memory, so every FFMA2 takes them as a uniform-register operand (no weight registers, no shared-memory
// traffic).  Pillars are visited class by class (same point count inside a warp) so warps do not diverge on n.
template <int NP, int GRP, int SLOT, class PointFn>
__device__ __forceinline__ void pfn_thread_store(PointFn pt, const float4 p0, int n, int max_pts, int a, int cz,
                                                 int cy, int cx, const PfnParams& pp, const CanvasGeom& cg,
                                                 __nv_bfloat16* canvas, long lo_off, long* dirty_slot) {
    float ctrx, ctry, ctrz;
    pillar_centre(pp, cz, cy, cx, ctrx, ctry, ctrz);
    f2 best[NP];
    pfn_eval<NP>([&](int j, int c) { const float2 v = c_pfn_k[SLOT][j][GRP * NP + c]; return f2{v.x, v.y}; }, pt, p0, n,
                 max_pts, ctrx, ctry, ctrz, best);
    const long row = canvas_row(cg, a, cy, cx);
    pfn_store<NP>(best, canvas + row * 64 + GRP * 2 * NP, lo_off);
    if (dirty_slot && GRP == 0) *dirty_slot = row;
}
// grp (warp-uniform: even / odd warps) -> compile-time channel half, so the constant-bank addresses are immediates
template <int NP, int SLOT, class PointFn>
__device__ __forceinline__ void pfn_thread_dispatch(int grp, PointFn pt, const float4 p0, int n, int max_pts, int a,
                                                    int cz, int cy, int cx, const PfnParams& pp, const CanvasGeom& cg,
                                                    __nv_bfloat16* canvas, long lo_off, long* dirty_slot) {
    static_assert(NP == 16, "two 32-channel halves per pillar");
    if (grp == 0) pfn_thread_store<NP, 0, SLOT>(pt, p0, n, max_pts, a, cz, cy, cx, pp, cg, canvas, lo_off, dirty_slot);
    else          pfn_thread_store<NP, 1, SLOT>(pt, p0, n, max_pts, a, cz, cy, cx, pp, cg, canvas, lo_off, dirty_slot);
}

// K4a: emit reference-format voxel tensors (8 lanes per voxel, voxels of all agents flattened) ---------
__global__ void __launch_bounds__(256) vox_emit_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                       int max_pts, float4* __restrict__ voxels,
                                                       int4* __restrict__ coords, int* __restrict__ num_points) {
    __shared__ int s_base[CB_MAX_AGENTS + 1];
    if (threadIdx.x == 0) {
        int b = 0;
        for (int a = 0; a < ao.n_agents; ++a) { s_base[a] = b; b += ws.nvox[a]; }
        s_base[ao.n_agents] = b;
    }
    __syncthreads();
    const int total = s_base[ao.n_agents];
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const int gg = blockIdx.x * 32 + grp, ng = gridDim.x * 32;
    for (int row = gg; row < total; row += ng) {
        int a = 0;
        while (a + 1 < ao.n_agents && row >= s_base[a + 1]) ++a;
        const int v = row - s_base[a];
        const int4 m = ws.vox_meta[(long)a * ws.vcap + v];
        const int n = m.y < max_pts ? m.y : max_pts;
        const float4* vp = ws.vp + ao.off[a] + m.x;
        for (int k = sub; k < max_pts; k += 8)
            voxels[(long)row * max_pts + k] = k < n ? vp[k] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (sub == 0) {
            coords[row] = make_int4(a, (m.z >> 24) & 0xFF, (m.z >> 12) & 0xFFF, m.z & 0xFFF);
            num_points[row] = n;
        }
    }
}

// =============================================================================================================
// v2 front-end of the fused points -> canvas path.  The canvas result does not depend on the ORDER of the voxels, only
// on which points each cell keeps (its first max_pts by index) and on which cells survive the max_voxels cap (the first
// max_voxels by first appearance), so the ordered scan over the points is only run when a cap is actually hit:
//   V1 (points)  cell id, count[cell]++ (arrival position), first[cell] = min index, the point goes straight into the
//                cell's slot array slots[cell][pos]; non-empty cells per agent are counted on the way
//   V2a/b        ONLY for agents with more than max_voxels non-empty cells (else the CTAs exit at once): ordered
//                per-chunk leader counts -> leader ranks -> cells past the cap are marked refused
//   V3 (cells)   non-empty cells -> work items {cell, n, x, y} bucketed by point-count class (CTA-aggregated atomics);
//                cells with more than max_pts points get an index list and are re-filled by V4a/b in index order
//   V4a/b        ONLY when such cells exist: index list fill, rank by index, first max_pts points re-written to the slots
//   V5           PFN + canvas store, one thread per (pillar, channel group)
// Slot order inside a pillar is the arrival order (run-dependent) for cells with <= max_pts points; the max is order-
// independent and the mean is accumulated in fp64 from the centred fp32 coordinates (exact: <= 32 addends of 24-bit
// mantissas within a 2^12 range), so the canvas is bit-reproducible from run to run.
// =============================================================================================================
constexpr int V2_BIG = 1 << 30;          // count[cell] >= V2_BIG: overflow cell being re-counted by V4a
constexpr int V2_S0 = 4;                // points per cell in the compact first-tier slot array
constexpr int V2_SCAL = 8;               // scal[0..4] class fill, [5] overflow list top, [6] any overflow cell

struct Vox2Ws {
    int* count;           // [n_agents*ncell]   points per cell; -1 = refused by the max_voxels cap
    int* scal;            // [V2_SCAL + n_agents]  scalars above, then non-empty cells per agent
    int* first;           // [n_agents*ncell]   min point index per cell; later the overflow-list base of big cells
    int* cellid;          // [sum_P]
    int* list;            // [sum_P]            index lists of the cells with more than max_pts points
    int* chunk_tot;       // [n_agents][max_chunks]
    int2* items;          // [NCLS][item_cap]   {global cell, n | x << 6 | y << 18}
    float* coef;          // [PFN_NCOEF][64]
    float4* slots;        // [n_agents*ncell][V2_S0]    the first V2_S0 points of every cell (64-byte records: the bulk of
                          //                            the scattered traffic stays inside a compact region)
    float4* slots2;       // [n_agents*ncell][S2]       points V2_S0 .. max_pts-1 of the few cells that have them
    int ncell, S0, S2, max_chunks, item_cap;
};

__device__ __forceinline__ float4* slot_ptr(const Vox2Ws& ws, long gc, int k) {
    return k < ws.S0 ? ws.slots + gc * ws.S0 + k : ws.slots2 + gc * ws.S2 + (k - ws.S0);
}

__device__ __forceinline__ int find_agent_bs(const AoView& ao, int i) {      // largest a with off[a] <= i
    int lo = 0, hi = ao.n_agents;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ao.off[mid] <= i) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(256) vox2_assign_kernel(const float4* __restrict__ pts,
                                                          const __grid_constant__ AgentOffsets ao_val,
                                                          const int* __restrict__ off_dev, const Geom g,
                                                          const Vox2Ws ws, int max_pts, int use_first) {
    pdl_launch_dependents();
    pdl_wait();
    const AoView ao = ao_view(ao_val, off_dev);
    const int total = ao.off[ao.n_agents];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int a_lo = find_agent_bs(ao, blockIdx.x * 256);          // CTA-uniform
    const int a_hi = find_agent_bs(ao, min(blockIdx.x * 256 + 255, total - 1));
    int a = a_lo, opened = 0;
    if (i < total) {
        if (a_lo != a_hi) a = find_agent_bs(ao, i);
        const float4 p = __ldg(pts + i);
        const float fx = floorf(__fdiv_rn(__fsub_rn(p.x, g.r0), g.v0));      // as vox_assign_kernel (serial generator)
        const float fy = floorf(__fdiv_rn(__fsub_rn(p.y, g.r1), g.v1));
        const float fz = floorf(__fdiv_rn(__fsub_rn(p.z, g.r2), g.v2));
        int cell = -1;
        if (fx >= 0.f && fx < (float)g.gx && fy >= 0.f && fy < (float)g.gy && fz >= 0.f && fz < (float)g.gz) {
            cell = ((int)fz * g.gy + (int)fy) * g.gx + (int)fx;
            const long gc = (long)a * ws.ncell + cell;
            const int pos = atomicAdd(ws.count + gc, 1);
            if (use_first) atomicMin(ws.first + gc, i - ao.off[a]);  // only the max_voxels cap needs it
            if (pos < max_pts) *slot_ptr(ws, gc, pos) = p;
            opened = pos == 0;
        }
        ws.cellid[i] = cell;
    }
    // non-empty cells per agent: one atomic per CTA when the CTA lies inside one agent (the usual case)
    if (a_lo == a_hi) {
        const int c = __syncthreads_count(opened);
        if (threadIdx.x == 0 && c > 0) atomicAdd(ws.scal + V2_SCAL + a_lo, c);
    } else if (opened) {
        atomicAdd(ws.scal + V2_SCAL + a, 1);
    }
}

// V2a/b: max_voxels cap.  "Leader" = the first point of its cell; voxel rank = number of leaders before it.
__device__ __forceinline__ int cap_leaders(const AoView& ao, const Vox2Ws& ws, int a, int chunk, long (&gc)[4]) {
    const int p0 = ao.off[a], np = ao.off[a + 1] - p0;
    int mask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = chunk * CHUNK + threadIdx.x * 4 + j;
        gc[j] = -1;
        if (i < np) {
            const int cell = ws.cellid[p0 + i];
            if (cell >= 0) {
                gc[j] = (long)a * ws.ncell + cell;
                if (ws.first[gc[j]] == i) mask |= 1 << j;
            }
        }
    }
    return mask;
}
__global__ void __launch_bounds__(256) vox2_cap_count_kernel(const __grid_constant__ AgentOffsets ao_val,
                                                             const int* __restrict__ off_dev, const Vox2Ws ws,
                                                             int max_voxels) {
    pdl_launch_dependents();
    pdl_wait();
    const AoView ao = ao_view(ao_val, off_dev);
    const int a = blockIdx.y, chunk = blockIdx.x;
    if (ws.scal[V2_SCAL + a] <= max_voxels) return;                 // uniform: this agent is under the cap
    long gc[4];
    const int mask = cap_leaders(ao, ws, a, chunk, gc);
    int ev, ec, tv, tc;
    block_scan2(__popc(mask), 0, ev, ec, tv, tc);
    if (threadIdx.x == 0) ws.chunk_tot[(long)a * ws.max_chunks + chunk] = tv;
}
__global__ void __launch_bounds__(256) vox2_cap_refuse_kernel(const __grid_constant__ AgentOffsets ao_val,
                                                              const int* __restrict__ off_dev, const Vox2Ws ws,
                                                              int max_voxels) {
    pdl_launch_dependents();
    pdl_wait();
    const AoView ao = ao_view(ao_val, off_dev);
    const int a = blockIdx.y, chunk = blockIdx.x;
    if (ws.scal[V2_SCAL + a] <= max_voxels) return;
    int part = 0;
    for (int c = threadIdx.x; c < chunk; c += 256) part += ws.chunk_tot[(long)a * ws.max_chunks + c];
    int ev, ec, before, tc;
    block_scan2(part, 0, ev, ec, before, tc);                       // leaders in the preceding chunks
    long gc[4];
    const int mask = cap_leaders(ao, ws, a, chunk, gc);
    int tv;
    block_scan2(__popc(mask), 0, ev, ec, tv, tc);
    int rank = before + ev;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (mask & (1 << j)) {
            if (rank >= max_voxels) ws.count[gc[j]] = -1;           // refused: max_voxels reached
            ++rank;
        }
}

// V3: one thread per 4 consecutive cells
__global__ void __launch_bounds__(256) vox2_cells_kernel(const Vox2Ws ws, unsigned total_cells, int gx, int max_pts) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_cnt[NCLS], s_base[NCLS];
    if (threadIdx.x < NCLS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned g0 = (blockIdx.x * 256u + threadIdx.x) * 4u;
    int cnt[4] = {0, 0, 0, 0};
    if (g0 + 3 < total_cells) {
        const int4 v = *reinterpret_cast<const int4*>(ws.count + g0);
        cnt[0] = v.x; cnt[1] = v.y; cnt[2] = v.z; cnt[3] = v.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (g0 + j < total_cells) cnt[j] = ws.count[g0 + j];
    }
    int lpos[4], cls[4], word[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) lpos[j] = -1;
    if (cnt[0] > 0 || cnt[1] > 0 || cnt[2] > 0 || cnt[3] > 0) {
        unsigned cell = g0 % (unsigned)ws.ncell;                   // nz == 1 on this path: cell = y * gx + x
        unsigned cy = cell / (unsigned)gx, cx = cell - cy * (unsigned)gx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (cnt[j] > 0) {
                const int n = cnt[j] < max_pts ? cnt[j] : max_pts;
                word[j] = n | ((int)cx << 6) | ((int)cy << 18);
                cls[j] = n <= 1 ? 0 : (n == 2 ? 1 : (n <= 4 ? 2 : (n <= 8 ? 3 : 4)));
                lpos[j] = atomicAdd(&s_cnt[cls[j]], 1);
                if (cnt[j] > max_pts) {                             // keep the first max_pts BY INDEX: V4a/b redo the slots
                    ws.first[g0 + j] = atomicAdd(ws.scal + 5, cnt[j]);
                    ws.count[g0 + j] = V2_BIG;
                    ws.scal[6] = 1;
                }
            }
            if (++cx == (unsigned)gx) { cx = 0; if (++cy * (unsigned)gx == (unsigned)ws.ncell) cy = 0; }   // next row / next agent
        }
    }
    __syncthreads();
    if (threadIdx.x < NCLS && s_cnt[threadIdx.x] > 0)
        s_base[threadIdx.x] = atomicAdd(ws.scal + threadIdx.x, s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (lpos[j] >= 0)
            ws.items[(long)cls[j] * ws.item_cap + s_base[cls[j]] + lpos[j]] = make_int2((int)(g0 + j), word[j]);
}

// V4a/b: cells with more than max_pts points (grid-stride; the whole grid exits at once when there are none)
__global__ void __launch_bounds__(256) vox2_big_fill_kernel(const __grid_constant__ AgentOffsets ao_val,
                                                            const int* __restrict__ off_dev, const Vox2Ws ws) {
    pdl_launch_dependents();
    pdl_wait();
    const AoView ao = ao_view(ao_val, off_dev);
    if (ws.scal[6] == 0) return;
    const int total = ao.off[ao.n_agents];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int cell = ws.cellid[i];
        if (cell < 0) continue;
        const int a = find_agent_bs(ao, i);
        const long gc = (long)a * ws.ncell + cell;
        if (ws.count[gc] < V2_BIG) continue;
        const int pos = atomicAdd(ws.count + gc, 1) - V2_BIG;
        ws.list[ws.first[gc] + pos] = i - ao.off[a];
    }
}
__global__ void __launch_bounds__(256) vox2_big_rank_kernel(const float4* __restrict__ pts,
                                                            const __grid_constant__ AgentOffsets ao_val,
                                                            const int* __restrict__ off_dev, const Vox2Ws ws,
                                                            int max_pts) {
    pdl_launch_dependents();
    pdl_wait();
    const AoView ao = ao_view(ao_val, off_dev);
    if (ws.scal[6] == 0) return;
    const int total = ao.off[ao.n_agents];
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int cell = ws.cellid[i];
        if (cell < 0) continue;
        const int a = find_agent_bs(ao, i);
        const long gc = (long)a * ws.ncell + cell;
        const int c = ws.count[gc];
        if (c < V2_BIG) continue;
        const int cnt = c - V2_BIG, li = i - ao.off[a];
        const int* lst = ws.list + ws.first[gc];
        int rank = 0;
        for (int k = 0; k < cnt; ++k) rank += (lst[k] < li) ? 1 : 0;
        if (rank < max_pts) *slot_ptr(ws, gc, rank) = __ldg(pts + i);
    }
}

// V5: PFN + scatter, thread per (pillar, channel group of 2*NP channels); the 64/(2*NP) groups of a pillar are
// consecutive warps of one CTA.  Work items are taken class-major so the pillars of a warp share a point-count class.
template <int NP, int SLOT>
__global__ void __launch_bounds__(256, 2) vox2_pfn_kernel(const Vox2Ws ws, int max_pts,
                                                                          const PfnParams pp, const CanvasGeom cg,
                                                                          __nv_bfloat16* canvas, long lo_off,
                                                                          long* dirty_rows, int* dirty_count) {
    constexpr int NG = 32 / NP;                                    // channel groups per pillar (2 or 4)
    constexpr int PW = 8 / NG;                                     // pillar-warps per CTA (4 or 2)
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int s_seg[NCLS + 1];
    if (threadIdx.x == 0) {
        int t = 0;
        for (int c = 0; c < NCLS; ++c) { s_seg[c] = t; t += ws.scal[c]; }
        s_seg[NCLS] = t;
        if (blockIdx.x == 0 && dirty_count) *dirty_count = t;
    }
    __syncthreads();
    const int total = s_seg[NCLS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp % NG;
    const int slot = (blockIdx.x * PW + warp / NG) * 32 + lane;
    const int nslot = gridDim.x * PW * 32;
    // Two-stage software pipeline over this thread's pillars: the work item of pillar i+2 and the first point of pillar
    // i+1 (whose item arrived an iteration ago) are in flight while pillar i is computed, so neither the item -> slot
    // address dependency nor the slot load latency sits on the critical path.
    auto load_item = [&](int gidx) {
        int c = 0;
#pragma unroll
        for (int k = 1; k < NCLS; ++k) c += (gidx >= s_seg[k]) ? 1 : 0;
        return ws.items[(long)c * ws.item_cap + (gidx - s_seg[c])];
    };
    int2 it_cur = make_int2(0, 1), it_nxt = make_int2(0, 1);
    float4 p_cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot < total) it_cur = load_item(slot);
    if (slot + nslot < total) it_nxt = load_item(slot + nslot);
    if (slot < total) p_cur = ws.slots[(long)it_cur.x * ws.S0];
    for (int gidx = slot; gidx < total; gidx += nslot) {
        int2 it_nn = it_nxt;
        if (gidx + 2 * nslot < total) it_nn = load_item(gidx + 2 * nslot);
        float4 p_nxt = p_cur;
        if (gidx + nslot < total) p_nxt = ws.slots[(long)it_nxt.x * ws.S0];
        const int gcell = it_cur.x, word = it_cur.y;
        const int n = word & 63, cx = (word >> 6) & 0xFFF, cy = (word >> 18) & 0xFFF;
        const int a = gcell / ws.ncell;
        long* dslot = dirty_rows ? dirty_rows + gidx : nullptr;
        pfn_thread_dispatch<NP, SLOT>(grp, [&](int k) { return *slot_ptr(ws, gcell, k); }, p_cur, n, max_pts, a, 0, cy, cx, pp, cg,
                                    canvas, lo_off, dslot);
        it_cur = it_nxt; it_nxt = it_nn; p_cur = p_nxt;
    }
}

// PFN + scatter from reference-format voxel tensors ----------------------------------------------
__global__ void __launch_bounds__(256, 2) pfn_scatter_kernel(const float4* __restrict__ voxels,
                                                          const int4* __restrict__ coords,
                                                          const int* __restrict__ num_points, int n_rows,
                                                          const int* __restrict__ n_rows_dev, int max_pts,
                                                          const PfnParams pp, const CanvasGeom cg, int n_agents,
                                                          __nv_bfloat16* canvas, long lo_off, long* dirty_rows,
                                                          int* dirty_count) {
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = threadIdx.x >> 3;
    const int gg = blockIdx.x * 32 + grp, ng = gridDim.x * 32;
    __shared__ __align__(16) float s_k[PFN_NCOEF][64];
    fill_pfn_table(s_k, pp);
    const int rows = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
    if (dirty_count && blockIdx.x == 0 && threadIdx.x == 0) *dirty_count = rows;
    for (int v = gg; v < rows; v += ng) {
        const int4 c = __ldg(coords + v);                     // [agent, z, y, x]
        int n = __ldg(num_points + v);
        n = n < max_pts ? n : max_pts;
        const bool ok = !(n < 1 || c.x < 0 || c.x >= n_agents || c.z < 0 || c.z >= cg.ny || c.w < 0 || c.w >= cg.nx);
        if (!ok) { if (dirty_rows && sub == 0) dirty_rows[v] = -1; continue; }
        const float4* vp = voxels + (long)v * max_pts;
        pfn_group_store([&](int k) { return __ldg(vp + k); }, n, max_pts, c.x, c.y, c.z, c.w, pp, s_k, cg, canvas, lo_off,
                        sub, dirty_rows ? dirty_rows + v : nullptr);
    }
}

// zero the canvas cells written by the previous frame (instead of a full-canvas memset)
__global__ void __launch_bounds__(256) canvas_clear_kernel(__nv_bfloat16* canvas, long lo_off,
                                                           const long* __restrict__ dirty_rows,
                                                           const int* __restrict__ dirty_count, int capacity) {
    const int sub = threadIdx.x & 7;
    const int gg = blockIdx.x * 32 + (threadIdx.x >> 3), ng = gridDim.x * 32;
    int n = *dirty_count;
    n = n < capacity ? n : capacity;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int i = gg; i < n; i += 4 * ng) {                          // 4 independent row loads in flight per group
        long row[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) row[u] = (i + u * ng < n) ? __ldg(dirty_rows + i + u * ng) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (row[u] < 0) continue;
            reinterpret_cast<uint4*>(canvas + row[u] * 64)[sub] = z;
            if (lo_off != 0) reinterpret_cast<uint4*>(canvas + lo_off + row[u] * 64)[sub] = z;
        }
    }
}

// ------------------------------------------------------------------------------------------ host
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int carve(VoxWs& ws, void* base, size_t bytes, int n_agents, int sum_points, const int32_t* grid, int max_voxels,
                 size_t* clear_bytes, size_t* total_bytes) {
    const long ncell = (long)grid[0] * grid[1] * grid[2];
    const int sp = sum_points > 0 ? sum_points : 1;
    const int vcap = max_voxels < sp ? max_voxels : sp;
    const int max_chunks = (sp + CHUNK - 1) / CHUNK;          // upper bound for any single agent
    uint8_t* p = (uint8_t*)base;
    size_t o = 0;
    auto take = [&](size_t n_bytes) { void* r = p ? p + o : nullptr; o += align256(n_bytes); return r; };
    ws.first = (int*)take((size_t)n_agents * ncell * 4);
    size_t o_first_end = o;
    ws.count = (int*)take((size_t)n_agents * ncell * 4);
    ws.cursor = (int*)take((size_t)n_agents * vcap * 4);
    size_t o_clear_end = o;
    ws.cell2vox = (int*)take((size_t)n_agents * ncell * 4);
    ws.vox_off = (int*)take((size_t)n_agents * (vcap + 1) * 4);
    ws.vox_meta = (int4*)take((size_t)n_agents * vcap * 16);
    ws.cellid = (int*)take((size_t)sp * 4);
    ws.list = (int*)take((size_t)sp * 4);
    ws.nvox = (int*)take(((size_t)n_agents + 1) * 4);
    ws.chunk_tot = (int2*)take((size_t)n_agents * max_chunks * 8);
    ws.vp = (float4*)take((size_t)sp * 16);
    ws.ncell = (int)ncell;
    ws.vcap = vcap;
    ws.max_chunks = max_chunks;
    if (clear_bytes) { clear_bytes[0] = o_first_end; clear_bytes[1] = o_clear_end - o_first_end; }
    if (total_bytes) *total_bytes = o;
    if (p && o > bytes) return CB_ERR_ARG;
    return CB_OK;
}

// Shared front half (K1..K3).  pt_offset is a HOST array.
static int run_front(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                     const float* vsize, const int32_t* grid, int max_pts, int max_voxels, void* workspace,
                     size_t workspace_bytes, cudaStream_t st, AgentOffsets& ao, VoxWs& ws, Geom& g, int* n_voxels_out,
                     int* dirty_count) {
    if (n_agents < 1 || n_agents > CB_MAX_AGENTS || max_pts < 1 || max_pts > 32 || max_voxels < 1) return CB_ERR_ARG;
    if (!workspace || ((uintptr_t)points & 15)) return CB_ERR_ARG;
    if (grid[0] > 4096 || grid[1] > 4096 || grid[2] > 128) return CB_ERR_ARG;        // packed voxel coordinates
    ao.n_agents = n_agents;
    for (int i = 0; i <= n_agents; ++i) ao.off[i] = pt_offset[i];
    if (ao.off[0] != 0) return CB_ERR_ARG;
    int max_np = 0;
    for (int i = 0; i < n_agents; ++i) {
        if (ao.off[i + 1] < ao.off[i]) return CB_ERR_ARG;
        if (ao.off[i + 1] - ao.off[i] > max_np) max_np = ao.off[i + 1] - ao.off[i];
    }
    const int total = ao.off[n_agents];
    size_t clr[2];
    int rc = carve(ws, workspace, workspace_bytes, n_agents, total, grid, max_voxels, clr, nullptr);
    if (rc) return rc;
    g.r0 = range[0]; g.r1 = range[1]; g.r2 = range[2];
    g.v0 = vsize[0]; g.v1 = vsize[1]; g.v2 = vsize[2];
    g.gx = grid[0]; g.gy = grid[1]; g.gz = grid[2];
    cudaError_t e;
    e = cudaMemsetAsync(ws.first, 0x7f, clr[0], st);              if (e) return (int)e;
    e = cudaMemsetAsync(ws.count, 0, clr[1], st);                 if (e) return (int)e;
    if (total > 0) {
        vox_assign_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float4*)points, ao, g, ws);
        CB_CHECK_LAUNCH();
    }
    const int n_chunks = max_np > 0 ? (max_np + CHUNK - 1) / CHUNK : 1;
    const dim3 cgrid((unsigned)n_chunks, (unsigned)n_agents);
    vox_chunk_count_kernel<<<cgrid, 256, 0, st>>>(ao, ws);
    CB_CHECK_LAUNCH();
    vox_chunk_scan_kernel<<<n_agents, 32, 0, st>>>(ws, n_chunks, max_voxels);
    CB_CHECK_LAUNCH();
    vox_total_kernel<<<1, 32, 0, st>>>(ws, n_agents, n_voxels_out, dirty_count);
    CB_CHECK_LAUNCH();
    if (total > 0) {
        vox_chunk_assign_kernel<<<cgrid, 256, 0, st>>>(ao, ws, g, max_voxels);
        CB_CHECK_LAUNCH();
        vox_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(ao, ws);
        CB_CHECK_LAUNCH();
        vox_rank_gather_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float4*)points, ao, ws, max_pts);
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

static int carve2(Vox2Ws& ws, void* base, size_t bytes, int n_agents, int sum_points, const int32_t* grid, int max_pts,
                  int max_voxels, size_t* clear_bytes, size_t* total_bytes) {
    const long ncell = (long)grid[0] * grid[1] * grid[2];
    const int sp = sum_points > 0 ? sum_points : 1;
    const long per_agent = max_voxels < ncell ? max_voxels : ncell;
    const long cap = (long)n_agents * per_agent < sp ? (long)n_agents * per_agent : sp;
    const int max_chunks = (sp + CHUNK - 1) / CHUNK;
    uint8_t* p = (uint8_t*)base;
    size_t o = 0;
    auto take = [&](size_t n_bytes) { void* r = p ? p + o : nullptr; o += align256(n_bytes); return r; };
    ws.count = (int*)take((size_t)n_agents * ncell * 4);
    ws.scal = (int*)take((size_t)(V2_SCAL + n_agents) * 4);
    size_t o_zero_end = o;
    ws.first = (int*)take((size_t)n_agents * ncell * 4);
    size_t o_first_end = o;
    ws.cellid = (int*)take((size_t)sp * 4);
    ws.list = (int*)take((size_t)sp * 4);
    ws.chunk_tot = (int*)take((size_t)n_agents * max_chunks * 4);
    ws.items = (int2*)take((size_t)NCLS * cap * 8);
    ws.coef = (float*)take((size_t)PFN_NCOEF * 64 * 4);
    ws.S0 = V2_S0;
    ws.slots = (float4*)take((size_t)n_agents * ncell * ws.S0 * 16);
    ws.S2 = max_pts > ws.S0 ? ((max_pts - ws.S0 + 1) & ~1) : 0;   // 32-byte aligned second-tier records
    ws.slots2 = (float4*)take((size_t)n_agents * ncell * ws.S2 * 16);
    ws.ncell = (int)ncell;
    ws.max_chunks = max_chunks;
    ws.item_cap = (int)cap;
    if (clear_bytes) { clear_bytes[0] = o_zero_end; clear_bytes[1] = o_first_end - o_zero_end; }
    if (total_bytes) *total_bytes = o;
    if (p && o > bytes) return CB_ERR_ARG;
    return CB_OK;
}

// v2 front half (V1..V4) of the fused path.  Offsets come either from a HOST array (pt_offset, baked into the kernel
// arguments) or from a DEVICE array (off_dev; then `total_cap` / `agent_cap` bound the grid sizes and the launch
// sequence depends only on those capacities, so a captured CUDA graph serves every frame).
static int run_front2(const float* points, const int32_t* pt_offset, const int32_t* off_dev, int total_cap, int agent_cap,
                      int n_agents, const float* range,
                      const float* vsize, const int32_t* grid, int max_pts, int max_voxels, void* workspace,
                      size_t workspace_bytes, cudaStream_t st, Vox2Ws& ws, const PfnParams& pp, int pfn_slot) {
    if (n_agents < 1 || n_agents > CB_MAX_AGENTS || max_pts < 1 || max_pts > 32 || max_voxels < 1) return CB_ERR_ARG;
    if (!workspace || ((uintptr_t)points & 15)) return CB_ERR_ARG;
    if (grid[0] > 4096 || grid[1] > 4096 || grid[2] != 1) return CB_ERR_ARG;          // packed work items; nz == 1
    AgentOffsets ao; Geom g;
    ao.n_agents = n_agents;
    int max_np = 0, total = 0, may_cap = 0;
    if (off_dev) {
        if (total_cap < 0 || agent_cap < 0) return CB_ERR_ARG;
        for (int i = 0; i <= n_agents; ++i) ao.off[i] = 0;
        total = total_cap; max_np = agent_cap < total_cap ? agent_cap : total_cap;
        may_cap = max_np > max_voxels;
    } else {
        for (int i = 0; i <= n_agents; ++i) ao.off[i] = pt_offset[i];
        if (ao.off[0] != 0) return CB_ERR_ARG;
        for (int i = 0; i < n_agents; ++i) {
            if (ao.off[i + 1] < ao.off[i]) return CB_ERR_ARG;
            if (ao.off[i + 1] - ao.off[i] > max_np) max_np = ao.off[i + 1] - ao.off[i];
        }
        total = ao.off[n_agents];
        // an agent with no more points than max_voxels can never hit the voxel cap: then `first` and V2a/b are not needed
        for (int i = 0; i < n_agents; ++i) may_cap |= (ao.off[i + 1] - ao.off[i]) > max_voxels;
    }
    size_t clr[2];
    int rc = carve2(ws, workspace, workspace_bytes, n_agents, total, grid, max_pts, max_voxels, clr, nullptr);
    if (rc) return rc;
    if ((long)n_agents * ws.ncell >= (1L << 31)) return CB_ERR_ARG;
    g.r0 = range[0]; g.r1 = range[1]; g.r2 = range[2];
    g.v0 = vsize[0]; g.v1 = vsize[1]; g.v2 = vsize[2];
    g.gx = grid[0]; g.gy = grid[1]; g.gz = grid[2];
    // PFN coefficient table -> constant bank first, so that the kernel chain below is kernel -> kernel only (programmatic
    // dependent launches: the launch latency of kernel n+1 overlaps the tail of kernel n)
    cudaError_t e;
    pfn_coef_kernel<<<1, 64, 0, st>>>(pp.w, pp.scale, pp.shift, ws.coef);
    CB_CHECK_LAUNCH();
    e = cudaMemcpyToSymbolAsync(c_pfn_k, ws.coef, sizeof(float) * PFN_NCOEF * 64,
                                sizeof(float) * PFN_NCOEF * 64 * (size_t)pfn_slot, cudaMemcpyDeviceToDevice, st);
    if (e) return (int)e;
    e = cudaMemsetAsync(ws.count, 0, clr[0], st);                 if (e) return (int)e;
    if (may_cap) { e = cudaMemsetAsync(ws.first, 0x7f, clr[1], st); if (e) return (int)e; }
    if (total > 0) {
        const unsigned pblocks = (unsigned)((total + 255) / 256);
        const unsigned sblocks = pblocks < 148u * 4u ? pblocks : 148u * 4u;
        vox2_assign_kernel<<<pblocks, 256, 0, st>>>((const float4*)points, ao, off_dev, g, ws, max_pts, may_cap);
        CB_CHECK_LAUNCH();
        if (may_cap) {
            const dim3 cgrid((unsigned)((max_np + CHUNK - 1) / CHUNK), (unsigned)n_agents);
            e = launch_pdl(vox2_cap_count_kernel, cgrid, dim3(256), 0, st, ao, off_dev, ws, max_voxels);   if (e) return (int)e;
            e = launch_pdl(vox2_cap_refuse_kernel, cgrid, dim3(256), 0, st, ao, off_dev, ws, max_voxels);  if (e) return (int)e;
        }
        const unsigned total_cells = (unsigned)n_agents * (unsigned)ws.ncell;
        e = launch_pdl(vox2_cells_kernel, dim3((total_cells + 1023u) / 1024u), dim3(256), 0, st, ws, total_cells, (int)grid[0],
                       max_pts);                                                                       if (e) return (int)e;
        e = launch_pdl(vox2_big_fill_kernel, dim3(sblocks), dim3(256), 0, st, ao, off_dev, ws);       if (e) return (int)e;
        e = launch_pdl(vox2_big_rank_kernel, dim3(sblocks), dim3(256), 0, st, (const float4*)points, ao, off_dev, ws, max_pts);
        if (e) return (int)e;
    }
    return CB_OK;
}

static CanvasGeom make_canvas_geom(int canvas_agents, int ny, int nx) {
    CanvasGeom cg;
    cg.ny = ny; cg.nx = nx;
    cg.Hq = (ny + 1) / 2 + 2;
    cg.Wq = (nx + 1) / 2 + 2;
    cg.plane_rows = (long)canvas_agents * cg.Hq * cg.Wq;
    return cg;
}

static PfnParams make_pfn(const float* w, const float* scale, const float* shift, const float* vsize,
                          const float* center_off) {
    PfnParams pp;
    pp.w = w; pp.scale = scale; pp.shift = shift;
    pp.vx = vsize[0]; pp.vy = vsize[1]; pp.vz = vsize[2];
    pp.offx = center_off[0]; pp.offy = center_off[1]; pp.offz = center_off[2];
    return pp;
}

}  // namespace cb

extern "C" size_t cb_voxelize_workspace_bytes(int n_agents, int sum_points, const int32_t* grid, int max_voxels) {
    cb::VoxWs ws;
    size_t total = 0, total2 = 0;
    cb::carve(ws, nullptr, 0, n_agents, sum_points, grid, max_voxels, nullptr, &total);
    if (grid[2] == 1) {                                            // the fused canvas path (v2 front-end, 32 slots per cell)
        cb::Vox2Ws ws2;
        cb::carve2(ws2, nullptr, 0, n_agents, sum_points, grid, 32, max_voxels, nullptr, &total2);
    }
    return total > total2 ? total : total2;
}

extern "C" int cb_voxelize(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                           const float* vsize, const int32_t* grid, int max_pts, int max_voxels, float* voxels,
                           int32_t* coords, int32_t* num_points, int32_t* n_voxels, void* workspace,
                           size_t workspace_bytes, void* stream) {
    using namespace cb;
    cudaStream_t st = (cudaStream_t)stream;
    AgentOffsets ao; VoxWs ws; Geom g;
    int rc = run_front(points, pt_offset, n_agents, range, vsize, grid, max_pts, max_voxels, workspace,
                       workspace_bytes, st, ao, ws, g, n_voxels, nullptr);
    if (rc) return rc;
    vox_emit_kernel<<<148 * 4, 256, 0, st>>>(ao, ws, max_pts, (float4*)voxels, (int4*)coords, num_points);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

// Constant-table slot of a caller, keyed by its workspace (one per engine): the same workspace always gets the same slot;
// a new workspace takes a free slot or, when all PFN_SLOTS are owned, the least recently used one.  Up to PFN_SLOTS engines
// can therefore run (or replay captured graphs) concurrently on different streams without sharing a table; within one
// stream the upload and the kernel are ordered, so sequential use is always safe.
static int pfn_slot_for(const void* workspace) {
    static std::mutex mu;
    static const void* owner[cb::PFN_SLOTS] = {};
    static unsigned long stamp[cb::PFN_SLOTS] = {};
    static unsigned long tick = 0;
    std::lock_guard<std::mutex> lock(mu);
    int lru = 0;
    for (int i = 0; i < cb::PFN_SLOTS; ++i) {
        if (owner[i] == workspace) { stamp[i] = ++tick; return i; }
        if (stamp[i] < stamp[lru]) lru = i;
    }
    owner[lru] = workspace;
    stamp[lru] = ++tick;
    return lru;
}

static int points_to_canvas_impl(const float* points, const int32_t* pt_offset, const int32_t* off_dev, int total_cap,
                                 int agent_cap, int n_agents, const float* range,
                                 const float* vsize, const int32_t* grid, int max_pts, int max_voxels,
                                 const float* w, const float* scale, const float* shift,
                                 const float* center_off, int canvas_agents, void* canvas_ps,
                                 int64_t lo_off, int64_t* dirty_rows, int32_t* dirty_count,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    using namespace cb;
    if (grid[2] != 1 || !canvas_ps || canvas_agents < n_agents) return CB_ERR_ARG;   // nz == 1 (point_pillar_scatter.py:13)
    cudaStream_t st = (cudaStream_t)stream;
    const CanvasGeom cg = make_canvas_geom(canvas_agents, grid[1], grid[0]);
    if (cg.plane_rows * 4 >= (1L << 31)) return CB_ERR_ARG;
    const PfnParams pp = make_pfn(w, scale, shift, vsize, center_off);
    cudaError_t ce;
    Vox2Ws ws2;
    const int pfn_slot = pfn_slot_for(workspace);
    int rc2 = run_front2(points, pt_offset, off_dev, total_cap, agent_cap, n_agents, range, vsize, grid, max_pts, max_voxels,
                         workspace, workspace_bytes, st, ws2, pp, pfn_slot);
    if (rc2) return rc2;
    auto go = [&](auto kern) {
        return launch_pdl(kern, dim3(148 * 2), dim3(256), 0, st, ws2, max_pts, pp, cg, (__nv_bfloat16*)canvas_ps, (long)lo_off,
                          (long*)dirty_rows, (int*)dirty_count);
    };
    static_assert(PFN_SLOTS == 4, "one instantiation per slot below");
    switch (pfn_slot) {
        case 0: ce = go(vox2_pfn_kernel<16, 0>); break;
        case 1: ce = go(vox2_pfn_kernel<16, 1>); break;
        case 2: ce = go(vox2_pfn_kernel<16, 2>); break;
        default: ce = go(vox2_pfn_kernel<16, 3>); break;
    }
    if (ce) return (int)ce;
    return CB_OK;
}

extern "C" int cb_points_to_canvas(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                                   const float* vsize, const int32_t* grid, int max_pts, int max_voxels,
                                   const float* w, const float* scale, const float* shift,
                                   const float* center_off, int canvas_agents, void* canvas_ps,
                                   int64_t lo_off, int64_t* dirty_rows, int32_t* dirty_count,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    if (!pt_offset) return CB_ERR_ARG;
    return points_to_canvas_impl(points, pt_offset, nullptr, 0, 0, n_agents, range, vsize, grid, max_pts, max_voxels, w, scale,
                                 shift, center_off, canvas_agents, canvas_ps, lo_off, dirty_rows, dirty_count, workspace,
                                 workspace_bytes, stream);
}

extern "C" int cb_points_to_canvas_dev(const float* points, const int32_t* pt_offset_dev, int n_agents, int point_capacity,
                                       int agent_capacity, const float* range,
                                       const float* vsize, const int32_t* grid, int max_pts, int max_voxels,
                                       const float* w, const float* scale, const float* shift,
                                       const float* center_off, int canvas_agents, void* canvas_ps,
                                       int64_t lo_off, int64_t* dirty_rows, int32_t* dirty_count,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    if (!pt_offset_dev) return CB_ERR_ARG;
    return points_to_canvas_impl(points, nullptr, pt_offset_dev, point_capacity, agent_capacity, n_agents, range, vsize, grid,
                                 max_pts, max_voxels, w, scale, shift, center_off, canvas_agents, canvas_ps, lo_off, dirty_rows,
                                 dirty_count, workspace, workspace_bytes, stream);
}

namespace cb {
struct I32Pack { int n; int v[CB_MAX_AGENTS + 1]; };
__global__ void upload_i32_kernel(const __grid_constant__ I32Pack p, int* __restrict__ dst) {
    const int i = threadIdx.x;
    if (i < p.n) dst[i] = p.v[i];
}
}  // namespace cb

extern "C" int cb_upload_i32(const int32_t* host_vals, int n, int32_t* dst_dev, void* stream) {
    if (!host_vals || !dst_dev || n < 1 || n > CB_MAX_AGENTS + 1) return CB_ERR_ARG;
    cb::I32Pack p;
    p.n = n;
    for (int i = 0; i < n; ++i) p.v[i] = host_vals[i];
    cb::upload_i32_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(p, dst_dev);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_pfn_scatter(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows,
                              const int32_t* n_voxels_dev, int max_pts, const float* w, const float* scale,
                              const float* shift, const float* vsize, const float* center_off, int n_agents,
                              int canvas_agents, int ny, int nx, void* canvas_ps, int64_t lo_off,
                              int64_t* dirty_rows, int32_t* dirty_count, void* stream) {
    using namespace cb;
    if (max_pts < 1 || max_pts > 32 || n_agents < 1 || canvas_agents < n_agents || !canvas_ps) return CB_ERR_ARG;
    if (n_rows <= 0) return CB_OK;
    const CanvasGeom cg = make_canvas_geom(canvas_agents, ny, nx);
    const PfnParams pp = make_pfn(w, scale, shift, vsize, center_off);
    int blocks = (n_rows + 31) / 32;
    if (blocks > 148 * 4) blocks = 148 * 4;
    pfn_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)voxels, (const int4*)coords, num_points, n_rows, n_voxels_dev, max_pts, pp, cg, n_agents,
        (__nv_bfloat16*)canvas_ps, (long)lo_off, (long*)dirty_rows, dirty_count);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_canvas_clear(void* canvas_ps, int64_t lo_off, const int64_t* dirty_rows, const int32_t* dirty_count,
                               int capacity, void* stream) {
    if (!canvas_ps || !dirty_rows || !dirty_count || capacity < 1) return CB_ERR_ARG;
    cb::canvas_clear_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)canvas_ps, (long)lo_off,
                                                                       (const long*)dirty_rows, dirty_count, capacity);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
