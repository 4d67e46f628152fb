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
//   K4   8 lanes per voxel: emit reference-format tensors, or PFN + canvas store.
#include <limits.h>
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

constexpr int CHUNK = 1024;       // points per scan chunk (256 threads x 4)
constexpr int NCLS = 5;           // pillar classes by point count: 1 | 2 | 3-4 | 5-8 | 9+

struct AgentOffsets { int n_agents; int off[CB_MAX_AGENTS + 1]; };

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
    int* perm;            // [n_agents][NCLS][vcap]  voxel ids bucketed by point-count class
    int* bucket;          // [n_agents][NCLS]        bucket fill counters
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
    // processing order of the fused PFN kernel: pillars bucketed by point-count class.  Positions are claimed with
    // shared-memory atomics inside the CTA and ONE global atomic per class and CTA.
    __shared__ int s_cnt[NCLS], s_gbase[NCLS];
    if (threadIdx.x < NCLS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int lpos[4], lcls[4], lpv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        lpos[j] = -1; lcls[j] = 0; lpv[j] = pv;
        if (lead[j]) {
            if (pv < max_voxels) {
                cell2vox[cell[j]] = pv;
                vox_off[pv] = pc;                           // CSR begin (used by the fill kernel)
                const int cx = cell[j] % g.gx, cyz = cell[j] / g.gx;
                const int cy = cyz % g.gy, cz = cyz / g.gy;
                vox_meta[pv] = make_int4(pc, cnt[j], cx | (cy << 12) | (cz << 24), chunk * CHUNK + (int)threadIdx.x * 4 + j);
                lcls[j] = cnt[j] <= 1 ? 0 : (cnt[j] == 2 ? 1 : (cnt[j] <= 4 ? 2 : (cnt[j] <= 8 ? 3 : 4)));
                lpos[j] = atomicAdd(&s_cnt[lcls[j]], 1);
            } else {
                cell2vox[cell[j]] = -1;                    // refused: max_voxels reached
            }
            pv += 1; pc += cnt[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < NCLS && s_cnt[threadIdx.x] > 0)
        s_gbase[threadIdx.x] = atomicAdd(ws.bucket + a * NCLS + threadIdx.x, s_cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (lpos[j] >= 0) ws.perm[((long)a * NCLS + lcls[j]) * ws.vcap + s_gbase[lcls[j]] + lpos[j]] = lpv[j];
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
// 8 lanes ("group") cooperate on one voxel.  sub = lane & 7.
struct PfnParams {
    const float* w; const float* scale; const float* shift;     // [64][10], [64], [64]
    float vx, vy, vz, offx, offy, offz;                         // voxel size, voxel/2 + range_min
};

struct PfnRegs { float w[8][10], sc[8], sh[8]; };               // channels 8*sub .. 8*sub+7
__device__ __forceinline__ void load_pfn(PfnRegs& r, const PfnParams& pp, int sub) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
        for (int j = 0; j < 10; ++j) r.w[c][j] = __ldg(pp.w + (8 * sub + c) * 10 + j);
        r.sc[c] = __ldg(pp.scale + 8 * sub + c);
        r.sh[c] = __ldg(pp.shift + 8 * sub + c);
    }
}

// Per-pillar geometry terms shared by both PFN mappings (pillar_vfe.py:118-132): mean over the valid slots
// (slot order) and the pillar centre = coord*voxel + (voxel/2 + range_min), fp32, unfused like the reference.
struct PillarTerms { float mx, my, mz, ctrx, ctry, ctrz; };
template <class PointFn>
__device__ __forceinline__ PillarTerms pillar_terms(PointFn pt, const float4& p0, int n, int cz, int cy, int cx,
                                                    const PfnParams& pp) {
    float sx = p0.x, sy = p0.y, sz = p0.z;
    for (int k = 1; k < n; ++k) { const float4 p = pt(k); sx += p.x; sy += p.y; sz += p.z; }
    const float fn = (float)n;
    PillarTerms t;
    t.mx = __fdiv_rn(sx, fn); t.my = __fdiv_rn(sy, fn); t.mz = __fdiv_rn(sz, fn);
    t.ctrx = __fadd_rn(__fmul_rn((float)cx, pp.vx), pp.offx);
    t.ctry = __fadd_rn(__fmul_rn((float)cy, pp.vy), pp.offy);
    t.ctrz = __fadd_rn(__fmul_rn((float)cz, pp.vz), pp.offz);
    return t;
}
__device__ __forceinline__ void point_features(const float4& p, const PillarTerms& t, float (&f)[10]) {
    f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = p.w;
    f[4] = p.x - t.mx; f[5] = p.y - t.my; f[6] = p.z - t.mz;
    f[7] = p.x - t.ctrx; f[8] = p.y - t.ctry; f[9] = p.z - t.ctrz;
}

// PFN of one pillar by one 8-lane group (staged path): PointFn(k) returns slot k (k < n, n >= 1).  Every lane walks
// all n points (broadcast loads) and produces 8 of the 64 channels; one 16-byte store per lane.
template <class PointFn>
__device__ __forceinline__ void pfn_group_store(PointFn pt, int n, int max_pts, int a, int cz, int cy, int cx,
                                                const PfnParams& pp, const PfnRegs& r, const CanvasGeom& cg,
                                                __nv_bfloat16* canvas, long lo_off, int sub, long* dirty_slot) {
    const float4 p0 = pt(0);
    const PillarTerms t = pillar_terms(pt, p0, n, cz, cy, cx, pp);
    float best[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) best[c] = (n < max_pts) ? fmaxf(r.sh[c], 0.f) : 0.f;   // zero-padded slots join the max
    for (int k = 0; k < n; ++k) {
        const float4 p = k == 0 ? p0 : pt(k);
        float f[10];
        point_features(p, t, f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float y = 0.f;
#pragma unroll
            for (int j = 0; j < 10; ++j) y = fmaf(r.w[c][j], f[j], y);
            best[c] = fmaxf(best[c], fmaf(y, r.sc[c], r.sh[c]));       // ReLU folded into the max (best >= 0)
        }
    }
    const long row = canvas_row(cg, a, cy, cx);
    uint4 hi;
    hi.x = pack_bf16(best[0], best[1]); hi.y = pack_bf16(best[2], best[3]);
    hi.z = pack_bf16(best[4], best[5]); hi.w = pack_bf16(best[6], best[7]);
    reinterpret_cast<uint4*>(canvas + row * 64)[sub] = hi;
    if (lo_off != 0) {
        uint4 lo;
        lo.x = pack_bf16(best[0] - bf16_lo(hi.x), best[1] - bf16_hi(hi.x));
        lo.y = pack_bf16(best[2] - bf16_lo(hi.y), best[3] - bf16_hi(hi.y));
        lo.z = pack_bf16(best[4] - bf16_lo(hi.z), best[5] - bf16_hi(hi.z));
        lo.w = pack_bf16(best[6] - bf16_lo(hi.w), best[7] - bf16_hi(hi.w));
        reinterpret_cast<uint4*>(canvas + lo_off + row * 64)[sub] = lo;
    }
    if (dirty_slot && sub == 0) *dirty_slot = row;
}

// Fused path: ONE THREAD per pillar and channel half; the PFN parameters live in constant memory, so every FFMA
// takes its weight as a constant-bank operand (no weight registers, no shared-memory traffic).  Pillars are visited
// class by class (same point count inside a warp, see vox_chunk_assign_kernel) so warps do not diverge on n.
__constant__ float c_pfn_w[64 * 10];
__constant__ float c_pfn_sc[64];
__constant__ float c_pfn_sh[64];

template <int HALF, class PointFn>
__device__ __forceinline__ void pfn_thread_store(PointFn pt, const float4 p0, int n, int max_pts, int a, int cz, int cy,
                                                 int cx, const PfnParams& pp, const CanvasGeom& cg, __nv_bfloat16* canvas,
                                                 long lo_off, long* dirty_slot) {
    const PillarTerms t = pillar_terms(pt, p0, n, cz, cy, cx, pp);
    float best[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) best[c] = (n < max_pts) ? fmaxf(c_pfn_sh[HALF * 32 + c], 0.f) : 0.f;
    for (int k = 0; k < n; ++k) {
        const float4 p = k == 0 ? p0 : pt(k);
        float f[10];
        point_features(p, t, f);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            float y = 0.f;
#pragma unroll
            for (int j = 0; j < 10; ++j) y = fmaf(c_pfn_w[(HALF * 32 + c) * 10 + j], f[j], y);
            best[c] = fmaxf(best[c], fmaf(y, c_pfn_sc[HALF * 32 + c], c_pfn_sh[HALF * 32 + c]));
        }
    }
    const long row = canvas_row(cg, a, cy, cx);
    uint4* dst = reinterpret_cast<uint4*>(canvas + row * 64 + HALF * 32);
    uint32_t hi[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) hi[c] = pack_bf16(best[2 * c], best[2 * c + 1]);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
    if (lo_off != 0) {
        uint4* dl = reinterpret_cast<uint4*>(canvas + lo_off + row * 64 + HALF * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * q + e;
                lo[e] = pack_bf16(best[2 * c] - bf16_lo(hi[c]), best[2 * c + 1] - bf16_hi(hi[c]));
            }
            dl[q] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (dirty_slot && HALF == 0) *dirty_slot = row;
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

// K4b: fused PFN + scatter from the regrouped points: thread per (pillar, channel half) -------------------
// Work items are the pillars of all agents flattened class-major (class, agent, index): every thread of the grid has
// work and the pillars of a warp share a point-count class.
__global__ void __launch_bounds__(256, 3) vox_pfn_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                         int max_pts, const PfnParams pp, const CanvasGeom cg,
                                                         __nv_bfloat16* canvas, long lo_off, long* dirty_rows) {
    __shared__ int s_base[CB_MAX_AGENTS + 1];                      // first output voxel row of each agent
    __shared__ int s_seg[NCLS * CB_MAX_AGENTS + 1];                // prefix over (class, agent) segments
    const int nseg = NCLS * ao.n_agents;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int a = 0; a < ao.n_agents; ++a) { s_base[a] = b; b += ws.nvox[a]; }
        int t = 0;
        for (int sgm = 0; sgm < nseg; ++sgm) {
            s_seg[sgm] = t;
            t += ws.bucket[(sgm % ao.n_agents) * NCLS + sgm / ao.n_agents];
        }
        s_seg[nseg] = t;
    }
    __syncthreads();
    const int total = s_seg[nseg];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = warp & 1;                                     // even warps: channels 0-31, odd warps: 32-63
    const int slot = (blockIdx.x * 4 + (warp >> 1)) * 32 + lane;
    const int nslot = gridDim.x * 4 * 32;
    // one pillar per thread and iteration; the next pillar's (perm -> metadata -> first point) chain is fetched while
    // the current one is computed
    struct Item { int a, v; int4 m; float4 p0; };
    auto fetch = [&](int gidx) {
        Item it;
        int lo = 0, hi = nseg;                                     // largest seg with s_seg[seg] <= gidx
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_seg[mid] <= gidx) lo = mid; else hi = mid; }
        const int c = lo / ao.n_agents;
        it.a = lo - c * ao.n_agents;
        it.v = ws.perm[((long)it.a * NCLS + c) * ws.vcap + (gidx - s_seg[lo])];
        it.m = __ldg(ws.vox_meta + (long)it.a * ws.vcap + it.v);
        it.p0 = ws.vp[ao.off[it.a] + it.m.x];
        return it;
    };
    Item cur;
    if (slot < total) cur = fetch(slot);
    for (int gidx = slot; gidx < total; gidx += nslot) {
        Item nxt = cur;
        if (gidx + nslot < total) nxt = fetch(gidx + nslot);
        const int a = cur.a;
        const int4 m = cur.m;
        const int n = m.y < max_pts ? m.y : max_pts;
        const float4* vpp = ws.vp + ao.off[a] + m.x;
        const int cz = (m.z >> 24) & 0xFF, cy = (m.z >> 12) & 0xFFF, cx = m.z & 0xFFF;
        long* dslot = dirty_rows ? dirty_rows + s_base[a] + cur.v : nullptr;
        if (half == 0) pfn_thread_store<0>([&](int k) { return vpp[k]; }, cur.p0, n, max_pts, a, cz, cy, cx, pp, cg, canvas, lo_off, dslot);
        else           pfn_thread_store<1>([&](int k) { return vpp[k]; }, cur.p0, n, max_pts, a, cz, cy, cx, pp, cg, canvas, lo_off, dslot);
        cur = nxt;
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
    PfnRegs r;
    load_pfn(r, pp, sub);
    const int rows = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
    if (dirty_count && blockIdx.x == 0 && threadIdx.x == 0) *dirty_count = rows;
    for (int v = gg; v < rows; v += ng) {
        const int4 c = __ldg(coords + v);                     // [agent, z, y, x]
        int n = __ldg(num_points + v);
        n = n < max_pts ? n : max_pts;
        const bool ok = !(n < 1 || c.x < 0 || c.x >= n_agents || c.z < 0 || c.z >= cg.ny || c.w < 0 || c.w >= cg.nx);
        if (!ok) { if (dirty_rows && sub == 0) dirty_rows[v] = -1; continue; }
        const float4* vp = voxels + (long)v * max_pts;
        pfn_group_store([&](int k) { return __ldg(vp + k); }, n, max_pts, c.x, c.y, c.z, c.w, pp, r, cg, canvas, lo_off,
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
    for (int i = gg; i < n; i += ng) {
        const long row = dirty_rows[i];
        if (row < 0) continue;
        reinterpret_cast<uint4*>(canvas + row * 64)[sub] = z;
        if (lo_off != 0) reinterpret_cast<uint4*>(canvas + lo_off + row * 64)[sub] = z;
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
    ws.bucket = (int*)take((size_t)n_agents * NCLS * 4);
    size_t o_clear_end = o;
    ws.cell2vox = (int*)take((size_t)n_agents * ncell * 4);
    ws.vox_off = (int*)take((size_t)n_agents * (vcap + 1) * 4);
    ws.vox_meta = (int4*)take((size_t)n_agents * vcap * 16);
    ws.cellid = (int*)take((size_t)sp * 4);
    ws.list = (int*)take((size_t)sp * 4);
    ws.nvox = (int*)take(((size_t)n_agents + 1) * 4);
    ws.chunk_tot = (int2*)take((size_t)n_agents * max_chunks * 8);
    ws.vp = (float4*)take((size_t)sp * 16);
    ws.perm = (int*)take((size_t)n_agents * NCLS * vcap * 4);
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
    size_t total = 0;
    cb::carve(ws, nullptr, 0, n_agents, sum_points, grid, max_voxels, nullptr, &total);
    return total;
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

extern "C" int cb_points_to_canvas(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                                   const float* vsize, const int32_t* grid, int max_pts, int max_voxels,
                                   const float* w, const float* scale, const float* shift,
                                   const float* center_off, int canvas_agents, void* canvas_ps,
                                   int64_t lo_off, int64_t* dirty_rows, int32_t* dirty_count,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    using namespace cb;
    if (grid[2] != 1 || !canvas_ps || canvas_agents < n_agents) return CB_ERR_ARG;   // nz == 1 (point_pillar_scatter.py:13)
    cudaStream_t st = (cudaStream_t)stream;
    AgentOffsets ao; VoxWs ws; Geom g;
    int rc = run_front(points, pt_offset, n_agents, range, vsize, grid, max_pts, max_voxels, workspace,
                       workspace_bytes, st, ao, ws, g, nullptr, dirty_count);
    if (rc) return rc;
    const CanvasGeom cg = make_canvas_geom(canvas_agents, grid[1], grid[0]);
    if (cg.plane_rows * 4 >= (1L << 31)) return CB_ERR_ARG;
    const PfnParams pp = make_pfn(w, scale, shift, vsize, center_off);
    cudaError_t ce;
    ce = cudaMemcpyToSymbolAsync(c_pfn_w, w, sizeof(float) * 640, 0, cudaMemcpyDeviceToDevice, st);     if (ce) return (int)ce;
    ce = cudaMemcpyToSymbolAsync(c_pfn_sc, scale, sizeof(float) * 64, 0, cudaMemcpyDeviceToDevice, st);  if (ce) return (int)ce;
    ce = cudaMemcpyToSymbolAsync(c_pfn_sh, shift, sizeof(float) * 64, 0, cudaMemcpyDeviceToDevice, st);  if (ce) return (int)ce;
    vox_pfn_kernel<<<148 * 3, 256, 0, st>>>(ao, ws, max_pts, pp, cg, (__nv_bfloat16*)canvas_ps, (long)lo_off,
                                            (long*)dirty_rows);
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
    cb::canvas_clear_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)canvas_ps, (long)lo_off,
                                                                       (const long*)dirty_rows, dirty_count, capacity);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
