// Pillar front-end on sm_100a: deterministic voxelisation (A1/A2), PillarVFE+PFN (A3/A4) and
// PointPillarScatter (A5), either staged through reference-format voxel tensors or fused
// points -> canvas.  HBM-bound integer/byte work: coalesced 16-byte point loads, one warp per pillar,
// 128-byte canvas-cell stores.  See include/coalign_b200.h for the reference lines each entry replaces.
//
// Determinism: spconv's generator is serial (voxel id = order of first appearance, a voxel keeps its
// first `max_pts` points, new voxels are refused after `max_voxels`).  We reproduce it bit-exactly with
//   K1  cell id per point, atomicMin(first point index per cell), per-cell count
//   K2  one CTA per agent: ordered scan over "leader" points (first[cell]==i) -> voxel ids, CSR offsets
//   K3  CSR fill (arbitrary order inside a cell)
//   K4  one warp per voxel: rank the cell's point indices, keep the `max_pts` smallest in order.
#include <limits.h>
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

struct AgentOffsets { int n_agents; int off[CB_MAX_AGENTS + 1]; };

struct VoxWs {            // workspace carve-up (device pointers)
    int* first;           // [n_agents][ncell]   min point index per cell (0x7f7f7f7f = empty)
    int* count;           // [n_agents][ncell]
    int* cursor;          // [n_agents][vcap]
    int* cell2vox;        // [n_agents][ncell]
    int* vox_off;         // [n_agents][vcap+1]
    int* vox_cell;        // [n_agents][vcap]
    int* cellid;          // [sum_P]
    int* list;            // [sum_P]
    int* nvox;            // [n_agents+1]
    int ncell, vcap;
};

struct Geom { float r0, r1, r2, v0, v1, v2; int gx, gy, gz; };

// PS canvas addressing (include/coalign_b200.h): 4 parity planes of PF-padded half-resolution maps
struct CanvasGeom {
    int ny, nx, Hq, Wq;   // Hq = ceil(ny/2)+2, Wq = ceil(nx/2)+2  (padded plane dims)
    long plane_rows;      // n_agents*Hq*Wq
};
__device__ __forceinline__ long canvas_row(const CanvasGeom& c, int a, int y, int x) {
    const int ph = (y & 1) * 2 + (x & 1);
    return (long)ph * c.plane_rows + ((long)a * c.Hq + (y >> 1) + 1) * c.Wq + (x >> 1) + 1;
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
__global__ void __launch_bounds__(1024) vox_scan_kernel(const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                        int max_voxels) {
    const int a = blockIdx.x;
    const int p0 = ao.off[a], np = ao.off[a + 1] - p0;
    const int* first = ws.first + (long)a * ws.ncell;
    const int* count = ws.count + (long)a * ws.ncell;
    int* cell2vox = ws.cell2vox + (long)a * ws.ncell;
    int* vox_off = ws.vox_off + (long)a * (ws.vcap + 1);
    int* vox_cell = ws.vox_cell + (long)a * ws.vcap;
    __shared__ int s_warp_v[32], s_warp_c[32];
    __shared__ int s_run_v, s_run_c;
    if (threadIdx.x == 0) { s_run_v = 0; s_run_c = 0; vox_off[0] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < np; base += 1024) {
        const int i = base + threadIdx.x;
        int cell = -1, lead = 0, cnt = 0;
        if (i < np) {
            cell = ws.cellid[p0 + i];
            if (cell >= 0 && first[cell] == i) { lead = 1; cnt = count[cell]; }
        }
        int v = lead, c = cnt;                                 // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int tv = __shfl_up_sync(0xffffffffu, v, d), tc = __shfl_up_sync(0xffffffffu, c, d);
            if (lane >= d) { v += tv; c += tc; }
        }
        if (lane == 31) { s_warp_v[warp] = v; s_warp_c[warp] = c; }
        __syncthreads();
        if (warp == 0) {
            int wv = s_warp_v[lane], wc = s_warp_c[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tv = __shfl_up_sync(0xffffffffu, wv, d), tc = __shfl_up_sync(0xffffffffu, wc, d);
                if (lane >= d) { wv += tv; wc += tc; }
            }
            s_warp_v[lane] = wv; s_warp_c[lane] = wc;          // inclusive over warps
        }
        __syncthreads();
        const int pre_v = s_run_v + (warp ? s_warp_v[warp - 1] : 0) + v - lead;   // exclusive prefix
        const int pre_c = s_run_c + (warp ? s_warp_c[warp - 1] : 0) + c - cnt;
        if (lead) {
            if (pre_v < max_voxels) {
                cell2vox[cell] = pre_v;
                vox_cell[pre_v] = cell;
                vox_off[pre_v + 1] = pre_c + cnt;              // CSR end of this voxel == begin of the next
            } else {
                cell2vox[cell] = -1;                           // refused: max_voxels reached
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_run_v += s_warp_v[31]; s_run_c += s_warp_c[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) ws.nvox[a] = s_run_v < max_voxels ? s_run_v : max_voxels;
}

__global__ void vox_total_kernel(const VoxWs ws, int n_agents, int* n_voxels_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int t = 0;
        for (int a = 0; a < n_agents; ++a) {
            if (n_voxels_out) n_voxels_out[a] = ws.nvox[a];
            t += ws.nvox[a];
        }
        ws.nvox[n_agents] = t;
        if (n_voxels_out) n_voxels_out[n_agents] = t;
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

// K4 helpers ----------------------------------------------------------------------------------
// The `max_pts` smallest point indices of voxel (a,v), ascending, into s_sorted[]; returns the total count.
__device__ __forceinline__ int voxel_sorted_points(const VoxWs& ws, const AgentOffsets& ao, int a, int v, int max_pts,
                                                   int* s_sorted, int lane) {
    const int* off = ws.vox_off + (long)a * (ws.vcap + 1);
    const int beg = off[v], cnt = off[v + 1] - beg;
    const int* lst = ws.list + ao.off[a] + beg;
    for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        const int mine = e < cnt ? lst[e] : INT_MAX;
        int rank = 0;
        for (int m0 = 0; m0 < cnt; m0 += 32) {
            const int other = (m0 + lane) < cnt ? lst[m0 + lane] : INT_MAX;
#pragma unroll
            for (int t = 0; t < 32; ++t) rank += (__shfl_sync(0xffffffffu, other, t) < mine) ? 1 : 0;
        }
        if (e < cnt && rank < max_pts) s_sorted[rank] = mine;
    }
    __syncwarp();
    return cnt;
}

struct PfnRegs {                 // per-lane slice of the PFN parameters: channels 2*lane, 2*lane+1
    float w[2][10], sc[2], sh[2];
};
__device__ __forceinline__ void load_pfn(PfnRegs& r, const float* w, const float* scale, const float* shift, int lane) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int j = 0; j < 10; ++j) r.w[c][j] = __ldg(w + (2 * lane + c) * 10 + j);
        r.sc[c] = __ldg(scale + 2 * lane + c);
        r.sh[c] = __ldg(shift + 2 * lane + c);
    }
}

// PFN of one pillar held by one warp: lane k carries point slot k (valid iff k < n, n >= 1).
// `off*` = voxel/2 + range_min.  Writes channels (2*lane, 2*lane+1) of canvas cell (a, cy, cx).
__device__ __forceinline__ void pfn_pillar_store(float4 p, int n, int max_pts, int a, int cz, int cy, int cx,
                                                 float vx, float vy, float vz, float offx, float offy, float offz,
                                                 const PfnRegs& r, const CanvasGeom& cg, __nv_bfloat16* canvas,
                                                 long lo_off, int lane) {
    if (lane >= n) p = make_float4(0.f, 0.f, 0.f, 0.f);
    float sx = p.x, sy = p.y, sz = p.z;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, d);
        sy += __shfl_xor_sync(0xffffffffu, sy, d);
        sz += __shfl_xor_sync(0xffffffffu, sz, d);
    }
    const float fn = (float)n;
    const float mx = __fdiv_rn(sx, fn), my = __fdiv_rn(sy, fn), mz = __fdiv_rn(sz, fn);
    // pillar centre = coord*voxel + (voxel/2 + range_min)  (pillar_vfe.py:87-89,124-132), fp32, unfused
    const float ctrx = __fadd_rn(__fmul_rn((float)cx, vx), offx);
    const float ctry = __fadd_rn(__fmul_rn((float)cy, vy), offy);
    const float ctrz = __fadd_rn(__fmul_rn((float)cz, vz), offz);
    float f[10];
    f[0] = p.x; f[1] = p.y; f[2] = p.z; f[3] = p.w;
    f[4] = p.x - mx; f[5] = p.y - my; f[6] = p.z - mz;
    f[7] = p.x - ctrx; f[8] = p.y - ctry; f[9] = p.z - ctrz;
    float best0 = 0.f, best1 = 0.f;                  // ReLU outputs are >= 0
    if (n < max_pts) {                               // zero-padded slots take part in the max (pillar_vfe.py:45-46)
        best0 = fmaxf(r.sh[0], 0.f);
        best1 = fmaxf(r.sh[1], 0.f);
    }
    for (int k = 0; k < n; ++k) {
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const float fj = __shfl_sync(0xffffffffu, f[j], k);
            y0 = fmaf(r.w[0][j], fj, y0);
            y1 = fmaf(r.w[1][j], fj, y1);
        }
        best0 = fmaxf(best0, fmaf(y0, r.sc[0], r.sh[0]));
        best1 = fmaxf(best1, fmaf(y1, r.sc[1], r.sh[1]));
    }
    const long row = canvas_row(cg, a, cy, cx);
    uint32_t* dst = reinterpret_cast<uint32_t*>(canvas + row * 64) + lane;
    const uint32_t hi = pack_bf16(best0, best1);
    *dst = hi;
    if (lo_off != 0) {
        uint32_t* dl = reinterpret_cast<uint32_t*>(canvas + lo_off + row * 64) + lane;
        *dl = pack_bf16(best0 - bf16_lo(hi), best1 - bf16_hi(hi));
    }
}

// K4a: emit reference-format voxel tensors --------------------------------------------------------
__global__ void __launch_bounds__(256) vox_emit_kernel(const float4* __restrict__ pts,
                                                       const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                       const Geom g, int max_pts, float4* __restrict__ voxels,
                                                       int4* __restrict__ coords, int* __restrict__ num_points) {
    __shared__ int s_sorted[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + wib, nw = gridDim.x * 8;
    int base = 0;
    for (int a = 0; a < ao.n_agents; ++a) {
        const int nv = ws.nvox[a];
        for (int v = gw; v < nv; v += nw) {
            const int cnt = voxel_sorted_points(ws, ao, a, v, max_pts, s_sorted[wib], lane);
            const int n = cnt < max_pts ? cnt : max_pts;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < n) p = __ldg(pts + ao.off[a] + s_sorted[wib][lane]);
            const long row = base + v;
            if (lane < max_pts) voxels[row * max_pts + lane] = p;
            if (lane == 0) {
                const int cell = ws.vox_cell[(long)a * ws.vcap + v];
                const int x = cell % g.gx, y = (cell / g.gx) % g.gy, z = cell / (g.gx * g.gy);
                coords[row] = make_int4(a, z, y, x);
                num_points[row] = n;
            }
            __syncwarp();
        }
        base += nv;
    }
}

// K4b: fused PFN + scatter straight from the CSR lists -------------------------------------------
__global__ void __launch_bounds__(256) vox_pfn_kernel(const float4* __restrict__ pts,
                                                      const __grid_constant__ AgentOffsets ao, const VoxWs ws,
                                                      const Geom g, int max_pts, const float* __restrict__ w,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      float offx, float offy, float offz,
                                                      const CanvasGeom cg, __nv_bfloat16* canvas, long lo_off) {
    __shared__ int s_sorted[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + wib, nw = gridDim.x * 8;
    PfnRegs r;
    load_pfn(r, w, scale, shift, lane);
    for (int a = 0; a < ao.n_agents; ++a) {
        const int nv = ws.nvox[a];
        for (int v = gw; v < nv; v += nw) {
            const int cnt = voxel_sorted_points(ws, ao, a, v, max_pts, s_sorted[wib], lane);
            const int n = cnt < max_pts ? cnt : max_pts;
            float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < n) p = __ldg(pts + ao.off[a] + s_sorted[wib][lane]);
            const int cell = ws.vox_cell[(long)a * ws.vcap + v];
            const int x = cell % g.gx, y = (cell / g.gx) % g.gy, z = cell / (g.gx * g.gy);
            pfn_pillar_store(p, n, max_pts, a, z, y, x, g.v0, g.v1, g.v2, offx, offy, offz, r, cg, canvas, lo_off, lane);
            __syncwarp();
        }
    }
}

// PFN + scatter from reference-format voxel tensors ----------------------------------------------
__global__ void __launch_bounds__(256) pfn_scatter_kernel(const float4* __restrict__ voxels,
                                                          const int4* __restrict__ coords,
                                                          const int* __restrict__ num_points, int n_rows,
                                                          const int* __restrict__ n_rows_dev, int max_pts,
                                                          const float* __restrict__ w, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float vx, float vy, float vz,
                                                          float offx, float offy, float offz, const CanvasGeom cg,
                                                          int n_agents, __nv_bfloat16* canvas, long lo_off) {
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    PfnRegs r;
    load_pfn(r, w, scale, shift, lane);
    const int rows = n_rows_dev ? min(n_rows, *n_rows_dev) : n_rows;
    for (int v = gw; v < rows; v += nw) {
        const int4 c = __ldg(coords + v);                     // [agent, z, y, x]
        int n = __ldg(num_points + v);
        n = n < max_pts ? n : max_pts;
        if (n < 1 || c.x < 0 || c.x >= n_agents || c.z < 0 || c.z >= cg.ny || c.w < 0 || c.w >= cg.nx) continue;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < n) p = __ldg(voxels + (long)v * max_pts + lane);
        pfn_pillar_store(p, n, max_pts, c.x, c.y, c.z, c.w, vx, vy, vz, offx, offy, offz, r, cg, canvas, lo_off, lane);
    }
}

// ------------------------------------------------------------------------------------------ host
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int carve(VoxWs& ws, void* base, size_t bytes, int n_agents, int sum_points, const int32_t* grid, int max_voxels,
                 size_t* clear_bytes, size_t* total_bytes) {
    const long ncell = (long)grid[0] * grid[1] * grid[2];
    const int vcap = max_voxels < sum_points ? max_voxels : (sum_points > 0 ? sum_points : 1);
    uint8_t* p = (uint8_t*)base;
    size_t o = 0;
    auto take = [&](size_t n_int) { void* r = p ? p + o : nullptr; o += align256(n_int * 4); return (int*)r; };
    ws.first = take((size_t)n_agents * ncell);
    size_t o_first_end = o;
    ws.count = take((size_t)n_agents * ncell);
    ws.cursor = take((size_t)n_agents * vcap);
    size_t o_clear_end = o;
    ws.cell2vox = take((size_t)n_agents * ncell);
    ws.vox_off = take((size_t)n_agents * (vcap + 1));
    ws.vox_cell = take((size_t)n_agents * vcap);
    ws.cellid = take((size_t)(sum_points > 0 ? sum_points : 1));
    ws.list = take((size_t)(sum_points > 0 ? sum_points : 1));
    ws.nvox = take((size_t)n_agents + 1);
    ws.ncell = (int)ncell;
    ws.vcap = vcap;
    if (clear_bytes) { clear_bytes[0] = o_first_end; clear_bytes[1] = o_clear_end - o_first_end; }
    if (total_bytes) *total_bytes = o;
    if (p && o > bytes) return CB_ERR_ARG;
    return CB_OK;
}

static size_t ws_bytes(int n_agents, int sum_points, const int32_t* grid, int max_voxels) {
    VoxWs ws;
    size_t total = 0;
    carve(ws, nullptr, 0, n_agents, sum_points, grid, max_voxels, nullptr, &total);
    return total;
}

// Shared front half (K1..K3).  pt_offset is a HOST array.
static int run_front(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                     const float* vsize, const int32_t* grid, int max_pts, int max_voxels, void* workspace,
                     size_t workspace_bytes, cudaStream_t st, AgentOffsets& ao, VoxWs& ws, Geom& g) {
    if (n_agents < 1 || n_agents > CB_MAX_AGENTS || max_pts < 1 || max_pts > 32 || max_voxels < 1) return CB_ERR_ARG;
    if (!workspace || ((uintptr_t)points & 15)) return CB_ERR_ARG;
    ao.n_agents = n_agents;
    for (int i = 0; i <= n_agents; ++i) ao.off[i] = pt_offset[i];
    if (ao.off[0] != 0) return CB_ERR_ARG;
    for (int i = 0; i < n_agents; ++i) if (ao.off[i + 1] < ao.off[i]) return CB_ERR_ARG;
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
    vox_scan_kernel<<<n_agents, 1024, 0, st>>>(ao, ws, max_voxels);
    CB_CHECK_LAUNCH();
    if (total > 0) {
        vox_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(ao, ws);
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

static CanvasGeom make_canvas_geom(int n_agents, int ny, int nx) {
    CanvasGeom cg;
    cg.ny = ny; cg.nx = nx;
    cg.Hq = (ny + 1) / 2 + 2;
    cg.Wq = (nx + 1) / 2 + 2;
    cg.plane_rows = (long)n_agents * cg.Hq * cg.Wq;
    return cg;
}

}  // namespace cb

extern "C" size_t cb_voxelize_workspace_bytes(int n_agents, int sum_points, const int32_t* grid, int max_voxels) {
    return cb::ws_bytes(n_agents, sum_points, grid, max_voxels);
}

extern "C" int cb_voxelize(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                           const float* vsize, const int32_t* grid, int max_pts, int max_voxels, float* voxels,
                           int32_t* coords, int32_t* num_points, int32_t* n_voxels, void* workspace,
                           size_t workspace_bytes, void* stream) {
    using namespace cb;
    cudaStream_t st = (cudaStream_t)stream;
    AgentOffsets ao; VoxWs ws; Geom g;
    int rc = run_front(points, pt_offset, n_agents, range, vsize, grid, max_pts, max_voxels, workspace,
                       workspace_bytes, st, ao, ws, g);
    if (rc) return rc;
    vox_total_kernel<<<1, 32, 0, st>>>(ws, n_agents, n_voxels);
    CB_CHECK_LAUNCH();
    vox_emit_kernel<<<148 * 4, 256, 0, st>>>((const float4*)points, ao, ws, g, max_pts, (float4*)voxels,
                                             (int4*)coords, num_points);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_points_to_canvas(const float* points, const int32_t* pt_offset, int n_agents, const float* range,
                                   const float* vsize, const int32_t* grid, int max_pts, int max_voxels,
                                   const float* w, const float* scale, const float* shift,
                                   const float* center_off, int canvas_agents, void* canvas_ps,
                                   int64_t lo_off, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace cb;
    if (grid[2] != 1 || !canvas_ps || canvas_agents < n_agents) return CB_ERR_ARG;       // PointPillarScatter asserts nz == 1 (point_pillar_scatter.py:13)
    cudaStream_t st = (cudaStream_t)stream;
    AgentOffsets ao; VoxWs ws; Geom g;
    int rc = run_front(points, pt_offset, n_agents, range, vsize, grid, max_pts, max_voxels, workspace,
                       workspace_bytes, st, ao, ws, g);
    if (rc) return rc;
    const CanvasGeom cg = make_canvas_geom(canvas_agents, grid[1], grid[0]);
    vox_pfn_kernel<<<148 * 4, 256, 0, st>>>((const float4*)points, ao, ws, g, max_pts, w, scale, shift,
                                            center_off[0], center_off[1], center_off[2], cg,
                                            (__nv_bfloat16*)canvas_ps, (long)lo_off);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_pfn_scatter(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows,
                              const int32_t* n_voxels_dev, int max_pts, const float* w, const float* scale,
                              const float* shift, const float* vsize, const float* center_off, int n_agents,
                              int canvas_agents, int ny, int nx, void* canvas_ps, int64_t lo_off, void* stream) {
    using namespace cb;
    if (max_pts < 1 || max_pts > 32 || n_agents < 1 || canvas_agents < n_agents || !canvas_ps) return CB_ERR_ARG;
    if (n_rows <= 0) return CB_OK;
    const CanvasGeom cg = make_canvas_geom(canvas_agents, ny, nx);
    const float offx = center_off[0], offy = center_off[1], offz = center_off[2];
    int blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pfn_scatter_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)voxels, (const int4*)coords, num_points, n_rows, n_voxels_dev, max_pts, w, scale, shift,
        vsize[0], vsize[1], vsize[2], offx, offy, offz, cg, n_agents, (__nv_bfloat16*)canvas_ps, (long)lo_off);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
