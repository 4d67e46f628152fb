// Training-step kernels around the tensor-core GEMMs (sm_100a): train-mode BatchNorm forward / backward, heads gradient
// packing, adjoint of the warp + attention fusion, train-mode PFN, parameter layout permutations, Adam.
// All HBM-bound element-wise / reduction work: 16-byte vector accesses, fp32 math, fp64 cross-block accumulation.
// Reference lines: include/coalign_b200.h ("Training step, device side").
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

// Programmatic dependent launch: these kernels run several waves of CTAs, so they do NOT trigger their dependents early
// (CTAs of the next kernel would occupy SM slots spinning in griddepcontrol.wait while later waves of this one still need
// them); griddepcontrol.wait is kept so that they may follow a kernel that does trigger early (the GEMM kernels).

// ------------------------------------------------------------------------------------------------ helpers
struct MapP {
    int n_img, Hp, Wp, c_total, c_mod, y_mode, y_pitch, y_ch_off, up_k, y_Hp, y_Wp, z_at_y, z_pitch;
    long y_plane_rows, rows_total;
};

static int fill_map(const cb_map* m, MapP& p) {
    if (!m || m->n_img < 1 || m->Hp < 3 || m->Wp < 3 || m->c_total < 8 || m->c_total % 8 || m->c_mod < 8 || m->c_mod % 8)
        return CB_ERR_ARG;
    if (m->c_total % m->c_mod || m->c_total > 2048 || m->c_mod > 256) return CB_ERR_ARG;
    if (m->y_mode != CB_OUT_PF && m->y_mode != CB_OUT_PS && m->y_mode != CB_OUT_UPSAMPLE) return CB_ERR_ARG;
    if (m->y_pitch % 8 || m->y_ch_off % 8) return CB_ERR_ARG;
    if (m->y_mode == CB_OUT_UPSAMPLE && (m->up_k < 1 || m->c_total != m->up_k * m->up_k * m->c_mod)) return CB_ERR_ARG;
    if (m->y_mode != CB_OUT_UPSAMPLE && m->c_total != m->c_mod) return CB_ERR_ARG;
    p.n_img = m->n_img; p.Hp = m->Hp; p.Wp = m->Wp; p.c_total = m->c_total; p.c_mod = m->c_mod; p.y_mode = m->y_mode;
    p.y_pitch = m->y_pitch; p.y_ch_off = m->y_ch_off; p.up_k = m->up_k; p.y_Hp = m->y_Hp; p.y_Wp = m->y_Wp;
    p.y_plane_rows = m->y_plane_rows;
    p.z_at_y = m->z_at_y; p.z_pitch = m->z_pitch;
    if (p.z_at_y && (p.z_pitch < p.c_mod || p.z_pitch % 8)) return CB_ERR_ARG;
    p.rows_total = (long)m->n_img * m->Hp * m->Wp;
    if (p.rows_total >= (1L << 31)) return CB_ERR_ARG;
    return CB_OK;
}

// z row q, column col (multiple of 8; 8 consecutive columns never straddle c_mod) -> element offset of the matching 8
// channels on the y side, or -1 when q is a halo row
__device__ __forceinline__ long y_offset(const MapP& m, unsigned q, int col) {
    const unsigned plane = (unsigned)(m.Hp * m.Wp);
    const unsigned n = q / plane, rem = q - n * plane;
    const unsigned hp = rem / (unsigned)m.Wp, wp = rem - hp * (unsigned)m.Wp;
    if (hp < 1u || hp > (unsigned)(m.Hp - 2) || wp < 1u || wp > (unsigned)(m.Wp - 2)) return -1;
    const int h = (int)hp - 1, w = (int)wp - 1;
    if (m.y_mode == CB_OUT_PF) return (long)q * m.y_pitch + m.y_ch_off + col;
    if (m.y_mode == CB_OUT_PS) {
        const int ph = (h & 1) * 2 + (w & 1);
        const long row = (long)ph * m.y_plane_rows + (long)((int)n * m.y_Hp + (h >> 1) + 1) * m.y_Wp + (w >> 1) + 1;
        return row * m.y_pitch + m.y_ch_off + col;
    }
    const int ab = col / m.c_mod, c = col - ab * m.c_mod;
    const int a = ab / m.up_k, b = ab - a * m.up_k;
    const long row = (long)((int)n * m.y_Hp + m.up_k * h + a + 1) * m.y_Wp + (m.up_k * w + b + 1);
    return row * m.y_pitch + m.y_ch_off + c;
}

// 8 bf16 (hi [+ lo plane]) -> 8 floats
__device__ __forceinline__ void load8(const __nv_bfloat16* base, long off, long lo_off, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + off));
    v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
    v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
    if (lo_off != 0) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(base + lo_off + off));
        v[0] += bf16_lo(l.x); v[1] += bf16_hi(l.x); v[2] += bf16_lo(l.y); v[3] += bf16_hi(l.y);
        v[4] += bf16_lo(l.z); v[5] += bf16_hi(l.z); v[6] += bf16_lo(l.w); v[7] += bf16_hi(l.w);
    }
}
__device__ __forceinline__ void store8(__nv_bfloat16* base, long off, long lo_off, const float (&v)[8]) {
    uint32_t hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) hi[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (lo_off != 0) {
        uint32_t lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) lo[e] = pack_bf16(v[2 * e] - bf16_lo(hi[e]), v[2 * e + 1] - bf16_hi(hi[e]));
        *reinterpret_cast<uint4*>(base + lo_off + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

constexpr int EW_THREADS = 256;
constexpr int EW_BLOCKS = 148 * 2;   // one resident wave of the 2-CTA/SM streaming kernels (several waves re-run their prologues and ramps)

// Per-block reduction of per-thread partial sums (8 channels each, two quantities) into fp64 global sums.
// s_part: [2][2048] floats of shared memory; slot of element j of this thread = tslot*8 + j with tslot = threadIdx.x
// (256 threads x 8 channels).  Thread c < c_mod then adds up every slot whose channel is c (fixed order: deterministic per
// block) and issues ONE fp64 atomic per quantity.  (Shared-memory float atomics compile to CAS spin loops on sm_100.)
__device__ __forceinline__ void block_reduce_to_global(float (&a0)[8], float (&a1)[8], int c_total, int c_mod, bool two,
                                                       float* s_part, double* sums) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s_part[threadIdx.x * 8 + j] = a0[j];
        if (two) s_part[2048 + threadIdx.x * 8 + j] = a1[j];
    }
    __syncthreads();
    // thread t owns chunk (t % cpr) -> columns [8*(t%cpr), +8); channel = column % c_mod
    const int cpr = c_total >> 3;
    const int rpi = EW_THREADS / cpr;                                 // row lanes per block (threads >= rpi*cpr hold zeros)
    for (int c = threadIdx.x; c < c_mod; c += EW_THREADS) {
        float t0 = 0.f, t1 = 0.f;
        for (int col = c; col < c_total; col += c_mod) {
            const int chunk = col >> 3, j = col & 7;
            for (int rl = 0; rl < rpi; ++rl) {
                const int t = rl * cpr + chunk;
                t0 += s_part[t * 8 + j];
                if (two) t1 += s_part[2048 + t * 8 + j];
            }
        }
        atomicAdd(sums + c, (double)t0);
        if (two) atomicAdd(sums + c_mod + c, (double)t1);
    }
}

// Thread layout shared by the row-wise kernels: a thread owns one 8-channel chunk (fixed) and walks rows with a
// division-free cursor (n, hp, wp advance by the decomposed row step); R rows are loaded back to back per iteration.
struct RowWalk { int chunk, cpr, row0, row_step; };
__device__ __forceinline__ RowWalk row_walk(int c_total) {
    RowWalk r;
    r.cpr = c_total >> 3;                                    // chunks per row (<= 256)
    const int rpi = EW_THREADS / r.cpr;                      // rows per block iteration (>= 1)
    r.chunk = threadIdx.x % r.cpr;
    const int rl = threadIdx.x / r.cpr;
    r.row0 = rl < rpi ? (int)blockIdx.x * rpi + rl : -1;          // -1: idle thread (256 not a multiple of cpr)
    r.row_step = gridDim.x * rpi;
    return r;
}
struct ColInfo { int col, cb, a, b; };                       // column, channel (col % c_mod), sub-pixel of CB_OUT_UPSAMPLE
__device__ __forceinline__ ColInfo col_info(const MapP& m, int col) {
    ColInfo c;
    c.col = col; c.cb = col % m.c_mod; c.a = c.b = 0;
    if (m.y_mode == CB_OUT_UPSAMPLE) { const int ab = col / m.c_mod; c.a = ab / m.up_k; c.b = ab - c.a * m.up_k; }
    return c;
}
struct RowCur { long q; int n, hp, wp, dn, dh, dw, step; };
__device__ __forceinline__ RowCur cur_init(const MapP& m, int q0, int step) {
    RowCur c;
    const int plane = m.Hp * m.Wp;
    c.q = q0; c.step = step;
    c.n = q0 / plane; int rem = q0 - c.n * plane; c.hp = rem / m.Wp; c.wp = rem - c.hp * m.Wp;
    c.dn = step / plane; rem = step - c.dn * plane; c.dh = rem / m.Wp; c.dw = rem - c.dh * m.Wp;
    return c;
}
__device__ __forceinline__ void cur_next(const MapP& m, RowCur& c) {
    c.q += c.step;
    c.wp += c.dw; if (c.wp >= m.Wp) { c.wp -= m.Wp; ++c.hp; }
    c.hp += c.dh; if (c.hp >= m.Hp) { c.hp -= m.Hp; ++c.n; }
    c.n += c.dn;
}
// y-side element offset of the cursor's row for this thread's 8 channels; -1 = halo row or past the end
__device__ __forceinline__ long y_off_cur(const MapP& m, const RowCur& c, const ColInfo& ci) {
    if (c.q >= m.rows_total || c.hp < 1 || c.hp > m.Hp - 2 || c.wp < 1 || c.wp > m.Wp - 2) return -1;
    const int h = c.hp - 1, w = c.wp - 1;
    if (m.y_mode == CB_OUT_PF) return c.q * m.y_pitch + m.y_ch_off + ci.col;
    if (m.y_mode == CB_OUT_PS) {
        const int ph = (h & 1) * 2 + (w & 1);
        const long row = (long)ph * m.y_plane_rows + (long)(c.n * m.y_Hp + (h >> 1) + 1) * m.y_Wp + (w >> 1) + 1;
        return row * m.y_pitch + m.y_ch_off + ci.col;
    }
    const long row = (long)(c.n * m.y_Hp + m.up_k * h + ci.a + 1) * m.y_Wp + (m.up_k * w + ci.b + 1);
    return row * m.y_pitch + m.y_ch_off + ci.cb;
}
__device__ __forceinline__ long z_off_cur(const MapP& m, const RowCur& c, const ColInfo& ci, long yo) {
    if (!m.z_at_y) return c.q * m.c_total + ci.col;
    return (yo - m.y_ch_off - ci.cb) / m.y_pitch * m.z_pitch + ci.cb;
}

// ------------------------------------------------------------------------------------------------ BatchNorm kernels
// All four are latency-bound streaming kernels: at 100-130 registers only 16 warps fit on an SM, so every thread keeps R
// rows x up to 3 tensors of 16-byte loads in flight (raw bf16 words, converted only when consumed): ~50-100 KB per SM, what
// HBM3e needs to stay busy (ncu of the first version: 16 KB in flight, 1.9 TB/s).  Loads are unconditional - rows past the
// end or halo rows re-read a valid dummy row and are masked - so that they issue back to back.
// Per-channel constants live in shared memory TRANSPOSED: channel c -> slot (c % 8) * (c_mod / 8) + c / 8, so that the 32
// lanes of a warp (consecutive 8-channel chunks) read consecutive words.  Channel-major storage made every such load an
// 8-way bank conflict and the kernels shared-memory bound (ncu: 9 wavefronts per 8-channel chunk in bn_bwd_apply,
// short-scoreboard / MIO-throttle stalls; profiles/r2_ncu_bn_small.csv).
__device__ __forceinline__ int tslot(int c, int nck) { return (c & 7) * nck + (c >> 3); }
template <bool LO> struct Raw8 { uint4 hi, lo; };
template <bool LO>
__device__ __forceinline__ void ld_raw(const __nv_bfloat16* base, long off, long lo_off, Raw8<LO>& r) {
    r.hi = __ldg(reinterpret_cast<const uint4*>(base + off));
    if (LO) r.lo = __ldg(reinterpret_cast<const uint4*>(base + lo_off + off));
}
template <bool LO>
__device__ __forceinline__ void cvt_raw(const Raw8<LO>& r, float (&v)[8]) {
    v[0] = bf16_lo(r.hi.x); v[1] = bf16_hi(r.hi.x); v[2] = bf16_lo(r.hi.y); v[3] = bf16_hi(r.hi.y);
    v[4] = bf16_lo(r.hi.z); v[5] = bf16_hi(r.hi.z); v[6] = bf16_lo(r.hi.w); v[7] = bf16_hi(r.hi.w);
    if (LO) {
        v[0] += bf16_lo(r.lo.x); v[1] += bf16_hi(r.lo.x); v[2] += bf16_lo(r.lo.y); v[3] += bf16_hi(r.lo.y);
        v[4] += bf16_lo(r.lo.z); v[5] += bf16_hi(r.lo.z); v[6] += bf16_lo(r.lo.w); v[7] += bf16_hi(r.lo.w);
    }
}

struct BnFin {                        // arguments of the BatchNorm finalize step (cb_bn_finalize)
    int c; double count; float eps, momentum;
    const float *gamma, *beta;
    float *running_mean, *running_var, *scale, *shift, *mean_out, *inv_out;
    int* counter;                     // cb_bn_stats_finalize: CTA ticket counter (zero between launches)
};
__device__ __forceinline__ void bn_finalize_channel(const BnFin& f, int i, double s0, double s1) {
    if (f.gamma == nullptr) {                                            // no BatchNorm: identity + bias
        f.scale[i] = 1.f;
        f.shift[i] = f.beta ? f.beta[i] : 0.f;
        return;
    }
    const double mean = s0 / f.count;
    double var = s1 / f.count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double inv = 1.0 / sqrt(var + (double)f.eps);
    const double sc = (double)f.gamma[i] * inv;
    f.scale[i] = (float)sc;
    f.shift[i] = (float)((double)f.beta[i] - mean * sc);
    if (f.mean_out) f.mean_out[i] = (float)mean;
    if (f.inv_out) f.inv_out[i] = (float)inv;
    if (f.running_mean) f.running_mean[i] = (float)((1.0 - f.momentum) * f.running_mean[i] + f.momentum * mean);
    if (f.running_var) {
        const double unb = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
        f.running_var[i] = (float)((1.0 - f.momentum) * f.running_var[i] + f.momentum * unb);
    }
}

// One cursor per thread walks its rows (row_step apart); R rows are loaded back to back per iteration, then consumed.
// Which tensors a launch reads is a template parameter, so a variant only holds registers for the loads it issues and the
// two-tensor variants run with R = 8 (64-128 B in flight per thread and tensor set).
template <int R, bool LO>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_stats_kernel(const __nv_bfloat16* __restrict__ z, long z_lo_off,
                                                              const MapP m, double* __restrict__ sums, const BnFin fin) {
    __shared__ float s_part[2 * 2048];
    __shared__ int s_last;
    pdl_wait();
    const RowWalk rw = row_walk(m.c_total);
    const ColInfo ci = col_info(m, rw.chunk * 8);
    float a0[8] = {}, a1[8] = {};
    if (rw.row0 >= 0) {
        RowCur c = cur_init(m, rw.row0, rw.row_step);
        while (c.q < m.rows_total) {
            Raw8<LO> raw[R];
            unsigned valid = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool ok = y_off_cur(m, c, ci) >= 0;                                   // halo rows hold zeros anyway
                valid |= (unsigned)ok << r;
                ld_raw<LO>(z, (c.q < m.rows_total ? c.q : 0) * m.c_total + ci.col, z_lo_off, raw[r]);
                cur_next(m, c);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float k = (valid >> r) & 1u ? 1.f : 0.f;
                float v[8];
                cvt_raw<LO>(raw[r], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float t = v[j] * k; a0[j] += t; a1[j] = fmaf(t, t, a1[j]); }
            }
        }
    }
    block_reduce_to_global(a0, a1, m.c_total, m.c_mod, true, s_part, sums);
    if (fin.counter == nullptr) return;
    // cb_bn_stats_finalize: the CTA that takes the last ticket sees every CTA's sums and closes the statistics
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fin.counter, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = threadIdx.x; i < fin.c; i += EW_THREADS) bn_finalize_channel(fin, i, __ldcg(sums + i), __ldcg(sums + fin.c + i));
    if (threadIdx.x == 0) *fin.counter = 0;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const BnFin f) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.c) return;
    bn_finalize_channel(f, i, f.gamma ? sums[i] : 0.0, f.gamma ? sums[f.c + i] : 0.0);
}

// EXTRA = a second BatchNorm branch (zb) and / or a residual is added
template <int R, bool LO, bool EXTRA>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_apply_kernel(
    const __nv_bfloat16* __restrict__ z, long z_lo_off, const float* __restrict__ scale, const float* __restrict__ shift,
    const __nv_bfloat16* __restrict__ zb, long zb_lo_off, const float* __restrict__ scale_b, const float* __restrict__ shift_b,
    const __nv_bfloat16* __restrict__ res, int res_pitch, long res_lo_off, int relu, const MapP m,
    __nv_bfloat16* __restrict__ y, long y_lo_off) {
    __shared__ float s_sc[256], s_sh[256], s_scb[256], s_shb[256];
    pdl_wait();
    const int nck = m.c_mod >> 3;
    for (int c = threadIdx.x; c < m.c_mod; c += EW_THREADS) {
        const int t = tslot(c, nck);
        s_sc[t] = scale[c]; s_sh[t] = shift[c];
        s_scb[t] = (EXTRA && zb) ? scale_b[c] : 0.f; s_shb[t] = (EXTRA && zb) ? shift_b[c] : 0.f;
    }
    __syncthreads();
    const RowWalk rw = row_walk(m.c_total);
    if (rw.row0 < 0) return;
    const ColInfo ci = col_info(m, rw.chunk * 8);
    const int col = ci.col;
    const float* sc = s_sc + (ci.cb >> 3);
    const float* sh = s_sh + (ci.cb >> 3);
    const float* scb = s_scb + (ci.cb >> 3);
    const float* shb = s_shb + (ci.cb >> 3);
    RowCur c = cur_init(m, rw.row0, rw.row_step);
    while (c.q < m.rows_total) {
        long yo[R];
        Raw8<LO> rz[R], rb[EXTRA ? R : 1], rr[EXTRA ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            yo[r] = y_off_cur(m, c, ci);
            const long qz = yo[r] >= 0 ? c.q : 0;
            ld_raw<LO>(z, qz * m.c_total + col, z_lo_off, rz[r]);
            if (EXTRA) {
                if (zb) ld_raw<LO>(zb, qz * m.c_total + col, zb_lo_off, rb[r]);
                if (res) ld_raw<LO>(res, qz * (long)res_pitch + col, res_lo_off, rr[r]);
            }
            cur_next(m, c);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (yo[r] >= 0) {
                float v[8], u[8];
                cvt_raw<LO>(rz[r], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j * nck], sh[j * nck]);
                if (EXTRA) {
                    if (zb) {
                        cvt_raw<LO>(rb[r], u);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += fmaf(u[j], scb[j * nck], shb[j * nck]);
                    }
                    if (res) {
                        cvt_raw<LO>(rr[r], u);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += u[j];
                    }
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                store8(y, yo[r], LO ? y_lo_off : 0, v);
            }
        }
    }
}

// ReLU mask: from y (the stored activation; YMASK) or - mask_scale != NULL, the plain conv + BN + ReLU case - recomputed from
// the z that is loaded anyway (sign of z*scale + shift), which saves one tensor read in both backward passes.
template <int R, bool LO, bool YMASK>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ dy, long dy_lo_off, const __nv_bfloat16* __restrict__ y, long y_lo_off, int relu,
    const __nv_bfloat16* __restrict__ z, long z_lo_off, const float* __restrict__ mean, const float* __restrict__ inv_std,
    const float* __restrict__ mask_scale, const float* __restrict__ mask_shift, const MapP m, double* __restrict__ sums) {
    __shared__ float s_part[2 * 2048];
    __shared__ float s_mu[256], s_ms[256], s_mh[256];
    pdl_wait();
    const RowWalk rw = row_walk(m.c_total);
    const ColInfo ci = col_info(m, rw.chunk * 8);
    const bool has_bn = mean != nullptr;
    const bool zmask = !YMASK && relu && mask_scale != nullptr && has_bn;
    const int nck = m.c_mod >> 3;
    for (int c = threadIdx.x; c < m.c_mod; c += EW_THREADS) {
        const int t = tslot(c, nck);
        s_mu[t] = has_bn ? mean[c] : 0.f;
        s_ms[t] = zmask ? mask_scale[c] : 0.f;
        s_mh[t] = zmask ? mask_shift[c] : 0.f;
    }
    __syncthreads();
    const float* mu = s_mu + (ci.cb >> 3);
    const float* ms = s_ms + (ci.cb >> 3);
    const float* mh = s_mh + (ci.cb >> 3);
    float a0[8] = {}, a1[8] = {};
    if (rw.row0 >= 0) {
        RowCur c = cur_init(m, rw.row0, rw.row_step);
        while (c.q < m.rows_total) {
            unsigned valid = 0;
            Raw8<LO> rg[R], ry[YMASK ? R : 1], rz[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const long yo = y_off_cur(m, c, ci);
                valid |= (unsigned)(yo >= 0) << r;
                const long yl = yo >= 0 ? yo : (long)(m.y_ch_off + ci.cb);
                ld_raw<LO>(dy, yl, dy_lo_off, rg[r]);
                if (YMASK) ld_raw<LO>(y, yl, y_lo_off, ry[r]);
                if (has_bn) ld_raw<LO>(z, yo >= 0 ? z_off_cur(m, c, ci, yo) : (long)ci.cb, z_lo_off, rz[r]);
                cur_next(m, c);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((valid >> r) & 1u) {
                    float g[8], v[8], yv[8];
                    cvt_raw<LO>(rg[r], g);
                    if (has_bn) cvt_raw<LO>(rz[r], v);
                    if (YMASK) cvt_raw<LO>(ry[r], yv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float gj = g[j];
                        if (zmask) gj = fmaf(v[j], ms[j * nck], mh[j * nck]) > 0.f ? gj : 0.f;
                        if (YMASK) gj = yv[j] > 0.f ? gj : 0.f;
                        a0[j] += gj;
                        if (has_bn) a1[j] = fmaf(gj, v[j] - mu[j * nck], a1[j]);       // x_hat = (z - mean) * inv_std: scaled below
                    }
                }
            }
        }
    }
    if (has_bn) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a1[j] *= inv_std[ci.cb + j];
    }
    block_reduce_to_global(a0, a1, m.c_total, m.c_mod, has_bn, s_part, sums);
}

template <int R, bool LO, bool YMASK>
__global__ void __launch_bounds__(EW_THREADS, 2) bn_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ dy, long dy_lo_off, const __nv_bfloat16* __restrict__ y, long y_lo_off, int relu,
    const __nv_bfloat16* __restrict__ z, long z_lo_off, const float* __restrict__ mean, const float* __restrict__ inv_std,
    const float* __restrict__ gamma, const float* __restrict__ mask_scale, const float* __restrict__ mask_shift,
    const double* __restrict__ sums, double count, const MapP m,
    __nv_bfloat16* __restrict__ dz, long dz_lo_off, __nv_bfloat16* __restrict__ dsum, long dsum_lo_off,
    float* __restrict__ d_gamma, float* __restrict__ d_beta) {
    // dz = gamma*inv * (g - s0/n - (z - mean)*inv * s1/n) = A*g + B*z + C per channel
    __shared__ float s_A[256], s_B[256], s_C[256], s_ms[256], s_mh[256];
    pdl_wait();
    const bool has_bn = mean != nullptr;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < m.c_mod; c += EW_THREADS) {
            if (d_beta) d_beta[c] = (float)sums[c];
            if (d_gamma && has_bn) d_gamma[c] = (float)sums[m.c_mod + c];
        }
    }
    const bool zmask = !YMASK && relu && mask_scale != nullptr && has_bn;
    const int nck = m.c_mod >> 3;
    for (int c = threadIdx.x; c < m.c_mod; c += EW_THREADS) {
        float A = 1.f, B = 0.f, Cc = 0.f;
        if (has_bn) {
            const double iv = (double)inv_std[c], gi = (double)gamma[c] * iv;
            const double k0 = sums[c] / count, k1 = sums[m.c_mod + c] / count;
            A = (float)gi;
            B = (float)(-gi * iv * k1);
            Cc = (float)(-gi * k0 + gi * iv * k1 * (double)mean[c]);
        }
        const int t = tslot(c, nck);
        s_A[t] = A; s_B[t] = B; s_C[t] = Cc;
        s_ms[t] = zmask ? mask_scale[c] : 0.f;
        s_mh[t] = zmask ? mask_shift[c] : 0.f;
    }
    __syncthreads();
    const RowWalk rw = row_walk(m.c_total);
    if (rw.row0 < 0) return;
    const ColInfo ci = col_info(m, rw.chunk * 8);
    const int col = ci.col;
    const float* cA = s_A + (ci.cb >> 3);
    const float* cB = s_B + (ci.cb >> 3);
    const float* cC = s_C + (ci.cb >> 3);
    const float* ms = s_ms + (ci.cb >> 3);
    const float* mh = s_mh + (ci.cb >> 3);
    RowCur c = cur_init(m, rw.row0, rw.row_step);
    while (c.q < m.rows_total) {
        int qv[R];                                        // GEMM row of each loaded row (rows_total < 2^31), -1 = masked
        Raw8<LO> rg[R], ry[YMASK ? R : 1], rz[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long yo = y_off_cur(m, c, ci);
            qv[r] = yo >= 0 ? (int)c.q : -1;
            const long yl = yo >= 0 ? yo : (long)(m.y_ch_off + ci.cb);
            ld_raw<LO>(dy, yl, dy_lo_off, rg[r]);
            if (YMASK) ld_raw<LO>(y, yl, y_lo_off, ry[r]);
            if (has_bn) ld_raw<LO>(z, yo >= 0 ? z_off_cur(m, c, ci, yo) : (long)ci.cb, z_lo_off, rz[r]);
            cur_next(m, c);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (qv[r] >= 0) {
                float g[8], v[8], yv[8];
                cvt_raw<LO>(rg[r], g);
                if (has_bn) cvt_raw<LO>(rz[r], v);
                if (YMASK) cvt_raw<LO>(ry[r], yv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (zmask) g[j] = fmaf(v[j], ms[j * nck], mh[j * nck]) > 0.f ? g[j] : 0.f;
                    if (YMASK) g[j] = yv[j] > 0.f ? g[j] : 0.f;
                }
                if (dsum) store8(dsum, (long)qv[r] * m.c_total + col, LO ? dsum_lo_off : 0, g);
                if (has_bn) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = fmaf(cA[j * nck], g[j], fmaf(cB[j * nck], v[j], cC[j * nck]));
                }
                store8(dz, (long)qv[r] * m.c_total + col, LO ? dz_lo_off : 0, g);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ heads gradient pack
struct HeadsP { const float* g[CB_MAX_HEADS]; int cn[CB_MAX_HEADS]; int c0[CB_MAX_HEADS]; int n_heads, total; };

__global__ void __launch_bounds__(256) heads_grad_pack_kernel(const HeadsP hd, int n, int H, int W,
                                                              __nv_bfloat16* __restrict__ out, long lo_off,
                                                              float* __restrict__ d_bias) {
    __shared__ float s_b[64];
    pdl_wait();
    if (threadIdx.x < 64) s_b[threadIdx.x] = 0.f;
    __syncthreads();
    const int Hp = H + 2, Wp = W + 2;
    const long rows = (long)n * Hp * Wp;
    float bsum[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bsum[j] = 0.f;
    for (long q = (long)blockIdx.x * 256 + threadIdx.x; q < rows; q += (long)gridDim.x * 256) {
        const int img = (int)(q / (Hp * Wp));
        const int rem = (int)(q - (long)img * Hp * Wp);
        const int hp = rem / Wp, wp = rem - hp * Wp;
        if (hp < 1 || hp > H || wp < 1 || wp > W) continue;
        const long pix = (long)(hp - 1) * W + (wp - 1);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
        for (int s = 0; s < CB_MAX_HEADS; ++s) {
            if (s < hd.n_heads) {
                const float* src = hd.g[s] + (long)img * hd.cn[s] * H * W + pix;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j >= hd.c0[s] && j < hd.c0[s] + hd.cn[s]) v[j] = __ldg(src + (long)(j - hd.c0[s]) * H * W);
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) bsum[j] += v[j];
        const float zero[8] = {};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float w8[8] = {v[8 * t], v[8 * t + 1], v[8 * t + 2], v[8 * t + 3], v[8 * t + 4], v[8 * t + 5], v[8 * t + 6], v[8 * t + 7]};
            store8(out, q * 64 + 8 * t, lo_off, w8);
            store8(out, q * 64 + 32 + 8 * t, lo_off, zero);
        }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float s = bsum[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((threadIdx.x & 31) == 0 && j < hd.total) atomicAdd(&s_b[j], s);
    }
    __syncthreads();
    if (threadIdx.x < hd.total) atomicAdd(d_bias + threadIdx.x, s_b[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ fusion backward
constexpr int FB_MAX_AGENTS = 8;
struct FuseBG {
    int H, W, C, in_ps, Hp, Wp;
    long plane_rows;
    float inv_sqrt_c;
};
__device__ __forceinline__ long fb_in_row(const FuseBG& g, int agent, int y, int x) {
    if (!g.in_ps) return ((long)agent * g.Hp + y + 1) * g.Wp + x + 1;
    const int ph = (y & 1) * 2 + (x & 1);
    return (long)ph * g.plane_rows + ((long)agent * g.Hp + (y >> 1) + 1) * g.Wp + (x >> 1) + 1;
}
__device__ __forceinline__ void red_add_v4f(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// LPP lanes per pixel, 8 channels per lane (C = 8*LPP).  A warp covers 32/LPP consecutive pixels of a row-major scan.
// Register budget (the first version held every agent's warped vector, tap record and gradient at once: 213-255 registers,
// one CTA per SM): only the warped vectors v_j stay live; the tap geometry is recomputed for the scatter pass (a dozen
// float ops per agent) and each agent's gradient is formed and scattered on the fly.
struct FbTap { int x0, y0; float wx, wy; };
__device__ __forceinline__ FbTap fb_tap(const double* __restrict__ A, double xs, double ys, const FuseBG& g) {
    FbTap t;
    const float gx = (float)(A[0] * xs + A[1] * ys + A[2]);
    const float gy = (float)(A[3] * xs + A[4] * ys + A[5]);
    const float ix = ((gx + 1.f) * g.W - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * g.H - 1.f) * 0.5f;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    t.wx = ix - fx0; t.wy = iy - fy0;
    t.x0 = (int)fminf(fmaxf(fx0, -2.f), (float)g.W + 1.f);
    t.y0 = (int)fminf(fmaxf(fy0, -2.f), (float)g.H + 1.f);
    return t;
}

template <int LPP, int MAXN>
__global__ void __launch_bounds__(256, 2) warp_att_fuse_bwd_kernel(
    const __nv_bfloat16* __restrict__ feat, long in_lo_off, const double* __restrict__ affine,
    const int* __restrict__ agent_off, int n_scenes, int L, const FuseBG g, int method,
    const __nv_bfloat16* __restrict__ dfused, long dfused_lo_off, float* __restrict__ dfeat) {
    pdl_wait();
    constexpr int PPW = 32 / LPP;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPP, cl = lane % LPP;                         // pixel within the warp, channel chunk
    const int c0 = cl * 8;
    const long total = (long)n_scenes * g.H * g.W;
    const long wid = (long)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (long)gridDim.x * 8;
    for (long base = wid * PPW; base < total; base += nw * PPW) {
        const long pix = base + sub;
        const bool live = pix < total;
        const long pc = live ? pix : total - 1;                          // idle lanes shadow the last pixel (no stores)
        const int b = (int)(pc / (g.H * g.W));
        const int rem = (int)(pc - (long)b * g.H * g.W);
        const int h = rem / g.W, w = rem - h * g.W;
        const int a0 = agent_off[b];
        int n = agent_off[b + 1] - a0;
        n = n < MAXN ? n : MAXN;
        const double xs = (2.0 * w + 1.0) / g.W - 1.0;
        const double ys = (2.0 * h + 1.0) / g.H - 1.0;
        const double* Ab = affine + (long)b * L * 6;
        float v[MAXN][8];
#pragma unroll
        for (int j = 0; j < MAXN; ++j) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
            if (j < n) {
                const FbTap t = fb_tap(Ab + j * 6, xs, ys, g);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
                    const float wt = ((k & 1) ? t.wx : 1.f - t.wx) * ((k >> 1) ? t.wy : 1.f - t.wy);
                    if (xx >= 0 && xx < g.W && yy >= 0 && yy < g.H) {
                        float u[8];
                        load8(feat, fb_in_row(g, a0 + j, yy, xx) * g.C + c0, in_lo_off, u);
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[j][e] = fmaf(wt, u[e], v[j][e]);
                    }
                }
            }
        }
        float go[8];
        load8(dfused, (((long)b * (g.H + 2) + h + 1) * (g.W + 2) + w + 1) * g.C + c0, dfused_lo_off, go);
        // per-agent coefficients of  dv_j = ca_j * go + cs_j * v_0  (+ d0 for the ego), or the arg-max mask of MaxFusion
        float ca[MAXN], cs[MAXN], d0[8];
        unsigned amask[MAXN];
#pragma unroll
        for (int e = 0; e < 8; ++e) d0[e] = 0.f;
#pragma unroll
        for (int j = 0; j < MAXN; ++j) { ca[j] = 0.f; cs[j] = 0.f; amask[j] = 0u; }
        if (method == 1) {                                               // MaxFusion: first arg-max agent per channel
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float best = v[0][e];
                int bj = 0;
#pragma unroll
                for (int j = 1; j < MAXN; ++j)
                    if (j < n && v[j][e] > best) { best = v[j][e]; bj = j; }
#pragma unroll
                for (int j = 0; j < MAXN; ++j) amask[j] |= (j == bj) ? (1u << e) : 0u;
            }
        } else {
            float da[MAXN];
            float smax = -INFINITY;
#pragma unroll
            for (int j = 0; j < MAXN; ++j) {
                da[j] = 0.f;
                if (j < n) {
                    float d = 0.f, e2 = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) { d = fmaf(v[0][e], v[j][e], d); e2 = fmaf(go[e], v[j][e], e2); }
#pragma unroll
                    for (int o = LPP / 2; o > 0; o >>= 1) {
                        d += __shfl_xor_sync(0xffffffffu, d, o);
                        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                    }
                    ca[j] = d * g.inv_sqrt_c;                            // score for now
                    da[j] = e2;
                    smax = fmaxf(smax, ca[j]);
                }
            }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < MAXN; ++j)
                if (j < n) { ca[j] = __expf(ca[j] - smax); den += ca[j]; }
            const float rden = 1.f / den;
            float tsum = 0.f;
#pragma unroll
            for (int j = 0; j < MAXN; ++j)
                if (j < n) { ca[j] *= rden; tsum = fmaf(ca[j], da[j], tsum); }           // ca = attention weights now
#pragma unroll
            for (int j = 0; j < MAXN; ++j) {
                if (j < n) {
                    cs[j] = ca[j] * (da[j] - tsum) * g.inv_sqrt_c;                         // d loss / d score_j
#pragma unroll
                    for (int e = 0; e < 8; ++e) d0[e] = fmaf(cs[j], v[j][e], d0[e]);
                }
            }
        }
        if (!live) continue;
#pragma unroll
        for (int j = 0; j < MAXN; ++j) {
            if (j < n) {
                float dv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (method == 1) dv[e] = (amask[j] >> e) & 1u ? go[e] : 0.f;
                    else dv[e] = fmaf(ca[j], go[e], cs[j] * v[0][e]) + (j == 0 ? d0[e] : 0.f);
                }
                const FbTap t = fb_tap(Ab + j * 6, xs, ys, g);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xx = t.x0 + (k & 1), yy = t.y0 + (k >> 1);
                    const float wt = ((k & 1) ? t.wx : 1.f - t.wx) * ((k >> 1) ? t.wy : 1.f - t.wy);
                    if (xx >= 0 && xx < g.W && yy >= 0 && yy < g.H && wt != 0.f) {
                        float* dst = dfeat + ((((long)(a0 + j) * g.H + yy) * g.W + xx) * g.C + c0);
                        red_add_v4f(dst, wt * dv[0], wt * dv[1], wt * dv[2], wt * dv[3]);
                        red_add_v4f(dst + 4, wt * dv[4], wt * dv[5], wt * dv[6], wt * dv[7]);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) grad_combine_kernel(const float* __restrict__ acc, const __nv_bfloat16* __restrict__ addend,
                                                           long addend_lo_off, int to_ps, int n_cap, int n, int H, int W, int C,
                                                           __nv_bfloat16* __restrict__ out, long out_lo_off) {
    pdl_wait();
    const int cpr = C >> 3;
    const long total = (long)n * H * W * cpr;
    const int Hp = to_ps ? (H + 1) / 2 + 2 : H + 2, Wp = to_ps ? (W + 1) / 2 + 2 : W + 2;
    const long plane_rows = (long)n_cap * Hp * Wp;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int ch = (int)(i % cpr);
        const long pix = i / cpr;
        const int w = (int)(pix % W);
        const long t = pix / W;
        const int h = (int)(t % H), a = (int)(t / H);
        long row;
        if (to_ps) {
            const int ph = (h & 1) * 2 + (w & 1);
            row = (long)ph * plane_rows + ((long)a * Hp + (h >> 1) + 1) * Wp + (w >> 1) + 1;
        } else {
            row = ((long)a * Hp + h + 1) * Wp + w + 1;
        }
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(acc + pix * C + ch * 8));
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(acc + pix * C + ch * 8) + 1);
        float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        if (addend) {
            float u[8];
            load8(addend, row * C + ch * 8, addend_lo_off, u);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += u[j];
        }
        store8(out, row * C + ch * 8, out_lo_off, v);
    }
}

// ------------------------------------------------------------------------------------------------ PFN (training)
struct PfnT { float vx, vy, vz, offx, offy, offz; };

// The 10 augmented features of the pillar's points (pillar_vfe.py:118-137), float32 arithmetic like the reference:
// lane = slot.  Returns them in f[10] (zeros for padded slots).
__device__ __forceinline__ void pfn_features(const float4* __restrict__ voxels, const int* __restrict__ coords, long m, int n,
                                             int max_pts, const PfnT& t, int lane, float (&f)[10]) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < n && lane < max_pts) p = __ldg(voxels + m * max_pts + lane);
    float sx = p.x, sy = p.y, sz = p.z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const float fn = (float)n;
    const float mx = sx / fn, my = sy / fn, mz = sz / fn;
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + m);      // [agent, z, y, x]
    const float cx = (float)c.w * t.vx + t.offx, cy = (float)c.z * t.vy + t.offy, cz = (float)c.y * t.vz + t.offz;
    const bool ok = lane < n;
    f[0] = ok ? p.x : 0.f; f[1] = ok ? p.y : 0.f; f[2] = ok ? p.z : 0.f; f[3] = ok ? p.w : 0.f;
    f[4] = ok ? p.x - mx : 0.f; f[5] = ok ? p.y - my : 0.f; f[6] = ok ? p.z - mz : 0.f;
    f[7] = ok ? p.x - cx : 0.f; f[8] = ok ? p.y - cy : 0.f; f[9] = ok ? p.z - cz : 0.f;
}

constexpr int PFN_NQ = 65;      // 10 feature sums + 55 upper-triangle Gram entries

__global__ void __launch_bounds__(256) pfn_train_stats_kernel(const float4* __restrict__ voxels, const int* __restrict__ coords,
                                                              const int* __restrict__ num_points, int n_rows_cap,
                                                              const int* __restrict__ n_voxels_dev, int max_pts, const PfnT t,
                                                              double* __restrict__ sums) {
    __shared__ float s_f[8][32][10];
    __shared__ double s_red[PFN_NQ];
    pdl_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int n_rows = n_voxels_dev ? *n_voxels_dev : n_rows_cap;
    if (n_rows > n_rows_cap) n_rows = n_rows_cap;
    for (int i = threadIdx.x; i < PFN_NQ; i += 256) s_red[i] = 0.0;
    __syncthreads();
    // lane L owns quantities L, L+32, L+64 (< 65): index -> (i, j) pair of the Gram matrix or a plain feature sum
    int qi[3], qj[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int qd = lane + 32 * r;
        qi[r] = -1; qj[r] = 0;
        if (qd < 10) { qi[r] = qd; qj[r] = -1; }
        else if (qd < PFN_NQ) {
            int k = qd - 10, i = 0;
            while (k >= 10 - i) { k -= 10 - i; ++i; }
            qi[r] = i; qj[r] = i + k;
        }
    }
    double acc[3] = {0.0, 0.0, 0.0};
    for (long m = (long)blockIdx.x * 8 + wib; m < n_rows; m += (long)gridDim.x * 8) {
        int n = __ldg(num_points + m);
        n = n < max_pts ? n : max_pts;
        float f[10];
        pfn_features(voxels, coords, m, n, max_pts, t, lane, f);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 10; ++e) s_f[wib][lane][e] = f[e];
        __syncwarp();
        for (int k = 0; k < n; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (qi[r] >= 0) {
                    const float a = s_f[wib][k][qi[r]];
                    const float bb = qj[r] >= 0 ? s_f[wib][k][qj[r]] : 1.f;
                    acc[r] += (double)a * (double)bb;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        if (qi[r] >= 0) atomicAdd(&s_red[lane + 32 * r], acc[r]);
    __syncthreads();
    for (int i = threadIdx.x; i < PFN_NQ; i += 256) atomicAdd(sums + i, s_red[i]);
}

// Gram entry (i, j) from the packed upper triangle
__device__ __forceinline__ double gram_at(const double* __restrict__ sums, int i, int j) {
    if (i > j) { const int t = i; i = j; j = t; }
    int off = 10;
    for (int r = 0; r < i; ++r) off += 10 - r;
    return sums[off + (j - i)];
}

__global__ void pfn_train_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ n_voxels_dev, int n_rows_cap,
                                          int max_pts, const float* __restrict__ w, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps, float momentum,
                                          float* __restrict__ running_mean, float* __restrict__ running_var,
                                          float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                          float* __restrict__ inv_out) {
    pdl_wait();
    const int c = threadIdx.x;
    if (c >= 64) return;
    int n_rows = n_voxels_dev ? *n_voxels_dev : n_rows_cap;
    if (n_rows > n_rows_cap) n_rows = n_rows_cap;
    const double count = (double)n_rows * (double)max_pts;               // every slot, padded ones included
    double wl[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) wl[i] = (double)w[c * 10 + i];
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < 10; ++i) {
        s1 += wl[i] * sums[i];
        for (int j = 0; j < 10; ++j) s2 += wl[i] * wl[j] * gram_at(sums, i, j);
    }
    const double mean = count > 0 ? s1 / count : 0.0;
    double var = count > 0 ? s2 / count - mean * mean : 0.0;
    if (var < 0.0) var = 0.0;
    const double inv = 1.0 / sqrt(var + (double)eps);
    const double sc = (double)gamma[c] * inv;
    scale[c] = (float)sc;
    shift[c] = (float)((double)beta[c] - mean * sc);
    if (mean_out) mean_out[c] = (float)mean;
    if (inv_out) inv_out[c] = (float)inv;
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
    if (running_var) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
}

struct CanvasG { int ny, nx, Hq, Wq; long plane_rows; };

// lane owns channels (lane, lane + 32); bsum layout: [0,64) d_beta, [64,128) d_gamma, [128, 128 + 640) sum dy*f [c][10]
__global__ void __launch_bounds__(256, 4) pfn_bwd_kernel(const float4* __restrict__ voxels, const int* __restrict__ coords,
                                                      const int* __restrict__ num_points, int n_rows_cap,
                                                      const int* __restrict__ n_voxels_dev, int max_pts,
                                                      const float* __restrict__ w, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, const float* __restrict__ mean,
                                                      const float* __restrict__ inv_std, const PfnT t,
                                                      const __nv_bfloat16* __restrict__ dcanvas, long dc_lo_off, const CanvasG cg,
                                                      double* __restrict__ bsum) {
    __shared__ float s_f[8][32][24];                                     // features [..][10]; reused for the final reduction
    pdl_wait();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int n_rows = n_voxels_dev ? *n_voxels_dev : n_rows_cap;
    if (n_rows > n_rows_cap) n_rows = n_rows_cap;
    // the kernel is bound by the latency of its dependent loads (count -> points -> coords -> gradient row): the linear
    // layer's weights live in shared memory (row stride 11: conflict-free) so that four CTAs fit on an SM
    __shared__ float s_w[64 * 11];
    for (int i = threadIdx.x; i < 640; i += 256) s_w[(i / 10) * 11 + (i % 10)] = w[i];
    __syncthreads();
    float sc[2], sh[2], mu[2], iv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int c = lane + 32 * r;
        sc[r] = scale[c]; sh[r] = shift[c]; mu[r] = mean[c]; iv[r] = inv_std[c];
    }
    float acc[2][12];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[r][i] = 0.f;
    for (long m = (long)blockIdx.x * 8 + wib; m < n_rows; m += (long)gridDim.x * 8) {
        int n = __ldg(num_points + m);
        n = n < max_pts ? n : max_pts;
        float f[10];
        pfn_features(voxels, coords, m, n, max_pts, t, lane, f);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 10; ++e) s_f[wib][lane][e] = f[e];
        __syncwarp();
        const int4 c4 = __ldg(reinterpret_cast<const int4*>(coords) + m);
        const int ya = c4.z, xa = c4.w, ag = c4.x;
        const int ph = (ya & 1) * 2 + (xa & 1);
        const long row = (long)ph * cg.plane_rows + ((long)ag * cg.Hq + (ya >> 1) + 1) * cg.Wq + (xa >> 1) + 1;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int c = lane + 32 * r;
            float g = __bfloat162float(dcanvas[row * 64 + c]);
            if (dc_lo_off != 0) g += __bfloat162float(dcanvas[dc_lo_off + row * 64 + c]);
            // max over the slots (torch.max: first index of the maximum): real points in slot order, then the padded value
            float best = -1.f, best_lin = 0.f;
            int bk = -1;
            for (int k = 0; k < n; ++k) {
                float lin = 0.f;
#pragma unroll
                for (int i = 0; i < 10; ++i) lin = fmaf(s_f[wib][k][i], s_w[c * 11 + i], lin);
                const float yv = fmaxf(fmaf(lin, sc[r], sh[r]), 0.f);
                if (yv > best) { best = yv; bk = k; best_lin = lin; }
            }
            if (n < max_pts) {
                const float yp = fmaxf(sh[r], 0.f);                      // zero features: linear = 0
                if (yp > best) { best = yp; bk = -1; best_lin = 0.f; }
            }
            if (best > 0.f && g != 0.f) {
                acc[r][0] += g;
                acc[r][1] = fmaf(g, (best_lin - mu[r]) * iv[r], acc[r][1]);
                if (bk >= 0) {
#pragma unroll
                    for (int i = 0; i < 10; ++i) acc[r][2 + i] = fmaf(g, s_f[wib][bk][i], acc[r][2 + i]);
                }
            }
        }
    }
    // per-warp fp32 partials (<= a few hundred pillars each) -> shared memory -> one fp64 global atomic per entry and CTA
    // (shared-memory double atomics are CAS loops on sm_100: no periodic flushes through them)
    __syncthreads();
    float* s_acc = &s_f[0][0][0];                                         // reuse: 8 warps x 768 floats = 24 KB
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 12; ++i) s_acc[wib * 768 + (lane + 32 * r) * 12 + i] = acc[r][i];
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 12; i += 256) {
        double t = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += (double)s_acc[w8 * 768 + i];
        const int c = i / 12, k = i - c * 12;
        const int dst = k == 0 ? c : (k == 1 ? 64 + c : 128 + c * 10 + (k - 2));
        atomicAdd(bsum + dst, t);
    }
}

__global__ void pfn_bwd_finalize_kernel(const double* __restrict__ st, const double* __restrict__ bsum,
                                        const int* __restrict__ n_voxels_dev, int n_rows_cap, int max_pts,
                                        const float* __restrict__ w, const float* __restrict__ gamma,
                                        const float* __restrict__ mean, const float* __restrict__ inv_std,
                                        float* __restrict__ d_w, float* __restrict__ d_gamma, float* __restrict__ d_beta) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 640) return;
    const int c = i / 10, f = i - c * 10;
    int n_rows = n_voxels_dev ? *n_voxels_dev : n_rows_cap;
    if (n_rows > n_rows_cap) n_rows = n_rows_cap;
    const double count = (double)n_rows * (double)max_pts;
    const double db = bsum[c], dg = bsum[64 + c], s1 = bsum[128 + c * 10 + f];
    double wg = 0.0;                                                     // (W G)[c][f]
    for (int k = 0; k < 10; ++k) wg += (double)w[c * 10 + k] * gram_at(st, k, f);
    const double inv = (double)inv_std[c], mu = (double)mean[c];
    const double X = inv * (wg - mu * st[f]);                            // sum x_hat_c * feat_f
    const double r = count > 0 ? (double)gamma[c] * inv * (s1 - db / count * st[f] - dg / count * X) : 0.0;
    d_w[c * 10 + f] = (float)r;
    if (f == 0) { d_gamma[c] = (float)dg; d_beta[c] = (float)db; }
}

// ------------------------------------------------------------------------------------------------ layout permutations
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ src, int R0, int K0, long rows, int K,
                                                          long s_r1, long s_r0, long s_k1, long s_k0,
                                                          __nv_bfloat16* __restrict__ dst, int dst_ld, int k_off, int lo_col_off) {
    pdl_wait();
    const long total = rows * K;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = i / K;
        const int k = (int)(i - r * K);
        const long r1 = r / R0, r0 = r - r1 * R0;
        const int k1 = k / K0, k0 = k - k1 * K0;
        const float v = __ldg(src + r1 * s_r1 + r0 * s_r0 + k1 * s_k1 + k0 * s_k0);
        const __nv_bfloat16 hi = __float2bfloat16(v);
        dst[r * dst_ld + k_off + k] = hi;
        if (lo_col_off != 0) dst[r * dst_ld + k_off + lo_col_off + k] = __float2bfloat16(v - __bfloat162float(hi));
    }
}

// One launch for all weight matrices.  Every CTA takes ONE contiguous range of the concatenated element space, so a thread
// changes job at most a couple of times: the job record is looked up (binary search) only on a change and kept in
// registers, and the index arithmetic is 32-bit.  (First version: per-element search + four 64-bit divisions, 0.42 ms for
// 26 M elements.)
__global__ void __launch_bounds__(256) pack_weights_batch_kernel(const cb_pack_job* __restrict__ jobs, int n_jobs, long total) {
    pdl_wait();
    long per = (total + gridDim.x - 1) / gridDim.x;
    per = (per + 255) / 256 * 256;
    const long beg = (long)blockIdx.x * per;
    const long end = beg + per < total ? beg + per : total;
    long cfirst = 0, cend = -1;
    const float* src = nullptr;
    __nv_bfloat16* dst = nullptr;
    unsigned K = 1, R0 = 1, K0 = 1;
    int s_r1 = 0, s_r0 = 0, s_k1 = 0, s_k0 = 0, dst_ld = 0, k_off = 0, lo_col = 0;
    for (long i = beg + threadIdx.x; i < end; i += 256) {
        if (i >= cend) {
            int lo = 0, hi = n_jobs;                                     // last job with first <= i
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (jobs[mid].first <= i) lo = mid; else hi = mid; }
            const cb_pack_job& j = jobs[lo];
            cfirst = j.first;
            cend = lo + 1 < n_jobs ? jobs[lo + 1].first : total;
            src = j.src; dst = (__nv_bfloat16*)j.dst;
            K = (unsigned)j.K; R0 = (unsigned)j.R0; K0 = (unsigned)j.K0;
            s_r1 = (int)j.s_r1; s_r0 = (int)j.s_r0; s_k1 = (int)j.s_k1; s_k0 = (int)j.s_k0;
            dst_ld = j.dst_ld; k_off = j.k_off; lo_col = j.lo_col_off;
        }
        const unsigned e = (unsigned)(i - cfirst);
        const unsigned r = e / K, k = e - r * K;
        const unsigned r1 = r / R0, r0 = r - r1 * R0;
        const unsigned k1 = k / K0, k0 = k - k1 * K0;
        const float v = __ldg(src + (long)((int)r1 * s_r1 + (int)r0 * s_r0 + (int)k1 * s_k1 + (int)k0 * s_k0));
        if (lo_col < 0) {                                                // fp32 destination (gradient permutation jobs)
            reinterpret_cast<float*>(dst)[(long)r * dst_ld + k_off + k] = v;
            continue;
        }
        const __nv_bfloat16 hi16 = __float2bfloat16(v);
        __nv_bfloat16* d = dst + (long)r * dst_ld + k_off + k;
        *d = hi16;
        if (lo_col != 0) d[lo_col] = __float2bfloat16(v - __bfloat162float(hi16));
    }
}

__global__ void __launch_bounds__(256) permute_f32_kernel(const float* __restrict__ src, int R0, int K0, long rows, int K,
                                                          long s_r1, long s_r0, long s_k1, long s_k0, float alpha,
                                                          float* __restrict__ dst) {
    pdl_wait();
    const long total = rows * K;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const long r = i / K;
        const int k = (int)(i - r * K);
        const long r1 = r / R0, r0 = r - r1 * R0;
        const int k1 = k / K0, k0 = k - k1 * K0;
        dst[i] = alpha * __ldg(src + r1 * s_r1 + r0 * s_r0 + k1 * s_k1 + k0 * s_k0);
    }
}

// ------------------------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long n, float lr, float b1, float b2, float eps,
                                                   float wd, float gscale, const int* __restrict__ step_dev) {
    pdl_wait();
    const int step = *step_dev;
    const float bc1 = 1.f - powf(b1, (float)step);
    const float bc2s = sqrtf(1.f - powf(b2, (float)step));
    const float step_size = lr / bc1;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
        const float pv = p[i];
        const float gv = fmaf(wd, pv, g[i] * gscale);
        const float mv = fmaf(1.f - b1, gv - m[i], m[i]);                 // m + (1-b1)(g - m)
        const float vv = fmaf(1.f - b2, gv * gv - v[i], v[i]);
        m[i] = mv; v[i] = vv;
        p[i] = pv - step_size * (mv / (sqrtf(vv) / bc2s + eps));
    }
}
__global__ void inc_step_kernel(int* step_dev) { if (threadIdx.x == 0 && blockIdx.x == 0) *step_dev += 1; }

static inline int ew_grid(long work_items) {
    long b = (work_items + EW_THREADS - 1) / EW_THREADS;
    if (b < 1) b = 1;
    return (int)(b < EW_BLOCKS ? b : EW_BLOCKS);
}

}  // namespace cb

using namespace cb;

#define BF(p) ((const __nv_bfloat16*)(p))
#define BFW(p) ((__nv_bfloat16*)(p))

static int bn_stats_launch(const void* z, int64_t z_lo_off, const cb_map* map, double* sums, const BnFin& fin, void* stream) {
    MapP m;
    int rc = fill_map(map, m);
    if (rc || !z || !sums) return rc ? rc : CB_ERR_ARG;
    const int rpi = EW_THREADS / (m.c_total >> 3);
    const dim3 grid(ew_grid((m.rows_total + rpi - 1) / rpi * EW_THREADS / 8));
    cudaError_t e = z_lo_off ? launch_pdl(bn_stats_kernel<4, true>, grid, dim3(EW_THREADS), 0, (cudaStream_t)stream, BF(z),
                                          (long)z_lo_off, m, sums, fin)
                             : launch_pdl(bn_stats_kernel<8, false>, grid, dim3(EW_THREADS), 0, (cudaStream_t)stream, BF(z),
                                          (long)z_lo_off, m, sums, fin);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_bn_stats(const void* z, int64_t z_lo_off, const cb_map* map, double* sums, void* stream) {
    BnFin fin{};
    return bn_stats_launch(z, z_lo_off, map, sums, fin, stream);
}

static int bn_fin_args(BnFin& f, const double* sums, int c, double count, float eps, float momentum, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                       float* inv_std) {
    if (c < 1 || !scale || !shift || (gamma && (!sums || !beta || count <= 0))) return CB_ERR_ARG;
    f.c = c; f.count = count; f.eps = eps; f.momentum = momentum; f.gamma = gamma; f.beta = beta;
    f.running_mean = running_mean; f.running_var = running_var; f.scale = scale; f.shift = shift; f.mean_out = mean;
    f.inv_out = inv_std; f.counter = nullptr;
    return CB_OK;
}

extern "C" int cb_bn_finalize(const double* sums, int c, double count, float eps, float momentum, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float* scale, float* shift,
                              float* mean, float* inv_std, void* stream) {
    BnFin f{};
    int rc = bn_fin_args(f, sums, c, count, eps, momentum, gamma, beta, running_mean, running_var, scale, shift, mean, inv_std);
    if (rc) return rc;
    cudaError_t e = launch_pdl(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, (cudaStream_t)stream, sums, f);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_bn_stats_finalize(const void* z, int64_t z_lo_off, const cb_map* map, double* sums, int32_t* counter,
                                    double count, float eps, float momentum, const float* gamma, const float* beta,
                                    float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                                    float* inv_std, void* stream) {
    if (!map || !counter || !gamma) return CB_ERR_ARG;
    BnFin f{};
    int rc = bn_fin_args(f, sums, map->c_mod, count, eps, momentum, gamma, beta, running_mean, running_var, scale, shift, mean,
                         inv_std);
    if (rc) return rc;
    f.counter = counter;
    return bn_stats_launch(z, z_lo_off, map, sums, f, stream);
}

extern "C" int cb_bn_apply(const void* z, int64_t z_lo_off, const float* scale, const float* shift, const void* z_b,
                           int64_t z_b_lo_off, const float* scale_b, const float* shift_b, const void* residual,
                           int32_t res_pitch, int64_t res_lo_off, int relu, const cb_map* map, void* y, int64_t y_lo_off,
                           void* stream) {
    MapP m;
    int rc = fill_map(map, m);
    if (rc) return rc;
    if (!z || !scale || !shift || !y || (z_b && (!scale_b || !shift_b)) || (residual && res_pitch % 8)) return CB_ERR_ARG;
    const int rpi = EW_THREADS / (m.c_total >> 3);
    const dim3 grid(ew_grid((m.rows_total + rpi - 1) / rpi * EW_THREADS / 4));
    const bool lo = z_lo_off || y_lo_off || z_b_lo_off || res_lo_off;
    const bool extra = z_b || residual;
    auto go = [&](auto kern) {
        return launch_pdl(kern, grid, dim3(EW_THREADS), 0, (cudaStream_t)stream, BF(z), (long)z_lo_off, scale, shift, BF(z_b),
                          (long)z_b_lo_off, scale_b, shift_b, BF(residual), (int)res_pitch, (long)res_lo_off, relu, m, BFW(y),
                          (long)y_lo_off);
    };
    cudaError_t e = lo ? (extra ? go(bn_apply_kernel<2, true, true>) : go(bn_apply_kernel<4, true, false>))
                       : (extra ? go(bn_apply_kernel<4, false, true>) : go(bn_apply_kernel<8, false, false>));
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_bn_bwd_reduce(const void* dy, int64_t dy_lo_off, const void* y, int64_t y_lo_off, int relu, const void* z,
                                int64_t z_lo_off, const float* mean, const float* inv_std, const float* mask_scale,
                                const float* mask_shift, const cb_map* map, double* sums, void* stream) {
    MapP m;
    int rc = fill_map(map, m);
    if (rc) return rc;
    const bool zmask = mask_scale && mask_shift && mean;
    if (!zmask) mask_scale = mask_shift = nullptr;
    if (!dy || !sums || (relu && !y && !zmask) || (mean && (!z || !inv_std))) return CB_ERR_ARG;
    const int rpi = EW_THREADS / (m.c_total >> 3);
    const dim3 grid(ew_grid((m.rows_total + rpi - 1) / rpi * EW_THREADS / 4));
    const bool lo = dy_lo_off || y_lo_off || z_lo_off;
    const bool ymask = relu && !zmask;
    auto go = [&](auto kern) {
        return launch_pdl(kern, grid, dim3(EW_THREADS), 0, (cudaStream_t)stream, BF(dy), (long)dy_lo_off, BF(y), (long)y_lo_off,
                          relu, BF(z), (long)z_lo_off, mean, inv_std, mask_scale, mask_shift, m, sums);
    };
    cudaError_t e = lo ? (ymask ? go(bn_bwd_reduce_kernel<2, true, true>) : go(bn_bwd_reduce_kernel<4, true, false>))
                       : (ymask ? go(bn_bwd_reduce_kernel<4, false, true>) : go(bn_bwd_reduce_kernel<6, false, false>));
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_bn_bwd_apply(const void* dy, int64_t dy_lo_off, const void* y, int64_t y_lo_off, int relu, const void* z,
                               int64_t z_lo_off, const float* mean, const float* inv_std, const float* gamma,
                               const float* mask_scale, const float* mask_shift,
                               const double* sums, double count, const cb_map* map, void* dz, int64_t dz_lo_off,
                               void* dsum_pf, int64_t dsum_lo_off, float* d_gamma, float* d_beta, void* stream) {
    MapP m;
    int rc = fill_map(map, m);
    if (rc) return rc;
    const bool zmask = mask_scale && mask_shift && mean;
    if (!zmask) mask_scale = mask_shift = nullptr;
    if (!dy || !dz || !sums || (relu && !y && !zmask) || (mean && (!z || !inv_std || !gamma || count <= 0))) return CB_ERR_ARG;
    const int rpi = EW_THREADS / (m.c_total >> 3);
    const dim3 grid(ew_grid((m.rows_total + rpi - 1) / rpi * EW_THREADS / 4));
    const bool lo = dy_lo_off || y_lo_off || z_lo_off || dz_lo_off || dsum_lo_off;
    const bool ymask = relu && !zmask;
    auto go = [&](auto kern) {
        return launch_pdl(kern, grid, dim3(EW_THREADS), 0, (cudaStream_t)stream, BF(dy), (long)dy_lo_off, BF(y), (long)y_lo_off,
                          relu, BF(z), (long)z_lo_off, mean, inv_std, gamma, mask_scale, mask_shift, sums, count, m, BFW(dz),
                          (long)dz_lo_off, BFW(dsum_pf), (long)dsum_lo_off, d_gamma, d_beta);
    };
    cudaError_t e = lo ? (ymask ? go(bn_bwd_apply_kernel<2, true, true>) : go(bn_bwd_apply_kernel<3, true, false>))
                       : (ymask ? go(bn_bwd_apply_kernel<4, false, true>) : go(bn_bwd_apply_kernel<6, false, false>));
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_heads_grad_pack(const float* const* grads, const int32_t* head_cn, int n_heads, int n, int H, int W,
                                  void* g_pf, int64_t g_lo_off, float* d_bias, void* stream) {
    if (!grads || !head_cn || n_heads < 1 || n_heads > CB_MAX_HEADS || !g_pf || !d_bias || n < 1) return CB_ERR_ARG;
    HeadsP hp;
    int c0 = 0;
    for (int s = 0; s < CB_MAX_HEADS; ++s) {
        hp.g[s] = s < n_heads ? grads[s] : nullptr;
        hp.cn[s] = s < n_heads ? head_cn[s] : 0;
        hp.c0[s] = c0;
        if (s < n_heads) { if (!grads[s] || head_cn[s] < 1) return CB_ERR_ARG; c0 += head_cn[s]; }
    }
    if (c0 > 32) return CB_ERR_ARG;
    hp.n_heads = n_heads; hp.total = c0;
    const long rows = (long)n * (H + 2) * (W + 2);
    cudaError_t e = launch_pdl(heads_grad_pack_kernel, dim3(ew_grid(rows)), dim3(256), 0, (cudaStream_t)stream, hp, n, H, W,
                               BFW(g_pf), (long)g_lo_off, d_bias);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_warp_att_fuse_bwd(const void* feat, int in_ps, int64_t in_lo_off, int sum_agents, const double* affine,
                                    const int32_t* agent_off, int n_scenes, int max_cav, int H, int W, int C, int method,
                                    const void* d_fused_pf, int64_t d_fused_lo_off, float* d_feat, void* stream) {
    if (!feat || !affine || !agent_off || !d_fused_pf || !d_feat || n_scenes < 1 || H < 1 || W < 1) return CB_ERR_ARG;
    if (C != 64 && C != 128 && C != 256) return CB_ERR_ARG;
    if (method != 0 && method != 1) return CB_ERR_ARG;
    FuseBG g;
    g.H = H; g.W = W; g.C = C; g.in_ps = in_ps;
    g.Hp = in_ps ? (H + 1) / 2 + 2 : H + 2;
    g.Wp = in_ps ? (W + 1) / 2 + 2 : W + 2;
    g.plane_rows = (long)sum_agents * g.Hp * g.Wp;
    g.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    const long total = (long)n_scenes * H * W;
    const int lpp = C / 8, ppw = 32 / lpp;
    long blocks = (total + 8L * ppw - 1) / (8L * ppw);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    if (blocks > 148 * 2) blocks = 148 * 2;                              // one resident wave (2 CTAs / SM)
#define CB_FB_LAUNCH(LPP_, MAXN_)                                                                                          \
    e = launch_pdl(warp_att_fuse_bwd_kernel<LPP_, MAXN_>, dim3((unsigned)blocks), dim3(256), 0, st, BF(feat), (long)in_lo_off, \
                   affine, agent_off, n_scenes, max_cav, g, method, BF(d_fused_pf), (long)d_fused_lo_off, d_feat)
    const bool small = max_cav <= 5;
    if (lpp == 8) { if (small) CB_FB_LAUNCH(8, 5); else CB_FB_LAUNCH(8, FB_MAX_AGENTS); }
    else if (lpp == 16) { if (small) CB_FB_LAUNCH(16, 5); else CB_FB_LAUNCH(16, FB_MAX_AGENTS); }
    else { if (small) CB_FB_LAUNCH(32, 5); else CB_FB_LAUNCH(32, FB_MAX_AGENTS); }
#undef CB_FB_LAUNCH
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_grad_combine(const float* acc, const void* addend, int64_t addend_lo_off, int to_ps, int n_cap, int n, int H,
                               int W, int C, void* out, int64_t out_lo_off, void* stream) {
    if (!acc || !out || n < 1 || n > n_cap || C % 8 || ((uintptr_t)acc & 15)) return CB_ERR_ARG;
    const long total = (long)n * H * W * (C / 8);
    cudaError_t e = launch_pdl(grad_combine_kernel, dim3(ew_grid(total / 2)), dim3(256), 0, (cudaStream_t)stream, acc, BF(addend),
                               (long)addend_lo_off, to_ps, n_cap, n, H, W, C, BFW(out), (long)out_lo_off);
    return e == cudaSuccess ? CB_OK : (int)e;
}

static PfnT make_pfnt(const float* vsize, const float* center_off) {
    PfnT t;
    t.vx = vsize[0]; t.vy = vsize[1]; t.vz = vsize[2];
    t.offx = center_off[0]; t.offy = center_off[1]; t.offz = center_off[2];
    return t;
}

extern "C" int cb_pfn_train_stats(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows_cap,
                                  const int32_t* n_voxels_dev, int max_pts, const float* vsize, const float* center_off,
                                  double* sums, void* stream) {
    if (!voxels || !coords || !num_points || !sums || n_rows_cap < 1 || max_pts < 1 || max_pts > 32) return CB_ERR_ARG;
    long blocks = ((long)n_rows_cap + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    cudaError_t e = launch_pdl(pfn_train_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                               (const float4*)voxels, (const int*)coords, (const int*)num_points, n_rows_cap,
                               (const int*)n_voxels_dev, max_pts, make_pfnt(vsize, center_off), sums);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_pfn_train_finalize(const double* sums, const int32_t* n_voxels_dev, int n_rows_cap, int max_pts,
                                     const float* w, const float* gamma, const float* beta, float eps, float momentum,
                                     float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                                     float* inv_std, void* stream) {
    if (!sums || !w || !gamma || !beta || !scale || !shift) return CB_ERR_ARG;
    cudaError_t e = launch_pdl(pfn_train_finalize_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, sums,
                               (const int*)n_voxels_dev, n_rows_cap, max_pts, w, gamma, beta, eps, momentum, running_mean,
                               running_var, scale, shift, mean, inv_std);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_pfn_bwd(const float* voxels, const int32_t* coords, const int32_t* num_points, int n_rows_cap,
                          const int32_t* n_voxels_dev, int max_pts, const float* w, const float* scale, const float* shift,
                          const float* mean, const float* inv_std, const float* vsize, const float* center_off,
                          const void* d_canvas_ps, int64_t d_canvas_lo_off, int canvas_agents, int ny, int nx, double* bsum,
                          void* stream) {
    if (!voxels || !coords || !num_points || !w || !scale || !shift || !mean || !inv_std || !d_canvas_ps || !bsum)
        return CB_ERR_ARG;
    if (n_rows_cap < 1 || max_pts < 1 || max_pts > 32) return CB_ERR_ARG;
    CanvasG cg;
    cg.ny = ny; cg.nx = nx; cg.Hq = (ny + 1) / 2 + 2; cg.Wq = (nx + 1) / 2 + 2;
    cg.plane_rows = (long)canvas_agents * cg.Hq * cg.Wq;
    long blocks = ((long)n_rows_cap + 7) / 8;
    if (blocks > 148 * 4) blocks = 148 * 4;
    cudaError_t e = launch_pdl(pfn_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const float4*)voxels,
                               (const int*)coords, (const int*)num_points, n_rows_cap, (const int*)n_voxels_dev, max_pts, w,
                               scale, shift, mean, inv_std, make_pfnt(vsize, center_off), BF(d_canvas_ps), (long)d_canvas_lo_off,
                               cg, bsum);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_pfn_bwd_finalize(const double* stats_sums, const double* bsum, const int32_t* n_voxels_dev, int n_rows_cap,
                                   int max_pts, const float* w, const float* gamma, const float* mean, const float* inv_std,
                                   float* d_w, float* d_gamma, float* d_beta, void* stream) {
    if (!stats_sums || !bsum || !w || !gamma || !mean || !inv_std || !d_w || !d_gamma || !d_beta) return CB_ERR_ARG;
    cudaError_t e = launch_pdl(pfn_bwd_finalize_kernel, dim3(5), dim3(128), 0, (cudaStream_t)stream, stats_sums, bsum,
                               (const int*)n_voxels_dev, n_rows_cap, max_pts, w, gamma, mean, inv_std, d_w, d_gamma, d_beta);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_pack_weight(const float* src, int R1, int R0, int K1, int K0, int64_t s_r1, int64_t s_r0, int64_t s_k1,
                              int64_t s_k0, void* dst, int dst_ld, int k_off, int lo_col_off, void* stream) {
    if (!src || !dst || R1 < 1 || R0 < 1 || K1 < 1 || K0 < 1 || dst_ld < K1 * K0 + k_off) return CB_ERR_ARG;
    const long rows = (long)R1 * R0;
    const int K = K1 * K0;
    cudaError_t e = launch_pdl(pack_weight_kernel, dim3(ew_grid(rows * K)), dim3(256), 0, (cudaStream_t)stream, src, R0, K0, rows,
                               K, (long)s_r1, (long)s_r0, (long)s_k1, (long)s_k0, BFW(dst), dst_ld, k_off, lo_col_off);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_pack_weights_batch(const cb_pack_job* jobs_dev, int n_jobs, int64_t total_elems, void* stream) {
    if (!jobs_dev || n_jobs < 1 || total_elems < 1) return CB_ERR_ARG;
    cudaError_t e = launch_pdl(pack_weights_batch_kernel, dim3(ew_grid(total_elems / 4)), dim3(256), 0, (cudaStream_t)stream,
                               jobs_dev, n_jobs, (long)total_elems);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_permute_f32(const float* src, int R1, int R0, int K1, int K0, int64_t s_r1, int64_t s_r0, int64_t s_k1,
                              int64_t s_k0, float alpha, float* dst, void* stream) {
    if (!src || !dst || R1 < 1 || R0 < 1 || K1 < 1 || K0 < 1) return CB_ERR_ARG;
    const long rows = (long)R1 * R0;
    const int K = K1 * K0;
    cudaError_t e = launch_pdl(permute_f32_kernel, dim3(ew_grid(rows * K)), dim3(256), 0, (cudaStream_t)stream, src, R0, K0, rows,
                               K, (long)s_r1, (long)s_r0, (long)s_k1, (long)s_k0, alpha, dst);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float grad_scale, int32_t* step_dev, int inc_step, void* stream) {
    if (!p || !g || !m || !v || !step_dev || n < 1) return CB_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (inc_step) {
        inc_step_kernel<<<1, 32, 0, st>>>(step_dev);
        CB_CHECK_LAUNCH();
    }
    cudaError_t e = launch_pdl(adam_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, st, p, g, m, v, (long)n, lr, beta1, beta2, eps,
                               weight_decay, grad_scale, (const int*)step_dev);
    return e == cudaSuccess ? CB_OK : (int)e;
}
