// Weight gradient of the convolutions on Blackwell tensor cores (sm_100a).
//
//   dW[m][box j][c] += sum_q dZ[q + a_off][m0 + m] * X_j[q + b_off_j][col_j + c]        q = flattened padded pixel (PF rows)
//
// i.e. the adjoint of the forward's row-shifted implicit GEMM (conv_tc.cu) with respect to the weights: for every filter
// tap one GEMM whose REDUCTION dimension is the pixel index.  Both operands are therefore "MN-major" for the tensor core
// (channels contiguous in memory, pixels along K): the [64 pixels][64 channels] TMA boxes (SWIZZLE_128B) are consumed
// directly through MN-major UMMA shared-memory descriptors - no transposition pass, no im2col.
//   M side: 128 output channels of dZ (two 64-channel boxes; columns past the tensor's pitch are zero-filled by TMA)
//   N side: up to four 64-channel boxes of X, each with its own row shift / tensor / destination, so that several filter
//           taps (Cin = 64) or several channel blocks (Cin >= 128) share one 128 x 256 accumulator
//   K side: pixels, split over CTAs (split-K); partial sums are added to the fp32 gradient with 16-byte vector reductions
// Persistent, warp specialised like the forward kernel: warp 8 TMA producer, warp 9 MMA issuer, warp 10 TMEM allocator,
// warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4); double-buffered accumulator.
// Autograd of /root/reference/opencood/models/sub_modules/resblock.py:53-69, base_bev_backbone_resnet.py:52-65,
// downsample_conv.py:18-24, point_pillar_baseline_multiscale.py:126-133 w.r.t. the conv weights (torch.nn.grad.conv2d_weight).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

constexpr int WG_KB = 64;                       // pixels per pipeline stage (4 UMMA K-steps of 16)
constexpr int WG_BOX_BYTES = WG_KB * 64 * 2;    // one [64 pixels][64 channels] bf16 box
constexpr int WG_A_BYTES = 2 * WG_BOX_BYTES;    // M = 128
constexpr int WG_B_BYTES = CB_WGRAD_MAX_BOXES * WG_BOX_BYTES;
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_STAGES = 4;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024;
constexpr int WG_TMEM_COLS = 512;               // 2 x 256 fp32 columns
constexpr int WG_THREADS = 384;
constexpr int WGW_PRODUCER = 8, WGW_MMA = 9, WGW_ALLOC = 10;

struct WgParams {
    long rows_total;
    int n_kblocks;          // ceil(rows_total / 64)
    int k_splits;
    int n_units;
    int n_combos;           // 1, or 3 in precise mode: (hi,hi) (lo,hi) (hi,lo)
    int x_lo_rows[2];       // row offset of the lo plane inside the X tensor maps (precise)
    float* dw;
    cb_wgrad_unit units[CB_WGRAD_MAX_UNITS];
};

// MN-major, SWIZZLE_128B operand descriptor: 64-element (128 B) channel groups `lbo` bytes apart, 8-pixel K groups 1024 B apart
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16)
__device__ __forceinline__ uint32_t make_idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dz, const __grid_constant__ CUtensorMap tmap_dz_lo,
                const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_x1,
                const __grid_constant__ WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[WG_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[WG_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    const uint32_t smem0 = raw_addr + pad;                     // 1 KiB aligned (SWIZZLE_128B atoms)
    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int total_items = p.n_units * p.k_splits;

    if (warp == WGW_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_dz); prefetch_tmap(&tmap_dz_lo); prefetch_tmap(&tmap_x0); prefetch_tmap(&tmap_x1);
    }
    if (warp == WGW_MMA && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == WGW_ALLOC) {
        tmem_alloc(&tmem_base_smem, WG_TMEM_COLS);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);

    // item = (K split, unit), split-major: CTAs that run at the same time work on the same pixel range (L2 reuse)
    // Producer and MMA issuer: the whole warp runs the loop, only the TMA / tcgen05 instructions sit under elect, so the
    // loop state stays in uniform registers (see the MMA-issuer note in conv_gemm_halo64_kernel).
    if (warp == WGW_PRODUCER) {
        {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                const int split = item / p.n_units, ui = item - split * p.n_units;
                const cb_wgrad_unit& u = p.units[ui];
                const int kb0 = (int)((long)p.n_kblocks * split / p.k_splits);
                const int kb1 = (int)((long)p.n_kblocks * (split + 1) / p.k_splits);
                const uint32_t bytes = (uint32_t)(WG_A_BYTES + u.n_boxes * WG_BOX_BYTES);
                for (int kb = kb0; kb < kb1; ++kb) {
                    for (int c = 0; c < p.n_combos; ++c) {
                        const uint32_t fb = full0 + stage * 8;
                        mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx_a(fb, bytes);
                            const uint32_t sa = smem0 + stage * WG_STAGE_BYTES;
                            const int q0 = kb * WG_KB;
                            const CUtensorMap* ta = (c == 1) ? &tmap_dz_lo : &tmap_dz;
                            tma_load_2d_a(sa, ta, fb, u.m0, q0 + u.a_row_off);
                            tma_load_2d_a(sa + WG_BOX_BYTES, ta, fb, u.m0 + 64, q0 + u.a_row_off);
                            for (int j = 0; j < u.n_boxes; ++j) {
                                const int sel = u.box[j].x_sel;
                                const int ro = u.box[j].row_off + (c == 2 ? p.x_lo_rows[sel] : 0);
                                tma_load_2d_a(sa + WG_A_BYTES + j * WG_BOX_BYTES, sel ? &tmap_x1 : &tmap_x0, fb,
                                              (int)u.box[j].col, q0 + ro);
                            }
                        }
                        __syncwarp();
                        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == WGW_MMA) {
        {
            const uint64_t desc0 = make_mn_sw128_desc(smem0, WG_BOX_BYTES);
            const uint32_t a_lo0 = (uint32_t)desc0, hi = (uint32_t)(desc0 >> 32);
            const uint32_t b_lo0 = (uint32_t)make_mn_sw128_desc(smem0 + WG_A_BYTES, WG_BOX_BYTES);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
                const int split = item / p.n_units, ui = item - split * p.n_units;
                const int nb = p.units[ui].n_boxes;
                const int kb0 = (int)((long)p.n_kblocks * split / p.k_splits);
                const int kb1 = (int)((long)p.n_kblocks * (split + 1) / p.k_splits);
                const uint32_t idesc = make_idesc_bf16_mn(128, 64 * nb);
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * 256;
                const int n_steps = (kb1 - kb0) * p.n_combos;
                for (int s = 0; s < n_steps; ++s) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (WG_STAGE_BYTES >> 4));
                    const uint32_t b_lo = b_lo0 + (uint32_t)(stage * (WG_STAGE_BYTES >> 4));
                    mbar_wait_a(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < WG_KB / 16; ++k) {
                            // 16 pixels further along K = 16 rows x 128 B = 2048 B (start-address field counts 16-byte units)
                            umma_bf16_lh(d_tmem, a_lo + (uint32_t)(k * 128), hi, b_lo + (uint32_t)(k * 128), hi, idesc,
                                         (s > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit_a(empty0 + stage * 8);
                        if (s + 1 == n_steps) umma_commit_a(tfull0 + buf * 8);
                    }
                    __syncwarp();
                    if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
                }
                if (n_steps == 0 && elect_one()) umma_commit_a(tfull0 + buf * 8);   // empty K range (epilogue skips it)
                __syncwarp();
            }
        }
    } else if (warp < 8) {
        const int q4 = warp & 3, half = warp >> 2;
        const int m = q4 * 32 + lane;                                   // TMEM lane = dZ channel within the unit
        int it = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            const int split = item / p.n_units, ui = item - split * p.n_units;
            const cb_wgrad_unit& u = p.units[ui];
            const int kb0 = (int)((long)p.n_kblocks * split / p.k_splits);
            const int kb1 = (int)((long)p.n_kblocks * (split + 1) / p.k_splits);
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait_a(tfull0 + buf * 8, acc_phase);
            tc_fence_after();
            if (kb1 > kb0) {
                const uint32_t t_row = tmem_base + buf * 256 + ((uint32_t)(q4 * 32) << 16);
                // 32-column chunks: this warp takes the chunks with (chunk & 1) == half
                for (int ch = half; ch < 2 * u.n_boxes; ch += 2) {
                    uint32_t r[32];
                    tmem_ld32(t_row + ch * 32, r);
                    tmem_ld_wait();
                    const int j = ch >> 1;
                    if (m < u.m_valid) {
                        float* dst = p.dw + u.box[j].out_off + (long)m * u.box[j].out_ld + (ch & 1) * 32;
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            red_add_v4(dst + 4 * t, __uint_as_float(r[4 * t]), __uint_as_float(r[4 * t + 1]),
                                       __uint_as_float(r[4 * t + 2]), __uint_as_float(r[4 * t + 3]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WGW_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, WG_TMEM_COLS);
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 wg_get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
    });
    return fn;
}

// bf16 [rows][pitch]; box = 64 channels x 64 rows, 128-byte swizzle, zero fill outside [0, rows) x [0, cols)
static int wg_make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems) {
    auto enc = wg_get_encode();
    if (!enc) return CB_ERR_DRIVER;
    if (((uintptr_t)base & 15) || (pitch_elems * 2) % 16 || rows < 1) return CB_ERR_ARG;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)WG_KB};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CB_OK : CB_ERR_DRIVER;
}

}  // namespace cb

extern "C" int cb_wgrad(const cb_wgrad_desc* d, int max_ctas, void* stream) {
    using namespace cb;
    if (!d || !d->dz_ptr || !d->x_ptr[0] || !d->dw) return CB_ERR_ARG;
    if (d->n_units < 1 || d->n_units > CB_WGRAD_MAX_UNITS || d->rows_total < 1) return CB_ERR_ARG;
    if (((uintptr_t)d->dw & 15)) return CB_ERR_ARG;
    static thread_local WgParams p;
    p.rows_total = d->rows_total;
    p.n_kblocks = (int)((d->rows_total + WG_KB - 1) / WG_KB);
    p.n_units = d->n_units;
    p.n_combos = d->dz_lo_ptr ? 3 : 1;
    p.x_lo_rows[0] = d->x_lo_rows[0]; p.x_lo_rows[1] = d->x_lo_rows[1];
    p.dw = d->dw;
    for (int i = 0; i < d->n_units; ++i) {
        const cb_wgrad_unit& u = d->units[i];
        if (u.n_boxes < 1 || u.n_boxes > CB_WGRAD_MAX_BOXES || u.m_valid < 1 || u.m_valid > 128 || u.m0 % 8) return CB_ERR_ARG;
        for (int j = 0; j < u.n_boxes; ++j) {
            if (u.box[j].x_sel > 1 || !d->x_ptr[u.box[j].x_sel] || u.box[j].col % 8 || (u.box[j].out_off & 3) ||
                (u.box[j].out_ld & 3))
                return CB_ERR_ARG;
        }
        p.units[i] = u;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cap = max_ctas > 0 ? max_ctas : sms;
    int ks = d->k_splits;
    if (ks <= 0) {                                   // exactly one wave of CTAs (units * splits <= SMs: a second, partial
        ks = cap / d->n_units;                       // wave would double the kernel's duration), >= 8 K blocks per work item
        if (ks < 1) ks = 1;
        const int lim = p.n_kblocks / 8 > 0 ? p.n_kblocks / 8 : 1;
        if (ks > lim) ks = lim;
    }
    if (ks > p.n_kblocks) ks = p.n_kblocks;
    p.k_splits = ks;
    CUtensorMap tdz, tdzl, tx0, tx1;
    int rc = wg_make_tmap(&tdz, d->dz_ptr, d->rows_total, d->dz_pitch, d->dz_pitch);
    if (rc) return rc;
    tdzl = tdz;
    if (d->dz_lo_ptr) {
        rc = wg_make_tmap(&tdzl, d->dz_lo_ptr, d->rows_total, d->dz_pitch, d->dz_pitch);
        if (rc) return rc;
    }
    rc = wg_make_tmap(&tx0, d->x_ptr[0], d->x_rows[0], d->x_pitch[0], d->x_pitch[0]);
    if (rc) return rc;
    tx1 = tx0;
    if (d->x_ptr[1]) {
        rc = wg_make_tmap(&tx1, d->x_ptr[1], d->x_rows[1], d->x_pitch[1], d->x_pitch[1]);
        if (rc) return rc;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int grid = p.n_units * p.k_splits;
    if (grid > cap) grid = cap;
    cudaError_t le = launch_pdl(wgrad_tc_kernel, dim3(grid), dim3(WG_THREADS), WG_SMEM_BYTES, (cudaStream_t)stream, tdz, tdzl,
                                tx0, tx1, p);
    return le == cudaSuccess ? CB_OK : (int)le;
}

// ------------------------------------------------------------------------------------------------------------------
// SIMT evaluation of the same descriptor (validation of the tensor-core kernel; tests only): one thread per
// (unit, m, box, channel), fp32 accumulation over all pixels.
// ------------------------------------------------------------------------------------------------------------------
namespace cb {
__global__ void wgrad_simt_kernel(const __nv_bfloat16* __restrict__ dz, const __nv_bfloat16* __restrict__ dz_lo, int dz_pitch,
                                  const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1, long x_rows0,
                                  long x_rows1, int x_pitch0, int x_pitch1, const __grid_constant__ WgParams p) {
    const int ui = blockIdx.y;
    const cb_wgrad_unit& u = p.units[ui];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per_m = u.n_boxes * 64;
    if (t >= u.m_valid * per_m) return;
    const int m = t / per_m, jc = t - m * per_m, j = jc >> 6, c = jc & 63;
    const int sel = u.box[j].x_sel;
    const __nv_bfloat16* x = sel ? x1 : x0;
    const long xr = sel ? x_rows1 : x_rows0;
    const int xp = sel ? x_pitch1 : x_pitch0;
    if (u.m0 + m >= dz_pitch) return;
    float acc = 0.f;
    for (long q = 0; q < p.rows_total; ++q) {
        const long qa = q + u.a_row_off, qb = q + u.box[j].row_off;
        if (qa < 0 || qa >= p.rows_total || qb < 0 || qb >= xr) continue;
        float a = __bfloat162float(dz[qa * dz_pitch + u.m0 + m]);
        float b = __bfloat162float(x[qb * xp + u.box[j].col + c]);
        if (p.n_combos == 3) {
            const float al = __bfloat162float(dz_lo[qa * dz_pitch + u.m0 + m]);
            const long qbl = qb + p.x_lo_rows[sel];
            const float bl = qbl < xr ? __bfloat162float(x[qbl * xp + u.box[j].col + c]) : 0.f;
            acc += a * b + al * b + a * bl;
        } else {
            acc += a * b;
        }
    }
    atomicAdd(p.dw + u.box[j].out_off + (long)m * u.box[j].out_ld + c, acc);
}
}  // namespace cb

extern "C" int cb_wgrad_simt(const cb_wgrad_desc* d, void* stream) {
    using namespace cb;
    if (!d || !d->dz_ptr || !d->x_ptr[0] || !d->dw || d->n_units < 1 || d->n_units > CB_WGRAD_MAX_UNITS) return CB_ERR_ARG;
    static thread_local WgParams p;
    p.rows_total = d->rows_total;
    p.n_kblocks = 0; p.k_splits = 1; p.n_units = d->n_units;
    p.n_combos = d->dz_lo_ptr ? 3 : 1;
    p.x_lo_rows[0] = d->x_lo_rows[0]; p.x_lo_rows[1] = d->x_lo_rows[1];
    p.dw = d->dw;
    for (int i = 0; i < d->n_units; ++i) p.units[i] = d->units[i];
    dim3 grid((128 * CB_WGRAD_MAX_BOXES * 64 + 255) / 256, d->n_units);
    wgrad_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)d->dz_ptr, (const __nv_bfloat16*)d->dz_lo_ptr, d->dz_pitch, (const __nv_bfloat16*)d->x_ptr[0],
        (const __nv_bfloat16*)d->x_ptr[1], d->x_rows[0], d->x_rows[1], d->x_pitch[0], d->x_pitch[1], p);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
