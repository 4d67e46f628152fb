// Device-side parameter block and epilogue shared by the tcgen05 conv GEMM (conv_tc.cu) and the SIMT
// validation kernel (conv_simt.cu).  See include/coalign_b200.h for the GEMM definition.
#pragma once
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

struct ConvParams {
    // row space
    int n_img, Hp, Wp, pad;
    long rows_total;
    int n_total, cout_mod, relu;
    // epilogue
    const float* bias;
    const __nv_bfloat16* residual;
    int res_pitch;
    long res_lo_off;
    __nv_bfloat16* out;
    int out_pitch, out_ch_off;
    long out_lo_off;
    int out_mode, up_k, out_Hp, out_Wp;
    long out_plane_rows;
    float* head_out[CB_MAX_HEADS];
    int head_c0[CB_MAX_HEADS], head_cn[CB_MAX_HEADS], n_heads;
    int n_ksteps;
    cb_kstep ksteps[CB_MAX_KSTEPS];
};

struct RowDest {
    long row;   // destination row (PF/PS/upsample) or -1 when the GEMM row is halo / out of range
    int n, h, w;
};

// Decode GEMM row q (flattened padded pixel, < 2^31) and compute where its result goes.
__device__ __forceinline__ RowDest decode_row(const ConvParams& p, long q64, int n0) {
    RowDest d;
    d.row = -1;
    d.n = d.h = d.w = 0;
    if (q64 >= p.rows_total) return d;
    const unsigned q = (unsigned)q64;
    const unsigned plane = (unsigned)(p.Hp * p.Wp);
    const unsigned n = q / plane;
    const unsigned rem = q - n * plane;
    const unsigned hp = rem / (unsigned)p.Wp;
    const unsigned wp = rem - hp * (unsigned)p.Wp;
    const unsigned pad = (unsigned)p.pad;
    if (hp < pad || hp > (unsigned)(p.Hp - 1) - pad || wp < pad || wp > (unsigned)(p.Wp - 1) - pad) return d;
    const int h = (int)(hp - pad), w = (int)(wp - pad);
    d.n = (int)n; d.h = h; d.w = w;
    if (p.out_mode == CB_OUT_HEADS || (p.out_mode == CB_OUT_PF && p.out_Hp == p.Hp && p.out_Wp == p.Wp)) {
        d.row = q64;
    } else if (p.out_mode == CB_OUT_PF) {      // destination PF map with its own (1-pixel) halo: address by pixel
        d.row = (long)((int)n * p.out_Hp + h + 1) * p.out_Wp + w + 1;
    } else if (p.out_mode == CB_OUT_PS) {
        const int ph = (h & 1) * 2 + (w & 1);
        d.row = (long)ph * p.out_plane_rows + (long)((int)n * p.out_Hp + (h >> 1) + 1) * p.out_Wp + (w >> 1) + 1;
    } else {  // CB_OUT_UPSAMPLE: GEMM column block n0 selects the (a,b) sub-pixel
        const int ab = n0 / p.cout_mod;
        const int a = ab / p.up_k, b = ab - a * p.up_k;
        d.row = (long)((int)n * p.out_Hp + p.up_k * h + a + 1) * p.out_Wp + (p.up_k * w + b + 1);
    }
    return d;
}

// Epilogue for 32 consecutive GEMM columns [col0, col0+32) of one row.  v = raw fp32 accumulators.
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const RowDest& d, long q, int col0,
                                               float (&v)[32]) {
    if (d.row < 0) return;
    const int c_base = col0 % p.cout_mod;   // chunk never straddles cout_mod (both multiples of 32)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + c_base + j);
    if (p.residual != nullptr) {
        const uint4* r = reinterpret_cast<const uint4*>(p.residual + q * (long)p.res_pitch + col0);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint4 u = __ldg(r + t);
            v[8 * t + 0] += bf16_lo(u.x); v[8 * t + 1] += bf16_hi(u.x);
            v[8 * t + 2] += bf16_lo(u.y); v[8 * t + 3] += bf16_hi(u.y);
            v[8 * t + 4] += bf16_lo(u.z); v[8 * t + 5] += bf16_hi(u.z);
            v[8 * t + 6] += bf16_lo(u.w); v[8 * t + 7] += bf16_hi(u.w);
        }
        if (p.res_lo_off != 0) {
            const uint4* rl = reinterpret_cast<const uint4*>(p.residual + p.res_lo_off + q * (long)p.res_pitch + col0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint4 u = __ldg(rl + t);
                v[8 * t + 0] += bf16_lo(u.x); v[8 * t + 1] += bf16_hi(u.x);
                v[8 * t + 2] += bf16_lo(u.y); v[8 * t + 3] += bf16_hi(u.y);
                v[8 * t + 4] += bf16_lo(u.z); v[8 * t + 5] += bf16_hi(u.z);
                v[8 * t + 6] += bf16_lo(u.w); v[8 * t + 7] += bf16_hi(u.w);
            }
        }
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
    }
    if (p.out_mode == CB_OUT_HEADS) {
        const int H = p.Hp - 2, W = p.Wp - 2;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = col0 + j;
#pragma unroll
            for (int s = 0; s < CB_MAX_HEADS; ++s) {
                if (s < p.n_heads && c >= p.head_c0[s] && c < p.head_c0[s] + p.head_cn[s]) {
                    p.head_out[s][(((long)d.n * p.head_cn[s] + (c - p.head_c0[s])) * H + d.h) * W + d.w] = v[j];
                }
            }
        }
        return;
    }
    __nv_bfloat16* o = p.out + d.row * (long)p.out_pitch + p.out_ch_off + c_base;
    uint4* o4 = reinterpret_cast<uint4*>(o);
    if (p.out_lo_off == 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint4 u;
            u.x = pack_bf16(v[8 * t + 0], v[8 * t + 1]);
            u.y = pack_bf16(v[8 * t + 2], v[8 * t + 3]);
            u.z = pack_bf16(v[8 * t + 4], v[8 * t + 5]);
            u.w = pack_bf16(v[8 * t + 6], v[8 * t + 7]);
            o4[t] = u;
        }
    } else {
        uint4* l4 = reinterpret_cast<uint4*>(o + p.out_lo_off);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = v[8 * t + 2 * e], b = v[8 * t + 2 * e + 1];
                hi[e] = pack_bf16(a, b);
                lo[e] = pack_bf16(a - bf16_lo(hi[e]), b - bf16_hi(hi[e]));
            }
            o4[t] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            l4[t] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// bf16x2 pack with the ReLU folded into the conversion (low half = a, high half = b)
__device__ __forceinline__ uint32_t pack_bf16_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %2, %1;" : "=r"(d) : "f"(a), "f"(b));
    return d;
}

// Fast epilogue of the tensor-core kernels (bf16 outputs, not the heads): bias from shared memory, residual
// already prefetched into registers, ReLU fused into the bf16 conversion.  The 32 rows x 64 B of the warp's chunk are
// transposed through a 2 KB per-warp shared-memory stage so that every store instruction writes 8 rows x 64 B of full
// sectors instead of 32 rows x 16 B (the scattered 16-byte stores cost 15-30 us per layer, profiles/r1_conv_analysis.md).
__device__ __forceinline__ void epilogue_chunk_fast(const ConvParams& p, int drow, int c_base,
                                                    const float* __restrict__ s_bias, const uint4 (&res)[4],
                                                    bool has_res, float (&v)[32], uint4* __restrict__ stage, int lane,
                                                    int dbg = 0) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias + c_base);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float4 b = b4[t];
        v[4 * t + 0] += b.x; v[4 * t + 1] += b.y; v[4 * t + 2] += b.z; v[4 * t + 3] += b.w;
    }
    if (has_res) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[8 * t + 0] += bf16_lo(res[t].x); v[8 * t + 1] += bf16_hi(res[t].x);
            v[8 * t + 2] += bf16_lo(res[t].y); v[8 * t + 3] += bf16_hi(res[t].y);
            v[8 * t + 4] += bf16_lo(res[t].z); v[8 * t + 5] += bf16_hi(res[t].z);
            v[8 * t + 6] += bf16_lo(res[t].w); v[8 * t + 7] += bf16_hi(res[t].w);
        }
    }
    if (dbg & 8) {
        // direct mode: the thread's 32 columns (64 B of one output row) leave as two 256-bit stores - full 32-byte
        // sectors without the shared-memory transposition (no smem traffic, no warp syncs)
        if (drow >= 0) {
            uint32_t w[16];
#pragma unroll
            for (int t = 0; t < 16; ++t)
                w[t] = p.relu ? pack_bf16_relu(v[2 * t], v[2 * t + 1]) : pack_bf16(v[2 * t], v[2 * t + 1]);
            __nv_bfloat16* o = p.out + (long)drow * p.out_pitch + p.out_ch_off + c_base;
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                         "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 16), "r"(w[8]), "r"(w[9]), "r"(w[10]),
                         "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
        }
        return;
    }
    const int sw = (lane >> 1) & 3;                         // 16-byte chunk swizzle: conflict-free writes and reads
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint4 u;
        if (p.relu) u = make_uint4(pack_bf16_relu(v[8 * t + 0], v[8 * t + 1]), pack_bf16_relu(v[8 * t + 2], v[8 * t + 3]),
                                   pack_bf16_relu(v[8 * t + 4], v[8 * t + 5]), pack_bf16_relu(v[8 * t + 6], v[8 * t + 7]));
        else        u = make_uint4(pack_bf16(v[8 * t + 0], v[8 * t + 1]), pack_bf16(v[8 * t + 2], v[8 * t + 3]),
                                   pack_bf16(v[8 * t + 4], v[8 * t + 5]), pack_bf16(v[8 * t + 6], v[8 * t + 7]));
        stage[lane * 4 + (t ^ sw)] = u;
    }
    __syncwarp();
    const int j = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int R = it * 8 + (lane >> 2);
        const uint4 val = stage[R * 4 + (j ^ ((R >> 1) & 3))];
        const int dr = __shfl_sync(0xffffffffu, drow, R);
        if (dr >= 0 && !(dbg & 2))
            reinterpret_cast<uint4*>(p.out + (long)dr * p.out_pitch + p.out_ch_off + c_base)[j] = val;
    }
    __syncwarp();
}

// Heads epilogue (CB_OUT_HEADS, one 32-column tile): fp32 NCHW stores, coalesced across the lanes of a warp
// (consecutive lanes = consecutive pixels of the same channel plane).
__device__ __forceinline__ void epilogue_heads_fast(const ConvParams& p, const RowDest& d, const float* __restrict__ s_bias,
                                                    float (&v)[32]) {
    if (d.row < 0) return;
    const int W = p.Wp - 2;
    const long HW = (long)(p.Hp - 2) * W;
    const long pix = (long)d.h * W + d.w;
#pragma unroll
    for (int s = 0; s < CB_MAX_HEADS; ++s) {
        if (s < p.n_heads) {
            const int c0 = p.head_c0[s], cn = p.head_cn[s];
            float* base = p.head_out[s] + (long)d.n * cn * HW + pix - (long)c0 * HW;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j >= c0 && j < c0 + cn) base[(long)j * HW] = v[j] + s_bias[j];
        }
    }
}

inline int fill_params(const cb_conv_desc* d, ConvParams& p) {
    if (d->n_ksteps < 1 || d->n_ksteps > CB_MAX_KSTEPS) return CB_ERR_ARG;
    if (d->block_n != 32 && d->block_n != 64 && d->block_n != 128 && d->block_n != 256) return CB_ERR_ARG;
    if (d->n_total % d->block_n != 0 || d->cout_mod % 32 != 0 || d->cout_mod <= 0) return CB_ERR_ARG;
    if (d->out_mode < 0 || d->out_mode > 3) return CB_ERR_ARG;
    if (d->out_mode != CB_OUT_HEADS && (d->out == nullptr || (d->out_pitch % 8) || (d->out_ch_off % 8))) return CB_ERR_ARG;
    if (d->out_mode == CB_OUT_UPSAMPLE && (d->up_k < 1 || d->cout_mod % d->block_n != 0)) return CB_ERR_ARG;
    if (d->residual && (d->res_pitch % 8)) return CB_ERR_ARG;
    if (d->in_pad < 0 || d->in_pad > 3) return CB_ERR_ARG;
    p.n_img = d->n_img; p.Hp = d->Hp; p.Wp = d->Wp; p.pad = d->in_pad ? d->in_pad : 1;
    if (p.pad != 1 && (d->residual != nullptr && (d->out_Hp != d->Hp || d->out_Wp != d->Wp) && false)) return CB_ERR_ARG;
    p.rows_total = (long)d->n_img * d->Hp * d->Wp;
    p.n_total = d->n_total; p.cout_mod = d->cout_mod; p.relu = d->relu;
    p.bias = d->bias;
    p.residual = (const __nv_bfloat16*)d->residual; p.res_pitch = d->res_pitch; p.res_lo_off = d->res_lo_off;
    p.out = (__nv_bfloat16*)d->out; p.out_pitch = d->out_pitch; p.out_ch_off = d->out_ch_off;
    p.out_lo_off = d->out_lo_off;
    p.out_mode = d->out_mode; p.up_k = d->up_k; p.out_Hp = d->out_Hp; p.out_Wp = d->out_Wp;
    if (d->out_mode == CB_OUT_PF && (d->out_Hp == 0 || d->out_Wp == 0)) { p.out_Hp = d->Hp; p.out_Wp = d->Wp; }
    p.out_plane_rows = d->out_plane_rows;
    for (int i = 0; i < CB_MAX_HEADS; ++i) { p.head_out[i] = d->head_out[i]; p.head_c0[i] = d->head_c0[i]; p.head_cn[i] = d->head_cn[i]; }
    p.n_heads = d->n_heads;
    p.n_ksteps = d->n_ksteps;
    for (int i = 0; i < d->n_ksteps; ++i) p.ksteps[i] = d->ksteps[i];
    return CB_OK;
}

}  // namespace cb
