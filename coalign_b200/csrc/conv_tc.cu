// Implicit-GEMM convolution on Blackwell tensor cores (sm_100a): TMA -> shared (SWIZZLE_128B) ->
// tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.  Persistent, warp specialised:
//   warp 0 : TMA producer (one elected thread)       warp 1 : MMA issuer (one elected thread)
//   warp 2 : TMEM allocator                          warps 4..7 : epilogue (TMEM lane quarter = warp%4)
// GEMM rows are flattened zero-padded output pixels (PF layout), so every filter tap is a constant row
// shift of a plain 2-D TMA box; stride-2 convolutions read phase-split (PS) inputs the same way.
// Replaces cuDNN conv2d/conv_transpose2d under
//   /root/reference/opencood/models/sub_modules/resblock.py:53-69 (BasicBlock),
//   base_bev_backbone_resnet.py:52-65,121-138 (deblocks), downsample_conv.py:18-24 (shrink header),
//   point_pillar_baseline_multiscale.py:126-133 (heads).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <mutex>
#include <cstdlib>
#include "conv_common.cuh"

namespace cb {

constexpr int BM = 128;       // GEMM rows per tile (= UMMA M, one TMEM lane per row)
constexpr int BK = 64;        // K per pipeline stage: 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;

// A pipeline stage holds KPS consecutive K-steps (each: 128x64 A tile + BNx64 weight tile).  Small-N tiles have
// little tensor work per K-step, so several K-steps share one barrier round trip (wait / expect_tx / commit).
template <int BN> struct TcCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int KSTEP_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int KPS = (BN == 256) ? 1 : 2;
    static constexpr int STAGE_BYTES = KPS * KSTEP_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // +1024: manual 1 KiB alignment
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;    // double-buffered accumulator
};


template <int BN, int KPS_OVR> struct TcCfgK : TcCfg<BN> {
    static constexpr int KPS = KPS_OVR ? KPS_OVR : TcCfg<BN>::KPS;
    static constexpr int STAGE_BYTES = KPS * TcCfg<BN>::KSTEP_BYTES;
};

// 12 warps.  The warp scheduler favours higher warp ids, so the two latency-critical single-thread roles get the
// highest ids on their sub-partitions and the 8 epilogue warps the lowest (TMEM lane quarter = warp % 4).
constexpr int TC_THREADS = 384;
constexpr int W_PRODUCER = 8, W_MMA = 9, W_ALLOC = 10;

// Epilogue of one accumulator tile for one thread: row q (TMEM lane), columns [c_lo, c_hi) of the BN-wide tile.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, const float* s_bias, uint4* stage, uint32_t t_row,
                                              long q, int n0, int c_lo, int c_hi, uint32_t tfull_bar, uint32_t acc_phase,
                                              int dbg) {
    const RowDest dst = decode_row(p, q, n0);
    const bool fast = (p.out_lo_off == 0) && (p.out_mode != CB_OUT_HEADS) && (p.res_lo_off == 0);
    const bool has_res = fast && p.residual != nullptr && dst.row >= 0 && !(dbg & 4);
    uint4 rcur[4] = {}, rnext[4] = {};
    const uint4* rptr = reinterpret_cast<const uint4*>(p.residual + q * (long)p.res_pitch + n0);
    if (has_res) {                                       // residual of the first chunk: issued before the MMA is done
#pragma unroll
        for (int t = 0; t < 4; ++t) rcur[t] = __ldg(rptr + (c_lo >> 3) + t);
    }
    mbar_wait_a(tfull_bar, acc_phase);
    tc_fence_after();
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
        if (has_res && c + 32 < c_hi) {
#pragma unroll
            for (int t = 0; t < 4; ++t) rnext[t] = __ldg(rptr + ((c + 32) >> 3) + t);
        }
        uint32_t r[32];
        tmem_ld32(t_row + c, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (dbg & 1) {                                   // experiment: no epilogue math / stores
            if (v[0] == 1.2345e-30f) p.out[0] = __float2bfloat16(v[1]);
        } else if (fast) {
            epilogue_chunk_fast(p, (int)dst.row, (n0 + c) % p.cout_mod, s_bias, rcur, has_res, v, stage,
                                (int)(threadIdx.x & 31), dbg);
        } else if (p.out_mode == CB_OUT_HEADS && BN == 32 && !p.relu && p.residual == nullptr) {
            epilogue_heads_fast(p, dst, s_bias, v);
        } else {
            epilogue_chunk(p, dst, q, n0 + c, v);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) rcur[t] = rnext[t];
    }
}

// PF-layout outputs: the half-tile (128 rows x GW columns) is assembled in shared memory (swizzled, conflict-free) and
// written with ONE TMA store - no per-thread global stores on the LSU/L1 path at all.  Halo / out-of-range rows are
// written as zeros, which is what the PF halo holds anyway.  `stage` = this column half's 16 KB buffer (1 KB aligned).
template <int BN>
__device__ __forceinline__ void epilogue_tile_tma(const ConvParams& p, const float* s_bias, uint4* stage,
                                                  const CUtensorMap* tmap_out, uint32_t t_row, long q, int m0, int n0,
                                                  int half, int q4, int lane, uint32_t tfull_bar, uint32_t acc_phase,
                                                  bool& store_pending) {
    constexpr int CW = BN / 2, GW = BN >= 128 ? 64 : 32, NG = CW / GW, CPG = GW / 32, CHR = GW / 8;
    const RowDest dst = decode_row(p, q, n0);
    const bool valid = dst.row >= 0;
    const bool has_res = p.residual != nullptr && valid;
    const uint4* rptr = reinterpret_cast<const uint4*>(p.residual + q * (long)p.res_pitch + n0 + half * CW);
    uint4 rcur[4] = {}, rnext[4] = {};
    if (has_res) {
#pragma unroll
        for (int t = 0; t < 4; ++t) rcur[t] = __ldg(rptr + t);
    }
    mbar_wait_a(tfull_bar, acc_phase);
    tc_fence_after();
    const int row_local = q4 * 32 + lane;
    const bool issuer = (q4 == 0 && lane == 0);
#pragma unroll 1
    for (int gi = 0; gi < NG; ++gi) {
        uint4 pk[CPG * 4];
#pragma unroll
        for (int cc = 0; cc < CPG; ++cc) {
            const int ci = gi * CPG + cc;                        // 32-column chunk index inside this half
            if (has_res && ci + 1 < NG * CPG) {
#pragma unroll
                for (int t = 0; t < 4; ++t) rnext[t] = __ldg(rptr + (ci + 1) * 4 + t);
            }
            uint32_t r[32];
            tmem_ld32(t_row + half * CW + ci * 32, r);
            tmem_ld_wait();
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + (n0 + half * CW + ci * 32) % p.cout_mod);
            float v[32];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float4 b = b4[t];
                v[4 * t + 0] = __uint_as_float(r[4 * t + 0]) + b.x; v[4 * t + 1] = __uint_as_float(r[4 * t + 1]) + b.y;
                v[4 * t + 2] = __uint_as_float(r[4 * t + 2]) + b.z; v[4 * t + 3] = __uint_as_float(r[4 * t + 3]) + b.w;
            }
            if (has_res) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    v[8 * t + 0] += bf16_lo(rcur[t].x); v[8 * t + 1] += bf16_hi(rcur[t].x);
                    v[8 * t + 2] += bf16_lo(rcur[t].y); v[8 * t + 3] += bf16_hi(rcur[t].y);
                    v[8 * t + 4] += bf16_lo(rcur[t].z); v[8 * t + 5] += bf16_hi(rcur[t].z);
                    v[8 * t + 6] += bf16_lo(rcur[t].w); v[8 * t + 7] += bf16_hi(rcur[t].w);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                uint4 u;
                if (p.relu) u = make_uint4(pack_bf16_relu(v[8 * t + 0], v[8 * t + 1]), pack_bf16_relu(v[8 * t + 2], v[8 * t + 3]),
                                           pack_bf16_relu(v[8 * t + 4], v[8 * t + 5]), pack_bf16_relu(v[8 * t + 6], v[8 * t + 7]));
                else        u = make_uint4(pack_bf16(v[8 * t + 0], v[8 * t + 1]), pack_bf16(v[8 * t + 2], v[8 * t + 3]),
                                           pack_bf16(v[8 * t + 4], v[8 * t + 5]), pack_bf16(v[8 * t + 6], v[8 * t + 7]));
                pk[cc * 4 + t] = valid ? u : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) rcur[t] = rnext[t];
        }
        if (store_pending && issuer) tma_store_wait_read0();     // previous store of this half has read the stage
        named_bar_sync(1 + half, 128);
        const int sw = GW == 64 ? (row_local & 7) : ((row_local >> 1) & 3);   // SWIZZLE_128B / SWIZZLE_64B chunk XOR
#pragma unroll
        for (int k = 0; k < CPG * 4; ++k) stage[row_local * CHR + (k ^ sw)] = pk[k];
        fence_proxy_async();                                     // generic-proxy writes -> visible to the TMA (async proxy)
        named_bar_sync(1 + half, 128);
        if (issuer) {
            tma_store_2d(tmap_out, smem_u32(stage), p.out_ch_off + (n0 + half * CW + gi * GW) % p.cout_mod, m0);
            tma_store_commit();
        }
        store_pending = true;
    }
}

template <int BN, int KPS_OVR = 0, int STAGES_OVR = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                    const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                    const __grid_constant__ ConvParams p, int m_tiles, int n_tiles, int dbg, int use_tma) {
    using Cfg = TcCfgK<BN, KPS_OVR>;
    constexpr int STAGES = STAGES_OVR ? STAGES_OVR : TcCfg<BN>::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[256];
    __shared__ __align__(1024) uint4 s_stage[2][1024];    // per column half: 128 rows x 128 B epilogue stage (TMA store) /
                                                          // per warp 2 KB slices for the transposed-STG path

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;                       // 1 KiB aligned (SWIZZLE_128B atoms)
    for (int i = threadIdx.x; i < 256; i += TC_THREADS) s_bias[i] = i < p.cout_mod ? p.bias[i] : 0.f;

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int total_tiles = m_tiles * n_tiles;
    const int nk = p.n_ksteps;

    if (warp == W_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_w);
        prefetch_tmap(&tmap_out);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_launch_dependents();          // the next kernel may start its own prologue as our CTAs retire
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                       // predecessor grid complete: activations / residual readable, outputs writable
    const uint32_t tmem_base = tmem_base_smem;

    // loop-invariant shared addresses / descriptors for the two single-thread hot loops
    const uint32_t smem_a0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);

    if (warp == W_PRODUCER) {
        // ------------------------------------------------------------------ TMA producer
        // (whole warp in the loop, elect only around the TMA instructions: see the MMA-issuer note in conv_gemm_halo64_kernel)
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
                const int m0 = mt * BM, n0 = nt * BN;
                for (int ks = 0; ks < nk; ks += Cfg::KPS) {
                    const int cnt = (nk - ks) < Cfg::KPS ? (nk - ks) : Cfg::KPS;
                    const uint32_t fb = full0 + stage * 8;
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx_a(fb, (uint32_t)cnt * Cfg::KSTEP_BYTES);
#pragma unroll
                        for (int j = 0; j < Cfg::KPS; ++j) {
                            if (j < cnt) {
                                const cb_kstep st = p.ksteps[ks + j];
                                const uint32_t sa = smem_a0 + stage * Cfg::STAGE_BYTES + j * Cfg::KSTEP_BYTES;
                                tma_load_2d_a(sa, st.a_sel ? &tmap_a1 : &tmap_a0, fb, (int)st.col, m0 + st.row_off);
                                tma_load_2d_a(sa + A_STAGE_BYTES, &tmap_w, fb, st.w_k, n0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            const uint64_t desc_a = make_sw128_desc(smem_a0);
            const uint32_t a_lo0 = (uint32_t)desc_a, hi = (uint32_t)(desc_a >> 32);
            const uint32_t b_lo0 = (uint32_t)make_sw128_desc(smem_a0 + A_STAGE_BYTES);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);       // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int ks = 0; ks < nk; ks += Cfg::KPS) {
                    const int cnt = (nk - ks) < Cfg::KPS ? (nk - ks) : Cfg::KPS;
                    // descriptor start-address field is (addr >> 4): stage / K-step strides and the 32-byte K advance add linearly
                    const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (Cfg::STAGE_BYTES >> 4));
                    const uint32_t b_lo = b_lo0 + (uint32_t)(stage * (Cfg::STAGE_BYTES >> 4));
                    mbar_wait_a(full0 + stage * 8, phase);           // TMA bytes landed
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < Cfg::KPS; ++j) {
                            if (j < cnt) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) {
                                    umma_bf16_lh(d_tmem, a_lo + (uint32_t)(j * (Cfg::KSTEP_BYTES >> 4) + 2 * k), hi,
                                                 b_lo + (uint32_t)(j * (Cfg::KSTEP_BYTES >> 4) + 2 * k), hi, idesc,
                                                 (ks > 0 || j > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        }
                        umma_commit_a(empty0 + stage * 8);           // frees the smem slot when the MMAs retire
                        if (ks + Cfg::KPS >= nk) umma_commit_a(tfull0 + buf * 8);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int q4 = warp & 3;                                     // TMEM lane quarter of this warp
        const int half = warp >> 2;                            // column half of the tile
        constexpr int CW = BN >= 64 ? BN / 2 : BN;                   // columns per warp (BN=32: half 1 idles)
        const int c_lo = half * CW, c_hi = (BN >= 64 || half == 0) ? c_lo + CW : c_lo;
        bool store_pending = false;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const long q = (long)mt * BM + q4 * 32 + lane;
            const uint32_t t_row = tmem_base + buf * BN + ((uint32_t)(q4 * 32) << 16);
            if (BN >= 64 && use_tma == 1)
                epilogue_tile_tma<(BN >= 64 ? BN : 64)>(p, s_bias, s_stage[half], &tmap_out, t_row, q, mt * BM, nt * BN, half,
                                                       q4, lane, tfull0 + buf * 8, acc_phase, store_pending);
            else
                epilogue_tile<BN>(p, s_bias, &s_stage[0][0] + warp * 128, t_row, q, nt * BN, c_lo, c_hi, tfull0 + buf * 8,
                                  acc_phase, dbg | (use_tma == 2 ? 8 : 0));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (store_pending && q4 == 0 && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// ----------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of 2 CTAs computes a 256 x BN tile.  Each CTA stages its own 128
// rows of A and HALF of the weight tile (BN/2 rows); the pair's tensor cores read both halves, so the shared-
// memory traffic per MAC drops (the single-CTA kernel is smem-bandwidth bound: TMA writes + MMA reads > 128 B/clk).
// Leader CTA (rank 0) issues every MMA; both CTAs run a TMA producer and an epilogue over their own TMEM lanes.
// ----------------------------------------------------------------------------------------------------------
template <int BN> struct Tc2Cfg {
    static constexpr int B_STAGE_BYTES = (BN / 2) * BK * 2;
    static constexpr int KSTEP_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;        // per CTA and K-step
    static constexpr int KPS = (BN == 256) ? 1 : 2;                          // K-steps per pipeline stage
    static constexpr int STAGE_BYTES = KPS * KSTEP_BYTES;
    static constexpr int STAGES = (BN == 256) ? 6 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN, int STAGES_OVR = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
conv_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                     const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                     const __grid_constant__ ConvParams p, int m_pairs, int n_tiles, int dbg, int use_tma) {
    using Cfg = Tc2Cfg<BN>;
    constexpr int STAGES = STAGES_OVR ? STAGES_OVR : Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];       // used in the leader CTA (both CTAs' TMA bytes land here)
    __shared__ __align__(8) uint64_t empty_bar[STAGES];      // per CTA; released by the leader's multicast commit
    __shared__ __align__(8) uint64_t tmem_full_bar[2];       // per CTA; multicast commit
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];      // leader's copy counts 16 epilogue warps (both CTAs)
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[256];
    __shared__ __align__(1024) uint4 s_stage[2][1024];

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < 256; i += TC_THREADS) s_bias[i] = i < p.cout_mod ? p.bias[i] : 0.f;

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int total_tiles = m_pairs * n_tiles;
    const int nk = p.n_ksteps;

    if (warp == W_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_w);
        prefetch_tmap(&tmap_out);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 16); }
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        tmem_alloc_2sm(&tmem_base_smem, Cfg::TMEM_COLS);
        tmem_relinquish_2sm();
    }
    pdl_launch_dependents();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_smem;

    const uint32_t smem_a0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);

    if (warp == W_PRODUCER) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
                const int mp = tile / n_tiles, nt = tile - mp * n_tiles;
                const int m0 = (mp * 2 + (int)rank) * BM, n0 = nt * BN + (int)rank * (BN / 2);
                for (int ks = 0; ks < nk; ks += Cfg::KPS) {
                    const int cnt = (nk - ks) < Cfg::KPS ? (nk - ks) : Cfg::KPS;
                    const uint32_t fb = (full0 + stage * 8) & 0xFEFFFFFFu;        // leader CTA's barrier
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx_a(full0 + stage * 8, 2u * (uint32_t)cnt * Cfg::KSTEP_BYTES);
#pragma unroll
                        for (int j = 0; j < Cfg::KPS; ++j) {
                            if (j < cnt) {
                                const cb_kstep st = p.ksteps[ks + j];
                                const uint32_t sa = smem_a0 + stage * Cfg::STAGE_BYTES + j * Cfg::KSTEP_BYTES;
                                tma_load_2d_2sm_a(sa, st.a_sel ? &tmap_a1 : &tmap_a0, fb, (int)st.col, m0 + st.row_off);
                                tma_load_2d_2sm_a(sa + A_STAGE_BYTES, &tmap_w, fb, st.w_k, n0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
            const uint64_t desc_a = make_sw128_desc(smem_a0);
            const uint32_t a_lo0 = (uint32_t)desc_a, hi = (uint32_t)(desc_a >> 32);
            const uint32_t b_lo0 = (uint32_t)make_sw128_desc(smem_a0 + A_STAGE_BYTES);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int ks = 0; ks < nk; ks += Cfg::KPS) {
                    const int cnt = (nk - ks) < Cfg::KPS ? (nk - ks) : Cfg::KPS;
                    const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (Cfg::STAGE_BYTES >> 4));
                    const uint32_t b_lo = b_lo0 + (uint32_t)(stage * (Cfg::STAGE_BYTES >> 4));
                    mbar_wait_a(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int j = 0; j < Cfg::KPS; ++j) {
                            if (j < cnt) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) {
                                    umma_bf16_2sm_lh(d_tmem, a_lo + (uint32_t)(j * (Cfg::KSTEP_BYTES >> 4) + 2 * k), hi,
                                                     b_lo + (uint32_t)(j * (Cfg::KSTEP_BYTES >> 4) + 2 * k), hi, idesc,
                                                     (ks > 0 || j > 0 || k > 0) ? 1u : 0u);
                                }
                            }
                        }
                        umma_commit_2sm_a(empty0 + stage * 8);
                        if (ks + Cfg::KPS >= nk) umma_commit_2sm_a(tfull0 + buf * 8);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows, 8 warps)
        const int q4 = warp & 3;
        const int half = warp >> 2;
        constexpr int CW = BN / 2;
        const int c_lo = half * CW, c_hi = c_lo + CW;
        bool store_pending = false;
        int it = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
            const int mp = tile / n_tiles, nt = tile - mp * n_tiles;
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const long q = (long)(mp * 2 + (int)rank) * BM + q4 * 32 + lane;
            const uint32_t t_row = tmem_base + buf * BN + ((uint32_t)(q4 * 32) << 16);
            if (use_tma == 1)
                epilogue_tile_tma<BN>(p, s_bias, s_stage[half], &tmap_out, t_row, q, (mp * 2 + (int)rank) * BM, nt * BN, half,
                                      q4, lane, tfull0 + buf * 8, acc_phase, store_pending);
            else
                epilogue_tile<BN>(p, s_bias, &s_stage[0][0] + warp * 128, t_row, q, nt * BN, c_lo, c_hi, tfull0 + buf * 8,
                                  acc_phase, (dbg & 7) | (use_tma == 2 ? 8 : 0));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[buf], 0);      // leader's barrier
        }
        if (store_pending && q4 == 0 && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Channel-major ("transposed") variant for Cout = 128 layers: the weight tile (128 output channels x 64) is the UMMA
// A operand and a 256-pixel activation tile the B operand, so D^T[channel][pixel] (128 TMEM lanes x 256 columns)
// accumulates 128 x 256 x 16 MACs per instruction instead of 128 x 128 x 16 - the pixel rows, which dominate the
// shared-memory traffic, are read once per 256-wide instruction.  Same descriptors / K-step tables as above; only
// the operand roles and the epilogue (TMEM lane = channel, so tiles are transposed back to pixel-major rows through
// a 2 KB per-warp shared-memory stage) differ.
// ----------------------------------------------------------------------------------------------------------
constexpr int TP = 256;                                   // pixels per tile (UMMA N)
struct TctCfg {
    static constexpr int P_BYTES = TP * BK * 2;           // activation tile of one K-step
    static constexpr int W_BYTES = 128 * BK * 2;          // weight tile of one K-step
    static constexpr int KSTEP_BYTES = P_BYTES + W_BYTES;
    static constexpr int STAGES = 4;
    static constexpr int SMEM_BYTES = STAGES * KSTEP_BYTES + 1024;
    static constexpr int TMEM_COLS = 2 * TP;
};

// Epilogue of the channel-major kernels (one of the 8 epilogue warps): TMEM lane = output channel, so every 32-pixel
// chunk is transposed back to pixel-major rows through the warp's 2 KB shared-memory stage.
__device__ __forceinline__ void tct_epilogue(const ConvParams& p, const float* s_bias, uint4* stage, uint32_t tmem_base,
                                             uint32_t tfull0, uint64_t* tmem_empty_bar, int m_tiles, int n_cblk, int warp,
                                             int lane) {
        // epilogue: warp -> channels [32*q4, +32) of the item's 128-channel block (TMEM lane quarter), pixel columns
        // [128*half, +128) in 4 chunks of 32
        const int q4 = warp & 3, half = warp >> 2;
        const uint16_t* stage16 = reinterpret_cast<const uint16_t*>(stage);
        uint16_t* stage16w = reinterpret_cast<uint16_t*>(stage);
        const int rsub = lane >> 2, csub = lane & 3;                 // row-in-8 / 16-byte chunk for the pixel-major accesses
        const bool has_res = p.residual != nullptr;
        int it = 0;
        for (int item = blockIdx.x; item < m_tiles * n_cblk; item += gridDim.x, ++it) {
            const int tile = item / n_cblk;
            const int c0 = (item - tile * n_cblk) * 128 + q4 * 32;    // first channel of this warp
            const float bias = s_bias[c0 + lane];
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const long pb0 = (long)tile * TP + half * 128;
            const uint32_t t_row = tmem_base + buf * TP + half * 128 + ((uint32_t)(q4 * 32) << 16);
            uint4 rcur[4] = {}, rnext[4] = {};
            if (has_res) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const long q = pb0 + i * 8 + rsub;
                    if (q < p.rows_total)
                        rcur[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + q * (long)p.res_pitch + c0) + csub);
                }
            }
            mbar_wait_a(tfull0 + buf * 8, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
                const long pb = pb0 + ci * 32;
                if (has_res && ci < 3) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long q = pb + 32 + i * 8 + rsub;
                        rnext[i] = make_uint4(0u, 0u, 0u, 0u);
                        if (q < p.rows_total)
                            rnext[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + q * (long)p.res_pitch + c0) + csub);
                    }
                }
                const int drow = (int)decode_row(p, pb + lane, 0).row;       // destination row of pixel (pb + lane), -1 = halo
                uint32_t r[32];
                tmem_ld32(t_row + ci * 32, r);
                float v[32];
                if (has_res) {                                       // pixel-major residual -> this thread's channel column
#pragma unroll
                    for (int i = 0; i < 4; ++i) stage[(i * 8 + rsub) * 4 + csub] = rcur[i];
                    __syncwarp();
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = __uint_as_float(r[j]) + bias + __uint_as_float((uint32_t)stage16[j * 32 + lane] << 16);
                    __syncwarp();
                } else {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const uint32_t pk = p.relu ? pack_bf16_relu(v[j], v[j + 1]) : pack_bf16(v[j], v[j + 1]);
                    stage16w[j * 32 + lane] = (uint16_t)(pk & 0xFFFFu);
                    stage16w[(j + 1) * 32 + lane] = (uint16_t)(pk >> 16);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int R = i * 8 + rsub;
                    const uint4 val = stage[R * 4 + csub];
                    const int dr = __shfl_sync(0xffffffffu, drow, R);
                    if (dr >= 0)
                        reinterpret_cast<uint4*>(p.out + (long)dr * p.out_pitch + p.out_ch_off + c0)[csub] = val;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tct_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                     const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ ConvParams p, int m_tiles,
                     int n_cblk) {
    using Cfg = TctCfg;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[256];
    __shared__ __align__(16) uint4 s_stage[8][128];       // per epilogue warp: 32 pixels x 64 B

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < 128 * n_cblk; i += TC_THREADS) s_bias[i] = p.bias[i];

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int nk = p.n_ksteps;

    if (warp == W_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_w);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_smem;

    const uint32_t smem_p0 = smem_u32(smem);
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);

    if (warp == W_PRODUCER) {
        // warp-uniform loops, elect only around the TMA / tcgen05 instructions (see conv_gemm_halo64_kernel)
        {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < m_tiles * n_cblk; item += gridDim.x) {
                const int tile = item / n_cblk, n0 = (item - tile * n_cblk) * 128;
                const int m0 = tile * TP;
                for (int ks = 0; ks < nk; ++ks) {
                    const uint32_t fb = full0 + stage * 8;
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx_a(fb, (uint32_t)Cfg::KSTEP_BYTES);
                        const cb_kstep st = p.ksteps[ks];
                        const uint32_t sa = smem_p0 + stage * Cfg::KSTEP_BYTES;
                        tma_load_2d_a(sa, st.a_sel ? &tmap_a1 : &tmap_a0, fb, (int)st.col, m0 + st.row_off);
                        tma_load_2d_a(sa + Cfg::P_BYTES, &tmap_w, fb, st.w_k, n0);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        {
            constexpr uint32_t idesc = make_idesc_bf16(128, TP);          // M = channels, N = pixels
            const uint64_t desc_p = make_sw128_desc(smem_p0);
            const uint32_t p_lo0 = (uint32_t)desc_p, hi = (uint32_t)(desc_p >> 32);
            const uint32_t w_lo0 = (uint32_t)make_sw128_desc(smem_p0 + Cfg::P_BYTES);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < m_tiles * n_cblk; item += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * TP;
                for (int ks = 0; ks < nk; ++ks) {
                    const uint32_t p_lo = p_lo0 + (uint32_t)(stage * (Cfg::KSTEP_BYTES >> 4));
                    const uint32_t w_lo = w_lo0 + (uint32_t)(stage * (Cfg::KSTEP_BYTES >> 4));
                    mbar_wait_a(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16_lh(d_tmem, w_lo + (uint32_t)(2 * k), hi, p_lo + (uint32_t)(2 * k), hi, idesc,
                                         (ks > 0 || k > 0) ? 1u : 0u);
                        umma_commit_a(empty0 + stage * 8);
                        if (ks + 1 >= nk) umma_commit_a(tfull0 + buf * 8);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        tct_epilogue(p, s_bias, s_stage[warp], tmem_base, tfull0, tmem_empty_bar, m_tiles, n_cblk, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Halo variant for Cout = 64 layers (level 0).  The pixel-major kernel above is shared-memory-bandwidth bound at
// N = 64 (profiles/r1_conv_analysis.md section 3: 24 KB of TMA writes + 24 KB of operand reads per 128-cycle K-step).
// Two changes cut the TMA writes ~4x:
//   * the three taps of one filter row (row shifts d-1, d, d+1 of the flattened PF pixel index) are served by ONE
//     TMA box of 130 rows; the MMA descriptors of the three taps start 0 / 128 / 256 bytes into that box
//     (SWIZZLE_128B is a function of the shared-memory address bits, so a row-shifted start stays consistent with
//     what the TMA wrote);
//   * the whole weight matrix (<= 10 K-steps x 8 KB) is loaded once per CTA and stays resident.
// Work items = groups of K-steps sharing one A box, built by the host from the K-step table (HaloItems).
// ----------------------------------------------------------------------------------------------------------
constexpr int HALO_ROWS = BM + 2;
constexpr int HALO_BYTES = HALO_ROWS * BK * 2;            // 16640 bytes per A box
constexpr int HALO_STRIDE = 17 * 1024;                    // stage pitch: keeps every stage 1 KB aligned
constexpr int HALO_MAX_ITEMS = 60;
struct HaloItems {
    int n;
    uint8_t first[HALO_MAX_ITEMS];                        // index of the item's first K-step
    uint8_t nsub[HALO_MAX_ITEMS];                         // 1..3 K-steps with consecutive row shifts
};
struct Halo64Cfg {
    static constexpr int BN = 64;
    static constexpr int W_BYTES = BN * BK * 2;           // 8 KB per K-step
    static constexpr int MAX_KSTEPS = 10;
    static constexpr int STAGES = 6;
    static constexpr int SMEM_BYTES = 1024 + STAGES * HALO_STRIDE + MAX_KSTEPS * W_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_halo64_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                        const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_out,
                        const __grid_constant__ ConvParams p, const __grid_constant__ HaloItems items, int m_tiles,
                        int bo_mode, int use_tma) {
    using Cfg = Halo64Cfg;
    constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[256];
    __shared__ __align__(1024) uint4 s_stage[2][1024];

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < 256; i += TC_THREADS) s_bias[i] = i < p.cout_mod ? p.bias[i] : 0.f;

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int nk = p.n_ksteps;
    const int n_items = items.n;

    if (warp == W_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_w);
        prefetch_tmap(&tmap_out);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        mbar_init(&w_bar, 1);
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const uint32_t smem_a0 = smem_u32(smem);
    const uint32_t smem_w0 = smem_a0 + STAGES * HALO_STRIDE;
    const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
    const uint32_t wbar = smem_u32(&w_bar);

    if (warp == W_PRODUCER) {
        // ------------------------------------------------------------------ TMA producer
        {
            // resident weights: independent of the predecessor kernel, so issued before the PDL wait
            if (elect_one()) {
                mbar_expect_tx_a(wbar, (uint32_t)nk * Cfg::W_BYTES);
                for (int j = 0; j < nk; ++j)
                    tma_load_2d_a(smem_w0 + j * Cfg::W_BYTES, &tmap_w, wbar, p.ksteps[j].w_k, 0);
            }
            __syncwarp();
            pdl_wait();
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
                const int m0 = tile * BM;
                for (int ii = 0; ii < n_items; ++ii) {
                    const cb_kstep st = p.ksteps[items.first[ii]];
                    const uint32_t fb = full0 + stage * 8;
                    mbar_wait_a(empty0 + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx_a(fb, (uint32_t)HALO_BYTES);
                        tma_load_2d_a(smem_a0 + stage * HALO_STRIDE, st.a_sel ? &tmap_a1 : &tmap_a0, fb, (int)st.col,
                                      m0 + st.row_off);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        // The WHOLE warp runs the loop and only the tcgen05 instructions sit under elect: stage / phase / descriptors are
        // then warp-uniform values the compiler keeps in uniform registers.  With the loop inside `if (elect_one())` the
        // same values live in per-thread registers and every UTCHMMA operand goes through R2UR: measured 50 cycles of
        // issue per MMA + ~290 cycles of bookkeeping per box against 32 tensor cycles per N = 64 MMA
        // (profiles/r2_exp_halo64_issue.txt) - the level-0 layers were bound by this one thread.
        {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            const uint64_t desc_a = make_sw128_desc(smem_a0), desc_w = make_sw128_desc(smem_w0);
            const uint32_t a_lo0 = (uint32_t)desc_a, w_lo0 = (uint32_t)desc_w, hi = (uint32_t)(desc_a >> 32);
            const uint32_t bo_shift = bo_mode ? (1u << 17) : 0u;     // matrix-base-offset field = bits 49..51
            mbar_wait_a(wbar, 0);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int ii = 0; ii < n_items; ++ii) {
                    const int ns = items.nsub[ii];
                    const uint32_t a_lo = a_lo0 + (uint32_t)(stage * (HALO_STRIDE >> 4));
                    const uint32_t w_lo = w_lo0 + (uint32_t)(items.first[ii] * (Cfg::W_BYTES >> 4));
                    mbar_wait_a(full0 + stage * 8, phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            // tap s of the filter row: same box, start shifted by s rows (128 B); bo_mode 1 additionally
                            // records the shift in the descriptor's matrix-base-offset field
                            if (s < ns) {
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k)
                                    umma_bf16_lh(d_tmem, a_lo + (uint32_t)(s * 8 + 2 * k), hi + (uint32_t)s * bo_shift,
                                                 w_lo + (uint32_t)(s * (Cfg::W_BYTES >> 4) + 2 * k), hi, idesc,
                                                 (ii > 0 || s > 0 || k > 0) ? 1u : 0u);
                            }
                        }
                        umma_commit_a(empty0 + stage * 8);
                        if (ii + 1 == n_items) umma_commit_a(tfull0 + buf * 8);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        pdl_wait();
        const int q4 = warp & 3;
        const int half = warp >> 2;
        constexpr int CW = BN / 2;
        const int c_lo = half * CW, c_hi = c_lo + CW;
        bool store_pending = false;
        int it = 0;
        for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const long q = (long)tile * BM + q4 * 32 + lane;
            const uint32_t t_row = tmem_base + buf * BN + ((uint32_t)(q4 * 32) << 16);
            if (use_tma == 1)
                epilogue_tile_tma<BN>(p, s_bias, s_stage[half], &tmap_out, t_row, q, tile * BM, 0, half, q4, lane,
                                      tfull0 + buf * 8, acc_phase, store_pending);
            else
                epilogue_tile<BN>(p, s_bias, &s_stage[0][0] + warp * 128, t_row, q, 0, c_lo, c_hi, tfull0 + buf * 8,
                                  acc_phase, use_tma == 2 ? 8 : 0);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (store_pending && q4 == 0 && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Channel-major kernel with halo boxes (Cout = 128 layers): as conv_gemm_tct_kernel, but the 256-pixel activation tile
// (UMMA B operand) of the three taps of one filter row is loaded once as a 258-row box (TMA boxes are limited to 256
// rows: 256 + 2) and the tap shift goes into the B descriptor's start address.  Activation boxes and weight tiles run
// through separate rings: 3 x 33 KB pixel boxes, 6 x 16 KB weight tiles.
// ----------------------------------------------------------------------------------------------------------
struct TctHaloCfg {
    static constexpr int P_ROWS = TP + 2;
    static constexpr int P_BYTES = P_ROWS * BK * 2;       // 33024
    static constexpr int P_STRIDE = 33 * 1024;
    static constexpr int W_BYTES = 128 * BK * 2;          // 16384
    static constexpr int PST = 3, WST = 6;
    static constexpr int SMEM_BYTES = 1024 + PST * P_STRIDE + WST * W_BYTES;
    static constexpr int TMEM_COLS = 2 * TP;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_gemm_tct_halo_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                          const __grid_constant__ CUtensorMap tmap_a0t, const __grid_constant__ CUtensorMap tmap_a1t,
                          const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ ConvParams p,
                          const __grid_constant__ HaloItems items, int m_tiles, int n_cblk, int bo_mode) {
    using Cfg = TctHaloCfg;
    constexpr int PST = Cfg::PST, WST = Cfg::WST;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t pfull_bar[PST];
    __shared__ __align__(8) uint64_t pempty_bar[PST];
    __shared__ __align__(8) uint64_t wfull_bar[WST];
    __shared__ __align__(8) uint64_t wempty_bar[WST];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float s_bias[256];
    __shared__ __align__(16) uint4 s_stage[8][128];

    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < 128 * n_cblk; i += TC_THREADS) s_bias[i] = p.bias[i];

    const int warp = warp_idx_uniform();
    const int lane = threadIdx.x & 31;
    const int n_items = items.n;

    if (warp == W_PRODUCER && lane == 0) {
        prefetch_tmap(&tmap_a0);
        prefetch_tmap(&tmap_a1);
        prefetch_tmap(&tmap_a0t);
        prefetch_tmap(&tmap_a1t);
        prefetch_tmap(&tmap_w);
    }
    if (warp == W_MMA && lane == 0) {
        for (int s = 0; s < PST; ++s) { mbar_init(&pfull_bar[s], 1); mbar_init(&pempty_bar[s], 1); }
        for (int s = 0; s < WST; ++s) { mbar_init(&wfull_bar[s], 1); mbar_init(&wempty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == W_ALLOC) {
        tmem_alloc(&tmem_base_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();
    const uint32_t tmem_base = tmem_base_smem;

    const uint32_t smem_p0 = smem_u32(smem);
    const uint32_t smem_w0 = smem_p0 + PST * Cfg::P_STRIDE;
    const uint32_t pfull0 = smem_u32(&pfull_bar[0]), pempty0 = smem_u32(&pempty_bar[0]);
    const uint32_t wfull0 = smem_u32(&wfull_bar[0]), wempty0 = smem_u32(&wempty_bar[0]);
    const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);

    if (warp == W_PRODUCER) {
        // warp-uniform loops, elect only around the TMA / tcgen05 instructions (see conv_gemm_halo64_kernel)
        {
            int ps = 0, ws = 0; uint32_t pphase = 0, wphase = 0;
            for (int item = blockIdx.x; item < m_tiles * n_cblk; item += gridDim.x) {
                const int tile = item / n_cblk, n0 = (item - tile * n_cblk) * 128;
                const int m0 = tile * TP;
                for (int ii = 0; ii < n_items; ++ii) {
                    const int first = items.first[ii], ns = items.nsub[ii];
                    const cb_kstep st = p.ksteps[first];
                    const uint32_t pb = pfull0 + ps * 8;
                    const uint32_t sa = smem_p0 + ps * Cfg::P_STRIDE;
                    mbar_wait_a(pempty0 + ps * 8, pphase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx_a(pb, (uint32_t)Cfg::P_BYTES);
                        tma_load_2d_a(sa, st.a_sel ? &tmap_a1 : &tmap_a0, pb, (int)st.col, m0 + st.row_off);
                        tma_load_2d_a(sa + TP * BK * 2, st.a_sel ? &tmap_a1t : &tmap_a0t, pb, (int)st.col,
                                      m0 + st.row_off + TP);
                    }
                    __syncwarp();
                    if (++ps == PST) { ps = 0; pphase ^= 1; }
                    for (int s = 0; s < ns; ++s) {
                        const uint32_t wb = wfull0 + ws * 8;
                        mbar_wait_a(wempty0 + ws * 8, wphase ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx_a(wb, (uint32_t)Cfg::W_BYTES);
                            tma_load_2d_a(smem_w0 + ws * Cfg::W_BYTES, &tmap_w, wb, p.ksteps[first + s].w_k, n0);
                        }
                        __syncwarp();
                        if (++ws == WST) { ws = 0; wphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        {
            constexpr uint32_t idesc = make_idesc_bf16(128, TP);          // M = channels, N = pixels
            const uint64_t desc_p = make_sw128_desc(smem_p0);
            const uint32_t p_lo0 = (uint32_t)desc_p, hi = (uint32_t)(desc_p >> 32);
            const uint32_t w_lo0 = (uint32_t)make_sw128_desc(smem_w0);
            const uint32_t bo_shift = bo_mode ? (1u << 17) : 0u;     // matrix-base-offset field = descriptor bits 49..51
            int ps = 0, ws = 0; uint32_t pphase = 0, wphase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < m_tiles * n_cblk; item += gridDim.x, ++it) {
                const int buf = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait_a(tempty0 + buf * 8, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * TP;
                for (int ii = 0; ii < n_items; ++ii) {
                    const int ns = items.nsub[ii];
                    const uint32_t p_lo = p_lo0 + (uint32_t)(ps * (Cfg::P_STRIDE >> 4));
                    mbar_wait_a(pfull0 + ps * 8, pphase);
                    for (int s = 0; s < ns; ++s) {
                        const uint32_t pd = p_lo + (uint32_t)(s * 8), pd_hi = hi + (uint32_t)s * bo_shift;
                        const uint32_t wd = w_lo0 + (uint32_t)(ws * (Cfg::W_BYTES >> 4));
                        mbar_wait_a(wfull0 + ws * 8, wphase);
                        tc_fence_after();
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_bf16_lh(d_tmem, wd + (uint32_t)(2 * k), hi, pd + (uint32_t)(2 * k), pd_hi, idesc,
                                             (ii > 0 || s > 0 || k > 0) ? 1u : 0u);
                            umma_commit_a(wempty0 + ws * 8);
                            if (s + 1 == ns) {
                                umma_commit_a(pempty0 + ps * 8);
                                if (ii + 1 == n_items) umma_commit_a(tfull0 + buf * 8);
                            }
                        }
                        __syncwarp();
                        if (++ws == WST) { ws = 0; wphase ^= 1; }
                    }
                    if (++ps == PST) { ps = 0; pphase ^= 1; }
                }
            }
        }
    } else if (warp < 8) {
        tct_epilogue(p, s_bias, s_stage[warp], tmem_base, tfull0, tmem_empty_bar, m_tiles, n_cblk, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_ALLOC) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// bf16 [rows][pitch] row-major; box = 64 columns x box_rows rows, 128-byte swizzle, zero OOB fill.
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_rows,
                     int box_cols = 64) {
    auto enc = get_encode();
    if (!enc) return CB_ERR_DRIVER;
    if (((uintptr_t)base & 15) || (pitch_elems * 2) % 16) return CB_ERR_ARG;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CB_OK : CB_ERR_DRIVER;
}

static int dbg_flags() { return cb::opt_get(CB_OPT_CONV_DEBUG); }

template <int BN, int KPS, int ST>
static int launch_st(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& tout,
                     int use_tma, const ConvParams& p, int m_tiles, int n_tiles, int max_ctas, cudaStream_t stream) {
    constexpr int SMEM = ST * KPS * TcCfg<BN>::KSTEP_BYTES + 1024;
    static_assert(SMEM <= 227 * 1024, "stage configuration exceeds shared memory");
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, KPS, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int grid = m_tiles * n_tiles;
    int cap = max_ctas > 0 ? max_ctas : 148;
    if (grid > cap) grid = cap;
    cudaError_t le = launch_pdl(conv_gemm_tc_kernel<BN, KPS, ST>, dim3(grid), dim3(TC_THREADS), SMEM, stream, a0, a1, w, tout,
                                p, m_tiles, n_tiles, dbg_flags() & 7, use_tma);
    if (le != cudaSuccess) return (int)le;
    return CB_OK;
}

template <int BN>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& tout, int use_tma,
                  const ConvParams& p, int m_tiles, int n_tiles, int max_ctas, cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    {                                                     // experiment: K-steps-per-stage / stage-count variants
        const int f = dbg_flags() >> 3;
        if (BN == 64 && f == 1) return launch_st<64, 1, 8>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 64 && f == 2) return launch_st<64, 2, 4>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 64 && f == 3) return launch_st<64, 4, 2>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 128 && f == 1) return launch_st<128, 1, 6>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 128 && f == 2) return launch_st<128, 3, 2>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 128 && f == 3) return launch_st<128, 2, 3>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
        if (BN == 256 && f == 2) return launch_st<256, 2, 2>(a0, a1, w, tout, use_tma, p, m_tiles, n_tiles, max_ctas, stream);
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = m_tiles * n_tiles;
    int cap = max_ctas > 0 ? max_ctas : sms;
    if (grid > cap) grid = cap;
    cudaError_t le = launch_pdl(conv_gemm_tc_kernel<BN, 0, 0>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, a0, a1,
                                w, tout, p, m_tiles, n_tiles, dbg_flags() & 7, use_tma);
    if (le != cudaSuccess) return (int)le;
    return CB_OK;
}


template <int BN>
static int launch2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& tout, int use_tma,
                   const ConvParams& p, int m_tiles, int n_tiles, int max_clusters, cudaStream_t stream) {
    using Cfg = Tc2Cfg<BN>;
    const int dbg = dbg_flags();
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_tc2_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(conv_gemm_tc2_kernel<BN, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            Cfg::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_pairs = (m_tiles + 1) / 2;
    int clusters = m_pairs * n_tiles;
    int cap = max_clusters > 0 ? max_clusters : sms / 2;
    if (clusters > cap) clusters = cap;
    cudaError_t le;
    if (dbg & 32)           // experiment: half the pipeline depth
        le = launch_pdl(conv_gemm_tc2_kernel<BN, 3>, dim3(2 * clusters), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, a0, a1, w,
                        tout, p, m_pairs, n_tiles, dbg, use_tma);
    else
        le = launch_pdl(conv_gemm_tc2_kernel<BN, 0>, dim3(2 * clusters), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, a0, a1, w,
                        tout, p, m_pairs, n_tiles, dbg, use_tma);
    if (le != cudaSuccess) return (int)le;
    return CB_OK;
}

}  // namespace cb


// Group K-steps that read the same channel block at consecutive row shifts (the taps of one 3x3 filter row in the
// flattened PF row space): each group is served by one halo box.
static int build_halo_items(const cb_conv_desc* d, cb::HaloItems& items) {
    items.n = 0;
    for (int i = 0; i < d->n_ksteps;) {
        int ns = 1;
        while (ns < 3 && i + ns < d->n_ksteps && d->ksteps[i + ns].a_sel == d->ksteps[i].a_sel &&
               d->ksteps[i + ns].col == d->ksteps[i].col && d->ksteps[i + ns].row_off == d->ksteps[i].row_off + ns)
            ++ns;
        if (items.n >= cb::HALO_MAX_ITEMS || i > 255) return CB_ERR_ARG;
        items.first[items.n] = (uint8_t)i;
        items.nsub[items.n] = (uint8_t)ns;
        ++items.n;
        i += ns;
    }
    return CB_OK;
}

// Direct epilogue (two 256-bit stores per thread and 32-column chunk instead of the shared-memory transposition):
// bf16 outputs whose rows and channel offsets keep every 64-byte chunk 32-byte aligned.  cb_set_option(CB_OPT_EPI_DIRECT, 0) disables it.
static bool epi_direct_ok(const cb_conv_desc* d) {
    const bool enabled = cb::opt_get(CB_OPT_EPI_DIRECT) != 0;
    return enabled && d->out_mode != CB_OUT_HEADS && d->out_lo_off == 0 && d->res_lo_off == 0 && d->out_pitch % 16 == 0 &&
           d->out_ch_off % 16 == 0 && ((uintptr_t)d->out % 32) == 0 && d->block_n >= 64;
}

static int conv_gemm_impl(const cb_conv_desc* d, int max_ctas, void* stream, bool pair) {
    using namespace cb;
    if (!d) return CB_ERR_ARG;
    static thread_local ConvParams p;
    int rc = fill_params(d, p);
    if (rc) return rc;
    for (int i = 0; i < d->n_ksteps; ++i) {
        const cb_kstep& s = d->ksteps[i];
        if (s.a_sel > 1 || d->a_ptr[s.a_sel] == nullptr) return CB_ERR_ARG;
        if (s.col % 8 || s.col + 64 > d->a_pitch[s.a_sel]) return CB_ERR_ARG;
        if (s.w_k % 8 || s.w_k < 0 || s.w_k + 64 > d->w_k_total) return CB_ERR_ARG;
    }
    CUtensorMap ta0, ta1, tw;
    rc = make_tmap(&ta0, d->a_ptr[0], d->a_rows[0], d->a_pitch[0], d->a_pitch[0], BM);
    if (rc) return rc;
    if (d->a_ptr[1]) {
        rc = make_tmap(&ta1, d->a_ptr[1], d->a_rows[1], d->a_pitch[1], d->a_pitch[1], BM);
        if (rc) return rc;
    } else {
        ta1 = ta0;
    }
    if (pair && (d->block_n < 64 || d->w_rows < d->block_n)) pair = false;      // heads (N=32) stay single-CTA
    rc = make_tmap(&tw, d->w_ptr, d->w_rows, d->w_k_total, d->w_k_total, pair ? d->block_n / 2 : d->block_n);
    if (rc) return rc;
    const int m_tiles = (int)((p.rows_total + BM - 1) / BM);
    const int n_tiles = d->n_total / d->block_n;
    cudaStream_t st = (cudaStream_t)stream;
    // PF-layout bf16 outputs go out through a TMA store (half-tile = 128 rows x 64 or 32 columns)
    CUtensorMap tout = tw;
    int use_tma = 0;
    {
        // TMA-store epilogue: measured slower than the transposed 64-byte STG path on the residual layers (BN=64: 127 vs
        // 107 us) and equal elsewhere (profiles/r1_exp_halo.txt), so it is opt-in (CB_OPT_TMA_STORE)
        const bool no_tma_store = cb::opt_get(CB_OPT_TMA_STORE) == 0;
        if (!no_tma_store && d->out_mode == CB_OUT_PF && d->out_lo_off == 0 && d->res_lo_off == 0 && d->block_n >= 64 &&
            !(dbg_flags() & 7) && p.pad == 1 && p.out_Hp == p.Hp && p.out_Wp == p.Wp) {
            rc = make_tmap(&tout, d->out, p.rows_total, d->out_pitch, d->out_pitch, BM, d->block_n >= 128 ? 64 : 32);
            if (rc) return rc;
            use_tma = 1;
        }
    }
    if (!use_tma && epi_direct_ok(d)) use_tma = 2;
    if (pair) {
        switch (d->block_n) {
            case 64: return launch2<64>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
            case 128: return launch2<128>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
            case 256: return launch2<256>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
        }
        return CB_ERR_ARG;
    }
    switch (d->block_n) {
        case 32: return launch<32>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
        case 64: return launch<64>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
        case 128: return launch<128>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
        case 256: return launch<256>(ta0, ta1, tw, tout, use_tma, p, m_tiles, n_tiles, max_ctas, st);
    }
    return CB_ERR_ARG;
}

extern "C" int cb_conv_gemm(const cb_conv_desc* d, int max_ctas, void* stream) {
    return conv_gemm_impl(d, max_ctas, stream, false);
}

extern "C" int cb_conv_gemm_pair(const cb_conv_desc* d, int max_clusters, void* stream) {
    return conv_gemm_impl(d, max_clusters, stream, true);
}

/* Channel-major tensor-core path for Cout = 128 layers (see conv_gemm_tct_kernel). */
extern "C" int cb_conv_gemm_t(const cb_conv_desc* d, int max_ctas, void* stream) {
    using namespace cb;
    if (!d) return CB_ERR_ARG;
    static thread_local ConvParams p;
    int rc = fill_params(d, p);
    if (rc) return rc;
    if ((d->n_total != 128 && d->n_total != 256) || d->cout_mod != d->n_total || d->w_rows < d->n_total) return CB_ERR_ARG;
    const int n_cblk = d->n_total / 128;                      // 128-channel blocks: items = (pixel tile, block), block fastest
    if (d->out_mode != CB_OUT_PF && d->out_mode != CB_OUT_PS) return CB_ERR_ARG;
    if (d->out_lo_off != 0 || d->res_lo_off != 0) return CB_ERR_ARG;
    if (p.rows_total >= (1L << 31) - 4 * TP) return CB_ERR_ARG;
    for (int i = 0; i < d->n_ksteps; ++i) {
        const cb_kstep& s = d->ksteps[i];
        if (s.a_sel > 1 || d->a_ptr[s.a_sel] == nullptr) return CB_ERR_ARG;
        if (s.col % 8 || s.col + 64 > d->a_pitch[s.a_sel]) return CB_ERR_ARG;
        if (s.w_k % 8 || s.w_k < 0 || s.w_k + 64 > d->w_k_total) return CB_ERR_ARG;
    }
    CUtensorMap ta0, ta1, tw;
    rc = make_tmap(&ta0, d->a_ptr[0], d->a_rows[0], d->a_pitch[0], d->a_pitch[0], TP);
    if (rc) return rc;
    if (d->a_ptr[1]) {
        rc = make_tmap(&ta1, d->a_ptr[1], d->a_rows[1], d->a_pitch[1], d->a_pitch[1], TP);
        if (rc) return rc;
    } else {
        ta1 = ta0;
    }
    rc = make_tmap(&tw, d->w_ptr, d->w_rows, d->w_k_total, d->w_k_total, 128);
    if (rc) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_tct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TctCfg::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = (int)((p.rows_total + TP - 1) / TP);
    int grid = m_tiles * n_cblk;
    const int cap = max_ctas > 0 ? max_ctas : sms;
    if (grid > cap) grid = cap;
    cudaError_t le = launch_pdl(conv_gemm_tct_kernel, dim3(grid), dim3(TC_THREADS), TctCfg::SMEM_BYTES, (cudaStream_t)stream,
                                ta0, ta1, tw, p, m_tiles, n_cblk);
    return le == cudaSuccess ? CB_OK : (int)le;
}

/* Halo / resident-weight tensor-core path for Cout = 64 layers (see conv_gemm_halo64_kernel). */
extern "C" int cb_conv_gemm_halo(const cb_conv_desc* d, int max_ctas, void* stream) {
    using namespace cb;
    if (!d) return CB_ERR_ARG;
    static thread_local ConvParams p;
    int rc = fill_params(d, p);
    if (rc) return rc;
    if (d->n_total != 64 || d->cout_mod != 64 || d->w_rows < 64) return CB_ERR_ARG;
    if (d->out_mode != CB_OUT_PF && d->out_mode != CB_OUT_PS) return CB_ERR_ARG;
    if (d->out_lo_off != 0 || d->res_lo_off != 0) return CB_ERR_ARG;
    if (d->n_ksteps > Halo64Cfg::MAX_KSTEPS) return CB_ERR_ARG;
    for (int i = 0; i < d->n_ksteps; ++i) {
        const cb_kstep& s = d->ksteps[i];
        if (s.a_sel > 1 || d->a_ptr[s.a_sel] == nullptr) return CB_ERR_ARG;
        if (s.col % 8 || s.col + 64 > d->a_pitch[s.a_sel]) return CB_ERR_ARG;
        if (s.w_k % 8 || s.w_k < 0 || s.w_k + 64 > d->w_k_total) return CB_ERR_ARG;
    }
    HaloItems items;
    if (build_halo_items(d, items)) return CB_ERR_ARG;
    CUtensorMap ta0, ta1, tw;
    rc = make_tmap(&ta0, d->a_ptr[0], d->a_rows[0], d->a_pitch[0], d->a_pitch[0], HALO_ROWS);
    if (rc) return rc;
    if (d->a_ptr[1]) {
        rc = make_tmap(&ta1, d->a_ptr[1], d->a_rows[1], d->a_pitch[1], d->a_pitch[1], HALO_ROWS);
        if (rc) return rc;
    } else {
        ta1 = ta0;
    }
    rc = make_tmap(&tw, d->w_ptr, d->w_rows, d->w_k_total, d->w_k_total, 64);
    if (rc) return rc;
    CUtensorMap tout = tw;
    int use_tma = 0;
    const bool no_tma_store = cb::opt_get(CB_OPT_TMA_STORE) == 0;
    if (d->out_mode == CB_OUT_PF && !no_tma_store && p.pad == 1 && p.out_Hp == p.Hp && p.out_Wp == p.Wp) {
        rc = make_tmap(&tout, d->out, p.rows_total, d->out_pitch, d->out_pitch, BM, 32);
        if (rc) return rc;
        use_tma = 1;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_halo64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Halo64Cfg::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = (int)((p.rows_total + BM - 1) / BM);
    int grid = m_tiles;
    const int cap = max_ctas > 0 ? max_ctas : sms;
    if (grid > cap) grid = cap;
    if (!use_tma && epi_direct_ok(d)) use_tma = 2;
    const int bo_mode = cb::opt_get(CB_OPT_HALO_BO);
    cudaError_t le = launch_pdl(conv_gemm_halo64_kernel, dim3(grid), dim3(TC_THREADS), Halo64Cfg::SMEM_BYTES,
                                (cudaStream_t)stream, ta0, ta1, tw, tout, p, items, m_tiles, bo_mode, use_tma);
    return le == cudaSuccess ? CB_OK : (int)le;
}

/* Channel-major path with halo boxes (see conv_gemm_tct_halo_kernel); same envelope as cb_conv_gemm_t. */
extern "C" int cb_conv_gemm_t_halo(const cb_conv_desc* d, int max_ctas, void* stream) {
    using namespace cb;
    if (!d) return CB_ERR_ARG;
    static thread_local ConvParams p;
    int rc = fill_params(d, p);
    if (rc) return rc;
    if ((d->n_total != 128 && d->n_total != 256) || d->cout_mod != d->n_total || d->w_rows < d->n_total) return CB_ERR_ARG;
    const int n_cblk = d->n_total / 128;                      // 128-channel blocks: items = (pixel tile, block), block fastest
    if (d->out_mode != CB_OUT_PF && d->out_mode != CB_OUT_PS) return CB_ERR_ARG;
    if (d->out_lo_off != 0 || d->res_lo_off != 0) return CB_ERR_ARG;
    if (p.rows_total >= (1L << 31) - 4 * TP) return CB_ERR_ARG;
    for (int i = 0; i < d->n_ksteps; ++i) {
        const cb_kstep& s = d->ksteps[i];
        if (s.a_sel > 1 || d->a_ptr[s.a_sel] == nullptr) return CB_ERR_ARG;
        if (s.col % 8 || s.col + 64 > d->a_pitch[s.a_sel]) return CB_ERR_ARG;
        if (s.w_k % 8 || s.w_k < 0 || s.w_k + 64 > d->w_k_total) return CB_ERR_ARG;
    }
    HaloItems items;
    if (build_halo_items(d, items)) return CB_ERR_ARG;
    CUtensorMap ta0, ta1, ta0t, ta1t, tw;
    rc = make_tmap(&ta0, d->a_ptr[0], d->a_rows[0], d->a_pitch[0], d->a_pitch[0], TP);
    if (rc) return rc;
    rc = make_tmap(&ta0t, d->a_ptr[0], d->a_rows[0], d->a_pitch[0], d->a_pitch[0], 2);
    if (rc) return rc;
    if (d->a_ptr[1]) {
        rc = make_tmap(&ta1, d->a_ptr[1], d->a_rows[1], d->a_pitch[1], d->a_pitch[1], TP);
        if (rc) return rc;
        rc = make_tmap(&ta1t, d->a_ptr[1], d->a_rows[1], d->a_pitch[1], d->a_pitch[1], 2);
        if (rc) return rc;
    } else {
        ta1 = ta0;
        ta1t = ta0t;
    }
    rc = make_tmap(&tw, d->w_ptr, d->w_rows, d->w_k_total, d->w_k_total, 128);
    if (rc) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_gemm_tct_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TctHaloCfg::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int m_tiles = (int)((p.rows_total + TP - 1) / TP);
    int grid = m_tiles * n_cblk;
    const int cap = max_ctas > 0 ? max_ctas : sms;
    if (grid > cap) grid = cap;
    const int bo_mode = cb::opt_get(CB_OPT_HALO_BO);
    cudaError_t le = launch_pdl(conv_gemm_tct_halo_kernel, dim3(grid), dim3(TC_THREADS), TctHaloCfg::SMEM_BYTES,
                                (cudaStream_t)stream, ta0, ta1, ta0t, ta1t, tw, p, items, m_tiles, n_cblk, bo_mode);
    return le == cudaSuccess ? CB_OK : (int)le;
}
