// Lift + splat of the camera model (sm_100a): depth soft-max (x) image features, summed into the BEV voxels their frustum
// points fall into - CamEncode.get_depth_dist + the outer product of CamEncode.forward
// (/root/reference/opencood/models/sub_modules/lss_submodule.py:60-61,134-136), LiftSplatShoot.get_geometry and voxel_pooling
// (/root/reference/opencood/models/lift_splat_shoot.py:80-169).  The (B*N, C, D, fH, fW) lifted tensor of the reference
// (472 MB per agent at the yaml's sizes), its sort by voxel rank and the cumsum trick are never materialised: one warp owns
// one image-feature pixel, keeps its C channels in registers (4 per lane and 128-channel group), walks the D depth bins of the
// ray, merges consecutive bins that land in the same voxel and adds `sum(prob) * feature` with 16-byte vector reductions into
// an fp32 channels-last BEV accumulator (512 contiguous bytes per warp-level reduction).  HBM/L2-atomic bound.
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

struct LiftGeom {
    int B, N, D, fH, fW, C;
    float dx[3], lo[3];          // voxel size, bx - dx/2
    int nx[3];
};

__device__ __forceinline__ void red_add_v4_ls(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.v4.f32.add [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// cam: [BN][24] = inverse(post_rots) (9, row major), post_trans (3), rots * inverse(intrins) (9), trans (3)
template <int G>     // 128-channel groups per pixel (C <= 128 * G)
__global__ void __launch_bounds__(256) lift_splat_kernel(const float* __restrict__ depth_logit, const float* __restrict__ feat,
                                                         const float* __restrict__ cam, const float* __restrict__ xs,
                                                         const float* __restrict__ ys, const float* __restrict__ ds,
                                                         const LiftGeom g, float* __restrict__ acc) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long warp = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long n_pix = (long)g.B * g.N * g.fH * g.fW;
    if (warp >= n_pix) return;
    const int w = (int)(warp % g.fW);
    long t = warp / g.fW;
    const int h = (int)(t % g.fH);
    const int bn = (int)(t / g.fH);
    const int b = bn / g.N;
    const long hw = (long)g.fH * g.fW, pix = (long)h * g.fW + w;
    // features of this pixel: lane owns channels 4*lane .. 4*lane+3 of every 128-channel group
    float4 f[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const int c0 = 128 * k + 4 * lane;
        f[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 < g.C) {
            const float* fp = feat + ((long)bn * g.C + c0) * hw + pix;
            f[k] = make_float4(__ldg(fp), __ldg(fp + hw), __ldg(fp + 2 * hw), __ldg(fp + 3 * hw));
        }
    }
    // soft-max over the depth bins (F.softmax: exp(x - max) / sum), bins d = lane, lane + 32
    const float* dl = depth_logit + (long)bn * g.D * hw + pix;
    float lg[2], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int d = lane + 32 * k;
        lg[k] = d < g.D ? __ldg(dl + (long)d * hw) : -INFINITY;
        mx = fmaxf(mx, lg[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float pr[2], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) { pr[k] = (lane + 32 * k) < g.D ? expf(lg[k] - mx) : 0.f; sum += pr[k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // voxel of every bin of this ray (lift_splat_shoot.py:92-105,121): float32, `.long()` truncates toward zero
    const float* cm = cam + (long)bn * 24;
    const float px = __fsub_rn(xs[w], cm[9]), py = __fsub_rn(ys[h], cm[10]);
    int vox[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int d = lane + 32 * k;
        vox[k] = -1;
        if (d < g.D) {
            const float pz = __fsub_rn(ds[d], cm[11]);
            const float q0 = cm[0] * px + cm[1] * py + cm[2] * pz;
            const float q1 = cm[3] * px + cm[4] * py + cm[5] * pz;
            const float q2 = cm[6] * px + cm[7] * py + cm[8] * pz;
            const float r0 = q0 * q2, r1 = q1 * q2;
            const float s0 = cm[12] * r0 + cm[13] * r1 + cm[14] * q2 + cm[21];
            const float s1 = cm[15] * r0 + cm[16] * r1 + cm[17] * q2 + cm[22];
            const float s2 = cm[18] * r0 + cm[19] * r1 + cm[20] * q2 + cm[23];
            const long i0 = (long)__fdiv_rn(__fsub_rn(s0, g.lo[0]), g.dx[0]);
            const long i1 = (long)__fdiv_rn(__fsub_rn(s1, g.lo[1]), g.dx[1]);
            const long i2 = (long)__fdiv_rn(__fsub_rn(s2, g.lo[2]), g.dx[2]);
            if (i0 >= 0 && i0 < g.nx[0] && i1 >= 0 && i1 < g.nx[1] && i2 >= 0 && i2 < g.nx[2])
                vox[k] = (int)((i1 * g.nx[0] + i0) * g.nx[2] + i2);           // (y, x, z): channels-last cell + z slot
        }
        pr[k] = pr[k] / sum;
    }
    // walk the ray: consecutive bins in the same voxel are merged into one reduction
    const long cell_stride = (long)g.nx[2] * g.C;                             // floats per (y, x) cell: [z][C]
    float* accb = acc + (long)b * g.nx[1] * g.nx[0] * cell_stride;
    int cur = -1;
    float wsum = 0.f;
    for (int d = 0; d <= g.D; ++d) {
        int v = -1;
        float p = 0.f;
        if (d < g.D) {
            v = __shfl_sync(0xffffffffu, vox[d >> 5], d & 31);
            p = __shfl_sync(0xffffffffu, pr[d >> 5], d & 31);
        }
        if (v != cur || d == g.D) {
            if (cur >= 0 && wsum != 0.f) {
                const int z = cur % g.nx[2];
                float* dst = accb + (long)(cur / g.nx[2]) * cell_stride + (long)z * g.C;
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const int c0 = 128 * k + 4 * lane;
                    if (c0 < g.C) red_add_v4_ls(dst + c0, wsum * f[k].x, wsum * f[k].y, wsum * f[k].z, wsum * f[k].w);
                }
            }
            cur = v;
            wsum = 0.f;
        }
        wsum += p;
    }
}

// fp32 channels-last BEV accumulator (n, H, W, C) -> bf16 PS layout with a `pad`-pixel halo (the stem's input)
__global__ void __launch_bounds__(256) nhwc_to_ps_pad_kernel(const float* __restrict__ src, int N, int C, int H, int W, int pad,
                                                             int n_cap, __nv_bfloat16* __restrict__ dst, long lo_off) {
    pdl_wait();
    const int c8 = C >> 3;
    const long total = (long)N * H * W * c8;
    const int Hq = (H + 1) / 2 + 2 * pad, Wq = (W + 1) / 2 + 2 * pad;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int ch = (int)(i % c8) * 8;
        long t = i / c8;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + (((long)n * H + h) * W + w) * C + ch));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + (((long)n * H + h) * W + w) * C + ch) + 1);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const int ph = (h & 1) * 2 + (w & 1);
        const long row = (long)ph * n_cap * Hq * Wq + ((long)n * Hq + (h >> 1) + pad) * Wq + (w >> 1) + pad;
        uint32_t hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(dst + row * C + ch) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (lo_off) {
            uint32_t lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) lo[e] = pack_bf16(v[2 * e] - bf16_lo(hi[e]), v[2 * e + 1] - bf16_hi(hi[e]));
            *reinterpret_cast<uint4*>(dst + lo_off + row * C + ch) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

}  // namespace cb

extern "C" int cb_lift_splat(const float* depth_logit, const float* feat, const float* cam_mats, const float* xs, const float* ys,
                             const float* ds, int B, int N, int D, int fH, int fW, int C, const float* dx, const float* bx,
                             const int32_t* nx, float* acc, void* stream) {
    using namespace cb;
    if (!depth_logit || !feat || !cam_mats || !xs || !ys || !ds || !dx || !bx || !nx || !acc) return CB_ERR_ARG;
    if (B < 1 || N < 1 || D < 1 || D > 64 || fH < 1 || fW < 1 || C < 4 || C % 4 || C > 256 || ((uintptr_t)acc & 15)) return CB_ERR_ARG;
    if ((long)nx[0] * nx[1] * nx[2] >= (1L << 31) || (nx[2] * C) % 4) return CB_ERR_ARG;
    LiftGeom g;
    g.B = B; g.N = N; g.D = D; g.fH = fH; g.fW = fW; g.C = C;
    for (int i = 0; i < 3; ++i) { g.dx[i] = dx[i]; g.lo[i] = bx[i] - dx[i] / 2.f; g.nx[i] = nx[i]; }
    const long n_pix = (long)B * N * fH * fW;
    const unsigned blocks = (unsigned)((n_pix + 7) / 8);
    cudaError_t e = C <= 128
        ? launch_pdl(lift_splat_kernel<1>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, depth_logit, feat, cam_mats, xs, ys, ds, g, acc)
        : launch_pdl(lift_splat_kernel<2>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, depth_logit, feat, cam_mats, xs, ys, ds, g, acc);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_nhwc_to_ps_pad(const float* src, int n, int c, int h, int w, int pad, int n_cap, void* dst, int64_t lo_off,
                                 void* stream) {
    if (!src || !dst || n < 1 || n > n_cap || c < 8 || c % 8 || pad < 1 || pad > 3 || ((uintptr_t)src & 15)) return CB_ERR_ARG;
    const long total = (long)n * h * w * (c / 8);
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaError_t e = cb::launch_pdl(cb::nhwc_to_ps_pad_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, n, c, h,
                                   w, pad, n_cap, (__nv_bfloat16*)dst, (long)lo_off);
    return e == cudaSuccess ? CB_OK : (int)e;
}
