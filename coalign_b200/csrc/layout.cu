// Layout helpers: dense NCHW float32 <-> PF / PS bf16 (tests, debugging, interop) + library probes.
#include "common.cuh"
#include "../../include/coalign_b200.h"

namespace cb {

__device__ __forceinline__ long layout_row(int ps, int n_img, int H, int W, int n, int h, int w) {
    if (!ps) return ((long)n * (H + 2) + h + 1) * (W + 2) + w + 1;
    const int Hq = (H + 1) / 2 + 2, Wq = (W + 1) / 2 + 2;
    const int ph = (h & 1) * 2 + (w & 1);
    return (long)ph * n_img * Hq * Wq + ((long)n * Hq + (h >> 1) + 1) * Wq + (w >> 1) + 1;
}

__global__ void nchw_to_layout_kernel(const float* __restrict__ src, int N, int C, int H, int W, int ps,
                                      __nv_bfloat16* __restrict__ dst, long lo_off) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over (n,h,w,c), c fastest
    const long total = (long)N * H * W * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    long t = i / C;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const float v = src[(((long)n * C + c) * H + h) * W + w];
    const long row = layout_row(ps, N, H, W, n, h, w);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[row * C + c] = hi;
    if (lo_off) dst[lo_off + row * C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void layout_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, long lo_off, int ps, int N, int C, int H,
                                      int W, int pitch, int ch_off, float* __restrict__ dst) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over (n,c,h,w), w fastest
    const long total = (long)N * C * H * W;
    if (i >= total) return;
    const int w = (int)(i % W);
    long t = i / W;
    const int h = (int)(t % H); t /= H;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const long row = layout_row(ps, N, H, W, n, h, w);
    float v = __bfloat162float(src[row * pitch + ch_off + c]);
    if (lo_off) v += __bfloat162float(src[lo_off + row * pitch + ch_off + c]);
    dst[i] = v;
}

// PS -> PF copy of (n, H, W, C) bf16 maps, 16 bytes per thread (exact; used where a stride-2 consumer wants the phase-
// split layout and a second consumer wants the padded-flat one, e.g. the per-level outputs of the single-agent model).
__global__ void ps_to_pf_kernel(const uint4* __restrict__ src, int n_cap, int n, int H, int W, int c8,
                                uint4* __restrict__ dst) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over (n, h, w, c8)
    const long total = (long)n * H * W * c8;
    if (i >= total) return;
    const int c = (int)(i % c8);
    long t = i / c8;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int a = (int)(t / H);
    dst[layout_row(0, n, H, W, a, h, w) * c8 + c] = __ldg(src + layout_row(1, n_cap, H, W, a, h, w) * c8 + c);
}

__global__ void nchw_to_ps_pad_kernel(const float* __restrict__ src, int N, int C, int H, int W, int pad, int n_cap,
                                      __nv_bfloat16* __restrict__ dst, long lo_off) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // over (n,h,w,c), c fastest
    const long total = (long)N * H * W * C;
    if (i >= total) return;
    const int c = (int)(i % C);
    long t = i / C;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const float v = src[(((long)n * C + c) * H + h) * W + w];
    const int Hq = (H + 1) / 2 + 2 * pad, Wq = (W + 1) / 2 + 2 * pad;
    const int ph = (h & 1) * 2 + (w & 1);
    const long row = (long)ph * n_cap * Hq * Wq + ((long)n * Hq + (h >> 1) + pad) * Wq + (w >> 1) + pad;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    dst[row * C + c] = hi;
    if (lo_off) dst[lo_off + row * C + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// 8 channels (16 bytes) per thread; align_corners=True: source coordinate = dst * (h-1)/(scale*h-1)
__global__ void __launch_bounds__(256) upsample_concat_kernel(const __nv_bfloat16* __restrict__ src, long src_lo_off, int n,
                                                              int h, int w, int c, int scale, __nv_bfloat16* __restrict__ dst,
                                                              long dst_lo_off, int dst_pitch, int dst_ch_off) {
    pdl_wait();
    const int c8 = c >> 3, H2 = h * scale, W2 = w * scale;
    const long total = (long)n * H2 * W2 * c8;
    const float ry = H2 > 1 ? (float)(h - 1) / (float)(H2 - 1) : 0.f, rx = W2 > 1 ? (float)(w - 1) / (float)(W2 - 1) : 0.f;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
        const int ch = (int)(i % c8) * 8;
        long t = i / c8;
        const int x = (int)(t % W2); t /= W2;
        const int y = (int)(t % H2);
        const int a = (int)(t / H2);
        float v[8];
        auto ld = [&](int yy, int xx, float (&o)[8]) {
            const long off = (((long)a * (h + 2) + yy + 1) * (w + 2) + xx + 1) * c + ch;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + off));
            o[0] = bf16_lo(u.x); o[1] = bf16_hi(u.x); o[2] = bf16_lo(u.y); o[3] = bf16_hi(u.y);
            o[4] = bf16_lo(u.z); o[5] = bf16_hi(u.z); o[6] = bf16_lo(u.w); o[7] = bf16_hi(u.w);
            if (src_lo_off) {
                const uint4 l = __ldg(reinterpret_cast<const uint4*>(src + src_lo_off + off));
                o[0] += bf16_lo(l.x); o[1] += bf16_hi(l.x); o[2] += bf16_lo(l.y); o[3] += bf16_hi(l.y);
                o[4] += bf16_lo(l.z); o[5] += bf16_hi(l.z); o[6] += bf16_lo(l.w); o[7] += bf16_hi(l.w);
            }
        };
        if (scale == 1) {
            ld(y, x, v);
        } else {
            // same arithmetic as ATen's upsample_bilinear2d (align_corners=True): fp32 source index, floor, lambda weights
            const float sy = ry * (float)y, sx = rx * (float)x;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
            const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
            float p00[8], p01[8], p10[8], p11[8];
            ld(y0, x0, p00); ld(y0, x1, p01); ld(y1, x0, p10); ld(y1, x1, p11);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = hy * (hx * p00[j] + lx * p01[j]) + ly * (hx * p10[j] + lx * p11[j]);
        }
        const long o = (((long)a * (H2 + 2) + y + 1) * (W2 + 2) + x + 1) * dst_pitch + dst_ch_off + ch;
        uint32_t hi[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) hi[e] = pack_bf16(v[2 * e], v[2 * e + 1]);
        *reinterpret_cast<uint4*>(dst + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (dst_lo_off) {
            uint32_t lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) lo[e] = pack_bf16(v[2 * e] - bf16_lo(hi[e]), v[2 * e + 1] - bf16_hi(hi[e]));
            *reinterpret_cast<uint4*>(dst + dst_lo_off + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

}  // namespace cb

extern "C" int cb_nchw_to_ps_pad(const float* src, int n, int c, int h, int w, int pad, int n_cap, void* dst, int64_t lo_off,
                                 void* stream) {
    const long total = (long)n * c * h * w;
    if (!src || !dst || total <= 0 || pad < 1 || pad > 3 || n > n_cap) return CB_ERR_ARG;
    cb::nchw_to_ps_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, n, c, h, w, pad, n_cap, (__nv_bfloat16*)dst, (long)lo_off);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_upsample_concat(const void* src_pf, int64_t src_lo_off, int n, int h, int w, int c, int scale, void* dst_pf,
                                  int64_t dst_lo_off, int dst_pitch, int dst_ch_off, void* stream) {
    if (!src_pf || !dst_pf || n < 1 || h < 1 || w < 1 || c < 8 || c % 8 || (scale != 1 && scale != 2)) return CB_ERR_ARG;
    if (dst_pitch % 8 || dst_ch_off % 8 || dst_ch_off + c > dst_pitch) return CB_ERR_ARG;
    if ((src_lo_off != 0) != (dst_lo_off != 0)) return CB_ERR_ARG;
    const long total = (long)n * h * scale * w * scale * (c / 8);
    long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaError_t e = cb::launch_pdl(cb::upsample_concat_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                                   (const __nv_bfloat16*)src_pf, (long)src_lo_off, n, h, w, c, scale, (__nv_bfloat16*)dst_pf,
                                   (long)dst_lo_off, dst_pitch, dst_ch_off);
    return e == cudaSuccess ? CB_OK : (int)e;
}

extern "C" int cb_ps_to_pf(const void* src_ps, int64_t src_lo_off, int n_cap, int n, int h, int w, int c, void* dst_pf,
                           int64_t dst_lo_off, void* stream) {
    if (!src_ps || !dst_pf || n < 1 || n > n_cap || h < 1 || w < 1 || c < 8 || c % 8) return CB_ERR_ARG;
    if ((src_lo_off != 0) != (dst_lo_off != 0) || src_lo_off % 8 || dst_lo_off % 8) return CB_ERR_ARG;
    const long total = (long)n * h * w * (c / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    cb::ps_to_pf_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src_ps, n_cap, n, h, w, c / 8, (uint4*)dst_pf);
    CB_CHECK_LAUNCH();
    if (src_lo_off != 0) {
        cb::ps_to_pf_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
            (const uint4*)((const __nv_bfloat16*)src_ps + src_lo_off), n_cap, n, h, w, c / 8,
            (uint4*)((__nv_bfloat16*)dst_pf + dst_lo_off));
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

extern "C" int cb_version(void) { return 100; }

extern "C" int cb_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -3;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return (major == 10 && minor == 0) ? 0 : -4;
}

extern "C" int cb_nchw_to_layout(const float* src, int n, int c, int h, int w, int to_ps, void* dst, int64_t lo_off,
                                 void* stream) {
    const long total = (long)n * c * h * w;
    if (total <= 0) return CB_ERR_ARG;
    cb::nchw_to_layout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        src, n, c, h, w, to_ps, (__nv_bfloat16*)dst, (long)lo_off);
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_layout_to_nchw(const void* src, int64_t lo_off, int from_ps, int n, int c, int h, int w, int pitch,
                                 int ch_off, float* dst, void* stream) {
    const long total = (long)n * c * h * w;
    if (total <= 0) return CB_ERR_ARG;
    cb::layout_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, (long)lo_off, from_ps, n, c, h, w, pitch, ch_off, dst);
    CB_CHECK_LAUNCH();
    return CB_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// option table (include/coalign_b200.h: cb_set_option)
// ---------------------------------------------------------------------------------------------------------------
#include <atomic>
namespace cb {
static std::atomic<int> g_opts[CB_OPT_COUNT] = {{0}, {1}, {0}, {0}, {9}, {0}, {0}, {0}};
int opt_get(int option) {
    if (option < 0 || option >= CB_OPT_COUNT) return 0;
    return g_opts[option].load(std::memory_order_relaxed);
}
}  // namespace cb

extern "C" int cb_set_option(int option, int value) {
    if (option < 0 || option >= CB_OPT_COUNT) return CB_ERR_ARG;
    return cb::g_opts[option].exchange(value, std::memory_order_relaxed);
}
extern "C" int cb_get_option(int option) {
    if (option < 0 || option >= CB_OPT_COUNT) return CB_ERR_ARG;
    return cb::opt_get(option);
}
