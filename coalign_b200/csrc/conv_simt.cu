// SIMT evaluation of a cb_conv_desc (fp32 accumulate over the same bf16 operands, same K-step table,
// same epilogue).  Validation kernel for conv_tc.cu - tests only, never on the model path.
#include "conv_common.cuh"

namespace cb {

struct SimtSrc {
    const __nv_bfloat16* a[2];
    long a_rows[2];
    int a_pitch[2];
    const __nv_bfloat16* w;
    int w_k_total;
};

// one thread: one GEMM row x 32 output columns
__global__ void __launch_bounds__(128)
conv_gemm_simt_kernel(const __grid_constant__ ConvParams p, const SimtSrc s, int chunks_per_row, int block_n) {
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long q = gid / chunks_per_row;
    const int col0 = (int)(gid - q * chunks_per_row) * 32;
    if (q >= p.rows_total) return;
    const int n0 = (col0 / block_n) * block_n;
    const RowDest dst = decode_row(p, q, n0);
    if (dst.row < 0) return;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    for (int ks = 0; ks < p.n_ksteps; ++ks) {
        const cb_kstep st = p.ksteps[ks];
        const long row = q + st.row_off;
        if (row < 0 || row >= s.a_rows[st.a_sel]) continue;        // TMA zero fill
        const __nv_bfloat16* ar = s.a[st.a_sel] + row * (long)s.a_pitch[st.a_sel] + st.col;
        for (int kk = 0; kk < 64; ++kk) {
            const float a = __bfloat162float(ar[kk]);
            if (a == 0.f) continue;
            const __nv_bfloat16* wr = s.w + (long)col0 * s.w_k_total + st.w_k + kk;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, __bfloat162float(wr[(long)j * s.w_k_total]), acc[j]);
        }
    }
    epilogue_chunk(p, dst, q, col0, acc);
}

}  // namespace cb

extern "C" int cb_conv_gemm_simt(const cb_conv_desc* d, void* stream) {
    using namespace cb;
    if (!d) return CB_ERR_ARG;
    static thread_local ConvParams p;
    int rc = fill_params(d, p);
    if (rc) return rc;
    SimtSrc s;
    for (int i = 0; i < 2; ++i) {
        s.a[i] = (const __nv_bfloat16*)d->a_ptr[i];
        s.a_rows[i] = d->a_rows[i];
        s.a_pitch[i] = d->a_pitch[i];
    }
    s.w = (const __nv_bfloat16*)d->w_ptr;
    s.w_k_total = d->w_k_total;
    const int chunks = d->n_total / 32;
    const long total = p.rows_total * chunks;
    const int threads = 128;
    const long blocks = (total + threads - 1) / threads;
    conv_gemm_simt_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(p, s, chunks, d->block_n);
    CB_CHECK_LAUNCH();
    return CB_OK;
}
