// K10 (BCE-with-logits heads), K11 (DeepFM FM+LR term), deterministic reductions and the small elementwise stages of
// the CrossNet backward.  All HBM-bound streaming kernels: 16-byte vectors where the layout allows, grids capped at a
// multiple of the SM count, two-stage (partials -> one CTA) reductions so results do not depend on atomics order.
#include "common.cuh"

namespace mapb {

constexpr int kRedBlocks = kNumSMs * 2;

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) t = sh[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

__global__ void __launch_bounds__(256) reduce_partial_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
    __shared__ float sh[8];
    float a = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a += x[i];
    a = block_sum_256(a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = a;
}
__global__ void __launch_bounds__(256) reduce_final_kernel(const float* __restrict__ partial, int nparts, float scale, float* out) {
    __shared__ float sh[8];
    float a = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 256) a += partial[i];
    a = block_sum_256(a, sh);
    if (threadIdx.x == 0) out[0] = a * scale;
}

// BCE: partial[block*3 + {0,1,2}] = {sum loss, #correct, sum labels}
__global__ void __launch_bounds__(256) bce_partial_kernel(const float* __restrict__ z, const float* __restrict__ y, int64_t n,
                                                          float inv_n, float* __restrict__ dz, float* __restrict__ partial) {
    __shared__ float sh[8];
    float ls = 0.f, cs = 0.f, ps = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float zi = z[i], yi = y[i];
        ls += softplusf(zi) - yi * zi;  // max(z,0) - z*y + log1p(exp(-|z|))
        const float s = sigmoidf(zi);
        cs += (((s > 0.5f) ? 1.f : 0.f) == yi) ? 1.f : 0.f;  // models.py:83
        ps += yi;
        if (dz != nullptr) dz[i] = (s - yi) * inv_n;
    }
    ls = block_sum_256(ls, sh);
    cs = block_sum_256(cs, sh);
    ps = block_sum_256(ps, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x * 3 + 0] = ls;
        partial[blockIdx.x * 3 + 1] = cs;
        partial[blockIdx.x * 3 + 2] = ps;
    }
}
__global__ void __launch_bounds__(256) bce_final_kernel(const float* __restrict__ partial, int nparts, float inv_n, float n_f,
                                                        float* __restrict__ stats) {
    __shared__ float sh[8];
    float a[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < nparts; i += 256) {
        a[0] += partial[i * 3 + 0];
        a[1] += partial[i * 3 + 1];
        a[2] += partial[i * 3 + 2];
    }
    for (int k = 0; k < 3; ++k) a[k] = block_sum_256(a[k], sh);
    if (threadIdx.x == 0) {
        stats[0] = a[0] * inv_n;
        stats[1] = a[1];
        stats[2] = a[2];
        stats[3] = n_f;
    }
}

// column sums: stage 1 — CTA (bx, by) sums rows [by*rows_per, ...) of columns bx*256 + tid
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ X, int64_t ldx, int64_t M, int N,
                                                             int rows_per, float* __restrict__ partial) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= N) return;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per;
    const int64_t r1 = (r0 + rows_per < M) ? r0 + rows_per : M;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int64_t r = r0;
    for (; r + 3 < r1; r += 4) {
        a0 += X[r * ldx + col];
        a1 += X[(r + 1) * ldx + col];
        a2 += X[(r + 2) * ldx + col];
        a3 += X[(r + 3) * ldx + col];
    }
    for (; r < r1; ++r) a0 += X[r * ldx + col];
    partial[(int64_t)blockIdx.y * N + col] = (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int nparts, int N, float* __restrict__ out) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= N) return;
    float a = 0.f;
    for (int p = 0; p < nparts; ++p) a += partial[(int64_t)p * N + col];
    out[col] = a;
}

// dU = G * X0 ; dX0_acc (+)= G * U     (autograd of Xi + X0 * U, code/layers.py:200)
__global__ void __launch_bounds__(256) cross_bwd_pre_kernel(const float* __restrict__ G, int64_t ldg, const float* __restrict__ X0,
                                                            int64_t ldx0, const float* __restrict__ U, int64_t ldu, int64_t M,
                                                            int N, int accumulate, float* __restrict__ dU, float* __restrict__ dX0) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int nv = N >> 2;  // host guarantees N % 4 == 0 and 16-byte aligned rows
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * nv; e += stride) {
        const int64_t m = e / nv;
        const int c = (int)(e - m * nv) * 4;
        const float4 g = *reinterpret_cast<const float4*>(G + m * ldg + c);
        const float4 x0 = *reinterpret_cast<const float4*>(X0 + m * ldx0 + c);
        const float4 u = *reinterpret_cast<const float4*>(U + m * ldu + c);
        *reinterpret_cast<float4*>(dU + m * N + c) = make_float4(g.x * x0.x, g.y * x0.y, g.z * x0.z, g.w * x0.w);
        float4 a = make_float4(g.x * u.x, g.y * u.y, g.z * u.z, g.w * u.w);
        if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dX0 + m * N + c);
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
        }
        *reinterpret_cast<float4*>(dX0 + m * N + c) = a;
    }
}

__global__ void __launch_bounds__(256) add3_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                                   const float* __restrict__ c, int64_t ldc, int64_t M, int N,
                                                   float* __restrict__ out, int64_t ldo) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * N; e += stride) {
        const int64_t m = e / N;
        const int n = (int)(e - m * N);
        float v = a[m * lda + n] + b[m * ldb + n];
        if (c != nullptr) v += c[m * ldc + n];
        out[m * ldo + n] = v;
    }
}

// dZ = dY * (Y > 0): standalone ReLU backward (the graph-captured step fuses this into the producing GEMM's epilogue)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y,
                                                       int64_t ldy, int64_t M, int N, float* __restrict__ out, int64_t ldo) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * N; e += stride) {
        const int64_t m = e / N;
        const int n = (int)(e - m * N);
        out[m * ldo + n] = (y[m * ldy + n] > 0.f) ? dy[m * lddy + n] : 0.f;
    }
}

// out[i] = x[i] * s[0]   (s lives on the device: upstream gradient of a scalar loss, no host sync)
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(const float* __restrict__ x, const float* __restrict__ s, int64_t n,
                                                              float* __restrict__ out) {
    const float sv = s[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] * sv;
}

__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, int64_t lds, int64_t M, int N,
                                                     float* __restrict__ dst, int64_t ldd) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < M * N; e += stride) {
        const int64_t m = e / N;
        const int n = (int)(e - m * N);
        dst[m * ldd + n] = src[m * lds + n];
    }
}

// out[i, :] = X[idx[i], :]  for an int64 [n_rows, F] matrix: the device-resident batcher (replaces DataLoader collate)
__global__ void __launch_bounds__(256) gather_rows_i64_kernel(const int64_t* __restrict__ X, int F, const int64_t* __restrict__ idx,
                                                              int64_t n, int64_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n * F; e += stride) {
        const int64_t r = e / F;
        out[e] = X[idx[r] * F + (e - r * F)];
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in, int64_t M, int64_t N,
                                                        float* __restrict__ out, int64_t ld_out) {
    __shared__ float tile[32][33];
    const int64_t n0 = (int64_t)blockIdx.x * 32, m0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8)
        if (m0 + r < M && n0 + tx < N) tile[r][tx] = in[(m0 + r) * ld_in + n0 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (n0 + r < N && m0 + tx < M) out[(n0 + r) * ld_out + m0 + tx] = tile[tx][r];
}

// DeepFM: out[b] = sum_f w[ids[b,f]] + bias + 0.5 * sum_d ((sum_f e)^2 - sum_f e^2).  One warp per sample.
__global__ void __launch_bounds__(256) fm_lr_fwd_kernel(const float* __restrict__ E, const int64_t* __restrict__ ids,
                                                        const float* __restrict__ w, const float* __restrict__ lr_bias, int64_t B,
                                                        int F, int D, float* __restrict__ out, int64_t ld_out) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    float fm = 0.f;
    for (int d = lane; d < D; d += 32) {
        float s = 0.f, sq = 0.f;
        for (int f = 0; f < F; ++f) {
            const float e = E[(b * F + f) * D + d];
            s += e;
            sq = fmaf(e, e, sq);
        }
        fm += s * s - sq;
    }
    float lr = 0.f;
    for (int f = lane; f < F; f += 32) lr += (ids != nullptr) ? __ldg(w + ids[b * F + f]) : __ldg(w + b * F + f);  // ids == NULL: w is per occurrence
    const float tot = warp_sum(0.5f * fm + lr);
    if (lane == 0) out[b * ld_out] = tot + lr_bias[0];
}

__global__ void __launch_bounds__(256) fm_lr_bwd_kernel(const float* __restrict__ E, const float* __restrict__ g, int64_t ld_g,
                                                        int64_t B, int F, int D, int accumulate, float* __restrict__ dE,
                                                        float* __restrict__ dw_occ) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float gb = g[b * ld_g];
    for (int d = lane; d < D; d += 32) {
        float s = 0.f;
        for (int f = 0; f < F; ++f) s += E[(b * F + f) * D + d];
        for (int f = 0; f < F; ++f) {
            const int64_t o = (b * F + f) * D + d;
            const float v = gb * (s - E[o]);
            dE[o] = accumulate ? dE[o] + v : v;
        }
    }
    if (dw_occ != nullptr)
        for (int f = lane; f < F; f += 32) dw_occ[b * F + f] = gb;
}

}  // namespace mapb

extern "C" size_t map_reduce_workspace_bytes(int64_t n) { (void)n; return (size_t)mapb::kRedBlocks * 3 * sizeof(float); }

extern "C" int map_reduce_sum_f32(const float* x, int64_t n, float scale, float* out, void* workspace, size_t workspace_bytes,
                                  map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(x && out && workspace && n >= 0, "map_reduce_sum_f32: bad argument");
    if (workspace_bytes < map_reduce_workspace_bytes(n)) { set_error("map_reduce_sum_f32: workspace too small"); return MAP_EWORKSPACE; }
    int blocks = (int)ceil_div(n > 0 ? n : 1, 256 * 4);
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    float* partial = static_cast<float*>(workspace);
    reduce_partial_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, n, partial);
    reduce_final_kernel<<<1, 256, 0, as_stream(stream)>>>(partial, blocks, scale, out);
    return check_launch("map_reduce_sum_f32");
}

extern "C" int map_bce_logits_fwd(const float* logits, const float* labels, int64_t n, float* stats_out, float* dlogits,
                                  void* workspace, size_t workspace_bytes, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(logits && labels && stats_out && workspace && n > 0, "map_bce_logits_fwd: bad argument");
    if (workspace_bytes < map_reduce_workspace_bytes(n)) { set_error("map_bce_logits_fwd: workspace too small"); return MAP_EWORKSPACE; }
    int blocks = (int)ceil_div(n, 256 * 4);
    if (blocks > kRedBlocks) blocks = kRedBlocks;
    float* partial = static_cast<float*>(workspace);
    const float inv_n = 1.0f / (float)n;
    bce_partial_kernel<<<blocks, 256, 0, as_stream(stream)>>>(logits, labels, n, inv_n, dlogits, partial);
    bce_final_kernel<<<1, 256, 0, as_stream(stream)>>>(partial, blocks, inv_n, (float)n, stats_out);
    return check_launch("map_bce_logits_fwd");
}

namespace mapb {
static int colsum_parts(int64_t M) {
    int parts = (int)ceil_div(M, 64);
    if (parts > 64) parts = 64;
    if (parts < 1) parts = 1;
    return parts;
}
}  // namespace mapb
extern "C" size_t map_colsum_workspace_bytes(int64_t M, int N) { return (size_t)mapb::colsum_parts(M) * (size_t)N * sizeof(float); }

extern "C" int map_colsum_f32(const float* X, int64_t ldx, int64_t M, int N, float* out, void* workspace, size_t workspace_bytes,
                              map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(X && out && workspace && M > 0 && N > 0, "map_colsum_f32: bad argument");
    if (workspace_bytes < map_colsum_workspace_bytes(M, N)) { set_error("map_colsum_f32: workspace too small"); return MAP_EWORKSPACE; }
    const int parts = colsum_parts(M);
    const int rows_per = (int)ceil_div(M, parts);
    float* partial = static_cast<float*>(workspace);
    dim3 grid((unsigned)ceil_div(N, 256), (unsigned)parts);
    colsum_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, ldx, M, N, rows_per, partial);
    colsum_final_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, as_stream(stream)>>>(partial, parts, N, out);
    return check_launch("map_colsum_f32");
}

extern "C" int map_cross_bwd_pre(const float* G, int64_t ldg, const float* X0, int64_t ldx0, const float* U, int64_t ldu,
                                 int64_t M, int N, int accumulate, float* dU, float* dX0_acc, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(G && X0 && U && dU && dX0_acc && M > 0 && N > 0, "map_cross_bwd_pre: bad argument");
    MAP_REQUIRE(N % 4 == 0 && ldg % 4 == 0 && ldx0 % 4 == 0 && ldu % 4 == 0 &&
                    ((((uintptr_t)G | (uintptr_t)X0 | (uintptr_t)U | (uintptr_t)dU | (uintptr_t)dX0_acc) & 15) == 0),
                "map_cross_bwd_pre: needs N %% 4 == 0 and 16-byte aligned rows");
    int64_t blocks = ceil_div(M * (N / 4), 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cross_bwd_pre_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(G, ldg, X0, ldx0, U, ldu, M, N, accumulate, dU, dX0_acc);
    return check_launch("map_cross_bwd_pre");
}

extern "C" int map_add3_f32(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, int64_t M, int N,
                            float* out, int64_t ldo, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(a && b && out && M > 0 && N > 0, "map_add3_f32: bad argument");
    int64_t blocks = ceil_div(M * N, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    add3_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a, lda, b, ldb, c, ldc, M, N, out, ldo);
    return check_launch("map_add3_f32");
}

extern "C" int map_transpose_f32(const float* in, int64_t ld_in, int64_t M, int64_t N, float* out, int64_t ld_out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(in && out && M > 0 && N > 0, "map_transpose_f32: bad argument");
    dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32));
    transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, ld_in, M, N, out, ld_out);
    return check_launch("map_transpose_f32");
}

extern "C" int map_fm_lr_fwd(const float* feat_embed, const int64_t* ids, const float* lr_w, const float* lr_bias, int64_t B, int F,
                             int D, float* out, int64_t ld_out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(feat_embed && lr_w && lr_bias && out && B > 0 && F > 0 && D > 0, "map_fm_lr_fwd: bad argument");
    fm_lr_fwd_kernel<<<(unsigned)ceil_div(B * 32, 256), 256, 0, as_stream(stream)>>>(feat_embed, ids, lr_w, lr_bias, B, F, D, out, ld_out);
    return check_launch("map_fm_lr_fwd");
}

extern "C" int map_fm_lr_bwd(const float* feat_embed, const float* g, int64_t ld_g, int64_t B, int F, int D, int accumulate,
                             float* d_embed, float* d_w_occ, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(feat_embed && g && d_embed && B > 0 && F > 0 && D > 0, "map_fm_lr_bwd: bad argument");
    fm_lr_bwd_kernel<<<(unsigned)ceil_div(B * 32, 256), 256, 0, as_stream(stream)>>>(feat_embed, g, ld_g, B, F, D, accumulate, d_embed, d_w_occ);
    return check_launch("map_fm_lr_bwd");
}

extern "C" int map_relu_bwd_f32(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t M, int N, float* out, int64_t ldo,
                                map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(dy && y && out && M > 0 && N > 0, "map_relu_bwd_f32: bad argument");
    int64_t blocks = ceil_div(M * N, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    relu_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(dy, lddy, y, ldy, M, N, out, ldo);
    return check_launch("map_relu_bwd_f32");
}

extern "C" int map_scale_by_scalar_f32(const float* x, const float* scalar_dev, int64_t n, float* out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(x && scalar_dev && out && n >= 0, "map_scale_by_scalar_f32: bad argument");
    if (n == 0) return MAP_OK;
    int64_t blocks = ceil_div(n, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    scale_by_scalar_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, scalar_dev, n, out);
    return check_launch("map_scale_by_scalar_f32");
}

extern "C" int map_copy2d_f32(const float* src, int64_t ld_src, int64_t M, int N, float* dst, int64_t ld_dst, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(src && dst && M > 0 && N > 0 && ld_src >= N && ld_dst >= N, "map_copy2d_f32: bad argument");
    int64_t blocks = ceil_div(M * N, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    copy2d_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, ld_src, M, N, dst, ld_dst);
    return check_launch("map_copy2d_f32");
}

extern "C" int map_gather_rows_i64(const int64_t* X, int64_t n_rows, int F, const int64_t* idx, int64_t n, int64_t* out,
                                   map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(n >= 0 && F > 0 && n_rows > 0, "map_gather_rows_i64: bad shape");
    if (n == 0) return MAP_OK;
    MAP_REQUIRE(X && idx && out, "map_gather_rows_i64: null pointer");
    int64_t blocks = ceil_div(n * F, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    gather_rows_i64_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(X, F, idx, n, out);
    return check_launch("map_gather_rows_i64");
}

// ------------------------------------------------------------------------------------------------ on-device AUC (Trainer.eval)
// ROC AUC of n scores = (sum of the average ranks of the positives - P(P+1)/2) / (P * (n - P)), ties sharing their average
// rank — what sklearn.metrics.roc_auc_score computes (code/trainer.py:196) — as a composition of the K2 kernels: scores ->
// order-preserving 32-bit keys -> map_dedup_ids (sorted tie groups: unique keys, segment starts) -> map_segment_reduce_rows
// with the labels as 1-float rows (positives per tie group) -> the rank sum below.
namespace mapb {

__global__ void __launch_bounds__(256) float_sort_keys_kernel(const float* __restrict__ x, int64_t n, int64_t* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t b = __float_as_uint(x[i] + 0.0f);                 // (-0.0 + 0.0 = +0.0: the two zeros tie)
        b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);           // ascending float order == ascending unsigned order
        keys[i] = (int64_t)b;
    }
}

// out[0] += sum_u pos[u] * (seg_start[u] + seg_start[u+1] + 1) / 2   (average 1-based rank of tie group u), out[1] += sum_u pos[u]
__global__ void __launch_bounds__(256) auc_rank_sum_kernel(const int32_t* __restrict__ seg_start, const float* __restrict__ pos,
                                                           const int32_t* __restrict__ n_unique, double* __restrict__ out) {
    __shared__ double sh[2][8];
    const int64_t U = *n_unique;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double r = 0.0, p = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < U; u += stride) {
        const double c = (double)pos[u];
        if (c != 0.0) {
            r += c * (0.5 * ((double)seg_start[u] + (double)seg_start[u + 1] + 1.0));
            p += c;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        p += __shfl_xor_sync(0xffffffffu, p, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sh[0][warp] = r;
        sh[1][warp] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double rs = 0.0, ps = 0.0;
        for (int w = 0; w < 8; ++w) {
            rs += sh[0][w];
            ps += sh[1][w];
        }
        if (rs != 0.0) atomicAdd(out, rs);
        if (ps != 0.0) atomicAdd(out + 1, ps);
    }
}

}  // namespace mapb

extern "C" int map_float_sort_keys(const float* x, int64_t n, int64_t* keys, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(x && keys && n > 0, "map_float_sort_keys: bad argument");
    int64_t blocks = ceil_div(n, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    float_sort_keys_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, keys);
    return check_launch("map_float_sort_keys");
}

extern "C" int map_auc_rank_sum(const int32_t* seg_start, const float* pos_count, const int32_t* n_unique, int64_t max_unique,
                                double* out2, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(seg_start && pos_count && n_unique && out2 && max_unique > 0, "map_auc_rank_sum: bad argument");
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(out2, 0, 2 * sizeof(double), st) != cudaSuccess) {
        set_error("map_auc_rank_sum: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MAP_ECUDA;
    }
    int64_t blocks = ceil_div(max_unique, 256);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    auc_rank_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(seg_start, pos_count, n_unique, out2);
    return check_launch("map_auc_rank_sum");
}
