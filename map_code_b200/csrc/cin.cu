// K17: Compressed Interaction Network of xDeepFM (reference code/layers.py:696-721).
//
// One CIN layer of the reference:  hadamard[b, h*M + m, d] = X0[b,h,d] * Xi[b,m,d]  (torch.einsum "bhd,bmd->bhmd"),
// X_next = Conv1d(F*M -> O, kernel 1)(hadamard) = W[O, F*M] . hadamard[b, :, d] + bias,  pooled[b, o] = sum_d X_next[b,o,d].
// The 1x1 convolution is a GEMM over the (sample, embedding-dim) PAIRS p = b*D + d:  Y[p, o] = sum_k Z[p, k] W[o, k] + bias[o] with
// Z[p, h*M + m] = X0[p, h] * Xi[p, m].  We keep every CIN activation pair-major ([B*D, channels], what the GEMM consumes and
// produces), so the contraction runs on the tensor-core GEMM (map_gemm_bf16s_group) with K padded to a multiple of 8 and the only
// extra kernels are the ones in this file: the [B, C, D] <-> [B*D, C] relayout of the embedding, the outer product and its
// autograd, and the sum over d (pooling) and its broadcast backward.  All exact fp32, deterministic.
#include "common.cuh"

namespace mapb {

// in [B, C, D] (row stride of a sample = C*D) -> out [(b*D + d), c] with leading dimension ld_out; or the reverse (to_pairs = 0):
// pairs [(b*D + d), c] -> out [B, C, D], optionally accumulating into out.
__global__ void __launch_bounds__(256) cin_relayout_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t B, int C, int D,
                                                           int64_t ld_pairs, int to_pairs, int accumulate) {
    extern __shared__ float tile[];   // [C][D + 1] of one sample
    const int Dp = D + 1;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (to_pairs) {
            for (int e = threadIdx.x; e < C * D; e += 256) tile[(e / D) * Dp + (e % D)] = src[b * C * D + e];
            __syncthreads();
            for (int e = threadIdx.x; e < C * D; e += 256) {
                const int d = e / C, c = e - d * C;
                dst[(b * D + d) * ld_pairs + c] = tile[c * Dp + d];
            }
        } else {
            for (int e = threadIdx.x; e < C * D; e += 256) {
                const int d = e / C, c = e - d * C;
                tile[c * Dp + d] = src[(b * D + d) * ld_pairs + c];
            }
            __syncthreads();
            for (int e = threadIdx.x; e < C * D; e += 256) {
                const float v = tile[(e / D) * Dp + (e % D)];
                float* o = dst + b * C * D + e;
                *o = accumulate ? *o + v : v;
            }
        }
    }
}

// Z[p, h*M + m] = X0[p, h] * Xi[p, m]; columns [F*M, ldz) are written as zeros (K padding of the GEMM)
__global__ void __launch_bounds__(256) cin_hadamard_fwd_kernel(const float* __restrict__ x0, int64_t ld_x0, const float* __restrict__ xi,
                                                               int64_t ld_xi, int64_t P, int F, int M, float* __restrict__ z, int64_t ldz) {
    extern __shared__ float sm[];   // x0 row [F], xi row [M]
    float* a = sm;
    float* c = sm + F;
    const int K = F * M;
    for (int64_t p = blockIdx.x; p < P; p += gridDim.x) {
        __syncthreads();
        for (int e = threadIdx.x; e < F; e += 256) a[e] = x0[p * ld_x0 + e];
        for (int e = threadIdx.x; e < M; e += 256) c[e] = xi[p * ld_xi + e];
        __syncthreads();
        float* row = z + p * ldz;
        for (int k = threadIdx.x; k < (int)ldz; k += 256) row[k] = k < K ? a[k / M] * c[k % M] : 0.f;
    }
}

// dX0[p, h] (+)= sum_m dZ[p, h*M + m] * Xi[p, m];  dXi[p, m] = sum_h dZ[p, h*M + m] * X0[p, h]
__global__ void __launch_bounds__(128) cin_hadamard_bwd_kernel(const float* __restrict__ dz, int64_t ldz, const float* __restrict__ x0,
                                                               int64_t ld_x0, const float* __restrict__ xi, int64_t ld_xi, int64_t P, int F,
                                                               int M, float* __restrict__ dx0, int64_t ld_dx0, int acc_dx0,
                                                               float* __restrict__ dxi, int64_t ld_dxi) {
    extern __shared__ float sm[];   // dz row [F*M], x0 row [F], xi row [M]
    const int K = F * M;
    float* g = sm;
    float* a = sm + K;
    float* c = a + F;
    for (int64_t p = blockIdx.x; p < P; p += gridDim.x) {
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += 128) g[k] = dz[p * ldz + k];
        for (int e = threadIdx.x; e < F; e += 128) a[e] = x0[p * ld_x0 + e];
        for (int e = threadIdx.x; e < M; e += 128) c[e] = xi[p * ld_xi + e];
        __syncthreads();
        for (int t = threadIdx.x; t < F + M; t += 128) {
            float s = 0.f;
            if (t < F) {
                for (int m = 0; m < M; ++m) s = fmaf(g[t * M + m], c[m], s);
                float* o = dx0 + p * ld_dx0 + t;
                *o = acc_dx0 ? *o + s : s;
            } else {
                const int m = t - F;
                for (int h = 0; h < F; ++h) s = fmaf(g[h * M + m], a[h], s);
                dxi[p * ld_dxi + m] = s;
            }
        }
    }
}

// pooled[b, o] = sum_d Y[b*D + d, o]   (X_i.sum(dim=-1), layers.py:719)
__global__ void __launch_bounds__(256) cin_pool_fwd_kernel(const float* __restrict__ y, int64_t ldy, int64_t B, int D, int O,
                                                           float* __restrict__ pooled, int64_t ld_pooled) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < B * O; e += stride) {
        const int64_t b = e / O;
        const int o = (int)(e - b * O);
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += y[(b * D + d) * ldy + o];
        pooled[b * ld_pooled + o] = s;
    }
}

// dY[b*D + d, o] = d_pooled[b, o] (+ d_next[b*D + d, o] if the layer feeds another one); columns [O, ld_dy) are zeroed
__global__ void __launch_bounds__(256) cin_pool_bwd_kernel(const float* __restrict__ d_pooled, int64_t ld_pooled, const float* __restrict__ d_next,
                                                           int64_t ld_next, int64_t B, int D, int O, float* __restrict__ dy, int64_t ld_dy) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = B * D * ld_dy;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t p = e / ld_dy;
        const int o = (int)(e - p * ld_dy);
        float v = 0.f;
        if (o < O) {
            v = d_pooled[(p / D) * ld_pooled + o];
            if (d_next != nullptr) v += d_next[p * ld_next + o];
        }
        dy[e] = v;
    }
}

static inline unsigned cin_grid(int64_t n, int per_sm) {
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    return (unsigned)(n < cap ? (n < 1 ? 1 : n) : cap);
}

}  // namespace mapb

using namespace mapb;

extern "C" int map_cin_relayout(const float* src, float* dst, int64_t B, int C, int D, int64_t ld_pairs, int to_pairs, int accumulate,
                                map_stream_t stream) {
    MAP_REQUIRE(src && dst && B >= 1 && C >= 1 && D >= 1 && ld_pairs >= C, "map_cin_relayout: bad argument");
    const size_t smem = (size_t)C * (D + 1) * sizeof(float);
    MAP_REQUIRE(smem <= 48 * 1024, "map_cin_relayout: C*(D+1) = %d floats exceed 48 KB of shared memory", C * (D + 1));
    cin_relayout_kernel<<<cin_grid(B, 8), 256, smem, as_stream(stream)>>>(src, dst, B, C, D, ld_pairs, to_pairs, accumulate);
    return check_launch("map_cin_relayout");
}

extern "C" int map_cin_hadamard_fwd(const float* x0, int64_t ld_x0, const float* xi, int64_t ld_xi, int64_t P, int F, int M, float* z,
                                    int64_t ldz, map_stream_t stream) {
    MAP_REQUIRE(x0 && xi && z && P >= 1 && F >= 1 && M >= 1 && ldz >= (int64_t)F * M && ld_x0 >= F && ld_xi >= M && ldz < (1ll << 31),
                "map_cin_hadamard_fwd: bad argument");
    const size_t smem = (size_t)(F + M) * sizeof(float);
    MAP_REQUIRE(smem <= 48 * 1024, "map_cin_hadamard_fwd: F + M too large");
    cin_hadamard_fwd_kernel<<<cin_grid(P, 16), 256, smem, as_stream(stream)>>>(x0, ld_x0, xi, ld_xi, P, F, M, z, ldz);
    return check_launch("map_cin_hadamard_fwd");
}

extern "C" int map_cin_hadamard_bwd(const float* dz, int64_t ldz, const float* x0, int64_t ld_x0, const float* xi, int64_t ld_xi, int64_t P,
                                    int F, int M, float* dx0, int64_t ld_dx0, int accumulate_dx0, float* dxi, int64_t ld_dxi,
                                    map_stream_t stream) {
    MAP_REQUIRE(dz && x0 && xi && dx0 && dxi && P >= 1 && F >= 1 && M >= 1 && ldz >= (int64_t)F * M && ld_x0 >= F && ld_xi >= M &&
                    ld_dx0 >= F && ld_dxi >= M,
                "map_cin_hadamard_bwd: bad argument");
    const size_t smem = ((size_t)F * M + F + M) * sizeof(float);
    MAP_REQUIRE(smem <= 48 * 1024, "map_cin_hadamard_bwd: F*M + F + M = %lld floats exceed 48 KB of shared memory", (long long)(smem / 4));
    cin_hadamard_bwd_kernel<<<cin_grid(P, 16), 128, smem, as_stream(stream)>>>(dz, ldz, x0, ld_x0, xi, ld_xi, P, F, M, dx0, ld_dx0,
                                                                                accumulate_dx0, dxi, ld_dxi);
    return check_launch("map_cin_hadamard_bwd");
}

extern "C" int map_cin_pool_fwd(const float* y, int64_t ldy, int64_t B, int D, int O, float* pooled, int64_t ld_pooled, map_stream_t stream) {
    MAP_REQUIRE(y && pooled && B >= 1 && D >= 1 && O >= 1 && ldy >= O && ld_pooled >= O, "map_cin_pool_fwd: bad argument");
    cin_pool_fwd_kernel<<<cin_grid(ceil_div(B * O, 256), 8), 256, 0, as_stream(stream)>>>(y, ldy, B, D, O, pooled, ld_pooled);
    return check_launch("map_cin_pool_fwd");
}

extern "C" int map_cin_pool_bwd(const float* d_pooled, int64_t ld_pooled, const float* d_next, int64_t ld_next, int64_t B, int D, int O,
                                float* dy, int64_t ld_dy, map_stream_t stream) {
    MAP_REQUIRE(d_pooled && dy && B >= 1 && D >= 1 && O >= 1 && ld_dy >= O && ld_pooled >= O && (d_next == nullptr || ld_next >= O),
                "map_cin_pool_bwd: bad argument");
    cin_pool_bwd_kernel<<<cin_grid(ceil_div(B * D * ld_dy, 256), 8), 256, 0, as_stream(stream)>>>(d_pooled, ld_pooled, d_next, ld_next, B, D, O,
                                                                                                  dy, ld_dy);
    return check_launch("map_cin_pool_bwd");
}
