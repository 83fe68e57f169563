// Exact-fp32 GEMM on CUDA cores with the fused epilogues of include/map_b200.h (MAP_EPI_*).
// Role: (1) skinny outputs that cannot fill a tensor-core tile (pred_rfd.2: N = num_fields, fc_out: N = 1);
//       (2) the fp32 cross-check of the TF32 tcgen05 kernel in tests.  The big GEMMs of the step run on
//       map_gemm_tf32_tcgen05 (gemm_tcgen05.cu).
#include "common.cuh"

namespace mapb {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float apply_epilogue(const map_gemm_args& a, float acc, int m, int n) {
    switch (a.epilogue) {
        case MAP_EPI_BIAS: return acc + a.bias[n];
        case MAP_EPI_BIAS_RELU: return fmaxf(acc + a.bias[n], 0.f);
        case MAP_EPI_CROSS: {
            const float u = acc + a.bias[n];
            a.aux_out[(int64_t)m * a.ld_aux_out + n] = u;
            return a.aux0[(int64_t)m * a.ld_aux0 + n] + a.aux1[(int64_t)m * a.ld_aux1 + n] * u;
        }
        case MAP_EPI_MUL_RELUMASK: return (a.aux0[(int64_t)m * a.ld_aux0 + n] > 0.f) ? acc : 0.f;
        case MAP_EPI_ADD: return acc + a.aux0[(int64_t)m * a.ld_aux0 + n];
        case MAP_EPI_ADD_MUL: {
            const float s = acc + a.aux0[(int64_t)m * a.ld_aux0 + n];
            a.aux_out[(int64_t)m * a.ld_aux_out + n] = s;
            return s * a.aux1[(int64_t)m * a.ld_aux1 + n];
        }
        case MAP_EPI_CROSS_BWD: {  // every (m, n) belongs to exactly one thread: the accumulation is a plain read-modify-write
            const float g = acc + (a.aux0 ? a.aux0[(int64_t)m * a.ld_aux0 + n] : 0.f);
            if (a.aux_out) a.aux_out[(int64_t)m * a.ld_aux_out + n] = g;
            float* d = a.acc_out + (int64_t)m * a.ld_acc_out + n;
            const float t = g * a.aux2[(int64_t)m * a.ld_aux2 + n];
            *d = a.acc_accumulate ? *d + t : t;
            return g * a.aux1[(int64_t)m * a.ld_aux1 + n];
        }
        case MAP_EPI_ADD3:
            return acc + a.aux0[(int64_t)m * a.ld_aux0 + n] + a.aux1[(int64_t)m * a.ld_aux1 + n] +
                   (a.aux2 ? a.aux2[(int64_t)m * a.ld_aux2 + n] : 0.f);
        default: return acc;
    }
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const map_gemm_args a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m, k;
            if (!a.trans_a) { k = t & 15; m = (t >> 4) + 16 * i; } else { m = t & 63; k = (t >> 6) + 4 * i; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < a.M && gk < a.K) v = a.trans_a ? a.A[(int64_t)gk * a.lda + gm] : a.A[(int64_t)gm * a.lda + gk];
            As[k][m] = v;
            int n;
            if (!a.trans_b) { k = t & 15; n = (t >> 4) + 16 * i; } else { n = t & 63; k = (t >> 6) + 4 * i; }
            const int gn = n0 + n;
            const int gk2 = k0 + k;
            v = 0.f;
            if (gn < a.N && gk2 < a.K) v = a.trans_b ? a.B[(int64_t)gk2 * a.ldb + gn] : a.B[(int64_t)gn * a.ldb + gk2];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }
    float csum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) {
                const float out = apply_epilogue(a, acc[i][j], m, n);
                a.C[(int64_t)m * a.ldc + n] = out;
                csum[j] += out;
            }
        }
    }
    if (a.colsum_out != nullptr) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) atomicAdd(a.colsum_out + n, csum[j]);
        }
    }
}

static_assert(sizeof(map_gemm_args) == 176, "map_gemm_args layout is part of the C ABI (mirrored by ctypes in _lib.py)");

int validate_gemm_args(const map_gemm_args* g, const char* who) {
    MAP_REQUIRE(g != nullptr, "%s: null args", who);
    MAP_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "%s: bad shape M=%d N=%d K=%d", who, g->M, g->N, g->K);
    MAP_REQUIRE(g->A && g->B && g->C, "%s: null matrix pointer", who);
    MAP_REQUIRE(g->lda >= (g->trans_a ? g->M : g->K) && g->ldb >= (g->trans_b ? g->N : g->K) && g->ldc >= g->N,
                "%s: leading dimension too small", who);
    switch (g->epilogue) {
        case MAP_EPI_NONE: break;
        case MAP_EPI_BIAS:
        case MAP_EPI_BIAS_RELU: MAP_REQUIRE(g->bias, "%s: epilogue needs bias", who); break;
        case MAP_EPI_CROSS: MAP_REQUIRE(g->bias && g->aux0 && g->aux1 && g->aux_out, "%s: EPI_CROSS needs bias, aux0 (Xi), aux1 (X0), aux_out (U)", who); break;
        case MAP_EPI_MUL_RELUMASK:
        case MAP_EPI_ADD: MAP_REQUIRE(g->aux0, "%s: epilogue needs aux0", who); break;
        case MAP_EPI_ADD_MUL: MAP_REQUIRE(g->aux0 && g->aux1 && g->aux_out, "%s: EPI_ADD_MUL needs aux0, aux1, aux_out", who); break;
        case MAP_EPI_CROSS_BWD: MAP_REQUIRE(g->aux1 && g->aux2 && g->acc_out, "%s: EPI_CROSS_BWD needs aux1 (X0), aux2 (U), acc_out", who); break;
        case MAP_EPI_ADD3: MAP_REQUIRE(g->aux0 && g->aux1, "%s: EPI_ADD3 needs aux0 and aux1", who); break;
        default: set_error("%s: unknown epilogue %d", who, g->epilogue); return MAP_EUNSUPPORTED;
    }
    return MAP_OK;
}

}  // namespace mapb

extern "C" int map_gemm_f32_simt(const map_gemm_args* args, map_stream_t stream) {
    using namespace mapb;
    const int rc = validate_gemm_args(args, "map_gemm_f32_simt");
    if (rc != MAP_OK) return rc;
    dim3 grid((unsigned)ceil_div(args->N, BN), (unsigned)ceil_div(args->M, BM));
    gemm_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(*args);
    return check_launch("map_gemm_f32_simt");
}
