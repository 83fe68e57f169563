// K1: embedding row gather.  HBM/L2-bound: 16-byte lanes, D/4 lanes per row, 4 independent rows in flight per thread.
// Replaces aten::embedding / index_select at code/layers.py:98, code/models.py:139, code/nce/index_linear.py:99-100.
#include "common.cuh"

namespace mapb {

constexpr int kGatherUnroll = 4;

// Flattened over 16-byte vectors: vector e belongs to row e / vpr, lane e % vpr.  Consecutive threads write consecutive
// 16-byte vectors of `out` (fully coalesced, streaming stores); the lanes of one row read one contiguous table row.
__global__ void __launch_bounds__(256) emb_gather_vec4_kernel(const float4* __restrict__ table, int64_t V, int vpr,
                                                              const int64_t* __restrict__ ids, int64_t n_vec,
                                                              float4* __restrict__ out, int32_t* oob_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; e < n_vec; e += stride * kGatherUnroll) {
        int64_t row[kGatherUnroll];
        int lane[kGatherUnroll];
        float4 val[kGatherUnroll];
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int64_t eu = e + u * stride;
            row[u] = -1;
            lane[u] = 0;
            if (eu < n_vec) {
                const int64_t r = eu / vpr;
                lane[u] = (int)(eu - r * vpr);
                row[u] = __ldg(ids + r);
            }
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row[u] >= 0 && row[u] < V) val[u] = __ldg(table + row[u] * vpr + lane[u]);
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int64_t eu = e + u * stride;
            if (eu < n_vec) {
                st_stream_f4(out + eu, val[u]);
                if (oob_flag != nullptr && (row[u] < 0 || row[u] >= V)) *oob_flag = 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) emb_gather_scalar_kernel(const float* __restrict__ table, int64_t V, int D,
                                                                const int64_t* __restrict__ ids, int64_t n_elem,
                                                                float* __restrict__ out, int32_t* oob_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += stride) {
        const int64_t r = e / D;
        const int d = (int)(e - r * D);
        const int64_t row = __ldg(ids + r);
        float v = 0.f;
        if (row >= 0 && row < V) v = __ldg(table + row * D + d);
        else if (oob_flag != nullptr) *oob_flag = 1;
        out[e] = v;
    }
}

}  // namespace mapb

extern "C" int map_emb_gather_f32(const float* table, int64_t V, int D, const int64_t* ids, int64_t n_ids, float* out,
                                  int32_t* oob_flag, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(V > 0 && D > 0 && n_ids >= 0, "map_emb_gather_f32: bad shape V=%lld D=%d n=%lld", (long long)V, D, (long long)n_ids);
    if (n_ids == 0) return MAP_OK;  // empty batch: nothing to launch (torch hands out null pointers for empty tensors)
    MAP_REQUIRE(table && ids && out, "map_emb_gather_f32: null pointer");
    const bool vec = (D % 4 == 0) && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec) {
        const int vpr = D / 4;
        const int64_t n_vec = n_ids * vpr;
        int64_t blocks = ceil_div(n_vec, 256 * kGatherUnroll);
        const int64_t cap = (int64_t)kNumSMs * 8;  // 8 resident CTAs of 256 threads per SM
        if (blocks > cap) blocks = cap;
        emb_gather_vec4_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float4*>(table), V, vpr, ids, n_vec, reinterpret_cast<float4*>(out), oob_flag);
    } else {
        const int64_t n_elem = n_ids * D;
        int64_t blocks = ceil_div(n_elem, 256);
        const int64_t cap = (int64_t)kNumSMs * 8;
        if (blocks > cap) blocks = cap;
        emb_gather_scalar_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, V, D, ids, n_elem, out, oob_flag);
    }
    return check_launch("map_emb_gather_f32");
}
