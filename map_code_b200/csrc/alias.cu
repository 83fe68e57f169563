// K5: Vose alias sampler.  map_alias_build is HOST code (replaces the 35 s Python loop of
// code/nce/alias_multinomial.py:40-73, bit-identical tables); map_alias_draw_philox replaces AliasMultinomial.draw
// (code/nce/alias_multinomial.py:81-97: random_ -> 2 index -> bernoulli -> select, ~8 launches) with one kernel.
#include <vector>

#include "common.cuh"

extern "C" int map_alias_build(const float* probs, int64_t V, float* out_prob, int64_t* out_alias) {
    MAP_REQUIRE(probs && out_prob && out_alias && V > 0, "map_alias_build: bad argument");
    std::vector<int64_t> smaller, larger;
    smaller.reserve((size_t)V);
    larger.reserve((size_t)V);
    const float scale = (float)V;  // K * prob evaluated in float32 like the reference's 0-dim tensor product
    for (int64_t i = 0; i < V; ++i) {
        volatile float q = scale * probs[i];
        out_prob[i] = q;
        out_alias[i] = 0;
        (q < 1.0f ? smaller : larger).push_back(i);
    }
    while (!smaller.empty() && !larger.empty()) {
        const int64_t s = smaller.back();
        smaller.pop_back();
        const int64_t l = larger.back();
        larger.pop_back();
        out_alias[s] = l;
        volatile float d = out_prob[l] - 1.0f;  // two separately rounded float32 ops (volatile: no contraction)
        volatile float q = d + out_prob[s];
        out_prob[l] = q;
        (q < 1.0f ? smaller : larger).push_back(l);
    }
    for (int64_t i : smaller) out_prob[i] = 1.0f;
    for (int64_t i : larger) out_prob[i] = 1.0f;
    return MAP_OK;
}

namespace mapb {
__global__ void __launch_bounds__(256) alias_draw_kernel(const float* __restrict__ prob, const int64_t* __restrict__ alias,
                                                         int64_t V, uint64_t seed, uint64_t offset, int64_t elem0, int64_t n,
                                                         const int64_t* __restrict__ step_dev, int64_t* __restrict__ out) {
    if (step_dev != nullptr) offset += 8ull * (uint64_t)(*step_dev);  // STREAMS_PER_STEP, see mask.cu
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const Philox4 r = philox_elem(seed, offset, (uint64_t)(elem0 + e));
        const int64_t kk = (int64_t)bounded64(r, (uint64_t)V);
        const float u = uniform24(r.w2);
        const float p = __ldg(prob + kk);
        out[e] = (u < p) ? kk : __ldg(alias + kk);
    }
}
}  // namespace mapb

extern "C" int map_alias_draw_philox(const float* prob, const int64_t* alias, int64_t V, uint64_t seed, uint64_t offset,
                                     int64_t elem0, int64_t n, const int64_t* step_dev, int64_t* out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(prob && alias && out && V > 0 && n >= 0 && elem0 >= 0, "map_alias_draw_philox: bad argument");
    if (n == 0) return MAP_OK;
    int64_t blocks = ceil_div(n, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    alias_draw_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(prob, alias, V, seed, offset, elem0, n, step_dev, out);
    return check_launch("map_alias_draw_philox");
}
