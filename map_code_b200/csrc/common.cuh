// Shared helpers for libmap_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/map_b200.h"

namespace mapb {

void set_error(const char* fmt, ...);
int check_launch(const char* what);  // returns MAP_OK or MAP_ECUDA (records cudaGetLastError text)

inline cudaStream_t as_stream(map_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define MAP_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            mapb::set_error(__VA_ARGS__);      \
            return MAP_EINVAL;                 \
        }                                      \
    } while (0)

constexpr int kNumSMs = 148;  // B200: grids are sized in multiples of this

// Row-sharded tables over NVLink peer memory (p2p.cu): base pointer of every rank's shard, passed to kernels by value.
constexpr int kMaxPeers = 8;  // one 8-GPU NVSwitch node
struct PeerTable {
    const void* p[kMaxPeers];
};
static inline int fill_peer_table(PeerTable* t, const void* const* ptrs, int R, const char* who) {
    MAP_REQUIRE(R >= 1 && R <= kMaxPeers, "%s: R=%d must be in [1, %d]", who, R, kMaxPeers);
    MAP_REQUIRE(ptrs != nullptr, "%s: null peer pointer table", who);
    for (int r = 0; r < kMaxPeers; ++r) t->p[r] = nullptr;
    for (int r = 0; r < R; ++r) {
        MAP_REQUIRE(ptrs[r] != nullptr, "%s: peer pointer %d is null", who, r);
        t->p[r] = ptrs[r];
    }
    return MAP_OK;
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- Philox4x32-10 (repo-wide stream convention)
struct Philox4 {
    uint32_t w0, w1, w2, w3;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// key = seed, counter = (elem_lo, elem_hi, offset_lo, offset_hi)
__device__ __forceinline__ Philox4 philox_elem(uint64_t seed, uint64_t offset, uint64_t elem) {
    return philox4x32_10((uint32_t)elem, (uint32_t)(elem >> 32), (uint32_t)offset, (uint32_t)(offset >> 32),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ uint64_t bounded64(const Philox4& r, uint64_t n) {
    return __umul64hi(((uint64_t)r.w1 << 32) | r.w0, n);
}
__device__ __forceinline__ uint32_t bounded32(uint32_t w, uint32_t n) { return __umulhi(w, n); }
__device__ __forceinline__ float uniform24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-08f; }

// ---------------------------------------------------------------- small device utilities
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {  // read-once data: do not pollute L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_f4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// numerically stable softplus(x) = max(x,0) + log1p(exp(-|x|))  (what BCEWithLogits computes)
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace mapb
