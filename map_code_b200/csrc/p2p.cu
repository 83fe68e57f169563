// K13 over NVLink peer memory: row-sharded tables addressed directly by every GPU of the node.
// Table row `id` lives on rank id % R at local row id / R.  Each rank allocates its shards and its per-step compact
// gradients with map_p2p_alloc (cudaMalloc + CUDA IPC handle); the other ranks map them with map_p2p_open, so the kernels
// below read rows of ANY shard with plain loads that travel over NVLink / NVSwitch:
//   forward   map_emb_gather_sharded_f32 / map_nce_fwd_sharded (nce.cu): the lookup IS the exchange — no id all-to-all,
//             no row all-to-all, no staging buffers;
//   backward  every rank reduces its own occurrences to a compact (unique id, row) list with the single-GPU pipeline;
//             after one barrier the OWNER pulls the entries it owns from all R lists (map_owned_compact -> sort with the
//             source rank in the low key bits -> map_segment_reduce_rows_ex over peer rows) and applies row-wise AdamW to
//             its shard.  Traffic per rank = the rows it needs, independent of R.
// The reference initialises NCCL (code/arguments.py:74) and never issues a collective; NCCL is kept here for the dense
// gradient all-reduce and for the two barriers per step.
#include <string.h>

#include "common.cuh"

namespace mapb {

constexpr int kGatherUnroll = 4;

// out[i, :] = shard[id % R][id / R, :]  — same lane mapping as emb_gather_vec4_kernel (embedding.cu)
__global__ void __launch_bounds__(256) emb_gather_sharded_vec4_kernel(const PeerTable shards, int R, int64_t V, int vpr,
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
            if (row[u] >= 0 && row[u] < V) {
                const int owner = (int)(row[u] % R);
                const float4* base = reinterpret_cast<const float4*>(shards.p[owner]);
                val[u] = __ldg(base + (row[u] / R) * vpr + lane[u]);
            }
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

__global__ void __launch_bounds__(256) emb_gather_sharded_scalar_kernel(const PeerTable shards, int R, int64_t V, int D,
                                                                        const int64_t* __restrict__ ids, int64_t n_elem,
                                                                        float* __restrict__ out, int32_t* oob_flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += stride) {
        const int64_t r = e / D;
        const int d = (int)(e - r * D);
        const int64_t row = __ldg(ids + r);
        float v = 0.f;
        if (row >= 0 && row < V) v = __ldg(reinterpret_cast<const float*>(shards.p[row % R]) + (row / R) * D + d);
        else if (oob_flag != nullptr) *oob_flag = 1;
        out[e] = v;
    }
}

struct PeerLists {
    const int64_t* uniq[kMaxPeers];
    const int32_t* n_unique[kMaxPeers];
};

// Scans the R per-rank unique-id lists (peer loads) and appends the entries this rank owns:
//   keys[k] = (id / R) * R_pow2 + s     (local row in the high bits, source rank s in the low `shift` bits: a stable sort
//                                        on the full key orders the contributions of a row by source rank -> deterministic sums)
//   src[k]  = s * cap + u               (which row of which rank's compact gradient)
// grid.y = source rank; warp-aggregated append.
__global__ void __launch_bounds__(256) owned_compact_kernel(const PeerLists lists, int R, int rank, int shift, int64_t cap,
                                                            int64_t* __restrict__ keys, int32_t* __restrict__ src,
                                                            int32_t* __restrict__ n_out) {
    const int s = blockIdx.y;
    const int64_t n_s = (int64_t)__ldg(lists.n_unique[s]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n_s; base += stride) {
        const int64_t u = base + threadIdx.x;
        int64_t id = -1;
        if (u < n_s) id = __ldg(lists.uniq[s] + u);
        const bool mine = id >= 0 && (int)(id % R) == rank;
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (m == 0) continue;
        int pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(n_out, __popc(m));
        pos0 = __shfl_sync(0xffffffffu, pos0, 0);
        if (mine) {
            const int k = pos0 + __popc(m & ((1u << lane) - 1u));
            keys[k] = ((id / R) << shift) | (int64_t)s;
            src[k] = (int32_t)(s * cap + u);
        }
    }
}

// Stream-ordered barrier over the R ranks of the node through peer memory (replaces a 4-byte NCCL all-reduce: ~35 us -> a few us).
// Every rank owns flags[site][kMaxPeers] (uint32, peer-visible, zero-initialised).  Thread s of the single warp publishes this
// rank's epoch in rank s's flags[site][rank] (release store at system scope: everything earlier kernels of this stream wrote is
// visible before it) and waits until rank s's epoch shows up in its own flags[site][s] (acquire loads).  The epoch lives in
// device memory and advances by one per call, so a captured graph replays the barrier verbatim.  A `site` is one barrier of
// the step's schedule: barriers of different sites may be in flight at the same time on different streams.
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) p2p_barrier_kernel(const PeerTable flags, int R, int rank, int site, unsigned* __restrict__ epochs,
                                                         unsigned* __restrict__ error_word) {
    const unsigned e = epochs[site] + 1u;
    __syncwarp();
    if ((int)threadIdx.x < R) {
        const int s = threadIdx.x;
        __threadfence_system();
        st_release_sys_u32(static_cast<unsigned*>(const_cast<void*>(flags.p[s])) + site * kMaxPeers + rank, e);
        const unsigned* mine = static_cast<const unsigned*>(flags.p[rank]) + site * kMaxPeers + s;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        // (epochs only grow; the signed difference keeps the comparison valid across a wrap)
        // A rank that does not arrive within 30 s of wall time is FATAL (error word for the host, then trap): continuing would
        // let the owner-side sums and AdamW run on unsynchronised peer data.
        while ((int)(ld_acquire_sys_u32(mine) - e) < 0) {
            if ((++spins & 0x3FFu) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                if (now - t0 > 30000000000ull) {
                    if (error_word != nullptr) atomicExch(error_word, 1u);
                    __threadfence_system();
                    __trap();
                }
            }
        }
        __threadfence_system();
    }
    __syncwarp();
    if (threadIdx.x == 0) epochs[site] = e;
}

// ---- dense-gradient all-reduce over peer memory (replaces the NCCL all-reduce of the gradient buckets).  Every rank copies its
// range of the flat gradient into a peer-visible send buffer; after a barrier
//   one shot : every rank sums the R copies of the whole range itself                         (small ranges: one barrier, R x bytes)
//   two shot : rank r sums the R copies of slice r IN PLACE in its own send buffer, barrier,
//              then every rank collects the R reduced slices from their owners                 (large ranges: 2 x bytes)
// The sum runs over the ranks in the same order (0 .. R-1) on every rank: the reduced gradient is bit-identical everywhere, so the
// replicated dense parameters stay bit-identical without a broadcast.  16-byte loads, kReduceUnroll independent peer loads in
// flight per thread and per rank (a remote load is a ~2-4 us round trip through the NVSwitch).
constexpr int kReduceUnroll = 4;

// peer data rewritten every step: never through the non-coherent path, never from a stale L1 line
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 r;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

__global__ void __launch_bounds__(512) p2p_reduce_kernel(const PeerTable send, int R, int64_t first4, int64_t count4, float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride * kReduceUnroll) {
        float4 acc[kReduceUnroll];
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int q = 0; q < R; ++q) {
            const float4* src = reinterpret_cast<const float4*>(send.p[q]) + first4;
            float4 v[kReduceUnroll];
#pragma unroll
            for (int u = 0; u < kReduceUnroll; ++u) {
                const int64_t iu = i + u * stride;
                v[u] = iu < count4 ? ld_volatile_f4(src + iu) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kReduceUnroll; ++u) {
                acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const int64_t iu = i + u * stride;
            if (iu < count4) out[iu] = acc[u];
        }
    }
}

// out[i] = send[i / slice4][first4 + i]: the reduced slices from their owners
__global__ void __launch_bounds__(512) p2p_gather_slices_kernel(const PeerTable send, int R, int64_t first4, int64_t count4, int64_t slice4,
                                                                float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride * kReduceUnroll) {
        float4 v[kReduceUnroll];
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const int64_t iu = i + u * stride;
            if (iu < count4) {
                int q = (int)(iu / slice4);
                if (q >= R) q = R - 1;
                v[u] = ld_volatile_f4(reinterpret_cast<const float4*>(send.p[q]) + first4 + iu);
            }
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const int64_t iu = i + u * stride;
            if (iu < count4) out[iu] = v[u];
        }
    }
}

}  // namespace mapb

extern "C" int map_p2p_barrier(const void* const* flag_ptrs, int R, int rank, int site, int n_sites, uint32_t* epochs,
                               uint32_t* error_word, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(rank >= 0 && rank < R && site >= 0 && site < n_sites && epochs != nullptr, "map_p2p_barrier: bad argument");
    PeerTable t;
    const int rc = fill_peer_table(&t, flag_ptrs, R, "map_p2p_barrier");
    if (rc != MAP_OK) return rc;
    p2p_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(t, R, rank, site, epochs, error_word);
    return check_launch("map_p2p_barrier");
}

extern "C" int map_p2p_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    using namespace mapb;
    MAP_REQUIRE(ptr && handle64 && bytes > 0, "map_p2p_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        set_error("map_p2p_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // zeroes are in place before any peer can map the buffer (barrier flags)
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        set_error("map_p2p_alloc: %s", cudaGetErrorString(e));
        cudaGetLastError();
        cudaFree(p);
        return MAP_ECUDA;
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return MAP_OK;
}

extern "C" int map_p2p_open(const unsigned char* handle64, void** ptr) {
    using namespace mapb;
    MAP_REQUIRE(ptr && handle64, "map_p2p_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("map_p2p_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    *ptr = p;
    return MAP_OK;
}

extern "C" int map_p2p_close(void* ptr) {
    using namespace mapb;
    if (ptr == nullptr) return MAP_OK;
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) {
        set_error("map_p2p_close: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    return MAP_OK;
}

extern "C" int map_p2p_free(void* ptr) {
    using namespace mapb;
    if (ptr == nullptr) return MAP_OK;
    const cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) {
        set_error("map_p2p_free: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    return MAP_OK;
}

extern "C" int map_emb_gather_sharded_f32(const void* const* shard_ptrs, int R, int64_t V, int D, const int64_t* ids,
                                          int64_t n_ids, float* out, int32_t* oob_flag, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(V > 0 && D > 0 && n_ids >= 0, "map_emb_gather_sharded_f32: bad shape V=%lld D=%d n=%lld", (long long)V, D, (long long)n_ids);
    PeerTable t;
    const int rc = fill_peer_table(&t, shard_ptrs, R, "map_emb_gather_sharded_f32");
    if (rc != MAP_OK) return rc;
    if (n_ids == 0) return MAP_OK;
    MAP_REQUIRE(ids && out, "map_emb_gather_sharded_f32: null pointer");
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (D % 4 == 0 && (uintptr_t)out % 16 == 0) {  // shards come from map_p2p_alloc: 256-byte aligned
        const int vpr = D / 4;
        const int64_t n_vec = n_ids * vpr;
        int64_t blocks = ceil_div(n_vec, 256 * kGatherUnroll);
        if (blocks > cap) blocks = cap;
        emb_gather_sharded_vec4_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(t, R, V, vpr, ids, n_vec,
                                                                                        reinterpret_cast<float4*>(out), oob_flag);
    } else {
        const int64_t n_elem = n_ids * D;
        int64_t blocks = ceil_div(n_elem, 256);
        if (blocks > cap) blocks = cap;
        emb_gather_sharded_scalar_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(t, R, V, D, ids, n_elem, out, oob_flag);
    }
    return check_launch("map_emb_gather_sharded_f32");
}

extern "C" int map_owned_compact(const void* const* uniq_ptrs, const void* const* n_unique_ptrs, int R, int rank, int64_t cap,
                                 int64_t* keys, int32_t* src, int32_t* n_out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(R >= 1 && R <= kMaxPeers && rank >= 0 && rank < R && cap > 0 && uniq_ptrs && n_unique_ptrs && keys && src && n_out,
                "map_owned_compact: bad argument");
    MAP_REQUIRE((int64_t)R * cap < ((int64_t)1 << 31), "map_owned_compact: R * cap must fit in int32");
    PeerLists l;
    for (int r = 0; r < kMaxPeers; ++r) {
        l.uniq[r] = nullptr;
        l.n_unique[r] = nullptr;
    }
    for (int r = 0; r < R; ++r) {
        MAP_REQUIRE(uniq_ptrs[r] && n_unique_ptrs[r], "map_owned_compact: peer pointer %d is null", r);
        l.uniq[r] = static_cast<const int64_t*>(uniq_ptrs[r]);
        l.n_unique[r] = static_cast<const int32_t*>(n_unique_ptrs[r]);
    }
    int shift = 0;
    while ((1 << shift) < R) ++shift;
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(n_out, 0, sizeof(int32_t), st) != cudaSuccess) {
        set_error("map_owned_compact: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MAP_ECUDA;
    }
    int64_t bx = ceil_div(cap, 256);
    if (bx > kNumSMs * 2) bx = kNumSMs * 2;
    owned_compact_kernel<<<dim3((unsigned)bx, (unsigned)R), 256, 0, st>>>(l, R, rank, shift, cap, keys, src, n_out);
    return check_launch("map_owned_compact");
}


static unsigned p2p_reduce_grid(int64_t count4) {
    int64_t blocks = mapb::ceil_div(count4, (int64_t)512 * mapb::kReduceUnroll);
    if (blocks > 64) blocks = 64;   // a small grid finds room beside the persistent GEMM CTAs (one per SM) and still keeps ~1 MB of peer loads in flight
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

extern "C" int map_p2p_reduce_f32(const void* const* send_ptrs, int R, int64_t first, int64_t count, float* out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(first >= 0 && count >= 0 && first % 4 == 0 && count % 4 == 0 && out != nullptr && ((uintptr_t)out & 15) == 0,
                "map_p2p_reduce_f32: first / count must be multiples of 4 floats and out 16-byte aligned");
    PeerTable t;
    const int rc = fill_peer_table(&t, send_ptrs, R, "map_p2p_reduce_f32");
    if (rc != MAP_OK) return rc;
    if (count == 0) return MAP_OK;
    p2p_reduce_kernel<<<p2p_reduce_grid(count / 4), 512, 0, as_stream(stream)>>>(t, R, first / 4, count / 4, reinterpret_cast<float4*>(out));
    return check_launch("map_p2p_reduce_f32");
}

extern "C" int map_p2p_gather_slices_f32(const void* const* send_ptrs, int R, int64_t first, int64_t count, int64_t slice, float* out,
                                         map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(first >= 0 && count >= 0 && slice > 0 && first % 4 == 0 && count % 4 == 0 && slice % 4 == 0 && out != nullptr &&
                    ((uintptr_t)out & 15) == 0,
                "map_p2p_gather_slices_f32: first / count / slice must be multiples of 4 floats and out 16-byte aligned");
    PeerTable t;
    const int rc = fill_peer_table(&t, send_ptrs, R, "map_p2p_gather_slices_f32");
    if (rc != MAP_OK) return rc;
    if (count == 0) return MAP_OK;
    p2p_gather_slices_kernel<<<p2p_reduce_grid(count / 4), 512, 0, as_stream(stream)>>>(t, R, first / 4, count / 4, slice / 4,
                                                                                       reinterpret_cast<float4*>(out));
    return check_launch("map_p2p_gather_slices_f32");
}

