// K2 (first half): deduplicate ids and reduce per-occurrence row gradients into a compact [U, D] gradient.
//   radix sort (id, occurrence)  ->  unique ids + segment starts  ->  segmented row sums.
// Replaces aten::embedding_dense_backward (thrust sort + dense [V,D] grad) behind code/layers.py:98 and the dense
// index_add behind code/nce/index_linear.py:99-100.  Everything is integer / HBM-L2 bound; no tensor cores.
#include <stdlib.h>

#include "common.cuh"

namespace mapb {

constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per CTA
constexpr int kSortWarps = kSortThreads / 32;

// The number of keys may live on the device (n_dev != nullptr: the owner-side merge of the row-sharded tables compacts a
// data-dependent number of entries): grids are sized for the capacity n, every kernel clamps to min(n, *n_dev).
__device__ __forceinline__ int64_t eff_n(int64_t n, const int32_t* n_dev) {
    if (n_dev == nullptr) return n;
    const int64_t d = (int64_t)__ldg(n_dev);
    return d < n ? d : n;
}

// ---------------------------------------------------------------------------------------------- sort
// exclusive scan of one value per thread over the 256 threads of the CTA (warp_tot: scratch); returns the exclusive prefix,
// *total = sum over the CTA
__device__ __forceinline__ uint32_t cta_excl_scan_256(uint32_t t, uint32_t* warp_tot, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __syncthreads();   // warp_tot may still be read by the previous use
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t x = warp_tot[w];
        if (w < warp) off += x;
        tot += x;
    }
    if (total != nullptr) *total = tot;
    return off + incl - t;
}

// one shared-memory atomic per distinct digit of the warp (the <mask> row alone is 8 % of the embedding ids: per-key atomics on
// one counter serialise)
__device__ __forceinline__ void warp_hist_add(uint32_t* sh, uint32_t digit, bool valid) {
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : 0xFFFFFFFFu);
    if (valid && (peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) atomicAdd(&sh[digit], (uint32_t)__popc(peers));
}

// Shared-memory state of one 2048-key tile
struct TileSmem {
    uint32_t warp_cnt[kSortWarps][kRadixBins];  // per-warp digit counts -> exclusive prefixes over the warps
    uint32_t tile_run[kRadixBins];              // digit counts of the tile
    uint32_t tile_excl[kRadixBins];             // their exclusive scan
    uint32_t gbase[kRadixBins];                 // global start of (digit, this tile)
    uint32_t warp_tot[kSortWarps];
    uint32_t keys_s[kSortTile];
    uint32_t vals_s[kSortTile];
};

// Loads a tile, ranks every key among the keys of the same digit that precede it, and reorders the tile by digit in shared
// memory (stable) so that the global writes of one digit are contiguous runs.  Element order inside the tile is
// (warp, round, lane) == ascending index: warp w owns the contiguous keys [256 w, 256 w + 256) and ranks them against its OWN
// histogram row with match.any — no CTA-wide synchronisation inside the 8 rounds (the earlier version met at 3 barriers per
// round) — then one pass over the 8 rows turns them into exclusive prefixes over the warps.
// On return: tile_run[d] = #keys of digit d, tile_excl[d] = exclusive scan, keys_s / vals_s = the tile sorted by digit.
template <typename LoadKey, typename LoadVal>
__device__ __forceinline__ void tile_rank_reorder(TileSmem& sm, int tile_n, int shift, LoadKey load_key, LoadVal load_val) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) sm.warp_cnt[w][tid] = 0;
    uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int li = warp * (kSortItems * 32) + it * 32 + lane;
        const bool valid = li < tile_n;
        key[it] = valid ? load_key(li) : 0u;
        val[it] = valid ? load_val(li) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int li = warp * (kSortItems * 32) + it * 32 + lane;
        const bool valid = li < tile_n;
        const uint32_t digit = valid ? ((key[it] >> shift) & (kRadixBins - 1)) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        const uint32_t before = valid ? sm.warp_cnt[warp][digit] : 0u;
        __syncwarp();
        if (valid && rank_in_warp == 0) sm.warp_cnt[warp][digit] = before + __popc(peers);
        __syncwarp();
        rank[it] = before + rank_in_warp;
    }
    __syncthreads();
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t c = sm.warp_cnt[w][tid];
        sm.warp_cnt[w][tid] = run;
        run += c;
    }
    sm.tile_run[tid] = run;
    sm.tile_excl[tid] = cta_excl_scan_256(run, sm.warp_tot, nullptr);
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; ++it) {
        const int li = warp * (kSortItems * 32) + it * 32 + lane;
        if (li < tile_n) {
            const uint32_t digit = (key[it] >> shift) & (kRadixBins - 1);
            const uint32_t lp = sm.tile_excl[digit] + sm.warp_cnt[warp][digit] + rank[it];
            sm.keys_s[lp] = key[it];
            sm.vals_s[lp] = val[it];
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- segmented row sums
constexpr int kSegTile = 8;  // sorted positions per lane-group; all 8 row loads are issued before the first add

template <int VEC>
struct RowVec;
template <>
struct RowVec<4> {
    float4 v;
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void fma(const RowVec& a, float s) {
        v.x = fmaf(a.v.x, s, v.x); v.y = fmaf(a.v.y, s, v.y); v.z = fmaf(a.v.z, s, v.z); v.w = fmaf(a.v.w, s, v.w);
    }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void atomic_add(float* p) const {
        atomicAdd(p + 0, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
    }
};
template <>
struct RowVec<1> {
    float v;
    __device__ __forceinline__ void zero() { v = 0.f; }
    __device__ __forceinline__ void load(const float* p) { v = *p; }
    __device__ __forceinline__ void fma(const RowVec& a, float s) { v = fmaf(a.v, s, v); }
    __device__ __forceinline__ void store(float* p) const { *p = v; }
    __device__ __forceinline__ void atomic_add(float* p) const { atomicAdd(p, v); }
};

// One group of `lanes` threads (lanes = D / VEC) walks kSegTile consecutive sorted positions.  A segment that lies entirely
// inside the tile is written with a plain store.  Segments that straddle tile borders (hot ids: <mask> = 3 collects B*L
// occurrences) are first merged ACROSS THE GROUPS OF THE CTA through shared memory — every group deposits at most a
// "head" and a "tail" partial — and only the per-CTA result goes to global memory: a plain store when the segment lies
// inside the CTA's range, otherwise one atomic per CTA (instead of one per 8 occurrences).
template <int VEC, bool HAS_MAP>
__global__ void __launch_bounds__(256) segment_reduce_kernel(const float* __restrict__ rows, int64_t ld_rows, int lanes,
                                                             const float* __restrict__ scale, int group,
                                                             const int32_t* __restrict__ occ_sorted,
                                                             const int32_t* __restrict__ seg_start,
                                                             const int32_t* __restrict__ n_unique, int64_t n,
                                                             float* __restrict__ grad, float* __restrict__ scalar_out,
                                                             int D, const int32_t* __restrict__ n_dev,
                                                             const int32_t* __restrict__ occ_map,
                                                             const PeerTable row_tab, int64_t rows_per_peer,
                                                             const int32_t* __restrict__ pos_seg) {
    __shared__ int slot_seg[2 * 256];                  // segment id of every head (2g) / tail (2g+1) slot, -1 = empty
    __shared__ float slot_sc[2 * 256];
    __shared__ __align__(16) float slot_vec[2 * 256 * VEC];  // [slot][lanes * VEC] with 2 * G * lanes * VEC <= 512 * VEC
    n = eff_n(n, n_dev);
    const int G = 256 / lanes;                         // groups per CTA
    const int g = threadIdx.x / lanes;
    const int lane = threadIdx.x - g * lanes;
    const bool in_group = g < G;
    const int64_t blk_t0 = (int64_t)blockIdx.x * G * kSegTile;
    const int64_t blk_t1 = (blk_t0 + (int64_t)G * kSegTile < n) ? blk_t0 + (int64_t)G * kSegTile : n;
    const int64_t t0 = blk_t0 + (int64_t)g * kSegTile;
    const bool active = in_group && t0 < n;
    const int64_t t1 = (t0 + kSegTile < n) ? t0 + kSegTile : n;
    const int U = *n_unique;
    const int rowlen = lanes * VEC;
    if (in_group && lane == 0) {
        slot_seg[2 * g] = -1;
        slot_seg[2 * g + 1] = -1;
    }
    if (active) {
        // segment index of every position of the tile, of the position before it and of the position after it (-1 = outside
        // the list).  With the position -> segment map of the dedup these are 10 independent loads issued together with the
        // occurrence loads; without it (plain map_segment_reduce_rows) they come from a binary search over seg_start and a
        // walk of dependent loads (a 20-step chain per group at 1e6 unique rows, and one more round trip per segment).
        int sg[kSegTile + 1];
        int sg_prev = -1;
        if (HAS_MAP) {
#pragma unroll
            for (int k = 0; k <= kSegTile; ++k) sg[k] = (t0 + k < n) ? __ldg(pos_seg + t0 + k) : -1;
            if (t0 > 0) sg_prev = __ldg(pos_seg + t0 - 1);
        } else {
            int lo = 0, hi = U - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if ((int64_t)seg_start[mid] <= t0) lo = mid; else hi = mid - 1;
            }
            int cur = lo;
            int64_t cur_end = seg_start[cur + 1];
            if ((int64_t)seg_start[cur] < t0) sg_prev = cur;
            else if (t0 > 0) sg_prev = cur - 1;
#pragma unroll
            for (int k = 0; k <= kSegTile; ++k) {
                if (t0 + k < n) {
                    if (t0 + k >= cur_end) {
                        ++cur;
                        cur_end = seg_start[cur + 1];
                    }
                    sg[k] = cur;
                } else {
                    sg[k] = -1;
                }
            }
        }
        RowVec<VEC> r[kSegTile];
        float sc[kSegTile];
#pragma unroll
        for (int k = 0; k < kSegTile; ++k) {
            sc[k] = 0.f;
            r[k].zero();
            if (t0 + k < t1) {
                int64_t occ = occ_sorted[t0 + k];
                if (occ_map != nullptr) occ = __ldg(occ_map + occ);   // compacted entry -> (peer, row) code
                sc[k] = (scale != nullptr) ? __ldg(scale + occ) : 1.f;
                if (rows_per_peer > 0) {  // rows live in R peer buffers (NVLink loads): code = peer * rows_per_peer + row
                    const int64_t peer = occ / rows_per_peer;
                    r[k].load(static_cast<const float*>(row_tab.p[peer]) + (occ - peer * rows_per_peer) * ld_rows + lane * VEC);
                } else {
                    r[k].load(rows + (occ / group) * ld_rows + lane * VEC);
                }
            }
        }
        RowVec<VEC> acc;
        acc.zero();
        float sacc = 0.f;
        bool started_here = sg_prev != sg[0];   // does the segment being accumulated begin inside this tile?
#pragma unroll
        for (int k = 0; k < kSegTile; ++k) {
            const int64_t pos = t0 + k;
            if (pos < t1) {
                const int s = sg[k];
                acc.fma(r[k], sc[k]);
                sacc += sc[k];
                const bool ends_here = sg[k + 1] != s;   // (also true at the end of the list: sg = -1 there)
                if (ends_here || pos + 1 == t1) {        // flush
                    if (started_here && ends_here) {      // interior segment of this tile
                        acc.store(grad + (int64_t)s * D + lane * VEC);
                        if (scalar_out != nullptr && lane == 0) scalar_out[s] = sacc;
                    } else {                              // straddles: head slot if it began before the tile, else tail slot
                        const int slot = started_here ? 2 * g + 1 : 2 * g;
                        acc.store(slot_vec + slot * rowlen + lane * VEC);
                        if (lane == 0) {
                            slot_seg[slot] = s;
                            slot_sc[slot] = sacc;
                        }
                    }
                    acc.zero();
                    sacc = 0.f;
                    started_here = true;                  // whatever follows begins inside the tile
                }
            }
        }
    }
    __syncthreads();
    if (!in_group) return;
    // merge equal-segment runs of slots (in position order; empty slots are skipped); the first slot of a run is its leader
    for (int h = 0; h < 2; ++h) {
        const int i = 2 * g + h;
        const int s = slot_seg[i];
        if (s < 0) continue;
        int j = i - 1;
        while (j >= 0 && slot_seg[j] < 0) --j;
        if (j >= 0 && slot_seg[j] == s) continue;  // not the leader
        RowVec<VEC> acc;
        acc.load(slot_vec + i * rowlen + lane * VEC);
        float sacc = slot_sc[i];
        for (int k = i + 1; k < 2 * G; ++k) {
            const int sk = slot_seg[k];
            if (sk < 0) continue;
            if (sk != s) break;
            RowVec<VEC> o;
            o.load(slot_vec + k * rowlen + lane * VEC);
            acc.fma(o, 1.f);
            sacc += slot_sc[k];
        }
        bool whole;   // does the segment lie entirely inside this CTA's range of positions?
        if (HAS_MAP) {
            const bool cut_front = blk_t0 > 0 && __ldg(pos_seg + blk_t0) == s && __ldg(pos_seg + blk_t0 - 1) == s;
            const bool cut_back = blk_t1 < n && __ldg(pos_seg + blk_t1 - 1) == s && __ldg(pos_seg + blk_t1) == s;
            whole = !cut_front && !cut_back;
        } else {
            whole = ((int64_t)seg_start[s] >= blk_t0) && ((int64_t)seg_start[s + 1] <= blk_t1);
        }
        float* dst = grad + (int64_t)s * D + lane * VEC;
        if (whole) acc.store(dst); else acc.atomic_add(dst);
        if (scalar_out != nullptr && lane == 0) {
            if (whole) scalar_out[s] = sacc; else atomicAdd(scalar_out + s, sacc);
        }
    }
}

// zero the first U rows of the compact gradient (only these can be touched by atomics)
__global__ void __launch_bounds__(256) zero_unique_rows_kernel(float* grad, float* scalar_out, int D,
                                                               const int32_t* __restrict__ n_unique, int64_t max_elems) {
    const int64_t U = *n_unique;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U * D && i < max_elems; i += stride) grad[i] = 0.f;
    if (scalar_out != nullptr)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += stride) scalar_out[i] = 0.f;
}

// With the position -> segment map only the rows of segments that a CTA border of segment_reduce_kernel cuts can receive
// atomics: one warp per border zeroes that one row (zeroing all U rows first cost U * D * 4 bytes of HBM writes, half as much
// again as the reduce itself reads at C5).
__global__ void __launch_bounds__(256) zero_cut_rows_kernel(float* grad, float* scalar_out, int D, const int32_t* __restrict__ pos_seg,
                                                            int64_t n, const int32_t* __restrict__ n_dev, int64_t per_cta,
                                                            int64_t n_borders) {
    n = eff_n(n, n_dev);
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t b = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) + 1; b <= n_borders; b += warps) {
        const int64_t p = b * per_cta;   // first position of CTA b
        if (p >= n) break;
        const int s = __ldg(pos_seg + p);
        if (__ldg(pos_seg + p - 1) != s) continue;
        for (int d = lane; d < D; d += 32) grad[(int64_t)s * D + d] = 0.f;
        if (scalar_out != nullptr && lane == 0) scalar_out[s] = 0.f;
    }
}

__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ grad, const int64_t* __restrict__ uniq,
                                                           const int32_t* __restrict__ n_unique, int D,
                                                           float* __restrict__ dense) {
    const int64_t U = *n_unique;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U * D; i += stride) {
        const int64_t u = i / D;
        dense[uniq[u] * D + (i - u * D)] = grad[i];
    }
}

// ---------------------------------------------------------------------------------------------- single-launch pipeline
// The whole of K2a (key extraction, every radix pass, head flags, unique ids / segment starts / position -> segment map) as
// ONE launch of persistent CTAs that meet at grid barriers.  At the step's sizes (n = 1.6e5 .. 3.2e5 keys) a pipeline of
// separate launches (prep, {histogram, scan, scatter} per pass, head count, scan, emit: 13 dependent launches of a few
// microseconds each, what this file did until r01e: 89 us for 1.6e5 keys) is launch/drain latency, not bandwidth; here a pass
// costs one barrier (41 us for the same keys):
//   H   CTA g builds the digit-0 histogram of its contiguous key range                              -> hist[0][g][256]
//   S_p thread b sums column b of hist[p] (all G rows: totals; rows < g: its own prefix), the CTA scans the 256 totals, then
//       ranks / reorders / scatters its tiles (tile_rank_reorder); while writing key k to position `pos` it also
//       counts k's NEXT digit into hist[p+1][owner CTA of pos] (global reductions), so pass p+1 needs no histogram sweep
//   U   head flags of the sorted keys: per-CTA counts -> barrier -> every CTA emits its range at its prefix
// G <= 2 CTAs per SM (256 threads, < 32 KB shared memory), so all CTAs are co-resident on an otherwise idle GPU; beside other
// kernels the late CTAs simply arrive late (nothing they wait for depends on this kernel).  Cross-CTA data is read with
// ld.global.cg (L2): the buffers are rewritten between passes, so the non-coherent / L1 paths must not be used for them.
constexpr int kPersistMaxCtas = 2 * kNumSMs;
constexpr unsigned long long kBarrierTimeoutNs = 5000000000ull;   // wall-clock bound of one grid-barrier wait (%globaltimer)

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long dedup_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// sync[0] = arrival counter (zeroed by the host before the launch), sync[1] = error word.
// A barrier that is not met within kBarrierTimeoutNs of WALL time (not a poll count: the bound does not depend on clocks,
// time-slicing or co-running kernels) is FATAL: the error word is set for the host and the kernel traps, so the launch fails
// through the normal CUDA error path and nothing downstream (segment sums, AdamW) ever consumes a partial sort.
__device__ __forceinline__ bool grid_barrier(unsigned* sync, unsigned target, int* ok_s) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(sync, 1u);
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_u32(sync) < target) {
            if ((++spins & 0x3FFu) == 0u) {
                const unsigned long long now = dedup_globaltimer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kBarrierTimeoutNs || ld_acquire_u32(sync + 1) != 0u) {
                    atomicExch(sync + 1, 1u);
                    __threadfence_system();
                    __trap();
                }
            }
        }
        __threadfence();
        *ok_s = 1;
    }
    __syncthreads();
    return *ok_s != 0;
}

struct PersistSmem : TileSmem {
    uint32_t carry;
    int ok;
};

__global__ void __launch_bounds__(kSortThreads, 3) dedup_persistent_kernel(
    const int64_t* __restrict__ ids, int64_t n_cap, const int32_t* __restrict__ n_dev, int passes, int seg_shift, uint32_t* keys_a,
    uint32_t* keys_b, uint32_t* vals_a, uint32_t* occ, uint32_t* hist, uint32_t* cnt, unsigned* sync, int64_t* __restrict__ uniq_ids,
    int32_t* __restrict__ seg_start, int32_t* __restrict__ n_unique, int32_t* __restrict__ pos_seg) {
    __shared__ PersistSmem sm;
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = blockIdx.x;
    const int64_t n = eff_n(n_cap, n_dev);
    const int64_t ntiles = (n + kSortTile - 1) / kSortTile;
    // with a device-side count (owner-side merge: the grid is sized for the capacity R * cap) only as many CTAs as there are
    // tiles take part: the others leave at once and the barriers / column sums span G rows only
    const int G = ((int64_t)gridDim.x < ntiles) ? (int)gridDim.x : (ntiles > 0 ? (int)ntiles : 1);
    if (g >= G) return;
    uint32_t tpc = (uint32_t)((ntiles + G - 1) / G);   // tiles per CTA
    if (tpc < 1) tpc = 1;
    const int64_t tile0 = (int64_t)g * tpc;
    const int64_t tile1 = (tile0 + tpc < ntiles) ? tile0 + tpc : ntiles;   // (tile0 >= tile1: this CTA only attends the barriers)
    unsigned epoch = 0;
    // phase timestamps of CTA 0 (globaltimer ns) behind the two sync words: start, H, barrier, then per pass (column sums,
    // tiles, barrier), head count, barrier, emit — read back through map_dedup_debug_offset (tuning aid, one store per phase)
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(sync + 16);
    int n_stamp = 0;
    auto stamp = [&]() {
        if (g == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            stamps[n_stamp] = t;
        }
        ++n_stamp;
    };
    stamp();

    // ---- H: digit-0 histogram of the own range (keys come straight from the int64 ids); rows of the later passes start at 0
    sm.tile_run[tid] = 0;
    for (int p = 1; p < passes; ++p) hist[((size_t)p * G + g) * kRadixBins + tid] = 0u;
    __syncthreads();
    for (int64_t t = tile0; t < tile1; ++t) {
        const int64_t base = t * kSortTile;
#pragma unroll
        for (int it = 0; it < kSortItems; ++it) {
            const int64_t i = base + it * kSortThreads + tid;
            const bool valid = i < n;
            warp_hist_add(sm.tile_run, valid ? ((uint32_t)ids[i] & (kRadixBins - 1)) : 0u, valid);
        }
    }
    __syncthreads();
    hist[(size_t)g * kRadixBins + tid] = sm.tile_run[tid];
    stamp();
    if (!grid_barrier(sync, (++epoch) * G, &sm.ok)) return;
    stamp();

    // ---- S_p
    for (int p = 0; p < passes; ++p) {
        const int shift = p * kRadixBits;
        const bool last = (p + 1 == passes);
        const uint32_t* kin = (p & 1) ? keys_a : keys_b;              // written by pass p-1 (unused for p == 0)
        uint32_t* kout = (p & 1) ? keys_b : keys_a;
        const uint32_t* vin = ((passes - p) & 1) ? vals_a : occ;      // pass p-1 wrote where ((passes-1-(p-1)) even ? occ : vals_a)
        uint32_t* vout = ((passes - 1 - p) & 1) ? vals_a : occ;       // the last pass lands in occ_sorted
        uint32_t* hnext = hist + (size_t)(p + 1) * G * kRadixBins;
        {   // column sums of this pass's histogram: thread = digit
            const uint32_t* h = hist + (size_t)p * G * kRadixBins + tid;
            uint32_t tot = 0, pre = 0;
            int r = 0;
            for (; r + 32 <= G; r += 32) {   // 32 independent L2 loads in flight per thread: the walk is latency-bound
                uint32_t c[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) c[k] = __ldcg(h + (size_t)(r + k) * kRadixBins);
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    tot += c[k];
                    if (r + k < g) pre += c[k];
                }
            }
            {   // tail: the same batch with a row guard
                uint32_t c[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) c[k] = (r + k < G) ? __ldcg(h + (size_t)(r + k) * kRadixBins) : 0u;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    tot += c[k];
                    if (r + k < g) pre += c[k];
                }
            }
            const uint32_t excl = cta_excl_scan_256(tot, sm.warp_tot, nullptr);
            sm.gbase[tid] = excl + pre;
        }
        __syncthreads();
        stamp();
        for (int64_t t = tile0; t < tile1; ++t) {
            const int64_t base = t * kSortTile;
            const int tile_n = (int)((n - base < kSortTile) ? (n - base) : kSortTile);
            if (p == 0)
                tile_rank_reorder(sm, tile_n, shift, [&](int li) { return (uint32_t)ids[base + li]; },
                                  [&](int li) { return (uint32_t)(base + li); });
            else
                tile_rank_reorder(sm, tile_n, shift, [&](int li) { return __ldcg(kin + base + li); },
                                  [&](int li) { return __ldcg(vin + base + li); });
            // coalesced runs to global memory + the next pass's histogram of the destination range (one reduction per distinct
            // (owner CTA, next digit) of the warp: hot keys would otherwise serialise on one L2 address)
            for (int i0 = 0; i0 < tile_n; i0 += kSortThreads) {
                const int i = i0 + tid;
                const bool valid = i < tile_n;
                uint32_t code = 0xFFFFFFFFu;
                if (valid) {
                    const uint32_t k = sm.keys_s[i];
                    const uint32_t digit = (k >> shift) & (kRadixBins - 1);
                    const uint32_t pos = sm.gbase[digit] + ((uint32_t)i - sm.tile_excl[digit]);
                    kout[pos] = k;
                    vout[pos] = sm.vals_s[i];
                    code = (uint32_t)((pos / kSortTile) / tpc) * kRadixBins + ((k >> (shift + kRadixBits)) & (kRadixBins - 1));
                }
                if (!last) {
                    const uint32_t peers = __match_any_sync(0xffffffffu, code);
                    if (valid && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(hnext + code, (uint32_t)__popc(peers));
                }
            }
            __syncthreads();
            sm.gbase[tid] += sm.tile_run[tid];
            __syncthreads();
        }
        stamp();
        if (!grid_barrier(sync, (++epoch) * G, &sm.ok)) return;
        stamp();
    }

    // ---- U: head flags.  sorted keys are in the buffer the last pass wrote
    const uint32_t* ks = ((passes - 1) & 1) ? keys_b : keys_a;
    uint32_t my_heads = 0;
    for (int64_t t = tile0; t < tile1; ++t) {
        const int64_t base = t * kSortTile + (int64_t)tid * kSortItems;
        uint32_t prev = (base > 0 && base < n) ? (__ldcg(ks + base - 1) >> seg_shift) : 0u;
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            if (base + k < n) {
                const uint32_t cur = __ldcg(ks + base + k) >> seg_shift;
                my_heads += (base + k == 0 || cur != prev) ? 1u : 0u;
                prev = cur;
            }
        }
    }
    {
        uint32_t total;
        (void)cta_excl_scan_256(my_heads, sm.warp_tot, &total);
        if (tid == 0) cnt[g] = total;
    }
    stamp();
    if (!grid_barrier(sync, (++epoch) * G, &sm.ok)) return;
    stamp();
    {
        uint32_t part = 0;
        for (int r = tid; r < g; r += kSortThreads) part += __ldcg(cnt + r);
        uint32_t total;
        (void)cta_excl_scan_256(part, sm.warp_tot, &total);
        if (tid == 0) sm.carry = total;
        __syncthreads();
    }
    for (int64_t t = tile0; t < tile1; ++t) {
        const int64_t base = t * kSortTile + (int64_t)tid * kSortItems;
        uint32_t seg[kSortItems];
        uint32_t f = 0, tsum = 0;
        uint32_t prev = (base > 0 && base < n) ? (__ldcg(ks + base - 1) >> seg_shift) : 0u;
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            seg[k] = 0;
            if (base + k < n) {
                const uint32_t cur = __ldcg(ks + base + k) >> seg_shift;
                seg[k] = cur;
                if (base + k == 0 || cur != prev) {
                    f |= 1u << k;
                    ++tsum;
                }
                prev = cur;
            }
        }
        uint32_t tile_total;
        uint32_t u = sm.carry + cta_excl_scan_256(tsum, sm.warp_tot, &tile_total);
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            if (base + k < n) {
                if (f & (1u << k)) {
                    uniq_ids[u] = (int64_t)seg[k];
                    seg_start[u] = (int32_t)(base + k);
                    ++u;
                }
                if (pos_seg != nullptr) pos_seg[base + k] = (int32_t)u - 1;
                if (base + k == n - 1) {  // the last element closes the list
                    seg_start[u] = (int32_t)n;
                    *n_unique = (int32_t)u;
                }
            }
        }
        __syncthreads();
        if (tid == 0) sm.carry += tile_total;
        __syncthreads();
    }
    stamp();
    if (n == 0 && g == 0 && tid == 0) {  // empty list (possible with a device-side count)
        seg_start[0] = 0;
        *n_unique = 0;
    }
}

struct DedupLayout {
    size_t keys_a, keys_b, vals_a, phist, pcnt, psync, total;
    int persist_ctas;
};
static DedupLayout dedup_layout(int64_t n) {
    DedupLayout L;
    const int64_t ntiles = ceil_div(n > 0 ? n : 1, kSortTile);
    // MAP_B200_DEDUP_CTAS caps the persistent grid: beside the one-CTA-per-SM persistent GEMMs a sort that wants every SM would
    // spin at its grid barriers on the SMs it got while the rest of its CTAs wait for a GEMM launch to end (and the next GEMM
    // launch waits for the spinning ones); a small grid fits on the SMs the GEMMs leave free (MAP_B200_GEMM_CLUSTERS)
    static int env_cap = -1;
    if (env_cap < 0) {
        const char* e = getenv("MAP_B200_DEDUP_CTAS");
        env_cap = (e != nullptr && atoi(e) >= 1) ? atoi(e) : 0;
        if (env_cap > kPersistMaxCtas) env_cap = kPersistMaxCtas;
    }
    // Up to 1 M keys (the id streams of a batch-4096 step: L2-resident, bound by the grid barriers, not by bandwidth) the sort
    // takes at most 96 CTAs: it then holds a third of the SMs instead of all of them while the dynamically scheduled GEMM launches
    // of the other streams keep working on the rest (C2 step 0.903 -> 0.881 ms, scripts/rounds/r2_36.sh; 64 CTAs: 0.890, 32: 0.920).
    // Larger streams (C5: 2.5 M keys, HBM-bound) keep the full grid.
    const int cap = env_cap > 0 ? env_cap : (n <= ((int64_t)1 << 20) ? 96 : kPersistMaxCtas);
    L.persist_ctas = ntiles < cap ? (int)ntiles : cap;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    L.keys_a = off; off += align((size_t)n * 4);
    L.keys_b = off; off += align((size_t)n * 4);
    L.vals_a = off; off += align((size_t)n * 4);
    L.phist = off; off += align((size_t)4 * L.persist_ctas * kRadixBins * 4);   // one histogram per pass (key_bits <= 32)
    L.pcnt = off; off += align((size_t)L.persist_ctas * 4);
    L.psync = off; off += 512;   // 2 sync words, then (at +64) up to 32 phase timestamps
    L.total = off;
    return L;
}

}  // namespace mapb

extern "C" size_t map_dedup_workspace_bytes(int64_t n_ids) { return mapb::dedup_layout(n_ids).total; }

extern "C" int map_dedup_ids_ex(const int64_t* ids, int64_t n, const int32_t* n_dev, int key_bits, int seg_shift, int64_t* uniq_ids,
                                int32_t* seg_start, int32_t* occ_sorted, int32_t* n_unique, int32_t* pos_seg, void* workspace,
                                size_t workspace_bytes, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(ids && uniq_ids && seg_start && occ_sorted && n_unique && workspace, "map_dedup_ids: null pointer");
    MAP_REQUIRE(n > 0 && n < (int64_t)1 << 31, "map_dedup_ids: n=%lld out of range", (long long)n);
    MAP_REQUIRE(key_bits >= 1 && key_bits <= 32, "map_dedup_ids: key_bits=%d", key_bits);
    MAP_REQUIRE(seg_shift >= 0 && seg_shift < key_bits, "map_dedup_ids: seg_shift=%d", seg_shift);
    const DedupLayout L = dedup_layout(n);
    if (workspace_bytes < L.total) {
        set_error("map_dedup_ids: workspace %zu < %zu", workspace_bytes, L.total);
        return MAP_EWORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    const int passes = (key_bits + kRadixBits - 1) / kRadixBits;
    unsigned* sync = reinterpret_cast<unsigned*>(ws + L.psync);
    if (cudaMemsetAsync(sync, 0, 8, st) != cudaSuccess) {   // arrival counter + error word of the grid barriers
        set_error("map_dedup_ids: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MAP_ECUDA;
    }
    dedup_persistent_kernel<<<L.persist_ctas, kSortThreads, 0, st>>>(
        ids, n, n_dev, passes, seg_shift, reinterpret_cast<uint32_t*>(ws + L.keys_a), reinterpret_cast<uint32_t*>(ws + L.keys_b),
        reinterpret_cast<uint32_t*>(ws + L.vals_a), reinterpret_cast<uint32_t*>(occ_sorted), reinterpret_cast<uint32_t*>(ws + L.phist),
        reinterpret_cast<uint32_t*>(ws + L.pcnt), sync, uniq_ids, seg_start, n_unique, pos_seg);
    return check_launch("map_dedup_ids");
}

/* byte offset, inside the workspace of map_dedup_ids for n_ids keys, of the uint64 phase timestamps the kernel's CTA 0 leaves
 * (globaltimer ns: start, histogram, barrier, per pass {column sums, tiles, barrier}, head count, barrier, emit) */
extern "C" size_t map_dedup_debug_offset(int64_t n_ids) { return mapb::dedup_layout(n_ids).psync + 64; }

extern "C" int map_dedup_ids(const int64_t* ids, int64_t n, int key_bits, int64_t* uniq_ids, int32_t* seg_start,
                             int32_t* occ_sorted, int32_t* n_unique, void* workspace, size_t workspace_bytes,
                             map_stream_t stream) {
    return map_dedup_ids_ex(ids, n, nullptr, key_bits, 0, uniq_ids, seg_start, occ_sorted, n_unique, nullptr, workspace, workspace_bytes,
                            stream);
}

extern "C" int map_segment_reduce_rows_ex(const float* rows, int64_t ld_rows, int D, const float* scale, int group,
                                          const int32_t* occ_sorted, const int32_t* seg_start, const int32_t* n_unique,
                                          const int32_t* pos_seg, int64_t n_ids, const int32_t* n_dev, const int32_t* occ_map,
                                          const float* const* row_ptrs, int n_peers, int64_t rows_per_peer, float* grad_compact,
                                          float* scalar_out, map_stream_t stream) {
    using namespace mapb;
    PeerTable row_tab{};
    if (row_ptrs != nullptr) {  // host array of device pointers -> by-value kernel parameter
        const int rc = fill_peer_table(&row_tab, reinterpret_cast<const void* const*>(row_ptrs), n_peers, "map_segment_reduce_rows_ex");
        if (rc != MAP_OK) return rc;
    } else {
        rows_per_peer = 0;
    }
    MAP_REQUIRE((rows || row_ptrs) && occ_sorted && seg_start && n_unique && grad_compact, "map_segment_reduce_rows: null pointer");
    MAP_REQUIRE(D >= 1 && group >= 1 && n_ids > 0, "map_segment_reduce_rows: bad shape D=%d group=%d n=%lld", D, group, (long long)n_ids);
    MAP_REQUIRE(row_ptrs == nullptr || (rows_per_peer > 0 && group == 1), "map_segment_reduce_rows: peer rows need rows_per_peer > 0, group == 1");
    cudaStream_t st = as_stream(stream);
    // peer buffers come from map_p2p_alloc (256-byte aligned), so only the local pointers need the alignment check
    const bool vec = (D % 4 == 0) && (ld_rows % 4 == 0) && ((uintptr_t)rows % 16 == 0) && ((uintptr_t)grad_compact % 16 == 0);
    const int lanes = vec ? D / 4 : D;
    MAP_REQUIRE(lanes <= 256, "map_segment_reduce_rows: D=%d too wide for one CTA", D);
    const int64_t per_cta = (int64_t)(256 / lanes) * kSegTile;   // sorted positions per CTA of segment_reduce_kernel
    const unsigned blocks = (unsigned)ceil_div(n_ids, per_cta);
    if (pos_seg != nullptr) {
        if (blocks > 1) {
            int64_t zb = ceil_div((int64_t)(blocks - 1) * 32, 256);
            if (zb > kNumSMs * 8) zb = kNumSMs * 8;
            zero_cut_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(grad_compact, scalar_out, D, pos_seg, n_ids, n_dev, per_cta, (int64_t)blocks - 1);
        }
    } else {
        const int64_t max_elems = n_ids * D;
        int64_t zb = ceil_div(max_elems, 256);
        if (zb > kNumSMs * 8) zb = kNumSMs * 8;
        zero_unique_rows_kernel<<<(unsigned)zb, 256, 0, st>>>(grad_compact, scalar_out, D, n_unique, max_elems);
    }
#define LAUNCH_SEGRED(V, M)                                                                                                     \
    segment_reduce_kernel<V, M><<<blocks, 256, 0, st>>>(rows, ld_rows, lanes, scale, group, occ_sorted, seg_start, n_unique, n_ids, \
                                                        grad_compact, scalar_out, D, n_dev, occ_map, row_tab, rows_per_peer, pos_seg)
    if (vec && pos_seg != nullptr) LAUNCH_SEGRED(4, true);
    else if (vec) LAUNCH_SEGRED(4, false);
    else if (pos_seg != nullptr) LAUNCH_SEGRED(1, true);
    else LAUNCH_SEGRED(1, false);
#undef LAUNCH_SEGRED
    return check_launch("map_segment_reduce_rows");
}

extern "C" int map_segment_reduce_rows(const float* rows, int64_t ld_rows, int D, const float* scale, int group,
                                       const int32_t* occ_sorted, const int32_t* seg_start, const int32_t* n_unique,
                                       int64_t n_ids, float* grad_compact, float* scalar_out, map_stream_t stream) {
    return map_segment_reduce_rows_ex(rows, ld_rows, D, scale, group, occ_sorted, seg_start, n_unique, nullptr, n_ids, nullptr, nullptr,
                                      nullptr, 0, 0, grad_compact, scalar_out, stream);
}

extern "C" int map_scatter_rows(const float* grad_compact, const int64_t* uniq_ids, const int32_t* n_unique,
                                int64_t max_unique, int D, float* dense, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(grad_compact && uniq_ids && n_unique && dense && D >= 1, "map_scatter_rows: bad argument");
    int64_t blocks = ceil_div(max_unique * D, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    scatter_rows_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(grad_compact, uniq_ids, n_unique, D, dense);
    return check_launch("map_scatter_rows");
}
