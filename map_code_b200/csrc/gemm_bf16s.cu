// K3 / K4 (default backend): split-bf16 tensor-core GEMM for sm_100a with fp32-level accuracy.
//
// Why not TF32: tcgen05 kind::tf32 truncates operands to 10 mantissa bits.  On this path the forward pre-activations feed
// ReLUs, and d(loss)/d(weights) is a DISCONTINUOUS function of them: a pre-activation whose sign differs from the reference's
// switches a whole unit's gradient on or off.  Measured at the benchmarked shape (B=4096, H=1000; profiles/r02a_parity_*):
// TF32 -> 3e-2 Frobenius error on the MLP weight gradients (north_star asks for 1e-3), exact fp32 with another summation
// order -> 2e-4..8e-4 (the reference's own reproducibility floor).  Every fp32 operand is therefore carried as bf16 PLANES
//     x = hi + lo (+ lo2),  hi = bf16(x), lo = bf16(x - hi), lo2 = bf16(x - hi - lo)
// written once by whoever produces the tensor (GEMM epilogues, the gather, AdamW), and a GEMM sums several bf16 MMAs into one
// fp32 TMEM accumulator:
//     terms = 3:  hi*hi + hi*lo + lo*hi                      (relative error ~2^-17 per product; all linear uses)
//     terms = 6:  + lo*lo + hi*lo2 + lo2*hi                  (~2^-24: the forward GEMMs whose outputs feed a ReLU)
// The tensor core's fp32 accumulator rounds toward zero: measured on B200, every MMA that lands in an accumulator costs ~2e-8 of its
// magnitude, systematically (terms = 6, K = 1624: 624 MMAs -> 1.1e-5, WORSE than terms = 3).  With terms = 6 the five small
// correction products therefore go to a SECOND TMEM accumulator (its rounding is relative to a 2^-8 times smaller sum) and only
// the K/16 hi*hi MMAs touch the main one; the epilogue adds the two in fp32.
// bf16 keeps fp32's exponent range, so there is no scaling and no overflow mode.  Per k-block of 64 elements a CTA stages 2-3
// planes of A and of its half of B (TMA, 128B swizzle) and the pair's leader issues terms x 4 MMAs (kind::f16, K = 16).
//
// Shape of the kernel (one launch = up to 4 INDEPENDENT problems = one level of the step's dependency graph):
//   * persistent: one CTA PAIR (cta_group::2, 256 x BLOCK_N tile, each CTA stages its 128 rows of A and half of B) per pair
//     of SMs walks the concatenated tile list of all problems;
//   * warp-specialised: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer (leader CTA), warps 2-5 epilogue;
//   * two TMEM accumulators (2 x 256 columns): the epilogue of tile i (TMEM -> registers -> smem transpose -> fused
//     bias / ReLU / CrossNet / ReLU-mask / residual -> global fp32 + bf16 planes for the next GEMM) overlaps the MMAs of tile i+1;
//   * operands in both majors straight from row-major storage: K-major tiles are {64 k, rows} boxes, MN-major tiles are
//     {64 mn, 64 k} boxes (dgrad reads W [N,K] as B^T, wgrad reads dY and X transposed), no transposed copies anywhere.
// Replaces the cuBLAS sgemm behind every nn.Linear of the step (code/layers.py:179-200, code/models.py:116-123), fwd + bwd.
#define MAP_GEMM_TWO_ACCUMULATORS 1
#define MAP_GEMM_STAGE_LD16 1
#include "gemm_common.cuh"

namespace mapb {

constexpr int kSBlockK = 64;               // bf16 elements per k-block = 128 bytes = one swizzle span
constexpr int kSUmmaK = 16;                // kind::f16: 32 bytes of K per tcgen05.mma
constexpr int kSPlaneA = kBlockM * 128;    // 16 KiB: one plane of a CTA's A tile
constexpr int kSChunk = 64 * 128;          // 8 KiB: one {64 mn, 64 k} box of an MN-major tile
constexpr int kSMaxStages = 6;
constexpr int kSMaxGroup = 4;
constexpr int kSAccCols = 256;             // TMEM columns per accumulator
constexpr int kSTmemCols = 2 * kSAccCols;
constexpr int kSEpiWarps = 8;              // two per TMEM lane quarter: the column chunks of a tile alternate between them
constexpr int kSThreads = 64 + 32 * kSEpiWarps;
constexpr int kSStagingBytes = kSEpiWarps * 32 * kStageLd16 * 4;   // per-warp transpose tiles (16-column chunks)
constexpr int kSSmemBudget = 227 * 1024 - 2048;

struct SProblem {
    GemmParams p;          // b_tile_bytes = bytes of ONE plane of this CTA's half of the B tile
    int pa, pb, terms;     // planes staged of A / B; MMAs per k-step (3 or 6; 6 = two accumulators, see above)
    int half_n;            // B columns staged by one CTA (block_n / 2)
    int n_tiles, m_pairs;  // tile grid (pair tiles)
    int n_pair_tiles;      // m_pairs * n_tiles * split_k
};

struct SGroupArgs {
    CUtensorMap tmap[2 * kSMaxGroup];   // A, B of problem i at [2i], [2i+1]; 3-D {inner, rows, plane}
    SProblem pr[kSMaxGroup];
    int count, total_tiles;
    // tile sequence of the launch: the problems in DESCENDING order of their per-tile cost (seq[i] = slot, seq_end[i] = exclusive
    // prefix sum of tiles); cluster c takes sequence positions c, 2n-1-c, 2n+c, 4n-1-c, ... (snake), so every cluster gets
    // a mix of expensive and cheap tiles and the launch ends within about one cheap tile of the average
    int seq[kSMaxGroup], seq_end[kSMaxGroup];
    int stages, stage_bytes;            // one ring geometry for the whole launch (the largest problem's planes)
    unsigned long long* trace;          // optional per-CTA timeline (map_gemm_bf16s_set_trace), nullptr in production
    // dynamic tile scheduler: sched[0] = next sequence position to hand out, sched[1] = clusters that have run dry.  Both are 0
    // between launches (the last cluster to run dry resets them), so a captured launch can be replayed without a memset node.
    // nullptr = static snake order (MAP_B200_GEMM_SCHED=static, A/B)
    unsigned* sched;
};

// trace record of one CTA: 64 x u64 = {globaltimer at entry, clock after setup, globaltimer at exit, smid | tiles << 32,
// then for each of its first 10 tiles: clock of the first TMA issue (producer), clock when the MMA issuer got the accumulator
// buffer, clock when the first stage had landed, clock when the last MMA was issued, clock when the accumulator was complete
// (epilogue warp 2), clock at the end of warp 2's epilogue}
constexpr int kSTraceWords = 64;
constexpr int kSTraceTiles = 10;

struct TileInfo {
    int gi, m_pair, n_tile, ks;
    int block_n, half_n, trans_a, trans_b, pa, pb, terms, nkb, kb0, b_plane_bytes;
    // two-accumulator tile (terms = 6) wider than 128 columns: main and corrections take one 256-column buffer each, so the tile
    // owns ALL of TMEM and its epilogue cannot overlap the next tile's MMAs.  Up to 128 columns both live in ONE buffer
    // (corrections 128 columns after the main accumulator) and the tile double-buffers like a single-accumulator one.
    bool wide_dual;
};

// TMEM accumulator buffers: a single-accumulator tile takes ONE of the two 256-column buffers (alternating, so its epilogue
// overlaps the next tile's MMAs), a two-accumulator tile takes BOTH (main in buffer 0, corrections in buffer 1).  MMA issuer and
// epilogue warps run the same bookkeeping: uses[b] = tiles that have used buffer b so far (its parity drives the barriers).
struct AccState {
    uint32_t uses[2];
    uint32_t next;
};
__device__ __forceinline__ uint32_t acc_pick(AccState& st, bool dual) {
    if (dual) { st.next = 0; return 0; }
    const uint32_t b = st.next;
    st.next ^= 1u;
    return b;
}

__device__ __forceinline__ TileInfo decode_tile(const SGroupArgs& g, int t) {
    TileInfo ti;
    int sp = 0, start = 0;
#pragma unroll
    for (int i = 0; i < kSMaxGroup - 1; ++i)
        if (t >= g.seq_end[i]) {
            sp = i + 1;
            start = g.seq_end[i];
        }
    const int gi = g.seq[sp];
    int nt, mp, nkb_split, kbt;
#define MAP_LOAD_PR(i)                                                                                              \
    {                                                                                                               \
        ti.block_n = g.pr[i].p.block_n; ti.half_n = g.pr[i].half_n; ti.trans_a = g.pr[i].p.trans_a;                 \
        ti.trans_b = g.pr[i].p.trans_b; ti.pa = g.pr[i].pa; ti.pb = g.pr[i].pb; ti.terms = g.pr[i].terms;          \
        ti.b_plane_bytes = g.pr[i].p.b_tile_bytes; nt = g.pr[i].n_tiles; mp = g.pr[i].m_pairs;                      \
        nkb_split = g.pr[i].p.num_k_blocks; kbt = g.pr[i].p.k_blocks_total;                                         \
    }
    if (gi == 0) MAP_LOAD_PR(0) else if (gi == 1) MAP_LOAD_PR(1) else if (gi == 2) MAP_LOAD_PR(2) else MAP_LOAD_PR(3)
#undef MAP_LOAD_PR
    const int local = t - start;
    ti.gi = gi;
    ti.m_pair = local % mp;          // m fastest: the pairs that run side by side share one B (weight) tile
    const int r = local / mp;
    ti.n_tile = r % nt;
    ti.ks = r / nt;
    ti.kb0 = ti.ks * nkb_split;
    int nkb = kbt - ti.kb0;
    if (nkb > nkb_split) nkb = nkb_split;
    ti.nkb = nkb;
    ti.wide_dual = ti.terms == 6 && ti.block_n > kSAccCols / 2;
    return ti;
}

// i-th tile of cluster c (of n): sequence position in snake order, or -1 when the cluster has no i-th tile
__device__ __forceinline__ int snake_tile(int i, int c, int n, int total) {
    const int t = i * n + ((i & 1) ? (n - 1 - c) : c);
    return t < total ? t : -1;
}

// ---- tile scheduler.  The pair's leader CTA decides which sequence position the cluster works on next and publishes it through
// a small ring (one copy in each CTA's shared memory) to every role of both CTAs: its own TMA producer, the peer's producer, the
// MMA issuer, the 8 + 8 epilogue warps.  Dynamic mode: positions come from a global counter in the launch's sequence order
// (descending per-tile cost = longest-processing-time-first list scheduling), so a cluster that starts late — its SMs were still
// held by a kernel of another stream — or that drew tiles with long epilogues simply takes fewer tiles; the static snake order
// made the whole launch wait for the unluckiest cluster (r02c traces: per-CTA lifetime median 86 us, max 112 us on the 4-problem
// backward levels).  -1 closes the sequence.
constexpr int kSchedDepth = 4;
constexpr int kSchedConsumers = 2 * kSEpiWarps + 2;   // leader: MMA issuer + epilogue warps; peer: TMA producer + epilogue warps

__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquires what a thread of the OTHER CTA released
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// one lane of the (converged) warp; the same lane every time
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
// consumer side: the j-th tile of this cluster (all calling threads get it); `arrive` = this thread reports the slot as read.
// `leader`: the caller runs in the leader CTA — publisher and barrier are in its own shared memory, CTA-scope ordering is enough
// (cluster-scope acquire / release on every tile boundary showed up as ~2k clk in the per-tile traces)
__device__ __forceinline__ void sched_release(const uint64_t* empty, int j, bool leader) {
    const uint32_t bar = smem_u32(&empty[j % kSchedDepth]);
    if (leader) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    else mbar_arrive_cta0(bar);
}
__device__ __forceinline__ int sched_read(const uint64_t* full, const uint64_t* empty, const int* tiles, int j, bool arrive, bool leader) {
    const int slot = j % kSchedDepth;
    if (leader) mbar_wait_quiet(smem_u32(&full[slot]), (uint32_t)(j / kSchedDepth) & 1u);
    else mbar_wait_cluster(smem_u32(&full[slot]), (uint32_t)(j / kSchedDepth) & 1u);
    const int t = *reinterpret_cast<const volatile int*>(&tiles[slot]);
    if (arrive) sched_release(empty, j, leader);
    return t;
}
// leader's producer: publish the cluster's j-th tile to both CTAs
__device__ __forceinline__ void sched_publish(uint64_t* full, uint64_t* empty, int* tiles, int j, int t) {
    const int slot = j % kSchedDepth;
    if (j >= kSchedDepth) mbar_wait_quiet(smem_u32(&empty[slot]), (uint32_t)(j / kSchedDepth - 1) & 1u);
    *reinterpret_cast<volatile int*>(&tiles[slot]) = t;
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_rank(smem_u32(&tiles[slot]), 1u)), "r"(t) : "memory");
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_rank(smem_u32(&full[slot]), 0u)) : "memory");
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_rank(smem_u32(&full[slot]), 1u)) : "memory");
}

template <int E0, int E1, int E2, int E3>
__global__ void __launch_bounds__(kSThreads, 1) gemm_bf16s_kernel(const __grid_constant__ SGroupArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t full_bar[kSMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kSMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(8) uint64_t sched_full[kSchedDepth];
    __shared__ __align__(8) uint64_t sched_empty[kSchedDepth];   // the leader's copy is the one in use
    __shared__ int sched_tile[kSchedDepth];
    __shared__ int sched_count;   // tiles this cluster processed (trace only)

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // (the shuffle makes it provably warp-uniform)
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const int cluster_id = (int)(blockIdx.x >> 1);
    const int n_clusters = (int)(gridDim.x >> 1);
    const int stages = g.stages;
    const uint32_t stage_bytes = (uint32_t)g.stage_bytes;
    const int total = g.total_tiles;
    unsigned long long* trace = g.trace != nullptr ? g.trace + (size_t)kSTraceWords * blockIdx.x : nullptr;
    if (trace != nullptr && threadIdx.x == 0) trace[0] = globaltimer_ns();

    // the cluster's first sequence position is requested before the setup below: the atomic's round trip hides behind it
    int t_first = 0;
    if (warp == 0 && lane == 0 && cta_rank == 0)
        t_first = g.sched != nullptr ? (int)atomicAdd(g.sched, 1u) : snake_tile(0, cluster_id, n_clusters, total);
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 2 * g.count; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmap[i]) : "memory");
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&tmem_full_bar[b]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[b]), 2 * kSEpiWarps);   // the epilogue warps of both CTAs of the pair (leader's barrier)
        }
        for (int d = 0; d < kSchedDepth; ++d) {
            mbar_init(smem_u32(&sched_full[d]), 1);
            mbar_init(smem_u32(&sched_empty[d]), kSchedConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // warp-collective in BOTH CTAs: the same columns are reserved on both SMs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)kSTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything of this CTA can arrive on them
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (trace != nullptr && threadIdx.x == 0) trace[1] = (unsigned long long)clock64();

    if (warp == 0) {
        {   // ===================== TMA producer (each CTA: its A rows, its half of B) =====================
            // whole warp, warp-uniform control flow, one elected lane issues (see the MMA issuer below: UTMALDG takes its operands from
            // uniform registers too)
            int s = 0;
            uint32_t ph = 0;
            // leader: source of the cluster's tile sequence (global counter or snake order); t_next = position of the tile after this one
            const bool dynamic = g.sched != nullptr;
            int t_next = __shfl_sync(0xffffffffu, t_first, 0);
            for (int tj = 0;; ++tj) {
                int t;
                if (cta_rank == 0) {
                    t = (t_next >= 0 && t_next < total) ? t_next : -1;
                    if (lane == 0) sched_publish(sched_full, sched_empty, sched_tile, tj, t);
                    __syncwarp();
                } else {
                    t = sched_read(sched_full, sched_empty, sched_tile, tj, lane == 0, false);
                    t = __shfl_sync(0xffffffffu, t, 0);
                }
                if (t < 0) break;
                const TileInfo ti = decode_tile(g, t);
                // the next position is requested a few k-blocks before this tile's loads end: late enough that a cluster does not
                // sit on a tile it will not start for a long time, early enough to hide the atomic's round trip (the value stays in
                // lane 0 until the loads are out: broadcasting it at once would stall the warp on the atomic)
                const int fetch_at = ti.nkb > 3 ? ti.nkb - 3 : 0;
                int fetched = -1;
                const CUtensorMap* tmap_a = &g.tmap[2 * ti.gi];
                const CUtensorMap* tmap_b = tmap_a + 1;
                const int m0 = (ti.m_pair * 2 + (int)cta_rank) * kBlockM;
                const int nb = ti.n_tile * ti.block_n + (int)cta_rank * ti.half_n;
                const uint32_t tx_bytes = (uint32_t)(ti.pa * kSPlaneA + ti.pb * ti.b_plane_bytes);
                const int b_chunks = ti.half_n >> 6;
                for (int i = 0; i < ti.nkb; ++i) {
                    if (cta_rank == 0 && i == fetch_at && lane == 0)
                        fetched = dynamic ? (int)atomicAdd(g.sched, 1u) : snake_tile(tj + 1, cluster_id, n_clusters, total);
                    mbar_wait_quiet(smem_u32(&empty_bar[s]), ph ^ 1u);
                    __syncwarp();
                    if (trace != nullptr && i == 0 && tj < kSTraceTiles && lane == 0) trace[4 + 6 * tj] = (unsigned long long)clock64();
                    const uint32_t fb = mapa_cta0(smem_u32(&full_bar[s]));
                    const uint32_t a_dst = smem_base + (uint32_t)s * stage_bytes;
                    const uint32_t b_dst = a_dst + (uint32_t)(ti.pa * kSPlaneA);
                    const int k0 = (ti.kb0 + i) * kSBlockK;
                    if (elect_one()) {
                        // both CTAs' loads complete on the LEADER's full barrier (its MMA reads both shared memories)
                        if (cta_rank == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2u * tx_bytes);
                        for (int pl = 0; pl < ti.pa; ++pl) {
                            if (!ti.trans_a) {
                                tma_load_3d_pair(a_dst + pl * kSPlaneA, tmap_a, fb, k0, m0, pl);                      // box {64 k, 128 m}
                            } else {
#pragma unroll
                                for (int c = 0; c < kBlockM / 64; ++c)
                                    tma_load_3d_pair(a_dst + pl * kSPlaneA + c * kSChunk, tmap_a, fb, m0 + c * 64, k0, pl);   // box {64 m, 64 k}
                            }
                        }
                        for (int pl = 0; pl < ti.pb; ++pl) {
                            if (!ti.trans_b) {
                                tma_load_3d_pair(b_dst + pl * ti.b_plane_bytes, tmap_b, fb, k0, nb, pl);              // box {64 k, half_n n}
                            } else {
                                for (int c = 0; c < b_chunks; ++c)
                                    tma_load_3d_pair(b_dst + pl * ti.b_plane_bytes + c * kSChunk, tmap_b, fb, nb + c * 64, k0, pl);  // box {64 n, 64 k}
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                t_next = __shfl_sync(0xffffffffu, fetched, 0);
            }
            if (cta_rank == 0 && dynamic && lane == 0) {   // this cluster has drawn its one position past the end; the last one to do so re-arms the counters
                if (atomicAdd(g.sched + 1, 1u) == (unsigned)(n_clusters - 1)) {
                    g.sched[0] = 0u;
                    g.sched[1] = 0u;
                    __threadfence();
                }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0) {  // ===================== MMA issuer (the pair's leader) =====================
            // The WHOLE warp walks the loop with warp-uniform control flow and one elected lane issues: tcgen05.mma / commit take
            // their descriptors from uniform registers, and inside an `if (lane == 0)` region the compiler cannot prove uniformity —
            // it wrapped every MMA in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY loop (~170 clk per MMA issued, r02c traces: the
            // main loop ran at the issue rate of this one thread, not at the tensor pipe's 128 clk).  Everything the descriptors
            // depend on is made provably uniform: the warp index and the values read from shared memory go through a shuffle.
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            int s = 0;
            uint32_t ph = 0;
            AccState acc{{0u, 0u}, 0u};
            for (int tj = 0;; ++tj) {
                int t = sched_read(sched_full, sched_empty, sched_tile, tj, lane == 0, true);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t < 0) break;
                const TileInfo ti = decode_tile(g, t);
                const bool dual = ti.terms == 6;
                const bool wide = ti.wide_dual;
                const uint32_t buf = acc_pick(acc, wide);
                // the epilogue of the tile that last used this accumulator has drained it
                mbar_wait_quiet(smem_u32(&tmem_empty_bar[buf]), (acc.uses[buf] & 1u) ^ 1u);
                if (wide) mbar_wait_quiet(smem_u32(&tmem_empty_bar[1]), (acc.uses[1] & 1u) ^ 1u);
                tcgen05_fence_after();
                if (trace != nullptr && tj < kSTraceTiles && lane == 0) trace[5 + 6 * tj] = (unsigned long long)clock64();
                const uint32_t idesc = make_idesc_bf16(ti.block_n, ti.trans_a, ti.trans_b, 2 * kBlockM);
                // K-major  (SWIZZLE_128B): 8-row groups 1024 B apart (SBO); one MMA consumes 32 B of every row -> +32 B per K step
                // MN-major (SWIZZLE_128B): 64-element chunks 8192 B apart (LBO); 8-k-row groups 1024 B apart (SBO); one MMA
                //                          consumes 16 k-rows -> +2048 B per K step
                // descriptor = constant high word | low word {start >> 4, LBO >> 4 << 16}: only the start address moves
                const uint32_t a_lbo = ti.trans_a ? (uint32_t)kSChunk : 16u, a_adv16 = ti.trans_a ? (2048u >> 4) : (32u >> 4);
                const uint32_t b_lbo = ti.trans_b ? (uint32_t)kSChunk : 16u, b_adv16 = ti.trans_b ? (2048u >> 4) : (32u >> 4);
                const uint32_t d_hi = (uint32_t)(make_smem_desc(0u, 0u, 1024u, 2u) >> 32);
                const uint32_t a_lo0 = (uint32_t)make_smem_desc(0u, a_lbo, 0u, 0u), b_lo0 = (uint32_t)make_smem_desc(0u, b_lbo, 0u, 0u);
                const uint32_t d_tmem = tmem_u + buf * (uint32_t)kSAccCols;
                // two-accumulator tiles: corrections in buffer 1 (wide) or in the upper half of the tile's own buffer
                const uint32_t d_corr = wide ? tmem_u + (uint32_t)kSAccCols : d_tmem + (uint32_t)(kSAccCols / 2);
                const uint32_t a_planes16 = (uint32_t)(ti.pa * kSPlaneA) >> 4;
                const uint32_t b_plane16 = (uint32_t)ti.b_plane_bytes >> 4;
                const int terms = ti.terms;
                for (int i = 0; i < ti.nkb; ++i) {
                    mbar_wait_quiet(smem_u32(&full_bar[s]), ph);
                    tcgen05_fence_after();
                    __syncwarp();
                    if (trace != nullptr && i == 0 && tj < kSTraceTiles && lane == 0) trace[6 + 6 * tj] = (unsigned long long)clock64();
                    const uint32_t a16 = (smem_base + (uint32_t)s * stage_bytes) >> 4;   // (all offsets are multiples of 16 bytes)
                    const uint32_t b16 = a16 + a_planes16;
                    if (elect_one()) {
                        // products ordered (a plane, b plane): (0,0) (0,1) (1,0) | (1,1) (0,2) (2,0)
                        for (int term = 0; term < terms; ++term) {
                            const uint32_t ia = (term == 2 || term == 3) ? 1u : (term == 5 ? 2u : 0u);
                            const uint32_t ib = (term == 1 || term == 3) ? 1u : (term == 4 ? 2u : 0u);
                            const uint32_t a_pl = a16 + ia * (uint32_t)(kSPlaneA >> 4);
                            const uint32_t b_pl = b16 + ib * b_plane16;
                            const uint32_t d = (dual && term > 0) ? d_corr : d_tmem;
                            const uint32_t first = (dual && term > 0) ? (uint32_t)(i | (term - 1)) : (uint32_t)(i | term);
#pragma unroll
                            for (int k = 0; k < kSBlockK / kSUmmaK; ++k) {
                                const uint64_t da = ((uint64_t)d_hi << 32) | (uint64_t)(a_lo0 | ((a_pl + k * a_adv16) & 0x3FFFu));
                                const uint64_t db = ((uint64_t)d_hi << 32) | (uint64_t)(b_lo0 | ((b_pl + k * b_adv16) & 0x3FFFu));
                                umma_bf16_pair(d, da, db, idesc, (first | (uint32_t)k) != 0u ? 1u : 0u);
                            }
                        }
                        tcgen05_commit_pair(smem_u32(&empty_bar[s]));   // frees the stage in both CTAs when these MMAs have read it
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                if (elect_one()) {
                    tcgen05_commit_pair(smem_u32(&tmem_full_bar[buf]));  // accumulator complete (both CTAs' epilogues)
                    if (wide) tcgen05_commit_pair(smem_u32(&tmem_full_bar[1]));
                }
                __syncwarp();
                if (trace != nullptr && tj < kSTraceTiles && lane == 0) trace[7 + 6 * tj] = (unsigned long long)clock64();
                ++acc.uses[buf];
                if (wide) ++acc.uses[1];
            }
        }
    } else {
        // ===================== epilogue warps (both CTAs): TMEM -> registers -> smem transpose -> fused epilogue -> global =====
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;  // which of the quarter's two warps: takes every second 16-column chunk
        float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)stages * stage_bytes) + (warp - 2) * (32 * kStageLd16);
        AccState acc{{0u, 0u}, 0u};
        int tj = 0;
        for (;; ++tj) {
            const int t = sched_read(sched_full, sched_empty, sched_tile, tj, false, cta_rank == 0);
            __syncwarp();   // every lane has read the slot before lane 0 hands it back
            if (lane == 0) sched_release(sched_empty, tj, cta_rank == 0);
            if (t < 0) {
                if (threadIdx.x == 64) sched_count = tj;
                break;
            }
            const TileInfo ti = decode_tile(g, t);
            const bool dual = ti.wide_dual;   // (buffer bookkeeping only: the epilogue finds the corrections through p.corr_cols)
            const uint32_t buf = acc_pick(acc, dual);
            unsigned long long* ttr = (trace != nullptr && tj < kSTraceTiles) ? trace + 4 + 6 * tj : nullptr;   // epi_tile stamps [5], [6]
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)kSAccCols;
            const int m0 = (ti.m_pair * 2 + (int)cta_rank) * kBlockM;
            const int n0 = ti.n_tile * ti.block_n;
            const int row_base = m0 + q * 32;
            const uint32_t fb = smem_u32(&tmem_full_bar[buf]);
            const uint32_t par = acc.uses[buf] & 1u;
            if (dual) {   // the corrections' barrier completes with the same commit sequence: consume its phase too
                mbar_wait(smem_u32(&tmem_full_bar[1]), acc.uses[1] & 1u);
            }
            // (epi_tile stamps ttr[5] when the accumulator is complete and ttr[6] at its end: words 8 and 9 (+6 per tile) of the record
            //  are written through the pointer shifted by -1 so that they land on [4 + 6 tj + 4] and [+5])
            unsigned long long* tt = ttr != nullptr ? ttr - 1 : nullptr;
            if (ti.gi == 0) epi_slot<E0, 16>(g.pr[0].p, stg, lane, lane_addr, row_base, m0, n0, fb, tt, par, part, 2);
            else if (ti.gi == 1) epi_slot<E1, 16>(g.pr[1].p, stg, lane, lane_addr, row_base, m0, n0, fb, tt, par, part, 2);
            else if (ti.gi == 2) epi_slot<E2, 16>(g.pr[2].p, stg, lane, lane_addr, row_base, m0, n0, fb, tt, par, part, 2);
            else epi_slot<E3, 16>(g.pr[3].p, stg, lane, lane_addr, row_base, m0, n0, fb, tt, par, part, 2);
            // this warp has read its quarter of the accumulator: hand the buffer back to the MMA issuer (leader CTA's barrier)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cta0(smem_u32(&tmem_empty_bar[buf]));
                if (dual) mbar_arrive_cta0(smem_u32(&tmem_empty_bar[1]));
            }
            ++acc.uses[buf];
            if (dual) ++acc.uses[1];
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        int my_tiles = 0;   // (the closing -1 has been published by now: the ring still holds the last kSchedDepth entries)
        if (g.sched == nullptr) {
            while (snake_tile(my_tiles, cluster_id, n_clusters, total) >= 0) ++my_tiles;
        } else {
            my_tiles = *reinterpret_cast<volatile int*>(&sched_count);
        }
        trace[2] = globaltimer_ns();
        trace[3] = (unsigned long long)smid | ((unsigned long long)my_tiles << 32);
    }
    cluster_sync_all();  // neither CTA retires (shared memory, barriers, TMEM) while the other may still reference it
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kSTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ split kernel
// fp32 [rows, cols] (row stride ld) -> n_planes bf16 planes [rows, ld_p]: hi, lo, lo2 (see the header of this file)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols4,
                                                         __nv_bfloat16* __restrict__ planes, int64_t ld_p, int64_t plane_stride, int n_planes) {
    const int64_t total = rows * cols4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols4;
        const int64_t c = (i - r * cols4) * 4;
        const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
        float r0 = v.x, r1 = v.y, r2 = v.z, r3 = v.w;
        __nv_bfloat16* dst = planes + r * ld_p + c;
        for (int pl = 0; pl < n_planes; ++pl) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(r0), h1 = __float2bfloat16_rn(r1), h2 = __float2bfloat16_rn(r2), h3 = __float2bfloat16_rn(r3);
            r0 -= __bfloat162float(h0); r1 -= __bfloat162float(h1); r2 -= __bfloat162float(h2); r3 -= __bfloat162float(h3);
            uint2 pk;
            pk.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pk.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            *reinterpret_cast<uint2*>(dst) = pk;
            dst += plane_stride;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFnS get_encode_fn_s() {
    static EncodeTiledFnS fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnS>(sym);
    }
    return fn;
}

// 3-D bf16 tensor map over the planes of one operand: {inner (contiguous), outer rows, plane}; box = {64, box_rows, 1}
static int make_tmap_planes(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int64_t plane_stride,
                            int n_planes, int box_rows) {
    EncodeTiledFnS enc = get_encode_fn_s();
    if (enc == nullptr) {
        set_error("map_gemm_bf16s: cuTensorMapEncodeTiled is not available from the driver");
        return MAP_ECUDA;
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)n_planes};
    const cuuint64_t gstride[2] = {(cuuint64_t)ld * 2u, (cuuint64_t)plane_stride * 2u};
    const cuuint32_t box[3] = {(cuuint32_t)kSBlockK, (cuuint32_t)box_rows, 1u};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("map_gemm_bf16s: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld planes=%d box_rows=%d", (int)r,
                  (long long)inner, (long long)outer, (long long)ld, n_planes, box_rows);
        return MAP_ECUDA;
    }
    return MAP_OK;
}

static bool aligned16s(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int bf16s_check(const map_gemm_split_args* a) {
    const map_gemm_args* g = &a->g;
    MAP_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "map_gemm_bf16s: bad shape M=%d N=%d K=%d", g->M, g->N, g->K);
    MAP_REQUIRE(g->N % 4 == 0, "map_gemm_bf16s: N=%d must be a multiple of 4", g->N);
    MAP_REQUIRE(a->a_planes && a->b_planes && g->C, "map_gemm_bf16s: null operand");
    MAP_REQUIRE(a->terms == 3 || a->terms == 6, "map_gemm_bf16s: terms=%d must be 3 or 6", a->terms);
    const int need = a->terms == 6 ? 3 : 2;
    MAP_REQUIRE(a->a_nplanes >= need && a->b_nplanes >= need, "map_gemm_bf16s: terms=%d needs %d planes of A and B (have %d, %d)", a->terms,
                need, a->a_nplanes, a->b_nplanes);
    MAP_REQUIRE(a->a_ld % 8 == 0 && a->b_ld % 8 == 0 && a->a_plane_stride % 8 == 0 && a->b_plane_stride % 8 == 0 && aligned16s(a->a_planes) &&
                    aligned16s(a->b_planes),
                "map_gemm_bf16s: operand planes need 16-byte aligned bases and row / plane strides that are multiples of 8 elements");
    MAP_REQUIRE(g->ldc % 4 == 0 && aligned16s(g->C), "map_gemm_bf16s: C alignment");
    MAP_REQUIRE(!g->bias || aligned16s(g->bias), "map_gemm_bf16s: bias alignment");
    MAP_REQUIRE(!g->aux0 || (aligned16s(g->aux0) && g->ld_aux0 % 4 == 0), "map_gemm_bf16s: aux0 alignment");
    MAP_REQUIRE(!g->aux1 || (aligned16s(g->aux1) && g->ld_aux1 % 4 == 0), "map_gemm_bf16s: aux1 alignment");
    MAP_REQUIRE(!g->aux2 || (aligned16s(g->aux2) && g->ld_aux2 % 4 == 0), "map_gemm_bf16s: aux2 alignment");
    MAP_REQUIRE(!g->aux_out || (aligned16s(g->aux_out) && g->ld_aux_out % 4 == 0), "map_gemm_bf16s: aux_out alignment");
    MAP_REQUIRE(!g->acc_out || (aligned16s(g->acc_out) && g->ld_acc_out % 4 == 0), "map_gemm_bf16s: acc_out alignment");
    MAP_REQUIRE(!g->colsum_out || aligned16s(g->colsum_out), "map_gemm_bf16s: colsum_out alignment");
    MAP_REQUIRE(g->epilogue >= MAP_EPI_NONE && g->epilogue <= MAP_EPI_ADD3, "map_gemm_bf16s: unknown epilogue %d", g->epilogue);
    if (a->c_planes != nullptr) {
        MAP_REQUIRE(a->c_nplanes >= 1 && a->c_nplanes <= 3 && a->c_ld % 4 == 0 && ((uintptr_t)a->c_planes & 7) == 0 && a->c_plane_stride % 4 == 0,
                    "map_gemm_bf16s: C planes need 8-byte aligned base, strides multiple of 4 elements, 1..3 planes");
    }
    return MAP_OK;
}

typedef void (*SKernelFn)(const SGroupArgs);
struct SCombo {
    int e[kSMaxGroup];  // epilogue kinds in DESCENDING order, -1 = unused slot
    SKernelFn fn;
};
#define MAP_SCOMBO(a, b, c, d) {{a, b, c, d}, gemm_bf16s_kernel<a, b, c, d>}
// the epilogue combinations the step issues (engine.py), plus every single problem
static const SCombo kSCombos[] = {
    MAP_SCOMBO(MAP_EPI_CROSS, MAP_EPI_BIAS_RELU, -1, -1),
    MAP_SCOMBO(MAP_EPI_CROSS_BWD, MAP_EPI_MUL_RELUMASK, -1, -1),   // head level of the by-field MFP encoder (its weight gradient is not a GEMM)
    MAP_SCOMBO(MAP_EPI_CROSS_BWD, MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, -1),
    MAP_SCOMBO(MAP_EPI_CROSS_BWD, MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, MAP_EPI_NONE),
    MAP_SCOMBO(MAP_EPI_CROSS_BWD, MAP_EPI_NONE, -1, -1),
    MAP_SCOMBO(MAP_EPI_ADD3, MAP_EPI_NONE, -1, -1),
    MAP_SCOMBO(MAP_EPI_ADD3, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    MAP_SCOMBO(MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, -1, -1),
    MAP_SCOMBO(MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    MAP_SCOMBO(MAP_EPI_NONE, MAP_EPI_NONE, -1, -1),
    MAP_SCOMBO(MAP_EPI_NONE, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    MAP_SCOMBO(MAP_EPI_NONE, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_BIAS, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_BIAS_RELU, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_CROSS, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_MUL_RELUMASK, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_ADD, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_ADD_MUL, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_CROSS_BWD, -1, -1, -1),
    MAP_SCOMBO(MAP_EPI_ADD3, -1, -1, -1),
};
#undef MAP_SCOMBO
constexpr int kNumSCombos = (int)(sizeof(kSCombos) / sizeof(kSCombos[0]));

static int find_scombo(const int* e, int count) {
    for (int c = 0; c < kNumSCombos; ++c) {
        bool ok = true;
        for (int i = 0; i < kSMaxGroup; ++i) ok = ok && kSCombos[c].e[i] == (i < count ? e[i] : -1);
        if (ok) return c;
    }
    return -1;
}

// ---- tile plan of one launch.  Measured (profiles/r02b_gemm_bf16s_levels.txt): in steady state the main loop is bound by L2 -> SM
// operand traffic (~30 B/clk/SM at the clocks this kernel runs at), not by the tensor pipe; one level of the step is only 1.5 - 3
// pair tiles per cluster, so the END of a launch is decided by tile quantisation.  The host therefore picks every problem's
// BLOCK_N by simulating the launch: per-tile cost = k-blocks x max(operand bytes / 30, MMA clocks) + the exposed part of the
// epilogue, tiles dealt to the 74 clusters in the kernel's own order (problems by descending tile cost, snake), makespan =
// busiest cluster.  K-major B takes any multiple of 16 (<= 256); MN-major B whole 64-column chunks per CTA half: 128 or 256.
struct TileModel {
    int block_n, n_tiles, m_pairs, split, nkb, tiles;
    double cost;      // time a cluster is busy with one tile when its epilogue overlaps the next tile's main loop
    double latency;   // main loop + the whole epilogue: what a tile costs when it is a cluster's LAST one
};
static const double kEpiCol[] = {15, 23, 25, 45, 60, 45, 60, 150, 47};   // epilogue clk per column: NONE, BIAS, BIAS_RELU, CROSS, RELUMASK, ADD, ADD_MUL, CROSS_BWD, ADD3

static int split_for(const map_gemm_split_args* a, int kb_total) {
    const map_gemm_args* g = &a->g;
    int split = 1;
    const bool allow_split = g->epilogue == MAP_EPI_NONE && g->colsum_out == nullptr && a->c_planes == nullptr;
    if (allow_split && kb_total >= 32) {
        split = (kb_total + 8) / 16;
        if (split > 16) split = 16;
    }
    if (const char* e = getenv("MAP_B200_SPLITK")) {
        const int sk = atoi(e);
        if (allow_split && sk >= 1 && sk <= 64) split = sk;
    }
    return split;
}

static TileModel model_tile(const map_gemm_split_args* a, int bn) {
    const map_gemm_args* g = &a->g;
    TileModel t{};
    t.block_n = bn;
    t.m_pairs = (int)ceil_div(g->M, 2 * kBlockM);
    t.n_tiles = (int)ceil_div(g->N, bn);
    const int kbt = (int)ceil_div(g->K, kSBlockK);
    t.split = split_for(a, kbt);
    t.nkb = (int)ceil_div(kbt, t.split);
    t.split = (int)ceil_div(kbt, t.nkb);
    t.tiles = t.m_pairs * t.n_tiles * t.split;
    const int planes = a->terms == 6 ? 3 : 2;
    const double bytes = planes * (kSPlaneA + bn / 2 * 128.0);
    const double mma = a->terms * 4 * 128.0 * bn / 256.0;
    const double per_kb = bytes / 30.0 > mma ? bytes / 30.0 : mma;
    const double epi = kEpiCol[g->epilogue] * bn + (a->c_planes ? 10.0 * a->c_nplanes * bn : 0.0);
    const bool wide_dual = a->terms == 6 && bn > kSAccCols / 2;   // owns all of TMEM: its whole epilogue is exposed
    t.cost = t.nkb * per_kb * (a->terms == 6 ? 1.1 : 1.0) + (wide_dual ? epi : 0.35 * epi) + 1500.0;
    t.latency = t.nkb * per_kb * (a->terms == 6 ? 1.1 : 1.0) + epi + 1500.0;
    return t;
}

// MAP_B200_GEMM_SCHED=static: snake order instead of the global tile counter (A/B)
static bool sched_dynamic() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("MAP_B200_GEMM_SCHED");
        v = (e != nullptr && strcmp(e, "static") == 0) ? 0 : 1;
    }
    return v != 0;
}

// position of a problem's tiles in the launch's sequence (descending).  Static snake: per-tile busy time.  Dynamic: the tile's whole
// latency — longest-processing-time-first, and the tiles that close the launch are the ones with the shortest exposed epilogue
// (split-K weight gradients: plain stores) instead of a CrossNet-backward tile with six streamed operands
static double seq_key(const TileModel& t, bool dynamic) { return dynamic ? t.latency : t.cost; }

static double simulate_makespan(const TileModel* tm, int count, int clusters) {
    int order[kSMaxGroup] = {0, 1, 2, 3};
    const bool dynamic = sched_dynamic();
    for (int i = 1; i < count; ++i)
        for (int j = i; j > 0 && seq_key(tm[order[j]], dynamic) > seq_key(tm[order[j - 1]], dynamic); --j) {
            const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp;
        }
    double load[kNumSMs / 2] = {};
    int pos = 0;
    for (int oi = 0; oi < count; ++oi) {
        const TileModel& t = tm[order[oi]];
        for (int k = 0; k < t.tiles; ++k, ++pos) {
            if (dynamic) {   // list scheduling: the next tile of the sequence goes to the cluster that runs dry first
                int c = 0;
                for (int i = 1; i < clusters; ++i) c = load[i] < load[c] ? i : c;
                load[c] += t.cost;
            } else {
                const int round = pos / clusters, c = pos % clusters;
                load[(round & 1) ? clusters - 1 - c : c] += t.cost;
            }
        }
    }
    double mx = 0;
    for (int c = 0; c < clusters; ++c) mx = load[c] > mx ? load[c] : mx;
    return mx;
}

struct PlanKey {
    int v[kSMaxGroup][8];
    int count, clusters;
};
struct PlanEntry {
    PlanKey key;
    int bn[kSMaxGroup];
};
static PlanEntry g_plan_cache[64];
static int g_plan_cache_n = 0;

static void plan_block_n_search(const map_gemm_split_args* sorted, int count, int clusters, int* out_bn);

// the search result only depends on the shapes: remembered per (shapes, epilogues, terms) signature (eager steps relaunch the same
// nine groups every step)
static void plan_block_n(const map_gemm_split_args* sorted, int count, int clusters, int* out_bn) {
    PlanKey k;
    memset(&k, 0, sizeof(k));
    k.count = count;
    k.clusters = clusters;
    for (int i = 0; i < count; ++i) {
        const map_gemm_args* g = &sorted[i].g;
        const int v[8] = {g->M, g->N, g->K, g->trans_a, g->trans_b, g->epilogue, sorted[i].terms,
                          (sorted[i].c_planes ? sorted[i].c_nplanes : 0) | (g->colsum_out ? 16 : 0)};
        memcpy(k.v[i], v, sizeof(v));
    }
    for (int e = 0; e < g_plan_cache_n; ++e)
        if (memcmp(&g_plan_cache[e].key, &k, sizeof(k)) == 0) {
            memcpy(out_bn, g_plan_cache[e].bn, sizeof(int) * kSMaxGroup);
            return;
        }
    plan_block_n_search(sorted, count, clusters, out_bn);
    if (getenv("MAP_B200_BLOCK_N") == nullptr && getenv("MAP_B200_SPLITK") == nullptr) {
        PlanEntry& e = g_plan_cache[g_plan_cache_n < 64 ? g_plan_cache_n++ : 63];
        e.key = k;
        memcpy(e.bn, out_bn, sizeof(int) * kSMaxGroup);
    }
}

static void plan_block_n_search(const map_gemm_split_args* sorted, int count, int clusters, int* out_bn) {
    int cand[kSMaxGroup][16], ncand[kSMaxGroup];
    for (int i = 0; i < kSMaxGroup; ++i) out_bn[i] = 0;
    for (int i = 0; i < count; ++i) {
        const map_gemm_args* g = &sorted[i].g;
        ncand[i] = 0;
        static const bool dual_wide = getenv("MAP_B200_DUAL_WIDE") != nullptr && atoi(getenv("MAP_B200_DUAL_WIDE")) != 0;
        // two-accumulator problems: at most 128 columns, so that main + corrections fit one TMEM buffer and the epilogue overlaps the
        // next tile's main loop (MAP_B200_DUAL_WIDE=1 lifts the limit: A/B)
        const int bn_max = (sorted[i].terms == 6 && !dual_wide) ? kSAccCols / 2 : 256;
        if (g->trans_b) {
            if (bn_max == 256) cand[i][ncand[i]++] = 256;
            cand[i][ncand[i]++] = 128;
        } else if (g->N <= 64) {
            cand[i][ncand[i]++] = (int)ceil_div(g->N, 16) * 16;
        } else {
            for (int bn = bn_max; bn >= 64; bn -= 16) {   // one candidate per distinct tile count: the narrowest width that still covers N
                const int nt = (int)ceil_div(g->N, bn);
                const int tight = (int)ceil_div((int)ceil_div(g->N, nt), 16) * 16;
                bool seen = false;
                for (int k = 0; k < ncand[i]; ++k) seen = seen || cand[i][k] == tight;
                if (!seen && ncand[i] < 16) cand[i][ncand[i]++] = tight;
            }
        }
        if (const char* e = getenv("MAP_B200_BLOCK_N")) {  // tuning override
            const int bn = atoi(e);
            if (bn >= 16 && bn <= 256 && bn % (g->trans_b ? 128 : 16) == 0) {
                cand[i][0] = bn;
                ncand[i] = 1;
            }
        }
    }
    int idx[kSMaxGroup] = {0, 0, 0, 0}, best[kSMaxGroup] = {0, 0, 0, 0};
    double best_ms = 1e300;
    for (;;) {
        TileModel tm[kSMaxGroup];
        for (int i = 0; i < count; ++i) tm[i] = model_tile(&sorted[i], cand[i][idx[i]]);
        const double ms = simulate_makespan(tm, count, clusters);
        if (ms < best_ms) {
            best_ms = ms;
            for (int i = 0; i < count; ++i) best[i] = idx[i];
        }
        int d = 0;
        while (d < count && ++idx[d] == ncand[d]) idx[d++] = 0;
        if (d == count) break;
    }
    for (int i = 0; i < count; ++i) out_bn[i] = cand[i][best[i]];
}

static unsigned long long* g_strace_buf = nullptr;
static int64_t g_strace_words = 0;

// counters of the dynamic tile scheduler: one {next, done} pair per launch, handed out round-robin.  A pair is 0 whenever no launch
// is using it (the kernel re-arms it), so captured launches replay without a memset; kSchedPairs launches later the pair is reused.
// Captured launches (replayed for the life of their graph) and eager launches draw from separate halves, so an eager launch on one
// stream can never share a pair with a graph replay that is in flight on another.
constexpr int kSchedPairs = 1024;
static unsigned* g_sched_buf = nullptr;
static int g_sched_pos[2] = {0, 0};   // [0] eager, [1] under stream capture

static int sched_counters(cudaStream_t st, unsigned** out) {
    *out = nullptr;
    if (!sched_dynamic()) return MAP_OK;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (g_sched_buf == nullptr) {
        MAP_REQUIRE(cs == cudaStreamCaptureStatusNone,
                    "map_gemm_bf16s: the first launch of the process allocates the tile-scheduler counters and cannot be captured; run it once eagerly");
        if (cudaMalloc(&g_sched_buf, kSchedPairs * 2 * sizeof(unsigned)) != cudaSuccess ||
            cudaMemset(g_sched_buf, 0, kSchedPairs * 2 * sizeof(unsigned)) != cudaSuccess) {
            set_error("map_gemm_bf16s: cannot allocate the tile-scheduler counters: %s", cudaGetErrorString(cudaGetLastError()));
            g_sched_buf = nullptr;
            return MAP_ECUDA;
        }
    }
    const int half = cs == cudaStreamCaptureStatusNone ? 0 : 1;
    *out = g_sched_buf + 2 * (half * (kSchedPairs / 2) + (g_sched_pos[half]++ % (kSchedPairs / 2)));
    return MAP_OK;
}

static int launch_sgroup(const map_gemm_split_args* sorted, int count, int combo, cudaStream_t st) {
    SGroupArgs ga{};
    ga.count = count;
    int total = 0, stage_bytes = 0;
    double cost[kSMaxGroup] = {0, 0, 0, 0};   // sequence key of every problem (seq_key)
    int clusters = kNumSMs / 2;
    if (const char* e = getenv("MAP_B200_GEMM_CLUSTERS")) {
        const int c = atoi(e);
        if (c >= 1 && c <= kNumSMs / 2) clusters = c;
    }
    int plan_bn[kSMaxGroup];
    plan_block_n(sorted, count, clusters, plan_bn);
    for (int i = 0; i < count; ++i) {
        const map_gemm_split_args* a = &sorted[i];
        const map_gemm_args* g = &a->g;
        SProblem& pr = ga.pr[i];
        GemmParams& p = pr.p;
        p.M = g->M; p.N = g->N; p.K = g->K;
        p.trans_a = g->trans_a ? 1 : 0;
        p.trans_b = g->trans_b ? 1 : 0;
        p.block_n = plan_bn[i];
        pr.half_n = p.block_n / 2;
        pr.terms = a->terms;
        pr.pa = pr.pb = (a->terms == 6) ? 3 : 2;
        p.epilogue = g->epilogue;
        p.C = g->C; p.ldc = g->ldc;
        p.bias = g->bias;
        p.aux0 = g->aux0; p.ld_aux0 = g->ld_aux0;
        p.aux1 = g->aux1; p.ld_aux1 = g->ld_aux1;
        p.aux_out = g->aux_out; p.ld_aux_out = g->ld_aux_out;
        p.aux2 = g->aux2; p.ld_aux2 = g->ld_aux2;
        p.acc_out = g->acc_out; p.ld_acc_out = g->ld_acc_out;
        p.acc_accumulate = g->acc_accumulate;
        p.colsum_out = g->colsum_out;
        p.corr_cols = (a->terms == 6) ? (p.block_n > kSAccCols / 2 ? kSAccCols : kSAccCols / 2) : 0;
        p.c_planes = reinterpret_cast<__nv_bfloat16*>(a->c_planes);
        p.ld_cp = a->c_ld; p.cp_stride = a->c_plane_stride; p.pc = a->c_planes ? a->c_nplanes : 0;
        p.b_tile_bytes = p.trans_b ? (pr.half_n / 64) * kSChunk : pr.half_n * 128;
        p.b_tile_bytes = (p.b_tile_bytes + 1023) & ~1023;   // planes start on swizzle-pattern boundaries
        const int sb = pr.pa * kSPlaneA + pr.pb * p.b_tile_bytes;
        if (sb > stage_bytes) stage_bytes = sb;
        p.tmem_cols = kSAccCols;
        p.k_blocks_total = (int)ceil_div(g->K, kSBlockK);
        // split-K (plain outputs only): a CTA pair's share of a long reduction (wgrad, K = batch) is ~16 k-blocks, like a dgrad tile
        const int split = split_for(a, p.k_blocks_total);
        p.num_k_blocks = (int)ceil_div(p.k_blocks_total, split);
        p.split_k = (int)ceil_div(p.k_blocks_total, p.num_k_blocks);
        p.trace = nullptr;
        pr.m_pairs = (int)ceil_div(g->M, 2 * kBlockM);
        pr.n_tiles = (int)ceil_div(g->N, p.block_n);
        pr.n_pair_tiles = pr.m_pairs * pr.n_tiles * p.split_k;
        total += pr.n_pair_tiles;
        cost[i] = seq_key(model_tile(a, p.block_n), sched_dynamic());
        int rc;
        // A planes: K-major [M rows][K contiguous] box {64, 128}; MN-major [K rows][M contiguous] box {64, 64}
        if (!p.trans_a) rc = make_tmap_planes(&ga.tmap[2 * i], a->a_planes, g->K, g->M, a->a_ld, a->a_plane_stride, pr.pa, kBlockM);
        else rc = make_tmap_planes(&ga.tmap[2 * i], a->a_planes, g->M, g->K, a->a_ld, a->a_plane_stride, pr.pa, kSBlockK);
        if (rc != MAP_OK) return rc;
        if (!p.trans_b) rc = make_tmap_planes(&ga.tmap[2 * i + 1], a->b_planes, g->K, g->N, a->b_ld, a->b_plane_stride, pr.pb, pr.half_n);
        else rc = make_tmap_planes(&ga.tmap[2 * i + 1], a->b_planes, g->N, g->K, a->b_ld, a->b_plane_stride, pr.pb, kSBlockK);
        if (rc != MAP_OK) return rc;
        if (p.split_k > 1) {
            if (cudaMemset2DAsync(g->C, (size_t)g->ldc * sizeof(float), 0, (size_t)g->N * sizeof(float), (size_t)g->M, st) != cudaSuccess) {
                set_error("map_gemm_bf16s: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MAP_ECUDA;
            }
        }
    }
    {
        int order[kSMaxGroup] = {0, 1, 2, 3};
        for (int i = 1; i < count; ++i)
            for (int j = i; j > 0 && cost[order[j]] > cost[order[j - 1]]; --j) {
                const int tmp = order[j]; order[j] = order[j - 1]; order[j - 1] = tmp;
            }
        int acc_tiles = 0;
        for (int i = 0; i < kSMaxGroup; ++i) {
            ga.seq[i] = i < count ? order[i] : 0;
            if (i < count) acc_tiles += ga.pr[order[i]].n_pair_tiles;
            ga.seq_end[i] = acc_tiles;
        }
    }
    ga.total_tiles = total;
    ga.stage_bytes = stage_bytes;
    int stages = (kSSmemBudget - 1024 - kSStagingBytes) / stage_bytes;
    if (stages > kSMaxStages) stages = kSMaxStages;
    if (const char* e = getenv("MAP_B200_STAGES")) {
        const int s_ = atoi(e);
        if (s_ >= 1 && s_ <= stages) stages = s_;
    }
    MAP_REQUIRE(stages >= 2, "map_gemm_bf16s: stage of %d bytes does not fit twice in shared memory", stage_bytes);
    ga.stages = stages;
    const size_t smem_bytes = (size_t)stages * stage_bytes + kSStagingBytes + 1024;
    const SKernelFn fn = kSCombos[combo].fn;
    static bool attr_set[kNumSCombos] = {};
    if (!attr_set[combo]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSSmemBudget) != cudaSuccess) {
            set_error("map_gemm_bf16s: cannot raise dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return MAP_ECUDA;
        }
        attr_set[combo] = true;
    }
    if (clusters > total) clusters = total;
    ga.trace = ((int64_t)2 * clusters * kSTraceWords <= g_strace_words) ? g_strace_buf : nullptr;
    {
        const int rc = sched_counters(st, &ga.sched);
        if (rc != MAP_OK) return rc;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kSThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, fn, ga);
    if (le != cudaSuccess) {
        set_error("map_gemm_bf16s: launch failed: %s", cudaGetErrorString(le));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    return check_launch("map_gemm_bf16s_group");
}

}  // namespace mapb

extern "C" int map_gemm_bf16s_set_trace(unsigned long long* dev_buf, int64_t capacity_u64) {
    mapb::g_strace_buf = dev_buf;
    mapb::g_strace_words = dev_buf ? capacity_u64 : 0;
    return MAP_OK;
}

extern "C" int map_split_bf16(const float* src, int64_t ld, int64_t rows, int64_t cols, uint16_t* planes, int64_t ld_p, int64_t plane_stride,
                              int n_planes, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(src && planes && rows > 0 && cols > 0 && n_planes >= 1 && n_planes <= 3, "map_split_bf16: bad argument");
    MAP_REQUIRE(cols % 4 == 0 && ld % 4 == 0 && ld_p % 4 == 0 && plane_stride % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)planes & 7) == 0,
                "map_split_bf16: cols / strides must be multiples of 4 elements, src 16-byte and planes 8-byte aligned");
    const int64_t total = rows * (cols / 4);
    int64_t blocks = ceil_div(total, 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, ld, rows, cols / 4, reinterpret_cast<__nv_bfloat16*>(planes), ld_p,
                                                                        plane_stride, n_planes);
    return check_launch("map_split_bf16");
}

extern "C" int map_gemm_bf16s_group(const map_gemm_split_args* args, int count, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(args != nullptr && count >= 1 && count <= kSMaxGroup, "map_gemm_bf16s_group: count=%d must be in [1, %d]", count, kSMaxGroup);
    cudaStream_t st = as_stream(stream);
    for (int i = 0; i < count; ++i) {
        const int rc = bf16s_check(&args[i]);
        if (rc != MAP_OK) return rc;
    }
    // canonical order: epilogue kind descending (stable), which is how the instantiated combinations are listed
    map_gemm_split_args sorted[kSMaxGroup];
    for (int i = 0; i < count; ++i) sorted[i] = args[i];
    for (int i = 1; i < count; ++i)
        for (int j = i; j > 0 && sorted[j].g.epilogue > sorted[j - 1].g.epilogue; --j) {
            const map_gemm_split_args tmp = sorted[j]; sorted[j] = sorted[j - 1]; sorted[j - 1] = tmp;
        }
    // greedy cover: the longest prefix of what is left that is an instantiated combination goes out as one launch
    int pos = 0;
    while (pos < count) {
        int e[kSMaxGroup], len = 0, combo = -1;
        for (int n = count - pos; n >= 1 && combo < 0; --n) {
            for (int i = 0; i < n; ++i) e[i] = sorted[pos + i].g.epilogue;
            combo = find_scombo(e, n);
            len = n;
        }
        MAP_REQUIRE(combo >= 0, "map_gemm_bf16s_group: no kernel for epilogue %d", sorted[pos].g.epilogue);
        const int rc = launch_sgroup(&sorted[pos], len, combo, st);
        if (rc != MAP_OK) return rc;
        pos += len;
    }
    return MAP_OK;
}
