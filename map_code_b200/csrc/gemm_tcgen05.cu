// K3 / K4: TF32 tensor-core GEMM for sm_100a — tcgen05.mma (kind::tf32, cta_group::1) fed by TMA, accumulator in TMEM,
// fused epilogues (bias / ReLU / CrossNet / ReLU-mask / residual) read back with tcgen05.ld.
// Replaces the cuBLAS sgemm behind every nn.Linear of the step (code/layers.py:179-200, code/models.py:116-123), forward
// and backward:   acc[m,n] = sum_k A[m,k] * B[n,k]
//   forward  Y  = X  . W^T      A = X  (K-major)            B = W  [N,K] (K-major)
//   dgrad    dX = dY . W        A = dY (K-major)            B = W  [N,K] read as [K_red=N][N_out=K]  -> trans_b = 1 (MN-major)
//   wgrad    dW = dY^T . X      A = dY [M,N] read transposed -> trans_a = 1 ; B = X [M,K] read transposed -> trans_b = 1
// Both majors are consumed straight from the row-major fp32 tensors: K-major tiles are [rows][32 floats] TMA boxes,
// MN-major tiles are [32 k-rows][32 floats] boxes; all boxes are 128 bytes wide with the 128B swizzle that the UMMA
// shared-memory descriptor names, so no transposed copies and no conversion passes exist anywhere in the step.
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue (TMEM lane quarter = warp%4).
// One 128 x BLOCK_N output tile per CTA (BLOCK_N runtime, multiple of 16, <= 256, chosen on the host so that the tile count
// lands on a multiple of the 148 SMs); 2 CTAs are co-resident per SM whenever the stage ring fits in half the shared
// memory, so one CTA's epilogue overlaps the other's main loop.  Small wgrad grids use split-K with fp32 red.global.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#include "gemm_common.cuh"

namespace mapb {

// ------------------------------------------------------------------------------------------------ the kernel
template <int EPI, bool SPLIT>
__global__ void __launch_bounds__(kGemmThreads) gemm_tf32_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                                         const __grid_constant__ CUtensorMap tmap_b,
                                                                         const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte aligned tile ring (the 128B swizzle pattern is anchored to 1024-byte boundaries)
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * kBlockM;
    const int n0 = blockIdx.x * p.block_n;
    const int kb0 = blockIdx.z * p.num_k_blocks;
    int nkb = p.k_blocks_total - kb0;
    if (nkb > p.num_k_blocks) nkb = p.num_k_blocks;
    unsigned long long* trace = nullptr;
    if (p.trace != nullptr)
        trace = p.trace + (size_t)kTraceWords * (blockIdx.x + gridDim.x * (blockIdx.y + (size_t)gridDim.y * blockIdx.z));
    if (trace != nullptr && threadIdx.x == 0) {
        trace[0] = globaltimer_ns();
        trace[1] = (unsigned long long)clock64();
    }

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation is warp-collective; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // Programmatic dependent launch: everything above (barrier init, tensor-map prefetch, TMEM allocation) may overlap the
    // tail of the previous kernel of the stream; nothing below may start before that kernel has completed and flushed
    // (every global read of this kernel — TMA operand loads and the epilogue's prefetches — is after this point).  The next
    // kernel of the stream is released right away: its CTAs become resident as ours retire and then wait at this same line.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (trace != nullptr && threadIdx.x == 0) trace[2] = (unsigned long long)clock64();

    if (warp == 0) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(kATileBytes + p.b_tile_bytes);
            const int b_chunks = (p.block_n + 31) >> 5;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                const uint32_t round = (uint32_t)(i / p.stages);
                mbar_wait(smem_u32(&empty_bar[s]), (round & 1u) ^ 1u);  // first pass through the ring falls through
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_expect_tx(fb, tx_bytes);
                const uint32_t a_dst = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                const uint32_t b_dst = a_dst + kATileBytes;
                const int k0 = (kb0 + i) * kBlockK;
                if (!p.trans_a) {
                    tma_load_2d(a_dst, &tmap_a, fb, k0, m0);                       // box {32 k, 128 m}
                } else {
#pragma unroll
                    for (int c = 0; c < kBlockM / 32; ++c) tma_load_2d(a_dst + c * 4096, &tmap_a, fb, m0 + c * 32, k0);  // box {32 m, 32 k}
                }
                if (!p.trans_b) {
                    tma_load_2d(b_dst, &tmap_b, fb, k0, n0);                       // box {32 k, block_n n}
                } else {
                    for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_dst + c * 4096, &tmap_b, fb, n0 + c * 32, k0);      // box {32 n, 32 k}
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane) =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.block_n, p.trans_a, p.trans_b);
            // K-major  (SWIZZLE_128B)        : 8-row groups 1024 B apart (SBO); one MMA consumes 32 B of every row
            //                                  -> advance the start address by 32 B per K step
            // MN-major (SWIZZLE_128B_BASE32B): 32-float chunks 4096 B apart (LBO); the swizzle atom is 4 k-rows x 128 B, so
            //                                  4-row groups are 512 B apart (SBO); one MMA consumes 8 k-rows
            //                                  -> advance the start address by 1024 B per K step
            const uint32_t a_lbo = p.trans_a ? 4096u : 16u, a_sbo = p.trans_a ? 512u : 1024u, a_adv = p.trans_a ? 1024u : 32u;
            const uint32_t b_lbo = p.trans_b ? 4096u : 16u, b_sbo = p.trans_b ? 512u : 1024u, b_adv = p.trans_b ? 1024u : 32u;
            const uint32_t a_lt = p.trans_a ? 1u : 2u, b_lt = p.trans_b ? 1u : 2u;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                const uint32_t round = (uint32_t)(i / p.stages);
                mbar_wait(smem_u32(&full_bar[s]), round & 1u);
                tcgen05_fence_after();
                if (trace != nullptr && i == 0) trace[3] = (unsigned long long)clock64();
                const uint32_t a_src = smem_base + (uint32_t)s * (uint32_t)p.stage_bytes;
                const uint32_t b_src = a_src + kATileBytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    const uint64_t da = make_smem_desc(a_src + k * a_adv, a_lbo, a_sbo, a_lt);
                    const uint64_t db = make_smem_desc(b_src + k * b_adv, b_lbo, b_sbo, b_lt);
                    umma_tf32(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
                }
                tcgen05_commit(smem_u32(&empty_bar[s]));  // frees the stage when these MMAs have read it
            }
            tcgen05_commit(smem_u32(&tmem_full_bar));     // accumulator complete
            if (trace != nullptr) trace[4] = (unsigned long long)clock64();
        }
    } else {
        // ===================== epilogue warps: TMEM -> registers -> smem transpose -> fused epilogue -> global =====
        // tcgen05.ld hands every thread one accumulator ROW; writing rows straight out would make each warp store touch 32
        // different lines.  Each warp therefore transposes 32x32 chunks through a private staging tile (the stage ring is
        // idle once tmem_full has fired: one tile per CTA) so that 8 lanes cover 128 contiguous bytes of one output row.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + q * (32 * kStageLd);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row_base = m0 + q * 32;
        EpiRegs e;
        const bool full_tile = (m0 + kBlockM <= p.M) && (n0 + p.block_n <= p.N);
        constexpr int CW = epi_chunk_w(EPI);
        if (p.block_n >= CW) epi_prefetch<EPI, CW>(p, lane, row_base, n0, full_tile, e);
        else epi_prefetch<EPI, 16>(p, lane, row_base, n0, full_tile, e);
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        tcgen05_fence_after();
        if (trace != nullptr && threadIdx.x == 64) trace[5] = (unsigned long long)clock64();
        int c = 0;
        for (; c + CW <= p.block_n; c += CW) {
            epi_chunk<EPI, SPLIT, CW>(p, stg, lane, lane_addr + (uint32_t)c, row_base, n0 + c, full_tile, e);
            if (c + 2 * CW <= p.block_n) epi_prefetch<EPI, CW>(p, lane, row_base, n0 + c + CW, full_tile, e);
            else if (CW == 32 && c + CW < p.block_n) epi_prefetch<EPI, 16>(p, lane, row_base, n0 + c + CW, full_tile, e);
        }
        if (CW == 32 && c < p.block_n)  // block_n % 32 == 16
            epi_chunk<EPI, SPLIT, 16>(p, stg, lane, lane_addr + (uint32_t)c, row_base, n0 + c, full_tile, e);
        if (trace != nullptr && threadIdx.x == 64) trace[6] = (unsigned long long)clock64();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        trace[7] = (globaltimer_ns() & 0xFFFFFFFFFFFFull) | ((unsigned long long)smid << 48);
    }
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ grouped kernel
// One launch = up to kMaxGroup INDEPENDENT GEMM problems (CrossNet layer i and MLP layer i, which both read X0; the dgrad
// and wgrad GEMMs that consume the same upstream gradient).  The grid is the concatenation of the problems' tile grids, one
// 128 x BLOCK_N tile per CTA, two co-resident CTAs per SM exactly like the single-problem kernel — the hardware block
// scheduler keeps every SM slot busy across problem boundaries, and the ~6 us launch / drain gap that each of these small
// GEMMs pays (4096 x 624 x 624 is ~10 us of kernel) is paid once per level of the step's dependency graph.
// The epilogue kind of every slot is a template parameter (E0..E3, -1 = unused) and the slot's parameters are read from the
// kernel parameter space with a STATIC index, so they stay constant-bank operands (a run-time indexed parameter block costs
// ~35 registers per epilogue thread and spills at the 168 registers that two 192-thread CTAs per SM allow).
// (A persistent one-CTA-per-SM variant with two TMEM accumulators was measured slower at these sizes: 2-4 tiles per CTA do
// not amortise its pipeline fill and quantise badly over 148 SMs; profiles/r01g_gemm_group.txt.)
constexpr int kMaxGroup = 4;

struct GroupArgs {
    CUtensorMap tmap[2 * kMaxGroup];  // A, B of slot i at [2i], [2i+1]
    GemmParams p[kMaxGroup];
    int tile_end[kMaxGroup];          // exclusive prefix sums of the per-slot tile counts
    int n_tiles[kMaxGroup];
    int m_tiles[kMaxGroup];
    int count, total_tiles;
    unsigned long long* trace;        // optional (map_gemm_set_trace): same record layout as the single-problem kernel
};

template <bool PAIR, int E0, int E1, int E2, int E3>
__global__ void __launch_bounds__(kGemmThreads, 2) gemm_tf32_group_kernel(const __grid_constant__ GroupArgs g) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // ---- which slot / tile is this CTA (all parameter reads below use static indices)
    const int t = blockIdx.x;
    int gi = 0, start = 0;
#pragma unroll
    for (int i = 0; i < kMaxGroup - 1; ++i)
        if (t >= g.tile_end[i]) {  // unused slots repeat total_tiles
            gi = i + 1;
            start = g.tile_end[i];
        }
    int block_n, trans_a, trans_b, num_k_blocks, k_blocks_total, stages, stage_bytes, b_tile_bytes, tmem_cols, nt, mt;
#define MAP_LOAD_SLOT(i)                                                                                      \
    {                                                                                                         \
        block_n = g.p[i].block_n; trans_a = g.p[i].trans_a; trans_b = g.p[i].trans_b;                        \
        num_k_blocks = g.p[i].num_k_blocks; k_blocks_total = g.p[i].k_blocks_total; stages = g.p[i].stages;   \
        stage_bytes = g.p[i].stage_bytes; b_tile_bytes = g.p[i].b_tile_bytes; tmem_cols = g.p[i].tmem_cols;   \
        nt = g.n_tiles[i]; mt = g.m_tiles[i];                                                                 \
    }
    if (gi == 0) MAP_LOAD_SLOT(0) else if (gi == 1) MAP_LOAD_SLOT(1) else if (gi == 2) MAP_LOAD_SLOT(2) else MAP_LOAD_SLOT(3)
#undef MAP_LOAD_SLOT
    const CUtensorMap* tmap_a = &g.tmap[2 * gi];
    const CUtensorMap* tmap_b = tmap_a + 1;
    const int local = t - start;
    int n_tile, m_tile, ks;
    if (PAIR) {  // m fastest (mt is even): CTAs 2i and 2i+1 of the grid = the two M halves of one 256 x block_n pair tile
        m_tile = local % mt;
        const int r = local / mt;
        n_tile = r % nt;
        ks = r / nt;
    } else {
        n_tile = local % nt;
        const int r = local / nt;
        m_tile = r % mt;
        ks = r / mt;
    }
    const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
    const int half_n = PAIR ? (block_n >> 1) : block_n;   // B columns this CTA stages
    const int m0 = m_tile * kBlockM, n0 = n_tile * block_n, kb0 = ks * num_k_blocks;
    int nkb = k_blocks_total - kb0;
    if (nkb > num_k_blocks) nkb = num_k_blocks;

    unsigned long long* trace = g.trace != nullptr ? g.trace + (size_t)kTraceWords * blockIdx.x : nullptr;
    if (trace != nullptr && threadIdx.x == 0) {
        trace[0] = globaltimer_ns();
        trace[1] = (unsigned long long)clock64();
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(tmap_b) : "memory");
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {  // warp-collective in BOTH CTAs of the pair: the same columns are reserved on both SMs
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)tmem_cols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything of this CTA can arrive on them
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (trace != nullptr && threadIdx.x == 0) trace[2] = (unsigned long long)clock64();

    if (warp == 0) {
        if (lane == 0) {  // ===================== TMA producer =====================
            const uint32_t tx_bytes = (uint32_t)(kATileBytes + b_tile_bytes);
            const int b_chunks = (half_n + 31) >> 5;
            const int nb = n0 + (int)cta_rank * half_n;   // first B column of this CTA's share
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                const uint32_t a_dst = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
                const uint32_t b_dst = a_dst + kATileBytes;
                const int k0 = (kb0 + i) * kBlockK;
                if (PAIR) {
                    // both CTAs' loads complete on the LEADER's full barrier (its MMA reads both shared memories); the leader
                    // expects the bytes of both
                    if (cta_rank == 0) mbar_expect_tx(smem_u32(&full_bar[s]), 2u * tx_bytes);
                    const uint32_t fb = mapa_cta0(smem_u32(&full_bar[s]));
                    if (!trans_a) {
                        tma_load_2d_pair(a_dst, tmap_a, fb, k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 32; ++c) tma_load_2d_pair(a_dst + c * 4096, tmap_a, fb, m0 + c * 32, k0);
                    }
                    if (!trans_b) {
                        tma_load_2d_pair(b_dst, tmap_b, fb, k0, nb);
                    } else {
                        for (int c = 0; c < b_chunks; ++c) tma_load_2d_pair(b_dst + c * 4096, tmap_b, fb, nb + c * 32, k0);
                    }
                } else {
                    const uint32_t fb = smem_u32(&full_bar[s]);
                    mbar_expect_tx(fb, tx_bytes);
                    if (!trans_a) {
                        tma_load_2d(a_dst, tmap_a, fb, k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < kBlockM / 32; ++c) tma_load_2d(a_dst + c * 4096, tmap_a, fb, m0 + c * 32, k0);
                    }
                    if (!trans_b) {
                        tma_load_2d(b_dst, tmap_b, fb, k0, n0);
                    } else {
                        for (int c = 0; c < b_chunks; ++c) tma_load_2d(b_dst + c * 4096, tmap_b, fb, n0 + c * 32, k0);
                    }
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {  // ===================== MMA issuer (the pair's leader) =====================
            const uint32_t idesc = make_idesc_tf32(block_n, trans_a, trans_b, PAIR ? 2 * kBlockM : kBlockM);
            const uint32_t a_lbo = trans_a ? 4096u : 16u, a_sbo = trans_a ? 512u : 1024u, a_adv = trans_a ? 1024u : 32u;
            const uint32_t b_lbo = trans_b ? 4096u : 16u, b_sbo = trans_b ? 512u : 1024u, b_adv = trans_b ? 1024u : 32u;
            const uint32_t a_lt = trans_a ? 1u : 2u, b_lt = trans_b ? 1u : 2u;
            int s = 0;
            uint32_t ph = 0;
            for (int i = 0; i < nkb; ++i) {
                mbar_wait(smem_u32(&full_bar[s]), ph);
                tcgen05_fence_after();
                if (trace != nullptr && i == 0) trace[3] = (unsigned long long)clock64();
                const uint32_t a_src = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
                const uint32_t b_src = a_src + kATileBytes;
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    const uint64_t da = make_smem_desc(a_src + k * a_adv, a_lbo, a_sbo, a_lt);
                    const uint64_t db = make_smem_desc(b_src + k * b_adv, b_lbo, b_sbo, b_lt);
                    if (PAIR) umma_tf32_pair(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
                    else umma_tf32(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
                }
                if (PAIR) tcgen05_commit_pair(smem_u32(&empty_bar[s])); else tcgen05_commit(smem_u32(&empty_bar[s]));
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (PAIR) tcgen05_commit_pair(smem_u32(&tmem_full_bar)); else tcgen05_commit(smem_u32(&tmem_full_bar));
            if (trace != nullptr) trace[4] = (unsigned long long)clock64();
        }
    } else {
        // ===================== epilogue warps (staging tiles alias the ring, idle once tmem_full has fired) =====================
        const int q = warp & 3;
        float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + q * (32 * kStageLd);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row_base = m0 + q * 32;
        const uint32_t fb = smem_u32(&tmem_full_bar);
        if (gi == 0) epi_slot<E0>(g.p[0], stg, lane, lane_addr, row_base, m0, n0, fb, trace);
        else if (gi == 1) epi_slot<E1>(g.p[1], stg, lane, lane_addr, row_base, m0, n0, fb, trace);
        else if (gi == 2) epi_slot<E2>(g.p[2], stg, lane, lane_addr, row_base, m0, n0, fb, trace);
        else epi_slot<E3>(g.p[3], stg, lane, lane_addr, row_base, m0, n0, fb, trace);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        trace[7] = (globaltimer_ns() & 0xFFFFFFFFFFFFull) | ((unsigned long long)smid << 48);
    }
    if (PAIR) cluster_sync_all();  // neither CTA retires (shared memory, barriers, TMEM) while the other may still reference it
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2-D fp32 tensor map: inner (contiguous) extent `inner`, `outer` rows of `ld` floats; box = {32 floats, box_rows}
static int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_rows, bool mn_major) {
    EncodeTiledFn enc = get_encode_fn();
    if (enc == nullptr) {
        set_error("map_gemm_tf32_tcgen05: cuTensorMapEncodeTiled is not available from the driver");
        return MAP_ECUDA;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("map_gemm_tf32_tcgen05: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box_rows=%d", (int)r,
                  (long long)inner, (long long)outer, (long long)ld, box_rows);
        return MAP_ECUDA;
    }
    return MAP_OK;
}

int validate_gemm_args(const map_gemm_args* g, const char* who);  // gemm_simt.cu

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

static int tf32_supported(const map_gemm_args* g, bool set_msg) {
#define UNSUP(...)                       \
    do {                                 \
        if (set_msg) set_error(__VA_ARGS__); \
        return 0;                        \
    } while (0)
    if (g->N % 4 != 0 || g->N < 16) UNSUP("map_gemm_tf32_tcgen05: N=%d must be a multiple of 4 and >= 16", g->N);
    if (g->lda % 4 != 0 || g->ldb % 4 != 0 || g->ldc % 4 != 0) UNSUP("map_gemm_tf32_tcgen05: lda/ldb/ldc must be multiples of 4 floats");
    if (!aligned16(g->A) || !aligned16(g->B) || !aligned16(g->C)) UNSUP("map_gemm_tf32_tcgen05: A/B/C must be 16-byte aligned");
    if (g->bias && !aligned16(g->bias)) UNSUP("map_gemm_tf32_tcgen05: bias must be 16-byte aligned");
    if (g->aux0 && (!aligned16(g->aux0) || g->ld_aux0 % 4 != 0)) UNSUP("map_gemm_tf32_tcgen05: aux0 alignment");
    if (g->aux1 && (!aligned16(g->aux1) || g->ld_aux1 % 4 != 0)) UNSUP("map_gemm_tf32_tcgen05: aux1 alignment");
    if (g->aux_out && (!aligned16(g->aux_out) || g->ld_aux_out % 4 != 0)) UNSUP("map_gemm_tf32_tcgen05: aux_out alignment");
    if (g->aux2 && (!aligned16(g->aux2) || g->ld_aux2 % 4 != 0)) UNSUP("map_gemm_tf32_tcgen05: aux2 alignment");
    if (g->acc_out && (!aligned16(g->acc_out) || g->ld_acc_out % 4 != 0)) UNSUP("map_gemm_tf32_tcgen05: acc_out alignment");
    if (g->colsum_out && !aligned16(g->colsum_out)) UNSUP("map_gemm_tf32_tcgen05: colsum_out must be 16-byte aligned");
#undef UNSUP
    return 1;
}

// Tile / split-K choice.  Measured on B200 (scripts/tune_gemm.py, profiles/r01b_tune_gemm.txt): at M = 4096 every GEMM of the
// step is bound by L2->SM operand traffic, and the best configurations all put ~256-296 CTAs in flight (two co-resident
// CTAs on each of the 148 SMs, one wave) with BLOCK_N <= 160 so that a 3-stage ring fits twice in shared memory.
// Model: time ~ waves * (k_blocks_per_cta * (128 + bn) [operand bytes per K block] + 16 * bn [epilogue]).
struct TileChoice {
    int block_n, split_k;
};
static TileChoice choose_tiles(int M, int N, int K, bool mn_major_b, bool allow_split) {
    const int m_tiles = (int)ceil_div(M, kBlockM);
    const int kb = (int)ceil_div(K, kBlockK);
    const int step = mn_major_b ? 32 : 16;  // MN-major B tiles are made of 32-float chunks
    const int slots = 2 * kNumSMs;
    TileChoice best{128, 1};
    double best_cost = 1e30;
    for (int bn = step; bn <= 160; bn += step) {
        const int n_tiles = (int)ceil_div(N, bn);
        const int tiles = m_tiles * n_tiles;
        int split = 1;
        if (allow_split && kb >= 32 && tiles < slots) {
            split = slots / tiles;
            if (split > kb / 8) split = kb / 8;
            if (split > 16) split = 16;
            if (split < 1) split = 1;
        }
        const int kb_cta = (int)ceil_div(kb, split);
        const int waves = (int)ceil_div((int64_t)tiles * split, slots);
        const double cost = (double)waves * ((double)kb_cta * (128 + bn) + 16.0 * bn) * (split > 1 ? 1.05 : 1.0);
        if (cost < best_cost * 0.999 || (cost < best_cost * 1.001 && bn > best.block_n)) {
            best_cost = cost;
            best = TileChoice{bn, split};
        }
    }
    return best;
}

static unsigned long long* g_trace_buf = nullptr;
static int64_t g_trace_records = 0;

}  // namespace mapb

extern "C" int map_gemm_set_trace(unsigned long long* dev_buf, int64_t capacity_records) {
    mapb::g_trace_buf = dev_buf;
    mapb::g_trace_records = dev_buf ? capacity_records : 0;
    return MAP_OK;
}

extern "C" int map_gemm_tf32_supported(const map_gemm_args* args) {
    if (args == nullptr || args->M <= 0 || args->N <= 0 || args->K <= 0) return 0;
    return mapb::tf32_supported(args, false);
}

extern "C" int map_gemm_tf32_tcgen05(const map_gemm_args* g, map_stream_t stream) {
    using namespace mapb;
    int rc = validate_gemm_args(g, "map_gemm_tf32_tcgen05");
    if (rc != MAP_OK) return rc;
    if (!tf32_supported(g, true)) return MAP_EUNSUPPORTED;
    cudaStream_t st = as_stream(stream);

    GemmParams p{};
    p.M = g->M; p.N = g->N; p.K = g->K;
    p.trans_a = g->trans_a ? 1 : 0;
    p.trans_b = g->trans_b ? 1 : 0;
    const TileChoice tc = choose_tiles(g->M, g->N, g->K, p.trans_b != 0, g->epilogue == MAP_EPI_NONE && g->colsum_out == nullptr);
    p.block_n = tc.block_n;
    if (const char* e = getenv("MAP_B200_BLOCK_N")) {  // tuning override (scripts/tune_gemm.py)
        int bn = atoi(e);
        const int step = p.trans_b ? 32 : 16;
        if (bn >= step && bn <= 256 && bn % step == 0) p.block_n = bn;
    }
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
    p.b_tile_bytes = p.trans_b ? ((p.block_n + 31) / 32) * 4096 : p.block_n * 128;
    p.stage_bytes = (kATileBytes + p.b_tile_bytes + 1023) & ~1023;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.block_n) p.tmem_cols <<= 1;
    p.k_blocks_total = (int)ceil_div(g->K, kBlockK);

    const int m_tiles = (int)ceil_div(g->M, kBlockM);
    const int n_tiles = (int)ceil_div(g->N, p.block_n);
    // split-K for small output grids with a long reduction (wgrad): partial tiles are reduced with 16-byte fp32 reductions
    p.split_k = tc.split_k;
    if (const char* e = getenv("MAP_B200_SPLITK")) {
        const int sk = atoi(e);
        if (sk >= 1 && sk <= 64 && (sk == 1 || g->epilogue == MAP_EPI_NONE)) p.split_k = sk;
    }
    p.num_k_blocks = (int)ceil_div(p.k_blocks_total, p.split_k);
    p.split_k = (int)ceil_div(p.k_blocks_total, p.num_k_blocks);

    // stage ring: aim for two co-resident CTAs (<= ~110 KB each); otherwise one CTA with a deeper ring
    const int budget2 = 110 * 1024, budget1 = 220 * 1024;
    int stages = budget2 / p.stage_bytes;
    if (stages < 3) stages = budget1 / p.stage_bytes;
    if (const char* e = getenv("MAP_B200_STAGES")) {
        const int st_ = atoi(e);
        if (st_ >= 1 && st_ * p.stage_bytes <= budget1) stages = st_;
    }
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages > p.num_k_blocks) stages = p.num_k_blocks > 1 ? p.num_k_blocks : 1;
    if (stages < 1) stages = 1;
    p.stages = stages;
    const size_t smem_bytes = (size_t)stages * p.stage_bytes + 1024;  // + alignment slack

    CUtensorMap tmap_a, tmap_b;
    if (!p.trans_a) rc = make_tmap(&tmap_a, g->A, g->K, g->M, g->lda, kBlockM, false);     // [M rows][K contiguous]
    else rc = make_tmap(&tmap_a, g->A, g->M, g->K, g->lda, kBlockK, true);                // [K rows][M contiguous]
    if (rc != MAP_OK) return rc;
    if (!p.trans_b) rc = make_tmap(&tmap_b, g->B, g->K, g->N, g->ldb, p.block_n, false);   // [N rows][K contiguous]
    else rc = make_tmap(&tmap_b, g->B, g->N, g->K, g->ldb, kBlockK, true);                // [K rows][N contiguous]
    if (rc != MAP_OK) return rc;

    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const GemmParams);
    constexpr int kNumKernels = 10;
    static const KernelFn kernels[kNumKernels] = {
        gemm_tf32_tcgen05_kernel<MAP_EPI_NONE, false>,     gemm_tf32_tcgen05_kernel<MAP_EPI_BIAS, false>,
        gemm_tf32_tcgen05_kernel<MAP_EPI_BIAS_RELU, false>, gemm_tf32_tcgen05_kernel<MAP_EPI_CROSS, false>,
        gemm_tf32_tcgen05_kernel<MAP_EPI_MUL_RELUMASK, false>, gemm_tf32_tcgen05_kernel<MAP_EPI_ADD, false>,
        gemm_tf32_tcgen05_kernel<MAP_EPI_ADD_MUL, false>,  gemm_tf32_tcgen05_kernel<MAP_EPI_CROSS_BWD, false>,
        gemm_tf32_tcgen05_kernel<MAP_EPI_ADD3, false>,     gemm_tf32_tcgen05_kernel<MAP_EPI_NONE, true>};
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < kNumKernels; ++i) {
            if (cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048) != cudaSuccess) {
                set_error("map_gemm_tf32_tcgen05: cannot raise dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
                return MAP_ECUDA;
            }
        }
        attr_set = true;
    }
    MAP_REQUIRE(g->epilogue >= MAP_EPI_NONE && g->epilogue <= MAP_EPI_ADD3, "map_gemm_tf32_tcgen05: unknown epilogue %d", g->epilogue);
    MAP_REQUIRE(p.split_k == 1 || (g->epilogue == MAP_EPI_NONE && g->colsum_out == nullptr),
                "map_gemm_tf32_tcgen05: split-K needs MAP_EPI_NONE without column sums");
    const KernelFn kernel = p.split_k > 1 ? kernels[kNumKernels - 1] : kernels[g->epilogue];
    if (p.split_k > 1) {
        if (cudaMemset2DAsync(g->C, (size_t)g->ldc * sizeof(float), 0, (size_t)g->N * sizeof(float), (size_t)g->M, st) != cudaSuccess) {
            set_error("map_gemm_tf32_tcgen05: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MAP_ECUDA;
        }
    }
    dim3 grid((unsigned)n_tiles, (unsigned)m_tiles, (unsigned)p.split_k);
    p.trace = ((int64_t)n_tiles * m_tiles * p.split_k <= g_trace_records) ? g_trace_buf : nullptr;
    static int use_pdl = -1;
    if (use_pdl < 0) {
        const char* e = getenv("MAP_B200_PDL");
        use_pdl = (e != nullptr && atoi(e) != 0) ? 1 : 0;  // opt-in: measured neutral-to-negative on the step (r01f)
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, kernel, tmap_a, tmap_b, p);
    if (le != cudaSuccess) {
        set_error("map_gemm_tf32_tcgen05: launch failed: %s", cudaGetErrorString(le));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    return check_launch("map_gemm_tf32_tcgen05");
}

// ------------------------------------------------------------------------------------------------ grouped launch (host)
namespace mapb {

typedef void (*GroupKernelFn)(const GroupArgs);
struct GroupCombo {
    int e[kMaxGroup];  // epilogue kinds in DESCENDING order, -1 = unused slot
    GroupKernelFn fn[2];  // [0] one CTA per tile, [1] CTA pairs (cta_group::2)
};
#define MAP_COMBO(a, b, c, d) {{a, b, c, d}, {gemm_tf32_group_kernel<false, a, b, c, d>, gemm_tf32_group_kernel<true, a, b, c, d>}}
// the epilogue combinations the step issues (engine.py): forward level (CrossNet + MLP layer), head dgrads + encoder wgrad,
// backward level (two dgrads + two wgrads), and the dgrad + wgrad pairs of the MLP-only backbones
static const GroupCombo kCombos[] = {
    MAP_COMBO(MAP_EPI_CROSS, MAP_EPI_BIAS_RELU, -1, -1),
    MAP_COMBO(MAP_EPI_CROSS_BWD, MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, -1),
    MAP_COMBO(MAP_EPI_CROSS_BWD, MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, MAP_EPI_NONE),
    MAP_COMBO(MAP_EPI_CROSS_BWD, MAP_EPI_NONE, -1, -1),
    MAP_COMBO(MAP_EPI_ADD3, MAP_EPI_NONE, -1, -1),
    MAP_COMBO(MAP_EPI_ADD3, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    MAP_COMBO(MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, -1, -1),
    MAP_COMBO(MAP_EPI_MUL_RELUMASK, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    MAP_COMBO(MAP_EPI_NONE, MAP_EPI_NONE, -1, -1),
    MAP_COMBO(MAP_EPI_NONE, MAP_EPI_NONE, MAP_EPI_NONE, -1),
    // single problems (pair mode only; without it they go through map_gemm_tf32_tcgen05)
    MAP_COMBO(MAP_EPI_NONE, -1, -1, -1),
    MAP_COMBO(MAP_EPI_BIAS, -1, -1, -1),
    MAP_COMBO(MAP_EPI_BIAS_RELU, -1, -1, -1),
    MAP_COMBO(MAP_EPI_MUL_RELUMASK, -1, -1, -1),
    MAP_COMBO(MAP_EPI_ADD3, -1, -1, -1),
};
#undef MAP_COMBO
constexpr int kNumCombos = (int)(sizeof(kCombos) / sizeof(kCombos[0]));

static int find_combo(const int* e, int count) {
    for (int c = 0; c < kNumCombos; ++c) {
        bool ok = true;
        for (int i = 0; i < kMaxGroup; ++i) ok = ok && kCombos[c].e[i] == (i < count ? e[i] : -1);
        if (ok) return c;
    }
    return -1;
}

// CTA-pair tiles: 256 x block_n per pair.  Operand bytes per K block and CTA are 128 x 32 (A) + block_n/2 x 32 (B half), so wide
// tiles are cheap: the widest block_n in [128, 256] that wastes the fewest padded columns; split-K so that a CTA's share of a
// long reduction (wgrad, K = batch) is about as long as a dgrad tile's (32 K blocks).
static TileChoice choose_tiles_pair(int N, int K, bool mn_major_b, bool allow_split) {
    const int step = mn_major_b ? 32 : 16;
    TileChoice best{step, 1};
    int best_cols = 1 << 30;
    for (int bn = 256; bn >= 128; bn -= step) {
        const int cols = (int)ceil_div(N, bn) * bn;
        if (cols < best_cols) {
            best_cols = cols;
            best.block_n = bn;
        }
    }
    if (N < 128) best.block_n = (int)ceil_div(N, step) * step;
    const int kb = (int)ceil_div(K, kBlockK);
    if (allow_split && kb >= 64) {
        best.split_k = (int)ceil_div(kb, 32);
        if (best.split_k > 16) best.split_k = 16;
    }
    return best;
}

// launches sorted[0..count) (epilogue kinds descending) as ONE grouped kernel; combo = index into kCombos
static int launch_group(const map_gemm_args* sorted, int count, int combo, bool pair, cudaStream_t st) {
    GroupArgs ga{};
    ga.count = count;
    int total = 0, max_smem = 0;
    for (int i = 0; i < count; ++i) {
        const map_gemm_args* g = &sorted[i];
        GemmParams& p = ga.p[i];
        const bool allow_split = g->epilogue == MAP_EPI_NONE && g->colsum_out == nullptr;
        const TileChoice tc = pair ? choose_tiles_pair(g->N, g->K, g->trans_b != 0, allow_split)
                                   : choose_tiles(g->M, g->N, g->K, g->trans_b != 0, allow_split);
        p.M = g->M; p.N = g->N; p.K = g->K;
        p.trans_a = g->trans_a ? 1 : 0;
        p.trans_b = g->trans_b ? 1 : 0;
        p.block_n = tc.block_n;
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
        const int half_n = pair ? p.block_n / 2 : p.block_n;   // B columns staged by one CTA
        p.b_tile_bytes = p.trans_b ? ((half_n + 31) / 32) * 4096 : half_n * 128;
        p.stage_bytes = (kATileBytes + p.b_tile_bytes + 1023) & ~1023;
        p.tmem_cols = 32;
        while (p.tmem_cols < p.block_n) p.tmem_cols <<= 1;
        p.k_blocks_total = (int)ceil_div(g->K, kBlockK);
        p.num_k_blocks = (int)ceil_div(p.k_blocks_total, tc.split_k);
        p.split_k = (int)ceil_div(p.k_blocks_total, p.num_k_blocks);
        int stages = (110 * 1024) / p.stage_bytes;   // two co-resident CTAs
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages > p.num_k_blocks) stages = p.num_k_blocks > 1 ? p.num_k_blocks : 1;
        if (stages < 1) stages = 1;
        p.stages = stages;
        if (stages * p.stage_bytes > max_smem) max_smem = stages * p.stage_bytes;
        p.trace = nullptr;
        ga.m_tiles[i] = (int)ceil_div(g->M, kBlockM);
        if (pair) ga.m_tiles[i] = (ga.m_tiles[i] + 1) & ~1;   // whole pairs: a padding tile is all out of bounds (zero-filled loads, no stores)
        ga.n_tiles[i] = (int)ceil_div(g->N, p.block_n);
        total += ga.m_tiles[i] * ga.n_tiles[i] * p.split_k;
        ga.tile_end[i] = total;
        int rc;
        if (!p.trans_a) rc = make_tmap(&ga.tmap[2 * i], g->A, g->K, g->M, g->lda, kBlockM, false);
        else rc = make_tmap(&ga.tmap[2 * i], g->A, g->M, g->K, g->lda, kBlockK, true);
        if (rc != MAP_OK) return rc;
        if (!p.trans_b) rc = make_tmap(&ga.tmap[2 * i + 1], g->B, g->K, g->N, g->ldb, half_n, false);
        else rc = make_tmap(&ga.tmap[2 * i + 1], g->B, g->N, g->K, g->ldb, kBlockK, true);
        if (rc != MAP_OK) return rc;
        if (p.split_k > 1) {
            if (cudaMemset2DAsync(g->C, (size_t)g->ldc * sizeof(float), 0, (size_t)g->N * sizeof(float), (size_t)g->M, st) != cudaSuccess) {
                set_error("map_gemm_tf32_group: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MAP_ECUDA;
            }
        }
    }
    for (int i = count; i < kMaxGroup; ++i) ga.tile_end[i] = total;
    ga.total_tiles = total;
    if (max_smem < 4 * 32 * kStageLd * 4) max_smem = 4 * 32 * kStageLd * 4;  // the epilogue staging tiles alias the ring
    const size_t smem_bytes = (size_t)max_smem + 1024;
    const GroupKernelFn fn = kCombos[combo].fn[pair ? 1 : 0];
    static bool attr_set[kNumCombos][2] = {};
    if (!attr_set[combo][pair ? 1 : 0]) {
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048) != cudaSuccess) {
            set_error("map_gemm_tf32_group: cannot raise dynamic shared memory limit: %s", cudaGetErrorString(cudaGetLastError()));
            return MAP_ECUDA;
        }
        attr_set[combo][pair ? 1 : 0] = true;
    }
    ga.trace = (total <= g_trace_records) ? g_trace_buf : nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)total);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, fn, ga);
    if (le != cudaSuccess) {
        set_error("map_gemm_tf32_group: launch failed: %s", cudaGetErrorString(le));
        cudaGetLastError();
        return MAP_ECUDA;
    }
    return check_launch("map_gemm_tf32_group");
}

// CTA pairs pay off when every problem has at least one full pair of M tiles and a tile-wide N
static bool want_pair(const map_gemm_args* a, int count) {
    static int enabled = -1;
    if (enabled < 0) {
        const char* e = getenv("MAP_B200_GEMM_PAIR");
        enabled = (e == nullptr || atoi(e) != 0) ? 1 : 0;
    }
    if (!enabled) return false;
    for (int i = 0; i < count; ++i)
        if (a[i].M < 2 * kBlockM || a[i].N < 64) return false;
    return true;
}

}  // namespace mapb

extern "C" int map_gemm_tf32_group(const map_gemm_args* args, int count, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(args != nullptr && count >= 1 && count <= kMaxGroup, "map_gemm_tf32_group: count=%d must be in [1, %d]", count, kMaxGroup);
    cudaStream_t st = as_stream(stream);
    for (int i = 0; i < count; ++i) {
        int rc = validate_gemm_args(&args[i], "map_gemm_tf32_group");
        if (rc != MAP_OK) return rc;
        if (!tf32_supported(&args[i], true)) return MAP_EUNSUPPORTED;
        MAP_REQUIRE(args[i].epilogue >= MAP_EPI_NONE && args[i].epilogue <= MAP_EPI_ADD3, "map_gemm_tf32_group: unknown epilogue %d", args[i].epilogue);
    }
    // canonical order: epilogue kind descending (stable), which is how the instantiated combinations are listed
    map_gemm_args sorted[kMaxGroup];
    for (int i = 0; i < count; ++i) sorted[i] = args[i];
    for (int i = 1; i < count; ++i)
        for (int j = i; j > 0 && sorted[j].epilogue > sorted[j - 1].epilogue; --j) {
            const map_gemm_args tmp = sorted[j]; sorted[j] = sorted[j - 1]; sorted[j - 1] = tmp;
        }
    // greedy cover: the longest prefix of what is left that is an instantiated combination goes out as one grouped launch,
    // a problem no combination starts with goes through the single-problem kernel
    int pos = 0;
    while (pos < count) {
        int e[kMaxGroup], len = 0, combo = -1;
        for (int n = count - pos; n >= 1 && combo < 0; --n) {
            for (int i = 0; i < n; ++i) e[i] = sorted[pos + i].epilogue;
            combo = find_combo(e, n);
            len = n;
        }
        const bool pair = combo >= 0 && want_pair(&sorted[pos], len);
        int rc;
        if (combo >= 0 && (len >= 2 || pair)) {
            rc = launch_group(&sorted[pos], len, combo, pair, st);
            pos += len;
        } else {
            rc = map_gemm_tf32_tcgen05(&sorted[pos], stream);
            pos += 1;
        }
        if (rc != MAP_OK) return rc;
    }
    return MAP_OK;
}
