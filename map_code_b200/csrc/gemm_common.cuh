// Shared pieces of the tcgen05 GEMM kernels (gemm_tcgen05.cu: TF32 single pass; gemm_bf16s.cu: split-bf16 multi pass):
// PTX wrappers (mbarrier, TMA, tcgen05), the per-problem parameter block and the fused epilogue (TMEM -> registers -> per-warp
// shared-memory transpose -> bias / ReLU / CrossNet / ReLU-mask / residual -> coalesced global stores).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mapb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;               // 32 fp32 = 128 bytes = one swizzle span
constexpr int kUmmaK = 8;                 // tf32: 32 bytes of K per tcgen05.mma
constexpr int kGemmThreads = 192;
constexpr int kATileBytes = kBlockM * kBlockK * 4;  // 16 KiB
constexpr int kMaxStages = 8;

struct GemmParams {
    int M, N, K;
    int block_n;          // multiple of 16
    int trans_a, trans_b;
    int num_k_blocks;     // per split
    int k_blocks_total;
    int stages;
    int stage_bytes;      // A tile + B tile, multiple of 1024
    int b_tile_bytes;
    int tmem_cols;        // power of two >= block_n, >= 32
    int epilogue;
    int split_k;
    float* C; int64_t ldc;
    const float* bias;
    const float* aux0; int64_t ld_aux0;
    const float* aux1; int64_t ld_aux1;
    float* aux_out; int64_t ld_aux_out;
    const float* aux2; int64_t ld_aux2;
    float* acc_out; int64_t ld_acc_out;
    int acc_accumulate;
    float* colsum_out;
    unsigned long long* trace;  // optional per-CTA phase timestamps (map_gemm_set_trace), nullptr in production
    // split-bf16 kernel only: the epilogue also writes C as `pc` bf16 planes (hi, lo[, lo2]: C = hi + lo (+ lo2) to 2^-17 / 2^-25),
    // the operand format of the GEMMs that consume C next (gemm_bf16s.cu); plane i at c_planes + i * cp_stride, row stride ld_cp
    __nv_bfloat16* c_planes; int64_t ld_cp; int64_t cp_stride; int pc;
    // split-bf16 kernel, terms = 6: the correction products are summed in a SECOND accumulator `corr_cols` TMEM columns after the
    // main one (0 = single accumulator); the epilogue adds the two in fp32
    int corr_cols;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// trace record of one CTA: 8 x u64 = {globaltimer at entry, clock at entry, clock after setup, clock when the first stage
// landed (MMA warp), clock when the last MMA was issued, clock when the accumulator was complete (epilogue warp 2), clock at
// the end of warp 2's epilogue, globaltimer at exit | smid << 48}
constexpr int kTraceWords = 8;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// A mis-programmed pipeline must not hang the GPU: after ~2 s of SM clocks the CTA traps (the launch then reports an
// error through the normal CUDA error path) instead of spinning forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("map_gemm_tf32_tcgen05: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}
// same without the message (a printf call site makes every live register caller-saved around it)
__device__ __forceinline__ void mbar_wait_quiet(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {  // arrives on `bar` when all prior tcgen05.mma retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one MMA of M = 256 spans the two SMs of a TPC; each CTA stages its own 128 rows of A and
// HALF of the B tile, so a 128 x N output tile per CTA costs (128 + N/2) x 32 floats per K block instead of (128 + N) x 32
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_cta0(uint32_t addr) {  // shared::cluster address of `addr` in CTA 0 of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the transaction bytes are credited to `bar_cluster`, a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {  // arrives on `bar` of BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], M = 128, N = block_n, K = 8 (tf32)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// UMMA shared-memory descriptor (PTX ISA "tcgen05 matrix descriptor"): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46)
// | version=1 [46,48) | layout_type [61,64): 2 = SWIZZLE_128B (16-byte atoms; K-major tiles),
//                                           1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only legal layout for MN-major
//                                               32-bit operands: the tensor core transposes at element granularity)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor: D=f32 [4,6)=1 | A=tf32 [7,10)=2 | B=tf32 [10,13)=2 | a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(int block_n, int a_mn_major, int b_mn_major, int m = kBlockM) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ epilogue
// Per warp and per chunk of W (32 or 16) accumulator columns: lane = (row-in-group rr, 4-column group cg); the warp walks its
// 32 rows in ITERS steps of RPI rows, so that W/4 lanes cover W*4 contiguous bytes of one output row (full 128-byte lines at
// W = 32).  Everything the fused epilogue READS from global memory (bias, aux0, aux1) is fetched into registers by
// epi_prefetch BEFORE the accumulator chunk is pulled out of TMEM: the loads of all rows are in flight together (one L2
// round trip per chunk instead of one per row), and chunk 0 is prefetched while the main loop is still running.
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4_nc(const float* p) {  // read-only for the lifetime of the kernel, streamed once
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ---- split-bf16 kernel (gemm_bf16s.cu): 3-D tensor maps {inner, rows, plane}, kind::f16 MMAs of CTA pairs
__device__ __forceinline__ void tma_load_3d_pair(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// instruction descriptor, kind::f16: D=f32 [4,6)=1 | A=bf16 [7,10)=1 | B=bf16 [10,13)=1 | a_major [15] | b_major [16] | N>>3 | M>>4
__device__ __forceinline__ uint32_t make_idesc_bf16(int block_n, int a_mn_major, int b_mn_major, int m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// arrive on an mbarrier of CTA 0 of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cta0(uint32_t bar_local_addr) {
    const uint32_t remote = mapa_cta0(bar_local_addr);
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

struct EpiRegs {
    float4 bias;
    float4 a0[8];
    float4 a1[8];
    float4 a2[8];
};

// The epilogue kind is a template parameter of the kernel: a run-time switch in the innermost loop compiled to an indirect
// branch through a constant-memory jump table per 16 output bytes and dominated the epilogue (r01d trace: ~9000 cycles).
__host__ __device__ constexpr bool epi_has_bias(int e) { return e == MAP_EPI_BIAS || e == MAP_EPI_BIAS_RELU || e == MAP_EPI_CROSS; }
__host__ __device__ constexpr bool epi_has_aux0(int e) { return e >= MAP_EPI_CROSS; }   // CROSS_BWD: optional (checked at run time)
__host__ __device__ constexpr bool epi_has_aux1(int e) {
    return e == MAP_EPI_CROSS || e == MAP_EPI_ADD_MUL || e == MAP_EPI_CROSS_BWD || e == MAP_EPI_ADD3;
}
__host__ __device__ constexpr bool epi_has_aux2(int e) { return e == MAP_EPI_CROSS_BWD || e == MAP_EPI_ADD3; }  // ADD3: optional
__host__ __device__ constexpr bool epi_has_aux_out(int e) { return e == MAP_EPI_CROSS || e == MAP_EPI_ADD_MUL || e == MAP_EPI_CROSS_BWD; }
// accumulator columns per epilogue chunk: three streamed operands per output only fit in registers at half width
__host__ __device__ constexpr int epi_chunk_w(int e) { return (e == MAP_EPI_CROSS_BWD || e == MAP_EPI_ADD3) ? 16 : 32; }

template <int EPI, int W>
__device__ __forceinline__ void epi_prefetch(const GemmParams& p, int lane, int row_base, int ncol0, bool full_tile, EpiRegs& e) {
    constexpr int LPR = W / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
    const int cg = lane % LPR, rr = lane / LPR;
    const int n = ncol0 + 4 * cg;
    const bool n_ok = full_tile || n < p.N;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (epi_has_bias(EPI)) e.bias = n_ok ? ld4_nc(p.bias + n) : z;
    if (epi_has_aux0(EPI)) {
        const bool have = (EPI != MAP_EPI_CROSS_BWD) || p.aux0 != nullptr;
        const float* src = p.aux0 + (int64_t)(row_base + rr) * p.ld_aux0 + n;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
            e.a0[i] = (have && n_ok && (full_tile || row_base + i * RPI + rr < p.M)) ? ld4_nc(src) : z;
            src += (int64_t)RPI * p.ld_aux0;
        }
    }
    if (epi_has_aux1(EPI)) {
        const float* src = p.aux1 + (int64_t)(row_base + rr) * p.ld_aux1 + n;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
            e.a1[i] = (n_ok && (full_tile || row_base + i * RPI + rr < p.M)) ? ld4_nc(src) : z;
            src += (int64_t)RPI * p.ld_aux1;
        }
    }
    if (epi_has_aux2(EPI)) {
        const bool have = (EPI != MAP_EPI_ADD3) || p.aux2 != nullptr;
        const float* src = p.aux2 + (int64_t)(row_base + rr) * p.ld_aux2 + n;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
            e.a2[i] = (have && n_ok && (full_tile || row_base + i * RPI + rr < p.M)) ? ld4_nc(src) : z;
            src += (int64_t)RPI * p.ld_aux2;
        }
    }
}

__device__ __forceinline__ void red_add4(float* dst, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// returns the value written to C (for the fused column sums)
template <int EPI, bool SPLIT>
__device__ __forceinline__ float4 epilogue_store4(float* dst, float* aux_dst, float* acc_dst, bool acc_accumulate, float4 acc,
                                                  const float4& b, const float4& x0_, const float4& x1_, const float4& x2_) {
    float4 out = acc;
    if (EPI == MAP_EPI_BIAS) {
        out = make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w);
    } else if (EPI == MAP_EPI_BIAS_RELU) {
        out = make_float4(fmaxf(acc.x + b.x, 0.f), fmaxf(acc.y + b.y, 0.f), fmaxf(acc.z + b.z, 0.f), fmaxf(acc.w + b.w, 0.f));
    } else if (EPI == MAP_EPI_CROSS) {  // aux0 = Xi, aux1 = X0
        const float4 u = make_float4(acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w);
        *reinterpret_cast<float4*>(aux_dst) = u;
        out = make_float4(fmaf(x1_.x, u.x, x0_.x), fmaf(x1_.y, u.y, x0_.y), fmaf(x1_.z, u.z, x0_.z), fmaf(x1_.w, u.w, x0_.w));
    } else if (EPI == MAP_EPI_MUL_RELUMASK) {
        out = make_float4(x0_.x > 0.f ? acc.x : 0.f, x0_.y > 0.f ? acc.y : 0.f, x0_.z > 0.f ? acc.z : 0.f, x0_.w > 0.f ? acc.w : 0.f);
    } else if (EPI == MAP_EPI_ADD) {
        out = make_float4(acc.x + x0_.x, acc.y + x0_.y, acc.z + x0_.z, acc.w + x0_.w);
    } else if (EPI == MAP_EPI_ADD_MUL) {
        const float4 sm = make_float4(acc.x + x0_.x, acc.y + x0_.y, acc.z + x0_.z, acc.w + x0_.w);
        *reinterpret_cast<float4*>(aux_dst) = sm;
        out = make_float4(sm.x * x1_.x, sm.y * x1_.y, sm.z * x1_.z, sm.w * x1_.w);
    } else if (EPI == MAP_EPI_CROSS_BWD) {  // aux0 = G above (or 0), aux1 = X0, aux2 = U below
        const float4 g = make_float4(acc.x + x0_.x, acc.y + x0_.y, acc.z + x0_.z, acc.w + x0_.w);
        if (aux_dst != nullptr) *reinterpret_cast<float4*>(aux_dst) = g;
        const float4 t = make_float4(g.x * x2_.x, g.y * x2_.y, g.z * x2_.z, g.w * x2_.w);
        if (acc_accumulate) red_add4(acc_dst, t); else *reinterpret_cast<float4*>(acc_dst) = t;
        out = make_float4(g.x * x1_.x, g.y * x1_.y, g.z * x1_.z, g.w * x1_.w);
    } else if (EPI == MAP_EPI_ADD3) {
        out = make_float4(acc.x + x0_.x + x1_.x + x2_.x, acc.y + x0_.y + x1_.y + x2_.y, acc.z + x0_.z + x1_.z + x2_.z,
                          acc.w + x0_.w + x1_.w + x2_.w);
    }
    if (SPLIT) {  // partial sums of a K split: one 16-byte fp32 reduction into the pre-zeroed output (EPI_NONE only)
        red_add4(dst, out);
    } else {
        *reinterpret_cast<float4*>(dst) = out;
    }
    return out;
}

// C as bf16 planes: hi = rn(x), lo = rn(x - hi), lo2 = rn(x - hi - lo) (the subtractions are exact in fp32)
// two floats -> packed bf16x2 (round to nearest even; one cvt.rn.bf16x2.f32), low half = a
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void store_planes4(const GemmParams& p, int64_t row, int n, const float4& o) {
    float r0 = o.x, r1 = o.y, r2 = o.z, r3 = o.w;
    __nv_bfloat16* dst = p.c_planes + row * p.ld_cp + n;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        if (pl < p.pc) {
            uint2 pk;
            pk.x = pack_bf16x2(r0, r1);
            pk.y = pack_bf16x2(r2, r3);
            *reinterpret_cast<uint2*>(dst) = pk;
            if (pl + 1 < p.pc) {   // residual for the next plane: x - float(bf16(x)) is exact in fp32
                r0 -= __uint_as_float(pk.x << 16); r1 -= __uint_as_float(pk.x & 0xFFFF0000u);
                r2 -= __uint_as_float(pk.y << 16); r3 -= __uint_as_float(pk.y & 0xFFFF0000u);
                dst += p.cp_stride;
            }
        }
    }
}

constexpr int kStageLd = 36;  // floats per staging row (144 B: 16-byte aligned, conflict-free 128-bit phases)
#ifdef MAP_GEMM_STAGE_LD16    // split-bf16 kernel: 8 epilogue warps with 16-column chunks only -> 20-float staging rows (2.5 KB per warp)
constexpr int kStageLd16 = 20;
#else
constexpr int kStageLd16 = kStageLd;
#endif

// one chunk: TMEM -> registers -> per-warp smem transpose -> fused epilogue on the prefetched operands -> global
template <int EPI, bool SPLIT, int W>
__device__ __forceinline__ void epi_chunk(const GemmParams& p, float* stg, int lane, uint32_t taddr, int row_base, int ncol0,
                                          bool full_tile, const EpiRegs& e) {
    constexpr int LPR = W / 4, RPI = 32 / LPR, ITERS = 32 / RPI;
    constexpr int LD = (W == 16) ? kStageLd16 : kStageLd;
    float v[W];
    if (W == 32) tmem_ld_x32(taddr, v); else tmem_ld_x16(taddr, v);
#ifdef MAP_GEMM_TWO_ACCUMULATORS   // compiled into the split-bf16 kernel only (the TF32 kernel runs 2 CTAs per SM at 168 registers)
    if (p.corr_cols != 0) {
        float w[W];
        if (W == 32) tmem_ld_x32(taddr + (uint32_t)p.corr_cols, w); else tmem_ld_x16(taddr + (uint32_t)p.corr_cols, w);
#pragma unroll
        for (int j = 0; j < W; ++j) v[j] += w[j];
    }
#endif
#pragma unroll
    for (int j = 0; j < W / 4; ++j)
        *reinterpret_cast<float4*>(stg + lane * LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
    const int cg = lane % LPR, rr = lane / LPR;
    const int n = ncol0 + 4 * cg;
    float* dst = p.C + (int64_t)(row_base + rr) * p.ldc + n;
    float* aux_dst = (epi_has_aux_out(EPI) && p.aux_out != nullptr) ? p.aux_out + (int64_t)(row_base + rr) * p.ld_aux_out + n : nullptr;
    float* acc_dst = (EPI == MAP_EPI_CROSS_BWD) ? p.acc_out + (int64_t)(row_base + rr) * p.ld_acc_out + n : nullptr;
    const bool acc_accumulate = p.acc_accumulate != 0;
    const float* src = stg + rr * LD + 4 * cg;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (full_tile) {  // interior tile (the common case): straight-line code, no per-row guards
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(src + i * RPI * LD);
            const float4 o = epilogue_store4<EPI, SPLIT>(dst, aux_dst, acc_dst, acc_accumulate, a, e.bias, e.a0[i], e.a1[i], e.a2[i]);
            cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
            if (!SPLIT && p.c_planes != nullptr) store_planes4(p, row_base + i * RPI + rr, n, o);
            dst += (int64_t)RPI * p.ldc;
            if (epi_has_aux_out(EPI) && aux_dst != nullptr) aux_dst += (int64_t)RPI * p.ld_aux_out;
            if (EPI == MAP_EPI_CROSS_BWD) acc_dst += (int64_t)RPI * p.ld_acc_out;
        }
    } else {
        const bool n_ok = n < p.N;
#pragma unroll
        for (int i = 0; i < ITERS; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(src + i * RPI * LD);
            if (n_ok && row_base + i * RPI + rr < p.M) {
                const float4 o = epilogue_store4<EPI, SPLIT>(dst, aux_dst, acc_dst, acc_accumulate, a, e.bias, e.a0[i], e.a1[i], e.a2[i]);
                cs.x += o.x; cs.y += o.y; cs.z += o.z; cs.w += o.w;
                if (!SPLIT && p.c_planes != nullptr) store_planes4(p, row_base + i * RPI + rr, n, o);
            }
            dst += (int64_t)RPI * p.ldc;
            if (epi_has_aux_out(EPI) && aux_dst != nullptr) aux_dst += (int64_t)RPI * p.ld_aux_out;
            if (EPI == MAP_EPI_CROSS_BWD) acc_dst += (int64_t)RPI * p.ld_acc_out;
        }
    }
    if (!SPLIT && p.colsum_out != nullptr) {  // column sums of C over this warp's 32 rows -> one 16-byte reduction per column group
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o);
            cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o);
            cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        }
        if (lane < LPR && (full_tile || n < p.N)) red_add4(p.colsum_out + n, cs);
    }
    __syncwarp();
}

// chunk width of the grouped kernel's epilogue: it keeps TWO prefetched operand sets in registers (see epi_tile), so every
// epilogue that streams two or more operands uses 16-column chunks (36-52 registers per set)
__host__ __device__ constexpr int epi_chunk_w_group(int e) { return epi_has_aux1(e) ? 16 : 32; }

template <int EPI, int CW>
__device__ __forceinline__ void epi_prefetch_any(const GemmParams& p, int lane, int row_base, int n0, int c, bool full_tile, EpiRegs& e) {
    if (CW == 16 || c + CW <= p.block_n) epi_prefetch<EPI, CW>(p, lane, row_base, n0 + c, full_tile, e);
    else epi_prefetch<EPI, 16>(p, lane, row_base, n0 + c, full_tile, e);   // trailing half chunk (block_n % 32 == 16)
}
template <int EPI, bool SPLIT, int CW>
__device__ __forceinline__ void epi_chunk_any(const GemmParams& p, float* stg, int lane, uint32_t lane_addr, int row_base, int n0, int c,
                                              bool full_tile, const EpiRegs& e) {
    if (CW == 16 || c + CW <= p.block_n) epi_chunk<EPI, SPLIT, CW>(p, stg, lane, lane_addr + (uint32_t)c, row_base, n0 + c, full_tile, e);
    else epi_chunk<EPI, SPLIT, 16>(p, stg, lane, lane_addr + (uint32_t)c, row_base, n0 + c, full_tile, e);
}

// Epilogue of one tile.  The operands of chunk i+1 (bias / aux0 / aux1 / aux2 rows) are requested BEFORE chunk i is pulled out
// of TMEM and written, into the other of two register sets: their L2 round trip (the whole cost of a chunk in the
// single-buffered version: ~2000 clk per 16-32 columns) overlaps the TMEM read, the transpose and the stores of chunk i.
// `part` / `nsplit`: the column chunks of the tile are dealt round-robin to the `nsplit` warps that share a TMEM lane quarter (the
// split-bf16 kernel runs two epilogue warps per quarter: twice the loads in flight); FORCE_CW overrides the chunk width.
template <int EPI, bool SPLIT, int FORCE_CW = 0>
__device__ __forceinline__ void epi_tile(const GemmParams& p, float* stg, int lane, uint32_t lane_addr, int row_base, int m0, int n0,
                                         uint32_t full_bar, unsigned long long* trace, uint32_t parity = 0, int part = 0, int nsplit = 1) {
    constexpr int CW = FORCE_CW != 0 ? FORCE_CW : epi_chunk_w_group(EPI);
    const bool full_tile = (m0 + kBlockM <= p.M) && (n0 + p.block_n <= p.N);
    const int bn = p.block_n;
    const int step = nsplit * CW;
    int c = part * CW;
    EpiRegs ea, eb;
    if (c < bn) epi_prefetch_any<EPI, CW>(p, lane, row_base, n0, c, full_tile, ea);
    mbar_wait(full_bar, parity);
    tcgen05_fence_after();
    if (trace != nullptr && threadIdx.x == 64) trace[5] = (unsigned long long)clock64();
    while (c < bn) {
        const int c1 = c + step;
        if (c1 < bn) epi_prefetch_any<EPI, CW>(p, lane, row_base, n0, c1, full_tile, eb);
        epi_chunk_any<EPI, SPLIT, CW>(p, stg, lane, lane_addr, row_base, n0, c, full_tile, ea);
        if (c1 >= bn) break;
        const int c2 = c1 + step;
        if (c2 < bn) epi_prefetch_any<EPI, CW>(p, lane, row_base, n0, c2, full_tile, ea);
        epi_chunk_any<EPI, SPLIT, CW>(p, stg, lane, lane_addr, row_base, n0, c1, full_tile, eb);
        c = c2;
    }
    if (trace != nullptr && threadIdx.x == 64) trace[6] = (unsigned long long)clock64();
}

template <int EPI, int FORCE_CW = 0>
__device__ __forceinline__ void epi_slot(const GemmParams& p, float* stg, int lane, uint32_t lane_addr, int row_base, int m0, int n0,
                                         uint32_t full_bar, unsigned long long* trace, uint32_t parity = 0, int part = 0, int nsplit = 1) {
    if constexpr (EPI == MAP_EPI_NONE) {
        if (p.split_k > 1) epi_tile<MAP_EPI_NONE, true, FORCE_CW>(p, stg, lane, lane_addr, row_base, m0, n0, full_bar, trace, parity, part, nsplit);
        else epi_tile<MAP_EPI_NONE, false, FORCE_CW>(p, stg, lane, lane_addr, row_base, m0, n0, full_bar, trace, parity, part, nsplit);
    } else if constexpr (EPI > MAP_EPI_NONE) {
        epi_tile<EPI, false, FORCE_CW>(p, stg, lane, lane_addr, row_base, m0, n0, full_bar, trace, parity, part, nsplit);
    }
}

}  // namespace mapb
