// K16: the MFP feature encoder evaluated BY FIELD (reference code/models.py:73-78).
//
// The reference computes enc = feat_encoder(final) for all F fields ([B, F*P], a [B,Kd] x [Kd, F*P] GEMM) and then keeps the L
// masked P-wide slices per sample (torch.gather).  Only those slices carry a gradient, so forward, dgrad and wgrad all reduce to
//     sel[n, :]   = final[b(n), :] . W_f(n)^T + bias_f(n)                 n = b*L + l, f(n) = masked_index[b, l]
//     dX[b, :]    = sum_l d_sel[b*L+l, :] . W_f(b,l)
//     dW_f        = sum_{n: f(n) = f} d_sel[n, :]^T (x) final[b(n), :]    db_f = sum_{n: f(n) = f} d_sel[n, :]
// with W_f = rows [f*P, (f+1)*P) of feat_encoder.weight: F/L times fewer flops (13x at mask_ratio 0.1) and no [B, F*P] tensors.
// Positions are bucketed by field on the device (stable counting sort, K16a) so that a CTA works on ONE W_f; the rows of `final`
// are gathered by index.  The contraction per field is [~B*L/F, Kd] x [Kd, P] with P = 32: gather-bound, exact fp32 FMAs on the
// CUDA cores (cp.async-staged, register-tiled), deterministic summation order.  K16d folds the L partial rows per sample and
// applies the first stage of the towers' backward (ReLU mask / CrossNetV2 stage / bias column sums / bf16 operand planes) in one
// pass — what the epilogues of the dense head dgrad GEMMs did.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mapb {

constexpr int kFeTM = 32;        // positions per tile
constexpr int kFeKC = 32;        // k (or position) elements per staged chunk
constexpr int kFeLd = 36;        // shared-memory row stride in floats: 16-byte aligned rows, conflict-free float4 reads of 8 rows
constexpr int kFeKG = 4;         // forward: K groups per CTA (in-CTA split-K, 64 threads each)
constexpr int kFeMaxFields = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void group_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ------------------------------------------------------------------------------------------------ K16a: bucket positions by field
// CTA f: start = #{n : mi[n] < f}; then the positions with mi[n] == f in ascending n (stable) -> perm[start ...].  One launch, no
// inter-CTA dependency.
__global__ void __launch_bounds__(1024) field_bucket_kernel(const int64_t* __restrict__ mi, int N, int F, int32_t* __restrict__ perm,
                                                            int32_t* __restrict__ fstart) {
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int c = 0;
    for (int n = threadIdx.x; n < N; n += 1024) c += (__ldg(mi + n) < (int64_t)f) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) warp_tot[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 32; ++w) t += warp_tot[w];
        s_base = t;
        fstart[f] = t;
    }
    __syncthreads();
    int running = s_base;
    for (int n0 = 0; n0 < N; n0 += 1024) {
        const int n = n0 + threadIdx.x;
        const bool hit = n < N && __ldg(mi + n) == (int64_t)f;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        __syncthreads();   // warp_tot is reused
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            woff += (w < warp) ? t : 0;
            tot += t;
        }
        if (hit) perm[running + woff + __popc(bal & ((1u << lane) - 1u))] = n;
        running += tot;
    }
    if (f == F - 1 && threadIdx.x == 0) fstart[F] = running;
}

// tile table of a launch: tstart[f] = number of position tiles of the fields before f (kFeTM positions per tile)
__device__ __forceinline__ void build_tile_starts(const int32_t* __restrict__ fstart, int F, int* tstart) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) tstart[f] = (fstart[f + 1] - fstart[f] + kFeTM - 1) / kFeTM;   // tiles of field f
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int f = 0; f < F; ++f) {
            const int n = tstart[f];
            tstart[f] = t;
            t += n;
        }
        tstart[F] = t;
    }
    __syncthreads();
}
__device__ __forceinline__ int field_of_tile(const int* tstart, int F, int tile) {
    int lo = 0, hi = F - 1;   // last f with tstart[f] <= tile
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tstart[mid] <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ K16b: forward
// CTA = one tile of 32 positions of one field x P outputs; 4 groups of 64 threads split the K loop (chunks of 32, cp.async double
// buffer per group); thread = 4 positions x P/8 outputs; the 4 partial tiles are summed in a fixed order.
template <int P>
__global__ void __launch_bounds__(64 * kFeKG, 3) field_enc_fwd_kernel(const float* __restrict__ X, int64_t ldx, int K, const float* __restrict__ W,
                                                                   int64_t ldw, const float* __restrict__ bias,
                                                                   const int32_t* __restrict__ perm, const int32_t* __restrict__ fstart, int L,
                                                                   int F, float* __restrict__ sel) {
    constexpr int PJ = P / 8;
    constexpr int ROWS = kFeTM + P;                 // staged rows per chunk: 32 gathered rows of X, then P rows of W_f
    constexpr int STAGE = ROWS * kFeLd;             // floats per stage
    extern __shared__ __align__(16) float smem[];   // [kFeKG][2][ROWS][kFeLd]
    __shared__ int tstart[kFeMaxFields + 1];
    __shared__ int rowb[kFeTM];
    build_tile_starts(fstart, F, tstart);
    const int total = tstart[F];
    const int g = threadIdx.x >> 6, lt = threadIdx.x & 63;
    const int tm = lt >> 3, tj = lt & 7;
    float* const gs = smem + (size_t)g * 2 * STAGE;
    const int nchunks = (K + kFeKC - 1) / kFeKC;
    const int ng = nchunks > g ? (nchunks - g + kFeKG - 1) / kFeKG : 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int f = field_of_tile(tstart, F, tile);
        const int s0 = fstart[f] + (tile - tstart[f]) * kFeTM;
        const int rows = min(kFeTM, fstart[f + 1] - s0);
        __syncthreads();   // previous tile's epilogue has read rowb / the partial tiles
        if (threadIdx.x < kFeTM) rowb[threadIdx.x] = perm[s0 + min((int)threadIdx.x, rows - 1)] / L;
        __syncthreads();
        const float* const Wf = W + (int64_t)f * P * ldw;
        auto issue = [&](int chunk, int st) {
            float* const dst = gs + (size_t)st * STAGE;
            const int k0 = chunk * kFeKC;
#pragma unroll
            for (int i = 0; i < ROWS * 8 / 64; ++i) {
                const int q = lt + 64 * i, r = q >> 3, k = k0 + (q & 7) * 4;
                const bool ok = k < K;
                const float* src = (r < kFeTM) ? X + (int64_t)rowb[r] * ldx : Wf + (int64_t)(r - kFeTM) * ldw;
                cp_async16(dst + r * kFeLd + (q & 7) * 4, ok ? src + k : src, ok);
            }
        };
        float acc[4][PJ];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < PJ; ++j) acc[i][j] = 0.f;
        if (ng > 0) issue(g, 0);
        cp_async_commit();
        for (int it = 0; it < ng; ++it) {
            if (it + 1 < ng) issue(g + kFeKG * (it + 1), (it + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
            group_bar(1 + g, 64);
            const float* const As = gs + (size_t)(it & 1) * STAGE;
            const float* const Ws = As + kFeTM * kFeLd;
#pragma unroll
            for (int k4 = 0; k4 < kFeKC / 4; ++k4) {
                float4 a[4], w[PJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(As + (tm + 8 * i) * kFeLd + k4 * 4);
#pragma unroll
                for (int j = 0; j < PJ; ++j) w[j] = *reinterpret_cast<const float4*>(Ws + (tj + 8 * j) * kFeLd + k4 * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < PJ; ++j) {
                        acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
                    }
            }
            group_bar(1 + g, 64);
        }
        cp_async_wait<0>();
        // partial tile of this group -> its own stage memory [32][P] (the group is past its last barrier: nobody reads the stages)
        float* const red = gs;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < PJ; ++j) red[(tm + 8 * i) * P + tj + 8 * j] = acc[i][j];
        __syncthreads();
        for (int e = threadIdx.x * 4; e < kFeTM * P; e += 64 * kFeKG * 4) {
            const int m = e / P, j = e - m * P;
            float4 s = *reinterpret_cast<const float4*>(bias + f * P + j);
#pragma unroll
            for (int gg = 0; gg < kFeKG; ++gg) {
                const float4 v = *reinterpret_cast<const float4*>(smem + (size_t)gg * 2 * STAGE + e);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            if (m < rows) *reinterpret_cast<float4*>(sel + (int64_t)perm[s0 + m] * P + j) = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K16c: dgrad per position
// dxpos[n, c] = sum_j d_sel[n, j] * W[f(n)*P + j, c].  CTA = 32 positions of one field x 128 columns; thread = 4 positions x 8 columns.
constexpr int kFeDC = 128;       // columns per dgrad tile
constexpr int kFeDLd = kFeDC + 4;
template <int P>
__global__ void __launch_bounds__(128) field_enc_dgrad_kernel(const float* __restrict__ d_sel, const float* __restrict__ W, int64_t ldw, int K,
                                                              const int32_t* __restrict__ perm, const int32_t* __restrict__ fstart, int F,
                                                              float* __restrict__ dxpos, int64_t ld_dx) {
    constexpr int DLD = P + 4;   // d_sel tile row stride
    __shared__ __align__(16) float ds_s[kFeTM * DLD];
    __shared__ __align__(16) float w_s[P * kFeDLd];
    __shared__ int tstart[kFeMaxFields + 1];
    __shared__ int rown[kFeTM];
    build_tile_starts(fstart, F, tstart);
    const int ncol = (K + kFeDC - 1) / kFeDC;
    const int64_t total = (int64_t)tstart[F] * ncol;
    const int tm = threadIdx.x >> 4, tc = threadIdx.x & 15;
    for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
        const int tile = (int)(t / ncol), c0 = (int)(t - (int64_t)tile * ncol) * kFeDC;
        const int f = field_of_tile(tstart, F, tile);
        const int s0 = fstart[f] + (tile - tstart[f]) * kFeTM;
        const int rows = min(kFeTM, fstart[f + 1] - s0);
        __syncthreads();
        if (threadIdx.x < kFeTM) rown[threadIdx.x] = perm[s0 + min((int)threadIdx.x, rows - 1)];
        __syncthreads();
        for (int q = threadIdx.x; q < kFeTM * (P / 4); q += 128) {
            const int r = q / (P / 4), p4 = q - r * (P / 4);
            *reinterpret_cast<float4*>(ds_s + r * DLD + p4 * 4) = *reinterpret_cast<const float4*>(d_sel + (int64_t)rown[r] * P + p4 * 4);
        }
        for (int q = threadIdx.x; q < P * (kFeDC / 4); q += 128) {
            const int j = q / (kFeDC / 4), c4 = q - j * (kFeDC / 4), c = c0 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < K) v = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(f * P + j) * ldw + c));
            *reinterpret_cast<float4*>(w_s + j * kFeDLd + c4 * 4) = v;
        }
        __syncthreads();
        float4 acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j4 = 0; j4 < P / 4; ++j4) {
            float4 d[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = *reinterpret_cast<const float4*>(ds_s + (tm + 8 * i) * DLD + j4 * 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 w0 = *reinterpret_cast<const float4*>(w_s + (j4 * 4 + jj) * kFeDLd + tc * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(w_s + (j4 * 4 + jj) * kFeDLd + tc * 4 + 64);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dv = jj == 0 ? d[i].x : jj == 1 ? d[i].y : jj == 2 ? d[i].z : d[i].w;
                    acc[i][0].x = fmaf(dv, w0.x, acc[i][0].x); acc[i][0].y = fmaf(dv, w0.y, acc[i][0].y);
                    acc[i][0].z = fmaf(dv, w0.z, acc[i][0].z); acc[i][0].w = fmaf(dv, w0.w, acc[i][0].w);
                    acc[i][1].x = fmaf(dv, w1.x, acc[i][1].x); acc[i][1].y = fmaf(dv, w1.y, acc[i][1].y);
                    acc[i][1].z = fmaf(dv, w1.z, acc[i][1].z); acc[i][1].w = fmaf(dv, w1.w, acc[i][1].w);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = tm + 8 * i;
            if (m >= rows) continue;
            float* const dst = dxpos + (int64_t)rown[m] * ld_dx + c0 + tc * 4;
            if (c0 + tc * 4 < K) st_stream_f4(reinterpret_cast<float4*>(dst), acc[i][0]);
            if (c0 + tc * 4 + 64 < K) st_stream_f4(reinterpret_cast<float4*>(dst + 64), acc[i][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K16d: fold + first backward stage
__device__ __forceinline__ void store_planes4(uint16_t* planes, int64_t ld, int64_t stride, int n_planes, int64_t row, int col, float4 v) {
    float r0 = v.x, r1 = v.y, r2 = v.z, r3 = v.w;
    uint16_t* dst = planes + row * ld + col;
    for (int pl = 0; pl < n_planes; ++pl) {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(r0, r1), h23 = __floats2bfloat162_rn(r2, r3);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h01);
        pk.y = *reinterpret_cast<const uint32_t*>(&h23);
        *reinterpret_cast<uint2*>(dst) = pk;
        r0 -= __uint_as_float(pk.x << 16); r1 -= __uint_as_float(pk.x & 0xFFFF0000u);
        r2 -= __uint_as_float(pk.y << 16); r3 -= __uint_as_float(pk.y & 0xFFFF0000u);
        dst += stride;
    }
}

constexpr int kFeRB = 8;   // rows per CTA of the fold (two in flight per thread: the kernel is a latency-bound stream)
__global__ void __launch_bounds__(128) head_bwd_fold_kernel(MapHeadBwdArgs a) {
    const int c = (blockIdx.y * 128 + threadIdx.x) * 4;
    if (c >= a.ncols) return;
    const int64_t b0 = (int64_t)blockIdx.x * kFeRB;
    const int64_t b1 = b0 + kFeRB < a.B ? b0 + kFeRB : a.B;
    const bool in_cross = a.cross_w > 0 && c >= a.cross_col0 && c < a.cross_col0 + a.cross_w;
    const bool in_relu = a.relu_w > 0 && c >= a.relu_col0 && c < a.relu_col0 + a.relu_w;
    const bool has_scalar = a.scalar_col >= c && a.scalar_col < c + 4;
    const int cc = c - (in_cross ? a.cross_col0 : in_relu ? a.relu_col0 : 0);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 cs = zero4;
    for (int64_t bb = b0; bb < b1; bb += 2) {
        float4 g[2], p[2], q[2];   // p = X0 (CrossNet) or y (ReLU); q = U (CrossNet)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t b = bb + r < b1 ? bb + r : bb;   // odd tail: the second lane repeats the first row and is not stored
            const float* src = a.dxpos + b * a.L * a.ld_dx + c;
            g[r] = ld_stream_f4(reinterpret_cast<const float4*>(src));
            for (int l = 1; l < a.L; ++l) {
                const float4 v = ld_stream_f4(reinterpret_cast<const float4*>(src + (int64_t)l * a.ld_dx));
                g[r].x += v.x; g[r].y += v.y; g[r].z += v.z; g[r].w += v.w;
            }
            p[r] = q[r] = zero4;
            if (in_cross) {
                p[r] = *reinterpret_cast<const float4*>(a.x0 + b * a.ld_x0 + cc);
                q[r] = *reinterpret_cast<const float4*>(a.u + b * a.ld_u + cc);
            } else if (in_relu) {
                p[r] = *reinterpret_cast<const float4*>(a.y + b * a.ld_y + cc);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t b = bb + r;
            if (b >= b1) break;
            const float4 gg = g[r];
            if (in_cross) {   // G = g ; dU = G * X0 ; dX0 = G * U   (MAP_EPI_CROSS_BWD with no layer above)
                *reinterpret_cast<float4*>(a.g_out + b * a.ld_g + cc) = gg;
                *reinterpret_cast<float4*>(a.dx0_out + b * a.ld_dx0 + cc) = make_float4(gg.x * q[r].x, gg.y * q[r].y, gg.z * q[r].z, gg.w * q[r].w);
                const float4 du = make_float4(gg.x * p[r].x, gg.y * p[r].y, gg.z * p[r].z, gg.w * p[r].w);
                *reinterpret_cast<float4*>(a.du_out + b * a.ld_du + cc) = du;
                if (a.du_planes != nullptr) store_planes4(a.du_planes, a.du_pl_ld, a.du_pl_stride, a.du_nplanes, b, cc, du);
                cs.x += du.x; cs.y += du.y; cs.z += du.z; cs.w += du.w;
            } else if (in_relu) {   // dZ = g * (y > 0)   (MAP_EPI_MUL_RELUMASK)
                const float4 dz = make_float4(p[r].x > 0.f ? gg.x : 0.f, p[r].y > 0.f ? gg.y : 0.f, p[r].z > 0.f ? gg.z : 0.f, p[r].w > 0.f ? gg.w : 0.f);
                *reinterpret_cast<float4*>(a.dz_out + b * a.ld_dz + cc) = dz;
                if (a.dz_planes != nullptr) store_planes4(a.dz_planes, a.dz_pl_ld, a.dz_pl_stride, a.dz_nplanes, b, cc, dz);
                cs.x += dz.x; cs.y += dz.y; cs.z += dz.z; cs.w += dz.w;
            }
            if (has_scalar) {
                const int o = a.scalar_col - c;
                a.scalar_out[b * a.ld_scalar] = o == 0 ? gg.x : o == 1 ? gg.y : o == 2 ? gg.z : gg.w;
            }
        }
    }
    float* bg = in_cross ? a.cross_bias_grad : in_relu ? a.relu_bias_grad : nullptr;
    if (bg != nullptr) {
        bg += cc;
        atomicAdd(bg + 0, cs.x); atomicAdd(bg + 1, cs.y); atomicAdd(bg + 2, cs.z); atomicAdd(bg + 3, cs.w);
    }
}

// ------------------------------------------------------------------------------------------------ K16e: wgrad per field
// dW[f*P + j, c] = sum_{s in field f} d_sel[perm[s], j] * X[b(perm[s]), c];  db[f*P + j] = sum_s d_sel[perm[s], j].
// CTA = (64 columns, field); loops over the field's positions in chunks of 16 (cp.async double buffer); thread = P/8 rows x 4 columns.
constexpr int kFeWC = 64, kFeWS = 16, kFeWLd = kFeWC + 4;
template <int P>
__global__ void __launch_bounds__(128) field_enc_wgrad_kernel(const float* __restrict__ d_sel, const float* __restrict__ X, int64_t ldx, int K,
                                                              const int32_t* __restrict__ perm, const int32_t* __restrict__ fstart, int L,
                                                              float* __restrict__ dW, int64_t ldw, float* __restrict__ dbias) {
    constexpr int PJ = P / 8;
    constexpr int DLD = P + 4;
    __shared__ __align__(16) float x_s[2][kFeWS * kFeWLd];
    __shared__ __align__(16) float d_s[2][kFeWS * DLD];
    const int f = blockIdx.y, c0 = blockIdx.x * kFeWC;
    const int sb = fstart[f], se = fstart[f + 1];
    const int tj = threadIdx.x >> 4, tc = threadIdx.x & 15;
    const int nch = (se - sb + kFeWS - 1) / kFeWS;
    auto issue = [&](int ch, int st) {
        const int s0 = sb + ch * kFeWS;
#pragma unroll
        for (int i = 0; i < kFeWS * (kFeWC / 4) / 128; ++i) {
            const int q = threadIdx.x + 128 * i, r = q >> 4, c = c0 + (q & 15) * 4;
            const bool ok = s0 + r < se && c < K;
            const float* src = X;
            if (ok) src = X + (int64_t)(perm[s0 + r] / L) * ldx + c;
            cp_async16(&x_s[st][r * kFeWLd + (q & 15) * 4], src, ok);
        }
        for (int q = threadIdx.x; q < kFeWS * (P / 4); q += 128) {
            const int r = q / (P / 4), p4 = q - r * (P / 4);
            const bool ok = s0 + r < se;
            const float* src = d_sel;
            if (ok) src = d_sel + (int64_t)perm[s0 + r] * P + p4 * 4;
            cp_async16(&d_s[st][r * DLD + p4 * 4], src, ok);
        }
    };
    float4 acc[PJ];
    float bacc[PJ];
#pragma unroll
    for (int j = 0; j < PJ; ++j) { acc[j] = make_float4(0.f, 0.f, 0.f, 0.f); bacc[j] = 0.f; }
    if (nch > 0) issue(0, 0);
    cp_async_commit();
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) issue(ch + 1, (ch + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* const xs = x_s[ch & 1];
        const float* const ds = d_s[ch & 1];
#pragma unroll
        for (int s = 0; s < kFeWS; ++s) {
            const float4 x = *reinterpret_cast<const float4*>(xs + s * kFeWLd + tc * 4);
            float dv[PJ];
            if constexpr (PJ == 1) {
                dv[0] = ds[s * DLD + tj];
            } else if constexpr (PJ == 2) {
                const float2 t = *reinterpret_cast<const float2*>(ds + s * DLD + tj * PJ);
                dv[0] = t.x; dv[1] = t.y;
            } else {
#pragma unroll
                for (int q = 0; q < PJ / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(ds + s * DLD + tj * PJ + q * 4);
                    dv[q * 4 + 0] = t.x; dv[q * 4 + 1] = t.y; dv[q * 4 + 2] = t.z; dv[q * 4 + 3] = t.w;
                }
            }
#pragma unroll
            for (int j = 0; j < PJ; ++j) {
                const float d = dv[j];
                acc[j].x = fmaf(d, x.x, acc[j].x); acc[j].y = fmaf(d, x.y, acc[j].y);
                acc[j].z = fmaf(d, x.z, acc[j].z); acc[j].w = fmaf(d, x.w, acc[j].w);
            }
            if (blockIdx.x == 0) {
#pragma unroll
                for (int j = 0; j < PJ; ++j) bacc[j] += dv[j];
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int j = 0; j < PJ; ++j) {
        const int row = f * P + tj * PJ + j;
        if (c0 + tc * 4 < K) *reinterpret_cast<float4*>(dW + (int64_t)row * ldw + c0 + tc * 4) = acc[j];
        if (dbias != nullptr && blockIdx.x == 0 && tc == 0) dbias[row] = bacc[j];
    }
}

static bool fe_p_ok(int P) { return P == 8 || P == 16 || P == 32 || P == 64; }

}  // namespace mapb

using namespace mapb;

extern "C" int map_field_enc_supported(int F, int P) { return (F >= 1 && F <= kFeMaxFields && fe_p_ok(P)) ? 1 : 0; }

extern "C" int map_field_bucket(const int64_t* masked_index, int64_t N, int F, int32_t* perm, int32_t* fstart, map_stream_t stream) {
    MAP_REQUIRE(masked_index && perm && fstart && N >= 1 && N < (1ll << 31) && F >= 1 && F <= kFeMaxFields, "map_field_bucket: bad argument");
    field_bucket_kernel<<<F, 1024, 0, as_stream(stream)>>>(masked_index, (int)N, F, perm, fstart);
    return check_launch("map_field_bucket");
}

#define FE_DISPATCH_P(P, ...)                                   \
    switch (P) {                                                \
        case 8: { constexpr int P_ = 8; __VA_ARGS__; break; }   \
        case 16: { constexpr int P_ = 16; __VA_ARGS__; break; } \
        case 32: { constexpr int P_ = 32; __VA_ARGS__; break; } \
        default: { constexpr int P_ = 64; __VA_ARGS__; break; } \
    }

extern "C" int map_field_enc_fwd(const float* X, int64_t ldx, int K, const float* W, int64_t ldw, const float* bias, const int32_t* perm,
                                 const int32_t* fstart, int64_t N, int L, int F, int P, float* sel, map_stream_t stream) {
    MAP_REQUIRE(X && W && bias && perm && fstart && sel && N >= 1 && L >= 1 && K >= 4, "map_field_enc_fwd: bad argument");
    MAP_REQUIRE(map_field_enc_supported(F, P), "map_field_enc_fwd: F <= %d and P in {8, 16, 32, 64} (got F=%d P=%d)", kFeMaxFields, F, P);
    MAP_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)bias & 15) == 0 &&
                    ((uintptr_t)sel & 15) == 0,
                "map_field_enc_fwd: K, ldx, ldw must be multiples of 4 floats and the pointers 16-byte aligned");
    const int64_t max_tiles = ceil_div(N, kFeTM) + F;
    FE_DISPATCH_P(P, {
        const size_t smem = (size_t)kFeKG * 2 * (kFeTM + P_) * kFeLd * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(field_enc_fwd_kernel<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            attr_set = true;
        }
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, field_enc_fwd_kernel<P_>, 64 * kFeKG, smem);
        if (per_sm < 1) per_sm = 1;
        const int64_t grid = max_tiles < (int64_t)kNumSMs * per_sm ? max_tiles : (int64_t)kNumSMs * per_sm;
        field_enc_fwd_kernel<P_><<<(unsigned)grid, 64 * kFeKG, smem, as_stream(stream)>>>(X, ldx, K, W, ldw, bias, perm, fstart, L, F, sel);
    });
    return check_launch("map_field_enc_fwd");
}

extern "C" int map_field_enc_dgrad(const float* d_sel, const float* W, int64_t ldw, int K, const int32_t* perm, const int32_t* fstart,
                                   int64_t N, int F, int P, float* dxpos, int64_t ld_dx, map_stream_t stream) {
    MAP_REQUIRE(d_sel && W && perm && fstart && dxpos && N >= 1 && K >= 4, "map_field_enc_dgrad: bad argument");
    MAP_REQUIRE(map_field_enc_supported(F, P), "map_field_enc_dgrad: F <= %d and P in {8, 16, 32, 64} (got F=%d P=%d)", kFeMaxFields, F, P);
    MAP_REQUIRE(K % 4 == 0 && ldw % 4 == 0 && ld_dx % 4 == 0 && ld_dx >= K && ((uintptr_t)d_sel & 15) == 0 && ((uintptr_t)W & 15) == 0 &&
                    ((uintptr_t)dxpos & 15) == 0,
                "map_field_enc_dgrad: K, ldw, ld_dx must be multiples of 4 floats and the pointers 16-byte aligned");
    const int64_t grid = (ceil_div(N, kFeTM) + F) * ceil_div(K, kFeDC);   // upper bound of the tile count: one CTA per tile
    FE_DISPATCH_P(P, { field_enc_dgrad_kernel<P_><<<(unsigned)grid, 128, 0, as_stream(stream)>>>(d_sel, W, ldw, K, perm, fstart, F, dxpos, ld_dx); });
    return check_launch("map_field_enc_dgrad");
}

extern "C" int map_field_enc_wgrad(const float* d_sel, const float* X, int64_t ldx, int K, const int32_t* perm, const int32_t* fstart, int64_t N,
                                   int L, int F, int P, float* dW, int64_t ldw, float* dbias, map_stream_t stream) {
    MAP_REQUIRE(d_sel && X && perm && fstart && dW && N >= 1 && L >= 1 && K >= 4, "map_field_enc_wgrad: bad argument");
    MAP_REQUIRE(map_field_enc_supported(F, P), "map_field_enc_wgrad: F <= %d and P in {8, 16, 32, 64} (got F=%d P=%d)", kFeMaxFields, F, P);
    MAP_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && ldw >= K && ((uintptr_t)d_sel & 15) == 0 && ((uintptr_t)X & 15) == 0 &&
                    ((uintptr_t)dW & 15) == 0,
                "map_field_enc_wgrad: K, ldx, ldw must be multiples of 4 floats and the pointers 16-byte aligned");
    const dim3 grid((unsigned)ceil_div(K, kFeWC), (unsigned)F);
    FE_DISPATCH_P(P, { field_enc_wgrad_kernel<P_><<<grid, 128, 0, as_stream(stream)>>>(d_sel, X, ldx, K, perm, fstart, L, dW, ldw, dbias); });
    return check_launch("map_field_enc_wgrad");
}

extern "C" int map_head_bwd_fold(const MapHeadBwdArgs* a, map_stream_t stream) {
    MAP_REQUIRE(a && a->dxpos && a->B >= 1 && a->L >= 1 && a->ncols >= 4 && a->ncols % 4 == 0 && a->ld_dx % 4 == 0 && a->ld_dx >= a->ncols &&
                    ((uintptr_t)a->dxpos & 15) == 0,
                "map_head_bwd_fold: bad argument");
    if (a->cross_w > 0) {
        MAP_REQUIRE(a->cross_col0 % 4 == 0 && a->cross_w % 4 == 0 && a->cross_col0 >= 0 && a->cross_col0 + a->cross_w <= a->ncols && a->x0 && a->u &&
                        a->g_out && a->du_out && a->dx0_out && a->ld_x0 % 4 == 0 && a->ld_u % 4 == 0 && a->ld_g % 4 == 0 && a->ld_du % 4 == 0 &&
                        a->ld_dx0 % 4 == 0 &&
                        (((uintptr_t)a->x0 | (uintptr_t)a->u | (uintptr_t)a->g_out | (uintptr_t)a->du_out | (uintptr_t)a->dx0_out) & 15) == 0,
                    "map_head_bwd_fold: CrossNet region: pointers / strides must be 16-byte aligned, bounds multiples of 4");
        MAP_REQUIRE(a->du_planes == nullptr || (a->du_nplanes >= 1 && a->du_nplanes <= 3 && a->du_pl_ld % 4 == 0 && a->du_pl_stride % 4 == 0 &&
                                                ((uintptr_t)a->du_planes & 7) == 0),
                    "map_head_bwd_fold: dU planes alignment");
    }
    if (a->relu_w > 0) {
        MAP_REQUIRE(a->relu_col0 % 4 == 0 && a->relu_w % 4 == 0 && a->relu_col0 >= 0 && a->relu_col0 + a->relu_w <= a->ncols && a->y && a->dz_out &&
                        a->ld_y % 4 == 0 && a->ld_dz % 4 == 0 && (((uintptr_t)a->y | (uintptr_t)a->dz_out) & 15) == 0,
                    "map_head_bwd_fold: ReLU region: pointers / strides must be 16-byte aligned, bounds multiples of 4");
        MAP_REQUIRE(a->dz_planes == nullptr || (a->dz_nplanes >= 1 && a->dz_nplanes <= 3 && a->dz_pl_ld % 4 == 0 && a->dz_pl_stride % 4 == 0 &&
                                                ((uintptr_t)a->dz_planes & 7) == 0),
                    "map_head_bwd_fold: dZ planes alignment");
    }
    MAP_REQUIRE(a->scalar_col < a->ncols && (a->scalar_col < 0 || a->scalar_out != nullptr), "map_head_bwd_fold: scalar column");
    const dim3 grid((unsigned)ceil_div(a->B, kFeRB), (unsigned)ceil_div(a->ncols / 4, 128));
    head_bwd_fold_kernel<<<grid, 128, 0, as_stream(stream)>>>(*a);
    return check_launch("map_head_bwd_fold");
}
