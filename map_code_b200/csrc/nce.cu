// K6 / K7: fused NCE head.  Replaces NCELoss.forward + IndexLinear._compute_sampled_logit + nce_loss /
// sampled_softmax_loss (code/nce/nce_loss.py:79-144,201-244; code/nce/index_linear.py:68-106).
// The "contraction" is a gather-dot (every position owns its K+1 rows): HBM/L2-bound, no tensor cores.
// One warp per position; LANES = P/4 lanes own one gathered row (16-byte loads), 32/LANES rows in flight per step.
// Forward and the input-side backward are produced in the same pass while the rows are in registers.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mapb {

// Row idx of the (possibly row-sharded) output tables: rank idx % R holds it at local row idx / R (R == 1: the plain table).
__device__ __forceinline__ const float* nce_row(const PeerTable& t, int R, int64_t idx, int width) {
    if (R == 1) return static_cast<const float*>(t.p[0]) + idx * width;
    const int64_t lr = idx / R;
    return static_cast<const float*>(t.p[idx - lr * R]) + lr * width;
}

template <int LANES>
__global__ void __launch_bounds__(256) nce_fwd_kernel(const float* __restrict__ input, int64_t N, int K,
                                                      const int64_t* __restrict__ target, const int64_t* __restrict__ noise,
                                                      const PeerTable emb_t, const PeerTable bias_t, int R,
                                                      const float* __restrict__ logq, float norm_term, int loss_type,
                                                      float grad_scale, float* __restrict__ logits,
                                                      int64_t* __restrict__ ids_out, float* __restrict__ loss_pos,
                                                      float* __restrict__ dz_out, float* __restrict__ d_input,
                                                      int32_t* __restrict__ acc_count) {
    constexpr int P = LANES * 4;
    constexpr int RPW = 32 / LANES;  // rows per warp step
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;   // which 16-byte slice of the row
    const int grp = lane / LANES;   // which row of the step
    const int K1 = K + 1;
    const float ln_k = logf((float)K);
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(input + n * P) + sub);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float loss = 0.f;
        float tgt_logit = 0.f, max_noise = -INFINITY;
        const int steps = (K1 + RPW - 1) / RPW;
        if (loss_type == MAP_NCE_LOSS_NCE && K1 <= 32) {
            // All K+1 ids of the position are fetched with ONE coalesced load (lane j holds id j); the row, bias and
            // log-noise loads of up to 8 steps are then issued back to back before the first use, so a position costs two
            // dependent memory round trips (ids, then rows) instead of 2 x steps (r01d: 36 us -> this form).
            int64_t my_idx = 0;
            if (lane < K1) my_idx = (lane == 0) ? __ldg(target + n) : __ldg(noise + n * K + (lane - 1));
            constexpr int BATCH = LANES < 8 ? LANES : 8;
            for (int it0 = 0; it0 < steps; it0 += BATCH) {
                float4 r[BATCH];
                float bq[BATCH], lq[BATCH];
                int64_t idx[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int j = (it0 + u) * RPW + grp;
                    idx[u] = __shfl_sync(0xffffffffu, my_idx, j & 31);
                    r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    bq[u] = 0.f;
                    lq[u] = 0.f;
                    if (it0 + u < steps && j < K1) {
                        r[u] = __ldg(reinterpret_cast<const float4*>(nce_row(emb_t, R, idx[u], P)) + sub);
                        bq[u] = __ldg(nce_row(bias_t, R, idx[u], 1));
                        lq[u] = __ldg(logq + idx[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int j = (it0 + u) * RPW + grp;
                    const bool valid = it0 + u < steps && j < K1;
                    float s = x.x * r[u].x + x.y * r[u].y + x.z * r[u].z + x.w * r[u].w;
                    s = group_sum<LANES>(s);
                    if (valid) {
                        const float lg = s + bq[u] - norm_term;                  // nce_loss.py:171-172
                        const float z = lg - lq[u] - ln_k;                        // nce_loss.py:222
                        const float y = (j == 0) ? 1.f : 0.f;
                        const float dz = (sigmoidf(z) - y) * grad_scale;
                        acc.x = fmaf(dz, r[u].x, acc.x); acc.y = fmaf(dz, r[u].y, acc.y);
                        acc.z = fmaf(dz, r[u].z, acc.z); acc.w = fmaf(dz, r[u].w, acc.w);
                        if (j == 0) tgt_logit = lg; else max_noise = fmaxf(max_noise, lg);
                        if (sub == 0) {
                            loss += softplusf(z) - y * z;                         // BCEWithLogits, nce_loss.py:227
                            logits[n * K1 + j] = lg;
                            dz_out[n * K1 + j] = dz;
                            if (ids_out != nullptr) ids_out[n * K1 + j] = idx[u];
                        }
                    }
                }
            }
        } else if (loss_type == MAP_NCE_LOSS_NCE) {
            for (int it = 0; it < steps; ++it) {
                const int j = it * RPW + grp;
                const bool valid = j < K1;
                int64_t idx = 0;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    idx = (j == 0) ? __ldg(target + n) : __ldg(noise + n * K + (j - 1));
                    r = __ldg(reinterpret_cast<const float4*>(nce_row(emb_t, R, idx, P)) + sub);
                }
                float s = x.x * r.x + x.y * r.y + x.z * r.z + x.w * r.w;
                s = group_sum<LANES>(s);
                if (valid) {
                    const float lg = s + __ldg(nce_row(bias_t, R, idx, 1)) - norm_term;     // nce_loss.py:171-172
                    const float z = lg - __ldg(logq + idx) - ln_k;           // nce_loss.py:222
                    const float y = (j == 0) ? 1.f : 0.f;
                    const float dz = (sigmoidf(z) - y) * grad_scale;
                    acc.x = fmaf(dz, r.x, acc.x); acc.y = fmaf(dz, r.y, acc.y);
                    acc.z = fmaf(dz, r.z, acc.z); acc.w = fmaf(dz, r.w, acc.w);
                    if (j == 0) tgt_logit = lg; else max_noise = fmaxf(max_noise, lg);
                    if (sub == 0) {
                        loss += softplusf(z) - y * z;                        // BCEWithLogits, nce_loss.py:227
                        logits[n * K1 + j] = lg;
                        dz_out[n * K1 + j] = dz;
                        if (ids_out != nullptr) ids_out[n * K1 + j] = idx;
                    }
                }
            }
        } else {  // sampled softmax (nce_loss.py:232-244): CE over (logit - logq) with label 0
            float mx = -INFINITY;
            for (int it = 0; it < steps; ++it) {
                const int j = it * RPW + grp;
                const bool valid = j < K1;
                int64_t idx = 0;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    idx = (j == 0) ? __ldg(target + n) : __ldg(noise + n * K + (j - 1));
                    r = __ldg(reinterpret_cast<const float4*>(nce_row(emb_t, R, idx, P)) + sub);
                }
                float s = x.x * r.x + x.y * r.y + x.z * r.z + x.w * r.w;
                s = group_sum<LANES>(s);
                if (valid) {
                    const float lg = s + __ldg(nce_row(bias_t, R, idx, 1)) - norm_term;
                    const float c = lg - __ldg(logq + idx);
                    mx = fmaxf(mx, c);
                    if (j == 0) tgt_logit = lg; else max_noise = fmaxf(max_noise, lg);
                    if (sub == 0) {
                        logits[n * K1 + j] = lg;
                        dz_out[n * K1 + j] = c;  // stash the corrected logit; overwritten with the gradient below
                        if (ids_out != nullptr) ids_out[n * K1 + j] = idx;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            __syncwarp();
            float se = 0.f;
            for (int j = lane; j < K1; j += 32) se += expf(dz_out[n * K1 + j] - mx);
            se = warp_sum(se);
            const float lse = mx + logf(se);
            const float c0 = dz_out[n * K1];
            __syncwarp();
            if (lane == 0) loss = lse - c0;
            for (int it = 0; it < steps; ++it) {
                const int j = it * RPW + grp;
                float dz = 0.f;
                if (j < K1) {
                    const int64_t idx = (j == 0) ? __ldg(target + n) : __ldg(noise + n * K + (j - 1));
                    const float4 r = __ldg(reinterpret_cast<const float4*>(nce_row(emb_t, R, idx, P)) + sub);
                    const float c = dz_out[n * K1 + j];
                    dz = (expf(c - lse) - ((j == 0) ? 1.f : 0.f)) * grad_scale;
                    acc.x = fmaf(dz, r.x, acc.x); acc.y = fmaf(dz, r.y, acc.y);
                    acc.z = fmaf(dz, r.z, acc.z); acc.w = fmaf(dz, r.w, acc.w);
                }
                __syncwarp();  // every lane of the group has read the stashed logit before it is overwritten
                if (j < K1 && sub == 0) dz_out[n * K1 + j] = dz;
            }
        }
        // combine across the RPW row groups (lanes with equal `sub`)
#pragma unroll
        for (int o = LANES; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        loss = warp_sum(loss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_noise = fmaxf(max_noise, __shfl_xor_sync(0xffffffffu, max_noise, o));
        const float t0 = __shfl_sync(0xffffffffu, tgt_logit, 0);  // lane 0 (grp 0, sub 0) saw j == 0
        if (d_input != nullptr && lane < LANES) reinterpret_cast<float4*>(d_input + n * P)[lane] = acc;
        if (lane == 0) {
            loss_pos[n] = loss;
            // argmax(dim=2) == 0 with first-occurrence tie-breaking <=> no noise logit is strictly larger (models.py:77)
            if (acc_count != nullptr && !(max_noise > t0)) atomicAdd(acc_count, 1);
        }
    }
}

// sel[n, :] = enc[(b*F + mi[n]) * P + :]   (torch.gather of the masked slices, code/models.py:75)
__global__ void __launch_bounds__(256) gather_slices_kernel(const float* __restrict__ enc, const int64_t* __restrict__ mi,
                                                            int64_t N, int L, int F, int P, float* __restrict__ sel) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * P; e += stride) {
        const int64_t n = e / P;
        const int p = (int)(e - n * P);
        const int64_t b = n / L;
        sel[e] = __ldg(enc + (b * F + mi[n]) * P + p);
    }
}

// autograd of the gather: scatter_add (duplicate masked fields of one sample accumulate)
__global__ void __launch_bounds__(256) scatter_add_slices_kernel(const float* __restrict__ d_sel, const int64_t* __restrict__ mi,
                                                                 int64_t N, int L, int F, int P, float* __restrict__ d_enc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * P; e += stride) {
        const int64_t n = e / P;
        const int p = (int)(e - n * P);
        const int64_t b = n / L;
        atomicAdd(d_enc + (b * F + mi[n]) * P + p, d_sel[e]);
    }
}

// The same scatter written as a DENSE pass over d_enc [B, F*P]: every output element is produced exactly once (sum over the l with
// masked_index[b, l] == f, zero elsewhere), so there is no pre-zeroing, no atomics and the result is deterministic; the optional
// bf16 planes (hi, lo) of d_enc — the operand format of the split-bf16 GEMMs that consume it — come out of the same pass.
__global__ void __launch_bounds__(256) expand_slices_kernel(const float* __restrict__ d_sel, const int64_t* __restrict__ mi, int64_t B, int L,
                                                            int F, int P4, float* __restrict__ d_enc, int64_t ld_enc,
                                                            uint32_t* __restrict__ planes, int64_t ld_p2, int64_t plane_stride2, int n_planes) {
    const int64_t per_row = (int64_t)F * P4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < B * per_row; e += stride) {
        const int64_t b = e / per_row;
        const int r = (int)(e - b * per_row);
        const int f = r / P4, p4 = r - f * P4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l)
            if (__ldg(mi + b * L + l) == f) {
                const float4 a = *reinterpret_cast<const float4*>(d_sel + ((b * L + l) * P4 + p4) * 4);
                v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
            }
        *reinterpret_cast<float4*>(d_enc + b * ld_enc + (int64_t)r * 4) = v;
        if (planes != nullptr) {
            float r0 = v.x, r1 = v.y, r2 = v.z, r3 = v.w;
            uint32_t* dst = planes + b * ld_p2 + (int64_t)r * 2;
            for (int pl = 0; pl < n_planes; ++pl) {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(r0, r1), h23 = __floats2bfloat162_rn(r2, r3);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01);
                pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(dst) = pk;
                r0 -= __uint_as_float(pk.x << 16); r1 -= __uint_as_float(pk.x & 0xFFFF0000u);
                r2 -= __uint_as_float(pk.y << 16); r3 -= __uint_as_float(pk.y & 0xFFFF0000u);
                dst += plane_stride2;
            }
        }
    }
}

// ids[n, 0] = target[n], ids[n, 1 + k] = noise[n, k]: the id list of the NCE tables' gradient (what nce_fwd also emits as ids_out),
// available as soon as the noise is drawn so that its sort can start before the forward pass
__global__ void __launch_bounds__(256) nce_ids_concat_kernel(const int64_t* __restrict__ target, const int64_t* __restrict__ noise,
                                                             int64_t N, int K, int64_t* __restrict__ ids) {
    const int K1 = K + 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < N * K1; e += stride) {
        const int64_t n = e / K1;
        const int j = (int)(e - n * K1);
        ids[e] = (j == 0) ? __ldg(target + n) : __ldg(noise + n * K + (j - 1));
    }
}

}  // namespace mapb

extern "C" int map_nce_ids_concat(const int64_t* target, const int64_t* noise, int64_t N, int K, int64_t* ids_out,
                                  map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(target && noise && ids_out && N >= 0 && K >= 1, "map_nce_ids_concat: bad argument");
    if (N == 0) return MAP_OK;
    int64_t blocks = ceil_div(N * (K + 1), 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    nce_ids_concat_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(target, noise, N, K, ids_out);
    return check_launch("map_nce_ids_concat");
}

static int nce_fwd_launch(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                          const mapb::PeerTable& emb_t, const mapb::PeerTable& bias_t, int R, const float* logprob_noise,
                          float norm_term, int loss_type, float grad_scale, float* logits, int64_t* ids_out, float* loss_pos,
                          float* dz, float* d_input, int32_t* acc_count, map_stream_t stream, const char* who) {
    using namespace mapb;
    MAP_REQUIRE(input && target && noise && logprob_noise && logits && loss_pos && dz, "%s: null pointer", who);
    MAP_REQUIRE(N >= 0 && K >= 1, "%s: bad shape N=%lld K=%d", who, (long long)N, K);
    MAP_REQUIRE(loss_type == MAP_NCE_LOSS_NCE || loss_type == MAP_NCE_LOSS_SAMPLED, "%s: unknown loss_type %d", who, loss_type);
    MAP_REQUIRE(((uintptr_t)input % 16 == 0) && (!d_input || (uintptr_t)d_input % 16 == 0), "%s: input/d_input must be 16-byte aligned", who);
    for (int r = 0; r < R; ++r) MAP_REQUIRE((uintptr_t)emb_t.p[r] % 16 == 0, "%s: emb must be 16-byte aligned", who);
    if (N == 0) return MAP_OK;
    int64_t blocks = ceil_div(N, 8);  // 8 warps (positions) per CTA
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cudaStream_t st = as_stream(stream);
#define LAUNCH(LN)                                                                                                           \
    nce_fwd_kernel<LN><<<(unsigned)blocks, 256, 0, st>>>(input, N, K, target, noise, emb_t, bias_t, R, logprob_noise, norm_term, \
                                                         loss_type, grad_scale, logits, ids_out, loss_pos, dz, d_input, acc_count)
    switch (P) {
        case 4: LAUNCH(1); break;
        case 8: LAUNCH(2); break;
        case 16: LAUNCH(4); break;
        case 32: LAUNCH(8); break;
        case 64: LAUNCH(16); break;
        case 128: LAUNCH(32); break;
        default:
            set_error("%s: proj_size P=%d not in {4,8,16,32,64,128}", who, P);
            return MAP_EUNSUPPORTED;
    }
#undef LAUNCH
    return check_launch(who);
}

extern "C" int map_nce_fwd(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                           const float* emb, const float* bias, const float* logprob_noise, int64_t V, float norm_term,
                           int loss_type, float grad_scale, float* logits, int64_t* ids_out, float* loss_pos, float* dz,
                           float* d_input, int32_t* acc_count, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(emb && bias && V > 0, "map_nce_fwd: null table");
    PeerTable emb_t{}, bias_t{};
    emb_t.p[0] = emb;
    bias_t.p[0] = bias;
    return nce_fwd_launch(input, N, P, K, target, noise, emb_t, bias_t, 1, logprob_noise, norm_term, loss_type, grad_scale, logits,
                          ids_out, loss_pos, dz, d_input, acc_count, stream, "map_nce_fwd");
}

extern "C" int map_nce_fwd_sharded(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                                   const void* const* emb_shards, const void* const* bias_shards, int R,
                                   const float* logprob_noise, float norm_term, int loss_type, float grad_scale, float* logits,
                                   int64_t* ids_out, float* loss_pos, float* dz, float* d_input, int32_t* acc_count,
                                   map_stream_t stream) {
    using namespace mapb;
    PeerTable emb_t, bias_t;
    int rc = fill_peer_table(&emb_t, emb_shards, R, "map_nce_fwd_sharded");
    if (rc != MAP_OK) return rc;
    rc = fill_peer_table(&bias_t, bias_shards, R, "map_nce_fwd_sharded");
    if (rc != MAP_OK) return rc;
    return nce_fwd_launch(input, N, P, K, target, noise, emb_t, bias_t, R, logprob_noise, norm_term, loss_type, grad_scale, logits,
                          ids_out, loss_pos, dz, d_input, acc_count, stream, "map_nce_fwd_sharded");
}

extern "C" int map_gather_slices(const float* enc, const int64_t* masked_index, int64_t N, int L, int F, int P, float* sel,
                                 map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(enc && masked_index && sel && L >= 1 && F >= 1 && P >= 1 && N >= 0, "map_gather_slices: bad argument");
    if (N == 0) return MAP_OK;
    int64_t blocks = ceil_div(N * P, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    gather_slices_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(enc, masked_index, N, L, F, P, sel);
    return check_launch("map_gather_slices");
}

extern "C" int map_scatter_add_slices(const float* d_input, const int64_t* masked_index, int64_t N, int L, int F, int P,
                                      float* d_enc, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(d_input && masked_index && d_enc && L >= 1 && F >= 1 && P >= 1 && N >= 0, "map_scatter_add_slices: bad argument");
    if (N == 0) return MAP_OK;
    int64_t blocks = ceil_div(N * P, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    scatter_add_slices_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_input, masked_index, N, L, F, P, d_enc);
    return check_launch("map_scatter_add_slices");
}

extern "C" int map_expand_slices(const float* d_input, const int64_t* masked_index, int64_t B, int L, int F, int P, float* d_enc,
                                 int64_t ld_enc, uint16_t* planes, int64_t ld_p, int64_t plane_stride, int n_planes, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(d_input && masked_index && d_enc && L >= 1 && F >= 1 && P >= 4 && P % 4 == 0 && B >= 0, "map_expand_slices: bad argument (P % 4 == 0)");
    MAP_REQUIRE(ld_enc % 4 == 0 && ((uintptr_t)d_enc & 15) == 0 && ((uintptr_t)d_input & 15) == 0, "map_expand_slices: alignment");
    MAP_REQUIRE(planes == nullptr || (n_planes >= 1 && n_planes <= 3 && ld_p % 4 == 0 && plane_stride % 4 == 0 && ((uintptr_t)planes & 7) == 0),
                "map_expand_slices: planes alignment");
    if (B == 0) return MAP_OK;
    int64_t blocks = ceil_div(B * F * (P / 4), 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    expand_slices_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_input, masked_index, B, L, F, P / 4, d_enc, ld_enc,
                                                                           reinterpret_cast<uint32_t*>(planes), ld_p / 2, plane_stride / 2, n_planes);
    return check_launch("map_expand_slices");
}

// ------------------------------------------------------------------------------------------------ full-softmax cross entropy
// IndexLinear.ce_loss (code/nce/index_linear.py:145-151): score = F.linear(input, emb.weight, bias) over the WHOLE vocabulary,
// loss[n] = logsumexp_v(score[n, v]) - score[n, target[n]].  The reference materialises score [N, V] (53 GB at N = 12288,
// V = 1.09 M); here the scores never leave the SM: a CTA owns 64 positions and one slice of the vocabulary, computes 64 x 64
// score tiles (register-tiled fp32 FMAs, both operand tiles transposed in shared memory) and folds them into a running
// (max, sum of exp) per position; a second kernel merges the vocabulary slices and subtracts the target's score.
namespace mapb {

constexpr int kCeTile = 64;     // positions per CTA and vocabulary rows per tile
constexpr int kCeMaxP = 64;

__global__ void __launch_bounds__(256) nce_full_ce_partial_kernel(const float* __restrict__ input, int64_t N, int P,
                                                                  const float* __restrict__ emb, const float* __restrict__ bias,
                                                                  int64_t V, int64_t rows_per_split, float2* __restrict__ partial,
                                                                  int n_splits) {
    __shared__ float xs[kCeMaxP][kCeTile + 1];   // xs[k][position]
    __shared__ float es[kCeMaxP][kCeTile + 1];   // es[k][vocabulary row]
    __shared__ float bs[kCeTile];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads, 4 x 4 scores each
    const int64_t n0 = (int64_t)blockIdx.x * kCeTile;
    const int split = blockIdx.y;
    const int64_t v_begin = (int64_t)split * rows_per_split;
    int64_t v_end = v_begin + rows_per_split;
    if (v_end > V) v_end = V;
    for (int e = threadIdx.x; e < kCeTile * P; e += 256) {
        const int r = e / P, k = e - r * P;
        xs[k][r] = (n0 + r < N) ? input[(n0 + r) * P + k] : 0.f;
    }
    float m[4], s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; s[i] = 0.f; }
    for (int64_t v0 = v_begin; v0 < v_end; v0 += kCeTile) {
        __syncthreads();
        for (int e = threadIdx.x; e < kCeTile * P; e += 256) {
            const int r = e / P, k = e - r * P;
            es[k][r] = (v0 + r < v_end) ? __ldg(emb + (v0 + r) * P + k) : 0.f;
        }
        if (threadIdx.x < kCeTile) bs[threadIdx.x] = (v0 + threadIdx.x < v_end) ? __ldg(bias + v0 + threadIdx.x) : -INFINITY;
        __syncthreads();
        float acc[4][4] = {};
        for (int k = 0; k < P; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = es[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float tile_max = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] += bs[tx * 4 + j];     // rows past the end carry bias = -inf: exp() = 0
                tile_max = fmaxf(tile_max, acc[i][j]);
            }
            const float m_new = fmaxf(m[i], tile_max);
            if (m_new > -INFINITY) {
                float add = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) add += __expf(acc[i][j] - m_new);
                s[i] = s[i] * __expf(m[i] - m_new) + add;
                m[i] = m_new;
            }
        }
    }
    // the 16 threads that share a position (same ty: one half warp) merge their running pairs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m[i], o, 16);
            const float s2 = __shfl_xor_sync(0xffffffffu, s[i], o, 16);
            const float mn = fmaxf(m[i], m2);
            if (mn > -INFINITY) {
                s[i] = s[i] * __expf(m[i] - mn) + s2 * __expf(m2 - mn);
                m[i] = mn;
            }
        }
        const int64_t n = n0 + ty * 4 + i;
        if (tx == 0 && n < N) partial[n * n_splits + split] = make_float2(m[i], s[i]);
    }
}

__global__ void __launch_bounds__(256) nce_full_ce_merge_kernel(const float* __restrict__ input, int64_t N, int P,
                                                                const float* __restrict__ emb, const float* __restrict__ bias,
                                                                int64_t V, const int64_t* __restrict__ target,
                                                                const float2* __restrict__ partial, int n_splits, float* __restrict__ loss_pos) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float m = -INFINITY;
    for (int sp = 0; sp < n_splits; ++sp) m = fmaxf(m, partial[n * n_splits + sp].x);
    float sum = 0.f;
    for (int sp = 0; sp < n_splits; ++sp) {
        const float2 p = partial[n * n_splits + sp];
        if (p.x > -INFINITY) sum += p.y * expf(p.x - m);
    }
    const int64_t t = target[n];
    float score = (t >= 0 && t < V) ? bias[t] : 0.f;
    if (t >= 0 && t < V)
        for (int k = 0; k < P; ++k) score = fmaf(input[n * P + k], emb[t * P + k], score);
    loss_pos[n] = m + logf(sum) - score;
}

static int ce_splits(int64_t N, int64_t V) {
    const int64_t pos_tiles = ceil_div(N, kCeTile);
    int64_t splits = ceil_div((int64_t)kNumSMs * 4, pos_tiles);
    const int64_t max_splits = ceil_div(V, kCeTile);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 4096) splits = 4096;
    return (int)splits;
}

}  // namespace mapb

extern "C" size_t map_nce_full_ce_workspace_bytes(int64_t N, int64_t V) {
    return (size_t)N * (size_t)mapb::ce_splits(N, V) * sizeof(float2);
}

extern "C" int map_nce_full_ce(const float* input, int64_t N, int P, const float* emb, const float* bias, int64_t V,
                               const int64_t* target, float* loss_pos, void* workspace, size_t workspace_bytes, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(input && emb && bias && target && loss_pos && workspace, "map_nce_full_ce: null pointer");
    MAP_REQUIRE(N >= 0 && V >= 1 && P >= 1 && P <= kCeMaxP, "map_nce_full_ce: bad shape N=%lld V=%lld P=%d (P <= %d)", (long long)N,
                (long long)V, P, kCeMaxP);
    if (N == 0) return MAP_OK;
    const int splits = ce_splits(N, V);
    if (workspace_bytes < (size_t)N * splits * sizeof(float2)) {
        set_error("map_nce_full_ce: workspace %zu < %zu", workspace_bytes, (size_t)N * splits * sizeof(float2));
        return MAP_EWORKSPACE;
    }
    int64_t rows_per_split = ceil_div(V, splits);
    rows_per_split = ceil_div(rows_per_split, kCeTile) * kCeTile;
    dim3 grid((unsigned)ceil_div(N, kCeTile), (unsigned)splits);
    cudaStream_t st = as_stream(stream);
    nce_full_ce_partial_kernel<<<grid, 256, 0, st>>>(input, N, P, emb, bias, V, rows_per_split, static_cast<float2*>(workspace), splits);
    nce_full_ce_merge_kernel<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(input, N, P, emb, bias, V, target, static_cast<const float2*>(workspace),
                                                                         splits, loss_pos);
    return check_launch("map_nce_full_ce");
}
