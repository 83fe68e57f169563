// K13: owner-side kernels of the row-sharded tables.  Table row `id` lives on rank id % R at local row
// id / R.  The exchange steps around these kernels are fixed-size NCCL collectives issued by map_code_b200/dist.py:
//   embedding fwd : all-gather(ids)  -> emb_gather_owned (zeros for foreign rows) -> reduce-scatter(rows)
//   embedding bwd : all-gather(dE)   -> owned_keys + dedup + segment_reduce + row-wise AdamW on the shard
//   NCE           : "move the query, not the rows": all-gather(queries, ids) -> nce_scores_owned -> reduce-scatter(scores)
//                   -> nce_loss_from_scores (local) -> all-gather(dz) -> nce_dinput_owned -> reduce-scatter(d_query);
//                   table gradients of the owned rows through the same dedup pipeline.
// Integer / gather work, HBM-L2 bound; same lane mapping as embedding.cu / nce.cu.
#include "common.cuh"

namespace mapb {

// out[i,:] = (ids[i] % R == rank) ? shard[ids[i] / R, :] : 0      (vector lanes: D % 4 == 0)
__global__ void __launch_bounds__(256) emb_gather_owned_kernel(const float4* __restrict__ shard, int64_t rows, int vpr,
                                                               const int64_t* __restrict__ ids, int64_t n_vec, int R, int rank,
                                                               float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_vec; e += stride) {
        const int64_t r = e / vpr;
        const int lane = (int)(e - r * vpr);
        const int64_t id = __ldg(ids + r);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0 && (int)(id % R) == rank) {
            const int64_t lr = id / R;
            if (lr < rows) v = __ldg(shard + lr * vpr + lane);
        }
        st_stream_f4(out + e, v);
    }
}

__global__ void __launch_bounds__(256) emb_gather_owned_scalar_kernel(const float* __restrict__ shard, int64_t rows, int D,
                                                                      const int64_t* __restrict__ ids, int64_t n_elem, int R,
                                                                      int rank, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += stride) {
        const int64_t r = e / D;
        const int64_t id = __ldg(ids + r);
        float v = 0.f;
        if (id >= 0 && (int)(id % R) == rank && id / R < rows) v = __ldg(shard + (id / R) * D + (e - r * D));
        out[e] = v;
    }
}

// keys[i] = owned ? ids[i] / R : sentinel   (the sentinel collects every foreign occurrence in ONE trailing segment)
__global__ void __launch_bounds__(256) owned_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int R, int rank,
                                                         int64_t sentinel, int64_t* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int64_t id = ids[i];
        keys[i] = (id >= 0 && (int)(id % R) == rank) ? id / R : sentinel;
    }
}

// partial[n,j] = owned(idx[n,j]) ? <q[n,:], emb_shard[idx/R,:]> + bias_shard[idx/R] : 0
template <int LANES>
__global__ void __launch_bounds__(256) nce_scores_owned_kernel(const float* __restrict__ q, int64_t N, int K1,
                                                               const int64_t* __restrict__ idx, const float* __restrict__ emb,
                                                               const float* __restrict__ bias, int R, int rank,
                                                               float* __restrict__ partial) {
    constexpr int P = LANES * 4;
    constexpr int RPW = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane % LANES, grp = lane / LANES;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(q + n * P) + sub);
        const int steps = (K1 + RPW - 1) / RPW;
        for (int it = 0; it < steps; ++it) {
            const int j = it * RPW + grp;
            const bool valid = j < K1;
            int64_t id = -1;
            bool own = false;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                id = __ldg(idx + n * K1 + j);
                own = (int)(id % R) == rank;
                if (own) r = __ldg(reinterpret_cast<const float4*>(emb + (id / R) * P) + sub);
            }
            float s = x.x * r.x + x.y * r.y + x.z * r.z + x.w * r.w;
            s = group_sum<LANES>(s);
            if (valid && sub == 0) partial[n * K1 + j] = own ? s + __ldg(bias + id / R) : 0.f;
        }
    }
}

// from complete scores: logits, per-position loss, dz (x grad_scale), accuracy count — the tail of nce_fwd_kernel.
// One warp per position, lane j handles columns j, j+32, ...
__global__ void __launch_bounds__(256) nce_loss_from_scores_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx,
                                                                   int64_t N, int K1, const float* __restrict__ logq,
                                                                   float norm_term, int loss_type, float grad_scale,
                                                                   float* __restrict__ logits, float* __restrict__ loss_pos,
                                                                   float* __restrict__ dz, int32_t* __restrict__ acc_count) {
    const int lane = threadIdx.x & 31;
    const float ln_k = logf((float)(K1 - 1));
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        float loss = 0.f, mx_noise = -INFINITY, mx_c = -INFINITY;
        for (int j = lane; j < K1; j += 32) {
            const float lg = scores[n * K1 + j] - norm_term;
            logits[n * K1 + j] = lg;
            const float c = lg - __ldg(logq + idx[n * K1 + j]);
            if (j > 0) mx_noise = fmaxf(mx_noise, lg);
            if (loss_type == MAP_NCE_LOSS_NCE) {
                const float z = c - ln_k;
                const float y = (j == 0) ? 1.f : 0.f;
                loss += softplusf(z) - y * z;
                dz[n * K1 + j] = (sigmoidf(z) - y) * grad_scale;
            } else {
                mx_c = fmaxf(mx_c, c);
            }
        }
        if (loss_type != MAP_NCE_LOSS_NCE) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx_c = fmaxf(mx_c, __shfl_xor_sync(0xffffffffu, mx_c, o));
            float se = 0.f;
            for (int j = lane; j < K1; j += 32) se += expf(logits[n * K1 + j] - __ldg(logq + idx[n * K1 + j]) - mx_c);
            se = warp_sum(se);
            const float lse = mx_c + logf(se);
            for (int j = lane; j < K1; j += 32) {
                const float c = logits[n * K1 + j] - __ldg(logq + idx[n * K1 + j]);
                dz[n * K1 + j] = (expf(c - lse) - ((j == 0) ? 1.f : 0.f)) * grad_scale;
                if (j == 0) loss = lse - c;
            }
        }
        loss = warp_sum(loss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx_noise = fmaxf(mx_noise, __shfl_xor_sync(0xffffffffu, mx_noise, o));
        if (lane == 0) {
            loss_pos[n] = loss;
            const float t0 = scores[n * K1] - norm_term;
            if (acc_count != nullptr && !(mx_noise > t0)) atomicAdd(acc_count, 1);
        }
    }
}

// d_q_partial[n,:] = sum over owned j of dz[n,j] * emb_shard[idx[n,j]/R, :]
template <int LANES>
__global__ void __launch_bounds__(256) nce_dinput_owned_kernel(const float* __restrict__ dz, int64_t N, int K1,
                                                               const int64_t* __restrict__ idx, const float* __restrict__ emb,
                                                               int R, int rank, float* __restrict__ d_q) {
    constexpr int P = LANES * 4;
    constexpr int RPW = 32 / LANES;
    const int lane = threadIdx.x & 31, sub = lane % LANES, grp = lane / LANES;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int steps = (K1 + RPW - 1) / RPW;
        for (int it = 0; it < steps; ++it) {
            const int j = it * RPW + grp;
            if (j < K1) {
                const int64_t id = __ldg(idx + n * K1 + j);
                if ((int)(id % R) == rank) {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(emb + (id / R) * P) + sub);
                    const float g = __ldg(dz + n * K1 + j);
                    acc.x = fmaf(g, r.x, acc.x); acc.y = fmaf(g, r.y, acc.y); acc.z = fmaf(g, r.z, acc.z); acc.w = fmaf(g, r.w, acc.w);
                }
            }
        }
#pragma unroll
        for (int o = LANES; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (lane < LANES) reinterpret_cast<float4*>(d_q + n * P)[lane] = acc;
    }
}

}  // namespace mapb

extern "C" int map_emb_gather_owned_f32(const float* shard, int64_t shard_rows, int D, const int64_t* ids, int64_t n_ids, int R,
                                        int rank, float* out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(shard_rows > 0 && D > 0 && n_ids >= 0 && R >= 1 && rank >= 0 && rank < R, "map_emb_gather_owned_f32: bad argument");
    if (n_ids == 0) return MAP_OK;
    MAP_REQUIRE(shard && ids && out, "map_emb_gather_owned_f32: null pointer");
    const bool vec = (D % 4 == 0) && ((uintptr_t)shard % 16 == 0) && ((uintptr_t)out % 16 == 0);
    int64_t blocks = ceil_div(n_ids * (vec ? D / 4 : D), 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (vec)
        emb_gather_owned_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(shard), shard_rows, D / 4,
                                                                                 ids, n_ids * (D / 4), R, rank, reinterpret_cast<float4*>(out));
    else
        emb_gather_owned_scalar_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(shard, shard_rows, D, ids, n_ids * D, R, rank, out);
    return check_launch("map_emb_gather_owned_f32");
}

extern "C" int map_owned_keys(const int64_t* ids, int64_t n, int R, int rank, int64_t sentinel, int64_t* keys, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(n >= 0 && R >= 1 && rank >= 0 && rank < R, "map_owned_keys: bad argument");
    if (n == 0) return MAP_OK;
    MAP_REQUIRE(ids && keys, "map_owned_keys: null pointer");
    owned_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(ids, n, R, rank, sentinel, keys);
    return check_launch("map_owned_keys");
}

#define MAP_DISPATCH_P(P, CALL)                                                       \
    switch (P) {                                                                      \
        case 4: { constexpr int LN = 1; CALL; } break;                                \
        case 8: { constexpr int LN = 2; CALL; } break;                                \
        case 16: { constexpr int LN = 4; CALL; } break;                               \
        case 32: { constexpr int LN = 8; CALL; } break;                               \
        case 64: { constexpr int LN = 16; CALL; } break;                              \
        case 128: { constexpr int LN = 32; CALL; } break;                             \
        default: mapb::set_error("proj_size P=%d not in {4,8,16,32,64,128}", P); return MAP_EUNSUPPORTED; \
    }

extern "C" int map_nce_scores_owned(const float* q, int64_t N, int P, int K1, const int64_t* idx, const float* emb_shard,
                                    const float* bias_shard, int R, int rank, float* partial, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(N >= 0 && K1 >= 2 && R >= 1 && rank >= 0 && rank < R, "map_nce_scores_owned: bad argument");
    if (N == 0) return MAP_OK;
    MAP_REQUIRE(q && idx && emb_shard && bias_shard && partial, "map_nce_scores_owned: null pointer");
    int64_t blocks = ceil_div(N, 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cudaStream_t st = as_stream(stream);
    MAP_DISPATCH_P(P, (nce_scores_owned_kernel<LN><<<(unsigned)blocks, 256, 0, st>>>(q, N, K1, idx, emb_shard, bias_shard, R, rank, partial)));
    return check_launch("map_nce_scores_owned");
}

extern "C" int map_nce_loss_from_scores(const float* scores, const int64_t* idx, int64_t N, int K1, const float* logprob_noise,
                                        float norm_term, int loss_type, float grad_scale, float* logits, float* loss_pos, float* dz,
                                        int32_t* acc_count, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(N >= 0 && K1 >= 2, "map_nce_loss_from_scores: bad argument");
    MAP_REQUIRE(loss_type == MAP_NCE_LOSS_NCE || loss_type == MAP_NCE_LOSS_SAMPLED, "map_nce_loss_from_scores: unknown loss_type %d", loss_type);
    if (N == 0) return MAP_OK;
    MAP_REQUIRE(scores && idx && logprob_noise && logits && loss_pos && dz, "map_nce_loss_from_scores: null pointer");
    int64_t blocks = ceil_div(N, 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    nce_loss_from_scores_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(scores, idx, N, K1, logprob_noise, norm_term, loss_type,
                                                                                grad_scale, logits, loss_pos, dz, acc_count);
    return check_launch("map_nce_loss_from_scores");
}

extern "C" int map_nce_dinput_owned(const float* dz, int64_t N, int P, int K1, const int64_t* idx, const float* emb_shard, int R,
                                    int rank, float* d_q, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(N >= 0 && K1 >= 2 && R >= 1 && rank >= 0 && rank < R, "map_nce_dinput_owned: bad argument");
    if (N == 0) return MAP_OK;
    MAP_REQUIRE(dz && idx && emb_shard && d_q, "map_nce_dinput_owned: null pointer");
    int64_t blocks = ceil_div(N, 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cudaStream_t st = as_stream(stream);
    MAP_DISPATCH_P(P, (nce_dinput_owned_kernel<LN><<<(unsigned)blocks, 256, 0, st>>>(dz, N, K1, idx, emb_shard, R, rank, d_q)));
    return check_launch("map_nce_dinput_owned");
}
