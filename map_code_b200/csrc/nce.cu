// K6 / K7: fused NCE head.  Replaces NCELoss.forward + IndexLinear._compute_sampled_logit + nce_loss /
// sampled_softmax_loss (code/nce/nce_loss.py:79-144,201-244; code/nce/index_linear.py:68-106).
// The "contraction" is a gather-dot (every position owns its K+1 rows): HBM/L2-bound, no tensor cores.
// One warp per position; LANES = P/4 lanes own one gathered row (16-byte loads), 32/LANES rows in flight per step.
// Forward and the input-side backward are produced in the same pass while the rows are in registers.
#include "common.cuh"

namespace mapb {

template <int LANES>
__global__ void __launch_bounds__(256) nce_fwd_kernel(const float* __restrict__ input, int64_t N, int K,
                                                      const int64_t* __restrict__ target, const int64_t* __restrict__ noise,
                                                      const float* __restrict__ emb, const float* __restrict__ bias,
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
        if (loss_type == MAP_NCE_LOSS_NCE) {
            for (int it = 0; it < steps; ++it) {
                const int j = it * RPW + grp;
                const bool valid = j < K1;
                int64_t idx = 0;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    idx = (j == 0) ? __ldg(target + n) : __ldg(noise + n * K + (j - 1));
                    r = __ldg(reinterpret_cast<const float4*>(emb + idx * P) + sub);
                }
                float s = x.x * r.x + x.y * r.y + x.z * r.z + x.w * r.w;
                s = group_sum<LANES>(s);
                if (valid) {
                    const float lg = s + __ldg(bias + idx) - norm_term;     // nce_loss.py:171-172
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
                    r = __ldg(reinterpret_cast<const float4*>(emb + idx * P) + sub);
                }
                float s = x.x * r.x + x.y * r.y + x.z * r.z + x.w * r.w;
                s = group_sum<LANES>(s);
                if (valid) {
                    const float lg = s + __ldg(bias + idx) - norm_term;
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
                    const float4 r = __ldg(reinterpret_cast<const float4*>(emb + idx * P) + sub);
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

}  // namespace mapb

extern "C" int map_nce_fwd(const float* input, int64_t N, int P, int K, const int64_t* target, const int64_t* noise,
                           const float* emb, const float* bias, const float* logprob_noise, int64_t V, float norm_term,
                           int loss_type, float grad_scale, float* logits, int64_t* ids_out, float* loss_pos, float* dz,
                           float* d_input, int32_t* acc_count, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(input && target && noise && emb && bias && logprob_noise && logits && loss_pos && dz, "map_nce_fwd: null pointer");
    MAP_REQUIRE(N >= 0 && K >= 1 && V > 0, "map_nce_fwd: bad shape N=%lld K=%d", (long long)N, K);
    MAP_REQUIRE(loss_type == MAP_NCE_LOSS_NCE || loss_type == MAP_NCE_LOSS_SAMPLED, "map_nce_fwd: unknown loss_type %d", loss_type);
    MAP_REQUIRE(((uintptr_t)input % 16 == 0) && ((uintptr_t)emb % 16 == 0) && (!d_input || (uintptr_t)d_input % 16 == 0),
                "map_nce_fwd: input/emb/d_input must be 16-byte aligned");
    if (N == 0) return MAP_OK;
    int64_t blocks = ceil_div(N, 8);  // 8 warps (positions) per CTA
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cudaStream_t st = as_stream(stream);
#define LAUNCH(LN)                                                                                                        \
    nce_fwd_kernel<LN><<<(unsigned)blocks, 256, 0, st>>>(input, N, K, target, noise, emb, bias, logprob_noise, norm_term, \
                                                         loss_type, grad_scale, logits, ids_out, loss_pos, dz, d_input, acc_count)
    switch (P) {
        case 4: LAUNCH(1); break;
        case 8: LAUNCH(2); break;
        case 16: LAUNCH(4); break;
        case 32: LAUNCH(8); break;
        case 64: LAUNCH(16); break;
        case 128: LAUNCH(32); break;
        default:
            set_error("map_nce_fwd: proj_size P=%d not in {4,8,16,32,64,128}", P);
            return MAP_EUNSUPPORTED;
    }
#undef LAUNCH
    return check_launch("map_nce_fwd");
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
