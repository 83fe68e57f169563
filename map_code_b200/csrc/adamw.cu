// K12 + second half of K2: AdamW with transformers==4.26.1 semantics (call site code/trainer.py:75-76,140).
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= step_size * m / (sqrt(v) + eps) ; then p -= lr * wd * p
// Hyper-parameters live in device memory so a captured CUDA graph sees the scheduled learning rate of every replay.
#include <math.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace mapb {

struct Hyper {
    float lr, step_size, beta1, beta2, eps, one_m_beta1, one_m_beta2;
};
// hyper[6], hyper[7] = 1-beta evaluated in double like the reference (`alpha=1.0 - beta1`, `value=1.0 - beta2` are Python
// floats): float32(1 - 0.999) differs from 1.f - float32(0.999) by 1.3e-5 relative.
__device__ __forceinline__ Hyper load_hyper(const float* h) { return Hyper{h[0], h[1], h[2], h[3], h[4], h[6], h[7]}; }

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const Hyper& h, float lr_wd) {
    m = h.beta1 * m + h.one_m_beta1 * g;
    v = h.beta2 * v + h.one_m_beta2 * g * g;
    const float denom = sqrtf(v) + h.eps;
    p = p - h.step_size * (m / denom);
    p = p - lr_wd * p;  // lr_wd = lr * weight_decay (0 => no-op), applied after the Adam move like HF AdamW
}

__device__ __forceinline__ double sched_lambda(int sched, int64_t step, int64_t warmup, int64_t total) {
    if (step < warmup) return (double)step / (double)(warmup > 1 ? warmup : 1);
    if (sched == MAP_SCHED_COSINE) {
        const int64_t den = (total - warmup) > 1 ? (total - warmup) : 1;
        const double progress = (double)(step - warmup) / (double)den;
        const double c = 0.5 * (1.0 + cos(3.14159265358979323846 * 0.5 * 2.0 * progress));
        return c > 0.0 ? c : 0.0;
    }
    return 1.0;
}

__global__ void hyper_step_kernel(float* hyper, int64_t* step_counter, double base_lr, double beta1, double beta2, double eps,
                                  int sched, int64_t warmup, int64_t total) {
    const int64_t s = *step_counter;  // steps completed so far == scheduler's current_step
    const double lr = base_lr * sched_lambda(sched, s, warmup, total);
    const int64_t t = s + 1;  // optimizer state["step"] after increment
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    hyper[0] = (float)lr;
    hyper[1] = (float)(lr * sqrt(bc2) / bc1);
    hyper[2] = (float)beta1;
    hyper[3] = (float)beta2;
    hyper[4] = (float)eps;
    hyper[5] = (float)t;
    hyper[6] = (float)(1.0 - beta1);
    hyper[7] = (float)(1.0 - beta2);
    *step_counter = t;
}

__global__ void hyper_set_kernel(float* hyper, double lr, double beta1, double beta2, double eps, int64_t t) {
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    hyper[0] = (float)lr;
    hyper[1] = (float)(lr * sqrt(bc2) / bc1);
    hyper[2] = (float)beta1;
    hyper[3] = (float)beta2;
    hyper[4] = (float)eps;
    hyper[5] = (float)t;
    hyper[6] = (float)(1.0 - beta1);
    hyper[7] = (float)(1.0 - beta2);
}

// grid = (chunks, n_tensors); each CTA handles one 4096-element chunk of one tensor
constexpr int kAdamChunk = 4096;
__global__ void __launch_bounds__(256) adamw_multi_tensor_kernel(const map_adamw_tensor* __restrict__ tensors,
                                                                 const float* __restrict__ hyper) {
    const map_adamw_tensor t = tensors[blockIdx.y];
    const int64_t begin = (int64_t)blockIdx.x * kAdamChunk;
    if (begin >= t.n) return;
    const Hyper h = load_hyper(hyper);
    const float lr_wd = h.lr * t.weight_decay;
    const int64_t end = (begin + kAdamChunk < t.n) ? begin + kAdamChunk : t.n;
    const bool vec = ((t.n & 3) == 0) && ((((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0);
    if (vec) {
        for (int64_t i = begin + threadIdx.x * 4; i < end; i += 256 * 4) {
            float4 p = *reinterpret_cast<float4*>(t.p + i);
            const float4 g = ld_stream_f4(reinterpret_cast<const float4*>(t.g + i));
            float4 m = *reinterpret_cast<float4*>(t.m + i);
            float4 v = *reinterpret_cast<float4*>(t.v + i);
            adamw_elem(p.x, g.x, m.x, v.x, h, lr_wd);
            adamw_elem(p.y, g.y, m.y, v.y, h, lr_wd);
            adamw_elem(p.z, g.z, m.z, v.z, h, lr_wd);
            adamw_elem(p.w, g.w, m.w, v.w, h, lr_wd);
            *reinterpret_cast<float4*>(t.p + i) = p;
            *reinterpret_cast<float4*>(t.m + i) = m;
            *reinterpret_cast<float4*>(t.v + i) = v;
            if (t.n_planes > 0) {   // refresh the bf16 planes the split-bf16 GEMMs read: hi = bf16(p), lo = bf16(p - hi), ...
                float r0 = p.x, r1 = p.y, r2 = p.z, r3 = p.w;
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(t.planes) + i;
                for (int pl = 0; pl < t.n_planes; ++pl) {
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(r0), h1 = __float2bfloat16_rn(r1), h2 = __float2bfloat16_rn(r2), h3 = __float2bfloat16_rn(r3);
                    r0 -= __bfloat162float(h0); r1 -= __bfloat162float(h1); r2 -= __bfloat162float(h2); r3 -= __bfloat162float(h3);
                    uint2 pk;
                    pk.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                    pk.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
                    *reinterpret_cast<uint2*>(dst) = pk;
                    dst += t.plane_stride;
                }
            }
            if (t.p_t != nullptr) {
                const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t e = i + k;
                    const int64_t r = e / t.cols;
                    t.p_t[(e - r * t.cols) * (int64_t)t.rows + r] = pv[k];
                }
            }
        }
    } else {
        for (int64_t i = begin + threadIdx.x; i < end; i += 256) {
            float p = t.p[i], m = t.m[i], v = t.v[i];
            adamw_elem(p, t.g[i], m, v, h, lr_wd);
            t.p[i] = p; t.m[i] = m; t.v[i] = v;
            if (t.p_t != nullptr) {
                const int64_t r = i / t.cols;
                t.p_t[(i - r * t.cols) * (int64_t)t.rows + r] = p;
            }
        }
    }
}

// sparse: lanes = D/4 (vector) or D (scalar) threads per touched row
template <int VEC>
__global__ void __launch_bounds__(256) adamw_sparse_rows_kernel(float* __restrict__ table, float* __restrict__ m_,
                                                                float* __restrict__ v_, int D, int lanes,
                                                                const int64_t* __restrict__ uniq,
                                                                const float* __restrict__ grad,
                                                                const int32_t* __restrict__ n_unique,
                                                                const float* __restrict__ hyper, float weight_decay) {
    const int64_t U = *n_unique;
    const Hyper h = load_hyper(hyper);
    const float lr_wd = h.lr * weight_decay;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < U * lanes; t += stride) {
        const int64_t u = t / lanes;
        const int lane = (int)(t - u * lanes);
        const int64_t off = uniq[u] * D + lane * VEC;
        const int64_t goff = u * D + lane * VEC;
        if (VEC == 4) {
            float4 p = *reinterpret_cast<float4*>(table + off);
            float4 m = *reinterpret_cast<float4*>(m_ + off);
            float4 v = *reinterpret_cast<float4*>(v_ + off);
            const float4 g = *reinterpret_cast<const float4*>(grad + goff);
            adamw_elem(p.x, g.x, m.x, v.x, h, lr_wd);
            adamw_elem(p.y, g.y, m.y, v.y, h, lr_wd);
            adamw_elem(p.z, g.z, m.z, v.z, h, lr_wd);
            adamw_elem(p.w, g.w, m.w, v.w, h, lr_wd);
            *reinterpret_cast<float4*>(table + off) = p;
            *reinterpret_cast<float4*>(m_ + off) = m;
            *reinterpret_cast<float4*>(v_ + off) = v;
        } else {
            float p = table[off], m = m_[off], v = v_[off];
            adamw_elem(p, grad[goff], m, v, h, lr_wd);
            table[off] = p; m_[off] = m; v_[off] = v;
        }
    }
}

// dense_exact: every row of the table is updated; the gradient of a row is looked up in the sorted unique-id list
// (binary search, L2-resident) and is 0 when the row was not touched this step.
template <int VEC>
__global__ void __launch_bounds__(256) adamw_dense_rows_kernel(float* __restrict__ table, float* __restrict__ m_,
                                                               float* __restrict__ v_, int64_t V, int D, int lanes,
                                                               const int64_t* __restrict__ uniq,
                                                               const float* __restrict__ grad,
                                                               const int32_t* __restrict__ n_unique,
                                                               const float* __restrict__ hyper, float weight_decay) {
    const int U = *n_unique;
    const Hyper h = load_hyper(hyper);
    const float lr_wd = h.lr * weight_decay;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < V * lanes; t += stride) {
        const int64_t row = t / lanes;
        const int lane = (int)(t - row * lanes);
        int lo = 0, hi = U;  // first index with uniq[idx] >= row
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(uniq + mid) < row) lo = mid + 1; else hi = mid;
        }
        const bool hit = (lo < U) && (__ldg(uniq + lo) == row);
        const int64_t off = row * D + lane * VEC;
        if (VEC == 4) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hit) g = *reinterpret_cast<const float4*>(grad + (int64_t)lo * D + lane * VEC);
            float4 p = *reinterpret_cast<float4*>(table + off);
            float4 m = *reinterpret_cast<float4*>(m_ + off);
            float4 v = *reinterpret_cast<float4*>(v_ + off);
            adamw_elem(p.x, g.x, m.x, v.x, h, lr_wd);
            adamw_elem(p.y, g.y, m.y, v.y, h, lr_wd);
            adamw_elem(p.z, g.z, m.z, v.z, h, lr_wd);
            adamw_elem(p.w, g.w, m.w, v.w, h, lr_wd);
            *reinterpret_cast<float4*>(table + off) = p;
            *reinterpret_cast<float4*>(m_ + off) = m;
            *reinterpret_cast<float4*>(v_ + off) = v;
        } else {
            const float g = hit ? grad[(int64_t)lo * D + lane] : 0.f;
            float p = table[off], m = m_[off], v = v_[off];
            adamw_elem(p, g, m, v, h, lr_wd);
            table[off] = p; m_[off] = m; v_[off] = v;
        }
    }
}

}  // namespace mapb

extern "C" int map_adamw_hyper_step(float* hyper, int64_t* step_counter, double base_lr, double beta1, double beta2, double eps,
                                    int sched, int64_t warmup_steps, int64_t total_steps, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(hyper && step_counter, "map_adamw_hyper_step: null pointer");
    MAP_REQUIRE(sched == MAP_SCHED_CONST || sched == MAP_SCHED_COSINE, "map_adamw_hyper_step: unknown schedule %d", sched);
    hyper_step_kernel<<<1, 1, 0, as_stream(stream)>>>(hyper, step_counter, base_lr, beta1, beta2, eps, sched, warmup_steps, total_steps);
    return check_launch("map_adamw_hyper_step");
}

extern "C" int map_adamw_hyper_set(float* hyper, double lr, double beta1, double beta2, double eps, int64_t step, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(hyper && step >= 1, "map_adamw_hyper_set: bad argument");
    hyper_set_kernel<<<1, 1, 0, as_stream(stream)>>>(hyper, lr, beta1, beta2, eps, step);
    return check_launch("map_adamw_hyper_set");
}

extern "C" int map_adamw_multi_tensor(const map_adamw_tensor* tensors_dev, int n_tensors, int64_t max_elems_per_tensor,
                                      const float* hyper, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(tensors_dev && hyper && n_tensors >= 1 && n_tensors <= 65535 && max_elems_per_tensor >= 1,
                "map_adamw_multi_tensor: bad argument");
    dim3 grid((unsigned)ceil_div(max_elems_per_tensor, kAdamChunk), (unsigned)n_tensors);
    adamw_multi_tensor_kernel<<<grid, 256, 0, as_stream(stream)>>>(tensors_dev, hyper);
    return check_launch("map_adamw_multi_tensor");
}

extern "C" int map_adamw_sparse_rows(float* table, float* m, float* v, int D, const int64_t* uniq_ids, const float* grad_compact,
                                     const int32_t* n_unique, int64_t max_unique, const float* hyper, float weight_decay,
                                     map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(table && m && v && uniq_ids && grad_compact && n_unique && hyper && D >= 1 && max_unique >= 1,
                "map_adamw_sparse_rows: bad argument");
    const bool vec = (D % 4 == 0) && ((((uintptr_t)table | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad_compact) & 15) == 0);
    const int lanes = vec ? D / 4 : D;
    int64_t blocks = ceil_div(max_unique * lanes, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (vec)
        adamw_sparse_rows_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, m, v, D, lanes, uniq_ids, grad_compact, n_unique, hyper, weight_decay);
    else
        adamw_sparse_rows_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, m, v, D, lanes, uniq_ids, grad_compact, n_unique, hyper, weight_decay);
    return check_launch("map_adamw_sparse_rows");
}

extern "C" int map_adamw_dense_rows_sparse_grad(float* table, float* m, float* v, int64_t V, int D, const int64_t* uniq_ids,
                                                const float* grad_compact, const int32_t* n_unique, const float* hyper,
                                                float weight_decay, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(table && m && v && uniq_ids && grad_compact && n_unique && hyper && D >= 1 && V >= 1,
                "map_adamw_dense_rows_sparse_grad: bad argument");
    const bool vec = (D % 4 == 0) && ((((uintptr_t)table | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad_compact) & 15) == 0);
    const int lanes = vec ? D / 4 : D;
    int64_t blocks = ceil_div(V * lanes, 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (vec)
        adamw_dense_rows_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, m, v, V, D, lanes, uniq_ids, grad_compact, n_unique, hyper, weight_decay);
    else
        adamw_dense_rows_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(table, m, v, V, D, lanes, uniq_ids, grad_compact, n_unique, hyper, weight_decay);
    return check_launch("map_adamw_dense_rows_sparse_grad");
}
