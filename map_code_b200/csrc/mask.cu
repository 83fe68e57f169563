// K8 / K9: Trainer.dynamic_mask on the device (code/trainer.py:217-266).  One warp per sample row: the row of ids is
// loaded coalesced into registers (F <= 64 -> two values per lane), every lane replays the L draws of its row, the lane
// that owns the drawn field applies the write -> sequential per row => last-writer-wins for duplicate fields, exactly
// like the reference's CPU scatter.  Integer work, bound by the B*F*8-byte id traffic.
#include "common.cuh"

namespace mapb {

constexpr int kMaxFields = 64;

// masked_index[b, l]; element index for Philox = (row0 + b) * L + l
// RNG offsets: offset_eff = offset + STREAMS_PER_STEP * (*step_dev) when step_dev != NULL, so a captured CUDA graph
// draws a fresh subsequence at every replay (the device step counter is advanced by map_adamw_hyper_step).
constexpr uint64_t kStreamsPerStep = 8;
__device__ __forceinline__ uint64_t eff_offset(uint64_t offset, const int64_t* step_dev) {
    return step_dev != nullptr ? offset + kStreamsPerStep * (uint64_t)(*step_dev) : offset;
}

__global__ void __launch_bounds__(256) mask_index_kernel(int64_t* __restrict__ mi, int64_t B, int L, int F, int method,
                                                         uint64_t seed, uint64_t offset, int64_t row0,
                                                         const int64_t* __restrict__ step_dev) {
    offset = eff_offset(offset, step_dev);
    if (method == MAP_SAMPLING_RANDINT) {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= B * L) return;
        const int64_t b = i / L;
        const int l = (int)(i - b * L);
        const Philox4 r = philox_elem(seed, offset, (uint64_t)(row0 + b) * (uint64_t)L + (uint64_t)l);
        mi[i] = (int64_t)bounded32(r.w0, (uint32_t)F);
    } else {  // randperm(F)[:L]: the first L steps of a Fisher-Yates shuffle, one thread per row
        const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (b >= B) return;
        unsigned char perm[kMaxFields];
        for (int f = 0; f < F; ++f) perm[f] = (unsigned char)f;
        for (int l = 0; l < L; ++l) {
            const Philox4 r = philox_elem(seed, offset, (uint64_t)(row0 + b) * (uint64_t)L + (uint64_t)l);
            const int j = l + (int)bounded32(r.w0, (uint32_t)(F - l));
            const unsigned char t = perm[l];
            perm[l] = perm[j];
            perm[j] = t;
            mi[b * L + l] = (int64_t)perm[l];
        }
    }
}

__global__ void __launch_bounds__(256) mfp_apply_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mi,
                                                        int64_t B, int F, int L, int64_t mask_id,
                                                        int64_t* __restrict__ ids_out, int64_t* __restrict__ labels) {
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int64_t* row = ids + b * F;
    int64_t v0 = (lane < F) ? row[lane] : 0;
    int64_t v1 = (lane + 32 < F) ? row[lane + 32] : 0;
    const int64_t o0 = v0, o1 = v1;
    for (int l = 0; l < L; ++l) {
        const int f = (int)mi[b * L + l];
        // labels are gathered from the ORIGINAL ids (gather happens before scatter, trainer.py:230-231)
        const int64_t src = (f < 32) ? __shfl_sync(0xffffffffu, o0, f & 31) : __shfl_sync(0xffffffffu, o1, f & 31);
        if (lane == 0) labels[b * L + l] = src;
        if (f == lane) v0 = mask_id;
        if (f == lane + 32) v1 = mask_id;
    }
    if (lane < F) ids_out[b * F + lane] = v0;
    if (lane + 32 < F) ids_out[b * F + lane + 32] = v1;
}

__global__ void __launch_bounds__(256) rfd_replace_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ mi,
                                                          int64_t B, int F, int L, int mode,
                                                          const int64_t* __restrict__ x_train, int64_t n_train,
                                                          const int64_t* __restrict__ idx_low,
                                                          const int64_t* __restrict__ idx_high, int64_t input_size,
                                                          uint64_t seed, uint64_t off_rep, uint64_t off_f2, int64_t row0,
                                                          const int64_t* __restrict__ step_dev,
                                                          int64_t* __restrict__ ids_out, float* __restrict__ labels,
                                                          int64_t* __restrict__ rep_out) {
    off_rep = eff_offset(off_rep, step_dev);
    off_f2 = eff_offset(off_f2, step_dev);
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int64_t* row = ids + b * F;
    const int64_t o0 = (lane < F) ? row[lane] : 0;
    const int64_t o1 = (lane + 32 < F) ? row[lane + 32] : 0;
    int64_t v0 = o0, v1 = o1;
    for (int l = 0; l < L; ++l) {
        const int f = (int)mi[b * L + l];
        const uint64_t elem = (uint64_t)(row0 + b) * (uint64_t)L + (uint64_t)l;
        const Philox4 r = philox_elem(seed, off_rep, elem);  // every lane computes the same draw (uniform branch)
        int64_t rep;
        if (mode == MAP_RFD_UNIGRAM) {
            const int64_t si = (int64_t)bounded64(r, (uint64_t)n_train);
            rep = __ldg(x_train + si * F + f);
        } else if (mode == MAP_RFD_UNIFORM) {
            const int64_t lo = __ldg(idx_low + f);
            rep = lo + (int64_t)bounded64(r, (uint64_t)(__ldg(idx_high + f) - lo));
        } else if (mode == MAP_RFD_WHOLE_UNIFORM) {
            rep = 10 + (int64_t)bounded64(r, (uint64_t)(input_size - 10));
        } else {  // MAP_RFD_WHOLE_UNIGRAM
            const int64_t si = (int64_t)bounded64(r, (uint64_t)n_train);
            const Philox4 r2 = philox_elem(seed, off_f2, elem);
            rep = __ldg(x_train + si * F + (int)bounded32(r2.w0, (uint32_t)F));
        }
        if (rep_out != nullptr && lane == 0) rep_out[b * L + l] = rep;
        if (f == lane) v0 = rep;
        if (f == lane + 32) v1 = rep;
    }
    if (lane < F) {
        ids_out[b * F + lane] = v0;
        labels[b * F + lane] = (v0 != o0) ? 1.f : 0.f;
    }
    if (lane + 32 < F) {
        ids_out[b * F + lane + 32] = v1;
        labels[b * F + lane + 32] = (v1 != o1) ? 1.f : 0.f;
    }
}

}  // namespace mapb

extern "C" int map_mask_index_philox(int64_t* masked_index, int64_t B, int L, int F, int sampling_method, uint64_t seed,
                                     uint64_t offset, int64_t row0, const int64_t* step_dev, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(masked_index && B >= 0 && L >= 0 && F >= 1 && F <= kMaxFields && L <= F, "map_mask_index_philox: bad shape B=%lld L=%d F=%d", (long long)B, L, F);
    if (sampling_method != MAP_SAMPLING_RANDINT && sampling_method != MAP_SAMPLING_NORMAL) {
        set_error("map_mask_index_philox: unknown sampling_method %d", sampling_method);  // trainer.py:226-227 NotImplementedError
        return MAP_EUNSUPPORTED;
    }
    if (B == 0 || L == 0) return MAP_OK;
    const int64_t threads = (sampling_method == MAP_SAMPLING_RANDINT) ? B * L : B;
    mask_index_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, as_stream(stream)>>>(masked_index, B, L, F, sampling_method, seed, offset, row0, step_dev);
    return check_launch("map_mask_index_philox");
}

extern "C" int map_mfp_mask_apply(const int64_t* ids, const int64_t* masked_index, int64_t B, int F, int L, int64_t mask_id,
                                  int64_t* ids_out, int64_t* labels, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(ids && ids_out && (L == 0 || (masked_index && labels)), "map_mfp_mask_apply: null pointer");
    MAP_REQUIRE(B >= 0 && F >= 1 && F <= kMaxFields && L >= 0, "map_mfp_mask_apply: bad shape");
    if (B == 0) return MAP_OK;
    mfp_apply_kernel<<<(unsigned)ceil_div(B * 32, 256), 256, 0, as_stream(stream)>>>(ids, masked_index, B, F, L, mask_id, ids_out, labels);
    return check_launch("map_mfp_mask_apply");
}

extern "C" int map_rfd_replace_philox(const int64_t* ids, const int64_t* masked_index, int64_t B, int F, int L, int mode,
                                      const int64_t* x_train, int64_t n_train, const int64_t* idx_low, const int64_t* idx_high,
                                      int64_t input_size, uint64_t seed, uint64_t offset_replace, uint64_t offset_field2,
                                      int64_t row0, const int64_t* step_dev, int64_t* ids_out, float* labels,
                                      int64_t* replace_feat_out, map_stream_t stream) {
    using namespace mapb;
    MAP_REQUIRE(ids && ids_out && labels && (L == 0 || masked_index), "map_rfd_replace_philox: null pointer");
    MAP_REQUIRE(B >= 0 && F >= 1 && F <= kMaxFields && L >= 0, "map_rfd_replace_philox: bad shape");
    switch (mode) {
        case MAP_RFD_UNIGRAM:
        case MAP_RFD_WHOLE_UNIGRAM:
            MAP_REQUIRE(x_train && n_train > 0, "map_rfd_replace_philox: Unigram modes need the training matrix");
            break;
        case MAP_RFD_UNIFORM:
            MAP_REQUIRE(idx_low && idx_high, "map_rfd_replace_philox: Uniform needs idx_low/idx_high");
            break;
        case MAP_RFD_WHOLE_UNIFORM:
            MAP_REQUIRE(input_size > 10, "map_rfd_replace_philox: Whole-Uniform needs input_size > 10");
            break;
        default:
            set_error("map_rfd_replace_philox: unknown RFD_replace mode %d", mode);  // trainer.py:261-262
            return MAP_EUNSUPPORTED;
    }
    if (B == 0) return MAP_OK;
    rfd_replace_kernel<<<(unsigned)ceil_div(B * 32, 256), 256, 0, as_stream(stream)>>>(
        ids, masked_index, B, F, L, mode, x_train, n_train, idx_low, idx_high, input_size, seed, offset_replace, offset_field2,
        row0, step_dev, ids_out, labels, replace_feat_out);
    return check_launch("map_rfd_replace_philox");
}
