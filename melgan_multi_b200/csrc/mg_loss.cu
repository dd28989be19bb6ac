// Fused loss reductions (models.py:138-167): feature_loss is 21 mean-|r - g| terms, the LSGAN losses are 3 (or 6) mean
// squares -- in the reference ~4 elementwise/reduction launches per term plus a host sync per logged value.  Here every
// term of one loss is a row of a table and ONE launch reduces all rows (a second, tiny one combines the per-CTA partial
// sums in a fixed order, so the result is bit-reproducible); the backward of all rows is one launch too.
//   mode 0: mean |a - b|        (feature_loss term, models.py:142)
//   mode 1: mean (1 - a)^2      (real term of discriminator_loss :151, generator_loss :165)
//   mode 2: mean a^2            (generated term of discriminator_loss :152)
#include "mg_common.cuh"

namespace mg {

constexpr int kLossMaxRows = 24;
constexpr int kLossChunk = 16384;  // elements per CTA

struct LossArgs {
    const float *a[kLossMaxRows];
    const float *b[kLossMaxRows];
    float *ga[kLossMaxRows];
    float *gb[kLossMaxRows];
    long long n[kLossMaxRows];
    int mode[kLossMaxRows];
    int first[kLossMaxRows + 1];  // first CTA of each row
    int count;
};

__device__ __forceinline__ float loss_term(int mode, float a, float b) {
    return mode == 0 ? fabsf(a - b) : mode == 1 ? (1.f - a) * (1.f - a) : a * a;
}

__device__ __forceinline__ int loss_row(const LossArgs &t, int cta) {
    int i = 0;
    while (i + 1 < t.count && cta >= t.first[i + 1]) ++i;
    return i;
}

__global__ void __launch_bounds__(256) loss_partial_kernel(const __grid_constant__ LossArgs t, float *__restrict__ partial) {
    const int i = loss_row(t, blockIdx.x);
    const long long base = (long long)(blockIdx.x - t.first[i]) * kLossChunk;
    const long long end = base + kLossChunk < t.n[i] ? base + kLossChunk : t.n[i];
    const float *a = t.a[i], *b = t.b[i];
    const int mode = t.mode[i];
    float s = 0.f;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | (mode == 0 ? reinterpret_cast<uintptr_t>(b) : 0)) & 15) == 0;
    if (vec) {  // base is a multiple of 4 elements: aligned float4 loads, scalar tail
        const long long end4 = base + ((end - base) & ~3ll);
        for (long long j = base + 4 * threadIdx.x; j < end4; j += 4 * 256) {
            const float4 x = *reinterpret_cast<const float4 *>(a + j);
            const float4 y = mode == 0 ? *reinterpret_cast<const float4 *>(b + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (loss_term(mode, x.x, y.x) + loss_term(mode, x.y, y.y)) + (loss_term(mode, x.z, y.z) + loss_term(mode, x.w, y.w));
        }
        for (long long j = end4 + threadIdx.x; j < end; j += 256) s += loss_term(mode, a[j], mode == 0 ? b[j] : 0.f);
    } else {
        for (long long j = base + threadIdx.x; j < end; j += 256) s += loss_term(mode, a[j], mode == 0 ? b[j] : 0.f);
    }
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}

// one warp per row: fixed-order combination of the row's partial sums -> out[row] = mean
__global__ void __launch_bounds__(32) loss_final_kernel(const __grid_constant__ LossArgs t, const float *__restrict__ partial,
                                                        float *__restrict__ out) {
    const int i = blockIdx.x;
    double s = 0.0;
    for (int c = t.first[i] + threadIdx.x; c < t.first[i + 1]; c += 32) s += (double)partial[c];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) out[i] = (float)(s / (double)t.n[i]);
}

// d out[i] / d a = gout[i] / n * {sign(a - b), -2 (1 - a), 2 a};  d / d b = -sign(a - b) (mode 0 only)
__global__ void __launch_bounds__(256) loss_backward_kernel(const __grid_constant__ LossArgs t, const float *__restrict__ gout) {
    const int i = loss_row(t, blockIdx.x);
    const long long base = (long long)(blockIdx.x - t.first[i]) * kLossChunk;
    const long long end = base + kLossChunk < t.n[i] ? base + kLossChunk : t.n[i];
    const float *a = t.a[i], *b = t.b[i];
    float *ga = t.ga[i], *gb = t.gb[i];
    const int mode = t.mode[i];
    const float sc = gout[i] / (float)t.n[i];
    for (long long j = base + threadIdx.x; j < end; j += 256) {
        const float x = a[j];
        if (mode == 0) {
            const float d = x - b[j];
            const float g = d > 0.f ? sc : d < 0.f ? -sc : 0.f;
            if (ga) ga[j] = g;
            if (gb) gb[j] = -g;
        } else if (ga) {
            ga[j] = mode == 1 ? -2.f * (1.f - x) * sc : 2.f * x * sc;
        }
    }
}

static int fill_args(LossArgs &t, const float *const *a, const float *const *b, const long long *n, const int *mode, int count,
                     const char *fn) {
    if (!a || !n || !mode || count < 1 || count > kLossMaxRows)
        return set_error(MG_ERR_INVALID_ARGUMENT, "%s: need 1..%d rows", fn, kLossMaxRows);
    long long ctas = 0;
    for (int i = 0; i < count; ++i) {
        if (!a[i] || n[i] < 1 || mode[i] < 0 || mode[i] > 2 || (mode[i] == 0 && (!b || !b[i])))
            return set_error(MG_ERR_INVALID_ARGUMENT, "%s: bad row %d", fn, i);
        t.a[i] = a[i]; t.b[i] = (mode[i] == 0) ? b[i] : nullptr; t.n[i] = n[i]; t.mode[i] = mode[i];
        t.ga[i] = t.gb[i] = nullptr;
        t.first[i] = (int)ctas;
        ctas += (n[i] + kLossChunk - 1) / kLossChunk;
        if (ctas > 0x7fffffffll) return set_error(MG_ERR_INVALID_ARGUMENT, "%s: too many elements", fn);
    }
    t.first[count] = (int)ctas;
    t.count = count;
    return MG_OK;
}

long long loss_num_ctas(const long long *n, int count) {
    long long c = 0;
    for (int i = 0; i < count; ++i) c += (n[i] + kLossChunk - 1) / kLossChunk;
    return c;
}

int launch_loss_forward(const float *const *a, const float *const *b, const long long *n, const int *mode, int count, float *out,
                        float *partial, cudaStream_t s) {
    LossArgs t;
    int rc = fill_args(t, a, b, n, mode, count, "mg_loss_forward");
    if (rc) return rc;
    loss_partial_kernel<<<t.first[count], 256, 0, s>>>(t, partial);
    MG_CUDA_TRY(cudaGetLastError());
    loss_final_kernel<<<count, 32, 0, s>>>(t, partial, out);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

int launch_loss_backward(const float *const *a, const float *const *b, const long long *n, const int *mode, int count,
                         const float *gout, float *const *ga, float *const *gb, cudaStream_t s) {
    LossArgs t;
    int rc = fill_args(t, a, b, n, mode, count, "mg_loss_backward");
    if (rc) return rc;
    for (int i = 0; i < count; ++i) {
        t.ga[i] = ga ? ga[i] : nullptr;
        t.gb[i] = (gb && mode[i] == 0) ? gb[i] : nullptr;
    }
    loss_backward_kernel<<<t.first[count], 256, 0, s>>>(t, gout);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

}  // namespace mg
