// Weight-norm fold + packing of all 30 generator layers in ONE launch.
//
// Replaces the per-layer weight_norm pre-forward hook of the reference (w = g * v / ||v||,
// norm over every axis but 0; models.py:16-28,46-59).  One CTA per norm row (4353 rows):
// the CTA reduces ||v_row||^2, then scatters g/||v|| * v into the packed layout of mg_layout.h.
#include "mg_common.cuh"
#include "mg_tc.cuh"

namespace mg {

struct PackArgs {
    const float *v[kNumLayers];
    const float *g[kNumLayers];
    const float *bias[kNumLayers];
};

struct RowTable {
    int first_row[kNumLayers + 1];
};

static RowTable make_row_table() {
    RowTable t;
    int r = 0;
    for (int l = 0; l < kNumLayers; ++l) {
        t.first_row[l] = r;
        r += layer_norm_rows(l);
    }
    t.first_row[kNumLayers] = r;
    return t;
}

__global__ void __launch_bounds__(128) pack_kernel(PackArgs a, RowTable rt, float *__restrict__ packed) {
    const int grow = blockIdx.x;
    int l = 0;
#pragma unroll 1
    while (grow >= rt.first_row[l + 1]) ++l;
    const int row = grow - rt.first_row[l];
    const LayerShape sh = layer_shape(l);
    const int inner = (sh.kind == 0 ? sh.cin : sh.cout) * sh.k;
    const float *__restrict__ vr = a.v[l] + (size_t)row * inner;

    float ss = 0.f;
    for (int j = threadIdx.x; j < inner; j += blockDim.x) {
        float e = vr[j];
        ss = fmaf(e, e, ss);
    }
    __shared__ float red[4];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    const float total = red[0] + red[1] + red[2] + red[3];
    const float scale = a.g[l][row] / sqrtf(total);

    float *__restrict__ wp = packed + weight_offset(l);
    if (sh.kind == 0) {
        // v[co=row][ci][k] -> wp[(ci*K + k)*Cout + co]
        for (int j = threadIdx.x; j < inner; j += blockDim.x) wp[(size_t)j * sh.cout + row] = scale * vr[j];
        if (l == 0) {
            // split-bf16 copy of conv_pre for conv_rows_tc_kernel (layout: mg_layout.h, conv_tc_weight_index)
            __nv_bfloat16 *tcw = reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<char *>(packed) + tc_region_start() + tc_pre_offset());
            for (int j = threadIdx.x; j < inner; j += blockDim.x) {
                const int ci = j / kPreK, tap = j - kPreK * ci;
                __nv_bfloat16 hi, lo;
                tc::split_bf16(scale * vr[j], hi, lo);
                tcw[conv_tc_weight_index(kMelBins, kPreK, kPreNG, row, ci, tap, 0)] = hi;
                tcw[conv_tc_weight_index(kMelBins, kPreK, kPreNG, row, ci, tap, 1)] = lo;
            }
        }
        if (l >= 5 && l <= 28) {
            // split-bf16 copy for the tensor-core ResBlock kernels (layout: mg_layout.h, tc_weight_index)
            __nv_bfloat16 *tcw = reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<char *>(packed) + tc_region_start() +
                                                                  tc_res_offset(l));
            for (int j = threadIdx.x; j < inner; j += blockDim.x) {
                const int ci = j / 3, tap = j - 3 * ci;
                __nv_bfloat16 hi, lo;
                tc::split_bf16(scale * vr[j], hi, lo);
                tcw[tc_weight_index(sh.cout, row, ci, tap, 0)] = hi;
                tcw[tc_weight_index(sh.cout, row, ci, tap, 1)] = lo;
            }
        }
    } else {
        // v[ci=row][co][k] -> wp[((ci*Cout + co)*S + k%S)*2 + k/S]
        const int S = sh.stride;
        for (int j = threadIdx.x; j < inner; j += blockDim.x) {
            const int co = j / sh.k, k = j - co * sh.k;
            wp[(((size_t)row * sh.cout + co) * S + (k % S)) * 2 + (k / S)] = scale * vr[j];
        }
        // split-bf16 copy for the tensor-core ConvT kernel (layout: mg_layout.h, up_weight_index)
        const int stage = l - 1;
        __nv_bfloat16 *tcw = reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<char *>(packed) + tc_region_start() +
                                                              tc_up_offset(stage));
        for (int j = threadIdx.x; j < inner; j += blockDim.x) {
            const int co = j / sh.k, k = j - co * sh.k;
            __nv_bfloat16 hi, lo;
            tc::split_bf16(scale * vr[j], hi, lo);
            tcw[up_weight_index(stage, row, co, k, 0)] = hi;
            tcw[up_weight_index(stage, row, co, k, 1)] = lo;
        }
        if (stage >= 2) {  // ... and for the ConvT fused into the ResBlock kernel (layout: mg_layout.h, upf_weight_index)
            __nv_bfloat16 *fw = reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<char *>(packed) + tc_region_start() +
                                                                 tc_upf_offset(stage));
            for (int j = threadIdx.x; j < inner; j += blockDim.x) {
                const int co = j / sh.k, k = j - co * sh.k;
                __nv_bfloat16 hi, lo;
                tc::split_bf16(scale * vr[j], hi, lo);
                fw[upf_weight_index(sh.cout, row, co, k, 0)] = hi;
                fw[upf_weight_index(sh.cout, row, co, k, 1)] = lo;
            }
        }
    }
    if (threadIdx.x == 0 && row < sh.cout) packed[bias_offset(l) + row] = a.bias[l][row];
}

int launch_pack(const float *const *v, const float *const *g, const float *const *bias, float *packed,
                cudaStream_t s) {
    PackArgs a;
    for (int l = 0; l < kNumLayers; ++l) {
        if (!v[l] || !g[l] || !bias[l]) return set_error(MG_ERR_INVALID_ARGUMENT, "mg_gen_pack: null tensor for layer %d", l);
        a.v[l] = v[l];
        a.g[l] = g[l];
        a.bias[l] = bias[l];
    }
    static const RowTable rt = make_row_table();
    pack_kernel<<<rt.first_row[kNumLayers], 128, 0, s>>>(a, rt, packed);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

}  // namespace mg
