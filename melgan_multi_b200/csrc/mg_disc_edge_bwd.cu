// Backward of the two one-channel-sided layers of a Discriminator (models.py:77,85 of the reference; autograd of :90,99):
//   conv_pre   Conv1d(1 -> 16, k15, pad 7)     x [Bt][1][L],    dz [Bt][16][L]
//   conv_post2 Conv1d(1024 -> 1, k3, pad 1)    x [Bt][1024][L], dz [Bt][1][L]
// Both are tiny contractions (240 / 3072 weights) over long activations -- bandwidth-bound reductions, nothing for the tensor
// cores -- that cuDNN runs as general convolutions behind nchw<->nhwc conversion kernels.  fp32 SIMT, fixed summation order
// (per-tile partials combined by a second launch: bit-reproducible), dw in the torch layout of the folded weight.
#include "mg_common.cuh"

namespace mg {

namespace {
constexpr int kPreTile = 512;                     // positions per CTA
constexpr int kPreHalo = 7;
constexpr int kPreRow = kPreTile + 2 * kPreHalo;  // staged positions t0 - 7 .. t0 + 518
constexpr int kPrePitch = kPreRow + 1;            // odd: the 16 channel rows start in different banks
}  // namespace

// CTA = (tile of 512 positions, item).  w: blob fp32 [tap 15][co 16].  partial: [tile][256] = dw [16][15] then db [16].
__global__ void __launch_bounds__(256) disc_pre_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ x,
                                                           const float *__restrict__ w, float *__restrict__ dx,
                                                           float *__restrict__ partial, int L, int tiles_per_item) {
    __shared__ float zs[16 * kPrePitch];
    __shared__ float xs[kPreRow];
    __shared__ float ws[240];
    const int tid = threadIdx.x;
    const int b = blockIdx.x / tiles_per_item, t0 = (blockIdx.x - b * tiles_per_item) * kPreTile;
    if (tid < 240) ws[tid] = w[tid];
    for (int i = tid; i < 16 * kPreRow; i += 256) {
        const int co = i / kPreRow, j = i - co * kPreRow, t = t0 - kPreHalo + j;
        zs[co * kPrePitch + j] = (t >= 0 && t < L) ? dz[((size_t)b * 16 + co) * L + t] : 0.f;
    }
    for (int j = tid; j < kPreRow; j += 256) {
        const int t = t0 - kPreHalo + j;
        xs[j] = (t >= 0 && t < L) ? x[(size_t)b * L + t] : 0.f;
    }
    __syncthreads();
    if (dx) {  // dx[p] = sum_co sum_k dz[co][p - k + 7] w[co][k]
#pragma unroll
        for (int h = 0; h < kPreTile / 256; ++h) {
            const int i = tid + 256 * h;
            if (t0 + i < L) {
                float acc = 0.f;
#pragma unroll 1
                for (int co = 0; co < 16; ++co) {
                    const float *zr = zs + co * kPrePitch + i + 2 * kPreHalo;  // position p - k + 7  <->  zr[-k]
#pragma unroll
                    for (int k = 0; k < 15; ++k) acc = fmaf(zr[-k], ws[k * 16 + co], acc);
                }
                dx[(size_t)b * L + t0 + i] = acc;
            }
        }
    }
    // dw[co][k] = sum_t dz[co][t] x[t + k - 7] over this tile; db[co] = sum_t dz[co][t]
    float acc = 0.f;
    const int n = min(kPreTile, L - t0);
    if (tid < 240) {
        const int co = tid / 15, k = tid - co * 15;
        const float *zr = zs + co * kPrePitch + kPreHalo, *xr = xs + k;
#pragma unroll 4
        for (int i = 0; i < n; ++i) acc = fmaf(zr[i], xr[i], acc);
    } else {
        const float *zr = zs + (tid - 240) * kPrePitch + kPreHalo;
        for (int i = 0; i < n; ++i) acc += zr[i];
    }
    partial[(size_t)blockIdx.x * 256 + tid] = acc;
}

__global__ void __launch_bounds__(256) disc_pre_bwd_combine_kernel(const float *__restrict__ partial, float *__restrict__ dw,
                                                                   float *__restrict__ db, int ntiles) {
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int c = 0; c < ntiles; ++c) s += partial[(size_t)c * 256 + tid];
    if (tid < 240) dw[tid] = s;
    else db[tid - 240] = s;
}

// dx[b][ci][p] = dz[b][p+1] w[ci][0] + dz[b][p] w[ci][1] + dz[b][p-1] w[ci][2].  w: blob fp32 [ci 1024][tap 3]; grid (1024 / 8, Bt)
__global__ void __launch_bounds__(256) disc_post2_dx_kernel(const float *__restrict__ dz, const float *__restrict__ w,
                                                            float *__restrict__ dx, int L) {
    const int b = blockIdx.y, ci = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const float w0 = w[ci * 3], w1 = w[ci * 3 + 1], w2 = w[ci * 3 + 2];
    const float *zr = dz + (size_t)b * L;
    float *out = dx + ((size_t)b * 1024 + ci) * L;
    for (int p = lane; p < L; p += 32) {
        const float zp = p + 1 < L ? __ldg(zr + p + 1) : 0.f, zm = p > 0 ? __ldg(zr + p - 1) : 0.f;
        out[p] = fmaf(zp, w0, fmaf(__ldg(zr + p), w1, zm * w2));
    }
}

// dw[ci][k] = sum_b sum_p x[b][ci][p] dz[b][p - k + 1]; one CTA per input channel, threads along the flattened (item, position)
// axis (independent loads in flight; a warp-per-channel loop would expose one L2 round trip per 32 positions); db = sum dz (CTA 0)
__global__ void __launch_bounds__(256) disc_post2_dw_kernel(const float *__restrict__ dz, const float *__restrict__ x,
                                                            float *__restrict__ dw, float *__restrict__ db, int Bt, int L) {
    const int ci = blockIdx.x, tid = threadIdx.x, n = Bt * L;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, sz = 0.f;
#pragma unroll 4
    for (int e = tid; e < n; e += 256) {
        const int b = e / L, p = e - b * L;
        const float v = __ldg(x + ((size_t)b * 1024 + ci) * L + p);
        const float z = __ldg(dz + e);
        a0 = fmaf(v, p + 1 < L ? __ldg(dz + e + 1) : 0.f, a0);
        a1 = fmaf(v, z, a1);
        a2 = fmaf(v, p > 0 ? __ldg(dz + e - 1) : 0.f, a2);
        sz += z;
    }
    __shared__ float red[4][8];
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = a0; red[1][tid >> 5] = a1; red[2][tid >> 5] = a2; red[3][tid >> 5] = sz; }
    __syncthreads();
    if (tid < 4) {
        float s = 0.f;
#pragma unroll
        for (int wp = 0; wp < 8; ++wp) s += red[tid][wp];
        if (tid < 3) dw[ci * 3 + tid] = s;
        else if (ci == 0) db[0] = s;
    }
}

size_t edge_bwd_workspace_bytes(int l, int Bt, int L) {
    return l == 0 ? (size_t)Bt * ((L + kPreTile - 1) / kPreTile) * 256 * sizeof(float) : 0;
}

// blob: one discriminator's packed weights; l = 0 (conv_pre) or 6 (conv_post2).  dx may be null (not needed); ws: l = 0 only.
int launch_disc_edge_backward(const void *blob, int l, const float *dz, const float *x, float *dx, float *dw, float *db, float *ws,
                              int Bt, int L, cudaStream_t s) {
    const float *w = reinterpret_cast<const float *>(blob) + d_weight_offset(l);
    if (l == 0) {
        const int tiles_per_item = (L + kPreTile - 1) / kPreTile, ntiles = Bt * tiles_per_item;
        disc_pre_bwd_kernel<<<ntiles, 256, 0, s>>>(dz, x, w, dx, ws, L, tiles_per_item);
        MG_CUDA_TRY(cudaGetLastError());
        disc_pre_bwd_combine_kernel<<<1, 256, 0, s>>>(ws, dw, db, ntiles);
        MG_CUDA_TRY(cudaGetLastError());
        return MG_OK;
    }
    if (Bt > 65535) return set_error(MG_ERR_INVALID_ARGUMENT, "discriminator batch %d exceeds 65535", Bt);
    if (dx) {
        disc_post2_dx_kernel<<<dim3(1024 / 8, Bt), 256, 0, s>>>(dz, w, dx, L);
        MG_CUDA_TRY(cudaGetLastError());
    }
    disc_post2_dw_kernel<<<1024, 256, 0, s>>>(dz, x, dw, db, Bt, L);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

}  // namespace mg
