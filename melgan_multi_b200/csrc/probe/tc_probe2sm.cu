// tc_probe2sm: validates tcgen05.mma.cta_group::2 (a CTA PAIR executing one M = 256 MMA) with the "row-linear K-major,
// no swizzle" operand layout of mg_tc.cuh before mg_res_tc.cu depends on it.  One dilated k=3 conv tile per CTA,
//     D_r[m][n] = sum_tap sum_ci W[tap][n][ci] * X_r[16 + m + (tap-1)*dil][ci],   m < 128, r = CTA rank (own X_r per CTA),
// issued by the LEADER CTA only.  Hypothesis under test (variant 0): each CTA's shared memory holds HALF of the B operand --
// rows [r*N/2, (r+1)*N/2) of W -- at the descriptor's address (same offsets in both CTAs), its own 128 rows of A, and its
// own 128 x N accumulator at the same TMEM address.  variant 1: every CTA holds ALL N rows of B (descriptor pitch N*16).
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe2sm tc_probe2sm.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mg_tc.cuh"

using namespace mg::tc;

constexpr int ROWS = 160;  // 16 slack + 128 + 16 slack

__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void mma2_commit(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

template <int C, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ out, int dil, int variant, int *status) {
    constexpr int KP = C / 8;
    constexpr int XPITCH = ROWS * 16;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t rank = cluster_ctarank();
    const int NB = variant == 0 ? N / 2 : N;  // rows of B in this CTA's shared memory
    const int WPITCH = NB * 16;
    uint8_t *Xh = smem, *Xl = Xh + KP * XPITCH;
    uint8_t *Wh = Xl + KP * XPITCH, *Wl = Wh + 3 * KP * WPITCH;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc2(&tmem_base_s, N < 32 ? 32 : N);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    const float *Xr = X + (size_t)rank * ROWS * C;
    for (int idx = tid; idx < ROWS * KP; idx += 128) {
        const int r = idx % ROWS, kp = idx / ROWS;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(Xr[r * C + kp * 8 + 2 * e], Xr[r * C + kp * 8 + 2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Xh + kp * XPITCH + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Xl + kp * XPITCH + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    for (int idx = tid; idx < 3 * NB * KP; idx += 128) {
        const int nl = idx % NB, kp = (idx / NB) % KP, tap = idx / (NB * KP);
        const int n = variant == 0 ? (int)rank * NB + nl : nl;
        const float *w = W + ((size_t)tap * N + n) * C + kp * 8;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(w[2 * e], w[2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Wh + (tap * KP + kp) * WPITCH + nl * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Wl + (tap * KP + kp) * WPITCH + nl * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync();  // both CTAs' operands are in place, both barriers exist
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (rank == 0 && warp == 1) {
        const uint32_t idesc = make_idesc_bf16(256, N);
        bool acc = false;
        for (int tap = 0; tap < 3; ++tap)
            for (int pass = 0; pass < 3; ++pass) {
                const uint8_t *xa = (pass == 1) ? Xl : Xh;
                const uint8_t *wb = (pass == 2) ? Wl : Wh;
                for (int k = 0; k < C / 16; ++k) {
                    const uint32_t a_addr = smem_u32(xa) + (2 * k) * XPITCH + (16 + (tap - 1) * dil) * 16;
                    const uint32_t b_addr = smem_u32(wb) + (tap * KP + 2 * k) * WPITCH;
                    if (elect_one()) mma2_bf16(tmem, make_desc(a_addr, XPITCH, 128), make_desc(b_addr, WPITCH, 128), idesc, acc);
                    acc = true;
                }
            }
        if (elect_one()) mma2_commit(&bar, 3);
    }
    const bool ok = mbar_wait(&bar, 0, 1u << 22);
    tc_fence_after();
    if (!ok) {
        if (tid == 0) atomicExch(status, 1 + (int)rank);
    } else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out[((size_t)rank * 128 + 32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tmem, N < 32 ? 32 : N);
}

template <int C, int N>
static int run_case(int dil, int variant) {
    std::vector<float> X((size_t)2 * ROWS * C), W((size_t)3 * N * C), out((size_t)256 * N, 0.f);
    srand(4321 + C + N);
    for (auto &v : X) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto &v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) / sqrtf((float)(3 * C));
    float *dX, *dW, *dO;
    int *dS, st = 0;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, out.size() * 4); cudaMemset(dS, 0, 4);
    const size_t smem = 2 * (C / 8) * ROWS * 16 + 2 * 3 * (C / 8) * (variant == 0 ? N / 2 : N) * 16;
    if (smem > 227 * 1024) { printf("2sm C=%d N=%d variant=%d: skipped (%zu bytes of shared memory)\n", C, N, variant, smem); return 0; }
    cudaFuncSetAttribute(probe2<C, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe2<C, N><<<2, 128, smem>>>(dX, dW, dO, dil, variant, dS);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("2sm C=%d N=%d dil=%d variant=%d: CUDA ERROR %s\n", C, N, dil, variant, cudaGetErrorString(e));
        return 2;
    }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr[2] = {0, 0}, maxref = 0;
    for (int r = 0; r < 2; ++r)
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int tap = 0; tap < 3; ++tap)
                    for (int c = 0; c < C; ++c)
                        s += (double)W[((size_t)tap * N + n) * C + c] * X[((size_t)r * ROWS + 16 + m + (tap - 1) * dil) * C + c];
                maxerr[r] = fmax(maxerr[r], fabs(s - out[((size_t)r * 128 + m) * N + n]));
                maxref = fmax(maxref, fabs(s));
            }
    const bool good = st == 0 && maxerr[0] / maxref < 2e-5 && maxerr[1] / maxref < 2e-5;
    printf("2sm C=%3d N=%3d dil=%d variant=%d: timeout=%d err(cta0)=%.3e err(cta1)=%.3e %s\n", C, N, dil, variant, st, maxerr[0] / maxref,
           maxerr[1] / maxref, good ? "OK" : "MISMATCH");
    cudaFree(dX); cudaFree(dW); cudaFree(dO); cudaFree(dS);
    return good ? 0 : 1;
}

int main() {
    int bad = 0;
    for (int variant = 0; variant < 2; ++variant) {
        int b = 0;
        b += run_case<64, 64>(1, variant);
        b += run_case<32, 32>(3, variant);
        b += run_case<128, 128>(9, variant);
        b += run_case<64, 256>(3, variant);
        printf("variant %d: %s\n", variant, b ? "FAILED" : "PASSED");
        if (variant == 0) bad = b;
    }
    printf(bad ? "PROBE2SM FAILED (%d)\n" : "PROBE2SM PASSED\n", bad);
    return bad ? 1 : 0;
}
