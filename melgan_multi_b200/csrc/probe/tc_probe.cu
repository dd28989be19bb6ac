// tc_probe: validates the tcgen05 building blocks of mg_tc.cuh on real hardware before the fused
// kernels depend on them.  Computes one dilated k=3 conv tile
//     D[m][n] = sum_tap sum_ci W[tap][n][ci] * X[16 + m + (tap-1)*dil][ci],   m < 128
// as 3 taps x 3 split-bf16 passes x C/16 UMMA instructions with the "row-linear K-major, no swizzle"
// operand layout, where a tap is nothing but a 16*dil byte shift of the A start address, and compares
// against a double-precision CPU result.  Standalone: nvcc -o tc_probe tc_probe.cu; ./tc_probe
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../mg_tc.cuh"

using namespace mg::tc;

constexpr int ROWS = 160;  // 16 slack + 128 + 16 slack

template <int C, int N>
__global__ void __launch_bounds__(128, 1)
probe_conv(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ out, int dil, int variant,
           int passes, int *status) {
    constexpr int KP = C / 8;                 // k-panels
    constexpr int XPITCH = ROWS * 16;         // bytes between k-panels of X
    constexpr int WPITCH = N * 16;            // bytes between k-panels of one tap of W
    constexpr int TCOLS = N < 32 ? 32 : N;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *Xh = smem, *Xl = Xh + KP * XPITCH;
    uint8_t *Wh = Xl + KP * XPITCH, *Wl = Wh + 3 * KP * WPITCH;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    // operands -> smem (hi/lo bf16, panel layout)
    for (int idx = tid; idx < ROWS * KP; idx += 128) {
        const int r = idx % ROWS, kp = idx / ROWS;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(X[r * C + kp * 8 + 2 * e], X[r * C + kp * 8 + 2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Xh + kp * XPITCH + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Xl + kp * XPITCH + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    for (int idx = tid; idx < 3 * N * KP; idx += 128) {
        const int n = idx % N, kp = (idx / N) % KP, tap = idx / (N * KP);
        const float *w = W + ((size_t)tap * N + n) * C + kp * 8;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(w[2 * e], w[2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Wh + (tap * KP + kp) * WPITCH + n * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Wl + (tap * KP + kp) * WPITCH + n * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        bool acc = false;
        for (int tap = 0; tap < 3; ++tap)
            for (int pass = 0; pass < passes; ++pass) {
                const uint8_t *xa = (pass == 1) ? Xl : Xh;  // passes: (xh,wh) (xl,wh) (xh,wl)
                const uint8_t *wb = (pass == 2) ? Wl : Wh;
                for (int k = 0; k < C / 16; ++k) {
                    const uint32_t a_addr = smem_u32(xa) + (2 * k) * XPITCH + (16 + (tap - 1) * dil) * 16;
                    const uint32_t b_addr = smem_u32(wb) + (tap * KP + 2 * k) * WPITCH;
                    uint64_t ad, bd;
                    if (variant == 0) {  // LBO = k-panel pitch, SBO = 8-row group pitch (128 B)
                        ad = make_desc(a_addr, XPITCH, 128);
                        bd = make_desc(b_addr, WPITCH, 128);
                    } else {             // swapped roles
                        ad = make_desc(a_addr, 128, XPITCH);
                        bd = make_desc(b_addr, 128, WPITCH);
                    }
                    mma_bf16(tmem, ad, bd, idesc, acc);
                    acc = true;
                }
            }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0, 1u << 22);
    tc_fence_after();
    if (!ok) {
        if (tid == 0) atomicExch(status, 1);
    } else {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + c0, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) out[(32 * warp + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int C, int N>
static int run_case(int dil, int variant, int passes) {
    std::vector<float> X((size_t)ROWS * C), W((size_t)3 * N * C), out((size_t)128 * N, 0.f);
    srand(1234 + C + N);
    for (auto &v : X) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto &v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) / sqrtf((float)(3 * C));
    float *dX, *dW, *dO;
    int *dS, st = 0;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, out.size() * 4); cudaMemset(dS, 0, 4);
    const size_t smem = 2 * (C / 8) * ROWS * 16 + 2 * 3 * (C / 8) * N * 16;
    cudaFuncSetAttribute(probe_conv<C, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_conv<C, N><<<1, 128, smem>>>(dX, dW, dO, dil, variant, passes, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("C=%d N=%d dil=%d variant=%d passes=%d: CUDA ERROR %s\n", C, N, dil, variant, passes, cudaGetErrorString(e));
        return 2;
    }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int tap = 0; tap < 3; ++tap)
                for (int c = 0; c < C; ++c) s += (double)W[((size_t)tap * N + n) * C + c] * X[(16 + m + (tap - 1) * dil) * C + c];
            maxerr = fmax(maxerr, fabs(s - out[m * N + n]));
            maxref = fmax(maxref, fabs(s));
        }
    printf("C=%3d N=%3d dil=%d variant=%d passes=%d: timeout=%d max|err|/max|ref| = %.3e %s\n", C, N, dil, variant, passes, st,
           maxerr / maxref, (st == 0 && maxerr / maxref < (passes == 3 ? 2e-5 : 2e-2)) ? "OK" : "MISMATCH");
    cudaFree(dX); cudaFree(dW); cudaFree(dO); cudaFree(dS);
    return (st == 0 && maxerr / maxref < (passes == 3 ? 2e-5 : 2e-2)) ? 0 : 1;
}

// Toeplitz operand: the k-panels of A OVERLAP (panel j of row m is the 16-byte unit m + j of one linear buffer, i.e. the
// descriptor's LBO is 16 bytes, equal to the row pitch).  This is what a strided / grouped conv with a channels-last
// input looks like (row m = a sliding window), and it needs no im2col if the hardware accepts it.
//     D[m][n] = sum_{j<6} sum_{e<8} U[m + j][e] * W[n][j*8 + e],   M = 128, N = 16, K = 48 (3 MMAs per pass)
__global__ void __launch_bounds__(128, 1)
probe_toeplitz(const float *__restrict__ U, const float *__restrict__ W, float *__restrict__ out, int *status) {
    constexpr int NU = 144, N = 16, KP = 6;
    __shared__ __align__(128) uint8_t Uh[NU * 16], Ul[NU * 16], Wh[KP * N * 16], Wl[KP * N * 16];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 32);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    for (int u = tid; u < NU; u += 128) {
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(U[u * 8 + 2 * e], U[u * 8 + 2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Uh + u * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Ul + u * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    for (int idx = tid; idx < KP * N; idx += 128) {
        const int n = idx % N, kp = idx / N;
        uint32_t h[4], l[4];
        for (int e = 0; e < 4; ++e) split2_bf16(W[n * 48 + kp * 8 + 2 * e], W[n * 48 + kp * 8 + 2 * e + 1], h[e], l[e]);
        *reinterpret_cast<uint4 *>(Wh + (kp * N + n) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(Wl + (kp * N + n) * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        bool acc = false;
        for (int pass = 0; pass < 3; ++pass)
            for (int k = 0; k < 3; ++k) {
                const uint32_t a_addr = smem_u32(pass == 1 ? Ul : Uh) + (2 * k) * 16;           // panel 2k of row 0 = unit 2k
                const uint32_t b_addr = smem_u32(pass == 2 ? Wl : Wh) + (2 * k) * (N * 16);
                mma_bf16(tmem, make_desc(a_addr, /*LBO: next k-panel = next unit*/ 16, 128), make_desc(b_addr, N * 16, 128), idesc, acc);
                acc = true;
            }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0, 1u << 22);
    tc_fence_after();
    if (!ok) {
        if (tid == 0) atomicExch(status, 1);
    } else {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[(32 * warp + lane) * N + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 32);
}

static int run_toeplitz() {
    std::vector<float> U(144 * 8), W(16 * 48), out(128 * 16, 0.f);
    srand(77);
    for (auto &v : U) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    for (auto &v : W) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) / 7.f;
    float *dU, *dW, *dO;
    int *dS, st = 0;
    cudaMalloc(&dU, U.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dU, U.data(), U.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    probe_toeplitz<<<1, 128>>>(dU, dW, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("toeplitz: CUDA ERROR %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            double s = 0;
            for (int j = 0; j < 6; ++j)
                for (int el = 0; el < 8; ++el) s += (double)U[(m + j) * 8 + el] * W[n * 48 + j * 8 + el];
            maxerr = fmax(maxerr, fabs(s - out[m * 16 + n]));
            maxref = fmax(maxref, fabs(s));
        }
    const bool good = st == 0 && maxerr / maxref < 2e-5;
    printf("toeplitz (LBO = 16 B, overlapping k-panels) N=16 K=48: timeout=%d max|err|/max|ref| = %.3e %s\n", st, maxerr / maxref,
           good ? "OK" : "MISMATCH");
    return good ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run_toeplitz();
    // which LBO/SBO assignment is right (variant 0 expected), single pass first
    for (int variant = 0; variant < 2; ++variant) bad += run_case<64, 64>(1, variant, 1) && variant == 0;
    for (int dil : {1, 3, 9}) {
        bad += run_case<64, 64>(dil, 0, 3);
        bad += run_case<32, 32>(dil, 0, 3);
        bad += run_case<128, 64>(dil, 0, 3);
        bad += run_case<64, 128>(dil, 0, 3);
    }
    bad += run_case<32, 16>(3, 0, 3);
    bad += run_case<16, 256>(3, 0, 3);
    printf(bad ? "PROBE FAILED (%d)\n" : "PROBE PASSED\n", bad);
    return bad ? 1 : 0;
}
