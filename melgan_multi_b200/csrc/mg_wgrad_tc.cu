// Weight gradient of Discriminator.conv_post1 (Conv1d 1024 -> 1024, k5, pad 2; models.py:84,96) on the tensor cores
// (tcgen05 + TMEM, split-bf16: fp32-grade, not TF32), plus its bias gradient.
//
//     dW[co][ci][tap] = sum_b sum_l dz[b][co][l] * x[b][ci][l + tap - 2]        db[co] = sum_b sum_l dz[b][co][l]
//
// GEMM view: M = 128 output channels (TMEM lane), N = 64 input channels per tap, five taps side by side in TMEM (5 x 64 = 320
// fp32 columns), K = the positions of all items (K-major operands: in the NCL layout a channel's positions ARE contiguous).
// One CTA = one (128 co) x (64 ci) tile of all five taps: 8 x 16 = 128 CTAs, each streaming the whole K extent once.
// A tap shifts x along K, i.e. by one bf16 INSIDE a 16-byte operand row -- not expressible as a descriptor offset -- so the
// converter warps write five shifted copies of the (smaller) x tile next to one copy of the dz tile; the zero padding of the
// reference at the two ends of every item is produced there, which is also what lets items follow each other along K.
//   stage (32 positions = 2 K16 steps): A = split(dz) [hi|lo][4 k-panels][128 rows][16 B] = 16 KB
//                                       B = split(x)  [tap][hi|lo][4 k-panels][64 rows][16 B] = 40 KB       x 3 stages
// Positions are padded per item to a multiple of 8 (one k-panel never straddles two items), the K extent to a multiple of 32.
#include "mg_common.cuh"
#include "mg_tc.cuh"

namespace mg {
using namespace tc;

namespace wg {
constexpr int C = 1024, NTAP = 5, PAD = 2;
constexpr int MT = 128, NTILE = 64;            // co rows / ci rows of a CTA
constexpr int NPANEL = 4;                      // k-panels (8 positions) per stage
constexpr int APANEL = MT * 16, BPANEL = NTILE * 16;
constexpr int AHALF = NPANEL * APANEL, BHALF = NPANEL * BPANEL;
constexpr int ASTAGE = 2 * AHALF, BSTAGE = NTAP * 2 * BHALF, STAGE = ASTAGE + BSTAGE;
constexpr int NSTAGE = 3;
constexpr int NCONV = 256;                     // converter threads: 512 A units (row, panel) + 256 B units per stage
constexpr int NT = NCONV + 32;
constexpr int TMEM_COLS = 512;                 // 5 x 64 accumulator columns (power-of-two allocation)
constexpr int SMEM_BYTES = NSTAGE * STAGE + (2 * NSTAGE + 1) * 8 + 16 + NCONV * 4;
static_assert(SMEM_BYTES + 1024 <= 227 * 1024, "shared memory budget");
}  // namespace wg

// 8 consecutive positions [p0, p0 + 8) of one channel row (length L, base `row`), zero outside [0, L)
__device__ __forceinline__ void load8(const float *__restrict__ row, int p0, int L, bool vec, bool live, float (&f)[8]) {
    if (vec) {  // L % 4 == 0 and the row base is 16-byte aligned: each float4 is wholly inside or wholly outside
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int p = p0 + 4 * q;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && p >= 0 && p < L) v = __ldg(reinterpret_cast<const float4 *>(row + p));
            f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (live && p0 + j >= 0 && p0 + j < L) ? __ldg(row + p0 + j) : 0.f;
    }
}

__device__ __forceinline__ void store_split8(uint8_t *hi_dst, uint8_t *lo_dst, const float *f) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2_bf16(f[2 * e], f[2 * e + 1], h[e], l[e]);
    *reinterpret_cast<uint4 *>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(wg::NT, 1)
post1_wgrad_tc_kernel(const float *__restrict__ x, const float *__restrict__ dz, float *__restrict__ dw, float *__restrict__ db,
                      int Bt, int L, int *__restrict__ status) {
    using namespace wg;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + NSTAGE * STAGE);
    uint64_t *empty = full + NSTAGE, *done = empty + NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    float *dbsum = reinterpret_cast<float *>(tmem_slot + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int co0 = blockIdx.x * MT, ci0 = blockIdx.y * NTILE;
    const int ppi = (L + 7) >> 3;                 // k-panels per item
    const int npanels = Bt * ppi;
    const int nstages = (npanels + NPANEL - 1) / NPANEL;

    if (warp == NCONV / 32) tmem_alloc(tmem_slot, TMEM_COLS);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NCONV / 32); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == NCONV / 32) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_bf16(MT, NTILE);
        const uint64_t adesc_t = desc_template(APANEL, 128), bdesc_t = desc_template(BPANEL, 128);
        const uint32_t base = smem_u32(smem);
        int s = 0, ph = 0;
        bool ok = true;
#pragma unroll 1
        for (int st = 0; st < nstages; ++st) {
            ok &= mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t a0 = base + s * STAGE, b0 = a0 + ASTAGE;
#pragma unroll 1
            for (int j = 0; j < NPANEL / 2; ++j) {
#pragma unroll 1
                for (int tap = 0; tap < NTAP; ++tap) {
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint64_t adesc = desc_at(adesc_t, a0 + (pass == 1) * AHALF + 2 * j * APANEL);
                        const uint64_t bdesc = desc_at(bdesc_t, b0 + (2 * tap + (pass == 2)) * BHALF + 2 * j * BPANEL);
                        const bool acc = !(st == 0 && j == 0 && pass == 0);
                        if (elect_one()) mma_bf16(tmem + tap * NTILE, adesc, bdesc, idesc, acc);
                    }
                }
            }
            if (elect_one()) mma_commit(&empty[s]);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        if (elect_one()) mma_commit(done);
        if (!ok && lane == 0) atomicExch(status, 26);
    } else {
        // ================= converter warps =================
        // A units: (row r = tid % 128, panel tid / 128) and the same row two panels further; B unit: (row tid % 64, panel tid / 64).
        // Lanes of a warp are consecutive rows of one panel: conflict-free 16-byte stores, and one (item, chunk) per warp.
        const bool vec = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dz)) & 15) == 0;
        const int ar = tid & (MT - 1), ap = tid >> 7;       // panels ap, ap + 2
        const int br = tid & (NTILE - 1), bp = tid >> 6;
        float acc_db = 0.f;
        int s = 0, ph = 0;
        bool ok = true;
#pragma unroll 1
        for (int st = 0; st < nstages; ++st) {
            // loads first (they do not depend on the ring), then wait for the slot
            float fa[2][8], fb[16];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = st * NPANEL + ap + 2 * u;
                const bool live = q < npanels;
                const int b = live ? q / ppi : 0, c = q - b * ppi;
                load8(dz + ((size_t)b * C + co0 + ar) * L, 8 * c, L, vec, live, fa[u]);
            }
            {
                const int q = st * NPANEL + bp;
                const bool live = q < npanels;
                const int b = live ? q / ppi : 0, c = q - b * ppi;
                const float *row = x + ((size_t)b * C + ci0 + br) * L;
                float lo4[8], hi4[8];
                load8(row, 8 * c - 4, L, vec, live, lo4);   // positions 8c-4 .. 8c+3
                load8(row, 8 * c + 4, L, vec, live, hi4);   // positions 8c+4 .. 8c+11
#pragma unroll
                for (int j = 0; j < 8; ++j) { fb[j] = lo4[j]; fb[8 + j] = hi4[j]; }
            }
            if (ok && !mbar_wait(&empty[s], ph ^ 1)) { ok = false; if (lane == 0) atomicExch(status, 27); }
            uint8_t *a = smem + s * STAGE, *bb = a + ASTAGE;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                uint8_t *dst = a + (ap + 2 * u) * APANEL + ar * 16;
                store_split8(dst, dst + AHALF, fa[u]);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc_db += fa[u][j];
            }
#pragma unroll
            for (int tap = 0; tap < NTAP; ++tap) {
                // X_tap[ci][l] = x[ci][l + tap - PAD]; fb[j] holds position 8c - 4 + j
                uint8_t *dst = bb + (2 * tap) * BHALF + bp * BPANEL + br * 16;
                store_split8(dst, dst + BHALF, &fb[4 + tap - PAD]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (++s == NSTAGE) { s = 0; ph ^= 1; }
        }
        dbsum[tid] = acc_db;
        // ================= epilogue: TMEM [co][tap][ci] -> dW [co][ci][tap] =================
        if (ok && !mbar_wait(done, 0)) { ok = false; if (lane == 0) atomicExch(status, 28); }
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;             // TMEM lane quadrant, half of the ci columns
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        float *out = dw + ((size_t)(co0 + q * 32 + lane) * C + ci0) * NTAP;
#pragma unroll 1
        for (int g = half * (NTILE / 16); g < (half + 1) * (NTILE / 16); ++g) {
            uint32_t w[NTAP][8];
#pragma unroll
            for (int tap = 0; tap < NTAP; ++tap) tmem_ld8(lane_addr + tap * NTILE + g * 8, w[tap]);
            tmem_ld_wait();
            float o[8 * NTAP];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int tap = 0; tap < NTAP; ++tap) o[i * NTAP + tap] = __uint_as_float(w[tap][i]);
            float4 *dst = reinterpret_cast<float4 *>(out + (size_t)g * 8 * NTAP);
#pragma unroll
            for (int v = 0; v < 8 * NTAP / 4; ++v) dst[v] = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (blockIdx.y == 0 && tid < MT) db[co0 + tid] = dbsum[tid] + dbsum[tid + MT];
    if (warp == NCONV / 32) tmem_dealloc(tmem, TMEM_COLS);
}

// x, dz [Bt][1024][L] -> dw [1024][1024][5], db [1024]
int launch_disc_post1_wgrad_tc(const float *x, const float *dz, float *dw, float *db, int Bt, int L, int *status, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        MG_CUDA_TRY(cudaFuncSetAttribute(post1_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wg::SMEM_BYTES));
        configured = true;
    }
    MG_CUDA_TRY(launch_ex(post1_wgrad_tc_kernel, dim3(wg::C / wg::MT, wg::C / wg::NTILE), dim3(wg::NT), wg::SMEM_BYTES, s, 1, false, x, dz,
                          dw, db, Bt, L, status));
    return MG_OK;
}

}  // namespace mg
