// LeakyReLU -> ConvTranspose1d on the tensor cores (tcgen05 + TMEM, split-bf16).
//
// Reference: Generator.forward, models.py:64-65 -- x = ups[i](F.leaky_relu(x)); ConvTranspose1d(Cin, Cout, K=2S, stride S,
// padding S/2), models.py:48-51.
//
// A transposed conv with K = 2S is, per output phase phi = (t + pad) mod S, a 2-tap conv over the INPUT positions:
//     out[co][S*s + phi - pad] = sum_ci  x[ci][s] * W[ci][co][phi]  +  x[ci][s-1] * W[ci][co][phi + S]
// GEMM view per phase: D_phi[s, co] = X[s, :] * W_phi0[co, :]^T + X[s-1, :] * W_phi1[co, :]^T with M = 128 input positions
// (TMEM lane = s), N = NG output channels, K = 16 per instruction.  The s-1 tap is the same A buffer read one row
// earlier (row-linear operand layout, mg_tc.cuh).  All S phases of a (row block, channel group) are resident in TMEM
// at once (S * NG columns per block), so the epilogue thread of input position s owns the S consecutive output
// samples [S*s - pad, S*s - pad + S) of each channel and stores them as contiguous, fully coalesced vectors.
//
// Rows are VIRTUAL input positions: the B batch items are concatenated with one zero row after each item
// (v = item*(Lin+1) + s, s in [0, Lin], row s = Lin is zero), so that x[-1] = x[Lin] = 0 falls out of the layout and short
// sequences (stage 0: Lin = 32) still fill 128-row blocks.
// One CTA = NB blocks of 128 virtual input positions x one group of NG output channels.  K (= Cin) is streamed in 16-channel
// chunks: the A chunk (LeakyReLU + hi/lo split of x) is produced in shared memory by the converter warps straight
// from the fp32 NCL input, the B slots (per chunk and phase: both taps, hi and lo) arrive by 1-D bulk TMA from the
// pre-packed blob (mg_layout.h).  Warp roles: converter/epilogue warps, TMA producer, MMA issuer.
#include "mg_common.cuh"
#include "mg_tc.cuh"

namespace mg {
using namespace tc;

template <int STAGE_>
struct UpCfg {
    static constexpr int STAGE = STAGE_;
    static constexpr int CIN = stage_cin(STAGE), COUT = stage_cout(STAGE), S = stage_stride(STAGE), PAD = stage_pad(STAGE);
    static constexpr int NG = up_ng(STAGE);
    static constexpr int NCG = COUT / NG;
    static constexpr int NB = (S == 8) ? 1 : 2;          // 128-row blocks per CTA
    static constexpr int COLS = NB * S * NG;              // TMEM columns in use
    static constexpr int TCOLS = COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    static constexpr int MINB = 2;                        // CTAs per SM: one CTA's loads / stores hide under the other's MMAs
    static constexpr int ROWS = 128 * NB;
    static constexpr int AROWS = ROWS + 8;                // row index i <-> input position r0 - 1 + i, i in [0, ROWS]
    static constexpr int APITCH = AROWS * 16;             // bytes between the two k-panels of a chunk
    static constexpr int ASLOT = 4 * APITCH;              // [half: hi, lo][k-panel: 2][AROWS][16 B]
    static constexpr int BSLOT = up_slot_bytes(STAGE);    // [tap][half][k-panel][NG][16 B]
    static constexpr int NSA = 3, NSB = (S == 8) ? 8 : 4;
    static constexpr int NCHUNK = CIN / 16;
    static constexpr int NWG = NB >= 2 ? 2 : 1;
    static constexpr int NCONV = 128 * NWG;               // converter / epilogue threads
    static constexpr bool BY_PHASE = (NB == 1);           // issuers split the S phases (NB == 1) or the NB row blocks
    static constexpr int NIW = BY_PHASE ? 4 : NB;         // MMA issuer warps
    static constexpr int NT = NCONV + 32 + 32 * NIW;
    static_assert(BY_PHASE ? (S % NIW == 0) : (NB % NIW == 0), "issuer split");
    static constexpr int SMEM_BYTES = NSA * ASLOT + NSB * BSLOT + (2 * NSA + 2 * NSB + 1) * 8 + 16;
    static_assert(MINB * TCOLS <= 512, "TMEM columns");
    static_assert(MINB * (SMEM_BYTES + 1024) <= 228 * 1024, "shared memory budget");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB)
convt_tc_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ packed, int Lin, int B,
                int *__restrict__ status) {
    constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, S = Cfg::S, PAD = Cfg::PAD, NG = Cfg::NG, NB = Cfg::NB;
    constexpr int ROWS = Cfg::ROWS, APITCH = Cfg::APITCH, ASLOT = Cfg::ASLOT, BSLOT = Cfg::BSLOT;
    constexpr int NSA = Cfg::NSA, NSB = Cfg::NSB, NCHUNK = Cfg::NCHUNK, NCONV = Cfg::NCONV, NWG = Cfg::NWG, NIW = Cfg::NIW;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *aring = smem, *bring = smem + NSA * ASLOT;
    uint64_t *fullA = reinterpret_cast<uint64_t *>(bring + NSB * BSLOT);
    uint64_t *emptyA = fullA + NSA, *fullB = emptyA + NSA, *emptyB = fullB + NSB, *done = emptyB + NSB;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * ROWS;  // first virtual row of the tile
    const int cg = blockIdx.y;
    const int Lout = Lin * S, Lv = Lin + 1;

    if (warp == 0) tmem_alloc(tmem_slot, Cfg::TCOLS);
    if (tid == 32) {
        for (int s = 0; s < NSA; ++s) { mbar_init(&fullA[s], NCONV); mbar_init(&emptyA[s], Cfg::NIW); }
        for (int s = 0; s < NSB; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], Cfg::BY_PHASE ? 1 : Cfg::NIW); }
        mbar_init(done, Cfg::NIW);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == NCONV / 32) {
        // ================= TMA producer: B slots (chunk, phase) =================
        if (lane == 0) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(packed) + tc_region_start() + tc_up_offset(Cfg::STAGE) +
                                 (size_t)cg * NCHUNK * S * BSLOT;
            int s = 0, ph = 0;
            bool ok = true;
            for (int i = 0; i < NCHUNK * S && ok; ++i) {
                if (!mbar_wait(&emptyB[s], ph ^ 1)) { ok = false; break; }
                mbar_arrive_expect_tx(&fullB[s], BSLOT);
                bulk_g2s(bring + s * BSLOT, src + (size_t)i * BSLOT, BSLOT, &fullB[s]);
                if (++s == NSB) { s = 0; ph ^= 1; }
            }
            if (!ok) atomicExch(status, 12);
        }
    } else if (warp > NCONV / 32) {
        // ================= MMA issuers (NIW warps; each runs the loop warp-uniform, one elected lane issues) ==========
        // A single issuing thread sustains about one tcgen05.mma per 50 cycles, slower than these N <= 64 MMAs execute,
        // so the accumulators are split across issuers: by 128-row block when NB > 1, else by output phase.
        const int iw = warp - (NCONV / 32 + 1);
        const uint32_t idesc = make_idesc_bf16(128, NG);
        const uint64_t adesc_t = desc_template(APITCH, 128), bdesc_t = desc_template(NG * 16, 128);
        const uint32_t aring_addr = smem_u32(aring), bring_addr = smem_u32(bring);
        int sa = 0, pha = 0;
        bool ok = true;  // a timed-out wait only raises the status word: control flow stays uniform
#pragma unroll 1
        for (int ch = 0; ch < NCHUNK; ++ch) {
            ok &= mbar_wait(&fullA[sa], pha);
            tc_fence_after();
            const uint64_t abase = desc_at(adesc_t, aring_addr + sa * ASLOT);
#pragma unroll 1
            for (int phi = (Cfg::BY_PHASE ? iw : 0); phi < S; phi += (Cfg::BY_PHASE ? NIW : 1)) {
                const int n = ch * S + phi, sb = n % NSB, phb = (n / NSB) & 1;
                ok &= mbar_wait(&fullB[sb], phb);
                tc_fence_after();
                const uint64_t bbase = desc_at(bdesc_t, bring_addr + sb * BSLOT);
#pragma unroll
                for (int tap = 0; tap < 2; ++tap)
#pragma unroll
                    for (int pass = 0; pass < 3; ++pass) {
                        const int ahalf = (pass == 1), bhalf = (pass == 2);
                        const uint64_t bdesc = bbase + (uint64_t)((((tap * 2 + bhalf) * 2) * NG * 16) >> 4);
#pragma unroll
                        for (int bi = 0; bi < (Cfg::BY_PHASE ? NB : NB / NIW); ++bi) {
                            const int blk = Cfg::BY_PHASE ? bi : iw + bi * NIW;
                            const uint64_t adesc = abase + (uint64_t)((ahalf * 2 * APITCH + (1 - tap) * 16) >> 4) + (uint64_t)(blk * 128);
                            const bool acc = !(ch == 0 && tap == 0 && pass == 0);
                            if (elect_one()) mma_bf16(tmem + (blk * S + phi) * NG, adesc, bdesc, idesc, acc);
                        }
                    }
                if (elect_one()) mma_commit(&emptyB[sb]);
            }
            if (elect_one()) mma_commit(&emptyA[sa]);
            if (++sa == NSA) { sa = 0; pha ^= 1; }
        }
        if (elect_one()) mma_commit(done);
        if (!ok && lane == 0) atomicExch(status, 13);
    } else {
        // ================= converter warps: A chunks = split(lrelu(x)) =================
        int sa = 0, pha = 0;
        bool ok = true;
        for (int ch = 0; ch < NCHUNK; ++ch) {
            if (ok && !mbar_wait(&emptyA[sa], pha ^ 1)) { ok = false; if (lane == 0) atomicExch(status, 14); }
            uint8_t *slot = aring + sa * ASLOT;
            for (int i = tid; i <= ROWS; i += NCONV) {
                const int v = r0 - 1 + i;
                const int item = v >= 0 ? v / Lv : 0, s = v - item * Lv;
                const bool inr = (v >= 0 && item < B && s < Lin);
                const float *xp = x + ((size_t)(inr ? item : 0) * CIN + ch * 16) * Lin + (inr ? s : 0);
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = inr ? lrelu(__ldg(xp + (size_t)j * Lin)) : 0.f;
                uint32_t h[8], l[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split2_bf16(f[2 * e], f[2 * e + 1], h[e], l[e]);
                *reinterpret_cast<uint4 *>(slot + i * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(slot + APITCH + i * 16) = make_uint4(h[4], h[5], h[6], h[7]);
                *reinterpret_cast<uint4 *>(slot + 2 * APITCH + i * 16) = make_uint4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<uint4 *>(slot + 3 * APITCH + i * 16) = make_uint4(l[4], l[5], l[6], l[7]);
            }
            fence_proxy_async();
            mbar_arrive(&fullA[sa]);
            if (++sa == NSA) { sa = 0; pha ^= 1; }
        }
        // ================= epilogue: D_phi[s, co] + bias -> out[co][S*s + phi - pad] =================
        if (ok && !mbar_wait(done, 0)) { ok = false; if (lane == 0) atomicExch(status, 15); }
        tc_fence_after();
        const int wg = warp >> 2, q = warp & 3;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        const float *bias = packed + bias_offset(1 + Cfg::STAGE) + cg * NG;
        for (int blk = wg; blk < NB; blk += NWG) {
            const int v = r0 + blk * 128 + q * 32 + lane;
            const int item = v / Lv, s = v - item * Lv;
            const bool row_ok = item < B;
            const int t0 = S * s - PAD;  // first output sample owned by this input position
            float *yb = y + ((size_t)(row_ok ? item : 0) * COUT + cg * NG) * Lout;
            if (S == 8) {
                const bool lo_ok = row_ok && s >= 1, hi_ok = row_ok && s <= Lin - 1;
#pragma unroll 1
                for (int j0 = 0; j0 < NG; j0 += 8) {
                    uint32_t v[8][8];
#pragma unroll
                    for (int phi = 0; phi < 8; ++phi) tmem_ld8(lane_addr + (blk * S + phi) * NG + j0, v[phi]);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float bj = __ldg(bias + j0 + j);
                        float *yp = yb + (size_t)(j0 + j) * Lout + t0;
                        if (lo_ok)
                            *reinterpret_cast<float4 *>(yp) = make_float4(__uint_as_float(v[0][j]) + bj, __uint_as_float(v[1][j]) + bj,
                                                                          __uint_as_float(v[2][j]) + bj, __uint_as_float(v[3][j]) + bj);
                        if (hi_ok)
                            *reinterpret_cast<float4 *>(yp + 4) = make_float4(__uint_as_float(v[4][j]) + bj, __uint_as_float(v[5][j]) + bj,
                                                                              __uint_as_float(v[6][j]) + bj, __uint_as_float(v[7][j]) + bj);
                    }
                }
            } else {  // S == 2: t0 = 2s - 1
                const bool lo_ok = row_ok && s >= 1, hi_ok = row_ok && s <= Lin - 1;
#pragma unroll 1
                for (int j0 = 0; j0 < NG; j0 += 16) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16(lane_addr + (blk * S + 0) * NG + j0, v0);
                    tmem_ld16(lane_addr + (blk * S + 1) * NG + j0, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float bj = __ldg(bias + j0 + j);
                        float *yp = yb + (size_t)(j0 + j) * Lout + t0;
                        if (lo_ok) yp[0] = __uint_as_float(v0[j]) + bj;
                        if (hi_ok) yp[1] = __uint_as_float(v1[j]) + bj;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, Cfg::TCOLS);
}

template <class Cfg>
static int launch_convt(const float *x, float *y, const float *packed, int B, int Lin, int *status, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        MG_CUDA_TRY(cudaFuncSetAttribute(convt_tc_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const long long vrows = (long long)B * (Lin + 1);  // Lin + 1 rows per item: position s = Lin feeds the last `pad` outputs
    dim3 grid((unsigned)((vrows + Cfg::ROWS - 1) / Cfg::ROWS), Cfg::NCG);
    convt_tc_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM_BYTES, s>>>(x, y, packed, Lin, B, status);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

// x [B][Cin][Lin] -> y [B][Cout][S*Lin], fp32 NCL, (Cin, Cout, S) of generator stage `stage`.
int launch_convt_tc(const float *x, float *y, const float *packed, int stage, int B, int Lin, int *status, cudaStream_t s) {
    switch (stage) {
        case 0: return launch_convt<UpCfg<0>>(x, y, packed, B, Lin, status, s);
        case 1: return launch_convt<UpCfg<1>>(x, y, packed, B, Lin, status, s);
        case 2: return launch_convt<UpCfg<2>>(x, y, packed, B, Lin, status, s);
        case 3: return launch_convt<UpCfg<3>>(x, y, packed, B, Lin, status, s);
    }
    return set_error(MG_ERR_INVALID_ARGUMENT, "launch_convt_tc: stage %d", stage);
}

}  // namespace mg
