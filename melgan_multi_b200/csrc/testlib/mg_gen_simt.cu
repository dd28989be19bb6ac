// Generator forward, fp32 SIMT path: conv_pre + four fused (LeakyReLU -> ConvTranspose1d -> ResBlock)
// stage kernels, the last one also fusing LeakyReLU -> conv_post -> tanh.
//
// Reference semantics: models.py:61-71 (Generator.forward) and :32-40 (ResBlock.forward).
//
// One CTA owns a tile of PTOT consecutive output positions of one batch item and keeps the whole
// stage on chip:
//   In  [CIN ][PIN ]  lrelu(input tile)                         (smem, aliases U)
//   R   [COUT][PTOT]  residual stream (ConvT output, then x + c2(...))   (smem, fp32)
//   U   [COUT][PTOT]  lrelu(c1(lrelu(R)))                                (smem, fp32)
// Weights never fit (4.7 MB for the 256-channel ResBlock), so each conv streams its folded weights
// from L2 through a double-buffered cp.async ring in chunks of input channels.  The tile carries a
// 16-position halo per side (1+1+3+1+9+1, the ResBlock receptive field; +3 when conv_post is fused),
// recomputed by neighbouring tiles, so HBM sees each stage input once and each stage output once.
// Positions outside [0, L) are forced to zero after EVERY conv, which is what per-layer zero padding
// means for a fused chain (long-context row).
//
// Thread map (all phases): warp (wc, wp) owns output channels [wc*COB, (wc+1)*COB) and positions
// wp*32*PB + lane + 32*j; lanes run along positions so activation reads are conflict-free and weight
// reads are warp-uniform broadcasts (float4 of 4 output channels).
#include "../mg_common.cuh"

namespace mg {

template <int CIN_, int COUT_, int S_, int PTOT_, int WC_, int WP_, int WBUF_, bool POST_, bool RES_ = true>
struct StageCfg {
    static constexpr int CIN = CIN_, COUT = COUT_, S = S_, PTOT = PTOT_, WC = WC_, WP = WP_, WBUF = WBUF_;
    static constexpr bool POST = POST_, RES = RES_;  // RES=false: ConvTranspose1d only (feeds the tensor-core ResBlock)
    static constexpr int PAD = S / 2;                   // ConvTranspose1d padding (4 for k16/s8, 1 for k4/s2)
    static constexpr int HALO = RES ? 16 + (POST ? 3 : 0) : 0;  // ResBlock receptive field (+ conv_post k7)
    static constexpr int PVALID = PTOT - 2 * HALO;
    static constexpr int NT = 32 * WC * WP;
    static constexpr int COB = COUT / WC;               // output channels per thread
    static constexpr int PB = PTOT / (32 * WP);         // positions per thread
    static constexpr int PIN = PTOT / S + 2;            // input positions the tile touches
    static constexpr int PADF = 16;                     // slack floats either side of R and U (dilation-9 taps)
    static constexpr int R_FLOATS = COUT * PTOT + 2 * PADF;
    static constexpr int IN_FLOATS = CIN * PIN;
    static constexpr int U_FLOATS = (RES && R_FLOATS > IN_FLOATS) ? R_FLOATS : (IN_FLOATS + 3) / 4 * 4;
    static constexpr int CIC_UP = WBUF / (COUT * 2 * S);  // input channels per ConvT weight chunk
    static constexpr int cic_res() {
        int c = 1;
        while (2 * c * 3 * COUT <= WBUF && 2 * c <= COUT) c *= 2;
        return c;
    }
    static constexpr int CIC_RES = cic_res();           // input channels per ResBlock weight chunk
    static constexpr size_t SMEM_BYTES = (size_t)(R_FLOATS + U_FLOATS + 2 * WBUF) * sizeof(float);
    static_assert(COUT % WC == 0 && COB % 4 == 0, "COB must be a multiple of 4 (float4 weight loads)");
    static_assert(PTOT % (32 * WP) == 0 && PTOT % S == 0 && 32 % S == 0, "tile shape");
    static_assert(CIC_UP >= 1 && CIN % CIC_UP == 0, "ConvT chunking");
    static_assert(RES || !POST, "conv_post fusion needs the ResBlock");
    static_assert(R_FLOATS % 4 == 0 && U_FLOATS % 4 == 0 && WBUF % 4 == 0, "16-byte alignment of smem regions");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(NT <= 1024, "block size");
};

template <int NT>
__device__ __forceinline__ void load_chunk(float *dst, const float *__restrict__ src, int nfloats, int tid) {
    for (int i = tid * 4; i < nfloats; i += NT * 4) cp_async16(dst + i, src + i);
}

// One k=3 conv of the ResBlock over the whole tile.
//   FIRST: dst U = lrelu(conv_DIL(lrelu(R)) + b)          (models.py:35-37)
//  !FIRST: dst R = conv_1(U) + b + R                       (models.py:38-39)
template <class Cfg, int DIL, bool FIRST>
__device__ __forceinline__ void conv3_phase(float *Rb, float *Ub, float *Wb, const float *__restrict__ wg,
                                            const float *__restrict__ bg, int tid, int co0, int pbase, int o,
                                            int Lout) {
    constexpr int C = Cfg::COUT, PTOT = Cfg::PTOT, COB = Cfg::COB, PB = Cfg::PB, CIC = Cfg::CIC_RES;
    constexpr int CH = CIC * 3 * C;  // floats per weight chunk
    constexpr int NCH = C / CIC;
    const float *src = FIRST ? Rb : Ub;

    float acc[COB][PB];
#pragma unroll
    for (int i = 0; i < COB; ++i)
#pragma unroll
        for (int j = 0; j < PB; ++j) acc[i][j] = 0.f;

    load_chunk<Cfg::NT>(Wb, wg, CH, tid);
    cp_async_commit();
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        if (c + 1 < NCH) load_chunk<Cfg::NT>(Wb + ((c + 1) & 1) * Cfg::WBUF, wg + (size_t)(c + 1) * CH, CH, tid);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float *Wc = Wb + (c & 1) * Cfg::WBUF;
#pragma unroll 2
        for (int cl = 0; cl < CIC; ++cl) {
            const float *row = src + (c * CIC + cl) * PTOT + pbase;
            float xs[3][PB];
#pragma unroll
            for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                for (int j = 0; j < PB; ++j) {
                    const float v = row[32 * j + (tap - 1) * DIL];
                    xs[tap][j] = FIRST ? lrelu(v) : v;
                }
            const float *wrow = Wc + cl * 3 * C + co0;
#pragma unroll
            for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                for (int q = 0; q < COB / 4; ++q) {
                    const float4 w = *reinterpret_cast<const float4 *>(wrow + tap * C + 4 * q);
#pragma unroll
                    for (int j = 0; j < PB; ++j) {
                        acc[4 * q + 0][j] = fmaf(w.x, xs[tap][j], acc[4 * q + 0][j]);
                        acc[4 * q + 1][j] = fmaf(w.y, xs[tap][j], acc[4 * q + 1][j]);
                        acc[4 * q + 2][j] = fmaf(w.z, xs[tap][j], acc[4 * q + 2][j]);
                        acc[4 * q + 3][j] = fmaf(w.w, xs[tap][j], acc[4 * q + 3][j]);
                    }
                }
        }
        __syncthreads();
    }
    // epilogue (every thread touches only its own (co, p) elements of dst)
    float *dst = FIRST ? Ub : Rb;
#pragma unroll
    for (int i = 0; i < COB; ++i) {
        const float b = bg[co0 + i];
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const int p = pbase + 32 * j;
            const int t = o + p;
            float v = acc[i][j] + b;
            float *d = dst + (co0 + i) * PTOT + p;
            if (FIRST) v = lrelu(v); else v += *d;
            *d = (t >= 0 && t < Lout) ? v : 0.f;
        }
    }
    __syncthreads();
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, 1)
gen_stage_kernel(const float *__restrict__ x, float *__restrict__ y, const float *__restrict__ packed,
                 int stage, int Lin) {
    constexpr int CIN = Cfg::CIN, COUT = Cfg::COUT, S = Cfg::S, PTOT = Cfg::PTOT, PIN = Cfg::PIN;
    constexpr int COB = Cfg::COB, PB = Cfg::PB, HALO = Cfg::HALO, PVALID = Cfg::PVALID, NT = Cfg::NT;
    extern __shared__ __align__(16) float smem[];
    float *Rb = smem + Cfg::PADF;
    float *In = smem + Cfg::R_FLOATS;            // aliases U (dead before U is first written)
    float *Ub = smem + Cfg::R_FLOATS + Cfg::PADF;
    float *Wb = smem + Cfg::R_FLOATS + Cfg::U_FLOATS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wc = warp / Cfg::WP, wp = warp % Cfg::WP;
    const int co0 = wc * COB;
    const int pbase = wp * (32 * PB) + lane;
    const int b = blockIdx.y;
    const int Lout = Lin * S;
    const int o = blockIdx.x * PVALID - HALO;  // global position of tile-local p = 0

    // ---- ConvTranspose1d as a 2-tap gather (models.py:65): out[t] = sum_ci x[s]*w[ci][co][phi] + x[s-1]*w[ci][co][phi+S],
    //      phi = (t+PAD) mod S, s = (t+PAD) div S.
    constexpr int BIAS = 64;  // keeps (o + PAD) non-negative for the div/mod below (HALO < 64*S)
    const int s_begin = (o + Cfg::PAD + BIAS * S) / S - BIAS - 1;
    const int tp0 = o + pbase + Cfg::PAD + BIAS * S;
    const int phi = tp0 % S;
    const int ih0 = tp0 / S - BIAS - s_begin;  // local index of s for j = 0; j adds 32/S

    if (tid < Cfg::PADF) {
        smem[tid] = 0.f;
        smem[Cfg::R_FLOATS - Cfg::PADF + tid] = 0.f;
    }
    {
        const float *xb = x + (size_t)b * CIN * Lin;
        for (int idx = tid; idx < CIN * PIN; idx += NT) {
            const int ci = idx / PIN, i = idx - ci * PIN;
            const int s = s_begin + i;
            In[idx] = (s >= 0 && s < Lin) ? lrelu(xb[(size_t)ci * Lin + s]) : 0.f;
        }
    }
    {
        const int lup = 1 + stage;
        const float *__restrict__ wg = packed + weight_offset(lup);
        const float *__restrict__ bg = packed + bias_offset(lup);
        constexpr int CIC = Cfg::CIC_UP;
        constexpr int CH = CIC * COUT * 2 * S;
        constexpr int NCH = CIN / CIC;
        float acc[COB][PB];
#pragma unroll
        for (int i = 0; i < COB; ++i)
#pragma unroll
            for (int j = 0; j < PB; ++j) acc[i][j] = 0.f;
        load_chunk<NT>(Wb, wg, CH, tid);
        cp_async_commit();
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            if (c + 1 < NCH) load_chunk<NT>(Wb + ((c + 1) & 1) * Cfg::WBUF, wg + (size_t)(c + 1) * CH, CH, tid);
            cp_async_commit();
            cp_async_wait<1>();
            __syncthreads();
            const float *Wc = Wb + (c & 1) * Cfg::WBUF;
#pragma unroll
            for (int cl = 0; cl < CIC; ++cl) {
                const float *inrow = In + (c * CIC + cl) * PIN + ih0;
                float xh[PB], xl[PB];
#pragma unroll
                for (int j = 0; j < PB; ++j) {
                    xh[j] = inrow[j * (32 / S)];
                    xl[j] = inrow[j * (32 / S) - 1];
                }
                const float *wrow = Wc + ((cl * COUT + co0) * S + phi) * 2;
#pragma unroll
                for (int i = 0; i < COB; ++i) {
                    const float2 w = *reinterpret_cast<const float2 *>(wrow + i * S * 2);
#pragma unroll
                    for (int j = 0; j < PB; ++j) acc[i][j] = fmaf(w.x, xh[j], fmaf(w.y, xl[j], acc[i][j]));
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < COB; ++i) {
            const float bb = bg[co0 + i];
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                const int p = pbase + 32 * j;
                const int t = o + p;
                Rb[(co0 + i) * PTOT + p] = (t >= 0 && t < Lout) ? acc[i][j] + bb : 0.f;
            }
        }
    }
    __syncthreads();
    if (Cfg::RES) {
    if (tid < Cfg::PADF) {  // In is dead: give U its zero slack
        Ub[-Cfg::PADF + tid] = 0.f;
        Ub[COUT * PTOT + tid] = 0.f;
    }
    __syncthreads();

    // ---- ResBlock (models.py:32-40): three (dilated conv, conv) pairs with residual adds
    {
        const int l1 = 5 + 6 * stage, l2 = l1 + 3;
#define MG_PAIR(J, DIL)                                                                                   \
        conv3_phase<Cfg, DIL, true>(Rb, Ub, Wb, packed + weight_offset(l1 + J), packed + bias_offset(l1 + J), tid, \
                                    co0, pbase, o, Lout);                                                  \
        conv3_phase<Cfg, 1, false>(Rb, Ub, Wb, packed + weight_offset(l2 + J), packed + bias_offset(l2 + J), tid,  \
                                   co0, pbase, o, Lout);
        MG_PAIR(0, 1)
        MG_PAIR(1, 3)
        MG_PAIR(2, 9)
#undef MG_PAIR
    }
    }

    if (Cfg::POST) {
        // ---- LeakyReLU -> conv_post (32->1, k7) -> tanh (models.py:67-69); y is audio [B][Lout]
        float *wpost = Wb;
        const float *__restrict__ wg = packed + weight_offset(29);
        for (int i = tid; i < COUT * kPostK; i += NT) wpost[i] = wg[i];
        const float bpost = packed[bias_offset(29)];
        __syncthreads();
        for (int p = HALO + tid; p < PTOT - HALO; p += NT) {
            const int t = o + p;
            if (t < Lout) {
                float acc = bpost;
#pragma unroll 4
                for (int ci = 0; ci < COUT; ++ci) {
                    const float *r = Rb + ci * PTOT + p - 3;
#pragma unroll
                    for (int k = 0; k < kPostK; ++k) acc = fmaf(wpost[ci * kPostK + k], lrelu(r[k]), acc);
                }
                y[(size_t)b * Lout + t] = tanhf(acc);
            }
        }
    } else {
        float *yb = y + (size_t)b * COUT * Lout;
        for (int idx = tid; idx < COUT * PVALID; idx += NT) {
            const int co = idx / PVALID, pv = idx - co * PVALID;
            const int t = o + HALO + pv;
            if (t < Lout) yb[(size_t)co * Lout + t] = Rb[co * PTOT + HALO + pv];
        }
    }
}

// conv_pre: Conv1d(80 -> 512, k7, pad 3) (models.py:46,62).  CTA = kPreTT frames x 128 output channels (one per thread);
// 0.5% of the generator FLOPs, so plain fp32 FMAs with the mel tile broadcast from shared memory.
constexpr int kPreTT = 16;
constexpr int kPreCG = 128;
__global__ void __launch_bounds__(kPreCG)
gen_pre_kernel(const float *__restrict__ mel, float *__restrict__ y, const float *__restrict__ packed, int T) {
    constexpr int XS = kPreTT + 8;  // row stride (floats), multiple of 4
    __shared__ __align__(16) float xs[kMelBins * XS];
    const int b = blockIdx.z, t0 = blockIdx.x * kPreTT, co = blockIdx.y * kPreCG + threadIdx.x;
    for (int idx = threadIdx.x; idx < kMelBins * XS; idx += kPreCG) {
        const int ci = idx / XS, i = idx - ci * XS;
        const int t = t0 + i - 3;
        xs[idx] = (i < kPreTT + 6 && t >= 0 && t < T) ? mel[((size_t)b * kMelBins + ci) * T + t] : 0.f;
    }
    __syncthreads();
    const float *__restrict__ wg = packed + weight_offset(0) + co;
    float acc[kPreTT];
    const float bias = packed[bias_offset(0) + co];
#pragma unroll
    for (int t = 0; t < kPreTT; ++t) acc[t] = bias;
#pragma unroll 2
    for (int ci = 0; ci < kMelBins; ++ci) {
        float w[kPreK];
#pragma unroll
        for (int k = 0; k < kPreK; ++k) w[k] = __ldg(wg + (size_t)(ci * kPreK + k) * kPreCout);
        float xr[XS];
#pragma unroll
        for (int q = 0; q < XS / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(xs + ci * XS + 4 * q);
            xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < kPreK; ++k)
#pragma unroll
            for (int t = 0; t < kPreTT; ++t) acc[t] = fmaf(w[k], xr[t + k], acc[t]);
    }
    float *yr = y + ((size_t)b * kPreCout + co) * T + t0;
#pragma unroll
    for (int t = 0; t < kPreTT; ++t)
        if (t0 + t < T) yr[t] = acc[t];
}

//                     CIN  COUT S  PTOT WC WP WBUF  POST
using Stage0 = StageCfg<512, 256, 8,  96, 16, 1, 4096, false>;
using Stage1 = StageCfg<256, 128, 8, 192,  8, 2, 4096, false>;
using Stage2 = StageCfg<128,  64, 2, 384,  8, 2, 2048, false>;
using Stage3 = StageCfg< 64,  32, 2, 768,  4, 4, 2048, true>;
// ConvTranspose1d-only variants (no halo, no U buffer)
using Up0 = StageCfg<512, 256, 8,  96, 16, 1, 4096, false, false>;
using Up1 = StageCfg<256, 128, 8, 192,  8, 2, 4096, false, false>;
using Up2 = StageCfg<128,  64, 2, 384,  8, 2, 2048, false, false>;
using Up3 = StageCfg< 64,  32, 2, 768,  4, 4, 2048, false, false>;

template <class Cfg>
static int launch_stage(const float *x, float *y, const float *packed, int stage, int B, int Lin, cudaStream_t s) {
    static bool configured = false;  // benign race: the attribute call is idempotent
    if (!configured) {
        MG_CUDA_TRY(cudaFuncSetAttribute(gen_stage_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES));
        configured = true;
    }
    const int Lout = Lin * Cfg::S;
    dim3 grid((Lout + Cfg::PVALID - 1) / Cfg::PVALID, B);
    gen_stage_kernel<Cfg><<<grid, Cfg::NT, Cfg::SMEM_BYTES, s>>>(x, y, packed, stage, Lin);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

int generator_simt_num_launches() { return 5; }
// fp32 SIMT stand-ins of the tensor-core ConvT / conv_pre kernels for the pipeline in mg_gen_pipeline.cu (MG_UP_PATH=simt)
int launch_up_simt(const float *x, float *y, const float *packed, int stage, int B, int Lin, cudaStream_t s) {
    switch (stage) {
        case 0: return launch_stage<Up0>(x, y, packed, 0, B, Lin, s);
        case 1: return launch_stage<Up1>(x, y, packed, 1, B, Lin, s);
        case 2: return launch_stage<Up2>(x, y, packed, 2, B, Lin, s);
        case 3: return launch_stage<Up3>(x, y, packed, 3, B, Lin, s);
    }
    return set_error(MG_ERR_INVALID_ARGUMENT, "launch_up_simt: stage %d", stage);
}
int launch_pre_simt(const float *mel, float *y, const float *packed, int B, int T, cudaStream_t s) {
    dim3 gpre((T + kPreTT - 1) / kPreTT, kPreCout / kPreCG, B);
    gen_pre_kernel<<<gpre, kPreCG, 0, s>>>(mel, y, packed, T);
    MG_CUDA_TRY(cudaGetLastError());
    return MG_OK;
}

int launch_generator_simt(const float *packed, const float *mel, float *audio, int B, int T, float *ws,
                          cudaStream_t s, cudaEvent_t *ev) {
#define MG_MARK(i) do { if (ev) MG_CUDA_TRY(cudaEventRecord(ev[i], s)); } while (0)
    float *a0 = ws + ws_offset(0, B, T);  // [B,512,T]
    float *a1 = ws + ws_offset(1, B, T);  // [B,256,8T]
    float *a2 = ws + ws_offset(2, B, T);  // [B,128,64T]
    float *a3 = ws + ws_offset(3, B, T);  // [B,64,128T]
    dim3 gpre((T + kPreTT - 1) / kPreTT, kPreCout / kPreCG, B);
    MG_MARK(0);
    gen_pre_kernel<<<gpre, kPreCG, 0, s>>>(mel, a0, packed, T);
    MG_CUDA_TRY(cudaGetLastError());
    int rc;
    MG_MARK(1);
    if ((rc = launch_stage<Stage0>(a0, a1, packed, 0, B, T, s))) return rc;
    MG_MARK(2);
    if ((rc = launch_stage<Stage1>(a1, a2, packed, 1, B, 8 * T, s))) return rc;
    MG_MARK(3);
    if ((rc = launch_stage<Stage2>(a2, a3, packed, 2, B, 64 * T, s))) return rc;
    MG_MARK(4);
    if ((rc = launch_stage<Stage3>(a3, audio, packed, 3, B, 128 * T, s))) return rc;
    MG_MARK(5);
#undef MG_MARK
    return MG_OK;
}

}  // namespace mg
