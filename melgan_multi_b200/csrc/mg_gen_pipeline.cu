// The generator's tensor-core pipeline: which kernels run, in what order, on which streams.
//   conv_pre -> 4 x [LeakyReLU -> ConvTranspose1d -> ResBlock] (-> LeakyReLU -> conv_post -> tanh), models.py:61-71,
// as 8 kernels per chain (mg_conv_tc.cu, mg_up_tc.cu, mg_res_tc.cu), the batch cut into concurrent slices.
#include <stdlib.h>

#include "mg_common.cuh"

namespace mg {

// Which stride-2 stages run their ConvT inside the ResBlock kernel: bit 0 = stage 2, bit 1 = stage 3.  Default: stage 3
// only (measured at config 2: up3 + res3 194 -> 176 us fused, up2 + res2 221 -> 241 us: at C = 64 the extra serial phases
// of the fused tile cost more than the separate ConvT kernel).  MG_GEN_FUSE_UP = 0 / 2 / 3 / 23 overrides (A/B runs).
int generator_tc_fused_up() {
    static const int mask = [] {
        const char *e = getenv("MG_GEN_FUSE_UP");
        if (!e) return 2;
        int m = 0;
        for (; *e; ++e) m |= (*e == '2') ? 1 : (*e == '3') ? 2 : 0;
        return m;
    }();
    return mask;
}
int generator_tc_num_launches() { return 9 - (generator_tc_fused_up() & 1) - ((generator_tc_fused_up() >> 1) & 1); }

// Tensor-core pipeline of one contiguous slice of the batch: conv_pre -> 4 x [ConvT (tcgen05) -> ResBlock (tcgen05)], the
// last ResBlock with LeakyReLU -> conv_post -> tanh fused into its epilogue.
// a0, a[0..2], u: this slice's part of the workspace buffers.
static int generator_tc_chain(const float *packed, const float *mel, float *audio, int B, int T, float *a0, float *const *a,
                              float *u, int *status, cudaStream_t s, cudaEvent_t *ev) {
#define MG_MARK(i) do { if (ev) MG_CUDA_TRY(cudaEventRecord(ev[i], s)); } while (0)
    int rc;
    MG_MARK(0);
    if ((rc = launch_gen_pre_tc(mel, a0, packed, B, T, status, s))) return rc;
    MG_MARK(1);
    if ((rc = launch_convt_tc(a0, u, packed, 0, B, T, status, s))) return rc;
    MG_MARK(2);
    if ((rc = launch_resblock_tc(u, a[0], packed, 0, B, 8 * T, status, s))) return rc;
    MG_MARK(3);
    if ((rc = launch_convt_tc(a[0], u, packed, 1, B, 8 * T, status, s))) return rc;
    MG_MARK(4);
    if ((rc = launch_resblock_tc(u, a[1], packed, 1, B, 64 * T, status, s))) return rc;
    MG_MARK(5);
    // stages 2 and 3 can run LeakyReLU -> ConvT(k4, s2) -> ResBlock (-> conv_post -> tanh) as ONE kernel reading the previous
    // stage's output, so that the ConvT output never goes to HBM (generator_tc_fused_up(): bit 0 = stage 2, bit 1 = stage 3)
    const int fuse = generator_tc_fused_up();
    int m = 6;
    if (fuse & 1) {
        if ((rc = launch_resblock_tc(a[1], a[2], packed, 12, B, 128 * T, status, s))) return rc;
    } else {
        if ((rc = launch_convt_tc(a[1], u, packed, 2, B, 64 * T, status, s))) return rc;
        MG_MARK(m); ++m;
        if ((rc = launch_resblock_tc(u, a[2], packed, 2, B, 128 * T, status, s))) return rc;
    }
    MG_MARK(m); ++m;
    if (fuse & 2) {
        if ((rc = launch_resblock_tc(a[2], audio, packed, 14, B, 256 * T, status, s))) return rc;
    } else {
        if ((rc = launch_convt_tc(a[2], u, packed, 3, B, 128 * T, status, s))) return rc;
        MG_MARK(m); ++m;
        // stage 4 = ResBlock 3 with LeakyReLU -> conv_post -> tanh fused into its final epilogue: writes the audio
        if ((rc = launch_resblock_tc(u, audio, packed, 4, B, 256 * T, status, s))) return rc;
    }
    MG_MARK(m);
#undef MG_MARK
    return MG_OK;
}

// Side streams for batch slices (forked from / joined into the caller's stream with events: the call stays asynchronous
// and stream-ordered for the caller).  Per host thread, like the rest of the library's state.
constexpr int kMaxDevices = 64;
struct SliceStreams {
    static constexpr int kMax = 8;
    cudaStream_t st[kMax - 1] = {};
    cudaEvent_t fork = nullptr, join[kMax - 1] = {};
    bool ready = false;
    int init() {
        if (ready) return MG_OK;
        for (int i = 0; i < kMax - 1; ++i) {
            MG_CUDA_TRY(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            MG_CUDA_TRY(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
        }
        MG_CUDA_TRY(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        ready = true;
        return MG_OK;
    }
};

// Whole generator.  The batch items are independent and every kernel's grid is a whole number of tiles per item (or per
// 128 virtual rows), so the batch is cut into `slices` contiguous parts whose nine-kernel chains run on forked streams:
// the block scheduler fills the SMs one chain's partial last wave leaves idle (stage 0 at config 2 is 192 one-per-SM
// tiles on 148 SMs) with the other chain's tiles.  Same kernels, same per-item arithmetic: results are bit-identical
// to the single-chain order.  ev != nullptr (per-kernel timing) keeps everything on one stream.
int generator_tc_slices(int B, int T) {
    static const int forced = [] {  // MG_GEN_SLICES=n pins the slice count (experiments); default: chosen from the shape
        const char *e = getenv("MG_GEN_SLICES");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : v > SliceStreams::kMax ? SliceStreams::kMax : v;
    }();
    // slicing pays while a slice still fills the machine: >= 512 mel frames per slice (stage-0 tiles ~ frames / 11)
    int slices = forced ? forced : 4;
    while (slices > 1 && ((!forced && (long long)B * T < 512ll * slices) || B < slices)) --slices;
    return slices;
}

// mel_host / audio_host (both or neither; pinned): the host-buffer entry point's copies, cut the same way -- each slice's
// stream uploads its mel slice before its chain and downloads its audio slice after it, so all but the last download
// overlap the other chains' kernels.
int launch_generator_tc(const float *packed, const float *mel, float *audio, int B, int T, float *ws, int *status,
                        cudaStream_t s, cudaEvent_t *ev, const float *mel_host, float *audio_host) {
    const int slices = ev ? 1 : generator_tc_slices(B, T);
    float *base[6];
    for (int i = 0; i < 6; ++i) base[i] = ws + ws_offset(i, B, T);
    const size_t per_item[6] = {(size_t)512 * T, (size_t)256 * 8 * T, (size_t)128 * 64 * T, (size_t)64 * 128 * T, 0, (size_t)8192 * T};
    const size_t mel_item = (size_t)kMelBins * T, audio_item = (size_t)256 * T;
    // one pool per (host thread, device): streams and events belong to the device that was current when they were created
    static thread_local SliceStreams pools[kMaxDevices];
    int dev = 0;
    MG_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return set_error(MG_ERR_INVALID_ARGUMENT, "launch_generator_tc: device ordinal %d", dev);
    SliceStreams &ss = pools[dev];
    int rc = MG_OK;
    if (slices > 1) {
        if ((rc = ss.init())) return rc;
        MG_CUDA_TRY(cudaEventRecord(ss.fork, s));
    }
    int forked = 0;  // side streams that wait on `fork` so far: all of them are joined back, also on the error path
    for (int k = 0, b0 = 0; k < slices && rc == MG_OK; ++k) {
        const int nb = B / slices + (k < B % slices);
        cudaStream_t q = k == 0 ? s : ss.st[k - 1];
        auto slice = [&]() -> int {
            if (k > 0) {
                MG_CUDA_TRY(cudaStreamWaitEvent(q, ss.fork, 0));
                forked = k;
            }
            if (mel_host)
                MG_CUDA_TRY(cudaMemcpyAsync(const_cast<float *>(mel) + b0 * mel_item, mel_host + b0 * mel_item,
                                            nb * mel_item * sizeof(float), cudaMemcpyHostToDevice, q));
            float *a[3] = {base[1] + b0 * per_item[1], base[2] + b0 * per_item[2], base[3] + b0 * per_item[3]};
            int r = generator_tc_chain(packed, mel + b0 * mel_item, audio + b0 * audio_item, nb, T, base[0] + b0 * per_item[0], a,
                                       base[5] + b0 * per_item[5], status, q, ev);
            if (r) return r;
            if (audio_host)
                MG_CUDA_TRY(cudaMemcpyAsync(audio_host + b0 * audio_item, audio + b0 * audio_item, nb * audio_item * sizeof(float),
                                            cudaMemcpyDeviceToHost, q));
            return MG_OK;
        };
        rc = slice();
        b0 += nb;
    }
    for (int k = 1; k <= forked; ++k) {  // join (best effort after an error: the caller's stream must not outrun a forked one)
        cudaError_t e = cudaEventRecord(ss.join[k - 1], ss.st[k - 1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, ss.join[k - 1], 0);
        if (e != cudaSuccess && rc == MG_OK) rc = set_error(MG_ERR_CUDA, "launch_generator_tc: joining slice %d: %s", k, cudaGetErrorString(e));
    }
    return rc;
}

}  // namespace mg
