/*
 * melgan_oracle.c -- CPU restatement of the reference's MelGAN hot path.  TEST INFRASTRUCTURE.
 *
 * This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The shipped path is the
 * CUDA library behind include/melgan_b200.h and fails loudly without a GPU.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the
 * restatement is pinned against outputs of the unmodified /root/reference/models.py executed
 * in the build container: tests/golden/make_golden.py (committed) imports the reference,
 * loads the seeded weights of melgan_multi_b200/synth.py and stores its outputs under
 * tests/golden/; tests/test_oracle.py checks this file against them.
 *
 * Every tensor is fp32, NCL-contiguous ([batch][channel][length]) like the reference's;
 * sums are accumulated in double so the oracle sits closer to the exact result than either
 * oneDNN or the GPU kernels do.
 *
 * Reference lines followed (all /root/reference/models.py):
 *   get_padding            :8-9      "same" padding (k*d - d)/2
 *   ResBlock.forward       :32-40    x = c2(lrelu(c1(lrelu(x)))) + x, three (c1, c2) pairs,
 *                                    c1 dilations 1/3/9 (:16-21), c2 dilation 1 (:23-28)
 *   Generator.forward      :61-71    conv_pre, 4 x (lrelu, ConvTranspose1d, ResBlock), lrelu,
 *                                    conv_post, tanh; layer shapes :46-59
 *   Discriminator.forward  :87-103   conv_pre, 4 grouped convs, conv_post1, conv_post2,
 *                                    lrelu after all but the last; 7 feature maps; shapes :77-85
 *   MultiScaleDiscriminator:114-117  AvgPool1d(4,2,pad 2) then AvgPool1d(4,4,pad 2), chained
 *                           :119-135 scale i sees pool_{i-1}(...pool_0(y))
 *   F.leaky_relu default slope 0.01 (:35,37,64,67,90,94,97)
 *   weight_norm(dim=0)     :16-28,46-59,77-85: w = g * v / ||v|| over all axes but 0, which is
 *                          per OUTPUT channel for Conv1d and per INPUT channel for
 *                          ConvTranspose1d (weight [C_in, C_out, K]).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define MGO_SLOPE 0.01f

/* w[i, :] = g[i] * v[i, :] / ||v[i, :]||_2 for i in [0, dim0); inner = product of other dims. */
void mgo_fold_weight_norm(const float *v, const float *g, float *w, int dim0, int inner)
{
    for (int i = 0; i < dim0; ++i) {
        double s = 0.0;
        for (int j = 0; j < inner; ++j) {
            double e = v[(size_t)i * inner + j];
            s += e * e;
        }
        double scale = (double)g[i] / sqrt(s);
        for (int j = 0; j < inner; ++j)
            w[(size_t)i * inner + j] = (float)(scale * (double)v[(size_t)i * inner + j]);
    }
}

int mgo_conv1d_out_len(int Lin, int K, int stride, int pad, int dil)
{
    return (Lin + 2 * pad - dil * (K - 1) - 1) / stride + 1;
}

/* torch.nn.Conv1d semantics.  w is [Cout][Cin/groups][K].  y is [B][Cout][Lout]. */
void mgo_conv1d(const float *x, const float *w, const float *bias, float *y,
                int B, int Cin, int Lin, int Cout, int K,
                int stride, int pad, int dil, int groups)
{
    const int Lout = mgo_conv1d_out_len(Lin, K, stride, pad, dil);
    const int cig = Cin / groups, cog = Cout / groups;
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * (size_t)(Lout > 0 ? Lout : 1));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            for (int co = 0; co < Cout; ++co) {
                const int grp = co / cog;
                for (int t = 0; t < Lout; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
                for (int c = 0; c < cig; ++c) {
                    const float *xr = x + ((size_t)b * Cin + (size_t)grp * cig + c) * Lin;
                    const float *wr = w + ((size_t)co * cig + c) * K;
                    for (int k = 0; k < K; ++k) {
                        const double wv = wr[k];
                        const int off = k * dil - pad; /* input index = t*stride + off */
                        int t0 = 0, t1 = Lout;
                        if (off < 0) t0 = (-off + stride - 1) / stride;
                        /* t*stride + off <= Lin-1 */
                        int tmax = (Lin - 1 - off) >= 0 ? (Lin - 1 - off) / stride + 1 : 0;
                        if (tmax < t1) t1 = tmax;
                        for (int t = t0; t < t1; ++t) acc[t] += wv * (double)xr[t * stride + off];
                    }
                }
                float *yr = y + ((size_t)b * Cout + co) * Lout;
                for (int t = 0; t < Lout; ++t) yr[t] = (float)acc[t];
            }
        }
        free(acc);
    }
}

/* torch.nn.ConvTranspose1d semantics (no output_padding, dilation 1, groups 1).
 * w is [Cin][Cout][K]; Lout = (Lin-1)*stride - 2*pad + K; scatter form t = s*stride - pad + k. */
void mgo_conv_transpose1d(const float *x, const float *w, const float *bias, float *y,
                          int B, int Cin, int Lin, int Cout, int K, int stride, int pad)
{
    const int Lout = (Lin - 1) * stride - 2 * pad + K;
#pragma omp parallel
    {
        double *acc = (double *)malloc(sizeof(double) * (size_t)(Lout > 0 ? Lout : 1));
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            for (int co = 0; co < Cout; ++co) {
                for (int t = 0; t < Lout; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
                for (int ci = 0; ci < Cin; ++ci) {
                    const float *xr = x + ((size_t)b * Cin + ci) * Lin;
                    const float *wr = w + ((size_t)ci * Cout + co) * K;
                    for (int k = 0; k < K; ++k) {
                        const double wv = wr[k];
                        for (int s = 0; s < Lin; ++s) {
                            const int t = s * stride - pad + k;
                            if (t >= 0 && t < Lout) acc[t] += wv * (double)xr[s];
                        }
                    }
                }
                float *yr = y + ((size_t)b * Cout + co) * Lout;
                for (int t = 0; t < Lout; ++t) yr[t] = (float)acc[t];
            }
        }
        free(acc);
    }
}

void mgo_leaky_relu(float *x, size_t n)
{
    for (size_t i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : x[i] * MGO_SLOPE;
}

void mgo_tanh(float *x, size_t n)
{
    for (size_t i = 0; i < n; ++i) x[i] = (float)tanh((double)x[i]);
}

int mgo_avgpool1d_out_len(int Lin, int k, int stride, int pad)
{
    return (Lin + 2 * pad - k) / stride + 1;
}

/* torch.nn.AvgPool1d(k, stride, padding=pad): count_include_pad=True, ceil_mode=False. */
void mgo_avgpool1d(const float *x, float *y, int rows, int Lin, int k, int stride, int pad)
{
    const int Lout = mgo_avgpool1d_out_len(Lin, k, stride, pad);
    for (int r = 0; r < rows; ++r)
        for (int t = 0; t < Lout; ++t) {
            double s = 0.0;
            for (int j = 0; j < k; ++j) {
                int i = t * stride - pad + j;
                if (i >= 0 && i < Lin) s += (double)x[(size_t)r * Lin + i];
            }
            y[(size_t)r * Lout + t] = (float)(s / (double)k);
        }
}

/* ------------------------------------------------------------------------------------- */
/* Generator.  w[30] / b[30] are FOLDED weights and biases in reference registration order:
 * 0 conv_pre; 1..4 ups; then per stage i: 5+6i+{0,1,2} convs1, 5+6i+{3,4,5} convs2; 29 post. */

static const int G_UP_CIN[4] = {512, 256, 128, 64};
static const int G_UP_K[4] = {16, 16, 4, 4};
static const int G_UP_S[4] = {8, 8, 2, 2};
static const int G_UP_P[4] = {4, 4, 1, 1};
static const int G_DIL[3] = {1, 3, 9};

/* stage_out: NULL or 6 pointers (each NULL or big enough): after conv_pre, after each of the
 * four ResBlocks, and the pre-tanh conv_post output.  Returns 0, or -1 on allocation failure. */
int mgo_generator_forward(const float *const *w, const float *const *b, const float *mel,
                          float *audio, int B, int T, float *const *stage_out)
{
    size_t maxel = (size_t)B * 512 * T; /* conv_pre output; later stages: C*L = 2048T,8192T,8192T,8192T */
    if ((size_t)B * 8192 * T > maxel) maxel = (size_t)B * 8192 * T;
    float *x = (float *)malloc(sizeof(float) * maxel);
    float *u = (float *)malloc(sizeof(float) * maxel);
    float *v = (float *)malloc(sizeof(float) * maxel);
    if (!x || !u || !v) { free(x); free(u); free(v); return -1; }

    int L = T;
    mgo_conv1d(mel, w[0], b[0], x, B, 80, L, 512, 7, 1, 3, 1, 1);
    if (stage_out && stage_out[0]) memcpy(stage_out[0], x, sizeof(float) * (size_t)B * 512 * L);
    for (int i = 0; i < 4; ++i) {
        const int cin = G_UP_CIN[i], c = cin / 2;
        mgo_leaky_relu(x, (size_t)B * cin * L);
        mgo_conv_transpose1d(x, w[1 + i], b[1 + i], u, B, cin, L, c, G_UP_K[i], G_UP_S[i], G_UP_P[i]);
        L = (L - 1) * G_UP_S[i] - 2 * G_UP_P[i] + G_UP_K[i];
        const size_t n = (size_t)B * c * L;
        /* u holds the residual stream */
        for (int j = 0; j < 3; ++j) {
            const int l1 = 5 + 6 * i + j, l2 = 5 + 6 * i + 3 + j;
            memcpy(x, u, sizeof(float) * n);
            mgo_leaky_relu(x, n);
            mgo_conv1d(x, w[l1], b[l1], v, B, c, L, c, 3, 1, G_DIL[j], G_DIL[j], 1);
            mgo_leaky_relu(v, n);
            mgo_conv1d(v, w[l2], b[l2], x, B, c, L, c, 3, 1, 1, 1, 1);
            for (size_t e = 0; e < n; ++e) u[e] = x[e] + u[e];
        }
        memcpy(x, u, sizeof(float) * n);
        if (stage_out && stage_out[1 + i]) memcpy(stage_out[1 + i], x, sizeof(float) * n);
    }
    mgo_leaky_relu(x, (size_t)B * 32 * L);
    mgo_conv1d(x, w[29], b[29], audio, B, 32, L, 1, 7, 1, 3, 1, 1);
    if (stage_out && stage_out[5]) memcpy(stage_out[5], audio, sizeof(float) * (size_t)B * L);
    mgo_tanh(audio, (size_t)B * L);
    free(x); free(u); free(v);
    return 0;
}

/* ------------------------------------------------------------------------------------- */
/* One Discriminator (models.py:87-103).  w[7]/b[7] folded, order: conv_pre, grouped 0..3,
 * conv_post1, conv_post2.  fmap[7] receive the seven feature maps (post-lrelu for the first
 * six, raw for the last, which is also the flattened logits).  fmap_len[7] receives lengths. */
static const int D_CIN[7] = {1, 16, 64, 256, 1024, 1024, 1024};
static const int D_COUT[7] = {16, 64, 256, 1024, 1024, 1024, 1};
static const int D_K[7] = {15, 41, 41, 41, 41, 5, 3};
static const int D_S[7] = {1, 4, 4, 4, 1, 1, 1};
static const int D_G[7] = {1, 4, 16, 64, 256, 1, 1};
static const int D_P[7] = {7, 20, 20, 20, 20, 2, 1};

void mgo_discriminator_lengths(int L, int *fmap_len)
{
    for (int l = 0; l < 7; ++l) {
        L = mgo_conv1d_out_len(L, D_K[l], D_S[l], D_P[l], 1);
        fmap_len[l] = L;
    }
}

void mgo_discriminator_forward(const float *const *w, const float *const *b, const float *y,
                               int B, int L, float *const *fmap)
{
    const float *in = y;
    for (int l = 0; l < 7; ++l) {
        mgo_conv1d(in, w[l], b[l], fmap[l], B, D_CIN[l], L, D_COUT[l], D_K[l], D_S[l], D_P[l], 1, D_G[l]);
        L = mgo_conv1d_out_len(L, D_K[l], D_S[l], D_P[l], 1);
        if (l < 6) mgo_leaky_relu(fmap[l], (size_t)B * D_COUT[l] * L);
        in = fmap[l];
    }
}
