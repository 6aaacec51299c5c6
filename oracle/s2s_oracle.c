/*
 * oracle/s2s_oracle.c -- CPU restatement of the seq2seq attention-ASR training hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA library in
 * seq2seq-attention-asr_b200/csrc and the timed "cpu_baseline" / "--impl reference" arm of
 * bench.py.  The product path never links, imports or calls anything in oracle/.
 *
 * PARITY PINNING: the reference (Ajay-Wong/seq2seq-attention-asr) is Lua on Torch7; neither
 * LuaJIT nor Torch7 exists in the build image, and all arithmetic lives in un-vendored,
 * un-pinned Torch7 packages (nn, nngraph, optim).  The only reproducible known-answer vectors
 * the reference holds are four notebook cells (Attention.ipynb:123-154, :257, :725-752,
 * :918-958); the oracle's primitives are pinned against those in tests/test_oracle_golden.py.
 * For every numeric output of the actual hot path (alpha, c, logp, gradients, NLL on given
 * weights) the reference holds no vector:  **parity unpinned** beyond those cells.  The
 * oracle's hand-written backward is instead cross-checked against an independent float64
 * autograd derivation (tests/torch_ref.py).
 *
 * Compiled twice by oracle/Makefile: -DREAL=float  -> liboracle_f32.so  (fp32 parity target,
 * CPU baseline) and -DREAL=double -> liboracle_f64.so (ground truth for tolerance checks).
 *
 * Every function follows the reference's per-utterance ("SGD" / non-batch) op order -- the
 * reference never batches: timit/timit.lua:240-289 loops utterances one at a time.
 * All reference citations are file:line under /root/reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Model configuration (mirrors timit/model_chorowski_baseline.lua:14-46 `model.*` fields)
 * cfg[] is a plain int array so ctypes can pass it without struct mirroring.
 * ------------------------------------------------------------------------------------------ */
enum { CFG_D = 0,   /* inputFrameSize (123) */
       CFG_H,       /* hiddenFrameSize == outputFrameSize (256) */
       CFG_NL,      /* number of bidirectional encoder layers (3) */
       CFG_S,       /* scoreDepth (512) */
       CFG_ST,      /* stateDepth (256) */
       CFG_V,       /* outputDepth / numPhonemes (62) */
       CFG_K,       /* hybridAttendFeatureMaps (0 = content-only) */
       CFG_KF,      /* hybridAttendFilterSize (10) */
       CFG_M,       /* mlpDepth (64) */
       CFG_MW,      /* maxout window (7) */
       CFG_MLP,     /* decoder MLP: 1 = Maxout-Linear (timit/model_chorowski_baseline.lua:56-57);
                       2 = Maxout-Linear-Maxout-Linear (librispeech/model_vgg.lua:76-80) */
       CFG_N };

/* Flat parameter layout (builder-defined; the Lua shim's parameters() returns tensors in this
 * order so getParameters() flattens to it).  Segment ids: */
enum { P_ENC = 0 /* enc: layer l, dir d, gate g -> P_ENC + (l*2+d)*3+g ; g: 0=z 1=r 2=h~ */ };
#define MAXSEG 64
typedef struct { int64_t off, rows, cols; } seg_t;
typedef struct {
    int D, H, NL, S, A, ST, V, K, KF, M, MW, MLP;
    seg_t enc[8][2][3];
    seg_t WV, bV, Ws, bs, WF, bF, U, bU, we, be, Wy, by, Wc, bc, Wj, bj, Gz, Gr, Gh, Wm, bm, Wl, bl, Wm2, bm2, Wo, bo;
    int64_t n;
} layout_t;

static void seg(seg_t* s, int64_t* off, int64_t rows, int64_t cols) { s->off = *off; s->rows = rows; s->cols = cols; *off += rows * cols; }

static void make_layout(const int* cfg, layout_t* Y) {
    memset(Y, 0, sizeof(*Y));
    Y->D = cfg[CFG_D]; Y->H = cfg[CFG_H]; Y->NL = cfg[CFG_NL]; Y->S = cfg[CFG_S]; Y->A = 2 * Y->H;
    Y->ST = cfg[CFG_ST]; Y->V = cfg[CFG_V]; Y->K = cfg[CFG_K]; Y->KF = cfg[CFG_KF]; Y->M = cfg[CFG_M]; Y->MW = cfg[CFG_MW];
    Y->MLP = cfg[CFG_MLP] == 2 ? 2 : 1;
    int64_t o = 0;
    for (int l = 0; l < Y->NL; l++) {
        int din = l == 0 ? Y->D : 2 * Y->H;                       /* model_chorowski_baseline.lua:22-31 */
        for (int d = 0; d < 2; d++) for (int g = 0; g < 3; g++) seg(&Y->enc[l][d][g], &o, Y->H, Y->H + din);   /* GRU.lua:23-26 */
    }
    seg(&Y->WV, &o, Y->S, Y->A); seg(&Y->bV, &o, Y->S, 1);        /* Attention.lua:44 (bias dead, kept: Q6) */
    seg(&Y->Ws, &o, Y->S, Y->ST); seg(&Y->bs, &o, Y->S, 1);       /* Attention.lua:66 */
    if (Y->K > 0) {
        seg(&Y->WF, &o, Y->K, Y->KF); seg(&Y->bF, &o, Y->K, 1);   /* Attention.lua:90 */
        seg(&Y->U, &o, Y->S, Y->K);  seg(&Y->bU, &o, Y->S, 1);    /* Attention.lua:91 */
    }
    seg(&Y->we, &o, 1, Y->S); seg(&Y->be, &o, 1, 1);              /* Attention.lua:110 */
    seg(&Y->Wy, &o, Y->ST, Y->V); seg(&Y->by, &o, Y->ST, 1);      /* Attention.lua:149 */
    seg(&Y->Wc, &o, Y->ST, Y->A); seg(&Y->bc, &o, Y->ST, 1);      /* Attention.lua:150 */
    seg(&Y->Wj, &o, Y->ST, 2 * Y->ST); seg(&Y->bj, &o, Y->ST, 1); /* Attention.lua:151 */
    seg(&Y->Gz, &o, Y->ST, 2 * Y->ST); seg(&Y->Gr, &o, Y->ST, 2 * Y->ST); seg(&Y->Gh, &o, Y->ST, 2 * Y->ST); /* model:50 */
    seg(&Y->Wm, &o, Y->M * Y->MW, Y->ST + Y->A); seg(&Y->bm, &o, Y->M * Y->MW, 1); /* model:56, Maxout.lua:15 */
    if (Y->MLP == 2) {                                             /* model_vgg.lua:78-79 */
        seg(&Y->Wl, &o, Y->M, Y->M); seg(&Y->bl, &o, Y->M, 1);
        seg(&Y->Wm2, &o, Y->M * Y->MW, Y->M); seg(&Y->bm2, &o, Y->M * Y->MW, 1);
    }
    seg(&Y->Wo, &o, Y->V, Y->M); seg(&Y->bo, &o, Y->V, 1);        /* model:57 / model_vgg.lua:80 */
    Y->n = o;
}

EXPORT int64_t orc_param_count(const int* cfg) { layout_t Y; make_layout(cfg, &Y); return Y.n; }

/* Writes (off, rows, cols) triples of every segment, in flat order; returns the count. */
EXPORT int orc_param_segments(const int* cfg, int64_t* out) {
    layout_t Y; make_layout(cfg, &Y); int n = 0;
#define PUT(s) do { if ((s).rows) { out[3*n] = (s).off; out[3*n+1] = (s).rows; out[3*n+2] = (s).cols; n++; } } while (0)
    for (int l = 0; l < Y.NL; l++) for (int d = 0; d < 2; d++) for (int g = 0; g < 3; g++) PUT(Y.enc[l][d][g]);
    PUT(Y.WV); PUT(Y.bV); PUT(Y.Ws); PUT(Y.bs); PUT(Y.WF); PUT(Y.bF); PUT(Y.U); PUT(Y.bU); PUT(Y.we); PUT(Y.be);
    PUT(Y.Wy); PUT(Y.by); PUT(Y.Wc); PUT(Y.bc); PUT(Y.Wj); PUT(Y.bj); PUT(Y.Gz); PUT(Y.Gr); PUT(Y.Gh);
    PUT(Y.Wm); PUT(Y.bm); PUT(Y.Wl); PUT(Y.bl); PUT(Y.Wm2); PUT(Y.bm2); PUT(Y.Wo); PUT(Y.bo);
#undef PUT
    return n;
}

/* ------------------------------------------------------------------------------------------
 * BLAS-1/2 helpers (the reference's addmv / addr calls, LinearZeroBias.lua:34,58,70)
 * ------------------------------------------------------------------------------------------ */
static inline real dotr(const real* a, const real* b, int n) {
    real acc = 0;
#pragma omp simd reduction(+ : acc)
    for (int i = 0; i < n; i++) acc += a[i] * b[i];
    return acc;
}
/* y (+)= W x ; W is rows x cols with leading dimension ld */
static void gemv(const real* W, int ld, int rows, int cols, const real* x, real* y, int acc) {
    for (int i = 0; i < rows; i++) { real v = dotr(W + (size_t)i * ld, x, cols); y[i] = acc ? y[i] + v : v; }
}
/* dx (+)= W^T dy */
static void gemv_t(const real* W, int ld, int rows, int cols, const real* dy, real* dx, int acc) {
    if (!acc) memset(dx, 0, sizeof(real) * cols);
    for (int i = 0; i < rows; i++) {
        const real g = dy[i]; const real* w = W + (size_t)i * ld;
        if (g == 0) continue;
#pragma omp simd
        for (int j = 0; j < cols; j++) dx[j] += g * w[j];
    }
}
/* dW += dy x^T */
static void ger(real* dW, int ld, int rows, int cols, const real* dy, const real* x) {
    for (int i = 0; i < rows; i++) {
        const real g = dy[i]; real* w = dW + (size_t)i * ld;
        if (g == 0) continue;
#pragma omp simd
        for (int j = 0; j < cols; j++) w[j] += g * x[j];
    }
}
static inline real sigm(real x) { return (real)1 / ((real)1 + (real)exp(-(double)x)); }
static inline real tanhr(real x) { return (real)tanh((double)x); }

/* ------------------------------------------------------------------------------------------
 * Primitive semantics pinned by the reference's notebook cells
 * ------------------------------------------------------------------------------------------ */
/* nn.TemporalConvolution (un-vendored Torch7 `nn`; call sites Attention.lua:66,90 and
 * TemporalConvolutionZeroBias.lua:39): cross-correlation, out[t] = b + W . vec(in[t*dW : t*dW+kW, :]),
 * weight [out, kW*in] frame-major.  Pinned by Attention.ipynb:123-154 (all rows = 15 40 65 90). */
EXPORT void orc_tconv_forward(const real* x, int L, int in, const real* W, const real* b, int out, int kW, int dW, real* y) {
    int Lo = (L - kW) / dW + 1;
    for (int t = 0; t < Lo; t++)
        for (int o = 0; o < out; o++) {
            real v = b ? b[o] : 0;
            v += dotr(W + (size_t)o * kW * in, x + (size_t)t * dW * in, kW * in);
            y[(size_t)t * out + o] = v;
        }
}
/* nn.Padding(dim=1, pad, nInputDim=2) on [L, F]: pad<0 prepends |pad| zero frames, pad>0 appends.
 * Pinned by Attention.ipynb:257 (0 0 1x10 0 0).  Call site Attention.lua:88. */
EXPORT void orc_padding(const real* x, int L, int F, int pad, real* y) {
    int p = pad < 0 ? -pad : pad;
    memset(y, 0, sizeof(real) * (size_t)(L + p) * F);
    memcpy(y + (pad < 0 ? (size_t)p * F : 0), x, sizeof(real) * (size_t)L * F);
}
/* nn.MM {a[1,L], h[L,A]} -> [1,A]  (Attention.lua:134).  Pinned by Attention.ipynb:918-958. */
EXPORT void orc_mm(const real* a, const real* b, int m, int k, int n, real* c) {
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
        real v = 0; for (int p = 0; p < k; p++) v += a[(size_t)i * k + p] * b[(size_t)p * n + j];
        c[(size_t)i * n + j] = v;
    }
}
/* nn.AddBias gradBias = sum of gradOutput over frames (AddBias.lua:60-93).  Pinned by
 * Attention.ipynb:725-752 (9 for L=9, 27 for batch 3x9).  Not on the hot path (never instantiated). */
EXPORT real orc_addbias_gradbias(const real* gradOutput, int nframes) {
    real v = 0; for (int i = 0; i < nframes; i++) v += gradOutput[i]; return v;
}

/* ------------------------------------------------------------------------------------------
 * nn.GRU step  (GRU.lua:22-30; weights LinearZeroBias [out, out+in], concat order {prev_h, x})
 * ------------------------------------------------------------------------------------------ */
typedef struct { const real *Wz, *Wr, *Wh; real *dWz, *dWr, *dWh; int in, out; } gru_w;

/* saves z, r, hc (candidate) for backward */
static void gru_step_fwd(const gru_w* g, const real* x, const real* hp, real* hn, real* z, real* r, real* hc, real* rh) {
    const int H = g->out, D = g->in, ld = H + D;
    for (int i = 0; i < H; i++) {
        z[i] = sigm(dotr(g->Wz + (size_t)i * ld, hp, H) + dotr(g->Wz + (size_t)i * ld + H, x, D));   /* GRU.lua:23 */
        r[i] = sigm(dotr(g->Wr + (size_t)i * ld, hp, H) + dotr(g->Wr + (size_t)i * ld + H, x, D));   /* GRU.lua:24 */
    }
    for (int i = 0; i < H; i++) rh[i] = r[i] * hp[i];                                                 /* GRU.lua:25 */
    for (int i = 0; i < H; i++)
        hc[i] = tanhr(dotr(g->Wh + (size_t)i * ld, rh, H) + dotr(g->Wh + (size_t)i * ld + H, x, D)); /* GRU.lua:26 */
    for (int i = 0; i < H; i++) hn[i] = ((real)1 - z[i]) * hp[i] + z[i] * hc[i];                      /* GRU.lua:27-30 */
}

/* dhn -> dx (overwritten), dhp (overwritten), dW accumulated.  scratch: 4*H reals */
static void gru_step_bwd(const gru_w* g, const real* x, const real* hp, const real* z, const real* r, const real* hc,
                         const real* dhn, real* dx, real* dhp, real* scratch) {
    const int H = g->out, D = g->in, ld = H + D;
    real *daz = scratch, *dar = scratch + H, *dah = scratch + 2 * H, *tmp = scratch + 3 * H;
    for (int i = 0; i < H; i++) {
        real dhc = dhn[i] * z[i];
        real dz = dhn[i] * (hc[i] - hp[i]);
        dhp[i] = dhn[i] * ((real)1 - z[i]);
        dah[i] = dhc * ((real)1 - hc[i] * hc[i]);
        daz[i] = dz * z[i] * ((real)1 - z[i]);
    }
    /* candidate linear: input [r*hp ; x] */
    for (int i = 0; i < H; i++) tmp[i] = r[i] * hp[i];
    for (int i = 0; i < H; i++) { if (dah[i] == 0) continue;
        real* w = g->dWh + (size_t)i * ld; const real gi = dah[i];
        for (int j = 0; j < H; j++) w[j] += gi * tmp[j];
        for (int j = 0; j < D; j++) w[H + j] += gi * x[j]; }
    gemv_t(g->Wh, ld, H, H, dah, tmp, 0);            /* d(r*hp) */
    gemv_t(g->Wh + H, ld, H, D, dah, dx, 0);
    for (int i = 0; i < H; i++) { dar[i] = tmp[i] * hp[i] * r[i] * ((real)1 - r[i]); dhp[i] += tmp[i] * r[i]; }
    for (int i = 0; i < H; i++) {
        real* wz = g->dWz + (size_t)i * ld; real* wr = g->dWr + (size_t)i * ld; const real gz = daz[i], gr = dar[i];
        for (int j = 0; j < H; j++) { wz[j] += gz * hp[j]; wr[j] += gr * hp[j]; }
        for (int j = 0; j < D; j++) { wz[H + j] += gz * x[j]; wr[H + j] += gr * x[j]; }
    }
    gemv_t(g->Wz, ld, H, H, daz, dhp, 1); gemv_t(g->Wr, ld, H, H, dar, dhp, 1);
    gemv_t(g->Wz + H, ld, H, D, daz, dx, 1); gemv_t(g->Wr + H, ld, H, D, dar, dx, 1);
}

EXPORT void orc_gru_step_forward(const real* Wz, const real* Wr, const real* Wh, int in, int out,
                                 const real* x, const real* hp, real* hn, real* z, real* r, real* hc) {
    gru_w g = { Wz, Wr, Wh, 0, 0, 0, in, out };
    real* rh = (real*)malloc(sizeof(real) * out);
    gru_step_fwd(&g, x, hp, hn, z, r, hc, rh); free(rh);
}
EXPORT void orc_gru_step_backward(const real* Wz, const real* Wr, const real* Wh, real* dWz, real* dWr, real* dWh, int in, int out,
                                  const real* x, const real* hp, const real* z, const real* r, const real* hc,
                                  const real* dhn, real* dx, real* dhp) {
    gru_w g = { Wz, Wr, Wh, dWz, dWr, dWh, in, out };
    real* s = (real*)malloc(sizeof(real) * 4 * out);
    gru_step_bwd(&g, x, hp, z, r, hc, dhn, dx, dhp, s); free(s);
}

/* ------------------------------------------------------------------------------------------
 * nn.RNN over a whole utterance (RNN.lua:120-201): unroll nn.GRU forward or reversed; output =
 * hidden sequence; zero initial state (Recurrent.lua:13,112).  x [L, in] (row stride ldx),
 * out [L, out] written with row stride ldo (so fwd/rev land in the JoinTable(2,2) halves,
 * model_chorowski_baseline.lua:24).  gates [L, 3*out] saved (z, r, hc).
 * ------------------------------------------------------------------------------------------ */
static void gru_seq_fwd(const gru_w* g, const real* x, int ldx, int L, int reverse, real* out, int ldo, real* gates) {
    const int H = g->out;
    real* zero = (real*)calloc(H, sizeof(real)); real* rh = (real*)malloc(sizeof(real) * H);
    const real* hp = zero;
    for (int s = 0; s < L; s++) {
        int t = reverse ? L - 1 - s : s;                           /* RNN.lua:142-145 */
        real* gt = gates + (size_t)t * 3 * H;
        gru_step_fwd(g, x + (size_t)t * ldx, hp, out + (size_t)t * ldo, gt, gt + H, gt + 2 * H, rh);  /* RNN.lua:155-162 */
        hp = out + (size_t)t * ldo;
    }
    free(zero); free(rh);
}
/* dout [L, out] (stride ldd) -> dx [L, in] (stride lddx; ACCUMULATED so the two directions sum as
 * nngraph does at the fan-out of the shared input, model_chorowski_baseline.lua:22-23) */
static void gru_seq_bwd(const gru_w* g, const real* x, int ldx, int L, int reverse, const real* out, int ldo, const real* gates,
                        const real* dout, int ldd, real* dx, int lddx) {
    const int H = g->out, D = g->in;
    real* zero = (real*)calloc(H, sizeof(real)); real* dh = (real*)calloc(H, sizeof(real));
    real* dhp = (real*)malloc(sizeof(real) * H); real* dxt = (real*)malloc(sizeof(real) * D); real* scr = (real*)malloc(sizeof(real) * 4 * H);
    for (int s = L - 1; s >= 0; s--) {                               /* RNN.lua:183 */
        int t = reverse ? L - 1 - s : s; int tp = reverse ? t + 1 : t - 1;
        const real* hp = (s == 0) ? zero : out + (size_t)tp * ldo;   /* RNN.lua:186-192 */
        const real* gt = gates + (size_t)t * 3 * H;
        for (int i = 0; i < H; i++) dh[i] += dout[(size_t)t * ldd + i];       /* RNN.lua:193-194 dEdy + dEdpy */
        gru_step_bwd(g, x + (size_t)t * ldx, hp, gt, gt + H, gt + 2 * H, dh, dxt, dhp, scr);
        if (dx) for (int j = 0; j < D; j++) dx[(size_t)t * lddx + j] += dxt[j];
        memcpy(dh, dhp, sizeof(real) * H);
    }
    free(zero); free(dh); free(dhp); free(dxt); free(scr);
}

EXPORT void orc_gru_seq_forward(const real* Wz, const real* Wr, const real* Wh, int in, int out, const real* x, int L, int reverse,
                                real* y, real* gates) {
    gru_w g = { Wz, Wr, Wh, 0, 0, 0, in, out };
    gru_seq_fwd(&g, x, in, L, reverse, y, out, gates);
}
EXPORT void orc_gru_seq_backward(const real* Wz, const real* Wr, const real* Wh, real* dWz, real* dWr, real* dWh, int in, int out,
                                 const real* x, int L, int reverse, const real* y, const real* gates, const real* dy, real* dx) {
    gru_w g = { Wz, Wr, Wh, dWz, dWr, dWh, in, out };
    memset(dx, 0, sizeof(real) * (size_t)L * in);
    gru_seq_bwd(&g, x, in, L, reverse, y, out, gates, dy, out, dx, in);
}

/* ------------------------------------------------------------------------------------------
 * nn.LSTM step (LSTM.lua:25-58): per gate Linear(in->out)+Linear(out->out), both WITH bias;
 * optional full-matrix peepholes Linear(out->out) (also with bias) on prev_c (i,f) / next_c (o).
 * Parameter block order: for gate in (i, f, g, o): Wx[out,in], bx[out], Wh[out,out], bh[out],
 * [Wc[out,out], bc[out] if peepholes and gate != g].
 * ------------------------------------------------------------------------------------------ */
static int64_t lstm_gate_size(int in, int out, int peep) { return (int64_t)out * in + out + (int64_t)out * out + out + (peep ? (int64_t)out * out + out : 0); }
EXPORT int64_t orc_lstm_param_count(int in, int out, int peepholes) {
    return 3 * lstm_gate_size(in, out, peepholes) + lstm_gate_size(in, out, 0);
}
static const real* lstm_gate_ptr(const real* P, int in, int out, int peep, int gate) {
    /* gates i(0), f(1) carry peepholes, g(2) never, o(3) carries */
    int64_t o = 0;
    for (int k = 0; k < gate; k++) o += lstm_gate_size(in, out, peep && k != 2);
    return P + o;
}
static void lstm_gate_pre(const real* P, int in, int out, int peep, const real* x, const real* h, const real* c, real* a) {
    const real* Wx = P; const real* bx = Wx + (size_t)out * in; const real* Wh = bx + out; const real* bh = Wh + (size_t)out * out;
    for (int i = 0; i < out; i++) a[i] = bx[i] + dotr(Wx + (size_t)i * in, x, in) + bh[i] + dotr(Wh + (size_t)i * out, h, out);
    if (peep) { const real* Wc = bh + out; const real* bc = Wc + (size_t)out * out;
        for (int i = 0; i < out; i++) a[i] += bc[i] + dotr(Wc + (size_t)i * out, c, out); }
}
/* saves gate activations acts[4*out] = (i, f, g, o) and tanh(next_c) is recomputed in bwd */
EXPORT void orc_lstm_step_forward(const real* P, int in, int out, int peep, const real* x, const real* hp, const real* cp,
                                  real* hn, real* cn, real* acts) {
    real *ig = acts, *fg = acts + out, *gg = acts + 2 * out, *og = acts + 3 * out;
    lstm_gate_pre(lstm_gate_ptr(P, in, out, peep, 0), in, out, peep, x, hp, cp, ig);      /* LSTM.lua:42 */
    lstm_gate_pre(lstm_gate_ptr(P, in, out, peep, 1), in, out, peep, x, hp, cp, fg);      /* LSTM.lua:43 */
    lstm_gate_pre(lstm_gate_ptr(P, in, out, peep, 2), in, out, 0, x, hp, cp, gg);         /* LSTM.lua:44 */
    for (int i = 0; i < out; i++) { ig[i] = sigm(ig[i]); fg[i] = sigm(fg[i]); gg[i] = tanhr(gg[i]); }
    for (int i = 0; i < out; i++) cn[i] = fg[i] * cp[i] + ig[i] * gg[i];                  /* LSTM.lua:45-46 */
    lstm_gate_pre(lstm_gate_ptr(P, in, out, peep, 3), in, out, peep, x, hp, cn, og);      /* LSTM.lua:47-50 */
    for (int i = 0; i < out; i++) { og[i] = sigm(og[i]); hn[i] = og[i] * tanhr(cn[i]); }  /* LSTM.lua:51 */
}
static void lstm_gate_bwd(const real* P, real* dP, int in, int out, int peep, const real* x, const real* h, const real* c,
                          const real* da, real* dx, real* dh, real* dc) {
    const real* Wx = P; const real* Wh = Wx + (size_t)out * in + out;
    real* dWx = dP; real* dbx = dWx + (size_t)out * in; real* dWh = dbx + out; real* dbh = dWh + (size_t)out * out;
    ger(dWx, in, out, in, da, x); ger(dWh, out, out, out, da, h);
    for (int i = 0; i < out; i++) { dbx[i] += da[i]; dbh[i] += da[i]; }
    gemv_t(Wx, in, out, in, da, dx, 1); gemv_t(Wh, out, out, out, da, dh, 1);
    if (peep) { const real* Wc = Wh + (size_t)out * out + out; real* dWc = dbh + out; real* dbc = dWc + (size_t)out * out;
        ger(dWc, out, out, out, da, c); for (int i = 0; i < out; i++) dbc[i] += da[i];
        gemv_t(Wc, out, out, out, da, dc, 1); }
}
/* (dhn, dcn) -> dx, dhp, dcp (overwritten); dP accumulated.  LSTM.lua:118-136 */
EXPORT void orc_lstm_step_backward(const real* P, real* dP, int in, int out, int peep, const real* x, const real* hp, const real* cp,
                                   const real* cn, const real* acts, const real* dhn, const real* dcn_in,
                                   real* dx, real* dhp, real* dcp) {
    const real *ig = acts, *fg = acts + out, *gg = acts + 2 * out, *og = acts + 3 * out;
    real* da = (real*)malloc(sizeof(real) * out); real* dcn = (real*)malloc(sizeof(real) * out);
    memset(dx, 0, sizeof(real) * in); memset(dhp, 0, sizeof(real) * out); memset(dcp, 0, sizeof(real) * out);
    for (int i = 0; i < out; i++) { real tc = tanhr(cn[i]);
        dcn[i] = (dcn_in ? dcn_in[i] : 0) + dhn[i] * og[i] * ((real)1 - tc * tc);
        da[i] = dhn[i] * tc * og[i] * ((real)1 - og[i]); }
#define GOFF(k) (lstm_gate_ptr(P, in, out, peep, k) - P)
    lstm_gate_bwd(P + GOFF(3), dP + GOFF(3), in, out, peep, x, hp, cn, da, dx, dhp, dcn);   /* out gate peeps at next_c */
    for (int i = 0; i < out; i++) da[i] = dcn[i] * ig[i] * ((real)1 - gg[i] * gg[i]);
    lstm_gate_bwd(P + GOFF(2), dP + GOFF(2), in, out, 0, x, hp, cp, da, dx, dhp, dcp);
    for (int i = 0; i < out; i++) da[i] = dcn[i] * cp[i] * fg[i] * ((real)1 - fg[i]);
    lstm_gate_bwd(P + GOFF(1), dP + GOFF(1), in, out, peep, x, hp, cp, da, dx, dhp, dcp);
    for (int i = 0; i < out; i++) da[i] = dcn[i] * gg[i] * ig[i] * ((real)1 - ig[i]);
    lstm_gate_bwd(P + GOFF(0), dP + GOFF(0), in, out, peep, x, hp, cp, da, dx, dhp, dcp);
#undef GOFF
    for (int i = 0; i < out; i++) dcp[i] += dcn[i] * fg[i];
    free(da); free(dcn);
}
/* nn.RNN(nn.LSTM) over an utterance: RNN.lua:153-164 passes {x, y, h}; LSTM returns {h, c} so
 * y = next_h, h = next_c.  Saves c sequence + acts. */
EXPORT void orc_lstm_seq_forward(const real* P, int in, int out, int peep, const real* x, int L, int reverse,
                                 real* y, real* cseq, real* acts) {
    real* zero = (real*)calloc(out, sizeof(real)); const real *hp = zero, *cp = zero;
    for (int s = 0; s < L; s++) { int t = reverse ? L - 1 - s : s;
        orc_lstm_step_forward(P, in, out, peep, x + (size_t)t * in, hp, cp, y + (size_t)t * out, cseq + (size_t)t * out, acts + (size_t)t * 4 * out);
        hp = y + (size_t)t * out; cp = cseq + (size_t)t * out; }
    free(zero);
}
EXPORT void orc_lstm_seq_backward(const real* P, real* dP, int in, int out, int peep, const real* x, int L, int reverse,
                                  const real* y, const real* cseq, const real* acts, const real* dy, real* dx) {
    real* zero = (real*)calloc(out, sizeof(real)); real* dh = (real*)calloc(out, sizeof(real)); real* dc = (real*)calloc(out, sizeof(real));
    real* dhp = (real*)malloc(sizeof(real) * out); real* dcp = (real*)malloc(sizeof(real) * out);
    for (int s = L - 1; s >= 0; s--) { int t = reverse ? L - 1 - s : s; int tp = reverse ? t + 1 : t - 1;
        const real* hp = s == 0 ? zero : y + (size_t)tp * out; const real* cp = s == 0 ? zero : cseq + (size_t)tp * out;
        for (int i = 0; i < out; i++) dh[i] += dy[(size_t)t * out + i];
        orc_lstm_step_backward(P, dP, in, out, peep, x + (size_t)t * in, hp, cp, cseq + (size_t)t * out, acts + (size_t)t * 4 * out,
                               dh, dc, dx + (size_t)t * in, dhp, dcp);
        memcpy(dh, dhp, sizeof(real) * out); memcpy(dc, dcp, sizeof(real) * out); }
    free(zero); free(dh); free(dc); free(dhp); free(dcp);
}

/* ------------------------------------------------------------------------------------------
 * nn.Attention decoder, teacher forced (Attention.lua:39-211, RNNAttention.lua:144-253)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int L, T;
    real *Vh;      /* [L,S]   Attention.lua:44 */
    real *alpha;   /* [T,L]   alpha_t            */
    real *s;       /* [T,ST]  s_t                */
    real *c;       /* [T,A]   c_t                */
    real *q;       /* [T,S]   ws output          */
    real *cin, *yin, *u;   /* [T,ST] each */
    real *gz, *gr, *gh;    /* [T,ST] decoder GRU gates */
    real *mpre;    /* [T,2,M*MW] maxout pre-activations (stage 1 | stage 2 when MLP == 2) */
    int  *midx;    /* [T,2,M]   argmax within window */
    real *mo;      /* [T,3,M]   maxout 1 | linear | maxout 2 */
    real *logp;    /* [T,V]   */
    real *F;       /* [T,L,K] location features (K>0) */
    real *pen;     /* [T] monotonic penalty value */
} dec_state;

static dec_state* dec_alloc(const layout_t* Y, int L, int T) {
    dec_state* d = (dec_state*)calloc(1, sizeof(dec_state)); d->L = L; d->T = T;
#define AL(p, n) d->p = (real*)calloc((size_t)(n) > 0 ? (size_t)(n) : 1, sizeof(real))
    AL(Vh, (size_t)L * Y->S); AL(alpha, (size_t)T * L); AL(s, (size_t)T * Y->ST); AL(c, (size_t)T * Y->A); AL(q, (size_t)T * Y->S);
    AL(cin, (size_t)T * Y->ST); AL(yin, (size_t)T * Y->ST); AL(u, (size_t)T * Y->ST);
    AL(gz, (size_t)T * Y->ST); AL(gr, (size_t)T * Y->ST); AL(gh, (size_t)T * Y->ST);
    AL(mpre, (size_t)T * 2 * Y->M * Y->MW); AL(mo, (size_t)T * 3 * Y->M); AL(logp, (size_t)T * Y->V);
    AL(F, (size_t)T * L * Y->K); AL(pen, T);
#undef AL
    d->midx = (int*)calloc((size_t)T * 2 * Y->M, sizeof(int));
    return d;
}
static void dec_free(dec_state* d) {
    free(d->Vh); free(d->alpha); free(d->s); free(d->c); free(d->q); free(d->cin); free(d->yin); free(d->u);
    free(d->gz); free(d->gr); free(d->gh); free(d->mpre); free(d->mo); free(d->logp); free(d->F); free(d->pen); free(d->midx); free(d);
}

static void pad_lr(int kf, int* pl, int* pr) {                   /* Attention.lua:77-85 */
    if (kf % 2 == 1) { *pl = (kf - 1) / 2; *pr = *pl; } else { *pl = kf / 2; *pr = *pl - 1; }
}

/* One decoder step forward (decoder_base_, Attention.lua:51-182).
 * prev alpha/s may be NULL (=zeros, Recurrent.lua:112); yprev < 0 means zeros_y (RNNAttention.lua:173).
 * dropmask: NULL or [ST+A] multiplicative mask on [s;c] before Maxout (model_chorowski_baseline_dropout.lua:56). */
static void dec_step_fwd(const layout_t* Y, const real* P, real lambda, const real* h, int L, const real* Vh,
                         const real* ap, const real* sp, int yprev, const real* dropmask,
                         real* alpha, real* s, real* c, real* q, real* F, real* cin, real* yin, real* u,
                         real* gz, real* gr, real* gh, real* mpre, int* midx, real* mo, real* logp, real* pen) {
    const int S = Y->S, A = Y->A, ST = Y->ST, V = Y->V, K = Y->K, KF = Y->KF, M = Y->M, MW = Y->MW;
    real* zs = (real*)calloc(ST > L ? ST : L, sizeof(real));
    if (!sp) sp = zs;
    /* Ws: TemporalConvolution(1,S,ST) on View(ST,1)(prev_s) == Linear ST->S with bias (Attention.lua:65-66) */
    for (int i = 0; i < S; i++) q[i] = P[Y->bs.off + i] + dotr(P + Y->Ws.off + (size_t)i * ST, sp, ST);
    /* location features (Attention.lua:86-91) */
    int pl = 0, pr = 0;
    if (K > 0) {
        pad_lr(KF, &pl, &pr);
        for (int l = 0; l < L; l++) for (int m = 0; m < K; m++) {
            real v = P[Y->bF.off + m];
            for (int j = 0; j < KF; j++) { int i = l + j - pl; if (i >= 0 && i < L && ap) v += P[Y->WF.off + (size_t)m * KF + j] * ap[i]; }
            F[(size_t)l * K + m] = v;
        }
    }
    /* Z = Ws + Vh (+ UF) ; e = w . tanh(Z) ; softmax  (Attention.lua:95-117) */
    real* e = (real*)malloc(sizeof(real) * L); real* zrow = (real*)malloc(sizeof(real) * S);
    const real* w = P + Y->we.off;
    for (int l = 0; l < L; l++) {
        for (int i = 0; i < S; i++) zrow[i] = q[i] + Vh[(size_t)l * S + i];
        if (K > 0) for (int i = 0; i < S; i++) zrow[i] += dotr(P + Y->U.off + (size_t)i * K, F + (size_t)l * K, K);
        real acc = 0; for (int i = 0; i < S; i++) acc += w[i] * tanhr(zrow[i]);
        e[l] = acc;
    }
    real mx = e[0]; for (int l = 1; l < L; l++) if (e[l] > mx) mx = e[l];
    real sum = 0; for (int l = 0; l < L; l++) { alpha[l] = (real)exp((double)(e[l] - mx)); sum += alpha[l]; }
    for (int l = 0; l < L; l++) alpha[l] /= sum;
    /* MonotonicAlignment fwd: identity, records penalty (MonotonicAlignment.lua:27-41) */
    { real ca = 0, cp = 0, tot = 0; for (int l = 0; l < L; l++) { ca += alpha[l]; cp += ap ? ap[l] : 0; tot += ca - cp; }
      *pen = lambda * (tot > 0 ? tot : 0); }
    /* context (Attention.lua:132-134) */
    for (int j = 0; j < A; j++) c[j] = 0;
    for (int l = 0; l < L; l++) { const real a = alpha[l]; const real* hl = h + (size_t)l * A; for (int j = 0; j < A; j++) c[j] += a * hl[j]; }
    /* recurrent input (Attention.lua:149-151) */
    for (int i = 0; i < ST; i++) yin[i] = P[Y->by.off + i] + (yprev >= 0 ? P[Y->Wy.off + (size_t)i * V + yprev] : 0);
    for (int i = 0; i < ST; i++) cin[i] = P[Y->bc.off + i] + dotr(P + Y->Wc.off + (size_t)i * A, c, A);
    for (int i = 0; i < ST; i++) u[i] = P[Y->bj.off + i] + dotr(P + Y->Wj.off + (size_t)i * 2 * ST, cin, ST) + dotr(P + Y->Wj.off + (size_t)i * 2 * ST + ST, yin, ST);
    /* decoder GRU (model_chorowski_baseline.lua:50; mem = Identity(prev_mem), :51) */
    gru_w g = { P + Y->Gz.off, P + Y->Gr.off, P + Y->Gh.off, 0, 0, 0, ST, ST };
    real* rh = (real*)malloc(sizeof(real) * ST);
    gru_step_fwd(&g, u, sp, s, gz, gr, gh, rh); free(rh);
    /* decoder MLP: JoinTable{s,c} -> [Dropout] -> Maxout -> Linear -> LogSoftMax (model:53-59) */
    real* sc = (real*)malloc(sizeof(real) * (ST + A));
    memcpy(sc, s, sizeof(real) * ST); memcpy(sc + ST, c, sizeof(real) * A);
    if (dropmask) for (int i = 0; i < ST + A; i++) sc[i] *= dropmask[i];
    for (int i = 0; i < M * MW; i++) mpre[i] = P[Y->bm.off + i] + dotr(P + Y->Wm.off + (size_t)i * (ST + A), sc, ST + A);
    for (int i = 0; i < M; i++) { int b = 0; for (int j = 1; j < MW; j++) if (mpre[i * MW + j] > mpre[i * MW + b]) b = j;   /* Maxout.lua:16-18 */
        midx[i] = b; mo[i] = mpre[i * MW + b]; }
    const real* top = mo;
    if (Y->MLP == 2) {                                             /* Linear(M,M) -> Maxout(M,M,7): model_vgg.lua:78-79 */
        real* l1 = mo + M; real* mo2 = mo + 2 * M; real* mpre2 = mpre + M * MW; int* midx2 = midx + M;
        for (int i = 0; i < M; i++) l1[i] = P[Y->bl.off + i] + dotr(P + Y->Wl.off + (size_t)i * M, mo, M);
        for (int i = 0; i < M * MW; i++) mpre2[i] = P[Y->bm2.off + i] + dotr(P + Y->Wm2.off + (size_t)i * M, l1, M);
        for (int i = 0; i < M; i++) { int b = 0; for (int j = 1; j < MW; j++) if (mpre2[i * MW + j] > mpre2[i * MW + b]) b = j;
            midx2[i] = b; mo2[i] = mpre2[i * MW + b]; }
        top = mo2;
    }
    real lmx = -INFINITY;
    for (int i = 0; i < V; i++) { logp[i] = P[Y->bo.off + i] + dotr(P + Y->Wo.off + (size_t)i * M, top, M); if (logp[i] > lmx) lmx = logp[i]; }
    double lse = 0; for (int i = 0; i < V; i++) lse += exp((double)(logp[i] - lmx));
    real lz = lmx + (real)log(lse);
    for (int i = 0; i < V; i++) logp[i] -= lz;
    free(e); free(zrow); free(sc); free(zs);
}

/* Teacher-forced forward over T steps.  labels[t] are 0-based class ids; the label fed at step t is
 * labels[t-1] (RNNAttention.lua:172-176).  dropmask NULL or [T, ST+A]. */
static void dec_forward(const layout_t* Y, const real* P, real lambda, const real* h, int L, const int* labels, int T,
                        const real* dropmask, dec_state* d) {
    const int S = Y->S, A = Y->A, ST = Y->ST;
    /* Vh = TemporalConvolutionZeroBias(A,S,1)(h)  (Attention.lua:44; bias forced to zero, TCZB.lua:38) */
    for (int l = 0; l < L; l++) gemv(P + Y->WV.off, A, S, A, h + (size_t)l * A, d->Vh + (size_t)l * S, 0);
    for (int t = 0; t < T; t++) {
        dec_step_fwd(Y, P, lambda, h, L, d->Vh, t ? d->alpha + (size_t)(t - 1) * L : NULL, t ? d->s + (size_t)(t - 1) * ST : NULL,
                     t ? labels[t - 1] : -1, dropmask ? dropmask + (size_t)t * (ST + A) : NULL,
                     d->alpha + (size_t)t * L, d->s + (size_t)t * ST, d->c + (size_t)t * A, d->q + (size_t)t * S,
                     d->F + (size_t)t * L * Y->K, d->cin + (size_t)t * ST, d->yin + (size_t)t * ST, d->u + (size_t)t * ST,
                     d->gz + (size_t)t * ST, d->gr + (size_t)t * ST, d->gh + (size_t)t * ST,
                     d->mpre + (size_t)t * 2 * Y->M * Y->MW, d->midx + (size_t)t * 2 * Y->M, d->mo + (size_t)t * 3 * Y->M, d->logp + (size_t)t * Y->V,
                     d->pen + t);
    }
}

/* Backward through the teacher-forced decoder.  dlogp [T,V]; G = flat gradient (accumulated);
 * dh [L,A] (overwritten with the gradient w.r.t. the annotations, incl. the Vh path). */
static void dec_backward(const layout_t* Y, const real* P, real* G, real lambda, const real* h, int L, const int* labels, int T,
                         const real* dropmask, const dec_state* d, const real* dlogp, real* dh) {
    const int S = Y->S, A = Y->A, ST = Y->ST, V = Y->V, K = Y->K, KF = Y->KF, M = Y->M, MW = Y->MW;
    real* dVh = (real*)calloc((size_t)L * S, sizeof(real));
    real* dalpha = (real*)calloc(L, sizeof(real));      /* carry: grad wrt alpha_t from step t+1 */
    real* dalpha_prev = (real*)malloc(sizeof(real) * L);
    real* ds = (real*)calloc(ST, sizeof(real));          /* carry: grad wrt s_t from step t+1 */
    real* ds_prev = (real*)malloc(sizeof(real) * ST);
    real* dlog = (real*)malloc(sizeof(real) * V); real* dmo = (real*)malloc(sizeof(real) * M); real* dm = (real*)malloc(sizeof(real) * M * MW);
    real* sc = (real*)malloc(sizeof(real) * (ST + A)); real* dsc = (real*)malloc(sizeof(real) * (ST + A));
    real* dc = (real*)malloc(sizeof(real) * A); real* du = (real*)malloc(sizeof(real) * ST); real* scr = (real*)malloc(sizeof(real) * 4 * ST);
    real* dcy = (real*)malloc(sizeof(real) * 2 * ST); real* cy = (real*)malloc(sizeof(real) * 2 * ST);
    real* de = (real*)malloc(sizeof(real) * L); real* dq = (real*)malloc(sizeof(real) * S); real* zrow = (real*)malloc(sizeof(real) * S);
    real* dF = (real*)malloc(sizeof(real) * ((size_t)L * K + 1)); real* zeros = (real*)calloc(ST > L ? ST : L, sizeof(real));
    memset(dh, 0, sizeof(real) * (size_t)L * A);
    gru_w g = { P + Y->Gz.off, P + Y->Gr.off, P + Y->Gh.off, G + Y->Gz.off, G + Y->Gr.off, G + Y->Gh.off, ST, ST };
    int pl = 0, pr = 0; if (K > 0) pad_lr(KF, &pl, &pr);
    const real* w = P + Y->we.off;

    for (int t = T - 1; t >= 0; t--) {                                  /* RNNAttention.lua:233 */
        const real* al = d->alpha + (size_t)t * L; const real* st = d->s + (size_t)t * ST; const real* ct = d->c + (size_t)t * A;
        const real* ap = t ? d->alpha + (size_t)(t - 1) * L : NULL; const real* sp = t ? d->s + (size_t)(t - 1) * ST : zeros;
        const real* dropm = dropmask ? dropmask + (size_t)t * (ST + A) : NULL;
        /* LogSoftMax bwd */
        real gs = 0; for (int i = 0; i < V; i++) gs += dlogp[(size_t)t * V + i];
        for (int i = 0; i < V; i++) dlog[i] = dlogp[(size_t)t * V + i] - (real)exp((double)d->logp[(size_t)t * V + i]) * gs;
        /* Linear M->V */
        const real* mo_t = d->mo + (size_t)t * 3 * M; const int* midx_t = d->midx + (size_t)t * 2 * M;
        ger(G + Y->Wo.off, M, V, M, dlog, Y->MLP == 2 ? mo_t + 2 * M : mo_t); for (int i = 0; i < V; i++) G[Y->bo.off + i] += dlog[i];
        gemv_t(P + Y->Wo.off, M, V, M, dlog, dmo, 0);
        if (Y->MLP == 2) {                                          /* Maxout 2 and Linear(M,M) backward (model_vgg.lua:78-79) */
            memset(dm, 0, sizeof(real) * M * MW);
            for (int i = 0; i < M; i++) dm[i * MW + midx_t[M + i]] = dmo[i];
            ger(G + Y->Wm2.off, M, M * MW, M, dm, mo_t + M); for (int i = 0; i < M * MW; i++) G[Y->bm2.off + i] += dm[i];
            real* dl1 = (real*)malloc(sizeof(real) * M);
            gemv_t(P + Y->Wm2.off, M, M * MW, M, dm, dl1, 0);
            ger(G + Y->Wl.off, M, M, M, dl1, mo_t); for (int i = 0; i < M; i++) G[Y->bl.off + i] += dl1[i];
            gemv_t(P + Y->Wl.off, M, M, M, dl1, dmo, 0);
            free(dl1);
        }
        /* Maxout bwd (gradient to the arg-max unit of each window) */
        memset(dm, 0, sizeof(real) * M * MW);
        for (int i = 0; i < M; i++) dm[i * MW + midx_t[i]] = dmo[i];
        memcpy(sc, st, sizeof(real) * ST); memcpy(sc + ST, ct, sizeof(real) * A);
        if (dropm) for (int i = 0; i < ST + A; i++) sc[i] *= dropm[i];
        ger(G + Y->Wm.off, ST + A, M * MW, ST + A, dm, sc); for (int i = 0; i < M * MW; i++) G[Y->bm.off + i] += dm[i];
        gemv_t(P + Y->Wm.off, ST + A, M * MW, ST + A, dm, dsc, 0);
        if (dropm) for (int i = 0; i < ST + A; i++) dsc[i] *= dropm[i];
        for (int i = 0; i < ST; i++) ds[i] += dsc[i];
        memcpy(dc, dsc + ST, sizeof(real) * A);
        /* decoder GRU bwd: ds (= dL/ds_t) -> du, ds_prev */
        gru_step_bwd(&g, d->u + (size_t)t * ST, sp, d->gz + (size_t)t * ST, d->gr + (size_t)t * ST, d->gh + (size_t)t * ST, ds, du, ds_prev, scr);
        /* Linear(2ST->ST) on JoinTable{c_in, y_in} (Attention.lua:151) */
        memcpy(cy, d->cin + (size_t)t * ST, sizeof(real) * ST); memcpy(cy + ST, d->yin + (size_t)t * ST, sizeof(real) * ST);
        ger(G + Y->Wj.off, 2 * ST, ST, 2 * ST, du, cy); for (int i = 0; i < ST; i++) G[Y->bj.off + i] += du[i];
        gemv_t(P + Y->Wj.off, 2 * ST, ST, 2 * ST, du, dcy, 0);
        /* c_in = Linear(A->ST)(c) ; y_in = Linear(V->ST)(prev_y) */
        ger(G + Y->Wc.off, A, ST, A, dcy, ct); for (int i = 0; i < ST; i++) G[Y->bc.off + i] += dcy[i];
        gemv_t(P + Y->Wc.off, A, ST, A, dcy, dc, 1);
        for (int i = 0; i < ST; i++) { G[Y->by.off + i] += dcy[ST + i]; if (t) G[Y->Wy.off + (size_t)i * V + labels[t - 1]] += dcy[ST + i]; }
        /* context bwd (Attention.lua:132-134): dalpha += h dc ; dh += alpha dc^T */
        for (int l = 0; l < L; l++) { dalpha[l] += dotr(h + (size_t)l * A, dc, A); const real a = al[l]; real* dhl = dh + (size_t)l * A;
            for (int j = 0; j < A; j++) dhl[j] += a * dc[j]; }
        /* MonotonicAlignment bwd (MonotonicAlignment.lua:49-75) */
        memset(dalpha_prev, 0, sizeof(real) * L);
        if (lambda != 0 && d->pen[t] > 0)
            for (int l = 0; l < L; l++) { real gd = lambda * (real)(L - l); dalpha[l] += gd; dalpha_prev[l] -= gd; }   /* (L+1 - l_1based) */
        /* SoftMax bwd */
        real dot = dotr(al, dalpha, L);
        for (int l = 0; l < L; l++) de[l] = al[l] * (dalpha[l] - dot);
        /* e / tanh / CAddTable bwd; recompute tanh(Z) */
        memset(dq, 0, sizeof(real) * S);
        const real* q = d->q + (size_t)t * S; const real* F = d->F + (size_t)t * L * K;
        for (int l = 0; l < L; l++) {
            for (int i = 0; i < S; i++) zrow[i] = q[i] + d->Vh[(size_t)l * S + i];
            if (K > 0) for (int i = 0; i < S; i++) zrow[i] += dotr(P + Y->U.off + (size_t)i * K, F + (size_t)l * K, K);
            real* dvl = dVh + (size_t)l * S; const real del = de[l];
            if (K > 0) for (int m = 0; m < K; m++) dF[(size_t)l * K + m] = 0;
            for (int i = 0; i < S; i++) {
                real th = tanhr(zrow[i]);
                G[Y->we.off + i] += del * th;                                  /* TCZB accGradParameters; gradBias zeroed (TCZB.lua:52) */
                real dz = del * w[i] * ((real)1 - th * th);
                dvl[i] += dz; dq[i] += dz;                                     /* RNNAttention.lua:247 ; ExpandAs.lua:32-37 */
                if (K > 0) for (int m = 0; m < K; m++) { G[Y->U.off + (size_t)i * K + m] += dz * F[(size_t)l * K + m]; dF[(size_t)l * K + m] += dz * P[Y->U.off + (size_t)i * K + m]; }
            }
        }
        if (K > 0) {
            for (int l = 0; l < L; l++) for (int m = 0; m < K; m++) { const real gfm = dF[(size_t)l * K + m];
                G[Y->bF.off + m] += gfm;
                for (int j = 0; j < KF; j++) { int i = l + j - pl; if (i >= 0 && i < L) {
                    if (ap) G[Y->WF.off + (size_t)m * KF + j] += gfm * ap[i];
                    dalpha_prev[i] += gfm * P[Y->WF.off + (size_t)m * KF + j]; } } }
        }
        /* Ws bwd */
        ger(G + Y->Ws.off, ST, S, ST, dq, sp); for (int i = 0; i < S; i++) G[Y->bs.off + i] += dq[i];
        gemv_t(P + Y->Ws.off, ST, S, ST, dq, ds_prev, 1);
        /* carries to step t-1 */
        memcpy(ds, ds_prev, sizeof(real) * ST); memcpy(dalpha, dalpha_prev, sizeof(real) * L);
    }
    /* Vh bwd (TCZB.lua:42-54): dW_V += dVh^T h ; dh += dVh W_V */
    for (int l = 0; l < L; l++) { ger(G + Y->WV.off, A, S, A, dVh + (size_t)l * S, h + (size_t)l * A);
        gemv_t(P + Y->WV.off, A, S, A, dVh + (size_t)l * S, dh + (size_t)l * A, 1); }
    free(dVh); free(dalpha); free(dalpha_prev); free(ds); free(ds_prev); free(dlog); free(dmo); free(dm); free(sc); free(dsc); free(dc);
    free(du); free(scr); free(dcy); free(cy); free(de); free(dq); free(zrow); free(dF); free(zeros);
}

/* ------------------------------------------------------------------------------------------
 * Whole model, one utterance (timit/timit.lua:262-282): encoder -> decoder -> nll -> backward.
 * X [L,D]; labels[T] (0-based); G accumulated.  Optional outputs may be NULL.
 * flags bit0 = normalizeNLL (timit.lua:268-272), bit1 = normalizeGrad (:279-281), bit2 = skip backward.
 * ------------------------------------------------------------------------------------------ */
static real utt_fwdbwd(const layout_t* Y, const real* P, real* G, real lambda, const real* X, int L, const int* labels, int T,
                       const real* dropmask, int flags, real* logp_out, real* alpha_out, real* annot_out, real* dX_out) {
    const int H = Y->H, A = Y->A, NL = Y->NL, V = Y->V;
    real* acts[9]; real* gates[8][2];
    acts[0] = (real*)X;
    for (int l = 0; l < NL; l++) {
        int din = l == 0 ? Y->D : A;
        acts[l + 1] = (real*)calloc((size_t)L * A, sizeof(real));
        for (int d = 0; d < 2; d++) {
            gates[l][d] = (real*)malloc(sizeof(real) * (size_t)L * 3 * H);
            gru_w g = { P + Y->enc[l][d][0].off, P + Y->enc[l][d][1].off, P + Y->enc[l][d][2].off, 0, 0, 0, din, H };
            gru_seq_fwd(&g, acts[l], din, L, d, acts[l + 1] + d * H, A, gates[l][d]);      /* JoinTable(2,2){fwd,rev}: model:24 */
        }
    }
    const real* h = acts[NL];
    dec_state* ds = dec_alloc(Y, L, T);
    dec_forward(Y, P, lambda, h, L, labels, T, dropmask, ds);
    real nll = 0;
    for (int t = 0; t < T; t++) nll -= ds->logp[(size_t)t * V + labels[t]];                /* timit.lua:269-271 */
    if (flags & 1) nll /= (real)T;
    if (logp_out) memcpy(logp_out, ds->logp, sizeof(real) * (size_t)T * V);
    if (alpha_out) memcpy(alpha_out, ds->alpha, sizeof(real) * (size_t)T * L);
    if (annot_out) memcpy(annot_out, h, sizeof(real) * (size_t)L * A);
    if (!(flags & 4)) {
        real* dlogp = (real*)calloc((size_t)T * V, sizeof(real));
        for (int t = 0; t < T; t++) dlogp[(size_t)t * V + labels[t]] = (flags & 2) ? -(real)1 / (real)T : -(real)1;   /* timit.lua:278-281 */
        real* dcur = (real*)malloc(sizeof(real) * (size_t)L * A);
        dec_backward(Y, P, G, lambda, h, L, labels, T, dropmask, ds, dlogp, dcur);
        for (int l = NL - 1; l >= 0; l--) {
            int din = l == 0 ? Y->D : A;
            real* dprev = (real*)calloc((size_t)L * din, sizeof(real));
            for (int d = 0; d < 2; d++) {
                gru_w g = { P + Y->enc[l][d][0].off, P + Y->enc[l][d][1].off, P + Y->enc[l][d][2].off,
                            G + Y->enc[l][d][0].off, G + Y->enc[l][d][1].off, G + Y->enc[l][d][2].off, din, H };
                gru_seq_bwd(&g, acts[l], din, L, d, acts[l + 1] + d * H, A, gates[l][d], dcur + d * H, A, dprev, din);
            }
            free(dcur); dcur = dprev;
        }
        if (dX_out) memcpy(dX_out, dcur, sizeof(real) * (size_t)L * Y->D);
        free(dcur); free(dlogp);
    }
    dec_free(ds);
    for (int l = 0; l < NL; l++) { free(acts[l + 1]); free(gates[l][0]); free(gates[l][1]); }
    return nll;
}

/* Minibatch exactly like timit/timit.lua:233-289: zero grads, loop utterances one at a time,
 * accumulate gradients; nll_out[b] = per-utterance NLL.  X [B, Lmax, D] padded, lengths[B];
 * labels [B, Tmax] padded, tlens[B].  Utterances are distributed over OpenMP threads with
 * thread-private gradient buffers summed at the end (the reference is single-threaded; this
 * only changes the floating-point summation order over utterances).
 * Optional per-utterance outputs: logp [B,Tmax,V], alpha [B,Tmax,Lmax], annot [B,Lmax,A].
 */
EXPORT void orc_model_fwdbwd(const int* cfg, const real* P, real* G, real lambda, const real* X, const int* lengths, int B, int Lmax,
                             const int* labels, const int* tlens, int Tmax, const real* dropmask, int flags, int nthreads,
                             real* nll_out, real* logp, real* alpha, real* annot, real* dX) {
    layout_t Y; make_layout(cfg, &Y);
    memset(G, 0, sizeof(real) * Y.n);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B;
    real** Gt = (real**)malloc(sizeof(real*) * nthreads);
    Gt[0] = G; for (int i = 1; i < nthreads; i++) Gt[i] = (real*)calloc(Y.n, sizeof(real));
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int b = 0; b < B; b++) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        int L = lengths ? lengths[b] : Lmax, T = tlens ? tlens[b] : Tmax;
        real* lp = logp ? (real*)malloc(sizeof(real) * (size_t)T * Y.V) : NULL;
        real* al = alpha ? (real*)malloc(sizeof(real) * (size_t)T * L) : NULL;
        nll_out[b] = utt_fwdbwd(&Y, P, Gt[tid], lambda, X + (size_t)b * Lmax * Y.D, L, labels + (size_t)b * Tmax, T,
                                dropmask ? dropmask + (size_t)b * Tmax * (Y.ST + Y.A) : NULL, flags, lp, al,
                                annot ? annot + (size_t)b * Lmax * Y.A : NULL, dX ? dX + (size_t)b * Lmax * Y.D : NULL);
        if (lp) { memcpy(logp + (size_t)b * Tmax * Y.V, lp, sizeof(real) * (size_t)T * Y.V); free(lp); }
        if (al) { for (int t = 0; t < T; t++) memcpy(alpha + ((size_t)b * Tmax + t) * Lmax, al + (size_t)t * L, sizeof(real) * L); free(al); }
    }
    for (int i = 1; i < nthreads; i++) { for (int64_t j = 0; j < Y.n; j++) G[j] += Gt[i][j]; free(Gt[i]); }
    free(Gt);
}

/* Stand-alone nn.Attention forward/backward on given annotations h [L,A] (Attention.lua:305-327).
 * Returns everything the introspection API exposes: logp [T,V], alpha [T,L] (Attention.lua:241),
 * Ws/q [T,S] (:248), Vh [L,S], penalty [T]. */
EXPORT void orc_attention_forward(const int* cfg, const real* P, real lambda, const real* h, int L, const int* labels, int T,
                                  const real* dropmask, real* logp, real* alpha, real* s, real* c, real* q, real* Vh, real* pen) {
    layout_t Y; make_layout(cfg, &Y);
    dec_state* d = dec_alloc(&Y, L, T);
    dec_forward(&Y, P, lambda, h, L, labels, T, dropmask, d);
    if (logp) memcpy(logp, d->logp, sizeof(real) * (size_t)T * Y.V);
    if (alpha) memcpy(alpha, d->alpha, sizeof(real) * (size_t)T * L);
    if (s) memcpy(s, d->s, sizeof(real) * (size_t)T * Y.ST);
    if (c) memcpy(c, d->c, sizeof(real) * (size_t)T * Y.A);
    if (q) memcpy(q, d->q, sizeof(real) * (size_t)T * Y.S);
    if (Vh) memcpy(Vh, d->Vh, sizeof(real) * (size_t)L * Y.S);
    if (pen) memcpy(pen, d->pen, sizeof(real) * T);
    dec_free(d);
}
EXPORT void orc_attention_backward(const int* cfg, const real* P, real* G, real lambda, const real* h, int L, const int* labels, int T,
                                   const real* dropmask, const real* dlogp, real* dh) {
    layout_t Y; make_layout(cfg, &Y);
    dec_state* d = dec_alloc(&Y, L, T);
    dec_forward(&Y, P, lambda, h, L, labels, T, dropmask, d);
    dec_backward(&Y, P, G, lambda, h, L, labels, T, dropmask, d, dlogp, dh);
    dec_free(d);
}

/* ------------------------------------------------------------------------------------------
 * Attention:BeamSearch (Attention.lua:332-438).  Returns the label sequence length; out[] 0-based.
 * Tie-break: lowest index first (torch.topk order is unverifiable in-tree).
 * ------------------------------------------------------------------------------------------ */
typedef struct { real* alpha; real* s; int* y; int ny; real p; } beam_t;
static void topk_desc(const real* v, int n, int k, int* idx) {
    for (int a = 0; a < k; a++) { int best = -1;
        for (int i = 0; i < n; i++) { int used = 0; for (int b = 0; b < a; b++) if (idx[b] == i) used = 1;
            if (!used && (best < 0 || v[i] > v[best])) best = i; }
        idx[a] = best; }
}
EXPORT int orc_beam_search(const int* cfg, const real* P, const real* h, int L, int eos, int Kb, int maxlen, int* out, real* out_logp) {
    layout_t Y; make_layout(cfg, &Y);
    const int S = Y.S, A = Y.A, ST = Y.ST, V = Y.V;
    real* Vh = (real*)malloc(sizeof(real) * (size_t)L * S);
    for (int l = 0; l < L; l++) gemv(P + Y.WV.off, A, S, A, h + (size_t)l * A, Vh + (size_t)l * S, 0);       /* Attention.lua:355 */
    real *c = (real*)malloc(sizeof(real) * A), *q = (real*)malloc(sizeof(real) * S), *F = (real*)malloc(sizeof(real) * ((size_t)L * Y.K + 1));
    real *cin = (real*)malloc(sizeof(real) * ST), *yin = (real*)malloc(sizeof(real) * ST), *u = (real*)malloc(sizeof(real) * ST);
    real *gz = (real*)malloc(sizeof(real) * ST), *gr = (real*)malloc(sizeof(real) * ST), *gh = (real*)malloc(sizeof(real) * ST);
    real *mpre = (real*)malloc(sizeof(real) * 2 * Y.M * Y.MW), *mo = (real*)malloc(sizeof(real) * 3 * Y.M); int* midx = (int*)malloc(sizeof(int) * 2 * Y.M);
    real pen;
    int cap = maxlen + 2;
    beam_t* beams = (beam_t*)calloc(Kb, sizeof(beam_t)); beam_t* nxt = (beam_t*)calloc(Kb, sizeof(beam_t));
    beam_t* fin = (beam_t*)calloc(Kb, sizeof(beam_t)); int nfin = 0, nb = 0;
    real* lp0 = (real*)malloc(sizeof(real) * V); real* a0 = (real*)malloc(sizeof(real) * L); real* s0 = (real*)malloc(sizeof(real) * ST);
    dec_step_fwd(&Y, P, 0, h, L, Vh, NULL, NULL, -1, NULL, a0, s0, c, q, F, cin, yin, u, gz, gr, gh, mpre, midx, mo, lp0, &pen);   /* :366 */
    int* idx = (int*)malloc(sizeof(int) * Kb);
    int K0 = Kb < V ? Kb : V;
    topk_desc(lp0, V, K0, idx);                                                             /* :370 */
    for (int k = 0; k < K0; k++) {
        beam_t b; b.y = (int*)malloc(sizeof(int) * cap); b.y[0] = idx[k]; b.ny = 1; b.p = lp0[idx[k]];
        b.alpha = (real*)malloc(sizeof(real) * L); b.s = (real*)malloc(sizeof(real) * ST);
        memcpy(b.alpha, a0, sizeof(real) * L); memcpy(b.s, s0, sizeof(real) * ST);
        if (idx[k] == eos) fin[nfin++] = b; else beams[nb++] = b;                            /* :377-387 */
    }
    int count = 0;
    real* pn = (real*)malloc(sizeof(real) * (size_t)Kb * V); real* an = (real*)malloc(sizeof(real) * (size_t)Kb * L); real* sn = (real*)malloc(sizeof(real) * (size_t)Kb * ST);
    while (nfin < K0 && count < maxlen) {                                                    /* :390 */
        count++;
        for (int k = 0; k < nb; k++) {                                                       /* :394-405 */
            dec_step_fwd(&Y, P, 0, h, L, Vh, beams[k].alpha, beams[k].s, beams[k].y[beams[k].ny - 1], NULL,
                         an + (size_t)k * L, sn + (size_t)k * ST, c, q, F, cin, yin, u, gz, gr, gh, mpre, midx, mo, pn + (size_t)k * V, &pen);
            for (int j = 0; j < V; j++) pn[(size_t)k * V + j] += beams[k].p;
        }
        int want = K0 - nfin; int kk = want < nb * V ? want : nb * V;
        topk_desc(pn, nb * V, kk, idx);                                                      /* :406 (topk of K, first K-finished used :413) */
        int nn = 0;
        for (int k = 0; k < kk; k++) {
            int i = idx[k] / V, j = idx[k] % V;                                              /* :407-408 */
            beam_t b; b.y = (int*)malloc(sizeof(int) * cap); memcpy(b.y, beams[i].y, sizeof(int) * beams[i].ny);
            b.y[beams[i].ny] = j; b.ny = beams[i].ny + 1; b.p = pn[(size_t)i * V + j];
            b.alpha = (real*)malloc(sizeof(real) * L); b.s = (real*)malloc(sizeof(real) * ST);
            memcpy(b.alpha, an + (size_t)i * L, sizeof(real) * L); memcpy(b.s, sn + (size_t)i * ST, sizeof(real) * ST);
            if (j == eos || count == maxlen) fin[nfin++] = b; else nxt[nn++] = b;            /* :418-426 */
        }
        for (int k = 0; k < nb; k++) { free(beams[k].y); free(beams[k].alpha); free(beams[k].s); }
        memcpy(beams, nxt, sizeof(beam_t) * nn); nb = nn;
    }
    int best = 0; for (int k = 1; k < nfin; k++) if (fin[k].p > fin[best].p) best = k;      /* :435 */
    int n = 0;
    if (nfin > 0) { n = fin[best].ny; memcpy(out, fin[best].y, sizeof(int) * n); if (out_logp) *out_logp = fin[best].p; }
    for (int k = 0; k < nb; k++) { free(beams[k].y); free(beams[k].alpha); free(beams[k].s); }
    for (int k = 0; k < nfin; k++) { free(fin[k].y); free(fin[k].alpha); free(fin[k].s); }
    free(beams); free(nxt); free(fin); free(idx); free(pn); free(an); free(sn); free(lp0); free(a0); free(s0);
    free(Vh); free(c); free(q); free(F); free(cin); free(yin); free(u); free(gz); free(gr); free(gh); free(mpre); free(mo); free(midx);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Weight noise (WeightNoise.lua:17-22) and adaptive weight noise (AdaptiveWeightNoise.lua:27-104).
 * eps_noise is an INJECTED standard-normal buffer (torch.randn cannot be reproduced).
 * ------------------------------------------------------------------------------------------ */
EXPORT void orc_weightnoise_sample(const real* w, const real* eps_noise, real sigma, int64_t n, real* sample) {
    for (int64_t i = 0; i < n; i++) sample[i] = eps_noise[i] * sigma + w[i];
}
/* weight = [mu ; s=log sigma^2] (2n) */
EXPORT void orc_awn_sample(const real* weight, const real* eps_noise, int64_t n, real* sample) {
    for (int64_t i = 0; i < n; i++) sample[i] = eps_noise[i] * (real)sqrt(exp((double)weight[n + i])) + weight[i];
}
static void awn_stats(const real* weight, int64_t n, double* amu, double* as2, double* sumsq, double* sums2, double* sums) {
    double m = 0; for (int64_t i = 0; i < n; i++) m += weight[i]; m /= (double)n;
    double q = 0, s2 = 0, ss = 0;
    for (int64_t i = 0; i < n; i++) { double d = weight[i] - m; q += d * d; s2 += exp((double)weight[n + i]); ss += weight[n + i]; }
    double a = s2 / (double)n + q / (double)n; if (a < 1e-12) a = 1e-12;       /* AdaptiveWeightNoise.lua:3,71 */
    *amu = m; *as2 = a; *sumsq = q; *sums2 = s2; *sums = ss;
}
/* returns L = lambda*KL + nll (AdaptiveWeightNoise.lua:63-80) */
EXPORT double orc_awn_forward(const real* weight, int64_t n, double lambda, double nll) {
    if (!(lambda > 0)) return nll;
    double amu, as2, q, s2, ss; awn_stats(weight, n, &amu, &as2, &q, &s2, &ss);
    double KL = 0.5 * ((double)n * log(as2) - ss);
    KL += 0.5 / as2 * q;
    KL += 0.5 / as2 * s2 - (double)n / 2;
    return lambda * KL + nll;
}
/* gradWeight [2n] overwritten (AdaptiveWeightNoise.lua:82-104); g = dNLL/dw [n] */
EXPORT void orc_awn_accgrad(const real* weight, const real* g, int64_t n, double lambda, real* gradWeight) {
    double amu = 0, as2 = 1, q, s2, ss;
    if (lambda > 0) awn_stats(weight, n, &amu, &as2, &q, &s2, &ss);
    for (int64_t i = 0; i < n; i++) {
        double sig2 = exp((double)weight[n + i]);
        double dLNds = 0.5 * (double)g[i] * (double)g[i] * sig2;
        if (lambda > 0) {
            gradWeight[i] = (real)(lambda * (weight[i] - amu) / as2 + g[i]);
            gradWeight[n + i] = (real)(lambda * 0.5 / as2 * sig2 - lambda * 0.5 + dLNds);
        } else { gradWeight[i] = g[i]; gradWeight[n + i] = (real)dLNds; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Gradient step (timit/timit.lua:291-315 + optim.adadelta [un-vendored Torch7 `optim`; published
 * algorithm: v = rho v + (1-rho) g^2 ; d = sqrt(a+eps)/sqrt(v+eps) g ; x -= d ; a = rho a + (1-rho) d^2])
 * noise: injected N(0,1) buffer or NULL.  Returns the pre-clip gradient norm (timit.lua:298).
 * ------------------------------------------------------------------------------------------ */
EXPORT double orc_grad_finalize(real* g, const real* p, int64_t n, int batch, double maxnorm, double wd, const real* noise, double noise_sigma) {
    if (batch > 1) for (int64_t i = 0; i < n; i++) g[i] /= (real)batch;                    /* timit.lua:292-295 */
    double nrm = 0; for (int64_t i = 0; i < n; i++) nrm += (double)g[i] * g[i]; nrm = sqrt(nrm);
    if (nrm > maxnorm) { real sc = (real)(maxnorm / nrm); for (int64_t i = 0; i < n; i++) g[i] *= sc; }   /* :300-302 */
    if (wd > 0) for (int64_t i = 0; i < n; i++) g[i] += (real)wd * p[i];                   /* :305-308 */
    if (noise) for (int64_t i = 0; i < n; i++) g[i] += noise[i] * (real)noise_sigma;       /* :311-315 */
    return nrm;
}
EXPORT void orc_adadelta(real* x, const real* g, real* v, real* a, int64_t n, double rho, double eps) {
    for (int64_t i = 0; i < n; i++) {
        v[i] = (real)rho * v[i] + (real)(1 - rho) * g[i] * g[i];
        real d = (real)sqrt((double)a[i] + eps) / (real)sqrt((double)v[i] + eps) * g[i];
        x[i] -= d;
        a[i] = (real)rho * a[i] + (real)(1 - rho) * d * d;
    }
}
/* TrainUtils.columnNormConstraint (TrainUtils.lua:63-85): per-ROW L2 norm (Q11); rows with
 * norm+1e-8 >= maxval are divided by (norm+1e-8)/maxval.  Returns -1 on NaN (reference error()s). */
EXPORT int orc_rownorm_constraint(real* W, int64_t rows, int64_t cols, double maxval) {
    for (int64_t i = 0; i < rows; i++) {
        double nr = 0; for (int64_t j = 0; j < cols; j++) nr += (double)W[i * cols + j] * W[i * cols + j];
        if (nr != nr) return -1;
        real norm = (real)sqrt(nr) + (real)1e-8;
        if (norm >= (real)maxval) { real div = norm / (real)maxval; for (int64_t j = 0; j < cols; j++) W[i * cols + j] /= div; }
    }
    return 0;
}

EXPORT int orc_sizeof_real(void) { return (int)sizeof(real); }
EXPORT int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
