// gru_seq.cuh -- persistent cluster GRU sequence kernels (gru_seq.cu)
#pragma once
#include "common.cuh"

namespace s2s {

struct GruSeqParams {
    const float* W;        // [ndir][3][H][ldw]  (forward: rows used as-is; backward: read transposed)
    int ldw;               // H + Din
    const float* xp;       // [B, Lmax, ndir*3H] time-batched input projections (forward only)
    const int* lengths;
    int B, Lmax, ndir, reverse0;   // reverse0: direction of dir index 0 (ndir == 1 case)
    float* y;              // [B, Lmax, ndir*H]
    float* save;           // [B, Lmax, ndir, 4H]: z | r | h~ | r*h_prev
    // backward
    const float* dy;       // [B, Lmax, ndir*H]
    float* dA;             // [B, Lmax, ndir*3H]: daz | dar | dah   (same column order as xp)
    float* hp_all;         // [B, Lmax, ndir, H]
    long long* clk;        // optional per-phase clock accumulators of CTA 0 (S2S_GRU_PROF)
    int dbg;               // timing experiments only (S2S_GRU_DBG): 1 = skip the DSMEM exchange, 2 = skip the mat-vec loops
};

// second-generation cluster kernels (gru_seq2.cu)
int gru_cluster2_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H);
// third-generation kernels: warp-specialised, software-pipelined sub-batches (gru_seq3.cu)
int gru_cluster3_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H);
// gru_seq4.cu: generation 3 with sub-batches of at most two utterances (three or four independent chains per cluster)
int gru_cluster4_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H);
// gru_seq5.cu: generation 3 with two units per lane over half the K-slice (half the shared-memory reads of the mat-vec)
int gru_cluster5_launch(s2s_ctx* ctx, bool backward, const GruSeqParams& p, int H);

// y [B,Lmax,ndir*H]; save [B,Lmax,ndir,4H] (z | r | h~ | r*h_prev)
int gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                    const int* lengths, int B, int Lmax, float* y, float* save);
int gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                     const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx, bool defer_wgrad = false);
// joins the weight-gradient branch of gru_seq_backward(defer_wgrad = true) calls back into the context's stream
int gru_seq_wgrad_join(s2s_ctx* ctx);
int zero_tail_rows(s2s_ctx* ctx, float* x, const int* lengths, int B, int Lmax, int W);

}  // namespace s2s
