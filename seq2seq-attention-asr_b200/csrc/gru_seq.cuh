// gru_seq.cuh -- persistent cluster GRU sequence kernels (gru_seq.cu)
#pragma once
#include "common.cuh"

namespace s2s {


// y [B,Lmax,ndir*H]; save [B,Lmax,ndir,4H] (z | r | h~ | r*h_prev)
int gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                    const int* lengths, int B, int Lmax, float* y, float* save);
int gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx,
                     const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx, bool defer_wgrad = false);
// joins the weight-gradient branch of gru_seq_backward(defer_wgrad = true) calls back into the context's stream
int gru_seq_wgrad_join(s2s_ctx* ctx);
int zero_tail_rows(s2s_ctx* ctx, float* x, const int* lengths, int B, int Lmax, int W);

}  // namespace s2s
