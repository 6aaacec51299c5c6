// model.cuh -- whole-model forward / backward (encoder stack -> nn.Attention -> NLL)
#pragma once
#include "common.cuh"
#include "decoder.cuh"

namespace s2s {

struct ModelState {
    bool valid = false;
    int B = 0, Lmax = 0, T = 0;
    Layout Y;
    const float* acts[9] = {nullptr};   // acts[0] = X (caller), acts[l+1] = output of encoder layer l [B,Lmax,A]
    float* saves[8] = {nullptr};        // per layer [B,Lmax,2,4H]
    float* logp = nullptr;              // [B,T,V] when the caller did not ask for it
};

int model_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* X, const int* lengths, int B, int Lmax,
                  const int* labels, const int* tlens, int T, const float* dropmask, float lambda, int flags, float* nll, float* logp,
                  bool backward_follows = false);
int model_backward(s2s_ctx* ctx, const Layout& Y, const float* P, float* G, const float* X, const int* lengths, int B, int Lmax,
                   const int* labels, const int* tlens, int T, const float* dropmask, float lambda, int flags, float* dX);

}  // namespace s2s
