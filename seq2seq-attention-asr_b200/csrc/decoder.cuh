// decoder.cuh -- state of the teacher-forced attention decoder (nn.Attention) kept between
// updateOutput and updateGradInput, plus the entry points used by model.cu / api.cu.
#pragma once
#include "attention.cuh"
#include "common.cuh"

namespace s2s {

struct DecoderState {
    bool valid = false;
    int B = 0, Lmax = 0, T = 0;
    Layout Y;
    float lambda = 0.f;
    bool has_drop = false;
    // saved by forward (persist arena); all [B, T, .] unless noted
    float* Vh = nullptr;      // [B, Lmax, S]            Attention.lua:44
    float* alpha = nullptr;   // [B, T, Lmax]            decoder:alpha()
    float* sc = nullptr;      // [B, T, ST+A]            {s_t, c_t}  (JoinTable order of model_chorowski_baseline.lua:53-54)
    float* q = nullptr;       // [B, T, S]               decoder:Ws()  (includes b_s, and U b_F on the location path)
    float* pen = nullptr;     // [B, T]
    float* cin = nullptr;     // [B, T, ST]
    float* yin = nullptr;     // [B, T, ST]
    float* su = nullptr;      // [B, T, 2ST]             {s_{t-1}, u_t}   (GRU.lua:22 concat order)
    float* rhu = nullptr;     // [B, T, 2ST]             {r * s_{t-1}, u_t}
    float* gates = nullptr;   // [B, T, 3ST]             z | r | h~
    float* mo = nullptr;      // [B, T, M]
    int* midx = nullptr;      // [B, T, M]
    float* l1 = nullptr;      // [B, T, M]   Linear(M,M) output          (MLP == 2, model_vgg.lua:78)
    float* mo2 = nullptr;     // [B, T, M]   second Maxout output
    int* midx2 = nullptr;     // [B, T, M]
    float* scm = nullptr;     // [B, T, ST+A] masked copy (dropout) or == sc
    float* logp = nullptr;    // [B, T, V]
    float* uw = nullptr;      // [KF, S]   U W_F   (location path)
    float* qbias = nullptr;   // [S]       b_s (+ U b_F)
    float* Wjc = nullptr;     // [ST, A]   W_j[:, :ST] . W_c  (the two input Linears of Attention.lua:150-151 folded)
    float* uy = nullptr;      // [B, T, ST]  W_j[:, ST:] y_in + folded biases (per-call scratch, forward only)
    // transposed copies of the recurrent-chain weights for the backward time loop (K-contiguous products), made by decoder_prepare when a
    // backward pass is known to follow (s2s_model_fwdbwd: under the encoder, on the side stream); decoder_backward makes them itself otherwise
    float *GhT = nullptr, *GzrT = nullptr, *WjcT = nullptr, *WsT = nullptr;
    bool tw_valid = false;
    bool prepared = false, prep_pending = false;      // decoder_prepare ran for (prep_B, prep_T, prep_n); its side-stream event ev[4] is not joined yet
    int prep_B = 0, prep_T = 0; int64_t prep_n = 0; uint64_t prep_epoch = 0, prep_epoch_scratch = 0;
    float* V1 = nullptr;      // [B, T, Lmax, KF]  d e_t / d alpha_{t-1} (location path), formed right after the forward loop when a backward
    bool v1_pending = false;  //                   pass is known to follow (decoder_forward(prefetch_v1)): ev[5] on side[1] marks it complete
    AttnScratch att;
};

int decoder_prepare(s2s_ctx* ctx, const Layout& Y, const float* P, const int* labels, int B, int T, bool backward_follows = false);
int decoder_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax,
                    const int* labels, const int* tlens, int T, const float* dropmask, float lambda, float* logp_out, bool prefetch_v1 = false);
int decoder_backward(s2s_ctx* ctx, const Layout& Y, const float* P, float* G, const float* h, const int* lengths, int B, int Lmax,
                     const int* labels, const int* tlens, int T, const float* dropmask, float lambda, const float* dlogp, float* dh,
                     bool defer_wgrad = false);

// decoder_cluster.cu: the time loop of decoder_forward as one persistent cluster kernel (ST = 256, S = A = 512; content or location-aware, KF <= 10)
int decoder_cluster_forward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax, const int* tlens,
                            int T, float lambda, const float* uy, DecoderState& d, bool* handled);
bool decoder_cluster_backward_eligible(const Layout& Y, int Lmax, float lambda);
int decoder_cluster_backward(s2s_ctx* ctx, const Layout& Y, const float* P, const float* h, const int* lengths, int B, int Lmax, int T, float lambda,
                             const DecoderState& d, const float* WsT, const float* GhT, const float* GzrT, const float* WjcT, const float* dsc,
                             const float* V1, float* dA, float* du_all, float* dc_all, float* dq_all, float* de_all, bool* handled);

}  // namespace s2s
