// attention.cuh -- launch wrappers of the attention-step kernels (attention.cu)
#pragma once
#include "common.cuh"

namespace s2s {

constexpr int ATT_R = 32;        // encoder frames per CTA (forward / backward step kernels)
constexpr int ATT_THREADS = 256;
constexpr int DVH_R = 16;        // encoder frames per CTA (deferred dVh kernel)

struct AttnScratch {             // sized for (B, Lmax, S, A); lives in the ctx arenas
    int nch = 0;                 // ceil(Lmax / ATT_R)
    float* E = nullptr;          // [B, Lmax]   raw energies of the current step
    float* part_ms = nullptr;    // [B, nch, 2] chunk max / chunk sum
    float* part_c = nullptr;     // [B, nch, A] chunk partial contexts
    float* dalpha = nullptr;     // [B, Lmax]   d alpha of the current backward step
    float* part_P = nullptr;     // [B, nch, 2, S]
    float* part_dot = nullptr;   // [B, nch]
    float* V1 = nullptr;         // [B, Lmax, KF] (location path, backward)
    unsigned* counters = nullptr;   // [B] zero between launches
};
int attn_scratch_alloc(s2s_ctx* ctx, Arena& arena, int B, int Lmax, int S, int A, int KF, bool backward, AttnScratch* sc);

struct AttnLoc {                 // location-aware term (Attention.lua:75-99), folded: Z += UW . alpha_pad + Ub
    int KF = 0, padl = 0;
    const float* uw = nullptr;   // [KF, S]  UW[j][i] = sum_m U[i,m] WF[m,j]
    const float* alpha_prev = nullptr;   // [B, *] previous alignment (NULL = zeros)
    int64_t ld_aprev = 0;
};

// forward step: e = w.tanh(q + Vh (+loc)); alpha = softmax_l(e) over l < L_b; c = sum_l alpha_l h_l
//   (Attention.lua:95-135).  q already includes the bias b_s (and U b_F for the location path).
// pen (nullable): lambda * max(0, sum_l (cumsum alpha_t - cumsum alpha_prev))  (MonotonicAlignment.lua:27-41)
int attn_step_fwd(s2s_ctx* ctx, const AttnScratch& sc, const float* Vh, const float* h, const float* q, int64_t ldq,
                  const float* w, const int* lengths, int B, int Lmax, int S, int A, const AttnLoc& loc,
                  float* alpha, int64_t ld_alpha, float* c, int64_t ld_c, float* pen, int64_t ld_pen, float lambda,
                  const float* alpha_prev_pen, int64_t ld_app, const int* tlens = nullptr, int tstep = 0);

// backward step (single pass over h and Vh):
//   dalpha = h.dc + dalpha_in ; de = alpha (dalpha - <alpha,dalpha>) ; dq = sum_l de_l w (1 - tanh^2 Z_l)
// pen_active (nullable) [B]: monotonic penalty > 0 flags; adds lambda*(L - l) to dalpha (MonotonicAlignment.lua:49-75)
int attn_step_bwd(s2s_ctx* ctx, const AttnScratch& sc, const float* Vh, const float* h, const float* q, int64_t ldq,
                  const float* w, const int* lengths, int B, int Lmax, int S, int A, const AttnLoc& loc,
                  const float* alpha, int64_t ld_alpha, const float* dc, int64_t ld_dc,
                  const float* dalpha_in, int64_t ld_dain, const float* pen, int64_t ld_pen, float lambda,
                  float* dq, int64_t ld_dq, float* de, int64_t ld_de, float* dalpha_prev, int64_t ld_dap);

// deferred accumulation over all T steps (replaces the per-step [L,S] read-modify-write of
// RNNAttention.lua:247):  dVh[b,l,:] = sum_t de_t[b,l] w (1 - tanh^2(q_t[b] + Vh[b,l] (+loc)))
//                         dwe[:]    += sum_{b,t,l} de_t[b,l] tanh(...)
//   q_all [B,T,S], de_all [B,T,Lmax], alpha_all [B,T,Lmax] (location path), tlens nullable.
int attn_dvh(s2s_ctx* ctx, const float* Vh, const float* q_all, const float* de_all, const float* w, const int* lengths,
             const int* tlens, int B, int Lmax, int T, int S, const AttnLoc& loc_all, float* dVh, float* dwe, float* duw);

// location path, all decoder steps at once: V1[b,t,l,j] = sum_s UW[j,s] w_s (1 - tanh^2 Z_t[l,s])  (t >= 1; row t = 0 is not written).
// The alignment carry of the backward time loop is dalpha_{t-1}[i] = sum_j de_t[x] V1[t,x,j], x = i - j + pad_left.
int attn_v1(s2s_ctx* ctx, const float* Vh, const float* q_all, const float* w, const float* uw, const float* alpha_all, const int* lengths,
            const int* tlens, int B, int Lmax, int T, int S, int KF, int padl, float* V1);

}  // namespace s2s
