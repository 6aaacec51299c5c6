-- s2s_ffi.lua -- LuaJIT FFI declarations of libs2s_b200.so (include/s2s_b200.h) and small helpers.
-- This is the binding a maintainer of the reference adds; the module shims in this directory
-- (Attention.lua, RNN.lua, ...) re-register the reference's torch classes on top of it.
-- NOT EXERCISED IN THE BUILD IMAGE: neither LuaJIT nor Torch7 is installed there; the same call
-- sequences are exercised through the ctypes mirror (seq2seq-attention-asr_b200/ops.py, nn.py).
local ffi = require 'ffi'

ffi.cdef[[
typedef struct s2s_ctx s2s_ctx;
typedef struct s2s_model_cfg { int D, H, NL, S, ST, V, K, KF, M, MW, MLP; } s2s_model_cfg;
int  s2s_ctx_create(int device, void* stream, s2s_ctx** out);
int  s2s_ctx_destroy(s2s_ctx* ctx);
int  s2s_ctx_set_stream(s2s_ctx* ctx, void* stream);
int  s2s_ctx_synchronize(s2s_ctx* ctx);
const char* s2s_last_error(void);
int64_t s2s_param_count(const s2s_model_cfg* cfg);
int     s2s_param_segments(const s2s_model_cfg* cfg, int64_t* out_host, int max);
int64_t s2s_decoder_param_offset(const s2s_model_cfg* cfg);
int s2s_tconv_zb_forward(s2s_ctx*, const float* x, int64_t rows, int in_, const float* W, int out, float* y);
int s2s_tconv_zb_backward(s2s_ctx*, const float* x, int64_t rows, int in_, const float* W, int out, const float* dy, float* dx, float* dW, float scale);
int s2s_linear_zb_forward(s2s_ctx*, const float* x, int64_t rows, int in_, const float* W, int out, float* y);
int s2s_linear_zb_backward(s2s_ctx*, const float* x, int64_t rows, int in_, const float* W, int out, const float* dy, float* dx, float* dW, float scale);
int64_t s2s_gru_seq_save_floats(int B, int Lmax, int H, int ndir);
int s2s_gru_seq_forward(s2s_ctx*, const float* W, int Din, int H, int ndir, int reverse, const float* x, int ldx, const int* lengths, int B, int Lmax, float* y, float* save);
int s2s_gru_seq_backward(s2s_ctx*, const float* W, float* dW, int Din, int H, int ndir, int reverse, const float* x, int ldx, const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx);
int s2s_gru_step_forward(s2s_ctx*, const float* W, int Din, int H, const float* x, const float* hprev, int B, float* hnext, float* gates);
int s2s_gru_step_backward(s2s_ctx*, const float* W, float* dW, int Din, int H, const float* x, const float* hprev, int B, const float* gates, const float* dhnext, float* dx, float* dhprev);
int s2s_dropout_mask(s2s_ctx*, float p, uint64_t seed, int64_t n, float* mask);
int s2s_lstm_step_forward(s2s_ctx*, const float* P, int Din, int H, int peepholes, const float* x, const float* hprev, const float* cprev, int B, float* hnext, float* cnext, float* acts);
int s2s_lstm_step_backward(s2s_ctx*, const float* P, float* dP, int Din, int H, int peepholes, const float* x, const float* hprev, const float* cprev, int B, const float* acts, const float* cnext, const float* dhnext, const float* dcnext, float* dx, float* dhprev, float* dcprev);
int s2s_edit_distance(const int* a, int na, const int* b, int nb, int* dist_host);
int s2s_nll_grad_seed(s2s_ctx*, const float* logp, const int* labels, const int* tlens, int B, int T, int V, int flags, float* nll, float* dlogp);
typedef struct s2s_vgg_cfg { int C1, C2, HID, OUT; } s2s_vgg_cfg;
int64_t s2s_vgg_param_count(const s2s_vgg_cfg* cfg, int F);
int s2s_vgg_out_len(int T);
int s2s_vgg_forward(s2s_ctx*, const s2s_vgg_cfg*, const float* P, const float* X, int B, int T, int F, float* h);
int s2s_vgg_backward(s2s_ctx*, const s2s_vgg_cfg*, const float* P, float* dP, int B, int T, int F, const float* dh, float* dX);
int s2s_attention_forward(s2s_ctx*, const s2s_model_cfg*, const float* P, const float* h, const int* lengths, int B, int Lmax, const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, float* logp);
int s2s_attention_backward(s2s_ctx*, const s2s_model_cfg*, const float* P, float* G, const float* h, const int* lengths, int B, int Lmax, const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, const float* dlogp, float* dh);
int s2s_attention_get(s2s_ctx*, int what, float* dst);
int s2s_attention_step(s2s_ctx*, const s2s_model_cfg*, const float* P, const float* h, const float* Vh, const int* lengths, int B, int Lmax, const int* yprev, const float* alpha_prev, const float* s_prev, float* alpha, float* s, float* logp);
int s2s_beam_search(s2s_ctx*, const s2s_model_cfg*, const float* P, const float* h, int L, int eos, int beam, int maxlen, int* out_host, int* n_out_host, float* logp_out_host);
int s2s_model_forward(s2s_ctx*, const s2s_model_cfg*, const float* P, const float* X, const int* lengths, int B, int Lmax, const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, int flags, float* nll, float* logp);
int s2s_model_fwdbwd(s2s_ctx*, const s2s_model_cfg*, const float* P, float* G, const float* X, const int* lengths, int B, int Lmax, const int* labels, const int* tlens, int Tmax, const float* dropmask, float lambda, int flags, float* nll, float* logp, float* dX);
int s2s_model_get_annotations(s2s_ctx*, float* dst);
int s2s_weightnoise_sample(s2s_ctx*, const float* w, const float* eps, uint64_t seed, float sigma, int64_t n, float* sample);
int s2s_awn_sample(s2s_ctx*, const float* weight, const float* eps, uint64_t seed, int64_t n, float* sample);
int s2s_awn_forward(s2s_ctx*, const float* weight, int64_t n, double lambda, double nll, double* L_host);
int s2s_awn_accgrad(s2s_ctx*, const float* weight, const float* g, int64_t n, double lambda, float* gradWeight);
int s2s_grad_finalize(s2s_ctx*, float* g, const float* p, int64_t n, int batch, double maxnorm, double wd, const float* noise, uint64_t seed, double noise_sigma, double* gradnorm_host);
int s2s_adadelta(s2s_ctx*, float* x, const float* g, float* v, float* a, int64_t n, double rho, double eps);
int s2s_rownorm_constraint(s2s_ctx*, float* W, int64_t rows, int64_t cols, double maxval, int* nan_host);
int s2s_model_rownorm_constraint(s2s_ctx*, const s2s_model_cfg*, float* P, double maxval, int* nan_host);
]]

local M = {}
M.C = ffi.load(os.getenv('S2S_B200_LIB') or 'libs2s_b200.so')
M.GET_ALPHA, M.GET_WS, M.GET_VH, M.GET_PENALTY, M.GET_STATE, M.GET_CONTEXT = 0, 1, 2, 3, 4, 5

-- one context per process, bound to cutorch's current device and its default stream
-- (timit/timit.lua:39 cutorch.setDevice(opt.device))
local ctxp = ffi.new('s2s_ctx*[1]')
function M.ctx()
   if M._ctx == nil then
      local dev = (cutorch and cutorch.getDevice() or 1) - 1
      M.check(M.C.s2s_ctx_create(dev, nil, ctxp))
      M._ctx = ctxp[0]
   end
   return M._ctx
end

-- Lua error() with the library's message, like the reference's assert/error paths
function M.check(rc)
   if rc ~= 0 then error(ffi.string(M.C.s2s_last_error()), 2) end
end

-- raw device pointer of a contiguous torch.CudaTensor / torch.CudaIntTensor
function M.fptr(t)
   if t == nil then return nil end
   assert(t:isContiguous(), 's2s: tensor must be contiguous')
   return ffi.cast('float*', t:data())
end
function M.iptr(t)
   if t == nil then return nil end
   assert(t:isContiguous(), 's2s: tensor must be contiguous')
   return ffi.cast('int*', t:data())
end

-- model.* fields of timit/model_chorowski_baseline.lua:14-46 -> s2s_model_cfg
function M.cfg(model)
   local c = ffi.new('s2s_model_cfg')
   c.D = model.inputFrameSize; c.H = model.hiddenFrameSize; c.NL = model.numEncoderLayers or 3
   c.S = model.scoreDepth; c.ST = model.stateDepth; c.V = model.outputDepth
   c.K = model.hybridAttendFeatureMaps or 0; c.KF = model.hybridAttendFilterSize or 10
   c.M = model.mlpDepth or 64; c.MW = model.maxoutWindow or 7
   return c
end

return M
