-- s2s_ffi.lua -- LuaJIT FFI declarations of libs2s_b200.so and small helpers.
-- This is the binding a maintainer of the reference adds; the module shims in this directory (Attention.lua, RNN.lua, ...)
-- re-register the reference's torch classes on top of it.  The cdef block below is GENERATED from include/s2s_b200.h
-- (tests/test_cabi_cpu.py checks that every function the header declares appears here with the same parameter count).
-- NOT EXERCISED IN THE BUILD IMAGE: neither LuaJIT nor Torch7 is installed there; the same call sequences are exercised
-- through the ctypes mirror (seq2seq-attention-asr_b200/ops.py, nn.py).
local ffi = require 'ffi'

ffi.cdef[[
typedef struct s2s_ctx s2s_ctx;
typedef struct s2s_model_cfg {
    int D;
    int H;
    int NL;
    int S;
    int ST;
    int V;
    int K;
    int KF;
    int M;
    int MW;
    int MLP;
} s2s_model_cfg;
int  s2s_ctx_create(int device, void* stream, s2s_ctx** out);
int  s2s_ctx_destroy(s2s_ctx* ctx);
int  s2s_ctx_set_stream(s2s_ctx* ctx, void* stream);
int  s2s_ctx_synchronize(s2s_ctx* ctx);
const char* s2s_last_error(void);
int  s2s_version(void);
int64_t s2s_ctx_launch_count(s2s_ctx* ctx);
int64_t s2s_ctx_kernel_count(s2s_ctx* ctx, int kernel_class);
int  s2s_ctx_set_graphs(s2s_ctx* ctx, int enable);
int  s2s_graph_begin(s2s_ctx* ctx);
int  s2s_graph_end(s2s_ctx* ctx, int* graph_id);
int  s2s_graph_launch(s2s_ctx* ctx, int graph_id);
int  s2s_graph_destroy(s2s_ctx* ctx, int graph_id);
int  s2s_ctx_profile(s2s_ctx* ctx, int enable);
int  s2s_ctx_profile_read(s2s_ctx* ctx, double* ms_host, int64_t* count_host, double* work_host);
int s2s_orthogonalize(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, float* bias);
int  s2s_dp_available(void);
int  s2s_dp_unique_id(void* id_host_128);
int  s2s_dp_init(s2s_ctx* ctx, int rank, int world, const void* id_host_128);
int  s2s_dp_rank(s2s_ctx* ctx);
int  s2s_dp_world(s2s_ctx* ctx);
int  s2s_dp_allreduce(s2s_ctx* ctx, float* G, int64_t n);
int  s2s_dp_broadcast(s2s_ctx* ctx, float* P, int64_t n, int root);
int  s2s_dp_set_overlap(s2s_ctx* ctx, int enable);
int  s2s_dp_destroy(s2s_ctx* ctx);
int64_t s2s_param_count(const s2s_model_cfg* cfg);
int     s2s_param_segments(const s2s_model_cfg* cfg, int64_t* out_host, int max);
int64_t s2s_decoder_param_offset(const s2s_model_cfg* cfg);
int s2s_tconv_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y);
int s2s_tconv_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out,
                          const float* dy, float* dx, float* dW, float scale);
int s2s_linear_zb_forward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out, float* y);
int s2s_linear_zb_backward(s2s_ctx* ctx, const float* x, int64_t rows, int in, const float* W, int out,
                           const float* dy, float* dx, float* dW, float scale);
int64_t s2s_gru_seq_save_floats(int B, int Lmax, int H, int ndir);
int s2s_gru_seq_forward(s2s_ctx* ctx, const float* W, int Din, int H, int ndir, int reverse,
                        const float* x, int ldx, const int* lengths, int B, int Lmax,
                        float* y, float* save);
int s2s_gru_seq_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, int ndir, int reverse,
                         const float* x, int ldx, const int* lengths, int B, int Lmax,
                         const float* y, const float* save, const float* dy, float* dx);
int s2s_gru_step_forward(s2s_ctx* ctx, const float* W, int Din, int H, const float* x, const float* hprev, int B,
                         float* hnext, float* gates);
int s2s_gru_step_backward(s2s_ctx* ctx, const float* W, float* dW, int Din, int H, const float* x, const float* hprev, int B,
                          const float* gates, const float* dhnext, float* dx, float* dhprev);
int s2s_lstm_step_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, const float* x, const float* hprev,
                          const float* cprev, int B, float* hnext, float* cnext, float* acts);
int s2s_lstm_step_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, const float* x,
                           const float* hprev, const float* cprev, int B, const float* acts, const float* cnext,
                           const float* dhnext, const float* dcnext, float* dx, float* dhprev, float* dcprev);
int64_t s2s_lstm_param_count(int in, int out, int peepholes);
int64_t s2s_lstm_seq_save_floats(int B, int Lmax, int H);
int s2s_lstm_seq_forward(s2s_ctx* ctx, const float* P, int Din, int H, int peepholes, int reverse,
                         const float* x, int ldx, const int* lengths, int B, int Lmax, float* y, float* save);
int s2s_lstm_seq_backward(s2s_ctx* ctx, const float* P, float* dP, int Din, int H, int peepholes, int reverse,
                          const float* x, int ldx, const int* lengths, int B, int Lmax,
                          const float* y, const float* save, const float* dy, float* dx);
int s2s_attention_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                          const float* h, const int* lengths, int B, int Lmax,
                          const int* labels, const int* tlens, int Tmax,
                          const float* dropmask, float lambda, float* logp);
int s2s_attention_backward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G,
                           const float* h, const int* lengths, int B, int Lmax,
                           const int* labels, const int* tlens, int Tmax,
                           const float* dropmask, float lambda, const float* dlogp, float* dh);
int s2s_attention_get(s2s_ctx* ctx, int what, float* dst);
int s2s_attention_step(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                       const float* h, const float* Vh, const int* lengths, int B, int Lmax,
                       const int* yprev, const float* alpha_prev, const float* s_prev,
                       float* alpha, float* s, float* logp);
int s2s_beam_search(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, const float* h, int L,
                    int eos, int beam, int maxlen, int* out_host, int* n_out_host, float* logp_out_host);
int s2s_model_forward(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P,
                      const float* X, const int* lengths, int B, int Lmax,
                      const int* labels, const int* tlens, int Tmax,
                      const float* dropmask, float lambda, int flags, float* nll, float* logp);
int s2s_model_fwdbwd(s2s_ctx* ctx, const s2s_model_cfg* cfg, const float* P, float* G,
                     const float* X, const int* lengths, int B, int Lmax,
                     const int* labels, const int* tlens, int Tmax,
                     const float* dropmask, float lambda, int flags,
                     float* nll, float* logp, float* dX);
int s2s_model_get_annotations(s2s_ctx* ctx, float* dst);
int s2s_labels_from_onehot(s2s_ctx* ctx, const float* onehot, int64_t rows, int V, int* labels);
int s2s_onehot(s2s_ctx* ctx, const int* labels, int64_t rows, int V, float* onehot);
int s2s_nll_grad_seed(s2s_ctx* ctx, const float* logp, const int* labels, const int* tlens, int B, int T, int V,
                      int flags, float* nll, float* dlogp);
int s2s_weightnoise_sample(s2s_ctx* ctx, const float* w, const float* eps, uint64_t seed, float sigma, int64_t n, float* sample);
int s2s_awn_sample(s2s_ctx* ctx, const float* weight, const float* eps, uint64_t seed, int64_t n, float* sample);
int s2s_awn_forward(s2s_ctx* ctx, const float* weight, int64_t n, double lambda, double nll, double* L_host);
int s2s_awn_accgrad(s2s_ctx* ctx, const float* weight, const float* g, int64_t n, double lambda, float* gradWeight);
int s2s_dropout_mask(s2s_ctx* ctx, float p, uint64_t seed, int64_t n, float* mask);
typedef struct s2s_vgg_cfg { int C1, C2, HID, OUT; } s2s_vgg_cfg;
int64_t s2s_vgg_param_count(const s2s_vgg_cfg* cfg, int F);
int s2s_vgg_out_len(int T);
int s2s_vgg_forward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, const float* X, int B, int T, int F, float* h);
int s2s_vgg_backward(s2s_ctx* ctx, const s2s_vgg_cfg* cfg, const float* P, float* dP, int B, int T, int F,
                     const float* dh, float* dX);
int s2s_grad_finalize(s2s_ctx* ctx, float* g, const float* p, int64_t n, int batch, double maxnorm, double wd,
                      const float* noise, uint64_t seed, double noise_sigma, double* gradnorm_host);
int s2s_adadelta(s2s_ctx* ctx, float* x, const float* g, float* v, float* a, int64_t n, double rho, double eps);
int s2s_rownorm_constraint(s2s_ctx* ctx, float* W, int64_t rows, int64_t cols, double maxval, int* nan_host);
int s2s_model_rownorm_constraint(s2s_ctx* ctx, const s2s_model_cfg* cfg, float* P, double maxval, int* nan_host);
int s2s_edit_distance(const int* a, int na, const int* b, int nb, int* dist_host);
int s2s_gemm_f32(s2s_ctx* ctx, int impl, int tA, int tB, int M, int N, int K, float alpha,
                 const float* A, int lda, const float* B, int ldb, float beta, float* C, int ldc, const float* bias);
int s2s_attn_step_forward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                          const int* lengths, int B, int Lmax, int S, int A, float* alpha, float* c);
int s2s_attn_step_backward(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                           const int* lengths, int B, int Lmax, int S, int A, const float* alpha,
                           const float* dc, const float* dalpha_in, float* dq, float* de);
int s2s_conv3_forward(s2s_ctx* ctx, const float* in, int64_t Mg, int Ww, int C, const float* Wp, const float* bias,
                      int N, float* out, int relu);
int s2s_conv3_dgrad(s2s_ctx* ctx, const float* dout, int64_t Mg, int Ww, int N, const float* WpT, int C, float* din);
int s2s_conv3_wgrad(s2s_ctx* ctx, const float* dout, const float* in, int64_t Mg, int Ww, int N, int C, float* dWp);
int s2s_attn_step_forward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                              const int* lengths, int B, int Lmax, int S, int A, int KF, const float* uw,
                              const float* alpha_prev, float* alpha, float* c);
int s2s_attn_step_backward_loc(s2s_ctx* ctx, const float* Vh, const float* h, const float* q, const float* w,
                               const int* lengths, int B, int Lmax, int S, int A, int KF, const float* uw,
                               const float* alpha_prev, const float* alpha, const float* dc, const float* dalpha_in,
                               float* dq, float* de, float* dalpha_prev);
]]

local M = {}
M.C = ffi.load(os.getenv('S2S_B200_LIB') or 'libs2s_b200.so')
M.GET_ALPHA, M.GET_WS, M.GET_VH, M.GET_PENALTY, M.GET_STATE, M.GET_CONTEXT = 0, 1, 2, 3, 4, 5
M.NORMALIZE_NLL, M.NORMALIZE_GRAD = 1, 2

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

-- raw device pointer of a contiguous torch.CudaTensor
function M.fptr(t)
   if t == nil then return nil end
   assert(torch.type(t) == 'torch.CudaTensor', 's2s: expected a torch.CudaTensor, got ' .. torch.type(t))
   assert(t:isContiguous(), 's2s: tensor must be contiguous')
   return ffi.cast('float*', t:data())
end

-- int32 device buffers.  cutorch builds of the reference's era have no CudaIntTensor, so an int buffer is a torch.CudaTensor
-- used as raw 4-byte slots: it is only ever written by the library (s2s_labels_from_onehot) and read back by the library.
function M.ibuffer(n) return torch.CudaTensor(n) end
function M.iptr(t)
   if t == nil then return nil end
   assert(torch.type(t) == 'torch.CudaTensor' and t:isContiguous(), 's2s: int buffers are contiguous torch.CudaTensor slots')
   return ffi.cast('int*', t:data())
end

-- one-hot labelmask [T,V] / [B,T,V] (timit/timit.lua:262) -> int labels on the device, no host round trip
function M.labels_of(onehot, buf)
   local V = onehot:size(onehot:nDimension())
   local rows = onehot:nElement() / V
   buf = buf or M.ibuffer(rows)
   if buf:nElement() ~= rows then buf:resize(rows) end
   M.check(M.C.s2s_labels_from_onehot(M.ctx(), M.fptr(onehot:contiguous()), rows, V, M.iptr(buf)))
   return buf
end

-- model.* fields of timit/model_chorowski_baseline.lua:14-46 -> s2s_model_cfg
function M.cfg(model)
   local c = ffi.new('s2s_model_cfg')
   c.D = model.inputFrameSize; c.H = model.hiddenFrameSize; c.NL = model.numEncoderLayers or 3
   c.S = model.scoreDepth; c.ST = model.stateDepth; c.V = model.outputDepth
   c.K = model.hybridAttendFeatureMaps or 0; c.KF = model.hybridAttendFilterSize or 10
   c.M = model.mlpDepth or 64; c.MW = model.maxoutWindow or 7; c.MLP = model.mlpStages or 1
   return c
end

-- leaf wrapper that gives a slice of a flat parameter vector the .weight / .bias / .gradWeight / .gradBias fields
-- TrainUtils.apply2graph looks for (TrainUtils.lua:137-184); a real nn.Module so listModules / findModules keep working
local Param, parent = torch.class('nn.S2SParam', 'nn.Module')
function Param:__init(name, weight, gradWeight, bias, gradBias)
   parent.__init(self)
   self.name, self.weight, self.gradWeight, self.bias, self.gradBias = name, weight, gradWeight, bias, gradBias
end
function Param:parameters()
   if self.bias then return {self.weight, self.bias}, {self.gradWeight, self.gradBias} end
   return {self.weight}, {self.gradWeight}
end
function Param:__tostring__() return 'nn.S2SParam(' .. self.name .. ')' end

return M
