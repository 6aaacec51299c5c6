-- Attention.lua (shim) -- nn.Attention with the reference's constructor and module protocol
-- (Attention.lua:15-24,214-327 of the reference), computing through libs2s_b200.so.
--   nn.Attention(decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
--                stateDepth, annotationDepth, outputDepth, monoAlignPenalty, penalty_lambda)
-- forward({h, y_onehot}) -> logp ; backward({h, y_onehot}, dlogp) -> {dh, nil}; parameter gradients are
-- accumulated inside updateGradInput exactly as the reference does (Attention.lua:325).
-- The module owns ONE flat parameter / gradient tensor in the library's layout (s2s_param_segments);
-- parameters() returns views of it in that order so getParameters() flattens to the same storage.
local s2s = require 's2s_ffi'
local ffi = require 'ffi'

local Attention, parent = torch.class('nn.Attention', 'nn.Module')

function Attention:__init(decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
                          stateDepth, annotationDepth, outputDepth, monoAlignPenalty, penalty_lambda)
   parent.__init(self)
   self.scoreDepth = scoreDepth
   self.hybridAttendFilterSize = hybridAttendFilterSize or 10
   self.hybridAttendFeatureMaps = hybridAttendFeatureMaps or 0
   self.stateDepth = stateDepth
   self.annotationDepth = annotationDepth
   self.outputDepth = outputDepth
   self.penalty_lambda = (monoAlignPenalty and penalty_lambda) or 0
   -- decoder_recurrent / decoder_mlp are accepted for signature compatibility: the GRU(2*st -> st) decoder and
   -- the Maxout(64x7) -> Linear -> LogSoftMax MLP of model_chorowski_baseline.lua:48-59 are built in.
   self.cfg = ffi.new('s2s_model_cfg')
   self.cfg.D = 1; self.cfg.H = annotationDepth / 2; self.cfg.NL = 1        -- encoder fields unused by the decoder calls
   self.cfg.S = scoreDepth; self.cfg.ST = stateDepth; self.cfg.V = outputDepth
   self.cfg.K = self.hybridAttendFeatureMaps; self.cfg.KF = self.hybridAttendFilterSize
   self.cfg.M = (decoder_mlp and decoder_mlp.mlpDepth) or 64; self.cfg.MW = (decoder_mlp and decoder_mlp.window) or 7
   local n = tonumber(s2s.C.s2s_param_count(self.cfg))
   self.off = tonumber(s2s.C.s2s_decoder_param_offset(self.cfg))
   self.flat = torch.CudaTensor(n):zero()          -- whole-model layout; the decoder segments start at self.off
   self.gradFlat = torch.CudaTensor(n):zero()
   self:reset()
end

function Attention:parameters()
   local segs = ffi.new('int64_t[?]', 3 * 128)
   local ns = s2s.C.s2s_param_segments(self.cfg, segs, 128)
   local p, g = {}, {}
   for i = 0, ns - 1 do
      local off, rows, cols = tonumber(segs[3 * i]), tonumber(segs[3 * i + 1]), tonumber(segs[3 * i + 2])
      if off >= self.off then
         p[#p + 1] = self.flat:narrow(1, off + 1, rows * cols):view(rows, cols)
         g[#g + 1] = self.gradFlat:narrow(1, off + 1, rows * cols):view(rows, cols)
      end
   end
   return p, g
end

function Attention:reset(stdv)
   local p = self:parameters()
   for _, w in ipairs(p) do
      local bound = stdv or 1 / math.sqrt(w:size(2) > 1 and w:size(2) or w:size(1))
      w:uniform(-bound, bound)
   end
end

local function shapes(x, y)
   if x:nDimension() == 2 then return 1, x:size(1), y:size(1) end          -- nonbatch ("SGD") mode, Attention.lua:308-311
   if x:nDimension() == 3 then return x:size(1), x:size(2), y:size(2) end
   error('x must be 2d or 3d')                                              -- Attention.lua:316
end

-- one-hot labelmask [T,V] / [B,T,V] -> int labels (argmax)
local function labels_of(y)
   local _, idx = y:max(y:nDimension())
   return idx:add(-1):int():cuda():contiguous()
end

function Attention:updateOutput(input)
   local x, y = unpack(input)
   local B, L, T = shapes(x, y)
   self.labels = labels_of(y)
   self.output = self.output:cuda():resize(y:size())
   s2s.check(s2s.C.s2s_attention_forward(s2s.ctx(), self.cfg, s2s.fptr(self.flat), s2s.fptr(x:contiguous()), nil, B, L,
                                         s2s.iptr(self.labels), nil, T, nil, self.penalty_lambda, s2s.fptr(self.output)))
   self.B, self.L, self.T = B, L, T
   return self.output
end

function Attention:updateGradInput(input, gradOutput)
   local x = input[1]
   self.gradH = self.gradH or torch.CudaTensor()
   self.gradH:resizeAs(x)
   s2s.check(s2s.C.s2s_attention_backward(s2s.ctx(), self.cfg, s2s.fptr(self.flat), s2s.fptr(self.gradFlat), s2s.fptr(x:contiguous()),
                                          nil, self.B, self.L, s2s.iptr(self.labels), nil, self.T, nil, self.penalty_lambda,
                                          s2s.fptr(gradOutput:contiguous()), s2s.fptr(self.gradH)))
   self.gradInput = {self.gradH}           -- the gradient w.r.t. the one-hot labels is not produced (callers discard it)
   return self.gradInput
end

function Attention:accGradParameters() end   -- accumulated inside updateGradInput (reference: Attention.lua:325)

local function getter(self, what, last)
   local out = self.B == 1 and torch.CudaTensor(self.T, last) or torch.CudaTensor(self.B, self.T, last)
   s2s.check(s2s.C.s2s_attention_get(s2s.ctx(), what, s2s.fptr(out)))
   return out
end
function Attention:alpha() return getter(self, s2s.GET_ALPHA, self.L) end         -- Attention.lua:241
function Attention:Ws() return getter(self, s2s.GET_WS, self.scoreDepth) end      -- Attention.lua:248
function Attention:penalty() return getter(self, s2s.GET_PENALTY, 1) end          -- Attention.lua:250
function Attention:setpenalty(penalty) self.penalty_lambda = (opt and opt.penalty) or penalty or self.penalty_lambda end  -- :252-274

-- Attention:BeamSearch(annotations, eos, K, maxseqlength)  (Attention.lua:332-438); eos is 1-based as in Lua
function Attention:BeamSearch(h, eos, K, maxseqlength)
   local L = h:size(1)
   local out = ffi.new('int[?]', maxseqlength + 2)
   local n, lp = ffi.new('int[1]'), ffi.new('float[1]')
   s2s.check(s2s.C.s2s_beam_search(s2s.ctx(), self.cfg, s2s.fptr(self.flat), s2s.fptr(h:contiguous()), L, eos - 1, K, maxseqlength, out, n, lp))
   local y = torch.IntTensor(n[0])
   for i = 0, n[0] - 1 do y[i + 1] = out[i] + 1 end
   return y, lp[0]
end
