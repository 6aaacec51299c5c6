-- Attention.lua (shim) -- nn.Attention with the reference's constructor and module protocol (Attention.lua:15-24,214-438 of the
-- reference), computing through libs2s_b200.so.
--   nn.Attention(decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
--                stateDepth, annotationDepth, outputDepth, monoAlignPenalty, penalty_lambda)
-- forward({h, y_onehot}) -> logp ; backward({h, y_onehot}, dlogp) -> {dh, dy = 0}; parameter gradients are accumulated inside
-- updateGradInput exactly as the reference does (Attention.lua:325).
-- The module owns ONE flat parameter / gradient tensor in the library's decoder-only layout (cfg.NL = 0, s2s_param_segments);
-- parameters() returns the SAME view tensors every time, in that order, so getParameters() re-points them into the caller's flat
-- storage and the library is always handed the first view's pointer (the views stay contiguous under nn.Module.flatten).
--
-- What is read from the two sub-graphs the model file builds (model_chorowski_baseline.lua:48-59, _dropout.lua:56, model_vgg.lua:70-80):
--   decoder_recurrent : must contain one nn.GRU(stateDepth, stateDepth); its weights initialise G_z, G_r, G_h
--   decoder_mlp       : nn.Maxout stages (one or two), an optional nn.Dropout(p) in front, nn.Linear layers; their sizes select
--                       cfg.M / cfg.MW / cfg.MLP, their weights initialise W_m, b_m, (W_l, b_l, W_m2, b_m2,) W_o, b_o
-- Anything else in those graphs (an LSTM decoder, a ReLU MLP: the inline model of timit/timit.lua:97-169) is rejected loudly.
require 'RNNAttention'
require 'Recurrent'
local s2s = require 's2s_ffi'
local ffi = require 'ffi'

local Attention, parent = torch.class('nn.Attention', 'nn.Module')

local SEGS = {'WV', 'bV', 'Ws', 'bs', 'WF', 'bF', 'U', 'bU', 'we', 'be', 'Wy', 'by', 'Wc', 'bc', 'Wj', 'bj', 'Gz', 'Gr', 'Gh', 'Wm', 'bm',
              'Wl', 'bl', 'Wm2', 'bm2', 'Wo', 'bo'}
local function segment_names(cfg)            -- order of make_layout (csrc/core.cu), decoder part
   local names = {}
   for _, n in ipairs(SEGS) do
      local loc = (n == 'WF' or n == 'bF' or n == 'U' or n == 'bU')
      local mlp2 = (n == 'Wl' or n == 'bl' or n == 'Wm2' or n == 'bm2')
      if (not loc or cfg.K > 0) and (not mlp2 or cfg.MLP == 2) then names[#names + 1] = n end
   end
   return names
end

local function find(graph, typename) return graph and graph.findModules and graph:findModules(typename) or {} end

function Attention:__init(decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
                          stateDepth, annotationDepth, outputDepth, monoAlignPenalty, penalty_lambda)
   parent.__init(self)
   self.scoreDepth = scoreDepth
   self.hybridAttendFilterSize = hybridAttendFilterSize or 10              -- Attention.lua:17
   self.hybridAttendFeatureMaps = hybridAttendFeatureMaps or 0
   self.stateDepth, self.annotationDepth, self.outputDepth = stateDepth, annotationDepth, outputDepth
   self.monoAlignPenalty = monoAlignPenalty
   self.penalty_lambda = (monoAlignPenalty and penalty_lambda) or 0       -- Attention.lua:122-126
   -- ---- inspect the caller's sub-graphs -------------------------------------------------------------------------------------
   local grus = find(decoder_recurrent, 'nn.GRU')
   assert(#grus == 1 and #find(decoder_recurrent, 'nn.LSTM') == 0,
          'nn.Attention (libs2s_b200): decoder_recurrent must wrap exactly one nn.GRU (LSTM decoders are not built)')
   assert(grus[1].diminput == stateDepth and grus[1].dimoutput == stateDepth, 'nn.Attention: decoder GRU must be GRU(stateDepth, stateDepth)')
   local maxouts, linears, drops = find(decoder_mlp, 'nn.Maxout'), {}, find(decoder_mlp, 'nn.Dropout')
   assert(#maxouts == 1 or #maxouts == 2, 'nn.Attention (libs2s_b200): decoder_mlp must be Maxout-Linear or Maxout-Linear-Maxout-Linear')
   for _, l in ipairs(find(decoder_mlp, 'nn.Linear')) do          -- the Linears that are NOT the inside of a Maxout
      local inside = false
      for _, m in ipairs(maxouts) do for _, ml in ipairs(m:findModules('nn.Linear')) do inside = inside or ml == l end end
      if not inside then linears[#linears + 1] = l end
   end
   assert(#linears == #maxouts, 'nn.Attention (libs2s_b200): one nn.Linear after every nn.Maxout expected')
   assert(maxouts[1].inputDim == stateDepth + annotationDepth, 'nn.Attention: first Maxout must read {s, c}')
   self.dropout = drops[1]                                         -- nn.Dropout on {s, c}: p and train/evaluate state are read per call
   local cfg = ffi.new('s2s_model_cfg')
   cfg.D = 1; cfg.H = annotationDepth / 2; cfg.NL = 0               -- decoder-only parameter vector
   cfg.S = scoreDepth; cfg.ST = stateDepth; cfg.V = outputDepth
   cfg.K = self.hybridAttendFeatureMaps; cfg.KF = self.hybridAttendFilterSize
   cfg.M = maxouts[1].outputDim; cfg.MW = maxouts[1].window; cfg.MLP = #maxouts
   if #maxouts == 2 then assert(maxouts[2].outputDim == cfg.M and maxouts[2].window == cfg.MW, 'nn.Attention: both Maxout stages must have the same shape') end
   self.cfg = cfg
   -- ---- flat storage and the leaf views -------------------------------------------------------------------------------------
   local n = tonumber(s2s.C.s2s_param_count(cfg))
   self.flat = torch.CudaTensor(n):zero()
   self.gradFlat = torch.CudaTensor(n):zero()
   local segs = ffi.new('int64_t[?]', 3 * 128)
   local ns = s2s.C.s2s_param_segments(cfg, segs, 128)
   local names = segment_names(cfg)
   assert(ns == #names, 'nn.Attention: segment table out of step with the library')
   self._p, self._g, self.seg = {}, {}, {}
   for i = 0, ns - 1 do
      local off, rows, cols = tonumber(segs[3 * i]), tonumber(segs[3 * i + 1]), tonumber(segs[3 * i + 2])
      local function view(t) local v = t:narrow(1, off + 1, rows * cols); if cols > 1 then v = v:view(rows, cols) end; return v end
      self._p[i + 1], self._g[i + 1] = view(self.flat), view(self.gradFlat)
      self.seg[names[i + 1]] = i + 1
   end
   -- leaves for TrainUtils.apply2graph (weight + bias pairs; row-norm constraint and orthogonalize see what the reference's graph shows)
   self.modules = {}
   local function leaf(w, b) self.modules[#self.modules + 1] = nn.S2SParam(w, self._p[self.seg[w]], self._g[self.seg[w]], b and self._p[self.seg[b]], b and self._g[self.seg[b]]) end
   leaf('WV', 'bV'); leaf('Ws', 'bs')
   if cfg.K > 0 then leaf('WF', 'bF'); leaf('U', 'bU') end
   leaf('we', 'be'); leaf('Wy', 'by'); leaf('Wc', 'bc'); leaf('Wj', 'bj'); leaf('Gz'); leaf('Gr'); leaf('Gh'); leaf('Wm', 'bm')
   if cfg.MLP == 2 then leaf('Wl', 'bl'); leaf('Wm2', 'bm2') end
   leaf('Wo', 'bo')
   self:reset()
   -- adopt the values the caller's modules were constructed with (same distribution, and a pre-loaded sub-module keeps its weights)
   local function put(name, t) self._p[self.seg[name]]:copy(t) end
   put('Gz', grus[1].modules[1].weight); put('Gr', grus[1].modules[2].weight); put('Gh', grus[1].modules[3].weight)
   local function maxout_linear(m) return m:findModules('nn.Linear')[1] end
   put('Wm', maxout_linear(maxouts[1]).weight); put('bm', maxout_linear(maxouts[1]).bias)
   if cfg.MLP == 2 then
      put('Wl', linears[1].weight); put('bl', linears[1].bias)
      put('Wm2', maxout_linear(maxouts[2]).weight); put('bm2', maxout_linear(maxouts[2]).bias)
   end
   put('Wo', linears[#linears].weight); put('bo', linears[#linears].bias)
   -- ---- the accessors callers use (timit/timit.lua:521, Attention.lua:354-365, ExtractAlpha.lua) ---------------------------------
   local me = self
   self.Vh = setmetatable({}, {__index = function(_, k) if k == 'output' then return me:_get(s2s.GET_VH) end end})     -- decoder.Vh.output
   self.decoder_base = nn.Recurrent(nn.S2SDecoderStep(self), {0, stateDepth, stateDepth})                             -- Attention.lua:186-188
   self.rnn = nn.RNNAttention(self.decoder_base, outputDepth, false)                                                  -- Attention.lua:203
   self.rnn.zeros_y = torch.CudaTensor(outputDepth):zero()
end

function Attention:parameters() return self._p, self._g end

function Attention:reset(stdv)               -- per-leaf reset() rules: U(+-1/sqrt(fan_in)) or U(+-stdv sqrt 3); ZeroBias biases stay 0
   for _, m in ipairs(self.modules) do
      local w = m.weight
      local fan_in = w:dim() == 2 and w:size(2) or w:size(1)
      if m.name == 'WF' then fan_in = self.cfg.KF end          -- TemporalConvolution(1, K, k): kW * inputFrameSize
      local b = stdv and stdv * math.sqrt(3) or 1 / math.sqrt(fan_in)
      w:uniform(-b, b)
      if m.bias then
         if m.name == 'WV' or m.name == 'U' or m.name == 'we' then m.bias:zero() else m.bias:uniform(-b, b) end   -- TemporalConvolutionZeroBias.lua:34
      end
   end
end

function Attention:training() self.train = true; if self.dropout then self.dropout:training() end end
function Attention:evaluate() self.train = false; if self.dropout then self.dropout:evaluate() end end
function Attention:cuda() return self end
function Attention:float() error('nn.Attention (libs2s_b200): CUDA only, there is no CPU path') end
Attention.double = Attention.float
function Attention:type(t)                    -- autoencoder:cuda() reaches every module through :type (timit/timit.lua:171)
   assert(t == nil or t == 'torch.CudaTensor', 'nn.Attention (libs2s_b200): CUDA only, there is no CPU path')
   return t and self or 'torch.CudaTensor'
end
function Attention:setT(T) self.T = T end

function Attention:_base() return s2s.fptr(self._p[1]), s2s.fptr(self._g[1]) end      -- W_V is the first decoder segment (offset 0 at NL = 0)

local function shapes(x, y)
   if x:nDimension() == 2 then return 1, x:size(1), y:size(1) end          -- nonbatch ("SGD") mode, Attention.lua:308-311
   if x:nDimension() == 3 then return x:size(1), x:size(2), y:size(2) end
   error('x must be 2d or 3d')                                              -- Attention.lua:316
end

function Attention:updateOutput(input)
   local x, y = unpack(input)
   local B, L, T = shapes(x, y)
   self.labels = s2s.labels_of(y, self.labels)
   self.output = torch.type(self.output) == 'torch.CudaTensor' and self.output or torch.CudaTensor()
   self.output:resize(y:size())
   local mask = nil
   if self.dropout and self.train ~= false and self.dropout.p > 0 then      -- nn.Dropout v2: mask / (1-p) in training mode only
      self.dropmask = self.dropmask or torch.CudaTensor()
      self.dropmask:resize(B * T, self.stateDepth + self.annotationDepth)
      self.dropseed = (self.dropseed or 0) + 1
      s2s.check(s2s.C.s2s_dropout_mask(s2s.ctx(), self.dropout.p, self.dropseed, self.dropmask:nElement(), s2s.fptr(self.dropmask)))
      mask = s2s.fptr(self.dropmask)
   end
   self.mask_used = mask
   local P = self:_base()
   s2s.check(s2s.C.s2s_attention_forward(s2s.ctx(), self.cfg, P, s2s.fptr(x:contiguous()), nil, B, L, s2s.iptr(self.labels), nil, T, mask,
                                         self.penalty_lambda, s2s.fptr(self.output)))
   self.B, self.L, self.T = B, L, T
   self.rnn.T, self.rnn.batchSize = T, x:nDimension() == 3 and B or 0
   return self.output
end

function Attention:updateGradInput(input, gradOutput)
   local x, y = unpack(input)
   self.gradH = self.gradH or torch.CudaTensor()
   self.gradH:resizeAs(x)
   local P, G = self:_base()
   s2s.check(s2s.C.s2s_attention_backward(s2s.ctx(), self.cfg, P, G, s2s.fptr(x:contiguous()), nil, self.B, self.L, s2s.iptr(self.labels), nil,
                                          self.T, self.mask_used, self.penalty_lambda, s2s.fptr(gradOutput:contiguous()), s2s.fptr(self.gradH)))
   -- the gradient w.r.t. the one-hot labels (RNNAttention.lua:248) is not produced: every caller discards it; zeros keep nngraph's sums valid
   self.gradY = self.gradY or torch.CudaTensor()
   self.gradY:resizeAs(y):zero()
   self.gradInput = {self.gradH, self.gradY}
   return self.gradInput
end

function Attention:accGradParameters() end   -- accumulated inside updateGradInput (reference: Attention.lua:325)

function Attention:_get(what)
   assert(self.B, 'forward must be run at least once')                      -- Attention.lua:219
   local last = ({[s2s.GET_ALPHA] = self.L, [s2s.GET_WS] = self.scoreDepth, [s2s.GET_PENALTY] = 1})[what]
   local out
   if what == s2s.GET_VH then out = self.rnn.batchSize == 0 and torch.CudaTensor(self.L, self.scoreDepth) or torch.CudaTensor(self.B, self.L, self.scoreDepth)
   else out = self.rnn.batchSize == 0 and torch.CudaTensor(self.T, last) or torch.CudaTensor(self.B, self.T, last) end
   s2s.check(s2s.C.s2s_attention_get(s2s.ctx(), what, s2s.fptr(out)))
   return out
end
function Attention:alpha() return self:_get(s2s.GET_ALPHA) end             -- Attention.lua:241
function Attention:Ws() return self:_get(s2s.GET_WS) end                   -- Attention.lua:248
function Attention:penalty() return self:_get(s2s.GET_PENALTY) end         -- Attention.lua:245
function Attention:setpenalty(penalty)                                     -- Attention.lua:252-274 (reads the global opt.penalty)
   assert(self.monoAlignPenalty, 'could not find penalty node')
   self.penalty_lambda = (opt and opt.penalty) or penalty or self.penalty_lambda
   print('setting penalty to ' .. self.penalty_lambda)
end

-- Attention:BeamSearch(annotations, eos, K, maxseqlength)  (Attention.lua:332-438); eos is 1-based as in Lua
function Attention:BeamSearch(h, eos, K, maxseqlength)
   local L = h:size(1)
   local out = ffi.new('int[?]', maxseqlength + 2)
   local n, lp = ffi.new('int[1]'), ffi.new('float[1]')
   local P = self:_base()
   s2s.check(s2s.C.s2s_beam_search(s2s.ctx(), self.cfg, P, s2s.fptr(h:contiguous()), L, eos - 1, K, maxseqlength, out, n, lp))
   local y = torch.IntTensor(n[0])
   for i = 0, n[0] - 1 do y[i + 1] = out[i] + 1 end
   return y, lp[0]
end

-- ---- decoder_base: ONE decoder step with explicit hidden state {alpha, s, mem}, the prototype BeamSearch drives in the reference
-- (Attention.lua:366,402: decoder_base:forward({{{Vh, h}, prev_y}, hidden}) -> {logp, {alpha, s, mem}}); single utterance or batch
local Step, sparent = torch.class('nn.S2SDecoderStep', 'nn.Module')
function Step:__init(att) sparent.__init(self); self.att = att end
function Step:parameters() return {}, {} end       -- the parameters belong to nn.Attention
function Step:updateOutput(input)
   local inp, hidden = unpack(input)
   local Vh_h, prev_y = unpack(inp)
   local Vh, h = unpack(Vh_h)
   local alpha_prev, s_prev = hidden[1], hidden[2]
   local att = self.att
   local B, L = 1, h:size(1)
   if h:nDimension() == 3 then B, L = h:size(1), h:size(2) end
   local yprev = nil
   if prev_y and prev_y:nElement() > 0 then yprev = s2s.iptr(s2s.labels_of(prev_y)) end      -- all-zero prev_y -> label -1 (t = 1)
   local sz = h:nDimension() == 3 and {B} or {}
   local function new(last) local s = {unpack(sz)}; s[#s + 1] = last; return torch.CudaTensor(unpack(s)) end
   local alpha, s, logp = new(L), new(att.stateDepth), new(att.outputDepth)
   local P = att:_base()
   local ap = (alpha_prev and alpha_prev:nElement() == B * L) and s2s.fptr(alpha_prev:contiguous()) or nil    -- dimhidden {0,..}: empty = zeros
   s2s.check(s2s.C.s2s_attention_step(s2s.ctx(), att.cfg, P, s2s.fptr(h:contiguous()), s2s.fptr(Vh:contiguous()), nil, B, L, yprev, ap,
                                      s_prev and s2s.fptr(s_prev:contiguous()) or nil, s2s.fptr(alpha), s2s.fptr(s), s2s.fptr(logp)))
   self.output = {logp, {alpha, s, hidden[3]}}     -- mem is an identity pass-through for GRU decoders (model_chorowski_baseline.lua:51)
   return self.output
end
function Step:updateGradInput() error('nn.S2SDecoderStep: decode-only; training runs through nn.Attention:backward') end
