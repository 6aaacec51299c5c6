-- LSTM.lua (shim) -- nn.LSTM(diminput, dimoutput, peepholes) with the reference's parameter set (LSTM.lua:25-60: per gate
-- Linear(in,out) + Linear(out,out), both with bias; optional full-matrix peepholes Linear(out,out) on prev_c (i, f) / next_c (o)),
-- stored as ONE flat tensor in the library's order (s2s_lstm_param_count) so nn.RNN hands the whole recurrence to
-- s2s_lstm_seq_forward / _backward with one pointer.  Stand-alone step (LSTM.lua:100-136): forward({x, prev_h, prev_c}) -> {h, c}.
local s2s = require 's2s_ffi'

local LSTM, parent = torch.class('nn.LSTM', 'nn.Module')

function LSTM:__init(diminput, dimoutput, peepholes)
   parent.__init(self)
   assert(diminput ~= nil, "diminput must be specified")      -- LSTM.lua:9
   assert(dimoutput ~= nil, "dimoutput must be specified")    -- LSTM.lua:10
   self.diminput, self.dimoutput, self.peepholes = diminput, dimoutput, peepholes or false
   local n = tonumber(s2s.C.s2s_lstm_param_count(diminput, dimoutput, self.peepholes and 1 or 0))
   self.weight = torch.CudaTensor(n)
   self.gradWeight = torch.CudaTensor(n):zero()
   self:reset()
end

function LSTM:reset(stdv)                    -- stock nn.Linear bounds: 1/sqrt(fan_in) of each layer is <= 1/sqrt(min fan-in)
   local b = stdv and stdv * math.sqrt(3) or 1 / math.sqrt(math.max(self.diminput, self.dimoutput))
   self.weight:uniform(-b, b)
end
function LSTM:parameters() return {self.weight}, {self.gradWeight} end
function LSTM:float() error('nn.LSTM (libs2s_b200): CUDA only, there is no CPU path') end
LSTM.double = LSTM.float
function LSTM:type(t)
   assert(t == nil or t == 'torch.CudaTensor', 'nn.LSTM (libs2s_b200): CUDA only, there is no CPU path')
   return t and self or 'torch.CudaTensor'
end
function LSTM:cuda() return self end

local function zeros_like_state(self, x)     -- LSTM.lua:93-99
   local B = x:dim() == 2 and x:size(1) or 1
   if not self.zeros or self.zeros:nElement() ~= B * self.dimoutput then self.zeros = torch.CudaTensor(B, self.dimoutput):zero() end
   return x:dim() == 2 and self.zeros or self.zeros:view(self.dimoutput)
end

function LSTM:updateOutput(input)            -- LSTM.lua:100-116
   local x, prev_h, prev_c = unpack(input)
   local B, H = x:dim() == 2 and x:size(1) or 1, self.dimoutput
   local z = zeros_like_state(self, x)
   prev_h, prev_c = prev_h or z, prev_c or z
   self.h = self.h or torch.CudaTensor(); self.c = self.c or torch.CudaTensor(); self.acts = self.acts or torch.CudaTensor()
   self.h:resizeAs(prev_h); self.c:resizeAs(prev_c); self.acts:resize(B, 4 * H)
   s2s.check(s2s.C.s2s_lstm_step_forward(s2s.ctx(), s2s.fptr(self.weight), self.diminput, H, self.peepholes and 1 or 0, s2s.fptr(x:contiguous()),
                                         s2s.fptr(prev_h:contiguous()), s2s.fptr(prev_c:contiguous()), B, s2s.fptr(self.h), s2s.fptr(self.c), s2s.fptr(self.acts)))
   self.output = {self.h, self.c}
   return self.output
end

function LSTM:updateGradInput(input, gradOutput)   -- LSTM.lua:118-136; weight gradients accumulate here (:133)
   local x, prev_h, prev_c = unpack(input)
   local dEdh, dEdc = unpack(gradOutput)
   assert(dEdh ~= nil, "dEdh should not be nil")
   assert(prev_c ~= nil or dEdc ~= nil, "prev_c and dEdc cannot both be nil")
   local B, H = x:dim() == 2 and x:size(1) or 1, self.dimoutput
   local z = zeros_like_state(self, x)
   prev_h, prev_c = prev_h or z, prev_c or z
   self.gx = self.gx or torch.CudaTensor(); self.gh = self.gh or torch.CudaTensor(); self.gc = self.gc or torch.CudaTensor()
   self.gx:resizeAs(x); self.gh:resizeAs(prev_h); self.gc:resizeAs(prev_c)
   s2s.check(s2s.C.s2s_lstm_step_backward(s2s.ctx(), s2s.fptr(self.weight), s2s.fptr(self.gradWeight), self.diminput, H, self.peepholes and 1 or 0,
                                          s2s.fptr(x:contiguous()), s2s.fptr(prev_h:contiguous()), s2s.fptr(prev_c:contiguous()), B, s2s.fptr(self.acts),
                                          s2s.fptr(self.c), s2s.fptr(dEdh:contiguous()), dEdc and s2s.fptr(dEdc:contiguous()) or nil,
                                          s2s.fptr(self.gx), s2s.fptr(self.gh), s2s.fptr(self.gc)))
   self.gradInput = {self.gx, self.gh, self.gc}
   return self.gradInput
end
function LSTM:accGradParameters() end
