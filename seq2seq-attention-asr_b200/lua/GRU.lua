-- GRU.lua (shim) -- nn.GRU(diminput, dimoutput), a subclass of nn.Recurrent like the reference's (GRU.lua:6-51):
--     z = sigmoid(W_z {h, x})   r = sigmoid(W_r {h, x})   h~ = tanh(W_h {r*h, x})   h' = (1-z) h + z h~        (GRU.lua:22-30)
-- three LinearZeroBias weights [out, out+in] (concat order {prev_h, x}, no biases), stored as ONE [3, out, out+in] tensor so that
-- nn.RNN hands the whole recurrence to the persistent cluster kernel with one pointer.  Used two ways by the reference:
--   * as the step module of nn.RNN (model_chorowski_baseline.lua:22-31): nn.RNN never calls the step, it calls s2s_gru_seq_*;
--   * as an nngraph node / stand-alone step, forward({x, prev_h}) -> h (model_chorowski_baseline.lua:50): s2s_gru_step_*.
require 'Recurrent'
local s2s = require 's2s_ffi'

local GRU, parent = torch.class('nn.GRU', 'nn.Recurrent')

function GRU:__init(diminput, dimoutput)     -- extra arguments are ignored, as in the reference (GRU.lua:8)
   assert(diminput ~= nil, "diminput must be specified")          -- GRU.lua:9-10
   assert(dimoutput ~= nil, "dimoutput must be specified")
   nn.Module.__init(self)
   self.diminput, self.dimoutput, self.dimhidden = diminput, dimoutput, dimoutput
   self.weight = torch.CudaTensor(3, dimoutput, dimoutput + diminput)
   self.gradWeight = torch.CudaTensor(3, dimoutput, dimoutput + diminput):zero()
   self.zeros_hidden = torch.zeros(dimoutput)
   -- the three LinearZeroBias leaves TrainUtils.apply2graph finds in the reference's GRU graph (row-norm constraint, orthogonalize)
   self.modules = {}
   for i, name in ipairs({'z', 'r', 'h'}) do
      self.modules[i] = nn.S2SParam('GRU.' .. name, self.weight[i], self.gradWeight[i])
   end
   self.recurrent = self
   self:reset()
end

function GRU:reset(stdv)                     -- LinearZeroBias.lua:12-29, once per gate
   if stdv then stdv = stdv * math.sqrt(3) else stdv = 1 / math.sqrt(self.dimoutput + self.diminput) end
   self.weight:uniform(-stdv, stdv)
end

function GRU:parameters()
   local p, g = {}, {}
   for i = 1, 3 do p[i] = self.modules[i].weight; g[i] = self.modules[i].gradWeight end
   return p, g
end
function GRU:training() self.train = true end
function GRU:evaluate() self.train = false end
function GRU:float() error('nn.GRU (libs2s_b200): CUDA only, there is no CPU path') end
GRU.double = GRU.float
function GRU:type(t)
   assert(t == nil or t == 'torch.CudaTensor', 'nn.GRU (libs2s_b200): CUDA only, there is no CPU path')
   return t and self or 'torch.CudaTensor'
end
function GRU:cuda() self.zeros_hidden = self.zeros_hidden:cuda(); return self end

-- the flat [3, out, out+in] weight may have been re-pointed gate by gate by getParameters(); the gates stay contiguous in
-- the flattened storage (nn.Module.flatten keeps tensors that shared a storage together), so gate z's pointer is the base
local function wptr(self) return s2s.fptr(self.modules[1].weight) end
local function gptr(self) return s2s.fptr(self.modules[1].gradWeight) end
GRU._wptr, GRU._gptr = wptr, gptr

function GRU:updateOutput(input)             -- one step: {x, prev_h} -> h   (Recurrent.lua:104-127 + GRU.lua:22-38)
   local x, prev_h = unpack(input)
   if type(x) == 'table' and #x == 1 then x = x[1] end
   local B = x:dim() == 2 and x:size(1) or 1
   self:resetZeros(x)
   prev_h = prev_h or self.zeros_hidden
   if x:dim() == 2 then self.output:resize(B, self.dimoutput) else self.output:resize(self.dimoutput) end
   self.gates = self.gates or torch.CudaTensor()
   self.gates:resize(B, 3 * self.dimoutput)
   s2s.check(s2s.C.s2s_gru_step_forward(s2s.ctx(), wptr(self), self.diminput, self.dimoutput, s2s.fptr(x:contiguous()),
                                        s2s.fptr(prev_h:contiguous()), B, s2s.fptr(self.output), s2s.fptr(self.gates)))
   return self.output
end

function GRU:resetZeros(x)
   local z = self.zeros_hidden
   if torch.type(z) ~= 'torch.CudaTensor' then z = z:cuda() end
   if x:dim() == 2 then
      if z:dim() ~= 2 or z:size(1) ~= x:size(1) then z:resize(x:size(1), self.dimoutput):zero() end
   elseif z:dim() ~= 1 then z:resize(self.dimoutput):zero() end
   self.zeros_hidden = z
end

function GRU:updateGradInput(input, gradOutput)     -- -> {dEdx, dEdph}; weight gradients accumulate here (Recurrent.lua:148)
   if type(gradOutput) == 'table' then gradOutput = gradOutput[1] end      -- GRU.lua:45-51
   local x, prev_h = unpack(input)
   if type(x) == 'table' and #x == 1 then x = x[1] end
   local B = x:dim() == 2 and x:size(1) or 1
   prev_h = prev_h or self.zeros_hidden
   self.gradX = self.gradX or torch.CudaTensor(); self.gradH = self.gradH or torch.CudaTensor()
   self.gradX:resizeAs(x); self.gradH:resizeAs(prev_h)
   s2s.check(s2s.C.s2s_gru_step_backward(s2s.ctx(), wptr(self), gptr(self), self.diminput, self.dimoutput, s2s.fptr(x:contiguous()),
                                         s2s.fptr(prev_h:contiguous()), B, s2s.fptr(self.gates), s2s.fptr(gradOutput:contiguous()),
                                         s2s.fptr(self.gradX), s2s.fptr(self.gradH)))
   self.gradInput = {self.gradX, self.gradH}
   return self.gradInput
end

function GRU:accGradParameters() end
