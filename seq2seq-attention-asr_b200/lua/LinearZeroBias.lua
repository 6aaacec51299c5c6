-- LinearZeroBias.lua (shim) -- nn.LinearZeroBias(inputSize, outputSize): nn.Linear without a bias (reference LinearZeroBias.lua:3-83),
-- computing through libs2s_b200.so (s2s_linear_zb_forward / _backward; large products run on the tcgen05 GEMM).
local s2s = require 's2s_ffi'

local LinearZeroBias, parent = torch.class('nn.LinearZeroBias', 'nn.Module')

function LinearZeroBias:__init(inputSize, outputSize)
   parent.__init(self)
   self.weight = torch.CudaTensor(outputSize, inputSize)          -- [out, in]  (LinearZeroBias.lua:6)
   self.gradWeight = torch.CudaTensor(outputSize, inputSize):zero()
   self:reset()
end

function LinearZeroBias:reset(stdv)                                -- LinearZeroBias.lua:12-29
   if stdv then stdv = stdv * math.sqrt(3) else stdv = 1 / math.sqrt(self.weight:size(2)) end
   self.weight:uniform(-stdv, stdv)
   return self
end

local function rows_of(input)
   if input:dim() == 1 then return 1 end
   if input:dim() == 2 then return input:size(1) end
   error('input must be vector or matrix')                         -- LinearZeroBias.lua:44
end

function LinearZeroBias:updateOutput(input)
   local rows, out, inp = rows_of(input), self.weight:size(1), self.weight:size(2)
   if input:dim() == 1 then self.output:resize(out) else self.output:resize(rows, out) end
   s2s.check(s2s.C.s2s_linear_zb_forward(s2s.ctx(), s2s.fptr(input:contiguous()), rows, inp, s2s.fptr(self.weight), out, s2s.fptr(self.output)))
   return self.output
end

function LinearZeroBias:updateGradInput(input, gradOutput)
   if not self.gradInput then return end
   local rows, out, inp = rows_of(input), self.weight:size(1), self.weight:size(2)
   self.gradInput:resizeAs(input)
   s2s.check(s2s.C.s2s_linear_zb_backward(s2s.ctx(), s2s.fptr(input:contiguous()), rows, inp, s2s.fptr(self.weight), out,
                                          s2s.fptr(gradOutput:contiguous()), s2s.fptr(self.gradInput), nil, 1))
   return self.gradInput
end

function LinearZeroBias:accGradParameters(input, gradOutput, scale)   -- gradWeight += scale * gradOutput^T input  (:67-74)
   local rows, out, inp = rows_of(input), self.weight:size(1), self.weight:size(2)
   s2s.check(s2s.C.s2s_linear_zb_backward(s2s.ctx(), s2s.fptr(input:contiguous()), rows, inp, s2s.fptr(self.weight), out,
                                          s2s.fptr(gradOutput:contiguous()), nil, s2s.fptr(self.gradWeight), scale or 1))
end

LinearZeroBias.sharedAccUpdateGradParameters = LinearZeroBias.accUpdateGradParameters

function LinearZeroBias:__tostring__()
   return torch.type(self) .. string.format('(%d -> %d)', self.weight:size(2), self.weight:size(1))
end
