-- TemporalConvolutionZeroBias.lua (shim) -- nn.TemporalConvolutionZeroBias(inputFrameSize, outputFrameSize, kW, dW)
-- for kW = 1 (the only use in the reference: Vh, UF, e; Attention.lua:44,91,110).  Bias kept and pinned to zero
-- (TemporalConvolutionZeroBias.lua:14-16,38,43,52-53) so parameter counts match.
local s2s = require 's2s_ffi'

local TCZB, parent = torch.class('nn.TemporalConvolutionZeroBias', 'nn.Module')

function TCZB:__init(inputFrameSize, outputFrameSize, kW, dW)
   parent.__init(self)
   assert((kW or 1) == 1 and (dW or 1) == 1, 'libs2s_b200: only kW = dW = 1 is on the hot path')
   self.inputFrameSize, self.outputFrameSize, self.kW, self.dW = inputFrameSize, outputFrameSize, 1, 1
   self.weight = torch.CudaTensor(outputFrameSize, inputFrameSize)
   self.bias = torch.CudaTensor(outputFrameSize):zero()
   self.gradWeight = torch.CudaTensor(outputFrameSize, inputFrameSize):zero()
   self.gradBias = torch.CudaTensor(outputFrameSize):zero()
   self:reset()
end

function TCZB:reset(stdv)                     -- TemporalConvolutionZeroBias.lua:21-35
   if stdv then stdv = stdv * math.sqrt(3) else stdv = 1 / math.sqrt(self.kW * self.inputFrameSize) end
   self.weight:uniform(-stdv, stdv)
   self.bias:zero()
end

function TCZB:updateOutput(input)
   local x = input:contiguous()
   local rows = x:nElement() / self.inputFrameSize
   local sz = x:size(); sz[#sz] = self.outputFrameSize
   self.output:resize(sz)
   s2s.check(s2s.C.s2s_tconv_zb_forward(s2s.ctx(), s2s.fptr(x), rows, self.inputFrameSize, s2s.fptr(self.weight), self.outputFrameSize, s2s.fptr(self.output)))
   return self.output
end

function TCZB:updateGradInput(input, gradOutput)
   self.gradInput:resizeAs(input)
   local rows = input:nElement() / self.inputFrameSize
   s2s.check(s2s.C.s2s_tconv_zb_backward(s2s.ctx(), s2s.fptr(input:contiguous()), rows, self.inputFrameSize, s2s.fptr(self.weight), self.outputFrameSize,
                                         s2s.fptr(gradOutput:contiguous()), s2s.fptr(self.gradInput), nil, 1))
   return self.gradInput
end

function TCZB:accGradParameters(input, gradOutput, scale)
   local rows = input:nElement() / self.inputFrameSize
   s2s.check(s2s.C.s2s_tconv_zb_backward(s2s.ctx(), s2s.fptr(input:contiguous()), rows, self.inputFrameSize, s2s.fptr(self.weight), self.outputFrameSize,
                                         s2s.fptr(gradOutput:contiguous()), nil, s2s.fptr(self.gradWeight), scale or 1))
   self.gradBias:zero()
end
