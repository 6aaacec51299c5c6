-- WeightNoise.lua (shim) -- nn.WeightNoise(parameters, sigma): fixed-sigma Gaussian weight noise (reference WeightNoise.lua:5-35);
-- Sample() = w + sigma * randn on the device with a counter-based Philox stream (one launch, no temporary).
local s2s = require 's2s_ffi'
local WeightNoise, parent = torch.class('nn.WeightNoise', 'nn.Module')

function WeightNoise:__init(parameters, sigma)
   parent.__init(self)
   self.sigma = sigma or 1e-3
   self.weight = parameters:clone()
   self.gradWeight = parameters:clone()
   self.sample = parameters:clone()
   self.calls = 0
end
function WeightNoise:getWeights() return self.weight end
function WeightNoise:Sample()                                   -- WeightNoise.lua:17-22
   self.calls = self.calls + 1
   s2s.check(s2s.C.s2s_weightnoise_sample(s2s.ctx(), s2s.fptr(self.weight), nil, self.calls, self.sigma, self.weight:nElement(), s2s.fptr(self.sample)))
   return self.sample
end
function WeightNoise:Mode() return self.weight end
function WeightNoise:updateOutput(nll) self.nll = nll; return self.nll end
function WeightNoise:accGradParameters(input, gradOutput) self.gradWeight:add(gradOutput) end
