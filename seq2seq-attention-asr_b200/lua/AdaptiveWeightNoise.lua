-- AdaptiveWeightNoise.lua (shim) -- nn.AdaptiveWeightNoise(parameters, lambda, sigma_init): Graves-2011 variational weight noise
-- (reference AdaptiveWeightNoise.lua:5-104).  weight = {mu ; s = log sigma^2} (2n); Sample() = mu + exp(s/2) randn;
-- updateOutput(nll) = lambda KL + nll with the empirical-Bayes Gaussian prior; accGradParameters gives d/dmu and d/ds.
-- Each is one fused launch of libs2s_b200.so (csrc/optim.cu).
local s2s = require 's2s_ffi'
local ffi = require 'ffi'
local AWN, parent = torch.class('nn.AdaptiveWeightNoise', 'nn.Module')

function AWN:__init(parameters, lambda, sigma_init)
   parent.__init(self)
   self.alpha_mu, self.alpha_sigma2 = 0, 1
   self.n = parameters:size(1)
   self.weight = torch.CudaTensor(2 * self.n)
   self:initialize(parameters, sigma_init or 1)                  -- default sigma_init = 1 (AdaptiveWeightNoise.lua:13)
   self.sample = parameters:clone():zero()
   self.gradWeight = self.weight:clone():zero()
   self.lambda = lambda or 1
   self.calls = 0
end
function AWN:initialize(mu_init, sigma_init)                     -- AdaptiveWeightNoise.lua:40-56
   local mu, s = unpack(self.weight:split(self.n))
   if type(mu_init) == 'number' then mu:fill(mu_init) else mu:copy(mu_init) end
   if type(sigma_init) == 'number' then s:fill(sigma_init ^ 2):log() else s:copy(sigma_init):pow(2):log() end
end
function AWN:getWeights() local mu = self.weight:narrow(1, 1, self.n); return mu, nil end
function AWN:Mode() return self.weight:narrow(1, 1, self.n) end
function AWN:Sample()                                            -- AdaptiveWeightNoise.lua:27-38
   self.calls = self.calls + 1
   s2s.check(s2s.C.s2s_awn_sample(s2s.ctx(), s2s.fptr(self.weight), nil, self.calls, self.n, s2s.fptr(self.sample)))
   return self.sample
end
function AWN:updateOutput(nll)                                   -- AdaptiveWeightNoise.lua:63-80
   self.nll = nll
   local L = ffi.new('double[1]')
   s2s.check(s2s.C.s2s_awn_forward(s2s.ctx(), s2s.fptr(self.weight), self.n, self.lambda, nll, L))
   self.L = L[0]
   return self.L
end
function AWN:accGradParameters(input, gradOutput)                -- AdaptiveWeightNoise.lua:82-104
   s2s.check(s2s.C.s2s_awn_accgrad(s2s.ctx(), s2s.fptr(self.weight), s2s.fptr(gradOutput:contiguous()), self.n, self.lambda, s2s.fptr(self.gradWeight)))
end
