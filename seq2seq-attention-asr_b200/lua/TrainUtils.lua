-- TrainUtils.lua (shim) -- the gradient step of timit/timit.lua:291-348 and TrainUtils.columnNormConstraint
-- (TrainUtils.lua:52-104) on the flat parameter / gradient vectors, plus the weight-noise samplers
-- (WeightNoise.lua:17-22, AdaptiveWeightNoise.lua:27-104).
local s2s = require 's2s_ffi'
local ffi = require 'ffi'
local T = {}

-- gradients:div(B); norm; clip; L2; gradient noise  -> returns the pre-clip norm (timit.lua:292-315)
function T.gradFinalize(gradients, parameters, batchSize, maxnorm, weightDecay, noiseSigma, seed)
   local nrm = ffi.new('double[1]')
   s2s.check(s2s.C.s2s_grad_finalize(s2s.ctx(), s2s.fptr(gradients), s2s.fptr(parameters), gradients:nElement(), batchSize, maxnorm or 1e20,
                                     weightDecay or 0, nil, seed or 0, noiseSigma or 0, nrm))
   return nrm[0]
end
-- optim.adadelta(opfunc, x, config, state) equivalent on precomputed gradients (timit.lua:338-342)
function T.adadelta(x, g, config, state)
   state.paramVariance = state.paramVariance or x.new(x:size()):zero()
   state.accDelta = state.accDelta or x.new(x:size()):zero()
   s2s.check(s2s.C.s2s_adadelta(s2s.ctx(), s2s.fptr(x), s2s.fptr(g), s2s.fptr(state.paramVariance), s2s.fptr(state.accDelta), x:nElement(),
                                config.rho or 0.9, config.eps or 1e-6))
end
function T.columnNormConstraint(m, maxval)
   if not m.weight then return end
   local nan = ffi.new('int[1]')
   s2s.check(s2s.C.s2s_rownorm_constraint(s2s.ctx(), s2s.fptr(m.weight), m.weight:size(1), m.weight:nElement() / m.weight:size(1), maxval or 1, nan))
   if nan[0] ~= 0 then __debug_module = m; error('found a nan, module saved to __debug_module') end   -- TrainUtils.lua:55-62
end
function T.columnNormConstraintModel(cfg, parameters, maxval)          -- columnNormConstraintGraph, timit.lua:346-348
   local nan = ffi.new('int[1]')
   s2s.check(s2s.C.s2s_model_rownorm_constraint(s2s.ctx(), cfg, s2s.fptr(parameters), maxval or 1, nan))
   if nan[0] ~= 0 then error('found a nan') end
end
function T.weightNoiseSample(weight, sigma, sample, seed)             -- WeightNoise:Sample()
   s2s.check(s2s.C.s2s_weightnoise_sample(s2s.ctx(), s2s.fptr(weight), nil, seed or 0, sigma, weight:nElement(), s2s.fptr(sample)))
   return sample
end
function T.awnSample(weight, sample, seed)                            -- AdaptiveWeightNoise:Sample()
   s2s.check(s2s.C.s2s_awn_sample(s2s.ctx(), s2s.fptr(weight), nil, seed or 0, sample:nElement(), s2s.fptr(sample)))
   return sample
end
function T.awnForward(weight, lambda, nll)                            -- AdaptiveWeightNoise:updateOutput
   local L = ffi.new('double[1]')
   s2s.check(s2s.C.s2s_awn_forward(s2s.ctx(), s2s.fptr(weight), weight:nElement() / 2, lambda, nll, L))
   return L[0]
end
function T.awnAccGrad(weight, g, lambda, gradWeight)                  -- AdaptiveWeightNoise:accGradParameters
   s2s.check(s2s.C.s2s_awn_accgrad(s2s.ctx(), s2s.fptr(weight), s2s.fptr(g), g:nElement(), lambda, s2s.fptr(gradWeight)))
end
return T
