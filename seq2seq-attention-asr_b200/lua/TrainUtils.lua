-- TrainUtils.lua (shim) -- the reference's TrainUtils table (TrainUtils.lua:202-213: orthogonalize, orthogonalizeGraph,
-- checkOrthogonalization, columnNormConstraint, columnNormConstraintGraph, checkColumnNormConstraint(Graph), apply2graph, getnorms,
-- checkoutput) with the per-module arithmetic in libs2s_b200.so, plus the flat-vector gradient step of timit/timit.lua:291-348.
-- Graph walking is host-side Lua as in the reference; the shim modules expose their matrices as nn.S2SParam leaves (.weight /
-- .bias views), so apply2graph visits exactly the leaves the reference's graphs show it.
require 'nn'
require 'nngraph'
local s2s = require 's2s_ffi'
local ffi = require 'ffi'
local T = {}

local function is_cuda(t) return torch.type(t) == 'torch.CudaTensor' end

-- m.weight (and m.bias as one more column) := orthonormal factor of its QR in the tall orientation  (TrainUtils.lua:5-26)
function T.orthogonalize(m)
   if not m.weight or m.weight:dim() ~= 2 then return end
   assert(is_cuda(m.weight) and m.weight:isContiguous(), 'orthogonalize (libs2s_b200): CUDA weights only')
   local bias = (m.bias and m.bias:nElement() == m.weight:size(1)) and s2s.fptr(m.bias) or nil
   s2s.check(s2s.C.s2s_orthogonalize(s2s.ctx(), s2s.fptr(m.weight), m.weight:size(1), m.weight:size(2), bias))
end

-- every ROW of m.weight whose L2 norm is >= maxval is divided by norm / maxval; NaN -> error  (TrainUtils.lua:52-104)
function T.columnNormConstraint(m, maxval)
   if not m.weight then return end
   local w = m.weight
   assert(is_cuda(w) and w:isContiguous(), 'columnNormConstraint (libs2s_b200): CUDA weights only')
   local rows = w:dim() == 2 and w:size(1) or 1                -- norm(2,2): one norm per row of a 2-D weight
   local nan = ffi.new('int[1]')
   s2s.check(s2s.C.s2s_rownorm_constraint(s2s.ctx(), s2s.fptr(w), rows, w:nElement() / rows, maxval or 1, nan))
   if nan[0] ~= 0 then
      print('\nmodule', m)
      __debug_module = m
      error('found a nan, module saved to __debug_module')     -- TrainUtils.lua:55-62
   end
end

function T.checkColumnNormConstraint(m) if m.weight and m.weight:dim() == 2 then print(m.weight:norm(2, 2)) end end

local function check_orth(m)                                    -- || w w^T - I || (or w^T w), TrainUtils.lua:29-50
   local w
   if m.weight and m.weight:dim() == 2 then
      w = m.weight
      if m.bias and m.bias:nElement() == w:size(1) then w = torch.cat(w, m.bias:view(m.bias:size(1), 1)) end
      local c = w:size(1) > w:size(2) and torch.mm(w:t(), w) or torch.mm(w, w:t())
      return (c - torch.eye(c:size(1)):typeAs(c)):norm()
   end
end

-- visit every leaf below `graph`: nn.gModule -> its forward nodes' modules, containers / wrappers -> .modules, else func(leaf)
-- (TrainUtils.lua:137-184)
function T.apply2graph(graph, func, toggleprint, prefix)
   prefix = prefix or ''
   local typename = torch.typename(graph) or ''
   local list
   if typename == 'nn.gModule' then list = graph.forwardnodes
   elseif graph.modules then list = graph.modules end
   local printname = graph.__tostring__ and graph:__tostring__() or typename
   if not list then
      if toggleprint and graph.weight then print(prefix .. printname) end
      local r = func(graph)
      if r and toggleprint then print(prefix, r) end
      return
   end
   if toggleprint then print(prefix .. printname) end
   for _, n in pairs(list) do
      local m = n
      if torch.typename(n) == 'nngraph.Node' then m = n.data.module end
      if m ~= nil then T.apply2graph(m, func, toggleprint, prefix .. '  ') end
   end
end

function T.orthogonalizeGraph(graph) T.apply2graph(graph, T.orthogonalize) end
function T.checkOrthogonalization(graph) T.apply2graph(graph, check_orth, true) end
function T.columnNormConstraintGraph(graph) T.apply2graph(graph, T.columnNormConstraint) end            -- timit/timit.lua:346-348
function T.checkColumnNormConstraintGraph(graph) T.apply2graph(graph, T.checkColumnNormConstraint, true) end

function T.getnorms(t)
   if type(t) == 'table' then
      local norms = {}
      for k, v in pairs(t) do norms[k] = T.getnorms(v) end
      return norms
   end
   return t:norm()
end
function T.checkoutput(m) if m.output then return T.getnorms(m.output) end end

-- ---- additions: the flat-vector gradient step of timit/timit.lua:291-348 as fused launches --------------------------------------
-- gradients:div(B); norm; clip; L2; gradient noise  -> returns the pre-clip norm (timit.lua:292-315)
function T.gradFinalize(gradients, parameters, batchSize, maxnorm, weightDecay, noiseSigma, seed)
   local nrm = ffi.new('double[1]')
   s2s.check(s2s.C.s2s_grad_finalize(s2s.ctx(), s2s.fptr(gradients), s2s.fptr(parameters), gradients:nElement(), batchSize, maxnorm or 1e20,
                                     weightDecay or 0, nil, seed or 0, noiseSigma or 0, nrm))
   return nrm[0]
end
-- optim.adadelta(opfunc, x, config, state) on precomputed gradients (timit.lua:338-342)
function T.adadelta(x, g, config, state)
   state.paramVariance = state.paramVariance or x.new(x:size()):zero()
   state.accDelta = state.accDelta or x.new(x:size()):zero()
   s2s.check(s2s.C.s2s_adadelta(s2s.ctx(), s2s.fptr(x), s2s.fptr(g), s2s.fptr(state.paramVariance), s2s.fptr(state.accDelta), x:nElement(),
                                config.rho or 0.9, config.eps or 1e-6))
end
-- data-parallel gradient sum over the ranks (NCCL over NVLink through the C ABI); no-op on one GPU
function T.allreduce(gradients) s2s.check(s2s.C.s2s_dp_allreduce(s2s.ctx(), s2s.fptr(gradients), gradients:nElement())) end

TrainUtils = T          -- the reference sets the global as well (TrainUtils.lua:202)
return T
