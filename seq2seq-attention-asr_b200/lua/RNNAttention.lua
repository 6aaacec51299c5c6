-- RNNAttention.lua (shim) -- nn.RNNAttention(recurrent, dimoutput, reverse): the teacher-forced decoder unroll with non-recurrent
-- inputs {Vh(h), h} shared by all steps (reference RNNAttention.lua:5-253).  In this library the whole unroll -- T steps of
-- attention + GRU + MLP, forward and backward, with the non-recurrent gradients accumulated once after the loop instead of
-- T read-modify-writes of [L,S] + [L,A] (RNNAttention.lua:247) -- is ONE persistent cluster kernel per direction of time, launched by
-- nn.Attention.  This class therefore carries the object surface callers touch (rnn.T, rnn.batchSize, rnn.zeros_y, setT, apply2clones,
-- parameters through the step) and runs a step-at-a-time forward for decoding; it does not clone the step per t.
local RNNAttention, parent = torch.class('nn.RNNAttention', 'nn.Module')

function RNNAttention:__init(recurrent, dimoutput, reverse)
   parent.__init(self)
   assert(recurrent ~= nil, "recurrent cannot be nil")                       -- RNNAttention.lua:8
   assert(dimoutput ~= nil, "recurrent must specify dimoutput")              -- RNNAttention.lua:9
   assert(not reverse, 'nn.RNNAttention (libs2s_b200): reverse decoding is not built (the reference never uses it: Attention.lua:203)')
   self.recurrent = recurrent
   self.dimoutput = dimoutput
   self.reverse = false
   self.output = torch.CudaTensor()
   self.rnn = {}                              -- no per-step clones exist
   self.zeros_y = torch.CudaTensor()
   self.modules = {self.recurrent}
   self.T = 0
end

function RNNAttention:parameters() return self.recurrent:parameters() end
function RNNAttention:training() self.recurrent:training() end
function RNNAttention:evaluate() self.recurrent:evaluate() end
function RNNAttention:cuda() return self end
function RNNAttention:setT(T) self.T = T end                                 -- RNNAttention.lua:132-134
function RNNAttention:apply2clones(func) func(self.recurrent) end            -- RNNAttention.lua:136-141 (the prototype is the only "clone")

-- forward({nonrecurrent = {Vh, h}, y}) -> [T, V] / [B, T, V] log-probabilities, teacher-forced, one step call per label
-- (RNNAttention.lua:144-185).  Decode / inspection path: training goes through nn.Attention:forward / :backward.
function RNNAttention:updateOutput(input)
   local nonrec, y = unpack(input)
   local batch = y:nDimension() == 3
   local sdim = batch and 2 or 1
   local T = self.T > 0 and self.T or y:size(sdim)
   self.sequence_dim, self.batchSize = sdim, batch and y:size(1) or 0
   local out, hidden, prev_y = {}, nil, nil
   for t = 1, T do
      prev_y = t > 1 and y:select(sdim, t - 1) or nil                         -- zeros at t = 1 (RNNAttention.lua:172-176)
      local o = self.recurrent:forward({{nonrec, prev_y or self.zeros_y}, hidden})
      out[t], hidden = o[1], o[2]
   end
   local sz = y:size(); self.output:resize(sz)
   for t = 1, T do self.output:select(sdim, t):copy(out[t]) end
   return self.output
end
function RNNAttention:updateGradInput() error('nn.RNNAttention (libs2s_b200): backward runs inside nn.Attention:backward (one fused time loop)') end
