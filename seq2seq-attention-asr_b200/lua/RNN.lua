-- RNN.lua (shim) -- nn.RNN(recurrent, reverse) over a whole [L,D] / [B,L,D] sequence (reference RNN.lua:5-201) for
-- recurrent = nn.GRU(diminput, dimoutput) or nn.LSTM(diminput, dimoutput, peepholes).  One library call runs the time-batched
-- input projection (tcgen05 GEMM) and the persistent cluster recurrence kernel; no per-frame clones exist, so the reference's
-- addClone / resetCloneParameters bookkeeping (RNN.lua:21-64) has nothing to do and float()/double() have no meaning.
local s2s = require 's2s_ffi'

local RNN, parent = torch.class('nn.RNN', 'nn.Module')

function RNN:__init(recurrent, reverse)
   parent.__init(self)
   assert(recurrent ~= nil, "recurrent cannot be nil")                       -- RNN.lua:8
   assert(recurrent.dimoutput ~= nil, "recurrent must specify dimoutput")    -- RNN.lua:9
   local kind = torch.typename(recurrent)
   assert(kind == 'nn.GRU' or kind == 'nn.LSTM', 'nn.RNN (libs2s_b200): the step module must be nn.GRU or nn.LSTM, got ' .. tostring(kind))
   self.recurrent = recurrent
   self.lstm = kind == 'nn.LSTM'
   self.dimoutput = recurrent.dimoutput
   self.T = 0
   self.reverse = reverse or false
   self.modules = {self.recurrent}
end

function RNN:parameters() return self.recurrent:parameters() end
function RNN:training() self.recurrent:training() end
function RNN:evaluate() self.recurrent:evaluate() end
function RNN:cuda() self.recurrent:cuda(); return self end
function RNN:float() error('nn.RNN (libs2s_b200): CUDA only, there is no CPU path') end
RNN.double = RNN.float
function RNN:type(t)
   assert(t == nil or t == 'torch.CudaTensor', 'nn.RNN (libs2s_b200): CUDA only, there is no CPU path')
   return t and self or 'torch.CudaTensor'
end

local function dims(x)
   if x:nDimension() == 2 then return 1, x:size(1), x:size(2), 1 end          -- RNN.lua:123-129: 2-D = one utterance
   if x:nDimension() == 3 then return x:size(1), x:size(2), x:size(3), 2 end
   error('input dimension must be 2D or 3D')
end

function RNN:updateOutput(input)
   local x = input:contiguous()
   local B, L, D, sdim = dims(x)
   local H, r, rev = self.dimoutput, self.recurrent, self.reverse and 1 or 0
   self.sequence_dim, self.T = sdim, L
   if sdim == 1 then self.output:resize(L, H) else self.output:resize(B, L, H) end
   self.save = self.save or torch.CudaTensor()
   if self.lstm then
      self.save:resize(tonumber(s2s.C.s2s_lstm_seq_save_floats(B, L, H)))
      s2s.check(s2s.C.s2s_lstm_seq_forward(s2s.ctx(), s2s.fptr(r.weight), D, H, r.peepholes and 1 or 0, rev, s2s.fptr(x), D, nil, B, L,
                                           s2s.fptr(self.output), s2s.fptr(self.save)))
   else
      self.save:resize(tonumber(s2s.C.s2s_gru_seq_save_floats(B, L, H, 1)))
      s2s.check(s2s.C.s2s_gru_seq_forward(s2s.ctx(), r:_wptr(), D, H, 1, rev, s2s.fptr(x), D, nil, B, L, s2s.fptr(self.output), s2s.fptr(self.save)))
   end
   self.B, self.L, self.D = B, L, D
   return self.output
end

function RNN:updateGradInput(input, gradOutput)
   assert(self.save and input:size(self.sequence_dim) == self.T, "sequence size of input must match self.T")   -- RNN.lua:171
   local H, r, rev = self.dimoutput, self.recurrent, self.reverse and 1 or 0
   self.gradInput:resizeAs(input)
   local x, dy = input:contiguous(), gradOutput:contiguous()
   if self.lstm then
      s2s.check(s2s.C.s2s_lstm_seq_backward(s2s.ctx(), s2s.fptr(r.weight), s2s.fptr(r.gradWeight), self.D, H, r.peepholes and 1 or 0, rev, s2s.fptr(x),
                                            self.D, nil, self.B, self.L, s2s.fptr(self.output), s2s.fptr(self.save), s2s.fptr(dy), s2s.fptr(self.gradInput)))
   else
      s2s.check(s2s.C.s2s_gru_seq_backward(s2s.ctx(), r:_wptr(), r:_gptr(), self.D, H, 1, rev, s2s.fptr(x), self.D, nil, self.B, self.L,
                                           s2s.fptr(self.output), s2s.fptr(self.save), s2s.fptr(dy), s2s.fptr(self.gradInput)))
   end
   return self.gradInput
end

function RNN:accGradParameters() end   -- accumulated inside updateGradInput, as in the reference wrappers (Recurrent.lua:148)
