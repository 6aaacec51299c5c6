-- RNN.lua (shim) -- nn.RNN(recurrent, reverse) over a whole [L,D] / [B,L,D] sequence (reference RNN.lua:5-201),
-- for recurrent = nn.GRU(diminput, dimoutput).  One library call runs the time-batched input projection
-- (tcgen05 GEMM) and the persistent cluster recurrence kernel; no per-frame clones exist.
local s2s = require 's2s_ffi'

local RNN, parent = torch.class('nn.RNN', 'nn.Module')

function RNN:__init(recurrent, reverse)
   parent.__init(self)
   assert(recurrent ~= nil, "recurrent cannot be nil")                       -- RNN.lua:8
   assert(recurrent.dimoutput ~= nil, "recurrent must specify dimoutput")    -- RNN.lua:9
   self.recurrent = recurrent
   self.dimoutput = recurrent.dimoutput
   self.reverse = reverse or false
   self.modules = {self.recurrent}
end

function RNN:parameters() return self.recurrent:parameters() end

function RNN:updateOutput(input)
   local x = input:contiguous()
   local B, L, D
   if x:nDimension() == 2 then B, L, D = 1, x:size(1), x:size(2)             -- RNN.lua:123-129
   elseif x:nDimension() == 3 then B, L, D = x:size(1), x:size(2), x:size(3)
   else error('input must be 2d or 3d') end
   local H = self.dimoutput
   if x:nDimension() == 2 then self.output:resize(L, H) else self.output:resize(B, L, H) end
   self.save = self.save or torch.CudaTensor()
   self.save:resize(tonumber(s2s.C.s2s_gru_seq_save_floats(B, L, H, 1)))
   s2s.check(s2s.C.s2s_gru_seq_forward(s2s.ctx(), s2s.fptr(self.recurrent.weight), D, H, 1, self.reverse and 1 or 0,
                                       s2s.fptr(x), D, nil, B, L, s2s.fptr(self.output), s2s.fptr(self.save)))
   self.B, self.L, self.D = B, L, D
   return self.output
end

function RNN:updateGradInput(input, gradOutput)
   assert(self.save, 'backward called before forward')                       -- RNN.lua:171
   self.gradInput:resizeAs(input)
   s2s.check(s2s.C.s2s_gru_seq_backward(s2s.ctx(), s2s.fptr(self.recurrent.weight), s2s.fptr(self.recurrent.gradWeight), self.D,
                                        self.dimoutput, 1, self.reverse and 1 or 0, s2s.fptr(input:contiguous()), self.D, nil, self.B,
                                        self.L, s2s.fptr(self.output), s2s.fptr(self.save), s2s.fptr(gradOutput:contiguous()),
                                        s2s.fptr(self.gradInput)))
   return self.gradInput
end

function RNN:accGradParameters() end   -- accumulated inside updateGradInput, as in the reference wrappers
