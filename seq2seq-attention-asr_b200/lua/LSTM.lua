-- LSTM.lua (shim) -- nn.LSTM(diminput, dimoutput, peepholes): parameter holder in the reference's order
-- (LSTM.lua:25-60: per gate Linear(in,out) + Linear(out,out), both with bias; full-matrix peepholes), stored as
-- one flat tensor so nn.RNN can hand it to s2s_lstm_seq_forward / _backward in one pointer.
local s2s = require 's2s_ffi'
local ffi = require 'ffi'
ffi.cdef[[
int64_t s2s_lstm_param_count(int in_, int out, int peepholes);
int64_t s2s_lstm_seq_save_floats(int B, int Lmax, int H);
int s2s_lstm_seq_forward(s2s_ctx*, const float* P, int Din, int H, int peepholes, int reverse, const float* x, int ldx, const int* lengths, int B, int Lmax, float* y, float* save);
int s2s_lstm_seq_backward(s2s_ctx*, const float* P, float* dP, int Din, int H, int peepholes, int reverse, const float* x, int ldx, const int* lengths, int B, int Lmax, const float* y, const float* save, const float* dy, float* dx);
]]
local LSTM, parent = torch.class('nn.LSTM', 'nn.Module')

function LSTM:__init(diminput, dimoutput, peepholes)
   parent.__init(self)
   assert(diminput ~= nil, "diminput must be specified")      -- LSTM.lua:9
   assert(dimoutput ~= nil, "dimoutput must be specified")    -- LSTM.lua:10
   self.diminput, self.dimoutput, self.peepholes = diminput, dimoutput, peepholes or false
   local n = tonumber(s2s.C.s2s_lstm_param_count(diminput, dimoutput, self.peepholes and 1 or 0))
   self.weight = torch.CudaTensor(n)
   self.gradWeight = torch.CudaTensor(n):zero()
   self:reset()
end
function LSTM:reset(stdv)
   self.weight:uniform(-(stdv or 1 / math.sqrt(self.dimoutput)), stdv or 1 / math.sqrt(self.dimoutput))
end
function LSTM:parameters() return {self.weight}, {self.gradWeight} end
-- nn.RNN(nn.LSTM(...)) dispatches on torch.typename(self.recurrent) == 'nn.LSTM' to the two calls above.
