-- Recurrent.lua (shim) -- nn.Recurrent(recurrent, dimhidden): gives a step module the {inp, prev_h} -> output contract with
-- lazily sized all-zero initial hidden state (reference Recurrent.lua:6-151).  Host-side glue only: no arithmetic happens
-- here, the wrapped module does the work (nn.GRU / nn.LSTM steps and the nn.Attention decoder step call the library).
local Recurrent, parent = torch.class('nn.Recurrent', 'nn.Module')

-- apply f to every leaf of a (nested) table of sizes / tensors, keeping the nesting
local function map(x, f, y)
   if type(x) == 'table' then
      local out = {}
      for i = 1, #x do out[i] = map(x[i], f, y and y[i]) end
      return out
   end
   return f(x, y)
end

function Recurrent:__init(recurrent, dimhidden)
   parent.__init(self)
   self.recurrent = recurrent
   self.modules = {recurrent}
   self.dimhidden = dimhidden
   self.zeros_hidden = map(dimhidden, function(n) return torch.zeros(n) end)     -- Recurrent.lua:13
end

function Recurrent:parameters() return self.recurrent:parameters() end
function Recurrent:training() self.recurrent:training() end
function Recurrent:evaluate() self.recurrent:evaluate() end

local function cast(self, typename)
   self.recurrent:type(typename)
   self.zeros_hidden = map(self.zeros_hidden, function(x) return x:type(typename) end)
   return self
end
function Recurrent:float() return cast(self, 'torch.FloatTensor') end
function Recurrent:double() return cast(self, 'torch.DoubleTensor') end
function Recurrent:cuda() return cast(self, 'torch.CudaTensor') end

local function batch_size(x)                                     -- 3-D input = batch mode, else "SGD" mode (Recurrent.lua:67-78)
   if type(x) == 'table' then return batch_size(x[1]) end
   return x:nDimension() == 3 and x:size(1) or 0
end

function Recurrent:resetZeros(inp)                                -- Recurrent.lua:79-102
   local B = batch_size(inp)
   self.zeros_hidden = map(self.zeros_hidden, function(z, dim)
      if B > 0 then
         if z:nDimension() ~= 2 or z:size(1) ~= B or z:size(2) ~= dim then z:resize(B, dim):zero() end
      else
         if z:nDimension() ~= 1 or z:size(1) ~= dim then z:resize(dim):zero() end
      end
      return z
   end, self.dimhidden)
end

function Recurrent:updateOutput(input)                            -- Recurrent.lua:104-127
   local inp, prev_h = unpack(input)
   if type(inp) == 'table' and #inp == 1 then inp = inp[1] end
   self:resetZeros(inp)
   prev_h = prev_h or self.zeros_hidden
   self.output = self.recurrent:forward({inp, prev_h})
   return self.output
end

function Recurrent:updateGradInput(input, gradOutput)             -- Recurrent.lua:129-151
   local inp, prev_h = unpack(input)
   local dEdy, dEdh
   if type(gradOutput) == 'table' then
      dEdy, dEdh = unpack(gradOutput)
      dEdh = dEdh or self.zeros_hidden
      gradOutput = {dEdy, dEdh}
      assert(prev_h ~= nil or dEdh ~= nil, "prev_h and dEdh cannot both be nil")
   else
      dEdy = gradOutput
   end
   assert(dEdy ~= nil, "dEdy cannot be nil")
   if type(inp) == 'table' and #inp == 1 then inp = inp[1] end
   prev_h = prev_h or self.zeros_hidden
   local dEdx, dEdph = unpack(self.recurrent:backward({inp, prev_h}, gradOutput))
   self.gradInput = {dEdx, dEdph}
   return self.gradInput
end
