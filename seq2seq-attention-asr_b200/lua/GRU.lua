-- GRU.lua (shim) -- nn.GRU(diminput, dimoutput): parameter holder with the reference's layout
-- (GRU.lua:8-43): three LinearZeroBias weights z, r, h~ of shape [out, out+in], concat order {prev_h, x},
-- stored contiguously as one [3, out, out+in] tensor so nn.RNN can hand it to the library in one pointer.
local GRU, parent = torch.class('nn.GRU', 'nn.Module')

function GRU:__init(diminput, dimoutput)     -- extra arguments are ignored, as in the reference (GRU.lua:8)
   parent.__init(self)
   self.diminput, self.dimoutput = diminput, dimoutput
   self.weight = torch.CudaTensor(3, dimoutput, dimoutput + diminput)
   self.gradWeight = torch.CudaTensor(3, dimoutput, dimoutput + diminput):zero()
   self:reset()
end

function GRU:reset(stdv)                     -- LinearZeroBias.lua:12-29
   stdv = stdv or 1 / math.sqrt(self.dimoutput + self.diminput)
   self.weight:uniform(-stdv, stdv)
end

function GRU:parameters()
   local p, g = {}, {}
   for i = 1, 3 do p[i] = self.weight[i]; g[i] = self.gradWeight[i] end
   return p, g
end
