"""Host-side mirror of the reference's Torch7 module surface, in Python because the image has no
LuaJIT/Torch7 (the Lua shims with the same structure are in lua/).  Same class names, constructor
arguments, method names and error behaviour as the reference:

    nn.TemporalConvolutionZeroBias(inF, outF, kW, dW)     TemporalConvolutionZeroBias.lua:3
    nn.LinearZeroBias(inF, outF)                          LinearZeroBias.lua:3
    nn.GRU(diminput, dimoutput)                           GRU.lua:8
    nn.RNN(recurrent, reverse)                            RNN.lua:5
    nn.Attention(decoder_recurrent, decoder_mlp, scoreDepth, filterSize, featureMaps,
                 stateDepth, annotationDepth, outputDepth, monoAlignPenalty, penaltyLambda)   Attention.lua:15-24
    nn.WeightNoise(parameters, sigma) / nn.AdaptiveWeightNoise(parameters, lambda, sigma_init)
    TrainUtils.columnNormConstraint(m, maxval)            TrainUtils.lua:52

Module protocol: updateOutput(input) -> self.output ; updateGradInput(input, gradOutput) -> self.gradInput ;
accGradParameters(input, gradOutput, scale) ; forward / backward / zeroGradParameters / parameters().
2-D inputs are single utterances ("SGD mode"), 3-D inputs are batches (RNN.lua:123-129, Attention.lua:308-316).
Every method is argument marshalling around one C-ABI call; nothing here computes.
"""
import math

import torch

from . import ops


class Module:
    def __init__(self, ctx):
        self.ctx = ctx
        self.output = None
        self.gradInput = None
        self.train = True

    def parameters(self):
        return [], []

    def forward(self, input):
        return self.updateOutput(input)

    def backward(self, input, gradOutput, scale=1.0):
        gi = self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput, scale)
        return gi

    def accGradParameters(self, input, gradOutput, scale=1.0):
        pass

    def zeroGradParameters(self):
        for g in self.parameters()[1]:
            g.zero_()

    def training(self):
        self.train = True

    def evaluate(self):
        self.train = False


class TemporalConvolutionZeroBias(Module):
    """kW = dW = 1 only (every use in the reference: Vh, UF, e -- Attention.lua:44,91,110)"""

    def __init__(self, ctx, inputFrameSize, outputFrameSize, kW=1, dW=1):
        super().__init__(ctx)
        if kW != 1 or dW != 1:
            raise ops.S2SError("libs2s_b200: only kW = dW = 1 is on the hot path")
        self.inputFrameSize, self.outputFrameSize = inputFrameSize, outputFrameSize
        self.weight = ctx.new(outputFrameSize, inputFrameSize)
        self.bias = ctx.zeros(outputFrameSize)            # pinned to zero (TemporalConvolutionZeroBias.lua:38)
        self.gradWeight = ctx.zeros(outputFrameSize, inputFrameSize)
        self.gradBias = ctx.zeros(outputFrameSize)
        self.reset()

    def reset(self, stdv=None):
        stdv = stdv or 1.0 / math.sqrt(self.inputFrameSize)
        self.weight.uniform_(-stdv, stdv)
        self.bias.zero_()

    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]

    def updateOutput(self, input):
        self.output = ops.tconv_zb_forward(self.ctx, input.contiguous(), self.weight)
        return self.output

    def updateGradInput(self, input, gradOutput):
        self.gradInput = ops.tconv_zb_backward(self.ctx, input.contiguous(), self.weight, gradOutput.contiguous())
        return self.gradInput

    def accGradParameters(self, input, gradOutput, scale=1.0):
        ops.tconv_zb_backward(self.ctx, input.contiguous(), self.weight, gradOutput.contiguous(), dW=self.gradWeight, scale=scale, need_dx=False)
        self.gradBias.zero_()                              # TemporalConvolutionZeroBias.lua:52-53


class LinearZeroBias(TemporalConvolutionZeroBias):
    def __init__(self, ctx, inputSize, outputSize):
        super().__init__(ctx, inputSize, outputSize)

    def parameters(self):
        return [self.weight], [self.gradWeight]


class GRU(Module):
    """Parameter holder: z, r, h~ LinearZeroBias weights [out, out+in], concat order {prev_h, x} (GRU.lua:22-26)"""

    def __init__(self, ctx, diminput, dimoutput, *ignored):
        super().__init__(ctx)
        self.diminput, self.dimoutput = diminput, dimoutput
        self.weight = ctx.new(3, dimoutput, dimoutput + diminput)
        self.gradWeight = ctx.zeros(3, dimoutput, dimoutput + diminput)
        self.reset()

    def reset(self, stdv=None):
        stdv = stdv or 1.0 / math.sqrt(self.dimoutput + self.diminput)
        self.weight.uniform_(-stdv, stdv)

    def parameters(self):
        return [self.weight[i] for i in range(3)], [self.gradWeight[i] for i in range(3)]

    # single-step protocol through nn.Recurrent: input {x, prev_h} -> h (Recurrent.lua:104-127, GRU.lua:22-38)
    def updateOutput(self, input):
        x, hp = (input if isinstance(input, (list, tuple)) else (input, None))
        self._batched = x.dim() == 2
        xb = x.contiguous() if self._batched else x.contiguous().unsqueeze(0)
        hpb = None if hp is None else (hp.contiguous() if self._batched else hp.contiguous().unsqueeze(0))
        hn, self._gates = ops.gru_step_forward(self.ctx, self.weight, xb, hpb)
        self._x, self._hp = xb, hpb
        self.output = hn if self._batched else hn[0]
        return self.output

    def updateGradInput(self, input, gradOutput):
        dhn = gradOutput.contiguous() if self._batched else gradOutput.contiguous().unsqueeze(0)
        dx, dhp, _ = ops.gru_step_backward(self.ctx, self.weight, self._x, self._hp, self._gates, dhn, dW=self.gradWeight)
        self.gradInput = [dx, dhp] if self._batched else [dx[0], dhp[0]]
        return self.gradInput


class LSTM(Module):
    """Parameter holder in the reference's order (LSTM.lua:25-60): per gate Linear(in,out)+Linear(out,out), both with
    bias, plus full-matrix peepholes when requested."""

    def __init__(self, ctx, diminput, dimoutput, peepholes=False):
        super().__init__(ctx)
        assert diminput is not None, "diminput must be specified"
        assert dimoutput is not None, "dimoutput must be specified"
        self.diminput, self.dimoutput, self.peepholes = diminput, dimoutput, bool(peepholes)
        n = ops.lstm_param_count(diminput, dimoutput, self.peepholes)
        self.weight = ctx.new(n)
        self.gradWeight = ctx.zeros(n)
        self.reset()

    def reset(self, stdv=None):
        stdv = stdv or 1.0 / math.sqrt(self.dimoutput)
        self.weight.uniform_(-stdv, stdv)

    def parameters(self):
        return [self.weight], [self.gradWeight]

    # single-step protocol (LSTM.lua:100-136): input {x, prev_h, prev_c} -> {next_h, next_c}
    def updateOutput(self, input):
        x = input[0]
        hp = input[1] if len(input) > 1 else None
        cp = input[2] if len(input) > 2 else None
        self._batched = x.dim() == 2
        up = (lambda t: None if t is None else (t.contiguous() if self._batched else t.contiguous().unsqueeze(0)))
        self._x, self._hp, self._cp = up(x), up(hp), up(cp)
        hn, cn, self._acts = ops.lstm_step_forward(self.ctx, self.weight, self._x, self.dimoutput, self._hp, self._cp, self.peepholes)
        self._cn = cn
        self.output = [hn, cn] if self._batched else [hn[0], cn[0]]
        return self.output

    def updateGradInput(self, input, gradOutput):
        up = (lambda t: None if t is None else (t.contiguous() if self._batched else t.contiguous().unsqueeze(0)))
        dhn = up(gradOutput[0]); dcn = up(gradOutput[1]) if len(gradOutput) > 1 else None
        dx, dhp, dcp, _ = ops.lstm_step_backward(self.ctx, self.weight, self._x, self.dimoutput, self._hp, self._cp, self._acts, self._cn, dhn, dcn,
                                                 self.peepholes, dP=self.gradWeight)
        self.gradInput = [dx, dhp, dcp] if self._batched else [dx[0], dhp[0], dcp[0]]
        return self.gradInput


class RNN(Module):
    def __init__(self, ctx, recurrent, reverse=False):
        super().__init__(ctx)
        assert recurrent is not None, "recurrent cannot be nil"
        assert getattr(recurrent, "dimoutput", None) is not None, "recurrent must specify dimoutput"
        self.recurrent, self.dimoutput, self.reverse = recurrent, recurrent.dimoutput, bool(reverse)
        self.modules = [recurrent]

    def parameters(self):
        return self.recurrent.parameters()

    def updateOutput(self, input, lengths=None):
        if input.dim() not in (2, 3):
            raise ops.S2SError("input must be 2d or 3d")
        x = input.contiguous().view(-1, input.shape[-2], input.shape[-1]) if input.dim() == 2 else input.contiguous()
        if isinstance(self.recurrent, LSTM):
            y, self._save = ops.lstm_seq_forward(self.ctx, self.recurrent.weight, x, self.dimoutput, peepholes=self.recurrent.peepholes,
                                                 lengths=lengths, reverse=self.reverse)
        else:
            y, self._save = ops.gru_seq_forward(self.ctx, self.recurrent.weight, x, lengths=lengths, ndir=1, reverse=self.reverse)
        self._y, self._lengths = y, lengths
        self.output = y[0] if input.dim() == 2 else y
        return self.output

    def updateGradInput(self, input, gradOutput):
        assert getattr(self, "_save", None) is not None, "backward called before forward"
        x = input.contiguous().view(-1, input.shape[-2], input.shape[-1])
        dy = gradOutput.contiguous().view(x.shape[0], x.shape[1], -1)
        if isinstance(self.recurrent, LSTM):
            dx, _ = ops.lstm_seq_backward(self.ctx, self.recurrent.weight, x, self._y, self._save, dy, self.dimoutput,
                                          peepholes=self.recurrent.peepholes, lengths=self._lengths, reverse=self.reverse,
                                          dP=self.recurrent.gradWeight)
        else:
            dx, _ = ops.gru_seq_backward(self.ctx, self.recurrent.weight, x, self._y, self._save, dy, lengths=self._lengths, ndir=1,
                                         reverse=self.reverse, dW=self.recurrent.gradWeight)
        self.gradInput = dx[0] if input.dim() == 2 else dx
        return self.gradInput


class Attention(Module):
    def __init__(self, ctx, decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
                 stateDepth, annotationDepth, outputDepth, monoAlignPenalty=False, penalty_lambda=0.0, mlpDepth=64, maxoutWindow=7):
        super().__init__(ctx)
        assert annotationDepth % 2 == 0
        self.cfg = dict(D=1, H=annotationDepth // 2, NL=1, S=scoreDepth, ST=stateDepth, V=outputDepth,
                        K=hybridAttendFeatureMaps or 0, KF=hybridAttendFilterSize or 10, M=mlpDepth, MW=maxoutWindow)
        self.scoreDepth, self.stateDepth, self.annotationDepth, self.outputDepth = scoreDepth, stateDepth, annotationDepth, outputDepth
        self.penalty_lambda = float(penalty_lambda) if monoAlignPenalty else 0.0
        n = ops.param_count(self.cfg)
        self.off = ops.decoder_param_offset(self.cfg)
        self.flat = ctx.zeros(n)
        self.gradFlat = ctx.zeros(n)
        self.reset()

    def _views(self, flat):
        return [flat[off:off + r * c].view(r, c) for off, r, c in ops.param_segments(self.cfg) if off >= self.off]

    def parameters(self):
        return self._views(self.flat), self._views(self.gradFlat)

    def reset(self, stdv=None):
        full = torch.from_numpy(ops.init_params(self.cfg, seed=1234)).to(self.flat.device)
        self.flat.copy_(full)

    @staticmethod
    def _labels(y):
        return y.argmax(dim=-1).to(torch.int32).contiguous()

    def updateOutput(self, input, lengths=None, tlens=None, dropmask=None):
        x, y = input
        if x.dim() not in (2, 3):
            raise ops.S2SError("x must be 2d or 3d")
        self._batched = x.dim() == 3
        h = x.contiguous() if self._batched else x.contiguous().unsqueeze(0)
        yl = self._labels(y if self._batched else y.unsqueeze(0))
        self._args = (h, yl, lengths, tlens, dropmask)
        logp = ops.attention_forward(self.ctx, self.cfg, self.flat, h, yl, lengths=lengths, tlens=tlens, dropmask=dropmask, lam=self.penalty_lambda)
        self.output = logp if self._batched else logp[0]
        return self.output

    def updateGradInput(self, input, gradOutput):
        h, yl, lengths, tlens, dropmask = self._args
        dlogp = gradOutput.contiguous() if self._batched else gradOutput.contiguous().unsqueeze(0)
        dh = ops.attention_backward(self.ctx, self.cfg, self.flat, self.gradFlat, h, yl, dlogp, lengths=lengths, tlens=tlens,
                                    dropmask=dropmask, lam=self.penalty_lambda)
        self.gradInput = [dh if self._batched else dh[0], None]
        return self.gradInput

    def _get(self, what, last):
        B, T = self._args[1].shape
        out = ops.attention_get(self.ctx, what, (B, T, last))
        return out if self._batched else out[0]

    def alpha(self):
        return self._get(ops.GET_ALPHA, self._args[0].shape[1])

    def Ws(self):
        return self._get(ops.GET_WS, self.scoreDepth)

    def penalty(self):
        return self._get(ops.GET_PENALTY, 1)

    def setpenalty(self, penalty):
        self.penalty_lambda = float(penalty)

    def BeamSearch(self, annotations, eos, K, maxseqlength):
        return ops.beam_search(self.ctx, self.cfg, self.flat, annotations.contiguous(), eos, beam=K, maxlen=maxseqlength)


class WeightNoise(Module):
    def __init__(self, ctx, parameters, sigma=1e-3):
        super().__init__(ctx)
        self.sigma = sigma
        self.weight = parameters.clone()
        self.gradWeight = torch.zeros_like(parameters)
        self.sample = torch.empty_like(parameters)
        self.seed = 0

    def Sample(self, eps=None):
        self.seed += 1
        self.sample = ops.weightnoise_sample(self.ctx, self.weight, self.sigma, eps=eps, seed=self.seed)
        return self.sample

    def Mode(self):
        return self.weight

    def updateOutput(self, nll):
        self.output = nll
        return nll

    def accGradParameters(self, input, gradOutput, scale=1.0):
        self.gradWeight.add_(gradOutput)


class AdaptiveWeightNoise(Module):
    """weight = [mu ; log sigma^2]  (AdaptiveWeightNoise.lua:5-56)"""

    def __init__(self, ctx, parameters, lam=1.0, sigma_init=0.075):
        super().__init__(ctx)
        self.n = parameters.numel()
        self.lam = lam
        self.weight = torch.cat([parameters, torch.full_like(parameters, math.log(sigma_init ** 2))])
        self.gradWeight = torch.zeros_like(self.weight)
        self.seed = 0

    def Sample(self, eps=None):
        self.seed += 1
        return ops.awn_sample(self.ctx, self.weight, eps=eps, seed=self.seed)

    def Mode(self):
        return self.weight[:self.n]

    def updateOutput(self, nll):
        self.output = ops.awn_forward(self.ctx, self.weight, self.lam, float(nll))
        return self.output

    def accGradParameters(self, input, gradOutput, scale=1.0):
        self.gradWeight = ops.awn_accgrad(self.ctx, self.weight, gradOutput.contiguous(), self.lam)


class TrainUtils:
    @staticmethod
    def columnNormConstraint(m, maxval=1.0):
        if getattr(m, "weight", None) is None:
            return
        w = m.weight.view(m.weight.shape[0], -1) if m.weight.dim() != 2 else m.weight
        if w.dim() == 2 and ops.rownorm_constraint(m.ctx, w, maxval):
            raise ops.S2SError("found a nan")            # TrainUtils.lua:55-62
