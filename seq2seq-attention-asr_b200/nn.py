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
    nn.Maxout / nn.Linear / nn.Dropout / nn.Sequential    shape holders for the decoder_mlp graph nn.Attention inspects
    TrainUtils.{apply2graph, columnNormConstraint(Graph), orthogonalize(Graph), optimConfigResets}   TrainUtils.lua:5-213, timit.lua:496-502

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
        self.modules = None          # children TrainUtils.apply2graph descends into (TrainUtils.lua:142-145)

    def findModules(self, cls):
        out = [self] if isinstance(self, cls) else []
        for m in self.modules or []:
            out += m.findModules(cls)
        return out

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
        for m in self.modules or []:
            m.training()

    def evaluate(self):
        self.train = False
        for m in self.modules or []:
            m.evaluate()


class Param(Module):
    """Leaf giving a slice of a flat parameter vector the .weight / .bias / .gradWeight / .gradBias fields that
    TrainUtils.apply2graph looks for (lua/s2s_ffi.lua nn.S2SParam)."""

    def __init__(self, ctx, name, weight, gradWeight, bias=None, gradBias=None):
        super().__init__(ctx)
        self.name, self.weight, self.gradWeight, self.bias, self.gradBias = name, weight, gradWeight, bias, gradBias

    def parameters(self):
        if self.bias is not None:
            return [self.weight, self.bias], [self.gradWeight, self.gradBias]
        return [self.weight], [self.gradWeight]


def _bound(stdv, fan_in):
    """reset(stdv) rule of LinearZeroBias.lua:12-29 / TemporalConvolutionZeroBias.lua:21-35: U(+-stdv sqrt 3) or U(+-1/sqrt(fan_in))"""
    return stdv * math.sqrt(3.0) if stdv else 1.0 / math.sqrt(fan_in)


# shape holders for the sub-graphs the model files hand to nn.Attention (model_chorowski_baseline.lua:53-59): they carry sizes and
# initial values only -- the decoder MLP runs time-batched inside the library
class Linear(Module):
    def __init__(self, ctx, inputSize, outputSize):
        super().__init__(ctx)
        b = 1.0 / math.sqrt(inputSize)
        self.weight = ctx.new(outputSize, inputSize).uniform_(-b, b)
        self.bias = ctx.new(outputSize).uniform_(-b, b)


class Maxout(Module):
    """Maxout.lua:5-21: Linear(in, out*window) -> max over groups of `window` consecutive units"""

    def __init__(self, ctx, inputDimension, outputDimension, window=4):
        super().__init__(ctx)
        assert inputDimension is not None, "must specify input dimension"
        assert outputDimension is not None, "must specify output dimension"
        self.inputDim, self.outputDim, self.window = inputDimension, outputDimension, window
        self.linear = Linear(ctx, inputDimension, outputDimension * window)
        self.modules = [self.linear]


class Dropout(Module):
    def __init__(self, ctx, p=0.5):
        super().__init__(ctx)
        self.p = p


class LogSoftMax(Module):
    pass


class Sequential(Module):
    def __init__(self, ctx, *mods):
        super().__init__(ctx)
        self.modules = list(mods)

    def add(self, m):
        self.modules.append(m)
        return self


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
        b = _bound(stdv, self.inputFrameSize)
        self.weight.uniform_(-b, b)
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
        # the three LinearZeroBias leaves of the reference's GRU graph (GRU.lua:23-26): row-norm constraint / orthogonalize act per gate
        self.modules = [Param(ctx, "GRU." + g, self.weight[i], self.gradWeight[i]) for i, g in enumerate("zrh")]
        self.reset()

    def reset(self, stdv=None):
        b = _bound(stdv, self.dimoutput + self.diminput)
        self.weight.uniform_(-b, b)

    def parameters(self):
        return [m.weight for m in self.modules], [m.gradWeight for m in self.modules]

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
        # the nn.Linear leaves of the reference's LSTM graph (LSTM.lua:25-36), as views of the flat vector
        self.modules = []
        for name, (wo, rows, cols), bo in ops.lstm_segments(diminput, dimoutput, self.peepholes):
            self.modules.append(Param(ctx, "LSTM." + name, self.weight[wo:wo + rows * cols].view(rows, cols), self.gradWeight[wo:wo + rows * cols].view(rows, cols),
                                      self.weight[bo:bo + rows], self.gradWeight[bo:bo + rows]))
        self.reset()

    def reset(self, stdv=None):
        for m in self.modules:                       # stock nn.Linear.reset: weight and bias U(+-1/sqrt(fan_in))
            b = _bound(stdv, m.weight.shape[1])
            m.weight.uniform_(-b, b); m.bias.uniform_(-b, b)

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
    """nn.Attention (Attention.lua:15-24,214-438).  decoder_recurrent must contain one GRU(stateDepth, stateDepth);
    decoder_mlp one or two Maxout stages (each followed by a Linear), optionally an nn.Dropout in front: their sizes select
    cfg.M / cfg.MW / cfg.MLP and their weights initialise the corresponding segments (lua/Attention.lua does the same)."""

    SEGS = ("WV", "bV", "Ws", "bs", "WF", "bF", "U", "bU", "we", "be", "Wy", "by", "Wc", "bc", "Wj", "bj", "Gz", "Gr", "Gh", "Wm", "bm",
            "Wl", "bl", "Wm2", "bm2", "Wo", "bo")

    def __init__(self, ctx, decoder_recurrent, decoder_mlp, scoreDepth, hybridAttendFilterSize, hybridAttendFeatureMaps,
                 stateDepth, annotationDepth, outputDepth, monoAlignPenalty=False, penalty_lambda=0.0):
        super().__init__(ctx)
        assert annotationDepth % 2 == 0
        grus = decoder_recurrent.findModules(GRU)
        if len(grus) != 1 or decoder_recurrent.findModules(LSTM):
            raise ops.S2SError("nn.Attention (libs2s_b200): decoder_recurrent must wrap exactly one nn.GRU (LSTM decoders are not built)")
        if grus[0].diminput != stateDepth or grus[0].dimoutput != stateDepth:
            raise ops.S2SError("nn.Attention: decoder GRU must be GRU(stateDepth, stateDepth)")
        maxouts = decoder_mlp.findModules(Maxout)
        inner = [m.linear for m in maxouts]
        linears = [l for l in decoder_mlp.findModules(Linear) if all(l is not i for i in inner)]
        drops = decoder_mlp.findModules(Dropout)
        if len(maxouts) not in (1, 2) or len(linears) != len(maxouts):
            raise ops.S2SError("nn.Attention (libs2s_b200): decoder_mlp must be Maxout-Linear or Maxout-Linear-Maxout-Linear")
        if maxouts[0].inputDim != stateDepth + annotationDepth:
            raise ops.S2SError("nn.Attention: first Maxout must read {s, c}")
        self.dropout = drops[0] if drops else None
        self.cfg = dict(D=1, H=annotationDepth // 2, NL=0, S=scoreDepth, ST=stateDepth, V=outputDepth,
                        K=hybridAttendFeatureMaps or 0, KF=hybridAttendFilterSize or 10, M=maxouts[0].outputDim, MW=maxouts[0].window,
                        MLP=len(maxouts))
        self.scoreDepth, self.stateDepth, self.annotationDepth, self.outputDepth = scoreDepth, stateDepth, annotationDepth, outputDepth
        self.monoAlignPenalty = bool(monoAlignPenalty)
        self.penalty_lambda = float(penalty_lambda) if monoAlignPenalty else 0.0
        n = ops.param_count(self.cfg)
        self.flat = ctx.zeros(n)
        self.gradFlat = ctx.zeros(n)
        names = [s for s in self.SEGS if (s not in ("WF", "bF", "U", "bU") or self.cfg["K"] > 0) and (s not in ("Wl", "bl", "Wm2", "bm2") or self.cfg["MLP"] == 2)]
        segs = ops.param_segments(self.cfg)
        assert len(segs) == len(names)
        self._p, self._g = {}, {}
        for name, (off, r, c) in zip(names, segs):
            v = (lambda t: t[off:off + r * c].view(r, c) if c > 1 else t[off:off + r * c])
            self._p[name], self._g[name] = v(self.flat), v(self.gradFlat)
        self._order = names
        pairs = [("WV", "bV"), ("Ws", "bs")] + ([("WF", "bF"), ("U", "bU")] if self.cfg["K"] > 0 else []) + \
                [("we", "be"), ("Wy", "by"), ("Wc", "bc"), ("Wj", "bj"), ("Gz", None), ("Gr", None), ("Gh", None), ("Wm", "bm")] + \
                ([("Wl", "bl"), ("Wm2", "bm2")] if self.cfg["MLP"] == 2 else []) + [("Wo", "bo")]
        self.modules = [Param(ctx, w, self._p[w], self._g[w], self._p.get(b), self._g.get(b)) for w, b in pairs]
        self.reset()
        # adopt the values the caller's modules were constructed with
        for name, src in zip(("Gz", "Gr", "Gh"), grus[0].modules):
            self._p[name].copy_(src.weight)
        self._p["Wm"].copy_(maxouts[0].linear.weight); self._p["bm"].copy_(maxouts[0].linear.bias)
        if self.cfg["MLP"] == 2:
            self._p["Wl"].copy_(linears[0].weight); self._p["bl"].copy_(linears[0].bias)
            self._p["Wm2"].copy_(maxouts[1].linear.weight); self._p["bm2"].copy_(maxouts[1].linear.bias)
        self._p["Wo"].copy_(linears[-1].weight); self._p["bo"].copy_(linears[-1].bias)
        self._dropseed = 0

    def parameters(self):
        return [self._p[k] for k in self._order], [self._g[k] for k in self._order]

    def reset(self, stdv=None):
        for m in self.modules:
            w = m.weight
            fan_in = self.cfg["KF"] if m.name == "WF" else (w.shape[1] if w.dim() == 2 else w.shape[0])
            b = _bound(stdv, fan_in)
            w.uniform_(-b, b)
            if m.bias is not None:
                if m.name in ("WV", "U", "we"):
                    m.bias.zero_()                    # TemporalConvolutionZeroBias.lua:34
                else:
                    m.bias.uniform_(-b, b)

    def updateOutput(self, input, lengths=None, tlens=None, dropmask=None):
        x, y = input
        if x.dim() not in (2, 3):
            raise ops.S2SError("x must be 2d or 3d")
        self._batched = x.dim() == 3
        h = x.contiguous() if self._batched else x.contiguous().unsqueeze(0)
        yl = ops.labels_from_onehot(self.ctx, (y if self._batched else y.unsqueeze(0)).contiguous())     # labelmask -> labels on the device
        B, T = yl.shape
        if dropmask is None and self.dropout is not None and self.train and self.dropout.p > 0:          # nn.Dropout v2, training mode only
            self._dropseed += 1
            dropmask = ops.dropout_mask(self.ctx, (B, T, self.stateDepth + self.annotationDepth), self.dropout.p, seed=self._dropseed)
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

    @property
    def Vh(self):
        """decoder.Vh.output (timit/timit.lua:521)"""
        h = self._args[0]
        out = ops.attention_get(self.ctx, ops.GET_VH, (h.shape[0], h.shape[1], self.scoreDepth))
        return type("VhNode", (), {"output": out if self._batched else out[0]})()

    def setpenalty(self, penalty):
        if not self.monoAlignPenalty:
            raise ops.S2SError("could not find penalty node")         # Attention.lua:263
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
    """weight = [mu ; log sigma^2]  (AdaptiveWeightNoise.lua:5-56); sigma_init defaults to 1 as in the reference (:13)"""

    def __init__(self, ctx, parameters, lam=1.0, sigma_init=1.0):
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
    """TrainUtils.lua:202-213 -- graph walkers + per-module weight post-processing (the arithmetic is in the library)"""

    @staticmethod
    def apply2graph(graph, func):
        """visit every leaf below `graph` (TrainUtils.lua:137-184): modules with children -> descend, else func(leaf)"""
        kids = getattr(graph, "modules", None)
        if kids:
            for m in kids:
                TrainUtils.apply2graph(m, func)
        else:
            func(graph)

    @staticmethod
    def columnNormConstraint(m, maxval=1.0):
        """every ROW of m.weight with L2 norm >= maxval is divided by norm / maxval (TrainUtils.lua:52-104: norm(2,2) per row)"""
        w = getattr(m, "weight", None)
        if w is None:
            return
        if w.dim() != 2:
            raise ops.S2SError("columnNormConstraint: apply it to the 2-D leaves (TrainUtils.columnNormConstraintGraph walks them); "
                               f"got a {w.dim()}-D weight on {type(m).__name__}")
        if ops.rownorm_constraint(m.ctx, w, maxval):
            raise ops.S2SError("found a nan")            # TrainUtils.lua:55-62

    @staticmethod
    def columnNormConstraintGraph(graph, maxval=1.0):     # timit/timit.lua:346-348
        TrainUtils.apply2graph(graph, lambda m: TrainUtils.columnNormConstraint(m, maxval))

    @staticmethod
    def orthogonalize(m):
        """m.weight (with m.bias as one more column) := orthonormal QR factor in the tall orientation (TrainUtils.lua:5-26)"""
        w = getattr(m, "weight", None)
        if w is None or w.dim() != 2:
            return
        b = getattr(m, "bias", None)
        ops.orthogonalize(m.ctx, w, b if (b is not None and b.numel() == w.shape[0]) else None)

    @staticmethod
    def orthogonalizeGraph(graph):                        # librispeech/exp0_scriptchecker.lua:49-52
        TrainUtils.apply2graph(graph, TrainUtils.orthogonalize)

    @staticmethod
    def optimConfigResets(optimConfig, resets, epoch):
        """the epoch-indexed optimiser schedule of timit/timit.lua:496-502: at the start of `epoch`, entries of resets[epoch]
        overwrite optimConfig in place"""
        if resets and epoch in resets:
            optimConfig.update(resets[epoch])
        return optimConfig
