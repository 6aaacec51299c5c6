"""seq2seq-attention-asr_b200 -- B200 (sm_100a) implementation of the seq2seq attention-ASR training
hot path of Ajay-Wong/seq2seq-attention-asr behind a C ABI (include/s2s_b200.h).

Python is only the host-side mirror of the reference's Lua module surface (the image has no
LuaJIT/Torch7): `ops` marshals torch CUDA tensors into the C ABI, `nn` mirrors the Torch7 module
protocol (updateOutput / updateGradInput / accGradParameters) of Attention.lua, RNN.lua, GRU.lua, ...
`data` / `h5` read the reference's HDF5 corpora and bucket them, `t7` reads / writes Torch7 checkpoints and `log.h5`.
The directory name contains hyphens; import it with
    importlib.import_module("seq2seq-attention-asr_b200")      or      import s2s_b200
"""
from . import _lib, data, dp, h5, nn, ops, t7  # noqa: F401
from ._lib import LIB_PATH, ModelCfg, S2SError, declared_symbols, load  # noqa: F401
from .ops import *  # noqa: F401,F403
from .ops import CHOROWSKI_TIMIT, Context  # noqa: F401
