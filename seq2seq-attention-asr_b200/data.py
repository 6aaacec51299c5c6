"""Data path of the training scripts: the reference's HDF5 corpora -> length-bucketed minibatches on the device.

The reference loads the whole file with torch-hdf5 (`hdf5.open(opt.datafile):all()`, timit/timit.lua:42-43), moves every utterance to
the GPU (`:cuda()`, :44-60) and then trains ONE UTTERANCE AT A TIME (timit.lua:240-289), building the label mask per utterance
(`labelmask = one-hot(Y)`, :262).  Here the same files are read (h5.py: the image has no libhdf5), utterances are grouped into
minibatches of similar length (padding wastes recurrence steps: a batch runs max(L_b) frames) and each batch goes to the device as
padded features + int32 labels + per-utterance lengths -- the inputs of s2s_model_fwdbwd; the label mask is generated on the
device by s2s_onehot when a caller wants the reference's {X, labelmask} pair.

    ds = TimitH5("timit_logmel.h5", "train")          # /train/<k>/{x, y, y39, start, finish}   preprocess_timit.py:356-363
    for batch in BucketedBatches(ds, batch_size=32, shuffle=True, seed=epoch):
        X, labels, lengths, tlens = batch.to_device(ctx)
"""
import numpy as np

from . import h5


class H5Utterances:
    """utterance store over one group of an HDF5 file: group -> {<k>: {x, y, ...}}"""

    def __init__(self, path, group="/", x="x", y="y", label_offset=0):
        self.f = h5.File(path)
        self.group = "/" + group.strip("/")
        self.keys = sorted(self.f.keys(self.group), key=lambda k: (len(k), k))      # "0", "1", ..., "10": numeric order
        self.xname, self.yname, self.label_offset = x, y, label_offset
        self._len = None

    def __len__(self):
        return len(self.keys)

    def __getitem__(self, i):
        base = f"{self.group}/{self.keys[i]}".replace("//", "/")
        x = np.asarray(self.f[f"{base}/{self.xname}"], dtype=np.float32)            # stored as float64 (preprocess_timit.py:275)
        y = np.asarray(self.f[f"{base}/{self.yname}"]).astype(np.int32).reshape(-1) - self.label_offset
        return x, y

    def lengths(self):
        """(frames, labels) of every utterance; read once (datasets are small next to the features, but the scan touches every header)"""
        if self._len is None:
            out = np.zeros((len(self), 2), dtype=np.int64)
            for i in range(len(self)):
                x, y = self[i]
                out[i] = (x.shape[-2] if x.ndim >= 2 else x.shape[0], y.shape[0])
            self._len = out
        return self._len


def TimitH5(path, split="train"):
    """timit/preprocess_timit.py:356-363: /<split>/<k>/{x [L,123], y [T] (0-based phoneme ids, EOS appended), y39, start, finish}"""
    return H5Utterances(path, split, "x", "y")


def LibriH5(path):
    """librispeech/preprocess.py:230-236: /<i>/{x, chars, words}; character labels are the targets (train.lua:97-101)"""
    return H5Utterances(path, "/", "x", "chars")


class Batch:
    def __init__(self, X, labels, lengths, tlens, index):
        self.X, self.labels, self.lengths, self.tlens, self.index = X, labels, lengths, tlens, index

    @property
    def frames(self):
        return int(self.lengths.sum())

    def to_device(self, ctx, pinned=None):
        """H2D of the padded batch (pinned staging when given); returns torch CUDA tensors"""
        import torch
        dev = ctx.device
        out = []
        for a in (self.X, self.labels, self.lengths, self.tlens):
            t = torch.from_numpy(np.ascontiguousarray(a))
            out.append(t.pin_memory().to(dev, non_blocking=True) if pinned else t.to(dev))
        return tuple(out)

    def labelmask(self, ctx, V):
        """the reference's one-hot label mask [B, T, V] (timit/timit.lua:262), generated on the device"""
        from . import ops
        import torch
        lab = torch.from_numpy(self.labels).to(ctx.device)
        return ops.onehot(ctx, lab, V)


def pad_batch(items, index, eos=None):
    """features zero-padded to the longest utterance (padding frames are ignored through `lengths`), labels padded with `eos`
    (ignored through `tlens`)"""
    B = len(items)
    Lmax = max(x.shape[0] for x, _ in items); Tmax = max(len(y) for _, y in items)
    D = items[0][0].shape[1]
    X = np.zeros((B, Lmax, D), dtype=np.float32)
    labels = np.full((B, Tmax), 0 if eos is None else eos, dtype=np.int32)
    lengths = np.zeros(B, dtype=np.int32); tlens = np.zeros(B, dtype=np.int32)
    for b, (x, y) in enumerate(items):
        X[b, :x.shape[0]] = x; labels[b, :len(y)] = y
        lengths[b] = x.shape[0]; tlens[b] = len(y)
    return Batch(X, labels, lengths, tlens, np.asarray(index))


class BucketedBatches:
    """Minibatches of utterances of similar length.  Utterances are sorted by frame count, cut into consecutive batches, and the
    ORDER OF THE BATCHES (not their composition) is shuffled per epoch -- the reference shuffles single utterances (torch.randperm,
    timit.lua:209) because its batch is a Python-style loop; a padded batch wants equal lengths.  `world`/`rank` give every
    data-parallel rank the same batch boundaries and a contiguous shard of each batch (dp.shard_bounds)."""

    def __init__(self, ds, batch_size, shuffle=True, seed=0, world=1, rank=0, eos=None, max_utterances=None):
        self.ds, self.bs, self.shuffle, self.seed, self.world, self.rank, self.eos = ds, batch_size, shuffle, seed, world, rank, eos
        lens = ds.lengths()[:, 0]
        n = len(lens) if max_utterances is None else min(len(lens), max_utterances)       # opt.maxnumsamples (timit.lua:207)
        order = np.argsort(lens[:n], kind="stable")
        self.batches = [order[i:i + batch_size] for i in range(0, n, batch_size)]

    def __len__(self):
        return len(self.batches)

    def padding_fraction(self):
        lens = self.ds.lengths()[:, 0]
        tot = sum(len(b) * lens[b].max() for b in self.batches)
        return 1.0 - sum(lens[b].sum() for b in self.batches) / tot

    def __iter__(self):
        from .dp import shard_bounds
        ids = np.arange(len(self.batches))
        if self.shuffle:
            np.random.default_rng(self.seed).shuffle(ids)
        for i in ids:
            idx = self.batches[i]
            lo, hi = shard_bounds(len(idx), self.world, self.rank)
            idx = idx[lo:hi]
            if len(idx) == 0:
                continue
            yield pad_batch([self.ds[j] for j in idx], idx, self.eos)


def write_timit_like(path, splits, seed=0, D=123, V=62):
    """a synthetic corpus in the layout of preprocess_timit.py:356-363 (zero-mean unit-variance features with 10 zero frames at each end,
    :274-276; labels with EOS = V - 1 appended): splits = {"train": [(L, T), ...], ...}"""
    rng = np.random.default_rng(seed)
    tree = {}
    for split, utts in splits.items():
        g = {}
        for k, (L, T) in enumerate(utts):
            x = rng.standard_normal((L, D))
            x[:10] = 0; x[-10:] = 0
            y = np.concatenate([rng.integers(0, V - 1, T - 1), [V - 1]]).astype(np.int64)
            start = np.sort(rng.integers(0, L, T - 1)).astype(np.int64)
            g[str(k)] = {"x": x, "y": y, "y39": (y % 39).astype(np.int64), "start": start, "finish": start + 1}
        tree[split] = g
    h5.write(path, tree)
    return tree
