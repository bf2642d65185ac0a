"""Shared plumbing of the plugin mirrors: one device context per process, a per-scene descriptor
cache so that each cloud's [5000,32,60] descriptor is read from disk and uploaded once (the reference
re-reads it three times per PAIR: test/matcher.py:66-67, test/estimator.py:106-107,:334-335)."""
import os
import numpy as np
import torch
from .. import ops

_ctx = {}


def context(cfg=None):
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    so3 = getattr(cfg, "SO3_related_files", None) if cfg is not None else None
    key = (dev, so3)
    if key not in _ctx:
        _ctx[key] = ops.Context(dev, so3_dir=so3)
        # extension knob (absent from the reference's parser -> default 0 = float32 reference arithmetic under Test.py):
        # 3 = the tcgen05 fp16 two-accumulator Gram for Des2R / the R-indicator (same argmax outside float32 near ties)
        cm = getattr(cfg, "corr_mode", None) if cfg is not None else None
        if cm is not None:
            _ctx[key].set_corr_mode(int(cm))
    return _ctx[key]


def make_non_exists_dir(fn):
    """utils/utils.py:9-11"""
    if not os.path.exists(fn):
        os.makedirs(fn)


def feature_dataset_name(dataset):
    """3dLomatch reuses the 3dmatch per-cloud directories (test/matcher.py:56-59)."""
    if dataset.name[0:4] == '3dLo':
        return f'3d{dataset.name[4:]}'
    return dataset.name


class CloudCache:
    """LRU of device-resident per-cloud tensors keyed by file path."""

    def __init__(self, ctx, capacity=64):
        self.ctx, self.capacity, self.d = ctx, capacity, {}

    def get(self, path, dtype=torch.float32):
        t = self.d.pop(path, None)
        if t is None:
            t = self.ctx.dev(np.load(path), dtype)
            if len(self.d) >= self.capacity:
                self.d.pop(next(iter(self.d)))
        self.d[path] = t
        return t
