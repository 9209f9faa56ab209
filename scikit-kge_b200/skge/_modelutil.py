"""Shared marshalling for the model hooks: host lists -> int32 SoA device tensors."""
import numpy as np
import torch

from . import _ext


def idx_tensor(a):
    if isinstance(a, torch.Tensor):
        return a.to(device=_ext.device(), dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.int32)).to(_ext.device())


def unzip_device(xys, with_ys=False):
    """[((s, o, p), y), ...] -> (s, o, p[, y]) CUDA tensors.  One host pass, one
    upload (the reference re-zips python tuples per batch: skge/util.py:104-110)."""
    a = np.array([x for x, _ in xys], dtype=np.int32).reshape(-1, 3)
    t = torch.from_numpy(np.ascontiguousarray(a.T)).to(_ext.device())
    out = (t[0].contiguous(), t[1].contiguous(), t[2].contiguous())
    if with_ys:
        y = torch.from_numpy(np.array([y for _, y in xys], dtype=np.float32)).to(_ext.device())
        return out + (y,)
    return out


def updater_args(updaters, first, second):
    """(opt code, lr, p2 of first, p2 of second) from the trainer's updaters."""
    u1, u2 = updaters[first], updaters[second]
    return u1.opt_code, float(u1.learning_rate), u1._state(), u2._state()
