"""chainer.serializers.save_npz / load_npz with Chainer's '/'-joined key schema, so that
`image_gen_epoch_N.npz` files are interchangeable with the reference's (train.py:139-144,190-192)."""
import numpy as np
import torch

from . import Parameter


def _collect(target):
    out = {}

    def ser(key, value):
        if isinstance(value, Parameter):
            out[key] = value.data.detach().float().cpu().numpy()
        else:
            obj, attr = value
            v = getattr(obj, attr)
            out[key] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)

    target.serialize(ser)
    return out


def save_npz(file, obj, compression=True):
    d = _collect(obj)
    (np.savez_compressed if compression else np.savez)(str(file), **d)


def load_npz(file, obj, path="", strict=True):
    with np.load(str(file)) as f:
        data = {k: f[k] for k in f.files}

    def de(key, value):
        k = path + key
        if k not in data:
            if strict:
                raise KeyError("%s not found in %s" % (k, file))
            return
        a = data[k]
        if isinstance(value, Parameter):
            if tuple(a.shape) != tuple(value.shape):
                raise ValueError("shape mismatch for %s: file %s, model %s" % (k, a.shape, value.shape))
            value.data = a.astype(np.float32)
        else:
            target, attr = value
            cur = getattr(target, attr)
            if torch.is_tensor(cur):
                cur.copy_(torch.from_numpy(np.asarray(a)).to(cur.dtype))
            elif isinstance(cur, np.ndarray):
                setattr(target, attr, a.astype(cur.dtype))
            else:
                setattr(target, attr, type(cur)(a))

    obj.serialize(de)
    if hasattr(obj, "_arena") and obj._arena is not None:
        obj._arena.refresh_bf16()
