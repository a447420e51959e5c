"""chainer.serializers.save_npz / load_npz with Chainer's '/'-joined key schema, so that
`image_gen_epoch_N.npz` files are interchangeable with the reference's (train.py:139-144,190-192), and — through the
same serializer objects — the full-trainer `snapshot_epoch_N.npz` of `extensions.snapshot` (train.py:137-138) that
`--resume` reads back (train.py:162-163): keys `updater/model:<name>/...`, `updater/optimizer:<name>/<param>/{m,v,t}`,
`updater/optimizer:<name>/{t,epoch}`, `updater/iterator:main/...`, `updater/iteration`.

A serializer is called as `serializer(key, value)` and indexed as `serializer['child']` (Chainer's
DictionarySerializer / NpzDeserializer protocol); `value` is a Parameter, a torch tensor (saved / restored in place),
or an `(object, attribute_name)` pair for scalars and persistents."""
import numpy as np
import torch

from . import Parameter


class _Saver(object):
    is_saver = True

    def __init__(self, target, path=""):
        self.target, self.path = target, path

    def __getitem__(self, key):
        return _Saver(self.target, self.path + key.strip("/") + "/")

    def __call__(self, key, value):
        k = self.path + key.lstrip("/")
        if isinstance(value, Parameter):
            self.target[k] = value.data.detach().float().cpu().numpy()
        elif torch.is_tensor(value):
            self.target[k] = value.detach().cpu().numpy()
        else:
            obj, attr = value
            v = getattr(obj, attr)
            self.target[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
        return value


class _Loader(object):
    is_saver = False

    def __init__(self, data, path="", strict=True, source=""):
        self.data, self.path, self.strict, self.source = data, path, strict, source

    def __getitem__(self, key):
        return _Loader(self.data, self.path + key.strip("/") + "/", self.strict, self.source)

    def __call__(self, key, value):
        k = self.path + key.lstrip("/")
        if k not in self.data:
            if self.strict:
                raise KeyError("%s not found in %s" % (k, self.source))
            return value
        a = self.data[k]
        if isinstance(value, Parameter):
            if tuple(a.shape) != tuple(value.shape):
                raise ValueError("shape mismatch for %s: file %s, model %s" % (k, a.shape, value.shape))
            value.data = a.astype(np.float32)
        elif torch.is_tensor(value):
            if tuple(a.shape) != tuple(value.shape):
                raise ValueError("shape mismatch for %s: file %s, target %s" % (k, a.shape, tuple(value.shape)))
            value.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(value.dtype))
        else:
            target, attr = value
            cur = getattr(target, attr)
            if torch.is_tensor(cur):
                cur.copy_(torch.from_numpy(np.asarray(a)).to(cur.dtype))
            elif isinstance(cur, np.ndarray):
                setattr(target, attr, a.astype(cur.dtype))
            elif isinstance(cur, bool):
                setattr(target, attr, bool(a))
            else:
                setattr(target, attr, type(cur)(a))
        return value


def _collect(target):
    out = {}
    target.serialize(_Saver(out))
    return out


def save_npz(file, obj, compression=True):
    d = _collect(obj)
    (np.savez_compressed if compression else np.savez)(str(file), **d)


def load_npz(file, obj, path="", strict=True):
    with np.load(str(file)) as f:
        data = {k: f[k] for k in f.files}
    obj.serialize(_Loader(data, path, strict, str(file)))
    for link in _links_of(obj):
        if getattr(link, "_arena", None) is not None:
            link._arena.refresh_bf16()


def _links_of(obj):
    """The models whose bf16 weight copies must follow a load: the object itself, or an updater's / trainer's models."""
    if hasattr(obj, "_arena"):
        return [obj]
    up = getattr(obj, "updater", obj)
    opts = getattr(up, "_optimizers", None)
    return [o.target for o in opts.values()] if opts else []
