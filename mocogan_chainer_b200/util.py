"""util.py of raahii/mocogan-chainer: the sample-logging side of training (util.py:13-51,89-115).

`log_tensorboard` is the only caller that runs the generator with chainer.config.train == False, i.e. through
F.fixed_batch_normalization with the running statistics.  Here the generator's output is turned
into uint8 clips AND the n x n grid video by ONE device pass (mcg_video_to_uint8 — `videos / 2 + 0.5`, x255, uint8
truncation and `to_grid`'s tiling fused), so the host only receives the frames it writes.  Writers take uint8 (C, H, W)
images: tensorboard's add_image stores uint8 arrays unscaled, which equals what it makes of the reference's [0, 1]
floats.  ffmpeg-based `save_video` is out of scope (no ffmpeg in the image); frames are written as images instead.
"""
import json
import os

import numpy as np

from . import chainer
from . import kernels as K


def to_sequence(video, horizontally=True):
    """util.py:13-28: frames (num, channel, height, width) -> one image with the frames side by side (or stacked)."""
    video = np.asarray(video)
    return np.concatenate(list(video), axis=2 if horizontally else 1)


def to_grid(videos_u8, size):
    """util.py:30-51 on device-resident uint8 clips (t, bs, c, h, w): only used when the clips did not come straight
    from `sample_videos` (which already returns the grid)."""
    import torch
    t, bs, c, h, w = videos_u8.shape
    grid = torch.zeros((t, c, size * h, size * w), dtype=videos_u8.dtype, device=videos_u8.device)
    for i in range(size):
        for j in range(size):
            if i * size + j < bs:
                grid[:, :, i * h:i * h + h, j * w:j * w + w] = videos_u8[:, i * size + j]
    return grid


def sample_videos(image_gen, num, train=False):
    """The body of util.py:92-103 on the device: generate `num` clips with `chainer.config.train = train` (False: fixed
    BatchNorm statistics) and return (videos uint8 (T, num, C, H, W), grid uint8 (T, C, n*H, n*W)) as device tensors."""
    n = int(np.sqrt(num))
    with chainer.using_config('train', train), chainer.no_backprop_mode():
        videos = image_gen(num)[0].data                       # (T, N, C, H, W) view of channels-last storage
    t, bs, c, h, w = videos.shape
    phys = videos.permute(0, 1, 3, 4, 2)
    if not phys.is_contiguous():
        phys = phys.contiguous()
    return K.video_to_uint8(phys.reshape(t * bs, 1, h, w, c), t, bs, True, n)


def log_tensorboard(image_gen, num, video_length, writer):
    """util.py:89-115.  Returns the extension: log(trainer_or_updater) writes four grid frames
    ('{:02d}th frame', indices np.linspace(0, video_length, 4, endpoint=False)) and the first min(num, 10) clips as
    frame strips ('video_{:02d}') under the updater's epoch.  (The reference indexes 10 clips unconditionally and
    fails for num < 10, App. B#13.)"""

    def log(trainer):
        updater = getattr(trainer, "updater", trainer)
        videos, grid = sample_videos(image_gen, num, train=False)
        frames = np.linspace(0, video_length, 4, endpoint=False, dtype=np.int64)
        grid_host = grid[frames.tolist()].cpu().numpy()          # only the four frames that are written cross PCIe
        for i, img in zip(frames, grid_host):
            writer.add_image('{:02d}th frame'.format(int(i)), img, updater.epoch)
        k = min(num, 10)
        vids = videos[:, :k].cpu().numpy()                        # (T, k, C, H, W)
        for i in range(k):
            writer.add_image('video_{:02d}'.format(i), to_sequence(vids[:, i]), updater.epoch)

    return log


class ImageLogWriter(object):
    """A SummaryWriter stand-in with the two methods the reference uses (add_scalar: updater.py:41,60; add_image:
    util.py:107,113) for environments without tensorboard: scalars go to scalars.jsonl, images to PNG files."""

    def __init__(self, logdir):
        self.logdir = str(logdir)
        os.makedirs(self.logdir, exist_ok=True)

    def add_scalar(self, tag, value, step):
        with open(os.path.join(self.logdir, "scalars.jsonl"), "a") as f:
            f.write(json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")

    def add_image(self, tag, img, step):
        from PIL import Image
        img = np.asarray(img)
        if img.dtype != np.uint8:
            img = (np.clip(img, 0, 1) * 255).astype(np.uint8)
        hwc = img.transpose(1, 2, 0)
        name = "%s_%06d.png" % (tag.replace(" ", "_").replace("/", "_"), int(step))
        Image.fromarray(hwc[:, :, 0] if hwc.shape[2] == 1 else hwc).save(os.path.join(self.logdir, name))

    def close(self):
        pass


def make_writer(logdir):
    """tensorboard's SummaryWriter when importable (train.py:103 uses tb_chainer's), else ImageLogWriter."""
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(str(logdir))
    except Exception:   # noqa: BLE001 — tensorboard missing or broken: fall back to plain files
        return ImageLogWriter(logdir)
