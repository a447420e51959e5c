"""generate_samples.py of raahii/mocogan-chainer on the B200-native generator (generate_samples.py:17-55).

Positional args and flags are the reference's.  As in the reference the generator runs with chainer.config.train ==
True, i.e. BatchNorm uses BATCH statistics at inference; dim_zl is inferred from the npz because the
reference's default-constructed generator cannot load MUG-trained weights (App. B#10).  mp4/jpg writing needs ffmpeg
(absent here): the uint8 videos (t, n, c, h, w) and the n x n grid video are saved as videos.npy / grid.npy instead."""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "mocogan_chainer_b200"

from . import chainer  # noqa: E402
from .model.net import ImageGenerator  # noqa: E402


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('model_weight')
    parser.add_argument('save_path')
    parser.add_argument('--num', '-n', type=int, default=36)
    parser.add_argument('--gpu', '-g', type=int, default=-1)
    return parser


def generate(gen, num, grid=True):
    """videos = gen(num)[0].data; ((v / 2 + 0.5) * 255).astype(uint8) (generate_samples.py:37-39) and the n x n grid
    video of util.py:30-51 (`to_grid`), both produced by ONE pass over the generator's output (mcg_video_to_uint8).
    Returns (videos uint8 (t, bs, c, h, w), grid uint8 (t, c, n*h, n*w) | None)."""
    from . import kernels as K
    with chainer.no_backprop_mode():
        videos = gen(num)[0].data                       # (t, bs, c, h, w), a view of channels-last storage
    t, bs, c, h, w = videos.shape
    phys = videos.permute(0, 1, 3, 4, 2)                # (t, bs, h, w, c): the storage order
    if not phys.is_contiguous():
        phys = phys.contiguous()
    n = int(round(np.sqrt(num)))
    return K.video_to_uint8(phys.reshape(t * bs, 1, h, w, c), t, bs, True, n if grid else 0)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if np.sqrt(args.num) % 1.0 != 0:
        raise ValueError('--num must be n^2 (n: natural number).')
    with np.load(args.model_weight) as f:
        in_size = f['g0/W_r/W'].shape[1]
        dim_zm = f['g0/W_r/W'].shape[0]
        n_hidden, c8 = f['dc1/W'].shape[0], f['dc1/W'].shape[1]
        out_channels = f['dc5/W'].shape[1]
    gen = ImageGenerator(dim_zc=n_hidden - dim_zm, dim_zm=dim_zm, dim_zl=in_size - dim_zm, out_channels=out_channels,
                         n_filters=c8 // 8)
    chainer.serializers.load_npz(args.model_weight, gen)
    print(">>> generating...")
    videos, grid = generate(gen, args.num)
    videos, grid = videos.cpu().numpy(), grid.cpu().numpy()
    print(">>> saving...")
    save_path = Path(args.save_path)
    save_path.mkdir(parents=True, exist_ok=True)
    np.save(save_path / 'videos.npy', videos)
    np.save(save_path / 'grid.npy', grid)               # (t, c, n*h, n*w): what the reference encodes as grid.mp4
    return videos


if __name__ == "__main__":
    main()
