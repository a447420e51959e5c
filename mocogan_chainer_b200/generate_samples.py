"""generate_samples.py of raahii/mocogan-chainer on the B200-native generator (generate_samples.py:17-55).

Positional args and flags are the reference's.  As in the reference the generator runs with chainer.config.train ==
True, i.e. BatchNorm uses BATCH statistics at inference; dim_zl is inferred from the npz because the
reference's default-constructed generator cannot load MUG-trained weights (App. B#10).  mp4/jpg writing needs ffmpeg
(absent here): the uint8 videos (t, n, c, h, w) are saved as videos.npy instead."""
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


def generate(gen, num):
    """videos = gen(num)[0].data; ((v / 2 + 0.5) * 255).astype(uint8)  — generate_samples.py:37-39."""
    import torch
    with chainer.no_backprop_mode():
        videos = gen(num)[0].data                       # (t, bs, c, h, w)
    return ((videos.float() / 2. + 0.5) * 255).to(torch.uint8)


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
    videos = generate(gen, args.num).cpu().numpy()
    print(">>> saving...")
    save_path = Path(args.save_path)
    save_path.mkdir(parents=True, exist_ok=True)
    np.save(save_path / 'videos.npy', videos)
    return videos


if __name__ == "__main__":
    main()
