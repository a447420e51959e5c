"""train.py of raahii/mocogan-chainer on the B200-native step (train.py:24-195).

Every reference flag is kept with its default (train.py:26-44), including --n_filters_idis/--n_filters_vdis which the
reference parses but never uses (App. B#4).  Additive flags: --synthetic, --dtype, --seed, --graph, --max_iter, --no_clip_cache.
The Chainer Trainer/extension plumbing is replaced by a plain loop with the same epoch-triggered actions: LogReport's
`log` file and the PrintReport line (train.py:146-151), model / trainer snapshots with the reference's file names
(train.py:135-144), and log_tensorboard's sample grids (train.py:154-160, util.py:89-115)."""
import json
import argparse
import os
import sys
import time
from datetime import datetime, timedelta, timezone
from pathlib import Path

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "mocogan_chainer_b200"

from . import chainer  # noqa: E402
from . import parallel  # noqa: E402
from . import random as mrandom  # noqa: E402
from . import util  # noqa: E402
from .datasets import MovingMnistDataset, MugDataset, SyntheticClipDataset  # noqa: E402
from .model.net import ImageDiscriminator, ImageGenerator, VideoDiscriminator  # noqa: E402
from .model.updater import Updater  # noqa: E402


def build_parser():
    parser = argparse.ArgumentParser(description='Train script')
    parser.add_argument('--gpu', '-g', type=int, default=-1, help='GPU ID (negative value indicates CPU)')
    parser.add_argument('--dataset_type', choices=['mug', 'mnist'], default='mug', help="dataset type")
    parser.add_argument('--dataset', default='data/dataset/train', help="dataset root path")
    parser.add_argument('--batchsize', type=int, default=100, help="batchsize")
    parser.add_argument('--max_epoch', type=int, default=1000, help="num learning epochs")
    parser.add_argument('--model', type=str, choices=['normal', 'cgan', 'infogan'], default="normal", help="MoCoGAN model")
    jst = timezone(timedelta(hours=9))
    parser.add_argument('--save_name', default=datetime.now(jst).strftime("%Y_%m%d_%H%M"),
                        help="save path for log, snapshot etc")
    parser.add_argument('--display_interval', type=int, default=1, help='interval of displaying log to console')
    parser.add_argument('--snapshot_interval', type=int, default=10, help='interval of snapshot')
    parser.add_argument('--log_tensorboard_interval', type=int, default=10,
                        help='interval of log to tensorboard (genenrate samples too)')
    parser.add_argument('--num_gen_samples', type=int, default=36, help='num generate samples')
    parser.add_argument('--dim_zc', type=int, default=50, help='number of dimensions of z content')
    parser.add_argument('--dim_zm', type=int, default=10, help='number of dimensions of z motion')
    parser.add_argument('--n_filters_gen', type=int, default=64, help='number of channelsof image generator')
    parser.add_argument('--n_filters_idis', type=int, default=64, help='number of channel of image discriminator')
    parser.add_argument('--n_filters_vdis', type=int, default=64, help='number of channel of video discriminator')
    parser.add_argument('--resume', '-r', default='', help='Resume the training from snapshot')
    # ---- additive
    parser.add_argument('--synthetic', type=int, default=0, help='use N synthetic clips instead of reading --dataset')
    parser.add_argument('--dtype', choices=['bf16', 'fp32'], default='bf16', help='bf16: tcgen05 path; fp32: strict path')
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--graph', action='store_true', help='replay the step as a CUDA graph')
    parser.add_argument('--max_iter', type=int, default=0, help='stop after this many iterations (0: --max_epoch rules)')
    parser.add_argument('--no_clip_cache', action='store_true',
                        help='feed float32 examples through SerialIterator as the reference does, instead of the uint8 clip cache')
    return parser


def build_models(model, dim_zc, dim_zm, num_labels, channel, n_filters_gen, video_length, use_noise, noise_sigma):
    """train.py:69-85 verbatim wiring (all three nets get n_filters_gen)."""
    if model == "normal":
        image_gen = ImageGenerator(dim_zc, dim_zm, num_labels, channel, n_filters_gen, video_length)
        image_dis = ImageDiscriminator(channel, 1, n_filters_gen, use_noise, noise_sigma)
        video_dis = VideoDiscriminator(channel, 1, n_filters_gen, use_noise, noise_sigma)
    elif model == "cgan":
        if num_labels == 0:
            raise ValueError("Called cgan model, but dataset has no label.")
        image_gen = ImageGenerator(dim_zc, dim_zm, num_labels, channel, n_filters_gen, video_length)
        image_dis = ImageDiscriminator(channel + num_labels, 1, n_filters_gen, use_noise, noise_sigma)
        video_dis = VideoDiscriminator(channel + num_labels, 1, n_filters_gen, use_noise, noise_sigma)
    elif model == "infogan":
        if num_labels == 0:
            raise ValueError("Called cgan model, but dataset has no label.")
        image_gen = ImageGenerator(dim_zc, dim_zm, num_labels, channel, n_filters_gen, video_length)
        image_dis = ImageDiscriminator(channel, 1 + num_labels, n_filters_gen, use_noise, noise_sigma)
        video_dis = VideoDiscriminator(channel, 1 + num_labels, n_filters_gen, use_noise, noise_sigma)
    else:
        raise ValueError(model)
    return image_gen, image_dis, video_dis


def make_optimizer(model, alpha=1e-3, beta1=0.9, beta2=0.999):
    optimizer = chainer.optimizers.Adam(alpha=alpha, beta1=beta1)   # beta2 is dropped, as in train.py:94
    optimizer.setup(model)
    optimizer.add_hook(chainer.optimizer.WeightDecay(1e-5), 'hook_dec')
    return optimizer


def main(argv=None):
    args = build_parser().parse_args(argv)
    if np.sqrt(args.num_gen_samples) % 1.0 != 0:
        raise ValueError('--num_gen_samples must be n^2 (n: natural number).')
    size, channel, video_length = 64, 3, 16      # train.py:48-50
    use_noise, noise_sigma = True, 0.2           # train.py:56-57
    num_labels = 6 if args.dataset_type == "mug" else 0
    rank, world = parallel.init_from_env()
    chainer.config.compute_dtype = args.dtype
    np.random.seed(args.seed)
    # device: the reference's default --gpu -1 means "CPU"; this build has no CPU path (the kernels are the product), so a
    # negative id is an error rather than a silent move to device 0.  Under torchrun every rank owns the GPU LOCAL_RANK
    # names, whatever --gpu says.
    if world > 1:
        device = int(os.environ.get("LOCAL_RANK", rank))
    elif args.gpu < 0:
        raise ValueError("--gpu {}: this build runs on a CUDA device only (no CPU fallback); pass --gpu 0".format(args.gpu))
    else:
        device = args.gpu
    chainer.cuda.get_device_from_id(device).use()
    if args.synthetic:
        train_dataset = SyntheticClipDataset(args.synthetic, channel, video_length, size, num_labels,
                                             seed=parallel.shard_seed(1234, rank))
    elif args.dataset_type == "mug":       # train.py:60-62
        train_dataset = MugDataset(args.dataset, video_length)
    elif args.dataset_type == "mnist":     # train.py:63-65
        train_dataset = MovingMnistDataset(args.dataset, video_length)
    else:
        raise NotImplementedError
    if hasattr(train_dataset, "clip_cache") and not args.no_clip_cache:
        # videos decoded once, uint8 batches assembled in pinned memory and normalised on the device
        train_iter = train_dataset.clip_cache(args.batchsize)
    else:
        train_iter = chainer.iterators.SerialIterator(train_dataset, args.batchsize)
    image_gen, image_dis, video_dis = build_models(args.model, args.dim_zc, args.dim_zm, num_labels, channel,
                                                   args.n_filters_gen, video_length, use_noise, noise_sigma)
    opts = {'image_gen': make_optimizer(image_gen, 2e-4, 5e-5, 0.999),
            'image_dis': make_optimizer(image_dis, 2e-4, 5e-5, 0.999),
            'video_dis': make_optimizer(video_dis, 2e-4, 5e-5, 0.999)}
    parallel.attach(list(opts.values()))
    mrandom.set_source(mrandom.DeviceRandom(seed=parallel.shard_seed(args.seed, rank), video_length=video_length))
    save_path = Path('result') / args.save_name
    writer = None
    if rank == 0:
        save_path.mkdir(parents=True, exist_ok=True)
        if args.log_tensorboard_interval > 0:
            writer = util.make_writer(Path('runs') / args.save_name)          # train.py:103
    updater = Updater(model=args.model, models=(image_gen, image_dis, video_dis), video_length=video_length,
                      img_size=size, channel=channel, dim_zl=num_labels, iterator=train_iter, tensorboard_writer=writer,
                      optimizer=opts, device=device, use_graph=args.graph)
    log_samples = util.log_tensorboard(image_gen, args.num_gen_samples, video_length, writer) if writer is not None else None
    log_entries = []
    trainer_state = chainer.training.TrainerState(updater)
    if args.resume:
        # train.py:162-163 `serializers.load_npz(args.resume, trainer)`: a full-trainer snapshot_epoch_N.npz restores the
        # models, Adam moments and step counts, the iterator position / order and the iteration count.  (Additive: a
        # pattern with {name} loads three per-model files instead.)
        if '{name}' in args.resume:
            for name, m in (('image_gen', image_gen), ('image_dis', image_dis), ('video_dis', video_dis)):
                chainer.serializers.load_npz(args.resume.format(name=name), m)
        else:
            chainer.serializers.load_npz(args.resume, trainer_state)
    if rank == 0:
        print('[ Training configuration ]')
        print('# minibatch size: {}  max epoch: {}  data size: {}  model: {}  dtype: {}  world: {}'.format(
            args.batchsize, args.max_epoch, len(train_dataset), args.model, args.dtype, world))
    t0 = time.time()
    while updater.epoch < args.max_epoch and not (args.max_iter and updater.iteration >= args.max_iter):
        updater.update()
        if updater.is_new_epoch and rank == 0:
            ep = updater.epoch
            if ep % args.display_interval == 0:
                ls = {k: float(v) for k, v in updater.losses.items()}
                print('epoch {:4d} iteration {:6d} image_gen/loss {:.4f} image_dis/loss {:.4f} video_dis/loss {:.4f} '
                      '({:.1f} it/s)'.format(ep, updater.iteration, ls.get('ImageGenerator', float('nan')),
                                             ls.get('ImageDiscriminator', float('nan')),
                                             ls.get('VideoDiscriminator', float('nan')),
                                             updater.iteration / (time.time() - t0)))
                # extensions.LogReport (train.py:146-147): result/<save_name>/log, one JSON entry per trigger
                log_entries.append({"epoch": ep, "iteration": updater.iteration, "elapsed_time": time.time() - t0,
                                    "image_gen/loss": ls.get('ImageGenerator'), "image_dis/loss": ls.get('ImageDiscriminator'),
                                    "video_dis/loss": ls.get('VideoDiscriminator')})
                with open(save_path / 'log', 'w') as f:
                    json.dump(log_entries, f, indent=4)
            if log_samples is not None and ep % args.log_tensorboard_interval == 0:
                log_samples(updater)                                              # train.py:154-160
            if ep % args.snapshot_interval == 0:
                chainer.serializers.save_npz(save_path / 'snapshot_epoch_{}.npz'.format(ep), trainer_state)   # train.py:137
                chainer.serializers.save_npz(save_path / 'image_gen_epoch_{}.npz'.format(ep), image_gen)
                chainer.serializers.save_npz(save_path / 'image_dis_epoch_{}.npz'.format(ep), image_dis)
                chainer.serializers.save_npz(save_path / 'video_dis_epoch_{}.npz'.format(ep), video_dis)
    if rank == 0:
        chainer.serializers.save_npz(save_path / 'image_gen_epoch_fianl.npz', image_gen)   # sic, train.py:190-192
        chainer.serializers.save_npz(save_path / 'image_dis_epoch_fianl.npz', image_dis)
        chainer.serializers.save_npz(save_path / 'video_dis_epoch_fianl.npz', video_dis)
        if writer is not None:
            writer.close()
    return updater


if __name__ == '__main__':
    main()
