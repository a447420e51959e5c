#!/usr/bin/env python
"""One generate_samples-style batch (BASELINE config 5: 256 clips x 32 frames) a few times — the command ncu's launch
list is pointed at for the generator path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print(bench.gen_frames_per_s(torch, bench.load_peaks(), iters=2))
