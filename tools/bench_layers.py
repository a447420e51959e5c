#!/usr/bin/env python
"""Times every tcgen05 convolution launch of the BASELINE config-2 step alone (bench.time_conv_layers) and prints a
table — the quick loop for kernel work (env MCG_TC_MT / MCG_TC_BN / MCG_TC_WMT / MCG_TC_WBN force a tile shape)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mocogan_chainer_b200 import kernels as K  # noqa: E402

peaks = bench.load_peaks()
rows = bench.time_conv_layers(K, torch, peaks)
tot = 0.0
for r in rows:
    tot += r["ms"] * r["calls_per_step"]
    print("%-8s %-14s %.4f ms x%d  %6.0f TF/s  %.2f" % (r["layer"], r["kernel"], r["ms"], r["calls_per_step"], r["tflops"],
                                                      r["frac_of_burst_peak"]))
print("sum over step: %.3f ms; tc error flag %d" % (tot, K.tc_error_flag()))
ntot = 0.0
for r in bench.time_narrow_layers(K, torch, peaks):
    ntot += r["ms"] * r["calls_per_step"]
    print("%-8s %-40s %.4f ms x%d  %6.0f GB/s  %.2f" % (r["layer"], r["kernel"], r["ms"], r["calls_per_step"], r["gb_per_s"],
                                                      r["frac_of_hbm_peak"]))
print("3-channel layers, sum over step: %.3f ms" % ntot)
