"""Per-layer timing of the tensor-core convolution engine at the bench's geometries (480x864 frame) under different tile /
operand-staging choices: CUDA-event time of the conv_tc launch alone (dvc_profile_conv), through dvc_debug_conv2d.

    python tools/conv_layer_bench.py [--out gpurun_out/conv_layers_r2.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch

import dvc
from dvc.synth import make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
V, Wn, C = dvc.NET_VGG, dvc.NET_WARP, dvc.NET_COLOR
# (label, net, name, cin, cout, H, W, kwargs)
LAYERS = [
    ("full 64->64 (conv1_2)", V, "conv1_2", 64, 64, 480, 864, dict(act=1)),
    ("full 128->128 (conv10_2 + tail)", C, "conv10_2", 128, 128, 480, 864, dict(act=2, slope=0.2, fuse_tail=True)),
    ("half 128->128 (conv2_2)", V, "conv2_2", 128, 128, 240, 432, dict(act=1)),
    ("quarter 256->256 (conv3_2)", V, "conv3_2", 256, 256, 120, 216, dict(act=1)),
    ("quarter 256->256 reflect+stats (res block)", Wn, "layer.0.conv1", 256, 256, 120, 216, dict(reflect=True, want_stats=True)),
    ("eighth 512->512 (conv4_2)", V, "conv4_2", 512, 512, 60, 108, dict(act=1)),
    ("eighth 512->512 dil 2 (conv5_2)", C, "conv5_2", 512, 512, 60, 108, dict(act=1, dil=2)),
    ("eighth->quarter upconv 512->256 (conv8_1)", C, "conv8_1.1", 512, 256, 60, 108, dict(act=1, upconv=True)),
    ("half->full upconv 128->128 (conv10_1)", C, "conv10_1.1", 128, 128, 240, 432, dict(act=1, upconv=True)),
]
VARIANTS = [("default", dict()), ("rowshare", dict(tc_rowshare=1)), ("bn128", dict(tc_force_bn=128)),
            ("bn128+rowshare", dict(tc_force_bn=128, tc_rowshare=1)), ("bn64+rowshare", dict(tc_force_bn=64, tc_rowshare=1)),
            ("bn256+rowshare", dict(tc_force_bn=256, tc_rowshare=1)), ("kc2", dict(tc_kc=2)), ("kc2+rowshare", dict(tc_kc=2, tc_rowshare=1))]
if os.environ.get("DVC_LAYER_EXPERIMENTS"):  # timing-only experiments (results are wrong): which resource binds the tile?
    VARIANTS = [("default", dict()), ("no-lo-plane", dict(tc_dbg=2)), ("rowshare", dict(tc_rowshare=1)),
                ("rowshare-noshift", dict(tc_rowshare=1, tc_dbg=1)), ("single-cta", dict(tc_cluster=1)), ("kc2", dict(tc_kc=2)),
                ("kc2+no-lo", dict(tc_kc=2, tc_dbg=2)), ("kc4", dict(tc_kc=4))]
lines = []
for label, net, name, cin, cout, H, W, kw in LAYERS:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1, cin, H, W, device="cuda", generator=g).abs()
    taps = 4 if kw.get("upconv") else 9
    flop = 2.0 * H * W * (4 if kw.get("upconv") else 1) * taps * cin * cout
    row = {"layer": label, "gflop": flop / 1e9}
    for vname, flags in VARIANTS:
        for k in ("tc_rowshare", "tc_force_bn", "tc_dbg"):
            ctx.debug_flag(k, flags.get(k, 0))
        ctx.debug_flag("tc_kc", flags.get("tc_kc", 1))
        ctx.debug_flag("tc_cluster", flags.get("tc_cluster", 2))
        ctx.debug_conv2d(net, name, x, cout, **kw)
        ctx.profile_conv(True)
        ctx.conv_profile(0, reset=True)
        for _ in range(args.reps):
            ctx.debug_conv2d(net, name, x, cout, **kw)
        torch.cuda.synchronize()
        n, ms, fl = ctx.conv_profile(0, reset=True)
        ctx.profile_conv(False)
        us = 1e3 * ms / args.reps
        row[vname] = {"us": us, "tflops": flop / us / 1e6}
    for k in ("tc_rowshare", "tc_force_bn", "tc_dbg"):
        ctx.debug_flag(k, 0)
    ctx.debug_flag("tc_kc", 1)
    ctx.debug_flag("tc_cluster", 2)
    lines.append(row)
    print(label, f"{row['gflop']:.1f} GF:", "  ".join(f"{v} {row[v]['us']:.1f}us ({row[v]['tflops']:.0f})" for v, _ in VARIANTS), flush=True)
if args.out:
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
