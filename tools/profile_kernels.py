"""Workload for `ncu --set full`: exemplar prologue + 2 frames at 480x864 on one stream (no overlap)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch
import dvc
from dvc.synth import make_lab, make_state_dict

ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
ctx.set_math(conv=dvc.MATH_TF32X3, corr={"tf32x3": dvc.MATH_TF32X3, "bf16x3": dvc.MATH_BF16X3, "fp16x3": dvc.MATH_FP16X3}[os.environ.get("DVC_CORR", "fp16x3")])
ctx.debug_flag("tc_kc", int(os.environ.get("DVC_KC", "1")))
ctx.debug_flag("tc_splits", int(os.environ.get("DVC_SPLITS", "1")))
ctx.debug_flag("corr_screen", int(os.environ.get("DVC_SCREEN", "1")))
T = float(os.environ.get("DVC_T", "1e-10"))
H, W = 480, 864
ctx.set_exemplar(make_lab(60, 1, H, W))
L = make_lab(61, 2, H, W)[:, 0:1].cuda()
last = torch.zeros(1, 3, H, W, device="cuda")
for t in range(2):
    ab = ctx.colorize_frames(L[t:t + 1], last, T)
    last = torch.cat((L[t:t + 1], ab), 1)
torch.cuda.synchronize()
print("done", float(ab.abs().mean()))
