"""BASELINE.json configs 4 and 5 on one GPU: 720p (736x1280) frame, and the correlation-only microbench at
feature maps 128^2 / 256^2 / 512^2 (C = 256) in both operand-split modes.  Prints one line per measurement."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch
import dvc
from dvc.synth import make_lab, make_state_dict

ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_TF32X3)

# ---- config 5: correlation only ----
ctx.profile_corr(True)
for side in (128, 256, 512):
    N = side * side
    g = torch.Generator(device="cuda").manual_seed(7)
    th = torch.nn.functional.normalize(torch.randn(1, 256, N, device="cuda", generator=g), dim=1)
    ph = torch.nn.functional.normalize(torch.randn(1, 256, N, device="cuda", generator=g), dim=1)
    V = torch.randn(1, N, 3, device="cuda", generator=g)
    for name, mode in (("fp16x3", dvc.MATH_FP16X3), ("tf32x3", dvc.MATH_TF32X3)):
        ctx.set_math(conv=dvc.MATH_TF32X3, corr=mode)
        for T in (1e-10, 0.01):
            if side == 512 and T > 1e-9 and name == "tf32x3":
                continue
            y, sim = ctx.corr_softmax_warp(th, ph, V, T); ctx.corr_mean_ms(True)
            reps = 3 if side < 512 else 1
            for _ in range(reps): ctx.corr_softmax_warp(th, ph, V, T)
            ms = ctx.corr_mean_ms(True)
            print(f"config5 corr-only features {side}x{side} N={N} {name} T={T:g}: {ms:.3f} ms  {2.0*N*N*259/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
    del th, ph, V
ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)

# ---- config 1: one 256x256 frame (latency) ----
H, W = 256, 256
ctx.set_exemplar(make_lab(4321, 1, H, W))
L1 = make_lab(1234, 1, H, W)[:, 0:1].cuda(); last1 = torch.zeros(1, 3, H, W, device="cuda")
for _ in range(3): ctx.colorize_frames(L1, last1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): ctx.colorize_frames(L1, last1)
e1.record(); torch.cuda.synchronize()
print(f"config1 256x256 frame (N=4096): {e0.elapsed_time(e1)/10:.3f} ms/frame on one stream", flush=True)

# ---- config 4: one 736x1280 frame (N = 58880) ----
H, W = 736, 1280
ctx.set_exemplar(make_lab(60, 1, H, W))
L = make_lab(61, 3, H, W)[:, 0:1].cuda(); last = torch.zeros(1, 3, H, W, device="cuda")
for _ in range(2): ab = ctx.colorize_frames(L[0:1], last)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
ctx.corr_mean_ms(True); e0.record()
for t in range(3): ab = ctx.colorize_frames(L[t:t + 1], last)
e1.record(); torch.cuda.synchronize()
print(f"config4 736x1280 frame (N=58880), one stream: {e0.elapsed_time(e1)/3:.2f} ms/frame, corr {ctx.corr_mean_ms(True):.3f} ms, finite={bool(torch.isfinite(ab).all())}", flush=True)
out = ctx.colorize_clip(L.contiguous()); torch.cuda.synchronize()
e0.record(); out = ctx.colorize_clip(L.contiguous()); e1.record(); torch.cuda.synchronize()
print(f"config4 736x1280 clip of 3 frames, two-stream pipeline: {e0.elapsed_time(e1)/3:.2f} ms/frame", flush=True)
