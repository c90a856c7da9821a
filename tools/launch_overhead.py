"""Host enqueue time vs device time of one fused frame (no profiling events): is the frame launch-bound?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch
import dvc
from dvc.synth import make_lab, make_state_dict

ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
H, W = 480, 864
ctx.set_exemplar(make_lab(60, 1, H, W))
L = make_lab(61, 8, H, W)[:, 0:1].cuda()
last = torch.zeros(1, 3, H, W, device="cuda")
for tail in (1, 0):
    ctx.debug_flag("tc_tail", tail)
    for t in range(2):
        ctx.colorize_frames(L[t:t + 1], last)
    torch.cuda.synchronize()
    n = 6
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ctx.launch_count(True)
    t0 = time.perf_counter(); e0.record()
    for t in range(n):
        ctx.colorize_frames(L[t:t + 1], last)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"tc_tail={tail}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/frame, device {e0.elapsed_time(e1) / n:.3f} ms/frame, "
          f"wall {1e3 * (t2 - t0) / n:.3f} ms/frame, launches/frame {ctx.launch_count() // n}")
