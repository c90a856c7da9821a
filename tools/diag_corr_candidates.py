"""Diagnostic (GPU): how many reference positions fall within a threshold of each query row's best score on the
bench's own frames -- sizes the candidate lists of the screened (one fp16 pass + exact re-scoring) T -> 0 correlation.
torch is used only to inspect the library's theta_hat / phi_hat operands; nothing here is on the product path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch

import bench
import dvc
from dvc.synth import make_state_dict

torch.backends.cuda.matmul.allow_tf32 = False
ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
H, W = bench.H, bench.W
N = (H // 4) * (W // 4)
ctx.set_exemplar(bench.synth_exemplar())
frames = bench.synth_frames(3, 1000).cuda()
last = torch.zeros(1, 3, H, W, device="cuda")
for t in range(3):
    ab = ctx.colorize_frames(frames[t:t + 1], last)
    torch.cuda.synchronize()
    th = ctx.debug_buffer("fr.theta", act=False)[: N * 256].view(N, 256)
    ph = ctx.debug_buffer("ex.phi", act=False)[: N * 256].view(N, 256)
    # what a single fp16 pass sees: hi planes of x * 2^14
    th_hi = (th * 16384).half().float() / 16384
    ph_hi = (ph * 16384).half().float() / 16384
    thrs = [1e-3, 5e-4, 2.5e-4, 1e-4, 3e-5]
    counts = {k: [] for k in thrs}
    err = 0.0
    for r0 in range(0, N, 4320):
        f = th[r0:r0 + 4320].double() @ ph.double().t()
        fh = (th_hi[r0:r0 + 4320] @ ph_hi.t()).double()
        err = max(err, float((f - fh).abs().max()))
        m = fh.max(1, keepdim=True).values
        for k in thrs:
            counts[k].append((fh >= m - k).sum(1))
    print(f"frame {t}: max |f - f_hi.hi| = {err:.2e}, row max in [{float(m.min()):.3f}, ...]")
    for k in thrs:
        c = torch.cat(counts[k]).float()
        qs = torch.quantile(c, torch.tensor([0.5, 0.9, 0.99, 0.999], device=c.device))
        print(f"  thr {k:g}: mean {float(c.mean()):.1f}  p50 {qs[0]:.0f} p90 {qs[1]:.0f} p99 {qs[2]:.0f} p99.9 {qs[3]:.0f} max {float(c.max()):.0f}"
              f"  rows>8: {float((c > 8).float().mean()) * 100:.2f}%  rows>32: {float((c > 32).float().mean()) * 100:.2f}%")
    last = torch.cat((frames[t:t + 1], ab), 1)
