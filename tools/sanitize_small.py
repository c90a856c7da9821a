"""Small end-to-end workload for compute-sanitizer (memcheck): module-level forwards + fused frame + clip at 32x48 / 40x64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch
import dvc
from dvc.synth import make_lab, make_state_dict

ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, make_state_dict(key, seed=0))
for (H, W) in ((32, 48), (40, 64)):
    ctx.set_exemplar(make_lab(60, 1, H, W))
    L = make_lab(61, 3, H, W)[:, 0:1].cuda()
    last = torch.zeros(1, 3, H, W, device="cuda")
    ab = ctx.colorize_frames(L[0:1], last)
    out = ctx.colorize_clip(L.contiguous().cpu().pin_memory())
    feats = ctx.vgg19_forward(torch.rand(1, 3, H, W, device="cuda"), ["r12", "r22", "r32", "r42", "r52"], True)
    x = torch.randn(1, 7, H, W, device="cuda")
    o = ctx.colorvidnet_forward(x)
    rgb = ctx.lab_to_rgb8(L[0:1], ab)
    lab = ctx.rgb8_to_lab(rgb)
    torch.cuda.synchronize()
    print(H, W, "ok", float(ab.abs().mean()), float(out.abs().mean()), float(o.abs().mean()))
print("done")
