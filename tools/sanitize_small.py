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
# round-2 kernels: screened / exact / softmax correlation (pairs and single CTAs, ties, overflowing lists), the static
# exemplar side, other feature depths, the contextual loss, the second phase-A stream, WLS filter, CenterPad resize
g = torch.Generator().manual_seed(3)
th = torch.nn.functional.normalize(torch.randn(1, 256, 300, generator=g), dim=1).cuda()
ph = torch.nn.functional.normalize(torch.randn(1, 256, 517, generator=g), dim=1)
ph[:, :, 100:160] = ph[:, :, 7:8]
ph = ph.cuda()
V = (torch.randn(1, 517, 3, generator=g) * 30).cuda()
for cluster in (2, 1):
    ctx.debug_flag("corr_cluster", cluster)
    for screen in (1, 0):
        ctx.debug_flag("corr_screen", screen)
        for T in (1e-10, 0.01):
            y, sim = ctx.corr_softmax_warp(th, ph, V, T)
ctx.debug_flag("corr_cluster", 2), ctx.debug_flag("corr_screen", 1)
ctx.debug_flag("corr_phi_static", 1)
ctx.corr_softmax_warp(th, ph, V, 1e-10), ctx.corr_softmax_warp(th, ph, V, 1e-10)
ctx.debug_flag("corr_phi_static", 0)
t2 = torch.nn.functional.normalize(torch.randn(1, 128, 200, generator=g), dim=1).cuda()
p2 = torch.nn.functional.normalize(torch.randn(1, 128, 260, generator=g), dim=1).cuda()
ctx.corr_softmax_warp(t2, p2, V[:, :260].contiguous(), 0.1), ctx.corr_softmax_warp(t2, p2, V[:, :260].contiguous(), 1e-10)
X = torch.relu(torch.randn(2, 128, 12, 16, generator=g)).cuda()
loss = ctx.contextual_loss_forward(X, torch.relu(torch.randn(2, 128, 12, 16, generator=g)).cuda() + 0.5 * X)
ctx.debug_flag("clip_astreams", 2)
ctx.colorize_clip(make_lab(70, 5, 40, 64)[:, 0:1].contiguous().pin_memory())
ctx.debug_flag("clip_astreams", 1)
l = make_lab(80, 1, 37, 50)[0, 0].cuda()
ab2 = ctx.fgs_filter(ctx.l_to_guide8(l), torch.randn(2, 37, 50, device="cuda") * 30)
img = (torch.rand(123, 211, 3, generator=g) * 255).to(torch.uint8).cuda()
small = ctx.centerpad_rgb8(img, (64, 96))
up = ctx.centerpad_rgb8(img[:40, :50].contiguous(), (64, 96))
torch.cuda.synchronize()
print("round-2 kernels ok", float(loss.mean()), float(ab2.abs().mean()), int(small.sum()), int(up.sum()))
print("done")
