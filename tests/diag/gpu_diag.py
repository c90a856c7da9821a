"""Scratch diagnostics for a GPU box: stage-by-stage error report + rough timings (not a bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import numpy as np, torch
import dvc
from oracle import dvc_oracle as O
from oracle.weights import make_lab, make_state_dict

sds = {k: make_state_dict(k, seed=0) for k in ("vgg", "warp", "color")}
ctx = dvc.get_context(0)
for net, key in ((dvc.NET_VGG, "vgg"), (dvc.NET_WARP, "warp"), (dvc.NET_COLOR, "color")):
    ctx.set_weights(net, sds[key])
print(torch.cuda.get_device_name(0))
G = lambda n: dict(np.load(os.path.join(ROOT, "tests", "golden", n + ".npz")))
KEYS = ["r12", "r22", "r32", "r42", "r52"]

def rel(a, b):
    return float(np.abs(a - b).max()), float(np.abs(b).max())

for name in ("small_32x48", "padbranch_40x64"):
    g = G(name)
    IA, IB, last = (torch.from_numpy(g[k]) for k in ("IA_lab", "IB_lab", "IA_last_lab"))
    outs = ctx.vgg19_forward(O.gray2rgb_batch(IA[:, 0:1]).cuda(), KEYS, True)
    for k, o in zip(KEYS, outs):
        print(name, "vgg", k, "err %.3e of max %.3e" % rel(o.cpu().numpy(), g[f"A_{k}"]))
    with torch.no_grad():
        fA = O.vgg19_forward(sds["vgg"], O.gray2rgb_batch(IA[:, 0:1])); fB = O.exemplar_features(sds["vgg"], IB)
        An = [O.feature_normalize(t).cuda() for t in fA[1:]]; Bn = [O.feature_normalize(t).cuda() for t in fB[1:]]
    y, sim = ctx.warpnet_forward(IB.cuda(), An, Bn, 1e-10)
    print(name, "warpnet sim err %.3e" % np.abs(sim.cpu().numpy()[:, :, ::4, ::4] - g["sim64"]).max(),
          "warp mismatches", int((y.cpu().numpy()[:, :, ::4, ::4] != g["warped32"]).sum()))
    y2, sim2, am = ctx.corr_softmax_warp(torch.from_numpy(g["theta_hat32"]).cuda(), torch.from_numpy(g["phi_hat32"]).cuda(),
                                         torch.from_numpy(g["V32"]).cuda(), 1e-10, want_argmax=True)
    print(name, "corr(golden operands) argmax mismatches", int((am.cpu().numpy() != g["argmax64"]).sum()),
          "sim err %.3e" % np.abs(sim2.cpu().numpy().ravel() - g["sim64"].ravel()).max())
    up = lambda a: torch.nn.functional.interpolate(torch.from_numpy(a), scale_factor=4, mode="nearest")
    x = torch.cat((IA[:, 0:1], up(g["warped32"])[:, 1:3], up(g["sim32"]), last), 1)
    out = ctx.colorvidnet_forward(x.cuda()).cpu().numpy()
    print(name, "colorvidnet(golden input) |out-ab64| %.3e  |out-ab32| %.3e  floor |ab32-ab64| %.3e" % (
        np.abs(out - g["ab64"]).max(), np.abs(out - g["ab32"]).max(), np.abs(g["ab32"] - g["ab64"]).max()))
    ctx.set_exemplar(IB)
    ab, warp, sim = ctx.colorize_frames(IA[:, 0:1].cuda(), last.cuda(), 1e-10, want_warp=True)
    print(name, "fused: sim err %.3e warp mismatches %d |ab-ab64| %.3e" % (
        np.abs(sim.cpu().numpy()[:, :, ::4, ::4] - g["sim64"]).max(),
        int((warp.cpu().numpy()[:, :, ::4, ::4] != g["warped32"]).sum()), np.abs(ab.cpu().numpy() - g["ab64"]).max()))

# rough timings at the bench size
H, W = 480, 864
IB = make_lab(60, 1, H, W); ctx.set_exemplar(IB)
L = make_lab(61, 1, H, W)[:, 0:1].cuda(); last = torch.zeros(1, 3, H, W, device="cuda")
for _ in range(2): ctx.colorize_frames(L, last)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
ctx.profile_corr(True); ctx.launch_count(True)
e0.record()
for _ in range(3): ctx.colorize_frames(L, last)
e1.record(); torch.cuda.synchronize()
print("480x864 fused frame: %.2f ms/frame, launches/frame %d, corr %.3f ms" % (e0.elapsed_time(e1) / 3, ctx.launch_count() // 3, ctx.corr_mean_ms()))
N = 25920
th = torch.nn.functional.normalize(torch.randn(1, 256, N, device="cuda"), dim=1); ph = torch.nn.functional.normalize(torch.randn(1, 256, N, device="cuda"), dim=1)
V = torch.randn(1, N, 3, device="cuda")
for T in (1e-10, 0.01):
    ctx.corr_softmax_warp(th, ph, V, T); ctx.corr_mean_ms(True)
    for _ in range(3): ctx.corr_softmax_warp(th, ph, V, T)
    ms = ctx.corr_mean_ms(True)
    print("corr N=25920 T=%g: %.3f ms -> %.1f TFLOP/s" % (T, ms, 2 * N * N * 259 / ms / 1e9))
