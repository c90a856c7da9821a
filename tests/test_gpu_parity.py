"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI against
(1) the committed golden vectors generated from the real reference, (2) the CPU oracle run here on the
same seeded inputs, (3) size-independent properties at the full 480x864 bench size.

Tolerances (floating point, stated per SURVEY.md §8c):
  VGG maps        max|d| <= 1e-4 * max|ref|            (fp32 summation-order noise)
  similarity      max|d| <= 2e-5
  warp (T->0)     identical argmax on every row whose fp64 top-2 gap > 1e-5 (tie-aware)
  final ab        max|ours - ref_fp64| <= max(1e-3, k * max|ref_fp32 - ref_fp64|): ColorVidNet with the seeded
                  random weights amplifies a 1e-6 input perturbation to ~1e-3 (measured), so the
                  reference's own fp32 forward sits 1e-3..2e-2 away from fp64; ours must be in the same band.
                  k = 1.25 for the default engine and the exact-fp32 CUDA-core engine; k = 2 only for
                  the debug variants of the tensor-core engine (single CTAs, 64-byte stages, split-K, tail rounds,
                  3xTF32 planes), which are not what ships.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import dvc_oracle as O
from oracle.weights import make_lab

pytestmark = pytest.mark.gpu
KEYS = ["r12", "r22", "r32", "r42", "r52"]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SHIPPING = ("fp32", "tf32x3", "tf32x3-bn256")  # engines gated at 1.25x the reference's own fp32 noise


def ab_gate(ours, g, variant="debug"):
    floor = np.abs(g["ab32"].astype(np.float64) - g["ab64"]).max()
    err = np.abs(ours.astype(np.float64) - g["ab64"]).max()
    return err, max(1e-3, (1.25 if variant in SHIPPING else 2.0) * floor)


@pytest.fixture(params=["fp32", "tf32x3", "tf32x3-bn256", "tf32x3-bn256-cluster1", "tf32x3-bn64", "tf32x3-nof16", "tf32x3-cluster1",
                        "tf32x3-cluster1-nof16", "tf32x3-cluster1-k64", "tf32x3-k64", "tf32x3-split3", "tf32x3-tail16", "tf32x3-tail"])
def conv_math(request, ctx):
    """Convolutions on CUDA cores (exact fp32, two-level accumulation) and on tcgen05 (3xTF32 operand split),
    the latter as single CTAs (64-byte and 128-byte K stages) and as CTA pairs (tcgen05.mma.cta_group::2).
    Layers with provably bounded inputs run 3xFP16 on scaled planes unless "-nof16" turns that off.
    "-bn256" pins the 256-channel tile (BN = 256: its own TMEM ring depth, stage count and two-loads-in-flight drain) on
    every layer with >= 256 output channels -- the tile the 480x864 bench runs on, which the launcher's heuristic
    never picks at the golden sizes; "-bn64" pins the narrow tile on every layer."""
    import dvc

    if request.param.startswith("tf32x3"):
        ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)
        ctx.debug_flag("tc_cluster", 1 if "cluster1" in request.param else 2)
        ctx.debug_flag("tc_force_bn", 256 if "bn256" in request.param else (64 if "bn64" in request.param else 0))
        ctx.debug_flag("tc_kbytes", 64 if request.param.endswith("k64") else 128)
        ctx.debug_flag("tc_splits", 3 if request.param.endswith("split3") else 1)
        ctx.debug_flag("tc_f16", 0 if request.param.endswith("nof16") else 1)
        # tail rounds of 256-channel launches on 128-channel tiles: forced at small sizes by pretending 16 pair slots
        ctx.debug_flag("tc_tail", 16 if request.param.endswith("tail16") else (1 if request.param.endswith("-tail") else 0))
    else:
        ctx.set_math(conv=dvc.MATH_FP32, corr=dvc.MATH_FP32)
    yield request.param
    ctx.debug_flag("tc_cluster", 2)
    ctx.debug_flag("tc_kbytes", 128)
    ctx.debug_flag("tc_splits", 1)
    ctx.debug_flag("tc_f16", 1)
    ctx.debug_flag("tc_tail", 0)
    ctx.debug_flag("tc_force_bn", 0)
    ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)


# ------------------------------------------------------------------------------------------ VGG19
@pytest.mark.parametrize("name", ["small_32x48", "padbranch_40x64"])
def test_vgg19_module_vs_golden(ctx, conv_math, name):
    g = load_golden(name)
    IA = torch.from_numpy(g["IA_lab"])
    x = O.gray2rgb_batch(IA[:, 0:1]).cuda()
    outs = ctx.vgg19_forward(x, KEYS, preprocess=True)
    for k, o in zip(KEYS, outs):
        ref = g[f"A_{k}"]
        assert o.shape == ref.shape
        err = np.abs(o.cpu().numpy() - ref).max()
        assert err <= 1e-4 * np.abs(ref).max(), (k, err, np.abs(ref).max())


def test_vgg19_all_keys_and_no_preprocess(ctx, conv_math, sds):
    x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(5))
    keys = ["r11", "p1", "r21", "r34", "p3", "r44", "r54", "p5"]
    with torch.no_grad():
        ref = O.vgg19_forward(sds["vgg"], x * 20 - 10, keys, preprocess=False)
    outs = ctx.vgg19_forward((x * 20 - 10).cuda(), keys, preprocess=False)
    for k, o, r in zip(keys, outs, ref):
        assert o.shape == r.shape, k
        assert (o.cpu() - r).abs().max() <= 1e-4 * r.abs().max() + 1e-9, k


# ------------------------------------------------------------------------------------------ K7
@pytest.fixture(params=["fp32", "tf32x3", "bf16x3", "fp16x3", "tf32x3-single", "fp16x3-single", "fp16x3-noscreen",
                        "fp16x3-noscreen-single"])
def corr_math(request, ctx):
    """Run the correlation tests on the CUDA-core kernel and on the tcgen05 operand-split modes, as CTA pairs
    (cta_group::2, the default) and as single CTAs.  fp16x3 at T -> 0 takes the screened path by default (one fp16 pass
    + exact fp32 re-scoring of the candidates); "-noscreen" pins the exact 3-pass kernel."""
    import dvc

    name = request.param.replace("-single", "").replace("-noscreen", "")
    mode = {"fp32": dvc.MATH_FP32, "tf32x3": dvc.MATH_TF32X3, "bf16x3": dvc.MATH_BF16X3, "fp16x3": dvc.MATH_FP16X3}[name]
    ctx.set_math(conv=dvc.MATH_TF32X3, corr=mode)
    ctx.debug_flag("corr_cluster", 1 if request.param.endswith("-single") else 2)
    ctx.debug_flag("corr_screen", 0 if "noscreen" in request.param else 1)
    yield name
    ctx.debug_flag("corr_cluster", 2)
    ctx.debug_flag("corr_screen", 1)
    ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)


@pytest.mark.parametrize("NA,NB,T", [(96, 96, 1e-10), (300, 517, 1e-10), (300, 517, 0.01), (1000, 130, 0.005),
                                     (5184, 5184, 1e-10)])
def test_corr_kernel_vs_oracle(ctx, corr_math, NA, NB, T):
    gen = torch.Generator().manual_seed(7)
    th = torch.nn.functional.normalize(torch.randn(1, 256, NA, generator=gen), dim=1)
    ph = torch.nn.functional.normalize(torch.randn(1, 256, NB, generator=gen), dim=1)
    V = torch.randn(1, NB, 3, generator=gen) * 30
    y, sim, am = ctx.corr_softmax_warp(th.cuda(), ph.cuda(), V.cuda(), T, want_argmax=True)
    yo, so, io = O.corr_softmax_warp(th.double(), ph.double(), V.double(), T, return_argmax=True)
    gap = O.top2_gap(th.double(), ph.double())
    # score error: fp32 FMA / 3xTF32 ~1e-7; 3xBF16 drops the lo.lo term (2^-16 relative per product) ~2e-6
    tol = 8e-6 if corr_math == "bf16x3" else 2e-6
    assert (sim.cpu().double() - so).abs().max() < tol
    if T < 1e-9:
        clear = gap[0] > 4 * tol
        assert (am.cpu()[0][clear] == io[0][clear]).all()
        assert torch.equal(y.cpu()[0][clear], V[0][io[0][clear]])  # one-hot: exact rows of V
    else:
        # softmax weights see the score error as exp(df / T)
        assert (y.cpu().double() - yo).abs().max() < (2e-2 if corr_math == "bf16x3" else 2e-3)


def test_corr_duplicated_exemplar_columns_average(ctx, corr_math):
    """Bit-equal maxima (duplicated phi columns: letterbox bars, flat exemplar regions): the reference's
    softmax(f / 1e-10) (NonlocalNet.py:486-497) averages the V rows of all of them -- so must the T -> 0 path."""
    gen = torch.Generator().manual_seed(9)
    NA, NB = 300, 700
    ph = torch.nn.functional.normalize(torch.randn(1, 256, NB, generator=gen), dim=1)
    dup = [3, 150, 151, 400, 699]            # five copies of column 3 spread over several 32-column chunks / tiles
    ph[:, :, dup] = ph[:, :, 3:4]
    ph[:, :, [20, 21]] = ph[:, :, 20:21]     # and a pair
    th = torch.nn.functional.normalize(torch.randn(1, 256, NA, generator=gen), dim=1)
    th[:, :, :40] = torch.nn.functional.normalize(ph[:, :, 3:4] + 0.05 * th[:, :, :40], dim=1)   # rows that pick the 5 copies
    th[:, :, 40:60] = torch.nn.functional.normalize(ph[:, :, 20:21] + 0.05 * th[:, :, 40:60], dim=1)
    V = torch.randn(1, NB, 3, generator=gen) * 30
    y, sim, am = ctx.corr_softmax_warp(th.cuda(), ph.cuda(), V.cuda(), 1e-10, want_argmax=True)
    yo, so, io = O.corr_softmax_warp(th.double(), ph.double(), V.double(), 1e-10, return_argmax=True)
    y = y.cpu().double()
    assert (y[0, :40] - V[0, dup].double().mean(0)).abs().max() < 1e-4
    assert (y[0, 40:60] - V[0, [20, 21]].double().mean(0)).abs().max() < 1e-4
    assert (am.cpu()[0, :40] == 3).all() and (am.cpu()[0, 40:60] == 20).all()  # lowest index of the tie
    # the matched rows score ~1.0 (a query that IS an exemplar column): the operand splits' relative error shows in full
    tol = {"bf16x3": 8e-6, "tf32x3": 4e-6}.get(corr_math, 2e-6)
    gap = O.top2_gap(th.double(), ph.double())[0]
    ok = (gap == 0) | (gap > 4 * tol)     # exact ties (averaged by the fp64 oracle too) or a clear winner
    assert (y[0][ok] - yo[0][ok]).abs().max() < 1e-3
    assert (sim.cpu().double() - so).abs().max() < tol


def test_corr_many_near_ties_overflow_the_candidate_lists(ctx, corr_math):
    """60 exemplar columns within 1e-6 of each other (plus 40 exact copies): far more candidates than a screening list
    holds, so the affected rows take the brute-force re-scoring path; results must not change."""
    gen = torch.Generator().manual_seed(10)
    NA, NB = 200, 900
    ph = torch.nn.functional.normalize(torch.randn(1, 256, NB, generator=gen), dim=1)
    base = ph[:, :, 7:8].clone()
    near = list(range(300, 360))
    ph[:, :, near] = torch.nn.functional.normalize(base + 1e-6 * torch.randn(1, 256, 60, generator=gen), dim=1)
    same = list(range(500, 540))
    ph[:, :, same] = ph[:, :, 333:334]
    th = torch.nn.functional.normalize(torch.randn(1, 256, NA, generator=gen), dim=1)
    th[:, :, :50] = torch.nn.functional.normalize(base + 0.05 * th[:, :, :50], dim=1)
    V = torch.randn(1, NB, 3, generator=gen) * 30
    y, sim, am = ctx.corr_softmax_warp(th.cuda(), ph.cuda(), V.cuda(), 1e-10, want_argmax=True)
    f = th[0].double().t() @ ph[0].double()
    m64, i64 = f.max(1)
    tol = {"bf16x3": 8e-6, "tf32x3": 4e-6}.get(corr_math, 2e-6)  # scores ~1.0 here, see the duplicated-columns test
    assert (sim.cpu().double()[0] - m64).abs().max() < tol
    # every reported argmax attains the fp64 maximum up to the score tolerance (the near ties are legal alternatives)
    assert (f.gather(1, am.cpu()[0].long().view(-1, 1))[:, 0] - m64).abs().max() < 2 * tol
    gap = O.top2_gap(th.double(), ph.double())[0]
    clear = gap > 4 * tol
    assert (am.cpu()[0][clear].long() == i64[clear]).all()


def test_corr_kernel_shared_exemplar_batch(ctx, corr_math):
    gen = torch.Generator().manual_seed(8)
    th = torch.nn.functional.normalize(torch.randn(3, 256, 200, generator=gen), dim=1)
    ph = torch.nn.functional.normalize(torch.randn(1, 256, 333, generator=gen), dim=1)
    V = torch.randn(1, 333, 3, generator=gen)
    y, sim = ctx.corr_softmax_warp(th.cuda(), ph.cuda(), V.cuda(), 1e-10)
    for b in range(3):
        yb, sb = ctx.corr_softmax_warp(th[b:b + 1].cuda(), ph.cuda(), V.cuda(), 1e-10)
        assert torch.equal(y[b:b + 1], yb) and torch.equal(sim[b:b + 1], sb)


def test_corr_static_exemplar_side(ctx, corr_math):
    """corr_phi_static: the exemplar side (transpose, V rows, operand planes) of the stand-alone entry is prepared once per
    (pointer, size) while the flag is set -- same bits as preparing it every call; re-arming the flag picks up new contents."""
    gen = torch.Generator().manual_seed(12)
    th = torch.nn.functional.normalize(torch.randn(1, 256, 400, generator=gen), dim=1).cuda()
    th2 = torch.nn.functional.normalize(torch.randn(1, 256, 400, generator=gen), dim=1).cuda()
    ph = torch.nn.functional.normalize(torch.randn(1, 256, 600, generator=gen), dim=1).cuda()
    V = (torch.randn(1, 600, 3, generator=gen) * 30).cuda()
    for T in (1e-10, 0.01):
        ref1, ref2 = ctx.corr_softmax_warp(th, ph, V, T), ctx.corr_softmax_warp(th2, ph, V, T)
        ctx.debug_flag("corr_phi_static", 1)
        try:
            a1 = ctx.corr_softmax_warp(th, ph, V, T)    # prepares the exemplar side
            a2 = ctx.corr_softmax_warp(th2, ph, V, T)   # reuses it
            assert all(torch.equal(x, y) for x, y in zip(ref1 + ref2, a1 + a2))
            ph.copy_(torch.nn.functional.normalize(torch.randn(1, 256, 600, generator=gen), dim=1))  # new exemplar, same buffer
            ctx.debug_flag("corr_phi_static", 1)        # re-arm: the next call prepares it afresh
            b1 = ctx.corr_softmax_warp(th, ph, V, T)
        finally:
            ctx.debug_flag("corr_phi_static", 0)
        assert all(torch.equal(x, y) for x, y in zip(ctx.corr_softmax_warp(th, ph, V, T), b1))
        assert not torch.equal(b1[1], a1[1])


@pytest.mark.parametrize("B,C,h,w", [(2, 256, 24, 32), (1, 128, 40, 40), (1, 512, 20, 24), (1, 256, 64, 64)])
def test_contextual_loss_forward_vs_oracle(ctx, B, C, h, w):
    """SURVEY.md §8f row 4 (value only): ContextualLoss_forward on K7 -- row maxima, then the online softmax with a per-row
    temperature -- against the fp64 evaluation of the reference's formula; the reference's own fp32 run is the yardstick."""
    import dvc

    ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)
    g = torch.Generator().manual_seed(C + h)
    X = torch.relu(torch.randn(B, C, h, w, generator=g))
    Y = torch.relu(torch.randn(B, C, h, w, generator=g) + 0.5 * X)   # VGG-like non-negative maps, correlated content
    for centering in (True, False):
        with torch.no_grad():
            ref64 = O.contextual_loss_forward(X.double(), Y.double(), 0.1, centering)
            ref32 = O.contextual_loss_forward(X, Y, 0.1, centering)
        out = ctx.contextual_loss_forward(X.cuda(), Y.cuda(), 0.1, centering).cpu().double()
        floor = (ref32.double() - ref64).abs().max().item()
        err = (out - ref64).abs().max().item()
        assert err <= max(2e-4 * ref64.abs().max().item(), 4 * floor), (err, floor, ref64)


def test_corr_other_feature_depths(ctx):
    """The stand-alone K7 entry at C != 256 (e.g. the 3x3-patch features of NonlocalWeightedAverage, NonlocalNet.py:95-108)."""
    import dvc

    ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)
    for C in (64, 128, 576):
        gen = torch.Generator().manual_seed(C)
        th = torch.nn.functional.normalize(torch.randn(1, C, 300, generator=gen), dim=1)
        ph = torch.nn.functional.normalize(torch.randn(1, C, 450, generator=gen), dim=1)
        V = torch.randn(1, 450, 3, generator=gen) * 30
        for T in (1e-10, 0.1):
            y, sim = ctx.corr_softmax_warp(th.cuda(), ph.cuda(), V.cuda(), T)
            yo, so = O.corr_softmax_warp(th.double(), ph.double(), V.double(), T)
            assert (sim.cpu().double() - so).abs().max() < 2e-6
            gap = O.top2_gap(th.double(), ph.double())[0]
            ok = gap > 1e-5 if T < 1e-9 else torch.ones_like(gap, dtype=torch.bool)
            assert (y.cpu().double()[0][ok] - yo[0][ok]).abs().max() < 2e-3


def test_corr_golden_operands(ctx):
    g = load_golden("small_32x48")
    y, sim, am = ctx.corr_softmax_warp(cu(g["theta_hat32"]), cu(g["phi_hat32"]), cu(g["V32"]), 1e-10, want_argmax=True)
    clear = g["gap64"][0] > 1e-5
    assert (am.cpu().numpy()[0][clear] == g["argmax64"][0][clear]).all()
    h, w = g["sim32"].shape[2:]
    assert np.abs(sim.cpu().numpy().reshape(1, 1, h, w) - g["sim32"]).max() < 2e-6


# ------------------------------------------------------------------------------------------ WarpNet
@pytest.mark.parametrize("name", ["small_32x48", "padbranch_40x64", "softmax_32x64", "softmax5_48x48", "batch2_32x32"])
def test_warpnet_module_vs_golden(ctx, conv_math, sds, name):
    g = load_golden(name)
    IA, IB = torch.from_numpy(g["IA_lab"]), torch.from_numpy(g["IB_lab"])
    T = float(g["temperature"])
    with torch.no_grad():
        fA = O.vgg19_forward(sds["vgg"], O.gray2rgb_batch(IA[:, 0:1]))
        fB = O.exemplar_features(sds["vgg"], IB)
        An = [O.feature_normalize(t).cuda() for t in fA[1:]]
        Bn = [O.feature_normalize(t).cuda() for t in fB[1:]]
    y, sim = ctx.warpnet_forward(IB.cuda(), An, Bn, T)
    y2, sim2 = ctx.warpnet_forward(IB.cuda(), An, Bn, T, reuse_exemplar=True)
    assert torch.equal(y, y2) and torch.equal(sim, sim2)  # cached exemplar side is bit-identical
    assert y.shape == (IA.shape[0], 3, IA.shape[2], IA.shape[3])
    ys, ss = y.cpu().numpy()[:, :, ::4, ::4], sim.cpu().numpy()[:, :, ::4, ::4]
    # nearest x4 up-sampling: every 4x4 block is constant (NonlocalNet.py:499-500)
    assert torch.equal(y, torch.nn.functional.interpolate(y[:, :, ::4, ::4], scale_factor=4, mode="nearest"))
    assert np.abs(ss - g["sim64"]).max() < 2e-5
    B = IA.shape[0]
    if T < 1e-9:
        clear = g["gap64"] > 1e-5
        m = np.broadcast_to(clear[:, None, :], (B, 3, clear.shape[1]))
        assert np.abs(ys.reshape(B, 3, -1)[m] - g["warped32"].reshape(B, 3, -1)[m]).max() < 1e-4
    else:
        floor = np.abs(g["warped32"].astype(np.float64) - g["warped64"]).max()
        assert np.abs(ys - g["warped64"]).max() <= max(1e-3, 2 * floor)


def test_warpnet_rejects_illegal_shapes(ctx):
    import dvc

    z = lambda c, h, w: torch.zeros(1, c, h, w, device="cuda")
    feats = [z(128, 24, 20), z(256, 12, 10), z(512, 6, 5), z(512, 3, 2)]
    with pytest.raises(dvc.DvcError):
        ctx.warpnet_forward(z(3, 48, 40), feats, feats, 1e-10)  # W % 16 != 0 (reference: RuntimeError at NonlocalNet.py:464)
    feats = [z(128, 16, 16), z(256, 8, 8), z(512, 4, 4), z(512, 2, 2)]
    with pytest.raises(dvc.DvcError):
        ctx.warpnet_forward(z(3, 32, 32), feats, feats, 1e-10, wta_scale_weight=0.5)


# ------------------------------------------------------------------------------------------ ColorVidNet
@pytest.mark.parametrize("name", ["small_32x48", "padbranch_40x64", "batch2_32x32"])
def test_colorvidnet_module_vs_golden(ctx, conv_math, name):
    g = load_golden(name)
    IA, last = torch.from_numpy(g["IA_lab"]), torch.from_numpy(g["IA_last_lab"])
    up = lambda a: torch.nn.functional.interpolate(torch.from_numpy(a), scale_factor=4, mode="nearest")
    x = torch.cat((IA[:, 0:1], up(g["warped32"])[:, 1:3], up(g["sim32"]), last), 1)
    out = ctx.colorvidnet_forward(x.cuda()).cpu().numpy()
    err, tol = ab_gate(out, g, conv_math)
    assert err <= tol, (err, tol)


# ------------------------------------------------------------------------------------------ fused frame path
@pytest.mark.parametrize("scale", [1e-4, 3e-2, 1.0, 4e2, 1e5])
def test_device_derived_scales_follow_the_input_magnitude(ctx, sds, scale):
    """The fp16 hi/lo planes of the conv -> ReLU -> conv chains get their exponent on the device from the measured
    max |input| (DynOut): the same network input scaled by 1e-4 ... 1e5 must come out as accurate as at scale 1
    (compared with the fp64 oracle on the same scaled input; the first InstanceNorm removes the scale, so the fp32
    reference's own distance to fp64 is the yardstick).  Also covers VGG19, whose outputs scale linearly."""
    import dvc

    ctx.set_math(conv=dvc.MATH_TF32X3, corr=dvc.MATH_FP16X3)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(1, 7, 32, 48, generator=g) * 30 * scale
    with torch.no_grad():
        ref64 = O.colorvidnet_forward({k: v.double() for k, v in sds["color"].items()}, x.double()).numpy()
        ref32 = O.colorvidnet_forward(sds["color"], x).numpy()
    out = ctx.colorvidnet_forward(x.cuda()).cpu().numpy()
    floor = np.abs(ref32.astype(np.float64) - ref64).max()
    err = np.abs(out.astype(np.float64) - ref64).max()
    assert np.isfinite(out).all() and err <= max(1e-3, 2.0 * floor), (scale, err, floor)
    rgb = (torch.rand(1, 3, 32, 48, generator=g) * scale)
    with torch.no_grad():
        f64 = O.vgg19_forward({k: v.double() for k, v in sds["vgg"].items()}, rgb.double(), preprocess=False)
    outs = ctx.vgg19_forward(rgb.cuda(), ["r12", "r22", "r32", "r42", "r52"], False)
    for o, r in zip(outs, f64):
        r = r.numpy()
        assert np.abs(o.cpu().numpy().astype(np.float64) - r).max() <= 1e-4 * max(np.abs(r).max(), 1e-30)


@pytest.mark.parametrize("name", ["small_32x48", "padbranch_40x64", "softmax_32x64", "softmax5_48x48", "default_216x384"])
def test_fused_frame_vs_golden(ctx, conv_math, name):
    g = load_golden(name)
    IA, IB, last = (torch.from_numpy(g[k]) for k in ("IA_lab", "IB_lab", "IA_last_lab"))
    T = float(g["temperature"])
    ctx.set_exemplar(IB)
    ab, warp, sim = ctx.colorize_frames(IA[:, 0:1].cuda(), last.cuda(), T, want_warp=True)
    ss = sim.cpu().numpy()[:, :, ::4, ::4]
    ys = warp.cpu().numpy()[:, :, ::4, ::4]
    assert np.abs(ss - g["sim64"]).max() < 2e-5
    if T < 1e-9:
        clear = g["gap64"] > 1e-5
        m = np.broadcast_to(clear[:, None, :], (1, 3, clear.shape[1]))
        nbad = int((np.abs(ys.reshape(1, 3, -1)[m] - g["warped32"].reshape(1, 3, -1)[m]) > 1e-4).sum())
        assert nbad == 0, nbad
    err, tol = ab_gate(ab.cpu().numpy(), g, conv_math)
    assert err <= tol, (err, tol)


def test_fused_clip_recurrence(ctx, conv_math):
    """dvc_colorize_clip == chaining dvc_colorize_frames with last = cat(L, ab) (test.py:96), bit for bit, and
    every frame matches the reference under teacher forcing."""
    g = load_golden("clip3_32x48")
    frames, IB, ref = torch.from_numpy(g["frames_lab"]), torch.from_numpy(g["IB_lab"]), g["ab32"]
    ctx.set_exemplar(IB)
    out = ctx.colorize_clip(frames[:, 0:1].contiguous().pin_memory())
    ctx.debug_flag("clip_astreams", 2)  # frames t+1 and t+2 in flight on two phase-A streams: same bits
    try:
        out2 = ctx.colorize_clip(frames[:, 0:1].contiguous().pin_memory())
    finally:
        ctx.debug_flag("clip_astreams", 1)
    assert torch.equal(out, out2)
    last = torch.zeros(1, 3, 32, 48, device="cuda")
    for t in range(frames.shape[0]):
        L = frames[t:t + 1, 0:1].cuda()
        ab = ctx.colorize_frames(L, last)
        assert torch.equal(ab.cpu(), out[t:t + 1])
        last = torch.cat((L, ab), 1)
    # free-running against the reference's free-running clip: frame 0 has no history and must match; afterwards the
    # recurrence is chaotic (two bit-different fp32 runs of the REFERENCE are 0.4 apart by frame 3, its 1- vs
    # 8-thread runs 1.5e-3 apart on a single frame), so the later frames are reported, not gated
    free = [float(np.abs(out[t].numpy() - ref[t]).max()) for t in range(frames.shape[0])]
    print(f"free-running |ab - reference| per frame ({conv_math}): " + ", ".join(f"{v:.2e}" for v in free))
    assert free[0] < 5e-3
    # teacher forcing against the reference's own outputs
    last = torch.zeros(1, 3, 32, 48, device="cuda")
    for t in range(frames.shape[0]):
        L = frames[t:t + 1, 0:1].cuda()
        ab = ctx.colorize_frames(L, last)
        assert np.abs(ab.cpu().numpy() - ref[t:t + 1]).max() < 5e-3
        last = torch.cat((L, torch.from_numpy(ref[t:t + 1]).cuda()), 1)


def test_fused_batch_equals_single(ctx, conv_math):
    IB = make_lab(40, 1, 32, 64)
    ctx.set_exemplar(IB)
    L = make_lab(41, 3, 32, 64)[:, 0:1].cuda()
    last = make_lab(42, 3, 32, 64).cuda()
    ab, warp, sim = ctx.colorize_frames(L, last, want_warp=True)
    for b in range(3):
        ab1, warp1, sim1 = ctx.colorize_frames(L[b:b + 1], last[b:b + 1], want_warp=True)
        # the pixel tiles of a batched call straddle image boundaries differently, so InstanceNorm statistics are
        # summed in a different order: same argmax, scores / ab equal up to that fp32 noise
        assert torch.equal(warp[b:b + 1], warp1) and (sim[b:b + 1] - sim1).abs().max() < 2e-6
        assert (ab[b:b + 1] - ab1).abs().max() < 5e-3


def test_exemplar_export_import_roundtrip(ctx):
    IB = make_lab(50, 1, 32, 48)
    ctx.set_exemplar(IB)
    L, last = make_lab(51, 1, 32, 48)[:, 0:1].cuda(), make_lab(52, 1, 32, 48).cuda()
    ab = ctx.colorize_frames(L, last)
    pack = ctx.exemplar_export(32, 48).clone()
    ctx.set_exemplar(make_lab(53, 1, 32, 48))
    assert not torch.equal(ctx.colorize_frames(L, last), ab)
    ctx.exemplar_import(pack, 32, 48)
    assert torch.equal(ctx.colorize_frames(L, last), ab)


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_size_480x864_properties(ctx, conv_math):
    """BASELINE config 2 size (480x854 padded to 480x864, N=25920): properties that need no full oracle run."""
    H, W = 480, 864
    IB = make_lab(60, 1, H, W)
    ctx.set_exemplar(IB)
    L = make_lab(61, 1, H, W)[:, 0:1].cuda()
    last = torch.zeros(1, 3, H, W, device="cuda")
    ab, warp, sim = ctx.colorize_frames(L, last, want_warp=True)
    assert torch.isfinite(ab).all() and ab.abs().max() <= 128.0
    assert sim.max() <= 1.0 + 1e-5 and sim.min() >= -1.0 - 1e-5
    # one-hot warp: every warped colour is exactly one row of the 4x4-pooled exemplar
    V = torch.nn.functional.avg_pool2d(IB, 4).view(3, -1).t().contiguous()
    rows = warp[0, :, ::4, ::4].reshape(3, -1).t().cpu()
    d = torch.cdist(rows[::97].double(), V.double()).min(dim=1).values
    assert d.max() < 1e-4
    # self-match: colourising the exemplar's own luminance must find itself (similarity == 1, identity argmax)
    ab2, warp2, sim2 = ctx.colorize_frames(IB[:, 0:1].cuda(), last, want_warp=True)
    # (gray version of the exemplar differs from its colour version, so only sanity-check the range)
    assert sim2.max() <= 1.0 + 1e-5
    # determinism
    ab3 = ctx.colorize_frames(L, last)
    assert torch.equal(ab, ab3)


def test_oracle_on_the_fly_64x64(ctx, conv_math, sds):
    """Seeded inputs not in the golden set, checked against the CPU oracle run on this machine."""
    IA, IB, last = make_lab(70, 1, 64, 64), make_lab(71, 1, 64, 64), make_lab(72, 1, 64, 64)
    sds64 = {k: O._cast(v, torch.float64) for k, v in sds.items()}
    ex = {}
    with torch.no_grad():
        fB = O.exemplar_features(sds64["vgg"], IB.double())
        ab64, warped64, sim64, _ = O.frame_colorization(sds64, IA.double(), IB.double(), last.double(), fB, extras=ex)
        fB32 = O.exemplar_features(sds["vgg"], IB)
        ab32, _, _, _ = O.frame_colorization(sds, IA, IB, last, fB32)
    gap = O.top2_gap(ex["theta_hat"], ex["phi_hat"])
    ctx.set_exemplar(IB)
    ab, warp, sim = ctx.colorize_frames(IA[:, 0:1].cuda(), last.cuda(), want_warp=True)
    assert (sim.cpu().double() - sim64).abs().max() < 2e-5
    clear = (gap > 1e-5).view(1, 1, 16, 16).expand(1, 3, 16, 16)
    # the pooled exemplar colours are fp32 sums on the GPU and fp64 sums in this oracle: compare with a tolerance
    # far below the distance between two different exemplar colours
    assert (warp.cpu()[:, :, ::4, ::4][clear].double() - warped64[:, :, ::4, ::4][clear]).abs().max() < 1e-4
    floor = (ab32.double() - ab64).abs().max().item()
    assert (ab.cpu().double() - ab64).abs().max().item() <= max(1e-3, 2 * floor)
