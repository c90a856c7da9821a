"""CPU tests of the pre / post-processing oracle (oracle/prepost_oracle.py) and of the host-side geometry helper.

The reference ships no vectors for these steps and neither opencv-contrib (ximgproc) nor scikit-image is installed here, so:
 * the Fast Global Smoother restatement is held against what the algorithm is defined to compute -- a float64 sparse direct
   solve of every line's WLS system (Min et al. 2014) -- and against the invariants of those systems;
 * skimage.transform.resize is restated over the scipy.ndimage calls it makes (scipy IS installed: that part is pinned),
   and the written-out formulas the CUDA kernel evaluates must reproduce scipy's result exactly after the uint8 truncation;
 * dvc.prepost.centerpad_geometry (one crop / pad offset pair for the device call) must reproduce CenterPad + CenterCrop
   (utils/util_distortion.py:217-258, test.py:45) including their int() truncations.
"""
import numpy as np
import pytest

from oracle import prepost_oracle as P


def test_fgs_restatement_solves_the_wls_systems():
    rng = np.random.default_rng(0)
    H, W = 40, 56
    g = (rng.random((H, W)) * 255).astype(np.uint8)
    g[10:20] = g[10:11]  # a flat band: weights exactly 1 along it
    src = (rng.standard_normal((2, H, W)) * 30).astype(np.float32)
    a = P.fgs_filter(g, src, 500, 4)
    b = P.fgs_reference_f64(g, src, 500, 4)
    assert np.abs(a - b).max() <= 2e-5 * np.abs(b).max()


def test_fgs_invariants():
    rng = np.random.default_rng(1)
    g = (rng.random((24, 40)) * 255).astype(np.uint8)
    const = np.full((1, 24, 40), 7.25, np.float32)
    assert np.abs(P.fgs_filter(g, const, 500, 4) - 7.25).max() < 1e-3          # constants are fixed points (fp32 Thomas, cond ~ 2000)
    src = (rng.standard_normal((1, 24, 40)) * 10).astype(np.float32)
    out = P.fgs_filter(g, src, 500, 4)
    assert abs(out.sum() - src.sum()) < 1e-3 * np.abs(src).sum()                # 1^T (I + lam L) = 1^T: the sum is kept
    assert np.abs(P.fgs_filter(g, src, 0.0, 4) - src).max() == 0                # lambda = 0 is the identity
    flat = np.zeros((24, 40), np.uint8)
    strong = P.fgs_filter(flat, src, 1e4, 4)
    assert strong.std() < 0.05 * src.std()                                       # flat guide + huge lambda -> nearly the mean
    edge = np.zeros((24, 40), np.uint8)
    edge[:, 20:] = 255                                                           # a hard edge in the guide blocks diffusion
    step = np.zeros((1, 24, 40), np.float32)
    step[:, :, 20:] = 100
    kept = P.fgs_filter(edge, step, 500, 4)
    assert np.abs(kept - step).max() < 1e-2   # exp(-255 / 4) ~ 2e-28 couples the halves; fp32 rounding of 100 +- dominates


def test_guide_from_luminance():
    l = np.array([-50.0, -49.999, 0.0, 12.3, 49.99, 50.0], np.float32)
    assert P.l_to_guide8(l).tolist() == [0, 0, 127, 158, 254, 255]


@pytest.mark.parametrize("hs,ws,size", [(270, 480, (108, 192)), (100, 200, (64, 96)), (200, 100, (64, 96)), (64, 96, (64, 96)),
                                        (50, 70, (64, 96)), (123, 457, (216, 384)), (48, 64, (432, 768))])
def test_resize_restated_equals_scipy_and_geometry_matches_centerpad(hs, ws, size):
    from dvc.prepost import centerpad_geometry

    rng = np.random.default_rng(hs * 1000 + ws)
    img = (rng.random((hs, ws, 3)) * 255).astype(np.uint8)
    ref = P.centerpad_transform(img, size, P.skimage_resize)       # scipy.ndimage arithmetic (pinned)
    mine = P.centerpad_transform(img, size, P.resize_restated)     # the formulas of csrc/prepost.cu
    assert ref.shape == (size[0], size[1], 3) and np.array_equal(ref, mine)
    Hr, Wr, oy, ox = centerpad_geometry(hs, ws, size)
    full = P.resize_restated(img, (Hr, Wr)) if (Hr, Wr) != (hs, ws) else img.astype(np.float64)
    comp = np.zeros((size[0], size[1], 3), np.uint8)
    ys, xs = np.arange(size[0]) + oy, np.arange(size[1]) + ox
    my, mx = (ys >= 0) & (ys < Hr), (xs >= 0) & (xs < Wr)
    comp[np.ix_(my, mx)] = np.clip(np.trunc(full[np.ix_(ys[my], xs[mx])]), 0, 255).astype(np.uint8)
    assert np.array_equal(comp, ref)
