"""GPU: the drop-in nn.Modules behind the reference's own per-frame glue.  The reference's
models/FrameColor.py is not available on the GPU box, so its 30 lines of glue are restated here exactly as
the oracle does (oracle.dvc_oracle.frame_colorization cites them)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import dvc_oracle as O

pytestmark = pytest.mark.gpu


def test_dropin_modules_run_reference_glue(sds):
    from models.ColorVidNet import ColorVidNet
    from models.NonlocalNet import VGG19_pytorch, WarpNet

    nonlocal_net, colornet, vggnet = WarpNet(1), ColorVidNet(7), VGG19_pytorch()
    vggnet.load_state_dict(sds["vgg"])                 # test.py:150
    nonlocal_net.load_state_dict(sds["warp"])          # test.py:158
    colornet.load_state_dict(sds["color"])             # test.py:159
    for m in (nonlocal_net, colornet, vggnet):
        for p in m.parameters():
            p.requires_grad = False
        m.eval()
        m.cuda()                                       # test.py:161-166
    g = load_golden("small_32x48")
    IA, IB, last = (torch.from_numpy(g[k]).cuda() for k in ("IA_lab", "IB_lab", "IA_last_lab"))
    with torch.no_grad():
        rgb = O.tensor_lab2rgb(torch.cat((O.uncenter_l(IB[:, 0:1]), IB[:, 1:3]), dim=1).cpu()).cuda()   # test.py:65
        features_B = vggnet(rgb, ["r12", "r22", "r32", "r42", "r52"], preprocess=True)                      # test.py:66
        for _ in range(2):  # second pass exercises the cached exemplar side
            IA_l = IA[:, 0:1]
            fA = vggnet(O.gray2rgb_batch(IA_l), ["r12", "r22", "r32", "r42", "r52"], preprocess=True)       # FrameColor.py:6-10
            An = [O.feature_normalize(t) for t in fA[1:]]
            Bn = [O.feature_normalize(t) for t in features_B[1:]]
            warped, sim = nonlocal_net(IB, *An, *Bn, temperature=1e-10)                                     # FrameColor.py:25-36
            ab = colornet(torch.cat((IA_l, warped[:, 1:3], sim, last), dim=1))                              # FrameColor.py:63-65
    floor = np.abs(g["ab32"].astype(np.float64) - g["ab64"]).max()
    err = np.abs(ab.cpu().numpy().astype(np.float64) - g["ab64"]).max()
    assert err <= max(1e-3, 2 * floor), (err, floor)
    assert np.abs(sim.cpu().numpy()[:, :, ::4, ::4] - g["sim64"]).max() < 2e-5


def test_dropin_refuses_cpu_tensors():
    import dvc
    from models.NonlocalNet import VGG19_pytorch

    with pytest.raises(dvc.DvcError):
        VGG19_pytorch()(torch.zeros(1, 3, 32, 32), ["r12"])
