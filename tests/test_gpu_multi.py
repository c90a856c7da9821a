"""GPU, world size 2: the query-row-sharded correlation with its fused all-gather over peer memory
(dvc/clip.py: RowShardedCorrelation) must reproduce the single-GPU kernel bit for bit -- every row of
NonlocalNet.py:477-498 is independent.  Skipped on boxes with one GPU."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, NB, T, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "deep-exemplar-based-video-colorization_b200"))
    import torch.distributed as dist
    import dvc
    from dvc.clip import RowShardedCorrelation

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ctx = dvc.get_context(rank)
    g = torch.Generator().manual_seed(7)
    th = torch.nn.functional.normalize(torch.randn(1, 256, N, generator=g), dim=1).cuda()
    ph = torch.nn.functional.normalize(torch.randn(1, 256, NB, generator=g), dim=1).cuda()
    V = torch.randn(1, NB, 3, generator=g).cuda()
    full_y, full_sim = ctx.corr_softmax_warp(th, ph, V, T)
    sharded = RowShardedCorrelation(ctx, N)
    y, sim = sharded(th, ph, V, T)
    y_again, sim_again = sharded(th, ph, V, T)  # second call: the other result set of the double buffer
    if T < 1e-9:  # one-hot rows: bit for bit
        ok = bool(torch.equal(y, full_y) and torch.equal(sim, full_sim) and torch.equal(y_again, y))
    else:  # softmax: a shard may pick another column-split count = another fp32 summation order of the same terms
        ok = bool((y - full_y).abs().max() < 1e-4 and torch.equal(sim, full_sim) and torch.equal(y_again, y))
    sharded.close()
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N,NB,T", [(1000, 517, 1e-10), (777, 900, 0.01)])
def test_row_sharded_correlation_two_gpus(tmp_path, N, NB, T):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, N, NB, T, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
