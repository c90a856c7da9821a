"""BASELINE config 4 on N GPUs (torchrun): the 736x1280 correlation (N = 58880 positions) with query rows sharded
over the ranks and the fused all-gather over peer memory; prints device time (max over ranks) and parity vs rank 0's
unsharded run.  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/corr_rowshard_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "deep-exemplar-based-video-colorization_b200"))
import torch
import torch.distributed as dist
import dvc
from dvc.clip import RowShardedCorrelation

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
ctx = dvc.get_context(local)
for side, T in ((184 * 320, 1e-10), (184 * 320, 0.01), (120 * 216, 1e-10)):
    N = side
    g = torch.Generator().manual_seed(7)
    th = torch.nn.functional.normalize(torch.randn(1, 256, N, generator=g), dim=1).cuda()
    ph = torch.nn.functional.normalize(torch.randn(1, 256, N, generator=g), dim=1).cuda()
    V = torch.randn(1, N, 3, generator=g).cuda()
    full_y, full_sim = ctx.corr_softmax_warp(th, ph, V, T)
    sh = RowShardedCorrelation(ctx, N)
    y, sim = sh(th, ph, V, T)  # warm-up + parity
    same = bool(torch.equal(y, full_y) and torch.equal(sim, full_sim))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 5
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ctx.corr_set_peer_outputs(sh._y4, sh._sim, sh.row0)
        ctx.corr_softmax_warp(th[:, :, sh.row0:sh.row1].contiguous(), ph, V, T)
    ctx.corr_set_peer_outputs()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    e0.record()
    for _ in range(reps):
        ctx.corr_softmax_warp(th, ph, V, T)
    e1.record(); torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1) / reps
    sh.close()
    if rank == 0:
        print(f"corr N={N} T={T:g}: 1 GPU {ms1:.3f} ms; {world} GPUs row-sharded + fused all-gather {float(ms):.3f} ms "
              f"(incl. transposes of the shard), identical={same}", flush=True)
if world > 1:
    dist.destroy_process_group()
