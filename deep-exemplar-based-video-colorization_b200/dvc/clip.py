"""Clip-level driver: segment sharding across the GPUs of one box and the once-per-clip exemplar broadcast.

Reference semantics: test.py:68-96 processes the frames of a clip in order and feeds frame
t-1's prediction into frame t, so a clip only shards into CONTIGUOUS segments, each treated like an
independent clip whose first frame starts from zeros (test.py:80).  The only data that must travel between
GPUs is the exemplar-side operand pack (phi_hat [N,256] + pooled Lab [N,4]); rank `src` computes it once and
`torch.distributed.broadcast` (NCCL over NVLink on the GPU box, gloo in the CPU tests) ships it.
"""
import torch
import torch.distributed as dist


def segment_bounds(n_frames, world_size, rank):
    """Contiguous, balanced segments: the first (n_frames % world_size) ranks take one extra frame."""
    if n_frames < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad segment request")
    base, extra = divmod(n_frames, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_exemplar(pack, src=0, group=None):
    """In-place broadcast of the flat exemplar operand pack (a CUDA tensor under NCCL, CPU under gloo)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(pack, src=src, group=group)
    return pack


def prepare_exemplar(ctx, IB_lab, H, W, src=0, group=None):
    """Rank `src` runs the exemplar prologue (test.py:57-66 + WarpNet's B side); everyone else receives the pack."""
    rank = dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0
    if rank == src:
        ctx.set_exemplar(IB_lab)
        pack = ctx.exemplar_export(H, W)
    else:
        pack = torch.empty(ctx.exemplar_pack_size(H, W), device=ctx.device, dtype=torch.float32)
    broadcast_exemplar(pack, src, group)
    if rank != src:
        ctx.exemplar_import(pack, H, W)
    return pack


def colorize_clip_sharded(ctx, host_L, IB_lab, temperature=1e-10, src=0, group=None):
    """Each rank colourises its contiguous segment of `host_L` [F,1,H,W]; returns (start, end, ab[start:end])."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    F_, _, H, W = host_L.shape
    prepare_exemplar(ctx, IB_lab, H, W, src, group)
    s, e = segment_bounds(F_, world, rank)
    if e == s:
        return s, e, torch.empty(0, 2, H, W)
    seg = host_L[s:e].contiguous()
    if not seg.is_pinned():
        seg = seg.pin_memory()
    return s, e, ctx.colorize_clip(seg, temperature)
