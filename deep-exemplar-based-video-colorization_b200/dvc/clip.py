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


class RowShardedCorrelation:
    """Single-frame scaling of K7 (NonlocalNet.py:477-498) over the GPUs of one box (BASELINE config 4).

    Every query row is independent, so rank r takes the contiguous rows `segment_bounds(N, world, r)` of theta_hat
    against the full phi_hat / V.  The result rows are not all-gathered afterwards: the kernel that finalises a row
    stores it into the full-size (y, sim) buffer of EVERY rank through peer-mapped pointers (CUDA IPC allocations,
    NVLink stores), so after one barrier each rank holds the complete result.  One process per GPU
    (torch.distributed: NCCL on the GPU box); world size <= 8.
    """

    def __init__(self, ctx, n_rows, group=None):
        self.ctx, self.group, self.N = ctx, group, int(n_rows)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        if self.world > 8:
            raise ValueError("RowShardedCorrelation: at most 8 ranks (one NVLink box)")
        self.row0, self.row1 = segment_bounds(self.N, self.world, self.rank)
        # my full-size result buffers: y4 [N][4] followed by sim [N], TWO sets used alternately by call parity: a fast
        # peer may already store the rows of call k+1 while this rank still copies the result of call k out of the
        # other set (it cannot reach call k+2 before this rank has passed call k+1's barrier, i.e. finished that copy)
        self._set_bytes = (self.N * 20 + 255) // 256 * 256  # float4 stores: every set starts 16-byte aligned
        self._calls = 0
        self._static = False
        self._own, handle = ctx.peer_buffer_create(2 * self._set_bytes)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, handle, group=group)
        else:
            handles[0] = handle
        self._opened = []
        self._y4, self._sim = [], []
        for r, h in enumerate(handles):
            base = self._own if r == self.rank else ctx.peer_buffer_open(h)
            if r != self.rank:
                self._opened.append(base)
            self._y4.append(base)
            self._sim.append(base + self.N * 16)

    def __call__(self, theta_hat, phi_hat, V, temperature, exemplar_unchanged=False):
        """theta_hat [1,256,N], phi_hat [1,256,NB], V [1,NB,3] (identical on every rank) -> (y [1,N,3], sim [1,N]).

        exemplar_unchanged=True: phi_hat / V are the same tensors with the same contents as in the previous call (the
        exemplar of a clip), so their operand planes are not prepared again."""
        if theta_hat.shape[0] != 1 or theta_hat.shape[2] != self.N:
            raise ValueError("RowShardedCorrelation: theta_hat must be [1,C,N]")
        ctx = self.ctx
        off = (self._calls & 1) * self._set_bytes
        self._calls += 1
        if self.row1 > self.row0:
            ctx.corr_set_peer_outputs([p + off for p in self._y4], [p + off for p in self._sim], self.row0)
            if not exemplar_unchanged or not self._static:
                ctx.debug_flag("corr_phi_static", 1)  # (re)arms the cache: this call prepares the exemplar side
                self._static = True
            try:
                ctx.corr_softmax_warp(theta_hat[:, :, self.row0:self.row1].contiguous(), phi_hat, V, temperature)
            finally:
                ctx.corr_set_peer_outputs()
        torch.cuda.synchronize(ctx.device)
        if self.world > 1:
            dist.barrier(group=self.group)  # every rank's rows have landed in every buffer
        flat = ctx.raw_view(self._own + off, self.N * 5)
        y = flat[: self.N * 4].view(1, self.N, 4)[:, :, :3].clone()
        sim = flat[self.N * 4:].view(1, self.N).clone()
        return y, sim

    def close(self):
        if self.world > 1:
            dist.barrier(group=self.group)  # nobody may still be writing into a buffer that is about to go away
        if self._static:
            self.ctx.debug_flag("corr_phi_static", 0)
            self._static = False
        for p in self._opened:
            self.ctx.peer_buffer_close(p)
        self._opened = []
        if self._own:
            self.ctx.peer_buffer_destroy(self._own)
            self._own = 0
