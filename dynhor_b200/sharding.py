"""Frame-range sharding of the joint optimisation across the GPUs of one box.

Per-frame pose parameters are independent; the silhouette term couples frames only through the constants
sum(keep_mask) and B (utils/losses.py:71,75; all-reduced once at setup) and the smoothness term couples frame b
only to b-1 and b+1 (utils/losses.py:81).  So each rank owns a contiguous frame range and, once per iteration,
swaps the 9 pose floats (rot6d + translation) of its first / last frame with its neighbours.  No other
data-path collective exists.  Works on any backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


class FrameShard:
    """Contiguous range [start, stop) of `B_total` frames owned by `rank` out of `world` ranks."""

    def __init__(self, rank=0, world=1, B_total=0):
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        if B_total < world:
            raise ValueError(f"cannot shard {B_total} frames over {world} ranks")
        self.rank, self.world, self.B_total = rank, world, B_total
        base, rem = divmod(B_total, world)
        self.start = rank * base + min(rank, rem)
        self.stop = self.start + base + (1 if rank < rem else 0)

    @property
    def B(self):
        return self.stop - self.start

    @property
    def has_prev(self):
        return self.rank > 0

    @property
    def has_next(self):
        return self.rank < self.world - 1

    def slice(self, seq):
        return seq[self.start:self.stop]

    def __repr__(self):
        return f"FrameShard(rank={self.rank}/{self.world}, frames=[{self.start},{self.stop}) of {self.B_total})"


def detect_shard(B_total):
    """Shard description from the default process group (single shard when torch.distributed is not set up)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return FrameShard(dist.get_rank(), dist.get_world_size(), B_total)
    return FrameShard(0, 1, B_total)


def exchange_halo(first_pose, last_pose, shard, halo_prev, halo_next, group=None):
    """Send this rank's first frame pose to rank-1 and last frame pose to rank+1; receive theirs into
    halo_prev / halo_next (9 floats each: rot6d row-major [3,2] then translation).  One grouped p2p batch."""
    if shard.world == 1:
        return
    ops = []
    if shard.has_prev:
        ops.append(dist.P2POp(dist.isend, first_pose, shard.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, halo_prev, shard.rank - 1, group))
    if shard.has_next:
        ops.append(dist.P2POp(dist.isend, last_pose, shard.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, halo_next, shard.rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def allreduce_sum_(t, shard, group=None):
    if shard.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_frames(local, shard, group=None):
    """Concatenate per-rank frame tensors [B_r, ...] into [B_total, ...] on every rank (ragged ranges allowed)."""
    if shard.world == 1:
        return local
    base = -(-shard.B_total // shard.world)
    pad = torch.zeros((base,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(shard.world)]
    dist.all_gather(outs, pad, group=group)
    parts = [outs[r][: FrameShard(r, shard.world, shard.B_total).B] for r in range(shard.world)]
    return torch.cat(parts, 0)
