"""Frame-range sharding of the joint optimisation across the GPUs of one box.

Per-frame pose parameters are independent; the silhouette term couples frames only through the constants
sum(keep_mask) and B (utils/losses.py:71,75; all-reduced once at setup) and the smoothness term couples frame b
only to b-1 and b+1 (utils/losses.py:81).  So each rank owns a contiguous frame range and, once per iteration,
swaps the 9 pose floats (rot6d + translation) of its first / last frame with its neighbours; with
optimize_object_scale the one shared parameter (jointopt.py:42-46) additionally needs its gradient summed over
all ranks.  No other data-path collective exists.  Works on any backend (nccl on GPUs, gloo in the CPU tests).

Ranges are cut by COST, not by count: a frame's rasterisation / backward time depends on what the object looks
like in it, and the one-iteration pose dependency between neighbouring ranges makes every rank step at the pace of
the slowest one.  `balanced_bounds` turns per-block timings of a one-iteration probe (dh_jointopt_probe) into
range boundaries of equal cost.
"""
import numpy as np
import torch
import torch.distributed as dist


class FrameShard:
    """Contiguous range [start, stop) of `B_total` frames owned by `rank` out of `world` ranks.  Without `bounds`
    the frames are split by count; `bounds` (world + 1 ascending frame numbers, 0 ... B_total) fixes every rank's
    range explicitly (cost-weighted partition)."""

    def __init__(self, rank=0, world=1, B_total=0, bounds=None):
        if not (0 <= rank < world):
            raise ValueError("rank out of range")
        if B_total < world:
            raise ValueError(f"cannot shard {B_total} frames over {world} ranks")
        self.rank, self.world, self.B_total = rank, world, B_total
        if bounds is None:
            base, rem = divmod(B_total, world)
            bounds = [r * base + min(r, rem) for r in range(world + 1)]
        bounds = [int(b) for b in bounds]
        if len(bounds) != world + 1 or bounds[0] != 0 or bounds[-1] != B_total or \
                any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
            raise ValueError(f"bad shard bounds {bounds} for {B_total} frames over {world} ranks")
        self.bounds = bounds
        self.start, self.stop = bounds[rank], bounds[rank + 1]

    @property
    def B(self):
        return self.stop - self.start

    @property
    def has_prev(self):
        return self.rank > 0

    @property
    def has_next(self):
        return self.rank < self.world - 1

    def of_rank(self, r):
        return FrameShard(r, self.world, self.B_total, self.bounds)

    def with_bounds(self, bounds):
        return FrameShard(self.rank, self.world, self.B_total, bounds)

    def slice(self, seq):
        return seq[self.start:self.stop]

    def __repr__(self):
        return f"FrameShard(rank={self.rank}/{self.world}, frames=[{self.start},{self.stop}) of {self.B_total})"


def detect_shard(B_total):
    """Shard description from the default process group (single shard when torch.distributed is not set up)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return FrameShard(dist.get_rank(), dist.get_world_size(), B_total)
    return FrameShard(0, 1, B_total)


PROBE_MIN_ITERATIONS = 64


def wants_probe(balance, num_iterations, world):
    """joint_optimize's partition policy.  The cost probe (two passes of the heavy kernels over the rank's equal-count
    range + an all-gather) costs about as much as five iterations and makes every iteration about a tenth faster
    (measured at 8 GPUs, 4096 frames: 3.2 instead of 3.6 ms), so "auto" asks for it from PROBE_MIN_ITERATIONS
    iterations on; "probe" always, "count" never; a single rank has nothing to cut."""
    if balance not in ("auto", "probe", "count"):
        raise ValueError(f"balance must be 'auto', 'probe' or 'count', not {balance!r}")
    if world <= 1 or balance == "count":
        return False
    return balance == "probe" or int(num_iterations) >= PROBE_MIN_ITERATIONS


def balanced_bounds(frame_cost, world):
    """Contiguous partition of len(frame_cost) frames into `world` ranges of (nearly) equal total cost.
    Boundary r is the frame at which the running cost crosses r/world of the total (rounded to the nearer frame);
    every range keeps at least one frame.  Deterministic: all ranks compute the same bounds from the same costs."""
    c = np.maximum(np.asarray(frame_cost, np.float64), 0.0)
    n = len(c)
    if n < world:
        raise ValueError(f"cannot shard {n} frames over {world} ranks")
    if not np.isfinite(c).all() or c.sum() <= 0.0:
        c = np.ones(n)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        i = int(np.searchsorted(cum, target))          # first prefix >= target
        if i > 0 and target - cum[i - 1] < cum[i] - target:
            i -= 1
        i = min(max(i, bounds[-1] + 1), n - (world - r))
        bounds.append(i)
    bounds.append(n)
    return bounds


def frame_costs_from_blocks(block_ms, start, stop, extra_ms=0.0):
    """Per-frame cost of frames [start, stop) from the probe's per-block times: the blocks are the equal cuts
    dh_jointopt_probe makes (block k = frames B*k/n ... B*(k+1)/n of the range), the cost is spread evenly inside a
    block; `extra_ms` (a kernel timed over the whole range) is spread evenly over all frames."""
    B = stop - start
    n = len(block_ms)
    cost = np.empty(B, np.float64)
    for k in range(n):
        b0, b1 = B * k // n, B * (k + 1) // n
        cost[b0:b1] = float(block_ms[k]) / max(b1 - b0, 1)
    return cost + float(extra_ms) / B


def rescale_costs(frame_cost, bounds, rank_ms):
    """Per-frame costs whose sum over every rank's current range equals that rank's MEASURED time per iteration
    (rank_ms, waits excluded): the probe's cost profile inside a range, the run's own clock between ranges."""
    c = np.asarray(frame_cost, np.float64).copy()
    for r, ms in enumerate(rank_ms):
        a, b = bounds[r], bounds[r + 1]
        c[a:b] *= float(ms) / max(c[a:b].sum(), 1e-30)
    return c


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


def allgather_equal(t, shard, group=None):
    """[n, ...] per rank (same shape everywhere) -> [world, n, ...] on every rank."""
    if shard.world == 1:
        return t.unsqueeze(0)
    outs = [torch.empty_like(t) for _ in range(shard.world)]
    dist.all_gather(outs, t.contiguous(), group=group)
    return torch.stack(outs, 0)


def allgather_frames(local, shard, group=None):
    """Concatenate per-rank frame tensors [B_r, ...] into [B_total, ...] on every rank (ragged ranges allowed)."""
    if shard.world == 1:
        return local
    sizes = [shard.bounds[r + 1] - shard.bounds[r] for r in range(shard.world)]
    base = max(sizes)
    pad = torch.zeros((base,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(shard.world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([outs[r][: sizes[r]] for r in range(shard.world)], 0)


# --------------------------------------------------------------------------------------------- exact sums
_TWO64 = 18446744073709551616.0


def fx128_from_float(x):
    """dh_core.h::fx_from_double on the host: (hi, lo) with value = hi + lo * 2^-64 (hi signed, floor)."""
    import math
    if not (abs(x) < 4.0e18):
        return 0, 0
    fl = math.floor(x)
    fr = x - fl
    if fr >= 1.0:
        return int(fl) + 1, 0
    return int(fl), int(fr * _TWO64)


def fx128_sum(parts):
    """Sum of (hi, lo) pairs (unsigned 64-bit words as python ints or a [n,2] int64 / uint64 array) -> float, the
    value k_finalize / k_scale_apply compute (wrap-around arithmetic on 128 bits, then one rounding)."""
    total = 0
    for hi, lo in parts:
        hi, lo = int(hi), int(lo)
        if hi >= 1 << 63:
            hi -= 1 << 64
        if lo < 0:
            lo += 1 << 64
        total += (hi << 64) + lo
    total = ((total + (1 << 127)) % (1 << 128)) - (1 << 127)
    hi, lo = total >> 64, total & ((1 << 64) - 1)
    return float(hi) + float(lo) * (1.0 / _TWO64)


# --------------------------------------------------------------------------------------------- peer mailboxes
class PeerMailboxes:
    """This rank's mailbox (include/dynhor_b200.h, "Mailbox layout") and the CUDA-IPC mappings of the mailboxes of
    all ranks of the box, set up ONCE per process and process group and reused by every later run: the handles
    travel in one tensor all_gather, nothing is freed or re-opened between runs, and runs are kept apart by their
    tick ranges (`reserve`).  `get` returns None -- on every rank alike -- when the ranks do not share a host or a
    pair of devices cannot reach each other directly; the caller then uses the host-driven exchange."""

    _cache = {}

    def __init__(self, mailbox, peers, shard):
        self.mailbox, self.peers, self.rank, self.world = mailbox, peers, shard.rank, shard.world
        self.next_tick = 4

    @classmethod
    def get(cls, shard, group=None):
        import ctypes
        import socket
        import zlib
        from . import _lib
        key = (torch.cuda.current_device(), id(group) if group is not None else 0, shard.world, shard.rank)
        if key in cls._cache:
            return cls._cache[key]
        lib = _lib.load()
        dev = torch.cuda.current_device()
        import os
        ok = shard.world <= _lib.MAX_RANKS and not os.environ.get("DH_FORCE_NO_P2P")   # (test knob: take the fallback)
        mb = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        if ok:
            _lib.check(lib.dh_dev_alloc(ctypes.byref(mb), 4 * _lib.MAILBOX_WORDS), "dh_dev_alloc")
            _lib.check(lib.dh_ipc_export(mb, handle), "dh_ipc_export")
        # one all_gather: IPC handle (64 B) + host id (4 B) + device index (1 B) of every rank
        host = zlib.crc32(socket.gethostname().encode()) & 0xFFFFFFFF
        rec = torch.tensor(list(handle.raw) + list(host.to_bytes(4, "little")) + [dev, int(ok)], dtype=torch.uint8,
                           device="cuda")
        recs = allgather_equal(rec, shard, group).cpu().numpy()
        ok = bool(recs[:, 69].all()) and len({bytes(r[64:68]) for r in recs}) == 1
        if ok:
            for r in range(shard.world):
                d = int(recs[r, 68])
                if r != shard.rank and d != dev and not torch.cuda.can_device_access_peer(dev, d):
                    ok = False
        flag = torch.tensor([int(ok)], dtype=torch.int32, device="cuda")
        if shard.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if mb.value:
                lib.dh_dev_free(mb)
            cls._cache[key] = None
            return None
        peers = []
        for r in range(shard.world):
            if r == shard.rank:
                peers.append(mb.value)
                continue
            ptr = ctypes.c_void_p()
            _lib.check(lib.dh_ipc_open(ctypes.create_string_buffer(bytes(recs[r, :64]), 64), ctypes.byref(ptr)),
                       "dh_ipc_open")
            peers.append(ptr.value)
        cls._cache[key] = cls(mb.value, peers, shard)
        return cls._cache[key]

    def reserve(self, n_ticks):
        """First tick of a run that will use at most n_ticks iterations.  Every rank makes the same calls in the
        same order, so the bases agree without communication.  A run has to be over on ALL ranks before the next one
        is seeded (any collective after it -- the history all-reduce, a barrier -- guarantees that): its last publish
        may share a slot with the next run's seed."""
        base = self.next_tick
        self.next_tick = base + int(n_ticks) + 4
        if self.next_tick > 2 ** 31 - 2 ** 24:  # pragma: no cover  (flags are int32; ~2e9 iterations per process)
            raise RuntimeError("PeerMailboxes: tick counter exhausted")
        return base

    def seed(self, side, tick, pose9):
        """Put the neighbour pose valid for `tick` into this rank's own slot (stream-ordered)."""
        import ctypes
        from . import _lib
        lib, st = _lib.load(), _lib.stream_ptr()
        slot = self.mailbox + 4 * ((side * 4 + (tick & 3)) * 16)
        flag = self.mailbox + 4 * (128 + side * 4 + (tick & 3))
        t = torch.tensor([tick], dtype=torch.int32, device=pose9.device)
        _lib.check(lib.dh_memcpy_d2d(ctypes.c_void_p(slot), _lib.ptr(pose9), 9 * 4, st), "dh_memcpy_d2d")
        _lib.check(lib.dh_memcpy_d2d(ctypes.c_void_p(flag), _lib.ptr(t), 4, st), "dh_memcpy_d2d")
        return t  # keep alive until the copy has run
