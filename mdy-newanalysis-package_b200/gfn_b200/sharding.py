"""Host-side logic of the frame-sharded multi-GPU driver (one process per GPU).

Frames are independent units of the g-function path: rank r analyses its own frames and
the only exchange is the sum of the small [8][histo_n] histograms.  On GPUs that sum is ONE
ncclAllReduce issued by libgfn (gfn_readout after gfn_comm_attach); the helpers here are the plumbing a
host framework needs around it, written against torch.distributed so they run on gloo (CPU tests) and
nccl (bench.py) alike.
"""
import numpy as np


def shard_frames(nframes, rank, world):
    """Contiguous block of frame indices of `rank`: sizes differ by at most one, blocks are ordered by
    rank, every frame is owned exactly once."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank %r of %r" % (rank, world))
    base, rem = divmod(int(nframes), world)
    begin = rank * base + min(rank, rem)
    return range(begin, begin + base + (1 if rank < rem else 0))


def broadcast_unique_id(dist, make_uid, device=None):
    """Rank 0 calls make_uid() (128 bytes, newanalysis.gfunction.comm_unique_id) and every rank gets it."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        uid = make_uid()
        if len(uid) != 128:
            raise ValueError("unique id must be 128 bytes")
        buf.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())


def allreduce_histogram(dist, hist, device=None):
    """Sum of per-rank un-normalised histograms through torch.distributed (fp64).  The product path
    does this with ncclAllReduce inside libgfn; this host version exists for frameworks that keep the
    histograms on the host and for the CPU tests of the sharding logic."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(hist, dtype=np.float64), device=device).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_max(dist, value, device=None):
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
