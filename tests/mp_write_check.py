"""Run under torchrun (one process per GPU) by tests/test_gpu_parity.py::test_multi_process_write_normalises_over_all_ranks:
every rank analyses its own block of frames of an NpT run, joins the NCCL communicator and calls write(); the files of
every rank must be the g-functions of the WHOLE run (histogram, frame count and mean volume summed over the ranks).
usage: mp_write_check.py <outdir>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mdy-newanalysis-package_b200"))

# torch is imported in main() only: the pytest process imports this module for frames() after libgfn has loaded the
# system NCCL, and torch's own libtorch_cuda.so must not be resolved against that one


def frames(nf, n, L, seed=97):
    rng = np.random.default_rng(seed)
    c = rng.random((nf, n, 3)) * L
    d = rng.standard_normal((nf, n, 3))
    boxes = L * (1.0 + 2e-3 * rng.standard_normal((nf, 3)))
    return c, d, boxes


def main():
    import torch
    import torch.distributed as dist
    import newanalysis.gfunction as G
    from gfn_b200 import sharding
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, L, nf = 1500, 40.0, 11                   # 11 frames over 2 ranks: 6 + 5
    c, d, boxes = frames(nf, n, L)
    mine = sharding.shard_frames(nf, rank, world)
    r = G.RDF(["all"], 0.0, 15.0, 0.05, n, n, n, n, L, devices=[local])
    r.attach_comm(sharding.broadcast_unique_id(dist, G.comm_unique_id, device="cuda"), world, rank)
    for f in mine:
        r.update_box(*[float(v) for v in boxes[f]])
        r.calcFrame(c[f], c[f], d[f], d[f])
    r.write(os.path.join(out, "rank%d" % rank))
    np.save(os.path.join(out, "rank%d_out.npy" % rank), r.histogram_out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
