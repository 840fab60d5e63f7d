"""N > 1 host logic on CPU: world_size-2 gloo processes shard frames, run the ORACLE on their shard
(the checker stands in for the GPU here - these tests are about the sharding / combination plumbing),
all-reduce the histograms and must reproduce the single-process result: counts exactly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_partition():
    from gfn_b200.sharding import shard_frames
    for nframes in (0, 1, 7, 8, 1000, 10001):
        for world in (1, 2, 3, 8):
            owned = [f for r in range(world) for f in shard_frames(nframes, r, world)]
            assert owned == list(range(nframes))
            sizes = [len(shard_frames(nframes, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_frames(10, 2, 2)


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "mdy-newanalysis-package_b200"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import oracle as O
    import cases as C
    from gfn_b200.sharding import shard_frames, allreduce_histogram, broadcast_unique_id, allreduce_max
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = broadcast_unique_id(dist, lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        assert allreduce_max(dist, float(rank)) == world - 1
        case = dict(C.CASES["ortho_npt"], frames=5)
        rdf = C.make_rdf(O.OracleRDF, case)
        for f in shard_frames(case["frames"], rank, world):
            c1, c2, d1, d2, newbox = C.build_inputs(case, f)
            rdf.update_box(*newbox)
            rdf.calcFrame(c1, c2, d1, d2)
        total = allreduce_histogram(dist, rdf.raw_sum())
        np.save(os.path.join(out_dir, "rank%d.npy" % rank), total)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    import socket
    import torch.multiprocessing as mp
    import oracle as O
    import cases as C
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    case = dict(C.CASES["ortho_npt"], frames=5)
    ref = C.drive(C.make_rdf(O.OracleRDF, case), case).raw_sum()
    hn = ref.size // 7
    for r in range(2):
        got = np.load(str(tmp_path / ("rank%d.npy" % r)))
        assert np.array_equal(got[:hn], ref[:hn])                       # counts exact for any rank count
        assert np.allclose(got, ref, rtol=0, atol=1e-12 * max(1.0, ref[:hn].max()))
