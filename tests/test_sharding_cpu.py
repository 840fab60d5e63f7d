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


def _traj_worker(rank, world, port, out_dir, path):
    """A trajectory sharded over ranks: every rank opens the DCD file with the native reader, takes its block of frames
    (host logic of RDF.calcTrajectory under torchrun), runs the oracle chain on it; histograms, frame counts and volume
    sums are all-reduced the way gfn_readout / gfn_allreduce_host do it on the GPUs."""
    for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "mdy-newanalysis-package_b200"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import oracle as O
    import cases as C
    from newanalysis.gfunction import DCDFile
    from gfn_b200.sharding import shard_frames, allreduce_histogram
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = C.ATOM_CASES["atoms_water_self"]
        masses, apr, rfa, charges = C.build_atoms(case)["topo1"]
        n = case["n1"]
        traj = DCDFile(path)
        o = O.OracleRDF(case["modes"], 0.0, case["hmax"], 0.05, n, n, n, n, case["L"])
        for f in shard_frames(traj.n_frames, rank, world):
            xyz, cell = traj.read(f)
            o.update_box(*[float(v) for v in cell[:3]])
            coor = np.ascontiguousarray(xyz, dtype=np.float64)
            com = O.com_by_residue_c(coor, masses, apr, rfa)
            dip = O.dip_by_residue_c(coor, charges, apr, rfa, com)
            o.calcFrame(com, com, dip, dip)
        total = allreduce_histogram(dist, o.raw_sum())
        scal = allreduce_histogram(dist, np.array([float(o.ctr), float(sum(o.volume_list))]))
        np.save(os.path.join(out_dir, "traj_rank%d.npy" % rank), np.concatenate([scal, total]))
    finally:
        dist.destroy_process_group()


def test_two_rank_trajectory_sharding(tmp_path):
    import socket
    import torch.multiprocessing as mp
    import cases as C
    import dcd as D
    import oracle as O
    case = C.ATOM_CASES["atoms_water_self"]
    a = C.build_atoms(case)
    nf, L = case["frames"], case["L"]
    cells = np.concatenate([L * (1 + 1e-3 * np.random.default_rng(3).standard_normal((nf, 3))), np.full((nf, 3), 90.0)], axis=1)
    path = str(tmp_path / "shard.dcd")
    D.write_dcd(path, a["pos1"], cells, big_endian=True)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_traj_worker, args=(2, port, str(tmp_path), path), nprocs=2, join=True)
    masses, apr, rfa, charges = a["topo1"]
    n = case["n1"]
    o = O.OracleRDF(case["modes"], 0.0, case["hmax"], 0.05, n, n, n, n, L)
    for f in range(nf):
        o.update_box(*[float(v) for v in cells[f, :3]])
        coor = np.ascontiguousarray(a["pos1"][f], dtype=np.float64)
        com = O.com_by_residue_c(coor, masses, apr, rfa)
        dip = O.dip_by_residue_c(coor, charges, apr, rfa, com)
        o.calcFrame(com, com, dip, dip)
    ref = o.raw_sum()
    hn = ref.size // 7
    for r in range(2):
        got = np.load(str(tmp_path / ("traj_rank%d.npy" % r)))
        assert got[0] == nf and np.isclose(got[1], sum(o.volume_list), rtol=1e-14)
        assert np.array_equal(got[2:2 + hn], ref[:hn])
        assert np.allclose(got[2:], ref, rtol=0, atol=1e-12 * max(1.0, ref[:hn].max()))
