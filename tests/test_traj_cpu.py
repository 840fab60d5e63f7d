"""CPU tests of the trajectory source (N4): libgfn.so's DCD reader (host code, no GPU needed) against the arrays
written by oracle/dcd.py and against the oracle's independent numpy reader."""
import os

import numpy as np
import pytest

import dcd as D


def DCDFile(path):
    from newanalysis.gfunction import DCDFile as F
    return F(path)


def _frames(nf, na, seed, L=40.0):
    rng = np.random.default_rng(seed)
    pos = (rng.random((nf, na, 3)) * L).astype(np.float32)
    cells = np.concatenate([L * (1 + 1e-3 * rng.standard_normal((nf, 3))), np.full((nf, 3), 90.0)], axis=1)
    return pos, cells


@pytest.mark.parametrize("big_endian", [False, True])
@pytest.mark.parametrize("with_cell", [False, True])
@pytest.mark.parametrize("fourth_dim", [False, True])
def test_reader_returns_what_was_written(tmp_path, big_endian, with_cell, fourth_dim):
    pos, cells = _frames(7, 113, 5)
    p = str(tmp_path / "t.dcd")
    D.write_dcd(p, pos, cells if with_cell else None, big_endian=big_endian, fourth_dim=fourth_dim,
                titles=("REMARKS one", "REMARKS two"))
    t = DCDFile(p)
    ora = D.DCD(p)
    assert (t.n_atoms, t.n_frames, t.has_cell) == (113, 7, with_cell) == (ora.n_atoms, ora.n_frames, ora.has_cell)
    assert len(t) == 7
    idx = np.array([5, 0, 112, 7, 7], dtype=np.int32)
    for f in (0, 3, 6):
        xyz, cell = t.read(f)
        oxyz, ocell = ora.read(f)
        assert xyz.dtype == np.float32 and np.array_equal(xyz, pos[f]) and np.array_equal(xyz, oxyz)
        if with_cell:
            assert np.array_equal(cell, cells[f]) and np.array_equal(cell, ocell)
        else:
            assert cell is None and ocell is None
        sub, _ = t.read(f, idx)
        assert np.array_equal(sub, pos[f][idx])
    t.close()
    with pytest.raises(ValueError, match="closed"):
        t.read(0)


def test_xplor_header_and_cosine_cells(tmp_path):
    pos, cells = _frames(3, 20, 6)
    p = str(tmp_path / "x.dcd")
    D.write_dcd(p, pos, None, charmm=False)
    t = DCDFile(p)
    assert (t.n_atoms, t.n_frames, t.has_cell) == (20, 3, False)
    assert np.array_equal(t.read(2)[0], pos[2])
    cells[:, 3:] = [[90.0, 90.0, 60.0]] * 3
    D.write_dcd(p, pos, cells, cell_as_cosines=True)
    t = DCDFile(p)
    got = t.read(1)[1]
    assert np.array_equal(got[:3], cells[1, :3]) and np.allclose(got[3:], [90.0, 90.0, 60.0], atol=1e-9)
    assert np.allclose(got, D.DCD(p).read(1)[1], atol=1e-12)


def test_truncated_and_overlong_header_counts(tmp_path):
    pos, cells = _frames(5, 50, 7)
    p = str(tmp_path / "t.dcd")
    D.write_dcd(p, pos, cells, header_frames=9)              # a run that was cut short: header promises more
    assert DCDFile(p).n_frames == 5
    D.write_dcd(p, pos, cells, header_frames=0)              # NAMD writes 0 until the run ends
    assert DCDFile(p).n_frames == 5
    size = os.path.getsize(p)
    with open(p, "r+b") as f:
        f.truncate(size - 100)                               # last frame incomplete
    t = DCDFile(p)
    assert t.n_frames == 4
    with pytest.raises(ValueError, match="out of range"):
        t.read(4)
    with pytest.raises(ValueError, match="out of range"):
        t.read(0, [50])


def test_refused_files(tmp_path):
    pos, cells = _frames(2, 10, 8)
    p = str(tmp_path / "t.dcd")
    D.write_dcd(p, pos, cells, fixed_atoms=3)
    with pytest.raises(ValueError, match="fixed atoms"):
        DCDFile(p)
    with open(p, "wb") as f:
        f.write(b"not a trajectory" * 20)
    with pytest.raises(ValueError, match="not a DCD"):
        DCDFile(p)
    with pytest.raises(ValueError, match="cannot open"):
        DCDFile(str(tmp_path / "missing.dcd"))
    D.write_dcd(p, pos, cells)
    raw = bytearray(open(p, "rb").read())
    t0 = DCDFile(p)
    # damage the leading marker of the X record of frame 1
    off = len(raw) - (56 + 3 * (8 + 40)) + 56
    raw[off] ^= 0x55
    open(p, "wb").write(bytes(raw))
    t = DCDFile(p)
    assert np.array_equal(t.read(0)[0], pos[0])
    with pytest.raises(ValueError, match="damaged"):
        t.read(1)
    del t0


@pytest.mark.parametrize("natoms", [2 ** 29 + 7, 2 ** 31 - 1, 50000])
def test_header_atom_count_is_checked_against_the_file(tmp_path, natoms):
    """natoms comes from the (untrusted) header: a count whose coordinate record would not fit an int32 marker, or
    whose frame is larger than what is left of the file, is refused at open instead of wrapping / reading wild."""
    import struct
    pos, _ = _frames(2, 10, 9)
    p = str(tmp_path / "t.dcd")
    D.write_dcd(p, pos)
    raw = bytearray(open(p, "rb").read())
    # the atom-count record: int32 4 | natoms | int32 4, right before frame 0
    off = len(raw) - 2 * 3 * (8 + 40) - 12
    assert struct.unpack("<iii", raw[off:off + 12]) == (4, 10, 4)
    raw[off + 4:off + 8] = struct.pack("<i", natoms)
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="less than one frame"):
        DCDFile(p)


def test_synth_writer_is_read_back_by_both_readers(tmp_path):
    """gfn_b200.synth.write_dcd (input generation for bench.py / examples) against the native reader and the oracle's."""
    from gfn_b200 import synth
    pos, cells = _frames(4, 257, 9)
    for with_cell in (False, True):
        p = str(tmp_path / ("s%d.dcd" % with_cell))
        synth.write_dcd(p, pos, cells if with_cell else None)
        t, ora = DCDFile(p), D.DCD(p)
        assert (t.n_atoms, t.n_frames, t.has_cell) == (257, 4, with_cell) == (ora.n_atoms, ora.n_frames, ora.has_cell)
        for f in range(4):
            xyz, cell = t.read(f)
            assert np.array_equal(xyz, pos[f]) and np.array_equal(ora.read(f)[0], pos[f])
            if with_cell:
                assert np.array_equal(cell, cells[f]) and np.array_equal(ora.read(f)[1], cells[f])
        # the same bytes as the checker's writer produces for this simplest variant, apart from the title text
        q = str(tmp_path / "o.dcd")
        D.write_dcd(q, pos, cells if with_cell else None, titles=("synthetic trajectory (gfn_b200.synth.write_dcd)",))
        assert open(p, "rb").read() == open(q, "rb").read()
