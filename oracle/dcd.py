"""oracle/dcd.py - numpy writer and reader of CHARMM / NAMD DCD trajectories (checker for csrc/gfn_traj.cpp).

TEST INFRASTRUCTURE.  Only tests/ may import this module; the product reads trajectories with libgfn.so
(gfn_traj_open_dcd / gfn_traj_read / gfn_enqueue_trajectory), bench.py and the examples write their synthetic input
files with gfn_b200.synth.write_dcd.

Parity status: UNPINNED.  The reference reads DCD files through MDAnalysis (docs/notebooks/rdf.ipynb cell 10:
`MDAnalysis.Universe(psf, base + "1mq_swm4.dcd")`), a third-party dependency (unpinned in the reference's setup.py /
devtools environment files) that is neither in /root/reference nor installed in this image, and the reference tree
ships no trajectory file.  What is restated here is the published record layout of the format (the one MDAnalysis'
libdcd and VMD's dcdplugin implement): Fortran unformatted records with 32-bit length markers,

  record 1   84 bytes: "CORD", icntrl[20] (int32): [0] frames, [1] first step, [2] save interval, [3] steps,
             [8] fixed atoms, [9] time step (float32 for CHARMM), [10] unit cell flag, [11] 4-D flag, [19] CHARMM version
  record 2   int32 ntitle, ntitle lines of 80 characters
  record 3   int32 natoms
  per frame  [6 float64: A, gamma, B, beta, alpha, C]  X[natoms] float32, Y[natoms] float32, Z[natoms] float32 [, W]

The writer and the reader are independent code paths (struct.pack vs numpy.frombuffer); the tests additionally check the
native reader against the arrays that were written, so a shared misunderstanding of the layout is the only thing they
cannot see.
"""
import struct

import numpy as np


def write_dcd(path, positions, cells=None, big_endian=False, charmm=True, fourth_dim=False, cell_as_cosines=False,
              header_frames=None, fixed_atoms=0, istart=0, nsavc=1, delta=0.002, titles=("written by oracle/dcd.py",)):
    """positions (nframes, natoms, 3) -> file.  cells (nframes, 6) as A, B, C, alpha, beta, gamma in degrees or None."""
    pos = np.asarray(positions, dtype=np.float32)
    nf, na = pos.shape[0], pos.shape[1]
    e = ">" if big_endian else "<"

    def rec(payload):
        return struct.pack(e + "i", len(payload)) + payload + struct.pack(e + "i", len(payload))

    ic = [0] * 20
    ic[0] = nf if header_frames is None else header_frames
    ic[1], ic[2], ic[3] = istart, nsavc, nf * nsavc
    ic[8] = fixed_atoms
    ic[10] = 1 if (cells is not None and charmm) else 0
    ic[11] = 1 if (fourth_dim and charmm) else 0
    ic[19] = 24 if charmm else 0
    body = b"CORD"
    for k in range(20):
        if k == 9 and charmm:
            body += struct.pack(e + "f", delta)
        elif k == 9:
            body += struct.pack(e + "d", delta)           # X-PLOR: a double occupying icntrl[9..10]
        elif k == 10 and not charmm:
            continue
        else:
            body += struct.pack(e + "i", ic[k])
    assert len(body) == 84
    with open(path, "wb") as f:
        f.write(rec(body))
        tt = struct.pack(e + "i", len(titles)) + b"".join(t.encode()[:80].ljust(80) for t in titles)
        f.write(rec(tt))
        f.write(rec(struct.pack(e + "i", na)))
        for i in range(nf):
            if ic[10]:
                A, B, C, al, be, ga = [float(v) for v in cells[i]]
                if cell_as_cosines:
                    al, be, ga = [np.cos(np.deg2rad(v)) for v in (al, be, ga)]
                f.write(rec(struct.pack(e + "6d", A, ga, B, be, al, C)))
            for k in range(3):
                f.write(rec(pos[i, :, k].astype(e + "f4").tobytes()))
            if ic[11]:
                f.write(rec(np.zeros(na, dtype=e + "f4").tobytes()))


class DCD(object):
    """Independent numpy reader: DCD(path).read(frame) -> (positions (natoms,3) float32, cell (6,) float64 or None)."""

    def __init__(self, path):
        self.buf = np.memmap(path, dtype=np.uint8, mode="r")
        first = int(np.frombuffer(self.buf[:4].tobytes(), "<i4")[0])
        self.e = "<" if first == 84 else ">"
        if int(np.frombuffer(self.buf[:4].tobytes(), self.e + "i4")[0]) != 84:
            raise ValueError("not a DCD file")
        ic = np.frombuffer(self.buf[8:88].tobytes(), self.e + "i4")
        charmm = ic[19] != 0
        self.header_frames = int(ic[0])
        self.has_cell = bool(charmm and ic[10])
        self.has_4d = bool(charmm and ic[11])
        off = 92
        tl = int(np.frombuffer(self.buf[off:off + 4].tobytes(), self.e + "i4")[0])
        off += 8 + tl
        self.n_atoms = int(np.frombuffer(self.buf[off + 4:off + 8].tobytes(), self.e + "i4")[0])
        off += 12
        self.first = off
        self.frame_bytes = (56 if self.has_cell else 0) + (4 if self.has_4d else 3) * (8 + 4 * self.n_atoms)
        present = (self.buf.shape[0] - off) // self.frame_bytes
        self.n_frames = min(present, self.header_frames) if self.header_frames > 0 else present

    def read(self, frame):
        o = self.first + frame * self.frame_bytes
        cell = None
        if self.has_cell:
            v = np.frombuffer(self.buf[o + 4:o + 52].tobytes(), self.e + "f8")
            al, be, ga = v[4], v[3], v[1]
            if abs(al) <= 1 and abs(be) <= 1 and abs(ga) <= 1:
                al, be, ga = [90.0 - np.rad2deg(np.arcsin(c)) for c in (al, be, ga)]
            cell = np.array([v[0], v[2], v[5], al, be, ga])
            o += 56
        n = self.n_atoms
        out = np.empty((n, 3), dtype=np.float32)
        for k in range(3):
            out[:, k] = np.frombuffer(self.buf[o + 4:o + 4 + 4 * n].tobytes(), self.e + "f4")
            o += 8 + 4 * n
        return out, cell
