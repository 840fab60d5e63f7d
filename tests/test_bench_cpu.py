"""bench.py's host logic and its reference arm, without a GPU: the exact-parity relation the timed runs are checked with,
the `config` object both arms print, and the JSON line of `--impl reference` (the reference's own calcHisto on the host
cores) at a reduced size."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mode_mask_matches_the_abi_bits(bench):
    """GFN_MODE_* of include/gfn.h: 000 = 1, 110 = 2, 101 = 4, 011 = 8, 220 = 16, 202 = 32, 022 = 64."""
    assert bench.mode_mask(["all"]) == 127
    assert bench.mode_mask(["000"]) == 1
    assert bench.mode_mask(["000", "110", "101", "011", "220"]) == 31
    import re
    hdr = open(os.path.join(ROOT, "include", "gfn.h")).read()
    for name, bit in (("000", 1), ("110", 2), ("101", 4), ("011", 8), ("220", 16), ("202", 32), ("022", 64)):
        assert re.search(r"#define\s+GFN_MODE_%s\s+%d\b" % (name, bit), hdr), name
        assert bench.mode_mask([name]) == bit


def test_public_config_is_the_same_object_for_both_arms(bench):
    """the driver compares the two arms' `config`; it holds the workload only (no run parameters)."""
    for name in ("cfg1", "cfg2", "cfg3", "cfg5"):
        c = bench.cfg(name)
        pc = bench.public_config(name, c)
        assert set(pc) == {"workload", "n", "L", "modes", "bins", "l2"}
        assert pc["workload"] == name and pc["n"] == c["n1"] and pc["bins"] == bench.HN
        assert "frames_per_step" not in pc
    assert bench.public_config("cfg3", bench.cfg("cfg3"))["n"] == 32768


def test_counts_ok_relation_against_the_oracle(bench):
    """The timed runs are checked with: g000 row sum == off_diag weight x accepted pairs (gfunction.pyx:131-134, :149).
    Here the relation itself is checked on oracle results for the three frame kinds, including quirk Q1 (two different
    selections of equal size: the i == j pairs of the triangle weigh 1)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    rng = np.random.default_rng(5)
    n, L = 300, 24.0
    c1 = rng.random((n, 3)) * L; c2 = rng.random((n, 3)) * L; c3 = rng.random((200, 3)) * L
    d = rng.standard_normal((n, 3))

    def accepted(a, b, tri):
        dx = b[None, :, :] - a[:, None, :]
        dx -= np.where(np.abs(dx) > L / 2, np.sign(dx) * L, 0.0)                # the reference's single shift (:144-146)
        dist = np.sqrt((dx * dx).sum(-1))
        ok = ~((dist < 0.0) | (dist > 10.0) | (np.floor(dist * 20.0) == 0))
        if tri:
            ok &= np.triu(np.ones_like(ok, dtype=bool))
        return float(ok.sum())

    # same selection on both sides: every accepted pair weighs 2
    o = O.OracleRDF(["000"], 0.0, 10.0, 0.05, n, n, n, n, L)
    o.calcFrame(c1, c1, d, d)
    g = o.raw_sum()[:200].sum()
    assert bench.counts_ok(g, accepted(c1, c1, True), n, n, True, 1)
    assert not bench.counts_ok(g + 1, accepted(c1, c1, True), n, n, True, 1)
    # rectangular: weight 1
    o = O.OracleRDF(["000"], 0.0, 10.0, 0.05, n, 200, n, 200, L)
    o.calcFrame(c1, c3, d, d[:200])
    assert bench.counts_ok(o.raw_sum()[:200].sum(), accepted(c1, c3, False), n, 200, False, 1)
    # two different selections of equal size (Q1): between 2 * accepted - n and 2 * accepted
    o = O.OracleRDF(["000"], 0.0, 10.0, 0.05, n, n, n, n, L)
    o.calcFrame(c1, c2, d, d)
    acc = accepted(c1, c2, True)
    g = o.raw_sum()[:200].sum()
    assert bench.counts_ok(g, acc, n, n, False, 1)
    assert 2 * acc - n <= g < 2 * acc                   # some i == j pair of the two selections is in range
    assert not bench.counts_ok(2 * acc + 2, acc, n, n, False, 1)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` at a reduced size (GFN_BENCH_N): one JSON line, impl = reference, the CPU baseline and
    a zero-copy e2e object, no GPU launches - and never a device: this runs in the CPU container."""
    env = dict(os.environ, GFN_BENCH_N="2048", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"].startswith("pair evals/sec") and d["unit"] == "1e9 pair evals/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["config"]["workload"] == "cfg3" and d["config"]["n"] == 2048
