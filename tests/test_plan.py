"""The balanced partition of the fused step (hydro_gen_b200/csrc/hg_plan.cuh, applied on the device by
k_plan_segments): the same functions compiled for the host (tests/host_emul).  Whatever durations the CTAs
report, the next plan must tile every strip exactly with segments of at least min_rows rows and use exactly
n_cta items; with a fixed cost landscape the cut converges to segments of equal cost within a few steps."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host_emul")], check=True)
    E = C.CDLL(os.path.join(HERE, "host_emul", "libhg_emul.so"))
    E.emul_plan.restype = C.c_int
    return E


def uniform_plan(n_cta, nstrips, row0, rows):
    items = []
    for k in range(nstrips):
        n = n_cta // nstrips + (1 if k < n_cta % nstrips else 0)
        for m in range(n):
            items.append((k, row0 + rows * m // n, row0 + rows * (m + 1) // n))
    return np.array(items, np.int32)


def replan(E, plan, ns, nstrips, row0, rows, min_rows):
    out = np.zeros_like(plan)
    rc = E.emul_plan(len(plan), nstrips, row0, rows, min_rows, plan.ctypes.data_as(C.c_void_p),
                     np.ascontiguousarray(ns, np.uint32).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def check_tiling(plan, nstrips, row0, rows, min_rows, n_cta):
    assert len(plan) == n_cta
    assert (np.diff(plan[:, 0]) >= 0).all()                    # grouped by strip, in order
    for s in range(nstrips):
        seg = plan[plan[:, 0] == s]
        assert len(seg) >= 1
        assert seg[0, 1] == row0 and seg[-1, 2] == row0 + rows
        assert (seg[1:, 1] == seg[:-1, 2]).all()               # no gap, no overlap
        assert ((seg[:, 2] - seg[:, 1]) >= min_rows).all()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_any_durations_give_a_valid_plan(emul, seed):
    rng = np.random.default_rng(seed)
    n_cta, nstrips, row0, rows, min_rows = 444, 36, 4096, 4096, 48     # a slab that does not start at row 0
    plan = uniform_plan(n_cta, nstrips, row0, rows)
    check_tiling(plan, nstrips, row0, rows, min_rows, n_cta)
    for it in range(12):
        kind = it % 4
        if kind == 0: ns = rng.integers(1, 10**6, n_cta)
        elif kind == 1: ns = np.where(rng.random(n_cta) < 0.5, 0, rng.integers(1, 10**9, n_cta))     # zeros and huge values
        elif kind == 2: ns = np.full(n_cta, 500000); ns[plan[:, 0] == 7] = 4 * 10**9 // 1000         # one strip 8000 x dearer
        else: ns = rng.integers(400000, 600000, n_cta)
        plan = replan(emul, plan, ns, nstrips, row0, rows, min_rows)
        check_tiling(plan, nstrips, row0, rows, min_rows, n_cta)


def test_converges_to_equal_cost_on_a_fixed_landscape(emul):
    """cost per row = a smooth function of (strip, row) + a fixed cost per segment (pipeline fill): after a few
    re-cuts the dearest segment is within a few per cent of the mean (one whole segment per strip is the grain)."""
    n_cta, nstrips, row0, rows, min_rows = 444, 36, 0, 4096, 48
    y = np.arange(rows)
    dens = np.array([(1.0 + 0.5 * np.sin(y / 700.0 + s) ** 2) * (1.0 + 0.3 * s / nstrips) + (0.6 if s in (0, 35) else 0.0) for s in range(nstrips)])
    cum = np.concatenate([np.zeros((nstrips, 1)), np.cumsum(dens, axis=1)], axis=1)

    def durations(plan):
        return np.array([1000.0 * (cum[s, b] - cum[s, a] + 17.0 * dens[s, a]) for s, a, b in plan])

    plan = uniform_plan(n_cta, nstrips, row0, rows)
    d0 = durations(plan)
    for _ in range(6):
        plan = replan(emul, plan, durations(plan), nstrips, row0, rows, min_rows)
        check_tiling(plan, nstrips, row0, rows, min_rows, n_cta)
    d = durations(plan)
    assert d0.max() / d0.mean() > 1.5                          # the uniform cut is badly unbalanced on this landscape
    assert d.max() / d.mean() < 1.10                           # the re-cut one is not
    counts = np.bincount(plan[:, 0], minlength=nstrips)
    assert counts[35] > counts[1]                              # dear strips get more segments
