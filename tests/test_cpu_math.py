"""CPU tests (no GPU): the product's device functions, compiled for the host by
tests/host_emul, must reproduce the oracle bit for bit; the defined-math header must
agree with libm; the monotonicity that the thermal marking threshold relies on holds."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle
from hydro_gen_b200 import _lib as hl
from tests.util import DT_TIME, FIELDS, SEED, assert_bit_equal, wet_world

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "host_emul")], check=True)
    E = C.CDLL(os.path.join(HERE, "host_emul", "libhg_emul.so"))
    E.emul_grid_step.restype = C.c_long
    return E


def planes_of(w):
    H, F, S = w.get(0), w.get(1), w.get(3)
    return [np.ascontiguousarray(a) for a in (H[..., 0], H[..., 1], H[..., 2], F[..., 0], F[..., 1], F[..., 2], F[..., 3], S[..., 0], S[..., 1])]


def emul_step(E, w, pl):
    arr = (C.c_void_p * 9)(*[p.ctypes.data for p in pl])
    er = hl.ErosionData.from_buffer_copy(bytes(w.erosion))
    return E.emul_grid_step(C.byref(er), w.W, w.H, arr)


def test_device_cell_math_matches_oracle_wet(emul):
    """flux, erosion, sediment gather, thermal x2, smoothing and rain of hg_cell.cuh /
    hg_noise.cuh vs the oracle over 24 wet steps (rain every 16) at 192x128."""
    w = wet_world(128, 200, width=192)
    mp = hl.MapSettingsData.from_buffer_copy(bytes(w.map))
    far = 0
    for _ in range(24):
        pl = planes_of(w)
        t = (w.steps + 1) * DT_TIME
        if (w.steps + 1) % 16 == 0:
            rn = hl.RainData.from_buffer_copy(bytes(w.rain))
            emul.emul_rain(C.byref(rn), C.byref(mp), C.c_float(t), w.W, w.H, pl[0].ctypes.data_as(C.c_void_p),
                           pl[1].ctypes.data_as(C.c_void_p), pl[2].ctypes.data_as(C.c_void_p))
        far += emul_step(emul, w, pl)
        w.step(t)
        for got, want, name in zip(pl, planes_of(w), "rock dirt water fL fR fT fB sed_r sed_d".split()):
            assert_bit_equal(got, want, name)
    assert far > 0   # the state exercises back-traces beyond +-1 cell
    w.close()


def test_device_thermal_marking_matches_oracle_steep(emul):
    """Lowered talus angles + roughened terrain: many marked neighbours in both layers, so the
    threshold form of the marking test and the single-atan sharpness are exercised."""
    n = 96
    w = oracle.World(n, seed=SEED); w.gen_heightmap()
    H = w.get(0)
    rng = np.random.default_rng(3)
    H[..., 0] += rng.random((n, n), dtype=np.float32) * 6
    H[..., 1] += rng.random((n, n), dtype=np.float32) * 2
    H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]
    w.set(0, H)
    for ka in ((0.9, 0.3), (1.3, 0.6), (0.05, 1.55), (-0.1, 2.0)):
        w.erosion.Kalpha[0], w.erosion.Kalpha[1] = ka
        for _ in range(3):
            pl = planes_of(w)
            emul_step(emul, w, pl)
            w.dispatch_grid()
            for got, want, name in zip(pl, planes_of(w), "rock dirt water fL fR fT fB sed_r sed_d".split()):
                assert_bit_equal(got, want, f"Kalpha={ka}: {name}")
    w.close()


def test_device_heightmap_matches_oracle(emul):
    for seed, variant in ((SEED, {}), (7.5, {"uplift": 1, "terrace": 5}), (99.0, {"mask_round": 1, "mask_slope": 1, "domain_warp": 2, "mask_exp": 0})):
        n = 96
        w = oracle.World(n, seed=seed)
        for k, v in variant.items():
            setattr(w.map, k, v)
        w.gen_heightmap()
        mp = hl.MapSettingsData.from_buffer_copy(bytes(w.map))
        rock = np.zeros((n, n), np.float32); dirt = np.zeros((n, n), np.float32)
        emul.emul_heightmap(C.byref(mp), n, n, rock.ctypes.data_as(C.c_void_p), dirt.ctypes.data_as(C.c_void_p))
        H = w.get(0)
        assert_bit_equal(rock, H[..., 0], f"rock {variant}")
        assert_bit_equal(dirt, H[..., 1], f"dirt {variant}")
        w.close()


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_defined_math_close_to_libm():
    L = oracle.lib()
    xs = np.concatenate([np.geomspace(1e-6, 1e6, 4000), np.linspace(0.3, 3.0, 4000)]).astype(np.float32)
    at = np.array([L.orc_atanf(float(x)) for x in xs], np.float32)
    assert _ulp_diff(at, np.arctan(xs.astype(np.float64)).astype(np.float32)).max() <= 4
    xe = np.linspace(-5, 5, 4001).astype(np.float32)
    ex = np.array([L.orc_expf(float(x)) for x in xe], np.float32)
    assert _ulp_diff(ex, np.exp(xe.astype(np.float64)).astype(np.float32)).max() <= 4
    xsn = np.concatenate([np.linspace(-20, 20, 4001), np.linspace(1e3, 4.3e5, 4001)]).astype(np.float32)
    sn = np.array([L.orc_sinf(float(x)) for x in xsn], np.float32)
    assert np.abs(sn - np.sin(xsn.astype(np.float64))).max() <= 3e-7


def test_atan_monotone_where_thresholds_live():
    """The exhaustive proof is scripts/check_atan_monotone.c (all 2^31 positive floats, ~1 min);
    here: every float in windows around the range-reduction breakpoints and around tan(Kalpha)
    for the default talus angles."""
    L = oracle.lib()
    for centre in (0.4142135623730950, 2.414213562373095, math.tan(0.6), math.tan(1.3), 1.0, 1e-3, 50.0):
        c = np.float32(centre).view(np.uint32)
        u = np.arange(int(c) - 3000, int(c) + 3000, dtype=np.uint32)
        y = np.array([L.orc_atanf(float(v)) for v in u.view(np.float32)], np.float32)
        assert np.all(np.diff(y) >= 0), centre


def test_struct_layouts_match_the_reference():
    """offsets of glsl/bindings.glsl:39-111"""
    E, M = hl.ErosionData, hl.MapSettingsData
    want = {"particle_count": 0, "Kc": 4, "Kalpha": 8, "Kconv": 16, "Ks": 24, "Kd": 32, "Ke": 40, "ENERGY_KEPT": 44,
            "Kspeed": 48, "G": 56, "d_t": 60, "density": 64, "init_volume": 68, "friction": 72, "inertia": 76,
            "min_volume": 80, "min_velocity": 84, "ttl": 88}
    for k, off in want.items():
        assert getattr(E, k).offset == off, k
    assert C.sizeof(E) == 96 and C.sizeof(hl.RainData) == 20 and C.sizeof(M) == 96
    assert M.hmap_dims.offset == 8 and M.seed.offset == 24 and M.octaves.offset == 44 and M.domain_warp.offset == 76
    assert M.terrace_scale.offset == 88
    assert oracle.PARTICLE_DTYPE.fields["sediment"][1] == 32 and oracle.PARTICLE_DTYPE.fields["to_kill"][1] == 40


def fused_emul_step(E, w, nt, seg, ws=0, two_lane=False, queued=False):
    """one step of the fused kernel body on the CPU; returns (planes, far cell indices)"""
    src = planes_of(w)
    dst = [np.zeros_like(p) for p in src]
    far = np.zeros(w.W * w.H, np.uint32)
    sa = (C.c_void_p * 9)(*[p.ctypes.data for p in src])
    da = (C.c_void_p * 9)(*[p.ctypes.data for p in dst])
    er = hl.ErosionData.from_buffer_copy(bytes(w.erosion))
    fn = E.emul_fusedq_step if queued else E.emul_fused2_step if two_lane else E.emul_fused_step
    fn.restype = C.c_long
    n = fn(C.byref(er), w.W, w.H, nt, seg, ws, sa, da, far.ctypes.data_as(C.c_void_p))
    assert n >= 0
    return dst, far[:n]


@pytest.mark.parametrize("nt,seg,shape,ws", [(32, 16, (72, 64), 0), (32, 64, (40, 136), 0), (128, 32, (192, 96), 0), (128, 128, (128, 160), 0), (224, 64, (264, 80), 0),
                                             (32, 16, (72, 64), 1), (32, 16, (72, 64), 2), (128, 32, (192, 96), 1), (128, 128, (128, 160), 2)])
def test_fused_kernel_body_emulated_matches_oracle(emul, nt, seg, shape, ws):
    """hg_fused_body.cuh — the code k_fused_step runs — executed thread by thread on the CPU with
    the kernel's own iteration plan (generic fill/drain + FREE steady state), against the oracle;
    ws = 1, 2: the warp-specialised split (k_fused_ws) with the two groups in either order: every plane bit-exact; the cells it defers to the far-fetch fix-up are exactly the
    cells whose back-trace leaves the +-1 window."""
    W, H = shape
    w = wet_world(H, 120, width=W, period=8)
    names = "rock dirt water fL fR fT fB sed_r sed_d".split()
    total_far = 0
    for _ in range(3):
        got, far = fused_emul_step(emul, w, nt, seg, ws)
        pl = planes_of(w)
        far_ref = emul_step(emul, w, pl)            # the unfused emulation counts the same cells
        w.step((w.steps + 1) * DT_TIME)
        want = planes_of(w)
        assert len(far) == far_ref
        mask = np.ones(W * H, bool); mask[far] = False
        for k, (g, x, name) in enumerate(zip(got, want, names)):
            if k >= 7:      # sediment of far cells is written by k_far_fixup, not by the main kernel
                g = np.where(mask.reshape(H, W), g, x)
            assert_bit_equal(g, x, f"nt={nt} seg={seg} {name}")
        total_far += len(far)
    w.close()


@pytest.mark.parametrize("nt,seg,shape,ws", [(32, 16, (72, 64), 1), (32, 64, (40, 136), 2), (32, 16, (72, 64), 3), (128, 32, (192, 96), 4),
                                             (128, 128, (128, 160), 2), (128, 48, (240, 64), 1)])
def test_queued_outflow_body_emulated_matches_oracle(emul, nt, seg, shape, ws):
    """hg_fused_body3.cuh -- the code k_fused_q runs: thermal threads only test their cell and queue the marked ones, a
    service role evaluates the outflow of the queued cells one iteration later -- executed thread by thread on the CPU
    with the three roles of an iteration in four different orders, against the oracle: every plane bit-exact.  The
    rings start poisoned, so a value consumed before it was produced would show."""
    W, H = shape
    w = wet_world(H, 120, width=W, period=8)
    names = "rock dirt water fL fR fT fB sed_r sed_d".split()
    for _ in range(3):
        got, far = fused_emul_step(emul, w, nt, seg, ws, queued=True)
        pl = planes_of(w)
        far_ref = emul_step(emul, w, pl)
        w.step((w.steps + 1) * DT_TIME)
        want = planes_of(w)
        assert len(far) == far_ref
        mask = np.ones(W * H, bool); mask[far] = False
        for k, (g, x, name) in enumerate(zip(got, want, names)):
            if k >= 7:
                g = np.where(mask.reshape(H, W), g, x)
            assert_bit_equal(g, x, f"queued nt={nt} seg={seg} ws={ws} {name}")
    w.close()


@pytest.mark.parametrize("nt,seg,shape,ws", [(32, 16, (72, 64), 0), (32, 64, (40, 136), 1), (32, 16, (104, 64), 2), (128, 32, (192, 96), 1),
                                             (128, 128, (256, 160), 2), (128, 48, (504, 64), 1)])
def test_two_lane_fused_body_emulated_matches_oracle(emul, nt, seg, shape, ws):
    """hg_fused_body2.cuh -- the code k_fused_ws2 runs: every thread advances column t of TWO adjacent strips in the
    two lanes of packed fp32 arithmetic (hg_v2.cuh, hg_cell2.cuh) -- executed thread by thread on the CPU (the packed
    operations as two scalar ones), single group and the warp-specialised split in either group order, against the
    oracle: every plane bit-exact, the deferred far cells exactly the cells whose back-trace leaves the +-1 window.
    Widths with a half-empty last strip pair, exactly two pairs, and one pair."""
    W, H = shape
    w = wet_world(H, 120, width=W, period=8)
    names = "rock dirt water fL fR fT fB sed_r sed_d".split()
    for _ in range(3):
        got, far = fused_emul_step(emul, w, nt, seg, ws, two_lane=True)
        pl = planes_of(w)
        far_ref = emul_step(emul, w, pl)
        w.step((w.steps + 1) * DT_TIME)
        want = planes_of(w)
        assert len(far) == far_ref
        mask = np.ones(W * H, bool); mask[far] = False
        for k, (g, x, name) in enumerate(zip(got, want, names)):
            if k >= 7:
                g = np.where(mask.reshape(H, W), g, x)
            assert_bit_equal(g, x, f"two-lane nt={nt} seg={seg} {name}")
    w.close()


def test_two_lane_thermal_marking_matches_oracle_steep(emul):
    """the packed thermal outflow (hg_thermal_outflow2: one lane hot, both hot, none) on roughened terrain with lowered
    talus angles, several parameter sets"""
    n = 96
    w = oracle.World(n, seed=SEED); w.gen_heightmap()
    H = w.get(0)
    rng = np.random.default_rng(3)
    H[..., 0] += rng.random((n, n), dtype=np.float32) * 6
    H[..., 1] += rng.random((n, n), dtype=np.float32) * 2
    H[..., 3] = H[..., 0] + H[..., 1] + H[..., 2]
    w.set(0, H)
    names = "rock dirt water fL fR fT fB sed_r sed_d".split()
    for ka in ((0.9, 0.3), (1.3, 0.6), (0.05, 1.55)):
        w.erosion.Kalpha[0], w.erosion.Kalpha[1] = ka
        for _ in range(2):
            got, far = fused_emul_step(emul, w, 32, 48, 1, two_lane=True)
            assert len(far) == 0
            w.dispatch_grid()
            for g, x, name in zip(got, planes_of(w), names):
                assert_bit_equal(g, x, f"two-lane Kalpha={ka}: {name}")
    w.close()


def test_table_form_simplex_equals_float_form(emul):
    """hg_simplex_tab (the rain kernel's gln_simplex with the five permutes as table lookups) returns the same bits as
    the float form for 2 M random points: map coordinates times the rain's noise scales, negative and large
    coordinates, and points on lattice lines."""
    rng = np.random.default_rng(7)
    n = 2_000_000
    xs = np.concatenate([rng.uniform(-70000, 70000, n // 2) * rng.choice([0.02, 0.04, 0.32, 2.56], n // 2),
                         rng.uniform(-300, 300, n // 4), np.round(rng.uniform(-600, 600, n // 4))]).astype(np.float32)
    ys = np.concatenate([rng.uniform(-70000, 70000, n // 2) * rng.choice([0.02, 0.04, 0.32, 2.56], n // 2),
                         rng.uniform(-300, 300, n // 4), np.round(rng.uniform(-600, 600, n // 4))]).astype(np.float32)
    emul.emul_simplex_tab_mismatches.restype = C.c_long
    bad = emul.emul_simplex_tab_mismatches(xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), C.c_long(len(xs)))
    assert bad == 0


def test_exact_rewrites_hold_on_a_sample(tmp_path):
    """The two exact rewrites of the second session, checked here on a sample (every 1009th / 211th bit pattern plus what the
    programs pin themselves) and exhaustively by the same programs without an argument (scripts/check_atan_one_division.c:
    0 mismatches over all 2^32 inputs; scripts/check_sqrt_threshold.c: 0 over all non-negative floats): hg_atanf's range
    reduction as one division, the momentum cut-off length(m) < 1e-12 as a comparison of the squared length, and the
    division by sqrt(2)f of the thermal outflow path as one multiply and two fmas (exact above 2.2e-32)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for src, arg, extra in (("check_atan_one_division.c", "1009", ["-fopenmp"]), ("check_sqrt_threshold.c", "211", []),
                            ("check_div_sqrt2.c", "499", ["-fopenmp"])):
        exe = str(tmp_path / src.replace(".c", ""))
        subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", *extra, os.path.join(root, "scripts", src), "-o", exe, "-lm"],
                       check=True, cwd=os.path.join(root, "scripts"))
        out = subprocess.run([exe, arg], capture_output=True, text=True, check=True).stdout
        assert " 0 mismatches" in out or "mismatches: 0" in out or "above 2e-32: 0" in out, out
