"""Pins oracle/deff_oracle.c (the CPU restatement) to the reference's own outputs recorded
in tests/golden/ by make_golden.py, and to the analytic known answers of the reference
documentation (doc 5.3).  CPU only."""
import numpy as np
import pytest

import _oracle as O
from golden.make_golden import kat_images

NCASES = 6


@pytest.mark.parametrize("k", range(NCASES))
@pytest.mark.parametrize("nphase", [2, 3])
def test_primitives_bit_exact(golden_prims, k, nphase):
    g = golden_prims
    img = g["img%d" % k]
    thr = 150 if nphase == 2 else 200
    Ds = 0.0 if nphase == 3 else 1e-3
    D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, 50.0)
    G, pf = O.floodfill((img > thr).astype(np.uint32))
    assert np.array_equal(G.astype(np.uint8), g["flood%d_p%d" % (k, nphase)])      # cuh:557-713
    assert pf == int(g["pathflag%d_p%d" % (k, nphase)])
    A, b = O.discretize(D, 0.25, 1.5, G if nphase == 3 else None)
    assert np.array_equal(A, g["A%d_p%d" % (k, nphase)])                            # cuh:715-902
    assert np.array_equal(b, g["b%d_p%d" % (k, nphase)])
    x0 = O.init_x(img.shape[1], img.shape[0], 0.25, 1.5)
    for maxit, pre in ((1, False), (37, False), (10001, False), (2500, True)):
        tag = "%d_p%d_it%d%s" % (k, nphase, maxit, "pre" if pre else "")
        r = O.jacobi(A, b, x0, D, 0.25, 1.5, 1e-7, maxit)
        assert r["iters"] == int(g["iters" + tag])
        # same arithmetic order, no FMA on either side: bit-identical (NaN-aware)
        assert np.array_equal(r["field"], g["x" + tag], equal_nan=True)             # cuh:69-92
        if not pre:     # JacobiGPUPreCond does not record deff (cuh:1144-1159)
            np.testing.assert_equal(r["deff_raw"], float(g["deff" + tag]))          # cuh:1252-1264


def test_bundled_00000_two_phase(golden_prims, golden_images):
    img = golden_images["00000"]
    r = O.solve_image(img, O.make_opts(Ds=1e-4, Df=1.0, nphase=2), O.MODE_2PH_BATCH, want_field=True)
    assert r["iters"] == [int(golden_prims["iters00000_2ph"])] == [100001]
    assert r["deff_raw"] == float(golden_prims["deff00000_2ph"])
    assert r["conv"] == float(golden_prims["conv00000_2ph"])
    assert np.array_equal(r["field"], golden_prims["x00000_2ph"])
    assert abs(r["deff"] - 0.1816910277372) < 1e-13


def _row(csv):
    return csv.strip().splitlines()[-1].split(",")


def test_driver_kats_match_reference_program(golden_drivers):
    """The oracle's driver restatement prints the same CSV fields as the reference program."""
    par, ser, wide, thin, p3 = kat_images()
    cases = [("kat_parallel_2ph_batch", par, dict(Ds=0.1, Df=1.0, nphase=2), O.MODE_2PH_BATCH, 0.37),
             ("kat_series_2ph_batch", ser, dict(Ds=0.1, Df=1.0, nphase=2), O.MODE_2PH_BATCH, 1 / (0.3 + 0.7 / 0.1)),
             ("kat_wide_2ph_batch", wide, dict(Ds=0.1, Df=1.0, nphase=2), O.MODE_2PH_BATCH, 1 / (0.5 + 0.5 / 0.1))]
    for name, img, kw, mode, analytic in cases:
        r = O.solve_image(img, O.make_opts(**kw), mode)
        row = _row(golden_drivers[name]["csv"])
        assert "%f" % r["porosity"] == row[1]
        assert "%d" % r["pathflag"] == row[2]
        assert "%f" % r["deff"] == row[3]
        assert "%f" % r["conv"] == row[6]
        assert abs(r["deff"] - analytic) / analytic < 1e-5      # doc 5.3 eq (7), (8)
        its = [l for l in golden_drivers[name]["stdout_key_lines"] if l.startswith("Iterations")]
        assert ["Iterations taken = %d" % i for i in r["iters"]] == its


@pytest.mark.slow
def test_driver_thin_phase_continuation(golden_drivers):
    par, ser, wide, thin, p3 = kat_images()
    r = O.solve_image(thin, O.make_opts(Ds=1.0, Df=1237500.0, nphase=2), O.MODE_2PH_SINGLE)
    g = golden_drivers["kat_thin_2ph_single"]
    its = [l for l in g["stdout_key_lines"] if l.startswith("Iterations")]
    assert ["Iterations taken = %d" % i for i in r["iters"]] == its
    assert r["iters"] == [70001, 100001, 110001, 70001]
    assert r["stage_D"] == [100.0, 10000.0, 1000000.0, 1237500.0]  # cuh:1762-1765
    row = _row(g["csv"])
    assert "%f" % r["deff"] == row[3] and "%f" % r["conv"] == row[6]
    assert abs(r["deff"] * 1237500.0 - 33.33246) < 2e-3            # doc 5.3.1: 33.33


def test_driver_three_phase(golden_drivers, golden_images):
    par, ser, wide, thin, p3 = kat_images()
    r = O.solve_image(p3, O.make_opts(), O.MODE_3PH)
    row = _row(golden_drivers["kat_parallel_3ph_single"]["csv"])
    assert "%f" % r["SVF"] == row[1] and "%f" % r["LVF"] == row[2] and "%d" % r["pathflag"] == row[3]
    assert "%1.3e" % r["deff"] == row[4] and "%1.3e" % r["conv"] == row[7]
    assert r["iters"] == [10001] * 7
    assert abs(r["deff"] - 371250.4) / 371250.4 < 1e-9             # doc 5.3.2 eq (9)
    assert abs(r["deff"] - 371250.399999912) < 1e-6


@pytest.mark.slow
def test_driver_three_phase_bundled(golden_drivers, golden_images):
    r = O.solve_image(golden_images["00000"], O.make_opts(), O.MODE_3PH)
    g = golden_drivers["bundled00000_3ph_single"]
    row = _row(g["csv"])
    assert "%f" % r["SVF"] == row[1] and "%f" % r["LVF"] == row[2] and "%d" % r["pathflag"] == row[3]
    assert "%1.3e" % r["deff"] == row[4] and "%1.3e" % r["conv"] == row[7]
    its = [l for l in g["stdout_key_lines"] if l.startswith("Iterations")]
    assert ["Iterations taken = %d" % i for i in r["iters"]] == its
    assert r["total_iters"] == 140007
    assert abs(r["deff"] - 224673.610442892) < 1e-6


def test_two_phase_Ds0_gives_nan(golden_drivers, golden_images):
    """Quirk Q13: Ds = 0 in 2-phase makes A0 = 0 in solids -> NaN Deff after one sweep."""
    r = O.solve_image(golden_images["00000"], O.make_opts(Ds=0.0, Df=1.0, nphase=2), O.MODE_2PH_BATCH)
    assert r["iters"] == [1] and np.isnan(r["deff"]) and np.isnan(r["conv"])
    assert _row(golden_drivers["bundled00000_2ph_Ds0_nan"]["csv"])[3] == "-nan"


def test_two_phase_single_small_Df_runs_no_stage(golden_drivers, golden_images):
    """Quirk Q8: SingleSim performs no solve when Df < 10 (cuh:1714, 1761)."""
    r = O.solve_image(golden_images["00000"], O.make_opts(Ds=1e-4, Df=1.0, nphase=2), O.MODE_2PH_SINGLE)
    assert r["nstages"] == 0 and r["total_iters"] == 0
    assert "Iterations taken" not in "".join(golden_drivers["bundled00000_2ph_single_Df1"]["stdout_key_lines"])


def test_weighted_harmonic_mean_zero():
    L = O.oracle()
    assert L.orc_weighted_harmonic_mean(0.5, 0.5, 0.0, 3.0) == 0.0     # cuh:358, w/0 = inf
    assert L.orc_weighted_harmonic_mean(0.5, 0.5, 0.0, 0.0) == 0.0
    assert L.orc_weighted_harmonic_mean(0.5, 0.5, 2.0, 2.0) == 2.0
