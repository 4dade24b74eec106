"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the golden vectors recorded from the reference, and -- at full config
sizes -- through size-independent properties.

Tolerances (north_star): |dDeff|/Deff <= 1e-4 with identical sweep counts; the field within the
reference's convergence criterion.  The device uses FMA contraction and a tree reduction for the
flux sums, the reference host code does not, so fields agree to ~1e-13 rather than bit-for-bit;
the tests assert the much tighter observed bounds next to the contractual ones.
"""
import os

import numpy as np
import pytest

import _oracle as O
import effectivediffusivityfvm_b200 as E
from golden.make_golden import kat_images

pytestmark = pytest.mark.gpu

DEFF_RTOL = 1e-4          # contractual (BASELINE.json north_star)
DEFF_RTOL_TIGHT = 1e-9    # observed headroom we do not want to lose silently
FIELD_ATOL = 1e-11


@pytest.fixture(scope="module")
def ctx():
    c = E.Deff2D(0)
    yield c
    c.close()


def blobs(seed, shape, levels=(0, 255), fracs=(0.6,), smooth=3):
    rng = np.random.default_rng(seed)
    z = rng.random(shape)
    for _ in range(smooth):
        z = (z + np.roll(z, 1, 0) + np.roll(z, -1, 0) + np.roll(z, 1, 1) + np.roll(z, -1, 1)) / 5
    qs = np.quantile(z, np.cumsum(fracs))
    out = np.full(shape, levels[-1], np.uint8)
    for lv, q in reversed(list(zip(levels[:-1], qs))):
        out[z < q] = lv
    return out


def rel(a, b):
    if np.isnan(a) and np.isnan(b):
        return 0.0
    return abs(a - b) / max(abs(b), 1e-300)


# ----------------------------------------------------------------------------- primitives

@pytest.mark.parametrize("nphase,Ds,Dg", [(2, 1e-3, 0.0), (3, 0.0, 50.0), (3, 0.02, 1237500.0)])
@pytest.mark.parametrize("shape,amp", [((12, 20), (1, 1)), ((17, 9), (1, 1)), ((33, 130), (1, 1)), ((16, 31), (2, 3)),
                                       ((40, 300), (1, 1)), ((7, 5), (4, 4))])
def test_assembly_sweeps_flux_vs_oracle(ctx, nphase, Ds, Dg, shape, amp):
    ampx, ampy = amp
    img = blobs(hash((shape, nphase)) % 1000, shape, levels=(0, 150, 255), fracs=(0.35, 0.35), smooth=1)
    CL, CR = 0.25, 1.5
    p = E.default_params(Ds=Ds, Df=1.0, Dg=Dg, amp_x=ampx, amp_y=ampy, CL=CL, CR=CR)
    ctx.set_kernel(1)
    ctx.domain_load(img, nphase, p)
    # oracle side
    D = O.fill_D(img, ampx, ampy, nphase, Ds, 1.0, Dg)
    G, pf = O.floodfill(O.grid_mask(img, ampx, ampy, 200 if nphase == 3 else 150))
    A, b = O.discretize(D, CL, CR, G if nphase == 3 else None)
    Ny, Nx = D.shape
    info = ctx.info()
    assert (info["Nx"], info["Ny"], info["pathflag"]) == (Nx, Ny, pf)
    # phase codes == the oracle's D classification, pinned == Grid in {1,2}
    codes = ctx.get_codes()
    Dtab = np.array([1.0, Ds, Dg])
    assert np.array_equal(Dtab[codes & 3], D)
    if nphase == 3:
        assert np.array_equal((codes & 4) != 0, (G == 1) | (G == 2))
    else:
        assert not np.any(codes & 4)
        assert info["porosity"] == O.oracle().orc_porosity(O._up(np.ascontiguousarray(img)), img.shape[1], img.shape[0])
    x0 = O.init_x(Nx, Ny, CL, CR)
    f0 = ctx.get_field()                                   # dead cells (A0 = 0, quirk Q13) read back as NaN
    assert np.array_equal(np.where(np.isnan(f0), x0, f0), x0)     # cuh:1732, bit-exact
    done = 0
    for n in (1, 2, 37):
        ctx.sweeps(n)
        done += n
        ref = O.sweeps(A, b, x0, done)
        got = ctx.get_field()
        assert np.array_equal(np.isnan(got), np.isnan(ref))          # dead cells (A0 = 0, quirk Q13)
        if not np.all(np.isnan(ref)):
            assert np.nanmax(np.abs(got - ref)) < 1e-13, (n, np.nanmax(np.abs(got - ref)))
        d_got, _ = ctx.flux()
        d_ref = O.flux_deff(ref, D, CL, CR)
        assert rel(d_got, d_ref) < 1e-12
    if nphase == 3:
        s, l = np.zeros(1), np.zeros(1)
        O.oracle().orc_fracts3(O._dp(D), Nx, Ny, Ds, 1.0, O._dp(s), O._dp(l))
        assert (info["SVF"], info["LVF"]) == (s[0], l[0])


@pytest.mark.parametrize("k", range(6))
@pytest.mark.parametrize("nphase", [2, 3])
def test_golden_primitives_from_reference(ctx, golden_prims, k, nphase):
    """Fields and Deff recorded from the reference's own JacobiGPU (kernel body on host threads)."""
    g = golden_prims
    img = g["img%d" % k]
    Ds = 0.0 if nphase == 3 else 1e-3
    p = E.default_params(Ds=Ds, Df=1.0, Dg=50.0, CL=0.25, CR=1.5)
    ctx.set_kernel(1)
    for maxit in (1, 37, 10001):
        ctx.domain_load(img, nphase, p)
        r = ctx.solve(1e-7, maxit)
        tag = "%d_p%d_it%d" % (k, nphase, maxit)
        assert r["iters"] == int(g["iters" + tag])
        ref = g["x" + tag]
        got = ctx.get_field()
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.nanmax(np.abs(got - ref)) < FIELD_ATOL
        dref = float(g["deff" + tag])
        if np.isnan(dref):
            assert np.isnan(r["deff_raw"])
        else:
            assert rel(r["deff_raw"], dref) < DEFF_RTOL_TIGHT


def test_residual_matches_reference_definition(ctx):
    img = blobs(3, (24, 40))
    p = E.default_params(Ds=1e-2, Df=1.0, CL=0.0, CR=1.0)
    ctx.set_kernel(1)
    ctx.domain_load(img, 2, p)
    ctx.sweeps(50)
    x = ctx.get_field()
    D = O.fill_D(img, 1, 1, 2, 1e-2, 1.0, 0.0)
    ref = O.oracle().orc_residual(24, 40, 0.0, 1.0, O._dp(x), O._dp(D))
    assert rel(ctx.residual(), ref) < 1e-10


@pytest.mark.parametrize("nphase", [2, 3])
def test_parity_with_reference_cuda_build(ctx, nphase):
    """The reference's own JacobiGPU + updateX_SOR, compiled unmodified by nvcc for sm_100a
    (oracle/_ref/libref_cuda.so), run on this GPU beside ours: same sweep count, Deff, field."""
    if O.reference("cuda") is None:
        pytest.skip("oracle/_ref/libref_cuda.so not built")
    img = blobs(40 + nphase, (72, 96), levels=(0, 150, 255), fracs=(0.3, 0.4))
    Ds, Dg, CL, CR = (0.0, 300.0, 0.0, 1.0) if nphase == 3 else (1e-3, 0.0, 0.0, 1.0)
    D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, Dg)
    G, _ = O.floodfill(O.grid_mask(img, 1, 1, 200 if nphase == 3 else 150))
    A, b = O.ref_discretize(D, CL, CR, G if nphase == 3 else None, kind="cuda")
    Ny, Nx = D.shape
    ref = O.ref_jacobi(A, b, O.init_x(Nx, Ny, CL, CR), D, CL, CR, 1e-5, 60001, kind="cuda")
    p = E.default_params(Ds=Ds, Df=1.0, Dg=Dg, CL=CL, CR=CR)
    for kernel in (1, 2):
        ctx.set_kernel(kernel, 4)
        ctx.domain_load(img, nphase, p)
        got = ctx.solve(1e-5, 60001)
        assert got["iters"] == ref["iters"]
        assert rel(got["deff_raw"], ref["deff_raw"]) < DEFF_RTOL_TIGHT
        f = ctx.get_field()
        assert np.array_equal(np.isnan(f), np.isnan(ref["field"]))
        assert np.nanmax(np.abs(f - ref["field"])) < FIELD_ATOL
    ctx.set_kernel(0)


def _maze(n, seed):
    """A tortuous open path: serpentine corridors with random gaps (many tile crossings)."""
    rng = np.random.default_rng(seed)
    img = np.full((n, n), 255, np.uint8)
    img[1::4, :] = 0                         # horizontal corridors
    for r in range(1, n - 4, 4):             # one connecting gap per corridor pair, alternating ends
        c = rng.integers(0, 8) if (r // 4) % 2 else n - 1 - rng.integers(0, 8)
        img[r:r + 5, c] = 0
    return img


@pytest.mark.parametrize("case", ["blobs", "quirk_solid_corner", "wall", "maze", "periodic", "amp", "percolation"])
def test_device_floodfill_equals_host_floodfill(ctx, case):
    """FloodFill by label propagation on the device (floodfill.cu) against the host FIFO flood,
    which is pinned to the reference's FloodFill (tests/test_host_logic.py): same PathFlag, same
    Grid (through the pinned bit of the 3-phase codes)."""
    amp = (1, 1)
    if case == "blobs":
        img = blobs(5, (300, 700), levels=(0, 150, 255), fracs=(0.3, 0.3))
    elif case == "quirk_solid_corner":
        img = blobs(6, (200, 333), levels=(0, 150, 255), fracs=(0.2, 0.3))
        img[0, 0] = 255
        img[:, 150:160] = 255                # a full-height wall: only the quirk's right-column seeds reach the right half
    elif case == "wall":
        img = blobs(7, (150, 400), levels=(0, 150, 255), fracs=(0.4, 0.3))
        img[0, 0] = 0
        img[:, 200:203] = 255
    elif case == "maze":
        img = _maze(400, 3)
    elif case == "periodic":
        img = np.full((130, 300), 255, np.uint8)
        img[:, 0:5] = 0
        img[0, :200] = 0                     # reachable only through the y-periodic wrap from the bottom row
        img[129, 150:290] = 0
        img[100:130, 289] = 0
        img[5, 199] = 0
    elif case == "amp":
        img = blobs(8, (90, 70), levels=(0, 150, 255), fracs=(0.3, 0.3))
        amp = (3, 4)
    else:
        img = np.where(np.random.default_rng(5).random((512, 512)) < 0.60, 0, 255).astype(np.uint8)
    for nphase in (3, 2):
        p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=50.0, amp_x=amp[0], amp_y=amp[1])
        res = {}
        for mode in (1, 2):
            ctx.set_floodfill(mode)
            ctx.domain_load(img, nphase, p)
            res[mode] = (ctx.info()["pathflag"], ctx.get_codes())
        ctx.set_floodfill(0)
        assert res[1][0] == res[2][0], (case, nphase)
        assert np.array_equal(res[1][1], res[2][1]), (case, nphase)
        # and both agree with the oracle's FloodFill
        G, pf = O.floodfill(O.grid_mask(img, amp[0], amp[1], 200 if nphase == 3 else 150))
        assert pf == res[2][0]
        if nphase == 3:
            assert np.array_equal((res[2][1] & 4) != 0, (G == 1) | (G == 2))


# ----------------------------------------------------------------------------- driver flows

def check_against_oracle(got, ref):
    assert got["iters"] == ref["iters"], (got["iters"], ref["iters"])
    assert got["pathflag"] == ref["pathflag"]
    for a, b in zip(got["stage_deff_raw"], ref["stage_deff_raw"]):
        assert rel(a, b) < DEFF_RTOL
        assert rel(a, b) < DEFF_RTOL_TIGHT
    if np.isnan(ref["deff"]):
        assert np.isnan(got["deff"]) and np.isnan(got["conv"])
    else:
        assert rel(got["deff"], ref["deff"]) <= DEFF_RTOL
        assert rel(got["deff"], ref["deff"]) < DEFF_RTOL_TIGHT
        assert abs(got["conv"] - ref["conv"]) < 1e-9
    assert got["porosity"] == ref["porosity"] and got["SVF"] == ref["SVF"] and got["LVF"] == ref["LVF"]


def test_kat_documented_cases(ctx, golden_drivers):
    """doc 5.3 known answers: same sweep counts and CSV fields as the reference program."""
    par, ser, wide, thin, p3 = kat_images()
    for name, img, analytic in (("kat_parallel_2ph_batch", par, 0.37), ("kat_series_2ph_batch", ser, 1 / (0.3 + 7.0)),
                                ("kat_wide_2ph_batch", wide, 1 / (0.5 + 5.0))):
        got = ctx.solve_image(img, E.default_params(Ds=0.1, Df=1.0, mode=E.MODE_2PH_BATCH))
        ref = O.solve_image(img, O.make_opts(Ds=0.1, Df=1.0, nphase=2), O.MODE_2PH_BATCH)
        check_against_oracle(got, ref)
        row = golden_drivers[name]["csv"].strip().splitlines()[-1].split(",")
        assert "%f" % got["porosity"] == row[1] and "%d" % got["pathflag"] == row[2] and "%f" % got["deff"] == row[3]
        assert "%f" % got["conv"] == row[6]
        assert rel(got["deff"], analytic) < 1e-5
    got = ctx.solve_image(p3, E.default_params())
    ref = O.solve_image(p3, O.make_opts(), O.MODE_3PH)
    check_against_oracle(got, ref)
    row = golden_drivers["kat_parallel_3ph_single"]["csv"].strip().splitlines()[-1].split(",")
    assert "%f" % got["SVF"] == row[1] and "%f" % got["LVF"] == row[2] and "%1.3e" % got["deff"] == row[4]
    assert rel(got["deff"], 371250.4) < 1e-9               # doc 5.3.2


def test_thin_phase_continuation(ctx):
    par, ser, wide, thin, p3 = kat_images()
    got = ctx.solve_image(thin, E.default_params(Ds=1.0, Df=1237500.0, mode=E.MODE_2PH_SINGLE))
    assert got["iters"] == [70001, 100001, 110001, 70001]
    assert got["stage_D"] == [100.0, 10000.0, 1000000.0, 1237500.0]
    for a, b in zip(got["stage_deff_raw"], [0.251889395650745 * 100, 0.00332259655936631 * 1e4,
                                            3.33322707579901e-05 * 1e6, 2.69353560643523e-05 * 1237500]):
        # contrast 1e6 over 110 001 sweeps: the device folds w/A0 into the face weights (one
        # rounding more per coefficient than cuh:89) and contracts to FMA; observed 1.6e-8
        assert rel(a, b) < 1e-6
    assert abs(got["deff"] * 1237500.0 - 33.33246) < 2e-3           # doc 5.3.1


def test_bundled_00000(ctx, golden_images, golden_prims, golden_drivers):
    img = golden_images["00000"]
    got = ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH), want_field=True)
    assert got["iters"] == [100001]
    assert rel(got["deff"], float(golden_prims["deff00000_2ph"])) < DEFF_RTOL_TIGHT
    assert rel(got["deff"], 0.1816910277372) < 1e-10
    assert np.max(np.abs(got["field"] - golden_prims["x00000_2ph"])) < FIELD_ATOL
    got3 = ctx.solve_image(img, E.default_params())                 # shipped input.txt defaults, config 1
    assert got3["iters"] == [80001] + [10001] * 6 and got3["total_iters"] == 140007
    assert rel(got3["deff"], 224673.610442892) < 1e-9
    row = golden_drivers["bundled00000_3ph_single"]["csv"].strip().splitlines()[-1].split(",")
    assert ["%f" % got3["SVF"], "%f" % got3["LVF"], "%d" % got3["pathflag"], "%1.3e" % got3["deff"]] == row[1:5]
    assert "%1.3e" % got3["conv"] == row[7]


def test_quirks(ctx, golden_images):
    img = golden_images["00000"]
    # Q13: 2-phase with Ds = 0 -> NaN Deff after one sweep, NaN in the solid cells of the map
    got = ctx.solve_image(img, E.default_params(Ds=0.0, Df=1.0, mode=E.MODE_2PH_BATCH), want_field=True)
    assert got["iters"] == [1] and np.isnan(got["deff"]) and np.isnan(got["conv"])
    ref = O.solve_image(img, O.make_opts(Ds=0.0, Df=1.0, nphase=2), O.MODE_2PH_BATCH, want_field=True)
    assert np.array_equal(np.isnan(got["field"]), np.isnan(ref["field"]))
    assert np.nanmax(np.abs(got["field"] - ref["field"])) < 1e-13
    # Q8: 2-phase single with Df < 10 performs no solve
    got = ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_SINGLE))
    assert got["nstages"] == 0 and got["total_iters"] == 0
    # MaxIter exit: Deff from the last check, field from the last sweep (Q4)
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=12345, tol=1e-12)
    got = ctx.solve_image(img, p, want_field=True)
    ref = O.solve_image(img, O.make_opts(Ds=1e-3, Df=1.0, nphase=2, max_iter=12345, tol=1e-12), O.MODE_2PH_BATCH, want_field=True)
    check_against_oracle(got, ref)
    assert got["iters"] == [12345]
    assert np.max(np.abs(got["field"] - ref["field"])) < FIELD_ATOL


def test_strict_reference_off_defines_the_quirky_cases(ctx, golden_images):
    """strict_reference = 0: Q8 one stage at Df instead of none, Q11 no right-column seeding,
    Q12 pixel == 150 solid in the FloodFill mask; everything else unchanged."""
    img = golden_images["00000"]
    p = E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_SINGLE, strict_reference=0)
    got = ctx.solve_image(img, p)
    ref = ctx.solve_image(img, E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH))
    assert got["nstages"] == 1 and got["iters"] == ref["iters"] and got["deff"] == ref["deff"]
    # a full-height solid wall with a solid corner pixel: the reference reports PathFlag 1 (Q11)
    wall = np.zeros((64, 96), np.uint8)
    wall[:, 40:44] = 255
    wall[0, 0] = 255
    for mode in (1, 2):
        ctx.set_floodfill(mode)
        ctx.domain_load(wall, 2, E.default_params(Ds=1e-3, Df=1.0))
        assert ctx.info()["pathflag"] == 1
        ctx.domain_load(wall, 2, E.default_params(Ds=1e-3, Df=1.0, strict_reference=0))
        assert ctx.info()["pathflag"] == 0
    # pixel == 150: open for the reference's FloodFill (> 150), solid in D (< 150 is fluid)
    grey = np.zeros((32, 48), np.uint8)
    grey[:, 20:24] = 150
    for mode in (1, 2):
        ctx.set_floodfill(mode)
        ctx.domain_load(grey, 2, E.default_params(Ds=1e-3, Df=1.0))
        assert ctx.info()["pathflag"] == 1
        ctx.domain_load(grey, 2, E.default_params(Ds=1e-3, Df=1.0, strict_reference=0))
        assert ctx.info()["pathflag"] == 0
    ctx.set_floodfill(0)


def test_mesh_amplification_and_custom_cadence(ctx):
    img = blobs(21, (20, 28), levels=(0, 150, 255), fracs=(0.3, 0.4))
    for mode, omode, kw in ((E.MODE_2PH_BATCH, O.MODE_2PH_BATCH, dict(Ds=1e-2, Df=1.0)),
                            (E.MODE_3PH, O.MODE_3PH, dict(Ds=0.0, Df=1.0, Dg=200.0))):
        p = E.default_params(mode=mode, amp_x=3, amp_y=2, check_every=500, max_iter=20000, **kw)
        got = ctx.solve_image(img, p, want_field=True)
        ref = O.solve_image(img, O.make_opts(ampx=3, ampy=2, check_every=500, max_iter=20000,
                                             nphase=3 if mode == E.MODE_3PH else 2, **kw), omode, want_field=True)
        check_against_oracle(got, ref)
        assert np.nanmax(np.abs(got["field"] - ref["field"])) < FIELD_ATOL


def test_batch_equals_serial(ctx):
    imgs = np.stack([blobs(100 + k, (32, 32)) for k in range(5)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=1000, max_iter=50000)
    got = ctx.solve_batch(imgs, p)
    for k in range(5):
        ref = O.solve_image(imgs[k], O.make_opts(Ds=1e-3, Df=1.0, nphase=2, check_every=1000, max_iter=50000), O.MODE_2PH_BATCH)
        check_against_oracle(got[k], ref)


def _same_result(a, b):
    """Packed and single-image solves use the same kernels and reduction order: bit-identical."""
    for k in ("iters", "nstages", "pathflag", "porosity", "SVF", "LVF", "total_iters", "n_cells", "stage_D"):
        assert a[k] == b[k], (k, a[k], b[k])
    for k in ("deff", "deff_raw", "conv"):
        assert a[k] == b[k] or (np.isnan(a[k]) and np.isnan(b[k])), (k, a[k], b[k])
    assert np.array_equal(np.array(a["stage_deff_raw"]), np.array(b["stage_deff_raw"]), equal_nan=True)


@pytest.mark.parametrize("slots", [0, 3])
def test_packed_batch_2phase_matches_single_image_and_oracle(ctx, slots):
    """K5: images of different porosity (different sweep counts) packed into one stack; with
    slots=3 finished images are replaced from the queue while the others keep iterating."""
    imgs = np.stack([blobs(200 + k, (40, 56), fracs=(0.45 + 0.04 * k,)) for k in range(9)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=500, max_iter=9000, tol=1e-3)
    ctx.set_batch_slots(slots)
    try:
        got = ctx.solve_batch(imgs, p, want_fields=True)
    finally:
        ctx.set_batch_slots(0)
    assert len({tuple(g["iters"]) for g in got}) > 1          # the images really stop at different checks
    for k in range(len(imgs)):
        one = ctx.solve_image(imgs[k], p, want_field=True)
        _same_result(got[k], one)
        assert np.array_equal(got[k]["field"], one["field"], equal_nan=True)
        ref = O.solve_image(imgs[k], O.make_opts(Ds=1e-3, Df=1.0, nphase=2, check_every=500, max_iter=9000, tol=1e-3), O.MODE_2PH_BATCH)
        check_against_oracle(got[k], ref)


@pytest.mark.parametrize("slots", [0, 2])
def test_packed_batch_3phase_stages_match_single_image(ctx, slots):
    """Every packed image walks its own pre-conditioning stages (cuh:1492-1597): images in
    different stages share the launches, the stage lives in the cell code."""
    imgs = np.stack([blobs(300 + k, (36, 44), levels=(0, 150, 255), fracs=(0.25 + 0.03 * k, 0.4)) for k in range(6)])
    p = E.default_params(Ds=0.0, Df=1.0, Dg=300.0, mode=E.MODE_3PH, check_every=300, max_iter=20000, amp_x=2, amp_y=1, tol=1e-4)
    ctx.set_batch_slots(slots)
    try:
        got = ctx.solve_batch(imgs, p, want_fields=True)
    finally:
        ctx.set_batch_slots(0)
    for k in range(len(imgs)):
        assert got[k]["nstages"] == 3 and got[k]["stage_D"] == [10.0, 100.0, 300.0]
        one = ctx.solve_image(imgs[k], p, want_field=True)
        _same_result(got[k], one)
        assert np.array_equal(got[k]["field"], one["field"], equal_nan=True)
    ref = O.solve_image(imgs[2], O.make_opts(Ds=0.0, Df=1.0, Dg=300.0, check_every=300, max_iter=20000, ampx=2, ampy=1, tol=1e-4), O.MODE_3PH)
    check_against_oracle(got[2], ref)


def test_packed_batch_quirks(ctx, golden_images):
    """Q13 (Ds = 0 in 2-phase: NaN after the first check) and MaxIter exits inside a packed batch."""
    img = golden_images["00000"]
    imgs = np.stack([img, img[::-1].copy(), img[:, ::-1].copy()])
    got = ctx.solve_batch(imgs, E.default_params(Ds=0.0, Df=1.0, mode=E.MODE_2PH_BATCH))
    for g in got:
        assert g["iters"] == [1] and np.isnan(g["deff"]) and np.isnan(g["conv"])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=12345, tol=1e-12)
    got = ctx.solve_batch(imgs, p, want_fields=True)
    for k in range(3):
        one = ctx.solve_image(imgs[k], p, want_field=True)
        assert got[k]["iters"] == [12345]
        _same_result(got[k], one)
        assert np.array_equal(got[k]["field"], one["field"], equal_nan=True)


def test_drop_in_program(ctx, tmp_path, golden_drivers):
    """input.txt in, CSV + CMAP out, through deff2d_run_input_file; compared with the files the
    reference program wrote for the same inputs (Time column excluded)."""
    par, ser, wide, thin, p3 = kat_images()
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with open("p3.jpg", "wb") as f:        # PGM content under a .jpg name, like the golden run
            f.write(b"P5\n100 100\n255\n" + p3.tobytes())
        g = golden_drivers["kat_parallel_3ph_single"]
        keys = ["Phases", "Ds", "Df", "Dg", "MeshAmpX", "MeshAmpY", "InputName", "CR", "CL", "OutputName", "printCMap",
                "CMapName", "Convergence", "MaxIter", "Verbose", "RunBatch", "NumImages"]
        base = dict(Phases=3, Ds=0, Df=1, Dg=1237500, MeshAmpX=1, MeshAmpY=1, InputName="p3.jpg", CR=1, CL=0,
                    OutputName="out.csv", printCMap=1, CMapName="CMAP.csv", Convergence="1e-5", MaxIter="5e5",
                    Verbose=0, RunBatch=0, NumImages=1)
        with open("input.txt", "w") as f:
            f.write("Input File:\n" + "\n".join("%s: %s" % (k, base[k]) for k in keys) + "\n")
        ctx.run_input_file("input.txt")
        mine = open("out.csv").read().strip().splitlines()
        ref = g["csv"].strip().splitlines()
        assert mine[0] == ref[0]
        a, b = mine[1].split(","), ref[1].split(",")
        assert a[:5] == b[:5] and a[6:7] == b[6:7] and a[8:] == b[8:]       # all but Time and conv
        assert abs(float(a[7]) - float(b[7])) < 1e-9
        cm = open("CMAP.csv").read().splitlines()
        assert cm[:6] == g["cmap_head"] and len(cm) == g["cmap_lines"] and cm[-3:] == g["cmap_tail"]
        # 3-phase batch of two images
        with open("00000.jpg", "wb") as f:
            f.write(b"P5\n100 100\n255\n" + p3.tobytes())
        with open("00001.jpg", "wb") as f:
            f.write(b"P5\n100 100\n255\n" + par.tobytes())
        base.update(RunBatch=1, NumImages=2, printCMap=0, OutputName="batch.csv")
        with open("input.txt", "w") as f:
            f.write("Input File:\n" + "\n".join("%s: %s" % (k, base[k]) for k in keys) + "\n")
        ctx.run_input_file("input.txt")
        mine = open("batch.csv").read().strip().splitlines()
        ref = golden_drivers["kat_parallel_3ph_batch"]["csv"].strip().splitlines()
        assert mine[0] == ref[0] and len(mine) == len(ref) == 3
        for ma, rb in zip(mine[1:], ref[1:]):
            a, b = ma.split(","), rb.split(",")
            assert a[:5] == b[:5] and a[6:7] == b[6:7] and a[8:] == b[8:]
    finally:
        os.chdir(cwd)


def test_drop_in_executable_on_bundled_jpeg(tmp_path, golden_dir, golden_drivers):
    """The `deff2d` executable in place of the reference's a.out: shipped input.txt (InputName
    changed to the bundled 00000.jpg, as in BASELINE config 1), JPEG decoded by the library's own
    decoder, CSV compared with the row the reference program wrote (Time excluded)."""
    import subprocess
    exe = os.path.join(os.path.dirname(E.LIB_PATH), "deff2d")
    assert os.path.exists(exe), "drop-in executable not built"
    z = np.load(os.path.join(golden_dir, "jpeg.npz"))
    (tmp_path / "00000.jpg").write_bytes(z["file_bundled_00000"].tobytes())
    (tmp_path / "input.txt").write_text(
        "Input Parameters:\nPhases: 3\nDs: 0\nDf: 1\nDg: 1237500\nMeshAmpX: 1\nMeshAmpY: 1\nInputName: 00000.jpg\n"
        "CR: 1\nCL: 0\nOutputName: TestOut.csv\nprintCMap: 1\nCMapName: CMAP_00000.csv\nConvergence: 1e-5\n"
        "MaxIter: 5e5\nVerbose: 1\nRunBatch: 0\nNumImages: 1\n")
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("Iterations taken = ") == 7 and "Iterations taken = 80001" in out.stdout
    mine = (tmp_path / "TestOut.csv").read_text().strip().splitlines()
    ref = golden_drivers["bundled00000_3ph_single"]["csv"].strip().splitlines()
    assert mine[0] == ref[0]
    a, b = mine[1].split(","), ref[1].split(",")
    assert a[:5] == b[:5] and a[6:] == b[6:]                     # everything but Time
    cm = (tmp_path / "CMAP_00000.csv").read_text().splitlines()
    assert cm[0] == "X,Y,C" and len(cm) == 128 * 128 + 1


def test_drop_in_executable_packed_batch(tmp_path):
    """RunBatch: 1 through the `deff2d` executable: equally sized images are solved packed (and,
    with the `Devices:` extension key, split over the GPUs present); the CSV rows equal the
    oracle's per-image results in the reference's %f format."""
    import subprocess
    exe = os.path.join(os.path.dirname(E.LIB_PATH), "deff2d")
    imgs = [blobs(400 + k, (40, 56), fracs=(0.5 + 0.05 * k,)) for k in range(5)]
    for k, im in enumerate(imgs):
        (tmp_path / ("%05d.jpg" % k)).write_bytes(b"P5\n56 40\n255\n" + im.tobytes())      # PGM content, .jpg name
    (tmp_path / "input.txt").write_text(
        "Input Parameters:\nPhases: 2\nDs: 0.01\nDf: 1\nDg: 0\nMeshAmpX: 1\nMeshAmpY: 1\nInputName: unused.jpg\n"
        "CR: 1\nCL: 0\nOutputName: batch.csv\nprintCMap: 0\nCMapName: unused.csv\nConvergence: 1e-4\n"
        "MaxIter: 100000\nVerbose: 0\nRunBatch: 1\nNumImages: 5\nDevices: 2\n")
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rows = (tmp_path / "batch.csv").read_text().strip().splitlines()
    assert rows[0] == "imgNum,porosity,PathFlag,Deff,Time,nElements,converge,ds,df" and len(rows) == 6
    for k, im in enumerate(imgs):
        ref = O.solve_image(im, O.make_opts(Ds=0.01, Df=1.0, nphase=2, tol=1e-4, max_iter=100000), O.MODE_2PH_BATCH)
        f = rows[1 + k].split(",")
        assert f[0] == str(k) and f[1] == "%f" % ref["porosity"] and f[2] == "%d" % ref["pathflag"]
        assert f[3] == "%f" % ref["deff"] and f[5] == "2240" and f[6] == "%f" % ref["conv"]
        assert f[7] == "%f" % 0.01 and f[8] == "%f" % 1.0


def test_drop_in_executable_batch_per_image_path_streams_rows(tmp_path):
    """RunBatch: 1 with Verbose: 1 (the reference's per-image stdout order is kept, so images are
    solved one at a time) and images of two sizes: rows are appended as each image finishes and the
    finished CSV equals what the packed path writes for the same inputs."""
    import subprocess
    exe = os.path.join(os.path.dirname(E.LIB_PATH), "deff2d")
    imgs = [blobs(500 + k, (40, 56) if k != 2 else (48, 40), fracs=(0.55,)) for k in range(4)]
    for k, im in enumerate(imgs):
        (tmp_path / ("%05d.jpg" % k)).write_bytes(b"P5\n%d %d\n255\n" % (im.shape[1], im.shape[0]) + im.tobytes())
    (tmp_path / "input.txt").write_text(
        "Phases: 2\nDs: 0.01\nDf: 1\nDg: 0\nMeshAmpX: 1\nMeshAmpY: 1\nInputName: unused.jpg\nCR: 1\nCL: 0\n"
        "OutputName: rows.csv\nprintCMap: 0\nCMapName: unused.csv\nConvergence: 1e-4\nMaxIter: 100000\nVerbose: 1\n"
        "RunBatch: 1\nNumImages: 4\n")
    out = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("Iterations taken = ") == 4
    rows = (tmp_path / "rows.csv").read_text().strip().splitlines()
    assert rows[0].startswith("imgNum,porosity") and len(rows) == 5
    for k, im in enumerate(imgs):
        ref = O.solve_image(im, O.make_opts(Ds=0.01, Df=1.0, nphase=2, tol=1e-4, max_iter=100000), O.MODE_2PH_BATCH)
        f = rows[1 + k].split(",")
        assert f[0] == str(k) and f[1] == "%f" % ref["porosity"] and f[3] == "%f" % ref["deff"] and f[5] == str(im.size)


# ----------------------------------------------------------------------------- full-size properties

def test_full_size_properties_config2(ctx, golden_images):
    """BASELINE config 2 (00042.jpg x4 = 4008 x 8028 cells): properties that need no oracle run."""
    img = golden_images["00042"]
    p = E.default_params(amp_x=4, amp_y=4)
    ctx.set_kernel(0)
    ctx.domain_load(img, 3, p)
    info = ctx.info()
    assert (info["Nx"], info["Ny"]) == (4008, 8028)
    # amplification keeps phase fractions: SVF/LVF equal the native-resolution values to rounding
    c1 = E.Deff2D(0)
    try:
        c1.domain_load(img, 3, E.default_params())
        i1 = c1.info()
        assert abs(info["SVF"] - i1["SVF"]) < 1e-9 and abs(info["LVF"] - i1["LVF"]) < 1e-9
        assert info["pathflag"] == i1["pathflag"]
    finally:
        c1.close()
    ctx.sweeps(40)
    f = ctx.get_field()
    assert np.all(np.isfinite(f))
    assert f.min() >= -1e-12 and f.max() <= 1.0 + 1e-12            # discrete maximum principle (weights >= 0, sum <= 1)
    codes = ctx.get_codes()
    assert np.all(f[(codes & 4) != 0] <= (1.0 / 3.0) ** 40 * (1 + 1e-10))   # pinned cells decay as x/3 per sweep (Q14)
    # the sweep is linear: sweeps(a*x) == a*sweeps(x) on pinned-free interior, checked through Deff
    d40, _ = ctx.flux()
    ctx.set_field(0.5 * f)
    ctx.sweeps(2)
    fa = ctx.get_field()
    ctx.set_field(f)
    ctx.sweeps(2)
    fb = ctx.get_field()
    # x -> 0.5 x is not a symmetry of the affine map (CR = 1 enters through the right face), but
    # the difference of two iterates is: S(f) - S(0.5 f) == 0.5 * (S(f) - S(0))
    ctx.set_field(np.zeros_like(f))
    ctx.sweeps(2)
    f0 = ctx.get_field()
    assert np.max(np.abs((fb - fa) - 0.5 * (fb - f0))) < 1e-13
    assert np.isfinite(d40) and d40 > 0


def test_full_size_config5_percolation_vs_oracle(ctx):
    """BASELINE config 5 (2048 x 2048 site percolation at p = 0.60, Ds/Df = 1e-4): every patch
    crosses phase interfaces.  The first 120 sweeps against the oracle on the full domain (the
    full 5e5-sweep solve is a bench case, not a test), tiled == streaming bit for bit."""
    from effectivediffusivityfvm_b200.datasets import c5_image
    img = c5_image()
    p = E.default_params(Ds=1e-4, Df=1.0, mode=E.MODE_2PH_BATCH)
    D = O.fill_D(img, 1, 1, 2, 1e-4, 1.0, 0.0)
    A, b = O.discretize(D, 0.0, 1.0, None)
    ref = O.sweeps(A, b, O.init_x(2048, 2048, 0.0, 1.0), 120)
    fields = {}
    for kernel, T in ((1, 1), (2, 4), (0, 0)):
        ctx.set_kernel(kernel, T)
        ctx.domain_load(img, 2, p)
        ctx.sweeps(120)
        fields[kernel] = ctx.get_field()
        assert np.max(np.abs(fields[kernel] - ref)) < 1e-13
        assert rel(ctx.flux()[0], O.flux_deff(ref, D, 0.0, 1.0)) < 1e-12
    assert np.array_equal(fields[1], fields[2]) and np.array_equal(fields[1], fields[0])
    _, pf = O.floodfill(O.grid_mask(img, 1, 1, 150))
    assert ctx.info()["pathflag"] == pf
    ctx.set_kernel(0)


def test_full_size_config4_properties(ctx):
    """BASELINE config 4 (one 16384 x 16384 two-phase domain, 268 M cells): no oracle run fits a
    test, so size-independent properties -- the tiled and the streaming kernel give bit-identical
    boundary flux and residual after the same sweeps, the maximum principle holds on the boundary
    columns, Deff decreases from its x0 value monotonically over the first checks."""
    from effectivediffusivityfvm_b200.datasets import c4_image
    img = np.tile(c4_image(4096), (4, 4))                       # the generator is periodic: a seamless 16384^2 medium in seconds
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH)
    out = {}
    for kernel, T in ((1, 1), (0, 0)):
        ctx.set_kernel(kernel, T)
        ctx.domain_load(img, 2, p)
        info = ctx.info()
        assert (info["Nx"], info["Ny"]) == (16384, 16384)
        vals = []
        for n in (1, 12, 27):
            ctx.sweeps(n)
            vals.append((ctx.flux(), ctx.residual()))
        out[kernel] = vals
    for a, b in zip(out[1], out[0]):
        assert a[0] == b[0]                                     # Deff, Q1, Q2: bit-identical from both kernels
        assert rel(a[1], b[1]) < 1e-12                          # the residual sum uses atomics: order varies
    d = [v[0][0] for v in out[0]]
    assert all(np.isfinite(d)) and d[0] > d[1] > d[2] > 0
    assert abs(info["porosity"] - 0.6) < 1e-3
    ctx.set_kernel(0)
    ctx.domain_load(blobs(1, (16, 16)), 2, p)                   # release nothing, but leave a small domain resident



# ----------------------------------------------------------------------------- config sizes against the oracle

def test_config2_full_size_20_sweeps_vs_oracle(ctx, golden_images):
    """BASELINE config 2 at full size (4008 x 8028 cells, 3-phase shipped defaults): codes, pinned mask and the field
    after 20 sweeps against the CPU oracle on the whole domain."""
    img = golden_images["00042"]
    p = E.default_params(amp_x=4, amp_y=4)
    D = O.fill_D(img, 4, 4, 3, 0.0, 1.0, 1237500.0)
    G, pf = O.floodfill(O.grid_mask(img, 4, 4, 200))
    A, b = O.discretize(D, 0.0, 1.0, G)
    ref = O.sweeps(A, b, O.init_x(4008, 8028, 0.0, 1.0), 20)
    del A, b
    ctx.set_kernel(0)
    ctx.domain_load(img, 3, p)
    assert ctx.info()["pathflag"] == pf
    codes = ctx.get_codes()
    assert np.array_equal((codes & 4) != 0, (G == 1) | (G == 2))
    ctx.sweeps(20)
    f = ctx.get_field()
    assert np.array_equal(np.isnan(f), np.isnan(ref))
    assert np.nanmax(np.abs(f - ref)) < 1e-12
    assert rel(ctx.flux()[0], O.flux_deff(ref, D, 0.0, 1.0)) < 1e-11


def test_packed_batch_of_config3_images_vs_oracle(ctx):
    """Eight BASELINE config 3 images (256 x 256, sigma 3 px, Ds/Df = 1e-3) through the packed batch mode against the
    CPU oracle image by image: sweep counts, PathFlag, porosity, Deff; MaxIter 30 001 bounds the oracle's time."""
    from effectivediffusivityfvm_b200.datasets import c3_image
    imgs = np.stack([c3_image(k) for k in range(8)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=1e-5, max_iter=30001)
    got = ctx.solve_batch(imgs, p, want_fields=True)
    for k in range(8):
        ref = O.solve_image(imgs[k], O.make_opts(Ds=1e-3, Df=1.0, nphase=2, max_iter=30001), O.MODE_2PH_BATCH, want_field=(k == 3))
        assert got[k]["iters"] == ref["iters"]
        assert got[k]["pathflag"] == ref["pathflag"] and got[k]["porosity"] == ref["porosity"]
        assert rel(got[k]["deff"], ref["deff"]) < DEFF_RTOL_TIGHT
        if k == 3:
            assert np.max(np.abs(got[k]["field"] - ref["field"])) < FIELD_ATOL


def test_stop_rule_at_the_tolerance_boundary(ctx):
    """cuh:1232 continues while `tol < fabs(change)`: a tolerance equal to |change| of a check stops there, the next
    representable smaller tolerance does not.  Against the oracle's loop with |change| within 1e-9 (relative) of tol on
    either side -- the two implementations' Deff agree to ~1e-13, so they must take the same decision there."""
    img = blobs(31, (64, 96))
    opts = dict(Ds=1e-2, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=500, max_iter=100000)
    ctx.domain_load(img, 2, E.default_params(**opts))
    probe = ctx.solve(1e-30, 6001)                       # Deff at sweeps 1, 501, ..., 6001
    tr = probe["trace"]
    k = 6                                                # the check after sweep 3001
    change = abs((tr[k - 1] - tr[k]) / tr[k - 1])
    stop_at = k * 500 + 1
    assert ctx.solve_image(img, E.default_params(tol=change, **opts))["iters"] == [stop_at]                  # tol == |change|: stop
    assert ctx.solve_image(img, E.default_params(tol=float(np.nextafter(change, 0.0)), **opts))["iters"][0] > stop_at
    for tol, stops in ((change * (1 + 1e-9), True), (change * (1 - 1e-9), False)):
        got = ctx.solve_image(img, E.default_params(tol=tol, **opts))
        ref = O.solve_image(img, O.make_opts(Ds=1e-2, Df=1.0, nphase=2, check_every=500, max_iter=100000, tol=tol), O.MODE_2PH_BATCH)
        assert got["iters"] == ref["iters"]
        assert (got["iters"] == [stop_at]) == stops
        assert rel(got["deff"], ref["deff"]) < DEFF_RTOL_TIGHT


def test_packed_batch_stream_late_images_and_early_results(ctx):
    """deff2d_solve_batch_stream: images that are not ready when a slot frees up are asked for again later (the solve
    goes on with what is resident), the stream may end before `count`, results arrive as images finish -- and equal
    the array call's bit for bit."""
    from effectivediffusivityfvm_b200.datasets import c3_image
    imgs = np.stack([c3_image(300 + k, 96) for k in range(9)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=40001, check_every=2000)
    ctx.set_batch_slots(4)
    want = ctx.solve_batch(imgs, p, want_fields=True)
    calls = {"n": 0, "refused": 0}
    got, order = {}, []

    def fetch(k, wait):
        calls["n"] += 1
        if k >= 9:
            return False                                  # the stream ends before the announced count of 20
        if not wait and k >= 2 and calls["n"] % 3 != 0:
            calls["refused"] += 1
            return None                                   # "still decoding"
        return imgs[k]

    def done(k, res, field):
        got[k] = (res, field)
        order.append(k)

    solved = ctx.solve_batch_stream(20, (96, 96), p, fetch, done, want_fields=True)
    ctx.set_batch_slots(0)
    assert solved == 9 and sorted(got) == list(range(9)) and calls["refused"] > 0
    for k in range(9):
        a, b = want[k], got[k][0]
        assert a["iters"] == b["iters"] and a["deff"] == b["deff"] and a["pathflag"] == b["pathflag"] and a["porosity"] == b["porosity"]
        assert np.array_equal(want[k]["field"], got[k][1])


# ----------------------------------------------------------------------------- K5 (cluster-resident) == K3 (streaming)

@pytest.mark.parametrize("shape", [(64, 64), (24, 16), (100, 130), (65, 64), (64, 65), (200, 70), (129, 255), (256, 256)])
@pytest.mark.parametrize("nphase", [2, 3])
def test_resident_kernel_is_bitwise_identical_to_streaming(ctx, shape, nphase):
    """A domain of up to 256 x 256 cells stays on chip for all sweeps of a call (one launch, a cluster of up to 4 x 4
    CTAs exchanging edges through distributed shared memory): same arithmetic as K2 / K3, so the same bits."""
    img = blobs(shape[0] * 7 + nphase, shape, levels=(0, 150, 255), fracs=(0.3, 0.4), smooth=2)
    p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=80.0, CL=0.25, CR=1.5)
    for n in (1, 2, 3, 29, 400):
        ctx.set_kernel(1)
        ctx.domain_load(img, nphase, p)
        ctx.sweeps(n)
        ref, dref = ctx.get_field(), ctx.flux()[0]
        ctx.set_kernel(0)
        ctx.set_resident(0)
        ctx.domain_load(img, nphase, p)
        l0 = ctx.kernel_launches
        ctx.sweeps(n)
        assert ctx.kernel_launches - l0 == 1
        assert np.array_equal(ctx.get_field(), ref, equal_nan=True)
        d = ctx.flux()[0]
        assert d == dref or (np.isnan(d) and np.isnan(dref))
    ctx.set_kernel(0)


def test_resident_packed_batch_matches_tiled_packed_batch(ctx):
    from effectivediffusivityfvm_b200.datasets import c3_image
    imgs = np.stack([c3_image(100 + k, 192) for k in range(12)])
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=40001)
    out = {}
    for mode in (1, 2):
        ctx.set_resident(mode)
        ctx.set_batch_slots(5)                           # slots are refilled while others are still running
        out[mode] = ctx.solve_batch(imgs, p)
    ctx.set_resident(0)
    ctx.set_batch_slots(0)
    for a, b in zip(out[1], out[2]):
        assert a["iters"] == b["iters"] and a["deff"] == b["deff"] and a["conv"] == b["conv"] and a["pathflag"] == b["pathflag"]


# ----------------------------------------------------------------------------- opt-in accelerated solver (non-parity mode)

@pytest.mark.parametrize("nphase,Ds,Dg", [(2, 1e-2, 0.0), (3, 0.0, 30.0)])
def test_chebyshev_solver_reaches_the_limit_of_the_reference_iteration(ctx, nphase, Ds, Dg):
    """solver = 1 (csrc/chebyshev.cu) solves the same linear system as the reference's damped Jacobi, so it must land
    on the limit of that iteration -- here reached by brute force, 400 000 reference sweeps on a small domain -- in a
    small fraction of the sweeps.  Non-parity mode: the reference's stop rule and its intermediate iterates do not apply."""
    img = blobs(9 + nphase, (48, 64), levels=(0, 150, 255), fracs=(0.3, 0.4), smooth=2)
    base = dict(Ds=Ds, Df=1.0, Dg=Dg, CL=0.2, CR=1.3)
    ctx.set_kernel(0)
    ctx.domain_load(img, nphase, E.default_params(**base))
    ctx.sweeps(400000)
    limit = ctx.flux()[0]
    field = ctx.get_field()
    ctx.domain_load(img, nphase, E.default_params(solver=1, residual_tol=1e-11, **base))
    got = ctx.solve(0.0, 100000)
    assert got["iters"] < 20000
    assert got["conv"] <= 1e-11                                   # the relative residual reached
    assert rel(got["deff_raw"], limit) < 1e-9
    f = ctx.get_field()
    live = np.isfinite(field) & np.isfinite(f)
    assert np.max(np.abs(f[live] - field[live])) < 1e-8
    # through the driver: same answer, stages and all
    mode = E.MODE_3PH if nphase == 3 else E.MODE_2PH_BATCH
    r = ctx.solve_image(img, E.default_params(solver=1, residual_tol=1e-11, mode=mode, max_iter=100000, **base))
    assert rel(r["deff_raw"], limit) < 1e-8                       # (the brute-force limit itself is only that converged)
    ctx.domain_load(blobs(1, (16, 16)), 2, E.default_params())     # leave a plain-solver domain resident

# ----------------------------------------------------------------------------- K2 (TMA tiled) == K3 (streaming)

@pytest.mark.parametrize("shape", [(40, 300), (131, 257), (64, 120), (24, 16), (300, 1000)])
@pytest.mark.parametrize("nphase", [2, 3])
def test_tiled_kernel_is_bitwise_identical_to_streaming(ctx, shape, nphase):
    """Overlapped temporal blocking is algebraically T plain sweeps; both kernels use the same
    FMA order, so the fields must agree bit for bit for every depth T and sweep count."""
    img = blobs(shape[0] * 7 + nphase, shape, levels=(0, 150, 255), fracs=(0.3, 0.4), smooth=2)
    p = E.default_params(Ds=0.0 if nphase == 3 else 1e-3, Df=1.0, Dg=80.0, CL=0.25, CR=1.5)
    ctx.set_kernel(1)
    ctx.domain_load(img, nphase, p)
    ctx.sweeps(29)
    ref = ctx.get_field()
    dref, _ = ctx.flux()
    for kernel in (2, 3, 4):               # the default layout, then both thread layouts by name (4 x 4 and 2 x 8 cells per thread)
        for T in range(1, 9):
            ctx.set_kernel(kernel, T)
            ctx.domain_load(img, nphase, p)
            ctx.sweeps(29)                 # 29 = k*T + remainder: exercises the tail launches too
            got = ctx.get_field()
            assert np.array_equal(got, ref, equal_nan=True), (kernel, T, np.nanmax(np.abs(got - ref)))
            d, _ = ctx.flux()
            assert d == dref or (np.isnan(d) and np.isnan(dref))
    ctx.set_kernel(0)


def test_tiled_kernel_full_solve_matches_oracle(ctx):
    img = blobs(77, (96, 160))
    for T in (2, 4, 8):
        ctx.set_kernel(2, T)
        got = ctx.solve_image(img, E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=2000, max_iter=200000))
        ref = O.solve_image(img, O.make_opts(Ds=1e-3, Df=1.0, nphase=2, check_every=2000, max_iter=200000), O.MODE_2PH_BATCH)
        check_against_oracle(got, ref)
    ctx.set_kernel(0)


# ----------------------------------------------------------------------------- multi-GPU slabs (needs >= 2 GPUs)

def test_slab_decomposition_matches_single_gpu():
    """Row slabs with NCCL halo exchange + flux all-reduce == the undecomposed domain (own rows
    bit-identical, Deff to 1e-12, same sweep counts); one process per GPU via torchrun."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "scripts", "slab_check.py"), "--size", "300x500"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert '"ok": true' in out.stdout


# ----------------------------------------------------------------------------- edge cases and error behaviour

def test_api_error_paths_and_degenerate_inputs(ctx):
    """Every entry point returns a status instead of crashing (reference quirks Q16, Q24): calls in
    the wrong order, bad arguments, empty batches, no-sweep solves."""
    L = ctx._L
    fresh = E.Deff2D(0)
    try:
        with pytest.raises(E.Deff2DError):
            fresh.sweeps(1)                                    # no domain loaded
        with pytest.raises(E.Deff2DError):
            fresh.flux()
        assert fresh.solve_batch(np.zeros((0, 8, 8), np.uint8), E.default_params(mode=E.MODE_2PH_BATCH)) == []
    finally:
        fresh.close()
    img = blobs(9, (24, 40))
    with pytest.raises(E.Deff2DError):
        ctx.solve_image(img, E.default_params(amp_x=0))        # cuh:1672-1675
    bad = E.default_params()
    bad.mode = 7
    with pytest.raises(E.Deff2DError):
        ctx.solve_image(img, bad)
    assert L.deff2d_set_kernel(ctx._h, 99, 0) < 0 and L.deff2d_set_floodfill(ctx._h, 5) < 0
    # tol >= 100: the reference loop body never runs (cuh:1232) -- 0 sweeps, Deff = the initial deffNew
    p = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, tol=100.0)
    got = ctx.solve_image(img, p)
    ref = O.solve_image(img, O.make_opts(Ds=1e-3, Df=1.0, nphase=2, tol=100.0), O.MODE_2PH_BATCH)
    assert got["iters"] == ref["iters"] == [0] and got["deff"] == ref["deff"]
    # the same inside a batch (falls back to the per-image path) and MaxIter = 1
    gb = ctx.solve_batch(np.stack([img, img]), p)
    assert [g["iters"] for g in gb] == [[0], [0]]
    p1 = E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, max_iter=1)
    got = ctx.solve_batch(np.stack([img, img[::-1].copy()]), p1)
    for k, im in enumerate((img, img[::-1].copy())):
        ref = O.solve_image(im, O.make_opts(Ds=1e-3, Df=1.0, nphase=2, max_iter=1), O.MODE_2PH_BATCH)
        assert got[k]["iters"] == ref["iters"] == [1] and rel(got[k]["deff"], ref["deff"]) < 1e-12
    ctx.domain_load(img, 2, E.default_params(Ds=1e-3, Df=1.0))
    f0 = ctx.get_field()
    ctx.sweeps(0)
    assert np.array_equal(ctx.get_field(), f0)


@pytest.mark.parametrize("shape", [(2, 2), (2, 37), (41, 2), (3, 3), (5, 300), (300, 5), (64, 64), (65, 63), (129, 127)])
def test_small_and_ragged_domains_all_kernels(ctx, shape):
    """Domains smaller than, equal to and just around one tile, thin strips included: the default
    (tiled, graph-replayed) path equals the streaming kernel bit for bit and the oracle to 1e-13."""
    img = blobs(shape[0] * 31 + shape[1], shape, levels=(0, 150, 255), fracs=(0.4, 0.3), smooth=1)
    for nphase, Ds, Dg in ((2, 1e-2, 0.0), (3, 0.0, 20.0)):
        p = E.default_params(Ds=Ds, Df=1.0, Dg=Dg, CL=0.1, CR=0.9)
        D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, Dg)
        G, _ = O.floodfill(O.grid_mask(img, 1, 1, 200 if nphase == 3 else 150))
        A, b = O.discretize(D, 0.1, 0.9, G if nphase == 3 else None)
        ref = O.sweeps(A, b, O.init_x(shape[1], shape[0], 0.1, 0.9), 75)
        out = {}
        for kernel, resident in ((1, 0), (0, 0), (0, 1)):      # streaming; default (cluster-resident at these sizes); tiled
            ctx.set_kernel(kernel)
            ctx.set_resident(resident)
            ctx.domain_load(img, nphase, p)
            ctx.sweeps(75)                                     # tiled: whole passes at the default depth + a remainder pass
            out[(kernel, resident)] = ctx.get_field()
        assert np.array_equal(out[(0, 0)], out[(1, 0)], equal_nan=True) and np.array_equal(out[(0, 1)], out[(1, 0)], equal_nan=True)
        assert np.array_equal(np.isnan(out[(0, 0)]), np.isnan(ref))
        if not np.all(np.isnan(ref)):
            assert np.nanmax(np.abs(out[(0, 0)] - ref)) < 1e-13
    ctx.set_kernel(0)
    ctx.set_resident(0)


def test_graph_replay_long_runs_match_streaming(ctx):
    """Runs long enough to go through the CUDA-graph path (32 passes per graph) with a remainder,
    interleaved with table changes (new stage) that must not be served by a stale graph."""
    img = blobs(123, (70, 90), levels=(0, 150, 255), fracs=(0.3, 0.4))
    p = E.default_params(Ds=0.0, Df=1.0, Dg=10.0)
    res = {}
    ctx.set_resident(1)                                        # a 70 x 90 domain would otherwise run cluster-resident
    for kernel in (1, 0):
        ctx.set_kernel(kernel)
        ctx.domain_load(img, 3, p)
        ctx.sweeps(700)
        ctx.set_D(0.0, 1.0, 1000.0)                            # next continuation stage: same graph shape, new table
        ctx.sweeps(600)
        ctx.set_field(np.nan_to_num(ctx.get_field()) * 0.5)     # dead cells (A0 = 0) read back as NaN: do not inject them
        ctx.sweeps(300)
        res[kernel] = (ctx.get_field(), ctx.flux())
    ctx.set_resident(0)
    assert np.array_equal(res[0][0], res[1][0], equal_nan=True) and res[0][1] == res[1][1]
    assert np.isfinite(res[0][1][0])
    ctx.set_kernel(0)


def test_single_process_multi_gpu_slabs_match_single_gpu(ctx):
    """deff2d_solve_image_slabs (csrc/multi.cpp): one host process, one thread per GPU, NCCL halo
    exchange -- same stage sequence, sweep counts, Deff (1e-12) and field as the single-GPU solve."""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    others = [E.Deff2D(d) for d in range(1, n)]
    try:
        img3 = blobs(71, (300, 260), levels=(0, 150, 255), fracs=(0.3, 0.4))
        img2 = blobs(72, (280, 200))
        cases = [(img3, E.default_params(Ds=0.0, Df=1.0, Dg=300.0, mode=E.MODE_3PH, check_every=400, max_iter=4000, tol=1e-4, amp_y=2)),
                 (img2, E.default_params(Ds=1e-3, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=500, max_iter=6000, tol=1e-4)),
                 (img2, E.default_params(Ds=1.0, Df=12000.0, mode=E.MODE_2PH_SINGLE, check_every=500, max_iter=3000, tol=1e-4))]
        for img, p in cases:
            one = ctx.solve_image(img, p, want_field=True)
            many = E.solve_image_slabs([ctx] + others, img, p, want_field=True)
            for k in ("iters", "nstages", "pathflag", "porosity", "SVF", "LVF", "total_iters", "n_cells", "stage_D"):
                assert many[k] == one[k], (k, many[k], one[k])
            assert rel(many["deff"], one["deff"]) < 1e-12 and abs(many["conv"] - one["conv"]) < 1e-12
            for a, b in zip(many["stage_deff_raw"], one["stage_deff_raw"]):
                assert rel(a, b) < 1e-12
            assert np.array_equal(np.isnan(many["field"]), np.isnan(one["field"]))
            assert np.nanmax(np.abs(many["field"] - one["field"])) < 1e-13
    finally:
        for c in others:
            c.close()


def test_opt_in_residual_stop_rule(ctx):
    """residual_tol > 0 (non-parity mode): the loop stops at the first check where the
    reference's Residual (cuh:451-494) is <= residual_tol, whatever the Deff change does."""
    img = blobs(55, (48, 64))
    base = dict(Ds=1e-2, Df=1.0, mode=E.MODE_2PH_BATCH, check_every=250, max_iter=100000)
    ref = ctx.solve_image(img, E.default_params(tol=1e-9, **base))            # reference rule, tight
    # the reference's Residual, kept verbatim, scales every face by dy/dx and plateaus (here near 8.4e-4)
    # instead of vanishing at the solution: the tolerance has to sit above that plateau
    rtol = 8.45e-4
    got = ctx.solve_image(img, E.default_params(tol=1e-9, residual_tol=rtol, **base))
    it = got["iters"][0]
    assert (it - 1) % 250 == 0 and 1 < it < ref["iters"][0]
    # the residual at the stopping check is below the tolerance, one check earlier it is not
    p = E.default_params(**base)
    ctx.domain_load(img, 2, p)
    ctx.sweeps(it - 250)
    assert ctx.residual() > rtol
    ctx.sweeps(250)
    assert ctx.residual() <= rtol
    assert rel(ctx.flux()[0], got["deff_raw"]) < 1e-12


def test_bundled_00042_stage_checkpoints_from_the_reference(ctx, golden_images):
    """Shipped input.txt verbatim on the bundled 00042.jpg (2.0 M cells, 3-phase): un-normalised Deff
    at every check of pre-conditioning stage 1 (DCG = 10, stops at sweep 150 001 because the signed
    change crosses zero) and at the first 11 checks of stage 2 (DCG = 100), as produced by the
    reference's own host code + kernel body."""
    img = golden_images["00042"]
    p = E.default_params()                                     # Ds 0, Df 1, Dg 1237500, tol 1e-5, MaxIter 5e5
    ctx.domain_load(img, 3, p)
    ctx.set_D(0.0, 1.0, 10.0)                                  # cuh:1492-1531
    r1 = ctx.solve(p.tol * 10, 1000000)                        # cuh:1501-1502
    stage1 = [9.255802317031, 3.436492926626, 3.377988028260, 3.346242512934, 3.325747837646, 3.311021646843,
              3.299831844933, 3.291133004473, 3.284350477833, 3.279116522299, 3.275166467026, 3.272294662892,
              3.270333924566, 3.269144735927, 3.268608828991, 3.268624930899]
    assert r1["iters"] == 150001 and len(r1["trace"]) == 16
    for a, b in zip(r1["trace"], stage1):
        assert rel(a, b) < 1e-9
    assert abs(r1["conv"] - (-4.926e-06)) < 1e-9
    ctx.set_D(0.0, 1.0, 100.0)                                 # stage 2, warm start
    r2 = ctx.solve(p.tol * 10, 100002)
    stage2 = [28.86247113599, 28.84390816332, 28.81427679676, 28.78700150716, 28.76275598923, 28.74128132848,
              28.72220433724, 28.70516255849, 28.68982560472, 28.67589684461, 28.66311119152]
    assert r2["iters"] == 100002 and len(r2["trace"]) == 11
    for a, b in zip(r2["trace"], stage2):
        assert rel(a, b) < 1e-9
