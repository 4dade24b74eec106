"""CPU tests of the host-side half of libdeff2d (no GPU compute): the coefficient LUT against
the oracle's materialised A/b, FloodFill and input parsing against the reference goldens, the
CSV/CMAP writers' formats, the image readers, and the C-ABI symbol table."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _emulate as EM
import _oracle as O
import effectivediffusivityfvm_b200 as E
from effectivediffusivityfvm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "deff2d.h")).read()
    declared = set(re.findall(r"\b(deff2d_[A-Za-z0-9_]+)\s*\(", hdr))
    L = _lib.lib()
    assert declared == set(L._declared), declared ^ set(L._declared)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (deff2d_[A-Za-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert L.deff2d_version() == 100


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.Deff2DError):
        E.Deff2D(0)


@pytest.mark.parametrize("nphase,Ds,Dg", [(2, 1e-3, 0.0), (2, 0.0, 0.0), (3, 0.0, 50.0), (3, 0.01, 1237500.0)])
@pytest.mark.parametrize("shape,CL,CR", [((12, 20), 0.0, 1.0), ((9, 17), 0.25, 1.5), ((1, 7), 0.0, 1.0), ((6, 2), 0.0, 1.0)])
def test_lut_matches_reference_matrix(nphase, Ds, Dg, shape, CL, CR):
    """Every weight the kernels will use equals (w/A0) * coefficient of the reference's A, b."""
    rng = np.random.default_rng(hash((nphase, shape)) % 2**32)
    Ny, Nx = shape
    if Ny == 1 or Nx == 1:
        pytest.skip("the reference reads out of bounds on 1-wide domains")
    img = np.choose(rng.integers(0, 3, size=shape), [0, 150, 255]).astype(np.uint8)
    omega = 2.0 / 3.0
    grid = None
    if nphase == 3:
        grid, _ = O.floodfill((img > 200).astype(np.uint32))
    D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, Dg)
    A, b = O.discretize(D, CL, CR, grid)
    A = A.reshape(Ny, Nx, 5)
    b = b.reshape(Ny, Nx)
    lut, dead = E.build_tables(Ds, 1.0, Dg, Nx, Ny, CL, CR, omega)
    codes = EM.phase_codes(img, nphase, grid=grid)
    idx = EM.lut_index(EM.pad_codes(codes), Nx, Ny)
    w = lut[idx]
    with np.errstate(divide="ignore", invalid="ignore"):
        dinv = omega / A[..., 0]
        exp = np.stack([dinv * -A[..., 1], dinv * -A[..., 2], dinv * -A[..., 3], dinv * -A[..., 4]], axis=-1)
        exp[:, 0, 0] = dinv[:, 0] * b[:, 0]          # Dirichlet faces ride on the ghost value 1.0
        exp[:, -1, 1] = dinv[:, -1] * b[:, -1]
    isdead = A[..., 0] == 0
    assert np.array_equal(dead[idx].astype(bool), isdead)
    exp[isdead] = 0.0
    pinned = (codes & 4) != 0
    assert np.all(w[pinned] == 0)
    exp[pinned] = 0.0                                  # identity row: x' = (1-w) x
    exp = exp + 0.0                                    # -0.0 -> +0.0
    assert np.array_equal(w + 0.0, exp)


@pytest.mark.parametrize("nphase,Ds,Dg", [(2, 1e-3, 0.0), (3, 0.01, 50.0)])
def test_compact_table_is_a_ranked_rearrangement_of_the_lut(nphase, Ds, Dg):
    """The planar table the tiled sweep gathers from holds exactly the weights of the 2048-entry table (include/deff2d.h:
    deff2d_compact_table): every possible neighbourhood has one slot, interior slots are a permutation of 0..31 /
    0..242, and in a 3-phase table single-phase neighbourhoods and single differing neighbours come first."""
    import ctypes as C
    from effectivediffusivityfvm_b200 import _lib
    L = _lib.lib()
    lut, _ = E.build_tables(Ds, 1.0, Dg, 64, 48, 0.0, 1.0, 2.0 / 3.0)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    clut = np.zeros(4 * 1024, dtype=np.float64)
    slot = np.zeros(2048, dtype=np.uint16)
    rc = L.deff2d_compact_table(lut.ctypes.data_as(C.POINTER(C.c_double)), nphase, clut.ctypes.data_as(C.POINTER(C.c_double)),
                                slot.ctypes.data_as(C.POINTER(C.c_uint16)))
    assert rc == 0
    clut = clut.reshape(4, 1024)
    lut = lut.reshape(2048, 4)
    idx = np.arange(2048)
    ph = [(idx >> s) & 3 for s in (0, 2, 4, 6, 8)]
    pinned = ((idx >> 10) & 1).astype(bool)
    possible = np.all([(q == 3) | (q < nphase) for q in ph], axis=0)
    assert np.all(slot[~possible] == 0xffff) and np.all(slot[possible] < 1024)
    inert = possible & (pinned | (ph[0] == 3))
    assert np.all(slot[inert] == 1023) and np.all(clut[:, 1023] == 0)
    live = possible & ~inert
    assert np.array_equal(clut[:, slot[live]].T, lut[live])            # same weights, bit for bit
    assert len(np.unique(slot[live])) == live.sum()                    # one slot per neighbourhood
    interior = live & np.all([q != 3 for q in ph[1:]], axis=0)
    n_int = nphase ** 5
    assert sorted(slot[interior]) == list(range(n_int))
    # the 32-bit halves the tiled sweep gathers on interface-rich media reassemble to the same doubles
    halves = np.zeros(8 * 1024, dtype=np.uint32)
    assert L.deff2d_split_table(clut.ctypes.data_as(C.POINTER(C.c_double)), 1, halves.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
    halves = halves.reshape(8, 1024).astype(np.uint64)
    assert np.array_equal((halves[4:] << np.uint64(32)) | halves[:4], np.ascontiguousarray(clut).view(np.uint64))
    if nphase == 3:                                                    # (2-phase: dense numbering, one line per centre phase)
        ndiff = sum((q != ph[0]).astype(int) for q in ph[1:])
        order = np.argsort(slot[interior])
        assert np.all(np.diff(ndiff[interior][order]) >= 0)            # ranked by the number of differing neighbours
        first_line = slot[interior] < 16
        assert np.all(ndiff[interior][first_line] <= 1) and np.all(first_line[ndiff[interior] == 0])
    else:
        assert np.array_equal(slot[interior], ph[0][interior] * 16 + (ph[1] | ph[2] << 1 | ph[3] << 2 | ph[4] << 3)[interior])


@pytest.mark.parametrize("nphase,Ds,Dg", [(2, 1e-3, 0.0), (3, 0.0, 50.0)])
def test_matrix_free_sweeps_match_oracle(nphase, Ds, Dg):
    """The LUT formulation run in numpy tracks the oracle's A/b sweeps to rounding."""
    rng = np.random.default_rng(5)
    Ny, Nx = 23, 37
    img = np.choose(rng.integers(0, 3, size=(Ny, Nx)), [0, 150, 255]).astype(np.uint8)
    grid = O.floodfill((img > 200).astype(np.uint32))[0] if nphase == 3 else None
    D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, Dg)
    A, b = O.discretize(D, 0.0, 1.0, grid)
    x0 = O.init_x(Nx, Ny, 0.0, 1.0)
    ref = O.sweeps(A, b, x0, 200)
    lut, dead = E.build_tables(Ds, 1.0, Dg, Nx, Ny, 0.0, 1.0, 2.0 / 3.0)
    pc = EM.pad_codes(EM.phase_codes(img, nphase, grid=grid))
    got = EM.sweep(EM.pad_field(x0), pc, lut, Nx, Ny, nsweeps=200)[1:Ny + 1, EM.XOFF:EM.XOFF + Nx]
    assert np.max(np.abs(got - ref)) < 1e-13


@pytest.mark.parametrize("k", range(6))
@pytest.mark.parametrize("nphase", [2, 3])
def test_floodfill_matches_reference(golden_prims, k, nphase):
    img = golden_prims["img%d" % k]
    thr = 150 if nphase == 2 else 200
    g, pf = E.floodfill((img > thr).astype(np.uint8))
    assert np.array_equal(g, golden_prims["flood%d_p%d" % (k, nphase)])
    assert pf == int(golden_prims["pathflag%d_p%d" % (k, nphase)])


def test_floodfill_large_random_vs_oracle():
    rng = np.random.default_rng(11)
    for p in (0.35, 0.45, 0.6):
        m = (rng.random((150, 190)) < p).astype(np.uint8)
        g, pf = E.floodfill(m)
        go, pfo = O.floodfill(m.astype(np.uint32))
        assert np.array_equal(g, go.astype(np.uint8)) and pf == pfo


def test_read_input_file_shipped_defaults(tmp_path):
    txt = ("Input File:\nPhases: 3\nDs: 0\nDf: 1\nDg: 1237500\nMeshAmpX: 1\nMeshAmpY: 1\nInputName: 00042.jpg\n"
           "CR: 1\nCL: 0\nOutputName: singleTest.csv\nprintCMap: 1\nCMapName: CMAP_00042.csv\nConvergence: 1e-5\n"
           "MaxIter: 5e5\nVerbose: 1\nRunBatch: 0\nNumImages: 500")      # Deff2DGPU/input.txt verbatim
    f = tmp_path / "input.txt"
    f.write_text(txt)
    inp = E.read_input_file(f)
    assert (inp.nphase, inp.batch, inp.num_images, inp.print_cmap) == (3, 0, 500, 1)
    assert (inp.p.Ds, inp.p.Df, inp.p.Dg) == (0.0, 1.0, 1237500.0)
    assert inp.p.max_iter == 500000 and inp.p.tol == 1e-5 and inp.p.verbose == 1       # "5e5" via double (cuh:299-300)
    assert inp.input_name == b"00042.jpg" and inp.output_name == b"singleTest.csv" and inp.cmap_name == b"CMAP_00042.csv"
    assert inp.p.mode == E.MODE_3PH
    # case-sensitive keys, unknown keys ignored (cuh:258-312)
    f.write_text("phases: 2\nPhases: 2\nSolver: fancy\nRunBatch: 1\nDf: 7\n")
    inp = E.read_input_file(f)
    assert inp.nphase == 2 and inp.p.mode == E.MODE_2PH_BATCH and inp.p.Df == 7.0
    with pytest.raises(E.Deff2DError):
        E.read_input_file(tmp_path / "missing.txt")


def test_csv_and_cmap_writers_match_reference_format(tmp_path, golden_drivers):
    L = _lib.lib()
    g = golden_drivers["bundled00000_3ph_single"]
    ref_row = g["csv"].strip().splitlines()
    inp = _lib.Input()
    inp.p = E.default_params()
    inp.nphase = 3
    inp.input_name = b"00000.jpg"
    out = tmp_path / "o.csv"
    inp.output_name = str(out).encode()
    r = _lib.Result()
    r.SVF, r.LVF, r.pathflag, r.deff, r.n_cells, r.conv = 0.653931, 0.0, 1, 224673.610442892, 16384, 1.368e-08
    r.solve_ms = float(ref_row[1].split(",")[5]) * 1000
    assert L.deff2d_write_csv_single(C.byref(inp), C.byref(r)) == 0
    assert L.deff2d_write_csv_single(C.byref(inp), C.byref(r)) == 0        # "a+": header repeats (Q18)
    lines = out.read_text().splitlines()
    assert lines[0] == ref_row[0] and lines[2] == ref_row[0]
    assert lines[1].split(",")[:5] == ref_row[1].split(",")[:5]
    assert lines[1].split(",")[6:] == ref_row[1].split(",")[6:]
    # 2-phase batch rows
    g2 = golden_drivers["bundled00000_2ph_batch"]["csv"].strip().splitlines()
    inp.nphase = 2
    inp.p.Ds, inp.p.Df = 1e-4, 1.0
    out2 = tmp_path / "b.csv"
    inp.output_name = str(out2).encode()
    rs = (_lib.Result * 1)()
    rs[0].porosity, rs[0].pathflag, rs[0].deff, rs[0].n_cells, rs[0].conv, rs[0].last_df = \
        0.3460693359375, 1, 0.18169102773720014, 16384, 6.546346763587775e-06, 1.0
    rs[0].solve_ms = float(g2[1].split(",")[4]) * 1000
    assert L.deff2d_write_csv_batch(C.byref(inp), rs, 1) == 0
    lines = out2.read_text().splitlines()
    assert lines[0] == g2[0]
    assert lines[1].split(",")[:4] == g2[1].split(",")[:4] and lines[1].split(",")[5:] == g2[1].split(",")[5:]
    # CMAP
    gc = golden_drivers["kat_parallel_3ph_single"]
    field = np.linspace(0, 1, 12).reshape(3, 4)
    cm = tmp_path / "c.csv"
    assert L.deff2d_write_cmap(str(cm).encode(), field.ctypes.data_as(_lib.c_double_p), 4, 3) == 0
    lines = cm.read_text().splitlines()
    assert lines[0] == gc["cmap_head"][0] == "X,Y,C"
    assert lines[1] == "0,0,0.000e+00" and lines[2].startswith("1,0,") and lines[5].startswith("0,1,")   # x inner, y outer
    assert len(lines) == 13
    assert re.fullmatch(r"\d+,\d+,-?\d\.\d{3}e[+-]\d\d", gc["cmap_head"][1])


def test_image_readers(tmp_path, golden_images):
    img = golden_images["00000"]
    p = tmp_path / "a.jpg"           # content sniffing: PGM bytes under a .jpg name
    with open(p, "wb") as f:
        f.write(b"P5\n# comment\n%d %d\n255\n" % (img.shape[1], img.shape[0]) + img.tobytes())
    got, ch = E.load_image(p)
    assert ch == 1 and np.array_equal(got, img)
    import cv2
    for params in ([], [cv2.IMWRITE_PNG_COMPRESSION, 9]):
        q = tmp_path / "b.png"
        cv2.imwrite(str(q), img, params)
        got, ch = E.load_image(q)
        assert ch == 1 and np.array_equal(got, img)
    rgb = np.stack([img, img // 2, 255 - img], axis=-1)
    cv2.imwrite(str(tmp_path / "c.png"), rgb[..., ::-1])
    got, ch = E.load_image(tmp_path / "c.png")
    assert ch == 3
    exp = ((rgb[..., 0].astype(int) * 77 + rgb[..., 1].astype(int) * 150 + rgb[..., 2].astype(int) * 29) >> 8).astype(np.uint8)
    assert np.array_equal(got, exp)
    with pytest.raises(E.Deff2DError):
        E.load_image(tmp_path / "nope.png")


def test_jpeg_decoder_matches_reference_decoder(tmp_path, golden_dir, golden_images):
    """The library's own JPEG decoder returns, pixel for pixel, what the reference's decoder
    (stbi_load(..., 1), Deff2D.cuh:342) returned for the same files: the two bundled images and
    generated baseline / progressive / restart-interval / optimised-Huffman files
    (tests/golden/jpeg.npz, made by tests/golden/make_golden.py)."""
    z = np.load(os.path.join(golden_dir, "jpeg.npz"))
    names = [k[5:] for k in z.files if k.startswith("file_")]
    assert len(names) >= 12
    for name in names:
        path = tmp_path / (name + ".jpg")
        path.write_bytes(z["file_" + name].tobytes())
        got, ch = E.load_image(path)
        want = golden_images[name[8:]] if name.startswith("bundled_") else z["pix_" + name]
        assert ch == 1
        assert got.shape == want.shape and np.array_equal(got, want), name


def test_jpeg_decoder_rejects_garbage(tmp_path):
    p = tmp_path / "bad.jpg"
    p.write_bytes(b"\xff\xd8\xff\xe0\x00\x10JFIF" + b"\x00" * 40)
    with pytest.raises(E.Deff2DError):
        E.load_image(p)
    p.write_bytes(b"\xff\xd8")
    with pytest.raises(E.Deff2DError):
        E.load_image(p)


def test_cmap_writer_parallel_blocks_and_npy(tmp_path):
    """createCMAP's text (cuh:497-524) from the block-parallel writer: byte-identical to per-cell
    C formatting, row blocks in order; the .npy companion round-trips through numpy.load."""
    L = _lib.lib()
    rng = np.random.default_rng(3)
    Ny, Nx = 77, 1030                                   # several row blocks, ragged last block
    f = rng.random((Ny, Nx)) * 10.0 ** rng.integers(-12, 3, size=(Ny, Nx))
    f[5, 7] = np.nan
    f[6, 8] = 0.0
    f[7, 9] = -1.5e-300
    p = tmp_path / "cmap.csv"
    assert L.deff2d_write_cmap(str(p).encode(), f.ctypes.data_as(_lib.c_double_p), Nx, Ny) == 0
    lines = p.read_text().split("\n")
    assert lines[0] == "X,Y,C" and lines[-1] == "" and len(lines) == Ny * Nx + 2
    libc = C.CDLL(None)
    libc.snprintf.restype = C.c_int
    buf = C.create_string_buffer(64)
    for (i, j) in [(0, 0), (0, Nx - 1), (5, 7), (6, 8), (7, 9), (63, 1029), (64, 0), (Ny - 1, Nx - 1)] + \
            [tuple(x) for x in rng.integers(0, [Ny, Nx], size=(200, 2))]:
        libc.snprintf(buf, 64, b"%d,%d,%1.3e", C.c_int(int(j)), C.c_int(int(i)), C.c_double(float(f[i, j])))
        assert lines[1 + i * Nx + j] == buf.value.decode(), (i, j)
    q = tmp_path / "field.npy"
    assert L.deff2d_write_field_npy(str(q).encode(), f.ctypes.data_as(_lib.c_double_p), Nx, Ny) == 0
    g = np.load(q)
    assert g.shape == (Ny, Nx) and g.dtype == np.float64 and np.array_equal(g, f, equal_nan=True)


def test_input_file_extension_keys(tmp_path):
    """Keys the reference parser does not know are ignored by it (cuh:261-311), so the extensions
    Devices: and FieldNpy: keep an input.txt usable by both programs."""
    p = tmp_path / "input.txt"
    p.write_text("Phases: 2\nDevices: 4\nFieldNpy: 1\nRunBatch: 1\nNumImages: 3\nSomethingElse: 9\n")
    inp = E.read_input_file(p)
    assert (inp.nphase, inp.devices, inp.field_npy, inp.batch, inp.num_images) == (2, 4, 1, 1, 3)
    p.write_text("Phases: 3\n")
    inp = E.read_input_file(p)
    assert (inp.devices, inp.field_npy) == (1, 0)


def test_jpeg_decoder_random_files_against_reference_decoder(tmp_path):
    """Freshly encoded grayscale JPEGs (random content, size, quality, baseline / progressive /
    optimised / restart markers) through the library's decoder and through the reference's own
    decoder built from /root/reference (oracle/_ref): identical pixels.  Skipped where the
    reference build or PIL is not available (the committed fixtures cover that case)."""
    if O.reference("cpu") is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built")
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(2026)
    for k in range(14):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 130))
        kind = k % 3
        if kind == 0:
            a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == 1:
            yy, xx = np.mgrid[0:h, 0:w]
            a = ((np.sin(xx / 7.0) + np.cos(yy / 5.0) + 2) * 63).astype(np.uint8)
        else:
            a = np.where(rng.random((h, w)) < 0.5, 30, 220).astype(np.uint8)
        kw = dict(quality=int(rng.integers(5, 101)))
        if k % 2:
            kw["progressive"] = True
        if k % 4 == 0:
            kw["optimize"] = True
        if k % 5 == 0:
            kw["restart_marker_blocks"] = int(rng.integers(1, 9))
        p = tmp_path / ("r%d.jpg" % k)
        Image.fromarray(a).save(p, "JPEG", **kw)
        got, ch = E.load_image(p)
        want, ch2 = O.ref_decode(str(p))
        assert ch == ch2 == 1 and np.array_equal(got, want), (k, h, w, kw)


def test_png_decoder_variants_against_reference_decoder(tmp_path):
    """PNG bit depths and colour types (8/16-bit gray, 1/2/4-bit gray, gray+alpha, palette) through
    the library's decoder and the reference's: same pixels, same channel count (the drivers accept
    only channel count 1, cuh:1665-1668)."""
    if O.reference("cpu") is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built")
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    a8 = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    a16 = rng.integers(0, 65536, (37, 53), dtype=np.uint16)
    cases = {"L": (Image.fromarray(a8), {}), "I16": (Image.fromarray(a16), {}), "bit1": (Image.fromarray(a8 > 127), {}),
             "LA": (Image.fromarray(np.stack([a8, 255 - a8], -1), "LA"), {}), "P": (Image.fromarray(a8).convert("P"), {}),
             "bit4": (Image.fromarray((a8 >> 4) << 4), {"bits": 4}), "bit2": (Image.fromarray((a8 >> 6) << 6), {"bits": 2})}
    for name, (im, kw) in cases.items():
        p = tmp_path / (name + ".png")
        im.save(p, "PNG", **kw)
        got, ch = E.load_image(p)
        want, ch2 = O.ref_decode(str(p))
        assert ch == ch2, name
        assert np.array_equal(got, want), name


def test_header_is_plain_c_and_links(tmp_path):
    """include/deff2d.h is the drop-in boundary: it must compile as C (no C++ or CUDA types in any
    signature) and a plain C program must link against libdeff2d.so and call the host-side entry
    points without a GPU."""
    src = tmp_path / "use_abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "deff2d.h"
int main(void) {
    deff2d_params p;
    deff2d_default_params(&p);
    if (deff2d_version() != DEFF2D_VERSION || p.Dg != 1237500.0 || p.mode != DEFF2D_MODE_3PH) return 2;
    unsigned char grid[12] = {0,0,1,0, 0,1,1,0, 0,0,0,0};
    int pf = deff2d_floodfill(grid, 4, 3);
    static double lut[2048 * 4];
    static unsigned char dead[2048];
    if (deff2d_build_tables(0.0, 1.0, 5.0, 8, 8, 0.0, 1.0, 0.0, lut, dead) != DEFF2D_OK) return 3;
    int ow, oh, tw, th;
    if (deff2d_tile_geometry(8, &ow, &oh, &tw, &th) != DEFF2D_OK) return 4;
    printf("%d %d %dx%d %dx%d %zu %zu %zu\n", pf, (int)dead[1 | 1 << 2 | 1 << 4 | 1 << 6 | 1 << 8], ow, oh, tw, th,
           sizeof(deff2d_result), sizeof(deff2d_params), sizeof(deff2d_input));
    return 0;
}
''')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-ldeff2d", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert out[0] == "1"                      # the open path along the bottom row reaches the last column
    assert out[1] == "1"                      # an all-solid neighbourhood with Ds = 0 has A0 = 0 (quirk Q13)
    assert out[2] == "48x48" and out[3] == "64x64"
    # the ctypes mirror matches the C layout
    assert [int(v) for v in out[4:7]] == [C.sizeof(_lib.Result), C.sizeof(_lib.Params), C.sizeof(_lib.Input)]


def test_csv_batch_rows_streamed_equal_bulk_write(tmp_path):
    """deff2d_append_csv_batch_row (header, then one flushed row per image) produces the same bytes
    as deff2d_write_csv_batch, for both phase counts."""
    L = _lib.lib()
    rng = np.random.default_rng(4)
    for nphase in (2, 3):
        res = (_lib.Result * 3)()
        for k in range(3):
            res[k].porosity, res[k].SVF, res[k].LVF = rng.random(3)
            res[k].deff, res[k].conv, res[k].solve_ms = rng.random() * 10.0 ** rng.integers(-3, 6), rng.normal() * 1e-6, rng.random() * 1e4
            res[k].pathflag, res[k].n_cells, res[k].last_df = int(rng.integers(0, 2)), 65536, 1.0
        files = []
        for mode in ("bulk", "rows"):
            inp = _lib.Input()
            L.deff2d_default_params(C.byref(inp.p))
            inp.nphase = nphase
            inp.p.Ds, inp.p.Df, inp.p.Dg = 1e-3, 1.0, 1237500.0
            path = tmp_path / ("%s%d.csv" % (mode, nphase))
            inp.output_name = str(path).encode()
            if mode == "bulk":
                assert L.deff2d_write_csv_batch(C.byref(inp), res, 3) == 0
            else:
                assert L.deff2d_append_csv_batch_row(C.byref(inp), -1, None) == 0
                for k in range(3):
                    assert L.deff2d_append_csv_batch_row(C.byref(inp), k, C.byref(res[k])) == 0
            files.append(path.read_bytes())
        assert files[0] == files[1] and files[0].count(b"\n") == 4


def test_fraction_accumulation_equals_the_literal_loop():
    """calcPorosity / calcFracts3D add 1.0/total once per cell (cuh:402, cuh:437); the library jumps whole binades of
    the running sum instead of looping, and must land on the same bits -- including totals that are powers of two
    (exact additions), ties, and counts that end in the middle of a binade."""
    L = _lib.lib()
    rng = np.random.default_rng(0)
    cases = [(0, 7), (1, 3), (63, 100), (64, 100), (65, 100), (3000, 10000), (10000, 10000), (4096, 4096), (1 << 20, 1 << 20),
             (12345, 1 << 16), (700001, 1000003), (2 * 10**6, 3 * 10**6), (16384, 128 * 128), (5678901, 2007 * 1002 * 16 // 4)]
    for _ in range(30):
        total = int(rng.integers(2, 4_000_000))
        cases.append((int(rng.integers(0, total + 1)), total))
    for count, total in cases:
        inc, s = 1.0 / total, 0.0
        for _ in range(count):
            s += inc
        got = L.deff2d_accumulate_fraction(count, total)
        assert got == s, (count, total, got, s)
    # far beyond what a loop is pleasant for: monotone, close to the closed form, and additive consistency
    big = L.deff2d_accumulate_fraction(200_000_000, 268_435_456)
    assert big == 200_000_000 / 268_435_456            # power-of-two total: every addition is exact
    v = L.deff2d_accumulate_fraction(160_000_000, 268_435_457)
    assert abs(v - 160_000_000 / 268_435_457) < 1e-7


def test_tga_and_colour_only_formats_against_reference_decoder(tmp_path):
    """The reference decoder's remaining formats (cuh:342): gray TGA (raw and RLE, both row orders) is the one that
    yields a 1-channel image and must decode pixel for pixel; colour TGA, BMP and GIF always come back with 3 / 4
    channels, which the drivers refuse (cuh:1665-1668) -- the library must report the same size and channel count."""
    if O.reference("cpu") is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built")
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    g = rng.integers(0, 256, (29, 41), dtype=np.uint8)
    g[5:20, 3:30] = 200                                   # runs for the RLE encoder
    rgb = rng.integers(0, 256, (29, 41, 3), dtype=np.uint8)

    def tga(path, arr, rle, top_down, gray=True):
        h, w = arr.shape[:2]
        itype = (3 if gray else 2) + (8 if rle else 0)
        hdr = bytes([0, 0, itype, 0, 0, 0, 0, 0, 0, 0, 0, 0, w & 255, w >> 8, h & 255, h >> 8, 8 if gray else 24, 0x20 if top_down else 0])
        rows = arr if top_down else arr[::-1]
        px = rows.reshape(h * w, -1)[:, ::-1] if not gray else rows.reshape(h * w, 1)      # BGR in the file
        body = bytearray()
        if not rle:
            body += px.tobytes()
        else:
            k, n = 0, h * w
            while k < n:
                run = 1
                while k + run < n and run < 128 and np.array_equal(px[k + run], px[k]):
                    run += 1
                if run > 1:
                    body += bytes([128 | (run - 1)]) + px[k].tobytes()
                    k += run
                else:
                    lit = 1
                    while k + lit < n and lit < 128 and not (k + lit + 1 < n and np.array_equal(px[k + lit], px[k + lit + 1])):
                        lit += 1
                    body += bytes([lit - 1]) + px[k:k + lit].tobytes()
                    k += lit
        path.write_bytes(hdr + bytes(body))

    files = []
    for rle in (False, True):
        for top in (False, True):
            p = tmp_path / ("g_%d_%d.tga" % (rle, top))
            tga(p, g, rle, top)
            files.append((p, True))
    p = tmp_path / "c.tga"
    tga(p, rgb, True, False, gray=False)
    files.append((p, True))
    for name, im, fmt in (("rgb.bmp", Image.fromarray(rgb), "BMP"), ("pal.bmp", Image.fromarray(g).convert("P"), "BMP"),
                          ("gray.bmp", Image.fromarray(g), "BMP"), ("a.gif", Image.fromarray(g), "GIF")):
        p = tmp_path / name
        im.save(p, fmt)
        files.append((p, False))
    for p, pixels in files:
        want, ch2 = O.ref_decode(str(p))
        L = _lib.lib()
        ptr = _lib.c_ubyte_p()
        W, H, ch = C.c_int(0), C.c_int(0), C.c_int(0)
        assert L.deff2d_load_image(str(p).encode(), C.byref(ptr), C.byref(W), C.byref(H), C.byref(ch)) == 0, p.name
        try:
            assert (H.value, W.value) == want.shape and ch.value == ch2, (p.name, H.value, W.value, ch.value, want.shape, ch2)
            if pixels:
                got = np.ctypeslib.as_array(ptr, shape=(H.value, W.value)).copy()
                assert np.array_equal(got, want), p.name
        finally:
            L.deff2d_free(ptr)
