#!/usr/bin/env python
"""Generates tests/golden/* from the REFERENCE ITSELF.

Run in the build container (needs /root/reference and oracle/_ref built by
oracle/build_ref.sh):   python tests/golden/make_golden.py

What it records (nothing here comes from the oracle restatement or from the CUDA path):
  images.npz        the two bundled sample images decoded by the reference's own decoder
                    (stb_image v2.26 through Deff2D.cuh:342), so that GPU-box tests do not
                    depend on a JPEG decoder
  primitives.npz    full-precision outputs of the reference's DiscretizeMatrix2D[_ImpSolid],
                    FloodFill and JacobiGPU[PreCond] (kernel body run on host threads by the
                    CUDA shim) on small seeded inputs
  drivers.json      CSV rows + stdout of the reference PROGRAM (its own main) on known-answer
                    cases of the reference documentation and on the bundled 00000.jpg
The committed outputs pin oracle/deff_oracle.c (tests/test_oracle_golden.py) and the CUDA
path (tests/test_gpu_*.py).
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402

REFDIR = "/root/reference/Deff2DGPU"


def write_pgm(path, img):
    """Lossless P5 content; the reference's decoder sniffs content, not the extension."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def input_txt(**kw):
    d = dict(Phases=3, Ds=0, Df=1, Dg=1237500, MeshAmpX=1, MeshAmpY=1, InputName="00000.jpg", CR=1, CL=0,
             OutputName="out.csv", printCMap=0, CMapName="CMAP.csv", Convergence="1e-5", MaxIter="5e5",
             Verbose=1, RunBatch=0, NumImages=1)
    d.update(kw)
    order = ["Phases", "Ds", "Df", "Dg", "MeshAmpX", "MeshAmpY", "InputName", "CR", "CL", "OutputName",
             "printCMap", "CMapName", "Convergence", "MaxIter", "Verbose", "RunBatch", "NumImages"]
    return "Input File:\n" + "\n".join("%s: %s" % (k, d[k]) for k in order) + "\n"


def run_program(files, **kw):
    """files: {name: uint8 image or bytes}.  Returns csv text, stdout, cmap text (or None)."""
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        for name, content in files.items():
            p = os.path.join(tmp, name)
            if isinstance(content, (bytes, bytearray)):
                with open(p, "wb") as f:
                    f.write(content)
            else:
                write_pgm(p, content)
        with open(os.path.join(tmp, "input.txt"), "w") as f:
            f.write(input_txt(**kw))
        out = subprocess.run([O.REF_CPU_EXE], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                             check=True, timeout=3600).stdout.decode()
        with open(os.path.join(tmp, kw.get("OutputName", "out.csv"))) as f:
            csv = f.read()
        cmap = None
        cm = os.path.join(tmp, kw.get("CMapName", "CMAP.csv"))
        if os.path.exists(cm):
            with open(cm) as f:
                cmap = f.read()
        return csv, out, cmap
    finally:
        shutil.rmtree(tmp)


def kat_images():
    par = np.full((100, 100), 255, np.uint8); par[:30, :] = 0           # doc 5.3 parallel, eps=0.3
    ser = np.full((100, 100), 255, np.uint8); ser[:, :30] = 0           # doc 5.3 series
    wide = np.full((50, 100), 255, np.uint8); wide[:, :50] = 0          # doc 5.3.3 wide domain
    thin = np.zeros((100, 100), np.uint8); thin[:, 48:51] = 255         # doc 5.3.1 thin phase
    p3 = np.zeros((100, 100), np.uint8); p3[:30, :] = 255; p3[30:70, :] = 150   # doc 5.3.2
    return par, ser, wide, thin, p3


def main():
    assert O.reference("cpu") is not None, "run oracle/build_ref.sh first"
    # ---- images
    img0, n0 = O.ref_decode(os.path.join(REFDIR, "00000.jpg"))
    img42, n42 = O.ref_decode(os.path.join(REFDIR, "00042.jpg"))
    assert n0 == 1 and n42 == 1
    np.savez_compressed(os.path.join(HERE, "images.npz"), img00000=img0, img00042=img42)

    # ---- primitives at full precision
    prim = {}
    rng = np.random.default_rng(20261018)
    cases = []
    for k, (Ny, Nx) in enumerate([(12, 20), (17, 9), (16, 16), (5, 31)]):
        ph = rng.integers(0, 3, size=(Ny, Nx))
        img = np.choose(ph, [0, 150, 255]).astype(np.uint8)
        cases.append(img)
    # a case whose cell (0,0) is solid (FloodFill right-column seeding quirk, Deff2D.cuh:601)
    q = cases[0].copy(); q[0, 0] = 255; q[:, 10] = 255
    cases.append(q)
    # a fully blocked domain with fluid (0,0)
    q2 = np.zeros((10, 14), np.uint8); q2[:, 7] = 255
    cases.append(q2)
    for k, img in enumerate(cases):
        prim["img%d" % k] = img
        for nphase, thr in ((2, 150), (3, 200)):
            Ds = 0.0 if nphase == 3 else 1e-3
            D = O.fill_D(img, 1, 1, nphase, Ds, 1.0, 50.0)
            G = (img > thr).astype(np.uint32)
            Gf, pf = O.ref_floodfill(G)
            prim["flood%d_p%d" % (k, nphase)] = Gf.astype(np.uint8)
            prim["pathflag%d_p%d" % (k, nphase)] = np.int32(pf)
            A, b = O.ref_discretize(D, 0.25, 1.5, Gf if nphase == 3 else None)
            prim["A%d_p%d" % (k, nphase)] = A
            prim["b%d_p%d" % (k, nphase)] = b
            x0 = O.init_x(img.shape[1], img.shape[0], 0.25, 1.5)
            for maxit, pre in ((1, False), (37, False), (10001, False), (2500, True)):
                r = O.ref_jacobi(A, b, x0, D, 0.25, 1.5, 1e-7, maxit, precond=pre)
                tag = "%d_p%d_it%d%s" % (k, nphase, maxit, "pre" if pre else "")
                prim["x" + tag] = r["field"]
                prim["deff" + tag] = np.float64(r["deff_raw"])
                prim["conv" + tag] = np.float64(r["conv"])
                prim["iters" + tag] = np.int64(r["iters"])
    # bundled 00000.jpg, 2-phase batch parameters of BASELINE.md, 17-digit Deff + field
    D = O.fill_D(img0, 1, 1, 2, 1e-4, 1.0, 0.0)
    A, b = O.ref_discretize(D, 0.0, 1.0)
    r = O.ref_jacobi(A, b, O.init_x(128, 128, 0.0, 1.0), D, 0.0, 1.0, 1e-5, 500000)
    prim["x00000_2ph"] = r["field"]
    prim["deff00000_2ph"] = np.float64(r["deff_raw"])
    prim["conv00000_2ph"] = np.float64(r["conv"])
    prim["iters00000_2ph"] = np.int64(r["iters"])
    np.savez_compressed(os.path.join(HERE, "primitives.npz"), **prim)

    # ---- program-level known answers
    par, ser, wide, thin, p3 = kat_images()
    drivers = {}

    def rec(name, files, **kw):
        csv, out, cmap = run_program(files, **kw)
        lines = [l for l in out.splitlines() if l.startswith(("Iterations taken", "DCF =", "Number", "Pre-Cond"))]
        drivers[name] = {"input": {k: str(v) for k, v in kw.items()}, "csv": csv, "stdout_key_lines": lines}
        if cmap is not None:
            cl = cmap.splitlines()
            drivers[name]["cmap_head"] = cl[:6]
            drivers[name]["cmap_lines"] = len(cl)
            drivers[name]["cmap_tail"] = cl[-3:]
        print(name, csv.strip().splitlines()[-1], flush=True)

    rec("kat_parallel_2ph_batch", {"00000.jpg": par}, Phases=2, Ds=0.1, Df=1, RunBatch=1, NumImages=1)
    rec("kat_series_2ph_batch", {"00000.jpg": ser}, Phases=2, Ds=0.1, Df=1, RunBatch=1, NumImages=1)
    rec("kat_wide_2ph_batch", {"00000.jpg": wide}, Phases=2, Ds=0.1, Df=1, RunBatch=1, NumImages=1)
    rec("kat_thin_2ph_single", {"thin.jpg": thin}, Phases=2, Ds=1, Df=1237500, InputName="thin.jpg")
    rec("kat_parallel_3ph_single", {"p3.jpg": p3}, InputName="p3.jpg", printCMap=1)
    rec("kat_parallel_3ph_batch", {"00000.jpg": p3, "00001.jpg": par}, RunBatch=1, NumImages=2, printCMap=0)
    with open(os.path.join(REFDIR, "00000.jpg"), "rb") as f:
        jpg0 = f.read()
    rec("bundled00000_3ph_single", {"00000.jpg": jpg0})
    rec("bundled00000_2ph_batch", {"00000.jpg": jpg0}, Phases=2, Ds="1e-4", Df=1, RunBatch=1, NumImages=1)
    rec("bundled00000_2ph_single_Df1", {"00000.jpg": jpg0}, Phases=2, Ds="1e-4", Df=1)
    rec("bundled00000_2ph_Ds0_nan", {"00000.jpg": jpg0}, Phases=2, Ds=0, Df=1, RunBatch=1, NumImages=1)
    with open(os.path.join(HERE, "drivers.json"), "w") as f:
        json.dump(drivers, f, indent=1, sort_keys=True)


def make_jpeg_fixtures(out_path=None):
    """JPEG files (bytes) + the pixels the reference's own decoder (stbi_load, Deff2D.cuh:342)
    returns for them: the two bundled images and generated baseline / progressive /
    restart-interval / optimised-Huffman grayscale files.  Needs /root/reference, PIL and
    oracle/_ref/libref_cpu.so; run here, committed as tests/golden/jpeg.npz."""
    import io
    import tempfile
    from PIL import Image
    import _oracle as O
    out_path = out_path or os.path.join(HERE, "jpeg.npz")
    files = {}
    for name in ("00000", "00042"):
        files["bundled_" + name] = open("/root/reference/Deff2DGPU/%s.jpg" % name, "rb").read()
    rng = np.random.default_rng(0)
    z = rng.random((67, 131))
    for _ in range(3):
        z = (z + np.roll(z, 1, 0) + np.roll(z, 1, 1)) / 3
    g = (255 * (z - z.min()) / (z.max() - z.min())).astype(np.uint8)
    noise = rng.integers(0, 256, (120, 203), dtype=np.uint8)
    two = np.where(z < np.median(z), 0, 255).astype(np.uint8)[:33, :17]
    cases = {"smooth_q90": (g, dict(quality=90)), "smooth_q30": (g, dict(quality=30)),
             "smooth_q100_prog": (g, dict(quality=100, progressive=True)),
             "smooth_q75_prog": (g, dict(quality=75, progressive=True)),
             "noise_q95": (noise, dict(quality=95)), "noise_q50_prog": (noise, dict(quality=50, progressive=True)),
             "twotone_opt": (two, dict(quality=85, optimize=True)),
             "noise_restart": (noise, dict(quality=80, restart_marker_blocks=7)),
             "smooth_restart_rows": (g, dict(quality=85, restart_marker_rows=1)),
             "tiny_1x1": (np.array([[200]], dtype=np.uint8), dict(quality=90)),
             "tiny_9x7": (noise[:7, :9].copy(), dict(quality=90))}
    for name, (arr, kw) in cases.items():
        buf = io.BytesIO()
        Image.fromarray(arr).save(buf, "JPEG", **kw)
        files["gen_" + name] = buf.getvalue()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for name, data in files.items():
            p = os.path.join(d, "x.jpg")
            with open(p, "wb") as f:
                f.write(data)
            pix, nch = O.ref_decode(p)
            assert nch == 1
            out["file_" + name] = np.frombuffer(data, dtype=np.uint8)
            if not name.startswith("bundled_"):      # the bundled images' pixels already live in images.npz
                out["pix_" + name] = pix
    np.savez_compressed(out_path, **out)
    return out_path


if __name__ == "__main__":
    main()
    make_jpeg_fixtures()
