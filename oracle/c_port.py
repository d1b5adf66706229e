"""Build + ctypes binding of oracle/mode_a.c, the plain-C restatement of the Mode A scorer.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``build()`` compiles it with gcc (no FMA contraction, so the
fp64 projection keeps cv2's operation order; OpenMP for the all-cores variant) into oracle/_build/libmodea.so;
``score`` has the signature and the result dict of oracle.mode_a.score.  Pinned to the reference's golden vectors
in tests/test_oracle_cpu.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mode_a.c")
LIB = os.path.join(HERE, "_build", "libmodea.so")
_lib = None


def build(force=False):
    """gcc -O2 -fopenmp -ffp-contract=off -shared; returns the library path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build the C restatement of the oracle")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [gcc, "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed on oracle/mode_a.c:\n" + res.stdout + res.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.modea_gray.restype = None
        lib.modea_gray.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        lib.modea_score.restype = C.c_int
        lib.modea_score.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_int]
        _lib = lib
    return _lib


def gray_from_rgb(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    out = np.empty(rgb.shape[:-1], dtype=np.uint8)
    load().modea_gray(rgb.ctypes.data, out.size, out.ctypes.data)
    return out


def score(gray, cams, c, ref, thr, wid=5, want_ncc=True, threads=0):
    """oracle.mode_a.score through the C restatement.  threads: 0 = all cores, 1 = the scalar port."""
    gray = np.ascontiguousarray(gray, dtype=np.uint8)
    V, H, W = gray.shape
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1, 3))
    ref = np.ascontiguousarray(np.asarray(ref).reshape(-1).astype(np.int32))
    N = c.shape[0]
    Rrt = np.ascontiguousarray(cams.R.reshape(V, 9))
    t = np.ascontiguousarray(cams.t.reshape(V, 3))
    k4 = np.ascontiguousarray(np.stack([cams.fx, cams.fy, cams.cx, cams.cy], axis=1))
    vis = np.zeros((N, V), dtype=np.uint8)
    avg = np.zeros(N)
    count = np.zeros(N, dtype=np.int32)
    xy = np.zeros((N, 2))
    ncc = np.zeros((N, V)) if want_ncc else None
    rc = load().modea_score(gray.ctypes.data, V, H, W, Rrt.ctypes.data, t.ctypes.data, k4.ctypes.data, N, c.ctypes.data,
                            ref.ctypes.data, float(thr), int(wid), vis.ctypes.data, avg.ctypes.data, count.ctypes.data,
                            xy.ctypes.data, ncc.ctypes.data if want_ncc else None, int(threads))
    if rc != 0:
        raise ValueError("modea_score: bad argument")
    return dict(x=xy[:, 0].copy(), y=xy[:, 1].copy(), ncc=ncc, vis=vis.astype(bool), count=count, avg=avg)
