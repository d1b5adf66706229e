"""The launcher runs the reference's unmodified main.py with the shim registered as ``MVS2``.
Only possible where the reference is mounted (the build container, which has no GPU): SfM
runs on the CPU as in the reference, and the MVS stage must then stop at the device boundary
with the loud no-fallback error -- never silently compute on the CPU."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_shim_module_surface():
    from mvs_b200 import MVS2
    for name in ["MyPatchHeapSort", "MyMatch", "ctNcc", "MyPatch", "CellTable", "DensePointsWithMVS2",
                 "is_patch_neighbor", "ray_plane_intersection", "patch_expansion"]:
        assert hasattr(MVS2, name) and name in MVS2.__all__
    import inspect
    assert list(inspect.signature(MVS2.MyPatch.__init__).parameters) == [
        "self", "centroid", "normal", "reference_img_index", "visible_set", "color", "dist", "patch_size"]
    assert list(inspect.signature(MVS2.MyPatch.photo_consistenecy_test).parameters) == [
        "self", "imgs", "par_K", "par_r", "par_t", "MIN_NCC"]
    assert list(inspect.signature(MVS2.patch_expansion).parameters) == [
        "args", "imgs", "initial_patches", "cells", "camera_pos", "visible_lower_bound"]
    assert list(inspect.signature(MVS2.DensePointsWithMVS2).parameters) == ["imgs", "global_set", "args"]


def test_celltable_matches_reference_semantics():
    import numpy as np
    from mvs_b200 import MVS2
    imgs = [np.zeros((480, 640, 3), np.uint8)] * 2
    ct = MVS2.CellTable(imgs, cell_size=2)
    assert ct.table[0].shape == (320, 240)                       # ceil((W-1)/cs), ceil((H-1)/cs), MVS2.py:88
    assert ct.is_vacant(0, 0, 0) and not ct.is_vacant(0, 320, 0) and not ct.is_vacant(0, 0, -1)
    assert ct.which_cell(11.9, 7.0) == (5, 3)
    assert list(ct.cell_center(5, 3)) == [11.0, 7.0]
    p = MVS2.MyPatch(np.zeros(3), np.zeros(3), 0, [[1, 11.9, 7.0], [0, 11.9, 7.0]], np.zeros(3), None)
    ct.fill_with_point(1, 11.9, 7.0, p)
    assert not ct.is_vacant(1, 5, 3) and ct.is_vacant(0, 5, 3)
    assert len(ct.Q_table[(1, 5, 3)]) == 2                       # appended once per V entry (MVS2.py:106-107)
    pts, cols = ct.reconstruct_from_Q()
    assert len(pts) == 1


def test_lazy_celltable_reconstructs_in_the_reference_scan_order(golden):
    """The device expansion hands its patches over as record ARRAYS; CellTable.reconstruct_from_Q must return them
    in exactly the order the reference's scan (MVS2.py:159-173: view, x-cell, y-cell, list position; every distinct
    patch once) produces on a fully materialised Q_table -- without building any MyPatch object."""
    import numpy as np
    from mvs_b200 import MVS2, records
    e = golden("dino12_expansion")
    V = e["vis"].shape[1]
    imgs = [np.zeros((240, 320, 3), np.uint8) + v for v in range(V)]
    ns = int(e["n_seeds"])
    keep = e["vis"].sum(1) > 0
    idx = np.nonzero(keep)[0]
    seeds, rest = idx[idx < ns], idx[idx >= ns]

    def build():
        ct = MVS2.CellTable(imgs, cell_size=2)
        for k in seeds:                                            # seeds arrive as objects through fill_with_point
            p = MVS2.MyPatch(e["c"][k], e["n"][k], int(e["ref"][k]),
                             [[int(v), float(e["xy"][k, 0]), float(e["xy"][k, 1])] for v in np.nonzero(e["vis"][k])[0]],
                             np.array([1, 2, 3]), None)
            for hit in p.V:
                ct.fill_with_point(hit[0], hit[1], hit[2], p)
        recs = records.make_records(V, e["c"][rest], e["n"][rest], e["xy"][rest], e["avg"][rest], e["ref"][rest], e["vis"][rest],
                                    px=np.stack([np.clip(e["xy"][rest, 0].astype(np.int32), 0, 319),
                                                 np.clip(e["xy"][rest, 1].astype(np.int32), 0, 239)], 1))
        half = len(recs) // 2                                      # two "rounds"
        for part in (recs[:half], recs[half:]):
            ct._add_records(part, MVS2._record_colors(part, imgs))
        return ct
    lazy = build()
    pts_a, col_a = lazy.reconstruct_from_Q()
    assert len(lazy._pending) == 2                                 # nothing was materialised
    eager = build()
    q = eager.Q_table                                              # materialises MyPatch objects
    assert not eager._pending and sum(len(v) for v in q.values()) > 0
    pts_b, col_b = eager.reconstruct_from_Q()
    assert len(pts_a) == len(pts_b) == len(seeds) + len(rest)
    assert np.array_equal(np.array(pts_a), np.array(pts_b))
    assert np.array_equal(np.array(col_a), np.array(col_b))
    # and the reference's own scan on the materialised table gives the same list
    seen, want = set(), []
    for v in range(V):
        t = eager.table[v]
        for i in range(t.shape[0]):
            for j in range(t.shape[1]):
                for p in q.get((v, i, j), []):
                    if id(p) not in seen:
                        seen.add(id(p))
                        want.append(p.c)
    assert np.array_equal(np.array(want), np.array(pts_b))


def test_context_fingerprint_sees_pixel_edits():
    import numpy as np
    from mvs_b200 import MVS2
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 255, (48, 64, 3)).astype(np.uint8) for _ in range(3)]
    a = MVS2._fingerprint(imgs)
    assert a == MVS2._fingerprint([im.copy() for im in imgs])
    imgs[1][0, 5, 1] ^= 1                                          # first row: always sampled
    assert MVS2._fingerprint(imgs) != a


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_launcher_runs_reference_main_up_to_the_device_boundary(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the end-to-end run is covered by the gpu tier")
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, "-m", "mvs_b200.launcher", "--reference-root", REF, "--", "-img_p",
                          os.path.join(REF, "dinoRing/"), "-par_p", os.path.join(REF, "dinoRing/dinoR_par.txt"),
                          "-t", "png", "-scale", "10"], cwd=tmp_path, env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode != 0
    assert "read images from" in res.stdout                      # main.py:9 ran
    assert "MvsError" in res.stderr and "no CPU fallback" in res.stderr
    assert not (tmp_path / "all_patches.ply").exists()
