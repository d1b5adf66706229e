"""Drop-in replacement for the reference's ``MVS2`` module (the only module main.py needs
from the MVS stage: ``from MVS2 import *`` at main.py:5, ``DensePointsWithMVS2`` at
main.py:30).  Same names, argument meaning and return values as MVS2.py; the scoring
(MVS2.py:62-77) and the expansion loop (MVS2.py:308-404) run on a B200 through
libmvsncc.so.  main.py, SFM.py, BundleAdjustment.py, GlobalSet.py and utils.py run
unchanged (see launcher.py).

Differences from the reference, all declared in DESIGN.md:
  * patch_expansion runs in synchronous rounds (every accepted patch is expanded once,
    against the round-start cell table) instead of the sequential FIFO;
  * nothing is plotted; the two PLY files are written by ply.export2ply (pyntcloud-free);
  * there is no CPU path: without the library or a B200 every scoring call raises.
"""
import heapq
import math
import os
import time
from collections import defaultdict

import numpy as np

from . import _lib
from .context import MvsContext, MvsError, _check
from .ply import export2ply
from .records import make_records, unpack_vis

__all__ = ["MyPatchHeapSort", "MyMatch", "ctNcc", "MyPatch", "CellTable", "DensePointsWithMVS2", "is_patch_neighbor",
           "ray_plane_intersection", "patch_expansion"]


# ---------------------------------------------------------------------------------------
# camera file (utils.py:56-81) and the Rodrigues round trip the reference applies implicitly
# ---------------------------------------------------------------------------------------
def read_pars(args):
    """Same contract as utils.read_pars: (dict K 3x3, dict R 3x3, dict t 3x1), index = line - 1."""
    par_K, par_r, par_t = {}, {}, {}
    with open(args.par_path, "r") as f:
        for i, line in enumerate(f.readlines()):
            if i == 0 or not line.strip():
                continue
            v = [float(x) for x in line.split()[1:]]
            par_K[i - 1] = np.array(v[0:9]).reshape(3, 3)
            par_r[i - 1] = np.array(v[9:18]).reshape(3, 3)
            par_t[i - 1] = np.array(v[18:21]).reshape(3, 1)
    return par_K, par_r, par_t


def _stack_pars(par_K, par_r, par_t, V):
    K = np.stack([np.asarray(par_K[i], dtype=np.float64).reshape(3, 3) for i in range(V)])
    R = np.stack([np.asarray(par_r[i], dtype=np.float64).reshape(3, 3) for i in range(V)])
    t = np.stack([np.asarray(par_t[i], dtype=np.float64).reshape(3) for i in range(V)])
    return K, R, t


def _roundtrip(R):
    """cv2.Rodrigues(cv2.Rodrigues(R)) when cv2 is importable (bit-identical projections to
    utils.py:241-244); otherwise None and the library computes it (agrees to ~1e-14)."""
    try:
        import cv2
    except Exception:
        return None
    return np.stack([cv2.Rodrigues(cv2.Rodrigues(r)[0])[0] for r in R])


_CTX = {"key": None, "ctx": None}


def _context(imgs, par_K, par_r, par_t):
    """The image stack and cameras live on the GPU once per (image list, cameras)."""
    V = len(imgs)
    K, R, t = _stack_pars(par_K, par_r, par_t, V)
    key = (id(imgs), V, imgs[0].shape, K.tobytes(), R.tobytes(), t.tobytes())
    if _CTX["key"] != key:
        if _CTX["ctx"] is not None:
            _CTX["ctx"].close()
        device = int(os.environ.get("MVS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _CTX["ctx"] = MvsContext(imgs, K, R, t, Rrt=_roundtrip(R), device=device)
        _CTX["key"] = key
    return _CTX["ctx"]


# ---------------------------------------------------------------------------------------
# small types of the reference's module surface
# ---------------------------------------------------------------------------------------
class MyPatchHeapSort(object):
    """MVS2.py:13-31: min-heap of patches keyed (dist, c0, c1, c2, R)."""

    def __init__(self, initial=None, key=lambda x: (x.dist, x.c[0], x.c[1], x.c[2], x.R)):
        self.key = key
        self._data = [(key(item), item) for item in initial] if initial else []
        heapq.heapify(self._data)

    def push(self, item):
        heapq.heappush(self._data, (self.key(item), item))

    def pop(self):
        return heapq.heappop(self._data)[1]

    def size(self):
        return len(self._data)


class MyMatch(object):
    """MVS2.py:33-37."""

    def __init__(self, src_point, dst_point, ncc_score):
        self.ncc_score = ncc_score
        self.src_point = src_point
        self.dst_point = dst_point


def ctNcc(desc1, desc2):
    """MVS2.py:39-43 for two uint8 descriptors, evaluated on the device (mvs_ncc_pairs)."""
    import ctypes as C
    a = np.ascontiguousarray(np.asarray(desc1).reshape(-1))
    b = np.ascontiguousarray(np.asarray(desc2).reshape(-1))
    if a.dtype != np.uint8 or b.dtype != np.uint8 or a.shape != b.shape:
        raise MvsError("ctNcc: descriptors must be uint8 arrays of equal length (getDescFeatures windows)")
    out = np.zeros(1)
    device = int(os.environ.get("MVS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    _check(_lib.load().mvs_ncc_pairs(device, 1, a.shape[0], C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data),
                                     C.c_void_p(out.ctypes.data), 0, None), "mvs_ncc_pairs")
    return float(out[0])


class MyPatch(object):
    """MVS2.py:45-77."""

    def __init__(self, centroid, normal, reference_img_index, visible_set, color, dist, patch_size=5):
        self.dist = dist
        self.c = centroid
        self.n = normal
        self.R = reference_img_index
        self.V = visible_set if visible_set is not None else []
        self.color = color
        self.patch_size = patch_size
        self.avg_ncc_score = 0

    def visible_ct(self):
        return len(self.V)

    def photo_consistenecy_test(self, imgs, par_K, par_r, par_t, MIN_NCC=0.7):
        """One hypothesis through the batched device scorer.  Appends [view, x, y] for every
        view with ncc > MIN_NCC and updates avg_ncc_score exactly as MVS2.py:72-76 does."""
        ctx = _context(imgs, par_K, par_r, par_t)
        out = ctx.score_host(np.asarray(self.c, dtype=np.float64).reshape(1, 3), [int(self.R)], min_ncc=MIN_NCC, wid=5)
        x, y = float(out["xy"][0, 0]), float(out["xy"][0, 1])
        views = np.nonzero(unpack_vis(out["vis_mask"], ctx.V)[0])[0]
        self.avg_ncc_score += float(out["avg"][0]) * len(views)
        for v in views:
            self.V.append([int(v), x, y])
        if self.visible_ct() > 0:
            self.avg_ncc_score /= self.visible_ct()
        return self.V


class CellTable(object):
    """MVS2.py:80-173: per-view vacancy grids (host mirror; the expansion keeps its own copy
    in HBM and writes it back here) and the cell -> patches map."""

    def __init__(self, imgs, cell_size=4.0):
        self.table = []
        self.Q_table = defaultdict(list)
        self.cell_size = cell_size
        for img in imgs:
            row, col = img.shape[0], img.shape[1]
            self.table.append(np.ones((math.ceil((col - 1) / cell_size), math.ceil((row - 1) / cell_size)), dtype=bool))

    def is_vacant(self, img_id, cell_i, cell_j):
        t = self.table[img_id]
        if cell_i >= t.shape[0] or cell_i < 0 or cell_j >= t.shape[1] or cell_j < 0:
            return False
        return t[cell_i][cell_j]

    def fill_with_point(self, img_id, col, row, patch):
        ci, cj = math.floor(col / self.cell_size), math.floor(row / self.cell_size)
        t = self.table[img_id]
        if ci >= t.shape[0] or col < 0 or cj >= t.shape[1] or row < 0:
            raise IndexError("CellTable.fill_with_point: (%r, %r) is outside view %d" % (col, row, img_id))
        t[ci][cj] = False
        for idx, l_col, l_row in patch.V:          # sic: keyed by img_id for every entry (MVS2.py:106-107)
            self.Q_table[(img_id, math.floor(l_col / self.cell_size), math.floor(l_row / self.cell_size))].append(patch)

    def show_table_non_zeros(self):
        for i, t in enumerate(self.table):
            print("img ", i, " size:", t.shape, " has seen:", t.shape[0] * t.shape[1] - np.count_nonzero(t))

    def which_cell(self, col, row):
        return math.floor(col / self.cell_size), math.floor(row / self.cell_size)

    def cell_center(self, cell_i, cell_j):
        return np.array([self.cell_size * (cell_i + 0.5), self.cell_size * (cell_j + 0.5)])

    def get_color(self, img, col, row):
        return img[int(row)][int(col)]

    def reconstruct_from_Q(self):
        """MVS2.py:159-173: every distinct patch once, in table-scan order."""
        import itertools
        # first occurrence in (view, x-cell, y-cell) scan order; MyPatch hashes by identity
        distinct = dict.fromkeys(itertools.chain.from_iterable(self.Q_table[key] for key in sorted(self.Q_table)))
        return [p.c for p in distinct], [p.color for p in distinct]


def is_patch_neighbor(patch, non_finished_patch, threshold=0.2):
    """MVS2.py:298-299."""
    d = np.asarray(patch.c) - np.asarray(non_finished_patch.c)
    return abs(np.dot(d, patch.n) + np.dot(d, non_finished_patch.n)) < threshold


def ray_plane_intersection(ray_origin, ray_direction, plane_center, plane_normal):
    """MVS2.py:302-306."""
    dot_out = np.dot(ray_direction, plane_normal)
    t = np.dot(plane_center - ray_origin, plane_normal) / dot_out
    return ray_origin + t * ray_direction


# ---------------------------------------------------------------------------------------
# the two callers of the scorer
# ---------------------------------------------------------------------------------------
def _triangulate(P1, P2, x1, x2):
    """utils.py:238-239 (cv2.triangulatePoints) for one correspondence -> homogeneous 4-vector."""
    try:
        import cv2
        return cv2.triangulatePoints(P1, P2, np.array([x1]).transpose(), np.array([x2]).transpose()).transpose()[0]
    except ImportError:
        A = np.stack([x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1], x2[0] * P2[2] - P2[0], x2[1] * P2[2] - P2[1]])
        return np.linalg.svd(A)[2][-1]


def _patches_to_records(patches, V):
    vis = np.zeros((len(patches), V), dtype=bool)
    xy = np.zeros((len(patches), 2))
    for k, p in enumerate(patches):
        for v, x, y in p.V:
            vis[k, int(v)] = True
            xy[k] = (x, y)
    return make_records(V, np.array([p.c for p in patches], dtype=np.float64).reshape(-1, 3),
                        np.array([p.n for p in patches], dtype=np.float64).reshape(-1, 3), xy,
                        np.array([p.avg_ncc_score for p in patches], dtype=np.float64), np.array([p.R for p in patches]),
                        vis)


def _records_to_patches(recs, imgs, V):
    """Device patch records -> MyPatch objects (the reference's carrier type).  Field extraction is
    done array-wise; only the object construction itself is per patch."""
    n = len(recs)
    if n == 0:
        return []
    vis = unpack_vis(recs["vis"], V)
    rows, views = np.nonzero(vis)                              # row-major: ascending view index inside a patch
    starts = np.searchsorted(rows, np.arange(n + 1))
    views = views.tolist()
    c = np.ascontiguousarray(recs["c"])
    nrm = np.ascontiguousarray(recs["n"])
    xy = recs["xy"].tolist()
    ref = recs["ref"].tolist()
    avg = recs["avg"].tolist()
    px = recs["px"]
    has_px = px[:, 0] >= 0
    colors = np.zeros((n, 3), dtype=imgs[0].dtype)
    if has_px.any():
        stack_ref = np.asarray(recs["ref"])[has_px]
        pr, pc = px[has_px, 1], px[has_px, 0]
        colors[has_px] = [imgs[r][y][x] for r, y, x in zip(stack_ref.tolist(), pr.tolist(), pc.tolist())]
    out = []
    for k in range(n):
        x, y = xy[k]
        p = MyPatch(c[k].copy(), nrm[k].copy(), ref[k], [[v, x, y] for v in views[starts[k]:starts[k + 1]]],
                    colors[k] if has_px[k] else np.zeros(3), None)
        p.avg_ncc_score = avg[k]
        out.append(p)
    return out


def patch_expansion(args, imgs, initial_patches, cells, camera_pos, visible_lower_bound):
    """MVS2.py:308-404 restructured into synchronous device rounds (rounds.py).  Mutates
    ``cells`` (vacancy grids + Q_table) like the reference; also returns the new patches."""
    from .rounds import DeviceBackend, RoundDriver
    par_K, par_r, par_t = read_pars(args)
    V = len(imgs)
    ctx = _context(imgs, par_K, par_r, par_t)
    table = np.stack([np.asarray(t, dtype=np.uint8) for t in cells.table])
    be = DeviceBackend(ctx, cell_size=int(cells.cell_size), scale=float(args.scale), bound=int(visible_lower_bound),
                       min_ncc=0.7, wid=5, table=table)
    rank, world, group = 0, 1, None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    # multi-GPU exchange: fused into the compaction over NVLink P2P stores unless MVS_EXCHANGE=collective
    drv = RoundDriver(be, rank=rank, world=world, group=group, timing=bool(os.environ.get("MVS_TIME_ROUNDS")),
                      exchange=os.environ.get("MVS_EXCHANGE", "p2p"))
    max_rounds = int(os.environ.get("MVS_MAX_ROUNDS", "100000"))
    max_iter = int(os.environ.get("MVS_MAX_ITERATIONS", "100000"))        # the reference's cap: iteration < 100000 (MVS2.py:321)
    max_patches = os.environ.get("MVS_MAX_PATCHES")
    accepted = drv.run(be.to_device(_patches_to_records(initial_patches, V)), max_rounds=max_rounds,
                       max_patches=int(max_patches) if max_patches else None, max_iterations=max_iter)
    drv.finish_timing()
    new_tab = be.table()
    for v in range(V):
        cells.table[v][...] = new_tab[v]
    new_patches = []
    for rnd in accepted:
        new_patches.extend(_records_to_patches(be.to_host(rnd), imgs, V))
    for p in new_patches:                                  # Q_table as MVS2.py:401-402 leaves it
        for hit in p.V:
            cells.Q_table[(hit[0], math.floor(hit[1] / cells.cell_size), math.floor(hit[2] / cells.cell_size))].extend(
                [p] * len(p.V))
    if getattr(args, "debug", False):
        for i, st in enumerate(drv.stats):
            print("round", i + 1, st)
    print("expansion rounds:", len(drv.stats), "new patches:", len(new_patches))
    patch_expansion.last_stats = drv.stats
    return new_patches


def DensePointsWithMVS2(imgs, global_set, args):
    """MVS2.py:176-295: seed patches from the SfM tracks, expansion, point cloud export."""
    t0 = time.time()
    par_K, par_r, par_t = read_pars(args)
    n_observations, n_world_points, legal_sets = global_set.getInfo()
    n_cameras = len(imgs)
    camera_pos = [-(par_r[i].transpose() @ par_t[i].reshape(3, -1)).reshape(-1) for i in range(n_cameras)]
    cells = CellTable(imgs, cell_size=args.cell_size)
    visible_lower_bound = 3 if n_cameras > 2 else 2

    # every (track, other observation) candidate of MVS2.py:223-250, scored in ONE batch at MIN_NCC 0.4
    cand, owner = [], []
    for ti, legal_set in enumerate(legal_sets):
        ref, base, O = None, None, None
        for ct, (cam, px, py) in enumerate(legal_set.point2d_list):
            if cam >= n_cameras:
                raise IndexError("track observation refers to view %d of %d" % (cam, n_cameras))
            pt = [float(px), float(py)]
            if ct == 0:
                ref, base, O = cam, pt, camera_pos[cam]
                P1 = par_K[ref] @ np.concatenate((par_r[ref], par_t[ref]), axis=1)
                continue
            P2 = par_K[cam] @ np.concatenate((par_r[cam], par_t[cam]), axis=1)
            un = _triangulate(P1, P2, base, pt)
            c = 0 * un[:-1] if un[-1] == 0 else un[:-1] / un[-1]
            dist = math.sqrt(((c - O) ** 2).sum())
            n = (O - c) / dist
            color = cells.get_color(imgs[cam], px, py)
            cand.append(MyPatch(c, n, ref, None, color, dist))
            owner.append(ti)
    initial_patches = []
    if cand:
        ctx = _context(imgs, par_K, par_r, par_t)
        out = ctx.score_host(np.array([p.c for p in cand]), np.array([p.R for p in cand], dtype=np.int32), min_ncc=0.4, wid=5)
        vis = unpack_vis(out["vis_mask"], ctx.V)
        for k, p in enumerate(cand):
            x, y = float(out["xy"][k, 0]), float(out["xy"][k, 1])
            p.V = [[int(v), x, y] for v in np.nonzero(vis[k])[0]]
            p.avg_ncc_score = float(out["avg"][k])
        by_track = defaultdict(list)
        for p, ti in zip(cand, owner):
            by_track[ti].append(p)
        for ti in sorted(by_track):                          # nearest-first, first with enough views wins (MVS2.py:253-260)
            heap = MyPatchHeapSort(by_track[ti])
            while heap.size() != 0:
                p = heap.pop()
                if p.visible_ct() >= visible_lower_bound:
                    initial_patches.append(p)
                    for hit in p.V:
                        cells.fill_with_point(hit[0], hit[1], hit[2], p)
                    break
    print("len of initial patches", len(initial_patches))
    export2ply(np.array([p.c for p in initial_patches]).reshape(-1, 3),
               np.array([p.color for p in initial_patches]).reshape(-1, 3), path="initial_patches")

    patch_expansion(args, imgs, initial_patches, cells, camera_pos, visible_lower_bound)

    print("reconstruct point cloud")
    points_3d, colors = cells.reconstruct_from_Q()
    print("Optimization took {0:.0f} seconds".format(time.time() - t0))
    print("points len:", len(points_3d))
    export2ply(np.array(points_3d).reshape(-1, 3), np.array(colors).reshape(-1, 3), path="all_patches")
