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
           "ray_plane_intersection", "patch_expansion"]        # (invalidate_context is this module's own helper, not exported by *)


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


_CTX = {"key": None, "ctx": None, "imgs": None, "pars": None}


def _fingerprint(imgs):
    """Cheap content fingerprint of the image list: a strided sample of every image plus the first and last
    rows.  Part of the context key, so that a NEW list that happens to reuse the id() of a freed one, or
    pixels edited in place, do not silently score against a stale GPU stack."""
    import zlib
    h = 0
    for im in imgs:
        a = np.asarray(im)
        flat = a.reshape(-1)
        h = zlib.crc32(np.ascontiguousarray(flat[:: max(1, flat.size // 4096)]).tobytes(), h)
        h = zlib.crc32(np.ascontiguousarray(a[0]).tobytes(), h)
        h = zlib.crc32(np.ascontiguousarray(a[-1]).tobytes(), h)
    return h


def invalidate_context():
    """Drop the cached GPU context (call after editing the pixels of a cached image list IN PLACE)."""
    if _CTX["ctx"] is not None:
        _CTX["ctx"].close()
    _CTX.update(key=None, ctx=None, imgs=None, pars=None)


def _context(imgs, par_K, par_r, par_t):
    """The image stack and cameras live on the GPU once per (image list, cameras).
    * The cache keeps a STRONG reference to the list, so its id() cannot be recycled by a new list while cached.
    * A list object seen for the first time is keyed on a content fingerprint as well (a different list with
      equal shape and cameras never reuses stale pixels).
    * The same list object again takes the fast path (identity + cameras; 70 us per call instead of 0.9 ms for
      re-fingerprinting 48 images): pixels edited IN PLACE are then not noticed -- call invalidate_context(), or
      set MVS_CTX_STRICT=1 to fingerprint every pixel on every call."""
    V = len(imgs)
    strict = bool(os.environ.get("MVS_CTX_STRICT"))
    pars = (par_K, par_r, par_t)
    if (not strict and _CTX["ctx"] is not None and imgs is _CTX["imgs"] and _CTX.get("pars") is not None and
            all(a is b for a, b in zip(pars, _CTX["pars"]))):
        return _CTX["ctx"]                                   # same list object, same parameter dicts
    K, R, t = _stack_pars(par_K, par_r, par_t, V)
    if strict:
        import zlib
        fp = 0
        for im in imgs:
            fp = zlib.crc32(np.ascontiguousarray(im).tobytes(), fp)
    else:
        fp = _fingerprint(imgs)
    key = (V, imgs[0].shape, fp, K.tobytes(), R.tobytes(), t.tobytes())
    if _CTX["key"] != key:
        if _CTX["ctx"] is not None:
            _CTX["ctx"].close()
        device = int(os.environ.get("MVS_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        _CTX["ctx"] = MvsContext(imgs, K, R, t, Rrt=_roundtrip(R), device=device)
        _CTX["key"] = key
    _CTX["imgs"] = imgs
    _CTX["pars"] = pars
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
    """MVS2.py:80-173: per-view vacancy grids (host mirror; the expansion keeps its own copy in HBM and
    writes it back here) and the cell -> patches map.  The patches the device expansion accepts arrive as
    record ARRAYS; ``Q_table`` materialises them into MyPatch objects only when somebody reads it, and
    ``reconstruct_from_Q`` / ``filter_out_outlier`` work on the arrays directly."""

    def __init__(self, imgs, cell_size=4.0):
        self.table = []
        self._Q = defaultdict(list)
        self._registry = {}                 # id(patch) -> patch, in first-insertion order (fill_with_point)
        self._pending = []                  # [(records, colors)] accepted by the device, not yet objects
        self._imgs = imgs
        self._backend = None                # set by patch_expansion: callable -> DeviceBackend (for the filter)
        self.cell_size = cell_size
        for img in imgs:
            row, col = img.shape[0], img.shape[1]
            self.table.append(np.ones((math.ceil((col - 1) / cell_size), math.ceil((row - 1) / cell_size)), dtype=bool))

    # -- the reference's attribute ---------------------------------------------------------------
    @property
    def Q_table(self):
        self._materialise()
        return self._Q

    @Q_table.setter
    def Q_table(self, value):
        self._pending = []
        self._Q = value

    def _materialise(self):
        for recs, colors in self._pending:
            for p in _records_to_patches(recs, self._imgs, len(self.table), colors=colors):
                self._registry[id(p)] = p
                for hit in p.V:          # Q_table as MVS2.py:401-402 leaves it (fill_with_point once per hit)
                    self._Q[(hit[0], math.floor(hit[1] / self.cell_size), math.floor(hit[2] / self.cell_size))].extend(
                        [p] * len(p.V))
        self._pending = []

    def _add_records(self, recs, colors):
        """Patches accepted by the device expansion (records in commit order)."""
        if len(recs):
            self._pending.append((recs, colors))

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
        self._materialise()
        self._registry.setdefault(id(patch), patch)
        for idx, l_col, l_row in patch.V:          # sic: keyed by img_id for every entry (MVS2.py:106-107)
            self._Q[(img_id, math.floor(l_col / self.cell_size), math.floor(l_row / self.cell_size))].append(patch)

    def show_table_non_zeros(self):
        for i, t in enumerate(self.table):
            print("img ", i, " size:", t.shape, " has seen:", t.shape[0] * t.shape[1] - np.count_nonzero(t))

    def which_cell(self, col, row):
        return math.floor(col / self.cell_size), math.floor(row / self.cell_size)

    def cell_center(self, cell_i, cell_j):
        return np.array([self.cell_size * (cell_i + 0.5), self.cell_size * (cell_j + 0.5)])

    def get_color(self, img, col, row):
        return img[int(row)][int(col)]

    def _object_first_keys(self):
        """Distinct MyPatch objects of the materialised part of Q with the first (view, ci, cj, position) at
        which the reference's scan (MVS2.py:159-173) meets them."""
        first = {}
        for key in sorted(self._Q):
            for pos, p in enumerate(self._Q[key]):
                if id(p) not in first:
                    first[id(p)] = (key[0], key[1], key[2], pos, p)
        return list(first.values())

    def reconstruct_from_Q(self):
        """MVS2.py:159-173: every distinct patch once, in table-scan order (first key = lowest view of its
        visible set; inside one list, insertion order).  Array path: no MyPatch objects are created for the
        device-accepted patches."""
        objs = self._object_first_keys()
        if not self._pending:
            return [o[4].c for o in objs], [o[4].color for o in objs]
        cs = self.cell_size
        V = len(self.table)
        kv = [o[0] for o in objs]; ki = [o[1] for o in objs]; kj = [o[2] for o in objs]; ko = [o[3] for o in objs]
        pts = [np.asarray(o[4].c, dtype=np.float64).reshape(3) for o in objs]
        cols = [np.asarray(o[4].color).reshape(3) for o in objs]
        P = [np.array(pts).reshape(-1, 3)]
        Ccol = [np.array(cols).reshape(-1, 3)]
        KV, KI, KJ, KO = [np.array(kv, np.int64)], [np.array(ki, np.int64)], [np.array(kj, np.int64)], [np.array(ko, np.int64)]
        base = 1 << 40                                        # device patches come after every object of a list
        for recs, colors in self._pending:
            vis = unpack_vis(recs["vis"], V)
            seen = vis.any(1)
            KV.append(np.where(seen, vis.argmax(1), V).astype(np.int64)[seen])
            KI.append(np.floor(recs["xy"][:, 0] / cs).astype(np.int64)[seen])
            KJ.append(np.floor(recs["xy"][:, 1] / cs).astype(np.int64)[seen])
            KO.append((base + np.arange(len(recs), dtype=np.int64))[seen])
            base += len(recs)
            P.append(np.ascontiguousarray(recs["c"])[seen])
            Ccol.append(np.asarray(colors)[seen])
        KV, KI, KJ, KO = (np.concatenate(x) for x in (KV, KI, KJ, KO))
        order = np.lexsort((KO, KJ, KI, KV))
        return list(np.concatenate(P)[order]), list(np.concatenate(Ccol)[order])

    def _all_records(self):
        """Every patch known to the table as records, in INSERTION order (objects first -- they were filled
        before the device expansion ran -- then the device-accepted arrays), with a back-reference."""
        V = len(self.table)
        objs = list(self._registry.values())
        parts = [_patches_to_records(objs, V)] if objs else []
        parts += [r for r, _ in self._pending]
        from .records import rec_dtype
        return objs, (np.concatenate(parts) if parts else np.zeros(0, dtype=rec_dtype(V)))

    def filter_out_outlier(self):
        """MVS2.py:132-158 on the device (mvs_cells_filter): removes every outlier patch from all its Q
        lists.  The reference leaves this pass disabled (MVS2.py:280-281); DensePointsWithMVS2 runs it when
        MVS_FILTER_OUTLIERS=1."""
        if self._backend is None:
            raise MvsError("CellTable.filter_out_outlier needs the device context of this table: it is available after "
                           "patch_expansion / DensePointsWithMVS2 ran on it (there is no CPU path)")
        objs, recs = self._all_records()
        be = self._backend(np.stack([np.asarray(t, dtype=np.uint8) for t in self.table]))
        removed, n_removed, n_empty = be.filter(recs)
        gone = set(id(o) for o, r in zip(objs, removed[:len(objs)]) if r)
        if gone:
            for key in list(self._Q):
                self._Q[key] = [q for q in self._Q[key] if id(q) not in gone]
            for i in gone:
                self._registry.pop(i, None)
        off = len(objs)
        kept = []
        for r, col in self._pending:
            m = ~removed[off:off + len(r)]
            off += len(r)
            kept.append((r[m], np.asarray(col)[m]))
        self._pending = [(r, c) for r, c in kept if len(r)]
        print("filter_out_outlier: removed", n_removed, "patches;", n_empty, "non-vacant cells had no patch left")
        return n_removed


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


def _record_colors(recs, imgs):
    """CellTable.get_color(imgs[ref], u, v) (MVS2.py:116-117,357) for device records: the pixel (px[0], px[1])
    of the reference view, gathered view by view (no per-patch Python)."""
    n = len(recs)
    colors = np.zeros((n, 3), dtype=np.asarray(imgs[0]).dtype)
    if n == 0:
        return colors
    px = recs["px"]
    ref = np.asarray(recs["ref"])
    has = px[:, 0] >= 0
    for v in np.unique(ref[has]):
        m = has & (ref == v)
        colors[m] = np.asarray(imgs[int(v)])[px[m, 1], px[m, 0]]
    return colors


def _records_to_patches(recs, imgs, V, colors=None):
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
    if colors is None:
        colors = _record_colors(recs, imgs)
    has_px = recs["px"][:, 0] >= 0
    out = []
    for k in range(n):
        x, y = xy[k]
        p = MyPatch(c[k].copy(), nrm[k].copy(), ref[k], [[v, x, y] for v in views[starts[k]:starts[k + 1]]],
                    colors[k] if has_px[k] else np.zeros(3), None)
        p.avg_ncc_score = avg[k]
        out.append(p)
    return out


def _dist_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


def patch_expansion(args, imgs, initial_patches, cells, camera_pos, visible_lower_bound):
    """MVS2.py:308-404 restructured into synchronous device rounds.  Mutates ``cells`` (vacancy grids +
    Q_table) like the reference.  Default: the whole loop runs inside ONE C call (mvs_expand_run: one host
    synchronisation per round; with several GPUs the accept decisions travel as a minimal wire over NVLink
    stores, one device-side barrier per round).  MVS_ROUNDS=stepwise selects the per-phase driver
    (rounds.RoundDriver: mvs_round_generate / score / commit, all-gather through torch.distributed)."""
    from .rounds import DeviceBackend, RoundDriver
    par_K, par_r, par_t = read_pars(args)
    V = len(imgs)
    ctx = _context(imgs, par_K, par_r, par_t)
    if float(cells.cell_size) != int(cells.cell_size) or int(cells.cell_size) < 1:
        raise MvsError("patch_expansion: the device cell table needs an integer cell_size >= 1 (got %r)" % (cells.cell_size,))
    table = np.stack([np.asarray(t, dtype=np.uint8) for t in cells.table])

    def backend(tab):
        return DeviceBackend(ctx, cell_size=int(cells.cell_size), scale=float(args.scale), bound=int(visible_lower_bound),
                             min_ncc=0.7, wid=5, table=tab)

    be = backend(table)
    cells._backend = backend
    rank, world, dist = _dist_world()
    max_rounds = int(os.environ.get("MVS_MAX_ROUNDS", "100000"))
    max_iter = int(os.environ.get("MVS_MAX_ITERATIONS", "100000"))        # the reference's cap: iteration < 100000 (MVS2.py:321)
    max_patches = os.environ.get("MVS_MAX_PATCHES")
    max_patches = int(max_patches) if max_patches else None
    timing = bool(os.environ.get("MVS_TIME_ROUNDS"))
    seeds = be.to_device(_patches_to_records(initial_patches, V))
    mode = os.environ.get("MVS_ROUNDS", "fused")
    if mode == "fused" and world > 1:
        # the fused exchange needs NCCL symmetric memory with peer access on ONE box: agree on it collectively,
        # otherwise every rank falls back to the stepwise driver with a collective all-gather
        from .rounds import agree_on_fused_exchange
        cap = int(os.environ.get("MVS_ROUND_CAPACITY", str(1 << 21)))
        # MVS_EXCHANGE_PARTS > 1: rounds whose shard reaches 2^17 candidates publish in position ranges under the scoring
        # (mvs_exchange_set_parts).  Off by default: measured on 8 GPUs it buys 2 % at best (exchange section)
        parts = int(os.environ.get("MVS_EXCHANGE_PARTS", "1"))
        ok, why = agree_on_fused_exchange(lambda: be.exchange_setup(cap, world, dist.group.WORLD, parts=parts), dist, be.device)
        if not ok:
            if rank == 0:
                print("patch_expansion: fused P2P exchange unavailable (%s); using the collective all-gather" % why)
            mode = "stepwise"
    if mode == "fused":
        stats, n_new = be.expand_run(seeds, max_rounds=max_rounds, max_iterations=max_iter, max_patches=max_patches,
                                     rank=rank, world=world, timing=timing)
        recs = be.expand_result(0, n_new)
    else:
        exchange = os.environ.get("MVS_EXCHANGE", "collective")
        drv = RoundDriver(be, rank=rank, world=world, group=None, timing=timing, exchange=exchange)
        accepted = drv.run(seeds, max_rounds=max_rounds, max_patches=max_patches, max_iterations=max_iter)
        drv.finish_timing()
        stats = drv.stats
        from .records import rec_dtype
        recs = np.concatenate([be.to_host(r) for r in accepted]) if accepted else np.zeros(0, dtype=rec_dtype(V))
    new_tab = be.table()
    for v in range(V):
        cells.table[v][...] = new_tab[v]
    cells._add_records(recs, _record_colors(recs, imgs))
    if getattr(args, "debug", False):
        for i, st in enumerate(stats):
            print("round", i + 1, st)
    print("expansion rounds:", len(stats), "new patches:", len(recs))
    patch_expansion.last_stats = stats
    patch_expansion.last_records = recs


def _flatten_tracks(legal_sets, n_cameras):
    """SfM tracks (GlobalSet.getInfo(): legal_set.point2d_list = [(view, x, y), ...]) as a flat observation
    array + offsets, the input layout of mvs_seed_stage."""
    obs, offsets = [], [0]
    for legal_set in legal_sets:
        for cam, px, py in legal_set.point2d_list:
            if cam >= n_cameras:
                raise IndexError("track observation refers to view %d of %d" % (cam, n_cameras))
            obs.append((float(cam), float(px), float(py)))
        offsets.append(len(obs))
    return np.array(obs, dtype=np.float64).reshape(-1, 3), np.array(offsets, dtype=np.int64)


def DensePointsWithMVS2(imgs, global_set, args):
    """MVS2.py:176-295: seed patches from the SfM tracks, expansion, point cloud export.  The seed stage
    (triangulation of every track candidate, scoring at MIN_NCC 0.4, nearest-first pick) runs on the device
    in one call (mvs_seed_stage); MVS_FILTER_OUTLIERS=1 also runs CellTable.filter_out_outlier, which the
    reference ships disabled (MVS2.py:280-281)."""
    from .rounds import DeviceBackend
    t0 = time.time()
    par_K, par_r, par_t = read_pars(args)
    n_observations, n_world_points, legal_sets = global_set.getInfo()
    n_cameras = len(imgs)
    camera_pos = [-(par_r[i].transpose() @ par_t[i].reshape(3, -1)).reshape(-1) for i in range(n_cameras)]
    cells = CellTable(imgs, cell_size=args.cell_size)
    visible_lower_bound = 3 if n_cameras > 2 else 2
    _, _, dist_mod = _dist_world()
    rank = dist_mod.get_rank() if dist_mod is not None else 0

    obs, offsets = _flatten_tracks(legal_sets, n_cameras)
    initial_patches = []
    if len(obs):
        ctx = _context(imgs, par_K, par_r, par_t)
        cs = int(args.cell_size) if float(args.cell_size) == int(args.cell_size) and args.cell_size >= 1 else 1
        be = DeviceBackend(ctx, cell_size=cs, scale=float(args.scale), bound=visible_lower_bound, min_ncc=0.7, wid=5)
        P = np.stack([par_K[v] @ np.concatenate((par_r[v], par_t[v].reshape(3, 1)), axis=1) for v in range(n_cameras)])   # utils.py:234-236
        recs = be.seed_stage(obs, offsets, P, min_ncc=0.4)
        # candidate number -> observation it was triangulated with (every observation but the first of its track)
        first = np.zeros(len(obs), dtype=bool)
        lens = np.diff(offsets)
        first[offsets[:-1][lens > 0]] = True
        cand_obs = np.nonzero(~first)[0]
        vis = unpack_vis(recs["vis"], n_cameras)
        for k in range(len(recs)):
            o = obs[cand_obs[int(recs["index"][k])]]
            c = np.array(recs["c"][k])
            ref = int(recs["ref"][k])
            x, y = float(recs["xy"][k, 0]), float(recs["xy"][k, 1])
            O = camera_pos[ref]
            p = MyPatch(c, np.array(recs["n"][k]), ref, [[int(v), x, y] for v in np.nonzero(vis[k])[0]],
                        cells.get_color(imgs[int(o[0])], o[1], o[2]), math.sqrt(((c - O) ** 2).sum()))
            p.avg_ncc_score = float(recs["avg"][k])
            initial_patches.append(p)
            for hit in p.V:
                cells.fill_with_point(hit[0], hit[1], hit[2], p)
    print("len of initial patches", len(initial_patches))
    if rank == 0:
        export2ply(np.array([p.c for p in initial_patches]).reshape(-1, 3),
                   np.array([p.color for p in initial_patches]).reshape(-1, 3), path="initial_patches")

    patch_expansion(args, imgs, initial_patches, cells, camera_pos, visible_lower_bound)

    if os.environ.get("MVS_FILTER_OUTLIERS"):
        print("filter outliers")
        cells.filter_out_outlier()
    print("reconstruct point cloud")
    points_3d, colors = cells.reconstruct_from_Q()
    print("Optimization took {0:.0f} seconds".format(time.time() - t0))
    print("points len:", len(points_3d))
    if rank == 0:
        export2ply(np.array(points_3d).reshape(-1, 3), np.array(colors).reshape(-1, 3), path="all_patches")
