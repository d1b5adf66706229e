"""Patch records: the binary layout of ``mvs_patch_record`` (include/mvs_ncc.h) as a NumPy
structured dtype -- the fields of the reference's MyPatch (MVS2.py:45-60) that cross the
C ABI and the GPU-to-GPU exchange."""
import numpy as np


def mask_words(V):
    return (V + 63) // 64


def rec_dtype(V):
    return np.dtype([("c", "<f8", 3), ("n", "<f8", 3), ("xy", "<f8", 2), ("avg", "<f8"), ("ref", "<i4"),
                     ("count", "<i4"), ("index", "<i8"), ("px", "<i4", 2), ("vis", "<u8", mask_words(V))])


def pack_vis(vis_bool):
    """[N,V] bool -> [N, ceil(V/64)] uint64 (bit v of the row = view v)."""
    vis_bool = np.asarray(vis_bool, dtype=bool)
    N, V = vis_bool.shape
    mw = mask_words(V)
    padded = np.zeros((N, mw * 64), dtype=np.uint8)
    padded[:, :V] = vis_bool
    return np.packbits(padded, axis=1, bitorder="little").view("<u8").reshape(N, mw)


def unpack_vis(mask, V):
    m = np.ascontiguousarray(mask, dtype=np.uint64).reshape(len(mask), -1)
    return np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")[:, :V].astype(bool)


def make_records(V, c, n, xy, avg, ref, vis_bool, index=None, px=None):
    N = len(c)
    r = np.zeros(N, dtype=rec_dtype(V))
    r["c"], r["n"], r["xy"], r["avg"], r["ref"] = c, n, xy, avg, ref
    r["vis"] = pack_vis(np.asarray(vis_bool).reshape(N, V))
    r["count"] = np.asarray(vis_bool).reshape(N, V).sum(1)
    r["index"] = np.arange(N) if index is None else index
    r["px"] = -1 if px is None else px
    return r
