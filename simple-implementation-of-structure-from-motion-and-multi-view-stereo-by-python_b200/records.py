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


def _align(x, a):
    return (x + a - 1) // a * a


def wire_geometry(V, capacity, parts=1):
    """Byte geometry of one source rank's region of the minimal wire (csrc/exchange.cu::wire_layout) for an inbox
    sized for ``capacity`` candidates per region with mvs_exchange_set_parts(``parts``)."""
    mw = mask_words(V)
    wb = 8 + 8 * mw
    ent_off = _align(16 + 8 * ((capacity + 31) // 32), 256)
    region = _align(ent_off + capacity * wb, 256)
    parts = max(int(parts), 1)
    part_cap = _align((capacity + parts - 1) // parts, 1024)
    sub = _align(ent_off + part_cap * wb, 256)
    if parts > 1:
        region = max(region, parts * sub)
    return dict(mw=mw, wb=wb, ent_off=ent_off, region_bytes=region, sub_bytes=sub, part_cap=part_cap)


def decode_wire(region, V, capacity, parts=1):
    """Decode one region (uint8 array) of the minimal wire, in either layout (the header names it): returns
    dict(n, used, kept [used], passed [n] bool, avg [n] f64 (NaN where not passed), vis [n, mw] u64, part [n] int
    (-1 where not passed)).  Raises AssertionError on an inconsistent wire (prefixes, overlapping parts)."""
    g = wire_geometry(V, capacity, parts)
    region = np.ascontiguousarray(region, dtype=np.uint8)
    hdr = region[:16].view("<i8")
    n = int(hdr[1] & ((1 << 48) - 1))
    used = max(int(hdr[1] >> 48), 1)
    nw = (n + 31) // 32
    passed = np.zeros(n, dtype=bool)
    avg = np.full(n, np.nan)
    vis = np.zeros((n, g["mw"]), dtype=np.uint64)
    part = np.full(n, -1, dtype=np.int32)
    kept = []
    for k in range(used):
        sub = region[k * g["sub_bytes"]:] if used > 1 else region
        h = sub[:16].view("<i8")
        assert int(h[1] & ((1 << 48) - 1)) == n and max(int(h[1] >> 48), 1) == used
        kk = int(h[0])
        words = sub[16:16 + 8 * nw].view("<u4").reshape(-1, 2)
        bits = np.unpackbits(words[:, 0].copy().view(np.uint8), bitorder="little")[:n].astype(bool)
        per_word = np.add.reduceat(bits.astype(np.int64), np.arange(0, n, 32)) if n else np.zeros(0, np.int64)
        assert np.array_equal(words[:, 1].astype(np.int64), np.cumsum(per_word) - per_word), "prefix words"
        assert int(bits.sum()) == kk, "header count"
        assert not (passed & bits).any(), "a candidate passed in two parts"
        ent = sub[g["ent_off"]: g["ent_off"] + g["wb"] * kk].view("<u8").reshape(kk, 1 + g["mw"])
        passed |= bits
        avg[bits] = ent[:, 0].copy().view("<f8")
        vis[bits] = ent[:, 1:]
        part[bits] = k
        kept.append(kk)
    return dict(n=n, used=used, kept=kept, passed=passed, avg=avg, vis=vis, part=part)
