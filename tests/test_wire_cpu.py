"""Host mirror of the minimal wire (records.wire_geometry / decode_wire) against a reference encoder written from
the format's description in include/mvs_ncc.h: both layouts (single; one sub-region per position range)."""
import numpy as np
import pytest

from mvs_b200 import records


def _encode(V, cap, parts_cfg, passed, avg, vis, part_of=None):
    """Region bytes of the wire for n = len(passed) candidates; part_of[i] = position range of candidate i
    (None: single layout)."""
    g = records.wire_geometry(V, cap, parts_cfg)
    n = len(passed)
    reg = np.full(g["region_bytes"], 0xCD, dtype=np.uint8)            # stale bytes everywhere
    used = 1 if part_of is None else parts_cfg
    for k in range(used):
        sub = reg[k * g["sub_bytes"]:] if used > 1 else reg
        sel = passed if part_of is None else passed & (part_of == k)
        bits = np.zeros(((n + 31) // 32) * 32, dtype=np.uint8)
        bits[:n] = sel
        w = np.packbits(bits, bitorder="little").view("<u4")
        per = bits.reshape(-1, 32).sum(1)
        words = np.stack([w, (np.cumsum(per) - per).astype(np.uint32)], axis=1).astype("<u4")
        sub[:16].view("<i8")[:] = [int(sel.sum()), n | ((used if used > 1 else 0) << 48)]
        sub[16:16 + words.nbytes] = words.reshape(-1).view(np.uint8)
        ent = np.concatenate([avg[sel].view("<u8")[:, None], vis[sel]], axis=1).astype("<u8")
        sub[g["ent_off"]: g["ent_off"] + ent.nbytes] = ent.reshape(-1).view(np.uint8)
    return reg


@pytest.mark.parametrize("V,parts", [(12, 1), (48, 3), (128, 2), (200, 8)])
def test_decode_wire_round_trip(V, parts):
    rng = np.random.default_rng(V)
    cap, n = 9000, 8765
    mw = records.mask_words(V)
    passed = rng.random(n) < 0.4
    avg = rng.random(n)
    vis = rng.integers(1, 1 << 62, (n, mw)).astype(np.uint64)
    part_of = None if parts == 1 else np.sort(rng.integers(0, parts, n))[rng.permutation(n)]
    reg = _encode(V, cap, parts, passed, avg, vis, part_of)
    w = records.decode_wire(reg, V, cap, parts)
    assert w["n"] == n and w["used"] == parts
    assert np.array_equal(w["passed"], passed)
    assert np.array_equal(w["avg"][passed], avg[passed]) and np.isnan(w["avg"][~passed]).all()
    assert np.array_equal(w["vis"][passed], vis[passed])
    if parts > 1:
        assert np.array_equal(w["part"][passed], part_of[passed])
        assert w["kept"] == [int((passed & (part_of == k)).sum()) for k in range(parts)]
    # a region sized for several parts still carries single-layout rounds (small shards)
    w1 = records.decode_wire(_encode(V, cap, parts, passed, avg, vis, None), V, cap, parts)
    assert w1["used"] == 1 and np.array_equal(w1["passed"], passed)


def test_decode_wire_rejects_inconsistent_prefix():
    V, cap, n = 12, 4096, 4000
    rng = np.random.default_rng(0)
    passed = rng.random(n) < 0.5
    reg = _encode(V, cap, 1, passed, rng.random(n), rng.integers(1, 99, (n, 1)).astype(np.uint64))
    reg[16 + 8 * 3 + 4] ^= 1                                           # the running prefix of word 3
    with pytest.raises(AssertionError):
        records.decode_wire(reg, V, cap, 1)


def test_geometry_partition_fits_the_region():
    for cap in (1, 1000, 1 << 20):
        for parts in (1, 2, 5, 8):
            g = records.wire_geometry(48, cap, parts)
            assert g["region_bytes"] % 256 == 0 and g["sub_bytes"] % 256 == 0
            assert g["part_cap"] * parts >= cap and g["part_cap"] % 1024 == 0
            if parts > 1:
                assert parts * g["sub_bytes"] <= g["region_bytes"]
