"""PLY writer for the point cloud the MVS stage exports.

Replaces utils.py:249-251 (export2ply), which builds a pandas DataFrame with columns
x, y, z, red, green, blue from np.hstack((points, colors)) -- all float64 -- and hands it to
pyntcloud (not installed here, so the container below is UNVERIFIED against its writer).  pyntcloud
writes a binary little-endian PLY with one vertex property per DataFrame column and the raw
records (here six float64 per vertex) behind the header; this writer emits that payload with the
properties declared as ``double``, which is what the 48-byte records are.  (From memory of
pyntcloud's io/ply.py -- not checkable here -- its header generator names a property by the first
letter of the dtype, i.e. ``float`` for float64 too, which would make its own float64 files
unreadable by standard PLY readers; if so, this writer deliberately does not reproduce that.)"""
import numpy as np


def export2ply(points, colors, path="output"):
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    colors = np.asarray(colors, dtype=np.float64).reshape(-1, 3)
    data = np.ascontiguousarray(np.hstack((points, colors)).astype("<f8"))
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\n"
              "property double x\nproperty double y\nproperty double z\n"
              "property double red\nproperty double green\nproperty double blue\nend_header\n") % len(data)
    with open(path + ".ply", "wb") as f:
        f.write(header.encode("ascii"))
        f.write(data.tobytes())


def read_ply(path):
    """Reads back what export2ply wrote (tests)."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    n = int([ln for ln in raw[:end].decode().splitlines() if ln.startswith("element vertex")][0].split()[-1])
    return np.frombuffer(raw[end:], dtype="<f8").reshape(n, 6)
