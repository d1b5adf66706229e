"""PLY writer for the point cloud the MVS stage exports.

Replaces utils.py:249-251 (export2ply), which builds a pandas DataFrame with columns
x, y, z, red, green, blue from np.hstack((points, colors)) -- all float64 -- and hands it to
pyntcloud (not installed here).  pyntcloud writes a binary little-endian PLY whose vertex
properties carry the DataFrame dtypes, i.e. six ``double`` properties; this writer emits the
same container directly."""
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
