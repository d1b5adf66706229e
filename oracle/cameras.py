"""Camera model of the reference's scoring path, restated in NumPy (fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows:
  * utils.py:56-81   read_pars   -- Middlebury ``*_par.txt`` layout
  * utils.py:241-244 projectPoint -- cv2.Rodrigues(R) then cv2.projectPoints
    (which rebuilds R' = Rodrigues(rvec)); only K[0,0], K[1,1], K[0,2], K[1,2]
    enter the projection, skew and distortion are ignored.

cv2 is a third-party wheel that is not part of /root/reference (cv2 4.13.0 in the
build container, unpinned by the reference).  The two cv2 algorithms on the path
are restated from OpenCV's published calib3d implementation:
  matrix -> rvec : R <- U Vt (SVD), r = (R21-R12, R02-R20, R10-R01),
                   s = |r|/2, c = (tr R - 1)/2, theta = acos(c), rvec = r*theta/(2 s)
  rvec -> matrix : R' = cos(th) I + (1-cos(th)) k k^T + sin(th) [k]x, k = rvec/th
tests/test_oracle_cpu.py checks the restatement against cv2 itself when cv2 is
importable (agreement ~1e-16 per entry, ~1e-12 px after projection).
"""
import numpy as np


def parse_par_text(text):
    """utils.py:56-81. Returns K [V,3,3], R [V,3,3], t [V,3] (fp64).

    First line = view count; each following line = name + 9 K + 9 R + 3 t,
    row-major.  View index = line index - 1."""
    lines = [ln for ln in text.splitlines() if ln.strip()]
    K, R, t = [], [], []
    for ln in lines[1:]:
        v = [float(x) for x in ln.split()[1:]]
        K.append(np.array(v[0:9]).reshape(3, 3))
        R.append(np.array(v[9:18]).reshape(3, 3))
        t.append(np.array(v[18:21]))
    return np.array(K), np.array(R), np.array(t)


def rodrigues_roundtrip(R):
    """R' = Rodrigues(Rodrigues(R)) as done implicitly by utils.py:242-243."""
    R = np.asarray(R, dtype=np.float64)
    U, _, Vt = np.linalg.svd(R)
    Q = U @ Vt
    r = np.array([Q[2, 1] - Q[1, 2], Q[0, 2] - Q[2, 0], Q[1, 0] - Q[0, 1]])
    s = np.sqrt((r * r).sum() * 0.25)
    c = (Q[0, 0] + Q[1, 1] + Q[2, 2] - 1.0) * 0.5
    c = min(1.0, max(-1.0, c))
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            rvec = np.zeros(3)
        else:
            tt = np.sqrt(np.maximum((np.diag(Q) + 1.0) * 0.5, 0.0))
            x = tt[0]
            y = tt[1] * (-1.0 if Q[0, 1] < 0 else 1.0)
            z = tt[2] * (-1.0 if Q[0, 2] < 0 else 1.0)
            if abs(x) < abs(y) and abs(x) < abs(z) and ((Q[1, 2] > 0) != (y * z > 0)):
                z = -z
            v = np.array([x, y, z])
            rvec = v * (theta / np.sqrt((v * v).sum()))
    else:
        rvec = r * (theta / (2.0 * s))
    th = np.sqrt((rvec * rvec).sum())
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = rvec / th
    cth, sth = np.cos(th), np.sin(th)
    kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return cth * np.eye(3) + (1.0 - cth) * np.outer(k, k) + sth * kx


class Cameras:
    """Per-view pinhole parameters the scorer uses (fp64)."""

    def __init__(self, K, R, t, roundtrip=True):
        self.K = np.asarray(K, dtype=np.float64).reshape(-1, 3, 3)
        self.R_file = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
        self.t = np.asarray(t, dtype=np.float64).reshape(-1, 3)
        self.V = self.K.shape[0]
        if roundtrip:
            self.R = np.stack([rodrigues_roundtrip(r) for r in self.R_file])
        else:
            self.R = self.R_file.copy()
        self.fx = self.K[:, 0, 0].copy()
        self.fy = self.K[:, 1, 1].copy()
        self.cx = self.K[:, 0, 2].copy()
        self.cy = self.K[:, 1, 2].copy()

    @classmethod
    def from_par_text(cls, text):
        return cls(*parse_par_text(text))

    def centres(self):
        """MVS2.py:188-189: camera_pos[i] = -(R_i^T t_i), with the FILE rotation."""
        return -np.einsum("vji,vj->vi", self.R_file, self.t)

    def project(self, c, view):
        """utils.py:241-244 for arrays: c [N,3], view [N] -> x [N], y [N] (fp64)."""
        c = np.asarray(c, dtype=np.float64).reshape(-1, 3)
        view = np.asarray(view).reshape(-1)
        R = self.R[view]
        t = self.t[view]
        # explicit left-to-right sums: the device kernel evaluates the same order
        X = R[:, 0, 0] * c[:, 0] + R[:, 0, 1] * c[:, 1] + R[:, 0, 2] * c[:, 2] + t[:, 0]
        Y = R[:, 1, 0] * c[:, 0] + R[:, 1, 1] * c[:, 1] + R[:, 1, 2] * c[:, 2] + t[:, 1]
        Z = R[:, 2, 0] * c[:, 0] + R[:, 2, 1] * c[:, 1] + R[:, 2, 2] * c[:, 2] + t[:, 2]
        # cv2.projectPoints: z = z ? 1./z : 1; x *= z; y *= z; u = x*fx + cx
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            iz = np.where(Z != 0.0, 1.0 / np.where(Z != 0.0, Z, 1.0), 1.0)
            x = (X * iz) * self.fx[view] + self.cx[view]
            y = (Y * iz) * self.fy[view] + self.cy[view]
        return x, y
