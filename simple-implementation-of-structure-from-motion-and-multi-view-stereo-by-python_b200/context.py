"""MvsContext: the image stack + cameras resident on one B200, and batched scoring.

Host-side mirror of the C ABI in include/mvs_ncc.h.  PyTorch is used only as
plumbing (device memory for outputs, the current stream); all arithmetic happens
in libmvsncc.so.  No fallback exists: without the library or without a B200 every
call raises.
"""
import ctypes as C

import numpy as np

from . import _lib

MODE_REFEXACT = 0
MODE_PMVS = 1
PMVS_REDUCE_TO_REFEXACT = 1


class MvsError(Exception):
    """Raised for every failure of the device path.  Deliberately NOT a RuntimeError: the
    reference's main.py swallows RuntimeError (main.py:43-46) and a missing GPU or library
    must stop the program loudly instead of printing one word."""


def _check(rc, what):
    if rc != 0:
        msg = _lib.load().mvs_last_error().decode("utf-8", "replace")
        raise MvsError(f"{what} failed (code {rc}): {msg}")


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def mask_words(V):
    return (V + 63) // 64


def unpack_vis(mask, V):
    """[N, ceil(V/64)] uint64 -> [N, V] bool (bit v of the row = view v)."""
    m = np.ascontiguousarray(mask, dtype=np.uint64).reshape(len(mask), -1)
    bits = np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :V].astype(bool)


class MvsContext:
    """Loads images (list/array of H x W x 3 uint8 RGB, main.py:7-20 layout) and
    cameras (K, R, t as utils.py:56-81 returns them) into HBM once."""

    def __init__(self, rgb, K, R, t, Rrt=None, device=0):
        self._h = None
        lib = _lib.load()
        if isinstance(rgb, (list, tuple)):
            rgb = np.stack([np.asarray(im) for im in rgb])
        self._rgb_torch = None
        on_device = 0
        if hasattr(rgb, "data_ptr"):                      # torch tensor on the device
            if not rgb.is_cuda or rgb.dtype.__str__() != "torch.uint8":
                raise MvsError("device image stack must be a CUDA uint8 tensor")
            rgb = rgb.contiguous()
            shape = tuple(rgb.shape)
            ptr = C.c_void_p(rgb.data_ptr())
            on_device = 1
            self._rgb_torch = rgb
        else:
            rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
            shape = rgb.shape
            ptr = _np_ptr(rgb)
        if len(shape) != 4 or shape[3] != 3:
            raise MvsError(f"image stack must be [V,H,W,3] uint8, got {shape}")
        V, H, W, _ = shape
        K = np.ascontiguousarray(np.asarray(K, dtype=np.float64).reshape(V, 9))
        R = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(V, 9))
        t = np.ascontiguousarray(np.asarray(t, dtype=np.float64).reshape(V, 3))
        if Rrt is not None:
            Rrt = np.ascontiguousarray(np.asarray(Rrt, dtype=np.float64).reshape(V, 9))
        h = C.c_void_p()
        _check(lib.mvs_create(C.byref(h), int(device), V, H, W, ptr, on_device, _np_ptr(K), _np_ptr(R), _np_ptr(Rrt),
                              _np_ptr(t)), "mvs_create")
        self._h = h
        self._rgb_torch = None
        self.V, self.H, self.W, self.device = V, H, W, int(device)
        self.K, self.R, self.t = K.reshape(V, 3, 3), R.reshape(V, 3, 3), t

    # -- lifetime ------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            _lib.load().mvs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- introspection ---------------------------------------------------------------
    def gray(self):
        out = np.empty((self.V, self.H, self.W), dtype=np.uint8)
        _check(_lib.load().mvs_download_gray(self._h, _np_ptr(out)), "mvs_download_gray")
        return out

    def cameras(self):
        rrt = np.empty((self.V, 3, 3))
        cen = np.empty((self.V, 3))
        _check(_lib.load().mvs_get_cameras(self._h, _np_ptr(rrt), _np_ptr(cen)), "mvs_get_cameras")
        return rrt, cen

    def launch_count(self):
        return int(_lib.load().mvs_launch_count(self._h))

    def profile(self, on=True):
        """Bracket every scoring kernel with CUDA events on its launch stream."""
        _check(_lib.load().mvs_profile_enable(self._h, 1 if on else 0), "mvs_profile_enable")

    def probe(self, on=True):
        """Swap K1 for the loads-only gather-ceiling probe (measurement only; outputs are not written)."""
        _check(_lib.load().mvs_profile_probe(self._h, 1 if on else 0), "mvs_profile_probe")

    def score_kernel_ms(self):
        """(mean ms, n) over the scoring kernels (K1 alone) launched since profile(True), at most
        the last 64; waits for them to finish."""
        ms, n = C.c_float(), C.c_int()
        _check(_lib.load().mvs_profile_score_ms(self._h, C.byref(ms), C.byref(n)), "mvs_profile_score_ms")
        return float(ms.value), int(n.value)

    # -- scoring ---------------------------------------------------------------------
    def score_host(self, c, ref, min_ncc=0.7, wid=5, nrm=None, mode=MODE_REFEXACT, want_ncc=False):
        """Host buffers in, host buffers out (copies + sync inside the call).
        Returns dict(vis_mask [N,mw] u64, avg [N] f64, count [N] i32, xy [N,2] f64[, ncc [N,V] f32])."""
        c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1, 3))
        N = c.shape[0]
        ref = np.ascontiguousarray(np.asarray(ref, dtype=np.int32).reshape(-1))
        if ref.shape[0] != N:
            raise MvsError("c and ref disagree on N")
        if nrm is not None:
            nrm = np.ascontiguousarray(np.asarray(nrm, dtype=np.float64).reshape(-1, 3))
        mw = mask_words(self.V)
        out = dict(vis_mask=np.zeros((N, mw), np.uint64), avg=np.zeros(N), count=np.zeros(N, np.int32),
                   xy=np.zeros((N, 2)))
        ncc = np.empty((N, self.V), np.float32) if want_ncc else None
        _check(_lib.load().mvs_score_batch(self._h, mode, N, _np_ptr(c), _np_ptr(nrm), _np_ptr(ref), float(min_ncc),
                                           int(wid), _np_ptr(out["vis_mask"]), _np_ptr(out["avg"]),
                                           _np_ptr(out["count"]), _np_ptr(out["xy"]), _np_ptr(ncc), 0, None),
               "mvs_score_batch")
        if want_ncc:
            out["ncc"] = ncc
        return out

    def score_device(self, c, ref, min_ncc=0.7, wid=5, nrm=None, mode=MODE_REFEXACT, out=None, want_ncc=False,
                     stream=None):
        """Device tensors in, device tensors out; only enqueues on the current stream.
        c [N,3] float64 cuda, ref [N] int32 cuda.  ``out`` may carry preallocated tensors."""
        import torch
        N = c.shape[0]
        dev = c.device
        mw = mask_words(self.V)
        if out is None:
            out = {}
        out.setdefault("vis_mask", torch.empty((N, mw), dtype=torch.int64, device=dev))
        out.setdefault("avg", torch.empty(N, dtype=torch.float64, device=dev))
        out.setdefault("count", torch.empty(N, dtype=torch.int32, device=dev))
        out.setdefault("xy", torch.empty((N, 2), dtype=torch.float64, device=dev))
        if want_ncc:
            out.setdefault("ncc", torch.empty((N, self.V), dtype=torch.float32, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        _check(_lib.load().mvs_score_batch(self._h, mode, N, p(c), p(nrm), p(ref), float(min_ncc), int(wid),
                                           p(out["vis_mask"]), p(out["avg"]), p(out["count"]), p(out["xy"]),
                                           p(out.get("ncc")), 1, C.c_void_p(stream)),
               "mvs_score_batch")
        return out

    # -- Mode B (north_star extension; spec = oracle/mode_b.py) --------------------------
    def score_pmvs_host(self, c, nrm, ref, min_ncc=0.7, mu=5, cand=None, flags=0, group=0, bound=3, want_ncc=False,
                        per_hypothesis=True):
        """Mode B through host buffers.  Returns the dict of score_host plus, when group > 1,
        best_idx [ceil(N/group)] i32 and best_avg f64; per_hypothesis=False skips every
        per-hypothesis output (only one record per selection set leaves the device)."""
        c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1, 3))
        N = c.shape[0]
        ref = np.ascontiguousarray(np.asarray(ref, dtype=np.int32).reshape(-1))
        if nrm is not None:
            nrm = np.ascontiguousarray(np.asarray(nrm, dtype=np.float64).reshape(-1, 3))
        mw = mask_words(self.V)
        if cand is not None:
            cand = np.ascontiguousarray(np.asarray(cand, dtype=np.uint64).reshape(N, mw))
        out = {}
        if per_hypothesis:
            out = dict(vis_mask=np.zeros((N, mw), np.uint64), avg=np.zeros(N), count=np.zeros(N, np.int32),
                       xy=np.zeros((N, 2)))
            if want_ncc:
                out["ncc"] = np.empty((N, self.V), np.float32)
        if group > 1:
            ns = (N + group - 1) // group
            out["best_idx"] = np.full(ns, -2, np.int32)
            out["best_avg"] = np.zeros(ns)
        g = out.get
        _check(_lib.load().mvs_score_pmvs(self._h, N, _np_ptr(c), _np_ptr(nrm), _np_ptr(ref), _np_ptr(cand), float(min_ncc),
                                          int(mu), int(flags), int(group), int(bound), _np_ptr(g("vis_mask")),
                                          _np_ptr(g("avg")), _np_ptr(g("count")), _np_ptr(g("xy")), _np_ptr(g("ncc")),
                                          _np_ptr(g("best_idx")), _np_ptr(g("best_avg")), 0, None), "mvs_score_pmvs")
        return out

    def score_pmvs_device(self, c, nrm, ref, min_ncc=0.7, mu=5, cand=None, flags=0, group=0, bound=3, out=None,
                          per_hypothesis=True, stream=None):
        """Mode B on device tensors (c, nrm [N,3] f64, ref [N] i32, cand [N,mw] i64); enqueues only."""
        import torch
        N = c.shape[0]
        dev = c.device
        mw = mask_words(self.V)
        out = {} if out is None else out
        if per_hypothesis:
            out.setdefault("vis_mask", torch.empty((N, mw), dtype=torch.int64, device=dev))
            out.setdefault("avg", torch.empty(N, dtype=torch.float64, device=dev))
            out.setdefault("count", torch.empty(N, dtype=torch.int32, device=dev))
            out.setdefault("xy", torch.empty((N, 2), dtype=torch.float64, device=dev))
        if group > 1:
            ns = (N + group - 1) // group
            out.setdefault("best_idx", torch.empty(ns, dtype=torch.int32, device=dev))
            out.setdefault("best_avg", torch.empty(ns, dtype=torch.float64, device=dev))
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        g = out.get
        _check(_lib.load().mvs_score_pmvs(self._h, N, p(c), p(nrm), p(ref), p(cand), float(min_ncc), int(mu), int(flags),
                                          int(group), int(bound), p(g("vis_mask")), p(g("avg")), p(g("count")),
                                          p(g("xy")), p(g("ncc")), p(g("best_idx")), p(g("best_avg")), 1,
                                          C.c_void_p(stream)), "mvs_score_pmvs")
        return out

    def select_best_device(self, avg, count, group, bound=3, stream=None):
        """Argmax over consecutive sets of an already scored device batch -> (best_idx i32, best_avg f64)."""
        import torch
        N = avg.shape[0]
        ns = (N + group - 1) // group
        bi = torch.empty(ns, dtype=torch.int32, device=avg.device)
        ba = torch.empty(ns, dtype=torch.float64, device=avg.device)
        if stream is None:
            stream = torch.cuda.current_stream(avg.device).cuda_stream
        _check(_lib.load().mvs_select_best(self._h, N, int(group), C.c_void_p(avg.data_ptr()), C.c_void_p(count.data_ptr()),
                                           int(bound), C.c_void_p(bi.data_ptr()), C.c_void_p(ba.data_ptr()),
                                           C.c_void_p(stream)), "mvs_select_best")
        return bi, ba
