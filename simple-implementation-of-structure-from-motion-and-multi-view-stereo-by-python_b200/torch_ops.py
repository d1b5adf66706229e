"""The C ABI of include/mvs_ncc.h as PyTorch custom operators (``torch.ops.mvs.*``) -- the PyTorch-extension
face of the boundary BASELINE.json's north_star asks for ("a thin C-ABI layer ... exposed as a PyTorch
extension because the repo is Python").

Zero-copy: the operators hand the tensors' device pointers and torch's CURRENT stream to libmvsncc.so; nothing
is staged through the host and nothing synchronises.  The context handle travels as an int64 (the
``mvs_ctx*``).  Importing this module registers the operators; every one of them fails loudly without the
library or a B200 (no CPU kernels are registered -- only shape functions for tracing):

    h = torch.ops.mvs.create(rgb_u8_cuda, K, R, t)                          # mvs_create      (main.py:7-20, utils.py:56-81)
    vis, avg, count, xy = torch.ops.mvs.score_batch(h, c, ref, 0.7, 5)      # mvs_score_batch (MVS2.py:62-77)
    idx, best = torch.ops.mvs.score_pmvs_select(h, c, n, ref, 0.7, 5, 64, 3)  # mvs_score_pmvs (north_star extension)
    torch.ops.mvs.destroy(h)
"""
import ctypes as C
from typing import Tuple

import torch
from torch import Tensor

from . import _lib
from .context import MvsError, _check


def _views(h: int) -> int:
    V = C.c_int()
    _check(_lib.load().mvs_get_info(C.c_void_p(h), C.byref(V), None, None, None), "mvs_get_info")
    return V.value


def _stream(t: Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dev_f64(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise MvsError(f"torch.ops.mvs: {what} must be a CUDA tensor (there is no CPU implementation)")
    return t.to(torch.float64).contiguous()


@torch.library.custom_op("mvs::create", mutates_args=())
def create(rgb: Tensor, K: Tensor, R: Tensor, t: Tensor) -> int:
    """rgb [V,H,W,3] uint8 CUDA, K/R [V,3,3], t [V,3] (any device, float64) -> context handle."""
    if not rgb.is_cuda or rgb.dtype != torch.uint8 or rgb.dim() != 4 or rgb.shape[3] != 3:
        raise MvsError("torch.ops.mvs.create: rgb must be a [V,H,W,3] uint8 CUDA tensor")
    rgb = rgb.contiguous()
    V, H, W, _ = rgb.shape
    Kh, Rh, th = (x.detach().to("cpu", torch.float64).contiguous() for x in (K.reshape(V, 9), R.reshape(V, 9), t.reshape(V, 3)))
    torch.cuda.current_stream(rgb.device).synchronize()           # mvs_create reads rgb on its own stream
    h = C.c_void_p()
    _check(_lib.load().mvs_create(C.byref(h), rgb.device.index or 0, V, H, W, C.c_void_p(rgb.data_ptr()), 1,
                                  C.c_void_p(Kh.data_ptr()), C.c_void_p(Rh.data_ptr()), None, C.c_void_p(th.data_ptr())),
           "mvs_create")
    return int(h.value)


@create.register_fake
def _(rgb, K, R, t):
    return 0


@torch.library.custom_op("mvs::destroy", mutates_args=())
def destroy(ctx: int) -> None:
    _lib.load().mvs_destroy(C.c_void_p(ctx))


@destroy.register_fake
def _(ctx):
    return None


@torch.library.custom_op("mvs::score_batch", mutates_args=())
def score_batch(ctx: int, c: Tensor, ref: Tensor, min_ncc: float, wid: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Mode A (MVS2.py:62-77) for c [N,3] f64, ref [N] i32 on the device ->
    vis_mask [N, ceil(V/64)] i64, avg [N] f64, count [N] i32, xy [N,2] f64."""
    c = _dev_f64(c, "c").reshape(-1, 3)
    ref = ref.to(torch.int32).contiguous()
    N, dev = c.shape[0], c.device
    mw = (_views(ctx) + 63) // 64
    vis = torch.empty((N, mw), dtype=torch.int64, device=dev)
    avg = torch.empty(N, dtype=torch.float64, device=dev)
    count = torch.empty(N, dtype=torch.int32, device=dev)
    xy = torch.empty((N, 2), dtype=torch.float64, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    _check(_lib.load().mvs_score_batch(C.c_void_p(ctx), 0, N, p(c), None, p(ref), float(min_ncc), int(wid), p(vis), p(avg), p(count),
                                       p(xy), None, 1, _stream(c)), "mvs_score_batch")
    return vis, avg, count, xy


@score_batch.register_fake
def _(ctx, c, ref, min_ncc, wid):
    N = c.shape[0]
    mw = 1                                                         # the mask width depends on the context; traced shapes assume V <= 64
    return (c.new_empty((N, mw), dtype=torch.int64), c.new_empty(N, dtype=torch.float64), c.new_empty(N, dtype=torch.int32),
            c.new_empty((N, 2), dtype=torch.float64))


@torch.library.custom_op("mvs::score_pmvs_select", mutates_args=())
def score_pmvs_select(ctx: int, c: Tensor, nrm: Tensor, ref: Tensor, min_ncc: float, mu: int, group: int,
                      bound: int) -> Tuple[Tensor, Tensor]:
    """Mode B with on-chip selection over consecutive sets of `group` hypotheses ->
    best_idx [ceil(N/group)] i32, best_avg f64 (mvs_score_pmvs; north_star extension)."""
    c = _dev_f64(c, "c").reshape(-1, 3)
    nrm = _dev_f64(nrm, "nrm").reshape(-1, 3)
    ref = ref.to(torch.int32).contiguous()
    N, dev = c.shape[0], c.device
    ns = (N + group - 1) // group
    bi = torch.empty(ns, dtype=torch.int32, device=dev)
    ba = torch.empty(ns, dtype=torch.float64, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr())
    _check(_lib.load().mvs_score_pmvs(C.c_void_p(ctx), N, p(c), p(nrm), p(ref), None, float(min_ncc), int(mu), 0, int(group), int(bound),
                                      None, None, None, None, None, p(bi), p(ba), 1, _stream(c)), "mvs_score_pmvs")
    return bi, ba


@score_pmvs_select.register_fake
def _(ctx, c, nrm, ref, min_ncc, mu, group, bound):
    ns = (c.shape[0] + group - 1) // group
    return c.new_empty(ns, dtype=torch.int32), c.new_empty(ns, dtype=torch.float64)
