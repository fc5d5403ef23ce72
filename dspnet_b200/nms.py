"""Drop-in replacements for the reference's NMS helpers, all running on the GPU.

Mirrors ``cython.cpu_nms.cpu_nms`` (cython/cpu_nms.pyx:17), ``cython.gpu_nms.gpu_nms`` (cython/gpu_nms.pyx:16) and
``detect/nms.py`` (``nms`` and the three ``*_nms_wrapper`` factories, :6-58): same arguments, same return value
(a Python list of kept ORIGINAL row indices in descending score order).  The three reference implementations use
two different comparison rules; each name keeps its own rule:

    cpu_nms          suppress iff float64(iou) >= thresh      (``rule='ge'``)
    gpu_nms, nms     suppress iff iou > float32(thresh)       (``rule='gt'``)

Unlike the reference's gpu_nms nothing is sorted, swept or allocated on the host.

Deliberate differences from the reference helpers:
  * everything is computed in float32 (what the Cython helpers do: their signatures are ``np.float32_t``);
    ``detect/nms.py::nms`` would keep float64 input in float64 -- float64 arrays are converted here;
  * equal scores are ordered like a stable ``argsort()[::-1]`` (ties: higher index first).  numpy's default argsort is
    not stable, so for tied scores the reference's own order is unspecified and parity is only defined for tie-free
    scores;
  * ``evalmap.postfilter`` pads / truncates to ``max_rows`` and returns the per-image counts; the reference's
    ``multi_solver.py:429`` raises when more than 200 rows survive -- compare ``counts`` with ``max_rows`` for that.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DspmbError

__all__ = ["cpu_nms", "gpu_nms", "nms", "nms_device", "py_nms_wrapper", "cpu_nms_wrapper", "gpu_nms_wrapper"]

_workspaces = {}


def nms_device(dets, thresh, rule="ge", class_col=-1, presorted=False):
    """NMS on a CUDA tensor ``dets`` (N, >=5) ``[x1, y1, x2, y2, score, ...]``; returns (keep int32 (N,), num_keep
    int32 (1,)) device tensors without synchronising.  ``class_col >= 5`` suppresses only within equal values of
    that column (the per-class mode of the sweep in BASELINE.json)."""
    if not (isinstance(dets, torch.Tensor) and dets.is_cuda and dets.dtype == torch.float32 and dets.dim() == 2):
        raise DspmbError(_lib.ERR_BAD_ARG, "nms_device: dets must be a 2-D float32 CUDA tensor")
    dets = dets.contiguous()
    n, dim = dets.shape
    dev = dets.device
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        nbytes = L.dspmb_nms_workspace_bytes(n)
        key = (dev.index, stream)
        ws = _workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            _workspaces[key] = ws
        _lib.check(L.dspmb_nms_f32(ctypes.c_void_p(dets.data_ptr()), n, dim, float(thresh), 0 if rule == "ge" else 1,
                                   int(class_col), int(bool(presorted)), ctypes.c_void_p(keep.data_ptr()),
                                   ctypes.c_void_p(num.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
                                   ctypes.c_void_p(stream)))
    return keep, num


def _host_nms(dets, thresh, rule, device_id=0, class_col=-1):
    if not torch.cuda.is_available():
        raise DspmbError(_lib.ERR_CUDA, "no CUDA device: dspnet_b200 has no CPU fallback")
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] < 5:
        raise DspmbError(_lib.ERR_BAD_ARG, "nms: dets must be (N, >=5) [x1, y1, x2, y2, score]")
    if dets.shape[0] == 0:
        return []
    with torch.cuda.device(device_id):
        d = torch.from_numpy(dets).cuda(non_blocking=True)
        keep, num = nms_device(d, thresh, rule=rule, class_col=class_col)
        k = int(num.item())
        return keep[:k].cpu().tolist()


def cpu_nms(dets, thresh):
    """cython/cpu_nms.pyx:17-68 semantics (``ovr >= thresh`` with a double threshold)."""
    return _host_nms(dets, thresh, "ge")


def gpu_nms(dets, thresh, device_id=0):
    """cython/gpu_nms.pyx:16-31 + cython/nms_kernel.cu semantics (``iou > thresh`` in float)."""
    return _host_nms(dets, thresh, "gt", device_id)


def nms(dets, thresh):
    """detect/nms.py:24-58 semantics (keeps ``ovr <= thresh``, i.e. the gpu_nms rule)."""
    return _host_nms(dets, thresh, "gt")


def py_nms_wrapper(thresh):
    def _nms(dets):
        return nms(dets, thresh)
    return _nms


def cpu_nms_wrapper(thresh):
    def _nms(dets):
        return cpu_nms(dets, thresh)
    return _nms


def gpu_nms_wrapper(thresh, device_id):
    def _nms(dets):
        return gpu_nms(dets, thresh, device_id)
    return _nms
