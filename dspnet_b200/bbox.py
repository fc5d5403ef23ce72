"""``bbox_overlaps_cython`` of the reference (cython/bbox.pyx:15-55) on the GPU.

Same name, argument order and result as the Cython helper: ``boxes`` (N, 4) and ``query_boxes`` (K, 4) float64
``[x1, y1, x2, y2]`` -> ``overlaps`` (N, K) float64 with the "+1" pixel convention.  numpy arrays in -> numpy array out
(host round trip); CUDA tensors in -> CUDA tensor out on the current stream.  There is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import DspmbError, _require_cuda


def bbox_overlaps_cython(boxes, query_boxes):
    _require_cuda()
    host = not (torch.is_tensor(boxes) and boxes.is_cuda)
    if host:
        dev = torch.device("cuda", torch.cuda.current_device())
        b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float64)).to(dev)
        q = torch.from_numpy(np.ascontiguousarray(query_boxes, dtype=np.float64)).to(dev)
    else:
        b = boxes.to(torch.float64).contiguous()
        q = torch.as_tensor(query_boxes, dtype=torch.float64, device=boxes.device).contiguous()
    if b.dim() != 2 or q.dim() != 2 or b.shape[1] != 4 or q.shape[1] != 4:
        raise DspmbError(_lib.ERR_BAD_ARG, "bbox_overlaps_cython: boxes (N, 4) and query_boxes (K, 4) expected")
    n, k = b.shape[0], q.shape[0]
    out = torch.empty((n, k), dtype=torch.float64, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(_lib.lib().dspmb_bbox_overlaps_f64(
            b.data_ptr(), n, q.data_ptr(), k, out.data_ptr(),
            ctypes.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)))
    return out.cpu().numpy() if host else out
