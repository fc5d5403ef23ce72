"""Pre-marshalled, allocation-free execution plans for the two batched operators.

``ops.MultiBoxDetection`` / ``ops.MultiBoxTarget`` validate, allocate outputs and marshal ~20 ctypes arguments on
every call, which costs more host time than the kernels take on a B200.  A plan does that once for a fixed
(B, A, C[, L]) and parameter set: ``run()`` is a single C-ABI call on the current stream, so it can sit in a
launch-bound loop or be captured into a CUDA graph (``capture()``).
"""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class DetectionPlan:
    """MultiBoxDetection (operator/multibox_detection-inl.h:47-72 parameters) for fixed shapes."""

    def __init__(self, B, A, C, device, clip=True, threshold=0.01, nms_threshold=0.5, force_suppress=False,
                 variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1, want_valid_count=False):
        self.B, self.A, self.C, self.device = B, A, C, torch.device(device)
        self._index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = _lib.lib()
        with torch.cuda.device(device):
            self.ws = torch.empty(max(self.lib.dspmb_detection_workspace_bytes(B, A, C), 256), dtype=torch.uint8,
                                  device=device)
        self.valid = torch.empty((B,), dtype=torch.int32, device=device) if want_valid_count else None
        self._var = _lib.float_array(variances)
        self._tail = (B, A, C, float(threshold), int(bool(clip)), self._var, float(nms_threshold),
                      int(bool(force_suppress)), int(nms_topk), _ptr(self.valid), _ptr(self.ws), self.ws.numel())
        self.launches_per_run = None  # set by the first run() from the library's own count

    def new_output(self):
        return torch.empty((self.B, self.A, 7), dtype=torch.float32, device=self.device)

    def new_valid_count(self):
        """(B,) int32 -- pass as run(..., valid=): rows at and beyond valid[b] of image b are untouched (-1)."""
        return torch.empty((self.B,), dtype=torch.int32, device=self.device)

    def run(self, cls_prob, loc_pred, anchor, out, stream=None, valid=None):
        """``valid``: per-call valid-count tensor instead of the plan's own (callers that keep several outputs in
        flight, e.g. beside the P2P gather, need one per output)."""
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if torch.cuda.current_device() != self._index:  # the library keys graphs / streams on the CURRENT device
            with torch.cuda.device(self.device):
                return self.run(cls_prob, loc_pred, anchor, out, s, valid)
        tail = self._tail if valid is None else self._tail[:9] + (valid.data_ptr(),) + self._tail[10:]
        rc = self.lib.dspmb_detection_f32(cls_prob.data_ptr(), loc_pred.data_ptr(), anchor.data_ptr(), out.data_ptr(),
                                          *tail, s)
        if rc:
            _lib.check(rc)
        if self.launches_per_run is None:
            self.launches_per_run = self.lib.dspmb_last_launch_count()
        return out


class DetectionHeadsPlan:
    """MultiBoxDetectionFromHeads (dspnet_b200.ops) for fixed shapes: the pointer tables of the per-scale heads are
    built once per set of head tensors (``bind``), a run is one C-ABI call."""

    def __init__(self, B, A, C, head_shapes, device, clip=True, threshold=0.01, nms_threshold=0.5, force_suppress=False,
                 variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1):
        """head_shapes: [(H, W, anchors_per_cell), ...] per scale."""
        self.B, self.A, self.C, self.device = B, A, C, torch.device(device)
        self._index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = _lib.lib()
        self.k = len(head_shapes)
        self._hw = _lib.int_array([v for (h, w, _) in head_shapes for v in (h, w)])
        self._na = _lib.int_array([n for (_, _, n) in head_shapes])
        with torch.cuda.device(device):
            nbytes = self.lib.dspmb_detection_heads_workspace_bytes(B, A, C, self._hw, self._na, self.k)
            self.ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        self._var = _lib.float_array(variances)
        self._tail = (B, A, C, float(threshold), int(bool(clip)), self._var, float(nms_threshold),
                      int(bool(force_suppress)), int(nms_topk), None, _ptr(self.ws), self.ws.numel())
        self.launches_per_run = None

    def new_output(self):
        return torch.empty((self.B, self.A, 7), dtype=torch.float32, device=self.device)

    def bind(self, cls_heads, loc_heads):
        """Pointer tables for one set of head tensors (keep the tensors alive while the binding is used)."""
        return ((ctypes.c_void_p * self.k)(*[h.data_ptr() for h in cls_heads]),
                (ctypes.c_void_p * self.k)(*[h.data_ptr() for h in loc_heads]))

    def run(self, binding, anchor, out, stream=None):
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if torch.cuda.current_device() != self._index:
            with torch.cuda.device(self.device):
                return self.run(binding, anchor, out, s)
        rc = self.lib.dspmb_detection_heads_f32(binding[0], binding[1], self._hw, self._na, self.k, anchor.data_ptr(),
                                                out.data_ptr(), *self._tail, s)
        if rc:
            _lib.check(rc)
        if self.launches_per_run is None:
            self.launches_per_run = self.lib.dspmb_last_launch_count()
        return out


class TargetPlan:
    """MultiBoxTarget (operator/multibox_target-inl.h:59-80 parameters) for fixed shapes."""

    def __init__(self, B, A, L, C, device, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                 negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2),
                 label_width=6, want_stats=True):
        self.B, self.A, self.L, self.C, self.device = B, A, L, C, torch.device(device)
        self._index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.lib = _lib.lib()
        with torch.cuda.device(device):
            self.ws = torch.empty(max(self.lib.dspmb_target_workspace_bytes(B, A, L, C), 256), dtype=torch.uint8,
                                  device=device)
        self.stats = torch.zeros((B, 4), dtype=torch.int32, device=device) if want_stats else None
        self._var = _lib.float_array(variances)
        self._head = (B, A, L, int(label_width), C, float(overlap_threshold), float(ignore_label),
                      float(negative_mining_ratio), float(negative_mining_thresh), int(minimum_negative_samples),
                      self._var, None)
        self._ws = (_ptr(self.ws), self.ws.numel())
        self.launches_per_run = None  # set by the first run() from the library's own count

    def new_outputs(self):
        d = self.device
        return (torch.empty((self.B, self.A * 5), dtype=torch.float32, device=d),
                torch.empty((self.B, self.A * 5), dtype=torch.float32, device=d),
                torch.empty((self.B, self.A), dtype=torch.float32, device=d))

    def new_stats(self):
        """(B, 4) int32 [num_valid_gt, num_positive, num_negative, num_bipartite] -- pass as run(..., stats=)."""
        return torch.zeros((self.B, 4), dtype=torch.int32, device=self.device)

    def run(self, anchor, label, cls_pred, outs, stream=None, stats=None):
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if torch.cuda.current_device() != self._index:
            with torch.cuda.device(self.device):
                return self.run(anchor, label, cls_pred, outs, s, stats)
        st = self.stats if stats is None else stats
        rc = self.lib.dspmb_target_f32(anchor.data_ptr(), label.data_ptr(), cls_pred.data_ptr(), outs[0].data_ptr(),
                                       outs[1].data_ptr(), outs[2].data_ptr(), *self._head, _ptr(st), *self._ws, s)
        if rc:
            _lib.check(rc)
        if self.launches_per_run is None:
            self.launches_per_run = self.lib.dspmb_last_launch_count()
        return outs

    def status(self):
        with torch.cuda.device(self.device):
            self._status()

    def _status(self):
        _lib.check(self.lib.dspmb_status(_ptr(self.ws), ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
