"""TEST INFRASTRUCTURE ONLY -- ctypes / import front-end of oracle/_ref/ (the reference's own sources compiled in
place by oracle/build_ref.py).  Same call signatures as oracle/oracle.py so tests can run either against the other.
``available()`` is False where neither /root/reference nor a prebuilt oracle/_ref/ exists."""
import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_lib = None
_nms = None


def available():
    return os.path.exists(os.path.join(_DIR, "libmultibox_ref.so"))


def nms_available():
    return bool(glob.glob(os.path.join(_DIR, "cpu_nms_ref*.so")))


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(os.path.join(_DIR, "libmultibox_ref.so"))
        fp = ctypes.POINTER(ctypes.c_float)
        L.ref_multibox_prior.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, fp, ctypes.c_int,
                                         ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int]
        L.ref_multibox_target.argtypes = [fp] * 6 + [ctypes.c_int] * 5 + [ctypes.c_float] * 4 + [ctypes.c_int, fp]
        L.ref_multibox_detection.argtypes = [fp] * 4 + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, fp,
                                                                           ctypes.c_float, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


class RefError(RuntimeError):
    def __init__(self, code):
        super().__init__("reference CHECK failed (code %d)" % code)
        self.code = code


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def multibox_prior(in_height, in_width, sizes=(1.0,), ratios=(1.0,), clip=False, steps=(-1.0, -1.0), offsets=(0.5, 0.5)):
    sizes, ratios = _f32(sizes), _f32(ratios)
    out = np.empty((1, in_height * in_width * (len(sizes) + len(ratios) - 1), 4), np.float32)
    rc = lib().ref_multibox_prior(_p(out), in_height, in_width, _p(sizes), len(sizes), _p(ratios), len(ratios),
                                  float(steps[0]), float(steps[1]), float(offsets[0]), float(offsets[1]), int(bool(clip)))
    if rc:
        raise RefError(rc)
    return out


def multibox_target(anchor, label, cls_pred, overlap_threshold=0.5, ignore_label=-1.0, negative_mining_ratio=-1.0,
                    negative_mining_thresh=0.5, minimum_negative_samples=0, variances=(0.1, 0.1, 0.2, 0.2)):
    anchor, label, cls_pred, var = _f32(anchor), _f32(label), _f32(cls_pred), _f32(variances)
    B, L, W = label.shape
    A, C = anchor.shape[1], cls_pred.shape[1]
    loc_target = np.empty((B, A * 5), np.float32)
    loc_mask = np.empty((B, A * 5), np.float32)
    cls_target = np.empty((B, A), np.float32)
    rc = lib().ref_multibox_target(_p(anchor), _p(label), _p(cls_pred), _p(loc_target), _p(loc_mask), _p(cls_target),
                                   B, A, L, W, C, overlap_threshold, ignore_label, negative_mining_ratio,
                                   negative_mining_thresh, minimum_negative_samples, _p(var))
    if rc:
        raise RefError(rc)
    return [loc_target, loc_mask, cls_target]


def multibox_detection(cls_prob, loc_pred, anchor, clip=True, threshold=0.01, background_id=0, nms_threshold=0.5,
                       force_suppress=False, variances=(0.1, 0.1, 0.2, 0.2), nms_topk=-1):
    cls_prob, loc_pred, anchor, var = _f32(cls_prob), _f32(loc_pred), _f32(anchor), _f32(variances)
    B, C, A = cls_prob.shape
    out = np.empty((B, A, 7), np.float32)
    rc = lib().ref_multibox_detection(_p(cls_prob), _p(loc_pred), _p(anchor), _p(out), B, A, C, threshold,
                                      int(bool(clip)), _p(var), nms_threshold, int(bool(force_suppress)), nms_topk)
    if rc:
        raise RefError(rc)
    return out


def cpu_nms(dets, thresh):
    """The reference's Cython cpu_nms (cython/cpu_nms.pyx) itself."""
    global _nms
    if _nms is None:
        path = glob.glob(os.path.join(_DIR, "cpu_nms_ref*.so"))[0]
        spec = importlib.util.spec_from_file_location("cpu_nms_ref", path)
        _nms = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_nms)
    return [int(i) for i in _nms.cpu_nms(np.ascontiguousarray(dets, dtype=np.float32), float(thresh))]


_bbox = None


def bbox_available():
    return bool(glob.glob(os.path.join(_DIR, "bbox_ref*.so")))


def bbox_overlaps_cython(boxes, query_boxes):
    """The reference's Cython bbox_overlaps_cython (cython/bbox.pyx:15-55) itself."""
    global _bbox
    if _bbox is None:
        path = glob.glob(os.path.join(_DIR, "bbox_ref*.so"))[0]
        spec = importlib.util.spec_from_file_location("bbox_ref", path)
        _bbox = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_bbox)
    return _bbox.bbox_overlaps_cython(np.ascontiguousarray(boxes, dtype=np.float64),
                                      np.ascontiguousarray(query_boxes, dtype=np.float64))


_gpu_nms = None


def gpu_nms_available():
    return os.path.exists(os.path.join(_DIR, "libgpu_nms_ref.so"))


def gpu_nms_sorted(sorted_dets, thresh, device_id=0):
    """The reference's own GPU NMS, `_nms` of cython/nms_kernel.cu:91-144 compiled unmodified for sm_100a (speed
    comparator only).  Same contract as the C++ function: rows already in descending score order, host pointers in and
    out, synchronous; returns the kept sorted positions."""
    global _gpu_nms
    if _gpu_nms is None:
        L = ctypes.CDLL(os.path.join(_DIR, "libgpu_nms_ref.so"))
        f = getattr(L, "_Z4_nmsPiS_PKfiifi")  # void _nms(int*, int*, const float*, int, int, float, int)
        f.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float),
                      ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        f.restype = None
        _gpu_nms = f
    d = np.ascontiguousarray(sorted_dets, dtype=np.float32)
    keep = np.empty(d.shape[0], np.int32)
    num = ctypes.c_int(0)
    _gpu_nms(keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.byref(num),
             d.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), d.shape[0], d.shape[1], float(thresh), device_id)
    return keep[: num.value]
