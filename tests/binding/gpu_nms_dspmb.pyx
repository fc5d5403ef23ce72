# cython: language_level=3
# The reference's cython/gpu_nms.pyx:16-31 bound to libdspmb instead of nms_kernel.cu: same Python signature
# (gpu_nms(dets, thresh, device_id=0) -> list of kept original indices), same host-side argsort, one C call.
# Only the extern block and the call differ from the reference file (INTEGRATION.md section 4).
import numpy as np
cimport numpy as np

cdef extern from "dspmb.h":
    int dspmb_nms_host(int *keep_out, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim,
                       float nms_overlap_thresh, int device_id)
    const char *dspmb_last_error()


def gpu_nms(np.ndarray[np.float32_t, ndim=2] dets, float thresh, np.int32_t device_id=0):
    cdef int boxes_num = dets.shape[0]
    cdef int boxes_dim = dets.shape[1]
    cdef int num_out = 0
    cdef np.ndarray[np.int32_t, ndim=1] keep = np.zeros(max(boxes_num, 1), dtype=np.int32)
    cdef np.ndarray[np.float32_t, ndim=1] scores = dets[:, 4]
    cdef np.ndarray[np.intp_t, ndim=1] order = scores.argsort()[::-1]
    cdef np.ndarray[np.float32_t, ndim=2] sorted_dets = np.ascontiguousarray(dets[order, :])
    cdef int rc = 0
    if boxes_num > 0:
        rc = dspmb_nms_host(<int *> &keep[0], &num_out, &sorted_dets[0, 0], boxes_num, boxes_dim, thresh, device_id)
    if rc != 0:
        raise RuntimeError(dspmb_last_error().decode())
    keep = keep[:num_out]
    return list(order[keep])
