"""TEST INFRASTRUCTURE ONLY.  Loads the reference's own detect/nms.py from /root/reference with its two Cython imports
(`cython.cpu_nms`, `cython.gpu_nms`, lines 2-3) stubbed, so that its numpy `nms()` (:24-58) can pin oracle.py_nms in the
build container; /root/reference does not exist on the GPU box."""
import importlib.util
import os
import sys
import types

REF = os.environ.get("DSPNET_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF, "detect", "nms.py"))


_mod = None


def module():
    global _mod
    if _mod is None:
        saved = {k: sys.modules.get(k) for k in ("cython", "cython.cpu_nms", "cython.gpu_nms")}
        pkg = types.ModuleType("cython")
        a, b = types.ModuleType("cython.cpu_nms"), types.ModuleType("cython.gpu_nms")
        a.cpu_nms = b.gpu_nms = None  # only the names are imported; nms() itself is pure numpy
        pkg.cpu_nms, pkg.gpu_nms = a, b
        sys.modules.update({"cython": pkg, "cython.cpu_nms": a, "cython.gpu_nms": b})
        try:
            spec = importlib.util.spec_from_file_location("ref_detect_nms", os.path.join(REF, "detect", "nms.py"))
            _mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_mod)
        finally:
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    return _mod


def nms(dets, thresh):
    return [int(i) for i in module().nms(dets, thresh)]
