"""TEST INFRASTRUCTURE ONLY.  Loads the reference's own train/metric.py (MultiBoxMetric) from /root/reference behind a
stand-in for the `mxnet` package (the class only needs `mx.metric.EvalMetric` as a base and `.asnumpy()` on its
inputs).  Used in the build container to pin oracle.multibox_training_outputs' statistics and dspnet_b200.loss
.MultiBoxMetric; /root/reference does not exist on the GPU box."""
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("DSPNET_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF, "train", "metric.py"))


class _EvalMetric(object):
    def __init__(self, name, *args, **kwargs):
        self.name = name
        self.reset()

    def reset(self):
        self.num_inst = 0
        self.sum_metric = 0.0


class ND(object):
    def __init__(self, a):
        self.a = np.asarray(a)

    def asnumpy(self):
        return self.a.copy()


_mod = None


def module():
    global _mod
    if _mod is None:
        mx = types.ModuleType("mxnet")
        mx.metric = types.ModuleType("mxnet.metric")
        mx.metric.EvalMetric = _EvalMetric
        saved = sys.modules.get("mxnet")
        sys.modules["mxnet"] = mx
        try:
            spec = importlib.util.spec_from_file_location("ref_train_metric", os.path.join(REF, "train", "metric.py"))
            _mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_mod)
        finally:
            if saved is None:
                del sys.modules["mxnet"]
            else:
                sys.modules["mxnet"] = saved
    return _mod


def multibox_metric(cls_prob, loc_loss, cls_label, eps=1e-8):
    """(names, values, sums, counts) of the reference's MultiBoxMetric after one update()."""
    m = module().MultiBoxMetric(eps=eps)
    m.update(None, [ND(cls_prob), ND(loc_loss), ND(cls_label)])
    names, values = m.get()
    return names, values, list(m.sum_metric), list(m.num_inst)
