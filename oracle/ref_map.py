"""TEST INFRASTRUCTURE ONLY.  Loads the reference's own evaluate/eval_metric.py (MApMetric / VOC07MApMetric) from
/root/reference behind a three-line stand-in for the `mxnet` package (the file only needs `mx.metric.EvalMetric` as a
base class and `.asnumpy()` on its inputs).  Used in the build container to pin oracle/map_oracle.py and to generate
tests/golden/map_golden.npz; /root/reference does not exist on the GPU box."""
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("DSPNET_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF, "evaluate", "eval_metric.py"))


class _EvalMetric(object):  # what mx.metric.EvalMetric provides to MApMetric: a name and reset()
    def __init__(self, name, *args, **kwargs):
        self.name = name
        self.reset()

    def reset(self):
        self.num_inst = 0
        self.sum_metric = 0.0


class _ND(object):
    """Stand-in for mx.nd.array: indexing returns another wrapper, asnumpy() the data."""

    def __init__(self, a):
        self.a = np.asarray(a)
        self.shape = self.a.shape

    def __getitem__(self, i):
        return _ND(self.a[i])

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
            spec = importlib.util.spec_from_file_location("ref_eval_metric", os.path.join(REF, "evaluate", "eval_metric.py"))
            _mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_mod)
        finally:
            if saved is None:
                del sys.modules["mxnet"]
            else:
                sys.modules["mxnet"] = saved
    return _mod


def run_metric(labels, preds, ovp_thresh=0.5, use_difficult=False, voc07=False, batches=1):
    """Feeds (B, L, W) labels and (B, M, 6+) predictions through the reference metric; returns
    (records {cid: (n, 2)}, counts {cid: int}, (name, value) of get())."""
    m = module()
    metric = (m.VOC07MApMetric if voc07 else m.MApMetric)(ovp_thresh, use_difficult)
    B = labels.shape[0]
    step = (B + batches - 1) // batches
    for s in range(0, B, step):
        metric.update([_ND(labels[s:s + step])], [_ND(preds[s:s + step])])
    records = {int(k): np.array(v, dtype=np.float64) for k, v in metric.records.items()}
    counts = {int(k): int(v) for k, v in metric.counts.items()}
    return records, counts, metric.get()
