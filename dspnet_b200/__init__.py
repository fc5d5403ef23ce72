"""dspnet_b200 -- B200-native (sm_100a) implementation of DSPNet's SSD-style multibox hot path.

Public surface (reference operator / helper names):

    from dspnet_b200 import MultiBoxPrior, MultiBoxTarget, MultiBoxDetection      # operator/multibox_*.{cc,cu}
    from dspnet_b200.nms import cpu_nms, gpu_nms, nms, gpu_nms_wrapper            # cython/*, detect/nms.py
    from dspnet_b200.symbol import multibox_anchors                               # symbol/common.py anchor branch
    from dspnet_b200.dist import ShardedMultiBox                                  # image-sharded multi-GPU driver
    from dspnet_b200.autograd import multibox_target, multibox_detection          # zero-gradient Backward (-inl.h)

All compute runs in libdspmb.so (hand-written CUDA behind the C ABI of include/dspmb.h); there is no CPU path.
"""
from .ops import (MultiBoxDetection, MultiBoxDetectionFromHeads, MultiBoxPrior, MultiBoxTarget, DspmbError,  # noqa: F401
                  multibox_prior_concat)

__version__ = "0.1.0"
