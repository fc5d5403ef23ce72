"""Autograd view of the three operators: forward through the CUDA path, Backward exactly as the reference declares it.

The reference's Backward methods only write zeros:
  * MultiBoxPrior      -- operator/multibox_prior-inl.h:131-143   grad(data) = 0
  * MultiBoxTarget     -- operator/multibox_target-inl.h:173-185  grad(cls_pred) = 0; anchor and label get no
                          gradient at all (DeclareBackwardDependency returns nothing, :251-256)
  * MultiBoxDetection  -- operator/multibox_detection-inl.h:109-125  grad(cls_prob) = grad(loc_pred) = grad(anchor) = 0
so a training graph that contains them (symbol/symbol_builder.py:73-94) back-propagates nothing through them.  These
functions give a torch graph the same behaviour: outputs are produced by libdspmb, gradients are zero tensors.
"""
import torch

from . import ops


def _zeros(t, needed):
    return torch.zeros_like(t) if needed else None


class _PriorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, kwargs):
        ctx.save_for_backward(data)
        return ops.MultiBoxPrior(data, **kwargs)

    @staticmethod
    def backward(ctx, grad_out):
        (data,) = ctx.saved_tensors
        return _zeros(data, ctx.needs_input_grad[0]), None


class _TargetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, label, cls_pred, kwargs):
        ctx.save_for_backward(cls_pred)
        return tuple(ops.MultiBoxTarget(anchor, label, cls_pred, **kwargs))

    @staticmethod
    def backward(ctx, *grads):
        (cls_pred,) = ctx.saved_tensors
        return None, None, _zeros(cls_pred, ctx.needs_input_grad[2]), None


class _DetectionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_prob, loc_pred, anchor, kwargs):
        ctx.save_for_backward(cls_prob, loc_pred, anchor)
        return ops.MultiBoxDetection(cls_prob, loc_pred, anchor, **kwargs)

    @staticmethod
    def backward(ctx, grad_out):
        cls_prob, loc_pred, anchor = ctx.saved_tensors
        need = ctx.needs_input_grad
        return _zeros(cls_prob, need[0]), _zeros(loc_pred, need[1]), _zeros(anchor, need[2]), None


def multibox_prior(data, **kwargs):
    """MultiBoxPrior on a CUDA tensor inside a torch graph; d(out)/d(data) = 0."""
    return _PriorFn.apply(data, kwargs)


def multibox_target(anchor, label, cls_pred, **kwargs):
    """MultiBoxTarget inside a torch graph -> (loc_target, loc_mask, cls_target); d/d(cls_pred) = 0."""
    for k in ("return_match", "return_stats"):
        if kwargs.get(k):
            raise ValueError("autograd.multibox_target returns the three reference outputs only (%s is not supported)" % k)
    return _TargetFn.apply(anchor, label, cls_pred, kwargs)


def multibox_detection(cls_prob, loc_pred, anchor, **kwargs):
    """MultiBoxDetection inside a torch graph; all three input gradients are zero."""
    if kwargs.get("return_valid_count"):
        raise ValueError("autograd.multibox_detection returns the reference output only")
    return _DetectionFn.apply(cls_prob, loc_pred, anchor, kwargs)
