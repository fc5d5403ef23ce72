// Intentionally empty: stands in for MXNet's src/operator/mshadow_op.h, which operator/multibox_target.cc:27
// includes as "../mshadow_op.h" but does not use on the CPU path.
