"""Hand-written sm_100a ops with autograd wrappers (kernels in csrc/kernels)."""
from .bn_act import bn_act, fused_bn_available  # noqa: F401
