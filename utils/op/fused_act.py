"""Mirror of reference utils/op/fused_act.py: same public names."""
from transeditor_b200.op import (  # noqa: F401
    FusedLeakyReLU, FusedLeakyReLUFunction, FusedLeakyReLUFunctionBackward, fused_leaky_relu,
)
