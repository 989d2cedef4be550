"""Drop-in `utils.op` (reference utils/op/__init__.py:1-2): the same three symbols, backed by
libte_b200.so instead of two JIT-built torch extensions.  `utils` itself stays a namespace
package, so `utils.sample`, `utils.distributed` ... still resolve from the TransEditor tree."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu  # noqa: F401
from .upfirdn2d import upfirdn2d  # noqa: F401
