"""Mirror of reference utils/op/upfirdn2d.py: same public names."""
from transeditor_b200.op import UpFirDn2d, UpFirDn2dBackward, upfirdn2d  # noqa: F401
