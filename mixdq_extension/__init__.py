"""Drop-in alias of the reference's `mixdq_extension` package (kernels/mixdq_extension):
the same module paths — `_C`, `op.quant`, `op.qlinear`, `op.qconv2d`, `nn.Linear`, `nn.Conv2d`,
`nn.utils` — resolved onto the B200-native implementation in `mixdq_b200`."""
from . import _C  # noqa: F401
