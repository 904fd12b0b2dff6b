from mixdq_b200.nn.linear import QuantizedLinear  # noqa: F401
