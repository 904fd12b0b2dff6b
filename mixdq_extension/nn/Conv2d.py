from mixdq_b200.nn.conv2d import QuantizedConv2d  # noqa: F401
