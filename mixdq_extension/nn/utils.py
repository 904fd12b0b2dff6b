from mixdq_b200.nn.utils import *  # noqa: F401,F403
from mixdq_b200.nn.utils import QParam, create_qparams_from_dtype, get_quant_para, dtype_to_bw  # noqa: F401
