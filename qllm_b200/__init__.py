"""qllm_b200 -- B200-native fused dequant-matmul engine behind QLLM's QuantLinear.forward.

Importing this package loads libb200q.so (hand-written sm_100a CUDA behind a C ABI, include/b200q.h);
if the library has not been built the import raises: there is no CPU or PyTorch fallback.
"""
from ._lib import lib, check, Layer  # noqa: F401  (raises ImportError when the .so is missing)
from .q_layers import (QuantLinearGPTQ, QuantLinearHQQ, QuantLinearMarlin, QuantLinearORT, WQLinear_GEMM, WQLinear_GEMV,  # noqa: F401
                       fuse_siblings, fused_mlp, linear_group, make_mixbits_quant_linear, select_quant_linear)

from .chain import DecodeChain  # noqa: F401,E402
from .repack import convert_layer, repack_to_new_mode  # noqa: F401,E402
from .loader import from_quantized, load_quant_config, save_quantized  # noqa: F401,E402  (qllm --load for this engine)

__version__ = "0.1.0"

# tuning / diagnostic switches for A/B runs: B200Q_OPTS="chain_window=6,chain_slots=8" (b200q_debug_set_option names)
import os as _os
for _kv in filter(None, _os.environ.get("B200Q_OPTS", "").split(",")):
    _k, _, _v = _kv.partition("=")
    check(lib.b200q_debug_set_option(_k.strip().encode(), float(_v)), "B200Q_OPTS " + _kv)


def patch_qllm():
    """Route an installed `qllm` through this engine: replaces the one dispatch function the reference
    uses to pick its QuantLinear class (qllm/utils/modelutils.py:44).  See INTEGRATION.md."""
    import qllm.utils.modelutils as mu
    mu.select_quant_linear = select_quant_linear
    mu.make_mixbits_quant_linear = make_mixbits_quant_linear
    return mu
