"""nasrec_b200 -- B200-native (sm_100a) implementation of NASRec's supernet hot path.

Drop-in for ``nasrec.supernet`` (SuperNet / SuperNetBlock / modules / config JSON);
see INTEGRATION.md.  The arithmetic is hand-written CUDA behind the C ABI in
``include/nasrec_b200.h`` -- there is no CPU or eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .supernet.supernet import SuperNet, SuperNetBlock, ops_config_lib, path_sampling_strategy_lib  # noqa: F401

__version__ = "0.2.0"


class precision:
    """``with nasrec_b200.precision("bf16"): ...`` -- GEMM arithmetic of everything run inside (process-wide switch):
    "fp32" = 3xTF32-split tensor-core products (fp32 parity, the default), "bf16" = the reference's mixed-precision
    path (train_utils.py:146,247-286 ``use_amp``): GEMM operands rounded to bfloat16, one product per k-step, fp32
    accumulation; master weights, Adagrad state, activations and all non-GEMM kernels stay fp32.  Stated tolerance
    (tests/test_gpu_bf16.py): logits within 2e-2 (RMS, relative) of the reference under torch.autocast(bfloat16)."""
    MODES = {"fp32": 3, "bf16": 2, "tf32": 1, "ffma": 0}

    def __init__(self, name: str):
        if name not in self.MODES:
            raise ValueError("precision must be one of %s" % sorted(self.MODES))
        self.mode = self.MODES[name]

    def __enter__(self):
        self.prev = _lib.LIB.gemm_mode()
        _lib.LIB.set_gemm_mode(self.mode)
        return self

    def __exit__(self, *exc):
        _lib.LIB.set_gemm_mode(self.prev)


def set_precision(name: str):
    """Process-wide GEMM arithmetic: see ``precision``."""
    _lib.LIB.set_gemm_mode(precision.MODES[name])


def install_as_nasrec():
    """Alias this package as ``nasrec`` so reference entry scripts import it unchanged
    (``from nasrec.supernet.supernet import SuperNet`` ...)."""
    import sys
    from . import supernet as _sn, utils as _ut
    from .supernet import modules as _m, supernet as _s, utils as _u
    from .utils import config as _c, data_pipes as _dp, io_utils as _io, lr_schedule as _lr, train_utils as _tu
    from . import search as _se
    sys.modules.setdefault("nasrec", sys.modules[__name__])
    sys.modules.setdefault("nasrec.supernet", _sn)
    sys.modules.setdefault("nasrec.supernet.supernet", _s)
    sys.modules.setdefault("nasrec.supernet.modules", _m)
    sys.modules.setdefault("nasrec.supernet.utils", _u)
    sys.modules.setdefault("nasrec.utils", _ut)
    sys.modules.setdefault("nasrec.utils.config", _c)
    sys.modules.setdefault("nasrec.utils.lr_schedule", _lr)
    sys.modules.setdefault("nasrec.utils.io_utils", _io)
    sys.modules.setdefault("nasrec.utils.train_utils", _tu)
    sys.modules.setdefault("nasrec.utils.data_pipes", _dp)       # transform half only (VanillaTransform*)
    sys.modules.setdefault("nasrec.searcher", _se)
    sys.modules.setdefault("nasrec.searcher.tokenizer", _se)     # Tokenizer
    sys.modules.setdefault("nasrec.searcher.searcher", _se)      # Searcher (in-process, GPU-resident)
