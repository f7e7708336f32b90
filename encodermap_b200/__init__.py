"""encodermap_b200 -- B200-native (sm_100a) kernels for EncoderMap's training hot path.

The package mirrors the module paths of the reference for the path it replaces:

    encodermap_b200.misc.distances        <- encodermap.misc.distances
    encodermap_b200.misc.backmapping      <- encodermap.misc.backmapping (TF part)
    encodermap_b200.encodermap_tf1        <- encodermap.encodermap_tf1.backmapping
    encodermap_b200.loss_functions        <- encodermap.loss_functions.loss_functions
    encodermap_b200.models.layers         <- encodermap.models.layers

Everything runs in libemk.so (hand-written CUDA, include/emk.h); there is no CPU fallback."""
from . import _lib  # noqa: F401
from ._lib import EmkError  # noqa: F401
from .parameters import ADCParameters, Parameters  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # heavier submodules are imported on first use so that `import encodermap_b200` works without a GPU
    import importlib

    if name in ("misc", "loss_functions", "models", "encodermap_tf1", "parallel", "tf_adapter", "graph", "_ops"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
