"""B200-native (sm_100a) implementation of the elektronn3 UNet / Predictor hot path.

``UNet`` is a drop-in for ``elektronn3.models.unet.UNet``, ``resunet.UNet`` for ``elektronn3.models.resunet.UNet`` and
``Predictor`` for
``elektronn3.inference.Predictor``; the arithmetic runs in hand-written CUDA kernels (libe3b.so, C ABI in
include/e3b.h).  There is no CPU or cuDNN fallback.
"""
from .unet import UNet  # noqa: F401
from .inference import Predictor  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .loss import DiceLoss  # noqa: F401
from . import resunet  # noqa: F401

__all__ = ['UNet', 'Predictor', 'GraphedTrainStep', 'DiceLoss', 'resunet']
