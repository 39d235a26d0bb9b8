"""B200-native fused view-synthesis loss for TripleDNet / FeatDepth / monodepth2 training.

Product path: hand-written sm_100a CUDA kernels in ``libtdl.so`` (C ABI: include/tdl.h),
bound with ctypes (``_lib``), wrapped as autograd Functions (``ops``) and exposed behind the
reference's ``compute_losses`` / ``generate_images_pred`` / ``generate_features_pred`` methods
(``losses``, ``nets``).  See DESIGN.md.
"""
from . import _lib, config, geometry, input_pipeline, losses, ops, registry, synth  # noqa: F401
from .config import Config  # noqa: F401
from .input_pipeline import GpuInputPipeline  # noqa: F401
from .losses import ViewSynthesisLossMixin  # noqa: F401
from .registry import MONO  # noqa: F401
