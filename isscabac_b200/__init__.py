"""isscabac_b200 -- B200-native many-stream CABAC engine behind the API surface of
christianrohlfing/ISScabac (the HM-derived CABAC_ArithmeticEncoder/Decoder, the MEX command
set and the MATLAB binarizer / context-selection layer).

Layout: csrc/ (CUDA kernels + C ABI, built into libisscabac.so), mex/ (the MATLAB mexFunction
shim), engine.py (device-tensor and host-buffer front ends), matlab_api.py (cabacWrapper /
SimpleCABACMex / cabacBinarizer ... mirrors of the MATLAB layer), coder.py (cabacEncode /
cabacDecode drivers of the ISS application), quantizer.py (quantizeWrapper: the dead-zone /
Lloyd-Max quantiser that produces the coder's symbols), multi_gpu.py (stream sharding over torch.distributed).
"""
from . import build as _build  # noqa: F401
from ._lib import CabacError, SymCfg, lib  # noqa: F401
from .engine import *  # noqa: F401,F403
