"""npvp_b200 - B200-native (sm_100a) implementation of the NPVP inference hot path.

Public surface mirrors the reference (XiYe20/NPVP): ``ResnetEncoder``, ``ResnetDecoder``, ``Predictor``
with identical constructors, forward signatures and state_dict keys; plus ``NPVPInference`` (a
Lightning-free stand-in for ``LitPredictor.forward``) and ``load_config`` for the reference YAMLs.
"""
from .autoencoder import ResnetEncoder, ResnetDecoder
from .predictor import Predictor

__all__ = ["ResnetEncoder", "ResnetDecoder", "Predictor", "NPVPInference", "load_config", "build_from_config"]
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("NPVPInference", "build_from_config"):
        from . import pipeline
        return getattr(pipeline, name)
    if name == "load_config":
        from .config import load_config
        return load_config
    raise AttributeError(name)
