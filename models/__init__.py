"""Drop-in package path: ``from models.Predictor import Predictor`` etc. resolve to the B200 implementation."""
from .ResNetAutoEncoder import ResnetEncoder, ResnetDecoder
from .Predictor import Predictor
from npvp_b200.layers import CoorGenerator, NRMLP, PosFeatFuser, EventEncoder, VidHRformerDecoderNAR, VidHRFormerEncoder

__all__ = ["ResnetEncoder", "ResnetDecoder", "Predictor", "CoorGenerator", "NRMLP", "PosFeatFuser", "EventEncoder",
           "VidHRformerDecoderNAR", "VidHRFormerEncoder"]
