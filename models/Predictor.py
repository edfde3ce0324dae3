"""Same import path as the reference's models/Predictor.py; the class lives in npvp_b200.predictor."""
from npvp_b200.predictor import Predictor
from npvp_b200.pipeline import NPVPInference as LitPredictorInference

__all__ = ["Predictor", "LitPredictorInference"]
