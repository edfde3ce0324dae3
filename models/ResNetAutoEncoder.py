"""Same import path as the reference's models/ResNetAutoEncoder.py; classes live in npvp_b200.autoencoder."""
from npvp_b200.autoencoder import ResnetEncoder, ResnetDecoder

__all__ = ["ResnetEncoder", "ResnetDecoder"]
