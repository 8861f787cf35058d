"""Per-image scale parameters (mirror of careless/models/scaling/image.py:9-63)."""
import numpy as np

from .nn import Scaler


class ImageScaler(Scaler):
    def __init__(self, max_images):
        self.max_images = int(max_images)
        self._scales = np.ones(self.max_images - 1, dtype=np.float32)      # image.py:21

    @property
    def scales(self):
        return np.concatenate(([1.], self._scales)).astype(np.float32)       # image.py:23-25


class HybridImageScaler(Scaler):
    def __init__(self, mlp_scaler, image_scaler):
        self.mlp_scaler = mlp_scaler
        self.image_scaler = image_scaler


class NeuralImageScaler(Scaler):
    """image.py:98-125 (per-image dense layers, --image-layers).  Not built yet in the CUDA path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("NeuralImageScaler (--image-layers > 0) is not implemented in careless_b200 yet")
