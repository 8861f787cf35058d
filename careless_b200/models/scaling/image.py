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

    def save_weights(self, path):
        self.mlp_scaler.save_weights(str(path) + "_mlp")
        np.savez(str(path) + "_image.npz", scales=self.image_scaler._scales)

    def load_weights(self, path):
        self.mlp_scaler.load_weights(str(path) + "_mlp")
        self.image_scaler._scales = np.load(str(path) + "_image.npz")["scales"].astype(np.float32)


class ImageLayer(Scaler):
    """image.py:66-96: a dense layer whose kernel (units, in) and bias depend on the image; identity / zero init."""

    def __init__(self, units, max_images):
        self.units, self.max_images = int(units), int(max_images)
        self.w = np.tile(np.eye(self.units, dtype=np.float32), (self.max_images, 1, 1))     # (n_images, units, in)
        self.b = np.zeros((self.max_images, self.units), dtype=np.float32)


class NeuralImageScaler(Scaler):
    """image.py:98-125 (--image-layers): MLP(metadata) -> per-image dense layers -> Dense(2) -> Normal.
    The per-image products run in the CUDA observation kernel on image-major rows (one image per tile)."""

    def __init__(self, image_layers, max_images, mlp_layers, mlp_width, leakiness=0.01, epsilon=1e-7, scale_bijector=None,
                 scale_multiplier=None):
        from .nn import MetadataScaler
        self.metadata_scaler = MetadataScaler(mlp_layers, mlp_width, leakiness, epsilon=epsilon, scale_bijector=scale_bijector,
                                              scale_multiplier=scale_multiplier)
        self.max_images = int(max_images)
        self.image_layers = [ImageLayer(mlp_width, max_images) for _ in range(int(image_layers))]

    def save_weights(self, path):
        self.metadata_scaler.save_weights(str(path) + "_mlp")
        np.savez(str(path) + "_image_layers.npz", flat=self.flat())

    def load_weights(self, path):
        self.metadata_scaler.load_weights(str(path) + "_mlp")
        self.from_flat(np.load(str(path) + "_image_layers.npz")["flat"])

    def flat(self):
        return np.concatenate([np.concatenate([l.w.reshape(-1), l.b.reshape(-1)]) for l in self.image_layers]) \
            if self.image_layers else np.zeros(0, dtype=np.float32)

    def from_flat(self, flat):
        off = 0
        for l in self.image_layers:
            l.w = np.asarray(flat[off:off + l.w.size], dtype=np.float32).reshape(l.w.shape); off += l.w.size
            l.b = np.asarray(flat[off:off + l.b.size], dtype=np.float32).reshape(l.b.shape); off += l.b.size
