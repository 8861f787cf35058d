"""MLP scale model (mirror of careless/models/scaling/nn.py:27-120).  Holds the keras-ordered weights;
forward/backward run fused in the CUDA observation kernel."""
import numpy as np

from ..base import BaseModel


class Scaler(BaseModel):
    trainable = True


class MetadataScaler(Scaler):
    def __init__(self, n_layers, width, leakiness=0.01, epsilon=1e-7, scale_bijector=None, scale_multiplier=None):
        if leakiness != 0.01:
            raise NotImplementedError("the CUDA kernel implements LeakyReLU(0.01), the only value the reference CLI uses")
        self.n_layers, self.width = int(n_layers), width
        self.epsilon = float(epsilon)
        # nn.py:14-18: softplus is the default of direct construction; the CLI passes exp (manager.py:457-463)
        self.scale_bijector = "softplus" if scale_bijector is None else str(scale_bijector).lower()
        if self.scale_bijector not in ("exp", "softplus"):
            raise ValueError(f"Unsupported scale bijector type, {scale_bijector}")
        self.scale_multiplier = scale_multiplier     # nn.py:84-87: additive tfb.Shift
        self.weights = None                          # built on first use (needs the metadata width)

    def build(self, n_meta):
        if self.weights is not None:
            return
        width = self.width if self.width is not None else n_meta
        self.width = int(width)
        ws, fan_in = [], n_meta
        for _ in range(self.n_layers):               # nn.py:55-68 identity kernels, zero bias
            ws += [np.eye(fan_in, width, dtype=np.float32), np.zeros(width, dtype=np.float32)]
            fan_in = width
        ws += [np.eye(fan_in, 2, dtype=np.float32), np.zeros(2, dtype=np.float32)]      # nn.py:72-79
        self.weights = ws

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        if len(ws) != len(self.weights) or any(a.shape != np.shape(b) for a, b in zip(self.weights, ws)):
            raise ValueError("weight shapes do not match the model")
        self.weights = [np.asarray(w, dtype=np.float32).copy() for w in ws]

    def flat(self):
        return np.concatenate([w.reshape(-1) for w in self.weights])

    def from_flat(self, flat):
        off, out = 0, []
        for w in self.weights:
            out.append(np.asarray(flat[off:off + w.size], dtype=np.float32).reshape(w.shape)); off += w.size
        self.weights = out

    def save_weights(self, path):
        np.savez(path if str(path).endswith(".npz") else str(path) + ".npz", *self.weights)

    def load_weights(self, path):
        w = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
        self.set_weights([w[k] for k in sorted(w.files, key=lambda s: int(s.split("_")[1]))])


class MLPScaler(MetadataScaler):
    pass
