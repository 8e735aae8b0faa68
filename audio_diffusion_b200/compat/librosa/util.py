import numpy as np


def normalize(S, norm=np.inf, axis=0):
    """librosa.util.normalize (default: max-abs along axis), used at scripts/train_unet.py:345 for logging only."""
    S = np.asarray(S)
    mag = np.abs(S).astype(float)
    length = np.max(mag, axis=axis, keepdims=True) if norm == np.inf else np.sum(mag ** norm, axis=axis, keepdims=True) ** (1.0 / norm)
    length[length < np.finfo(float).tiny] = 1.0
    return S / length
