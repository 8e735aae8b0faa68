"""Import surface of `librosa` for the unchanged reference package (audiodiffusion/__init__.py:5 imports
librosa.beat.beat_track; scripts/train_unet.py:22 imports librosa.util.normalize). The mel arithmetic itself is NOT
routed through this shim: use `audio_diffusion_b200.Mel` (exposed as `diffusers.Mel`), whose kernels replace
librosa.feature.melspectrogram / power_to_db / db_to_power / feature.inverse.mel_to_audio (mel.py:145-167)."""
from . import beat, util  # noqa: F401

__version__ = "0.10.2+b200shim"


def _moved(name):
    def f(*a, **k):
        raise NotImplementedError(f"librosa.{name}: use audio_diffusion_b200.Mel (CUDA) — no CPU path in this build")
    return f


load = _moved("load")
power_to_db = _moved("power_to_db")
db_to_power = _moved("db_to_power")
