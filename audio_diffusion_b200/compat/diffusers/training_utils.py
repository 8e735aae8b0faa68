"""EMAModel as used at scripts/train_unet.py:185-190, :265-266, :292-301 — the engine's implementation."""
from audio_diffusion_b200.training import EMAModel  # noqa: F401
