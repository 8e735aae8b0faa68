from audio_diffusion_b200.pipeline import BaseOutput  # noqa: F401
