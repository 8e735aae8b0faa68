from audio_diffusion_b200.mel import Mel  # noqa: F401  (scripts/train_unet.py:19, scripts/audio_to_images.py:10)
