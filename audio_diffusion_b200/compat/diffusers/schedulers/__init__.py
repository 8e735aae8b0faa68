from audio_diffusion_b200.schedulers import DDIMScheduler, DDPMScheduler  # noqa: F401
