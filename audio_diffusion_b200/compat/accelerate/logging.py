"""accelerate.logging.get_logger (scripts/train_unet.py:27): a stdlib logger that only speaks on the main process."""
import logging
import os


class _MainProcessAdapter(logging.LoggerAdapter):
    def log(self, level, msg, *args, main_process_only=True, **kwargs):
        if main_process_only and int(os.environ.get("RANK", "0")) != 0:
            return
        super().log(level, msg, *args, **kwargs)


def get_logger(name: str, log_level: str = None):
    logger = logging.getLogger(name)
    if log_level is not None:
        logger.setLevel(log_level.upper())
    return _MainProcessAdapter(logger, {})
