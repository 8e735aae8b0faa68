"""accelerate.utils.ProjectConfiguration as constructed at scripts/train_unet.py:44."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class ProjectConfiguration:
    project_dir: Optional[str] = None
    logging_dir: Optional[str] = None
    automatic_checkpoint_naming: bool = False
    total_limit: Optional[int] = None
