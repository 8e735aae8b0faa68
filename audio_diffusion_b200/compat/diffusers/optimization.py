"""get_scheduler as called at scripts/train_unet.py:174-179 (cosine / linear / constant with warm-up)."""
import math

from torch.optim.lr_scheduler import LambdaLR


def get_scheduler(name, optimizer, num_warmup_steps=0, num_training_steps=None):
    name = str(name)

    def warm(step):
        return float(step) / float(max(1, num_warmup_steps)) if step < num_warmup_steps else None

    if name == "constant":
        return LambdaLR(optimizer, lambda s: 1.0)
    if name == "constant_with_warmup":
        return LambdaLR(optimizer, lambda s: warm(s) if warm(s) is not None else 1.0)
    if name == "linear":
        def f(s):
            w = warm(s)
            if w is not None:
                return w
            return max(0.0, float(num_training_steps - s) / float(max(1, num_training_steps - num_warmup_steps)))
        return LambdaLR(optimizer, f)
    if name == "cosine":
        def f(s):
            w = warm(s)
            if w is not None:
                return w
            p = float(s - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 2.0 * 0.5 * p)))
        return LambdaLR(optimizer, f)
    raise ValueError(f"unsupported lr scheduler {name}")
