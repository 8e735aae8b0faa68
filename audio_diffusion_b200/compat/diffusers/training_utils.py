"""EMAModel as used at scripts/train_unet.py:185-190,265-266,292-301 (inv_gamma / power / max_value schedule)."""
import copy

import torch


class EMAModel:
    def __init__(self, model, update_after_step=0, inv_gamma=1.0, power=2 / 3, min_value=0.0, max_value=0.9999, **kw):
        self.averaged_model = copy.deepcopy(model).eval()
        self.averaged_model.requires_grad_(False)
        self.update_after_step = update_after_step
        self.inv_gamma, self.power, self.min_value, self.max_value = inv_gamma, power, min_value, max_value
        self.decay = 0.0
        self.optimization_step = 0

    def get_decay(self, optimization_step):
        step = max(0, optimization_step - self.update_after_step - 1)
        value = 1 - (1 + step / self.inv_gamma) ** -self.power
        if step <= 0:
            return 0.0
        return max(self.min_value, min(value, self.max_value))

    @torch.no_grad()
    def step(self, new_model):
        self.decay = self.get_decay(self.optimization_step)
        ema = dict(self.averaged_model.named_parameters())
        for name, p in new_model.named_parameters():
            if p.requires_grad:
                ema[name].sub_((1 - self.decay) * (ema[name] - p.to(ema[name].device)))
            else:
                ema[name].copy_(p)
        self.optimization_step += 1
