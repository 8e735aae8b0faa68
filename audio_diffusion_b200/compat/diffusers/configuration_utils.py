"""ConfigMixin / register_to_config as used by audiodiffusion/mel.py:22,58 (config dict + save/load of <config_name>)."""
import functools
import inspect
import json
import os


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class ConfigMixin:
    config_name = None

    def register_to_config(self, **kw):
        self.config = _Cfg(**kw)

    def save_config(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, self.config_name), "w") as f:
            json.dump(dict(self.config), f, indent=2)

    save_pretrained = save_config

    @classmethod
    def from_config(cls, path_or_dict):
        cfg = path_or_dict
        if isinstance(cfg, str):
            with open(os.path.join(cfg, cls.config_name)) as f:
                cfg = json.load(f)
        sig = inspect.signature(cls.__init__).parameters
        return cls(**{k: v for k, v in cfg.items() if k in sig})

    from_pretrained = from_config


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        ConfigMixin.register_to_config(self, **cfg)
        init(self, *args, **kwargs)
    return wrapper
