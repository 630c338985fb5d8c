"""The few `diffuser.utils` symbols the hot path and its callers touch
(reference: diffuser/utils/{arrays,config,progress}.py)."""
from .arrays import to_np, to_torch, to_device, apply_dict
from .config import Config


def print_color(s, *args, c='r', **kwargs):
    print(s)


class Silent:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, attr):
        return lambda *args, **kwargs: None


Progress = Silent
