"""The few `diffuser.utils` symbols the hot path and its callers touch
(reference: diffuser/utils/{arrays,config,progress}.py)."""
from .._overlay import extend as _extend, lazy_getattr as _lazy

# overlay mode: the reference's utils modules (arrays, config, serialization, setup, training_pbdiff ...) are used as
# they are -- none of them is on the hot path -- and the symbols its __init__ star-imports resolve lazily
_OVERLAY = _extend(__path__, "utils", reference_first=True)
from .arrays import to_np, to_torch, to_device, apply_dict
from .config import Config
__getattr__ = _lazy(__name__, ("serialization", "train_utils", "progress", "setup", "config", "rendering", "arrays", "colab",
                                "eval_utils", "training_pbdiff", "rm2d_render", "luo_utils"))


def print_color(s, *args, c='r', **kwargs):
    print(s)


class Silent:
    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, attr):
        return lambda *args, **kwargs: None


Progress = Silent
