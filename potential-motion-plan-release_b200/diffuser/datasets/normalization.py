"""LimitsNormalizer / DatasetNormalizer with constant limits, as the planners use them
(reference: diffuser/datasets/normalization.py:137-162 and :13-90, norm_const_dict path)."""
import numpy as np


class LimitsNormalizer:
    """maps [min, max] to [-1, 1]"""

    def __init__(self, mins, maxs):
        self.mins = np.asarray(mins, np.float32)
        self.maxs = np.asarray(maxs, np.float32)

    def normalize(self, x):
        x = (x - self.mins) / (self.maxs - self.mins)
        return 2 * x - 1

    def unnormalize(self, x, eps=0):
        if x.max() > 1 + eps or x.min() < -1 - eps:
            x = np.clip(x, -1, 1)
        x = (x + 1) / 2.
        return x * (self.maxs - self.mins) + self.mins


class DatasetNormalizer:
    """per-key normalizers built from constants: {'observations': (mins, maxs), ...}"""

    def __init__(self, norm_const_dict, action_dim=0):
        self.normalizers = {k: LimitsNormalizer(*v) for k, v in norm_const_dict.items()}
        self.observation_dim = len(self.normalizers['observations'].mins)
        self.action_dim = action_dim

    def normalize(self, x, key):
        return self.normalizers[key].normalize(x)

    def unnormalize(self, x, key):
        return self.normalizers[key].unnormalize(x)
