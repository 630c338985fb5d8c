"""numpy <-> torch helpers (reference interface: diffuser/utils/arrays.py:13-68)."""
import numpy as np
import torch

DTYPE = torch.float
DEVICE = 'cuda:0'


def to_np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else x


def to_torch(x, dtype=None, device=None):
    dtype = dtype or DTYPE
    device = device or DEVICE
    if isinstance(x, dict):
        return {k: to_torch(v, dtype, device) for k, v in x.items()}
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype)
    if isinstance(x, list) and len(x) and torch.is_tensor(x[0]):
        return torch.stack(x, dim=0).to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


def to_device(x, device=DEVICE):
    if isinstance(x, dict):
        return {k: to_device(v, device) for k, v in x.items()}
    return x.to(device)


def apply_dict(fn, d, *args, **kwargs):
    return {k: fn(v, *args, **kwargs) for k, v in d.items()}
