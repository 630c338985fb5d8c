"""normalisation constants of the reference (diffuser/datasets/__init__.py:8-61)"""
import numpy as np

from .._overlay import extend as _extend, lazy_getattr as _lazy

_extend(__path__, "datasets", reference_first=True)   # overlay: the reference's own normalizers / preprocessing (its callers use their signatures)
from .normalization import LimitsNormalizer, DatasetNormalizer
__getattr__ = _lazy(__name__, ("data_api", "sequence", "preprocessing", "buffer", "normalization"))

_K7 = [2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 3.05432619099]
KUKA7D_MIN = -np.array(_K7, dtype=np.float32)
KUKA7D_MAX = np.array(_K7, dtype=np.float32)
DUALKUKA14D_MIN = -np.array(_K7 + _K7, dtype=np.float32)
DUALKUKA14D_MAX = np.array(_K7 + _K7, dtype=np.float32)
RAND_MAZE_MIN = np.array([0., 0.], dtype=np.float32)
RAND_MAZESIZE_55 = np.array([5., 5.], dtype=np.float32)
