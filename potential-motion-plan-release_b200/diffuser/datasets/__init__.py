"""normalisation constants of the reference (diffuser/datasets/__init__.py:8-61)"""
import numpy as np

from .normalization import LimitsNormalizer, DatasetNormalizer

_K7 = [2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 2.96705972839, 2.09439510239, 3.05432619099]
KUKA7D_MIN = -np.array(_K7, dtype=np.float32)
KUKA7D_MAX = np.array(_K7, dtype=np.float32)
DUALKUKA14D_MIN = -np.array(_K7 + _K7, dtype=np.float32)
DUALKUKA14D_MAX = np.array(_K7 + _K7, dtype=np.float32)
RAND_MAZE_MIN = np.array([0., 0.], dtype=np.float32)
RAND_MAZESIZE_55 = np.array([5., 5.], dtype=np.float32)
