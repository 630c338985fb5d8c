from .temporal_cond import TemporalUnet_WCond
from .diffusion_pb import GaussianDiffusionPB
