"""Drop-in model classes at the reference's dotted paths (diffuser.models.*): the pickled
`Config._class` objects of the reference resolve to these."""
from .diffusion_pb import GaussianDiffusionPB
from .temporal_cond import TemporalUnet_WCond

__all__ = ["TemporalUnet_WCond", "GaussianDiffusionPB"]
