"""Drop-in model classes at the reference's dotted paths (diffuser.models.*): the pickled
`Config._class` objects of the reference resolve to these."""
from .._overlay import extend as _extend

_extend(__path__, "models", reference_first=False)   # e.g. diffuser.models.vit_vanilla for code that imports it directly
from .diffusion_pb import GaussianDiffusionPB
from .temporal_cond import TemporalUnet_WCond

__all__ = ["TemporalUnet_WCond", "GaussianDiffusionPB"]
