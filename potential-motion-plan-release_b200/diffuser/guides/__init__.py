from .._overlay import extend as _extend

_extend(__path__, "guides", reference_first=False)   # planners / helpers the mirror does not re-implement
from .policies import Policy, Trajectories
from .policies_compose import PolicyCompose
from .rm2d_colchk import RM2DCollisionChecker
from .kuka_colchk import KukaCollisionChecker
from .kuka_replan import KukaDiffusionReplanner
