"""Drop-in mirror of the reference's `diffuser` package for the planning hot path only.
Same module paths / class names / constructor kwargs / state-dict keys as the reference, so
its pickled configs and checkpoints load unchanged; the compute goes to libpmp_b200.so."""

from ._overlay import extend as _extend

_extend(__path__, "", reference_first=False)     # un-mirrored sub-packages / modules come from the reference checkout
