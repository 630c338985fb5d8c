"""Import shim that makes the *reference* model / sampler / Maze2D geometry
importable on CPU inside the build container.

Only usable where /root/reference exists (never on the GPU box).  Used by
tests/golden/gen_golden.py to freeze golden vectors and by the `not gpu`
tests that cross-check the oracle restatement against the live reference.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("PMP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "diffuser", "models"))


class _Silent:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


_loaded = {}


def load_models():
    """Returns (TemporalUnet_WCond, GaussianDiffusionPB) of the reference."""
    if "models" in _loaded:
        return _loaded["models"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    # the product mirror also defines a top-level `diffuser`; the two must
    # never be mixed inside one interpreter
    if "diffuser" in sys.modules and not getattr(sys.modules["diffuser"], "__file__", "").startswith(REF_ROOT):
        raise RuntimeError("a non-reference `diffuser` package is already imported")
    sys.path.insert(0, REF_ROOT)
    for n in ("colorama", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, MagicMock())
    import diffuser  # noqa: empty __init__

    u = types.ModuleType("diffuser.utils")
    u.__path__ = [os.path.join(REF_ROOT, "diffuser", "utils")]
    u.print_color = lambda s, *a, **k: None
    u.Progress = u.Silent = _Silent
    u.to_np = lambda x: x.detach().cpu().numpy()
    sys.modules["diffuser.utils"] = u
    diffuser.utils = u
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
    _loaded["models"] = (TemporalUnet_WCond, GaussianDiffusionPB)
    return _loaded["models"]


def load_maze2d_geometry():
    """Returns (Point2DRobot, RectangleWall, RectangleWallGroup, AbsStaticMaze2DEnv)
    of the reference with the package __init__ side effects (gym registration,
    pybullet) bypassed."""
    if "geom" in _loaded:
        return _loaded["geom"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    root = os.path.join(REF_ROOT, "pb_diff_envs", "pb_diff_envs")
    for n in ("pybullet", "pybullet_data", "colorama", "matplotlib", "matplotlib.pyplot",
              "matplotlib.patches", "gym", "gym.spaces", "gym.utils"):
        sys.modules.setdefault(n, MagicMock())
    for sub in ("", ".environment", ".robot", ".objects", ".utils"):
        name = "pb_diff_envs" + sub
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(root, *sub.strip(".").split("."))] if sub else [root]
            sys.modules[name] = m
    from pb_diff_envs.robot.point2d_robot import Point2DRobot
    from pb_diff_envs.environment.rand_rec_group import RectangleWall, RectangleWallGroup
    from pb_diff_envs.environment.abs_maze2d_env import AbsStaticMaze2DEnv
    _loaded["geom"] = (Point2DRobot, RectangleWall, RectangleWallGroup, AbsStaticMaze2DEnv)
    return _loaded["geom"]


class quiet:
    """context manager silencing the reference's print() chatter"""

    def __enter__(self):
        import contextlib, io
        self._cm = contextlib.redirect_stdout(io.StringIO())
        self._cm.__enter__()
        return self

    def __exit__(self, *a):
        return self._cm.__exit__(*a)
