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
    # the product mirror also defines a top-level `diffuser`: park its modules while the reference is imported
    # and put them back afterwards (the reference classes keep their own module globals), so that both can be
    # exercised by one test process in any order
    parked = {k: sys.modules.pop(k) for k in list(sys.modules)
              if (k == "diffuser" or k.startswith("diffuser.")) and
              not (getattr(sys.modules[k], "__file__", None) or "").startswith(REF_ROOT)}
    sys.path.insert(0, REF_ROOT)
    try:
        for n in ("colorama", "matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(n, MagicMock())
        import diffuser  # noqa: empty __init__
        import torch

        u = types.ModuleType("diffuser.utils")
        u.__path__ = [os.path.join(REF_ROOT, "diffuser", "utils")]
        u.print_color = lambda s, *a, **k: None
        u.Progress = u.Silent = u.Config = _Silent
        u.to_np = lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else x      # utils/arrays.py:13-16
        u.to_torch = lambda x, dtype=None, device=None: torch.as_tensor(x, dtype=dtype or torch.float32, device=device or "cpu")
        sys.modules["diffuser.utils"] = u
        diffuser.utils = u
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
    finally:
        sys.path.remove(REF_ROOT)
        # leave no reference module behind under the shared name: `quiet()` re-installs them around reference calls
        _loaded["ref_modules"] = {k: sys.modules.pop(k) for k in list(sys.modules)
                                  if k == "diffuser" or k.startswith("diffuser.")}
        sys.modules.update(parked)
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


def _import_ref(*names):
    """imports reference modules by name; third-party modules they import but the Maze2D path never touches
    (imageio ...) are mocked as they come up"""
    load_models(); load_maze2d_geometry()
    import importlib
    with quiet():
        sys.path.insert(0, REF_ROOT)
        try:
            for _ in range(30):
                try:
                    return [importlib.import_module(n) for n in names]
                except ModuleNotFoundError as e:
                    if e.name.startswith(("diffuser", "pb_diff_envs")):
                        raise
                    sys.modules[e.name] = MagicMock()
        finally:
            sys.path.remove(REF_ROOT)
    raise RuntimeError("could not import %s" % (names,))


def load_replanner():
    """Returns (KukaDiffusionReplanner, RM2DCollisionChecker, LimitsNormalizer) of the reference (kuka_replan.py,
    rm2d_colchk.py, datasets/normalization.py)"""
    if "replan" not in _loaded:
        rp, cc, nm = _import_ref("diffuser.guides.kuka_replan", "diffuser.guides.rm2d_colchk",
                                 "diffuser.datasets.normalization")
        _loaded["replan"] = (rp.KukaDiffusionReplanner, cc.RM2DCollisionChecker, nm.LimitsNormalizer)
    return _loaded["replan"]


def load_composed_planner():
    """Returns ComposedEnvPlanner of the reference (comp_plan_env.py)"""
    if "comp_planner" not in _loaded:
        _loaded["comp_planner"] = _import_ref("diffuser.guides.comp_plan_env")[0].ComposedEnvPlanner
    return _loaded["comp_planner"]


class quiet:
    """context manager for calls into the reference: silences its print() chatter and, for the duration of the
    call, makes `diffuser` resolve to the reference modules again (lazy imports inside reference code)"""

    def __enter__(self):
        import contextlib, io
        self._cm = contextlib.redirect_stdout(io.StringIO())
        self._cm.__enter__()
        ref = _loaded.get("ref_modules", {})
        self._parked = {k: sys.modules.pop(k) for k in list(sys.modules)
                        if (k == "diffuser" or k.startswith("diffuser.")) and sys.modules[k] is not ref.get(k)}
        sys.modules.update(ref)
        return self

    def __exit__(self, *a):
        ref = _loaded.get("ref_modules", {})
        for k in [k for k in sys.modules if k == "diffuser" or k.startswith("diffuser.")]:
            if k not in ref:
                ref[k] = sys.modules[k]      # lazily imported reference submodule
            del sys.modules[k]
        sys.modules.update(self._parked)
        return self._cm.__exit__(*a)
