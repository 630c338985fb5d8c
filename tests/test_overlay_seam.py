"""The Python seam of the drop-in: the reference's own pickled Configs resolve to the mirror's
classes, and with PMP_REFERENCE_ROOT set the mirror OVERLAYS the reference checkout -- un-mirrored modules
(utils.serialization, guides.plan_pb_diff_helper, rm2d_plan_env ...) come from the reference, hot-path modules from
the mirror -- so the reference's planners run on top of the B200 classes unmodified.  CPU only (nothing is computed)."""
import os
import pickle
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "potential-motion-plan-release_b200")
LOGDIR = os.path.join(ROOT, "tests", "golden", "logdir_m2d")
REF = os.environ.get("PMP_REFERENCE_ROOT", "/root/reference")


def _run(code, env_extra):
    env = dict(os.environ)
    env.pop("PMP_REFERENCE_ROOT", None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_reference_written_config_pickles_resolve_to_the_mirror_classes():
    """tests/golden/logdir_m2d/*.pkl were written by the REFERENCE's Config class with the reference's model classes
    (tests/golden/gen_golden.py:gen_logdir, like scripts/train.py:64-102); stand-alone mirror, no reference on the path"""
    out = _run("""
        import sys, pickle
        sys.path.insert(0, %r)
        import diffuser.models as M, diffuser.utils as U
        mc = pickle.load(open(%r + "/model_config.pkl", "rb"))
        dc = pickle.load(open(%r + "/diffusion_config.pkl", "rb"))
        assert type(mc) is U.Config and mc._class is M.TemporalUnet_WCond and dc._class is M.GaussianDiffusionPB
        model = mc()
        diffusion = dc(model)
        assert type(model) is M.TemporalUnet_WCond and type(diffusion) == M.GaussianDiffusionPB     # the `==` test of policies.py:86
        assert diffusion.n_timesteps == 100 and model.n_timesteps == 100 and diffusion.horizon == 48
        assert sys.modules[type(model).__module__].__file__.startswith(%r)
        print("ok", len(model.state_dict()))
    """ % (PKG, LOGDIR, LOGDIR, PKG), {})
    assert out.strip().startswith("ok")


# third-party packages the reference's planners import but this image lacks: any import below these names yields a mock
_AUTOMOCK = """
import sys, types, importlib.abc, importlib.machinery
from unittest.mock import MagicMock
_NAMES = ("h5py", "gym", "tap", "termcolor", "colorama", "matplotlib", "imageio", "pybullet", "pybullet_data", "git",
          "mediapy", "timm", "skvideo", "mujoco_py", "d4rl", "transforms3d", "pyquaternion", "omegaconf", "torch_geometric")
class _MockModule(types.ModuleType):
    # CamelCase attributes are real (empty) classes so that reference code can derive from them; the rest are mocks
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (object,), {}) if name[:1].isupper() else MagicMock(name=self.__name__ + "." + name)
        setattr(self, name, v)
        return v
class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _NAMES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
    def create_module(self, spec):
        m = _MockModule(spec.name)
        m.__path__ = []
        return m
    def exec_module(self, module):
        pass
for n in _NAMES:
    try:
        __import__(n)
    except Exception:
        pass
sys.meta_path.append(_Finder())
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "diffuser", "guides")), reason="reference checkout not present")
def test_overlay_runs_the_reference_planner_modules_on_the_mirror_classes():
    out = _run(_AUTOMOCK + """
import sys
sys.path.insert(0, %r + "/pb_diff_envs"); sys.path.insert(0, %r)
import diffuser, diffuser.utils as U, diffuser.models as M, diffuser.guides as G
ref, pkg = %r, %r
f = lambda mod: sys.modules[mod.__module__ if hasattr(mod, "__module__") else mod.__name__].__file__
# un-mirrored modules: the reference's own files
import diffuser.utils.serialization as S, diffuser.utils.config as C, diffuser.datasets.normalization as N
assert S.__file__.startswith(ref) and C.__file__.startswith(ref) and N.__file__.startswith(ref)
assert U.load_config.__module__ == "diffuser.utils.serialization" and U.get_latest_epoch is S.get_latest_epoch
# hot-path modules: the mirror
import diffuser.guides.policies as P, diffuser.guides.policies_compose as PC, diffuser.guides.rm2d_colchk as RC
import diffuser.models.temporal_cond as TC, diffuser.models.diffusion_pb as DP
for m in (P, PC, RC, TC, DP):
    assert m.__file__.startswith(pkg), m.__file__
# the reference's planner front end imports on top of them and binds the mirror's Policy
import diffuser.guides.plan_pb_diff_helper as H
assert H.__file__.startswith(ref) and H.Policy is P.Policy and hasattr(H, "DiffusionPlanner")
import diffuser.guides.rm2d_plan_env as RP
assert RP.__file__.startswith(ref) and hasattr(RP, "RandMaze2DEnvPlanner")
# a reference-written logdir, loaded with the reference's own loader functions, instantiates the mirror's classes
mc = U.load_config(%r, "model_config.pkl"); dc = U.load_config(%r, "diffusion_config.pkl")
assert type(mc) is C.Config
model = mc(); diffusion = dc(model)
assert type(model) is M.TemporalUnet_WCond and type(diffusion) == M.GaussianDiffusionPB
assert f(type(model)).startswith(pkg) and f(type(diffusion)).startswith(pkg)
print("ok")
""" % (REF, PKG, REF, PKG, LOGDIR, LOGDIR), {"PMP_REFERENCE_ROOT": REF})
    assert out.strip().endswith("ok")


def test_diffusion_schedule_length_reaches_the_model():
    """n_timesteps != 100: forward() and the sampling loops must agree on one engine / one time-embedding table"""
    sys.path.insert(0, PKG)
    sys.path.insert(0, ROOT)
    from oracle import synth
    from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
    a = synth.arch("m2d")
    model = TemporalUnet_WCond(**a["unet"])
    assert model.n_timesteps == 100
    kw = dict(a["diffusion"]); kw["n_timesteps"] = 1000
    d = GaussianDiffusionPB(model, **kw)
    assert d.n_timesteps == 1000 and model.n_timesteps == 1000 and d.betas.shape == (1000,)
    ts = d.ddim_set_timesteps(8)
    assert int(ts[0]) == 875 and d.step_coefs_ddim(ts, 0.0).shape == (8, 8)
