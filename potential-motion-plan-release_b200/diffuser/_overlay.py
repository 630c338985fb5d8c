"""Overlay of the mirror package on the reference checkout.

With PMP_REFERENCE_ROOT=/path/to/potential-motion-plan-release set, every `diffuser.*` module this mirror does NOT
re-implement (utils.serialization / setup / training_pbdiff, guides.plan_pb_diff_helper, rm2d_plan_env,
datasets.preprocessing ...) resolves to the reference's own file, while the hot-path modules
(models.*, guides.policies*, guides.*_colchk, guides.kuka_replan) resolve to the mirror -- so the reference's
planners and entry scripts (scripts/plan_rm2d.py, DiffusionPlanner) run unmodified on the B200 classes, and its pickled
Configs (which name `diffuser.models.temporal_cond.TemporalUnet_WCond` etc. by dotted path) unpickle to them.
Without the variable the mirror stands alone, as before."""
import importlib
import os


def reference_dir(sub=""):
    root = os.environ.get("PMP_REFERENCE_ROOT")
    if not root:
        return None
    d = os.path.join(root, "diffuser", sub) if sub else os.path.join(root, "diffuser")
    return d if os.path.isdir(d) else None


def extend(path_list, sub, reference_first):
    """adds the reference's directory of sub-package `sub` to a package __path__"""
    d = reference_dir(sub)
    if d and d not in path_list:
        if reference_first:
            path_list.insert(0, d)
        else:
            path_list.append(d)
    return d is not None


def lazy_getattr(package, submodules):
    """module-level __getattr__: the reference's package __init__ star-imports its sub-modules (and with them gym,
    matplotlib, h5py ...); here a missing attribute is looked up in those sub-modules on first use instead"""
    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        if reference_dir() is None:
            raise AttributeError("module %r has no attribute %r (set PMP_REFERENCE_ROOT to overlay the reference checkout)"
                                 % (package, name))
        last = None
        for sub in submodules:
            try:
                mod = importlib.import_module("%s.%s" % (package, sub))
            except ImportError as e:      # a third-party dependency of that sub-module is missing: keep looking
                last = e
                continue
            if hasattr(mod, name):
                return getattr(mod, name)
        raise AttributeError("module %r has no attribute %r%s" % (package, name, " (%s)" % last if last else ""))
    return __getattr__


def reference_fallback(module_name):
    """module-level __getattr__ for a mirror module that shadows a reference module of the same name: a symbol the
    mirror does not define (dynamic-maze helpers, renderers ... outside the hot path) is taken from the reference's own
    file, loaded under a private name, so `from diffuser.guides.rm2d_colchk import check_single_dyn_traj` keeps working"""
    cache = {}

    def __getattr__(name):
        if name.startswith("__"):
            raise AttributeError(name)
        if "mod" not in cache:
            parts = module_name.split(".")           # diffuser.<sub>.<module>
            d = reference_dir(os.path.join(*parts[1:-1])) if len(parts) > 2 else reference_dir()
            path = os.path.join(d, parts[-1] + ".py") if d else None
            if not path or not os.path.isfile(path):
                raise AttributeError("module %r has no attribute %r" % (module_name, name))
            import importlib.util
            import sys
            spec = importlib.util.spec_from_file_location(module_name + "__reference", path)
            mod = importlib.util.module_from_spec(spec)
            mod.__package__ = ".".join(parts[:-1])     # relative imports inside the reference file keep working
            sys.modules[spec.name] = mod
            spec.loader.exec_module(mod)
            cache["mod"] = mod
        try:
            return getattr(cache["mod"], name)
        except AttributeError:
            raise AttributeError("module %r has no attribute %r" % (module_name, name))
    return __getattr__
