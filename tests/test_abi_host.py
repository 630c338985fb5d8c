"""CPU: the C-ABI library loads and exports every symbol include/pmp_b200.h declares, fails
loudly without a GPU (no CPU fallback), and the drop-in host classes mirror the reference
interface (state-dict keys, buffers, coefficient tables, pickled Config)."""
import ctypes
import io
import os
import pickle
import re

import numpy as np
import pytest
import torch

import pmp_b200
from pmp_b200 import _lib
from oracle import refshim, synth, unet_ref, sampler_ref
from util import ddpm_coefs, ddim_coefs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pmp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pmp_b200.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libpmp_b200.so does not export %s" % s
    assert sorted(_lib.EXPORTS) == syms   # the ctypes table binds exactly the declared ABI
    assert lib.pmp_abi_version() == 1


def test_release_library_has_no_work_skipping_switches():
    """the diagnostic switches that skip work or pick other kernels (PMP_TC_DEBUG, PMP_TC_TIMERS, PMP_KEEP_FP32,
    PMP_NO_UNCOND_DEDUP, PMP_PROFILE_DUMP ...) exist only in `make diag` builds: no such name is in the shipped library"""
    raw = open(_lib.LIB_PATH, "rb").read()
    for name in (b"PMP_TC_DEBUG", b"PMP_TC_TIMERS", b"PMP_TC_TIMER_MATCH", b"PMP_KEEP_FP32", b"PMP_NO_UNCOND_DEDUP",
                 b"PMP_PROFILE_DUMP", b"PMP_TC_AR", b"PMP_TC_TEAM", b"PMP_TC_WIDE"):
        assert name not in raw, name.decode()


def test_conservative_kuka_sphere_table_contains_every_link_box():
    """sphere_table(conservative=True): the union of a link's spheres contains the link's collision-mesh AABB (so a path the
    checker accepts with it is collision free for the mesh too); the default table is the tight, non-conservative model"""
    from diffuser.guides.kuka_colchk import sphere_table, KUKA_LINK_AABB
    rng = np.random.default_rng(0)
    for cons, must_cover in ((True, True), (False, False)):
        tab = sphere_table(conservative=cons)
        assert 1 <= len(tab) <= 32 and set(tab[:, 0].astype(int)) == set(range(len(KUKA_LINK_AABB)))
        covered_all = True
        for l, (mn, mx) in enumerate(KUKA_LINK_AABB):
            mn, mx = np.array(mn), np.array(mx)
            pts = np.concatenate([rng.uniform(mn, mx, (4000, 3)),
                                  np.array([[x, y, z] for x in (mn[0], mx[0]) for y in (mn[1], mx[1]) for z in (mn[2], mx[2])])])
            sp = tab[tab[:, 0] == l]
            d = np.linalg.norm(pts[:, None, :] - sp[None, :, 1:4], axis=2) - sp[None, :, 4]
            covered_all &= bool((d.min(1) <= 1e-6).all())
        assert covered_all == must_cover


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert pmp_b200.device_count() == 0
    a = synth.arch("m2d")
    sd = unet_ref.synth_weights(a["unet"], 1)
    with pytest.raises(pmp_b200.PmpError):
        pmp_b200.UnetEngine(a["unet"], sd, "cpu")
    arch = _lib.make_arch(a["unet"])
    h = ctypes.c_void_p()
    rc = pmp_b200.lib().pmp_model_create(ctypes.byref(arch), 0, ctypes.byref(h))
    assert rc == -2 and b"cuda" in pmp_b200.lib().pmp_last_error().lower()
    from diffuser.models import TemporalUnet_WCond
    m = TemporalUnet_WCond(**a["unet"]).eval()
    with pytest.raises(pmp_b200.PmpError):
        m(torch.zeros(2, 48, 2), None, torch.zeros(2, dtype=torch.long), torch.zeros(2, 12), use_dropout=False)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "potential-motion-plan-release_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


@pytest.mark.parametrize("fam", synth.FAMILIES)
def test_dropin_unet_state_dict(fam):
    from diffuser.models import TemporalUnet_WCond
    a = synth.arch(fam)
    torch.manual_seed(0)
    m = TemporalUnet_WCond(**a["unet"])
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == list(unet_ref.param_shapes(a["unet"]).items())
    m.load_state_dict(unet_ref.synth_weights(a["unet"], 1), strict=True)
    if refshim.available():
        # same seed -> bit-identical random init as the reference constructor
        import subprocess, sys, json
        code = ("import sys,torch;sys.path.insert(0,%r);from oracle import refshim,synth;U,G=refshim.load_models();"
                "a=synth.arch(%r);\nwith refshim.quiet():\n torch.manual_seed(0);m=U(**a['unet'])\n"
                "print(synth.state_dict_digest(m.state_dict()))" % (ROOT, fam))
        ref_digest = subprocess.check_output([sys.executable, "-c", code]).decode().split()[-1]
        torch.manual_seed(0)
        assert synth.state_dict_digest(TemporalUnet_WCond(**a["unet"]).state_dict()) == ref_digest


def test_dropin_diffusion_buffers_and_coefs():
    from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
    a = synth.arch("m2d")
    d = GaussianDiffusionPB(TemporalUnet_WCond(**a["unet"]), **a["diffusion"])
    tb = sampler_ref.schedule(100)
    for k, v in tb.items():
        assert torch.equal(getattr(d, k), v), k
    keys = [k for k in d.state_dict() if not k.startswith("model.")]
    assert keys == list(tb.keys()) + ["loss_fn.weights"]
    assert d.loss_fn.weights.shape == (48, 2) and float(d.loss_fn.weights[0].sum()) == 0 and float(d.loss_fn.weights[-1].sum()) == 0
    ts = list(range(99, -1, -1))
    assert np.array_equal(d.step_coefs_ddpm(ts), ddpm_coefs(tb, ts))
    for n, eta in ((8, 0.0), (24, 1.0), (10, 0.0)):
        t2 = d.ddim_set_timesteps(n)
        assert np.array_equal(t2, sampler_ref.ddim_timesteps(n, 100))
        assert np.array_equal(d.step_coefs_ddim(t2, eta, n), ddim_coefs(tb, t2, n, 100, eta))
    # the training step is CUDA only: on a CPU model `loss` fails loudly instead of falling back to torch autograd
    import pmp_b200
    d.train()
    x = torch.zeros(2, 48, 2)
    with pytest.raises(pmp_b200.PmpError):
        d.loss(x, {0: x[:, 0], 47: x[:, -1]}, torch.zeros(2, 1, 12))


def test_config_pickle_roundtrip():
    from diffuser.utils import Config
    a = synth.arch("kuka")
    cfg = Config("models.TemporalUnet_WCond", **a["unet"])
    blob = pickle.dumps(cfg)
    cfg2 = pickle.loads(blob)
    m = cfg2()
    assert type(m).__module__ == "diffuser.models.temporal_cond" and type(m).__name__ == "TemporalUnet_WCond"
    dcfg = pickle.loads(pickle.dumps(Config("models.GaussianDiffusionPB", **a["diffusion"])))
    d = dcfg(m)
    from diffuser.models import GaussianDiffusionPB
    assert type(d) == GaussianDiffusionPB and d.n_timesteps == 100 and d.var_temp == 1.0


def test_unsupported_configs_fail_loudly():
    from diffuser.models import TemporalUnet_WCond
    a = synth.arch("m2d")
    bad = dict(a["unet"]); bad["network_config"] = dict(bad["network_config"], energy_mode=False)
    with pytest.raises(Exception):
        TemporalUnet_WCond(**bad)
    bad = dict(a["unet"]); bad["network_config"] = dict(bad["network_config"], wallLoc_encoder="mlp")
    with pytest.raises(Exception):
        TemporalUnet_WCond(**bad)


def test_normalizer_matches_oracle():
    from diffuser.datasets import LimitsNormalizer, KUKA7D_MIN, KUKA7D_MAX
    n = LimitsNormalizer(KUKA7D_MIN, KUKA7D_MAX)
    x = (np.random.default_rng(0).normal(0, 0.8, (5, 52, 7))).astype(np.float32)
    assert np.array_equal(n.unnormalize(x), synth.unnormalize(x, KUKA7D_MIN, KUKA7D_MAX))
    q = np.random.default_rng(1).uniform(KUKA7D_MIN, KUKA7D_MAX, (4, 7)).astype(np.float32)
    assert np.array_equal(n.normalize(q), synth.normalize(q, KUKA7D_MIN, KUKA7D_MAX))


def test_workloads_match_oracle():
    """bench.py draws its synthetic inputs from the product-side copy of the generators"""
    from pmp_b200 import workloads
    for fam in synth.FAMILIES:
        assert workloads.arch(fam) == synth.arch(fam)
        a, b = workloads.make_problems(fam, 7, 5), synth.make_problems(fam, 7, 5)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (fam, k)
    a = synth.arch("m2d")
    sd1 = unet_ref.synth_weights(a["unet"], 3)
    sd2 = workloads.synth_state_dict({k: torch.empty(s) for k, s in unet_ref.param_shapes(a["unet"]).items()}, 3)
    assert synth.state_dict_digest(sd1) == workloads.state_dict_digest(sd2)


def test_problem_file_ingest_npz_roundtrip(tmp_path):
    """pmp_b200.ingest: the reference's problem-file dict (problems, infos/wall_locations) -> planner arrays
    (kuka_plan_utils.py:36-57, kuka_plan_env.py:285-313); HDF5 is refused loudly where h5py is missing"""
    from pmp_b200 import ingest
    rng = np.random.default_rng(0)
    n_env, n_prob, nw = 5, 7, 3
    cen = rng.uniform(0.7, 4.3, (n_env, 1, nw, 2))
    d = {"problems": rng.uniform(0, 5, (n_env, n_prob, 2, 2)).astype(np.float32),
         "infos/wall_locations": np.repeat(cen, n_prob, 1).reshape(n_env, n_prob, nw * 2).astype(np.float32),
         "maze_idx": np.arange(n_env)}
    p = str(tmp_path / "m2d-problems.npz")
    ingest.save_problems_npz(d, p)
    got = ingest.load_problems(p)
    assert set(got) == set(d) and all(np.array_equal(got[k], d[k]) for k in d)
    s, g, w, c = ingest.problem_arrays(got, [3, 1], prob_permaze=4, wall_dim=2, prob_start_idx=2)
    assert s.shape == (2, 4, 2) and g.shape == (2, 4, 2) and w.shape == (2, 4, 6) and c.shape == (2, 3, 2)
    assert np.array_equal(s[0], d["problems"][3, 2:6, 0]) and np.array_equal(g[1], d["problems"][1, 2:6, 1])
    assert np.array_equal(w[0, 0], d["infos/wall_locations"][3, 2]) and np.allclose(c[0], cen[3, 0])
    # training-set layout [n_env, n_prob, nw, >= wall_dim]: only the centre coordinates are kept
    d4 = dict(d); d4["infos/wall_locations"] = np.concatenate([np.repeat(cen, n_prob, 1), np.ones((n_env, n_prob, nw, 2))], -1)
    _, _, w4, c4 = ingest.problem_arrays(d4, [0], 2, 2)
    assert w4.shape == (1, 2, 6) and np.allclose(c4[0], cen[0, 0])
    with pytest.raises(ValueError):
        ingest.problem_arrays(got, [0], prob_permaze=99, wall_dim=2)
    with pytest.raises((OSError, ValueError)):
        ingest.load_problems(str(tmp_path / "x-problems.hdf5"))       # missing file: an error, never an empty dict


def test_hdf5_problem_file_is_read_like_its_npz_twin(tmp_path):
    """tests/golden/m2d_synth-problems.hdf5 (HDF5 in the layout h5py's defaults write: superblock 0, old-style nested
    groups, gzip-chunked / contiguous / compact datasets; make_h5_fixture.py) through ingest.load_problems -- h5py when
    installed, else the built-in reader pmp_b200/h5min.py -- gives the arrays of its .npz twin, bit for bit, and the
    planner arrays cut from both are identical.  Unsupported / damaged files fail loudly."""
    from pmp_b200 import ingest, h5min
    gdir = os.path.join(ROOT, "tests", "golden")
    h5 = ingest.load_problems(os.path.join(gdir, "m2d_synth-problems.hdf5"))
    nz = ingest.load_problems(os.path.join(gdir, "m2d_synth-problems.npz"))
    assert set(nz) <= set(h5) and {"maze_idx", "infos/is_sol_kp"} <= set(h5)
    for k in nz:
        assert h5[k].dtype == nz[k].dtype and np.array_equal(h5[k], nz[k]), k
    assert h5["maze_idx"].dtype == np.int64 and np.array_equal(h5["maze_idx"][:, 0], np.arange(h5["problems"].shape[0]))
    a = ingest.problem_arrays(h5, np.arange(4), 7, 2)
    b = ingest.problem_arrays(nz, np.arange(4), 7, 2)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # the reader itself (also when h5py is present)
    flat = h5min.read_flat(os.path.join(gdir, "m2d_synth-problems.hdf5"))
    assert all(np.array_equal(flat[k], nz[k]) for k in nz)
    raw = bytearray(open(os.path.join(gdir, "m2d_synth-problems.hdf5"), "rb").read())
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(b"not an hdf5 file" + bytes(raw[16:]))
    with pytest.raises(h5min.H5FormatError):
        h5min.read_flat(str(bad))
    raw[8] = 3                                   # superblock version 3 (libver='latest'): outside the built subset
    bad.write_bytes(bytes(raw))
    with pytest.raises(h5min.H5FormatError):
        h5min.read_flat(str(bad))


def test_concave_layout_ingest_matches_reference_fixture(tmp_path):
    """pmp_b200.ingest concave helpers against tests/golden/concave_layouts.npz (layouts drawn and converted by the
    reference's own RandRectangleGroup_43, gen_golden.py concave) and against the loop restatement; bit-exact"""
    from pmp_b200 import ingest
    from oracle import ingest_ref
    from util import golden
    z = golden("concave_layouts.npz")
    layouts, tokens, hext = z["layouts"], z["tokens"], float(z["hext"])
    assert np.array_equal(ingest_ref.centers_and_modes(layouts), tokens)
    got = ingest.concave_tokens_from_boxes(layouts)
    assert got.dtype == np.float64 and np.array_equal(got, tokens)
    assert np.array_equal(got[..., 2], z["sampled_modes"] + 1)            # mode = deleted quadrant + 1
    # (centre, mode) -> squares: the generator's order, equal to the saved layout up to the rounding of its centres
    boxes = ingest.concave_boxes(tokens[..., :2], tokens[..., 2], hext)
    assert boxes.shape == layouts.shape and np.abs(boxes - layouts).max() < 1e-12
    assert np.array_equal(ingest.concave_tokens_from_boxes(boxes)[..., 2], tokens[..., 2])
    for e in range(3):
        for o in range(7):
            sq = ingest_ref.squares_of(tokens[e, o, :2], hext, int(tokens[e, o, 2]) - 1)
            assert np.array_equal(sq, boxes[e, 3 * o:3 * o + 3])
    # the two files the generator leaves behind: <name>_eval.npy and <name>_43_eval.pkl
    npy = str(tmp_path / "grp_eval.npy"); np.save(npy, layouts)
    assert np.array_equal(ingest.load_wall_layouts(npy), layouts)
    pkl = str(tmp_path / "grp_43_eval.pkl")
    with open(pkl, "wb") as f:
        pickle.dump({"centers_43": tokens[..., :2].copy(), "mode_list": tokens[..., 2].astype(np.int64)}, f)
    cen, modes = ingest.load_concave_43(pkl)
    assert np.array_equal(cen, tokens[..., :2]) and np.array_equal(modes, tokens[..., 2].astype(np.int64))
    wl = ingest.concave_wall_locations(cen, modes, n_probs=5)
    assert wl.shape == (24, 5, 7, 3) and np.array_equal(wl, ingest_ref.val_wall_locations(cen, modes.astype(np.float64), 5))
    d = {"problems": np.zeros((24, 5, 2, 2), np.float32), "infos/wall_locations": wl}
    _, _, w, c = ingest.problem_arrays(d, [2, 7], prob_permaze=3, wall_dim=3)
    assert w.shape == (2, 3, 21) and np.array_equal(w[1, 0], tokens[7].astype(np.float32).ravel())
    assert c.shape == (2, 7, 3) and np.array_equal(c[0], tokens[2])
    # malformed input fails loudly
    bad = layouts.copy(); bad[0, 0, 1] += 0.01; bad[0, 0, 0] += 0.01
    with pytest.raises(ValueError):
        ingest.concave_tokens_from_boxes(bad)
    with pytest.raises(ValueError):
        ingest.concave_boxes(tokens[..., :2], np.zeros((24, 7)), hext)
    evil = str(tmp_path / "evil.pkl")
    with open(evil, "wb") as f:
        pickle.dump({"centers_43": os.path.join}, f)
    with pytest.raises(pickle.UnpicklingError):
        ingest.load_concave_43(evil)


def test_concave_problem_arrays_sources_agree():
    """the three places a concave layout can come from (problem file squares, the .npy, the _43 pickle) give the same
    model input and checker squares (plan_pb_diff_helper.py:128-134)"""
    from pmp_b200 import ingest
    from util import golden
    z = golden("concave_layouts.npz")
    layouts, tokens = z["layouts"], z["tokens"]
    n_env, n_p = len(layouts), 4
    rng = np.random.default_rng(1)
    probs = rng.uniform(0, 5, (n_env, n_p, 2, 2)).astype(np.float32)
    raw = {"problems": probs, "infos/wall_locations": np.repeat(layouts[:, None], n_p, 1)}
    tok = {"problems": probs, "infos/wall_locations": np.repeat(tokens[:, None], n_p, 1)}
    flat = {"problems": probs, "infos/wall_locations": np.zeros((n_env, n_p, 21), np.float32)}
    idx = [5, 0, 11]
    ref = ingest.concave_problem_arrays(raw, idx, 3, prob_start_idx=1)
    assert ref[0].shape == (3, 3, 2) and ref[2].shape == (3, 3, 21) and ref[3].shape == (3, 21, 2)
    assert np.array_equal(ref[0], probs[idx][:, 1:4, 0]) and np.array_equal(ref[1], probs[idx][:, 1:4, 1])
    assert np.array_equal(ref[2][0, 2], tokens[5].astype(np.float32).ravel()) and np.array_equal(ref[3], layouts[idx])
    cm = (tokens[..., :2], tokens[..., 2].astype(np.int64))
    for d, kw in ((flat, dict(layouts=layouts)), (flat, dict(centers_modes=cm, layouts=layouts)), (raw, dict(centers_modes=cm))):
        got = ingest.concave_problem_arrays(d, idx, 3, prob_start_idx=1, **kw)
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    for d, kw in ((tok, {}), (flat, dict(centers_modes=cm))):          # no squares on file: centre +- hext
        got = ingest.concave_problem_arrays(d, idx, 3, prob_start_idx=1, hext=0.25, **kw)
        assert all(np.array_equal(a, b) for a, b in zip(got[:3], ref[:3])) and np.abs(got[3] - ref[3]).max() < 1e-12
    with pytest.raises(ValueError):
        ingest.concave_problem_arrays(flat, idx, 3)


def _import_plan_env_host():
    """diffuser.guides.plan_env_b200.compute_avg_result is host arithmetic; importing the module needs no GPU"""
    from diffuser.guides.plan_env_b200 import compute_avg_result
    return compute_avg_result


def test_result_statistics_match_reference_golden():
    """compute_avg_result against what the reference's plan_kuka_permaze_stat + compute_kuka_avg_result
    (kuka_plan_utils.py:110-213) returned for the chosen paths of tests/golden/pick_multi_horizon.npz"""
    from util import golden
    f = _import_plan_env_host()
    z = golden("pick_multi_horizon.npz")
    E, n_p = z["collision_cnts"].shape
    res = f(z["collision_cnts"].ravel(), z["success"].ravel(), z["num_colck"].ravel(), E, n_p, float(z["batch_runtime_total"]))
    keys = [k[5:] for k in z.files if k.startswith("stat_")]
    assert set(keys) == set(res) and len(keys) == 10
    for k in keys:
        assert float(res[k]) == float(z["stat_" + k]), (k, res[k], float(z["stat_" + k]))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_result_statistics_match_live_reference():
    """same, on random per-environment outcomes with sizes whose rates are not exact binary fractions"""
    f = _import_plan_env_host()
    PU = refshim._import_ref("diffuser.guides.kuka_plan_utils")[0]
    rng = np.random.default_rng(3)
    for n_env, n_prob in ((1, 1), (7, 3), (13, 21), (100, 20)):
        ncol = rng.integers(0, 3, (n_env, n_prob)) * (rng.random((n_env, n_prob)) < 0.6)
        nor = rng.random((n_env, n_prob)) < 0.7
        rep = rng.random((n_env, n_prob)) < 0.2
        colck = rng.integers(1, 5000, (n_env, n_prob))
        t = rng.uniform(0.1, 2.0, n_env)
        per_env = []
        for e in range(n_env):
            ok = ncol[e] == 0
            per_env.append(dict(num_samples=n_prob, traj_srate=ok.sum() / n_prob, no_collision_rate=ok.sum() / n_prob,
                                se_valid_rate=np.array([True] * n_prob).sum() / n_prob, batch_runtime=t[e],
                                no_replan_suc_list=list(nor[e]), replan_suc_list=list(rep[e]), num_colck_list=list(colck[e])))
        with refshim.quiet():
            ref = PU.compute_kuka_avg_result(per_env)
        got = f(ncol.ravel(), nor.ravel(), colck.ravel(), n_env, n_prob, sum(t.tolist()), repl_suc=rep.ravel())
        for k in ref:
            if k in ("avg_batch_time", "avg_sample_time"):      # the reference adds the per-maze times in its own order
                assert abs(got[k] - ref[k]) <= 1e-12 * abs(ref[k])
            else:
                assert float(got[k]) == float(ref[k]), (k, got[k], ref[k])


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_problem_batch_layout_matches_live_reference():
    """problem file -> the [n_prob * n_samples] batch the sampler sees: ingest.problem_arrays + the problem-major
    repeat of plan_envs against the reference's rm2d_kuka_preproc (preprocessing.py:36-67), kuka_obs_target_list
    (kuka_plan_utils.py:36-57), fit_obs_target_tensor and get_wallLoc_tensor_single (kuka_plan_env.py:264-316)"""
    from pmp_b200 import ingest
    PU, PP, KP = refshim._import_ref("diffuser.guides.kuka_plan_utils", "diffuser.datasets.preprocessing",
                                     "diffuser.guides.kuka_plan_env")
    rng = np.random.default_rng(6)
    for D, nw, wd_file, wd in ((2, 6, 2, 2), (7, 4, 6, 3), (14, 5, 6, 3)):
        n_env, n_prob_file, n_prob, n_s = 5, 9, 4, 3
        raw = {"problems": rng.uniform(-3, 3, (n_env, n_prob_file, 2, D)).astype(np.float32),
               "infos/wall_locations": rng.uniform(-1, 1, (n_env, n_prob_file, nw, wd_file)).astype(np.float32)}
        ref_d = {k: v.copy() for k, v in raw.items()}

        class E_:
            name = "kuka" if D > 2 else "maze2d"
            is_eval = True
        with refshim.quiet():
            PP.rm2d_kuka_preproc(ref_d, wloc_select=tuple(range(wd)), env=E_)       # in place: [n_env, n_p, nw*wd]
        for maze_idx in (0, 3):
            pl = object.__new__(KP.KukaEnvPlanner)
            pl.problems_dict = ref_d; pl.maze_idx = maze_idx
            with refshim.quiet():
                o, t = PU.kuka_obs_target_list(ref_d, maze_idx, n_prob)
                o_t, t_t = pl.fit_obs_target_tensor(o, t, n_s)
                w_t = pl.get_wallLoc_tensor_single(0, n_prob, n_s)
            s_, g_, w_, _ = ingest.problem_arrays(raw, [maze_idx], n_prob, wd)
            rep = lambda a: np.repeat(a.reshape(n_prob, -1), n_s, axis=0)              # plan_envs' batch layout
            assert np.array_equal(rep(s_), o_t.numpy()) and np.array_equal(rep(g_), t_t.numpy())
            assert np.array_equal(rep(w_), w_t.numpy())


def test_replan_failed_rounds_equal_sequential_reference_order():
    """plan_env_b200.replan_failed (round r repairs candidate r of all still-failed problems at once) against the
    sequential restatement of KukaEnvPlanner.replan_pred_trajs (kuka_plan_env.py:207-260); the replanner is a stand-in
    whose replan_batch runs the pinned oracle per row, so only the orchestration is under test (no GPU)"""
    from diffuser.guides.plan_env_b200 import replan_failed
    from oracle import replan_ref
    E_, n_prob, n_s, hext, n_trial = 4, 5, 7, 0.5, 2
    pr = synth.make_problems("m2d", E_, 51)
    rng = np.random.default_rng(14)
    n_p = E_ * n_prob
    env_p = np.repeat(np.arange(E_), n_prob)
    horizons = (40, 48)

    def resample(trial, traj):          # deterministic function of the trajectory only
        al = np.linspace(0, 1, len(traj))[:, None]
        return np.clip(traj + (0.45 if trial else -0.3) * np.sin(np.pi * al) ** 2, 0.02, 4.98).astype(np.float32)
    cands = {}
    suc = rng.random(n_p) < 0.3
    for p in np.nonzero(~suc)[0]:
        s, g = rng.uniform(0.3, 4.7, 2), rng.uniform(0.3, 4.7, 2)
        cands[int(p)] = []
        for j in range(n_s):
            H = horizons[j % 2]
            al = np.linspace(0, 1, H)[:, None]
            cands[int(p)].append(np.clip(s * (1 - al) + g * al + rng.normal(0, 0.08, (H, 2)) * np.sin(np.pi * al), 0.02, 4.98).astype(np.float32))

    class Stub:
        calls = 0
        def replan_batch(self, trajs, walls, geom):
            Stub.calls += 1
            assert len({len(t) for t in trajs}) == 1 and len(walls) == len(trajs) == len(geom["env_id"])
            out = [replan_ref.replan_collision_section(t, replan_ref.Maze2DChecker(geom["cen"][e], geom["hext"]), resample, n_trial)
                   for t, e in zip(trajs, geom["env_id"])]
            return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]
    walls_p = np.zeros((n_p, 12), np.float32)
    got = replan_failed(Stub(), cands, walls_p, env_p, dict(robot="maze2d", cen=pr["boxes"], hext=hext), suc)
    ref = replan_ref.replan_pred_trajs(cands, suc, lambda p: replan_ref.Maze2DChecker(pr["boxes"][env_p[p]], hext), resample, n_trial)
    assert set(got[0]) == set(ref[0]) and all(np.array_equal(got[0][p], ref[0][p]) for p in ref[0])
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    n_failed = int((~suc).sum())
    assert 0 < len(ref[0]) < n_failed and (got[1] >= suc).all()
    assert Stub.calls <= 6 * len(horizons)                     # rounds x lengths, not problems x candidates


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_normalizer_matches_live_reference():
    """LimitsNormalizer (constant limits) against the reference class (datasets/normalization.py:137-162) built with the
    same limits through its min_max argument; float32 / float64 inputs, in and out of range"""
    from diffuser.datasets import LimitsNormalizer, KUKA7D_MIN, KUKA7D_MAX, RAND_MAZE_MIN, RAND_MAZESIZE_55
    Ref = refshim.load_replanner()[2]
    rng = np.random.default_rng(2)
    for lo, hi in ((KUKA7D_MIN, KUKA7D_MAX), (RAND_MAZE_MIN, RAND_MAZESIZE_55)):
        D = len(lo)
        ref = Ref(np.stack([lo, hi]).astype(np.float32), (np.asarray(lo, np.float32), np.asarray(hi, np.float32)))
        got = LimitsNormalizer(lo, hi)
        for dt in (np.float32, np.float64):
            q = rng.uniform(lo, hi, (6, 48, D)).astype(dt)
            assert np.array_equal(got.normalize(q), ref.normalize(q)) and got.normalize(q).dtype == ref.normalize(q).dtype
            for x in (rng.uniform(-1, 1, (6, 48, D)).astype(dt), rng.normal(0, 0.9, (6, 48, D)).astype(dt)):
                a, b = got.unnormalize(x), ref.unnormalize(x)
                assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_compose_wall_split_matches_live_reference():
    """PolicyCompose.wallLoc_preproc against the reference method (policies_compose.py:95-181) for the three
    composition types, including the jitter-padded one- and two-obstacle second model (same torch seed)"""
    from diffuser.guides import PolicyCompose
    RefPC = refshim._import_ref("diffuser.guides.policies_compose")[0].PolicyCompose
    cases = [("direct", [6, 3], 2), ("direct", [6, 1], 2), ("direct", [6, 2], 2), ("direct", [4, 3, 2], 3),
             ("direct", [3, 1, 3], 2), ("share", [6, 2], 2), ("share-model2fix", [6, 3], 2), ("share", [4, 1], 3)]
    for comp_type, nwc, wd in cases:
        objs = []
        for cls in (RefPC, PolicyCompose):
            o = object.__new__(cls)
            o.comp_type = comp_type; o.num_walls_c = torch.tensor(nwc); o.wall_dim = wd; o.n_models = len(nwc)
            o.is_dyn_env = False; o.wall_is_dyn = None
            o.w_split_sizes = (tuple((o.num_walls_c * wd).to(torch.int32)) if cls is RefPC
                               else tuple(int(v) for v in (o.num_walls_c * wd)))
            objs.append(o)
        n_total = sum(nwc) if comp_type == "direct" else (nwc[0] + nwc[1])
        w = torch.randn(5, n_total * wd, generator=torch.Generator().manual_seed(1))
        torch.manual_seed(4)
        with refshim.quiet():
            ref = objs[0].wallLoc_preproc(w.clone())
        torch.manual_seed(4)
        got = objs[1].wallLoc_preproc(w.clone())
        assert len(ref) == len(got), (comp_type, nwc)
        for a, b in zip(ref, got):
            assert torch.equal(a, b), (comp_type, nwc)


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_replanner_helpers_match_live_reference():
    """section helpers of the batched replanner (product and oracle copies) against the reference functions
    (kuka_replan.py:176-253), exhaustively over the section positions of the planning horizons"""
    from diffuser.guides import kuka_replan as prod
    from oracle import replan_ref
    R = refshim._import_ref("diffuser.guides.kuka_replan")[0]
    for H in (40, 48, 56):
        for st in range(H):
            for end in range(st, min(H, st + 30)):
                if end - st + 1 >= H - 1:
                    continue
                ref = R.KukaDiffusionReplanner.get_expanded_subtraj(None, H, st, end)
                assert np.array_equal(prod.expanded_subtraj(H, st, end), ref)
                assert np.array_equal(replan_ref.expanded_subtraj(H, st, end), ref)
    rng = np.random.default_rng(0)
    for _ in range(200):
        arr = np.nonzero(rng.random(48) < rng.uniform(0.05, 0.9))[0]
        if not len(arr):
            continue
        ref = R.split_sorted_array_into_subarrays(arr)
        for f in (prod.split_sorted_array_into_subarrays, replan_ref.split_runs):
            got = f(arr)
            assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(got, ref))
    for D in (2, 7, 14):
        assert prod.get_se_thres(np.zeros((4, D))) == R.get_se_thres(np.zeros((4, D))) == replan_ref.se_thres(np.zeros((4, D)))


@pytest.mark.skipif(not refshim.available(), reason="reference tree not present")
def test_solution_interp_matches_live_reference():
    """sol_interp_pair (product) / sol_interp (oracle) against the reference's SolutionInterp
    (pb_diff_envs/utils/kuka_utils_luo.py:10-55) for two-waypoint inputs (what steerTo_repl feeds it) and, for the
    oracle, longer waypoint lists; float32 and float64 ends, including nearly coincident ones"""
    from diffuser.guides.kuka_replan import sol_interp_pair
    from oracle import replan_ref
    SI = refshim._import_ref("pb_diff_envs.utils.kuka_utils_luo")[0].SolutionInterp
    rng = np.random.default_rng(5)
    for density in (3, 8):
        ref_f = SI(density=density)
        for D in (2, 7, 14):
            for dt in (np.float32, np.float64):
                for scale in (1.0, 0.05, 1e-4):
                    a = rng.uniform(-2, 2, D).astype(dt); b = (a + rng.normal(0, scale, D)).astype(dt)
                    with refshim.quiet():
                        ref = np.asarray(ref_f([a, b]))
                    got = sol_interp_pair(a, b, density)
                    assert got.shape == ref.shape and np.array_equal(got, ref), (density, D, dt, scale)
                    assert np.array_equal(replan_ref.sol_interp([a, b], density), ref)
                pts = [rng.uniform(-2, 2, D).astype(dt) for _ in range(5)]
                with refshim.quiet():
                    ref = np.asarray(ref_f(pts))
                assert np.array_equal(replan_ref.sol_interp(pts, density), ref)


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) on a two-plan sample: one JSON line with the contract's
    keys, the CPU baseline described (kind / cores / sample) and the checker leg beside it; no GPU involved"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-plans", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "plans/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["gpu_launches"] == 0 and d["config"]["plans_per_step"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "plans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "DDPM-100" in cb["sample"] and cb["value"] == d["value"]
    assert cb["checker"]["unit"] == "trajectories/s" and cb["checker"]["value"] > 0 and cb["checker"]["cores"] == 1
