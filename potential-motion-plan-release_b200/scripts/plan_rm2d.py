"""plan_rm2d entry point (reference: scripts/plan_rm2d.py + plan_rm2d.sh) on the B200 hot path.

The reference script needs its gym environments, HDF5 problem files and trained checkpoints, none
of which exist here, so this entry point plans SYNTHETIC static-Maze2D problem sets with the
reference's evaluation hyper-parameters (scripts/plan_rm2d.py:57-70: 20 problems x 10 samples per
environment, 8 DDIM steps, cond_w 2.0, horizon 48).  Pass --logdir <dir with model_config.pkl,
diffusion_config.pkl, state_<epoch>.pt> to plan with a reference checkpoint instead of random
weights.

    python scripts/plan_rm2d.py --family m2d --plan_n_maze 50
"""
import argparse
import glob
import json
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from pmp_b200 import workloads
from diffuser.models import TemporalUnet_WCond, GaussianDiffusionPB
from diffuser.guides import Policy
from diffuser.guides.plan_env_b200 import plan_envs
from diffuser.datasets import DatasetNormalizer


def latest_checkpoint(logdir):
    """state_<epoch>.pt with the highest epoch number, like get_latest_epoch (utils/serialization.py:27-33)"""
    best, path = -1, None
    for f in glob.glob(os.path.join(logdir, 'state_*.pt')):
        try:
            ep = int(os.path.basename(f)[len('state_'):-len('.pt')])
        except ValueError:
            continue
        if ep > best:
            best, path = ep, f
    if path is None:
        raise FileNotFoundError('no state_<epoch>.pt under %s' % logdir)
    return path


def load_diffusion(family, logdir, device):
    a = workloads.arch(family)
    if logdir:
        # the two config pickles are trusted input (pickled class objects, exactly like the reference's
        # utils/serialization.py:98-129); the checkpoint holds tensors only and is loaded as such
        model = pickle.load(open(os.path.join(logdir, 'model_config.pkl'), 'rb'))()
        diffusion = pickle.load(open(os.path.join(logdir, 'diffusion_config.pkl'), 'rb'))(model)
        diffusion.load_state_dict(torch.load(latest_checkpoint(logdir), map_location='cpu', weights_only=True)['ema'])
    else:
        torch.manual_seed(0)
        model = TemporalUnet_WCond(**a['unet'])
        diffusion = GaussianDiffusionPB(model, **a['diffusion'])
    return a, diffusion.to(device).eval()


def make_replanner(args, diffusion, norm, robot):
    """KukaDiffusionReplanner + the planners' collision checker settings (comp_plan_env.py:166-177)"""
    if not args.do_replan:
        return None
    from diffuser.guides import RM2DCollisionChecker, KukaCollisionChecker
    from diffuser.guides.kuka_replan import KukaDiffusionReplanner
    rp = KukaDiffusionReplanner(diffusion, norm, dict(use_ddim=True, dn_steps=args.repl_dn_steps, n_replan_trial=5),
                                args.device)
    if robot == 'maze2d':
        chk = RM2DCollisionChecker(normalizer=None, device=args.device)
        chk.use_collision_eps = True
        chk.collision_eps = 0.08
    else:
        chk = KukaCollisionChecker(device=args.device, conservative=bool(args.kuka_conservative))
    rp.set_collision_checker(chk)
    return rp


def main(argv=None, robot='maze2d'):
    ap = argparse.ArgumentParser()
    ap.add_argument('--family', default='m2d' if robot == 'maze2d' else 'kuka', choices=workloads.FAMILIES)
    ap.add_argument('--plan_n_maze', type=int, default=10)
    ap.add_argument('--n_prob_env', type=int, default=20)
    ap.add_argument('--samples_perprob', type=int, default=10 if robot == 'maze2d' else 20)
    ap.add_argument('--ddim_steps', type=int, default=8)
    ap.add_argument('--use_ddim', type=int, default=1)
    ap.add_argument('--cond_w', type=float, default=2.0)
    ap.add_argument('--logdir', default=None)
    ap.add_argument('--problems', default=None,
                    help='problem file of the reference (<dataset>-problems.hdf5, needs h5py) or its .npz conversion '
                         '(pmp_b200.ingest); default: synthetic problems')
    ap.add_argument('--hext', type=float, default=None, help='half extent of the obstacles in --problems')
    ap.add_argument('--layouts', default=None, help='concave mazes: rand_rec2dgrp/<env_list_name>_eval.npy')
    ap.add_argument('--concave43', default=None, help='concave mazes: rand_rec2dgrp/<env_list_name>_43_eval.pkl')
    ap.add_argument('--do_replan', type=int, default=0,
                    help='repair problems whose candidates all collide with the batched replanner (reference: do_replan)')
    ap.add_argument('--repl_dn_steps', type=int, default=3)
    ap.add_argument('--kuka_conservative', type=int, default=0,
                    help='Kuka / Dual Kuka: score with spheres that contain the link boxes (accepted paths are collision free '
                         'for the meshes as well; lower success rates) instead of the tight sphere model')
    ap.add_argument('--device', default='cuda:0')
    args = ap.parse_args(argv)
    a, diffusion = load_diffusion(args.family, args.logdir, args.device)
    diffusion.condition_guidance_w = args.cond_w
    diffusion.ddim_num_inference_steps = args.ddim_steps
    H = a['horizon']
    n_env, n_prob = args.plan_n_maze, args.n_prob_env
    starts, goals, walls, boxes = [], [], [], []
    if args.problems:
        from pmp_b200 import ingest
        pd = ingest.load_problems(args.problems)
        wd = 2 if robot == 'maze2d' else 3
        if args.family == 'm2d_concave':
            # (cx, cy, mode) rows from the generator's pickle / layout, squares for the checker (ingest.py)
            s_, g_, w_, c_ = ingest.concave_problem_arrays(
                pd, np.arange(n_env), n_prob,
                centers_modes=ingest.load_concave_43(args.concave43) if args.concave43 else None,
                layouts=ingest.load_wall_layouts(args.layouts) if args.layouts else None,
                hext=args.hext if args.hext is not None else 0.25)
        else:
            s_, g_, w_, c_ = ingest.problem_arrays(pd, np.arange(n_env), n_prob, wd)
        lo, hi = workloads.limits(args.family)
        norm = DatasetNormalizer({'observations': (lo, hi)})
        policy = Policy(diffusion, norm, use_ddim=bool(args.use_ddim))
        hext = args.hext if args.hext is not None else workloads.make_problems(args.family, 1, 0)['hext']
        res, _ = plan_envs(policy, s_, g_, w_, c_, hext, H, samples_perprob=args.samples_perprob, robot=robot,
                           n_arms=2 if args.family == 'dualkuka' else 1,
                           replanner=make_replanner(args, diffusion, norm, robot), kuka_conservative=bool(args.kuka_conservative))
        res['config'] = vars(args)
        print(json.dumps(res))
        return res
    for e in range(n_env):
        # one obstacle layout per environment, n_prob start/goal pairs in it
        pr = workloads.make_problems(args.family, 1, 5000 + e)
        lay = dict(boxes=pr['boxes'][0], walls=pr['walls'][0], hext=pr['hext'])
        rng = np.random.default_rng(7000 + e)
        s, g = [], []
        for _ in range(n_prob):
            if robot == 'maze2d':
                si, gi = workloads.maze2d_start_goal(rng, lay['boxes'], lay['hext'])
            else:
                si = rng.uniform(pr['lo'], pr['hi']).astype(np.float32)
                gi = rng.uniform(pr['lo'], pr['hi']).astype(np.float32)
            s.append(si); g.append(gi)
        starts.append(s); goals.append(g)
        walls.append(np.repeat(lay['walls'][None], n_prob, 0)); boxes.append(lay['boxes'])
    lo, hi = workloads.limits(args.family)
    norm = DatasetNormalizer({'observations': (lo, hi)})
    policy = Policy(diffusion, norm, use_ddim=bool(args.use_ddim))
    res, _ = plan_envs(policy, np.array(starts, np.float32), np.array(goals, np.float32), np.array(walls, np.float32),
                       np.array(boxes), pr['hext'], H, samples_perprob=args.samples_perprob, robot=robot,
                       n_arms=2 if args.family == 'dualkuka' else 1,
                       replanner=make_replanner(args, diffusion, norm, robot), kuka_conservative=bool(args.kuka_conservative))
    res['config'] = vars(args)
    print(json.dumps(res))
    return res


if __name__ == '__main__':
    main()
