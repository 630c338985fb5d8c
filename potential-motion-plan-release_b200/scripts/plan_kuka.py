"""plan_kuka entry point (reference: scripts/plan_kuka.py + plan_kuka.sh): Kuka 7-DoF / Dual-Kuka
14-DoF planning on synthetic voxel environments with the analytic sphere-link collision model
(declared deviation from the reference's pybullet meshes).  Evaluation hyper-parameters follow
scripts/plan_kuka.py:51-77 (20 problems x 20 samples, 8 / 10 DDIM steps, horizon 52 / 56).

    python scripts/plan_kuka.py --family kuka --plan_n_maze 20
    python scripts/plan_kuka.py --family dualkuka --ddim_steps 10
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from plan_rm2d import main  # noqa: E402

if __name__ == '__main__':
    main(robot='kuka')
