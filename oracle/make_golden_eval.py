"""Generates tests/golden/eval_metrics.npz from the reference's OWN convert_dir_vec_to_pose (scripts/utils/data_utils.py:77-98):
the function and its bone table are extracted from the source file with `ast` and executed here (importing the module would need
librosa), then the metric lines of evaluate_testset (scripts/train.py:293-310) are applied verbatim.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_eval"""
import ast
import os

import numpy as np

from .make_golden import OUT

SRC = '/root/reference/scripts/utils/data_utils.py'
MEAN_DIR_VEC = [0.0154009, -0.9690125, -0.0884354, -0.0022264, -0.8655276, 0.4342174, -0.0035145, -0.8755367, -0.4121039, -0.9236511, 0.3061306,
                -0.0012415, -0.5155854, 0.8129665, 0.0871897, 0.2348464, 0.1846561, 0.8091402, 0.9271948, 0.2960011, -0.013189, 0.5233978, 0.8092403,
                0.0725451, -0.2037076, 0.1924306, 0.8196916]          # config/multimodal_context.yml:13 (mean_dir_vec)


def reference_convert():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if (isinstance(n, ast.Assign) and getattr(n.targets[0], 'id', '') == 'dir_vec_pairs')
            or (isinstance(n, ast.FunctionDef) and n.name == 'convert_dir_vec_to_pose')]
    assert len(keep) == 2
    ns = {'np': np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, 'exec'), ns)
    return ns['convert_dir_vec_to_pose']


def main():
    conv = reference_convert()
    rng = np.random.Generator(np.random.PCG64(11))
    B, T, n_pre = 5, 34, 4
    target = (np.cumsum(0.02 * rng.standard_normal((B, T, 27)), axis=1) + 0.1 * rng.standard_normal((B, 1, 27))).astype(np.float32)
    out = (target + 0.05 * rng.standard_normal((B, T, 27))).astype(np.float32)
    # train.py:293-310, verbatim apart from variable names
    out_dir_vec = out.copy().astype(np.float64) + np.array(MEAN_DIR_VEC).squeeze()
    out_joint_poses = conv(out_dir_vec)
    target_vec = target.copy().astype(np.float64) + np.array(MEAN_DIR_VEC).squeeze()
    target_poses = conv(target_vec)
    diff = out_joint_poses[:, n_pre:] - target_poses[:, n_pre:]
    mae_val = np.mean(np.absolute(diff))
    target_acc = np.diff(target_poses, n=2, axis=1)
    out_acc = np.diff(out_joint_poses, n=2, axis=1)
    accel = np.mean(np.abs(target_acc - out_acc))
    np.savez(os.path.join(OUT, 'eval_metrics.npz'), out=out, target=target, mean_dir_vec=np.array(MEAN_DIR_VEC), n_pre=n_pre,
             joint_poses=out_joint_poses, mae=np.float64(mae_val), accel=np.float64(accel),
             l1=np.float64(np.mean(np.abs(out - target), dtype=np.float64)))
    print('mae', mae_val, 'accel', accel)


if __name__ == '__main__':
    main()
