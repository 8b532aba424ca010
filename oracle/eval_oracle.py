"""CPU oracle for the validation metrics of evaluate_testset (scripts/train.py:234-329).  TEST INFRASTRUCTURE ONLY.

numpy float64 restatement of
  * convert_dir_vec_to_pose           scripts/utils/data_utils.py:14-15,77-98  (9 direction vectors -> 10 joint positions along the
                                       fixed-bone-length kinematic tree, root at the origin)
  * the per-batch metrics             scripts/train.py:283,293-310: L1 loss of the direction vectors, MAE of the joint coordinates over
                                       the generated frames (>= n_pre_poses), mean |second time difference| mismatch ("accel").
Parity pinning: oracle/make_golden_eval.py extracts `dir_vec_pairs` and `convert_dir_vec_to_pose` from the reference source file (the
module itself cannot be imported here: it pulls in librosa) and executes THEM on seeded inputs -> tests/golden/eval_metrics.npz."""
import numpy as np

# (parent joint, child joint, bone length)  data_utils.py:14-15
DIR_VEC_PAIRS = [(0, 1, 0.26), (1, 2, 0.18), (2, 3, 0.14), (1, 4, 0.22), (4, 5, 0.36), (5, 6, 0.33), (1, 7, 0.22), (7, 8, 0.36), (8, 9, 0.33)]


def convert_dir_vec_to_pose(vec: np.ndarray) -> np.ndarray:
    """[..., 27] or [..., 9, 3] direction vectors -> [..., 10, 3] joint positions (data_utils.py:77-98)."""
    vec = np.asarray(vec, dtype=np.float64)
    if vec.shape[-1] != 3:
        vec = vec.reshape(vec.shape[:-1] + (-1, 3))
    pos = np.zeros(vec.shape[:-2] + (10, 3))
    for j, (a, b, length) in enumerate(DIR_VEC_PAIRS):
        pos[..., b, :] = pos[..., a, :] + length * vec[..., j, :]
    return pos


def batch_metrics(out_dir_vec: np.ndarray, target_vec: np.ndarray, mean_dir_vec: np.ndarray, n_pre_poses: int):
    """(l1 loss, joint MAE, accel) of one batch exactly as train.py:283,293-310 computes them ([B,T,27] float32 inputs)."""
    l1 = float(np.mean(np.abs(out_dir_vec.astype(np.float32) - target_vec.astype(np.float32)), dtype=np.float64))
    out = out_dir_vec.astype(np.float64) + np.asarray(mean_dir_vec, dtype=np.float64).squeeze()
    tgt = target_vec.astype(np.float64) + np.asarray(mean_dir_vec, dtype=np.float64).squeeze()
    op, tp = convert_dir_vec_to_pose(out), convert_dir_vec_to_pose(tgt)
    mae = float(np.mean(np.abs(op[:, n_pre_poses:] - tp[:, n_pre_poses:])))
    accel = float(np.mean(np.abs(np.diff(tp, n=2, axis=1) - np.diff(op, n=2, axis=1))))
    return l1, mae, accel
