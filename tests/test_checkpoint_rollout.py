"""End-to-end parity of env + encoder + heads with the reference's SHIPPED checkpoints (the only reference
artefacts that exercise left shift + ESA mask + features + encoder together).

Greedy rollouts of the 100 shipped test instances (generator seed 3) with
  `iotj`   = tester/IoTJ_MAPPO/PPO_{operation,machine}_actor_J6M6E2_1000.pth  -> result CSV row 15 (`PPO-G`)
  `n12800` = trained_model/can_use/No_lr_decay/PPO_{job,machine}_actor_J6M6E2_top1.pth -> CSV row 17 (`new12800`)
(tests/golden/policy_golden.npz holds the weights and the CSV rows; gen_policy_golden.py made it) through
validate.greedy_validate -- one batch of 100 on the device, per-instance BatchNorm statistics as the reference's
batch-1 validation has them -- against
  (a) the authors' CSV rows (their GPU), and
  (b) the unmodified reference run on a B200 (tests/golden/policy_reference_b200.npz, written by
      profiles/checkpoint_parity.py: trainer/validate.py with the reference env and networks).
A greedy rollout is a chain of arg-max decisions, so FP32 rounding differences between GPUs / kernels can flip a
near-tie and change that instance's whole schedule: the reference itself reproduces 100 / 99 of the 100 CSV instances
on a B200.  Measured here (profiles/r02_checkpoint_parity.json): FP32 path 97 / 100 and 100 / 100 instances bit-identical
to the CSV, TF32 (tcgen05) path 97 / 97; mean objective within 0.12 %.  Tolerances below: FP32 >= 95 identical, TF32 >= 90
identical, mean objective within 0.5 %."""
import importlib
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
val = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.validate")


def _objective(x):
    return 0.4 * x[:, 0] + 0.4 * (x[:, 1] + x[:, 3]) + 0.2 * x[:, 2]


@pytest.mark.parametrize("tag,row", [("iotj", "csv15"), ("n12800", "csv17")])
@pytest.mark.parametrize("precision,min_same", [("fp32", 95), ("tf32", 90)])
def test_greedy_rollouts_with_shipped_checkpoints_reproduce_the_result_csv(tag, row, precision, min_same):
    g = np.load(os.path.join(GOLD, "policy_golden.npz"))
    pd = np.load(os.path.join(GOLD, "pdr_golden.npz"))
    inst = {k: pd[k] for k in ("t", "p", "transT", "edge")}
    sd_op, sd_m = val.load_actor_state_dicts(g, tag)          # strict: every key of the shipped state_dict is consumed
    job = enc.JobActor(sd_op, 6, 6, precision=precision)
    mch = enc.MachineActor(sd_m, 6, precision=precision)
    out = val.greedy_validate(job, mch, inst)
    mine, csv = out["final4"], g[row]
    same = int((mine == csv).all(axis=1).sum())
    assert same >= min_same, (tag, precision, same)
    rel = abs(_objective(mine).mean() - _objective(csv).mean()) / _objective(csv).mean()
    assert rel < 5e-3, (tag, precision, rel)
    np.testing.assert_allclose(out["objective"], _objective(mine), rtol=1e-12)
    ref = np.load(os.path.join(GOLD, "policy_reference_b200.npz"))[tag]
    same_ref = int((mine == ref).all(axis=1).sum())
    assert same_ref >= min_same - 1, (tag, precision, same_ref)
    # the reference on the B200 against the authors' numbers: the yardstick for what "same" can mean here
    assert int((ref == csv).all(axis=1).sum()) >= 99


def test_batch_statistics_over_the_whole_batch_are_a_different_policy():
    """Why per-instance BatchNorm groups matter: with statistics over all 100 instances (the training-rollout convention)
    no instance reproduces the batch-1 validation of the reference."""
    g = np.load(os.path.join(GOLD, "policy_golden.npz"))
    pd = np.load(os.path.join(GOLD, "pdr_golden.npz"))
    inst = {k: pd[k] for k in ("t", "p", "transT", "edge")}
    sd_op, sd_m = val.load_actor_state_dicts(g, "n12800")
    job, mch = enc.JobActor(sd_op, 6, 6), enc.MachineActor(sd_m, 6)
    whole = val.greedy_validate(job, mch, inst, per_instance_batchnorm=False)["final4"]
    assert int((whole == g["csv17"]).all(axis=1).sum()) < 50
    assert np.isfinite(whole).all()
