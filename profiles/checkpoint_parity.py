"""End-to-end parity of the encoder path with the reference's SHIPPED checkpoints.

Runs on the GPU box.  For the 100 shipped test instances (generator seed 3) and both shipped J6M6E2 checkpoint pairs
(`PPO-G` = tester/IoTJ_MAPPO/*_1000.pth -> CSV row 15; `new12800` = trained_model/can_use/No_lr_decay/*_top1.pth -> CSV
row 17) it compares the final (makespan, processing energy / N, transport, idle) of greedy rollouts from
  csv        the authors' shipped result rows (their GPU)
  reference  the UNMODIFIED reference on this box: trainer/validate.py:60-297 with the reference env (CPU, networkx) and
             the reference networks (FP32 on this GPU) -- only when baseline/_ref is staged
  fp32       validate.greedy_validate with JobActor / MachineActor(precision="fp32"), one batch of 100, per-instance BatchNorm
  tf32       the same with the tcgen05 TF32 actors (one instance per forward)
and writes profiles/r02_checkpoint_parity.json (+ tests/golden/policy_reference_b200.npz when the reference ran)."""
import contextlib
import importlib
import io
import json
import os
import random
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
enc = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.encoder")
val = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.validate")


def cmp(a, b):
    """exact matches (all four costs bit-equal), max relative deviation of the objective, mean objectives."""
    obj = lambda x: 0.4 * x[:, 0] + 0.4 * (x[:, 1] + x[:, 3]) + 0.2 * x[:, 2]
    same = int((a == b).all(axis=1).sum())
    oa, ob = obj(a), obj(b)
    return {"instances_identical": same, "of": int(a.shape[0]), "max_rel_objective_dev": float(np.max(np.abs(oa - ob) / ob)),
            "mean_objective": [float(oa.mean()), float(ob.mean())]}


def run_reference(ds, tag):
    from oracle import ref_harness as rh

    if not rh.reference_available():
        return None
    paths = {"iotj": rh.shipped_checkpoint_paths(6, 6, 2, 1000),
             "n12800": tuple(os.path.join(rh.REF_ROOT, "trained_model/can_use/No_lr_decay/PPO_%s_actor_J6M6E2_top1.pth" % k)
                             for k in ("job", "machine"))}[tag]
    ppo, args = rh.make_reference_ppo(6, 6, "cuda:0", *paths)
    with contextlib.redirect_stdout(io.StringIO()):
        from trainer import validate
    res = []
    for i in range(100):
        random.seed(i)
        with contextlib.redirect_stdout(io.StringIO()):
            _, final4, _ = validate.validate_cost_gcn_jointActor_GAT(ppo, False, ds, i, "random", greedy=True, args=args)
        res.append(final4)
    return np.array(res, dtype=np.float64)


def main():
    g = np.load(os.path.join(GOLD, "policy_golden.npz"))
    pd = np.load(os.path.join(GOLD, "pdr_golden.npz"))
    inst = {k: pd[k] for k in ("t", "p", "transT", "edge")}
    ds = types.SimpleNamespace(**inst)
    report, refs = {}, {}
    for tag, row in (("iotj", "csv15"), ("n12800", "csv17")):
        sd_op, sd_m = val.load_actor_state_dicts(g, tag)
        csv = g[row]
        out = {}
        ref = run_reference(ds, tag)
        if ref is not None:
            refs[tag] = ref
            out["reference_on_b200_vs_csv"] = cmp(ref, csv)
        for prec in ("fp32", "tf32"):
            job = enc.JobActor(sd_op, 6, 6, precision=prec)
            mch = enc.MachineActor(sd_m, 6, precision=prec)
            mine = val.greedy_validate(job, mch, inst)["final4"]
            out[prec + "_vs_csv"] = cmp(mine, csv)
            if ref is not None:
                out[prec + "_vs_reference_on_b200"] = cmp(mine, ref)
            if prec == "fp32":
                whole = val.greedy_validate(job, mch, inst, per_instance_batchnorm=False)["final4"]
                out["fp32_batch_statistics_over_all_100_vs_csv"] = cmp(whole, csv)
        report[tag] = out
        print(tag, json.dumps(out, indent=1))
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "r02_checkpoint_parity.json"), "w"), indent=1)
    if refs:
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "policy_reference_b200.npz"), **refs)


if __name__ == "__main__":
    main()
