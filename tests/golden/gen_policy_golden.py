"""Golden data for the end-to-end policy rollouts (CSV rows 15 `PPO-G` and 17 `new12800`).

Run in the build container (reads /root/reference):

    python tests/golden/gen_policy_golden.py

Writes policy_golden.npz:
  csv15, csv17 [100,4]   rows 15 / 17 of results/test_results/Real_{MK,PT,TT,IT}_J6_M6_E2_Seed3_Weight442.csv: final
                         (makespan, processing energy / N, transport time, idle time) of the greedy policy rollouts the
                         authors ran on their GPU for the 100 shipped test instances (test_all.py:660-667)
  iotj/op/<key>, iotj/mch/<key>     tester/IoTJ_MAPPO/PPO_{operation,machine}_actor_J6M6E2_1000.pth   (row 15's weights)
  n12800/op/<key>, n12800/mch/<key> trained_model/can_use/No_lr_decay/PPO_{job,machine}_actor_J6M6E2_top1.pth (row 17's)
The checkpoints are the reference's shipped artefacts (data, not source); they make the GPU test self-contained.
"""
import csv
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MTFJSP_REFERENCE_ROOT", "/root/reference")


def main():
    out = {}
    for row in (15, 17):
        a = np.zeros((100, 4))
        for k, name in enumerate(("MK", "PT", "TT", "IT")):
            rows = list(csv.reader(open(os.path.join(REF, "results/test_results/Real_%s_J6_M6_E2_Seed3_Weight442.csv" % name))))
            a[:, k] = [float(x) for x in rows[row][:100]]
        out["csv%d" % row] = a
    pairs = {"iotj": ("tester/IoTJ_MAPPO/PPO_operation_actor_J6M6E2_1000.pth", "tester/IoTJ_MAPPO/PPO_machine_actor_J6M6E2_1000.pth"),
             "n12800": ("trained_model/can_use/No_lr_decay/PPO_job_actor_J6M6E2_top1.pth",
                        "trained_model/can_use/No_lr_decay/PPO_machine_actor_J6M6E2_top1.pth")}
    for tag, (op, mch) in pairs.items():
        for part, path in (("op", op), ("mch", mch)):
            sd = torch.load(os.path.join(REF, path), map_location="cpu")
            for k, v in sd.items():
                out["%s/%s/%s" % (tag, part, k)] = v.numpy()
    path = os.path.join(HERE, "policy_golden.npz")
    np.savez_compressed(path, **out)
    print("%s %.0f KB, %d arrays" % (path, os.path.getsize(path) / 1024, len(out)))


if __name__ == "__main__":
    main()
