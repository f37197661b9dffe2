"""Batched PDR evaluation: the rule orders equal the ones the reference's own rule code produced
(recorded in tests/golden/pdr_golden.npz) and the device rollouts reproduce the shipped result CSV rows bit for bit."""
import importlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pdr_golden.npz")


def test_rule_orders_match_reference_rule_code():
    rules = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rules")
    g = np.load(GOLD)
    t, p = g["t"], g["p"]
    names = [str(x) for x in g["rule_names"]]
    for r, name in enumerate(names):
        o, m = name.split("+")
        order = rules.op_rule(o, t, p, 6, 6)
        mch = rules.machine_rule(m, t, p)
        np.testing.assert_array_equal(order, g["ops"][r].astype(np.int64), err_msg=name)
        np.testing.assert_array_equal(mch[np.arange(100)[:, None], order], g["mch"][r].astype(np.int64), err_msg=name)


@pytest.mark.gpu
def test_batched_rule_rollouts_reproduce_shipped_csv_rows():
    pytest.importorskip("torch")
    rules = importlib.import_module("e2e-mappo-for-mt-fjsp_b200.rules")
    g = np.load(GOLD)
    out = rules.evaluate_rules(g["t"], g["p"], g["transT"], g["edge"], 6, 6, 2)
    assert out["names"] == [str(x) for x in g["rule_names"]]
    np.testing.assert_array_equal(out["costs"], g["gold"])
