"""Shared helpers for the parity tests."""
import os

import numpy as np

from zbot_lab_b200.assets import zbot_6s as Z
from zbot_lab_b200.utils import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["mdp_v2_n64", "mdp_v2_n7", "mdp_v2_n300"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = syn.synth_mdp_case(int(g["seed"]), int(g["n"]), int(g["steps"]))
    return g, case


def make_mdp_oracle(n, origins):
    from oracle.mdp_oracle import MdpOracle
    return MdpOracle(n, origins, syn.reset_tables(), syn.index_sets(),
                     np.tile(np.asarray(Z.DEFAULT_JOINT_POS, np.float32), (n, 1)))


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(|b|, floor)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
