"""shim of ``isaaclab_rl.rsl_rl`` (train.py:93, play.py:66)."""
import copy
import os

import torch

from zbot_lab_b200.envs.rsl_rl_wrapper import RslRlVecEnvWrapper  # noqa: F401
from zbot_lab_b200.tasks.zbot6b_direct.walking_v2_cfg import (  # noqa: F401
    PPORunnerCfgV2 as RslRlOnPolicyRunnerCfg,
    RslRlPpoActorCriticCfg,
    RslRlPpoAlgorithmCfg,
)
from zbot_lab_b200.utils.configclass import Cfg as RslRlBaseRunnerCfg  # noqa: F401


class _ExportedActor(torch.nn.Module):
    def __init__(self, policy, normalizer=None):
        super().__init__()
        self.actor = copy.deepcopy(policy.actor).cpu()
        self.normalizer = copy.deepcopy(normalizer).cpu() if normalizer is not None else torch.nn.Identity()

    def forward(self, x):
        return self.actor(self.normalizer(x))


def export_policy_as_jit(policy, normalizer, path: str, filename: str = "policy.pt"):
    """TorchScript export of the actor MLP (play.py:173)."""
    os.makedirs(path, exist_ok=True)
    m = _ExportedActor(policy, normalizer).eval()
    torch.jit.script(m).save(os.path.join(path, filename))


def export_policy_as_onnx(policy, normalizer, path: str, filename: str = "policy.onnx", verbose: bool = False):
    """ONNX export (play.py:175).  The ``onnx`` package is not installed in this image; when the exporter is
    unavailable a note is written next to where the file would be instead of failing the play loop."""
    os.makedirs(path, exist_ok=True)
    m = _ExportedActor(policy, normalizer).eval()
    n_in = m.actor[0].in_features
    try:
        torch.onnx.export(m, torch.zeros(1, n_in), os.path.join(path, filename), input_names=["obs"],
                          output_names=["actions"], opset_version=17, dynamo=False)
    except Exception as e:  # noqa: BLE001
        with open(os.path.join(path, filename + ".unavailable.txt"), "w") as f:
            f.write(f"ONNX export unavailable in this environment: {type(e).__name__}: {e}\n")
