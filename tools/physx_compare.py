#!/usr/bin/env python
"""Dump a 50-step PhysX trajectory of ``zbot-6b-walking-v2`` for the third parity level of BASELINE.json's north_star
("dynamics must track the reference's PhysX CPU articulation within a stated joint-position and base-pose tolerance over a
fixed 50-step horizon from identical initial states").

PhysX / Isaac Sim are closed and absent from the build and benchmark boxes, so this repo cannot produce that fixture
itself.  This script is the recipe for anyone who has an Isaac Sim + Isaac Lab install with the reference checked out:

    ./isaaclab.sh -p /path/to/this/repo/tools/physx_compare.py --reference /path/to/zbot_lab \
        --out /path/to/this/repo/tests/golden/physx_v2_traj.npz [--num_envs 64] [--device cpu]

It (1) builds the reference's own task (`gym.make("zbot-6b-walking-v2")`, CPU PhysX pipeline by default), (2) writes the
repo's seeded synthetic articulation states (`zbot_lab_b200.utils.synthetic.synth_sim_state`, seed 1234: the states
bench.py and the dynamics tests use) into the simulation, (3) steps 50 control steps with the repo's seeded action stream
through the reference's `env.step`, and (4) saves, per step, what the reference's MDP reads: root pose / velocity, joint
positions / velocities, `applied_torque`, the 12-link poses, the feet rows of `net_forces_w_history`, rewards and flags.

`tests/test_physx_fixture.py` consumes the file when it exists (and is skipped otherwise): it replays the same states and
actions through the CUDA step and reports / bounds the per-step deviation.  Nothing in this file is imported by the product.
"""
import argparse
import os
import sys

import numpy as np

HORIZON = 50
SEED = 1234


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--reference", required=True, help="checkout of crowznl/zbot_lab (its source/zbot is pip-installed or on PYTHONPATH)")
    p.add_argument("--out", required=True)
    p.add_argument("--num_envs", type=int, default=64)
    p.add_argument("--device", default="cpu", help="cpu = PhysX CPU pipeline (the north_star's comparator); cuda:0 also works")
    args, rest = p.parse_known_args()

    from isaaclab.app import AppLauncher                                   # noqa: E402  (Isaac Lab must be importable)
    app = AppLauncher(headless=True, device=args.device).app

    import gymnasium as gym                                                # noqa: E402
    import torch                                                           # noqa: E402
    sys.path.insert(0, os.path.join(args.reference, "source", "zbot"))
    import zbot.tasks  # noqa: F401,E402  (registers zbot-6b-walking-v2)
    from isaaclab_tasks.utils import parse_env_cfg                         # noqa: E402

    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, here)
    from zbot_lab_b200.utils import synthetic as syn                       # noqa: E402  (pure numpy)

    n = args.num_envs
    cfg = parse_env_cfg("zbot-6b-walking-v2", device=args.device, num_envs=n)
    cfg.seed = SEED
    env = gym.make("zbot-6b-walking-v2", cfg=cfg).unwrapped
    env.reset()
    rng = np.random.default_rng(SEED)
    st = syn.synth_sim_state(rng, n)                                       # env-LOCAL root pose + joints
    actions = rng.normal(0.0, 1.0, (HORIZON, n, 6)).astype(np.float32)
    dev = env.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    robot = env._robot
    ids = robot._ALL_INDICES
    root = torch.cat([t(st["root_pos"]) + env._terrain.env_origins, t(st["root_quat"])], -1)
    robot.write_root_pose_to_sim(root, ids)
    robot.write_root_velocity_to_sim(torch.cat([t(st["root_lin_vel"]), t(st["root_ang_vel"])], -1), ids)
    robot.write_joint_state_to_sim(t(st["joint_pos"]), t(st["joint_vel"]), None, ids)
    env.scene.write_data_to_sim()
    env.sim.forward()
    env.episode_length_buf[:] = 0
    env._get_observations()                                                # refresh the cached ("stale") tensors from the new state

    out = {"seed": SEED, "num_envs": n, "horizon": HORIZON, "sim_dt": float(env.physics_dt), "decimation": int(cfg.decimation),
           "joint_names": np.array(robot.joint_names), "body_names": np.array(robot.body_names),
           "sensor_body_names": np.array(env._contact_sensor.body_names),
           "init/" + "root_pos": st["root_pos"], "init/root_quat": st["root_quat"], "init/root_lin_vel": st["root_lin_vel"],
           "init/root_ang_vel": st["root_ang_vel"], "init/joint_pos": st["joint_pos"], "init/joint_vel": st["joint_vel"],
           "actions": actions, "env_origins": env._terrain.env_origins.cpu().numpy()}
    rec = {k: [] for k in ("root_pos", "root_quat", "root_lin_vel", "root_ang_vel", "joint_pos", "joint_vel", "applied_torque",
                           "body_link_pos", "body_link_quat", "feet_force_hist", "last_air_time", "obs", "rew", "terminated", "truncated")}
    feet_ids = env._feet_ids
    for k in range(HORIZON):
        obs, rew, term, trunc, _ = env.step(t(actions[k]))
        d = robot.data
        org = env._terrain.env_origins
        rec["root_pos"].append((d.root_pos_w - org).cpu().numpy())
        rec["root_quat"].append(d.root_quat_w.cpu().numpy())
        rec["root_lin_vel"].append(d.root_lin_vel_w.cpu().numpy())
        rec["root_ang_vel"].append(d.root_ang_vel_w.cpu().numpy())
        rec["joint_pos"].append(d.joint_pos.cpu().numpy())
        rec["joint_vel"].append(d.joint_vel.cpu().numpy())
        rec["applied_torque"].append(d.applied_torque.cpu().numpy())
        rec["body_link_pos"].append((d.body_link_pos_w - org.unsqueeze(1)).cpu().numpy())
        rec["body_link_quat"].append(d.body_link_quat_w.cpu().numpy())
        rec["feet_force_hist"].append(env._contact_sensor.data.net_forces_w_history[:, :, feet_ids].cpu().numpy())
        rec["last_air_time"].append(env._contact_sensor.data.last_air_time[:, feet_ids].cpu().numpy())
        rec["obs"].append(obs["policy"].cpu().numpy())
        rec["rew"].append(rew.cpu().numpy())
        rec["terminated"].append(term.cpu().numpy())
        rec["truncated"].append(trunc.cpu().numpy())
    for k, v in rec.items():
        out["traj/" + k] = np.stack(v)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.startswith("traj/")})
    env.close()
    app.close()


if __name__ == "__main__":
    main()
