"""TEST INFRASTRUCTURE / CPU BASELINE -- ctypes loader of ``oracle/cpu_port.cpp``.

Builds ``oracle/_build/libzbot_cpu_port.so`` on demand with g++ (portable ``-march=x86-64-v3``)
and exposes the float / double CPU instantiations of the per-env step math.  Only
``tests/``, ``__graft_entry__`` and ``bench.py``'s cpu_baseline / reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from zbot_lab_b200.native import M_NUM_RAND, ZbotCfg, make_cfg  # noqa: F401  (struct definition / constants only)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libzbot_cpu_port.so")
SRCS = [os.path.join(HERE, "cpu_port.cpp"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_core.h"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_layout.h"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_model_constants.h"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_pair.h"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_halves.h"),
        os.path.join(ROOT, "zbot_lab_b200", "csrc", "zbot_h2.h"),
        os.path.join(ROOT, "include", "zbot_b200.h")]


def build(force: bool = False) -> str:
    stale = force or not os.path.isfile(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRCS)
    if stale:
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
        cmd = [cxx, "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c++17", "-ffp-contract=off",
               "-shared", "-o", SO, SRCS[0]]
        subprocess.run(cmd, check=True, cwd=HERE)
    return SO


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.zbot_port_state_word.argtypes = [C.c_char_p]
    return _LIB


def set_threads(n: int | None = None) -> int:
    """Use ``n`` OpenMP threads (default: every host core); returns the thread count in effect."""
    return int(lib().zbot_port_set_threads(C.c_int(n if n else (os.cpu_count() or 1))))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PortTerrain(C.Structure):
    """Host twin of ``zbot_bind_terrain`` (struct PortTerrain in cpu_port.cpp)."""
    _fields_ = [("heights", C.c_void_p), ("nx", C.c_int), ("ny", C.c_int), ("x0", C.c_float), ("y0", C.c_float),
                ("cell", C.c_float), ("tile_origins", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int),
                ("tile_size", C.c_float), ("env_origins", C.c_void_p), ("curriculum", C.c_int)]


def make_port_terrain(terrain, env_origins4: np.ndarray, curriculum: bool):
    """``terrain`` = zbot_lab_b200.terrain.Terrain; ``env_origins4`` (N, 4) float32, updated in place by the steps."""
    assert env_origins4.dtype == np.float32 and env_origins4.flags.c_contiguous and env_origins4.shape[1] == 4
    pt = PortTerrain(_p(terrain.heights), terrain.heights.shape[0], terrain.heights.shape[1], terrain.x0, terrain.y0, terrain.cell,
                     _p(terrain.origins), terrain.rows, terrain.cols, float(terrain.cfg.size[0]), _p(env_origins4), int(bool(curriculum)))
    pt._keep = (terrain, env_origins4)
    return pt


class PortEnv:
    """N envs stepped on the CPU with the kernel's own arithmetic (float32 or float64)."""

    def __init__(self, n, dtype=np.float32, cfg: ZbotCfg | None = None):
        self.n = n
        self.dtype = np.dtype(dtype)
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        self.cfg = cfg or make_cfg(n)
        self.snake = (self.cfg.task == 1)
        self.v4 = (self.cfg.task == 2)
        self.mtask = (self.cfg.task == 3)
        self.state = np.zeros((n, 80), self.dtype)
        self.ep_len = np.zeros(n, np.int64)
        self.export_words = (lib().zbot_port_snake_export_words() if self.snake else 69 if self.v4 else 72 if self.mtask
                             else getattr(lib(), "zbot_port_export_words_" + self.sfx)())
        self.reset_all()

    def word(self, name):
        w = lib().zbot_port_state_word(name.encode())
        assert w >= 0, name
        return w

    def field(self, name, width):
        w = self.word(name)
        return self.state[:, w:w + width]

    def reset_all(self):
        from zbot_lab_b200.assets import zbot_6s as Z
        if self.snake:
            from zbot_lab_b200.assets import zbot_d_6s as S
            m = S.model_f32()
            self.state[:] = 0
            self.field("root_pos", 3)[:] = m.default_root_pos
            self.field("root_quat", 4)[:] = m.default_root_quat
            self.field("joint_speed_limit", 1)[:] = np.pi
            self.ep_len[:] = 0
            return
        if self.mtask:
            from zbot_lab_b200.assets import zbot_6s_v2 as V
            m = V.model_f32()
            self.state[:] = 0
            self.field("root_pos", 3)[:] = m.default_root_pos
            self.field("root_quat", 4)[:] = m.default_root_quat
            self.field("joint_pos", 6)[:] = m.default_joint_pos
            self.field("joint_speed_limit", 1)[:] = self.cfg.contact_mu      # per-env friction coefficient
            lp, _ = V.default_link_poses()
            self.field("feet_down_pos_last", 6)[:] = np.concatenate([lp[V.link_index("foot0")], lp[V.link_index("foot1")]])
            self.ep_len[:] = 0
            return
        self.state[:] = 0
        self.field("root_pos", 3)[:] = Z.model_f32().default_root_pos
        self.field("root_quat", 4)[:] = Z.DEFAULT_ROOT_QUAT
        self.field("joint_pos", 6)[:] = Z.model_f32().default_joint_pos
        self.field("joint_speed_limit", 1)[:] = 1.0
        lp, _ = Z.default_link_poses()
        self.field("feet_down_pos_last", 6)[:] = np.concatenate([lp[0], lp[11]])
        self.ep_len[:] = 0

    def set_sim_state(self, st):
        for k, (name, w) in {"root_pos": ("root_pos", 3), "root_quat": ("root_quat", 4),
                             "root_lin_vel": ("root_lin_vel", 3), "root_ang_vel": ("root_ang_vel", 3),
                             "joint_pos": ("joint_pos", 6), "joint_vel": ("joint_vel", 6)}.items():
            self.field(name, w)[:] = st[k]

    def step(self, actions, export=False, rnd=None):
        n = self.n
        a = np.ascontiguousarray(actions, self.dtype)
        if self.mtask:
            obs = np.zeros((n, 25), self.dtype)
            rew, term, trunc = np.zeros(n, self.dtype), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
            rs = np.zeros((n, 20), self.dtype)      # 16 term slots + is_terminated sum, base_height / feet_close / illegal_contact flags
            ex = np.zeros((n, self.export_words), self.dtype) if export else None
            r = np.ascontiguousarray(rnd, self.dtype)
            assert r.shape == (n, M_NUM_RAND)
            pt = getattr(self, "terrain", None)
            if pt is not None:
                rc = getattr(lib(), "zbot_port_m_step_terrain_" + self.sfx)(
                    C.byref(self.cfg), _p(self.state), _p(self.ep_len), _p(a), _p(r), _p(obs), _p(rew), _p(term), _p(trunc),
                    _p(rs), _p(ex), C.c_int(n), C.byref(pt))
            else:
                rc = getattr(lib(), "zbot_port_m_step_" + self.sfx)(
                    C.byref(self.cfg), _p(self.state), _p(self.ep_len), _p(a), _p(r), _p(obs), _p(rew), _p(term), _p(trunc),
                    _p(rs), _p(ex), C.c_int(n))
            assert rc == 0, rc
            return obs, rew, term.astype(bool), trunc.astype(bool), rs, ex
        if self.v4:
            obs = np.zeros((n, 24), self.dtype)
            rew, term, trunc = np.zeros(n, self.dtype), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
            rs = np.zeros((n, 16), self.dtype)
            ex = np.zeros((n, self.export_words), self.dtype) if export else None
            r = np.ascontiguousarray(rnd, self.dtype)
            assert r.shape == (n, 10)
            rc = getattr(lib(), "zbot_port_v4_step_" + self.sfx)(
                C.byref(self.cfg), _p(self.state), _p(self.ep_len), _p(a), _p(r), _p(obs), _p(rew), _p(term), _p(trunc),
                _p(rs), _p(ex), C.c_int(n))
            assert rc == 0, rc
            return obs, rew, term.astype(bool), trunc.astype(bool), rs, ex
        obs = np.zeros((n, 23), self.dtype)
        rew = np.zeros(n, self.dtype)
        term = np.zeros(n, np.uint8)
        trunc = np.zeros(n, np.uint8)
        rs = np.zeros((n, 16), self.dtype)
        ex = np.zeros((n, self.export_words), self.dtype) if export else None
        fn = getattr(lib(), ("zbot_port_snake_step_" if self.snake else "zbot_port_step_") + self.sfx)
        rc = fn(C.byref(self.cfg), _p(self.state), _p(self.ep_len), _p(a), _p(obs), _p(rew), _p(term),
                _p(trunc), _p(rs), _p(ex), C.c_int(n))
        assert rc == 0, rc
        return obs, rew, term.astype(bool), trunc.astype(bool), rs, ex


def substeps(sim: np.ndarray, target: np.ndarray, nsub: int, cfg: ZbotCfg | None = None, snake: bool = False,
             model: str | None = None, terrain: "PortTerrain | None" = None):
    """sim [N][25] (in/out), target [N][6] -> forces [N][7][3], applied torque [N][6]."""
    n = sim.shape[0]
    dt = sim.dtype
    sfx = "f32" if dt == np.float32 else "f64"
    cfg = cfg or make_cfg(n)
    forces = np.zeros((n, 7, 3), dt)
    tau = np.zeros((n, 6), dt)
    target = np.ascontiguousarray(target, dt)
    name = ("zbot_port_substeps_m_" if model == "m" else "zbot_port_substeps_halves_" if model == "halves" else
            "zbot_port_substeps_h2_" if model == "h2" else
            "zbot_port_substeps_snake_" if snake else "zbot_port_substeps_")
    if terrain is not None:
        assert model == "m"
        rc = getattr(lib(), "zbot_port_substeps_m_terrain_" + sfx)(C.byref(cfg), _p(sim), _p(target), _p(forces), _p(tau),
                                                                   C.c_int(n), C.c_int(nsub), C.byref(terrain))
    else:
        rc = getattr(lib(), name + sfx)(C.byref(cfg), _p(sim), _p(target), _p(forces), _p(tau),
                                        C.c_int(n), C.c_int(nsub))
    assert rc == 0
    return forces, tau


def substeps_pair(sim: np.ndarray, target: np.ndarray, nsub: int, cfg: ZbotCfg | None = None, snake: bool = False):
    """float32 substeps through the two-lane instantiation (T = F2, csrc/zbot_pair.h): envs (2i, 2i+1) share a call chain."""
    n = sim.shape[0]
    assert sim.dtype == np.float32
    cfg = cfg or make_cfg(n)
    forces = np.zeros((n, 7, 3), np.float32)
    tau = np.zeros((n, 6), np.float32)
    target = np.ascontiguousarray(target, np.float32)
    rc = lib().zbot_port_substeps_pair_f32(C.byref(cfg), _p(sim), _p(target), _p(forces), _p(tau), C.c_int(n), C.c_int(nsub),
                                           C.c_int(1 if snake else 0))
    assert rc == 0
    return forces, tau


def link_view(sim: np.ndarray):
    n = sim.shape[0]
    dt = sim.dtype
    sfx = "f32" if dt == np.float32 else "f64"
    pos = np.zeros((n, 12, 3), dt)
    quat = np.zeros((n, 12, 4), dt)
    vel = np.zeros((n, 12, 3), dt)
    getattr(lib(), "zbot_port_link_view_" + sfx)(_p(sim), _p(pos), _p(quat), _p(vel), C.c_int(n))
    return pos, quat, vel


def pack_sim(st: dict, dtype=np.float64) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate(
        [st["root_pos"], st["root_quat"], st["root_lin_vel"], st["root_ang_vel"], st["joint_pos"],
         st["joint_vel"]], axis=-1), dtype)
