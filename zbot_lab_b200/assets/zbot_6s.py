"""Physical model of the 6-module ZBOT serial chain (``ZBOT_6S_CFG``).

Single source of truth for the constants the CUDA step kernel, the CPU oracle and
the host-side env share.  Mirrors the *parameters* of

* ``/root/reference/source/zbot/zbot/assets/zbot_cfg.py:621-669`` (``ZBOT_6S_CFG``:
  init pose, implicit PD gains, effort limit), and
* ``/root/reference/source/zbot/zbot/assets/zbot_assets/zbot_6s_new.usd`` (link
  masses / CoMs / inertias, joint frames, collision extents; values as decoded in
  SURVEY.md Appendix A and cross-checked against the ASCII sibling
  ``zbot_6s_v04.usda:110-113,192-195,178-181``).

Everything derived here is float64; :func:`model_f32` rounds the derived table to
float32 so the float (GPU), double (CPU port) and numpy (oracle) code paths use
bit-identical constants.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------- #
# names / ordering
# --------------------------------------------------------------------------- #
#: articulation (kinematic traversal) order of the 12 links -- SURVEY A.1
LINK_NAMES = ("foot_0", "b1", "a2", "b2", "a3", "b3", "base", "b4", "a5", "b5", "a6", "foot_1")
#: ContactSensor body order = USD prim order under /zbot -- SURVEY A.1
SENSOR_BODY_NAMES = ("b1", "a2", "b2", "a3", "b3", "b4", "a5", "b5", "a6", "foot_0", "foot_1", "base")
JOINT_NAMES = ("joint1", "joint2", "joint3", "joint4", "joint5", "joint6")

NUM_LINKS = 12
NUM_JOINTS = 6
NUM_BODIES = 7  # reduced chain: fixed b_k -> a_{k+1} pairs merged

# link -> (reduced body index, is the second ("a") half of a merged pair)
LINK_TO_BODY = {
    "foot_0": (0, False),
    "b1": (1, False), "a2": (1, True),
    "b2": (2, False), "a3": (2, True),
    "b3": (3, False), "base": (3, True),
    "b4": (4, False), "a5": (4, True),
    "b5": (5, False), "a6": (5, True),
    "foot_1": (6, False),
}
A_TYPE_LINKS = ("foot_0", "a2", "a3", "base", "a5", "a6")
B_TYPE_LINKS = ("b1", "b2", "b3", "b4", "b5", "foot_1")

# --------------------------------------------------------------------------- #
# raw USD constants (zbot_6s_v04.usda:110-113, 192-195 -- identical in zbot_6s_new.usd)
# --------------------------------------------------------------------------- #
LINK_MASS = 0.25042
LINK_SPACING = 0.053  # every link frame is (0, 0, 0.053*k), identity orientation, at q = 0

A_COM = (-0.0082592, -5.1063e-8, 0.028345)
A_DIAG_INERTIA = (0.000220404, 0.00019972, 0.00029235598)
A_PRINCIPAL_AXES = (0.93171555, -8.355497e-7, 0.36318883, -0.0000021434983)  # wxyz

B_COM = (-0.011593, -5.1063e-8, 0.023274)
B_DIAG_INERTIA = (0.00022040308, 0.00019972, 0.0002923569)
B_PRINCIPAL_AXES = (0.9997794, -4.83283e-8, -0.021004679, 0.0000023001876)  # wxyz

SIN45 = math.sqrt(0.5)
#: joint axis in BOTH the parent and child link frame: (+-sin45, 0, cos45); "+" for joint1,3,5
JOINT_AXIS_SIGN = (+1.0, -1.0, +1.0, -1.0, +1.0, -1.0)

# actuator (ImplicitActuatorCfg, zbot_cfg.py:658-668)
KP = 50.0
KD = 5.0
EFFORT_LIMIT = 20.0

# initial state (zbot_cfg.py:641-655)
DEFAULT_ROOT_POS = (0.0, -0.06, 0.0)
DEFAULT_ROOT_QUAT = (1.0, 0.0, 0.0, 0.0)
DEFAULT_JOINT_POS = (0.312, 0.837, -2.02, 2.02, -0.837, -0.312)

# collision geometry
FOOT_DISC_RADIUS = 0.05
FOOT0_SOLE_Z = 0.0      # foot_0 ("a" link): sole = link-frame z = 0 plane
FOOT1_SOLE_Z = 0.053    # foot_1 ("b" link): sole = link-frame z = +0.053 plane
BODY_SPHERE_RADIUS = 0.05
BODY_SPHERE_Z = 0.053   # sphere centre of a merged b+a body, in the b-link frame (the a-link origin)
NUM_FOOT_POINTS = 4     # rim points per foot disc

GRAVITY = 9.81

# ---- ground-contact model (OUR choice; PhysX's is closed -- see DESIGN.md "Contact model") ----
# Each contact point carries a linearly-implicit soft constraint  f = F0 - dt*K*a_point,
# K = diag(beta, beta, gamma) in the world frame (ground normal = +z):
#   normal : spring k = ALPHA*ERP/dt on the PREDICTED penetration (pen - dt*v_z'), force cap
#            ALPHA*VDEP (max depenetration velocity, zbot_cfg.py:633), damper d = ALPHA*(1-ERP)
#            ramped in over RAMP metres of penetration; active iff the predictor force > 0, so the
#            law is continuous in position and velocity (no on/off jump at first touch)
#   tangent: regularised Coulomb, viscous coefficient beta = min(BETA_MAX, MU*f_n0/|v_t|)
CONTACT_ALPHA = 1000.0     # N s/m
CONTACT_ERP = 0.2
CONTACT_VDEP = 1.0         # m/s  (RigidBodyPropertiesCfg.max_depenetration_velocity)
CONTACT_BETA_MAX = 3000.0  # N s/m
CONTACT_MU = 1.0           # static = dynamic friction 1.0 x 1.0, "multiply" combine (env_v2.py:50-68)
CONTACT_RAMP = 5.0e-4      # m
CONTACT_VT_EPS = 1.0e-6    # m/s
CONTACT_MARGIN = 0.02      # m: a point may activate speculatively (predicted penetration) within this gap
SIM_DT = 1.0 / 200.0       # env_v2.py:48
DECIMATION = 4             # env_v2.py:40


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def quat_to_mat(q) -> np.ndarray:
    w, x, y, z = (float(v) for v in q)
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ]
    )


def _link_inertia(diag, axes) -> np.ndarray:
    R = quat_to_mat(axes)
    return R @ np.diag(diag) @ R.T


def _skew(v) -> np.ndarray:
    x, y, z = v
    return np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]], dtype=np.float64)


def _merge(m1, c1, I1, m2, c2, I2):
    """Parallel-axis merge of two rigid bodies expressed in one frame."""
    m = m1 + m2
    c = (m1 * c1 + m2 * c2) / m
    I = np.zeros((3, 3))
    for mi, ci, Ii in ((m1, c1, I1), (m2, c2, I2)):
        d = ci - c
        I += Ii + mi * (d @ d * np.eye(3) - np.outer(d, d))
    return m, c, I


@dataclass(frozen=True)
class ZbotModel:
    """Reduced 7-body model.  All arrays float64 holding float32-representable values."""

    body_mass: np.ndarray      # (7,)
    body_com: np.ndarray       # (7,3)   CoM in the body frame
    body_inertia: np.ndarray   # (7,3,3) about the CoM, body frame
    joint_pos: np.ndarray      # (6,3)   joint k+1 origin in the PARENT body (k) frame
    joint_axis: np.ndarray     # (6,3)   unit axis, same in parent and child frame
    foot_points: np.ndarray    # (2,P,3) rim points: [0] in body-0 frame, [1] in body-6 frame
    sphere_centre: np.ndarray  # (5,3)   sphere centre of bodies 1..5 (body frame)
    sphere_radius: float
    # per-LINK quantities needed by the MDP (link frame offset inside its body, CoM in link frame)
    link_offset: np.ndarray    # (12,3)  link-frame origin in its body frame (articulation order)
    link_com: np.ndarray       # (12,3)  CoM in the link frame
    link_body: np.ndarray      # (12,) int
    default_joint_pos: np.ndarray  # (6,)
    default_root_pos: np.ndarray   # (3,)
    kp: float
    kd: float
    effort_limit: float
    gravity: float
    # generalisations shared with the snake robot (zbot_d_6s.py); defaults describe the v2 robot
    name: str = "zbot_6s_new"
    default_root_quat: np.ndarray = None   # (4,) wxyz
    link_rot: np.ndarray = None            # (12,4) link-frame orientation relative to its body frame
    contact_list: np.ndarray = None        # (P,5): body, local x,y,z, drop (world-z offset, = sphere radius)
    link_names: tuple = LINK_NAMES
    sensor_body_names: tuple = SENSOR_BODY_NAMES
    self_pairs: tuple = ()                 # (link_i, link_j) pairs of the filtered self-contact sensors
    link_centre: np.ndarray = None         # (12,3) sphere centre used for the self-contact check (body frame)


def build_model(dtype=np.float32) -> ZbotModel:
    """Derive the reduced-coordinate model; round every table to ``dtype``."""
    Ia = _link_inertia(A_DIAG_INERTIA, A_PRINCIPAL_AXES)
    Ib = _link_inertia(B_DIAG_INERTIA, B_PRINCIPAL_AXES)
    ca = np.array(A_COM)
    cb = np.array(B_COM)
    off = np.array([0.0, 0.0, LINK_SPACING])

    mass = np.zeros(7)
    com = np.zeros((7, 3))
    inertia = np.zeros((7, 3, 3))
    mass[0], com[0], inertia[0] = LINK_MASS, ca, Ia
    mass[6], com[6], inertia[6] = LINK_MASS, cb, Ib
    mm, cm, Im = _merge(LINK_MASS, cb, Ib, LINK_MASS, ca + off, Ia)
    for k in range(1, 6):
        mass[k], com[k], inertia[k] = mm, cm, Im

    jpos = np.zeros((6, 3))
    jpos[0] = off            # joint1 in foot_0 frame: (0,0,0.053)
    jpos[1:] = 2.0 * off     # joints 2..6 in the b_k frame: a-link origin (0.053) + 0.053
    jaxis = np.array([[s * SIN45, 0.0, SIN45] for s in JOINT_AXIS_SIGN])

    P = NUM_FOOT_POINTS
    ang = 2.0 * math.pi * (np.arange(P) + 0.0) / P
    rim = np.stack([FOOT_DISC_RADIUS * np.cos(ang), FOOT_DISC_RADIUS * np.sin(ang), np.zeros(P)], -1)
    rim[np.abs(rim) < 1e-17] = 0.0
    foot_points = np.stack([rim + np.array([0, 0, FOOT0_SOLE_Z]), rim + np.array([0, 0, FOOT1_SOLE_Z])])
    sphere_centre = np.tile(np.array([0.0, 0.0, BODY_SPHERE_Z]), (5, 1))

    link_offset = np.zeros((12, 3))
    link_com = np.zeros((12, 3))
    link_body = np.zeros(12, dtype=np.int64)
    for i, name in enumerate(LINK_NAMES):
        b, second = LINK_TO_BODY[name]
        link_body[i] = b
        link_offset[i] = off if second else 0.0
        link_com[i] = ca if name in A_TYPE_LINKS else cb

    def r(x):
        return np.asarray(x, dtype=dtype).astype(np.float64)

    inertia = 0.5 * (inertia + inertia.transpose(0, 2, 1))
    # CAD noise: |Ixy|, |Iyz| < 1e-10 kg m^2 and CoM_y = -5.1e-8 m are zeroed so the kernel can
    # use the sparse (Ixx, Iyy, Izz, Ixz) / (cx, 0, cz) forms; oracle and kernel share this table.
    inertia[:, 0, 1] = inertia[:, 1, 0] = 0.0
    inertia[:, 1, 2] = inertia[:, 2, 1] = 0.0
    com[:, 1] = 0.0
    link_com[:, 1] = 0.0
    contact_list = []
    for f, b in ((0, 0), (1, 6)):
        for j in range(P):
            contact_list.append([b, *foot_points[f, j], 0.0])
    for b in range(1, 6):
        contact_list.append([b, *sphere_centre[b - 1], BODY_SPHERE_RADIUS])
    link_rot = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (12, 1))
    return ZbotModel(
        default_root_quat=np.array(DEFAULT_ROOT_QUAT, dtype=np.float64), link_rot=link_rot,
        contact_list=r(np.array(contact_list)),
        body_mass=r(mass), body_com=r(com), body_inertia=r(inertia),
        # joint_axis is the UNIT vector (+-sqrt(1/2), 0, sqrt(1/2)): kept in double so the joint
        # quaternions stay unit in the float64 instantiations; float code rounds it on use.
        joint_pos=r(jpos), joint_axis=jaxis, foot_points=r(foot_points),
        sphere_centre=r(sphere_centre), sphere_radius=float(dtype(BODY_SPHERE_RADIUS)),
        link_offset=r(link_offset), link_com=r(link_com), link_body=link_body,
        default_joint_pos=r(DEFAULT_JOINT_POS), default_root_pos=r(DEFAULT_ROOT_POS),
        kp=KP, kd=KD, effort_limit=EFFORT_LIMIT, gravity=float(dtype(GRAVITY)),
    )


_MODEL = None


def model_f32() -> ZbotModel:
    global _MODEL
    if _MODEL is None:
        _MODEL = build_model(np.float32)
    return _MODEL


# index helpers (resolve by NAME, never by assuming the two orders coincide)
def link_index(name: str) -> int:
    return LINK_NAMES.index(name)


def sensor_index(name: str) -> int:
    return SENSOR_BODY_NAMES.index(name)


def find_bodies(pattern: str, names=LINK_NAMES):
    """``Articulation.find_bodies``-style full-match regex lookup -> (ids, names)."""
    import re

    rx = re.compile(pattern)
    ids = [i for i, n in enumerate(names) if rx.fullmatch(n)]
    return ids, [names[i] for i in ids]


# --------------------------------------------------------------------------- #
# host-side kinematics (numpy, float64) -- used for constant tables (default-pose
# link poses) and by tests; the per-step kinematics run inside the CUDA kernel.
# --------------------------------------------------------------------------- #
def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    w1, x1, y1, z1 = (a[..., i] for i in range(4))
    w2, x2, y2, z2 = (b[..., i] for i in range(4))
    return np.stack(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ],
        axis=-1,
    )


def quat_rotate(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    xyz = q[..., 1:]
    t = 2.0 * np.cross(xyz, v)
    return v + q[..., 0:1] * t + np.cross(xyz, t)


def fk_links(root_pos, root_quat, joint_pos, model: ZbotModel | None = None):
    """Forward kinematics of all 12 links (articulation order).

    root_pos (...,3), root_quat (...,4 wxyz), joint_pos (...,6) -> link_pos (...,12,3),
    link_quat (...,12,4).  child = parent o T(joint_pos) o Rot(axis, q)   
    """
    m = model or model_f32()
    root_pos = np.asarray(root_pos, np.float64)
    root_quat = np.asarray(root_quat, np.float64)
    joint_pos = np.asarray(joint_pos, np.float64)
    bpos = [root_pos]
    bquat = [root_quat]
    for k in range(6):
        p = bpos[-1] + quat_rotate(bquat[-1], np.broadcast_to(m.joint_pos[k], root_pos.shape))
        half = 0.5 * joint_pos[..., k : k + 1]
        qj = np.concatenate([np.cos(half), np.sin(half) * m.joint_axis[k]], axis=-1)
        bpos.append(p)
        bquat.append(quat_mul(bquat[-1], qj))
    lp, lq = [], []
    for i in range(12):
        b = int(m.link_body[i])
        lp.append(bpos[b] + quat_rotate(bquat[b], np.broadcast_to(m.link_offset[i], root_pos.shape)))
        lq.append(quat_mul(bquat[b], np.broadcast_to(m.link_rot[i], bquat[b].shape)))
    return np.stack(lp, axis=-2), np.stack(lq, axis=-2)


def default_link_poses():
    """Link poses at the ``ZBOT_6S_CFG`` init state, relative to the env origin."""
    m = model_f32()
    return fk_links(m.default_root_pos, np.array(DEFAULT_ROOT_QUAT), m.default_joint_pos, m)
