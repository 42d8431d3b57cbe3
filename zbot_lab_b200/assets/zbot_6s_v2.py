"""Physical model of the 6-module ZBOT biped of the manager-based task (``ZBOT_6S_V2_CFG``).

Mirrors the parameters of ``/root/reference/source/zbot/zbot/assets/zbot_cfg.py:959-1005`` (init pose (0,0,0.2545),
identity rotation, joint1/2/3/7/8/9 = 2.02 / -0.837 / -0.312 / -2.02 / 0.837 / 0.312, implicit PD kp 20 / kd 0.5 /
effort 20) and of the USD it names, ``zbot_assets/zbot_6s_v09.usd`` -- a binary crate decoded with
``tools/usdc_dump.py`` (build-container tool); the decoded values are the tables below.

Structure of that file: articulation root = link ``base``; ``base =FixedJoint0=> a7``; one branch
``base -joint1-> b1 => a2 -joint2-> b2 => a3 -joint3-> foot0`` running along -y, the other
``a7 -joint7-> b7 => a8 -joint8-> b8 => a9 -joint9-> foot1`` along +y; every link frame has identity orientation at
q = 0 except ``base`` (rotated -64 deg about y); link origins 0.053 m apart on the y axis; revolute axis = joint-frame Y,
with ``localRot`` = rotations about z, i.e. the axis lies in the link xy-plane at 45 deg to the chain.  The authored mass
properties are the SAME numbers as in ``zbot_6s_new.usd`` (CoM (cx, 0, cz), principal axes about y) although the chain
now runs along y: they are used AS AUTHORED (PhysX does), so the CoM sits 28 mm off the chain axis and the bodies are not
symmetric about the chain's plane any more -> the reduced model carries full CoM vectors and full inertia tensors.

It is the same kinematic chain as the walking robot of ``zbot_6s.py``: seen from the sole of ``foot0`` the joints sit at
0.053, 0.159, ... 0.583 m, total length 0.636 m, joint axes alternate (+-sqrt(1/2), 0, sqrt(1/2)) in the *chain frame*
C = (x_c, y_c, z_c) = (-x_link, z_link, y_link).  The reduced model therefore keeps the walking robot's convention --
root body = ``foot0`` with its origin at the sole centre, body frames = chain frame, chain joint k = USD joints
(joint3, joint2, joint1, joint7, joint8, joint9)[k] with the SAME angle value (the three joints below the base are
traversed child -> parent, the axis vector flips with the direction, the angle does not) -- and carries the constant
link rotations (``link_rot``) for what the MDP reads per link.  The free-floating dynamics do not depend on which body is
called the root; the articulation root of the USD (``base``) only fixes what ``root_*`` means in the MDP.
"""
from __future__ import annotations

import math

import numpy as np

from . import zbot_6s as Z

# --------------------------------------------------------------------------------------------------------------
# decoded from zbot_6s_v09.usd (tools/usdc_dump.py); link xformOp:translate / orient at q = 0, world = /zbot frame
# --------------------------------------------------------------------------------------------------------------
#: articulation (breadth-first from the root, Isaac Lab / PhysX convention [IL-upstream]) body order
LINK_NAMES = ("base", "a7", "b1", "b7", "a2", "a8", "b2", "b8", "a3", "a9", "foot0", "foot1")
#: Isaac Lab joint order (breadth-first) [IL-upstream]: what ``joint_pos`` / actions columns mean
JOINT_NAMES = ("joint1", "joint7", "joint2", "joint8", "joint3", "joint9")
#: USD joint of chain joint k (chain = foot0 -> foot1)
CHAIN_JOINTS = ("joint3", "joint2", "joint1", "joint7", "joint8", "joint9")
#: column of the Isaac Lab joint vector that holds chain joint k
CHAIN_TO_IL = tuple(JOINT_NAMES.index(n) for n in CHAIN_JOINTS)

LINK_Y = {"foot0": -0.265, "a3": -0.212, "b2": -0.159, "a2": -0.106, "b1": -0.053, "base": 0.0, "a7": 0.0,
          "b7": 0.053, "a8": 0.106, "b8": 0.159, "a9": 0.212, "foot1": 0.265}
BASE_LINK_QUAT = (0.848048096156426, 0.0, -0.5299192642332049, 0.0)      # /zbot/base.xformOp:orient (wxyz)
A_TYPE_LINKS = ("base", "a7", "a2", "a3", "a8", "a9")
B_TYPE_LINKS = ("b1", "b2", "b7", "b8", "foot0", "foot1")
#: reduced chain body of every link (fixed joints merged)
LINK_TO_BODY = {"foot0": 0, "a3": 1, "b2": 1, "a2": 2, "b1": 2, "base": 3, "a7": 3, "b7": 4, "a8": 4, "b8": 5, "a9": 5,
                "foot1": 6}
JOINT_Y = (-0.265, -0.159, -0.053, 0.053, 0.159, 0.265)                   # chain joints 0..5 on the world y axis
SOLE_Y = (-0.318, 0.318)                                                  # flat disc r = 0.05 of foot0 / foot1 (mesh points)

KP = 20.0
KD = 0.5
EFFORT_LIMIT = 20.0
DEFAULT_ROOT_POS = (0.0, 0.0, 0.2545)            # pose of the articulation root link `base`
DEFAULT_ROOT_QUAT = (1.0, 0.0, 0.0, 0.0)
DEFAULT_JOINT_POS_USD = {"joint1": 2.02, "joint2": -0.837, "joint3": -0.312, "joint7": -2.02, "joint8": 0.837, "joint9": 0.312}

#: chain frame axes written in link / world coordinates at q = 0 (columns x_c, y_c, z_c)
CHAIN_IN_WORLD = np.array([[-1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0]])


def _mat_to_quat(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[1 + i] = 0.25 * s
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def build_model(dtype=np.float32) -> Z.ZbotModel:
    C = CHAIN_IN_WORLD
    Ia = Z._link_inertia(Z.A_DIAG_INERTIA, Z.A_PRINCIPAL_AXES)     # same authored numbers as zbot_6s_new.usd
    Ib = Z._link_inertia(Z.B_DIAG_INERTIA, Z.B_PRINCIPAL_AXES)
    ca, cb = np.array(Z.A_COM), np.array(Z.B_COM)
    # body origins on the world y axis: body 0 at the foot0 sole centre, body k >= 1 at chain joint k-1
    body_y = (SOLE_Y[0],) + JOINT_Y
    parts = {b: [] for b in range(7)}
    link_offset, link_com, link_rot, link_body = np.zeros((12, 3)), np.zeros((12, 3)), np.zeros((12, 4)), np.zeros(12, np.int64)
    for i, name in enumerate(LINK_NAMES):
        b = LINK_TO_BODY[name]
        Rl = Z.quat_to_mat(BASE_LINK_QUAT) if name == "base" else np.eye(3)     # link frame in the world at q = 0
        c_l, I_l = (ca, Ia) if name in A_TYPE_LINKS else (cb, Ib)
        org_w = np.array([0.0, LINK_Y[name] - body_y[b], 0.0])                 # link origin relative to the body origin
        Rcl = C.T @ Rl                                                         # link axes in the chain (= body) frame
        parts[b].append((Z.LINK_MASS, C.T @ org_w + Rcl @ c_l, Rcl @ I_l @ Rcl.T))
        link_offset[i], link_com[i], link_rot[i], link_body[i] = C.T @ org_w, c_l, _mat_to_quat(Rcl), b
    mass, com, inertia = np.zeros(7), np.zeros((7, 3)), np.zeros((7, 3, 3))
    for b in range(7):
        if len(parts[b]) == 1:
            mass[b], com[b], inertia[b] = parts[b][0]
        else:
            (m1, c1, I1), (m2, c2, I2) = parts[b]
            mass[b], com[b], inertia[b] = Z._merge(m1, c1, I1, m2, c2, I2)
    inertia = 0.5 * (inertia + inertia.transpose(0, 2, 1))

    jpos = np.zeros((6, 3))
    for k in range(6):
        jpos[k] = (0.0, 0.0, JOINT_Y[k] - body_y[k])                           # 0.053, then 0.106: same as zbot_6s.py
    jaxis = np.array([[s * Z.SIN45, 0.0, Z.SIN45] for s in Z.JOINT_AXIS_SIGN])

    P = Z.NUM_FOOT_POINTS
    ang = 2.0 * math.pi * np.arange(P) / P
    rim = np.stack([Z.FOOT_DISC_RADIUS * np.cos(ang), Z.FOOT_DISC_RADIUS * np.sin(ang), np.zeros(P)], -1)
    rim[np.abs(rim) < 1e-17] = 0.0
    # soles in the body frames: body 0 origin IS the foot0 sole centre; foot1's sole is 0.053 beyond joint9
    foot_points = np.stack([rim, rim + np.array([0.0, 0.0, SOLE_Y[1] - JOINT_Y[5]])])
    sphere_centre = np.tile(np.array([0.0, 0.0, Z.BODY_SPHERE_Z]), (5, 1))
    contact_list = []
    for f, b in ((0, 0), (1, 6)):
        for j in range(P):
            contact_list.append([b, *foot_points[f, j], 0.0])
    for b in range(1, 6):
        contact_list.append([b, *sphere_centre[b - 1], Z.BODY_SPHERE_RADIUS])

    def r(x):
        return np.asarray(x, dtype=dtype).astype(np.float64)

    q_chain = np.array([DEFAULT_JOINT_POS_USD[n] for n in CHAIN_JOINTS])
    m = Z.ZbotModel(
        name="zbot_6s_v09", default_root_quat=np.array([1.0, 0.0, 0.0, 0.0]), link_rot=link_rot,
        contact_list=r(np.array(contact_list)), body_mass=r(mass), body_com=r(com), body_inertia=r(inertia),
        joint_pos=r(jpos), joint_axis=jaxis, foot_points=r(foot_points), sphere_centre=r(sphere_centre),
        sphere_radius=float(dtype(Z.BODY_SPHERE_RADIUS)), link_offset=r(link_offset), link_com=r(link_com),
        link_body=link_body, default_joint_pos=r(q_chain), default_root_pos=np.zeros(3), kp=KP, kd=KD,
        effort_limit=EFFORT_LIMIT, gravity=float(dtype(Z.GRAVITY)), link_names=LINK_NAMES,
        sensor_body_names=LINK_NAMES)
    # the chain root (foot0 sole frame) pose that puts the `base` LINK at DEFAULT_ROOT_POS with identity rotation
    lp, lq = Z.fk_links(np.zeros(3), np.array([1.0, 0.0, 0.0, 0.0]), q_chain, m)
    ib = LINK_NAMES.index("base")
    q_inv = lq[ib] * np.array([1.0, -1.0, -1.0, -1.0])
    root_quat = q_inv / np.linalg.norm(q_inv)
    root_pos = np.array(DEFAULT_ROOT_POS) - Z.quat_rotate(root_quat, lp[ib])
    object.__setattr__(m, "default_root_pos", r(root_pos))
    object.__setattr__(m, "default_root_quat", r(root_quat) / np.linalg.norm(r(root_quat)))
    return m


_MODEL = None


def model_f32() -> Z.ZbotModel:
    global _MODEL
    if _MODEL is None:
        _MODEL = build_model(np.float32)
    return _MODEL


def link_index(name: str) -> int:
    return LINK_NAMES.index(name)


def default_link_poses():
    """Link poses at the ``ZBOT_6S_V2_CFG`` init state (base link at (0,0,0.2545), identity), env-local."""
    m = model_f32()
    return Z.fk_links(m.default_root_pos, m.default_root_quat, m.default_joint_pos, m)
