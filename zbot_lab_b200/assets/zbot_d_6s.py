"""Physical model of the 6-module ZBOT in its *snake* configuration (``ZBOT_D_6S_CFG``).

Mirrors the parameters of ``/root/reference/source/zbot/zbot/assets/zbot_cfg.py:109-168``
(init pose (0,0,0.05) / rot (0.707,0,-0.707,0), implicit PD kp 20 / kd 0.5 / effort 20) and of the USD it
names, ``zbot_6s_v03.usd``.  That file is a binary crate; its structure is described in SURVEY.md A.6 and is
the one of the ASCII sibling ``zbot_assets/zbot_6s_v04.usda`` (links a1,b1,...,a6,b6; revolute
``localRot0 = (0.92388,0,0.38268,0)``, ``localRot1 = identity``; fixed ``localPos0 = (-0.037477,0,0.037477)``;
link xforms at ``zbot_6s_v04.usda:117-119,196-198,263-265,340-342,...``), which is what is decoded here.

It is the SAME physical chain as the walking robot (same joint positions and axes in the chain frame): only
the link frames differ -- "b" links are rotated +45 deg about y, every second module by 180 deg about z --
and the "b" / even-"a" mass properties are therefore rotated with their frames.  The reduced model keeps
the walking robot's body frames (origin at the joint, chain orientation) and carries the constant link
rotations separately (``link_rot``) for the quantities the MDP reads per link.
"""
from __future__ import annotations

import math

import numpy as np

from . import zbot_6s as Z

LINK_NAMES = ("a1", "b1", "a2", "b2", "a3", "b3", "a4", "b4", "a5", "b5", "a6", "b6")
JOINT_NAMES = Z.JOINT_NAMES

KP = 20.0
KD = 0.5
EFFORT_LIMIT = 20.0
DEFAULT_ROOT_POS = (0.0, 0.0, 0.05)
DEFAULT_ROOT_QUAT_RAW = (0.707, 0.0, -0.707, 0.0)   # zbot_cfg.py:140 (normalised below)
DEFAULT_JOINT_POS = (0.0,) * 6

SPHERE_RADIUS = 0.05
#: filtered self-contact sensors (``…/zbot6_direct/zbot_direct_6dof_snake_v0.py:23-48``): (sensor link, filter links)
SELF_CONTACT_SENSORS = (
    ("a1", ("b4", "a5", "b5", "a6", "b6")),
    ("b6", ("a3", "b2", "a2", "b1")),
    ("b1", ("a5", "b5", "a6")),
    ("a6", ("b2", "a2")),
)

_RY45 = np.array([math.cos(math.pi / 8), 0.0, math.sin(math.pi / 8), 0.0])
_RZ180 = np.array([0.0, 0.0, 0.0, 1.0])
_ID = np.array([1.0, 0.0, 0.0, 0.0])


def _link_rot(name: str) -> np.ndarray:
    k = int(name[1])
    even = (k % 2 == 0)
    if name[0] == "a":
        return _RZ180 if even else _ID
    return Z.quat_mul(_RZ180, _RY45) if even else _RY45


def build_model(dtype=np.float32) -> Z.ZbotModel:
    Ia = Z._link_inertia(Z.A_DIAG_INERTIA, Z.A_PRINCIPAL_AXES)
    Ib = Z._link_inertia(Z.B_DIAG_INERTIA, Z.B_PRINCIPAL_AXES)
    ca, cb = np.array(Z.A_COM), np.array(Z.B_COM)
    off = np.array([0.0, 0.0, Z.LINK_SPACING])

    link_body = np.array([0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6])
    link_rot = np.stack([_link_rot(n) for n in LINK_NAMES])
    link_offset = np.zeros((12, 3))
    link_com = np.zeros((12, 3))
    link_inertia_body = []
    for i, n in enumerate(LINK_NAMES):
        second = (n[0] == "a" and i > 0)          # a_{k+1} sits on top of b_k inside a merged body
        link_offset[i] = off if second else 0.0
        link_com[i] = ca if n[0] == "a" else cb
        R = Z.quat_to_mat(link_rot[i])
        I = R @ (Ia if n[0] == "a" else Ib) @ R.T
        link_inertia_body.append((Z.LINK_MASS, link_offset[i] + R @ link_com[i], I))

    mass, com, inertia = np.zeros(7), np.zeros((7, 3)), np.zeros((7, 3, 3))
    mass[0], com[0], inertia[0] = link_inertia_body[0]
    mass[6], com[6], inertia[6] = link_inertia_body[11]
    for k in range(1, 6):
        (m1, c1, I1), (m2, c2, I2) = link_inertia_body[2 * k - 1], link_inertia_body[2 * k]
        mass[k], com[k], inertia[k] = Z._merge(m1, c1, I1, m2, c2, I2)

    jpos = np.zeros((6, 3))
    jpos[0] = off
    jpos[1:] = 2.0 * off
    jaxis = np.array([[s * Z.SIN45, 0.0, Z.SIN45] for s in Z.JOINT_AXIS_SIGN])

    # ground contact: a line of r = 0.05 spheres on the chain axis, one per 0.053 m (joint centres, link
    # junctions and both ends) -- the cylinders lie on their sides in this task
    contact = [[0, 0.0, 0.0, 0.0, SPHERE_RADIUS]]
    for b in range(1, 6):
        contact.append([b, 0.0, 0.0, 0.0, SPHERE_RADIUS])
        contact.append([b, 0.0, 0.0, Z.LINK_SPACING, SPHERE_RADIUS])
    contact.append([6, 0.0, 0.0, 0.0, SPHERE_RADIUS])
    contact.append([6, 0.0, 0.0, Z.LINK_SPACING, SPHERE_RADIUS])

    # self-contact spheres: one r = 0.05 sphere at the middle of each half-module cylinder (body frame)
    link_centre = np.zeros((12, 3))
    for i, n in enumerate(LINK_NAMES):
        if n[0] == "a":
            link_centre[i] = link_offset[i] + np.array([0.0, 0.0, 0.044])     # a: z in [0, 0.088]
        else:
            link_centre[i] = np.array([0.0, 0.0, 0.009])                       # b: z in [-0.035, 0.053]
    pairs = tuple((LINK_NAMES.index(s), LINK_NAMES.index(f)) for s, fs in SELF_CONTACT_SENSORS for f in fs)

    def r(x):
        return np.asarray(x, dtype=dtype).astype(np.float64)

    inertia = 0.5 * (inertia + inertia.transpose(0, 2, 1))
    inertia[:, 0, 1] = inertia[:, 1, 0] = 0.0      # stay zero under Ry / Rz(180) up to CAD noise
    inertia[:, 1, 2] = inertia[:, 2, 1] = 0.0
    com[:, 1] = 0.0
    link_com[:, 1] = 0.0
    q0 = np.array(DEFAULT_ROOT_QUAT_RAW)
    q0 = q0 / np.linalg.norm(q0)
    return Z.ZbotModel(
        name="zbot_6s_v03", body_mass=r(mass), body_com=r(com), body_inertia=r(inertia),
        joint_pos=r(jpos), joint_axis=jaxis, foot_points=r(np.zeros((2, 4, 3))), sphere_centre=r(np.zeros((5, 3))),
        sphere_radius=float(dtype(SPHERE_RADIUS)), link_offset=r(link_offset), link_com=r(link_com),
        link_body=link_body, default_joint_pos=r(DEFAULT_JOINT_POS), default_root_pos=r(DEFAULT_ROOT_POS),
        default_root_quat=q0, link_rot=link_rot, contact_list=r(np.array(contact)),
        link_names=LINK_NAMES, sensor_body_names=LINK_NAMES, self_pairs=pairs, link_centre=r(link_centre),
        kp=KP, kd=KD, effort_limit=EFFORT_LIMIT, gravity=float(dtype(Z.GRAVITY)),
    )


_MODEL = None


def model_f32() -> Z.ZbotModel:
    global _MODEL
    if _MODEL is None:
        _MODEL = build_model(np.float32)
    return _MODEL
