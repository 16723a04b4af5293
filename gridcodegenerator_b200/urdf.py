"""Minimal URDF front-end: URDF text -> :class:`Robot`.

Replaces the external URDFParser package (reference README.md:8, not vendored)
for the subset the reference supports: a fixed base, 1-DoF revolute /
continuous / prismatic joints, fixed joints merged into their parent link, joint
ids assigned in DFS pre-order (SURVEY.md section 8c "Implicit invariants").

The reference needs a ONE-HOT motion subspace S (helpers/_topology_helpers.py:247 takes
``S.tolist().index(1)``), i.e. a joint about a positive principal axis of its own frame.
Real URDFs also use negative axes (``0 -1 0``: half of the Atlas arm joints) and, rarely,
skew ones.  Such a joint keeps its meaning and gets a one-hot S by re-defining the child
frame: with R_a the rotation that takes the URDF axis a onto the nearest positive principal
axis e_k, the joint frame F is replaced by F' = R_a F (tree transform E0' = R_a E0, same
origin), and everything that hangs off the child link - its inertia, fixed children, the
origins of the next joints - is expressed in F' through the constant transform F' -> F.
q keeps the URDF's sign convention.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

from .robot import Robot, rpy_to_R, spatial_inertia, xform

_ROBOT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "robots")


def _floats(s: Optional[str], default):
    if s is None:
        return np.array(default, dtype=np.float64)
    return np.array([float(t) for t in s.split()], dtype=np.float64)


def _origin(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.zeros(3), np.zeros(3)
    return _floats(o.get("xyz"), [0, 0, 0]), _floats(o.get("rpy"), [0, 0, 0])


def _link_inertia(link) -> np.ndarray:
    inertial = link.find("inertial")
    if inertial is None:
        return np.zeros((6, 6))
    xyz, rpy = _origin(inertial)
    mass = float(inertial.find("mass").get("value"))
    it = inertial.find("inertia")
    g = lambda k: float(it.get(k, "0"))
    Ic = np.array([[g("ixx"), g("ixy"), g("ixz")],
                   [g("ixy"), g("iyy"), g("iyz")],
                   [g("ixz"), g("iyz"), g("izz")]])
    Rc = rpy_to_R(rpy)
    return spatial_inertia(mass, xyz, Rc @ Ic @ Rc.T)


def axis_alignment(axis) -> tuple:
    """(k, R_a): principal axis index k and the rotation with R_a @ a = e_k for the unit vector a = axis/|axis|.
    k is the component of largest magnitude; a negative one is first flipped by a half turn about the
    next principal axis (exact for the common ``0 -1 0`` case), the rest is the minimal rotation."""
    a = np.asarray(axis, dtype=np.float64)
    nrm = np.linalg.norm(a)
    if nrm < 1e-12:
        raise ValueError("zero joint axis")
    a = a / nrm
    k = int(np.argmax(np.abs(a)))
    F = np.eye(3)
    if a[k] < 0:
        m = (k + 1) % 3
        F = -np.eye(3)
        F[m, m] = 1.0                      # half turn about e_m: det = +1, e_k -> -e_k
    a1 = F @ a
    t = np.zeros(3)
    t[k] = 1.0
    v, c = np.cross(a1, t), float(a1 @ t)
    vx = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    R2 = np.eye(3) + vx + vx @ vx / (1.0 + c)          # c = |a_k| >= 1/sqrt(3)
    R = R2 @ F
    R[np.abs(R) < 1e-15] = 0.0
    return k, R


def parse_urdf_string(text: str, name: Optional[str] = None) -> Robot:
    root = ET.fromstring(text)
    links: Dict[str, ET.Element] = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children: Dict[str, List[ET.Element]] = {k: [] for k in links}
    child_links = set()
    for j in joints:
        children[j.find("parent").get("link")].append(j)
        child_links.add(j.find("child").get("link"))
    roots = [k for k in links if k not in child_links]
    if len(roots) != 1:
        raise ValueError("URDF must have exactly one root link, found %r" % roots)

    parent: List[int] = []
    S_ind: List[int] = []
    E0: List[np.ndarray] = []
    r0: List[np.ndarray] = []
    Imats: List[np.ndarray] = []
    damping: List[float] = []
    jnames: List[str] = []
    lnames: List[str] = []
    base_I = np.zeros((6, 6))

    def visit(link_name: str, owner: int, X_owner_to_link: np.ndarray):
        """owner = id of the moving joint whose body this link belongs to (-1 = base);
        X_owner_to_link = motion transform from the owner's frame to this link's frame."""
        nonlocal base_I
        I_here = _link_inertia(links[link_name])
        I_in_owner = X_owner_to_link.T @ I_here @ X_owner_to_link
        if owner == -1:
            base_I = base_I + I_in_owner
        else:
            Imats[owner] = Imats[owner] + I_in_owner
        for j in children[link_name]:
            xyz, rpy = _origin(j)
            R = rpy_to_R(rpy)
            X_T = xform(R.T, xyz) @ X_owner_to_link       # owner frame -> joint (child) frame
            jtype = j.get("type")
            child = j.find("child").get("link")
            if jtype == "fixed":
                visit(child, owner, X_T)
                continue
            if jtype not in ("revolute", "continuous", "prismatic"):
                raise NotImplementedError("joint type %r is not supported" % jtype)
            ax = j.find("axis")
            axis = _floats(ax.get("xyz") if ax is not None else None, [1, 0, 0])
            k_axis, R_a = axis_alignment(axis)              # identity for a positive principal axis
            X_T = xform(R_a, np.zeros(3)) @ X_T             # joint frame F -> F' = R_a F (same origin)
            dyn = j.find("dynamics")
            jid = len(parent)
            parent.append(owner)
            S_ind.append(k_axis + (3 if jtype == "prismatic" else 0))
            # recover (E, r) of the composed tree transform
            E = X_T[:3, :3]
            rx = -E.T @ X_T[3:, :3]
            E0.append(E.copy())
            r0.append(np.array([rx[2, 1], rx[0, 2], rx[1, 0]]))
            Imats.append(np.zeros((6, 6)))
            damping.append(float(dyn.get("damping", "0")) if dyn is not None else 0.0)
            jnames.append(j.get("name"))
            lnames.append(child)
            visit(child, jid, xform(R_a.T, np.zeros(3)))    # the child link lives in F: F' -> F

    visit(roots[0], -1, np.eye(6))
    return Robot(name or root.get("name", "robot"), parent, S_ind, E0, r0, Imats, damping,
                 jnames, lnames, base_I)


def load_urdf(path: str, name: Optional[str] = None) -> Robot:
    with open(path, "r") as f:
        return parse_urdf_string(f.read(), name)


NAMED_ROBOTS = ("iiwa14", "hyq", "atlas", "chain64", "mixed5", "pchain4")


def load_named_robot(name: str) -> Robot:
    """Synthetic URDFs of the named topologies (no real URDFs exist offline,
    SURVEY.md section 0); generated by robots/make_synthetic_urdfs.py."""
    path = os.path.join(_ROBOT_DIR, name + ".urdf")
    if not os.path.exists(path):
        raise FileNotFoundError("no URDF for robot %r under %s" % (name, _ROBOT_DIR))
    return load_urdf(path, name)
