"""Minimal URDF front-end: URDF text -> :class:`Robot`.

Replaces the external URDFParser package (reference README.md:8, not vendored)
for the subset the reference supports: a fixed base, 1-DoF revolute /
continuous / prismatic joints about a positive principal axis, fixed joints
merged into their parent link, joint ids assigned in DFS pre-order
(SURVEY.md section 8c "Implicit invariants").
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
            hot = [k for k in range(3) if abs(axis[k] - 1.0) < 1e-9]
            if len(hot) != 1 or abs(np.abs(axis).sum() - 1.0) > 1e-9:
                raise NotImplementedError(
                    "joint %s: only positive principal axes are supported (S must be one-hot, "
                    "reference helpers/_topology_helpers.py:247), got %r" % (j.get("name"), axis.tolist()))
            dyn = j.find("dynamics")
            jid = len(parent)
            parent.append(owner)
            S_ind.append(hot[0] + (3 if jtype == "prismatic" else 0))
            # recover (E, r) of the composed tree transform
            E = X_T[:3, :3]
            rx = -E.T @ X_T[3:, :3]
            E0.append(E.copy())
            r0.append(np.array([rx[2, 1], rx[0, 2], rx[1, 0]]))
            Imats.append(np.zeros((6, 6)))
            damping.append(float(dyn.get("damping", "0")) if dyn is not None else 0.0)
            jnames.append(j.get("name"))
            lnames.append(child)
            visit(child, jid, np.eye(6))

    visit(roots[0], -1, np.eye(6))
    return Robot(name or root.get("name", "robot"), parent, S_ind, E0, r0, Imats, damping,
                 jnames, lnames, base_I)


def load_urdf(path: str, name: Optional[str] = None) -> Robot:
    with open(path, "r") as f:
        return parse_urdf_string(f.read(), name)


NAMED_ROBOTS = ("iiwa14", "hyq", "atlas", "chain64", "mixed5")


def load_named_robot(name: str) -> Robot:
    """Synthetic URDFs of the named topologies (no real URDFs exist offline,
    SURVEY.md section 0); generated by robots/make_synthetic_urdfs.py."""
    path = os.path.join(_ROBOT_DIR, name + ".urdf")
    if not os.path.exists(path):
        raise FileNotFoundError("no URDF for robot %r under %s" % (name, _ROBOT_DIR))
    return load_urdf(path, name)
