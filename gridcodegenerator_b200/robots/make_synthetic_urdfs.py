#!/usr/bin/env python3
"""Writes the synthetic URDFs of the named topologies (iiwa14, hyq, atlas, chain64).

No real URDFs exist offline, so these are "synthetic URDFs of the
named topology" as BASELINE.json allows.  iiwa14 uses the commonly published
iiwa_description numbers (recalled, unverified); the others use
fixed plausible masses/geometry.  All joints use positive principal axes.
Run:  python gridcodegenerator_b200/robots/make_synthetic_urdfs.py
"""
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PI = math.pi


def _fmt(v):
    return " ".join(repr(float(x)) for x in v)


def link_xml(name, mass, com, inertia_diag, inertia_off=(0.0, 0.0, 0.0)):
    ixx, iyy, izz = inertia_diag
    ixy, ixz, iyz = inertia_off
    return (
        '  <link name="%s">\n    <inertial>\n      <origin xyz="%s" rpy="0 0 0"/>\n'
        '      <mass value="%r"/>\n'
        '      <inertia ixx="%r" ixy="%r" ixz="%r" iyy="%r" iyz="%r" izz="%r"/>\n'
        "    </inertial>\n  </link>\n"
        % (name, _fmt(com), float(mass), ixx, ixy, ixz, iyy, iyz, izz))


def joint_xml(name, jtype, parent, child, xyz, rpy, axis, damping=0.0):
    s = '  <joint name="%s" type="%s">\n    <parent link="%s"/>\n    <child link="%s"/>\n' % (
        name, jtype, parent, child)
    s += '    <origin xyz="%s" rpy="%s"/>\n' % (_fmt(xyz), _fmt(rpy))
    if jtype != "fixed":
        s += '    <axis xyz="%s"/>\n' % _fmt(axis)
        s += '    <limit lower="-3.14" upper="3.14" effort="300" velocity="10"/>\n'
        s += '    <dynamics damping="%r"/>\n' % float(damping)
    s += "  </joint>\n"
    return s


def write(name, body):
    with open(os.path.join(HERE, name + ".urdf"), "w") as f:
        f.write('<?xml version="1.0"?>\n<robot name="%s">\n%s</robot>\n' % (name, body))


AX = {"x": (1, 0, 0), "y": (0, 1, 0), "z": (0, 0, 1)}


def iiwa14():
    origins = [((0, 0, 0.1575), (0, 0, 0)), ((0, 0, 0.2025), (PI / 2, 0, PI)),
               ((0, 0.2045, 0), (PI / 2, 0, PI)), ((0, 0, 0.2155), (PI / 2, 0, 0)),
               ((0, 0.1845, 0), (-PI / 2, PI, 0)), ((0, 0, 0.2155), (PI / 2, 0, 0)),
               ((0, 0.081, 0), (-PI / 2, PI, 0))]
    links = [(5.76, (0, -0.03, 0.12), (0.033, 0.0333, 0.0123)),
             (6.35, (0.0003, 0.059, 0.042), (0.0305, 0.0304, 0.011)),
             (3.5, (0, 0.03, 0.13), (0.025, 0.0238, 0.0076)),
             (3.5, (0, 0.067, 0.034), (0.017, 0.0164, 0.006)),
             (3.5, (0.0001, 0.021, 0.076), (0.01, 0.0087, 0.00449)),
             (1.8, (0, 0.0006, 0.0004), (0.0049, 0.0047, 0.0036)),
             (1.2, (0, 0, 0.02), (0.001, 0.001, 0.001))]
    body = link_xml("iiwa_link_0", 5.0, (-0.1, 0, 0.07), (0.05, 0.06, 0.03))
    for i, ((xyz, rpy), (m, c, I)) in enumerate(zip(origins, links), start=1):
        body += link_xml("iiwa_link_%d" % i, m, c, I)
        body += joint_xml("iiwa_joint_%d" % i, "revolute", "iiwa_link_%d" % (i - 1), "iiwa_link_%d" % i,
                          xyz, rpy, AX["z"], damping=0.0)
    # a fixed end-effector flange exercises fixed-joint merging
    body += link_xml("iiwa_link_ee", 0.3, (0, 0, 0.02), (0.0002, 0.0002, 0.0003))
    body += joint_xml("iiwa_joint_ee", "fixed", "iiwa_link_7", "iiwa_link_ee", (0, 0, 0.045), (0, 0, 0), None)
    write("iiwa14", body)


def hyq():
    body = link_xml("trunk", 53.433, (0.056, 0.0215, 0.00358), (1.5725, 8.5015, 9.1954), (0.0397, 0.6366, 0.0275))
    legs = {"LF": (0.3735, 0.207, 0.0), "RF": (0.3735, -0.207, 0.0),
            "LH": (-0.3735, 0.207, 0.0), "RH": (-0.3735, -0.207, 0.0)}
    for leg, hip in legs.items():
        sy = 1.0 if leg[0] == "L" else -1.0
        body += link_xml(leg + "_hipassembly", 2.93, (0.04263, 0.0, 0.16931 * 0.2), (0.05071, 0.13466, 0.08875),
                         (-3.6e-4 * sy, 0.02262 * 0.1, -5.1e-4))
        body += joint_xml(leg + "_HAA", "revolute", "trunk", leg + "_hipassembly", hip,
                          (0.0, PI / 2 if leg[1] == "F" else -PI / 2, 0.0), AX["x"])
        body += link_xml(leg + "_upperleg", 2.638, (0.15074, -0.02625 * sy, 0.0), (0.00368, 0.02719, 0.02811),
                         (2.19e-3 * sy, -1.0e-4, 3.5e-5))
        body += joint_xml(leg + "_HFE", "revolute", leg + "_hipassembly", leg + "_upperleg",
                          (0.08, 0.0, 0.0), (sy * PI / 2, 0.0, 0.0), AX["y"])
        body += link_xml(leg + "_lowerleg", 0.881, (0.1254, 0.0005 * sy, -0.0001), (0.00047, 0.01256, 0.01233),
                         (3.0e-5 * sy, -1.0e-5, 0.0))
        body += joint_xml(leg + "_KFE", "revolute", leg + "_upperleg", leg + "_lowerleg",
                          (0.35, 0.0, 0.0), (0.0, 0.0, 0.0), AX["y"])
        body += link_xml(leg + "_foot", 0.05, (0.0, 0.0, 0.0), (1e-5, 1e-5, 1e-5))
        body += joint_xml(leg + "_foot_joint", "fixed", leg + "_lowerleg", leg + "_foot",
                          (0.33, 0.0, 0.0), (0.0, 0.0, 0.0), None)
    write("hyq", body)


def atlas():
    """Atlas-like 30 DoF: back(3) -> {l_arm 7, neck 1, r_arm 7}, l_leg 6, r_leg 6."""
    body = link_xml("pelvis", 17.9, (0.011, 0.0, 0.027), (0.125, 0.095, 0.117), (0.0008, 0.0007, -0.0005))

    def chain(prefix, parent_link, specs):
        nonlocal body
        prev = parent_link
        for k, (axis, xyz, rpy, mass, com, inertia) in enumerate(specs):
            link = "%s_%d" % (prefix, k)
            body += link_xml(link, mass, com, inertia, (1e-4 * (k + 1), -2e-4, 3e-4 / (k + 1)))
            body += joint_xml("%s_j%d" % (prefix, k), "revolute", prev, link, xyz, rpy, AX[axis])
            prev = link
        return prev

    utorso = chain("back", "pelvis", [
        ("z", (-0.0125, 0, 0), (0, 0, 0), 2.27, (-0.011, 0, 0.075), (0.0039, 0.0034, 0.0017)),
        ("y", (0, 0, 0.162), (0, 0, 0), 0.8, (-0.007, 0, 0.012), (0.0004, 0.0007, 0.0008)),
        ("x", (0, 0, 0.05), (0, 0, 0), 84.4, (-0.062, 0.002, 0.306), (1.58, 1.60, 0.72)),
    ])

    def arm(side):
        s = 1.0 if side == "l" else -1.0
        return [
            ("z", (0.1406, 0.2256 * s, 0.4776), (0, 0, 0), 4.47, (0, -0.003 * s, 0.091), (0.0087, 0.0024, 0.0087)),
            ("x", (0, 0.11 * s, -0.245), (0, 0, 0), 3.9, (0, -0.019 * s, 0.0), (0.0041, 0.0110, 0.0088)),
            ("y", (0, 0.187 * s, 0.016), (0, 0, 0), 4.42, (0, -0.048 * s, 0.084), (0.0063, 0.0043, 0.0040)),
            ("x", (0, 0.119 * s, 0.0092), (0, 0, 0), 3.39, (0, -0.027 * s, 0.0), (0.0028, 0.0046, 0.0041)),
            ("y", (0, 0.2955 * s, 0), (0, 0, 0), 2.51, (0, 0.005 * s, 0.0), (0.0011, 0.0023, 0.0021)),
            ("x", (0, 0.0, 0), (0, 0, PI / 2 * s), 0.71, (0, 0.0, 0.0), (0.0005, 0.0004, 0.0006)),
            ("y", (0, 0.12 * s, 0), (0, 0, 0), 2.26, (0, 0.09 * s, 0.0), (0.0035, 0.0015, 0.0031)),
        ]

    chain("l_arm", utorso, arm("l"))
    chain("neck", utorso, [("y", (0.2546, 0, 0.6215), (0, 0, 0), 1.42, (-0.075, 0, 0.03), (0.0040, 0.0042, 0.0036))])
    chain("r_arm", utorso, arm("r"))

    def leg(side):
        s = 1.0 if side == "l" else -1.0
        return [
            ("z", (0, 0.089 * s, 0), (0, 0, 0), 2.41, (0.0, 0, 0.0), (0.0012, 0.0016, 0.0016)),
            ("x", (0, 0, 0), (0, 0, 0), 0.69, (0.0, 0, 0.0), (0.0008, 0.0009, 0.0011)),
            ("y", (0.05, 0.0225 * s, -0.066), (0, 0, 0), 8.2, (0, 0.0, -0.21), (0.09, 0.09, 0.02)),
            ("y", (-0.05, 0, -0.374), (0, 0, 0), 4.5, (0.001, 0, -0.187), (0.077, 0.076, 0.01)),
            ("y", (0, 0, -0.422), (0, 0, 0), 0.125, (0, 0, 0), (1e-4, 1e-4, 1e-4)),
            ("x", (0, 0, 0), (0, 0, 0), 2.41, (0.027, 0, -0.067), (0.002, 0.007, 0.008)),
        ]

    chain("l_leg", "pelvis", leg("l"))
    chain("r_leg", "pelvis", leg("r"))
    write("atlas", body)


def chain64():
    body = link_xml("base", 1.0, (0, 0, 0), (0.01, 0.01, 0.01))
    prev = "base"
    for i in range(64):
        link = "link_%d" % i
        body += link_xml(link, 0.5, (0.0, 0.01, 0.03), (0.0008, 0.0008, 0.0003), (1e-5, 0.0, -2e-5))
        rpy = (PI / 2, 0, 0) if i % 2 else (-PI / 2, 0, 0)
        body += joint_xml("joint_%d" % i, "revolute", prev, link, (0.0, 0.0, 0.06) if i % 2 else (0.0, -0.06, 0.0),
                          rpy, AX["z"])
        prev = link
    write("chain64", body)


def mixed5():
    """Small branched robot that exercises what the named robots do not: prismatic joints, oblique
    X_tree rotations (dense E0), non-zero damping, products of inertia, a fixed link in the middle."""
    body = link_xml("base", 2.0, (0, 0, 0), (0.02, 0.02, 0.02))
    specs = [  # name, type, parent link, axis, xyz, rpy, damping, mass, com, inertia diag, inertia off
        ("j0", "revolute", "base", "z", (0.0, 0.0, 0.1), (0.3, -0.2, 0.5), 0.4, 3.0, (0.01, 0.02, 0.08), (0.03, 0.025, 0.01), (1e-3, -2e-3, 5e-4)),
        ("j1", "prismatic", "l_j0", "y", (0.05, 0.0, 0.2), (-0.4, 0.1, 0.9), 0.1, 1.5, (0.0, 0.05, 0.0), (0.01, 0.004, 0.01), (2e-4, 1e-4, -3e-4)),
        ("j2", "revolute", "l_j1", "x", (0.0, 0.15, 0.0), (0.7, 0.6, -0.3), 0.0, 1.0, (0.03, 0.0, 0.02), (0.004, 0.006, 0.005), (0.0, 2e-4, 1e-4)),
        ("j3", "prismatic", "l_j1_plate", "x", (0.1, -0.05, 0.02), (0.2, 0.0, 1.2), 0.25, 0.8, (0.04, 0.01, 0.0), (0.002, 0.003, 0.003), (1e-4, 0.0, 0.0)),
        ("j4", "revolute", "l_j3", "y", (0.12, 0.0, 0.0), (-0.9, 0.3, 0.1), 0.05, 0.6, (0.0, 0.0, 0.05), (0.002, 0.002, 0.0008), (0.0, 0.0, 1e-4)),
    ]
    for name, jtype, parent, axis, xyz, rpy, damp, m, com, Id, Io in specs:
        if name == "j3":      # a fixed plate between j1's link and j3 exercises fixed-joint merging mid-tree
            body += link_xml("l_j1_plate", 0.4, (0.0, 0.01, 0.0), (0.001, 0.001, 0.0015))
            body += joint_xml("j1_plate_fix", "fixed", "l_j1", "l_j1_plate", (0.02, 0.03, -0.01), (0.1, 0.2, 0.3), None)
        body += link_xml("l_" + name, m, com, Id, Io)
        body += joint_xml(name, jtype, parent, "l_" + name, xyz, rpy, AX[axis], damping=damp)
    write("mixed5", body)


if __name__ == "__main__":
    mixed5()
    iiwa14()
    hyq()
    atlas()
    chain64()
    print("wrote mixed5 / iiwa14 / hyq / atlas / chain64 URDFs into", HERE)
