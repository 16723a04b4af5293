"""URDF front-end (the URDFParser stand-in): numbering, fixed-joint merging, conventions."""
import numpy as np
import pytest

from gridcodegenerator_b200 import load_named_robot, parse_urdf_string
from gridcodegenerator_b200.robot import rot_axis, rpy_to_R, spatial_inertia, xform


def test_named_topologies():
    assert load_named_robot("iiwa14").parent == [-1, 0, 1, 2, 3, 4, 5]
    assert load_named_robot("hyq").parent == [-1, 0, 1, -1, 3, 4, -1, 6, 7, -1, 9, 10]
    atlas = load_named_robot("atlas")
    assert atlas.n == 30 and atlas.get_max_bfs_level() + 1 == 10 and atlas.get_total_ancestor_count() == 120
    c = load_named_robot("chain64")
    assert c.is_serial_chain() and c.are_Ss_identical(list(range(64))) and c.get_total_subtree_count() == 2080
    m = load_named_robot("mixed5")
    assert m.parent == [-1, 0, 1, 1, 3] and m.S_ind == [2, 4, 0, 3, 1] and m.damping[0] == 0.4


def test_robot_api_surface_matches_what_the_reference_calls():
    r = load_named_robot("hyq")
    assert r.get_num_pos() == 12 and r.get_parent_id(4) == 3 and r.get_parent_id_array()[0] == -1
    assert r.get_ids_by_bfs_level(0) == [0, 3, 6, 9] and r.get_max_bfs_width() == 4
    assert r.get_ancestors_by_id(2) == [0, 1] and r.get_subtree_by_id(3) == [3, 4, 5]
    a = r.get_ancestors_by_id(2)
    a.append(99)                                   # the oracle mutates the list it gets (_test.py:355-356)
    assert r.get_ancestors_by_id(2) == [0, 1]
    assert r.has_repeated_parents([0, 3]) and not r.has_repeated_parents([1, 4])
    assert r.get_unique_parent_ids([1, 2, 4]) == [0, 1, 3]
    assert r.get_is_ancestor_of(0, 2) and r.get_is_in_subtree_of(2, 0) and not r.get_is_in_subtree_of(3, 0)
    assert list(r.get_S_by_id(1)) == [0, 1, 0, 0, 0, 0]
    assert len(r.get_Imats_ordered_by_id()) == 13 and r.get_Imats_dict_by_id()[5] is r.get_Imat_by_id(5)
    assert r.get_joint_by_id(0).get_name() == "LF_HAA" and r.get_link_by_id(2).get_name() == "LF_lowerleg"
    X = r.get_Xmat_Func_by_id(1)(0.3)
    assert X.shape == (6, 6) and np.allclose(X[:3, 3:], 0) and np.allclose(X[:3, :3], X[3:, 3:])


def test_sympy_xmats_agree_with_numeric_ones():
    sp = pytest.importorskip("sympy")
    r = load_named_robot("mixed5")
    for i, Xs in enumerate(r.get_Xmats_ordered_by_id()):
        f = sp.lambdify(sp.Symbol("theta"), Xs, "numpy")
        assert np.allclose(np.array(f(0.37), dtype=float), r.Xmat(i, 0.37), atol=1e-12)


URDF = """<robot name="t">
  <link name="base"/>
  <link name="a"><inertial><origin xyz="0.1 0 0" rpy="0 0 0"/><mass value="2"/>
    <inertia ixx="0.1" ixy="0" ixz="0" iyy="0.2" iyz="0" izz="0.3"/></inertial></link>
  <link name="tool"><inertial><origin xyz="0 0.05 0" rpy="0.3 0 0"/><mass value="0.5"/>
    <inertia ixx="0.01" ixy="0.001" ixz="0" iyy="0.02" iyz="0" izz="0.03"/></inertial></link>
  <link name="b"><inertial><origin xyz="0 0 0.1"/><mass value="1"/>
    <inertia ixx="0.01" ixy="0" ixz="0" iyy="0.01" iyz="0" izz="0.01"/></inertial></link>
  <joint name="j0" type="revolute"><parent link="base"/><child link="a"/><origin xyz="0 0 0.5" rpy="0.1 0.2 0.3"/>
    <axis xyz="0 1 0"/><dynamics damping="0.7"/></joint>
  <joint name="fix" type="fixed"><parent link="a"/><child link="tool"/><origin xyz="0.2 0 0" rpy="0 0.4 0"/></joint>
  <joint name="j1" type="prismatic"><parent link="tool"/><child link="b"/><origin xyz="0 0.1 0" rpy="0 0 0.5"/>
    <axis xyz="0 0 1"/></joint>
</robot>"""


def test_fixed_joint_merging_and_conventions():
    r = parse_urdf_string(URDF)
    assert r.parent == [-1, 0] and r.S_ind == [1, 5] and r.damping == [0.7, 0.0]
    # X_tree of j0: E0 = R(rpy)^T, r0 = xyz
    assert np.allclose(r.E0[0], rpy_to_R([0.1, 0.2, 0.3]).T) and np.allclose(r.r0[0], [0, 0, 0.5])
    # the tool is merged into link a through the fixed transform
    X_fix = xform(rpy_to_R([0, 0.4, 0]).T, [0.2, 0, 0])
    Rc = rpy_to_R([0.3, 0, 0])
    I_tool = spatial_inertia(0.5, [0, 0.05, 0], Rc @ np.array([[0.01, 0.001, 0], [0.001, 0.02, 0], [0, 0, 0.03]]) @ Rc.T)
    I_a = spatial_inertia(2.0, [0.1, 0, 0], np.diag([0.1, 0.2, 0.3]))
    assert np.allclose(r.Imats[0], I_a + X_fix.T @ I_tool @ X_fix)
    # j1 hangs off the fixed link: its tree transform is composed through the fixed joint
    X_T1 = xform(rpy_to_R([0, 0, 0.5]).T, [0, 0.1, 0]) @ X_fix
    assert np.allclose(r.Xmat(1, 0.0), X_T1)
    # prismatic: E constant, translation along the child z axis
    assert np.allclose(r.Xmat(1, 0.25), xform(np.eye(3), [0, 0, 0.25]) @ X_T1)
    # revolute about y: X = blkdiag(ry, ry) X_tree
    ry = rot_axis(1, 0.6)
    XJ = np.zeros((6, 6)); XJ[:3, :3] = ry; XJ[3:, 3:] = ry
    assert np.allclose(r.Xmat(0, 0.6), XJ @ xform(r.E0[0], r.r0[0]))


def _urdf_fk(text, q):
    """World pose (R, p) of every moving joint's URDF child frame by plain URDF semantics:
    T_child = T_parent * T_origin * Rot(axis, q)  or  * Trans(axis * q) - independent of Robot/xform."""
    import xml.etree.ElementTree as ET
    root = ET.fromstring(text)
    joints = {j.find("child").get("link"): j for j in root.findall("joint")}
    order = [j for j in root.findall("joint")]
    pose = {}
    children = {l.get("name") for l in root.findall("link")} - set(joints)
    base = children.pop()
    pose[base] = (np.eye(3), np.zeros(3))
    out, qi = [], 0
    pending = list(order)
    while pending:
        for j in list(pending):
            par = j.find("parent").get("link")
            if par not in pose:
                continue
            pending.remove(j)
            Rp, pp = pose[par]
            o = j.find("origin")
            xyz = np.array([float(t) for t in (o.get("xyz") or "0 0 0").split()]) if o is not None else np.zeros(3)
            rpy = [float(t) for t in (o.get("rpy") or "0 0 0").split()] if o is not None else [0, 0, 0]
            R, p = Rp @ rpy_to_R(rpy), pp + Rp @ xyz
            if j.get("type") != "fixed":
                a = np.array([float(t) for t in j.find("axis").get("xyz").split()])
                a = a / np.linalg.norm(a)
                if j.get("type") == "prismatic":
                    p = p + R @ (a * q[qi])
                else:
                    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                    R = R @ (np.eye(3) + np.sin(q[qi]) * K + (1 - np.cos(q[qi])) * K @ K)
                out.append((qi, R, p))
                qi += 1
            pose[j.find("child").get("link")] = (R, p)
    return out


SKEW = URDF.replace('<axis xyz="0 1 0"/>', '<axis xyz="0 -1 0"/>').replace('<axis xyz="0 0 1"/>', '<axis xyz="0.2 -0.3 0.9"/>')


@pytest.mark.parametrize("text", [URDF, SKEW, URDF.replace('<axis xyz="0 0 1"/>', '<axis xyz="-1 0 0"/>')])
def test_negative_and_skew_axes_keep_the_urdf_kinematics(text):
    """Joints about negative / non-principal axes get a one-hot S by re-defining the child frame
    (gridcodegenerator_b200/urdf.py): the frame origins and the joint axes in the world must be exactly what
    plain URDF forward kinematics gives, for the URDF's own sign of q."""
    from gridcodegenerator_b200.urdf import axis_alignment
    r = parse_urdf_string(text)
    assert all(0 <= k < 6 for k in r.S_ind) and r.n == 2
    q = np.array([0.7, -0.4])
    fk = _urdf_fk(text, q)
    X = np.eye(6)
    for i, R_w, p_w in fk:                            # serial chain j0 -> j1
        X = r.Xmat(i, q[i]) @ X                       # base -> F'_i
        E = X[:3, :3]
        rx = -E.T @ X[3:, :3]
        assert np.allclose([rx[2, 1], rx[0, 2], rx[1, 0]], p_w, atol=1e-12)       # frame origin in the world
        k = r.S_ind[i] % 3
        import xml.etree.ElementTree as ET
        axes = [np.array([float(t) for t in j.find("axis").get("xyz").split()]) for j in ET.fromstring(text).findall("joint")
                if j.get("type") != "fixed"]
        a_w = R_w @ (axes[i] / np.linalg.norm(axes[i]))
        assert np.allclose(E.T[:, k], a_w, atol=1e-12)                             # e_k of F' is the URDF axis
        kk, R_a = axis_alignment(axes[i])
        assert kk == k and np.allclose(R_a.T @ E, R_w.T, atol=1e-12)               # F' = R_a F
        assert np.isclose(np.linalg.det(R_a), 1.0)


def test_flipped_axis_is_the_mirrored_joint():
    """axis -e_y with q is the same mechanism as axis +e_y with -q: equal inverse dynamics up to the sign of
    that joint's torque (checked through the oracle-independent traced RNEA)."""
    from gridcodegenerator_b200.algorithms import TRACERS
    a = parse_urdf_string(URDF.replace('<axis xyz="0 1 0"/>', '<axis xyz="0 -1 0"/>'))
    b = parse_urdf_string(URDF)
    rng = np.random.default_rng(0)
    q, qd, qdd = rng.uniform(-1, 1, (3, 2))
    sgn = np.array([-1.0, 1.0])

    def rnea(robot, q, qd, qdd):
        ins = {"gravity": np.array([9.81])}
        for i in range(2):
            ins["q%d" % i], ins["qd%d" % i], ins["qdd%d" % i] = (np.array([x[i]]) for x in (q, qd, qdd))
        return TRACERS["id_qdd"](robot).evaluate(ins)["c"][0]

    ca, cb = rnea(a.with_damping(0.0), q, qd, qdd), rnea(b.with_damping(0.0), sgn * q, sgn * qd, sgn * qdd)
    assert np.allclose(ca, sgn * cb, atol=1e-12)


def test_unsupported_urdfs_fail_loudly():
    with pytest.raises(ValueError):
        parse_urdf_string(URDF.replace('<axis xyz="0 1 0"/>', '<axis xyz="0 0 0"/>'))
    with pytest.raises(NotImplementedError):
        parse_urdf_string(URDF.replace('type="prismatic"', 'type="floating"'))
    with pytest.raises(ValueError):
        parse_urdf_string(URDF.replace('<link name="base"/>', '<link name="base"/><link name="stray"/>'))
