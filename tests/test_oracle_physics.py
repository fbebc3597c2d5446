"""Known-answer tests pinning the CPU physics oracle (oracle/rigid_oracle.c).

The reference ships no vectors for this stage (its engine is an absent third-party dependency: "parity unpinned",
DESIGN.md), so the oracle is pinned against physics instead: free fall, conservation of linear / angular momentum
and energy in free flight (checks CRBA + RNE + integration against an independent numpy kinematics), static
equilibrium (contact force = m g), joint limits, and fp32-vs-fp64 agreement.
"""
import numpy as np
import pytest

from _kin_check import momentum_energy
from hcr_genesis_lr_cl_b200.robot_model import GO2_DOF_NAMES, TRON1_PF_DOF_NAMES, load_robot_model
from oracle.physics import P_GRAV, PhysicsOracle, default_params

Q0 = np.array([0, 0.8, -1.5] * 4, float)


def _state(n, z=10.0):
    st = np.zeros((n, 13))
    st[:, 2] = z
    st[:, 3] = 1
    return st


def _env(n, nj, arm=0.1):
    envp = np.zeros((n, 5))
    envp[:, 4] = 1
    jp = np.zeros((n, 3 * nj))
    jp[:, :nj] = arm
    return envp, jp


def test_free_fall_matches_closed_form():
    m = load_robot_model("go2", GO2_DOF_NAMES)
    o = PhysicsOracle(m, default_params(dt=0.005))
    st, q, qd = _state(1), Q0[None].copy(), np.zeros((1, 12))
    envp, jp = _env(1, 12)
    # hold the pose with a stiff PD so the body falls as one rigid object
    for _ in range(100):
        tau = 200 * (Q0 - q) - 5 * qd
        o.substep(st, q, qd, tau, envp, jp)
    t = 100 * 0.005
    assert abs(st[0, 9] + 9.81 * t) < 1e-6                 # semi-implicit Euler: v = -g t exactly
    assert abs(st[0, 2] - (10 - 0.5 * 9.81 * t * (t + 0.005))) < 1e-5
    assert np.abs(st[0, 10:13]).max() < 1e-6 and np.abs(st[0, 7:9]).max() < 1e-6


@pytest.mark.parametrize("robot,names", [("go2", GO2_DOF_NAMES), ("tron1_pf", TRON1_PF_DOF_NAMES)])
def test_momentum_and_energy_conservation_in_free_flight(robot, names):
    m = load_robot_model(robot, names, armature=0.0)
    nj = len(names)
    prm = default_params(dt=0.0005)
    prm[P_GRAV] = 0.0
    o = PhysicsOracle(m, prm)
    rng = np.random.default_rng(0)
    st = _state(1)
    qq = rng.normal(size=4)
    st[0, 3:7] = qq / np.linalg.norm(qq)
    st[0, 7:13] = rng.normal(size=6)
    q = rng.uniform(-0.3, 0.3, (1, nj)) + (Q0[None] if robot == "go2" else 0)
    qd = rng.normal(size=(1, nj)) * 3
    envp, jp = _env(1, nj, arm=0.0)
    P0, L0, E0 = momentum_energy(m, st[0], q[0], qd[0])
    for _ in range(200):
        o.substep(st, q, qd, np.zeros((1, nj)), envp, jp)
    P1, L1, E1 = momentum_energy(m, st[0], q[0], qd[0])
    assert np.abs(P1 - P0).max() < 1e-3 * np.abs(P0).max()
    assert np.abs(L1 - L0).max() < 1e-3 * np.abs(L0).max()
    assert abs(E1 - E0) < 1e-3 * E0


def test_standing_contact_force_equals_weight():
    m = load_robot_model("go2", GO2_DOF_NAMES)
    o = PhysicsOracle(m, default_params(dt=0.005, iters=50))
    st, q, qd = _state(1, z=0.35), Q0[None].copy(), np.zeros((1, 12))
    envp, jp = _env(1, 12)
    for _ in range(1200):
        tau = 60 * (Q0 - q) - 2.0 * qd
        lf, nc = o.substep(st, q, qd, tau, envp, jp)
    assert nc[0] >= 4
    assert abs(lf[0, :, 2].sum() - m.total_mass * 9.81) < 0.01 * m.total_mass * 9.81
    assert np.abs(st[0, 7:13]).max() < 1e-2
    assert np.abs(lf[0, :, :2].sum(0)).max() < 1.0               # no net horizontal force at rest
    assert 0.2 < st[0, 2] < 0.4


def test_joint_limit_rows_hold_the_joint():
    m = load_robot_model("go2", GO2_DOF_NAMES)
    o = PhysicsOracle(m, default_params(dt=0.005))
    st, q, qd = _state(1), Q0[None].copy(), np.zeros((1, 12))
    envp, jp = _env(1, 12)
    hi = m.dof_limits[0, 1]
    for _ in range(300):
        tau = np.zeros((1, 12))
        tau[0, 0] = 20.0                                           # push the FR hip into its upper limit
        tau[0, 1:] = (200 * (Q0 - q) - 5 * qd)[0, 1:]
        o.substep(st, q, qd, tau, envp, jp)
    assert hi - 1e-3 < q[0, 0] < hi + 0.05


def test_fp32_oracle_tracks_fp64():
    m = load_robot_model("go2", GO2_DOF_NAMES)
    rng = np.random.default_rng(1)
    out = {}
    for prec in ("f64", "f32"):
        o = PhysicsOracle(m, default_params(dt=0.005), precision=prec)
        st, q, qd = _state(4, z=0.3), np.tile(Q0, (4, 1)), np.zeros((4, 12))
        envp, jp = _env(4, 12)
        r = np.random.default_rng(5)
        for _ in range(8):
            tau = 20 * (Q0 + r.uniform(-0.2, 0.2, (4, 12)) - q) - 0.5 * qd
            o.substep(st, q, qd, tau, envp, jp)
        out[prec] = (st.copy(), q.copy(), qd.copy())
    assert np.abs(out["f32"][0][:, :7] - out["f64"][0][:, :7]).max() < 1e-5
    assert np.abs(out["f32"][1] - out["f64"][1]).max() < 1e-5
    assert np.abs(out["f32"][2] - out["f64"][2]).max() < 5e-3


def test_heightfield_contact_normal_on_a_slope():
    """Robot resting on a 30 % heightfield slope (triangle normals + friction cone): static equilibrium."""
    m = load_robot_model("go2", GO2_DOF_NAMES)
    rows = cols = 64
    hf = (np.arange(rows)[:, None] * np.ones((1, cols)) * 6).astype(np.int16)      # dh/dx = 6*0.005/0.1 = 0.3
    prm = default_params(dt=0.005, hscale=0.1, vscale=0.005, border=0.0)
    o = PhysicsOracle(m, prm, hf)
    st, q, qd = _state(1, z=0.0), Q0[None].copy(), np.zeros((1, 12))
    st[0, 0:2] = 3.0
    st[0, 2] = 0.3 * 3.0 + 0.27
    envp, jp = _env(1, 12)
    tot, contacts = np.zeros(3), 0
    for i in range(1200):
        tau = 60 * (Q0 - q) - 2 * qd
        lf, nc = o.substep(st, q, qd, tau, envp, jp)
        if i >= 900:
            tot += lf[0].sum(0)
            contacts += int(nc[0])
    assert contacts > 0
    # mu = 1 > slope 0.3: friction holds the robot, so the mean contact force balances gravity (vertical, = m g)
    assert abs(tot[2] / 300 - m.total_mass * 9.81) < 0.1 * m.total_mass * 9.81
    assert abs(tot[0]) < 0.1 * tot[2] and abs(tot[1]) < 0.1 * tot[2]
    assert np.abs(st[0, 7:9]).max() < 0.05


def test_trimesh_surface_matches_the_reference_triangles():
    """terrain.mesh_type "trimesh": the oracle's collision surface == the triangles the reference's own
    convert_heightfield_to_trimesh builds (terrain_utils.py:835-902; golden from tools/make_golden_trimesh.py, which imports
    it unchanged).  The heightfield mode cuts each cell along the other diagonal, so it must NOT match on twisted cells."""
    import os
    from oracle.physics import terrain_query
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trimesh_surface.npz"))
    prm = default_params(hscale=float(g["hscale"]), vscale=float(g["vscale"]), border=0.0, trimesh=True)
    h, n = terrain_query(prm, g["hf"], g["xy"])
    assert np.abs(h - g["height"]).max() < 1e-6 and np.abs(n - g["normal"]).max() < 1e-6      # the reference stores float32 vertices
    prm_hf = default_params(hscale=float(g["hscale"]), vscale=float(g["vscale"]), border=0.0, trimesh=False)
    h2, _ = terrain_query(prm_hf, g["hf"], g["xy"])
    assert np.abs(h2 - g["height"]).max() > 1e-3
    # the same query restated from the committed triangles (no reference needed at test time)
    v, t = g["vertices"].astype(np.float64), g["triangles"]
    for k in range(0, len(g["xy"]), 37):
        p = g["xy"][k]
        best, zb = -1.0, None
        for a, b, c in v[t]:
            T = np.array([[b[0] - a[0], c[0] - a[0]], [b[1] - a[1], c[1] - a[1]]])
            l1, l2 = np.linalg.solve(T, p - a[:2])
            m = min(l1, l2, 1 - l1 - l2)
            if m > best:
                best, zb = m, a[2] * (1 - l1 - l2) + b[2] * l1 + c[2] * l2
        assert abs(zb - h[k]) < 1e-6          # vertices are stored as float32 in the fixture


@pytest.mark.parametrize("robot,names", [("go2", GO2_DOF_NAMES), ("tron1_pf", TRON1_PF_DOF_NAMES)])
def test_work_energy_balance_with_joint_torques_and_gravity(robot, names):
    """dE/dt = tau . qd for the floating-base tree in flight (E = kinetic + potential from the independent numpy
    kinematics): pins M^-1, the RNE bias incl. gravity and the applied-torque path of the oracle together."""
    m = load_robot_model(robot, names, armature=0.0)
    nj = len(names)
    dt = 0.0002
    o = PhysicsOracle(m, default_params(dt=dt))
    rng = np.random.default_rng(5)
    st = _state(1, z=3.0)
    qq = rng.normal(size=4)
    st[0, 3:7] = qq / np.linalg.norm(qq)
    st[0, 7:13] = rng.normal(size=6)
    q = rng.uniform(-0.1, 0.1, (1, nj)) + (Q0[None] if robot == "go2" else 0)
    qd = rng.normal(size=(1, nj))
    envp, jp = _env(1, nj, arm=0.0)
    tau = rng.uniform(-1, 1, (1, nj)) * 0.13 * m.effort[None, :]   # small / short enough that no joint reaches a limit in the window
    E0 = momentum_energy(m, st[0], q[0], qd[0], gravity=9.81)[2]
    work = 0.0
    lim = m.dof_limits
    for _ in range(100):
        qd_before = qd[0].copy()
        o.substep(st, q, qd, tau, envp, jp)
        work += float(tau[0] @ (0.5 * (qd_before + qd[0]))) * dt       # trapezoid of the joint power
    E1 = momentum_energy(m, st[0], q[0], qd[0], gravity=9.81)[2]
    assert (q[0] > lim[:, 0]).all() and (q[0] < lim[:, 1]).all()
    assert abs(work) > 0.5                                             # the torques did inject a visible amount of energy
    assert abs((E1 - E0) - work) < 0.02 * abs(work) + 0.02, (E1 - E0, work)
