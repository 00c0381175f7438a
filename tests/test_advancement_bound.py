"""The motion bound behind conservative advancement (csrc/check_motion.cu, csrc/capi.cu): between neighbouring
tested states of an edge no link-sphere centre moves by more than  travel = sum_i |to_i - from_i| * rho_i / nd,
rho_i = (chain length from joint i's frame to the sphere's frame) + |local offset|, maximised over the spheres
the joint carries.  Checked against the oracle's fp64 forward kinematics on random edges (CPU only)."""
import numpy as np


def rho_of(dh, link, xyzr):
    dof = dh.shape[0]
    seg = np.sqrt(dh[:, 0] ** 2 + dh[:, 1] ** 2)          # |translation| of DH row m
    loc = np.sqrt((xyzr[:, :3] ** 2).sum(1))
    rho = np.zeros(dof)
    for i in range(dof):
        for j, l in enumerate(link):
            if l >= i + 1:
                rho[i] = max(rho[i], seg[i:l].sum() + loc[j])
    return rho


def test_travel_bounds_sphere_motion(O, synth, world_arrays):
    dh, link, xyzr, sph, box = world_arrays
    w = O.World(dh, link, xyzr, sph, box)
    rho = rho_of(dh, link, xyzr)
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(300):
        a = rng.uniform(-np.pi, np.pi, 7)
        b = a + rng.normal(size=7) * rng.uniform(0.01, 2.0)
        nd = int(np.ceil(np.linalg.norm(b - a) / 0.01))
        travel = (np.abs(b - a) * rho).sum() / nd
        j = int(rng.integers(0, nd))
        c0 = w.fk(a + (b - a) * (j / nd))[1]
        c1 = w.fk(a + (b - a) * ((j + 1) / nd))[1]
        moved = np.sqrt(((c1 - c0) ** 2).sum(1)).max()
        assert moved <= travel * (1 + 1e-9) + 1e-15
        worst = max(worst, moved / travel)
    assert 0.05 < worst <= 1.0                            # the bound is not vacuous either


def test_single_joint_motion_reaches_a_good_part_of_rho(O, synth, world_arrays):
    # turning one joint alone: displacement of the farthest sphere <= rho_i * angle, and rho_i is within a small
    # factor of what some configuration actually achieves (a loose bound would waste the advancement)
    dh, link, xyzr, sph, box = world_arrays
    w = O.World(dh, link, xyzr, sph, box)
    rho = rho_of(dh, link, xyzr)
    rng = np.random.default_rng(4)
    tight = 0
    for i in range(7):
        best = 0.0
        for _ in range(200):
            q = rng.uniform(-np.pi, np.pi, 7)
            q2 = q.copy()
            q2[i] += 1e-3
            moved = np.sqrt(((w.fk(q2)[1] - w.fk(q)[1]) ** 2).sum(1)).max()
            assert moved <= rho[i] * 1e-3 * (1 + 1e-6)
            best = max(best, moved / 1e-3)
        tight += best >= 0.3 * rho[i]
    assert tight >= 5       # spheres on a joint's own axis (the wrist roll) never reach its rho
