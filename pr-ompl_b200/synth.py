"""Synthetic inputs of the BASELINE.json configs (arm, obstacles, clouds, edges, queries).

The reference takes the robot and the environment as opaque user callbacks
(StateValidityChecker::isValid at pr_ompl/src/RRTConnect.cpp:198, FK/IK callbacks at
frechet/include/NNFrechet.hpp:43-59); BASELINE.json names a "synthetic 7-DOF arm" with
"sphere-approximated forward kinematics" against "sphere/box obstacles".  This module only
GENERATES those inputs as plain numpy arrays (SURVEY.md §8(d) "synthetic inputs"); it
computes nothing on the hot path.
"""
import numpy as np

PI = float(np.pi)
DOF = 7


def default_arm():
    """7-DOF arm, standard DH rows (a, d, cos(alpha), sin(alpha), theta0) and 16 link spheres.

    Returns (dh[7,5] f64, sphere_link[16] i32, sphere_xyzr[16,4] f64).
    """
    d = [0.36, 0.0, 0.42, 0.0, 0.40, 0.0, 0.126]
    sa = [-1.0, 1.0, 1.0, -1.0, -1.0, 1.0, 0.0]
    ca = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    dh = np.zeros((DOF, 5), dtype=np.float64)
    for i in range(DOF):
        dh[i] = (0.0, d[i], ca[i], sa[i], 0.0)
    spheres = [
        (0, 0.0, 0.0, 0.08, 0.10), (0, 0.0, 0.0, 0.22, 0.09),
        (1, 0.0, 0.0, 0.0, 0.09),
        (2, 0.0, 0.0, 0.10, 0.075), (2, 0.0, 0.0, 0.21, 0.075), (2, 0.0, 0.0, 0.32, 0.075),
        (3, 0.0, 0.0, 0.0, 0.08),
        (4, 0.0, 0.0, 0.10, 0.065), (4, 0.0, 0.0, 0.20, 0.065), (4, 0.0, 0.0, 0.30, 0.065),
        (5, 0.0, 0.0, 0.0, 0.07),
        (6, 0.0, 0.0, 0.06, 0.06),
        (7, 0.0, 0.0, 0.0, 0.05), (7, 0.0, 0.0, 0.06, 0.045), (7, 0.0, 0.0, 0.12, 0.04),
        (7, 0.0, 0.0, 0.17, 0.03),
    ]
    link = np.array([s[0] for s in spheres], dtype=np.int32)
    xyzr = np.array([s[1:] for s in spheres], dtype=np.float64)
    return dh, link, xyzr


def obstacles(seed=3, n_spheres=128, n_boxes=128, shell=(0.55, 1.25), size=(0.03, 0.10),
              keepout_radius=0.32):
    """Sphere and axis-aligned box obstacles, uniform in a shell around the shoulder (0,0,0.36).

    Obstacles whose bounding sphere would touch the fixed base column (|xy| < keepout_radius
    for z < 0.6) are re-drawn, otherwise no configuration could be valid.
    Returns (obs_spheres[n,4] f64 (x,y,z,R), obs_boxes[n,6] f64 (cx,cy,cz,hx,hy,hz)).
    """
    rng = np.random.default_rng(seed)
    centre = np.array([0.0, 0.0, 0.36])

    def draw_centre(extent):
        while True:
            v = rng.normal(size=3)
            v /= np.linalg.norm(v)
            # uniform in volume between the two radii
            u = rng.uniform(shell[0] ** 3, shell[1] ** 3) ** (1.0 / 3.0)
            c = centre + u * v
            if c[2] < 0.0:
                continue
            if np.hypot(c[0], c[1]) - extent < keepout_radius and c[2] - extent < 0.6:
                continue
            return c

    sph = np.zeros((n_spheres, 4))
    for i in range(n_spheres):
        r = rng.uniform(*size)
        sph[i, :3] = draw_centre(r)
        sph[i, 3] = r
    box = np.zeros((n_boxes, 6))
    for i in range(n_boxes):
        h = rng.uniform(size[0], size[1], size=3)
        box[i, :3] = draw_centre(float(np.linalg.norm(h)))
        box[i, 3:] = h
    return sph, box


def f32_representable(x):
    """Round to float32 and back (clouds are fp32-representable doubles)."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def cloud(seed, n, dim=DOF, low=-PI, high=PI, dup_fraction=0.0):
    """n points i.i.d. U[low,high]^dim (fp32-representable), optionally with exact duplicates."""
    rng = np.random.default_rng(seed)
    x = f32_representable(rng.uniform(low, high, size=(n, dim)))
    if dup_fraction > 0.0 and n > 1:
        m = int(n * dup_fraction)
        src = rng.integers(0, n, size=m)
        dst = rng.integers(0, n, size=m)
        x[dst] = x[src]
    return np.ascontiguousarray(x)


def edges(seed, n, dim=DOF, low=-PI, high=PI, max_len=1.0):
    """Edge batch of the checkMotion config: a ~ U[low,high]^dim, b = a + L*u, u uniform on the
    unit sphere, L ~ U(0, max_len]."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(low, high, size=(n, dim))
    u = rng.normal(size=(n, dim))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    L = max_len * (1.0 - rng.uniform(0.0, 1.0, size=(n, 1)))
    b = a + L * u
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def fk_pose(dh, q):
    """End-effector pose (3x4) of the DH arm at configuration q (numpy; input generation only)."""
    T = np.eye(4)
    for i in range(dh.shape[0]):
        a, d, ca, sa, th0 = dh[i]
        ct, st = np.cos(q[i] + th0), np.sin(q[i] + th0)
        L = np.array([[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st], [0.0, sa, ca, d],
                      [0.0, 0.0, 0.0, 1.0]])
        T = T @ L
    return T[:3, :]


def lin_int_space(mn, mx, n):
    """frechet/src/util.cpp:3-21 -- (int)(min + i*delta); may stop short of max."""
    delta = (mx - mn) / float(n - 1)
    return np.array([int(mn + i * delta) for i in range(n)], dtype=np.int64)


def nnf_layer_counts(seed, R, W, M):
    """IK budget per super-sampled reference index, as NNFrechet draws it (frechet/src/NNFrechet.cpp:310-343,
    418-448): both end waypoints get M, the other W*M - 2M samples fall on indices 1..R-2 by
    std::uniform_int_distribution<int> on std::default_random_engine.  Restated for libstdc++ (minstd_rand0:
    x <- 16807 x mod 2^31-1; the distribution scales the generator's range down with rejection); input
    generation only -- pinned against the reference's own draws in tests/test_oracle_nnf.py."""
    x = seed % 2147483647 or 1
    counts = np.zeros(R, dtype=np.uint32)
    counts[0] = counts[-1] = M
    a, b = 1, R - 2
    urngrange = 2147483646 - 1
    uerange = (b - a) + 1
    scaling = urngrange // uerange
    past = uerange * scaling
    for _ in range(W * M - 2 * M):
        while True:
            x = (x * 16807) % 2147483647
            ret = x - 1
            if ret < past:
                break
        counts[a + ret // scaling] += 1
    return counts


def nnf_reference(seed, n_poses, dh, dof=DOF, amplitude=0.6):
    """A smooth joint trajectory (n_poses x dof) and the end-effector poses (n_poses x 12) it traces:
    the reference path handed to NNFrechet::setRefPath."""
    rng = np.random.default_rng(seed)
    q0 = rng.uniform(-1.0, 1.0, size=dof)
    amp = rng.uniform(0.3, 1.0, size=dof) * amplitude
    freq = rng.uniform(0.5, 1.5, size=dof)
    phase = rng.uniform(0, 2 * PI, size=dof)
    t = np.linspace(0.0, 1.0, n_poses)[:, None]
    q = q0 + amp * np.sin(2 * PI * freq * t + phase)
    poses = np.stack([fk_pose(dh, qi).reshape(12) for qi in q])
    return q, poses


def nnf_ik_states(seed, ref_configs, layer_counts, sigma=0.15):
    """Synthetic IK: the trajectory configuration of the waypoint plus N(0, sigma^2) joint noise, in
    NN-graph vertex order (layer 0, layer R-1, layers 1..R-2).  Legal because the planner
    recomputes every pose with FK (frechet/src/NNFrechet.cpp:295)."""
    rng = np.random.default_rng(seed)
    R = len(layer_counts)
    order = [0, R - 1] + list(range(1, R - 1)) if R > 1 else [0]
    rows = []
    for r in order:
        c = int(layer_counts[r])
        if c:
            rows.append(ref_configs[r] + sigma * rng.normal(size=(c, ref_configs.shape[1])))
    return np.ascontiguousarray(np.vstack(rows)) if rows else np.zeros((0, ref_configs.shape[1]))
