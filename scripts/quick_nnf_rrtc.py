"""Ad-hoc scale runs of the NNF and batched-RRTConnect paths (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm()
what = sys.argv[1]
if what == "nnf":
    W, M, k, D = int(sys.argv[2]), int(sys.argv[3]), 5, int(sys.argv[4])
    iters = int(sys.argv[5])
    nobs = int(os.environ.get("NOBS", "32")); tseed = int(os.environ.get("TSEED", "4"))
    sph, box = synth.obstacles(3, n_spheres=nobs, n_boxes=nobs)
    gw = P.World(dh, link, xyzr, sph, box)
    R = (W - 1) * (D + 1) + 1
    q, poses = synth.nnf_reference(tseed, 2 * R + 5, dh)
    print('traj valid frac', gw.is_valid_batch(q).mean())
    ids = synth.lin_int_space(0, len(q) - 1, R)
    rng = np.random.default_rng(1)
    counts = np.zeros(R, dtype=np.uint32); counts[0] = counts[-1] = M
    counts[1:-1] += np.bincount(rng.integers(1, R - 1, W * M - 2 * M), minlength=R)[1:-1].astype(np.uint32)
    ik = synth.nnf_ik_states(2, q[ids], counts)
    t = time.time(); g = P.Nnf(gw, poses[ids], counts, ik, num_nn=k, discretization=D); torch.cuda.synchronize(); t1 = time.time() - t
    print(f"nnf W={W} M={M} D={D}: R={R} V={g.n_vertices} E={g.n_edges} TPG nodes={R*g.n_vertices:.3e} setup {t1:.3f}s", flush=True)
    if len(sys.argv) > 6: g.set_mode(int(sys.argv[6]))
    t = time.time(); r = g.solve(max_lazy_iters=iters); t2 = time.time() - t
    print(f"solve: iters={r['lazy_iters']} solved={r['solved']} err={r['final_error']:.4f} goal={r['goal_cost']:.4g} eval={r['n_evaluated']} coll={r['n_collisions']} {t2:.3f}s  last dp {g.last_dp_ms():.3f} ms band={g.band_info()}", flush=True)
else:
    B = int(sys.argv[2]); iters = int(sys.argv[3])
    sph, box = synth.obstacles(3)
    gw = P.World(dh, link, xyzr, sph, box)
    q = synth.cloud(4, 8 * B)
    v = gw.is_valid_batch(q)
    qv = q[v == 1]
    starts, goals = qv[0:2 * B:2].copy(), qv[1:2 * B:2].copy()
    for ext in ("ext-con", "con-con"):
        rb = P.RrtcBatch(gw, B, -np.pi, np.pi, extension_type=ext, seed=4, max_iters=iters, max_nodes=2048)
        rb.run(starts, goals)
        t = time.time(); rb.run(starts, goals); dt = time.time() - t
        s = rb.summary()
        print(f"rrtc B={B} {ext}: {dt*1e3:.1f} ms  {B/dt:.1f} plans/s solved={np.mean(s['status']==6):.3f} "
              f"mean iters={s['iterations'].mean():.1f} mean nodes={(s['n_start']+s['n_goal']).mean():.1f} max nodes={max(s['n_start'].max(), s['n_goal'].max())}", flush=True)
