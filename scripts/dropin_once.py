"""Drop-in A alone: the reference's unmodified RRTConnect over the GPU structures, one query after the other
(bench.py's extra.dropin_a)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
sys.argv = [sys.argv[0]]
args = bench.parse()
env = bench.Env(args, 0, 0, 1)
print(json.dumps(bench.measure_dropin_a(env, args, True)))
