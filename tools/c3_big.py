"""Config 3 at the BASELINE shape through bench.run_config3 (8192 envs x 256^2 by default)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
h = bench.Harness(1, "cuda:0")
print(json.dumps(bench.run_config3(h, "cuda:0", n, steps, 2)))
