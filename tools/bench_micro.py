"""Quick A/B: milliseconds per 64-episode micro-batch (graph replay) of the bench workload; prints one line."""
import json, subprocess, sys, os
env = dict(os.environ)
for kv in sys.argv[1:]:
    k, v = kv.split("=", 1)
    env[k] = v
out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-extras", "--no-cpu-baseline",
                      "--global-episodes", "1024", "--e2e-steps", "1"], capture_output=True, text=True, env=env).stdout
d = json.loads(out.strip().splitlines()[-1])
r = d["roofline"]
print(" ".join(sys.argv[1:]) or "default", "| ms/micro %.3f  eps/s %.0f  tensor-kernels %.3f ms frac %.3f  tuple %.3f ms" %
      (d["config"]["ms_per_micro_batch"], d["value"], r["kernel_ms_per_micro_batch"], r["frac"],
       d["roofline_tuple"]["kernel_ms_per_micro_batch"]))
