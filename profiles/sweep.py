"""Kernel tuning sweep (run on the GPU box): times bench.py for alternative builds / kernels / block sizes.
usage: python profiles/sweep.py name=ENV1=V1,ENV2=V2[:extra bench args] ..."""
import json, os, subprocess, sys
base = ["python", "bench.py", "--steps", "20", "--warmup", "5", "--no-e2e", "--no-cpu-baseline"]
for spec in sys.argv[1:]:
    name, rest = spec.split("=", 1)
    envs, _, extra = rest.partition(":")
    env = dict(os.environ)
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run(base + extra.split(), env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(f"{name:28s} {d['ms_per_step']:8.3f} ms  {d['value']/1e9:7.2f} G/s  frac {d['roofline']['frac']:.3f}", flush=True)
    except Exception as e:
        print(name, "FAILED", r.stderr[-300:], flush=True)
