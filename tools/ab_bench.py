"""Same-box A/B of bench configurations (developer tool): python tools/ab_bench.py "ENV=V ENV2=V" "ENV=W" ...
Runs `bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras` once per configuration, twice round-robin."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfgs = sys.argv[1:] or [""]
res = {c: [] for c in cfgs}
for rep in range(2):
    for c in cfgs:
        env = dict(os.environ)
        for kv in c.split():
            k, v = kv.split("=")
            env[k] = v
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline",
                              "--no-extras"], env=env, capture_output=True, text=True).stdout
        try:
            d = json.loads(out.strip().splitlines()[-1])
            res[c].append((d["ms_per_step"], d["value"], d["e2e"]["value"]))
        except Exception as e:
            res[c].append(("fail", str(e)[:80], out[-200:]))
for c, v in res.items():
    print(f"{c or '(default)':50s}", "  ".join(f"{a[0]:.1f} ms {a[1]:.1f} / {a[2]:.1f}" if a[0] != "fail" else str(a) for a in v), flush=True)
