"""Kernel tuning aid: build variants of libisscabac.so with extra -D flags (here, on the CPU box),
then time each of them on the GPU box with the bench's device-resident C3 workload.

  python tools/tune.py build  name1:DEF1,DEF2  name2:DEF3 ...     (here)
  python tools/tune.py run [--bins N] [--steps K]                  (under gpurun)

`run` benches every lib*.so under isscabac_b200/_variants plus the default library and prints one line
per variant: encode / decode Gbins/s and kernel milliseconds.  Every variant goes through bench.py's own
checks (round trip, finish flags, byte identity of the first streams against the reference engine)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    if sys.argv[1] == "build":
        from isscabac_b200 import build as B
        for old in glob.glob(os.path.join(ROOT, "isscabac_b200", "_variants", "lib*.so")):
            os.remove(old)
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition(":")
            print(B.build_variant(name, [d for d in defs.split(",") if d]))
        return
    extra = sys.argv[2:]
    libs = [None] + sorted(glob.glob(os.path.join(ROOT, "isscabac_b200", "_variants", "lib*.so")))
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env["ISSCABAC_LIB"] = lib
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu", "--no-configs", "--steps", "10"] + extra,
                           capture_output=True, text=True, env=env)
        name = os.path.basename(lib) if lib else "default"
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(json.dumps({"variant": name, "enc": round(d["encode_gbins"], 1), "dec": round(d["decode_gbins"], 1),
                              "ms": {k: round(v, 3) for k, v in d["kernel_ms"].items()}}), flush=True)
        except Exception:
            print(json.dumps({"variant": name, "error": (r.stderr or r.stdout)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
