"""Where the wall time of the group farm goes (config4, every visible GPU): process start, CUDA contexts, groups.
  python profiles/farm_timing.py [n_groups] [jobs]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from allhic_b200 import farm, synth, _ffi
from concurrent.futures import ProcessPoolExecutor
n_groups = int(sys.argv[1]) if len(sys.argv) > 1 else 48
jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
outdir = os.path.join(os.environ.get("TMPDIR", "/tmp"), "allhic_b200_config4")
os.makedirs(outdir, exist_ok=True)
def _make(c):
    ids, clm, _ = synth.materialize(c["n"], seed=c["seed"], outdir=outdir)
    return ids, clm
if __name__ == "__main__":
    with ProcessPoolExecutor(max_workers=16) as ex:
        groups = list(ex.map(_make, synth.config4_groups()[:n_groups]))
    ndev = _ffi.lib().ahx_device_count()
    with tempfile.NamedTemporaryFile("w", suffix=".list", delete=False) as f:
        f.write("".join("%s %s\n" % g for g in groups))
    env = dict(os.environ, AHX_FARM_TIMING="1")
    t0 = time.perf_counter()
    p = subprocess.run([farm.BINARY, "optimize-many", f.name, "--gpus", str(ndev), "--jobs", str(jobs)], capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    lines = [l for l in p.stderr.splitlines() if l.startswith("farm:")]
    ctx = sorted(float(l.split(" at ")[1].split()[0]) for l in lines if "has its context" in l)
    grp = [(float(l.split(" from ")[1].split()[0]), float(l.split(" to ")[1].split()[0])) for l in lines if " group " in l]
    print("wall %.2f s, rc %d; contexts ready: first %.2f median %.2f last %.2f s; groups: first start %.2f, last end %.2f, longest %.2f s, median %.2f s"
          % (wall, p.returncode, ctx[0], ctx[len(ctx) // 2], ctx[-1], min(a for a, b in grp), max(b for a, b in grp), max(b - a for a, b in grp), sorted(b - a for a, b in grp)[len(grp) // 2]))
