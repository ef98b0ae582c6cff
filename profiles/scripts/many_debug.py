"""Debug helper: `optimize-many --jobs J` on a small synthetic genome; prints the failing groups' log lines."""
import os, subprocess, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_host import _write_genome
from allhic_b200 import farm
d = pathlib.Path(tempfile.mkdtemp())
clmf, refiles = _write_genome(d, [60, 45, 80, 30])
(d / "list").write_text("".join("%s %s\n" % (r, clmf) for r in refiles))
for jobs in sys.argv[1:] or ["4"]:
    r = subprocess.run([farm.BINARY, "optimize-many", str(d / "list"), "--ngen", "300", "--jobs", jobs], capture_output=True, text=True, timeout=600)
    print("jobs", jobs, "rc", r.returncode, "success", r.stderr.count("Success"))
    print("\n".join(l for l in r.stderr.splitlines() if "CRITIC" in l or "rror" in l)[:3000])
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("AHX_")))
