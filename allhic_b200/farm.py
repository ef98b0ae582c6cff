"""One linkage group per GPU, no collective: the reference's own suggestion for
polyploid partitions ("groups are independent and may be run as k separate
processes", allhic.go:170-173; `pipeline` runs them sequentially,
allhic.go:285-293).  Each group is a full `allhic-b200 optimize` run (the C++
host over the C ABI) pinned to one GPU; groups are handed out
longest-processing-time-first (weight ~ N^2) to whichever GPU is free.
"""
import os
import re
import subprocess
import threading
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "bin", "allhic-b200")


def _count_contigs(refile):
    with open(refile) as f:
        return sum(1 for line in f if line.strip() and not line.startswith("#"))


def optimize_groups(groups, devices, flags=(), binary=BINARY, log_dir=None):
    """groups: list of (refile, clmfile); devices: CUDA device indices to farm over.
    Returns a list of dicts (group, device, seconds, final_score, n_contigs, ok) in input order."""
    order = sorted(range(len(groups)), key=lambda g: -_count_contigs(groups[g][0]) ** 2)
    lock = threading.Lock()
    results = [None] * len(groups)

    def worker(dev):
        while True:
            with lock:
                if not order:
                    return
                g = order.pop(0)
            refile, clmfile = groups[g]
            t0 = time.perf_counter()
            # the process sees its GPU only: CUDA start-up enumerates every visible device, which costs
            # seconds per process on an 8-GPU box
            env = dict(os.environ)
            vis = [x for x in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
            env["CUDA_VISIBLE_DEVICES"] = vis[dev] if dev < len(vis) else str(dev)
            p = subprocess.run([binary, "optimize", refile, clmfile, "--device", "0", *flags],
                               capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            scores = re.findall(r"max_score=([0-9.eE+-]+)", p.stdout)
            if log_dir:
                with open(os.path.join(log_dir, "group%02d.stderr" % g), "w") as f:
                    f.write(p.stderr)
            results[g] = dict(group=g, device=dev, seconds=dt, n_contigs=_count_contigs(refile),
                              final_score=float(scores[-1]) if scores else None,
                              ok=p.returncode == 0 and "Success" in p.stderr)

    threads = [threading.Thread(target=worker, args=(d,)) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results
