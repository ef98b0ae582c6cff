"""One linkage group per GPU, no collective: the reference's own suggestion for
polyploid partitions ("groups are independent and may be run as k separate
processes", allhic.go:170-173; `pipeline` runs them sequentially,
allhic.go:285-293).  Each group is a full `allhic-b200 optimize` run (the C++
host over the C ABI) pinned to one GPU; groups are handed out
longest-processing-time-first (weight ~ N^2) to whichever GPU is free.

`jobs=J` (default 6) runs ONE `allhic-b200 optimize-many --gpus G --jobs J`
process for the whole box: a worker thread per (GPU, job slot), each with its
own context and stream.  At the CLI's default population (100) a group's
persistent GA kernel occupies 25 of a B200's 148 SMs, so ~6 groups run side by
side on one GPU; CUDA starts once instead of once per GPU process (eight
processes starting together take ~5 s each), and groups that name the same
genome-wide clmfile share one parse of it.  `jobs=0` keeps the older modes:
`persistent=True` one `optimize-many -` process per GPU fed over stdin,
`persistent=False` one `optimize` process per group.  All modes leave
identical .tour files.
"""
import os
import re
import subprocess
import threading
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(_HERE, "bin", "allhic-b200")


def partition_refiles(contigsfile, k):
    """The per-group REfiles `allhic partition <counts_RE.txt> <pairs.txt> k` leaves behind (Partitioner.splitRE,
    partition.go:205-215): RemoveExt(contigsfile) + ".<k>g<j>.txt", j = 1..k."""
    root, _ = os.path.splitext(contigsfile)
    return ["%s.%dg%d.txt" % (root, k, j + 1) for j in range(k)]


def pipeline_groups(contigsfile, k, clmfile):
    """The (REfile, clmfile) list `allhic pipeline` optimizes one after the other (allhic.go:285-293): every group of
    the k-way partition against the same genome-wide CLM (the host binary parses that file once for all of them).
    Feed it to optimize_groups()."""
    refiles = partition_refiles(contigsfile, k)
    missing = [r for r in refiles if not os.path.exists(r)]
    if missing:
        raise FileNotFoundError("partition output not found: %s" % ", ".join(missing))
    return [(r, clmfile) for r in refiles]


def _count_contigs(refile):
    with open(refile) as f:
        return sum(1 for line in f if line.strip() and not line.startswith("#"))


def _device_env(dev):
    # the process sees its GPU only: CUDA start-up enumerates every visible device, which costs
    # seconds per process on an 8-GPU box
    env = dict(os.environ)
    vis = [x for x in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
    env["CUDA_VISIBLE_DEVICES"] = vis[dev] if dev < len(vis) else str(dev)
    return env


def _persistent_worker(dev, groups, order, lock, results, flags, binary, log_dir):
    """One `optimize-many -` process on device `dev`: write a group to its stdin, read its stdout up
    to the AHX_GROUP_DONE line, repeat.  A worker that dies fails the group it was on and is restarted."""
    proc = None
    errf = open(os.path.join(log_dir, "device%d.stderr" % dev), "w") if log_dir else subprocess.DEVNULL
    try:
        while True:
            with lock:
                if not order:
                    break
                g = order.pop(0)
                seq = len(groups) - len(order) - 1          # dispatch order: 0 = first group handed out
            refile, clmfile = groups[g]
            if proc is None or proc.poll() is not None:
                proc = subprocess.Popen([binary, "optimize-many", "-", "--device", "0", *flags], stdin=subprocess.PIPE,
                                        stdout=subprocess.PIPE, stderr=errf, text=True, bufsize=1, env=_device_env(dev))
            t0 = time.perf_counter()
            out, ok, answered = [], False, False
            try:
                proc.stdin.write("%s %s\n" % (refile, clmfile))
                proc.stdin.flush()
                for line in proc.stdout:
                    if line.startswith("AHX_GROUP_DONE ") or line.startswith("AHX_GROUP_FAILED "):   # FAILED: log.Fatal of this group only
                        ok, answered = line.startswith("AHX_GROUP_DONE "), True
                        break
                    out.append(line)
            except (BrokenPipeError, OSError):
                pass
            if not answered:                            # the worker died on this group: start a fresh one for the next
                proc.kill()
                proc.wait()
                proc = None
            dt = time.perf_counter() - t0
            scores = re.findall(r"max_score=([0-9.eE+-]+)", "".join(out))
            results[g] = dict(group=g, device=dev, seq=seq, seconds=dt, n_contigs=_count_contigs(refile),
                              final_score=float(scores[-1]) if scores else None, ok=ok)
    finally:
        if proc is not None and proc.poll() is None:
            try:
                proc.stdin.close()
                proc.wait(timeout=60)
            except Exception:
                proc.kill()
        if log_dir:
            errf.close()


def _one_process(groups, devices, flags, binary, log_dir, jobs):
    """`optimize-many --gpus G --jobs J` over a list file; parses its per-group blocks of stdout."""
    import tempfile
    env = dict(os.environ)
    vis = [x for x in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip()]
    env["CUDA_VISIBLE_DEVICES"] = ",".join(vis[d] if d < len(vis) else str(d) for d in devices)
    with tempfile.NamedTemporaryFile("w", suffix=".list", delete=False) as f:
        f.write("".join("%s %s\n" % g for g in groups))
        listfile = f.name
    errf = open(os.path.join(log_dir, "optimize-many.stderr"), "w") if log_dir else subprocess.DEVNULL
    results = [None] * len(groups)
    try:
        proc = subprocess.Popen([binary, "optimize-many", listfile, "--device", "0", "--gpus", str(len(devices)), "--jobs", str(jobs), *flags],
                                stdout=subprocess.PIPE, stderr=errf, text=True, bufsize=1, env=env)
        block, seq = [], 0
        for line in proc.stdout:
            m = re.match(r"AHX_GROUP_(DONE|FAILED) (\d+) (\S+) device=(\d+) seconds=([0-9.]+)", line)
            if not m:
                block.append(line)
                continue
            g = int(m.group(2))
            scores = re.findall(r"max_score=([0-9.eE+-]+)", "".join(block))
            results[g] = dict(group=g, device=devices[int(m.group(4))], seq=seq, seconds=float(m.group(5)), n_contigs=_count_contigs(groups[g][0]),
                              final_score=float(scores[-1]) if scores else None, ok=m.group(1) == "DONE")
            block, seq = [], seq + 1
        proc.wait()
    finally:
        os.unlink(listfile)
        if log_dir:
            errf.close()
    for g in range(len(groups)):
        if results[g] is None:
            results[g] = dict(group=g, device=None, seq=None, seconds=0.0, n_contigs=_count_contigs(groups[g][0]), final_score=None, ok=False)
    return results


def optimize_groups(groups, devices, flags=(), binary=BINARY, log_dir=None, persistent=True, jobs=6):
    """groups: list of (refile, clmfile); devices: CUDA device indices to farm over.
    Returns a list of dicts (group, device, seq = dispatch order, seconds, final_score, n_contigs, ok) in input order."""
    if jobs and jobs > 0:
        return _one_process(groups, list(devices), flags, binary, log_dir, jobs)
    order = sorted(range(len(groups)), key=lambda g: -_count_contigs(groups[g][0]) ** 2)
    lock = threading.Lock()
    results = [None] * len(groups)
    if persistent:
        threads = [threading.Thread(target=_persistent_worker, args=(d, groups, order, lock, results, flags, binary, log_dir))
                   for d in devices]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        return results

    def worker(dev):
        while True:
            with lock:
                if not order:
                    return
                g = order.pop(0)
                seq = len(groups) - len(order) - 1
            refile, clmfile = groups[g]
            t0 = time.perf_counter()
            env = _device_env(dev)
            p = subprocess.run([binary, "optimize", refile, clmfile, "--device", "0", *flags],
                               capture_output=True, text=True, env=env)
            dt = time.perf_counter() - t0
            scores = re.findall(r"max_score=([0-9.eE+-]+)", p.stdout)
            if log_dir:
                with open(os.path.join(log_dir, "group%02d.stderr" % g), "w") as f:
                    f.write(p.stderr)
            results[g] = dict(group=g, device=dev, seq=seq, seconds=dt, n_contigs=_count_contigs(refile),
                              final_score=float(scores[-1]) if scores else None,
                              ok=p.returncode == 0 and "Success" in p.stderr)

    threads = [threading.Thread(target=worker, args=(d,)) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return results
