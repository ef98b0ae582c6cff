"""Seeded synthetic linkage groups in the reference's wire formats.

The reference ships no generator; this one follows SURVEY.md 8(d): it mimics
what `allhic extract` writes for one chromosome-sized group
(extract.go:470-531 for the four oriented link distances, base.go:44-54 for
MinLinkDist/MaxLinkDist/MinLinks, extract.go:523 for the >=MinLinks filter)
so that the product and the oracle ingest the *same* `.ids` / `.clm` text
through their own parsers.  Ground truth order = identity, all '+'.

BASELINE.json configs -> (N, seed):  config2 (500, 4202), config3 (2000, 4203),
config4 group g (U[700,900], 4204+g), config5 (10000, 4205).
"""
import os

import numpy as np

MIN_LINK_DIST = 1 << 11   # base.go:44
MAX_LINK_DIST = 1 << 27   # base.go:54
MIN_LINKS = 3             # base.go:51

CONFIGS = {
    "config2": dict(n=500, seed=4202, npop=1000),
    "config3": dict(n=2000, seed=4203, npop=8000),
    "config5": dict(n=10000, seed=4205, npop=8000),
}


def config4_groups():
    """48 linkage groups, N_g ~ U[700,900], seeds 4204+g."""
    rng = np.random.Generator(np.random.PCG64(4204))
    ns = rng.integers(700, 901, size=48)
    return [dict(n=int(n), seed=4204 + g, npop=1000) for g, n in enumerate(ns)]


def simulate_read_pairs(n, seed, pairs_per_contig=300, mean_size=100_000):
    """One chromosome of n contigs laid end to end in true order and R = pairs_per_contig * n read pairs with a
    1/d separation spectrum on [MinLinkDist, MaxLinkDist]: what the BAM that `extract` reads would hold.
    Returns dict(sizes, G, ai, apos, bi, bpos): the inter-contig pairs as (contig, 0-based position) of read and
    mate, read on the left."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = np.clip(rng.exponential(mean_size, n), 1_000, 2_000_000).astype(np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    G = int(starts[-1])
    R = pairs_per_contig * n
    x = rng.integers(0, G, size=R)
    d = np.exp(rng.uniform(np.log(MIN_LINK_DIST), np.log(MAX_LINK_DIST), size=R)).astype(np.int64)
    y = x + d
    keep = y < G
    x, y = x[keep], y[keep]
    ai = np.searchsorted(starts, x, side="right") - 1
    bi = np.searchsorted(starts, y, side="right") - 1
    inter = ai != bi
    x, y, ai, bi = x[inter], y[inter], ai[inter], bi[inter]
    return dict(sizes=sizes, G=G, ai=ai, apos=x - starts[ai], bi=bi, bpos=y - starts[bi])


def generate(n, seed, pairs_per_contig=300, mean_size=100_000):
    """Returns dict(names, sizes, pa, pb, dists) where dists[e] is a (4, nlinks)
    int64 array in file orientation order ++, +-, -+, -- for pair (pa[e] < pb[e]).
    The CLM lines are what `extract` writes for the simulated read pairs (extract.go:476-531; checked against a
    restatement of that loop by tests/test_extract_semantics.py), except that a pair's distances are written in
    ascending order (Go: BAM record order -- GoldenArray is a histogram, so the order only moves SumLog's rounding)
    and a zero distance (read and mate both at position 0 of abutting contigs) is written as 1 (ln 0 = -Inf)."""
    sim = simulate_read_pairs(n, seed, pairs_per_contig, mean_size)
    sizes, G, ai, bi, apos, bpos = sim["sizes"], sim["G"], sim["ai"], sim["bi"], sim["apos"], sim["bpos"]
    L1 = sizes[ai]
    L2 = sizes[bi]
    four = np.stack([(L1 - apos) + bpos, (L1 - apos) + (L2 - bpos), apos + bpos, apos + (L2 - bpos)], axis=0)
    order = np.lexsort((bi, ai))
    ai, bi, four = ai[order], bi[order], four[:, order]
    key = ai.astype(np.int64) * n + bi
    uniq, first, counts = np.unique(key, return_index=True, return_counts=True)
    sel = counts >= MIN_LINKS
    pa = (uniq[sel] // n).astype(np.int32)
    pb = (uniq[sel] % n).astype(np.int32)
    dists = []
    for f, c in zip(first[sel], counts[sel]):
        blk = np.sort(four[:, f:f + c], axis=1)
        dists.append(np.maximum(blk, 1))
    names = ["tig%05d" % i for i in range(n)]
    return dict(names=names, sizes=sizes, pa=pa, pb=pb, dists=dists, G=G)


def write(group, prefix):
    """Writes <prefix>.ids (REfile, clm.go:141-157) and <prefix>.clm
    (clm.go:22-31, extract.go:529-530).  Returns (idsfile, clmfile)."""
    ids = prefix + ".ids"
    clm = prefix + ".clm"
    with open(ids, "w") as f:
        f.write("#Contig\tRECounts\tLength\n")
        for name, s in zip(group["names"], group["sizes"]):
            f.write("%s\t%d\t%d\n" % (name, max(1, int(s) // 256), int(s)))
    tags = ["++", "+-", "-+", "--"]
    names = group["names"]
    with open(clm, "w") as f:
        out = []
        for a, b, blk in zip(group["pa"], group["pb"], group["dists"]):
            for o in range(4):
                out.append("%s%s %s%s\t%d\t%s\n" % (names[a], tags[o][0], names[b], tags[o][1], blk.shape[1],
                                                     " ".join(map(str, blk[o].tolist()))))
        f.write("".join(out))
    return ids, clm


def materialize(name_or_n, seed=None, outdir=None, **kw):
    """Generate (or reuse) a named config; returns (idsfile, clmfile, info)."""
    if isinstance(name_or_n, str):
        cfg = CONFIGS[name_or_n]
        n, seed, tag = cfg["n"], cfg["seed"], name_or_n
    else:
        n, tag = int(name_or_n), "n%d_s%d" % (int(name_or_n), seed)
    outdir = outdir or os.path.join(os.environ.get("TMPDIR", "/tmp"), "allhic_b200_synth")
    os.makedirs(outdir, exist_ok=True)
    prefix = os.path.join(outdir, "%s_seed%d" % (tag, seed))
    if not (os.path.exists(prefix + ".ids") and os.path.exists(prefix + ".clm") and os.path.exists(prefix + ".done")):
        g = generate(n, seed, **kw)
        write(g, prefix)
        with open(prefix + ".done", "w") as f:
            f.write("E=%d G=%d\n" % (len(g["pa"]), g["G"]))
    return prefix + ".ids", prefix + ".clm", dict(n=n, seed=seed)
