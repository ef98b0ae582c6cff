"""SURVEY.md 8 (f4), upstream producers: what `allhic extract` writes into the .clm (extract.go:445-531) and how
`partition` names the per-group REfiles (partition.go:205-215).  The synthetic generator behind configs 2-5 claims to
mimic the former; here it is checked against a restatement of the reference's loop (oracle/extract_ref.py) on the same
simulated read pairs.  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import extract_ref  # noqa: E402
from allhic_b200 import farm, synth  # noqa: E402


def _records(sim, names, rng, noise=True):
    """BAM-like records for the simulated inter-contig pairs, either mate first (extract swaps so that the smaller
    contig index leads, extract.go:476-481), plus records the reference must drop."""
    recs = []
    for a, ap, b, bp in zip(sim["ai"], sim["apos"], sim["bi"], sim["bpos"]):
        if rng.random() < 0.5:
            recs.append((names[a], names[b], int(ap), int(bp), 60, 0x41))
        else:
            recs.append((names[b], names[a], int(bp), int(ap), 30, 0x81))
    if noise:
        n = len(names)
        for _ in range(500):
            a, b = rng.integers(0, n, 2)
            kind = rng.integers(0, 5)
            if kind == 0:
                recs.append((names[a], names[b], 10, 20, 0, 0x41))                       # MapQ == 0
            elif kind == 1:
                recs.append((names[a], names[b], 10, 20, 60, int(rng.choice([0x4, 0x100, 0x200, 0x400, 0x800]))))   # flags & 3844
            elif kind == 2:
                recs.append(("alien_contig", names[b], 10, 20, 60, 0x41))               # unknown reference
            elif kind == 3:
                recs.append((names[a], "alien_contig", 10, 20, 60, 0x41))               # unknown mate reference
            else:
                recs.append((names[a], names[a], 100, 100 + int(rng.integers(0, 5000)), 60, 0x41))   # intra-contig
        order = rng.permutation(len(recs))
        recs = [recs[i] for i in order]
    return recs


def _parse(lines):
    out = {}
    for ln in lines:
        tag, n, d = ln.rstrip("\n").split("\t")
        d = sorted(max(1, int(x)) for x in d.split(" "))
        assert len(d) == int(n)
        assert tag not in out
        out[tag] = d
    return out


def test_synthetic_clm_is_what_extract_writes(tmp_path):
    n, seed = 150, 77
    sim = synth.simulate_read_pairs(n, seed)
    g = synth.generate(n, seed)
    ids, clmf = synth.write(g, str(tmp_path / "s"))
    contigs = [(nm, int(sz)) for nm, sz in zip(g["names"], g["sizes"])]
    rng = np.random.default_rng(5)
    lines, intra = extract_ref.extract_contig_links(contigs, _records(sim, g["names"], rng))
    want = _parse(lines)
    got = _parse(open(clmf).readlines())
    assert len(want) > 1000 and set(got) == set(want)
    assert all(got[k] == want[k] for k in want)
    # every kept pair has all four orientations with the same link count, smaller contig index first
    for a, b, blk in zip(g["pa"], g["pb"], g["dists"]):
        assert a < b and blk.shape[0] == 4 and blk.shape[1] >= extract_ref.MIN_LINKS
    # intra-contig links never reach the clm; only those >= MinLinkDist are kept at all
    assert all(x >= extract_ref.MIN_LINK_DIST for c in intra for x in c)


def test_extract_restatement_known_answer():
    """Hand-computed: two contigs of 1000 and 500 bp, read at 900 of A, mate at 100 of B (extract.go:483-491)."""
    contigs = [("A", 1000), ("B", 500)]
    recs = [("A", "B", 900, 100, 60, 0)] * 3 + [("B", "A", 100, 900, 60, 0)]
    lines, _ = extract_ref.extract_contig_links(contigs, recs)
    assert lines == ["A+ B+\t4\t200 200 200 200\n", "A+ B-\t4\t500 500 500 500\n",
                     "A- B+\t4\t1000 1000 1000 1000\n", "A- B-\t4\t1300 1300 1300 1300\n"]
    lines, _ = extract_ref.extract_contig_links(contigs, recs[:2])          # below MinLinks: nothing is written
    assert lines == []


def test_partition_refile_names(tmp_path):
    c = str(tmp_path / "sample.counts_GATC.txt")
    want = [str(tmp_path / ("sample.counts_GATC.3g%d.txt" % j)) for j in (1, 2, 3)]
    assert farm.partition_refiles(c, 3) == want == extract_ref.partition_refiles(c, 3)
    for w in want:
        open(w, "w").write("#Contig\tRECounts\tLength\n")
    assert farm.pipeline_groups(c, 3, "x.clm") == [(w, "x.clm") for w in want]
