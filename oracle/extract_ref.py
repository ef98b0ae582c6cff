"""TEST INFRASTRUCTURE -- CPU restatement of the CLM writer of `allhic extract`
(extract.go:406-538, Extracter.extractContigLinks), the upstream producer of the
`.clm` text that `optimize` ingests.  Only tests may import this.

It restates, with the reference's own rules and order of tests, what happens to
one BAM record (extract.go:445-492) and how the collected links are written
(extract.go:507-531).  The BAM container itself (biogo/hts, go.mod) is not
restated: a record here is the tuple of fields the loop reads --
(ref name, mate ref name, pos, mate pos, mapq, flags).

Parity unpinned (no Go toolchain in this image): the restatement follows the
source line by line; it is used to check that allhic_b200/synth.py writes what
`extract` would write for the same read pairs.
"""
from collections import OrderedDict

MIN_LINK_DIST = 1 << 11   # base.go:44  MinLinkDist (intra-contig links only, extract.go:469)
MIN_LINKS = 3             # base.go:51  MinLinks


def extract_contig_links(contigs, records, min_links=MIN_LINKS):
    """contigs: list of (name, length) in `.ids` order (readFastaAndWriteRE, extract.go:112-180);
    records: iterable of (ref, mate_ref, pos, mate_pos, mapq, flags).
    Returns (clm_lines, intra_links) -- clm_lines in the insertion order of the contig pairs (Go iterates its map in
    random order, extract.go:510: compare as a set); intra_links[i] = list of intra-contig link sizes of contig i."""
    contig_to_idx = {name: i for i, (name, _) in enumerate(contigs)}
    intra = [[] for _ in contigs]
    contig_pairs = OrderedDict()
    for ref, mate_ref, pos, mate_pos, mapq, flags in records:
        # Filtering: Unmapped | Secondary | QCFail | Duplicate | Supplementary   (extract.go:445-448)
        if mapq == 0 or (flags & 3844) != 0:
            continue
        # Make sure we have these contig ids                                      (extract.go:450-459)
        ai = contig_to_idx.get(ref)
        if ai is None:
            continue
        bi = contig_to_idx.get(mate_ref)
        if bi is None:
            continue
        apos, bpos = pos, mate_pos
        # An intra-contig link                                                     (extract.go:468-474)
        if ai == bi:
            link = abs(apos - bpos)
            if link >= MIN_LINK_DIST:
                intra[ai].append(link)
            continue
        # An inter-contig link: smaller contig index first                        (extract.go:476-481)
        if ai > bi:
            ai, bi = bi, ai
            apos, bpos = bpos, apos
        L1, L2 = contigs[ai][1], contigs[bi][1]
        apos2, bpos2 = L1 - apos, L2 - bpos                                       # extract.go:483-491
        ApBp = apos2 + bpos
        ApBm = apos2 + bpos2
        AmBp = apos + bpos
        AmBm = apos + bpos2
        contig_pairs.setdefault((ai, bi), []).append((ApBp, ApBm, AmBp, AmBm))
    tags = ["++", "+-", "-+", "--"]                                               # extract.go:509
    lines = []
    for (ai, bi), links in contig_pairs.items():
        for i in range(4):
            links_with_dir = [link[i] for link in links]                          # extract.go:512-515 (no sort, no unique)
            n_links = len(links_with_dir)
            if n_links < min_links:                                               # extract.go:521-523
                continue
            at, bt = contigs[ai][0], contigs[bi][0]
            lines.append("%s%s %s%s\t%d\t%s\n" % (at, tags[i][0], bt, tags[i][1], n_links,
                                                   " ".join(map(str, links_with_dir))))   # extract.go:527-530
    return lines, intra


def partition_refiles(contigsfile, k):
    """Names of the per-group REfiles `partition` writes (Partitioner.splitRE, partition.go:205-215):
    RemoveExt(contigsfile) + ".%dg%d.txt" % (K, j + 1); RemoveExt strips the last extension (base.go:125-127)."""
    import os
    root, _ = os.path.splitext(contigsfile)
    return ["%s.%dg%d.txt" % (root, k, j + 1) for j in range(k)]
