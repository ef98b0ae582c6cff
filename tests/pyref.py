"""Independent pure-Python restatement of the reference's optimize scoring path
(TEST INFRASTRUCTURE; small inputs only).

Written separately from oracle/allhic_oracle.c, with Python dicts standing in
for the Go maps, so that a transcription slip in either restatement shows up as
a disagreement.  Citations are to the reference repo (tanghaibao/allhic).
Python's math.log is libm's, not Go's FDLIBM port: sums involving logs are
compared with a 1e-13 relative tolerance, everything else exactly.
"""
import math

LB, UB, BB = 18, 29, 12                      # base.go:30-34
PHI = 0.4812118250596684                     # base.go:36
LIMIT = 10000000                             # evaluate.go:22
GR = [5778, 9349, 15127, 24476, 39603, 64079, 103682, 167761, 271443, 439204, 710647, 1149851]  # base.go:101


def go_round(x):                             # base.go:130-135
    return math.ceil(x - 0.5) if x < 0 else math.floor(x + 0.5)


def golden_array(a):                         # base.go:156-167
    counts = [0] * BB
    for x in a:
        c = int(go_round(math.log(float(x)) / PHI))
        c = LB if c < LB else (UB if c > UB else c)
        counts[c - LB] += 1
    return counts


def sum_log(a):                              # base.go:138-144
    s = 0.0
    for v in a:
        s += math.log(float(v))
    return s


def rr(b):                                   # clm.go:160-165
    return '+' if b == '-' else '-'


class PyCLM:
    def __init__(self, clmfile, refile):     # clm.go:121-134
        self.names, self.sizes, self.tig2idx = [], [], {}
        self.contacts, self.oriented = {}, {}
        for line in open(refile):            # clm.go:141-157
            words = line.split()
            if not words or words[0][0] == '#':
                continue
            self.tig2idx[words[0]] = len(self.names)
            self.names.append(words[0])
            self.sizes.append(int(words[-1]))
        for row in open(clmfile):            # clm.go:168-240
            row = row.strip()
            if not row:
                continue
            words = row.split('\t')
            atig, btig = words[0].split(' ')[:2]
            at, ao, bt, bo = atig[:-1], atig[-1], btig[:-1], btig[-1]
            links = [int(x) for x in words[2].split(' ')]
            if at not in self.tig2idx or bt not in self.tig2idx:
                continue
            ai, bi = self.tig2idx[at], self.tig2idx[bt]
            g = golden_array(links)
            md = sum_log(links)
            strand = 1 if ao == bo else -1
            c = (strand, len(links), md)
            if (ai, bi) in self.contacts:
                if md < self.contacts[(ai, bi)][2]:
                    self.contacts[(ai, bi)] = c
            else:
                self.contacts[(ai, bi)] = c
            self.oriented[(ai, bi, ao, bo)] = g
            self.oriented[(bi, ai, rr(bo), rr(ao))] = g
        self.N = len(self.names)
        self.signs = ['+'] * self.N
        self.tour = list(range(self.N))

    def M(self):                             # clm.go:437-447
        P = [[0] * self.N for _ in range(self.N)]
        for (ai, bi), c in self.contacts.items():
            P[ai][bi] = c[1]
            P[bi][ai] = c[1]
        return P

    def evaluate(self, tour):                # evaluate.go:116-143
        M = self.M() if not hasattr(self, "_M") else self._M
        self._M = M
        mid, cum = [], 0.0
        for t in tour:
            s = float(self.sizes[t])
            mid.append(cum + s / 2)
            cum += s
        score = 0.0
        n = len(tour)
        for i in range(n):
            a = tour[i]
            for j in range(i + 1, n):
                dist = mid[j] - mid[i]
                if dist > LIMIT:
                    break
                score -= float(M[a][tour[j]]) / dist
        return score

    def evaluate_q(self, tour, signs):       # orientation.go:137-193
        cumsize, cum = [], 0
        for t in tour:
            cumsize.append(cum)
            cum += self.sizes[t]
        score = 0.0
        n = len(tour)
        for i in range(n):
            a = tour[i]
            for j in range(i + 1, n):
                b = tour[j]
                g = self.oriented.get((a, b, signs[a], signs[b]))
                if g is None:
                    continue
                dist = cumsize[j - 1] - cumsize[i]
                if dist > LIMIT:
                    break
                for k in range(BB):
                    score -= float(g[k]) * math.log(float(GR[k] + dist))
        return score
