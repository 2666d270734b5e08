"""Shared helpers for the tests: synthetic read generation (SURVEY §8d config 3 generator) and comparisons."""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return x, z ^ (z >> 31)


class Rng:
    def __init__(self, seed):
        self.s = seed & MASK64

    def next(self):
        self.s, v = splitmix64(self.s)
        return v

    def below(self, n):
        return self.next() % n


COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def synth_read(i, vrecs, jrecs, length=150, seed=20251019, p_err=0.003, with_n=False):
    """One synthetic TCR read: 60 random nt + V (3' trimmed) + random insert + J (5' trimmed) + 120 random nt, a
    uniformly placed window containing >= 20 nt of V or >= 12 nt of J, random strand, iid substitutions."""
    g = Rng(seed + i)
    rnd = lambda n: "".join("ACGT"[g.below(4)] for _ in range(n))
    v = vrecs[g.below(len(vrecs))]
    j = jrecs[g.below(len(jrecs))]
    v = v[:len(v) - g.below(7)]
    j = j[g.below(7):]
    ins = rnd(g.below(19))
    pre = rnd(60)
    post = rnd(120)
    tpl = pre + v + ins + j + post
    v0, v1 = len(pre), len(pre) + len(v)
    j0, j1 = v1 + len(ins), v1 + len(ins) + len(j)
    for _ in range(64):
        st = g.below(max(1, len(tpl) - length + 1))
        en = st + length
        ov_v = max(0, min(en, v1) - max(st, v0))
        ov_j = max(0, min(en, j1) - max(st, j0))
        if ov_v >= 20 or ov_j >= 12:
            break
    read = list(tpl[st:st + length])
    qual = ["I"] * len(read)
    for k in range(len(read)):
        r = g.below(100000)
        if r < int(p_err * 100000):
            read[k] = "ACGT"[g.below(4)]
            qual[k] = "#"
        elif r < 5000 + int(p_err * 100000):
            qual[k] = "G"
    if with_n and g.below(50) == 0:
        read[g.below(len(read))] = "N"
    s = "".join(read)
    if g.below(2):
        s = revcomp(s)
        qual.reverse()
    return f"s{i}", s, "".join(qual)


def synth_reads(n, vrecs, jrecs, length=150, start=0, **kw):
    names, seqs, quals = [], [], []
    for i in range(start, start + n):
        a, b, c = synth_read(i, vrecs, jrecs, length, **kw)
        names.append(a); seqs.append(b); quals.append(c)
    return names, seqs, quals


def random_reads(n, length, seed):
    rng = np.random.default_rng(seed)
    arr = rng.integers(0, 4, size=(n, length))
    return ["".join("ACGT"[x] for x in row) for row in arr]


def myread_tuple(r):
    return (r.seq, r.qual, r.vs, r.js, r.cdr3, bool(r.able))
