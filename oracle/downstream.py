"""TEST-ONLY port of the parts of the reference that STAY in Julia: the grouping by
CDR3 length, `nbsorb` error correction (Jtool.jl:396-590), the Accumulator / Prob / Dregion / CSV formatting
(catt.jl:579-687).  It exists so that the oracle's and the CUDA path's Vpart/Jpart can be compared with the
reference's bundled golden CSV without Julia.  Never imported by the product.
"""
from __future__ import annotations

import csv
import math
from collections import Counter

CODON = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
_NT = {"A": 0, "C": 1, "G": 2, "T": 3}


def translate(nt: str) -> str:
    out = []
    for p in range(0, len(nt) - 2, 3):
        try:
            out.append(CODON[_NT[nt[p]] * 16 + _NT[nt[p + 1]] * 4 + _NT[nt[p + 2]]])
        except KeyError:
            out.append("X")
    return "".join(out)


def max_value(p, L, n0):                                   # Jtool.jl:437-440
    tmp = (3 * math.sqrt(p * (1 - p)) + math.sqrt(9 * p * (1 - p) + 4 * (1 - p) ** L * n0)) / (2 * (1 - p) ** L)
    return tmp ** 2


def error_number(p, L, nt, x, alpha=3.0):                  # Jtool.jl:442-448
    ps = math.comb(L, x) * p ** x * (1 - p) ** (L - x) if L >= x >= 0 else 0.0
    return ps * nt + alpha * math.sqrt(nt * ps * (1 - ps))


def junfen(n, pb, alpha=2.576):                            # Jtool.jl:450-456
    if n < 1:
        return 0.0
    p = 1 / pb
    return max(n * p + alpha * math.sqrt(n * p * (1 - p)), 1.0)


def err_cnt_seq(s):                                        # Jtool.jl:462-464 (sums over .seq, sic)
    return sum(10 ** (-(ord(ch) - 33) / 10) for ch in s)


def MQ(x, y):                                              # Jtool.jl:467-469
    return "".join(max(a, b) for a, b in zip(x, y))


def hamming_g3(s1, s2, upper=10, skip=0):                  # Jtool.jl:65-74
    cnt = 0
    for i in range(skip, len(s1) - skip):
        if s1[i] != s2[i]:
            cnt += 1
            if cnt > upper:
                return False
    return True


def link_rlg(leaf, roots, uplimit):                        # Jtool.jl:396-421
    cnt, holder = 0, 1
    for idy, (root, _rc) in enumerate(roots, start=1):
        diffs = [abs(u - leaf[1]) for u in uplimit[idy - 1]]
        upper = diffs.index(min(diffs)) + 1
        if hamming_g3(leaf[0][0], root[0], upper, skip=3):
            cnt += 1
            if cnt > 1:
                return cnt, holder
            holder = idy
    return cnt, holder


def nbsorb(lgt, val, err_rate, sc_mode=False, link_fn=None, near_fn=None, groups=None):   # Jtool.jl:475-590; val = [(cdr3, qual), ...]
    # link_fn(lgt, root_seqs, root_up, leaf_seqs, leaf_counts) -> (statu[], holder[]) and near_fn(lgt, query_seqs, target_seqs) -> flag[]
    # stand in for the two scan loops (Jtool.jl:524-531, 553-558) when a test wants another implementation of them checked.
    # groups = [((cdr3, merged qual), count), ...] in cdr3 order replaces the sort / run-length / MQ step (Jtool.jl:481-489) when
    # a test wants the library's clonotype table checked as the source of those groups.
    if not val and not groups:
        return set(), {}
    if groups is not None:
        ss = list(groups)
    else:
        val = sorted(val, key=lambda x: x[0])
        ss = []
        i = 0
        while i < len(val):
            j = i
            q = val[i][1]
            while j + 1 < len(val) and val[j + 1][0] == val[i][0]:
                j += 1
                q = MQ(q, val[j][1])
            ss.append(((val[i][0], q), j - i + 1))
            i = j + 1
    ss.sort(key=lambda x: -x[1])
    p, L = err_rate, lgt - 6
    nt = math.floor(max_value(p, L, ss[0][1]))
    thresold = math.floor(junfen(error_number(p, L, nt, 1), L * 3 ** 1, 3.291))
    if sc_mode:
        return {x[0][0] for x in ss if x[1] > thresold}, {}
    np_ = math.floor(max_value(p, L, thresold))
    unconfirm = junfen(error_number(p, L, np_, 1), L * 3 ** 1, 3.291)
    maybe = [x for x in ss if unconfirm >= x[1]]
    leafs = [x for x in ss if unconfirm < x[1] < thresold]
    roots = [x for x in ss if x[1] >= thresold]
    if not roots:
        return {x[0][0] for x in leafs}, {}
    conf = {x[0][0] for x in roots}
    root_up = []
    for root in roots:
        nt = math.floor(max_value(p, L, root[1]))
        root_up.append([junfen(error_number(p, L, nt, x, 3.291), L * 3 ** x, 3.291) for x in range(1, 8)])
    trees = {}
    if link_fn is not None and leafs:
        st, hd = link_fn(lgt, [r[0][0] for r in roots], root_up, [x[0][0] for x in leafs], [x[1] for x in leafs])
        linked = list(zip([int(a) for a in st], [int(b) for b in hd]))
    else:
        linked = [link_rlg(leaf, roots, root_up) for leaf in leafs]
    for leaf, (statu, ridx) in zip(leafs, linked):
        if statu == 0:
            conf.add(leaf[0][0])
        elif statu == 1:
            trees.setdefault(ridx, []).append(leaf)
    highconf = leafs + roots
    jiuji = set()
    near = near_fn(lgt, [x[0][0] for x in maybe], [c[0][0] for c in highconf]) if near_fn is not None and maybe else None
    for im, x in enumerate(maybe):
        flag = False
        if err_cnt_seq(x[0][0]) > 1:
            flag = True
        elif near is not None:
            flag = bool(near[im])
        else:
            for confm in highconf:
                if hamming_g3(x[0][0], confm[0][0], 3, skip=3):
                    flag = True
                    break
        if err_cnt_seq(x[0][0]) < 1 and not flag:
            jiuji.add(x[0][0])
    seq2seq = {}
    for ridx, yezi in trees.items():
        root, rc = roots[ridx - 1]
        tfm = {}
        rg = range(3, lgt - 3)
        for (lseq, lqual), _c in yezi:
            for pos in rg:
                if pos < len(lqual):
                    item = (lseq[pos], pos)
                    pr = 10 ** (-(ord(lqual[pos]) - 33) / 10)
                    tfm[item] = tfm.get(item, 0) + math.log(pr)
        volm = math.floor(max_value(p, L, rc)) - rc

        def keyf(x):
            prod = 1.0
            for pos in rg:
                prod *= tfm.get((x[0][0][pos], pos), 0)
            return x[1] * prod
        yezi = sorted(yezi, key=keyf, reverse=True)
        for (lseq, _q), cnt in yezi:
            if cnt <= volm:
                volm -= cnt
                seq2seq[lseq] = root[0]
            else:
                conf.add(lseq)
    return conf | jiuji, seq2seq


def load_prob(path):
    with open(path) as fh:
        rows = list(csv.DictReader(fh))
    return {col: {r["AA"]: float(r[col]) for r in rows} for col in ("-4", "-3", "-2", "1", "2", "3")}


def prob_of(aa, tab):                                      # catt.jl:656-669
    try:
        v = (tab["1"][aa[1]] * tab["2"][aa[2]] * tab["3"][aa[3]] *
             tab["-2"][aa[-2]] * tab["-3"][aa[-3]] * tab["-4"][aa[-4]])
    except (KeyError, IndexError):
        return float("nan")
    return round(v, 5)


def clonotype_table(Vpart, Jpart, the_err, dlist=None, best_d=None, sc_mode=False, prob_tab=None, link_fn=None, near_fn=None):
    """catt.jl:583-686 from the filtered Vpart/Jpart to output rows.
    Returns Counter{(AAseq, NNseq, Vregion, Jregion, Dregion) -> Frequency} (row order is Julia Dict order in the
    reference, so rows are compared as a multiset)."""
    V = [r for r in Vpart if r.cdr3 != "None"]
    J = [r for r in Jpart if r.cdr3 != "None"]
    conf, seq2seq, res = set(), {}, []
    for lgt in range(21, 106):
        grp = [r for r in V if len(r.cdr3) == lgt] + [r for r in J if len(r.cdr3) == lgt]
        if not grp:
            continue
        cf, c2c = nbsorb(lgt, [(r.cdr3, r.qual) for r in grp], the_err, sc_mode, link_fn, near_fn)
        conf |= cf
        seq2seq.update(c2c)
        res += grp
    rows = Counter()
    for r in res:
        if r.cdr3 in conf or r.cdr3 in seq2seq:
            aa = translate(seq2seq.get(r.cdr3, r.cdr3))
            rows[(r.vs, aa, r.js, r.cdr3)] += 1
    out = Counter()
    for (vs, aa, js, nn), n in rows.items():
        d = best_d(nn, dlist) if dlist else ""
        out[(aa, nn, vs, js, d)] += n
    return out


def clonotype_table_from_rows(rows, the_err, dlist=None, best_d=None, sc_mode=False, link_fn=None, near_fn=None):
    """The same output rows from a clonotype table [(cdr3, merged qual, vs, js, count), ...] ordered by (len, cdr3, ...) - the
    library's catt_result_clonotypes - instead of from the reads: nbsorb's groups are the per-cdr3 sums / MQ maxima of the rows
    (Jtool.jl:481-489), the output counter (catt.jl:644) the rows themselves."""
    conf, seq2seq = set(), {}
    by_len = {}
    for r in rows:
        by_len.setdefault(len(r[0]), []).append(r)
    for lgt in range(21, 106):
        grp = by_len.get(lgt)
        if not grp:
            continue
        groups = []
        for cdr3, qual, _, _, count in grp:                  # rows of one cdr3 are adjacent
            if groups and groups[-1][0][0] == cdr3:
                (c, q), n = groups[-1]
                groups[-1] = ((c, MQ(q, qual)), n + count)
            else:
                groups.append(((cdr3, qual), count))
        cf, c2c = nbsorb(lgt, None, the_err, sc_mode, link_fn, near_fn, groups=groups)
        conf |= cf
        seq2seq.update(c2c)
    out = Counter()
    for cdr3, _, vs, js, count in rows:
        if 21 <= len(cdr3) <= 105 and (cdr3 in conf or cdr3 in seq2seq):
            aa = translate(seq2seq.get(cdr3, cdr3))
            d = best_d(cdr3, dlist) if dlist else ""
            out[(aa, cdr3, vs, js, d)] += count
    return out


def load_golden_csv(path):
    out = Counter()
    probs = {}
    with open(path) as fh:
        for r in csv.DictReader(fh):
            key = (r["AAseq"], r["NNseq"], r["Vregion"], r["Jregion"], r.get("Dregion", ""))
            out[key] += int(r["Frequency"])
            probs[key] = float(r["Prob"])
    return out, probs
