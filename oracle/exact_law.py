"""Test infrastructure only (oracle/): exact output laws on small graphs.

`p_ref(dump, k)`  -- the reference UGS law of SURVEY.md section 0.1 / Appendix A.5-A.6:
    P(S) = b[r]/Z * sum over growth orders of prod 1/|cut_i|, computed by exhaustive recursion over the
    growth process of rand_grow (reference sampler.cpp:36-85) from every viable root (sampler.cpp:117-173).
`connected_subsets(rows, k)` -- all connected induced k-subsets (the uniform law's support), the job the
    reference's uniform_sampler does by C(n,k) enumeration (uniform_sampler.cpp).
`p_epsilon(rows, k, eps)` -- the accepted-sample law of the epsilon_uniform sampler (Appendix A.9).
Pure Python: small cases only.
"""
from __future__ import annotations

from collections import defaultdict
from itertools import combinations


def rows_from_dump(d):
    n = len(d["order"])
    return [list(map(int, d["indices"][int(d["indptr"][v]):int(d["indptr"][v + 1])])) for v in range(n)]


def p_ref(d, k: int):
    """dict {frozenset(S) (complete) or tuple (partial row): probability} under the reference law."""
    rows = rows_from_dump(d)
    n = len(rows)
    index_of = [int(x) for x in d["index_of"]]
    order = [int(x) for x in d["order"]]
    b = [float(x) for x in d["bucket_b"]]
    sd = [int(x) for x in d["suffix_deg"]]
    Z = sum(b)
    if Z > 0:
        roots = {vi: b[vi] / Z for vi in range(n) if b[vi] > 0}
    else:
        el = [vi for vi in range(n) if sd[vi] > 0] or list(range(n))
        roots = {vi: 1.0 / len(el) for vi in el}
    out = defaultdict(float)

    def rec(S, root_vi, p):
        if len(S) == k:
            out[frozenset(S)] += p
            return
        cut = sorted({w for u in S for w in rows[u] if index_of[w] >= root_vi and w not in S})
        if not cut:
            out[tuple(S)] += p  # partial row (only at relaxation levels 1/2)
            return
        for w in cut:
            rec(S + [w], root_vi, p / len(cut))

    for vi, pr in roots.items():
        rec([order[vi]], vi, pr)
    return dict(out)


def connected_subsets(rows, k: int):
    n = len(rows)
    adj = [set(r) - {v} for v, r in enumerate(rows)]
    res = []
    for comb in combinations(range(n), k):
        s = set(comb)
        seen = {comb[0]}
        stack = [comb[0]]
        while stack:
            u = stack.pop()
            for w in adj[u]:
                if w in s and w not in seen:
                    seen.add(w)
                    stack.append(w)
        if len(seen) == k:
            res.append(frozenset(comb))
    return res


def chi_square(counts: dict, law: dict, total: int):
    """Pearson statistic, degrees of freedom, and #observations outside the law's support."""
    stat = 0.0
    for key, p in law.items():
        e = p * total
        o = counts.get(key, 0)
        stat += (o - e) ** 2 / e
    extra = sum(c for key, c in counts.items() if key not in law)
    return stat, len(law) - 1, extra


def p_epsilon(rows, k: int, eps: float, max_attempts=None):
    """Exact law of one output row of the epsilon_uniform sampler (SURVEY.md Appendix A.9,
    epsilon_uniform_sampler.cpp:18-84,225-257). rows = adjacency lists WITH duplicates.
    Returns {frozenset(S): prob, (): prob of an all -1 row}."""
    n = len(rows)
    if max_attempts is None:
        max_attempts = max(10, int(10.0 / eps))
    acc = defaultdict(float)  # per-attempt accepted mass per node set

    def rec(S, frontier, w, p, depth):
        if len(S) == k:
            a = min(1.0, eps / (w + eps))
            acc[frozenset(S)] += p * a
            return
        if not frontier or depth > 100 * k:
            return  # growth failure: the attempt is wasted
        F = len(frontier)
        for fi, u in enumerate(frontier):
            cand = [v for v in rows[u] if v not in S]
            if not cand:
                rec(S, frontier[:fi] + frontier[fi + 1:], w, p / F, depth + 1)
                continue
            C = len(cand)
            mult = defaultdict(int)
            for v in cand:
                mult[v] += 1
            for v, c in mult.items():
                rec(S + [v], frontier + [v], w * (1.0 / (F + 1)) * (1.0 / C), p / F * c / C, depth + 1)

    if n >= k:
        for s in range(n):
            rec([s], [s], 1.0 / n, 1.0 / n, 0)
    A = sum(acc.values())
    out = {}
    if A > 0:
        scale = (1.0 - (1.0 - A) ** max_attempts) / A
        for key, v in acc.items():
            out[key] = v * scale
    out[()] = (1.0 - A) ** max_attempts if A < 1 else 0.0
    return out


def p_rwr(rows, k: int, p_restart: float, steps: int = 4000):
    """Law of one output row of rwr_sampler (rwr_sampler.cpp:165-204) as {tuple(visiting order): prob, (): failure},
    ignoring the n*k*10 step cap (its truncated mass is far below the test resolution on the small cases it is used
    for, except where the walk cannot reach k vertices at all -- that mass lands in () here as well).
    rows = adjacency lists WITH duplicates. The walk inside the current set S is a Markov chain with restart; the
    distribution of the first vertex outside S it steps on is accumulated over `steps` transitions."""
    n = len(rows)
    out = defaultdict(float)
    if n < k or n == 0:
        return {(): 1.0}

    def exits(seq, cur):
        """{new vertex: prob} for the walk at `cur` with visited sequence seq (seq[0] = seed)."""
        S = set(seq)
        seed = seq[0]
        dist = {cur: 1.0}
        ex = defaultdict(float)
        for _ in range(steps):
            nd = defaultdict(float)
            for u, pu in dist.items():
                if pu < 1e-18:
                    continue
                if not rows[u]:
                    nd[seed] += pu
                    continue
                nd[seed] += pu * p_restart
                share = pu * (1.0 - p_restart) / len(rows[u])
                if share == 0.0:
                    continue
                for v in rows[u]:
                    if v in S:
                        nd[v] += share
                    else:
                        ex[v] += share
            dist = nd
            if sum(dist.values()) < 1e-15:
                break
        return ex, sum(dist.values())

    def rec(seq, cur, p):
        if len(seq) == k:
            out[tuple(seq)] += p
            return
        ex, stuck = exits(seq, cur)
        out[()] += p * stuck
        for v, pv in ex.items():
            rec(seq + [v], v, p * pv)

    for s in range(n):
        rec([s], s, 1.0 / n)
    return {key: v for key, v in out.items() if v > 0}
