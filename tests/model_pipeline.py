"""Executable model (pure Python, small inputs) of the *device* pipeline's
algorithm: bucket-sequential / round-parallel induced sorting at level 0 and
rank-pair doubling on the reduced string.  It exists to validate the parallel
formulation (step semantics, validity ranges, fill accounting) against the
oracle on a CPU, and to produce expected intermediate arrays for the GPU
kernel unit tests.  TEST INFRASTRUCTURE ONLY -- never imported by suffix_b200.
"""


def classify(T):
    """S-type bit per position, LMS bit per position (reference semantics:
    last char is L, src/table.rs:592-615)."""
    n = len(T)
    S = [0] * n
    for i in range(n - 2, -1, -1):
        if T[i] < T[i + 1]:
            S[i] = 1
        elif T[i] > T[i + 1]:
            S[i] = 0
        else:
            S[i] = S[i + 1]
    lms = [0] * n
    for i in range(1, n):
        lms[i] = 1 if (S[i] and not S[i - 1]) else 0
    return S, lms


def bucket_tables(T, S, sigma=256):
    Lc = [0] * sigma
    Sc = [0] * sigma
    for c, s in zip(T, S):
        if s:
            Sc[c] += 1
        else:
            Lc[c] += 1
    bstart = [0] * (sigma + 1)
    for c in range(sigma):
        bstart[c + 1] = bstart[c] + Lc[c] + Sc[c]
    return Lc, Sc, bstart


def _step(T, SA, src, logical, lo, hi, spass, bstart, fill):
    """One partition step over a logical list of *physical slot indices*
    `logical` into `src`: every entry s with s>0 and lo <= T[s-1] <= hi emits
    s-1 into bucket d=T[s-1], stably.  L pass fills heads upward, S pass fills
    tails downward.  Returns per-destination counts (fill is updated)."""
    cnt = {}
    for p in logical:
        s = src[p]
        if s == 0:
            continue
        d = T[s - 1]
        if d < lo or d > hi:
            continue
        k = fill[d] + cnt.get(d, 0)
        if spass:
            SA[bstart[d + 1] - 1 - k] = s - 1
        else:
            SA[bstart[d] + k] = s - 1
        cnt[d] = cnt.get(d, 0) + 1
    for d, v in cnt.items():
        fill[d] += v
    return cnt


def induce(T, SA, lms_list, lms_off, Lc, Sc, bstart, sigma=256, stats=None):
    """L pass then S pass.  lms_list is grouped by first char with offsets
    lms_off[c]; within a group the order is the caller's (text order in stage
    1, sorted order in stage 2)."""
    n = len(T)
    # ---- L pass (reference P6-P7 / P19: head inserts, left-to-right) ----
    fill = [0] * sigma
    c_last = T[n - 1]
    SA[bstart[c_last]] = n - 1            # seed: suffix n-1 is L by definition
    fill[c_last] = 1
    steps = 0
    for c in range(sigma):
        if bstart[c + 1] == bstart[c]:
            continue
        begin, end = 0, fill[c]
        while end > begin:                # chain rounds inside bucket c
            _step(T, SA, SA, range(bstart[c] + begin, bstart[c] + end), c, sigma - 1,
                  False, bstart, fill)
            steps += 1
            begin, end = end, fill[c]
        assert fill[c] == Lc[c], (c, fill[c], Lc[c])
        if lms_off[c + 1] > lms_off[c] and c + 1 <= sigma - 1:
            _step(T, SA, lms_list, range(lms_off[c], lms_off[c + 1]), c + 1, sigma - 1,
                  False, bstart, fill)
            steps += 1
    # ---- S pass (reference P8-P9 / P20: tail inserts, right-to-left) ----
    fill = [0] * sigma
    for c in range(sigma - 1, -1, -1):
        if bstart[c + 1] == bstart[c]:
            continue
        begin, end = 0, fill[c]
        while end > begin:
            top = bstart[c + 1] - 1
            _step(T, SA, SA, range(top - begin, top - end, -1), 0, c, True, bstart, fill)
            steps += 1
            begin, end = end, fill[c]
        assert fill[c] == Sc[c], (c, fill[c], Sc[c])
        if Lc[c] > 0 and c > 0:
            top = bstart[c] + Lc[c] - 1
            _step(T, SA, SA, range(top, bstart[c] - 1, -1), 0, c - 1, True, bstart, fill)
            steps += 1
    if stats is not None:
        stats["steps"] = stats.get("steps", 0) + steps


def _multiround_chain(T, SA, c, begin, end, R, spass, bstart, fill):
    """The cascade step of k_induce6 (induce6.cuh, R = CAS_R = 5): one partition step that plays up
    to R chain rounds of bucket c at once.  Entry e_j of the list [begin,end) with
    l_j = min(run of c to its left, R) emits e_j - r into round r (1 <= r <= l_j) of the
    chain region, round-major then list order; if l_j < R its terminal e_j - l_j
    induces into bucket d = T[e_j - l_j - 1] (valid range as in a normal chain step),
    ordered by (l_j, j).  Returns the new `begin` (first unprocessed slot)."""
    sigma = len(fill)
    F = fill[c]
    def slot(d, k):
        return bstart[d + 1] - 1 - k if spass else bstart[d] + k
    ents = [SA[slot(c, k)] for k in range(begin, end)]
    runs = []
    for e in ents:
        l = 0
        while l < R and e - 1 - l >= 0 and T[e - 1 - l] == c:
            l += 1
        runs.append(l)
    a = [sum(1 for l in runs if l >= r) for r in range(0, R + 1)]      # a[r], r >= 1
    off = F
    for r in range(1, R + 1):
        k = 0
        for e, l in zip(ents, runs):
            if l >= r:
                SA[slot(c, off + k)] = e - r
                k += 1
        off += a[r]
    new_begin = F + sum(a[1:R])
    fill[c] = F + sum(a[1:R + 1])
    # terminals, ordered by (l, j)
    order = sorted(range(len(ents)), key=lambda j: (runs[j], j))
    cnt = {}
    for j in order:
        l, e = runs[j], ents[j]
        if l == R:
            continue                      # still inside its run: continues in the next step
        x = e - l
        if x == 0:
            continue
        d = T[x - 1]
        ok = (d < c) if spass else (d > c)
        if not ok:
            continue
        k = fill[d] + cnt.get(d, 0)
        SA[slot(d, k)] = x - 1
        cnt[d] = cnt.get(d, 0) + 1
    for d, v in cnt.items():
        fill[d] += v
    return new_begin


def induce_multiround(T, SA, lms_list, lms_off, Lc, Sc, bstart, R, sigma=256):
    """induce() with the chain rounds of every bucket played R at a time."""
    n = len(T)
    fill = [0] * sigma
    SA[bstart[T[n - 1]]] = n - 1
    fill[T[n - 1]] = 1
    for c in range(sigma):
        if bstart[c + 1] == bstart[c]:
            continue
        begin = 0
        while fill[c] > begin:
            begin = _multiround_chain(T, SA, c, begin, fill[c], R, False, bstart, fill)
        assert fill[c] == Lc[c]
        if lms_off[c + 1] > lms_off[c] and c + 1 <= sigma - 1:
            _step(T, SA, lms_list, range(lms_off[c], lms_off[c + 1]), c + 1, sigma - 1, False, bstart, fill)
    fill = [0] * sigma
    for c in range(sigma - 1, -1, -1):
        if bstart[c + 1] == bstart[c]:
            continue
        begin = 0
        while fill[c] > begin:
            begin = _multiround_chain(T, SA, c, begin, fill[c], R, True, bstart, fill)
        assert fill[c] == Sc[c]
        if Lc[c] > 0 and c > 0:
            top = bstart[c] + Lc[c] - 1
            _step(T, SA, SA, range(top, bstart[c] - 1, -1), 0, c - 1, True, bstart, fill)


def induce_paircount(T, S, lms, SA, lms_list, lms_off, Lc, Sc, bstart, sigma=256):
    """PLANNED device formulation for large alphabets (NOTES_ROUND1.md, idea c).  With the
    (source bucket c -> destination bucket d) pair counts known from the classifier, the
    start of every region "entries induced from part X of bucket c into bucket d" is known
    a priori.  All LMS-sourced inductions (L pass) and all L-part-sourced inductions (S
    pass) then are ONE upfront, fully parallel partition each, and a bucket only needs its
    chain rounds: half the grid-synchronised steps of the per-bucket two-list scheme."""
    n = len(T)
    CL, CM, CS, CLS = {}, {}, {}, {}
    for i in range(1, n):                          # the classifier would count these pairs
        key = (T[i], T[i - 1])                     # (source bucket, destination bucket)
        if not S[i] and not S[i - 1]:
            CL[key] = CL.get(key, 0) + 1           # L(c) -> d   (L pass, d >= c)
        if lms[i]:
            CM[key] = CM.get(key, 0) + 1           # LMS(c) -> d (L pass, d > c)
        if S[i] and S[i - 1]:
            CS[key] = CS.get(key, 0) + 1           # S(c) -> d   (S pass, d <= c)
        if not S[i] and S[i - 1]:
            CLS[key] = CLS.get(key, 0) + 1         # L(c) -> d   (S pass, d < c)
    chars = [c for c in range(sigma) if bstart[c + 1] > bstart[c]]

    # ------------------------------------------------------------------ L pass
    baseL, baseM, chain0 = {}, {}, {}
    for d in chars:
        off = 1 if d == T[n - 1] else 0            # the seed n-1 sits first in its bucket
        for c in chars:
            if c >= d:
                break
            baseL[(c, d)] = off; off += CL.get((c, d), 0)
            baseM[(c, d)] = off; off += CM.get((c, d), 0)
        chain0[d] = off
        assert off + CL.get((d, d), 0) == Lc[d], (d, off, Lc[d])
    SA[bstart[T[n - 1]]] = n - 1
    cur = {}
    # upfront: every LMS suffix induces its predecessor (independent of all buckets' state)
    for c in chars:
        for k in range(lms_off[c], lms_off[c + 1]):
            s = lms_list[k]
            d = T[s - 1]
            SA[bstart[d] + baseM[(c, d)] + cur.get(("M", c, d), 0)] = s - 1
            cur[("M", c, d)] = cur.get(("M", c, d), 0) + 1
    # per bucket: chain rounds only
    for c in chars:
        begin, end = 0, chain0[c]
        nxt = chain0[c]
        while end > begin:
            for p in range(begin, end):
                s = SA[bstart[c] + p]
                if s == 0:
                    continue
                d = T[s - 1]
                if d == c:
                    SA[bstart[c] + nxt] = s - 1; nxt += 1
                elif d > c:
                    SA[bstart[d] + baseL[(c, d)] + cur.get(("L", c, d), 0)] = s - 1
                    cur[("L", c, d)] = cur.get(("L", c, d), 0) + 1
            begin, end = end, nxt
        assert nxt == Lc[c], (c, nxt, Lc[c])

    # ------------------------------------------------------------------ S pass (tail-relative slots)
    baseS, baseLS, chain0 = {}, {}, {}
    for d in chars:
        off = 0
        for c in reversed(chars):
            if c <= d:
                break
            baseS[(c, d)] = off; off += CS.get((c, d), 0)
            baseLS[(c, d)] = off; off += CLS.get((c, d), 0)
        chain0[d] = off
        assert off + CS.get((d, d), 0) == Sc[d], (d, off, Sc[d])
    cur = {}
    tail = lambda d, k: bstart[d + 1] - 1 - k
    # upfront: every L-part entry (right to left inside its bucket) induces an S-type predecessor
    for c in chars:
        for p in range(bstart[c] + Lc[c] - 1, bstart[c] - 1, -1):
            s = SA[p]
            if s == 0:
                continue
            d = T[s - 1]
            if d < c:
                SA[tail(d, baseLS[(c, d)] + cur.get(("LS", c, d), 0))] = s - 1
                cur[("LS", c, d)] = cur.get(("LS", c, d), 0) + 1
    for c in reversed(chars):
        begin, end = 0, chain0[c]
        nxt = chain0[c]
        while end > begin:
            for k in range(begin, end):
                s = SA[tail(c, k)]
                if s == 0:
                    continue
                d = T[s - 1]
                if d == c:
                    SA[tail(c, nxt)] = s - 1; nxt += 1
                elif d < c:
                    SA[tail(d, baseS[(c, d)] + cur.get(("S", c, d), 0))] = s - 1
                    cur[("S", c, d)] = cur.get(("S", c, d), 0) + 1
            begin, end = end, nxt
        assert nxt == Sc[c], (c, nxt, Sc[c])


def lms_equal(T, S, lms, a, b):
    """LMS-substring equality, src/table.rs:802-820 semantics."""
    n = len(T)
    i, j = a, b
    while i < n and j < n:
        if T[i] != T[j] or S[i] != S[j]:
            return False
        if i > a and (lms[i] or lms[j]):
            return True
        i += 1
        j += 1
    return False


def doubling_sa(R, kgram=None):
    """Suffix array of the reduced string by iterated rank/rename with
    group-start ranks; only non-singleton groups stay active.  The first
    refinement sorts by a k-gram of dense names (depth 1 -> k), later rounds by
    (group, rank[i+h]) pairs with h doubling -- the device's scheme
    (DESIGN.md 2.2); kgram=None picks k = floor(64 / bits(names+1)) like the device."""
    m = len(R)
    sa = sorted(range(m), key=lambda i: R[i])              # stable radix sort by name
    rank = [0] * m
    grp = [0] * m
    for p in range(m):
        grp[p] = p if (p == 0 or R[sa[p]] != R[sa[p - 1]]) else grp[p - 1]
        rank[sa[p]] = grp[p] + 1
    def active_of(keys):
        k = len(keys)
        return [q for q in range(k)
                if not ((q == 0 or keys[q] != keys[q - 1]) and (q == k - 1 or keys[q + 1] != keys[q]))]
    act = active_of([grp[p] for p in range(m)])
    apos = act[:]
    asuf = [sa[p] for p in act]
    agrp = [grp[p] for p in act]
    h = 1
    rounds = 0
    if kgram is None:
        bw = max(1, (max(R) + 1).bit_length()) if m else 1
        kgram = min(8, 64 // bw)
    while apos:
        rounds += 1
        first = rounds == 1 and kgram >= 2
        if first:
            keys = [tuple((R[asuf[k] + j] + 1) if asuf[k] + j < m else 0 for j in range(kgram)) for k in range(len(apos))]
        else:
            keys = [(agrp[k], rank[asuf[k] + h] if asuf[k] + h < m else 0) for k in range(len(apos))]
        order = sorted(range(len(apos)), key=lambda k: keys[k])
        skeys = [keys[k] for k in order]
        ssuf = [asuf[k] for k in order]
        ngrp = [0] * len(apos)
        for k in range(len(apos)):
            sa[apos[k]] = ssuf[k]
            ngrp[k] = apos[k] if (k == 0 or skeys[k] != skeys[k - 1]) else ngrp[k - 1]
        for k in range(len(apos)):
            rank[ssuf[k]] = ngrp[k] + 1
        keep = active_of(skeys)
        apos = [apos[k] for k in keep]
        asuf = [ssuf[k] for k in keep]
        agrp = [ngrp[k] for k in keep]
        h = kgram if first else h * 2
    return sa, rounds


def lms_direct_sort(T, lmspos, kc, max_rounds=None, stats=None):
    """Model of suffix_b200/csrc/lms_sort.cuh: LMS suffixes ordered by windows of kc
    characters (zero-padded past the end of the text), first sort fed in DESCENDING
    text position, every sort stable, truncated members (window runs past n) final
    where the stable sort leaves them and forming groups of their own.  Returns the
    LMS suffixes in suffix order, or None when max_rounds is exhausted."""
    n = len(T)
    codes = sorted(set(T))
    code = {c: k for k, c in enumerate(codes)}            # dense, order preserving, 0-based

    def window(p):                                        # p <= n
        return tuple(code[T[p + i]] if p + i < n else 0 for i in range(kc))

    items = list(reversed(lmspos))
    lst = sorted(items, key=window)                       # Python's sort is stable
    m = len(lst)

    def groups(seq_keys, seq_pos, span):
        """heads by the kernel's rule: key differs, or either neighbour truncated."""
        heads = []
        for j in range(len(seq_pos)):
            tr = lambda p: p + span > n
            heads.append(j == 0 or seq_keys[j] != seq_keys[j - 1] or tr(seq_pos[j - 1]) or tr(seq_pos[j]))
        return heads

    keys = [window(p) for p in lst]
    heads = groups(keys, lst, kc)
    # active = members of groups with more than one element: (slot, pos, grp)
    def active_of(heads, slots, poss):
        out = []
        g = None
        for j in range(len(poss)):
            if heads[j]:
                g = slots[j]
            tail = (j + 1 == len(poss)) or heads[j + 1]
            if not (heads[j] and tail):
                out.append((slots[j], poss[j], g))
        return out

    act = active_of(heads, list(range(m)), lst)
    h = kc
    rounds = 1
    while act:
        if max_rounds is not None and rounds >= max_rounds:
            return None
        rounds += 1
        slots = [a[0] for a in act]
        ks = [(a[2], window(a[1] + h)) for a in act]      # a[1] + h <= n by the truncation rule
        order = sorted(range(len(act)), key=lambda j: ks[j])
        poss = [act[j][1] for j in order]
        ks = [ks[j] for j in order]
        for sl, p in zip(slots, poss):
            lst[sl] = p
        heads = groups(ks, poss, h + kc)
        act = active_of(heads, slots, poss)
        h += kc
    if stats is not None:
        stats["direct_rounds"] = rounds
    return lst


def build_sa(T, sigma=256, stats=None, multiround=0, paircount=False, direct_kc=0):
    """Whole pipeline; T is a list/bytes of ints < sigma.  direct_kc > 0: the LMS suffixes are
    sorted directly by character windows (lms_direct_sort) instead of stage-1 induce + naming."""
    T = list(T)
    n = len(T)
    if n == 0:
        return []
    if n == 1:
        return [0]
    S, lms = classify(T)
    Lc, Sc, bstart = bucket_tables(T, S, sigma)
    lmspos = [i for i in range(n) if lms[i]]
    m = len(lmspos)
    lms_cnt = [0] * sigma
    for p in lmspos:
        lms_cnt[T[p]] += 1
    lms_off = [0] * (sigma + 1)
    for c in range(sigma):
        lms_off[c + 1] = lms_off[c] + lms_cnt[c]
    SA = [None] * n
    sorted_lms = None
    if m > 0 and direct_kc:
        sorted_lms = lms_direct_sort(T, lmspos, direct_kc, stats=stats)
    if sorted_lms is not None:
        pass
    elif m > 0:
        # stage 1: LMS grouped by first char (stable counting sort, text order)
        grouped = sorted(lmspos, key=lambda p: T[p])
        if paircount:
            induce_paircount(T, S, lms, SA, grouped, lms_off, Lc, Sc, bstart, sigma)
        elif multiround:
            induce_multiround(T, SA, grouped, lms_off, Lc, Sc, bstart, multiround, sigma)
        else:
            induce(T, SA, grouped, lms_off, Lc, Sc, bstart, sigma, stats)
        assert sorted(SA) == list(range(n))
        sorted_sub = [s for s in SA if lms[s]]
        # naming
        names = [0] * m
        name = -1
        for i, s in enumerate(sorted_sub):
            if i == 0 or not lms_equal(T, S, lms, s, sorted_sub[i - 1]):
                name += 1
            names[i] = name
        nnames = name + 1
        text_rank = {p: k for k, p in enumerate(lmspos)}
        R = [0] * m
        for i, s in enumerate(sorted_sub):
            R[text_rank[s]] = names[i]
        if nnames == m:
            sa_r = [0] * m
            for k in range(m):
                sa_r[R[k]] = k
            rounds = 0
        else:
            sa_r, rounds = doubling_sa(R)
        if stats is not None:
            stats.update(m=m, names=nnames, rounds=rounds)
        sorted_lms = [lmspos[k] for k in sa_r]
    else:
        sorted_lms = []
    SA = [None] * n
    if paircount:
        induce_paircount(T, S, lms, SA, sorted_lms, lms_off, Lc, Sc, bstart, sigma)
    elif multiround:
        induce_multiround(T, SA, sorted_lms, lms_off, Lc, Sc, bstart, multiround, sigma)
    else:
        induce(T, SA, sorted_lms, lms_off, Lc, Sc, bstart, sigma, stats)
    return SA
