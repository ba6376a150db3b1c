"""TEST INFRASTRUCTURE — a slow numpy/python model of the GPU setitem pipeline.

It restates, array by array, what the CUDA kernels in pixiu_b200/csrc/encode.cu
compute (SA -> LCP -> LPF M(s) -> P flags -> pair rule -> runs -> leftmost
pointer -> emission), so the *formulas* can be checked against the oracle on the
CPU before the kernels are run on a GPU.  O(n^2)-ish; small inputs only.
"""
import numpy as np


def build_text(docs):
    """text bytes with one 0 separator after each doc; returns text, is_sep, rec_start, rec_id, dist_end"""
    parts, starts = [], []
    pos = 0
    for d in docs:
        starts.append(pos)
        parts.append(np.frombuffer(d, dtype=np.uint8))
        parts.append(np.zeros(1, dtype=np.uint8))
        pos += len(d) + 1
    text = np.concatenate(parts)
    n = len(text)
    is_sep = np.zeros(n, dtype=bool)
    rec_id = np.zeros(n, dtype=np.int64)
    dist = np.zeros(n, dtype=np.int64)
    for r, (s, d) in enumerate(zip(starts, docs)):
        is_sep[s + len(d)] = True
        rec_id[s:s + len(d) + 1] = r
        dist[s:s + len(d) + 1] = np.arange(len(d), -1, -1)
    return text, is_sep, np.array(starts + [pos]), rec_id, dist


def suffix_array(text, is_sep):
    n = len(text)
    # symbol = byte+1, separator = unique lowest symbols ordered by position
    def key(i):
        out = []
        while True:
            if is_sep[i]:
                out.append((0, i))
                return out
            out.append((int(text[i]) + 1, 0))
            i += 1
    return np.array(sorted(range(n), key=key), dtype=np.int64)


def lcp_array(text, dist, sa):
    n = len(sa)
    lcp = np.zeros(n, dtype=np.int64)
    for j in range(1, n):
        a, b = sa[j - 1], sa[j]
        lim = min(dist[a], dist[b])
        h = 0
        while h < lim and text[a + h] == text[b + h]:
            h += 1
        lcp[j] = h
    return lcp


def encode_window(docs, first_new=0, strict251=False):
    """returns list of encoded records for docs[first_new:] (all docs form one chunk)"""
    text, is_sep, rec_start, rec_id, dist = build_text(docs)
    n = len(text)
    sa = suffix_array(text, is_sep)
    rank = np.zeros(n, dtype=np.int64)
    rank[sa] = np.arange(n)
    lcp = lcp_array(text, dist, sa)

    # K6: M(s) = max over the nearest smaller-position neighbours on both sides in SA order
    M = np.zeros(n, dtype=np.int64)
    for s in range(n):
        if is_sep[s]:
            continue
        r = rank[s]
        best = 0
        m = 1 << 30
        j = r
        while j > 0:                      # previous smaller value, tracking min lcp over (j, r]
            m = min(m, lcp[j])
            j -= 1
            if sa[j] < s:
                best = max(best, m)
                break
        m = 1 << 30
        j = r
        while j + 1 < n:
            j += 1
            m = min(m, lcp[j])
            if sa[j] < s:
                best = max(best, m)
                break
        M[s] = best
    reach = np.arange(n) + M

    # K7: P flags = image of reach (+ separators), then the escape-pair rule
    P = np.zeros(n + 1, dtype=bool)
    P[reach[~is_sep]] = True
    P[np.nonzero(is_sep)[0]] = True
    P = P[:n]
    C = ~P
    # position of the last non-251 byte at or before i (max-scan)
    idx = np.arange(n)
    last_non = np.maximum.accumulate(np.where(text != 251, idx, -1))
    C2 = C.copy()
    for i in range(n):
        if is_sep[i]:
            continue
        partner = -1
        if text[i] == 251:
            k = i - last_non[i]           # 1-based index inside the 251 run
            partner = i + 1 if k % 2 == 1 else i - 1
        elif i > 0 and text[i - 1] == 251 and (i - 1 - last_non[i - 1]) % 2 == 1:
            partner = i - 1               # second byte of a pair 251,x
        if partner >= 0 and not is_sep[partner] and not (C[i] and C[partner]):
            C2[i] = False
        if partner >= 0 and is_sep[partner]:
            C2[i] = C[i]
    C = C2
    Pf = ~C
    prevP = np.maximum.accumulate(np.where(Pf, idx, -1))
    nextP = np.minimum.accumulate(np.where(Pf, idx, n)[::-1])[::-1]

    # per-position output contribution
    contrib = np.zeros(n, dtype=np.int64)
    for i in range(n):
        if is_sep[i]:
            continue
        if Pf[i]:
            contrib[i] = 1
        else:
            rl = nextP[i] - prevP[i] - 1
            if rl <= 6:
                contrib[i] = 1
            elif i == nextP[i] - 1:
                contrib[i] = 8 if (rl > 255 or (rl == 251 and not strict251)) else 6
    off = np.concatenate([[0], np.cumsum(contrib)])

    out = []
    for d in range(first_new, len(docs)):
        a, b = rec_start[d], rec_start[d] + len(docs[d])
        enc = bytearray(off[b] - off[a])
        for i in range(a, b):
            o = off[i] - off[a]
            if contrib[i] == 1:
                enc[o] = text[i]
            elif contrib[i] > 1:
                rl = nextP[i] - prevP[i] - 1
                # s*(i) = min{s : reach(s) > i} within the record (binary search; reach is monotone)
                lo, hi = a, i + 1
                while lo < hi:
                    mid = (lo + hi) // 2
                    if reach[mid] > i:
                        hi = mid
                    else:
                        lo = mid + 1
                sstar = lo
                E = i + 1 - sstar
                assert E >= rl
                r = rank[sstar]
                lo_i = r
                while lcp[lo_i] >= E:     # lcp[0] = 0 stops it
                    lo_i -= 1
                hi_i = r
                while hi_i + 1 < n and lcp[hi_i + 1] >= E:
                    hi_i += 1
                left = sa[lo_i:hi_i + 1].min()
                assert left < sstar
                src = rec_id[left]
                to = left + E - rec_start[src]
                if contrib[i] == 8:
                    frm = to - rl
                    enc[o:o + 8] = bytes([251, 1, src & 255, src >> 8, to & 255, to >> 8, frm & 255, frm >> 8])
                else:
                    enc[o:o + 6] = bytes([251, rl, src & 255, src >> 8, to & 255, to >> 8])
        out.append(bytes(enc))
    return out
