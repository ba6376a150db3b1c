"""Deterministic synthetic workloads for the BASELINE.json configs.

All generators return *packed* batches ``(keys u8[], key_off i64[n+1], vals u8[],
val_off i64[n+1])`` — the layout the C ABI takes.  Randomness is a counter-based
splitmix64 so the bytes do not depend on the numpy version.

C1  gen_urls_kv      10k URL keys, ~100 B values
C2  gen_html_pages   HTML-like pages (bytes 33..126, <= 60,000 B), URL keys
C3  gen_nested       1 KB records, each = previous record with one byte swept
C4  gen_urls_kv with n = 10 M and ~200 B values
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(counter: np.ndarray, seed: int) -> np.ndarray:
    """counter-based splitmix64: u64[n] -> u64[n]"""
    with np.errstate(over="ignore"):
        z = counter.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed * 0xD1342543DE82EF95 & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


class Rng:
    """tiny stream wrapper over splitmix64"""

    def __init__(self, seed: int):
        self.seed = seed
        self.ctr = 0

    def u64(self, n: int) -> np.ndarray:
        r = splitmix64(np.arange(self.ctr, self.ctr + n, dtype=np.uint64), self.seed)
        self.ctr += n
        return r

    def below(self, n: int, hi) -> np.ndarray:
        return (self.u64(n) >> np.uint64(11)).astype(np.int64) % np.asarray(hi, dtype=np.int64)

    def uniform(self, n: int) -> np.ndarray:
        return (self.u64(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)

    def normal(self, n: int) -> np.ndarray:
        u1 = np.maximum(self.uniform(n), 1e-12)
        u2 = self.uniform(n)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)


def ragged_gather(pool: np.ndarray, starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """concatenate pool[starts[i]:starts[i]+lens[i]] for all i (vectorised)"""
    lens = lens.astype(np.int64)
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=pool.dtype)
    ends = np.cumsum(lens)
    idx = np.arange(total, dtype=np.int64) + np.repeat(starts.astype(np.int64) - (ends - lens), lens)
    return pool[idx]


def pack(items: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        np.cumsum([len(x) for x in items], out=off[1:])
    data = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if off[-1] else np.zeros(0, dtype=np.uint8)
    return data, off


def unpack(data: np.ndarray, off: np.ndarray) -> list[bytes]:
    b = data.tobytes()
    return [b[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def _vocab(rng: Rng, n_words: int) -> list[bytes]:
    lens = 3 + rng.below(n_words, 8)
    letters = rng.below(int(lens.sum()), 26)
    out, p = [], 0
    for L in lens.tolist():
        out.append(bytes((97 + letters[p:p + L]).astype(np.uint8)))
        p += L
    return out


_HOSTS = [b"news", b"sports", b"ent", b"finance", b"tech", b"auto"]


def _urls_packed(rng: Rng, n: int) -> tuple[np.ndarray, np.ndarray]:
    """unique keys http://{host}.qq.com/a/{yyyymmdd}/{6 or 8 digits}.htm, packed (vectorised: 10 M keys in seconds)"""
    host = rng.below(n, len(_HOSTS))
    day = rng.below(n, 28) + 1
    month = rng.below(n, 12) + 1
    # the serial is a permutation-like function of i => keys are unique
    if n <= 1000000:
        serial, nd = (np.arange(n, dtype=np.int64) * 7919 + 104729) % 1000000, 6
    else:
        serial, nd = np.arange(n, dtype=np.int64), 8
    hl = np.array([len(h) for h in _HOSTS], dtype=np.int64)[host]
    tail_len = len(b".qq.com/a/2016") + 4 + 1 + nd + len(b".htm")
    klen = 7 + hl + tail_len
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(klen, out=off[1:])
    out = np.empty(int(off[-1]), dtype=np.uint8)
    hpool, hoff = pack(_HOSTS)
    # fixed-width tail of every key
    tail = np.empty((n, tail_len), dtype=np.uint8)
    pre = np.frombuffer(b".qq.com/a/2016", dtype=np.uint8)
    tail[:, :len(pre)] = pre
    c = len(pre)
    tail[:, c] = 48 + month // 10
    tail[:, c + 1] = 48 + month % 10
    tail[:, c + 2] = 48 + day // 10
    tail[:, c + 3] = 48 + day % 10
    tail[:, c + 4] = ord("/")
    sv = serial.copy()
    for d in range(nd - 1, -1, -1):
        tail[:, c + 5 + d] = 48 + sv % 10
        sv //= 10
    tail[:, c + 5 + nd:] = np.frombuffer(b".htm", dtype=np.uint8)
    head = np.frombuffer(b"http://", dtype=np.uint8)
    base = off[:-1]
    for j in range(7):
        out[base + j] = head[j]
    for h in range(len(_HOSTS)):
        sel = np.nonzero(host == h)[0]
        if len(sel):
            for j in range(len(_HOSTS[h])):
                out[base[sel] + 7 + j] = _HOSTS[h][j]
    tb = base + 7 + hl
    for j in range(tail_len):
        out[tb + j] = tail[:, j]
    return out, off


def _urls(rng: Rng, n: int) -> list[bytes]:
    return unpack(*_urls_packed(rng, n))


def gen_urls_kv(n: int = 10000, seed: int = 1, val_words: int = 12):
    """C1 (and C4 with n=10M, val_words≈28): URL keys, `<title>` + words values."""
    rng = Rng(seed)
    vocab = _vocab(rng, 2000)
    kd, ko = _urls_packed(rng, n)
    vpool, voff = pack(vocab)
    vlen = np.diff(voff)
    head = np.frombuffer(b"<title>", dtype=np.uint8)
    # value i = "<title>" + words joined by single spaces
    sp = np.frombuffer(b" ", dtype=np.uint8)
    pool = np.concatenate([vpool, head, sp])
    h_at, s_at = len(vpool), len(vpool) + len(head)
    ctr0 = rng.ctr
    rng.ctr += n * val_words

    def rows(r0, r1):  # the values of records [r0, r1): the random stream is counter-based, so blocks are independent
        m = r1 - r0
        u = splitmix64(np.arange(ctr0 + r0 * val_words, ctr0 + r1 * val_words, dtype=np.uint64), seed)
        w = ((u >> np.uint64(11)).astype(np.int64) % len(vocab)).reshape(m, val_words)
        starts = np.empty((m, 2 * val_words), dtype=np.int64)
        lens = np.empty((m, 2 * val_words), dtype=np.int64)
        starts[:, 0], lens[:, 0] = h_at, len(head)
        starts[:, 1::2], lens[:, 1::2] = voff[w], vlen[w]
        starts[:, 2::2], lens[:, 2::2] = s_at, 1
        return ragged_gather(pool, starts.ravel(), lens.ravel()), lens.sum(axis=1)

    B = 250000
    blocks = [(a, min(n, a + B)) for a in range(0, n, B)]
    if len(blocks) > 1:  # (numpy releases the GIL inside its loops: the 10 M-record corpus of C4 builds on all cores)
        from concurrent.futures import ThreadPoolExecutor
        import os

        with ThreadPoolExecutor(max_workers=min(len(blocks), os.cpu_count() or 1)) as ex:
            parts = list(ex.map(lambda ab: rows(*ab), blocks))
    else:
        parts = [rows(*ab) for ab in blocks]
    vals = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, dtype=np.uint8)
    val_off = np.zeros(n + 1, dtype=np.int64)
    if parts:
        np.cumsum(np.concatenate([p[1] for p in parts]), out=val_off[1:])
    return kd, ko, vals, val_off


def gen_html_pages(n_pages: int = 10000, seed: int = 2, max_len: int = 60000, mean_len: float = 39600.0,
                   mix=(420, 455, 462, 580)):
    """C2: HTML-like ASCII pages with whitespace stripped (bytes 33..126), URL keys.

    A page belongs to one of 8 site templates.  It is a fixed header (a long run of
    the site's tag fragments), a body of paragraphs (vocabulary words, inline tags,
    `<ahref="URL">` links) and a fixed footer, truncated to min(max_len, lognormal).
    """
    rng = Rng(seed)
    vocab = _vocab(rng, 2000)
    n_sites, n_frag = 8, 300
    cls = _vocab(rng, 400)
    tags = [b"div", b"span", b"li", b"ul", b"p", b"a", b"td", b"tr", b"h2", b"h3", b"em", b"table"]
    pieces: list[bytes] = list(vocab)  # piece ids [0, 2000) = words
    frag_base = len(pieces)
    pick = rng.below(n_sites * n_frag * 4, 1 << 30)
    q = 0
    for s in range(n_sites):
        for f in range(n_frag):
            t = tags[pick[q] % len(tags)]
            c1 = cls[pick[q + 1] % len(cls)]
            c2 = cls[pick[q + 2] % len(cls)]
            k = pick[q + 3] % 4
            q += 4
            if k == 0:
                pieces.append(b'<%sclass="%s-%s"id="s%d_%d">' % (t, c1, c2, s, f))
            elif k == 1:
                pieces.append(b'</%s><%sclass="%s">' % (t, t, c1))
            elif k == 2:
                pieces.append(b'<%sstyle="margin:%dpx;color:#%06x">' % (t, f % 40, (pick[q - 1] * 2654435761) % (1 << 24)))
            else:
                pieces.append(b'</%s></div><!--%s-->' % (t, c2))
    link_base = len(pieces)
    n_links = 4000
    links = _urls(Rng(seed + 1000), n_links)
    pieces.extend(b'<ahref="%s">' % u for u in links)
    punct_base = len(pieces)
    pieces.extend([b",", b".", b"</p><p>", b"&nbsp;", b":", b"!", b"</a>", b"<br/>"])
    # "unique-ish" tokens: a large pool of random alphanumerics (ids, numbers, hashes)
    rand_base, n_rand = len(pieces), 200000
    an = np.frombuffer(b"0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ_-%", dtype=np.uint8)
    rl = 4 + rng.below(n_rand, 9)
    rbytes = an[rng.below(int(rl.sum()), len(an))]
    pool0, poff0 = pack(pieces)
    pool = np.concatenate([pool0, rbytes])
    poff = np.concatenate([poff0, poff0[-1] + np.cumsum(rl)])
    plen = np.diff(poff)
    w_word, w_frag, w_link, w_punct = mix

    keys = _urls(rng, n_pages)
    target = np.minimum(max_len, np.exp(np.log(mean_len) - 0.18 + 0.6 * rng.normal(n_pages))).astype(np.int64)
    target = np.maximum(target, 2000)
    site = rng.below(n_pages, n_sites)
    out_chunks, lens_out = [], np.zeros(n_pages, dtype=np.int64)
    hdr_n, ftr_n = 60, 25
    avg_piece = 7.2
    for i in range(n_pages):
        T = int(target[i])
        s = int(site[i])
        nb = int(T / avg_piece) + 64
        r = rng.below(nb, 1000)
        rid = rng.below(nb, 1 << 30)
        body = np.where(r < w_word, rid % 2000,                                # vocabulary word
               np.where(r < w_frag, frag_base + s * n_frag + rid % n_frag,     # site tag fragment
               np.where(r < w_link, link_base + rid % n_links,                 # link
               np.where(r < w_punct, punct_base + rid % 8,                     # punctuation / inline tag
                        rand_base + rid % n_rand))))                          # unique-ish token
        ids = np.concatenate([frag_base + s * n_frag + np.arange(hdr_n), body,
                              frag_base + s * n_frag + n_frag - ftr_n + np.arange(ftr_n)])
        page = ragged_gather(pool, poff[ids], plen[ids])[:T]
        out_chunks.append(page)
        lens_out[i] = len(page)
    vals = np.concatenate(out_chunks) if out_chunks else np.zeros(0, dtype=np.uint8)
    val_off = np.zeros(n_pages + 1, dtype=np.int64)
    np.cumsum(lens_out, out=val_off[1:])
    kd, ko = pack(keys)
    return kd, ko, vals, val_off


def gen_nested(n: int = 1000000, seed: int = 3, rec_len: int = 1000, sprinkle: bool = True):
    """C3: record r = record r-1 with the byte at (13*(r-1)) % rec_len replaced by a
    random uppercase letter (deep back-reference chains); every 16th record carries
    a short self-periodic run (overlapping self references)."""
    rng = Rng(seed)
    alpha = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz<>/=\"'.,;:-_()[]", dtype=np.uint8)[:41]
    base = alpha[rng.below(rec_len, len(alpha))]
    letters = (65 + rng.below(n + 1, 26)).astype(np.uint8)
    inv13 = pow(13, -1, rec_len)
    p = np.arange(rec_len, dtype=np.int64)
    first_mut = (1 + (p * inv13) % rec_len).astype(np.int32)          # first record index mutating position p
    vals = np.empty((n, rec_len), dtype=np.uint8)
    B = 16384

    def block(r0):
        r = np.arange(r0, min(n, r0 + B), dtype=np.int32)[:, None]
        last = r - ((r - first_mut[None, :]) % rec_len)     # latest mutation of p at or before r
        vals[r0:r0 + len(r)] = np.where(last >= 1, letters[np.maximum(last, 0)], base[None, :])

    if n > 4 * B:  # (numpy releases the GIL inside its loops)
        from concurrent.futures import ThreadPoolExecutor
        import os

        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            list(ex.map(block, range(0, n, B)))
    else:
        for r0 in range(0, n, B):
            block(r0)
    if sprinkle:
        rows = np.arange(0, n, 16)
        per = 1 + rng.below(len(rows), 9)
        ln = 20 + rng.below(len(rows), 100)
        at = rng.below(len(rows), rec_len - 130)
        for j, r in enumerate(rows.tolist()):
            a, P, L = int(at[j]), int(per[j]), int(ln[j])
            seg = vals[r, a:a + P].copy()
            vals[r, a:a + L] = np.tile(seg, L // P + 1)[:L]
    kd = np.empty((n, 12), dtype=np.uint8)
    kd[:, :3] = np.frombuffer(b"key", dtype=np.uint8)
    iv = np.arange(n, dtype=np.int64)
    for d in range(8, -1, -1):
        kd[:, 3 + d] = 48 + iv % 10
        iv //= 10
    kd, ko = kd.reshape(-1), np.arange(n + 1, dtype=np.int64) * 12
    val_off = np.arange(n + 1, dtype=np.int64) * rec_len
    return kd, ko, vals.reshape(-1), val_off
