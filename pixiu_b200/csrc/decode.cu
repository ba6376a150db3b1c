// getitem hot path: batched decode of PiXiu-encoded records.
//
// Replaces the recursive generator PXSGen::operator() (proj/PiXiuStr.h:110-198), which
// re-scans the referenced record from its first byte for every back reference and bubbles
// each byte through one coroutine per nesting level, by ONE data-flow kernel over a flat
// *decoded arena* (the records of every touched chunk, back to back, u32-addressed):
//   K10 k_decode_tiles one warp per 2 KiB decode tile (tile descriptors make every tile
//                      independently parsable), tiles handed out by ticket in arena order:
//                      251-dispatch of PiXiuStr.h:142-160 in parallel, literal bytes go straight
//                      to the arena, every reference token becomes a segment (destination,
//                      source, length; self-overlapping references keep their period) that is
//                      copied as soon as its source bytes are final.  Finality is tracked per
//                      byte in a bitmap (1 bit / byte), so the dependency depth is the nesting
//                      depth of the bytes, not of tiles or records.
//   K12 k_copy_records only when the caller's layout differs from the arena order.
// Traffic per decoded byte: the encoded byte or the source byte read once, the byte written
// once, 1/8 byte of bitmap - no per-byte pointer arrays.
// Waiting is deadlock-free: a source always precedes its destination in the arena and tickets
// follow arena order, so the tile owning a source is finished or held by a resident warp.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>

#include "index.h"
#include "store.h"

namespace pixiu {

constexpr int DEC_WARPS = 4;
constexpr uint32_t ENC_MAX = TILE + 16;           // encoded bytes a tile can span
constexpr uint32_t STG_PAD = 16;                  // free bytes in front of the staged range (reads just before it stay in bounds)
constexpr uint32_t STG_BYTES = STG_PAD + 16 + ENC_MAX + 16;  // pad + 16-byte alignment slack + range + token read-ahead
constexpr uint32_t STG_WORDS = (STG_BYTES + 31) / 32 * 8;  // whole 32-byte bitmap words
constexpr uint32_t BM_WORDS = 96;                 // bitmaps: three words per lane
// every reference segment but the first and the last of a tile puts >= 7 decoded bytes into the tile
constexpr uint32_t SEG_MAX = TILE / 7 + 4;
// copy pieces: every reference segment is cut into pieces of at most 32 bytes (a piece is ready when one 32-bit
// window of the "final" bitmap is all ones); periodic segments with a period >= 32 are also cut where they wrap.
// A tile keeps at most PIECE_MAX pieces (typical tiles have ~80); the segments beyond that go straight to k_resolve.
constexpr uint32_t PIECE_MAX = 256;
constexpr uint32_t PIECE_ROWS = (PIECE_MAX + 31) / 32;
constexpr uint32_t SPIN_LIMIT = 1u << 22;
constexpr uint32_t GIVEUP_SPINS = 64;   // polls without any progress after which a tile hands its open pieces to k_resolve
constexpr int RESOLVE_HOPS = 48;
static_assert(STG_BYTES <= BM_WORDS * 32 && TILE + 4 <= (BM_WORDS - 1) * 32, "bitmaps too small");

struct DecodeView {
    const uint8_t *enc;
    const uint64_t *enc_off;
    const uint32_t *enc_len, *dec_len, *first, *tile_base, *tile_desc;
    const uint32_t *arena_off;  // per record: offset of its decoded bytes in the arena
    uint8_t *arena;
    uint32_t *fin;              // per arena byte: 1 bit, set = the byte holds its final value
    uint32_t *gup;              // per arena byte: 1 bit, set = the byte was handed to k_resolve (its source is in ptr)
    uint32_t *ptr;              // per arena byte: source position, written only for handed-over bytes
};

struct ParseBits {                  // dead once the heads are listed: shares its storage with the pieces
    uint32_t b251[BM_WORDS];        // bit p: staged byte p is a 251 that can start a token
    uint32_t cst[BM_WORDS];         // bit p: that 251 surely starts a token (no 251 among the 7 bytes before it)
    uint32_t headb[BM_WORDS];       // bit p: a reference token starts at p
};
struct Pieces {
    // meta: u (12 bits) | (n - 1) << 12 (5 bits) | period << 17 (5 bits, 0 = plain copy) | phase << 22
    uint32_t src[PIECE_MAX];        // arena position of the first source byte (periodic: of the period's byte 0)
    uint32_t meta[PIECE_MAX];
};

struct alignas(16) WarpSmem {
    uint32_t stg[STG_WORDS];        // staged encoded bytes: byte e0 + k of the record sits at staged position soff + k
    union {
        ParseBits ps;
        Pieces pc;
    };
    uint16_t heads[SEG_MAX + 8];    // staged positions of the reference heads, ascending
    uint32_t startb[BM_WORDS];      // bit u: a segment starts at output byte u - mis (word-aligned "u" coordinates)
    uint32_t finw[BM_WORDS];        // bit u: output byte u - mis is a literal (final once phase 4 has stored it)
    uint16_t wprefix[BM_WORDS];     // segment starts before word w of startb
    // governors of the literal bytes: entry 0 governs the tile's start (the tail of a token that began in the
    // previous tile, or empty), entries 1.. follow the start bits.  lo16: end of the segment (u); hi16: literals
    // after it sit at staged position (u - mis) + this
    uint32_t seg_ed[SEG_MAX + 1];
    uint32_t pend[PIECE_ROWS];      // per row of 32 pieces: lanes whose piece is not copied yet
};

__device__ __forceinline__ uint32_t nib251(uint32_t w) {
    return ((__vcmpeq4(w, 0xFBFBFBFBu) & 0x08040201u) * 0x01010101u) >> 24;
}

__device__ __forceinline__ uint32_t nibnz(uint32_t w) {  // 4-bit mask of the nonzero bytes of w
    return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t bytemask4(uint32_t m) {  // 4-bit mask -> 0xFF per selected byte
    return (((m & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
}

// publish the new bits of a tile's bitmap (finw, u coordinates) in the arena's bitmap (tiles share words: OR);
// every lane looks after the global words lane, lane + 32 and lane + 64 of the tile's slice
__device__ __forceinline__ void flush_word(const uint32_t *finw, uint32_t *lm, uint32_t sh, uint32_t ngw, uint32_t j, uint32_t &pub) {
    if (j < ngw) {
        const uint32_t lo = finw[j], hi = j > 0 ? finw[j - 1] : 0u;
        const uint32_t v = sh ? ((lo << sh) | (hi >> (32 - sh))) : lo;
        const uint32_t nv = v & ~pub;
        if (nv) {
            atomicOr(&lm[j], nv);
            pub |= nv;
        }
    }
}
__device__ __forceinline__ void flush_final(const uint32_t *finw, uint32_t *lm, uint32_t sh, uint32_t ngw, uint32_t lane,
                                            uint32_t &pub0, uint32_t &pub1, uint32_t &pub2) {
    flush_word(finw, lm, sh, ngw, lane, pub0);
    flush_word(finw, lm, sh, ngw, lane + 32, pub1);
    flush_word(finw, lm, sh, ngw, lane + 64, pub2);
}

// K10: one warp per 2 KiB tile of decoded output, tiles taken by ticket in arena order.
//   1. stage the tile's encoded bytes in shared memory (16-byte loads)
//   2. bitmap of the 251s; a 251 with no 251 among the 7 bytes before it surely starts a token: one walk per such
//      cluster start classifies the tokens of its cluster (PiXiuStr.h:142-160 dispatch) and marks reference heads
//   3. heads in order -> segments: a scan of (decoded - encoded) token bytes gives every reference's output position,
//      because literals map 1:1
//   4. literal bytes: one aligned 32-bit store per output word straight from the staged bytes, then the literal bits
//      of the tile are published in the arena's "final" bitmap
//   5. reference segments: a segment is copied once its source range is final (sources always precede their
//      destination in the arena, and tickets are handed out in arena order, so every source belongs to a tile that
//      is finished or held by a resident warp: waiting cannot deadlock); copied ranges are published the same way
__global__ void __launch_bounds__(DEC_WARPS * 32, 8)
k_decode_tiles(DecodeView V, const uint32_t *__restrict__ work_tile, const uint32_t *__restrict__ work_rec,
               uint32_t n_work, uint32_t *__restrict__ ctr, uint32_t giveup_spins, uint32_t piece_cap,
               uint32_t sleep_after, uint32_t sleep_ns, unsigned long long *__restrict__ trace) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1;
    const uint32_t FULL = 0xffffffffu;
    uint32_t *err = ctr;
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(ctr + 1, 1u);
    w = __shfl_sync(FULL, w, 0);
    if (w >= n_work) return;
    if (trace && lane == 0) trace[4 * (size_t) w] = globaltimer_ns();
    const uint32_t gt = work_tile[w], g = work_rec[w];
    const uint32_t t = gt - V.tile_base[g];
    const uint32_t dl = V.dec_len[g], el = V.enc_len[g];
    const uint32_t t0 = t * TILE, t1 = min(dl, t0 + TILE), nbytes = t1 - t0;
    const uint32_t rec_base = V.arena_off[g], chunk_first = V.first[g];
    const uint8_t *encp = V.enc + V.enc_off[g];
    uint32_t desc = V.tile_desc[gt];
    const uint32_t e0 = desc & 0xffff;
    uint32_t skip = desc >> 16;
    const bool raw_first = skip == 0xFFFF;  // first enc byte is the 2nd half of an escape pair
    if (raw_first) skip = 0;
    uint32_t e_end = el;
    if (t1 < dl) {
        uint32_t d2 = V.tile_desc[gt + 1];
        e_end = d2 & 0xffff;
        uint32_t sk2 = d2 >> 16;
        if (sk2 != 0 && sk2 != 0xFFFF) e_end += (encp[e_end + 1] == 1) ? 8u : 6u;
    }
    const uint32_t ne = e_end - e0;
    const uint32_t mis = (uint32_t) ((uintptr_t) (V.arena + rec_base + t0) & 3);
    const uint32_t nu = nbytes + mis;
    const uint32_t B0 = rec_base + t0 - mis;  // arena position of u = 0
    uint32_t pub0 = 0, pub1 = 0, pub2 = 0;  // bits of the tile already published (global words lane, lane+32, lane+64)
    uint32_t *const lm = V.fin + (B0 >> 5);
    const uint32_t fsh = B0 & 31, ngw = (fsh + nu + 31) / 32;
    if (ne > ENC_MAX || e_end > el) {
        if (lane == 0) atomicExch(err, 4u);
        return;
    }
    // ---- 1. stage the encoded bytes (the compressed arena has slack past its end) ----
    const uint8_t *gsrc = encp + e0;
    const uint32_t a16 = (uint32_t) ((uintptr_t) gsrc & 15);
    const uint32_t soff = STG_PAD + a16, nstg = soff + ne;
    {
        const uint4 *g4 = reinterpret_cast<const uint4 *>(gsrc - a16);
        uint4 *s4 = reinterpret_cast<uint4 *>(S.stg);
        const uint32_t n16 = (a16 + ne + 8 + 15) >> 4;
        for (uint32_t j = lane; j < n16; j += 32) s4[1 + j] = g4[j];
        if (lane < 4) S.stg[lane] = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            S.ps.headb[lane + 32 * k] = 0;
            S.startb[lane + 32 * k] = 0;
            S.finw[lane + 32 * k] = 0;
        }
    }
    __syncwarp();
    const uint8_t *SB = reinterpret_cast<const uint8_t *>(S.stg);
    // ---- 2a. bitmap of the 251s (32 staged bytes per lane and step) ----
    {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(S.stg);
        const uint32_t lo = soff + (raw_first ? 1u : 0u);  // the raw first byte is a plain literal
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t wi = lane + 32 * k, base = wi * 32;
            uint32_t bits = 0;
            if (base < nstg) {
                const uint4 A = s4[2 * wi], B = s4[2 * wi + 1];
                bits = nib251(A.x) | (nib251(A.y) << 4) | (nib251(A.z) << 8) | (nib251(A.w) << 12) | (nib251(B.x) << 16) |
                       (nib251(B.y) << 20) | (nib251(B.z) << 24) | (nib251(B.w) << 28);
                if (base < lo) bits &= (lo - base >= 32) ? 0u : (0xFFFFFFFFu << (lo - base));
                if (base + 32 > nstg) bits &= 0xFFFFFFFFu >> (base + 32 - nstg);
            }
            S.ps.b251[wi] = bits;
        }
    }
    __syncwarp();
    // ---- 2b. cluster starts, then one walk per cluster ----
    uint32_t cs[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t wi = lane + 32 * k;
        const uint32_t b = S.ps.b251[wi], pb = wi ? S.ps.b251[wi - 1] : 0u;
        unsigned long long y = (((unsigned long long) b << 32) | pb) << 1;
        y |= y << 1;
        y |= y << 2;
        y |= y << 3;  // OR of the shifts 1..7
        cs[k] = b & ~(uint32_t) (y >> 32);
        S.ps.cst[wi] = cs[k];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint32_t st = cs[k];
        const uint32_t base = (lane + 32 * k) * 32;
        while (st) {
            uint32_t e = base + __ffs(st) - 1;
            st &= st - 1;
            while (true) {
                if (e + 1 >= nstg) break;  // first half of an escape pair cut by the tile boundary: a literal
                const uint32_t nx = SB[e + 1];
                uint32_t tl;
                if (nx == 0 || nx == 251 || nx == 2) {
                    tl = 2;
                } else if (nx == 1) {
                    tl = 8;
                } else if (nx > 6) {
                    tl = 6;
                } else {
                    atomicExch(err, 5u);  // 3..6: invalid (assert(false), PiXiuStr.h:193)
                    break;
                }
                if (tl > 2) atomicOr(&S.ps.headb[e >> 5], 1u << (e & 31));
                e += tl;
                if (e >= nstg) break;
                // the next 251 of this cluster lies within 7 bytes of the token's end (further ones start their own)
                const uint32_t wq = e >> 5, sh = e & 31;
                const uint32_t x = __funnelshift_r(S.ps.b251[wq], S.ps.b251[wq + 1], sh) & 0x7Fu;
                if (!x) break;
                const uint32_t cx = __funnelshift_r(S.ps.cst[wq], S.ps.cst[wq + 1], sh);
                const uint32_t f = __ffs(x) - 1;
                if ((cx >> f) & 1u) break;  // that one is a cluster start: its own walk handles it
                e += f;
            }
        }
    }
    __syncwarp();
    // ---- 2c. reference heads in ascending order (lane l owns bitmap words 3l .. 3l+2) ----
    uint32_t nheads;
    {
        uint32_t h[3] = {S.ps.headb[3 * lane], S.ps.headb[3 * lane + 1], S.ps.headb[3 * lane + 2]};
        const uint32_t cnt = __popc(h[0]) + __popc(h[1]) + __popc(h[2]);
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, inc, d);
            if ((int) lane >= d) inc += o;
        }
        nheads = __shfl_sync(FULL, inc, 31);
        if (nheads > SEG_MAX) {
            if (lane == 0) atomicExch(err, 6u);
            return;
        }
        uint32_t slot = inc - cnt;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            uint32_t hv = h[k];
            while (hv) {
                S.heads[slot++] = (uint16_t) ((3 * lane + k) * 32 + __ffs(hv) - 1);
                hv &= hv - 1;
            }
        }
        if (lane == 0) S.seg_ed[0] = mis | (soff << 16);  // entry 0: literals of the tile's start map 1:1 onto the staged bytes
    }
    __syncwarp();
    // ---- 3. heads -> segments.  D = sum over the references so far of (decoded - encoded) token bytes; a token
    //         at staged position p starts at output byte (p - soff) + D - skip ----
    uint32_t nseg = 0;  // governor entries 1..nseg
    uint32_t npiece = 0;
    bool full = false;  // the piece table is full: later segments are handed over
    int D = 0;
    for (uint32_t c0 = 0; c0 < nheads; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool head = c < nheads;
        uint32_t p = 0, idx = 0, from = 0, tl = 0;
        int d = 0;
        if (head) {
            p = S.heads[c];
            const uint32_t b1 = SB[p + 1];
            const bool big = b1 == 1;
            idx = SB[p + 2] | (SB[p + 3] << 8);
            const uint32_t to = SB[p + 4] | (SB[p + 5] << 8);
            from = big ? (uint32_t) (SB[p + 6] | (SB[p + 7] << 8)) : to - b1;
            if (to <= from || from > 0xFFFF) {
                atomicExch(err, 5u);
                from = to;
            }
            tl = to - from;
            d = (int) tl - (big ? 8 : 6);
        }
        int inc = d;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            int o = __shfl_up_sync(FULL, inc, dd);
            if ((int) lane >= dd) inc += o;
        }
        const int rel = (int) (p - soff) + D + (inc - d) - (int) skip;  // output position of the token's first byte
        const uint32_t k0 = rel < 0 ? (uint32_t) (-rel) : 0u;
        const uint32_t k1 = (int) tl + rel > (int) nbytes ? (uint32_t) max((int) nbytes - rel, 0) : tl;
        const bool emit = head && k0 < k1;
        const bool initial = emit && rel < 0;  // the tail of a token that began in the previous tile (only the first head)
        const uint32_t em = __ballot_sync(FULL, emit && !initial);
        uint32_t sbase = 0, per = 0, ks = k0, us = 0, len = 0, np = 0;
        if (emit) {
            const uint32_t sidx = initial ? 0u : 1u + nseg + __popc(em & lt);
            const uint32_t src_g = chunk_first + idx;
            if (src_g == g) {  // self reference (PiXiuStr.h:168-181): overlapping copies repeat with this period
                const uint32_t period = (uint32_t) ((int) t0 + rel) - from;
                sbase = rec_base + from;
                per = tl > period ? period : 0u;
                if ((int) t0 + rel <= (int) from) atomicExch(err, 3u);
            } else {
                if (src_g > g) atomicExch(err, 3u);
                sbase = (src_g > g ? 0u : V.arena_off[src_g]) + from;
            }
            len = k1 - k0;
            if (per) {
                ks = k0 % per;
                if (ks + len <= per) per = 0;  // this part does not wrap: a plain copy out of the first period
            }
            us = (uint32_t) (rel + (int) k0) + mis;
            if (sidx <= SEG_MAX) {
                S.seg_ed[sidx] = (us + len) | (((uint32_t) ((int) (soff + skip) - (D + inc)) & 0xFFFFu) << 16);
                if (!initial) atomicOr(&S.startb[us >> 5], 1u << (us & 31));
            }
            // pieces of this segment: a short plain segment is one piece; longer ones are cut at the 32-byte
            // boundaries of the SOURCE, so that a single word of the "final" bitmap decides whether a piece is ready
            if (per != 0 && per < 32) {
                np = (len + 31) >> 5;
            } else if (per == 0) {
                const uint32_t A = sbase + ks;
                np = len <= 32 ? 1u : ((A + len - 1) >> 5) - (A >> 5) + 1;
            } else {
                for (uint32_t pos = ks, rem = len; rem; np++) {
                    const uint32_t n = min(min(32u - ((sbase + pos) & 31u), per - pos), rem);
                    pos = pos + n == per ? 0u : pos + n;
                    rem -= n;
                }
            }
        }
        uint32_t pinc = np;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, pinc, dd);
            if ((int) lane >= dd) pinc += o;
        }
        uint32_t slot = npiece + pinc - np;
        const bool spill = np && (full || slot + np > piece_cap);  // (from the first spilling lane on, every later segment spills)
        const uint32_t sb = __ballot_sync(FULL, spill);
        if (spill) {
            // no room: the whole segment goes to k_resolve (source pointers + hand-over bits)
            const uint32_t B = B0 + us;
            for (uint32_t j = 0; j < len; j++) V.ptr[B + j] = sbase + (per ? (ks + j) % per : ks + j);
            for (uint32_t wj = B >> 5; wj <= (B + len - 1) >> 5; wj++) {
                uint32_t bits = 0xFFFFFFFFu;
                if (wj == B >> 5) bits <<= (B & 31);
                if (wj == (B + len - 1) >> 5) bits &= 0xFFFFFFFFu >> (31 - ((B + len - 1) & 31));
                atomicOr(V.gup + wj, bits);
            }
            atomicAdd(ctr + 2, 1u);
        }
        if (np && !spill) {
            if (per != 0 && per < 32) {
                for (uint32_t o = 0; o < len; o += 32, slot++) {
                    S.pc.src[slot] = sbase;
                    S.pc.meta[slot] = (us + o) | ((min(32u, len - o) - 1) << 12) | (per << 17) | (((ks + o) % per) << 22);
                }
            } else if (per == 0 && len <= 32) {
                S.pc.src[slot] = sbase + ks;
                S.pc.meta[slot] = us | ((len - 1) << 12);
            } else {
                const uint32_t wrap = per ? per : 0xFFFFFFFFu;
                for (uint32_t pos = ks, o = 0; o < len; slot++) {
                    const uint32_t n = min(min(32u - ((sbase + pos) & 31u), wrap - pos), len - o);
                    S.pc.src[slot] = sbase + pos;
                    S.pc.meta[slot] = (us + o) | ((n - 1) << 12);
                    pos = pos + n == wrap ? 0u : pos + n;
                    o += n;
                }
            }
        }
        nseg += __popc(em);
        npiece = sb ? __shfl_sync(FULL, slot, __ffs(sb) - 1) : npiece + __shfl_sync(FULL, pinc, 31);
        if (sb) full = true;
        D += __shfl_sync(FULL, inc, 31);
    }
    if ((int) ne + D < (int) (skip + nbytes) || nseg > SEG_MAX) {
        if (lane == 0) atomicExch(err, 6u);
        return;
    }
    __syncwarp();
    {  // segment starts before each bitmap word
        const uint32_t c0 = __popc(S.startb[3 * lane]), c1 = __popc(S.startb[3 * lane + 1]), c2 = __popc(S.startb[3 * lane + 2]);
        uint32_t inc = c0 + c1 + c2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(FULL, inc, d);
            if ((int) lane >= d) inc += o;
        }
        const uint32_t b = inc - (c0 + c1 + c2);
        S.wprefix[3 * lane] = (uint16_t) b;
        S.wprefix[3 * lane + 1] = (uint16_t) (b + c0);
        S.wprefix[3 * lane + 2] = (uint16_t) (b + c0 + c1);
        const uint32_t nrows0 = (npiece + 31) / 32;
        if (lane < PIECE_ROWS) S.pend[lane] = lane < nrows0 ? (32 * (lane + 1) <= npiece ? 0xFFFFFFFFu : (1u << (npiece - 32 * lane)) - 1) : 0u;
    }
    __syncwarp();
    // ---- 4. literal bytes, one output word per lane and step.  Bytes of a word that belong to reference segments
    //         are stored as zero (= "not final yet", see phase 5); only ZERO-valued literals need a bit in the bitmap ----
    uint8_t *dstu = V.arena + B0;
    uint32_t anyz = 0;
    for (uint32_t u0 = 0; u0 < nu; u0 += 128) {
        const uint32_t u = u0 + 4 * lane;
        uint32_t m = 0, wv = 0, hz = 0;
        if (u < nu) {
            const uint32_t sw = S.startb[u >> 5], sh = u & 31;
            const uint32_t gi = S.wprefix[u >> 5] + __popc(sw & ((1u << sh) - 1));  // governing entry of the word's first byte
            const uint32_t nibs = (sw >> sh) & 0xFu;
            const uint32_t ed = S.seg_ed[gi];
            const uint32_t lo = max(u, ed & 0xFFFFu);
            const uint32_t hi = min(nibs ? u + __ffs(nibs) - 1 : u + 4, nu);
            if (lo < hi) {
                m = ((1u << (hi - u)) - 1) & ~((1u << (lo - u)) - 1);
                const uint32_t bm = bytemask4(m);
                const uint32_t q = min((u - mis + (ed >> 16)) & 0xFFFFu, STG_BYTES - 8);
                wv = __funnelshift_r(S.stg[q >> 2], S.stg[(q >> 2) + 1], 8 * (q & 3)) & bm;
                hz = (wv - 0x01010101u) & ~wv & 0x80808080u & bm;  // (superset of) the zero-valued literal bytes
                if (u >= mis && u + 4 <= nu) {
                    *reinterpret_cast<uint32_t *>(dstu + u) = wv;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if ((m >> i) & 1u) dstu[u + i] = (uint8_t) (wv >> (8 * i));
                }
            }
        }
        if (__any_sync(FULL, hz != 0)) {  // rare: text has no zero bytes beyond the key terminator
            uint32_t v = (m & ~nibnz(wv)) << (4 * (lane & 7));
            v |= __shfl_xor_sync(FULL, v, 1);
            v |= __shfl_xor_sync(FULL, v, 2);
            v |= __shfl_xor_sync(FULL, v, 4);
            if ((lane & 7) == 0) S.finw[(u0 >> 5) + (lane >> 3)] = v;
            anyz = 1;
        }
    }
    __syncwarp();
    if (anyz) flush_final(S.finw, lm, fsh, ngw, lane, pub0, pub1, pub2);
    if (trace && lane == 0) trace[4 * (size_t) w + 1] = globaltimer_ns();
    // ---- 5. copy pieces, one per lane and row.  Finality travels IN BAND: the arena starts zeroed and a byte is only
    //         ever stored with its final value, so a nonzero byte is final the moment it is visible; a final byte
    //         whose value is zero is announced by its bit in V.fin (the bit IS the value, the byte itself needs no
    //         store).  A piece loads its source words (L2-coherent relaxed loads), copies the bytes that are final,
    //         remembers them in its done mask and retries the rest: no flag round trip, no fence, and the critical
    //         path is the nesting depth of BYTES (measured before: flag + MEMBAR.ALL.GPU per hop ~ 5 us).  The
    //         literal phase is over: the done masks take seg_ed's storage. ----
    static_assert(SEG_MAX + 1 >= PIECE_MAX, "done masks do not fit in seg_ed");
    uint32_t *const pdone = S.seg_ed;
    for (uint32_t k = lane; k < npiece; k += 32) pdone[k] = 0;
    __syncwarp();
    uint32_t remaining = npiece, spins = 0, sweep = 3;  // (the first sweep examines every piece)
    const uint32_t nrows = (npiece + 31) / 32;
    for (; remaining; sweep++) {
        uint32_t any = 0;
        for (uint32_t row = 0; row < nrows; row++) {
            const uint32_t pm = S.pend[row];
            if (!pm) continue;
            bool giveup = false, complete = false;
            uint32_t meta = 0, a = 0, dn = 0, avail = 0;
            if ((pm >> lane) & 1u) {
                const uint32_t slot = row * 32 + lane;
                meta = S.pc.meta[slot];
                a = S.pc.src[slot];
                const uint32_t per = (meta >> 17) & 31u, np = ((meta >> 12) & 31u) + 1, us = meta & 0xFFFu;
                const uint32_t n = per ? per : np;  // source bytes
                const uint32_t need = 0xFFFFFFFFu >> (32 - n);
                dn = per ? 0u : pdone[slot];
                uint32_t seen = 0;  // source bytes found final in this sweep
                // a waiting piece costs one load per sweep: the source byte behind its first open byte; the whole
                // window is examined when that byte has arrived (and every fourth sweep, for bytes out of order and
                // for zero-valued ones).  Sweeps are what a dependency hop costs, and 32 warps per SM share the issue
                // slots: the full examination of every piece in every sweep made a hop ~12 us.
                bool look = (sweep & 3u) == 3u;
                if (!look) {
                    // (first open byte on even sweeps, last one on odd sweeps: looking at the first byte only makes a
                    // byte wait for everything left of it in its piece, and with a source window that slides from
                    // record to record that running maximum chains every record to its predecessor)
                    const uint32_t open = need & ~dn;
                    const uint32_t pj = a + ((sweep & 1u) ? 31u - (uint32_t) __clz(open) : (uint32_t) __ffs(open) - 1u);
                    const uint32_t w0 = ld_poll_u32(reinterpret_cast<const uint32_t *>(V.arena + (pj & ~3u)));
                    look = ((w0 >> (8 * (pj & 3))) & 0xFFu) != 0;
                }
                if (look) {
                    const uint32_t sa = a & 3, nsw = (sa + n + 3) >> 2;  // aligned source words (<= 9)
                    const uint32_t *wp0 = reinterpret_cast<const uint32_t *>(V.arena + (a - sa));
                    uint32_t sw[10];
#pragma unroll
                    for (int q = 0; q < 5; q++) sw[q] = (uint32_t) q < nsw ? ld_poll_u32(wp0 + q) : 0u;
#pragma unroll
                    for (int q = 5; q < 10; q++) sw[q] = 0;
                    if (nsw > 5) {
#pragma unroll
                        for (int q = 5; q < 9; q++) sw[q] = (uint32_t) q < nsw ? ld_poll_u32(wp0 + q) : 0u;
                    }
                    // fast path (the common case): nothing copied yet and no zero among the source bytes, i.e. the whole
                    // piece is final -> straight word copy.  (w - 0x01010101) & ~w & 0x80808080 flags zero bytes; bytes
                    // outside [sa, sa + n) are forced nonzero.
                    bool fast = false;
                    if (!per && dn == 0) {
                        const uint32_t end = sa + n, L = (end - 1) >> 2, e8 = 8 * (end & 3);
                        const uint32_t mlo = (1u << (8 * sa)) - 1, mhi = e8 ? 0xFFFFFFFFu << e8 : 0u;
                        uint32_t z = 0;
#pragma unroll
                        for (int q = 0; q < 9; q++) {
                            uint32_t w = sw[q];
                            if (q == 0) w |= mlo;
                            if ((uint32_t) q == L) w |= mhi;
                            if ((uint32_t) q > L) w = 0xFFFFFFFFu;
                            z |= (w - 0x01010101u) & ~w & 0x80808080u;
                        }
                        fast = z == 0;
                    }
                    uint32_t zout = 0;  // copied bytes whose value is zero (destination coordinates of the piece)
                    if (fast) {
                        uint8_t *dst = dstu + us;
                        const uint32_t hbe = min((4u - (us & 3)) & 3u, n);  // head bytes up to a destination word boundary
                        const uint32_t x0 = __funnelshift_r(sw[0], sw[1], 8 * sa);
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            if ((uint32_t) i < hbe) dst[i] = (uint8_t) (x0 >> (8 * i));
                        const uint32_t rem = n - hbe, m = rem >> 2, tb = rem & 3, so = sa + hbe;
                        const uint32_t sh = 8 * (so & 3);
                        if (so >> 2) {  // (0 or 1)
#pragma unroll
                            for (int q = 0; q < 9; q++) sw[q] = sw[q + 1];
                        }
                        uint32_t *dw = reinterpret_cast<uint32_t *>(dst + hbe);
                        uint32_t xt = 0;
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            const uint32_t x = __funnelshift_r(sw[t], sw[t + 1], sh);
                            if ((uint32_t) t < m) dw[t] = x;
                            if ((uint32_t) t == m) xt = x;
                        }
#pragma unroll
                        for (int i = 0; i < 3; i++)
                            if ((uint32_t) i < tb) reinterpret_cast<uint8_t *>(dw + m)[i] = (uint8_t) (xt >> (8 * i));
                        avail = need;
                        seen = need;
                        dn = need;
                        complete = true;
                    } else {
                        unsigned long long nzm = 0;
#pragma unroll
                        for (int q = 0; q < 5; q++) nzm |= (unsigned long long) nibnz(sw[q]) << (4 * q);
                        if (nsw > 5) {
#pragma unroll
                            for (int q = 5; q < 9; q++) nzm |= (unsigned long long) nibnz(sw[q]) << (4 * q);
                        }
                        const uint32_t nz = (uint32_t) (nzm >> sa) & need;  // source bytes seen nonzero: final
                        uint32_t zf = 0;  // source bytes that are final zeros
                        const uint32_t zc = need & ~nz & ~dn;
                        if (zc) {
                            const uint32_t wi = a >> 5, bs = a & 31;
                            uint32_t bits = ld_relaxed_u32(V.fin + wi) >> bs;
                            if (bs + n > 32) bits |= ld_relaxed_u32(V.fin + wi + 1) << (32 - bs);
                            zf = zc & bits;
                        }
                        seen = nz | zf;
                        avail = seen & ~dn;
                        if (per) avail = avail == need ? need : 0u;  // short periods are copied in one go
                        if (avail && per) {
                            uint8_t *dst = dstu + us;
                            const uint32_t ph = meta >> 22;
                            for (uint32_t j = 0; j < np; j += 4) {  // four loads in flight
                                uint8_t bv[4];
#pragma unroll
                                for (int i = 0; i < 4; i++) bv[i] = __ldcg(V.arena + a + (ph + j + i) % per);
#pragma unroll
                                for (int i = 0; i < 4; i++)
                                    if (j + i < np) {
                                        dst[j + i] = bv[i];
                                        if (bv[i] == 0) zout |= 1u << (j + i);
                                    }
                            }
                            dn = 0xFFFFFFFFu >> (32 - np);
                            complete = true;
                        } else if (avail) {
                            // destination word t holds piece bytes [4t - da, 4t - da + 4); its source bytes straddle the
                            // aligned source words t + c and t + c + 1 (c = -1 when the source sits further left in its word)
                            const uint32_t da = us & 3;
                            const int delta = (int) sa - (int) da;
                            const uint32_t sh = 8u * (uint32_t) (delta & 3);
                            if (delta < 0) {
#pragma unroll
                                for (int q = 9; q > 0; q--) sw[q] = sw[q - 1];
                                sw[0] = 0;
                            }
                            uint32_t *dw = reinterpret_cast<uint32_t *>(dstu + us - da);
                            const unsigned long long am = (unsigned long long) avail << da;  // bytes to store, word coordinates
                            const uint32_t ndw = (da + n + 3) >> 2;
#pragma unroll
                            for (int t = 0; t < 9; t++) {
                                const uint32_t vm = (uint32_t) (am >> (4 * t)) & 0xFu;
                                if ((uint32_t) t < ndw && vm) {
                                    const uint32_t x = __funnelshift_r(sw[t], sw[t + 1], sh);
                                    if (vm == 0xFu) {
                                        dw[t] = x;
                                    } else {
#pragma unroll
                                        for (int i = 0; i < 4; i++)
                                            if ((vm >> i) & 1u) reinterpret_cast<uint8_t *>(dw + t)[i] = (uint8_t) (x >> (8 * i));
                                    }
                                }
                            }
                            zout = zf;
                            dn |= avail;
                            pdone[slot] = dn;
                            complete = dn == need;
                        }
                    }
                    if (zout) {  // announce the zero-valued bytes just copied
                        const uint32_t B = B0 + us, bs = B & 31;
                        red_relaxed_or_u32(V.fin + (B >> 5), zout << bs);
                        if (bs && (zout >> (32 - bs))) red_relaxed_or_u32(V.fin + (B >> 5) + 1, zout >> (32 - bs));
                    }
                }
                // a missing byte that its own tile handed to k_resolve will not become final in this kernel
                if (!avail && (spins & 7u) == 7u) {
                    const uint32_t miss = need & ~seen & ~dn;
                    const uint32_t wi = a >> 5, bs = a & 31;
                    giveup = spins >= giveup_spins || (ld_relaxed_u32(V.gup + wi) & (miss << bs)) != 0 ||
                             (bs + n > 32 && (ld_relaxed_u32(V.gup + wi + 1) & (miss >> (32 - bs))) != 0);
                }
            }
            uint32_t gb = __ballot_sync(FULL, giveup);
            const uint32_t gmask = gb;
            if (gb) {
                // hand the pieces over: per byte source pointers (self-overlapping references folded onto their first
                // period), one coalesced row of pointers per piece
                while (gb) {
                    const int r = __ffs(gb) - 1;
                    gb &= gb - 1;
                    const uint32_t rm = __shfl_sync(FULL, meta, r), ra = __shfl_sync(FULL, a, r);
                    const uint32_t n = ((rm >> 12) & 31u) + 1, per = (rm >> 17) & 31u;
                    if (lane < n) V.ptr[B0 + (rm & 0xFFFu) + lane] = ra + (per ? ((rm >> 22) + lane) % per : lane);
                }
                if (giveup) {
                    const uint32_t B = B0 + (meta & 0xFFFu), n = ((meta >> 12) & 31u) + 1;
                    const uint32_t bits = (0xFFFFFFFFu >> (32 - n)) & ~dn, bs = B & 31;  // the bytes still open
                    atomicOr(V.gup + (B >> 5), bits << bs);
                    if (bs + n > 32 && (bits >> (32 - bs))) atomicOr(V.gup + (B >> 5) + 1, bits >> (32 - bs));
                }
                if (lane == 0) atomicAdd(ctr + 2, (uint32_t) __popc(gmask));
                remaining -= __popc(gmask);
                any = 1;
            }
            const uint32_t cb = __ballot_sync(FULL, complete);
            if (__ballot_sync(FULL, avail != 0)) any = 1;
            if (cb | gmask) {
                if (lane == 0) S.pend[row] = pm & ~(cb | gmask);
                remaining -= __popc(cb);
                __syncwarp();
            }
        }
        __syncwarp();
        if (!any) {
            if (++spins > SPIN_LIMIT || ld_relaxed_u32(err) != 0) {
                if (lane == 0) atomicCAS(err, 0u, 8u);
                return;
            }
            if (spins > sleep_after) __nanosleep(spins > 256 ? 400 : sleep_ns);
        } else {
            spins = 0;
        }
    }
    if (trace && lane == 0) {  // measurement aid (PIXIU_DEC_TRACE_FILE): entry, end of the literal phase, done, sweeps
        trace[4 * (size_t) w + 2] = globaltimer_ns();
        trace[4 * (size_t) w + 3] = sweep;
    }
}

// K11 (only when K10 handed pieces over): a warp per 1 KiB of arena looks at its 32 words of the hand-over bitmap
// and, for every word with bits set, resolves the 32 bytes one per lane: each handed-over byte chases its pointer
// chain to a final byte.  Chains longer than RESOLVE_HOPS park their progress in the pointer array and the kernel is
// re-run: concurrent shortening makes the remaining rounds logarithmic in the nesting depth.  Final bytes are only
// read, never written, so no thread waits on another.
__global__ void __launch_bounds__(256)
k_resolve(uint32_t n, uint8_t *__restrict__ arena, uint32_t *__restrict__ ptr, const uint32_t *__restrict__ fin,
          uint32_t *__restrict__ gup, uint32_t *__restrict__ unfinished, uint32_t *__restrict__ err) {
    const uint32_t lane = lane_id();
    const uint32_t w0 = ((blockIdx.x * 256 + threadIdx.x) >> 5) * 32;  // first bitmap word of the warp
    const uint32_t nwords = (n + 31) >> 5;
    bool pending = false;
    const uint32_t mine = w0 + lane < nwords ? gup[w0 + lane] : 0u;
    uint32_t todo = __ballot_sync(0xffffffffu, mine != 0);
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t bits = __shfl_sync(0xffffffffu, mine, k);
        const uint32_t i = (w0 + k) * 32 + lane;
        bool done = true;
        if ((bits >> lane) & 1u) {
            uint32_t p = ptr[i];
            done = false;
            for (int h = 0; h < RESOLVE_HOPS; h++) {
                if (p >= i) {  // sources always precede their byte in the arena: corrupt input
                    atomicExch(err, 7u);
                    done = true;
                    p = i;
                    break;
                }
                const uint8_t bv = __ldcg(arena + p);  // final = nonzero, or a zero announced in the bitmap
                if (bv != 0 || ((fin[p >> 5] >> (p & 31)) & 1u)) {
                    arena[i] = bv;
                    done = true;
                    break;
                }
                p = ptr[p];
            }
            if (p != i) ptr[i] = p;  // the final origin, or an ancestor further up the chain
            pending |= !done;
        }
        // bytes resolved for good leave the bitmap (later rounds skip them; chains through them end at ptr -> final)
        const uint32_t still = __ballot_sync(0xffffffffu, !done);
        if (lane == 0 && still != bits) gup[w0 + k] = still;
    }
    if (__syncthreads_or(pending) && threadIdx.x == 0) atomicAdd(unfinished, 1u);
}

// K12: arena -> caller layout, one warp per requested record
__global__ void __launch_bounds__(256)
k_copy_records(uint32_t n, const uint32_t *__restrict__ recs, const uint64_t *__restrict__ out_off,
               const uint32_t *__restrict__ arena_off, const uint32_t *__restrict__ dec_len,
               const uint8_t *__restrict__ arena, uint8_t *__restrict__ out) {
    uint32_t w = (blockIdx.x * 256 + threadIdx.x) >> 5;
    if (w >= n) return;
    uint32_t g = recs[w];
    const uint8_t *src = arena + arena_off[g];
    uint8_t *dst = out + out_off[w];
    uint32_t len = dec_len[g];
    for (uint32_t j = lane_id(); j < len; j += 32) dst[j] = src[j];
}


// Work list of a decode call, built on the device.  The host describes the touched chunks (a few hundred ranges at
// most); everything per record or per tile is derived here from the record tables already resident in HBM.
struct DecRange {
    uint32_t first, last;      // records [first, last] of one chunk form a contiguous part of the arena
    uint32_t tile_lo, ntiles;  // their decode tiles (global tile ids are consecutive inside a chunk)
    uint32_t rec_cum, tile_cum;  // records / tiles of the ranges before this one
    uint32_t arena_base, pad;
};

// arena offset of every record of the ranges
__global__ void __launch_bounds__(256)
k_dec_aoff(uint32_t n_rec, uint32_t n_ranges, const DecRange *__restrict__ R, const uint64_t *__restrict__ dec_prefix,
           uint32_t *__restrict__ aoff) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rec) return;
    uint32_t lo = 0, hi = n_ranges;  // last range with rec_cum <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (R[mid].rec_cum <= i) lo = mid;
        else hi = mid;
    }
    const DecRange r = R[lo];
    const uint32_t g = r.first + (i - r.rec_cum);
    aoff[g] = r.arena_base + (uint32_t) (dec_prefix[g] - dec_prefix[r.first]);
}

// (tile, record) of every work item.  Tiles of one chunk keep their order (the data-flow argument needs it); chunks
// are interleaved round-robin, so that as many dependency chains as there are chunks advance side by side: the
// k-th tile of range c goes to position sum_c' min(ntiles_c', k) + #{c' < c : ntiles_c' > k}.
__global__ void __launch_bounds__(256)
k_dec_work(uint32_t n_work, uint32_t n_ranges, const DecRange *__restrict__ R, const uint32_t *__restrict__ tile_base,
           uint32_t *__restrict__ work_tile, uint32_t *__restrict__ work_rec) {
    const uint32_t j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n_work) return;
    uint32_t lo = 0, hi = n_ranges;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (R[mid].tile_cum <= j) lo = mid;
        else hi = mid;
    }
    const DecRange r = R[lo];
    const uint32_t k = j - r.tile_cum, gt = r.tile_lo + k;
    uint32_t pos = 0;
    for (uint32_t c = 0; c < n_ranges; c++) {
        const uint32_t nt = R[c].ntiles;
        pos += min(nt, k) + ((c < lo && nt > k) ? 1u : 0u);
    }
    uint32_t a = r.first, b = r.last + 1;  // last record with tile_base <= gt
    while (b - a > 1) {
        const uint32_t mid = (a + b) >> 1;
        if (tile_base[mid] <= gt) a = mid;
        else b = mid;
    }
    work_tile[pos] = gt;
    work_rec[pos] = a;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
// The arena of one decode pass is addressed with 32 bits.  A request whose touched chunks decode to more than
// DEC_ARENA_LIMIT bytes is split by chunk (chunks are self-contained) into passes that each stay below it.
void Store::decode_records(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off) {
    if (recs.empty()) return;
    if (mirror_from < n_records()) flush_mirrors();  // the running sum below is built from the host mirrors
    const char *lim_env = getenv("PIXIU_DEC_ARENA_LIMIT");  // test knob
    const uint64_t LIMIT = lim_env ? (uint64_t) atoll(lim_env) : (7ull << 29);  // 3.5 GiB
    // arena bytes per touched chunk: records [first, max requested]
    std::map<uint32_t, uint32_t> cmax;
    uint32_t last_f = 0xFFFFFFFFu, *last_max = nullptr;
    for (uint32_t g : recs) {
        const uint32_t f = h_first[g];
        if (f != last_f) {
            auto it = cmax.find(f);
            if (it == cmax.end()) it = cmax.emplace(f, g).first;
            last_f = f;
            last_max = &it->second;
        }
        if (g > *last_max) *last_max = g;
    }
    if (h_dec_prefix.empty()) h_dec_prefix.push_back(0);
    while (h_dec_prefix.size() < n_records() + 1) h_dec_prefix.push_back(h_dec_prefix.back() + h_dec_len[h_dec_prefix.size() - 1]);
    uint64_t total = 0;
    for (auto &cm : cmax) total += h_dec_prefix[cm.second + 1] - h_dec_prefix[cm.first];
    if (total <= LIMIT) {
        decode_pass(recs, d_out, out_off, &cmax);
        return;
    }
    std::map<uint32_t, uint32_t> group_of;  // chunk first record -> pass
    uint32_t ng = 0;
    uint64_t acc = 0;
    for (auto &cm : cmax) {
        const uint64_t b = h_dec_prefix[cm.second + 1] - h_dec_prefix[cm.first];
        if (acc && acc + b > LIMIT) {
            ng++;
            acc = 0;
        }
        group_of[cm.first] = ng;
        acc += b;
    }
    ng++;
    std::vector<std::vector<uint32_t>> grecs(ng);
    std::vector<std::vector<uint64_t>> goffs(ng);
    for (size_t i = 0; i < recs.size(); i++) {
        const uint32_t gi = group_of[h_first[recs[i]]];
        grecs[gi].push_back(recs[i]);
        goffs[gi].push_back(out_off[i]);
    }
    double ms_sum = 0;
    for (uint32_t gi = 0; gi < ng; gi++) {
        goffs[gi].push_back(0);  // (only the first recs.size() entries are read; keeps the "n + 1" shape)
        decode_pass(grecs[gi], d_out, goffs[gi], nullptr);
        ms_sum += last_get_ms;
    }
    last_get_ms = ms_sum;
}

void Store::decode_pass(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off,
                        const std::map<uint32_t, uint32_t> *known_max) {
    if (recs.empty()) return;
    const auto t_h0 = std::chrono::steady_clock::now();
    // per touched chunk: records [first, max requested] form the arena
    std::map<uint32_t, uint32_t> own_max;  // chunk first record -> max requested record
    if (!known_max) {
        std::map<uint32_t, uint32_t> &chunk_max = own_max;
        uint32_t last_f = 0xFFFFFFFFu, *last_max = nullptr;
        for (uint32_t g : recs) {
            const uint32_t f = h_first[g];
            if (f != last_f) {
                auto it = chunk_max.find(f);
                if (it == chunk_max.end()) it = chunk_max.emplace(f, g).first;
                last_f = f;
                last_max = &it->second;
            }
            if (g > *last_max) *last_max = g;
        }
    }
    const std::map<uint32_t, uint32_t> &chunk_max = known_max ? *known_max : own_max;
    const size_t NR = n_records();
    // running sum of the decoded lengths (host and device copies grow with the store)
    if (h_dec_prefix.empty()) h_dec_prefix.push_back(0);
    while (h_dec_prefix.size() < NR + 1) h_dec_prefix.push_back(h_dec_prefix.back() + h_dec_len[h_dec_prefix.size() - 1]);
    if (dec_prefix_synced < NR + 1) {
        d_dec_prefix.reserve_keep(NR + 1, dec_prefix_synced, st);
        PX_CUDA(cudaMemcpyAsync(d_dec_prefix.p + dec_prefix_synced, h_dec_prefix.data() + dec_prefix_synced,
                                (NR + 1 - dec_prefix_synced) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        dec_prefix_synced = NR + 1;
    }
    std::vector<DecRange> ranges;
    ranges.reserve(chunk_max.size());
    uint64_t arena_bytes = 0, n_work = 0, n_rec = 0;
    double alg_bytes = 0;
    for (auto &cm : chunk_max) {
        const uint32_t f = cm.first, m = cm.second;
        const uint32_t tile_lo = h_tile_base[f], tile_hi = h_tile_base[m] + div_up<uint32_t>(h_dec_len[m], TILE);
        if (arena_bytes >= 0xFFFFFF00ull) break;
        ranges.push_back(DecRange{f, m, tile_lo, tile_hi - tile_lo, (uint32_t) n_rec, (uint32_t) n_work, (uint32_t) arena_bytes, 0u});
        arena_bytes += h_dec_prefix[m + 1] - h_dec_prefix[f];
        n_work += tile_hi - tile_lo;
        n_rec += m - f + 1;
        alg_bytes += (double) (h_enc_off[m] + h_enc_len[m] - h_enc_off[f]) + (double) (h_dec_prefix[m + 1] - h_dec_prefix[f]);
    }
    if (arena_bytes >= 0xFFFFFF00ull) throw std::runtime_error("decode: arena of one call exceeds 4 GiB; split the batch");
    // direct mode: the request is exactly the arena order and the caller's layout is packed the same way
    bool direct = ((uintptr_t) d_out & 127) == 0 && recs.size() == n_rec;
    if (direct) {
        size_t ri = 0;
        for (const DecRange &r : ranges) {
            const uint64_t base = (uint64_t) r.arena_base - h_dec_prefix[r.first];
            for (uint32_t g = r.first; g <= r.last && direct; g++, ri++)
                direct = recs[ri] == g && out_off[ri] == base + h_dec_prefix[g];
            if (!direct) break;
        }
    }
    if (!direct) dec_scratch.reserve_discard(arena_bytes + 256);  // same packed layout, private buffer
    uint8_t *arena = direct ? d_out : dec_scratch.p;
    const size_t bm_words = arena_bytes / 32 + 64;
    dec_flags.reserve_discard(2 * bm_words);                  // zero-byte bitmap, then the "handed over" bitmap
    dec_ptr.reserve_discard(arena_bytes + 64);                // source pointers (touched only for handed-over bytes)
    dec_aoff.reserve_discard(NR + 1);
    dec_work.reserve_discard(2 * n_work + 2);
    dec_ranges.reserve_discard(ranges.size() * sizeof(DecRange) / sizeof(uint32_t) + 8);
    DecRange *d_ranges = reinterpret_cast<DecRange *>(dec_ranges.p);
    PX_CUDA(cudaMemcpyAsync(d_ranges, ranges.data(), ranges.size() * sizeof(DecRange), cudaMemcpyHostToDevice, st));
    k_dec_aoff<<<(unsigned) div_up<uint64_t>(n_rec, 256), 256, 0, st>>>((uint32_t) n_rec, (uint32_t) ranges.size(), d_ranges,
                                                                        d_dec_prefix.p, dec_aoff.p);
    k_dec_work<<<(unsigned) div_up<uint64_t>(n_work, 256), 256, 0, st>>>((uint32_t) n_work, (uint32_t) ranges.size(), d_ranges,
                                                                         d_tile_base.p, dec_work.p, dec_work.p + n_work);
    launches += 2;
    dec_ctr.reserve_discard(64);
    PX_CUDA(cudaMemsetAsync(dec_ctr.p, 0, 64 * sizeof(uint32_t), st));  // [0] error, [1] tile ticket, [2] pieces handed over
    PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, 2 * bm_words * sizeof(uint32_t), st));
    DecodeView V{d_enc.ptr(), d_enc_off.p, d_enc_len.p, d_dec_len.p, d_first.p, d_tile_base.p, d_tile_desc.p,
                 dec_aoff.p, arena, dec_flags.p, dec_flags.p + bm_words, dec_ptr.p};
    const size_t smem = sizeof(WarpSmem) * DEC_WARPS;
    // (per device and cheap: no process-wide "already done" flag, a process may drive several GPUs)
    PX_CUDA(cudaFuncSetAttribute(k_decode_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const char *gs = getenv("PIXIU_GIVEUP_SPINS");  // tuning knob
    const uint32_t giveup_spins = gs ? (uint32_t) atoi(gs) : GIVEUP_SPINS;
    const char *pcs = getenv("PIXIU_PIECE_CAP");    // test knob: forces the spill path of the piece table
    const uint32_t piece_cap = pcs ? std::min<uint32_t>((uint32_t) atoi(pcs), PIECE_MAX) : PIECE_MAX;
    const char *sa_ = getenv("PIXIU_SLEEP_AFTER"), *sn_ = getenv("PIXIU_SLEEP_NS");  // tuning knobs of the poll back-off
    const uint32_t sleep_after = sa_ ? (uint32_t) atoi(sa_) : 16u, sleep_ns = sn_ ? (uint32_t) atoi(sn_) : 64u;
    if (getenv("PIXIU_TRACE"))
        fprintf(stderr, "[decode] host work list %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_h0).count());
    unsigned long long *d_trace = nullptr;
    const char *trace_file = getenv("PIXIU_DEC_TRACE_FILE");  // measurement aid: per-tile entry / parsed / done times + sweeps
    if (trace_file) {
        PX_CUDA(cudaMalloc(&d_trace, 4 * n_work * sizeof(unsigned long long)));
        PX_CUDA(cudaMemsetAsync(d_trace, 0, 4 * n_work * sizeof(unsigned long long), st));
    }
    PX_CUDA(cudaEventRecord(ev0, st));
    prof.begin(PC_DECODE, st);
    PX_CUDA(cudaMemsetAsync(arena, 0, arena_bytes, st));  // zero = "not final yet" (k_decode_tiles, phase 5): timed with the kernel
    k_decode_tiles<<<(unsigned) div_up<uint64_t>(n_work, DEC_WARPS), DEC_WARPS * 32, smem, st>>>(
        V, dec_work.p, dec_work.p + n_work, (uint32_t) n_work, dec_ctr.p, giveup_spins, piece_cap, sleep_after, sleep_ns, d_trace);
    int nl = 1;
    uint32_t h_ctr[4] = {0, 0, 0, 0};
    PX_CUDA(cudaMemcpyAsync(h_ctr, dec_ctr.p, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    last_handed_over = h_ctr[2];
    if (d_trace) {  // file: u64 n_work, u32 record[n_work] (ticket order), u64 {entry, parsed, done (ns), sweeps}[n_work]
        std::vector<unsigned long long> ht(4 * n_work);
        std::vector<uint32_t> hr(n_work);
        PX_CUDA(cudaMemcpy(ht.data(), d_trace, ht.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        PX_CUDA(cudaMemcpy(hr.data(), dec_work.p + n_work, n_work * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trace_file, "wb")) {
            const unsigned long long nw = n_work;
            fwrite(&nw, 8, 1, f);
            fwrite(hr.data(), 4, n_work, f);
            fwrite(ht.data(), 8, ht.size(), f);
            fclose(f);
        }
        cudaFree(d_trace);
    }
    if (getenv("PIXIU_TRACE")) fprintf(stderr, "[decode] %llu tiles, %u pieces handed to k_resolve\n", (unsigned long long) n_work, h_ctr[2]);
    if (h_ctr[0] == 0 && h_ctr[2] != 0) {
        // deep reference chains: the pieces the data-flow pass handed over are resolved by pointer chasing
        for (int round = 0; round < 40; round++) {
            k_resolve<<<(unsigned) div_up<uint64_t>(div_up<uint64_t>(arena_bytes, 32), 256), 256, 0, st>>>(
                (uint32_t) arena_bytes, arena, dec_ptr.p, V.fin, V.gup, dec_ctr.p + 3 + round, dec_ctr.p);
            nl++;
            PX_CUDA(cudaMemcpyAsync(h_ctr, dec_ctr.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PX_CUDA(cudaMemcpyAsync(h_ctr + 3, dec_ctr.p + 3 + round, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PX_CUDA(cudaStreamSynchronize(st));
            if (h_ctr[0] || h_ctr[3] == 0) break;
        }
        if (h_ctr[0] == 0 && h_ctr[3] != 0) h_ctr[0] = 9;  // chains did not resolve
    }
    if (!direct) {
        // requested records -> caller layout
        DevBuf<uint32_t> &d_recs = dec_reqs;
        d_recs.reserve_discard(recs.size() + 1);
        dec_loc.reserve_discard(recs.size() + 1);
        PX_CUDA(cudaMemcpyAsync(d_recs.p, recs.data(), recs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PX_CUDA(cudaMemcpyAsync(dec_loc.p, out_off.data(), recs.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        k_copy_records<<<(unsigned) div_up<uint64_t>((uint64_t) recs.size() * 32, 256), 256, 0, st>>>(
            (uint32_t) recs.size(), d_recs.p, dec_loc.p, dec_aoff.p, d_dec_len.p, arena, d_out);
        nl++;
    }
    PX_LAUNCH_CHECK();
    prof.end(st, alg_bytes, nl);
    launches += nl;
    PX_CUDA(cudaEventRecord(ev1, st));
    PX_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_get_ms = ms;
    prof.collect();
    if (h_ctr[0]) throw std::runtime_error("decode: kernel reported error " + std::to_string(h_ctr[0]));
}

// Host-side token walk of one encoded record: validates it, returns its decoded length and
// appends its tile descriptors (import path only; setitem builds them on the GPU).
static int64_t parse_record(const uint8_t *e, uint32_t n, uint32_t self_idx, const std::vector<uint32_t> &dec_len_of,
                            std::vector<uint32_t> &desc) {
    uint32_t d = 0, next_tile = 0;
    auto mark = [&](uint32_t tok_e, uint32_t tok_d, uint32_t tok_len, bool pair_second) {
        // tiles whose first byte falls inside [tok_d, tok_d + tok_len)
        while ((uint64_t) next_tile * TILE < (uint64_t) tok_d + tok_len) {
            uint32_t skip = next_tile * TILE - tok_d;
            desc.push_back(tok_e | ((pair_second ? 0xFFFFu : skip) << 16));
            next_tile++;
        }
    };
    for (uint32_t i = 0; i < n;) {
        uint8_t b = e[i];
        if (b != 251) {
            mark(i, d, 1, false);
            d++;
            i++;
            continue;
        }
        if (i + 1 >= n) return -1;
        uint8_t nx = e[i + 1];
        if (nx == 0 || nx == 251 || nx == 2) {
            mark(i, d, 1, false);
            mark(i + 1, d + 1, 1, true);
            d += 2;
            i += 2;
            continue;
        }
        uint32_t idx, to, from, adv;
        if (nx == 1) {
            if (i + 8 > n) return -1;
            idx = e[i + 2] | (e[i + 3] << 8);
            to = e[i + 4] | (e[i + 5] << 8);
            from = e[i + 6] | (e[i + 7] << 8);
            adv = 8;
        } else if (nx > 6) {
            if (i + 6 > n) return -1;
            idx = e[i + 2] | (e[i + 3] << 8);
            to = e[i + 4] | (e[i + 5] << 8);
            if (to < nx) return -1;
            from = to - nx;
            adv = 6;
        } else {
            return -1;
        }
        if (to <= from) return -1;
        if (idx == self_idx) {
            if (from >= d) return -1;
        } else if (idx > self_idx || to > dec_len_of[idx]) {
            return -1;
        }
        mark(i, d, to - from, false);
        d += to - from;
        i += adv;
        if (d > MAX_DOC) return -1;
    }
    return d;
}

int64_t Store::import_chunk(int64_t n, const uint8_t *enc, const int64_t *enc_off) {
    if (n <= 0 || n > (int64_t) MAX_CHUNK_RECS) return PIXIU_EINVAL;
    if (win_open) close_window();
    std::vector<uint32_t> dl(n), descs;
    std::vector<uint32_t> tbase(n);
    uint64_t tiles = n_tiles;
    for (int64_t r = 0; r < n; r++) {
        int64_t len = enc_off[r + 1] - enc_off[r];
        if (len <= 0 || len > (int64_t) MAX_DOC) return PIXIU_ECORRUPT;
        tbase[r] = (uint32_t) tiles;
        size_t before = descs.size();
        int64_t d = parse_record(enc + enc_off[r], (uint32_t) len, (uint32_t) r, dl, descs);
        if (d <= 0) return PIXIU_ECORRUPT;
        dl[r] = (uint32_t) d;
        if (descs.size() - before != div_up<uint32_t>((uint32_t) d, TILE)) return PIXIU_EINTERNAL;
        tiles += descs.size() - before;
    }
    const size_t g0 = n_records();
    const uint64_t bytes = (uint64_t) (enc_off[n] - enc_off[0]);
    dirty = true;  // validation is over: from here on the record tables change
    grow_record_tables(g0 + n, enc_bytes + bytes, tiles);
    chunk_first.push_back((uint32_t) g0);
    chunk_count.push_back((uint32_t) n);
    for (int64_t r = 0; r < n; r++) {
        h_enc_off.push_back(enc_bytes + (uint64_t) (enc_off[r] - enc_off[0]));
        h_enc_len.push_back((uint32_t) (enc_off[r + 1] - enc_off[r]));
        h_dec_len.push_back(dl[r]);
        h_first.push_back((uint32_t) g0);
        h_tile_base.push_back(tbase[r]);
        h_live.push_back(1);
    }
    PX_CUDA(cudaMemcpyAsync(d_enc.ptr() + enc_bytes, enc + enc_off[0], bytes, cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(d_enc_off.p + g0, h_enc_off.data() + g0, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(d_enc_len.p + g0, h_enc_len.data() + g0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(d_dec_len.p + g0, h_dec_len.data() + g0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(d_first.p + g0, h_first.data() + g0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(d_tile_base.p + g0, h_tile_base.data() + g0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (!descs.empty())
        PX_CUDA(cudaMemcpyAsync(d_tile_desc.p + n_tiles, descs.data(), descs.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaStreamSynchronize(st));
    enc_bytes += bytes;
    n_tiles = tiles;
    mirror_from = n_records();
    return (int64_t) chunk_first.size() - 1;
}

}  // namespace pixiu
