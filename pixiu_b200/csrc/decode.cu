// getitem hot path: batched decode of PiXiu-encoded records.
//
// Replaces the recursive generator PXSGen::operator() (proj/PiXiuStr.h:110-198), which re-scans the referenced
// record from its first byte for every back reference and bubbles each byte through one coroutine per nesting
// level, by TWO kernels over a flat *decoded arena* (the records of every touched chunk, back to back,
// u32-addressed), one warp per 2 KiB decode tile (tile descriptors make every tile independently parsable):
//   K10 k_decode_literals  no dependencies, pure throughput.  A tile's encoded bytes are staged in shared memory by one
//                          bulk-async copy (cp.async.bulk onto a per-warp mbarrier).  The 251-dispatch of
//                          PiXiuStr.h:142-160 runs in parallel
//                          over the staged encoded bytes (bitmap of the 251s, one lane per token cluster), a scan of
//                          (decoded - encoded) token bytes places every reference token, and the tile is then written
//                          ONCE, each lane assembling a 64-byte strip of output words straight from the staged bytes:
//                          literal bytes with their value, reference bytes as zero = "not final yet".  Every reference
//                          token leaves copy *pieces* of at most 32 bytes (destination, source, length; self-overlapping
//                          references keep their period) in a packed table in HBM.
//   K11 k_decode_copies    the data flow: tiles by ticket in arena order, one lane per piece, a piece copies as soon as
//                          its source bytes are final.  Finality travels IN BAND: a nonzero byte is final the moment it is
//                          visible; a final byte whose value is zero is announced in a 1-bit-per-byte bitmap that stays
//                          all-zero between calls (the words a call dirties are listed and cleared again).  The dependency
//                          depth is the nesting depth of BYTES, not of tiles or records; there is no flag, no fence and no
//                          memset of the arena.
//   K12 k_copy_records     only when the caller's layout differs from the arena order.
// Why two kernels: a single data-flow kernel that parses, places and copies (the first version of this round, ~6,000
// SASS instructions of mostly straight-line code with warps in every phase at once) saturated the GPC instruction cache
// (ncu: gcc__cache_requests_type_instruction 93 % of peak, stall_no_inst the top stall) long before any data path.  Split
// this way each kernel is a compact loop nest, and the polling kernel is a few hundred instructions.
// Traffic per decoded byte: the encoded byte read once, the byte written once (reference bytes twice), 8 bytes of
// piece table written and read per reference token, the source bytes of every reference read once.
// Waiting (K11) is deadlock-free: a source always precedes its destination in the arena and tickets follow arena
// order, so every byte a warp waits for belongs to a tile that is finished or held by a resident warp.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>

#include "index.h"
#include "store.h"

namespace pixiu {

constexpr int DEC_WARPS = 4;
#ifndef PIXIU_DEC_MINB
#define PIXIU_DEC_MINB 7     // resident CTAs per SM the decode kernels are compiled for (register budget 65536 / (128 x this))
#endif
constexpr uint32_t ENC_MAX = TILE + 16;           // encoded bytes a tile can span
#ifndef PIXIU_DEC_TMA
#define PIXIU_DEC_TMA 1           // stage a tile's encoded bytes with one bulk-async copy instead of a load/store loop
#endif
constexpr uint32_t STG_PAD = 16;                  // free bytes in front of the staged range (reads just before it stay in bounds)
constexpr uint32_t STG_BYTES = STG_PAD + 16 + ENC_MAX + 16;  // pad + 16-byte alignment slack + range + token read-ahead
constexpr uint32_t STG_WORDS = (STG_BYTES + 31) / 32 * 8;  // whole 32-byte bitmap words
constexpr uint32_t BM_WORDS = 96;                 // bitmaps: three words per lane
// every reference token but the first and the last of a tile puts >= 7 decoded bytes into the tile
constexpr uint32_t SEG_MAX = TILE / 7 + 4;
constexpr uint32_t TOK_MAX = SEG_MAX + 8;         // token table entries of a tile (+ sentinel)
// copy pieces of a tile: one per reference token of up to 32 bytes, one per 32 bytes of a longer one
constexpr uint32_t PIECE_BUF = SEG_MAX + TILE / 32 + 8;
// K11 copies the pieces of a tile in batches of at most PEND_MAX (four rows of 32 lanes)
constexpr uint32_t PEND_MAX = 128;
constexpr uint32_t PEND_ROWS = PEND_MAX / 32;
constexpr uint32_t SPIN_LIMIT = 1u << 22;
constexpr uint32_t STRIP = 64;                    // output bytes a lane assembles
// plain references of UNIT_MIN bytes or more are copied by the whole warp, UNIT_BYTES at a time, one destination word
// per lane (a "unit"); everything else by one lane per piece of at most 32 bytes
constexpr uint32_t UNIT_MIN = 64, UNIT_BYTES = 120;
constexpr uint32_t META_UNIT = 0x80000000u;       // piece meta: the piece is a unit, (n - 1) takes bits 12..18
static_assert(STG_BYTES <= BM_WORDS * 32 && TILE + 4 <= (BM_WORDS - 1) * 32, "bitmaps too small");

// the touched part of one chunk: records [first, last] form a contiguous part of the arena
struct DecRange {
    uint32_t first, last;      // records
    uint32_t tile_lo, ntiles;  // their decode tiles (global tile ids are consecutive inside a chunk)
    uint32_t rec_cum, tile_cum;  // records / tiles of the ranges before this one
    uint32_t arena_base, pad;
};

// one work item = one tile, everything its warp needs to start in ONE 48-byte read (k_dec_work derives it from the
// record tables, so that the warp does not walk a chain of dependent table loads)
struct alignas(16) TileJob {
    uint64_t enc_pos;       // offset in the compressed arena of the first staged byte
    uint32_t rec_base;      // arena offset of the tile's record
    uint32_t g;             // the record
    uint32_t t0_nbytes;     // first decoded byte of the tile in its record (lo16), decoded bytes of the tile (hi16)
    uint32_t ne_skip;       // staged encoded bytes (lo16); bytes of the first token to skip (hi16; 0xFFFF: raw first byte)
    uint32_t chunk_first;   // global id of record 0 of the chunk (back references carry chunk-local indices)
    uint32_t range;         // index of the DecRange
    uint32_t pad0, pad1, pad2, pad3;
};
static_assert(sizeof(TileJob) == 48, "TileJob layout");

struct DecodeView {
    const uint8_t *enc;
    const TileJob *jobs;        // in ticket order
    const uint32_t *arena_off;  // per record: offset of its decoded bytes in the arena
    uint8_t *arena;
    uint32_t *fin;              // per arena byte: 1 bit, set = the byte is final AND its value is zero
    uint32_t *dirty;            // words of `fin` this call has set bits in (cleared again by k_fin_clean)
    uint32_t dirty_cap;
    uint2 *pieces;              // packed piece table: {source arena position, meta}
    uint32_t pieces_cap;
    uint2 *phead;               // per tile (ticket order): {first piece, number of pieces}
};

// what the out-of-line helpers need (passed by value: a reference to the kernel's parameter block would force a copy
// of it into local memory)
struct PubCtx {
    uint8_t *arena;
    uint32_t *fin, *dirty, *ctr;
    uint32_t dirty_cap;
};

// counters of a decode call (dec_ctr)
enum { DC_ERR = 0, DC_TICKET = 1, DC_DIRTY = 2, DC_PIECES = 3, DC_TICKET2 = 4, DC_SWEEPS = 5 };

struct ParseBits {                  // dead once the heads are listed: shares its storage with the token table
    uint32_t b251[BM_WORDS];        // bit p: staged byte p is a 251 that can start a token
    uint32_t cst[BM_WORDS];         // bit p: that 251 surely starts a token (no 251 among the 7 bytes before it)
    uint16_t cl[SEG_MAX + 8];       // staged positions of those cluster starts, ascending
};
// reference tokens of a tile in output order, in "u" coordinates (u = output byte of the tile + misalignment of the
// tile's first byte, so that u = 0 is a word boundary of the arena); entry nt is a sentinel
struct TokTab {
    int16_t ts[TOK_MAX];            // first output byte of token i (clipped to the tile)
    int16_t te[TOK_MAX];            // one past its last output byte (clipped)
    int16_t dl[TOK_MAX];            // the literals in front of token i: staged position = u + dl[i]
};
struct Pending {
    // meta: u (12 bits) | (n - 1) << 12 (5 bits) | period << 17 (5 bits, 0 = plain copy) | phase << 22
    uint32_t src[PEND_MAX];         // arena position of the first source byte (periodic: of the period's byte 0)
    uint32_t meta[PEND_MAX];
    uint32_t done[PEND_MAX];        // bytes of the piece already copied
};

struct alignas(16) WarpSmem {       // K10
    uint32_t stg[STG_WORDS];        // staged encoded bytes: byte e0 + k of the record sits at staged position soff + k
    union {
        ParseBits ps;
        TokTab tt;
    };
    uint16_t heads[SEG_MAX + 8];    // staged positions of the reference heads, ascending
    uint2 piece[PIECE_BUF];         // the tile's copy pieces, flushed to the packed table at the end
    uint16_t first_tok[TILE / STRIP + 4];  // per strip: the token that governs its first byte (u <= 2062: 33 strips)
    uint64_t mbar;                  // the staging copy of a tile completes here (cp.async.bulk transaction bytes)
};
struct alignas(16) CopySmem {       // K11
    Pending pc;
    uint32_t pend[PEND_ROWS];       // per row of 32 pieces: lanes whose piece is not copied yet
    uint32_t upend[PEND_ROWS];      // per row: slots that hold a unit (copied by the whole warp) not finished yet
};

__device__ __forceinline__ uint32_t nib251(uint32_t w) {
    return ((__vcmpeq4(w, 0xFBFBFBFBu) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t nibz(uint32_t w) {   // 4-bit mask of the zero bytes of w
    return ((__vcmpeq4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t nibnz(uint32_t w) {  // 4-bit mask of the nonzero bytes of w
    return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}

__device__ __forceinline__ uint32_t atom_relaxed_or_u32(uint32_t *p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.gpu.global.or.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// stores that other warps poll: strong (relaxed, device scope) so that the poll and the store are both morally
// strong operations; a byte is only ever stored with its final value (K10 has left zero = "not final yet")
__device__ __forceinline__ void st_pub_u8(uint8_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u8 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ void st_pub_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v));
}

// announce final bytes whose value is zero: bit j of `zmask` = arena byte B + j (j < 32)
__device__ __noinline__ void publish_zeros(PubCtx P, uint32_t B, uint32_t zmask) {
    const uint32_t bs = B & 31;
    uint32_t wi = B >> 5, bits = zmask << bs;
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
        if (bits) {
            if (atom_relaxed_or_u32(P.fin + wi, bits) == 0) {  // first bit in this word: remember it for the clean-up
                const uint32_t k = atomicAdd(P.ctr + DC_DIRTY, 1u);
                if (k < P.dirty_cap) P.dirty[k] = wi;
            }
        }
        bits = bs ? zmask >> (32 - bs) : 0u;
        wi++;
    }
}

// One lane writes n (1..32) bytes at dst from a source given as nine aligned 32-bit words: byte j of the source is
// byte sa + j of w0..w8.  Head bytes up to a word boundary of the destination, whole words, tail bytes.  ONE copy of
// this code serves the literal runs, the retired references and the pending pieces (the kernel's working set has to
// fit the instruction cache: the first version inlined it five times and stalled on instruction fetch).
__device__ __noinline__ void put_bytes(uint8_t *dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                       uint32_t w6, uint32_t w7, uint32_t w8, uint32_t sa, uint32_t n) {
    uint32_t sw[10] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, 0u};
    const uint32_t hbe = min((4u - ((uint32_t) (uintptr_t) dst & 3u)) & 3u, n);  // head bytes up to a destination word boundary
    const uint32_t x0 = __funnelshift_r(sw[0], sw[1], 8 * sa);
#pragma unroll
    for (int i = 0; i < 3; i++)
        if ((uint32_t) i < hbe) st_pub_u8(dst + i, (x0 >> (8 * i)) & 0xFFu);
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
        if ((uint32_t) t < m) st_pub_u32(dw + t, x);
        if ((uint32_t) t == m) xt = x;
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
        if ((uint32_t) i < tb) st_pub_u8(reinterpret_cast<uint8_t *>(dw + m) + i, (xt >> (8 * i)) & 0xFFu);
}

// Copy the pending pieces of a tile: one piece per lane and row, sweeping until all are done.  A piece loads its
// source words (L2-coherent relaxed loads), copies the bytes that are final (nonzero, or zero and announced in the
// bitmap), remembers them in its done mask and retries the rest: no flag round trip, no fence, and the critical path
// is the nesting depth of BYTES.  Returns the number of sweeps, or 0xFFFFFFFF on a time-out / foreign error.
__device__ __noinline__ uint32_t drain_pending(CopySmem &S, PubCtx P, uint32_t B0, uint32_t npiece, uint32_t sleep_after,
                                               uint32_t sleep_ns, uint32_t sweep_gap) {
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lane = lane_id();
    uint8_t *const dstu = P.arena + B0;
    const uint32_t nrows = (npiece + 31) / 32;
    uint32_t remaining = 0, uremaining = 0;
    for (uint32_t row = 0; row < PEND_ROWS; row++) {   // which slots are lane pieces, which are units
        const uint32_t k = row * 32 + lane;
        const bool valid = k < npiece, unit = valid && (S.pc.meta[k] & META_UNIT);
        if (valid) S.pc.done[k] = 0;
        const uint32_t vm = __ballot_sync(FULL, valid && !unit), um = __ballot_sync(FULL, unit);
        if (lane == 0) {
            S.pend[row] = vm;
            S.upend[row] = um;
        }
        remaining += __popc(vm);
        uremaining += __popc(um);
    }
    __syncwarp();
    uint32_t spins = 0, sweep = 3;  // (the first sweep examines every piece)
    for (; remaining | uremaining; sweep++) {
        uint32_t any = 0;
        // A waiting piece costs one load per sweep: the source byte behind its first open byte (its last one on odd
        // sweeps: looking at the first byte only makes a byte wait for everything left of it in its piece, and with a
        // source window that slides from record to record that running maximum chains every record to its predecessor);
        // the whole window is examined when that byte has arrived and every fourth sweep (bytes out of order,
        // zero-valued ones).  The polls of all rows are issued before any of them is consumed: one L2 round trip per
        // sweep, not one per row (the sweep period is what a dependency hop costs).
        uint32_t pollw[PEND_ROWS];
        const bool full_look = (sweep & 3u) == 3u;
#pragma unroll
        for (uint32_t row = 0; row < PEND_ROWS; row++) {
            pollw[row] = 1u;
            if (!full_look && row < nrows && ((S.pend[row] >> lane) & 1u)) {
                const uint32_t slot = row * 32 + lane;
                const uint32_t meta = S.pc.meta[slot], a = S.pc.src[slot];
                const uint32_t per = (meta >> 17) & 31u, n = per ? per : ((meta >> 12) & 31u) + 1;
                const uint32_t open = (0xFFFFFFFFu >> (32 - n)) & ~(per ? 0u : S.pc.done[slot]);
                const uint32_t pj = a + ((sweep & 1u) ? 31u - (uint32_t) __clz(open) : (uint32_t) __ffs(open) - 1u);
                pollw[row] = (ld_poll_u32(reinterpret_cast<const uint32_t *>(P.arena + (pj & ~3u))) >> (8 * (pj & 3))) & 0xFFu;
            }
        }
        static_assert(PEND_ROWS == 4, "row select below");
#pragma unroll 1
        for (uint32_t row = 0; row < nrows; row++) {
            const uint32_t pm = S.pend[row];
            if (!pm) continue;
            const uint32_t pw = row == 0 ? pollw[0] : row == 1 ? pollw[1] : row == 2 ? pollw[2] : pollw[3];
            bool complete = false;
            uint32_t avail = 0;
            if ((pm >> lane) & 1u) {
                const uint32_t slot = row * 32 + lane;
                const uint32_t meta = S.pc.meta[slot], a = S.pc.src[slot];
                const uint32_t per = (meta >> 17) & 31u, np = ((meta >> 12) & 31u) + 1, us = meta & 0xFFFu;
                const uint32_t n = per ? per : np;  // source bytes
                const uint32_t need = 0xFFFFFFFFu >> (32 - n);
                uint32_t dn = per ? 0u : S.pc.done[slot];
                const bool look = full_look || pw != 0;
                if (look) {
                    const uint32_t sa = a & 3, nsw = (sa + n + 3) >> 2;  // aligned source words (<= 9)
                    const uint32_t *wp0 = reinterpret_cast<const uint32_t *>(P.arena + (a - sa));
                    uint32_t sw[9];
#pragma unroll
                    for (int q = 0; q < 9; q++) sw[q] = (uint32_t) q < nsw ? ld_poll_u32(wp0 + q) : 0u;
                    // source bytes seen nonzero are final.  The common case - every byte of the range nonzero - is settled by
                    // a cheap zero-byte test over the words (bytes outside the range forced nonzero); the exact mask only
                    // when some byte is still zero
                    uint32_t nz = need;
                    {
                        const uint32_t end = sa + n, L = (end - 1) >> 2, e8 = 8 * (end & 3);
                        uint32_t z = 0;
#pragma unroll
                        for (int q = 0; q < 9; q++) {
                            uint32_t w = sw[q];
                            if (q == 0) w |= (1u << (8 * sa)) - 1;
                            if ((uint32_t) q == L && e8) w |= 0xFFFFFFFFu << e8;
                            if ((uint32_t) q > L) w = 0xFFFFFFFFu;
                            z |= (w - 0x01010101u) & ~w & 0x80808080u;
                        }
                        if (z) {
                            unsigned long long nzm = 0;
#pragma unroll
                            for (int q = 0; q < 9; q++) nzm |= (unsigned long long) nibnz(sw[q]) << (4 * q);
                            nz = (uint32_t) (nzm >> sa) & need;
                        }
                    }
                    uint32_t zf = 0;  // source bytes that are final zeros
                    const uint32_t zc = need & ~nz & ~dn;
                    if (zc) {
                        const uint32_t wi = a >> 5, bs = a & 31;
                        uint32_t bits = ld_relaxed_u32(P.fin + wi) >> bs;
                        if (bs + n > 32) bits |= ld_relaxed_u32(P.fin + wi + 1) << (32 - bs);
                        zf = zc & bits;
                    }
                    avail = (nz | zf) & ~dn;
                    uint32_t zout = 0;  // copied bytes whose value is zero (destination coordinates of the piece)
                    if (per) {
                        avail = avail == need ? need : 0u;  // short periods are copied in one go
                        if (avail) {
                            uint8_t *dst = dstu + us;
                            const uint32_t ph = meta >> 22;
#pragma unroll 1
                            for (uint32_t j = 0; j < np; j += 4) {  // four loads in flight (the period is final: plain L2 loads)
                                uint32_t bv[4];
#pragma unroll
                                for (int i = 0; i < 4; i++) bv[i] = __ldcg(P.arena + a + (ph + j + i) % per);
#pragma unroll
                                for (int i = 0; i < 4; i++)
                                    if (j + i < np) {
                                        st_pub_u8(dst + j + i, bv[i]);
                                        if (bv[i] == 0) zout |= 1u << (j + i);
                                    }
                            }
                            complete = true;
                        }
                    } else if (avail == need && dn == 0) {
                        // the common case: the whole piece is final -> straight copy
                        put_bytes(dstu + us, sw[0], sw[1], sw[2], sw[3], sw[4], sw[5], sw[6], sw[7], sw[8], sa, n);
                        zout = zf;
                        complete = true;
                    } else if (avail) {
                        // destination word t holds piece bytes [4t - da, 4t - da + 4); its source bytes straddle the
                        // aligned source words t + c and t + c + 1 (c = -1 when the source sits further left in its word)
                        const uint32_t da = us & 3;
                        const int delta = (int) sa - (int) da;
                        const uint32_t sh = 8u * (uint32_t) (delta & 3);
                        uint32_t prev = 0;
                        uint32_t *dw = reinterpret_cast<uint32_t *>(dstu + us - da);
                        const unsigned long long am = (unsigned long long) avail << da;  // bytes to store, word coordinates
                        const uint32_t ndw = (da + n + 3) >> 2;
#pragma unroll
                        for (int t = 0; t < 9; t++) {
                            // (delta < 0: word t straddles source words t - 1 and t)
                            const uint32_t lo_w = delta < 0 ? prev : sw[t], hi_w = delta < 0 ? sw[t] : (t < 8 ? sw[t + 1] : 0u);
                            prev = sw[t];
                            const uint32_t vm = (uint32_t) (am >> (4 * t)) & 0xFu;
                            if ((uint32_t) t < ndw && vm) {
                                const uint32_t x = __funnelshift_r(lo_w, hi_w, sh);
                                if (vm == 0xFu) {
                                    st_pub_u32(dw + t, x);
                                } else {
#pragma unroll
                                    for (int i = 0; i < 4; i++)
                                        if ((vm >> i) & 1u) st_pub_u8(reinterpret_cast<uint8_t *>(dw + t) + i, (x >> (8 * i)) & 0xFFu);
                                }
                            }
                        }
                        zout = zf & avail;
                        dn |= avail;
                        S.pc.done[slot] = dn;
                        complete = dn == need;
                    }
                    if (zout) publish_zeros(P, B0 + us, zout);
                }
            }
            const uint32_t cb = __ballot_sync(FULL, complete);
            if (__ballot_sync(FULL, avail != 0)) any = 1;
            if (cb) {
                if (lane == 0) S.pend[row] = pm & ~cb;
                remaining -= __popc(cb);
                __syncwarp();
            }
        }
        // ---- units: the whole warp, one destination word per lane; a word is stored as soon as its source bytes are
        //      final (progress per word, not per unit: a long reference into a region that is itself being copied
        //      follows it word by word).  The units are examined one after the other, each a little later than the one
        //      before: sampling them all at once (loads of several units in flight together) was measured and needs
        //      more sweeps - on deep chains the sweeps, not the load latency, are what costs ----
        if (uremaining) {
#pragma unroll 1
            for (uint32_t row = 0; row < nrows; row++) {
                uint32_t um = S.upend[row];
#pragma unroll 1
                while (um) {
                    const uint32_t slot = row * 32 + (uint32_t) __ffs(um) - 1u;
                    um &= um - 1;
                    const uint32_t meta = S.pc.meta[slot], A = S.pc.src[slot], dn = S.pc.done[slot];
                    const uint32_t B = B0 + (meta & 0xFFFu), n = ((meta >> 12) & 0x7Fu) + 1;
                    const uint32_t da = B & 3u, nw = (da + n + 3) >> 2;   // destination words (<= 31)
                    bool ready = false;
                    if (lane < nw && !((dn >> lane) & 1u)) {
                        const int j0 = 4 * (int) lane - (int) da;      // unit byte held by the word's byte 0 (< 0 in word 0 when da > 0)
                        const uint32_t lo = j0 < 0 ? (uint32_t) (-j0) : 0u, hi = min(4u, (uint32_t) ((int) n - j0));
                        const uint32_t sp = A + (uint32_t) (j0 + (int) lo);   // source position of the first valid byte
                        const uint32_t *wp = reinterpret_cast<const uint32_t *>(P.arena + (sp & ~3u));
                        const uint32_t w0 = ld_poll_u32(wp), w1 = ((sp & 3u) + (hi - lo) > 4u) ? ld_poll_u32(wp + 1) : 0u;
                        const uint32_t x = __funnelshift_r(w0, w1, 8 * (sp & 3u)) << (8 * lo);
                        const uint32_t vb = ((1u << hi) - 1) & ~((1u << lo) - 1);      // valid bytes of the word (4 bits)
                        const uint32_t zb = nibz(x) & vb;                                 // valid bytes that read zero
                        ready = zb == 0;
                        if (!ready) {   // zero: final only if announced in the bitmap (bits of source bytes sp ...)
                            const uint32_t wi = sp >> 5, bs = sp & 31;
                            uint32_t bits = ld_relaxed_u32(P.fin + wi) >> bs;
                            if (bs > 28) bits |= ld_relaxed_u32(P.fin + wi + 1) << (32 - bs);
                            ready = (((zb >> lo) & ~bits) & 0xFu) == 0;
                        }
                        if (ready) {
                            uint8_t *d = P.arena + (B - da) + 4 * lane;
                            if (vb == 0xFu) {
                                st_pub_u32(reinterpret_cast<uint32_t *>(d), x);
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; i++)
                                    if ((vb >> i) & 1u) st_pub_u8(d + i, (x >> (8 * i)) & 0xFFu);
                            }
                            if (zb) publish_zeros(P, (B - da) + 4 * lane, zb);
                        }
                    }
                    const uint32_t rb = __ballot_sync(FULL, ready);
                    if (rb) {
                        any = 1;
                        const uint32_t nd = dn | rb;
                        if (nd == (nw >= 32 ? 0xFFFFFFFFu : (1u << nw) - 1)) {
                            if (lane == 0) S.upend[row] &= ~(1u << (slot & 31));
                            uremaining--;
                        } else if (lane == 0) {
                            S.pc.done[slot] = nd;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (!any) {
            if (++spins > SPIN_LIMIT || ld_relaxed_u32(P.ctr + DC_ERR) != 0) {
                if (lane == 0) atomicCAS(P.ctr + DC_ERR, 0u, 8u);
                return 0xFFFFFFFFu;
            }
            if (spins > sleep_after) __nanosleep(spins > 256 ? 400 : sleep_ns);
        } else {
            spins = 0;
            if (sweep_gap && (remaining | uremaining)) __nanosleep(sweep_gap);
        }
    }
    return sweep - 3;
}


// K10: one warp per 2 KiB tile of decoded output; no waiting anywhere.
//   1. stage the tile's encoded bytes in shared memory (one bulk-async copy per tile, completion on the warp's mbarrier)
//   2. bitmap of the 251s; a 251 with no 251 among the 7 bytes before it surely starts a token: these cluster starts
//      are listed, and one lane per cluster walks its tokens (PiXiuStr.h:142-160 dispatch) and counts / lists the
//      reference heads (a cluster is almost always a single token)
//   3. heads in order, 32 per round, one lane per reference token: a scan of (decoded - encoded) token bytes gives its
//      output position (literals map 1:1); the lane enters the token in the tile's token table and its copy pieces
//      (<= 32 bytes each) in the piece buffer
//   4. the tile is written once: every lane assembles a 64-byte strip word by word, walking the token table from the
//      token that governs the strip's first byte - literal bytes come straight from the staged bytes (one unaligned
//      32-bit read per word and run), reference bytes are zero; zero-valued literals are announced in the bitmap
//   5. the pieces go to the packed table (one atomic add per tile)
__global__ void __launch_bounds__(DEC_WARPS * 32, PIXIU_DEC_MINB)
k_decode_literals(DecodeView V, uint32_t n_work, uint32_t *__restrict__ ctr) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t FULL = 0xffffffffu;
    uint32_t *err = ctr + DC_ERR;
    const PubCtx P{V.arena, V.fin, V.dirty, ctr, V.dirty_cap};
    // persistent warps: every warp takes tiles by ticket, TICKETS at a time (nothing here waits on another tile, so the
    // order does not matter and one atomic serves several tiles), until none is left
    constexpr uint32_t TICKETS = 4;
    uint32_t w = 0, w_end = 0;
    if (lane < 4) S.stg[lane] = 0;   // the pad in front of the staged bytes stays zero
#if PIXIU_DEC_TMA
    const uint32_t mbar = (uint32_t) __cvta_generic_to_shared(&S.mbar);
    const uint32_t stg_dst = (uint32_t) __cvta_generic_to_shared(S.stg + 4);
    uint32_t mphase = 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
#endif
#pragma unroll 1
    for (;; w++) {
    __syncwarp();
    if (w == w_end) {
        if (lane == 0) w = atomicAdd(ctr + DC_TICKET, TICKETS);
        w = __shfl_sync(FULL, w, 0);
        w_end = w + TICKETS;
    }
    if (w >= n_work) return;
    // ---- the job: two 16-byte reads ----
    const uint4 j0 = __ldg(reinterpret_cast<const uint4 *>(V.jobs + w)), j1 = __ldg(reinterpret_cast<const uint4 *>(V.jobs + w) + 1);
    const uint8_t *gsrc = V.enc + (((uint64_t) j0.y << 32) | j0.x);
    const uint32_t rec_base = j0.z, g = j0.w;
    const uint32_t t0 = j1.x & 0xFFFFu, nbytes = j1.x >> 16;
    const uint32_t ne = j1.y & 0xFFFFu;
    uint32_t skip = j1.y >> 16;
    const bool raw_first = skip == 0xFFFF;  // first enc byte is the 2nd half of an escape pair
    if (raw_first) skip = 0;
    const uint32_t chunk_first = j1.z;
    const uint32_t mis = (rec_base + t0) & 15u;  // (the arena itself is at least 16-byte aligned)
    const uint32_t B0 = rec_base + t0 - mis;     // arena position of u = 0 (16-byte aligned); u = output byte + mis
    const uint32_t nu = nbytes + mis;
    bool failed = false;
    if (ne > ENC_MAX) {   // (k_dec_work flags an inconsistent descriptor this way)
        if (lane == 0) atomicExch(err, 4u);
        failed = true;
    }
    // ---- 1. stage the encoded bytes (the compressed arena has slack past its end) ----
    const uint32_t a16 = (uint32_t) ((uintptr_t) gsrc & 15);
    const uint32_t soff = STG_PAD + a16, nstg = soff + ne;
    uint32_t nheads = 0;
    if (!failed) {
        const uint32_t n16 = (a16 + ne + 8 + 15) >> 4;
#if PIXIU_DEC_TMA
        // one bulk copy (TMA, cp.async.bulk) per tile; its bytes land on the warp's mbarrier.  The warp's reads of
        // the previous tile are behind the __syncwarp at the top of the loop; the proxy fence orders them before the
        // asynchronous write
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(n16 * 16) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             stg_dst), "l"(gsrc - a16), "r"(n16 * 16), "r"(mbar)
                         : "memory");
        }
        {
            uint32_t ok = 0, spins = 0;
            while (true) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok)
                             : "r"(mbar), "r"(mphase)
                             : "memory");
                if (ok) break;
                if (++spins > (1u << 22)) {   // (never seen: a copy that does not land would hang the warp)
                    atomicExch(err, 9u);
                    failed = true;
                    break;
                }
            }
            mphase ^= 1u;
            failed = __any_sync(FULL, failed);
        }
#else
        const uint4 *g4 = reinterpret_cast<const uint4 *>(gsrc - a16);
        uint4 *s4 = reinterpret_cast<uint4 *>(S.stg);
#pragma unroll 1
        for (uint32_t j = lane; j < n16; j += 32) s4[1 + j] = g4[j];
        __syncwarp();
#endif
    }
    const uint8_t *SB = reinterpret_cast<const uint8_t *>(S.stg);
    if (!failed) {
        // ---- 2a. bitmap of the 251s (32 staged bytes per lane and step) ----
        {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(S.stg);
            const uint32_t lo = soff + (raw_first ? 1u : 0u);  // the raw first byte is a plain literal
#pragma unroll 1
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
        // ---- 2b. cluster starts (lane l owns the consecutive bitmap words 3l .. 3l+2), listed in ascending order ----
        uint32_t ncl;
        {
            uint32_t cs[3], cnt = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint32_t wi = 3 * lane + k;
                const uint32_t b = S.ps.b251[wi], pb = wi ? S.ps.b251[wi - 1] : 0u;
                unsigned long long y = (((unsigned long long) b << 32) | pb) << 1;
                y |= y << 1;
                y |= y << 2;
                y |= y << 3;  // OR of the shifts 1..7
                cs[k] = b & ~(uint32_t) (y >> 32);
                cnt += __popc(cs[k]);
            }
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(FULL, inc, d);
                if ((int) lane >= d) inc += o;
            }
            ncl = __shfl_sync(FULL, inc, 31);
            __syncwarp();
            uint32_t slot = inc - cnt;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                S.ps.cst[3 * lane + k] = cs[k];
                uint32_t hv = cs[k];
                while (hv) {
                    if (slot < SEG_MAX + 8) S.ps.cl[slot] = (uint16_t) ((3 * lane + k) * 32 + __ffs(hv) - 1);
                    slot++;
                    hv &= hv - 1;
                }
            }
            if (ncl > SEG_MAX + 8) {   // (cluster starts are >= 8 bytes apart: cannot happen on a tile of <= 2064 bytes)
                if (lane == 0) atomicExch(err, 6u);
                failed = true;
                ncl = 0;
            }
        }
        __syncwarp();
        // ---- 2c. one lane per cluster: walk its tokens, count the reference heads (pass 0), list them (pass 1) ----
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < ncl; c0 += 32) {
            const bool has = c0 + lane < ncl;
            const uint32_t e_first = has ? S.ps.cl[c0 + lane] : 0u;
            uint32_t nh = 0, h0 = 0, hbase = 0;
            bool again = false;   // (a cluster with more than one reference head is walked a second time to list them)
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                uint32_t e = e_first, k = 0;
                while (has && (pass == 0 || again)) {
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
                    if (tl > 2) {
                        if (pass && hbase + k < SEG_MAX + 8) S.heads[hbase + k] = (uint16_t) e;
                        if (k == 0) h0 = e;
                        k++;
                    }
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
                if (pass == 0) {
                    nh = k;
                    uint32_t inc = nh;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        uint32_t o = __shfl_up_sync(FULL, inc, d);
                        if ((int) lane >= d) inc += o;
                    }
                    hbase = nheads + inc - nh;
                    nheads += __shfl_sync(FULL, inc, 31);
                    if (nh == 1 && hbase < SEG_MAX + 8) S.heads[hbase] = (uint16_t) h0;
                    again = nh > 1;
                    if (!__any_sync(FULL, again)) break;
                }
            }
        }
        if (nheads > SEG_MAX) {
            if (lane == 0) atomicExch(err, 6u);
            failed = true;
            nheads = 0;
        }
        __syncwarp();   // (from here on ParseBits is dead: its storage holds the pending pieces)
    }
    // ---- 3. rounds of 32 reference tokens, one lane each ----
    uint32_t npiece = 0;        // pieces in the buffer
    int D = 0;                  // sum of (decoded - encoded) bytes of the reference tokens so far
    int carry_out = 0;          // output position where the previous token ended (literals before the first one start at 0)
    uint32_t carry_stg = soff;  // staged position behind the previous token
    const uint32_t nstrips = (nu + STRIP - 1) / STRIP;
    for (uint32_t s = lane; s < nstrips; s += 32) S.first_tok[s] = (uint16_t) nheads;   // (the sentinel governs what no token claims)
    // arena offset of the source record of this lane's token, fetched one round ahead (a dependent L2 round trip
    // that would otherwise sit in the middle of every round)
    auto source_base = [&](uint32_t c) -> uint32_t {
        if (c >= nheads) return 0u;
        const uint32_t p2 = S.heads[c] + 2u;
        const uint32_t src_g = chunk_first + (__funnelshift_r(S.stg[p2 >> 2], S.stg[(p2 >> 2) + 1], 8 * (p2 & 3)) & 0xFFFFu);
        return src_g < g ? V.arena_off[src_g] : 0u;
    };
    uint32_t aoff_cur = failed ? 0u : source_base(lane);
    __syncwarp();
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < nheads && !failed; c0 += 32) {
        const uint32_t c = c0 + lane;
        const bool head = c < nheads;
        const uint32_t aoff_next = source_base(c + 32);
        uint32_t p = 0, idx = 0, from = 0, tl = 0, elen = 0;
        int d = 0;
        if (head) {
            p = S.heads[c];
            // the token's eight bytes: two unaligned 32-bit reads of the staged words
            const uint32_t wq = p >> 2, sh = 8 * (p & 3);
            const uint32_t a0 = S.stg[wq], a1 = S.stg[wq + 1], a2 = S.stg[wq + 2];
            const uint32_t f0 = __funnelshift_r(a0, a1, sh), f1 = __funnelshift_r(a1, a2, sh);
            const uint32_t b1 = (f0 >> 8) & 0xFFu;
            const bool big = b1 == 1;
            idx = f0 >> 16;
            const uint32_t to = f1 & 0xFFFFu;
            from = big ? f1 >> 16 : to - b1;
            if (to <= from || from > 0xFFFF) {
                atomicExch(err, 5u);
                from = to;
            }
            tl = to - from;
            elen = big ? 8u : 6u;
            d = (int) tl - (int) elen;
        }
        int inc = d;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            int o = __shfl_up_sync(FULL, inc, dd);
            if ((int) lane >= dd) inc += o;
        }
        const int rel = (int) (p - soff) + D + (inc - d) - (int) skip;  // output position of the token's first byte
        const int end_out = rel + (int) tl;
        const uint32_t end_stg = p + elen;
        int pe_out = __shfl_up_sync(FULL, end_out, 1);
        uint32_t pe_stg = __shfl_up_sync(FULL, end_stg, 1);
        if (lane == 0) {
            pe_out = carry_out;
            pe_stg = carry_stg;
        }
        // the token in the tile: output bytes [k0, k1) of it
        const uint32_t k0 = rel < 0 ? (uint32_t) (-rel) : 0u;
        const uint32_t k1 = end_out > (int) nbytes ? (uint32_t) max((int) nbytes - rel, 0) : tl;
        const bool emit = head && k0 < k1;
        // ---- token table (every head, also one that puts nothing into the tile) and the strips it governs ----
        if (head) {
            const int us0 = min(max(rel, 0), (int) nbytes) + (int) mis, ue0 = min(max(end_out, 0), (int) nbytes) + (int) mis;
            S.tt.ts[c] = (int16_t) us0;
            S.tt.te[c] = (int16_t) ue0;
            S.tt.dl[c] = (int16_t) ((int) pe_stg - pe_out - (int) mis);
            // token c governs the positions [end of token c - 1, end of token c): the strips that start in there
            const int gov0 = c == 0 ? 0 : min(max(pe_out, 0), (int) nbytes) + (int) mis;
            for (uint32_t s = ((uint32_t) gov0 + STRIP - 1) / STRIP; s * STRIP < (uint32_t) ue0; s++) S.first_tok[s] = (uint16_t) c;
        }
        // ---- copy pieces ----
        uint32_t sbase = 0, per = 0, ks = k0, us = 0, len = 0;
        if (emit) {
            const uint32_t src_g = chunk_first + idx;
            if (src_g == g) {  // self reference (PiXiuStr.h:168-181): overlapping copies repeat with this period
                const uint32_t period = (uint32_t) ((int) t0 + rel) - from;
                sbase = rec_base + from;
                per = tl > period ? period : 0u;
                if ((int) t0 + rel <= (int) from) atomicExch(err, 3u);
            } else {
                if (src_g > g) atomicExch(err, 3u);
                sbase = aoff_cur + from;
            }
            len = k1 - k0;
            if (per) {
                ks = k0 % per;
                if (ks + len <= per) per = 0;  // this part does not wrap: a plain copy out of the first period
            }
            us = (uint32_t) (rel + (int) k0) + mis;
        }
        {
            // pieces of a token: 32 destination bytes each (a period >= 32 that wraps: also cut where it wraps)
            const bool wrapper = emit && per >= 32, units = emit && per == 0 && len >= UNIT_MIN;
            uint32_t np = !emit || wrapper ? 0u : units ? (len + UNIT_BYTES - 1) / UNIT_BYTES : (len + 31) >> 5;
            if (wrapper) {
                for (uint32_t pos = ks, o = 0; o < len; np++) {
                    const uint32_t n = min(min(32u, per - pos), len - o);
                    pos = pos + n == per ? 0u : pos + n;
                    o += n;
                }
            }
            uint32_t pinc = np;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) {
                uint32_t o = __shfl_up_sync(FULL, pinc, dd);
                if ((int) lane >= dd) pinc += o;
            }
            const uint32_t tot = __shfl_sync(FULL, pinc, 31);
            if (npiece + tot > PIECE_BUF) {
                if (lane == 0) atomicExch(err, 6u);
                failed = true;
            } else {
                uint32_t slot = npiece + pinc - np;
                if (wrapper) {
                    for (uint32_t pos = ks, o = 0; o < len; slot++) {
                        const uint32_t n = min(min(32u, per - pos), len - o);
                        S.piece[slot] = make_uint2(sbase + pos, (us + o) | ((n - 1) << 12));
                        pos = pos + n == per ? 0u : pos + n;
                        o += n;
                    }
                } else if (units) {
                    for (uint32_t o = 0; o < len; o += UNIT_BYTES, slot++) {
                        const uint32_t n = min(UNIT_BYTES, len - o);
                        S.piece[slot] = make_uint2(sbase + ks + o, (us + o) | ((n - 1) << 12) | META_UNIT);
                    }
                } else {
                    for (uint32_t o = 0; o < len; o += 32, slot++) {
                        const uint32_t n = min(32u, len - o);
                        S.piece[slot] = per ? make_uint2(sbase, (us + o) | ((n - 1) << 12) | (per << 17) | (((ks + o) % per) << 22))
                                            : make_uint2(sbase + ks + o, (us + o) | ((n - 1) << 12));
                    }
                }
                npiece += tot;
            }
        }
        // carries for the next round
        aoff_cur = aoff_next;
        const uint32_t last = min(nheads - c0, 32u) - 1;
        carry_out = __shfl_sync(FULL, end_out, last);
        carry_stg = __shfl_sync(FULL, end_stg, last);
        D += __shfl_sync(FULL, inc, last);
    }
    if (!failed) {
        if ((int) ne + D < (int) (skip + nbytes)) {
            if (lane == 0) atomicExch(err, 6u);
            failed = true;
        }
    }
    if (lane == 0) {   // sentinel: behind the last token everything is literal
        S.tt.ts[nheads] = (int16_t) nu;
        S.tt.te[nheads] = (int16_t) nu;
        S.tt.dl[nheads] = (int16_t) ((int) carry_stg - carry_out - (int) mis);
    }
    __syncwarp();
    // ---- 4. the tile is written once, a 64-byte strip per lane, 16 bytes (one store) at a time ----
    if (!failed) {
        uint8_t *const dstu = V.arena + B0;
#pragma unroll 1
        for (uint32_t s = lane; s < nstrips; s += 32) {
            uint32_t i = S.first_tok[s];
            int ts_i = S.tt.ts[i], te_i = S.tt.te[i], dl_i = S.tt.dl[i];
            const int s_end = (int) min((s + 1) * STRIP, nu);
#pragma unroll 1
            for (int pos = (int) (s * STRIP); pos < s_end; pos += 16) {
                const int lo = max(pos, (int) mis), end = min(pos + 16, s_end);   // bytes [lo, end) of this group belong to the tile
                uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0;   // the group's four words
                uint32_t lit = 0;                          // its literal bytes (one bit each)
                int cur = lo;
                while (true) {
                    if (cur < ts_i) {   // literals in front of token i: bytes [cur, min(ts_i, end)), staged at u + dl_i
                        const int se = min(ts_i, end);
                        const uint32_t bm = (0xFFFFu >> (16 - (se - pos))) & (0xFFFFu << (cur - pos));
                        const uint32_t q = (uint32_t) (pos + dl_i), wq = q >> 2, sh = 8 * (q & 3);
                        const uint32_t a0 = S.stg[wq], a1 = S.stg[wq + 1], a2 = S.stg[wq + 2], a3 = S.stg[wq + 3], a4 = S.stg[wq + 4];
                        // (nibble -> byte mask: 0xF -> 0xFFFFFFFF, 0x3 -> 0x0000FFFF, ...)
                        x0 |= __funnelshift_r(a0, a1, sh) & ((((bm & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
                        x1 |= __funnelshift_r(a1, a2, sh) & (((((bm >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
                        x2 |= __funnelshift_r(a2, a3, sh) & (((((bm >> 8) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
                        x3 |= __funnelshift_r(a3, a4, sh) & (((((bm >> 12) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
                        lit |= bm;
                        cur = se;
                    }
                    if (cur < te_i) cur = min(te_i, end);   // bytes of token i: zero = "not final yet"
                    if (cur >= end) break;
                    i++;
                    ts_i = S.tt.ts[i];
                    te_i = S.tt.te[i];
                    dl_i = S.tt.dl[i];
                }
                if (lo == pos && end == pos + 16) {
                    *reinterpret_cast<uint4 *>(dstu + pos) = make_uint4(x0, x1, x2, x3);
                } else {
                    for (int k = lo; k < end; k++) {
                        const uint32_t xw = (k - pos) < 4 ? x0 : (k - pos) < 8 ? x1 : (k - pos) < 12 ? x2 : x3;
                        dstu[k] = (uint8_t) (xw >> (8 * ((k - pos) & 3)));
                    }
                }
                // zero-valued literal bytes are final: announce them
                const uint32_t zb = (nibz(x0) | (nibz(x1) << 4) | (nibz(x2) << 8) | (nibz(x3) << 12)) & lit;
                if (zb) publish_zeros(P, B0 + (uint32_t) pos, zb);
            }
        }
    }
    // ---- 5. the pieces go to the packed table ----
    uint32_t pbase = 0;
    if (lane == 0) {
        pbase = failed ? 0u : atomicAdd(ctr + DC_PIECES, npiece);
        if (!failed && pbase + npiece > V.pieces_cap) {
            atomicExch(err, 10u);
            failed = true;
        }
        V.phead[w] = make_uint2(pbase, failed ? 0u : npiece);
    }
    pbase = __shfl_sync(FULL, pbase, 0);
    failed = __shfl_sync(FULL, (int) failed, 0) != 0;
    if (!failed)
        for (uint32_t k = lane; k < npiece; k += 32) V.pieces[pbase + k] = S.piece[k];
    }  // next ticket
}

// K11: the data flow.  Tiles by ticket in arena order; the pieces of a tile are copied in batches of up to 128 (one lane
// per piece and row, see drain_pending); K10 has written every literal and zeroed every reference byte of the whole
// arena before this kernel starts, so polling needs no flag.
__global__ void __launch_bounds__(DEC_WARPS * 32, 8)
k_decode_copies(DecodeView V, uint32_t n_work, uint32_t *__restrict__ ctr, uint32_t piece_cap, uint32_t sleep_after,
                uint32_t sleep_ns, uint32_t sweep_gap, unsigned long long *__restrict__ trace) {
    __shared__ CopySmem smem[DEC_WARPS];
    CopySmem &S = smem[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t FULL = 0xffffffffu;
    const PubCtx P{V.arena, V.fin, V.dirty, ctr, V.dirty_cap};
#pragma unroll 1
    for (;;) {
        __syncwarp();
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(ctr + DC_TICKET2, 1u);
        w = __shfl_sync(FULL, w, 0);
        if (w >= n_work) return;
        if (trace && lane == 0) trace[4 * (size_t) w] = globaltimer_ns();
        const uint2 ph = V.phead[w];
        const uint4 j0 = __ldg(reinterpret_cast<const uint4 *>(V.jobs + w)), j1 = __ldg(reinterpret_cast<const uint4 *>(V.jobs + w) + 1);
        const uint32_t at = j0.z + (j1.x & 0xFFFFu);   // arena position of the tile's first byte
        const uint32_t B0 = at - (at & 15u);
        uint32_t sweeps = 0;
#pragma unroll 1
        for (uint32_t b0 = 0; b0 < ph.y; b0 += piece_cap) {
            const uint32_t n = min(piece_cap, ph.y - b0);
            for (uint32_t k = lane; k < n; k += 32) {
                const uint2 r = V.pieces[ph.x + b0 + k];
                S.pc.src[k] = r.x;
                S.pc.meta[k] = r.y;
            }
            __syncwarp();
            const uint32_t sw = drain_pending(S, P, B0, n, sleep_after, sleep_ns, sweep_gap);
            if (sw == 0xFFFFFFFFu) return;   // (time-out or foreign error: the flag is set)
            sweeps += sw;
        }
        if (trace && lane == 0) {  // measurement aid (PIXIU_DEC_TRACE_FILE): entry, -, done, sweeps
            trace[4 * (size_t) w + 1] = trace[4 * (size_t) w];
            trace[4 * (size_t) w + 2] = globaltimer_ns();
            trace[4 * (size_t) w + 3] = sweeps;
        }
    }
}

// clears the words of the zero-byte bitmap a decode call has set bits in (the bitmap stays all-zero between calls)
__global__ void __launch_bounds__(256)
k_fin_clean(uint32_t *__restrict__ fin, const uint32_t *__restrict__ dirty, const uint32_t *__restrict__ ctr, uint32_t cap) {
    const uint32_t n = min(ctr[DC_DIRTY], cap);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) fin[dirty[i]] = 0;
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

// One TileJob per work item.  Tiles of one chunk keep their order (the data-flow argument needs it); chunks are
// interleaved round-robin, so that as many dependency chains as there are chunks advance side by side: the k-th tile
// of range c goes to position sum_c' min(ntiles_c', k) + #{c' < c : ntiles_c' > k}.
__global__ void __launch_bounds__(256)
k_dec_work(uint32_t n_work, uint32_t n_ranges, const DecRange *__restrict__ R, const uint32_t *__restrict__ tile_base,
           const uint32_t *__restrict__ tile_desc, const uint64_t *__restrict__ enc_off, const uint32_t *__restrict__ enc_len,
           const uint32_t *__restrict__ dec_len, const uint32_t *__restrict__ first, const uint32_t *__restrict__ aoff,
           const uint8_t *__restrict__ enc, TileJob *__restrict__ jobs) {
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
    const uint32_t g = a, t = gt - tile_base[g];
    const uint32_t dl = dec_len[g], el = enc_len[g];
    const uint32_t t0 = t * TILE, t1 = min(dl, t0 + TILE);
    const uint32_t desc = tile_desc[gt];
    const uint32_t e0 = desc & 0xffff;
    uint32_t e_end = el;
    if (t1 < dl) {
        const uint32_t d2 = tile_desc[gt + 1];
        e_end = d2 & 0xffff;
        const uint32_t sk2 = d2 >> 16;
        if (sk2 != 0 && sk2 != 0xFFFF) e_end += (enc[enc_off[g] + e_end + 1] == 1) ? 8u : 6u;
    }
    uint32_t ne = e_end - e0;
    if (e_end > el || e_end < e0 || ne > ENC_MAX) ne = 0xFFFFu;  // inconsistent descriptor: the tile's warp reports it
    TileJob J;
    J.enc_pos = enc_off[g] + e0;
    J.rec_base = aoff[g];
    J.g = g;
    J.t0_nbytes = t0 | ((t1 - t0) << 16);
    J.ne_skip = ne | (desc & 0xFFFF0000u);
    J.chunk_first = first[g];
    J.range = lo;
    J.pad0 = J.pad1 = J.pad2 = J.pad3 = 0;
    jobs[pos] = J;
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
// The arena of one decode pass is addressed with 32 bits.  A request whose touched chunks decode to more than
// DEC_ARENA_LIMIT bytes is split by chunk (chunks are self-contained) into passes that each stay below it.
void Store::decode_records(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off) {
    if (recs.empty()) return;
    if (mirror_from < n_records()) flush_mirrors();  // the running sum below is built from the host mirrors
    const uint64_t LIMIT = knobs.dec_arena_limit;  // 3.5 GiB unless the test knob says otherwise
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
    // zero-byte bitmap: all-zero between calls (a (re)allocation zeroes it once; a call clears the words it dirtied)
    const size_t bm_words = arena_bytes / 32 + 64;
    if (bm_words > dec_flags.cap) {
        dec_flags.reserve_discard(bm_words);
        PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, dec_flags.cap * sizeof(uint32_t), st));
    }
    const size_t dirty_cap = std::max<size_t>(arena_bytes / 128, 4096);
    dec_dirty.reserve_discard(dirty_cap);
    dec_aoff.reserve_discard(NR + 1);
    dec_work.reserve_discard((n_work + 1) * (sizeof(TileJob) / sizeof(uint32_t)));
    // piece table: one piece per reference token (a token is >= 6 encoded bytes) plus one per 32 decoded bytes of the long
    // ones, plus slack per tile; 8 B per piece, and 8 B per tile for the table's directory
    uint64_t enc_span = 0;
    for (const DecRange &r : ranges) enc_span += h_enc_off[r.last] + h_enc_len[r.last] - h_enc_off[r.first];
    const uint64_t pieces_cap = std::min<uint64_t>(enc_span / 6 + arena_bytes / 32 + 2 * n_work + 64, 0xFFFFFF00ull);
    dec_pieces.reserve_discard(pieces_cap);
    dec_phead.reserve_discard(n_work + 1);
    dec_ranges.reserve_discard(ranges.size() * sizeof(DecRange) / sizeof(uint32_t) + 8);
    DecRange *d_ranges = reinterpret_cast<DecRange *>(dec_ranges.p);
    PX_CUDA(cudaMemcpyAsync(d_ranges, ranges.data(), ranges.size() * sizeof(DecRange), cudaMemcpyHostToDevice, st));
    k_dec_aoff<<<(unsigned) div_up<uint64_t>(n_rec, 256), 256, 0, st>>>((uint32_t) n_rec, (uint32_t) ranges.size(), d_ranges,
                                                                        d_dec_prefix.p, dec_aoff.p);
    TileJob *d_jobs = reinterpret_cast<TileJob *>(dec_work.p);
    k_dec_work<<<(unsigned) div_up<uint64_t>(n_work, 256), 256, 0, st>>>(
        (uint32_t) n_work, (uint32_t) ranges.size(), d_ranges, d_tile_base.p, d_tile_desc.p, d_enc_off.p, d_enc_len.p, d_dec_len.p,
        d_first.p, dec_aoff.p, d_enc.ptr(), d_jobs);
    launches += 2;
    dec_ctr.reserve_discard(64);
    PX_CUDA(cudaMemsetAsync(dec_ctr.p, 0, 64 * sizeof(uint32_t), st));
    DecodeView V{d_enc.ptr(), d_jobs, dec_aoff.p, arena, dec_flags.p, dec_dirty.p, (uint32_t) dirty_cap,
                 reinterpret_cast<uint2 *>(dec_pieces.p), (uint32_t) pieces_cap, reinterpret_cast<uint2 *>(dec_phead.p)};
    const size_t smem = sizeof(WarpSmem) * DEC_WARPS;
    // (per device and cheap: no process-wide "already done" flag, a process may drive several GPUs)
    PX_CUDA(cudaFuncSetAttribute(k_decode_literals, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    const uint32_t piece_cap = std::min<uint32_t>(std::max<uint32_t>(knobs.piece_cap, 1u), PEND_MAX);
    if (knobs.trace)
        fprintf(stderr, "[decode] host work list %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_h0).count());
    unsigned long long *d_trace = nullptr;
    const char *trace_file = knobs.dec_trace_file.empty() ? nullptr : knobs.dec_trace_file.c_str();  // measurement aid
    if (trace_file) {
        PX_CUDA(cudaMalloc(&d_trace, 4 * n_work * sizeof(unsigned long long)));
        PX_CUDA(cudaMemsetAsync(d_trace, 0, 4 * n_work * sizeof(unsigned long long), st));
    }
    PX_CUDA(cudaEventRecord(ev0, st));
    prof.begin(PC_DECODE, st);
    if (!dec_sms) {
        int dev = 0, sms = 0;
        PX_CUDA(cudaGetDevice(&dev));
        PX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        dec_sms = (uint32_t) sms;
    }
    // persistent warps: one CTA per resident CTA slot of every SM, fewer for small calls
    const uint64_t ctas = div_up<uint64_t>(n_work, DEC_WARPS);
    prof.begin(PC_DECODE_LIT, st);
    k_decode_literals<<<(unsigned) std::min<uint64_t>(ctas, (uint64_t) dec_sms * PIXIU_DEC_MINB), DEC_WARPS * 32, smem, st>>>(
        V, (uint32_t) n_work, dec_ctr.p);
    prof.end(st, alg_bytes, 1);
    prof.begin(PC_DECODE_COPY, st);
    k_decode_copies<<<(unsigned) std::min<uint64_t>(ctas, (uint64_t) dec_sms * std::min<uint32_t>(std::max<uint32_t>(knobs.copy_ctas, 1u), 8u)), DEC_WARPS * 32, 0, st>>>(
        V, (uint32_t) n_work, dec_ctr.p, piece_cap, knobs.sleep_after, knobs.sleep_ns, knobs.sweep_gap, d_trace);
    prof.end(st, 0.0, 1);
    k_fin_clean<<<64, 256, 0, st>>>(dec_flags.p, dec_dirty.p, dec_ctr.p, (uint32_t) dirty_cap);
    int nl = 3;
    uint32_t h_ctr[8] = {0};
    PX_CUDA(cudaMemcpyAsync(h_ctr, dec_ctr.p, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
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
    last_pending_pieces = h_ctr[DC_PIECES];
    last_drains = h_ctr[DC_SWEEPS];
    if (h_ctr[DC_DIRTY] > dirty_cap)  // more dirty words than the list holds (zero-heavy data): clear the whole bitmap
        PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, dec_flags.cap * sizeof(uint32_t), st));
    if (d_trace) {  // file: u64 n_work, u32 record[n_work] (ticket order), u64 {entry, parsed, done (ns), sweeps}[n_work]
        std::vector<unsigned long long> ht(4 * n_work);
        std::vector<uint32_t> hr(n_work);
        PX_CUDA(cudaMemcpy(ht.data(), d_trace, ht.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        {
            std::vector<TileJob> hj(n_work);
            PX_CUDA(cudaMemcpy(hj.data(), d_jobs, n_work * sizeof(TileJob), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n_work; i++) hr[i] = hj[i].g;
        }
        if (FILE *f = fopen(trace_file, "wb")) {
            const unsigned long long nw = n_work;
            fwrite(&nw, 8, 1, f);
            fwrite(hr.data(), 4, n_work, f);
            fwrite(ht.data(), 8, ht.size(), f);
            fclose(f);
        }
        cudaFree(d_trace);
    }
    if (knobs.trace)
        fprintf(stderr, "[decode] %llu tiles, %u copy pieces, %u dirty bitmap words\n", (unsigned long long) n_work, h_ctr[DC_PIECES],
                h_ctr[DC_DIRTY]);
    if (h_ctr[DC_ERR]) {
        // (the bitmap may hold bits of tiles that never finished)
        PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, dec_flags.cap * sizeof(uint32_t), st));
        throw std::runtime_error("decode: kernel reported error " + std::to_string(h_ctr[DC_ERR]));
    }
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
