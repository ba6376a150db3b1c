// getitem hot path: batched decode of PiXiu-encoded records.
//
// Replaces the recursive generator PXSGen::operator() (proj/PiXiuStr.h:110-198), which
// re-scans the referenced record from its first byte for every back reference and bubbles
// each byte through one coroutine per nesting level, by two dependency-free phases over a
// flat *decoded arena* (the records of every touched chunk, back to back, u32-addressed):
//   K10 k_token_scan   one warp per 2 KiB decode tile (tile descriptors make every tile
//                      independently parsable): 251-dispatch of PiXiuStr.h:142-160 in parallel,
//                      literal bytes go straight to the arena, every referenced byte gets a
//                      source pointer (arena position; self-overlapping references are folded
//                      onto their first period) and a literal bitmap is written;
//   K11 k_resolve      every non-literal byte chases its pointer chain to a literal; chains
//                      longer than RESOLVE_HOPS park their progress in the pointer array and the
//                      kernel is re-run — concurrent shortening makes the remaining rounds
//                      logarithmic in the nesting depth (deep chains, BASELINE config 3);
//   K12 k_copy_records only when the caller's layout differs from the arena order.
// No kernel ever waits on another thread's output, so there are no spin loops and no
// ordering hazards: pointers only ever move to an ancestor on the same chain, and bytes are
// only read from literal positions, which are final after K10.
#include <algorithm>
#include <cstring>
#include <map>

#include "index.h"
#include "store.h"

namespace pixiu {

constexpr int DEC_WARPS = 4;
constexpr uint32_t ENC_MAX = TILE + 16;
constexpr int RESOLVE_HOPS = 48;
enum : uint8_t { K_LIT = 0, K_COV = 1, K_SREF = 2, K_BREF = 3 };

struct DecodeView {
    const uint8_t *enc;
    const uint64_t *enc_off;
    const uint32_t *enc_len, *dec_len, *first, *tile_base, *tile_desc;
    const uint32_t *arena_off;  // per record: offset of its decoded bytes in the arena
    uint8_t *arena;
    uint32_t *ptr;              // per arena byte: source position (non-literal bytes only)
    uint32_t *litmap;           // per arena byte: 1 bit, set = literal
};

// every reference segment but the first and the last of a tile puts >= 7 decoded bytes into the tile
constexpr uint32_t SEG_MAX = TILE / 7 + 4;
constexpr uint32_t ENC_WORDS = (ENC_MAX + 16) / 4;
constexpr uint32_t UB_WORDS = TILE / 32 + 6;  // bitmaps over the tile's output in word-aligned ("u") coordinates
constexpr uint16_t CAND_HEAD = 0x8000;        // candidate flag: this 251 starts a reference token

struct alignas(16) WarpSmem {
    uint32_t encw[ENC_WORDS];       // encoded bytes, re-aligned: byte p of the tile's range is byte p of encw
    uint16_t cands[ENC_MAX + 16];   // positions of the 251s, ascending (| CAND_HEAD)
    // reference segments clipped to the tile, in output order
    uint32_t seg_base[SEG_MAX];   // arena position of the token's source byte 0
    uint16_t seg_k0[SEG_MAX];     // first token byte inside the tile
    uint16_t seg_per[SEG_MAX];    // period of a self-overlapping reference (0: none)
    uint16_t seg_rel[SEG_MAX];    // tile-relative output position of byte k0
    uint16_t seg_len[SEG_MAX];
    uint16_t seg_delta[SEG_MAX];  // literals after the segment: encoded position = output position + delta (mod 2^16)
    uint32_t startb[UB_WORDS];    // bit u: a segment starts at output byte u - mis
    uint32_t lit[UB_WORDS];       // literal bitmap of the tile
    uint16_t wprefix[UB_WORDS];   // segment starts before word w of startb
};

// K10: one warp per 2 KiB tile of decoded output.
//   1. stage the tile's encoded bytes in shared memory
//   2. list the 251s; walk each cluster of 251s from its first (certain) token head to classify reference heads
//      (PiXiuStr.h:142-160 dispatch); a scan over the heads gives every reference's output position, because
//      literals map 1:1:  out(p) = p + sum over earlier references (token bytes decoded - token bytes encoded)
//   3. per output byte (4 per lane, aligned to the arena's words): inside a reference segment -> source pointer
//      (coalesced), else the literal is copied straight from the staged bytes; literal bitmap for k_resolve
__global__ void __launch_bounds__(DEC_WARPS * 32)
k_token_scan(DecodeView V, const uint32_t *__restrict__ work_tile, const uint32_t *__restrict__ work_rec,
             uint32_t n_work, uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t lt = (1u << lane) - 1;
    const uint32_t w = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (w >= n_work) return;
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
    if (ne > ENC_MAX || e_end > el) {
        if (lane == 0) atomicExch(err, 4u);
        return;
    }
    // ---- 1. stage the encoded bytes re-aligned to words (the arena has slack past its end) ----
    const uint8_t *gsrc = encp + e0;
    const uint32_t a0 = (uint32_t) ((uintptr_t) gsrc & 3);
    const uint32_t *gw = reinterpret_cast<const uint32_t *>(gsrc - a0);
    const uint32_t nw = ((ne + 3) >> 2) + 2;  // two extra words: token fields are read up to 7 bytes past a head
    for (uint32_t j = lane; j < nw; j += 32) S.encw[j] = a0 ? __funnelshift_r(gw[j], gw[j + 1], 8 * a0) : gw[j];
    S.startb[lane] = 0;
    S.startb[32 + lane] = 0;
    if (lane < UB_WORDS - 64) S.startb[64 + lane] = 0;
    __syncwarp();
    const uint8_t *EB = reinterpret_cast<const uint8_t *>(S.encw);
    // ---- 2a. positions of the 251s, 4 bytes per lane ----
    uint32_t ncand = 0;
    for (uint32_t r0 = 0; r0 < ne; r0 += 128) {
        const uint32_t p = r0 + 4 * lane;
        const uint32_t wd = p < ne ? S.encw[p >> 2] : 0u;
        const uint32_t nval = p < ne ? min(ne - p, 4u) : 0u;
        uint32_t m = __vcmpeq4(wd, 0xFBFBFBFBu);  // 0xFF per byte that is 251
        m &= nval >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nval)) - 1);
        if (raw_first && p == 0) m &= ~0xFFu;  // the raw first byte is a plain literal
        const uint32_t cnt = __popc(m) >> 3;
        if (__ballot_sync(0xffffffffu, cnt != 0)) {
            uint32_t inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int) lane >= d) inc += o;
            }
            uint32_t slot = ncand + inc - cnt;
#pragma unroll
            for (int bb = 0; bb < 4; bb++)
                if ((m >> (8 * bb)) & 1u) S.cands[slot++] = (uint16_t) (p + bb);
            ncand += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncwarp();
    // ---- 2b. cluster walks, one lane per cluster: a 251 with no 251 among the 7 bytes before it surely starts
    //          a token, and the next 251 of the same cluster lies within 7 bytes of the previous token's end ----
    for (uint32_t c0 = 0; c0 < ncand; c0 += 32) {
        uint32_t ci = c0 + lane;
        if (ci >= ncand) continue;
        uint32_t e = S.cands[ci] & 0xFFFu;
        if (ci != 0 && (uint32_t) (S.cands[ci - 1] & 0xFFF) + 7 >= e) continue;
        while (true) {
            if (e + 1 >= ne) break;  // first half of an escape pair cut by the tile boundary: a literal
            const uint32_t nx = EB[e + 1];
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
            if (tl > 2) S.cands[ci] = (uint16_t) (e | CAND_HEAD);
            e += tl;
            ci++;
            while (ci < ncand && S.cands[ci] < e) ci++;  // 251s inside the token
            if (ci >= ncand || (uint32_t) S.cands[ci] >= e + 7) break;
            e = S.cands[ci] & 0xFFFu;
        }
    }
    __syncwarp();
    // ---- 2c. reference heads in order -> segments.  D = sum over the references so far of (decoded - encoded)
    //          token bytes; a token at encoded position p starts at output byte p + D - skip ----
    const uint32_t mis = (uint32_t) ((uintptr_t) (V.arena + rec_base + t0) & 3);
    uint32_t nseg = 0;
    int D = 0;
    for (uint32_t c0 = 0; c0 < ncand; c0 += 32) {
        const uint32_t c = c0 + lane;
        const uint32_t cv = c < ncand ? S.cands[c] : 0u;
        const bool head = (cv & CAND_HEAD) != 0;
        if (__ballot_sync(0xffffffffu, head) == 0) continue;
        const uint32_t p = cv & 0xFFF;
        uint32_t idx = 0, from = 0, tl = 0;
        int d = 0;
        if (head) {
            const bool big = EB[p + 1] == 1;
            idx = EB[p + 2] | (EB[p + 3] << 8);
            const uint32_t to = EB[p + 4] | (EB[p + 5] << 8);
            from = big ? (uint32_t) (EB[p + 6] | (EB[p + 7] << 8)) : to - EB[p + 1];
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
            int o = __shfl_up_sync(0xffffffffu, inc, dd);
            if ((int) lane >= dd) inc += o;
        }
        const int rel = (int) p + D + (inc - d) - (int) skip;  // output position of the token's first byte
        const uint32_t k0 = rel < 0 ? (uint32_t) (-rel) : 0u;
        const uint32_t k1 = (int) tl + rel > (int) nbytes ? (uint32_t) max((int) nbytes - rel, 0) : tl;
        const bool emit = head && k0 < k1;
        const uint32_t em = __ballot_sync(0xffffffffu, emit);
        if (emit) {
            const uint32_t sidx = nseg + __popc(em & lt);
            const uint32_t src_g = chunk_first + idx;
            uint32_t sbase, per = 0;
            if (src_g == g) {  // self reference (PiXiuStr.h:168-181): overlapping copies repeat with this period
                const uint32_t period = (uint32_t) ((int) t0 + rel) - from;
                sbase = rec_base + from;
                per = tl > period ? period : 0u;
            } else {
                if (src_g > g) atomicExch(err, 3u);
                sbase = (src_g > g ? 0u : V.arena_off[src_g]) + from;
            }
            if (sidx < SEG_MAX) {
                const uint32_t srel = (uint32_t) (rel + (int) k0);
                S.seg_base[sidx] = sbase;
                S.seg_k0[sidx] = (uint16_t) k0;
                S.seg_per[sidx] = (uint16_t) per;
                S.seg_rel[sidx] = (uint16_t) srel;
                S.seg_len[sidx] = (uint16_t) (k1 - k0);
                S.seg_delta[sidx] = (uint16_t) ((int) skip - (D + inc));
                atomicOr(&S.startb[(srel + mis) >> 5], 1u << ((srel + mis) & 31));
            }
        }
        nseg += __popc(em);
        D += __shfl_sync(0xffffffffu, inc, 31);
    }
    if ((int) ne + D < (int) (skip + nbytes) || nseg > SEG_MAX) {
        if (lane == 0) atomicExch(err, 6u);
        return;
    }
    __syncwarp();
    // segment starts before each bitmap word (UB_WORDS <= 96: three words per lane)
    {
        const uint32_t c0 = __popc(S.startb[lane]);
        const uint32_t c1 = __popc(S.startb[32 + lane]);
        const uint32_t c2 = lane < UB_WORDS - 64 ? __popc(S.startb[64 + lane]) : 0u;
        uint32_t i0 = c0, i1 = c1, i2 = c2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o0 = __shfl_up_sync(0xffffffffu, i0, d), o1 = __shfl_up_sync(0xffffffffu, i1, d),
                     o2 = __shfl_up_sync(0xffffffffu, i2, d);
            if ((int) lane >= d) i0 += o0, i1 += o1, i2 += o2;
        }
        const uint32_t t0s = __shfl_sync(0xffffffffu, i0, 31), t1s = __shfl_sync(0xffffffffu, i1, 31);
        S.wprefix[lane] = (uint16_t) (i0 - c0);
        S.wprefix[32 + lane] = (uint16_t) (t0s + i1 - c1);
        if (lane < UB_WORDS - 64) S.wprefix[64 + lane] = (uint16_t) (t0s + t1s + i2 - c2);
    }
    __syncwarp();
    // ---- 3. per output byte, 4 per lane.  u = output position + mis, so that u = 0 is an aligned arena word ----
    const uint32_t nu = nbytes + mis;
    uint8_t *dstu = V.arena + rec_base + t0 - mis;
    uint32_t *gptru = V.ptr + rec_base + t0 - mis;
    for (uint32_t u0 = 0; u0 < nu; u0 += 128) {
        const uint32_t u = u0 + 4 * lane;
        const uint32_t sw = S.startb[u >> 5], sh = u & 31;
        uint32_t cnt = S.wprefix[u >> 5] + __popc(sw & ((1u << sh) - 1));
        const uint32_t nibs = (sw >> sh) & 0xFu;
        const uint32_t lo = u < mis ? mis - u : 0u;
        const uint32_t hi = u + 4 <= nu ? 4u : (nu > u ? nu - u : 0u);
        uint32_t litn = 0;
        bool generic = hi != 0;
        if (nibs == 0 && lo == 0 && hi == 4) {  // one segment (or none) governs all four bytes
            const uint32_t jj = u - mis;
            uint32_t q = jj;
            bool all_lit = true;
            generic = false;
            if (cnt) {
                const uint32_t sgi = cnt - 1;
                const uint32_t o = jj - S.seg_rel[sgi], len = S.seg_len[sgi];
                if (o + 4 <= len) {
                    all_lit = false;
                    const uint32_t kk = S.seg_k0[sgi] + o, per = S.seg_per[sgi], sb = S.seg_base[sgi];
                    uint4 pv;
                    if (per) pv = make_uint4(sb + kk % per, sb + (kk + 1) % per, sb + (kk + 2) % per, sb + (kk + 3) % per);
                    else pv = make_uint4(sb + kk, sb + kk + 1, sb + kk + 2, sb + kk + 3);
                    *reinterpret_cast<uint4 *>(gptru + u) = pv;
                } else if (o >= len) {
                    q = min((jj + S.seg_delta[sgi]) & 0xFFFFu, ENC_MAX);
                } else {
                    all_lit = false;
                    generic = true;
                }
            }
            if (all_lit) {
                const uint32_t qa = q >> 2;
                *reinterpret_cast<uint32_t *>(dstu + u) = __funnelshift_r(S.encw[qa], S.encw[qa + 1], 8 * (q & 3));
                litn = 0xFu;
            }
        }
        if (generic) {
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                cnt += (nibs >> i) & 1u;
                if (i < lo || i >= hi) continue;
                const uint32_t jj = u + i - mis;
                uint32_t q = jj;
                bool isref = false;
                if (cnt) {
                    const uint32_t sgi = cnt - 1;
                    const uint32_t o = jj - S.seg_rel[sgi];
                    if (o < S.seg_len[sgi]) {
                        isref = true;
                        const uint32_t kk = S.seg_k0[sgi] + o, per = S.seg_per[sgi];
                        gptru[u + i] = S.seg_base[sgi] + (per ? kk % per : kk);
                    } else {
                        q = min((jj + S.seg_delta[sgi]) & 0xFFFFu, ENC_MAX);
                    }
                }
                if (!isref) {
                    dstu[u + i] = EB[q];
                    litn |= 1u << i;
                }
            }
        }
        uint32_t v = litn << (4 * (lane & 7));
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        v |= __shfl_xor_sync(0xffffffffu, v, 4);
        if ((lane & 7) == 0) S.lit[(u0 >> 5) + (lane >> 3)] = v;
    }
    __syncwarp();
    // ---- 4. the tile's slice of the literal bitmap (zero-initialised map, OR-ed in: tiles share words) ----
    {
        const uint32_t B0 = rec_base + t0 - mis, sh = B0 & 31;
        uint32_t *lm = V.litmap + (B0 >> 5);
        const uint32_t nsw = (nu + 31) / 32, ngw = (sh + nu + 31) / 32;
        for (uint32_t j = lane; j < ngw; j += 32) {
            uint32_t lo = j < nsw ? S.lit[j] : 0u, hi = (j > 0 && sh) ? S.lit[j - 1] : 0u;
            uint32_t v = sh ? ((lo << sh) | (hi >> (32 - sh))) : lo;
            if (v) atomicOr(&lm[j], v);
        }
    }
}

// K11: four arena bytes per thread; every non-literal byte chases its pointer chain to a literal
__global__ void __launch_bounds__(256)
k_resolve(uint32_t n, uint8_t *__restrict__ arena, uint32_t *__restrict__ ptr, const uint32_t *__restrict__ litmap,
          uint32_t *__restrict__ unfinished, uint32_t *__restrict__ err) {
    const uint32_t i4 = (blockIdx.x * 256 + threadIdx.x) * 4;
    bool pending = false;
    if (i4 < n) {
        const uint32_t bits = (litmap[i4 >> 5] >> (i4 & 31)) & 0xFu;
        if (bits != 0xFu) {
            uint32_t word = *reinterpret_cast<const uint32_t *>(arena + i4);
            uint32_t changed = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t i = i4 + b;
                if (((bits >> b) & 1u) || i >= n) continue;
                uint32_t p = ptr[i];
                bool done = false;
                for (int h = 0; h < RESOLVE_HOPS; h++) {
                    if (p >= i) {  // sources always precede their byte in the arena: corrupt input
                        atomicExch(err, 7u);
                        done = true;
                        p = i;
                        break;
                    }
                    if ((litmap[p >> 5] >> (p & 31)) & 1u) {
                        word = (word & ~(0xFFu << (8 * b))) | ((uint32_t) arena[p] << (8 * b));
                        changed |= 1u << b;
                        done = true;
                        break;
                    }
                    p = ptr[p];
                }
                if (p != i) ptr[i] = p;  // the literal origin, or an ancestor further up the chain
                pending |= !done;
            }
            if (changed) {
                if (i4 + 3 < n) {
                    *reinterpret_cast<uint32_t *>(arena + i4) = word;
                } else {
                    for (int b = 0; b < 4; b++)
                        if ((changed >> b) & 1u) arena[i4 + b] = (uint8_t) (word >> (8 * b));
                }
            }
        }
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

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
void Store::decode_records(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off) {
    if (recs.empty()) return;
    // per touched chunk: records [first, max requested] form the arena
    std::map<uint32_t, uint32_t> chunk_max;  // chunk first record -> max requested record
    for (uint32_t g : recs) {
        uint32_t f = h_first[g];
        auto it = chunk_max.find(f);
        if (it == chunk_max.end()) chunk_max[f] = g;
        else it->second = std::max(it->second, g);
    }
    const size_t NR = n_records();
    std::vector<uint32_t> aoff(NR, 0);
    std::vector<uint32_t> wt, wr;
    uint64_t arena_bytes = 0;
    uint32_t lo_g = 0xFFFFFFFFu, hi_g = 0;
    double alg_bytes = 0;
    // direct mode: the request is exactly the arena order and the caller's layout is packed the same way
    bool direct = ((uintptr_t) d_out & 127) == 0;
    size_t ri = 0;
    for (auto &cm : chunk_max)
        for (uint32_t g = cm.first; g <= cm.second; g++) {
            if (direct) {
                if (ri < recs.size() && recs[ri] == g && out_off[ri] == arena_bytes) ri++;
                else direct = false;
            }
            aoff[g] = (uint32_t) arena_bytes;
            uint32_t nt = div_up<uint32_t>(h_dec_len[g], TILE);
            for (uint32_t t = 0; t < nt; t++) {
                wt.push_back(h_tile_base[g] + t);
                wr.push_back(g);
            }
            arena_bytes += h_dec_len[g];
            alg_bytes += (double) h_enc_len[g] + h_dec_len[g];
            lo_g = std::min(lo_g, g);
            hi_g = std::max(hi_g, g);
        }
    if (ri != recs.size()) direct = false;
    if (arena_bytes >= 0xFFFFFF00ull) throw std::runtime_error("decode: arena of one call exceeds 4 GiB; split the batch");
    if (!direct) dec_scratch.reserve_discard(arena_bytes + 256);  // same packed layout, private buffer
    const uint64_t n_work = wt.size();
    uint8_t *arena = direct ? d_out : dec_scratch.p;
    dec_loc.reserve_discard(1);  // (unused in this scheme)
    dec_flags.reserve_discard(arena_bytes / 32 + 64);        // literal bitmap
    dec_ptr.reserve_discard(arena_bytes + 64);                // source pointers
    dec_aoff.reserve_discard(NR + 1);
    PX_CUDA(cudaMemcpyAsync(dec_aoff.p + lo_g, aoff.data() + lo_g, (size_t) (hi_g - lo_g + 1) * sizeof(uint32_t),
                            cudaMemcpyHostToDevice, st));
    dec_work.reserve_discard(2 * n_work + 2);
    PX_CUDA(cudaMemcpyAsync(dec_work.p, wt.data(), n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(dec_work.p + n_work, wr.data(), n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    dec_ctr.reserve_discard(64);
    PX_CUDA(cudaMemsetAsync(dec_ctr.p, 0, 64 * sizeof(uint32_t), st));
    PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, (arena_bytes / 32 + 2) * sizeof(uint32_t), st));
    DecodeView V{d_enc.ptr(), d_enc_off.p, d_enc_len.p, d_dec_len.p, d_first.p, d_tile_base.p, d_tile_desc.p,
                 dec_aoff.p, arena, dec_ptr.p, dec_flags.p};
    const size_t smem = sizeof(WarpSmem) * DEC_WARPS;
    // (per device and cheap: no process-wide "already done" flag, a process may drive several GPUs)
    PX_CUDA(cudaFuncSetAttribute(k_token_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    PX_CUDA(cudaEventRecord(ev0, st));
    prof.begin(PC_DECODE, st);
    k_token_scan<<<(unsigned) div_up<uint64_t>(n_work, DEC_WARPS), DEC_WARPS * 32, smem, st>>>(
        V, dec_work.p, dec_work.p + n_work, (uint32_t) n_work, dec_ctr.p);
    int nl = 1;
    // resolve rounds: one is enough unless chains are deeper than RESOLVE_HOPS
    uint32_t h_ctr[2] = {0, 0};
    for (int round = 0; round < 40; round++) {
        k_resolve<<<(unsigned) div_up<uint64_t>(div_up<uint64_t>(arena_bytes, 4), 256), 256, 0, st>>>((uint32_t) arena_bytes, arena, dec_ptr.p,
                                                                               dec_flags.p, dec_ctr.p + 1 + round, dec_ctr.p);
        nl++;
        PX_CUDA(cudaMemcpyAsync(h_ctr, dec_ctr.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaMemcpyAsync(h_ctr + 1, dec_ctr.p + 1 + round, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        PX_CUDA(cudaStreamSynchronize(st));
        if (h_ctr[0] || h_ctr[1] == 0) break;
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
    if (h_ctr[1]) throw std::runtime_error("decode: reference chains did not resolve");
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
