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

constexpr int DEC_WARPS = 8;
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

struct WarpSmem {
    uint8_t enc[ENC_MAX];
    uint8_t kind[ENC_MAX];
    uint8_t out[TILE];
    uint32_t lit[TILE / 32];
    uint16_t queue[256];
    uint32_t qn;
};

// bytes [k0, k1) of a reference token (token-relative): write their source pointers
// rel0: tile-relative decoded offset of the token's first byte (may be negative)
__device__ __forceinline__ void emit_ref_ptrs(const DecodeView &V, uint32_t g, uint32_t rec_base, uint32_t t0, int rel0,
                                              uint32_t idx, uint32_t from, uint32_t k0, uint32_t k1, uint32_t step,
                                              uint32_t lane_off, uint32_t *err) {
    uint32_t src_g = V.first[g] + idx;
    uint32_t *dst = V.ptr + rec_base + t0;
    if (src_g == g) {
        // self reference (PiXiuStr.h:168-181): overlapping copies repeat with period = token start - from
        uint32_t period = (uint32_t) ((int) t0 + rel0) - from;
        uint32_t base = rec_base + from;
        if (from + k1 <= from + period) {
            for (uint32_t k = k0 + lane_off; k < k1; k += step) dst[rel0 + (int) k] = base + k;
        } else {
            for (uint32_t k = k0 + lane_off; k < k1; k += step) dst[rel0 + (int) k] = base + (k % period);
        }
    } else {
        if (src_g > g) {
            atomicExch(err, 3u);
            return;
        }
        uint32_t base = V.arena_off[src_g] + from;
        for (uint32_t k = k0 + lane_off; k < k1; k += step) dst[rel0 + (int) k] = base + k;
    }
}

__device__ __forceinline__ uint32_t tok_dlen(const WarpSmem &S, uint32_t p) {
    uint8_t k = S.kind[p];
    if (k == K_LIT) return 1;
    if (k == K_SREF) return S.enc[p + 1];
    if (k == K_BREF)
        return (uint32_t) (S.enc[p + 4] | (S.enc[p + 5] << 8)) - (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
    return 0;
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
k_token_scan(DecodeView V, const uint32_t *__restrict__ work_tile, const uint32_t *__restrict__ work_rec,
             uint32_t n_work, uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * DEC_WARPS + (threadIdx.x >> 5);
    if (w >= n_work) return;
    const uint32_t gt = work_tile[w], g = work_rec[w];
    const uint32_t t = gt - V.tile_base[g];
    const uint32_t dl = V.dec_len[g], el = V.enc_len[g];
    const uint32_t t0 = t * TILE, t1 = min(dl, t0 + TILE), nbytes = t1 - t0;
    const uint32_t rec_base = V.arena_off[g];
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
    // ---- 1. stage encoded bytes; default token kinds ----
    for (uint32_t p = lane; p < ne; p += 32) {
        uint8_t b = encp[e0 + p];
        S.enc[p] = b;
        S.kind[p] = b == 251 ? K_COV : K_LIT;
    }
    for (uint32_t j = lane; j < TILE / 32; j += 32) S.lit[j] = 0;
    if (lane == 0) {
        S.qn = 0;
        if (raw_first) S.kind[0] = K_LIT;
    }
    __syncwarp();
    // ---- 2. token heads: a 251 with no 251 among the 7 bytes before it surely starts a token;
    //         its owner walks the cluster of nearby 251s (PiXiuStr.h:142-160 dispatch) ----
    const uint32_t pstart = raw_first ? 1u : 0u;
    for (uint32_t p = pstart + lane; p < ne; p += 32) {
        if (S.enc[p] != 251) continue;
        bool certain = true;
        for (uint32_t q = (p >= pstart + 7 ? p - 7 : pstart); q < p; q++) certain &= S.enc[q] != 251;
        if (!certain) continue;
        uint32_t e = p;
        while (true) {
            if (e + 1 >= ne) {  // first half of an escape pair cut by the tile boundary
                S.kind[e] = K_LIT;
                break;
            }
            uint32_t nx = S.enc[e + 1];
            uint32_t tl;
            if (nx == 0 || nx == 251 || nx == 2) {
                S.kind[e] = K_LIT;
                S.kind[e + 1] = K_LIT;
                tl = 2;
            } else if (nx == 1) {
                S.kind[e] = K_BREF;
                tl = 8;
            } else if (nx > 6) {
                S.kind[e] = K_SREF;
                tl = 6;
            } else {
                atomicExch(err, 5u);  // 3..6: invalid (assert(false), PiXiuStr.h:193)
                break;
            }
            if (tl > 2)
                for (uint32_t q = e + 1; q < e + tl && q < ne; q++) S.kind[q] = K_COV;
            e += tl;
            // the next 251 of the same cluster lies within 7 bytes of the last one seen
            uint32_t q = e;
            while (q < ne && q < e + 7 && S.enc[q] != 251) q++;
            if (q >= ne || q >= e + 7) break;
            e = q;
        }
    }
    __syncwarp();
    // ---- 3. decoded offset of every token: per-lane strips + warp scan ----
    const uint32_t strip = (ne + 31) / 32;
    const uint32_t p0 = min(lane * strip, ne), p1 = min(p0 + strip, ne);
    uint32_t sum = 0;
    for (uint32_t p = p0; p < p1; p++) sum += tok_dlen(S, p);
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if ((int) lane >= d) inc += o;
    }
    uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (total < skip + nbytes) {
        if (lane == 0) atomicExch(err, 6u);
        return;
    }
    // ---- 4. emit: literal bytes to the tile buffer, reference bytes to the pointer array ----
    int rel = (int) (inc - sum) - (int) skip;  // tile-relative decoded offset at p0
    for (uint32_t p = p0; p < p1; p++) {
        uint8_t k = S.kind[p];
        if (k == K_LIT) {
            if (rel >= 0 && rel < (int) nbytes) {
                S.out[rel] = S.enc[p];
                atomicOr(&S.lit[rel >> 5], 1u << (rel & 31));
            }
            rel += 1;
        } else if (k == K_SREF || k == K_BREF) {
            uint32_t idx = S.enc[p + 2] | (S.enc[p + 3] << 8);
            uint32_t to = S.enc[p + 4] | (S.enc[p + 5] << 8);
            uint32_t from = k == K_SREF ? to - S.enc[p + 1] : (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
            uint32_t tl = to - from;
            uint32_t k0 = rel < 0 ? (uint32_t) (-rel) : 0u;
            uint32_t k1 = (int) tl + rel > (int) nbytes ? (uint32_t) ((int) nbytes - rel) : tl;
            if (k0 < k1) {
                uint32_t qi = 256;
                if (k1 - k0 > 48) qi = atomicAdd(&S.qn, 1u);
                if (qi < 256) S.queue[qi] = (uint16_t) p;
                else emit_ref_ptrs(V, g, rec_base, t0, rel, idx, from, k0, k1, 1, 0, err);
            }
            rel += (int) tl;
        }
    }
    __syncwarp();
    // long references: the whole warp writes each one (coalesced)
    {
        uint32_t qn = min(S.qn, 256u);
        for (uint32_t qi = 0; qi < qn; qi++) {
            uint32_t p = S.queue[qi];
            uint32_t owner = p / strip;
            uint32_t d = __shfl_sync(0xffffffffu, inc - sum, owner);
            for (uint32_t q = owner * strip; q < p; q++) d += tok_dlen(S, q);
            int r0 = (int) d - (int) skip;
            uint8_t k = S.kind[p];
            uint32_t idx = S.enc[p + 2] | (S.enc[p + 3] << 8);
            uint32_t to = S.enc[p + 4] | (S.enc[p + 5] << 8);
            uint32_t from = k == K_SREF ? to - S.enc[p + 1] : (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
            uint32_t tl = to - from;
            uint32_t k0 = r0 < 0 ? (uint32_t) (-r0) : 0u;
            uint32_t k1 = (int) tl + r0 > (int) nbytes ? (uint32_t) ((int) nbytes - r0) : tl;
            emit_ref_ptrs(V, g, rec_base, t0, r0, idx, from, k0, k1, 32, lane, err);
        }
    }
    __syncwarp();
    // ---- 5. store the tile's bytes (literal positions are final, the rest is filled by k_resolve)
    //         and its slice of the literal bitmap; records are packed, so nothing is aligned ----
    {
        uint8_t *dst = V.arena + rec_base + t0;
        uint32_t head = (uint32_t) ((4 - ((uintptr_t) dst & 3)) & 3);
        if (head > nbytes) head = nbytes;
        if (lane < head) dst[lane] = S.out[lane];
        uint32_t nwords = (nbytes - head) >> 2;
        uint32_t *dw = (uint32_t *) (dst + head);
        for (uint32_t j = lane; j < nwords; j += 32) {
            const uint8_t *sb = S.out + head + 4 * j;
            dw[j] = (uint32_t) sb[0] | ((uint32_t) sb[1] << 8) | ((uint32_t) sb[2] << 16) | ((uint32_t) sb[3] << 24);
        }
        uint32_t tail0 = head + 4 * nwords;
        if (tail0 + lane < nbytes) dst[tail0 + lane] = S.out[tail0 + lane];
        // bitmap: global bit position B0 = rec_base + t0 (zero-initialised map, OR-ed in)
        const uint32_t B0 = rec_base + t0, sh = B0 & 31;
        uint32_t *lm = V.litmap + (B0 >> 5);
        const uint32_t nsw = (nbytes + 31) / 32, ngw = (sh + nbytes + 31) / 32;
        for (uint32_t j = lane; j < ngw; j += 32) {
            uint32_t lo = j < nsw ? S.lit[j] : 0u, hi = (j > 0 && sh) ? S.lit[j - 1] : 0u;
            uint32_t v = sh ? ((lo << sh) | (hi >> (32 - sh))) : lo;
            if (v) atomicOr(&lm[j], v);
        }
    }
}

// K11: chase every non-literal byte to its literal origin
__global__ void __launch_bounds__(256)
k_resolve(uint32_t n, uint8_t *__restrict__ arena, uint32_t *__restrict__ ptr, const uint32_t *__restrict__ litmap,
          uint32_t *__restrict__ unfinished, uint32_t *__restrict__ err) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    bool pending = false;
    if (i < n && !((litmap[i >> 5] >> (i & 31)) & 1u)) {
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
                arena[i] = arena[p];
                done = true;
                break;
            }
            p = ptr[p];
        }
        if (p != i) ptr[i] = p;  // the literal origin, or an ancestor further up the chain
        pending = !done;
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
    DecodeView V{d_enc.p, d_enc_off.p, d_enc_len.p, d_dec_len.p, d_first.p, d_tile_base.p, d_tile_desc.p,
                 dec_aoff.p, arena, dec_ptr.p, dec_flags.p};
    const size_t smem = sizeof(WarpSmem) * DEC_WARPS;
    static bool attr_set = false;
    if (!attr_set) {
        PX_CUDA(cudaFuncSetAttribute(k_token_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    PX_CUDA(cudaEventRecord(ev0, st));
    prof.begin(PC_DECODE, st);
    k_token_scan<<<(unsigned) div_up<uint64_t>(n_work, DEC_WARPS), DEC_WARPS * 32, smem, st>>>(
        V, dec_work.p, dec_work.p + n_work, (uint32_t) n_work, dec_ctr.p);
    int nl = 1;
    // resolve rounds: one is enough unless chains are deeper than RESOLVE_HOPS
    uint32_t h_ctr[2] = {0, 0};
    for (int round = 0; round < 40; round++) {
        k_resolve<<<(unsigned) div_up<uint64_t>(arena_bytes, 256), 256, 0, st>>>((uint32_t) arena_bytes, arena, dec_ptr.p,
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
    PX_CUDA(cudaMemcpyAsync(d_enc.p + enc_bytes, enc + enc_off[0], bytes, cudaMemcpyHostToDevice, st));
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
    return (int64_t) chunk_first.size() - 1;
}

}  // namespace pixiu
