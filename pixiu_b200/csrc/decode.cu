// getitem hot path: batched decode of PiXiu-encoded records.
//
// Replaces the recursive generator PXSGen::operator() (proj/PiXiuStr.h:110-198), which
// re-scans the referenced record from its first byte for every back reference, by a
// persistent data-flow kernel over 2 KiB decode tiles:
//   * tiles are handed out in (chunk, record, tile) order by a ticket counter, which is a
//     topological order of the reference DAG (a record only references earlier records
//     of its chunk, or earlier bytes of itself);
//   * a warp parses its tile's tokens in parallel (251-dispatch of PiXiuStr.h:142-160),
//     gathers literal bytes and bytes of already finished tiles (waiting on their
//     per-tile flags), resolves references into the tile itself by pointer jumping in
//     shared memory, then stores the tile and publishes its flag.
// Every decoded byte is written once and every encoded byte read once; referenced bytes
// come from tiles written moments earlier (L2 resident for window-sized chunks).
#include <algorithm>
#include <cstring>
#include <map>

#include "index.h"
#include "store.h"

namespace pixiu {

constexpr int DEC_WARPS = 8;
constexpr uint32_t ENC_MAX = TILE + 16;
constexpr uint16_t SRC_RESOLVED = 0xFFFF;
constexpr uint32_t DEC_SPIN_LIMIT = 1u << 26;
enum : uint8_t { K_LIT = 0, K_COV = 1, K_SREF = 2, K_BREF = 3 };

struct DecodeView {
    const uint8_t *enc;
    const uint64_t *enc_off;
    const uint32_t *enc_len, *dec_len, *first, *tile_base, *tile_desc;
    const uint64_t *loc;   // per record: device address of its decoded bytes
    uint32_t *flags;       // per tile
    uint32_t epoch;
};

struct WarpSmem {
    uint8_t enc[ENC_MAX];
    uint8_t kind[ENC_MAX];
    uint8_t out[TILE];
    uint16_t src[TILE];
    uint16_t queue[256];
    uint32_t qn;
};

__device__ __forceinline__ bool wait_tiles(const DecodeView &V, uint32_t src_g, uint32_t a, uint32_t b, uint32_t *err) {
    // wait until decoded bytes [a, b) of record src_g are published
    uint32_t tb = V.tile_base[src_g];
    for (uint32_t t = a / TILE; t <= (b - 1) / TILE; t++) {
        uint32_t spins = 0;
        while (ld_acquire_u32(V.flags + tb + t) != V.epoch) {
            if (++spins > DEC_SPIN_LIMIT) {
                atomicExch(err, 1u);
                return false;
            }
            __nanosleep(20);
        }
    }
    return true;
}

// bytes [k0, k1) of a reference token (token-relative) -> tile buffer / in-tile pointers
// rel0: tile-relative decoded offset of the token's first byte (may be negative)
__device__ __forceinline__ void emit_ref_bytes(const DecodeView &V, WarpSmem &S, uint32_t g, uint32_t t0, int rel0,
                                               uint32_t idx, uint32_t from, uint32_t k0, uint32_t k1, uint32_t step,
                                               uint32_t lane_off, uint32_t *err) {
    uint32_t src_g = V.first[g] + idx;
    if (src_g == g) {
        // self reference (PiXiuStr.h:168-181): bytes come from earlier output of this record;
        // overlapping copies repeat with period = token start - from
        uint32_t dst_abs = (uint32_t) ((int) t0 + rel0);
        uint32_t period = dst_abs - from;
        uint32_t ext_hi = 0;  // highest external byte needed (exclusive), for the flag wait
        for (uint32_t k = k0 + lane_off; k < k1; k += step) {
            uint32_t so = from + (k % period);
            if (so < t0) ext_hi = max(ext_hi, so + 1);
        }
        if (ext_hi) wait_tiles(V, g, from, ext_hi, err);
        const uint8_t *base = (const uint8_t *) V.loc[g];
        for (uint32_t k = k0 + lane_off; k < k1; k += step) {
            uint32_t so = from + (k % period);
            int rel = rel0 + (int) k;
            if (so >= t0) {
                S.src[rel] = (uint16_t) (so - t0);
            } else {
                S.out[rel] = __ldcg(base + so);
                S.src[rel] = SRC_RESOLVED;
            }
        }
    } else {
        if (src_g > g) {
            atomicExch(err, 3u);
            return;
        }
        if (k0 + lane_off < k1) wait_tiles(V, src_g, from + k0, from + k1, err);
        const uint8_t *base = (const uint8_t *) V.loc[src_g] + from;
        for (uint32_t k = k0 + lane_off; k < k1; k += step) {
            int rel = rel0 + (int) k;
            S.out[rel] = __ldcg(base + k);
            S.src[rel] = SRC_RESOLVED;
        }
    }
}

__global__ void __launch_bounds__(DEC_WARPS * 32)
k_decode_tiles(DecodeView V, const uint32_t *__restrict__ work_tile, const uint32_t *__restrict__ work_rec,
               uint32_t n_work, uint32_t *__restrict__ ctr /* [0] ticket, [1] err */) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    uint32_t *err = ctr + 1;

    while (true) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(ctr, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= n_work) break;
        const uint32_t gt = work_tile[w], g = work_rec[w];
        const uint32_t t = gt - V.tile_base[g];
        const uint32_t dl = V.dec_len[g], el = V.enc_len[g];
        const uint32_t t0 = t * TILE, t1 = min(dl, t0 + TILE), nbytes = t1 - t0;
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
            continue;
        }
        // ---- 1. stage encoded bytes; default token kinds ----
        for (uint32_t p = lane; p < ne; p += 32) {
            uint8_t b = encp[e0 + p];
            S.enc[p] = b;
            S.kind[p] = b == 251 ? K_COV : K_LIT;
        }
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
                for (uint32_t q = e + 1; q < e + tl && q < ne; q++)
                    if (tl > 2) S.kind[q] = K_COV;
                e += tl;
                // next 251 of the same cluster lies within 7 bytes of the last one seen
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
        for (uint32_t p = p0; p < p1; p++) {
            uint8_t k = S.kind[p];
            if (k == K_LIT) sum += 1;
            else if (k == K_SREF) sum += S.enc[p + 1];
            else if (k == K_BREF)
                sum += (uint32_t) (S.enc[p + 4] | (S.enc[p + 5] << 8)) - (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
        }
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if ((int) lane >= d) inc += o;
        }
        uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
        if (total < skip + nbytes) {
            if (lane == 0) atomicExch(err, 6u);
            continue;
        }
        // ---- 4. emit: literals and short references by the owning lane, long references queued ----
        int rel = (int) (inc - sum) - (int) skip;  // tile-relative decoded offset at p0
        for (uint32_t p = p0; p < p1; p++) {
            uint8_t k = S.kind[p];
            if (k == K_LIT) {
                if (rel >= 0 && rel < (int) nbytes) {
                    S.out[rel] = S.enc[p];
                    S.src[rel] = SRC_RESOLVED;
                }
                rel += 1;
            } else if (k == K_SREF || k == K_BREF) {
                uint32_t idx = S.enc[p + 2] | (S.enc[p + 3] << 8);
                uint32_t to = S.enc[p + 4] | (S.enc[p + 5] << 8);
                uint32_t from = k == K_SREF ? to - S.enc[p + 1] : (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
                uint32_t tl = to - from;
                // clip the token to the tile
                uint32_t k0 = rel < 0 ? (uint32_t) (-rel) : 0u;
                uint32_t k1 = (int) tl + rel > (int) nbytes ? (uint32_t) ((int) nbytes - rel) : tl;
                if (k0 < k1) {
                    if (k1 - k0 > 48) {
                        uint32_t qi = atomicAdd(&S.qn, 1u);
                        if (qi < 256) S.queue[qi] = (uint16_t) p;
                        else emit_ref_bytes(V, S, g, t0, rel, idx, from, k0, k1, 1, 0, err);
                    } else {
                        emit_ref_bytes(V, S, g, t0, rel, idx, from, k0, k1, 1, 0, err);
                    }
                }
                rel += (int) tl;
            }
        }
        __syncwarp();
        // long references: the whole warp copies each one; their offsets are recomputed from the strips
        {
            uint32_t qn = min(S.qn, 256u);
            for (uint32_t qi = 0; qi < qn; qi++) {
                uint32_t p = S.queue[qi];
                // decoded offset of token p: prefix of its strip owner + in-strip walk (done by every lane)
                uint32_t owner = p / strip;
                uint32_t base = __shfl_sync(0xffffffffu, inc - sum, owner);
                uint32_t d = base;
                for (uint32_t q = owner * strip; q < p; q++) {
                    uint8_t k = S.kind[q];
                    if (k == K_LIT) d += 1;
                    else if (k == K_SREF) d += S.enc[q + 1];
                    else if (k == K_BREF)
                        d += (uint32_t) (S.enc[q + 4] | (S.enc[q + 5] << 8)) - (uint32_t) (S.enc[q + 6] | (S.enc[q + 7] << 8));
                }
                int r0 = (int) d - (int) skip;
                uint8_t k = S.kind[p];
                uint32_t idx = S.enc[p + 2] | (S.enc[p + 3] << 8);
                uint32_t to = S.enc[p + 4] | (S.enc[p + 5] << 8);
                uint32_t from = k == K_SREF ? to - S.enc[p + 1] : (uint32_t) (S.enc[p + 6] | (S.enc[p + 7] << 8));
                uint32_t tl = to - from;
                uint32_t k0 = r0 < 0 ? (uint32_t) (-r0) : 0u;
                uint32_t k1 = (int) tl + r0 > (int) nbytes ? (uint32_t) ((int) nbytes - r0) : tl;
                emit_ref_bytes(V, S, g, t0, r0, idx, from, k0, k1, 32, lane, err);
            }
        }
        __syncwarp();
        // ---- 5. references into this very tile: pointer jumping in shared memory ----
        for (int round = 0; round < 16; round++) {
            bool pending = false;
            for (uint32_t r = lane; r < ((nbytes + 31) & ~31u); r += 32) {
                uint16_t s = r < nbytes ? S.src[r] : SRC_RESOLVED;
                uint16_t ss = SRC_RESOLVED;
                uint8_t sv = 0;
                if (s != SRC_RESOLVED) {
                    ss = S.src[s];
                    sv = S.out[s];
                }
                __syncwarp();
                if (s != SRC_RESOLVED) {
                    if (ss == SRC_RESOLVED) {
                        S.out[r] = sv;
                        S.src[r] = SRC_RESOLVED;
                    } else {
                        S.src[r] = ss;
                        pending = true;
                    }
                }
                __syncwarp();
            }
            if (!__any_sync(0xffffffffu, pending)) break;
        }
        // ---- 6. store the tile, publish ----
        {
            uint8_t *dst = (uint8_t *) V.loc[g] + t0;
            uint32_t head = (uint32_t) ((4 - ((uintptr_t) dst & 3)) & 3);
            if (head > nbytes) head = nbytes;
            if (lane < head) dst[lane] = S.out[lane];
            uint32_t nwords = (nbytes - head) >> 2;
            uint32_t *dw = (uint32_t *) (dst + head);
            for (uint32_t j = lane; j < nwords; j += 32) {
                const uint8_t *s = S.out + head + 4 * j;
                dw[j] = (uint32_t) s[0] | ((uint32_t) s[1] << 8) | ((uint32_t) s[2] << 16) | ((uint32_t) s[3] << 24);
            }
            uint32_t tail0 = head + 4 * nwords;
            if (tail0 + lane < nbytes) dst[tail0 + lane] = S.out[tail0 + lane];
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(V.flags + gt, V.epoch);
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
void Store::decode_records(const std::vector<uint32_t> &recs, uint8_t *d_out, const std::vector<uint64_t> &out_off) {
    if (recs.empty()) return;
    // per touched chunk: decode records [first, max requested]
    std::map<uint32_t, uint32_t> chunk_max;  // chunk first record -> max requested record
    for (uint32_t g : recs) {
        uint32_t f = h_first[g];
        auto it = chunk_max.find(f);
        if (it == chunk_max.end()) chunk_max[f] = g;
        else it->second = std::max(it->second, g);
    }
    const size_t NR = n_records();
    std::vector<uint64_t> loc;  // only entries of touched ranges are meaningful
    loc.assign(NR, 0);
    std::vector<std::pair<uint32_t, uint32_t>> dups;  // (request index, first request index of the same record)
    std::vector<int64_t> first_req(0);
    std::map<uint32_t, uint32_t> req_of;  // record -> first request index
    for (uint32_t i = 0; i < recs.size(); i++) {
        auto it = req_of.find(recs[i]);
        if (it == req_of.end()) {
            req_of[recs[i]] = i;
            loc[recs[i]] = (uint64_t) (uintptr_t) d_out + out_off[i];
        } else {
            dups.push_back({i, it->second});
        }
    }
    uint64_t scratch = 0;
    uint64_t n_work = 0;
    for (auto &cm : chunk_max)
        for (uint32_t g = cm.first; g <= cm.second; g++) {
            if (!loc[g]) scratch += (h_dec_len[g] + 15u) & ~15ull;
            n_work += div_up<uint32_t>(h_dec_len[g], TILE);
        }
    dec_scratch.reserve_discard(scratch + 16);
    uint64_t so = 0;
    std::vector<uint32_t> wt, wr;
    wt.reserve(n_work);
    wr.reserve(n_work);
    uint32_t lo_g = 0xFFFFFFFFu, hi_g = 0;
    for (auto &cm : chunk_max)
        for (uint32_t g = cm.first; g <= cm.second; g++) {
            if (!loc[g]) {
                loc[g] = (uint64_t) (uintptr_t) dec_scratch.p + so;
                so += (h_dec_len[g] + 15u) & ~15ull;
            }
            uint32_t nt = div_up<uint32_t>(h_dec_len[g], TILE);
            for (uint32_t t = 0; t < nt; t++) {
                wt.push_back(h_tile_base[g] + t);
                wr.push_back(g);
            }
            lo_g = std::min(lo_g, g);
            hi_g = std::max(hi_g, g);
        }
    dec_loc.reserve_discard(NR + 1);
    PX_CUDA(cudaMemcpyAsync(dec_loc.p + lo_g, loc.data() + lo_g, (size_t) (hi_g - lo_g + 1) * sizeof(uint64_t),
                            cudaMemcpyHostToDevice, st));
    dec_work.reserve_discard(2 * n_work + 2);
    PX_CUDA(cudaMemcpyAsync(dec_work.p, wt.data(), n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PX_CUDA(cudaMemcpyAsync(dec_work.p + n_work, wr.data(), n_work * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (dec_flags.cap < n_tiles + 1) {
        dec_flags.reserve_discard(n_tiles + 1);
        PX_CUDA(cudaMemsetAsync(dec_flags.p, 0, dec_flags.cap * sizeof(uint32_t), st));
        dec_epoch = 0;
    }
    dec_epoch++;
    dec_ctr.reserve_discard(4);
    PX_CUDA(cudaMemsetAsync(dec_ctr.p, 0, 4 * sizeof(uint32_t), st));
    DecodeView V{d_enc.p, d_enc_off.p, d_enc_len.p, d_dec_len.p, d_first.p, d_tile_base.p, d_tile_desc.p,
                 dec_loc.p, dec_flags.p, dec_epoch};
    const size_t smem = sizeof(WarpSmem) * DEC_WARPS;
    static bool attr_set = false;
    if (!attr_set) {
        PX_CUDA(cudaFuncSetAttribute(k_decode_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 2;
    uint32_t grid = (uint32_t) std::min<uint64_t>((uint64_t) sms * per_sm, div_up<uint64_t>(n_work, DEC_WARPS));
    PX_CUDA(cudaEventRecord(ev0, st));
    k_decode_tiles<<<grid, DEC_WARPS * 32, smem, st>>>(V, dec_work.p, dec_work.p + n_work, (uint32_t) n_work, dec_ctr.p);
    PX_LAUNCH_CHECK();
    launches++;
    for (auto &d : dups)
        PX_CUDA(cudaMemcpyAsync(d_out + out_off[d.first], d_out + out_off[d.second], h_dec_len[recs[d.first]],
                                cudaMemcpyDeviceToDevice, st));
    PX_CUDA(cudaEventRecord(ev1, st));
    uint32_t h_ctr[2];
    PX_CUDA(cudaMemcpyAsync(h_ctr, dec_ctr.p, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
    PX_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    PX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    last_get_ms = ms;
    if (h_ctr[1]) throw std::runtime_error("decode: kernel reported error " + std::to_string(h_ctr[1]));
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
