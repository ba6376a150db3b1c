// Block-min trees and the nearest-smaller search over them (used by the match finder, encode.cu, and by the bulk
// build of the CritBit tree, index.cu).
#pragma once
#include "store.h"

namespace pixiu {

// ---------------------------------------------------------------------------------
// block-min trees over sa and lcp
// ---------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
k_tree_level(const uint32_t *__restrict__ in_a, const uint32_t *__restrict__ in_l, uint32_t n_in,
             uint32_t *__restrict__ out_a, uint32_t *__restrict__ out_l, uint32_t n_out) {
    uint32_t o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    uint32_t b = o * TREE_B, e = min(b + TREE_B, n_in);
    uint32_t ma = 0xFFFFFFFFu, ml = 0xFFFFFFFFu;
    for (uint32_t j = b; j < e; j++) {
        ma = min(ma, in_a[j]);
        ml = min(ml, in_l[j]);
    }
    out_a[o] = ma;
    out_l[o] = ml;
}

// first tree level from the interleaved leaves
static __global__ void __launch_bounds__(256)
k_tree_level_leaf(const uint2 *__restrict__ leaf, uint32_t n_in, uint32_t *__restrict__ out_a, uint32_t *__restrict__ out_l,
                  uint32_t n_out) {
    uint32_t o = blockIdx.x * 256 + threadIdx.x;
    if (o >= n_out) return;
    uint32_t b = o * TREE_B, e = min(b + TREE_B, n_in);
    uint32_t ma = 0xFFFFFFFFu, ml = 0xFFFFFFFFu;
    for (uint32_t j = b; j < e; j++) {
        const uint2 x = leaf[j];
        ma = min(ma, x.x);
        ml = min(ml, x.y & 0xFFFFu);
    }
    out_a[o] = ma;
    out_l[o] = ml;
}

// Generic nearest-smaller search over the block-min tree.
//  LEFT : visits j = start-1, start-2, ...      RIGHT: visits j = start+1, start+2, ...
//  stops at the first visited j with KEY[j] < thr and returns it (else -1 / n).
//  acc = min of ACC over the visited elements; INCL decides whether the found element counts.
//  The walk gives up early (returns not-found) once acc <= floor_ — the caller cannot improve.
//  KEYA: KEY is the sa tree (ACC the lcp tree); otherwise KEY is the lcp tree (ACC the sa tree).
template <bool LEFT, bool INCL, bool KEYA>
__device__ __forceinline__ int64_t tree_search(const MinTree &T, uint32_t start, uint32_t thr, uint32_t &acc,
                                               int64_t floor_ /* -1: never give up */) {
    auto KEY = [&](int lv, uint32_t j) { return KEYA ? T.a[lv][j] : T.l[lv][j]; };
    auto ACC = [&](int lv, uint32_t j) { return KEYA ? T.l[lv][j] : T.a[lv][j]; };
    // level 0: key and accumulated value of one entry (one 8-byte load when the leaves are interleaved)
    auto LEAF = [&](uint32_t j, uint32_t &k, uint32_t &a) {
        if (T.leaf) {
            const uint2 x = T.leaf[j];
            k = KEYA ? x.x : (x.y & 0xFFFFu);   // (the upper half of .y carries dist[sa], see MinTree::leaf)
            a = KEYA ? (x.y & 0xFFFFu) : x.x;
        } else {
            k = KEY(0, j);
            a = ACC(0, j);
        }
    };
    int lv = 0;
    int64_t pos = start;  // in units of level lv; entries beyond pos (in walk direction) are unvisited
    const int64_t NOTFOUND = LEFT ? -1 : (int64_t) T.size[0];
    int64_t hit = -1;
    // ---- ascend ----
    while (true) {
        bool found = false;
        if (LEFT) {
            while (pos % TREE_B != 0) {
                pos--;
                if (lv == 0) {
                    uint32_t k, a;
                    LEAF((uint32_t) pos, k, a);
                    if (INCL) acc = min(acc, a);
                    if (k < thr) return pos;
                    if (!INCL) acc = min(acc, a);
                } else {
                    uint32_t k = KEY(lv, (uint32_t) pos);
                    if (k < thr) { found = true; hit = pos; break; }
                    acc = min(acc, ACC(lv, (uint32_t) pos));
                }
                if ((int64_t) acc <= floor_) return NOTFOUND;
            }
            if (found) break;
            if (pos == 0) return NOTFOUND;
            pos /= TREE_B;
        } else {
            while ((pos + 1) % TREE_B != 0 && pos + 1 < (int64_t) T.size[lv]) {
                pos++;
                if (lv == 0) {
                    uint32_t k, a;
                    LEAF((uint32_t) pos, k, a);
                    if (INCL) acc = min(acc, a);
                    if (k < thr) return pos;
                    if (!INCL) acc = min(acc, a);
                } else {
                    uint32_t k = KEY(lv, (uint32_t) pos);
                    if (k < thr) { found = true; hit = pos; break; }
                    acc = min(acc, ACC(lv, (uint32_t) pos));
                }
                if ((int64_t) acc <= floor_) return NOTFOUND;
            }
            if (found) break;
            if (pos + 1 >= (int64_t) T.size[lv]) return NOTFOUND;
            pos /= TREE_B;
        }
        lv++;
        if (lv >= T.nlev) return NOTFOUND;
    }
    // ---- descend into entry `hit` of level lv (lv >= 1): the answer is inside ----
    while (lv > 0) {
        int64_t b = hit * TREE_B, e = b + TREE_B < (int64_t) T.size[lv - 1] ? b + TREE_B : (int64_t) T.size[lv - 1];
        lv--;
        bool found = false;
        if (LEFT) {
            for (int64_t c = e - 1; c >= b; c--) {
                if (lv == 0) {
                    uint32_t k, a;
                    LEAF((uint32_t) c, k, a);
                    if (INCL) acc = min(acc, a);
                    if (k < thr) return c;
                    if (!INCL) acc = min(acc, a);
                } else {
                    uint32_t k = KEY(lv, (uint32_t) c);
                    if (k < thr) { hit = c; found = true; break; }
                    acc = min(acc, ACC(lv, (uint32_t) c));
                }
                if ((int64_t) acc <= floor_) return NOTFOUND;
            }
        } else {
            for (int64_t c = b; c < e; c++) {
                if (lv == 0) {
                    uint32_t k, a;
                    LEAF((uint32_t) c, k, a);
                    if (INCL) acc = min(acc, a);
                    if (k < thr) return c;
                    if (!INCL) acc = min(acc, a);
                } else {
                    uint32_t k = KEY(lv, (uint32_t) c);
                    if (k < thr) { hit = c; found = true; break; }
                    acc = min(acc, ACC(lv, (uint32_t) c));
                }
                if ((int64_t) acc <= floor_) return NOTFOUND;
            }
        }
        if (!found) return NOTFOUND;  // cannot happen: the block minimum promised a hit
    }
    return NOTFOUND;
}

}  // namespace pixiu
