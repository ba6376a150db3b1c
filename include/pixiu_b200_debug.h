/* Test hooks of libpixiu_b200.so — used by tests/ only, not part of the drop-in surface. */
#ifndef PIXIU_B200_DEBUG_H
#define PIXIU_B200_DEBUG_H
#include "pixiu_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* GPU radix sort (radix_sort.cuh) of host (key,value) pairs on bits [0,end_bit); vals == NULL sorts indices */
int pixiu_debug_sort_pairs(int device, uint64_t *keys, uint32_t *vals, int64_t n, int end_bit, uint32_t *vals_out);
/* copy an internal array of the last encode of the open window (sa, rank, lcp, reach, off, prevp, nextp: u32;
 * text, flagp, flagc: u8; dist: u16); returns the window length */
int64_t pixiu_debug_window_array(pixiu_store *s, const char *name, void *out, int64_t cap_bytes);
/* cudaMemcpy for tests that play the collective of the multi-GPU mode: kind 1 = D2H, 2 = H2D */
int pixiu_debug_memcpy(void *dst, const void *src, int64_t bytes, int kind);
/* MemPool::nth / used_num the reference would show for the open window (PIXIU_ROTATE_REFERENCE only) */
int pixiu_debug_pool_state(pixiu_store *s, int32_t *nth, int32_t *used);
/* inner nodes visited by each key's CritBit walk (host index; feeds the algorithmic-bytes figure of the lookup bench) */
int pixiu_debug_index_depth(pixiu_store *s, int64_t n, const uint8_t *keys, const int64_t *key_off, int32_t *depth);
/* change a tuning / test knob of a live store ("dec_arena_limit", "piece_cap", "sleep_after", "sleep_ns", "trace", ...;
 * the environment variables PIXIU_<NAME> set the same knobs when the store is created) */
int pixiu_debug_set_knob(pixiu_store *s, const char *name, int64_t value);
/* pending (polled) pieces and drain passes of the last decode call */
int pixiu_debug_decode_counters(pixiu_store *s, int64_t *pending_pieces, int64_t *drains);
#ifdef __cplusplus
}
#endif
#endif
