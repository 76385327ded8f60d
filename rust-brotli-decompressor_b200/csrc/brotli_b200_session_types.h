// brotli_b200_session_types.h -- plain structs shared by the kernels (brotli_decode_core.cuh) and the host runtime.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace brotli_b200 {

// Device-side record of one streaming session (one BrotliDecoderState): what BrotliState carries from one
// BrotliDecompressStream call to the next (src/state.rs:156-278).  The host fills the first block before every launch,
// the kernel fills the results and keeps the checkpoint: the decoder state at the last point a later launch can
// continue from -- the start of the stream, a metablock boundary, the command about to be decoded by the checked loop
// inside a compressed metablock (its prefix-code tables and context maps stay in the session's own arena), or the
// middle of an uncompressed metablock.  Positions (pos, next_flush, flushed, budget, out_cap) are relative to `out`,
// bitpos to `in`: the host slides both windows and rebases these fields.  All zero = start of stream.
struct ResumeState {
  // ---- host -> kernel ----
  const uint8_t* in;        // the stream's bytes not yet behind the checkpoint
  uint64_t in_size;
  uint8_t* out;             // output window
  uint64_t out_cap;
  uint64_t budget;          // output bytes the caller can take (DecompressStream: delivered + available_out)
  uint8_t* arena;           // ArenaLayout::kBytes owned by the session
  const uint8_t* dict;      // custom LZ77 dictionary of the session or nullptr
  uint64_t dict_size;
  uint32_t allow_large_window;
  // ---- kernel -> host ----
  int32_t code;             // BrotliDecoderErrorCode
  uint64_t decoded;         // output position reached (the flush point when at_flush)
  uint64_t used;            // input bytes consumed (whole bytes, as BrotliBitReaderUnload leaves them)
  uint64_t flushed_now;     // output the reference has handed out at ring flush points (what a fatal error leaves)
  uint32_t at_flush;        // NeedsMoreOutput at a ring flush point: `budget` is too small for position `decoded`
  uint32_t hit_cap;         // the decoder ran into out_cap: the outcome is not the reference's, repeat with a larger window
  uint32_t forced_flush_error;  // the error came out of the forced flush of NeedsMoreInput (src/decode.rs:2838-2850): the reference
                                // leaves *available_in / *next_in as they were at the start of the call
  // ---- checkpoint ----
  uint32_t kind;            // 0 start of stream, 1 metablock boundary, 2 inside a compressed metablock, 3 inside an uncompressed one
  uint64_t bitpos;
  uint64_t rbsize, next_flush, flushed;
  uint32_t pos;
  int32_t mlen;
  int32_t d0, d1, d2, d3;
  uint32_t wbits, large_window, rb_allocated, full_ring, is_last;
  // kind 2
  uint32_t nbt_l, nbt_c, nbt_d, bl_l, bl_c, bl_d, rbt_l0, rbt_l1, rbt_c0, rbt_c1, rbt_d0, rbt_d1;
  uint32_t n_lit_trees, n_dist_trees, npostfix, ndirect, dist_alphabet, dist_max_symbol, tables_used;
  uint32_t state, ins_rem, copy_len, implicit_dist, dist_ctx;
};

// One piece of a session launch's staging traffic (fresh input in, new output out, window slides).
struct SessionCopy { const uint8_t* src; uint8_t* dst; uint64_t n; };

}  // namespace brotli_b200
