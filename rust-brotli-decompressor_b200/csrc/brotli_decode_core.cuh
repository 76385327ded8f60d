// brotli_decode_core.cuh -- warp-per-stream Brotli decoder for sm_100a (B200).
//
// One warp decodes one stream.  The serial part of the format (bit window -> prefix-code
// symbol -> command) is executed warp-uniformly: every lane holds the same decoder state in
// registers, table lookups are same-address (broadcast) loads, so no shuffles are needed to
// distribute (insert_len, copy_len, distance).  The data-parallel part -- LZ77 backreference
// copies, literal/uncompressed runs, dictionary words, table replication -- is lane-strided.
// The output buffer IS the sliding window (no ring buffer, no second copy of every byte);
// the reference's ring-buffer flush points are emulated arithmetically because they decide
// `decoded_size` on errors.
//
// Reference path restated here (file:line under /root/reference):
//   bit window        src/bit_reader/mod.rs:135-338        -> struct BitReader
//   symbol decode     src/decode.rs:377-398                -> decode_symbol
//   prefix codes      src/decode.rs:516-1013, src/huffman/mod.rs:196-471 -> read_huffman_code
//   context maps      src/decode.rs:1096-1128,1272-1428    -> decode_context_map
//   block switching   src/decode.rs:1469-1658              -> switch_*_block
//   command loop      src/decode.rs:2330-2744              -> process_commands<SAFE>
//   distance          src/decode.rs:2017-2131              -> read_distance
//   dictionary        src/decode.rs:2593-2640, src/transform.rs:720-795 -> emit_dictionary_word
//   stream driver     src/decode.rs:2779-3403, :152-372    -> decode_stream
//
// The same source also compiles for the host with warp width 1 (BROTLI_B200_HOSTSIM); that
// build exists only so tests/ can exercise this logic in the GPU-less dev container.  It is
// never part of the product library.
#pragma once
#include <stdint.h>
#include <stddef.h>

#include "brotli_b200_session_types.h"

#if defined(BROTLI_B200_HOSTSIM)
#include <string.h>
#define BD_DEV inline
#define BD_COLD static
#define BD_CONST_TABLE static const
struct uint2 { uint32_t x, y; };
namespace brotli_b200 {
namespace hw {
constexpr uint32_t kWarp = 1;
static inline uint32_t lane() { return 0; }
static inline void syncwarp() {}
static inline bool all(bool p) { return p; }
static inline uint32_t funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
  s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t brev(uint32_t v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
  return (v >> 16) | (v << 16);
}
static inline uint32_t ldg32(const uint32_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint8_t ldg8(const uint8_t* p) { return *p; }
// "shared memory" references are plain addresses on the host
typedef uintptr_t sref_t;
static inline sref_t to_sref(const void* p) { return (sref_t)p; }
static inline uint32_t lds8(sref_t a) { return *(const uint8_t*)a; }
static inline uint32_t lds16(sref_t a) { return *(const uint16_t*)a; }
static inline uint32_t lds32(sref_t a) { return *(const uint32_t*)a; }
static inline uint2 lds64(sref_t a) { return *(const uint2*)a; }
static inline sref_t lds_ref(sref_t a) { return *(const sref_t*)a; }
}  // namespace hw
}  // namespace brotli_b200
#else
#define BD_DEV __device__ __forceinline__
#define BD_COLD static __device__ __noinline__  /* per-metablock / per-stream code: kept out of the command loop's register budget */
#define BD_CONST_TABLE __device__ const
namespace brotli_b200 {
namespace hw {
constexpr uint32_t kWarp = 32;
BD_DEV uint32_t lane() { return threadIdx.x & 31u; }
BD_DEV void syncwarp() { __syncwarp(); }
BD_DEV bool all(bool p) { return __all_sync(0xffffffffu, p); }
BD_DEV uint32_t funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
BD_DEV uint32_t brev(uint32_t v) { return __brev(v); }
BD_DEV uint32_t ldg32(const uint32_t* p) { return __ldg(p); }
BD_DEV uint8_t ldg8(const uint8_t* p) { return __ldg(p); }
// 32-bit shared-state-space addresses: ld.shared with 32-bit address arithmetic instead of generic
// 64-bit loads.  The asm statements are not volatile so the compiler may schedule them freely; every
// address data-depends on bits read after the tables were written.
typedef uint32_t sref_t;
BD_DEV sref_t to_sref(const void* p) { return (sref_t)__cvta_generic_to_shared(p); }
BD_DEV uint32_t lds8(sref_t a) { uint32_t v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint32_t lds16(sref_t a) { uint32_t v; asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint32_t lds32(sref_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint2 lds64(sref_t a) { uint2 v; asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
BD_DEV sref_t lds_ref(sref_t a) { return lds32(a); }
}  // namespace hw
}  // namespace brotli_b200
#endif

namespace brotli_b200 {

// ---- RFC 7932 constant tables (generated; see tables/gen_tables.py) ----
namespace tbl {
#define BROTLI_TABLE_ATTR BD_CONST_TABLE
#include "../../tables/brotli_tables.h"
#undef BROTLI_TABLE_ATTR
// src/decode.rs:55-61
BD_CONST_TABLE uint8_t kCodeLengthCodeOrder[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
BD_CONST_TABLE uint8_t kCodeLengthPrefixLength[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
BD_CONST_TABLE uint8_t kCodeLengthPrefixValue[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};
}  // namespace tbl

// BrotliDecoderErrorCode, src/state.rs:22-65 (c/brotli/decode.h:69-111)
enum Status : int32_t {
  kSuccess = 1, kNeedsMoreInput = 2, kNeedsMoreOutput = 3,
  kErrExuberantNibble = -1, kErrReserved = -2, kErrExuberantMetaNibble = -3, kErrSimpleHuffmanAlphabet = -4,
  kErrSimpleHuffmanSame = -5, kErrClSpace = -6, kErrHuffmanSpace = -7, kErrContextMapRepeat = -8,
  kErrBlockLength1 = -9, kErrBlockLength2 = -10, kErrTransform = -11, kErrDictionary = -12, kErrWindowBits = -13,
  kErrPadding1 = -14, kErrPadding2 = -15, kErrDistance = -16, kErrInvalidArguments = -20, kErrUnreachable = -31,
  // internal (never leaves process_commands)
  kNeedSafe = 100, kRetryFast = 101, kMetablockDone = 102
};

// ---- per-warp scratch arena in global memory (L1/L2 resident while a stream decodes) ----
// Table entries are u16: leaf = symbol << 4 | code_len; root pointer = sub_offset << 4 | (8 + sub_bits).
constexpr uint32_t kRootBits = 8;
constexpr uint32_t kMaxLitTable = 630;    // alphabet 256, src/huffman/mod.rs:15-25
constexpr uint32_t kMaxCmdTable = 1080;   // alphabet 704
constexpr uint32_t kMaxDistTable = 920;   // alphabet <= 544 (large window max symbol)
constexpr uint32_t kMaxBlockTypeTable = 632;  // alphabet <= 258
constexpr uint32_t kMaxBlockLenTable = 396;   // alphabet 26
constexpr uint32_t kMaxCtxMapTable = 646;     // alphabet <= 272
constexpr uint32_t kMaxAlphabet = 1152;

struct ArenaLayout {
  static constexpr size_t kCtxMapLit = 0;                               // u8[16384]
  static constexpr size_t kCtxMapDist = kCtxMapLit + 16384;             // u8[1024]
  static constexpr size_t kCtxModes = kCtxMapDist + 1024;               // u8[256]
  static constexpr size_t kMtf = kCtxModes + 256;                       // u8[256]
  static constexpr size_t kCodeLen = kMtf + 256;                        // u8[kMaxAlphabet]
  static constexpr size_t kSorted = kCodeLen + kMaxAlphabet;            // u16[kMaxAlphabet]
  static constexpr size_t kTreeOffLit = kSorted + 2 * kMaxAlphabet;     // u32[256]
  static constexpr size_t kTreeOffCmd = kTreeOffLit + 1024;             // u32[256]
  static constexpr size_t kTreeOffDist = kTreeOffCmd + 1024;            // u32[256]
  static constexpr size_t kBlockTypeTrees = kTreeOffDist + 1024;        // u16[3][632]
  static constexpr size_t kBlockLenTrees = kBlockTypeTrees + 2 * 3 * kMaxBlockTypeTable;  // u16[3][396]
  static constexpr size_t kCtxMapTree = kBlockLenTrees + 2 * 3 * kMaxBlockLenTable;       // u16[646]
  static constexpr size_t kPtrLit = (kCtxMapTree + 2 * kMaxCtxMapTable + 63) & ~size_t(63);  // const u16*[256]
  static constexpr size_t kPtrCmd = kPtrLit + 256 * 8;                  // const u16*[256]
  static constexpr size_t kPtrDist = kPtrCmd + 256 * 8;                 // const u16*[256]
  static constexpr size_t kTables = (kPtrDist + 256 * 8 + 63) & ~size_t(63);
  static constexpr size_t kTableEntries = 256 * (size_t)(kMaxLitTable + kMaxCmdTable + kMaxDistTable);
  static constexpr size_t kBytes = (kTables + 2 * kTableEntries + 255) & ~size_t(255);
};

// Small per-warp scratch that wants dynamic indexing; lives in shared memory on the GPU.
struct WarpScratch {
  uint16_t cl_tab[32];    // code-length code lookup: value << 8 | bits
  uint16_t count[16];     // histogram of code lengths
  uint16_t offs[16];
  uint8_t cl_cl[18];      // code length code lengths
  uint8_t pad[2];
  uint8_t word[72];       // dictionary word staging (<= 5 + 24 + 8 bytes, + uppercase overrun)
};

// Per-warp region of shared memory: everything the command loop touches per symbol.  A metablock's
// prefix-code tables are built in the global arena and then promoted here while they fit (insert&copy
// and distance tables first -- every command needs them -- then literal tables); what does not fit
// is decoded straight from the arena through the same generic pointers.
struct WarpShared {
  static constexpr uint32_t kLitMapBytes = 256;   // literal context map of <= 4 block types
  static constexpr uint32_t kDistMapBytes = 64;   // distance context map of <= 16 block types
  static constexpr uint32_t kLitPtrs = 32, kDistPtrs = 16, kCmdPtrs = 4, kDistLut = 64;
  WarpScratch ws;
  uint8_t lit_map[kLitMapBytes];
  uint8_t dist_map[kDistMapBytes];
  const uint16_t* lit_ptrs[kLitPtrs];
  const uint16_t* dist_ptrs[kDistPtrs];
  const uint16_t* cmd_ptrs[kCmdPtrs];
  const uint16_t* cur_dist[4];  // distance tables of the current distance block type, by distance context
  // shared-space mirrors used by process_commands_shared (valid when Decoder::all_shared)
  hw::sref_t cur_dist_s[4];
  hw::sref_t lit_s[kLitPtrs];
  uint32_t dist_lut[kDistLut];  // distance symbol 16 + i -> (distance base << 5) | extra bits, see build_dist_lut
  // uint16_t tables[] follow (size chosen at launch)
};

// Read-only lookup data shared by all warps of a CTA (shared memory on the GPU).
struct SharedLuts {
  const uint2* cmd_lut;     // [704] .x = insert_base | insert_extra<<16 | ctx<<24 | implicit<<26 ; .y = copy_base | copy_extra<<16
  const uint8_t* ctx_lut;   // [2048]
  const uint8_t* dictionary;
};

BD_DEV uint2 pack_cmd_lut(uint32_t code) {
  const tbl::BrotliCmdLutElement& e = tbl::kBrotliCmdLut[code];
  uint2 r;
  r.x = (uint32_t)e.insert_len_offset | ((uint32_t)e.insert_len_extra_bits << 16) | ((uint32_t)e.context << 24) |
        ((e.distance_code == 0 ? 1u : 0u) << 26);
  r.y = (uint32_t)e.copy_len_offset | ((uint32_t)e.copy_len_extra_bits << 16);
  return r;
}

// ======================= bit reader =======================
// 96 bits of look-ahead (lo, hi, nx) over 32-bit aligned words of the input; bp = bits of `lo`
// already consumed.  peek() always yields 32 valid bits.  Bits past the end of the stream read
// as zero and the position keeps counting, so truncation is detected by position, not by
// special-casing every read (same observable behaviour as the reference's byte-wise "safe"
// reader, src/bit_reader/mod.rs:228-242,362-374).
struct BitReader {
  const uint32_t* w;
  const uint8_t* bytes;
  uint32_t lo, hi, nx;
  uint32_t k;           // word index of `lo`
  uint32_t bp;          // 0..31
  uint32_t first_full;  // words [first_full, end_full) lie wholly inside the stream
  uint32_t end_full;
  uint32_t lead;        // bytes between the aligned base and the first stream byte (0..3)
  uint64_t end_byte;    // lead + stream size

  BD_DEV uint32_t load_checked(uint32_t j) const {
    if (j >= first_full && j < end_full) return hw::ldg32(w + j);
    uint32_t v = 0;
    uint64_t b0 = (uint64_t)j * 4;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      uint64_t bi = b0 + b;
      if (bi >= lead && bi < end_byte) v |= (uint32_t)hw::ldg8(bytes + bi) << (8 * b);
    }
    return v;
  }
  BD_DEV void seek_byte(uint64_t byte_off) {
    uint64_t a = lead + byte_off;
    k = (uint32_t)(a >> 2);
    bp = (uint32_t)(a & 3) * 8;
    lo = load_checked(k); hi = load_checked(k + 1); nx = load_checked(k + 2);
  }
  BD_DEV void init(const uint8_t* p, uint64_t size) {
    uintptr_t a = (uintptr_t)p;
    lead = (uint32_t)(a & 3);
    bytes = p - lead;
    w = (const uint32_t*)bytes;
    end_byte = lead + size;
    first_full = lead ? 1u : 0u;
    end_full = (uint32_t)(end_byte >> 2);
    if (end_full < first_full) end_full = first_full;
    seek_byte(0);
  }
  BD_DEV uint32_t peek() const { return hw::funnelshift_r(lo, hi, bp); }
  template <bool FAST>
  BD_DEV void skip(uint32_t n) {  // n <= 32
    bp += n;
    if (bp >= 32) {
      lo = hi; hi = nx; k++;
      nx = FAST ? hw::ldg32(w + k + 2) : load_checked(k + 2);
      bp -= 32;
    }
  }
  template <bool FAST>
  BD_DEV uint32_t read(uint32_t n) {  // n <= 31
    uint32_t v = peek() & ((1u << n) - 1u);
    skip<FAST>(n);
    return v;
  }
  BD_DEV uint64_t bitpos() const { return ((uint64_t)k << 5) + bp; }
  BD_DEV bool overrun() const { return bitpos() > end_byte * 8; }
  BD_DEV uint64_t byte_pos() const { return (bitpos() >> 3) - lead; }  // valid when byte aligned
  BD_DEV uint64_t size() const { return end_byte - lead; }
  // src/bit_reader/mod.rs:378-385; false if the padding bits are not zero
  BD_DEV bool jump_to_byte_boundary() {
    uint32_t pad = (8u - (bp & 7u)) & 7u;
    return pad == 0 || read<false>(pad) == 0;
  }
  // true while the unchecked fast path may run: >= `words` whole words remain beyond `nx`
  BD_DEV bool fast_ok(uint32_t words) const { return k + 3 + words <= end_full; }
};

// ======================= prefix-code symbol decode (src/decode.rs:377-398) =======================
BD_DEV uint32_t decode_symbol(const uint16_t* tab, uint32_t bits, uint32_t& len) {
  uint32_t e = tab[bits & 0xFFu];
  uint32_t l = e & 15u;
  if (l > kRootBits) {
    uint32_t wbits = l - kRootBits;
    e = tab[(e >> 4) + ((bits >> kRootBits) & ((1u << wbits) - 1u))];
    l = kRootBits + (e & 15u);
  }
  len = l;
  return e >> 4;
}

// ======================= decoder state (warp-uniform) =======================
enum CmdState : uint32_t { kCmdBegin = 0, kCmdInner = 1, kCmdPostLiterals = 2 };

struct Decoder {
  BitReader br;
  uint8_t* out;
  uint32_t cap;           // output capacity (clamped to < 2^32 - 64)
  uint32_t pos;           // bytes produced so far
  int32_t mlen;           // meta_block_remaining_len
  uint32_t wbits;
  uint32_t max_backward;  // (1 << wbits) - 16
  uint32_t large_window;  // stream carries the large-window marker (src/decode.rs:152-187)
  // custom LZ77 dictionary (BrotliState::new_with_custom_dictionary, src/state.rs:400-411): bytes that logically
  // precede the output.  cdict points at the part the window can reach (its last cdict_size bytes,
  // src/decode.rs:1831-1838); cdict_limit = max_backward - the UNCLIPPED length (src/decode.rs:2954-2955).
  const uint8_t* cdict;
  uint32_t cdict_size;
  int64_t cdict_limit;
  uint32_t cdict_given;
  // ring-buffer flush emulation (src/decode.rs:1693-1738,1808-1871)
  uint32_t rb_allocated;
  uint64_t rbsize;
  uint64_t next_flush;    // absolute position of the next flush event
  uint64_t flushed;       // bytes the reference would have handed to the caller before a fatal error
  uint64_t discarded;     // literals decoded past the capacity while a command overshoots MLEN (SAFE only)
  uint32_t full_ring;
  // streaming sessions / exact re-decodes: the caller can take `budget` output bytes in total; a ring-buffer flush point
  // beyond it stops the decoder there with NeedsMoreOutput, as WriteRingBuffer does (src/decode.rs:1693-1738)
  uint64_t budget;
  uint32_t at_flush;      // stopped at next_flush for lack of budget
  uint32_t is_last;       // ISLAST of the current metablock
  uint32_t tables_used;   // entries of the arena's table space the current metablock uses
  // last four distances, d0 most recent (src/state.rs:295-296)
  int32_t d0, d1, d2, d3;
  // block-split state per category: literal, command, distance (src/state.rs:146-154)
  uint32_t nbt_l, nbt_c, nbt_d;
  uint32_t bl_l, bl_c, bl_d;
  uint32_t rbt_l0, rbt_l1, rbt_c0, rbt_c1, rbt_d0, rbt_d1;
  uint32_t n_lit_trees, n_dist_trees;
  uint32_t npostfix, ndirect;     // ndirect includes the 16 short codes
  uint32_t dist_alphabet, dist_max_symbol;
  // current trees
  const uint16_t* cmd_tree;
  const uint16_t* lit_tree;       // valid when trivial_ctx
  uint32_t trivial_ctx;
  uint32_t ctx_slice;             // literal block type << 6
  uint32_t ctx_mode_off;          // mode * 512 into ctx_lut
  uint32_t dist_slice;            // distance block type << 2
  // command in flight (for the fast -> safe hand-over)
  uint32_t state;
  uint32_t ins_rem;               // literals still to emit
  uint32_t copy_len;
  uint32_t implicit_dist;         // 1: reuse last distance, no symbol
  uint32_t dist_ctx;
  // arena views
  uint8_t* arena;
  uint16_t* tables;
  WarpScratch* ws;
  SharedLuts luts;
  // shared-memory residency of the current metablock's tables
  WarpShared* sh;
  uint16_t* stab;                 // shared table storage
  uint32_t stab_cap, stab_used;   // in entries
  const uint8_t* map_lit;         // literal context map: sh->lit_map or the arena copy
  const uint8_t* map_dist;
  const uint16_t* const* lit_ptrs;   // tree index -> table
  const uint16_t* const* cmd_ptrs;
  const uint16_t* const* dist_ptrs;
  uint32_t n_unpromoted;          // tables of this metablock left in the global arena
  uint32_t all_shared;            // every table, map and pointer array of this metablock is in shared memory
  uint32_t use_dist_lut;          // sh->dist_lut covers the whole distance alphabet

  BD_DEV uint8_t* ctx_map_lit() const { return arena + ArenaLayout::kCtxMapLit; }
  BD_DEV uint8_t* ctx_map_dist() const { return arena + ArenaLayout::kCtxMapDist; }
  BD_DEV uint8_t* ctx_modes() const { return arena + ArenaLayout::kCtxModes; }
  BD_DEV uint32_t* tree_off_lit() const { return (uint32_t*)(arena + ArenaLayout::kTreeOffLit); }
  BD_DEV uint32_t* tree_off_cmd() const { return (uint32_t*)(arena + ArenaLayout::kTreeOffCmd); }
  BD_DEV uint32_t* tree_off_dist() const { return (uint32_t*)(arena + ArenaLayout::kTreeOffDist); }
  BD_DEV uint16_t* block_type_tree(int cat) const { return (uint16_t*)(arena + ArenaLayout::kBlockTypeTrees) + cat * kMaxBlockTypeTable; }
  BD_DEV uint16_t* block_len_tree(int cat) const { return (uint16_t*)(arena + ArenaLayout::kBlockLenTrees) + cat * kMaxBlockLenTable; }
  BD_DEV uint16_t* ctx_map_tree() const { return (uint16_t*)(arena + ArenaLayout::kCtxMapTree); }
};

BD_DEV uint32_t bit_width(uint32_t x) { uint32_t r = 0; while (x) { x >>= 1; r++; } return r; }  // Log2Floor, src/decode.rs:502-509

// ======================= prefix-code descriptions -> lookup tables =======================
// Builds the 2-level table for code lengths clen[0..n) (count[] = histogram, complete code).
// Same shape as BrotliBuildHuffmanTable (src/huffman/mod.rs:273-386): 8-bit root, 2nd-level
// tables as wide as the longest code under their root prefix.  The serial walk over the sorted
// symbols is warp-uniform; the replication of each entry is lane-strided.
BD_DEV int build_table(Decoder& d, uint16_t* tab, uint32_t cap_entries, uint32_t n, uint32_t& table_size) {
  const uint32_t lane = hw::lane();
  const uint8_t* clen = d.arena + ArenaLayout::kCodeLen;
  uint16_t* sorted = (uint16_t*)(d.arena + ArenaLayout::kSorted);
  uint16_t* count = d.ws->count;
  uint16_t* offs = d.ws->offs;
  uint32_t max_len = 0, acc = 0;
  for (uint32_t l = 1; l <= 15; l++) { offs[l] = (uint16_t)acc; acc += count[l]; if (count[l]) max_len = l; }
  for (uint32_t s = 0; s < n; s++) {  // counting sort by (length, symbol)
    uint32_t l = clen[s];
    if (l) { sorted[offs[l]] = (uint16_t)s; offs[l]++; }
  }
  uint32_t code = 0, idx = 0;
  const uint32_t root_len = max_len < kRootBits ? max_len : kRootBits;
  for (uint32_t l = 1; l <= root_len; l++) {
    const uint32_t reps = 256u >> l;
    for (uint32_t j = count[l]; j != 0; j--) {
      const uint32_t e = ((uint32_t)sorted[idx++] << 4) | l;
      const uint32_t rev = hw::brev(code) >> (32 - l);
      for (uint32_t r = lane; r < reps; r += hw::kWarp) tab[rev + (r << l)] = (uint16_t)e;
      code++;
    }
    code <<= 1;
  }
  uint32_t next_free = 256;
  if (max_len > kRootBits) {
    uint32_t cur_prefix = 0xFFFFFFFFu, sub_off = 0, sub_bits = 0;
    for (uint32_t l = kRootBits + 1; l <= max_len; l++) {
      const uint32_t sl = l - kRootBits;
      while (count[l] != 0) {
        const uint32_t prefix = code >> sl;
        if (prefix != cur_prefix) {
          // NextTableBitSize, src/huffman/mod.rs:181-193
          uint32_t len2 = l; int32_t left = 1 << (len2 - kRootBits);
          while (len2 < 15) { left -= count[len2]; if (left <= 0) break; len2++; left <<= 1; }
          sub_bits = len2 - kRootBits;
          sub_off = next_free;
          next_free += 1u << sub_bits;
          if (next_free > cap_entries) return kErrUnreachable;
          cur_prefix = prefix;
          tab[hw::brev(prefix) >> 24] = (uint16_t)((sub_off << 4) | (kRootBits + sub_bits));
        }
        const uint32_t e = ((uint32_t)sorted[idx++] << 4) | sl;
        const uint32_t rev = hw::brev(code & ((1u << sl) - 1u)) >> (32 - sl);
        const uint32_t reps = 1u << (sub_bits - sl);
        for (uint32_t r = lane; r < reps; r += hw::kWarp) tab[sub_off + rev + (r << sl)] = (uint16_t)e;
        code++;
        count[l]--;
      }
      code <<= 1;
    }
  }
  table_size = next_free;
  hw::syncwarp();
  return kSuccess;
}

// ReadHuffmanCode, src/decode.rs:868-1013 (+ :516-556, :565-658, :661-853).  Truncation is
// detected by the caller through br.overrun(); reads past the end see zero bits, every loop
// below still terminates and stays inside its arrays.
BD_COLD int read_huffman_code(Decoder& d, uint32_t alphabet_size, uint32_t max_symbol, uint16_t* tab, uint32_t cap_entries,
                             uint32_t& table_size) {
  BitReader& br = d.br;
  const uint32_t lane = hw::lane();
  alphabet_size &= 0x7ff;
  const uint32_t hskip = br.read<false>(2);
  if (hskip == 1) {  // simple code: 1..4 symbols
    const uint32_t nsym = br.read<false>(2) + 1;
    const uint32_t max_bits = bit_width(alphabet_size - 1);
    uint32_t v[4] = {0, 0, 0, 0};
#pragma unroll
    for (uint32_t i = 0; i < 4; i++) {
      if (i < nsym) {
        v[i] = br.read<false>(max_bits);
        if (v[i] >= max_symbol) return br.overrun() ? kNeedsMoreInput : kErrSimpleHuffmanAlphabet;
      }
    }
#pragma unroll
    for (uint32_t i = 0; i < 3; i++)
#pragma unroll
      for (uint32_t j = i + 1; j < 4; j++)
        if (j < nsym && v[i] == v[j]) return br.overrun() ? kNeedsMoreInput : kErrSimpleHuffmanSame;
    uint32_t tree_select = 0;
    if (nsym == 4) tree_select = br.read<false>(1);
    // BrotliBuildSimpleHuffmanTable, src/huffman/mod.rs:390-471 (pattern of period <= 8, replicated)
    uint32_t p[8];
    uint32_t period;
#define BD_SWAP(a, b) { uint32_t t_ = a; a = b; b = t_; }
    if (nsym == 1) {
      period = 1; p[0] = v[0] << 4;
    } else if (nsym == 2) {
      if (v[1] < v[0]) BD_SWAP(v[0], v[1]);
      period = 2; p[0] = (v[0] << 4) | 1; p[1] = (v[1] << 4) | 1;
    } else if (nsym == 3) {
      if (v[2] < v[1]) BD_SWAP(v[1], v[2]);
      period = 4; p[0] = p[2] = (v[0] << 4) | 1; p[1] = (v[1] << 4) | 2; p[3] = (v[2] << 4) | 2;
    } else if (!tree_select) {
      for (int i = 0; i < 3; i++) for (int j = i + 1; j < 4; j++) if (v[j] < v[i]) BD_SWAP(v[i], v[j]);
      period = 4; p[0] = (v[0] << 4) | 2; p[2] = (v[1] << 4) | 2; p[1] = (v[2] << 4) | 2; p[3] = (v[3] << 4) | 2;
    } else {
      if (v[3] < v[2]) BD_SWAP(v[2], v[3]);
      period = 8;
      p[0] = p[2] = p[4] = p[6] = (v[0] << 4) | 1;
      p[1] = p[5] = (v[1] << 4) | 2; p[3] = (v[2] << 4) | 3; p[7] = (v[3] << 4) | 3;
    }
#undef BD_SWAP
    if (256 > cap_entries) return kErrUnreachable;
    for (uint32_t i = lane; i < 256; i += hw::kWarp) {
      uint32_t q = i & (period - 1), e = p[0];
#pragma unroll
      for (uint32_t t = 1; t < 8; t++) if (q == t) e = p[t];
      tab[i] = (uint16_t)e;
    }
    table_size = 256;
    hw::syncwarp();
    return kSuccess;
  }
  // ---- complex code: code-length code lengths (src/decode.rs:801-853) ----
  WarpScratch& ws = *d.ws;
  for (uint32_t i = 0; i < 16; i++) ws.count[i] = 0;
  for (uint32_t i = 0; i < 18; i++) ws.cl_cl[i] = 0;
  uint32_t space = 32, num_codes = 0, last_sym = 0;
  for (uint32_t i = hskip; i < 18; i++) {
    const uint32_t ix = br.peek() & 15u;
    const uint32_t v = tbl::kCodeLengthPrefixValue[ix];
    br.skip<false>(tbl::kCodeLengthPrefixLength[ix]);
    const uint32_t sym = tbl::kCodeLengthCodeOrder[i];
    ws.cl_cl[sym] = (uint8_t)v;
    if (v != 0) {
      space -= 32u >> v; num_codes++; ws.count[v]++; last_sym = sym;
      if (space - 1u >= 32u) break;
    }
  }
  if (!(num_codes == 1 || space == 0)) return br.overrun() ? kNeedsMoreInput : kErrClSpace;
  // BrotliBuildCodeLengthsHuffmanTable, src/huffman/mod.rs:196-271
  if (num_codes == 1) {
    for (uint32_t i = 0; i < 32; i++) ws.cl_tab[i] = (uint16_t)(last_sym << 8);
  } else {
    uint32_t code = 0;
    for (uint32_t l = 1; l <= 5; l++) {
      for (uint32_t s = 0; s < 18; s++) {
        if (ws.cl_cl[s] == l) {
          const uint32_t rev = hw::brev(code) >> (32 - l);
          for (uint32_t r = rev; r < 32; r += 1u << l) ws.cl_tab[r] = (uint16_t)((s << 8) | l);
          code++;
        }
      }
      code <<= 1;
    }
  }
  // ---- symbol code lengths (src/decode.rs:565-731) ----
  uint8_t* clen = d.arena + ArenaLayout::kCodeLen;
  for (uint32_t i = 0; i < 16; i++) ws.count[i] = 0;
  for (uint32_t i = lane; i < max_symbol; i += hw::kWarp) clen[i] = 0;
  hw::syncwarp();
  uint32_t symbol = 0, prev_code_len = 8, repeat = 0, repeat_code_len = 0;
  space = 32768;
  while (symbol < max_symbol && space > 0) {
    const uint32_t bits = br.peek();
    const uint32_t p = ws.cl_tab[bits & 31u];
    const uint32_t code_len = p >> 8;
    if (code_len < 16) {
      br.skip<false>(p & 0xFF);
      repeat = 0;
      if (code_len != 0) {
        clen[symbol] = (uint8_t)code_len;
        prev_code_len = code_len;
        space -= 32768u >> code_len;
        ws.count[code_len]++;
      }
      symbol++;
    } else {  // ProcessRepeatedCodeLength, src/decode.rs:600-658
      const uint32_t extra_bits = code_len - 14;
      uint32_t repeat_delta = (bits >> (p & 0xFF)) & ((1u << extra_bits) - 1u);
      br.skip<false>((p & 0xFF) + extra_bits);
      const uint32_t new_len = code_len == 16 ? prev_code_len : 0;
      if (repeat_code_len != new_len) { repeat = 0; repeat_code_len = new_len; }
      const uint32_t old_repeat = repeat;
      if (repeat > 0) { repeat -= 2; repeat <<= extra_bits; }
      repeat += repeat_delta + 3;
      repeat_delta = repeat - old_repeat;
      if (symbol + repeat_delta > max_symbol) { symbol = max_symbol; space = 0xFFFFF; break; }
      if (repeat_code_len != 0) {
        for (uint32_t r = lane; r < repeat_delta; r += hw::kWarp) clen[symbol + r] = (uint8_t)repeat_code_len;
        space -= repeat_delta << (15 - repeat_code_len);
        ws.count[repeat_code_len] = (uint16_t)(ws.count[repeat_code_len] + repeat_delta);
      }
      symbol += repeat_delta;
    }
  }
  if (space != 0) return br.overrun() ? kNeedsMoreInput : kErrHuffmanSpace;
  hw::syncwarp();
  return build_table(d, tab, cap_entries, max_symbol, table_size);
}

// ======================= small header pieces =======================
// DecodeVarLenUint8, src/decode.rs:193-241
BD_DEV uint32_t read_varlen_uint8(BitReader& br) {
  if (!br.read<false>(1)) return 0;
  const uint32_t n = br.read<false>(3);
  if (n == 0) return 1;
  return (1u << n) + br.read<false>(n);
}

// ReadBlockLength, src/decode.rs:1016-1026
template <bool FAST>
BD_DEV uint32_t read_block_length(BitReader& br, const uint16_t* tree) {
  uint32_t len;
  const uint32_t code = decode_symbol(tree, br.peek(), len);
  br.skip<FAST>(len);
  const uint32_t nbits = tbl::kBrotliBlockLengthNBits[code];
  return tbl::kBrotliBlockLengthOffset[code] + br.read<FAST>(nbits);
}

// DecodeBlockTypeAndLength, src/decode.rs:1469-1524 (caller has checked num_types > 1)
template <bool FAST>
BD_DEV void read_block_switch(Decoder& d, int cat, uint32_t num_types, uint32_t& rb0, uint32_t& rb1, uint32_t& block_length) {
  uint32_t len;
  uint32_t block_type = decode_symbol(d.block_type_tree(cat), d.br.peek(), len);
  d.br.skip<FAST>(len);
  block_length = read_block_length<FAST>(d.br, d.block_len_tree(cat));
  if (block_type == 1) block_type = rb1 + 1;
  else if (block_type == 0) block_type = rb0;
  else block_type -= 2;
  if (block_type >= num_types) block_type -= num_types;
  rb0 = rb1; rb1 = block_type;
}

// PrepareLiteralDecoding + DetectTrivialLiteralBlockTypes, src/decode.rs:1525-1570
BD_DEV void prepare_literal_decoding(Decoder& d) {
  const uint32_t block_type = d.rbt_l1;
  d.ctx_slice = block_type << 6;
  const uint8_t* map = d.map_lit + d.ctx_slice;
  const uint32_t sample = map[0];
  bool same = true;
  for (uint32_t j = hw::lane(); j < 64; j += hw::kWarp) same = same && (map[j] == sample);
  d.trivial_ctx = hw::all(same) ? 1u : 0u;
  d.lit_tree = d.lit_ptrs[sample];
  d.ctx_mode_off = (uint32_t)(d.ctx_modes()[block_type] & 3u) * 512u;
}

// InverseMoveToFrontTransform, src/decode.rs:1096-1128 (warp-uniform, serial)
BD_COLD void inverse_move_to_front(Decoder& d, uint8_t* v, uint32_t n) {
  uint8_t* mtf = d.arena + ArenaLayout::kMtf;
  for (uint32_t i = hw::lane(); i < 256; i += hw::kWarp) mtf[i] = (uint8_t)i;
  hw::syncwarp();
  for (uint32_t i = 0; i < n; i++) {
    uint32_t index = v[i];
    const uint8_t value = mtf[index];
    v[i] = value;
    for (; index > 0; index--) mtf[index] = mtf[index - 1];
    mtf[0] = value;
  }
  hw::syncwarp();
}

// DecodeContextMap, src/decode.rs:1272-1428
BD_COLD int decode_context_map(Decoder& d, uint32_t map_size, uint8_t* map, uint32_t& num_htrees) {
  BitReader& br = d.br;
  const uint32_t lane = hw::lane();
  num_htrees = read_varlen_uint8(br) + 1;
  for (uint32_t i = lane; i < map_size; i += hw::kWarp) map[i] = 0;
  hw::syncwarp();
  if (num_htrees <= 1) return kSuccess;
  uint32_t max_rle = 0;
  const uint32_t b5 = br.peek() & 31u;
  if (b5 & 1) { max_rle = (b5 >> 1) + 1; br.skip<false>(5); } else { br.skip<false>(1); }
  const uint32_t alphabet = num_htrees + max_rle;
  uint32_t tsize;
  int r = read_huffman_code(d, alphabet, alphabet, d.ctx_map_tree(), kMaxCtxMapTable, tsize);
  if (r != kSuccess) return r;
  const uint16_t* tree = d.ctx_map_tree();
  uint32_t idx = 0;
  while (idx < map_size) {
    uint32_t len;
    const uint32_t code = decode_symbol(tree, br.peek(), len);
    br.skip<false>(len);
    if (code == 0) { idx++; continue; }  // already zero
    if (code > max_rle) { map[idx++] = (uint8_t)(code - max_rle); continue; }
    const uint32_t reps = (1u << code) + br.read<false>(code);
    if (idx + reps > map_size) return br.overrun() ? kNeedsMoreInput : kErrContextMapRepeat;
    idx += reps;  // zeros
    if (br.overrun()) return kNeedsMoreInput;
  }
  hw::syncwarp();
  if (br.read<false>(1)) inverse_move_to_front(d, map, map_size);
  return kSuccess;
}

// ======================= output primitives =======================
// LZ77 copy of `len` bytes from `dist` back with byte-serial (overlapping) semantics,
// src/decode.rs:2641-2720.  Caller guarantees pos + len <= cap and dist <= pos.
BD_DEV void warp_copy(uint8_t* out, uint32_t pos, uint32_t dist, uint32_t len) {
  const uint32_t lane = hw::lane();
  hw::syncwarp();  // earlier stores by other lanes must be visible to the loads below
  uint8_t* dst = out + pos;
  const uint8_t* src = dst - dist;
  if (dist >= len) {  // no overlap at all
    for (uint32_t i = lane; i < len; i += hw::kWarp) dst[i] = src[i];
  } else if (dist >= hw::kWarp) {  // overlap only across 32-byte steps
    for (uint32_t base = 0; base < len; base += hw::kWarp) {
      const uint32_t i = base + lane;
      if (i < len) dst[i] = src[i];
      hw::syncwarp();
    }
  } else {  // periodic pattern: every source byte lies before pos
    for (uint32_t i = lane; i < len; i += hw::kWarp) dst[i] = src[i % dist];
  }
}

// Static dictionary word + transform (src/decode.rs:2597-2620, src/transform.rs:720-795) into ws.word.
BD_COLD uint32_t build_dictionary_word(Decoder& d, uint32_t offset, uint32_t wlen, uint32_t transform_idx) {
  uint8_t* o = d.ws->word;
  const uint8_t* dict = d.luts.dictionary + offset;
  const uint8_t* prefix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[transform_idx * 3]];
  const uint32_t t = tbl::kBrotliTransforms[transform_idx * 3 + 1];
  const uint8_t* suffix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[transform_idx * 3 + 2]];
  uint32_t idx = 0;
  while (prefix[idx]) { o[idx] = prefix[idx]; idx++; }
  int32_t len = (int32_t)wlen;
  int32_t skip = t < tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 ? 0 : (int32_t)t - (tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 - 1);
  if (skip > len) skip = len;
  dict += skip; len -= skip;
  if (t <= tbl::BROTLI_TRANSFORM_OMIT_LAST_9) len -= (int32_t)t;
  const uint32_t body = idx;
  for (int32_t i = 0; i < len; i++) o[idx++] = hw::ldg8(dict + i);
  if (t == tbl::BROTLI_TRANSFORM_UPPERCASE_FIRST || t == tbl::BROTLI_TRANSFORM_UPPERCASE_ALL) {
    // ToUpperCase, src/transform.rs:720-735.  The reference may touch up to two bytes past the
    // word; those bytes are then overwritten by the suffix / later output, so confine the
    // effect to the staging buffer (zero-filled tail is harmless).
    o[idx] = 0; o[idx + 1] = 0;
    uint32_t q = body; int32_t left = len;
    do {
      uint32_t step;
      if (o[q] < 0xc0) { if (o[q] >= 'a' && o[q] <= 'z') o[q] ^= 32; step = 1; }
      else if (o[q] < 0xe0) { o[q + 1] ^= 32; step = 2; }
      else { o[q + 2] ^= 5; step = 3; }
      q += step; left -= (int32_t)step;
    } while (t == tbl::BROTLI_TRANSFORM_UPPERCASE_ALL && left > 0);
  }
  for (uint32_t i = 0; suffix[i]; i++) o[idx++] = suffix[i];
  hw::syncwarp();
  return idx;
}

// Distance tables of the current distance block type, indexed by the 2-bit distance context
// (src/decode.rs:2186-2188: dist_htree_index = dist_context_map[slice + context]).
BD_DEV void refresh_cur_dist(Decoder& d) {
  if (hw::lane() < 4 || hw::kWarp == 1) {
    for (uint32_t c = hw::kWarp == 1 ? 0 : hw::lane(); c < 4; c += hw::kWarp == 1 ? 1 : 4)
    {
      const uint16_t* t = d.dist_ptrs[d.map_dist[d.dist_slice + c]];
      d.sh->cur_dist[c] = t;
      d.sh->cur_dist_s[c] = hw::to_sref(t);
    }
  }
  hw::syncwarp();
}

// ======================= distance (src/decode.rs:2017-2131) =======================
template <bool FAST>
BD_DEV int32_t read_distance(Decoder& d, uint32_t& push_to_ring) {
  BitReader& br = d.br;
  const uint32_t tree_idx = d.map_dist[d.dist_slice + d.dist_ctx];
  const uint16_t* tree = d.dist_ptrs[tree_idx];
  uint32_t len;
  const uint32_t code = decode_symbol(tree, br.peek(), len);
  br.skip<FAST>(len);
  d.bl_d--;
  push_to_ring = 1;
  if (code < 16) {  // TakeDistanceFromRingBuffer
    if (code == 0) { push_to_ring = 0; return d.d0; }
    // codes 1..3: d1..d3; 4..9: d0 -1,+1,-2,+2,-3,+3; 10..15: d1 likewise
    int32_t base;
    if (code < 4) return code == 1 ? d.d1 : (code == 2 ? d.d2 : d.d3);
    const uint32_t c = code - 4;
    base = c < 6 ? d.d0 : d.d1;
    const uint32_t m = c < 6 ? c : c - 6;
    const int32_t delta = (int32_t)(m >> 1) + 1;
    int32_t v = (m & 1) ? base + delta : base - delta;
    if (!(m & 1) && v <= 0) v = 0x7fffffff;  // src/decode.rs:2042-2046
    return v;
  }
  int32_t distval = (int32_t)code - (int32_t)d.ndirect;
  uint32_t dc = code;
  if (distval >= 0) {
    const uint32_t postfix = (uint32_t)distval & ((1u << d.npostfix) - 1u);
    const uint32_t hcode = (uint32_t)distval >> d.npostfix;
    const uint32_t nbits = (hcode >> 1) + 1;
    const uint32_t bits = br.read<FAST>(nbits);
    const uint32_t offset = ((2u + (hcode & 1u)) << nbits) - 4u;
    dc = ((offset + bits) << d.npostfix) + postfix + d.ndirect;
  }
  return (int32_t)(dc - 16u + 1u);
}

// ======================= command loop (src/decode.rs:2330-2744) =======================
// FAST (SAFE=false): unchecked word loads, no per-byte limits; entered only when the input has
// >= 16 whole words of slack and the command cannot reach the output capacity or a flush point.
// SAFE: bounds-aware loads, truncation checked before anything is emitted, output split at the
// capacity and at the emulated ring-buffer flush points.  SAFE handles one command and returns
// kRetryFast so that long streams drop back to the fast loop.
BD_DEV int flush_event(Decoder& d) {  // WriteRingBuffer at pos >= ringbuffer_size, src/decode.rs:1693-1738
  if (d.mlen < 0) return kErrBlockLength1;
  if (d.next_flush > d.budget) { d.at_flush = 1; return kNeedsMoreOutput; }  // num_written < to_write, :1718-1729
  d.flushed = d.next_flush;
  d.next_flush = d.full_ring ? d.next_flush + d.rbsize : ~(uint64_t)0;
  return kSuccess;
}

template <bool SAFE>
BD_COLD int process_commands(Decoder& d) {
  BitReader& br = d.br;
  const uint32_t lane = hw::lane();
  constexpr bool FAST = !SAFE;
  for (;;) {
    if (d.state == kCmdBegin) {
      if (FAST && !br.fast_ok(16)) return kNeedSafe;
      if (d.bl_c == 0) {  // DecodeCommandBlockSwitch, src/decode.rs:1609-1621
        if (d.nbt_c <= 1) return kNeedsMoreInput;  // reference quirk, src/decode.rs:2368-2373 with :1479-1481
        read_block_switch<FAST>(d, 1, d.nbt_c, d.rbt_c0, d.rbt_c1, d.bl_c);
        d.cmd_tree = d.cmd_ptrs[d.rbt_c1];
      }
      // ReadCommandInternal, src/decode.rs:2134-2189
      uint32_t len;
      const uint32_t sym = decode_symbol(d.cmd_tree, br.peek(), len);
      br.skip<FAST>(len);
      const uint2 lut = d.luts.cmd_lut[sym];
      const uint32_t ie = (lut.x >> 16) & 0xFF, ce = lut.y >> 16;
      uint32_t insert_len = lut.x & 0xFFFF;
      if (ie) insert_len += br.read<FAST>(ie);
      d.copy_len = (lut.y & 0xFFFF) + br.read<FAST>(ce);
      d.implicit_dist = (lut.x >> 26) & 1;
      d.dist_ctx = (lut.x >> 24) & 3;
      d.bl_c--;
      if (SAFE && br.overrun()) return kNeedsMoreInput;
      d.ins_rem = insert_len;
      d.mlen -= (int32_t)insert_len;
      d.state = insert_len ? kCmdInner : kCmdPostLiterals;
      if (FAST) {
        // everything this command can write must stay clear of cap and of the next flush point
        const uint64_t limit = d.next_flush < d.cap ? d.next_flush : d.cap;
        if ((uint64_t)d.pos + insert_len + (d.copy_len > 40 ? d.copy_len : 40) >= limit) return kNeedSafe;
      }
    }
    if (d.state == kCmdInner) {
      uint32_t i = d.ins_rem;
      if (d.trivial_ctx) {  // src/decode.rs:2393-2462
        do {
          if (FAST && !br.fast_ok(4)) { d.ins_rem = i; return kNeedSafe; }
          if (d.bl_l == 0) {
            if (d.nbt_l > 1) {
              read_block_switch<FAST>(d, 0, d.nbt_l, d.rbt_l0, d.rbt_l1, d.bl_l);
              prepare_literal_decoding(d);
              if (!d.trivial_ctx) break;
            } else if (SAFE) { d.ins_rem = i; return kNeedsMoreInput; }
          }
          uint32_t len;
          const uint32_t lit = decode_symbol(d.lit_tree, br.peek(), len);
          br.skip<FAST>(len);
          bool keep = true;
          if (SAFE) {
            if (br.overrun()) { d.ins_rem = i; return kNeedsMoreInput; }
            // The reference decodes into its ring buffer and only notices a full output buffer at
            // a flush; a truncated input or an MLEN overshoot (BLOCK_LENGTH_1/2) is reported first.
            // So past the capacity keep consuming this literal run without storing it.
            if (d.pos >= d.cap) keep = false;
          }
          if (keep) { if (lane == 0) d.out[d.pos] = (uint8_t)lit; }
          if (d.bl_l == 0) return kErrWindowBits;  // src/decode.rs:2434-2438
          d.bl_l--;
          if (keep) d.pos++; else d.discarded++;
          i--;
          if (SAFE && (uint64_t)d.pos + d.discarded == d.next_flush) { d.ins_rem = i; int r = flush_event(d); if (r != kSuccess) return r; }
        } while (i != 0);
      }
      if (i != 0) {  // context-dependent literals, src/decode.rs:2463-2551
        hw::syncwarp();
        uint32_t p1 = d.pos >= 1 ? d.out[d.pos - 1] : 0, p2 = d.pos >= 2 ? d.out[d.pos - 2] : 0;
        if (d.cdict_given && d.pos < 2) { p1 = 0; p2 = 0; }  // custom_dict_avoid_context_seed, src/decode.rs:2466-2476
        const uint8_t* map = d.map_lit;
        const uint16_t* const* lit_ptrs = d.lit_ptrs;
        do {
          if (FAST && !br.fast_ok(4)) { d.ins_rem = i; return kNeedSafe; }
          if (d.bl_l == 0) {
            if (d.nbt_l > 1) {
              read_block_switch<FAST>(d, 0, d.nbt_l, d.rbt_l0, d.rbt_l1, d.bl_l);
              prepare_literal_decoding(d);
              if (d.trivial_ctx) break;
            } else if (SAFE) { d.ins_rem = i; return kNeedsMoreInput; }
          }
          const uint32_t context = d.luts.ctx_lut[d.ctx_mode_off + p1] | d.luts.ctx_lut[d.ctx_mode_off + 256 + p2];
          const uint16_t* tree = lit_ptrs[map[d.ctx_slice + context]];
          uint32_t len;
          const uint32_t lit = decode_symbol(tree, br.peek(), len);
          br.skip<FAST>(len);
          bool keep = true;
          if (SAFE) {
            if (br.overrun()) { d.ins_rem = i; return kNeedsMoreInput; }
            if (d.pos >= d.cap) keep = false;
          }
          p2 = p1; p1 = lit;
          if (keep) { if (lane == 0) d.out[d.pos] = (uint8_t)lit; }
          if (d.bl_l == 0) return kErrWindowBits;  // src/decode.rs:2522-2526
          d.bl_l--;
          if (keep) d.pos++; else d.discarded++;
          i--;
          if (SAFE && (uint64_t)d.pos + d.discarded == d.next_flush) { d.ins_rem = i; int r = flush_event(d); if (r != kSuccess) return r; }
        } while (i != 0);
      }
      d.ins_rem = i;
      if (i != 0) continue;  // literal block switch flipped trivial <-> contextual: re-dispatch
      if (SAFE && d.discarded != 0 && d.mlen >= 0) return kNeedsMoreOutput;
      if (d.mlen <= 0) return kMetablockDone;
      d.state = kCmdPostLiterals;
    }
    // ---- kCmdPostLiterals: distance, then copy or dictionary word (src/decode.rs:2559-2689) ----
    if (FAST && !br.fast_ok(8)) return kNeedSafe;
    int32_t dist;
    uint32_t push = 0;
    if (d.implicit_dist) {
      dist = d.d0;
    } else {
      if (d.bl_d == 0 && d.nbt_d > 1) {  // DecodeDistanceBlockSwitch, src/decode.rs:1643-1658
        read_block_switch<FAST>(d, 2, d.nbt_d, d.rbt_d0, d.rbt_d1, d.bl_d);
        d.dist_slice = d.rbt_d1 << 2;
        refresh_cur_dist(d);
      } else if (d.bl_d == 0 && SAFE) {
        return kNeedsMoreInput;
      }
      dist = read_distance<FAST>(d, push);
      if (SAFE && br.overrun()) return kNeedsMoreInput;
    }
    // src/decode.rs:2583-2589 (and :1799-1801, :3314-3316: the full-size ring has wrapped by the time pos passes max_backward)
    const uint32_t max_distance = (int64_t)d.pos < d.cdict_limit ? d.pos + d.cdict_size : d.max_backward;
    const uint32_t copy_len = d.copy_len;
    if (dist > (int32_t)max_distance) {  // static dictionary, src/decode.rs:2593-2640
      if (dist > 0x7FFFFFFC) return kErrDistance;
      if (copy_len < 4 || copy_len > 24) return kErrDictionary;
      const uint32_t word_id = (uint32_t)dist - max_distance - 1;
      const uint32_t shift = tbl::kBrotliDictSizeBitsByLength[copy_len];
      const uint32_t word_idx = word_id & ((1u << shift) - 1u);
      const uint32_t transform_idx = word_id >> shift;
      if (transform_idx >= 121) return kErrTransform;
      const uint32_t offset = tbl::kBrotliDictOffsetsByLength[copy_len] + word_idx * copy_len;
      uint32_t n;
      if (transform_idx == 0) {
        n = copy_len;
        hw::syncwarp();
        for (uint32_t i = lane; i < n; i += hw::kWarp) d.ws->word[i] = hw::ldg8(d.luts.dictionary + offset + i);
        hw::syncwarp();
      } else {
        hw::syncwarp();
        n = build_dictionary_word(d, offset, copy_len, transform_idx);
      }
      uint32_t room = n;
      if (SAFE && d.cap - d.pos < n) room = d.cap - d.pos;
      for (uint32_t i = lane; i < room; i += hw::kWarp) d.out[d.pos + i] = d.ws->word[i];
      if (room < n) {
        if (d.mlen - (int32_t)n < 0)  // overshoots MLEN: the reference fails before it ever flushes (see literal loop)
          return (uint64_t)d.pos + d.discarded + n >= d.next_flush ? kErrBlockLength1 : kErrBlockLength2;
        d.pos += room; return kNeedsMoreOutput;
      }
      d.pos += n;
      d.mlen -= (int32_t)n;
      if (SAFE && d.pos >= d.next_flush) { int r = flush_event(d); if (r != kSuccess) return r; }
    } else {
      if (dist <= 0) return kErrUnreachable;  // cannot happen for symbols < max_symbol; keeps the copy in bounds
      if (push) { d.d3 = d.d2; d.d2 = d.d1; d.d1 = d.d0; d.d0 = dist; }
      d.mlen -= (int32_t)copy_len;
      if (FAST && (uint32_t)dist <= d.pos) {
        warp_copy(d.out, d.pos, (uint32_t)dist, copy_len);
        d.pos += copy_len;
      } else {
        uint32_t left = copy_len;
        while (left > 0) {
          if (d.pos >= d.cap) {
            if (d.mlen < 0) return (uint64_t)d.pos + d.discarded + left >= d.next_flush ? kErrBlockLength1 : kErrBlockLength2;
            return kNeedsMoreOutput;
          }
          uint64_t seg = left;
          if (seg > d.cap - d.pos) seg = d.cap - d.pos;
          if (seg > d.next_flush - d.pos) seg = d.next_flush - d.pos;
          if ((uint32_t)dist > d.pos) {  // the source starts in the custom dictionary, which precedes the output
            const uint32_t back = (uint32_t)dist - d.pos;  // <= cdict_size by the max_distance test above
            if (seg > back) seg = back;
            hw::syncwarp();
            for (uint32_t i = lane; i < (uint32_t)seg; i += hw::kWarp) d.out[d.pos + i] = hw::ldg8(d.cdict + (d.cdict_size - back) + i);
          } else
          warp_copy(d.out, d.pos, (uint32_t)dist, (uint32_t)seg);
          d.pos += (uint32_t)seg; left -= (uint32_t)seg;
          if (d.pos == d.next_flush) { int r = flush_event(d); if (r != kSuccess) return r; }
        }
      }
    }
    d.state = kCmdBegin;
    if (d.mlen <= 0) return kMetablockDone;
    if (SAFE) return kRetryFast;
  }
}

// ======================= command loop, fast path =======================
// The loop the kernel spends its time in.  All hot state is copied into locals (registers) and
// written back on exit.  It handles whole commands whose input lies in whole words well inside the
// stream and whose output stays clear of the capacity and of the emulated flush points; anything
// else -- block switches, the tail of the stream, the last bytes of the output -- makes it write its
// state back and return kNeedSafe, and process_commands<true> finishes that command.
BD_COLD uint32_t emit_dictionary_word(Decoder& d, uint32_t pos, uint32_t word_id, uint32_t copy_len, int& err) {
  const uint32_t lane = hw::lane();
  const uint32_t shift = tbl::kBrotliDictSizeBitsByLength[copy_len];
  const uint32_t word_idx = word_id & ((1u << shift) - 1u);
  const uint32_t transform_idx = word_id >> shift;
  if (transform_idx >= 121) { err = kErrTransform; return 0; }
  const uint32_t offset = tbl::kBrotliDictOffsetsByLength[copy_len] + word_idx * copy_len;
  uint32_t n;
  hw::syncwarp();
  if (transform_idx == 0) {
    n = copy_len;
    for (uint32_t i = lane; i < n; i += hw::kWarp) d.out[pos + i] = hw::ldg8(d.luts.dictionary + offset + i);
  } else {
    n = build_dictionary_word(d, offset, copy_len, transform_idx);
    for (uint32_t i = lane; i < n; i += hw::kWarp) d.out[pos + i] = d.ws->word[i];
  }
  err = kSuccess;
  return n;
}

BD_DEV int process_commands_fast(Decoder& d) {
  const uint32_t lane = hw::lane();
  BitReader br = d.br;
  if (!br.fast_ok(3)) return kNeedSafe;
  uint8_t* const out = d.out;
  uint32_t pos = d.pos;
  int32_t mlen = d.mlen;
  int32_t d0 = d.d0, d1 = d.d1, d2 = d.d2, d3 = d.d3;
  uint32_t bl_c = d.bl_c, bl_l = d.bl_l, bl_d = d.bl_d;
  const uint16_t* const cmd_tree = d.cmd_tree;
  const uint16_t* const lit_tree = d.lit_tree;
  const uint32_t trivial_ctx = d.trivial_ctx;
  const uint2* const cmd_lut = d.luts.cmd_lut;
  const uint16_t* const* const cur_dist = d.sh->cur_dist;
  const uint32_t max_backward = d.max_backward;
  const uint32_t npostfix = d.npostfix, ndirect = d.ndirect;
  const uint64_t lim64 = d.next_flush < d.cap ? d.next_flush : d.cap;
  const uint32_t limit = lim64 > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)lim64;
  // command in flight (only meaningful when we bail out in the middle of one)
  uint32_t state = kCmdBegin, ins_rem = 0, copy_len = 0, implicit_dist = 0, dist_ctx = 0;
  int ret;
  for (;;) {
    // ---- insert&copy symbol and its extra bits (ReadCommandInternal, src/decode.rs:2134-2189) ----
    if (bl_c == 0) { ret = kNeedSafe; break; }
    uint32_t len;
    const uint32_t sym = decode_symbol(cmd_tree, br.peek(), len);
    br.skip<true>(len);
    const uint2 lut = cmd_lut[sym];
    const uint32_t ie = (lut.x >> 16) & 0xFF, ce = lut.y >> 16;
    uint32_t insert_len = lut.x & 0xFFFF;
    if (ie) insert_len += br.read<true>(ie);
    copy_len = lut.y & 0xFFFF;
    if (ce) copy_len += br.read<true>(ce);
    implicit_dist = (lut.x >> 26) & 1;
    dist_ctx = (lut.x >> 24) & 3;
    bl_c--;
    mlen -= (int32_t)insert_len;
    ins_rem = insert_len;
    state = insert_len ? kCmdInner : kCmdPostLiterals;
    // one guard per command: input words for the literals + distance + the next command's symbol,
    // output room for the literals and the copy, and no literal block switch inside the run
    if (br.k + 16u + (insert_len >> 1) > br.end_full || insert_len > bl_l ||
        (uint64_t)pos + insert_len + (copy_len > 40 ? copy_len : 40) >= limit) { ret = kNeedSafe; break; }
    // ---- literals (src/decode.rs:2391-2558) ----
    if (insert_len) {
      bl_l -= insert_len;
      if (trivial_ctx) {
        uint32_t i = 0;
        do {
          uint32_t l2;
          const uint32_t lit = decode_symbol(lit_tree, br.peek(), l2);
          br.skip<true>(l2);
          if (lane == 0) out[pos + i] = (uint8_t)lit;
        } while (++i != insert_len);
      } else {
        hw::syncwarp();
        uint32_t p1 = pos >= 1 ? out[pos - 1] : 0, p2 = pos >= 2 ? out[pos - 2] : 0;
        const uint8_t* const map = d.map_lit + d.ctx_slice;
        const uint8_t* const ctx_lut = d.luts.ctx_lut + d.ctx_mode_off;
        const uint16_t* const* const lit_ptrs = d.lit_ptrs;
        uint32_t i = 0;
        do {
          const uint32_t context = ctx_lut[p1] | ctx_lut[256 + p2];
          const uint16_t* tree = lit_ptrs[map[context]];
          uint32_t l2;
          const uint32_t lit = decode_symbol(tree, br.peek(), l2);
          br.skip<true>(l2);
          p2 = p1; p1 = lit;
          if (lane == 0) out[pos + i] = (uint8_t)lit;
        } while (++i != insert_len);
      }
      pos += insert_len;
      ins_rem = 0;
      state = kCmdPostLiterals;
      if (mlen <= 0) { ret = kMetablockDone; break; }
    }
    // ---- distance (src/decode.rs:2066-2131) ----
    int32_t dist = d0;
    uint32_t push = 0;
    if (!implicit_dist) {
      if (bl_d == 0) { ret = kNeedSafe; break; }
      bl_d--;
      uint32_t l3;
      const uint32_t code = decode_symbol(cur_dist[dist_ctx], br.peek(), l3);
      br.skip<true>(l3);
      push = 1;
      if (code >= ndirect) {
        const uint32_t distval = code - ndirect;
        const uint32_t postfix = distval & ((1u << npostfix) - 1u);
        const uint32_t hcode = distval >> npostfix;
        const uint32_t nbits = (hcode >> 1) + 1;
        const uint32_t extra = br.read<true>(nbits);
        const uint32_t offset = ((2u + (hcode & 1u)) << nbits) - 4u;
        dist = (int32_t)(((offset + extra) << npostfix) + postfix + ndirect - 15u);
      } else if (code >= 16) {
        dist = (int32_t)(code - 15u);
      } else if (code == 0) {
        push = 0;
      } else if (code < 4) {
        dist = code == 1 ? d1 : (code == 2 ? d2 : d3);
      } else {  // last / second-to-last distance -3..+3, src/decode.rs:2017-2049
        const uint32_t c = code - 4;
        const int32_t base = c < 6 ? d0 : d1;
        const uint32_t m = c < 6 ? c : c - 6;
        const int32_t delta = (int32_t)(m >> 1) + 1;
        dist = (m & 1) ? base + delta : base - delta;
        if (!(m & 1) && dist <= 0) dist = 0x7fffffff;
      }
    }
    // ---- copy or static dictionary word (src/decode.rs:2583-2689) ----
    const uint32_t max_distance = pos < max_backward ? pos : max_backward;
    if (dist > (int32_t)max_distance) {
      if (dist > 0x7FFFFFFC) { ret = kErrDistance; break; }
      if (copy_len < 4 || copy_len > 24) { ret = kErrDictionary; break; }
      int err;
      const uint32_t n = emit_dictionary_word(d, pos, (uint32_t)dist - max_distance - 1, copy_len, err);
      if (err != kSuccess) { ret = err; break; }
      pos += n;
      mlen -= (int32_t)n;
    } else {
      if (dist <= 0) { ret = kErrUnreachable; break; }
      if (push) { d3 = d2; d2 = d1; d1 = d0; d0 = dist; }
      mlen -= (int32_t)copy_len;
      warp_copy(out, pos, (uint32_t)dist, copy_len);
      pos += copy_len;
    }
    state = kCmdBegin;
    if (mlen <= 0) { ret = kMetablockDone; break; }
  }
  d.br = br;
  d.pos = pos; d.mlen = mlen;
  d.d0 = d0; d.d1 = d1; d.d2 = d2; d.d3 = d3;
  d.bl_c = bl_c; d.bl_l = bl_l; d.bl_d = bl_d;
  d.state = state; d.ins_rem = ins_rem; d.copy_len = copy_len; d.implicit_dist = implicit_dist; d.dist_ctx = dist_ctx;
  return ret;
}

// ======================= command loop, shared-memory path =======================
// Same contract as process_commands_fast, for metablocks whose tables all live in the warp's shared
// memory (Decoder::all_shared, the common case): table and LUT reads are ld.shared with 32-bit
// addresses, the rare second-level lookup is out of line, distance symbols go through dist_lut.
BD_COLD uint32_t second_level_shared(hw::sref_t tab, uint32_t e, uint32_t bits) {
  const uint32_t wbits = (e & 15u) - kRootBits;
  const uint32_t e2 = hw::lds16(tab + (((e >> 4) + ((bits >> kRootBits) & ((1u << wbits) - 1u))) << 1));
  return e2 + kRootBits;  // symbol << 4 | total length (<= 15)
}

BD_DEV uint32_t decode_symbol_shared(hw::sref_t tab, uint32_t bits, uint32_t& len) {
  uint32_t e = hw::lds16(tab + ((bits & 0xFFu) << 1));
  if ((e & 15u) > kRootBits) e = second_level_shared(tab, e, bits);
  len = e & 15u;
  return e >> 4;
}

#if defined(BROTLI_B200_HOSTSIM)
#define BD_LIKELY(x) (x)
#define BD_UNLIKELY(x) (x)
#else
#define BD_LIKELY(x) __builtin_expect(!!(x), 1)
#define BD_UNLIKELY(x) __builtin_expect(!!(x), 0)
#endif

BD_DEV int process_commands_shared(Decoder& d) {
  const uint32_t lane = hw::lane();
  BitReader br = d.br;
  if (!br.fast_ok(3)) return kNeedSafe;
  uint8_t* const out = d.out;
  uint32_t pos = d.pos;
  int32_t mlen = d.mlen;
  int32_t d0 = d.d0, d1 = d.d1, d2 = d.d2, d3 = d.d3;
  uint32_t bl_c = d.bl_c, bl_l = d.bl_l, bl_d = d.bl_d;
  const hw::sref_t cmd_tree = hw::to_sref(d.cmd_tree);
  const hw::sref_t lit_tree = hw::to_sref(d.lit_tree);
  const hw::sref_t cmd_lut = hw::to_sref(d.luts.cmd_lut);
  const hw::sref_t cur_dist = hw::to_sref(d.sh->cur_dist_s);
  const hw::sref_t dist_lut = hw::to_sref(d.sh->dist_lut);
  const uint32_t trivial_ctx = d.trivial_ctx;
  const uint32_t use_dist_lut = d.use_dist_lut;
  const uint32_t max_backward = d.max_backward;
  const uint32_t npostfix = d.npostfix, ndirect = d.ndirect;
  uint64_t lim64 = d.next_flush < d.cap ? d.next_flush : d.cap;
  if (lim64 > 0xF0000000ull) lim64 = 0xF0000000ull;  // keeps pos + insert_len + copy_len inside 32 bits
  const uint32_t limit = (uint32_t)lim64;
  const uint32_t k_end = br.end_full;
  uint32_t insert_len = 0, copy_len = 0, cmd_bits = 0;  // cmd_bits: implicit-distance flag (bit 26) and distance context (bits 24-25)
  // Short backreference copies are split in two: the byte is loaded when the command is decoded and
  // stored when the NEXT copy (or anything that reads the output) comes up, so the load's round
  // trip to L2/HBM overlaps the decode of the following command instead of stalling the warp.
  uint8_t* pend_ptr = out;
  uint32_t pend_val = 0;
  bool pend = false;
#define BD_FLUSH_PENDING() do { if (pend) { *pend_ptr = (uint8_t)pend_val; pend = false; } } while (0)
  int ret;
  for (;;) {
    // ---- insert&copy symbol and its extra bits (ReadCommandInternal, src/decode.rs:2134-2189) ----
    if (BD_UNLIKELY(bl_c == 0)) goto bail_begin;
    {
      uint32_t len;
      const uint32_t sym = decode_symbol_shared(cmd_tree, br.peek(), len);
      br.skip<true>(len);
      const uint2 lut = hw::lds64(cmd_lut + (sym << 3));
      cmd_bits = lut.x;
      insert_len = lut.x & 0xFFFF;
      copy_len = lut.y & 0xFFFF;
      if (BD_UNLIKELY((lut.x & 0xFF0000u) | (lut.y >> 16))) {
        if (lut.x & 0xFF0000u) insert_len += br.read<true>((lut.x >> 16) & 0xFF);
        if (lut.y >> 16) copy_len += br.read<true>(lut.y >> 16);
      }
    }
    bl_c--;
    mlen -= (int32_t)insert_len;
    // one guard per command: input words for the literals + distance + the next command's symbol,
    // output room for the literals and the copy, and no literal block switch inside the run
    if (BD_UNLIKELY(br.k + 16u + (insert_len >> 1) > k_end || insert_len > bl_l ||
                    pos + insert_len + (copy_len > 40 ? copy_len : 40) >= limit))
      goto bail_command;
    // ---- literals (src/decode.rs:2391-2558) ----
    if (insert_len) {
      bl_l -= insert_len;
      if (BD_LIKELY(trivial_ctx)) {
        uint32_t i = 0;
#pragma unroll 1
        do {
          uint32_t l2;
          const uint32_t lit = decode_symbol_shared(lit_tree, br.peek(), l2);
          br.skip<true>(l2);
          if (lane == 0) out[pos + i] = (uint8_t)lit;
        } while (++i != insert_len);
      } else {
        BD_FLUSH_PENDING();
        hw::syncwarp();
        uint32_t p1 = pos >= 1 ? out[pos - 1] : 0, p2 = pos >= 2 ? out[pos - 2] : 0;
        const hw::sref_t map = hw::to_sref(d.map_lit + d.ctx_slice);
        const hw::sref_t ctx_lut = hw::to_sref(d.luts.ctx_lut + d.ctx_mode_off);
        const hw::sref_t lit_s = hw::to_sref(d.sh->lit_s);
        uint32_t i = 0;
#pragma unroll 1
        do {
          const uint32_t context = hw::lds8(ctx_lut + p1) | hw::lds8(ctx_lut + 256 + p2);
          const hw::sref_t tree = hw::lds_ref(lit_s + hw::lds8(map + context) * (uint32_t)sizeof(hw::sref_t));
          uint32_t l2;
          const uint32_t lit = decode_symbol_shared(tree, br.peek(), l2);
          br.skip<true>(l2);
          p2 = p1; p1 = lit;
          if (lane == 0) out[pos + i] = (uint8_t)lit;
        } while (++i != insert_len);
      }
      pos += insert_len;
      insert_len = 0;
      if (BD_UNLIKELY(mlen <= 0)) goto done_after_literals;
    }
    // ---- distance (src/decode.rs:2066-2131) ----
    {
      int32_t dist = d0;
      uint32_t push = 0;
      if (!(cmd_bits & (1u << 26))) {
        if (BD_UNLIKELY(bl_d == 0)) goto bail_command;
        bl_d--;
        const hw::sref_t tree = hw::lds_ref(cur_dist + ((cmd_bits >> 24) & 3u) * (uint32_t)sizeof(hw::sref_t));
        const uint32_t bits = br.peek();
        uint32_t l3;
        const uint32_t code = decode_symbol_shared(tree, bits, l3);
        push = 1;
        if (code >= 16) {
          uint32_t base, nbits;
          if (BD_LIKELY(use_dist_lut)) {
            const uint32_t e = hw::lds32(dist_lut + ((code - 16u) << 2));
            nbits = e & 31u; base = e >> 5;
          } else if (code >= ndirect) {
            const uint32_t distval = code - ndirect;
            const uint32_t hcode = distval >> npostfix;
            nbits = (hcode >> 1) + 1;
            base = ((((2u + (hcode & 1u)) << nbits) - 4u) << npostfix) + (distval & ((1u << npostfix) - 1u)) + ndirect - 15u;
          } else {
            nbits = 0; base = code - 15u;
          }
          if (BD_LIKELY(l3 + nbits <= 32)) {  // symbol and extra bits from the one 32-bit peek
            const uint32_t extra = (bits >> l3) & ((1u << nbits) - 1u);
            br.skip<true>(l3 + nbits);
            dist = (int32_t)(base + (extra << npostfix));
          } else {
            br.skip<true>(l3);
            dist = (int32_t)(base + (br.read<true>(nbits) << npostfix));
          }
        } else {
          br.skip<true>(l3);
          if (code == 0) {
            push = 0;
          } else if (code < 4) {
            dist = code == 1 ? d1 : (code == 2 ? d2 : d3);
          } else {  // last / second-to-last distance -3..+3, src/decode.rs:2017-2049
            const uint32_t c = code - 4;
            const int32_t b = c < 6 ? d0 : d1;
            const uint32_t m = c < 6 ? c : c - 6;
            const int32_t delta = (int32_t)(m >> 1) + 1;
            dist = (m & 1) ? b + delta : b - delta;
            if (!(m & 1) && dist <= 0) dist = 0x7fffffff;
          }
        }
      }
      // ---- copy or static dictionary word (src/decode.rs:2583-2689) ----
      const uint32_t max_distance = pos < max_backward ? pos : max_backward;
      if (BD_UNLIKELY((uint32_t)dist > max_distance)) {  // also catches dist <= 0 (impossible for symbols < max_symbol)
        if (dist <= 0) { ret = kErrUnreachable; goto fail; }
        if (dist > 0x7FFFFFFC) { ret = kErrDistance; goto fail; }
        if (copy_len < 4 || copy_len > 24) { ret = kErrDictionary; goto fail; }
        int err;
        const uint32_t n = emit_dictionary_word(d, pos, (uint32_t)dist - max_distance - 1, copy_len, err);
        if (err != kSuccess) { ret = err; goto fail; }
        pos += n;
        mlen -= (int32_t)n;
      } else {
        if (push) { d3 = d2; d2 = d1; d1 = d0; d0 = dist; }
        mlen -= (int32_t)copy_len;
        BD_FLUSH_PENDING();
        hw::syncwarp();  // literal and copy stores issued so far are visible to the loads below
        if (BD_LIKELY(copy_len <= hw::kWarp && (uint32_t)dist >= copy_len)) {  // one byte per lane, no overlap
          if (lane < copy_len) {
            pend_ptr = out + (pos + lane);
            pend_val = *(pend_ptr - (uint32_t)dist);
            pend = true;
          }
        } else {
          warp_copy(out, pos, (uint32_t)dist, copy_len);
        }
        pos += copy_len;
      }
    }
    if (BD_UNLIKELY(mlen <= 0)) goto done_after_copy;
  }
bail_begin:  // before reading a command
  d.state = kCmdBegin; d.ins_rem = 0;
  ret = kNeedSafe;
  goto writeback;
bail_command:  // command read; literals (if any) and the distance are still to come
  d.state = insert_len ? kCmdInner : kCmdPostLiterals; d.ins_rem = insert_len;
  ret = kNeedSafe;
  goto writeback;
done_after_literals:
  d.state = kCmdPostLiterals; d.ins_rem = 0;
  ret = kMetablockDone;
  goto writeback;
done_after_copy:
  d.state = kCmdBegin; d.ins_rem = 0;
  ret = kMetablockDone;
  goto writeback;
fail:
  d.state = kCmdPostLiterals; d.ins_rem = 0;
writeback:
  BD_FLUSH_PENDING();
  hw::syncwarp();
#undef BD_FLUSH_PENDING
  d.br = br;
  d.pos = pos; d.mlen = mlen;
  d.d0 = d0; d.d1 = d1; d.d2 = d2; d.d3 = d3;
  d.bl_c = bl_c; d.bl_l = bl_l; d.bl_d = bl_d;
  d.copy_len = copy_len; d.implicit_dist = (cmd_bits >> 26) & 1; d.dist_ctx = (cmd_bits >> 24) & 3;
  return ret;
}

// ======================= metablock header (src/decode.rs:243-372, :3046-3288) =======================
BD_DEV int read_block_split_header(Decoder& d, int cat, uint32_t& num_types, uint32_t& block_length) {
  num_types = read_varlen_uint8(d.br) + 1;
  block_length = 1u << 24;
  if (num_types < 2) return kSuccess;
  uint32_t tsize;
  int r = read_huffman_code(d, num_types + 2, num_types + 2, d.block_type_tree(cat), kMaxBlockTypeTable, tsize);
  if (r != kSuccess) return r;
  r = read_huffman_code(d, 26, 26, d.block_len_tree(cat), kMaxBlockLenTable, tsize);
  if (r != kSuccess) return r;
  block_length = read_block_length<false>(d.br, d.block_len_tree(cat));
  return kSuccess;
}

// BrotliMaxDistanceSymbol, src/decode.rs:2766-2777
BD_DEV uint32_t max_distance_symbol(uint32_t ndirect, uint32_t npostfix) {
  const uint32_t bound = npostfix == 0 ? 0u : npostfix == 1 ? 4u : npostfix == 2 ? 12u : 28u;
  const uint32_t diff = npostfix == 0 ? 73u : npostfix == 1 ? 126u : npostfix == 2 ? 228u : 424u;
  const uint32_t postfix = 1u << npostfix;
  if (ndirect < bound) return ndirect + diff + postfix;
  if (ndirect > bound + postfix) return ndirect + diff;
  return bound + diff + postfix;
}

// Distance symbols >= 16 as a lookup: dist = (lut >> 5) + (extra << NPOSTFIX) with (lut & 31) extra
// bits -- the closed form of src/decode.rs:2100-2128 tabulated per metablock (NPOSTFIX / NDIRECT are
// metablock parameters).  Only used when the whole alphabet fits the table and every base fits 27 bits.
BD_DEV void build_dist_lut(Decoder& d) {
  bool ok = d.dist_alphabet <= 16 + WarpShared::kDistLut;
  if (ok) {
    for (uint32_t i = hw::lane(); i < WarpShared::kDistLut; i += hw::kWarp) {
      const uint32_t code = 16 + i;
      uint32_t base = code - 15u, nbits = 0;
      if (code >= d.ndirect) {
        const uint32_t distval = code - d.ndirect;
        const uint32_t postfix = distval & ((1u << d.npostfix) - 1u);
        const uint32_t hcode = distval >> d.npostfix;
        nbits = (hcode >> 1) + 1;
        const uint64_t b = ((uint64_t)(((2u + (hcode & 1u)) << nbits) - 4u) << d.npostfix) + postfix + d.ndirect - 15u;
        if (code < d.dist_alphabet && (b >> 27) != 0) ok = false;
        base = (uint32_t)b;
      }
      d.sh->dist_lut[i] = (base << 5) | (nbits & 31u);
    }
  }
  d.use_dist_lut = hw::all(ok) ? 1u : 0u;
}

// HuffmanTreeGroupDecode, src/decode.rs:1130-1219.  Each table is built in the arena; tree_off[i] is its offset in
// the arena's table space.
BD_DEV int read_tree_group(Decoder& d, uint32_t ntrees, uint32_t alphabet, uint32_t max_symbol, uint32_t* tree_off, uint32_t& next_off) {
  for (uint32_t i = 0; i < ntrees; i++) {
    uint32_t tsize = 0;
    const uint32_t room = (uint32_t)(ArenaLayout::kTableEntries - next_off);
    int r = read_huffman_code(d, alphabet, max_symbol, d.tables + next_off, room, tsize);
    if (r != kSuccess) return r;
    tree_off[i] = next_off;
    next_off += tsize;
  }
  hw::syncwarp();
  return kSuccess;
}

// Tables of one group -> ptrs[]: while a table fits (keeping `reserve` entries free for the groups that follow) it is
// copied into the warp's shared-memory table storage and ptrs[i] points there, else ptrs[i] points into the arena.
BD_DEV void promote_tree_group(Decoder& d, uint32_t ntrees, const uint32_t* tree_off, uint32_t group_end, const uint16_t** ptrs,
                               uint32_t reserve) {
  const uint32_t lane = hw::lane();
  for (uint32_t i = 0; i < ntrees; i++) {
    const uint32_t off = tree_off[i];
    const uint32_t tsize = (i + 1 < ntrees ? tree_off[i + 1] : group_end) - off;
    const uint16_t* gtab = d.tables + off;
    const uint16_t* where = gtab;
    if (d.stab_used + tsize + reserve <= d.stab_cap) {
      uint16_t* stab = d.stab + d.stab_used;
      for (uint32_t j = lane; j < tsize; j += hw::kWarp) stab[j] = gtab[j];
      d.stab_used += tsize;
      where = stab;
    } else {
      d.n_unpromoted++;
    }
    ptrs[i] = where;
  }
  hw::syncwarp();
}

// Where the current metablock's lookup structures live (shared memory when small enough, else the arena), from what
// the header parse left in the arena: context maps, tree offsets, tables.  Also what a streaming session re-creates
// when it continues inside a metablock (the arena belongs to the session; shared memory does not survive a launch).
BD_COLD void setup_metablock_views(Decoder& d) {
  WarpShared* sh = d.sh;
  const uint16_t** lit_ptrs = d.n_lit_trees <= WarpShared::kLitPtrs ? sh->lit_ptrs : (const uint16_t**)(d.arena + ArenaLayout::kPtrLit);
  const uint16_t** cmd_ptrs = d.nbt_c <= WarpShared::kCmdPtrs ? sh->cmd_ptrs : (const uint16_t**)(d.arena + ArenaLayout::kPtrCmd);
  const uint16_t** dist_ptrs = d.n_dist_trees <= WarpShared::kDistPtrs ? sh->dist_ptrs : (const uint16_t**)(d.arena + ArenaLayout::kPtrDist);
  d.lit_ptrs = lit_ptrs; d.cmd_ptrs = cmd_ptrs; d.dist_ptrs = dist_ptrs;
  d.map_lit = d.ctx_map_lit(); d.map_dist = d.ctx_map_dist();
  const uint32_t lane = hw::lane();
  if ((d.nbt_l << 6) <= WarpShared::kLitMapBytes) {
    for (uint32_t i = lane; i < (d.nbt_l << 6); i += hw::kWarp) sh->lit_map[i] = d.ctx_map_lit()[i];
    d.map_lit = sh->lit_map;
  }
  if ((d.nbt_d << 2) <= WarpShared::kDistMapBytes) {
    for (uint32_t i = lane; i < (d.nbt_d << 2); i += hw::kWarp) sh->dist_map[i] = d.ctx_map_dist()[i];
    d.map_dist = sh->dist_map;
  }
  d.stab_used = 0;
  d.n_unpromoted = 0;
  const uint32_t reserve_dist = (d.n_dist_trees < 8 ? d.n_dist_trees : 8u) * 256u;
  const uint32_t reserve_cmd = (d.nbt_c < 2 ? d.nbt_c : 2u) * 1024u;
  promote_tree_group(d, d.n_lit_trees, d.tree_off_lit(), d.tree_off_cmd()[0], lit_ptrs, reserve_cmd + reserve_dist);
  promote_tree_group(d, d.nbt_c, d.tree_off_cmd(), d.tree_off_dist()[0], cmd_ptrs, reserve_dist);
  promote_tree_group(d, d.n_dist_trees, d.tree_off_dist(), d.tables_used, dist_ptrs, 0);
  hw::syncwarp();
  prepare_literal_decoding(d);
  d.dist_slice = d.rbt_d1 << 2;
  refresh_cur_dist(d);
  d.cmd_tree = d.cmd_ptrs[d.rbt_c1];
  // shared-space mirrors for process_commands_shared
  d.all_shared = d.n_unpromoted == 0 && d.map_lit == sh->lit_map && d.map_dist == sh->dist_map && lit_ptrs == sh->lit_ptrs &&
                 cmd_ptrs == sh->cmd_ptrs && dist_ptrs == sh->dist_ptrs;
  if (d.all_shared) {
    for (uint32_t i = lane; i < d.n_lit_trees; i += hw::kWarp) sh->lit_s[i] = hw::to_sref(lit_ptrs[i]);
  }
  build_dist_lut(d);
  hw::syncwarp();
}

// Everything between MLEN and the first command of a compressed metablock.
BD_COLD int read_compressed_metablock_header(Decoder& d) {
  BitReader& br = d.br;
  int r;
  if ((r = read_block_split_header(d, 0, d.nbt_l, d.bl_l)) != kSuccess) return r;
  if ((r = read_block_split_header(d, 1, d.nbt_c, d.bl_c)) != kSuccess) return r;
  if ((r = read_block_split_header(d, 2, d.nbt_d, d.bl_d)) != kSuccess) return r;
  d.rbt_l0 = 1; d.rbt_l1 = 0; d.rbt_c0 = 1; d.rbt_c1 = 0; d.rbt_d0 = 1; d.rbt_d1 = 0;  // src/state.rs:430-435
  const uint32_t bits = br.read<false>(6);  // src/decode.rs:3141-3153
  d.npostfix = bits & 3;
  d.ndirect = 16 + ((bits >> 2) << d.npostfix);
  uint8_t* modes = d.ctx_modes();
  for (uint32_t i = 0; i < d.nbt_l; i++) modes[i] = (uint8_t)br.read<false>(2);  // ReadContextModes, :1991-2015
  if (br.overrun()) return kNeedsMoreInput;
  if ((r = decode_context_map(d, d.nbt_l << 6, d.ctx_map_lit(), d.n_lit_trees)) != kSuccess) return r;
  if (br.overrun()) return kNeedsMoreInput;
  const uint32_t num_direct_codes = d.ndirect - 16;  // src/decode.rs:3189-3200
  d.dist_alphabet = 16 + num_direct_codes + ((d.large_window ? 62u : 24u) << (d.npostfix + 1));
  d.dist_max_symbol = d.large_window ? max_distance_symbol(num_direct_codes, d.npostfix) : d.dist_alphabet;
  if ((r = decode_context_map(d, d.nbt_d << 2, d.ctx_map_dist(), d.n_dist_trees)) != kSuccess) return r;
  if (br.overrun()) return kNeedsMoreInput;
  uint32_t next_off = 0;
  if ((r = read_tree_group(d, d.n_lit_trees, 256, 256, d.tree_off_lit(), next_off)) != kSuccess) return r;
  if ((r = read_tree_group(d, d.nbt_c, 704, 704, d.tree_off_cmd(), next_off)) != kSuccess) return r;
  if ((r = read_tree_group(d, d.n_dist_trees, d.dist_alphabet, d.dist_max_symbol, d.tree_off_dist(), next_off)) != kSuccess) return r;
  if (br.overrun()) return kNeedsMoreInput;
  d.tables_used = next_off;
  hw::syncwarp();
  setup_metablock_views(d);
  d.state = kCmdBegin;
  return kSuccess;
}

// "canny" ring-buffer sizing, src/decode.rs:1808-1871 -- only its size matters here.
BD_DEV void allocate_ring_emulation(Decoder& d, uint32_t is_last, uint32_t is_uncompressed) {
  if (is_uncompressed) {  // peek at the header of the next metablock: ISLAST + ISLASTEMPTY
    const uint64_t at = d.br.byte_pos() + (uint64_t)d.mlen;
    if (at < d.br.size() && (hw::ldg8(d.br.bytes + d.br.lead + at) & 3) == 3) is_last = 1;
  }
  uint64_t size = (uint64_t)1 << d.wbits;
  if (is_last)
    while (size >= ((uint64_t)d.cdict_size + (uint64_t)d.mlen + 16) * 2 && size > 32) size >>= 1;  // src/decode.rs:1843-1847
  d.rbsize = size;
  d.full_ring = size == ((uint64_t)1 << d.wbits);
  d.next_flush = size;
  d.rb_allocated = 1;
}

// CopyUncompressedBlockToOutput, src/decode.rs:1754-1806
BD_COLD int copy_uncompressed(Decoder& d) {
  const uint32_t lane = hw::lane();
  uint64_t in_pos = d.br.byte_pos();
  const uint64_t in_size = d.br.size();
  const uint8_t* in = d.br.bytes + d.br.lead;
  int result = kSuccess;
  for (;;) {
    uint64_t n = in_size > in_pos ? in_size - in_pos : 0;
    if (n > (uint64_t)d.mlen) n = (uint64_t)d.mlen;
    if (n > d.next_flush - d.pos) n = d.next_flush - d.pos;
    uint64_t n2 = n;
    if (n2 > d.cap - d.pos) n2 = d.cap - d.pos;
    for (uint64_t i = lane; i < n2; i += hw::kWarp) d.out[d.pos + i] = hw::ldg8(in + in_pos + i);
    d.pos += (uint32_t)n2; d.mlen -= (int32_t)n2; in_pos += n2;
    if (n2 < n) {
      // Output full.  The reference keeps copying into its ring buffer and reports a truncated
      // input first unless a flush point (which needs output space) comes before the input ends.
      const uint64_t in_left = in_size - in_pos;
      const uint64_t vend = (uint64_t)d.pos + (in_left < (uint64_t)d.mlen ? in_left : (uint64_t)d.mlen);
      result = (in_left < (uint64_t)d.mlen && !(d.full_ring && vend >= d.next_flush)) ? kNeedsMoreInput : kNeedsMoreOutput;
      break;
    }
    if (d.full_ring && d.pos == d.next_flush) {
      int r = flush_event(d);
      if (r != kSuccess) { result = r; break; }
      continue;
    }
    result = d.mlen == 0 ? kSuccess : kNeedsMoreInput;
    break;
  }
  d.br.seek_byte(in_pos);
  return result;
}

// ======================= stream driver (src/decode.rs:2779-3403) =======================
// (ResumeState -- the device-side record of a streaming session -- and SessionCopy: brotli_b200_session_types.h)
BD_DEV void save_checkpoint(Decoder& d, ResumeState* rs, uint32_t kind) {
  hw::syncwarp();
  if (hw::lane() == 0) {
    rs->kind = kind;
    rs->bitpos = d.br.bitpos() - 8 * (uint64_t)d.br.lead;
    rs->rbsize = d.rbsize; rs->next_flush = d.next_flush; rs->flushed = d.flushed;
    rs->pos = d.pos; rs->mlen = d.mlen; rs->d0 = d.d0; rs->d1 = d.d1; rs->d2 = d.d2; rs->d3 = d.d3;
    rs->wbits = d.wbits; rs->large_window = d.large_window; rs->rb_allocated = d.rb_allocated; rs->full_ring = d.full_ring;
    rs->is_last = d.is_last;
    if (kind == 2) {
      rs->nbt_l = d.nbt_l; rs->nbt_c = d.nbt_c; rs->nbt_d = d.nbt_d; rs->bl_l = d.bl_l; rs->bl_c = d.bl_c; rs->bl_d = d.bl_d;
      rs->rbt_l0 = d.rbt_l0; rs->rbt_l1 = d.rbt_l1; rs->rbt_c0 = d.rbt_c0; rs->rbt_c1 = d.rbt_c1; rs->rbt_d0 = d.rbt_d0; rs->rbt_d1 = d.rbt_d1;
      rs->n_lit_trees = d.n_lit_trees; rs->n_dist_trees = d.n_dist_trees; rs->npostfix = d.npostfix; rs->ndirect = d.ndirect;
      rs->dist_alphabet = d.dist_alphabet; rs->dist_max_symbol = d.dist_max_symbol; rs->tables_used = d.tables_used;
      rs->state = d.state; rs->ins_rem = d.ins_rem; rs->copy_len = d.copy_len; rs->implicit_dist = d.implicit_dist; rs->dist_ctx = d.dist_ctx;
    }
  }
  hw::syncwarp();
}

// Returns the BrotliDecoderErrorCode; *decoded_size follows the reference's flush rules:
// success / NeedsMoreInput -> everything decoded, NeedsMoreOutput -> capacity (or the flush point the budget of a
// session cannot pass), fatal -> only what the ring buffer had flushed (multiples of its size).
BD_DEV int decode_stream(Decoder& d, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap, uint32_t allow_large_window,
                         uint64_t* decoded_size, uint64_t* in_used, const uint8_t* custom_dict = nullptr, uint64_t custom_dict_size = 0,
                         ResumeState* resume = nullptr) {
  BitReader& br = d.br;
  br.init(in, in_size);
  d.out = out;
  d.cap = out_cap > 0xFFFFFF00ull ? 0xFFFFFF00u : (uint32_t)out_cap;
  d.pos = 0; d.mlen = 0; d.rb_allocated = 0; d.rbsize = 0; d.next_flush = ~(uint64_t)0; d.flushed = 0; d.full_ring = 0; d.discarded = 0;
  d.d0 = 4; d.d1 = 11; d.d2 = 15; d.d3 = 16;
  d.large_window = 0;
  d.budget = resume ? resume->budget : ~(uint64_t)0;
  d.at_flush = 0; d.is_last = 0; d.tables_used = 0;
  const uint32_t enter = resume ? resume->kind : 0u;  // where this launch picks the stream up
  int result;
  do {
    if (in_size >= ((uint64_t)1 << 32)) { result = kErrInvalidArguments; break; }  // src/decode.rs:2799-2801
    if (in_size == 0 && enter == 0) { result = kNeedsMoreInput; break; }
    if (enter != 0) {  // continue from the session's checkpoint
      d.wbits = resume->wbits; d.large_window = resume->large_window;
      d.pos = resume->pos; d.d0 = resume->d0; d.d1 = resume->d1; d.d2 = resume->d2; d.d3 = resume->d3;
      d.rb_allocated = resume->rb_allocated; d.full_ring = resume->full_ring;
      d.rbsize = resume->rbsize; d.next_flush = resume->next_flush; d.flushed = resume->flushed;
      d.mlen = resume->mlen; d.is_last = resume->is_last;
      br.seek_byte(resume->bitpos >> 3);
      br.skip<false>((uint32_t)(resume->bitpos & 7));
    } else
    // DecodeWindowBits, src/decode.rs:152-187,2940-2951
    if (br.read<false>(1) == 0) {
      d.wbits = 16;
    } else {
      uint32_t n = br.read<false>(3);
      if (n != 0) {
        d.wbits = 17 + n;
      } else {
        n = br.read<false>(3);
        if (n == 1) {
          if (!allow_large_window || br.read<false>(1) == 1) { result = kErrWindowBits; break; }
          d.large_window = 1;
          d.wbits = br.read<false>(6);
          if (br.overrun()) { result = kNeedsMoreInput; break; }
          if (d.wbits < 10 || d.wbits > 30) { result = kErrWindowBits; break; }
        } else {
          d.wbits = n != 0 ? 8 + n : 17;
        }
      }
    }
    d.max_backward = (1u << d.wbits) - 16;
    // custom dictionary: the window reaches its last max_backward bytes (src/decode.rs:1831-1838)
    d.cdict_given = custom_dict_size != 0;
    d.cdict_size = custom_dict_size > d.max_backward ? d.max_backward : (uint32_t)custom_dict_size;
    d.cdict = custom_dict + (custom_dict_size - d.cdict_size);
    d.cdict_limit = (int64_t)d.max_backward - (int64_t)custom_dict_size;
    uint32_t inside = enter >= 2 ? enter : 0u;  // 2 / 3: the first pass of the loop continues inside a metablock
    for (;;) {  // metablocks
      uint32_t is_uncompressed = 0, is_metadata = 0;
      result = kSuccess;
      if (inside == 0) {
        // DecodeMetaBlockLength, src/decode.rs:243-372
        const uint32_t is_last = br.read<false>(1);
        d.is_last = is_last;
        d.mlen = 0;
        if (is_last && br.read<false>(1)) {
          // ISLASTEMPTY: nothing else in this metablock
        } else {
          const uint32_t nib = br.read<false>(2);
          if (nib == 3) {
            is_metadata = 1;
            if (br.read<false>(1) != 0) { result = kErrReserved; }
            else {
              const uint32_t nbytes = br.read<false>(2);
              uint32_t v = 0;
              for (uint32_t i = 0; i < nbytes; i++) {
                const uint32_t b = br.read<false>(8);
                if (i + 1 == nbytes && nbytes > 1 && b == 0) { result = kErrExuberantMetaNibble; break; }
                v |= b << (i * 8);
              }
              if (nbytes != 0) d.mlen = (int32_t)v + 1;
            }
          } else {
            const uint32_t nn = nib + 4;
            uint32_t v = 0;
            for (uint32_t i = 0; i < nn; i++) {
              const uint32_t b = br.read<false>(4);
              if (i + 1 == nn && nn > 4 && b == 0) { result = kErrExuberantNibble; break; }
              v |= b << (i * 4);
            }
            if (result == kSuccess) {
              if (!is_last) is_uncompressed = br.read<false>(1);
              d.mlen = (int32_t)v + 1;
            }
          }
        }
        if (br.overrun()) result = kNeedsMoreInput;
        if (result != kSuccess) break;
        if ((is_metadata || is_uncompressed) && !br.jump_to_byte_boundary()) {  // src/decode.rs:2990-2994
          result = br.overrun() ? kNeedsMoreInput : kErrPadding2; break;
        }
        if (br.overrun()) { result = kNeedsMoreInput; break; }
      } else {
        is_uncompressed = inside == 3;
      }
      if (is_metadata) {  // src/decode.rs:3031-3045
        const uint64_t at = br.byte_pos(), left = br.size() - at;
        if (left < (uint64_t)d.mlen) { result = kNeedsMoreInput; break; }
        br.seek_byte(at + (uint64_t)d.mlen);
        d.mlen = 0;
      } else if (d.mlen != 0 || inside != 0) {
        if (!d.rb_allocated) allocate_ring_emulation(d, d.is_last, is_uncompressed);
        if (is_uncompressed) {
          result = copy_uncompressed(d);
          if (resume && (result == kNeedsMoreInput || result == kNeedsMoreOutput)) save_checkpoint(d, resume, 3);
          if (result != kSuccess) break;
        } else {
          if (inside == 2) {  // the tables of this metablock are in the session's arena
            d.nbt_l = resume->nbt_l; d.nbt_c = resume->nbt_c; d.nbt_d = resume->nbt_d;
            d.bl_l = resume->bl_l; d.bl_c = resume->bl_c; d.bl_d = resume->bl_d;
            d.rbt_l0 = resume->rbt_l0; d.rbt_l1 = resume->rbt_l1; d.rbt_c0 = resume->rbt_c0; d.rbt_c1 = resume->rbt_c1;
            d.rbt_d0 = resume->rbt_d0; d.rbt_d1 = resume->rbt_d1;
            d.n_lit_trees = resume->n_lit_trees; d.n_dist_trees = resume->n_dist_trees; d.npostfix = resume->npostfix; d.ndirect = resume->ndirect;
            d.dist_alphabet = resume->dist_alphabet; d.dist_max_symbol = resume->dist_max_symbol; d.tables_used = resume->tables_used;
            setup_metablock_views(d);
            d.state = resume->state; d.ins_rem = resume->ins_rem; d.copy_len = resume->copy_len;
            d.implicit_dist = resume->implicit_dist; d.dist_ctx = resume->dist_ctx;
          } else {
            result = read_compressed_metablock_header(d);
            if (result == kSuccess && br.overrun()) result = kNeedsMoreInput;
            if (result != kSuccess) break;
          }
          for (;;) {  // ProcessCommands / SafeProcessCommands, src/decode.rs:3289-3298
            // (streams with a custom dictionary take the checked loop for every command: the register-resident
            // loops assume that every distance stays inside the output region)
            // (a session that continues in the middle of a command lets the checked loop finish that command first)
            result = (d.cdict_given || d.state != kCmdBegin) ? kNeedSafe : (d.all_shared ? process_commands_shared(d) : process_commands_fast(d));
            if (result == kNeedSafe) {
              // the checked loop is where a decode can stop (end of input, ring flush point): a session continues here
              if (resume) save_checkpoint(d, resume, 2);
              result = process_commands<true>(d);
            }
            if (result != kRetryFast) break;
          }
          if (result != kMetablockDone) break;
          result = kSuccess;
        }
      }
      inside = 0;
      // BROTLI_STATE_METABLOCK_DONE, src/decode.rs:3345-3381
      if (d.mlen < 0) { result = kErrBlockLength2; break; }
      if (!d.is_last) {
        if (resume && !br.overrun()) save_checkpoint(d, resume, 1);  // a complete metablock: the next launch starts here
        continue;
      }
      if (!br.jump_to_byte_boundary()) { result = br.overrun() ? kNeedsMoreInput : kErrPadding2; break; }
      if (br.overrun()) { result = kNeedsMoreInput; break; }
      result = kSuccess;
      break;
    }
  } while (0);
  hw::syncwarp();
  // On NeedsMoreInput the reference flushes its ring buffer (src/decode.rs:2834-2846); that
  // flush fails with BLOCK_LENGTH_1 when the command in flight has overshot MLEN (:1709-1711).
  uint32_t forced_flush_error = 0;
  if (result == kNeedsMoreInput && d.rb_allocated && d.mlen < 0) { result = kErrBlockLength1; forced_flush_error = 1; }
  if (d.at_flush) *decoded_size = d.next_flush;
  else if (result == kSuccess || result == kNeedsMoreInput || result == kNeedsMoreOutput) *decoded_size = d.pos;
  else *decoded_size = d.flushed;
  // input bytes consumed (whole bytes; on success the reference un-reads its look-ahead, src/decode.rs:3374-3376)
  uint64_t used = ((br.bitpos() + 7) >> 3) - br.lead;
  *in_used = used < in_size ? used : in_size;
  if (resume && hw::lane() == 0) {
    resume->code = result; resume->decoded = *decoded_size; resume->used = *in_used; resume->flushed_now = d.flushed; resume->at_flush = d.at_flush;
    resume->hit_cap = (d.pos >= d.cap || d.discarded != 0) ? 1u : 0u;
    resume->forced_flush_error = forced_flush_error;
  }
  return result;
}

// One launch's worth of a streaming session: everything comes from (and goes back to) its ResumeState.
BD_DEV int decode_session(Decoder& d, ResumeState* rs) {
  uint8_t* const own_arena = d.arena;
  if (rs->arena) d.arena = rs->arena;  // (a one-launch record -- the exact re-decode of a one-shot -- uses the warp's arena)
  d.tables = (uint16_t*)(d.arena + ArenaLayout::kTables);
  uint64_t decoded = 0, used = 0;
  const int code = decode_stream(d, rs->in, rs->in_size, rs->out, rs->out_cap, rs->allow_large_window, &decoded, &used, rs->dict, rs->dict_size, rs);
  d.arena = own_arena;
  d.tables = (uint16_t*)(d.arena + ArenaLayout::kTables);
  return code;
}

}  // namespace brotli_b200
