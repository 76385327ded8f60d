/* oracle/brotli_oracle.c -- CPU restatement of dropbox/rust-brotli-decompressor (v5.0.3).
 *
 * TEST INFRASTRUCTURE ONLY (see brotli_oracle.h).  Plain C99, one translation unit.
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * The restatement keeps the reference's data structures that decide observable behaviour:
 * the 64-bit LSB-first bit window, the 2-level (root 8 bit) HuffmanCode tables, the ring
 * buffer with its "canny" sizing and flush points (they decide decoded_size on errors), and
 * the fast/safe split of the command loop.  What it drops is resumability: a one-shot call
 * owns all input and the whole output buffer, so NeedsMoreInput / NeedsMoreOutput are terminal
 * (exactly how brotli_decode() at src/lib.rs:446-468 uses BrotliDecompressStream).
 */
#include "brotli_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../tables/brotli_tables.h"

extern const uint8_t kBrotliDictionaryData[BROTLI_DICTIONARY_SIZE]; /* tables/brotli_dictionary.c */

/* ---- BrotliDecoderErrorCode, src/state.rs:22-65 ---- */
enum {
  E_NO_ERROR = 0, E_SUCCESS = 1, E_NEEDS_MORE_INPUT = 2, E_NEEDS_MORE_OUTPUT = 3,
  E_EXUBERANT_NIBBLE = -1, E_RESERVED = -2, E_EXUBERANT_META_NIBBLE = -3, E_SIMPLE_HUFFMAN_ALPHABET = -4,
  E_SIMPLE_HUFFMAN_SAME = -5, E_CL_SPACE = -6, E_HUFFMAN_SPACE = -7, E_CONTEXT_MAP_REPEAT = -8,
  E_BLOCK_LENGTH_1 = -9, E_BLOCK_LENGTH_2 = -10, E_TRANSFORM = -11, E_DICTIONARY = -12, E_WINDOW_BITS = -13,
  E_PADDING_1 = -14, E_PADDING_2 = -15, E_DISTANCE = -16, E_DICTIONARY_NOT_SET = -19, E_INVALID_ARGUMENTS = -20,
  E_ALLOC_CONTEXT_MODES = -21, E_ALLOC_TREE_GROUPS = -22, E_ALLOC_CONTEXT_MAP = -25, E_ALLOC_RING_BUFFER_1 = -26,
  E_ALLOC_RING_BUFFER_2 = -27, E_ALLOC_BLOCK_TYPE_TREES = -30, E_UNREACHABLE = -31
};

/* src/state.rs:533-578.  Note CL_SPACE prints "ERROR_FORMAT_FL_SPACE" (:547). */
const char* oracle_error_string(int c) {
  switch (c) {
    case E_NO_ERROR: return "NO_ERROR";
    case E_SUCCESS: return "SUCCESS";
    case E_NEEDS_MORE_INPUT: return "NEEDS_MORE_INPUT";
    case E_NEEDS_MORE_OUTPUT: return "NEEDS_MORE_OUTPUT";
    case E_EXUBERANT_NIBBLE: return "ERROR_FORMAT_EXUBERANT_NIBBLE";
    case E_RESERVED: return "ERROR_FORMAT_RESERVED";
    case E_EXUBERANT_META_NIBBLE: return "ERROR_FORMAT_EXUBERANT_META_NIBBLE";
    case E_SIMPLE_HUFFMAN_ALPHABET: return "ERROR_FORMAT_SIMPLE_HUFFMAN_ALPHABET";
    case E_SIMPLE_HUFFMAN_SAME: return "ERROR_FORMAT_SIMPLE_HUFFMAN_SAME";
    case E_CL_SPACE: return "ERROR_FORMAT_FL_SPACE";
    case E_HUFFMAN_SPACE: return "ERROR_FORMAT_HUFFMAN_SPACE";
    case E_CONTEXT_MAP_REPEAT: return "ERROR_FORMAT_CONTEXT_MAP_REPEAT";
    case E_BLOCK_LENGTH_1: return "ERROR_FORMAT_BLOCK_LENGTH_1";
    case E_BLOCK_LENGTH_2: return "ERROR_FORMAT_BLOCK_LENGTH_2";
    case E_TRANSFORM: return "ERROR_FORMAT_TRANSFORM";
    case E_DICTIONARY: return "ERROR_FORMAT_DICTIONARY";
    case E_WINDOW_BITS: return "ERROR_FORMAT_WINDOW_BITS";
    case E_PADDING_1: return "ERROR_FORMAT_PADDING_1";
    case E_PADDING_2: return "ERROR_FORMAT_PADDING_2";
    case E_DISTANCE: return "ERROR_FORMAT_DISTANCE";
    case E_DICTIONARY_NOT_SET: return "ERROR_DICTIONARY_NOT_SET";
    case E_INVALID_ARGUMENTS: return "ERROR_INVALID_ARGUMENTS";
    case E_ALLOC_CONTEXT_MODES: return "ERROR_ALLOC_CONTEXT_MODES";
    case E_ALLOC_TREE_GROUPS: return "ERROR_ALLOC_TREE_GROUPS";
    case E_ALLOC_CONTEXT_MAP: return "ERROR_ALLOC_CONTEXT_MAP";
    case E_ALLOC_RING_BUFFER_1: return "ERROR_ALLOC_RING_BUFFER_1";
    case E_ALLOC_RING_BUFFER_2: return "ERROR_ALLOC_RING_BUFFER_2";
    case E_ALLOC_BLOCK_TYPE_TREES: return "ERROR_ALLOC_BLOCK_TYPE_TREES";
    case E_UNREACHABLE: return "ERROR_UNREACHABLE";
    default: return "ERROR_UNREACHABLE";
  }
}

/* ---- constants, src/decode.rs:41-61,129-132; src/huffman/mod.rs:10-26 ---- */
#define kBrotliWindowGap 16
#define kBrotliLargeMinWbits 10
#define kBrotliLargeMaxWbits 30
#define kBrotliMaxAllowedDistance 0x7FFFFFFC
#define kDefaultCodeLength 8
#define kCodeLengthRepeatCode 16
#define kNumLiteralCodes 256
#define kNumInsertAndCopyCodes 704
#define kNumBlockLengthCodes 26
#define kLiteralContextBits 6
#define kDistanceContextBits 2
#define HUFFMAN_TABLE_BITS 8
#define HUFFMAN_TABLE_MASK 0xff
#define CODE_LENGTH_CODES 18
#define NUM_DISTANCE_SHORT_CODES 16
#define BROTLI_MAX_DISTANCE_BITS 24
#define BROTLI_LARGE_MAX_DISTANCE_BITS 62
#define HUFFMAN_MAX_CODE_LENGTH 15
#define HUFFMAN_MAX_CODE_LENGTHS_SIZE 704
#define HUFFMAN_MAX_TABLE_SIZE 1080
#define HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH 5
static const uint8_t kCodeLengthCodeOrder[CODE_LENGTH_CODES] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
static const uint8_t kCodeLengthPrefixLength[16] = {2, 2, 2, 3, 2, 2, 2, 4, 2, 2, 2, 3, 2, 2, 2, 4};
static const uint8_t kCodeLengthPrefixValue[16] = {0, 4, 3, 2, 0, 4, 3, 1, 0, 4, 3, 2, 0, 4, 3, 5};

typedef OracleHuffmanCode HC;
typedef OracleBitReader BR;

/* ======================= bit reader: src/bit_reader/mod.rs ======================= */
static inline uint32_t BitMask(uint32_t n) { return n >= 32 ? 0xFFFFFFFFu : ((1u << n) - 1u); } /* :23-34 */
static inline uint32_t GetAvailableBits(const BR* br) { return 64u - br->bit_pos_; }              /* :88-90 */
static inline uint32_t GetRemainingBytes(const BR* br) { return br->avail_in + (GetAvailableBits(br) >> 3); } /* :92 */
static inline int CheckInputAmount(const BR* br, uint32_t num) { return br->avail_in >= num; }    /* :96 */
static inline uint32_t Load32LE(const uint8_t* in, uint32_t p) {                                   /* :110-118 */
  return (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24);
}
static inline uint64_t Load64LE(const uint8_t* in, uint32_t p) {                                   /* :121-131 */
  return (uint64_t)Load32LE(in, p) | ((uint64_t)Load32LE(in, p + 4) << 32);
}
/* BrotliFillBitWindow (:135-171) and ...CompileTimeNbits (:174-219); 64-bit reg_t, unaligned reads. */
static inline void FillBitWindow(BR* br, uint32_t n_bits, const uint8_t* in) {
  if (n_bits <= 8 && br->bit_pos_ >= 56) {
    br->val_ >>= 56; br->bit_pos_ ^= 56; br->val_ |= Load64LE(in, br->next_in) << 8; br->avail_in -= 7; br->next_in += 7;
  } else if (n_bits <= 16 && br->bit_pos_ >= 48) {
    br->val_ >>= 48; br->bit_pos_ ^= 48; br->val_ |= Load64LE(in, br->next_in) << 16; br->avail_in -= 6; br->next_in += 6;
  } else if (br->bit_pos_ >= 32) {
    br->val_ >>= 32; br->bit_pos_ ^= 32; br->val_ |= (uint64_t)Load32LE(in, br->next_in) << 32; br->avail_in -= 4; br->next_in += 4;
  }
}
static inline void FillBitWindowCT(BR* br, uint32_t n_bits, const uint8_t* in) {
  if (n_bits <= 8) {
    if (br->bit_pos_ >= 56) { br->val_ >>= 56; br->bit_pos_ ^= 56; br->val_ |= Load64LE(in, br->next_in) << 8; br->avail_in -= 7; br->next_in += 7; }
  } else if (n_bits <= 16) {
    if (br->bit_pos_ >= 48) { br->val_ >>= 48; br->bit_pos_ ^= 48; br->val_ |= Load64LE(in, br->next_in) << 16; br->avail_in -= 6; br->next_in += 6; }
  } else if (br->bit_pos_ >= 32) {
    br->val_ >>= 32; br->bit_pos_ ^= 32; br->val_ |= (uint64_t)Load32LE(in, br->next_in) << 32; br->avail_in -= 4; br->next_in += 4;
  }
}
static inline void FillBitWindow16(BR* br, const uint8_t* in) { FillBitWindowCT(br, 17, in); } /* :221-223 */
static inline int PullByte(BR* br, const uint8_t* in) {                                         /* :228-242 */
  if (br->avail_in == 0) return 0;
  br->val_ >>= 8; br->val_ |= (uint64_t)in[br->next_in] << 56; br->bit_pos_ -= 8; br->avail_in -= 1; br->next_in += 1;
  return 1;
}
static inline uint64_t GetBitsUnmasked(const BR* br) { return br->bit_pos_ >= 64 ? 0 : br->val_ >> br->bit_pos_; } /* :247 */
static inline uint32_t Get16BitsUnmasked(BR* br, const uint8_t* in) {                             /* :253-257 */
  FillBitWindowCT(br, 16, in); return (uint32_t)(GetBitsUnmasked(br) & 0xffffffffu);
}
static inline uint32_t GetBits(BR* br, uint32_t n, const uint8_t* in) {                           /* :261-264 */
  FillBitWindow(br, n, in); return (uint32_t)GetBitsUnmasked(br) & BitMask(n);
}
static inline int SafeGetBits(BR* br, uint32_t n, uint32_t* val, const uint8_t* in) {             /* :275-287 */
  while (GetAvailableBits(br) < n) if (!PullByte(br, in)) return 0;
  *val = (uint32_t)GetBitsUnmasked(br) & BitMask(n); return 1;
}
static inline void DropBits(BR* br, uint32_t n) { br->bit_pos_ += n; }                            /* :291 */
static inline void TakeBits(BR* br, uint32_t n, uint32_t* val) {                                  /* :309-315 */
  *val = (uint32_t)GetBitsUnmasked(br) & BitMask(n); DropBits(br, n);
}
static inline uint32_t ReadBits(BR* br, uint32_t n, const uint8_t* in) {                          /* :318-338 */
  uint32_t v; FillBitWindow(br, n, in); TakeBits(br, n, &v); return v;
}
static inline int SafeReadBits(BR* br, uint32_t n, uint32_t* val, const uint8_t* in) {            /* :362-374 */
  while (GetAvailableBits(br) < n) if (!PullByte(br, in)) return 0;
  TakeBits(br, n, val); return 1;
}
static inline int JumpToByteBoundary(BR* br) {                                                    /* :378-385 */
  uint32_t pad_bits_count = GetAvailableBits(br) & 7, pad_bits = 0;
  if (pad_bits_count) TakeBits(br, pad_bits_count, &pad_bits);
  return pad_bits == 0;
}
static int PeekByte(BR* br, uint32_t offset, const uint8_t* in) {                                 /* :391-403 */
  uint32_t available_bits = GetAvailableBits(br), bytes_left = available_bits >> 3;
  if (offset < bytes_left) return (int)((GetBitsUnmasked(br) >> (offset << 3)) & 0xFF);
  offset -= bytes_left;
  if (offset < br->avail_in) return in[br->next_in + offset];
  return -1;
}
static void CopyBytes(uint8_t* dest, BR* br, uint32_t num, const uint8_t* in) {                   /* :408-422 */
  uint32_t offset = 0;
  while (GetAvailableBits(br) >= 8 && num > 0) { dest[offset++] = (uint8_t)GetBitsUnmasked(br); DropBits(br, 8); num--; }
  memcpy(dest + offset, in + br->next_in, num);
  br->avail_in -= num; br->next_in += num;
}
static inline int WarmupBitReader(BR* br, const uint8_t* in) {                                    /* :429-446 */
  if (GetAvailableBits(br) == 0 && !PullByte(br, in)) return 0;
  return 1;
}
typedef struct { uint64_t val_; uint32_t bit_pos_, next_in, avail_in; } BRState;                   /* :57-86 */
static inline BRState BRSave(const BR* br) { BRState m = {br->val_, br->bit_pos_, br->next_in, br->avail_in}; return m; }
static inline void BRRestore(BR* br, const BRState* m) { br->val_ = m->val_; br->bit_pos_ = m->bit_pos_; br->next_in = m->next_in; br->avail_in = m->avail_in; }

/* exported KAT shims */
int oracle_br_warmup(BR* br, const uint8_t* in) { return WarmupBitReader(br, in); }
int oracle_br_safe_read_bits(BR* br, uint32_t n, uint32_t* v, const uint8_t* in) { return SafeReadBits(br, n, v, in); }
uint32_t oracle_br_read_bits(BR* br, uint32_t n, const uint8_t* in) { return ReadBits(br, n, in); }
uint32_t oracle_br_read_constant_n_bits(BR* br, uint32_t n, const uint8_t* in) { /* :341-360 */
  uint32_t v; FillBitWindowCT(br, n, in); TakeBits(br, n, &v); return v;
}
uint32_t oracle_br_get16_bits_unmasked(BR* br, const uint8_t* in) { return Get16BitsUnmasked(br, in); }

/* ======================= Huffman table builders: src/huffman/mod.rs ======================= */
static inline uint32_t ReverseBits8(uint32_t num) { /* kReverseBits, :135-161 */
  num = ((num & 0xF0) >> 4) | ((num & 0x0F) << 4);
  num = ((num & 0xCC) >> 2) | ((num & 0x33) << 2);
  num = ((num & 0xAA) >> 1) | ((num & 0x55) << 1);
  return num;
}
#define REVERSE_BITS_LOWEST 0x80u
static inline void ReplicateValue(HC* table, uint32_t offset, int step, int end, HC code) { /* :165-176 */
  do { end -= step; table[offset + (uint32_t)end] = code; } while (end > 0);
}
static inline int NextTableBitSize(const uint16_t* count, int len, int root_bits) {          /* :181-193 */
  int left = 1 << (len - root_bits);
  while (len < HUFFMAN_MAX_CODE_LENGTH) {
    left -= count[len];
    if (left <= 0) break;
    len++; left <<= 1;
  }
  return len - root_bits;
}
/* BrotliBuildCodeLengthsHuffmanTable, :196-271 */
void oracle_build_code_lengths_huffman_table(HC* table, const uint8_t* code_lengths, const uint16_t* count) {
  int sorted[18], offset[HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH + 1];
  int symbol = -1, bits, step;
  const int table_size = 1 << HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH;
  uint32_t key, key_step;
  for (bits = 1; bits <= HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH; bits++) { symbol += count[bits]; offset[bits] = symbol; }
  offset[0] = 17; /* zero-length symbols go last */
  symbol = 18;
  do { symbol--; sorted[offset[code_lengths[symbol]]--] = symbol; } while (symbol != 0);
  if (offset[0] == 0) { /* all but one symbol have zero length */
    HC code; code.bits = 0; code.value = (uint16_t)sorted[0];
    for (key = 0; key < (uint32_t)table_size; key++) table[key] = code;
    return;
  }
  key = 0; key_step = REVERSE_BITS_LOWEST; symbol = 0; bits = 1; step = 2;
  do {
    HC code; int bits_count = count[bits];
    code.bits = (uint8_t)bits; code.value = 0;
    for (; bits_count != 0; bits_count--) {
      code.value = (uint16_t)sorted[symbol++];
      ReplicateValue(table, ReverseBits8(key), step, table_size, code);
      key += key_step;
    }
    step <<= 1; key_step >>= 1;
  } while (++bits <= HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH);
}
/* BrotliBuildHuffmanTable, :273-386.  symbol_lists is indexed at negative offsets from
 * symbol_lists_offset (per-length list heads at [-16..-1], links at [0..)). */
uint32_t oracle_build_huffman_table(HC* root_table, int root_bits, const uint16_t* symbol_lists, size_t slo, uint16_t* count) {
  HC code; int max_length = -1, table_bits, table_size, total_size, bits, step, len;
  uint32_t table_free_offset = 0, key, key_step, sub_key, sub_key_step;
  code.bits = 0; code.value = 0;
  while (symbol_lists[(ptrdiff_t)slo + max_length] == 0xFFFF) max_length--;
  max_length += HUFFMAN_MAX_CODE_LENGTH + 1;
  table_bits = root_bits; table_size = 1 << table_bits; total_size = table_size;
  if (table_bits > max_length) { table_bits = max_length; table_size = 1 << table_bits; }
  key = 0; key_step = REVERSE_BITS_LOWEST; bits = 1; step = 2;
  do {
    int symbol = bits - (HUFFMAN_MAX_CODE_LENGTH + 1), bits_count = count[bits];
    code.bits = (uint8_t)bits;
    for (; bits_count != 0; bits_count--) {
      symbol = symbol_lists[(ptrdiff_t)slo + symbol];
      code.value = (uint16_t)symbol;
      ReplicateValue(root_table, table_free_offset + ReverseBits8(key), step, table_size, code);
      key += key_step;
    }
    step <<= 1; key_step >>= 1;
  } while (++bits <= table_bits);
  while (total_size != table_size) { /* replicate the partial root table */
    memcpy(&root_table[table_free_offset + (uint32_t)table_size], &root_table[table_free_offset], (size_t)table_size * sizeof(HC));
    table_size <<= 1;
  }
  key_step = REVERSE_BITS_LOWEST >> (root_bits - 1);
  sub_key = REVERSE_BITS_LOWEST << 1; sub_key_step = REVERSE_BITS_LOWEST; step = 2;
  for (len = root_bits + 1; len <= max_length; len++) {
    int symbol = len - (HUFFMAN_MAX_CODE_LENGTH + 1);
    for (; count[len] != 0; count[len]--) {
      if (sub_key == (REVERSE_BITS_LOWEST << 1)) {
        table_free_offset += (uint32_t)table_size;
        table_bits = NextTableBitSize(count, len, root_bits);
        table_size = 1 << table_bits; total_size += table_size;
        sub_key = ReverseBits8(key); key += key_step;
        root_table[sub_key].bits = (uint8_t)(table_bits + root_bits);
        root_table[sub_key].value = (uint16_t)(table_free_offset - sub_key);
        sub_key = 0;
      }
      code.bits = (uint8_t)(len - root_bits);
      symbol = symbol_lists[(ptrdiff_t)slo + symbol];
      code.value = (uint16_t)symbol;
      ReplicateValue(root_table, table_free_offset + ReverseBits8(sub_key), step, table_size, code);
      sub_key += sub_key_step;
    }
    step <<= 1; sub_key_step >>= 1;
  }
  return (uint32_t)total_size;
}
/* BrotliBuildSimpleHuffmanTable, :390-471.  num_symbols is NSYM-1 (0..3), 4 = NSYM 4 with tree-select. */
uint32_t oracle_build_simple_huffman_table(HC* table, int root_bits, const uint16_t* val, size_t val_len, uint32_t num_symbols) {
  uint32_t table_size = 1, goal_size = 1u << root_bits, i, k;
  if (num_symbols == 0) {
    table[0].bits = 0; table[0].value = val[0];
  } else if (num_symbols == 1) {
    table[0].bits = 1; table[1].bits = 1;
    if (val[1] > val[0]) { table[0].value = val[0]; table[1].value = val[1]; }
    else { table[0].value = val[1]; table[1].value = val[0]; }
    table_size = 2;
  } else if (num_symbols == 2) {
    table[0].bits = 1; table[0].value = val[0]; table[2].bits = 1; table[2].value = val[0];
    if (val[2] > val[1]) { table[1].value = val[1]; table[3].value = val[2]; }
    else { table[1].value = val[2]; table[3].value = val[1]; }
    table[1].bits = 2; table[3].bits = 2; table_size = 4;
  } else if (num_symbols == 3) {
    uint16_t m[4]; m[0] = val[0]; m[1] = val[1]; m[2] = val[2]; m[3] = val_len > 3 ? val[3] : 65535;
    for (i = 0; i < 3; i++) for (k = i + 1; k < 4; k++) if (m[k] < m[i]) { uint16_t t = m[k]; m[k] = m[i]; m[i] = t; }
    for (i = 0; i < 4; i++) table[i].bits = 2;
    table[0].value = m[0]; table[2].value = m[1]; table[1].value = m[2]; table[3].value = m[3];
    table_size = 4;
  } else { /* num_symbols == 4 */
    uint16_t m[4]; m[0] = val[0]; m[1] = val[1]; m[2] = val[2]; m[3] = val[3];
    if (m[3] < m[2]) { uint16_t t = m[3]; m[3] = m[2]; m[2] = t; }
    for (i = 0; i < 7; i++) { table[i].value = m[0]; table[i].bits = (uint8_t)(1 + (i & 1)); }
    table[1].value = m[1]; table[3].value = m[2]; table[5].value = m[1]; table[7].value = m[3];
    table[3].bits = 3; table[7].bits = 3; table_size = 8;
  }
  while (table_size != goal_size) { memcpy(&table[table_size], &table[0], table_size * sizeof(HC)); table_size <<= 1; }
  return goal_size;
}

/* ======================= dictionary transforms: src/transform.rs:720-795 ======================= */
static int ToUpperCase(uint8_t* p) {
  if (p[0] < 0xc0) { if (p[0] >= 'a' && p[0] <= 'z') p[0] ^= 32; return 1; }
  if (p[0] < 0xe0) { p[1] ^= 32; return 2; }
  p[2] ^= 5; return 3;
}
int oracle_transform_dictionary_word(uint8_t* dst, const uint8_t* word, int len, int transform) {
  int idx = 0, i = 0, skip;
  const uint8_t* prefix = &kBrotliPrefixSuffix[kBrotliTransforms[transform * 3]];
  const uint8_t t = kBrotliTransforms[transform * 3 + 1];
  const uint8_t* suffix = &kBrotliPrefixSuffix[kBrotliTransforms[transform * 3 + 2]];
  while (prefix[idx]) { dst[idx] = prefix[idx]; idx++; }
  skip = t < BROTLI_TRANSFORM_OMIT_FIRST_1 ? 0 : t - (BROTLI_TRANSFORM_OMIT_FIRST_1 - 1);
  if (skip > len) skip = len;
  word += skip; len -= skip;
  if (t <= BROTLI_TRANSFORM_OMIT_LAST_9) len -= t;
  while (i < len) { dst[idx++] = word[i++]; }
  if (t == BROTLI_TRANSFORM_UPPERCASE_FIRST) {
    ToUpperCase(&dst[idx - len]);
  } else if (t == BROTLI_TRANSFORM_UPPERCASE_ALL) {
    uint8_t* up = &dst[idx - len];
    while (len > 0) { int step = ToUpperCase(up); up += step; len -= step; }
  }
  for (i = 0; suffix[i]; i++) dst[idx++] = suffix[i];
  return idx;
}

/* ======================= decoder state: src/state.rs:156-278,279-450 ======================= */
enum RunningState { /* src/state.rs:68-94 */
  ST_UNINITED, ST_LARGE_WINDOW_BITS, ST_INITIALIZE, ST_METABLOCK_BEGIN, ST_METABLOCK_HEADER, ST_METABLOCK_HEADER_2,
  ST_CONTEXT_MODES, ST_COMMAND_BEGIN, ST_COMMAND_INNER, ST_COMMAND_POST_DECODE_LITERALS, ST_COMMAND_POST_WRAP_COPY,
  ST_UNCOMPRESSED, ST_METADATA, ST_COMMAND_INNER_WRITE, ST_METABLOCK_DONE, ST_COMMAND_POST_WRITE_1,
  ST_COMMAND_POST_WRITE_2, ST_HUFFMAN_CODE_0, ST_HUFFMAN_CODE_1, ST_HUFFMAN_CODE_2, ST_HUFFMAN_CODE_3, ST_CONTEXT_MAP_1, ST_CONTEXT_MAP_2, ST_TREE_GROUP, ST_DONE
};

typedef struct HGroup { /* HuffmanTreeGroup, src/huffman/mod.rs:51-72 */
  uint32_t* htrees; HC* codes; uint16_t alphabet_size, max_symbol, num_htrees;
} HGroup;

typedef struct State {
  const uint8_t* input;
  uint8_t* output; size_t available_out, output_offset, total_out;
  int state, loop_counter;
  BR br;
  int pos, max_backward_distance, max_backward_distance_minus_custom_dict_size, max_distance;
  int ringbuffer_size, ringbuffer_mask, dist_rb_idx, dist_rb[4];
  uint8_t* ringbuffer; size_t ringbuffer_len;
  uint16_t htree_command_index;
  const uint8_t* context_lookup;
  size_t context_map_slice_index, dist_context_map_slice_index;
  HGroup literal_hgroup, insert_copy_hgroup, distance_hgroup;
  int trivial_literal_context, distance_context, meta_block_remaining_len;
  uint32_t block_length[3], num_block_types[3], block_type_rb[6];
  HC *block_type_trees, *block_len_trees;
  uint32_t distance_postfix_bits, num_direct_distance_codes; int distance_postfix_mask;
  uint32_t num_dist_htrees; uint8_t* dist_context_map;
  uint8_t literal_htree_index, dist_htree_index;
  int copy_length, distance_code;
  size_t rb_roundtrips, partial_pos_out;
  /* ReadHuffmanCode scratch */
  uint32_t symbol, repeat, space, prev_code_len, repeat_code_len;
  HC table[32];
  uint16_t symbols_lists_array[HUFFMAN_MAX_CODE_LENGTH + 1 + 2048];
  int next_symbol[32];
  uint8_t code_length_code_lengths[18];
  uint16_t code_length_histo[16];
  HC* context_map_table;
  uint32_t mtf_upper_bound; uint8_t mtf[256];
  const uint8_t* custom_dict; ptrdiff_t custom_dict_size; int custom_dict_avoid_context_seed;
  uint8_t is_last_metablock, is_uncompressed, is_metadata, size_nibbles;
  uint32_t window_bits; int large_window, canny_ringbuffer_allocation, should_wrap_ringbuffer;
  uint32_t num_literal_htrees; uint8_t *context_map, *context_modes;
  uint32_t trivial_literal_contexts[8];
  /* sub-states that let every header function resume where NeedsMoreInput interrupted it, src/state.rs:96-154,160-178 */
  int substate_decode_uint8, substate_metablock_header, substate_huffman, substate_tree_group, substate_context_map,
      substate_read_block_length;
  uint32_t sub_loop_counter, htree_index, htree_next_offset, context_index, max_run_length_prefix, code, block_length_index;
  /* carry-over of an interrupted read across calls, src/state.rs:169-170 */
  uint8_t buffer[8]; uint32_t buffer_length;
  int error_code;
} State;
#define SYMBOL_LISTS_INDEX (HUFFMAN_MAX_CODE_LENGTH + 1) /* src/state.rs:352 */

static void HGroupReset(HGroup* g) { free(g->htrees); free(g->codes); memset(g, 0, sizeof(*g)); }
static int HGroupInit(HGroup* g, uint16_t alphabet_size, uint16_t max_symbol, uint16_t ntrees) { /* huffman/mod.rs:61-72 */
  HGroupReset(g);
  g->alphabet_size = alphabet_size; g->max_symbol = max_symbol; g->num_htrees = ntrees;
  g->htrees = (uint32_t*)calloc(ntrees ? ntrees : 1, sizeof(uint32_t));
  g->codes = (HC*)calloc((size_t)(ntrees ? ntrees : 1) * HUFFMAN_MAX_TABLE_SIZE, sizeof(HC));
  return g->htrees && g->codes;
}

/* BrotliStateMetablockBegin, src/state.rs:422-450 */
static void StateMetablockBegin(State* s) {
  int i;
  s->meta_block_remaining_len = 0;
  for (i = 0; i < 3; i++) { s->block_length[i] = 1u << 24; s->num_block_types[i] = 1; s->block_type_rb[2 * i] = 1; s->block_type_rb[2 * i + 1] = 0; }
  free(s->context_map); s->context_map = NULL;
  free(s->context_modes); s->context_modes = NULL;
  free(s->dist_context_map); s->dist_context_map = NULL;
  s->context_map_slice_index = 0; s->literal_htree_index = 0; s->dist_context_map_slice_index = 0; s->dist_htree_index = 0;
  s->context_lookup = &kBrotliContextLookup[0];
  HGroupReset(&s->literal_hgroup); HGroupReset(&s->insert_copy_hgroup); HGroupReset(&s->distance_hgroup);
}
/* BrotliStateCleanupAfterMetablock, src/state.rs:451-463 */
static void StateCleanupAfterMetablock(State* s) {
  free(s->context_map); s->context_map = NULL;
  free(s->context_modes); s->context_modes = NULL;
  free(s->dist_context_map); s->dist_context_map = NULL;
  HGroupReset(&s->literal_hgroup); HGroupReset(&s->insert_copy_hgroup); HGroupReset(&s->distance_hgroup);
}
static void StateCleanup(State* s) { /* src/state.rs:465-480 */
  StateCleanupAfterMetablock(s);
  free(s->ringbuffer); free(s->block_type_trees); free(s->block_len_trees); free(s->context_map_table);
}

/* ======================= src/decode.rs ======================= */
/* DecodeWindowBits, :152-187 */
static int DecodeWindowBits(int* s_large_window, uint32_t* window_bits, BR* br) {
  uint32_t n; int large_window = *s_large_window;
  *s_large_window = 0;
  TakeBits(br, 1, &n);
  if (n == 0) { *window_bits = 16; return E_SUCCESS; }
  TakeBits(br, 3, &n);
  if (n != 0) { *window_bits = 17 + n; return E_SUCCESS; }
  TakeBits(br, 3, &n);
  if (n == 1) {
    if (large_window) {
      TakeBits(br, 1, &n);
      if (n == 1) return E_WINDOW_BITS;
      *s_large_window = 1; return E_SUCCESS;
    }
    return E_WINDOW_BITS;
  }
  if (n != 0) { *window_bits = 8 + n; return E_SUCCESS; }
  *window_bits = 17; return E_SUCCESS;
}
/* DecodeVarLenUint8, :193-241 (sub-states BROTLI_STATE_DECODE_UINT8_{NONE,SHORT,LONG}) */
static int DecodeVarLenUint8(State* s, uint32_t* value) {
  uint32_t bits;
  for (;;) {
    switch (s->substate_decode_uint8) {
      case 0:
        if (!SafeReadBits(&s->br, 1, &bits, s->input)) return E_NEEDS_MORE_INPUT;
        if (bits == 0) { *value = 0; return E_SUCCESS; }
        s->substate_decode_uint8 = 1;
        /* fall through */
      case 1:
        if (!SafeReadBits(&s->br, 3, &bits, s->input)) { s->substate_decode_uint8 = 1; return E_NEEDS_MORE_INPUT; }
        if (bits == 0) { *value = 1; s->substate_decode_uint8 = 0; return E_SUCCESS; }
        *value = bits; /* the output value is the temporary storage and persists across calls */
        s->substate_decode_uint8 = 2;
        /* fall through */
      default:
        if (!SafeReadBits(&s->br, *value, &bits, s->input)) { s->substate_decode_uint8 = 2; return E_NEEDS_MORE_INPUT; }
        *value = (1u << *value) + bits;
        s->substate_decode_uint8 = 0;
        return E_SUCCESS;
    }
  }
}
/* DecodeMetaBlockLength, :243-372 (sub-states BROTLI_STATE_METABLOCK_HEADER_*) */
enum { MH_NONE, MH_EMPTY, MH_NIBBLES, MH_SIZE, MH_UNCOMPRESSED, MH_RESERVED, MH_BYTES, MH_METADATA };
static int DecodeMetaBlockLength(State* s) {
  uint32_t bits; int i;
  for (;;) {
    switch (s->substate_metablock_header) {
      case MH_NONE:
        if (!SafeReadBits(&s->br, 1, &bits, s->input)) return E_NEEDS_MORE_INPUT;
        s->is_last_metablock = (uint8_t)bits; s->meta_block_remaining_len = 0; s->is_uncompressed = 0; s->is_metadata = 0;
        if (!s->is_last_metablock) { s->substate_metablock_header = MH_NIBBLES; continue; }
        s->substate_metablock_header = MH_EMPTY;
        /* fall through */
      case MH_EMPTY:
        if (!SafeReadBits(&s->br, 1, &bits, s->input)) return E_NEEDS_MORE_INPUT; /* ISLASTEMPTY */
        if (bits) { s->substate_metablock_header = MH_NONE; return E_SUCCESS; }
        s->substate_metablock_header = MH_NIBBLES;
        /* fall through */
      case MH_NIBBLES:
        if (!SafeReadBits(&s->br, 2, &bits, s->input)) return E_NEEDS_MORE_INPUT; /* MNIBBLES */
        s->size_nibbles = (uint8_t)(bits + 4);
        s->loop_counter = 0;
        if (bits == 3) { s->is_metadata = 1; s->substate_metablock_header = MH_RESERVED; continue; }
        s->substate_metablock_header = MH_SIZE;
        /* fall through */
      case MH_SIZE:
        for (i = s->loop_counter; i < s->size_nibbles; i++) {
          if (!SafeReadBits(&s->br, 4, &bits, s->input)) { s->loop_counter = i; return E_NEEDS_MORE_INPUT; }
          if (i + 1 == s->size_nibbles && s->size_nibbles > 4 && bits == 0) return E_EXUBERANT_NIBBLE;
          s->meta_block_remaining_len |= (int)(bits << (i * 4));
        }
        s->substate_metablock_header = MH_UNCOMPRESSED;
        /* fall through */
      case MH_UNCOMPRESSED:
        if (!s->is_last_metablock && !s->is_metadata) {
          if (!SafeReadBits(&s->br, 1, &bits, s->input)) return E_NEEDS_MORE_INPUT;
          s->is_uncompressed = (uint8_t)bits;
        }
        s->meta_block_remaining_len += 1;
        s->substate_metablock_header = MH_NONE;
        return E_SUCCESS;
      case MH_RESERVED:
        if (!SafeReadBits(&s->br, 1, &bits, s->input)) return E_NEEDS_MORE_INPUT;
        if (bits != 0) return E_RESERVED;
        s->substate_metablock_header = MH_BYTES;
        /* fall through */
      case MH_BYTES:
        if (!SafeReadBits(&s->br, 2, &bits, s->input)) return E_NEEDS_MORE_INPUT; /* MSKIPBYTES */
        if (bits == 0) { s->substate_metablock_header = MH_NONE; return E_SUCCESS; }
        s->size_nibbles = (uint8_t)bits;
        s->substate_metablock_header = MH_METADATA;
        /* fall through */
      default: /* MH_METADATA */
        for (i = s->loop_counter; i < s->size_nibbles; i++) {
          if (!SafeReadBits(&s->br, 8, &bits, s->input)) { s->loop_counter = i; return E_NEEDS_MORE_INPUT; }
          if (i + 1 == s->size_nibbles && s->size_nibbles > 1 && bits == 0) return E_EXUBERANT_META_NIBBLE;
          s->meta_block_remaining_len |= (int)(bits << (i * 8));
        }
        s->substate_metablock_header = MH_UNCOMPRESSED;
        continue;
    }
  }
}
/* DecodeSymbol :377-391, ReadSymbol :395-398 */
static inline uint32_t DecodeSymbol(uint32_t bits, const HC* table, BR* br) {
  uint32_t table_index = bits & HUFFMAN_TABLE_MASK;
  HC e = table[table_index];
  if (e.bits > HUFFMAN_TABLE_BITS) {
    uint32_t nbits = e.bits - HUFFMAN_TABLE_BITS;
    DropBits(br, HUFFMAN_TABLE_BITS);
    table_index += e.value;
    e = table[table_index + ((bits >> HUFFMAN_TABLE_BITS) & BitMask(nbits))];
  }
  DropBits(br, e.bits);
  return e.value;
}
static inline uint32_t ReadSymbol(const HC* table, BR* br, const uint8_t* in) { return DecodeSymbol(Get16BitsUnmasked(br, in), table, br); }
/* SafeDecodeSymbol :402-441, SafeReadSymbol :443-456 */
static int SafeDecodeSymbol(const HC* table, BR* br, uint32_t* result) {
  uint32_t available_bits = GetAvailableBits(br), val, table_index;
  HC e, sub;
  if (available_bits == 0) {
    if (table[0].bits == 0) { *result = table[0].value; return 1; }
    return 0;
  }
  val = (uint32_t)GetBitsUnmasked(br);
  table_index = val & HUFFMAN_TABLE_MASK;
  e = table[table_index];
  if (e.bits <= HUFFMAN_TABLE_BITS) {
    if (e.bits <= available_bits) { DropBits(br, e.bits); *result = e.value; return 1; }
    return 0;
  }
  if (available_bits <= HUFFMAN_TABLE_BITS) return 0;
  val = (val & BitMask(e.bits)) >> HUFFMAN_TABLE_BITS;
  available_bits -= HUFFMAN_TABLE_BITS;
  sub = table[table_index + e.value + val];
  if (available_bits < sub.bits) return 0;
  DropBits(br, HUFFMAN_TABLE_BITS + sub.bits);
  *result = sub.value;
  return 1;
}
static int SafeReadSymbol(const HC* table, BR* br, uint32_t* result, const uint8_t* in) {
  uint32_t val;
  if (SafeGetBits(br, 15, &val, in)) { *result = DecodeSymbol(val, table, br); return 1; }
  return SafeDecodeSymbol(table, br, result);
}
static uint32_t Log2Floor(uint32_t x) { uint32_t r = 0; while (x) { x >>= 1; r++; } return r; } /* :502-509 */

/* ReadSimpleHuffmanSymbols, :516-556 (resumes at sub_loop_counter) */
enum { HS_NONE, HS_SIMPLE_SIZE, HS_SIMPLE_READ, HS_SIMPLE_BUILD, HS_COMPLEX, HS_LENGTH_SYMBOLS };
static int ReadSimpleHuffmanSymbols(uint32_t alphabet_size, uint32_t max_symbol, State* s) {
  uint32_t max_bits = Log2Floor(alphabet_size - 1), i, k, num_symbols = s->symbol;
  for (i = s->sub_loop_counter; i <= num_symbols; i++) {
    uint32_t v;
    if (!SafeReadBits(&s->br, max_bits, &v, s->input)) { s->sub_loop_counter = i; s->substate_huffman = HS_SIMPLE_READ; return E_NEEDS_MORE_INPUT; }
    if (v >= max_symbol) return E_SIMPLE_HUFFMAN_ALPHABET;
    s->symbols_lists_array[i] = (uint16_t)v;
  }
  for (i = 0; i < num_symbols; i++)
    for (k = i + 1; k <= num_symbols; k++)
      if (s->symbols_lists_array[i] == s->symbols_lists_array[k]) return E_SIMPLE_HUFFMAN_SAME;
  return E_SUCCESS;
}
/* ProcessSingleCodeLength, :565-589 */
static inline void ProcessSingleCodeLength(uint32_t code_len, uint32_t* symbol, uint32_t* repeat, uint32_t* space,
                                           uint32_t* prev_code_len, uint16_t* symbol_lists /* at index offset */,
                                           uint16_t* code_length_histo, int* next_symbol) {
  *repeat = 0;
  if (code_len != 0) {
    symbol_lists[next_symbol[code_len]] = (uint16_t)*symbol;
    next_symbol[code_len] = (int)*symbol;
    *prev_code_len = code_len;
    *space -= 32768u >> code_len;
    code_length_histo[code_len]++;
  }
  (*symbol)++;
}
/* ProcessRepeatedCodeLength, :600-658 */
static inline void ProcessRepeatedCodeLength(uint32_t code_len, uint32_t repeat_delta, uint32_t alphabet_size, uint32_t* symbol,
                                             uint32_t* repeat, uint32_t* space, uint32_t* prev_code_len,
                                             uint32_t* repeat_code_len, uint16_t* symbol_lists, uint16_t* code_length_histo,
                                             int* next_symbol) {
  uint32_t old_repeat, extra_bits, new_len;
  if (code_len == kCodeLengthRepeatCode) { extra_bits = 2; new_len = *prev_code_len; } else { extra_bits = 3; new_len = 0; }
  if (*repeat_code_len != new_len) { *repeat = 0; *repeat_code_len = new_len; }
  old_repeat = *repeat;
  if (*repeat > 0) { *repeat -= 2; *repeat <<= extra_bits; }
  *repeat += repeat_delta + 3;
  repeat_delta = *repeat - old_repeat;
  if (*symbol + repeat_delta > alphabet_size) { *symbol = alphabet_size; *space = 0xFFFFF; return; }
  if (*repeat_code_len != 0) {
    uint32_t last = *symbol + repeat_delta;
    int next = next_symbol[*repeat_code_len];
    do { symbol_lists[next] = (uint16_t)*symbol; next = (int)*symbol; (*symbol)++; } while (*symbol != last);
    next_symbol[*repeat_code_len] = next;
    *space -= repeat_delta << (15 - *repeat_code_len);
    code_length_histo[*repeat_code_len] = (uint16_t)(code_length_histo[*repeat_code_len] + repeat_delta);
  } else {
    *symbol += repeat_delta;
  }
}
/* ReadSymbolCodeLengths (fast), :661-731 */
static int ReadSymbolCodeLengths(uint32_t alphabet_size, State* s) {
  uint32_t symbol = s->symbol, repeat = s->repeat, space = s->space, prev_code_len = s->prev_code_len, repeat_code_len = s->repeat_code_len;
  uint16_t* lists = &s->symbols_lists_array[SYMBOL_LISTS_INDEX];
  if (!WarmupBitReader(&s->br, s->input)) return E_NEEDS_MORE_INPUT;
  while (symbol < alphabet_size && space > 0) {
    HC p; uint32_t code_len;
    if (!CheckInputAmount(&s->br, 4)) {
      s->symbol = symbol; s->repeat = repeat; s->prev_code_len = prev_code_len; s->repeat_code_len = repeat_code_len; s->space = space;
      return E_NEEDS_MORE_INPUT;
    }
    FillBitWindow16(&s->br, s->input);
    p = s->table[GetBitsUnmasked(&s->br) & BitMask(HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH)];
    DropBits(&s->br, p.bits);
    code_len = p.value;
    if (code_len < kCodeLengthRepeatCode) {
      ProcessSingleCodeLength(code_len, &symbol, &repeat, &space, &prev_code_len, lists, s->code_length_histo, s->next_symbol);
    } else {
      uint32_t extra_bits = code_len == kCodeLengthRepeatCode ? 2 : 3;
      uint32_t repeat_delta = (uint32_t)GetBitsUnmasked(&s->br) & BitMask(extra_bits);
      DropBits(&s->br, extra_bits);
      ProcessRepeatedCodeLength(code_len, repeat_delta, alphabet_size, &symbol, &repeat, &space, &prev_code_len, &repeat_code_len,
                                lists, s->code_length_histo, s->next_symbol);
    }
  }
  s->space = space;
  return E_SUCCESS;
}
/* SafeReadSymbolCodeLengths, :733-797 */
static int SafeReadSymbolCodeLengths(uint32_t alphabet_size, State* s) {
  uint16_t* lists = &s->symbols_lists_array[SYMBOL_LISTS_INDEX];
  while (s->symbol < alphabet_size && s->space > 0) {
    uint32_t code_len, bits = 0, available_bits = GetAvailableBits(&s->br);
    HC p;
    if (available_bits != 0) bits = (uint32_t)GetBitsUnmasked(&s->br);
    p = s->table[bits & BitMask(HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH)];
    if (p.bits > available_bits) { if (!PullByte(&s->br, s->input)) return E_NEEDS_MORE_INPUT; continue; }
    code_len = p.value;
    if (code_len < kCodeLengthRepeatCode) {
      DropBits(&s->br, p.bits);
      ProcessSingleCodeLength(code_len, &s->symbol, &s->repeat, &s->space, &s->prev_code_len, lists, s->code_length_histo, s->next_symbol);
    } else {
      uint32_t extra_bits = code_len - 14, repeat_delta = (bits >> p.bits) & BitMask(extra_bits);
      if (available_bits < p.bits + extra_bits) { if (!PullByte(&s->br, s->input)) return E_NEEDS_MORE_INPUT; continue; }
      DropBits(&s->br, p.bits + extra_bits);
      ProcessRepeatedCodeLength(code_len, repeat_delta, alphabet_size, &s->symbol, &s->repeat, &s->space, &s->prev_code_len,
                                &s->repeat_code_len, lists, s->code_length_histo, s->next_symbol);
    }
  }
  return E_SUCCESS;
}
/* ReadCodeLengthCodeLengths, :801-853.  sub_loop_counter holds HSKIP on entry and the resume index afterwards. */
static int ReadCodeLengthCodeLengths(State* s) {
  uint32_t num_codes = s->repeat, space = s->space, i;
  for (i = s->sub_loop_counter; i < CODE_LENGTH_CODES; i++) {
    uint8_t code_len_idx = kCodeLengthCodeOrder[i];
    uint32_t ix = 0, v;
    if (!SafeGetBits(&s->br, 4, &ix, s->input)) {
      uint32_t available_bits = GetAvailableBits(&s->br);
      ix = available_bits != 0 ? ((uint32_t)GetBitsUnmasked(&s->br) & 0xF) : 0;
      if (kCodeLengthPrefixLength[ix] > available_bits) {
        s->sub_loop_counter = i; s->repeat = num_codes; s->space = space; s->substate_huffman = HS_COMPLEX;
        return E_NEEDS_MORE_INPUT;
      }
    }
    v = kCodeLengthPrefixValue[ix];
    DropBits(&s->br, kCodeLengthPrefixLength[ix]);
    s->code_length_code_lengths[code_len_idx] = (uint8_t)v;
    if (v != 0) {
      space -= 32u >> v; num_codes++; s->code_length_histo[v]++;
      if (space - 1u >= 32u) break; /* space is 0 or wrapped */
    }
  }
  if (!(num_codes == 1 || space == 0)) return E_CL_SPACE;
  return E_SUCCESS;
}
/* ReadHuffmanCode, :868-1013 (sub-states BROTLI_STATE_HUFFMAN_*) */
static int ReadHuffmanCode(uint32_t alphabet_size, uint32_t max_symbol, HC* table, uint32_t* opt_table_size, State* s) {
  uint32_t table_size; int r, i;
  alphabet_size &= 0x7ff;
  for (;;) {
    switch (s->substate_huffman) {
      case HS_NONE:
        if (!SafeReadBits(&s->br, 2, &s->sub_loop_counter, s->input)) return E_NEEDS_MORE_INPUT;
        if (s->sub_loop_counter != 1) { /* 0: no skipping, 2 / 3: skip that many code lengths */
          s->space = 32; s->repeat = 0;
          memset(s->code_length_histo, 0, sizeof(uint16_t) * (HUFFMAN_MAX_CODE_LENGTH_CODE_LENGTH + 1));
          memset(s->code_length_code_lengths, 0, sizeof(s->code_length_code_lengths));
          s->substate_huffman = HS_COMPLEX;
          continue;
        }
        s->substate_huffman = HS_SIMPLE_SIZE;
        /* fall through */
      case HS_SIMPLE_SIZE:
        if (!SafeReadBits(&s->br, 2, &s->symbol, s->input)) { s->substate_huffman = HS_SIMPLE_SIZE; return E_NEEDS_MORE_INPUT; } /* NSYM-1 */
        s->sub_loop_counter = 0;
        s->substate_huffman = HS_SIMPLE_READ;
        /* fall through */
      case HS_SIMPLE_READ:
        r = ReadSimpleHuffmanSymbols(alphabet_size, max_symbol, s);
        if (r != E_SUCCESS) return r;
        s->substate_huffman = HS_SIMPLE_BUILD;
        /* fall through */
      case HS_SIMPLE_BUILD:
        if (s->symbol == 3) {
          uint32_t bits;
          if (!SafeReadBits(&s->br, 1, &bits, s->input)) { s->substate_huffman = HS_SIMPLE_BUILD; return E_NEEDS_MORE_INPUT; }
          s->symbol += bits;
        }
        table_size = oracle_build_simple_huffman_table(table, HUFFMAN_TABLE_BITS, s->symbols_lists_array,
                                                       sizeof(s->symbols_lists_array) / sizeof(uint16_t), s->symbol);
        if (opt_table_size) *opt_table_size = table_size;
        s->substate_huffman = HS_NONE;
        return E_SUCCESS;
      case HS_COMPLEX:
        r = ReadCodeLengthCodeLengths(s);
        if (r != E_SUCCESS) return r;
        oracle_build_code_lengths_huffman_table(s->table, s->code_length_code_lengths, s->code_length_histo);
        memset(s->code_length_histo, 0, sizeof(s->code_length_histo));
        for (i = 0; i <= HUFFMAN_MAX_CODE_LENGTH; i++) {
          s->next_symbol[i] = i - (HUFFMAN_MAX_CODE_LENGTH + 1);
          s->symbols_lists_array[SYMBOL_LISTS_INDEX + i - (HUFFMAN_MAX_CODE_LENGTH + 1)] = 0xFFFF;
        }
        s->symbol = 0; s->prev_code_len = kDefaultCodeLength; s->repeat = 0; s->repeat_code_len = 0; s->space = 32768;
        s->substate_huffman = HS_LENGTH_SYMBOLS;
        /* fall through */
      default: /* HS_LENGTH_SYMBOLS */
        r = ReadSymbolCodeLengths(max_symbol, s);
        if (r == E_NEEDS_MORE_INPUT) r = SafeReadSymbolCodeLengths(max_symbol, s);
        if (r != E_SUCCESS) return r;
        if (s->space != 0) return E_HUFFMAN_SPACE;
        table_size = oracle_build_huffman_table(table, HUFFMAN_TABLE_BITS, s->symbols_lists_array, SYMBOL_LISTS_INDEX, s->code_length_histo);
        if (opt_table_size) *opt_table_size = table_size;
        s->substate_huffman = HS_NONE;
        return E_SUCCESS;
    }
  }
}
/* ReadBlockLength :1016-1026 */
static inline uint32_t ReadBlockLength(const HC* table, BR* br, const uint8_t* in) {
  uint32_t code = ReadSymbol(table, br, in);
  return kBrotliBlockLengthOffset[code] + ReadBits(br, kBrotliBlockLengthNBits[code], in);
}
/* SafeReadBlockLength{Index,FromIndex} :1031-1070: the index survives a failed read of the extra bits
 * (BROTLI_STATE_READ_BLOCK_LENGTH_SUFFIX) */
static int SafeReadBlockLength(State* s, uint32_t* result, const HC* table, BR* br, const uint8_t* in) {
  uint32_t index, bits;
  if (s->substate_read_block_length == 0) { if (!SafeReadSymbol(table, br, &index, in)) return 0; }
  else index = s->block_length_index;
  if (!SafeReadBits(br, kBrotliBlockLengthNBits[index], &bits, in)) { s->block_length_index = index; s->substate_read_block_length = 1; return 0; }
  *result = kBrotliBlockLengthOffset[index] + bits;
  s->substate_read_block_length = 0;
  return 1;
}
/* InverseMoveToFrontTransform :1096-1128 */
static void InverseMoveToFrontTransform(uint8_t* v, uint32_t v_len, uint8_t* mtf, uint32_t* mtf_upper_bound) {
  uint32_t upper_bound = *mtf_upper_bound, i;
  for (i = 0; i <= upper_bound; i++) mtf[i] = (uint8_t)i;
  upper_bound = 0;
  for (i = 0; i < v_len; i++) {
    int index = v[i];
    uint8_t value = mtf[index];
    upper_bound |= v[i];
    v[i] = value;
    for (; index > 0; index--) mtf[index] = mtf[index - 1];
    mtf[0] = value;
  }
  *mtf_upper_bound = upper_bound;
}
/* HuffmanTreeGroupDecode :1130-1219 (resumes at htree_index / htree_next_offset) */
static int HuffmanTreeGroupDecode(HGroup* g, State* s) {
  if (s->substate_tree_group == 0) { s->htree_next_offset = 0; s->htree_index = 0; s->substate_tree_group = 1; }
  while (s->htree_index < g->num_htrees) {
    uint32_t table_size = 0;
    int r = ReadHuffmanCode(g->alphabet_size, g->max_symbol, g->codes + s->htree_next_offset, &table_size, s);
    if (r != E_SUCCESS) return r;
    g->htrees[s->htree_index] = s->htree_next_offset;
    s->htree_next_offset += table_size;
    s->htree_index++;
  }
  s->substate_tree_group = 0;
  return E_SUCCESS;
}
/* DecodeContextMap(Inner) :1272-1465 (sub-states BROTLI_STATE_CONTEXT_MAP_*) */
enum { CM_NONE, CM_READ_PREFIX, CM_HUFFMAN, CM_DECODE, CM_TRANSFORM };
static int DecodeContextMap(uint32_t context_map_size, uint32_t* num_htrees, uint8_t** context_map_arg, State* s) {
  uint32_t bits, alphabet_size; int r;
  for (;;) {
    switch (s->substate_context_map) {
      case CM_NONE:
        r = DecodeVarLenUint8(s, num_htrees);
        if (r != E_SUCCESS) return r;
        (*num_htrees)++;
        s->context_index = 0;
        free(*context_map_arg);
        *context_map_arg = (uint8_t*)calloc(context_map_size ? context_map_size : 1, 1);
        if (!*context_map_arg) return E_ALLOC_CONTEXT_MAP;
        if (*num_htrees <= 1) return E_SUCCESS;
        s->substate_context_map = CM_READ_PREFIX;
        /* fall through */
      case CM_READ_PREFIX:
        if (!SafeGetBits(&s->br, 5, &bits, s->input)) return E_NEEDS_MORE_INPUT;
        if (bits & 1) { s->max_run_length_prefix = (bits >> 1) + 1; DropBits(&s->br, 5); }
        else { s->max_run_length_prefix = 0; DropBits(&s->br, 1); }
        s->substate_context_map = CM_HUFFMAN;
        /* fall through */
      case CM_HUFFMAN:
        alphabet_size = *num_htrees + s->max_run_length_prefix;
        r = ReadHuffmanCode(alphabet_size, alphabet_size, s->context_map_table, NULL, s);
        if (r != E_SUCCESS) return r;
        s->code = 0xFFFF;
        s->substate_context_map = CM_DECODE;
        /* fall through */
      case CM_DECODE: {
        uint32_t context_index = s->context_index, max_run_length_prefix = s->max_run_length_prefix, code = s->code;
        uint8_t* context_map = *context_map_arg;
        int rle_code_goto = code != 0xFFFF; /* an RLE code whose extra bits were cut off by the end of the input */
        while (rle_code_goto || context_index < context_map_size) {
          uint32_t reps;
          if (!rle_code_goto) {
            if (!SafeReadSymbol(s->context_map_table, &s->br, &code, s->input)) { s->code = 0xFFFF; s->context_index = context_index; return E_NEEDS_MORE_INPUT; }
            if (code == 0) { context_map[context_index++] = 0; continue; }
            if (code > max_run_length_prefix) { context_map[context_index++] = (uint8_t)(code - max_run_length_prefix); continue; }
          }
          rle_code_goto = 0;
          if (!SafeReadBits(&s->br, code, &reps, s->input)) { s->code = code; s->context_index = context_index; return E_NEEDS_MORE_INPUT; }
          reps += 1u << code;
          if (context_index + reps > context_map_size) return E_CONTEXT_MAP_REPEAT;
          do { context_map[context_index++] = 0; } while (--reps);
        }
        s->substate_context_map = CM_TRANSFORM;
      }
        /* fall through */
      default: /* CM_TRANSFORM */
        if (!SafeReadBits(&s->br, 1, &bits, s->input)) { s->substate_context_map = CM_TRANSFORM; return E_NEEDS_MORE_INPUT; }
        if (bits) InverseMoveToFrontTransform(*context_map_arg, context_map_size, s->mtf, &s->mtf_upper_bound);
        s->substate_context_map = CM_NONE;
        return E_SUCCESS;
    }
  }
}
/* DecodeBlockTypeAndLength :1469-1524 */
static int DecodeBlockTypeAndLength(int safe, State* s, int tree_type) {
  uint32_t max_block_type = s->num_block_types[tree_type], block_type = 0;
  const HC* type_tree = &s->block_type_trees[tree_type * HUFFMAN_MAX_TABLE_SIZE];
  const HC* len_tree = &s->block_len_trees[tree_type * HUFFMAN_MAX_TABLE_SIZE];
  uint32_t* rb = &s->block_type_rb[tree_type * 2];
  if (max_block_type <= 1) return 0;
  if (!safe) {
    block_type = ReadSymbol(type_tree, &s->br, s->input);
    s->block_length[tree_type] = ReadBlockLength(len_tree, &s->br, s->input);
  } else {
    BRState memento = BRSave(&s->br);
    uint32_t block_length_out = 0;
    if (!SafeReadSymbol(type_tree, &s->br, &block_type, s->input)) return 0;
    if (!SafeReadBlockLength(s, &block_length_out, len_tree, &s->br, s->input)) { s->substate_read_block_length = 0; BRRestore(&s->br, &memento); return 0; }
    s->block_length[tree_type] = block_length_out;
  }
  if (block_type == 1) block_type = rb[1] + 1;
  else if (block_type == 0) block_type = rb[0];
  else block_type -= 2;
  if (block_type >= max_block_type) block_type -= max_block_type;
  rb[0] = rb[1]; rb[1] = block_type;
  return 1;
}
/* DetectTrivialLiteralBlockTypes :1525-1553 */
static void DetectTrivialLiteralBlockTypes(State* s) {
  uint32_t i, j;
  memset(s->trivial_literal_contexts, 0, sizeof(s->trivial_literal_contexts));
  for (i = 0; i < s->num_block_types[0]; i++) {
    size_t offset = (size_t)i << kLiteralContextBits; uint32_t error = 0, sample = s->context_map[offset];
    for (j = 0; j < (1u << kLiteralContextBits); j++) error |= s->context_map[offset + j] ^ sample;
    if (error == 0) s->trivial_literal_contexts[i >> 5] |= 1u << (i & 31);
  }
}
/* PrepareLiteralDecoding :1554-1570 */
static void PrepareLiteralDecoding(State* s) {
  uint32_t block_type = s->block_type_rb[1];
  s->context_map_slice_index = (size_t)block_type << kLiteralContextBits;
  s->trivial_literal_context = (int)((s->trivial_literal_contexts[block_type >> 5] >> (block_type & 31)) & 1);
  s->literal_htree_index = s->context_map[s->context_map_slice_index];
  s->context_lookup = &kBrotliContextLookup[(s->context_modes[block_type] & 3) * 512];
}
static int DecodeLiteralBlockSwitch(int safe, State* s) { /* :1575-1588 */
  if (!DecodeBlockTypeAndLength(safe, s, 0)) return 0;
  PrepareLiteralDecoding(s); return 1;
}
static int DecodeCommandBlockSwitch(int safe, State* s) { /* :1609-1621 */
  if (!DecodeBlockTypeAndLength(safe, s, 1)) return 0;
  s->htree_command_index = (uint16_t)s->block_type_rb[3]; return 1;
}
static int DecodeDistanceBlockSwitch(int safe, State* s) { /* :1643-1658 */
  if (!DecodeBlockTypeAndLength(safe, s, 2)) return 0;
  s->dist_context_map_slice_index = (size_t)s->block_type_rb[5] << kDistanceContextBits;
  s->dist_htree_index = s->dist_context_map[s->dist_context_map_slice_index + (size_t)s->distance_context];
  return 1;
}
/* UnwrittenBytes :1679-1692 */
static size_t UnwrittenBytes(const State* s, int wrap) {
  size_t pos = (wrap && s->pos > s->ringbuffer_size) ? (size_t)s->ringbuffer_size : (size_t)s->pos;
  return s->rb_roundtrips * (size_t)s->ringbuffer_size + pos - s->partial_pos_out;
}
/* WriteRingBuffer :1693-1738 */
static int WriteRingBuffer(State* s, int force) {
  size_t to_write = UnwrittenBytes(s, 1), num_written = s->available_out, start_index;
  if (num_written > to_write) num_written = to_write;
  if (s->meta_block_remaining_len < 0) return E_BLOCK_LENGTH_1;
  start_index = s->partial_pos_out & (size_t)s->ringbuffer_mask;
  if (num_written) memcpy(s->output + s->output_offset, s->ringbuffer + start_index, num_written);
  s->output_offset += num_written; s->available_out -= num_written;
  s->partial_pos_out += num_written; s->total_out = s->partial_pos_out;
  if (num_written < to_write) {
    if (s->ringbuffer_size == (1 << s->window_bits) || force) return E_NEEDS_MORE_OUTPUT;
    return E_SUCCESS;
  }
  if (s->ringbuffer_size == (1 << s->window_bits) && s->pos >= s->ringbuffer_size) {
    s->pos -= s->ringbuffer_size; s->rb_roundtrips++; s->should_wrap_ringbuffer = s->pos != 0;
  }
  return E_SUCCESS;
}
/* WrapRingBuffer :1740-1752 */
static void WrapRingBuffer(State* s) {
  if (s->should_wrap_ringbuffer) { memcpy(s->ringbuffer, s->ringbuffer + s->ringbuffer_size, (size_t)s->pos); s->should_wrap_ringbuffer = 0; }
}
/* CopyUncompressedBlockToOutput :1754-1806 */
static int CopyUncompressedBlockToOutput(State* s) {
  for (;;) {
    uint32_t remaining = GetRemainingBytes(&s->br);
    uint32_t nbytes = remaining < (uint32_t)s->meta_block_remaining_len ? remaining : (uint32_t)s->meta_block_remaining_len;
    int r;
    if (nbytes > (uint32_t)(s->ringbuffer_size - s->pos)) nbytes = (uint32_t)(s->ringbuffer_size - s->pos);
    CopyBytes(s->ringbuffer + s->pos, &s->br, nbytes, s->input);
    s->pos += (int)nbytes; s->meta_block_remaining_len -= (int)nbytes;
    if (s->pos < (1 << s->window_bits)) return s->meta_block_remaining_len == 0 ? E_SUCCESS : E_NEEDS_MORE_INPUT;
    r = WriteRingBuffer(s, 0);
    if (r != E_SUCCESS) return r;
    if (s->ringbuffer_size == (1 << s->window_bits)) s->max_distance = s->max_backward_distance;
  }
}
/* canny sizing rule of BrotliAllocateRingBuffer, :1843-1850 */
int oracle_ringbuffer_size(int window_bits, int is_last, int canny, int64_t custom_dict_size, int meta_block_remaining_len) {
  int64_t size = (int64_t)1 << window_bits;
  if (is_last && canny)
    while (size >= (custom_dict_size + (int64_t)meta_block_remaining_len + 16) * 2 && size > 32) size >>= 1;
  if (size > ((int64_t)1 << window_bits)) size = (int64_t)1 << window_bits;
  return (int)size;
}
/* BrotliAllocateRingBuffer :1808-1871 */
static int AllocateRingBuffer(State* s) {
  const int kRingBufferWriteAheadSlack = 42;
  int is_last = s->is_last_metablock;
  size_t max_dict_size; const uint8_t* custom_dict = s->custom_dict;
  s->ringbuffer_size = 1 << s->window_bits;
  if (s->is_uncompressed) {
    int next_block_header = PeekByte(&s->br, (uint32_t)s->meta_block_remaining_len, s->input);
    if (next_block_header != -1 && (next_block_header & 3) == 3) is_last = 1;
  }
  max_dict_size = (size_t)s->ringbuffer_size - 16;
  if ((size_t)s->custom_dict_size > max_dict_size) { custom_dict += (size_t)s->custom_dict_size - max_dict_size; s->custom_dict_size = (ptrdiff_t)max_dict_size; }
  s->ringbuffer_size = oracle_ringbuffer_size((int)s->window_bits, is_last, s->canny_ringbuffer_allocation, s->custom_dict_size, s->meta_block_remaining_len);
  s->ringbuffer_mask = s->ringbuffer_size - 1;
  s->ringbuffer_len = (size_t)s->ringbuffer_size + kRingBufferWriteAheadSlack + BROTLI_MAX_DICTIONARY_WORD_LENGTH;
  s->ringbuffer = (uint8_t*)malloc(s->ringbuffer_len);
  if (!s->ringbuffer) return 0;
  s->ringbuffer[s->ringbuffer_size - 1] = 0; s->ringbuffer[s->ringbuffer_size - 2] = 0;
  if (s->custom_dict_size) memcpy(s->ringbuffer + ((size_t)(-s->custom_dict_size) & (size_t)s->ringbuffer_mask), custom_dict, (size_t)s->custom_dict_size);
  return 1;
}
/* ReadContextModes :1991-2015 (resumes at loop_counter) */
static int ReadContextModes(State* s) {
  uint32_t i, bits;
  for (i = (uint32_t)s->loop_counter; i < s->num_block_types[0]; i++) {
    if (!SafeReadBits(&s->br, 2, &bits, s->input)) { s->loop_counter = (int)i; return E_NEEDS_MORE_INPUT; }
    s->context_modes[i] = (uint8_t)bits;
  }
  return E_SUCCESS;
}
/* TakeDistanceFromRingBuffer :2017-2049 */
static void TakeDistanceFromRingBuffer(State* s) {
  if (s->distance_code == 0) {
    s->dist_rb_idx--; s->distance_code = s->dist_rb[s->dist_rb_idx & 3]; s->distance_context = 1;
  } else {
    int distance_code = s->distance_code << 1;
    const uint32_t kDistanceShortCodeIndexOffset = 0xaaafff1bu, kDistanceShortCodeValueOffset = 0xfa5fa500u;
    int v = (s->dist_rb_idx + ((int32_t)kDistanceShortCodeIndexOffset >> distance_code)) & 3;
    s->distance_code = s->dist_rb[v];
    v = (int)(kDistanceShortCodeValueOffset >> distance_code) & 3;
    if ((distance_code & 3) != 0) s->distance_code += v;
    else { s->distance_code -= v; if (s->distance_code <= 0) s->distance_code = 0x7fffffff; }
  }
}
static inline int SafeReadBits0(BR* br, uint32_t n, uint32_t* val, const uint8_t* in) { /* :2051-2062 */
  if (n) return SafeReadBits(br, n, val, in);
  *val = 0; return 1;
}
/* ReadDistanceInternal :2066-2131 */
static int ReadDistanceInternal(int safe, State* s) {
  int distval; BRState memento; const HC* tree = s->distance_hgroup.codes + s->distance_hgroup.htrees[s->dist_htree_index];
  memset(&memento, 0, sizeof(memento));
  if (!safe) {
    s->distance_code = (int)ReadSymbol(tree, &s->br, s->input);
  } else {
    uint32_t code = 0;
    memento = BRSave(&s->br);
    if (!SafeReadSymbol(tree, &s->br, &code, s->input)) return 0;
    s->distance_code = (int)code;
  }
  s->distance_context = 0;
  if ((s->distance_code & ~0xf) == 0) { TakeDistanceFromRingBuffer(s); s->block_length[2]--; return 1; }
  distval = s->distance_code - (int)s->num_direct_distance_codes;
  if (distval >= 0) {
    uint32_t nbits, bits = 0; int postfix; int64_t offset;
    postfix = distval & s->distance_postfix_mask;
    distval >>= s->distance_postfix_bits;
    nbits = ((uint32_t)distval >> 1) + 1;
    if (safe) {
      if (!SafeReadBits0(&s->br, nbits, &bits, s->input)) { s->distance_code = -1; BRRestore(&s->br, &memento); return 0; }
    } else {
      bits = ReadBits(&s->br, nbits, s->input);
    }
    offset = (int64_t)(int32_t)((uint32_t)(2 + (distval & 1)) << nbits) - 4; /* wrapping i32 arithmetic, :2124 */
    s->distance_code = (int)(uint32_t)(((uint64_t)(offset + (int64_t)bits) << s->distance_postfix_bits) + (uint64_t)(int64_t)postfix + (uint64_t)s->num_direct_distance_codes);
  }
  s->distance_code = (int)((uint32_t)s->distance_code - NUM_DISTANCE_SHORT_CODES + 1);
  s->block_length[2]--;
  return 1;
}
/* ReadCommandInternal :2134-2189 */
static int ReadCommandInternal(int safe, State* s, int* insert_length) {
  uint32_t cmd_code = 0, insert_len_extra = 0, copy_length = 0; BRState memento;
  const HC* tree = s->insert_copy_hgroup.codes + s->insert_copy_hgroup.htrees[s->htree_command_index];
  const BrotliCmdLutElement* v;
  memset(&memento, 0, sizeof(memento));
  if (!safe) cmd_code = ReadSymbol(tree, &s->br, s->input);
  else { memento = BRSave(&s->br); if (!SafeReadSymbol(tree, &s->br, &cmd_code, s->input)) return 0; }
  v = &kBrotliCmdLut[cmd_code];
  s->distance_code = v->distance_code; s->distance_context = v->context;
  s->dist_htree_index = s->dist_context_map[s->dist_context_map_slice_index + (size_t)s->distance_context];
  *insert_length = v->insert_len_offset;
  if (!safe) {
    if (v->insert_len_extra_bits) insert_len_extra = ReadBits(&s->br, v->insert_len_extra_bits, s->input);
    copy_length = ReadBits(&s->br, v->copy_len_extra_bits, s->input);
  } else if (!SafeReadBits0(&s->br, v->insert_len_extra_bits, &insert_len_extra, s->input) ||
             !SafeReadBits0(&s->br, v->copy_len_extra_bits, &copy_length, s->input)) {
    BRRestore(&s->br, &memento); return 0;
  }
  s->copy_length = (int)copy_length + v->copy_len_offset;
  s->block_length[1]--;
  *insert_length += (int)insert_len_extra;
  return 1;
}
/* memmove16 :2203-2228 */
static inline void memmove16(uint8_t* data, uint32_t dst, uint32_t src) { uint8_t t[16]; memcpy(t, data + src, 16); memcpy(data + dst, t, 16); }

/* ProcessCommandsInternal :2330-2744.  `safe`=0 is ProcessCommands (:2746), 1 SafeProcessCommands (:2755). */
static int ProcessCommandsInternal(int safe, State* s) {
  int pos, i, result = E_SUCCESS;
  const uint8_t* in = s->input;
  if ((!safe && !CheckInputAmount(&s->br, 28)) || (!safe && !WarmupBitReader(&s->br, in))) return E_NEEDS_MORE_INPUT;
  pos = s->pos; i = s->loop_counter;
  for (;;) {
    if (s->state == ST_COMMAND_BEGIN) {
      if (!safe && !CheckInputAmount(&s->br, 28)) { result = E_NEEDS_MORE_INPUT; break; }
      if (s->block_length[1] == 0) {
        if (!DecodeCommandBlockSwitch(safe, s)) { result = E_NEEDS_MORE_INPUT; break; } /* :2370 (also when NBLTYPESI==1) */
        continue;
      }
      if (!ReadCommandInternal(safe, s, &i) && safe) { result = E_NEEDS_MORE_INPUT; break; }
      if (i == 0) { s->state = ST_COMMAND_POST_DECODE_LITERALS; continue; }
      s->meta_block_remaining_len -= i;
      s->state = ST_COMMAND_INNER;
    }
    if (s->state == ST_COMMAND_INNER) {
      int inner_return = 0, inner_continue = 0;
      if (s->trivial_literal_context) { /* :2393-2462 (preloaded-symbol variant folded into ReadSymbol) */
        const HC* literal_htree = s->literal_hgroup.codes + s->literal_hgroup.htrees[s->literal_htree_index];
        for (;;) {
          uint32_t literal = 0;
          if (!safe && !CheckInputAmount(&s->br, 28)) { result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
          if (s->block_length[0] == 0) {
            if (!DecodeLiteralBlockSwitch(safe, s) && safe) { result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
            literal_htree = s->literal_hgroup.codes + s->literal_hgroup.htrees[s->literal_htree_index];
            if (!s->trivial_literal_context) { s->state = ST_COMMAND_INNER; inner_continue = 1; break; }
          }
          if (!safe) literal = ReadSymbol(literal_htree, &s->br, in);
          else if (!SafeReadSymbol(literal_htree, &s->br, &literal, in)) { result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
          s->ringbuffer[pos] = (uint8_t)literal;
          if (s->block_length[0] == 0) { result = E_WINDOW_BITS; inner_return = 1; break; } /* :2434-2438 */
          s->block_length[0]--;
          pos++;
          if (pos == s->ringbuffer_size) { s->state = ST_COMMAND_INNER_WRITE; i--; inner_return = 1; break; }
          if (--i == 0) break;
        }
      } else { /* :2463-2551 */
        uint8_t p1 = s->ringbuffer[(pos - 1) & s->ringbuffer_mask], p2 = s->ringbuffer[(pos - 2) & s->ringbuffer_mask];
        if (s->custom_dict_avoid_context_seed && pos < 2) { p1 = 0; p2 = 0; }
        if (pos > 1) s->custom_dict_avoid_context_seed = 0;
        for (;;) {
          uint32_t literal = 0; uint8_t context; const HC* hc;
          if (!safe && !CheckInputAmount(&s->br, 28)) { s->state = ST_COMMAND_INNER; result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
          if (s->block_length[0] == 0) {
            if (!DecodeLiteralBlockSwitch(safe, s) && safe) { result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
            if (s->trivial_literal_context) { s->state = ST_COMMAND_INNER; inner_continue = 1; break; }
          }
          context = s->context_lookup[p1] | s->context_lookup[256 + p2];
          hc = s->literal_hgroup.codes + s->literal_hgroup.htrees[s->context_map[s->context_map_slice_index + context]];
          p2 = p1;
          if (!safe) literal = ReadSymbol(hc, &s->br, in);
          else if (!SafeReadSymbol(hc, &s->br, &literal, in)) { result = E_NEEDS_MORE_INPUT; inner_return = 1; break; }
          p1 = (uint8_t)literal;
          s->ringbuffer[pos] = p1;
          if (s->block_length[0] == 0) { result = E_WINDOW_BITS; inner_return = 1; break; } /* :2522-2526 */
          s->block_length[0]--;
          pos++;
          if (pos == s->ringbuffer_size) { s->state = ST_COMMAND_INNER_WRITE; i--; inner_return = 1; break; }
          if (--i == 0) break;
        }
      }
      if (inner_return) break;
      if (inner_continue) continue;
      if (s->meta_block_remaining_len <= 0) { s->state = ST_METABLOCK_DONE; break; }
      s->state = ST_COMMAND_POST_DECODE_LITERALS;
    }
    if (s->state == ST_COMMAND_POST_DECODE_LITERALS) {
      if (s->distance_code >= 0) { /* implicit distance: reuse last, :2560-2565 */
        s->distance_context = s->distance_code != 0 ? 0 : 1;
        s->dist_rb_idx--;
        s->distance_code = s->dist_rb[s->dist_rb_idx & 3];
      } else {
        if (s->block_length[2] == 0) {
          if (!DecodeDistanceBlockSwitch(safe, s) && safe) { result = E_NEEDS_MORE_INPUT; break; }
        }
        if (!ReadDistanceInternal(safe, s) && safe) { result = E_NEEDS_MORE_INPUT; break; }
      }
      if (s->max_distance != s->max_backward_distance) { /* :2583-2589 */
        if (pos < s->max_backward_distance_minus_custom_dict_size) s->max_distance = pos + (int)s->custom_dict_size;
        else s->max_distance = s->max_backward_distance;
      }
      i = s->copy_length;
      if (s->distance_code > s->max_distance) { /* static dictionary, :2593-2640 */
        if (s->distance_code > kBrotliMaxAllowedDistance) { s->pos = pos; s->loop_counter = i; return E_DISTANCE; }
        if (i >= BROTLI_MIN_DICTIONARY_WORD_LENGTH && i <= BROTLI_MAX_DICTIONARY_WORD_LENGTH) {
          int offset = (int)kBrotliDictOffsetsByLength[i];
          int word_id = s->distance_code - s->max_distance - 1;
          int shift = kBrotliDictSizeBitsByLength[i];
          int mask = (int)BitMask((uint32_t)shift);
          int word_idx = word_id & mask, transform_idx = word_id >> shift;
          s->dist_rb_idx += s->distance_context;
          offset += word_idx * i;
          if (transform_idx < BROTLI_NUM_TRANSFORMS) {
            int len = i; const uint8_t* word = &kBrotliDictionaryData[offset];
            if (transform_idx == 0) memcpy(&s->ringbuffer[pos], word, (size_t)len);
            else len = oracle_transform_dictionary_word(&s->ringbuffer[pos], word, len, transform_idx);
            pos += len; s->meta_block_remaining_len -= len;
            if (pos >= s->ringbuffer_size) { s->state = ST_COMMAND_POST_WRITE_1; break; }
          } else { result = E_TRANSFORM; break; }
        } else { result = E_DICTIONARY; break; }
      } else { /* LZ77 copy, :2641-2680 */
        uint32_t src_start, dst_start, dst_end, src_end;
        s->dist_rb[s->dist_rb_idx & 3] = s->distance_code; s->dist_rb_idx++;
        s->meta_block_remaining_len -= i;
        src_start = (uint32_t)((pos - s->distance_code) & s->ringbuffer_mask);
        dst_start = (uint32_t)pos; dst_end = (uint32_t)pos + (uint32_t)i; src_end = src_start + (uint32_t)i;
        memmove16(s->ringbuffer, dst_start, src_start);
        if (src_end > (uint32_t)pos && dst_end > src_start) { s->state = ST_COMMAND_POST_WRAP_COPY; continue; }
        if (dst_end >= (uint32_t)s->ringbuffer_size || src_end >= (uint32_t)s->ringbuffer_size) { s->state = ST_COMMAND_POST_WRAP_COPY; continue; }
        pos += i;
        if (i > 16) {
          if (i > 32) memmove(s->ringbuffer + dst_start + 16, s->ringbuffer + src_start + 16, (size_t)(i - 16));
          else memmove16(s->ringbuffer, dst_start + 16, src_start + 16);
        }
      }
      if (s->meta_block_remaining_len <= 0) { s->state = ST_METABLOCK_DONE; break; }
      s->state = ST_COMMAND_BEGIN;
      continue;
    }
    if (s->state == ST_COMMAND_POST_WRAP_COPY) { /* :2690-2720 */
      int wrap_guard = s->ringbuffer_size - pos, inner_return = 0;
      while (i > 0) {
        i--;
        s->ringbuffer[pos] = s->ringbuffer[(pos - s->distance_code) & s->ringbuffer_mask];
        pos++;
        if (--wrap_guard == 0) { s->state = ST_COMMAND_POST_WRITE_2; inner_return = 1; break; }
      }
      if (inner_return) break;
      i--;
      if (s->meta_block_remaining_len <= 0) { s->state = ST_METABLOCK_DONE; break; }
      s->state = ST_COMMAND_BEGIN;
      continue;
    }
    result = E_UNREACHABLE; break;
  }
  s->pos = pos; s->loop_counter = i;
  return result;
}
/* BrotliMaxDistanceSymbol :2766-2777 */
static uint32_t MaxDistanceSymbol(uint32_t ndirect, uint32_t npostfix) {
  static const uint32_t bound[4] = {0, 4, 12, 28}, diff[4] = {73, 126, 228, 424};
  uint32_t postfix = 1u << npostfix;
  if (ndirect < bound[npostfix]) return ndirect + diff[npostfix] + postfix;
  if (ndirect > bound[npostfix] + postfix) return ndirect + diff[npostfix];
  return bound[npostfix] + diff[npostfix] + postfix;
}

/* BrotliBitReaderUnload, src/bit_reader/mod.rs:295-306 */
static void BitReaderUnload(BR* br) {
  uint32_t unused_bytes = GetAvailableBits(br) >> 3, unused_bits = unused_bytes << 3;
  br->avail_in += unused_bytes; br->next_in -= unused_bytes;
  if (unused_bits == 64) br->val_ = 0; else br->val_ <<= unused_bits;
  br->bit_pos_ += unused_bits;
}

/* BrotliDecompressStream :2779-3403.  Resumable: everything the decoder needs to continue lives in State, including
 * the 8-byte carry-over of an interrupted read (s->buffer, :2813-2833,:2848-2916).  *input_offset / *available_in and
 * *output_offset / *available_out / *total_out move exactly as the reference moves them. */
static int DecompressStream(State* s, size_t* available_in, size_t* input_offset, const uint8_t* xinput,
                            size_t* available_out, size_t* output_offset, uint8_t* output, size_t* total_out) {
  int result = E_SUCCESS;
  if (s->error_code < 0) return s->error_code; /* is_fatal: sticky, :2796-2798 (error_code is not touched) */
  if ((uint64_t)*available_in >= ((uint64_t)1 << 32) || (uint64_t)*input_offset >= ((uint64_t)1 << 32)) { /* :2799-2804 */
    s->error_code = E_INVALID_ARGUMENTS; return E_INVALID_ARGUMENTS;
  }
  s->output = output; s->available_out = *available_out; s->output_offset = *output_offset; s->total_out = *total_out;
  if (s->buffer_length == 0) {
    s->input = xinput; s->br.avail_in = (uint32_t)*available_in; s->br.next_in = (uint32_t)*input_offset;
  } else { /* :2818-2832: bytes are added to the carry-over one at a time below; they are copied up front */
    size_t copy_len = sizeof(s->buffer) - s->buffer_length;
    result = E_NEEDS_MORE_INPUT;
    if (copy_len > *available_in) copy_len = *available_in;
    if (copy_len) memcpy(s->buffer + s->buffer_length, xinput + *input_offset, copy_len);
    s->input = s->buffer; s->br.next_in = 0;
  }
#define STREAM_SYNC_OUT() do { *available_out = s->available_out; *output_offset = s->output_offset; *total_out = s->total_out; } while (0)
  for (;;) {
    if (result != E_SUCCESS) {
      if (result == E_NEEDS_MORE_INPUT) {
        if (s->ringbuffer) { /* :2838-2850 flush what was decoded; only a fatal outcome replaces the result */
          int r = WriteRingBuffer(s, 1);
          if (r < 0) { result = r; break; }
        }
        if (s->buffer_length != 0) { /* reading from the carry-over, :2851-2886 */
          if (s->br.avail_in == 0) { /* the interrupted read is complete: back to the caller's input */
            s->buffer_length = 0;
            result = E_SUCCESS;
            s->input = xinput; s->br.avail_in = (uint32_t)*available_in; s->br.next_in = (uint32_t)*input_offset;
            continue;
          } else if (*available_in != 0) { /* one more byte from the caller and retry */
            result = E_SUCCESS;
            s->buffer[s->buffer_length] = xinput[*input_offset];
            s->buffer_length++;
            s->br.avail_in = s->buffer_length;
            (*input_offset)++; (*available_in)--;
            continue;
          }
          break; /* cannot finish the read and there is no more input */
        } else { /* :2887-2899 the caller's input ran out: its unread tail goes to the carry-over */
          *input_offset = s->br.next_in; *available_in = s->br.avail_in;
          while (*available_in != 0) {
            s->buffer[s->buffer_length] = xinput[*input_offset];
            s->buffer_length++; (*input_offset)++; (*available_in)--;
          }
          break;
        }
      } else { /* failure or NeedsMoreOutput, :2902-2915 */
        if (s->buffer_length != 0) {
          s->buffer_length = 0; /* the carry-over was consumed and produced output */
        } else {
          BitReaderUnload(&s->br);
          *available_in = s->br.avail_in; *input_offset = s->br.next_in;
        }
      }
      break;
    }
    switch (s->state) {
      case ST_UNINITED: /* :2921-2939 */
        if (!WarmupBitReader(&s->br, s->input)) { result = E_NEEDS_MORE_INPUT; break; }
        result = DecodeWindowBits(&s->large_window, &s->window_bits, &s->br);
        if (result != E_SUCCESS) break;
        s->state = s->large_window ? ST_LARGE_WINDOW_BITS : ST_INITIALIZE;
        break;
      case ST_LARGE_WINDOW_BITS: /* :2940-2951 */
        if (!SafeReadBits(&s->br, 6, &s->window_bits, s->input)) { result = E_NEEDS_MORE_INPUT; break; }
        if (s->window_bits < kBrotliLargeMinWbits || s->window_bits > kBrotliLargeMaxWbits) { result = E_WINDOW_BITS; break; }
        s->state = ST_INITIALIZE;
        break;
      case ST_INITIALIZE: /* :2952-2973 */
        s->max_backward_distance = (1 << s->window_bits) - kBrotliWindowGap;
        s->max_backward_distance_minus_custom_dict_size = (int)((ptrdiff_t)s->max_backward_distance - s->custom_dict_size);
        s->block_type_trees = (HC*)calloc(3 * HUFFMAN_MAX_TABLE_SIZE, sizeof(HC));
        s->block_len_trees = (HC*)calloc(3 * HUFFMAN_MAX_TABLE_SIZE, sizeof(HC));
        if (!s->block_type_trees || !s->block_len_trees) { result = E_ALLOC_BLOCK_TYPE_TREES; break; }
        s->state = ST_METABLOCK_BEGIN;
        break;
      case ST_METABLOCK_BEGIN: /* :2974-2979 */
        StateMetablockBegin(s);
        s->state = ST_METABLOCK_HEADER;
        break;
      case ST_METABLOCK_HEADER: /* :2980-3014 */
        result = DecodeMetaBlockLength(s);
        if (result != E_SUCCESS) break;
        if ((s->is_metadata || s->is_uncompressed) && !JumpToByteBoundary(&s->br)) { result = E_PADDING_2; break; }
        if (s->is_metadata) { s->state = ST_METADATA; break; }
        if (s->meta_block_remaining_len == 0) { s->state = ST_METABLOCK_DONE; break; }
        if (!s->ringbuffer && !AllocateRingBuffer(s)) { result = E_ALLOC_RING_BUFFER_2; break; }
        if (s->is_uncompressed) { s->state = ST_UNCOMPRESSED; break; }
        s->loop_counter = 0;
        s->state = ST_HUFFMAN_CODE_0;
        break;
      case ST_UNCOMPRESSED: /* :3015-3030 */
        result = CopyUncompressedBlockToOutput(s);
        if (result != E_SUCCESS) break;
        s->state = ST_METABLOCK_DONE;
        break;
      case ST_METADATA: /* :3031-3045 */
        while (s->meta_block_remaining_len > 0) {
          uint32_t bits;
          if (!SafeReadBits(&s->br, 8, &bits, s->input)) { result = E_NEEDS_MORE_INPUT; break; }
          s->meta_block_remaining_len--;
        }
        if (result == E_SUCCESS) s->state = ST_METABLOCK_DONE;
        break;
      case ST_HUFFMAN_CODE_0: { /* :3046-3070 */
        int k = s->loop_counter;
        if (k >= 3) { s->state = ST_METABLOCK_HEADER_2; break; }
        result = DecodeVarLenUint8(s, &s->num_block_types[k]);
        if (result != E_SUCCESS) break;
        s->num_block_types[k]++;
        if (s->num_block_types[k] < 2) { s->loop_counter++; break; }
        s->state = ST_HUFFMAN_CODE_1;
        break;
      }
      case ST_HUFFMAN_CODE_1: { /* :3071-3092 */
        uint32_t alphabet_size = s->num_block_types[s->loop_counter] + 2;
        result = ReadHuffmanCode(alphabet_size, alphabet_size, &s->block_type_trees[s->loop_counter * HUFFMAN_MAX_TABLE_SIZE], NULL, s);
        if (result != E_SUCCESS) break;
        s->state = ST_HUFFMAN_CODE_2;
        break;
      }
      case ST_HUFFMAN_CODE_2: /* :3093-3111 */
        result = ReadHuffmanCode(kNumBlockLengthCodes, kNumBlockLengthCodes, &s->block_len_trees[s->loop_counter * HUFFMAN_MAX_TABLE_SIZE], NULL, s);
        if (result != E_SUCCESS) break;
        s->state = ST_HUFFMAN_CODE_3;
        break;
      case ST_HUFFMAN_CODE_3: { /* :3112-3139 */
        uint32_t block_length_out = 0;
        if (!SafeReadBlockLength(s, &block_length_out, &s->block_len_trees[s->loop_counter * HUFFMAN_MAX_TABLE_SIZE], &s->br, s->input)) { result = E_NEEDS_MORE_INPUT; break; }
        s->block_length[s->loop_counter] = block_length_out;
        s->loop_counter++;
        s->state = ST_HUFFMAN_CODE_0;
        break;
      }
      case ST_METABLOCK_HEADER_2: { /* :3141-3163 */
        uint32_t bits;
        if (!SafeReadBits(&s->br, 6, &bits, s->input)) { result = E_NEEDS_MORE_INPUT; break; }
        s->distance_postfix_bits = bits & 3; bits >>= 2;
        s->num_direct_distance_codes = NUM_DISTANCE_SHORT_CODES + (bits << s->distance_postfix_bits);
        s->distance_postfix_mask = (int)BitMask(s->distance_postfix_bits);
        free(s->context_modes);
        s->context_modes = (uint8_t*)calloc(s->num_block_types[0], 1);
        if (!s->context_modes) { result = E_ALLOC_CONTEXT_MODES; break; }
        s->loop_counter = 0;
        s->state = ST_CONTEXT_MODES;
        break;
      }
      case ST_CONTEXT_MODES: /* :3164-3172 */
        result = ReadContextModes(s);
        if (result != E_SUCCESS) break;
        s->state = ST_CONTEXT_MAP_1;
        break;
      case ST_CONTEXT_MAP_1: /* :3173-3187 */
        result = DecodeContextMap(s->num_block_types[0] << kLiteralContextBits, &s->num_literal_htrees, &s->context_map, s);
        if (result != E_SUCCESS) break;
        DetectTrivialLiteralBlockTypes(s);
        s->state = ST_CONTEXT_MAP_2;
        break;
      case ST_CONTEXT_MAP_2: { /* :3188-3266 */
        uint32_t num_direct_codes = s->num_direct_distance_codes - NUM_DISTANCE_SHORT_CODES;
        uint32_t num_distance_codes = NUM_DISTANCE_SHORT_CODES + num_direct_codes +
            ((s->large_window ? BROTLI_LARGE_MAX_DISTANCE_BITS : BROTLI_MAX_DISTANCE_BITS) << (s->distance_postfix_bits + 1));
        uint32_t max_distance_symbol = s->large_window ? MaxDistanceSymbol(num_direct_codes, s->distance_postfix_bits) : num_distance_codes;
        result = DecodeContextMap(s->num_block_types[2] << kDistanceContextBits, &s->num_dist_htrees, &s->dist_context_map, s);
        if (result != E_SUCCESS) break;
        if (!HGroupInit(&s->literal_hgroup, kNumLiteralCodes, kNumLiteralCodes, (uint16_t)s->num_literal_htrees) ||
            !HGroupInit(&s->insert_copy_hgroup, kNumInsertAndCopyCodes, kNumInsertAndCopyCodes, (uint16_t)s->num_block_types[1]) ||
            !HGroupInit(&s->distance_hgroup, (uint16_t)num_distance_codes, (uint16_t)max_distance_symbol, (uint16_t)s->num_dist_htrees)) {
          s->error_code = E_UNREACHABLE; STREAM_SYNC_OUT(); return E_UNREACHABLE;
        }
        s->loop_counter = 0;
        s->state = ST_TREE_GROUP;
        break;
      }
      case ST_TREE_GROUP: /* :3267-3288 */
        result = HuffmanTreeGroupDecode(s->loop_counter == 0 ? &s->literal_hgroup : s->loop_counter == 1 ? &s->insert_copy_hgroup : &s->distance_hgroup, s);
        if (result != E_SUCCESS) break;
        if (++s->loop_counter >= 3) {
          PrepareLiteralDecoding(s);
          s->dist_context_map_slice_index = 0; s->htree_command_index = 0;
          s->state = ST_COMMAND_BEGIN;
        }
        break;
      case ST_COMMAND_BEGIN: case ST_COMMAND_INNER: case ST_COMMAND_POST_DECODE_LITERALS: case ST_COMMAND_POST_WRAP_COPY: /* :3289-3298 */
        result = ProcessCommandsInternal(0, s);
        if (result == E_NEEDS_MORE_INPUT) result = ProcessCommandsInternal(1, s);
        break;
      case ST_COMMAND_INNER_WRITE: case ST_COMMAND_POST_WRITE_1: case ST_COMMAND_POST_WRITE_2: /* :3299-3344 */
        result = WriteRingBuffer(s, 0);
        if (result != E_SUCCESS) break;
        WrapRingBuffer(s);
        if (s->ringbuffer_size == (1 << s->window_bits)) s->max_distance = s->max_backward_distance;
        if (s->state == ST_COMMAND_POST_WRITE_1) {
          s->state = s->meta_block_remaining_len <= 0 ? ST_METABLOCK_DONE : ST_COMMAND_BEGIN;
        } else if (s->state == ST_COMMAND_POST_WRITE_2) {
          s->state = ST_COMMAND_POST_WRAP_COPY;
        } else if (s->loop_counter == 0) {
          s->state = s->meta_block_remaining_len <= 0 ? ST_METABLOCK_DONE : ST_COMMAND_POST_DECODE_LITERALS;
        } else {
          s->state = ST_COMMAND_INNER;
        }
        break;
      case ST_METABLOCK_DONE: /* :3345-3381 */
        if (s->meta_block_remaining_len < 0) { result = E_BLOCK_LENGTH_2; break; }
        StateCleanupAfterMetablock(s);
        if (!s->is_last_metablock) { s->state = ST_METABLOCK_BEGIN; break; }
        if (!JumpToByteBoundary(&s->br)) { result = E_PADDING_2; break; }
        if (s->buffer_length == 0) { /* :3374-3378 hand the unread whole bytes back */
          BitReaderUnload(&s->br);
          *available_in = s->br.avail_in; *input_offset = s->br.next_in;
        }
        s->state = ST_DONE;
        break;
      case ST_DONE: /* :3382-3397 */
        if (s->ringbuffer) { result = WriteRingBuffer(s, 1); if (result != E_SUCCESS) break; }
        s->error_code = result; STREAM_SYNC_OUT();
        return result;
      default:
        s->error_code = E_UNREACHABLE; STREAM_SYNC_OUT();
        return E_UNREACHABLE;
    }
  }
  s->error_code = result; /* SaveErrorCode!, :89-102 */
  STREAM_SYNC_OUT();
  return result;
#undef STREAM_SYNC_OUT
}

/* make_brotli_state!, src/state.rs:279-388 (+ new_with_custom_dictionary :400-411, new_strict :413-420) */
static State* StateCreate(int large_window, const uint8_t* custom_dict, size_t custom_dict_len) {
  State* s = (State*)calloc(1, sizeof(State));
  if (!s) return NULL;
  s->state = ST_UNINITED;
  s->dist_rb[0] = 16; s->dist_rb[1] = 15; s->dist_rb[2] = 11; s->dist_rb[3] = 4;
  s->context_lookup = &kBrotliContextLookup[0];
  s->mtf_upper_bound = 255; s->canny_ringbuffer_allocation = 1; s->large_window = large_window;
  s->custom_dict = custom_dict; s->custom_dict_size = (ptrdiff_t)custom_dict_len; s->custom_dict_avoid_context_seed = custom_dict_len != 0;
  s->context_map_table = (HC*)calloc(HUFFMAN_MAX_TABLE_SIZE, sizeof(HC));
  s->br.val_ = 0; s->br.bit_pos_ = 64; /* BrotliInitBitReader, bit_reader:424-427 */
  if (!s->context_map_table) { free(s); return NULL; }
  return s;
}
static int ResultOfCode(int code) { /* SaveErrorCode!, :89-102 */
  return code == E_SUCCESS ? ORACLE_RESULT_SUCCESS : code == E_NEEDS_MORE_INPUT ? ORACLE_NEEDS_MORE_INPUT
       : code == E_NEEDS_MORE_OUTPUT ? ORACLE_NEEDS_MORE_OUTPUT : ORACLE_RESULT_FAILURE;
}

/* brotli_decode (src/lib.rs:446-468) + BrotliDecoderReturnInfo::new (src/lib.rs:344-369) */
OracleReturnInfo oracle_brotli_decode_ex(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_cap,
                                         int large_window, const uint8_t* custom_dict, size_t custom_dict_len) {
  OracleReturnInfo ret; State* s = StateCreate(large_window, custom_dict, custom_dict_len); int code; const char* msg;
  size_t available_in = input_len, input_offset = 0, available_out = output_cap, output_offset = 0, total_out = 0;
  memset(&ret, 0, sizeof(ret));
  if (!s) { ret.result = ORACLE_RESULT_FAILURE; ret.error_code = E_UNREACHABLE; return ret; }
  code = DecompressStream(s, &available_in, &input_offset, input, &available_out, &output_offset, output, &total_out);
  ret.error_code = code;
  ret.result = ResultOfCode(code);
  ret.decoded_size = output_offset;
  msg = oracle_error_string(code);
  strncpy(ret.error, msg, sizeof(ret.error) - 1);
  StateCleanup(s);
  free(s);
  return ret;
}
OracleReturnInfo oracle_brotli_decode(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_cap) {
  return oracle_brotli_decode_ex(input, input_len, output, output_cap, 1, NULL, 0);
}

/* ---- the streaming call: one State across many BrotliDecompressStream calls (src/decode.rs:2779-2790) ---- */
struct OracleStream { State* s; uint8_t* dict; };
OracleStream* oracle_stream_create(int large_window, const uint8_t* custom_dict, size_t custom_dict_len) {
  OracleStream* o = (OracleStream*)calloc(1, sizeof(OracleStream));
  if (!o) return NULL;
  if (custom_dict_len) { o->dict = (uint8_t*)malloc(custom_dict_len); if (!o->dict) { free(o); return NULL; } memcpy(o->dict, custom_dict, custom_dict_len); }
  o->s = StateCreate(large_window, o->dict, custom_dict_len);
  if (!o->s) { free(o->dict); free(o); return NULL; }
  return o;
}
int oracle_stream_decompress(OracleStream* o, size_t* available_in, size_t* input_offset, const uint8_t* input,
                             size_t* available_out, size_t* output_offset, uint8_t* output, size_t* total_out) {
  return ResultOfCode(DecompressStream(o->s, available_in, input_offset, input, available_out, output_offset, output, total_out));
}
int oracle_stream_error_code(const OracleStream* o) { return o->s->error_code; }
void oracle_stream_destroy(OracleStream* o) {
  if (!o) return;
  StateCleanup(o->s); free(o->s); free(o->dict); free(o);
}

/* ---- threaded batch driver (CPU baseline only) ---- */
typedef struct BatchJob {
  size_t begin, end; const uint8_t* in; const uint64_t* in_off; uint8_t* out; const uint64_t* out_off; uint64_t* out_len; int32_t* codes;
} BatchJob;
static void* BatchWorker(void* arg) {
  BatchJob* j = (BatchJob*)arg; size_t i;
  for (i = j->begin; i < j->end; i++) {
    OracleReturnInfo r = oracle_brotli_decode(j->in + j->in_off[i], (size_t)(j->in_off[i + 1] - j->in_off[i]),
                                              j->out + j->out_off[i], (size_t)(j->out_off[i + 1] - j->out_off[i]));
    if (j->out_len) j->out_len[i] = r.decoded_size;
    if (j->codes) j->codes[i] = r.error_code;
  }
  return NULL;
}
int oracle_brotli_decode_batch(size_t n, const uint8_t* in, const uint64_t* in_off, uint8_t* out, const uint64_t* out_off,
                               uint64_t* out_len, int32_t* codes, int threads) {
  pthread_t* tids; BatchJob* jobs; int t, started = 0; size_t i = 0;
  uint64_t total, acc = 0;
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  tids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  jobs = (BatchJob*)calloc((size_t)threads, sizeof(BatchJob));
  if (!tids || !jobs) { free(tids); free(jobs); return -1; }
  total = n ? (in_off[n] - in_off[0]) + (out_off[n] - out_off[0]) : 0;
  for (t = 0; t < threads; t++) { /* contiguous slices balanced by in+out bytes */
    uint64_t target = total / (uint64_t)threads * (uint64_t)(t + 1);
    BatchJob* j = &jobs[t];
    j->begin = i;
    if (t == threads - 1) i = n;
    else while (i < n && acc < target) { acc += (in_off[i + 1] - in_off[i]) + (out_off[i + 1] - out_off[i]); i++; }
    j->end = i; j->in = in; j->in_off = in_off; j->out = out; j->out_off = out_off; j->out_len = out_len; j->codes = codes;
    if (pthread_create(&tids[t], NULL, BatchWorker, j) != 0) { BatchWorker(j); tids[t] = 0; } else started++;
  }
  for (t = 0; t < threads; t++) if (tids[t]) pthread_join(tids[t], NULL);
  (void)started;
  free(tids); free(jobs);
  return 0;
}
