/* oracle/brotli_oracle.h -- CPU restatement of dropbox/rust-brotli-decompressor's decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA decoder: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load
 * or call it.  Nothing under rust-brotli-decompressor_b200/ links or includes it.
 *
 * Parity is PINNED: tests/test_oracle_*.py check this code against the reference's own
 * fixtures (the testdata compressed/original pairs), inline vectors (src/test.rs,
 * the src/bin test files, c/main.c), the Huffman-builder table dumps (src/huffman/tests.rs), the
 * bit-reader KATs (src/bit_reader/mod.rs:450-632) and the 256 one-byte streams
 * (src/bin/tests.rs:76-80), and differentially against the system libbrotlidec 1.1.0
 * (the C decoder the Rust crate is a port of).  The Rust crate itself cannot be built in
 * this image (no rustc/cargo).
 */
#ifndef BROTLI_ORACLE_H_
#define BROTLI_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* BrotliResult, src/decode.rs:33-40 */
enum { ORACLE_RESULT_FAILURE = 0, ORACLE_RESULT_SUCCESS = 1, ORACLE_NEEDS_MORE_INPUT = 2, ORACLE_NEEDS_MORE_OUTPUT = 3 };

/* HuffmanCode, src/huffman/mod.rs:28-33 (#[repr(C)]: u16 value, u8 bits, 1 byte padding) */
typedef struct OracleHuffmanCode { uint16_t value; uint8_t bits; } OracleHuffmanCode;

/* BrotliBitReader, src/bit_reader/mod.rs:36-41 */
typedef struct OracleBitReader { uint64_t val_; uint32_t bit_pos_; uint32_t next_in; uint32_t avail_in; } OracleBitReader;

/* Mirror of BrotliDecoderReturnInfo (src/lib.rs:336-370) for the one-shot entry. */
typedef struct OracleReturnInfo {
  size_t decoded_size;
  char error[256];
  int result;      /* BrotliResult */
  int error_code;  /* BrotliDecoderErrorCode, src/state.rs:22-65 */
} OracleReturnInfo;

/* One-shot decode == brotli_decode (src/lib.rs:446-468): BrotliState::new (large_window=true),
 * one BrotliDecompressStream call over the whole input/output. */
OracleReturnInfo oracle_brotli_decode(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_cap);

/* Same with options: large_window (src/state.rs:394 vs :416) and a custom LZ77 dictionary
 * (BrotliState::new_with_custom_dictionary, src/state.rs:400-411). */
OracleReturnInfo oracle_brotli_decode_ex(const uint8_t* input, size_t input_len, uint8_t* output, size_t output_cap,
                                         int large_window, const uint8_t* custom_dict, size_t custom_dict_len);

/* Decode a batch on `threads` host threads (static split balanced by in+out bytes); used as the
 * timed CPU baseline.  Streams are in[in_off[i]..in_off[i+1]) -> out[out_off[i]..out_off[i+1]). */
int oracle_brotli_decode_batch(size_t n, const uint8_t* in, const uint64_t* in_off, uint8_t* out,
                               const uint64_t* out_off, uint64_t* out_len, int32_t* codes, int threads);

/* The reference's resumable call: one BrotliState across many BrotliDecompressStream calls (src/decode.rs:2779-2790;
 * BrotliState persists between calls, src/state.rs:156-278).  The arguments move exactly as the reference moves them:
 * *input_offset / *available_in by the bytes consumed, *output_offset / *available_out by the bytes produced,
 * *total_out = bytes produced since the start of the stream.  Returns a BrotliResult. */
typedef struct OracleStream OracleStream;
OracleStream* oracle_stream_create(int large_window, const uint8_t* custom_dict, size_t custom_dict_len);
int oracle_stream_decompress(OracleStream* o, size_t* available_in, size_t* input_offset, const uint8_t* input,
                             size_t* available_out, size_t* output_offset, uint8_t* output, size_t* total_out);
int oracle_stream_error_code(const OracleStream* o); /* BrotliDecoderErrorCode of the last call */
void oracle_stream_destroy(OracleStream* o);

const char* oracle_error_string(int code); /* BrotliDecoderErrorStr, src/state.rs:533-578 */

/* ---- pieces exported for the reference's known-answer tests ---- */
void oracle_build_code_lengths_huffman_table(OracleHuffmanCode* table, const uint8_t* code_lengths, const uint16_t* count);
uint32_t oracle_build_huffman_table(OracleHuffmanCode* root_table, int root_bits, const uint16_t* symbol_lists,
                                    size_t symbol_lists_offset, uint16_t* count);
uint32_t oracle_build_simple_huffman_table(OracleHuffmanCode* table, int root_bits, const uint16_t* val, size_t val_len,
                                           uint32_t num_symbols);
int oracle_transform_dictionary_word(uint8_t* dst, const uint8_t* word, int len, int transform);
int oracle_br_warmup(OracleBitReader* br, const uint8_t* input);
int oracle_br_safe_read_bits(OracleBitReader* br, uint32_t n_bits, uint32_t* val, const uint8_t* input);
uint32_t oracle_br_read_bits(OracleBitReader* br, uint32_t n_bits, const uint8_t* input);
uint32_t oracle_br_read_constant_n_bits(OracleBitReader* br, uint32_t n_bits, const uint8_t* input);
uint32_t oracle_br_get16_bits_unmasked(OracleBitReader* br, const uint8_t* input);
/* canny ring-buffer sizing, src/decode.rs:1843-1850 (KAT at :1888-1892) */
int oracle_ringbuffer_size(int window_bits, int is_last, int canny, int64_t custom_dict_size, int meta_block_remaining_len);

#ifdef __cplusplus
}
#endif
#endif /* BROTLI_ORACLE_H_ */
