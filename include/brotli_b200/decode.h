/* include/brotli_b200/decode.h -- C ABI of libbrotli_b200.so, the B200-native batched Brotli decoder.
 *
 * This is the drop-in boundary for the decode path of dropbox/rust-brotli-decompressor
 * (crate brotli-decompressor 5.0.3).  Every entry point below either has the exact name,
 * argument order, enum values and struct layout of the symbol the crate exports through its
 * `ffi-api` feature (reference: c/brotli/decode.h and src/ffi/mod.rs; file:line cited per item),
 * or is a batch extension (prefix BrotliB200) whose per-stream semantics are those of
 * BrotliDecoderDecompress.  All decoding runs in hand-written sm_100a CUDA kernels; there is
 * no CPU decode path in this library: without a usable CUDA device every decode call fails
 * with BROTLI_DECODER_ERROR_UNREACHABLE and the message "brotli_b200: no CUDA device".
 *
 * Plain C: pointers and sizes only.
 */
#ifndef BROTLI_B200_DECODE_H_
#define BROTLI_B200_DECODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define BROTLI_B200_API
#else
#define BROTLI_B200_API __attribute__((visibility("default")))
#endif

/* reference: c/brotli/decode.h:40-49 == src/ffi/interface.rs:16-22 */
typedef enum {
  BROTLI_DECODER_RESULT_ERROR = 0,
  BROTLI_DECODER_RESULT_SUCCESS = 1,
  BROTLI_DECODER_RESULT_NEEDS_MORE_INPUT = 2,
  BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT = 3
} BrotliDecoderResult;

/* reference: c/brotli/decode.h:69-111 == src/state.rs:22-65 */
typedef enum {
  BROTLI_DECODER_NO_ERROR = 0,
  BROTLI_DECODER_SUCCESS = 1,
  BROTLI_DECODER_NEEDS_MORE_INPUT = 2,
  BROTLI_DECODER_NEEDS_MORE_OUTPUT = 3,
  BROTLI_DECODER_ERROR_FORMAT_EXUBERANT_NIBBLE = -1,
  BROTLI_DECODER_ERROR_FORMAT_RESERVED = -2,
  BROTLI_DECODER_ERROR_FORMAT_EXUBERANT_META_NIBBLE = -3,
  BROTLI_DECODER_ERROR_FORMAT_SIMPLE_HUFFMAN_ALPHABET = -4,
  BROTLI_DECODER_ERROR_FORMAT_SIMPLE_HUFFMAN_SAME = -5,
  BROTLI_DECODER_ERROR_FORMAT_CL_SPACE = -6,
  BROTLI_DECODER_ERROR_FORMAT_HUFFMAN_SPACE = -7,
  BROTLI_DECODER_ERROR_FORMAT_CONTEXT_MAP_REPEAT = -8,
  BROTLI_DECODER_ERROR_FORMAT_BLOCK_LENGTH_1 = -9,
  BROTLI_DECODER_ERROR_FORMAT_BLOCK_LENGTH_2 = -10,
  BROTLI_DECODER_ERROR_FORMAT_TRANSFORM = -11,
  BROTLI_DECODER_ERROR_FORMAT_DICTIONARY = -12,
  BROTLI_DECODER_ERROR_FORMAT_WINDOW_BITS = -13,
  BROTLI_DECODER_ERROR_FORMAT_PADDING_1 = -14,
  BROTLI_DECODER_ERROR_FORMAT_PADDING_2 = -15,
  BROTLI_DECODER_ERROR_FORMAT_DISTANCE = -16,
  BROTLI_DECODER_ERROR_DICTIONARY_NOT_SET = -19,
  BROTLI_DECODER_ERROR_INVALID_ARGUMENTS = -20,
  BROTLI_DECODER_ERROR_ALLOC_CONTEXT_MODES = -21,
  BROTLI_DECODER_ERROR_ALLOC_TREE_GROUPS = -22,
  BROTLI_DECODER_ERROR_ALLOC_CONTEXT_MAP = -25,
  BROTLI_DECODER_ERROR_ALLOC_RING_BUFFER_1 = -26,
  BROTLI_DECODER_ERROR_ALLOC_RING_BUFFER_2 = -27,
  BROTLI_DECODER_ERROR_ALLOC_BLOCK_TYPE_TREES = -30,
  BROTLI_DECODER_ERROR_UNREACHABLE = -31
} BrotliDecoderErrorCode;
#define BROTLI_LAST_ERROR_CODE BROTLI_DECODER_ERROR_UNREACHABLE

/* reference: c/brotli/decode.h:143-157 == src/ffi/interface.rs:8-13 */
typedef enum BrotliDecoderParameter {
  BROTLI_DECODER_PARAM_DISABLE_RING_BUFFER_REALLOCATION = 0,
  BROTLI_DECODER_PARAM_LARGE_WINDOW = 1
} BrotliDecoderParameter;

/* reference: c/brotli/decode.h:30-33 == src/huffman/mod.rs:28-33 (only used to type a scratch argument) */
typedef struct HuffmanCodeStruct {
  uint16_t value;
  uint8_t bits;
} HuffmanCode;

/* reference: c/brotli/decode.h:127-132 == src/lib.rs:336-342 (#[repr(C)]) */
typedef struct BrotliDecoderReturnInfoStruct {
  size_t decoded_size;
  char error[256]; /* NUL-terminated BrotliDecoderErrorStr(code), no BROTLI_DECODER_ prefix (src/lib.rs:361-367) */
  BrotliDecoderResult result;
  BrotliDecoderErrorCode code;
} BrotliDecoderReturnInfo;

/* reference: c/brotli/types.h brotli_alloc_func / brotli_free_func == src/ffi/interface.rs:41-46 */
typedef void* (*brotli_alloc_func)(void* opaque, size_t size);
typedef void (*brotli_free_func)(void* opaque, void* address);

typedef struct BrotliDecoderStateStruct BrotliDecoderState;

/* ------------------------------------------------------------------------------------------
 * One-shot entry points (batch of one stream on the GPU).
 * ------------------------------------------------------------------------------------------ */

/* Replaces BrotliDecoderDecompress, c/brotli/decode.h:215-219, src/ffi/mod.rs:262-292.
 * *decoded_size: in = capacity of decoded_buffer, out = bytes written.  SUCCESS only when the
 * whole stream decoded into the buffer; corrupt, truncated and too-small-output all give ERROR
 * (src/ffi/mod.rs:279-283).  decoded_size NULL or misaligned -> ERROR (:269-271).  Trailing bytes
 * after the last metablock are ignored; large-window streams are accepted (src/state.rs:394). */
BROTLI_B200_API BrotliDecoderResult BrotliDecoderDecompress(size_t encoded_size, const uint8_t* encoded_buffer,
                                                            size_t* decoded_size, uint8_t* decoded_buffer);

/* Replaces BrotliDecoderDecompressWithReturnInfo, c/brotli/decode.h:221-225, src/ffi/mod.rs:245-260.
 * result is the BrotliResult of the single BrotliDecompressStream call (src/lib.rs:446-468): 2 when
 * the input is truncated, 3 when decoded_buffer is too small. */
BROTLI_B200_API BrotliDecoderReturnInfo BrotliDecoderDecompressWithReturnInfo(size_t encoded_size,
                                                                              const uint8_t* encoded_buffer,
                                                                              size_t decoded_size,
                                                                              uint8_t* decoded_buffer);

/* Replaces BrotliDecoderDecompressPrealloc, c/brotli/decode.h:227-238, src/ffi/mod.rs:178-223.  The
 * scratch buffers are validated like the reference does (NULL with non-zero length, misaligned or
 * wrapping -> ERROR_INVALID_ARGUMENTS, src/ffi/mod.rs:45-81) and otherwise unused: decoder state
 * lives in device memory owned by the library. */
BROTLI_B200_API BrotliDecoderReturnInfo BrotliDecoderDecompressPrealloc(
    size_t encoded_size, const uint8_t* encoded_buffer, size_t decoded_size, uint8_t* decoded_buffer,
    size_t scratch_u8_size, uint8_t* scratch_u8_buffer, size_t scratch_u32_size, uint32_t* scratch_u32_buffer,
    size_t scratch_hc_size, HuffmanCode* scratch_hc_buffer);

/* ------------------------------------------------------------------------------------------
 * Streaming entry points (GPU backed: each call re-submits the bytes received so far as a batch
 * of one and hands out the newly produced suffix of the output).
 * ------------------------------------------------------------------------------------------ */

/* c/brotli/decode.h:188-189, src/ffi/mod.rs:108-153: both callbacks or neither; large_window off. */
BROTLI_B200_API BrotliDecoderState* BrotliDecoderCreateInstance(brotli_alloc_func alloc_func,
                                                                brotli_free_func free_func, void* opaque);
/* c/brotli/decode.h:167-168, src/ffi/mod.rs:155-176: only before the first byte is consumed. */
BROTLI_B200_API int BrotliDecoderSetParameter(BrotliDecoderState* state, BrotliDecoderParameter param,
                                              uint32_t value);
/* c/brotli/decode.h:196, src/ffi/mod.rs:532-543 */
BROTLI_B200_API void BrotliDecoderDestroyInstance(BrotliDecoderState* state);
/* c/brotli/decode.h:278-280, src/ffi/mod.rs:389-463: NULL in any of the first five -> ERROR with
 * ERROR_INVALID_ARGUMENTS; total_out may be NULL. */
BROTLI_B200_API BrotliDecoderResult BrotliDecoderDecompressStream(BrotliDecoderState* state, size_t* available_in,
                                                                  const uint8_t** next_in, size_t* available_out,
                                                                  uint8_t** next_out, size_t* total_out);
/* src/ffi/mod.rs:466-479: same with plain in/out arrays; consumed/produced counts come back through
 * available_in / available_out. */
BROTLI_B200_API BrotliDecoderResult BrotliDecoderDecompressStreaming(BrotliDecoderState* state, size_t* available_in,
                                                                     const uint8_t* next_in, size_t* available_out,
                                                                     uint8_t* next_out);
BROTLI_B200_API int BrotliDecoderHasMoreOutput(const BrotliDecoderState* state);             /* decode.h:289-290 */
BROTLI_B200_API const uint8_t* BrotliDecoderTakeOutput(BrotliDecoderState* state, size_t* size); /* decode.h:320-321 */
BROTLI_B200_API int BrotliDecoderIsUsed(const BrotliDecoderState* state);                    /* decode.h:333 */
BROTLI_B200_API int BrotliDecoderIsFinished(const BrotliDecoderState* state);                /* decode.h:343-344 */
BROTLI_B200_API BrotliDecoderErrorCode BrotliDecoderGetErrorCode(const BrotliDecoderState* state); /* decode.h:357-358 */
BROTLI_B200_API const char* BrotliDecoderGetErrorString(const BrotliDecoderState* state);   /* decode.h:371-372 */
BROTLI_B200_API const char* BrotliDecoderErrorString(BrotliDecoderErrorCode c);              /* decode.h:377 */
BROTLI_B200_API uint32_t BrotliDecoderVersion(void);                                         /* decode.h:384: 0x1000f00 */
/* src/ffi/mod.rs:492-530 */
BROTLI_B200_API uint8_t* BrotliDecoderMallocU8(BrotliDecoderState* state, size_t size);
BROTLI_B200_API void BrotliDecoderFreeU8(BrotliDecoderState* state, uint8_t* data, size_t size);
BROTLI_B200_API size_t* BrotliDecoderMallocUsize(BrotliDecoderState* state, size_t size);
BROTLI_B200_API void BrotliDecoderFreeUsize(BrotliDecoderState* state, size_t* data, size_t size);

/* ------------------------------------------------------------------------------------------
 * Batch extension (new; the reference decodes one stream per call).  Stream i of a batch is
 * decoded exactly like BrotliDecoderDecompressWithReturnInfo(in_i, out_i): codes[i] is the
 * BrotliDecoderErrorCode (1 = success, 2 = truncated input, 3 = output too small, <0 = corrupt)
 * and out_len[i] the reference's decoded_size for that outcome.
 *
 * Packed layout: stream i's compressed bytes are in_bytes[in_off[i] .. in_off[i+1]) and its
 * output region is out_bytes[out_off[i] .. out_off[i+1]) (the region size is its capacity).
 * Return value: 0 on success, otherwise a negative BrotliDecoderErrorCode for an argument or
 * CUDA failure (message via BrotliB200LastError()).
 * ------------------------------------------------------------------------------------------ */

/* Device-resident batch: every pointer is a device pointer on the current CUDA device;
 * `cuda_stream` is a cudaStream_t (NULL = default stream).  Asynchronous: returns after the
 * launch; results are valid once the stream has been synchronised.  No host<->device copies.
 * The library's scratch buffers (bail list, stream order) grow on the first call of a larger batch: that call may
 * allocate device memory (an implicit synchronisation), so the entry is not meant for CUDA graph capture; launches of
 * different host threads / streams are ordered one after the other (they share the decoders' table arenas).  A stream
 * that is corrupt beyond a too small capacity keeps NeedsMoreOutput(3) here. */
BROTLI_B200_API int BrotliB200DecompressBatchDevice(size_t n, const uint8_t* d_in_bytes, const uint64_t* d_in_off,
                                                    uint8_t* d_out_bytes, const uint64_t* d_out_off,
                                                    uint64_t* d_out_len, int32_t* d_codes, void* cuda_stream);

/* Host-resident packed batch: copies in_bytes/offsets to the device (chunked and overlapped with
 * decoding), decodes, copies the output regions and the per-stream results back.  Synchronous. */
BROTLI_B200_API int BrotliB200DecompressBatchPacked(size_t n, const uint8_t* in_bytes, const uint64_t* in_off,
                                                    uint8_t* out_bytes, const uint64_t* out_off, uint64_t* out_len,
                                                    int32_t* codes);

/* Host-resident scattered batch, the direct generalisation of BrotliDecoderDecompress:
 * out_size[i] is in = capacity / out = bytes written; results[i] is SUCCESS or ERROR with the
 * one-shot's mapping; codes may be NULL. */
BROTLI_B200_API int BrotliB200DecompressBatch(size_t n, const uint8_t* const* in, const size_t* in_size,
                                              uint8_t* const* out, size_t* out_size, BrotliDecoderResult* results,
                                              BrotliDecoderErrorCode* codes);

/* Custom LZ77 dictionary: `dictionary` logically precedes the output of every stream of the call, as in the
 * reference's BrotliState::new_with_custom_dictionary (src/state.rs:400-411; reached through
 * BrotliDecompressCustomDict, src/lib.rs:105-131, Decompressor::new_with_custom_dict, src/reader.rs:105, and the
 * CLI's -dict, src/bin/brotli-decompressor.rs:239-257 -- the reference's C FFI has no entry for it).  Only the last
 * (1 << WBITS) - 16 bytes of the dictionary are reachable (src/decode.rs:1831-1838).  Result fields as for
 * BrotliDecoderDecompressWithReturnInfo. */
BROTLI_B200_API BrotliDecoderReturnInfo BrotliB200DecompressWithDictionary(size_t encoded_size, const uint8_t* encoded_buffer,
                                                                           size_t decoded_size, uint8_t* decoded_buffer,
                                                                           const uint8_t* dictionary, size_t dictionary_size);
/* Streaming form: attaches a custom dictionary to a decoder state before its first input byte -- what
 * Decompressor::new_with_custom_dict (src/reader.rs:105) and DecompressorWriter::new_with_custom_dictionary
 * (src/writer.rs:117) do through BrotliState::new_with_custom_dictionary, which also accepts large-window streams
 * (src/state.rs:400-411).  The bytes are copied.  Returns 1, or 0 if the state has been used already or the arguments
 * are invalid. */
BROTLI_B200_API int BrotliB200DecoderSetCustomDictionary(BrotliDecoderState* state, const uint8_t* dictionary,
                                                         size_t dictionary_size);
/* Multiplexed streaming: the i-th element of every array is one BrotliDecoderDecompressStream call (src/ffi/mod.rs:389-463)
 * on states[i] -- same argument movement, same result, call by call -- and ONE decode launch serves all n states (each
 * state keeps a sliding window of its stream, its prefix-code tables and the decoder's checkpoint in device memory, so a
 * call continues where the previous one stopped, inside a metablock included).  total_out may be NULL.  A state must not
 * appear twice in one call.  Returns 0, or a negative code for an argument or CUDA failure (the affected states fail). */
BROTLI_B200_API int BrotliB200DecoderDecompressStreamBatch(size_t n, BrotliDecoderState* const* states, size_t* available_in,
                                                           const uint8_t** next_in, size_t* available_out, uint8_t** next_out,
                                                           size_t* total_out, BrotliDecoderResult* results);
/* BrotliB200DecompressBatchPacked with one custom dictionary (host memory) shared by all streams of the batch. */
BROTLI_B200_API int BrotliB200DecompressBatchPackedWithDictionary(size_t n, const uint8_t* in_bytes, const uint64_t* in_off,
                                                                  uint8_t* out_bytes, const uint64_t* out_off, uint64_t* out_len,
                                                                  int32_t* codes, const uint8_t* dictionary, size_t dictionary_size);

/* Per-stream 64-bit checksums of device-resident output regions: sums[i] =
 * (sum over j < len[i] of mix((b_j + 1) * (0x9E3779B97F4A7C15 + 2j)) * 0xBF58476D1CE4E5B9) ^ (len[i] * 0x94D049BB133111EB),
 * mix(x) = x ^ (x >> 29), all mod 2^64.  Used to verify bit-exactness of full-size batches without
 * moving them to the host. */
BROTLI_B200_API int BrotliB200ChecksumBatchDevice(size_t n, const uint8_t* d_bytes, const uint64_t* d_off,
                                                  const uint64_t* d_len, uint64_t* d_sums, void* cuda_stream);

/* Number of kernels this library has launched in this process (decode + checksum). */
BROTLI_B200_API uint64_t BrotliB200KernelLaunchCount(void);
/* Device time of the most recent decode kernel launched through BrotliB200DecompressBatchPacked, ms. */
BROTLI_B200_API double BrotliB200LastKernelMs(void);
/* Device time of the decode kernels of the current device's BrotliB200DecompressBatch* launches since the last
 * reset, measured with CUDA events recorded on the launching stream around each kernel (at most the 64 most
 * recent launches are kept).  Synchronises the device.  lane_ms / exact_ms: summed duration of the
 * lane-per-stream kernel and of the exact warp-per-stream kernel; launches: decode calls covered; bailed: streams
 * the lane kernel handed to the exact kernel in the most recent call.  reset != 0 clears the record. */
BROTLI_B200_API int BrotliB200KernelTimes(double* lane_ms, double* exact_ms, uint32_t* launches, uint32_t* bailed, int reset);
/* Tuning knobs of the current device (tests and benchmarks force a decode path with them): "lane_min_streams" -- batches
 * smaller than this skip the lane-per-stream kernel (default 6000, from the measured latency table); "small_geometry" --
 * 0/1, the 8-warp geometry for batches below one wave; "sort_streams" -- 0/1, longest-first order; "lane_slot_bytes" -- a
 * smaller table slot per lane than the geometry allows (0 = all of it).  Returns 1 if set. */
BROTLI_B200_API int BrotliB200SetTuning(const char* name, uint64_t value);
/* Last library-level error message of the calling thread ("" if none). */
BROTLI_B200_API const char* BrotliB200LastError(void);
/* Resident decoding warps per launch on the current device (148 SMs x warps per SM on a B200). */
BROTLI_B200_API int BrotliB200ResidentWarps(void);
/* Warps per SM of the lane-per-stream kernel that decoded the most recent batch on the current device (the geometry is
 * chosen per batch, for uniform batches on the device itself); 0 if that batch did not use it.  Synchronises the device. */
BROTLI_B200_API int BrotliB200LastLaneGeometry(void);
/* Frees the per-device scratch arenas and staging buffers. */
BROTLI_B200_API void BrotliB200Shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* BROTLI_B200_DECODE_H_ */
