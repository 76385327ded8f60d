// brotli_b200_runtime.h -- interface between the host runtime (brotli_b200_host.cpp) and the
// CUDA translation unit (brotli_b200_kernels.cu).  Internal; the public ABI is
// include/brotli_b200/decode.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace brotli_b200 {

constexpr int kWarpsPerCta = 4;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kMinCtasPerSm = 8;  // 32 decoding warps per SM

// One launch decodes streams [0, n) of a packed batch (see decode.h, "Packed layout").
struct BatchArgs {
  const uint8_t* in;
  const uint64_t* in_off;   // [n + 1]
  uint8_t* out;
  const uint64_t* out_off;  // [n + 1]
  uint64_t* out_len;        // [n]
  int32_t* codes;           // [n]
  uint64_t* in_used;        // optional [n]: compressed bytes consumed
  const uint32_t* order;    // optional [n]: ticket t decodes stream order[t] (longest first)
  uint32_t* ticket;         // device counter, reset by the launcher
  uint8_t* arena;           // resident_warps * arena_bytes_per_warp()
  const uint8_t* dictionary;  // RFC 7932 static dictionary in device memory
  uint32_t n;
  uint32_t large_window;    // accept the large-window header (one-shot: 1, src/state.rs:394)
};

size_t arena_bytes_per_warp();
int query_resident_ctas(int device);
cudaError_t launch_decode_batch(const BatchArgs& a, int ctas, cudaStream_t stream);
cudaError_t launch_checksum_batch(uint32_t n, const uint8_t* bytes, const uint64_t* off, const uint64_t* len, uint64_t* sums,
                                  cudaStream_t stream);

}  // namespace brotli_b200
