// brotli_b200_runtime.h -- interface between the host runtime (brotli_b200_host.cpp) and the
// CUDA translation unit (brotli_b200_kernels.cu).  Internal; the public ABI is
// include/brotli_b200/decode.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace brotli_b200 {

// One persistent CTA per SM; each warp owns kWarpSharedBytes of shared memory (scratch + promoted
// prefix-code tables).  24 warps x 9 KiB + the two LUTs fill the 227 KB an sm_100 CTA may use, and
// 768 threads leave 80 registers per thread for the command loop.
#ifndef BROTLI_B200_WARPS_PER_CTA
#define BROTLI_B200_WARPS_PER_CTA 24
#endif
#ifndef BROTLI_B200_WARP_SHARED_BYTES
#define BROTLI_B200_WARP_SHARED_BYTES 9216
#endif
constexpr int kWarpsPerCta = BROTLI_B200_WARPS_PER_CTA;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
constexpr int kMinCtasPerSm = 1;
constexpr uint32_t kWarpSharedBytes = BROTLI_B200_WARP_SHARED_BYTES;
constexpr uint32_t kDynamicSharedBytes = kWarpsPerCta * kWarpSharedBytes;
// the wide geometry of the same kernel (see brotli_b200_kernels.cu): more, slightly slower warps
constexpr int kWarpsPerCtaWide = 28;
constexpr uint32_t kWarpSharedBytesWide = 8000;
constexpr int kMaxWarpsPerCta = kWarpsPerCtaWide > kWarpsPerCta ? kWarpsPerCtaWide : kWarpsPerCta;

// Lane kernel (brotli_b200_lane_kernel.cu): one stream per lane, one persistent CTA per SM; the SM's
// shared memory is split into one private table slot per lane, so fewer warps mean wider root tables.
#ifndef BROTLI_B200_LANE_WARPS_PER_CTA
#define BROTLI_B200_LANE_WARPS_PER_CTA 20
#endif
constexpr int kLaneWarpsPerCta = BROTLI_B200_LANE_WARPS_PER_CTA;

struct ResumeState;  // brotli_b200_session_types.h: device-side record of a streaming session (buffers, results, checkpoint)
struct SessionCopy;  // brotli_b200_session_types.h: one piece of a session launch's staging traffic

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
  const uint32_t* n_ptr;    // optional: the stream count lives in device memory (fallback pass over a bail list)
  const uint8_t* custom_dict;  // optional custom LZ77 dictionary shared by the batch (device memory), src/state.rs:400-411
  uint64_t custom_dict_size;
  ResumeState* sessions;    // optional [n]: streaming sessions -- stream t is wholly described by sessions[t] (device memory)
};

// Extra arguments of the lane kernel.
struct LaneArgs {
  uint8_t* arena;        // resident lanes * lane_arena_bytes_per_lane()
  uint32_t* bail_count;  // device counter, reset by the launcher
  uint32_t* bail_list;   // [n] stream indices the lane kernel gave up; decoded next by the exact kernel
  const uint8_t* xdict;  // expanded static dictionary, xdict_bytes(), filled by launch_build_xdict
  uint32_t slot_bytes;   // shared-memory slot per lane
  uint32_t chunk;        // streams a warp takes per ticket (1..32)
  const uint8_t* cdict;  // custom LZ77 dictionary of the batch or nullptr; 16 readable bytes on either side
  uint64_t cdict_len;
  // geometry chosen on the device (launch_choose_lane_geometry): a launch whose id is not the chosen one exits at once
  const uint32_t* geom_choice = nullptr;
  uint32_t geom_id = 0;
};

size_t arena_bytes_per_warp();
size_t resume_state_bytes();
int query_resident_ctas(int device);
cudaError_t launch_decode_batch(const BatchArgs& a, int ctas, cudaStream_t stream);
cudaError_t launch_session_copy(const SessionCopy* d_pieces, uint32_t n, cudaStream_t stream);
size_t lane_arena_bytes_per_lane();
uint32_t lane_slot_bytes(int warps);
size_t order_temp_bytes(uint32_t n);
cudaError_t launch_order_by_size(uint32_t n, const uint64_t* in_off, uint32_t* scratch, size_t temp_bytes, cudaStream_t stream);
size_t xdict_bytes();
cudaError_t launch_build_xdict(const uint8_t* dictionary, uint8_t* xdict, cudaStream_t stream);
int query_lane_resident_ctas(int device, int warps);
bool lane_kernel_takes_dictionary(int warps);
cudaError_t launch_decode_lane(const BatchArgs& a, const LaneArgs& la, int ctas, int warps, cudaStream_t stream, bool reset_counters = true);
// Geometry by wave fit, decided on the device after the longest-first sort (`sorted_keys`: compressed sizes in 256-byte
// buckets, descending): lanes[0] is the default geometry, lanes[1..n_geom) the alternatives with fewer resident lanes
// (0 = not available).  *choice = index of the geometry to run.  Only uniform batches (smallest stream at least half
// the largest) are fitted: their waves run in lock-step, which is what makes the fit matter.
cudaError_t launch_choose_lane_geometry(uint32_t n, const uint32_t* sorted_keys, const uint32_t* lanes, uint32_t n_geom, uint32_t* choice, cudaStream_t stream);
cudaError_t launch_checksum_batch(uint32_t n, const uint8_t* bytes, const uint64_t* off, const uint64_t* len, uint64_t* sums,
                                  cudaStream_t stream);

}  // namespace brotli_b200
