// brotli_b200_host.cpp -- host runtime and C ABI of libbrotli_b200.so (include/brotli_b200/decode.h).
//
// The reference's host side is Rust; rustc is not available in the build image, so the host layer
// above the CUDA kernels is C++ and keeps the reference's names and semantics:
//   BrotliDecoderDecompress{,WithReturnInfo,Prealloc}   src/ffi/mod.rs:178-292  (-> brotli_decode, src/lib.rs:446-468)
//   BrotliDecoderDecompressStream & state queries       src/ffi/mod.rs:108-176,389-590
// plus the batch entry points.  There is no CPU decoder in this file or in anything it links:
// every byte is decoded by brotli_decode_batch_kernel; without a CUDA device calls fail loudly.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/brotli_b200/decode.h"
#include "brotli_b200_runtime.h"
// the kernel's session record (ResumeState, SessionCopy) and the session logic
#include "brotli_b200_session_types.h"
#include "brotli_b200_session.h"

extern "C" const uint8_t kBrotliDictionaryData[];  // tables/brotli_dictionary.c (122 784 bytes)

namespace {

using brotli_b200::BatchArgs;

constexpr size_t kDictionaryBytes = 122784;
constexpr int kMaxDevices = 16;
constexpr size_t kPipelineChunkBytes = 256u << 20;  // in+out bytes per pipeline stage of the host-batch path

thread_local std::string tl_error;
std::atomic<uint64_t> g_launches{0};
std::atomic<double> g_last_kernel_ms{0.0};

void set_error(const std::string& s) { tl_error = s; }

const char* error_name(int c) {  // BrotliDecoderErrorStr, src/state.rs:533-578
  switch (c) {
    case 0: return "NO_ERROR";
    case 1: return "SUCCESS";
    case 2: return "NEEDS_MORE_INPUT";
    case 3: return "NEEDS_MORE_OUTPUT";
    case -1: return "ERROR_FORMAT_EXUBERANT_NIBBLE";
    case -2: return "ERROR_FORMAT_RESERVED";
    case -3: return "ERROR_FORMAT_EXUBERANT_META_NIBBLE";
    case -4: return "ERROR_FORMAT_SIMPLE_HUFFMAN_ALPHABET";
    case -5: return "ERROR_FORMAT_SIMPLE_HUFFMAN_SAME";
    case -6: return "ERROR_FORMAT_FL_SPACE";  // sic, src/state.rs:547
    case -7: return "ERROR_FORMAT_HUFFMAN_SPACE";
    case -8: return "ERROR_FORMAT_CONTEXT_MAP_REPEAT";
    case -9: return "ERROR_FORMAT_BLOCK_LENGTH_1";
    case -10: return "ERROR_FORMAT_BLOCK_LENGTH_2";
    case -11: return "ERROR_FORMAT_TRANSFORM";
    case -12: return "ERROR_FORMAT_DICTIONARY";
    case -13: return "ERROR_FORMAT_WINDOW_BITS";
    case -14: return "ERROR_FORMAT_PADDING_1";
    case -15: return "ERROR_FORMAT_PADDING_2";
    case -16: return "ERROR_FORMAT_DISTANCE";
    case -19: return "ERROR_DICTIONARY_NOT_SET";
    case -20: return "ERROR_INVALID_ARGUMENTS";
    case -21: return "ERROR_ALLOC_CONTEXT_MODES";
    case -22: return "ERROR_ALLOC_TREE_GROUPS";
    case -25: return "ERROR_ALLOC_CONTEXT_MAP";
    case -26: return "ERROR_ALLOC_RING_BUFFER_1";
    case -27: return "ERROR_ALLOC_RING_BUFFER_2";
    case -30: return "ERROR_ALLOC_BLOCK_TYPE_TREES";
    default: return "ERROR_UNREACHABLE";
  }
}

// Grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, n); want = n; }  // (clear the failed attempt from the runtime's last-error slot)
    if (e != cudaSuccess) cudaGetLastError();
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Everything the library owns on one GPU.  Created on first use of that device.
struct DeviceCtx {
  std::mutex mu;       // serialises users of the staging buffers / arena / streams
  bool ready = false;
  int device = -1;
  int ctas = 0;
  int lane_ctas = 0;           // 0: lane kernel disabled (BROTLI_B200_LANE=0): every stream takes the exact kernel
  int lane_warps = 0;          // warps per CTA of the lane kernel
  uint32_t lane_slot_bytes = 0;
  // second geometry for batches below one wave: fewer, fuller warps with wider table slots (profiles/r02/lat_*.jsonl)
  int small_warps = 0, small_ctas = 0;
  uint32_t small_slot_bytes = 0;
  // further geometries with fewer resident lanes than the default, taken when a batch falls into waves better with them
  // (profiles/r02/waves.txt): {warps, ctas, slot bytes}; ctas == 0: not available
  struct LaneGeom { int warps = 0, ctas = 0; uint32_t slot_bytes = 0; };
  LaneGeom alt[2];
  bool geom_lanes_uploaded = false;
  int last_lane_warps = 0;       // geometry of the most recent lane launch; -1: chosen on the device (ticket + 24)
  size_t lane_min_streams = 0; // batches smaller than this go straight to the warp-per-stream kernel
  uint8_t* lane_arena = nullptr;
  uint8_t* xdict = nullptr;    // expanded static dictionary of the lane kernel
  uint32_t* bail_count = nullptr;
  DevBuf bail_list;
  DevBuf order;                // longest-first order of a batch (lane kernel), see launch_order_by_size
  bool sort_streams = true;    // BROTLI_B200_SORT=0: decode in batch order
  uint8_t* arena = nullptr;
  uint8_t* dictionary = nullptr;
  uint32_t* ticket = nullptr;
  cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
  // per-launch kernel timing (BrotliB200KernelTimes): [i][0..2] = before lane kernel, between, after exact kernel
  static constexpr int kTimedLaunches = 64;
  cudaEvent_t ev_t[kTimedLaunches][3] = {};
  uint32_t timed_count = 0;
  std::mutex launch_mu;          // orders launches: decode kernels share the scratch arena and the ticket
  cudaEvent_t ev_arena = nullptr;  // completion of the most recent decode launch
  bool arena_busy = false;
  DevBuf in, out, in_off, out_off, out_len, codes, in_used;
  DevBuf cdict;            // custom LZ77 dictionary of the batch in flight
  DevBuf sess_states, sess_pieces, sess_blob;  // staging of a session launch: ResumeState[], SessionCopy[], bytes
  void* sess_pin_up = nullptr; size_t sess_pin_up_cap = 0;      // pinned host staging of the sessions' fresh input ...
  void* sess_pin_down = nullptr; size_t sess_pin_down_cap = 0;  // ... and of their new output
  // session buffers are recycled through power-of-two size classes (a state opens with three device buffers; thousands of
  // states must not mean thousands of cudaMalloc calls per second): free lists per class, small classes cut from slabs
  std::vector<uint8_t*> sess_free[48];
  std::vector<void*> sess_slabs;
  std::unordered_map<uint8_t*, int> sess_class;  // live block -> size class
  DevBuf redo_out;         // output windows of the exact re-decode of NeedsMoreOutput one-shots
  // pinned host staging of the scattered-batch entry (BrotliB200DecompressBatch): grow-only
  void* pin_in = nullptr; size_t pin_in_cap = 0;
  void* pin_out = nullptr; size_t pin_out_cap = 0;
  std::mutex pin_mu;
};

DeviceCtx g_ctx[kMaxDevices];
constexpr int kDefaultL2FetchBytes = 0;  // see acquire_ctx

#define CU_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      set_error(std::string("brotli_b200: CUDA error: ") + cudaGetErrorString(e_) + " at " #expr); \
      return BROTLI_DECODER_ERROR_UNREACHABLE;                                                    \
    }                                                                                             \
  } while (0)

// Returns the context of the current device (initialising it), or nullptr with the error set.
DeviceCtx* acquire_ctx() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  int count = 0;
  if (e == cudaSuccess) e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1 || dev >= kMaxDevices) {
    cudaGetLastError();
    set_error("brotli_b200: no CUDA device (this library has no CPU decode path)");
    return nullptr;
  }
  DeviceCtx* c = &g_ctx[dev];
  std::lock_guard<std::mutex> lock(c->mu);
  if (c->ready) return c;
  c->device = dev;
  c->ctas = brotli_b200::query_resident_ctas(dev);
  if (c->ctas <= 0) { set_error("brotli_b200: occupancy query failed (kernel image not loadable on this device?)"); return nullptr; }
  const size_t arena_bytes = (size_t)c->ctas * brotli_b200::kMaxWarpsPerCta * brotli_b200::arena_bytes_per_warp();
  bool ok = cudaMalloc((void**)&c->arena, arena_bytes) == cudaSuccess &&
            cudaMalloc((void**)&c->dictionary, kDictionaryBytes + 64) == cudaSuccess &&
            cudaMalloc((void**)&c->ticket, 256) == cudaSuccess &&
            cudaMemcpy(c->dictionary, kBrotliDictionaryData, kDictionaryBytes, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->ev_k0) == cudaSuccess && cudaEventCreate(&c->ev_k1) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_arena, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error(std::string("brotli_b200: device initialisation failed: ") + cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  // L2 -> DRAM fetch granularity.  The decoders gather 4..16 bytes per backreference from windows that are, summed
  // over the resident streams, far larger than L2; a smaller fetch unit cuts the over-fetch of every such miss.
  // BROTLI_B200_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity for the device, 0 leaves it alone.
  {
    const char* fetch_env = getenv("BROTLI_B200_L2_FETCH");
    const int fetch = fetch_env ? atoi(fetch_env) : kDefaultL2FetchBytes;
    if (fetch == 32 || fetch == 64 || fetch == 128) {
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch) != cudaSuccess) cudaGetLastError();  // a hint: not fatal
    }
  }
  const char* lane_env = getenv("BROTLI_B200_LANE");
  if (!(lane_env && lane_env[0] == '0')) {
    const char* warps_env = getenv("BROTLI_B200_LANE_WARPS");
    c->lane_warps = warps_env ? atoi(warps_env) : brotli_b200::kLaneWarpsPerCta;
    c->lane_ctas = brotli_b200::query_lane_resident_ctas(dev, c->lane_warps);
    if (c->lane_ctas <= 0) { set_error("brotli_b200: occupancy query of the lane kernel failed"); return nullptr; }
    c->lane_slot_bytes = brotli_b200::lane_slot_bytes(c->lane_warps);
    if (const char* slot_env = getenv("BROTLI_B200_LANE_SLOT_BYTES")) {  // experiments: a smaller table slot than the geometry allows
      const uint32_t v = ((uint32_t)atoi(slot_env) / 4u) | 1u;            // (an odd number of words, see lane_slot_bytes)
      if (v * 4u >= 80u && v * 4u <= c->lane_slot_bytes) c->lane_slot_bytes = v * 4u;
    }
    const size_t lane_bytes = (size_t)c->lane_ctas * c->lane_warps * 32 * brotli_b200::lane_arena_bytes_per_lane();
    if (cudaMalloc((void**)&c->lane_arena, lane_bytes) != cudaSuccess) {
      set_error(std::string("brotli_b200: lane arena allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    if (cudaMalloc((void**)&c->xdict, brotli_b200::xdict_bytes()) != cudaSuccess ||
        brotli_b200::launch_build_xdict(c->dictionary, c->xdict, c->s_compute) != cudaSuccess ||
        cudaStreamSynchronize(c->s_compute) != cudaSuccess) {
      set_error(std::string("brotli_b200: expanded dictionary build failed: ") + cudaGetErrorString(cudaGetLastError()));
      return nullptr;
    }
    g_launches.fetch_add(1);
    // Batch-size dependent routing, from the measured latency table (profiles/r02/lat_*.jsonl): a stream
    // takes ~25-32 ms on a lane however few streams there are, ~5-17 ms on a warp of the exact kernel -- below ~6000
    // streams the exact kernel alone is faster; below one wave of the default geometry 8 warps per SM with full warps
    // beat 14 warps with half-empty ones.
    c->lane_min_streams = getenv("BROTLI_B200_LANE_MIN") ? (size_t)atoll(getenv("BROTLI_B200_LANE_MIN")) : 6000;
    if (!warps_env && !(getenv("BROTLI_B200_LANE_SMALL") && getenv("BROTLI_B200_LANE_SMALL")[0] == '0')) {
      c->small_warps = 8;
      c->small_ctas = brotli_b200::query_lane_resident_ctas(dev, c->small_warps);
      c->small_slot_bytes = brotli_b200::lane_slot_bytes(c->small_warps);
      if (c->small_ctas <= 0) c->small_warps = 0;
      if (!(getenv("BROTLI_B200_LANE_FIT") && getenv("BROTLI_B200_LANE_FIT")[0] == '0') && c->lane_warps == 20) {
        const int alt_warps[2] = {14, 16};
        for (int k = 0; k < 2; k++) {
          c->alt[k].warps = alt_warps[k];
          c->alt[k].ctas = brotli_b200::query_lane_resident_ctas(dev, alt_warps[k]);
          c->alt[k].slot_bytes = brotli_b200::lane_slot_bytes(alt_warps[k]);
          if (c->alt[k].ctas <= 0) c->alt[k].ctas = 0;
        }
      }
    }
    c->bail_count = c->ticket + 16;  // same 256-byte allocation as the tickets
    const char* sort_env = getenv("BROTLI_B200_SORT");
    c->sort_streams = !(sort_env && sort_env[0] == '0');
  }
  c->ready = true;
  return c;
}

// ---- device-resident batch -------------------------------------------------------------------
int decode_device(DeviceCtx* c, size_t n, const uint8_t* d_in, const uint64_t* d_in_off, uint8_t* d_out,
                  const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_codes, uint64_t* d_in_used, uint32_t large_window,
                  cudaStream_t stream, const uint8_t* d_dict = nullptr, uint64_t dict_size = 0, brotli_b200::ResumeState* d_sessions = nullptr) {
  if (n == 0) return 0;
  if (n >= ((uint64_t)1 << 31)) { set_error("brotli_b200: batch too large"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
  BatchArgs a;
  a.in = d_in; a.in_off = d_in_off; a.out = d_out; a.out_off = d_out_off; a.out_len = d_out_len; a.codes = d_codes;
  a.in_used = d_in_used;
  a.order = nullptr; a.ticket = c->ticket; a.arena = c->arena; a.dictionary = c->dictionary;
  a.n = (uint32_t)n; a.large_window = large_window; a.n_ptr = nullptr;
  a.custom_dict = dict_size ? d_dict : nullptr; a.custom_dict_size = d_dict ? dict_size : 0;
  a.sessions = d_sessions;
  std::lock_guard<std::mutex> lock(c->launch_mu);
  if (c->arena_busy) CU_TRY(cudaStreamWaitEvent(stream, c->ev_arena, 0));  // launches on other streams must not overlap
  const uint32_t slot = c->timed_count % DeviceCtx::kTimedLaunches;
  for (int e = 0; e < 3; e++) if (!c->ev_t[slot][e]) CU_TRY(cudaEventCreate(&c->ev_t[slot][e]));
  CU_TRY(cudaEventRecord(c->ev_t[slot][0], stream));
  // (a streaming session is one stream: exact kernel; a batch with a custom dictionary takes the lane kernel's dictionary
  // instance when the configured geometry has one)
  c->last_lane_warps = 0;
  if (c->lane_ctas > 0 && a.sessions == nullptr && n >= c->lane_min_streams &&
      (a.custom_dict_size == 0 || brotli_b200::lane_kernel_takes_dictionary(c->lane_warps))) {
    // optimistic pass: one stream per lane; whatever it gives up lands on the bail list
    CU_TRY(c->bail_list.reserve(n * sizeof(uint32_t)));
    brotli_b200::LaneArgs la;
    int lane_warps = c->lane_warps, lane_ctas = c->lane_ctas;
    la.slot_bytes = c->lane_slot_bytes;
    if (c->small_warps && a.custom_dict_size == 0 && n <= (size_t)c->small_ctas * c->small_warps * 32) {
      lane_warps = c->small_warps; lane_ctas = c->small_ctas; la.slot_bytes = c->small_slot_bytes;
    }
    // geometry by wave fit, decided on the device after the sort (uniform batches only): see launch_choose_lane_geometry
    const bool fit = lane_warps == c->lane_warps && a.custom_dict_size == 0 && c->sort_streams && n >= 64 &&
                     c->lane_slot_bytes == brotli_b200::lane_slot_bytes(c->lane_warps) && (c->alt[0].ctas > 0 || c->alt[1].ctas > 0);
    la.arena = c->lane_arena; la.bail_count = c->bail_count; la.bail_list = (uint32_t*)c->bail_list.p;
    la.xdict = c->xdict;
    la.cdict = a.custom_dict; la.cdict_len = a.custom_dict_size;
    // small batches: fewer streams per warp, spread over all resident warps
    const size_t total_warps = (size_t)lane_ctas * lane_warps;
    const size_t per_warp = (n + total_warps - 1) / total_warps;
    la.chunk = per_warp >= 32 ? 32u : (uint32_t)per_warp;
    if (c->sort_streams && n >= 64) {
      const size_t temp = brotli_b200::order_temp_bytes((uint32_t)n);
      CU_TRY(c->order.reserve(4 * n * sizeof(uint32_t) + temp + 256));
      CU_TRY(brotli_b200::launch_order_by_size((uint32_t)n, d_in_off, (uint32_t*)c->order.p, temp, stream));
      g_launches.fetch_add(2);
      a.order = (const uint32_t*)c->order.p + 3 * n;
    }
    if (fit) {
      // one launch per candidate geometry; all but the chosen one exit at once (the choice is made on the device, so
      // the call stays asynchronous)
      uint32_t* const d_geom = c->ticket + 24;  // [0] choice, [1..3] resident lanes per candidate (same 256-byte allocation as the tickets)
      uint32_t lanes[3] = {(uint32_t)(lane_ctas * lane_warps * 32), c->alt[0].ctas > 0 ? (uint32_t)(c->alt[0].ctas * c->alt[0].warps * 32) : 0u,
                           c->alt[1].ctas > 0 ? (uint32_t)(c->alt[1].ctas * c->alt[1].warps * 32) : 0u};
      if (!c->geom_lanes_uploaded) {
        CU_TRY(cudaMemcpyAsync(d_geom + 1, lanes, sizeof(lanes), cudaMemcpyHostToDevice, stream));
        c->geom_lanes_uploaded = true;
      }
      CU_TRY(brotli_b200::launch_choose_lane_geometry((uint32_t)n, (const uint32_t*)c->order.p + n, d_geom + 1, 3, d_geom, stream));
      CU_TRY(cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), stream));
      CU_TRY(cudaMemsetAsync(la.bail_count, 0, sizeof(uint32_t), stream));
      la.geom_choice = d_geom;
      c->last_lane_warps = -1;
      for (uint32_t g = 0; g < 3; g++) {
        if (lanes[g] == 0) continue;
        brotli_b200::LaneArgs lg = la;
        lg.geom_id = g;
        const int gw = g == 0 ? lane_warps : c->alt[g - 1].warps, gc = g == 0 ? lane_ctas : c->alt[g - 1].ctas;
        if (g != 0) lg.slot_bytes = c->alt[g - 1].slot_bytes;
        const size_t gwarps = (size_t)gc * gw, gper = (n + gwarps - 1) / gwarps;
        lg.chunk = gper >= 32 ? 32u : (uint32_t)gper;
        CU_TRY(brotli_b200::launch_decode_lane(a, lg, gc, gw, stream, false));
        g_launches.fetch_add(1);
      }
      g_launches.fetch_add(1);
    } else {
      CU_TRY(brotli_b200::launch_decode_lane(a, la, lane_ctas, lane_warps, stream));
      g_launches.fetch_add(1);
      c->last_lane_warps = lane_warps;
    }
    // exact pass over the bail list (usually empty): one warp per stream, full reference semantics
    a.order = la.bail_list; a.n_ptr = la.bail_count; a.ticket = c->ticket + 8;
  }
  CU_TRY(cudaEventRecord(c->ev_t[slot][1], stream));
  CU_TRY(brotli_b200::launch_decode_batch(a, c->ctas, stream));
  CU_TRY(cudaEventRecord(c->ev_t[slot][2], stream));
  c->timed_count++;
  CU_TRY(cudaEventRecord(c->ev_arena, stream));
  c->arena_busy = true;
  g_launches.fetch_add(1);
  return 0;
}

// ---- host-resident packed batch: chunked H2D -> decode -> D2H pipeline ------------------------
// Chunks are contiguous stream ranges of about kPipelineChunkBytes (in + out).  Copies run on
// their own streams and are ordered against the decode stream with events, so the H2D of chunk
// k+1 and the D2H of chunk k-1 overlap the decode of chunk k.  Decode kernels stay on one stream
// because they share the per-warp scratch arena.
int redo_needs_more_output(DeviceCtx* c, size_t n, const uint8_t* in_bytes, const uint64_t* in_off, const uint64_t* out_off, const uint8_t* d_in,
                           uint64_t* out_len, int32_t* codes, const uint8_t* d_dict, size_t dict_size);
int decode_host_packed(DeviceCtx* c, size_t n, const uint8_t* in_bytes, const uint64_t* in_off, uint8_t* out_bytes,
                       const uint64_t* out_off, uint64_t* out_len, int32_t* codes, uint64_t* in_used, uint32_t large_window,
                       const uint8_t* dict = nullptr, size_t dict_size = 0) {
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lock(c->mu);
  if (dict_size) {
    CU_TRY(c->cdict.reserve(dict_size + 64));  // 32 bytes of slack on either side: the lane kernel reads whole aligned 16-byte blocks
    CU_TRY(cudaMemcpyAsync((uint8_t*)c->cdict.p + 32, dict, dict_size, cudaMemcpyHostToDevice, c->s_h2d));  // ordered before the first chunk's h2d event
  }
  const uint64_t in_base = in_off[0], out_base = out_off[0];
  const uint64_t in_total = in_off[n] - in_base, out_total = out_off[n] - out_base;
  for (size_t i = 0; i < n; i++)
    if (in_off[i + 1] < in_off[i] || out_off[i + 1] < out_off[i]) { set_error("brotli_b200: offsets must be non-decreasing"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
  CU_TRY(c->in.reserve(in_total + 16));
  CU_TRY(c->out.reserve(out_total + 16));
  CU_TRY(c->in_off.reserve((n + 1) * 8));
  CU_TRY(c->out_off.reserve((n + 1) * 8));
  CU_TRY(c->out_len.reserve(n * 8));
  CU_TRY(c->codes.reserve(n * 4));
  if (in_used) CU_TRY(c->in_used.reserve(n * 8));
  uint8_t* d_in = (uint8_t*)c->in.p - in_base;   // absolute offsets index the device mirrors directly
  uint8_t* d_out = (uint8_t*)c->out.p - out_base;
  CU_TRY(cudaMemcpyAsync(c->in_off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));
  CU_TRY(cudaMemcpyAsync(c->out_off.p, out_off, (n + 1) * 8, cudaMemcpyHostToDevice, c->s_h2d));

  struct Chunk { size_t b, e; cudaEvent_t h2d, done; cudaEvent_t d2h = nullptr; };
  static const bool trace = getenv("BROTLI_B200_PIPE_TRACE") != nullptr;  // per-chunk timeline on stderr
  std::vector<Chunk> chunks;
  for (size_t b = 0; b < n;) {
    size_t e = b; uint64_t acc = 0;
    // A chunk should fill every resident lane of the lane kernel (a launch takes about as long for a few streams as for
    // one stream per lane), but nothing can leave before the first chunk is up and decoded, and the D2H stream -- the
    // bottleneck: 2.6x the bytes of the H2D stream -- should never wait for a decode.  So the chunks ramp up: 1/4 and
    // 1/2 of the resident lanes, then full waves (BROTLI_B200_PIPE_RAMP=k: 1/2^k, ..., 1/2, 1; 0: the plan of round 1,
    // a single 0.4 chunk first).  Measured on the headline call, profiles/r02/e2e_ramp.txt; a remainder below a quarter
    // wave joins the chunk before it.
    const size_t lanes = c->lane_ctas > 0 ? (size_t)c->lane_ctas * c->lane_warps * 32 : 1;
    size_t min_streams = lanes;
    {
      static const int ramp = getenv("BROTLI_B200_PIPE_RAMP") && getenv("BROTLI_B200_PIPE_RAMP")[0] ? atoi(getenv("BROTLI_B200_PIPE_RAMP")) : 2;
      const int ci = (int)chunks.size();
      if (ramp > 0) { if (ci < ramp) min_streams >>= (ramp - ci); }
      else if (b == 0) min_streams = min_streams * 2 / 5;
      if (min_streams == 0) min_streams = 1;
    }
    while (e < n && (e == b || acc < kPipelineChunkBytes || e - b < min_streams)) { acc += (in_off[e + 1] - in_off[e]) + (out_off[e + 1] - out_off[e]); e++; }
    if (n - e < lanes / 4) e = n;
    chunks.push_back(Chunk{b, e, nullptr, nullptr, nullptr});
    b = e;
  }
  int rc = 0;
  for (auto& k : chunks) {
    const unsigned flags = trace ? cudaEventDefault : cudaEventDisableTiming;
    if (trace) cudaEventCreate(&k.d2h);
    if (cudaEventCreateWithFlags(&k.h2d, flags) != cudaSuccess ||
        cudaEventCreateWithFlags(&k.done, flags) != cudaSuccess) { rc = BROTLI_DECODER_ERROR_UNREACHABLE; set_error("brotli_b200: cudaEventCreate failed"); break; }
  }
  bool timed = false;
  for (size_t ci = 0; ci < chunks.size() && rc == 0; ci++) {
    Chunk& k = chunks[ci];
    const uint64_t i0 = in_off[k.b], i1 = in_off[k.e], o0 = out_off[k.b], o1 = out_off[k.e];
    cudaError_t e = cudaSuccess;
    if (i1 > i0) e = cudaMemcpyAsync(d_in + i0, in_bytes + i0, i1 - i0, cudaMemcpyHostToDevice, c->s_h2d);
    if (e == cudaSuccess) e = cudaEventRecord(k.h2d, c->s_h2d);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_compute, k.h2d, 0);
    if (e == cudaSuccess && ci == 0) { e = cudaEventRecord(c->ev_k0, c->s_compute); timed = true; }
    if (e != cudaSuccess) { set_error(std::string("brotli_b200: H2D stage failed: ") + cudaGetErrorString(e)); rc = BROTLI_DECODER_ERROR_UNREACHABLE; break; }
    rc = decode_device(c, k.e - k.b, d_in, (const uint64_t*)c->in_off.p + k.b, d_out, (const uint64_t*)c->out_off.p + k.b,
                       (uint64_t*)c->out_len.p + k.b, (int32_t*)c->codes.p + k.b, in_used ? (uint64_t*)c->in_used.p + k.b : nullptr,
                       large_window, c->s_compute, dict_size ? (const uint8_t*)c->cdict.p + 32 : nullptr, dict_size);
    if (rc != 0) break;
    if (ci + 1 == chunks.size()) cudaEventRecord(c->ev_k1, c->s_compute);
    e = cudaEventRecord(k.done, c->s_compute);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->s_d2h, k.done, 0);
    if (e == cudaSuccess && o1 > o0) e = cudaMemcpyAsync(out_bytes + o0, d_out + o0, o1 - o0, cudaMemcpyDeviceToHost, c->s_d2h);
    if (trace && k.d2h) cudaEventRecord(k.d2h, c->s_d2h);
    if (e != cudaSuccess) { set_error(std::string("brotli_b200: D2H stage failed: ") + cudaGetErrorString(e)); rc = BROTLI_DECODER_ERROR_UNREACHABLE; break; }
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_len, c->out_len.p, n * 8, cudaMemcpyDeviceToHost, c->s_d2h);
    if (e == cudaSuccess) e = cudaMemcpyAsync(codes, c->codes.p, n * 4, cudaMemcpyDeviceToHost, c->s_d2h);
    if (e == cudaSuccess && in_used) e = cudaMemcpyAsync(in_used, c->in_used.p, n * 8, cudaMemcpyDeviceToHost, c->s_d2h);
    if (e != cudaSuccess) { set_error(std::string("brotli_b200: result copy failed: ") + cudaGetErrorString(e)); rc = BROTLI_DECODER_ERROR_UNREACHABLE; }
  }
  cudaError_t e1 = cudaStreamSynchronize(c->s_h2d), e2 = cudaStreamSynchronize(c->s_compute), e3 = cudaStreamSynchronize(c->s_d2h);
  if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
    cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
    set_error(std::string("brotli_b200: batch failed on the device: ") + cudaGetErrorString(e));
    rc = BROTLI_DECODER_ERROR_UNREACHABLE;
  }
  if (rc == 0 && timed) { float ms = 0; if (cudaEventElapsedTime(&ms, c->ev_k0, c->ev_k1) == cudaSuccess) g_last_kernel_ms.store(ms); }
  if (rc == 0) rc = redo_needs_more_output(c, n, in_bytes, in_off, out_off, d_in, out_len, codes, dict_size ? (const uint8_t*)c->cdict.p + 32 : nullptr, dict_size);
  if (trace && rc == 0 && !chunks.empty()) {
    for (size_t ci = 0; ci < chunks.size(); ci++) {
      float th = 0, td = 0, to = 0;
      cudaEventElapsedTime(&th, chunks[0].h2d, chunks[ci].h2d);  // relative to the end of the first H2D copy
      cudaEventElapsedTime(&td, chunks[0].h2d, chunks[ci].done);
      if (chunks[ci].d2h) cudaEventElapsedTime(&to, chunks[0].h2d, chunks[ci].d2h);
      fprintf(stderr, "brotli_b200 pipe: chunk %zu streams %zu  h2d_done %.1f ms  decode_done %.1f ms  d2h_done %.1f ms\n", ci,
              chunks[ci].e - chunks[ci].b, th, td, to);
    }
  }
  for (auto& k : chunks) { if (k.h2d) cudaEventDestroy(k.h2d); if (k.done) cudaEventDestroy(k.done); if (k.d2h) cudaEventDestroy(k.d2h); }
  return rc;
}

// ---- streaming sessions and exact re-decodes: the device backend of brotli_b200_session.h ----------------
// All of it runs on s_compute under c->mu (the staging buffers are the context's).
struct CudaDev {
  DeviceCtx* c;
  // Size-class allocator over cudaMalloc: every live block has its class in a host-side map (address -> class).
  std::unordered_map<uint8_t*, int>& classes() { return c->sess_class; }
  uint8_t* alloc(size_t n) {
    int k = 12;  // 4 KiB minimum
    while (((size_t)1 << k) < n) k++;
    if (k >= 48) return nullptr;
    auto& fl = c->sess_free[k];
    if (fl.empty()) {
      const size_t sz = (size_t)1 << k;
      const size_t per_slab = sz <= ((size_t)4 << 20) ? (((size_t)32 << 20) / sz > 64 ? 64 : ((size_t)32 << 20) / sz) : 1;  // up to 32 MB / 64 blocks per cudaMalloc
      void* p = nullptr;
      if (cudaMalloc(&p, sz * per_slab) != cudaSuccess) {
        cudaGetLastError();
        if (per_slab == 1 || cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        c->sess_slabs.push_back(p);
        fl.push_back((uint8_t*)p);
      } else {
        c->sess_slabs.push_back(p);
        for (size_t j = 0; j < per_slab; j++) fl.push_back((uint8_t*)p + j * sz);
      }
    }
    uint8_t* r = fl.back(); fl.pop_back();
    classes()[r] = k;
    return r;
  }
  void release(uint8_t* p) {
    if (!p) return;
    auto it = classes().find(p);
    if (it == classes().end()) return;
    c->sess_free[it->second].push_back(p);
    classes().erase(it);
  }
  uint8_t* host_up(size_t bytes) {
    if (bytes > c->sess_pin_up_cap) {
      if (c->sess_pin_up) cudaFreeHost(c->sess_pin_up);
      c->sess_pin_up = nullptr; c->sess_pin_up_cap = 0;
      const size_t want = bytes + bytes / 2 + 65536;
      if (cudaHostAlloc(&c->sess_pin_up, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      c->sess_pin_up_cap = want;
    }
    return (uint8_t*)c->sess_pin_up;
  }
  size_t arena_bytes() const { return brotli_b200::arena_bytes_per_warp(); }
  int upload(uint8_t* d, const uint8_t* h, size_t n) {
    return cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : BROTLI_DECODER_ERROR_UNREACHABLE;
  }
  // pieces (with their blob-relative side rebased onto the device blob) -> device, then the copy kernel
  int copy_pieces(std::vector<brotli_b200::SessionCopy>& v) {
    CU_TRY(c->sess_pieces.reserve(v.size() * sizeof(brotli_b200::SessionCopy)));
    CU_TRY(cudaMemcpyAsync(c->sess_pieces.p, v.data(), v.size() * sizeof(brotli_b200::SessionCopy), cudaMemcpyHostToDevice, c->s_compute));
    CU_TRY(brotli_b200::launch_session_copy((const brotli_b200::SessionCopy*)c->sess_pieces.p, (uint32_t)v.size(), c->s_compute));
    g_launches.fetch_add(1);
    return 0;
  }
  int run(brotli_b200::ResumeState* arr, uint32_t n, const uint8_t* blob, size_t blob_bytes, const brotli_b200::SessionCopy* scatter, uint32_t n_scatter) {
    if (n_scatter) {
      CU_TRY(c->sess_blob.reserve(blob_bytes + 16));
      CU_TRY(cudaMemcpyAsync(c->sess_blob.p, blob, blob_bytes, cudaMemcpyHostToDevice, c->s_compute));
      std::vector<brotli_b200::SessionCopy> v(scatter, scatter + n_scatter);
      for (auto& k : v) k.src = (const uint8_t*)c->sess_blob.p + (uintptr_t)k.src;
      int rc = copy_pieces(v);
      if (rc != 0) return rc;
      // (no synchronisation: the blob is the pinned staging buffer of host_up(), reused only after this call's final
      // synchronisation, and a copy from the pageable `v` is staged before cudaMemcpyAsync returns)
    }
    CU_TRY(c->sess_states.reserve((size_t)n * sizeof(brotli_b200::ResumeState)));
    CU_TRY(cudaMemcpyAsync(c->sess_states.p, arr, (size_t)n * sizeof(brotli_b200::ResumeState), cudaMemcpyHostToDevice, c->s_compute));
    int rc = decode_device(c, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1u, c->s_compute, nullptr, 0,
                           (brotli_b200::ResumeState*)c->sess_states.p);
    if (rc != 0) return rc;
    CU_TRY(cudaMemcpyAsync(arr, c->sess_states.p, (size_t)n * sizeof(brotli_b200::ResumeState), cudaMemcpyDeviceToHost, c->s_compute));
    CU_TRY(cudaStreamSynchronize(c->s_compute));
    return 0;
  }
  const uint8_t* gather(const brotli_b200::SessionCopy* pieces, uint32_t n, size_t bytes) {
    if (bytes + 16 > c->sess_pin_down_cap) {
      if (c->sess_pin_down) cudaFreeHost(c->sess_pin_down);
      c->sess_pin_down = nullptr; c->sess_pin_down_cap = 0;
      const size_t want = bytes + bytes / 2 + 65536;
      if (cudaHostAlloc(&c->sess_pin_down, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      c->sess_pin_down_cap = want;
    }
    if (c->sess_blob.reserve(bytes + 16) != cudaSuccess) return nullptr;
    std::vector<brotli_b200::SessionCopy> v(pieces, pieces + n);
    for (auto& k : v) k.dst = (uint8_t*)c->sess_blob.p + (uintptr_t)k.dst;
    if (copy_pieces(v) != 0) return nullptr;
    if (cudaMemcpyAsync(c->sess_pin_down, c->sess_blob.p, bytes, cudaMemcpyDeviceToHost, c->s_compute) != cudaSuccess ||
        cudaStreamSynchronize(c->s_compute) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (const uint8_t*)c->sess_pin_down;
  }
  int move(const brotli_b200::SessionCopy* pieces, uint32_t n) {
    std::vector<brotli_b200::SessionCopy> v(pieces, pieces + n);
    // (stream order is all the later launches need; the host never reads these buffers)
    return copy_pieces(v);
  }
};

// The reference decodes a whole ring buffer ahead of the caller's buffer, so a stream that is corrupt (or ends) beyond a too
// small output capacity reports the corruption (or NeedsMoreInput), not NeedsMoreOutput (src/decode.rs:1693-1738: the
// capacity is only noticed at a ring flush point).  The batch kernels cannot write past a stream's region; streams they
// leave at NeedsMoreOutput are decoded again here into scratch windows that reach the next flush point, with the
// capacity as the decoder's budget: code and decoded_size then are exactly the reference's.  `d_in` holds the batch's
// input (absolute offsets), results are patched in the host arrays.  Runs under c->mu.
int redo_needs_more_output(DeviceCtx* c, size_t n, const uint8_t* in_bytes, const uint64_t* in_off, const uint64_t* out_off, const uint8_t* d_in,
                           uint64_t* out_len, int32_t* codes, const uint8_t* d_dict, size_t dict_size) {
  static const bool enabled = !(getenv("BROTLI_B200_EXACT_REDO") && getenv("BROTLI_B200_EXACT_REDO")[0] == '0');
  if (!enabled) return 0;
  std::vector<uint32_t> idx;
  for (size_t i = 0; i < n; i++) if (codes[i] == BROTLI_DECODER_NEEDS_MORE_OUTPUT) idx.push_back((uint32_t)i);
  if (idx.empty()) return 0;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
  const size_t group_limit = free_b / 2 > ((size_t)4 << 30) ? ((size_t)4 << 30) : free_b / 2;  // scratch per launch
  CudaDev dev{c};
  size_t k = 0;
  while (k < idx.size()) {
    std::vector<brotli_b200::ResumeState> arr;
    std::vector<uint32_t> who;
    size_t scratch = 0;
    for (; k < idx.size(); k++) {
      const uint32_t i = idx[k];
      const uint8_t* p = in_bytes + in_off[i];
      const size_t sz = (size_t)(in_off[i + 1] - in_off[i]);
      const uint64_t cap = out_off[i + 1] - out_off[i];
      // WBITS (src/decode.rs:152-187): the ring buffer is at most 1 << wbits, flush points are its multiples
      uint32_t wbits = 16;
      if (sz >= 1 && (p[0] & 1)) {
        const uint32_t n3 = (p[0] >> 1) & 7;
        if (n3) wbits = 17 + n3;
        else { const uint32_t m = (p[0] >> 4) & 7; wbits = m == 1 ? (sz >= 2 ? (uint32_t)(p[1] & 63) : 30u) : (m ? 8 + m : 17); }
      }
      if (wbits > 30) wbits = 30;
      const uint64_t ring = (uint64_t)1 << wbits;
      const uint64_t need = (cap / ring + 1) * ring + 64;  // first flush point beyond the capacity
      if (need > group_limit) continue;                      // no room for the scratch window: the first pass's answer stands
      if (scratch + need + 16 > group_limit && !arr.empty()) break;
      brotli_b200::ResumeState r;
      memset(&r, 0, sizeof(r));
      r.in = d_in + in_off[i]; r.in_size = sz;
      r.out = (uint8_t*)(uintptr_t)scratch;  // offset for now
      r.out_cap = need; r.budget = cap;
      r.arena = nullptr; r.dict = d_dict; r.dict_size = dict_size; r.allow_large_window = 1;
      arr.push_back(r); who.push_back(i);
      scratch += (need + 15) & ~(uint64_t)15;
    }
    if (arr.empty()) continue;
    CU_TRY(c->redo_out.reserve(scratch + 64));
    for (auto& r : arr) r.out = (uint8_t*)c->redo_out.p + (uintptr_t)r.out;
    int rc = dev.run(arr.data(), (uint32_t)arr.size(), nullptr, 0, nullptr, 0);
    if (rc != 0) return rc;
    for (size_t j = 0; j < arr.size(); j++) {
      const brotli_b200::ResumeState& r = arr[j];
      const uint32_t i = who[j];
      const uint64_t cap = out_off[i + 1] - out_off[i];
      if (r.hit_cap) continue;  // (cannot happen: the window reaches the flush point)
      if (r.code < 0) { codes[i] = r.code; out_len[i] = r.flushed_now < cap ? r.flushed_now : cap; }
      else if (r.code == BROTLI_DECODER_NEEDS_MORE_INPUT) { codes[i] = r.code; out_len[i] = r.decoded < cap ? r.decoded : cap; }
      // NeedsMoreOutput at the flush point, or the stream ends beyond the capacity: NeedsMoreOutput with a full buffer stands
    }
  }
  return 0;
}

// is_valid_slice_ptr, src/ffi/mod.rs:45-60
template <typename T>
bool valid_slice(const T* p, size_t len, size_t align) {
  if (len == 0) return true;
  if (p == nullptr) return false;
  if (((uintptr_t)p) % align != 0) return false;
  if (len > (size_t)INTPTR_MAX / sizeof(T)) return false;
  return (uintptr_t)p + len * sizeof(T) >= (uintptr_t)p;
}

BrotliDecoderReturnInfo make_info(int result, int code, size_t decoded, const char* msg) {
  BrotliDecoderReturnInfo r;
  memset(&r, 0, sizeof(r));
  r.decoded_size = decoded;
  r.result = (BrotliDecoderResult)result;
  r.code = (BrotliDecoderErrorCode)code;
  strncpy(r.error, msg ? msg : error_name(code), sizeof(r.error) - 1);
  return r;
}

BrotliDecoderReturnInfo invalid_arguments_info() {  // src/ffi/mod.rs:83-106
  return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_INVALID_ARGUMENTS, 0, nullptr);
}

// One stream, host buffers: no chunk pipeline, one launch, and only the decoded bytes come back (the caller's buffer is
// written exactly as far as the reference writes it: c/main.c passes a capacity larger than its buffer).
int decode_one(DeviceCtx* c, const uint8_t* in, size_t in_size, uint8_t* out, size_t out_cap, uint64_t* out_len, int32_t* code, uint64_t* used,
               uint32_t large_window, const uint8_t* dict, size_t dict_size) {
  std::lock_guard<std::mutex> lock(c->mu);
  const uint8_t* d_dict = nullptr;
  if (dict_size) {
    CU_TRY(c->cdict.reserve(dict_size + 64));
    CU_TRY(cudaMemcpyAsync((uint8_t*)c->cdict.p + 32, dict, dict_size, cudaMemcpyHostToDevice, c->s_compute));
    d_dict = (const uint8_t*)c->cdict.p + 32;
  }
  CU_TRY(c->in.reserve(in_size + 16));
  CU_TRY(c->out.reserve(out_cap + 16));
  CU_TRY(c->in_off.reserve(64));
  CU_TRY(c->out_len.reserve(64));
  // meta block in one device buffer: in_off[2] | out_off[2] | out_len | in_used | code
  uint64_t meta[8] = {0, (uint64_t)in_size, 0, (uint64_t)out_cap, 0, 0, 0, 0};
  uint64_t* d_meta = (uint64_t*)c->in_off.p;
  CU_TRY(cudaMemcpyAsync(d_meta, meta, sizeof(meta), cudaMemcpyHostToDevice, c->s_compute));
  if (in_size) CU_TRY(cudaMemcpyAsync(c->in.p, in, in_size, cudaMemcpyHostToDevice, c->s_compute));
  int rc = decode_device(c, 1, (const uint8_t*)c->in.p, d_meta, (uint8_t*)c->out.p, d_meta + 2, d_meta + 4, (int32_t*)(d_meta + 6), d_meta + 5,
                         large_window, c->s_compute, d_dict, dict_size);
  if (rc != 0) return rc;
  CU_TRY(cudaMemcpyAsync(meta, d_meta, sizeof(meta), cudaMemcpyDeviceToHost, c->s_compute));
  CU_TRY(cudaStreamSynchronize(c->s_compute));
  *out_len = meta[4]; *used = meta[5]; *code = (int32_t)(uint32_t)meta[6];
  if (*out_len > out_cap) *out_len = out_cap;
  if (*out_len) {
    CU_TRY(cudaMemcpyAsync(out, c->out.p, (size_t)*out_len, cudaMemcpyDeviceToHost, c->s_compute));
    CU_TRY(cudaStreamSynchronize(c->s_compute));
  }
  if (*code == BROTLI_DECODER_NEEDS_MORE_OUTPUT) {
    const uint64_t in_off[2] = {0, (uint64_t)in_size}, out_off[2] = {0, (uint64_t)out_cap};
    rc = redo_needs_more_output(c, 1, in, in_off, out_off, (const uint8_t*)c->in.p, out_len, code, d_dict, dict_size);
  }
  return rc;
}

// brotli_decode (src/lib.rs:446-468) on the GPU: a batch of one.
BrotliDecoderReturnInfo one_shot(const uint8_t* in, size_t in_size, uint8_t* out, size_t out_cap, uint32_t large_window,
                                 uint64_t* in_used, const uint8_t* dict = nullptr, size_t dict_size = 0) {
  DeviceCtx* c = acquire_ctx();
  if (!c) return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, tl_error.c_str());
  if (in_size >= ((uint64_t)1 << 32)) return invalid_arguments_info();  // src/decode.rs:2799-2812
  uint64_t out_len = 0, used = 0;
  int32_t code = 0;
  static const uint8_t kNothing[1] = {0};
  uint8_t dummy_out[1];
  int rc = decode_one(c, in ? in : kNothing, in_size, out ? out : dummy_out, out_cap, &out_len, &code, &used, large_window, dict, dict_size);
  if (rc != 0) return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, tl_error.c_str());
  if (in_used) *in_used = used;
  int result = code == 1 ? 1 : (code == 2 ? 2 : (code == 3 ? 3 : 0));  // BrotliResult
  return make_info(result, code, (size_t)out_len, nullptr);
}

}  // namespace

// =========================================== C ABI ===========================================
extern "C" {

BrotliDecoderResult BrotliDecoderDecompress(size_t encoded_size, const uint8_t* encoded_buffer, size_t* decoded_size,
                                            uint8_t* decoded_buffer) {
  try {
    if (!valid_slice(decoded_size, 1, alignof(size_t))) return BROTLI_DECODER_RESULT_ERROR;
    BrotliDecoderReturnInfo r = BrotliDecoderDecompressWithReturnInfo(encoded_size, encoded_buffer, *decoded_size, decoded_buffer);
    *decoded_size = r.decoded_size;
    return r.result == BROTLI_DECODER_RESULT_SUCCESS ? BROTLI_DECODER_RESULT_SUCCESS : BROTLI_DECODER_RESULT_ERROR;
  } catch (...) {
    if (decoded_size) *decoded_size = 0;
    return BROTLI_DECODER_RESULT_ERROR;
  }
}

BrotliDecoderReturnInfo BrotliDecoderDecompressWithReturnInfo(size_t encoded_size, const uint8_t* encoded_buffer,
                                                              size_t decoded_size, uint8_t* decoded_buffer) {
  try {
    if (!valid_slice(encoded_buffer, encoded_size, 1) || !valid_slice(decoded_buffer, decoded_size, 1)) return invalid_arguments_info();
    return one_shot(encoded_buffer, encoded_size, decoded_buffer, decoded_size, 1u, nullptr);
  } catch (const std::exception& e) {
    return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, e.what());
  } catch (...) {
    return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, "brotli_b200: unknown exception");
  }
}

BrotliDecoderReturnInfo BrotliDecoderDecompressPrealloc(size_t encoded_size, const uint8_t* encoded_buffer, size_t decoded_size,
                                                        uint8_t* decoded_buffer, size_t scratch_u8_size, uint8_t* scratch_u8_buffer,
                                                        size_t scratch_u32_size, uint32_t* scratch_u32_buffer,
                                                        size_t scratch_hc_size, HuffmanCode* scratch_hc_buffer) {
  if (!valid_slice(scratch_u8_buffer, scratch_u8_size, 1) || !valid_slice(scratch_u32_buffer, scratch_u32_size, alignof(uint32_t)) ||
      !valid_slice(scratch_hc_buffer, scratch_hc_size, alignof(uint16_t)))
    return invalid_arguments_info();
  return BrotliDecoderDecompressWithReturnInfo(encoded_size, encoded_buffer, decoded_size, decoded_buffer);
}

// ---- streaming state: a device-resident session (brotli_b200_session.h) ----------------------
struct BrotliDecoderStateStruct {
  brotli_alloc_func alloc_func;
  brotli_free_func free_func;
  void* opaque;
  brotli_b200::Session sess;     // windows of the stream in device memory, checkpoint, pending output
  int session_device;            // CUDA device that owns the session buffers (-1: none yet)
  char error[256];
};

BrotliDecoderState* BrotliDecoderCreateInstance(brotli_alloc_func alloc_func, brotli_free_func free_func, void* opaque) {
  if ((alloc_func == nullptr) != (free_func == nullptr)) return nullptr;  // src/ffi/mod.rs:132-135
  void* mem = alloc_func ? alloc_func(opaque, sizeof(BrotliDecoderStateStruct)) : malloc(sizeof(BrotliDecoderStateStruct));
  if (!mem) return nullptr;
  BrotliDecoderStateStruct* s = new (mem) BrotliDecoderStateStruct();
  s->alloc_func = alloc_func; s->free_func = free_func; s->opaque = opaque;
  s->sess.large_window = false;  // src/ffi/mod.rs:127
  s->session_device = -1;
  s->error[0] = 0;
  return s;
}

void BrotliDecoderDestroyInstance(BrotliDecoderState* s) {
  if (!s) return;
  if (s->session_device >= 0) {
    DeviceCtx* c = &g_ctx[s->session_device];
    std::lock_guard<std::mutex> lock(c->mu);
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(s->session_device);
    CudaDev dev{c};
    brotli_b200::SessionRunner<CudaDev>(dev).destroy(s->sess);
    cudaSetDevice(prev);
  }
  brotli_free_func f = s->free_func; void* opaque = s->opaque;
  s->~BrotliDecoderStateStruct();
  if (f) f(opaque, s); else free(s);
}

int BrotliDecoderSetParameter(BrotliDecoderState* s, BrotliDecoderParameter param, uint32_t value) {
  if (!s || s->sess.used || s->sess.stop != brotli_b200::Session::kFresh) return 0;  // src/ffi/mod.rs:163-166
  switch (param) {
    case BROTLI_DECODER_PARAM_DISABLE_RING_BUFFER_REALLOCATION: return 1;  // no ring buffer exists on the GPU path
    case BROTLI_DECODER_PARAM_LARGE_WINDOW: s->sess.large_window = value != 0; return 1;
    default: return 0;
  }
}

static BrotliDecoderResult stream_fail(BrotliDecoderState* s, int code, const char* msg) {
  s->sess.stop = brotli_b200::Session::kFailed; s->sess.code = code;
  strncpy(s->error, msg ? msg : error_name(code), sizeof(s->error) - 1);
  return BROTLI_DECODER_RESULT_ERROR;
}

// n BrotliDecoderDecompressStream calls, one per state, served by ONE decode launch (src/ffi/mod.rs:389-463 per state).
int BrotliB200DecoderDecompressStreamBatch(size_t n, BrotliDecoderState* const* states, size_t* available_in, const uint8_t** next_in,
                                           size_t* available_out, uint8_t** next_out, size_t* total_out, BrotliDecoderResult* results) {
  try {
    if (n == 0) return 0;
    if (!states || !available_in || !next_in || !available_out || !next_out || !results) { set_error("brotli_b200: null batch array"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
    std::vector<brotli_b200::Session*> ss; std::vector<brotli_b200::StreamCall> cs; std::vector<size_t> who;
    DeviceCtx* c = nullptr;
    for (size_t i = 0; i < n; i++) {
      BrotliDecoderState* s = states[i];
      results[i] = BROTLI_DECODER_RESULT_ERROR;
      if (!s) continue;
      if ((available_in[i] && !next_in[i]) || (available_out[i] && !next_out[i])) { stream_fail(s, BROTLI_DECODER_ERROR_INVALID_ARGUMENTS, nullptr); continue; }
      if (!c) { c = acquire_ctx(); if (!c) { for (size_t j = 0; j < n; j++) if (states[j]) stream_fail(states[j], BROTLI_DECODER_ERROR_UNREACHABLE, tl_error.c_str()); return BROTLI_DECODER_ERROR_UNREACHABLE; } }
      if (s->session_device < 0) s->session_device = c->device;
      if (s->session_device != c->device) {  // the session's buffers live on the device of its first call
        stream_fail(s, BROTLI_DECODER_ERROR_UNREACHABLE, "brotli_b200: decoder state used from another CUDA device"); continue;
      }
      ss.push_back(&s->sess);
      cs.push_back(brotli_b200::StreamCall{&available_in[i], &next_in[i], &available_out[i], &next_out[i], total_out ? &total_out[i] : nullptr, -1});
      who.push_back(i);
    }
    if (ss.empty()) return 0;
    int rc;
    {
      std::lock_guard<std::mutex> lock(c->mu);
      CudaDev dev{c};
      rc = brotli_b200::SessionRunner<CudaDev>(dev).stream_calls(ss.data(), cs.data(), ss.size(), BROTLI_DECODER_ERROR_UNREACHABLE);
    }
    for (size_t k = 0; k < who.size(); k++) {
      BrotliDecoderState* s = states[who[k]];
      results[who[k]] = (BrotliDecoderResult)cs[k].result;
      if (cs[k].result == BROTLI_DECODER_RESULT_ERROR && !s->error[0])
        strncpy(s->error, s->sess.code == BROTLI_DECODER_ERROR_UNREACHABLE && rc != 0 ? tl_error.c_str() : error_name(s->sess.code), sizeof(s->error) - 1);
    }
    return rc;
  } catch (...) { set_error("brotli_b200: exception"); return BROTLI_DECODER_ERROR_UNREACHABLE; }
}

BrotliDecoderResult BrotliDecoderDecompressStream(BrotliDecoderState* s, size_t* available_in, const uint8_t** next_in,
                                                  size_t* available_out, uint8_t** next_out, size_t* total_out) {
  if (!s) return BROTLI_DECODER_RESULT_ERROR;
  try {
    if (!available_in || !next_in || !available_out || !next_out) return stream_fail(s, BROTLI_DECODER_ERROR_INVALID_ARGUMENTS, nullptr);  // src/ffi/mod.rs:397-407
    BrotliDecoderResult r = BROTLI_DECODER_RESULT_ERROR;
    BrotliDecoderState* one[1] = {s};
    BrotliB200DecoderDecompressStreamBatch(1, one, available_in, next_in, available_out, next_out, total_out, &r);
    return r;
  } catch (const std::exception& e) {
    return stream_fail(s, BROTLI_DECODER_ERROR_UNREACHABLE, e.what());
  } catch (...) {
    return stream_fail(s, BROTLI_DECODER_ERROR_UNREACHABLE, "brotli_b200: unknown exception");
  }
}

BrotliDecoderResult BrotliDecoderDecompressStreaming(BrotliDecoderState* s, size_t* available_in, const uint8_t* next_in,
                                                     size_t* available_out, uint8_t* next_out) {
  if (!available_in || !available_out) return BrotliDecoderDecompressStream(s, nullptr, nullptr, nullptr, nullptr, nullptr);
  const uint8_t* in = next_in; uint8_t* out = next_out;
  return BrotliDecoderDecompressStream(s, available_in, &in, available_out, &out, nullptr);
}

int BrotliDecoderHasMoreOutput(const BrotliDecoderState* s) {
  return s && s->sess.stop != brotli_b200::Session::kFailed && s->sess.pending_bytes() != 0 ? 1 : 0;
}

const uint8_t* BrotliDecoderTakeOutput(BrotliDecoderState* s, size_t* size) {
  if (!s || !size) return nullptr;
  const size_t avail = s->sess.pending_bytes();
  const size_t n = *size == 0 ? avail : (*size < avail ? *size : avail);
  if (s->sess.stop == brotli_b200::Session::kFailed || n == 0) { *size = 0; return nullptr; }
  const uint8_t* p = s->sess.pending.data() + s->sess.pending_off;  // stays valid until the next call on this state
  s->sess.pending_off += n; s->sess.delivered += n; *size = n;
  return p;
}

int BrotliDecoderIsUsed(const BrotliDecoderState* s) { return s && s->sess.used ? 1 : 0; }
int BrotliDecoderIsFinished(const BrotliDecoderState* s) {
  return s && s->sess.stop == brotli_b200::Session::kDone && s->sess.pending_bytes() == 0 ? 1 : 0;
}
BrotliDecoderErrorCode BrotliDecoderGetErrorCode(const BrotliDecoderState* s) { return (BrotliDecoderErrorCode)(s ? s->sess.code : 0); }
const char* BrotliDecoderGetErrorString(const BrotliDecoderState* s) { return s && s->error[0] ? s->error : ""; }
const char* BrotliDecoderErrorString(BrotliDecoderErrorCode c) { return error_name((int)c); }
uint32_t BrotliDecoderVersion(void) { return 0x1000f00; }  // src/ffi/mod.rs:588-590

uint8_t* BrotliDecoderMallocU8(BrotliDecoderState* s, size_t size) {
  return (uint8_t*)(s && s->alloc_func ? s->alloc_func(s->opaque, size) : malloc(size));
}
void BrotliDecoderFreeU8(BrotliDecoderState* s, uint8_t* data, size_t) {
  if (s && s->free_func) s->free_func(s->opaque, data); else free(data);
}
size_t* BrotliDecoderMallocUsize(BrotliDecoderState* s, size_t size) {
  return (size_t*)(s && s->alloc_func ? s->alloc_func(s->opaque, size * sizeof(size_t)) : malloc(size * sizeof(size_t)));
}
void BrotliDecoderFreeUsize(BrotliDecoderState* s, size_t* data, size_t) {
  if (s && s->free_func) s->free_func(s->opaque, data); else free(data);
}

// ---- batch extension -------------------------------------------------------------------------
int BrotliB200DecompressBatchDevice(size_t n, const uint8_t* d_in_bytes, const uint64_t* d_in_off, uint8_t* d_out_bytes,
                                    const uint64_t* d_out_off, uint64_t* d_out_len, int32_t* d_codes, void* cuda_stream) {
  try {
    if (n == 0) return 0;
    if (!d_in_off || !d_out_off || !d_out_len || !d_codes) { set_error("brotli_b200: null batch array"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
    DeviceCtx* c = acquire_ctx();
    if (!c) return BROTLI_DECODER_ERROR_UNREACHABLE;
    return decode_device(c, n, d_in_bytes, d_in_off, d_out_bytes, d_out_off, d_out_len, d_codes, nullptr, 1u, (cudaStream_t)cuda_stream);
  } catch (...) { set_error("brotli_b200: exception"); return BROTLI_DECODER_ERROR_UNREACHABLE; }
}

int BrotliB200DecompressBatchPacked(size_t n, const uint8_t* in_bytes, const uint64_t* in_off, uint8_t* out_bytes,
                                    const uint64_t* out_off, uint64_t* out_len, int32_t* codes) {
  try {
    if (n == 0) return 0;
    if (!in_off || !out_off || !out_len || !codes || !in_bytes || !out_bytes) { set_error("brotli_b200: null batch array"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
    DeviceCtx* c = acquire_ctx();
    if (!c) return BROTLI_DECODER_ERROR_UNREACHABLE;
    return decode_host_packed(c, n, in_bytes, in_off, out_bytes, out_off, out_len, codes, nullptr, 1u);
  } catch (...) { set_error("brotli_b200: exception"); return BROTLI_DECODER_ERROR_UNREACHABLE; }
}

int BrotliB200DecompressBatch(size_t n, const uint8_t* const* in, const size_t* in_size, uint8_t* const* out, size_t* out_size,
                              BrotliDecoderResult* results, BrotliDecoderErrorCode* codes) {
  try {
    if (n == 0) return 0;
    if (!in || !in_size || !out || !out_size || !results) { set_error("brotli_b200: null batch array"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
    DeviceCtx* c = acquire_ctx();
    if (!c) return BROTLI_DECODER_ERROR_UNREACHABLE;
    // pack the scattered streams into one blob (8-byte aligned slots), decode, scatter back
    std::vector<uint64_t> in_off(n + 1), out_off(n + 1), out_len(n);
    std::vector<int32_t> cds(n);
    uint64_t ia = 0, oa = 0;
    for (size_t i = 0; i < n; i++) {
      if (!valid_slice(in[i], in_size[i], 1) || !valid_slice(out[i], out_size[i], 1)) { set_error("brotli_b200: invalid stream pointer"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
      in_off[i] = ia; out_off[i] = oa; ia += in_size[i]; oa += out_size[i];
    }
    in_off[n] = ia; out_off[n] = oa;
    // packed staging in PINNED host memory (grow-only, owned by the context): the pipeline's copies then run as true
    // asynchronous DMA instead of the driver's staged pageable path
    std::lock_guard<std::mutex> pin_lock(c->pin_mu);
    auto pin_reserve = [](void** p, size_t* cap, size_t need) -> bool {
      if (need <= *cap) return true;
      if (*p) cudaFreeHost(*p);
      *p = nullptr; *cap = 0;
      const size_t want = need + need / 4 + 4096;
      if (cudaHostAlloc(p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
      *cap = want;
      return true;
    };
    if (!pin_reserve(&c->pin_in, &c->pin_in_cap, ia + 1) || !pin_reserve(&c->pin_out, &c->pin_out_cap, oa + 1)) {
      set_error("brotli_b200: pinned staging allocation failed"); return BROTLI_DECODER_ERROR_UNREACHABLE;
    }
    uint8_t* in_blob = (uint8_t*)c->pin_in; uint8_t* out_blob = (uint8_t*)c->pin_out;
    for (size_t i = 0; i < n; i++) if (in_size[i]) memcpy(in_blob + in_off[i], in[i], in_size[i]);
    int rc = decode_host_packed(c, n, in_blob, in_off.data(), out_blob, out_off.data(), out_len.data(), cds.data(), nullptr, 1u);
    if (rc != 0) return rc;
    for (size_t i = 0; i < n; i++) {
      if (out_len[i]) memcpy(out[i], out_blob + out_off[i], out_len[i]);
      out_size[i] = (size_t)out_len[i];
      results[i] = cds[i] == 1 ? BROTLI_DECODER_RESULT_SUCCESS : BROTLI_DECODER_RESULT_ERROR;
      if (codes) codes[i] = (BrotliDecoderErrorCode)cds[i];
    }
    return 0;
  } catch (...) { set_error("brotli_b200: exception"); return BROTLI_DECODER_ERROR_UNREACHABLE; }
}

// ---- custom LZ77 dictionary (BrotliState::new_with_custom_dictionary, src/state.rs:400-411; src/lib.rs:105-131) ----
// Streaming form: the dictionary belongs to the state, as in Decompressor::new_with_custom_dict (src/reader.rs:105) and
// DecompressorWriter::new_with_custom_dictionary (src/writer.rs:117).  Only before the first input byte.
int BrotliB200DecoderSetCustomDictionary(BrotliDecoderState* s, const uint8_t* dictionary, size_t dictionary_size) {
  if (!s || s->sess.used || s->sess.stop != brotli_b200::Session::kFresh || s->sess.d_dict || (dictionary_size && !dictionary)) return 0;
  try { s->sess.dict.assign(dictionary, dictionary + dictionary_size); } catch (...) { return 0; }
  s->sess.large_window = true;  // BrotliState::new_with_custom_dictionary, src/state.rs:400-411
  return 1;
}

BrotliDecoderReturnInfo BrotliB200DecompressWithDictionary(size_t encoded_size, const uint8_t* encoded_buffer, size_t decoded_size,
                                                           uint8_t* decoded_buffer, const uint8_t* dictionary, size_t dictionary_size) {
  try {
    if (!valid_slice(encoded_buffer, encoded_size, 1) || !valid_slice(decoded_buffer, decoded_size, 1) ||
        !valid_slice(dictionary, dictionary_size, 1))
      return invalid_arguments_info();
    return one_shot(encoded_buffer, encoded_size, decoded_buffer, decoded_size, 1u, nullptr, dictionary, dictionary_size);
  } catch (const std::exception& e) {
    return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, e.what());
  } catch (...) {
    return make_info(BROTLI_DECODER_RESULT_ERROR, BROTLI_DECODER_ERROR_UNREACHABLE, 0, "brotli_b200: unknown exception");
  }
}

int BrotliB200DecompressBatchPackedWithDictionary(size_t n, const uint8_t* in_bytes, const uint64_t* in_off, uint8_t* out_bytes,
                                                  const uint64_t* out_off, uint64_t* out_len, int32_t* codes, const uint8_t* dictionary,
                                                  size_t dictionary_size) {
  try {
    if (n == 0) return 0;
    if (!in_off || !out_off || !out_len || !codes || !in_bytes || !out_bytes || (dictionary_size && !dictionary)) {
      set_error("brotli_b200: null batch array"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS;
    }
    DeviceCtx* c = acquire_ctx();
    if (!c) return BROTLI_DECODER_ERROR_UNREACHABLE;
    return decode_host_packed(c, n, in_bytes, in_off, out_bytes, out_off, out_len, codes, nullptr, 1u, dictionary, dictionary_size);
  } catch (...) { set_error("brotli_b200: exception"); return BROTLI_DECODER_ERROR_UNREACHABLE; }
}

int BrotliB200ChecksumBatchDevice(size_t n, const uint8_t* d_bytes, const uint64_t* d_off, const uint64_t* d_len, uint64_t* d_sums,
                                  void* cuda_stream) {
  if (n == 0) return 0;
  if (n > 0xFFFFFFF0ull || !d_off || !d_len || !d_sums) { set_error("brotli_b200: bad checksum arguments"); return BROTLI_DECODER_ERROR_INVALID_ARGUMENTS; }
  if (!acquire_ctx()) return BROTLI_DECODER_ERROR_UNREACHABLE;
  CU_TRY(brotli_b200::launch_checksum_batch((uint32_t)n, d_bytes, d_off, d_len, d_sums, (cudaStream_t)cuda_stream));
  g_launches.fetch_add(1);
  return 0;
}

uint64_t BrotliB200KernelLaunchCount(void) { return g_launches.load(); }
double BrotliB200LastKernelMs(void) { return g_last_kernel_ms.load(); }

int BrotliB200KernelTimes(double* lane_ms, double* exact_ms, uint32_t* launches, uint32_t* bailed, int reset) {
  DeviceCtx* c = acquire_ctx();
  if (!c) return BROTLI_DECODER_ERROR_UNREACHABLE;
  std::lock_guard<std::mutex> lock(c->launch_mu);
  CU_TRY(cudaDeviceSynchronize());
  double lane = 0, exact = 0;
  const uint32_t n = c->timed_count < (uint32_t)DeviceCtx::kTimedLaunches ? c->timed_count : (uint32_t)DeviceCtx::kTimedLaunches;
  for (uint32_t i = 0; i < n; i++) {
    float a = 0, b = 0;
    CU_TRY(cudaEventElapsedTime(&a, c->ev_t[i][0], c->ev_t[i][1]));
    CU_TRY(cudaEventElapsedTime(&b, c->ev_t[i][1], c->ev_t[i][2]));
    lane += a; exact += b;
  }
  if (lane_ms) *lane_ms = lane;
  if (exact_ms) *exact_ms = exact;
  if (launches) *launches = n;
  if (bailed) {
    *bailed = 0;
    if (c->bail_count) CU_TRY(cudaMemcpy(bailed, c->bail_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  if (reset) c->timed_count = 0;
  return 0;
}
const char* BrotliB200LastError(void) { return tl_error.c_str(); }

// Tuning knobs of the current device's context (tests and benchmarks force a path with them).
int BrotliB200SetTuning(const char* name, uint64_t value) {
  DeviceCtx* c = acquire_ctx();
  if (!c || !name) return 0;
  std::lock_guard<std::mutex> lock(c->launch_mu);
  if (strcmp(name, "lane_min_streams") == 0) { c->lane_min_streams = (size_t)value; return 1; }
  if (strcmp(name, "small_geometry") == 0) {
    if (value == 0) { c->small_warps = 0; return 1; }
    c->small_warps = 8;
    c->small_ctas = brotli_b200::query_lane_resident_ctas(c->device, 8);
    c->small_slot_bytes = brotli_b200::lane_slot_bytes(8);
    if (c->small_ctas <= 0) { c->small_warps = 0; return 0; }
    return 1;
  }
  if (strcmp(name, "sort_streams") == 0) { c->sort_streams = value != 0; return 1; }
  if (strcmp(name, "lane_slot_bytes") == 0) {  // table slot of the default geometry: 0 = all the geometry allows; smaller slots for tests
    const uint32_t full = brotli_b200::lane_slot_bytes(c->lane_warps);
    const uint32_t v = (((uint32_t)value / 4u) | 1u) * 4u;  // an odd number of words (see lane_slot_bytes)
    if (value == 0) { c->lane_slot_bytes = full; return 1; }
    if (v < 80u || v > full) return 0;
    c->lane_slot_bytes = v;
    return 1;
  }
  return 0;
}

int BrotliB200LastLaneGeometry(void) {
  DeviceCtx* c = acquire_ctx();
  if (!c) return -1;
  std::lock_guard<std::mutex> lock(c->launch_mu);
  if (c->last_lane_warps >= 0) return c->last_lane_warps;
  uint32_t choice = 0;
  if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(&choice, c->ticket + 24, sizeof(choice), cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  if (choice == 0) return c->lane_warps;
  return choice <= 2 ? c->alt[choice - 1].warps : -1;
}

int BrotliB200ResidentWarps(void) {
  DeviceCtx* c = acquire_ctx();
  return c ? c->ctas * brotli_b200::kWarpsPerCta : -1;
}

void BrotliB200Shutdown(void) {  // (open decoder states keep their own device buffers: destroy them first)
  for (int d = 0; d < kMaxDevices; d++) {
    DeviceCtx* c = &g_ctx[d];
    std::lock_guard<std::mutex> lock(c->mu);
    if (!c->ready) continue;
    int prev = 0; cudaGetDevice(&prev); cudaSetDevice(c->device);
    cudaFree(c->arena); cudaFree(c->dictionary); cudaFree(c->ticket);
    if (c->lane_arena) cudaFree(c->lane_arena);
    if (c->xdict) cudaFree(c->xdict);
    c->xdict = nullptr;
    c->lane_arena = nullptr; c->lane_ctas = 0; c->bail_list.release(); c->order.release();
    c->sess_states.release(); c->sess_pieces.release(); c->sess_blob.release(); c->redo_out.release(); c->cdict.release();
    if (c->sess_pin_up) cudaFreeHost(c->sess_pin_up);
    if (c->sess_pin_down) cudaFreeHost(c->sess_pin_down);
    c->sess_pin_up = c->sess_pin_down = nullptr; c->sess_pin_up_cap = c->sess_pin_down_cap = 0;
    for (void* p : c->sess_slabs) cudaFree(p);
    c->sess_slabs.clear();
    for (auto& fl : c->sess_free) fl.clear();
    c->sess_class.clear();
    c->in.release(); c->out.release(); c->in_off.release(); c->out_off.release(); c->out_len.release(); c->codes.release(); c->in_used.release();
    if (c->pin_in) cudaFreeHost(c->pin_in);
    if (c->pin_out) cudaFreeHost(c->pin_out);
    c->pin_in = c->pin_out = nullptr; c->pin_in_cap = c->pin_out_cap = 0;
    cudaStreamDestroy(c->s_compute); cudaStreamDestroy(c->s_h2d); cudaStreamDestroy(c->s_d2h);
    for (auto& t : c->ev_t) for (auto& e : t) if (e) { cudaEventDestroy(e); e = nullptr; }
    c->timed_count = 0;
    cudaEventDestroy(c->ev_k0); cudaEventDestroy(c->ev_k1); cudaEventDestroy(c->ev_arena); c->arena_busy = false;
    c->ready = false;
    cudaSetDevice(prev);
  }
}

}  // extern "C"
