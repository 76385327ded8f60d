// brotli_b200_lane_kernel.cu -- brotli_decode_lane_kernel: one stream per lane, 32 streams per warp
// (see brotli_decode_lane.cuh).  Streams this optimistic path gives up are appended to a bail list and
// decoded afterwards by the exact warp-per-stream kernel (brotli_b200_kernels.cu).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <cub/device/device_radix_sort.cuh>

#include "brotli_b200_runtime.h"
#include "brotli_decode_lane.cuh"

namespace brotli_b200 {

// One persistent CTA per SM.  Static shared memory: the command LUT and the literal-context LUT, read by
// all lanes.  Dynamic shared memory: one private slot per lane (slot header + root tables).
// DICT: the batch carries a custom LZ77 dictionary (its own instance: batches without one run the code they always ran).
template <int WARPS, bool DICT = false>
__global__ void __launch_bounds__(WARPS * 32, 1) brotli_decode_lane_kernel(BatchArgs a, LaneArgs la) {
  uint2* const s_cmd_lut = lane::g_cmd_lut;
  uint8_t* const s_ctx_lut = lane::g_ctx_lut;
  uint32_t* const s_word_info = lane::g_word_info;
  uint32_t* const s_transform_info = lane::g_transform_info;
  extern __shared__ __align__(16) uint8_t s_dyn[];
  if (la.geom_choice && *la.geom_choice != la.geom_id) return;  // another geometry's launch decodes this batch
  for (uint32_t i = threadIdx.x; i < 704; i += blockDim.x) s_cmd_lut[i] = pack_cmd_lut(i);
  for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s_ctx_lut[i] = tbl::kBrotliContextLookup[i];
  {
    const lane::XDictLayout x = lane::xdict_layout();
    for (uint32_t i = threadIdx.x; i < 25; i += blockDim.x) s_word_info[i] = lane::pack_word_info(x, i);
    for (uint32_t i = threadIdx.x; i < BROTLI_NUM_TRANSFORMS; i += blockDim.x) s_transform_info[i] = lane::pack_transform_info(i);
  }
  __syncthreads();

  const uint32_t lane_id = threadIdx.x & 31u;
  const uint64_t glane = (uint64_t)blockIdx.x * (WARPS * 32) + threadIdx.x;
  uint8_t* const arena = la.arena + glane * lane::ArenaLayout::kBytes;

  lane::LaneCtx c;
  // dynamic shared memory: the input rings of all lanes (two 16-byte blocks each, block-interleaved so that
  // lanes spread over the banks), then one table slot per lane
  c.ring = hw::to_sref(s_dyn + (size_t)threadIdx.x * 16);
  c.ring_stride = WARPS * 32 * 16;
  c.hist = hw::to_sref(s_dyn + (size_t)WARPS * 32 * 32 + (size_t)threadIdx.x * 32);  // 32-byte output history ring
  // cp.async landing zones: 48 bytes per lane (copy source, next phase-A entry) at a stride that spreads a warp's 16-byte
  // writes over all banks, then the 16-byte blocks of the next phase-C entry
  c.stage = hw::to_sref(s_dyn + (size_t)WARPS * 32 * 64 + (size_t)threadIdx.x * 48);
  c.stage_c = hw::to_sref(s_dyn + (size_t)WARPS * 32 * 112 + (size_t)threadIdx.x * 16);
  c.slot = hw::to_sref(s_dyn + (size_t)WARPS * 32 * 128 + (size_t)threadIdx.x * la.slot_bytes);
  c.stab = c.slot + lane::kSlotHeaderBytes;
  c.E = (la.slot_bytes - lane::kSlotHeaderBytes) / 2;
  c.gtab = (uint16_t*)(arena + lane::ArenaLayout::kTab);
  c.ctx_lit = arena + lane::ArenaLayout::kCtxLit;
  c.ctx_dist = arena + lane::ArenaLayout::kCtxDist;
  c.ctx_modes = arena + lane::ArenaLayout::kCtxModes;
  c.cmd_lut = hw::to_sref(s_cmd_lut);
  c.ctx_lut = hw::to_sref(s_ctx_lut);
  c.dictionary = a.dictionary;
  c.xdict = la.xdict;
  c.word_info = hw::to_sref(s_word_info);
  c.transform_info = hw::to_sref(s_transform_info);
  c.cdict = DICT ? la.cdict : nullptr;
  c.cdict_len = DICT ? la.cdict_len : 0;

  const uint32_t chunk = la.chunk;  // streams a warp takes per ticket (32 unless the batch is small)
  for (;;) {
    uint32_t base = 0;
    if (lane_id == 0) base = atomicAdd(a.ticket, chunk);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= a.n) break;
    const uint32_t t = base + lane_id;
    const bool active = lane_id < chunk && t < a.n;
    uint32_t i = 0;
    uint64_t in0 = 0, in1 = 0, out0 = 0, out1 = 0;
    if (active) {
      i = a.order ? a.order[t] : t;
      in0 = a.in_off[i]; in1 = a.in_off[i + 1];
      out0 = a.out_off[i]; out1 = a.out_off[i + 1];
    }
    uint64_t decoded = 0, used = 0;
    // the whole warp decodes together: one stream per lane, one prefix-code symbol per lane and iteration
    const uint32_t r = lane::decode_streams<WARPS * 32 * 16, DICT>(c, active, a.in + in0, in1 - in0, a.out + out0, out1 - out0, &decoded, &used);
    if (active) {
      if (r == lane::kStDone) {
        a.out_len[i] = decoded;
        a.codes[i] = kSuccess;
        if (a.in_used) a.in_used[i] = used;
      } else {
        la.bail_list[atomicAdd(la.bail_count, 1u)] = i;
      }
    }
    __syncwarp();
  }
}

// One thread per (word, transform): fills the expanded dictionary (see xdict_layout) once per device.
__global__ void brotli_build_xdict_kernel(const uint8_t* dictionary, uint8_t* xdict) {
  const lane::XDictLayout x = lane::xdict_layout();
  const uint32_t len = BROTLI_MIN_DICTIONARY_WORD_LENGTH + blockIdx.y;
  const uint32_t n = BROTLI_NUM_TRANSFORMS << lane::dict_size_bits(len);
  for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const uint32_t idx = e / BROTLI_NUM_TRANSFORMS, t = e % BROTLI_NUM_TRANSFORMS;
    lane::build_xdict_entry(xdict + x.base[len] + (size_t)e * lane::xdict_stride(len),
                            dictionary + tbl::kBrotliDictOffsetsByLength[len] + idx * len, len, t);
  }
}

#if BD_LANE_WAIT_HIST
// MEASUREMENT ONLY: adds up the per-warp wait histograms (4 sites x 64 buckets, then 4 cycle sums at [256..259]) into out[264] and clears them.
extern "C" __attribute__((visibility("default"))) int BrotliB200ProbeWaitHist(unsigned long long* out) {
  const size_t n = (size_t)lane::kWaitHistWarps * lane::kWaitHistRow;
  unsigned long long* h = (unsigned long long*)calloc(n, sizeof(unsigned long long));
  if (!h || cudaDeviceSynchronize() != cudaSuccess) return 0;
  if (cudaMemcpyFromSymbol(h, lane::g_wait_hist_all, n * sizeof(unsigned long long)) != cudaSuccess) return 0;
  for (uint32_t i = 0; i < lane::kWaitHistRow; i++) out[i] = 0;
  for (size_t w = 0; w < lane::kWaitHistWarps; w++) for (uint32_t i = 0; i < lane::kWaitHistRow; i++) out[i] += h[w * lane::kWaitHistRow + i];
  memset(h, 0, n * sizeof(unsigned long long));
  const int ok = cudaMemcpyToSymbol(lane::g_wait_hist_all, h, n * sizeof(unsigned long long)) == cudaSuccess;
  free(h);
  return ok;
}
#endif

size_t xdict_bytes() { return (size_t)lane::xdict_layout().total + 64; }

cudaError_t launch_build_xdict(const uint8_t* dictionary, uint8_t* xdict, cudaStream_t stream) {
  brotli_build_xdict_kernel<<<dim3(64, BROTLI_MAX_DICTIONARY_WORD_LENGTH - BROTLI_MIN_DICTIONARY_WORD_LENGTH + 1), 256, 0, stream>>>(dictionary, xdict);
  return cudaGetLastError();
}

// ---- longest-first order ----------------------------------------------------------------------------
// The rounds a warp needs are the maximum over its 32 streams, so streams that take similar work should share
// a warp, and the longest should start first.  Compressed size predicts the work well (correlation 0.9 on
// text, and it separates the stream families of a mixed batch): streams are ordered by descending
// compressed size in 256-byte buckets (stable, so equal buckets keep batch order).
__global__ void brotli_order_keys_kernel(uint32_t n, const uint64_t* in_off, uint32_t* keys, uint32_t* vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint64_t sz = (in_off[i + 1] - in_off[i]) >> 8;
    keys[i] = sz > 0xFFFFFFu ? 0xFFFFFFu : (uint32_t)sz;
    vals[i] = i;
  }
}

size_t order_temp_bytes(uint32_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                            (uint32_t*)nullptr, (int)n, 0, 24);
  return bytes;
}

// scratch: 4 * n uint32 (keys in/out, values in/out) followed by order_temp_bytes(n); the order lands in scratch + 3 * n
cudaError_t launch_order_by_size(uint32_t n, const uint64_t* in_off, uint32_t* scratch, size_t temp_bytes, cudaStream_t stream) {
  uint32_t* keys_a = scratch, *keys_b = scratch + n, *vals_a = scratch + 2 * (size_t)n, *vals_b = scratch + 3 * (size_t)n;
  brotli_order_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, in_off, keys_a, vals_a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  return cub::DeviceRadixSort::SortPairsDescending((void*)(scratch + 4 * (size_t)n), temp_bytes, keys_a, keys_b, vals_a, vals_b, (int)n, 0, 24, stream);
}

// Geometry by wave fit (see brotli_b200_runtime.h).  Full waves cost the same per stream at 14..24 warps per SM (the
// kernel is bound by the memory system), a partial last wave costs at least a stream's latency -- about half a wave of
// the default geometry -- however few streams it holds.  Cost in lane-slots: full waves + max(that floor, 0.8 x the
// rest); a batch that fits one wave of an alternative takes the first such; the default keeps a 7 % bonus (the model
// is good to a few per cent: profiles/r02/waves.txt, fit.txt).
__global__ void brotli_lane_geometry_kernel(uint32_t n, const uint32_t* sorted_keys, const uint32_t* lanes, uint32_t n_geom, uint32_t* choice) {
  uint32_t pick = 0;
  const uint32_t kmax = sorted_keys[0], kmin = sorted_keys[n - 1];
  if (2 * kmin >= kmax && kmin != 0) {
    auto cost = [](uint64_t streams, uint64_t l) {
      const uint64_t full = streams / l, rest = streams - full * l;
      const uint64_t floor_slots = 51000;
      const uint64_t tail = rest == 0 ? 0 : (rest * 4 / 5 > floor_slots ? rest * 4 / 5 : floor_slots);
      return full * l + tail;
    };
    uint64_t best = cost(n, lanes[0]) * 93 / 100;
    for (uint32_t k = 1; k < n_geom; k++) {
      if (lanes[k] == 0) continue;
      const uint64_t ck = n <= lanes[k] ? 0 : cost(n, lanes[k]);
      if (ck < best) { best = ck; pick = k; }
    }
  }
  *choice = pick;
}

cudaError_t launch_choose_lane_geometry(uint32_t n, const uint32_t* sorted_keys, const uint32_t* lanes, uint32_t n_geom, uint32_t* choice, cudaStream_t stream) {
  brotli_lane_geometry_kernel<<<1, 1, 0, stream>>>(n, sorted_keys, lanes, n_geom, choice);
  return cudaGetLastError();
}

namespace {
template <int WARPS>
int lane_occupancy(uint32_t dyn) {
  int per_sm = 0;
  if (cudaFuncSetAttribute(brotli_decode_lane_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brotli_decode_lane_kernel<WARPS>, WARPS * 32, dyn) != cudaSuccess) return -1;
  return per_sm;
}
}  // namespace

size_t lane_arena_bytes_per_lane() { return lane::ArenaLayout::kBytes; }

// Slot size for `warps` warps per CTA: what is left of the SM's shared memory, an odd number of 32-bit
// words so that equal offsets in different lanes' slots fall into different banks.
uint32_t lane_slot_bytes(int warps) {
  // static shared memory of the kernel: the command LUT (5632), the context LUT (2048), word / transform info (584)
  const uint32_t static_bytes = 704 * 8 + 2048 + 4 * (25 + BROTLI_NUM_TRANSFORMS) + 64;
  uint32_t per_lane = (232448u - static_bytes) / (uint32_t)(warps * 32) - 128u;  // 128: the lane's input ring (32), output history ring (32) and cp.async landing zone (64)
  uint32_t words = per_lane / 4;
  if ((words & 1u) == 0) words--;
  return words * 4;
}

int query_lane_resident_ctas(int device, int warps) {
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  const uint32_t dyn = (lane_slot_bytes(warps) + 128u) * (uint32_t)(warps * 32);
  int per_sm = -1;
  switch (warps) {
    case 4: per_sm = lane_occupancy<4>(dyn); break;
    case 8: per_sm = lane_occupancy<8>(dyn); break;
    case 12: per_sm = lane_occupancy<12>(dyn); break;
    case 14: per_sm = lane_occupancy<14>(dyn); break;
    case 16: per_sm = lane_occupancy<16>(dyn); break;
    case 20: per_sm = lane_occupancy<20>(dyn); break;
    case 24: per_sm = lane_occupancy<24>(dyn); break;
    case 28: per_sm = lane_occupancy<28>(dyn); break;
    case 32: per_sm = lane_occupancy<32>(dyn); break;
    default: return -1;
  }
  if (per_sm < 1) return -1;
  return per_sm * sms;
}

// the dictionary instance exists for the default geometry only
bool lane_kernel_takes_dictionary(int warps) { return warps == kLaneWarpsPerCta; }

cudaError_t launch_decode_lane(const BatchArgs& a, const LaneArgs& la, int ctas, int warps, cudaStream_t stream, bool reset_counters) {
  if (la.cdict_len != 0 && !lane_kernel_takes_dictionary(warps)) return cudaErrorInvalidValue;
  cudaError_t e = cudaSuccess;
  if (reset_counters) {
    e = cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(la.bail_count, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
  }
  const uint32_t dyn = (la.slot_bytes + 128u) * (uint32_t)(warps * 32);
  if (la.cdict_len != 0) {  // (the default geometry: checked above)
    if (cudaFuncSetAttribute(brotli_decode_lane_kernel<kLaneWarpsPerCta, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) return cudaGetLastError();
    brotli_decode_lane_kernel<kLaneWarpsPerCta, true><<<ctas, kLaneWarpsPerCta * 32, dyn, stream>>>(a, la);
    return cudaGetLastError();
  }
  switch (warps) {
    case 4: brotli_decode_lane_kernel<4><<<ctas, 128, dyn, stream>>>(a, la); break;
    case 8: brotli_decode_lane_kernel<8><<<ctas, 256, dyn, stream>>>(a, la); break;
    case 12: brotli_decode_lane_kernel<12><<<ctas, 384, dyn, stream>>>(a, la); break;
    case 14: brotli_decode_lane_kernel<14><<<ctas, 448, dyn, stream>>>(a, la); break;
    case 16: brotli_decode_lane_kernel<16><<<ctas, 512, dyn, stream>>>(a, la); break;
    case 20: brotli_decode_lane_kernel<20><<<ctas, 640, dyn, stream>>>(a, la); break;
    case 24: brotli_decode_lane_kernel<24><<<ctas, 768, dyn, stream>>>(a, la); break;
    case 28: brotli_decode_lane_kernel<28><<<ctas, 896, dyn, stream>>>(a, la); break;
    case 32: brotli_decode_lane_kernel<32><<<ctas, 1024, dyn, stream>>>(a, la); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace brotli_b200
