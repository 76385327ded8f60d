// brotli_b200_kernels.cu -- sm_100a kernels of the batched Brotli decoder and their launchers.
//
//   brotli_decode_batch_kernel   persistent warps; each warp pulls stream indices from an atomic
//                                ticket and decodes one stream at a time (brotli_decode_core.cuh)
//   brotli_checksum_batch_kernel per-stream 64-bit checksum of decoded regions (parity at scale)
//
// Host code reaches these only through the launch_* functions declared in brotli_b200_runtime.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "brotli_decode_core.cuh"
#include "brotli_b200_runtime.h"

namespace brotli_b200 {

// Shared memory of one CTA: the two read-only LUTs (static) + one WarpShared region with its table
// storage per warp (dynamic, kWarpSharedBytes each).
static_assert(sizeof(WarpShared) % 8 == 0, "table storage must stay 8-byte aligned");
static_assert(kWarpSharedBytes > sizeof(WarpShared) + 2 * 2048, "per-warp region too small for any table");
static_assert(kWarpSharedBytesWide > sizeof(WarpShared) + 2 * 2048, "per-warp region too small for any table");

// Two geometries of the same kernel: WARPS warps per CTA with WSB bytes of shared memory each.  The default (24 x 9216)
// has the faster warps; the wide one (28 x 8000: 4144 resident warps) takes batches whose streams then fit into fewer
// waves -- config C4's 4096 streams of 16 MiB run in one wave instead of two (66 vs 47 GB/s).
template <int WARPS, uint32_t WSB>
__global__ void __launch_bounds__(WARPS * 32, kMinCtasPerSm) brotli_decode_batch_kernel(BatchArgs a) {
  constexpr uint32_t kSharedTableEntries = (WSB - sizeof(WarpShared)) / 2;
  __shared__ uint2 s_cmd_lut[704];
  __shared__ __align__(16) uint8_t s_ctx_lut[2048];
  extern __shared__ __align__(16) uint8_t s_dyn[];
  for (uint32_t i = threadIdx.x; i < 704; i += blockDim.x) s_cmd_lut[i] = pack_cmd_lut(i);
  for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s_ctx_lut[i] = tbl::kBrotliContextLookup[i];
  __syncthreads();

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t gwarp = (uint64_t)blockIdx.x * WARPS + warp;

  Decoder d;
  d.arena = a.arena + gwarp * ArenaLayout::kBytes;
  d.tables = (uint16_t*)(d.arena + ArenaLayout::kTables);
  d.sh = (WarpShared*)(s_dyn + (size_t)warp * WSB);
  d.ws = &d.sh->ws;
  d.stab = (uint16_t*)(d.sh + 1);
  d.stab_cap = kSharedTableEntries;
  d.stab_used = 0;
  d.luts.cmd_lut = s_cmd_lut;
  d.luts.ctx_lut = s_ctx_lut;
  d.luts.dictionary = a.dictionary;

  const uint32_t n = a.n_ptr ? *a.n_ptr : a.n;
  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(a.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= n) break;
    if (a.sessions) {  // streaming sessions: stream t is described by (and reports into) its ResumeState
      decode_session(d, a.sessions + t);
      continue;
    }
    const uint32_t i = a.order ? a.order[t] : t;
    const uint64_t in0 = a.in_off[i], in1 = a.in_off[i + 1];
    const uint64_t out0 = a.out_off[i], out1 = a.out_off[i + 1];
    uint64_t decoded = 0, used = 0;
    const int code = decode_stream(d, a.in + in0, in1 - in0, a.out + out0, out1 - out0, a.large_window, &decoded, &used, a.custom_dict,
                                   a.custom_dict_size, nullptr);
    if (lane == 0) {
      a.out_len[i] = decoded;
      a.codes[i] = code;
      if (a.in_used) a.in_used[i] = used;
    }
  }
}

// Order-independent 64-bit sum of position-keyed byte hashes, one warp per stream: byte j
// contributes mix((b + 1) * (K1 + 2j)) * K2, so any changed, moved or missing byte shows.
// tests/ and bench.py recompute it with numpy over the oracle's output.
__global__ void brotli_checksum_batch_kernel(uint32_t n, const uint8_t* bytes, const uint64_t* off, const uint64_t* len,
                                             uint64_t* sums) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t gwarp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i = gwarp; i < n; i += nwarps) {
    const uint8_t* p = bytes + off[i];
    const uint64_t L = len[i];
    uint64_t acc = 0;
    for (uint64_t j = lane; j < L; j += 32) {
      uint64_t x = (uint64_t)p[j] + 1u;
      x *= 0x9E3779B97F4A7C15ull + 2 * j;  // position-dependent odd multiplier
      x ^= x >> 29;
      acc += x * 0xBF58476D1CE4E5B9ull;
    }
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) sums[i] = acc ^ (L * 0x94D049BB133111EBull);
  }
}

// Moves the pieces of a session launch: fresh input from the launch's staging blob to the sessions' input buffers
// before the decode, new output from the sessions' windows to the staging blob after it.  One CTA per piece.
__global__ void brotli_session_copy_kernel(const SessionCopy* pieces, uint32_t n) {
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const SessionCopy c = pieces[i];
    for (uint64_t j = threadIdx.x; j < c.n; j += blockDim.x) c.dst[j] = c.src[j];
  }
}

cudaError_t launch_session_copy(const SessionCopy* d_pieces, uint32_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  brotli_session_copy_kernel<<<n < 1184 ? n : 1184, 256, 0, stream>>>(d_pieces, n);
  return cudaGetLastError();
}

size_t resume_state_bytes() { return sizeof(ResumeState); }

size_t arena_bytes_per_warp() { return ArenaLayout::kBytes; }

namespace {
constexpr uint32_t kDynamicSharedBytesWide = kWarpsPerCtaWide * kWarpSharedBytesWide;
bool g_wide_ok = false;  // the wide geometry is resident with one CTA per SM as well (checked once per process)
}

int query_resident_ctas(int device) {
  int per_sm = 0, sms = 0;
  auto* k = brotli_decode_batch_kernel<kWarpsPerCta, kWarpSharedBytes>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynamicSharedBytes) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreadsPerCta, kDynamicSharedBytes) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  if (per_sm < 1) per_sm = 1;
  auto* kw = brotli_decode_batch_kernel<kWarpsPerCtaWide, kWarpSharedBytesWide>;
  int per_sm_wide = 0;
  g_wide_ok = cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDynamicSharedBytesWide) == cudaSuccess &&
              cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_wide, kw, kWarpsPerCtaWide * 32, kDynamicSharedBytesWide) == cudaSuccess &&
              per_sm_wide >= 1 && per_sm == 1;
  if (!g_wide_ok) cudaGetLastError();
  return per_sm * sms;
}

cudaError_t launch_decode_batch(const BatchArgs& a, int ctas, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  // a small batch whose size the host knows does not need every resident CTA (launch and drain cost) -- but its streams
  // should still spread over the SMs: at most four per CTA
  if (!a.n_ptr) {
    const int need = (int)((a.n + 3) / 4);
    if (need < ctas) ctas = need < 1 ? 1 : need;
    // Geometry by wave count: a wave of the wide geometry takes ~1.17x as long as one of the default (latency table,
    // profiles/r02), so it is taken when it saves a wave (4144 instead of 3552 streams per wave on 148 SMs).
    static const bool allow_wide = !(getenv("BROTLI_B200_EXACT_WIDE") && getenv("BROTLI_B200_EXACT_WIDE")[0] == '0');
    const uint64_t per_wave = (uint64_t)ctas * kWarpsPerCta, per_wave_wide = (uint64_t)ctas * kWarpsPerCtaWide;
    const uint64_t waves = (a.n + per_wave - 1) / per_wave, waves_wide = (a.n + per_wave_wide - 1) / per_wave_wide;
    if (g_wide_ok && allow_wide && a.n > per_wave && waves_wide * 117 < waves * 100) {
      brotli_decode_batch_kernel<kWarpsPerCtaWide, kWarpSharedBytesWide><<<ctas, kWarpsPerCtaWide * 32, kDynamicSharedBytesWide, stream>>>(a);
      return cudaGetLastError();
    }
  }
  brotli_decode_batch_kernel<kWarpsPerCta, kWarpSharedBytes><<<ctas, kThreadsPerCta, kDynamicSharedBytes, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_checksum_batch(uint32_t n, const uint8_t* bytes, const uint64_t* off, const uint64_t* len, uint64_t* sums,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  uint32_t blocks = (n + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  brotli_checksum_batch_kernel<<<blocks, 256, 0, stream>>>(n, bytes, off, len, sums);
  return cudaGetLastError();
}

}  // namespace brotli_b200
