// tests/hostsim/hostsim.cpp -- TEST-ONLY host build of the CUDA decoder's logic.
//
// The dev container has no GPU, so the device source (csrc/brotli_decode_core.cuh) is also
// compiled here for the host with a warp width of 1 (BROTLI_B200_HOSTSIM).  CPU tests use it to
// check the kernel's control flow against the oracle before spending GPU time.  It is NOT a
// fallback: nothing in the product library or its C ABI can reach this file.
#define BROTLI_B200_HOSTSIM 1
#include "../../rust-brotli-decompressor_b200/csrc/brotli_decode_core.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../rust-brotli-decompressor_b200/csrc/brotli_b200_session.h"

extern "C" const uint8_t kBrotliDictionaryData[];

static uint32_t g_last_stab_used, g_last_all_shared, g_last_unpromoted;
// shared-table occupancy of the last metablock decoded (sizing experiments for kWarpSharedBytes)
extern "C" void hostsim_last_table_stats(uint32_t* stab_used, uint32_t* all_shared, uint32_t* unpromoted) {
  *stab_used = g_last_stab_used; *all_shared = g_last_all_shared; *unpromoted = g_last_unpromoted;
}

extern "C" int hostsim_decode_dict(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, const uint8_t* dict,
                                   size_t dict_size, uint64_t* decoded);
extern "C" int hostsim_decode(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, uint64_t* decoded) {
  return hostsim_decode_dict(in, in_size, out, cap, large_window, nullptr, 0, decoded);
}
extern "C" int hostsim_decode_dict(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, const uint8_t* dict,
                                   size_t dict_size, uint64_t* decoded) {
  using namespace brotli_b200;
  static std::vector<uint2> cmd_lut;
  if (cmd_lut.empty()) { cmd_lut.resize(704); for (uint32_t i = 0; i < 704; i++) cmd_lut[i] = pack_cmd_lut(i); }
  std::vector<uint8_t> arena(ArenaLayout::kBytes);
  // "shared memory" of the simulated warp: WarpShared followed by table storage (same carve-up as the kernel)
  const uint32_t stab_entries = 4096;
  std::vector<uint64_t> shmem((sizeof(WarpShared) + 2 * stab_entries + 7) / 8);
  WarpShared* sh = (WarpShared*)shmem.data();
  WarpScratch& ws = sh->ws;
  Decoder d;
  memset(&d, 0, sizeof(d));
  d.sh = sh;
  d.stab = (uint16_t*)(sh + 1);
  d.stab_cap = stab_entries;
  d.arena = arena.data();
  d.tables = (uint16_t*)(arena.data() + ArenaLayout::kTables);
  d.ws = &ws;
  d.luts.cmd_lut = cmd_lut.data();
  d.luts.ctx_lut = tbl::kBrotliContextLookup;
  d.luts.dictionary = kBrotliDictionaryData;
  uint64_t used = 0;
  int rc = decode_stream(d, in, in_size, out, cap, (uint32_t)large_window, decoded, &used, dict, dict_size, nullptr);
  g_last_stab_used = d.stab_used; g_last_all_shared = d.all_shared; g_last_unpromoted = d.n_unpromoted;
  return rc;
}

// ---- streaming sessions: the product's session logic (csrc/brotli_b200_session.h) over a host "device" ----
namespace {
using namespace brotli_b200;
struct HostDev {
  size_t live_bytes = 0, peak_bytes = 0; uint64_t launches = 0;
  std::vector<uint2> cmd_lut;
  std::vector<uint64_t> shmem;
  std::vector<uint8_t> own_arena;
  HostDev() {
    cmd_lut.resize(704); for (uint32_t i = 0; i < 704; i++) cmd_lut[i] = pack_cmd_lut(i);
    shmem.resize((sizeof(WarpShared) + 2 * 4096 + 7) / 8);
    own_arena.resize(ArenaLayout::kBytes);
  }
  uint8_t* alloc(size_t n) {
    uint8_t* p = (uint8_t*)malloc(n + 16);
    if (!p) return nullptr;
    *(size_t*)p = n; live_bytes += n; if (live_bytes > peak_bytes) peak_bytes = live_bytes;
    memset(p + 16, 0xA5, n);  // device memory is not zeroed either
    return p + 16;
  }
  void release(uint8_t* p) { if (!p) return; live_bytes -= *(size_t*)(p - 16); free(p - 16); }
  int upload(uint8_t* d, const uint8_t* h, size_t n) { memcpy(d, h, n); return 0; }
  size_t arena_bytes() const { return ArenaLayout::kBytes; }
  int run(ResumeState* arr, uint32_t n, const uint8_t* blob, size_t, const SessionCopy* scatter, uint32_t n_scatter) {
    for (uint32_t i = 0; i < n_scatter; i++) memcpy(scatter[i].dst, blob + (uintptr_t)scatter[i].src, scatter[i].n);
    launches++;
    for (uint32_t i = 0; i < n; i++) {
      WarpShared* sh = (WarpShared*)shmem.data();
      Decoder d;
      memset(&d, 0, sizeof(d));
      d.sh = sh; d.stab = (uint16_t*)(sh + 1); d.stab_cap = 4096; d.ws = &sh->ws;
      d.arena = own_arena.data(); d.tables = (uint16_t*)(d.arena + ArenaLayout::kTables);
      d.luts.cmd_lut = cmd_lut.data(); d.luts.ctx_lut = tbl::kBrotliContextLookup; d.luts.dictionary = kBrotliDictionaryData;
      memset(shmem.data(), 0xEE, shmem.size() * 8);  // shared memory does not survive a launch
      decode_session(d, &arr[i]);
      if (getenv("HOSTSIM_TRACE")) fprintf(stderr, "launch: in_size %llu budget %llu cap %llu -> code %d decoded %llu used %llu at_flush %u hit_cap %u | ck kind %u bitpos %llu pos %u mlen %d state %u ins_rem %u bl %u %u %u\n",
        (unsigned long long)arr[i].in_size, (unsigned long long)arr[i].budget, (unsigned long long)arr[i].out_cap, arr[i].code, (unsigned long long)arr[i].decoded,
        (unsigned long long)arr[i].used, arr[i].at_flush, arr[i].hit_cap, arr[i].kind, (unsigned long long)arr[i].bitpos, arr[i].pos, arr[i].mlen, arr[i].state, arr[i].ins_rem,
        arr[i].bl_l, arr[i].bl_c, arr[i].bl_d);
    }
    return 0;
  }
  std::vector<uint8_t> up, down;
  uint8_t* host_up(size_t bytes) { if (up.size() < bytes) up.resize(bytes); return up.data(); }
  const uint8_t* gather(const SessionCopy* p, uint32_t n, size_t bytes) {
    if (down.size() < bytes + 1) down.resize(bytes + 1);
    for (uint32_t i = 0; i < n; i++) memcpy(down.data() + (uintptr_t)p[i].dst, p[i].src, p[i].n);
    return down.data();
  }
  int move(const SessionCopy* p, uint32_t n) { for (uint32_t i = 0; i < n; i++) memmove(p[i].dst, p[i].src, p[i].n); return 0; }
};
HostDev g_dev;
}  // namespace

extern "C" void* hostsim_stream_create(int large_window, const uint8_t* dict, size_t dict_size) {
  Session* s = new Session();
  s->large_window = large_window != 0;
  if (dict_size) s->dict.assign(dict, dict + dict_size);
  return s;
}
extern "C" void hostsim_stream_destroy(void* h) {
  Session* s = (Session*)h;
  SessionRunner<HostDev> r(g_dev);
  r.destroy(*s);
  delete s;
}
// n concurrent BrotliDecoderDecompressStream calls (one launch); per call: in/out buffers and the moved counters
extern "C" int hostsim_stream_calls(size_t n, void** handles, const uint8_t** in, size_t* in_size, uint8_t** out, size_t* out_cap,
                                    size_t* consumed, size_t* produced, size_t* total_out, int* results, int* codes) {
  std::vector<Session*> ss(n);
  std::vector<StreamCall> cs(n);
  std::vector<const uint8_t*> next_in(n);
  std::vector<uint8_t*> next_out(n);
  std::vector<size_t> avail_in(n), avail_out(n);
  for (size_t i = 0; i < n; i++) {
    ss[i] = (Session*)handles[i];
    next_in[i] = in[i]; next_out[i] = out[i]; avail_in[i] = in_size[i]; avail_out[i] = out_cap[i];
    cs[i] = StreamCall{&avail_in[i], &next_in[i], &avail_out[i], &next_out[i], &total_out[i], -1};
  }
  SessionRunner<HostDev> r(g_dev);
  const int rc = r.stream_calls(ss.data(), cs.data(), n, -31);
  for (size_t i = 0; i < n; i++) {
    consumed[i] = in_size[i] - avail_in[i]; produced[i] = out_cap[i] - avail_out[i];
    results[i] = cs[i].result; codes[i] = ss[i]->code;
  }
  return rc;
}
extern "C" void hostsim_stream_stats(size_t* live_bytes, size_t* peak_bytes, uint64_t* launches) {
  *live_bytes = g_dev.live_bytes; *peak_bytes = g_dev.peak_bytes; *launches = g_dev.launches;
}
