// tests/hostsim/hostsim.cpp -- TEST-ONLY host build of the CUDA decoder's logic.
//
// The dev container has no GPU, so the device source (csrc/brotli_decode_core.cuh) is also
// compiled here for the host with a warp width of 1 (BROTLI_B200_HOSTSIM).  CPU tests use it to
// check the kernel's control flow against the oracle before spending GPU time.  It is NOT a
// fallback: nothing in the product library or its C ABI can reach this file.
#define BROTLI_B200_HOSTSIM 1
#include "../../rust-brotli-decompressor_b200/csrc/brotli_decode_core.cuh"

#include <stdlib.h>
#include <vector>

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
extern "C" int hostsim_decode_resume(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, void* resume_state,
                                     uint64_t* decoded);
extern "C" size_t hostsim_resume_state_bytes() { return sizeof(brotli_b200::ResumeState); }
static void* g_resume = nullptr;
static const uint8_t* g_resume_dict = nullptr;
static size_t g_resume_dict_size = 0;
extern "C" void hostsim_session_dictionary(const uint8_t* dict, size_t dict_size) { g_resume_dict = dict; g_resume_dict_size = dict_size; }
extern "C" int hostsim_decode_resume(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, void* resume_state,
                                     uint64_t* decoded) {
  g_resume = resume_state;
  int rc = hostsim_decode_dict(in, in_size, out, cap, large_window, g_resume_dict, g_resume_dict_size, decoded);
  g_resume = nullptr;
  return rc;
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
  int rc = decode_stream(d, in, in_size, out, cap, (uint32_t)large_window, decoded, &used, dict, dict_size, (ResumeState*)g_resume);
  g_last_stab_used = d.stab_used; g_last_all_shared = d.all_shared; g_last_unpromoted = d.n_unpromoted;
  return rc;
}
