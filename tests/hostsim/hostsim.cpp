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

extern "C" int hostsim_decode(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, int large_window, uint64_t* decoded) {
  using namespace brotli_b200;
  static std::vector<uint2> cmd_lut;
  if (cmd_lut.empty()) { cmd_lut.resize(704); for (uint32_t i = 0; i < 704; i++) cmd_lut[i] = pack_cmd_lut(i); }
  std::vector<uint8_t> arena(ArenaLayout::kBytes);
  WarpScratch ws;
  Decoder d;
  memset(&d, 0, sizeof(d));
  d.arena = arena.data();
  d.tables = (uint16_t*)(arena.data() + ArenaLayout::kTables);
  d.ws = &ws;
  d.luts.cmd_lut = cmd_lut.data();
  d.luts.ctx_lut = tbl::kBrotliContextLookup;
  d.luts.dictionary = kBrotliDictionaryData;
  uint64_t used = 0;
  return decode_stream(d, in, in_size, out, cap, (uint32_t)large_window, decoded, &used);
}
