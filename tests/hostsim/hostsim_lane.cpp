// tests/hostsim/hostsim_lane.cpp -- TEST-ONLY host build of the lane-per-stream path of the CUDA decoder
// (csrc/brotli_decode_lane.cuh).  A lane is scalar code, so the host build runs exactly the device logic
// for one stream.  Returns 1 on success or 1000 when the optimistic path gives the stream up (the exact
// kernel then decodes it on the device).  Not a fallback: never linked into the product.
#define BROTLI_B200_HOSTSIM 1
#include "../../rust-brotli-decompressor_b200/csrc/brotli_decode_lane.cuh"

#include <stdlib.h>
#include <vector>

extern "C" const uint8_t kBrotliDictionaryData[];

extern "C" int hostsim_lane_decode_dict(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, uint32_t table_entries, uint64_t* decoded,
                                        uint64_t* used, const uint8_t* dict, size_t dict_size);
extern "C" int hostsim_lane_decode(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, uint32_t table_entries, uint64_t* decoded,
                                   uint64_t* used) {
  return hostsim_lane_decode_dict(in, in_size, out, cap, table_entries, decoded, used, nullptr, 0);
}
// 0: the throughput configuration of the command loop (what the large geometries run), 1: the latency configuration of the small ones
static int g_latency_config = 0;
extern "C" void hostsim_lane_set_latency_config(int on) { g_latency_config = on; }
extern "C" int hostsim_lane_decode_dict(const uint8_t* in, size_t in_size, uint8_t* out, size_t cap, uint32_t table_entries, uint64_t* decoded,
                                        uint64_t* used, const uint8_t* dict, size_t dict_size) {
  using namespace brotli_b200;
  static std::vector<uint2> cmd_lut;
  if (cmd_lut.empty()) { cmd_lut.resize(704); for (uint32_t i = 0; i < 704; i++) cmd_lut[i] = pack_cmd_lut(i); }
  std::vector<uint64_t> arena((lane::ArenaLayout::kBytes + 7) / 8);
  std::vector<uint64_t> slot((lane::kSlotHeaderBytes + 2 * (size_t)table_entries + 7) / 8 + 1);
  uint8_t* a = (uint8_t*)arena.data();
  lane::LaneCtx c;
  c.slot = hw::to_sref(slot.data());
  c.stab = c.slot + lane::kSlotHeaderBytes;
  c.E = table_entries;
  c.gtab = (uint16_t*)(a + lane::ArenaLayout::kTab);
  c.ctx_lit = a + lane::ArenaLayout::kCtxLit;
  c.ctx_dist = a + lane::ArenaLayout::kCtxDist;
  c.ctx_modes = a + lane::ArenaLayout::kCtxModes;
  // two 16-byte input blocks; the device copies whole aligned blocks, so give the stream padding on both sides
  // while keeping its address modulo 16
  alignas(16) static uint8_t ring[48];  // blocks at +0 and +16 (stride 16) or +0 and +32 (stride 32: the latency instance)
  alignas(16) static uint8_t hist[32];
  alignas(16) static uint8_t stage[64];
  c.hist = hw::to_sref(hist);
  c.stage = hw::to_sref(stage);
  c.stage_c = hw::to_sref(stage + 48);
  c.ring = hw::to_sref(ring);
  c.ring_stride = g_latency_config ? 32 : 16;
  std::vector<uint8_t> padded(in_size + 64 + 16);
  uint8_t* pin = padded.data() + 32;
  pin += (((uintptr_t)in & 15u) - ((uintptr_t)pin & 15u)) & 15u;
  memcpy(pin, in, in_size);
  in = pin;
  c.cmd_lut = hw::to_sref(cmd_lut.data());
  c.ctx_lut = hw::to_sref(tbl::kBrotliContextLookup);
  c.dictionary = kBrotliDictionaryData;
  // expanded dictionary and its two index tables (built by brotli_build_xdict_kernel on the device)
  static std::vector<uint8_t> xdict;
  static uint32_t word_info[25], transform_info[BROTLI_NUM_TRANSFORMS];
  if (xdict.empty()) {
    const lane::XDictLayout x = lane::xdict_layout();
    xdict.resize(x.total + 64);
    for (uint32_t len = 0; len < 25; len++) {
      if (lane::dict_size_bits(len) != (len >= 4 ? tbl::kBrotliDictSizeBitsByLength[len] : 0u)) abort();
      word_info[len] = lane::pack_word_info(x, len);
    }
    for (uint32_t t = 0; t < BROTLI_NUM_TRANSFORMS; t++) transform_info[t] = lane::pack_transform_info(t);
    for (uint32_t len = BROTLI_MIN_DICTIONARY_WORD_LENGTH; len <= BROTLI_MAX_DICTIONARY_WORD_LENGTH; len++)
      for (uint32_t idx = 0; idx < (1u << tbl::kBrotliDictSizeBitsByLength[len]); idx++)
        for (uint32_t t = 0; t < BROTLI_NUM_TRANSFORMS; t++)
          lane::build_xdict_entry(xdict.data() + x.base[len] + (size_t)(idx * BROTLI_NUM_TRANSFORMS + t) * lane::xdict_stride(len),
                                  kBrotliDictionaryData + tbl::kBrotliDictOffsetsByLength[len] + idx * len, len, t);
  }
  c.xdict = xdict.data();
  c.word_info = hw::to_sref(word_info);
  c.transform_info = hw::to_sref(transform_info);
  uint64_t d = 0, u = 0;
  std::vector<uint8_t> dict_padded(dict_size + 64);
  if (dict_size) memcpy(dict_padded.data() + 32, dict, dict_size);
  c.cdict = dict_padded.data() + 32;
  c.cdict_len = dict_size;
  const uint32_t r = g_latency_config ? (dict_size ? lane::decode_streams<32, true>(c, true, in, in_size, out, cap, &d, &u)
                                                   : lane::decode_streams<32, false>(c, true, in, in_size, out, cap, &d, &u))
                                      : (dict_size ? lane::decode_streams<16, true>(c, true, in, in_size, out, cap, &d, &u)
                                                   : lane::decode_streams<16, false>(c, true, in, in_size, out, cap, &d, &u));
  *decoded = d;
  if (used) *used = u;
  return r == lane::kStDone ? 1 : 1000;
}
