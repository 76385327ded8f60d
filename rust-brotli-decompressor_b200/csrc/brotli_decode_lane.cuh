// brotli_decode_lane.cuh -- one stream per LANE: 32 independent streams per warp (sm_100a).
//
// The warp-per-stream decoder (brotli_decode_core.cuh) issues the serial prefix-code chain once per
// warp, so a B200 is limited by instruction issue long before HBM.  A batch of many independent
// streams has a better mapping: every lane owns a whole stream, so one issued instruction advances 32
// bit windows / table lookups / output cursors at once.  What makes that work on the SM:
//
//   * phased rounds -- the warp runs rounds of fixed phases: [A: command or literal symbol] [P: retire the copy
//     requested last round] [C: distance symbol, copy set-up, request for the copy's source bytes].  Each
//     phase's code is issued once per round for all lanes that are at that point of their stream, so a
//     lane normally completes one whole insert-and-copy command per round; warp votes at the phase
//     boundaries keep the 32 streams converged.
//   * software pipelining with cp.async: what a phase needs from L2/HBM is requested a phase earlier --
//     backreference (and dictionary) source bytes at the end of a round for phase P of the next, the
//     second-level table entry of the next command/literal symbol right after the distance symbol is read,
//     that of the distance symbol right after the command is read -- and lands in per-lane shared staging.
//   * per-lane prefix-code tables with a SMALL root level in shared memory (a private slot per lane;
//     32 random addresses over 32 banks cost ~3 wavefronts) and the second level in a per-lane global
//     arena.  Root widths are chosen per metablock so all roots fit the slot.
//   * output through a 4-byte write combiner: literals and copies are appended to a partial word in a
//     register and leave as aligned 32-bit stores; copy sources are read as aligned words and re-aligned
//     with funnel shifts.  Loads after stores of the SAME thread are ordered by the hardware, so no
//     synchronisation is needed for the data.
//
// This path is OPTIMISTIC (like a fast path in front of the reference's "safe" path): it decodes
// well-formed streams with compressed metablocks and a regular window.  Anything else -- corrupt or
// truncated input, too small an output region, uncompressed / metadata metablocks, large-window
// headers, table sets larger than the arena -- makes the lane give its stream up ("bail"); the host
// runs exactly those streams through brotli_decode_batch_kernel, which owns the reference's
// error-code and decoded_size semantics.  Every validity check of the reference is still made here;
// a failed check bails instead of reporting a code.
//
// Reference path restated (file:line under /root/reference):
//   DecodeWindowBits src/decode.rs:152-187, DecodeMetaBlockLength :243-372, DecodeVarLenUint8 :193-241,
//   ReadHuffmanCode :868-1013 (+ :516-556, :565-658, :661-731, :801-853), table shape after
//   src/huffman/mod.rs:273-471 (canonical codes, bit-reversed keys; the root width is ours),
//   DecodeContextMap :1272-1428, InverseMoveToFrontTransform :1096-1128, DecodeBlockTypeAndLength
//   :1469-1524, PrepareLiteralDecoding :1554-1570, ReadCommandInternal :2134-2189, ReadDistanceInternal
//   :2066-2131, TakeDistanceFromRingBuffer :2017-2049, ProcessCommandsInternal :2330-2744, dictionary
//   :2593-2640 + src/transform.rs:720-795, METABLOCK_DONE :3345-3381.
#pragma once
#include "brotli_decode_core.cuh"

#ifndef BD_NS
#define BD_NS brotli_b200
#endif
namespace BD_NS {
namespace lane {

// ---- per-lane storage geometry ----
// shared slot: cur_dist[4] (root of the distance tree per distance context), then E u16 table entries; in
// metablocks with context-modelled literals the top 32 entries hold the 64-entry context map of the current
// literal block type instead
constexpr uint32_t kSlotHeaderBytes = 16;
constexpr uint32_t kCtxMapEntries = 32;      // table entries the 64-byte context map takes from the top of the slot when needed
constexpr uint32_t kGlobalTab = 16384;      // u16 entries of virtual table space behind the shared slot (12-bit pointers x 4)
constexpr uint32_t kMaxBlockTypes = 64;     // per category handled here (more: bail)
constexpr uint32_t kBlockRootBits = 6;
#ifndef BD_LANE_EXTRA_LITERALS
#define BD_LANE_EXTRA_LITERALS 1
#endif
constexpr uint32_t kMaxExtraLiterals = BD_LANE_EXTRA_LITERALS;  // literals a lane may add to its phase-A literal per round
// (measured and rejected: 2 takes 14 % of the rounds off a text stream, but the extra code is issued in every round of
// every warp -- headline +-0, C3 -2 %, C5 -7 %: the kernel is bound by instructions per round, not by rounds)
#ifndef BD_LANE_CMD_LITERALS
#define BD_LANE_CMD_LITERALS 0
#endif
constexpr uint32_t kCmdLiterals = BD_LANE_CMD_LITERALS;  // literals a lane may decode in the round of their command
// Latency configuration: the small geometries (at most 12 warps per SM) only run batches below one wave of lanes, where
// the batch takes as long as one stream takes on its lane (rounds x the round's dependent chain) and issue slots are idle:
// there, more work per round is what counts, not fewer instructions per round -- more literals per literal round, and a
// command's first literals in the command's own round.
#ifndef BD_LANE_EXTRA_LITERALS_LAT
#define BD_LANE_EXTRA_LITERALS_LAT 4
#endif
#ifndef BD_LANE_CMD_LITERALS_LAT
#define BD_LANE_CMD_LITERALS_LAT 3
#endif
#ifndef BD_LANE_BURST_LAT
#define BD_LANE_BURST_LAT 0   /* literal bursts (see kBurst) in the latency configuration */
#endif
#ifndef BD_LANE_BURST_LANES_LAT
#define BD_LANE_BURST_LANES_LAT 6
#endif
#if defined(BROTLI_B200_HOSTSIM)
#define BD_LANE_IS_LATENCY_STRIDE(s) ((s) == 32u)  /* the host build's second instance (tests/hostsim) */
#else
#define BD_LANE_IS_LATENCY_STRIDE(s) ((s) <= 12u * 32u * 16u)
#endif
// Literal bursts: while at least kBurstLanes lanes of the warp are inside a literal run, up to kBurst extra literal-only
// iterations follow phase A (table look-ups synchronous, nothing else of the round issued).  A round costs ~650
// instructions whatever its lanes do; a burst iteration ~60, so literal-heavy streams (binary data, 4 KiB responses)
// need far fewer rounds.
#ifndef BD_LANE_BURST
#define BD_LANE_BURST 0
#endif
// narrowest root a group may get in the shared slot, and the weights of the three symbol kinds in the root-width allocation
// root widths: all combinations instead of the greedy widening
#ifndef BD_LANE_ROOTS_EXHAUSTIVE
#define BD_LANE_ROOTS_EXHAUSTIVE 0
#endif
// only the current block type's distance trees in the shared slot (see refresh_cur_dist)
#ifndef BD_LANE_DIST_CACHE
#define BD_LANE_DIST_CACHE 1
#endif
// cache slots of the distance-tree cache: as many as a block type really uses (1), or always four (0: the first version)
#ifndef BD_LANE_DIST_SLOTS_EXACT
#define BD_LANE_DIST_SLOTS_EXACT 1
#endif
// MEASUREMENT ONLY (output is wrong): every backreference reads its source from the lane's most recent output lines, which
// are still in L2 -- what the kernel would run at if copy sources cost no DRAM transaction
#ifndef BD_LANE_PROBE_NEAR
#define BD_LANE_PROBE_NEAR 0
#endif
// Phase A's literal words leave for global memory in phase P, behind the wait for the copy chunk: a global store issued
// while the chunk's cp.async is still in flight holds the warp's memory pipeline until the chunk has landed (measured:
// dropping either the stores or the chunk requests takes 24 % off the kernel, dropping both 28 %, and where the chunk
// reads from does not matter), so phase A ran behind the chunk's latency instead of beside it.
#ifndef BD_LANE_DEFER_A_STORES
#define BD_LANE_DEFER_A_STORES 1
#endif
#ifndef BD_LANE_WAIT_ALL_AT_P
#define BD_LANE_WAIT_ALL_AT_P 0
#endif
// The copy chunk requested at the end of a round is retired (phase P) behind phase C1 and the next-A look-ahead of the
// next round instead of right behind phase A: it has ~60 % of a round to arrive instead of ~40 %.  cp.async groups
// complete in order, so nothing younger than the chunk may be waited for before phase P: the distance look-ahead of
// phase A becomes a plain (predicated, top-level) global load into a register, tracked by its own scoreboard.
#ifndef BD_LANE_LATE_RETIRE
#define BD_LANE_LATE_RETIRE 0
#endif
// Look-ahead entries loaded into a register (predicated ld.global at the top level of the round, own scoreboard) instead of
// cp.async into the staging area: 1 the distance symbol's, 2 the next phase-A symbol's as well.  Fewer cp.async requests
// share the path into shared memory with the copy chunks.
#ifndef BD_LANE_LA_REG
#define BD_LANE_LA_REG 0
#endif
#define BD_LANE_DIST_LA_REG (BD_LANE_LATE_RETIRE || BD_LANE_LA_REG >= 1)
#define BD_LANE_NEXT_LA_REG (BD_LANE_LA_REG >= 2)
// The copy chunk's source blocks are loaded into registers (two predicated ld.global.cg.v4 at the top level of the round, each
// with its own scoreboard) and put into the landing zone when phase P retires them, instead of cp.async.  Measured: 110.1
// against 87.9 ms on the headline probe (profiles/r02/r08s_chunk_ldg.txt) -- register-destination loads held across a round
// are what cp.async replaced in round 1, and they still lose.
#ifndef BD_LANE_CHUNK_LDG
#define BD_LANE_CHUNK_LDG 0
#endif
// MEASUREMENT ONLY (results stay right): every warp times its three cp.async waits and its rounds with clock64() and adds
// them to a histogram in device memory (half-octave buckets of cycles; read by BrotliB200ProbeWaitHist, lane kernel TU)
#ifndef BD_LANE_WAIT_HIST
#define BD_LANE_WAIT_HIST 0
#endif
// MEASUREMENT ONLY (output is wrong; control flow of streams without literal contexts is unchanged): bit 0 drops the global
// output stores, bit 1 the copy-source requests, bit 2 the history-ring mirror stores -- what do these instructions cost?
#ifndef BD_LANE_PROBE_ABLATE
#define BD_LANE_PROBE_ABLATE 0
#endif
// roots that live in the arena are looked up asynchronously as well (their root entry is requested a phase ahead)
#ifndef BD_LANE_ASYNC_ARENA_ROOTS
#define BD_LANE_ASYNC_ARENA_ROOTS 1
#endif
#ifndef BD_LANE_ARENA_LANES
#define BD_LANE_ARENA_LANES 2  /* lanes of a warp with a tree group in the arena from which the warp takes that loop instance */
#endif
// per-metablock table construction without the sort arrays (temporaries in the shared staging area)
// per-metablock table construction with its small temporaries (count[], the code-length lookup, offs[]) in the shared
// landing zones and the bit window in registers; length-ordered fill as before (0: everything in local memory)
#ifndef BD_LANE_HEADER_SMEM
#define BD_LANE_HEADER_SMEM 1
#endif
#ifndef BD_LANE_HEADER_V2
#define BD_LANE_HEADER_V2 0
#endif
// one input-block request per round instead of one test per bit skip
#ifndef BD_LANE_SKIP_LITE
#define BD_LANE_SKIP_LITE 0
#endif
#ifndef BD_LANE_NARROW_ROOTS
#define BD_LANE_NARROW_ROOTS 0
#endif
#ifndef BD_LANE_W_LIT
#define BD_LANE_W_LIT 1
#endif
#ifndef BD_LANE_W_CMD
#define BD_LANE_W_CMD 1
#endif
#ifndef BD_LANE_W_DIST
#define BD_LANE_W_DIST 1
#endif
#ifndef BD_LANE_BURST_LANES
#define BD_LANE_BURST_LANES 12
#endif
constexpr uint32_t kBurst = BD_LANE_BURST, kBurstLanes = BD_LANE_BURST_LANES;
struct ArenaLayout {
  static constexpr size_t kTab = 0;                                   // u16[kGlobalTab]
  static constexpr size_t kCtxLit = kTab + 2 * (size_t)kGlobalTab;    // u8[64 * kMaxBlockTypes]
  static constexpr size_t kCtxDist = kCtxLit + 64 * (size_t)kMaxBlockTypes;  // u8[4 * kMaxBlockTypes]
  static constexpr size_t kCtxModes = kCtxDist + 4 * (size_t)kMaxBlockTypes;  // u8[kMaxBlockTypes]
  static constexpr size_t kBytes = (kCtxModes + kMaxBlockTypes + 255) & ~size_t(255);
};

enum : int { kLaneOk = 0, kLaneDone = 1, kLaneBail = 2, kLaneNext = 3 };  // kLaneNext: the metablock had no commands (raw bytes / metadata) and is done
constexpr uint32_t kMaxLaneRaw = 1u << 18;  // uncompressed metablocks up to this size are copied by the lane itself (one active lane): larger ones go to the exact kernel

#if defined(BROTLI_B200_HOSTSIM)
static inline void sts8(hw::sref_t a, uint32_t v) { *(uint8_t*)a = (uint8_t)v; }
static inline void sts16(hw::sref_t a, uint32_t v) { *(uint16_t*)a = (uint16_t)v; }
static inline void sts32(hw::sref_t a, uint32_t v) { *(uint32_t*)a = v; }
static inline uint32_t vlds16(hw::sref_t a) { return *(const uint16_t*)a; }
static inline uint32_t vlds32(hw::sref_t a) { return *(const uint32_t*)a; }
static inline uint2 vlds64(hw::sref_t a) { return *(const uint2*)a; }
static inline uint32_t vlds8(hw::sref_t a) { return *(const uint8_t*)a; }
static inline uint32_t funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t s) {
  if (s >= 32) return hi;
  return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
static inline uint32_t funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {  // high word of (hi:lo) << s, s in 0..31
  s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline void st32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void st32_if(bool cond, uint8_t* p, uint32_t v) { if (cond) memcpy(p, &v, 4); }
static inline void sts32_if(bool cond, hw::sref_t a, uint32_t v) { if (cond) *(uint32_t*)a = v; }
static inline void warp_sync() {}
static inline bool warp_any(bool p) { return p; }
static inline uint32_t warp_count(bool p) { return p ? 32u : 0u; }  /* the one simulated lane stands for a full warp */
#define BD_PIN32(x) ((void)0)
#define BD_PIN64(x) ((void)0)
static inline void ld32_if(bool cond, const uint8_t* p, uint32_t& dst) { if (cond) memcpy(&dst, p, 4); }
static inline void probe_ldg16_if(bool, const void*) {}
static inline void ldg128_if(bool cond, const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  if (cond) { uint32_t t[4]; memcpy(t, p, 16); a = t[0]; b = t[1]; c = t[2]; d = t[3]; }
}
static inline void sts128(hw::sref_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { uint32_t t[4] = {a, b, c, d}; memcpy((void*)dst, t, 16); }
static inline void ld16_if(bool cond, const uint16_t* p, uint32_t& dst) { if (cond) dst = *p; }
// asynchronous 16-byte copy global -> "shared": immediate on the host
static inline void cp_async16(hw::sref_t dst, const uint8_t* src) { memcpy((void*)dst, src, 16); }
static inline void cp_async16_if(bool cond, hw::sref_t dst, const void* src) { if (cond) memcpy((void*)dst, src, 16); }
static inline uint64_t l2_policy_stream() { return 0; }
static inline uint64_t l2_policy_keep() { return 0; }
static inline void cp_async16_hint(hw::sref_t dst, const uint8_t* src, uint64_t) { memcpy((void*)dst, src, 16); }
static inline void cp_async16_if_hint(bool cond, hw::sref_t dst, const void* src, uint64_t) { if (cond) memcpy((void*)dst, src, 16); }
static inline void st32_stream(uint8_t* p, uint32_t v, uint64_t) { memcpy(p, &v, 4); }
static inline void cp_async_commit() {}
static inline void cp_async_wait_all_but_latest() {}
static inline void cp_async_wait_all() {}
#else
// volatile: these loads follow stores to the same slot (table fill) made through asm as well
BD_DEV void sts8(hw::sref_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
BD_DEV void sts16(hw::sref_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); }
BD_DEV void sts32(hw::sref_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
BD_DEV uint32_t vlds16(hw::sref_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint32_t vlds32(hw::sref_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint32_t vlds8(hw::sref_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
BD_DEV uint2 vlds64(hw::sref_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
BD_DEV uint32_t funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_rc(lo, hi, s); }
BD_DEV uint32_t funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
BD_DEV uint32_t ld32(const uint8_t* p) { return *(const uint32_t*)p; }
#ifndef BD_LANE_L2_HINTS
#define BD_LANE_L2_HINTS 4
#endif
// modes: 0 no hints; 1 input / output / copy sources evict-first, table arena evict-last; 2 as 1, copy sources
// without a hint; 3 as 1, output stores without a hint; 4 (default) as 1, table arena without a hint; 5 as 4, output
// stores evict-LAST (does a lane's recent output stay in L2 for its short-distance copies?)
#if BD_LANE_PROBE_ABLATE & 1
BD_DEV void st32(uint8_t*, uint32_t) {}
#elif BD_LANE_L2_HINTS == 5
// (not volatile, no inputs: the compiler keeps one copy of the policy per function)
BD_DEV uint64_t l2_policy_keep_pure() { uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
BD_DEV void st32(uint8_t* p, uint32_t v) { asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_keep_pure()) : "memory"); }
#elif BD_LANE_L2_HINTS == 1 || BD_LANE_L2_HINTS == 2 || BD_LANE_L2_HINTS == 4
BD_DEV void st32(uint8_t* p, uint32_t v) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }  // streaming: evict first
#else
BD_DEV void st32(uint8_t* p, uint32_t v) { *(uint32_t*)p = v; }
#endif
// predicated stores that stay predicated (no branch around a one-instruction body)
BD_DEV void sts32_if(bool cond, hw::sref_t a, uint32_t v) {
#if BD_LANE_PROBE_ABLATE & 4
  return;
#endif
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t}" ::"r"((uint32_t)cond), "r"(a), "r"(v) : "memory");
}
BD_DEV void st32_if(bool cond, uint8_t* p, uint32_t v) {
#if BD_LANE_PROBE_ABLATE & 1
  (void)cond; (void)p; (void)v;
#elif BD_LANE_L2_HINTS == 5
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.L2::cache_hint.u32 [%1], %2, %3;\n\t}" ::"r"((uint32_t)cond), "l"(p), "r"(v), "l"(l2_policy_keep_pure()) : "memory");
#elif BD_LANE_L2_HINTS == 1 || BD_LANE_L2_HINTS == 2 || BD_LANE_L2_HINTS == 4
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.cs.u32 [%1], %2;\n\t}" ::"r"((uint32_t)cond), "l"(p), "r"(v) : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.u32 [%1], %2;\n\t}" ::"r"((uint32_t)cond), "l"(p), "r"(v) : "memory");
#endif
}
// LDGSTS: the input stream reaches shared memory without passing through a register, so nothing in the
// decode loop ever waits on (or moves) an in-flight global load of compressed bytes
BD_DEV void cp_async16(hw::sref_t dst, const uint8_t* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
// L2 eviction policies: streaming data (compressed input, output words, backreference sources) is marked
// evict-first, the per-lane table arena evict-last, so that the tables stay L2-resident under the stream
BD_DEV uint64_t l2_policy_stream() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
BD_DEV uint64_t l2_policy_keep() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
BD_DEV void cp_async16_hint(hw::sref_t dst, const uint8_t* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
BD_DEV void cp_async16_if_hint(bool cond, hw::sref_t dst, const void* src, uint64_t pol) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p cp.async.cg.shared.global.L2::cache_hint [%1], [%2], 16, %3;\n\t}" ::"r"((uint32_t)cond), "r"(dst), "l"(src), "l"(pol) : "memory");
}
BD_DEV void st32_stream(uint8_t* p, uint32_t v, uint64_t pol) { asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory"); }
BD_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// predicated 16-byte copy (stays predicated: no branch, see ld16_if).  .cg: served by L2, where this thread's
// earlier stores are, never by a possibly stale L1 line.
BD_DEV void cp_async16_if(bool cond, hw::sref_t dst, const void* src) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p cp.async.cg.shared.global [%1], [%2], 16;\n\t}" ::"r"((uint32_t)cond), "r"(dst), "l"(src) : "memory");
}
// NOTE: the hardware counts committed groups per WARP (one dependency counter), so commits and waits with a
// nonzero count are only placed where the whole warp is converged.
BD_DEV void cp_async_wait_all_but_latest() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
BD_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
BD_DEV void warp_sync() { __syncwarp(); }
// Pin a loop-invariant copy in a register: without this the compiler re-reads such values from the
// (address-taken) structs in local memory every round, and with the L1 full of streaming data each of
// those "free" reloads is an L2 round trip on the round's critical path.
// Predicated global loads that stay predicated (no branch).  ptxas tracks outstanding loads per control-flow
// path; a load issued inside an if-block is waited for where that block rejoins the main path, which would
// put the whole L2/HBM latency back on the round's critical path.  Without a branch there is no such join,
// and the wait moves to the first real use (a round later).
BD_DEV void ld32_if(bool cond, const uint8_t* p, uint32_t& dst) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.global.u32 %0, [%2];\n\t}" : "+r"(dst) : "r"((uint32_t)cond), "l"(p) : "memory");
}
BD_DEV void ld16_if(bool cond, const uint16_t* p, uint32_t& dst) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p ld.global.u16 %0, [%2];\n\t}" : "+r"(dst) : "r"((uint32_t)cond), "l"(p) : "memory");
}
#define BD_PIN32(x) asm volatile("" : "+r"(x))
#define BD_PIN64(x) asm volatile("" : "+l"(x))
// MEASUREMENT ONLY: a predicated 16-byte global load whose result is dropped (BD_LANE_PROBE_ABLATE bit 3)
BD_DEV void probe_ldg16_if(bool cond, const void* p) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 a<4>;\n\tsetp.ne.u32 p, %0, 0;\n\t@p ld.global.cg.v4.u32 {a0, a1, a2, a3}, [%1];\n\t}" ::"r"((uint32_t)cond), "l"(p) : "memory");
}
BD_DEV void ldg128_if(bool cond, const void* p, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}"
               : "+r"(a), "+r"(b), "+r"(c), "+r"(d) : "r"((uint32_t)cond), "l"(p) : "memory");
}
BD_DEV void sts128(hw::sref_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
BD_DEV bool warp_any(bool p) { return __any_sync(0xffffffffu, p); }  // all 32 lanes take part
BD_DEV uint32_t warp_count(bool p) { return __popc(__ballot_sync(0xffffffffu, p)); }
#endif

BD_DEV uint32_t mask_bits(uint32_t n) { return (1u << n) - 1u; }  // n <= 31
// x & mask_bits(n) in one instruction (SGXT.U32); n >= 32 keeps x
#if defined(BROTLI_B200_HOSTSIM)
static inline uint32_t low_bits(uint32_t x, uint32_t n) { return n >= 32 ? x : x & ((1u << n) - 1u); }
#else
BD_DEV uint32_t low_bits(uint32_t x, uint32_t n) { uint32_t r; asm("szext.clamp.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }
#endif

#if !defined(BROTLI_B200_HOSTSIM)
// LUTs read by all lanes of the CTA (filled by the kernel before its first __syncthreads()).  At namespace scope so
// that the command loop addresses them as link-time constants instead of carrying four pointers in registers.
__shared__ uint2 g_cmd_lut[704];                       // pack_cmd_lut
__shared__ __align__(16) uint8_t g_ctx_lut[2048];      // kBrotliContextLookup
__shared__ uint32_t g_word_info[25];                   // pack_word_info
__shared__ uint32_t g_transform_info[BROTLI_NUM_TRANSFORMS];  // pack_transform_info
#endif

#if BD_LANE_WAIT_HIST && !defined(BROTLI_B200_HOSTSIM)
constexpr uint32_t kWaitHistWarps = 4096, kWaitHistRow = 4 * 64 + 8;  // per warp (no contention): [site][bucket] counts, then the four cycle sums
__device__ unsigned long long g_wait_hist_all[kWaitHistWarps * kWaitHistRow];
BD_DEV void wait_hist_add(uint32_t site, long long dt) {
  if ((threadIdx.x & 31u) != 0) return;
  unsigned long long* const g_wait_hist = g_wait_hist_all + (size_t)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % kWaitHistWarps) * kWaitHistRow;
  const unsigned long long d = dt < 0 ? 0ull : (unsigned long long)dt;
  const uint32_t lg = 63u - (uint32_t)__clzll((long long)(d + 1));          // floor(log2(d + 1))
  const uint32_t half = lg == 0 ? 0u : (uint32_t)(((d + 1) >> (lg - 1)) & 1u);  // upper half of the octave
  uint32_t b = 2u * lg + half; if (b > 63u) b = 63u;
  atomicAdd(&g_wait_hist[site * 64 + b], 1ull);
  atomicAdd(&g_wait_hist[256 + site], d);
}
#define LN_WAIT_T0() const long long wt0_ = clock64()
#define LN_WAIT_T1(SITE) wait_hist_add(SITE, clock64() - wt0_)
#else
#define LN_WAIT_T0() ((void)0)
#define LN_WAIT_T1(SITE) ((void)0)
#endif

// Constant per-lane context (where this lane's storage lives).
struct LaneCtx {
  hw::sref_t slot;       // shared: kSlotHeaderBytes header, then E u16 table entries
  hw::sref_t stab;       // slot + kSlotHeaderBytes
  uint32_t E;            // entries of virtual table space that live in the shared slot
  uint16_t* gtab;        // arena: virtual entries E .. E + kGlobalTab
  uint8_t* ctx_lit;
  uint8_t* ctx_dist;
  uint8_t* ctx_modes;
  hw::sref_t hist;       // shared: 32-byte ring mirroring this lane's most recent output words
  hw::sref_t stage_c;    // shared, 16-byte aligned: the block holding the second-level table entry of the next phase-C symbol (its
                         // own array, 16 bytes per lane: with stage at a 48-byte stride a warp's 16-byte cp.async writes are free
                         // of bank conflicts; at 64 bytes per lane sixteen lanes shared each bank group)
  hw::sref_t stage;      // shared, 16-byte aligned: [0..31] two 16-byte blocks of copy source, [32..47] / [48..63] the
                         // block holding the second-level table entry of the next phase-A / phase-C symbol
  hw::sref_t ring;       // shared: this lane's first 16-byte input block buffer; the second one is ring_stride further
  uint32_t ring_stride;
  hw::sref_t cmd_lut;    // shared: uint2[704], pack_cmd_lut
  hw::sref_t ctx_lut;    // shared: u8[2048]
  const uint8_t* dictionary;  // RFC 7932 dictionary (source of the expanded table)
  const uint8_t* xdict;  // expanded static dictionary (every word under every transform), see xdict_layout
  hw::sref_t word_info;       // shared: u32[25], pack_word_info
  const uint8_t* cdict;       // custom LZ77 dictionary of the batch (kDict instances only; 16 readable bytes on either side)
  uint64_t cdict_len;
  hw::sref_t transform_info;  // shared: u32[121], pack_transform_info
};

// Word j of the input for the bit window (per-metablock code: synchronous).  When j enters a new 16-byte block,
// the block after it is requested into the ring half whose words are all in registers already, and everything
// requested so far is waited for.  Blocks past the end of the stream repeat the last one.  The command loop
// has its own asynchronous version of this (LN_SKIP).
BD_DEV uint32_t ring_next(const uint8_t* gin, hw::sref_t ring, uint32_t ring_stride, uint32_t last_blk, uint32_t j) {
  if ((j & 3u) == 0) {
    const uint32_t b = (j >> 2) + 1;
    cp_async16(ring + (b & 1u) * ring_stride, gin + 16 * (size_t)(b < last_blk ? b : last_blk));
    cp_async_commit();
    cp_async_wait_all();
  }
  return vlds32(ring + ((j >> 2) & 1u) * ring_stride + (j & 3u) * 4u);
}

// Register copy of a lane's bit window for the loops of the per-metablock code.  The Lane record itself lives in
// local memory (its address is passed around), so every store through a byte pointer or an asm statement with a
// memory clobber would force its fields back to memory and in again -- an L2 round trip each.
struct BitWin {
  const uint8_t* gin;
  hw::sref_t ring;
  uint32_t ring_stride, lo, hi, nx, k, bp, last_blk;
  uint64_t end_bit;
  BD_DEV uint32_t peek() const { return hw::funnelshift_r(lo, hi, bp); }
  BD_DEV void skip(uint32_t n) {
    bp += n;
    if (bp >= 32) {
      lo = hi; hi = nx; k++;
      nx = ring_next(gin, ring, ring_stride, last_blk, k + 2);
      bp -= 32;
    }
  }
  BD_DEV uint32_t read(uint32_t n) {  // n <= 25
    const uint32_t v = peek() & mask_bits(n);
    skip(n);
    return v;
  }
  BD_DEV bool overrun() const { return (uint64_t)k * 32 + bp > end_bit; }
};

// Decoder state of one lane's stream.  Lives in local memory for the per-metablock (cold) code; the
// command loop works on register copies.
struct Lane {
  // bit window: 96 bits of look-ahead (lo, hi, nx = words k, k+1, k+2 of the 16-byte aligned input) in
  // registers, fed from a two-block ring in shared memory that cp.async fills one block ahead
  const uint8_t* gin;    // 16-byte aligned base of the input
  hw::sref_t ring;
  uint32_t ring_stride;
  uint32_t lo, hi, nx, k, bp;
  uint32_t k_max;        // last word holding stream bytes
  uint32_t last_blk;     // last 16-byte block holding stream bytes; copies never go past it
  uint64_t end_bit;      // 8 * (lead + size), relative to the aligned base
  uint32_t lead;
  // output write combiner; positions are biased by (out & 3) so that word boundaries are absolute
  uint8_t* out_al;
  uint32_t posb, acc, bias, capb;
  // stream / metablock
  uint32_t wbits, max_backward, is_last;
  // custom dictionary as this stream's window sees it (src/decode.rs:1831-1838, :2954-2955), kDict instances only
  const uint8_t* cdict_end;  // one past the dictionary's last byte
  uint32_t cdict_size;       // reachable bytes: min(length, max_backward)
  int32_t cdict_limit;       // positions below it have max_distance = pos + cdict_size
  int32_t mlen;
  int32_t d0, d1, d2, d3;
  uint32_t bl[3], nbt[3], rb[6];   // category order: 0 literal, 1 command, 2 distance (reference order)
  uint32_t npostfix, ndirect, dist_alphabet;
  uint32_t n_lit, n_dist;
  uint32_t rbits[3], root[3];  // per group (0 literal, 1 command, 2 distance): root width, root base (virtual index)
  uint32_t trivial_lo, trivial_hi;      // bit i: literal block type i uses one tree for all 64 contexts
  uint32_t trivial, lit_tree, ctx_mode_off, ctx_slice, cmd_tree, dist_slice;
  uint32_t dist_home;    // 0, or (distance-tree cache) the virtual index in the arena where all distance trees' roots live
  uint32_t dist_slots;   // cache slots: the largest number of distinct trees any distance block type uses (1..4)
  uint32_t cold_next;    // next free virtual index of the arena part of the table space
  uint32_t e_tab;        // entries of the shared slot available to tables in this metablock (E, or E - kCtxMapEntries)

  BD_DEV uint32_t peek() const { return hw::funnelshift_r(lo, hi, bp); }
  BD_DEV void skip(uint32_t n) {
    bp += n;
    if (bp >= 32) {
      lo = hi; hi = nx; k++;
      nx = ring_next(gin, ring, ring_stride, last_blk, k + 2);
      bp -= 32;
    }
  }
  BD_DEV uint32_t read(uint32_t n) {  // n <= 25
    const uint32_t v = peek() & mask_bits(n);
    skip(n);
    return v;
  }
  BD_DEV bool overrun() const { return (uint64_t)k * 32 + bp > end_bit; }
  BD_DEV uint32_t pos() const { return posb - bias; }
  BD_DEV BitWin win() const {
    BitWin b;
    b.gin = gin; b.ring = ring; b.ring_stride = ring_stride; b.lo = lo; b.hi = hi; b.nx = nx; b.k = k; b.bp = bp;
    b.last_blk = last_blk; b.end_bit = end_bit;
    return b;
  }
  BD_DEV void put(const BitWin& b) { lo = b.lo; hi = b.hi; nx = b.nx; k = b.k; bp = b.bp; }
};

// ---- virtual table space ----
BD_DEV uint32_t tab_load(const LaneCtx& c, uint32_t v) {
  return v < c.E ? vlds16(c.stab + (v << 1)) : (uint32_t)c.gtab[v - c.E];
}
BD_DEV void tab_store(const LaneCtx& c, uint32_t v, uint32_t e) {
  if (v < c.E) sts16(c.stab + (v << 1), e); else c.gtab[v - c.E] = (uint16_t)e;
}

// One symbol of the tree whose root (2^rbits entries) starts at virtual index root_v.
// Entry = symbol << 4 | code length; length > rbits marks a pointer: its second-level table starts at
// virtual index E + 4 * value and is indexed by the next (length - rbits) bits.
template <class Reader>
BD_DEV uint32_t decode_generic(const LaneCtx& c, Reader& L, uint32_t root_v, uint32_t rbits) {
  const uint32_t bits = L.peek();
  uint32_t e = tab_load(c, root_v + (bits & mask_bits(rbits)));
  uint32_t len = e & 15u;
  if (len > rbits) {
    e = c.gtab[((e >> 4) << 2) + ((bits >> rbits) & mask_bits(len - rbits))];
    len = e & 15u;
  }
  L.skip(len);
  return e >> 4;
}

// ---- output write combiner ----
// Every word that leaves for global memory is mirrored into a 32-byte per-lane history ring in shared memory
// (hist + (position & 31)): short-distance copies read their source from there instead of waiting for a
// just-stored byte to come back from L2.
// first word of an unaligned region: bytes below the region are not ours (out of line: once per stream at most)
BD_COLD void store_head_bytes(uint8_t* out_al, uint32_t bias, uint32_t wpos, uint32_t word) {
  for (uint32_t j = 0; j < 4; j++) if (wpos + j >= bias) out_al[wpos + j] = (uint8_t)(word >> (8 * j));
}
// Experiment prepared for the next round, OFF (the default build is instruction-for-instruction what was measured):
// BD_LANE_HEAD_PER_ROUND=1 takes the unaligned-head test out of every output store.  The stores skip the region's
// first word when the region starts inside it, and the command loop writes that word's bytes (from the history ring)
// right after the phase that completed it -- LN_HEAD_CHECK, three places per round instead of every store site.
#ifndef BD_LANE_HEAD_PER_ROUND
#define BD_LANE_HEAD_PER_ROUND 1
#endif
BD_DEV void store_word_if(bool cond, uint8_t* out_al, uint32_t bias, hw::sref_t hist, uint32_t wpos, uint32_t word) {
  sts32_if(cond, hist + (wpos & 28u), word);
#if BD_LANE_HEAD_PER_ROUND
  st32_if(cond && wpos >= bias, out_al + wpos, word);
  return;
#endif
  if (BD_UNLIKELY(cond && wpos < bias)) {
    store_head_bytes(out_al, bias, wpos, word);
  } else {
    st32_if(cond, out_al + wpos, word);
  }
}
// v holds exactly n (1..4) valid low bytes, the rest is zero
BD_DEV void append(uint8_t* out_al, uint32_t bias, hw::sref_t hist, uint32_t& posb, uint32_t& acc, uint32_t v, uint32_t n) {
  const uint32_t a = posb & 3u, sh = a * 8u;
  const uint32_t word = acc | (v << sh);
  const bool full = a + n >= 4;
  store_word_if(full, out_al, bias, hist, posb & ~3u, word);
  acc = full ? funnelshift_rc(v, 0u, 32u - sh) : word;
  posb += n;
}
// As append(), but a completed word only goes to the history ring; its global store is left to the caller
// (dw_pos: the word's position, 0xFFFFFFFF = none pending; a second completed word sends the first one off).
BD_DEV void append_defer(uint8_t* out_al, uint32_t bias, hw::sref_t hist, uint32_t& posb, uint32_t& acc, uint32_t v, uint32_t n,
                         uint32_t& dw_pos, uint32_t& dw_word) {
  const uint32_t a = posb & 3u, sh = a * 8u;
  const uint32_t word = acc | (v << sh);
  const bool full = a + n >= 4;
  sts32_if(full, hist + (posb & 28u), word);
  if (BD_UNLIKELY(full && dw_pos != 0xFFFFFFFFu)) st32_if(dw_pos >= bias, out_al + dw_pos, dw_word);
  dw_pos = full ? posb & ~3u : dw_pos;
  dw_word = full ? word : dw_word;
  acc = full ? funnelshift_rc(v, 0u, 32u - sh) : word;
  posb += n;
}
// v_hi:v_lo holds exactly n (1..8) valid low bytes, the rest is zero
BD_DEV void append8(uint8_t* out_al, uint32_t bias, hw::sref_t hist, uint32_t& posb, uint32_t& acc, uint32_t v_lo, uint32_t v_hi, uint32_t n) {
  const uint32_t a = posb & 3u, sh = a * 8u, wpos = posb & ~3u;
  const uint32_t x0 = acc | (v_lo << sh);
  const uint32_t x1 = funnelshift_l(v_lo, v_hi, sh);
  const uint32_t t = a + n;
  const bool s0 = t >= 4, s1 = t >= 8;
#if BD_LANE_PROBE_ABLATE & 16
  sts32_if(s0, hist + (wpos & 28u), x0);
  sts32_if(s1, hist + ((wpos + 4) & 28u), x1);
#else
  store_word_if(s0, out_al, bias, hist, wpos, x0);
  sts32_if(s1, hist + ((wpos + 4) & 28u), x1);
  st32_if(s1, out_al + wpos + 4, x1);
#endif
  acc = s1 ? funnelshift_rc(v_hi, 0u, 32u - sh) : (s0 ? x1 : x0);
  posb += n;
}
BD_DEV void flush_partial(uint8_t* out_al, uint32_t bias, uint32_t posb, uint32_t acc) {
  const uint32_t a = posb & 3u, wpos = posb & ~3u;
  for (uint32_t j = 0; j < a; j++) if (wpos + j >= bias) out_al[wpos + j] = (uint8_t)(acc >> (8 * j));
}
// Four recent output bytes sp .. sp+3 (sp + 3 < posb, sp within the last 28 bytes): from the partial word and the
// history ring.
BD_DEV uint32_t recent4(hw::sref_t hist, uint32_t posb, uint32_t acc, uint32_t sp) {
  const uint32_t wcur = posb & ~3u, w0p = sp & ~3u, w1p = w0p + 4;
  const uint32_t w0 = w0p == wcur ? acc : vlds32(hist + (w0p & 28u));
  const uint32_t w1 = w1p == wcur ? acc : (w1p < wcur ? vlds32(hist + (w1p & 28u)) : 0u);
  return hw::funnelshift_r(w0, w1, 8 * (sp & 3u));
}
// Last two output bytes (p1 = newest), zero before the start of the stream: the literal context.
BD_DEV void last_two(hw::sref_t hist, uint32_t bias, uint32_t posb, uint32_t acc, uint32_t& p1, uint32_t& p2) {
  const uint32_t a = posb & 3u;
  uint32_t x;  // the four bytes before posb, newest in the top byte
  if (a >= 2) {
    x = acc << (32 - 8 * a);
  } else {
    const uint32_t prev = vlds32(hist + (((posb & ~3u) - 4u) & 28u));  // the word before the partial one
    x = a ? (acc << 24) | (prev >> 8) : prev;
  }
  const uint32_t have = posb - bias;
  p1 = have >= 1 ? x >> 24 : 0u;
  p2 = have >= 2 ? (x >> 16) & 0xFFu : 0u;
}

// ======================= per-metablock (cold) code =======================

// DecodeVarLenUint8, src/decode.rs:193-241
BD_DEV uint32_t read_varlen8(Lane& L) {
  if (L.read(1) == 0) return 0;
  const uint32_t n = L.read(3);
  if (n == 0) return 1;
  return (1u << n) + L.read(n);
}

BD_DEV uint32_t bit_width(uint32_t x) { uint32_t r = 0; while (x) { x >>= 1; r++; } return r; }  // Log2Floor of src/decode.rs:502-509

#if BD_LANE_HEADER_V2
// ---- prefix-code tables from code lengths ----
// Temporaries live in the lane's 64-byte cp.async landing zone (c.stage), which is idle outside the command loop:
// [0..31] count[16] (u16: symbols per code length), [32..63] first the 5-bit lookup of the code-length code (u8[32]),
// then next_code[16] (u16: the next canonical code of each length).  Local memory only holds the code lengths
// themselves (one byte per symbol, read back four at a time): every access there is an L2 round trip.
#error "BD_LANE_HEADER_V2 keeps 64 bytes of temporaries in c.stage, which is 48 bytes per lane now"
BD_DEV hw::sref_t tmp_count(const LaneCtx& c, uint32_t l) { return c.stage + 2u * l; }
BD_DEV hw::sref_t tmp_next(const LaneCtx& c, uint32_t l) { return c.stage + 32u + 2u * l; }

// Roots and second-level tables of one prefix code, canonical codes in bit-reversed order: the shape of
// BrotliBuildHuffmanTable (src/huffman/mod.rs:273-386) with our root width.  count[] (shared) holds the symbols per
// length; the code is complete (checked by the caller).  Root of 2^rbits entries at virtual index root_v; codes
// longer than the root go to second-level tables allocated from L.cold_next (always in the arena), one per root
// prefix, as wide as the longest code under that prefix.  Prepares next_code[]; the entries themselves are written
// by place_symbol for every symbol in increasing order (which is canonical order within a length).
BD_DEV int begin_table(const LaneCtx& c, uint32_t& cold_next, uint32_t root_v, uint32_t rbits) {
  // ascending lengths: first code of each length; provisional root entries (sub-table width) of the long prefixes
  uint32_t f = 0, p0 = 0;
  for (uint32_t l = 1; l <= 15; l++) {
    const uint32_t n = vlds16(tmp_count(c, l));
    sts16(tmp_next(c, l), f);
    if (l > rbits && n != 0) {
      const uint32_t sh = l - rbits;
      const uint32_t plo = f >> sh, phi = (f + n - 1) >> sh;
      for (uint32_t pfx = plo; pfx <= phi; pfx++) tab_store(c, root_v + (rbits ? hw::brev(pfx) >> (32 - rbits) : 0u), sh);
    }
    f += n;
    if (l == rbits) p0 = f;  // prefixes below p0 are codes no longer than the root
    f <<= 1;
  }
  // sub-tables in prefix order
  for (uint32_t pfx = p0; pfx < (1u << rbits); pfx++) {
    const uint32_t v = root_v + (rbits ? hw::brev(pfx) >> (32 - rbits) : 0u);
    const uint32_t sub_w = tab_load(c, v);
    const uint32_t sub_v = cold_next - c.E;  // index into the arena part
    const uint32_t sub_size = sub_w < 2 ? 4u : 1u << sub_w;  // every arena allocation is a multiple of four entries
    if (sub_v + sub_size > kGlobalTab) return kLaneBail;
    cold_next += sub_size;
    tab_store(c, v, ((sub_v >> 2) << 4) | (rbits + sub_w));
  }
  return kLaneOk;
}
BD_DEV void place_symbol(const LaneCtx& c, uint32_t root_v, uint32_t rbits, uint32_t sym, uint32_t l) {
  const hw::sref_t nc = tmp_next(c, l);
  const uint32_t code = vlds16(nc);
  sts16(nc, code + 1);
  const uint32_t rev = hw::brev(code) >> (32 - l);
  const uint32_t e = (sym << 4) | l;
  if (l <= rbits) {
    for (uint32_t t = rev; t < (1u << rbits); t += 1u << l) tab_store(c, root_v + t, e);
  } else {
    const uint32_t ptr = tab_load(c, root_v + (rev & mask_bits(rbits)));
    const uint32_t sub_v = (ptr >> 4) << 2, sub_w = (ptr & 15u) - rbits;
    for (uint32_t t = rev >> rbits; t < (1u << sub_w); t += 1u << (l - rbits)) c.gtab[sub_v + t] = (uint16_t)e;
  }
}

// ReadHuffmanCode, src/decode.rs:868-1013: one prefix-code description -> lookup structure.
BD_DEV int read_huffman_code_impl(const LaneCtx& c, BitWin& L, uint32_t& cold_next, uint32_t alphabet_size, uint32_t max_symbol, uint32_t root_v, uint32_t rbits) {
  for (uint32_t l = 0; l < 16; l += 2) sts32(c.stage + 2u * l, 0u);  // count[]
  const uint32_t hskip = L.read(2);
  if (hskip == 1) {  // simple code: NSYM 1..4 explicit symbols (ReadSimpleHuffmanSymbols, :516-556)
    const uint32_t nsym = L.read(2) + 1;
    const uint32_t max_bits = bit_width(alphabet_size - 1);
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < nsym; i++) {
      s[i] = L.read(max_bits);
      if (s[i] >= max_symbol) return kLaneBail;
    }
    for (uint32_t i = 0; i + 1 < nsym; i++)
      for (uint32_t k = i + 1; k < nsym; k++) if (s[i] == s[k]) return kLaneBail;
    if (nsym == 1) {
      for (uint32_t t = 0; t < (1u << rbits); t++) tab_store(c, root_v + t, s[0] << 4);
      return kLaneOk;
    }
    // canonical codes sorted by (length, value) reproduce BrotliBuildSimpleHuffmanTable (src/huffman/mod.rs:390-471)
    uint32_t len[4] = {1, 1, 0, 0};
    if (nsym == 3) { len[1] = 2; len[2] = 2; }
    if (nsym == 4) {
      if (L.read(1)) { len[0] = 1; len[1] = 2; len[2] = 3; len[3] = 3; } else { len[0] = len[1] = len[2] = len[3] = 2; }
    }
    for (uint32_t i = 0; i < nsym; i++) { const hw::sref_t cn = tmp_count(c, len[i]); sts16(cn, vlds16(cn) + 1); }
    if (begin_table(c, cold_next, root_v, rbits) != kLaneOk) return kLaneBail;
    // symbols in increasing order (four at most: selection by value)
    uint32_t done = 0;
    for (uint32_t n = 0; n < nsym; n++) {
      uint32_t best = 4;
      for (uint32_t i = 0; i < nsym; i++) if (!((done >> i) & 1u) && (best == 4 || s[i] < s[best])) best = i;
      done |= 1u << best;
      place_symbol(c, root_v, rbits, s[best], len[best]);
    }
    return kLaneOk;
  }
  // complex code: code-length code lengths (ReadCodeLengthCodeLengths, :801-853), four bits each in cl_lo (symbols
  // 0..7), cl_mid (8..15), cl_hi (16, 17)
  uint32_t cl_lo = 0, cl_mid = 0, cl_hi = 0;
  uint32_t space = 32, num_codes = 0;
  for (uint32_t i = hskip; i < 18; i++) {
    const uint32_t ix = L.peek() & 15u;
    // kCodeLengthPrefixLength / kCodeLengthPrefixValue (src/decode.rs:59-61) packed four bits per entry
    const uint32_t plen = (uint32_t)(0x4222322242223222ull >> (ix * 4)) & 15u;
    const uint32_t v = (uint32_t)(0x5340234013402340ull >> (ix * 4)) & 15u;
    L.skip(plen);
    const uint32_t sy = tbl::kCodeLengthCodeOrder[i];
    const uint32_t vv = v << ((sy & 7u) * 4u);
    if (sy < 8) cl_lo |= vv; else if (sy < 16) cl_mid |= vv; else cl_hi |= vv;
    if (v != 0) {
      space -= 32u >> v;
      num_codes++;
      if (space - 1u >= 32u) break;  // space is 0 or wrapped
    }
  }
  if (!(num_codes == 1 || space == 0)) return kLaneBail;
  // 5-bit lookup of the code-length code (BrotliBuildCodeLengthsHuffmanTable, src/huffman/mod.rs:196-271): symbol << 3 | length
  const hw::sref_t cl_tab = c.stage + 32u;
#define BD_CL_CL(sy) ((((sy) < 8 ? cl_lo : ((sy) < 16 ? cl_mid : cl_hi)) >> (((sy) & 7u) * 4u)) & 15u)
  if (num_codes == 1) {
    uint32_t only = 0;
    for (uint32_t i = 0; i < 18; i++) if (BD_CL_CL(i)) only = i;
    for (uint32_t t = 0; t < 32; t += 4) sts32(cl_tab + t, (only << 3) * 0x01010101u);
  } else {
    // symbols per length (five bits each), then canonical codes in symbol order
    uint32_t cnt = 0;
    for (uint32_t sy = 0; sy < 18; sy++) { const uint32_t l = BD_CL_CL(sy); if (l) cnt += 1u << (5u * (l - 1u)); }
    uint32_t next = 0, f = 0;  // next code of each length (six bits each)
    for (uint32_t l = 1; l <= 5; l++) {
      next |= f << (6u * (l - 1u));
      f = (f + ((cnt >> (5u * (l - 1u))) & 31u)) << 1;
    }
    for (uint32_t sy = 0; sy < 18; sy++) {
      const uint32_t l = BD_CL_CL(sy);
      if (!l) continue;
      const uint32_t code = (next >> (6u * (l - 1u))) & 63u;
      next += 1u << (6u * (l - 1u));
      const uint32_t rev = hw::brev(code) >> (32 - l);
      for (uint32_t t = rev; t < 32; t += 1u << l) sts8(cl_tab + t, (sy << 3) | l);
    }
  }
#undef BD_CL_CL
  // symbol code lengths with repeat codes (ReadSymbolCodeLengths, :661-731; Process*CodeLength :565-658)
  uint32_t cl_words[176];  // one byte per symbol, read back four at a time
  uint8_t* const cl = (uint8_t*)cl_words;
  uint32_t symbol = 0, prev_len = 8, repeat = 0, repeat_len = 0;
  space = 32768;
  if (max_symbol > 704) return kLaneBail;
  while (symbol < max_symbol && space > 0) {
    const uint32_t p = vlds8(cl_tab + (L.peek() & 31u));
    L.skip(p & 7u);
    const uint32_t code_len = p >> 3;
    if (code_len < 16) {
      repeat = 0;
      cl[symbol] = (uint8_t)code_len;
      if (code_len != 0) {
        prev_len = code_len;
        space -= 32768u >> code_len;
        const hw::sref_t cn = tmp_count(c, code_len);
        sts16(cn, vlds16(cn) + 1);
      }
      symbol++;
    } else {
      const uint32_t extra_bits = code_len - 14;
      uint32_t delta = L.read(extra_bits);
      const uint32_t new_len = code_len == 16 ? prev_len : 0;
      if (repeat_len != new_len) { repeat = 0; repeat_len = new_len; }
      const uint32_t old_repeat = repeat;
      if (repeat > 0) { repeat -= 2; repeat <<= extra_bits; }
      repeat += delta + 3;
      delta = repeat - old_repeat;
      if (symbol + delta > max_symbol) { space = 0xFFFFF; break; }
      for (uint32_t j = 0; j < delta; j++) cl[symbol + j] = (uint8_t)repeat_len;
      symbol += delta;
      if (repeat_len != 0) {
        space -= delta << (15 - repeat_len);
        const hw::sref_t cn = tmp_count(c, repeat_len);
        sts16(cn, vlds16(cn) + delta);
      }
    }
  }
  if (space != 0) return kLaneBail;
  if (begin_table(c, cold_next, root_v, rbits) != kLaneOk) return kLaneBail;  // (overwrites the code-length lookup)
  // symbols in increasing order; the words holding their lengths are loaded two iterations ahead
  const uint32_t nw = (symbol + 3) >> 2;
  uint32_t w0 = cl_words[0], w1 = nw > 1 ? cl_words[1] : 0u;
  for (uint32_t i = 0; i < nw; i++) {
    const uint32_t w2 = i + 2 < nw ? cl_words[i + 2] : 0u;
    uint32_t w = w0;
    if (i + 1 == nw && (symbol & 3u) != 0) w &= mask_bits(8u * (symbol & 3u));  // bytes past the last symbol were never written
    for (uint32_t j = 0; w != 0; j++, w >>= 8) {
      const uint32_t l = w & 0xFFu;
      if (l) place_symbol(c, root_v, rbits, 4 * i + j, l);
    }
    w0 = w1; w1 = w2;
  }
  return kLaneOk;
}

BD_COLD int read_huffman_code(const LaneCtx& c, Lane& L, uint32_t alphabet_size, uint32_t max_symbol, uint32_t root_v, uint32_t rbits) {
  BitWin b = L.win();
  uint32_t cold_next = L.cold_next;
  const int r = read_huffman_code_impl(c, b, cold_next, alphabet_size, max_symbol, root_v, rbits);
  L.put(b);
  L.cold_next = cold_next;
  return r;
}

#else
#if BD_LANE_HEADER_SMEM
// ---- prefix-code tables from code lengths, length-ordered fill, small temporaries in shared memory ----
// The lane's cp.async landing zones are idle outside the command loop: c.stage[0..31] holds count[16] (u16: symbols per
// code length); the 32 bytes made of c.stage[32..47] and c.stage_c[0..15] hold first the 5-bit lookup of the code-length
// code (u8[32]) and then offs[16] (u16: where the next symbol of each length goes in sorted[]).  Local memory keeps the
// two big arrays (code length per symbol, symbols sorted by length); every access there is an L2 round trip, and the
// small arrays used to be indexed dynamically, i.e. lived there too.  The bit window is a register copy (BitWin).
BD_DEV hw::sref_t hc_count(const LaneCtx& c, uint32_t l) { return c.stage + 2u * l; }
BD_DEV hw::sref_t hc_tmp(const LaneCtx& c, uint32_t b) { return b < 16u ? c.stage + 32u + b : c.stage_c + (b - 16u); }

// Fill the lookup structure of one prefix code from its symbols sorted by (length, value); count[] in shared memory.
// Root of 2^rbits entries at root_v; longer codes go to second-level tables allocated from cold_next (always in the
// arena).  Same shape as BrotliBuildHuffmanTable (src/huffman/mod.rs:273-386).
BD_DEV int fill_table(const LaneCtx& c, uint32_t& cold_next, const uint32_t* sorted_w, uint32_t root_v, uint32_t rbits) {
  uint32_t code = 0, idx = 0;
  uint32_t pair = 0;  // sorted[] is read two symbols per (local-memory) load
  uint32_t cur_prefix = 0xFFFFFFFFu, sub_w = 0, sub_v = 0;
  for (uint32_t l = 1; l <= 15; l++) {
    for (uint32_t j = vlds16(hc_count(c, l)); j != 0; j--) {
      if ((idx & 1u) == 0) pair = sorted_w[idx >> 1];
      const uint32_t sym = (idx & 1u) ? pair >> 16 : pair & 0xFFFFu;
      idx++;
      const uint32_t rev = hw::brev(code) >> (32 - l);
      const uint32_t e = (sym << 4) | l;
      if (l <= rbits) {
        for (uint32_t t = rev; t < (1u << rbits); t += 1u << l) tab_store(c, root_v + t, e);
      } else {
        const uint32_t prefix = code >> (l - rbits);
        if (prefix != cur_prefix) {  // first code under a new root slot: size its second-level table (NextTableBitSize, :181-193)
          cur_prefix = prefix;
          // symbols not yet placed: j of length l (this one included), all of every longer length
          int32_t left = (1 << (l - rbits)) - (int32_t)j;
          uint32_t ll = l;
          while (ll < 15 && left > 0) {
            ll++; left <<= 1;
            if (ll < 15) left -= (int32_t)vlds16(hc_count(c, ll));
          }
          sub_w = ll - rbits;
          sub_v = cold_next - c.E;  // index into the arena part
          const uint32_t sub_size = sub_w < 2 ? 4u : 1u << sub_w;  // every arena allocation is a multiple of four entries
          if (sub_v + sub_size > kGlobalTab) return kLaneBail;
          cold_next += sub_size;
          tab_store(c, root_v + (rev & mask_bits(rbits)), ((sub_v >> 2) << 4) | (rbits + sub_w));
        }
        for (uint32_t t = rev >> rbits; t < (1u << sub_w); t += 1u << (l - rbits)) c.gtab[sub_v + t] = (uint16_t)e;
      }
      code++;
    }
    code <<= 1;
  }
  return kLaneOk;
}

// ReadHuffmanCode, src/decode.rs:868-1013: one prefix-code description -> lookup structure.
BD_DEV int read_huffman_code_impl(const LaneCtx& c, BitWin& L, uint32_t& cold_next, uint32_t alphabet_size, uint32_t max_symbol,
                                  uint32_t root_v, uint32_t rbits) {
  uint32_t sorted_w[352];  // u16 sorted[704]: symbols by (length, value)
  uint16_t* const sorted = (uint16_t*)sorted_w;
  for (uint32_t l = 0; l < 16; l += 2) sts32(hc_count(c, l), 0u);
  const uint32_t hskip = L.read(2);
  if (hskip == 1) {  // simple code: NSYM 1..4 explicit symbols (ReadSimpleHuffmanSymbols, :516-556)
    const uint32_t nsym = L.read(2) + 1;
    const uint32_t max_bits = bit_width(alphabet_size - 1);
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < nsym; i++) {
      s[i] = L.read(max_bits);
      if (s[i] >= max_symbol) return kLaneBail;
    }
    for (uint32_t i = 0; i + 1 < nsym; i++)
      for (uint32_t k = i + 1; k < nsym; k++) if (s[i] == s[k]) return kLaneBail;
    if (nsym == 1) {
      for (uint32_t t = 0; t < (1u << rbits); t++) tab_store(c, root_v + t, s[0] << 4);
      return kLaneOk;
    }
    // canonical codes sorted by (length, value) reproduce BrotliBuildSimpleHuffmanTable (src/huffman/mod.rs:390-471)
    uint32_t len[4] = {1, 1, 0, 0};
    if (nsym == 3) { len[1] = 2; len[2] = 2; }
    if (nsym == 4) {
      if (L.read(1)) { len[0] = 1; len[1] = 2; len[2] = 3; len[3] = 3; } else { len[0] = len[1] = len[2] = len[3] = 2; }
    }
    uint32_t n = 0;
    for (uint32_t l = 1; l <= 3; l++) {
      const uint32_t first = n;
      for (uint32_t i = 0; i < nsym; i++) if (len[i] == l) {
        uint32_t q = n++;
        while (q > first && sorted[q - 1] > s[i]) { sorted[q] = sorted[q - 1]; q--; }
        sorted[q] = (uint16_t)s[i];
        sts16(hc_count(c, l), vlds16(hc_count(c, l)) + 1u);
      }
    }
    asm volatile("" ::: "memory");  // (sorted[] was written as u16, fill_table reads it as u32 pairs)
    return fill_table(c, cold_next, sorted_w, root_v, rbits);
  }
  // complex code: code-length code lengths (ReadCodeLengthCodeLengths, :801-853)
  uint8_t cl_cl[18];
  for (uint32_t i = 0; i < 18; i++) cl_cl[i] = 0;
  uint32_t space = 32, num_codes = 0;
  for (uint32_t i = hskip; i < 18; i++) {
    const uint32_t ix = L.peek() & 15u;
    // kCodeLengthPrefixLength / kCodeLengthPrefixValue (src/decode.rs:59-61) packed four bits per entry
    const uint32_t plen = (uint32_t)(0x4222322242223222ull >> (ix * 4)) & 15u;
    const uint32_t v = (uint32_t)(0x5340234013402340ull >> (ix * 4)) & 15u;
    L.skip(plen);
    cl_cl[tbl::kCodeLengthCodeOrder[i]] = (uint8_t)v;
    if (v != 0) {
      space -= 32u >> v;
      num_codes++;
      if (space - 1u >= 32u) break;  // space is 0 or wrapped
    }
  }
  if (!(num_codes == 1 || space == 0)) return kLaneBail;
  // 5-bit lookup of the code-length code (BrotliBuildCodeLengthsHuffmanTable, src/huffman/mod.rs:196-271): symbol << 3 | length
  if (num_codes == 1) {
    uint32_t only = 0;
    for (uint32_t i = 0; i < 18; i++) if (cl_cl[i]) only = i;
    for (uint32_t t = 0; t < 32; t++) sts8(hc_tmp(c, t), only << 3);
  } else {
    uint32_t code = 0;
    for (uint32_t l = 1; l <= 5; l++) {
      for (uint32_t sy = 0; sy < 18; sy++) if (cl_cl[sy] == l) {
        const uint32_t rev = hw::brev(code) >> (32 - l);
        for (uint32_t t = rev; t < 32; t += 1u << l) sts8(hc_tmp(c, t), (sy << 3) | l);
        code++;
      }
      code <<= 1;
    }
  }
  // symbol code lengths with repeat codes (ReadSymbolCodeLengths, :661-731; Process*CodeLength :565-658)
  uint32_t cl_w[176];  // u8 cl[704]: code length per symbol, read back four at a time
  uint8_t* const cl = (uint8_t*)cl_w;
  uint32_t symbol = 0, prev_len = 8, repeat = 0, repeat_len = 0;
  space = 32768;
  if (max_symbol > 704) return kLaneBail;
  while (symbol < max_symbol && space > 0) {
    const uint32_t p = vlds8(hc_tmp(c, L.peek() & 31u));
    L.skip(p & 7u);
    const uint32_t code_len = p >> 3;
    if (code_len < 16) {
      repeat = 0;
      cl[symbol] = (uint8_t)code_len;
      if (code_len != 0) {
        prev_len = code_len;
        space -= 32768u >> code_len;
        const hw::sref_t cn = hc_count(c, code_len);
        sts16(cn, vlds16(cn) + 1u);
      }
      symbol++;
    } else {
      const uint32_t extra_bits = code_len - 14;
      uint32_t delta = L.read(extra_bits);
      const uint32_t new_len = code_len == 16 ? prev_len : 0;
      if (repeat_len != new_len) { repeat = 0; repeat_len = new_len; }
      const uint32_t old_repeat = repeat;
      if (repeat > 0) { repeat -= 2; repeat <<= extra_bits; }
      repeat += delta + 3;
      delta = repeat - old_repeat;
      if (symbol + delta > max_symbol) { space = 0xFFFFF; break; }
      for (uint32_t j = 0; j < delta; j++) cl[symbol + j] = (uint8_t)repeat_len;
      symbol += delta;
      if (repeat_len != 0) {
        space -= delta << (15 - repeat_len);
        const hw::sref_t cn = hc_count(c, repeat_len);
        sts16(cn, vlds16(cn) + delta);
      }
    }
  }
  if (space != 0) return kLaneBail;
  asm volatile("" ::: "memory");  // (cl[] was written as bytes, the loop below reads it as words)
  // offs[] takes the place of the code-length lookup
  uint32_t o = 0;
  for (uint32_t l = 1; l <= 15; l++) { sts16(hc_tmp(c, 2u * l), o); o += vlds16(hc_count(c, l)); }
  for (uint32_t base = 0; base < symbol; base += 4) {
    uint32_t w = cl_w[base >> 2];
    if (base + 4 > symbol) w &= mask_bits(8u * (symbol - base));  // bytes past the last symbol were never written
    for (uint32_t sy = base; w != 0; sy++, w >>= 8) {
      const uint32_t l = w & 0xFFu;
      if (l) {
        const hw::sref_t on = hc_tmp(c, 2u * l);
        const uint32_t at = vlds16(on);
        sorted[at] = (uint16_t)sy;
        sts16(on, at + 1u);
      }
    }
  }
  asm volatile("" ::: "memory");  // (sorted[] was written as u16, fill_table reads it as u32 pairs; cl[] likewise above)
  return fill_table(c, cold_next, sorted_w, root_v, rbits);
}

BD_COLD int read_huffman_code(const LaneCtx& c, Lane& L, uint32_t alphabet_size, uint32_t max_symbol, uint32_t root_v, uint32_t rbits) {
  BitWin b = L.win();
  uint32_t cold_next = L.cold_next;
  const int r = read_huffman_code_impl(c, b, cold_next, alphabet_size, max_symbol, root_v, rbits);
  L.put(b);
  L.cold_next = cold_next;
  return r;
}

#else
// Fill the lookup structure of one prefix code from its symbols sorted by (length, value).
// count[l] = symbols of length l.  Root of 2^rbits entries at root_v; longer codes go to second-level
// tables allocated from L.cold_next (always in the arena).  Same shape as BrotliBuildHuffmanTable (src/huffman/mod.rs:273-386).
BD_DEV int fill_table(const LaneCtx& c, Lane& L, const uint16_t* sorted, const uint16_t* count, uint32_t root_v, uint32_t rbits) {
  uint16_t rem[16];
  for (uint32_t l = 0; l < 16; l++) rem[l] = count[l];
  uint32_t code = 0, idx = 0;
  uint32_t cur_prefix = 0xFFFFFFFFu, sub_w = 0, sub_v = 0;
  for (uint32_t l = 1; l <= 15; l++) {
    for (uint32_t j = count[l]; j != 0; j--) {
      const uint32_t sym = sorted[idx++];
      const uint32_t rev = hw::brev(code) >> (32 - l);
      const uint32_t e = (sym << 4) | l;
      if (l <= rbits) {
        for (uint32_t t = rev; t < (1u << rbits); t += 1u << l) tab_store(c, root_v + t, e);
      } else {
        const uint32_t prefix = code >> (l - rbits);
        if (prefix != cur_prefix) {  // first code under a new root slot: size its second-level table (NextTableBitSize, :181-193)
          cur_prefix = prefix;
          int32_t left = 1 << (l - rbits);
          uint32_t ll = l;
          while (ll < 15) {
            left -= (int32_t)rem[ll];
            if (left <= 0) break;
            ll++; left <<= 1;
          }
          sub_w = ll - rbits;
          sub_v = L.cold_next - c.E;  // index into the arena part
          const uint32_t sub_size = sub_w < 2 ? 4u : 1u << sub_w;  // every arena allocation is a multiple of four entries
          if (sub_v + sub_size > kGlobalTab) return kLaneBail;
          L.cold_next += sub_size;
          tab_store(c, root_v + (rev & mask_bits(rbits)), ((sub_v >> 2) << 4) | (rbits + sub_w));
        }
        for (uint32_t t = rev >> rbits; t < (1u << sub_w); t += 1u << (l - rbits)) c.gtab[sub_v + t] = (uint16_t)e;
      }
      code++;
      rem[l]--;
    }
    code <<= 1;
  }
  return kLaneOk;
}

// ReadHuffmanCode, src/decode.rs:868-1013: one prefix-code description -> lookup structure.
BD_COLD int read_huffman_code(const LaneCtx& c, Lane& L, uint32_t alphabet_size, uint32_t max_symbol, uint32_t root_v, uint32_t rbits) {
  uint16_t sorted[704];
  uint16_t count[16];
  for (uint32_t l = 0; l < 16; l++) count[l] = 0;
  const uint32_t hskip = L.read(2);
  if (hskip == 1) {  // simple code: NSYM 1..4 explicit symbols (ReadSimpleHuffmanSymbols, :516-556)
    const uint32_t nsym = L.read(2) + 1;
    const uint32_t max_bits = bit_width(alphabet_size - 1);
    uint32_t s[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < nsym; i++) {
      s[i] = L.read(max_bits);
      if (s[i] >= max_symbol) return kLaneBail;
    }
    for (uint32_t i = 0; i + 1 < nsym; i++)
      for (uint32_t k = i + 1; k < nsym; k++) if (s[i] == s[k]) return kLaneBail;
    if (nsym == 1) {
      for (uint32_t t = 0; t < (1u << rbits); t++) tab_store(c, root_v + t, s[0] << 4);
      return kLaneOk;
    }
    // canonical codes sorted by (length, value) reproduce BrotliBuildSimpleHuffmanTable (src/huffman/mod.rs:390-471)
    uint32_t len[4] = {1, 1, 0, 0};
    if (nsym == 3) { len[1] = 2; len[2] = 2; }
    if (nsym == 4) {
      if (L.read(1)) { len[0] = 1; len[1] = 2; len[2] = 3; len[3] = 3; } else { len[0] = len[1] = len[2] = len[3] = 2; }
    }
    uint32_t n = 0;
    for (uint32_t l = 1; l <= 3; l++) {
      const uint32_t first = n;
      for (uint32_t i = 0; i < nsym; i++) if (len[i] == l) {
        uint32_t q = n++;
        while (q > first && sorted[q - 1] > s[i]) { sorted[q] = sorted[q - 1]; q--; }
        sorted[q] = (uint16_t)s[i];
        count[l]++;
      }
    }
    return fill_table(c, L, sorted, count, root_v, rbits);
  }
  // complex code: code-length code lengths (ReadCodeLengthCodeLengths, :801-853)
  uint8_t cl_cl[18];
  for (uint32_t i = 0; i < 18; i++) cl_cl[i] = 0;
  uint32_t space = 32, num_codes = 0;
  for (uint32_t i = hskip; i < 18; i++) {
    const uint32_t ix = L.peek() & 15u;
    // kCodeLengthPrefixLength / kCodeLengthPrefixValue (src/decode.rs:59-61) packed four bits per entry
    const uint32_t plen = (uint32_t)(0x4222322242223222ull >> (ix * 4)) & 15u;
    const uint32_t v = (uint32_t)(0x5340234013402340ull >> (ix * 4)) & 15u;
    L.skip(plen);
    cl_cl[tbl::kCodeLengthCodeOrder[i]] = (uint8_t)v;
    if (v != 0) {
      space -= 32u >> v;
      num_codes++;
      if (space - 1u >= 32u) break;  // space is 0 or wrapped
    }
  }
  if (!(num_codes == 1 || space == 0)) return kLaneBail;
  // 5-bit lookup of the code-length code (BrotliBuildCodeLengthsHuffmanTable, src/huffman/mod.rs:196-271)
  uint8_t cl_tab[32];  // symbol << 3 | length
  if (num_codes == 1) {
    uint32_t only = 0;
    for (uint32_t i = 0; i < 18; i++) if (cl_cl[i]) only = i;
    for (uint32_t t = 0; t < 32; t++) cl_tab[t] = (uint8_t)(only << 3);
  } else {
    uint32_t code = 0;
    for (uint32_t l = 1; l <= 5; l++) {
      for (uint32_t sy = 0; sy < 18; sy++) if (cl_cl[sy] == l) {
        const uint32_t rev = hw::brev(code) >> (32 - l);
        for (uint32_t t = rev; t < 32; t += 1u << l) cl_tab[t] = (uint8_t)((sy << 3) | l);
        code++;
      }
      code <<= 1;
    }
  }
  // symbol code lengths with repeat codes (ReadSymbolCodeLengths, :661-731; Process*CodeLength :565-658)
  uint8_t cl[704];
  uint32_t symbol = 0, prev_len = 8, repeat = 0, repeat_len = 0;
  space = 32768;
  if (max_symbol > 704) return kLaneBail;
  while (symbol < max_symbol && space > 0) {
    const uint32_t p = cl_tab[L.peek() & 31u];
    L.skip(p & 7u);
    const uint32_t code_len = p >> 3;
    if (code_len < 16) {
      repeat = 0;
      cl[symbol] = (uint8_t)code_len;
      if (code_len != 0) {
        prev_len = code_len;
        space -= 32768u >> code_len;
        count[code_len]++;
      }
      symbol++;
    } else {
      const uint32_t extra_bits = code_len - 14;
      uint32_t delta = L.read(extra_bits);
      const uint32_t new_len = code_len == 16 ? prev_len : 0;
      if (repeat_len != new_len) { repeat = 0; repeat_len = new_len; }
      const uint32_t old_repeat = repeat;
      if (repeat > 0) { repeat -= 2; repeat <<= extra_bits; }
      repeat += delta + 3;
      delta = repeat - old_repeat;
      if (symbol + delta > max_symbol) { space = 0xFFFFF; break; }
      for (uint32_t j = 0; j < delta; j++) cl[symbol + j] = (uint8_t)repeat_len;
      symbol += delta;
      if (repeat_len != 0) {
        space -= delta << (15 - repeat_len);
        count[repeat_len] = (uint16_t)(count[repeat_len] + delta);
      }
    }
  }
  if (space != 0) return kLaneBail;
  uint16_t offs[16];
  uint32_t o = 0;
  for (uint32_t l = 1; l <= 15; l++) { offs[l] = (uint16_t)o; o += count[l]; }
  for (uint32_t sy = 0; sy < symbol; sy++) {
    const uint32_t l = cl[sy];
    if (l) sorted[offs[l]++] = (uint16_t)sy;
  }
  return fill_table(c, L, sorted, count, root_v, rbits);
}

#endif
#endif

// A tree that lives wholly in the arena part of the table space (block-switch and context-map codes).
BD_DEV int read_arena_tree(const LaneCtx& c, Lane& L, uint32_t alphabet, uint32_t& root_v) {
  root_v = L.cold_next;
  if (root_v + (1u << kBlockRootBits) > c.E + kGlobalTab) return kLaneBail;
  L.cold_next += 1u << kBlockRootBits;
  return read_huffman_code(c, L, alphabet, alphabet, root_v, kBlockRootBits);
}

// ReadBlockLength, src/decode.rs:1016-1026
BD_DEV uint32_t read_block_length(const LaneCtx& c, Lane& L, uint32_t root_v) {
  const uint32_t code = decode_generic(c, L, root_v, kBlockRootBits);
  if (code >= 26) return 0;  // cannot happen for a 26-symbol alphabet
  return tbl::kBrotliBlockLengthOffset[code] + L.read(tbl::kBrotliBlockLengthNBits[code]);
}

// DecodeContextMap, src/decode.rs:1272-1428 (+ InverseMoveToFrontTransform :1096-1128)
BD_COLD int decode_context_map(const LaneCtx& c, Lane& L, uint32_t size, uint32_t& ntrees, uint8_t* map) {
  ntrees = read_varlen8(L) + 1;
  if (ntrees <= 1) {
    for (uint32_t i = 0; i < size; i++) map[i] = 0;
    return kLaneOk;
  }
  uint32_t rle_max = 0;
  const uint32_t b5 = L.peek() & 31u;
  if (b5 & 1u) { rle_max = (b5 >> 1) + 1; L.skip(5); } else { L.skip(1); }
  const uint32_t saved_cold = L.cold_next;
  uint32_t root_v;
  if (read_arena_tree(c, L, ntrees + rle_max, root_v) != kLaneOk) return kLaneBail;
  {
    BitWin b = L.win();  // (the byte stores to the map would otherwise send the bit window through local memory)
    uint32_t i = 0;
    while (i < size) {
      const uint32_t code = decode_generic(c, b, root_v, kBlockRootBits);
      if (code == 0) { map[i++] = 0; continue; }
      if (code > rle_max) { map[i++] = (uint8_t)(code - rle_max); continue; }
      uint32_t reps = (1u << code) + b.read(code);
      if (i + reps > size) return kLaneBail;
      do { map[i++] = 0; } while (--reps);
      if (b.overrun()) return kLaneBail;
    }
    L.put(b);
  }
  L.cold_next = saved_cold;  // the map's own code is not needed any more
  if (L.read(1)) {
    uint8_t mtf[256];
    for (uint32_t j = 0; j < 256; j++) mtf[j] = (uint8_t)j;
    for (uint32_t j = 0; j < size; j++) {
      uint32_t index = map[j];
      const uint8_t value = mtf[index];
      map[j] = value;
      for (; index > 0; index--) mtf[index] = mtf[index - 1];
      mtf[0] = value;
    }
  }
  return kLaneOk;
}

// Virtual root index of tree i of group g.
BD_DEV uint32_t tree_root(const Lane& L, uint32_t g, uint32_t i) { return L.root[g] + (i << L.rbits[g]); }

// Distance-context -> root of its tree for the current distance block type, kept in the shared slot.
// With the distance-tree cache (L.dist_home != 0: more distance trees than any one block type uses) only the trees of the
// current block type's four distance contexts are in the shared slot: L.dist_slots cache slots of one root each, filled
// here from the trees' homes in the arena (second-level pointers are arena indices, so a root can be copied anywhere).
// Contexts that share a tree share a slot -- the q5..q9 encoder gives every distance block type ONE tree for all four
// contexts, so the whole distance group usually costs a single root.
BD_DEV void refresh_cur_dist(const LaneCtx& c, const Lane& L) {
  uint32_t tr[4], slot_of[4];
  uint32_t used = 0;  // cache slots taken by this block type (at most L.dist_slots: the maximum over all block types)
  for (uint32_t ctx = 0; ctx < 4; ctx++) {
    const uint32_t t = c.ctx_dist[L.dist_slice + ctx];
    tr[ctx] = t;
    if (L.dist_home == 0) { sts32(c.slot + ctx * 4, tree_root(L, 2, t)); continue; }
    uint32_t same = ctx;
    for (uint32_t k = 0; k < ctx; k++) if (tr[k] == t && same == ctx) same = k;
    slot_of[ctx] = same != ctx ? slot_of[same] : used++;
    const uint32_t dst_v = L.root[2] + (slot_of[ctx] << L.rbits[2]);
    sts32(c.slot + ctx * 4, dst_v);
    if (same != ctx) continue;
    const uint16_t* src = c.gtab + (L.dist_home - c.E) + (t << L.rbits[2]);
    const uint32_t n = 1u << L.rbits[2];
    if (n >= 2) {
      for (uint32_t j = 0; j < n; j += 2) sts32(c.stab + ((dst_v + j) << 1), *(const uint32_t*)(src + j));  // (both sides 4-byte aligned)
    } else {
      sts16(c.stab + (dst_v << 1), src[0]);
    }
  }
}

// PrepareLiteralDecoding, src/decode.rs:1554-1570
BD_DEV void prepare_literal(const LaneCtx& c, Lane& L) {
  const uint32_t bt = L.rb[1];
  L.ctx_slice = bt << 6;
  L.trivial = ((bt < 32 ? L.trivial_lo >> bt : L.trivial_hi >> (bt - 32)) & 1u);
  L.lit_tree = c.ctx_lit[L.ctx_slice];
  L.ctx_mode_off = (uint32_t)(c.ctx_modes[bt] & 3u) * 512u;
  if (!L.trivial) {  // the command loop reads the block type's context map from the shared slot
    for (uint32_t j = 0; j < 64; j += 4) sts32(c.stab + 2 * L.e_tab + j, ld32(c.ctx_lit + L.ctx_slice + j));
  }
}

struct BlockTrees { uint32_t type_root[3], len_root[3]; };

// DecodeBlockTypeAndLength + Decode{Literal,Command,Distance}BlockSwitch, src/decode.rs:1469-1658.
// cat: 0 literal, 1 command, 2 distance.
BD_COLD int block_switch(const LaneCtx& c, Lane& L, uint32_t cat, const BlockTrees& trees) {
  // the roots are read here, not at the call site: the caller's loop must not touch local memory
  const uint32_t type_root = trees.type_root[cat], len_root = trees.len_root[cat];
  const uint32_t nbt = L.nbt[cat];
  if (nbt < 2) return kLaneBail;  // the counter of a single block type can only run out in a corrupt stream
  uint32_t bt = decode_generic(c, L, type_root, kBlockRootBits);
  L.bl[cat] = read_block_length(c, L, len_root);
  uint32_t* rb = &L.rb[2 * cat];
  if (bt == 1) bt = rb[1] + 1;
  else if (bt == 0) bt = rb[0];
  else bt -= 2;
  if (bt >= nbt) bt -= nbt;
  if (bt >= nbt) return kLaneBail;
  rb[0] = rb[1]; rb[1] = bt;
  if (cat == 0) prepare_literal(c, L);
  else if (cat == 1) L.cmd_tree = bt;
  else { L.dist_slice = bt << 2; refresh_cur_dist(c, L); }
  return L.overrun() ? kLaneBail : kLaneOk;
}


// Bit window at byte `byte_off` of the (aligned) input: what stream_begin does for the first byte.
BD_DEV void seek_byte(Lane& L, uint64_t byte_off) {
  L.k = (uint32_t)(byte_off >> 2);
  L.bp = 8 * (uint32_t)(byte_off & 3u);
  cp_async_wait_all();  // nothing older may still be landing in the ring
  const uint32_t b0 = L.k >> 2;
  cp_async16(L.ring + (b0 & 1u) * L.ring_stride, L.gin + 16 * (size_t)(b0 < L.last_blk ? b0 : L.last_blk));
  cp_async16(L.ring + ((b0 + 1) & 1u) * L.ring_stride, L.gin + 16 * (size_t)(b0 + 1 < L.last_blk ? b0 + 1 : L.last_blk));
  cp_async_commit();
  cp_async_wait_all();
  L.lo = vlds32(L.ring + ((L.k >> 2) & 1u) * L.ring_stride + (L.k & 3u) * 4u);
  L.hi = vlds32(L.ring + (((L.k + 1) >> 2) & 1u) * L.ring_stride + ((L.k + 1) & 3u) * 4u);
  L.nx = vlds32(L.ring + (((L.k + 2) >> 2) & 1u) * L.ring_stride + ((L.k + 2) & 3u) * 4u);
  if (((L.k + 2) >> 2) != b0) {  // the window already reaches into block b0 + 1: block b0 + 2 must be on its way (ring_next's invariant)
    const uint32_t b2 = b0 + 2;
    cp_async16(L.ring + (b2 & 1u) * L.ring_stride, L.gin + 16 * (size_t)(b2 < L.last_blk ? b2 : L.last_blk));
    cp_async_commit();
    cp_async_wait_all();
  }
}

// Bit window back at word `k`, bit `bp` of the (aligned) input (a position this lane has been at before).
BD_DEV void seek_bit(Lane& L, uint32_t k, uint32_t bp) {
  seek_byte(L, (uint64_t)k * 4);
  L.bp = bp;
}

// Byte-align the bit window (JumpToByteBoundary, src/bit_reader/mod.rs:378-385); false if the padding bits are not zero.
BD_DEV bool jump_to_byte_boundary(Lane& L) {
  const uint32_t pad = (8u - (L.bp & 7u)) & 7u;
  return pad == 0 || L.read(pad) == 0;
}

// ISUNCOMPRESSED metablock (CopyUncompressedBlockToOutput, src/decode.rs:1754-1806) for a well-formed stream: `n` raw bytes
// from the byte-aligned bit position to the output, through the write combiner (so that later copies and literal contexts
// see them in the history ring), then the bit window continues behind them.
BD_COLD int copy_raw(const LaneCtx& c, Lane& L, uint32_t n) {
  const uint64_t bo = ((uint64_t)L.k * 32 + L.bp) >> 3;  // byte offset from the aligned base
  if (bo + n > (L.end_bit >> 3)) return kLaneBail;       // truncated input: the exact kernel's business
  if (n > L.capb - L.posb || n > kMaxLaneRaw) return kLaneBail;
  uint32_t posb = L.posb, acc = L.acc;
  const uint32_t posb0 = posb;
  const uint8_t* src = L.gin + bo;
  uint32_t i = 0;
  // up to an aligned output word (and past the region's first word, whose bytes below the region are not ours)
  while (i < n && ((posb & 3u) != 0 || posb < 4)) {
    append(L.out_al, L.bias, c.hist, posb, acc, src[i], 1);
    i++;
#if BD_LANE_HEAD_PER_ROUND
    if (L.bias != 0 && posb0 < 4 && posb == 4) store_head_bytes(L.out_al, L.bias, 0, vlds32(c.hist));
#endif
  }
  // whole words: aligned input words re-aligned with a funnel shift (never past the word holding the last stream byte)
  const uint32_t* w = (const uint32_t*)L.gin;
  while (i + 4 <= n) {
    const uint64_t a = bo + i;
    const uint32_t j = (uint32_t)(a >> 2), sh = 8 * (uint32_t)(a & 3u);
    const uint32_t w0 = w[j], w1 = sh ? w[j + 1 <= L.k_max ? j + 1 : L.k_max] : 0u;
    append(L.out_al, L.bias, c.hist, posb, acc, hw::funnelshift_r(w0, w1, sh), 4);
    i += 4;
  }
  while (i < n) { append(L.out_al, L.bias, c.hist, posb, acc, src[i], 1); i++; }
  (void)posb0;
  L.posb = posb; L.acc = acc;
  seek_byte(L, bo + n);
  return kLaneNext;
}

// Metablock header up to the first command: src/decode.rs:2980-3288.
BD_COLD int metablock_begin(const LaneCtx& c, Lane& L, BlockTrees& bt) {
  const uint32_t is_last = L.read(1);
  L.is_last = is_last;
  L.mlen = 0;
  if (is_last && L.read(1)) return kLaneDone;  // ISLASTEMPTY
  const uint32_t nib = L.read(2);
  if (nib == 3) {  // metadata metablock: reserved bit, MSKIPBYTES, MSKIPLEN - 1, padding, then the bytes to skip (:3031-3045)
    if (L.read(1) != 0) return kLaneBail;
    const uint32_t nbytes = L.read(2);
    uint32_t skip = 0;
    for (uint32_t i = 0; i < nbytes; i++) {
      const uint32_t b = L.read(8);
      if (i + 1 == nbytes && nbytes > 1 && b == 0) return kLaneBail;
      skip |= b << (i * 8);
    }
    if (nbytes != 0) skip += 1;
    if (!jump_to_byte_boundary(L) || L.overrun()) return kLaneBail;
    const uint64_t bo = ((uint64_t)L.k * 32 + L.bp) >> 3;
    if (bo + skip > (L.end_bit >> 3)) return kLaneBail;
    if (skip != 0) seek_byte(L, bo + skip);
    return kLaneNext;
  }
  const uint32_t nn = nib + 4;
  uint32_t v = 0;
  for (uint32_t i = 0; i < nn; i++) {
    const uint32_t b = L.read(4);
    if (i + 1 == nn && nn > 4 && b == 0) return kLaneBail;
    v |= b << (i * 4);
  }
  if (!is_last && L.read(1)) {  // uncompressed metablock
    if (!jump_to_byte_boundary(L) || L.overrun()) return kLaneBail;
    return copy_raw(c, L, v + 1);
  }
  L.mlen = (int32_t)v + 1;
  L.cold_next = c.E;
  // block types and lengths per category (HUFFMAN_CODE_0..3, :3046-3140)
  for (uint32_t k = 0; k < 3; k++) {
    L.nbt[k] = read_varlen8(L) + 1;
    L.bl[k] = 1u << 24;
    L.rb[2 * k] = 1; L.rb[2 * k + 1] = 0;
    bt.type_root[k] = bt.len_root[k] = 0;
    if (L.nbt[k] >= 2) {
      if (L.nbt[k] > kMaxBlockTypes) return kLaneBail;
      if (read_arena_tree(c, L, L.nbt[k] + 2, bt.type_root[k]) != kLaneOk) return kLaneBail;
      if (read_arena_tree(c, L, 26, bt.len_root[k]) != kLaneOk) return kLaneBail;
      L.bl[k] = read_block_length(c, L, bt.len_root[k]);
    }
    if (L.overrun()) return kLaneBail;
  }
  const uint32_t pb = L.read(6);
  L.npostfix = pb & 3u;
  L.ndirect = 16u + ((pb >> 2) << L.npostfix);
  for (uint32_t i = 0; i < L.nbt[0]; i++) c.ctx_modes[i] = (uint8_t)L.read(2);  // ReadContextModes, :1991-2015
  if (decode_context_map(c, L, L.nbt[0] << 6, L.n_lit, c.ctx_lit) != kLaneOk) return kLaneBail;
  // DetectTrivialLiteralBlockTypes, :1525-1553
  L.trivial_lo = L.trivial_hi = 0;
  for (uint32_t i = 0; i < L.nbt[0]; i++) {
    const uint8_t* m = c.ctx_lit + (i << 6);
    uint32_t diff = 0;
    for (uint32_t j = 1; j < 64; j++) diff |= (uint32_t)(m[j] ^ m[0]);
    if (diff == 0) { if (i < 32) L.trivial_lo |= 1u << i; else L.trivial_hi |= 1u << (i - 32); }
  }
  if (decode_context_map(c, L, L.nbt[2] << 2, L.n_dist, c.ctx_dist) != kLaneOk) return kLaneBail;
  if (L.overrun()) return kLaneBail;
  L.dist_alphabet = L.ndirect + (48u << L.npostfix);  // 16 + NDIRECT + (24 << (NPOSTFIX + 1)), :3189-3194
  // context-modelled literals in this metablock: their block type's context map lives at the top of the slot
  {
    bool all_trivial = true;
    for (uint32_t i = 0; i < L.nbt[0]; i++) all_trivial = all_trivial && (((i < 32 ? L.trivial_lo >> i : L.trivial_hi >> (i - 32)) & 1u) != 0);
    if (!all_trivial && c.E < 2 * kCtxMapEntries) return kLaneBail;
    L.e_tab = all_trivial ? c.E : c.E - kCtxMapEntries;
  }
  // Root widths.  Groups (0 literal, 1 command, 2 distance) share the slot.  Their comfortable minimum is 4 / 4 / 3
  // bits; when that does not fit, the largest group is narrowed bit by bit -- down to 0 bits: one entry per tree, a
  // pointer to a second-level table that covers the whole code, which the command loop's look-ahead fetches
  // asynchronously like any other second-level entry -- so a group only leaves the slot (synchronous look-ups in the
  // arena) when it has more trees than the slot has entries.  Narrow roots make the second-level tables larger; if
  // they overflow the arena, the tree groups are read again with the roots in the arena (second attempt).
  // The groups that stay are then widened where it pays most: widening group g from R to R' costs
  // ntrees * (2^R' - 2^R) entries and saves the second-level look-ups of the symbols whose codes are R+1..R' bits
  // long.  Typical shares of symbols with codes longer than R (per cent; text and binary corpora)
  // stand in for the streams' own statistics.  Jumps over several widths are considered at once (the saving per
  // entry is not monotonic: the first bits of a root save nothing).
  // Distance-tree cache: a block type uses at most four distance trees (one per distance context) -- in practice one:
  // the encoder's distance context maps send all four contexts of a block type to the same tree.  When the metablock
  // has more trees than any block type uses, the shared slot only holds that many cache slots (see refresh_cur_dist)
  // and the group is sized -- and widened -- accordingly; all trees are built at their home in the arena.
  uint32_t dist_slots = 1;
  for (uint32_t b = 0; b < L.nbt[2]; b++) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < 4; i++) {
      bool seen = false;
      for (uint32_t j = 0; j < i; j++) seen = seen || c.ctx_dist[4 * b + j] == c.ctx_dist[4 * b + i];
      k += seen ? 0u : 1u;
    }
    if (k > dist_slots) dist_slots = k;
  }
  if (!BD_LANE_DIST_SLOTS_EXACT) dist_slots = 4;
  L.dist_slots = dist_slots;
  const bool dist_cache = BD_LANE_DIST_CACHE && L.n_dist > dist_slots;
  const uint32_t ntrees_all[3] = {L.n_lit, L.nbt[1], L.n_dist};
  const uint32_t ntrees[3] = {L.n_lit, L.nbt[1], dist_cache ? dist_slots : L.n_dist};
  const uint32_t rmin[3] = {4, 4, 3}, rmax[3] = {8, 8, 7};
  const uint32_t alpha[3] = {256, 704, L.dist_alphabet};
  const uint32_t save_lo = L.lo, save_hi = L.hi, save_nx = L.nx, save_k = L.k, save_bp = L.bp, save_cold = L.cold_next;
  for (uint32_t attempt = 0;; attempt++) {
    uint32_t rb[3] = {rmin[0], rmin[1], rmin[2]};
    bool shared[3] = {true, true, true};
    bool narrowed = false;
    if (attempt == 0 && BD_LANE_NARROW_ROOTS) {
      for (;;) {
        uint32_t total = 0, g = 3, best = 0;
        for (uint32_t i = 0; i < 3; i++) {
          const uint32_t sz = ntrees[i] << rb[i];
          total += sz;
          if (rb[i] != 0 && sz >= best) { best = sz; g = i; }
        }
        if (total <= L.e_tab || g == 3) break;
        rb[g]--;
        narrowed = true;
      }
    }
    for (;;) {
      uint32_t total = 0, g = 3, best = 0;
      for (uint32_t i = 0; i < 3; i++) if (shared[i]) {
        const uint32_t sz = ntrees[i] << rb[i];
        total += sz;
        if (sz >= best) { best = sz; g = i; }
      }
      if (total <= L.e_tab || g == 3) break;
      shared[g] = false;
#ifdef BD_LANE_SPILL_NARROW
      rb[g] = rmin[g];
#else
      rb[g] = (ntrees[g] << rmax[g]) <= kGlobalTab / 2 ? rmax[g] : rmin[g];  // the arena has room for wide roots
#endif
    }
    static const uint8_t kLongShare[3][9] = {{100, 100, 99, 97, 66, 36, 18, 9, 3},   // literal:  R = 0..8
                                             {100, 96, 78, 51, 37, 25, 16, 10, 6},   // command
                                             {100, 98, 92, 78, 32, 12, 5, 3, 1}};    // distance
    static const uint8_t kGroupWeight[3] = {BD_LANE_W_LIT, BD_LANE_W_CMD, BD_LANE_W_DIST};
#if BD_LANE_ROOTS_EXHAUSTIVE
    // all combinations of widths of the groups in the slot (at most 5^3): fewest expected second-level look-ups, then
    // fewest entries
    {
      uint32_t best_score = 0xFFFFFFFFu, best_size = 0, pick[3] = {rb[0], rb[1], rb[2]};
      const uint32_t lo0 = rb[0], lo1 = rb[1], lo2 = rb[2];
      const uint32_t hi0 = shared[0] ? rmax[0] : lo0, hi1 = shared[1] ? rmax[1] : lo1, hi2 = shared[2] ? rmax[2] : lo2;
      for (uint32_t r0 = lo0; r0 <= hi0; r0++) for (uint32_t r1 = lo1; r1 <= hi1; r1++) for (uint32_t r2 = lo2; r2 <= hi2; r2++) {
        const uint32_t size = (shared[0] ? ntrees[0] << r0 : 0u) + (shared[1] ? ntrees[1] << r1 : 0u) + (shared[2] ? ntrees[2] << r2 : 0u);
        if (size > L.e_tab) continue;
        const uint32_t score = (shared[0] ? kLongShare[0][r0] * kGroupWeight[0] : 0u) + (shared[1] ? kLongShare[1][r1] * kGroupWeight[1] : 0u) +
                               (shared[2] ? kLongShare[2][r2] * kGroupWeight[2] : 0u);
        if (score < best_score || (score == best_score && size < best_size)) { best_score = score; best_size = size; pick[0] = r0; pick[1] = r1; pick[2] = r2; }
      }
      rb[0] = pick[0]; rb[1] = pick[1]; rb[2] = pick[2];
    }
#else
    for (;;) {
      uint32_t total = 0, g = 3, to = 0;
      uint32_t best_num = 0, best_den = 1;  // benefit / cost of the best step, compared as fractions
      for (uint32_t i = 0; i < 3; i++) if (shared[i]) total += ntrees[i] << rb[i];
      for (uint32_t i = 0; i < 3; i++) if (shared[i]) {
        for (uint32_t r2 = rb[i] + 1; r2 <= rmax[i]; r2++) {
          const uint32_t extra = (ntrees[i] << r2) - (ntrees[i] << rb[i]);
          if (total + extra > L.e_tab) break;
          const uint32_t num = ((uint32_t)kLongShare[i][rb[i]] - (uint32_t)kLongShare[i][r2]) * kGroupWeight[i];
          if (num * best_den > best_num * extra) { best_num = num; best_den = extra; g = i; to = r2; }
        }
      }
      if (g == 3) break;
      rb[g] = to;
    }
#endif
    uint32_t next_shared = 0;
    const uint32_t order[3] = {1, 0, 2};
    bool fits = true;
    for (uint32_t oi = 0; oi < 3; oi++) {
      const uint32_t g = order[oi];
      const uint32_t sz = ntrees[g] << rb[g];
      if (shared[g]) {
        L.root[g] = next_shared;
        next_shared += sz;
      } else {
        if (L.cold_next + sz > c.E + kGlobalTab) { fits = false; break; }
        L.root[g] = L.cold_next;
        L.cold_next += sz;
      }
      L.rbits[g] = rb[g];
    }
    L.dist_home = 0;
    if (fits && dist_cache && shared[2]) {  // (a distance group that went to the arena as a whole needs no cache)
      const uint32_t sz = L.n_dist << rb[2];
      if (L.cold_next + sz > c.E + kGlobalTab) fits = false;
      else { L.dist_home = L.cold_next; L.cold_next += sz; }
    } else if (fits && dist_cache) {
      // the whole group is in the arena after all: it was placed as dist_slots trees, it has n_dist
      const uint32_t extra = (L.n_dist - dist_slots) << rb[2];
      if (L.root[2] + (dist_slots << rb[2]) != L.cold_next || L.cold_next + extra > c.E + kGlobalTab) fits = false;
      else L.cold_next += extra;
    }
    // HuffmanTreeGroupDecode x3, :1130-1219
    int r = fits ? kLaneOk : kLaneBail;
    for (uint32_t g = 0; g < 3 && r == kLaneOk; g++) {
      for (uint32_t i = 0; i < ntrees_all[g] && r == kLaneOk; i++) {
        const uint32_t rv = (g == 2 && L.dist_home != 0) ? L.dist_home + (i << L.rbits[2]) : tree_root(L, g, i);
        r = read_huffman_code(c, L, alpha[g], alpha[g], rv, L.rbits[g]);
        if (r == kLaneOk && L.overrun()) r = kLaneBail;
      }
    }
    if (r == kLaneOk) break;
    if (!narrowed) return kLaneBail;
    // narrow roots did not work out (most likely: second-level tables beyond the arena): once more, the classic way
    L.cold_next = save_cold;
    seek_bit(L, save_k, save_bp);
    if (L.lo != save_lo || L.hi != save_hi || L.nx != save_nx) return kLaneBail;  // (cannot happen: the input does not change)
  }
#ifdef BD_LANE_MB_STATS
  BD_LANE_MB_STATS(c, L);
#endif
  prepare_literal(c, L);
  L.cmd_tree = 0;
  L.dist_slice = 0;
  refresh_cur_dist(c, L);
  return kLaneOk;
}

// ======================= expanded static dictionary =======================
// Every (word, transform) pair of the RFC 7932 dictionary is materialised once per device, so a
// dictionary reference in the command loop is a plain word copy from a table instead of byte-serial
// transform logic (src/transform.rs:720-795) running with one active lane.  Entry of word `idx` of length
// L under transform t: xdict + xdict_base(L) + (idx * 121 + t) * xdict_stride(L), holding the
// prefix | transformed word | suffix bytes (at most L + 13), zero padded.
#if defined(BROTLI_B200_HOSTSIM)
#define BD_HD inline
#else
#define BD_HD __host__ __device__ __forceinline__
#endif
BD_HD uint32_t xdict_stride(uint32_t len) { return (len + 16u) & ~3u; }
// kBrotliDictSizeBitsByLength (src/dictionary/mod.rs:3-18) for lengths 4..24, four bits per entry, so that the
// layout is computable on the host as well (tests/hostsim checks it against the generated table)
BD_HD uint32_t dict_size_bits(uint32_t len) {
  if (len < 4 || len > 24) return 0;
  return len < 20 ? (uint32_t)(0x7877899AAAAABBAAull >> ((len - 4) * 4)) & 15u : (0x55667u >> ((len - 20) * 4)) & 15u;
}
struct XDictLayout {
  uint32_t base[25];
  uint32_t total;
};
BD_HD XDictLayout xdict_layout() {
  XDictLayout x;
  uint32_t off = 0;
  for (uint32_t l = 0; l < 25; l++) {
    x.base[l] = off;
    if (l >= BROTLI_MIN_DICTIONARY_WORD_LENGTH) off += (BROTLI_NUM_TRANSFORMS << dict_size_bits(l)) * xdict_stride(l);
  }
  x.total = off;
  return x;
}
// word info [len]: (xdict_base(len) / 4) << 4 | size bits;  transform info [t]: prefix length | suffix length << 4 |
// (bytes the transform cuts from the word: omit-first / omit-last count) << 8
BD_DEV uint32_t pack_word_info(const XDictLayout& x, uint32_t len) {
  return ((x.base[len] >> 2) << 4) | dict_size_bits(len);
}
BD_DEV uint32_t pack_transform_info(uint32_t t) {
  const uint8_t* prefix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[t * 3]];
  const uint8_t* suffix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[t * 3 + 2]];
  uint32_t plen = 0, slen = 0;
  while (prefix[plen]) plen++;
  while (suffix[slen]) slen++;
  const uint32_t type = tbl::kBrotliTransforms[t * 3 + 1];
  const uint32_t cut = type <= tbl::BROTLI_TRANSFORM_OMIT_LAST_9 ? type : (type < tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 ? 0u : type - (tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 - 1));
  return plen | (slen << 4) | (cut << 8);
}
// Bytes of the word itself that survive transform `type` (omit-first / omit-last).
BD_DEV uint32_t transformed_word_length(uint32_t len, uint32_t type) {
  uint32_t skip = type < tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 ? 0u : type - (tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 - 1);
  if (skip > len) skip = len;
  int32_t wl = (int32_t)(len - skip);
  if (type <= tbl::BROTLI_TRANSFORM_OMIT_LAST_9) wl -= (int32_t)type;
  return wl > 0 ? (uint32_t)wl : 0u;
}
// TransformDictionaryWord, src/transform.rs:743-795 (+ ToUpperCase :720-741): writes one table entry.
BD_DEV uint32_t build_xdict_entry(uint8_t* dst, const uint8_t* word, uint32_t len, uint32_t t) {
  const uint8_t* prefix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[t * 3]];
  const uint32_t type = tbl::kBrotliTransforms[t * 3 + 1];
  const uint8_t* suffix = &tbl::kBrotliPrefixSuffix[tbl::kBrotliTransforms[t * 3 + 2]];
  uint32_t n = 0;
  while (*prefix) dst[n++] = *prefix++;
  uint32_t skip = type < tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 ? 0u : type - (tbl::BROTLI_TRANSFORM_OMIT_FIRST_1 - 1);
  if (skip > len) skip = len;
  const uint32_t wlen = transformed_word_length(len, type);
  word += skip;
  uint32_t upper = (type == tbl::BROTLI_TRANSFORM_UPPERCASE_FIRST || type == tbl::BROTLI_TRANSFORM_UPPERCASE_ALL) ? 1u : 0u;
  uint32_t i = 0;
  while (i < wlen) {  // changes ToUpperCase would make past the word's end are never visible in the output
    const uint32_t c0 = word[i];
    if (c0 < 0xC0 || !upper) {
      dst[n++] = (uint8_t)((upper && c0 >= 'a' && c0 <= 'z') ? c0 ^ 32u : c0);
      i += 1;
    } else if (c0 < 0xE0) {
      dst[n++] = (uint8_t)c0;
      if (i + 1 < wlen) dst[n++] = word[i + 1] ^ 32u;
      i += 2;
    } else {
      dst[n++] = (uint8_t)c0;
      if (i + 1 < wlen) dst[n++] = word[i + 1];
      if (i + 2 < wlen) dst[n++] = word[i + 2] ^ 5u;
      i += 3;
    }
    if (type == tbl::BROTLI_TRANSFORM_UPPERCASE_FIRST) upper = 0;
  }
  while (*suffix) dst[n++] = *suffix++;
  for (uint32_t j = n; j < xdict_stride(len); j++) dst[j] = 0;
  return n;
}

// ======================= the command loop: phased rounds =======================
#ifndef BD_LANE_COUNT
#define BD_LANE_COUNT(i) ((void)0)  /* path statistics hook of the host build (tests/hostsim) */
#endif
#ifndef BD_LANE_DIST_STATS
#define BD_LANE_DIST_STATS(ud, len, dictword) ((void)0)  /* copy distance statistics hook of the host build */
#endif
#ifndef BD_LANE_LA_STATS
#define BD_LANE_LA_STATS(slot, in, two) ((void)0)  /* look-ahead statistics hook of the host build (profiles/hostsim_tables.py) */
#endif
enum : uint32_t { kStIdle = 0, kStHeader = 1, kStCommands = 2, kStFinish = 3, kStDone = 4, kStBail = 5 };
enum : uint32_t { kPhCmd = 0, kPhLit = 1, kPhDist = 2, kPhCopy = 3 };  // what a lane's stream needs next

// Restates ProcessCommandsInternal (src/decode.rs:2330-2744) for one metablock.  The WHOLE WARP calls
// this together; lanes with run == false only take part in the votes.  On return st is kStHeader
// (metablock complete) or kStBail for every lane that ran.
// kStride: distance between the two 16-byte blocks of a lane's block-interleaved input ring (16 x the CTA's
// threads), a compile-time constant of the kernel instance.
// kDict: the batch has a custom LZ77 dictionary (a separate kernel instance, so that streams without one pay nothing).
template <uint32_t kStride, bool kDict, bool kArena>
BD_DEV void run_commands(const LaneCtx& c, Lane& L, const BlockTrees& bt, bool run, uint32_t& st) {
  // (these shadow the namespace-scope defaults: see the latency configuration)
  constexpr bool kLatency = BD_LANE_IS_LATENCY_STRIDE(kStride);
  constexpr uint32_t kMaxExtraLiterals = kLatency ? BD_LANE_EXTRA_LITERALS_LAT : BD_LANE_EXTRA_LITERALS;
  constexpr uint32_t kCmdLiterals = kLatency ? BD_LANE_CMD_LITERALS_LAT : BD_LANE_CMD_LITERALS;
  constexpr uint32_t kBurst = kLatency ? BD_LANE_BURST_LAT : BD_LANE_BURST;
  constexpr uint32_t kBurstLanes = kLatency ? BD_LANE_BURST_LANES_LAT : BD_LANE_BURST_LANES;
  // register copies of the hot state
  const uint8_t* gin = nullptr;
  uint32_t lo = 0, hi = 0, nx = 0, k = 0, bp = 0, k_max = 0, last_blk = 0;
  hw::sref_t ring = c.ring;
  constexpr uint32_t ring_stride = kStride;
  uint8_t* out_al = nullptr;
  uint32_t bias = 0, capb = 0, posb = 0, acc = 0;
  int32_t mlen = 0;
  int32_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
  uint32_t bl_l = 0, bl_c = 0, bl_d = 0;
  uint32_t max_backward = 0, npostfix = 0, ndirect = 0;
  const uint8_t* cdict_end = nullptr;
  uint32_t cdict_size = 0;
  int32_t cdict_limit = 0;
  uint32_t r_lit = 0, r_cmd = 0, r_dist = 0, root_lit = 0;
  uint32_t cmd_tv = 0, lit_tv = 0, trivial = 0;  // trees of the current command / literal block type
  hw::sref_t ctx_lut = 0, ctx_map = 0;
  uint32_t E = c.E;
  hw::sref_t stab = c.stab;
  const uint16_t* gtab = c.gtab;
  hw::sref_t slot = c.slot;
  const uint8_t* xdict = c.xdict;
  hw::sref_t hist = c.hist, stage = c.stage;
#if defined(BROTLI_B200_HOSTSIM)
  const hw::sref_t stage_c = c.stage_c;
#else
  const hw::sref_t stage_c = ring + kStride * 7u;  // (ring = dynamic shared memory + 16 x thread; see the kernel's layout)
#endif
#define LN_STAGE(SLOT) ((SLOT) == 48u ? stage_c : stage + (SLOT))
#if defined(BROTLI_B200_HOSTSIM)
  const hw::sref_t cmd_lut = c.cmd_lut, word_info = c.word_info, transform_info = c.transform_info, ctx_lut_base = c.ctx_lut;
#else
  const hw::sref_t cmd_lut = hw::to_sref(g_cmd_lut), word_info = hw::to_sref(g_word_info), transform_info = hw::to_sref(g_transform_info);
  const hw::sref_t ctx_lut_base = hw::to_sref(g_ctx_lut);
#endif
  BD_PIN32(ring); BD_PIN32(E); BD_PIN32(stab); BD_PIN64(gtab); BD_PIN32(slot); BD_PIN64(xdict); BD_PIN32(hist); BD_PIN32(stage);

#define LN_TREES()                                                          \
  do {                                                                      \
    cmd_tv = tree_root(L, 1, L.cmd_tree);                                   \
    trivial = L.trivial;                                                    \
    lit_tv = tree_root(L, 0, L.lit_tree);                                   \
    ctx_lut = ctx_lut_base + L.ctx_mode_off;                                \
    ctx_map = stab + 2 * L.e_tab;                                           \
    BD_PIN32(cmd_tv); BD_PIN32(trivial); BD_PIN32(lit_tv); BD_PIN32(ctx_lut); BD_PIN32(ctx_map); \
  } while (0)

  if (run) {
    gin = L.gin; lo = L.lo; hi = L.hi; nx = L.nx; k = L.k; bp = L.bp; k_max = L.k_max; last_blk = L.last_blk;
    out_al = L.out_al; bias = L.bias; capb = L.capb; posb = L.posb; acc = L.acc;
    mlen = L.mlen;
    d0 = L.d0; d1 = L.d1; d2 = L.d2; d3 = L.d3;
    bl_l = L.bl[0]; bl_c = L.bl[1]; bl_d = L.bl[2];
    max_backward = L.max_backward; npostfix = L.npostfix; ndirect = L.ndirect;
    if (kDict) { cdict_end = L.cdict_end; cdict_size = L.cdict_size; cdict_limit = L.cdict_limit; }
    r_lit = L.rbits[0]; r_cmd = L.rbits[1]; r_dist = L.rbits[2]; root_lit = L.root[0];
    LN_TREES();
  }
  BD_PIN64(gin); BD_PIN32(k_max); BD_PIN32(last_blk); BD_PIN64(out_al); BD_PIN32(bias); BD_PIN32(capb);
  BD_PIN32(max_backward); BD_PIN32(npostfix); BD_PIN32(ndirect); BD_PIN32(r_lit); BD_PIN32(r_cmd); BD_PIN32(r_dist); BD_PIN32(root_lit);
  const bool ran = run;
  uint32_t klim = (((k + 2) >> 2) + 1) << 2;  // 4 x the highest input block requested so far (see LN_SKIP, LN_KLIM)
  uint32_t ph = kPhCmd;
  uint32_t ins = 0, copy_len = 0, cmd_bits = 0;
  uint32_t p1 = 0, p2 = 0;
  bool ctx_fresh = false;     // p1/p2 hold the last two output bytes (non-trivial literal contexts)
  // copy chunk in flight: up to 16 source bytes starting pend_off bytes into the two aligned 16-byte blocks
  // on their way (cp.async) to stage[0..31]
  uint32_t pend_n = 0, pend_off = 0;
  uint32_t cw0 = 0, cw1 = 0, cw2 = 0, cw3 = 0, cw4 = 0, cw5 = 0, cw6 = 0, cw7 = 0;  // BD_LANE_CHUNK_LDG: the chunk's two source blocks
  (void)cw0; (void)cw1; (void)cw2; (void)cw3; (void)cw4; (void)cw5; (void)cw6; (void)cw7;
  const uint8_t* csrc = nullptr;  // next source byte of the copy being made
  uint32_t crem = 0;              // its remaining bytes
  bool dhave = false;

#if BD_LANE_L2_HINTS && !defined(BROTLI_B200_HOSTSIM)
  uint64_t pol_stream = l2_policy_stream(), pol_keep = l2_policy_keep();
  BD_PIN64(pol_stream); BD_PIN64(pol_keep);
#if BD_LANE_L2_HINTS == 4 || BD_LANE_L2_HINTS == 5
#define LN_CP16_IF_KEEP(COND, DST, SRC) cp_async16_if(COND, DST, SRC)
#else
#define LN_CP16_IF_KEEP(COND, DST, SRC) cp_async16_if_hint(COND, DST, SRC, pol_keep)
#endif
#define LN_CP16_IF_STREAM(COND, DST, SRC) cp_async16_if_hint(COND, DST, SRC, pol_stream)
#if BD_LANE_L2_HINTS == 2
#define LN_CP16_IF_SRC(COND, DST, SRC) cp_async16_if(COND, DST, SRC)
#else
#define LN_CP16_IF_SRC(COND, DST, SRC) cp_async16_if_hint(COND, DST, SRC, pol_stream)
#endif
#else
#define LN_CP16_IF_KEEP(COND, DST, SRC) cp_async16_if(COND, DST, SRC)
#define LN_CP16_IF_STREAM(COND, DST, SRC) cp_async16_if(COND, DST, SRC)
#define LN_CP16_IF_SRC(COND, DST, SRC) cp_async16_if(COND, DST, SRC)
#endif
#define LN_PEEK() hw::funnelshift_r(lo, hi, bp)
#if BD_LANE_DEFER_A_STORES
#define LN_APPEND_A(V) append_defer(out_al, bias, hist, posb, acc, (V), 1, dw_pos, dw_word)
#define LN_FLUSH_A() do { st32_if(!(BD_LANE_PROBE_ABLATE & 32) && dw_pos != 0xFFFFFFFFu && dw_pos >= bias, out_al + dw_pos, dw_word); dw_pos = 0xFFFFFFFFu; } while (0)
#else
#define LN_APPEND_A(V) append(out_al, bias, hist, posb, acc, (V), 1)
#define LN_FLUSH_A() ((void)0)
#endif
#if BD_LANE_HEAD_PER_ROUND
// the region's first word was completed since position P0 (fewer than 32 bytes ago: it is still in the history ring)
#define LN_HEAD_CHECK(P0) do { if (BD_UNLIKELY(bias != 0 && (P0) < 4 && posb >= 4)) store_head_bytes(out_al, bias, 0, vlds32(hist)); } while (0)
#else
#define LN_HEAD_CHECK(P0) ((void)0)
#endif
#if BD_LANE_SKIP_LITE
// Bits consumed; on a word boundary the window takes word k + 3 from the ring (speculatively loaded: the load is
// unconditional, only its use depends on the boundary).  Input blocks are requested ONCE per round (LN_INPUT_BLOCK,
// after phase C1, where the round's bit position is final): the block after the one nx is in, so two blocks -- at
// least 97 bits -- are ahead of every round.  klim = 4 x the highest block requested: a lane that eats through all
// of it within one round (k > klim on a word boundary) fetches the next block on the spot.
#define LN_SKIP(n)                                                                               \
  do {                                                                                           \
    bp += (n);                                                                                   \
    const bool adv_ = bp >= 32;  /* branch-free word shift: nearly every round some lane needs it */ \
    if (BD_UNLIKELY(adv_ && k > klim)) {                                                         \
      const uint32_t b_ = (k + 3) >> 2;                                                          \
      cp_async16(ring + (b_ & 1u) * ring_stride, gin + 16 * (size_t)(b_ < last_blk ? b_ : last_blk)); \
      cp_async_commit(); cp_async_wait_all();                                                    \
      klim += 4;                                                                                 \
    }                                                                                            \
    const uint32_t j_ = k + 3;                                                                   \
    const uint32_t nw_ = vlds32(ring + ((j_ >> 2) & 1u) * ring_stride + (j_ & 3u) * 4u);         \
    k += adv_ ? 1u : 0u;                                                                         \
    lo = adv_ ? hi : lo; hi = adv_ ? nx : hi; nx = adv_ ? nw_ : nx;                              \
    bp &= 31u;                                                                                   \
  } while (0)
#define LN_KLIM() ((((k + 2) >> 2) + 1) << 2)
#define LN_INPUT_BLOCK(RUN)                                                                      \
  do {                                                                                           \
    const uint32_t nb_ = ((k + 2) >> 2) + 1;                                                     \
    const bool need_ = (RUN) && (nb_ << 2) > klim;                                               \
    LN_CP16_IF_STREAM(need_, ring + (nb_ & 1u) * ring_stride, gin + 16 * (size_t)(nb_ < last_blk ? nb_ : last_blk)); \
    klim = need_ ? nb_ << 2 : klim;                                                              \
  } while (0)
/* the per-metablock bit reader expects the block after nx's to be there */
#define LN_BLOCK_SWITCH_RING()                                                        \
  do {                                                                                \
    const uint32_t nb_ = ((k + 2) >> 2) + 1;                                          \
    if ((nb_ << 2) > klim) {                                                          \
      cp_async16(ring + (nb_ & 1u) * ring_stride, gin + 16 * (size_t)(nb_ < last_blk ? nb_ : last_blk)); \
      cp_async_commit(); cp_async_wait_all();                                         \
    }                                                                                 \
  } while (0)
#else
// Bits consumed; on a word boundary the window takes word k + 2 from the ring.  When that word starts a new
// 16-byte block, the block after it is requested with cp.async -- not committed here: the round's convergent
// commit covers it, and the convergent wait at the start of the next round completes it long before its words
// are needed.  Only a lane that eats more than a whole block within one round has to commit and wait itself.
#define LN_SKIP(n)                                                                               \
  do {                                                                                           \
    bp += (n);                                                                                   \
    const bool adv_ = bp >= 32;  /* branch-free word shift: nearly every round some lane needs it */ \
    k += adv_ ? 1u : 0u;                                                                         \
    const uint32_t j_ = k + 2;                                                                   \
    const bool blk_ = adv_ && (j_ & 3u) == 0;                                                    \
    if (BD_UNLIKELY(blk_ && blk_seen)) { cp_async_commit(); cp_async_wait_all(); }               \
    const uint32_t b_ = (j_ >> 2) + 1;                                                           \
    LN_CP16_IF_STREAM(blk_, ring + (b_ & 1u) * ring_stride, gin + 16 * (size_t)(b_ < last_blk ? b_ : last_blk)); \
    blk_seen = blk_seen || blk_;                                                                 \
    const uint32_t nw_ = vlds32(ring + ((j_ >> 2) & 1u) * ring_stride + (j_ & 3u) * 4u);         \
    lo = adv_ ? hi : lo; hi = adv_ ? nx : hi; nx = adv_ ? nw_ : nx;                              \
    bp &= 31u;                                                                                   \
  } while (0)
#define LN_KLIM() 0u
#define LN_INPUT_BLOCK(RUN) ((void)0)
#define LN_BLOCK_SWITCH_RING() ((void)0)
#endif
#define LN_SAVE()                                                                                   \
  do {                                                                                              \
    L.lo = lo; L.hi = hi; L.nx = nx; L.k = k; L.bp = bp; L.posb = posb; L.acc = acc; L.mlen = mlen;  \
    L.d0 = d0; L.d1 = d1; L.d2 = d2; L.d3 = d3; L.bl[0] = bl_l; L.bl[1] = bl_c; L.bl[2] = bl_d;     \
  } while (0)
// block switch of category CAT (0 literal, 1 command, 2 distance); rare, so the state takes a round trip
// through local memory (:2367-2372, :2413-2424, :2567-2571)
#define LN_BLOCK_SWITCH(CAT)                                                          \
  do {                                                                                \
    LN_BLOCK_SWITCH_RING();                                                           \
    LN_SAVE();                                                                        \
    const int r_ = block_switch(c, L, CAT, bt);                                       \
    lo = L.lo; hi = L.hi; nx = L.nx; k = L.k; bp = L.bp;                                       \
    klim = LN_KLIM();  /* (the per-metablock bit reader keeps the block after nx's requested) */ \
    bl_l = L.bl[0]; bl_c = L.bl[1]; bl_d = L.bl[2];                                   \
    LN_TREES();                                                                       \
    ctx_fresh = false;                                                                \
    if (r_ != kLaneOk) ev = kStBail;                                                  \
  } while (0)
// One symbol of the tree rooted at TV (root width TR): sets BITS (the 32-bit peek), LEN and SYM, or -- when
// the code is longer than the root -- requests the second-level entry and sets WAIT: the lane retries
// this phase next round with the entry in `de` (DecodeSymbol, :377-391, over our table shape).
#define LN_DECODE(TV, TR, BITS, LEN, SYM)                                                        \
  do {                                                                                           \
    BITS = LN_PEEK();                                                                            \
    const uint32_t v_ = (TV) + low_bits(BITS, TR);                                           \
    uint32_t e_ = vlds16(stab + ((v_ < E ? v_ : 0u) << 1));                                      \
    ld16_if(v_ >= E, gtab + (v_ - E), e_);  /* root outside the shared slot */                    \
    const bool need2_ = (e_ & 15u) > (TR);                                                       \
    const uint32_t sub_ = need2_ ? (e_ & 15u) - (TR) : 0u;                                       \
    ld16_if(need2_, gtab + ((e_ >> 4) << 2) + low_bits(BITS >> (TR), sub_), e_);           \
    LEN = e_ & 15u; SYM = e_ >> 4;                                                               \
  } while (0)
// request the next copy chunk: min(crem, 16) bytes from csrc on; only 16-byte blocks that hold source bytes
// are touched
#define LN_ISSUE_CHUNK(ISS)                                                \
  do {                                                                     \
    const bool iss_ = (ISS);                                               \
    const uint32_t nn_ = crem < 16 ? crem : 16u;                           \
    const uint32_t off_ = (uint32_t)(uintptr_t)csrc & 15u;                 \
    if (BD_LANE_CHUNK_LDG) {                                                                               \
      ldg128_if(iss_, csrc - off_, cw0, cw1, cw2, cw3);                                                    \
      ldg128_if(iss_ && off_ + nn_ > 16, csrc - off_ + 16, cw4, cw5, cw6, cw7);                            \
    } else {                                                                                               \
    LN_CP16_IF_SRC(iss_ && !(BD_LANE_PROBE_ABLATE & 2), stage, csrc - off_);                              \
    LN_CP16_IF_SRC(iss_ && !(BD_LANE_PROBE_ABLATE & 2) && off_ + nn_ > 16, stage + 16, csrc - off_ + 16); \
    }                                                                                                      \
    if (BD_LANE_PROBE_ABLATE & 8) { probe_ldg16_if(iss_, csrc - off_); probe_ldg16_if(iss_ && off_ + nn_ > 16, csrc - off_ + 16); } \
    if (iss_) { pend_n = nn_; pend_off = off_; crem -= nn_; csrc += 16; }  \
  } while (0)
// append the chunk in flight to the output
#define LN_RETIRE_CHUNK()                                                  \
  do {                                                                     \
    if (BD_LANE_CHUNK_LDG) { sts128(stage, cw0, cw1, cw2, cw3); sts128(stage + 16, cw4, cw5, cw6, cw7); } \
    const hw::sref_t sw_ = stage + (pend_off & 12u);                       \
    const uint32_t s8_ = (pend_off & 3u) * 8u;                             \
    const uint32_t pw0_ = vlds32(sw_), pw1_ = vlds32(sw_ + 4), pw2_ = vlds32(sw_ + 8); \
    uint32_t v0_ = hw::funnelshift_r(pw0_, pw1_, s8_);                     \
    uint32_t v1_ = hw::funnelshift_r(pw1_, pw2_, s8_);                     \
    /* byte masks without branches: low_bits keeps everything from 32 bits on and nothing at 0 bits */ \
    const uint32_t b0_ = 8u * (pend_n < 8 ? pend_n : 8u);                  \
    v0_ = low_bits(v0_, b0_);                                              \
    v1_ = low_bits(v1_, b0_ - (b0_ < 32 ? b0_ : 32u));                     \
    append8(out_al, bias, hist, posb, acc, v0_, v1_, b0_ >> 3);            \
    if (pend_n > 8) {                                                      \
      BD_LANE_COUNT(3);                                                    \
      const uint32_t pw3_ = vlds32(sw_ + 12), pw4_ = vlds32(sw_ + 16);     \
      uint32_t v2_ = hw::funnelshift_r(pw2_, pw3_, s8_);                   \
      uint32_t v3_ = hw::funnelshift_r(pw3_, pw4_, s8_);                   \
      const uint32_t b1_ = 8u * (pend_n - 8);                              \
      v2_ = low_bits(v2_, b1_);                                            \
      v3_ = low_bits(v3_, b1_ - (b1_ < 32 ? b1_ : 32u));                   \
      append8(out_al, bias, hist, posb, acc, v2_, v3_, b1_ >> 3);          \
    }                                                                      \
    pend_n = 0;                                                            \
  } while (0)

  // Look-ahead for a symbol that will be decoded a phase later: the root look-up is done now and, when the code
  // is longer than the root, its second-level entry (the aligned 16-byte block holding it) is requested with
  // cp.async into stage[SLOT..SLOT+15].  PV/PE/PSEL: valid flag, entry (bit 31: take it from the stage), byte
  // offset in the block.  In the kArena instance of the loop -- taken by a warp when one of its lanes has a tree
  // group in the arena (it did not fit the shared slot) -- a root in the arena has its ROOT entry requested the same
  // way (bits 31 and 30 of PE); the few codes longer than such a (wide) root then cost one synchronous load in
  // LN_TAKE instead of two dependent ones for every symbol.  Warps without such a lane run the instance without
  // these instructions (they are on the round's critical path: -1.8 % on the headline batch when always there).
#define LN_LOOKAHEAD(TV, TR, SLOT, PV, PE, PSEL)                                                 \
  do {                                                                                           \
    const uint32_t bits_ = LN_PEEK();                                                            \
    const uint32_t v_ = (TV) + low_bits(bits_, TR);                                          \
    const bool in_ = v_ < E;                                                                     \
    const uint32_t e_ = vlds16(stab + ((in_ ? v_ : 0u) << 1));                                   \
    const bool two_ = in_ && (e_ & 15u) > (TR);                                                  \
    const uint32_t sub_ = two_ ? (e_ & 15u) - (TR) : 0u;                                         \
    const bool ar_ = kArena && !in_;  /* root in the arena: request the root entry */            \
    const uint32_t i2_ = ar_ ? v_ - E : ((e_ >> 4) << 2) + low_bits(bits_ >> (TR), sub_);  \
    BD_LANE_LA_STATS(SLOT, in_, two_);                                                           \
    LN_CP16_IF_KEEP(two_ || ar_, LN_STAGE(SLOT), gtab + (i2_ & ~7u));                            \
    PV = kArena || in_; PE = ar_ ? 0xC0000000u : (two_ ? 0x80000000u : e_); PSEL = (two_ || ar_) ? (i2_ & 7u) << 1 : PSEL; \
  } while (0)
// entry of a looked-ahead symbol (its group has been waited for)
#define LN_TAKE(SLOT, PV, PE, PSEL, TR, BITS, LEN, SYM)                                          \
  do {                                                                                           \
    BITS = LN_PEEK();                                                                            \
    const uint32_t es_ = vlds16(LN_STAGE(SLOT) + PSEL);  /* unconditional: no branch */          \
    uint32_t e_ = (PE & 0x80000000u) ? es_ : PE;                                                 \
    if (kArena && BD_UNLIKELY((PE & 0x40000000u) != 0 && (e_ & 15u) > (TR))) {  /* arena root, long code */ \
      BD_LANE_COUNT(9);                                                                          \
      e_ = gtab[((e_ >> 4) << 2) + low_bits(BITS >> (TR), (e_ & 15u) - (TR))];                   \
    }                                                                                            \
    PV = false;                                                                                  \
    LEN = e_ & 15u; SYM = e_ >> 4;                                                               \
  } while (0)

  // The same look-ahead with the entry loaded into a register: the root look-up happens here (inside phase A's branch),
  // the load itself (PIDX: arena index, PNEED) is issued by LN_LOOKAHEAD_REG_LOAD at the top level of the round -- a
  // load issued inside a branch is waited for where the branch rejoins.
#define LN_LOOKAHEAD_REG(TV, TR, PV, PE, PIDX, PNEED)                                            \
  do {                                                                                           \
    const uint32_t bits_ = LN_PEEK();                                                            \
    const uint32_t v_ = (TV) + low_bits(bits_, TR);                                              \
    const bool in_ = v_ < E;                                                                     \
    const uint32_t e_ = vlds16(stab + ((in_ ? v_ : 0u) << 1));                                   \
    const bool two_ = in_ && (e_ & 15u) > (TR);                                                  \
    const uint32_t sub_ = two_ ? (e_ & 15u) - (TR) : 0u;                                         \
    const bool ar_ = kArena && !in_;                                                             \
    PIDX = ar_ ? v_ - E : ((e_ >> 4) << 2) + low_bits(bits_ >> (TR), sub_);                      \
    BD_LANE_LA_STATS(48u, in_, two_);                                                            \
    PNEED = two_ || ar_;                                                                         \
    PV = kArena || in_; PE = ar_ ? 0xC0000000u : (two_ ? 0x80000000u : e_);                      \
  } while (0)
#define LN_TAKE_REG(PLD, PV, PE, TR, BITS, LEN, SYM)                                             \
  do {                                                                                           \
    BITS = LN_PEEK();                                                                            \
    uint32_t e_ = (PE & 0x80000000u) ? PLD : PE;                                                 \
    if (kArena && BD_UNLIKELY((PE & 0x40000000u) != 0 && (e_ & 15u) > (TR))) {  /* arena root, long code */ \
      BD_LANE_COUNT(9);                                                                          \
      e_ = gtab[((e_ >> 4) << 2) + low_bits(BITS >> (TR), (e_ & 15u) - (TR))];                   \
    }                                                                                            \
    PV = false;                                                                                  \
    LEN = e_ & 15u; SYM = e_ >> 4;                                                               \
  } while (0)

  // look-ahead state: pa_* for the symbol phase A will decode next, pc_* for the distance symbol of phase C
  bool pa_valid = false, pc_valid = false;
  uint32_t pa_e = 0, pa_sel = 0, pc_e = 0, pc_sel = 0;
#if BD_LANE_DIST_LA_REG
  uint32_t pc_ld = 0;  // second-level (or arena root) entry of the distance symbol, loaded by phase A's look-ahead
#endif
#if BD_LANE_NEXT_LA_REG
  uint32_t pa_ld = 0;  // the same for the next phase-A symbol
#endif

#if BD_LANE_WAIT_HIST && !defined(BROTLI_B200_HOSTSIM)
  long long round_t0_ = 0;
#endif
  while (warp_any(run)) {
    uint32_t ev = kStCommands;  // kStHeader: metablock complete; kStBail: give the stream up
#if !BD_LANE_SKIP_LITE
    bool blk_seen = false;       // this lane has requested an input block in this round (see LN_SKIP)
#endif
    bool lit_fresh = false;      // this lane's literal run was announced by a command read in this round
#if BD_LANE_HEAD_PER_ROUND
    const uint32_t posb0 = posb;
#endif
    uint32_t lit_pack = 0, lit_n = 0;  // literals decoded in phase A of their command's round, appended in phase P
    uint32_t dw_pos = 0xFFFFFFFFu, dw_word = 0;  // output word completed by phase A's literals, stored in phase P (BD_LANE_DEFER_A_STORES)
#if BD_LANE_DIST_LA_REG
    uint32_t pc_idx = 0;    // arena index of the entry the distance look-ahead loads
    bool pc_need = false;
#endif
#if BD_LANE_NEXT_LA_REG
    uint32_t pa_idx = 0;
    bool pa_need = false;
#endif
    // groups still pending here: [next-A look-ahead, copy chunk] of the previous round; phase A needs the first
#if BD_LANE_WAIT_HIST && !defined(BROTLI_B200_HOSTSIM)
    { const long long now_ = clock64(); if (round_t0_ != 0) wait_hist_add(3, now_ - round_t0_); round_t0_ = now_; }
#endif
    { LN_WAIT_T0(); cp_async_wait_all_but_latest(); LN_WAIT_T1(0); }
#ifdef BD_LANE_ROUND_STATS
    BD_LANE_ROUND_STATS(ph, pa_valid);
#endif

    // ---- phase A: one symbol for every lane at a command boundary (insert&copy command and its extra bits,
    //      ReadCommandInternal :2134-2189) or inside a literal run (:2391-2551): the table look-up and the bit
    //      skip are shared, only the short tails differ ----
    if (run && (ph == kPhCmd || ph == kPhLit)) {
      const bool is_lit = ph == kPhLit;
      if (BD_UNLIKELY((is_lit ? bl_l : bl_c) == 0)) { LN_BLOCK_SWITCH(is_lit ? 0u : 1u); pa_valid = false; }
      if (ev == kStCommands) {
        uint32_t bits, len, sym;
        if (BD_LIKELY(pa_valid)) {
#if BD_LANE_NEXT_LA_REG
          LN_TAKE_REG(pa_ld, pa_valid, pa_e, (is_lit ? r_lit : r_cmd), bits, len, sym);
#else
          LN_TAKE(32u, pa_valid, pa_e, pa_sel, (is_lit ? r_lit : r_cmd), bits, len, sym);
#endif
        } else {
          uint32_t tv = is_lit ? lit_tv : cmd_tv;
          const uint32_t tr = is_lit ? r_lit : r_cmd;
          if (is_lit && !trivial) {  // tree by the context of the last two bytes (:2500-2507)
            if (!ctx_fresh) { last_two(hist, bias, posb, acc, p1, p2); ctx_fresh = true; if (kDict && posb - bias < 2) { p1 = 0; p2 = 0; } }
            const uint32_t cx = vlds8(ctx_lut + p1) | vlds8(ctx_lut + 256 + p2);
            tv = root_lit + (vlds8(ctx_map + cx) << r_lit);
          }
          BD_LANE_COUNT(6);
          LN_DECODE(tv, tr, bits, len, sym);
#ifdef BD_LANE_FALLBACK_STATS
          BD_LANE_FALLBACK_STATS(0, is_lit);
#endif
        }
        uint32_t nskip = len;
        if (is_lit) {
          bl_l--;
          LN_APPEND_A(sym);
          p2 = p1; p1 = sym;
          --ins;
          // Further literals of the run in the same round while they cost no trip to the arena: one tree for all
          // contexts, its root in the shared slot, the code no longer than the root and inside the 32-bit peek.
          if (trivial) {
            for (uint32_t rep = 0; rep < kMaxExtraLiterals; rep++) {
              if (ins == 0 || bl_l == 0 || nskip + r_lit > 32) break;
              const uint32_t v2 = lit_tv + low_bits(bits >> nskip, r_lit);
              const uint32_t e2 = vlds16(stab + ((v2 < E ? v2 : 0u) << 1));
              if (v2 >= E || (e2 & 15u) > r_lit) break;
              bl_l--;
              LN_APPEND_A(e2 >> 4);
              --ins;
              nskip += e2 & 15u;
            }
          }
          if (ins == 0) {
            if (mlen <= 0) ev = kStHeader;  // a trailing insert without a copy ends the metablock (:2552-2556)
            else ph = kPhDist;
          }
        } else {
          const uint2 lut = vlds64(cmd_lut + (sym << 3));
          cmd_bits = lut.x;
          ins = lut.x & 0xFFFFu;
          copy_len = lut.y & 0xFFFFu;
          const uint32_t ie = (lut.x >> 16) & 0xFFu, ce = lut.y >> 16;
          if (BD_LIKELY(len + ie + ce <= 32)) {  // symbol and both extra fields from the one 32-bit peek
            const uint32_t x = bits >> len;
            ins += low_bits(x, ie);
            copy_len += low_bits(x >> ie, ce);
            nskip = len + ie + ce;
          } else {
            BD_LANE_COUNT(4);
            LN_SKIP(len);
            if (ie) { ins += low_bits(LN_PEEK(), ie); LN_SKIP(ie); }
            bits = LN_PEEK();  // (the literals below continue from this peek)
            copy_len += low_bits(bits, ce);
            nskip = ce;
          }
          bl_c--;
          mlen -= (int32_t)ins;
          ctx_fresh = false;
          ph = kPhDist;
          if (ins != 0) {
            // The command's first literals are decoded in the command's own round, while they come out of the same
            // 32-bit peek: one tree for all contexts, root in the shared slot, codes no longer than the root (a short
            // insert -- 98 % are <= 4 -- then costs no round of its own).  They are appended in phase P, behind the
            // previous command's copy chunk, which is still in flight.  The bounds phase P checks for a literal run
            // are checked here first.
            if (kCmdLiterals != 0 && trivial && mlen >= 0 && ins + pend_n <= capb - posb) {
              for (uint32_t rep = 0; rep < kCmdLiterals; rep++) {
                if (ins == 0 || bl_l == 0 || nskip + r_lit > 32) break;
                const uint32_t v2 = lit_tv + low_bits(bits >> nskip, r_lit);
                const uint32_t e2 = vlds16(stab + ((v2 < E ? v2 : 0u) << 1));
                if (v2 >= E || (e2 & 15u) > r_lit) break;
                bl_l--;
                lit_pack |= (e2 >> 4) << (8 * lit_n);
                lit_n++;
                --ins;
                nskip += e2 & 15u;
              }
            }
            if (ins != 0) { ph = kPhLit; lit_fresh = true; }
            else if (mlen <= 0) ev = kStHeader;  // a trailing insert without a copy ends the metablock (:2552-2556)
          }
        }
        LN_SKIP(nskip);
        // look ahead for the distance symbol this lane decodes in phase C of this round
        if (ph == kPhDist && ev == kStCommands && !(cmd_bits & (1u << 26)) && bl_d != 0) {
          const uint32_t tvd = vlds32(slot + ((cmd_bits >> 24) & 3u) * 4u);
#if BD_LANE_DIST_LA_REG
          LN_LOOKAHEAD_REG(tvd, r_dist, pc_valid, pc_e, pc_idx, pc_need);
#else
          LN_LOOKAHEAD(tvd, r_dist, 48u, pc_valid, pc_e, pc_sel);
#endif
        }
      }
    }
    // ---- literal burst (see kBurst): more literals for the lanes inside a run while enough of the warp is ----
    if (kBurst != 0) {
      for (uint32_t b = 0; b < kBurst; b++) {
        // (a lane whose run was only just announced by its command waits for phase P's bounds check of this round)
        const bool lit = run && ev == kStCommands && ph == kPhLit && ins != 0 && bl_l != 0 && !lit_fresh;
        if (warp_count(lit) < kBurstLanes) break;
        if (lit) {
          uint32_t tv = lit_tv;
          if (!trivial) {
            if (!ctx_fresh) { last_two(hist, bias, posb, acc, p1, p2); ctx_fresh = true; if (kDict && posb - bias < 2) { p1 = 0; p2 = 0; } }
            const uint32_t cx = vlds8(ctx_lut + p1) | vlds8(ctx_lut + 256 + p2);
            tv = root_lit + (vlds8(ctx_map + cx) << r_lit);
          }
          uint32_t bits, len, sym;
          LN_DECODE(tv, r_lit, bits, len, sym);
          (void)bits;
          bl_l--;
          append(out_al, bias, hist, posb, acc, sym, 1);
          p2 = p1; p1 = sym;
          --ins;
          LN_SKIP(len);
          pa_valid = false;
          if (ins == 0) {
            if (mlen <= 0) ev = kStHeader;  // a trailing insert without a copy ends the metablock (:2552-2556)
            else ph = kPhDist;
          }
        }
      }
    }
    warp_sync();
    // ---- phase P: retire the copy chunk requested at the end of the previous round ----
#define LN_PHASE_P()                                                                                       \
  do {                                                                                                     \
    LN_FLUSH_A();  /* (dw_pos is only ever set by a running lane) */                                       \
    if (run && pend_n != 0) LN_RETIRE_CHUNK();                                                             \
    if (kCmdLiterals != 0 && run) {  /* literals decoded in their command's round (phase A) */             \
      append(out_al, bias, hist, posb, acc, lit_pack, lit_n);                                              \
    }                                                                                                      \
    if (BD_LANE_HEAD_PER_ROUND && run) LN_HEAD_CHECK(posb0);  /* phases A and P append at most 19 bytes */ \
    /* overshooting the metablock (BLOCK_LENGTH) or the output region: the exact decoder's business */     \
    if (run && ph == kPhLit && BD_UNLIKELY(mlen < 0 || ins > capb - posb)) ev = kStBail;                   \
  } while (0)
#if BD_LANE_DIST_LA_REG
    // the distance look-ahead's load, at the top level of the round (see LN_LOOKAHEAD_REG)
    ld16_if(pc_need, gtab + pc_idx, pc_ld);
#endif
#if !BD_LANE_LATE_RETIRE
    cp_async_commit();  // group: what phase A requested (distance look-ahead, input blocks)
#if BD_LANE_WAIT_ALL_AT_P
    cp_async_wait_all();  // experiment: nothing in flight while phase P stores
#else
    { LN_WAIT_T0(); cp_async_wait_all_but_latest(); LN_WAIT_T1(1); }  // the chunk (and everything older); phase A's requests stay in flight
#endif
    LN_PHASE_P();
    warp_sync();
    { LN_WAIT_T0(); cp_async_wait_all(); LN_WAIT_T1(2); }  // the distance look-ahead
#endif

    // ---- phase C1: distance symbol (ReadDistanceInternal :2066-2131, TakeDistanceFromRingBuffer :2017-2049) ----
    int32_t dist = d0;
    uint32_t push = 0;
    const bool go = run && ph == kPhDist && ev == kStCommands;  // this lane makes its copy in this round
    if (go && !(cmd_bits & (1u << 26))) {  // explicit distance symbol
      if (BD_UNLIKELY(bl_d == 0)) { LN_BLOCK_SWITCH(2); pc_valid = false; }
      if (ev == kStCommands) {
        uint32_t bits, len, sym;
        if (BD_LIKELY(pc_valid)) {
#if BD_LANE_DIST_LA_REG
          LN_TAKE_REG(pc_ld, pc_valid, pc_e, r_dist, bits, len, sym);
#else
          LN_TAKE(48u, pc_valid, pc_e, pc_sel, r_dist, bits, len, sym);
#endif
        } else {
          const uint32_t tv = vlds32(slot + ((cmd_bits >> 24) & 3u) * 4u);
          BD_LANE_COUNT(7);
          LN_DECODE(tv, r_dist, bits, len, sym);
#ifdef BD_LANE_FALLBACK_STATS
          BD_LANE_FALLBACK_STATS(1, 0);
#endif
        }
        bl_d--;
        // one straight-line computation for both kinds of distance symbol, one bit skip
        const bool longcode = sym >= 16;
        const bool direct = sym < ndirect;
        const uint32_t distval = sym - ndirect;
        const uint32_t hcode = distval >> npostfix;
        const uint32_t nbits = (longcode && !direct) ? (hcode >> 1) + 1 : 0u;
        const uint32_t base = direct ? sym - 15u
                                     : ((((2u + (hcode & 1u)) << nbits) - 4u) << npostfix) + low_bits(distval, npostfix) + ndirect - 15u;
        uint32_t extra, nskip = len + nbits;
        if (BD_UNLIKELY(nskip > 32)) {
          BD_LANE_COUNT(5);
          LN_SKIP(len);
          extra = low_bits(LN_PEEK(), nbits);
          nskip = nbits;
        } else {
          extra = low_bits(bits >> len, nbits);
        }
        LN_SKIP(nskip);
        // last distances (sym 0..3) and last / second-to-last distance -3..+3 (sym 4..15), :2017-2049
        const uint32_t cc = sym - 4;
        const uint32_t m = cc < 6 ? cc : cc - 6;
        const int32_t delta = (int32_t)(m >> 1) + 1;
        const int32_t rb0 = sym == 1 ? d1 : (sym == 2 ? d2 : (sym == 3 ? d3 : d0));
        const int32_t rb4 = cc < 6 ? d0 : d1;
        int32_t sd = sym < 4 ? rb0 : ((m & 1) ? rb4 + delta : rb4 - delta);
        if (sym >= 4 && !(m & 1) && sd <= 0) sd = 0x7fffffff;
        dist = longcode ? (int32_t)(base + (extra << npostfix)) : sd;
        push = sym != 0 ? 1u : 0u;
      }
    }
    // ---- look ahead for the symbol phase A decodes next: the bit position of every lane is final for this round.
    //      Lanes that just read their distance will be at a command boundary; lanes inside a literal run
    //      (one tree for all contexts) continue with the literal tree ----
    if (run && !pa_valid && ev == kStCommands) {
      const bool next_cmd = go || ph == kPhCmd;
      const bool next_lit = ph == kPhLit;
      // (late retire: the literal context of a lane whose copy chunk is still to be appended is not known yet)
      if ((next_cmd ? bl_c != 0 : (next_lit && bl_l != 0)) && !(BD_LANE_LATE_RETIRE && !next_cmd && !trivial && pend_n != 0)) {
        uint32_t tv = next_cmd ? cmd_tv : lit_tv;
        if (!next_cmd && !trivial) {  // tree by the context of the last two bytes (:2500-2507); phase P is over
          if (!ctx_fresh) { last_two(hist, bias, posb, acc, p1, p2); ctx_fresh = true; if (kDict && posb - bias < 2) { p1 = 0; p2 = 0; } }
          const uint32_t cx = vlds8(ctx_lut + p1) | vlds8(ctx_lut + 256 + p2);
          tv = root_lit + (vlds8(ctx_map + cx) << r_lit);
        }
#if BD_LANE_NEXT_LA_REG
        LN_LOOKAHEAD_REG(tv, next_cmd ? r_cmd : r_lit, pa_valid, pa_e, pa_idx, pa_need);
#else
        LN_LOOKAHEAD(tv, next_cmd ? r_cmd : r_lit, 32u, pa_valid, pa_e, pa_sel);
#endif
      }
    }
#if BD_LANE_NEXT_LA_REG
    ld16_if(pa_need, gtab + pa_idx, pa_ld);
#endif
    LN_INPUT_BLOCK(run);
    warp_sync();
    cp_async_commit();  // group: next-A look-ahead and the round's input block
#if BD_LANE_LATE_RETIRE
    cp_async_wait_all_but_latest();  // the chunk (and everything older); the group just committed stays in flight
    LN_PHASE_P();
    warp_sync();
#endif

    // ---- phase C2: the copy or static dictionary word (:2583-2689); its source bytes are requested below ----
    if (go && ev == kStCommands) {
      BD_LANE_COUNT(8);
      const uint32_t pos = posb - bias;
      // src/decode.rs:2583-2589 (with a dictionary: pos + its reachable size until that passes max_backward)
      const uint32_t max_distance = kDict ? ((int32_t)pos < cdict_limit ? pos + cdict_size : max_backward) : (pos < max_backward ? pos : max_backward);
      crem = 0;
      if (BD_UNLIKELY((uint32_t)dist > max_distance)) {
        // static dictionary: the transformed word is an entry of the expanded table
        BD_LANE_COUNT(2);
        BD_LANE_DIST_STATS(0u, copy_len, true);
        if (dist <= 0 || dist > 0x7FFFFFFC || copy_len < 4 || copy_len > 24 || k > k_max + 2) {
          ev = kStBail;
        } else {
          const uint32_t wi = vlds32(word_info + copy_len * 4u);
          const uint32_t shift = wi & 15u;
          const uint32_t word_id = (uint32_t)dist - max_distance - 1u;
          const uint32_t t = word_id >> shift;
          if (t >= BROTLI_NUM_TRANSFORMS) {
            ev = kStBail;
          } else {
            const uint32_t ti = vlds32(transform_info + t * 4u);
            const uint32_t cut = ti >> 8;  // == copy_len - transformed_word_length(copy_len, type), saturated
            const uint32_t n = (ti & 15u) + ((ti >> 4) & 15u) + (copy_len > cut ? copy_len - cut : 0u);
            if (n > capb - posb || n == 0) {  // (an empty word makes no progress: left to the exact decoder)
              ev = kStBail;
            } else {
              csrc = xdict + ((size_t)(wi >> 4) << 2) + (size_t)(low_bits(word_id, shift) * BROTLI_NUM_TRANSFORMS + t) * xdict_stride(copy_len);
              crem = n;
              mlen -= (int32_t)n;
            }
          }
        }
      } else {
        if (push) { d3 = d2; d2 = d1; d1 = d0; d0 = dist; }
        if (BD_UNLIKELY(copy_len > capb - posb)) {
          ev = kStBail;
        } else {
          mlen -= (int32_t)copy_len;
          crem = copy_len;
          uint32_t ud = (uint32_t)dist;
          BD_LANE_DIST_STATS(ud, copy_len, false);
#if BD_LANE_PROBE_NEAR == 1
          if (pos >= 64) ud = 20u + (ud & 15u);   // (may share a 32-byte sector with the word being written: a partial sector)
#elif BD_LANE_PROBE_NEAR == 2
          if (pos >= 256) ud = 64u + (ud & 63u);  // whole sectors written two to sixteen rounds ago
#elif BD_LANE_PROBE_NEAR == 3
          if (pos >= 1024) ud = 256u + (ud & 255u);
#elif BD_LANE_PROBE_NEAR == 4
          if (pos >= 4096) ud = 1024u + (ud & 1023u);
#elif BD_LANE_PROBE_NEAR == 5
          if (pos >= 16384) ud = 4096u + (ud & 4095u);
#elif BD_LANE_PROBE_NEAR == 6
          if (ud < 4096 && pos >= 8192) ud += 4096u;  // no copy reads anything written in the last 4 KiB
#endif
          if (kDict && ud > pos) {
            // the source starts in the custom dictionary, which logically precedes the output: a plain copy from the
            // dictionary's tail when it also ends there; a copy that runs on into the output is the exact kernel's
            if (ud - pos < copy_len) { ev = kStBail; crem = 0; }
            csrc = cdict_end - (ud - pos);
          } else {
          if (BD_UNLIKELY(ud < 20)) {
            BD_LANE_COUNT(0);
            // Short period: copy byte-wise until the period can be widened to >= 20 (a copy at distance d
            // equals a copy at distance k*d once k*d bytes are out); chunks do the rest.
            const uint32_t wide = ud * ((19u + ud) / ud);
            const uint32_t m = crem < wide ? crem : wide;
            // up to four bytes per step, sources from the partial word and the history ring (never a load from
            // global memory); while the period is shorter than four bytes it is widened as the bytes come out
            uint32_t dd = ud;
            for (uint32_t left = m; left != 0;) {
              BD_LANE_COUNT(1);
              uint32_t n4 = dd < 4 ? dd : 4u;
              if (n4 > left) n4 = left;
              uint32_t v = recent4(hist, posb, acc, posb - dd);
              if (n4 < 4) v = low_bits(v, 8 * n4);
              append(out_al, bias, hist, posb, acc, v, n4);
              left -= n4;
              if (dd < 4) dd += ud;
            }
            crem -= m;
            ud = wide;
          }
          // Everything below the current output word is in memory, and a distance >= 20 keeps the sixteen
          // source bytes of every chunk below that word at the time the chunk is loaded.
          csrc = out_al + (posb - ud);
          }
        }
      }
      if (ev == kStCommands) {
        if (mlen <= 0) ev = kStHeader;  // end of the metablock; the copy is drained after the loop
        else ph = kPhCopy;
      }
    }
#if BD_LANE_HEAD_PER_ROUND
    if (go) LN_HEAD_CHECK(posb0);  // the short-distance path (distance <= position < 4: at most 21 bytes)
#endif
    warp_sync();
    LN_ISSUE_CHUNK(run && crem != 0 && ev != kStBail);
    cp_async_commit();  // group: the copy chunk
    // last chunk in flight (or nothing to copy: zero-length dictionary output): the next command can be decoded
    if (run && ph == kPhCopy && crem == 0 && ev == kStCommands) ph = kPhCmd;
    if (run && ev != kStCommands) {  // this lane's registers stay as they are until the whole warp is through
      run = false;
      st = ev;
    }
  }
  if (ran && st == kStHeader) {
    // the metablock's last copy may still be in flight
    while (pend_n != 0) {
      cp_async_wait_all();
#if BD_LANE_HEAD_PER_ROUND
      const uint32_t posb1 = posb;
#endif
      LN_RETIRE_CHUNK();
#if BD_LANE_HEAD_PER_ROUND
      LN_HEAD_CHECK(posb1);
#endif
      LN_ISSUE_CHUNK(crem != 0);
      cp_async_commit();
    }
  }
  cp_async_wait_all();  // input blocks requested in the last round: the per-metablock code reads the ring right away
  if (ran) LN_SAVE();
#undef LN_PEEK
#undef LN_APPEND_A
#undef LN_STAGE
#undef LN_PHASE_P
#undef LN_LOOKAHEAD_REG
#undef LN_TAKE_REG
#undef LN_FLUSH_A
#undef LN_HEAD_CHECK
#undef LN_CP16_IF_KEEP
#undef LN_CP16_IF_STREAM
#undef LN_CP16_IF_SRC
#undef LN_SKIP
#undef LN_KLIM
#undef LN_INPUT_BLOCK
#undef LN_BLOCK_SWITCH_RING
#undef LN_SAVE
#undef LN_TREES
#undef LN_BLOCK_SWITCH
#undef LN_DECODE
#undef LN_LOOKAHEAD
#undef LN_TAKE
#undef LN_ISSUE_CHUNK
#undef LN_RETIRE_CHUNK
}

// Stream header: bit window and output cursor set-up, DecodeWindowBits (src/decode.rs:152-187).
BD_DEV uint32_t stream_begin(const LaneCtx& c, Lane& L, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap) {
  if (in_size == 0 || in_size >= ((uint64_t)1 << 31)) return kStBail;
  const uintptr_t ia = (uintptr_t)in;
  L.lead = (uint32_t)(ia & 15u);
  L.gin = in - L.lead;
  L.ring = c.ring;
  L.ring_stride = c.ring_stride;
  L.end_bit = 8 * ((uint64_t)L.lead + in_size);
  L.k_max = (uint32_t)((L.lead + in_size - 1) >> 2);
  L.last_blk = L.k_max >> 2;
  L.k = L.lead >> 2;
  L.bp = 8 * (L.lead & 3u);
  // blocks 0 and 1 now; from here on ring_next keeps one block of look-ahead in flight
  cp_async_wait_all();  // nothing of the previous stream may still be landing in the ring
  cp_async16(L.ring, L.gin);
  cp_async16(L.ring + L.ring_stride, L.gin + 16 * (size_t)(1 < L.last_blk ? 1 : L.last_blk));
  cp_async_commit();
  cp_async_wait_all();
  L.lo = vlds32(L.ring + ((L.k >> 2) & 1u) * L.ring_stride + (L.k & 3u) * 4u);
  L.hi = vlds32(L.ring + (((L.k + 1) >> 2) & 1u) * L.ring_stride + ((L.k + 1) & 3u) * 4u);
  L.nx = vlds32(L.ring + (((L.k + 2) >> 2) & 1u) * L.ring_stride + ((L.k + 2) & 3u) * 4u);
  if (L.k + 2 >= 4) {  // the window already reaches into block 1: block 2 must be on its way (block 0 is all in registers)
    cp_async16(L.ring, L.gin + 16 * (size_t)(2 < L.last_blk ? 2 : L.last_blk));
    cp_async_commit();
    cp_async_wait_all();
  }
  const uintptr_t oa = (uintptr_t)out;
  L.bias = (uint32_t)(oa & 3u);
  L.out_al = out - L.bias;
  const uint64_t cap = out_cap > 0xF0000000ull ? 0xF0000000ull : out_cap;
  L.capb = (uint32_t)cap + L.bias;
  L.posb = L.bias;
  L.acc = 0;
  L.d0 = 4; L.d1 = 11; L.d2 = 15; L.d3 = 16;
  L.mlen = 0; L.is_last = 0;
  if (L.read(1) == 0) {
    L.wbits = 16;
  } else {
    uint32_t n = L.read(3);
    if (n != 0) {
      L.wbits = 17 + n;
    } else {
      n = L.read(3);
      if (n == 1) return kStBail;  // large-window marker (or invalid)
      L.wbits = n != 0 ? 8 + n : 17;
    }
  }
  L.max_backward = (1u << L.wbits) - 16;
  L.cdict_size = c.cdict_len > L.max_backward ? L.max_backward : (uint32_t)c.cdict_len;
  L.cdict_end = c.cdict + c.cdict_len;
  L.cdict_limit = c.cdict_len > L.max_backward ? 0 : (int32_t)(L.max_backward - (uint32_t)c.cdict_len);
  return kStHeader;
}

// After the last metablock: final padding must be zero (src/decode.rs:3365-3373); flush the write combiner.
BD_DEV uint32_t stream_finish(Lane& L, uint64_t* decoded, uint64_t* used) {
  const uint32_t pad = (8u - (L.bp & 7u)) & 7u;
  if (pad && L.read(pad) != 0) return kStBail;
  if (L.overrun()) return kStBail;
  flush_partial(L.out_al, L.bias, L.posb, L.acc);
  *decoded = L.pos();
  const uint64_t bitpos = (uint64_t)L.k * 32 + L.bp;
  *used = ((bitpos + 7) >> 3) - L.lead;
  return kStDone;
}

// One stream per lane, the whole warp together (lanes without a stream pass active == false).
// Returns kStDone (decoded; sizes written) or kStBail (hand the stream to the exact kernel).
template <uint32_t kStride, bool kDict>
BD_DEV uint32_t decode_streams(const LaneCtx& c, bool active, const uint8_t* in, uint64_t in_size, uint8_t* out, uint64_t out_cap,
                               uint64_t* decoded, uint64_t* used) {
  Lane L;
  BlockTrees bt;
  uint32_t st = kStIdle;
  if (active) st = stream_begin(c, L, in, in_size, out, out_cap);
  for (;;) {
    while (st == kStHeader) {  // (metablocks without commands -- raw bytes, metadata -- are done on the spot)
      const int r = metablock_begin(c, L, bt);
      if (r == kLaneNext) { st = L.is_last ? kStFinish : (L.overrun() ? kStBail : kStHeader); continue; }
      st = r == kLaneOk ? kStCommands : (r == kLaneDone ? kStFinish : kStBail);
    }
    warp_sync();
    if (!warp_any(st == kStCommands)) break;
    // (two instances of the loop: see LN_LOOKAHEAD)
    const bool in_arena = st == kStCommands && (L.root[0] >= c.E || L.root[1] >= c.E || L.root[2] >= c.E);
    // (a few such lanes do not pay for the longer look-ahead code of the other lanes: measured on the headline batch,
    // where 3 % of the streams have their distance trees in the arena)
    if (BD_LANE_ASYNC_ARENA_ROOTS && warp_count(in_arena) >= BD_LANE_ARENA_LANES) run_commands<kStride, kDict, true>(c, L, bt, st == kStCommands, st);
    else run_commands<kStride, kDict, false>(c, L, bt, st == kStCommands, st);
    // METABLOCK_DONE, src/decode.rs:3345-3381: BLOCK_LENGTH_2 (:3356-3359) / truncated input
    if (st == kStHeader && (L.mlen < 0 || L.overrun())) st = kStBail;
    if (st == kStHeader && L.is_last) st = kStFinish;
  }
  if (st == kStFinish) st = stream_finish(L, decoded, used);
  return st;
}

}  // namespace lane
}  // namespace BD_NS
