// brotli_b200_session.h -- BrotliDecompressStream semantics (src/decode.rs:2779-2916, src/ffi/mod.rs:389-463) over
// resumable launches of the warp-per-stream kernel.
//
// A Session is what a BrotliDecoderState owns: a sliding window of the stream's input and output in device memory, a
// table arena, the decoder's checkpoint (ResumeState, brotli_decode_core.cuh) and the decoded bytes the caller has not
// taken yet.  stream_calls() does for n states what one BrotliDecompressStream call does for one, with ONE decode
// launch for all of them:
//
//   reference                                              here
//   ---------------------------------------------------   ------------------------------------------------------------
//   decodes until the input runs out: NeedsMoreInput,      the kernel continues from the checkpoint, reports the position
//   all input consumed, ring buffer flushed (forced)       reached; new bytes come down, min(available_out, ...) delivered
//   ring buffer full (pos reaches its size) and the        the kernel stops at the emulated flush point when it lies beyond
//   caller's buffer cannot take it: NeedsMoreOutput,       the budget delivered + available_out; input consumed up to the
//   input consumed up to the bit position                  command in flight; later calls hand out what is pending first
//   end of stream: Success once everything is written      same; trailing input stays with the caller
//
// Memory per session is bounded by the window: input behind the checkpoint and output more than one window behind
// it are dropped (the buffers slide; positions in the ResumeState are rebased).
//
// The logic is a template over the device backend so that tests/hostsim can run the very same code against the
// host build of the kernel (no GPU in the dev container); the product instantiates it with CUDA (brotli_b200_host.cpp).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

#include "brotli_b200_session_types.h"

namespace brotli_b200 {

// BrotliDecoderResult / BrotliResult values (c/brotli/decode.h:40-49)
enum : int { kResError = 0, kResSuccess = 1, kResNeedsMoreInput = 2, kResNeedsMoreOutput = 3 };

struct StreamCall {  // the arguments of one BrotliDecoderDecompressStream call and its result
  size_t* available_in; const uint8_t** next_in; size_t* available_out; uint8_t** next_out; size_t* total_out;
  int result;
};

struct Session {
  // ---- device ----
  uint8_t* d_in[2] = {nullptr, nullptr}; size_t d_in_cap[2] = {0, 0}; int in_sel = 0;
  uint64_t in_base = 0; size_t in_len = 0;   // stream offsets [in_base, in_base + in_len) sit at d_in[in_sel]
  uint8_t* d_out[2] = {nullptr, nullptr}; size_t d_out_cap[2] = {0, 0}; int out_sel = 0;
  uint64_t out_base = 0;                     // stream position of d_out[out_sel][0]
  uint8_t* d_arena = nullptr;
  uint8_t* d_dict = nullptr;                 // 32 bytes of slack on either side
  std::vector<uint8_t> dict;
  ResumeState rs;                            // host mirror; the checkpoint lives here between launches
  // ---- host ----
  std::vector<uint8_t> pending; size_t pending_off = 0;  // decoded bytes the caller has not taken yet
  uint64_t delivered = 0;   // output bytes handed to the caller
  uint64_t fetched = 0;     // output positions below it have come down from the device
  uint64_t consumed = 0;    // input bytes taken from the caller
  enum Stop { kFresh, kNeedIn, kAtFlush, kDone, kFailed } stop = kFresh;
  uint64_t flush_at = 0;    // kAtFlush: the flush point the decoder waits at
  uint32_t grow = 0;        // output capacity granted beyond the budget, doubled whenever a launch ran into it
  int code = 0;             // BrotliDecoderErrorCode of the last call
  bool large_window = false, used = false;
  Session() { memset(&rs, 0, sizeof(rs)); }
  size_t pending_bytes() const { return pending.size() - pending_off; }
};

// Dev: uint8_t* alloc(size_t) / void release(uint8_t*) / int upload(dst, src, n) /
//      uint8_t* host_up(bytes)     staging memory on the host for the batch's fresh input (pinned on a GPU)
//      int run(ResumeState* sessions, n, const uint8_t* blob, blob_bytes, SessionCopy* scatter, n_scatter)   (blob = host_up(); scatter[i].src = offset into blob)
//      const uint8_t* gather(SessionCopy* pieces, n, bytes)   new output -> one host blob (pieces[i].dst = offset into it), nullptr on error
//      int move(SessionCopy* pieces, n)                                                                       device -> device
//      size_t arena_bytes()
template <class Dev>
struct SessionRunner {
  Dev& dev;
  std::vector<SessionCopy> pre_moves;  // buffer growth of this batch: executed as one device-to-device pass before the decode
  explicit SessionRunner(Dev& d) : dev(d) {}

  // hand pending bytes to the caller (positions below `limit` only)
  static void deliver(Session& s, StreamCall& c, uint64_t limit = ~(uint64_t)0) {
    size_t n = s.pending_bytes();
    if (n > *c.available_out) n = *c.available_out;
    if (limit < s.delivered + n) n = limit > s.delivered ? (size_t)(limit - s.delivered) : 0;
    if (n) {
      memcpy(*c.next_out, s.pending.data() + s.pending_off, n);
      *c.next_out += n; *c.available_out -= n; s.pending_off += n; s.delivered += n;
    }
    if (s.pending_off == s.pending.size()) { s.pending.clear(); s.pending_off = 0; }
    else if (s.pending_off > (1u << 20) && s.pending_off > s.pending.size() / 2) {
      s.pending.erase(s.pending.begin(), s.pending.begin() + (ptrdiff_t)s.pending_off); s.pending_off = 0;
    }
    if (c.total_out) *c.total_out = (size_t)s.delivered;
  }

  // Host copies of one batch (caller input -> staging blob, staging blob -> caller output): collected while the
  // calls are walked in order, then made in parallel (they are independent and dominate a batch of many sessions).
  struct HostCopy { uint8_t* dst; const uint8_t* src; size_t n; };
  std::vector<HostCopy> jobs;
  void run_jobs() {
    const long nj = (long)jobs.size();
#pragma omp parallel for schedule(dynamic, 16) if (nj >= 64)
    for (long j = 0; j < nj; j++) memcpy(jobs[(size_t)j].dst, jobs[(size_t)j].src, jobs[(size_t)j].n);
    jobs.clear();
  }
  // hand the caller what is pending, then this call's new bytes (src, n) straight from the staging blob: only what
  // the caller's buffer cannot take is kept in `pending`
  void deliver_new(Session& s, StreamCall& c, const uint8_t* src, size_t n, uint64_t limit = ~(uint64_t)0) {
    if (n == 0 || s.pending_bytes() != 0) {
      if (n) s.pending.insert(s.pending.end(), src, src + n);
      deliver(s, c, limit);
      return;
    }
    size_t m = n;
    if (m > *c.available_out) m = *c.available_out;
    if (limit < s.delivered + m) m = limit > s.delivered ? (size_t)(limit - s.delivered) : 0;
    if (m) {
      jobs.push_back(HostCopy{*c.next_out, src, m});
      *c.next_out += m; *c.available_out -= m; s.delivered += m;
    }
    if (m < n) s.pending.insert(s.pending.end(), src + m, src + n);
    if (c.total_out) *c.total_out = (size_t)s.delivered;
  }

  bool reserve(uint8_t*& p, size_t& cap, size_t need) {
    if (need <= cap) return true;
    const size_t want = (need + 4095) & ~(size_t)4095;  // (callers add the slack they want: the two halves of a pair must not leapfrog)
    uint8_t* q = dev.alloc(want);
    if (!q) return false;
    if (p) dev.release(p);
    p = q; cap = want;
    return true;
  }
  // grow the active buffer of a ping-pong pair, keeping its first `keep` bytes
  bool grow_keep(uint8_t* (&p)[2], size_t (&cap)[2], int& sel, size_t need, size_t keep) {
    if (need <= cap[sel]) return true;
    const int o = sel ^ 1;
    if (!reserve(p[o], cap[o], need + need / 2)) return false;
    if (keep) pre_moves.push_back(SessionCopy{p[sel], p[o], keep});  // (the old half stays allocated: it is the source)
    sel = o;
    return true;
  }

  int fail(Session& s, StreamCall& c, int code) {
    s.stop = Session::kFailed; s.code = code;
    c.result = kResError;
    return kResError;
  }

  // n concurrent BrotliDecoderDecompressStream calls, one per state.  Returns 0, or a device error (then every call of
  // the batch that needed the device has failed with `device_error_code`).
  int stream_calls(Session** ss, StreamCall* cs, size_t n, int device_error_code) {
    std::vector<uint32_t> todo;       // calls that need the decoder
    std::vector<uint64_t> budget(n, 0);
    jobs.clear();
    for (size_t i = 0; i < n; i++) {
      Session& s = *ss[i]; StreamCall& c = cs[i];
      c.result = -1;
      if (s.stop == Session::kFailed) { c.result = kResError; continue; }  // sticky, src/decode.rs:2796-2798
      if (*c.available_in >= ((uint64_t)1 << 32)) { fail(s, c, -20); continue; }  // :2799-2801
      budget[i] = s.delivered + *c.available_out;
      // calls the decoder cannot make progress in only write what the ring buffer holds (WriteRingBuffer at the state the
      // last call stopped in; the forced flush of NeedsMoreInput, src/decode.rs:2838-2850)
      if (s.stop == Session::kDone) { deliver(s, c); c.result = s.pending_bytes() ? kResNeedsMoreOutput : kResSuccess; s.code = c.result; continue; }
      if (s.stop == Session::kAtFlush && s.flush_at > budget[i]) { deliver(s, c); c.result = kResNeedsMoreOutput; s.code = 3; continue; }
      if ((s.stop == Session::kNeedIn || s.stop == Session::kFresh) && *c.available_in == 0) { deliver(s, c); c.result = kResNeedsMoreInput; s.code = 2; continue; }
      todo.push_back((uint32_t)i);
    }
    if (todo.empty()) return 0;
    // ---- fresh input: one blob for the whole batch ----
    size_t blob_need = 0;
    for (uint32_t i : todo) blob_need += *cs[i].available_in;
    uint8_t* const blob = dev.host_up(blob_need + 16);
    size_t blob_size = 0;
    if (!blob) { for (uint32_t i : todo) fail(*ss[i], cs[i], device_error_code); return device_error_code; }
    std::vector<SessionCopy> scatter;
    std::vector<size_t> fresh(n, 0);
    for (uint32_t i : todo) {
      Session& s = *ss[i]; StreamCall& c = cs[i];
      const size_t k = *c.available_in;
      fresh[i] = k;
      if (k) s.used = true;
      if (!grow_keep(s.d_in, s.d_in_cap, s.in_sel, s.in_len + k + 16, s.in_len)) { fail(s, c, -26); continue; }
      if (!s.d_arena) { s.d_arena = dev.alloc(dev.arena_bytes()); if (!s.d_arena) { fail(s, c, -22); continue; } }
      if (!s.dict.empty() && !s.d_dict) {
        s.d_dict = dev.alloc(s.dict.size() + 64);
        if (!s.d_dict || dev.upload(s.d_dict + 32, s.dict.data(), s.dict.size()) != 0) { fail(s, c, -26); continue; }
      }
      if (k) {
        scatter.push_back(SessionCopy{(const uint8_t*)(uintptr_t)blob_size, s.d_in[s.in_sel] + s.in_len, k});
        jobs.push_back(HostCopy{blob + blob_size, *c.next_in, k});
        blob_size += k;
        s.in_len += k;
      }
    }
    run_jobs();
    // ---- decode; a launch that ran into the end of its output window is repeated with a larger one ----
    std::vector<uint32_t> round;
    for (uint32_t i : todo) if (cs[i].result == -1) round.push_back(i);
    std::vector<ResumeState> arr;
    bool first = true;
    while (!round.empty()) {
      arr.resize(round.size());
      std::vector<uint32_t> live;
      for (uint32_t i : round) {
        Session& s = *ss[i]; StreamCall& c = cs[i];
        // window: everything up to the budget plus `grow` (the decoder runs ahead of the caller by up to one ring buffer)
        if (s.grow == 0) s.grow = 1u << 16;
        const uint64_t want_abs = budget[i] + s.grow + 64;
        uint64_t have = s.fetched > s.out_base ? s.fetched - s.out_base : 0;  // bytes of the window a later launch may read
        if (s.rs.kind != 0 && s.rs.pos > have) have = s.rs.pos;                  // (everything below the checkpoint)
        if (!grow_keep(s.d_out, s.d_out_cap, s.out_sel, (size_t)(want_abs - s.out_base), (size_t)have)) { fail(s, c, -26); continue; }
        live.push_back(i);
      }
      arr.resize(live.size());
      for (size_t k = 0; k < live.size(); k++) {
        Session& s = *ss[live[k]];
        ResumeState& r = s.rs;
        r.in = s.d_in[s.in_sel]; r.in_size = s.in_len;
        r.out = s.d_out[s.out_sel];
        uint64_t cap = budget[live[k]] + s.grow + 64 - s.out_base;
        if (cap > s.d_out_cap[s.out_sel]) cap = s.d_out_cap[s.out_sel];
        r.out_cap = cap;
        r.budget = budget[live[k]] - s.out_base;
        r.arena = s.d_arena;
        r.dict = s.d_dict ? s.d_dict + 32 : nullptr; r.dict_size = s.dict.size();
        r.allow_large_window = s.large_window ? 1u : 0u;
        arr[k] = r;
      }
      if (live.empty()) { pre_moves.clear(); break; }
      if (!pre_moves.empty()) {
        const int mrc = dev.move(pre_moves.data(), (uint32_t)pre_moves.size());
        pre_moves.clear();
        if (mrc != 0) { for (uint32_t i : live) fail(*ss[i], cs[i], device_error_code); return mrc; }
      }
      const int rc = dev.run(arr.data(), (uint32_t)arr.size(), first ? blob : nullptr, first ? blob_size : 0,
                             first ? scatter.data() : nullptr, first ? (uint32_t)scatter.size() : 0u);
      first = false;
      if (rc != 0) { for (uint32_t i : live) fail(*ss[i], cs[i], device_error_code); return rc; }
      round.clear();
      for (size_t k = 0; k < live.size(); k++) {
        Session& s = *ss[live[k]];
        s.rs = arr[k];
        // ran into the end of the window before a stop the reference knows: same checkpoint, more room
        if (s.rs.hit_cap && s.rs.code != 1) {
          if (s.grow >= (1u << 31)) { fail(s, cs[live[k]], -26); continue; }
          s.grow *= 4;
          round.push_back(live[k]);
        }
      }
    }
    // ---- new output comes down: one blob for the whole batch ----
    std::vector<SessionCopy> gather;
    std::vector<uint32_t> got;
    std::vector<const uint8_t*> new_ptr(n, nullptr);
    std::vector<size_t> new_n(n, 0);
    size_t out_bytes = 0;
    for (uint32_t i : todo) {
      Session& s = *ss[i];
      if (cs[i].result != -1) continue;
      const ResumeState& r = s.rs;
      uint64_t upto = s.out_base + (r.code < 0 ? r.flushed_now : r.decoded);  // a fatal error leaves only what was flushed
      if (upto > s.fetched) {
        const size_t k = (size_t)(upto - s.fetched);
        gather.push_back(SessionCopy{s.d_out[s.out_sel] + (s.fetched - s.out_base), (uint8_t*)(uintptr_t)out_bytes, k});
        got.push_back(i);
        out_bytes += k;
      }
    }
    if (!gather.empty()) {
      const uint8_t* down = dev.gather(gather.data(), (uint32_t)gather.size(), out_bytes);
      if (!down) { for (uint32_t i : todo) if (cs[i].result == -1) fail(*ss[i], cs[i], device_error_code); return device_error_code; }
      for (size_t k = 0; k < got.size(); k++) {
        Session& s = *ss[got[k]];
        new_ptr[got[k]] = down + (uintptr_t)gather[k].dst;
        new_n[got[k]] = gather[k].n;
        s.fetched += gather[k].n;
      }
    }
    // ---- results, input accounting, sliding windows ----
    std::vector<SessionCopy> moves;
    for (uint32_t i : todo) {
      Session& s = *ss[i]; StreamCall& c = cs[i];
      if (c.result != -1) continue;
      const ResumeState& r = s.rs;
      s.code = r.code;
      const uint64_t before = s.consumed;            // == in_base + in_len - fresh[i]
      uint64_t used_abs = s.in_base + r.used;
      if (r.code == 2) used_abs = s.in_base + s.in_len;  // NeedsMoreInput swallows the whole input, src/decode.rs:2887-2899
      // An error out of the forced flush of NeedsMoreInput leaves the caller's counters where they were (the reference
      // breaks out before it writes them back, :2843-2846; bytes it had taken into its 8-byte carry-over -- at most 7 -- are
      // the one thing this does not reproduce).
      if (r.code < 0 && r.forced_flush_error) used_abs = before;
      if (used_abs < before) used_abs = before;
      if (used_abs > s.in_base + s.in_len) used_abs = s.in_base + s.in_len;
      const size_t adv = (size_t)(used_abs - before);
      *c.next_in += adv; *c.available_in -= adv;
      s.consumed = used_abs;
      s.in_len = (size_t)(used_abs - s.in_base);     // what the caller keeps will come again
      if (r.code == 2) { s.stop = Session::kNeedIn; deliver_new(s, c, new_ptr[i], new_n[i]); c.result = kResNeedsMoreInput; }
      else if (r.code == 3) { s.stop = Session::kAtFlush; s.flush_at = s.out_base + r.decoded; deliver_new(s, c, new_ptr[i], new_n[i]); c.result = kResNeedsMoreOutput; }
      else if (r.code == 1) { s.stop = Session::kDone; deliver_new(s, c, new_ptr[i], new_n[i]); c.result = s.pending_bytes() ? kResNeedsMoreOutput : kResSuccess; }
      else { deliver_new(s, c, new_ptr[i], new_n[i], s.out_base + r.flushed_now); fail(s, c, r.code); }  // only what passed a flush point: nothing is flushed at the error
      if (s.stop == Session::kFailed || s.stop == Session::kDone) continue;
      // slide the input window: bytes behind the checkpoint are never read again
      const uint64_t cut = s.rs.bitpos >> 3;
      if (cut >= 4096 && cut >= s.in_len / 2 && cut <= s.in_len) {
        const size_t rem = s.in_len - (size_t)cut;
        const int o = s.in_sel ^ 1;
        if (reserve(s.d_in[o], s.d_in_cap[o], rem + 4096)) {
          if (rem) moves.push_back(SessionCopy{s.d_in[s.in_sel] + cut, s.d_in[o], rem});
          s.in_sel = o; s.in_base += cut; s.in_len = rem; s.rs.bitpos -= cut * 8;
        }
      }
      // slide the output window: copies reach back one window from the checkpoint at most, and everything below
      // `fetched` is already on the host
      if (s.rs.kind != 0) {
        const uint64_t window = (uint64_t)1 << s.rs.wbits;
        const uint64_t ck = s.rs.pos;  // relative
        if (ck > 2 * window + (1u << 16)) {
          const uint64_t drop = ck - window;  // keep [ck - window, ...)
          const uint64_t keep = (s.fetched - s.out_base) - drop;
          const int o = s.out_sel ^ 1;
          if (reserve(s.d_out[o], s.d_out_cap[o], s.d_out_cap[s.out_sel])) {
            if (keep) moves.push_back(SessionCopy{s.d_out[s.out_sel] + drop, s.d_out[o], (size_t)keep});
            s.out_sel = o; s.out_base += drop;
            s.rs.pos -= (uint32_t)drop;
            s.rs.next_flush = s.rs.next_flush == ~(uint64_t)0 ? s.rs.next_flush : s.rs.next_flush - drop;
            s.rs.flushed = s.rs.flushed > drop ? s.rs.flushed - drop : 0;
          }
        }
      }
    }
    run_jobs();  // (the staging blob is the device backend's: it stays valid until the next batch)
    if (!moves.empty()) {
      const int rc = dev.move(moves.data(), (uint32_t)moves.size());
      if (rc != 0) return rc;
    }
    return 0;
  }

  void destroy(Session& s) {
    for (int k = 0; k < 2; k++) { if (s.d_in[k]) dev.release(s.d_in[k]); if (s.d_out[k]) dev.release(s.d_out[k]); s.d_in[k] = s.d_out[k] = nullptr; }
    if (s.d_arena) dev.release(s.d_arena);
    if (s.d_dict) dev.release(s.d_dict);
    s.d_arena = s.d_dict = nullptr;
  }
};

}  // namespace brotli_b200
