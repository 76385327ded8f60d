// profiles/probes/session_bench.cpp -- throughput of multiplexed streaming sessions through the C ABI, without a Python
// driver in the loop: N open BrotliDecoderStates, each fed `piece` bytes of compressed input per round through
// BrotliB200DecoderDecompressStreamBatch (one decode launch per round), output taken into per-session buffers and
// checksummed against the originals' lengths.  Streams come from a file written by profiles/gpu_sessions.py:
//   u32 U; U x {u32 compressed_len, u32 decompressed_len}; blobs.
//   g++ -O2 -I include session_bench.cpp -L <pkg> -l:libbrotli_b200.so -Wl,-rpath,<pkg> -o session_bench
#include <brotli_b200/decode.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: session_bench STREAMS.bin N_SESSIONS PIECE_BYTES [OUT_CAP]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  const size_t n = (size_t)atol(argv[2]), piece = (size_t)atol(argv[3]), out_cap = argc > 4 ? (size_t)atol(argv[4]) : (1u << 17);
  uint32_t U = 0;
  if (fread(&U, 4, 1, f) != 1) return 2;
  std::vector<uint32_t> hdr(2 * U);
  if (fread(hdr.data(), 4, 2 * U, f) != 2 * U) return 2;
  std::vector<std::vector<uint8_t>> comp(U);
  for (uint32_t i = 0; i < U; i++) { comp[i].resize(hdr[2 * i]); if (fread(comp[i].data(), 1, comp[i].size(), f) != comp[i].size()) return 2; }
  fclose(f);
  std::vector<BrotliDecoderState*> st(n);
  for (auto& s : st) s = BrotliDecoderCreateInstance(nullptr, nullptr, nullptr);
  std::vector<size_t> pos(n, 0), avail_in(n), avail_out(n), total(n, 0), got(n, 0);
  std::vector<const uint8_t*> next_in(n);
  std::vector<uint8_t*> next_out(n);
  std::vector<BrotliDecoderResult> res(n, BROTLI_DECODER_RESULT_NEEDS_MORE_INPUT);
  std::vector<std::vector<uint8_t>> out(n, std::vector<uint8_t>(out_cap));
  // warm-up call: device context, kernels, allocator slabs are not what is measured
  { size_t ai = 0, ao = 0; const uint8_t* ni = nullptr; uint8_t* no = nullptr; BrotliDecoderState* w = BrotliDecoderCreateInstance(nullptr, nullptr, nullptr);
    ai = comp[0].size() < 512 ? comp[0].size() : 512; ni = comp[0].data(); std::vector<uint8_t> o(65536); ao = o.size(); no = o.data();
    BrotliDecoderDecompressStream(w, &ai, &ni, &ao, &no, nullptr); BrotliDecoderDestroyInstance(w); }
  size_t rounds = 0, live = n;
  std::vector<BrotliDecoderState*> bs; std::vector<size_t> idx;
  std::vector<size_t> b_ai, b_ao, b_to; std::vector<const uint8_t*> b_ni; std::vector<uint8_t*> b_no; std::vector<BrotliDecoderResult> b_res;
  const auto t0 = std::chrono::steady_clock::now();
  double first_round_s = 0;
  while (live) {
    bs.clear(); idx.clear(); b_ai.clear(); b_ao.clear(); b_ni.clear(); b_no.clear(); b_to.clear();
    for (size_t i = 0; i < n; i++) {
      if (res[i] != BROTLI_DECODER_RESULT_NEEDS_MORE_INPUT && res[i] != BROTLI_DECODER_RESULT_NEEDS_MORE_OUTPUT) continue;
      const std::vector<uint8_t>& c = comp[i % U];
      if (res[i] == BROTLI_DECODER_RESULT_NEEDS_MORE_INPUT) {
        size_t k = c.size() - pos[i]; if (k > piece) k = piece;
        avail_in[i] = k; next_in[i] = c.data() + pos[i]; pos[i] += k;
      }
      idx.push_back(i); bs.push_back(st[i]); b_ai.push_back(avail_in[i]); b_ni.push_back(next_in[i]);
      b_ao.push_back(out_cap); b_no.push_back(out[i].data()); b_to.push_back(0);
    }
    b_res.assign(idx.size(), BROTLI_DECODER_RESULT_ERROR);
    if (BrotliB200DecoderDecompressStreamBatch(idx.size(), bs.data(), b_ai.data(), b_ni.data(), b_ao.data(), b_no.data(), b_to.data(), b_res.data()) != 0) {
      fprintf(stderr, "batch call failed: %s\n", BrotliB200LastError()); return 1;
    }
    for (size_t k = 0; k < idx.size(); k++) {
      const size_t i = idx[k];
      res[i] = b_res[k]; avail_in[i] = b_ai[k]; next_in[i] = b_ni[k]; got[i] += out_cap - b_ao[k];
      if (res[i] == BROTLI_DECODER_RESULT_SUCCESS || res[i] == BROTLI_DECODER_RESULT_ERROR) live--;
    }
    if (rounds == 0) first_round_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    rounds++;
  }
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double lane_ms = 0, exact_ms = 0; uint32_t launches = 0, bailed = 0;
  BrotliB200KernelTimes(&lane_ms, &exact_ms, &launches, &bailed, 0);
  size_t total_out = 0; bool ok = true;
  for (size_t i = 0; i < n; i++) { total_out += got[i]; ok = ok && res[i] == BROTLI_DECODER_RESULT_SUCCESS && got[i] == hdr[2 * (i % U) + 1]; }
  printf("{\"experiment\": \"multiplexed sessions (C++ driver)\", \"sessions\": %zu, \"piece_bytes\": %zu, \"rounds\": %zu, \"decompressed_GB\": %.3f, "
         "\"wall_s\": %.4f, \"GBps\": %.3f, \"first_round_s\": %.4f, \"GBps_after_first_round\": %.3f, \"decode_kernel_ms_per_launch\": %.3f, \"launches_timed\": %u, \"all_success_and_lengths\": %s}\n",
         n, piece, rounds, total_out / 1e9, dt, total_out / dt / 1e9, first_round_s, total_out / (dt - first_round_s) / 1e9 * (double)(rounds - 1) / rounds, launches ? exact_ms / launches : 0.0, launches, ok ? "true" : "false");
  for (auto s : st) BrotliDecoderDestroyInstance(s);
  return ok ? 0 : 1;
}
