// profiles/probes/gather_probe.cu -- what the memory system gives the lane kernel's access pattern, with no decode work.
// Every lane owns a private 64 KiB window (the output region of its stream; W x 148 x 32 windows = 4..7 GB, far beyond
// L2) and repeats: one asynchronous 16-byte gather (cp.async.cg -> LDGSTS) from a pseudo-random earlier position of its
// own window -- a backreference source -- waited for one iteration later, plus a 4-byte streaming store at its write
// cursor (or the same bytes as 16- / 32-byte stores: "stores" 2 / 3).  G gathers per lane and iteration stay in flight (the kernel has 1..3: copy source, table entry, input block).
// Reported: gathers/s, the DRAM bytes they cost at 64 bytes per miss, and ns per iteration: the ceiling of
// "one 16-byte gather per lane and round" whatever the instruction stream does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu && ./gather_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); exit(1); } } while (0)

__device__ __forceinline__ uint32_t sref(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int G, int STORES>
__global__ void gather_kernel(uint8_t* win, uint32_t window, uint32_t iters, uint32_t* out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint64_t glane = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t* base = win + glane * window;
  uint8_t* slot = smem + (size_t)threadIdx.x * 16 * G * 2;
  uint32_t x = (uint32_t)glane * 2654435761u + 12345u, acc = 0, pos = 64;
  for (uint32_t it = 0; it < iters; it++) {
    uint8_t* dst = slot + (it & 1) * 16 * G;
#pragma unroll
    for (int g = 0; g < G; g++) {
      x = x * 1664525u + 1013904223u;
      const uint32_t off = ((x >> 8) % (pos > 64 ? pos : 64u)) & ~15u;  // an earlier position of this lane's window
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sref(dst + 16 * g)), "l"(base + off) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const uint32_t* w = (const uint32_t*)(slot + ((it + 1) & 1) * 16 * G);
#pragma unroll
    for (int g = 0; g < G; g++) acc += w[4 * g];
    // STORES: 1 = one 4-byte store per iteration (what the kernel does); 2 = one 16-byte store every 4th iteration;
    // 3 = one 32-byte sector (two 16-byte stores) every 8th iteration -- the same bytes, fewer and wider requests
    if (STORES == 1) { asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(base + pos), "r"(acc) : "memory"); }
    if (STORES == 2 && (pos & 15u) == 12) { asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(base + (pos & ~15u)), "r"(acc) : "memory"); }
    if (STORES == 3 && (pos & 31u) == 28) {
      asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(base + (pos & ~31u)), "r"(acc) : "memory");
      asm volatile("st.global.cs.v4.u32 [%0], {%1, %1, %1, %1};" ::"l"(base + (pos & ~31u) + 16), "r"(acc) : "memory");
    }
    pos += 4; if (pos >= window) pos = 64;
  }
  out[glane] = acc;
}

template <int G, int STORES>
void run(int warps, uint8_t* d_win, uint32_t window, uint32_t iters, uint32_t* d_out) {
  const int threads = warps * 32;
  const size_t smem = (size_t)threads * 16 * G * 2;
  CK(cudaFuncSetAttribute(gather_kernel<G, STORES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  gather_kernel<G, STORES><<<148, threads, smem>>>(d_win, window, iters / 8, d_out);  // warm-up
  CK(cudaEventRecord(e0));
  gather_kernel<G, STORES><<<148, threads, smem>>>(d_win, window, iters, d_out);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double gathers = (double)148 * threads * iters * G;
  printf("{\"warps_per_sm\": %d, \"lanes\": %d, \"gathers_in_flight_per_lane\": %d, \"stores\": %d, \"ms\": %.2f, \"Ggathers_per_s\": %.2f, "
         "\"dram_GBps_at_64B_per_gather\": %.0f, \"ns_per_iteration\": %.0f}\n",
         warps, 148 * threads, G, STORES, ms, gathers / ms / 1e6, gathers * 64 / ms / 1e6, ms * 1e6 / iters);
}

int main(int argc, char** argv) {
  const uint32_t window = 65536, iters = 4000;
  const size_t bytes = (size_t)148 * 32 * 32 * window;  // up to 32 warps per SM
  uint8_t* d_win; uint32_t* d_out;
  CK(cudaMalloc(&d_win, bytes)); CK(cudaMemset(d_win, 1, bytes));
  CK(cudaMalloc(&d_out, (size_t)148 * 1024 * 4));
  if (argc >= 2) {  // one line: the lane kernel's pattern (two gathers in flight, 4-byte stores) at `warps` per SM -- bench.py
    const int warps = atoi(argv[1]);
    if (warps < 1 || warps > 32) return 2;
    run<2, 1>(warps, d_win, window, iters, d_out);
    return 0;
  }
  for (int warps : {8, 14, 20, 24, 32}) {
    run<1, 0>(warps, d_win, window, iters, d_out);
    run<1, 1>(warps, d_win, window, iters, d_out);
    run<2, 1>(warps, d_win, window, iters, d_out);
    run<3, 1>(warps, d_win, window, iters, d_out);
    run<2, 2>(warps, d_win, window, iters, d_out);
    run<2, 3>(warps, d_win, window, iters, d_out);
  }
  return 0;
}
