// profiles/probes/tma_probe.cu -- measured basis of the "no TMA in the lane kernel" decision.
// The lane kernel's global->shared traffic is one 16-byte gather per lane and request (input blocks, copy sources, table
// entries): per-lane addresses, per-lane destinations, tens of thousands of independent streams.  This probe moves exactly
// that pattern -- every lane streams its own 24 KB region through a private two-stage shared-memory ring -- once with the
// per-thread asynchronous copy the kernel uses (cp.async.cg 16 -> LDGSTS) and once with the bulk-copy engine
// (cp.async.bulk global->shared + mbarrier complete_tx; one bulk copy per lane and block, block sizes 16..256 bytes).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); exit(1); } } while (0)

constexpr int kWarps = 14, kThreads = kWarps * 32;

__device__ __forceinline__ uint32_t sref(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// B bytes per lane and step through cp.async 16-byte pieces
template <int B>
__global__ void __launch_bounds__(kThreads, 1) ldgsts_kernel(const uint8_t* in, uint64_t per_lane, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint64_t glane = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
  const uint8_t* src = in + glane * per_lane;
  uint8_t* slot = smem + (size_t)threadIdx.x * 2 * B;
  const uint32_t steps = (uint32_t)(per_lane / B);
  uint32_t acc = 0;
  for (int j = 0; j < B; j += 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sref(slot + j)), "l"(src + j) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (uint32_t s = 0; s < steps; s++) {
    if (s + 1 < steps) {
      uint8_t* dst = slot + ((s + 1) & 1) * B;
      for (int j = 0; j < B; j += 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sref(dst + j)), "l"(src + (uint64_t)(s + 1) * B + j) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const uint32_t* w = (const uint32_t*)(slot + (s & 1) * B);
    for (int j = 0; j < B / 16; j++) acc += w[j * 4];
  }
  out[glane] = acc;
}

// B bytes per lane and step through ONE bulk copy per lane; completion through a per-warp mbarrier per stage
template <int B>
__global__ void __launch_bounds__(kThreads, 1) bulk_kernel(const uint8_t* in, uint64_t per_lane, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar[kWarps][2];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t glane = (uint64_t)blockIdx.x * kThreads + threadIdx.x;
  const uint8_t* src = in + glane * per_lane;
  uint8_t* slot = smem + (size_t)threadIdx.x * 2 * B;
  const uint32_t steps = (uint32_t)(per_lane / B);
  if (lane == 0) {
    for (int k = 0; k < 2; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sref(&mbar[warp][k])) : "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t acc = 0;
  auto issue = [&](uint32_t s) {
    const uint32_t mb = sref(&mbar[warp][s & 1]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(32u * B) : "memory");
    __syncwarp();
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sref(slot + (s & 1) * B)), "l"(src + (uint64_t)s * B), "r"((uint32_t)B), "r"(mb) : "memory");
  };
  issue(0);
  for (uint32_t s = 0; s < steps; s++) {
    if (s + 1 < steps) issue(s + 1);
    const uint32_t mb = sref(&mbar[warp][s & 1]), parity = (s >> 1) & 1;
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    }
    const uint32_t* w = (const uint32_t*)(slot + (s & 1) * B);
    for (int j = 0; j < B / 16; j++) acc += w[j * 4];
    __syncwarp();  // every lane has read the stage before it is refilled
  }
  out[glane] = acc;
}

template <typename K>
float time_kernel(K kernel, int ctas, size_t dyn, const uint8_t* in, uint64_t per_lane, uint32_t* out) {
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; w++) kernel<<<ctas, kThreads, dyn>>>(in, per_lane, out);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; r++) kernel<<<ctas, kThreads, dyn>>>(in, per_lane, out);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 5;
}

int main() {
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint64_t per_lane = 24576;  // ~ one compressed 64 KiB text stream
  const uint64_t lanes = (uint64_t)sms * kThreads, total = lanes * per_lane;
  uint8_t* in; uint32_t* out;
  CK(cudaMalloc(&in, total)); CK(cudaMalloc(&out, lanes * 4));
  CK(cudaMemset(in, 1, total));
  uint32_t* h = (uint32_t*)malloc(lanes * 4);
  printf("{\"probe\": \"per-lane streaming global->shared, %llu lanes x %llu B = %.2f GB\", \"results\": [", (unsigned long long)lanes, (unsigned long long)per_lane, total / 1e9);
  bool first = true;
#define RUN(NAME, KERNEL, B) do { \
    float ms = time_kernel(KERNEL<B>, sms, (size_t)kThreads * 2 * B, in, per_lane, out); \
    CK(cudaMemcpy(h, out, lanes * 4, cudaMemcpyDeviceToHost)); \
    bool ok = true; for (uint64_t i = 0; i < lanes; i += 997) ok = ok && h[i] == (uint32_t)(per_lane / 16) * 0x01010101u; \
    printf("%s{\"engine\": \"%s\", \"block_bytes\": %d, \"ms\": %.3f, \"GBps\": %.1f, \"ok\": %s}", first ? "" : ", ", NAME, B, ms, total / ms / 1e6, ok ? "true" : "false"); first = false; } while (0)
  RUN("cp.async 16 B (LDGSTS)", ldgsts_kernel, 16);
  RUN("cp.async 16 B (LDGSTS)", ldgsts_kernel, 64);
  RUN("cp.async 16 B (LDGSTS)", ldgsts_kernel, 128);
  RUN("cp.async.bulk (TMA)", bulk_kernel, 16);
  RUN("cp.async.bulk (TMA)", bulk_kernel, 64);
  RUN("cp.async.bulk (TMA)", bulk_kernel, 128);
  RUN("cp.async.bulk (TMA)", bulk_kernel, 256);
  printf("]}\n");
  return 0;
}
