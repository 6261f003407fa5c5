// Micro-benchmark of the MMA-issuer loop: what one pipeline iteration costs the issuing thread.
#include <cstdio>
#include <cstdlib>
#include "../www2023tiger_b200/csrc/umma.cuh"

// MODE bits: 1 = commit to a barrier each iteration, 2 = try_wait on an already-complete barrier, 4 = tcgen05.fence::after
template <int MODE, int NMMA>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar, bar_done[8], bar_ready;
  __shared__ uint32_t slot;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) f[i] = 1.0f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar_ready, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_done[i], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar_ready);   // phase 0 of bar_ready is complete from now on
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = __shfl_sync(0xffffffffu, slot, 0);
  if (threadIdx.x < 32) {
    const uint32_t base = smem_addr_u32(smem);
    const uint32_t b_lo = umma_desc_lo(base + 64 * 1024, 32);
    const uint32_t idesc = umma_idesc_tf32(128, 32);
    const uint32_t a_t = taddr + 448;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (MODE & 2) mbar_wait(&bar_ready, 0);
      if (MODE & 8) {
        uint32_t ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_addr_u32(&bar_ready)), "r"(0u) : "memory");
        } while (!ok);
      }
      if (MODE & 16) {   // only lane 0 polls, result broadcast
        uint32_t ok = 1;
        if ((threadIdx.x & 31) == 0) {
          do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_addr_u32(&bar_ready)), "r"(0u) : "memory");
          } while (!ok);
        }
        __syncwarp();
      }
      if (MODE & 4) tc_fence_after_sync();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < NMMA; ++j) umma_tf32_ts(taddr + (uint32_t)((j & 1) * 32), a_t + (j & 3) * 8, b_lo + (j & 3) * 64, idesc, 1u);
        if (MODE & 1) umma_commit(&bar_done[r & 7]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(taddr, 512);
}

template <int MODE, int NMMA>
void run(long long* d, const char* what) {
  cudaFuncSetAttribute(bench<MODE, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 256;
  bench<MODE, NMMA><<<1, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h;
  cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%-46s %d MMAs (N=32, TS): %7.1f cycles/iteration\n", what, NMMA, (double)h / reps);
}

int main() {
  long long* d;
  cudaMalloc(&d, sizeof(long long));
  run<0, 4>(d, "MMAs only");
  run<1, 4>(d, "MMAs + commit");
  run<3, 4>(d, "wait(complete) + MMAs + commit");
  run<7, 4>(d, "wait + fence::after + MMAs + commit");
  run<7, 2>(d, "wait + fence::after + MMAs + commit");
  run<7, 8>(d, "wait + fence::after + MMAs + commit");
  run<7, 16>(d, "wait + fence::after + MMAs + commit");
  run<6, 4>(d, "wait + fence::after + MMAs (no commit)");
  run<4, 4>(d, "fence::after + MMAs");
  run<9, 4>(d, "test_wait(all lanes) + MMAs + commit");
  run<17, 4>(d, "test_wait(lane 0) + MMAs + commit");
  return 0;
}
