// Micro-benchmark of tcgen05.mma kind::tf32 issue patterns (one CTA, operands resident in smem / TMEM).
// Straight-line groups of 16 MMAs issued by the elected lane of a converged warp.
#include <cstdio>
#include <cstdlib>
#include "../www2023tiger_b200/csrc/umma.cuh"

template <int N, int NACC, bool TS>
__global__ void __launch_bounds__(128, 1) bench(int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) f[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = __shfl_sync(0xffffffffu, slot, 0);
  if (threadIdx.x < 32) {
    const uint32_t base = smem_addr_u32(smem);
    const uint32_t a_lo = umma_desc_lo(base, 128), b_lo = umma_desc_lo(base + 64 * 1024, N);
    const uint32_t idesc = umma_idesc_tf32(128, N);
    const uint32_t a_t = taddr + 448;   // activations in TMEM (TS form)
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t d = taddr + (uint32_t)((j % NACC) * N);
          if (TS) umma_tf32_ts(d, a_t + (j & 3) * 8, b_lo + (j & 3) * 2 * N, idesc, 1u);
          else umma_tf32_lo(d, a_lo + (j & 3) * 256, b_lo + (j & 3) * 2 * N, idesc, 1u);
        }
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

template <int N, int NACC, bool TS>
void run(long long* d) {
  cudaFuncSetAttribute(bench<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 64;
  bench<N, NACC, TS><<<1, 128, 200 * 1024>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h;
  cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
  printf("%s N %3d accumulators %d: %6.1f cycles/MMA (math floor %d)\n", TS ? "TS" : "SS", N, NACC,
         (double)h / (reps * 16), 128 * N / 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, sizeof(long long));
  run<32, 1, false>(d); run<32, 2, false>(d); run<32, 4, false>(d);
  run<32, 1, true>(d);  run<32, 2, true>(d);  run<32, 4, true>(d);
  run<64, 1, false>(d); run<64, 4, false>(d); run<64, 1, true>(d); run<64, 4, true>(d);
  run<96, 1, false>(d); run<96, 4, false>(d); run<96, 1, true>(d); run<96, 4, true>(d);
  run<128, 1, false>(d); run<128, 2, false>(d); run<128, 1, true>(d); run<128, 2, true>(d);
  return 0;
}
