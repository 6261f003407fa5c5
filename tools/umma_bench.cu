// Micro-benchmark of tcgen05.mma kind::tf32 issue patterns (one CTA per SM, operands resident in smem).
#include <cstdio>
#include <cstdlib>
#include "../www2023tiger_b200/csrc/umma.cuh"

struct Cfg { int n; int n_acc; int reps; int layout; int a_rows; int commit_every; int kind; int m; };

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* f = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) f[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t taddr = slot;
  if (threadIdx.x == 0) {
    const uint32_t base = smem_addr_u32(smem);
    const uint32_t a_addr = base, b_addr = base + 64 * 1024;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint64_t lt = 0;
    if (c.layout == 0) { a_lbo = 128 * 16; a_sbo = 128; b_lbo = c.n * 16; b_sbo = 128; }
    else if (c.layout == 1) { a_lbo = 128; a_sbo = 256; b_lbo = 128; b_sbo = 256; }
    else { a_lbo = 16; a_sbo = 1024; b_lbo = 16; b_sbo = 1024; lt = 2ull << 61; }   // SWIZZLE_128B K-major
    uint32_t idesc = umma_idesc_tf32(c.m, c.n);
    if (c.kind == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
    uint32_t phase = 0;
    const long long t0 = clock64();
    for (int r = 0; r < c.reps; ++r) {
      const uint32_t step = c.layout == 2 ? 32u : 4096u;
      const uint64_t ad = umma_smem_desc(a_addr + (r & 3) * step, a_lbo, a_sbo) | lt;
      const uint64_t bd = umma_smem_desc(b_addr + (r & 3) * step, b_lbo, b_sbo) | lt;
      const uint32_t dcol = taddr + (uint32_t)((r % c.n_acc) * c.n);
      const uint32_t acc = r >= c.n_acc;
      if (c.kind == 0) umma_tf32(dcol, ad, bd, idesc, acc);
      else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dcol), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      if (c.commit_every > 0 && (r + 1) % c.commit_every == 0) {
        umma_commit(&bar);
        mbar_wait(&bar, phase);
        phase ^= 1;
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, phase);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(taddr, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ns[] = {32, 128, 256};
  for (int kind = 0; kind < 2; ++kind)
    for (int m : {128, 64})
      for (int layout = 0; layout < 3; ++layout)
        for (int n : ns) {
          Cfg c{n, 1, 384, layout, 128, 0, kind, m};
          bench<<<1, 128, 200 * 1024>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h;
          cudaMemcpy(&h, d, sizeof(long long), cudaMemcpyDeviceToHost);
          printf("kind %s M %3d layout %d N %3d: %7.1f cycles/MMA (floor %d)\n", kind ? "bf16" : "tf32", m, layout, n,
                 (double)h / c.reps, 128 * n / 256);
        }
  return 0;
}
