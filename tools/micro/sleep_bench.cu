// Developer aid: how long do __nanosleep(t) and mbarrier.try_wait (with / without a suspend-time hint) really block?
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(long long* out)
{
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
  __syncthreads();
  const unsigned ns[6] = {100, 300, 1000, 2000, 10000, 100000};
  for (int i = 0; i < 6; i++) {
    long long t0 = clock64();
    for (int r = 0; r < 16; r++) __nanosleep(ns[i]);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[i] = (t1 - t0) / 16;
  }
  {
    long long t0 = clock64();
    uint32_t ok;
    for (int r = 0; r < 16; r++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) out[6] = (t1 - t0) / 16;
  }
  const unsigned hint[3] = {1000, 10000, 1000000};
  for (int i = 0; i < 3; i++) {
    long long t0 = clock64();
    uint32_t ok;
    for (int r = 0; r < 16; r++)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u), "r"(hint[i]) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) out[7 + i] = (t1 - t0) / 16;
  }
}
int main()
{
  long long* d; cudaMalloc(&d, 80);
  k<<<1, 32>>>(d); cudaDeviceSynchronize();
  k<<<1, 32>>>(d); cudaDeviceSynchronize();
  long long h[10]; cudaMemcpy(h, d, 80, cudaMemcpyDeviceToHost);
  const char* n[10] = {"nanosleep 100", "nanosleep 300", "nanosleep 1000", "nanosleep 2000", "nanosleep 10000", "nanosleep 100000",
                       "try_wait (no hint)", "try_wait hint 1000", "try_wait hint 10000", "try_wait hint 1000000"};
  for (int i = 0; i < 10; i++) printf("%-24s %8lld cycles\n", n[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
