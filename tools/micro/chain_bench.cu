// micro-benchmark: dependent FADD chain latency and the in-place running-sum loop (developer aid)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gen2_uhf_rfid_reader_b200/csrc/rx_fused.cuh"
using namespace rfid_b200;

__global__ void k_regchain(float* out, long long* cyc, float x)
{
  float acc = x;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; i++) acc = __fadd_rn(acc, x);
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_smemchain(float* out, long long* cyc, int nlanes)
{
  __shared__ __align__(16) float buf[3][128 + 32];
  for (int i = threadIdx.x; i < 3 * 160; i += 32) (&buf[0][0])[i] = 1e-3f * i;
  __syncwarp();
  float acc = 0.f;
  int lane = threadIdx.x;
  long long t0 = clock64();
  if (lane < nlanes) chain_inplace(buf[lane], 128, acc);
  __syncwarp();
  long long t1 = clock64();
  out[threadIdx.x] = acc + buf[0][5];
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
}

int main()
{
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  long long h[4];
  for (int rep = 0; rep < 2; rep++) {
    k_regchain<<<1, 32>>>(out, cyc, 1.5f);
    k_smemchain<<<1, 32>>>(out, cyc, 3);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("register chain 128 FADD: %lld cycles (%.2f/step); smem running sum 128 steps, 3 lanes: %lld cycles (%.2f/step)\n", h[0], h[0] / 128.0, h[1], h[1] / 128.0);
  }
  k_smemchain<<<1, 32>>>(out, cyc, 1);
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
  printf("smem running sum, 1 lane: %lld cycles (%.2f/step)\n", h[1], h[1] / 128.0);
  return 0;
}
