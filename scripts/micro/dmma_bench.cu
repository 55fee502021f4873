// Microbenchmark: fp64 throughput of DFMA vs DMMA (mma.sync f64) on sm_100a.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double * out, int iters)
{
  double a0 = threadIdx.x*1e-9, a1 = a0+1, a2 = a0+2, a3 = a0+3, a4 = a0+4, a5 = a0+5, a6 = a0+6, a7 = a0+7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i)
  {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x*blockDim.x+threadIdx.x] = a0+a1+a2+a3+a4+a5+a6+a7;
}
__global__ void dmma884_kernel(double * out, int iters)
{
  double a = threadIdx.x*1e-9, b = 1.0000001;
  double c[8][2];
  for (int q = 0; q < 8; ++q) { c[q][0] = q; c[q][1] = q+0.5; }
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
  }
  double s = 0; for (int q = 0; q < 8; ++q) s += c[q][0]+c[q][1];
  out[blockIdx.x*blockDim.x+threadIdx.x] = s;
}
__global__ void dmma16816_kernel(double * out, int iters)
{
  double a[8], b[4];
  for (int q = 0; q < 8; ++q) a[q] = threadIdx.x*1e-9+q;
  for (int q = 0; q < 4; ++q) b[q] = 1.0000001+q;
  double c[4][4];
  for (int q = 0; q < 4; ++q) for (int r = 0; r < 4; ++r) c[q][r] = q+r;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
        : "+d"(c[q][0]), "+d"(c[q][1]), "+d"(c[q][2]), "+d"(c[q][3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0; for (int q = 0; q < 4; ++q) for (int r = 0; r < 4; ++r) s += c[q][r];
  out[blockIdx.x*blockDim.x+threadIdx.x] = s;
}
template <class F> float timeit(F f)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
  double * out; cudaMalloc(&out, 148*8*1024*sizeof(double));
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2)
  {
    const int threads = warps*32, blocks = 148;
    float ms = timeit([&]{ dfma_kernel<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DFMA        %8.2f TFLOP/s\n", warps, 2.0*8*iters*(double)threads*blocks/ms*1e-9);
    ms = timeit([&]{ dmma884_kernel<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA m8n8k4  %8.2f TFLOP/s\n", warps, 2.0*8*8*4*8*iters*(double)warps*blocks/ms*1e-9);
    ms = timeit([&]{ dmma16816_kernel<<<blocks, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA m16n8k16 %8.2f TFLOP/s\n", warps, 2.0*16*8*16*4*iters*(double)warps*blocks/ms*1e-9);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
