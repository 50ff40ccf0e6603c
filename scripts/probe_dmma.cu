// probe_dmma.cu — fp64 tensor-pipe throughput on B200 by mma.sync shape (the TRSM and the Cholesky updates are built on DMMA):
// every warp runs 8 independent accumulator chains; report cycles per instruction and FMA / clk / SM at 4, 8, 16 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/probe_dmma scripts/probe_dmma.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int SHAPE>
__global__ void k(double *out, int iters, long long *cyc) {
  double d[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = threadIdx.x * 1e-9 + i;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + 1e-6 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = 1.0 - 1e-6 * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a[0]), "d"(b[0]));
      else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(d[i][0]), "+d"(d[i][1]), "+d"(d[i][2]), "+d"(d[i][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int SHAPE>
void run(const char *name, int fma_per_inst, int dep_test) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    k<SHAPE><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<SHAPE><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_inst_sm = (double)h / (iters * 8.0 * warps);
    printf("%-10s %2d warps/SM: %7.2f cycles per warp-instruction issued SM-wide -> %6.1f FMA/clk/SM (%5.1f TFLOP/s at 1.965 GHz x 148)\n", name, warps,
           per_inst_sm, fma_per_inst / per_inst_sm, 2.0 * fma_per_inst / per_inst_sm * 1.965e9 * 148 / 1e12);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("m8n8k4", 256, 0);
  run<1>("m16n8k4", 512, 0);
  run<2>("m16n8k8", 1024, 0);
  run<3>("m16n8k16", 2048, 0);
  return 0;
}
