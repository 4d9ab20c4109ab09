// FP64 pipe microbenchmarks for B200 (sm_100a): the roofline denominator for the
// fused PDIPM kernels is the FP64 arithmetic pipe, which MEASURED_PEAKS.json does not hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
// Prints one JSON line: DFMA TFLOP/s, DMMA(m8n8k4) TFLOP/s, FFMA TFLOP/s, dependent-chain
// latencies (cycles) of DFMA, ddiv, dsqrt, 64-bit shuffle, and 64-bit smem load.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 12345.678f) out[0] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// dependent-chain latency probes, one warp
__global__ void lat_kernel(long long* out, double* sink, double x0) {
  __shared__ double sm[64];
  sm[threadIdx.x] = x0 + threadIdx.x; sm[threadIdx.x + 32] = 1.0;
  __syncthreads();
  const int N = 256;
  double x = x0 + 1e-3 * threadIdx.x;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) x = fma(x, 1.0000001, 1e-9);
  t1 = clock64(); if (threadIdx.x == 0) out[0] = (t1 - t0);
  // ddiv
  double y = x;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) y = 3.0000001 / y;
  t1 = clock64(); if (threadIdx.x == 0) out[1] = (t1 - t0);
  // dsqrt
  double z = fabs(y) + 2.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) z = sqrt(z) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) out[2] = (t1 - t0);
  // 64-bit shuffle
  double w = z;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) w = __shfl_sync(0xffffffffu, w, (threadIdx.x + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) out[3] = (t1 - t0);
  // smem pointer chase (64-bit)
  int idx = threadIdx.x;
  __shared__ int chase[64];
  chase[threadIdx.x] = (threadIdx.x + 7) & 31; chase[threadIdx.x + 32] = 0;
  __syncthreads();
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) idx = chase[idx];
  t1 = clock64(); if (threadIdx.x == 0) out[4] = (t1 - t0);
  // rsqrt-based: drsqrt
  double r = fabs(w) + 2.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) r = rsqrt(r) + 1.5;
  t1 = clock64(); if (threadIdx.x == 0) out[5] = (t1 - t0);
  // DMMA dependent chain
  double c0 = r, c1 = w;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(1e-9), "d"(1e-9));
  t1 = clock64(); if (threadIdx.x == 0) out[6] = (t1 - t0);
  sink[threadIdx.x] = x + y + z + w + idx + r + c0 + c1 + sm[threadIdx.x];
}

// smem bandwidth: 64-bit conflict-free loads
__global__ void lds_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  int base = threadIdx.x;
  for (int it = 0; it < iters; it++) {
    acc0 += sm[(base) & 4095];
    acc1 += sm[(base + 1024) & 4095];
    acc2 += sm[(base + 2048) & 4095];
    acc3 += sm[(base + 3072) & 4095];
    base += 32;
  }
  if (acc0 + acc1 + acc2 + acc3 == 1.2345) out[0] = acc0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  double* d; CK(cudaMalloc(&d, 1 << 20));
  long long* dl; CK(cudaMalloc(&dl, 64 * sizeof(long long)));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 4096;
  int blocks = sms * 8, threads = 256;
  auto flops = [&](double per_thread_per_iter) { return (double)blocks * threads * iters * per_thread_per_iter; };
  double best_dfma = 0, best_dmma = 0, best_ffma = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0); dfma_kernel<8><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    double tf = flops(8 * 2) / (ms * 1e-3) / 1e12; if (tf > best_dfma) best_dfma = tf;
    cudaEventRecord(e0); dmma_kernel<8><<<blocks, threads>>>(d, iters, 1e-9, 1e-9); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    // one m8n8k4 = 8*8*4 FMA per warp = 16 flop per thread
    tf = flops(8 * 16) / (ms * 1e-3) / 1e12; if (tf > best_dmma) best_dmma = tf;
    cudaEventRecord(e0); ffma_kernel<8><<<blocks, threads>>>((float*)d, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    tf = flops(8 * 2) / (ms * 1e-3) / 1e12; if (tf > best_ffma) best_ffma = tf;
  }
  // sustained DFMA for ~2 s to see power-capped clocks
  double sustained = 0;
  {
    int reps = 0; cudaEventRecord(e0);
    for (; reps < 400; reps++) dfma_kernel<8><<<blocks, threads>>>(d, iters * 4, 1.0000001, 1e-9);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    sustained = (double)reps * blocks * threads * (iters * 4.0) * 16 / (ms * 1e-3) / 1e12;
  }
  lat_kernel<<<1, 32>>>(dl, d, 1.5); CK(cudaDeviceSynchronize());
  long long hl[8]; CK(cudaMemcpy(hl, dl, sizeof(hl), cudaMemcpyDeviceToHost));
  // smem bandwidth
  double lds_gbs = 0;
  {
    int it2 = 8192; int thr = 1024;
    cudaEventRecord(e0); lds_kernel<<<sms, thr, 32768>>>(d, it2); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    cudaEventRecord(e0); lds_kernel<<<sms, thr, 32768>>>(d, it2); cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    lds_gbs = (double)sms * thr * it2 * 4 * 8 / (ms * 1e-3) / 1e9;
  }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, \"dfma_tflops\": %.2f, \"dfma_tflops_sustained\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, \"ffma_tflops\": %.2f, "
         "\"lat_cycles\": {\"dfma\": %.1f, \"ddiv\": %.1f, \"dsqrt_plus_add\": %.1f, \"shfl64\": %.1f, \"lds32_chase\": %.1f, \"drsqrt_plus_add\": %.1f, \"dmma\": %.1f}, \"smem_lds64_gbs\": %.0f}\n",
         prop.name, sms, clk, best_dfma, sustained, best_dmma, best_ffma,
         hl[0] / 256.0, hl[1] / 256.0, hl[2] / 256.0, hl[3] / 256.0, hl[4] / 256.0, hl[5] / 256.0, hl[6] / 256.0, lds_gbs);
  return 0;
}
