// peaks.cu -- measured denominators for the rooflines quoted in DESIGN.md / bench.py:
// FP64 FMA (DFMA), FP64 tensor (DMMA m8n8k4), FP32 FMA issue rates and HBM copy bandwidth.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/peaks tools/peaks.cu && gpurun_out/peaks
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <typename T, int CH>
__global__ void k_fma(int iters, T *out, T a, T b) {
  T acc[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c] = (T)(threadIdx.x + c);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) acc[c] = acc[c] * a + b;
  }
  T s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += acc[c];
  if (s == (T)123456789) out[0] = s;
}

template <int CH>
__global__ void k_dmma(int iters, double *out, double a, double b) {
  double c0[CH], c1[CH];
#pragma unroll
  for (int c = 0; c < CH; c++) { c0[c] = threadIdx.x; c1[c] = c; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; c++) s += c0[c] + c1[c];
  if (s == 123456789.0) out[0] = s;
}

__global__ void k_copy(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// shared-memory read bandwidth: every lane reads a distinct 16-byte word per instruction
__global__ void k_lds(int iters, double *out) {
  __shared__ double2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(i, 1.0);
  __syncthreads();
  double s = 0;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) { const double2 v = sm[(idx + 32 * u) & 1023]; s += v.x; idx += (int)v.y; }
  }
  if (s == 123456789.0) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < reps; r++) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *dout;
  CK(cudaMalloc(&dout, 64));
  const int iters = 4096, threads = 256, ctas = sms * 8;
  printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
  {
    float ms = time_ms([&] { k_fma<double, 8><<<ctas, threads>>>(iters, dout, 1.0000001, 1e-9); }, 5);
    printf(", \"dfma_tfma_s\": %.3f", (double)ctas * threads * iters * 8 / (ms * 1e-3) / 1e12);
  }
  {
    float ms = time_ms([&] { k_fma<float, 8><<<ctas, threads>>>(iters, (float *)dout, 1.0000001f, 1e-9f); }, 5);
    printf(", \"ffma_tfma_s\": %.3f", (double)ctas * threads * iters * 8 / (ms * 1e-3) / 1e12);
  }
  {
    float ms = time_ms([&] { k_dmma<8><<<ctas, threads>>>(iters / 4, dout, 1.0000001, 1e-9); }, 5);
    // one m8n8k4 = 256 FMA per warp = 8 FMA per lane
    printf(", \"dmma_tfma_s\": %.3f", (double)ctas * threads * (iters / 4) * 8 * 8 / (ms * 1e-3) / 1e12);
  }
  {
    float ms = time_ms([&] { k_lds<<<ctas, threads>>>(512, dout); }, 5);
    printf(", \"lds128_tb_s\": %.3f", (double)ctas * threads * 512 * 8 * 16 / (ms * 1e-3) / 1e12);
  }
  {
    const size_t bytes = (size_t)2 << 30;
    float4 *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    float ms = time_ms([&] { k_copy<<<sms * 16, 512>>>(a, b, bytes / 16); }, 10);
    printf(", \"hbm_copy_gb_s\": %.1f", 2.0 * bytes / (ms * 1e-3) / 1e9);
    float ms2 = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 10);
    printf(", \"hbm_memcpy_gb_s\": %.1f", 2.0 * bytes / (ms2 * 1e-3) / 1e9);
  }
  printf(", \"sm_clock_khz_prop\": %d}\n", p.clockRate);
  return 0;
}
