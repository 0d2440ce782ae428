// tools/ubench_f32x2.cu -- does Blackwell's packed fp32 (mul/add.rn.f32x2) double the
// non-fused FP32 rate on B200?  Informs the FIR inner loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void k(float *out, int iters, float seed)
{
  // 8 independent chains per thread
  if (MODE == 0) {            // scalar: mul + add (non fused), 16 flop-instr per iter
    float a[8], t = seed;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { float p = __fmul_rn(a[i], t); a[i] = __fadd_rn(a[i], p); }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else if (MODE == 1) {     // packed: mul2 + add2, 8 independent 64-bit chains = 32 flop per iter in 16 instr
    unsigned long long a[8], t;
    float2 tt = make_float2(seed, seed * 0.5f);
    t = *reinterpret_cast<unsigned long long *>(&tt);
#pragma unroll
    for (int i = 0; i < 8; i++) { float2 v = make_float2(seed + i, seed - i); a[i] = *reinterpret_cast<unsigned long long *>(&v); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { unsigned long long p = mul2(a[i], t); a[i] = add2(a[i], p); }
    }
    unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  } else if (MODE == 2) {     // scalar FFMA
    float a[8], t = seed;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { a[i] = fmaf(a[i], t, 1.0f); a[i] = fmaf(a[i], t, 0.5f); }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {                    // packed FFMA2
    unsigned long long a[8], t, c;
    float2 tt = make_float2(seed, seed * 0.5f), cc = make_float2(1.0f, 0.5f);
    t = *reinterpret_cast<unsigned long long *>(&tt);
    c = *reinterpret_cast<unsigned long long *>(&cc);
#pragma unroll
    for (int i = 0; i < 8; i++) { float2 v = make_float2(seed + i, seed - i); a[i] = *reinterpret_cast<unsigned long long *>(&v); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { a[i] = fma2(a[i], t, c); a[i] = fma2(a[i], t, c); }
    }
    unsigned long long s = 0; for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
  }
}

template <int MODE>
void run(const char *name, double flop_per_instr)
{
  float *out; cudaMalloc(&out, 148 * 8 * 512 * sizeof(float));
  const int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 512>>>(out, 16, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 512>>>(out, iters, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 148.0 * 8 * 512 * iters * 16;     // thread-level instructions
  printf("%-14s %8.3f ms  %7.2f T thread-instr/s  %7.2f TFLOP/s(ops)\n", name, ms, instr / ms / 1e9, instr * flop_per_instr / ms / 1e9);
  cudaFree(out);
}

int main()
{
  run<0>("FMUL+FADD", 1);
  run<1>("FMUL2+FADD2", 2);
  run<2>("FFMA", 2);
  run<3>("FFMA2", 4);
  return 0;
}
