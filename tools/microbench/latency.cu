// Developer tool: dependent-issue latency (cycles per op, one warp) of the operations the render kernel's serial
// chains are made of.  nvcc -arch=sm_100a -fmad=false -O3 -o latency latency.cu && ./latency
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 2048
template <class F>
__device__ __forceinline__ long long timed(F f) {
  long long t0 = clock64();
  f();
  long long t1 = clock64();
  return t1 - t0;
}

__global__ void k(double* sink, long long* out, double a, double b, float fa, int ia, unsigned mask_in) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)(i & 31);
  __syncthreads();
  double x = a + threadIdx.x * 1e-9, y = b;
  float fx = fa;
  int ix = ia + threadIdx.x;
  unsigned m = mask_in;
  long long t;
  int r = 0;
#define RUN(name, body)                                  \
  t = timed([&] {                                        \
    _Pragma("unroll 16") for (int i = 0; i < N; ++i) { body; } \
  });                                                    \
  if (threadIdx.x == 0) out[r] = t;                      \
  ++r;
  RUN("dadd", x = x + y)
  RUN("dmul", x = x * y)
  RUN("dfma", x = fma(x, y, y))
  RUN("dadd+dmul", x = (x + y) * y)
  RUN("ffma", fx = fmaf(fx, fa, fa))
  RUN("fmnmx", fx = fmaxf(fminf(fx, fa), 0.5f * fa))
  RUN("iadd", ix = ix + ia)
  RUN("imad", ix = ix * ia + 3)
  RUN("dsqrt", x = sqrt(x + y))
  RUN("ddiv", x = y / (x + 1.0))
  RUN("drcp", x = 1.0 / (x + 1.0))
  RUN("frcp", fx = __frcp_rn(fx + 1.0f))
  RUN("d2f+f2d", x = (double)__double2float_rn(x) + y)
  RUN("lds", ix = (int)sm[(ix & 1023)] + ia)
  RUN("shfl", ix = __shfl_sync(0xffffffffu, ix, (ix + 1) & 31) + 1)
  RUN("redux_min", ix = (int)__reduce_min_sync(0xffffffffu, (unsigned)ix) + (int)threadIdx.x)
  RUN("ballot+popc", ix = __popc(__ballot_sync(0xffffffffu, (ix & 1) != 0)) + (int)threadIdx.x)
  RUN("ffs", ix = __ffs(ix | 0x100) + ia)
  RUN("popc", ix = __popc(ix) + ia)
  RUN("syncwarp", __syncwarp(); ix += 1)
  RUN("any", ix += __any_sync(0xffffffffu, ix > 5) ? 1 : 2)
  if (threadIdx.x == 0) out[63] = r;
  sink[threadIdx.x] = x + fx + ix + m;
}

int main() {
  const char* names[] = {"dadd", "dmul", "dfma", "dadd+dmul(2 ops)", "ffma", "fmnmx(2 ops)", "iadd", "imad", "dsqrt(+dadd)",
                         "ddiv(+dadd)", "drcp(+dadd)", "frcp(+fadd)", "d2f+f2d+dadd", "lds(+cvt+iadd)", "shfl(+iadd)",
                         "redux_min(+iadd)", "ballot+popc(+iadd)", "ffs(+or+iadd)", "popc(+iadd)", "syncwarp(+iadd)",
                         "any(+sel+iadd)"};
  double* sink;
  long long* out;
  cudaMalloc(&sink, 1024 * 8);
  cudaMalloc(&out, 64 * 8);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(sink, out, 1.0000001, 0.9999999, 1.0001f, 3, 0xf0f0u);
  cudaDeviceSynchronize();
  long long h[64];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < (int)h[63]; ++i) printf("%-22s %7.2f cycles/iter\n", names[i], (double)h[i] / N);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
