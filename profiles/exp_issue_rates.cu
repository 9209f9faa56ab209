// Microbenchmark: warp-instruction issue rates of the candidate inner products of the L1 sweep
// (FADD + FADD|.|  vs  VABSDIFF (int32 |a-b|+c in one instruction)  vs a mix of the two).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scikit-kge_b200/build/exp/issue_rates profiles/exp_issue_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const int *in, int *out, int iters) {
  // 8 x 8 accumulators per thread like the sweep; operands change every iteration (from registers)
  int q[8], e[8];
  for (int i = 0; i < 8; ++i) { q[i] = in[threadIdx.x + 32 * i]; e[i] = in[threadIdx.x + 32 * (8 + i)]; }
  float accf[8][8];
  unsigned acci[8][8];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { accf[i][j] = 0.f; acci[i][j] = 0u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MODE == 0) {
          accf[i][j] += fabsf(__int_as_float(q[i]) - __int_as_float(e[j]));
        } else if (MODE == 1) {
          acci[i][j] = __sad(q[i], e[j], acci[i][j]);
        } else {
          // two thirds of the elements on the integer path, one third on the fp32 path
          if ((i * 8 + j) % 3 != 2) acci[i][j] = __sad(q[i], e[j], acci[i][j]);
          else accf[i][j] += fabsf(__int_as_float(q[i]) - __int_as_float(e[j]));
        }
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) { q[i] += 0x01000001; e[i] ^= it; }
  }
  float sf = 0.f; unsigned si = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { sf += accf[i][j]; si += acci[i][j]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (int)sf + (int)si;
}

template <int MODE>
void run(const char *name, const int *in, int *out) {
  const int iters = 20000, blocks = 148 * 2;
  k<MODE><<<blocks, 256>>>(in, out, 100);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, 256>>>(in, out, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double elems = (double)blocks * 256 * 64.0 * iters;
  printf("%-28s %8.3f ms  %.3f T elements/s  (= %.1f elements / clk / SM at 1.965 GHz)\n", name, ms, elems / ms / 1e9,
         elems / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
  int *in, *out;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 1, 4096 * 4);
  cudaMalloc(&out, 148 * 2 * 256 * 4);
  run<0>("FADD + FADD|.| (fp32)", in, out);
  run<1>("VABSDIFF (int32)", in, out);
  run<2>("2/3 VABSDIFF + 1/3 fp32", in, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
