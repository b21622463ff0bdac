#include <cstdio>
#include <cuda_runtime.h>
// dependent-chain latencies and 1-warp throughputs on B200
__global__ void k_dfma_lat(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dfma_tp(double* out, long long* cyc, int n) {
  double a0 = out[0], a1 = out[1], a2 = out[2], a3 = out[3], a4 = out[4], a5 = out[5], a6 = out[6], a7 = out[7], b = out[8], c = out[9];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c); }
  long long t1 = clock64();
  out[10] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_rcp_lat(double* out, long long* cyc, int n) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = __drcp_rn(a) + 1.0; }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_div_lat(double* out, long long* cyc, int n) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = 1.0 / a + 1.0; }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_sqrt_lat(double* out, long long* cyc, int n) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = sqrt(a) + 1.0; }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_chain(double* out, long long* cyc, int n) {
  __shared__ double s[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s[i] = (double)((i * 7 + 1) & 255);
  __syncthreads();
  int idx = threadIdx.x & 31;
  double acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { double v = s[idx]; acc = fma(v, 1.0000001, acc); idx = (int)v & 255; __syncwarp(); }
  long long t1 = clock64();
  out[3] = acc; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc, int n) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { __syncthreads(); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_sts_lds_sync(double* out, long long* cyc, int n) {   // the elimination step pattern: LDS, DFMA, STS, syncwarp
  __shared__ double s[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s[i] = 1.0 + i * 1e-3;
  __syncthreads();
  const int l = threadIdx.x & 31;
  long long t0 = clock64();
  if (threadIdx.x < 32) for (int i = 0; i < n; ++i) {
    const int q = i & 7;
    double p = s[q * 15], m = s[l * 14 + q] * p, w = s[64 + l * 3 + q];
    s[256 + l + ((i & 1) << 5)] = fma(-m, w, s[256 + l + ((i & 1) << 5)]);
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_gld_chain(const int* next, double* out, long long* cyc, int n) {
  int idx = blockIdx.x * 4099 % (1 << 24);
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) idx = next[idx];
  long long t1 = clock64();
  out[3] = idx; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class F> void run(const char* name, F launch, int grid, int n, int per) {
  long long* cyc; cudaMalloc(&cyc, grid * sizeof(long long));
  launch(cyc); cudaDeviceSynchronize(); launch(cyc); cudaDeviceSynchronize();
  long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < grid; ++i) s += h[i];
  printf("%-28s grid %5d : %.1f cycles per op\n", name, grid, s / grid / n / per);
  cudaFree(cyc); delete[] h;
}
int main() {
  double* out; cudaMalloc(&out, 64 * sizeof(double));
  double h[16]; for (int i = 0; i < 16; ++i) h[i] = 1.0 + 1e-9 * i; cudaMemcpy(out, h, sizeof(h), cudaMemcpyHostToDevice);
  const int n = 2000;
  for (int grid : {1, 148 * 7}) {
    for (int thr : {32, 64}) {
      printf("--- grid %d threads %d\n", grid, thr);
      run("dfma dependent", [&](long long* c) { k_dfma_lat<<<grid, thr>>>(out, c, n); }, grid, n, 4);
      run("dfma 8 independent", [&](long long* c) { k_dfma_tp<<<grid, thr>>>(out, c, n); }, grid, n, 8);
      run("drcp_rn + add dependent", [&](long long* c) { k_rcp_lat<<<grid, thr>>>(out, c, n); }, grid, n, 1);
      run("1/x + add dependent", [&](long long* c) { k_div_lat<<<grid, thr>>>(out, c, n); }, grid, n, 1);
      run("sqrt + add dependent", [&](long long* c) { k_sqrt_lat<<<grid, thr>>>(out, c, n); }, grid, n, 1);
      run("lds->fma->idx + syncwarp", [&](long long* c) { k_lds_chain<<<grid, thr>>>(out, c, n); }, grid, n, 1);
      run("syncthreads", [&](long long* c) { k_bar<<<grid, thr>>>(out, c, n); }, grid, n, 1);
      run("elim-step pattern", [&](long long* c) { k_sts_lds_sync<<<grid, thr>>>(out, c, n); }, grid, n, 1);
    }
  }
  // global pointer chase over 64 MB and 1 GB (L2 / DRAM latency)
  for (size_t words : {(size_t)1 << 22, (size_t)1 << 28}) {
    int* nx; cudaMalloc(&nx, words * sizeof(int));
    int* hn = new int[words];
    for (size_t i = 0; i < words; ++i) hn[i] = (int)((i * 1103515245ull + 12345ull) % words);
    cudaMemcpy(nx, hn, words * sizeof(int), cudaMemcpyHostToDevice);
    for (int grid : {1, 148 * 7}) {
      char nm[64]; snprintf(nm, 64, "gld chase %zu MB", words * 4 >> 20);
      run(nm, [&](long long* c) { k_gld_chain<<<grid, 32>>>(nx, out, c, 500); }, grid, 500, 1);
    }
    cudaFree(nx); delete[] hn;
  }
  return 0;
}
