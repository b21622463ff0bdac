// Host-side conversion between the kernels' knot-minor block layout (BatchPtrs, common.cuh) and the dense knot-major arrays of
// the test hook gusto_get_blocks (include/gusto_b200.h): f[B][N][NX], A[B][N][NX][NX] row-major, g[B][N][NX],
// rows[B][N][n_obs][5].  Plain host code, shared by the C ABI and the test-only host simulation.
#pragma once
#include "common.cuh"

namespace gusto {

template <int M>
inline void blocks_unpack(int B, int N, int n_obs, const double* fc, const double* Ac, const double* gc, const double* rc,
                          double* f, double* A, double* g, double* rows) {
  using T = Traits<M>;
  constexpr int NX = T::NX, ANZ = T::ANZ;
  const size_t np = g_np(N), fs = (size_t)n_obs * np;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < N; ++k) {
      const size_t gk = (size_t)b * N + k;
      for (int i = 0; i < NX; ++i) {
        if (f && fc) f[gk * NX + i] = fc[((size_t)b * NX + i) * np + k];
        if (g && gc) g[gk * NX + i] = gc[((size_t)b * NX + i) * np + k];
      }
      if (A && Ac) {
        for (int i = 0; i < NX * NX; ++i) A[gk * NX * NX + i] = 0.0;
        for (int e = 0; e < ANZ; ++e) A[gk * NX * NX + T::a_row(e) * NX + T::a_col(e)] = Ac[((size_t)b * ANZ + e) * np + k];
      }
      if (rows && rc)
        for (int i = 0; i < n_obs; ++i)
          for (int q = 0; q < 5; ++q) rows[(gk * n_obs + i) * 5 + q] = rc[(size_t)b * 5 * fs + q * fs + (size_t)i * np + k];
    }
}

template <int M>
inline void blocks_pack(int B, int N, int n_obs, const double* f, const double* A, const double* g, const double* rows,
                        double* fc, double* Ac, double* gc, double* rc) {
  using T = Traits<M>;
  constexpr int NX = T::NX, ANZ = T::ANZ;
  const size_t np = g_np(N), fs = (size_t)n_obs * np;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < N; ++k) {
      const size_t gk = (size_t)b * N + k;
      for (int i = 0; i < NX; ++i) {
        fc[((size_t)b * NX + i) * np + k] = f[gk * NX + i];
        gc[((size_t)b * NX + i) * np + k] = g[gk * NX + i];
      }
      for (int e = 0; e < ANZ; ++e) Ac[((size_t)b * ANZ + e) * np + k] = A[gk * NX * NX + T::a_row(e) * NX + T::a_col(e)];
      for (int i = 0; i < n_obs; ++i)
        for (int q = 0; q < 5; ++q) rc[(size_t)b * 5 * fs + q * fs + (size_t)i * np + k] = rows[(gk * n_obs + i) * 5 + q];
    }
}

template <int M> constexpr int blocks_anz() { return Traits<M>::ANZ; }

}  // namespace gusto
