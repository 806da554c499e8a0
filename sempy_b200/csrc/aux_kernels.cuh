// Set-up and single-element helper kernels: the unassembled operator diagonal (Jacobi), the batched
// gradient / gradient_transpose / kron of sempy/gradient.py and sempy/kron.py.
#pragma once

#include "common.cuh"

namespace semb {

// Unassembled diagonal of the element operator.
__global__ void diag_kernel(const double* __restrict__ G6, double* __restrict__ diag, int64_t E, int Nq,
                            const __grid_constant__ DMat Dp) {
  const int64_t Np = (int64_t)Nq * Nq * Nq;
  const int64_t n = E * Np;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = q / Np;
    const int node = (int)(q - e * Np);
    const int i = node % Nq, j = (node / Nq) % Nq, k = node / (Nq * Nq);
    const double* ge = G6 + e * 6 * Np;
    double s = 0.0;
    for (int m = 0; m < Nq; ++m) {
      const double dmi = Dp.d[m * Nq + i], dmj = Dp.d[m * Nq + j], dmk = Dp.d[m * Nq + k];
      s += dmi * dmi * ge[0 * Np + k * Nq * Nq + j * Nq + m];
      s += dmj * dmj * ge[3 * Np + k * Nq * Nq + m * Nq + i];
      s += dmk * dmk * ge[5 * Np + m * Nq * Nq + j * Nq + i];
    }
    const double dii = Dp.d[i * Nq + i], djj = Dp.d[j * Nq + j], dkk = Dp.d[k * Nq + k];
    s += 2.0 * (dii * djj * ge[1 * Np + node] + dii * dkk * ge[2 * Np + node] + djj * dkk * ge[4 * Np + node]);
    diag[q] = s;
  }
}

// Batched sempy/gradient.py:6-23 / :39-56 for any Nq (set-up helper; one CTA per element).
template <bool TRANSPOSE>
__global__ void gradient_kernel(const double* __restrict__ a0, const double* __restrict__ a1,
                                const double* __restrict__ a2, double* __restrict__ o0, double* __restrict__ o1,
                                double* __restrict__ o2, int Nq, const __grid_constant__ DMat Dp) {
  extern __shared__ double sm[];
  const int Np = Nq * Nq * Nq;
  const int64_t e = blockIdx.x;
  double* s0 = sm;
  double* s1 = sm + Np;
  double* s2 = sm + 2 * Np;
  for (int q = threadIdx.x; q < Np; q += blockDim.x) {
    s0[q] = a0[e * Np + q];
    if (TRANSPOSE) {
      s1[q] = a1[e * Np + q];
      s2[q] = a2[e * Np + q];
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Np; q += blockDim.x) {
    const int i = q % Nq, j = (q / Nq) % Nq, k = q / (Nq * Nq);
    if (!TRANSPOSE) {
      double r = 0.0, s = 0.0, t = 0.0;
      for (int m = 0; m < Nq; ++m) {
        r = fma(Dp.d[i * Nq + m], s0[k * Nq * Nq + j * Nq + m], r);
        s = fma(Dp.d[j * Nq + m], s0[k * Nq * Nq + m * Nq + i], s);
        t = fma(Dp.d[k * Nq + m], s0[m * Nq * Nq + j * Nq + i], t);
      }
      o0[e * Np + q] = r;
      o1[e * Np + q] = s;
      o2[e * Np + q] = t;
    } else {
      double r = 0.0, s = 0.0, t = 0.0;
      for (int m = 0; m < Nq; ++m) {
        r = fma(Dp.d[m * Nq + i], s0[k * Nq * Nq + j * Nq + m], r);
        s = fma(Dp.d[m * Nq + j], s1[k * Nq * Nq + m * Nq + i], s);
        t = fma(Dp.d[m * Nq + k], s2[m * Nq * Nq + j * Nq + i], t);
      }
      o0[e * Np + q] = (r + s) + t;
    }
  }
}

// Batched sempy/kron.py:17-43: out = (Sz (x) Sy (x) Sx) U, one CTA per input.
__global__ void kron_kernel(const double* __restrict__ Sz, int nz, int mz, const double* __restrict__ Sy, int ny,
                            int my, const double* __restrict__ Sx, int nx, int mx, const double* __restrict__ U,
                            double* __restrict__ out) {
  extern __shared__ double sm[];
  double* sU = sm;                       // mz*my*mx
  double* t1 = sU + mz * my * mx;        // mz*my*nx
  double* t2 = t1 + mz * my * nx;        // mz*ny*nx
  const int64_t e = blockIdx.x;
  const int nin = mz * my * mx, nout = nz * ny * nx;
  for (int q = threadIdx.x; q < nin; q += blockDim.x) sU[q] = U[e * nin + q];
  __syncthreads();
  for (int q = threadIdx.x; q < mz * my * nx; q += blockDim.x) {
    const int c = q % nx, ij = q / nx;
    double s = 0.0;
    for (int k = 0; k < mx; ++k) s = fma(Sx[c * mx + k], sU[ij * mx + k], s);
    t1[q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < mz * ny * nx; q += blockDim.x) {
    const int c = q % nx, b = (q / nx) % ny, i = q / (nx * ny);
    double s = 0.0;
    for (int j = 0; j < my; ++j) s = fma(Sy[b * my + j], t1[(i * my + j) * nx + c], s);
    t2[q] = s;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nout; q += blockDim.x) {
    const int bc = q % (ny * nx), a = q / (ny * nx);
    double s = 0.0;
    for (int i = 0; i < mz; ++i) s = fma(Sz[a * mz + i], t2[i * ny * nx + bc], s);
    out[e * nout + q] = s;
  }
}

}  // namespace semb
