// Mesh set-up on the device ("next" row 1): global numbering by coordinate
// sort and the geometric factors.  Both reproduce the reference rules:
//   numbering  sempy/mesh.py:389-435  stable lexicographic sort by exact (x, y, z); new id when the
//              SQUARED distance to the next sorted point exceeds tol2; ids from 1
//   factors    sempy/mesh.py:304-356  G = (grad a . grad b) * B * J, same operation order
#pragma once

#include <cub/cub.cuh>

#include "common.cuh"

namespace semb {

// IEEE double -> unsigned key with the same order; -0.0 and +0.0 map to the same key (they compare
// equal in the reference's sort).
__device__ __forceinline__ unsigned long long sortable_key(double v) {
  if (v == 0.0) v = 0.0;  // canonical +0
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void make_keys_kernel(const double* __restrict__ c, const int64_t* __restrict__ perm,
                                 unsigned long long* __restrict__ keys, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    keys[q] = sortable_key(c[perm ? perm[q] : q]);
}
__global__ void iota_kernel(int64_t* __restrict__ v, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) v[q] = q;
}
// flag[q] = 1 where sorted point q starts a new global node (q = 0 counts as a start)
__global__ void gap_flags_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                 const int64_t* __restrict__ order, double tol2, int* __restrict__ flag, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (q == 0) {
      flag[q] = 1;
      continue;
    }
    const int64_t a = order[q - 1], b = order[q];
    const double dx = x[a] - x[b], dy = y[a] - y[b], dz = z[a] - z[b];
    // (dx^2 + dy^2) + dz^2, no contraction: the comparison must not depend on FMA rounding
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    flag[q] = d2 > tol2 ? 1 : 0;
  }
}
__global__ void scatter_ids_kernel(const int64_t* __restrict__ order, const int* __restrict__ idscan,
                                   int32_t* __restrict__ gid, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    gid[order[q]] = idscan[q];  // inclusive scan of the start flags = 1-based id
}

// Geometric factors of one element per CTA: coordinates staged in shared memory, nine derivatives per
// node with D from the parameter block, then the reference's formulas in the reference's order.
__global__ void geom_factors_kernel(const double* __restrict__ xe, const double* __restrict__ ye,
                                    const double* __restrict__ ze, double* __restrict__ G6, double* __restrict__ jaco,
                                    int Nq, const __grid_constant__ DMat Dp, const __grid_constant__ DMat Wp) {
  extern __shared__ double sm[];
  const int nn = Nq * Nq, Np = nn * Nq;
  const int64_t e = blockIdx.x;
  double* sx = sm;
  double* sy = sm + Np;
  double* sz = sm + 2 * Np;
  for (int q = threadIdx.x; q < Np; q += blockDim.x) {
    sx[q] = xe[e * Np + q];
    sy[q] = ye[e * Np + q];
    sz[q] = ze[e * Np + q];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < Np; q += blockDim.x) {
    const int i = q % Nq, j = (q / Nq) % Nq, k = q / nn;
    double xr = 0, xs = 0, xt = 0, yr = 0, ys = 0, yt = 0, zr = 0, zs = 0, zt = 0;
    for (int m = 0; m < Nq; ++m) {
      const double di = Dp.d[i * Nq + m], dj = Dp.d[j * Nq + m], dk = Dp.d[k * Nq + m];
      const int qr = k * nn + j * Nq + m, qs = k * nn + m * Nq + i, qt = m * nn + j * Nq + i;
      xr = fma(di, sx[qr], xr);
      yr = fma(di, sy[qr], yr);
      zr = fma(di, sz[qr], zr);
      xs = fma(dj, sx[qs], xs);
      ys = fma(dj, sy[qs], ys);
      zs = fma(dj, sz[qs], zs);
      xt = fma(dk, sx[qt], xt);
      yt = fma(dk, sy[qt], yt);
      zt = fma(dk, sz[qt], zt);
    }
    const double J = xr * (ys * zt - yt * zs) - yr * (xs * zt - xt * zs) + zr * (xs * yt - ys * xt);
    const double rx = (ys * zt - yt * zs) / J, sxx = (yt * zr - yr * zt) / J, tx = (yr * zs - ys * zr) / J;
    const double ry = -(zt * xs - zs * xt) / J, syy = -(zr * xt - zt * xr) / J, ty = -(zs * xr - zr * xs) / J;
    const double rz = (xs * yt - xt * ys) / J, szz = -(xr * yt - xt * yr) / J, tz = (xr * ys - xs * yr) / J;
    const double B = (Wp.d[i] * Wp.d[j]) * Wp.d[k];
    double* g = G6 + e * 6 * (int64_t)Np;
    g[0 * Np + q] = (rx * rx + ry * ry + rz * rz) * B * J;
    g[1 * Np + q] = (rx * sxx + ry * syy + rz * szz) * B * J;
    g[2 * Np + q] = (rx * tx + ry * ty + rz * tz) * B * J;
    g[3 * Np + q] = (sxx * sxx + syy * syy + szz * szz) * B * J;
    g[4 * Np + q] = (sxx * tx + syy * ty + szz * tz) * B * J;
    g[5 * Np + q] = (tx * tx + ty * ty + tz * tz) * B * J;
    if (jaco) jaco[e * Np + q] = J;
  }
}

}  // namespace semb
