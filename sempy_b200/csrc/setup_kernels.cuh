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

// gslib-style set-up from ids (GS.setup, sempy/mesh.py:437-438): keys for the stable sort by id, and the
// segment starts of the sorted order (a new segment where the id changes; id 0 = "not shared": every
// such entry is its own segment, as in gslib)
__global__ void id_keys_kernel(const int32_t* __restrict__ ids, unsigned int* __restrict__ keys, int64_t n, int* __restrict__ info) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t id = ids[q];
    if (id < 0) atomicOr(info, 1);
    keys[q] = (unsigned int)id;
  }
}
__global__ void id_flags_kernel(const unsigned int* __restrict__ sorted_keys, int* __restrict__ flag, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    flag[q] = (q == 0 || sorted_keys[q] != sorted_keys[q - 1] || sorted_keys[q] == 0u) ? 1 : 0;
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

// ---------------------------------------------------------------------------
// Gather-scatter schedule on the device, from the numbering arrays (global_to_local, global_start) of
// sempy/mesh.py:422-423.  Only segments with more than one local copy are kept, grouped by length
// (2 | 4 | 8 | anything else) and ordered by their lowest local index inside a group, the copies of a
// segment in ascending local index (the fixed summation order of both the push and the pull kernels).
// ---------------------------------------------------------------------------
constexpr unsigned long long SEG_CLASS_SHIFT = 40;

// key = (class << 40) | lowest local index; class 4 = single-copy segment (sorted to the end, dropped)
__global__ void seg_keys_kernel(const int64_t* __restrict__ g2l, const int64_t* __restrict__ gstart, int64_t nseg, int64_t n,
                                unsigned long long* __restrict__ keys, int32_t* __restrict__ vals, int* __restrict__ info) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = gstart[s], len = gstart[s + 1] - lo;
    unsigned long long cls = 4, first = 0;
    if (len < 1 || lo < 0 || lo + len > n) {
      atomicOr(info + 0, 1);  // malformed offsets
    } else {
      int64_t mn = n;
      bool bad = false;
      for (int64_t q = lo; q < lo + len; ++q) {
        const int64_t id = g2l[q];
        bad |= id < 0 || id >= n;
        mn = id < mn ? id : mn;
      }
      if (bad) {
        atomicOr(info + 0, 2);  // local index out of range
        mn = 0;
      }
      if (len > 1) {
        cls = len == 2 ? 0 : len == 4 ? 1 : len == 8 ? 2 : 3;
        first = (unsigned long long)mn;
        if (len > 15) atomicMax(info + 1, (int)(len > 0x7fffffff ? 0x7fffffff : len));
      }
    }
    keys[s] = (cls << SEG_CLASS_SHIFT) | first;
    vals[s] = (int32_t)s;
  }
}

// bounds[c] = first sorted position whose class is >= c (c = 0..4); bounds[4] = number of shared segments
__global__ void seg_bounds_kernel(const unsigned long long* __restrict__ keys, int64_t nseg, int64_t* __restrict__ bounds) {
  const int c = threadIdx.x;
  if (c > 4) return;
  const unsigned long long target = (unsigned long long)c << SEG_CLASS_SHIFT;
  int64_t lo = 0, hi = nseg;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  bounds[c] = lo;
}

// lengths of the irregular segments (sorted positions [b3, b4)) for the CSR offsets
__global__ void seg_var_len_kernel(const int32_t* __restrict__ vals, const int64_t* __restrict__ gstart, int64_t b3, int64_t b4,
                                   int32_t* __restrict__ len_out) {
  for (int64_t t = b3 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < b4; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = vals[t];
    len_out[t - b3] = (int32_t)(gstart[s + 1] - gstart[s]);
  }
}

struct SegLayout {
  int64_t b[5];     // class boundaries in the sorted order
  int64_t off[4];   // start of each class in ids (multiples of 4 ints)
};

// writes the ids of every shared segment (ascending) and the links of the pull form (vec_kernels.cuh)
__global__ void seg_emit_kernel(const int32_t* __restrict__ vals, const int64_t* __restrict__ g2l,
                                const int64_t* __restrict__ gstart, const SegLayout lay, const int32_t* __restrict__ var_start,
                                int32_t* __restrict__ ids, uint32_t* __restrict__ link) {
  const int64_t total = lay.b[4];
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = vals[t];
    const int64_t lo = gstart[s];
    const int len = (int)(gstart[s + 1] - lo);
    const int cls = t < lay.b[1] ? 0 : t < lay.b[2] ? 1 : t < lay.b[3] ? 2 : 3;
    const int64_t dst = cls < 3 ? lay.off[cls] + (t - lay.b[cls]) * len : lay.off[3] + var_start[t - lay.b[3]];
    int32_t* out = ids + dst;
    // insertion sort straight into the output (segments are short)
    for (int c = 0; c < len; ++c) {
      const int32_t id = (int32_t)g2l[lo + c];
      int pos = c;
      while (pos > 0 && out[pos - 1] > id) {
        out[pos] = out[pos - 1];
        --pos;
      }
      out[pos] = id;
    }
    if (len == 2) {  // pair links of the pull form: each copy points at the other (GS_LINK_PAIR)
      link[out[0]] = 0x80000000u | (uint32_t)out[1];
      link[out[1]] = 0x80000000u | (uint32_t)out[0];
    } else {  // longer segments: every copy points at its segment's slot in the compact sums (GS_LINK_LONG)
      const uint32_t lk = 0x40000000u | (uint32_t)(t - lay.b[1]);
      for (int c = 0; c < len; ++c) link[out[c]] = lk;
    }
  }
}

// ---------------------------------------------------------------------------
// Dirichlet mask of the bounding box on the device (Mesh.setup_mask, sempy/mesh.py:449-477): bit = 1 where
// any coordinate is within tol of the box (the reference's mask is 0 there).
// ---------------------------------------------------------------------------
struct BBox {
  double lo[3], hi[3];
};
__global__ void mask_bits_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                 int64_t n, const BBox bb, double tol, uint32_t* __restrict__ bits, int64_t nwords) {
  // one thread per node, one warp per 32-bit word
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nwords * 32; q += stride) {
    bool m = false;
    if (q < n) {
      const double cx = x[q], cy = y[q], cz = z[q];
      m = fabs(cx - bb.lo[0]) < tol || fabs(cx - bb.hi[0]) < tol || fabs(cy - bb.lo[1]) < tol || fabs(cy - bb.hi[1]) < tol ||
          fabs(cz - bb.lo[2]) < tol || fabs(cz - bb.hi[2]) < tol;
    }
    const unsigned int w = __ballot_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) bits[q >> 5] = w;
  }
}
// x[q] = 0 where bit q is set   (Mesh.apply_mask, sempy/mesh.py:487-494)
__global__ void mask_apply_bits_kernel(double* __restrict__ x, const uint32_t* __restrict__ bits, int64_t n) {
  const int64_t nwords = (n + 31) >> 5;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t b = bits[w];
    while (b) {
      const int k = __ffs(b) - 1;
      b &= b - 1;
      const int64_t q = w * 32 + k;
      if (q < n) x[q] = 0.0;
    }
  }
}
// mask as the reference stores it: 0.0 at Dirichlet nodes, 1.0 elsewhere
__global__ void mask_expand_kernel(const uint32_t* __restrict__ bits, double* __restrict__ mask, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    mask[q] = ((bits[q >> 5] >> (q & 31)) & 1u) ? 0.0 : 1.0;
}

}  // namespace semb
