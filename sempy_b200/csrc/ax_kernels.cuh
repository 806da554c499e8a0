// K1: the fused FP64 spectral-element Poisson operator  w_e = D^T G D u_e.
//
// Replaces the per-element Python loop of sempy/elliptic.py:9-27 (gradient ->
// 3x3 G -> gradient_transpose, sempy/gradient.py:6-23,39-56), with the
// Dirichlet mask (sempy/mesh.py:487-494) and the local part of the pAp dot
// (sempy/elliptic.py:51) folded into the epilogue.
//
// HBM-bound by design: 64 B per node (u 8 + six G components 48 + w 8) against
// 12*Nq+15 flop, so no tensor cores; the work is to keep shared-memory traffic
// and FP64 issue below the HBM time.  See DESIGN.md "K1".
#pragma once

#include "common.cuh"

namespace semb {

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  // G is read exactly once per operator application: streaming (evict-first) loads.
  return __ldcs(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

template <int NT>
__device__ __forceinline__ void block_sum_to(double v, double* dst) {
  __shared__ double s_red[NT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NT / 32; ++q) s += s_red[q];
    *dst = s;
  }
}

// ---------------------------------------------------------------------------
// Nq = 8 (N = 7): one warp per element, whole element staged in shared memory.
//
// Three ownership layouts of the 8x8x8 nodes, all by the same 32 lanes:
//   T: lane owns two adjacent k-columns  (i = 2*ip, 2*ip+1; j)      -> t-derivative in registers
//   R: lane owns two r-lines             (j, k), 8 consecutive i    -> r-derivative in registers
//   S: lane owns two s-lines             (i, k), 8 values of j      -> s-derivative in registers
// D is a kernel parameter, so all 64 DFMAs of a line use constant-bank
// operands.  Between layouts the data crosses shared memory once
// (15 eight-byte accesses per node in total instead of ~32 for a plane sweep).
// Padding: row stride 10 doubles, plane stride 88 doubles makes every access
// pattern below bank-conflict free (R: LDS.128 by 8 lanes with distinct j;
// S: LDS.64 by 16 lanes = 8 i x 2 k; T: LDS.128 by 8 lanes = 4 ip x {j, j+4}).
// ---------------------------------------------------------------------------
constexpr int AX8_RS = 10;
constexpr int AX8_PS = 88;
constexpr int AX8_ARR = 8 * AX8_PS;  // doubles per staged array
constexpr int AX8_WARPS = 4;         // elements per CTA
constexpr int AX8_SMEM = AX8_WARPS * 2 * AX8_ARR * (int)sizeof(double);

// D is replicated once per phase in the parameter block (DMat8x).  With a
// single copy the compiler CSEs the loads of all 64 entries across the five
// phases, tries to keep them live for the whole kernel, runs out of uniform
// registers and spills them into vector registers (255 regs + local memory).
// Private copies make every phase load its entries just in time with LDCU
// into uniform registers, which DFMA reads directly.
// Both lines of a lane go through the contraction together, so each entry of
// D is fetched (LDCU) once and consumed at once by two DFMAs: few uniform
// registers live, no UR spills.
template <bool TRANSPOSE>
__device__ __forceinline__ void line8x2(const double (&D)[64], const double (&a)[2][8], double (&o)[2][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const double d = TRANSPOSE ? D[m * 8 + i] : D[i * 8 + m];
      s0 = fma(d, a[0][m], s0);
      s1 = fma(d, a[1][m], s1);
    }
    o[0][i] = s0;
    o[1][i] = s1;
  }
}

template <bool MASK, bool DOT>
__global__ void __launch_bounds__(AX8_WARPS * 32, 4)
ax8_kernel(const double* __restrict__ u, const double* __restrict__ G6, double* __restrict__ w,
           const unsigned long long* __restrict__ maskbits, double* __restrict__ partials, int64_t e0,
           int64_t e1, const __grid_constant__ DMat8x Dp, const int* __restrict__ done, int l2_prefetch) {
  if (done != nullptr && *done) return;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* A = smem + warp * (2 * AX8_ARR);
  double* B = A + AX8_ARR;
  // No branch around the element body: a warp past the end redoes the last
  // element (same values to the same addresses) and contributes 0 to the dot.
  // Keeping the warp provably converged lets the compiler use the uniform
  // datapath (LDCU/UR) for D and plain WARPSYNCs.
  const int64_t e_raw = e0 + (int64_t)blockIdx.x * AX8_WARPS + warp;
  const bool live = e_raw < e1;
  const int64_t e = live ? e_raw : e1 - 1;
  double dot = 0.0;

  {
    // T layout
    const int ip = lane & 3;
    const int tj = ((lane >> 2) & 1) * 4 + (lane >> 3);
    const int t_g = tj * 8 + 2 * ip;            // offset inside a k-plane, global
    const int t_s = tj * AX8_RS + 2 * ip;       // offset inside a k-plane, shared
    const double* ue = u + e * 512;
    const double* ge = G6 + e * (6 * 512);

    // The register loads of G run one plane (3 KB per warp) ahead of their use: with 16 warps per SM
    // that is ~30 KB in flight per SM, short of what 6.5 TB/s x DRAM latency needs (measured: 79 % of
    // the HBM peak).  One lane therefore asks the L2 for the element's whole 24 KB block of G up
    // front (TMA-style bulk prefetch, no registers, no shared memory); the plane loads then hit L2.
    // Measured 206 -> 178 us per launch at 32^3 elements (92 % of the measured HBM peak).
    if (l2_prefetch && lane == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge), "r"(6 * 512 * 8) : "memory");

    // P1: u -> registers (T) and shared
    double2 uc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) uc[k] = __ldg(reinterpret_cast<const double2*>(ue + k * 64 + t_g));
    // first plane of G requested now so it overlaps the r/s passes
    double2 g[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) g[c] = ld_stream2(ge + c * 512 + t_g);
#pragma unroll
    for (int k = 0; k < 8; ++k) sts2(A + k * AX8_PS + t_s, uc[k]);
    __syncwarp();

    // line ownership (R and S layouts): lines L0 = lane, L1 = lane + 32
    const int lk0 = lane >> 3, lq = lane & 7;          // plane of line 0 (line 1: +4), index inside the plane
    double* colA = A + lk0 * AX8_PS + lq;               // S layout: (i = lq, k), stride AX8_RS over j
    double* colB = B + lk0 * AX8_PS + lq;
    double* rowA = A + lk0 * AX8_PS + lq * AX8_RS;      // R layout: (j = lq, k), contiguous over i
    constexpr int L1 = 4 * AX8_PS;                      // offset of the lane's second line

    // P2a: s-derivative, A -> B
    {
      double a[2][8], o[2][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        a[0][m] = colA[m * AX8_RS];
        a[1][m] = colA[L1 + m * AX8_RS];
      }
      line8x2<false>(Dp.d[0], a, o);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
    }
    __syncwarp();
    // P2b: r-derivative, in place on A (each line is private to its lane)
    {
      double a[2][8], o[2][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 v0 = lds2(rowA + 2 * c), v1 = lds2(rowA + L1 + 2 * c);
        a[0][2 * c] = v0.x;
        a[0][2 * c + 1] = v0.y;
        a[1][2 * c] = v1.x;
        a[1][2 * c + 1] = v1.y;
      }
      line8x2<false>(Dp.d[1], a, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts2(rowA + 2 * c, make_double2(o[0][2 * c], o[0][2 * c + 1]));
        sts2(rowA + L1 + 2 * c, make_double2(o[1][2 * c], o[1][2 * c + 1]));
      }
    }
    __syncwarp();

    // P3: t-derivative in registers, apply G, start D^T in t; wr -> A, ws -> B
    double2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double2 gn[6];
      if (k < 7) {
#pragma unroll
        for (int c = 0; c < 6; ++c) gn[c] = ld_stream2(ge + c * 512 + (k + 1) * 64 + t_g);
      }
      const double2 ur = lds2(A + k * AX8_PS + t_s);
      const double2 us = lds2(B + k * AX8_PS + t_s);
      double2 ut = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        ut.x = fma(Dp.d[2][k * 8 + m], uc[m].x, ut.x);
        ut.y = fma(Dp.d[2][k * 8 + m], uc[m].y, ut.y);
      }
      double2 wr, ws, wt;
      wr.x = g[0].x * ur.x + g[1].x * us.x + g[2].x * ut.x;
      wr.y = g[0].y * ur.y + g[1].y * us.y + g[2].y * ut.y;
      ws.x = g[1].x * ur.x + g[3].x * us.x + g[4].x * ut.x;
      ws.y = g[1].y * ur.y + g[3].y * us.y + g[4].y * ut.y;
      wt.x = g[2].x * ur.x + g[4].x * us.x + g[5].x * ut.x;
      wt.y = g[2].y * ur.y + g[4].y * us.y + g[5].y * ut.y;
      sts2(A + k * AX8_PS + t_s, wr);
      sts2(B + k * AX8_PS + t_s, ws);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        acc[kk].x = fma(Dp.d[2][k * 8 + kk], wt.x, acc[kk].x);
        acc[kk].y = fma(Dp.d[2][k * 8 + kk], wt.y, acc[kk].y);
      }
      if (k < 7) {
#pragma unroll
        for (int c = 0; c < 6; ++c) g[c] = gn[c];
      }
      __syncwarp();  // also keeps the compiler from hoisting every plane's G loads (register blow-up)
    }

    // P4: D^T along r (A, in place) and along s (B, in place)
    {
      double a[2][8], o[2][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 v0 = lds2(rowA + 2 * c), v1 = lds2(rowA + L1 + 2 * c);
        a[0][2 * c] = v0.x;
        a[0][2 * c + 1] = v0.y;
        a[1][2 * c] = v1.x;
        a[1][2 * c + 1] = v1.y;
      }
      line8x2<true>(Dp.d[3], a, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts2(rowA + 2 * c, make_double2(o[0][2 * c], o[0][2 * c + 1]));
        sts2(rowA + L1 + 2 * c, make_double2(o[1][2 * c], o[1][2 * c + 1]));
      }
    }
    {
      double a[2][8], o[2][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        a[0][m] = colB[m * AX8_RS];
        a[1][m] = colB[L1 + m * AX8_RS];
      }
      line8x2<true>(Dp.d[4], a, o);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
    }
    __syncwarp();

    // P5: sum the three parts, mask, local dot, store (T layout, coalesced 512 B per plane)
    double* we = w + e * 512;
    const int bit = tj * 8 + 2 * ip;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double2 r2 = lds2(A + k * AX8_PS + t_s);
      const double2 s2 = lds2(B + k * AX8_PS + t_s);
      double2 o;
      o.x = (r2.x + s2.x) + acc[k].x;
      o.y = (r2.y + s2.y) + acc[k].y;
      if (MASK) {
        const unsigned long long mb = __ldg(maskbits + e * 8 + k) >> bit;
        if (mb & 1ull) o.x = 0.0;
        if (mb & 2ull) o.y = 0.0;
      }
      if (DOT) dot = fma(o.x, uc[k].x, fma(o.y, uc[k].y, dot));
      *reinterpret_cast<double2*>(we + k * 64 + t_g) = o;
    }
  }
  if (DOT) block_sum_to<AX8_WARPS * 32>(live ? dot : 0.0, partials + blockIdx.x);
}

// ---------------------------------------------------------------------------
// Any Nq in [2, 16]: one CTA of Nq^2 threads per element, k-plane sweep with
// the u column and the D^T_t accumulators in registers.  Correct for every
// order; not tuned (the tuned path is ax8_kernel; Nq = 16 is next).
// ---------------------------------------------------------------------------
template <int NQ, bool MASK, bool DOT>
__global__ void __launch_bounds__(NQ* NQ)
ax_generic_kernel(const double* __restrict__ u, const double* __restrict__ G6, double* __restrict__ w,
                  const uint32_t* __restrict__ maskbits, double* __restrict__ partials, int64_t e0, int64_t e1,
                  const __grid_constant__ DMat Dp, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  constexpr int NP = NQ * NQ * NQ;
  constexpr int NT = NQ * NQ;
  __shared__ double sD[NQ][NQ + 1];
  __shared__ double su[NQ][NQ + 1];
  __shared__ double swr[NQ][NQ + 1];
  __shared__ double sws[NQ][NQ + 1];
  __shared__ double s_part[NT];
  const int i = threadIdx.x % NQ, j = threadIdx.x / NQ;
  const int64_t e = e0 + blockIdx.x;
  sD[j][i] = Dp.d[j * NQ + i];
  const double* ue = u + e * NP;
  const double* ge = G6 + e * (int64_t)(6 * NP);
  double uc[NQ], acc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    uc[k] = ue[k * NT + j * NQ + i];
    acc[k] = 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    su[j][i] = uc[k];
    __syncthreads();
    double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
    for (int m = 0; m < NQ; ++m) {
      ur = fma(sD[i][m], su[j][m], ur);
      us = fma(sD[j][m], su[m][i], us);
      ut = fma(sD[k][m], uc[m], ut);
    }
    const int node = k * NT + j * NQ + i;
    const double grr = ge[node], grs = ge[NP + node], grt = ge[2 * NP + node];
    const double gss = ge[3 * NP + node], gst = ge[4 * NP + node], gtt = ge[5 * NP + node];
    swr[j][i] = grr * ur + grs * us + grt * ut;
    sws[j][i] = grs * ur + gss * us + gst * ut;
    const double wt = grt * ur + gst * us + gtt * ut;
#pragma unroll
    for (int kk = 0; kk < NQ; ++kk) acc[kk] = fma(sD[k][kk], wt, acc[kk]);
    __syncthreads();
    double ar = 0.0, as = 0.0;
#pragma unroll
    for (int m = 0; m < NQ; ++m) {
      ar = fma(sD[m][i], swr[j][m], ar);
      as = fma(sD[m][j], sws[m][i], as);
    }
    acc[k] += ar + as;
    __syncthreads();
  }
  double dot = 0.0;
  double* we = w + e * NP;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const int node = k * NT + j * NQ + i;
    double o = acc[k];
    if (MASK) {
      const int64_t gi = e * NP + node;
      if ((maskbits[gi >> 5] >> (gi & 31)) & 1u) o = 0.0;
    }
    if (DOT) dot = fma(o, uc[k], dot);
    we[node] = o;
  }
  if (DOT) {
    // partial warps exist for most Nq, so no shuffles here: fixed-order sum through shared memory
    s_part[threadIdx.x] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int q = 0; q < NT; ++q) s += s_part[q];
      partials[blockIdx.x] = s;
    }
  }
}

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
