// K1 for the higher orders (Nq = 9 ... 16, i.e. N = 8 ... 15): the same three-layout, contract-in-registers
// design as ax8_kernel, with one CTA of Nq^2 threads per element and one line per thread.
//
//   T: thread (i, j) owns the k-column           -> t-derivative and D^T_t in registers, G applied here
//   R: thread (j, k) owns the r-line (Nq contiguous values)
//   S: thread (i, k) owns the s-line (stride = one padded row)
// The element (two staged arrays of Nq planes) lives in shared memory: 74 KB at Nq = 16, so two CTAs
// (16 warps) per SM -- the "shared-memory pressure" of BASELINE config 3.  Row padding makes the
// R-line accesses conflict-free (RS/2 odd for even Nq -> LDS.128; RS odd for odd Nq -> LDS.64); S and
// T accesses are contiguous across lanes.  D is replicated per phase in the parameter block (10 KB at
// Nq = 16; see line8x2 in ax_kernels.cuh for why) and read through uniform registers.
#pragma once

#include <cstring>

#include "ax_common.cuh"

namespace semb {

template <int NQ>
struct DMatL {
  double d[5][NQ * NQ];
};

template <int NQ>
struct AxlCfg {
  static constexpr bool V2 = (NQ % 2 == 0);
  static constexpr int RS = V2 ? ((NQ % 4 == 2) ? NQ : NQ + 2) : NQ;
  static constexpr int PS = NQ * RS;
  static constexpr int ARR = NQ * PS;
  static constexpr int SMEM = 2 * ARR * (int)sizeof(double);
  static constexpr int NT = NQ * NQ;
  // registers, not shared memory, set the residency: 128 registers per thread need <= 512 threads per SM
  static constexpr int MINB = (512 / NT) > 0 ? (512 / NT) : 1;
};

template <int NQ, bool TRANSPOSE>
__device__ __forceinline__ void line_contract(const double (&D)[NQ * NQ], const double (&a)[NQ], double (&o)[NQ]) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NQ; ++m) s = fma(TRANSPOSE ? D[m * NQ + i] : D[i * NQ + m], a[m], s);
    o[i] = s;
  }
}

template <int NQ, bool MASK, bool DOT, int DIR>
__global__ void __launch_bounds__(AxlCfg<NQ>::NT, AxlCfg<NQ>::MINB)
axl_kernel(const __grid_constant__ AxArgs a, const __grid_constant__ DMatL<NQ> Dp) {
  using C = AxlCfg<NQ>;
  constexpr int NT = C::NT, NP = NQ * NQ * NQ, RS = C::RS, PS = C::PS;
  if (a.done != nullptr && *a.done) return;
  extern __shared__ __align__(16) double smem[];
  double* A = smem;
  double* B = smem + C::ARR;
  const int tid = threadIdx.x;
  const int64_t e = a.e0 + blockIdx.x;
  const double* ue = a.u + e * NP;
  const double* ge = a.G6 + e * (int64_t)(6 * NP);
  const int l2_prefetch = a.l2_prefetch;

  // bulk L2 prefetch of the element's block of G (see ax8_kernel)
  if (l2_prefetch) {
    if ((NP % 2) == 0) {
      if (tid < 6) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge + tid * NP), "r"(NP * 8) : "memory");
    } else if (tid == 0) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge), "r"(6 * NP * 8) : "memory");
    }
  }

  // T layout: column (ti, tj)
  const int ti = tid % NQ, tj = tid / NQ;
  const int t_g = tj * NQ + ti;   // inside a k-plane, global
  const int t_s = tj * RS + ti;   // inside a k-plane, shared
  // R layout: line (j = ti, k = tj) ; S layout: line (i = ti, k = tj)
  double* rowA = A + tj * PS + ti * RS;
  double* colA = A + tj * PS + ti;
  double* colB = B + tj * PS + ti;

  // P1: u -> registers (T) and shared; first plane of G
  double uc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) uc[k] = __ldg(ue + k * NT + t_g);
  if (DIR != 0) {
    // CG direction update folded into the load phase: p <- z + beta p, written back and used as u
    const double beta = a.s->beta;
    const double* re = a.r + e * NP;
    const double* de = a.dinv + e * NP;
    double* pe = a.p_out + e * NP;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const double rv = __ldg(re + k * NT + t_g);
      const double dv = DIR == 2 ? __ldg(de + k * NT + t_g) : 0.0;
      uc[k] = dir_update<DIR>(uc[k], rv, dv, beta);
      pe[k * NT + t_g] = uc[k];
    }
  }
  double g[6], gn[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) g[c] = __ldcs(ge + c * NP + t_g);
#pragma unroll
  for (int k = 0; k < NQ; ++k) A[k * PS + t_s] = uc[k];
  __syncthreads();

  // P2a: s-derivative, A -> B
  {
    double a[NQ], o[NQ];
#pragma unroll
    for (int m = 0; m < NQ; ++m) a[m] = colA[m * RS];
    line_contract<NQ, false>(Dp.d[0], a, o);
#pragma unroll
    for (int m = 0; m < NQ; ++m) colB[m * RS] = o[m];
  }
  __syncthreads();
  // P2b: r-derivative, in place on A (each line is private to its thread)
  {
    double a[NQ], o[NQ];
    if (C::V2) {
#pragma unroll
      for (int c = 0; c < NQ / 2; ++c) {
        const double2 v = lds2(rowA + 2 * c);
        a[2 * c] = v.x;
        a[2 * c + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < NQ; ++c) a[c] = rowA[c];
    }
    line_contract<NQ, false>(Dp.d[1], a, o);
    if (C::V2) {
#pragma unroll
      for (int c = 0; c < NQ / 2; ++c) sts2(rowA + 2 * c, make_double2(o[2 * c], o[2 * c + 1]));
    } else {
#pragma unroll
      for (int c = 0; c < NQ; ++c) rowA[c] = o[c];
    }
  }
  __syncthreads();

  // P3: t-derivative in registers, G, start of D^T_t; wr -> A, ws -> B (own locations only)
  double acc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) acc[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    if (k < NQ - 1) {
#pragma unroll
      for (int c = 0; c < 6; ++c) gn[c] = __ldcs(ge + c * NP + (k + 1) * NT + t_g);
    }
    const double ur = A[k * PS + t_s];
    const double us = B[k * PS + t_s];
    double ut = 0.0;
#pragma unroll
    for (int m = 0; m < NQ; ++m) ut = fma(Dp.d[2][k * NQ + m], uc[m], ut);
    const double wr = g[0] * ur + g[1] * us + g[2] * ut;
    const double ws = g[1] * ur + g[3] * us + g[4] * ut;
    const double wt = g[2] * ur + g[4] * us + g[5] * ut;
    A[k * PS + t_s] = wr;
    B[k * PS + t_s] = ws;
#pragma unroll
    for (int kk = 0; kk < NQ; ++kk) acc[kk] = fma(Dp.d[2][k * NQ + kk], wt, acc[kk]);
    if (k < NQ - 1) {
#pragma unroll
      for (int c = 0; c < 6; ++c) g[c] = gn[c];
    }
    __syncwarp();  // scheduling fence: keeps the compiler from hoisting every plane's G loads
  }
  __syncthreads();

  // P4: D^T along r (A, in place) and along s (B, in place)
  {
    double a[NQ], o[NQ];
    if (C::V2) {
#pragma unroll
      for (int c = 0; c < NQ / 2; ++c) {
        const double2 v = lds2(rowA + 2 * c);
        a[2 * c] = v.x;
        a[2 * c + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < NQ; ++c) a[c] = rowA[c];
    }
    line_contract<NQ, true>(Dp.d[3], a, o);
    if (C::V2) {
#pragma unroll
      for (int c = 0; c < NQ / 2; ++c) sts2(rowA + 2 * c, make_double2(o[2 * c], o[2 * c + 1]));
    } else {
#pragma unroll
      for (int c = 0; c < NQ; ++c) rowA[c] = o[c];
    }
  }
  {
    double a[NQ], o[NQ];
#pragma unroll
    for (int m = 0; m < NQ; ++m) a[m] = colB[m * RS];
    line_contract<NQ, true>(Dp.d[4], a, o);
#pragma unroll
    for (int m = 0; m < NQ; ++m) colB[m * RS] = o[m];
  }
  __syncthreads();

  // P5: sum the three parts, mask, local dot, store
  double dot = 0.0;
  double* we = a.w + e * NP;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    double o = (A[k * PS + t_s] + B[k * PS + t_s]) + acc[k];
    if (MASK) {
      const int64_t gi = e * NP + k * NT + t_g;
      if ((__ldg(a.maskbits + (gi >> 5)) >> (gi & 31)) & 1u) o = 0.0;
    }
    if (DOT) dot = fma(o, uc[k], dot);
    we[k * NT + t_g] = o;
  }
  if (DOT) {
    const double cta = block_sum_nt<NT>(dot);
    double total;
    if (grid_reduce<NT>(a.red, cta, a.cta_off + (int)blockIdx.x, a.ncta_total, total) && tid == 0) a.s->pAp = total;
  }
}

template <int NQ, bool MASK, bool DOT, int DIR>
static int launch_axl_t(const AxArgs& a, const DMatL<NQ>& dl, cudaStream_t st, int device) {
  using Cfg = AxlCfg<NQ>;
  static bool attr_done[64] = {false};
  if (!attr_done[device & 63]) {
    SEMB_CUDA(cudaFuncSetAttribute(axl_kernel<NQ, MASK, DOT, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    SEMB_CUDA(cudaFuncSetAttribute(axl_kernel<NQ, MASK, DOT, DIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[device & 63] = true;
  }
  axl_kernel<NQ, MASK, DOT, DIR><<<(unsigned)(a.e1 - a.e0), Cfg::NT, Cfg::SMEM, st>>>(a, dl);
  SEMB_CUDA(cudaGetLastError());
  return 0;
}

template <int NQ>
static int launch_axl_nq(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  DMatL<NQ> dl;
  for (int q = 0; q < 5; ++q)
    for (int r = 0; r < NQ; ++r) std::memcpy(&dl.d[q][r * NQ], &D.d[r * NQ], sizeof(double) * NQ);
  SEMB_CHECK((dot && (dir == 1 || dir == 2)) || (!dot && dir == 0), "launch_axl: unsupported variant");
  if (!dot) return mask ? launch_axl_t<NQ, true, false, 0>(a, dl, st, device) : launch_axl_t<NQ, false, false, 0>(a, dl, st, device);
  if (dir == 1) return mask ? launch_axl_t<NQ, true, true, 1>(a, dl, st, device) : launch_axl_t<NQ, false, true, 1>(a, dl, st, device);
  return mask ? launch_axl_t<NQ, true, true, 2>(a, dl, st, device) : launch_axl_t<NQ, false, true, 2>(a, dl, st, device);
}

}  // namespace semb
