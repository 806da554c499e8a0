// K1 for every order but N = 7 (Nq = 3 ... 7 and 9 ... 16): the same three-layout, contract-in-registers
// design as ax8_kernel, with Nq^2 threads per element and one line per thread and layout.  One element per
// CTA from Nq = 9 up; below that EPB elements share a CTA (Nq = 4: 10 x 16 threads) so that the warps are full.
//
//   T: thread (i, j) owns the k-column           -> t-derivative / D^T_t in registers, G applied here
//   R: thread (j, k) owns the r-line (Nq contiguous values)
//   S: thread (i, k) owns the s-line (stride = one padded row)
// The element lives in shared memory as THREE staged arrays of Nq planes (110 KB at Nq = 16, two CTAs =
// 16 warps per SM -- the "shared-memory pressure" of BASELINE config 3):
//   P1  u -> registers (T columns) and A   (+ the folded direction update); requests enter the memory system in the
//       order they are needed: the loads of P1, then the bulk L2 prefetch of G (AX_PF_POS below)
//   P2  r-line of A -> C (u_r) and s-line of A -> B (u_s) in ONE pass: both lines use the same entry of D
//       for the same output index, so every D operand fetched (LDCU -> uniform register) feeds two DFMAs
//   P3  u_t = D_t u in registers; per plane: G, (w_r, w_s, w_t) -> (A, B, C) at the thread's own T location;
//       p.Ap accumulated as (grad u)^T G (grad u) = u_r w_r + u_s w_s + u_t w_t (no u needed later)
//   P4  D^T along r (A, in place) and along s (B, in place) in one pass sharing D again; D^T along t from
//       C into registers
//   P5  sum, mask, store
// Compared with accumulating D^T_t inside P3 and keeping u for the dot product (round 1: 128 registers
// and 240 ... 1048 B of spills per thread with the CG variants, 6 barriers), nothing is live across a
// phase but the u column in P1-P3 and the t-result in P4-P5: no spills, 4 barriers.
// Row padding makes the R-line accesses conflict-free (RS/2 odd for even Nq -> LDS.128; RS odd for odd
// Nq -> LDS.64); S and T accesses are contiguous across lanes.  D is replicated per phase in the
// parameter block (8 KB at Nq = 16; see line8x2 in k1_ax8.cu for why) and read through uniform registers.
#pragma once

#include <cstring>

#include "ax_common.cuh"

#ifndef AXL_GPD
#define AXL_GPD 0  // planes of G in flight ahead of the plane being used; 0 = per-order default (build knob)
#endif
// Where the bulk L2 prefetch of the element's G is issued (build knob; -1 = the measured best):
//   0 first thing (rounds 1-2), 1 behind the loads of P1, 2 after P1, 3 first half of every component behind the
//   loads of P1 and the second half after the first barrier.
// The loads of P1 are what the CTA waits for first; queued behind 196 KB of prefetch (Nq = 16) they arrive late.
// Plain operator, Nq = 16, 16.8 M nodes: 230 (0) -> 218 (1) / 220 (3) us; PCG iteration 479 -> 456 (1) / 448 (3) us.
#ifndef AX_PF_POS
#define AX_PF_POS -1
#endif

namespace semb {

template <int NQ>
struct DMatL {
  double d[4][NQ * NQ];
};

template <int NQ>
struct AxlCfg {
  static constexpr bool V2 = (NQ % 2 == 0);
  static constexpr int RS = V2 ? ((NQ % 4 == 2) ? NQ : NQ + 2) : NQ;
  static constexpr int PS = NQ * RS;
  static constexpr int ARR = NQ * PS;
  static constexpr int EPB = (160 / (NQ * NQ)) > 0 ? (160 / (NQ * NQ)) : 1;  // elements per CTA
  static constexpr int SMEM = EPB * 3 * ARR * (int)sizeof(double);
  static constexpr int NT = EPB * NQ * NQ;
  // registers and shared memory set the residency: <= 512 threads of 128 registers, <= 227 KB per SM
  static constexpr int BY_REGS = (512 / NT) > 0 ? (512 / NT) : 1;
  static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 1024) > 0 ? (227 * 1024) / (SMEM + 1024) : 1;
  static constexpr int MINB = BY_REGS < BY_SMEM ? BY_REGS : BY_SMEM;
  // planes of G requested ahead of the one in use (ring of GPD + 1 register sets).  Measured, plain operator
  // at ~16 M nodes: Nq = 11: 205 / 198 / 191 / 188 us for 1 / 2 / 3 / 4 planes (4 planes: spills in the CG
  // variant), Nq = 13: 217 / 205 / 207 / 213, Nq = 16: 231 / 227 / 227 / 227; the low orders are untested: 1.
  static constexpr int GPD_DEFAULT = NQ >= 13 ? 2 : (NQ >= 9 ? 3 : 1);
  static constexpr int GPD_WANT = AXL_GPD > 0 ? AXL_GPD : GPD_DEFAULT;
  static constexpr int GPD = GPD_WANT < NQ - 1 ? GPD_WANT : NQ - 1;
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

// One pass over an r-line (contiguous, `row_in` -> `row_out`) and an s-line (stride RS, `col_in` -> `col_out`)
// with the same D: out[o] = sum_m D[o][m] in[m] (or D[m][o] with TRANSPOSE).  Results are stored as soon as
// they are complete (two outputs at a time where the row store can be 128 bits wide).
template <int NQ, bool TRANSPOSE>
__device__ __forceinline__ void line_pair(const double (&D)[NQ * NQ], const double* row_in, double* row_out,
                                          const double* col_in, double* col_out) {
  using C = AxlCfg<NQ>;
  constexpr int RS = C::RS;
  double ar[NQ], as[NQ];
  if (C::V2) {
#pragma unroll
    for (int c = 0; c < NQ / 2; ++c) {
      const double2 v = lds2(row_in + 2 * c);
      ar[2 * c] = v.x;
      ar[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < NQ; ++c) ar[c] = row_in[c];
  }
#pragma unroll
  for (int m = 0; m < NQ; ++m) as[m] = col_in[m * RS];
  constexpr int STEP = C::V2 ? 2 : 1;
#pragma unroll
  for (int o = 0; o < NQ; o += STEP) {
    double sr[STEP], ss[STEP];
#pragma unroll
    for (int q = 0; q < STEP; ++q) sr[q] = ss[q] = 0.0;
#pragma unroll
    for (int m = 0; m < NQ; ++m) {
#pragma unroll
      for (int q = 0; q < STEP; ++q) {
        const double d = TRANSPOSE ? D[m * NQ + o + q] : D[(o + q) * NQ + m];
        sr[q] = fma(d, ar[m], sr[q]);
        ss[q] = fma(d, as[m], ss[q]);
      }
    }
    if (C::V2) sts2(row_out + o, make_double2(sr[0], sr[STEP - 1]));
    else row_out[o] = sr[0];
#pragma unroll
    for (int q = 0; q < STEP; ++q) col_out[(o + q) * RS] = ss[q];
  }
}

template <int NQ, bool MASK, bool DOT, int DIR>
__global__ void __launch_bounds__(AxlCfg<NQ>::NT, AxlCfg<NQ>::MINB)
axl_kernel(const __grid_constant__ AxArgs a, const __grid_constant__ DMatL<NQ> Dp) {
  using C = AxlCfg<NQ>;
  constexpr int NT = C::NT, NL = NQ * NQ, NP = NQ * NQ * NQ, RS = C::RS, PS = C::PS;
  pdl_launch_dependents();
  pdl_wait();
  if (a.done != nullptr && *a.done) return;
  extern __shared__ __align__(16) double smem[];
  const int el = C::EPB > 1 ? (int)threadIdx.x / NL : 0;      // element slot of this thread inside the CTA
  const int tid = C::EPB > 1 ? (int)threadIdx.x - el * NL : (int)threadIdx.x;
  double* A = smem + el * (3 * C::ARR);
  double* B = A + C::ARR;
  double* Cc = A + 2 * C::ARR;
  // a slot past the end redoes the last element (all threads take part in the barriers) and stores nothing
  const int64_t e_raw = a.e0 + (int64_t)blockIdx.x * C::EPB + el;
  const bool live = e_raw < a.e1;
  const int64_t e = live ? e_raw : a.e1 - 1;
  const double* ue = a.u + e * NP;
  const double* ge = a.G6 + e * (int64_t)(6 * NP);

  // bulk L2 prefetch of the element's block of G (see ax8_kernel)
  // half = 0: the whole block; 1 / 2: first / second half of every component (position 3)
  auto prefetch_g = [&](int half = 0) {
    if (a.l2_prefetch) {
      if ((NP % 4) == 0) {
        const int bytes = half == 0 ? NP * 8 : NP * 4;
        const double* src = ge + tid * NP + (half == 2 ? NP / 2 : 0);
        if (tid < 6) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
      } else if ((NP % 2) == 0) {
        if (tid < 6 && half != 2) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge + tid * NP), "r"(NP * 8) : "memory");
      } else if (tid == 0 && half != 2) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge), "r"(6 * NP * 8) : "memory");
      }
    }
  };
  constexpr int PF = AX_PF_POS >= 0 ? AX_PF_POS : (DIR != 0 ? 3 : 1);
  if (PF == 0) prefetch_g();

  // T layout: column (ti, tj)
  const int ti = tid % NQ, tj = tid / NQ;
  const int t_g = tj * NQ + ti;   // inside a k-plane, global
  const int t_s = tj * RS + ti;   // inside a k-plane, shared
  // R layout: line (j = ti, k = tj) ; S layout: line (i = ti, k = tj)
  const int row = tj * PS + ti * RS;
  const int col = tj * PS + ti;

  // P1: u -> registers (T) and shared; first plane of G
  double uc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) uc[k] = __ldg(ue + k * NL + t_g);
  if (DIR != 0) {
    // CG direction update folded into the load phase: p <- z + beta p, written back and used as u
    const double beta = a.s->beta;
    const double* re = a.r + e * NP;
    const double* de = a.dinv + e * NP;
    double* pe = a.p_out + e * NP;
    double rv[NQ], dv[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      rv[k] = __ldg(re + k * NL + t_g);
      dv[k] = DIR == 2 ? __ldg(de + k * NL + t_g) : 0.0;
    }
    if (PF == 1) prefetch_g();
    if (PF == 3) prefetch_g(1);
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      uc[k] = dir_update<DIR>(uc[k], rv[k], dv[k], beta);
      if (live) st_out(pe + k * NL + t_g, uc[k]);
    }
  } else if (PF == 1) {
    prefetch_g();
  } else if (PF == 3) {
    prefetch_g(1);
  }
  if (PF == 2) prefetch_g();
  // G runs GPD planes ahead of its use in a ring of GPD + 1 register sets (plane 0 is requested here,
  // planes 1 .. GPD-1 behind P2 where the registers are free again)
  constexpr int GPD = AxlCfg<NQ>::GPD, GS = GPD + 1;
  double g[GS][6];
#pragma unroll
  for (int c = 0; c < 6; ++c) g[0][c] = __ldcs(ge + c * NP + t_g);
#pragma unroll
  for (int k = 0; k < NQ; ++k) A[k * PS + t_s] = uc[k];
  __syncthreads();
  if (PF == 3) prefetch_g(2);

  // P2: u_r (A rows -> C rows) and u_s (A columns -> B columns), one shared sweep over D
  line_pair<NQ, false>(Dp.d[0], A + row, Cc + row, A + col, B + col);
#pragma unroll
  for (int q = 1; q < GPD; ++q) {
#pragma unroll
    for (int c = 0; c < 6; ++c) g[q][c] = __ldcs(ge + c * NP + q * NL + t_g);
  }
  __syncthreads();

  // P3: u_t in registers, G, (w_r, w_s, w_t) -> (A, B, C) at the thread's own location; p.Ap
  double dot = 0.0;
  {
    double ut[NQ];
    line_contract<NQ, false>(Dp.d[1], uc, ut);
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      if (k + GPD < NQ) {
#pragma unroll
        for (int c = 0; c < 6; ++c) g[(k + GPD) % GS][c] = __ldcs(ge + c * NP + (k + GPD) * NL + t_g);
      }
      const double(&gk)[6] = g[k % GS];
      const double ur = Cc[k * PS + t_s];
      const double us = B[k * PS + t_s];
      const double wr = gk[0] * ur + gk[1] * us + gk[2] * ut[k];
      const double ws = gk[1] * ur + gk[3] * us + gk[4] * ut[k];
      const double wt = gk[2] * ur + gk[4] * us + gk[5] * ut[k];
      if (DOT) dot = fma(ur, wr, fma(us, ws, fma(ut[k], wt, dot)));
      A[k * PS + t_s] = wr;
      B[k * PS + t_s] = ws;
      Cc[k * PS + t_s] = wt;
      __syncwarp();  // scheduling fence: keeps the compiler from hoisting every plane's G loads
    }
  }
  __syncthreads();

  // P4: D^T along r (A rows, in place) and along s (B columns, in place) sharing D; D^T along t (C) -> registers
  line_pair<NQ, true>(Dp.d[2], A + row, A + row, B + col, B + col);
  double acc[NQ];
  {
    double at[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) at[k] = Cc[k * PS + t_s];
    line_contract<NQ, true>(Dp.d[3], at, acc);
  }
  __syncthreads();

  // P5: sum the three parts, mask, store.  The mask words of all planes are fetched first: behind the stores
  // to w the compiler cannot move them (possible aliasing) and every plane would expose a load latency
  // (ncu: 22 % of the stall samples sat in this loop).
  double* we = a.w + e * NP;
  uint32_t mw[NQ];
  if (MASK) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) mw[k] = __ldg(a.maskbits + ((e * NP + k * NL + t_g) >> 5));
  }
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    double o = (A[k * PS + t_s] + B[k * PS + t_s]) + acc[k];
    if (MASK) {
      const int64_t gi = e * NP + k * NL + t_g;
      if ((mw[k] >> (gi & 31)) & 1u) {
        // the reference's dot is p . mask(Ap) (sempy/elliptic.py:49-51): take the masked rows out of
        // (grad u)^T G (grad u) = u . Au.  Rare path (boundary nodes only); u comes back from memory.
        if (DOT) dot = fma(-o, (DIR != 0 ? a.p_out : a.u)[gi], dot);
        o = 0.0;
      }
    }
    if (live) st_out(we + k * NL + t_g, o);
  }
  if (DOT) {
    const double cta = block_sum_nt<NT>(live ? dot : 0.0);
    // one partial per CTA; summed in index order by a passenger CTA of the launch that follows
    if (threadIdx.x == 0) a.red.part[a.cta_off + (int)blockIdx.x] = cta;
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
  SEMB_CUDA(launch_maybe_pdl(axl_kernel<NQ, MASK, DOT, DIR>, dim3((unsigned)((a.e1 - a.e0 + Cfg::EPB - 1) / Cfg::EPB)), dim3(Cfg::NT),
                             Cfg::SMEM, st, a.pdl != 0, a, dl));
  return 0;
}

template <int NQ>
static int launch_axl_nq(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  DMatL<NQ> dl;
  for (int q = 0; q < 4; ++q)
    for (int r = 0; r < NQ; ++r) std::memcpy(&dl.d[q][r * NQ], &D.d[r * NQ], sizeof(double) * NQ);
  SEMB_CHECK((dot && (dir == 1 || dir == 2)) || (!dot && dir == 0), "launch_axl: unsupported variant");
  if (!dot) return mask ? launch_axl_t<NQ, true, false, 0>(a, dl, st, device) : launch_axl_t<NQ, false, false, 0>(a, dl, st, device);
  if (dir == 1) return mask ? launch_axl_t<NQ, true, true, 1>(a, dl, st, device) : launch_axl_t<NQ, false, true, 1>(a, dl, st, device);
  return mask ? launch_axl_t<NQ, true, true, 2>(a, dl, st, device) : launch_axl_t<NQ, false, true, 2>(a, dl, st, device);
}

}  // namespace semb
