// Fast-diagonalisation (FDM) preconditioner, the multi-element form of the reference's single-element
// sketch in example.py:71-114 (1-D generalised eigenproblems A s = lambda B s), :155-171 (fast_kron) and
// :186-190 (precon_fdm: R, S^T (x) S^T (x) S^T, 1/diag, S (x) S (x) S, R^T).
//
// z = M^-1 r with M block diagonal:
//   * one block per element INTERIOR (the nodes 1 .. Nq-2 in each direction; they belong to one element
//     only, so the assembled operator's block is the element's own): the separable operator of the
//     element's bounding box  B(x)B(x)A + B(x)A(x)B + A(x)B(x)B  with homogeneous Dirichlet data on the
//     element surface, inverted exactly as
//         (S(x)S(x)S) diag( scale / (ax l_a + ay l_b + az l_c) ) (S(x)S(x)S)^T
//     where (S, l) solve the interior 1-D problem  A_hat S = B_hat S diag(l),  S^T B_hat S = I,  and
//     ax = (2/Lx)^2, ..., scale = 8 / (Lx Ly Lz) carry the element's edge lengths.  Exact for straight
//     box elements, a separable approximation for deformed ones;
//   * point Jacobi (the assembled, masked diagonal) on the element surfaces.
// M is symmetric positive definite, needs no overlap and no exchange between elements or GPUs.
// The kernel also forms the partial of sum rmult r z (sempy/iterative.py:59,73) and, in the CTA that
// finishes last, the grid total, the all-rank sum and the CG scalar step -- the same tail as the update
// kernel.
#pragma once

#include "common.cuh"
#include "reduce.cuh"
#include "vec_kernels.cuh"

namespace semb {

constexpr int FDM_NT = 128;

struct FdmArgs {
  const double* r;
  double* z;
  const double* dinv;          // Jacobi part (surface nodes)
  const unsigned char* mult;   // copy counts (weights of the dot)
  const double* S;             // [m*m] row-major S[node][mode], then [m] eigenvalues
  const double* esc;           // (E,4): ax, ay, az, scale
  CgScalars* s;
  double* hist;
  GridRed red;
  int ncta;
  int finish;                  // 0: only s->rdotr_new = sum; 1: + all-rank sum and CG scalar step
  int check_done;              // skip when s->done (inside the CG loop)
};

template <int NQ>
__global__ void __launch_bounds__(FDM_NT)
fdm_apply_kernel(const FdmArgs a, const LlArgs ll) {
  if (a.check_done && a.s->done) return;
  constexpr int M = NQ - 2 > 0 ? NQ - 2 : 0;
  constexpr int NP = NQ * NQ * NQ, MP = M * M * M;
  constexpr int MD = M > 0 ? M : 1;  // divisor (the M = 0 instantiation never executes the divisions)
  __shared__ double sA[MP > 0 ? MP : 1];
  __shared__ double sB[MP > 0 ? MP : 1];
  __shared__ double sS[M * M > 0 ? M * M : 1];
  __shared__ double sL[M > 0 ? M : 1];
  const int tid = threadIdx.x;
  const int64_t e = blockIdx.x;
  const double* re = a.r + e * NP;
  double* ze = a.z + e * NP;
  for (int q = tid; q < M * M; q += FDM_NT) sS[q] = a.S[q];
  for (int q = tid; q < M; q += FDM_NT) sL[q] = a.S[M * M + q];
  double dot = 0.0;
  // surface: point Jacobi; interior: stage r
  for (int q = tid; q < NP; q += FDM_NT) {
    const int i = q % NQ, j = (q / NQ) % NQ, k = q / (NQ * NQ);
    const bool inner = i > 0 && i < NQ - 1 && j > 0 && j < NQ - 1 && k > 0 && k < NQ - 1;
    const double rv = re[q];
    if (inner) {
      sA[((k - 1) * M + (j - 1)) * M + (i - 1)] = rv;
    } else {
      const double zv = a.dinv[e * NP + q] * rv;
      ze[q] = zv;
      dot = fma(rv / (double)a.mult[e * NP + q], zv, dot);
    }
  }
  __syncthreads();
  if (MP > 0) {
    const double ax = a.esc[4 * e + 0], ay = a.esc[4 * e + 1], az = a.esc[4 * e + 2], sc = a.esc[4 * e + 3];
    // forward: (S^T (x) S^T (x) S^T) r
    for (int q = tid; q < MP; q += FDM_NT) {  // x: sB[k][j][a] = sum_i S[i][a] sA[k][j][i]
      const int m = q % MD, kj = q / MD;
      double v = 0.0;
#pragma unroll
      for (int i = 0; i < M; ++i) v = fma(sS[i * M + m], sA[kj * M + i], v);
      sB[q] = v;
    }
    __syncthreads();
    for (int q = tid; q < MP; q += FDM_NT) {  // y: sA[k][b][a] = sum_j S[j][b] sB[k][j][a]
      const int m = q % MD, b = (q / MD) % MD, k = q / (MD * MD);
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j) v = fma(sS[j * M + b], sB[(k * M + j) * M + m], v);
      sA[q] = v;
    }
    __syncthreads();
    for (int q = tid; q < MP; q += FDM_NT) {  // z and the diagonal: sB[c][b][a]
      const int m = q % MD, b = (q / MD) % MD, cidx = q / (MD * MD);
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) v = fma(sS[k * M + cidx], sA[(k * M + b) * M + m], v);
      sB[q] = v * (sc / (ax * sL[m] + ay * sL[b] + az * sL[cidx]));
    }
    __syncthreads();
    // back: (S (x) S (x) S)
    for (int q = tid; q < MP; q += FDM_NT) {  // x: sA[c][b][i] = sum_a S[i][a] sB[c][b][a]
      const int i = q % MD, cb = q / MD;
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < M; ++m) v = fma(sS[i * M + m], sB[cb * M + m], v);
      sA[q] = v;
    }
    __syncthreads();
    for (int q = tid; q < MP; q += FDM_NT) {  // y: sB[c][j][i] = sum_b S[j][b] sA[c][b][i]
      const int i = q % MD, j = (q / MD) % MD, cidx = q / (MD * MD);
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < M; ++b) v = fma(sS[j * M + b], sA[(cidx * M + b) * M + i], v);
      sB[q] = v;
    }
    __syncthreads();
    for (int q = tid; q < MP; q += FDM_NT) {  // z: out[k][j][i] = sum_c S[k][c] sB[c][j][i]
      const int i = q % MD, j = (q / MD) % MD, k = q / (MD * MD);
      double v = 0.0;
#pragma unroll
      for (int cidx = 0; cidx < M; ++cidx) v = fma(sS[k * M + cidx], sB[(cidx * M + j) * M + i], v);
      const int node = ((k + 1) * NQ + (j + 1)) * NQ + (i + 1);
      ze[node] = v;
      dot = fma(re[node], v, dot);  // interior nodes have one copy: weight 1
    }
  }
  const double cta = block_sum_nt<FDM_NT>(dot);
  double total;
  if (!grid_reduce<FDM_NT>(a.red, cta, (int)blockIdx.x, a.ncta, total)) return;
  if (tid == 0) a.s->rdotr_new = total;
  if (!a.finish) return;
  if (ll.nranks > 1) {
    __syncthreads();
    ll_allreduce_cta(&a.s->rdotr_new, 1, ll.peers, ll.rank, ll.nranks, ll.seq, ll.ll_error);
  }
  if (tid == 0) cg_step_scalars(a.s, a.hist);
}

// ---------------------------------------------------------------------------
// Tuned form (interior size M = Nq - 2 >= 2): the three-layout, contract-in-registers scheme of the operator
// kernels.  Thread (a, b) of an element owns, in turn, the x-line (j = b, k = a), the y-line (i = b, k = a)
// and the z-column (i = b, j = a) of the M^3 interior block, which lives in ONE padded shared array: every
// 1-D pass reads a whole line into registers, multiplies by S^T or S with the entries of S coming from the
// kernel parameter block (uniform registers) and writes the line back in place; forward-z, the eigenvalue
// scaling and backward-z happen in registers without touching shared memory.  EPB elements share a CTA so
// that small orders still fill their warps (N = 7: 4 elements x 36 lines).
// ---------------------------------------------------------------------------
template <int M>
struct FdmS {
  double s[6][M * M];  // one private copy of S per use (see line8x2 in k1_ax8.cu for why)
  double lam[M];
};

template <int M, bool TRANSPOSE>
__device__ __forceinline__ void fdm_line(const double (&S)[M * M], const double (&in)[M], double (&out)[M]) {
#pragma unroll
  for (int o = 0; o < M; ++o) {
    double v = 0.0;
#pragma unroll
    for (int m = 0; m < M; ++m) v = fma(TRANSPOSE ? S[m * M + o] : S[o * M + m], in[m], v);
    out[o] = v;
  }
}

template <int NQ>
struct FdmCfg {
  static constexpr int M = NQ - 2;
  static constexpr int L = M * M;                                   // lines per direction
  static constexpr int EPB = (160 / L) > 0 ? (160 / L) : 1;         // elements per CTA
  static constexpr int NT = EPB * L;
  static constexpr int RS = (M % 2) ? M : M + 1;                    // odd row stride: conflict-free x-lines
  static constexpr int PS = M * RS;
  static constexpr int ARR = M * PS;
};

template <int NQ>
__global__ void __launch_bounds__(FdmCfg<NQ>::NT)
fdm_lines_kernel(const FdmArgs a, const LlArgs ll, const __grid_constant__ FdmS<NQ - 2> P, int64_t E) {
  using C = FdmCfg<NQ>;
  constexpr int M = C::M, L = C::L, NT = C::NT, RS = C::RS, PS = C::PS, NP = NQ * NQ * NQ;
  if (a.check_done && a.s->done) return;
  __shared__ double sm[C::EPB * C::ARR];
  const int tid = threadIdx.x;
  const int el = tid / L, l = tid - el * L;
  const int ta = l / M, tb = l - ta * M;
  const int64_t e = (int64_t)blockIdx.x * C::EPB + el;
  const bool live = e < E;
  double* A = sm + el * C::ARR;
  const double* re = a.r + e * NP;
  double* ze = a.z + e * NP;
  double dot = 0.0;
  // load: interior -> shared, surface -> point Jacobi.  Four strided trips at a time with all their loads
  // (r, dinv and the copy count of every node, surface or not) issued before the first use -- ncu on the plain
  // loop: 61 % long-scoreboard stalls at the first use of each load, 21 % of the DRAM peak -- and the weight
  // 1 / copies from a small table instead of an FP64 division per surface node.
  __shared__ double s_rcp[16];
  if (tid < 16) s_rcp[tid] = 1.0 / (double)(tid > 0 ? tid : 1);
  __syncthreads();
  constexpr int LB = 4;
  for (int q0 = l; q0 < NP; q0 += LB * L) {
    double rv[LB], dv[LB];
    int mc[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int q = q0 + u * L;
      const bool ok = live && q < NP;
      rv[u] = ok ? __ldg(re + q) : 0.0;
      dv[u] = ok ? __ldg(a.dinv + e * NP + q) : 0.0;
      mc[u] = ok ? (int)__ldg(a.mult + e * NP + q) : 1;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int q = q0 + u * L;
      if (!(live && q < NP)) continue;
      const int i = q % NQ, j = (q / NQ) % NQ, k = q / (NQ * NQ);
      const bool inner = i > 0 && i < NQ - 1 && j > 0 && j < NQ - 1 && k > 0 && k < NQ - 1;
      if (inner) {
        A[(k - 1) * PS + (j - 1) * RS + (i - 1)] = rv[u];
      } else {
        const double zv = dv[u] * rv[u];
        ze[q] = zv;
        const double w = mc[u] < 16 ? s_rcp[mc[u]] : 1.0 / (double)mc[u];
        dot = fma(w * rv[u], zv, dot);
      }
    }
  }
  __syncthreads();
  double* row = A + ta * PS + tb * RS;   // x-line (j = tb, k = ta)
  double* col = A + ta * PS + tb;        // y-line (i = tb, k = ta), stride RS
  double* pil = A + ta * RS + tb;        // z-column (i = tb, j = ta), stride PS
  double in[M], out[M];
  // forward x, y
#pragma unroll
  for (int q = 0; q < M; ++q) in[q] = row[q];
  fdm_line<M, true>(P.s[0], in, out);
#pragma unroll
  for (int q = 0; q < M; ++q) row[q] = out[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < M; ++q) in[q] = col[q * RS];
  fdm_line<M, true>(P.s[1], in, out);
#pragma unroll
  for (int q = 0; q < M; ++q) col[q * RS] = out[q];
  __syncthreads();
  // forward z, eigenvalue scaling, backward z -- all in registers
#pragma unroll
  for (int q = 0; q < M; ++q) in[q] = pil[q * PS];
  fdm_line<M, true>(P.s[2], in, out);
  {
    const double ax = live ? a.esc[4 * e + 0] : 1.0, ay = live ? a.esc[4 * e + 1] : 1.0;
    const double az = live ? a.esc[4 * e + 2] : 1.0, sc = live ? a.esc[4 * e + 3] : 0.0;
    const double base = ax * P.lam[tb] + ay * P.lam[ta];
#pragma unroll
    for (int q = 0; q < M; ++q) out[q] = out[q] * (sc / (base + az * P.lam[q]));
  }
  fdm_line<M, false>(P.s[3], out, in);
#pragma unroll
  for (int q = 0; q < M; ++q) pil[q * PS] = in[q];
  __syncthreads();
  // backward y, x
#pragma unroll
  for (int q = 0; q < M; ++q) in[q] = col[q * RS];
  fdm_line<M, false>(P.s[4], in, out);
#pragma unroll
  for (int q = 0; q < M; ++q) col[q * RS] = out[q];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < M; ++q) in[q] = row[q];
  fdm_line<M, false>(P.s[5], in, out);
#pragma unroll
  for (int q = 0; q < M; ++q) row[q] = out[q];
  __syncthreads();
  // store the interior (coalesced) and its part of r.z (interior nodes have one copy: weight 1)
  for (int q0 = l; q0 < NP; q0 += LB * L) {
    double rv[LB];
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int q = q0 + u * L;
      rv[u] = (live && q < NP) ? __ldg(re + q) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int q = q0 + u * L;
      if (!(live && q < NP)) continue;
      const int i = q % NQ, j = (q / NQ) % NQ, k = q / (NQ * NQ);
      if (i > 0 && i < NQ - 1 && j > 0 && j < NQ - 1 && k > 0 && k < NQ - 1) {
        const double zv = A[(k - 1) * PS + (j - 1) * RS + (i - 1)];
        ze[q] = zv;
        dot = fma(rv[u], zv, dot);
      }
    }
  }
  const double cta = block_sum_nt<NT>(dot);
  double total;
  if (!grid_reduce<NT>(a.red, cta, (int)blockIdx.x, a.ncta, total)) return;
  if (tid == 0) a.s->rdotr_new = total;
  if (!a.finish) return;
  if (ll.nranks > 1) {
    __syncthreads();
    ll_allreduce_cta(&a.s->rdotr_new, 1, ll.peers, ll.rank, ll.nranks, ll.seq, ll.ll_error);
  }
  if (tid == 0) cg_step_scalars(a.s, a.hist);
}

}  // namespace semb
