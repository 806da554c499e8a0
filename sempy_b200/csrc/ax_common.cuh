// Shared by the three K1 kernel families (k1_ax8.cu, k1_axl.cu, k1_generic.cu): the argument block and
// the host-side launchers semb200.cu calls.
//
// K1 is the fused FP64 spectral-element Poisson operator  w_e = D^T G D u_e  (sempy/elliptic.py:9-27:
// gradient -> 3x3 G -> gradient_transpose, sempy/gradient.py:6-23,39-56) with, depending on the
// template flags,
//   DIR  : the CG direction update folded into the load phase (sempy/elliptic.py:70,
//          sempy/iterative.py:83):  p <- z + beta p  with z = r (DIR = 1) or dinv * r (DIR = 2); the new p
//          is written back and used as u -- the separate direction kernel and one read of p disappear;
//   MASK : the Dirichlet mask in the epilogue (sempy/mesh.py:487-494);
//   DOT  : the local p.Ap of sempy/elliptic.py:51, reduced over the whole grid in fixed order by the CTA
//          that finishes last (reduce.cuh) and stored in s->pAp.
// HBM-bound: 64 B per node (u 8 + six G components 48 + w 8; +24 with DIR) against 12*Nq+15 flop.
#pragma once

#include "common.cuh"
#include "reduce.cuh"

namespace semb {

struct AxArgs {
  const double* u;   // input p (E*Np); with DIR it holds the OLD direction and is updated in place
  double* p_out;     // == u when DIR != 0
  const double* r;   // DIR: residual
  const double* dinv;  // DIR == 2: inverse Jacobi diagonal
  const double* G6;  // (E,6,Np)
  double* w;         // output
  const uint32_t* maskbits;
  double* eb;        // nullptr, or the compact edge buffer (E, 12*Nq-16): see eb_slot
  CgScalars* s;      // DIR: beta; DOT: pAp
  const int* done;   // nullptr or &s->done
  GridRed red;       // DOT
  int cta_off;       // DOT: number of this launch's first CTA in the logical grid
  int ncta_total;    // DOT: CTAs of the logical grid (several launches may share it)
  int64_t e0, e1;    // element range of this launch
  int l2_prefetch;
};

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  // G is read exactly once per operator application: streaming (evict-first) loads.
  return __ldcs(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// p <- z + beta p for one value (the same expression in every kernel that forms it)
template <int DIR>
__device__ __forceinline__ double dir_update(double p_old, double r, double dinv, double beta) {
  const double z = DIR == 2 ? dinv * r : r;
  return fma(beta, p_old, z);
}

// Compact "edge buffer".  On a conforming hex mesh the nodes with more than two copies are the element's
// edge and vertex nodes (12*Nq - 16 of Nq^3).  The operator kernel writes those values of its output a second
// time into a dense per-element block, so that the gather-scatter of the longer segments reads fully used
// sectors instead of one 8-byte value per 32-byte sector of the vector.  Slot of node (i, j, k), or -1:
//   [k = 0 perimeter | k = Nq-1 perimeter | corners (i, j in {0, Nq-1}) of the planes k = 1 .. Nq-2]
//   perimeter of a plane: row j = 0 (i), row j = Nq-1 (Nq + i), then (i = 0, i = Nq-1) pairs of the rows between.
__host__ __device__ __forceinline__ int eb_count(int Nq) { return 12 * Nq - 16; }
__host__ __device__ __forceinline__ int eb_slot(int i, int j, int k, int Nq) {
  const int L = Nq - 1, P = 4 * Nq - 4;
  const bool bi = i == 0 || i == L, bj = j == 0 || j == L, bk = k == 0 || k == L;
  if (bk) {
    if (!(bi || bj)) return -1;
    const int base = k == 0 ? 0 : P;
    if (j == 0) return base + i;
    if (j == L) return base + Nq + i;
    return base + 2 * Nq + (j - 1) * 2 + (i == L ? 1 : 0);
  }
  if (!(bi && bj)) return -1;
  return 2 * P + (k - 1) * 4 + (j == L ? 2 : 0) + (i == L ? 1 : 0);
}

// Host launchers (one per kernel family; defined in the k1_*.cu files).  dir: 0, 1 or 2.
int launch_ax8(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);
int launch_axl(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);
int launch_ax_generic(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st);
// CTAs the launch over [e0, e1) uses (= partials it contributes with DOT)
int ax_num_ctas(int Nq, int64_t e0, int64_t e1);

}  // namespace semb
