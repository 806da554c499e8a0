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
//   DOT  : the local p.Ap of sempy/elliptic.py:51 as per-warp / per-CTA partial sums; a passenger CTA of the
//          launch that follows (the gather-scatter) adds them in index order and stores s->pAp.
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
  CgScalars* s;      // DIR: beta; DOT: pAp
  const int* done;   // nullptr or &s->done
  GridRed red;       // DOT
  int cta_off;       // DOT: number of this launch's first CTA in the logical grid
  int ncta_total;    // DOT: CTAs of the logical grid (several launches may share it)
  int64_t e0, e1;    // element range of this launch
  int l2_prefetch;
  int pdl;           // host side only: launch with the programmatic-stream-serialization attribute
};

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  // G is read exactly once per operator application: streaming (evict-first) loads.
  return __ldcs(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

// Output stores (w, and p with DIR): AX_ST_CS = 1 marks them evict-first -- by the time another kernel reads
// them the rest of this launch (>= 1 GB at the bench sizes) has passed through the L2.
#ifndef AX_ST_CS
#define AX_ST_CS 0
#endif
__device__ __forceinline__ void st_out(double* p, double v) {
  if (AX_ST_CS) __stcs(p, v);
  else *p = v;
}
__device__ __forceinline__ void st_out2(double* p, double2 v) {
  if (AX_ST_CS) __stcs(reinterpret_cast<double2*>(p), v);
  else *reinterpret_cast<double2*>(p) = v;
}

// p <- z + beta p for one value (the same expression in every kernel that forms it)
template <int DIR>
__device__ __forceinline__ double dir_update(double p_old, double r, double dinv, double beta) {
  const double z = DIR == 2 ? dinv * r : r;
  return fma(beta, p_old, z);
}

// Host launchers (one per kernel family; defined in the k1_*.cu files).  dir: 0, 1 or 2.
int launch_ax8(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);
int launch_axl(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);
int launch_ax_generic(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st);
// CTAs the launch over [e0, e1) uses; with DOT every CTA writes ax_partials_per_cta(Nq) partial sums of p.Ap
// to red.part, CTA number cta_off + blockIdx.x first
int ax_num_ctas(int Nq, int64_t e0, int64_t e1);
int ax_partials_per_cta(int Nq);

}  // namespace semb
