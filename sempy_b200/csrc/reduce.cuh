// Deterministic (fixed-order, atomic-free sums) block and grid reductions shared by the operator and
// the CG vector kernels.  The only atomics are ticket counters that elect the CTA which finishes last;
// the sums themselves never depend on which CTA that is, so every dot product of the CG loop is bitwise
// repeatable (sempy/elliptic.py:33,37,51,60 are plain np.dot calls: any fixed order is a valid one).
#pragma once

#include "common.cuh"

namespace semb {

// Sum over the CTA; the result is returned in every thread.  NT threads; full warps use shuffles, a
// trailing partial warp goes through shared memory in thread order.
template <int NT>
__device__ __forceinline__ double block_sum_nt(double v) {
  constexpr int NW = (NT + 31) / 32;
  __shared__ double s_w[NW];
  __shared__ double s_out;
  const int tid = threadIdx.x;
  if (NT % 32 == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) s_w[tid >> 5] = v;
    __syncthreads();
  } else {
    __shared__ double s_all[NT];
    s_all[tid] = v;
    __syncthreads();
    if ((tid & 31) == 0) {
      double s = 0.0;
      const int hi = (tid + 32 < NT) ? tid + 32 : NT;
      for (int q = tid; q < hi; ++q) s += s_all[q];
      s_w[tid >> 5] = s;
    }
    __syncthreads();
  }
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NW; ++q) s += s_w[q];
    s_out = s;
  }
  __syncthreads();
  const double r = s_out;
  __syncthreads();  // s_out / s_w may be reused by the next call
  return r;
}

// Two-level grid reduction.  CTA number `cta` of `ncta` (one logical grid, possibly spread over several
// launches on one stream) contributes `value` (taken from thread 0).  CTAs are grouped by RED_GROUP
// consecutive numbers; the CTA that completes a group sums the group's partials in index order, the CTA
// that completes the last group sums the group sums in index order.  Returns true in the single CTA
// that holds the grand total (in every thread of it).  Tickets are left at zero for the next use.
constexpr int RED_GROUP = 64;
struct GridRed {
  double* part;        // [ncta]
  double* part2;       // [ceil(ncta / RED_GROUP)]
  unsigned int* tick;  // [ceil(ncta / RED_GROUP) + 1]; the last entry is the top-level ticket
};

template <int NT>
__device__ __forceinline__ bool grid_reduce(const GridRed& g, double value, int cta, int ncta, double& total) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int grp = cta / RED_GROUP;
  const int ngrp = (ncta + RED_GROUP - 1) / RED_GROUP;
  if (tid == 0) {
    __stcg(g.part + cta, value);
    __threadfence();  // the partial is visible before the ticket
    const int gsize = min(RED_GROUP, ncta - grp * RED_GROUP);
    s_last = atomicAdd(g.tick + grp, 1u) == (unsigned int)(gsize - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double v = 0.0;
  for (int q = tid; q < RED_GROUP; q += NT) {
    const int src = grp * RED_GROUP + q;
    if (src < ncta) v += __ldcg(g.part + src);  // NT >= RED_GROUP: one term per thread; else strided, fixed
  }
  v = block_sum_nt<NT>(v);
  if (tid == 0) {
    __stcg(g.part2 + grp, v);
    g.tick[grp] = 0u;
    __threadfence();
    s_last = atomicAdd(g.tick + ngrp, 1u) == (unsigned int)(ngrp - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double w = 0.0;
  for (int q = tid; q < ngrp; q += NT) w += __ldcg(g.part2 + q);
  w = block_sum_nt<NT>(w);
  if (tid == 0) g.tick[ngrp] = 0u;
  total = w;
  return true;
}

}  // namespace semb
