// EXPERIMENTAL (off by default, knob "ax_fused_gs"): the N = 7 operator kernel and the gather-scatter
// in ONE persistent kernel, so that the direct-stiffness sum reads the operator output while it is still
// in L2 instead of sweeping the vector a second time (the separate pass moves 18 B/node of sector traffic
// at 48 % of the DRAM peak, profiles/r01_summary.md).
//
// Every warp is a worker: it takes work items from a global counter in increasing order.  Item e < E:
// compute element e exactly as ax8_kernel does, publish elem_flag[e] = tag.  Every item e >= lag also
// CLOSES element e - lag: it completes the segments whose highest copy lives in that element
// (closure_tables.h).  All other copies of such a segment live in lower-numbered elements, which were
// dequeued earlier, so the only waits are on elements that some running warp already owns: no cycles,
// no dependence on co-residency.  The lag keeps the waits rare (the dependencies finished long ago) while
// the data (lag * 4 KB, about 10 MB) is still in the 126 MB L2.  Items E .. E+lag-1 only close.
// Reads of other warps' output bypass L1 (ld.global.cg); flags are published after __threadfence().
#pragma once

#include "ax_kernels.cuh"

namespace semb {

struct FusedGsTables {
  const int32_t* base;
  const uint16_t* n2;
  const uint16_t* n4;
  const uint16_t* n8;
  const int32_t* ids;
  const int32_t* dep_start;
  const int32_t* dep_elem;
};

constexpr unsigned long long FUSED_TIMEOUT_NS = 20000000000ull;

__device__ __forceinline__ unsigned long long fused_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// wait until element d has been published in this launch; false on timeout
__device__ __forceinline__ bool fused_wait(const unsigned int* elem_flag, int d, unsigned int tag, int* err) {
  unsigned int spins = 0;
  unsigned long long t0 = 0;
  while (*(volatile const unsigned int*)(elem_flag + d) != tag) {
    if ((++spins & 255u) == 0) {
      const unsigned long long now = fused_now();
      if (t0 == 0) t0 = now;
      if (now - t0 > FUSED_TIMEOUT_NS || *(volatile int*)err) {
        *(volatile int*)err = 1;
        return false;
      }
    }
  }
  return true;
}

// complete the segments closed by element ec (all 32 lanes of the warp take part)
__device__ __forceinline__ void fused_close(double* __restrict__ w, const FusedGsTables& gt, int ec,
                                            const unsigned int* elem_flag, unsigned int tag, int* err, int lane) {
  const int d0 = gt.dep_start[ec], d1 = gt.dep_start[ec + 1];
  bool ok = true;
  for (int q = d0 + lane; q < d1; q += 32) ok &= fused_wait(elem_flag, gt.dep_elem[q], tag, err);
  if (lane == 0) ok &= fused_wait(elem_flag, ec, tag, err);
  __threadfence();
  __syncwarp();
  if (!__all_sync(0xffffffffu, ok)) return;
  const int32_t* ids = gt.ids + gt.base[ec];
  const int n2 = gt.n2[ec], n4 = gt.n4[ec], n8 = gt.n8[ec];
  for (int t = lane; t < n2; t += 32) {
    const int2 id = __ldg(reinterpret_cast<const int2*>(ids) + t);
    const double s = __ldcg(w + id.x) + __ldcg(w + id.y);
    w[id.x] = s;
    w[id.y] = s;
  }
  const int32_t* q4 = ids + 2 * n2;
  for (int t = lane; t < n4; t += 32) {
    const int2 a = __ldg(reinterpret_cast<const int2*>(q4) + 2 * t), b = __ldg(reinterpret_cast<const int2*>(q4) + 2 * t + 1);
    const double s = ((__ldcg(w + a.x) + __ldcg(w + a.y)) + __ldcg(w + b.x)) + __ldcg(w + b.y);
    w[a.x] = s;
    w[a.y] = s;
    w[b.x] = s;
    w[b.y] = s;
  }
  const int32_t* q8 = q4 + 4 * n4;
  for (int t = lane; t < n8; t += 32) {
    int id[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int2 v = __ldg(reinterpret_cast<const int2*>(q8) + 4 * t + c);
      id[2 * c] = v.x;
      id[2 * c + 1] = v.y;
    }
    double s = __ldcg(w + id[0]);
#pragma unroll
    for (int c = 1; c < 8; ++c) s += __ldcg(w + id[c]);
#pragma unroll
    for (int c = 0; c < 8; ++c) w[id[c]] = s;
  }
}

template <bool MASK, bool DOT>
__global__ void __launch_bounds__(AX8_WARPS * 32, 4)
ax8_fused_kernel(const double* __restrict__ u, const double* __restrict__ G6, double* __restrict__ w,
                 const unsigned long long* __restrict__ maskbits, double* __restrict__ partials, double* __restrict__ dot_out,
                 int E, const __grid_constant__ DMat8x Dp, const int* __restrict__ done, const FusedGsTables gt,
                 unsigned int* __restrict__ work, unsigned int* __restrict__ elem_flag, unsigned int tag, int lag,
                 int* __restrict__ err) {
  if (done != nullptr && *done) return;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* A = smem + warp * (2 * AX8_ARR);
  double* B = A + AX8_ARR;
  const int ip = lane & 3;
  const int tj = ((lane >> 2) & 1) * 4 + (lane >> 3);
  const int t_g = tj * 8 + 2 * ip;
  const int t_s = tj * AX8_RS + 2 * ip;
  const int lk0 = lane >> 3, lq = lane & 7;
  double* colA = A + lk0 * AX8_PS + lq;
  double* colB = B + lk0 * AX8_PS + lq;
  double* rowA = A + lk0 * AX8_PS + lq * AX8_RS;
  constexpr int L1 = 4 * AX8_PS;
  const int total = E + lag;
  double dot = 0.0;

  // work[0]: next item; work[1]: finished-warp ticket
  int item = 0;
  if (lane == 0) item = (int)atomicAdd(work, 1u);
  item = __shfl_sync(0xffffffffu, item, 0);
  if (item < E && lane == 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G6 + (int64_t)item * (6 * 512)), "r"(6 * 512 * 8) : "memory");

  while (item < total) {
    // take the next item now: its index (and its block of G, via the L2 prefetch) arrive while this one runs
    int nxt = 0;
    if (lane == 0) nxt = (int)atomicAdd(work, 1u);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    if (nxt < E && lane == 0)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G6 + (int64_t)nxt * (6 * 512)), "r"(6 * 512 * 8) : "memory");

    if (item < E) {
      const int64_t e = item;
      const double* ue = u + e * 512;
      const double* ge = G6 + e * (6 * 512);
      // P1
      double2 uc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) uc[k] = __ldg(reinterpret_cast<const double2*>(ue + k * 64 + t_g));
      double2 g[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) g[c] = ld_stream2(ge + c * 512 + t_g);
#pragma unroll
      for (int k = 0; k < 8; ++k) sts2(A + k * AX8_PS + t_s, uc[k]);
      __syncwarp();
      // P2a
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
      // P2b
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
      // P3
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
        __syncwarp();
      }
      // P4
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
      // P5
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
      // publish: every lane's stores are device-visible before lane 0 raises the flag
      __threadfence();
      __syncwarp();
      if (lane == 0) *(volatile unsigned int*)(elem_flag + item) = tag;
    }
    if (item >= lag) fused_close(w, gt, item - lag, elem_flag, tag, err, lane);
    item = nxt;
  }

  if (DOT) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  // per-warp partial, then the warp that finishes last adds them in a fixed order and re-arms the counters
  const unsigned int nwarps = gridDim.x * AX8_WARPS;
  int last = 0;
  if (lane == 0) {
    if (DOT) partials[blockIdx.x * AX8_WARPS + warp] = dot;
    __threadfence();
    last = atomicAdd(work + 1, 1u) == nwarps - 1;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    __threadfence();
    if (DOT) {
      double s = 0.0;
      for (unsigned int q = lane; q < nwarps; q += 32) s += __ldcg(partials + q);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) *dot_out = s;
    }
    if (lane == 0) {
      work[0] = 0;
      work[1] = 0;
    }
  }
}

}  // namespace semb
