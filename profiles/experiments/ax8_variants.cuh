// Tuning variants of the Nq = 8 fused Ax kernel (see ax_kernels.cuh for the
// layouts).  Selected at run time with SEMB_AX_VARIANT for A/B measurements
// on the GPU; the winner becomes the default in launch_ax.
//
//   bit 0  S-lines as adjacent pairs: LDS.128/STS.128 instead of 2 x LDS.64 (less MIO pressure)
//   bit 1  P3 fetches ur/us of plane k+1 before computing plane k (short-scoreboard stalls)
//   bit 2  CTA dot partial without __syncthreads (last warp to arrive adds the four warp sums)
//   bit 3  t-direction kept in place (ut -> wt), D^T_t applied in P5; u is not held in registers
//          after P1 (the dot re-reads it, L2 hit); G prefetched two planes ahead
//   bit 4  bulk L2 prefetch (cp.async.bulk.prefetch.L2) of the element's 24 KB of G at kernel start
//   bit 5  (with bit 4) request the element pf_dist ahead instead (G and u): later warps start warm
#pragma once

#include "ax_kernels.cuh"

namespace semb {

template <bool MASK, bool DOT, int VAR>
__global__ void __launch_bounds__(AX8_WARPS * 32, 4)
ax8v_kernel(const double* __restrict__ u, const double* __restrict__ G6, double* __restrict__ w,
            const unsigned long long* __restrict__ maskbits, double* __restrict__ partials, int64_t e0,
            int64_t e1, const __grid_constant__ DMat8x Dp, const int* __restrict__ done, int pf_dist) {
  constexpr bool S128 = VAR & 1, PFRS = VAR & 2, NOSYNC = VAR & 4, INPLACE = VAR & 8, L2PF = VAR & 16, L2AHEAD = VAR & 32;
  if (done != nullptr && *done) return;
  extern __shared__ __align__(16) double smem[];
  __shared__ double s_wsum[AX8_WARPS];
  __shared__ unsigned int s_arrived;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (DOT && NOSYNC) {
    if (threadIdx.x == 0) s_arrived = 0;
    __syncthreads();
  }
  double* A = smem + warp * (2 * AX8_ARR);
  double* B = A + AX8_ARR;
  const int64_t e_raw = e0 + (int64_t)blockIdx.x * AX8_WARPS + warp;
  const bool live = e_raw < e1;
  const int64_t e = live ? e_raw : e1 - 1;
  double dot = 0.0;

  const int ip = lane & 3;
  const int tj = ((lane >> 2) & 1) * 4 + (lane >> 3);
  const int t_g = tj * 8 + 2 * ip;
  const int t_s = tj * AX8_RS + 2 * ip;
  const double* ue = u + e * 512;
  const double* ge = G6 + e * (6 * 512);

  // bit 4: one lane asks the L2 for the element's whole G block (24 KB, contiguous) right away, so
  // that the per-plane register loads of P3 find it in L2 instead of paying the DRAM latency
  if (L2PF && lane == 0) {
    // bit 5: the block was already requested by the warp that ran pf_dist elements earlier; request the
    // one pf_dist elements ahead (G and u) so that a later warp starts with everything in L2
    if (!L2AHEAD || e_raw < e0 + pf_dist)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge), "r"(6 * 512 * 8) : "memory");
    if (L2AHEAD && e_raw + pf_dist < e1) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(G6 + (e_raw + pf_dist) * (6 * 512)), "r"(6 * 512 * 8) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(u + (e_raw + pf_dist) * 512), "r"(512 * 8) : "memory");
    }
  }
  // P1
  double2 uc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) uc[k] = __ldg(reinterpret_cast<const double2*>(ue + k * 64 + t_g));
  double2 g[6], gn[6], gn2[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) g[c] = ld_stream2(ge + c * 512 + t_g);
  if (INPLACE) {
#pragma unroll
    for (int c = 0; c < 6; ++c) gn[c] = ld_stream2(ge + c * 512 + 64 + t_g);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sts2(A + k * AX8_PS + t_s, uc[k]);
  __syncwarp();

  const int lk0 = lane >> 3, lq = lane & 7;
  double* rowA = A + lk0 * AX8_PS + lq * AX8_RS;
  constexpr int L1 = 4 * AX8_PS;
  // S layout, scalar form: lines (i = lq, k = lk0) and (i = lq, k = lk0 + 4)
  double* colA = A + lk0 * AX8_PS + lq;
  double* colB = B + lk0 * AX8_PS + lq;
  // S layout, pair form: lines (i = 2c, k) and (i = 2c + 1, k), c = lane & 3, k = lane >> 2
  double* pairA = A + (lane >> 2) * AX8_PS + 2 * (lane & 3);
  double* pairB = B + (lane >> 2) * AX8_PS + 2 * (lane & 3);

  // P2a: s-derivative A -> B
  {
    double a[2][8], o[2][8];
    if (S128) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const double2 v = lds2(pairA + m * AX8_RS);
        a[0][m] = v.x;
        a[1][m] = v.y;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        a[0][m] = colA[m * AX8_RS];
        a[1][m] = colA[L1 + m * AX8_RS];
      }
    }
    line8x2<false>(Dp.d[0], a, o);
    if (S128) {
#pragma unroll
      for (int m = 0; m < 8; ++m) sts2(pairB + m * AX8_RS, make_double2(o[0][m], o[1][m]));
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
    }
  }
  __syncwarp();
  // P2b: r-derivative in place on A
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
  double2 acc[8];  // !INPLACE: D^T_t accumulators;  INPLACE: ut, overwritten by wt
  if (INPLACE) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double2 t = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        t.x = fma(Dp.d[2][k * 8 + m], uc[m].x, t.x);
        t.y = fma(Dp.d[2][k * 8 + m], uc[m].y, t.y);
      }
      acc[k] = t;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_double2(0.0, 0.0);
  }
  double2 ur = lds2(A + t_s), us = lds2(B + t_s);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (INPLACE) {
      if (k < 6) {
#pragma unroll
        for (int c = 0; c < 6; ++c) gn2[c] = ld_stream2(ge + c * 512 + (k + 2) * 64 + t_g);
      }
    } else if (k < 7) {
#pragma unroll
      for (int c = 0; c < 6; ++c) gn[c] = ld_stream2(ge + c * 512 + (k + 1) * 64 + t_g);
    }
    double2 urn, usn;
    if (PFRS) {
      if (k < 7) {
        urn = lds2(A + (k + 1) * AX8_PS + t_s);
        usn = lds2(B + (k + 1) * AX8_PS + t_s);
      }
    } else if (k > 0) {
      ur = lds2(A + k * AX8_PS + t_s);
      us = lds2(B + k * AX8_PS + t_s);
    }
    double2 ut;
    if (INPLACE) {
      ut = acc[k];
    } else {
      ut = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        ut.x = fma(Dp.d[2][k * 8 + m], uc[m].x, ut.x);
        ut.y = fma(Dp.d[2][k * 8 + m], uc[m].y, ut.y);
      }
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
    if (INPLACE) {
      acc[k] = wt;
    } else {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        acc[kk].x = fma(Dp.d[2][k * 8 + kk], wt.x, acc[kk].x);
        acc[kk].y = fma(Dp.d[2][k * 8 + kk], wt.y, acc[kk].y);
      }
    }
    if (INPLACE) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        g[c] = gn[c];
        gn[c] = gn2[c];
      }
    } else if (k < 7) {
#pragma unroll
      for (int c = 0; c < 6; ++c) g[c] = gn[c];
    }
    if (PFRS && k < 7) {
      ur = urn;
      us = usn;
    }
    __syncwarp();
  }

  // the dot needs u again in P5: with INPLACE it is re-read (L2 hit) while P4 computes
  double2 u5[8];
  if (DOT && INPLACE) {
#pragma unroll
    for (int k = 0; k < 8; ++k) u5[k] = __ldg(reinterpret_cast<const double2*>(ue + k * 64 + t_g));
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
    if (S128) {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const double2 v = lds2(pairB + m * AX8_RS);
        a[0][m] = v.x;
        a[1][m] = v.y;
      }
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        a[0][m] = colB[m * AX8_RS];
        a[1][m] = colB[L1 + m * AX8_RS];
      }
    }
    line8x2<true>(Dp.d[4], a, o);
    if (S128) {
#pragma unroll
      for (int m = 0; m < 8; ++m) sts2(pairB + m * AX8_RS, make_double2(o[0][m], o[1][m]));
    } else {
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
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
    double2 t2;
    if (INPLACE) {
      t2 = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        t2.x = fma(Dp.d[2][m * 8 + k], acc[m].x, t2.x);
        t2.y = fma(Dp.d[2][m * 8 + k], acc[m].y, t2.y);
      }
    } else {
      t2 = acc[k];
    }
    double2 o;
    o.x = (r2.x + s2.x) + t2.x;
    o.y = (r2.y + s2.y) + t2.y;
    if (MASK) {
      const unsigned long long mb = __ldg(maskbits + e * 8 + k) >> bit;
      if (mb & 1ull) o.x = 0.0;
      if (mb & 2ull) o.y = 0.0;
    }
    if (DOT) {
      const double2 uk = INPLACE ? u5[k] : uc[k];
      dot = fma(o.x, uk.x, fma(o.y, uk.y, dot));
    }
    *reinterpret_cast<double2*>(we + k * 64 + t_g) = o;
  }

  if (DOT) {
    if (!live) dot = 0.0;
    if (NOSYNC) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (lane == 0) {
        s_wsum[warp] = dot;
        __threadfence_block();
        if (atomicAdd(&s_arrived, 1u) == AX8_WARPS - 1) {
          __threadfence_block();
          const volatile double* ws = s_wsum;
          partials[blockIdx.x] = ((ws[0] + ws[1]) + ws[2]) + ws[3];  // fixed order -> repeatable
        }
      }
    } else {
      block_sum_to<AX8_WARPS * 32>(dot, partials + blockIdx.x);
    }
  }
}

}  // namespace semb
