// K1 for Nq = 8 (N = 7, BASELINE configs 2, 4, 5): one warp per element, whole element staged in shared
// memory; see ax_common.cuh for what the kernel fuses and DESIGN.md "K1" for the measurements.
//
// Three ownership layouts of the 8x8x8 nodes, all by the same 32 lanes:
//   T: lane owns two adjacent k-columns  (i = 2*ip, 2*ip+1; j)      -> t-derivative in registers
//   R: lane owns two r-lines             (j, k), 8 consecutive i    -> r-derivative in registers
//   S: lane owns two s-lines             (i, k), 8 values of j      -> s-derivative in registers
// D is a kernel parameter, so all 64 DFMAs of a line use constant-bank operands (LDCU -> uniform
// registers).  Between layouts the data crosses shared memory once (15 eight-byte accesses per node in
// total instead of ~32 for a plane sweep).  Padding: row stride 10 doubles, plane stride 88 doubles
// makes every access pattern below bank-conflict free (R: LDS.128 by 8 lanes with distinct j; S: LDS.64
// by 16 lanes = 8 i x 2 k; T: LDS.128 by 8 lanes = 4 ip x {j, j+4}).
#include <cstring>

#include "ax_common.cuh"

namespace semb {

constexpr int AX8_RS = 10;
constexpr int AX8_PS = 88;
constexpr int AX8_ARR = 8 * AX8_PS;  // doubles per staged array
#ifndef AX8_WARPS_N
#define AX8_WARPS_N 4
#endif
// Where the L2 prefetch of G is issued (build knob, see k1_axl.cuh): 0 = first thing, 1 = behind the loads of P1
// (default: 183.2 -> 182.4 us plain, 421 -> 415 us per PCG iteration at 32^3), 2 = after P1.
#ifndef AX_PF_POS
#define AX_PF_POS -1
#endif
constexpr int AX8_PF = AX_PF_POS < 0 ? 1 : (AX_PF_POS == 3 ? 1 : AX_PF_POS);
#ifndef AX8_GPD
#define AX8_GPD 1  // planes of G in flight ahead of the plane being used (build knob)
#endif
constexpr int AX8_WARPS = AX8_WARPS_N;  // elements (= warps) per CTA; 16 warps per SM whatever the split
constexpr int AX8_SMEM = AX8_WARPS * 2 * AX8_ARR * (int)sizeof(double);

// D is replicated once per phase in the parameter block.  With a single copy the compiler CSEs the
// loads of all 64 entries across the five phases, tries to keep them live for the whole kernel, runs
// out of uniform registers and spills them into vector registers (255 regs + local memory).  Private
// copies make every phase load its entries just in time with LDCU into uniform registers, which DFMA
// reads directly.  Both lines of a lane go through the contraction together, so each entry of D is
// fetched once and consumed at once by two DFMAs: few uniform registers live, no UR spills.
struct DMat8x {
  double d[5][64];
};

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

template <bool MASK, bool DOT, int DIR>
__global__ void __launch_bounds__(AX8_WARPS * 32, 16 / AX8_WARPS)
ax8_kernel(const __grid_constant__ AxArgs a, const __grid_constant__ DMat8x Dp) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.done != nullptr && *a.done) return;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* A = smem + warp * (2 * AX8_ARR);
  double* B = A + AX8_ARR;
  // No branch around the element body: a warp past the end recomputes the last element but stores
  // nothing and contributes 0 to the dot.  Keeping the warp provably converged lets the compiler use
  // the uniform datapath (LDCU/UR) for D and plain WARPSYNCs.
  const int64_t e_raw = a.e0 + (int64_t)blockIdx.x * AX8_WARPS + warp;
  const bool live = e_raw < a.e1;
  const int64_t e = live ? e_raw : a.e1 - 1;
  double dot = 0.0;

  {
    // T layout
    const int ip = lane & 3;
    const int tj = ((lane >> 2) & 1) * 4 + (lane >> 3);
    const int t_g = tj * 8 + 2 * ip;            // offset inside a k-plane, global
    const int t_s = tj * AX8_RS + 2 * ip;       // offset inside a k-plane, shared
    const double* ue = a.u + e * 512;
    const double* ge = a.G6 + e * (6 * 512);

    // The register loads of G run one plane (3 KB per warp) ahead of their use: with 16 warps per SM
    // that is ~30 KB in flight per SM, short of what 6.5 TB/s x DRAM latency needs (measured: 79 % of
    // the HBM peak).  One lane therefore asks the L2 for the element's whole 24 KB block of G up
    // front (TMA-style bulk prefetch, no registers, no shared memory); the plane loads then hit L2.
    // Measured 206 -> 178 us per launch at 32^3 elements (92 % of the measured HBM peak).
    auto prefetch_g = [&]() {
      if (a.l2_prefetch && lane == 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ge), "r"(6 * 512 * 8) : "memory");
    };
    if (AX8_PF == 0) prefetch_g();

    // P1: u -> registers (T) and shared; with DIR the direction update happens here
    double2 uc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) uc[k] = __ldg(reinterpret_cast<const double2*>(ue + k * 64 + t_g));
    if (DIR != 0) {
      const double beta = a.s->beta;
      const double* re = a.r + e * 512;
      const double* de = a.dinv + e * 512;
      double2 rv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) rv[k] = __ldg(reinterpret_cast<const double2*>(re + k * 64 + t_g));
      if (DIR == 2) {
        double2 dv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) dv[k] = __ldg(reinterpret_cast<const double2*>(de + k * 64 + t_g));
        if (AX8_PF == 1) prefetch_g();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uc[k].x = dir_update<2>(uc[k].x, rv[k].x, dv[k].x, beta);
          uc[k].y = dir_update<2>(uc[k].y, rv[k].y, dv[k].y, beta);
        }
      } else {
        if (AX8_PF == 1) prefetch_g();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uc[k].x = dir_update<1>(uc[k].x, rv[k].x, 0.0, beta);
          uc[k].y = dir_update<1>(uc[k].y, rv[k].y, 0.0, beta);
        }
      }
      if (live) {
        double* pe = a.p_out + e * 512;
#pragma unroll
        for (int k = 0; k < 8; ++k) st_out2(pe + k * 64 + t_g, uc[k]);
      }
    }
    if (AX8_PF == 1 && DIR == 0) prefetch_g();
    if (AX8_PF == 2) prefetch_g();
    // first plane of G requested now so it overlaps the r/s passes; ring of AX8_GPD + 1 register sets
    constexpr int GPD = AX8_GPD, GS = GPD + 1;
    double2 g[GS][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) g[0][c] = ld_stream2(ge + c * 512 + t_g);
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
      double x[2][8], o[2][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        x[0][m] = colA[m * AX8_RS];
        x[1][m] = colA[L1 + m * AX8_RS];
      }
      line8x2<false>(Dp.d[0], x, o);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
    }
    __syncwarp();
    // P2b: r-derivative, in place on A (each line is private to its lane)
    {
      double x[2][8], o[2][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 v0 = lds2(rowA + 2 * c), v1 = lds2(rowA + L1 + 2 * c);
        x[0][2 * c] = v0.x;
        x[0][2 * c + 1] = v0.y;
        x[1][2 * c] = v1.x;
        x[1][2 * c + 1] = v1.y;
      }
      line8x2<false>(Dp.d[1], x, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts2(rowA + 2 * c, make_double2(o[0][2 * c], o[0][2 * c + 1]));
        sts2(rowA + L1 + 2 * c, make_double2(o[1][2 * c], o[1][2 * c + 1]));
      }
    }
    __syncwarp();

    // P3: t-derivative in registers, apply G, start D^T in t; wr -> A, ws -> B
#pragma unroll
    for (int q = 1; q < GPD; ++q) {
#pragma unroll
      for (int c = 0; c < 6; ++c) g[q][c] = ld_stream2(ge + c * 512 + q * 64 + t_g);
    }
    double2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_double2(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k + GPD < 8) {
#pragma unroll
        for (int c = 0; c < 6; ++c) g[(k + GPD) % GS][c] = ld_stream2(ge + c * 512 + (k + GPD) * 64 + t_g);
      }
      const double2(&gk)[6] = g[k % GS];
      const double2 ur = lds2(A + k * AX8_PS + t_s);
      const double2 us = lds2(B + k * AX8_PS + t_s);
      double2 ut = make_double2(0.0, 0.0);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        ut.x = fma(Dp.d[2][k * 8 + m], uc[m].x, ut.x);
        ut.y = fma(Dp.d[2][k * 8 + m], uc[m].y, ut.y);
      }
      double2 wr, ws, wt;
      wr.x = gk[0].x * ur.x + gk[1].x * us.x + gk[2].x * ut.x;
      wr.y = gk[0].y * ur.y + gk[1].y * us.y + gk[2].y * ut.y;
      ws.x = gk[1].x * ur.x + gk[3].x * us.x + gk[4].x * ut.x;
      ws.y = gk[1].y * ur.y + gk[3].y * us.y + gk[4].y * ut.y;
      wt.x = gk[2].x * ur.x + gk[4].x * us.x + gk[5].x * ut.x;
      wt.y = gk[2].y * ur.y + gk[4].y * us.y + gk[5].y * ut.y;
      sts2(A + k * AX8_PS + t_s, wr);
      sts2(B + k * AX8_PS + t_s, ws);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        acc[kk].x = fma(Dp.d[2][k * 8 + kk], wt.x, acc[kk].x);
        acc[kk].y = fma(Dp.d[2][k * 8 + kk], wt.y, acc[kk].y);
      }
      __syncwarp();  // also keeps the compiler from hoisting every plane's G loads (register blow-up)
    }

    // P4: D^T along r (A, in place) and along s (B, in place)
    {
      double x[2][8], o[2][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double2 v0 = lds2(rowA + 2 * c), v1 = lds2(rowA + L1 + 2 * c);
        x[0][2 * c] = v0.x;
        x[0][2 * c + 1] = v0.y;
        x[1][2 * c] = v1.x;
        x[1][2 * c + 1] = v1.y;
      }
      line8x2<true>(Dp.d[3], x, o);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sts2(rowA + 2 * c, make_double2(o[0][2 * c], o[0][2 * c + 1]));
        sts2(rowA + L1 + 2 * c, make_double2(o[1][2 * c], o[1][2 * c + 1]));
      }
    }
    {
      double x[2][8], o[2][8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        x[0][m] = colB[m * AX8_RS];
        x[1][m] = colB[L1 + m * AX8_RS];
      }
      line8x2<true>(Dp.d[4], x, o);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        colB[m * AX8_RS] = o[0][m];
        colB[L1 + m * AX8_RS] = o[1][m];
      }
    }
    __syncwarp();

    // P5: sum the three parts, mask, local dot, store (T layout, coalesced 512 B per plane)
    double* we = a.w + e * 512;
    const unsigned long long* mbits = reinterpret_cast<const unsigned long long*>(a.maskbits);
    const int bit = tj * 8 + 2 * ip;
    // all eight mask words first: behind the stores to w the compiler cannot move them (possible aliasing),
    // and one exposed load latency per plane adds up
    unsigned long long mw[8];
    if (MASK) {
#pragma unroll
      for (int k = 0; k < 8; ++k) mw[k] = __ldg(mbits + e * 8 + k);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double2 r2 = lds2(A + k * AX8_PS + t_s);
      const double2 s2 = lds2(B + k * AX8_PS + t_s);
      double2 o;
      o.x = (r2.x + s2.x) + acc[k].x;
      o.y = (r2.y + s2.y) + acc[k].y;
      if (MASK) {
        const unsigned long long mb = mw[k] >> bit;
        if (mb & 1ull) o.x = 0.0;
        if (mb & 2ull) o.y = 0.0;
      }
      if (DOT) dot = fma(o.x, uc[k].x, fma(o.y, uc[k].y, dot));
      if (live) st_out2(we + k * 64 + t_g, o);
    }
  }
  if (DOT) {
    // One partial per warp (lanes added in a fixed shuffle order), no barrier, no atomic: the warps of a CTA
    // leave independently.  The partials are summed in index order by a passenger CTA of the launch that
    // follows (gs_fixed_kernel / reduce_partials_kernel) -- the in-kernel grid reduction tried before cost a
    // fence, a ticket atomic and two barriers per CTA (15 % of this kernel's stall samples, ncu).
    double v = live ? dot : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) a.red.part[(int64_t)(a.cta_off + (int)blockIdx.x) * AX8_WARPS + warp] = v;
  }
}

int ax8_num_ctas(int64_t e0, int64_t e1) { return (int)((e1 - e0 + AX8_WARPS - 1) / AX8_WARPS); }
int ax8_partials_per_cta() { return AX8_WARPS; }

template <bool MASK, bool DOT, int DIR>
static int launch_ax8_t(const AxArgs& a, const DMat8x& d8, cudaStream_t st, int device) {
  static bool attr_done[64] = {false};  // per instantiation and device
  if (!attr_done[device & 63]) {
    SEMB_CUDA(cudaFuncSetAttribute(ax8_kernel<MASK, DOT, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, AX8_SMEM));
    SEMB_CUDA(cudaFuncSetAttribute(ax8_kernel<MASK, DOT, DIR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done[device & 63] = true;
  }
  SEMB_CUDA(launch_maybe_pdl(ax8_kernel<MASK, DOT, DIR>, dim3((unsigned)ax8_num_ctas(a.e0, a.e1)), dim3(AX8_WARPS * 32), AX8_SMEM, st,
                             a.pdl != 0, a, d8));
  return 0;
}

// The variants the library uses: the plain operator (semb_ax) and the CG form (dot + direction update).
int launch_ax8(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  DMat8x d8;
  for (int q = 0; q < 5; ++q) std::memcpy(d8.d[q], D.d, sizeof(d8.d[q]));
  SEMB_CHECK((dot && (dir == 1 || dir == 2)) || (!dot && dir == 0), "launch_ax8: unsupported variant");
  if (!dot) return mask ? launch_ax8_t<true, false, 0>(a, d8, st, device) : launch_ax8_t<false, false, 0>(a, d8, st, device);
  if (dir == 1) return mask ? launch_ax8_t<true, true, 1>(a, d8, st, device) : launch_ax8_t<false, true, 1>(a, d8, st, device);
  return mask ? launch_ax8_t<true, true, 2>(a, d8, st, device) : launch_ax8_t<false, true, 2>(a, d8, st, device);
}

}  // namespace semb
