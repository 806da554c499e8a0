// K1 as a plain k-plane sweep, one CTA of Nq^2 threads per element with the u column and the D^T_t
// accumulators in registers: correct for every order, used for N = 1 (Nq = 2) only -- all other orders go
// through k1_ax8.cu / k1_axl.cuh.  See ax_common.cuh for the fusions.
#include <cstring>

#include "ax_common.cuh"

namespace semb {

template <int NQ, bool MASK, bool DOT, int DIR>
__global__ void __launch_bounds__(NQ* NQ)
ax_generic_kernel(const __grid_constant__ AxArgs a, const __grid_constant__ DMat Dp) {
  if (a.done != nullptr && *a.done) return;
  constexpr int NP = NQ * NQ * NQ;
  constexpr int NT = NQ * NQ;
  __shared__ double sD[NQ][NQ + 1];
  __shared__ double su[NQ][NQ + 1];
  __shared__ double swr[NQ][NQ + 1];
  __shared__ double sws[NQ][NQ + 1];
  const int i = threadIdx.x % NQ, j = threadIdx.x / NQ;
  const int64_t e = a.e0 + blockIdx.x;
  sD[j][i] = Dp.d[j * NQ + i];
  const double* ue = a.u + e * NP;
  const double* ge = a.G6 + e * (int64_t)(6 * NP);
  double uc[NQ], acc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    uc[k] = ue[k * NT + j * NQ + i];
    acc[k] = 0.0;
  }
  if (DIR != 0) {
    const double beta = a.s->beta;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int64_t q = e * NP + k * NT + j * NQ + i;
      uc[k] = dir_update<DIR>(uc[k], a.r[q], DIR == 2 ? a.dinv[q] : 0.0, beta);
      a.p_out[q] = uc[k];
    }
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
  double* we = a.w + e * NP;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const int node = k * NT + j * NQ + i;
    double o = acc[k];
    if (MASK) {
      const int64_t gi = e * NP + node;
      if ((a.maskbits[gi >> 5] >> (gi & 31)) & 1u) o = 0.0;
    }
    if (DOT) dot = fma(o, uc[k], dot);
    we[node] = o;
  }
  if (DOT) {
    const double cta = block_sum_nt<NT>(dot);
    if (threadIdx.x == 0) a.red.part[a.cta_off + (int)blockIdx.x] = cta;  // summed by the next launch's passenger CTA
  }
}

template <int NQ, bool MASK, bool DOT, int DIR>
static int launch_generic_t(const AxArgs& a, const DMat& D, cudaStream_t st) {
  ax_generic_kernel<NQ, MASK, DOT, DIR><<<(unsigned)(a.e1 - a.e0), NQ * NQ, 0, st>>>(a, D);
  SEMB_CUDA(cudaGetLastError());
  return 0;
}

template <int NQ>
static int launch_generic_nq(const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st) {
  SEMB_CHECK((dot && (dir == 1 || dir == 2)) || (!dot && dir == 0), "launch_ax_generic: unsupported variant");
  if (!dot) return mask ? launch_generic_t<NQ, true, false, 0>(a, D, st) : launch_generic_t<NQ, false, false, 0>(a, D, st);
  if (dir == 1) return mask ? launch_generic_t<NQ, true, true, 1>(a, D, st) : launch_generic_t<NQ, false, true, 1>(a, D, st);
  return mask ? launch_generic_t<NQ, true, true, 2>(a, D, st) : launch_generic_t<NQ, false, true, 2>(a, D, st);
}

int launch_ax_generic(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st) {
  switch (Nq) {
    case 2: return launch_generic_nq<2>(a, D, mask, dot, dir, st);
    case 3: return launch_generic_nq<3>(a, D, mask, dot, dir, st);
    case 4: return launch_generic_nq<4>(a, D, mask, dot, dir, st);
    case 5: return launch_generic_nq<5>(a, D, mask, dot, dir, st);
    case 6: return launch_generic_nq<6>(a, D, mask, dot, dir, st);
    case 7: return launch_generic_nq<7>(a, D, mask, dot, dir, st);
    default: break;
  }
  set_error("launch_ax_generic: unsupported Nq " + std::to_string(Nq));
  return 2;
}

int ax8_num_ctas(int64_t e0, int64_t e1);
int ax8_partials_per_cta();
// partials of p.Ap that one CTA of the operator launch writes (ax_common.cuh)
int ax_partials_per_cta(int Nq) { return Nq == 8 ? ax8_partials_per_cta() : 1; }
int ax_num_ctas(int Nq, int64_t e0, int64_t e1) {
  if (e1 <= e0) return 0;
  if (Nq == 8) return ax8_num_ctas(e0, e1);
  const int epb = (Nq >= 3 && 160 / (Nq * Nq) > 0) ? 160 / (Nq * Nq) : 1;  // AxlCfg<NQ>::EPB
  return (int)((e1 - e0 + epb - 1) / epb);
}

}  // namespace semb
