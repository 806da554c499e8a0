// K1 for Nq = 13 ... 16 (k1_axl.cuh; N = 15 is BASELINE config 3).
#include "k1_axl.cuh"

namespace semb {
int launch_axl_lo(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);

int launch_axl(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  switch (Nq) {
    case 13: return launch_axl_nq<13>(a, D, mask, dot, dir, st, device);
    case 14: return launch_axl_nq<14>(a, D, mask, dot, dir, st, device);
    case 15: return launch_axl_nq<15>(a, D, mask, dot, dir, st, device);
    case 16: return launch_axl_nq<16>(a, D, mask, dot, dir, st, device);
    default: return launch_axl_lo(Nq, a, D, mask, dot, dir, st, device);
  }
}
}  // namespace semb
