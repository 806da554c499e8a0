// K1 for Nq = 9 ... 12 (k1_axl.cuh); split from the higher orders only to compile in parallel.
#include "k1_axl.cuh"

namespace semb {
int launch_axl_small(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device);

int launch_axl_lo(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  switch (Nq) {
    case 9: return launch_axl_nq<9>(a, D, mask, dot, dir, st, device);
    case 10: return launch_axl_nq<10>(a, D, mask, dot, dir, st, device);
    case 11: return launch_axl_nq<11>(a, D, mask, dot, dir, st, device);
    case 12: return launch_axl_nq<12>(a, D, mask, dot, dir, st, device);
    default: return launch_axl_small(Nq, a, D, mask, dot, dir, st, device);
  }
}
}  // namespace semb
