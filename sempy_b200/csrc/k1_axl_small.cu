// K1 for Nq = 3 ... 7 (N = 2 ... 6): k1_axl.cuh with several elements per CTA; a translation unit of its
// own only to compile in parallel.
#include "k1_axl.cuh"

namespace semb {
int launch_axl_small(int Nq, const AxArgs& a, const DMat& D, bool mask, bool dot, int dir, cudaStream_t st, int device) {
  switch (Nq) {
    case 3: return launch_axl_nq<3>(a, D, mask, dot, dir, st, device);
    case 4: return launch_axl_nq<4>(a, D, mask, dot, dir, st, device);
    case 5: return launch_axl_nq<5>(a, D, mask, dot, dir, st, device);
    case 6: return launch_axl_nq<6>(a, D, mask, dot, dir, st, device);
    case 7: return launch_axl_nq<7>(a, D, mask, dot, dir, st, device);
    default: break;
  }
  set_error("launch_axl_small: unsupported Nq " + std::to_string(Nq));
  return 2;
}
}  // namespace semb
