// Shared declarations of libsemb200: error handling, the context, device scalars.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/semb200.h"

namespace semb {

void set_error(const std::string& msg);

#define SEMB_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t err__ = (call);                                                              \
    if (err__ != cudaSuccess) {                                                              \
      ::semb::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " (" +       \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                    \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// NCCL is bound at run time (dlopen), not at link time: a process that also
// imports torch must end up with ONE libnccl.so.2 -- torch's bundled one when
// torch is loaded, the system one otherwise.  Linking -lnccl would pull the
// system library in first and break a later `import torch`.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
const NcclApi& nccl();  // loads on first use; ok == false if no NCCL library could be bound

#define SEMB_NCCL(call)                                                                      \
  do {                                                                                       \
    if (!::semb::nccl().ok) {                                                                \
      ::semb::set_error("NCCL library (libnccl.so.2) could not be loaded");                  \
      return 3;                                                                              \
    }                                                                                        \
    ncclResult_t err__ = ::semb::nccl().call;                                                \
    if (err__ != ncclSuccess) {                                                              \
      ::semb::set_error(std::string("nccl" #call) + ": " + ::semb::nccl().GetErrorString(err__) + " (" + \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                    \
      return 3;                                                                              \
    }                                                                                        \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Inside semb_cg on one GPU the three kernels of an iteration depend on each other only through "the previous
// grid has completed".  Launched with the programmatic-stream-serialization attribute, a kernel may become
// resident while its predecessor drains: every kernel of the chain first lets ITS dependents be scheduled
// (pdl_launch_dependents) and then waits until the predecessor has completed and its writes are visible
// (pdl_wait) before it reads anything the predecessor produced -- including the `done` flag.  Launch latency
// and the ramp-up of the next grid overlap the tail of the previous one.  Both instructions are no-ops in a
// kernel that was launched normally.
#ifdef __CUDACC__
// SEMB_PDL_EARLY = 1 lets the dependents become resident as soon as every CTA of this grid has started.  Measured:
// the waiting CTAs of the NEXT grids (the operator kernel's need 128 registers x 128 threads and 45 KB each) take the
// slots the running grid's later CTAs need: 416 -> 494 us per PCG iteration at 32^3 N=7.  Off: dependents are
// released when this grid's CTAs exit (only the launch latency is hidden).
#ifndef SEMB_PDL_EARLY
#define SEMB_PDL_EARLY 0
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if SEMB_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

#define SEMB_CHECK(cond, msg)              \
  do {                                     \
    if (!(cond)) {                         \
      ::semb::set_error(msg);              \
      return 2;                            \
    }                                      \
  } while (0)

#define SEMB_TRY(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != 0) return rc__; \
  } while (0)

constexpr int kMaxNq = 16;
constexpr int kNumSMs = 148;  // B200

// D by value in the kernel parameter space (constant bank): after unrolling, every D operand of a
// DFMA is fetched with LDCU into a uniform register and costs no vector register.
struct DMat {
  double d[kMaxNq * kMaxNq];  // row-major, leading dimension Nq
};

// Device-resident scalars of one CG solve (no host round trip per iteration).
struct CgScalars {
  double rdotr;      // current r.r (or r.z), global
  double rdotr_new;  // written by the update kernel, promoted by cg_finalize
  double pAp;
  double alpha;
  double beta;
  double TOL;
  double norm_b;
  int niter;
  int done;
  int maxit;
  int pad;
};

// One class of gather-scatter segments: nseg segments of equal length len,
// local indices stored segment-major at ids[off + s*len + c].
struct GsClass {
  int len;
  int64_t nseg;
  int64_t off;
};

constexpr int LL_MAXR = 16;  // ranks supported by the peer-memory path (one NVSwitch domain)
constexpr int LL_MAXV = 4;   // doubles per scalar all-reduce
struct HaloPeers {
  int nneigh;
  int rank_of[LL_MAXR];
  int64_t off[LL_MAXR + 1];   // my send blocks (entries)
  int64_t peer_off[LL_MAXR];  // start of my block in that neighbour's halo part (entries)
};

// Pageable host memory <-> device through a ring of pinned chunks filled / drained by a few host threads
// (a plain cudaMemcpy from pageable memory stages through ONE driver thread at ~7 GB/s, measured 37 ms
// for the 2 x 134 MB of a 32^3-element solve; the ring overlaps several memcpy threads with the DMA).
constexpr int HS_THREADS = 8;  // at most; hs_copy uses hardware threads / ranks of them, between 2 and 8
constexpr size_t HS_CHUNK = (size_t)4 << 20;  // bytes per chunk
struct HostStager {
  bool ready = false;
  cudaStream_t stream[HS_THREADS] = {};
  char* slot[HS_THREADS][2] = {};
  cudaEvent_t ev[HS_THREADS][2] = {};
  cudaEvent_t fence = nullptr;
};

struct Profile {
  bool on = false;
  int mask = 0xffff;
  std::vector<cudaEvent_t> start[SEMB_K_COUNT], stop[SEMB_K_COUNT];
  size_t used[SEMB_K_COUNT] = {0, 0, 0, 0};
  double total_ms[SEMB_K_COUNT] = {0, 0, 0, 0};
  int64_t launches[SEMB_K_COUNT] = {0, 0, 0, 0};
};

}  // namespace semb

struct semb_ctx {
  int device = 0;
  int64_t E = 0;
  int Nq = 0;
  int64_t Np = 0;
  int64_t n = 0;  // E*Np
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;

  semb::DMat D;
  bool have_D = false;
  double* G6 = nullptr;  // (E,6,Np)
  bool have_G = false;

  // gather-scatter schedule (shared segments only)
  int32_t* gs_ids = nullptr;
  int64_t gs_nids = 0;
  std::vector<semb::GsClass> gs_classes;  // fixed-length classes (2,4,8,...)
  int32_t* gs_var_start = nullptr;        // CSR offsets of the irregular rest
  int64_t gs_var_nseg = 0;
  int64_t gs_var_off = 0;
  bool have_gs = false;
  // pull form: link[q] = GS_LINK_PAIR | the other copy's local index, GS_LINK_LONG | slot of the node's
  // segment in gs_sums, or 0 (no other copy; once a halo is set also the nodes completed by the halo kernels)
  uint32_t* gs_link = nullptr;
  double* gs_sums = nullptr;  // compact sums of the segments with more than two copies
  bool pull_ok = false;       // fewer than 2^30 such segments
  double* rmult = nullptr;
  unsigned char* mult = nullptr;  // copy count per node (rmult = 1/mult), read by the CG update kernel

  // Dirichlet mask: 1 bit per local node (1 = zeroed)
  uint32_t* maskbits = nullptr;
  bool have_mask = false;

  // scratch / CG state
  double* partials = nullptr;   // per-CTA partial sums
  double* partials2 = nullptr;  // per-group sums (reduce.cuh)
  int64_t partials_cap = 0;
  unsigned int* ticket = nullptr;         // group tickets of the grid reductions (reduce.cuh)
  unsigned int* ticket_update = nullptr;  // the CG update kernel's own ticket (inside the same allocation)
  double* pass_part = nullptr;            // slice sums / ticket of the p.Ap passengers (vec_kernels.cuh)
  unsigned int* pass_tick = nullptr;
  int* info = nullptr;                    // 4 device ints for set-up checks
  semb::CgScalars* scal = nullptr;
  double* hist = nullptr;
  int64_t hist_cap = 0;
  double* vec[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // r,p,Ap,x,dinv,scratch,p'

  double* stage[2] = {nullptr, nullptr};  // pooled staging of HOST vector arguments (never freed per call)
  semb::HostStager hs;                    // pinned ring for pageable HOST arguments
  int* h_pinned = nullptr;  // pinned {done, niter}

  // multi-GPU: NCCL communicator, interface ("halo") tables of this rank's element slab
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_if_done = nullptr, ev_gathered = nullptr, ev_exchanged = nullptr;
  bool have_halo = false;
  std::vector<int> nb_rank;        // neighbour ranks
  std::vector<int64_t> nb_off;     // start of each neighbour's block in the send region (nneigh+1)
  int64_t n_send = 0;              // entries sent (= received)
  int64_t* send_start = nullptr;   // CSR over send entries -> local copies to sum
  int32_t* send_ids = nullptr;
  int64_t n_if = 0;                // interface nodes of this rank
  int64_t* src_start = nullptr;    // CSR per interface node -> offsets into hbuf, ascending rank
  int32_t* src_off = nullptr;
  int64_t* copy_start = nullptr;   // CSR per interface node -> local copies
  int32_t* copy_ids = nullptr;
  double* hbuf = nullptr;          // [send region | receive region]
  int64_t n_if_elems = 0;          // elements [0, n_if_elems) touch the interface

  // peer-memory path (NVLink): every rank maps every other rank's "LL" region (cudaIpc) and writes
  // {32-bit data, 32-bit sequence} words straight into it; the reader spins on the sequence number
  // (NCCL's LL idea), so neither the scalar all-reduce nor the halo exchange needs a NCCL call.
  bool p2p = false;
  unsigned long long* ll_local = nullptr;  // this rank's region
  std::vector<unsigned long long*> ll_peer;  // every rank's region as mapped here ([rank] = ll_local)
  unsigned long long** ll_peer_dev = nullptr;
  int* ll_error = nullptr;                 // device flag raised when a peer-memory wait times out
  int64_t ll_halo_cap = 0;                 // doubles per parity in the halo part
  semb::HaloPeers halo_peers;              // neighbour ranks, my send blocks, where they land remotely
  unsigned int seq_scalar = 0, seq_halo = 0;  // host-side sequence numbers (same on all ranks)

  // FDM preconditioner (fdm_kernels.cuh): 1-D eigenvectors / eigenvalues of the interior problem and the
  // per-element scale factors
  bool fdm_ready = false;
  double* fdm_S = nullptr;     // [m*m S | m lambda], m = Nq - 2
  double* fdm_invL = nullptr;  // (E, 4): ax, ay, az, scale
  semb::DMat fdm_Sh;           // host copies for the kernel parameter block
  double fdm_lam[semb::kMaxNq];

  semb::Profile prof;
  int64_t launch_count = 0;
};
