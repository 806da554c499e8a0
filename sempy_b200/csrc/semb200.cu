// libsemb200: C ABI (include/semb200.h) over the sm_100a kernels.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <utility>

#include "aux_kernels.cuh"
#include "ax_common.cuh"
#include "common.cuh"
#include "fdm_kernels.cuh"
#include "setup_kernels.cuh"
#include "vec_kernels.cuh"

namespace semb {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// Tuning knobs for A/B measurements (environment at load time, or semb_set_tuning at run time).
// None of them changes results (gs_pull: the same sums, a + b commutes bitwise).
static int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v ? atoi(v) : fallback;
}
// "p2p": 0 = NCCL for the halo exchange and the scalar all-reduces instead of the peer-memory path
static int g_p2p_on = env_int("SEMB_P2P", 1);
// "ax_l2_prefetch": 0 = no bulk L2 prefetch of G in the tuned Ax kernels
static int g_ax_l2_prefetch = env_int("SEMB_AX_L2_PREFETCH", 1);
// "gs_pull": 0 = the whole gather-scatter as a separate in-place pass (push form) instead of summing the
// two-copy segments inside the consumer (CG update kernel)
static int g_gs_pull = env_int("SEMB_GS_PULL", 1);
// "pdl": 1 = the kernels of a CG iteration on one GPU are chained by programmatic dependent launch (common.cuh)
static int g_pdl = env_int("SEMB_PDL", 1);

const NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    // an already loaded libnccl.so.2 (torch's) wins; otherwise load whatever the system provides
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define SEMB_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
    SEMB_SYM(GetUniqueId, "ncclGetUniqueId");
    SEMB_SYM(CommInitRank, "ncclCommInitRank");
    SEMB_SYM(CommDestroy, "ncclCommDestroy");
    SEMB_SYM(GroupStart, "ncclGroupStart");
    SEMB_SYM(GroupEnd, "ncclGroupEnd");
    SEMB_SYM(Send, "ncclSend");
    SEMB_SYM(Recv, "ncclRecv");
    SEMB_SYM(AllReduce, "ncclAllReduce");
    SEMB_SYM(GetErrorString, "ncclGetErrorString");
#undef SEMB_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv &&
           a.AllReduce && a.GetErrorString;
    return a;
  }();
  return api;
}

static inline int vec_grid(int64_t n) {
  int64_t g = (n + VEC_NT - 1) / VEC_NT;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Brackets a launch with CUDA events when profiling is on.
struct ProfScope {
  semb_ctx* c;
  int kind;
  cudaEvent_t stop = nullptr;
  cudaStream_t st;
  ProfScope(semb_ctx* ctx, int k, cudaStream_t stream = nullptr) : c(ctx), kind(k), st(stream) {
    if (c == nullptr) return;  // bracket switched off by the caller
    if (st == nullptr) st = c->stream;
    if (!c->prof.on || !(c->prof.mask & (1 << k))) return;
    Profile& p = c->prof;
    if (p.used[kind] == p.start[kind].size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      p.start[kind].push_back(a);
      p.stop[kind].push_back(b);
    }
    cudaEventRecord(p.start[kind][p.used[kind]], st);
    stop = p.stop[kind][p.used[kind]];
    p.used[kind]++;
    p.launches[kind]++;
  }
  ~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

static int ensure_vec(semb_ctx* c, int which) {
  if (c->vec[which] == nullptr) SEMB_CUDA(cudaMalloc(&c->vec[which], sizeof(double) * (size_t)std::max<int64_t>(c->n, 2)));
  return 0;
}

static int bind_device(semb_ctx* c) {
  SEMB_CUDA(cudaSetDevice(c->device));
  return 0;
}

// 128-bit accesses of the vector kernels need 16-byte aligned device vectors (a contiguous torch view
// with an odd element offset is not): refuse instead of faulting inside a kernel.
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static void drop_jacobi(semb_ctx* c) {
  if (c->vec[4]) {  // cached Jacobi diagonal is stale
    cudaFree(c->vec[4]);
    c->vec[4] = nullptr;
  }
  c->fdm_ready = false;
}

// ---- K1 dispatch -----------------------------------------------------------
// One launch over the elements [e0, e1).  dot: the launch is part of a logical grid of ncta_total CTAs
// whose fixed-order sum of p.Ap ends up in s->pAp (reduce.cuh); dir: 1 / 2 fold the CG direction update
// p <- r + beta p / p <- dinv r + beta p into the load phase.
// p_out (dir != 0): where the updated direction goes (nullptr: in place)
static int launch_ax(semb_ctx* c, const double* p, double* ap, bool mask, bool dot, int dir, int64_t e0, int64_t e1,
                     const double* r, const double* dinv, int cta_off, int ncta_total, const int* done,
                     cudaStream_t st = nullptr, bool prof = true, double* p_out = nullptr, bool pdl = false) {
  if (e1 <= e0) return 0;
  if (st == nullptr) st = c->stream;
  ProfScope ps(prof ? c : nullptr, SEMB_K_AX, st);
  c->launch_count++;
  AxArgs a;
  a.u = p;
  a.p_out = p_out != nullptr ? p_out : const_cast<double*>(p);
  a.r = r;
  a.dinv = dinv;
  a.G6 = c->G6;
  a.w = ap;
  a.maskbits = c->maskbits;
  a.s = c->scal;
  a.done = done;
  a.red = GridRed{c->partials, c->partials2, c->ticket};
  a.cta_off = cta_off;
  a.ncta_total = ncta_total;
  a.e0 = e0;
  a.e1 = e1;
  a.l2_prefetch = g_ax_l2_prefetch;
  a.pdl = pdl ? 1 : 0;
  if (c->Nq == 8) return launch_ax8(a, c->D, mask, dot, dir, st, c->device);
  if (c->Nq >= 3) return launch_axl(c->Nq, a, c->D, mask, dot, dir, st, c->device);
  return launch_ax_generic(c->Nq, a, c->D, mask, dot, dir, st);
}

// ---- K2 dispatch -----------------------------------------------------------
// in-place (push) form
// pull: the two-copy segments are left to the consumer and the longer ones only have their sums formed
// (compact form, c->gs_sums), nothing is written back to x (pull form, gs_linked)
// pass_n > 0: the pass_n partial sums of p.Ap in c->partials are added up (fixed order) into *pass_out by a
// passenger CTA of the gather-scatter launch, or by a launch of their own when there is nothing to gather
static int launch_gs(semb_ctx* c, double* x, const int* done, bool pull = false, int pass_n = 0, double* pass_out = nullptr,
                     bool pdl = false) {
  const bool skip_pairs = pull;
  double* sums = pull ? c->gs_sums : nullptr;
  ProfScope ps(c, SEMB_K_GS);
  GsPlan plan{0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const GsClass& k : c->gs_classes) {
    if (k.len == 2) { if (!skip_pairs) { plan.n2 = k.nseg; plan.off2 = k.off; } }
    else if (k.len == 4) { plan.n4 = k.nseg; plan.off4 = k.off; }
    else if (k.len == 8) { plan.n8 = k.nseg; plan.off8 = k.off; }
  }
  // blocks per class in proportion to their loads+stores; each thread takes 4 / 2 / 1 segments per trip
  auto blocks_for = [](int64_t nseg, int per_thread) {
    if (nseg == 0) return 0;
    const int64_t want = (nseg + (int64_t)VEC_NT * per_thread - 1) / ((int64_t)VEC_NT * per_thread);
    return (int)std::min<int64_t>(want, (int64_t)kNumSMs * 8);
  };
  plan.b2 = blocks_for(plan.n2, GS_U2);
  plan.b4 = blocks_for(plan.n4, GS_U4);
  plan.b8 = blocks_for(plan.n8, GS_U8);
  const int total_blocks = plan.b2 + plan.b4 + plan.b8;
  if (total_blocks > 0) {
    c->launch_count++;
    GsPassenger pass{pass_n > 0 ? c->partials : nullptr, pass_n, pass_out,
                     GridRed{c->pass_part, c->pass_part + 32, c->pass_tick}};
    SEMB_CUDA(launch_maybe_pdl(gs_fixed_kernel, dim3(total_blocks + (pass_n > 0 ? GS_PASSENGERS : 0)), dim3(VEC_NT), 0, c->stream, pdl,
                               x, c->gs_ids, plan, done, sums, pass));
  } else if (pass_n > 0) {
    c->launch_count++;
    reduce_partials_kernel<<<1, RED_NT, 0, c->stream>>>(c->partials, pass_n, 0, 1, pass_out, done);
    SEMB_CUDA(cudaGetLastError());
  }
  if (c->gs_var_nseg > 0) {
    c->launch_count++;
    gs_var_kernel<<<vec_grid(c->gs_var_nseg), VEC_NT, 0, c->stream>>>(x, c->gs_ids + c->gs_var_off, c->gs_var_start,
                                                                      c->gs_var_nseg, done,
                                                                      sums != nullptr ? sums + plan.n4 + plan.n8 : nullptr);
    SEMB_CUDA(cudaGetLastError());
  }
  return 0;
}
static inline bool use_pull(const semb_ctx* c) { return c->pull_ok && g_gs_pull != 0; }

// ---- multi-GPU halo --------------------------------------------------------
static inline bool use_p2p(const semb_ctx* c) { return c->p2p && g_p2p_on; }

static HaloPeers make_halo_peers(const semb_ctx* c, const std::vector<int64_t>& peer_off) {
  HaloPeers hp;
  std::memset(&hp, 0, sizeof(hp));
  hp.nneigh = (int)c->nb_rank.size();
  for (int q = 0; q < hp.nneigh; ++q) {
    hp.rank_of[q] = c->nb_rank[q];
    hp.off[q] = c->nb_off[q];
    hp.peer_off[q] = peer_off[q];
  }
  hp.off[hp.nneigh] = c->n_send;
  return hp;
}

static LlArgs ll_args(semb_ctx* c, unsigned int seq) {
  return LlArgs{c->ll_peer_dev, c->rank, c->nranks, seq, c->ll_error};
}

// gather this rank's partial sums; with the peer-memory path this also delivers them (seq = the
// exchange's sequence number, advanced by the caller once per exchange)
static int halo_gather(semb_ctx* c, const double* x, cudaStream_t st, const int* done) {
  if (c->n_send == 0) return 0;
  c->launch_count++;
  if (use_p2p(c)) {
    halo_gather_p2p_kernel<<<vec_grid(c->n_send), VEC_NT, 0, st>>>(x, c->send_start, c->send_ids, c->n_send, c->hbuf,
                                                                   c->ll_peer_dev, c->halo_peers, c->ll_halo_cap,
                                                                   c->seq_halo, done);
  } else {
    halo_gather_kernel<<<vec_grid(c->n_send), VEC_NT, 0, st>>>(x, c->send_start, c->send_ids, c->n_send, c->hbuf, done);
  }
  SEMB_CUDA(cudaGetLastError());
  return 0;
}
static int halo_exchange(semb_ctx* c, cudaStream_t st) {
  if (c->nb_rank.empty() || use_p2p(c)) return 0;
  SEMB_CHECK(c->comm != nullptr, "no NCCL communicator (semb_comm_init was called without an id) and the peer-memory "
                                 "transport is not attached or switched off");
  SEMB_NCCL(GroupStart());
  for (size_t q = 0; q < c->nb_rank.size(); ++q) {
    const int64_t off = c->nb_off[q], cnt = c->nb_off[q + 1] - off;
    SEMB_NCCL(Send(c->hbuf + off, (size_t)cnt, ncclDouble, c->nb_rank[q], c->comm, st));
    SEMB_NCCL(Recv(c->hbuf + c->n_send + off, (size_t)cnt, ncclDouble, c->nb_rank[q], c->comm, st));
  }
  SEMB_NCCL(GroupEnd());
  return 0;
}
// pass_vals (peer-memory path only): scalars all-reduced by a passenger CTA of the same launch
static int halo_scatter(semb_ctx* c, double* x, cudaStream_t st, const int* done, double* pass_vals = nullptr,
                        int pass_count = 0) {
  if (use_p2p(c)) {
    if (c->n_if == 0 && pass_vals == nullptr) return 0;
    unsigned int pseq = 0;
    if (pass_vals != nullptr) pseq = ++c->seq_scalar;
    const int blocks = (c->n_if > 0 ? vec_grid(c->n_if) : 0) + (pass_vals != nullptr ? 1 : 0);
    c->launch_count++;
    halo_scatter_p2p_kernel<<<blocks, VEC_NT, 0, st>>>(x, c->hbuf, c->ll_local + LL_SCALAR_WORDS, c->n_send, c->src_start,
                                                       c->src_off, c->copy_start, c->copy_ids, c->n_if, c->ll_halo_cap,
                                                       c->seq_halo, done, c->ll_error, pass_vals, pass_count,
                                                       ll_args(c, pseq));
  } else {
    if (c->n_if == 0) return 0;
    c->launch_count++;
    halo_scatter_kernel<<<vec_grid(c->n_if), VEC_NT, 0, st>>>(x, c->hbuf, c->src_start, c->src_off, c->copy_start,
                                                             c->copy_ids, c->n_if, done);
  }
  SEMB_CUDA(cudaGetLastError());
  return 0;
}
// Complete in-place gather-scatter of one vector on the context's stream (no overlap).
static int gs_full(semb_ctx* c, double* x, const int* done) {
  if (c->have_halo) {
    c->seq_halo++;
    SEMB_TRY(halo_gather(c, x, c->stream, done));
    SEMB_TRY(halo_exchange(c, c->stream));
  }
  SEMB_TRY(launch_gs(c, x, done));
  if (c->have_halo) SEMB_TRY(halo_scatter(c, x, c->stream, done));
  return 0;
}
static int allreduce_sum(semb_ctx* c, double* dev, int count) {
  if (c->nranks == 1) return 0;
  if (use_p2p(c) && count <= LL_MAXV) {
    c->seq_scalar++;
    c->launch_count++;
    ll_allreduce_kernel<<<1, LL_MAXR * LL_MAXV, 0, c->stream>>>(dev, count, c->ll_peer_dev, c->rank, c->nranks,
                                                               c->seq_scalar, c->ll_error);
    SEMB_CUDA(cudaGetLastError());
    return 0;
  }
  SEMB_CHECK(c->comm != nullptr, "no NCCL communicator (semb_comm_init was called without an id) and the peer-memory "
                                 "transport is not attached or switched off");
  SEMB_NCCL(AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, c->comm, c->stream));
  return 0;
}

// after a stream synchronisation: did a peer-memory wait time out?
static int check_ll_error(semb_ctx* c) {
  if (!c->p2p || c->ll_error == nullptr) return 0;
  int flag = 0;
  SEMB_CUDA(cudaMemcpy(&flag, c->ll_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    set_error("peer-memory exchange timed out: a neighbouring rank did not deliver (rank failure or table mismatch)");
    return 4;
  }
  return 0;
}

static int reduce_partials(semb_ctx* c, int64_t n, int64_t stride, int nout, double* out, const int* done) {
  c->launch_count++;
  reduce_partials_kernel<<<1, RED_NT, 0, c->stream>>>(c->partials, n, stride, nout, out, done);
  SEMB_CUDA(cudaGetLastError());
  return 0;
}

// partial sums, group sums and tickets of the grid reductions (reduce.cuh); the CG update kernel's own
// ticket is the last entry
static int ensure_partials(semb_ctx* c) {
  const int64_t need = std::max<int64_t>(c->Nq >= 2 ? c->E : 0, 2 * (int64_t)kNumSMs * 8) + 16;
  if (c->partials_cap < need) {
    cudaFree(c->partials);
    cudaFree(c->partials2);
    cudaFree(c->ticket);
    const int64_t groups = need / RED_GROUP + 2;
    SEMB_CUDA(cudaMalloc(&c->partials, sizeof(double) * (size_t)need));
    SEMB_CUDA(cudaMalloc(&c->partials2, sizeof(double) * (size_t)groups));
    SEMB_CUDA(cudaMalloc(&c->ticket, sizeof(unsigned int) * (size_t)(groups + 2)));
    SEMB_CUDA(cudaMemset(c->ticket, 0, sizeof(unsigned int) * (size_t)(groups + 2)));
    c->partials_cap = need;
    c->ticket_update = c->ticket + groups + 1;
    if (c->pass_part == nullptr) {
      SEMB_CUDA(cudaMalloc(&c->pass_part, sizeof(double) * 64));
      SEMB_CUDA(cudaMalloc(&c->pass_tick, sizeof(unsigned int) * 8));
      SEMB_CUDA(cudaMemset(c->pass_tick, 0, sizeof(unsigned int) * 8));
    }
  }
  return 0;
}

// ---- host <-> device copies of caller vectors -------------------------------------------------
static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
static int hs_init(semb_ctx* c) {
  HostStager& h = c->hs;
  if (h.ready) return 0;
  for (int t = 0; t < HS_THREADS; ++t) {
    SEMB_CUDA(cudaStreamCreateWithFlags(&h.stream[t], cudaStreamNonBlocking));
    for (int q = 0; q < 2; ++q) {
      SEMB_CUDA(cudaMallocHost(&h.slot[t][q], HS_CHUNK));
      SEMB_CUDA(cudaEventCreateWithFlags(&h.ev[t][q], cudaEventDisableTiming));
    }
  }
  SEMB_CUDA(cudaEventCreateWithFlags(&h.fence, cudaEventDisableTiming));
  h.ready = true;
  return 0;
}
static void hs_free(semb_ctx* c) {
  HostStager& h = c->hs;
  if (!h.ready) return;
  for (int t = 0; t < HS_THREADS; ++t) {
    for (int q = 0; q < 2; ++q) {
      cudaFreeHost(h.slot[t][q]);
      cudaEventDestroy(h.ev[t][q]);
    }
    cudaStreamDestroy(h.stream[t]);
  }
  cudaEventDestroy(h.fence);
  h.ready = false;
}
// Copy `bytes` between pageable host memory and the device, ordered after the work already queued on
// c->stream; returns when the data has arrived (host side: complete; device side: c->stream waits for it).
static int hs_copy(semb_ctx* c, void* dev, void* host, size_t bytes, bool to_device) {
  if (bytes < 2 * HS_CHUNK || host_is_pinned(host)) {
    if (to_device) SEMB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
    else {
      SEMB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
      SEMB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return 0;
  }
  SEMB_TRY(hs_init(c));
  HostStager& h = c->hs;
  SEMB_CUDA(cudaEventRecord(h.fence, c->stream));
  const size_t nchunk = (bytes + HS_CHUNK - 1) / HS_CHUNK;
  // one memcpy thread moves ~10 GB/s, PCIe ~55 GB/s: up to 8 threads, fewer when several ranks share the host
  const int NTH = std::max(2, std::min(HS_THREADS, (int)std::thread::hardware_concurrency() / std::max(1, c->nranks)));
  cudaError_t errs[HS_THREADS];
  std::thread workers[HS_THREADS];
  for (int t = 0; t < NTH; ++t) {
    errs[t] = cudaSuccess;
    workers[t] = std::thread([&, t]() {
      cudaError_t e = cudaSetDevice(c->device);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(h.stream[t], h.fence, 0);
      auto span = [&](size_t ch, size_t& off, size_t& len) {
        off = ch * HS_CHUNK;
        len = std::min(HS_CHUNK, bytes - off);
      };
      size_t off, len;
      if (to_device) {
        int use = 0;
        for (size_t ch = t; ch < nchunk && e == cudaSuccess; ch += NTH, ++use) {
          const int q = use & 1;
          if (use >= 2) e = cudaEventSynchronize(h.ev[t][q]);  // the DMA out of this slot has finished
          if (e != cudaSuccess) break;
          span(ch, off, len);
          std::memcpy(h.slot[t][q], static_cast<char*>(host) + off, len);
          e = cudaMemcpyAsync(static_cast<char*>(dev) + off, h.slot[t][q], len, cudaMemcpyHostToDevice, h.stream[t]);
          if (e == cudaSuccess) e = cudaEventRecord(h.ev[t][q], h.stream[t]);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h.stream[t]);
      } else {
        // DMA of chunk k+1 into the other slot runs while chunk k is copied out
        size_t ch = t;
        int use = 0;
        if (ch < nchunk) {
          span(ch, off, len);
          e = cudaMemcpyAsync(h.slot[t][0], static_cast<char*>(dev) + off, len, cudaMemcpyDeviceToHost, h.stream[t]);
          if (e == cudaSuccess) e = cudaEventRecord(h.ev[t][0], h.stream[t]);
        }
        for (; ch < nchunk && e == cudaSuccess; ch += NTH, ++use) {
          const int q = use & 1;
          const size_t nxt = ch + NTH;
          if (nxt < nchunk) {
            size_t o2, l2;
            span(nxt, o2, l2);
            e = cudaMemcpyAsync(h.slot[t][q ^ 1], static_cast<char*>(dev) + o2, l2, cudaMemcpyDeviceToHost, h.stream[t]);
            if (e == cudaSuccess) e = cudaEventRecord(h.ev[t][q ^ 1], h.stream[t]);
            if (e != cudaSuccess) break;
          }
          e = cudaEventSynchronize(h.ev[t][q]);
          if (e != cudaSuccess) break;
          span(ch, off, len);
          std::memcpy(static_cast<char*>(host) + off, h.slot[t][q], len);
        }
      }
      errs[t] = e;
    });
  }
  for (int t = 0; t < NTH; ++t) workers[t].join();
  for (int t = 0; t < NTH; ++t) SEMB_CUDA(errs[t]);
  return 0;
}

// Host <-> device staging of caller vectors through two pooled buffers of the context (allocated on
// first use, kept): a reference-style loop that calls mesh.dssum(x) with NumPy arrays pays two copies per
// call, not an allocation.
struct Staged {
  double* dev = nullptr;
};
static int stage_in(semb_ctx* c, const double* src, int mem, int64_t n, Staged& s, bool copy, int slot) {
  if (mem == SEMB_DEVICE) {
    SEMB_CHECK(aligned16(src), "device vectors must be 16-byte aligned (a tensor view with an odd element offset is not)");
    s.dev = const_cast<double*>(src);
    return 0;
  }
  if (c->stage[slot] == nullptr) SEMB_CUDA(cudaMalloc(&c->stage[slot], sizeof(double) * (size_t)std::max<int64_t>(c->n, 2)));
  s.dev = c->stage[slot];
  if (copy) SEMB_TRY(hs_copy(c, s.dev, const_cast<double*>(src), sizeof(double) * (size_t)n, true));
  return 0;
}
static int stage_out(semb_ctx* c, double* dst, int mem, int64_t n, Staged& s) {
  if (mem == SEMB_DEVICE) return 0;
  SEMB_TRY(hs_copy(c, s.dev, dst, sizeof(double) * (size_t)n, false));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return check_ll_error(c);
}

// rmult = 1 / gs(ones)   (sempy/mesh.py:440-442); across ranks when a halo is set
static int build_rmult(semb_ctx* c) {
  if (!c->rmult) SEMB_CUDA(cudaMalloc(&c->rmult, sizeof(double) * (size_t)c->n));
  if (!c->mult) SEMB_CUDA(cudaMalloc(&c->mult, (size_t)c->n + 16));
  SEMB_CUDA(cudaMemsetAsync(c->info, 0, 4 * sizeof(int), c->stream));
  fill_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(c->rmult, 1.0, c->n);
  SEMB_TRY(gs_full(c, c->rmult, nullptr));
  reciprocal_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(c->rmult, c->mult, c->n, c->info);
  SEMB_CUDA(cudaGetLastError());
  int info[4] = {0, 0, 0, 0};
  SEMB_CUDA(cudaMemcpyAsync(info, c->info, sizeof(info), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  SEMB_TRY(check_ll_error(c));
  SEMB_CHECK(info[0] == 0, "gather-scatter: a node has more than 255 copies (the CG update kernel stores one byte per node)");
  drop_jacobi(c);
  return 0;
}

static int build_jacobi(semb_ctx* c) {
  SEMB_TRY(ensure_vec(c, 4));
  double* d = c->vec[4];
  c->launch_count++;
  diag_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(c->G6, d, c->E, c->Nq, c->D);
  SEMB_CUDA(cudaGetLastError());
  SEMB_TRY(gs_full(c, d, nullptr));
  c->launch_count++;
  jacobi_invert_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(d, c->have_mask ? c->maskbits : nullptr, c->n);
  SEMB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
static int upload(T** dst, const T* src, size_t count) {
  cudaFree(*dst);
  *dst = nullptr;
  SEMB_CUDA(cudaMalloc(dst, sizeof(T) * std::max<size_t>(count, 1)));
  if (count) SEMB_CUDA(cudaMemcpy(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice));
  // a pageable-memory cudaMemcpy may return before its DMA has landed, and the context's stream is
  // non-blocking (no implicit ordering with the default stream): make the table resident here
  SEMB_CUDA(cudaDeviceSynchronize());
  return 0;
}

// scoped device allocation for set-up temporaries
struct DevTmp {
  void* p = nullptr;
  ~DevTmp() { cudaFree(p); }
  int alloc(size_t bytes) {
    SEMB_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return 0;
  }
  template <typename T>
  T* as() { return static_cast<T*>(p); }
};
// HOST arrays are uploaded into a temporary, DEVICE arrays used in place
template <typename T>
static int dev_view(const T* src, int mem, size_t count, DevTmp& tmp, const T** out) {
  if (mem == SEMB_DEVICE) {
    *out = src;
    return 0;
  }
  SEMB_TRY(tmp.alloc(sizeof(T) * count));
  SEMB_CUDA(cudaMemcpy(tmp.p, src, sizeof(T) * count, cudaMemcpyHostToDevice));
  *out = tmp.as<T>();
  return 0;
}

}  // namespace semb

using namespace semb;

extern "C" {

int semb_version(void) { return 200; }
const char* semb_last_error(void) { return g_last_error.c_str(); }

int semb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

int semb_create(semb_ctx** out, int device, int64_t num_elems, int Nq) {
  SEMB_CHECK(out != nullptr, "semb_create: out is NULL");
  SEMB_CHECK(num_elems > 0, "semb_create: num_elems must be positive");
  SEMB_CHECK(Nq >= 2 && Nq <= kMaxNq, "semb_create: Nq must be in [2,16]");
  SEMB_CHECK((int64_t)num_elems * Nq * Nq * Nq < (int64_t)2147483647, "semb_create: more than 2^31 local nodes per GPU");
  SEMB_CUDA(cudaSetDevice(device));
  semb_ctx* c = new semb_ctx();
  c->device = device;
  c->E = num_elems;
  c->Nq = Nq;
  c->Np = (int64_t)Nq * Nq * Nq;
  c->n = c->E * c->Np;
  SEMB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  SEMB_CUDA(cudaMalloc(&c->scal, sizeof(CgScalars) + 8 * sizeof(double)));
  SEMB_CUDA(cudaMemset(c->scal, 0, sizeof(CgScalars) + 8 * sizeof(double)));
  SEMB_CUDA(cudaMallocHost(&c->h_pinned, 64));
  SEMB_CUDA(cudaMalloc(&c->info, 4 * sizeof(int)));
  SEMB_CUDA(cudaMemset(c->info, 0, 4 * sizeof(int)));
  if (ensure_partials(c)) return 1;
  SEMB_CUDA(cudaDeviceSynchronize());  // the memsets above ran on the default stream
  *out = c;
  return 0;
}

// gslib-style handle (GS(comm), sempy/mesh.py:437): a context that only gather-scatters vectors of n entries
int semb_gs_create(semb_ctx** out, int device, int64_t n) {
  SEMB_CHECK(out != nullptr, "semb_gs_create: out is NULL");
  SEMB_CHECK(n > 0 && n < (int64_t)2147483647, "semb_gs_create: n must be in [1, 2^31)");
  SEMB_CUDA(cudaSetDevice(device));
  semb_ctx* c = new semb_ctx();
  c->device = device;
  c->E = n;
  c->Nq = 1;
  c->Np = 1;
  c->n = n;
  SEMB_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  SEMB_CUDA(cudaMalloc(&c->scal, sizeof(CgScalars) + 8 * sizeof(double)));
  SEMB_CUDA(cudaMemset(c->scal, 0, sizeof(CgScalars) + 8 * sizeof(double)));
  SEMB_CUDA(cudaMallocHost(&c->h_pinned, 64));
  SEMB_CUDA(cudaMalloc(&c->info, 4 * sizeof(int)));
  SEMB_CUDA(cudaMemset(c->info, 0, 4 * sizeof(int)));
  if (ensure_partials(c)) return 1;
  SEMB_CUDA(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

int semb_destroy(semb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaFree(c->G6);
  cudaFree(c->gs_ids);
  cudaFree(c->gs_link);
  cudaFree(c->gs_sums);
  cudaFree(c->gs_var_start);
  cudaFree(c->rmult);
  cudaFree(c->mult);
  cudaFree(c->maskbits);
  cudaFree(c->partials);
  cudaFree(c->partials2);
  cudaFree(c->pass_part);
  cudaFree(c->pass_tick);
  cudaFree(c->scal);
  cudaFree(c->ticket);
  cudaFree(c->info);
  cudaFree(c->hist);
  cudaFree(c->fdm_S);
  cudaFree(c->fdm_invL);
  for (double* v : c->vec) cudaFree(v);
  for (double* v : c->stage) cudaFree(v);
  hs_free(c);
  cudaFreeHost(c->h_pinned);
  for (int k = 0; k < SEMB_K_COUNT; ++k) {
    for (cudaEvent_t e : c->prof.start[k]) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof.stop[k]) cudaEventDestroy(e);
  }
  cudaFree(c->send_start);
  cudaFree(c->send_ids);
  cudaFree(c->src_start);
  cudaFree(c->src_off);
  cudaFree(c->copy_start);
  cudaFree(c->copy_ids);
  cudaFree(c->hbuf);
  for (int r = 0; r < (int)c->ll_peer.size(); ++r)
    if (r != c->rank && c->ll_peer[r]) cudaIpcCloseMemHandle(c->ll_peer[r]);
  cudaFree(c->ll_local);
  cudaFree(c->ll_peer_dev);
  cudaFree(c->ll_error);
  if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->ev_if_done) cudaEventDestroy(c->ev_if_done);
  if (c->ev_gathered) cudaEventDestroy(c->ev_gathered);
  if (c->ev_exchanged) cudaEventDestroy(c->ev_exchanged);
  cudaStreamDestroy(c->own_stream);
  delete c;
  return 0;
}

int semb_set_stream(semb_ctx* c, void* s) {
  SEMB_CHECK(c != nullptr, "semb_set_stream: ctx is NULL");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return 0;
}

int semb_synchronize(semb_ctx* c) {
  SEMB_CHECK(c != nullptr, "semb_synchronize: ctx is NULL");
  SEMB_TRY(bind_device(c));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return check_ll_error(c);
}

int semb_set_derivative(semb_ctx* c, const double* D) {
  SEMB_CHECK(c && D, "semb_set_derivative: NULL argument");
  std::memset(c->D.d, 0, sizeof(c->D.d));
  std::memcpy(c->D.d, D, sizeof(double) * c->Nq * c->Nq);
  c->have_D = true;
  drop_jacobi(c);
  return 0;
}

// (E,3,3,Np) -> (E,6,Np) on the device
__global__ void g9_to_g6_kernel(const double* __restrict__ g9, double* __restrict__ g6, int64_t E, int64_t Np) {
  const int64_t n = E * 6 * Np;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = q / (6 * Np);
    const int c = (int)((q / Np) % 6);
    const int64_t node = q % Np;
    const int src = c == 0 ? 0 : c == 1 ? 1 : c == 2 ? 2 : c == 3 ? 4 : c == 4 ? 5 : 8;
    g6[q] = g9[(e * 9 + src) * Np + node];
  }
}

int semb_set_geom(semb_ctx* c, const double* G, int layout, int mem) {
  SEMB_CHECK(c && G, "semb_set_geom: NULL argument");
  SEMB_CHECK(layout == SEMB_G9 || layout == SEMB_G6, "semb_set_geom: layout must be SEMB_G9 or SEMB_G6");
  SEMB_TRY(bind_device(c));
  const size_t bytes6 = sizeof(double) * 6 * (size_t)c->n;
  if (!c->G6) SEMB_CUDA(cudaMalloc(&c->G6, bytes6));
  if (layout == SEMB_G6) {
    SEMB_CUDA(cudaMemcpyAsync(c->G6, G, bytes6, mem == SEMB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                              c->stream));
  } else {
    // stream the 9-component array through a bounded device buffer, element chunks at a time
    const int64_t chunkE = std::max<int64_t>(1, std::min<int64_t>(c->E, (int64_t)(256u << 20) / (9 * 8 * c->Np)));
    double* tmp = nullptr;
    if (mem == SEMB_HOST) SEMB_CUDA(cudaMalloc(&tmp, sizeof(double) * 9 * (size_t)(chunkE * c->Np)));
    for (int64_t e0 = 0; e0 < c->E; e0 += chunkE) {
      const int64_t ne = std::min(chunkE, c->E - e0);
      const double* src = G + e0 * 9 * c->Np;
      if (mem == SEMB_HOST) {
        SEMB_CUDA(cudaMemcpyAsync(tmp, src, sizeof(double) * 9 * (size_t)(ne * c->Np), cudaMemcpyHostToDevice, c->stream));
        src = tmp;
      }
      g9_to_g6_kernel<<<vec_grid(ne * 6 * c->Np), VEC_NT, 0, c->stream>>>(src, c->G6 + e0 * 6 * c->Np, ne, c->Np);
      SEMB_CUDA(cudaGetLastError());
    }
    SEMB_CUDA(cudaStreamSynchronize(c->stream));
    if (tmp) cudaFree(tmp);
  }
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  c->have_G = true;
  drop_jacobi(c);
  return 0;
}

int semb_setup_geom(semb_ctx* c, const double* w, const double* xe, const double* ye, const double* ze, double* jaco,
                    int mem) {
  SEMB_CHECK(c && w && xe && ye && ze, "semb_setup_geom: NULL argument");
  SEMB_CHECK(c->have_D, "semb_setup_geom: call semb_set_derivative first");
  SEMB_TRY(bind_device(c));
  DMat wm;
  std::memset(wm.d, 0, sizeof(wm.d));
  std::memcpy(wm.d, w, sizeof(double) * c->Nq);
  const size_t cnt = (size_t)c->n;
  DevTmp tx, ty, tz, tj;
  const double *dx, *dy, *dz;
  SEMB_TRY(dev_view(xe, mem, cnt, tx, &dx));
  SEMB_TRY(dev_view(ye, mem, cnt, ty, &dy));
  SEMB_TRY(dev_view(ze, mem, cnt, tz, &dz));
  double* dj = jaco;
  if (jaco != nullptr && mem == SEMB_HOST) {
    SEMB_TRY(tj.alloc(sizeof(double) * cnt));
    dj = tj.as<double>();
  }
  if (!c->G6) SEMB_CUDA(cudaMalloc(&c->G6, sizeof(double) * 6 * cnt));
  const int Np = (int)c->Np;
  SEMB_CUDA(cudaFuncSetAttribute(geom_factors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 3 * 4096)));
  geom_factors_kernel<<<(unsigned)c->E, std::min(256, ((Np + 31) / 32) * 32), sizeof(double) * 3 * Np, c->stream>>>(
      dx, dy, dz, c->G6, dj, c->Nq, c->D, wm);
  SEMB_CUDA(cudaGetLastError());
  if (jaco != nullptr && mem == SEMB_HOST)
    SEMB_CUDA(cudaMemcpyAsync(jaco, dj, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c->stream));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  c->have_G = true;
  drop_jacobi(c);
  return 0;
}

int semb_get_geom(semb_ctx* c, double* G6, int mem) {
  SEMB_CHECK(c && G6, "semb_get_geom: NULL argument");
  SEMB_CHECK(c->have_G, "semb_get_geom: geometric factors are not set");
  SEMB_TRY(bind_device(c));
  SEMB_CUDA(cudaMemcpyAsync(G6, c->G6, sizeof(double) * 6 * (size_t)c->n,
                            mem == SEMB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int semb_set_gather_scatter(semb_ctx* c, const int64_t* g2l, const int64_t* gstart, int64_t nglobal, int mem) {
  SEMB_CHECK(c && g2l && gstart, "semb_set_gather_scatter: NULL argument");
  SEMB_CHECK(nglobal > 0 && nglobal <= c->n, "semb_set_gather_scatter: nglobal out of range");
  SEMB_TRY(bind_device(c));
  const int64_t n = c->n, ns = nglobal;
  DevTmp t_g2l, t_gst;
  const int64_t *d_g2l, *d_gst;
  SEMB_TRY(dev_view(g2l, mem, (size_t)n, t_g2l, &d_g2l));
  SEMB_TRY(dev_view(gstart, mem, (size_t)ns + 1, t_gst, &d_gst));
  int64_t ends[2] = {-1, -1};
  SEMB_CUDA(cudaMemcpy(&ends[0], d_gst, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SEMB_CUDA(cudaMemcpy(&ends[1], d_gst + ns, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SEMB_CHECK(ends[0] == 0 && ends[1] == n, "semb_set_gather_scatter: global_start must run from 0 to E*Np");

  // 1. one sort key per segment: (length class, lowest local index)
  DevTmp k0, k1, v0, v1, info, bounds, cubtmp, vlen, vstart;
  SEMB_TRY(k0.alloc(8 * (size_t)ns));
  SEMB_TRY(k1.alloc(8 * (size_t)ns));
  SEMB_TRY(v0.alloc(4 * (size_t)ns));
  SEMB_TRY(v1.alloc(4 * (size_t)ns));
  SEMB_TRY(info.alloc(4 * sizeof(int)));
  SEMB_TRY(bounds.alloc(5 * sizeof(int64_t)));
  SEMB_CUDA(cudaMemset(info.p, 0, 4 * sizeof(int)));
  seg_keys_kernel<<<vec_grid(ns), VEC_NT>>>(d_g2l, d_gst, ns, n, k0.as<unsigned long long>(), v0.as<int32_t>(), info.as<int>());
  SEMB_CUDA(cudaGetLastError());
  size_t tb = 0;
  SEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.as<unsigned long long>(), k1.as<unsigned long long>(),
                                            v0.as<int32_t>(), v1.as<int32_t>(), (int)ns, 0, 43));
  SEMB_TRY(cubtmp.alloc(tb));
  SEMB_CUDA(cub::DeviceRadixSort::SortPairs(cubtmp.p, tb, k0.as<unsigned long long>(), k1.as<unsigned long long>(),
                                            v0.as<int32_t>(), v1.as<int32_t>(), (int)ns, 0, 43));
  seg_bounds_kernel<<<1, 32>>>(k1.as<unsigned long long>(), ns, bounds.as<int64_t>());
  SEMB_CUDA(cudaGetLastError());
  SegLayout lay;
  int h_info[4];
  SEMB_CUDA(cudaMemcpy(lay.b, bounds.p, sizeof(lay.b), cudaMemcpyDeviceToHost));
  SEMB_CUDA(cudaMemcpy(h_info, info.p, sizeof(h_info), cudaMemcpyDeviceToHost));
  SEMB_CHECK((h_info[0] & 1) == 0, "semb_set_gather_scatter: empty segment or offsets out of range");
  SEMB_CHECK((h_info[0] & 2) == 0, "semb_set_gather_scatter: local index out of range");
  const int32_t* sorted = v1.as<int32_t>();

  // 2. layout of the id table: [len 2 | len 4 | len 8 | other], every class starting at a multiple of 4
  const int64_t n2 = lay.b[1] - lay.b[0], n4 = lay.b[2] - lay.b[1], n8 = lay.b[3] - lay.b[2], nv = lay.b[4] - lay.b[3];
  auto align4 = [](int64_t v) { return (v + 3) & ~(int64_t)3; };
  lay.off[0] = 0;
  lay.off[1] = align4(lay.off[0] + 2 * n2);
  lay.off[2] = align4(lay.off[1] + 4 * n4);
  lay.off[3] = align4(lay.off[2] + 8 * n8);
  int64_t nvar_ids = 0;
  cudaFree(c->gs_var_start);
  c->gs_var_start = nullptr;
  SEMB_CUDA(cudaMalloc(&c->gs_var_start, sizeof(int32_t) * (size_t)(nv + 1)));
  if (nv > 0) {
    SEMB_TRY(vlen.alloc(sizeof(int32_t) * (size_t)(nv + 1)));
    SEMB_CUDA(cudaMemset(vlen.p, 0, sizeof(int32_t) * (size_t)(nv + 1)));
    seg_var_len_kernel<<<vec_grid(nv), VEC_NT>>>(sorted, d_gst, lay.b[3], lay.b[4], vlen.as<int32_t>());
    SEMB_CUDA(cudaGetLastError());
    size_t sb = 0;
    SEMB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, sb, vlen.as<int32_t>(), c->gs_var_start, (int)(nv + 1)));
    DevTmp scantmp;
    SEMB_TRY(scantmp.alloc(sb));
    SEMB_CUDA(cub::DeviceScan::ExclusiveSum(scantmp.p, sb, vlen.as<int32_t>(), c->gs_var_start, (int)(nv + 1)));
    int32_t last = 0;
    SEMB_CUDA(cudaMemcpy(&last, c->gs_var_start + nv, sizeof(int32_t), cudaMemcpyDeviceToHost));
    nvar_ids = last;
  } else {
    SEMB_CUDA(cudaMemset(c->gs_var_start, 0, sizeof(int32_t)));
  }
  const int64_t nids = align4(lay.off[3] + nvar_ids);
  SEMB_CHECK(nids < (int64_t)2147483647, "semb_set_gather_scatter: id table too long");

  // 3. ids (copies of a segment in ascending local index) and the links of the pull form
  cudaFree(c->gs_ids);
  cudaFree(c->gs_link);
  cudaFree(c->gs_sums);
  c->gs_ids = nullptr;
  c->gs_link = nullptr;
  c->gs_sums = nullptr;
  c->pull_ok = n4 + n8 + nv < ((int64_t)1 << 30);
  SEMB_CUDA(cudaMalloc(&c->gs_sums, sizeof(double) * (size_t)std::max<int64_t>(n4 + n8 + nv, 2)));
  SEMB_CUDA(cudaMalloc(&c->gs_ids, sizeof(int32_t) * (size_t)std::max<int64_t>(nids, 4)));
  SEMB_CUDA(cudaMemset(c->gs_ids, 0, sizeof(int32_t) * (size_t)std::max<int64_t>(nids, 4)));
  SEMB_CUDA(cudaMalloc(&c->gs_link, sizeof(uint32_t) * (size_t)(n + 2)));
  SEMB_CUDA(cudaMemset(c->gs_link, 0, sizeof(uint32_t) * (size_t)(n + 2)));
  if (lay.b[4] > 0) {
    seg_emit_kernel<<<vec_grid(lay.b[4]), VEC_NT>>>(sorted, d_g2l, d_gst, lay, c->gs_var_start, c->gs_ids, c->gs_link);
    SEMB_CUDA(cudaGetLastError());
  }
  SEMB_CUDA(cudaDeviceSynchronize());  // tables complete before the context's (non-blocking) stream uses them
  c->gs_classes.clear();
  c->gs_classes.push_back(GsClass{2, n2, lay.off[0]});
  c->gs_classes.push_back(GsClass{4, n4, lay.off[1]});
  c->gs_classes.push_back(GsClass{8, n8, lay.off[2]});
  c->gs_var_off = lay.off[3];
  c->gs_var_nseg = nv;
  c->gs_nids = nids;
  c->have_gs = true;
  return build_rmult(c);
}

// GS.setup(ids) of gslib as SemPy drives it (sempy/mesh.py:437-438): entries with the same positive id are
// copies of one node; id 0 marks an entry that takes no part.  The ids are grouped on the device (stable
// radix sort by id, segment starts where the id changes) and handed to semb_set_gather_scatter.
int semb_gs_setup(semb_ctx* c, const int32_t* ids, int mem) {
  SEMB_CHECK(c && ids, "semb_gs_setup: NULL argument");
  SEMB_TRY(bind_device(c));
  const int64_t n = c->n;
  DevTmp t_ids, k0, k1, i0, i1, fl, st, cnt, tmp, info;
  const int32_t* d_ids;
  SEMB_TRY(dev_view(ids, mem, (size_t)n, t_ids, &d_ids));
  SEMB_TRY(k0.alloc(4 * (size_t)n));
  SEMB_TRY(k1.alloc(4 * (size_t)n));
  SEMB_TRY(i0.alloc(8 * (size_t)n));
  SEMB_TRY(i1.alloc(8 * (size_t)n));
  SEMB_TRY(fl.alloc(4 * (size_t)n));
  SEMB_TRY(st.alloc(8 * ((size_t)n + 1)));
  SEMB_TRY(cnt.alloc(8));
  SEMB_TRY(info.alloc(4 * sizeof(int)));
  SEMB_CUDA(cudaMemset(info.p, 0, 4 * sizeof(int)));
  const int grid = vec_grid(n);
  id_keys_kernel<<<grid, VEC_NT>>>(d_ids, k0.as<unsigned int>(), n, info.as<int>());
  iota_kernel<<<grid, VEC_NT>>>(i0.as<int64_t>(), n);
  size_t sort_bytes = 0, sel_bytes = 0;
  SEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int64_t>(),
                                            i1.as<int64_t>(), (int)n));
  SEMB_CUDA(cub::DeviceSelect::Flagged(nullptr, sel_bytes, i0.as<int64_t>(), fl.as<int>(), st.as<int64_t>(), cnt.as<int64_t>(), (int)n));
  size_t tb = std::max(sort_bytes, sel_bytes);
  SEMB_TRY(tmp.alloc(tb));
  SEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, k0.as<unsigned int>(), k1.as<unsigned int>(), i0.as<int64_t>(),
                                            i1.as<int64_t>(), (int)n));
  id_flags_kernel<<<grid, VEC_NT>>>(k1.as<unsigned int>(), fl.as<int>(), n);
  iota_kernel<<<grid, VEC_NT>>>(i0.as<int64_t>(), n);  // sorted positions
  SEMB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, i0.as<int64_t>(), fl.as<int>(), st.as<int64_t>(), cnt.as<int64_t>(), (int)n));
  SEMB_CUDA(cudaGetLastError());
  int64_t ng = 0;
  int h_info[4];
  SEMB_CUDA(cudaMemcpy(&ng, cnt.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SEMB_CUDA(cudaMemcpy(h_info, info.p, sizeof(h_info), cudaMemcpyDeviceToHost));
  SEMB_CHECK(h_info[0] == 0, "semb_gs_setup: negative id");
  SEMB_CUDA(cudaMemcpy(st.as<int64_t>() + ng, &n, sizeof(int64_t), cudaMemcpyHostToDevice));
  return semb_set_gather_scatter(c, i1.as<int64_t>(), st.as<int64_t>(), ng, SEMB_DEVICE);
}

int semb_set_mask(semb_ctx* c, const int64_t* mask_ids, int64_t n_mask) {
  SEMB_CHECK(c && (mask_ids || n_mask == 0) && n_mask >= 0, "semb_set_mask: bad argument");
  SEMB_TRY(bind_device(c));
  const size_t words = ((size_t)c->n + 63) / 64 * 2;
  std::vector<uint32_t> bits(words, 0u);
  for (int64_t q = 0; q < n_mask; ++q) {
    const int64_t id = mask_ids[q];
    SEMB_CHECK(id >= 0 && id < c->n, "semb_set_mask: id out of range");
    bits[id >> 5] |= 1u << (id & 31);
  }
  SEMB_TRY(upload(&c->maskbits, bits.data(), words));
  c->have_mask = true;
  drop_jacobi(c);
  return 0;
}

int semb_setup_mask_box(semb_ctx* c, const double* xe, const double* ye, const double* ze, const double* bbox6, double tol,
                        int mem) {
  SEMB_CHECK(c && xe && ye && ze && bbox6, "semb_setup_mask_box: NULL argument");
  SEMB_TRY(bind_device(c));
  const size_t cnt = (size_t)c->n;
  DevTmp tx, ty, tz;
  const double *dx, *dy, *dz;
  SEMB_TRY(dev_view(xe, mem, cnt, tx, &dx));
  SEMB_TRY(dev_view(ye, mem, cnt, ty, &dy));
  SEMB_TRY(dev_view(ze, mem, cnt, tz, &dz));
  const int64_t words = (c->n + 63) / 64 * 2;
  cudaFree(c->maskbits);
  c->maskbits = nullptr;
  SEMB_CUDA(cudaMalloc(&c->maskbits, sizeof(uint32_t) * (size_t)words));
  BBox bb;
  for (int a = 0; a < 3; ++a) {
    bb.lo[a] = bbox6[2 * a];
    bb.hi[a] = bbox6[2 * a + 1];
  }
  mask_bits_kernel<<<vec_grid(words * 32), VEC_NT, 0, c->stream>>>(dx, dy, dz, c->n, bb, tol, c->maskbits, words);
  SEMB_CUDA(cudaGetLastError());
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  c->have_mask = true;
  drop_jacobi(c);
  return 0;
}

int semb_get_mask(semb_ctx* c, double* mask, int mem) {
  SEMB_CHECK(c && mask, "semb_get_mask: NULL argument");
  SEMB_CHECK(c->have_mask, "semb_get_mask: the mask is not set");
  SEMB_TRY(bind_device(c));
  Staged v;
  SEMB_TRY(stage_in(c, mask, mem, c->n, v, false, 0));
  mask_expand_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(c->maskbits, v.dev, c->n);
  SEMB_CUDA(cudaGetLastError());
  SEMB_TRY(stage_out(c, mask, mem, c->n, v));
  if (mem == SEMB_DEVICE) SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int semb_minmax(int device, int64_t n, const double* x, double* lo_hi, int mem) {
  SEMB_CHECK(n > 0 && x && lo_hi, "semb_minmax: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  DevTmp tx, out, tmp;
  const double* dx;
  SEMB_TRY(dev_view(x, mem, (size_t)n, tx, &dx));
  SEMB_TRY(out.alloc(2 * sizeof(double)));
  size_t tb = 0, tb2 = 0;
  SEMB_CUDA(cub::DeviceReduce::Min(nullptr, tb, dx, out.as<double>(), (int)n));
  SEMB_CUDA(cub::DeviceReduce::Max(nullptr, tb2, dx, out.as<double>() + 1, (int)n));
  tb = std::max(tb, tb2);
  SEMB_TRY(tmp.alloc(tb));
  SEMB_CUDA(cub::DeviceReduce::Min(tmp.p, tb, dx, out.as<double>(), (int)n));
  SEMB_CUDA(cub::DeviceReduce::Max(tmp.p, tb, dx, out.as<double>() + 1, (int)n));
  SEMB_CUDA(cudaMemcpy(lo_hi, out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int semb_get_rmult(semb_ctx* c, double* rmult, int mem) {
  SEMB_CHECK(c && rmult, "semb_get_rmult: NULL argument");
  SEMB_CHECK(c->have_gs, "semb_get_rmult: call semb_set_gather_scatter first");
  SEMB_TRY(bind_device(c));
  SEMB_CUDA(cudaMemcpyAsync(rmult, c->rmult, sizeof(double) * (size_t)c->n,
                            mem == SEMB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int semb_ax(semb_ctx* c, const double* p, double* ap, int flags, int mem) {
  SEMB_CHECK(c && p && ap, "semb_ax: NULL argument");
  SEMB_CHECK(c->have_D && c->have_G, "semb_ax: derivative matrix and geometric factors must be set first");
  const bool mask = flags & SEMB_AX_MASK, gs = flags & SEMB_AX_GS;
  SEMB_CHECK(!mask || c->have_mask, "semb_ax: SEMB_AX_MASK without semb_set_mask");
  SEMB_CHECK(!gs || c->have_gs, "semb_ax: SEMB_AX_GS without semb_set_gather_scatter");
  SEMB_CHECK(mem == SEMB_HOST || p != ap, "semb_ax: in-place application is not supported");
  SEMB_TRY(bind_device(c));
  Staged in, out;
  SEMB_TRY(stage_in(c, p, mem, c->n, in, true, 0));
  SEMB_TRY(stage_in(c, ap, mem, c->n, out, false, 1));
  SEMB_TRY(launch_ax(c, in.dev, out.dev, mask, false, 0, 0, c->E, nullptr, nullptr, 0, 0, nullptr));
  if (gs) SEMB_TRY(gs_full(c, out.dev, nullptr));
  SEMB_TRY(stage_out(c, ap, mem, c->n, out));
  return 0;
}

int semb_gs(semb_ctx* c, double* x, int mem) {
  SEMB_CHECK(c && x, "semb_gs: NULL argument");
  SEMB_CHECK(c->have_gs, "semb_gs: call semb_set_gather_scatter first");
  SEMB_TRY(bind_device(c));
  Staged v;
  SEMB_TRY(stage_in(c, x, mem, c->n, v, true, 0));
  SEMB_TRY(gs_full(c, v.dev, nullptr));
  SEMB_TRY(stage_out(c, x, mem, c->n, v));
  return 0;
}

int semb_mask(semb_ctx* c, double* x, int mem) {
  SEMB_CHECK(c && x, "semb_mask: NULL argument");
  SEMB_CHECK(c->have_mask, "semb_mask: call semb_set_mask first");
  SEMB_TRY(bind_device(c));
  Staged v;
  SEMB_TRY(stage_in(c, x, mem, c->n, v, true, 0));
  c->launch_count++;
  mask_apply_bits_kernel<<<vec_grid((c->n + 31) / 32), VEC_NT, 0, c->stream>>>(v.dev, c->maskbits, c->n);
  SEMB_CUDA(cudaGetLastError());
  SEMB_TRY(stage_out(c, x, mem, c->n, v));
  return 0;
}

int semb_dot(semb_ctx* c, const double* x, const double* y, int weighted, double* out, int mem) {
  SEMB_CHECK(c && x && y && out, "semb_dot: NULL argument");
  SEMB_CHECK(!weighted || c->have_gs, "semb_dot: weighted dot needs semb_set_gather_scatter");
  SEMB_TRY(bind_device(c));
  Staged a, b;
  SEMB_TRY(stage_in(c, x, mem, c->n, a, true, 0));
  if (mem == SEMB_HOST && y == x) b.dev = a.dev;
  else SEMB_TRY(stage_in(c, y, mem, c->n, b, true, 1));
  const int grid = vec_grid(c->n);
  c->launch_count++;
  if (weighted) dot_kernel<true><<<grid, VEC_NT, 0, c->stream>>>(a.dev, b.dev, c->rmult, c->n, c->partials);
  else dot_kernel<false><<<grid, VEC_NT, 0, c->stream>>>(a.dev, b.dev, nullptr, c->n, c->partials);
  SEMB_CUDA(cudaGetLastError());
  double* scratch = reinterpret_cast<double*>(c->scal + 1);
  SEMB_TRY(reduce_partials(c, grid, 0, 1, scratch, nullptr));
  SEMB_TRY(allreduce_sum(c, scratch, 1));
  SEMB_CUDA(cudaMemcpyAsync(out, scratch, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int semb_diag(semb_ctx* c, double* diag, int mem) {
  SEMB_CHECK(c && diag, "semb_diag: NULL argument");
  SEMB_CHECK(c->have_D && c->have_G, "semb_diag: derivative matrix and geometric factors must be set first");
  SEMB_TRY(bind_device(c));
  Staged v;
  SEMB_TRY(stage_in(c, diag, mem, c->n, v, false, 0));
  c->launch_count++;
  diag_kernel<<<vec_grid(c->n), VEC_NT, 0, c->stream>>>(c->G6, v.dev, c->E, c->Nq, c->D);
  SEMB_CUDA(cudaGetLastError());
  SEMB_TRY(stage_out(c, diag, mem, c->n, v));
  return 0;
}

// ---- batched single-element pieces ------------------------------------------
static int make_dmat(DMat& dm, const double* D, int Nq) {
  SEMB_CHECK(D != nullptr && Nq >= 1 && Nq <= kMaxNq, "D is NULL or Nq outside [1,16]");
  std::memset(dm.d, 0, sizeof(dm.d));
  std::memcpy(dm.d, D, sizeof(double) * Nq * Nq);
  return 0;
}

struct TmpBuf {
  double* dev = nullptr;
  bool owned = false;
  ~TmpBuf() {
    if (owned && dev) cudaFree(dev);
  }
  int in(const double* src, int mem, size_t count, bool copy) {
    if (mem == SEMB_DEVICE) {
      dev = const_cast<double*>(src);
      return 0;
    }
    SEMB_CUDA(cudaMalloc(&dev, sizeof(double) * std::max<size_t>(count, 1)));
    owned = true;
    if (copy) SEMB_CUDA(cudaMemcpy(dev, src, sizeof(double) * count, cudaMemcpyHostToDevice));
    return 0;
  }
  int out(double* dst, int mem, size_t count) {
    if (mem == SEMB_DEVICE) return 0;
    SEMB_CUDA(cudaMemcpy(dst, dev, sizeof(double) * count, cudaMemcpyDeviceToHost));
    return 0;
  }
};

int semb_gradient(int device, int Nq, const double* D, int64_t nelem, const double* U, double* Ur, double* Us,
                  double* Ut, int mem) {
  SEMB_CHECK(U && Ur && Us && Ut && nelem > 0, "semb_gradient: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  DMat dm;
  SEMB_TRY(make_dmat(dm, D, Nq));
  const size_t cnt = (size_t)nelem * Nq * Nq * Nq;
  TmpBuf u, r, s, t;
  SEMB_TRY(u.in(U, mem, cnt, true));
  SEMB_TRY(r.in(Ur, mem, cnt, false));
  SEMB_TRY(s.in(Us, mem, cnt, false));
  SEMB_TRY(t.in(Ut, mem, cnt, false));
  const int Np = Nq * Nq * Nq;
  const int nt = std::min(256, ((Np + 31) / 32) * 32);
  const size_t smem = sizeof(double) * 3 * Np;
  SEMB_CUDA(cudaFuncSetAttribute(gradient_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 3 * 4096)));
  gradient_kernel<false><<<(unsigned)nelem, nt, smem>>>(u.dev, nullptr, nullptr, r.dev, s.dev, t.dev, Nq, dm);
  SEMB_CUDA(cudaGetLastError());
  SEMB_CUDA(cudaDeviceSynchronize());
  SEMB_TRY(r.out(Ur, mem, cnt));
  SEMB_TRY(s.out(Us, mem, cnt));
  SEMB_TRY(t.out(Ut, mem, cnt));
  return 0;
}

int semb_gradient_transpose(int device, int Nq, const double* D, int64_t nelem, const double* Wr, const double* Ws,
                            const double* Wt, double* out, int mem) {
  SEMB_CHECK(Wr && Ws && Wt && out && nelem > 0, "semb_gradient_transpose: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  DMat dm;
  SEMB_TRY(make_dmat(dm, D, Nq));
  const size_t cnt = (size_t)nelem * Nq * Nq * Nq;
  TmpBuf a, b, cc, o;
  SEMB_TRY(a.in(Wr, mem, cnt, true));
  SEMB_TRY(b.in(Ws, mem, cnt, true));
  SEMB_TRY(cc.in(Wt, mem, cnt, true));
  SEMB_TRY(o.in(out, mem, cnt, false));
  const int Np = Nq * Nq * Nq;
  const int nt = std::min(256, ((Np + 31) / 32) * 32);
  const size_t smem = sizeof(double) * 3 * Np;
  SEMB_CUDA(cudaFuncSetAttribute(gradient_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 3 * 4096)));
  gradient_kernel<true><<<(unsigned)nelem, nt, smem>>>(a.dev, b.dev, cc.dev, o.dev, nullptr, nullptr, Nq, dm);
  SEMB_CUDA(cudaGetLastError());
  SEMB_CUDA(cudaDeviceSynchronize());
  SEMB_TRY(o.out(out, mem, cnt));
  return 0;
}

int semb_kron(int device, const double* Sz, int nz, int mz, const double* Sy, int ny, int my, const double* Sx, int nx,
              int mx, int64_t nelem, const double* U, double* out, int mem) {
  SEMB_CHECK(Sz && Sy && Sx && U && out && nelem > 0, "semb_kron: bad argument");
  SEMB_CHECK(nz > 0 && ny > 0 && nx > 0 && mz > 0 && my > 0 && mx > 0, "semb_kron: non-positive dimension");
  const size_t smem = sizeof(double) * ((size_t)mz * my * mx + (size_t)mz * my * nx + (size_t)mz * ny * nx);
  SEMB_CHECK(smem <= 200 * 1024, "semb_kron: factors too large for one CTA's shared memory");
  SEMB_CUDA(cudaSetDevice(device));
  const size_t nin = (size_t)mz * my * mx, nout = (size_t)nz * ny * nx;
  TmpBuf sz, sy, sx, u, o;
  // the factors are always HOST arrays (tiny)
  SEMB_TRY(sz.in(Sz, SEMB_HOST, (size_t)nz * mz, true));
  SEMB_TRY(sy.in(Sy, SEMB_HOST, (size_t)ny * my, true));
  SEMB_TRY(sx.in(Sx, SEMB_HOST, (size_t)nx * mx, true));
  SEMB_TRY(u.in(U, mem, (size_t)nelem * nin, true));
  SEMB_TRY(o.in(out, mem, (size_t)nelem * nout, false));
  SEMB_CUDA(cudaFuncSetAttribute(kron_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  kron_kernel<<<(unsigned)nelem, 256, smem>>>(sz.dev, nz, mz, sy.dev, ny, my, sx.dev, nx, mx, u.dev, o.dev);
  SEMB_CUDA(cudaGetLastError());
  SEMB_CUDA(cudaDeviceSynchronize());
  SEMB_TRY(o.out(out, mem, (size_t)nelem * nout));
  return 0;
}

// ---- mesh set-up on the device ---------------------------------------------------
int semb_numbering(int device, int64_t n, const double* x, const double* y, const double* z, double tol2, int32_t* gid,
                   int64_t* g2l, int64_t* gstart, int64_t* nglobal, int mem) {
  SEMB_CHECK(n > 0 && x && y && z && gid && g2l && gstart && nglobal, "semb_numbering: bad argument");
  SEMB_CHECK(n < (int64_t)2147483647, "semb_numbering: more than 2^31 nodes");
  SEMB_CUDA(cudaSetDevice(device));
  TmpBuf bx, by, bz;
  SEMB_TRY(bx.in(x, mem, (size_t)n, true));
  SEMB_TRY(by.in(y, mem, (size_t)n, true));
  SEMB_TRY(bz.in(z, mem, (size_t)n, true));
  struct Dev {
    void* p = nullptr;
    ~Dev() { cudaFree(p); }
    int alloc(size_t bytes) {
      SEMB_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
      return 0;
    }
  } k0, k1, i0, i1, fl, sc, st, tmp, cnt;
  SEMB_TRY(k0.alloc(8 * (size_t)n));
  SEMB_TRY(k1.alloc(8 * (size_t)n));
  SEMB_TRY(i0.alloc(8 * (size_t)n));
  SEMB_TRY(i1.alloc(8 * (size_t)n));
  unsigned long long* keys[2] = {(unsigned long long*)k0.p, (unsigned long long*)k1.p};
  int64_t* idx[2] = {(int64_t*)i0.p, (int64_t*)i1.p};
  const int grid = vec_grid(n);
  iota_kernel<<<grid, VEC_NT>>>(idx[0], n);
  size_t tmp_bytes = 0;
  SEMB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys[0], keys[1], idx[0], idx[1], (int)n));
  size_t scan_bytes = 0, sel_bytes = 0;
  SEMB_TRY(fl.alloc(4 * (size_t)n));
  SEMB_TRY(sc.alloc(4 * (size_t)n));
  SEMB_TRY(st.alloc(8 * ((size_t)n + 1)));
  SEMB_TRY(cnt.alloc(8));
  SEMB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int*)fl.p, (int*)sc.p, (int)n));
  SEMB_CUDA(cub::DeviceSelect::Flagged(nullptr, sel_bytes, idx[1], (int*)fl.p, (int64_t*)st.p, (int64_t*)cnt.p, (int)n));
  SEMB_TRY(tmp.alloc(std::max(tmp_bytes, std::max(scan_bytes, sel_bytes))));
  size_t tb = std::max(tmp_bytes, std::max(scan_bytes, sel_bytes));
  // least-significant key first: z, then y, then x (radix sort is stable) == np.lexsort((z, y, x))
  const double* coord[3] = {bz.dev, by.dev, bx.dev};
  int cur = 0;
  for (int pass = 0; pass < 3; ++pass) {
    make_keys_kernel<<<grid, VEC_NT>>>(coord[pass], idx[cur], keys[0], n);
    SEMB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys[0], keys[1], idx[cur], idx[1 - cur], (int)n));
    cur = 1 - cur;
  }
  const int64_t* order = idx[cur];
  gap_flags_kernel<<<grid, VEC_NT>>>(bx.dev, by.dev, bz.dev, order, tol2, (int*)fl.p, n);
  SEMB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, (int*)fl.p, (int*)sc.p, (int)n));
  TmpBuf dummy;
  // ids back in local order (int32), stored in the free key buffer
  int32_t* gid_dev = (int32_t*)keys[0];
  scatter_ids_kernel<<<grid, VEC_NT>>>(order, (int*)sc.p, gid_dev, n);
  // segment starts: sorted positions whose flag is set
  iota_kernel<<<grid, VEC_NT>>>(idx[1 - cur], n);
  SEMB_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, idx[1 - cur], (int*)fl.p, (int64_t*)st.p, (int64_t*)cnt.p, (int)n));
  SEMB_CUDA(cudaGetLastError());
  int64_t ng = 0;
  SEMB_CUDA(cudaMemcpy(&ng, cnt.p, sizeof(int64_t), cudaMemcpyDeviceToHost));
  SEMB_CUDA(cudaMemcpy((int64_t*)st.p + ng, &n, sizeof(int64_t), cudaMemcpyHostToDevice));
  const cudaMemcpyKind kind = mem == SEMB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  SEMB_CUDA(cudaMemcpy(gid, gid_dev, sizeof(int32_t) * (size_t)n, kind));
  SEMB_CUDA(cudaMemcpy(g2l, order, sizeof(int64_t) * (size_t)n, kind));
  SEMB_CUDA(cudaMemcpy(gstart, st.p, sizeof(int64_t) * ((size_t)ng + 1), kind));
  *nglobal = ng;
  return 0;
}

int semb_geom_factors(int device, int Nq, const double* D, const double* w, int64_t nelem, const double* xe,
                      const double* ye, const double* ze, double* G6, double* jaco, int mem) {
  SEMB_CHECK(D && w && xe && ye && ze && G6 && nelem > 0, "semb_geom_factors: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  DMat dm, wm;
  SEMB_TRY(make_dmat(dm, D, Nq));
  std::memset(wm.d, 0, sizeof(wm.d));
  std::memcpy(wm.d, w, sizeof(double) * Nq);
  const size_t cnt = (size_t)nelem * Nq * Nq * Nq;
  TmpBuf bx, by, bz, bg, bj;
  SEMB_TRY(bx.in(xe, mem, cnt, true));
  SEMB_TRY(by.in(ye, mem, cnt, true));
  SEMB_TRY(bz.in(ze, mem, cnt, true));
  SEMB_TRY(bg.in(G6, mem, 6 * cnt, false));
  if (jaco) SEMB_TRY(bj.in(jaco, mem, cnt, false));
  const int Np = Nq * Nq * Nq;
  const size_t smem = sizeof(double) * 3 * Np;
  SEMB_CUDA(cudaFuncSetAttribute(geom_factors_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 3 * 4096)));
  geom_factors_kernel<<<(unsigned)nelem, std::min(256, ((Np + 31) / 32) * 32), smem>>>(bx.dev, by.dev, bz.dev, bg.dev,
                                                                                      jaco ? bj.dev : nullptr, Nq, dm, wm);
  SEMB_CUDA(cudaGetLastError());
  SEMB_CUDA(cudaDeviceSynchronize());
  SEMB_TRY(bg.out(G6, mem, 6 * cnt));
  if (jaco) SEMB_TRY(bj.out(jaco, mem, cnt));
  return 0;
}

}  // extern "C"

// ---- solvers ---------------------------------------------------------------
namespace semb {
template <int NQ>
static int launch_fdm_nq(semb_ctx* c, FdmArgs a, const LlArgs& ll) {
  if (NQ >= 4) {
    // tuned form: lines in registers, S and the eigenvalues from the parameter block
    constexpr int NQT = NQ >= 4 ? NQ : 4;  // (keeps the M < 2 instantiations out of the tuned kernel)
    constexpr int M = NQT - 2;
    using Cfg = FdmCfg<NQT>;
    FdmS<M> P;
    for (int u = 0; u < 6; ++u)
      for (int q = 0; q < M * M; ++q) P.s[u][q] = c->fdm_Sh.d[q];
    for (int q = 0; q < M; ++q) P.lam[q] = c->fdm_lam[q];
    a.ncta = (int)((c->E + Cfg::EPB - 1) / Cfg::EPB);
    fdm_lines_kernel<NQT><<<(unsigned)a.ncta, Cfg::NT, 0, c->stream>>>(a, ll, P, c->E);
  } else {
    fdm_apply_kernel<NQ><<<(unsigned)c->E, FDM_NT, 0, c->stream>>>(a, ll);
  }
  SEMB_CUDA(cudaGetLastError());
  return 0;
}
// z = M^-1 r (fdm_kernels.cuh) + sum rmult r z -> s->rdotr_new (+ all-rank sum and scalar step with finish)
static int launch_fdm(semb_ctx* c, const double* r, double* z, double* hist, int finish, int check_done) {
  ProfScope ps(c, SEMB_K_DIRECTION);
  c->launch_count++;
  FdmArgs a;
  a.r = r;
  a.z = z;
  a.dinv = c->vec[4];
  a.mult = c->mult;
  a.S = c->fdm_S;
  a.esc = c->fdm_invL;
  a.s = c->scal;
  a.hist = hist;
  a.red = GridRed{c->partials, c->partials2, c->ticket};
  a.ncta = (int)c->E;
  a.finish = finish;
  a.check_done = check_done;
  LlArgs ll = ll_args(c, 0);
  ll.nranks = 1;
  if (finish && c->nranks > 1 && use_p2p(c)) ll = ll_args(c, ++c->seq_scalar);
  switch (c->Nq) {
#define SEMB_CASE_F(Q) \
  case Q:              \
    return launch_fdm_nq<Q>(c, a, ll);
    SEMB_CASE_F(2) SEMB_CASE_F(3) SEMB_CASE_F(4) SEMB_CASE_F(5) SEMB_CASE_F(6) SEMB_CASE_F(7) SEMB_CASE_F(8) SEMB_CASE_F(9)
    SEMB_CASE_F(10) SEMB_CASE_F(11) SEMB_CASE_F(12) SEMB_CASE_F(13) SEMB_CASE_F(14) SEMB_CASE_F(15) SEMB_CASE_F(16)
#undef SEMB_CASE_F
    default:
      break;
  }
  set_error("fdm: unsupported Nq");
  return 2;
}
}  // namespace semb

extern "C" {

int semb_set_fdm(semb_ctx* c, const double* S, const double* lambda, const double* elem_scale) {
  SEMB_CHECK(c && S && lambda && elem_scale, "semb_set_fdm: NULL argument");
  SEMB_TRY(bind_device(c));
  const int m = c->Nq - 2;
  std::vector<double> tab((size_t)std::max(m * m + m, 1), 0.0);
  for (int q = 0; q < m * m; ++q) tab[q] = S[q];
  for (int q = 0; q < m; ++q) {
    SEMB_CHECK(lambda[q] > 0.0, "semb_set_fdm: eigenvalues must be positive");
    tab[m * m + q] = lambda[q];
  }
  std::memset(c->fdm_Sh.d, 0, sizeof(c->fdm_Sh.d));
  for (int q = 0; q < m * m; ++q) c->fdm_Sh.d[q] = S[q];
  for (int q = 0; q < m; ++q) c->fdm_lam[q] = lambda[q];
  SEMB_TRY(upload(&c->fdm_S, tab.data(), tab.size()));
  SEMB_TRY(upload(&c->fdm_invL, elem_scale, 4 * (size_t)c->E));
  c->fdm_ready = true;
  return 0;
}

int semb_cg(semb_ctx* c, const double* b, double* x, double tol, int maxit, int precond, int mem, int* niter,
            double* history) {
  SEMB_CHECK(c && b && x && niter, "semb_cg: NULL argument");
  SEMB_CHECK(c->have_D && c->have_G && c->have_gs, "semb_cg: derivative, geometric factors and gather-scatter must be set");
  SEMB_CHECK(precond == SEMB_PRECOND_NONE || precond == SEMB_PRECOND_JACOBI || precond == SEMB_PRECOND_FDM,
             "semb_cg: unknown preconditioner");
  SEMB_CHECK(precond != SEMB_PRECOND_FDM || c->fdm_ready, "semb_cg: SEMB_PRECOND_FDM without semb_set_fdm");
  SEMB_CHECK(maxit >= 0, "semb_cg: maxit must be non-negative");
  SEMB_CHECK(mem == SEMB_HOST || (aligned16(b) && aligned16(x)), "semb_cg: device vectors must be 16-byte aligned");
  SEMB_TRY(bind_device(c));
  const bool jac = precond == SEMB_PRECOND_JACOBI, fdm = precond == SEMB_PRECOND_FDM;
  const int64_t n = c->n;
  for (int v = 0; v < 3; ++v) SEMB_TRY(ensure_vec(c, v));
  SEMB_TRY(ensure_vec(c, 6));
  // two direction buffers: iteration k reads p_{k-1} from pbuf[(k+1)&1] and writes p_k to pbuf[k&1], so that
  // the update kernel of an odd iteration still finds p_{k-1} (the solution is advanced every second iteration)
  double* pbuf[2] = {c->vec[1], c->vec[6]};
  double *r = c->vec[0], *Ap = c->vec[2];
  const double* bdev = b;
  double* xdev = x;
  if (mem == SEMB_HOST) {
    SEMB_TRY(ensure_vec(c, 3));
    Staged sb;
    SEMB_TRY(stage_in(c, b, mem, n, sb, true, 0));
    bdev = sb.dev;
    xdev = c->vec[3];
  }
  if ((jac || fdm) && c->vec[4] == nullptr) SEMB_TRY(build_jacobi(c));
  const double* dinv = c->vec[4];
  double* z = nullptr;
  if (fdm) {
    SEMB_TRY(ensure_vec(c, 5));
    z = c->vec[5];
  }
  if (history != nullptr && maxit > 0) {
    if (c->hist_cap < 5 * (int64_t)maxit) {
      cudaFree(c->hist);
      c->hist = nullptr;
      SEMB_CUDA(cudaMalloc(&c->hist, sizeof(double) * 5 * (size_t)maxit));
      c->hist_cap = 5 * (int64_t)maxit;
    }
  }
  double* hist = (history != nullptr && maxit > 0) ? c->hist : nullptr;
  CgScalars* s = c->scal;
  double* two = reinterpret_cast<double*>(c->scal + 1);
  const int grid = vec_grid(n);
  const bool mask = c->have_mask;
  const bool multi = c->nranks > 1;
  const bool p2p = multi && use_p2p(c);
  const bool pull = use_pull(c);
  // programmatic dependent launch along K1 -> K2' -> K3 -> K1 ...: one GPU, no FDM launch in between
  const bool pdl = g_pdl != 0 && !multi && !fdm;

  // x = 0, r = b, p = 0; |b|_w^2 and r.z                          (elliptic.py:31-46, iterative.py:48-63)
  c->launch_count += 2;
  if (jac) cg_init_kernel<true><<<grid, VEC_NT, 0, c->stream>>>(bdev, xdev, r, pbuf[1], c->rmult, dinv, n, c->partials);
  else cg_init_kernel<false><<<grid, VEC_NT, 0, c->stream>>>(bdev, xdev, r, pbuf[1], c->rmult, nullptr, n, c->partials);
  SEMB_CUDA(cudaGetLastError());
  SEMB_TRY(reduce_partials(c, grid, grid, jac ? 2 : 1, two, nullptr));
  SEMB_TRY(allreduce_sum(c, two, jac ? 2 : 1));
  if (fdm) {
    SEMB_TRY(launch_fdm(c, r, z, nullptr, 0, 0));
    SEMB_TRY(allreduce_sum(c, &s->rdotr_new, 1));
  }
  cg_init_scalars_kernel<<<1, 1, 0, c->stream>>>(s, two, tol, maxit, jac ? 1 : fdm ? 2 : 0);
  SEMB_CUDA(cudaGetLastError());

  const int* done = &s->done;
  // direction update folded into the operator's load phase: p = z + beta p with z = r, dinv r, or the FDM output
  const int dir = jac ? 2 : 1;
  const double* zsrc = fdm ? z : r;
  const int64_t nif = c->have_halo ? c->n_if_elems : 0;
  const int cta1 = ax_num_ctas(c->Nq, 0, nif), cta2 = ax_num_ctas(c->Nq, nif, c->E);
  const int npart = (cta1 + cta2) * ax_partials_per_cta(c->Nq);  // partial sums of p.Ap per operator application
  int issued = 0;
  int chunk = 1;
  while (true) {
    SEMB_CUDA(cudaMemcpyAsync(c->h_pinned, &s->niter, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->h_pinned[2] = 0;
    if (c->p2p && c->ll_error)
      SEMB_CUDA(cudaMemcpyAsync(c->h_pinned + 2, c->ll_error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SEMB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->h_pinned[2]) {
      set_error("peer-memory exchange timed out: a neighbouring rank did not deliver (rank failure or table mismatch)");
      return 4;
    }
    if (c->h_pinned[1] || issued >= maxit) break;
    const int todo = std::min(chunk, maxit - issued);
    for (int it = 0; it < todo; ++it) {
      const int k = issued + it;  // iteration number (0-based)
      const double* p_old = pbuf[(k + 1) & 1];
      double* p = pbuf[k & 1];
      // p = z + beta p; Ap = mask(A p); pAp = p.Ap (local copies)         (elliptic.py:70,48-51)
      // Ap = dssum(Ap)                                                  (elliptic.py:54)
      if (c->have_halo && p2p) {
        // Peer-memory path, no NCCL call.  The elements that touch the slab interface run on a second,
        // high-priority stream, followed there by the kernel that sums this rank's interface copies and
        // writes them straight into the neighbours' memory; the interior elements run at the same time
        // on the main stream (the two launches share one fixed-order p.Ap reduction).  The main stream
        // then sums the longer local segments, picks the neighbours' sums up from local memory
        // (halo_scatter) and, in a passenger CTA of that launch, all-reduces pAp.
        {
          // timed as ONE logical operator application: fork -> join of the two concurrent launches
          ProfScope ps(c, SEMB_K_AX);
          SEMB_CUDA(cudaEventRecord(c->ev_if_done, c->stream));  // p, beta of the previous iteration are final
          SEMB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_if_done, 0));
          SEMB_TRY(launch_ax(c, p_old, Ap, mask, true, dir, 0, nif, zsrc, dinv, 0, cta1 + cta2, done, c->comm_stream, false, p));
          c->seq_halo++;
          SEMB_TRY(halo_gather(c, Ap, c->comm_stream, done));
          SEMB_CUDA(cudaEventRecord(c->ev_gathered, c->comm_stream));
          SEMB_TRY(launch_ax(c, p_old, Ap, mask, true, dir, nif, c->E, zsrc, dinv, cta1, cta1 + cta2, done, nullptr, false, p));
          SEMB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gathered, 0));
        }
        SEMB_TRY(launch_gs(c, Ap, done, pull, npart, &s->pAp));
        SEMB_TRY(halo_scatter(c, Ap, c->stream, done, &s->pAp, 1));
      } else if (c->have_halo) {
        // NCCL fallback: same order, the exchange on a second stream
        SEMB_TRY(launch_ax(c, p_old, Ap, mask, true, dir, 0, nif, zsrc, dinv, 0, cta1 + cta2, done, nullptr, true, p));
        SEMB_CUDA(cudaEventRecord(c->ev_if_done, c->stream));
        SEMB_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_if_done, 0));
        c->seq_halo++;
        SEMB_TRY(halo_gather(c, Ap, c->comm_stream, done));
        SEMB_CUDA(cudaEventRecord(c->ev_gathered, c->comm_stream));
        SEMB_TRY(halo_exchange(c, c->comm_stream));
        SEMB_CUDA(cudaEventRecord(c->ev_exchanged, c->comm_stream));
        SEMB_TRY(launch_ax(c, p_old, Ap, mask, true, dir, nif, c->E, zsrc, dinv, cta1, cta1 + cta2, done, nullptr, true, p));
        SEMB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gathered, 0));
        SEMB_TRY(launch_gs(c, Ap, done, pull, npart, &s->pAp));
        // one communicator: its two NCCL operations must not overlap, so the all-reduce follows the exchange
        SEMB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_exchanged, 0));
        SEMB_TRY(allreduce_sum(c, &s->pAp, 1));
        SEMB_TRY(halo_scatter(c, Ap, c->stream, done));
      } else {
        SEMB_TRY(launch_ax(c, p_old, Ap, mask, true, dir, 0, c->E, zsrc, dinv, 0, cta2, done, nullptr, true, p, pdl));
        SEMB_TRY(launch_gs(c, Ap, done, pull, npart, &s->pAp, pdl));
      }
      // x += alpha p (every second iteration, two terms at once); r -= alpha dssum(Ap); r.z; scalar step
      // (elliptic.py:56-71)
      {
        ProfScope ps(c, SEMB_K_UPDATE);
        c->launch_count++;
        // the CTA that finishes last sums the partials, all-reduces them over the ranks (peer memory) and
        // does the scalar step; with NCCL or FDM that tail runs in separate launches
        const bool tail = !fdm && (!multi || p2p);
        unsigned int* ticket = tail ? c->ticket_update : nullptr;
        LlArgs ll = ll_args(c, 0);
        ll.nranks = 1;
        if (tail && p2p) ll = ll_args(c, ++c->seq_scalar);
#define SEMB_UPDATE_X(J, P, X)                                                                                        \
  SEMB_CUDA(launch_maybe_pdl(cg_update_kernel<J, P, X>, dim3(grid), dim3(VEC_NT), 0, c->stream, pdl, xdev, r, p, p_old, Ap, \
                             c->mult, dinv, c->gs_link, c->gs_sums, s, n, c->partials, ticket, hist, ll))
#define SEMB_UPDATE(J, P)                  \
  do {                                     \
    if (k & 1) SEMB_UPDATE_X(J, P, 2);     \
    else SEMB_UPDATE_X(J, P, 0);           \
  } while (0)
        if (jac && pull) SEMB_UPDATE(true, true);
        else if (jac) SEMB_UPDATE(true, false);
        else if (pull) SEMB_UPDATE(false, true);
        else SEMB_UPDATE(false, false);
#undef SEMB_UPDATE
#undef SEMB_UPDATE_X
        SEMB_CUDA(cudaGetLastError());
      }
      if (fdm) {
        // z = M^-1 r, r.z, all-rank sum and scalar step in the same launch (NCCL: separate launches)
        SEMB_TRY(launch_fdm(c, r, z, hist, (!multi || p2p) ? 1 : 0, 1));
        if (multi && !p2p) {
          SEMB_TRY(allreduce_sum(c, &s->rdotr_new, 1));
          c->launch_count++;
          cg_step_scalars_kernel<<<1, 1, 0, c->stream>>>(s, hist);
        }
      } else if (multi && !p2p) {
        SEMB_TRY(reduce_partials(c, grid, 0, 1, &s->rdotr_new, done));
        SEMB_TRY(allreduce_sum(c, &s->rdotr_new, 1));
        c->launch_count++;
        cg_step_scalars_kernel<<<1, 1, 0, c->stream>>>(s, hist);
      }
      SEMB_CUDA(cudaGetLastError());
    }
    issued += todo;
    chunk = std::min(chunk * 2, 16);
  }
  *niter = c->h_pinned[0];
  if (*niter & 1) {  // the last iteration's term of the solution is still pending
    c->launch_count++;
    cg_flush_x_kernel<<<grid, VEC_NT, 0, c->stream>>>(xdev, pbuf[(*niter - 1) & 1], s, n);
    SEMB_CUDA(cudaGetLastError());
  }
  if (history != nullptr && *niter > 0)
    SEMB_CUDA(cudaMemcpyAsync(history, c->hist, sizeof(double) * 5 * (size_t)(*niter), cudaMemcpyDeviceToHost, c->stream));
  if (mem == SEMB_HOST) SEMB_TRY(hs_copy(c, xdev, x, sizeof(double) * (size_t)n, false));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  return check_ll_error(c);
}

int semb_cg_generic(int device, int64_t n, semb_apply_fn A, void* A_user, semb_apply_fn Minv, void* Minv_user,
                    const double* b, double* x, double tol, int maxit, int mem, int* niter, double* history) {
  SEMB_CHECK(A && b && x && niter && n > 0, "semb_cg_generic: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  const bool pre = Minv != nullptr;
  double* buf = nullptr;  // r, p, Ap, z, xs, bs
  SEMB_CUDA(cudaMalloc(&buf, sizeof(double) * 6 * (size_t)n));
  struct Guard {
    double* p;
    double* q;
    ~Guard() {
      cudaFree(p);
      cudaFree(q);
    }
  } guard{buf, nullptr};
  double *r = buf, *p = buf + n, *Ap = buf + 2 * n, *z = buf + 3 * n, *xs = buf + 4 * n, *bs = buf + 5 * n;
  const int grid = vec_grid(n);
  double* sc = nullptr;  // partials[grid] + scalars: [0]=rdotr [1]=pAp [2]=rdotr_new [3]=norm_b
  SEMB_CUDA(cudaMalloc(&sc, sizeof(double) * (grid + 8)));
  guard.q = sc;
  double* partials = sc + 8;
  const double* bdev = b;
  double* xdev = x;
  if (mem == SEMB_HOST) {
    SEMB_CUDA(cudaMemcpy(bs, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    bdev = bs;
    xdev = xs;
  }
  auto dot = [&](const double* a, const double* bb, double* out_dev, double* out_host) -> int {
    dot_kernel<false><<<grid, VEC_NT>>>(a, bb, nullptr, n, partials);
    reduce_partials_kernel<<<1, RED_NT>>>(partials, grid, 0, 1, out_dev, nullptr);
    SEMB_CUDA(cudaGetLastError());
    SEMB_CUDA(cudaMemcpy(out_host, out_dev, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
  };
  SEMB_CUDA(cudaMemset(xdev, 0, sizeof(double) * (size_t)n));
  SEMB_CUDA(cudaMemcpy(r, bdev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice));
  double norm_b = 0.0, rdotr = 0.0, pAp = 0.0, rnew = 0.0;
  SEMB_TRY(dot(bdev, bdev, sc + 3, &norm_b));
  const double TOL = std::max(tol * tol * norm_b, tol * tol);
  if (pre) {
    SEMB_CUDA(cudaDeviceSynchronize());
    SEMB_CHECK(Minv(Minv_user, r, z, n) == 0, "semb_cg_generic: Minv callback failed");
    SEMB_TRY(dot(r, z, sc + 0, &rdotr));
    SEMB_CUDA(cudaMemcpy(p, z, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice));
  } else {
    rdotr = norm_b;
    SEMB_CUDA(cudaMemcpy(sc + 0, sc + 3, sizeof(double), cudaMemcpyDeviceToDevice));
    SEMB_CUDA(cudaMemcpy(p, r, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice));
  }
  int it = 0;
  const bool early = !pre && rdotr < 1.0e-20;  // iterative.py:15 (cg only)
  while (!early && it < maxit && rdotr > TOL) {
    SEMB_CUDA(cudaDeviceSynchronize());
    SEMB_CHECK(A(A_user, p, Ap, n) == 0, "semb_cg_generic: operator callback failed");
    SEMB_TRY(dot(p, Ap, sc + 1, &pAp));
    axpy_pair_kernel<<<grid, VEC_NT>>>(xdev, r, p, Ap, sc + 0, sc + 1, n);
    SEMB_CUDA(cudaGetLastError());
    if (pre) {
      SEMB_CUDA(cudaDeviceSynchronize());
      SEMB_CHECK(Minv(Minv_user, r, z, n) == 0, "semb_cg_generic: Minv callback failed");
      SEMB_TRY(dot(r, z, sc + 2, &rnew));
    } else {
      SEMB_TRY(dot(r, r, sc + 2, &rnew));
    }
    xpby_kernel<<<grid, VEC_NT>>>(p, pre ? z : r, sc + 2, sc + 0, n);
    SEMB_CUDA(cudaGetLastError());
    if (history) {
      double* h = history + 5 * (size_t)it;
      h[0] = rdotr;
      h[1] = rnew;
      h[2] = rdotr / pAp;
      h[3] = rnew / rdotr;
      h[4] = pAp;
    }
    SEMB_CUDA(cudaMemcpy(sc + 0, sc + 2, sizeof(double), cudaMemcpyDeviceToDevice));
    rdotr = rnew;
    ++it;
  }
  *niter = it;
  if (mem == SEMB_HOST) SEMB_CUDA(cudaMemcpy(x, xdev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
  SEMB_CUDA(cudaDeviceSynchronize());
  return 0;
}


// ---- multi-GPU -------------------------------------------------------------
int semb_comm_unique_id(char* id128) {
  SEMB_CHECK(id128 != nullptr, "semb_comm_unique_id: NULL argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  SEMB_NCCL(GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return 0;
}

int semb_comm_init(semb_ctx* c, int nranks, int rank, const char* id128) {
  SEMB_CHECK(c && nranks >= 1 && rank >= 0 && rank < nranks, "semb_comm_init: bad argument");
  SEMB_CHECK(c->comm == nullptr && c->comm_stream == nullptr, "semb_comm_init: communicator already set");
  SEMB_TRY(bind_device(c));
  if (id128 != nullptr) {  // NULL: peer-memory transport only, no NCCL communicator (saves its creation time)
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    SEMB_NCCL(CommInitRank(&c->comm, nranks, id, rank));
  }
  c->nranks = nranks;
  c->rank = rank;
  int prio_lo = 0, prio_hi = 0;
  SEMB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  SEMB_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi));  // interface work first
  SEMB_CUDA(cudaEventCreateWithFlags(&c->ev_if_done, cudaEventDisableTiming));
  SEMB_CUDA(cudaEventCreateWithFlags(&c->ev_gathered, cudaEventDisableTiming));
  SEMB_CUDA(cudaEventCreateWithFlags(&c->ev_exchanged, cudaEventDisableTiming));
  return 0;
}

int semb_set_halo(semb_ctx* c, int nneigh, const int* nb_rank, const int64_t* nb_count, const int64_t* send_start,
                  const int32_t* send_ids, int64_t n_if, const int64_t* src_start, const int32_t* src_off,
                  const int64_t* copy_start, const int32_t* copy_ids, int64_t n_if_elems) {
  SEMB_CHECK(c != nullptr && nneigh >= 0 && n_if >= 0, "semb_set_halo: bad argument");
  SEMB_CHECK(c->have_gs, "semb_set_halo: call semb_set_gather_scatter first");
  SEMB_CHECK(nneigh == 0 || c->comm_stream != nullptr, "semb_set_halo: call semb_comm_init first");
  SEMB_CHECK(nneigh <= LL_MAXR, "semb_set_halo: too many neighbours");
  SEMB_CHECK(n_if_elems >= 0 && n_if_elems <= c->E, "semb_set_halo: n_if_elems out of range");
  SEMB_TRY(bind_device(c));
  c->nb_rank.assign(nb_rank, nb_rank + nneigh);
  c->nb_off.assign(1, 0);
  for (int q = 0; q < nneigh; ++q) {
    SEMB_CHECK(nb_rank[q] >= 0 && nb_rank[q] < c->nranks && nb_rank[q] != c->rank && nb_count[q] > 0,
               "semb_set_halo: bad neighbour entry");
    c->nb_off.push_back(c->nb_off.back() + nb_count[q]);
  }
  c->n_send = c->nb_off.back();
  c->n_if = n_if;
  c->n_if_elems = n_if_elems;
  if (c->n_send > 0) {
    SEMB_CHECK(send_start && send_ids && src_start && src_off && copy_start && copy_ids, "semb_set_halo: NULL table");
    for (int64_t q = 0; q < send_start[c->n_send]; ++q)
      SEMB_CHECK(send_ids[q] >= 0 && send_ids[q] < c->n, "semb_set_halo: send id out of range");
    for (int64_t q = 0; q < copy_start[n_if]; ++q)
      SEMB_CHECK(copy_ids[q] >= 0 && copy_ids[q] < c->n, "semb_set_halo: copy id out of range");
    SEMB_TRY(upload(&c->send_start, send_start, (size_t)c->n_send + 1));
    SEMB_TRY(upload(&c->send_ids, send_ids, (size_t)send_start[c->n_send]));
    SEMB_TRY(upload(&c->src_start, src_start, (size_t)n_if + 1));
    SEMB_TRY(upload(&c->src_off, src_off, (size_t)src_start[n_if]));
    SEMB_TRY(upload(&c->copy_start, copy_start, (size_t)n_if + 1));
    SEMB_TRY(upload(&c->copy_ids, copy_ids, (size_t)copy_start[n_if]));
    cudaFree(c->hbuf);
    c->hbuf = nullptr;
    SEMB_CUDA(cudaMalloc(&c->hbuf, sizeof(double) * 2 * (size_t)c->n_send));
    // the halo kernels complete the interface nodes: take them out of the pull form of the local schedule
    if (c->gs_link != nullptr) {
      clear_links_kernel<<<vec_grid(copy_start[n_if]), VEC_NT, 0, c->stream>>>(c->gs_link, c->copy_ids, copy_start[n_if]);
      SEMB_CUDA(cudaGetLastError());
      SEMB_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  c->have_halo = c->nranks > 1;
  // New tables invalidate the peer-memory layout on EVERY rank (a collective decision: a per-rank
  // fallback would leave one side on NCCL while the other spins on peer words): NCCL until all ranks
  // have gone through semb_p2p_alloc / semb_p2p_attach again.
  c->p2p = false;
  // multiplicities now count the copies on the other ranks too: over NCCL here, or -- without a NCCL
  // communicator -- at the end of semb_p2p_attach, once the peer-memory transport can carry the exchange
  if (c->comm == nullptr && c->nranks > 1) return 0;
  return build_rmult(c);
}

static void close_peer_mappings(semb_ctx* c) {
  for (int r = 0; r < (int)c->ll_peer.size(); ++r)
    if (r != c->rank && c->ll_peer[r]) cudaIpcCloseMemHandle(c->ll_peer[r]);
  c->ll_peer.clear();
}

int semb_p2p_alloc(semb_ctx* c, int64_t halo_capacity, char* handle64) {
  SEMB_CHECK(c && handle64, "semb_p2p_alloc: NULL argument");
  SEMB_CHECK(halo_capacity >= c->n_send, "semb_p2p_alloc: halo_capacity must cover this rank's interface entries");
  SEMB_CHECK(c->comm_stream != nullptr && c->have_halo, "semb_p2p_alloc: call semb_comm_init and semb_set_halo first");
  SEMB_CHECK(c->nranks <= LL_MAXR, "semb_p2p_alloc: more ranks than the peer-memory path supports");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  SEMB_TRY(bind_device(c));
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  c->p2p = false;
  // Every (re-)attachment starts from a fresh, zeroed region and a fresh handle: the capacity fixes the
  // parity offsets that the WRITERS compute inside this region, and stale sequence numbers of an earlier
  // layout must not be mistaken for data.  Old mappings of the peers' regions are closed first.
  close_peer_mappings(c);
  cudaFree(c->ll_local);
  c->ll_local = nullptr;
  c->ll_halo_cap = std::max<int64_t>(halo_capacity, 1);
  if (c->ll_error == nullptr) SEMB_CUDA(cudaMalloc(&c->ll_error, sizeof(int)));
  SEMB_CUDA(cudaMemset(c->ll_error, 0, sizeof(int)));
  const size_t words = (size_t)LL_SCALAR_WORDS + 2 * 2 * (size_t)c->ll_halo_cap;
  SEMB_CUDA(cudaMalloc(&c->ll_local, words * sizeof(unsigned long long)));
  SEMB_CUDA(cudaMemset(c->ll_local, 0, words * sizeof(unsigned long long)));
  SEMB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  SEMB_CUDA(cudaIpcGetMemHandle(&h, c->ll_local));
  std::memcpy(handle64, &h, sizeof(h));
  return 0;
}

int semb_p2p_attach(semb_ctx* c, const char* handles, const int64_t* peer_recv_off) {
  SEMB_CHECK(c && handles && (peer_recv_off || c->nb_rank.empty()), "semb_p2p_attach: NULL argument");
  SEMB_CHECK(c->ll_local != nullptr, "semb_p2p_attach: call semb_p2p_alloc first");
  SEMB_TRY(bind_device(c));
  close_peer_mappings(c);
  c->ll_peer.assign(c->nranks, nullptr);
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) {
      c->ll_peer[r] = c->ll_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * (size_t)r, sizeof(h));
    void* ptr = nullptr;
    SEMB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->ll_peer[r] = static_cast<unsigned long long*>(ptr);
  }
  SEMB_TRY(upload(&c->ll_peer_dev, c->ll_peer.data(), c->ll_peer.size()));
  std::vector<int64_t> po(peer_recv_off, peer_recv_off + c->nb_rank.size());
  for (size_t q = 0; q < po.size(); ++q)
    SEMB_CHECK(po[q] >= 0 && po[q] + (c->nb_off[q + 1] - c->nb_off[q]) <= c->ll_halo_cap,
               "semb_p2p_attach: a block does not fit the neighbour's receive region");
  c->halo_peers = make_halo_peers(c, po);
  c->seq_scalar = 0;
  c->seq_halo = 0;
  c->p2p = true;
  return build_rmult(c);  // collective: every rank is inside semb_p2p_attach
}

// ---- memory helpers --------------------------------------------------------
int semb_malloc(int device, void** ptr, int64_t bytes) {
  SEMB_CHECK(ptr && bytes >= 0, "semb_malloc: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  SEMB_CUDA(cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 1)));
  return 0;
}
int semb_free(int device, void* ptr) {
  SEMB_CUDA(cudaSetDevice(device));
  SEMB_CUDA(cudaFree(ptr));
  return 0;
}
// pinned (page-locked) host memory: results allocated here come back from the device by direct DMA
int semb_host_alloc(void** ptr, int64_t bytes) {
  SEMB_CHECK(ptr && bytes >= 0, "semb_host_alloc: bad argument");
  SEMB_CUDA(cudaMallocHost(ptr, (size_t)std::max<int64_t>(bytes, 1)));
  return 0;
}
int semb_host_free(void* ptr) {
  SEMB_CUDA(cudaFreeHost(ptr));
  return 0;
}
int semb_memcpy(int device, void* dst, const void* src, int64_t bytes, int kind) {
  SEMB_CHECK(dst && src && bytes >= 0 && kind >= 0 && kind <= 2, "semb_memcpy: bad argument");
  SEMB_CUDA(cudaSetDevice(device));
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  SEMB_CUDA(cudaMemcpy(dst, src, (size_t)bytes, k));
  return 0;
}

// ---- measurement -----------------------------------------------------------
static int prof_flush(semb_ctx* c) {
  SEMB_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < SEMB_K_COUNT; ++k) {
    for (size_t q = 0; q < c->prof.used[k]; ++q) {
      float ms = 0.f;
      SEMB_CUDA(cudaEventElapsedTime(&ms, c->prof.start[k][q], c->prof.stop[k][q]));
      c->prof.total_ms[k] += ms;
    }
    c->prof.used[k] = 0;
  }
  return 0;
}
int semb_profile_enable(semb_ctx* c, int on) {
  SEMB_CHECK(c != nullptr, "semb_profile_enable: ctx is NULL");
  SEMB_TRY(bind_device(c));
  if (!on && c->prof.on) SEMB_TRY(prof_flush(c));
  c->prof.on = on != 0;
  c->prof.mask = on > 1 ? (on >> 1) : 0xffff;  // on = 1: all kernels; on = 1 | (mask << 1): only those in mask
  return 0;
}
int semb_profile_reset(semb_ctx* c) {
  SEMB_CHECK(c != nullptr, "semb_profile_reset: ctx is NULL");
  SEMB_TRY(bind_device(c));
  SEMB_TRY(prof_flush(c));
  for (int k = 0; k < SEMB_K_COUNT; ++k) {
    c->prof.total_ms[k] = 0.0;
    c->prof.launches[k] = 0;
  }
  return 0;
}
int semb_profile_get(semb_ctx* c, int kernel, double* total_ms, int64_t* launches) {
  SEMB_CHECK(c && total_ms && launches && kernel >= 0 && kernel < SEMB_K_COUNT, "semb_profile_get: bad argument");
  SEMB_TRY(bind_device(c));
  SEMB_TRY(prof_flush(c));
  *total_ms = c->prof.total_ms[kernel];
  *launches = c->prof.launches[kernel];
  return 0;
}
int64_t semb_launch_count(semb_ctx* c) { return c ? c->launch_count : -1; }

int semb_set_tuning(const char* name, int value) {
  SEMB_CHECK(name != nullptr, "semb_set_tuning: NULL name");
  if (std::strcmp(name, "ax_l2_prefetch") == 0) {
    g_ax_l2_prefetch = value;
    return 0;
  }
  if (std::strcmp(name, "gs_pull") == 0) {
    g_gs_pull = value;
    return 0;
  }
  if (std::strcmp(name, "pdl") == 0) {
    g_pdl = value;
    return 0;
  }
  if (std::strcmp(name, "p2p") == 0) {
    g_p2p_on = value;
    return 0;
  }
  set_error(std::string("semb_set_tuning: unknown knob ") + name);
  return 2;
}

}  // extern "C"
