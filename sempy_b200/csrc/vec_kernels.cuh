// K2 (gather-scatter), K3/K4 (fused CG vector updates) and the scalar kernels.
#pragma once

#include "common.cuh"
#include "reduce.cuh"

namespace semb {

constexpr int VEC_NT = 256;

// ---------------------------------------------------------------------------
// K2: direct-stiffness summation, atomic-free.
//
// Replaces gslib's gs(x, gs_double, gs_add) as SemPy calls it
// (sempy/mesh.py:444-447; semantics stated by python_gather_scatter,
// sempy/loopy/loopy_kernels.py:26-39).  The schedule holds only segments with
// more than one local copy, grouped by length so that one thread owns one
// segment, fetches its indices with a single vector load, sums the copies in
// ascending local index (fixed order -> bitwise repeatable) and writes the sum
// back to every copy.  Traffic: 8 + 4 + 8 bytes per shared local node.
// ---------------------------------------------------------------------------
template <int L>
__device__ __forceinline__ void gs_load_ids(const int32_t* __restrict__ ids, int64_t s, int (&id)[L]) {
  if (L == 2) {
    const int2 v = __ldg(reinterpret_cast<const int2*>(ids) + s);
    id[0] = v.x;
    id[1] = v.y;
  } else {
#pragma unroll
    for (int c = 0; c < L / 4; ++c) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(ids) + s * (L / 4) + c);
      id[4 * c] = v.x;
      id[4 * c + 1] = v.y;
      id[4 * c + 2] = v.z;
      id[4 * c + 3] = v.w;
    }
  }
}

// U segments per thread, all index loads first, then all gathers, then sums and stores: the pass is
// latency-bound (ncu: 48 % DRAM at exactly the sector-level traffic), so memory-level parallelism per
// thread is what counts.
// sums != nullptr (compact form): the sum of segment s goes to sums[s] only and x is left untouched -- the
// consumer picks it up through its node links -- which halves the scattered sector traffic of the pass.
template <int L, int U>
__device__ __forceinline__ void gs_segments(double* __restrict__ x, const int32_t* __restrict__ ids, int64_t first,
                                            int64_t stride, int64_t nseg, double* __restrict__ sums = nullptr) {
  int id[U][L];
  double v[U][L];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t s = first + u * stride;
    if (s < nseg) gs_load_ids<L>(ids, s, id[u]);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (first + u * stride < nseg) {
#pragma unroll
      for (int c = 0; c < L; ++c) v[u][c] = x[id[u][c]];
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (first + u * stride < nseg) {
      double sum = v[u][0];
#pragma unroll
      for (int c = 1; c < L; ++c) sum += v[u][c];
      if (sums != nullptr) {
        sums[first + u * stride] = sum;
      } else {
#pragma unroll
        for (int c = 0; c < L; ++c) x[id[u][c]] = sum;
      }
    }
  }
}

// All fixed-length classes in ONE launch: the grid is split into three block ranges (len 2 | 4 | 8), so
// a block works on one class only.  (Three launches cost two extra drain/launch bubbles and a nearly
// empty len-8 grid: 8.8 us for 0.13 waves, profiles/r01_summary.md.)
// segments per thread and trip (memory-level parallelism of the latency-bound pass)
#ifndef GS_U2
#define GS_U2 4
#define GS_U4 2
#define GS_U8 1
#endif
struct GsPlan {
  int64_t n2, n4, n8;        // segments per class
  int64_t off2, off4, off8;  // start of each class in ids (multiples of 4 ints)
  int b2, b4, b8;            // blocks per class
};

// sums (compact form, len-4 and len-8 classes only): segment s of class 4 -> sums[s], of class 8 ->
// sums[n4 + s]; the irregular rest follows at sums[n4 + n8 + s] (gs_var_kernel).
// pass.n > 0: the first GS_PASSENGERS blocks of the grid (scheduled first; as the last blocks they would start
// when the others are finishing and put their sums on the critical path) are passengers that add the
// pass.n partial sums of p.Ap left by the operator launch and store the total in *pass.out: every passenger
// sums a fixed slice (16 loads in flight per thread: one CTA alone is latency-bound at ~0.4 ns per partial,
// 100 us for the 262 144 partials of a 64^3-element mesh), the one that finishes last adds the slice sums
// in index order (reduce.cuh).
constexpr int GS_PASSENGERS = 16;
struct GsPassenger {
  const double* part;  // nullptr: no passengers
  int n;
  double* out;
  GridRed red;         // scratch of the slice reduction (not the array `part` points into)
};

__device__ __forceinline__ void gs_passenger(const GsPassenger& pass, int slice) {
  constexpr int PU = 16;
  const int len = (pass.n + GS_PASSENGERS - 1) / GS_PASSENGERS;
  const int lo = slice * len, hi = min(pass.n, lo + len);
  double v[PU];
#pragma unroll
  for (int u = 0; u < PU; ++u) v[u] = 0.0;
  int t = lo + threadIdx.x;
  for (; t + (PU - 1) * VEC_NT < hi; t += PU * VEC_NT) {
#pragma unroll
    for (int u = 0; u < PU; ++u) v[u] += __ldcg(pass.part + t + u * VEC_NT);
  }
  for (; t < hi; t += VEC_NT) v[0] += __ldcg(pass.part + t);
#pragma unroll
  for (int w = PU / 2; w > 0; w >>= 1) {
#pragma unroll
    for (int u = 0; u < w; ++u) v[u] += v[u + w];
  }
  const double mine = block_sum_nt<VEC_NT>(v[0]);
  double total;
  if (grid_reduce<VEC_NT>(pass.red, mine, slice, GS_PASSENGERS, total) && threadIdx.x == 0) *pass.out = total;
}

__global__ void __launch_bounds__(VEC_NT)
gs_fixed_kernel(double* __restrict__ x, const int32_t* __restrict__ ids, const GsPlan plan,
                const int* __restrict__ done, double* __restrict__ sums, const GsPassenger pass) {
  pdl_launch_dependents();
  pdl_wait();
  if (done != nullptr && *done) return;
  int b = blockIdx.x;
  if (pass.part != nullptr) {
    b -= GS_PASSENGERS;
    if (b < 0) {
      gs_passenger(pass, b + GS_PASSENGERS);
      return;
    }
  }
  if (b < plan.b2) {
    const int64_t stride = (int64_t)plan.b2 * VEC_NT;
    for (int64_t t = (int64_t)b * VEC_NT + threadIdx.x; t < plan.n2; t += GS_U2 * stride)
      gs_segments<2, GS_U2>(x, ids + plan.off2, t, stride, plan.n2);
    return;
  }
  b -= plan.b2;
  if (b < plan.b4) {
    const int64_t stride = (int64_t)plan.b4 * VEC_NT;
    for (int64_t t = (int64_t)b * VEC_NT + threadIdx.x; t < plan.n4; t += GS_U4 * stride)
      gs_segments<4, GS_U4>(x, ids + plan.off4, t, stride, plan.n4, sums);
    return;
  }
  b -= plan.b4;
  const int64_t stride = (int64_t)plan.b8 * VEC_NT;
  for (int64_t t = (int64_t)b * VEC_NT + threadIdx.x; t < plan.n8; t += GS_U8 * stride)
    gs_segments<8, GS_U8>(x, ids + plan.off8, t, stride, plan.n8, sums != nullptr ? sums + plan.n4 : nullptr);
}

// Segments of any other length (unstructured meshes): CSR offsets.
__global__ void __launch_bounds__(VEC_NT)
gs_var_kernel(double* __restrict__ x, const int32_t* __restrict__ ids, const int32_t* __restrict__ start,
              int64_t nseg, const int* __restrict__ done, double* __restrict__ sums) {
  if (done != nullptr && *done) return;
  for (int64_t s = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; s < nseg; s += (int64_t)gridDim.x * VEC_NT) {
    const int lo = start[s], hi = start[s + 1];
    double sum = 0.0;
    for (int q = lo; q < hi; ++q) sum += x[ids[q]];
    if (sums != nullptr) {
      sums[s] = sum;
    } else {
      for (int q = lo; q < hi; ++q) x[ids[q]] = sum;
    }
  }
}

// ---------------------------------------------------------------------------
// Pull form.  The consumer of a gather-scattered vector completes the sums while it streams through the
// vector, guided by one 32-bit link per node:
//   0                       : no other copy (or a node that the halo kernels complete): the value as it is
//   GS_LINK_PAIR | index    : exactly one other copy (73 % of the shared nodes of a hex mesh at N = 7, the
//                             face-interior nodes): own value + the partner's UNSUMMED value.  a + b commutes
//                             in IEEE arithmetic, so both copies get the bitwise identical sum.
//   GS_LINK_LONG | segment  : more copies (edges, vertices: 16 % of the nodes): the sum of the node's segment,
//                             formed BEFORE the consumer by the compact form of gs_fixed_kernel / gs_var_kernel
//                             (copies added in ascending local index, one 8-byte store per segment into a
//                             small L2-resident array instead of a scattered write-back to every copy).
// One predicated gather per node, no index table, no divergence, nothing written back to the vector: the
// in-place pass over the len-2 class disappears and the rest moves half its sectors.  Inside semb_cg the
// consumer is the update kernel (r -= alpha * dssum(Ap)).  The pair gathers hit L2: the partner sits in a
// neighbouring element that the same pass streams through.
// ---------------------------------------------------------------------------
constexpr uint32_t GS_LINK_PAIR = 0x80000000u;
constexpr uint32_t GS_LINK_LONG = 0x40000000u;
constexpr uint32_t GS_LINK_INDEX = 0x3fffffffu;

// The other operand of a linked node -- the partner's value or the segment's sum -- or 0.0 without a link.
// A predicated load (no branch): the gathers of a thread's nodes are issued back to back with its
// streaming loads instead of being serialised in divergent branches ahead of them.
__device__ __forceinline__ double gs_other(const double* __restrict__ v, const double* __restrict__ sums, uint32_t link) {
  const double* src = (link & GS_LINK_PAIR) ? v + (link & ~GS_LINK_PAIR) : sums + (link & GS_LINK_INDEX);
  double o;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.f64 %0, [%1];\n\t}"
      : "=d"(o)
      : "l"(src), "r"(link));
  return o;
}
// value of dssum(v) at a node with link `link`, own (unsummed) value `own` and other operand `other`
__device__ __forceinline__ double gs_combine(uint32_t link, double own, double other) {
  return (link & GS_LINK_PAIR) ? own + other : (link & GS_LINK_LONG) ? other : own;
}
__device__ __forceinline__ double gs_linked(const double* __restrict__ v, const double* __restrict__ sums, uint32_t link,
                                            double own) {
  return gs_combine(link, own, gs_other(v, sums, link));
}

// link[copy] = 0 for the copies of interface nodes (their sum is completed by the halo kernels)
__global__ void clear_links_kernel(uint32_t* __restrict__ link, const int32_t* __restrict__ copy_ids, int64_t ncopy) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ncopy; q += (int64_t)gridDim.x * blockDim.x)
    link[copy_ids[q]] = 0u;
}

// ---------------------------------------------------------------------------
// Multi-GPU part of the gather-scatter (elements are split into one slab per
// GPU; nodes on the slab interfaces have copies on more than one rank).
//   halo_gather : hbuf[t] = sum of this rank's copies of send entry t
//   (NCCL send/recv of the send region into the neighbours' receive regions)
//   halo_scatter: every copy of interface node i <- sum over the sharing ranks,
//                 added in ascending rank order on every rank (bitwise equal
//                 copies on all GPUs).
// Takes over what gslib's pairwise exchange would do under MPI (the reference
// passes comm = 0 and never runs it, sempy/mesh.py:437).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(VEC_NT)
halo_gather_kernel(const double* __restrict__ x, const int64_t* __restrict__ start, const int32_t* __restrict__ ids,
                   int64_t nent, double* __restrict__ hbuf, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  for (int64_t t = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; t < nent; t += (int64_t)gridDim.x * VEC_NT) {
    double s = 0.0;
    for (int64_t q = start[t]; q < start[t + 1]; ++q) s += x[ids[q]];
    hbuf[t] = s;
  }
}

__global__ void __launch_bounds__(VEC_NT)
halo_scatter_kernel(double* __restrict__ x, const double* __restrict__ hbuf, const int64_t* __restrict__ src_start,
                    const int32_t* __restrict__ src_off, const int64_t* __restrict__ copy_start,
                    const int32_t* __restrict__ copy_ids, int64_t n_if, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  for (int64_t i = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; i < n_if; i += (int64_t)gridDim.x * VEC_NT) {
    double s = 0.0;
    for (int64_t q = src_start[i]; q < src_start[i + 1]; ++q) s += hbuf[src_off[q]];
    for (int64_t q = copy_start[i]; q < copy_start[i + 1]; ++q) x[copy_ids[q]] = s;
  }
}

// ---------------------------------------------------------------------------
// Peer-memory ("LL") transport over NVLink.  A double travels as two 8-byte words {32 data bits,
// 32-bit sequence number}; 8-byte stores are atomic, so a reader that sees the expected sequence
// number in both words has the value -- no fences, no separate flag, no NCCL call.
// Region layout per rank (in 8-byte words): scalars [2 parity][LL_MAXR ranks][LL_MAXV values][2],
// then the halo part [2 parity][ll_halo_cap doubles][2].
// ---------------------------------------------------------------------------
constexpr int64_t LL_SCALAR_WORDS = 2 * LL_MAXR * LL_MAXV * 2;

__device__ __forceinline__ void ll_store(unsigned long long* dst, double v, unsigned int seq) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  const unsigned int lo = (unsigned int)bits, hi = (unsigned int)(bits >> 32);
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(lo), "r"(seq) : "memory");
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + 1), "r"(hi), "r"(seq) : "memory");
}
// A peer that died (or a table bug) must not hang the GPU: after LL_TIMEOUT_NS without the expected
// sequence number the wait gives up, raises *ll_error and yields NaN; the host turns that into an error.
// The limit is generous because ranks legitimately drift apart by seconds during host-side mesh set-up.
constexpr unsigned long long LL_TIMEOUT_NS = 60000000000ull;
__device__ __forceinline__ unsigned long long ll_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ double ll_load(const unsigned long long* src, unsigned int seq, int* ll_error) {
  unsigned int lo = 0, hi = 0, f0, f1;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  while (true) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(f0) : "l"(src) : "memory");
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(hi), "=r"(f1) : "l"(src + 1) : "memory");
    if (f0 == seq && f1 == seq) break;
    if ((++spins & 1023u) == 0) {
      const unsigned long long now = ll_now();
      if (t0 == 0) t0 = now;
      if (now - t0 > LL_TIMEOUT_NS || *(volatile int*)ll_error) {
        *(volatile int*)ll_error = 1;
        return __longlong_as_double(0x7ff8000000000000ll);
      }
    }
  }
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

// In-place sum of `count` doubles over all ranks; every rank adds in ascending rank order, so the
// result is bitwise identical everywhere.  Runs on the first LL_MAXR*LL_MAXV threads of one CTA
// (all threads of the CTA must call it).
__device__ __forceinline__ void ll_allreduce_cta(double* vals, int count, unsigned long long* const* __restrict__ peers,
                                                 int rank, int nranks, unsigned int seq, int* ll_error) {
  __shared__ double s_v[LL_MAXR][LL_MAXV];
  const int r = threadIdx.x / LL_MAXV, cidx = threadIdx.x % LL_MAXV;
  const int64_t slot = (int64_t)(seq & 1) * (LL_MAXR * LL_MAXV * 2);
  const bool mine = threadIdx.x < LL_MAXR * LL_MAXV && r < nranks && cidx < count;
  if (mine) {
    // my value cidx -> rank r's region, slot [parity][my rank][cidx]
    ll_store(peers[r] + slot + ((int64_t)rank * LL_MAXV + cidx) * 2, vals[cidx], seq);
    s_v[r][cidx] = ll_load(peers[rank] + slot + ((int64_t)r * LL_MAXV + cidx) * 2, seq, ll_error);
  }
  __syncthreads();
  if (threadIdx.x < count) {
    double sum = 0.0;
    for (int q = 0; q < nranks; ++q) sum += s_v[q][threadIdx.x];
    vals[threadIdx.x] = sum;
  }
  __syncthreads();
}
__global__ void __launch_bounds__(LL_MAXR* LL_MAXV)
ll_allreduce_kernel(double* vals, int count, unsigned long long* const* __restrict__ peers, int rank, int nranks,
                    unsigned int seq, int* ll_error) {
  ll_allreduce_cta(vals, count, peers, rank, nranks, seq, ll_error);
}

// halo_gather + remote write: the partial sum of send entry t goes to hbuf[t] (own use) and, as LL
// words, straight into the neighbour's halo part over NVLink.
__global__ void __launch_bounds__(VEC_NT)
halo_gather_p2p_kernel(const double* __restrict__ x, const int64_t* __restrict__ start, const int32_t* __restrict__ ids,
                       int64_t nent, double* __restrict__ hbuf, unsigned long long* const* __restrict__ peers,
                       const HaloPeers hp, int64_t halo_cap, unsigned int seq, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  const int64_t par = (int64_t)(seq & 1) * halo_cap * 2;
  for (int64_t t = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; t < nent; t += (int64_t)gridDim.x * VEC_NT) {
    double s = 0.0;
    for (int64_t q = start[t]; q < start[t + 1]; ++q) s += x[ids[q]];
    hbuf[t] = s;
    int nb = 0;
    while (nb + 1 < hp.nneigh && t >= hp.off[nb + 1]) ++nb;
    unsigned long long* dst = peers[hp.rank_of[nb]] + LL_SCALAR_WORDS + par + (hp.peer_off[nb] + (t - hp.off[nb])) * 2;
    ll_store(dst, s, seq);
  }
}

// Arguments of the in-kernel all-rank sum (peer memory); nranks <= 1 switches it off.
struct LlArgs {
  unsigned long long* const* peers;
  int rank, nranks;
  unsigned int seq;
  int* ll_error;
};

// halo_scatter reading the neighbours' partial sums from this rank's own LL region (spins until they
// have arrived); same ascending-rank summation order as the NCCL form.  With pass_vals != nullptr the
// launch carries one more CTA (the last) that all-reduces pass_count scalars in place while the others
// wait for their halo words: inside semb_cg that is the pAp sum, so the two exchanges of the operator
// phase share one launch and overlap.
__global__ void __launch_bounds__(VEC_NT)
halo_scatter_p2p_kernel(double* __restrict__ x, const double* __restrict__ hbuf, const unsigned long long* __restrict__ ll_halo,
                        int64_t n_send, const int64_t* __restrict__ src_start, const int32_t* __restrict__ src_off,
                        const int64_t* __restrict__ copy_start, const int32_t* __restrict__ copy_ids, int64_t n_if,
                        int64_t halo_cap, unsigned int seq, const int* __restrict__ done, int* ll_error,
                        double* pass_vals, int pass_count, const LlArgs pass) {
  if (done != nullptr && *done) return;
  int nblk = gridDim.x;
  if (pass_vals != nullptr) {
    nblk -= 1;
    if ((int)blockIdx.x == nblk) {
      ll_allreduce_cta(pass_vals, pass_count, pass.peers, pass.rank, pass.nranks, pass.seq, pass.ll_error);
      return;
    }
  }
  const unsigned long long* base = ll_halo + (int64_t)(seq & 1) * halo_cap * 2;
  for (int64_t i = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; i < n_if; i += (int64_t)nblk * VEC_NT) {
    double s = 0.0;
    for (int64_t q = src_start[i]; q < src_start[i + 1]; ++q) {
      const int64_t off = src_off[q];
      s += off < n_send ? hbuf[off] : ll_load(base + (off - n_send) * 2, seq, ll_error);
    }
    for (int64_t q = copy_start[i]; q < copy_start[i + 1]; ++q) x[copy_ids[q]] = s;
  }
}

// x[mask_ids] = 0   (Mesh.apply_mask, sempy/mesh.py:487-494, as an index list)
__global__ void mask_kernel(double* __restrict__ x, const int64_t* __restrict__ ids, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    x[ids[q]] = 0.0;
}

// ---------------------------------------------------------------------------
// Reductions: every CTA writes one partial; a single CTA sums them in a fixed
// order (bitwise repeatable, no atomics).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double s_red[VEC_NT / 32];
  __shared__ double s_out;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < VEC_NT / 32; ++q) s += s_red[q];
    s_out = s;
  }
  __syncthreads();
  return s_out;
}

// out[q] = sum(partials[q*stride .. q*stride + n)) for q < nout (one CTA of 1024 threads, fixed order).
constexpr int RED_NT = 1024;
__global__ void __launch_bounds__(RED_NT)
reduce_partials_kernel(const double* __restrict__ partials, int64_t n, int64_t stride, int nout,
                       double* __restrict__ out, const int* __restrict__ done) {
  if (done != nullptr && *done) return;
  __shared__ double s_red[RED_NT / 32];
  for (int q = 0; q < nout; ++q) {
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    const double* src = partials + q * stride;
    int64_t t = threadIdx.x;
    for (; t + 3 * RED_NT < n; t += 4 * RED_NT) {
      v0 += src[t];
      v1 += src[t + RED_NT];
      v2 += src[t + 2 * RED_NT];
      v3 += src[t + 3 * RED_NT];
    }
    for (; t < n; t += RED_NT) v0 += src[t];
    double v = (v0 + v1) + (v2 + v3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      double w = s_red[threadIdx.x];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
      if (threadIdx.x == 0) out[q] = w;
    }
    __syncthreads();
  }
}

// dot with optional weight: partials[b] = sum_b w x y
template <bool WEIGHTED>
__global__ void __launch_bounds__(VEC_NT)
dot_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ wgt, int64_t n,
           double* __restrict__ partials) {
  double acc = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q < n; q += (int64_t)gridDim.x * VEC_NT)
    acc = fma(WEIGHTED ? wgt[q] * x[q] : x[q], y[q], acc);
  const double s = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// ---------------------------------------------------------------------------
// CG set-up:  x = 0, r = b, p = 0 (the first operator launch forms p = z + 0 * p with z = Minv r);
// partial sums of |b|_w^2 and r.z  (sempy/elliptic.py:31-46; sempy/iterative.py:48-63 for the PCG form).
// ---------------------------------------------------------------------------
template <bool JACOBI>
__global__ void __launch_bounds__(VEC_NT)
cg_init_kernel(const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
               const double* __restrict__ rmult, const double* __restrict__ dinv, int64_t n,
               double* __restrict__ partials) {
  double nb = 0.0, rz = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q < n; q += (int64_t)gridDim.x * VEC_NT) {
    const double bq = b[q];
    const double wb = rmult[q] * bq;
    nb = fma(wb, bq, nb);
    if (JACOBI) rz = fma(wb, dinv[q] * bq, rz);
    x[q] = 0.0;
    r[q] = bq;
    p[q] = 0.0;
  }
  const double s0 = block_sum(nb);
  if (threadIdx.x == 0) partials[blockIdx.x] = s0;
  if (JACOBI) {
    __syncthreads();
    const double s1 = block_sum(rz);
    if (threadIdx.x == 0) partials[gridDim.x + blockIdx.x] = s1;
  }
}

// two[0] = |b|_w^2, two[1] = r.z (precond 1 = Jacobi); precond 2 (FDM): r.z is already in s->rdotr_new
__global__ void cg_init_scalars_kernel(CgScalars* s, const double* two, double tol, int maxit, int precond) {
  const int jacobi = precond != 0;
  const double norm_b = two[0];
  const double rdotr = precond == 2 ? s->rdotr_new : precond == 1 ? two[1] : norm_b;
  s->norm_b = norm_b;
  s->TOL = fmax(tol * tol * norm_b, tol * tol);
  s->rdotr = rdotr;
  s->rdotr_new = rdotr;
  s->pAp = 0.0;
  s->alpha = 0.0;
  s->beta = 0.0;
  s->niter = 0;
  s->maxit = maxit;
  // elliptic.py:43 early-out (CG only; pcg has none) and the loop condition :47 / iterative.py:61
  const bool early = (!jacobi) && (rdotr < 1.0e-20);
  s->done = (early || !(0 < maxit && rdotr > s->TOL)) ? 1 : 0;
}

// After the update's reduction: beta, history, iteration count, stop test
// (sempy/elliptic.py:47,59-71).
__device__ __forceinline__ void cg_step_scalars(CgScalars* s, double* hist) {
  const double r0 = s->rdotr, r1 = s->rdotr_new;
  const double alpha = r0 / s->pAp;
  const double beta = r1 / r0;
  if (hist != nullptr) {
    double* h = hist + 5 * (int64_t)s->niter;
    h[0] = r0;
    h[1] = r1;
    h[2] = alpha;
    h[3] = beta;
    h[4] = s->pAp;
  }
  s->alpha = alpha;
  s->beta = beta;
  s->rdotr = r1;
  s->niter += 1;
  s->done = !(s->niter < s->maxit && r1 > s->TOL) ? 1 : 0;
}

// The CTA that finishes last adds the per-CTA partials (fixed order -> repeatable whichever CTA it is),
// sums over the ranks when there are several (peer-memory all-reduce inside this CTA) and performs the
// scalar step, so no reduction launch follows the update.  ticket == nullptr (NCCL fallback) leaves
// all of that to separate launches.
__device__ __forceinline__ void cg_update_finish(CgScalars* s, const double* partials, unsigned int* ticket,
                                                 double* hist, const LlArgs& ll) {
  if (ticket == nullptr) return;
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    __threadfence();  // this CTA's partial is visible before its ticket
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double v = 0.0;
  for (unsigned int t = threadIdx.x; t < gridDim.x; t += VEC_NT) v += __ldcg(partials + t);
  const double tot = block_sum(v);
  if (threadIdx.x == 0) {
    s->rdotr_new = tot;
    *ticket = 0;
  }
  if (ll.nranks > 1) {
    __syncthreads();
    ll_allreduce_cta(&s->rdotr_new, 1, ll.peers, ll.rank, ll.nranks, ll.seq, ll.ll_error);
  }
  if (threadIdx.x == 0) cg_step_scalars(s, hist);
}

// ---------------------------------------------------------------------------
// K3: x += alpha p ; r -= alpha dssum(Ap) ; partial of sum rmult r r (or rmult r dinv r)
// (sempy/elliptic.py:52-60; iterative.py:66-73).  With PULL the pairwise part of the gather-scatter of Ap
// (elliptic.py:54) is completed here: nodes with exactly one other copy add the partner's UNSUMMED operator
// output on the fly, nodes of longer segments take their segment's sum from the compact array (gs_linked),
// so Ap is neither re-read by a separate full pass nor written back.
// 57 B per node with Jacobi (+4 link with PULL, + the partner / segment-sum sectors from L2).
// ---------------------------------------------------------------------------
// The weight rmult = 1/multiplicity is read as a one-byte copy count (1 B instead of 8 B per node) and
// turned back into the identical double through a small shared-memory table of 1.0/c (a division in
// the loop costs a MUFU + Newton sequence with a slow-path call per element and made the kernel
// 16 % slower than reading the 8-byte weights).
// The solution is advanced every SECOND iteration: x += alpha_{k-1} p_{k-1} + alpha_k p_k (XMODE 2, odd k;
// the operator kernel alternates between two direction buffers, so p_{k-1} still exists), nothing for even
// k (XMODE 0); semb_cg applies a pending term after the loop (cg_flush_x_kernel).  Same operations in the
// same order as one update per iteration (x2 = fma(a1, p1, fma(a0, p0, x0))), so the same bits, for 8 B per
// node and iteration less traffic (x is read and written half as often).  XMODE 1: one update per iteration.
template <bool JACOBI, bool PULL, int XMODE>
__global__ void __launch_bounds__(VEC_NT, 4)
cg_update_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                 const double* __restrict__ p_prev, const double* __restrict__ Ap, const unsigned char* __restrict__ mult, const double* __restrict__ dinv,
                 const uint32_t* __restrict__ link, const double* __restrict__ sums, CgScalars* __restrict__ s, int64_t n,
                 double* __restrict__ partials, unsigned int* __restrict__ ticket, double* __restrict__ hist,
                 const LlArgs ll) {
  pdl_launch_dependents();
  pdl_wait();
  if (s->done) return;
  __shared__ double s_rcp[256];
  s_rcp[threadIdx.x] = 1.0 / (double)(threadIdx.x > 0 ? threadIdx.x : 1);  // VEC_NT == 256 entries
  __syncthreads();
  const double alpha = s->rdotr / s->pAp;
  const double alpha_prev = s->alpha;  // of the previous iteration (XMODE 2)
  double acc = 0.0;
  const int64_t n2 = n >> 1;
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  const double2* p2 = reinterpret_cast<const double2*>(p);
  const double2* pp2 = reinterpret_cast<const double2*>(p_prev);
  const double2* a2 = reinterpret_cast<const double2*>(Ap);
  const uchar2* m2 = reinterpret_cast<const uchar2*>(mult);
  const double2* d2 = reinterpret_cast<const double2*>(dinv);
  const uint2* l2 = reinterpret_cast<const uint2*>(link);
  // U chunks per trip with all their loads (links first, then the gathers that depend on them, then the
  // streams) issued before the first use: the iterations that do not touch x move only 37 B per node and
  // were latency-bound at one chunk per trip (ncu: 56 % of the DRAM peak, 88 % long-scoreboard stalls).
  constexpr int U = XMODE == 0 ? 2 : 1;
  const int64_t stride = (int64_t)gridDim.x * VEC_NT;
  for (int64_t q0 = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q0 < n2; q0 += U * stride) {
    uint2 lk[U];
    double2 ov[U], av[U], rv[U], dv[U];
    uchar2 mc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      lk[u] = make_uint2(0u, 0u);
      if (PULL && q0 + u * stride < n2) lk[u] = __ldg(l2 + q0 + u * stride);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ov[u] = make_double2(0.0, 0.0);
      if (PULL) {
        ov[u].x = gs_other(Ap, sums, lk[u].x);
        ov[u].y = gs_other(Ap, sums, lk[u].y);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = q0 + u * stride;
      if (q < n2) {
        av[u] = a2[q];
        mc[u] = m2[q];
        rv[u] = r2[q];
        if (JACOBI) dv[u] = d2[q];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = q0 + u * stride;
      if (q >= n2) continue;
      if (XMODE != 0) {
        const double2 pv = p2[q];
        double2 xv = x2[q];
        if (XMODE == 2) {
          const double2 qv = pp2[q];
          xv.x = xv.x + alpha_prev * qv.x;
          xv.y = xv.y + alpha_prev * qv.y;
        }
        xv.x = xv.x + alpha * pv.x;
        xv.y = xv.y + alpha * pv.y;
        x2[q] = xv;
      }
      if (PULL) {
        av[u].x = gs_combine(lk[u].x, av[u].x, ov[u].x);
        av[u].y = gs_combine(lk[u].y, av[u].y, ov[u].y);
      }
      rv[u].x = rv[u].x - alpha * av[u].x;
      rv[u].y = rv[u].y - alpha * av[u].y;
      r2[q] = rv[u];
      const double wx = s_rcp[mc[u].x], wy = s_rcp[mc[u].y];
      if (JACOBI) {
        acc = fma(wx * rv[u].x, dv[u].x * rv[u].x, acc);
        acc = fma(wy * rv[u].y, dv[u].y * rv[u].y, acc);
      } else {
        acc = fma(wx * rv[u].x, rv[u].x, acc);
        acc = fma(wy * rv[u].y, rv[u].y, acc);
      }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t q = n - 1;
    double av = Ap[q];
    if (PULL) av = gs_linked(Ap, sums, link[q], av);
    if (XMODE != 0) {
      double xv = x[q];
      if (XMODE == 2) xv = xv + alpha_prev * p_prev[q];
      x[q] = xv + alpha * p[q];
    }
    const double rv = r[q] - alpha * av;
    r[q] = rv;
    acc = fma(s_rcp[mult[q]] * rv, JACOBI ? dinv[q] * rv : rv, acc);
  }
  const double sum = block_sum(acc);
  if (threadIdx.x == 0) partials[blockIdx.x] = sum;
  cg_update_finish(s, partials, ticket, hist, ll);
}

// x += alpha p for the term that is still pending when the loop ends after an odd number of iterations
__global__ void __launch_bounds__(VEC_NT)
cg_flush_x_kernel(double* __restrict__ x, const double* __restrict__ p, const CgScalars* __restrict__ s, int64_t n) {
  const double alpha = s->alpha;
  for (int64_t q = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q < n; q += (int64_t)gridDim.x * VEC_NT)
    x[q] = x[q] + alpha * p[q];
}

__global__ void cg_step_scalars_kernel(CgScalars* s, double* hist) {
  if (s->done) return;
  cg_step_scalars(s, hist);
}

// dinv = mask ? 0 : 1/diag   (Jacobi preconditioner on assembled diagonal)
__global__ void jacobi_invert_kernel(double* __restrict__ d, const uint32_t* __restrict__ maskbits, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const bool masked = maskbits != nullptr && ((maskbits[q >> 5] >> (q & 31)) & 1u);
    d[q] = masked ? 0.0 : 1.0 / d[q];
  }
}

// rmult = 1 / count  (sempy/mesh.py:440-442): fill with ones, gather-scatter, invert
__global__ void fill_kernel(double* __restrict__ x, double v, int64_t n) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) x[q] = v;
}
// info[0] is raised when a node has more than 255 copies (the update kernel's weights are one byte)
__global__ void reciprocal_kernel(double* __restrict__ x, unsigned char* __restrict__ count, int64_t n, int* __restrict__ info) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const double c = x[q];  // exact small integer: sum of ones
    if (c > 255.0) atomicOr(info, 1);
    count[q] = (unsigned char)(c < 255.0 ? c : 255.0);
    x[q] = 1.0 / c;
  }
}

// ---- generic-operator CG pieces (sempy/iterative.py) ----------------------
// x += a p ; r -= a Ap  with a = num/den read from device memory
__global__ void __launch_bounds__(VEC_NT)
axpy_pair_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                 const double* __restrict__ Ap, const double* num, const double* den, int64_t n) {
  const double a = *num / *den;
  for (int64_t q = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q < n; q += (int64_t)gridDim.x * VEC_NT) {
    x[q] = x[q] + a * p[q];
    r[q] = r[q] - a * Ap[q];
  }
}
// p = z + (num/den) p
__global__ void __launch_bounds__(VEC_NT)
xpby_kernel(double* __restrict__ p, const double* __restrict__ z, const double* num, const double* den, int64_t n) {
  const double bcoef = *num / *den;
  for (int64_t q = (int64_t)blockIdx.x * VEC_NT + threadIdx.x; q < n; q += (int64_t)gridDim.x * VEC_NT)
    p[q] = z[q] + bcoef * p[q];
}

}  // namespace semb
