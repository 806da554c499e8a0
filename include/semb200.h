/* semb200 -- C ABI of the B200-native SemPy Poisson hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, an opaque context,
 * int status returns (0 = ok, non-zero = error; text via semb_last_error()),
 * no exceptions and no torch types across it.  The Python package
 * `sempy_b200` binds it with ctypes; INTEGRATION.md shows the stub a SemPy
 * maintainer would add.  Every entry point names the reference interface
 * (file:line under the SemPy checkout) whose work it takes over.
 *
 * Conventions
 *   - All vectors are "local" element-major FP64 arrays of length E*Np,
 *     Np = Nq^3, node index inside an element = k*Nq^2 + j*Nq + i, i fastest
 *     (sempy/gradient.py:12-21, sempy/mass.py:31).
 *   - `mem` says where caller-owned vector pointers live: SEMB_HOST (the
 *     library stages through device scratch and copies back, synchronously)
 *     or SEMB_DEVICE (used in place, asynchronously on the context's stream).
 *   - One context per GPU; calls on one context are serialised on its stream.
 *     The context owns D, G, the gather-scatter and mask tables, CG work
 *     vectors and scalars.  Set-up arrays are always HOST pointers unless a
 *     `mem` argument says otherwise.
 */
#ifndef SEMB200_H
#define SEMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct semb_ctx semb_ctx;

enum { SEMB_HOST = 0, SEMB_DEVICE = 1 };
/* layout of the geometric factors handed to semb_set_geom */
enum { SEMB_G9 = 9 /* (E,3,3,Np), mesh.get_geom() */, SEMB_G6 = 6 /* (E,6,Np): rr,rs,rt,ss,st,tt */ };
/* flags of semb_ax */
enum { SEMB_AX_MASK = 1 /* zero Dirichlet nodes */, SEMB_AX_GS = 2 /* then gather-scatter */ };
/* preconditioner of semb_cg */
enum { SEMB_PRECOND_NONE = 0, SEMB_PRECOND_JACOBI = 1, SEMB_PRECOND_FDM = 2 };
/* kernels that can be timed with semb_profile_* */
enum { SEMB_K_AX = 0 /* fused operator (+ direction update, mask, pAp) */, SEMB_K_GS = 1 /* gather-scatter pass */,
       SEMB_K_UPDATE = 2 /* CG update (+ pulled gather-scatter, r.z, scalar step) */,
       SEMB_K_DIRECTION = 3 /* preconditioner application (FDM) */, SEMB_K_COUNT = 4 };

int semb_version(void);
/* Text of the last error on the calling thread ("" if none). */
const char* semb_last_error(void);
/* Number of CUDA devices visible, or a negative status. */
int semb_device_count(void);

/* ---- context ------------------------------------------------------------
 * The reference's plug-in seam is the duck-typed mesh object (get_num_elems,
 * Nq, Np, get_geom, get_rmult, apply_mask, dssum; sempy/elliptic.py:10-14,31)
 * plus the gslib handle GS(comm).setup(ids) (sempy/mesh.py:437-438).  A
 * context holds the device-resident form of exactly that state. */
int semb_create(semb_ctx** out, int device, int64_t num_elems, int Nq);
int semb_destroy(semb_ctx* ctx);
/* Run on this CUDA stream (cudaStream_t as void*); NULL = the context's own
 * non-blocking stream.  For the legacy default stream pass cudaStreamLegacy (0x1). */
int semb_set_stream(semb_ctx* ctx, void* cuda_stream);
int semb_synchronize(semb_ctx* ctx);

/* D (Nq*Nq, row-major): reference_derivative_matrix(N), sempy/derivative.py:34.
 * Passed in so both paths use identical bits (gradient.py:7 rebuilds it). */
int semb_set_derivative(semb_ctx* ctx, const double* D);
/* Geometric factors of Mesh.calc_geometric_factors / get_geom
 * (sempy/mesh.py:304-356, 514-517).  Stored on the device as 6 components. */
int semb_set_geom(semb_ctx* ctx, const double* G, int layout, int mem);
/* Gather-scatter tables = Mesh.get_global_to_local_map() (sempy/mesh.py:422-423,
 * 479-480): local indices sorted by global id and the nglobal+1 segment
 * offsets (HOST or DEVICE arrays, `mem`).  Replaces gslib's GS.setup
 * (mesh.py:437-438); the schedule (segments with more than one copy, grouped by
 * length, copies in ascending local index, plus the partner links of the two-copy
 * segments) is built on the device.  Also builds rmult = 1/multiplicity
 * (mesh.py:440-442). */
int semb_set_gather_scatter(semb_ctx* ctx, const int64_t* global_to_local, const int64_t* global_start,
                            int64_t nglobal, int mem);
/* The reference's gslib handle on its own (sempy/mesh.py:437-438: gs = GS(0);
 * gs.setup(global_id int32, 1-based)): semb_gs_create makes a context that only
 * gather-scatters vectors of n entries (semb_gs, semb_get_rmult); semb_gs_setup
 * groups the ids on the device -- equal positive ids are copies of one node, id 0
 * takes no part -- and builds the same schedule as semb_set_gather_scatter. */
int semb_gs_create(semb_ctx** ctx, int device, int64_t n);
int semb_gs_setup(semb_ctx* ctx, const int32_t* ids, int mem);
/* Dirichlet nodes = Mesh.get_mask_ids() (sempy/mesh.py:482-485), HOST ids. */
int semb_set_mask(semb_ctx* ctx, const int64_t* mask_ids, int64_t n_mask);
/* Mesh.setup_mask on the device (sempy/mesh.py:449-477): a node is a Dirichlet
 * node when any coordinate lies within tol of the bounding box
 * bbox6 = {xmin, xmax, ymin, ymax, zmin, zmax} (HOST array; the box of the WHOLE
 * mesh when it is split over ranks).  Coordinates are (E*Np) arrays, `mem`. */
int semb_setup_mask_box(semb_ctx* ctx, const double* xe, const double* ye, const double* ze, const double* bbox6,
                        double tol, int mem);
/* The mask as the reference stores it (mesh.mask): 0.0 at Dirichlet nodes, 1.0
 * elsewhere, E*Np doubles. */
int semb_get_mask(semb_ctx* ctx, double* mask, int mem);
/* lo_hi[0] = min, lo_hi[1] = max of n doubles (bounding box of setup_mask,
 * sempy/mesh.py:451-457); lo_hi is a HOST array. */
int semb_minmax(int device, int64_t n, const double* x, double* lo_hi, int mem);
/* Mesh.calc_geometric_factors (sempy/mesh.py:304-356) straight into the context:
 * G = (grad a . grad b) * B * J from the coordinates (E*Np each, `mem`), B the
 * tensor product of the GLL weights w[Nq] (HOST); jaco (optional, `mem`) receives
 * the Jacobian.  Needs semb_set_derivative.  Nothing of G crosses PCIe. */
int semb_setup_geom(semb_ctx* ctx, const double* w, const double* xe, const double* ye, const double* ze, double* jaco,
                    int mem);
/* Copy the six stored components (E,6,Np) out: mesh.get_geom(), sempy/mesh.py:514-517. */
int semb_get_geom(semb_ctx* ctx, double* G6, int mem);
/* Copy rmult (E*Np doubles) out; mesh.get_rmult(), sempy/mesh.py:511. */
int semb_get_rmult(semb_ctx* ctx, double* rmult, int mem);

/* ---- operators ----------------------------------------------------------
 * ap = A_local p per element: D^T G D, no assembly (sempy/elliptic.py:9-27,
 * gradient.py:6-23,39-56).  flags add Mesh.apply_mask (mesh.py:487-494) and
 * Mesh.dssum (mesh.py:444-447) as elliptic_cg applies them (elliptic.py:48-54). */
int semb_ax(semb_ctx* ctx, const double* p, double* ap, int flags, int mem);
/* In-place direct-stiffness sum; Mesh.dssum -> gslib gs(x, gs_double, gs_add),
 * sempy/mesh.py:444-447; spec in sempy/loopy/loopy_kernels.py:26-39. */
int semb_gs(semb_ctx* ctx, double* x, int mem);
/* In-place x[mask_ids] = 0; Mesh.apply_mask, sempy/mesh.py:487-494. */
int semb_mask(semb_ctx* ctx, double* x, int mem);
/* out = sum_i w_i x_i y_i with w = rmult (weighted=1) or 1 (weighted=0):
 * the dots of sempy/elliptic.py:33,37,51,60. `out` is a HOST double. */
int semb_dot(semb_ctx* ctx, const double* x, const double* y, int weighted, double* out, int mem);
/* Unassembled diagonal of the element operator; input of
 * the Jacobi preconditioner. */
int semb_diag(semb_ctx* ctx, double* diag, int mem);

/* Batched single-element pieces, any E_b elements of order Nq-1 with the D
 * given: sempy/gradient.py:6-23 and :39-56.  Arrays are (E_b, Nq^3). */
int semb_gradient(int device, int Nq, const double* D, int64_t nelem, const double* U, double* Ur, double* Us,
                  double* Ut, int mem);
int semb_gradient_transpose(int device, int Nq, const double* D, int64_t nelem, const double* Wr,
                            const double* Ws, const double* Wt, double* out, int mem);
/* (Sz (x) Sy (x) Sx) U for nelem inputs, C order; sempy/kron.py:17-43.
 * S? are (n?, m?) row-major; U is (nelem, mz*my*mx); out (nelem, nz*ny*nx). */
int semb_kron(int device, const double* Sz, int nz, int mz, const double* Sy, int ny, int my, const double* Sx,
              int nx, int mx, int64_t nelem, const double* U, double* out, int mem);

/* ---- mesh set-up on the device (inputs of the path) -------------------------
 * Global numbering of n local nodes by coordinate, the rule of
 * Mesh.establish_global_numbering (sempy/mesh.py:389-435): stable lexicographic
 * sort by exact (x, y, z), a new id whenever the SQUARED distance to the next
 * sorted point exceeds tol2 (the reference uses 1e-12), ids from 1.  Outputs:
 * gid int32[n] (mesh.global_id), g2l int64[n] (global_to_local), gstart
 * int64[<= n+1] (global_start, *nglobal + 1 entries).  Bit-identical to the
 * reference's Python sort for the same coordinates. */
int semb_numbering(int device, int64_t n, const double* x, const double* y, const double* z, double tol2, int32_t* gid,
                   int64_t* g2l, int64_t* gstart, int64_t* nglobal, int mem);
/* Geometric factors of Mesh.calc_geometric_factors (sempy/mesh.py:304-356) for
 * nelem elements: G6 (nelem,6,Np) = (grad a . grad b) * B * J with B the tensor
 * product of the GLL weights w[Nq], and (optional) the Jacobian (nelem,Np). */
int semb_geom_factors(int device, int Nq, const double* D, const double* w, int64_t nelem, const double* xe,
                      const double* ye, const double* ze, double* G6, double* jaco, int mem);

/* ---- solvers ------------------------------------------------------------
 * elliptic_cg(mesh, b, tol, maxit, verbose), sempy/elliptic.py:30-73, run
 * entirely on the device: weighted dots, pAp before the gather-scatter, stop
 * when rdotr <= max(tol^2 |b|_w^2, tol^2), early-out rdotr < 1e-20.
 * precond = SEMB_PRECOND_JACOBI runs the PCG of sempy/iterative.py:45-85 with
 * Minv = mask/dssum(diag) on the same local vectors (stop on r.z).
 * history: NULL or HOST array of 5*maxit doubles, filled per iteration with
 * rdotr0, rdotr, alpha, beta, pAp (the values the reference prints with
 * verbose=1, elliptic.py:63-68). */
int semb_cg(semb_ctx* ctx, const double* b, double* x, double tol, int maxit, int precond, int mem, int* niter,
            double* history);
/* Fast-diagonalisation preconditioner for semb_cg(precond = SEMB_PRECOND_FDM), the
 * multi-element form of the reference's sketch (example.py:71-114 eigenproblems,
 * :155-171 fast_kron, :186-190 precon_fdm): per element interior
 * (S(x)S(x)S) diag(scale/(ax l_a + ay l_b + az l_c)) (S(x)S(x)S)^T, point Jacobi on
 * the element surfaces.  S: (Nq-2)^2 row-major [node][mode] eigenvectors of the
 * interior 1-D problem A s = l B s with S^T B S = I; lambda: Nq-2 eigenvalues;
 * elem_scale: (E,4) = ax, ay, az, scale per element.  All HOST arrays. */
int semb_set_fdm(semb_ctx* ctx, const double* S, const double* lambda, const double* elem_scale);

/* cg(A,b,...) / pcg(A,Minv,b,...) of sempy/iterative.py:4-42,45-85 for any
 * operator: the vector updates and dots run on the device, A and Minv are
 * callbacks that receive DEVICE pointers (n doubles in, n doubles out) and
 * return 0 on success.  Minv = NULL selects cg.  Synchronisation contract: the
 * loop's own kernels run on the legacy default stream and the device is idle
 * (cudaDeviceSynchronize) when a callback is entered; a callback may use any
 * stream but must have its result complete in d_out when it returns (e.g. end
 * with semb_synchronize on the context it used). */
typedef int (*semb_apply_fn)(void* user, const double* d_in, double* d_out, int64_t n);
int semb_cg_generic(int device, int64_t n, semb_apply_fn A, void* A_user, semb_apply_fn Minv, void* Minv_user,
                    const double* b, double* x, double tol, int maxit, int mem, int* niter, double* history);

/* ---- multi-GPU ----------------------------------------------------------
 * One process and one context per GPU; the elements are split into one
 * contiguous slab per rank (the reference is single-process: it hands gslib
 * comm = 0, sempy/mesh.py:437, and never partitions).  Nodes on the slab
 * interfaces have copies on several ranks: semb_gs / semb_ax(GS) / semb_cg
 * complete their sums with a neighbour exchange (peer memory over NVLink once
 * semb_p2p_attach has run, NCCL before that and as the fallback; overlapped with
 * the interior elements inside semb_cg) and the CG dots with an all-reduce.
 *   semb_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller
 *     broadcasts it (torch.distributed, MPI, a file...).
 *   semb_comm_init: collective over all ranks.  id128 == NULL: no NCCL
 *     communicator at all (its creation takes seconds on 8 GPUs): the context
 *     then relies on the peer-memory transport, multiplicities are completed
 *     inside semb_p2p_attach, and an operation that would need NCCL fails.
 *   semb_set_halo: interface tables of this rank, built by the host-side
 *     partitioner (sempy_b200/distributed.py).  Send entries are grouped by
 *     neighbour in an order both sides share; entry t is the sum of the local
 *     copies send_ids[send_start[t] .. send_start[t+1]).  The exchange buffer is
 *     [send entries | entries received, same layout].  Interface node i gets
 *     the sum of buffer slots src_off[src_start[i] .. src_start[i+1]) (listed in
 *     ascending rank order on every rank, so all GPUs add in the same order)
 *     written to its local copies copy_ids[copy_start[i] .. copy_start[i+1]).
 *     Elements [0, n_if_elems) are the ones that touch the interface; they are
 *     computed first so that the exchange overlaps the rest. */
int semb_comm_unique_id(char* id128);
int semb_comm_init(semb_ctx* ctx, int nranks, int rank, const char* id128);
int semb_set_halo(semb_ctx* ctx, int nneigh, const int* nb_rank, const int64_t* nb_count, const int64_t* send_start,
                  const int32_t* send_ids, int64_t n_if, const int64_t* src_start, const int32_t* src_off,
                  const int64_t* copy_start, const int32_t* copy_ids, int64_t n_if_elems);

/* Peer-memory transport (NVLink / NVSwitch, one node): after semb_set_halo each rank allocates its
 * receive region and publishes a 64-byte CUDA IPC handle (semb_p2p_alloc); the caller gathers all
 * ranks' handles (nranks * 64 bytes, rank order; halo_capacity = the largest interface entry count
 * of any rank, the same value on every rank) and, for each of this rank's neighbours q (in the
 * order given to semb_set_halo), the entry offset at which this rank's block starts in q's receive
 * layout (= the sum of q's nb_count over q's neighbours listed before this rank), and hands both to
 * semb_p2p_attach.  From then on the halo exchange is written straight into the neighbour's memory by
 * the gather kernel and the CG dots are summed by a one-CTA peer-memory all-reduce; NCCL stays as the
 * fallback (semb_set_tuning("p2p", 0)).  A wait for a peer gives up after 60 s: the call that
 * synchronises next returns an error instead of hanging the GPU. */
int semb_p2p_alloc(semb_ctx* ctx, int64_t halo_capacity, char* handle64);
int semb_p2p_attach(semb_ctx* ctx, const char* handles, const int64_t* peer_recv_off);

/* ---- device memory helpers for callers without their own allocator ------ */
int semb_malloc(int device, void** ptr, int64_t bytes);
int semb_free(int device, void* ptr);
int semb_memcpy(int device, void* dst, const void* src, int64_t bytes, int kind /*0 h2d, 1 d2h, 2 d2d*/);
/* Pinned host memory for result vectors (device -> host by direct DMA).  HOST
 * arguments in pageable memory are also accepted everywhere: they are staged
 * through a ring of pinned chunks by a few host threads. */
int semb_host_alloc(void** ptr, int64_t bytes);
int semb_host_free(void* ptr);

/* ---- measurement --------------------------------------------------------
 * With profiling on, every launch of the named kernels is bracketed by CUDA
 * events on the context's stream; semb_profile_get returns the summed device
 * time and the launch count since the last reset.  on = 1 brackets every
 * kernel kind; on = 1 | (mask << 1) only the kinds whose bit (1 << SEMB_K_*) is
 * set in mask (each bracket costs two event records on the stream). */
int semb_profile_enable(semb_ctx* ctx, int on);
int semb_profile_reset(semb_ctx* ctx);
int semb_profile_get(semb_ctx* ctx, int kernel, double* total_ms, int64_t* launches);
/* Kernels launched by this context since creation (all kinds). */
int64_t semb_launch_count(semb_ctx* ctx);
/* Process-wide tuning knobs for A/B measurements; results do not depend on
 * them.  "ax_l2_prefetch": 1 (default) / 0, bulk L2 prefetch of the geometric
 * factors in the tuned Ax kernels.  "p2p": 1 (default) / 0, peer-memory transport
 * vs NCCL for the halo exchange and the scalar all-reduces.  "gs_pull": 1 (default) / 0,
 * inside semb_cg the two-copy segments of the gather-scatter are summed by the CG
 * update kernel while it streams (only the longer segments keep an in-place pass)
 * vs one in-place pass over all segments.  "pdl": 1 (default) / 0, the kernels of
 * a CG iteration on one GPU are chained by programmatic dependent launch. */
int semb_set_tuning(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* SEMB200_H */
