// gsfm_ra.cu -- B200 (sm_100a) robust rotation averaging behind the C ABI of include/gsfm_ra.h.
//
// Replaces the Ceres solve inside GSfMNonlinearRotationEstimator
// (reference src/GSfM_nonlinear_rotation_estimator.cpp:24-80, 201-309): per-edge residuals,
// SO(3) Jacobians, covariance whitening and robust reweighting (K1), the block-3x3 normal
// equations, and a block-Jacobi PCG whose SpMV is K2 -- all resident in HBM, fp64 throughout.
//
// Data layout.  The view graph is stored as HALF-EDGES: every edge (i,j)
// appears once in row i and once in row j, sorted by (row, col); a row's half-edges are
// contiguous, so per-view sums (diagonal block, gradient) are segmented reductions over a
// contiguous range and need no atomics.  The half-edge array is cut into equal contiguous RANGES,
// one per resident warp of the kernel that walks it, and every range at row boundaries into
// SEGMENTS; records of 32 half-edges are planar (SoA) and travel by bulk async copy (TMA).  The
// normal-equation matrix lives in the BODY tangent frame (see so3_device.cuh): H = D^T Ht D with
// D = blockdiag(Jr(omega_i)), so the per-edge kernel never touches the per-view factors; the
// Euclidean (angle-axis) Levenberg-Marquardt of Ceres is reproduced exactly by transforming the
// LM diagonal per view.
//
// The same solver runs robust TRANSLATION averaging (include/gsfm_pa.h; reference
// src/GSfM_nonlinear_position_estimator.cpp): error type GSFM_RA_POSITION_BASELINE, D = I.
#include "ra_common.cuh"
#include "ra_structure.cuh"
#include "ra_edges.cuh"
#include "ra_pcg.cuh"
#include "ra_dense.cuh"
#include "ra_api_kernels.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int make_dev_loss(const gsfm_ra_loss* in, DevLoss* L) {
  std::memset(L, 0, sizeof(*L));
  if (!in) { L->kind = kLossTrivial; L->scale = 1.0; return 0; }
  if (in->kind < 0 || in->kind > GSFM_RA_LOSS_TABULATED) { set_error("unknown loss kind %d", in->kind); return GSFM_RA_ERR_INVALID; }
  // composition rho(s) = f(g(s)) (ComposedLoss, scripts/loss_functions.py:250-265): g is one of the closed-form kinds
  const double iscale = (in->inner_scale == 0.0) ? 1.0 : in->inner_scale;
  if (in->inner_kind != GSFM_RA_LOSS_TRIVIAL || iscale != 1.0) {
    if (in->inner_kind < 0 || in->inner_kind >= GSFM_RA_LOSS_MAGSAC3) {
      set_error("inner loss kind %d of a composition has no closed form on the device (tabulate the composed object instead)", in->inner_kind);
      return GSFM_RA_ERR_UNSUPPORTED;
    }
    if (in->kind == GSFM_RA_LOSS_TABULATED) { set_error("a tabulated loss cannot be composed: tabulate the composed object"); return GSFM_RA_ERR_UNSUPPORTED; }
    if (in->inner_kind != GSFM_RA_LOSS_TRIVIAL && !(in->inner_p[0] > 0.0)) { set_error("inner loss parameter p[0] must be > 0"); return GSFM_RA_ERR_INVALID; }
    if ((in->inner_kind == GSFM_RA_LOSS_TOLERANT || in->inner_kind == GSFM_RA_LOSS_GEMANMCCLURE) && !(in->inner_p[1] > 0.0)) {
      set_error("inner loss parameter p[1] must be > 0"); return GSFM_RA_ERR_INVALID;
    }
    L->composed = 1; L->inner_kind = in->inner_kind; L->ip0 = in->inner_p[0]; L->ip1 = in->inner_p[1]; L->iscale = iscale;
    L->isq0 = in->inner_p[0] * in->inner_p[0]; L->iinv_sq0 = 1.0 / L->isq0;
  }
  if (in->kind == GSFM_RA_LOSS_TABULATED) {
    const int M = in->table_per_octave;
    if (!in->table || in->table_octaves < 1 || M < 1 || M > 4096 || (M & (M - 1)) != 0) {
      set_error("tabulated loss: table must be non-NULL, octaves >= 1, per_octave a power of two <= 4096"); return GSFM_RA_ERR_INVALID;
    }
    int lg = 0;
    while ((1 << lg) < M) ++lg;
    L->kind = in->kind; L->flags = in->flags; L->scale = (in->scale == 0.0) ? 1.0 : in->scale;
    L->tab_min_exp = in->table_min_exp; L->tab_octaves = in->table_octaves; L->tab_log2_per_octave = lg;
    L->tab_rows = 2 + in->table_octaves * M;
    L->table = nullptr;  // device copy attached by attach_loss_table
    return 0;
  }
  L->kind = in->kind; L->flags = in->flags; L->p0 = in->p[0]; L->p1 = in->p[1];
  L->scale = (in->scale == 0.0) ? 1.0 : in->scale;
  L->sq0 = in->p[0] * in->p[0];
  L->inv_sq0 = 1.0 / L->sq0;
  const bool needs_p0 = in->kind != GSFM_RA_LOSS_TRIVIAL && in->kind != GSFM_RA_LOSS_TABULATED;
  if (needs_p0 && !(in->p[0] > 0.0)) { set_error("loss parameter p[0] must be > 0"); return GSFM_RA_ERR_INVALID; }
  if ((in->kind == GSFM_RA_LOSS_TOLERANT || in->kind == GSFM_RA_LOSS_GEMANMCCLURE) && !(in->p[1] > 0.0)) {
    set_error("loss parameter p[1] must be > 0"); return GSFM_RA_ERR_INVALID;
  }
  if (in->kind >= GSFM_RA_LOSS_MAGSAC3) {
    // include/gamma_values.cpp:6-11, 384-389, 780-785
    double C, quant, gk; int nu, n;
    if (in->kind == GSFM_RA_LOSS_MAGSAC3) { nu = 3; C = 4.029720004054876e-01; quant = 3.368214175218727; gk = 3.439485560754856e-03; n = 36843; }
    else if (in->kind == GSFM_RA_LOSS_MAGSAC4) { nu = 4; C = 2.525252525252525e-01; quant = 3.643721193503644e+00; gk = 3.611260617758625e-03; n = 38683; }
    else { nu = 9; C = 3.837828575290349e-03; quant = 4.654674460524809e+00; gk = 3.344206155099048e-02; n = 48553; }
    const double sigma = in->p[0];
    L->nu = nu; L->table_size = n;
    L->sq_sigma = sigma * sigma;
    L->sq_sigma_max_2 = 2.0 * L->sq_sigma;
    L->cubed_sigma = L->sq_sigma * sigma;
    L->clamp_s = quant * quant * L->sq_sigma;
    const double dof = (nu - 1.0) / 2.0;
    L->Ctd = C * std::pow(2.0, dof);
    L->one_over_sigma = L->Ctd / sigma;
    L->gamma_k = gk;
    L->weight_zero = L->one_over_sigma * (std::tgamma(dof) - gk);
    L->expo = nu / 2.0 - 1.5;
  }
  return 0;
}

// Device copy of a tabulated loss (rows of rho, rho', rho''); `buf` owns it.
template <typename Buf>
int attach_loss_table(const gsfm_ra_loss* in, DevLoss* L, Buf* buf, cudaStream_t st) {
  if (!in || in->kind != GSFM_RA_LOSS_TABULATED) return 0;
  const size_t n = 3 * (size_t)L->tab_rows;
  RA_TRY(buf->alloc(n));
  CUDA_TRY(cudaMemcpyAsync(buf->p, in->table, n * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));  // the caller's table may be freed as soon as the call returns
  L->table = buf->p;
  return 0;
}

bool type_needs_cov(int t) { return t == 3 || t == 6 || t == 7 || t == 8; }
bool type_is_position(int t) { return t == GSFM_RA_POSITION_BASELINE; }

int check_problem(const gsfm_ra_problem* p) {
  if (!p) { set_error("problem is NULL"); return GSFM_RA_ERR_INVALID; }
  if (p->num_views == 0 || p->num_edges == 0) { set_error("empty problem (views=%u edges=%llu)", p->num_views, (unsigned long long)p->num_edges); return GSFM_RA_ERR_INVALID; }
  if (!p->edge_i || !p->edge_j || !p->omega_ij) { set_error("edge arrays are NULL"); return GSFM_RA_ERR_INVALID; }
  if (p->num_views >= kSideBit) { set_error("too many views"); return GSFM_RA_ERR_INVALID; }
  if (p->num_edges >= (1ull << 31)) { set_error("too many edges for 32-bit half-edge offsets"); return GSFM_RA_ERR_UNSUPPORTED; }
  if (type_is_position(p->error_type)) {
    if (!p->orientation) { set_error("GSFM_RA_POSITION_BASELINE needs the global orientations"); return GSFM_RA_ERR_INVALID; }
    if (p->fixed_view >= (int64_t)p->num_views) { set_error("fixed_view %lld out of range", (long long)p->fixed_view); return GSFM_RA_ERR_INVALID; }
    return 0;
  }
  if (p->error_type < GSFM_RA_QUATERNION_NORM || p->error_type > GSFM_RA_ANGLE_AXIS_COVNORM) { set_error("unknown error_type %d", p->error_type); return GSFM_RA_ERR_INVALID; }
  if (type_needs_cov(p->error_type) && !p->cov6) { set_error("error_type %d needs cov6", p->error_type); return GSFM_RA_ERR_INVALID; }
  return 0;
}

int select_device(int device, int* out) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device visible: libgsfm_ra has no CPU fallback");
    return GSFM_RA_ERR_NO_DEVICE;
  }
  if (device < 0) { CUDA_TRY(cudaGetDevice(&device)); }
  if (device >= n) { set_error("device %d out of range (%d visible)", device, n); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(device));
  *out = device;
  return 0;
}

// Device buffers come from the device's stream-ordered memory pool (cudaMallocAsync) on the stream that
// is "current" for this thread while a solver is being built; the pool's release threshold is raised so a
// second gsfm_ra_solve() call reuses the memory of the first instead of paying cudaMalloc/cudaFree again.
thread_local cudaStream_t g_alloc_stream = nullptr;
thread_local bool g_alloc_async = false;

struct AllocScope {
  cudaStream_t prev_s;
  bool prev_a;
  explicit AllocScope(cudaStream_t st) : prev_s(g_alloc_stream), prev_a(g_alloc_async) { g_alloc_stream = st; g_alloc_async = true; }
  ~AllocScope() { g_alloc_stream = prev_s; g_alloc_async = prev_a; }
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool async = false;
  cudaStream_t st = nullptr;
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return 0;
    if (g_alloc_async) {
      async = true; st = g_alloc_stream;
      CUDA_TRY(cudaMallocAsync(&p, count * sizeof(T), st));
    } else {
      async = false;
      CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    }
    return 0;
  }
  void release() {
    if (p) { if (async) cudaFreeAsync(p, st); else cudaFree(p); }
    p = nullptr; n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// Owns the solver's stream; declared FIRST in the solver so it is destroyed LAST, after every DevBuf has
// queued its cudaFreeAsync on it.
struct StreamHolder {
  cudaStream_t s = nullptr;
  ~StreamHolder() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};

// Per-device facts that are expensive to query: cached for the life of the process.
struct DeviceInfo {
  bool ready = false;
  int sm_count = 0, coop = 0, occ_k2[3] = {1, 1, 1};  // resident blocks of the persistent PCG kernel: [0] 4-, [1] 6-, [2] 9-double records
  int l2_bytes = 0;
};
DeviceInfo g_device_info[64];

inline unsigned grid_for(uint64_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

}  // namespace

// ------------------------------------------------------------------------------------------
// the resident solver
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so the single-GPU library has no link dependency on it.  When the
// host process already loaded a libnccl.so.2 (e.g. PyTorch's bundled one) the same copy is reused.
// ------------------------------------------------------------------------------------------
namespace ncclx {
constexpr int kUniqueIdBytes = 128;
struct UniqueId { char internal[kUniqueIdBytes]; };
typedef void* Comm;
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(Comm*, int, UniqueId, int);
typedef int (*AllReduce_t)(const void*, void*, size_t, int /*dtype*/, int /*op*/, Comm, cudaStream_t);
typedef int (*CommDestroy_t)(Comm);
typedef const char* (*GetErrorString_t)(int);
constexpr int kFloat64 = 8;  // ncclDouble
constexpr int kSum = 0;      // ncclSum
struct Api {
  void* handle = nullptr;
  GetUniqueId_t GetUniqueId = nullptr;
  CommInitRank_t CommInitRank = nullptr;
  AllReduce_t AllReduce = nullptr;
  CommDestroy_t CommDestroy = nullptr;
  GetErrorString_t GetErrorString = nullptr;
};
Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* env = std::getenv("GSFM_RA_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (a.handle) break;
    }
    if (a.handle) {
      a.GetUniqueId = (GetUniqueId_t)dlsym(a.handle, "ncclGetUniqueId");
      a.CommInitRank = (CommInitRank_t)dlsym(a.handle, "ncclCommInitRank");
      a.AllReduce = (AllReduce_t)dlsym(a.handle, "ncclAllReduce");
      a.CommDestroy = (CommDestroy_t)dlsym(a.handle, "ncclCommDestroy");
      a.GetErrorString = (GetErrorString_t)dlsym(a.handle, "ncclGetErrorString");
      if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) a.handle = nullptr;
    }
  }
  return a.handle ? &a : nullptr;
}
}  // namespace ncclx

#define NCCL_TRY(expr)                                                                                     \
  do {                                                                                                     \
    int r__ = (expr);                                                                                      \
    if (r__ != 0) {                                                                                        \
      set_error("%s failed: %s", #expr, ncclx::api()->GetErrorString ? ncclx::api()->GetErrorString(r__) : "?"); \
      return GSFM_RA_ERR_CUDA;                                                                             \
    }                                                                                                      \
  } while (0)

// A balanced work partition of the half-edge array for one kernel class: num_warps equal
// contiguous ranges (a multiple of 32 half-edges each), every range cut at row boundaries into
// segments.  num_warps = SMs x resident blocks x warps per block of THAT kernel, so the kernel runs
// as exactly one full wave.
struct Partition {
  uint32_t num_warps = 0, num_segs = 0, grid = 0, span = 0;  // span = half-edges per warp (multiple of 32)
  DevBuf<uint32_t> warp_seg_ptr, seg_row, seg_begin, seg_len, node_seg_ptr;
};

struct gsfm_ra_solver {
  StreamHolder stream_holder;  // first member: destroyed last
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t N = 0;
  uint64_t E = 0;       // edges of this shard
  uint64_t edge_begin = 0;
  uint64_t H = 0;       // half-edges of this shard (2E)
  int rank = 0, world = 1;
  int error_type = 4;
  int blk = 6;          // doubles per stored off-diagonal block (4: compact scalar-weight stencil, 6: symmetric Laplacian stencil, 9: general)
  bool scalar_u = false;  // the whitening factor is a scalar multiple of the identity (types 2, 4, 5, 7, 8)
  gsfm_ra_options opt;
  DevLoss loss;
  int64_t launches = 0;
  int64_t linear_unconverged = 0;  // PCG solves that hit pcg_max_iterations above pcg_rtol
  bool cooperative = true;  // persistent PCG kernel available
  ncclx::Comm comm = nullptr;  // edge-sharded exchange (world > 1)
  // fused exchange: one cudaMalloc'ed block of LLCell per rank (layout in ra_common.cuh), every peer's block mapped here
  LLCell* xchg = nullptr;
  void* peer_base[kMaxPeers] = {};
  bool peers_connected = false;
  bool peers_ipc = false;    // peer_base[] came from cudaIpcOpenMemHandle (closed on destruction), not from in-process peer access

  // structure
  DevBuf<uint32_t> he_col, he_row, he_edge, iso;
  uint32_t n_iso = 0;
  ColBlocks cbk = {};    // column blocks of the half-edge order (ra_structure.cuh)
  bool slice = false;    // the persistent PCG kernel stages a block's slice of z in shared memory
  Partition pk1, pk2;  // K1 (edge kernel) and K2 (SpMV / PCG) partitions
  // per half-edge constants, planar
  DevBuf<double> loss_table;  // device copy of a tabulated loss (GSFM_RA_LOSS_TABULATED)
  DevBuf<double> inrec;  // K1's input records (k_setup_halfedges)
  int ku = 1;            // whitening entries per half-edge in the input records: 1 (scalar) or 6 (upper triangle)
  // edge-order copies for the API kernels
  DevBuf<uint32_t> d_ei, d_ej;
  DevBuf<double> d_omega_ij, d_cov6, d_weight, d_orient;  // d_orient: translation averaging, the global orientations [N][3]
  uint32_t fixed_view = kNoFixedView;                     // translation averaging: the view held constant
  // linearisation, double buffered: [cur] is the accepted point, [cur^1] the candidate
  DevBuf<double> omega[2], node_q[2], node_JL[2], val[2], ediag[2];
  // lin[b] = [Hd 6N | gt 3N | cost, bad]: everything one evaluation sums over edges, contiguous so the
  // edge-sharded solver reduces it across GPUs with ONE all-reduce
  DevBuf<double> lin[2];
  double* Hd_p[2] = {nullptr, nullptr};
  double* gt_p[2] = {nullptr, nullptr};
  DevBuf<double> part;
  int cur = 0;
  // PCG
  DevBuf<double> scale, Dblk, Minv, x, r, z, p, q, sv, bvec, y, ypart, delta, dense_A, dense_work, dense_winv;  // z, p: stride 4
  DevBuf<double> slots;
  DevBuf<unsigned long long> bar_slots;  // grid barrier + reduction slots of the persistent PCG kernel (see grid_bar_sum2)
  bool fused_step = false;               // set while a trust-region batch is enqueued: PCG kernel runs prologue + epilogue
  uint32_t keep8 = 0;                    // L2 residency of the matrix stream, see l2_policy_evict_last
  DevBuf<unsigned> counter, row_cnt;
  DevBuf<DevScalars> sc;
  DevScalars* h_sc = nullptr;  // pinned
  HostMailbox* mailbox = nullptr;      // pinned + mapped; device alias below
  HostMailbox* mailbox_dev = nullptr;
  unsigned mailbox_seq = 0;
  // One trust-region batch as a CUDA graph (one per linearisation buffer): replayed with a single launch, its kernels
  // run back to back instead of waiting for the host to enqueue them one by one.  What varies between replays lives in
  // it_params (refreshed by the graph's first node from h_it_params).
  DevBuf<IterParams> it_params;
  IterParams* h_it_params = nullptr;   // pinned
  cudaGraphExec_t batch_graph[2] = {nullptr, nullptr};
  int64_t batch_launches = 0;
  bool graph_params = false;           // true while a batch is being captured: kernels read mu / seq from it_params
  int graph_state = 0;                 // 0 untried, 1 in use, -1 unavailable (direct launches)
  unsigned long long* prof_buf = nullptr;  // device, set only by gsfm_ra_solver_time_kernels

  // trust-region state (host)
  bool linearized = false;
  bool scale_ready = false;
  double radius = 1e4, decrease_factor = 2.0, x_cost = 0.0, x_norm = 0.0, gmax = 0.0;
  int iteration = 0, invalid_steps = 0;
  bool last_successful = false;
  int termination = GSFM_RA_TERM_NONE;
  double initial_cost = 0.0;
  // timing accumulators (ms, CUDA events)
  double ms_setup = 0, ms_assemble = 0, ms_linear = 0, ms_cost = 0;

  ~gsfm_ra_solver() {
    if (comm && ncclx::api()) ncclx::api()->CommDestroy(comm);
    for (int r = 0; r < kMaxPeers; ++r) if (peers_ipc && peer_base[r] && r != rank) cudaIpcCloseMemHandle(peer_base[r]);
    if (xchg) cudaFree(xchg);
    if (h_sc) cudaFreeHost(h_sc);
    if (mailbox) cudaFreeHost(mailbox);
    if (h_it_params) cudaFreeHost(h_it_params);
    for (auto& g : batch_graph) if (g) cudaGraphExecDestroy(g);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    // the stream itself is destroyed by stream_holder after the buffers have been returned to the pool
  }

  bool sharded() const { return world > 1; }
  // exchange mode (ra_common.cuh): direct up to 2 ranks, owner-reduce above; GSFM_RA_EXCHANGE=owner / direct forces one (tests)
  bool owner_mode() const {
    const char* e = std::getenv("GSFM_RA_EXCHANGE");
    if (e && std::strcmp(e, "owner") == 0) return world > 1;
    if (e && std::strcmp(e, "direct") == 0) return false;
    return world > 2;
  }
  // QUATERNION_COSINE: parameters live on the manifold (left-multiplicative update, local coordinates delta = phi/2)
  bool manifold() const { return error_type <= GSFM_RA_QUATERNION_COSINE; }
  bool position() const { return type_is_position(error_type); }   // translation averaging: Euclidean 3-vectors, D = I
  int param_kind() const { return position() ? 2 : (manifold() ? 1 : 0); }  // node_prep_view / apply_view
  bool general() const { return error_type < GSFM_RA_QUATERNION_COSINE; }  // two-block residuals, 9-double records
  int allreduce(double* buf, size_t count) {
    if (!comm) { set_error("sharded solver used before gsfm_ra_solver_comm_init"); return GSFM_RA_ERR_INVALID; }
    NCCL_TRY(ncclx::api()->AllReduce(buf, buf, count, ncclx::kFloat64, ncclx::kSum, comm, stream));
    return 0;
  }

  int fetch_scalars() {
    CUDA_TRY(cudaMemcpyAsync(h_sc, sc.p, sizeof(DevScalars), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
  }

  // Wait for the mailbox of the batch published as `seq` (see HostMailbox) and take its scalars.  Polls host memory;
  // cudaStreamQuery every few thousand polls catches a failed launch, and a finished stream without the flag is an error.
  int fetch_mailbox(unsigned seq) {
    unsigned spins = 0;
    while (mailbox->seq != seq) {
      if ((++spins & 0xfffu) == 0) {
        const cudaError_t q = cudaStreamQuery(stream);
        if (q == cudaSuccess) {
          if (mailbox->seq == seq) break;
          set_error("trust-region batch finished without publishing its scalars");
          return GSFM_RA_ERR_CUDA;
        }
        if (q != cudaErrorNotReady) { set_error("CUDA failure while waiting for a trust-region batch: %s", cudaGetErrorString(q)); return GSFM_RA_ERR_CUDA; }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    std::memcpy(h_sc, const_cast<DevScalars*>(&mailbox->sc), sizeof(DevScalars));
    return 0;
  }
  // kernel times of the batch whose scalars were just fetched (%globaltimer stamps written by its kernels)
  void account_batch() {
    if (h_sc->t_linear_end >= h_sc->t_begin) ms_linear += 1e-6 * (double)(h_sc->t_linear_end - h_sc->t_begin);
    if (h_sc->t_end >= h_sc->t_linear_end) ms_assemble += 1e-6 * (double)(h_sc->t_end - h_sc->t_linear_end);
  }

  int rec_doubles() const { return blk * 32 + 16; }
  int smem_bytes() const { return spmv_smem_bytes(blk); }
  // persistent PCG kernel: + the shared-memory copy of z for small graphs (GSFM_RA_NO_SLICE=1 disables it, A/B runs)
  int pcg_smem_bytes() const { return spmv_smem_bytes(blk) + (slice ? (int)(24u * cbk.cbsize + 16u) : 0); }
  // K1 is specialised on (Jacobian?, residual kind, scalar weight?, loss): the common losses get their own instantiation
  // (no switch, fewer registers), everything else runs the generic one.
  typedef void (*K1Fn)(const K1Args);
  template <bool JAC, int RES, bool SCAL>
  K1Fn pick_k1_loss() const {
    const bool plain = loss.scale == 1.0 && !loss.composed;
    if (plain && loss.kind == kLossCauchy) return k_edges<JAC, RES, SCAL, kLossCauchy>;
    if (plain && loss.kind == kLossSoftLOne) return k_edges<JAC, RES, SCAL, kLossSoftLOne>;
    if (plain && loss.kind == kLossHuber) return k_edges<JAC, RES, SCAL, kLossHuber>;
    if (plain && loss.kind == kLossMagsac3) return k_edges<JAC, RES, SCAL, kLossMagsac3>;
    return k_edges<JAC, RES, SCAL, -1>;
  }
  K1Fn pick_k1(bool jacobian) const {
    if (manifold()) return jacobian ? pick_k1_loss<true, 1, true>() : pick_k1_loss<false, 1, true>();
    if (position()) return jacobian ? pick_k1_loss<true, 2, true>() : pick_k1_loss<false, 2, true>();
    if (scalar_u) return jacobian ? pick_k1_loss<true, 0, true>() : pick_k1_loss<false, 0, true>();
    return jacobian ? pick_k1_loss<true, 0, false>() : pick_k1_loss<false, 0, false>();
  }
  int k1_smem_bytes() const { return kWarpsPerBlock * kStages * in_rec_doubles(ku) * 8 + kWarpsPerBlock * kStages * 8; }
  template <bool JAC, int TYPE>
  void launch_edges_g(int b, double* val_out) {
    k_edges_general<JAC, TYPE><<<pk1.grid, kBlock, 0, stream>>>(pk1.num_warps, H, pk1.warp_seg_ptr.p, pk1.seg_row.p, pk1.seg_begin.p, pk1.seg_len.p,
                                                                he_col.p, inrec.p, node_q[b].p, loss, val_out, part.p);
  }
  void launch_edges(int b, bool jacobian, double* val_out) {
    if (error_type == GSFM_RA_QUATERNION_NORM) { if (jacobian) launch_edges_g<true, 0>(b, val_out); else launch_edges_g<false, 0>(b, nullptr); return; }
    if (error_type == GSFM_RA_ROTATION_MAT_FNORM) { if (jacobian) launch_edges_g<true, 1>(b, val_out); else launch_edges_g<false, 1>(b, nullptr); return; }
    K1Args A;
    A.num_warps = pk1.num_warps; A.warp_span = pk1.span; A.H = H;
    A.warp_seg_ptr = pk1.warp_seg_ptr.p; A.seg_begin = pk1.seg_begin.p; A.seg_len = pk1.seg_len.p;
    A.inrec = inrec.p; A.node_q = node_q[b].p; A.val = val_out; A.part = part.p; A.loss = loss; A.fixed = fixed_view;
    pick_k1(jacobian)<<<pk1.grid, kBlock, k1_smem_bytes(), stream>>>(A);
  }
  void launch_spmv(int b, const double* x4, int check_done) {
    if (blk == 4)
      k_spmv<4><<<pk2.grid, kPcgBlock, smem_bytes(), stream>>>(pk2.num_warps, H, pk2.span, pk2.warp_seg_ptr.p, pk2.seg_begin.p, pk2.seg_len.p, val[b].p,
                                                            x4, ypart.p, sc.p, check_done, keep8);
    else if (blk == 6)
      k_spmv<6><<<pk2.grid, kPcgBlock, smem_bytes(), stream>>>(pk2.num_warps, H, pk2.span, pk2.warp_seg_ptr.p, pk2.seg_begin.p, pk2.seg_len.p, val[b].p,
                                                            x4, ypart.p, sc.p, check_done, keep8);
    else
      k_spmv<9><<<pk2.grid, kPcgBlock, smem_bytes(), stream>>>(pk2.num_warps, H, pk2.span, pk2.warp_seg_ptr.p, pk2.seg_begin.p, pk2.seg_len.p, val[b].p,
                                                            x4, ypart.p, sc.p, check_done, keep8);
  }

  // ---- evaluation at omega[b]: node prep, K1 (or K1c), node finalize -------------------------
  int evaluate(int b, bool jacobian, bool publish = false, bool prepped = false) {
    HostMailbox* mb = publish ? mailbox_dev : nullptr;
    const IterParams* ip_dev = graph_params ? it_params.p : nullptr;  // graph replay: the sequence number comes from device memory
    const unsigned mseq = (publish && !graph_params) ? ++mailbox_seq : 0u;
    if (!prepped) {
      k_node_prep<<<grid_for(N), kBlock, 0, stream>>>(N, omega[b].p, node_q[b].p, node_JL[b].p, slots.p, counter.p, sc.p, param_kind());
      launches += 1;
    }
    launch_edges(b, jacobian, val[b].p);
    double* tail = lin[b].p + 9ull * N;
    const int co = jacobian ? 0 : 1;
    if (!sharded()) {
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 0, tail,
                                                           slots.p, counter.p, sc.p, mb, mseq, publish ? ip_dev : nullptr);
    } else if (peers_connected && jacobian) {
      // edge-sharded, fused exchange: the reduction of [Hd | gt | cost] across GPUs happens inside the kernel (peer memory)
      PeerPtrs pp;
      for (int r = 0; r < kMaxPeers; ++r) pp.p[r] = (LLCell*)peer_base[r];
      k_node_finalize_ll<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, slots.p, counter.p,
                                                              sc.p, mb, mseq, publish ? ip_dev : nullptr, pp, world, rank, owner_mode() ? 1 : 0);
    } else {
      // edge-sharded, NCCL: local sums -> ONE all-reduce of [Hd | gt | cost, bad] -> per-view post-processing
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 1, tail,
                                                           slots.p, counter.p, sc.p, nullptr, 0u, nullptr);
      if (jacobian) RA_TRY(allreduce(lin[b].p, 9ull * N + 2));
      else RA_TRY(allreduce(tail, 2));
      k_node_finalize<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk1.node_seg_ptr.p, part.p, node_JL[b].p, Hd_p[b], gt_p[b], ediag[b].p, co, 2, tail,
                                                           slots.p, counter.p, sc.p, mb, mseq, publish ? ip_dev : nullptr);
      launches += 2;
    }
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }

  // y = (Ht offdiag + diag_blocks) x on linearisation b (separate-kernel path); xin has stride 4, yout stride 3.
  int spmv(int b, const double* xin, double* yout, const double* diag_blocks) {
    launch_spmv(b, xin, 0);
    if (!sharded()) {
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, diag_blocks, xin, yout, nullptr, 1, slots.p, counter.p, sc.p);
    } else {
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, nullptr, xin, yout, nullptr, 1, slots.p, counter.p, sc.p);
      RA_TRY(allreduce(yout, 3ull * N));
      k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, diag_blocks, xin, yout, yout, 1, slots.p, counter.p, sc.p);
      launches += 2;
    }
    launches += 2;
    CUDA_TRY(cudaGetLastError());
    return 0;
  }

  PrepareArgs prepare_args(int b, double mu, const double* user_damp, const double* user_b) {
    PrepareArgs A;
    A.mu = mu; A.lo = opt.min_lm_diagonal; A.hi = opt.max_lm_diagonal;
    A.ediag = ediag[b].p; A.scale = scale.p; A.node_JL = node_JL[b].p; A.Hd = Hd_p[b]; A.gt = gt_p[b]; A.user_damp = user_damp; A.user_b = user_b;
    A.Dblk = Dblk.p; A.Minv = Minv.p; A.x = x.p; A.r = r.p; A.z = z.p; A.p = p.p; A.q = q.p; A.bvec = bvec.p;
    return A;
  }
  ApplyArgs apply_args(int b, int c) {
    ApplyArgs A;
    A.node_JL = node_JL[b].p; A.xt = x.p; A.bvec = bvec.p; A.res = r.p; A.Dblk = Dblk.p; A.Hd = Hd_p[b]; A.gt = gt_p[b]; A.omega = omega[b].p;
    A.cand = omega[c].p; A.delta_out = delta.p; A.manifold = param_kind();
    return A;
  }
  PcgParams pcg_params(int b, double rtol, int max_iter) {
    PcgParams P;
    P.N = N; P.num_warps = pk2.num_warps; P.n_iso = n_iso; P.max_iter = max_iter; P.H = H; P.rtol2 = rtol * rtol;
    P.keep8 = keep8;
    P.cbk = cbk;
    if (!slice) P.cbk.ncb = 0;
    P.warp_seg_ptr = pk2.warp_seg_ptr.p; P.seg_row = pk2.seg_row.p; P.seg_begin = pk2.seg_begin.p; P.seg_len = pk2.seg_len.p;
    P.node_seg_ptr = pk2.node_seg_ptr.p; P.iso = iso.p; P.warp_span = pk2.span;
    P.val = val[b].p; P.Dblk = Dblk.p; P.Minv = Minv.p;
    P.x = x.p; P.r = r.p; P.z = z.p; P.p = p.p; P.q = q.p; P.s = sv.p; P.ypart = ypart.p;
    P.row_cnt = row_cnt.p; P.bar_slots = bar_slots.p; P.slots = slots.p; P.counter = counter.p; P.sc = sc.p; P.prof = prof_buf;
    P.fused = 0; P.ip = nullptr; P.cand_q = nullptr; P.cand_JL = nullptr;
    std::memset(&P.prep, 0, sizeof(P.prep)); std::memset(&P.apply, 0, sizeof(P.apply));
    P.world = peers_connected ? world : 1; P.rank = rank;
    P.owner_mode = owner_mode() ? 1 : 0;
    for (int r = 0; r < kMaxPeers; ++r) P.peer[r] = (LLCell*)peer_base[r];
    return P;
  }

  // Block-Jacobi PCG on (Ht + Lam) xt = bt for linearisation b; on return (stream order) x = xt, r = the
  // residual bt - (Ht + Lam) xt and bvec = bt.  One cooperative launch; no host synchronisation.
  int pcg_enqueue(int b, double mu, const double* user_damp, const double* user_b, double rtol, int max_iter) {
    const bool fuse = fused_step && !user_damp && !user_b;  // a trust-region batch: prologue and epilogue live in the PCG kernel
    if (!fuse) {
      k_prepare_solve<<<grid_for(N), kBlock, 0, stream>>>(N, prepare_args(b, mu, user_damp, user_b), slots.p, counter.p, sc.p,
                                                           graph_params ? it_params.p : nullptr);
      launches += 1;
    }
    if (opt.linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY) {
      const uint32_t n = 3 * N, np = (n + 1 + kNB - 1) / kNB * kNB;  // room for the right-hand-side row
      if (dense_A.n < (size_t)np * np) {
        AllocScope scope(stream);
        RA_TRY(dense_A.alloc((size_t)np * np));
        RA_TRY(dense_work.alloc(np));
        RA_TRY(dense_winv.alloc((size_t)np * kNB));
      }
      CUDA_TRY(cudaMemsetAsync(dense_A.p, 0, (size_t)np * np * sizeof(double), stream));
      k_dense_assemble<<<grid_for(std::max<uint64_t>(H, np)), kBlock, 0, stream>>>(H, N, np, blk, he_row.p, he_col.p, val[b].p, Dblk.p, r.p, dense_A.p);
      uint32_t n_arg = n, np_arg = np;
      double* A_arg = dense_A.p; double* x_arg = x.p; double* wi_arg = dense_winv.p; double* w_arg = dense_work.p; DevScalars* sc_arg = sc.p;
      void* args[] = {&n_arg, &np_arg, &A_arg, &x_arg, &wi_arg, &w_arg, &sc_arg};
      // all SMs: the trailing update of panel kb has (nblk-kb)(nblk-kb-1)/2 tiles to spread
      int occ = 1;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dense_cholesky_solve, kBlock, 0));
      const unsigned grid = (unsigned)(sm_count * std::max(1, std::min(occ, 2)));
      CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_dense_cholesky_solve, dim3(grid), dim3(kBlock), args, 0, stream));
      CUDA_TRY(cudaMemsetAsync(r.p, 0, 3ull * N * sizeof(double), stream));  // exact solve: residual 0
      launches += 2;
      return 0;
    }
    if (cooperative && (!sharded() || peers_connected)) {
      PcgParams P = pcg_params(b, rtol, max_iter);
      P.fused = fuse ? 1 : 0;
      if (fuse) {
        P.prep = prepare_args(b, mu, nullptr, nullptr);
        P.apply = apply_args(b, b ^ 1);
        P.cand_q = node_q[b ^ 1].p; P.cand_JL = node_JL[b ^ 1].p;
        P.ip = graph_params ? it_params.p : nullptr;
      }
      void* args[] = {&P};
      const void* fn = (blk == 4) ? (const void*)k_pcg_persistent<4> : (blk == 6) ? (const void*)k_pcg_persistent<6> : (const void*)k_pcg_persistent<9>;
      CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(pk2.grid), dim3(kPcgBlock), args, pcg_smem_bytes(), stream));
      launches += 1;
      return 0;
    }
    // fallback (device without cooperative launch, or NCCL exchange): textbook PCG from separate kernels, polled
    const int poll = 8;
    const double rtol2 = rtol * rtol;
    int enq = 0;
    while (true) {
      for (int k = 0; k < poll && enq < max_iter; ++k, ++enq) {
        launch_spmv(b, p.p, 1);
        if (!sharded()) {
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, Dblk.p, p.p, y.p, nullptr, 0, slots.p, counter.p, sc.p);
        } else {
          // the ONE collective of a CG step: all-reduce of the 3N partial matvec
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, nullptr, p.p, y.p, nullptr, 1, slots.p, counter.p, sc.p);
          RA_TRY(allreduce(y.p, 3ull * N));
          k_spmv_finish<<<grid_for(N), kBlock, 0, stream>>>(N, cbk.ncb, pk2.node_seg_ptr.p, ypart.p, Dblk.p, p.p, y.p, y.p, 0, slots.p, counter.p, sc.p);
          launches += 2;
        }
        k_pcg_update<<<grid_for(N), kBlock, 0, stream>>>(N, Minv.p, p.p, y.p, x.p, r.p, z.p, rtol2, max_iter, slots.p, counter.p, sc.p);
        k_pcg_direction<<<grid_for(4ull * N), kBlock, 0, stream>>>(4 * N, z.p, p.p, sc.p);
        launches += 4;
      }
      CUDA_TRY(cudaGetLastError());
      RA_TRY(fetch_scalars());
      if (h_sc->pcg_done || enq >= max_iter) break;
    }
    return 0;
  }

  double elapsed(cudaEvent_t a, cudaEvent_t b2) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b2);
    return ms;
  }
  double elapsed_since(cudaEvent_t a) {
    cudaEventRecord(ev[1], stream);
    cudaEventSynchronize(ev[1]);
    return elapsed(a, ev[1]);
  }
};

namespace {

int device_info(int device, DeviceInfo** out) {
  DeviceInfo& d = g_device_info[device & 63];
  if (!d.ready) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&d.coop, cudaDevAttrCooperativeLaunch, device));
    CUDA_TRY(cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, device));
    CUDA_TRY(cudaFuncSetAttribute(k_pcg_persistent<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(4) + 24 * kSliceMaxViews + 16));
    CUDA_TRY(cudaFuncSetAttribute(k_pcg_persistent<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(6) + 24 * kSliceMaxViews + 16));
    CUDA_TRY(cudaFuncSetAttribute(k_pcg_persistent<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(9) + 24 * kSliceMaxViews + 16));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(4)));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(6)));
    CUDA_TRY(cudaFuncSetAttribute(k_spmv<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, spmv_smem_bytes(9)));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_k2[0], k_pcg_persistent<4>, kPcgBlock, spmv_smem_bytes(4) + 24 * kSliceMaxViews + 16));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_k2[1], k_pcg_persistent<6>, kPcgBlock, spmv_smem_bytes(6) + 24 * kSliceMaxViews + 16));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_k2[2], k_pcg_persistent<9>, kPcgBlock, spmv_smem_bytes(9) + 24 * kSliceMaxViews + 16));
    // keep freed blocks in the pool: the next solver reuses them
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    d.ready = true;
  }
  *out = &d;
  return 0;
}

template <typename T>
int exclusive_scan(const T* in, T* out, size_t n, cudaStream_t st) {
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, st));
  DevBuf<unsigned char> tmp;
  RA_TRY(tmp.alloc(bytes + 16));
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, in, out, (int)n, st));
  return 0;
}

// Upload the (shard of the) problem and build every device structure; one host sync at the end.
int build_solver(const gsfm_ra_problem* prob, const gsfm_ra_options* options, int rank, int world, gsfm_ra_solver** out) {
  RA_TRY(check_problem(prob));
  if (!options || !out) { set_error("options/out is NULL"); return GSFM_RA_ERR_INVALID; }
  if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank/world"); return GSFM_RA_ERR_INVALID; }
  const uint32_t N = prob->num_views;
  const double tb0 = now_ms();
  double tb_prev = tb0;
  auto lap = [&](const char* what) {
    if (options->verbose >= 2) { const double t = now_ms(); std::fprintf(stderr, "[gsfm_ra] setup %-28s %8.2f ms\n", what, t - tb_prev); tb_prev = t; }
  };
  int device = 0;
  RA_TRY(select_device(options->device, &device));
  DeviceInfo* di = nullptr;
  RA_TRY(device_info(device, &di));
  std::unique_ptr<gsfm_ra_solver> s(new gsfm_ra_solver());
  s->device = device;
  s->opt = *options;
  s->rank = rank; s->world = world;
  s->error_type = prob->error_type;
  s->scalar_u = !(prob->error_type == GSFM_RA_ANGLE_AXIS_COVARIANCE || prob->error_type == GSFM_RA_ANGLE_AXIS_COV_INLIERS);
  s->ku = s->scalar_u ? 1 : 6;
  // scalar-weight angle-axis stencils (types 4, 5, 7, 8) are stored in the compact 4-double form (-DGSFM_RA_NO_COMPACT: 6)
  s->blk = s->general() ? 9 : (kCompactScalarStencil && s->scalar_u && !s->manifold() && !s->position()) ? 4 : 6;
  if (s->position() && prob->fixed_view >= 0) s->fixed_view = (uint32_t)prob->fixed_view;
  RA_TRY(make_dev_loss(&options->loss, &s->loss));
  s->sm_count = di->sm_count;
  CUDA_TRY(cudaStreamCreateWithFlags(&s->stream_holder.s, cudaStreamNonBlocking));
  s->stream = s->stream_holder.s;
  AllocScope alloc_scope(s->stream);
  RA_TRY(attach_loss_table(&options->loss, &s->loss, &s->loss_table, s->stream));
  for (auto& e : s->ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDefault));
  CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
  lap("context/stream");

  // shard: a contiguous range of the caller's edge list
  const uint64_t e0 = prob->num_edges * (uint64_t)rank / world, e1 = prob->num_edges * (uint64_t)(rank + 1) / world;
  const uint64_t E = e1 - e0, H = 2 * E;
  if (E == 0) { set_error("rank %d of %d owns no edges", rank, world); return GSFM_RA_ERR_INVALID; }
  s->N = N; s->E = E; s->H = H; s->edge_begin = e0;
  cudaStream_t st = s->stream;

  auto up32 = [&](DevBuf<uint32_t>& d, const uint32_t* src, size_t n) -> int {
    RA_TRY(d.alloc(n));
    if (n == 0) return 0;
    CUDA_TRY(cudaMemcpyAsync(d.p, src, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    return 0;
  };
  RA_TRY(up32(s->d_ei, prob->edge_i + e0, E));
  RA_TRY(up32(s->d_ej, prob->edge_j + e0, E));
  // The structure build below needs only the two index arrays.  The per-edge PAYLOAD (measurements, covariances, weights: 24 to
  // 80 bytes per edge against 8 of indices) is uploaded on a second stream while the keys are sorted and the partitions built;
  // K0, its first reader, waits for it.  (GSFM_RA_NO_COPY_STREAM=1: everything on the solver's stream, A/B runs.)
  struct CopyStream {
    cudaStream_t s = nullptr;
    cudaEvent_t ready = nullptr, done = nullptr;
    ~CopyStream() {
      if (s) cudaStreamSynchronize(s);
      if (ready) cudaEventDestroy(ready);
      if (done) cudaEventDestroy(done);
      if (s) cudaStreamDestroy(s);
    }
  } cs;
  const bool overlap_upload = std::getenv("GSFM_RA_NO_COPY_STREAM") == nullptr;
  cudaStream_t up = st;
  if (overlap_upload) {
    CUDA_TRY(cudaStreamCreateWithFlags(&cs.s, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&cs.ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&cs.done, cudaEventDisableTiming));
    up = cs.s;
  }
  RA_TRY(s->d_omega_ij.alloc(3 * E));
  if (prob->cov6) RA_TRY(s->d_cov6.alloc(6 * E));
  if (prob->edge_weight) RA_TRY(s->d_weight.alloc(E));
  if (s->position()) RA_TRY(s->d_orient.alloc(3ull * N));
  if (overlap_upload) {  // the buffers come from the pool in the order of the solver's stream
    CUDA_TRY(cudaEventRecord(cs.ready, st));
    CUDA_TRY(cudaStreamWaitEvent(up, cs.ready, 0));
  }
  CUDA_TRY(cudaMemcpyAsync(s->d_omega_ij.p, prob->omega_ij + 3 * e0, 3 * E * sizeof(double), cudaMemcpyHostToDevice, up));
  if (prob->cov6) CUDA_TRY(cudaMemcpyAsync(s->d_cov6.p, prob->cov6 + 6 * e0, 6 * E * sizeof(double), cudaMemcpyHostToDevice, up));
  if (prob->edge_weight) CUDA_TRY(cudaMemcpyAsync(s->d_weight.p, prob->edge_weight + e0, E * sizeof(double), cudaMemcpyHostToDevice, up));
  if (s->position()) CUDA_TRY(cudaMemcpyAsync(s->d_orient.p, prob->orientation, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, up));
  if (overlap_upload) CUDA_TRY(cudaEventRecord(cs.done, up));
  lap("enqueue uploads");

  // ---- half-edges sorted by (column block, row, col): keys -> radix sort -> unpack (ra_structure.cuh) ---------------
  // column blocks: as few as keep a block's slice of the gathered vector within the shared-memory budget; none (one block,
  // plain row-major order, gather from L2) beyond kMaxColBlocks * kSliceMaxViews views or with GSFM_RA_NO_SLICE=1
  {
    // Measured (profiles/r02_col_blocks.txt): at 10k views / 1M edges four column blocks cut the L2 gather traffic but quadruple
    // the segments (row sums to reduce, publish and collect) and the CG step goes from 31 to 49 us, K1 from 60 to 69 us -- so
    // several blocks are an experiment switch (GSFM_RA_COL_BLOCKS=n), and the default is ONE block: graphs of <= kSliceMaxViews
    // views gather from shared memory, larger ones from L2.
    uint32_t ncb = 1;
    if (const char* e = std::getenv("GSFM_RA_COL_BLOCKS")) ncb = (uint32_t)std::min(kMaxColBlocks, std::max(1, std::atoi(e)));
    if (std::getenv("GSFM_RA_NO_SLICE")) ncb = 1;
    s->cbk.ncb = ncb;
    s->cbk.cbsize = (N + ncb - 1) / ncb;
    s->slice = s->cbk.cbsize <= (uint32_t)kSliceMaxViews && !std::getenv("GSFM_RA_NO_SLICE");
  }
  const uint32_t ncb = s->cbk.ncb, cbsize = s->cbk.cbsize, NP = ncb * N;
  DevBuf<int> d_err;
  DevBuf<uint64_t> keys_a, keys_b;
  DevBuf<uint32_t> vals_a, vals_b, pieceptr, flags_ne, flags_iso, nz_rank, iso_rank;
  RA_TRY(d_err.alloc(4 + kMaxColBlocks + 1));
  CUDA_TRY(cudaMemsetAsync(d_err.p, 0, (4 + kMaxColBlocks + 1) * sizeof(int), st));
  RA_TRY(keys_a.alloc(H)); RA_TRY(keys_b.alloc(H)); RA_TRY(vals_a.alloc(H)); RA_TRY(vals_b.alloc(H));
  k_build_keys<<<grid_for(E), kBlock, 0, st>>>(E, N, cbsize, s->d_ei.p, s->d_ej.p, keys_a.p, vals_a.p, d_err.p);
  {
    int bits = 1;
    while (bits < 64 && ((uint64_t)NP * N - 1) >> bits) ++bits;
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_a.p, keys_b.p, vals_a.p, vals_b.p, (int)H, 0, bits, st));
    DevBuf<unsigned char> tmp;
    RA_TRY(tmp.alloc(bytes + 16));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys_a.p, keys_b.p, vals_a.p, vals_b.p, (int)H, 0, bits, st));
  }
  RA_TRY(s->he_col.alloc(H)); RA_TRY(s->he_row.alloc(H)); RA_TRY(s->he_edge.alloc(H));
  k_unpack_keys<<<grid_for(H), kBlock, 0, st>>>(H, N, keys_b.p, vals_b.p, s->he_row.p, s->he_col.p, s->he_edge.p, d_err.p);
  RA_TRY(pieceptr.alloc(NP + 2)); RA_TRY(flags_ne.alloc(NP + 2)); RA_TRY(nz_rank.alloc(NP + 2)); RA_TRY(flags_iso.alloc(N + 2)); RA_TRY(iso_rank.alloc(N + 2));
  k_pieceptr<<<grid_for(NP + 1), kBlock, 0, st>>>(NP, N, H, keys_b.p, pieceptr.p);
  k_piece_flags<<<grid_for(NP + 1), kBlock, 0, st>>>(NP, pieceptr.p, flags_ne.p);
  k_iso_flags<<<grid_for(N + 1), kBlock, 0, st>>>(N, ncb, pieceptr.p, flags_iso.p);
  RA_TRY(exclusive_scan(flags_ne.p, nz_rank.p, NP + 1, st));
  RA_TRY(exclusive_scan(flags_iso.p, iso_rank.p, N + 1, st));
  RA_TRY(s->iso.alloc(N));
  k_iso_fill<<<grid_for(N), kBlock, 0, st>>>(N, flags_iso.p, iso_rank.p, s->iso.p);
  CUDA_TRY(cudaMemcpyAsync(d_err.p + 1, iso_rank.p + N, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));  // n_iso
  for (uint32_t cb = 0; cb <= ncb; ++cb)  // first half-edge of every column block
    CUDA_TRY(cudaMemcpyAsync(d_err.p + 4 + cb, pieceptr.p + (size_t)cb * N, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  s->launches += 7;

  // ---- balanced partitions, one per kernel class, sized to exactly one resident wave of that kernel ----
  const int occ_k2 = di->occ_k2[s->blk == 4 ? 0 : s->blk == 6 ? 1 : 2];
  s->cooperative = di->coop != 0 && occ_k2 > 0;
  if (s->opt.linear_solver == GSFM_RA_SOLVER_AUTO)
    s->opt.linear_solver = (N <= GSFM_RA_AUTO_DENSE_MAX_VIEWS && s->cooperative && world == 1) ? GSFM_RA_SOLVER_DENSE_CHOLESKY : GSFM_RA_SOLVER_PCG;
  // Edge-sharded: the range length and the launch grid are sized from the LARGEST shard (ceil(E_total / world) edges), so
  // every rank launches the same grid even when the shards differ by an edge -- the replicated grid-wide sums then add in
  // the same order on every rank and the replicas stay bit-identical.  (A rank with fewer half-edges leaves its last
  // warps idle.)
  const uint64_t H_sizing = 2 * ((prob->num_edges + (uint64_t)world - 1) / (uint64_t)world);
  auto make = [&](Partition& P, int blocks_per_sm, int warps_per_block) -> int {
    const uint64_t max_warps = (uint64_t)s->sm_count * std::max(1, blocks_per_sm) * warps_per_block;
    uint64_t per = (H_sizing + max_warps - 1) / max_warps;
    per = std::max<uint64_t>(64, (per + 31) / 32 * 32);  // at least 64 half-edges per warp, whole records
    const uint32_t nw = (uint32_t)std::max<uint64_t>(1, (H + per - 1) / per);
    const uint32_t nw_sizing = (uint32_t)std::max<uint64_t>(1, (H_sizing + per - 1) / per);
    P.num_warps = nw; P.span = (uint32_t)per;
    P.num_segs = nw + NP;  // upper bound: every range start + every piece start opens one segment
    P.grid = (nw_sizing + warps_per_block - 1) / warps_per_block;
    DevBuf<uint32_t> nseg, cnt;
    RA_TRY(nseg.alloc(nw + 2)); RA_TRY(cnt.alloc(NP + 2));
    RA_TRY(P.warp_seg_ptr.alloc(nw + 2)); RA_TRY(P.seg_row.alloc(P.num_segs)); RA_TRY(P.seg_begin.alloc(P.num_segs));
    RA_TRY(P.seg_len.alloc(P.num_segs)); RA_TRY(P.node_seg_ptr.alloc(NP + 2));
    k_part_count<<<grid_for(nw + 1), kBlock, 0, st>>>(nw, (uint32_t)per, H, N, cbsize, s->he_row.p, s->he_col.p, nz_rank.p, nseg.p);
    RA_TRY(exclusive_scan(nseg.p, P.warp_seg_ptr.p, nw + 1, st));
    k_part_fill<<<grid_for(nw), kBlock, 0, st>>>(nw, (uint32_t)per, H, N, cbsize, s->he_row.p, s->he_col.p, pieceptr.p, P.warp_seg_ptr.p, P.seg_row.p, P.seg_begin.p,
                                                 P.seg_len.p);
    k_node_seg_count<<<grid_for(NP + 1), kBlock, 0, st>>>(NP, (uint32_t)per, pieceptr.p, cnt.p);
    RA_TRY(exclusive_scan(cnt.p, P.node_seg_ptr.p, NP + 1, st));
    s->launches += 5;
    return 0;
  };
  int occ_k1 = 1;
  if (!s->general()) {
    // the two K1 instantiations this solver will launch (with / without Jacobian): opt in to their shared memory, size the
    // partition for the occupancy of the heavier one
    for (int jac = 0; jac < 2; ++jac) CUDA_TRY(cudaFuncSetAttribute((const void*)s->pick_k1(jac != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, s->k1_smem_bytes()));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_k1, (const void*)s->pick_k1(true), kBlock, s->k1_smem_bytes()));
  }
  RA_TRY(make(s->pk1, occ_k1, kWarpsPerBlock));
  RA_TRY(make(s->pk2, std::min(occ_k2, kPcgBlocksPerSM), kPcgWarps));
  // the cooperative grid must be fully resident; node loops are grid-strided so any size works
  s->pk2.grid = std::min<uint32_t>(s->pk2.grid, (uint32_t)s->sm_count * std::max(1, std::min(occ_k2, kPcgBlocksPerSM)));
  {
    // L2 residency of the matrix stream: pin what 62 % of the L2 can hold when the stream is larger than that but still
    // comparable to the L2 (GSFM_RA_L2_KEEP=0..7 overrides, 0 disables)
    const double stream = (double)((H + 31) / 32) * s->rec_doubles() * 8.0, budget = 0.62 * (double)di->l2_bytes;
    uint32_t k = 0;
    if (stream > budget) k = (uint32_t)std::min(7.0, std::floor(8.0 * budget / stream));
    if (const char* e = std::getenv("GSFM_RA_L2_KEEP")) k = (uint32_t)std::min(7, std::max(0, std::atoi(e)));
    s->keep8 = k;
  }
  lap("enqueue structure build");

  // ---- per half-edge constants (K0) and the solver's working set ---------------------------------
  {
    const size_t nrec_in = (size_t)((H + 31) / 32), rd = (size_t)in_rec_doubles(s->ku);
    RA_TRY(s->inrec.alloc(nrec_in * rd));
    CUDA_TRY(cudaMemsetAsync(s->inrec.p + (nrec_in - 1) * rd, 0, rd * 8, st));  // lanes past H in the last record
  }
  if (overlap_upload) CUDA_TRY(cudaStreamWaitEvent(st, cs.done, 0));  // the payload is on the device from here on
  k_setup_halfedges<<<grid_for(H), kBlock, 0, st>>>(H, s->ku, s->he_edge.p, s->he_row.p, s->he_col.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p,
                                                    prob->error_type, s->inrec.p, s->d_ei.p, s->d_orient.p);
  s->launches += 1;
  const size_t nrec = (size_t)((H + 31) / 32);
  for (int b = 0; b < 2; ++b) {
    RA_TRY(s->omega[b].alloc(3ull * N));
    RA_TRY(s->node_q[b].alloc(4ull * N));
    RA_TRY(s->node_JL[b].alloc(9ull * N));
    RA_TRY(s->val[b].alloc(nrec * s->rec_doubles()));
    // only the last (partial) record has lanes K1 never writes
    CUDA_TRY(cudaMemsetAsync(s->val[b].p + (nrec - 1) * s->rec_doubles(), 0, (size_t)s->rec_doubles() * 8, st));
    RA_TRY(s->lin[b].alloc(9ull * N + 2));
    s->Hd_p[b] = s->lin[b].p;
    s->gt_p[b] = s->lin[b].p + 6ull * N;
    RA_TRY(s->ediag[b].alloc(3ull * N));
    CUDA_TRY(cudaMemsetAsync(s->omega[b].p, 0, 3ull * N * sizeof(double), st));
  }
  k_embed_cols<<<grid_for(H), kBlock, 0, st>>>(H, s->blk, s->he_col.p, s->val[0].p, s->val[1].p);
  s->launches += 1;
  RA_TRY(s->part.alloc((size_t)s->pk1.num_segs * kPartStride));
  RA_TRY(s->ypart.alloc((size_t)s->pk2.num_segs * 3));
  for (DevBuf<double>* d : {&s->scale, &s->x, &s->r, &s->q, &s->sv, &s->bvec, &s->y, &s->delta}) RA_TRY(d->alloc(3ull * N));
  RA_TRY(s->z.alloc(4ull * N));
  RA_TRY(s->p.alloc(4ull * N));
  RA_TRY(s->bar_slots.alloc(2 * 4 * (size_t)s->pk2.grid + 8));
  CUDA_TRY(cudaMemsetAsync(s->bar_slots.p, 0, (2 * 4 * (size_t)s->pk2.grid + 8) * sizeof(unsigned long long), st));
  RA_TRY(s->row_cnt.alloc(N));
  CUDA_TRY(cudaMemsetAsync(s->row_cnt.p, 0, (size_t)N * sizeof(unsigned), st));
  RA_TRY(s->Dblk.alloc(6ull * N));
  RA_TRY(s->Minv.alloc(6ull * N));
  RA_TRY(s->slots.alloc((size_t)std::max<unsigned>(std::max(grid_for(3ull * N), grid_for(E)), 64) * 4 + 64));
  RA_TRY(s->counter.alloc(4));
  RA_TRY(s->sc.alloc(1));
  CUDA_TRY(cudaMemsetAsync(s->counter.p, 0, 4 * sizeof(unsigned), st));
  CUDA_TRY(cudaMemsetAsync(s->sc.p, 0, sizeof(DevScalars), st));
  CUDA_TRY(cudaMallocHost(&s->h_sc, sizeof(DevScalars)));
  CUDA_TRY(cudaHostAlloc(&s->mailbox, sizeof(HostMailbox), cudaHostAllocMapped));
  std::memset(s->mailbox, 0, sizeof(HostMailbox));
  CUDA_TRY(cudaHostGetDevicePointer(&s->mailbox_dev, s->mailbox, 0));
  CUDA_TRY(cudaMallocHost(&s->h_it_params, sizeof(IterParams)));
  RA_TRY(s->it_params.alloc(1));
  k_jacobi_scale<<<grid_for(3ull * N), kBlock, 0, st>>>(3 * N, s->ediag[0].p, s->scale.p, 0);
  s->launches += 1;
  CUDA_TRY(cudaGetLastError());
  lap("enqueue K0 + allocations");
  // the only synchronisation of the build: input checks and the isolated-view count
  int h_err[4 + kMaxColBlocks + 1] = {0};
  CUDA_TRY(cudaMemcpyAsync(h_err, d_err.p, (4 + kMaxColBlocks + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  lap("wait for the device");
  if (h_err[0] == 1) { set_error("an edge is out of range or a self loop"); return GSFM_RA_ERR_INVALID; }
  if (h_err[0] == 2) { set_error("duplicate edge: the same view pair appears twice"); return GSFM_RA_ERR_INVALID; }
  s->n_iso = (uint32_t)h_err[1];
  for (uint32_t cb = 0; cb <= ncb; ++cb) s->cbk.begin[cb] = (uint32_t)h_err[4 + cb];
  if (options->verbose >= 2) std::fprintf(stderr, "[gsfm_ra] setup total %.2f ms\n", now_ms() - tb0);
  s->ms_setup = s->elapsed_since(s->ev[0]);
  s->radius = options->initial_trust_region_radius;
  *out = s.release();
  return 0;
}

void reset_trust_region(gsfm_ra_solver* s) {
  s->linearized = false;
  s->scale_ready = false;
  s->radius = s->opt.initial_trust_region_radius;
  s->decrease_factor = 2.0;
  s->iteration = 0;
  s->invalid_steps = 0;
  s->last_successful = false;
  s->termination = GSFM_RA_TERM_NONE;
}

void push_trace(gsfm_ra_summary* sum, const gsfm_ra_iteration& it) {
  if (sum && sum->trace && sum->trace_size < sum->trace_capacity) sum->trace[sum->trace_size++] = it;
}

// One trust-region batch at linearisation buffer b: damping + PCG initialisation, the whole PCG solve, the step and the
// candidate point, and the speculative linearisation of the candidate (node prep, K1, node finalize + host mailbox).
int enqueue_batch(gsfm_ra_solver* s, int b, double mu) {
  const int c = b ^ 1;
  const uint32_t N = s->N;
  // single GPU / fused exchange with the persistent PCG kernel: damping + PCG initialisation run as the kernel's prologue,
  // step + candidate + its per-view preparation as its epilogue -> the batch is PCG, K1, node finalize
  const bool fuse = s->opt.linear_solver == GSFM_RA_SOLVER_PCG && s->cooperative && (!s->sharded() || s->peers_connected) && !std::getenv("GSFM_RA_NO_FUSE");
  if (!fuse) CUDA_TRY(cudaMemsetAsync(&s->sc.p->bad, 0, sizeof(int), s->stream));
  s->fused_step = fuse;
  const int rc = s->pcg_enqueue(b, mu, nullptr, nullptr, s->opt.pcg_rtol, s->opt.pcg_max_iterations);
  s->fused_step = false;
  RA_TRY(rc);
  if (!fuse) {
    k_apply_step<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->apply_args(b, c), s->slots.p, s->counter.p, s->sc.p);
    s->launches += 1;
  }
  RA_TRY(s->evaluate(c, true, true, fuse));
  return 0;
}

// Run the batch: as a CUDA graph replay where possible (single GPU, persistent PCG kernel) -- one launch, the kernels run
// back to back -- else as direct launches.  GSFM_RA_NO_GRAPH=1 forces direct launches.
int run_batch(gsfm_ra_solver* s, int b) {
  if (s->graph_state == 0) {
    const bool eligible = s->opt.linear_solver == GSFM_RA_SOLVER_PCG && s->cooperative && (!s->sharded() || s->peers_connected) && !std::getenv("GSFM_RA_NO_GRAPH");
    s->graph_state = eligible ? 1 : -1;
  }
  if (s->graph_state == 1 && !s->batch_graph[b]) {
    // capture this buffer's batch once
    const int64_t l0 = s->launches;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      s->graph_params = true;
      int rc = (cudaMemcpyAsync(s->it_params.p, s->h_it_params, sizeof(IterParams), cudaMemcpyHostToDevice, s->stream) == cudaSuccess) ? 0 : 1;
      if (rc == 0) rc = enqueue_batch(s, b, 0.0);
      s->graph_params = false;
      const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
      ok = rc == 0 && e == cudaSuccess && graph != nullptr;
      if (ok) ok = cudaGraphInstantiate(&s->batch_graph[b], graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
    }
    s->batch_launches = s->launches - l0;
    s->launches = l0;
    if (!ok) {
      (void)cudaGetLastError();
      s->batch_graph[b] = nullptr;
      s->graph_state = -1;
      if (s->opt.verbose) std::fprintf(stderr, "[gsfm_ra] CUDA graph capture of the trust-region batch failed; using direct launches\n");
    }
  }
  if (s->graph_state == 1) {
    s->h_it_params->mu = s->radius;
    s->h_it_params->seq = ++s->mailbox_seq;
    CUDA_TRY(cudaGraphLaunch(s->batch_graph[b], s->stream));
    s->launches += s->batch_launches;
    return 0;
  }
  return enqueue_batch(s, b, s->radius);
}

// Ceres-1.14 trust-region loop, same order of checks as oracle/ra_oracle.cc.
// A candidate point is linearised speculatively (full K1 into the other buffer): an accepted step
// then needs no second evaluation, a rejected one simply keeps the current buffer.
int iterate(gsfm_ra_solver* s, int max_new_iterations, gsfm_ra_summary* sum) {
  const gsfm_ra_options& o = s->opt;
  const double t_start = now_ms();
  const int64_t launches0 = s->launches;
  const double asm0 = s->ms_assemble, lin0 = s->ms_linear, cost0 = s->ms_cost;
  const uint32_t N = s->N;
  int succ = 0, unsucc = 0, unconverged = 0;
  int64_t lin_total = 0;
  if (sum) sum->trace_size = 0;
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->linearized) {
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    CUDA_TRY(cudaMemsetAsync(&s->sc.p->bad, 0, sizeof(int), s->stream));
    RA_TRY(s->evaluate(s->cur, true));
    RA_TRY(s->fetch_scalars());
    s->ms_assemble += s->elapsed_since(s->ev[0]);
    if (s->h_sc->bad) { set_error("non-finite cost or Jacobian at the initial point"); s->termination = GSFM_RA_TERM_FAILURE; return GSFM_RA_ERR_NUMERIC; }
    s->x_cost = s->h_sc->cost; s->gmax = s->h_sc->gmax; s->x_norm = std::sqrt(s->h_sc->xnorm2);
    s->initial_cost = s->x_cost;
    s->linearized = true;
    if (!s->scale_ready) {
      k_jacobi_scale<<<grid_for(3ull * N), kBlock, 0, s->stream>>>(3 * N, s->ediag[s->cur].p, s->scale.p, o.jacobi_scaling);
      s->launches += 1;
      s->scale_ready = true;
    }
    gsfm_ra_iteration it0;
    std::memset(&it0, 0, sizeof(it0));
    it0.cost = s->x_cost; it0.gradient_max_norm = s->gmax; it0.trust_region_radius = s->radius;
    push_trace(sum, it0);
    if (s->gmax <= o.gradient_tolerance) s->termination = GSFM_RA_TERM_GRADIENT_TOLERANCE;
  }
  int done_here = 0;
  while (s->termination == GSFM_RA_TERM_NONE && done_here < max_new_iterations) {
    if (s->iteration >= o.max_num_iterations) { s->termination = GSFM_RA_TERM_MAX_ITERATIONS; break; }
    if (s->last_successful && s->gmax <= o.gradient_tolerance) { s->termination = GSFM_RA_TERM_GRADIENT_TOLERANCE; break; }
    if (s->radius <= o.min_trust_region_radius) { s->termination = GSFM_RA_TERM_MIN_RADIUS; break; }
    ++s->iteration;
    ++done_here;
    gsfm_ra_iteration it;
    std::memset(&it, 0, sizeof(it));
    it.iteration = s->iteration;
    const int b = s->cur, c = s->cur ^ 1;
    // ---- one trust-region iteration = one stream-ordered batch, ONE host synchronisation -------
    //   k_prepare_solve -> k_pcg_persistent (whole PCG + Ht x) -> k_apply_step (step, candidate)
    //   -> speculative linearisation of the candidate (k_node_prep, K1, k_node_finalize)
    RA_TRY(run_batch(s, b));
    RA_TRY(s->fetch_mailbox(s->mailbox_seq));
    s->account_batch();
    if (s->h_sc->bad == 2) { set_error("multi-GPU exchange timed out: a peer rank did not reach the same CG step"); s->termination = GSFM_RA_TERM_FAILURE; return GSFM_RA_ERR_CUDA; }
    const bool breakdown = s->h_sc->pcg_breakdown != 0;
    const int lin_it = s->h_sc->pcg_iter;
    const double lin_res = (s->h_sc->bb > 0.0) ? std::sqrt(s->h_sc->rr / s->h_sc->bb) : 0.0;
    it.linear_iterations = lin_it; it.linear_residual = lin_res;
    lin_total += lin_it;
    if (o.linear_solver == GSFM_RA_SOLVER_PCG && !breakdown && lin_it >= o.pcg_max_iterations && lin_res > o.pcg_rtol) ++unconverged;
    const double model_change = -s->h_sc->dg - 0.5 * s->h_sc->dHd;
    it.model_cost_change = model_change;
    // `bad` also covers a non-finite candidate evaluation; a non-finite STEP shows up in step2
    const bool step_finite = std::isfinite(s->h_sc->step2) && std::isfinite(model_change);
    bool valid = !breakdown && step_finite && model_change > 0.0;
    it.step_is_valid = valid;
    if (!valid) {
      it.cost = s->x_cost; it.gradient_max_norm = s->gmax;
      if (++s->invalid_steps >= 5) { s->termination = GSFM_RA_TERM_INVALID_STEPS; it.trust_region_radius = s->radius; push_trace(sum, it); break; }
      s->radius *= 0.5;
      it.trust_region_radius = s->radius;
      s->last_successful = false;
      ++unsucc;
      push_trace(sum, it);
      continue;
    }
    s->invalid_steps = 0;
    double cand_cost = s->h_sc->cost;
    const bool cand_bad = s->h_sc->bad != 0 || !std::isfinite(cand_cost);
    if (cand_bad) cand_cost = DBL_MAX;
    it.candidate_cost = cand_cost;
    it.step_norm = std::sqrt(s->h_sc->step2);
    it.cost_change = s->x_cost - cand_cost;
    it.relative_decrease = it.cost_change / model_change;
    it.gradient_max_norm = s->gmax;
    if (it.step_norm <= o.parameter_tolerance * (s->x_norm + o.parameter_tolerance)) {
      s->termination = GSFM_RA_TERM_PARAMETER_TOLERANCE; it.cost = s->x_cost; it.trust_region_radius = s->radius; push_trace(sum, it); break;
    }
    if (std::fabs(it.cost_change) <= o.function_tolerance * s->x_cost) {
      s->termination = GSFM_RA_TERM_FUNCTION_TOLERANCE; it.cost = s->x_cost; it.trust_region_radius = s->radius; push_trace(sum, it); break;
    }
    if (it.relative_decrease > o.min_relative_decrease && !cand_bad) {
      s->cur = c;
      s->x_cost = cand_cost;
      s->gmax = s->h_sc->gmax;
      s->x_norm = std::sqrt(s->h_sc->xnorm2);
      it.gradient_max_norm = s->gmax;
      it.step_is_successful = 1;
      const double q = it.relative_decrease;
      s->radius = s->radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * q - 1.0, 3));
      s->radius = std::min(o.max_trust_region_radius, s->radius);
      s->decrease_factor = 2.0;
      s->last_successful = true;
      ++succ;
      it.cost = s->x_cost;
    } else {
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->last_successful = false;
      ++unsucc;
      it.cost = cand_cost;
    }
    it.trust_region_radius = s->radius;
    push_trace(sum, it);
    if (o.verbose)
      std::fprintf(stderr, "[gsfm_ra] it %3d cost %.12e dcost %+.3e |g| %.3e |step| %.3e q %.3e radius %.3e pcg %d (%.1e) %s\n", s->iteration, s->x_cost,
                   it.cost_change, it.gradient_max_norm, it.step_norm, it.relative_decrease, s->radius, lin_it, lin_res, it.step_is_successful ? "ok" : "rejected");
  }
  if (sum) {
    sum->termination = s->termination;
    sum->num_iterations = s->iteration;
    sum->num_successful_steps = succ;
    sum->num_unsuccessful_steps = unsucc;
    sum->total_linear_iterations = lin_total;
    sum->num_linear_unconverged = unconverged;
    sum->n_gpus_used = s->world;
    sum->initial_cost = s->initial_cost;
    sum->final_cost = s->x_cost;
    sum->ms_setup = s->ms_setup;
    sum->ms_assemble = s->ms_assemble - asm0;
    sum->ms_linear = s->ms_linear - lin0;
    sum->ms_cost = s->ms_cost - cost0;
    sum->ms_total = now_ms() - t_start;
    sum->kernel_launches = s->launches - launches0;
  }
  return 0;
}

// The exchange block of a sharded solver: plain cudaMalloc (IPC handles cannot be taken from the async pool), zeroed (tag 0 =
// "nothing yet": sequence numbers start at 1).
int alloc_exchange_block(gsfm_ra_solver* s) {
  if (s->xchg) return 0;
  const size_t bytes = ll_cells_total(s->N, s->world) * sizeof(LLCell);
  CUDA_TRY(cudaMalloc(&s->xchg, bytes));
  CUDA_TRY(cudaMemset(s->xchg, 0, bytes));
  return 0;
}

// RAII temp solver for the one-shot API calls
struct TempSolver {
  gsfm_ra_solver* s = nullptr;
  ~TempSolver() { delete s; }
};

int make_temp(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, int device, TempSolver* t) {
  gsfm_ra_options o;
  gsfm_ra_default_options(&o);
  if (loss) o.loss = *loss;
  o.device = device;
  RA_TRY(build_solver(problem, &o, 0, 1, &t->s));
  if (omega) RA_TRY(gsfm_ra_solver_set_rotations(t->s, omega));
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int gsfm_ra_abi_version(void) { return GSFM_RA_ABI_VERSION; }
const char* gsfm_ra_last_error(void) { return g_last_error.c_str(); }
int gsfm_ra_residual_dim(int32_t error_type) {
  return error_type == GSFM_RA_QUATERNION_NORM ? 4 : (error_type == GSFM_RA_ROTATION_MAT_FNORM ? 9 : 3);
}
int gsfm_ra_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

void gsfm_ra_default_options(gsfm_ra_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->loss.kind = GSFM_RA_LOSS_TRIVIAL;
  o->loss.scale = 1.0;
  o->max_num_iterations = 200;
  o->jacobi_scaling = 1;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->linear_solver = GSFM_RA_SOLVER_AUTO;
  o->pcg_max_iterations = 500;
  o->pcg_rtol = 1e-10;
  o->num_threads = 0;
  o->device = -1;
  o->verbose = 0;
  o->n_gpus = 0;
}

int gsfm_ra_solver_create(const gsfm_ra_problem* problem, const gsfm_ra_options* options, gsfm_ra_solver** out) {
  return build_solver(problem, options, 0, 1, out);
}
int gsfm_ra_solver_create_sharded(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int32_t rank, int32_t world_size,
                                  gsfm_ra_solver** out) {
  return build_solver(problem, options, rank, world_size, out);
}
void gsfm_ra_solver_destroy(gsfm_ra_solver* solver) {
  if (!solver) return;
  cudaSetDevice(solver->device);
  delete solver;
}
int gsfm_ra_solver_set_rotations(gsfm_ra_solver* s, const double* omega) {
  if (!s || !omega) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaMemcpyAsync(s->omega[s->cur].p, omega, 3ull * s->N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  reset_trust_region(s);
  return 0;
}
int gsfm_ra_solver_get_rotations(gsfm_ra_solver* s, double* omega) {
  if (!s || !omega) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  CUDA_TRY(cudaMemcpyAsync(omega, s->omega[s->cur].p, 3ull * s->N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}
int gsfm_ra_solver_reset(gsfm_ra_solver* s) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  reset_trust_region(s);
  return 0;
}
int gsfm_ra_solver_iterate(gsfm_ra_solver* s, int32_t num_iterations, gsfm_ra_summary* summary) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->opt.linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY) {
    if (s->N > 2730) { set_error("dense Cholesky is limited to 2730 views (8190 unknowns); use PCG"); return GSFM_RA_ERR_UNSUPPORTED; }
    if (s->sharded() || !s->cooperative) { set_error("dense Cholesky needs a single cooperative-launch device"); return GSFM_RA_ERR_UNSUPPORTED; }
  } else if (s->opt.linear_solver != GSFM_RA_SOLVER_PCG) {
    set_error("unknown linear solver %d", s->opt.linear_solver);
    return GSFM_RA_ERR_INVALID;
  }
  return iterate(s, num_iterations, summary);
}
int gsfm_ra_comm_unique_id(uint8_t* id) {
  if (!id) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (!ncclx::api()) { set_error("libnccl.so.2 could not be loaded (set GSFM_RA_NCCL_LIB)"); return GSFM_RA_ERR_UNSUPPORTED; }
  ncclx::UniqueId u;
  NCCL_TRY(ncclx::api()->GetUniqueId(&u));
  std::memcpy(id, u.internal, ncclx::kUniqueIdBytes);
  return 0;
}
int gsfm_ra_solver_comm_init(gsfm_ra_solver* s, const uint8_t* id) {
  if (!s || !id) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->world == 1) return 0;
  if (!ncclx::api()) { set_error("libnccl.so.2 could not be loaded (set GSFM_RA_NCCL_LIB)"); return GSFM_RA_ERR_UNSUPPORTED; }
  CUDA_TRY(cudaSetDevice(s->device));
  ncclx::UniqueId u;
  std::memcpy(u.internal, id, ncclx::kUniqueIdBytes);
  NCCL_TRY(ncclx::api()->CommInitRank(&s->comm, s->world, u, s->rank));
  return 0;
}
int gsfm_ra_solver_ipc_export(gsfm_ra_solver* s, uint8_t* handle) {
  if (!s || !handle) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  RA_TRY(alloc_exchange_block(s));
  std::memset(handle, 0, GSFM_RA_IPC_HANDLE_BYTES);
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, s->xchg));
  static_assert(sizeof(h) <= GSFM_RA_IPC_HANDLE_BYTES, "handle size");
  std::memcpy(handle, &h, sizeof(h));
  return 0;
}
int gsfm_ra_solver_ipc_import(gsfm_ra_solver* s, const uint8_t* handles) {
  if (!s || !handles) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (s->world > kMaxPeers) { set_error("at most %d ranks", kMaxPeers); return GSFM_RA_ERR_UNSUPPORTED; }
  if (!s->xchg) { set_error("call gsfm_ra_solver_ipc_export first"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) { s->peer_base[r] = s->xchg; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * GSFM_RA_IPC_HANDLE_BYTES, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(&s->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  s->peers_connected = true;
  s->peers_ipc = true;
  return 0;
}
int gsfm_ra_solver_edge_range(const gsfm_ra_solver* s, uint64_t* e0, uint64_t* e1) {
  if (!s) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (e0) *e0 = s->edge_begin;
  if (e1) *e1 = s->edge_begin + s->E;
  return 0;
}

void* gsfm_ra_solver_cuda_stream(gsfm_ra_solver* s) { return s ? (void*)s->stream : nullptr; }

int gsfm_ra_solver_info(const gsfm_ra_solver* s, int32_t* out) {
  if (!s || !out) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  out[0] = s->opt.linear_solver;
  out[1] = s->blk * 8 + 4;
  out[2] = (int32_t)s->pk2.grid;
  out[3] = kPcgBlock;
  out[4] = (int32_t)s->keep8;
  out[5] = s->world;
  out[6] = s->graph_state == 1 ? 1 : 0;
  out[7] = 0;
  return 0;
}

int gsfm_ra_measure_stream(uint64_t bytes, int32_t repeats, int32_t device, double* gb_per_s) {
  if (!gb_per_s || repeats < 1 || bytes < (1u << 20)) { set_error("bad argument (bytes >= 1 MiB, repeats >= 1)"); return GSFM_RA_ERR_INVALID; }
  int dev = 0;
  RA_TRY(select_device(device, &dev));
  DeviceInfo* di = nullptr;
  RA_TRY(device_info(dev, &di));
  constexpr int kCB = Chunk<6>::kBytes;  // the chunk size of the symmetric 6-double records
  const int smem = kPcgWarps * kStages2 * kCB + kPcgWarps * kStages2 * 8;
  CUDA_TRY(cudaFuncSetAttribute(k_stream_probe<kCB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const uint32_t nchunks = (uint32_t)(bytes / kCB);
  DevBuf<unsigned char> buf;
  DevBuf<double> sink;
  RA_TRY(buf.alloc((size_t)nchunks * kCB));
  RA_TRY(sink.alloc(8));
  CUDA_TRY(cudaMemset(buf.p, 0, (size_t)nchunks * kCB));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  const unsigned grid = (unsigned)(di->sm_count * kPcgBlocksPerSM);
  for (int k = 0; k < 3; ++k) k_stream_probe<kCB><<<grid, kPcgBlock, smem>>>(buf.p, nchunks, sink.p);  // warm: the buffer is now L2 resident if it fits
  CUDA_TRY(cudaEventRecord(e0));
  for (int k = 0; k < repeats; ++k) k_stream_probe<kCB><<<grid, kPcgBlock, smem>>>(buf.p, nchunks, sink.p);
  CUDA_TRY(cudaEventRecord(e1));
  CUDA_TRY(cudaEventSynchronize(e1));
  CUDA_TRY(cudaGetLastError());
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *gb_per_s = (double)nchunks * kCB * repeats / (ms * 1e-3) / 1e9;
  return 0;
}

int gsfm_ra_solver_time_kernels(gsfm_ra_solver* s, int32_t repeats, double* out_ms) {
  if (!s || !out_ms || repeats < 1) { set_error("bad argument"); return GSFM_RA_ERR_INVALID; }
  CUDA_TRY(cudaSetDevice(s->device));
  if (!s->linearized) { set_error("time_kernels needs a linearised solver (call iterate first)"); return GSFM_RA_ERR_INVALID; }
  const int b = s->cur, c = s->cur ^ 1;  // scratch output goes to the candidate buffer
  auto timed = [&](auto&& launch, double* ms) -> int {
    launch();  // warm
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    for (int k = 0; k < repeats; ++k) launch();
    CUDA_TRY(cudaGetLastError());
    *ms = s->elapsed_since(s->ev[0]) / repeats;
    return 0;
  };
  RA_TRY(timed([&] { s->launch_edges(b, true, s->val[c].p); }, &out_ms[0]));
  RA_TRY(timed([&] { s->launch_edges(b, false, nullptr); }, &out_ms[1]));
  CUDA_TRY(cudaMemsetAsync(s->z.p, 0, 4ull * s->N * sizeof(double), s->stream));
  RA_TRY(timed([&] { s->launch_spmv(b, s->z.p, std::getenv("GSFM_RA_DEBUG_NOGATHER") ? 2 : 0); }, &out_ms[2]));
  // one PCG iteration inside the persistent kernel: (time of R iterations) / R with rtol = 0
  {
    const int R = std::max(8, (int)repeats);
    RA_TRY(s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R));  // warm
    CUDA_TRY(cudaEventRecord(s->ev[0], s->stream));
    RA_TRY(s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R));
    RA_TRY(s->fetch_scalars());
    const int done_it = std::max(1, s->h_sc->pcg_iter);
    out_ms[3] = s->elapsed_since(s->ev[0]) / done_it;
    if (std::getenv("GSFM_RA_PROFILE_PHASES")) {
      DevBuf<unsigned long long> pb;
      RA_TRY(pb.alloc(8));
      CUDA_TRY(cudaMemsetAsync(pb.p, 0, 64, s->stream));
      s->prof_buf = pb.p;
      const int rc = s->pcg_enqueue(b, s->radius, nullptr, nullptr, 0.0, R);
      s->prof_buf = nullptr;
      RA_TRY(rc);
      unsigned long long hp[8];
      CUDA_TRY(cudaMemcpyAsync(hp, pb.p, 64, cudaMemcpyDeviceToHost, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));
      const double n = std::max<double>(1.0, (double)hp[6]);
      std::fprintf(stderr, "[gsfm_ra] PCG phases of block 0, us per iteration: spmv_pass %.2f | barrier+sum %.2f | vector update %.2f | barrier+sum %.2f  "
                   "(%d iterations)\n", hp[0] / n / 1e3, hp[1] / n / 1e3, hp[3] / n / 1e3, hp[4] / n / 1e3, (int)hp[6]);
    }
  }
  // restore the linearisation-dependent partials (K1 scratch wrote `part`): re-run the finalize inputs
  RA_TRY(s->evaluate(b, true));
  RA_TRY(s->fetch_scalars());
  return 0;
}

// How many devices a one-shot solve shards over (options.n_gpus, include/gsfm_ra.h).  Multi-GPU needs the PCG path (the dense
// factorisation is a single-device kernel) and peer access between all the devices; -1 asks for as many devices as keep
// >= GSFM_RA_MIN_EDGES_PER_GPU edges each (below that the per-step exchange costs more than the pass it saves).
static int resolve_world(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int* dev0_out) {
  int want = options->n_gpus;
  if (want == 0 || want == 1 || !problem) return 1;
  if (options->linear_solver == GSFM_RA_SOLVER_DENSE_CHOLESKY) return 1;
  if (options->linear_solver == GSFM_RA_SOLVER_AUTO && problem->num_views <= GSFM_RA_AUTO_DENSE_MAX_VIEWS) return 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) { cudaGetLastError(); return 1; }
  int dev0 = options->device;
  if (dev0 < 0 && cudaGetDevice(&dev0) != cudaSuccess) return 1;
  const int avail = std::min(ndev - dev0, kMaxPeers);
  uint64_t min_edges = GSFM_RA_MIN_EDGES_PER_GPU;
  if (const char* e = std::getenv("GSFM_RA_MIN_EDGES_PER_GPU")) min_edges = std::max<long long>(1, std::atoll(e));  // tests
  if (want < 0) want = (int)std::min<uint64_t>((uint64_t)avail, problem->num_edges / min_edges);
  want = std::max(1, std::min(want, avail));
  for (int a = 0; a < want; ++a)
    for (int b = 0; b < want; ++b) {
      int ok = 1;
      if (a != b && (cudaDeviceCanAccessPeer(&ok, dev0 + a, dev0 + b) != cudaSuccess || !ok)) { cudaGetLastError(); return 1; }
    }
  *dev0_out = dev0;
  return want;
}

// One process, W devices: one host thread per device, each with its own resident solver on its shard of the edges; the
// exchange blocks are wired by peer access (no IPC, no NCCL, no host-side hand-shake).  Every thread runs the same
// trust-region loop on bit-identical replicated scalars, so they take the same decisions without talking to each other.
static int solve_multi(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int dev0, int W, double* omega_inout,
                       gsfm_ra_summary* summary) {
  const double t_enter = now_ms();
  std::vector<gsfm_ra_solver*> sv(W, nullptr);
  std::vector<int> rc(W, 0);
  std::vector<std::string> err(W);
  auto parallel = [&](auto&& body) {
    std::vector<std::thread> th;
    for (int r = 0; r < W; ++r) th.emplace_back([&, r] { rc[r] = body(r); if (rc[r] != 0) err[r] = g_last_error; });
    for (auto& t : th) t.join();
    for (int r = 0; r < W; ++r)
      if (rc[r] != 0 && rc[r] != GSFM_RA_ERR_NUMERIC) { set_error("device %d: %s", dev0 + r, err[r].c_str()); return rc[r]; }
    return rc[0];
  };
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  struct Cleanup {  // destroys the solvers on their own devices, then puts the caller's current device back
    std::vector<gsfm_ra_solver*>& v;
    int restore;
    ~Cleanup() { for (auto* s : v) if (s) { cudaSetDevice(s->device); delete s; } cudaSetDevice(restore); }
  } cleanup{sv, prev_dev};
  int st = parallel([&](int r) -> int {
    gsfm_ra_options o = *options;
    o.device = dev0 + r; o.n_gpus = 0;
    if (r != 0) o.verbose = 0;
    RA_TRY(build_solver(problem, &o, r, W, &sv[r]));
    RA_TRY(alloc_exchange_block(sv[r]));
    for (int q = 0; q < W; ++q) {
      if (q == r) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(dev0 + q, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", dev0 + r, dev0 + q, cudaGetErrorString(e)); return GSFM_RA_ERR_CUDA; }
      cudaGetLastError();
    }
    return 0;
  });
  if (st != 0) { cudaSetDevice(prev_dev); return st; }
  for (int r = 0; r < W; ++r) {
    for (int q = 0; q < W; ++q) sv[r]->peer_base[q] = sv[q]->xchg;
    sv[r]->peers_connected = true;
  }
  const double t_built = now_ms();
  st = parallel([&](int r) -> int {
    RA_TRY(gsfm_ra_solver_set_rotations(sv[r], omega_inout));
    gsfm_ra_summary local;
    std::memset(&local, 0, sizeof(local));
    return gsfm_ra_solver_iterate(sv[r], options->max_num_iterations + 1, r == 0 ? summary : &local);
  });
  int out = st;
  const double t_solved = now_ms();
  if (st == 0 || st == GSFM_RA_ERR_NUMERIC) {
    const int g = gsfm_ra_solver_get_rotations(sv[0], omega_inout);
    if (g != 0) out = g;
  }
  if (options->verbose)
    std::fprintf(stderr, "[gsfm_ra] %d devices: build + connect %.2f ms, solve %.2f ms (%d iterations)\n", W, t_built - t_enter, t_solved - t_built,
                 summary ? summary->num_iterations : -1);
  cudaSetDevice(prev_dev);
  return out;
}

int gsfm_ra_solve(const gsfm_ra_problem* problem, const gsfm_ra_options* options, double* omega_inout, gsfm_ra_summary* summary) {
  if (!omega_inout) { set_error("omega_inout is NULL"); return GSFM_RA_ERR_INVALID; }
  if (!options) { set_error("options is NULL"); return GSFM_RA_ERR_INVALID; }
  const double t0 = now_ms();
  int dev0 = 0;
  const int W = resolve_world(problem, options, &dev0);
  if (W > 1) {
    RA_TRY(check_problem(problem));
    if (options->verbose) std::fprintf(stderr, "[gsfm_ra] sharding %llu edges over devices %d..%d\n", (unsigned long long)problem->num_edges, dev0, dev0 + W - 1);
    const int rc = solve_multi(problem, options, dev0, W, omega_inout, summary);
    if (summary) summary->ms_total = now_ms() - t0;
    return rc;
  }
  TempSolver t;
  RA_TRY(build_solver(problem, options, 0, 1, &t.s));
  RA_TRY(gsfm_ra_solver_set_rotations(t.s, omega_inout));
  const int rc = gsfm_ra_solver_iterate(t.s, options->max_num_iterations + 1, summary);
  if (rc != 0 && rc != GSFM_RA_ERR_NUMERIC) return rc;
  RA_TRY(gsfm_ra_solver_get_rotations(t.s, omega_inout));
  if (summary) summary->ms_total = now_ms() - t0;
  return rc;
}

int gsfm_ra_solve_sigma_consensus(const gsfm_ra_problem* problem, const gsfm_ra_options* options, int32_t iters_num, double sigma_max,
                                  double* omega_inout, gsfm_ra_summary* summary) {
  if (!omega_inout || !options || !problem) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (problem->error_type != GSFM_RA_ANGLE_AXIS) { set_error("sigma consensus runs on GSFM_RA_ANGLE_AXIS (PairwiseRotationError with a scalar weight)"); return GSFM_RA_ERR_INVALID; }
  if (!(sigma_max > 0.0) || iters_num < 1) { set_error("sigma_max must be > 0 and iters_num >= 1"); return GSFM_RA_ERR_INVALID; }
  const double t0 = now_ms();
  gsfm_ra_problem q = *problem;
  q.edge_weight = nullptr;
  TempSolver t;
  RA_TRY(build_solver(&q, options, 0, 1, &t.s));
  gsfm_ra_solver* s = t.s;
  RA_TRY(gsfm_ra_solver_set_rotations(s, omega_inout));
  const uint64_t E = s->E, H = s->H;
  {
    AllocScope scope(s->stream);
    RA_TRY(s->d_weight.alloc(E));
  }
  CUDA_TRY(cudaMemsetAsync(s->d_weight.p, 0, E * sizeof(double), s->stream));  // last_weights start at 0
  // include/gamma_values.cpp:6-11 (nu = 3)
  const double C3 = 4.029720004054876e-01, gamma_k = 3.439485560754856e-03, table_size = 36843.0;
  const double sq2 = sigma_max * sigma_max * 2.0, one_over_sigma = C3 * 2.0 / sigma_max, weight_zero = one_over_sigma * (1.0 - gamma_k);
  gsfm_ra_summary total;
  std::memset(&total, 0, sizeof(total));
  if (summary) { total.trace = summary->trace; total.trace_capacity = summary->trace_capacity; }
  int rc = 0;
  for (int it = 0; it < iters_num; ++it) {
    const int b = s->cur;
    k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[b].p, s->node_q[b].p, s->node_JL[b].p, s->slots.p, s->counter.p, s->sc.p, 0);
    k_sigma_weights<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->node_q[b].p, one_over_sigma, sq2, gamma_k,
                                                           weight_zero, table_size, s->d_weight.p, s->slots.p, s->counter.p, s->sc.p);
    k_setup_halfedges<<<grid_for(H), kBlock, 0, s->stream>>>(H, s->ku, s->he_edge.p, s->he_row.p, s->he_col.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p,
                                                             s->error_type, s->inrec.p);
    s->launches += 3;
    CUDA_TRY(cudaGetLastError());
    RA_TRY(s->fetch_scalars());
    // the reference divides by view_pairs.size(), the pairs it skipped included (rotation_estimator.cpp:419-424)
    const double diff = s->h_sc->dg / (double)(problem->total_pair_count > 0 ? (uint64_t)problem->total_pair_count : E);
    reset_trust_region(s);
    gsfm_ra_summary s1;
    std::memset(&s1, 0, sizeof(s1));
    if (total.trace && total.trace_size < total.trace_capacity) { s1.trace = total.trace + total.trace_size; s1.trace_capacity = total.trace_capacity - total.trace_size; }
    rc = gsfm_ra_solver_iterate(s, options->max_num_iterations + 1, &s1);
    if (rc != 0 && rc != GSFM_RA_ERR_NUMERIC) return rc;
    if (it == 0) total.initial_cost = s1.initial_cost;
    total.final_cost = s1.final_cost; total.termination = s1.termination;
    total.num_iterations += s1.num_iterations; total.num_successful_steps += s1.num_successful_steps;
    total.num_unsuccessful_steps += s1.num_unsuccessful_steps; total.total_linear_iterations += s1.total_linear_iterations;
    total.ms_assemble += s1.ms_assemble; total.ms_linear += s1.ms_linear; total.ms_cost += s1.ms_cost; total.kernel_launches += s1.kernel_launches + 3;
    total.trace_size += s1.trace_size; total.num_linear_unconverged += s1.num_linear_unconverged; total.n_gpus_used = 1;
    total.outer_iterations = it + 1;
    total.last_weight_change = diff;
    if (rc != 0 || diff <= 1e-7) break;
  }
  RA_TRY(gsfm_ra_solver_get_rotations(s, omega_inout));
  total.ms_setup = s->ms_setup;
  total.ms_total = now_ms() - t0;
  if (summary) *summary = total;
  return rc;
}

int gsfm_ra_eval_edges(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* r, double* jac_i, double* jac_j,
                       double* rho, int32_t device) {
  if (!omega) { set_error("omega is NULL"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint64_t E = s->E;
  DevBuf<double> dr, dji, djj, drho;
  const uint64_t d = (uint64_t)gsfm_ra_residual_dim(s->error_type);
  if (r) RA_TRY(dr.alloc(d * E));
  if (jac_i) RA_TRY(dji.alloc(3 * d * E));
  if (jac_j) RA_TRY(djj.alloc(3 * d * E));
  if (rho) RA_TRY(drho.alloc(3 * E));
  k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[0].p, s->node_q[0].p, s->node_JL[0].p, s->slots.p, s->counter.p, s->sc.p, s->param_kind());
  if (s->error_type == GSFM_RA_QUATERNION_NORM)
    k_eval_edges_general<0><<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_weight.p, s->node_q[0].p, s->node_JL[0].p,
                                                                   s->loss, dr.p, dji.p, djj.p, drho.p);
  else if (s->error_type == GSFM_RA_ROTATION_MAT_FNORM)
    k_eval_edges_general<1><<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_weight.p, s->node_q[0].p, s->node_JL[0].p,
                                                                   s->loss, dr.p, dji.p, djj.p, drho.p);
  else
    k_eval_edges<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->d_cov6.p, s->d_weight.p, s->error_type,
                                                        s->node_q[0].p, s->node_JL[0].p, s->loss, dr.p, dji.p, djj.p, drho.p, s->d_orient.p);
  CUDA_TRY(cudaGetLastError());
  if (r) CUDA_TRY(cudaMemcpyAsync(r, dr.p, d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (jac_i) CUDA_TRY(cudaMemcpyAsync(jac_i, dji.p, 3 * d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (jac_j) CUDA_TRY(cudaMemcpyAsync(jac_j, djj.p, 3 * d * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  if (rho) CUDA_TRY(cudaMemcpyAsync(rho, drho.p, 3 * E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

int gsfm_ra_whiten(const gsfm_ra_problem* problem, double* U, int32_t device) {
  RA_TRY(check_problem(problem));
  if (!U) { set_error("U is NULL"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  const uint64_t E = problem->num_edges;
  DevBuf<double> dc, dw, du;
  if (problem->cov6) { RA_TRY(dc.alloc(6 * E)); CUDA_TRY(cudaMemcpy(dc.p, problem->cov6, 6 * E * sizeof(double), cudaMemcpyHostToDevice)); }
  if (problem->edge_weight) { RA_TRY(dw.alloc(E)); CUDA_TRY(cudaMemcpy(dw.p, problem->edge_weight, E * sizeof(double), cudaMemcpyHostToDevice)); }
  RA_TRY(du.alloc(9 * E));
  k_whiten_edges<<<grid_for(E), kBlock>>>(E, dc.p, dw.p, problem->error_type, du.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(U, du.p, 9 * E * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int gsfm_ra_assemble(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* cost, double* gradient, double* hdiag,
                     uint32_t* rowptr, uint32_t* col, double* val, int32_t device) {
  if (!omega) { set_error("omega is NULL"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  RA_TRY(s->evaluate(0, true));
  RA_TRY(s->fetch_scalars());
  if (cost) *cost = s->h_sc->cost;
  const uint32_t N = s->N;
  const uint64_t H = s->H;
  DevBuf<double> dval, dh, dg;
  DevBuf<uint32_t> dcol;
  if (val || col) {
    RA_TRY(dval.alloc(9 * H));
    RA_TRY(dcol.alloc(H));
    // block-CSR order of the API = (row, col); a column-blocked layout is exported through the sorted permutation
    DevBuf<uint64_t> ka, kb;
    DevBuf<uint32_t> ia, ib;
    const uint32_t* order = nullptr;
    if (s->cbk.ncb > 1) {
      RA_TRY(ka.alloc(H)); RA_TRY(kb.alloc(H)); RA_TRY(ia.alloc(H)); RA_TRY(ib.alloc(H));
      k_rowmajor_keys<<<grid_for(H), kBlock, 0, s->stream>>>(H, N, s->he_row.p, s->he_col.p, ka.p, ia.p);
      size_t bytes = 0;
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ka.p, kb.p, ia.p, ib.p, (int)H, 0, 64, s->stream));
      DevBuf<unsigned char> tmp;
      RA_TRY(tmp.alloc(bytes + 16));
      CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ka.p, kb.p, ia.p, ib.p, (int)H, 0, 64, s->stream));
      CUDA_TRY(cudaStreamSynchronize(s->stream));  // tmp goes out of scope
      order = ib.p;
    }
    k_export_blocks<<<grid_for(H), kBlock, 0, s->stream>>>(H, s->blk, s->he_row.p, s->he_col.p, s->val[0].p, s->node_JL[0].p, order, dval.p, dcol.p);
    CUDA_TRY(cudaGetLastError());
    if (val) CUDA_TRY(cudaMemcpyAsync(val, dval.p, 9 * H * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (col) CUDA_TRY(cudaMemcpyAsync(col, dcol.p, H * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
  }
  if (hdiag || gradient) {
    RA_TRY(dh.alloc(9ull * N));
    RA_TRY(dg.alloc(3ull * N));
    k_export_nodes<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->Hd_p[0], s->gt_p[0], s->node_JL[0].p, dh.p, dg.p);
    CUDA_TRY(cudaGetLastError());
    if (hdiag) CUDA_TRY(cudaMemcpyAsync(hdiag, dh.p, 9ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (gradient) CUDA_TRY(cudaMemcpyAsync(gradient, dg.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (rowptr) {
    std::vector<uint32_t> he_row(H);
    CUDA_TRY(cudaMemcpy(he_row.data(), s->he_row.p, H * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    std::fill(rowptr, rowptr + N + 1, 0u);
    for (uint64_t h = 0; h < H; ++h) rowptr[he_row[h] + 1]++;
    for (uint32_t a = 0; a < N; ++a) rowptr[a + 1] += rowptr[a];
  }
  return 0;
}

int gsfm_ra_cost(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, double* cost, int32_t device) {
  if (!omega || !cost) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  RA_TRY(t.s->evaluate(0, false));
  RA_TRY(t.s->fetch_scalars());
  *cost = t.s->h_sc->cost;
  return 0;
}

int gsfm_ra_spmv(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, const double* damping, const double* x, double* y,
                 int32_t device) {
  if (!omega || !x || !y) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint32_t N = s->N;
  RA_TRY(s->evaluate(0, true));
  DevBuf<double> dx, dd;
  RA_TRY(dx.alloc(3ull * N));
  CUDA_TRY(cudaMemcpyAsync(dx.p, x, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (damping) { RA_TRY(dd.alloc(3ull * N)); CUDA_TRY(cudaMemcpyAsync(dd.p, damping, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream)); }
  // y = Jl^T Ht (Jl x) + damping .* x
  k_node_apply<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->node_JL[0].p, dx.p, 0, nullptr, nullptr, s->p.p, 4);
  RA_TRY(s->spmv(0, s->p.p, s->y.p, s->Hd_p[0]));
  k_node_apply<<<grid_for(N), kBlock, 0, s->stream>>>(N, s->node_JL[0].p, s->y.p, 1, dd.p, dx.p, s->delta.p, 3);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(y, s->delta.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

int gsfm_ra_pcg(const gsfm_ra_problem* problem, const gsfm_ra_loss* loss, const double* omega, const double* damping, const double* b, double rtol,
                int32_t max_iterations, double* x, int32_t* iterations, double* rel_residual, int32_t device) {
  if (!omega || !b || !x) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  TempSolver t;
  RA_TRY(make_temp(problem, loss, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint32_t N = s->N;
  RA_TRY(s->evaluate(0, true));
  DevBuf<double> db, dd;
  RA_TRY(db.alloc(3ull * N));
  RA_TRY(dd.alloc(3ull * N));
  CUDA_TRY(cudaMemcpyAsync(db.p, b, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  if (damping) CUDA_TRY(cudaMemcpyAsync(dd.p, damping, 3ull * N * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  else CUDA_TRY(cudaMemsetAsync(dd.p, 0, 3ull * N * sizeof(double), s->stream));
  RA_TRY(s->pcg_enqueue(0, 1.0, dd.p, db.p, rtol, max_iterations));
  RA_TRY(s->fetch_scalars());
  const int it = s->h_sc->pcg_iter;
  const double res = (s->h_sc->bb > 0.0) ? std::sqrt(s->h_sc->rr / s->h_sc->bb) : 0.0;
  // x = Jl^-1 xt
  {
    ApplyArgs A;
    std::memset(&A, 0, sizeof(A));
    A.node_JL = s->node_JL[0].p; A.xt = s->x.p; A.delta_out = s->delta.p;
    k_apply_step<<<grid_for(N), kBlock, 0, s->stream>>>(N, A, s->slots.p, s->counter.p, s->sc.p);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(x, s->delta.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if (iterations) *iterations = it;
  if (rel_residual) *rel_residual = res;
  return 0;
}

int gsfm_ra_eval_loss(const gsfm_ra_loss* loss, const double* s_in, uint64_t n, double* out, int32_t device) {
  if (!loss || !s_in || !out) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  DevLoss L;
  RA_TRY(make_dev_loss(loss, &L));
  if (n == 0) return 0;
  DevBuf<double> ds, dout, dtab;
  RA_TRY(attach_loss_table(loss, &L, &dtab, nullptr));
  RA_TRY(ds.alloc(n));
  RA_TRY(dout.alloc(3 * n));
  CUDA_TRY(cudaMemcpy(ds.p, s_in, n * sizeof(double), cudaMemcpyHostToDevice));
  k_eval_loss<<<grid_for(n), kBlock>>>(n, ds.p, L, dout.p);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(out, dout.p, 3 * n * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int gsfm_ra_filter_view_pairs(const gsfm_ra_problem* problem, const double* omega, double max_degrees, uint8_t* keep, double* angle_rad,
                              int32_t device) {
  if (!omega || !problem) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (!(max_degrees >= 0.0)) { set_error("max_relative_rotation_difference_degrees must be >= 0"); return GSFM_RA_ERR_INVALID; }
  gsfm_ra_problem p2 = *problem;
  p2.error_type = GSFM_RA_ANGLE_AXIS;
  TempSolver t;
  RA_TRY(make_temp(&p2, nullptr, omega, device, &t));
  gsfm_ra_solver* s = t.s;
  const uint64_t E = s->E;
  DevBuf<uint8_t> dk;
  DevBuf<double> da;
  RA_TRY(dk.alloc(E));
  RA_TRY(da.alloc(E));
  const double thr = max_degrees * M_PI / 180.0;
  k_node_prep<<<grid_for(s->N), kBlock, 0, s->stream>>>(s->N, s->omega[0].p, s->node_q[0].p, s->node_JL[0].p, s->slots.p, s->counter.p, s->sc.p, s->param_kind());
  k_filter_pairs<<<grid_for(E), kBlock, 0, s->stream>>>(E, s->d_ei.p, s->d_ej.p, s->d_omega_ij.p, s->node_q[0].p, thr * thr, dk.p, da.p);
  CUDA_TRY(cudaGetLastError());
  if (keep) CUDA_TRY(cudaMemcpyAsync(keep, dk.p, E, cudaMemcpyDeviceToHost, s->stream));
  if (angle_rad) CUDA_TRY(cudaMemcpyAsync(angle_rad, da.p, E * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return 0;
}

// ---- translation averaging (include/gsfm_pa.h): the same solver on error type GSFM_RA_POSITION_BASELINE ------------------
void gsfm_pa_default_options(gsfm_ra_options* o) {
  if (!o) return;
  gsfm_ra_default_options(o);
  o->max_num_iterations = 400;           // NonlinearPositionEstimator::Options::max_num_iterations
  o->loss.kind = GSFM_RA_LOSS_HUBER;     // new ceres::HuberLoss(options_.robust_loss_width), position_estimator.cpp:330
  o->loss.p[0] = 0.1;
}

int gsfm_pa_as_ra_problem(const gsfm_pa_problem* p, gsfm_ra_problem* out) {
  if (!p || !out) { set_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  if (p->error_type != GSFM_PA_BASELINE && p->error_type != GSFM_PA_COVARIANCE) { set_error("unknown position error type %d", p->error_type); return GSFM_RA_ERR_INVALID; }
  if (!p->position_2 || !p->orientation) { set_error("position_2 / orientation is NULL"); return GSFM_RA_ERR_INVALID; }
  std::memset(out, 0, sizeof(*out));
  out->num_views = p->num_views; out->num_edges = p->num_edges;
  out->edge_i = p->edge_i; out->edge_j = p->edge_j;
  out->omega_ij = p->position_2; out->edge_weight = p->edge_weight;
  out->error_type = GSFM_RA_POSITION_BASELINE;
  out->orientation = p->orientation;
  out->fixed_view = p->fixed_view;
  return 0;
}

int gsfm_pa_solve(const gsfm_pa_problem* problem, const gsfm_ra_options* options, double* positions_inout, gsfm_ra_summary* summary) {
  gsfm_ra_problem q;
  RA_TRY(gsfm_pa_as_ra_problem(problem, &q));
  return gsfm_ra_solve(&q, options, positions_inout, summary);
}

int gsfm_pa_eval_edges(const gsfm_pa_problem* problem, const gsfm_ra_loss* loss, const double* positions, double* r, double* jac_i, double* jac_j,
                       double* rho, int32_t device) {
  gsfm_ra_problem q;
  RA_TRY(gsfm_pa_as_ra_problem(problem, &q));
  return gsfm_ra_eval_edges(&q, loss, positions, r, jac_i, jac_j, rho, device);
}

int gsfm_pa_cost(const gsfm_pa_problem* problem, const gsfm_ra_loss* loss, const double* positions, double* cost, int32_t device) {
  gsfm_ra_problem q;
  RA_TRY(gsfm_pa_as_ra_problem(problem, &q));
  return gsfm_ra_cost(&q, loss, positions, cost, device);
}

}  // extern "C"

#include "gsfm_graph.cuh"
