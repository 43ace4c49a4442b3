// gsfm_graph.cuh -- view-graph shaping on the device: the steps immediately BEFORE the rotation-averaging solve
// (SURVEY.md section 8 f, rank 1).  Included by gsfm_ra.cu (shares its helpers); fp64 / u32 as everywhere.
//
//   gsfm_ra_filter_initial_view_graph   FilterInitialViewGraph, reference src/GSfM_global_reconstruction_estimator.cpp:369-390
//                                       (drop view pairs with fewer verified matches than the threshold) followed by
//                                       RemoveDisconnectedViewPairs, T/sfm/view_graph/remove_disconnected_view_pairs.cc:48-
//                                       (keep the largest connected component)
//   gsfm_ra_init_orientations_mst       OrientationsFromMaximumSpanningTree,
//                                       T/sfm/view_graph/orientations_from_maximum_spanning_tree.cc:109-178: Kruskal on
//                                       num_verified_matches (T/math/graph/minimum_spanning_tree.h:70-98), then
//                                       R_neighbor = (src < nbr ? R_rel : R_rel^T) R_src from the root (:60-83)
//
// Both are hash-map / union-find walks on the host in the reference.  Here:
//   * connected components: min-label hooking over the edge list + pointer jumping until nothing changes -- the final
//     label of a component is its smallest view index, whatever the order the hardware applies the atomicMin's in;
//   * maximum spanning tree: the edges get a strict total order (weight descending, then i, then j -- the order the host
//     restatement's Kruskal sorts by), under which the tree is UNIQUE; Boruvka rounds (every component picks its best
//     outgoing edge with one atomicMin on the edge's rank, mutual picks keep the smaller root) build exactly that tree in
//     <= log2(N) edge-parallel rounds;
//   * orientations: level-synchronous propagation over the N-1 tree edges by one CTA (a view is reached through exactly
//     one tree edge, so there is no race), rotations composed as unit quaternions.
// Integer atomics only: every result is deterministic.
#pragma once

namespace graphk {

constexpr uint32_t kNone = 0xffffffffu;

__global__ void k_iota(uint32_t n, uint32_t* __restrict__ a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
__global__ void k_fill(uint32_t n, uint32_t* __restrict__ a, uint32_t v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__device__ __forceinline__ uint32_t find_root(const uint32_t* parent, uint32_t v) {
  uint32_t p = parent[v];
  while (p != v) { v = p; p = parent[v]; }
  return v;
}

// ---- connected components -------------------------------------------------------------------
// hook the larger root under the smaller one (parent links only ever decrease: no cycles)
__global__ void k_cc_hook(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const uint8_t* __restrict__ keep,
                          uint32_t* parent, uint32_t* changed) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || (keep && !keep[e])) return;
  const uint32_t ru = find_root(parent, ei[e]), rv = find_root(parent, ej[e]);
  if (ru == rv) return;
  const uint32_t hi = ru > rv ? ru : rv, lo = ru > rv ? rv : ru;
  atomicMin(&parent[hi], lo);
  *changed = 1u;
}
__global__ void k_cc_compress(uint32_t N, uint32_t* parent) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) parent[v] = find_root(parent, v);
}
__global__ void k_filter_edges(uint64_t E, const int32_t* __restrict__ matches, int32_t min_matches, uint8_t* __restrict__ keep) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) keep[e] = matches[e] >= min_matches ? 1 : 0;
}
__global__ void k_touch(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const uint8_t* __restrict__ keep,
                        uint32_t* __restrict__ touched) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || !keep[e]) return;
  touched[ei[e]] = 1u;
  touched[ej[e]] = 1u;
}
__global__ void k_cc_count(uint32_t N, const uint32_t* __restrict__ label, const uint32_t* __restrict__ touched, uint32_t* cnt) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N && touched[v]) atomicAdd(&cnt[label[v]], 1u);
}
// largest count, ties -> smallest label: one packed atomicMax on (count << 32 | ~label)
__global__ void k_cc_best(uint32_t N, const uint32_t* __restrict__ cnt, unsigned long long* best) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N && cnt[v]) atomicMax(best, ((unsigned long long)cnt[v] << 32) | (unsigned long long)(~v));
}
__global__ void k_cc_apply_views(uint32_t N, const uint32_t* __restrict__ label, const uint32_t* __restrict__ touched,
                                 const unsigned long long* __restrict__ best, uint8_t* __restrict__ view_keep) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const uint32_t b = ~(uint32_t)(*best & 0xffffffffull);
  view_keep[v] = (*best != 0ull && touched[v] && label[v] == b) ? 1 : 0;
}
__global__ void k_cc_apply_edges(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej,
                                 const uint8_t* __restrict__ view_keep, uint8_t* __restrict__ keep) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) keep[e] = (keep[e] && view_keep[ei[e]] && view_keep[ej[e]]) ? 1 : 0;
}

// ---- Boruvka --------------------------------------------------------------------------------
// keys for the rank sort: first by (i, j) (64 bit), then stably by the inverted weight (32 bit)
__global__ void k_mst_keys(uint64_t E, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej, const int32_t* __restrict__ w,
                           unsigned long long* __restrict__ key_ij, uint32_t* __restrict__ key_w, uint32_t* __restrict__ ids) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  key_ij[e] = ((unsigned long long)ei[e] << 32) | ej[e];
  // weight descending == (0x7fffffff - w) ascending; negative weights sort last
  key_w[e] = (uint32_t)(0x7fffffff - (w[e] < 0 ? -1 : w[e])) ;
  ids[e] = (uint32_t)e;
}
__global__ void k_gather_u32(uint64_t n, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[idx[k]];
}
// every component's best (lowest rank) outgoing edge; by_rank[r] = edge id of rank r
__global__ void k_mst_pick(uint64_t E, const uint32_t* __restrict__ by_rank, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej,
                           const uint32_t* __restrict__ comp, uint32_t* best) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= E) return;
  const uint32_t e = by_rank[r];
  const uint32_t cu = comp[ei[e]], cv = comp[ej[e]];
  if (cu == cv) return;
  atomicMin(&best[cu], (uint32_t)r);
  atomicMin(&best[cv], (uint32_t)r);
}
// hook every component that picked an edge under the component at the edge's other end; a mutual pick (both ends chose
// the same edge) keeps the smaller root.  Marks the tree edges.
__global__ void k_mst_hook(uint32_t N, const uint32_t* __restrict__ by_rank, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej,
                           const uint32_t* __restrict__ comp, const uint32_t* __restrict__ best, uint32_t* parent, uint8_t* __restrict__ in_tree,
                           uint32_t* changed) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N || comp[c] != c || best[c] == kNone) return;  // roots only
  const uint32_t e = by_rank[best[c]];
  const uint32_t cu = comp[ei[e]], cv = comp[ej[e]];
  const uint32_t other = (cu == c) ? cv : cu;
  in_tree[e] = 1;
  *changed = 1u;
  if (best[other] == best[c] && c < other) return;  // mutual pick: this root stays
  parent[c] = other;
}
__global__ void k_mst_relabel(uint32_t N, const uint32_t* __restrict__ parent, uint32_t* __restrict__ comp) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) comp[v] = find_root(parent, comp[v]);
}
__global__ void k_compact_tree(uint64_t E, const uint8_t* __restrict__ in_tree, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ ei,
                               const uint32_t* __restrict__ ej, uint32_t* __restrict__ tree_edges, uint32_t* __restrict__ tree_lo,
                               uint32_t* __restrict__ tree_hi) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E || !in_tree[e]) return;
  const uint32_t a = ei[e], b = ej[e], t = pos[e];
  tree_edges[t] = (uint32_t)e;
  tree_lo[t] = a < b ? a : b;
  tree_hi[t] = a < b ? b : a;
}
__global__ void k_flag_u8_to_u32(uint64_t n, const uint8_t* __restrict__ f, uint32_t* __restrict__ o) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) o[k] = f[k] ? 1u : 0u;
}

// ---- orientation propagation over the tree (one CTA, level synchronous) --------------------------
// q[v] = unit quaternion (w, x, y, z) of R_v; done[v] flags.  An edge (i, j), i < j, stores R_j = R_rel R_i
// (TwoViewInfo::rotation_2, T/sfm/twoview_info.h:123-126): reaching j from i multiplies by q_rel, i from j by its conjugate.
__global__ void __launch_bounds__(1024) k_tree_propagate(uint32_t T, const uint32_t* __restrict__ tree_edges, const uint32_t* __restrict__ tree_lo,
                                                         const uint32_t* __restrict__ tree_hi, uint8_t* __restrict__ used,
                                                         const double* __restrict__ omega_ij, double* q, uint32_t* done, uint32_t* rounds_out) {
  __shared__ int progress;
  uint32_t rounds = 0;
  while (true) {
    if (threadIdx.x == 0) progress = 0;
    __syncthreads();
    int mine = 0;
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
      if (used[t]) continue;  // this thread's own earlier write
      const uint32_t lo = tree_lo[t], hi = tree_hi[t];
      const uint32_t dl = ((volatile uint32_t*)done)[lo], dh = ((volatile uint32_t*)done)[hi];
      // level synchronous: only views finished in an EARLIER round (flag < current round + 2) may be sources
      const bool l_ok = dl != 0u && dl <= rounds + 1u, h_ok = dh != 0u && dh <= rounds + 1u;
      if (l_ok == h_ok) continue;
      const uint32_t e = tree_edges[t];
      used[t] = 1;
      const Q4 qr = aa_to_quat(omega_ij[3 * (size_t)e], omega_ij[3 * (size_t)e + 1], omega_ij[3 * (size_t)e + 2]);
      const uint32_t src = l_ok ? lo : hi, dst = l_ok ? hi : lo;
      const Q4 qs{q[4 * (size_t)src], q[4 * (size_t)src + 1], q[4 * (size_t)src + 2], q[4 * (size_t)src + 3]};
      const Q4 qd = l_ok ? qmul(qr, qs) : qmul(qconj(qr), qs);
      q[4 * (size_t)dst] = qd.w; q[4 * (size_t)dst + 1] = qd.x; q[4 * (size_t)dst + 2] = qd.y; q[4 * (size_t)dst + 3] = qd.z;
      __threadfence_block();
      ((volatile uint32_t*)done)[dst] = rounds + 2u;
      mine = 1;
    }
    if (mine) progress = 1;
    __syncthreads();
    const int any = progress;
    __syncthreads();
    ++rounds;
    if (!any) break;
  }
  if (threadIdx.x == 0 && rounds_out) *rounds_out = rounds;
}
__global__ void k_quat_to_omega(uint32_t N, const double* __restrict__ q, const uint32_t* __restrict__ done, double* __restrict__ omega) {
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  if (!done[v]) {
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    omega[3 * (size_t)v] = nan; omega[3 * (size_t)v + 1] = nan; omega[3 * (size_t)v + 2] = nan;
    return;
  }
  double e[3], t2, c;
  quat_log(Q4{q[4 * (size_t)v], q[4 * (size_t)v + 1], q[4 * (size_t)v + 2], q[4 * (size_t)v + 3]}, e, &t2, &c);
  omega[3 * (size_t)v] = e[0]; omega[3 * (size_t)v + 1] = e[1]; omega[3 * (size_t)v + 2] = e[2];
}

}  // namespace graphk

namespace {

// Labels of the connected components over the kept edges: label[v] = smallest view index of v's component.
int connected_components(uint32_t N, uint64_t E, const uint32_t* d_ei, const uint32_t* d_ej, const uint8_t* d_keep, uint32_t* d_label,
                         uint32_t* d_flag, cudaStream_t st, int* rounds_out) {
  graphk::k_iota<<<grid_for(N), kBlock, 0, st>>>(N, d_label);
  int rounds = 0;
  while (true) {
    CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st));
    graphk::k_cc_hook<<<grid_for(E), kBlock, 0, st>>>(E, d_ei, d_ej, d_keep, d_label, d_flag);
    graphk::k_cc_compress<<<grid_for(N), kBlock, 0, st>>>(N, d_label);
    uint32_t h = 0;
    CUDA_TRY(cudaMemcpyAsync(&h, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ++rounds;
    if (!h) break;
    if (rounds > 64) { set_error("connected components did not converge"); return GSFM_RA_ERR_NUMERIC; }
  }
  if (rounds_out) *rounds_out = rounds;
  return 0;
}

int check_edges_host(uint32_t N, uint64_t E, const uint32_t* ei, const uint32_t* ej) {
  if (N == 0) { set_error("no views"); return GSFM_RA_ERR_INVALID; }
  if (E >= 0x7fffffffull) { set_error("at most 2^31 - 2 edges"); return GSFM_RA_ERR_UNSUPPORTED; }
  if (E && (!ei || !ej)) { set_error("NULL edge arrays"); return GSFM_RA_ERR_INVALID; }
  for (uint64_t e = 0; e < E; ++e)
    if (ei[e] >= N || ej[e] >= N || ei[e] == ej[e]) { set_error("edge %llu is out of range or a self loop", (unsigned long long)e); return GSFM_RA_ERR_INVALID; }
  return 0;
}

}  // namespace

extern "C" {

int gsfm_ra_filter_initial_view_graph(uint32_t num_views, uint64_t num_edges, const uint32_t* edge_i, const uint32_t* edge_j,
                                      const int32_t* num_verified_matches, int32_t min_num_two_view_inliers, uint8_t* edge_keep,
                                      uint8_t* view_keep, int32_t device) {
  if (!edge_keep || !view_keep) { set_error("NULL output"); return GSFM_RA_ERR_INVALID; }
  RA_TRY(check_edges_host(num_views, num_edges, edge_i, edge_j));
  if (num_edges && !num_verified_matches) { set_error("num_verified_matches is NULL"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  const uint32_t N = num_views;
  const uint64_t E = num_edges;
  if (E == 0) { std::memset(view_keep, 0, N); return 0; }
  StreamHolder sh;
  CUDA_TRY(cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking));
  cudaStream_t st = sh.s;
  {
    AllocScope scope(st);
    DevBuf<uint32_t> d_ei, d_ej, label, touched, cnt, flag;
    DevBuf<int32_t> d_m;
    DevBuf<uint8_t> d_keep, d_vkeep;
    DevBuf<unsigned long long> best;
    RA_TRY(d_ei.alloc(E)); RA_TRY(d_ej.alloc(E)); RA_TRY(d_m.alloc(E)); RA_TRY(d_keep.alloc(E));
    RA_TRY(label.alloc(N)); RA_TRY(touched.alloc(N)); RA_TRY(cnt.alloc(N)); RA_TRY(flag.alloc(1)); RA_TRY(d_vkeep.alloc(N)); RA_TRY(best.alloc(1));
    CUDA_TRY(cudaMemcpyAsync(d_ei.p, edge_i, E * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ej.p, edge_j, E * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_m.p, num_verified_matches, E * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(touched.p, 0, (size_t)N * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(cnt.p, 0, (size_t)N * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(best.p, 0, sizeof(unsigned long long), st));
    graphk::k_filter_edges<<<grid_for(E), kBlock, 0, st>>>(E, d_m.p, min_num_two_view_inliers, d_keep.p);
    graphk::k_touch<<<grid_for(E), kBlock, 0, st>>>(E, d_ei.p, d_ej.p, d_keep.p, touched.p);
    RA_TRY(connected_components(N, E, d_ei.p, d_ej.p, d_keep.p, label.p, flag.p, st, nullptr));
    graphk::k_cc_count<<<grid_for(N), kBlock, 0, st>>>(N, label.p, touched.p, cnt.p);
    graphk::k_cc_best<<<grid_for(N), kBlock, 0, st>>>(N, cnt.p, best.p);
    graphk::k_cc_apply_views<<<grid_for(N), kBlock, 0, st>>>(N, label.p, touched.p, best.p, d_vkeep.p);
    graphk::k_cc_apply_edges<<<grid_for(E), kBlock, 0, st>>>(E, d_ei.p, d_ej.p, d_vkeep.p, d_keep.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(edge_keep, d_keep.p, E, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(view_keep, d_vkeep.p, N, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return 0;
}

int gsfm_ra_init_orientations_mst(uint32_t num_views, uint64_t num_edges, const uint32_t* edge_i, const uint32_t* edge_j,
                                  const double* omega_ij, const int32_t* edge_weight, int64_t root, double* omega_out,
                                  uint8_t* edge_in_tree, int32_t* rounds_out, int32_t device) {
  if (!omega_out) { set_error("omega_out is NULL"); return GSFM_RA_ERR_INVALID; }
  RA_TRY(check_edges_host(num_views, num_edges, edge_i, edge_j));
  if (num_edges && (!omega_ij || !edge_weight)) { set_error("omega_ij / edge_weight is NULL"); return GSFM_RA_ERR_INVALID; }
  if (root >= (int64_t)num_views) { set_error("root is not a view"); return GSFM_RA_ERR_INVALID; }
  int dev;
  RA_TRY(select_device(device, &dev));
  const uint32_t N = num_views;
  const uint64_t E = num_edges;
  if (root < 0) {  // the smallest view index that has an edge (the host restatement's choice)
    uint32_t r = N;
    for (uint64_t e = 0; e < E; ++e) r = std::min(r, std::min(edge_i[e], edge_j[e]));
    root = (r == N) ? 0 : r;
  }
  StreamHolder sh;
  CUDA_TRY(cudaStreamCreateWithFlags(&sh.s, cudaStreamNonBlocking));
  cudaStream_t st = sh.s;
  int rounds = 0;
  {
    AllocScope scope(st);
    DevBuf<uint32_t> d_ei, d_ej, ids_a, ids_b, kw_a, kw_b, comp, parent, best, flag, tree_edges, tflag, tpos, done, d_rounds;
    DevBuf<int32_t> d_w;
    DevBuf<unsigned long long> kij_a, kij_b;
    DevBuf<uint32_t> tree_lo, tree_hi;
    DevBuf<uint8_t> in_tree, used;
    DevBuf<double> d_wij, q, d_omega;
    RA_TRY(comp.alloc(N)); RA_TRY(parent.alloc(N)); RA_TRY(best.alloc(N)); RA_TRY(flag.alloc(1)); RA_TRY(done.alloc(N)); RA_TRY(d_rounds.alloc(1));
    RA_TRY(q.alloc(4ull * N)); RA_TRY(d_omega.alloc(3ull * N));
    CUDA_TRY(cudaMemsetAsync(done.p, 0, (size_t)N * sizeof(uint32_t), st));
    CUDA_TRY(cudaMemsetAsync(q.p, 0, 4ull * N * sizeof(double), st));
    uint32_t T = 0;
    if (E) {
      RA_TRY(d_ei.alloc(E)); RA_TRY(d_ej.alloc(E)); RA_TRY(d_w.alloc(E)); RA_TRY(d_wij.alloc(3 * E)); RA_TRY(in_tree.alloc(E));
      RA_TRY(ids_a.alloc(E)); RA_TRY(ids_b.alloc(E)); RA_TRY(kw_a.alloc(E)); RA_TRY(kw_b.alloc(E)); RA_TRY(kij_a.alloc(E)); RA_TRY(kij_b.alloc(E));
      RA_TRY(tflag.alloc(E + 1)); RA_TRY(tpos.alloc(E + 1)); RA_TRY(tree_edges.alloc(N)); RA_TRY(tree_lo.alloc(N)); RA_TRY(tree_hi.alloc(N));
      RA_TRY(used.alloc(N));
      CUDA_TRY(cudaMemsetAsync(used.p, 0, N, st));
      CUDA_TRY(cudaMemcpyAsync(d_ei.p, edge_i, E * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_ej.p, edge_j, E * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_w.p, edge_weight, E * sizeof(int32_t), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemcpyAsync(d_wij.p, omega_ij, 3 * E * sizeof(double), cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaMemsetAsync(in_tree.p, 0, E, st));
      // rank of every edge in the order (weight desc, i, j): stable LSD -- sort by (i, j), then by the inverted weight
      graphk::k_mst_keys<<<grid_for(E), kBlock, 0, st>>>(E, d_ei.p, d_ej.p, d_w.p, kij_a.p, kw_a.p, ids_a.p);
      {
        size_t bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kij_a.p, kij_b.p, ids_a.p, ids_b.p, (int)E, 0, 64, st));
        DevBuf<unsigned char> tmp;
        RA_TRY(tmp.alloc(bytes + 16));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kij_a.p, kij_b.p, ids_a.p, ids_b.p, (int)E, 0, 64, st));
      }
      graphk::k_gather_u32<<<grid_for(E), kBlock, 0, st>>>(E, ids_b.p, kw_a.p, kw_b.p);  // weights in (i, j) order
      {
        size_t bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kw_b.p, kw_a.p, ids_b.p, ids_a.p, (int)E, 0, 32, st));
        DevBuf<unsigned char> tmp;
        RA_TRY(tmp.alloc(bytes + 16));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kw_b.p, kw_a.p, ids_b.p, ids_a.p, (int)E, 0, 32, st));
      }
      const uint32_t* by_rank = ids_a.p;  // by_rank[r] = edge id of rank r
      graphk::k_iota<<<grid_for(N), kBlock, 0, st>>>(N, comp.p);
      graphk::k_iota<<<grid_for(N), kBlock, 0, st>>>(N, parent.p);
      while (true) {
        graphk::k_fill<<<grid_for(N), kBlock, 0, st>>>(N, best.p, graphk::kNone);
        CUDA_TRY(cudaMemsetAsync(flag.p, 0, sizeof(uint32_t), st));
        graphk::k_mst_pick<<<grid_for(E), kBlock, 0, st>>>(E, by_rank, d_ei.p, d_ej.p, comp.p, best.p);
        graphk::k_mst_hook<<<grid_for(N), kBlock, 0, st>>>(N, by_rank, d_ei.p, d_ej.p, comp.p, best.p, parent.p, in_tree.p, flag.p);
        graphk::k_mst_relabel<<<grid_for(N), kBlock, 0, st>>>(N, parent.p, comp.p);
        uint32_t h = 0;
        CUDA_TRY(cudaMemcpyAsync(&h, flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!h) break;
        if (++rounds > 64) { set_error("Boruvka did not converge"); return GSFM_RA_ERR_NUMERIC; }
      }
      // compact the tree edges (edge order), then propagate from the root
      graphk::k_flag_u8_to_u32<<<grid_for(E), kBlock, 0, st>>>(E, in_tree.p, tflag.p);
      CUDA_TRY(cudaMemsetAsync(tflag.p + E, 0, sizeof(uint32_t), st));
      RA_TRY(exclusive_scan(tflag.p, tpos.p, E + 1, st));
      CUDA_TRY(cudaMemcpyAsync(&T, tpos.p + E, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (T >= N) { set_error("spanning forest has %u edges for %u views", T, N); return GSFM_RA_ERR_NUMERIC; }
      graphk::k_compact_tree<<<grid_for(E), kBlock, 0, st>>>(E, in_tree.p, tpos.p, d_ei.p, d_ej.p, tree_edges.p, tree_lo.p, tree_hi.p);
    }
    // root: identity, flag 1
    const double qid[4] = {1.0, 0.0, 0.0, 0.0};
    const uint32_t one = 1u;
    CUDA_TRY(cudaMemcpyAsync(q.p + 4ull * (uint64_t)root, qid, sizeof(qid), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(done.p + root, &one, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    if (T) graphk::k_tree_propagate<<<1, 1024, 0, st>>>(T, tree_edges.p, tree_lo.p, tree_hi.p, used.p, d_wij.p, q.p, done.p, d_rounds.p);
    graphk::k_quat_to_omega<<<grid_for(N), kBlock, 0, st>>>(N, q.p, done.p, d_omega.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(omega_out, d_omega.p, 3ull * N * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (edge_in_tree && E) CUDA_TRY(cudaMemcpyAsync(edge_in_tree, in_tree.p, E, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (rounds_out) *rounds_out = rounds;
  return 0;
}

}  // extern "C"
