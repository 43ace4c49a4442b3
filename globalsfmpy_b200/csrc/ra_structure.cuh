// ra_structure.cuh -- K0: half-edge records, CSR structure and balanced warp partitions, built on the device
// Part of libgsfm_ra (one translation unit, see gsfm_ra.cu); reference citations sit next to each kernel.
#pragma once
#include "ra_common.cuh"
namespace {

// ------------------------------------------------------------------------------------------
// K0 setup: per half-edge, gather the edge's measurement and weight into K1's INPUT RECORDS, one record per 32
// consecutive half-edges:
//   { double qij[4][32]; double U[kU][32]; uint32_t col[32]; uint32_t row[32]; }      128 B aligned
//   qij = unit quaternion of omega_ij, U = whitening (rotation_estimator.cpp:251-288), kU = 6 (upper triangle) or 1
//   (scalar weight); col carries the side bit.  1536 B (kU = 1) / 2816 B (kU = 6) per record: K1 pulls a record with
//   ONE bulk async copy.
// ------------------------------------------------------------------------------------------
__device__ __host__ __forceinline__ int in_rec_doubles(int ku) { return (4 + ku) * 32 + 32; }

__global__ void k_setup_halfedges(uint64_t H, int ku, const uint32_t* __restrict__ he_edge, const uint32_t* __restrict__ he_row,
                                  const uint32_t* __restrict__ he_col, const double* __restrict__ omega_ij, const double* __restrict__ cov6,
                                  const double* __restrict__ weight, int error_type, double* __restrict__ inrec,
                                  const uint32_t* __restrict__ ei = nullptr, const double* __restrict__ orientation = nullptr) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const uint64_t k = he_edge[h];
  // translation averaging (orientation != null): the measurement slot carries the pair's translation direction rotated into
  // the global frame by the FIRST view's orientation (position_estimator.cpp:271-272), omega_ij holds TwoViewInfo::position_2
  const Q4 q = orientation ? rotated_translation(orientation + 3 * (size_t)ei[k], omega_ij + 3 * k)
                           : aa_to_quat(omega_ij[3 * k], omega_ij[3 * k + 1], omega_ij[3 * k + 2]);
  double* rec = inrec + (size_t)(h >> 5) * in_rec_doubles(ku);
  const int lane = (int)(h & 31);
  rec[lane] = q.w; rec[32 + lane] = q.x; rec[64 + lane] = q.y; rec[96 + lane] = q.z;
  double c6[6] = {0, 0, 0, 0, 0, 0};
  if (cov6) for (int t = 0; t < 6; ++t) c6[t] = cov6[6 * k + t];
  double u[6];
  whiten(error_type, c6, weight ? weight[k] : 1.0, u);
  for (int t = 0; t < ku; ++t) rec[(4 + t) * 32 + lane] = u[t];
  uint32_t* idx = reinterpret_cast<uint32_t*>(rec + (4 + ku) * 32);
  idx[lane] = he_col[h];
  idx[32 + lane] = he_row[h];
}

// ------------------------------------------------------------------------------------------
// Structure build on the device (one-time per problem): half-edge keys -> radix sort -> pieces,
// piece pointers, duplicate / range checks, balanced warp partitions and their segments.
//
// COLUMN BLOCKS.  The views are cut into ncb blocks of cbsize consecutive views and the half-edges are sorted by
// (column block of col, row, col).  A PIECE is the run of half-edges of one row inside one column block; piece id
// p = cb * N + row, so everything that used to be "per row" in the partition machinery is "per piece" and a row's pieces are
// p = row, N + row, 2N + row, ...  Why: the per-half-edge gather x[col] of the SpMV moves 32 B per half-edge through the L2,
// as much as the matrix stream itself; with the columns of a CTA's range confined to one block, that block's slice of x
// (<= kSliceMaxViews * 24 B) is staged in shared memory once per pass and the gather never leaves the SM.  ncb = 1 (graphs of
// more than kMaxColBlocks * kSliceMaxViews views, where the matrix stream is HBM-bound anyway) is the plain (row, col) order.
// ------------------------------------------------------------------------------------------
constexpr int kMaxColBlocks = 8;
struct ColBlocks {
  uint32_t ncb, cbsize;                     // column block of view v: v / cbsize
  uint32_t begin[kMaxColBlocks + 1];        // first half-edge of every block (begin[ncb] = H)
};
__device__ __host__ __forceinline__ uint32_t piece_of(uint32_t row, uint32_t col, uint32_t N, uint32_t cbsize) { return (col / cbsize) * N + row; }

__global__ void k_build_keys(uint64_t E, uint32_t N, uint32_t cbsize, const uint32_t* __restrict__ ei, const uint32_t* __restrict__ ej,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, int* err) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= E) return;
  uint32_t i = ei[k], j = ej[k];
  if (i >= N || j >= N || i == j) { atomicMax(err, 1); i = 0; j = (N > 1) ? 1 : 0; }
  keys[2 * k] = (uint64_t)piece_of(i, j, N, cbsize) * N + j;     vals[2 * k] = (uint32_t)k;
  keys[2 * k + 1] = (uint64_t)piece_of(j, i, N, cbsize) * N + i; vals[2 * k + 1] = (uint32_t)k | kSideBit;
}
__global__ void k_unpack_keys(uint64_t H, uint32_t N, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint32_t* __restrict__ he_row,
                              uint32_t* __restrict__ he_col, uint32_t* __restrict__ he_edge, int* err) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const uint64_t key = keys[h];
  const uint32_t v = vals[h];
  he_row[h] = (uint32_t)((key / N) % N);
  he_col[h] = (uint32_t)(key % N) | (v & kSideBit);
  he_edge[h] = v & ~kSideBit;
  if (h > 0 && keys[h - 1] == key) atomicMax(err, 2);  // the same view pair twice
}
// pieceptr[p] = first half-edge of piece p, p = 0 .. NP (NP = ncb * N pieces)
__global__ void k_pieceptr(uint32_t NP, uint32_t N, uint64_t H, const uint64_t* __restrict__ keys, uint32_t* __restrict__ pieceptr) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > NP) return;
  const uint64_t target = (uint64_t)p * N;
  uint64_t lo = 0, hi = H;
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
  pieceptr[p] = (uint32_t)lo;
}
__global__ void k_piece_flags(uint32_t NP, const uint32_t* __restrict__ pieceptr, uint32_t* __restrict__ nonempty) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > NP) return;
  nonempty[p] = (p < NP && pieceptr[p + 1] > pieceptr[p]) ? 1u : 0u;
}
// views without any half-edge (in this shard): every piece of the row is empty
__global__ void k_iso_flags(uint32_t N, uint32_t ncb, const uint32_t* __restrict__ pieceptr, uint32_t* __restrict__ isoflag) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > N) return;
  uint32_t any = 0;
  if (r < N)
    for (uint32_t cb = 0; cb < ncb; ++cb) any |= (pieceptr[cb * N + r + 1] > pieceptr[cb * N + r]) ? 1u : 0u;
  isoflag[r] = (r < N) ? 1u - any : 0u;
}
__global__ void k_iso_fill(uint32_t N, const uint32_t* __restrict__ isoflag, const uint32_t* __restrict__ iso_rank, uint32_t* __restrict__ iso) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < N && isoflag[r]) iso[iso_rank[r]] = r;
}
__global__ void k_part_count(uint32_t nw, uint32_t per, uint64_t H, uint32_t N, uint32_t cbsize, const uint32_t* __restrict__ he_row,
                             const uint32_t* __restrict__ he_col, const uint32_t* __restrict__ nz_rank, uint32_t* __restrict__ nseg) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nw) return;
  if (w == nw) { nseg[w] = 0; return; }
  const uint64_t lo = (uint64_t)w * per, hi = min(H, lo + per);
  const uint32_t p0 = piece_of(he_row[lo], he_col[lo] & ~kSideBit, N, cbsize), p1 = piece_of(he_row[hi - 1], he_col[hi - 1] & ~kSideBit, N, cbsize);
  nseg[w] = nz_rank[p1] - nz_rank[p0] + 1;  // non-empty pieces met by the range = its segments
}
__global__ void k_part_fill(uint32_t nw, uint32_t per, uint64_t H, uint32_t N, uint32_t cbsize, const uint32_t* __restrict__ he_row,
                            const uint32_t* __restrict__ he_col, const uint32_t* __restrict__ pieceptr, const uint32_t* __restrict__ warp_seg_ptr,
                            uint32_t* __restrict__ seg_row, uint32_t* __restrict__ seg_begin, uint32_t* __restrict__ seg_len) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const uint64_t lo = (uint64_t)w * per, hi = min(H, lo + per);
  uint32_t t = warp_seg_ptr[w];
  uint64_t h = lo;
  uint32_t p = piece_of(he_row[lo], he_col[lo] & ~kSideBit, N, cbsize);
  while (h < hi) {
    while (pieceptr[p + 1] <= h) ++p;
    const uint64_t end = min(hi, (uint64_t)pieceptr[p + 1]);
    const uint32_t row = p % N, cb = p / N;
    // bit 31: NOT the first segment of its row -- the piece began in an earlier range, or the row has half-edges in an
    // earlier column block.  The warp that holds a row's first segment owns the row (it adds the pieces and finishes it).
    bool later = h > pieceptr[p];
    for (uint32_t c = 0; c < cb && !later; ++c) later = pieceptr[c * N + row + 1] > pieceptr[c * N + row];
    seg_row[t] = row | (later ? kSideBit : 0u);
    seg_begin[t] = (uint32_t)h;
    seg_len[t] = (uint32_t)(end - h);
    ++t;
    h = end;
  }
}
// segments per piece (the ranges it spans); their exclusive scan, node_seg_ptr[p], is the id of the piece's first segment
__global__ void k_node_seg_count(uint32_t NP, uint32_t per, const uint32_t* __restrict__ pieceptr, uint32_t* __restrict__ cnt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > NP) return;
  uint32_t c = 0;
  if (p < NP && pieceptr[p + 1] > pieceptr[p]) c = (pieceptr[p + 1] - 1) / per - pieceptr[p] / per + 1;
  cnt[p] = c;
}

// Column indices live inside the chunk records of both block buffers (written once).
__global__ void k_embed_cols(uint64_t H, int blk, const uint32_t* __restrict__ he_col, double* rec0, double* rec1) {
  const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const size_t w = (size_t)(h >> 5) * (blk * 32 + 16) + blk * 32;
  reinterpret_cast<uint32_t*>(rec0 + w)[h & 31] = he_col[h];
  reinterpret_cast<uint32_t*>(rec1 + w)[h & 31] = he_col[h];
}

}  // namespace