// gsfm_io.cpp -- native readers / writer for the on-disk formats either side of the rotation-averaging path
// (SURVEY.md section 8 f, rank 3).  Host-only C++ (no CUDA), part of libgsfm_ra.so, C ABI in include/gsfm_ra.h.
//
//   covariance_rot.txt   reference src/uncertainty.cpp:200-229 (read_covariance), :164-198 (store_covariance_rot):
//                        two header lines, then `id1 id2` + 9 doubles bit-cast to uint64 and printed in decimal
//                        (C00 C11 C22 C01 C02 C12 R0 R1 R2)
//   1DSfM dataset        thirdparty/TheiaSfM/src/theia/io/read_1dsfm.cc:93-412: cc.txt (views of the largest connected
//                        component), list.txt (one view per line, the line index is the view id), tracks.txt (per track
//                        `n (view feature) x n`), EGs.txt (`id1 id2 R[9 row-major] t[3]`)
//                        rotation_2 = angle-axis of S R^T S, position_2 = S t, S = diag(1, -1, -1)   (:309-333)
//                        num_verified_matches of a pair = number of tracks that see both views (as the Python host side
//                        of this repo derives it; Theia leaves the count to the matcher)
// Arrays are returned in malloc'ed memory the caller releases with gsfm_ra_free.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/gsfm_ra.h"

namespace gsfm_io {

void set_io_error(const std::string& msg);  // defined in gsfm_ra.cu (thread-local last error)

namespace {

bool read_file(const std::string& path, std::string* out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  const size_t got = n > 0 ? std::fread(&(*out)[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  out->resize(got);
  return true;
}

// whitespace-separated token scanner over a buffer
struct Scanner {
  const char* p;
  const char* end;
  explicit Scanner(const std::string& s) : p(s.data()), end(s.data() + s.size()) {}
  void skip_ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p; }
  void skip_line() { while (p < end && *p != '\n') ++p; if (p < end) ++p; }
  bool at_end() { skip_ws(); return p >= end; }
  bool next_u64(uint64_t* v) {
    skip_ws();
    if (p >= end || *p < '0' || *p > '9') return false;
    char* q = nullptr;
    *v = std::strtoull(p, &q, 10);
    if (q == p) return false;
    p = q;
    return true;
  }
  bool next_double(double* v) {
    skip_ws();
    if (p >= end) return false;
    char* q = nullptr;
    *v = std::strtod(p, &q);
    if (q == p) return false;
    p = q;
    return true;
  }
};

template <typename T>
T* to_malloc(const std::vector<T>& v) {
  T* out = (T*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
  return out;
}

// ceres::RotationMatrixToAngleAxis = RotationMatrixToQuaternion (trace / largest-diagonal branches) followed by
// QuaternionToAngleAxis (angle in [0, pi]); R row-major.
void matrix_to_angle_axis(const double* R, double* w) {
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i + 1] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j + 1] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k + 1] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = std::sqrt(s2);
    const double two_theta = 2.0 * (q[0] < 0.0 ? std::atan2(-s, -q[0]) : std::atan2(s, q[0]));
    const double k = two_theta / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    w[0] = q[1] * 2.0; w[1] = q[2] * 2.0; w[2] = q[3] * 2.0;
  }
}

struct PairHash {
  size_t operator()(uint64_t k) const { return (size_t)(k * 0x9e3779b97f4a7c15ull >> 16); }
};

}  // namespace
}  // namespace gsfm_io

using namespace gsfm_io;

extern "C" {

void gsfm_ra_free(void* p) { std::free(p); }

int gsfm_ra_read_covariance_rot(const char* path, uint64_t* count, uint32_t** view_id1, uint32_t** view_id2, double** cov6, double** rot) {
  if (!path || !count || !view_id1 || !view_id2 || !cov6 || !rot) { set_io_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  std::string buf;
  if (!read_file(path, &buf)) { set_io_error(std::string("cannot open ") + path); return GSFM_RA_ERR_INVALID; }
  Scanner sc(buf);
  sc.skip_line();
  sc.skip_line();  // two header lines (uncertainty.cpp:208-209)
  std::vector<uint32_t> a, b;
  std::vector<double> c6, r3;
  while (!sc.at_end()) {
    uint64_t i1, i2, bits[9];
    if (!sc.next_u64(&i1) || !sc.next_u64(&i2)) { set_io_error(std::string("malformed entry in ") + path); return GSFM_RA_ERR_INVALID; }
    for (int k = 0; k < 9; ++k)
      if (!sc.next_u64(&bits[k])) { set_io_error(std::string("truncated entry in ") + path); return GSFM_RA_ERR_INVALID; }
    a.push_back((uint32_t)i1); b.push_back((uint32_t)i2);
    for (int k = 0; k < 9; ++k) {
      double v;
      std::memcpy(&v, &bits[k], sizeof(v));  // the file stores the IEEE bit pattern
      (k < 6 ? c6 : r3).push_back(v);
    }
  }
  *count = a.size();
  *view_id1 = to_malloc(a); *view_id2 = to_malloc(b); *cov6 = to_malloc(c6); *rot = to_malloc(r3);
  if (!*view_id1 || !*view_id2 || !*cov6 || !*rot) { set_io_error("out of memory"); return GSFM_RA_ERR_INVALID; }
  return 0;
}

int gsfm_ra_write_covariance_rot(const char* path, uint64_t count, const uint32_t* view_id1, const uint32_t* view_id2, const double* cov6,
                                 const double* rot) {
  if (!path || (count && (!view_id1 || !view_id2 || !cov6 || !rot))) { set_io_error("NULL argument"); return GSFM_RA_ERR_INVALID; }
  FILE* f = std::fopen(path, "w");
  if (!f) { set_io_error(std::string("cannot create ") + path); return GSFM_RA_ERR_INVALID; }
  std::fprintf(f, "# Stored as uint64, should convert to double first.\n# view_id1 view_id2 C00 C11 C22 C01 C02 C12 R0 R1 R2\n");
  for (uint64_t e = 0; e < count; ++e) {
    std::fprintf(f, "%u %u", view_id1[e], view_id2[e]);
    for (int k = 0; k < 9; ++k) {
      const double v = k < 6 ? cov6[6 * e + k] : rot[3 * e + (k - 6)];
      uint64_t bits;
      std::memcpy(&bits, &v, sizeof(bits));
      std::fprintf(f, " %llu", (unsigned long long)bits);
    }
    std::fprintf(f, " \n");
  }
  std::fclose(f);
  return 0;
}

int gsfm_ra_read_1dsfm(const char* dataset_directory, uint32_t* num_listed_views, uint64_t* num_views, uint32_t** view_ids,
                       double** focal_length_priors, uint64_t* num_pairs, uint32_t** view_id1, uint32_t** view_id2, double** rotation_2,
                       double** position_2, int32_t** num_verified_matches) {
  if (!dataset_directory || !num_views || !view_ids || !num_pairs || !view_id1 || !view_id2 || !rotation_2) {
    set_io_error("NULL argument");
    return GSFM_RA_ERR_INVALID;
  }
  const std::string dir = std::string(dataset_directory) + "/";
  std::string buf;
  // cc.txt: the views to keep (read_1dsfm.cc:93-111)
  if (!read_file(dir + "cc.txt", &buf)) { set_io_error("cannot open " + dir + "cc.txt"); return GSFM_RA_ERR_INVALID; }
  std::unordered_set<uint32_t> cc;
  {
    Scanner sc(buf);
    uint64_t v;
    while (sc.next_u64(&v)) cc.insert((uint32_t)v);
  }
  // list.txt: line index = view id; optional "0 focal" after the image name (:113-160)
  if (!read_file(dir + "list.txt", &buf)) { set_io_error("cannot open " + dir + "list.txt"); return GSFM_RA_ERR_INVALID; }
  std::vector<uint32_t> ids;
  std::vector<double> focal;
  uint32_t listed = 0;
  {
    size_t pos = 0;
    while (pos < buf.size()) {
      size_t eol = buf.find('\n', pos);
      if (eol == std::string::npos) eol = buf.size();
      const std::string line = buf.substr(pos, eol - pos);
      pos = eol + 1;
      if (line.find_first_not_of(" \t\r") == std::string::npos) continue;
      const uint32_t vid = listed++;
      if (!cc.count(vid)) continue;
      double f = 0.0;
      {
        Scanner ls(line);
        ls.skip_ws();
        while (ls.p < ls.end && *ls.p != ' ' && *ls.p != '\t') ++ls.p;  // the image name
        double flag;
        if (ls.next_double(&flag) && !ls.next_double(&f)) f = 0.0;
      }
      ids.push_back(vid);
      focal.push_back(f);
    }
  }
  std::unordered_set<uint32_t> alive(ids.begin(), ids.end());
  // EGs.txt (:299-373)
  if (!read_file(dir + "EGs.txt", &buf)) { set_io_error("cannot open " + dir + "EGs.txt"); return GSFM_RA_ERR_INVALID; }
  std::vector<uint32_t> a, b;
  std::vector<double> rot, posv;
  std::unordered_map<uint64_t, uint32_t, PairHash> pair_index;
  {
    Scanner sc(buf);
    while (!sc.at_end()) {
      uint64_t i1, i2;
      double v[12];
      if (!sc.next_u64(&i1) || !sc.next_u64(&i2)) { set_io_error("malformed line in " + dir + "EGs.txt"); return GSFM_RA_ERR_INVALID; }
      for (int k = 0; k < 12; ++k)
        if (!sc.next_double(&v[k])) { set_io_error("truncated line in " + dir + "EGs.txt"); return GSFM_RA_ERR_INVALID; }
      if (!alive.count((uint32_t)i1) || !alive.count((uint32_t)i2)) continue;
      // rotation = S R^T S: entry (r, c) = s_r s_c R(c, r)
      static const double s[3] = {1.0, -1.0, -1.0};
      double M[9], w[3];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[3 * r + c] = s[r] * s[c] * v[3 * c + r];
      matrix_to_angle_axis(M, w);
      if (i1 == i2) continue;  // ViewGraph::AddEdge ignores self loops (view_graph.cc:133-151)
      const uint64_t key = ((uint64_t)std::min(i1, i2) << 32) | std::max(i1, i2);
      const auto found = pair_index.find(key);
      if (found != pair_index.end()) {  // a repeated pair overwrites the stored TwoViewInfo, as edges_[pair] = info does
        const uint32_t at = found->second;
        a[at] = (uint32_t)i1; b[at] = (uint32_t)i2;
        for (int k = 0; k < 3; ++k) { rot[3 * at + k] = w[k]; posv[3 * at + k] = s[k] * v[9 + k]; }
        continue;
      }
      pair_index[key] = (uint32_t)a.size();
      a.push_back((uint32_t)i1); b.push_back((uint32_t)i2);
      for (int k = 0; k < 3; ++k) { rot.push_back(w[k]); posv.push_back(s[k] * v[9 + k]); }
    }
  }
  // tracks.txt (:162-297): per track the views that see it; a pair's verified matches = tracks seen by both
  std::vector<int32_t> matches(a.size(), 0);
  if (num_verified_matches) {
    if (!read_file(dir + "tracks.txt", &buf)) { set_io_error("cannot open " + dir + "tracks.txt"); return GSFM_RA_ERR_INVALID; }
    Scanner sc(buf);
    uint64_t n_tracks = 0;
    if (!sc.next_u64(&n_tracks)) { set_io_error("malformed header in " + dir + "tracks.txt"); return GSFM_RA_ERR_INVALID; }
    std::vector<uint32_t> views;
    for (uint64_t t = 0; t < n_tracks; ++t) {
      uint64_t n;
      if (!sc.next_u64(&n)) break;
      views.clear();
      for (uint64_t k = 0; k < n; ++k) {
        uint64_t v, feat;
        if (!sc.next_u64(&v) || !sc.next_u64(&feat)) { set_io_error("truncated track in " + dir + "tracks.txt"); return GSFM_RA_ERR_INVALID; }
        views.push_back((uint32_t)v);
      }
      std::sort(views.begin(), views.end());
      if (std::adjacent_find(views.begin(), views.end()) != views.end()) continue;  // a view seen twice: AddTrack rejects the track
      for (size_t x = 0; x < views.size(); ++x)
        for (size_t y = x + 1; y < views.size(); ++y) {
          const auto it = pair_index.find(((uint64_t)views[x] << 32) | views[y]);
          if (it != pair_index.end()) ++matches[it->second];
        }
    }
  }
  if (num_listed_views) *num_listed_views = listed;
  *num_views = ids.size();
  *view_ids = to_malloc(ids);
  if (focal_length_priors) *focal_length_priors = to_malloc(focal);
  *num_pairs = a.size();
  *view_id1 = to_malloc(a); *view_id2 = to_malloc(b); *rotation_2 = to_malloc(rot);
  if (position_2) *position_2 = to_malloc(posv);
  if (num_verified_matches) *num_verified_matches = to_malloc(matches);
  return 0;
}

}  // extern "C"
