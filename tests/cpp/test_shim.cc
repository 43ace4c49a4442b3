// Exercises include/gsfm_rotation_estimator.hpp (the C++ mirror of theia::RotationEstimator /
// GSfMNonlinearRotationEstimator) with stand-in container types shaped like Theia's and Eigen's, on a synthetic
// view graph in the pattern of T/sfm/global_pose_estimation/robust_rotation_estimator_test.cc:150-243
// (random orientations scaled by 0.2, relative rotations R_j R_i^T with angular noise, ring + random pairs).
// Usage: test_shim [--expect-no-device]; prints one line per check, exit code 0 = all passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <unistd.h>
#include <unordered_map>

#include "../../include/gsfm_position_estimator.hpp"
#include "../../include/gsfm_rotation_estimator.hpp"

namespace {

typedef uint32_t ViewId;                          // T/sfm/types.h:47
typedef std::pair<ViewId, ViewId> ViewIdPair;
struct PairHash { size_t operator()(const ViewIdPair& p) const { return (size_t)p.first * 1000003u ^ (size_t)p.second; } };
struct Vector3d {                                 // stand-in for Eigen::Vector3d
  double v[3] = {0, 0, 0};
  double& operator[](int k) { return v[k]; }
  const double& operator[](int k) const { return v[k]; }
};
struct Matrix3d {                                 // stand-in for Eigen::Matrix3d
  double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double operator()(int r, int c) const { return m[3 * r + c]; }
  double& operator()(int r, int c) { return m[3 * r + c]; }
};
struct TwoViewInfo { Vector3d rotation_2, position_2; int num_verified_matches = 0; };  // T/sfm/twoview_info.h:54-98
typedef std::unordered_map<ViewIdPair, TwoViewInfo, PairHash> ViewPairs;
typedef std::unordered_map<ViewId, Vector3d> Orientations;
typedef std::unordered_map<ViewIdPair, std::pair<Matrix3d, Vector3d>, PairHash> CovarianceMap;

void Exp(const Vector3d& w, double* R) {
  const double t = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double a = t > 1e-12 ? std::sin(t) / t : 1.0, b = t > 1e-12 ? (1 - std::cos(t)) / (t * t) : 0.5;
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double kk = 0;
      for (int m = 0; m < 3; ++m) kk += K[3 * r + m] * K[3 * m + c];
      R[3 * r + c] = (r == c) + a * K[3 * r + c] + b * kk;
    }
}
Vector3d Log(const double* R) {
  const double c = std::min(1.0, std::max(-1.0, 0.5 * (R[0] + R[4] + R[8] - 1)));
  const double t = std::acos(c);
  Vector3d w;
  const double k = t > 1e-9 ? t / (2 * std::sin(t)) : 0.5;
  w[0] = k * (R[7] - R[5]); w[1] = k * (R[2] - R[6]); w[2] = k * (R[3] - R[1]);
  return w;
}
void Mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
void MulT(const double* A, const double* B, double* C) {  // A B^T
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
}

int g_failed = 0;
void Check(bool ok, const char* what) {
  std::printf("%s  %s\n", ok ? "ok  " : "FAIL", what);
  if (!ok) ++g_failed;
}

// mean angular distance after removing the gauge with view 0 (both solutions share the same fixed relation only up to a
// global rotation): compare R_k R_0^T
double MeanError(const Orientations& est, const Orientations& gt, int n) {
  double R0e[9], R0g[9];
  Exp(est.at(0), R0e); Exp(gt.at(0), R0g);
  double sum = 0;
  for (int k = 1; k < n; ++k) {
    double Re[9], Rg[9], A[9], B[9], D[9];
    Exp(est.at(k), Re); Exp(gt.at(k), Rg);
    MulT(Re, R0e, A); MulT(Rg, R0g, B); MulT(A, B, D);
    const Vector3d d = Log(D);
    sum += std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  }
  return sum / (n - 1);
}

}  // namespace

int main(int argc, char** argv) {
  const bool expect_no_device = argc > 1 && std::strcmp(argv[1], "--expect-no-device") == 0;
  const int n = 60, extra = 400;
  std::mt19937 rng(56);  // the seed of Theia's rotation tests
  std::uniform_real_distribution<double> uni(-1.0, 1.0);
  std::normal_distribution<double> gauss(0.0, 1.0);
  Orientations gt, init, cam;  // cam: ground-truth camera positions (translation averaging)
  for (int k = 0; k < n; ++k) { Vector3d w; for (int t = 0; t < 3; ++t) w[t] = 0.2 * uni(rng) * 3.0; gt[k] = w; }
  for (int k = 0; k < n; ++k) { Vector3d c; for (int t = 0; t < 3; ++t) c[t] = 10.0 * uni(rng); cam[k] = c; }
  ViewPairs pairs;
  CovarianceMap covs;
  auto add = [&](ViewId a, ViewId b) {
    if (a == b) return;
    if (a > b) std::swap(a, b);
    if (pairs.count({a, b})) return;
    double Ra[9], Rb[9], Rab[9], N[9], Rn[9];
    Exp(gt[a], Ra); Exp(gt[b], Rb); MulT(Rb, Ra, Rab);
    Vector3d noise; for (int t = 0; t < 3; ++t) noise[t] = gauss(rng) * (1.0 * M_PI / 180.0) / std::sqrt(3.0);
    Exp(noise, N); Mul(N, Rab, Rn);
    TwoViewInfo info; info.rotation_2 = Log(Rn); info.num_verified_matches = 100;
    {  // position_2 = R_a (c_b - c_a) / |c_b - c_a| with half a degree of noise (T/sfm/twoview_info.h: position of camera 2 in frame 1)
      double d[3], nn = 0;
      for (int t = 0; t < 3; ++t) { d[t] = cam[b][t] - cam[a][t]; nn += d[t] * d[t]; }
      for (int t = 0; t < 3; ++t) d[t] = d[t] / std::sqrt(nn) + gauss(rng) * (0.5 * M_PI / 180.0);
      nn = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int r = 0; r < 3; ++r) info.position_2[r] = (Ra[3 * r] * d[0] + Ra[3 * r + 1] * d[1] + Ra[3 * r + 2] * d[2]) / nn;
    }
    pairs[{a, b}] = info;
    Matrix3d S; S.m[0] = S.m[4] = S.m[8] = 1e-8 * (0.5 + std::fabs(uni(rng)));  // isotropic-ish covariance, 1e8 Sigma ~ 1
    covs[{a, b}] = {S, Vector3d()};
  };
  for (int k = 0; k < n; ++k) add(k, (k + 1) % n);
  std::uniform_int_distribution<int> pick(0, n - 1);
  for (int k = 0; k < extra; ++k) add(pick(rng), pick(rng));
  // initial guess: chain the ring edges from view 0
  init[0] = gt[0];
  for (int k = 1; k < n; ++k) {
    double Rp[9], Rr[9], R[9];
    Exp(init[k - 1], Rp); Exp(pairs[{(ViewId)(k - 1), (ViewId)k}].rotation_2, Rr); Mul(Rr, Rp, R);
    init[k] = Log(R);
  }

  typedef gsfm_b200::GSfMNonlinearRotationEstimator<ViewPairs, Orientations> Estimator;
  Estimator est(0.1);
  gsfm_b200::RotationEstimator<ViewPairs, Orientations>* base = &est;  // the plugin interface

  {  // empty inputs: false, as rotation_estimator.cpp:29-40
    Orientations none; ViewPairs nopairs; Orientations some = init;
    Check(!base->EstimateRotations(pairs, &none), "EstimateRotations returns false without initial orientations");
    Check(!base->EstimateRotations(nopairs, &some), "EstimateRotations returns false without view pairs");
  }
  {  // covariance_rot.txt through the library's native writer / reader (host-only: runs without a device too)
    const char* tmp = std::getenv("TMPDIR");
    const std::string dir = std::string(tmp ? tmp : "/tmp") + "/gsfm_shim_" + std::to_string((long)getpid());
    Check(system(("mkdir -p " + dir).c_str()) == 0, "scratch directory");
    std::string err;
    Check(gsfm_b200::StoreCovarianceRot(dir, covs, &err), "StoreCovarianceRot");
    CovarianceMap back;
    Check(gsfm_b200::ReadCovariance(dir, &back, &err) && back.size() == covs.size(), "ReadCovariance");
    bool same = true;
    for (const auto& kv : covs) {
      const auto it = back.find(kv.first);
      if (it == back.end()) { same = false; break; }
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) same = same && std::memcmp(&kv.second.first.m[3 * r + c], &it->second.first.m[3 * (r < c ? r : c) + (r < c ? c : r)], 8) == 0;
    }
    Check(same, "  covariances survive the round trip bit for bit (upper triangle mirrored)");
    Check(!gsfm_b200::ReadCovariance(dir + "/nowhere", &back, &err) && err.find("cannot open") != std::string::npos, "  a missing file fails loudly");
    if (system(("rm -rf " + dir).c_str()) != 0) std::printf("      (could not remove %s)\n", dir.c_str());
  }
  typedef gsfm_b200::GSfMNonlinearPositionEstimator<ViewPairs, Orientations> PosEstimator;
  if (expect_no_device) {
    PosEstimator pe;
    Orientations c;
    Check(!pe.EstimatePositions(pairs, gt, &c) && pe.last_error().find("no CUDA device") != std::string::npos, "no CUDA device: EstimatePositions fails loudly");
  }
  if (expect_no_device) {
    Orientations o = init;
    Check(!base->EstimateRotations(pairs, &o), "no CUDA device: the call fails loudly (no CPU fallback)");
    Check(est.last_error().find("no CUDA device") != std::string::npos, "last_error names the missing device");
    Check(std::memcmp(&o.at(5), &init.at(5), sizeof(Vector3d)) == 0, "orientations untouched on failure");
    return g_failed ? 1 : 0;
  }

  const double e0 = MeanError(init, gt, n);
  {
    Orientations o = init;
    Check(base->EstimateRotations(pairs, &o), "EstimateRotations (SoftLOne 0.1, ANGLE_AXIS)");
    const double e = MeanError(o, gt, n);
    std::printf("      mean error vs ground truth: init %.4f deg -> %.4f deg, %d iterations, cost %.6g -> %.6g\n", e0 * 180 / M_PI, e * 180 / M_PI,
                est.summary().num_iterations, est.summary().initial_cost, est.summary().final_cost);
    Check(e < 0.5 * M_PI / 180.0 && e < e0, "  converges below 0.5 deg");
    Check(est.summary().final_cost < est.summary().initial_cost, "  cost decreases");
  }
  for (int type : {GSFM_RA_QUATERNION_COSINE, GSFM_RA_QUATERNION_NORM, GSFM_RA_ROTATION_MAT_FNORM}) {
    Orientations o = init;
    Check(est.EstimateRotationsWithCustomizedLoss(pairs, &o, gsfm_b200::HuberLoss(0.1), 4, type), "EstimateRotationsWithCustomizedLoss (quaternion type)");
    Check(MeanError(o, gt, n) < 0.5 * M_PI / 180.0, "  converges below 0.5 deg");
  }
  for (int type : {GSFM_RA_ANGLE_AXIS_COVARIANCE, GSFM_RA_ANGLE_AXIS, GSFM_RA_ANGLE_AXIS_INLIERS, GSFM_RA_ANGLE_AXIS_COV_INLIERS, GSFM_RA_ANGLE_AXIS_COVTRACE,
                   GSFM_RA_ANGLE_AXIS_COVNORM}) {
    Orientations o = init;
    auto matches = [&](const ViewIdPair& p) { return (double)pairs.at(p).num_verified_matches; };
    Check(est.EstimateRotationsWithCustomizedLossAndCovariance(pairs, &o, gsfm_b200::MAGSACWeightBasedLoss(2.0), 4, covs, type, matches),
          "EstimateRotationsWithCustomizedLossAndCovariance (angle-axis type)");
    Check(MeanError(o, gt, n) < 0.5 * M_PI / 180.0, "  converges below 0.5 deg");
  }
  {  // an edge without covariance is skipped, an orientation that is absent removes its edges (:239-247)
    CovarianceMap fewer = covs; fewer.erase(fewer.begin());
    Orientations o = init; o.erase(7);
    Check(est.EstimateRotationsWithCustomizedLossAndCovariance(pairs, &o, gsfm_b200::CauchyLoss(0.5), 1, fewer, GSFM_RA_ANGLE_AXIS_COVARIANCE),
          "skipping rules: missing covariance / missing orientation");
    Check(o.size() == (size_t)n - 1 && o.count(7) == 0, "  absent view stays absent");
  }
  {
    Orientations o = init;
    Check(est.EstimateRotationsWithSigmaConsensus(pairs, &o, gsfm_b200::TrivialLoss(), 1, 5, 0.05), "EstimateRotationsWithSigmaConsensus");
    Check(MeanError(o, gt, n) < 0.5 * M_PI / 180.0 && est.summary().outer_iterations >= 1, "  converges below 0.5 deg");
  }
  {  // a borrowed ceres::LossFunction-shaped object (what the reference's estimator takes, and what the pybind11 trampoline
     // pyLossFunction is) through the tabulating adapter, against the closed-form loss of the same formula; and ComposedLoss
    struct CeresCauchy {  // ceres::CauchyLoss(a): rho = b log(1 + s / b)
      double b, c;
      explicit CeresCauchy(double a) : b(a * a), c(1.0 / (a * a)) {}
      void Evaluate(double s, double* rho) const {
        const double sum = 1.0 + s * c, inv = 1.0 / sum;
        rho[0] = b * std::log(sum); rho[1] = inv > 2.2250738585072014e-308 ? inv : 2.2250738585072014e-308; rho[2] = -c * (inv * inv);
      }
    } ceres_loss(0.1);
    gsfm_b200::TabulatedLoss tab(ceres_loss);
    Orientations o1 = init, o2 = init;
    Check(est.EstimateRotationsWithCustomizedLossAndCovariance(pairs, &o1, tab, 1, covs, GSFM_RA_ANGLE_AXIS), "TabulatedLoss adapter (any object with Evaluate(s, rho[3]))");
    Check(est.EstimateRotationsWithCustomizedLossAndCovariance(pairs, &o2, gsfm_b200::CauchyLoss(0.1), 1, covs, GSFM_RA_ANGLE_AXIS), "  closed-form CauchyLoss");
    const double d = MeanError(o1, o2, n);
    std::printf("      tabulated vs closed form: mean difference %.3e rad\n", d);
    Check(d < 1e-6, "  the tabulated object reproduces the closed form to 1e-6 rad");
    Orientations o3 = init;
    Check(est.EstimateRotationsWithCustomizedLossAndCovariance(pairs, &o3, gsfm_b200::ComposedLoss(gsfm_b200::CauchyLoss(0.5), gsfm_b200::HuberLoss(0.05)), 1, covs,
                                                               GSFM_RA_ANGLE_AXIS), "ComposedLoss(Cauchy, Huber)");
    Check(MeanError(o3, gt, n) < 0.5 * M_PI / 180.0, "  converges below 0.5 deg");
  }
  {  // the steps before the solve, on the device: initial view-graph filter and spanning-tree initialisation
    ViewPairs vp = pairs;
    vp[{(ViewId)0, (ViewId)1}].num_verified_matches = 5;        // below the threshold: removed (the ring keeps 0 and 1 connected)
    TwoViewInfo island; island.num_verified_matches = 100;       // a two-view island: not the largest component
    vp[{(ViewId)1000, (ViewId)1001}] = island;
    std::string err;
    Check(gsfm_b200::FilterInitialViewGraph(&vp, 30, &err), "FilterInitialViewGraph");
    Check(vp.count({0, 1}) == 0 && vp.count({1000, 1001}) == 0 && vp.size() == pairs.size() - 1, "  weak pair and island removed, the rest kept");
    Orientations o;
    Check(gsfm_b200::OrientationsFromMaximumSpanningTree(vp, &o, &err), "OrientationsFromMaximumSpanningTree");
    Check(o.size() == (size_t)n, "  every view of the component gets an orientation");
    // the relative rotations carry 1 degree of noise: a tree path of a few edges stays within a few degrees of the truth
    const double e = MeanError(o, gt, n);
    std::printf("      spanning-tree initialisation: mean error vs ground truth %.3f deg\n", e * 180 / M_PI);
    Check(e < 10.0 * M_PI / 180.0, "  initialisation is consistent with the measurements");
    Check(base->EstimateRotations(vp, &o) && MeanError(o, gt, n) < 0.5 * M_PI / 180.0, "  and the solve converges from it");
  }
  {  // translation averaging behind theia::PositionEstimator (include/gsfm_position_estimator.hpp)
    PosEstimator pe;
    gsfm_b200::PositionEstimator<ViewPairs, Orientations>* pbase = &pe;
    Orientations none, c;
    ViewPairs nopairs;
    Check(!pbase->EstimatePositions(nopairs, gt, &c) && !pbase->EstimatePositions(pairs, none, &c), "EstimatePositions returns false for empty inputs");
    Check(pbase->EstimatePositions(pairs, gt, &c), "EstimatePositions (Huber 0.1, BASELINE, every camera starts at the origin)");
    Check(c.size() == (size_t)n && pe.fixed_view() == 0, "  a position for every oriented view; the smallest id is the constant one");
    Check(c.at(0)[0] == 0.0 && c.at(0)[1] == 0.0 && c.at(0)[2] == 0.0, "  the constant view stays at the origin");
    // the gauge left is translation (fixed by view 0) and SCALE: compare after the least-squares scale about view 0
    auto mean_err = [&](const Orientations& est_c) {
      double num = 0, den = 0;
      for (int k = 1; k < n; ++k)
        for (int t = 0; t < 3; ++t) { const double g = cam[k][t] - cam[0][t]; num += g * est_c.at(k)[t]; den += est_c.at(k)[t] * est_c.at(k)[t]; }
      const double sc = num / den;
      double e = 0;
      for (int k = 1; k < n; ++k) {
        double d2 = 0;
        for (int t = 0; t < 3; ++t) { const double d = sc * est_c.at(k)[t] - (cam[k][t] - cam[0][t]); d2 += d * d; }
        e += std::sqrt(d2);
      }
      return e / (n - 1);
    };
    const double e = mean_err(c);
    std::printf("      mean position error vs ground truth (scene of +-10): %.4f, %d iterations, cost %.6g -> %.6g\n", e, pe.summary().num_iterations,
                pe.summary().initial_cost, pe.summary().final_cost);
    Check(e < 0.3 && pe.summary().final_cost < 0.01 * pe.summary().initial_cost, "  recovers the camera positions up to the similarity gauge");
    Orientations c2;
    Check(pe.EstimatePositions(pairs, gt, &c2, gsfm_b200::PositionErrorType::BASELINE, gsfm_b200::CauchyLoss(0.1)) && mean_err(c2) < 0.3,
          "EstimatePositions with a customized loss (Cauchy 0.1)");
    Orientations fewer = gt; fewer.erase(9);
    Orientations c3;
    Check(pe.EstimatePositions(pairs, fewer, &c3) && c3.size() == (size_t)n - 1 && c3.count(9) == 0, "  a view without orientation gets no position and its pairs are skipped");
    PosEstimator::Options po; po.min_num_points_per_view = 10;
    PosEstimator pt(po);
    Check(!pt.EstimatePositions(pairs, gt, &c3) && !pt.last_error().empty(), "  point-to-camera constraints are declined loudly");
  }
  std::printf("%s\n", g_failed ? "FAILED" : "ALL OK");
  return g_failed ? 1 : 0;
}
