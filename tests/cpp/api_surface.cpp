// api_surface.cpp -- CPU-only check of the pwn:: class surface (include/pwn/pwn.h): every constructor default of
// SURVEY.md Appendix B, every setter/getter pair the reference's callers use, and the host-side members that never
// touch the device (per-point project / unProject / projectInterval, scale, MultiPointProjector::computeImageSize,
// v2t / t2v, Isometry3f algebra, .pwn save / load).  Nothing here creates a pwn::Context, so it runs without a GPU.
// Built and run by tests/test_host_cpp.py::test_api_surface_cpu.
#include <cstdio>
#include <sstream>

#include "pwn/pwn.h"

using namespace pwn;

static int g_fail = 0;
#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);        \
      g_fail++;                                                          \
    }                                                                    \
  } while (0)
static bool near(float a, float b, float tol) { return std::fabs(a - b) <= tol; }

int main() {
  // ---- constructor defaults (reference file:line in SURVEY.md Appendix B) ----
  {
    PinholePointProjector p;  // pointprojector.cpp:6-13, pinholepointprojector.cpp:5-13
    CHECK(p.minDistance() == 0.01f && p.maxDistance() == 6.0f);
    CHECK(p.imageRows() == 0 && p.imageCols() == 0);
    CHECK(p.baseline() == 0.075f && p.alpha() == 0.1f);
    CHECK(p.cameraMatrix()(0, 0) == 1.0f && p.cameraMatrix()(1, 1) == 1.0f && p.cameraMatrix()(0, 2) == 0.5f &&
          p.cameraMatrix()(1, 2) == 0.5f && p.cameraMatrix()(2, 2) == 1.0f);
    CHECK(p.transform().matrix() == Matrix4f::Identity());
    StatsCalculatorIntegralImage s;  // statscalculatorintegralimage.cpp:6-12
    CHECK(s.worldRadius() == 0.1f && s.minImageRadius() == 10 && s.maxImageRadius() == 30 && s.minPoints() == 50 &&
          s.curvatureThreshold() == 0.02f);
    PointInformationMatrixCalculator pi;  // informationmatrixcalculator.h:105-110
    NormalInformationMatrixCalculator ni;  // informationmatrixcalculator.h:140-145
    CHECK(pi.curvatureThreshold() == 0.02f && ni.curvatureThreshold() == 0.02f);
    CHECK(pi.flatInformationMatrix()(0, 0) == 1000.0f && pi.flatInformationMatrix()(1, 1) == 1.0f &&
          pi.flatInformationMatrix()(2, 2) == 1.0f && pi.flatInformationMatrix()(3, 3) == 0.0f);
    CHECK(ni.flatInformationMatrix()(0, 0) == 100.0f && ni.flatInformationMatrix()(1, 1) == 100.0f &&
          ni.flatInformationMatrix()(2, 2) == 100.0f);
    CHECK(ni.nonFlatInformationMatrix()(0, 0) == 1.0f && pi.nonFlatInformationMatrix()(2, 2) == 1.0f);
    CorrespondenceFinder c;  // correspondencefinder.cpp:9-18
    CHECK(c.inlierDistanceThreshold() == 0.5f && c.squaredThreshold() == 0.25f);
    CHECK(c.inlierNormalAngularThreshold() == cosf((float)M_PI / 6));
    CHECK(c.flatCurvatureThreshold() == 0.02f && c.inlierCurvatureRatioThreshold() == 1.3f);
    CHECK(c.numCorrespondences() == 0 && c.imageRows() == 0 && c.imageCols() == 0);
    Linearizer l;  // linearizer.cpp:9-15
    CHECK(l.inlierMaxChi2() == 9e3f && l.robustKernel() && l.aligner() == 0);
    Aligner a;  // aligner.cpp:13-32
    CHECK(a.outerIterations() == 10 && a.innerIterations() == 1 && a.minInliers() == 100);
    CHECK(a.translationalMinEigenRatio() == 50.0f && a.rotationalMinEigenRatio() == 50.0f);
    CHECK(a.projector() == 0 && a.linearizer() == 0 && a.correspondenceFinder() == 0);
    CHECK(a.referenceCloud() == 0 && a.currentCloud() == 0 && !a.debug());
    CHECK(a.initialGuess().matrix() == Matrix4f::Identity() && a.sensorOffset().matrix() == Matrix4f::Identity());
    Merger m;  // merger.cpp:5-13
    CHECK(m.distanceThreshold() == 0.1f && m.maxPointDepth() == 10.0f && m.normalThreshold() == cosf(10 * M_PI / 180.0f));
    CHECK(m.imageSize().x() == 0 && m.imageSize().y() == 0 && m.depthImageConverter() == 0);
    VoxelCalculator v;
    CHECK(v.resolution() == 0.01f);
    Stats st;  // stats.h:21-27: identity, no points; curvature() of the empty Stats evaluates to 0 / 1e-9 = 0
    CHECK(st.n() == 0 && st(0, 0) == 1.0f && st(3, 3) == 1.0f && st.curvature() == 0.0f);
    Point pt;
    Normal nr;
    CHECK(pt[3] == 1.0f && nr[3] == 0.0f);
    Correspondence co;
    CHECK(co.referenceIndex == -1 && co.currentIndex == -1);
  }

  // ---- setters / getters ----
  {
    Aligner a;
    Linearizer l;
    CorrespondenceFinder c;
    PinholePointProjector p;
    Cloud ref, cur;
    a.setProjector(&p);
    a.setLinearizer(&l);
    a.setCorrespondenceFinder(&c);
    CHECK(a.projector() == &p && a.linearizer() == &l && a.correspondenceFinder() == &c && l.aligner() == &a);
    a.setOuterIterations(7);
    a.setInnerIterations(2);
    a.setMinInliers(33);
    a.setTranslationalMinEigenRatio(5.0f);
    a.setRotationalMinEigenRatio(6.0f);
    a.setDebug(true);
    CHECK(a.outerIterations() == 7 && a.innerIterations() == 2 && a.minInliers() == 33 && a.debug());
    CHECK(a.translationalMinEigenRatio() == 5.0f && a.rotationalMinEigenRatio() == 6.0f);
    Isometry3f g;
    g.setTranslation(0.1f, 0.2f, 0.3f);
    g.matrix()(3, 0) = 9.0f;  // the setters rewrite the last row (aligner.h:115-130)
    a.setInitialGuess(g);
    CHECK(a.initialGuess().matrix()(3, 0) == 0.0f && a.initialGuess().matrix()(0, 3) == 0.1f);
    a.setSensorOffset(g);
    CHECK(a.referenceSensorOffset().matrix()(1, 3) == 0.2f && a.currentSensorOffset().matrix()(2, 3) == 0.3f);
    // setReferenceCloud / setCurrentCloud clear the priors (aligner.h:60-80)
    Matrix6f info = Matrix6f::Identity();
    a.addRelativePrior(g, info);
    a.addAbsolutePrior(g, g, info);
    a.setReferenceCloud(&ref);
    a.setCurrentCloud(&cur);
    CHECK(a.referenceCloud() == &ref && a.currentCloud() == &cur);
    nicp_align_params ap = a.abiAlignParams();
    CHECK(ap.outer_iterations == 7 && ap.inner_iterations == 2 && ap.robust_kernel == 1 && ap.inlier_max_chi2 == 9e3f);
    c.setInlierDistanceThreshold(1.0f);
    c.setInlierNormalAngularThreshold(0.95f);
    c.setFlatCurvatureThreshold(0.03f);
    c.setInlierCurvatureRatioThreshold(1.5f);
    c.setImageSize(12, 16);
    CHECK(c.squaredThreshold() == 1.0f && c.inlierNormalAngularThreshold() == 0.95f && c.flatCurvatureThreshold() == 0.03f &&
          c.inlierCurvatureRatioThreshold() == 1.5f);
    CHECK(c.imageRows() == 12 && c.imageCols() == 16 && c.referenceIndexImage().rows == 12 && c.currentIndexImage().cols == 16);
    l.setInlierMaxChi2(1e3f);
    l.setRobustKernel(false);
    l.setT(g);
    CHECK(l.inlierMaxChi2() == 1e3f && !l.robustKernel() && l.T().matrix()(3, 0) == 0.0f);
    StatsCalculatorIntegralImage s;
    s.setWorldRadius(0.2f); s.setMinImageRadius(3); s.setMaxImageRadius(6); s.setMinPoints(10); s.setCurvatureThreshold(0.2f);
    PointInformationMatrixCalculator pi;
    NormalInformationMatrixCalculator ni;
    DepthImageConverterIntegralImage conv(&p, &s, &pi, &ni);
    CHECK(conv.projector() == &p && conv.statsCalculator() == &s && conv.pointInformationMatrixCalculator() == &pi &&
          conv.normalInformationMatrixCalculator() == &ni);
    nicp_stats_params sp = conv.abiStatsParams();
    CHECK(sp.world_radius == 0.2f && sp.min_image_radius == 3 && sp.max_image_radius == 6 && sp.min_points == 10 &&
          sp.curvature_threshold == 0.2f && sp.omega_curvature_threshold == 0.02f && sp.flat_omega_p[0] == 1000.0f &&
          sp.flat_omega_n[2] == 100.0f && sp.nonflat_omega_n[1] == 1.0f);
    Merger m;
    m.setImageSize(60, 80);
    m.setDepthImageConverter(&conv);
    m.setDistanceThreshold(0.2f); m.setNormalThreshold(0.9f); m.setMaxPointDepth(5.0f);
    CHECK(m.imageSize()[0] == 60 && m.imageSize()[1] == 80 && m.depthImageConverter() == &conv && m.distanceThreshold() == 0.2f &&
          m.normalThreshold() == 0.9f && m.maxPointDepth() == 5.0f);
    cur.traversabilityVector().push_back(1);
    CHECK(cur.traversabilityVector().size() == 1);
    cur.clear();
    CHECK(cur.traversabilityVector().empty() && cur.size() == 0);
  }

  // ---- PinholePointProjector, host side ----
  {
    PinholePointProjector p;
    Matrix3f K = Matrix3f::Identity();
    K(0, 0) = 525.0f; K(1, 1) = 525.0f; K(0, 2) = 319.5f; K(1, 2) = 239.5f;  // pwn_simple_aligner.cpp:225-229
    p.setCameraMatrix(K);
    p.setImageSize(480, 640);
    p.setMinDistance(0.5f);
    p.setMaxDistance(4.5f);
    // K * K^-1 = I
    Matrix3f I3 = K * p.inverseCameraMatrix();
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) CHECK(near(I3(r, c), r == c ? 1.0f : 0.0f, 1e-5f));
    // identity pose: KRt = [K 0], iKRt = [K^-1 0] (pinholepointprojector.cpp:17-31)
    CHECK(p.KRt()(0, 0) == 525.0f && p.KRt()(0, 2) == 319.5f && p.KRt()(3, 3) == 1.0f && p.KRt()(0, 3) == 0.0f);
    CHECK(near(p.iKRt()(0, 0), 1.0f / 525.0f, 1e-9f) && near(p.iKRt()(0, 2), -319.5f / 525.0f, 1e-6f));
    // unProject -> project returns the pixel and the depth (pinholepointprojector.h:224-251); x is the column
    Isometry3f T;
    T.setTranslation(0.05f, -0.02f, 0.1f);
    Vector6f v;
    v[0] = 0.05f; v[1] = -0.02f; v[2] = 0.1f; v[3] = 0.01f; v[4] = -0.02f; v[5] = 0.015f;
    p.setTransform(v2t(v));
    int bad = 0;
    for (int y = 0; y < 480; y += 37)
      for (int x = 0; x < 640; x += 41) {
        const float d = 0.6f + 0.003f * (x + y);
        Point q;
        CHECK(p.unProject(q, x, y, d));
        int x2 = -1, y2 = -1;
        float d2;
        CHECK(p.project(x2, y2, d2, q));
        if (x2 != x || y2 != y || !near(d2, d, 2e-5f)) bad++;
      }
    CHECK(bad == 0);
    Point q;
    int xi, yi;
    float di;
    CHECK(!p.unProject(q, 10, 10, 0.4f) && !p.unProject(q, 10, 10, 4.6f));  // outside [minDistance, maxDistance]
    p.setTransform(Isometry3f::Identity());
    CHECK(!p.project(xi, yi, di, Point(0.0f, 0.0f, 0.2f)) && !p.project(xi, yi, di, Point(0.0f, 0.0f, 5.0f)));
    CHECK(p.project(xi, yi, di, Point(0.0f, 0.0f, 1.0f)) && xi == 320 && yi == 240 && di == 1.0f);  // round half away from zero
    // _projectInterval (pinholepointprojector.h:264-274): int(max(fx, fy) * r / d), -1 out of range
    CHECK(p.projectInterval(0, 0, 1.0f, 0.1f) == 52 && p.projectInterval(0, 0, 3.5f, 0.1f) == 15 &&
          p.projectInterval(0, 0, 0.4f, 0.1f) == -1 && p.projectInterval(0, 0, 4.6f, 0.1f) == -1);
    // scale (pinholepointprojector.cpp:149-154): first two rows of K, truncated image size
    p.scale(0.25f);
    CHECK(p.cameraMatrix()(0, 0) == 131.25f && p.cameraMatrix()(1, 2) == 59.875f && p.cameraMatrix()(2, 2) == 1.0f);
    CHECK(p.imageRows() == 120 && p.imageCols() == 160);
    nicp_projector ap = p.abiProjector();
    CHECK(ap.rows == 120 && ap.cols == 160 && ap.K[0] == 131.25f && ap.K[6] == 79.875f && ap.min_distance == 0.5f);
  }

  // ---- MultiPointProjector, host side (multipointprojector.cpp:7-18, .h:14-27) ----
  {
    MultiPointProjector mp;
    PinholePointProjector cam[3];
    Isometry3f off;
    mp.addPointProjector(&cam[0], off, 160, 120);
    mp.addPointProjector(&cam[1], off, 120, 100);
    int rows = 0, cols = 0;
    mp.computeImageSize(rows, cols);
    CHECK(rows == 160 && cols == 220 && mp.numProjectors() == 2);
    CHECK(cam[0].imageRows() == 160 && cam[0].imageCols() == 120);  // the child's setImageSize(width, height)
    mp.setPointProjector(&cam[2], off, 120, 100, 1);
    CHECK(cam[2].imageRows() == 120 && cam[2].imageCols() == 100);
    Isometry3f T;
    T.setTranslation(1.0f, 2.0f, 3.0f);
    mp.setTransform(T);
    CHECK(cam[0].transform().matrix()(0, 3) == 1.0f && cam[2].transform().matrix()(2, 3) == 3.0f);
    nicp_multi_projector abi = mp.abiMultiProjector();
    CHECK(abi.num_cameras == 2 && abi.camera[0].rows == 160 && abi.camera[0].cols == 120 && abi.camera[1].cols == 100);
    mp.clearProjectors();
    CHECK(mp.numProjectors() == 0);
  }

  // ---- bm_se3.h through the library's host helpers; Isometry3f algebra ----
  {
    Vector6f v;
    v[0] = 0.3f; v[1] = -0.2f; v[2] = 0.5f; v[3] = 0.05f; v[4] = -0.1f; v[5] = 0.08f;
    Isometry3f T = v2t(v);
    Vector6f w = t2v(T);
    for (int i = 0; i < 6; i++) CHECK(near(w[i], v[i], 2e-7f));
    Matrix3f R = T.linear(), RtR = R.transpose() * R;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) CHECK(near(RtR(r, c), r == c ? 1.0f : 0.0f, 1e-6f));
    Isometry3f P = T * T.inverse();
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) CHECK(near(P.matrix()(r, c), r == c ? 1.0f : 0.0f, 1e-6f));
    v[3] = v[4] = v[5] = 0.0f;
    CHECK(v2t(v).linear() == Matrix3f::Identity());
  }

  // ---- Cloud::save / Cloud::load (cloud.cpp:25-133), host only ----
  for (int binary = 0; binary < 2; binary++) {
    Cloud c;
    const int n = 7;
    c.points().resize(n);
    c.normals().resize(n);
    c.stats().resize(n);
    c.pointInformationMatrix().resize(n);
    c.normalInformationMatrix().resize(n);
    for (int i = 0; i < n; i++) {
      c.points()[i] = Point(0.25f * i, -0.5f * i, 1.0f + 0.125f * i);
      c.normals()[i] = Normal(0.0f, 0.6f, -0.8f);
      for (int k = 0; k < 3; k++) c.stats()[i](k, 3) = c.points()[i][k];
      c.stats()[i]._n = 50 + i;
      c.stats()[i]._eigenValues(0) = 0.001f * i;
      c.stats()[i]._eigenValues(1) = 0.5f;
      c.stats()[i]._eigenValues(2) = 0.75f;
    }
    Isometry3f T;
    T.setTranslation(0.5f, 0.25f, -0.125f);
    std::stringstream ss;
    CHECK(c.save(ss, T, 1, binary != 0));
    if (!binary) CHECK(ss.str().compare(0, 18, "PWNCLOUD 7 0\n0.5 0") == 0);
    Cloud d;
    Isometry3f T2;
    CHECK(d.load(T2, ss));
    CHECK(d.size() == (size_t)n && T2.matrix() == T.matrix());
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < 4; k++) CHECK(d.points()[i][k] == c.points()[i][k] && d.normals()[i][k] == c.normals()[i][k]);
      for (int k = 0; k < 3; k++) CHECK(d.stats()[i](k, 3) == c.stats()[i](k, 3));
      if (binary) CHECK(d.stats()[i].n() == 50 + i && d.stats()[i].eigenValues()(2) == 0.75f);
    }
    // step = 2 keeps every other point (cloud.cpp:93-101)
    std::stringstream s2;
    CHECK(c.save(s2, T, 2, binary != 0));
    Cloud e;
    CHECK(e.load(T2, s2));
    CHECK(e.size() == 3 && e.points()[1][0] == c.points()[2][0]);
  }
  {
    Cloud c;
    Isometry3f T;
    std::stringstream ss("NOTACLOUD 3 0\n");
    CHECK(!c.load(T, ss));
  }

  if (g_fail) {
    std::printf("api_surface: %d check(s) failed\n", g_fail);
    return 1;
  }
  std::printf("api_surface ok\n");
  return 0;
}
