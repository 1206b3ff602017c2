// integration/pwn_b200/b200_pwn.h -- the binding a g2o_frontend maintainer adds (option A).
//
// It is written against the REFERENCE'S OWN headers (g2o_frontend/pwn_core/*.h, Eigen, OpenCV) and the C-ABI of this
// library (include/nicp_b200.h) and subclasses the two virtuals the trackers call:
//     pwn::DepthImageConverter::compute()   depthimageconverter.h:47   -> B200DepthImageConverter
//     pwn::Aligner::align()                 aligner.h:308              -> B200Aligner
// exactly like the reference's earlier GPU attempt did (class CuAligner : public Aligner, pwn_cuda/cualigner.h:8-15).
// Everything else -- projector, stats calculator, information-matrix calculators, finder, lineariser objects, their
// setters, the BOSS wrappers in pwn_boss, the trackers -- stays the reference's code and keeps configuring these two
// objects the way it always did.
//
// Clouds: pwn::Cloud is a bundle of host std::vectors with no hooks, so the device mirror is kept on the side, keyed by
// the Cloud* the callers pass around.  compute() builds the cloud on the GPU and (by default) copies it back into the host
// vectors, so every reference caller that reads cloud.points() / normals() / stats() keeps working; align() uses the
// device mirror when the host cloud still is what compute() produced and re-uploads it otherwise (a cloud loaded from a
// file, merged, transformed or voxelised on the host).
//
// Compiled in this repository by oracle/build_ref_pwn_core.sh against the reference headers (with the Eigen / OpenCV
// stand-ins of oracle/shim, because the image has neither library) into oracle/_ref/drop_in_demo; with the real Eigen
// and OpenCV it compiles unchanged (only data(), rows, cols, operator() of those types are used).
#pragma once
#include <sys/time.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "g2o_frontend/pwn_core/aligner.h"
#include "g2o_frontend/pwn_core/depthimageconverterintegralimage.h"
#include "g2o_frontend/pwn_core/pinholepointprojector.h"
#include "g2o_frontend/pwn_core/statscalculatorintegralimage.h"
#include "nicp_b200.h"

namespace pwn {

class B200Context {
 public:
  explicit B200Context(int device = 0) : _ctx(0) { check(nicp_create(device, &_ctx), "nicp_create"); }
  ~B200Context() {
    for (std::map<const Cloud *, Mirror>::iterator it = _mirrors.begin(); it != _mirrors.end(); ++it) nicp_cloud_destroy(it->second.dev);
    nicp_destroy(_ctx);
  }
  nicp_context *handle() { return _ctx; }
  static void check(int rc, const char *what) {
    if (rc != NICP_OK) throw std::runtime_error(std::string(what) + ": " + nicp_last_error());
  }

  // device cloud able to hold `capacity` points for this host cloud (created / grown on demand)
  nicp_cloud *acquire(const Cloud *cloud, int capacity) {
    Mirror &m = _mirrors[cloud];
    if (!m.dev || m.capacity < capacity) {
      if (m.dev) nicp_cloud_destroy(m.dev);
      m.dev = 0;
      check(nicp_cloud_create(_ctx, capacity > 0 ? capacity : 1, &m.dev), "nicp_cloud_create");
      m.capacity = capacity > 0 ? capacity : 1;
    }
    return m.dev;
  }
  // remember what the host cloud looked like when the mirror was last in step with it
  void stamp(const Cloud *cloud) {
    Mirror &m = _mirrors[cloud];
    m.size = cloud->points().size();
    m.data = cloud->points().empty() ? 0 : (const void *)&cloud->points()[0];
    m.first = cloud->points().empty() ? Point() : cloud->points()[0];
    m.last = cloud->points().empty() ? Point() : cloud->points()[cloud->points().size() - 1];
  }
  // the mirror of a cloud, uploaded from the host vectors if there is none or the host cloud has changed since
  nicp_cloud *mirror(const Cloud *cloud) {
    std::map<const Cloud *, Mirror>::iterator it = _mirrors.find(cloud);
    const size_t n = cloud->points().size();
    if (it != _mirrors.end() && it->second.dev && it->second.size == n && n > 0 &&
        it->second.data == (const void *)&cloud->points()[0] && it->second.first == cloud->points()[0] &&
        it->second.last == cloud->points()[n - 1])
      return it->second.dev;
    nicp_cloud *dev = acquire(cloud, (int)n);
    std::vector<float> p(4 * n), nr(4 * n), cv(n), op(6 * n), on(6 * n);
    for (size_t i = 0; i < n; i++) {
      for (int k = 0; k < 4; k++) p[4 * i + k] = cloud->points()[i](k);
      if (i < cloud->normals().size())
        for (int k = 0; k < 4; k++) nr[4 * i + k] = cloud->normals()[i](k);
      if (i < cloud->stats().size()) cv[i] = cloud->stats()[i].curvature();
      if (i < cloud->pointInformationMatrix().size()) sym6(cloud->pointInformationMatrix()[i], &op[6 * i]);
      if (i < cloud->normalInformationMatrix().size()) sym6(cloud->normalInformationMatrix()[i], &on[6 * i]);
    }
    static const float one[4] = {0, 0, 0, 1};
    check(nicp_cloud_upload(_ctx, dev, (int)n, n ? &p[0] : one, n ? &nr[0] : 0, n ? &cv[0] : 0, n ? &op[0] : 0, n ? &on[0] : 0),
          "nicp_cloud_upload");
    stamp(cloud);
    return dev;
  }
  void forget(const Cloud *cloud) {
    std::map<const Cloud *, Mirror>::iterator it = _mirrors.find(cloud);
    if (it == _mirrors.end()) return;
    if (it->second.dev) nicp_cloud_destroy(it->second.dev);
    _mirrors.erase(it);
  }

 private:
  struct Mirror {
    nicp_cloud *dev;
    int capacity;
    size_t size;
    const void *data;
    Point first, last;
    Mirror() : dev(0), capacity(0), size(0), data(0) {}
  };
  static void sym6(const Eigen::Matrix4f &m, float *o) {
    o[0] = m(0, 0); o[1] = m(0, 1); o[2] = m(0, 2); o[3] = m(1, 1); o[4] = m(1, 2); o[5] = m(2, 2);
  }
  B200Context(const B200Context &);
  nicp_context *_ctx;
  std::map<const Cloud *, Mirror> _mirrors;
};

inline nicp_projector b200Projector(const PinholePointProjector &pp, int rows, int cols) {
  nicp_projector p;
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) p.K[3 * c + r] = pp.cameraMatrix()(r, c);  // column-major, like Eigen
  p.rows = rows;
  p.cols = cols;
  p.min_distance = pp.minDistance();
  p.max_distance = pp.maxDistance();
  return p;
}

// DepthImageConverterIntegralImage::compute (depthimageconverterintegralimage.cpp:15-55) on the GPU
class B200DepthImageConverter : public DepthImageConverterIntegralImage {
 public:
  B200DepthImageConverter(B200Context *context, PointProjector *projector_ = 0, StatsCalculator *statsCalculator_ = 0,
                          PointInformationMatrixCalculator *pointInformationMatrixCalculator_ = 0,
                          NormalInformationMatrixCalculator *normalInformationMatrixCalculator_ = 0)
      // DepthImageConverter is a VIRTUAL base of DepthImageConverterIntegralImage (depthimageconverterintegralimage.h:15):
      // the most derived class initialises it
      : DepthImageConverter(projector_, statsCalculator_, pointInformationMatrixCalculator_, normalInformationMatrixCalculator_),
        DepthImageConverterIntegralImage(projector_, statsCalculator_, pointInformationMatrixCalculator_,
                                         normalInformationMatrixCalculator_),
        _context(context), _hostMirror(true) {}
  // false: leave the host vectors of the cloud empty (a pure GPU pipeline; align() only needs the device mirror)
  void setHostMirror(bool v) { _hostMirror = v; }

  virtual void compute(Cloud &cloud, const DepthImage &depthImage, const Eigen::Isometry3f &sensorOffset = Eigen::Isometry3f::Identity()) {
    PinholePointProjector *pp = dynamic_cast<PinholePointProjector *>(_projector);
    StatsCalculatorIntegralImage *sc = dynamic_cast<StatsCalculatorIntegralImage *>(_statsCalculator);
    if (!pp || !sc || !_pointInformationMatrixCalculator || !_normalInformationMatrixCalculator)
      throw std::runtime_error("B200DepthImageConverter: needs a PinholePointProjector, a StatsCalculatorIntegralImage and both information matrix calculators");
    const int rows = depthImage.rows, cols = depthImage.cols;
    // what the reference does to its own objects on the way (depthimageconverterintegralimage.cpp:34-39)
    cloud.clear();
    _projector->setImageSize(rows, cols);
    _projector->setTransform(Eigen::Isometry3f::Identity());
    _indexImage.create(rows, cols);
    nicp_projector p = b200Projector(*pp, rows, cols);
    nicp_stats_params s;
    s.world_radius = sc->worldRadius();
    s.min_image_radius = sc->minImageRadius();
    s.max_image_radius = sc->maxImageRadius();
    s.min_points = sc->minPoints();
    s.curvature_threshold = sc->curvatureThreshold();
    s.omega_curvature_threshold = _pointInformationMatrixCalculator->curvatureThreshold();
    InformationMatrix fp = _pointInformationMatrixCalculator->flatInformationMatrix();
    InformationMatrix fn = _normalInformationMatrixCalculator->flatInformationMatrix();
    InformationMatrix nn = _normalInformationMatrixCalculator->nonFlatInformationMatrix();
    for (int i = 0; i < 3; i++) { s.flat_omega_p[i] = fp(i, i); s.flat_omega_n[i] = fn(i, i); s.nonflat_omega_n[i] = nn(i, i); }
    // cv::Mat_ rows are contiguous for images the callers create(); stage a copy otherwise
    std::vector<float> staged;
    const float *src = &depthImage(0, 0);
    if (rows > 1 && &depthImage(1, 0) != src + cols) {
      staged.resize((size_t)rows * cols);
      for (int r = 0; r < rows; r++)
        for (int c = 0; c < cols; c++) staged[(size_t)r * cols + c] = depthImage(r, c);
      src = &staged[0];
    }
    nicp_context *ctx = _context->handle();
    nicp_cloud *dev = _context->acquire(&cloud, rows * cols);
    std::vector<int> index((size_t)rows * cols);
    B200Context::check(nicp_depth_to_cloud(ctx, src, &p, &s, sensorOffset.matrix().data(), _hostMirror ? 1 : 0, dev, &index[0]),
                       "nicp_depth_to_cloud");
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) _indexImage(r, c) = index[(size_t)r * cols + c];
    sc->intervalImage().create(rows, cols);
    B200Context::check(nicp_last_interval_image(ctx, &index[0]), "nicp_last_interval_image");
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) sc->intervalImage()(r, c) = index[(size_t)r * cols + c];
    if (_hostMirror) {
      download(ctx, dev, cloud);
      // the reference's converter also fills cloud.gaussians() (unProject with the sensor model, pinholepointprojector.cpp:
      // 93-133, then transformInPlace(sensorOffset)); Merger::merge indexes them, so a CPU Merger must find them here
      B200Context::check(nicp_cloud_compute_gaussians(ctx, dev, src, &p, pp->baseline(), pp->alpha(), sensorOffset.matrix().data()),
                         "nicp_cloud_compute_gaussians");
      const int n = nicp_cloud_size(dev);
      std::vector<float> g((size_t)NICP_GAUSS_FLOATS * (n > 0 ? n : 1));
      std::vector<int> gf(n > 0 ? n : 1);
      if (n) B200Context::check(nicp_cloud_download_gaussians(ctx, dev, &g[0], &gf[0]), "nicp_cloud_download_gaussians");
      cloud.gaussians().resize(n);
      for (int i = 0; i < n; i++) {
        const float *q = &g[(size_t)NICP_GAUSS_FLOATS * i];
        Eigen::Vector3f mean(q[0], q[1], q[2]);
        Eigen::Matrix3f cov;
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 3; r++) cov(r, c) = q[3 + 3 * c + r];
        cloud.gaussians()[i] = Gaussian3f(mean, cov, false);
      }
    }
    _context->stamp(&cloud);
  }

 protected:
  static void download(nicp_context *ctx, nicp_cloud *dev, Cloud &cloud) {
    const int n = nicp_cloud_size(dev);
    std::vector<float> p(4 * (size_t)n), nr(4 * (size_t)n), cv(n), op(6 * (size_t)n), on(6 * (size_t)n), s16(16 * (size_t)n), ev(3 * (size_t)n);
    std::vector<int> cnt(n);
    if (n) {
      B200Context::check(nicp_cloud_download(ctx, dev, &p[0], &nr[0], &cv[0], &op[0], &on[0]), "nicp_cloud_download");
      B200Context::check(nicp_cloud_download_stats(ctx, dev, &s16[0], &ev[0], &cnt[0]), "nicp_cloud_download_stats");
    }
    cloud.points().resize(n);
    cloud.normals().resize(n);
    cloud.stats().resize(n);
    cloud.pointInformationMatrix().resize(n);
    cloud.normalInformationMatrix().resize(n);
    for (int i = 0; i < n; i++) {
      Stats &st = cloud.stats()[i];
      for (int k = 0; k < 4; k++) { cloud.points()[i](k) = p[4 * (size_t)i + k]; cloud.normals()[i](k) = nr[4 * (size_t)i + k]; }
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) st(r, c) = s16[16 * (size_t)i + 4 * c + r];
      st.setEigenValues(Eigen::Vector3f(ev[3 * (size_t)i], ev[3 * (size_t)i + 1], ev[3 * (size_t)i + 2]));
      st.setN(cnt[i]);
      st.setCurvature(cv[i]);
      unsym6(&op[6 * (size_t)i], cloud.pointInformationMatrix()[i]);
      unsym6(&on[6 * (size_t)i], cloud.normalInformationMatrix()[i]);
    }
  }
  static void unsym6(const float *o, InformationMatrix &m) {
    m.setZero();
    m(0, 0) = o[0]; m(0, 1) = m(1, 0) = o[1]; m(0, 2) = m(2, 0) = o[2];
    m(1, 1) = o[3]; m(1, 2) = m(2, 1) = o[4]; m(2, 2) = o[5];
  }
  B200Context *_context;
  bool _hostMirror;
};

// The finder's correspondence count and the lineariser's H / b / error / inliers are protected members without setters;
// align() has to leave them as the reference's align() does.  A derived class may form a pointer to a protected member of
// its base and apply it to any object of the base type.
struct B200FinderAccess : public CorrespondenceFinder {
  static void setNumCorrespondences(CorrespondenceFinder &f, int n) { f.*(&B200FinderAccess::_numCorrespondences) = n; }
};
struct B200LinearizerAccess : public Linearizer {
  static void set(Linearizer &l, const float *H, const float *b, float error, int inliers) {
    Matrix6f &Hm = l.*(&B200LinearizerAccess::_H);
    Vector6f &bm = l.*(&B200LinearizerAccess::_b);
    for (int c = 0; c < 6; c++)
      for (int r = 0; r < 6; r++) Hm(r, c) = H[6 * c + r];
    for (int r = 0; r < 6; r++) bm(r) = b[r];
    l.*(&B200LinearizerAccess::_error) = error;
    l.*(&B200LinearizerAccess::_inliers) = inliers;
  }
};

// Aligner::align (aligner.cpp:49-150) on the GPU
class B200Aligner : public Aligner {
 public:
  explicit B200Aligner(B200Context *context) : Aligner(), _context(context), _frameInlierDepthThreshold(50.0f) {}
  // threshold of PwnMatcherBase::matchClouds' image statistics (pwn_matcher_base.cpp:15), returned in lastResult()
  void setFrameInlierDepthThreshold(float v) { _frameInlierDepthThreshold = v; }
  const nicp_align_result &lastResult() const { return _last; }

  virtual void align() {
    PinholePointProjector *pp = dynamic_cast<PinholePointProjector *>(_projector);
    if (!pp || !_linearizer || !_correspondenceFinder || !_referenceCloud || !_currentCloud)
      throw std::runtime_error("B200Aligner: needs a PinholePointProjector, a linearizer, a correspondence finder and both clouds");
    // Aligner::addRelativePrior / addAbsolutePrior (aligner.cpp:34-40) -> nicp_prior[]
    std::vector<nicp_prior> priors(_priors.size());
    for (size_t j = 0; j < _priors.size(); j++) {
      const SE3AbsolutePrior *ab = dynamic_cast<const SE3AbsolutePrior *>(_priors[j]);
      nicp_prior &q = priors[j];
      q.kind = ab ? 1 : 0;
      const Eigen::Isometry3f reference = ab ? ab->referenceTransform() : Eigen::Isometry3f::Identity();
      for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) {
          q.mean[4 * c + r] = _priors[j]->mean().matrix()(r, c);
          q.reference[4 * c + r] = reference.matrix()(r, c);
        }
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) q.information[6 * c + r] = _priors[j]->information()(r, c);
    }
    struct timeval tvStart, tvEnd;
    gettimeofday(&tvStart, 0);
    const int rows = pp->imageRows(), cols = pp->imageCols();
    nicp_projector p = b200Projector(*pp, rows, cols);
    nicp_align_params a;
    a.inlier_distance_threshold = _correspondenceFinder->inlierDistanceThreshold();
    a.inlier_normal_angular_threshold = _correspondenceFinder->inlierNormalAngularThreshold();
    a.flat_curvature_threshold = _correspondenceFinder->flatCurvatureThreshold();
    a.inlier_curvature_ratio_threshold = _correspondenceFinder->inlierCurvatureRatioThreshold();
    a.inlier_max_chi2 = _linearizer->inlierMaxChi2();
    a.robust_kernel = _linearizer->robustKernel() ? 1 : 0;
    a.outer_iterations = _outerIterations;
    a.inner_iterations = _innerIterations;
    nicp_context *ctx = _context->handle();
    nicp_cloud *ref = _context->mirror(_referenceCloud), *cur = _context->mirror(_currentCloud);
    B200Context::check(nicp_align(ctx, ref, cur, &p, &a, _referenceSensorOffset.matrix().data(), _currentSensorOffset.matrix().data(),
                                  _initialGuess.matrix().data(), priors.empty() ? 0 : &priors[0], (int)priors.size(),
                                  _frameInlierDepthThreshold, &_last),
                       "nicp_align");
    for (int c = 0; c < 4; c++)
      for (int r = 0; r < 4; r++) _T.matrix()(r, c) = _last.T[4 * c + r];
    for (int c = 0; c < 6; c++)
      for (int r = 0; r < 6; r++) _omega(r, c) = _last.omega[6 * c + r];
    _error = _last.error;
    _inliers = _last.inliers;
    _translationalEigenRatio = _last.translational_eigen_ratio;
    _rotationalEigenRatio = _last.rotational_eigen_ratio;
    // the state the callers read after align() (pwn_matcher_base.cpp:156-171, pwn_apps/pwn_cloud_aligner.cpp:683-715)
    CorrespondenceFinder *cf = _correspondenceFinder;
    cf->setImageSize(rows, cols);
    cf->referenceDepthImage().create(rows, cols);
    cf->currentDepthImage().create(rows, cols);
    const size_t P = (size_t)rows * cols;
    std::vector<int> ri(P), ci(P), corr(2 * P, -1);
    std::vector<float> rd(P), cd(P);
    float H[36], b[6];
    B200Context::check(nicp_align_get_state(ctx, &ri[0], &rd[0], &ci[0], &cd[0], &corr[0], H, b), "nicp_align_get_state");
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        const size_t i = (size_t)r * cols + c;
        cf->referenceIndexImage()(r, c) = ri[i];
        cf->currentIndexImage()(r, c) = ci[i];
        cf->referenceDepthImage()(r, c) = rd[i];
        cf->currentDepthImage()(r, c) = cd[i];
      }
    cf->correspondences().assign(P, Correspondence());
    for (int k = 0; k < _last.num_correspondences; k++) cf->correspondences()[k] = Correspondence(corr[2 * (size_t)k], corr[2 * (size_t)k + 1]);
    B200FinderAccess::setNumCorrespondences(*cf, _last.num_correspondences);
    // the lineariser: T = T^-1 and the H / b of _computeStatistics' linearisation at the final T (aligner.cpp:165-170);
    // error / inliers are those of the last loop iteration (the record has no separate pair for the final linearisation)
    _linearizer->setT(_T.inverse());
    B200LinearizerAccess::set(*_linearizer, H, b, _last.error, _last.inliers);
    _projector->setTransform(_T * _referenceSensorOffset);
    gettimeofday(&tvEnd, 0);
    _totalTime = (tvEnd.tv_sec - tvStart.tv_sec) * 1000.0 + (tvEnd.tv_usec - tvStart.tv_usec) * 0.001;
  }

 protected:
  B200Context *_context;
  float _frameInlierDepthThreshold;
  nicp_align_result _last;
};

}  // namespace pwn
