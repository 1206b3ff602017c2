// ref_pwn_core.cpp -- TEST INFRASTRUCTURE (oracle side): an extern "C" face over the REFERENCE'S OWN pwn_core sources
// (g2o_frontend/pwn_core/*.cpp, compiled from /root/reference by oracle/build_ref_pwn_core.sh against the Eigen / OpenCV
// stand-ins of oracle/shim/) -> oracle/_ref/libpwn_core_ref.so.  tests/test_reference_pwn_core.py compares the oracle
// (oracle/pwn_oracle.c, the checker of every GPU parity test) with it stage by stage.  See oracle/shim/Eigen/Core for
// what this pins (the reference's control flow, indexing, gates, accumulation order, threading) and what it cannot
// (Eigen's numerical kernels, which the stand-in delegates to the oracle's restatements).
// Nothing under g2o_frontend_b200/ or include/ uses this file.
#include <omp.h>

#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <vector>

// the wrapper reads the raw state of Gaussian3f (which of its two forms is valid) and Merger::_collapsedIndices; the
// reference only exposes them through accessors that convert lazily / not at all
#define protected public
#include "g2o_frontend/pwn_core/aligner.h"
#include "g2o_frontend/pwn_core/cloud.h"
#include "g2o_frontend/pwn_core/correspondencefinder.h"
#include "g2o_frontend/pwn_core/depthimageconverterintegralimage.h"
#include "g2o_frontend/pwn_core/informationmatrixcalculator.h"
#include "g2o_frontend/pwn_core/linearizer.h"
#include "g2o_frontend/pwn_core/merger.h"
#include "g2o_frontend/pwn_core/multipointprojector.h"
#include "g2o_frontend/pwn_core/pinholepointprojector.h"
#include "g2o_frontend/pwn_core/pwn_static.h"
#include "g2o_frontend/pwn_core/statscalculatorintegralimage.h"
#include "g2o_frontend/pwn_core/voxelcalculator.h"

using namespace pwn;

namespace {
Eigen::Isometry3f iso(const float *T) {  // column-major 4x4
  Eigen::Isometry3f X;
  if (T) std::memcpy(X.matrix().data(), T, 16 * sizeof(float));
  return X;
}
Eigen::Matrix3f mat3(const float *K) {
  Eigen::Matrix3f M;
  std::memcpy(M.data(), K, 9 * sizeof(float));
  return M;
}
void setup_projector(PinholePointProjector &p, const float *K, int rows, int cols, float minD, float maxD) {
  p.setCameraMatrix(mat3(K));
  p.setImageSize(rows, cols);
  p.setMinDistance(minD);
  p.setMaxDistance(maxD);
}
struct FinderParams { float dist, ncos, flat, ratio; };
void setup_finder(CorrespondenceFinder &f, const float *p, int rows, int cols) {
  f.setInlierDistanceThreshold(p[0]);
  f.setInlierNormalAngularThreshold(p[1]);
  f.setFlatCurvatureThreshold(p[2]);
  f.setInlierCurvatureRatioThreshold(p[3]);
  f.setImageSize(rows, cols);
  f.referenceDepthImage().create(rows, cols);
  f.currentDepthImage().create(rows, cols);
}
}  // namespace

extern "C" {

void refcore_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

// pwn_static.cpp:54-68, :5-36
void refcore_depth_u16_to_f32(const unsigned short *raw, int rows, int cols, float scale, float *out) {
  cv::Mat src(rows, cols, CV_16UC1), dst;
  std::memcpy(src.data, raw, sizeof(unsigned short) * (size_t)rows * cols);
  DepthImage_convert_16UC1_to_32FC1(dst, src, scale);
  std::memcpy(out, dst.data, sizeof(float) * (size_t)rows * cols);
}
void refcore_depth_f32_to_u16(const float *depth, int rows, int cols, float scale, unsigned short *out) {
  cv::Mat src(rows, cols, CV_32FC1), dst;
  std::memcpy(src.data, depth, sizeof(float) * (size_t)rows * cols);
  DepthImage_convert_32FC1_to_16UC1(dst, src, scale);
  std::memcpy(out, dst.data, sizeof(unsigned short) * (size_t)rows * cols);
}
void refcore_depth_scale(const float *depth, int rows, int cols, int step, float maxDepthCov, float *out) {
  DepthImage src(rows, cols), dst;
  std::memcpy(src.data, depth, sizeof(float) * (size_t)rows * cols);
  DepthImage_scale(dst, src, step, maxDepthCov);
  std::memcpy(out, dst.data, sizeof(float) * (size_t)dst.rows * dst.cols);
}

// PinholePointProjector::_updateMatrices through setTransform + one unProject / project of probe points is indirect;
// the matrices themselves are protected, so the stage tests below exercise them through project / unProject.

// PinholePointProjector::_updateMatrices (pinholepointprojector.cpp:17-31): the two matrices are protected members
void refcore_update_matrices(const float K[9], const float T[16], float KRt[16], float iKRt[16]) {
  PinholePointProjector p;
  p.setCameraMatrix(mat3(K));
  p.setTransform(iso(T));
  std::memcpy(KRt, p._KRt.data(), sizeof(float) * 16);
  std::memcpy(iKRt, p._iKRt.data(), sizeof(float) * 16);
}
// v2t / t2v (bm_se3.h:36-52)
void refcore_v2t(const float v[6], float T[16]) {
  Vector6f x;
  std::memcpy(x.data(), v, sizeof(float) * 6);
  Eigen::Isometry3f X = v2t(x);
  std::memcpy(T, X.matrix().data(), sizeof(float) * 16);
}
void refcore_t2v(const float T[16], float v[6]) {
  Vector6f x = t2v(iso(T));
  std::memcpy(v, x.data(), sizeof(float) * 6);
}

// ---- clouds -------------------------------------------------------------------------------------------------
// DepthImageConverterIntegralImage::compute (depthimageconverterintegralimage.cpp:15-55) with the reference's own
// projector / stats calculator / information matrix calculators.  Returns a Cloud*.
// statsParams = worldRadius, minImageRadius, maxImageRadius, minPoints, curvatureThreshold, omegaCurvatureThreshold
void *refcore_depth_to_cloud(const float *depth, int rows, int cols, const float K[9], float minD, float maxD,
                             const float statsParams[6], const float sensorOffset[16], int *index, int *interval,
                             float *integral10) {
  PinholePointProjector projector;
  setup_projector(projector, K, rows, cols, minD, maxD);
  StatsCalculatorIntegralImage stats;
  stats.setWorldRadius(statsParams[0]);
  stats.setMinImageRadius((int)statsParams[1]);
  stats.setMaxImageRadius((int)statsParams[2]);
  stats.setMinPoints((int)statsParams[3]);
  stats.setCurvatureThreshold(statsParams[4]);
  PointInformationMatrixCalculator pim;
  NormalInformationMatrixCalculator nim;
  pim.setCurvatureThreshold(statsParams[5]);
  nim.setCurvatureThreshold(statsParams[5]);
  DepthImageConverterIntegralImage converter(&projector, &stats, &pim, &nim);
  DepthImage d(rows, cols);
  std::memcpy(d.data, depth, sizeof(float) * (size_t)rows * cols);
  Cloud *cloud = new Cloud();
  converter.compute(*cloud, d, iso(sensorOffset));
  if (index) std::memcpy(index, converter.indexImage().data, sizeof(int) * (size_t)rows * cols);
  if (interval) std::memcpy(interval, stats.intervalImage().data, sizeof(int) * (size_t)rows * cols);
  if (integral10) {
    // PointIntegralImage is stored transposed: coeff(image column, image row) (pointintegralimage.cpp:12)
    PointIntegralImage &I = stats.integralImage();
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        const PointAccumulator &a = I.coeffRef(c, r);
        float *o = integral10 + ((size_t)r * cols + c) * 10;
        o[0] = a.sum()(3);
        o[1] = a.sum()(0); o[2] = a.sum()(1); o[3] = a.sum()(2);
        o[4] = a.squaredSum()(0, 0); o[5] = a.squaredSum()(0, 1); o[6] = a.squaredSum()(0, 2);
        o[7] = a.squaredSum()(1, 1); o[8] = a.squaredSum()(1, 2); o[9] = a.squaredSum()(2, 2);
      }
  }
  return cloud;
}
void refcore_cloud_free(void *h) { delete (Cloud *)h; }
int refcore_cloud_size(void *h) { return (int)((Cloud *)h)->points().size(); }
// any output may be NULL; stats16 / omega16 column-major 4x4
void refcore_cloud_get(void *h, float *points4, float *normals4, float *stats16, float *eigvals3, int *statsN,
                       float *curvature, float *omegaP16, float *omegaN16) {
  Cloud &c = *(Cloud *)h;
  for (size_t i = 0; i < c.points().size(); i++) {
    if (points4) std::memcpy(points4 + 4 * i, c.points()[i].data(), 16);
    if (normals4) std::memcpy(normals4 + 4 * i, c.normals()[i].data(), 16);
    if (stats16) std::memcpy(stats16 + 16 * i, c.stats()[i].data(), 64);
    if (eigvals3) std::memcpy(eigvals3 + 3 * i, c.stats()[i].eigenValues().data(), 12);
    if (statsN) statsN[i] = c.stats()[i].n();
    if (curvature) curvature[i] = c.stats()[i].curvature();
    if (omegaP16) std::memcpy(omegaP16 + 16 * i, c.pointInformationMatrix()[i].data(), 64);
    if (omegaN16) std::memcpy(omegaN16 + 16 * i, c.normalInformationMatrix()[i].data(), 64);
  }
}
// Cloud::save / Cloud::load (cloud.cpp:25-133); sizes of the objects the binary mode dumps raw
int refcore_cloud_save(void *h, const char *path, const float T[16], int step, int binary) {
  return ((Cloud *)h)->save(path, iso(T), step, binary != 0) ? 1 : 0;
}
void *refcore_cloud_load(const char *path, float T[16]) {
  Cloud *c = new Cloud();
  Eigen::Isometry3f X;
  if (!c->load(X, path)) {
    delete c;
    return 0;
  }
  std::memcpy(T, X.matrix().data(), sizeof(float) * 16);
  return c;
}
void refcore_object_sizes(int sizes[3]) {
  sizes[0] = (int)sizeof(Point);
  sizes[1] = (int)sizeof(Normal);
  sizes[2] = (int)sizeof(Stats);
}
// Cloud::transformInPlace (cloud.cpp:173-186), Cloud::add (cloud.cpp:145-171)
void refcore_cloud_transform(void *h, const float T[16]) { ((Cloud *)h)->transformInPlace(iso(T)); }
void refcore_cloud_add(void *dst, void *src, const float T[16]) { ((Cloud *)dst)->add(*(Cloud *)src, iso(T)); }

// PinholePointProjector::project (pinholepointprojector.cpp:33-66) with the projector at pose T
void refcore_project(void *h, const float K[9], const float T[16], int rows, int cols, float minD, float maxD, int *index,
                     float *depth) {
  PinholePointProjector projector;
  setup_projector(projector, K, rows, cols, minD, maxD);
  projector.setTransform(iso(T));
  IntImage ii;
  DepthImage di;
  projector.project(ii, di, ((Cloud *)h)->points());
  std::memcpy(index, ii.data, sizeof(int) * (size_t)rows * cols);
  std::memcpy(depth, di.data, sizeof(float) * (size_t)rows * cols);
}
// PinholePointProjector::unProject (3-argument form, pinholepointprojector.cpp:68-91) at pose T; returns the count
int refcore_unproject(const float *depth, int rows, int cols, const float K[9], const float T[16], float minD, float maxD,
                      float *points4, int *index) {
  PinholePointProjector projector;
  setup_projector(projector, K, rows, cols, minD, maxD);
  projector.setTransform(iso(T));
  DepthImage d(rows, cols);
  std::memcpy(d.data, depth, sizeof(float) * (size_t)rows * cols);
  PointVector pts;
  IntImage ii;
  projector.unProject(pts, ii, d);
  for (size_t i = 0; i < pts.size(); i++) std::memcpy(points4 + 4 * i, pts[i].data(), 16);
  std::memcpy(index, ii.data, sizeof(int) * (size_t)rows * cols);
  return (int)pts.size();
}

// CorrespondenceFinder::compute (correspondencefinder.cpp:20-118) on given index images, then Linearizer::update
// (linearizer.cpp:17-115) at the same T.  finderParams = distance, normal cos, flat curvature, curvature ratio.
// corr receives numCorrespondences (referenceIndex, currentIndex) pairs.  H column-major 6x6.
int refcore_correspond_linearize(void *href, void *hcur, const int *refIndex, const int *curIndex, int rows, int cols,
                                 const float T[16], const float finderParams[4], float maxChi2, int robust, int *corr,
                                 float H[36], float b[6], float *error, int *inliers) {
  CorrespondenceFinder finder;
  setup_finder(finder, finderParams, rows, cols);
  std::memcpy(finder.referenceIndexImage().data, refIndex, sizeof(int) * (size_t)rows * cols);
  std::memcpy(finder.currentIndexImage().data, curIndex, sizeof(int) * (size_t)rows * cols);
  finder.compute(*(Cloud *)href, *(Cloud *)hcur, iso(T));
  const int n = finder.numCorrespondences();
  if (corr)
    for (int i = 0; i < n; i++) {
      corr[2 * i] = finder.correspondences()[i].referenceIndex;
      corr[2 * i + 1] = finder.correspondences()[i].currentIndex;
    }
  Linearizer linearizer;
  Aligner aligner;
  aligner.setReferenceCloud((Cloud *)href);
  aligner.setCurrentCloud((Cloud *)hcur);
  aligner.setCorrespondenceFinder(&finder);
  aligner.setLinearizer(&linearizer);
  linearizer.setAligner(&aligner);
  linearizer.setInlierMaxChi2(maxChi2);
  linearizer.setRobustKernel(robust != 0);
  linearizer.setT(iso(T));
  linearizer.update();
  Matrix6f Hm = linearizer.H();
  Vector6f bm = linearizer.b();
  if (H) std::memcpy(H, Hm.data(), sizeof(float) * 36);
  if (b) std::memcpy(b, bm.data(), sizeof(float) * 6);
  if (error) *error = linearizer.error();
  if (inliers) *inliers = linearizer.inliers();
  return n;
}

// Aligner::align (aligner.cpp:49-150).  priors: numPriors records of [kind (0 relative, 1 absolute), mean 16, reference 16,
// information 36] floats.  Outputs: T, omega (6x6), mean (6), error, inliers, the two eigen-ratios, and the finder's state
// after the last iteration (index / depth images, correspondences).
int refcore_align(void *href, void *hcur, const float K[9], int rows, int cols, float minD, float maxD,
                  const float finderParams[4], float maxChi2, int robust, int outer, int inner, const float guess[16],
                  const float refOffset[16], const float curOffset[16], const float *priors, int numPriors, float T[16],
                  float omega[36], float *error, int *inliers, float ratios[2], int *refIndex, float *refDepth, int *curIndex,
                  float *curDepth, int *corr) {
  PinholePointProjector projector;
  setup_projector(projector, K, rows, cols, minD, maxD);
  CorrespondenceFinder finder;
  setup_finder(finder, finderParams, rows, cols);
  Linearizer linearizer;
  Aligner aligner;
  aligner.setProjector(&projector);
  aligner.setReferenceCloud((Cloud *)href);
  aligner.setCurrentCloud((Cloud *)hcur);
  aligner.setCorrespondenceFinder(&finder);
  aligner.setLinearizer(&linearizer);
  linearizer.setAligner(&aligner);
  linearizer.setInlierMaxChi2(maxChi2);
  linearizer.setRobustKernel(robust != 0);
  aligner.setOuterIterations(outer);
  aligner.setInnerIterations(inner);
  aligner.setInitialGuess(iso(guess));
  aligner.setReferenceSensorOffset(iso(refOffset));
  aligner.setCurrentSensorOffset(iso(curOffset));
  for (int j = 0; j < numPriors; j++) {
    const float *p = priors + 69 * j;
    Matrix6f info;
    std::memcpy(info.data(), p + 33, sizeof(float) * 36);
    if (p[0] == 0.0f)
      aligner.addRelativePrior(iso(p + 1), info);
    else
      aligner.addAbsolutePrior(iso(p + 17), iso(p + 1), info);
  }
  aligner.align();
  std::memcpy(T, aligner.T().matrix().data(), sizeof(float) * 16);
  if (omega) std::memcpy(omega, aligner.omega().data(), sizeof(float) * 36);
  if (error) *error = aligner.error();
  if (inliers) *inliers = aligner.inliers();
  if (ratios) { ratios[0] = aligner.translationalEigenRatio(); ratios[1] = aligner.rotationalEigenRatio(); }
  const size_t P = (size_t)rows * cols;
  if (refIndex) std::memcpy(refIndex, finder.referenceIndexImage().data, sizeof(int) * P);
  if (refDepth) std::memcpy(refDepth, finder.referenceDepthImage().data, sizeof(float) * P);
  if (curIndex) std::memcpy(curIndex, finder.currentIndexImage().data, sizeof(int) * P);
  if (curDepth) std::memcpy(curDepth, finder.currentDepthImage().data, sizeof(float) * P);
  const int n = finder.numCorrespondences();
  if (corr)
    for (int i = 0; i < n; i++) {
      corr[2 * i] = finder.correspondences()[i].referenceIndex;
      corr[2 * i + 1] = finder.correspondences()[i].currentIndex;
    }
  return n;
}

// ---- MultiPointProjector (BASELINE config 5) ------------------------------------------------------------------
// What Aligner::align executes for a MultiPointProjector: the call through PointProjector* binds to the const virtual
// PointProjector::project (pointprojector.cpp:17-40: the base-class z-buffer, x is the row, empty depth 0, the images
// pre-sized by the caller), which calls the per-point MultiPointProjector::project (multipointprojector.cpp:157-205);
// MultiPointProjector's own image-level project is non-const and only hides it.  cams: per camera K (9), sensor offset
// (16), width, height, minD, maxD = 29 floats.
void refcore_multi_project(void *h, const float *cams, int numCams, const float T[16], int rows, int cols, int *index,
                           float *depth) {
  MultiPointProjector multi;
  for (int i = 0; i < numCams; i++) {
    const float *q = cams + 29 * i;
    PinholePointProjector *p = new PinholePointProjector();
    p->setCameraMatrix(mat3(q));
    p->setMinDistance(q[27]);
    p->setMaxDistance(q[28]);
    multi.addPointProjector(p, iso(q + 9), (int)q[25], (int)q[26]);
    p->setImageSize((int)q[25], (int)q[26]);  // what ChildProjectorInfo's constructor does (multipointprojector.h:61-78)
  }
  multi.setTransform(iso(T));
  const PointProjector *base = &multi;
  IntImage ii(rows, cols);
  DepthImage di;
  base->project(ii, di, ((Cloud *)h)->points());
  for (int r = 0; r < rows; r++) {
    std::memcpy(index + (size_t)r * cols, &ii(r, 0), sizeof(int) * cols);
    std::memcpy(depth + (size_t)r * cols, &di(r, 0), sizeof(float) * cols);
  }
  multi.clearProjectors();
}

// ---- local-map maintenance -----------------------------------------------------------------------------------
// Gaussian3f per point as the oracle lays it out: mean 3, covariance 9 (column-major), information vector 3,
// information matrix 9; flags bit 0 = _momentsUpdated, bit 1 = _infoUpdated (basemath/gaussian.h)
int refcore_cloud_gaussians(void *h, float *gauss24, int *flags) {
  Cloud &c = *(Cloud *)h;
  for (size_t i = 0; i < c.gaussians().size(); i++) {
    const Gaussian3f &g = c.gaussians()[i];
    float *o = gauss24 + 24 * i;
    std::memcpy(o, g._mean.data(), 12);
    std::memcpy(o + 3, g._covarianceMatrix.data(), 36);
    std::memcpy(o + 12, g._informationVector.data(), 12);
    std::memcpy(o + 15, g._informationMatrix.data(), 36);
    flags[i] = (g._momentsUpdated ? 1 : 0) | (g._infoUpdated ? 2 : 0);
  }
  return (int)c.gaussians().size();
}
// Merger::merge (merger.cpp:15-119) with a Merger of image size rows x cols whose converter holds a pinhole projector
int refcore_merge(void *h, const float K[9], const float T[16], int rows, int cols, float minD, float maxD,
                  float distanceThreshold, float normalThreshold, float maxPointDepth, int *collapsed) {
  Cloud &c = *(Cloud *)h;
  PinholePointProjector projector;
  setup_projector(projector, K, rows, cols, minD, maxD);
  DepthImageConverterIntegralImage converter(&projector, 0, 0, 0);
  Merger merger;
  merger.setDepthImageConverter(&converter);
  merger.setImageSize(rows, cols);
  merger.setDistanceThreshold(distanceThreshold);
  merger.setNormalThreshold(normalThreshold);
  merger.setMaxPointDepth(maxPointDepth);
  std::streambuf *old = std::cerr.rdbuf(0);  // the reference reports to stderr
  merger.merge(&c, iso(T));
  std::cerr.rdbuf(old);
  if (collapsed) std::memcpy(collapsed, merger._collapsedIndices.data(), sizeof(int) * merger._collapsedIndices.size());
  return (int)c.points().size();
}
// VoxelCalculator::compute (voxelcalculator.cpp:15-73)
int refcore_voxelize(void *h, float resolution) {
  Cloud &c = *(Cloud *)h;
  VoxelCalculator v;
  std::streambuf *old = std::cout.rdbuf(0);
  v.compute(c, resolution);
  std::cout.rdbuf(old);
  return (int)c.points().size();
}

}  // extern "C"
