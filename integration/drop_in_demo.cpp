// integration/drop_in_demo.cpp -- the drop-in claim, executed: ONE program written against the reference's own classes
// (g2o_frontend/pwn_core, the flow of pwn_core/pwn_simple_aligner.cpp:28-188: projector / stats calculator /
// information-matrix calculators / converter / finder / lineariser / aligner, configured through the reference's setters)
// that runs the pipeline twice -- once with the reference's pwn::DepthImageConverterIntegralImage + pwn::Aligner (its CPU
// code, compiled from /root/reference), once with B200DepthImageConverter + B200Aligner (integration/pwn_b200/b200_pwn.h
// -> C-ABI -> CUDA) plugged into the SAME pointers -- and prints both results as one JSON object.
//
//   drop_in_demo depthA.f32 depthB.f32 rows cols fx fy cx cy [minImageRadius maxImageRadius minPoints inlierDistance [cpu|gpu|both [priors]]]
//
// depth files: rows*cols float32, metres.  Built by oracle/build_ref_pwn_core.sh into oracle/_ref/drop_in_demo (it contains
// reference code, so it lives with the other compiled reference artefacts); run by tests/test_vs_reference_gpu.py.
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "pwn_b200/b200_pwn.h"

using namespace pwn;

struct Result {
  Eigen::Isometry3f T;
  int nRef, nCur, inliers, numCorr;
  float error;
  double ms;
  long refIndexSum;
  size_t gaussians;
  double gaussianSum;  // sum over the reference cloud's gaussians of |mean| and |covariance| entries
};

static bool readDepth(const char *path, int rows, int cols, DepthImage &d) {
  std::ifstream is(path, std::ios::binary);
  if (!is) return false;
  std::vector<float> buf((size_t)rows * cols);
  is.read((char *)&buf[0], sizeof(float) * buf.size());
  if (!is) return false;
  d.create(rows, cols);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) d(r, c) = buf[(size_t)r * cols + c];
  return true;
}

// odometry / IMU style priors as pwn_tracker2 adds them (pwn_tracker.cpp:150-160): one relative, one absolute
static void addPriors(Aligner *aligner) {
  Vector6f v;
  v << 0.03f, -0.02f, 0.05f, 0.002f, 0.017f, 0.001f;
  Eigen::Isometry3f mean = v2t(v);
  Matrix6f info = Matrix6f::Identity() * 2000.0f;
  info(0, 1) = info(1, 0) = 150.0f;  // not diagonal, not symmetric-by-accident: a transposed copy would show
  info(3, 5) = info(5, 3) = -300.0f;
  aligner->addRelativePrior(mean, info);
  Vector6f w;
  w << 0.5f, 0.1f, -0.2f, 0.0f, 0.0871557f, 0.0f;
  Eigen::Isometry3f reference = v2t(w);
  aligner->addAbsolutePrior(reference, reference * mean, info * 0.5f);
}

static bool g_withPriors = false;

// everything below only sees the base-class pointers, like a tracker does
static Result run(DepthImageConverter *converter, Aligner *aligner, PinholePointProjector *projector, const DepthImage &dA,
                  const DepthImage &dB) {
  Cloud reference, current;
  converter->compute(reference, dA, Eigen::Isometry3f::Identity());
  converter->compute(current, dB, Eigen::Isometry3f::Identity());
  projector->setImageSize(dA.rows, dA.cols);
  aligner->correspondenceFinder()->setImageSize(dA.rows, dA.cols);
  aligner->setReferenceCloud(&reference);
  aligner->setCurrentCloud(&current);
  aligner->setInitialGuess(Eigen::Isometry3f::Identity());
  aligner->setSensorOffset(Eigen::Isometry3f::Identity());
  if (g_withPriors) addPriors(aligner);  // after setReferenceCloud / setCurrentCloud, which clear them (aligner.h:60-80)
  aligner->align();
  Result r;
  r.T = aligner->T();
  r.nRef = (int)reference.points().size();
  r.nCur = (int)current.points().size();
  r.inliers = aligner->inliers();
  r.error = aligner->error();
  r.numCorr = aligner->correspondenceFinder()->numCorrespondences();
  r.ms = aligner->totalTime();
  r.gaussians = reference.gaussians().size();
  r.gaussianSum = 0.0;
  for (size_t i = 0; i < reference.gaussians().size(); i++) {
    const Eigen::Vector3f m = reference.gaussians()[i].mean();
    const Eigen::Matrix3f C = reference.gaussians()[i].covarianceMatrix();
    for (int k = 0; k < 3; k++) r.gaussianSum += std::fabs((double)m(k));
    for (int k = 0; k < 9; k++) r.gaussianSum += std::fabs((double)C.data()[k]);
  }
  r.refIndexSum = 0;
  const IntImage &ri = aligner->correspondenceFinder()->referenceIndexImage();
  for (int y = 0; y < ri.rows; y++)
    for (int x = 0; x < ri.cols; x++) r.refIndexSum += ri(y, x) >= 0;
  return r;
}

static void print(const char *name, const Result &r) {
  std::printf("\"%s\": {\"T\": [", name);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) std::printf("%s%.9g", (i || j) ? ", " : "", r.T.matrix()(i, j));
  std::printf("], \"reference_points\": %d, \"current_points\": %d, \"inliers\": %d, \"num_correspondences\": %d, \"error\": %.9g, "
              "\"reference_pixels\": %ld, \"gaussians\": %zu, \"gaussian_sum\": %.17g, \"align_ms\": %.3f}",
              r.nRef, r.nCur, r.inliers, r.numCorr, r.error, r.refIndexSum, r.gaussians, r.gaussianSum, r.ms);
}

int main(int argc, char **argv) {
  if (argc < 9) {
    std::fprintf(stderr, "usage: %s depthA.f32 depthB.f32 rows cols fx fy cx cy [minR maxR minPoints inlierDistance [cpu|gpu|both]]\n", argv[0]);
    return 2;
  }
  const int rows = std::atoi(argv[3]), cols = std::atoi(argv[4]);
  DepthImage dA, dB;
  if (!readDepth(argv[1], rows, cols, dA) || !readDepth(argv[2], rows, cols, dB)) {
    std::fprintf(stderr, "cannot read the depth images\n");
    return 2;
  }
  const int minR = argc > 9 ? std::atoi(argv[9]) : 10, maxR = argc > 10 ? std::atoi(argv[10]) : 30, minPts = argc > 11 ? std::atoi(argv[11]) : 50;
  const float inlierDistance = argc > 12 ? (float)std::atof(argv[12]) : 1.0f;
  const std::string which = argc > 13 ? argv[13] : "both";
  const bool withPriors = argc > 14 && std::string(argv[14]) == "priors";
  omp_set_num_threads(1);  // the reference drops rows % threads rows and correspondences % threads terms; 1 = none
  g_withPriors = withPriors;

  // the reference's objects, configured as pwn_simple_aligner.cpp does from pwn_aligner_1_1.conf
  PinholePointProjector projector;
  Eigen::Matrix3f K;
  K << (float)std::atof(argv[5]), 0.0f, (float)std::atof(argv[7]), 0.0f, (float)std::atof(argv[6]), (float)std::atof(argv[8]), 0.0f, 0.0f, 1.0f;
  projector.setCameraMatrix(K);
  projector.setMinDistance(0.5f);
  projector.setMaxDistance(4.5f);
  StatsCalculatorIntegralImage statsCalculator;
  statsCalculator.setWorldRadius(0.1f);
  statsCalculator.setMinImageRadius(minR);
  statsCalculator.setMaxImageRadius(maxR);
  statsCalculator.setMinPoints(minPts);
  statsCalculator.setCurvatureThreshold(0.2f);
  PointInformationMatrixCalculator pointInformationMatrixCalculator;
  NormalInformationMatrixCalculator normalInformationMatrixCalculator;
  pointInformationMatrixCalculator.setCurvatureThreshold(0.02f);
  normalInformationMatrixCalculator.setCurvatureThreshold(0.02f);
  CorrespondenceFinder correspondenceFinder;
  correspondenceFinder.setInlierDistanceThreshold(inlierDistance);
  correspondenceFinder.setInlierNormalAngularThreshold(0.95f);
  correspondenceFinder.setFlatCurvatureThreshold(0.02f);
  correspondenceFinder.setInlierCurvatureRatioThreshold(1.3f);
  Linearizer linearizer;
  linearizer.setInlierMaxChi2(9e3f);
  linearizer.setRobustKernel(true);

  std::printf("{");
  bool first = true;
  if (which != "gpu") {
    DepthImageConverterIntegralImage converter(&projector, &statsCalculator, &pointInformationMatrixCalculator, &normalInformationMatrixCalculator);
    Aligner aligner;
    aligner.setProjector(&projector);
    aligner.setLinearizer(&linearizer);
    linearizer.setAligner(&aligner);
    aligner.setCorrespondenceFinder(&correspondenceFinder);
    aligner.setOuterIterations(10);
    aligner.setInnerIterations(1);
    print("reference_cpu", run(&converter, &aligner, &projector, dA, dB));
    first = false;
  }
  if (which != "cpu") {
    try {
      B200Context context(0);
      B200DepthImageConverter converter(&context, &projector, &statsCalculator, &pointInformationMatrixCalculator, &normalInformationMatrixCalculator);
      B200Aligner aligner(&context);
      aligner.setProjector(&projector);
      aligner.setLinearizer(&linearizer);
      linearizer.setAligner(&aligner);
      aligner.setCorrespondenceFinder(&correspondenceFinder);
      aligner.setOuterIterations(10);
      aligner.setInnerIterations(1);
      Result r = run(&converter, &aligner, &projector, dA, dB);
      if (!first) std::printf(", ");
      first = false;
      print("b200", r);
      // second alignment of the same pair: clouds resident, kernels warm
      Result r2 = run(&converter, &aligner, &projector, dA, dB);
      std::printf(", \"b200_second_align_ms\": %.3f", r2.ms);
    } catch (const std::exception &e) {
      std::printf("%s\"b200_error\": \"%s\"", first ? "" : ", ", e.what());
      std::printf("}\n");
      return 3;
    }
  }
  std::printf("}\n");
  return 0;
}
