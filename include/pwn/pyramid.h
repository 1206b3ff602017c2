// pwn/pyramid.h -- coarse-to-fine NICP (BASELINE config 2: 160x120 / 320x240 / 640x480).
//
// The reference has no pyramid class ("no pyramid exists in the reference"); it scales
// by integer down-sampling (DepthImage_scale, pwn_static.cpp:5-36; PinholePointProjector::scale,
// pinholepointprojector.cpp:149-154; PwnMatcherBase::makeCloud, pwn_tracker2/pwn_matcher_base.cpp:46-75).
// A pyramid is the composition of exactly those calls per level, with the previous level's T as the
// next level's initial guess -- a thin host loop over the existing device entry points.
#pragma once
#include "pwn.h"

namespace pwn {

struct PyramidLevel {
  int step;                       // DepthImage_scale step (4, 2, 1)
  int minImageRadius, maxImageRadius, minPoints;
  float inlierDistanceThreshold;
  int outerIterations;
};

class PyramidAligner {
 public:
  // converter / aligner carry every parameter that is not per level (their projector must be a
  // PinholePointProjector; its camera matrix is the full-resolution one)
  PyramidAligner(DepthImageConverterIntegralImage *converter, Aligner *aligner) : _converter(converter), _aligner(aligner) {}
  void addLevel(const PyramidLevel &l) { _levels.push_back(l); }
  const std::vector<Isometry3f> &levelTransforms() const { return _levelT; }
  const std::vector<int> &levelInliers() const { return _levelInliers; }

  Isometry3f align(const RawDepthImage &reference, const RawDepthImage &current, const Matrix3f &cameraMatrix,
                   const Isometry3f &sensorOffset, const Isometry3f &initialGuess, float depthScale = 0.001f) {
    PinholePointProjector *cp = dynamic_cast<PinholePointProjector *>(_converter->projector());
    PinholePointProjector *ap = dynamic_cast<PinholePointProjector *>(_aligner->projector());
    StatsCalculatorIntegralImage *sc = dynamic_cast<StatsCalculatorIntegralImage *>(_converter->statsCalculator());
    if (!cp || !ap || !sc) throw std::runtime_error("PyramidAligner: pinhole projectors / integral-image stats required");
    Isometry3f T = initialGuess;
    _levelT.clear();
    _levelInliers.clear();
    for (size_t li = 0; li < _levels.size(); li++) {
      const PyramidLevel &L = _levels[li];
      DepthImage dRef, dCur;
      DepthImage_convertAndScale(dRef, reference, L.step, depthScale);
      DepthImage_convertAndScale(dCur, current, L.step, depthScale);
      // PwnMatcherBase::makeCloud: scaled camera matrix and image size
      cp->setCameraMatrix(cameraMatrix);
      cp->setImageSize(reference.rows, reference.cols);
      cp->scale(1.0f / L.step);
      sc->setMinImageRadius(L.minImageRadius);
      sc->setMaxImageRadius(L.maxImageRadius);
      sc->setMinPoints(L.minPoints);
      Cloud ref, cur;
      _converter->compute(ref, dRef, sensorOffset);
      _converter->compute(cur, dCur, sensorOffset);
      ap->setCameraMatrix(cameraMatrix);
      ap->setImageSize(reference.rows, reference.cols);
      ap->scale(1.0f / L.step);
      _aligner->correspondenceFinder()->setImageSize(ap->imageRows(), ap->imageCols());
      _aligner->correspondenceFinder()->setInlierDistanceThreshold(L.inlierDistanceThreshold);
      _aligner->setOuterIterations(L.outerIterations);
      _aligner->setSensorOffset(sensorOffset);
      _aligner->setReferenceCloud(&ref);
      _aligner->setCurrentCloud(&cur);
      _aligner->setInitialGuess(T);
      _aligner->align();
      T = _aligner->T();
      _levelT.push_back(T);
      _levelInliers.push_back(_aligner->inliers());
    }
    return T;
  }

 protected:
  DepthImageConverterIntegralImage *_converter;
  Aligner *_aligner;
  std::vector<PyramidLevel> _levels;
  std::vector<Isometry3f> _levelT;
  std::vector<int> _levelInliers;
};

}  // namespace pwn
