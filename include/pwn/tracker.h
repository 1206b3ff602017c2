// pwn/tracker.h -- sequential depth-frame tracking (BASELINE config 3) over the pwn:: classes.
//
// Mirrors the control flow of pwn_tracker::PwnTracker::processFrame
// (g2o_frontend/pwn_tracker/pwn_tracker.cpp:106-282) and PwnMatcherBase::makeCloud
// (g2o_frontend/pwn_tracker2/pwn_matcher_base.cpp:46-75) without the BOSS map / cache / callback plumbing
// (out of scope): every frame is converted to a cloud and aligned against the current
// KEYFRAME cloud with the guess prevT^-1 * globalT * initialGuess; the keyframe is replaced when
// inliers / (rows*cols) drops below newFrameInliersFraction (0.4, pwn_tracker.cpp:33,164-167); the global
// rotation is re-orthogonalised every 50 frames (:153-158).
#pragma once
#include "pwn.h"

namespace pwn {

class SequentialTracker {
 public:
  SequentialTracker(DepthImageConverterIntegralImage *converter, Aligner *aligner)
      : _converter(converter), _aligner(aligner), _scale(2), _newFrameInliersFraction(0.4f), _previousCloud(0),
        _counter(0), _numKeyframes(0), _lastInliers(0), _lastWasKeyframe(false) {}
  ~SequentialTracker() { delete _previousCloud; }
  void setScale(int scale_) { _scale = scale_; }                                  // PwnMatcherBase::_scale
  void setNewFrameInliersFraction(float v) { _newFrameInliersFraction = v; }      // pwn_tracker.cpp:33
  const Isometry3f &globalT() const { return _globalT; }
  int numKeyframes() const { return _numKeyframes; }
  int lastInliers() const { return _lastInliers; }
  bool lastWasKeyframe() const { return _lastWasKeyframe; }
  const Isometry3f &lastRelative() const { return _lastRelative; }

  // PwnMatcherBase::makeCloud
  Cloud *makeCloud(int &r, int &c, Matrix3f &cameraMatrix, const Isometry3f &sensorOffset, const RawDepthImage &raw,
                   float depthScale) {
    PinholePointProjector *projector = dynamic_cast<PinholePointProjector *>(_converter->projector());
    float invScale = 1.0f / _scale;
    Matrix3f scaledCameraMatrix = cameraMatrix * invScale;
    scaledCameraMatrix(2, 2) = 1.0f;
    projector->setCameraMatrix(scaledCameraMatrix);
    projector->setImageSize(raw.rows / _scale, raw.cols / _scale);
    DepthImage scaledImage;
    DepthImage_convertAndScale(scaledImage, raw, _scale, depthScale);
    cameraMatrix = projector->cameraMatrix();
    r = projector->imageRows();
    c = projector->imageCols();
    Cloud *cloud = new Cloud;
    _converter->compute(*cloud, scaledImage, sensorOffset);
    return cloud;
  }

  // PwnTracker::processFrame
  void processFrame(const RawDepthImage &raw, const Isometry3f &sensorOffset, const Matrix3f &cameraMatrix_,
                    const Isometry3f &initialGuess = Isometry3f::Identity(), float depthScale = 0.001f) {
    int r, c;
    Matrix3f scaledCameraMatrix = cameraMatrix_;
    Cloud *currentCloud = makeCloud(r, c, scaledCameraMatrix, sensorOffset, raw, depthScale);
    _lastWasKeyframe = false;
    if (_previousCloud) {
      _aligner->setCurrentSensorOffset(sensorOffset);
      _aligner->setCurrentCloud(currentCloud);
      _aligner->setReferenceSensorOffset(_previousCloudOffset);
      _aligner->setReferenceCloud(_previousCloud);
      _aligner->correspondenceFinder()->setImageSize(r, c);
      PinholePointProjector *alprojector = dynamic_cast<PinholePointProjector *>(_aligner->projector());
      alprojector->setCameraMatrix(scaledCameraMatrix);
      alprojector->setImageSize(r, c);
      Isometry3f guess = _previousCloudTransform.inverse() * _globalT * initialGuess;
      _aligner->setInitialGuess(guess);
      _aligner->align();
      _lastInliers = _aligner->inliers();
      _lastRelative = _aligner->T();
      if (_aligner->inliers() > 0)
        _globalT = _previousCloudTransform * _aligner->T();
      else
        _globalT = _globalT * guess;
      if (!(_counter % 50)) {  // pwn_tracker.cpp:153-158: R -= 0.5 R (R^T R - I)
        Matrix3f R = _globalT.linear();
        Matrix3f E = R.transpose() * R;
        for (int i = 0; i < 3; i++) E(i, i) -= 1.0f;
        Matrix3f corr = R * E;
        _globalT.setLinear(R - corr * 0.5f);
      }
      _globalT.fixLastRow();
      int maxInliers = r * c;
      float inliersFraction = (float)_aligner->inliers() / (float)maxInliers;
      if (inliersFraction < _newFrameInliersFraction) {
        _numKeyframes++;
        delete _previousCloud;
        _previousCloud = currentCloud;
        _previousCloudTransform = _globalT;
        _lastWasKeyframe = true;
      } else {
        delete currentCloud;
      }
    } else {
      _previousCloud = currentCloud;
      _previousCloudTransform = _globalT;
      _previousCloudOffset = sensorOffset;
      _numKeyframes++;
      _lastWasKeyframe = true;
    }
    _counter++;
  }

 protected:
  DepthImageConverterIntegralImage *_converter;
  Aligner *_aligner;
  int _scale;
  float _newFrameInliersFraction;
  Cloud *_previousCloud;
  Isometry3f _previousCloudTransform, _previousCloudOffset, _globalT, _lastRelative;
  int _counter, _numKeyframes, _lastInliers;
  bool _lastWasKeyframe;
};

}  // namespace pwn
