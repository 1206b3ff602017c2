// ref_pwn_cuda.cu -- TEST INFRASTRUCTURE (oracle side): a thin extern "C" face over the REFERENCE'S OWN CUDA
// implementation of the NICP iteration, g2o_frontend/pwn_cuda/cudaaligner_rk.cu (+ cudasla.cu, cudautils.cu,
// cudamatrix.cuh), compiled UNMODIFIED from where it lies under /root/reference by oracle/Makefile into
// oracle/_ref/libpwn_cuda_ref.so.  pwn_cuda is the reference's earlier GPU attempt (disabled in its CMake, written for
// Fermi/Kepler) but it depends on nothing except the CUDA runtime -- no Eigen, no OpenCV -- so, unlike pwn_core, it
// builds here.  What it pins:
//   * the reference's own float32 restatements of bm_se3.h (cudasla.cu:137-200: matBuildSkew, transformInverse,
//     _v2t, _t2v), which are __host__ __device__ and run on the CPU;
//   * AlignerContext::processCorrespondence (cudaaligner_rk.cu:558-678): the correspondence gates (normal angle,
//     distance, curvature ratio), the robust kernel and the per-correspondence Htt / Htr / Hrr / bt / br terms of
//     Linearizer::update.  It is a __device__ member; this file re-declares __device__ as host+device BEFORE including
//     the reference source so that the very same code also runs on the CPU (tests/test_reference_pwn_cuda.py);
//   * on a GPU, the whole reference iteration (z-buffer projection, fused gate + linearise kernel, block reduction)
//     through its flat API createContext / initComputation / simpleIteration (cudaaligner.h:59-80).
// Nothing under g2o_frontend_b200/ or include/ uses this file.  No reference source is copied into the repository.
#undef __device__
#define __device__ __location__(host) __location__(device)
#include "cudaaligner_rk.cu"  // -I /root/reference/g2o_frontend/pwn_cuda

#include <cstring>

extern "C" {

// ---- cudasla.cu helpers on the host ---------------------------------------------------------------------
void refcuda_v2t(const float v[6], float m[16]) { pwn::_v2t(m, v); }
void refcuda_t2v(const float m[16], float v[6]) { pwn::_t2v(v, m); }
void refcuda_transform_inverse(const float s[16], float d[16]) { pwn::transformInverse(d, s); }
void refcuda_skew(const float v[4], float m[16]) { pwn::matBuildSkew(m, v); }

// ---- AlignerContext::processCorrespondence on the host ----------------------------------------------------
// One reference point / current point.  params = distanceThreshold (squared), normalThreshold, flatCurvatureThreshold,
// minCurvatureRatio, maxCurvatureRatio, inlierThreshold (max chi2).  out56 = Htt(16) Htr(16) Hrr(16) bt(4) br(4) in the
// reference's layout (column-major 4x4; bt[3] = 1, br[3] = chi2 when accepted).  Returns what the member returns.
int refcuda_process_correspondence(const float T[16], const float refPoint[4], const float refNormal[4], float refCurvature,
                                   const float curPoint[4], const float curNormal[4], float curCurvature,
                                   const float omegaP[16], const float omegaN[16], const float params[6], int robustKernel,
                                   float out56[56], float *error) {
  pwn::AlignerContext c;
  std::memset(&c, 0, sizeof c);
  float rp[4], rn[4], cp[4], cn[4], oP[16], oN[16], rc = refCurvature, cc = curCurvature;
  std::memcpy(rp, refPoint, sizeof rp); std::memcpy(rn, refNormal, sizeof rn);
  std::memcpy(cp, curPoint, sizeof cp); std::memcpy(cn, curNormal, sizeof cn);
  std::memcpy(oP, omegaP, sizeof oP); std::memcpy(oN, omegaN, sizeof oN);
  c._referencePoints.map(4, 1, rp);
  c._referenceNormals.map(4, 1, rn);
  c._referenceCurvatures = &rc;
  c._currentPoints.map(4, 1, cp);
  c._currentNormals.map(4, 1, cn);
  c._currentCurvatures = &cc;
  c._currentOmegaPs.map(16, 1, oP);
  c._currentOmegaNs.map(16, 1, oN);
  c._distanceThreshold = params[0];
  c._normalThreshold = params[1];
  c._flatCurvatureThreshold = params[2];
  c._minCurvatureRatio = params[3];
  c._maxCurvatureRatio = params[4];
  c._inlierThreshold = params[5];
  c._robustKernel = robustKernel != 0;
  std::memcpy(c._transform, T, sizeof c._transform);
  float e = 0.0f;
  std::memset(out56, 0, 56 * sizeof(float));
  int r = c.processCorrespondence(&e, out56, out56 + 32, out56 + 16, out56 + 48, out56 + 52, 0, 0);
  if (error) *error = e;
  return r;
}

// ---- the reference's flat GPU API (needs a device) ----------------------------------------------------------
// status = the AlignerStatus operation code (0 = Ok)
int refcuda_create(void **ctx, int maxReferencePoints, int maxCurrentPoints, int rows, int cols) {
  pwn::AlignerContext *c = 0;
  pwn::AlignerStatus s = pwn::createContext(&c, maxReferencePoints, maxCurrentPoints, rows, cols);
  *ctx = c;
  return (int)s._operation;
}
int refcuda_destroy(void *ctx) { return (int)pwn::destroyContext((pwn::AlignerContext *)ctx)._operation; }
// params as above; the reference keeps them as public members of the context
void refcuda_set_params(void *ctx, const float params[6], int robustKernel) {
  pwn::AlignerContext *c = (pwn::AlignerContext *)ctx;
  c->_distanceThreshold = params[0];
  c->_normalThreshold = params[1];
  c->_flatCurvatureThreshold = params[2];
  c->_minCurvatureRatio = params[3];
  c->_maxCurvatureRatio = params[4];
  c->_inlierThreshold = params[5];
  c->_robustKernel = robustKernel != 0;
  // initComputation / simpleIteration upload _cudaHostContext, a copy made by init(): keep it in step
  if (c->_cudaHostContext) {
    c->_cudaHostContext->_distanceThreshold = params[0];
    c->_cudaHostContext->_normalThreshold = params[1];
    c->_cudaHostContext->_flatCurvatureThreshold = params[2];
    c->_cudaHostContext->_minCurvatureRatio = params[3];
    c->_cudaHostContext->_maxCurvatureRatio = params[4];
    c->_cudaHostContext->_inlierThreshold = params[5];
    c->_cudaHostContext->_robustKernel = robustKernel != 0;
  }
}
int refcuda_init_computation(void *ctx, const float K9[9], const float sensorOffset[16], float *refPoints4, float *refNormals4,
                             float *refCurvatures, int numRef, float *curPoints4, float *curNormals4, float *curCurvatures,
                             float *curOmegaP16, float *curOmegaN16, int numCur) {
  return (int)pwn::initComputation((pwn::AlignerContext *)ctx, K9, sensorOffset, refPoints4, refNormals4, refCurvatures, numRef,
                                   curPoints4, curNormals4, curCurvatures, curOmegaP16, curOmegaN16, numCur)._operation;
}
// one iteration at `transform` (applied to the reference points, like Linearizer::T()): projection of the reference
// cloud, fused gate + linearise kernel, reduction.  Hb56 = Htt Htr Hrr bt br as above (bt[3] = inliers, br[3] = chi2).
int refcuda_iteration(void *ctx, const float transform[16], float Hb56[56], int *inliers, float *error) {
  pwn::AlignerContext *c = (pwn::AlignerContext *)ctx;
  float T[16];
  std::memcpy(T, transform, sizeof T);
  int inl = 0;
  float err = 0.0f;
  int s = (int)pwn::simpleIteration(c, T, &inl, &err)._operation;
  if (Hb56) pwn::getHb(c, Hb56, Hb56 + 16, Hb56 + 32, Hb56 + 48, Hb56 + 52);
  if (inliers) *inliers = inl;
  if (error) *error = err;
  return s;
}
// the index images the last iteration worked on (device -> host), rows*cols ints each
int refcuda_get_indices(void *ctx, int *referenceIndices, int *currentIndices) {
  pwn::AlignerContext *c = (pwn::AlignerContext *)ctx;
  size_t bytes = sizeof(int) * (size_t)c->_rows * c->_cols;
  cudaError_t e = cudaSuccess;
  if (referenceIndices) e = cudaMemcpy(referenceIndices, c->_cudaHostContext->_referenceIndices.values(), bytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && currentIndices) e = cudaMemcpy(currentIndices, c->_cudaHostContext->_currentIndices.values(), bytes, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? 0 : 4;
}

}  // extern "C"
