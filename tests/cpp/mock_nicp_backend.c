/* mock_nicp_backend.c -- a TEST DOUBLE, not a CPU fallback.  It defines the dozen nicp_* entry points that the reference-side
 * binding (integration/pwn_b200/b200_pwn.h) calls and answers them with the oracle, so that the binding's own logic -- buffer
 * layouts, column-major conversions, device-mirror bookkeeping, the state it publishes into the reference's finder / lineariser
 * objects -- can be exercised end to end on a machine without a GPU.  It is compiled by ONE test
 * (tests/test_reference_pwn_core.py::test_drop_in_binding_logic_with_a_mock_backend) into a temporary directory and put in
 * front of the real library with LD_LIBRARY_PATH for that one subprocess.  It is never built by build(), never installed next
 * to the product, and nothing under g2o_frontend_b200/, include/ or integration/ knows it exists: the product has no CPU path
 * (tests/test_abi.py::test_fails_loudly_without_gpu). */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "nicp_b200.h"
#include "pwn_oracle.h"
struct nicp_cloud { int n, cap; float *points, *normals, *statsM, *eig, *curv, *oP, *oN; int *statsN; int has_stats; };
struct nicp_context { int lastRows, lastCols; int *interval; int *refIndex, *curIndex, *corr; float *refDepth, *curDepth; int P, ncorr; float H[36], b[6]; };
const char *nicp_last_error(void) { return "mock backend"; }
int nicp_create(int d, nicp_context **c) { *c = calloc(1, sizeof **c); return 0; }
void nicp_destroy(nicp_context *c) { free(c); }
int nicp_cloud_create(nicp_context *ctx, int cap, nicp_cloud **out) {
  nicp_cloud *c = calloc(1, sizeof *c); c->cap = cap;
  c->points = calloc(cap, 16); c->normals = calloc(cap, 16); c->statsM = calloc(cap, 64); c->eig = calloc(cap, 12); c->curv = calloc(cap, 4);
  c->oP = calloc(cap, 64); c->oN = calloc(cap, 64); c->statsN = calloc(cap, 4); *out = c; return 0; }
void nicp_cloud_destroy(nicp_cloud *c) { if (!c) return; free(c->points); free(c->normals); free(c->statsM); free(c->eig); free(c->curv); free(c->oP); free(c->oN); free(c->statsN); free(c); }
int nicp_cloud_size(const nicp_cloud *c) { return c->n; }
int nicp_depth_to_cloud(nicp_context *ctx, const float *depth, const nicp_projector *p, const nicp_stats_params *sp, const float so[16], int keep, nicp_cloud *c, int *index) {
  orc_stats_params q; memset(&q, 0, sizeof q);
  q.worldRadius = sp->world_radius; q.minImageRadius = sp->min_image_radius; q.maxImageRadius = sp->max_image_radius; q.minPoints = sp->min_points;
  q.curvatureThreshold = sp->curvature_threshold; q.omegaCurvatureThreshold = sp->omega_curvature_threshold;
  for (int i = 0; i < 3; i++) { q.flatOmegaP[i] = sp->flat_omega_p[i]; q.nonFlatOmegaP[i] = 1; q.flatOmegaN[i] = sp->flat_omega_n[i]; q.nonFlatOmegaN[i] = sp->nonflat_omega_n[i]; }
  int P = p->rows * p->cols; free(ctx->interval); ctx->interval = malloc(4 * P); ctx->lastRows = p->rows; ctx->lastCols = p->cols;
  int *idx = index ? index : malloc(4 * P); float *integ = malloc(40 * (size_t)P);
  c->n = orc_depth_to_cloud(depth, p->rows, p->cols, p->K, p->min_distance, p->max_distance, &q, so, c->points, c->normals, c->statsM, c->eig, c->statsN, c->curv, c->oP, c->oN, idx, ctx->interval, integ);
  free(integ); if (!index) free(idx); c->has_stats = keep; return 0; }
int nicp_last_interval_image(nicp_context *ctx, int *out) { memcpy(out, ctx->interval, 4 * ctx->lastRows * ctx->lastCols); return 0; }
static const int SYM[6] = {0, 4, 8, 5, 9, 10};
int nicp_cloud_download(nicp_context *ctx, const nicp_cloud *c, float *p, float *n, float *cv, float *op, float *on) {
  if (p) memcpy(p, c->points, 16 * c->n); if (n) memcpy(n, c->normals, 16 * c->n); if (cv) memcpy(cv, c->curv, 4 * c->n);
  for (int i = 0; i < c->n; i++) for (int k = 0; k < 6; k++) { if (op) op[6*i+k] = c->oP[16*i+SYM[k]]; if (on) on[6*i+k] = c->oN[16*i+SYM[k]]; }
  return 0; }
int nicp_cloud_download_stats(nicp_context *ctx, const nicp_cloud *c, float *s, float *e, int *n) {
  if (!c->has_stats) return 1; memcpy(s, c->statsM, 64 * c->n); memcpy(e, c->eig, 12 * c->n); memcpy(n, c->statsN, 4 * c->n); return 0; }
int nicp_cloud_upload(nicp_context *ctx, nicp_cloud *c, int n, const float *p, const float *nr, const float *cv, const float *op, const float *on) {
  c->n = n; memcpy(c->points, p, 16 * n); memcpy(c->normals, nr, 16 * n); memcpy(c->curv, cv, 4 * n); memset(c->oP, 0, 64 * n); memset(c->oN, 0, 64 * n);
  static const int RC[6][2] = {{0,0},{0,1},{0,2},{1,1},{1,2},{2,2}};
  for (int i = 0; i < n; i++) for (int k = 0; k < 6; k++) { int r = RC[k][0], cc = RC[k][1]; c->oP[16*i+4*cc+r] = c->oP[16*i+4*r+cc] = op[6*i+k]; c->oN[16*i+4*cc+r] = c->oN[16*i+4*r+cc] = on[6*i+k]; }
  return 0; }
int nicp_align(nicp_context *ctx, const nicp_cloud *r, const nicp_cloud *c, const nicp_projector *p, const nicp_align_params *a, const float ro[16], const float co[16], const float g[16], const nicp_prior *pr, int np, float thr, nicp_align_result *res) {
  orc_align_params q; memset(&q, 0, sizeof q); q.outerIterations = a->outer_iterations; q.innerIterations = a->inner_iterations; memcpy(q.K, p->K, 36); q.rows = p->rows; q.cols = p->cols; q.minD = p->min_distance; q.maxD = p->max_distance;
  memcpy(q.refSensorOffset, ro, 64); memcpy(q.curSensorOffset, co, 64); memcpy(q.initialGuess, g, 64);
  q.corr.inlierDistanceThreshold = a->inlier_distance_threshold; q.corr.inlierNormalAngularThreshold = a->inlier_normal_angular_threshold; q.corr.flatCurvatureThreshold = a->flat_curvature_threshold; q.corr.inlierCurvatureRatioThreshold = a->inlier_curvature_ratio_threshold;
  q.inlierMaxChi2 = a->inlier_max_chi2; q.robustKernel = a->robust_kernel; q.numThreads = 1;
  int P = p->rows * p->cols; ctx->P = P; free(ctx->refIndex); free(ctx->curIndex); free(ctx->corr); free(ctx->refDepth); free(ctx->curDepth);
  ctx->refIndex = malloc(4*P); ctx->curIndex = malloc(4*P); ctx->corr = malloc(8*P); ctx->refDepth = malloc(4*P); ctx->curDepth = malloc(4*P);
  orc_align_result o; orc_align(r->n, r->points, r->normals, r->curv, c->n, c->points, c->normals, c->curv, c->oP, c->oN, &q, &o, ctx->refIndex, ctx->refDepth, ctx->curIndex, ctx->curDepth, ctx->corr, 0);
  memset(res, 0, sizeof *res); memcpy(res->T, o.T, 64); memcpy(res->omega, o.omega, 144); res->error = o.error; res->inliers = o.inliers; res->num_correspondences = o.numCorrespondences;
  res->translational_eigen_ratio = o.translationalRatio; res->rotational_eigen_ratio = o.rotationalRatio; ctx->ncorr = o.numCorrespondences; memcpy(ctx->H, o.H, 144); memcpy(ctx->b, o.b, 24); return 0; }
int nicp_align_get_state(nicp_context *ctx, int *ri, float *rd, int *ci, float *cd, int *corr, float H[36], float b[6]) {
  int P = ctx->P; if (ri) memcpy(ri, ctx->refIndex, 4*P); if (rd) memcpy(rd, ctx->refDepth, 4*P); if (ci) memcpy(ci, ctx->curIndex, 4*P); if (cd) memcpy(cd, ctx->curDepth, 4*P);
  if (corr) memcpy(corr, ctx->corr, 8 * ctx->ncorr); if (H) memcpy(H, ctx->H, 144); if (b) memcpy(b, ctx->b, 24); return 0; }
