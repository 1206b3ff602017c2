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
struct nicp_cloud { int n, cap; float *points, *normals, *statsM, *eig, *curv, *oP, *oN; int *statsN; int has_stats; float *gauss; int *gflags; int has_gauss; };
int orc_voxelize(const float *points, int n, float resolution, int strict, int *representatives); /* oracle/voxel_oracle.cpp */
struct nicp_context { int lastRows, lastCols; int *interval; int *refIndex, *curIndex, *corr; float *refDepth, *curDepth; int P, ncorr; float H[36], b[6]; };
const char *nicp_last_error(void) { return "mock backend"; }
int nicp_create(int d, nicp_context **c) { *c = calloc(1, sizeof **c); return 0; }
void nicp_destroy(nicp_context *c) { free(c); }
int nicp_cloud_create(nicp_context *ctx, int cap, nicp_cloud **out) {
  nicp_cloud *c = calloc(1, sizeof *c); c->cap = cap;
  c->points = calloc(cap, 16); c->normals = calloc(cap, 16); c->statsM = calloc(cap, 64); c->eig = calloc(cap, 12); c->curv = calloc(cap, 4);
  c->oP = calloc(cap, 64); c->oN = calloc(cap, 64); c->statsN = calloc(cap, 4); c->gauss = calloc(cap, 96); c->gflags = calloc(cap, 4); *out = c; return 0; }
void nicp_cloud_destroy(nicp_cloud *c) { if (!c) return; free(c->points); free(c->normals); free(c->statsM); free(c->eig); free(c->curv); free(c->oP); free(c->oN); free(c->statsN); free(c->gauss); free(c->gflags); free(c); }
int nicp_cloud_size(const nicp_cloud *c) { return c->n; }
int nicp_depth_to_cloud(nicp_context *ctx, const float *depth, const nicp_projector *p, const nicp_stats_params *sp, const float so[16], int keep, nicp_cloud *c, int *index) {
  orc_stats_params q; memset(&q, 0, sizeof q);
  q.worldRadius = sp->world_radius; q.minImageRadius = sp->min_image_radius; q.maxImageRadius = sp->max_image_radius; q.minPoints = sp->min_points;
  q.curvatureThreshold = sp->curvature_threshold; q.omegaCurvatureThreshold = sp->omega_curvature_threshold;
  for (int i = 0; i < 3; i++) { q.flatOmegaP[i] = sp->flat_omega_p[i]; q.nonFlatOmegaP[i] = 1; q.flatOmegaN[i] = sp->flat_omega_n[i]; q.nonFlatOmegaN[i] = sp->nonflat_omega_n[i]; }
  int P = p->rows * p->cols; free(ctx->interval); ctx->interval = malloc(4 * P); ctx->lastRows = p->rows; ctx->lastCols = p->cols;
  int *idx = index ? index : malloc(4 * P); float *integ = malloc(40 * (size_t)P);
  c->n = orc_depth_to_cloud(depth, p->rows, p->cols, p->K, p->min_distance, p->max_distance, &q, so, c->points, c->normals, c->statsM, c->eig, c->statsN, c->curv, c->oP, c->oN, idx, ctx->interval, integ);
  free(integ); if (!index) free(idx); c->has_stats = keep; c->has_gauss = 0; return 0; }
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
  orc_prior priors[8]; if (np > 8) return 1;
  for (int j = 0; j < np; j++) { priors[j].kind = pr[j].kind; memcpy(priors[j].mean, pr[j].mean, 64); orc_iso_inverse(pr[j].reference, priors[j].refInv); memcpy(priors[j].info, pr[j].information, 144); }
  q.numPriors = np; q.priors = np ? priors : 0;
  int P = p->rows * p->cols; ctx->P = P; free(ctx->refIndex); free(ctx->curIndex); free(ctx->corr); free(ctx->refDepth); free(ctx->curDepth);
  ctx->refIndex = malloc(4*P); ctx->curIndex = malloc(4*P); ctx->corr = malloc(8*P); ctx->refDepth = malloc(4*P); ctx->curDepth = malloc(4*P);
  orc_align_result o; orc_align(r->n, r->points, r->normals, r->curv, c->n, c->points, c->normals, c->curv, c->oP, c->oN, &q, &o, ctx->refIndex, ctx->refDepth, ctx->curIndex, ctx->curDepth, ctx->corr, 0);
  memset(res, 0, sizeof *res); memcpy(res->T, o.T, 64); memcpy(res->omega, o.omega, 144); res->error = o.error; res->inliers = o.inliers; res->num_correspondences = o.numCorrespondences;
  res->translational_eigen_ratio = o.translationalRatio; res->rotational_eigen_ratio = o.rotationalRatio; ctx->ncorr = o.numCorrespondences; memcpy(ctx->H, o.H, 144); memcpy(ctx->b, o.b, 24); return 0; }
int nicp_align_get_state(nicp_context *ctx, int *ri, float *rd, int *ci, float *cd, int *corr, float H[36], float b[6]) {
  int P = ctx->P; if (ri) memcpy(ri, ctx->refIndex, 4*P); if (rd) memcpy(rd, ctx->refDepth, 4*P); if (ci) memcpy(ci, ctx->curIndex, 4*P); if (cd) memcpy(cd, ctx->curDepth, 4*P);
  if (corr) memcpy(corr, ctx->corr, 8 * ctx->ncorr); if (H) memcpy(H, ctx->H, 144); if (b) memcpy(b, ctx->b, 24); return 0; }

/* ---- what the pwn:: host classes and the CLI driver (include/pwn/pwn.h, g2o_frontend_b200/host) call on top ---- */
int nicp_synchronize(nicp_context *c) { return 0; }
int nicp_is_verification_build(void) { return 1; }
long long nicp_launch_count(const nicp_context *c) { return 0; }
void nicp_update_matrices(const float K[9], const float T[16], float KRt[16], float iKRt[16]) {
  float a[16], b[16]; orc_update_matrices(K, T, a, b); if (KRt) memcpy(KRt, a, 64); if (iKRt) memcpy(iKRt, b, 64); }
void nicp_v2t(const float v[6], float T[16]) { orc_v2t(v, T); }
void nicp_t2v(const float T[16], float v[6]) { orc_t2v(T, v); }
int nicp_depth_prepare(nicp_context *ctx, const uint16_t *raw, int rows, int cols, float scale, int step, float maxCov, float *out) {
  if (step < 1) step = 1;
  float *tmp = malloc(4 * (size_t)rows * cols); orc_depth_u16_to_f32(raw, rows * cols, scale, tmp);
  if (step > 1) orc_depth_scale(tmp, rows, cols, step, maxCov, out); else memcpy(out, tmp, 4 * (size_t)rows * cols);
  free(tmp); return 0; }
int nicp_project(nicp_context *ctx, const nicp_cloud *c, const float KRt[16], int rows, int cols, float minD, float maxD, int *index, float *depth) {
  int *ii = index ? index : malloc(4 * (size_t)rows * cols); float *dd = depth ? depth : malloc(4 * (size_t)rows * cols);
  orc_project(c->points, c->n, rows, cols, KRt, minD, maxD, ii, dd); if (!index) free(ii); if (!depth) free(dd); return 0; }
int nicp_cloud_append(nicp_context *ctx, nicp_cloud *d, const nicp_cloud *s, const float T[16]) {
  if (d->n + s->n > d->cap) return 1;
  int k = d->n, n = s->n;
  memcpy(d->points + 4 * k, s->points, 16 * n); memcpy(d->normals + 4 * k, s->normals, 16 * n); memcpy(d->statsM + 16 * k, s->statsM, 64 * n);
  memcpy(d->oP + 16 * k, s->oP, 64 * n); memcpy(d->oN + 16 * k, s->oN, 64 * n); memcpy(d->eig + 3 * k, s->eig, 12 * n);
  memcpy(d->curv + k, s->curv, 4 * n); memcpy(d->statsN + k, s->statsN, 4 * n);
  orc_cloud_transform(T, n, d->points + 4 * k, d->normals + 4 * k, d->statsM + 16 * k, d->oP + 16 * k, d->oN + 16 * k);
  if (s->has_gauss) { memcpy(d->gauss + 24 * k, s->gauss, 96 * n); memcpy(d->gflags + k, s->gflags, 4 * n); orc_gaussians_transform(T, n, d->gauss + 24 * k, d->gflags + k); d->has_gauss = 1; }
  d->n += n; return 0; }
int nicp_cloud_compute_gaussians(nicp_context *ctx, nicp_cloud *c, const float *depth, const nicp_projector *p, float baseline, float alpha, const float so[16]) {
  float eye[16] = {1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1}, KRt[16], iKRt[16]; orc_update_matrices(p->K, eye, KRt, iKRt);
  size_t P = (size_t)p->rows * p->cols; float *pts = malloc(16 * P); int *idx = malloc(4 * P);
  int n = orc_unproject_gaussians(depth, p->rows, p->cols, p->K, iKRt, p->min_distance, p->max_distance, baseline, alpha, pts, idx, c->gauss, c->gflags);
  orc_gaussians_transform(so, n, c->gauss, c->gflags); free(pts); free(idx); c->has_gauss = 1; return n == c->n ? 0 : 1; }
int nicp_cloud_has_gaussians(const nicp_cloud *c) { return c && c->has_gauss; }
static void compact(nicp_cloud *c, const int *keepIdx, int k) {  /* eig / curv / statsN follow the kept points */
  for (int i = 0; i < k; i++) { int j = keepIdx[i]; memmove(c->eig + 3 * i, c->eig + 3 * j, 12); c->curv[i] = c->curv[j]; c->statsN[i] = c->statsN[j]; } }
int nicp_merge(nicp_context *ctx, nicp_cloud *c, const nicp_projector *p, const float T[16], const nicp_merge_params *m, int *collapsed, int *newSize) {
  int n = c->n; int *col = malloc(4 * (size_t)(n > 0 ? n : 1)), *keep = malloc(4 * (size_t)(n > 0 ? n : 1));
  int k = orc_merge(n, c->points, c->normals, c->statsM, c->oP, c->oN, c->gauss, c->gflags, p->rows, p->cols, p->K, T, p->min_distance, p->max_distance,
                    m->distance_threshold, m->normal_threshold, m->max_point_depth, col);
  int q = 0; for (int i = 0; i < n; i++) if (col[i] < 0 || col[i] == i) keep[q++] = i;
  compact(c, keep, k); if (collapsed) memcpy(collapsed, col, 4 * (size_t)n); c->n = k; if (newSize) *newSize = k; free(col); free(keep); return q == k ? 0 : 1; }
int nicp_voxelize(nicp_context *ctx, nicp_cloud *c, float res, int *reps, int *newSize) {
  int n = c->n; int *rep = malloc(4 * (size_t)(n > 0 ? n : 1)); int k = orc_voxelize(c->points, n, res, 1, rep);
  nicp_cloud *t; nicp_cloud_create(ctx, k > 0 ? k : 1, &t);
  for (int i = 0; i < k; i++) { int j = rep[i];
    memcpy(t->points + 4 * i, c->points + 4 * j, 16); memcpy(t->normals + 4 * i, c->normals + 4 * j, 16); memcpy(t->statsM + 16 * i, c->statsM + 16 * j, 64);
    memcpy(t->oP + 16 * i, c->oP + 16 * j, 64); memcpy(t->oN + 16 * i, c->oN + 16 * j, 64); memcpy(t->eig + 3 * i, c->eig + 3 * j, 12);
    t->curv[i] = c->curv[j]; t->statsN[i] = c->statsN[j]; memcpy(t->gauss + 24 * i, c->gauss + 24 * j, 96); t->gflags[i] = c->gflags[j]; }
  memcpy(c->points, t->points, 16 * (size_t)k); memcpy(c->normals, t->normals, 16 * (size_t)k); memcpy(c->statsM, t->statsM, 64 * (size_t)k);
  memcpy(c->oP, t->oP, 64 * (size_t)k); memcpy(c->oN, t->oN, 64 * (size_t)k); memcpy(c->eig, t->eig, 12 * (size_t)k); memcpy(c->curv, t->curv, 4 * (size_t)k);
  memcpy(c->statsN, t->statsN, 4 * (size_t)k); memcpy(c->gauss, t->gauss, 96 * (size_t)k); memcpy(c->gflags, t->gflags, 4 * (size_t)k);
  if (reps) memcpy(reps, rep, 4 * (size_t)k); c->n = k; if (newSize) *newSize = k; nicp_cloud_destroy(t); free(rep); return 0; }
/* entry points the mock does not model: loud failures */
int nicp_cloud_transform(nicp_context *c, nicp_cloud *d, const float T[16]) { return 1; }
int nicp_cloud_download_gaussians(nicp_context *c, const nicp_cloud *d, float *g, int *f) {
  if (!d->has_gauss) return 1; if (g) memcpy(g, d->gauss, 96 * (size_t)d->n); if (f) memcpy(f, d->gflags, 4 * (size_t)d->n); return 0; }
int nicp_cloud_upload_gaussians(nicp_context *c, nicp_cloud *d, const float *g, const int *f) { return 1; }
int nicp_unproject(nicp_context *c, const float *d, int r, int co, const float *m, float a, float b, nicp_cloud *cl, int *i) { return 1; }
int nicp_project_intervals(nicp_context *c, const float *d, const nicp_projector *p, float w, int *i) { return 1; }
int nicp_raw_depth_to_cloud(nicp_context *c, const uint16_t *r, int a, int b, float s, int st, float m, const nicp_projector *p, const nicp_stats_params *sp, const float *so, int k, nicp_cloud *cl, int *i) { return 1; }
int nicp_correspond_linearize(nicp_context *c, const nicp_cloud *r, const nicp_cloud *cu, const int *ri, const int *ci, int ro, int co, const float *T, const nicp_align_params *ap, float *H, float *b, float *e, int *i, int *n, int *im) { return 1; }
int nicp_linearize(nicp_context *c, const nicp_cloud *r, const nicp_cloud *cu, const int *co, int n, const float *T, const nicp_align_params *ap, float *H, float *b, float *e, int *i) { return 1; }
void nicp_multi_image_size(const nicp_multi_projector *m, int *r, int *c) { if (r) *r = 0; if (c) *c = 0; }
int nicp_multi_depth_to_cloud(nicp_context *c, const float *d, const nicp_multi_projector *m, const nicp_stats_params *s, const float *so, int k, nicp_cloud *cl, int *i) { return 1; }
int nicp_multi_project(nicp_context *c, const nicp_cloud *cl, const nicp_multi_projector *m, const float *T, int *i, float *d) { return 1; }
int nicp_multi_align(nicp_context *c, const nicp_cloud *r, const nicp_cloud *cu, const nicp_multi_projector *m, const nicp_align_params *a, const float *ro, const float *co, const float *g, const nicp_prior *p, int n, float t, nicp_align_result *res) { return 1; }
int nicp_align_get_trace(nicp_context *c, float *t, int n) { return 1; }
int nicp_align_batch(nicp_context *c, int n, const nicp_cloud *const *r, const nicp_cloud *const *cu, const nicp_projector *p, const nicp_align_params *a, const float *ro, const float *co, const float *g, float t, nicp_align_result *res) { return 1; }
/* stage-level virtuals (pwn.h references them from the vtables of the statistics / information-matrix calculators) */
static void mock_stats_params(const nicp_stats_params *sp, orc_stats_params *q) {
  memset(q, 0, sizeof *q);
  q->worldRadius = sp->world_radius; q->minImageRadius = sp->min_image_radius; q->maxImageRadius = sp->max_image_radius; q->minPoints = sp->min_points;
  q->curvatureThreshold = sp->curvature_threshold; q->omegaCurvatureThreshold = sp->omega_curvature_threshold;
  for (int i = 0; i < 3; i++) { q->flatOmegaP[i] = sp->flat_omega_p[i]; q->nonFlatOmegaP[i] = 1; q->flatOmegaN[i] = sp->flat_omega_n[i]; q->nonFlatOmegaN[i] = sp->nonflat_omega_n[i]; }
}
int nicp_stats_compute(nicp_context *ctx, const float *points4, int n, const int *index, const int *interval, int rows, int cols, const nicp_stats_params *sp, float *normals4, float *stats16, float *eig3, int *cnt, float *curv) {
  orc_stats_params q; mock_stats_params(sp, &q);
  float *integ = malloc(40 * (size_t)rows * cols);
  float *nr = normals4 ? normals4 : malloc(16 * (size_t)(n + 1)), *s = stats16 ? stats16 : malloc(64 * (size_t)(n + 1)), *e = eig3 ? eig3 : malloc(12 * (size_t)(n + 1)), *cv = curv ? curv : malloc(4 * (size_t)(n + 1));
  int *c = cnt ? cnt : malloc(4 * (size_t)(n + 1));
  orc_integral_image(index, points4, rows, cols, integ);
  orc_stats(integ, index, interval, points4, rows, cols, n, &q, nr, s, e, c, cv);
  free(integ); if (!normals4) free(nr); if (!stats16) free(s); if (!eig3) free(e); if (!curv) free(cv); if (!cnt) free(c);
  return 0; }
int nicp_information_compute(nicp_context *ctx, int n, const float *normals4, const float *stats16, const float *eig3, const float *curv, const nicp_stats_params *sp, float *op6, float *on6) {
  orc_stats_params q; mock_stats_params(sp, &q);
  float *oP = malloc(64 * (size_t)(n + 1)), *oN = malloc(64 * (size_t)(n + 1));
  orc_information(normals4, stats16, eig3, curv, n, &q, oP, oN);
  static const int at[6] = {0, 4, 8, 5, 9, 10}; /* (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) of a column-major 4x4 */
  for (int i = 0; i < n; i++) for (int k = 0; k < 6; k++) { if (op6) op6[6 * (size_t)i + k] = oP[16 * (size_t)i + at[k]]; if (on6) on6[6 * (size_t)i + k] = oN[16 * (size_t)i + at[k]]; }
  free(oP); free(oN); return 0; }
int nicp_raw_depth_to_cloud_batch(nicp_context *c, int n, const uint16_t *const *r, int a, int b, float s, int st, float m, const nicp_projector *p, const nicp_stats_params *sp, const float *so, int k, nicp_cloud *const *cl) { return 1; }
int nicp_align_batch_priors(nicp_context *c, int n, const nicp_cloud *const *r, const nicp_cloud *const *cu, const nicp_projector *p, const nicp_align_params *a, const float *ro, const float *co, const float *g, const nicp_prior *pr, const int *po, float t, nicp_align_result *res) { return 1; }
