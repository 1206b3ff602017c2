// solve_sections.cu -- where the single thread of k_reduce_solve's dense step spends its cycles (align.cu solve_step):
// the same device functions (nicp_math.cuh), one thread, clock64() between the sections.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../g2o_frontend_b200/csrc -I../../include -o solve_sections solve_sections.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "nicp_math.cuh"
using namespace nicp;

__global__ void k(const float *Hin, const float *bin, const float *invTin, const float *K, const float *off, float *out, long long *cyc) {
  float H[36], b[6], invT[16];
  long long t0 = clock64();
  for (int i = 0; i < 36; i++) H[i] = Hin[i];
  for (int i = 0; i < 6; i++) b[i] = bin[i];
  for (int i = 0; i < 16; i++) invT[i] = invTin[i];
  for (int i = 0; i < 36; i++) out[100 + i] = H[i];  // the stores of st->H / st->b
  for (int i = 0; i < 6; i++) out[140 + i] = b[i];
  long long t1 = clock64();
  for (int d = 0; d < 6; d++) NM6(H, d, d) = fadd(fadd(NM6(H, d, d), 1.0f), 1000.0f);
  float nb[6], dx[6], dT[16];
  for (int k = 0; k < 6; k++) nb[k] = -b[k];
  ldlt_solve6(H, nb, dx);
  out[0] = dx[0] + dx[1] + dx[2] + dx[3] + dx[4] + dx[5];
  long long t2 = clock64();
  v2t(dx, dT);
  iso_mul(dT, invT, invT);
  out[1] = invT[0] + invT[5] + invT[12];
  long long t3 = clock64();
  float T[16], v[6], tmp[16];
  iso_inverse(invT, T);
  t2v(T, v);
  v2t(v, T);
  fix_last_row(T);
  for (int k2 = 0; k2 < 16; k2++) out[200 + k2] = T[k2];
  long long t4 = clock64();
  iso_inverse(T, invT);
  fix_last_row(invT);
  for (int k2 = 0; k2 < 16; k2++) out[220 + k2] = invT[k2];
  iso_mul(T, off, tmp);
  float Tc[16], KRt[16];
  iso_mul(tmp, off, Tc);
  compute_KRt(K, Tc, KRt);
  for (int k2 = 0; k2 < 16; k2++) out[240 + k2] = KRt[k2];
  long long t5 = clock64();
  cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
}

int main() {
  float hH[36], hb[6], hT[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.01f, 0.02f, 0.03f, 1}, hK[9] = {525, 0, 0, 0, 525, 0, 319.5f, 239.5f, 1};
  float hoff[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) hH[c * 6 + r] = (r == c ? 5000.0f + 300 * r : 10.0f * (r + c));
  for (int i = 0; i < 6; i++) hb[i] = 3.0f * (i + 1);
  float *dH, *db, *dT, *dK, *doff, *dout;
  long long *dc;
  cudaMalloc(&dH, sizeof hH); cudaMalloc(&db, sizeof hb); cudaMalloc(&dT, sizeof hT); cudaMalloc(&dK, sizeof hK);
  cudaMalloc(&doff, sizeof hoff); cudaMalloc(&dout, 1024 * 4); cudaMalloc(&dc, 5 * 8);
  cudaMemcpy(dH, hH, sizeof hH, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof hb, cudaMemcpyHostToDevice);
  cudaMemcpy(dT, hT, sizeof hT, cudaMemcpyHostToDevice); cudaMemcpy(dK, hK, sizeof hK, cudaMemcpyHostToDevice);
  cudaMemcpy(doff, hoff, sizeof hoff, cudaMemcpyHostToDevice);
  long long c[5];
  for (int rep = 0; rep < 3; rep++) {
    k<<<1, 1>>>(dH, db, dT, dK, doff, dout, dc);
    cudaDeviceSynchronize();
    cudaMemcpy(c, dc, sizeof c, cudaMemcpyDeviceToHost);
    printf("rep %d cycles: load+store H/b %lld | ldlt %lld | v2t+mul %lld | inverse+t2v+v2t+store %lld | inverse+KRt+stores %lld | total %lld\n",
           rep, c[0], c[1], c[2], c[3], c[4], c[0] + c[1] + c[2] + c[3] + c[4]);
  }
  return 0;
}
