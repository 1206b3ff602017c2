// corr_lin.cuh -- the fused CorrespondenceFinder::compute + Linearizer::update kernel, round-2 structure.
//
//   CorrespondenceFinder::compute   correspondencefinder.cpp:20-118
//   Linearizer::update              linearizer.cpp:17-115
//   PwnMatcherBase::matchClouds     pwn_tracker2/pwn_matcher_base.cpp:167-196 (image statistics, MODE 1)
//
// One warp owns a 96-pixel tile (3 pixels per lane) of a GROUP of pairs that share their current cloud -- the shape of
// loop-closure candidate verification, where one frame is matched against many candidates (pwn_closer.cpp:92-105).
// The current side of the tile (index image, point, normal + curvature, Omega_P / Omega_N) is loaded ONCE and kept in
// registers / shared memory while the warp walks the pairs of the group; per pair it only fetches the reference
// z-buffer words and the reference point / normal gathers, so the current cloud stops being re-read (and its addresses
// re-computed) for every pair.  A single alignment is a group of one and runs the same code: the partial row of a
// (pair, tile) does not depend on the group it was computed in.
//
// The Linearizer term runs its point half and its normal half in the two lanes of Blackwell's packed FP32 instructions
// (fma.rn.f32x2 / mul / add -> FFMA2 / FMUL2 / FADD2, sm_100): Omega_P / Omega_N are stored interleaved (Omega3) so a
// 128-bit load leaves (P_ij, N_ij) in an aligned register pair.  Measured on a B200 (tools/microbench/fp32_pipes.cu):
// FFMA2 issues at half the rate of FFMA for twice the work, i.e. the same FP32 rate for half the issue slots, and the
// kernel was bound by issue slots (profiles/r2_summary.md).  Every operation of the term is spelled out (no contraction
// left to the compiler), so both builds and every instantiation produce the same H and b bits.
#pragma once
#include "nicp_internal.cuh"

namespace nicp {

// ---- packed FP32 pairs ------------------------------------------------------------------------------------------
typedef unsigned long long f32x2;  // (lo, hi)
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(f32x2 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return a;
}
__device__ __forceinline__ float hi_of(f32x2 v) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  return b;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// accumulators of one (pair, tile): scalar sums for what only the point half feeds, packed (point, normal) sums for Hrr / br
struct TermAcc {
  float htt[6];    // sum Omega_P (xx xy xz yy yz zz)
  float htr[9];    // sum Omega_P * S_p, row-major
  f32x2 hrr[6];    // (S_p^T Omega_P S_p, S_n^T Omega_N S_n) upper triangle
  float bt[3];
  f32x2 br[3];
  float err, inl;
};
__device__ __forceinline__ void term_clear(TermAcc &A) {
#pragma unroll
  for (int i = 0; i < 6; i++) { A.htt[i] = 0.0f; A.hrr[i] = 0ull; }
#pragma unroll
  for (int i = 0; i < 9; i++) A.htr[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; i++) { A.bt[i] = 0.0f; A.br[i] = 0ull; }
  A.err = 0.0f;
  A.inl = 0.0f;
}

// one correspondence (linearizer.cpp:56-89).  R = transformed reference (point, normal) per axis, packed; cp / cn the
// current point / normal; w0..w2 the interleaved information matrices: A=(a,g) B=(b,h) C=(c,i) D=(d,j) E=(e,k) F=(f,l)
// with Omega_P = [a b c; b d e; c e f], Omega_N = [g h i; h j k; i k l].
__device__ __forceinline__ void term_add(TermAcc &acc, f32x2 Rx, f32x2 Ry, f32x2 Rz, float4 cp, float4 cn, f32x2 A, f32x2 B,
                                         f32x2 C, f32x2 D, f32x2 E, f32x2 F, float maxChi2, int robust) {
  // errors (rp - cp, rn - cn)
  const f32x2 E0 = pk(__fsub_rn(lo_of(Rx), cp.x), __fsub_rn(hi_of(Rx), cn.x));
  const f32x2 E1 = pk(__fsub_rn(lo_of(Ry), cp.y), __fsub_rn(hi_of(Ry), cn.y));
  const f32x2 E2 = pk(__fsub_rn(lo_of(Rz), cp.z), __fsub_rn(hi_of(Rz), cn.z));
  // Omega e, rows: (a e0 + b e1) + c e2 ...
  const f32x2 W0 = fma2(C, E2, fma2(B, E1, mul2(A, E0)));
  const f32x2 W1 = fma2(E, E2, fma2(D, E1, mul2(B, E0)));
  const f32x2 W2 = fma2(F, E2, fma2(E, E1, mul2(C, E0)));
  const f32x2 chi2 = fma2(E2, W2, fma2(E1, W1, mul2(E0, W0)));
  const float chi = __fadd_rn(lo_of(chi2), hi_of(chi2));
  float ks = 1.0f;
  if (chi > maxChi2) {
    if (!robust) return;
    ks = sqrtf(__fdividef(maxChi2, chi));
  }
  acc.inl = __fadd_rn(acc.inl, 1.0f);
  acc.err = __fmaf_rn(ks, chi, acc.err);
  // skew(v) = -2 [v]x (bm_se3.h:54-66): P = 2 v, NP = -2 v
  const f32x2 two = pk(2.0f, 2.0f), mtwo = pk(-2.0f, -2.0f);
  const f32x2 PX = mul2(Rx, two), PY = mul2(Ry, two), PZ = mul2(Rz, two);
  const f32x2 NX = mul2(Rx, mtwo), NY = mul2(Ry, mtwo), NZ = mul2(Rz, mtwo);
  // M = Omega S (point half: Omega_P S_p, normal half: Omega_N S_n)
  const f32x2 m00 = fma2(C, PY, mul2(B, NZ)), m01 = fma2(A, PZ, mul2(C, NX)), m02 = fma2(B, PX, mul2(A, NY));
  const f32x2 m10 = fma2(E, PY, mul2(D, NZ)), m11 = fma2(B, PZ, mul2(E, NX)), m12 = fma2(D, PX, mul2(B, NY));
  const f32x2 m20 = fma2(F, PY, mul2(E, NZ)), m21 = fma2(C, PZ, mul2(F, NX)), m22 = fma2(E, PX, mul2(C, NY));
  acc.htt[0] = __fadd_rn(acc.htt[0], lo_of(A)); acc.htt[1] = __fadd_rn(acc.htt[1], lo_of(B));
  acc.htt[2] = __fadd_rn(acc.htt[2], lo_of(C)); acc.htt[3] = __fadd_rn(acc.htt[3], lo_of(D));
  acc.htt[4] = __fadd_rn(acc.htt[4], lo_of(E)); acc.htt[5] = __fadd_rn(acc.htt[5], lo_of(F));
  acc.htr[0] = __fadd_rn(acc.htr[0], lo_of(m00)); acc.htr[1] = __fadd_rn(acc.htr[1], lo_of(m01));
  acc.htr[2] = __fadd_rn(acc.htr[2], lo_of(m02)); acc.htr[3] = __fadd_rn(acc.htr[3], lo_of(m10));
  acc.htr[4] = __fadd_rn(acc.htr[4], lo_of(m11)); acc.htr[5] = __fadd_rn(acc.htr[5], lo_of(m12));
  acc.htr[6] = __fadd_rn(acc.htr[6], lo_of(m20)); acc.htr[7] = __fadd_rn(acc.htr[7], lo_of(m21));
  acc.htr[8] = __fadd_rn(acc.htr[8], lo_of(m22));
  // Hrr = S^T M, upper triangle; S^T rows: (0,-tz,ty) (tz,0,-tx) (-ty,tx,0) with t = 2 v
  acc.hrr[0] = fma2(NZ, m10, fma2(PY, m20, acc.hrr[0]));
  acc.hrr[1] = fma2(NZ, m11, fma2(PY, m21, acc.hrr[1]));
  acc.hrr[2] = fma2(NZ, m12, fma2(PY, m22, acc.hrr[2]));
  acc.hrr[3] = fma2(NX, m21, fma2(PZ, m01, acc.hrr[3]));
  acc.hrr[4] = fma2(NX, m22, fma2(PZ, m02, acc.hrr[4]));
  acc.hrr[5] = fma2(NY, m02, fma2(PX, m12, acc.hrr[5]));
  acc.bt[0] = __fmaf_rn(ks, lo_of(W0), acc.bt[0]);
  acc.bt[1] = __fmaf_rn(ks, lo_of(W1), acc.bt[1]);
  acc.bt[2] = __fmaf_rn(ks, lo_of(W2), acc.bt[2]);
  const f32x2 KS = pk(ks, ks);
  acc.br[0] = fma2(KS, fma2(PY, W2, mul2(NZ, W1)), acc.br[0]);
  acc.br[1] = fma2(KS, fma2(PZ, W0, mul2(NX, W2)), acc.br[1]);
  acc.br[2] = fma2(KS, fma2(PX, W1, mul2(NY, W0)), acc.br[2]);
}

// the 32 reduction slots of a (pair, tile) from the accumulators (slot layout: A_* in nicp_internal.cuh)
__device__ __forceinline__ void term_slots(const TermAcc &A, float (&v)[kAccum]) {
#pragma unroll
  for (int i = 0; i < 6; i++) v[A_HTT + i] = A.htt[i];
#pragma unroll
  for (int i = 0; i < 9; i++) v[A_HTR + i] = A.htr[i];
#pragma unroll
  for (int i = 0; i < 6; i++) v[A_HRR + i] = __fadd_rn(lo_of(A.hrr[i]), hi_of(A.hrr[i]));
#pragma unroll
  for (int i = 0; i < 3; i++) v[A_BT + i] = A.bt[i];
#pragma unroll
  for (int i = 0; i < 3; i++) v[A_BR + i] = __fadd_rn(lo_of(A.br[i]), hi_of(A.br[i]));
  v[A_ERR] = A.err;
  v[A_INL] = A.inl;
}

// per-lane slots (lane l, pixel slot k -> [k * 32 + l]): nothing here is shared between lanes, shared memory is used as
// a register file extension for what must survive the walk over the pairs of the group
struct GroupSmem {
  float4 om[3][96];  // Omega_P / Omega_N of the current point (Omega3 layout), written by cp.async
  float4 cp[96];     // current point (w unused)
  float4 cn[96];     // current normal; w = curvature clamped to flatCurvatureThreshold, or -1 for a zero normal (MODE 0)
};

// MODE 0: correspondence gates + linearise at state->invT; the correspondence image is written only when asked for.
// MODE 1: linearise over the stored correspondence image (inner iterations > 0, _computeStatistics) and, if imgStats,
//         accumulate the matchClouds image statistics (slots 29..31).
template <int MODE, int MINB>
__global__ void __launch_bounds__(32, MINB) k_corr_lin_group(const PairDesc *__restrict__ desc, const PairGroup *__restrict__ groups,
                                                             int parity, int epoch, int writeCorr, AlignConsts ac, int numPixels,
                                                             int imgStats, float imgThreshold, int groupFast, int curEpoch) {
  constexpr int TK = 3, NT = 32, TILE = NT * TK;
  __shared__ GroupSmem S;
  // group-fastest block order: the CTAs resident together work on the same tile of different groups, so the reference
  // z-buffer rows and the tile's current-cloud lines they touch stay close in L2
  const int groupId = groupFast ? blockIdx.x : blockIdx.y;
  const int tileId = groupFast ? blockIdx.y : blockIdx.x;
  const PairGroup G = groups[groupId];
  const int lane = threadIdx.x;
  const int base = tileId * TILE;

  // ---- the current side of the tile, once per group ----
  const PairDesc &D0 = desc[G.first];
  const int *__restrict__ curIndex = D0.curIndex;
  const float4 *__restrict__ curPoints = D0.curPoints;
  const float4 *__restrict__ curNormals = D0.curNormals;
  const float4 *__restrict__ curOmega = D0.curOmega;
  // (read-only global loads spelled __ldg: the pointers come out of the descriptor, where the compiler cannot see their
  // address space and would emit generic loads)
  int ci[TK];
#pragma unroll
  for (int k = 0; k < TK; k++) {
    const int pix = base + k * NT + lane;
    ci[k] = pix < numPixels ? __ldg(curIndex + pix) : -1;
  }
  unsigned long long zc[TK];
  if (MODE == 1 && imgStats) {
    const unsigned long long *__restrict__ zcur = D0.curZ;
#pragma unroll
    for (int k = 0; k < TK; k++) {
      const int pix = base + k * NT + lane;
      zc[k] = pix < numPixels ? __ldg(zcur + pix) : kEmptyZ;
    }
  }
  bool curOk[TK];
  {
    float4 cpl[TK], cnl[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) {
      curOk[k] = ci[k] >= 0;
      cpl[k] = make_float4(0.f, 0.f, 0.f, 1.f);
      cnl[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (curOk[k]) {
        const float4 *om = curOmega + 3 * (size_t)ci[k];
        cp_async16(&S.om[0][k * NT + lane], om);
        cp_async16(&S.om[1][k * NT + lane], om + 1);
        cp_async16(&S.om[2][k * NT + lane], om + 2);
        cnl[k] = __ldg(curNormals + ci[k]);
        cpl[k] = __ldg(curPoints + ci[k]);
      }
    }
    // what the gates need from the current side (correspondencefinder.cpp:69, :87-93), prepared once per group
#pragma unroll
    for (int k = 0; k < TK; k++) {
      if (MODE == 0) {
        // a zero current normal rejects the pixel for every pair (the pixel still counts as "both indices valid")
        if (dot3(cnl[k].x, cnl[k].y, cnl[k].z, cnl[k].x, cnl[k].y, cnl[k].z) == 0.0f) cnl[k].w = -1.0f;  // curvature >= 0
        else if (cnl[k].w < ac.flatCurvature) cnl[k].w = ac.flatCurvature;
      }
      S.cp[k * NT + lane] = cpl[k];
      S.cn[k * NT + lane] = cnl[k];
    }
  }
  unsigned short c16[TK];
#pragma unroll
  for (int k = 0; k < TK; k++) {
    c16[k] = 0;
    if (MODE == 1 && imgStats) {
      const float dc = z_depth(zc[k], curEpoch, FLT_MAX);
      c16[k] = dc < FLT_MAX ? (unsigned short)(int)fmul(1000.0f, dc) : 0;
    }
  }
  bool omegaReady = false;

  // ---- the pairs of the group ----
  for (int g = 0; g < G.count; g++) {
    const PairDesc &D = desc[G.first + g];
    Affine T;  // state->invT (column-major), four 128-bit read-only loads
    {
      const float4 *tp = reinterpret_cast<const float4 *>(D.state->invT);
      const float4 c0 = __ldg(tp), c1 = __ldg(tp + 1), c2 = __ldg(tp + 2), c3 = __ldg(tp + 3);
      T.r[0][0] = c0.x; T.r[1][0] = c0.y; T.r[2][0] = c0.z;
      T.r[0][1] = c1.x; T.r[1][1] = c1.y; T.r[2][1] = c1.z;
      T.r[0][2] = c2.x; T.r[1][2] = c2.y; T.r[2][2] = c2.z;
      T.r[0][3] = c3.x; T.r[1][3] = c3.y; T.r[2][3] = c3.z;
    }
    const float4 *__restrict__ refPoints = D.refPoints;
    const float4 *__restrict__ refNormals = D.refNormals;
    const unsigned long long *__restrict__ zref = D.refZ[parity];
    int *__restrict__ corrImage = D.corrImage;

    int ri[TK];
    unsigned long long zr[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) {
      const int pix = base + k * NT + lane;
      ri[k] = -1;
      zr[k] = kEmptyZ;
      if (pix < numPixels) {
        if (MODE == 0) {
          ri[k] = z_index(__ldg(zref + pix), epoch);
        } else {
          ri[k] = __ldg(corrImage + pix);
          if (imgStats) zr[k] = __ldg(zref + pix);
        }
      }
    }
    float4 rp0[TK], rn0[TK];
    bool ok[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) {
      ok[k] = ri[k] >= 0 && curOk[k];
      if (ok[k]) {
        rn0[k] = __ldg(refNormals + ri[k]);
        rp0[k] = __ldg(refPoints + ri[k]);
      }
    }
    float midx = 0.0f, imgSum = 0.0f, imgNz = 0.0f, imgInl = 0.0f;
    if (MODE == 1 && imgStats) {
#pragma unroll
      for (int k = 0; k < TK; k++) {
        // DepthImage_convert_32FC1_to_16UC1 + mask + bitwise (abs diff & 255.0f) (pwn_matcher_base.cpp:167-190)
        const float dr = z_depth(zr[k], epoch, FLT_MAX);
        const unsigned short r16 = dr < FLT_MAX ? (unsigned short)(int)fmul(1000.0f, dr) : 0;
        if (c16[k] > 0 && r16 > 0) {
          const float df = fabsf(fsub((float)c16[k], (float)r16));
          const float dm = __uint_as_float(__float_as_uint(df) & 0x437F0000u);
          imgNz += 1.0f;
          if (dm < imgThreshold) imgInl += 1.0f;
          imgSum += dm;
        }
      }
    }

    // ---- per pixel: transform, gates in the reference's order, then the Linearizer term of an accepted pixel in place
    // (the thread that owns the pixel accumulates it; a pixel slot nobody in the warp accepted is skipped warp-wide) ----
    if (!omegaReady) {
      cp_async_wait_all();
      omegaReady = true;
    }
    TermAcc acc;
    term_clear(acc);
    float nc = 0.0f;
#pragma unroll
    for (int k = 0; k < TK; k++) {
      bool good = ok[k];
      f32x2 Rx = 0ull, Ry = 0ull, Rz = 0ull;
      float4 cpk = make_float4(0.f, 0.f, 0.f, 1.f), cnk = make_float4(0.f, 0.f, 0.f, 0.f);
      if (good) {
        cpk = S.cp[k * NT + lane];
        cnk = S.cn[k * NT + lane];
        float rpx, rpy, rpz, rnx, rny, rnz;
        xform_point(T, rp0[k].x, rp0[k].y, rp0[k].z, rpx, rpy, rpz);
        xform_normal(T, rn0[k].x, rn0[k].y, rn0[k].z, rnx, rny, rnz);
        if (MODE == 0) {
          midx += 1.0f;
          // correspondencefinder.cpp:69 zero normals, :78 normal angle, :84 distance, :87-99 curvature ratio
          if (cnk.w < 0.0f || dot3(rn0[k].x, rn0[k].y, rn0[k].z, rn0[k].x, rn0[k].y, rn0[k].z) == 0.0f) good = false;
          if (good && dot3(cnk.x, cnk.y, cnk.z, rnx, rny, rnz) < ac.normalThreshold) good = false;
          if (good) {
            const float dx = fsub(cpk.x, rpx), dy = fsub(cpk.y, rpy), dz = fsub(cpk.z, rpz);
            if (dot3(dx, dy, dz, dx, dy, dz) > ac.squaredThreshold) good = false;
          }
          if (good) {
            float rc = rn0[k].w;
            const float cc = cnk.w;  // already clamped
            if (rc < ac.flatCurvature) rc = ac.flatCurvature;
            // (rc + 1e-5) / (cc + 1e-5) in double, rounded to float; identical operands give exactly 1.  A float32
            // pre-test decides unless the quotient lies within 1e-4 (relative) of a threshold.
            if (rc != cc) {
              const float q = __fdividef(rc + 1e-5f, cc + 1e-5f);
              const float lo = ac.minRatio * (1.0f - 1e-4f), hi = ac.maxRatio * (1.0f + 1e-4f);
              const float loIn = ac.minRatio * (1.0f + 1e-4f), hiIn = ac.maxRatio * (1.0f - 1e-4f);
              if (q < lo || q > hi) {
                good = false;
              } else if (!(q > loIn && q < hiIn)) {
                const float ratio = (float)(((double)rc + 1e-5) / ((double)cc + 1e-5));
                if (ratio < ac.minRatio || ratio > ac.maxRatio) good = false;
              }
            }
          }
        }
        Rx = pk(rpx, rnx);
        Ry = pk(rpy, rny);
        Rz = pk(rpz, rnz);
      }
      if (MODE == 0 && writeCorr) {
        const int pix = base + k * NT + lane;
        if (pix < numPixels) corrImage[pix] = good ? ri[k] : -1;
      }
      if (__any_sync(0xffffffffu, good)) {
        if (good) {
          const ulonglong2 w0 = *reinterpret_cast<const ulonglong2 *>(&S.om[0][k * NT + lane]);
          const ulonglong2 w1 = *reinterpret_cast<const ulonglong2 *>(&S.om[1][k * NT + lane]);
          const ulonglong2 w2 = *reinterpret_cast<const ulonglong2 *>(&S.om[2][k * NT + lane]);
          term_add(acc, Rx, Ry, Rz, cpk, cnk, w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, ac.maxChi2, ac.robust);
          nc += 1.0f;
        }
      }
    }
    float v[kAccum];
    term_slots(acc, v);
    if (MODE == 0) {
      v[A_NCORR] = nc;
      v[A_MIDX] = midx;
      v[31] = 0.0f;
    } else {
      v[29] = imgSum;
      v[30] = imgNz;
      v[31] = imgInl;
    }
    const float tot = warp_transpose_reduce(v, lane);
    D.partials[(size_t)tileId * kAccum + lane] = tot;
  }
  if (!omegaReady) cp_async_wait_all();
}

}  // namespace nicp
