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

__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
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
// OPAQUE: the zeros come out of an asm statement, so the compiler cannot fold "0 + term" of the first pixel slot into a
// move and every slot updates the sums in place (used where the slots sit behind warp-uniform branches)
template <bool OPAQUE>
__device__ __forceinline__ void term_clear(TermAcc &A) {
  float z = 0.0f;
  f32x2 z2 = 0ull;
  if (OPAQUE) {
    asm volatile("mov.f32 %0, 0f00000000;" : "=f"(z));
    asm volatile("mov.b64 %0, 0;" : "=l"(z2));
  }
#pragma unroll
  for (int i = 0; i < 6; i++) { A.htt[i] = z; A.hrr[i] = z2; }
#pragma unroll
  for (int i = 0; i < 9; i++) A.htr[i] = z;
#pragma unroll
  for (int i = 0; i < 3; i++) { A.bt[i] = z; A.br[i] = z2; }
  A.err = z;
  A.inl = z;
}

// one correspondence (linearizer.cpp:56-89), executed by every lane of the warp: `w` is 1 for the lanes whose pixel was
// accepted and 0 for the others, whose Omega is multiplied away (everything below is linear in Omega, and every input is
// finite -- the shared-memory slots are zero-filled at kernel start), so there is no divergent control flow around the
// accumulators.  R = transformed reference (point, normal) per axis, packed; E0..E2 = the errors (rp - cp, rn - cn) per
// axis, packed;
// A=(a,g) B=(b,h) C=(c,i) D=(d,j) E=(e,k) F=(f,l) with Omega_P = [a b c; b d e; c e f], Omega_N = [g h i; h j k; i k l].
template <bool ROBUST, bool MASKED = true>
__device__ __forceinline__ void term_add_e(TermAcc &acc, float w, f32x2 Rx, f32x2 Ry, f32x2 Rz, f32x2 E0, f32x2 E1, f32x2 E2,
                                           f32x2 A, f32x2 B, f32x2 C, f32x2 D, f32x2 E, f32x2 F, float maxChi2, int robust) {
  if (MASKED) {  // (w == 1 for every caller with MASKED = false: the product with 1 is exact, the bits do not change)
    const f32x2 W = pk(w, w);
    A = mul2(A, W); B = mul2(B, W); C = mul2(C, W); D = mul2(D, W); E = mul2(E, W); F = mul2(F, W);
  }
  // Omega e, rows: (a e0 + b e1) + c e2 ...
  f32x2 W0 = fma2(C, E2, fma2(B, E1, mul2(A, E0)));
  f32x2 W1 = fma2(E, E2, fma2(D, E1, mul2(B, E0)));
  f32x2 W2 = fma2(F, E2, fma2(E, E1, mul2(C, E0)));
  const f32x2 chi2 = fma2(E2, W2, fma2(E1, W1, mul2(E0, W0)));
  float chi = __fadd_rn(lo_of(chi2), hi_of(chi2));
  float ks = 1.0f;
  if (ROBUST) {
    // sqrt(maxChi2 / chi) (linearizer.cpp:66-71) as r * rsqrt(r): two special-function results and a select, where sqrtf
    // brings a divergent region with its own slow-path call into every pixel slot; 2 ulp from the rounded square root
    // (the scale only weighs outlier terms in b and the error; H / b tolerance 1e-4).  chi = 0: inf * 0 is selected away.
    const float r = __fdividef(maxChi2, chi);
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(r));
    ks = chi > maxChi2 ? __fmul_rn(r, rs) : 1.0f;
  } else {
    // without the robust kernel such a correspondence is skipped altogether (linearizer.cpp:66-71): multiplied away
    const float keep = chi > maxChi2 ? 0.0f : 1.0f;
    const f32x2 K2 = pk(keep, keep);
    A = mul2(A, K2); B = mul2(B, K2); C = mul2(C, K2); D = mul2(D, K2); E = mul2(E, K2); F = mul2(F, K2);
    W0 = mul2(W0, K2); W1 = mul2(W1, K2); W2 = mul2(W2, K2);
    chi = __fmul_rn(chi, keep);
    w = __fmul_rn(w, keep);
  }
  (void)robust;
  acc.inl = __fadd_rn(acc.inl, w);
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
// the same from the current point / normal (per-pair kernel)
template <bool ROBUST, bool MASKED = true>
__device__ __forceinline__ void term_add(TermAcc &acc, float w, f32x2 Rx, f32x2 Ry, f32x2 Rz, float4 cp, float4 cn, f32x2 A, f32x2 B,
                                         f32x2 C, f32x2 D, f32x2 E, f32x2 F, float maxChi2, int robust) {
  term_add_e<ROBUST, MASKED>(acc, w, Rx, Ry, Rz, sub2(Rx, pk(cp.x, cn.x)), sub2(Ry, pk(cp.y, cn.y)), sub2(Rz, pk(cp.z, cn.z)), A, B, C,
                             D, E, F, maxChi2, robust);
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

// per-lane slots (lane l, pixel slot k -> [k * 32 + l]): nothing here is shared between lanes except T; shared memory is
// used as a register file extension -- for what must survive the walk over the pairs of the group, and as the landing
// zone of the reference gathers, which are issued with cp.async one pair AHEAD of the pair being computed (no registers
// are held while they are in flight)
template <int MAXG>
struct GroupSmem {  // shared by the warps of a CTA: the current side of the tile and T of every pair
  float4 om[3][96];     // Omega_P / Omega_N of the current point (Omega3 layout), cp.async
  float4 c0[96];        // current point / normal interleaved like nicp_cloud::pn: (px, nx, py, ny)
  float4 c1[96];        // (pz, nz, 1, curvature clamped to flatCurvatureThreshold, or -1 for a zero normal (MODE 0))
  float4 T[MAXG][4];    // state->invT of every pair of the group (column-major), fetched by lane g in the prologue
};
struct GatherStage {    // nicp_cloud::pn records of the pair being computed / the next pair, cp.async
  float4 p0[96];        // reference (px, nx, py, ny)
  float4 p1[96];        // reference (pz, nz, 1, curvature)
};                      // (a consumed stage doubles as the scratch of the shared-memory reduction)
static_assert(sizeof(GatherStage) >= 32 * 20 * sizeof(float), "reduction scratch fits a gather stage");

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ const float4 *shfl_ptr(const float4 *p, int src) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  unsigned int lo = __shfl_sync(0xffffffffu, (unsigned int)v, src), hi = __shfl_sync(0xffffffffu, (unsigned int)(v >> 32), src);
  return reinterpret_cast<const float4 *>(((unsigned long long)hi << 32) | lo);
}

// MODE 0: correspondence gates + linearise at state->invT; the correspondence image is written only when asked for.
// MODE 1: linearise over the stored correspondence image (inner iterations > 0, _computeStatistics) and, if imgStats,
//         accumulate the matchClouds image statistics (slots 29..31).
// Software pipeline over the pairs of the group (one exposed memory round trip per GROUP, not per pair):
//   while pair g is computed out of shared memory, the point / normal gathers of pair g + 1 are landing there (cp.async)
//   and the z-buffer words of pair g + 2 are on their way into registers.  No address is loaded inside the loop: the
//   slot buffers are affine in the slot index (SlotBases), the cloud pointers and T of all pairs are fetched by lane g
//   in the prologue and handed out with shuffles / shared memory.
// Deterministic cross-lane sums through shared memory, in two rounds of 16 values so that the scratch (32 rows x 20
// floats = 2560 bytes) fits into the gather stage the pair has just consumed: lane l parks 16 values as row l (padded
// rows: conflict-free 128-bit stores); lane j adds rows 0..15 of column j & 15 if j < 16, rows 16..31 otherwise, top to
// bottom, and one shuffle joins the halves.  After the two rounds lane j holds the total of value j.
// 8 STS.128 + 32 LDS + 32 FADD + 2 SHFL instead of the register butterfly's 31 SHFL + 62 FSEL + 31 FADD.
__device__ __forceinline__ float warp_transpose_reduce_smem(const float (&v)[kAccum], int lane, float *scratch) {
  float result = 0.0f;
  const int col = lane & 15, rowBase = lane & 16;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    float4 *row = reinterpret_cast<float4 *>(scratch + lane * 20);
#pragma unroll
    for (int q = 0; q < 4; q++) row[q] = make_float4(v[16 * r + 4 * q], v[16 * r + 4 * q + 1], v[16 * r + 4 * q + 2], v[16 * r + 4 * q + 3]);
    __syncwarp();
    float t[16];
#pragma unroll
    for (int l = 0; l < 16; l++) t[l] = scratch[(rowBase + l) * 20 + col];
    float sum = t[0];
#pragma unroll
    for (int l = 1; l < 16; l++) sum = __fadd_rn(sum, t[l]);
    const float other = __shfl_xor_sync(0xffffffffu, sum, 16);
    // rows 0..15 first, then rows 16..31, on both halves of the warp
    const float total = (lane & 16) ? __fadd_rn(other, sum) : __fadd_rn(sum, other);
    if ((lane >> 4) == r) result = total;
    __syncwarp();
  }
  return result;
}

// VAR bit 0: 1 = every pixel slot of a non-empty tile runs straight-line, 0 = a slot no lane needs is skipped behind a
//            warp-uniform branch;  bit 1: 1 = cross-lane sums through shared memory, 0 = register butterfly.
// NW warps per CTA share the current side of the tile and split the pairs of the group (warp w takes pairs w, w + NW, ...):
// less shared memory per warp, so more warps fit an SM.
template <int MODE, int MINB, bool PACKED, bool ROBUST, int VAR, int NW>
__global__ void __launch_bounds__(32 * NW, MINB) k_corr_lin_group(const PairDesc *__restrict__ desc, const PairGroup *__restrict__ groups,
                                                             SlotBases B, int epoch, int writeCorr, AlignConsts ac, int numPixels,
                                                             int imgStats, float imgThreshold, int groupFast, int curEpoch) {
  constexpr int TK = 3, NT = 32, TILE = NT * TK;
  constexpr bool STRAIGHT = (VAR & 1) != 0, SMEMRED = (VAR & 2) != 0;
  __shared__ GroupSmem<(NW > 1 ? kMaxGroup : kMaxGroup / 2)> S;  // one-warp CTAs walk at most kMaxGroup / 2 pairs
  __shared__ GatherStage stages[NW][2];
  const int warp = NW > 1 ? (int)(threadIdx.x >> 5) : 0;
  GatherStage *const stage = stages[warp];
  // group-fastest block order: the CTAs resident together work on the same tile of different groups, so the reference
  // z-buffer rows and the tile's current-cloud lines they touch stay close in L2
  const int groupId = groupFast ? blockIdx.x : blockIdx.y;
  const int tileId = groupFast ? blockIdx.y : blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int base = tileId * TILE;
  const bool wantZ = MODE == 0 || imgStats;  // the reference z-buffer words are needed (index / depth)
  PairGroup G;
  {
    const uint4 *gp = reinterpret_cast<const uint4 *>(groups + groupId);
    const uint4 g0 = __ldg(gp), g1 = __ldg(gp + 1), g2 = __ldg(gp + 2);
    G.first = (int)g0.x; G.count = (int)g0.y; G.curSlot = (int)g0.z; G.pad = 0;
    G.curPoints = reinterpret_cast<const float4 *>(((unsigned long long)g1.y << 32) | g1.x);
    G.curNormals = reinterpret_cast<const float4 *>(((unsigned long long)g1.w << 32) | g1.z);
    G.curOmega = reinterpret_cast<const float4 *>(((unsigned long long)g2.y << 32) | g2.x);
    G.curPN = reinterpret_cast<const float4 *>(((unsigned long long)g2.w << 32) | g2.z);
  }

  // what identifies the reference side of a pixel for one pair: the z-buffer word (MODE 0: index + epoch; MODE 1 with
  // image statistics: depth) and / or the stored correspondence (MODE 1)
  struct RefKey {
    unsigned long long z[TK];
    int ri[TK];
  };
  auto load_key = [&](int g, RefKey &key) {
    const size_t off = (size_t)(G.first + g) * (size_t)B.slotPixels;
#pragma unroll
    for (int k = 0; k < TK; k++) {
      const int pix = base + k * NT + lane;
      key.z[k] = kEmptyZ;
      key.ri[k] = -1;
      if (pix < numPixels) {
        if (wantZ) key.z[k] = __ldg(B.refZ + off + pix);
        if (MODE == 1) key.ri[k] = __ldg(B.corrImage + off + pix);
      }
    }
  };
  auto issue_gathers = [&](const float4 *refPN, const int (&ri)[TK], const bool (&curOk)[TK], int st) {
#pragma unroll
    for (int k = 0; k < TK; k++) {
      if (ri[k] >= 0 && curOk[k]) {  // the two halves of the point's 32-byte sector
        cp_async16(&stage[st].p0[k * NT + lane], refPN + 2 * (size_t)ri[k]);
        cp_async16(&stage[st].p1[k * NT + lane], refPN + 2 * (size_t)ri[k] + 1);
      }
    }
  };

  // every shared-memory slot of this lane holds finite values from the start (the per-pixel code below is branch free
  // and multiplies what a rejected lane read by zero)
  {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < TK; k++) {
      if (warp == 0) { S.om[0][k * NT + lane] = z4; S.om[1][k * NT + lane] = z4; S.om[2][k * NT + lane] = z4; }
      stage[0].p0[k * NT + lane] = z4; stage[1].p0[k * NT + lane] = z4;
      stage[0].p1[k * NT + lane] = z4; stage[1].p1[k * NT + lane] = z4;
    }
  }
  // ---- prologue: current side of the tile (once per group), keys of pairs 0 and 1, pointers and T of every pair ----
  const int *__restrict__ curIndex = B.curIndex + (size_t)G.curSlot * (size_t)B.slotPixels;
  int ci[TK];
#pragma unroll
  for (int k = 0; k < TK; k++) {
    const int pix = base + k * NT + lane;
    ci[k] = pix < numPixels ? __ldg(curIndex + pix) : -1;
  }
  RefKey keyNext;  // keys of the pair whose gathers are issued next
  const bool mine = warp < G.count;  // this warp has at least one pair
  if (mine) load_key(warp, keyNext);
  else {
#pragma unroll
    for (int k = 0; k < TK; k++) { keyNext.z[k] = kEmptyZ; keyNext.ri[k] = -1; }
  }
  // lane g fetches what pair g of the group needs: its cloud pointers (kept, handed out by shuffle) and its T (to shared
  // memory); count <= kMaxGroup
  const float4 *myRefPN = nullptr;
  if (lane < G.count) {
    myRefPN = desc[G.first + lane].refPN;
    if (warp == 0) {
      const float4 *tp = reinterpret_cast<const float4 *>(B.state[G.first + lane].invT);
      S.T[lane][0] = __ldg(tp);
      S.T[lane][1] = __ldg(tp + 1);
      S.T[lane][2] = __ldg(tp + 2);
      S.T[lane][3] = __ldg(tp + 3);
    }
  }
  // nothing of the current cloud projects into this tile: every pair of the group gets a row of zeros (MODE 1 with image
  // statistics still has to look at the reference z-buffer)
  if (!(MODE == 1 && imgStats) && !__any_sync(0xffffffffu, ci[0] >= 0 || ci[1] >= 0 || ci[2] >= 0)) {
    for (int g = warp; g < G.count; g += NW) {
      B.partials[(size_t)(G.first + g) * (size_t)B.partialStride + (size_t)tileId * kAccum + lane] = 0.0f;
      if (MODE == 0 && writeCorr) {
        int *__restrict__ corrImage = B.corrImage + (size_t)(G.first + g) * (size_t)B.slotPixels;
#pragma unroll
        for (int k = 0; k < TK; k++) {
          const int pix = base + k * NT + lane;
          if (pix < numPixels) corrImage[pix] = -1;
        }
      }
    }
    return;
  }
  unsigned long long zc[TK];
  if (MODE == 1 && imgStats) {
    const unsigned long long *__restrict__ zcur = B.curZ + (size_t)G.curSlot * (size_t)B.slotPixels;
#pragma unroll
    for (int k = 0; k < TK; k++) {
      const int pix = base + k * NT + lane;
      zc[k] = pix < numPixels ? __ldg(zcur + pix) : kEmptyZ;
    }
  }
  bool curOk[TK];
  int riCur[TK];                 // reference indices of the pair being computed
  unsigned long long zCur[TK];   // its z-buffer words (image statistics)
  {
    float4 c0l[TK], c1l[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) {
      curOk[k] = ci[k] >= 0;
      c0l[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      c1l[k] = make_float4(0.f, 0.f, 1.f, 0.f);
      if (curOk[k] && warp == 0) {
        const float4 *om = G.curOmega + 3 * (size_t)ci[k];
        cp_async16(&S.om[0][k * NT + lane], om);
        cp_async16(&S.om[1][k * NT + lane], om + 1);
        cp_async16(&S.om[2][k * NT + lane], om + 2);
        c0l[k] = __ldg(G.curPN + 2 * (size_t)ci[k]);
        c1l[k] = __ldg(G.curPN + 2 * (size_t)ci[k] + 1);
      }
    }
    // gathers of pair 0 ride in the same cp.async group as Omega
#pragma unroll
    for (int k = 0; k < TK; k++) {
      riCur[k] = MODE == 0 ? z_index(keyNext.z[k], epoch) : keyNext.ri[k];
      zCur[k] = keyNext.z[k];
    }
    issue_gathers(shfl_ptr(myRefPN, warp), riCur, curOk, 0);
    cp_async_commit();
    if (warp + NW < G.count) load_key(warp + NW, keyNext);
    // what the gates need from the current side (correspondencefinder.cpp:69, :87-93), prepared once per group
#pragma unroll
    for (int k = 0; k < TK; k++) {
      if (MODE == 0) {
        // a zero current normal rejects the pixel for every pair (the pixel still counts as "both indices valid")
        const float nx = c0l[k].y, ny = c0l[k].w, nz = c1l[k].y;
        if (dot3(nx, ny, nz, nx, ny, nz) == 0.0f) c1l[k].w = -1.0f;  // curvature >= 0
        else if (c1l[k].w < ac.flatCurvature) c1l[k].w = ac.flatCurvature;
      }
      if (warp == 0) {
        S.c0[k * NT + lane] = c0l[k];
        S.c1[k * NT + lane] = c1l[k];
      }
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
  if (NW > 1) {
    // the shared current side (cp / cn stores, Omega cp.async of warp 0) and T become visible to the other warps
    if (warp == 0) cp_async_wait_group<0>();
    __syncthreads();
  } else {
    __syncwarp();  // S.T of every pair visible to every lane
  }

  // ---- the pairs of the group ----
  for (int g = warp, it = 0; g < G.count; g += NW, it++) {
    const int st = it & 1;
    // pair g + 1: decode its keys (loaded one iteration ago), send its gathers into the other stage; pair g + 2: keys
    int riNext[TK];
    unsigned long long zNext[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) {
      riNext[k] = -1;
      zNext[k] = kEmptyZ;
    }
    if (g + NW < G.count) {
#pragma unroll
      for (int k = 0; k < TK; k++) {
        riNext[k] = MODE == 0 ? z_index(keyNext.z[k], epoch) : keyNext.ri[k];
        zNext[k] = keyNext.z[k];
      }
      issue_gathers(shfl_ptr(myRefPN, g + NW), riNext, curOk, st ^ 1);
      if (g + 2 * NW < G.count) load_key(g + 2 * NW, keyNext);
    }
    cp_async_commit();         // (an empty group when there is no next pair)
    cp_async_wait_group<1>();  // everything but the group just committed has landed: Omega and the gathers of pair g

    // state->invT of pair g (column-major, broadcast reads from shared memory), the rotation entries duplicated into both
    // halves of a packed register
    const float4 t0 = S.T[g][0], t1 = S.T[g][1], t2 = S.T[g][2], t3 = S.T[g][3];
    const f32x2 T00 = pk(t0.x, t0.x), T01 = pk(t1.x, t1.x), T02 = pk(t2.x, t2.x);
    const f32x2 T10 = pk(t0.y, t0.y), T11 = pk(t1.y, t1.y), T12 = pk(t2.y, t2.y);
    const f32x2 T20 = pk(t0.z, t0.z), T21 = pk(t1.z, t1.z), T22 = pk(t2.z, t2.z);
    const float T03 = t3.x, T13 = t3.y, T23 = t3.z;
    const f32x2 ONE = pk(ac.one, ac.one);
    auto sum2 = [&](f32x2 a, f32x2 b) { return fma2(b, ONE, a); };  // a + b, exactly rounded, never contracted
    int *__restrict__ corrImage = B.corrImage + (size_t)(G.first + g) * (size_t)B.slotPixels;

    // a tile where no lane has a pixel of this pair contributes a row of zeros (decided before any sum is live)
    bool ok0[TK];
#pragma unroll
    for (int k = 0; k < TK; k++) ok0[k] = riCur[k] >= 0 && curOk[k];
    const bool tileAny = __any_sync(0xffffffffu, ok0[0] || ok0[1] || ok0[2]) || (MODE == 1 && imgStats);
    float tot = 0.0f;
    if (tileAny) {
      float midx = 0.0f, imgSum = 0.0f, imgNz = 0.0f, imgInl = 0.0f;
      if (MODE == 1 && imgStats) {
#pragma unroll
        for (int k = 0; k < TK; k++) {
          // DepthImage_convert_32FC1_to_16UC1 + mask + bitwise (abs diff & 255.0f) (pwn_matcher_base.cpp:167-190)
          const float dr = z_depth(zCur[k], epoch, FLT_MAX);
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
      // ---- per pixel, STRAIGHT-LINE for every lane: transform, gates in the reference's order, Linearizer term weighted
      // by 1 (accepted) or 0.  No branch surrounds the sums, so they stay in place in their registers; a lane without a
      // pixel works on the zeros / stale finite values of its slots and its Omega is multiplied away. ----
      TermAcc acc;
      term_clear<!STRAIGHT>(acc);
      float sacc[kAccum];  // scalar formulation of the term (comparison variant, PACKED = false)
      if (!PACKED) {
#pragma unroll
        for (int i = 0; i < kAccum; i++) sacc[i] = 0.0f;
      }
      float nc = 0.0f;
#pragma unroll
      for (int k = 0; k < TK; k++) {
        if (!STRAIGHT) {
          if (!__any_sync(0xffffffffu, ok0[k])) {  // warp-uniform: no lane has a pixel in this slot
            if (MODE == 0 && writeCorr) {
              const int pix = base + k * NT + lane;
              if (pix < numPixels) corrImage[pix] = -1;
            }
            continue;
          }
        }
        // reference and current point + normal, interleaved: (px, nx, py, ny) (pz, nz, 1, curvature)
        const ulonglong2 r0 = *reinterpret_cast<const ulonglong2 *>(&stage[st].p0[k * NT + lane]);
        const ulonglong2 r1 = *reinterpret_cast<const ulonglong2 *>(&stage[st].p1[k * NT + lane]);
        const ulonglong2 q0 = *reinterpret_cast<const ulonglong2 *>(&S.c0[k * NT + lane]);
        const ulonglong2 q1 = *reinterpret_cast<const ulonglong2 *>(&S.c1[k * NT + lane]);
        const f32x2 Px = r0.x, Py = r0.y, Pz = r1.x, Cx = q0.x, Cy = q0.y, Cz = q1.x;
        // T (p, 1) in the low halves and T (n, 0) in the high halves, row by row in the order of xform_point /
        // xform_normal: ((m0 v0 + m1 v1) + m2 v2) [+ m3]; products and sums rounded separately, so each half is bit for
        // bit what the scalar functions return.  The sums are written b * 1 + a with a run-time 1: ptxas contracts
        // mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (unlike the scalar .rn forms), which rounds once.
        f32x2 Rx = sum2(sum2(mul2(T00, Px), mul2(T01, Py)), mul2(T02, Pz));
        f32x2 Ry = sum2(sum2(mul2(T10, Px), mul2(T11, Py)), mul2(T12, Pz));
        f32x2 Rz = sum2(sum2(mul2(T20, Px), mul2(T21, Py)), mul2(T22, Pz));
        Rx = pk(fadd(lo_of(Rx), T03), hi_of(Rx));
        Ry = pk(fadd(lo_of(Ry), T13), hi_of(Ry));
        Rz = pk(fadd(lo_of(Rz), T23), hi_of(Rz));
        const f32x2 E0 = sub2(Rx, Cx), E1 = sub2(Ry, Cy), E2 = sub2(Rz, Cz);  // (rp - cp, rn - cn)
        bool good = ok0[k];
        if (MODE == 0) {
          midx += ok0[k] ? 1.0f : 0.0f;
          const float rnx0 = hi_of(Px), rny0 = hi_of(Py), rnz0 = hi_of(Pz);
          const float cc = hi_of(q1.y);  // current curvature, already clamped; -1 = zero current normal
          // correspondencefinder.cpp:69 zero normals, :78 normal angle, :84 distance, :87-99 curvature ratio
          good = good && !(cc < 0.0f) && dot3(rnx0, rny0, rnz0, rnx0, rny0, rnz0) != 0.0f;
          good = good && !(dot3(hi_of(Cx), hi_of(Cy), hi_of(Cz), hi_of(Rx), hi_of(Ry), hi_of(Rz)) < ac.normalThreshold);
          const float dx = lo_of(E0), dy = lo_of(E1), dz = lo_of(E2);  // (rp - cp)^2 = (cp - rp)^2 bit for bit
          good = good && !(dot3(dx, dy, dz, dx, dy, dz) > ac.squaredThreshold);
          float rc = hi_of(r1.y);
          if (rc < ac.flatCurvature) rc = ac.flatCurvature;
          // (rc + 1e-5) / (cc + 1e-5) in double, rounded to float; identical operands give exactly 1.  A float32
          // pre-test decides unless the quotient lies within 1e-4 (relative) of a threshold.
          const float q = __fdividef(rc + 1e-5f, cc + 1e-5f);
          const float lo = ac.minRatio * (1.0f - 1e-4f), hi = ac.maxRatio * (1.0f + 1e-4f);
          const float loIn = ac.minRatio * (1.0f + 1e-4f), hiIn = ac.maxRatio * (1.0f - 1e-4f);
          const bool differ = rc != cc;
          good = good && !(differ && (q < lo || q > hi));
          if (good && differ && !(q > loIn && q < hiIn)) {  // rare: the exact quotient decides
            const float ratio = (float)(((double)rc + 1e-5) / ((double)cc + 1e-5));
            if (ratio < ac.minRatio || ratio > ac.maxRatio) good = false;
          }
        }
        if (MODE == 0 && writeCorr) {
          const int pix = base + k * NT + lane;
          if (pix < numPixels) corrImage[pix] = good ? riCur[k] : -1;
        }
        const float w = good ? 1.0f : 0.0f;
        nc += w;
        if (!STRAIGHT) {
          if (!__any_sync(0xffffffffu, good)) continue;  // warp-uniform: every lane of this slot was rejected
        }
        if (PACKED) {
          const ulonglong2 w0 = *reinterpret_cast<const ulonglong2 *>(&S.om[0][k * NT + lane]);
          const ulonglong2 w1 = *reinterpret_cast<const ulonglong2 *>(&S.om[1][k * NT + lane]);
          const ulonglong2 w2 = *reinterpret_cast<const ulonglong2 *>(&S.om[2][k * NT + lane]);
          term_add_e<ROBUST>(acc, w, Rx, Ry, Rz, E0, E1, E2, w0.x, w0.y, w1.x, w1.y, w2.x, w2.y, ac.maxChi2, ac.robust);
        } else if (good) {
          accumulate_term(sacc, lo_of(Rx), lo_of(Ry), lo_of(Rz), hi_of(Rx), hi_of(Ry), hi_of(Rz),
                          make_float4(lo_of(Cx), lo_of(Cy), lo_of(Cz), 1.0f), make_float4(hi_of(Cx), hi_of(Cy), hi_of(Cz), 0.0f),
                          S.om[0][k * NT + lane], S.om[1][k * NT + lane], S.om[2][k * NT + lane], ac.maxChi2, ac.robust);
        }
      }
      float v[kAccum];
      if (PACKED) {
        term_slots(acc, v);
      } else {
#pragma unroll
        for (int i = 0; i < kAccum; i++) v[i] = sacc[i];
      }
      if (MODE == 0) {
        v[A_NCORR] = nc;
        v[A_MIDX] = midx;
        v[31] = 0.0f;
      } else {
        v[29] = imgSum;
        v[30] = imgNz;
        v[31] = imgInl;
      }
      if (SMEMRED) {
        tot = warp_transpose_reduce_smem(v, lane, reinterpret_cast<float *>(&stage[st]));
        // the scratch goes back to being a gather stage: finite values everywhere (rejected lanes read their slots)
      } else {
        tot = warp_transpose_reduce(v, lane);
      }
    } else if (MODE == 0 && writeCorr) {
#pragma unroll
      for (int k = 0; k < TK; k++) {
        const int pix = base + k * NT + lane;
        if (pix < numPixels) corrImage[pix] = -1;
      }
    }
    B.partials[(size_t)(G.first + g) * (size_t)B.partialStride + (size_t)tileId * kAccum + lane] = tot;
#pragma unroll
    for (int k = 0; k < TK; k++) {
      riCur[k] = riNext[k];
      zCur[k] = zNext[k];
    }
  }
  cp_async_wait_group<0>();
}

}  // namespace nicp
