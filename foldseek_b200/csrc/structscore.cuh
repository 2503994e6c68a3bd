// structscore.cuh — LDDT and TM-score of the traced alignments (structurealign.cpp:376-412).
//
// Reference behaviour reproduced:
//   LDDTCalculator::initQuery / computeLDDTScore     src/commons/LDDT.cpp:87-215, LDDT.h:102-119
//   TMaligner::computeAppoximateTMscore               src/commons/TMaligner.cpp:50-104
//     standard_TMscore / detailed_search_standard     lib/tmalign/TMalign.cpp:1366-1427, 589-626
//     TMscore8_search_standard, score_fun8            lib/tmalign/TMalign.cpp:394-554, 76-141
//     KabschFast -> kabsch_quat_soa_avx               lib/tmalign/TMalign.cpp:191-222, lib/tmalign/Kabsch.h:474-758
//     BasicFunction::do_rotation                      lib/tmalign/basic_fun.h:73-122
// Float parity: every float operation is issued with an explicit rounding intrinsic in the order the reference binary executes
// it (gcc contracts the vector mul+add pairs: rotation = t + fma(u2,z, fma(u0,x, u1*y)); squared distance = fma(dz,dz, fma(dx,dx,
// dy*dy)); covariance = fma(-mean1, sum2, cross)); the AVX Kabsch sums keep their 4-lane / hadd association; the TM sum over the
// aligned columns is sequential like the reference's scalar loop. The eigenvalue / quaternion part runs in double.
//
// Mapping: one warp per traced pair. Aligned coordinates live in shared memory (13 floats per aligned column); element-wise work
// (gather, rotate, distances, compaction) is spread over the 32 lanes; the order-sensitive reductions are 4-lane chains (Kabsch:
// lane = quantity group x chain) and one sequential chain (TM sum, lane 0).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef SS_EXPLICIT_DOUBLE
#define SS_EXPLICIT_DOUBLE 1   // 0: let nvcc contract the double section itself (same results on sm_100a with CUDA 12.9, tools/ss_compare.py)
#endif

namespace ss {

constexpr int SS_FLOATS_PER_COL = 13;   // xtm ytm xt (9) + di term (2) + iAli kAli (2)

struct SsPair {
    uint64_t qCaOff, tCaOff;    // residue offsets into the query / target coordinate arrays
    uint64_t btOff;             // backtrace (device buffer of the traceback)
    uint32_t qLen, tLen;
    int32_t qStart, tStart;
    uint32_t btLen;
    uint32_t nAligned;          // number of 'M' columns (host counted)
    uint64_t qNormOff;          // into qNorm (LDDT 1/count per query residue)
    // parameter_set4search / standard_TMscore parameters (computed on the host with the reference's double libm calls)
    float Lnorm, scoreD8, d0Search, d0SearchClipped, d0Std;
    float pad;
};

struct SsOut {
    double tm;                  // max(detailed search, standard TM-score) like TMaligner.cpp:101
    float lddt;                 // avgLddtScore
    int32_t flags;              // bit 0: a Kabsch rotation was NaN (the reference falls back to the classic Kabsch(); not reproduced)
};

struct SsArgs {
    const SsPair *pairs;
    uint32_t nPairs;
    uint32_t *counter;
    const float *qx, *qy, *qz, *tx, *ty, *tz;
    const float *qNorm;
    const char *bt;
    SsOut *out;
    uint32_t colsPerWarp;       // shared-memory capacity (aligned columns) of one warp
    uint32_t jobsPerPair;       // warps that share one pair (independent search starts / LDDT rows), combined through `scratch`
    float *scratch;             // per pair: colsPerWarp LDDT column sums + 8 words (best1, best2, nan flag, jobs done), zeroed by the host
};

// ---- LDDT norm per query residue: 1 / #{r != q : dist(q, r) < 15}, INF when empty (LDDT.cpp:102-116) ----
__device__ __forceinline__ float lddt_dist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

__global__ void lddt_norm_kernel(const float *x, const float *y, const float *z, const uint64_t *off, const uint32_t *len, uint32_t nq, float *norm) {
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const uint64_t o = off[q];
    const uint32_t L = len[q];
    for (uint32_t col = threadIdx.x; col < L; col += blockDim.x) {
        const float cx = x[o + col], cy = y[o + col], cz = z[o + col];
        uint32_t cnt = 0;
        for (uint32_t row = 0; row < L; row++)
            cnt += (row != col) && (lddt_dist(x[o + row], y[o + row], z[o + row], cx, cy, cz) < 15.0f);
        norm[o + col] = cnt ? __fdiv_rn(1.0f, (float) cnt) : __int_as_float(0x7f800000);
    }
}

struct Warp {
    int lane;
    float *xtm[3], *ytm[3], *xt[3], *di, *term;
    int *iAli, *kAli;
    int n;

    // kabsch_quat_soa_avx over the columns idx[0..cnt) (idx == NULL: columns first..first+cnt). Returns false on a NaN rotation.
    __device__ bool kabsch(const int *idx, int first, int cnt, float t[3], float u[3][3]) const {
        const int j = lane & 3, g = lane >> 2;
        float a0 = 0.0f, a1 = 0.0f;
        for (int k = j; k < cnt; k += 4) {
            const int c = idx ? idx[k] : first + k;
            const float x1 = xtm[0][c], y1 = xtm[1][c], z1 = xtm[2][c], x2 = ytm[0][c], y2 = ytm[1][c], z2 = ytm[2][c];
            float v0, v1;
            switch (g) {
                case 0: v0 = x1; v1 = x2; break;
                case 1: v0 = __fmul_rn(x1, x2); v1 = 0.0f; break;
                case 2: v0 = y1; v1 = y2; break;
                case 3: v0 = z1; v1 = z2; break;
                case 4: v0 = __fmul_rn(x1, y2); v1 = __fmul_rn(y1, z2); break;
                case 5: v0 = __fmul_rn(x1, z2); v1 = __fmul_rn(z1, x2); break;
                case 6: v0 = __fmul_rn(y1, x2); v1 = __fmul_rn(z1, y2); break;
                default: v0 = __fmul_rn(y1, y2); v1 = __fmul_rn(z1, z2); break;
            }
            a0 = __fadd_rn(a0, v0);
            a1 = __fadd_rn(a1, v1);
        }
        // hadd tree: (l0 + l1) + (l2 + l3)
        a0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, 1)); a0 = __fadd_rn(a0, __shfl_xor_sync(0xffffffffu, a0, 2));
        a1 = __fadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, 1)); a1 = __fadd_rn(a1, __shfl_xor_sync(0xffffffffu, a1, 2));
        const float s1x = __shfl_sync(0xffffffffu, a0, 0), s2x = __shfl_sync(0xffffffffu, a1, 0), sxx = __shfl_sync(0xffffffffu, a0, 4);
        const float s1y = __shfl_sync(0xffffffffu, a0, 8), s2y = __shfl_sync(0xffffffffu, a1, 8), s1z = __shfl_sync(0xffffffffu, a0, 12), s2z = __shfl_sync(0xffffffffu, a1, 12);
        const float sxy = __shfl_sync(0xffffffffu, a0, 16), syz = __shfl_sync(0xffffffffu, a1, 16), sxz = __shfl_sync(0xffffffffu, a0, 20), szx = __shfl_sync(0xffffffffu, a1, 20);
        const float syx = __shfl_sync(0xffffffffu, a0, 24), szy = __shfl_sync(0xffffffffu, a1, 24), syy = __shfl_sync(0xffffffffu, a0, 28), szz = __shfl_sync(0xffffffffu, a1, 28);
        const float inv = __fdiv_rn(1.0f, (float) cnt);
        const float m1x = __fmul_rn(s1x, inv), m1y = __fmul_rn(s1y, inv), m1z = __fmul_rn(s1z, inv);
        const float m2x = __fmul_rn(s2x, inv), m2y = __fmul_rn(s2y, inv), m2z = __fmul_rn(s2z, inv);
        float r[9];
        r[0] = __fsub_rn(sxx, __fmul_rn(m1x, s2x));
        r[1] = __fmaf_rn(-m1x, s2y, sxy); r[2] = __fmaf_rn(-m1x, s2z, sxz);
        r[3] = __fmaf_rn(-m1y, s2x, syx); r[4] = __fmaf_rn(-m1y, s2y, syy);
        r[5] = __fmaf_rn(-m2z, s1y, syz); r[6] = __fmaf_rn(-m2x, s1z, szx);
        r[7] = __fmaf_rn(-m2y, s1z, szy); r[8] = __fmaf_rn(-m2z, s1z, szz);
#define SS_DOT3(a0_, a1_, a2_, b0_, b1_, b2_) __fadd_rn(__fadd_rn(0.0f, __fmul_rn(a0_, b0_)), __fadd_rn(__fmul_rn(a1_, b1_), __fmul_rn(a2_, b2_)))
        const float r0r0 = SS_DOT3(r[0], r[1], r[2], r[0], r[1], r[2]), r0r1 = SS_DOT3(r[0], r[1], r[2], r[3], r[4], r[5]);
        const float r1r1 = SS_DOT3(r[3], r[4], r[5], r[3], r[4], r[5]), r1r2 = SS_DOT3(r[3], r[4], r[5], r[6], r[7], r[8]);
        const float r0r2 = SS_DOT3(r[0], r[1], r[2], r[6], r[7], r[8]), r2r2 = SS_DOT3(r[6], r[7], r[8], r[6], r[7], r[8]);
        const float c1 = __fsub_rn(__fmul_rn(r[4], r[8]), __fmul_rn(r[7], r[5]));
        const float c2 = __fsub_rn(__fmul_rn(r[5], r[6]), __fmul_rn(r[8], r[3]));
        const float c3 = __fsub_rn(__fmul_rn(r[3], r[7]), __fmul_rn(r[6], r[4]));
        const float detf = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(r[0], c1)), __fadd_rn(__fmul_rn(r[1], c2), __fmul_rn(r[2], c3)));
#undef SS_DOT3
#if SS_EXPLICIT_DOUBLE
        // the reference binary's fused sequence (oracle/fs_oracle.cpp: tmo::eig_rot), independent of the compiler's contraction choices;
        // verified on a B200: 93 of 93 TM values of tests/test_structscore_gpu.py bit-identical to the oracle (tools/ss_compare.py)
        double ud[3][3];
        {
            const double det = detf;
            const double rr0 = r0r0, rr1 = r0r1, rr2 = r1r1, rr3 = r0r2, rr4 = r1r2, rr5 = r2r2;
            const double inv3 = 1.0 / 3.0, SQRT3 = 1.732050807568877;
            const double spur = __dmul_rn(__dadd_rn(__dadd_rn(rr0, rr2), rr5), inv3);
            double tt = __dmul_rn(rr2, rr5);
            tt = __fma_rn(-rr4, rr4, tt); tt = __fma_rn(rr5, rr0, tt); tt = __fma_rn(-rr3, rr3, tt);
            tt = __fma_rn(rr0, rr2, tt); tt = __fma_rn(-rr1, rr1, tt);
            const double cof = __dmul_rn(tt, inv3);
            double e0 = spur, e1 = spur, e2 = spur;
            const double h = (spur > 0) ? __fma_rn(spur, spur, -cof) : -1.0;
            if (h > 0) {
                const double detsq = __dmul_rn(det, det);
                double gg = __fma_rn(cof, spur, -detsq);
                gg = __fma_rn(gg, 0.5, -__dmul_rn(spur, h));
                const double sqrth = __dsqrt_rn(h);
                double d1 = __fma_rn(__dmul_rn(h, h), h, -__dmul_rn(gg, gg));
                d1 = (d1 < 0) ? __dmul_rn(atan2(0.0, -gg), inv3) : __dmul_rn(atan2(__dsqrt_rn(d1), -gg), inv3);
                const double cc = cos(d1), sn = sin(d1);
                const double cth = __dmul_rn(sqrth, cc), k3 = __dmul_rn(sqrth, SQRT3);
                e0 = __fma_rn(cth, 2.0, spur);
                e1 = __dadd_rn(__fma_rn(k3, sn, -cth), spur);
                e2 = __dadd_rn(-__fma_rn(k3, sn, cth), spur);
            }
            e0 = (e0 < 0) ? 0 : __dsqrt_rn(e0); e1 = (e1 < 0) ? 0 : __dsqrt_rn(e1); e2 = (e2 < 0) ? 0 : __dsqrt_rn(e2);
            const double ev = (det < 0) ? __dsub_rn(__dadd_rn(e0, e1), e2) : __dadd_rn(e2, __dadd_rn(e0, e1));
            const double r00 = r[0], r01 = r[1], r02 = r[2], r10 = r[3], r11 = r[4], r12 = r[5], r20 = r[6], r21 = r[7], r22 = r[8];
            const double b01 = __dsub_rn(r12, r21), b02 = __dsub_rn(r20, r02), b03 = __dsub_rn(r01, r10);
            const double b12 = __dadd_rn(r01, r10), b13 = __dadd_rn(r20, r02), b23 = __dadd_rn(r12, r21);
            const double b11 = __dsub_rn(__dsub_rn(__dsub_rn(r00, r11), r22), ev), b22 = __dsub_rn(__dsub_rn(__dsub_rn(r11, r00), r22), ev);
            const double b33 = __dsub_rn(__dadd_rn(__dsub_rn(-r00, r11), r22), ev);
            const double p2233 = __fma_rn(b22, b33, -__dmul_rn(b23, b23)), p1233 = __fma_rn(b12, b33, -__dmul_rn(b13, b23));
            const double p1223 = __fma_rn(b12, b23, -__dmul_rn(b13, b22)), p0232 = __fma_rn(b02, b23, -__dmul_rn(b03, b22));
            const double p0233 = __fma_rn(b33, b02, -__dmul_rn(b03, b23)), p0231 = __fma_rn(b02, b13, -__dmul_rn(b03, b12));
            const double q0 = __fma_rn(b13, p1223, __fma_rn(b11, p2233, -__dmul_rn(b12, p1233)));
            const double q1 = __fma_rn(-b13, p0232, __fma_rn(-p2233, b01, __dmul_rn(b12, p0233)));
            const double q2 = __fma_rn(b13, p0231, __fma_rn(p1233, b01, -__dmul_rn(b11, p0233)));
            const double q3 = __fma_rn(-b12, p0231, __fma_rn(-b01, p1223, __dmul_rn(b11, p0232)));
            const double q00 = __dmul_rn(q0, q0), q11 = __dmul_rn(q1, q1), q22 = __dmul_rn(q2, q2), q33 = __dmul_rn(q3, q3);
            const double il = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dadd_rn(q00, q11), q22), q33));
            const double q12 = __dmul_rn(q1, q2), q13 = __dmul_rn(q1, q3), q23 = __dmul_rn(q2, q3);
            const double xj = __dmul_rn(q11, il), az = __dmul_rn(__dmul_rn(q0, q3), il), ay = __dmul_rn(__dmul_rn(q0, q2), il), ax = __dmul_rn(__dmul_rn(q0, q1), il);
            const double ajmxj = __fma_rn(q00, il, -xj);
            ud[0][0] = __fma_rn(-q33, il, __fma_rn(-q22, il, __fma_rn(q00, il, xj)));
            ud[0][1] = __dmul_rn(2.0, __fma_rn(q12, il, az)); ud[0][2] = __dmul_rn(2.0, __fma_rn(q13, il, -ay));
            ud[1][0] = __dmul_rn(2.0, __fma_rn(q12, il, -az)); ud[1][1] = __fma_rn(-q33, il, __fma_rn(q22, il, ajmxj));
            ud[1][2] = __dmul_rn(2.0, __fma_rn(q23, il, ax)); ud[2][0] = __dmul_rn(2.0, __fma_rn(q13, il, ay));
            ud[2][1] = __dmul_rn(2.0, __fma_rn(q23, il, -ax)); ud[2][2] = __fma_rn(q33, il, __fma_rn(-q22, il, ajmxj));
        }
#else
        const double det = detf, detsq = det * det;
        const double rr0 = r0r0, rr1 = r0r1, rr2 = r1r1, rr3 = r0r2, rr4 = r1r2, rr5 = r2r2;
        const double inv3 = 1.0 / 3.0, SQRT3 = 1.732050807568877;
        const double spur = (rr0 + rr2 + rr5) * inv3;
        const double cof = (rr2 * rr5 - rr4 * rr4 + rr0 * rr5 - rr3 * rr3 + rr0 * rr2 - rr1 * rr1) * inv3;
        double e0 = spur, e1 = spur, e2 = spur;
        const double h = (spur > 0) ? spur * spur - cof : -1.0;
        if (h > 0) {
            const double gg = (spur * cof - detsq) * 0.5 - spur * h;
            const double sqrth = sqrt(h);
            double d1 = h * h * h - gg * gg;
            d1 = (d1 < 0) ? atan2(0.0, -gg) * inv3 : atan2(sqrt(d1), -gg) * inv3;
            const double cth = sqrth * cos(d1), sth = sqrth * SQRT3 * sin(d1);
            e0 += cth + cth; e1 += -cth + sth; e2 += -cth - sth;
        }
        e0 = (e0 < 0) ? 0 : sqrt(e0); e1 = (e1 < 0) ? 0 : sqrt(e1); e2 = (e2 < 0) ? 0 : sqrt(e2);
        const double ev = (det < 0) ? e0 + e1 - e2 : e0 + e1 + e2;
        const double m00 = r[0], m01 = r[1], m02 = r[2], m10 = r[3], m11 = r[4], m12 = r[5], m20 = r[6], m21 = r[7], m22 = r[8];
        const double b01 = m12 - m21, b02 = m20 - m02, b03 = m01 - m10;
        double b11 = m00 - m11 - m22; const double b12 = m01 + m10, b13 = m20 + m02;
        double b22 = -m00 + m11 - m22; const double b23 = m12 + m21; double b33 = -m00 - m11 + m22;
        b11 -= ev; b22 -= ev; b33 -= ev;
        const double p2233 = b22 * b33 - b23 * b23, p1233 = b12 * b33 - b13 * b23, p1223 = b12 * b23 - b13 * b22;
        const double p0232 = b02 * b23 - b03 * b22, p0233 = b02 * b33 - b03 * b23, p0231 = b02 * b13 - b03 * b12;
        const double q0 = b11 * p2233 - b12 * p1233 + b13 * p1223, q1 = -b01 * p2233 + b12 * p0233 - b13 * p0232;
        const double q2 = b01 * p1233 - b11 * p0233 + b13 * p0231, q3 = -b01 * p1223 + b11 * p0232 - b12 * p0231;
        const double il = 1.0 / (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
        const double aj = q0 * q0 * il, xj = q1 * q1 * il, yj = q2 * q2 * il, zj = q3 * q3 * il;
        const double xy = q1 * q2 * il, az = q0 * q3 * il, zx = q3 * q1 * il, ay = q0 * q2 * il, yz = q2 * q3 * il, ax = q0 * q1 * il;
        double ud[3][3];
        ud[0][0] = aj + xj - yj - zj; ud[0][1] = 2.0 * (xy + az); ud[0][2] = 2.0 * (zx - ay);
        ud[1][0] = 2.0 * (xy - az); ud[1][1] = aj - xj + yj - zj; ud[1][2] = 2.0 * (yz + ax);
        ud[2][0] = 2.0 * (zx + ay); ud[2][1] = 2.0 * (yz - ax); ud[2][2] = aj - xj - yj + zj;
#endif
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const float n0 = (float) (-ud[0][i]), n1 = (float) (-ud[1][i]), n2 = (float) (-ud[2][i]);
            const float c2i = i == 0 ? m2x : i == 1 ? m2y : m2z;
            t[i] = __fadd_rn(__fadd_rn(__fmul_rn(n0, m1x), __fmul_rn(n1, m1y)), __fadd_rn(__fmul_rn(n2, m1z), __fmul_rn(c2i, 1.0f)));
            u[i][0] = (float) ud[0][i]; u[i][1] = (float) ud[1][i]; u[i][2] = (float) ud[2][i];
            ok = ok && !(isnan(u[i][0]) || isnan(u[i][1]) || isnan(u[i][2]));
        }
        return ok;
    }

    __device__ void rotate(const float t[3], const float u[3][3]) const {
        for (int i = lane; i < n; i += 32) {
            const float x = xtm[0][i], y = xtm[1][i], z = xtm[2][i];
#pragma unroll
            for (int r = 0; r < 3; r++) xt[r][i] = __fadd_rn(t[r], __fmaf_rn(u[r][2], z, __fmaf_rn(u[r][0], x, __fmul_rn(u[r][1], y))));
        }
        __syncwarp();
    }

    // score_fun8: returns n_cut, fills iAli, score1 = sum / Lnorm (Lnorm = n inside the search)
    __device__ int score(float d, float scoreD8, float d0, float &score1) const {
        const float d02 = __fmul_rn(d0, d0), cut = __fmul_rn(scoreD8, scoreD8);
        for (int i = lane; i < n; i += 32) {
            const float dx = __fsub_rn(xt[0][i], ytm[0][i]), dy = __fsub_rn(xt[1][i], ytm[1][i]), dz = __fsub_rn(xt[2][i], ytm[2][i]);
            const float dd = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            di[i] = dd;
            term[i] = (dd < cut) ? __fdiv_rn(1.0f, __fadd_rn(1.0f, __fdiv_rn(dd, d02))) : 0.0f;
        }
        __syncwarp();
        float sum = 0.0f;
        if (lane == 0) {
            int i = 0;
            for (; i + 8 <= n; i += 8) {
                const float v0 = term[i], v1 = term[i + 1], v2 = term[i + 2], v3 = term[i + 3], v4 = term[i + 4], v5 = term[i + 5], v6 = term[i + 6], v7 = term[i + 7];
                sum = __fadd_rn(sum, v0); sum = __fadd_rn(sum, v1); sum = __fadd_rn(sum, v2); sum = __fadd_rn(sum, v3);
                sum = __fadd_rn(sum, v4); sum = __fadd_rn(sum, v5); sum = __fadd_rn(sum, v6); sum = __fadd_rn(sum, v7);
            }
            for (; i < n; i++) sum = __fadd_rn(sum, term[i]);
        }
        sum = __shfl_sync(0xffffffffu, sum, 0);
        float dTmp = __fmul_rn(d, d);
        int inc = 0, nCut;
        for (;;) {
            nCut = 0;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const bool in = i < n && di[i] < dTmp;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) iAli[nCut + __popc(m & ((1u << lane) - 1u))] = i;
                nCut += __popc(m);
            }
            __syncwarp();
            if (nCut < 3 && n > 3) { inc++; const double dinc = (double) d + inc * 0.5; dTmp = (float) (dinc * dinc); }
            else break;
        }
        score1 = __fdiv_rn(sum, (float) n);
        return nCut;
    }

    // TMscore8_search_standard with simplify_step 40
    // The (fragment length, start) combinations are independent searches whose maximum is taken: job `jobIdx` of `nJobs` runs every
    // nJobs-th of them (the caller combines the partial maxima).
    __device__ float search(float localD0Search, float scoreD8, float d0, bool &nanSeen, int jobIdx, int nJobs) const {
        float t[3], u[3][3];
        const int nIt = 20, nInitMax = 6;
        int Lini[6], LiniMin = 4;
        if (n < LiniMin) LiniMin = n;
        int nInit = 0, ii;
        for (ii = 0; ii < nInitMax - 1; ii++) {
            nInit++;
            Lini[ii] = n >> ii;            // (int)(Lali / pow(2.0, i)) for i <= 4
            if (Lini[ii] <= LiniMin) { Lini[ii] = LiniMin; break; }
        }
        if (ii == nInitMax - 1) { nInit++; Lini[ii] = LiniMin; }
        float scoreMax = -1.0f, sc;
        int combo = 0;
        for (int iInit = 0; iInit < nInit; iInit++) {
            const int Lfrag = Lini[iInit], iLmax = n - Lfrag;
            int i = 0;
            for (;;) {
                if ((combo++ % nJobs) != jobIdx) {
                    if (i < iLmax) { i += 40; if (i > iLmax) i = iLmax; continue; }
                    else break;
                }
                if (!kabsch(NULL, i, Lfrag, t, u)) nanSeen = true;
                rotate(t, u);
                int nCut = score(__fsub_rn(localD0Search, 1.0f), scoreD8, d0, sc);
                if (sc > scoreMax) scoreMax = sc;
                int ka = Lfrag;          // k_ali = i .. i + Lfrag - 1
                for (int it = 0; it < nIt; it++) {
                    // k_ali <- i_ali
                    for (int k = lane; k < nCut; k += 32) kAli[k] = iAli[k];
                    __syncwarp();
                    ka = nCut;
                    if (!kabsch(kAli, 0, nCut, t, u)) nanSeen = true;
                    rotate(t, u);
                    nCut = score(__fadd_rn(localD0Search, 1.0f), scoreD8, d0, sc);
                    if (sc > scoreMax) scoreMax = sc;
                    if (nCut == ka) {
                        bool same = true;
                        for (int k = lane; k < nCut; k += 32) same = same && (iAli[k] == kAli[k]);
                        if (__all_sync(0xffffffffu, same)) break;
                    }
                }
                if (i < iLmax) { i += 40; if (i > iLmax) i = iLmax; }
                else break;
            }
        }
        return scoreMax;
    }
};

__global__ void __launch_bounds__(256) structscore_kernel(SsArgs a) {
    extern __shared__ __align__(16) float ssmem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *base = ssmem + (size_t) warp * a.colsPerWarp * SS_FLOATS_PER_COL;
    const uint32_t cap = a.colsPerWarp;
    Warp w;
    w.lane = lane;
    for (int r = 0; r < 3; r++) { w.xtm[r] = base + (size_t) r * cap; w.ytm[r] = base + (size_t) (3 + r) * cap; w.xt[r] = base + (size_t) (6 + r) * cap; }
    w.di = base + (size_t) 9 * cap; w.term = base + (size_t) 10 * cap;
    w.iAli = reinterpret_cast<int *>(base + (size_t) 11 * cap); w.kAli = reinterpret_cast<int *>(base + (size_t) 12 * cap);
    const uint32_t J = a.jobsPerPair;
    for (;;) {
        uint32_t job = 0;
        if (lane == 0) job = atomicAdd(a.counter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= a.nPairs * J) break;
        const uint32_t slot = job / J;
        const int jx = (int) (job - slot * J);
        float *gDi = a.scratch + (size_t) slot * (cap + 8);
        int *gMeta = reinterpret_cast<int *>(gDi + cap);      // [0] best score of search 1 (float bits), [1] search 2, [2] NaN seen, [3] jobs done
        const SsPair p = a.pairs[slot];
        // ---- aligned columns from the backtrace (constructAlignHashes / invmap): iAli = query index, kAli = target index ----
        int n = 0;
        {
            int qi = p.qStart, ti = p.tStart;
            const char *bt = a.bt + p.btOff;
            for (uint32_t b0 = 0; b0 < p.btLen; b0 += 32) {
                const uint32_t b = b0 + lane;
                const char c = b < p.btLen ? bt[b] : 0;
                const unsigned mM = __ballot_sync(0xffffffffu, c == 'M'), mI = __ballot_sync(0xffffffffu, c == 'I'), mD = __ballot_sync(0xffffffffu, c == 'D');
                const unsigned lower = (1u << lane) - 1u;
                if (c == 'M') {
                    const int k = n + __popc(mM & lower);
                    w.iAli[k] = qi + __popc((mM | mI) & lower);
                    w.kAli[k] = ti + __popc((mM | mD) & lower);
                }
                n += __popc(mM); qi += __popc(mM | mI); ti += __popc(mM | mD);
            }
        }
        __syncwarp();
        w.n = n;
        for (int k = lane; k < n; k += 32) {
            const uint64_t qo = p.qCaOff + (uint64_t) w.iAli[k], to = p.tCaOff + (uint64_t) w.kAli[k];
            w.ytm[0][k] = a.qx[qo]; w.ytm[1][k] = a.qy[qo]; w.ytm[2][k] = a.qz[qo];
            w.xtm[0][k] = a.tx[to]; w.xtm[1][k] = a.ty[to]; w.xtm[2][k] = a.tz[to];
            w.term[k] = a.qNorm[p.qNormOff + (uint64_t) w.iAli[k]];
            w.di[k] = 0.0f;
        }
        __syncwarp();
        // ---- LDDT: every unordered pair of aligned columns whose query distance is below the cutoff (LDDT.cpp:158-215) ----
        // the per-column sums are multiples of 0.25, exact in float in any order; this job takes every J-th row, di[] collects its part
        for (int aIdx = jx; aIdx < n; aIdx += (int) J) {
            const float ax = w.ytm[0][aIdx], ay = w.ytm[1][aIdx], az = w.ytm[2][aIdx];
            const float bx = w.xtm[0][aIdx], by = w.xtm[1][aIdx], bz = w.xtm[2][aIdx];
            float mine = 0.0f;
            for (int b = aIdx + 1 + lane; b < n; b += 32) {
                const float dq = lddt_dist(ax, ay, az, w.ytm[0][b], w.ytm[1][b], w.ytm[2][b]);
                if (!(dq < 15.0f)) continue;
                const float dt = lddt_dist(bx, by, bz, w.xtm[0][b], w.xtm[1][b], w.xtm[2][b]);
                const float dl = fabsf(__fsub_rn(dq, dt));
                const float s = 0.25f * (float) ((dl < 0.5f) + (dl < 1.0f) + (dl < 2.0f) + (dl < 4.0f));
                mine += s;
                w.di[b] += s;         // column b belongs to this lane only within this pass (b = aIdx + 1 + lane + 32 m)
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            __syncwarp();
            if (lane == 0) w.di[aIdx] += mine;
            __syncwarp();
        }
        for (int k = lane; k < n; k += 32) { const float v = w.di[k]; if (v != 0.0f) atomicAdd(gDi + k, v); }
        // ---- TM-score: this job's share of the independent searches ----
        bool nanSeen = false;
        const float s1 = w.search(p.d0Std, p.scoreD8, p.d0Std, nanSeen, jx, (int) J);                 // standard_TMscore: local_d0_search = d0_input = d0
        const float s2 = w.search(p.d0SearchClipped, p.scoreD8, p.d0Search, nanSeen, jx, (int) J);   // detailed_search_standard
        int done = 0;
        if (lane == 0) {
            if (s1 >= 0.0f) atomicMax(gMeta + 0, __float_as_int(s1));    // non-negative floats order like their bit patterns
            if (s2 >= 0.0f) atomicMax(gMeta + 1, __float_as_int(s2));
            if (nanSeen) atomicOr(gMeta + 2, 1);
            __threadfence();
            done = atomicAdd(gMeta + 3, 1);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == (int) J - 1) {          // the last job of the pair finishes it
            __threadfence();
            // the searches used iAli / term as scratch: rebuild the query index of every aligned column for the LDDT normalisation
            {
                int n2 = 0, qi = p.qStart;
                const char *bt = a.bt + p.btOff;
                for (uint32_t b0 = 0; b0 < p.btLen; b0 += 32) {
                    const uint32_t b = b0 + lane;
                    const char c = b < p.btLen ? bt[b] : 0;
                    const unsigned mM = __ballot_sync(0xffffffffu, c == 'M'), mI = __ballot_sync(0xffffffffu, c == 'I');
                    const unsigned lower = (1u << lane) - 1u;
                    if (c == 'M') w.term[n2 + __popc(mM & lower)] = a.qNorm[p.qNormOff + (uint64_t) (qi + __popc((mM | mI) & lower))];
                    n2 += __popc(mM); qi += __popc(mM | mI);
                }
                __syncwarp();
            }
            if (lane == 0) {                // LDDTScoreResult (LDDT.h:102-119): sequential float sum, NaN columns dropped
                float sum = 0.0f; int len = n;
                for (int k = 0; k < n; k++) {
                    const float v = __fmul_rn(__ldcg(gDi + k), w.term[k]);
                    if (isnan(v)) len--; else sum = __fadd_rn(sum, v);
                }
                const float lddt = __fdiv_rn(sum, (float) len);
                const float b1 = __int_as_float(__ldcg(gMeta + 0)), b2raw = __int_as_float(__ldcg(gMeta + 1));
                const double tmAln = (double) b1 * (double) n / (1.0 * (double) p.Lnorm);
                const float b2 = __fdiv_rn(__fmul_rn(b2raw, (float) n), p.Lnorm);
                SsOut o; o.tm = fmax((double) b2, tmAln); o.lddt = lddt; o.flags = __ldcg(gMeta + 2) ? 1 : 0; a.out[slot] = o;
            }
        }
        __syncwarp();
    }
}

}  // namespace ss
