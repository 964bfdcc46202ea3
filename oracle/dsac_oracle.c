/*
 * TEST INFRASTRUCTURE ONLY -- tier-2 CPU oracle / CPU baseline for the DSAC* RGB forward path.
 *
 * Plain C99 + OpenMP restatement of /root/reference/dsacstar/dsacstar.cpp:63-178 and the
 * helpers it calls in /root/reference/dsacstar/dsacstar_util.h.  Nothing in the product
 * (crossloc_b200/, dsacstar/, networks/, loss/) links or loads this file; it is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference has no tests or golden vectors for this path and its own
 * extension needs OpenCV 3.4.2 C++ (absent here).  The OpenCV routines it calls are
 * restated from their published algorithms and validated numerically against the tier-1
 * oracle (oracle/dsac_oracle_py.py, which calls cv2 4.13) in tests/test_oracle.py:
 *   cv::projectPoints  -> ora_project        (x' = R X + t; z = z ? 1/z : 1; u = f x' z + cx)
 *   cv::Rodrigues      -> ora_rodrigues
 *   cv::solvePnP P3P   -> ora_p3p            (3-point law-of-cosines quartic, 4th point picks the root)
 *   cv::solvePnP ITERATIVE(useExtrinsicGuess) -> ora_lm  (CvLevMarq state machine, 20 iterations max,
 *                                               eps = FLT_EPSILON, lambda = 10^k starting at k = -3)
 *
 * Reference function            file:line                    here
 *   createSampling              dsacstar_util.h:59-76        cell_px
 *   sampleHypotheses            dsacstar_util.h:135-221      sample_hypothesis
 *   getReproErrs                dsacstar_util.h:356-446      repro_errs
 *   getHypScores                dsacstar_util.h:316-343      hyp_score
 *   softMax / draw(false)       dsacstar_util.h:684-752      select_best
 *   refineHyp                   dsacstar_util.h:522-597      refine_hyp
 *   pose2trans                  dsacstar_util.h:759-770      pose2trans
 *   irand / ThreadRand          thread_rand.cpp:13-71        replaced by Philox4x32-10 (crossloc_b200/rng.py)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_MAX_REF_STEPS 100 /* dsacstar.cpp:47 */
#define ORA_EPS 0.00000001    /* dsacstar_util.h:45 */

/* ------------------------------------------------------------------ RNG (crossloc_b200/rng.py) */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void ora_sample_cells(uint64_t seed, uint32_t image, uint32_t hyp, uint32_t tr, int w, int h, int32_t cells[8])
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[8];
    uint32_t c0[4] = {tr, hyp, image, 0}, c1[4] = {tr, hyp, image, 1};
    philox4x32_10(c0, key, r);
    philox4x32_10(c1, key, r + 4);
    for (int j = 0; j < 4; j++) {
        cells[2 * j] = (int32_t)(((uint64_t)r[2 * j] * (uint32_t)w) >> 32);
        cells[2 * j + 1] = (int32_t)(((uint64_t)r[2 * j + 1] * (uint32_t)h) >> 32);
    }
}

/* ------------------------------------------------------------------ small linear algebra */
void ora_rodrigues(const double r[3], double R[9])
{
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) {
        R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
        return;
    }
    double c = cos(th), s = sin(th), c1 = 1 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

/* rotation matrix -> axis-angle (cv::Rodrigues mat->vec for a proper rotation) */
static void rot2rvec(const double R[9], double r[3])
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1 ? 1 : (c < -1 ? -1 : c);
    double th = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1.0 : 1.0);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1.0 : 1.0);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        th /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * th; r[1] = ry * th; r[2] = rz * th;
        return;
    }
    double vth = 1 / (2 * s) * th;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

/* cv::projectPoints for one point, no distortion */
static void ora_project(const double R[9], const double t[3], double f, double cx, double cy,
                        double X, double Y, double Z, double *u, double *v)
{
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    *u = f * (x * z) + cx;
    *v = f * (y * z) + cy;
}

/* ------------------------------------------------------------------ quartic */
static double cbrt_signed(double x) { return x < 0 ? -pow(-x, 1.0 / 3) : pow(x, 1.0 / 3); }

/* one real root (the largest) of y^3 + p y^2 + q y + r = 0 */
static double cubic_largest_root(double p, double q, double r)
{
    double a = q - p * p / 3, b = r + 2 * p * p * p / 27 - p * q / 3, sh = -p / 3;
    double disc = b * b / 4 + a * a * a / 27;
    if (disc > 0) {
        double sd = sqrt(disc);
        return cbrt_signed(-b / 2 + sd) + cbrt_signed(-b / 2 - sd) + sh;
    }
    if (a >= 0) return sh; /* triple root */
    double m = 2 * sqrt(-a / 3), arg = 3 * b / (a * m);
    arg = arg > 1 ? 1 : (arg < -1 ? -1 : arg);
    return m * cos(acos(arg) / 3) + sh; /* k = 0 branch is the largest */
}

/* Newton (Bairstow) refinement of a quadratic factor x^2 + p x + q of the monic quartic
 * x^4 + a x^3 + b x^2 + c x + d: drives the remainder r x + s of the division to zero.  Keeps close
 * root pairs accurate where Ferrari's closed form loses half of the digits.                        */
static void bairstow_refine(double a, double b, double c, double d, double *p, double *q)
{
    for (int it = 0; it < 4; it++) {
        double al = a - *p, be = b - *p * al - *q;
        double r = c - *p * be - *q * al, s = d - *q * be;
        double rp = -be - *p * (*p - al) + *q, rq = *p - al;
        double sp = -*q * (*p - al), sq = *q - be;
        double det = rp * sq - rq * sp;
        if (!(fabs(det) > 1e-300)) return;
        double dp = (-r * sq + s * rq) / det, dq = (-s * rp + r * sp) / det;
        if (!(dp == dp) || !(dq == dq)) return;
        *p += dp;
        *q += dq;
    }
}

/* real roots of c4 x^4 + ... + c0: Ferrari's factorisation into two quadratics via the resolvent
 * cubic, each factor refined by Bairstow, then one Newton polish per simple root                   */
static int quartic_real_roots(const double c[5], double roots[4])
{
    if (fabs(c[4]) < 1e-300) return 0;
    double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    double y = cubic_largest_root(-b, a * cc - 4 * d, -a * a * d + 4 * b * d - cc * cc);
    double R2 = a * a / 4 - b + y, tol = 1e-9 * (fabs(a * a / 4) + fabs(b) + fabs(y) + 1e-300);
    if (R2 < -tol) return 0;
    double Rr = R2 > 0 ? sqrt(R2) : 0, D2, E2;
    if (Rr > sqrt(tol)) {
        double w = (4 * a * b - 8 * cc - a * a * a) / (4 * Rr);
        D2 = 3 * a * a / 4 - R2 - 2 * b + w;
        E2 = 3 * a * a / 4 - R2 - 2 * b - w;
    } else {
        double w = y * y - 4 * d;
        w = w > 0 ? 2 * sqrt(w) : 0;
        D2 = 3 * a * a / 4 - 2 * b + w;
        E2 = 3 * a * a / 4 - 2 * b - w;
    }
    /* quadratic factors x^2 + p x + q holding the root pairs -a/4 +- R/2 +- sqrt(D2 | E2)/2 */
    double m1 = -a / 4 + Rr / 2, m2 = -a / 4 - Rr / 2;
    double pq[2][2] = {{-2 * m1, m1 * m1 - D2 / 4}, {-2 * m2, m2 * m2 - E2 / 4}};
    int n = 0;
    for (int k = 0; k < 2; k++) {
        double p = pq[k][0], q = pq[k][1];
        bairstow_refine(a, b, cc, d, &p, &q);
        double disc = p * p - 4 * q, dtol = 1e-10 * (p * p + fabs(4 * q) + 1e-300);
        if (disc < -dtol) continue;
        double sd = disc > 0 ? sqrt(disc) : 0;
        /* numerically stable quadratic roots */
        double t = -0.5 * (p + (p >= 0 ? sd : -sd));
        double x1 = t, x2 = (t != 0) ? q / t : -p - t;
        if (sd == 0) x1 = x2 = -p / 2;
        roots[n++] = x1;
        roots[n++] = x2;
    }
    for (int i = 0; i < n; i++) {
        double x = roots[i];
        for (int it = 0; it < 2; it++) {
            double fx = (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0];
            double dfx = ((4 * c[4] * x + 3 * c[3]) * x + 2 * c[2]) * x + c[1];
            double scale = fabs(c[4] * x * x * x) + fabs(c[3] * x * x) + fabs(c[2] * x) + fabs(c[1]);
            if (!(fabs(dfx) > 1e-7 * scale)) break; /* (near-)multiple root: keep the Bairstow value */
            double nx = x - fx / dfx;
            if (!(nx == nx)) break;
            x = nx;
        }
        roots[i] = x;
    }
    return n;
}

/* ------------------------------------------------------------------ P3P
 * Restates the contract of cv::solvePnP(..., SOLVEPNP_P3P) as used at dsacstar_util.h:185-193:
 * exactly four correspondences; the first three give up to four poses, the fourth selects the
 * one with the smallest squared reprojection error; returns 0 when the 3-point problem has no
 * admissible solution.  obj: 4x3 (float inputs widened), img: 4x2 pixels.                      */
int ora_p3p(const double obj[12], const double img[8], double f, double cx, double cy, double rvec[3], double tvec[3])
{
    double fv[3][3];
    for (int i = 0; i < 3; i++) {
        double x = (img[2 * i] - cx) / f, y = (img[2 * i + 1] - cy) / f, n = sqrt(x * x + y * y + 1);
        fv[i][0] = x / n; fv[i][1] = y / n; fv[i][2] = 1 / n;
    }
    const double *X1 = obj, *X2 = obj + 3, *X3 = obj + 6;
    double d23[3] = {X2[0] - X3[0], X2[1] - X3[1], X2[2] - X3[2]};
    double d13[3] = {X1[0] - X3[0], X1[1] - X3[1], X1[2] - X3[2]};
    double d12[3] = {X1[0] - X2[0], X1[1] - X2[1], X1[2] - X2[2]};
    double a2 = d23[0] * d23[0] + d23[1] * d23[1] + d23[2] * d23[2];
    double b2 = d13[0] * d13[0] + d13[1] * d13[1] + d13[2] * d13[2];
    double c2 = d12[0] * d12[0] + d12[1] * d12[1] + d12[2] * d12[2];
    if (!(a2 > 0) || !(b2 > 0) || !(c2 > 0)) return 0;
    double ca = fv[1][0] * fv[2][0] + fv[1][1] * fv[2][1] + fv[1][2] * fv[2][2];
    double cb = fv[0][0] * fv[2][0] + fv[0][1] * fv[2][1] + fv[0][2] * fv[2][2];
    double cg = fv[0][0] * fv[1][0] + fv[0][1] * fv[1][1] + fv[0][2] * fv[1][2];

    /* s2 = u s1, s3 = v s1;  u = N(v) / D(v) from the difference of the two cosine-law ratios */
    double k = (a2 - c2) / b2;
    double N[3] = {k + 1, -2 * k * cb, k - 1};      /* n0 + n1 v + n2 v^2 */
    double D[2] = {2 * cg, -2 * ca};                /* d0 + d1 v */
    /* b2 (D^2 + N^2 - 2 cg N D) - c2 (1 + v^2 - 2 cb v) D^2 = 0 */
    double DD[3] = {D[0] * D[0], 2 * D[0] * D[1], D[1] * D[1]};
    double NN[5] = {N[0] * N[0], 2 * N[0] * N[1], N[1] * N[1] + 2 * N[0] * N[2], 2 * N[1] * N[2], N[2] * N[2]};
    double ND[4] = {N[0] * D[0], N[0] * D[1] + N[1] * D[0], N[1] * D[1] + N[2] * D[0], N[2] * D[1]};
    double Q[3] = {1, -2 * cb, 1};
    double QD[5] = {Q[0] * DD[0], Q[0] * DD[1] + Q[1] * DD[0], Q[0] * DD[2] + Q[1] * DD[1] + Q[2] * DD[0],
                    Q[1] * DD[2] + Q[2] * DD[1], Q[2] * DD[2]};
    double r = c2 / b2, poly[5];
    for (int i = 0; i < 5; i++) {
        double dd = i < 3 ? DD[i] : 0, nd = i < 4 ? ND[i] : 0;
        poly[i] = dd + NN[i] - 2 * cg * nd - r * QD[i];
    }
    double roots[4];
    int nr = quartic_real_roots(poly, roots);

    int found = 0;
    double best = 0;
    for (int i = 0; i < nr; i++) {
        double v = roots[i];
        if (!(v > 0)) continue;
        double den = D[0] + D[1] * v;
        if (fabs(den) < 1e-12) continue;
        double u = (N[0] + N[1] * v + N[2] * v * v) / den;
        if (!(u > 0)) continue;
        double q = 1 + v * v - 2 * v * cb;
        if (!(q > 0)) continue;
        double s1 = sqrt(b2 / q), s2 = u * s1, s3 = v * s1;
        /* Newton polish of the three cosine-law equations in the depths themselves */
        for (int it = 0; it < 2; it++) {
            double F1 = s2 * s2 + s3 * s3 - 2 * s2 * s3 * ca - a2;
            double F2 = s1 * s1 + s3 * s3 - 2 * s1 * s3 * cb - b2;
            double F3 = s1 * s1 + s2 * s2 - 2 * s1 * s2 * cg - c2;
            double J12 = 2 * s2 - 2 * s3 * ca, J13 = 2 * s3 - 2 * s2 * ca;
            double J21 = 2 * s1 - 2 * s3 * cb, J23 = 2 * s3 - 2 * s1 * cb;
            double J31 = 2 * s1 - 2 * s2 * cg, J32 = 2 * s2 - 2 * s1 * cg;
            /* J = [[0,J12,J13],[J21,0,J23],[J31,J32,0]] */
            double det = J12 * J23 * J31 + J13 * J21 * J32;
            if (!(fabs(det) > 1e-12 * (fabs(J12 * J23 * J31) + fabs(J13 * J21 * J32)) + 1e-300)) break;
            double d1 = (F1 * (-J23 * J32) - J12 * (F2 * 0 - J23 * F3) + J13 * (F2 * J32 - 0 * F3)) / det;
            double d2 = (0 * (F2 * 0 - J23 * F3) - F1 * (J21 * 0 - J23 * J31) + J13 * (J21 * F3 - F2 * J31)) / det;
            double d3 = (0 * (0 * F3 - F2 * J32) - J12 * (J21 * F3 - F2 * J31) + F1 * (J21 * J32 - 0 * J31)) / det;
            if (!(d1 == d1) || !(d2 == d2) || !(d3 == d3)) break;
            s1 -= d1; s2 -= d2; s3 -= d3;
        }
        if (!(s1 > 0) || !(s2 > 0) || !(s3 > 0)) continue;
        double P1[3], P2[3], P3[3];
        for (int j = 0; j < 3; j++) { P1[j] = s1 * fv[0][j]; P2[j] = s2 * fv[1][j]; P3[j] = s3 * fv[2][j]; }
        /* rigid transform from two congruent triangles via orthonormal frames */
        double e1[3], e2[3], e3[3], g1[3], g2[3], g3[3], tmp[3];
        for (int j = 0; j < 3; j++) { e1[j] = X2[j] - X1[j]; tmp[j] = X3[j] - X1[j]; }
        cross3(e1, tmp, e3);
        double n1 = norm3(e1), n3 = norm3(e3);
        if (n1 < 1e-300 || n3 < 1e-300) continue;
        for (int j = 0; j < 3; j++) { e1[j] /= n1; e3[j] /= n3; }
        cross3(e3, e1, e2);
        for (int j = 0; j < 3; j++) { g1[j] = P2[j] - P1[j]; tmp[j] = P3[j] - P1[j]; }
        cross3(g1, tmp, g3);
        n1 = norm3(g1); n3 = norm3(g3);
        if (n1 < 1e-300 || n3 < 1e-300) continue;
        for (int j = 0; j < 3; j++) { g1[j] /= n1; g3[j] /= n3; }
        cross3(g3, g1, g2);
        double Rm[9], t[3];
        for (int ri = 0; ri < 3; ri++)
            for (int cj = 0; cj < 3; cj++) Rm[3 * ri + cj] = g1[ri] * e1[cj] + g2[ri] * e2[cj] + g3[ri] * e3[cj];
        for (int j = 0; j < 3; j++) t[j] = P1[j] - (Rm[3 * j] * X1[0] + Rm[3 * j + 1] * X1[1] + Rm[3 * j + 2] * X1[2]);
        /* 4th point disambiguation */
        const double *X4 = obj + 9;
        double x = Rm[0] * X4[0] + Rm[1] * X4[1] + Rm[2] * X4[2] + t[0];
        double y = Rm[3] * X4[0] + Rm[4] * X4[1] + Rm[5] * X4[2] + t[1];
        double z = Rm[6] * X4[0] + Rm[7] * X4[1] + Rm[8] * X4[2] + t[2];
        double du = cx + f * x / z - img[6], dv = cy + f * y / z - img[7];
        double e = du * du + dv * dv;
        if (!(e == e)) continue;
        if (!found || e < best) {
            found = 1; best = e;
            rot2rvec(Rm, rvec);
            tvec[0] = t[0]; tvec[1] = t[1]; tvec[2] = t[2];
        }
    }
    return found;
}

/* ------------------------------------------------------------------ LM (solvePnP ITERATIVE with guess) */
static int solve6(double A[36], double b[6], double x[6])
{
    int p[6];
    for (int i = 0; i < 6; i++) p[i] = i;
    for (int c = 0; c < 6; c++) {
        int best = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(A[6 * p[r] + c]) > fabs(A[6 * p[best] + c])) best = r;
        int tmp = p[c]; p[c] = p[best]; p[best] = tmp;
        double piv = A[6 * p[c] + c];
        if (fabs(piv) < 1e-300) return 0;
        for (int r = c + 1; r < 6; r++) {
            double m = A[6 * p[r] + c] / piv;
            if (m == 0) continue;
            for (int k2 = c; k2 < 6; k2++) A[6 * p[r] + k2] -= m * A[6 * p[c] + k2];
            b[p[r]] -= m * b[p[c]];
        }
    }
    for (int c = 5; c >= 0; c--) {
        double s = b[p[c]];
        for (int k2 = c + 1; k2 < 6; k2++) s -= A[6 * p[c] + k2] * x[k2];
        x[c] = s / A[6 * p[c] + c];
    }
    return 1;
}

/* residuals (+ optional J^T J, J^T e) of the pixel reprojection error over n points */
static double lm_accumulate(const double prm[6], int n, const float *obj, const float *img, double f, double cx,
                            double cy, double *JtJ, double *Jte)
{
    double R[9], r[3] = {prm[0], prm[1], prm[2]}, t[3] = {prm[3], prm[4], prm[5]};
    ora_rodrigues(r, R);
    double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double M[9]; /* d(R X)/dr = -R [X]x M  with  M = (r r^T + (R^T - I)[r]x) / |r|^2, identity for r -> 0 */
    if (th2 < 1e-24) {
        M[0] = M[4] = M[8] = 1; M[1] = M[2] = M[3] = M[5] = M[6] = M[7] = 0;
    } else {
        double A[9] = {R[0] - 1, R[3], R[6], R[1], R[4] - 1, R[7], R[2], R[5], R[8] - 1}; /* R^T - I */
        double K[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = r[i] * r[j];
                for (int k2 = 0; k2 < 3; k2++) s += A[3 * i + k2] * K[3 * k2 + j];
                M[3 * i + j] = s / th2;
            }
    }
    if (JtJ) { memset(JtJ, 0, 36 * sizeof(double)); memset(Jte, 0, 6 * sizeof(double)); }
    double sq = 0;
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
        double iz = z ? 1. / z : 1;
        double eu = f * x * iz + cx - img[2 * i], ev = f * y * iz + cy - img[2 * i + 1];
        sq += eu * eu + ev * ev;
        if (!JtJ) continue;
        /* d(u,v)/dp */
        double a0 = f * iz, a2 = -f * x * iz * iz, b2 = -f * y * iz * iz;
        /* dp/dr = -R [X]x M */
        double Xx[9] = {0, -Z, Y, Z, 0, -X, -Y, X, 0}, RX[9], dpdr[9];
        for (int ii = 0; ii < 3; ii++)
            for (int jj = 0; jj < 3; jj++)
                RX[3 * ii + jj] = R[3 * ii] * Xx[jj] + R[3 * ii + 1] * Xx[3 + jj] + R[3 * ii + 2] * Xx[6 + jj];
        for (int ii = 0; ii < 3; ii++)
            for (int jj = 0; jj < 3; jj++)
                dpdr[3 * ii + jj] = -(RX[3 * ii] * M[jj] + RX[3 * ii + 1] * M[3 + jj] + RX[3 * ii + 2] * M[6 + jj]);
        double Ju[6], Jv[6];
        for (int j = 0; j < 3; j++) {
            Ju[j] = a0 * dpdr[j] + a2 * dpdr[6 + j];
            Jv[j] = a0 * dpdr[3 + j] + b2 * dpdr[6 + j];
        }
        Ju[3] = a0; Ju[4] = 0; Ju[5] = a2;
        Jv[3] = 0; Jv[4] = a0; Jv[5] = b2;
        for (int r2 = 0; r2 < 6; r2++) {
            for (int c = 0; c < 6; c++) JtJ[6 * r2 + c] += Ju[r2] * Ju[c] + Jv[r2] * Jv[c];
            Jte[r2] += Ju[r2] * eu + Jv[r2] * ev;
        }
    }
    return sqrt(sq);
}

static int lm_step(const double JtJ[36], const double Jte[6], int lambdaLg10, const double prev[6], double prm[6])
{
    double A[36], b[6], x[6], lambda = exp(lambdaLg10 * log(10.));
    memcpy(A, JtJ, sizeof(A));
    memcpy(b, Jte, sizeof(b));
    for (int i = 0; i < 6; i++) A[7 * i] *= 1. + lambda;
    if (!solve6(A, b, x)) return 0;
    for (int i = 0; i < 6; i++) prm[i] = prev[i] - x[i];
    return 1;
}

/* CvLevMarq::update driven as in cvFindExtrinsicCameraParams2 (max_iter 20, eps FLT_EPSILON) */
int ora_lm(int n, const float *obj, const float *img, double f, double cx, double cy, double rvec[3], double tvec[3])
{
    double prm[6] = {rvec[0], rvec[1], rvec[2], tvec[0], tvec[1], tvec[2]}, prev[6], JtJ[36], Jte[6];
    int lambdaLg10 = -3, iters = 0;
    double prevErr = DBL_MAX, errNorm;
    for (;;) {
        /* CALC_J: J and err at prm */
        double e0 = lm_accumulate(prm, n, obj, img, f, cx, cy, JtJ, Jte);
        memcpy(prev, prm, sizeof(prev));
        if (!lm_step(JtJ, Jte, lambdaLg10, prev, prm)) return 0;
        if (iters == 0) prevErr = e0;
        /* CHECK_ERR loop */
        for (;;) {
            errNorm = lm_accumulate(prm, n, obj, img, f, cx, cy, NULL, NULL);
            if (errNorm > prevErr && ++lambdaLg10 <= 16) {
                if (!lm_step(JtJ, Jte, lambdaLg10, prev, prm)) return 0;
                continue;
            }
            break;
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double dn = 0, pn = 0;
        for (int i = 0; i < 6; i++) { dn += (prm[i] - prev[i]) * (prm[i] - prev[i]); pn += prev[i] * prev[i]; }
        if (++iters >= 20 || sqrt(dn) / (sqrt(pn) + DBL_EPSILON) < FLT_EPSILON) break;
        prevErr = errNorm;
    }
    for (int i = 0; i < 3; i++) { rvec[i] = prm[i]; tvec[i] = prm[3 + i]; }
    return 1;
}

/* ------------------------------------------------------------------ DSAC* pieces */
static inline void cell_px(int x, int y, int S, int *px, int *py)
{
    *px = x * S + S / 2; /* dsacstar_util.h:70-72 */
    *py = y * S + S / 2;
}

/* dsacstar_util.h:356-446, calcJ = false; errs is [Hc*Wc] row-major */
static void repro_errs(const float *coords, int Hc, int Wc, int S, const double rvec[3], const double tvec[3], float f,
                       float cx, float cy, float maxReproj, float *errs)
{
    double R[9];
    ora_rodrigues(rvec, R);
    int n = Hc * Wc;
    for (int y = 0; y < Hc; y++)
        for (int x = 0; x < Wc; x++) {
            int i = y * Wc + x, px, py;
            cell_px(x, y, S, &px, &py);
            double u, v;
            ora_project(R, tvec, f, cx, cy, coords[i], coords[n + i], coords[2 * n + i], &u, &v);
            float du = (float)px - (float)u, dv = (float)py - (float)v; /* Point2f - Point2f */
            float l = (float)sqrt((double)du * du + (double)dv * dv);   /* (float) cv::norm */
            errs[i] = l < maxReproj ? l : maxReproj;
        }
}

/* dsacstar_util.h:316-343 for one hypothesis */
static double hyp_score(const float *errs, int Hc, int Wc, float thr, float alpha)
{
    float beta = 5 / thr;
    double s = 0;
    for (int x = 0; x < Wc; x++)
        for (int y = 0; y < Hc; y++) {
            double soft = beta * (errs[y * Wc + x] - thr);
            soft = 1 / (1 + exp(-soft));
            s += 1 - soft;
        }
    return s * (alpha / Wc / Hc);
}

/* dsacstar_util.h:135-221 for one hypothesis; returns number of tries used */
static unsigned sample_hypothesis(const float *coords, int Hc, int Wc, int S, float f, float cx, float cy, float thr,
                                  unsigned maxTries, uint64_t seed, uint32_t image, uint32_t h,
                                  const int32_t *forced, double rvec[3], double tvec[3], int32_t cells[8])
{
    int n = Hc * Wc;
    unsigned t = 0;
    rvec[0] = rvec[1] = rvec[2] = tvec[0] = tvec[1] = tvec[2] = 0;
    while (t < maxTries) {
        if (forced) memcpy(cells, forced, 8 * sizeof(int32_t));
        else ora_sample_cells(seed, image, h, t, Wc, Hc, cells);
        t++;
        double obj[12], img[8];
        for (int j = 0; j < 4; j++) {
            int x = cells[2 * j], y = cells[2 * j + 1], px, py;
            cell_px(x, y, S, &px, &py);
            img[2 * j] = px; img[2 * j + 1] = py;
            obj[3 * j] = coords[y * Wc + x]; obj[3 * j + 1] = coords[n + y * Wc + x]; obj[3 * j + 2] = coords[2 * n + y * Wc + x];
        }
        if (!ora_p3p(obj, img, f, cx, cy, rvec, tvec)) {
            rvec[0] = rvec[1] = rvec[2] = tvec[0] = tvec[1] = tvec[2] = 0; /* safeSolvePnP, dsacstar_util.h:114-116 */
            if (forced) break;
            continue;
        }
        double R[9];
        ora_rodrigues(rvec, R);
        int outlier = 0;
        for (int j = 0; j < 4; j++) {
            double u, v;
            ora_project(R, tvec, f, cx, cy, obj[3 * j], obj[3 * j + 1], obj[3 * j + 2], &u, &v);
            float du = (float)img[2 * j] - (float)u, dv = (float)img[2 * j + 1] - (float)v;
            if (sqrt((double)du * du + (double)dv * dv) < thr) continue; /* strict <, dsacstar_util.h:210 */
            outlier = 1;
            break;
        }
        if (!outlier || forced) break;
    }
    return t;
}

/* dsacstar_util.h:522-597 */
static int refine_hyp(const float *coords, int Hc, int Wc, int S, float f, float cx, float cy, float thr,
                      float maxReproj, const float *errs0, double rvec[3], double tvec[3], int32_t *counts)
{
    int n = Hc * Wc, steps = 0;
    float *errs = (float *)malloc(n * sizeof(float)), *obj = (float *)malloc(3 * n * sizeof(float)),
          *img = (float *)malloc(2 * n * sizeof(float));
    memcpy(errs, errs0, n * sizeof(float));
    unsigned best = 4;
    for (int step = 0; step < ORA_MAX_REF_STEPS; step++) {
        unsigned m = 0;
        for (int x = 0; x < Wc; x++)
            for (int y = 0; y < Hc; y++)
                if (errs[y * Wc + x] < thr) {
                    int px, py, i = y * Wc + x;
                    cell_px(x, y, S, &px, &py);
                    img[2 * m] = (float)px; img[2 * m + 1] = (float)py;
                    obj[3 * m] = coords[i]; obj[3 * m + 1] = coords[n + i]; obj[3 * m + 2] = coords[2 * n + i];
                    m++;
                }
        if (counts) counts[steps] = (int32_t)m;
        steps++;
        if (m <= best) break;
        best = m;
        double r2[3] = {rvec[0], rvec[1], rvec[2]}, t2[3] = {tvec[0], tvec[1], tvec[2]};
        if (!ora_lm((int)m, obj, img, f, cx, cy, r2, t2)) break;
        memcpy(rvec, r2, sizeof(r2));
        memcpy(tvec, t2, sizeof(t2));
        repro_errs(coords, Hc, Wc, S, rvec, tvec, f, cx, cy, maxReproj, errs);
    }
    free(errs); free(obj); free(img);
    return steps;
}

/* dsacstar_util.h:759-770: inverse of [R t; 0 1], written row-major as float */
static void pose2trans(const double rvec[3], const double tvec[3], float out[16])
{
    double R[9];
    ora_rodrigues(rvec, R);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) out[4 * i + j] = (float)R[3 * j + i];
        out[4 * i + 3] = (float)-(R[i] * tvec[0] + R[3 + i] * tvec[1] + R[6 + i] * tvec[2]);
    }
    out[12] = out[13] = out[14] = 0; out[15] = 1;
}

/*
 * dsacstar.cpp:63-178 for one image.
 *   coords          [3, Hc, Wc] planar float
 *   forced_samples  nullable [hyps, 4, 2] (x, y) cells: one try per hypothesis, no acceptance loop
 *   out_*           nullable debug outputs: best index, scores [hyps], hypotheses [hyps, 6] (rvec, tvec),
 *                   tries [hyps], refine inlier counts [ORA_MAX_REF_STEPS] (unused entries = -1),
 *                   refined rvec/tvec [6]
 */
int ora_forward_rgb(const float *coords, int Hc, int Wc, float *out_pose, int hyps, float thr, float focal, float cx,
                    float cy, float alpha, float maxReproj, int S, uint64_t seed, uint32_t image, unsigned maxTries,
                    const int32_t *forced_samples, int do_refine, int32_t *out_best, double *out_scores,
                    double *out_hyps, int32_t *out_tries, int32_t *out_counts, double *out_rt)
{
    int n = Hc * Wc;
    double *rv = (double *)malloc(sizeof(double) * 3 * hyps), *tv = (double *)malloc(sizeof(double) * 3 * hyps);
    double *scores = (double *)malloc(sizeof(double) * hyps);
    float *errs = (float *)malloc(sizeof(float) * (size_t)n * hyps);
#pragma omp parallel for schedule(dynamic)
    for (int h = 0; h < hyps; h++) {
        int32_t cells[8];
        unsigned t = sample_hypothesis(coords, Hc, Wc, S, focal, cx, cy, thr, maxTries, seed, image, (uint32_t)h,
                                       forced_samples ? forced_samples + 8 * h : NULL, rv + 3 * h, tv + 3 * h, cells);
        if (out_tries) out_tries[h] = (int32_t)t;
    }
#pragma omp parallel for schedule(dynamic)
    for (int h = 0; h < hyps; h++) {
        repro_errs(coords, Hc, Wc, S, rv + 3 * h, tv + 3 * h, focal, cx, cy, maxReproj, errs + (size_t)n * h);
        scores[h] = hyp_score(errs + (size_t)n * h, Hc, Wc, thr, alpha);
    }
    /* softMax + draw(false): first maximal probability among those >= EPS */
    double mx = scores[0], sum = 0;
    for (int h = 1; h < hyps; h++) if (scores[h] > mx) mx = scores[h];
    for (int h = 0; h < hyps; h++) sum += exp(scores[h] - mx);
    int best = 0;
    double bestp = -1;
    for (int h = 0; h < hyps; h++) {
        double p = exp(scores[h] - mx) / sum;
        if (p < ORA_EPS) continue;
        if (bestp < 0 || p > bestp) { bestp = p; best = h; }
    }
    double r[3] = {rv[3 * best], rv[3 * best + 1], rv[3 * best + 2]}, t[3] = {tv[3 * best], tv[3 * best + 1], tv[3 * best + 2]};
    if (out_counts) for (int i = 0; i < ORA_MAX_REF_STEPS; i++) out_counts[i] = -1;
    if (do_refine) refine_hyp(coords, Hc, Wc, S, focal, cx, cy, thr, maxReproj, errs + (size_t)n * best, r, t, out_counts);
    pose2trans(r, t, out_pose);
    if (out_best) *out_best = best;
    if (out_scores) memcpy(out_scores, scores, sizeof(double) * hyps);
    if (out_hyps) for (int h = 0; h < hyps; h++) for (int j = 0; j < 3; j++) { out_hyps[6 * h + j] = rv[3 * h + j]; out_hyps[6 * h + 3 + j] = tv[3 * h + j]; }
    if (out_rt) for (int j = 0; j < 3; j++) { out_rt[j] = r[j]; out_rt[3 + j] = t[j]; }
    free(rv); free(tv); free(scores); free(errs);
    return 0;
}

int ora_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
