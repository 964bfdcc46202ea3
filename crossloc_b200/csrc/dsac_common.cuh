// Device-side geometry for the DSAC* pose solver (sm_100a).
//
// Everything here is double precision on float inputs, as the reference's OpenCV calls are
// (SURVEY.md section 8a rows a9-a14).  Reference call sites restated:
//   cv::projectPoints   /root/reference/dsacstar/dsacstar_util.h:199-205, 395-401   -> project_point
//   cv::Rodrigues       /root/reference/dsacstar/dsacstar_util.h:762                -> rodrigues
//   cv::solvePnP (P3P)  /root/reference/dsacstar/dsacstar_util.h:185-193            -> p3p_solve
//   irand               /root/reference/dsacstar/thread_rand.cpp:32-42, 68-71       -> sample_cells (Philox4x32-10,
//                                                                                     spec in crossloc_b200/rng.py)
// HOSTDEV lets tests/host_math_test.cu compile the same functions for the CPU to check them
// against cv2 without a GPU; the shipped library only ever calls them from kernels.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define HOSTDEV __host__ __device__ __forceinline__
#else
#define HOSTDEV inline
#endif

namespace cl {

struct Pose {
    double r[3];  // axis-angle (OpenCV rvec), scene -> camera
    double t[3];  // translation (OpenCV tvec)
};

// ---------------------------------------------------------------------------- RNG
HOSTDEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0; c1 = (uint32_t)p1; c2 = n2; c3 = (uint32_t)p0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Four (x, y) cells of try `tr` of hypothesis `hyp` of image `image`; drawn with replacement,
// x and y independently (dsacstar_util.h:168-173).
HOSTDEV void sample_cells(uint64_t seed, uint32_t image, uint32_t hyp, uint32_t tr, int w, int h, int cells[8])
{
    uint32_t r[8];
    philox4x32_10(tr, hyp, image, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    philox4x32_10(tr, hyp, image, 1u, (uint32_t)seed, (uint32_t)(seed >> 32), r + 4);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cells[2 * j] = (int)(((uint64_t)r[2 * j] * (uint32_t)w) >> 32);
        cells[2 * j + 1] = (int)(((uint64_t)r[2 * j + 1] * (uint32_t)h) >> 32);
    }
}

// ---------------------------------------------------------------------------- rotations
HOSTDEV void rodrigues(const double r[3], double R[9])
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < DBL_EPSILON) {
        R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
        return;
    }
    const double c = cos(th), s = sin(th), c1 = 1 - c, x = r[0] / th, y = r[1] / th, z = r[2] / th;
    R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

HOSTDEV void rot_to_rvec(const double R[9], double r[3])
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
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
    const double vth = 1 / (2 * s) * th;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// cv::projectPoints without distortion: z = z ? 1/z : 1 (no cheirality test)
HOSTDEV void project_point(const double R[9], const double t[3], double f, double cx, double cy, double X, double Y,
                           double Z, double& u, double& v)
{
    const double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    const double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    u = f * (x * z) + cx;
    v = f * (y * z) + cy;
}

// Clamped float reprojection error of one cell, as getReproErrs computes it
// (dsacstar_util.h:438-443): Point2f difference, double norm, cast to float, min with maxReproj.
HOSTDEV float repro_error(const double R[9], const double t[3], float f, float cx, float cy, float X, float Y, float Z,
                          int px, int py, float max_reproj)
{
    double u, v;
    project_point(R, t, f, cx, cy, X, Y, Z, u, v);
    const float du = (float)px - (float)u, dv = (float)py - (float)v;
    const float l = (float)sqrt((double)du * du + (double)dv * dv);
    return l < max_reproj ? l : max_reproj;   // NaN propagates as max_reproj, like std::min(l, max) does
}

// ---------------------------------------------------------------------------- quartic
HOSTDEV double cubic_largest_root(double p, double q, double r)
{
    const double a = q - p * p / 3, b = r + 2 * p * p * p / 27 - p * q / 3, sh = -p / 3;
    const double disc = b * b / 4 + a * a * a / 27;
    if (disc > 0) {
        const double sd = sqrt(disc);
        return cbrt(-b / 2 + sd) + cbrt(-b / 2 - sd) + sh;
    }
    if (a >= 0) return sh;
    const double m = 2 * sqrt(-a / 3);
    double arg = 3 * b / (a * m);
    arg = arg > 1 ? 1 : (arg < -1 ? -1 : arg);
    return m * cos(acos(arg) / 3) + sh;
}

HOSTDEV void bairstow_refine(double a, double b, double c, double d, double& p, double& q)
{
    for (int it = 0; it < 4; it++) {
        const double al = a - p, be = b - p * al - q;
        const double r = c - p * be - q * al, s = d - q * be;
        const double rp = -be - p * (p - al) + q, rq = p - al;
        const double sp = -q * (p - al), sq = q - be;
        const double det = rp * sq - rq * sp;
        if (!(fabs(det) > 1e-300)) return;
        const double dp = (-r * sq + s * rq) / det, dq = (-s * rp + r * sp) / det;
        if (!(dp == dp) || !(dq == dq)) return;
        p += dp;
        q += dq;
    }
}

// Real roots of c[4] x^4 + ... + c[0]: Ferrari factorisation into two quadratics through the
// resolvent cubic, Bairstow refinement of each factor, one Newton polish per simple root.
HOSTDEV int quartic_real_roots(const double c[5], double roots[4])
{
    if (fabs(c[4]) < 1e-300) return 0;
    const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
    const double y = cubic_largest_root(-b, a * cc - 4 * d, -a * a * d + 4 * b * d - cc * cc);
    const double R2 = a * a / 4 - b + y, tol = 1e-9 * (fabs(a * a / 4) + fabs(b) + fabs(y) + 1e-300);
    if (R2 < -tol) return 0;
    const double Rr = R2 > 0 ? sqrt(R2) : 0;
    double D2, E2;
    if (Rr > sqrt(tol)) {
        const double w = (4 * a * b - 8 * cc - a * a * a) / (4 * Rr);
        D2 = 3 * a * a / 4 - R2 - 2 * b + w;
        E2 = 3 * a * a / 4 - R2 - 2 * b - w;
    } else {
        double w = y * y - 4 * d;
        w = w > 0 ? 2 * sqrt(w) : 0;
        D2 = 3 * a * a / 4 - 2 * b + w;
        E2 = 3 * a * a / 4 - 2 * b - w;
    }
    const double m1 = -a / 4 + Rr / 2, m2 = -a / 4 - Rr / 2;
    int n = 0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double p = k == 0 ? -2 * m1 : -2 * m2;
        double q = k == 0 ? m1 * m1 - D2 / 4 : m2 * m2 - E2 / 4;
        bairstow_refine(a, b, cc, d, p, q);
        const double disc = p * p - 4 * q, dtol = 1e-10 * (p * p + fabs(4 * q) + 1e-300);
        if (disc < -dtol) continue;
        const double sd = disc > 0 ? sqrt(disc) : 0;
        const double t = -0.5 * (p + (p >= 0 ? sd : -sd));
        double x1 = t, x2 = (t != 0) ? q / t : -p - t;
        if (sd == 0) x1 = x2 = -p / 2;
        roots[n++] = x1;
        roots[n++] = x2;
    }
    for (int i = 0; i < n; i++) {
        double x = roots[i];
        for (int it = 0; it < 2; it++) {
            const double fx = (((c[4] * x + c[3]) * x + c[2]) * x + c[1]) * x + c[0];
            const double dfx = ((4 * c[4] * x + 3 * c[3]) * x + 2 * c[2]) * x + c[1];
            const double scale = fabs(c[4] * x * x * x) + fabs(c[3] * x * x) + fabs(c[2] * x) + fabs(c[1]);
            if (!(fabs(dfx) > 1e-7 * scale)) break;
            const double nx = x - fx / dfx;
            if (!(nx == nx)) break;
            x = nx;
        }
        roots[i] = x;
    }
    return n;
}

HOSTDEV void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
HOSTDEV double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// ---------------------------------------------------------------------------- P3P
// Contract of cv::solvePnP(SOLVEPNP_P3P) on exactly four correspondences: the first three give up
// to four poses (law-of-cosines quartic in the depth ratio), the fourth picks the one with the
// smallest squared reprojection error; false when no admissible solution exists.
HOSTDEV bool p3p_solve(const double obj[12], const double img[8], double f, double cx, double cy, Pose& out)
{
    double fv[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double x = (img[2 * i] - cx) / f, y = (img[2 * i + 1] - cy) / f, n = sqrt(x * x + y * y + 1);
        fv[i][0] = x / n; fv[i][1] = y / n; fv[i][2] = 1 / n;
    }
    const double* X1 = obj; const double* X2 = obj + 3; const double* X3 = obj + 6; const double* X4 = obj + 9;
    double a2 = 0, b2 = 0, c2 = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        a2 += (X2[j] - X3[j]) * (X2[j] - X3[j]);
        b2 += (X1[j] - X3[j]) * (X1[j] - X3[j]);
        c2 += (X1[j] - X2[j]) * (X1[j] - X2[j]);
    }
    if (!(a2 > 0) || !(b2 > 0) || !(c2 > 0)) return false;
    const double ca = fv[1][0] * fv[2][0] + fv[1][1] * fv[2][1] + fv[1][2] * fv[2][2];
    const double cb = fv[0][0] * fv[2][0] + fv[0][1] * fv[2][1] + fv[0][2] * fv[2][2];
    const double cg = fv[0][0] * fv[1][0] + fv[0][1] * fv[1][1] + fv[0][2] * fv[1][2];

    // s2 = u s1, s3 = v s1, u = N(v) / D(v);  quartic: b2 (D^2 + N^2 - 2 cg N D) - c2 (1 + v^2 - 2 cb v) D^2
    const double k = (a2 - c2) / b2;
    const double N0 = k + 1, N1 = -2 * k * cb, N2 = k - 1, D0 = 2 * cg, D1 = -2 * ca;
    const double DD[3] = {D0 * D0, 2 * D0 * D1, D1 * D1};
    const double NN[5] = {N0 * N0, 2 * N0 * N1, N1 * N1 + 2 * N0 * N2, 2 * N1 * N2, N2 * N2};
    const double ND[4] = {N0 * D0, N0 * D1 + N1 * D0, N1 * D1 + N2 * D0, N2 * D1};
    const double Q1 = -2 * cb;
    const double QD[5] = {DD[0], DD[1] + Q1 * DD[0], DD[2] + Q1 * DD[1] + DD[0], Q1 * DD[2] + DD[1], DD[2]};
    const double rr = c2 / b2;
    double poly[5];
    poly[0] = DD[0] + NN[0] - 2 * cg * ND[0] - rr * QD[0];
    poly[1] = DD[1] + NN[1] - 2 * cg * ND[1] - rr * QD[1];
    poly[2] = DD[2] + NN[2] - 2 * cg * ND[2] - rr * QD[2];
    poly[3] = NN[3] - 2 * cg * ND[3] - rr * QD[3];
    poly[4] = NN[4] - rr * QD[4];
    double roots[4];
    const int nr = quartic_real_roots(poly, roots);

    bool found = false;
    double best = 0;
    for (int i = 0; i < nr; i++) {
        const double v = roots[i];
        if (!(v > 0)) continue;
        const double den = D0 + D1 * v;
        if (fabs(den) < 1e-12) continue;
        const double u = (N0 + N1 * v + N2 * v * v) / den;
        if (!(u > 0)) continue;
        const double q = 1 + v * v - 2 * v * cb;
        if (!(q > 0)) continue;
        double s1 = sqrt(b2 / q), s2 = u * s1, s3 = v * s1;
        for (int it = 0; it < 2; it++) {   // Newton polish of the cosine-law equations in the depths
            const double F1 = s2 * s2 + s3 * s3 - 2 * s2 * s3 * ca - a2;
            const double F2 = s1 * s1 + s3 * s3 - 2 * s1 * s3 * cb - b2;
            const double F3 = s1 * s1 + s2 * s2 - 2 * s1 * s2 * cg - c2;
            const double J12 = 2 * s2 - 2 * s3 * ca, J13 = 2 * s3 - 2 * s2 * ca;
            const double J21 = 2 * s1 - 2 * s3 * cb, J23 = 2 * s3 - 2 * s1 * cb;
            const double J31 = 2 * s1 - 2 * s2 * cg, J32 = 2 * s2 - 2 * s1 * cg;
            const double det = J12 * J23 * J31 + J13 * J21 * J32;
            if (!(fabs(det) > 1e-12 * (fabs(J12 * J23 * J31) + fabs(J13 * J21 * J32)) + 1e-300)) break;
            const double d1 = (F1 * (-J23 * J32) + J12 * J23 * F3 + J13 * F2 * J32) / det;
            const double d2 = (F1 * J23 * J31 + J13 * (J21 * F3 - F2 * J31)) / det;
            const double d3 = (-J12 * (J21 * F3 - F2 * J31) + F1 * J21 * J32) / det;
            if (!(d1 == d1) || !(d2 == d2) || !(d3 == d3)) break;
            s1 -= d1; s2 -= d2; s3 -= d3;
        }
        if (!(s1 > 0) || !(s2 > 0) || !(s3 > 0)) continue;
        double P1[3], e1[3], e2[3], e3[3], g1[3], g2[3], g3[3], tmp[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            P1[j] = s1 * fv[0][j];
            g1[j] = s2 * fv[1][j] - P1[j];
            tmp[j] = s3 * fv[2][j] - P1[j];
        }
        cross3(g1, tmp, g3);
        double n1 = norm3(g1), n3 = norm3(g3);
        if (n1 < 1e-300 || n3 < 1e-300) continue;
#pragma unroll
        for (int j = 0; j < 3; j++) { g1[j] /= n1; g3[j] /= n3; }
        cross3(g3, g1, g2);
#pragma unroll
        for (int j = 0; j < 3; j++) { e1[j] = X2[j] - X1[j]; tmp[j] = X3[j] - X1[j]; }
        cross3(e1, tmp, e3);
        n1 = norm3(e1); n3 = norm3(e3);
        if (n1 < 1e-300 || n3 < 1e-300) continue;
#pragma unroll
        for (int j = 0; j < 3; j++) { e1[j] /= n1; e3[j] /= n3; }
        cross3(e3, e1, e2);
        double Rm[9], t[3];
#pragma unroll
        for (int ri = 0; ri < 3; ri++)
#pragma unroll
            for (int cj = 0; cj < 3; cj++) Rm[3 * ri + cj] = g1[ri] * e1[cj] + g2[ri] * e2[cj] + g3[ri] * e3[cj];
#pragma unroll
        for (int j = 0; j < 3; j++) t[j] = P1[j] - (Rm[3 * j] * X1[0] + Rm[3 * j + 1] * X1[1] + Rm[3 * j + 2] * X1[2]);
        const double x = Rm[0] * X4[0] + Rm[1] * X4[1] + Rm[2] * X4[2] + t[0];
        const double yy = Rm[3] * X4[0] + Rm[4] * X4[1] + Rm[5] * X4[2] + t[1];
        const double z = Rm[6] * X4[0] + Rm[7] * X4[1] + Rm[8] * X4[2] + t[2];
        const double du = cx + f * x / z - img[6], dv = cy + f * yy / z - img[7];
        const double e = du * du + dv * dv;
        if (!(e == e)) continue;
        if (!found || e < best) {
            found = true;
            best = e;
            rot_to_rvec(Rm, out.r);
            out.t[0] = t[0]; out.t[1] = t[1]; out.t[2] = t[2];
        }
    }
    return found;
}

// ---------------------------------------------------------------------------- 6x6 solve (LM step)
HOSTDEV bool solve6(double A[36], double b[6], double x[6])
{
    int p[6] = {0, 1, 2, 3, 4, 5};
    for (int c = 0; c < 6; c++) {
        int best = c;
        for (int r = c + 1; r < 6; r++)
            if (fabs(A[6 * p[r] + c]) > fabs(A[6 * p[best] + c])) best = r;
        const int tmp = p[c]; p[c] = p[best]; p[best] = tmp;
        const double piv = A[6 * p[c] + c];
        if (fabs(piv) < 1e-300) return false;
        for (int r = c + 1; r < 6; r++) {
            const double m = A[6 * p[r] + c] / piv;
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
    return true;
}

// M with d(R(r) X)/dr = -R [X]x M,  M = (r r^T + (R^T - I)[r]x) / |r|^2  (identity as r -> 0)
HOSTDEV void rotation_jacobian_factor(const double r[3], const double R[9], double M[9])
{
    const double th2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (th2 < 1e-24) {
        M[0] = M[4] = M[8] = 1; M[1] = M[2] = M[3] = M[5] = M[6] = M[7] = 0;
        return;
    }
    const double A[9] = {R[0] - 1, R[3], R[6], R[1], R[4] - 1, R[7], R[2], R[5], R[8] - 1};
    const double K[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = r[i] * r[j];
#pragma unroll
            for (int k2 = 0; k2 < 3; k2++) s += A[3 * i + k2] * K[3 * k2 + j];
            M[3 * i + j] = s / th2;
        }
}

}  // namespace cl
