// Per-hypothesis algebra of the DSAC* backward pass, double precision throughout as the
// reference's cv::Mat_<double> code is.  Reference code restated (all under /root/reference/dsacstar/):
//   dsacstar_loss.h:47-84          calcAngularDistance, loss           -> pose_loss
//   dsacstar_loss.h:96-212         dLoss                               -> pose_loss_jacobian
//   dsacstar_util.h:777-790        trans2pose                          -> trans_to_pose
//   dsacstar_derivative.h:50-109   dProjectdObj                        -> d_project_d_obj
//   dsacstar_util.h:412-432        rows of the refinement Jacobian     -> residual_jacobian_row
//   dsacstar.cpp:408               (J^T J).inv(cv::DECOMP_SVD)         -> sym6_pinv
// HOSTDEV (dsac_common.cuh) lets tests/host_math_harness.cpp compile the same functions for the CPU.
#pragma once
#include "dsac_common.cuh"

namespace cl {

constexpr double kBwdEps = 0.00000001;        // EPS, dsacstar_util.h:45
constexpr double kBwdPiShort = 3.1415926;     // PI, dsacstar_util.h:46 (the forward loss uses this one)
constexpr double kBwdPi = 3.1415926535897932384626433832795;   // CV_PI (dLoss uses this one)
constexpr double kBwdMaxLoss = 10000000.0;    // MAXLOSS, dsacstar_loss.h:35
constexpr double kBwdProbThresh = 0.001;      // PROB_THRESH, dsacstar_derivative.h:36

// dR[9 * i + k] = d R_k / d r_i (k row-major): dR/dr_i = R [m_i]x with m_i the i-th column of the factor M of
// rotation_jacobian_factor (d(R X)/dr = -R [X]x M).  cv::Rodrigues' 3 x 9 Jacobian holds the same numbers.
HOSTDEV void rodrigues_jacobian(const double r[3], const double R[9], double dR[27])
{
    double M[9];
    rotation_jacobian_factor(r, R, M);
    for (int i = 0; i < 3; i++) {
        const double m0 = M[i], m1 = M[3 + i], m2 = M[6 + i];
        const double K[9] = {0, -m2, m1, m2, 0, -m0, -m1, m0, 0};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                dR[9 * i + 3 * a + b] = R[3 * a] * K[b] + R[3 * a + 1] * K[3 + b] + R[3 * a + 2] * K[6 + b];
    }
}

// Inverse of a 3 x 3 matrix by cofactors; false when singular.
HOSTDEV bool inv3(const double A[9], double I[9])
{
    const double c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
    if (det == 0) return false;
    const double id = 1 / det;
    I[0] = c0 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c1 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c2 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return true;
}

// trans2pose: (rvec, tvec) of the inverse of a 4 x 4 transform whose last row is 0 0 0 1.  cv::Rodrigues replaces the
// rotation block by its nearest rotation (U V^T of its SVD) before taking the axis; the Newton iteration
// Q <- (Q + Q^-T) / 2 converges to the same polar factor.
HOSTDEV void trans_to_pose(const double T[16], double rt[6])
{
    const double A[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    double Q[9];
    if (!inv3(A, Q)) {
        for (int j = 0; j < 6; j++) rt[j] = 0;
        return;
    }
    for (int j = 0; j < 3; j++) rt[3 + j] = -(Q[3 * j] * T[3] + Q[3 * j + 1] * T[7] + Q[3 * j + 2] * T[11]);
    for (int it = 0; it < 8; it++) {
        double I[9];
        if (!inv3(Q, I)) break;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) Q[3 * a + b] = 0.5 * (Q[3 * a + b] + I[3 * b + a]);
    }
    rot_to_rvec(Q, rt);
}

// Camera-to-world transform of a scene-to-camera (rvec, tvec): pose2trans, dsacstar_util.h:759-770.
HOSTDEV void pose_to_trans(const double rt[6], double Rc[9], double tc[3])
{
    double R[9];
    rodrigues(rt, R);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) Rc[3 * i + j] = R[3 * j + i];
        tc[i] = -(R[i] * rt[3] + R[3 + i] * rt[4] + R[6 + i] * rt[5]);
    }
}

// loss(estTrans, gtTrans): w_rot * angle (degrees, with the short PI) + w_trans * distance, soft-clamped above cut.
HOSTDEV double pose_loss(const double rt[6], const double gt[16], double w_rot, double w_trans, double cut)
{
    double R1[9], t1[3];
    pose_to_trans(rt, R1, t1);
    // trace(rot2 * rot1^T) = sum_ij rot2_ij rot1_ij
    double trace = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) trace += gt[4 * i + j] * R1[3 * i + j];
    trace = trace > 3.0 ? 3.0 : (trace < -1.0 ? -1.0 : trace);
    const double rot_err = 180 * acos((trace - 1.0) / 2.0) / kBwdPiShort;
    const double d0 = t1[0] - gt[3], d1 = t1[1] - gt[7], d2 = t1[2] - gt[11];
    const double t_err = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    double val = w_rot * rot_err + w_trans * t_err;
    if (val > cut) val = sqrt(cut * val);
    return val < kBwdMaxLoss ? val : kBwdMaxLoss;
}

// dLoss(est, gt): 1 x 6 Jacobian of the loss of the INVERTED poses w.r.t. the estimate's (rvec, tvec).  The clamped
// branch takes sqrt(loss), not sqrt(cut * loss), exactly as dsacstar_loss.h:137-141 does.
HOSTDEV void pose_loss_jacobian(const double est[6], const double gt[6], double w_rot, double w_trans, double cut,
                                double jac[6])
{
    double R1[9], R2[9], dR[27];
    rodrigues(est, R1);
    rodrigues(gt, R2);
    rodrigues_jacobian(est, R1, dR);
    for (int j = 0; j < 6; j++) jac[j] = 0;
    double trace = 0;
    for (int k = 0; k < 9; k++) trace += R1[k] * R2[k];   // trace(rot1 * rot2^T)
    trace = trace > 3.0 ? 3.0 : (trace < -1.0 ? -1.0 : trace);
    const double rot_err = 180 * acos((trace - 1.0) / 2.0) / kBwdPi;
    double it1[3], it2[3];
    for (int i = 0; i < 3; i++) {
        it1[i] = R1[i] * est[3] + R1[3 + i] * est[4] + R1[6 + i] * est[5];
        it2[i] = R2[i] * gt[3] + R2[3 + i] * gt[4] + R2[6 + i] * gt[5];
    }
    const double e0 = it1[0] - it2[0], e1 = it1[1] - it2[1], e2 = it1[2] - it2[2];
    const double t_err = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    double val = w_rot * rot_err + w_trans * t_err;
    bool cut_loss = false;
    if (val > cut) { val = sqrt(val); cut_loss = true; }
    if (val > kBwdMaxLoss) return;
    if ((t_err + rot_err) < kBwdEps) return;
    const double d[3] = {e0 / t_err, e1 / t_err, e2 / t_err};
    double out[6];
    // translation part: d * invRot1, invRot1[i][j] = R1[j][i]
    for (int j = 0; j < 3; j++) out[3 + j] = (d[0] * R1[3 * j] + d[1] * R1[3 * j + 1] + d[2] * R1[3 * j + 2]) * w_trans;
    const double ang = 180 / kBwdPi * -1 / sqrt(3 - trace * trace + 2 * trace);
    for (int q = 0; q < 3; q++) {
        const double* D = dR + 9 * q;
        // d invT1_i / dr_q = sum_j t_j d R1[j][i] / dr_q
        double s = 0;
        for (int i = 0; i < 3; i++) s += d[i] * (est[3] * D[i] + est[4] * D[3 + i] + est[5] * D[6 + i]);
        double tr = 0;
        for (int k = 0; k < 9; k++) tr += R2[k] * D[k];
        out[q] = s * w_trans + ang * tr * w_rot;
    }
    if (cut_loss)
        for (int j = 0; j < 6; j++) out[j] *= 0.5 / val;
    for (int j = 0; j < 6; j++)
        if (out[j] != out[j]) return;   // NaN anywhere: zero Jacobian
    for (int j = 0; j < 6; j++) jac[j] = out[j];
}

// dProjectdObj: derivative of the reprojection error of one point w.r.t. its scene coordinate (zero when the point
// lies in the camera plane or its error exceeds max_reproj).
HOSTDEV void d_project_d_obj(float ptx, float pty, float X, float Y, float Z, const double R[9], const double t[3],
                             double f, double ppx, double ppy, float max_reproj, double out[3])
{
    out[0] = out[1] = out[2] = 0;
    const double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    const double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    const double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    if (fabs(z) < kBwdEps) return;
    const double px = f * x / z + ppx, py = f * y / z + ppy;
    const double ex = ptx - px, ey = pty - py;
    double err = sqrt(ex * ex + ey * ey);
    if (err > max_reproj) return;
    err += kBwdEps;
    for (int j = 0; j < 3; j++) {
        const double pxd = f * R[j] / z - f * x / z / z * R[6 + j];
        const double pyd = f * R[3 + j] / z - f * y / z / z * R[6 + j];
        out[j] = 0.5 / err * (2 * ex * -pxd + 2 * ey * -pyd);
    }
}

// Precomputed rotation terms of cv::projectPoints' Jacobian w.r.t. (rvec, tvec) for one pose.
struct ProjJac {
    double R[9], RM[9], t[3];
};

HOSTDEV void proj_jac_setup(const double rt[6], ProjJac& pj)
{
    double M[9];
    rodrigues(rt, pj.R);
    rotation_jacobian_factor(rt, pj.R, M);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++)
            pj.RM[3 * i + j] = pj.R[3 * i] * M[j] + pj.R[3 * i + 1] * M[3 + j] + pj.R[3 * i + 2] * M[6 + j];
        pj.t[i] = rt[3 + i];
    }
}

// One row of jacobeanHyp / jacobeanR: d |proj - pt| / d (rvec, tvec) (dsacstar_util.h:412-432, dsacstar.cpp:391-405).
// The projection is rounded to float before the difference (cv::Point2f); rows with an error above max_reproj stay 0.
HOSTDEV void residual_jacobian_row(const ProjJac& pj, double f, double cx, double cy, float X, float Y, float Z,
                                   int ptx, int pty, float max_reproj, double row[6])
{
    for (int j = 0; j < 6; j++) row[j] = 0;
    const double Xw = X, Yw = Y, Zw = Z;
    const double qx = pj.R[0] * Xw + pj.R[1] * Yw + pj.R[2] * Zw;
    const double qy = pj.R[3] * Xw + pj.R[4] * Yw + pj.R[5] * Zw;
    const double qz = pj.R[6] * Xw + pj.R[7] * Yw + pj.R[8] * Zw;
    const double x = qx + pj.t[0], y = qy + pj.t[1], z = qz + pj.t[2];
    const double iz = z ? 1. / z : 1;
    const double u = f * (x * iz) + cx, v = f * (y * iz) + cy;
    const float du = (float)u - (float)ptx, dv = (float)v - (float)pty;
    double err = sqrt((double)du * du + (double)dv * dv);
    if (err < kBwdEps) err = kBwdEps;
    if (err > max_reproj) return;
    const double nu = 1 / err * du, nv = 1 / err * dv;
    const double a0 = f * iz, a2 = -f * x * iz * iz, b2 = -f * y * iz * iz;
    for (int j = 0; j < 3; j++) {
        const double dx = -(-qz * pj.RM[3 + j] + qy * pj.RM[6 + j]);
        const double dy = -(qz * pj.RM[j] - qx * pj.RM[6 + j]);
        const double dz = -(-qy * pj.RM[j] + qx * pj.RM[3 + j]);
        row[j] = nu * (a0 * dx + a2 * dz) + nv * (a0 * dy + b2 * dz);
    }
    row[3] = nu * a0;
    row[4] = nv * a0;
    row[5] = nu * a2 + nv * b2;
}

// Pseudo-inverse of a symmetric 6 x 6 matrix the way cv::invert(DECOMP_SVD) forms it: cyclic Jacobi rotations, then
// sum v v^T / lambda over the eigenvalues with |lambda| > 2 DBL_EPSILON sum |lambda| (SVBkSb's threshold).
HOSTDEV void sym6_pinv(const double Ain[36], double P[36])
{
    double A[36], V[36];
    for (int i = 0; i < 36; i++) { A[i] = Ain[i]; V[i] = 0; }
    for (int i = 0; i < 6; i++) V[7 * i] = 1;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0, diag = 0;
        for (int i = 0; i < 6; i++) {
            diag += A[7 * i] * A[7 * i];
            for (int j = i + 1; j < 6; j++) off += A[6 * i + j] * A[6 * i + j];
        }
        if (off <= 1e-60 * diag || off == 0) break;
        for (int p = 0; p < 5; p++)
            for (int q = p + 1; q < 6; q++) {
                const double apq = A[6 * p + q];
                if (apq == 0) continue;
                const double theta = (A[7 * q] - A[7 * p]) / (2 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 6; k++) {
                    const double akp = A[6 * k + p], akq = A[6 * k + q];
                    A[6 * k + p] = c * akp - s * akq;
                    A[6 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 6; k++) {
                    const double apk = A[6 * p + k], aqk = A[6 * q + k];
                    A[6 * p + k] = c * apk - s * aqk;
                    A[6 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 6; k++) {
                    const double vkp = V[6 * k + p], vkq = V[6 * k + q];
                    V[6 * k + p] = c * vkp - s * vkq;
                    V[6 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    double sum = 0;
    for (int i = 0; i < 6; i++) sum += fabs(A[7 * i]);
    const double thr = sum * (DBL_EPSILON * 2);
    for (int i = 0; i < 36; i++) P[i] = 0;
    for (int e = 0; e < 6; e++) {
        const double w = A[7 * e];
        if (fabs(w) <= thr) continue;
        const double iw = 1 / w;
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) P[6 * a + b] += V[6 * a + e] * V[6 * b + e] * iw;
    }
}

}  // namespace cl
