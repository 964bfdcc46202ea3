// Host build of crossloc_b200/csrc/dsac_common.cuh so that the CUDA solver's geometry (P3P, quartic,
// Rodrigues, RNG) can be checked against cv2 on a machine without a GPU.  Test-only; never shipped.
#include "../crossloc_b200/csrc/dsac_common.cuh"
#include "../crossloc_b200/csrc/dsac_backward_math.cuh"

extern "C" {
int hm_p3p(const double* obj, const double* img, double f, double cx, double cy, double* r, double* t)
{
    cl::Pose p;
    bool ok = cl::p3p_solve(obj, img, f, cx, cy, p);
    for (int j = 0; j < 3; j++) { r[j] = ok ? p.r[j] : 0; t[j] = ok ? p.t[j] : 0; }
    return ok;
}
void hm_sample_cells(uint64_t seed, uint32_t image, uint32_t hyp, uint32_t tr, int w, int h, int* cells)
{
    cl::sample_cells(seed, image, hyp, tr, w, h, cells);
}
void hm_rodrigues(const double* r, double* R) { cl::rodrigues(r, R); }
void hm_rot_to_rvec(const double* R, double* r) { cl::rot_to_rvec(R, r); }
int hm_quartic(const double* c, double* roots) { return cl::quartic_real_roots(c, roots); }
float hm_repro_error(const double* R, const double* t, float f, float cx, float cy, float X, float Y, float Z, int px,
                     int py, float max_reproj)
{
    return cl::repro_error(R, t, f, cx, cy, X, Y, Z, px, py, max_reproj);
}
// ---- backward pass algebra (dsac_backward_math.cuh)
double hm_pose_loss(const double* rt, const double* gt16, double w_rot, double w_trans, double cut)
{
    return cl::pose_loss(rt, gt16, w_rot, w_trans, cut);
}
void hm_pose_loss_jacobian(const double* est, const double* gt, double w_rot, double w_trans, double cut, double* jac)
{
    cl::pose_loss_jacobian(est, gt, w_rot, w_trans, cut, jac);
}
void hm_trans_to_pose(const double* T, double* rt) { cl::trans_to_pose(T, rt); }
void hm_d_project_d_obj(float ptx, float pty, float X, float Y, float Z, const double* R, const double* t, double f,
                        double ppx, double ppy, float max_reproj, double* out)
{
    cl::d_project_d_obj(ptx, pty, X, Y, Z, R, t, f, ppx, ppy, max_reproj, out);
}
void hm_residual_row(const double* rt, double f, double cx, double cy, float X, float Y, float Z, int ptx, int pty,
                     float max_reproj, double* row)
{
    cl::ProjJac pj;
    cl::proj_jac_setup(rt, pj);
    cl::residual_jacobian_row(pj, f, cx, cy, X, Y, Z, ptx, pty, max_reproj, row);
}
void hm_sym6_pinv(const double* A, double* P) { cl::sym6_pinv(A, P); }
}
