// Host build of crossloc_b200/csrc/dsac_common.cuh so that the CUDA solver's geometry (P3P, quartic,
// Rodrigues, RNG) can be checked against cv2 on a machine without a GPU.  Test-only; never shipped.
#include "../crossloc_b200/csrc/dsac_common.cuh"

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
}
