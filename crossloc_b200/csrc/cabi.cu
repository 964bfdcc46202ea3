// extern "C" entry points declared in include/crossloc_b200.h -- common part and the DSAC* solver.
#include "../../include/crossloc_b200.h"

#include <memory>
#include <utility>

#include "cabi_common.h"
#include "dsac.h"

namespace cl {

thread_local std::string g_last_error;

namespace {
std::mutex g_ws_mutex;
std::map<std::pair<int, cudaStream_t>, std::unique_ptr<Workspace>> g_workspaces;
bool g_dsac_timing = false;
}  // namespace

Workspace& workspace_for(cudaStream_t stream)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    auto& slot = g_workspaces[std::make_pair(dev, stream)];
    if (!slot) {
        slot.reset(new Workspace);
        slot->stream = stream;
    }
    return *slot;
}

void release_workspaces()
{
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    for (auto it = g_workspaces.begin(); it != g_workspaces.end();) {
        if (it->first.first == dev) {
            {
                std::lock_guard<std::mutex> use(it->second->mu);
                it->second->release();
            }
            it = g_workspaces.erase(it);
        } else {
            ++it;
        }
    }
}

}  // namespace cl

extern "C" const char* cl_version(void) { return "crossloc_b200 0.1.0 sm_100a"; }

extern "C" const char* cl_last_error(void) { return cl::g_last_error.c_str(); }

extern "C" int cl_dsac_forward_rgb(const float* coords, int B, int Hc, int Wc, float* out_pose, int hyps, float thr,
                                   const float* focal, float cx, float cy, float alpha, float max_reproj,
                                   int subsample, uint64_t seed, uint32_t image_base, uint32_t max_tries, int refine,
                                   const int32_t* forced_samples, int32_t* out_best, double* out_scores,
                                   double* out_hyps, int32_t* out_tries, int32_t* out_counts, double* out_rt,
                                   void* cuda_stream)
{
    using namespace cl;
    if (!coords || !out_pose || !focal) return fail(-1, "cl_dsac_forward_rgb: coords, out_pose and focal must not be NULL");
    if (B < 0 || Hc <= 0 || Wc <= 0 || hyps <= 0 || subsample <= 0)
        return fail(-1, "cl_dsac_forward_rgb: invalid sizes B=%d Hc=%d Wc=%d hyps=%d subsample=%d", B, Hc, Wc, hyps, subsample);
    if (B > 65535) return fail(-1, "cl_dsac_forward_rgb: B=%d exceeds the 65535 images one launch takes", B);
    if (max_tries == 0) return fail(-1, "cl_dsac_forward_rgb: max_tries must be >= 1");
    if (B == 0) return 0;

    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    Workspace& ws = workspace_for(stream);   // intermediates are private to (device, stream): re-entrant across streams
    std::lock_guard<std::mutex> lock(ws.mu);
    Stager st(ws, stream);
    const size_t n = (size_t)Hc * Wc;

    DsacArgs a{};
    a.B = B; a.Hc = Hc; a.Wc = Wc; a.hyps = hyps; a.thr = thr; a.cx = cx; a.cy = cy; a.alpha = alpha;
    a.max_reproj = max_reproj; a.S = subsample; a.seed = seed; a.image_base = image_base; a.max_tries = max_tries;
    a.refine = refine;
    CL_CUDA(st.in("dsac.coords", coords, (size_t)B * 3 * n, &a.coords));
    CL_CUDA(st.in("dsac.focal", focal, (size_t)B, &a.focal));
    CL_CUDA(st.in("dsac.forced", forced_samples, (size_t)B * hyps * 8, &a.forced));
    CL_CUDA(st.out("dsac.pose", out_pose, (size_t)B * 16, &a.out_pose));
    CL_CUDA(st.out("dsac.best", out_best, (size_t)B, &a.out_best));
    CL_CUDA(st.out("dsac.scores", out_scores, (size_t)B * hyps, &a.scores, /*always=*/true));
    CL_CUDA(st.out("dsac.hyps", out_hyps, (size_t)B * hyps * 6, &a.hyp_rt, /*always=*/true));
    CL_CUDA(st.out("dsac.tries", out_tries, (size_t)B * hyps, &a.tries));
    CL_CUDA(st.out("dsac.counts", out_counts, (size_t)B * 100, &a.out_counts));
    CL_CUDA(st.out("dsac.rt", out_rt, (size_t)B * 6, &a.out_rt));
    void* errs;
    CL_CUDA(ws.get("dsac.errs", (size_t)B * n * sizeof(float), &errs));
    a.errs = static_cast<float*>(errs);

    cudaEvent_t* ev = nullptr;
    if (g_dsac_timing) {
        std::array<cudaEvent_t, 4> q{};
        for (auto& e : q) CL_CUDA(cudaEventCreate(&e));
        ws.timing_log.push_back(q);
        ev = ws.timing_log.back().data();
    }
    CL_CUDA(dsac_forward_launch(a, stream, ev));
    CL_CUDA(st.finish());
    return 0;
}

extern "C" int cl_dsac_backward_rgb(const float* coords, int B, int Hc, int Wc, float* grad, const float* gt_pose,
                                    int hyps, float thr, const float* focal, float cx, float cy, float w_rot,
                                    float w_trans, float soft_clamp, float alpha, float max_reproj, int subsample,
                                    uint64_t seed, uint32_t image_base, uint32_t max_tries,
                                    const int32_t* forced_samples, double* out_loss, double* out_probs,
                                    double* out_losses, double* out_hyps, double* out_ref_rt, int32_t* out_tries,
                                    int32_t* out_cells, void* cuda_stream)
{
    using namespace cl;
    if (!coords || !grad || !gt_pose || !focal || !out_loss)
        return fail(-1, "cl_dsac_backward_rgb: coords, grad, gt_pose, focal and out_loss must not be NULL");
    if (B < 0 || Hc <= 0 || Wc <= 0 || hyps <= 0 || subsample <= 0)
        return fail(-1, "cl_dsac_backward_rgb: invalid sizes B=%d Hc=%d Wc=%d hyps=%d subsample=%d", B, Hc, Wc, hyps, subsample);
    if (B > 65535) return fail(-1, "cl_dsac_backward_rgb: B=%d exceeds the 65535 images one launch takes", B);
    if (max_tries == 0) return fail(-1, "cl_dsac_backward_rgb: max_tries must be >= 1");
    if (B == 0) return 0;

    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    Workspace& ws = workspace_for(stream);
    std::lock_guard<std::mutex> lock(ws.mu);
    Stager st(ws, stream);
    const size_t n = (size_t)Hc * Wc, bh = (size_t)B * hyps;

    DsacBwdArgs a{};
    DsacArgs& f = a.fwd;
    f.B = B; f.Hc = Hc; f.Wc = Wc; f.hyps = hyps; f.thr = thr; f.cx = cx; f.cy = cy; f.alpha = alpha;
    f.max_reproj = max_reproj; f.S = subsample; f.seed = seed; f.image_base = image_base; f.max_tries = max_tries;
    f.refine = 1;
    a.w_rot = w_rot; a.w_trans = w_trans; a.soft_clamp = soft_clamp;
    CL_CUDA(st.in("dsac.coords", coords, (size_t)B * 3 * n, &f.coords));
    CL_CUDA(st.in("dsac.focal", focal, (size_t)B, &f.focal));
    CL_CUDA(st.in("dsac.forced", forced_samples, bh * 8, &f.forced));
    CL_CUDA(st.in("dsacb.gt", gt_pose, (size_t)B * 16, &a.gt_pose));
    CL_CUDA(st.inout("dsacb.grad", grad, (size_t)B * 3 * n, &a.grad));
    CL_CUDA(st.out("dsacb.loss", out_loss, (size_t)B, &a.out_loss));
    CL_CUDA(st.out("dsac.scores", (double*)nullptr, bh, &f.scores, /*always=*/true));
    CL_CUDA(st.out("dsac.hyps", out_hyps, bh * 6, &f.hyp_rt, /*always=*/true));
    CL_CUDA(st.out("dsac.tries", out_tries, bh, &f.tries));
    CL_CUDA(st.out("dsacb.cells", out_cells, bh * 8, &f.out_cells, /*always=*/true));
    CL_CUDA(st.out("dsacb.probs", out_probs, bh, &a.probs, /*always=*/true));
    CL_CUDA(st.out("dsacb.losses", out_losses, bh, &a.losses, /*always=*/true));
    CL_CUDA(st.out("dsacb.ref", out_ref_rt, bh * 6, &a.ref_rt, /*always=*/true));
    void* p;
    CL_CUDA(ws.get("dsacb.dloss", bh * 6 * sizeof(double), &p));
    a.dloss = static_cast<double*>(p);
    CL_CUDA(ws.get("dsacb.accepted", bh * sizeof(int32_t), &p));
    a.accepted = static_cast<int32_t*>(p);
    CL_CUDA(ws.get("dsacb.errs", bh * n * sizeof(float), &p));
    a.errs = static_cast<float*>(p);
    CL_CUDA(ws.get("dsacb.inlier", bh * n, &p));
    a.inlier = static_cast<uint8_t*>(p);
    CL_CUDA(ws.get("dsacb.hyp_grad", bh * n * 3 * sizeof(double), &p));
    a.hyp_grad = static_cast<double*>(p);

    CL_CUDA(dsac_backward_launch(a, stream));
    CL_CUDA(st.finish());
    return 0;
}

extern "C" int cl_dsac_timing(int enable, void* cuda_stream, float* ms, int* solves)
{
    using namespace cl;
    if (ms) {
        Workspace& ws = workspace_for(static_cast<cudaStream_t>(cuda_stream));
        std::lock_guard<std::mutex> lock(ws.mu);
        ms[0] = ms[1] = ms[2] = 0.f;
        for (auto& q : ws.timing_log) {
            CL_CUDA(cudaEventSynchronize(q[3]));
            for (int i = 0; i < 3; i++) {
                float t = 0.f;
                CL_CUDA(cudaEventElapsedTime(&t, q[i], q[i + 1]));
                ms[i] += t;
            }
        }
        if (solves) *solves = (int)ws.timing_log.size();
        ws.drop_timing();
    }
    g_dsac_timing = enable != 0;
    return 0;
}

extern "C" int cl_release_workspaces(void)
{
    cl::release_workspaces();
    return 0;
}
