// TEST INFRASTRUCTURE: host build of the kernel's evaluation core (boxfusion_b200/csrc/bf_refine_eval.cuh, the very
// source nvcc compiles into bf_refine_kernel) so that tests/test_eval_core_host.py can compare it bit for bit with the
// CPU oracle without a GPU.  Built with g++ -O2 -ffp-contract=off (the counterpart of nvcc -fmad=false).
// Nothing here is shipped or measured.
#include "../../boxfusion_b200/csrc/bf_refine_eval.cuh"

extern "C" {

// `rolled` selects the kernel's code-size variant (compact rolled loops for the saturated regime, or fully unrolled).
// bf_evaluate_kernel's loops on the host: fitness[P] of one box against V views (box_fusion.py:413-461).
// stats[0] += evaluations, stats[1] += evaluations that needed at least one exact fallback test,
// stats[2] += fallback tests, stats[3] |= candidate overflow.
void bfh_evaluate(const float* box6, const float* t_c /*[V,16]*/, const float* pst /*[P,6]*/, int P, int n_eval,
                  const float* rot9, const float* poses /*[V,16]*/, int V, float fx, float cx, float fy, float cy,
                  const float* search6, float img_h, float img_w, float* fitness /*[P]*/, long long* stats /*[4]*/, int rolled) {
    bf_view* views = new bf_view[V];
    for (int v = 0; v < V; ++v) bf_view_stage(views[v], poses + 16 * v, t_c + 16 * v, img_w, img_h);
    long long n_ev = 0, n_fb_ev = 0, n_fb = 0;
    int any_over = 0;
#pragma omp parallel for schedule(static) reduction(+ : n_ev, n_fb_ev, n_fb) reduction(| : any_over)
    for (int p = 0; p < P; ++p) {
        float value = 0.0f, count = 0.0f;
        if (p < n_eval) {
            float c[8][3];
            bf_particle_corners(box6, pst + 6 * p, search6, rot9, c);
            for (int v = 0; v < V; ++v) {
                int over = 0, fb = 0;
                value += rolled ? bf_eval_view<true>(c, views[v], fx, cx, fy, cy, img_w, img_h, &over, &fb)
                                : bf_eval_view<false>(c, views[v], fx, cx, fy, cy, img_w, img_h, &over, &fb);
                count += 1;
                n_ev += 1; n_fb_ev += (fb > 0); n_fb += fb; any_over |= over;
            }
        }
        fitness[p] = value / (count + 1e-6f);
    }
    if (stats) { stats[0] += n_ev; stats[1] += n_fb_ev; stats[2] += n_fb; stats[3] |= any_over; }
    delete[] views;
}

// box_fusion.py:380-398 on two raw 8-point sets (a = particle side, b = observation side), through the same
// bf_hull8 / bf_view_finish / bf_hull_iou path the kernel takes.
float bfh_iou_points(const float* a16, const float* b16, float img_w, float img_h, int* fallbacks, int* overflow, int rolled) {
    bf_view vw;
    float pose[16] = {0};
    bf_view_stage(vw, pose, b16, img_w, img_h);
    P2 uv[8];
    for (int k = 0; k < 8; ++k) { uv[k].x = a16[2 * k]; uv[k].y = a16[2 * k + 1]; }
    P2 hm[16];
    for (int k = 0; k < 16; ++k) { hm[k].x = 0.f; hm[k].y = 0.f; }
    const int n0 = rolled ? bf_hull8<true>(uv, hm) : bf_hull8<false>(uv, hm);
    P2 h0[8];
    for (int k = 0; k < 8; ++k) h0[k] = hm[k];
    int fb = 0, over = 0;
    bf_divrange dk = bf_divrange_init();
    const float iou = rolled ? bf_hull_iou<true>(h0, hm, n0, vw, &over, &fb, dk) : bf_hull_iou<false>(h0, hm, n0, vw, &over, &fb, dk);
    if (fallbacks) *fallbacks = fb;
    if (overflow) *overflow = over;
    return iou;
}

}  // extern "C"
