// Host build of easysfm_b200/csrc/two_view_math.cuh for the CPU tests (tests/test_two_view.py): the SAME functions the CUDA kernels of
// two_view.cu call, compiled with g++, so that the minimal solver, the sampler and the pose pieces can be compared with the numpy oracle
// without a GPU.  Test infrastructure: nothing in the product loads this library.
#include "../../easysfm_b200/csrc/two_view_math.cuh"

using namespace esfm::tv;

extern "C" {
void tvh_sample(unsigned long long seed, unsigned long long pair, unsigned long long hyp, unsigned m, int* out) {
    int idx[5];
    sample_indices(seed, pair, hyp, m, idx);
    for (int k = 0; k < 5; ++k) out[k] = idx[k];
}
int tvh_five_point(const double* q1, const double* q2, double* out) {
    double a[5][2], b[5][2], Es[10][9];
    for (int i = 0; i < 5; ++i) { a[i][0] = q1[2 * i]; a[i][1] = q1[2 * i + 1]; b[i][0] = q2[2 * i]; b[i][1] = q2[2 * i + 1]; }
    const int n = five_point(a, b, Es);
    for (int k = 0; k < n; ++k)
        for (int e = 0; e < 9; ++e) out[9 * k + e] = Es[k][e];
    return n;
}
double tvh_sampson(const double* E, double x1, double y1, double x2, double y2) { return sampson_error(E, x1, y1, x2, y2); }
int tvh_update_iters(double p, double ep, int mp, int mx) { return ransac_update_num_iters(p, ep, mp, mx); }
void tvh_decompose(const double* E, double* R1, double* R2, double* t) { decompose_essential(E, R1, R2, t); }
void tvh_triangulate(const double* R, const double* t, double ax, double ay, double bx, double by, double* X) {
    double Q[4];
    triangulate_dlt(R, t, ax, ay, bx, by, Q);
    for (int i = 0; i < 4; ++i) X[i] = Q[i];
}
int tvh_cheirality(const double* R, const double* t, double ax, double ay, double bx, double by, double dist) { return cheirality_ok(R, t, ax, ay, bx, by, dist) ? 1 : 0; }
}
