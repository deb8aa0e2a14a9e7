// TEST HARNESS (not product code): compiles divshot_b200/csrc/aux_normal_ops.h — the per-Gaussian normal arithmetic the
// kernels of aux_outputs.cu call — for the host, so tests/aux_ref.py can use it in the oracle-level restatement that
// tests/test_aux_outputs.py checks against float64 autograd.
#include "aux_normal_ops.h"

extern "C" {
void t_normals_forward(const float* quats, const float* scales, const float* means, const float* view, int activated, long long N,
                       float* n_v, int* axis, float* flip) {
    for (long long i = 0; i < N; i++)
        dvs_aux::normal_forward(quats + 4 * i, scales + 3 * i, means + 3 * i, view, activated != 0, n_v + 3 * i, axis + i, flip + i);
}
void t_normals_backward(const float* quats, const int* axis, const float* flip, const float* view, int activated, long long N,
                        const float* dn_v, float* dq) {
    for (long long i = 0; i < N; i++)
        dvs_aux::normal_backward(quats + 4 * i, axis[i], flip[i], view, activated != 0, dn_v + 3 * i, dq + 4 * i);
}
}
