// TEST HARNESS (not product code): compiles divshot_b200/csrc/sh_grad_ops.h — the arithmetic of the kernel in
// sh_exchange.cu — for the host, so tests/test_sh_exchange.py can check the factored SH gradient against the oracle's
// per-view dL/dshN without a GPU.
#include <cstring>

#include "sh_grad_ops.h"

extern "C" void t_sh_grad_from_dsh0(const float* means, const float* campos /*[V,3]*/, const float* dsh0_all /*[V,N,3]*/, long long N,
                                    int V, int deg, int KR, float* out /*[N,KR,3]*/) {
    for (long long i = 0; i < N; i++) {
        float acc[45];
        std::memset(acc, 0, sizeof acc);
        for (int v = 0; v < V; v++) dvs_shx::accumulate_view(deg, means + 3 * i, campos + 3 * v, dsh0_all + ((size_t)v * N + i) * 3, acc);
        for (int k = 0; k < 3 * KR; k++) out[(size_t)i * 3 * KR + k] = k < 45 ? acc[k] : 0.f;
    }
}
