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

// The kernel of sh_exchange.cu thread by thread: the same two phase functions, the same tile loop, 128 "threads" per CTA and
// `grid` CTAs; the shared-memory rows are a plain array.  Lets the CPU suite (and ASan) see the indexing the GPU will run.
extern "C" void t_sh_exchange_kernel_emulation(const float* means, const float* campos, const float* dsh0_all, long long N, int V, int deg,
                                               int KR, float* out, int grid) {
    const int T = 128;
    alignas(16) static float rows[128 * 45];
    const dvs_shx::ExchangeArgs a{means, campos, dsh0_all, N, V, deg, 3 * KR, out, (int)((reinterpret_cast<unsigned long long>(out) & 15ull) == 0)};
    const long long n_tiles = (N + T - 1) / T;
    for (int block = 0; block < grid; block++)
        for (long long tile = block; tile < n_tiles; tile += grid) {
            const long long base = tile * T;
            const int cnt = (int)(N - base < T ? N - base : T);
            for (int k = 0; k < 128 * 45; k++) rows[k] = -12345.0f;  // stale shared memory must never reach the output
            for (int tid = 0; tid < T; tid++) dvs_shx::exchange_compute(a, rows, tid, base, cnt);
            for (int tid = 0; tid < T; tid++) dvs_shx::exchange_store(a, rows, tid, T, base, cnt);
        }
}
