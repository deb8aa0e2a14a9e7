// TEST INFRASTRUCTURE — extern "C" handle on the REFERENCE's own model reader/writer library.
//
// This file is ours; it only *calls* the reference.  `make -C oracle ref` compiles it together with the reference
// sources WHERE THEY LIE under /root/reference (external/tinygsplat/tiny_gsplat.cpp, external/spz/src/*.cc — never
// copied into this repo) into oracle/_ref/libtinygsplat_ref.so.  The tests use it to pin the model writers/readers of
// divshot_b200/csrc/model_io.cpp against the real reference: same bytes out of the writers, same
// values out of the readers.  Only tests/ may load it.
//
// Row layout handed back by ref_load = the reference's `tinygsplat::RichPoint` (tiny_gsplat.hpp:262-269):
// 59 floats = pos[3], shs[48], opacity, scale[3], rot[4].
#include <csetjmp>
#include <csignal>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "tiny_gsplat.hpp"

namespace {
struct Cloud {
    std::vector<glm::vec3> pos, scales;
    std::vector<std::array<float, 48>> shs;
    std::vector<std::array<float, 3>> sh0;
    std::vector<std::array<float, 45>> shn;
    std::vector<glm::vec4> rot;
    std::vector<float> opac;
    std::vector<uint8_t> degrees;
};

// Inputs use the trainer's tensor layouts (diverse/source/assets/gaussian_model.cpp:43-68): sh0[N,3], shN[N,15,3].
// `shs` is assembled exactly as the reference's own caller does (gaussian_model.cpp:428-438).
Cloud make_cloud(int64_t N, const float* pos, const float* sh0, const float* shn, const float* opac, const float* scales,
                 const float* rot, const uint8_t* degrees) {
    Cloud c;
    c.pos.resize(N); c.scales.resize(N); c.shs.resize(N); c.sh0.resize(N); c.shn.resize(N); c.rot.resize(N);
    c.opac.assign(opac, opac + N);
    c.degrees.resize(N);
    for (int64_t i = 0; i < N; i++) {
        c.pos[i] = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        c.scales[i] = glm::vec3(scales[3 * i], scales[3 * i + 1], scales[3 * i + 2]);
        c.rot[i] = glm::vec4(rot[4 * i], rot[4 * i + 1], rot[4 * i + 2], rot[4 * i + 3]);
        for (int k = 0; k < 3; k++) c.sh0[i][k] = c.shs[i][k] = sh0[3 * i + k];
        for (int k = 0; k < 45; k++) c.shn[i][k] = c.shs[i][3 + k] = shn[45 * i + k];
        c.degrees[i] = degrees ? degrees[i] : 3;
    }
    return c;
}
}  // namespace

// The reference's load_ply has no return statement (tiny_gsplat.cpp:632-722).  g++ ends such a function with a trap
// (`ud2`, -funreachable-traps, the -O0 default the Makefile also spells out) AFTER all of its work is done and its
// worker threads have joined.  The unmodified function is therefore run under a SIGILL handler that jumps back here;
// `points` (caller-owned, in memory) is complete at that moment.  Locals of load_ply are leaked — test infrastructure.
namespace {
sigjmp_buf g_jb;
void on_sigill(int) { siglongjmp(g_jb, 1); }
void call_reference_load_ply(const std::string& p, std::vector<tinygsplat::RichPoint>& pts, bool& aa) {
    struct sigaction sa {}, old {};
    sa.sa_handler = on_sigill;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGILL, &sa, &old);
    if (sigsetjmp(g_jb, 1) == 0) (void)tinygsplat::load_ply(p, pts, aa);
    sigaction(SIGILL, &old, nullptr);
}
}  // namespace

extern "C" {

// format: 1 ply, 2 splat, 3 compressed ply, 4 dvsplat, 5 spz, 6 reduced ply (same numbering as include/dvs_model_io.h);
// 7 = the reduced-PLY writer with halfFloat = true (not reachable from the reference's dispatch; feeds the reader test)
__attribute__((visibility("default"))) int ref_save(int format, const char* path, long long N, const float* pos,
                                                    const float* sh0, const float* shn, const float* opac,
                                                    const float* scales, const float* rot, const uint8_t* degrees,
                                                    int antialiased) {
    Cloud c = make_cloud(N, pos, sh0, shn, opac, scales, rot, degrees);
    const std::string p(path);
    bool ok = false;
    switch (format) {
        case 1: ok = tinygsplat::save_ply(p, c.pos, c.scales, c.shs, c.rot, c.opac, antialiased != 0); break;
        case 2: ok = tinygsplat::save_splat(p, c.pos, c.scales, c.shs, c.rot, c.opac); break;
        case 3: ok = tinygsplat::save_compress_ply(p, c.pos, c.scales, c.shs, c.rot, c.opac, antialiased != 0); break;
        case 4: ok = tinygsplat::save_dvs_splat(p, c.pos, c.scales, c.sh0, c.shn, c.rot, c.opac, c.degrees); break;
        case 5: ok = tinygsplat::save_spz_splats(p, c.pos, c.scales, c.shs, c.rot, c.opac, antialiased != 0); break;
        case 6: ok = tinygsplat::save_reduced_ply(p, c.pos, c.scales, c.sh0, c.shn, c.rot, c.opac, c.degrees); break;  // gaussian_model.cpp:447
        case 7: ok = tinygsplat::save_reduced_ply(p, c.pos, c.scales, c.sh0, c.shn, c.rot, c.opac, c.degrees, {}, false, true); break;
        default: return -2;
    }
    return ok ? 0 : -1;
}

// Returns the number of points in the file (or <0); fills rows[min(n,cap)][59].
// The readers' return values are not used (load_ply has none): success is "points came back".
__attribute__((visibility("default"))) long long ref_load(int format, const char* path, float* rows, long long cap,
                                                          int* antialiased) {
    static_assert(sizeof(tinygsplat::RichPoint) == 59 * sizeof(float), "RichPoint is 59 packed floats");
    std::vector<tinygsplat::RichPoint> pts;
    bool aa = false;
    const std::string p(path);
    switch (format) {
        case 1: call_reference_load_ply(p, pts, aa); break;
        case 2: (void)tinygsplat::load_splat(p, pts); break;
        case 3: (void)tinygsplat::load_compress_ply(p, pts, aa); break;
        case 4: (void)tinygsplat::load_dvs_splat(p, pts); break;
        case 5: (void)tinygsplat::load_spz_splats(p, pts, aa); break;
        case 6: case 7: (void)tinygsplat::load_reduced_ply(p, pts); break;
        default: return -2;
    }
    if (antialiased) *antialiased = aa ? 1 : 0;
    const long long n = (long long)pts.size();
    if (rows && n > 0) std::memcpy(rows, pts.data(), sizeof(tinygsplat::RichPoint) * (size_t)std::min(n, cap));
    return n;
}
}
