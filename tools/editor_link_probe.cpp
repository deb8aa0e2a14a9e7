// editor_link_probe.cpp — uses the trainer the way the reference EDITOR does (application/editor/source/editor.cpp:
// 846-855 construct + loadTrainData + trainSetup, :1426-1654 trainStep / getters): linked against libgstrain.so, calling
// the GaussianTrainerScene class directly instead of the nine dlsym'd C symbols of the CLI.  It trains a few iterations,
// then takes the model across the trainer -> viewer boundary both ways:
//   (a) the reference's way: six getGaussian*Cpu() vectors (editor.cpp:1559-1566), written raw to <out>.raw
//   (b) the fused device pack (requestViewerPack / acquireViewerPack), written to <out>.pack
// tests/test_viewer_pack.py quantises (a) with the reference's own CPU code and expects the bytes of (b).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "gaussian_trainer_scene.hpp"

static void put(FILE* f, const void* p, size_t n) { if (std::fwrite(p, 1, n, f) != n) { std::perror("fwrite"); std::exit(7); } }

int main(int argc, char** argv) {
    const std::string data = argc > 1 ? argv[1] : "synthetic:N=20000,W=320,H=240,views=4,deg=1";
    const int iters = argc > 2 ? std::atoi(argv[2]) : 30;
    const std::string out = argc > 3 ? argv[3] : "/tmp/editor_link_probe";
    GaussianTrainConfig cfg;
    cfg.sourcePath = data; cfg.numIters = iters;
    GaussianTrainerScene scene(cfg, -1);
    if (!scene.loadTrainData(data)) { std::fprintf(stderr, "loadTrainData failed\n"); return 4; }
    scene.trainSetup();
    GaussianViewerPack early;
    if (scene.acquireViewerPack(early, false)) { std::fprintf(stderr, "a pack before any request\n"); return 5; }
    for (int i = 0; i < iters; i++) {
        scene.trainStep();
        if (i == iters / 2) scene.requestViewerPack();  // an older snapshot: must be superseded by the last request
    }
    scene.requestViewerPack();
    GaussianViewerPack pk;
    if (!scene.acquireViewerPack(pk, true)) { std::fprintf(stderr, "acquireViewerPack failed\n"); return 5; }
    if (pk.iteration != scene.getCurrentIterations() || pk.count != scene.getNumGaussians()) { std::fprintf(stderr, "stale pack\n"); return 6; }
    const long long n = pk.count;
    {
        FILE* f = std::fopen((out + ".pack").c_str(), "wb");
        if (!f) return 7;
        put(f, &n, 8); put(f, pk.bboxMin, 12); put(f, pk.bboxMax, 12);
        put(f, pk.gaussians, (size_t)n * 32); put(f, pk.colors, (size_t)n * 8); put(f, pk.sh, (size_t)n * 64);
        std::fclose(f);
    }
    {
        FILE* f = std::fopen((out + ".raw").c_str(), "wb");
        if (!f) return 7;
        put(f, &n, 8);
        const auto pos = scene.getGaussianPositionCpu(); const auto sc = scene.getGaussianScalingsCpu();
        const auto rot = scene.getGaussianRotationsCpu(); const auto op = scene.getGaussianOpcaitiesCpu();
        const auto sh0 = scene.getGaussianSH0Cpu(); const auto shn = scene.getGaussianSHNCpu();
        put(f, pos.data(), pos.size() * 4); put(f, sc.data(), sc.size() * 4); put(f, rot.data(), rot.size() * 4);
        put(f, op.data(), op.size() * 4); put(f, sh0.data(), sh0.size() * 4); put(f, shn.data(), shn.size() * 4);
        std::fclose(f);
    }
    std::printf("packed %lld gaussians after %d iterations, bbox [%g %g %g] .. [%g %g %g]\n", n, pk.iteration, pk.bboxMin[0],
                pk.bboxMin[1], pk.bboxMin[2], pk.bboxMax[0], pk.bboxMax[1], pk.bboxMax[2]);
    return 0;
}
