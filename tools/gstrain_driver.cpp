// gstrain_driver.cpp — minimal stand-in for the reference CLI loop (application/diverseshot-cli/source/gs_train.cpp:20-179):
// dlopen("libgstrain.so"), resolve the nine symbols by name, create_splat -> load_train_data -> train_step loop ->
// save_splat_model -> delete_splat -> gstrain_destroy.  Used by tests/test_plugin.py on the GPU box, where the
// reference sources (and therefore the real CLI build) may be absent.  Prints the loss trajectory.
// Optional trailing key=value arguments set schedule fields the CLI takes as flags (main.cpp:19-70):
// warmup= refineEvery= refineStop= resetAlphaEvery= capMax= strategy= (densifyStrategy: 0 ADC, 1 MCMC, 2 ADC+)
// visibleAdam= revisedOpacity= enableBg= (0 / 1); modelType= (0 3DGS, 1 2DGS); normalLoss= (0 / 1); loadItr= (create_splat's second argument: resume from the model at <out> at that
// iteration, main.cpp:40-41); lossCheck=0 (exit 0 even if the loss did not fall by 20 %: short resume / timing runs).
// Data parallel: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK (or DVS_*) set, e.g. under
// `python -m torch.distributed.run --no-python`; every rank runs this same loop, rank 0's output file is the model.
// Prints the loss trajectory and the training rate (iterations/s over the loop, wall clock).
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "gaussian_trainer_scene.hpp"

int main(int argc, char** argv) {
    const std::string data = argc > 1 ? argv[1] : "synthetic:N=20000,W=320,H=240,views=4,deg=1";
    const int iters = argc > 2 ? std::atoi(argv[2]) : 300;
    const std::string out = argc > 3 ? argv[3] : "/tmp/gstrain_driver.ply";
    void* h = dlopen("libgstrain.so", RTLD_LAZY | RTLD_LOCAL);
    if (!h) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) { std::fprintf(stderr, "missing symbol %s\n", n); std::exit(3); } return p; };
    auto init = (void (*)())sym("gstrain_init");
    auto create = (void* (*)(const GaussianTrainConfig&, int))sym("create_splat");
    auto load = (bool (*)(GaussianTrainerScene*, const std::string&))sym("load_train_data");
    auto step = (void (*)(GaussianTrainerScene*))sym("train_step");
    auto save = (void (*)(GaussianTrainerScene*))sym("save_splat_model");
    auto mesh = (void (*)(GaussianTrainerScene*))sym("export_mesh");
    auto del = (void (*)(GaussianTrainerScene*))sym("delete_splat");
    auto cur = (int (*)(GaussianTrainerScene*))sym("get_cur_step");
    auto destroy = (void (*)())sym("gstrain_destroy");
    (void)mesh;
    init();
    GaussianTrainConfig cfg;
    cfg.sourcePath = data; cfg.modelPath = out; cfg.numIters = iters; cfg.verbose = true;
    int load_itr = -1, loss_check = 1;
    for (int a = 4; a < argc; a++) {
        const std::string kv = argv[a];
        const size_t eq = kv.find('=');
        if (eq == std::string::npos) { std::fprintf(stderr, "expected key=value, got %s\n", argv[a]); return 6; }
        const std::string k = kv.substr(0, eq);
        const int v = std::atoi(kv.c_str() + eq + 1);
        if (k == "warmup") cfg.warmupLength = v;
        else if (k == "refineEvery") cfg.refineEvery = v;
        else if (k == "refineStop") cfg.refineStopIter = v;
        else if (k == "resetAlphaEvery") cfg.resetAlphaEvery = v;
        else if (k == "capMax") cfg.capMax = v;
        else if (k == "strategy") cfg.densifyStrategy = v;
        else if (k == "visibleAdam") cfg.visibleAdam = v != 0;
        else if (k == "revisedOpacity") cfg.revisedOpacity = v != 0;
        else if (k == "enableBg") cfg.enableBg = v != 0;
        else if (k == "modelType") cfg.modelType = v;
        else if (k == "packLevel") cfg.packLevel = v;
        else if (k == "normalLoss") cfg.normalConsistencyLoss = v != 0;
        else if (k == "loadItr") load_itr = v;
        else if (k == "lossCheck") loss_check = v;
        else if (k == "verbose") cfg.verbose = v != 0;
        else if (k == "numIters") cfg.numIters = v;  // the schedule's length when it differs from the steps run here
        else { std::fprintf(stderr, "unknown option %s\n", k.c_str()); return 6; }
    }
    auto* scene = (GaussianTrainerScene*)create(cfg, load_itr);
    if (!load(scene, data)) { std::fprintf(stderr, "load_train_data failed\n"); return 4; }
    float first = -1.f, last = -1.f;
    const int start_step = cur(scene);
    // the rate is measured after a few warm-up steps (first-use allocations, NCCL's lazy connection set-up)
    const int timed_from = start_step + std::min(20, std::max(0, (iters - start_step) / 4));
    auto t0 = std::chrono::steady_clock::now();
    while (true) {
        const int i = cur(scene);
        if (i >= iters) break;
        if (i == timed_from) t0 = std::chrono::steady_clock::now();
        step(scene);
        last = scene->getCurrentLoss();
        if (i == 0) first = last;
        if (i % 50 == 0) std::printf("iter %d loss %.6f\n", i, last);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const char* rk = std::getenv("DVS_RANK") ? std::getenv("DVS_RANK") : std::getenv("RANK");
    if (!rk || std::atoi(rk) == 0) save(scene);  // data parallel: every rank holds the same model, rank 0 writes it
    std::printf("steps %d first_loss %.6f last_loss %.6f its_per_s %.1f (from step %d)\n", cur(scene), first, last,
                secs > 0 ? (cur(scene) - timed_from) / secs : 0.0, start_step);
    del(scene);
    destroy();
    dlclose(h);
    return (!loss_check || last < 0.8f * first) ? 0 : 5;
}
