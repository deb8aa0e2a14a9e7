// editor_api_probe.cpp — compiles the call patterns the reference EDITOR uses on GaussianTrainerScene against our authored
// header, with the reference's own glm on the include path, and links them against libgstrain.so.  Each block cites the
// reference lines whose expression shapes it repeats (argument and result types are what matters: glm::mat4 / glm::quat /
// glm::vec3 results fed to glm functions, glm vectors passed in).  Compiling + linking is the CPU-side check; run on a
// GPU it also prints a small report that tests/test_zz_staged_editor_api.py reads.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <glm/gtc/quaternion.hpp>

#include "gaussian_trainer_scene.hpp"

#ifndef GSTRAIN_HAS_GLM
#error "the editor surface needs glm on the include path"
#endif
// entt keeps the trainer as a component (editor.cpp:2024 add_component<GaussianTrainerScene>(trainConfig, -1))
static_assert(std::is_move_constructible<GaussianTrainerScene>::value && std::is_move_assignable<GaussianTrainerScene>::value, "entt component");
static_assert(sizeof(GsPoint3D) == 16, "update_from_pos_color reads 16-byte records (gaussian_model.cpp:70-95)");

int main(int argc, char** argv) {
    const std::string data = argc > 1 ? argv[1] : "synthetic:N=5000,W=160,H=120,views=3,deg=1";
    const std::string out = argc > 2 ? argv[2] : "/tmp/editor_api_probe";
    GaussianTrainConfig trainConfig;
    trainConfig.numIters = 40;
    std::vector<GaussianTrainerScene> pool;  // a movable component store
    pool.emplace_back(trainConfig, -1);
    GaussianTrainerScene moved(std::move(pool[0]));
    pool.clear();
    GaussianTrainerScene& gs_train = moved;
    gs_train.setModelPath(out + ".ply");                                                   // editor.cpp:2024
    if (!gs_train.loadTrainData(data)) return 4;                                           // :2034
    gs_train.trainSetup();                                                                 // :2035
    if (gs_train.getCurrentTrainingStatus() == TrainingStatus::Preprocess_Done)            // :1455-1457
        gs_train.setTrainingStatus(TrainingStatus::Training);
    // editor.cpp:850-856 — frustum of every training camera
    for (auto i = 0; i < gs_train.getNumCameras(); i++) {
        auto projection = gs_train.getCameraProjection(i);
        auto viewR = gs_train.getCameraRotation(i);
        auto viewT = gs_train.getCameraPos(i);
        auto view = glm::translate(glm::identity<glm::mat4>(), viewT) * glm::mat4_cast(viewR);
        const glm::mat4 vp = projection * glm::inverse(view);
        // the pose built this way must be the inverse of the world->camera matrix the rasterizer uses
        const auto V = gs_train.getCameraView(i);
        const glm::mat4 W2C = glm::make_mat4(V.data());
        const glm::mat4 id = W2C * view;
        float err = 0.f;
        for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) err = std::max(err, std::abs(id[c][r] - (c == r ? 1.f : 0.f)));
        std::printf("camera %d pose_error %.3g vp00 %.6f\n", i, err, vp[0][0]);
        if (err > 1e-4f) return 8;
    }
    // editor.cpp:1486-1487, inspector_panel.cpp:909-933, scene_view_panel.cpp:1393-1394 — focus region
    gs_train.focus_region_position = glm::vec3(0.0f);
    gs_train.focus_region_scale = glm::vec3(1.0f);
    glm::vec3 focus_pos = gs_train.focus_region_position;
    glm::vec3 focus_rot = gs_train.focus_region_rotation;
    glm::vec3 focus_scale = gs_train.focus_region_scale;
    focus_pos.x = 1.5f; focus_rot.z = 90.f; focus_scale.y = 2.f;
    gs_train.updateFocusRegion(focus_pos, focus_rot, focus_scale);
    auto [a, b] = gs_train.getFocusRegion();
    auto focusTransform = gs_train.getFocusRegionTransform();
    const glm::vec4 ex = focusTransform * glm::vec4(1, 0, 0, 1);  // Rz(90) * S: x axis -> +y, then translated by 1.5 in x
    std::printf("focus box [%g %g %g]..[%g %g %g] ex [%g %g %g]\n", a.x, a.y, a.z, b.x, b.y, b.z, ex.x, ex.y, ex.z);
    if (std::abs(ex.x - 1.5f) > 1e-5f || std::abs(ex.y - 1.f) > 1e-5f) return 9;
    // editor.cpp:1523-1527 — the point cloud shown before training
    const auto& points3d = gs_train.getPoints3D(0);
    const unsigned char* pos_color_h = (const unsigned char*)points3d.data();
    std::printf("points3d %zu first %g %g %g rgb %d %d %d\n", points3d.size(), *(const float*)(pos_color_h + 0), *(const float*)(pos_color_h + 4),
                *(const float*)(pos_color_h + 8), pos_color_h[12], pos_color_h[13], pos_color_h[14]);
    // img2d_dataset_panel.cpp:113-115
    auto splat_imge = gs_train.getSplatImageView(1);
    std::printf("image %s %dx%d alpha %d\n", splat_imge.name.c_str(), splat_imge.width, splat_imge.height, splat_imge.data[3]);
    // editor.cpp:1603-1649 — the training thread
    gs_train.startTrain();
    while (gs_train.getCurrentIterations() < gs_train.getTrainConfig().numIters) {
        if (gs_train.isTerminate()) break;
        if (gs_train.isTrain()) gs_train.trainStep();
    }
    std::printf("phase %s progress %.2f elapsed %.3f remaining %.3f loss %.5f\n", gs_train.getCurrentTrainingPhaseName().c_str(),
                gs_train.getProgressOnCurrentPhase(), gs_train.getTrainingElpasedTime(), gs_train.getEstimateTrainingTime(), gs_train.getCurrentLoss());
    // inspector_panel.cpp:1037-1044 — ApplyEdit: the editor hands back an edited model (here: every second splat deleted)
    {
        const auto p = gs_train.getGaussianPositionCpu(); const auto r = gs_train.getGaussianRotationsCpu();
        const auto s = gs_train.getGaussianScalingsCpu(); const auto o = gs_train.getGaussianOpcaitiesCpu();
        const auto c0 = gs_train.getGaussianSH0Cpu(); const auto cn = gs_train.getGaussianSHNCpu();
        std::vector<glm::vec3> position, scale; std::vector<glm::vec4> rotation; std::vector<float> opacity;
        std::vector<std::array<float, 3>> sh0; std::vector<std::array<float, 45>> shn;
        for (size_t i = 0; i < o.size(); i += 2) {
            position.emplace_back(p[3 * i], p[3 * i + 1], p[3 * i + 2]); scale.emplace_back(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
            rotation.emplace_back(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]); opacity.push_back(o[i]);
            sh0.push_back({c0[3 * i], c0[3 * i + 1], c0[3 * i + 2]});
            std::array<float, 45> row; for (int k = 0; k < 45; k++) row[k] = cn[45 * i + k];
            shn.push_back(row);
        }
        gs_train.updateTensorFromGaussianData(position, rotation, scale, opacity, sh0, shn);
        const auto back = gs_train.getGaussianPositionCpu();
        const bool same = back.size() == 3 * position.size() && back[0] == position[0].x && back[back.size() - 1] == position.back().z;
        std::printf("edited model %lld gaussians roundtrip %s\n", (long long)gs_train.getNumGaussians(), same ? "ok" : "MISMATCH");
        if (!same) return 10;
        gs_train.trainStep();  // trains on after the edit
    }
    // inspector_panel.cpp:789,837 — strategy switch and reset
    gs_train.getTrainConfig().densifyStrategy = 0;
    gs_train.setDensifyStrategy(0);
    gs_train.resetGaussian();
    std::printf("after reset: %lld gaussians iteration %d phase %s\n", (long long)gs_train.getNumGaussians(), gs_train.getCurrentIterations(),
                gs_train.getCurrentTrainingPhaseName().c_str());
    if ((size_t)gs_train.getNumGaussians() != points3d.size() || gs_train.getCurrentIterations() != 0) return 11;
    // editor.cpp:3512,3535,1629
    if (!gs_train.saveCameraDatas(out + "_cameras.json") || !gs_train.exportSparsePointCloud(out + "_points.ply")) return 12;
    gs_train.saveGaussianModel();
    gs_train.pruenIteraions.emplace_back();                                                // inspector_panel.cpp:956
    gs_train.ShowTrainView = true;
    std::printf("ok\n");
    return 0;
}
