/*
 * gaussian_trainer_scene.hpp — public header of the `gstrain` trainer plugin.
 *
 * The reference includes this file (application/diverseshot-cli/source/gs_train.cpp:3,
 * application/editor/source/editor.cpp) from `${GSTRAIN_INCLUDE_DIR}` = diverse_utils/gstrain/src
 * (CMakeLists.txt:101, application/diverseshot-cli/CMakeLists.txt:52) but does not ship it — the trainer is
 * closed (README.md:32,46).  This header is authored from the *usage* in the reference:
 *   - every GaussianTrainConfig field assigned at gs_train.cpp:50-99 and editor.cpp:1750-2020,2206;
 *   - GSPackLevel::{PackF32ToU8,PackTileID} (gs_train.cpp:94-96, editor.cpp:1581);
 *   - the nine C symbols resolved at gs_train.cpp:24-179 (signatures from the typedefs there);
 *   - the GaussianTrainerScene methods the editor calls (editor.cpp:846-855,1426-1654,2023-2045).
 * Defaults are the CLI defaults (application/diverseshot-cli/source/main.cpp:12-70).
 *
 * ABI note: create_splat / load_train_data pass C++ objects by reference across the .so boundary
 * (gs_train.cpp:105-110), so the plugin and its caller must be built against this same header and the
 * same libstdc++ ABI.
 */
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

// The editor-facing methods speak glm (editor.cpp:852-856 passes their results to glm::translate / glm::mat4_cast /
// maths::Frustum).  They are INLINE wrappers over glm-free exported methods, present only when glm is on the include
// path (it is for both reference targets: ${GLM_INCLUDE_DIR}, application/diverseshot-cli/CMakeLists.txt:42), so the
// class layout and the exported symbols are the same with or without it.
#if defined(__has_include) && !defined(GSTRAIN_NO_GLM)
#if __has_include(<glm/glm.hpp>) && __has_include(<glm/gtc/quaternion.hpp>)
#include <glm/glm.hpp>
#include <glm/gtc/quaternion.hpp>
#include <glm/gtc/type_ptr.hpp>
#define GSTRAIN_HAS_GLM 1
#endif
#endif

// The editor links the trainer instead of dlopen-ing it (application/editor/source/editor.cpp:846-855,1426-1654 call the
// class directly), so the class is exported; the CLI only needs the nine C symbols at the bottom.
#if defined(_WIN32)
#define GSTRAIN_API
#else
#define GSTRAIN_API __attribute__((visibility("default")))
#endif

enum GSPackLevel : int { PackNone = 0, PackF32ToU8 = 1, PackTileID = 2 };

struct GaussianTrainConfig {
    // paths
    std::string sourcePath, modelPath, cameraPosePath, pointCloudPath;
    // schedule (main.cpp:19-33)
    int numIters = 30000;
    int warmupLength = 500;
    int refineEvery = 100;
    int resetAlphaEvery = 3000;
    int refineStopIter = 15000;
    int refineScale2dStopIter = 0;
    int modelType = 0;        // 0 = 3DGS, 1 = 2DGS
    int densifyStrategy = 1;  // 0 = ADC, 1 = MCMC, 2 = ADC+
    int pruneInterval = 0;
    int pruneStrategy = 0;
    int maxImageWidth = 2048, maxImageHeight = 2048, maxImageCount = 0;
    int capMax = 3000000;
    int packLevel = PackF32ToU8;
    int meshResolution = 0, resolutionSchedule = 0, datasetType = 0, cameraModel = 0, quality = 0;
    int mapperType = 0, videoStrategy = 0, videoFps = 0;
    // learning rates / thresholds (3DGS defaults where the CLI leaves them unset)
    float growGrad2d = 2e-4f;
    float ssimWeight = 0.2f;
    float noiselr = 1e5f;
    float poslrInit = 1.6e-4f, poslrFinal = 1.6e-6f;
    float rotationlr = 1e-3f, scalinglr = 5e-3f, featurelr = 2.5e-3f, opacitylr = 5e-2f;
    float min_opacity = 0.005f, pruneOpacity = 0.005f, pruneScale3d = 0.1f, pruneScale2d = 0.15f;
    // switches
    bool revisedOpacity = false, progressiveTrain = false, useAbsGrad = true, pixelGradScale = false;
    bool verbose = false, mipAntiliased = false, exportMesh = false, useMask = false;
    bool normalConsistencyLoss = false, bestQuality = false, visibleAdam = false, enableBg = false;
    bool enableFocusRegion = false, cullSH = false, singleCamera = false, outputSparsePoints = false;
};

enum class TrainingStatus : int {
    Loading_Prepare = 0,
    Loading_Data,
    Colmap_Sfm,
    Preprocess_Done,
    Training,
    Training_Done,
    GS2Mesh,
    Loading_Failed,
};

// Trainer -> viewer hand-off (include/dvs_viewer_pack.h).  The three buffers that
// GaussianModel::create_gpu_buffer (diverse/source/assets/gaussian_model.cpp:115-212) quantises on the CPU from the six
// getGaussian*Cpu() vectors, produced on the device instead and delivered in pinned host memory, byte-identical:
// a maintainer copies them straight into the mapped gaussians_buf / gaussians_sh_0_buf / gaussians_sh_n_buf.
struct GaussianViewerPack {
    const void* gaussians = nullptr;  // [count] x 32 B  Gaussian           (gaussian_model.h:46-50)
    const void* colors = nullptr;     // [count] x  8 B  PackedVertexColor  (gaussian_model.h:64)
    const void* sh = nullptr;         // [count] x 64 B  PackedVertexSH     (gaussian_model.h:52-58)
    int64_t count = 0;
    float bboxMin[3] = {0, 0, 0}, bboxMax[3] = {0, 0, 0};  // local_bounding_box (gaussian_model.cpp:292-299)
    int iteration = -1;               // training iteration the snapshot was taken after
};

// Three floats that convert from and to any vec3-like type (x, y, z members; V(x, y, z) constructor) — glm::vec3 in the
// editor: `gs_train.focus_region_position = glm::vec3(0.0f)` (editor.cpp:1486), `glm::vec3 p = gsTrain->focus_region_position`
// (inspector_panel.cpp:909), `gs->updateFocusRegion(focus_pos, focus_rot, focus_scale)` (:933).
struct GsVec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    GsVec3() = default;
    GsVec3(float a, float b, float c) : x(a), y(b), z(c) {}
    template <class V, class = decltype(std::declval<const V&>().x + std::declval<const V&>().y + std::declval<const V&>().z)>
    GsVec3(const V& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
    template <class V, class = decltype(V(0.f, 0.f, 0.f).x)>
    operator V() const { return V(x, y, z); }
};
// One SfM / initialisation point: 16 bytes, the record GaussianModel::update_from_pos_color reads
// (gaussian_model.cpp:70-95: three floats, then r, g, b bytes); editor.cpp:1523-1527 casts getPoints3D(0).data() to it.
struct GsPoint3D {
    float x, y, z;
    uint8_t r, g, b, a;
};
// A training image for the dataset panel (img2d_dataset_panel.cpp:113-115: width, height, data as R8G8B8A8, name).
struct GsImageView {
    int width = 0, height = 0;
    const uint8_t* data = nullptr;
    std::string name;
};

struct GaussianTrainerImpl;  // B200 rasterizer context + device-resident parameters (gstrain.cu)

class GSTRAIN_API GaussianTrainerScene {
public:
    GaussianTrainerScene(const GaussianTrainConfig& config, int loadItr);
    ~GaussianTrainerScene();
    GaussianTrainerScene(const GaussianTrainerScene&) = delete;
    GaussianTrainerScene& operator=(const GaussianTrainerScene&) = delete;
    // the editor keeps the trainer as an entt component (editor.cpp:2024 add_component<GaussianTrainerScene>(config, -1)),
    // which needs a movable type
    GaussianTrainerScene(GaussianTrainerScene&& o) noexcept;
    GaussianTrainerScene& operator=(GaussianTrainerScene&& o) noexcept;

    bool loadTrainData(const std::string& path);
    void trainSetup();
    void trainStep();
    void startTrain() { train_ = true; }
    void pauseTrain() { train_ = false; }
    bool& isTrain() { return train_; }
    bool isTerminate() const { return terminate_; }
    bool isPruningSplat() const { return false; }
    void saveGaussianModel();
    void exportMesh(const std::string& path);
    void setModelPath(const std::string& p) { config_.modelPath = p; }
    void setTrainingStatus(TrainingStatus s) { status_ = s; }
    TrainingStatus getCurrentTrainingStatus() const { return status_; }
    int getCurrentIterations() const { return curIteration; }
    int& maxIteriaons() { return config_.numIters; }
    float getCurrentLoss() const { return loss_; }
    GaussianTrainConfig& getTrainConfig() { return config_; }
    int64_t getNumGaussians() const;
    // CPU copies of the stored (raw) parameters, layouts of gaussian_model.cpp:43-68
    std::vector<float> getGaussianPositionCpu() const;
    std::vector<float> getGaussianSH0Cpu() const;
    std::vector<float> getGaussianSHNCpu() const;
    std::vector<float> getGaussianOpcaitiesCpu() const;
    std::vector<float> getGaussianScalingsCpu() const;
    std::vector<float> getGaussianRotationsCpu() const;
    // Viewer hand-off without the 236 B/Gaussian round trip and the CPU quantisation pass (see GaussianViewerPack).
    // requestViewerPack queues the pack kernel behind the training work already queued and a device->host copy on a
    // side stream, and returns at once; acquireViewerPack hands out the newest finished snapshot (wait = true blocks
    // for the newest requested one).  The pointers stay valid until the next-but-one request.
    void requestViewerPack();
    bool acquireViewerPack(GaussianViewerPack& out, bool wait);
    int getNumCameras() const;
    std::array<float, 16> getCameraProjectionFlat(int i) const;  // the perspective matrix alone, flat [4c+r] (column-major)
    std::array<float, 16> getCameraView(int i) const;            // world -> camera, flat [4c+r]
    void getCameraRotationWXYZ(int i, float q[4]) const;         // camera -> world rotation as a unit quaternion
    void getCameraPosXYZ(int i, float p[3]) const;               // camera centre in world space

    // ---- the rest of the surface the editor calls (SURVEY.md §8-B "methods used by the editor")
    bool resumeFromModelFile(const std::string& path);  // what create_splat(config, loadItr >= 0) does after load_train_data
    void resetGaussian();                    // back to the initial parameters, optimizer state and iteration 0 (inspector_panel.cpp:837,1017)
    void setDensifyStrategy(int strategy);   // 0 ADC, 1 MCMC, 2 ADC+ (inspector_panel.cpp:789); statistics restart
    float getProgressOnCurrentPhase() const;               // 0..1 (scene_view_panel.cpp:1022)
    std::string getCurrentTrainingPhaseName() const;       // scene_view_panel.cpp:1056, inspector_panel.cpp:997
    float getTrainingElpasedTime() const;                  // seconds spent in trainStep (inspector_panel.cpp:998)
    float getEstimateTrainingTime() const;                 // seconds remaining at the current rate
    // editor -> trainer: replace the model by edited host arrays (raw parameters, the getters' layouts); the optimizer
    // state restarts.  n may differ from the current count (inspector_panel.cpp:1037-1044 after splat editing).
    void updateTensorFromHost(const float* pos, const float* rot, const float* scale, const float* opacity, const float* sh0,
                              const float* shn, int64_t n);
    const std::vector<GsPoint3D>& getPoints3D(int which) const;   // the initialisation point cloud (editor.cpp:1523)
    GsImageView getSplatImageView(int id);                        // training image `id` as RGBA8 (img2d_dataset_panel.cpp:113)
    bool saveCameraDatas(const std::string& jsonPath) const;      // editor.cpp:3512: cameras.json (id, img_name, width, height, position, rotation, fx, fy)
    bool exportSparsePointCloud(const std::string& plyPath) const;  // editor.cpp:3535: x y z + red green blue
    void updateFocusRegion(const GsVec3& position, const GsVec3& rotationDegrees, const GsVec3& scale);
    void getFocusRegionMinMax(float mn[3], float mx[3]) const;   // the region's box before its transform (the points' bounding box)
    void getFocusRegionTransformFlat(float m[16]) const;          // T * Rz * Ry * Rx * S, flat [4c+r]

#ifdef GSTRAIN_HAS_GLM
    glm::mat4 getCameraProjection(int i) const { return glm::make_mat4(getCameraProjectionFlat(i).data()); }  // editor.cpp:852
    glm::quat getCameraRotation(int i) const { float q[4]; getCameraRotationWXYZ(i, q); return glm::quat(q[0], q[1], q[2], q[3]); }  // :854
    glm::vec3 getCameraPos(int i) const { float p[3]; getCameraPosXYZ(i, p); return glm::vec3(p[0], p[1], p[2]); }  // :855
    std::pair<glm::vec3, glm::vec3> getFocusRegion() const {  // scene_view_panel.cpp:1393
        float a[3], b[3]; getFocusRegionMinMax(a, b);
        return {glm::vec3(a[0], a[1], a[2]), glm::vec3(b[0], b[1], b[2])};
    }
    glm::mat4 getFocusRegionTransform() const { float m[16]; getFocusRegionTransformFlat(m); return glm::make_mat4(m); }  // :1394
    // inspector_panel.cpp:1037-1044: the GaussianModel accessors' types (gaussian_model.h:88-95,134-139)
    void updateTensorFromGaussianData(const std::vector<glm::vec3>& pos, const std::vector<glm::vec4>& rot,
                                      const std::vector<glm::vec3>& scale, const std::vector<float>& opacity,
                                      const std::vector<std::array<float, 3>>& sh0, const std::vector<std::array<float, 45>>& shn) {
        static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::vec4) == 16, "packed glm vectors");
        updateTensorFromHost(pos.empty() ? nullptr : &pos[0].x, rot.empty() ? nullptr : &rot[0].x, scale.empty() ? nullptr : &scale[0].x,
                             opacity.data(), sh0.empty() ? nullptr : sh0[0].data(), shn.empty() ? nullptr : shn[0].data(), (int64_t)pos.size());
    }
#else
    std::array<float, 16> getCameraProjection(int i) const { return getCameraProjectionFlat(i); }
#endif

    bool ShowTrainView = false;
    int curIteration = 0;
    std::vector<int> pruenIteraions;
    GsVec3 focus_region_position{0.f, 0.f, 0.f}, focus_region_rotation{0.f, 0.f, 0.f}, focus_region_scale{1.f, 1.f, 1.f};

private:
    int64_t plannedCapacity(int64_t N) const;  // arena size: capMax when the schedule reaches the refinement window
    GaussianTrainConfig config_;
    TrainingStatus status_ = TrainingStatus::Loading_Prepare;
    bool train_ = true, terminate_ = false;
    float loss_ = 0.f;
    GaussianTrainerImpl* impl_ = nullptr;
};

// The nine symbols the unmodified CLI resolves with dlsym (gs_train.cpp:24,105-110,144-150,178-179).
extern "C" {
void gstrain_init();
void* create_splat(const GaussianTrainConfig& config, int loadItr);
bool load_train_data(GaussianTrainerScene* scene, const std::string& path);
void train_step(GaussianTrainerScene* scene);
void save_splat_model(GaussianTrainerScene* scene);
void export_mesh(GaussianTrainerScene* scene);
void delete_splat(GaussianTrainerScene* scene);
int get_cur_step(GaussianTrainerScene* scene);
void gstrain_destroy();
}
