// stillleben_shim.hpp — the reference's C++ surface for the render path, implemented over the C ABI of slb.h.
//
// A C++ caller written against the reference's headers (include/stillleben/{context,mesh,object,scene,render_pass}.h)
// keeps its code: the same class names, method names, argument meaning and error behaviour, served by libslb.so
// instead of Magnum / OpenGL:
//
//   auto context = sl::Context::CreateCUDA(0);                      // src/context.cpp:411
//   auto mesh = sl::Mesh::fromData(context, meshData);              // (mesh import is not on the render path: the caller
//   mesh->centerBBox(); mesh->scaleToBBoxDiagonal(0.5f);            //  hands in the consolidated 68-byte vertex stream)
//   auto object = std::make_shared<sl::Object>(); object->setMesh(mesh);
//   sl::Scene scene(context, sl::ViewportSize(640, 480));
//   scene.addObject(object); scene.setCameraLookAt({4, 0, 0}, {0, 0, 0});
//   sl::RenderPass pass; auto ret = pass.render(scene);            // src/render_pass.cpp:303-796
//   auto ids = ret->vertexIndex.image();                            // host copy, like CUDATexture / Image2D read-back
//
// Matrices are sl::Matrix4: column-major float[16] exactly like Magnum::Matrix4::data(). Header-only C++17; link with
// stillleben_b200/libslb.so. tests/cpp/shim_client.cpp replays the reference's "vertex indices" test case
// (tests/basic.cpp:375-453) through this header.
#pragma once
#include <slb.h>

#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace sl {

// ---- minimal linear algebra with Magnum's conventions (column-major, m[col][row]) ---------------------------------------
struct Vector2 { float x = 0, y = 0; };
struct Vector3 {
    float x = 0, y = 0, z = 0;
    Vector3() = default;
    Vector3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    Vector3 operator+(const Vector3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vector3 operator-(const Vector3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vector3 operator-() const { return {-x, -y, -z}; }
    Vector3 operator*(float s) const { return {x * s, y * s, z * s}; }
    float dot(const Vector3& o) const { return x * o.x + y * o.y + z * o.z; }
    float length() const { return std::sqrt(dot(*this)); }
    Vector3 normalized() const { const float l = length(); return {x / l, y / l, z / l}; }
    static Vector3 cross(const Vector3& a, const Vector3& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
};
using Color3 = Vector3;
struct Color4 { float r = 1, g = 1, b = 1, a = 1; };
struct Matrix4 {
    float m[16];                                                      // column-major: element (row r, column c) = m[c * 4 + r]
    Matrix4() { std::memset(m, 0, sizeof m); m[0] = m[5] = m[10] = m[15] = 1.0f; }
    float& at(int r, int c) { return m[c * 4 + r]; }
    float at(int r, int c) const { return m[c * 4 + r]; }
    const float* data() const { return m; }
    static Matrix4 translation(const Vector3& t) { Matrix4 r; r.at(0, 3) = t.x; r.at(1, 3) = t.y; r.at(2, 3) = t.z; return r; }
    static Matrix4 scaling(float s) { Matrix4 r; r.at(0, 0) = r.at(1, 1) = r.at(2, 2) = s; return r; }
    Matrix4 operator*(const Matrix4& o) const {
        Matrix4 r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                float s = 0;
                for (int k = 0; k < 4; ++k) s += at(i, k) * o.at(k, j);
                r.at(i, j) = s;
            }
        return r;
    }
    Vector3 transformPoint(const Vector3& p) const {
        const float w = at(3, 0) * p.x + at(3, 1) * p.y + at(3, 2) * p.z + at(3, 3);
        return {(at(0, 0) * p.x + at(0, 1) * p.y + at(0, 2) * p.z + at(0, 3)) / w, (at(1, 0) * p.x + at(1, 1) * p.y + at(1, 2) * p.z + at(1, 3)) / w,
                (at(2, 0) * p.x + at(2, 1) * p.y + at(2, 2) * p.z + at(2, 3)) / w};
    }
    Vector3 translation() const { return {at(0, 3), at(1, 3), at(2, 3)}; }
    Matrix4 invertedRigid() const {
        Matrix4 r;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.at(i, j) = at(j, i);
        const Vector3 t = translation();
        for (int i = 0; i < 3; ++i) r.at(i, 3) = -(r.at(i, 0) * t.x + r.at(i, 1) * t.y + r.at(i, 2) * t.z);
        return r;
    }
};
struct Range3D {
    Vector3 min_, max_;
    Vector3 min() const { return min_; }
    Vector3 max() const { return max_; }
    Vector3 center() const { return (min_ + max_) * 0.5f; }
    Vector3 size() const { return max_ - min_; }
};
struct ViewportSize { int x, y; ViewportSize(int w, int h) : x(w), y(h) {} };

class Exception : public std::runtime_error { using std::runtime_error::runtime_error; };

// ---- context (include/stillleben/context.h) -----------------------------------------------------------------------------
class Context {
public:
    using Ptr = std::shared_ptr<Context>;
    // The reference's Create() brings up GL without CUDA interop; here rendering always runs on a CUDA device.
    static Ptr Create(const std::string& = {}) { return CreateCUDA(0); }
    static Ptr CreateCUDA(unsigned int device, const std::string& = {}) {
        slb_ctx* h = nullptr;
        if (slb_ctx_create((int)device, &h) != SLB_OK) return {};     // context.cpp:411-470 returns an empty pointer on failure
        return Ptr(new Context(h));
    }
    ~Context() { slb_ctx_destroy(m_h); }
    slb_ctx* handle() const { return m_h; }
    int device() const { return slb_ctx_device(m_h); }
    void check(int rc, const char* what) const {
        if (rc != SLB_OK) throw Exception(std::string(what) + ": " + slb_last_error(m_h));
    }
private:
    explicit Context(slb_ctx* h) : m_h(h) {}
    slb_ctx* m_h;
};

// ---- mesh (include/stillleben/mesh.h; import happens outside: the caller supplies the consolidated stream) --------------
struct MeshData {
    std::vector<uint8_t> vertices;                                    // n * 68 bytes (src/mesh_tools/consolidate.cpp:53-61)
    std::vector<uint32_t> indices;
    std::vector<slb_submesh> submeshes;
    std::vector<slb_material> materials;
    Range3D bbox;
    size_t numVertices() const { return vertices.size() / SLB_VERTEX_STRIDE; }
};
class Mesh {
public:
    enum class Scale { Exact, OrderOfMagnitude };
    static std::shared_ptr<Mesh> fromData(const Context::Ptr& ctx, MeshData data, const std::string& name = "memory") {
        auto m = std::shared_ptr<Mesh>(new Mesh(ctx, std::move(data), name));
        return m;
    }
    ~Mesh() { if (m_h) slb_mesh_destroy(m_ctx->handle(), m_h); }
    void load(bool = true, bool = true) {}                            // data is already in memory
    void loadVisual() {                                               // src/mesh.cpp:624-745: upload on first use
        if (m_h) return;
        const float lo[3] = {m_data.bbox.min_.x, m_data.bbox.min_.y, m_data.bbox.min_.z}, hi[3] = {m_data.bbox.max_.x, m_data.bbox.max_.y, m_data.bbox.max_.z};
        m_ctx->check(slb_mesh_upload(m_ctx->handle(), m_data.vertices.data(), (uint32_t)m_data.numVertices(), m_data.indices.data(),
                                     (uint32_t)m_data.indices.size(), m_data.submeshes.data(), (uint32_t)m_data.submeshes.size(),
                                     m_data.materials.data(), (uint32_t)m_data.materials.size(), nullptr, 0, lo, hi, &m_h),
                     "Mesh::loadVisual");
    }
    const Context::Ptr& context() const { return m_ctx; }
    const std::string& filename() const { return m_name; }
    Range3D bbox() const { return {m_pretransform.transformPoint(m_data.bbox.min_), m_pretransform.transformPoint(m_data.bbox.max_)}; }   // mesh.cpp:1075-1081
    void centerBBox() { m_rigid = Matrix4::translation(-m_data.bbox.center()); updatePretransform(); }                                       // mesh.cpp:1000-1005
    void scaleToBBoxDiagonal(float target, Scale mode = Scale::Exact) {                                                                       // mesh.cpp:1007-1027
        const float scale = target / m_data.bbox.size().length();
        m_scale = mode == Scale::Exact ? scale : std::pow(10.0f, std::round(std::log10(scale)));
        updatePretransform();
    }
    const Matrix4& pretransform() const { return m_pretransform; }
    unsigned int classIndex() const { return m_classIndex; }
    void setClassIndex(unsigned int index) {
        if (index > 0xFFFFu) throw std::invalid_argument("Mesh::setClassIndex(): out of range");    // mesh.cpp:1083-1089
        m_classIndex = index;
    }
    size_t numVertices() const { return m_data.numVertices(); }
    const slb_mesh* slbHandle() { loadVisual(); return m_h; }
private:
    Mesh(const Context::Ptr& ctx, MeshData d, std::string name) : m_ctx(ctx), m_data(std::move(d)), m_name(std::move(name)) {}
    void updatePretransform() { m_pretransform = Matrix4::scaling(m_scale) * m_rigid; }               // mesh.cpp:1029-1032
    Context::Ptr m_ctx;
    MeshData m_data;
    std::string m_name;
    slb_mesh* m_h = nullptr;
    Matrix4 m_rigid, m_pretransform;
    float m_scale = 1.0f;
    unsigned int m_classIndex = 1;                                    // mesh.h:300
};

// ---- object (include/stillleben/object.h) -------------------------------------------------------------------------------
class Object {
public:
    void setMesh(const std::shared_ptr<Mesh>& mesh) { m_mesh = mesh; }
    const std::shared_ptr<Mesh>& mesh() const { return m_mesh; }
    void setPose(const Matrix4& pose) { m_pose = pose; }
    const Matrix4& pose() const { return m_pose; }
    void setInstanceIndex(unsigned int index) {
        if (index > 0xFFFFu) throw std::invalid_argument("Object::setInstanceIndex(): out of range");   // object.cpp:376-382
        m_instanceIndex = index;
    }
    unsigned int instanceIndex() const { return m_instanceIndex; }
    void setMetallic(float v) { m_metallic = v; }
    void setRoughness(float v) { m_roughness = v; }
    float metallic() const { return m_metallic; }
    float roughness() const { return m_roughness; }
    void setCastsShadows(bool v) { m_castsShadows = v; }
    bool castsShadows() const { return m_castsShadows; }
private:
    std::shared_ptr<Mesh> m_mesh;
    Matrix4 m_pose;
    unsigned int m_instanceIndex = 0;
    float m_metallic = -1.0f, m_roughness = -1.0f;                    // object.h:277-278: < 0 = the material's value
    bool m_castsShadows = true;
};

// ---- scene (include/stillleben/scene.h) ---------------------------------------------------------------------------------
class Scene {
public:
    Scene(const Context::Ptr& ctx, const ViewportSize& viewport) : m_ctx(ctx), m_viewport(viewport) {
        setCameraFromFOV(58.0f * 3.14159265358979f / 180.0f);         // scene.cpp:138
        m_lightColors[0] = {300.0f, 300.0f, 300.0f};                  // scene.h:225-230
    }
    const Context::Ptr& context() const { return m_ctx; }
    ViewportSize viewport() const { return m_viewport; }
    void setCameraPose(const Matrix4& pose) { m_cameraPose = pose; }
    const Matrix4& cameraPose() const { return m_cameraPose; }
    void setCameraLookAt(const Vector3& position, const Vector3& lookAt, const Vector3& up = {0.0f, 0.0f, 1.0f}) {   // scene.cpp:205-215
        const Vector3 z = (lookAt - position).normalized(), x = Vector3::cross(z, up).normalized(), y = Vector3::cross(z, x).normalized();
        Matrix4 m;
        m.at(0, 0) = x.x; m.at(1, 0) = x.y; m.at(2, 0) = x.z;
        m.at(0, 1) = y.x; m.at(1, 1) = y.y; m.at(2, 1) = y.z;
        m.at(0, 2) = z.x; m.at(1, 2) = z.y; m.at(2, 2) = z.z;
        m.at(0, 3) = position.x; m.at(1, 3) = position.y; m.at(2, 3) = position.z;
        m_cameraPose = m;
    }
    void setCameraIntrinsics(float fx, float fy, float cx, float cy) {                                                 // scene.cpp:222-253
        const float n = 0.1f, f = 10.0f, W = (float)m_viewport.x, H = (float)m_viewport.y;
        const float L = -cx * n / fx, R = (W - cx) * n / fx, T = -cy * n / fy, B = (H - cy) * n / fy;
        Matrix4 P;
        std::memset(P.m, 0, sizeof P.m);
        P.m[0] = 2.0f * n / (R - L);
        P.m[5] = 2.0f * n / (B - T);
        P.m[8] = (R + L) / (L - R); P.m[9] = (T + B) / (T - B); P.m[10] = (f + n) / (f - n); P.m[11] = 1.0f;
        P.m[14] = (2.0f * f * n) / (n - f);
        m_projection = P;
    }
    void setCameraFromFOV(float fovRad) {                                                                              // scene.cpp:260-271
        const float fx = (float)m_viewport.x / (2.0f * std::tan(fovRad / 2.0f));
        setCameraIntrinsics(fx, fx, (float)m_viewport.x / 2.0f, (float)m_viewport.y / 2.0f);
    }
    void setCameraProjection(const Matrix4& P) { m_projection = P; }
    const Matrix4& projectionMatrix() const { return m_projection; }
    void addObject(const std::shared_ptr<Object>& obj) {
        if (obj->instanceIndex() == 0) obj->setInstanceIndex((unsigned int)m_objects.size() + 1);                       // scene.cpp:285-287
        m_objects.push_back(obj);
    }
    const std::vector<std::shared_ptr<Object>>& objects() const { return m_objects; }
    void loadVisual() { for (auto& o : m_objects) o->mesh()->loadVisual(); }
    void setLightDirections(const std::array<Vector3, 3>& d) { m_lightDirections = d; }
    void setLightColors(const std::array<Color3, 3>& c) { m_lightColors = c; }
    const std::array<Vector3, 3>& lightDirections() const { return m_lightDirections; }
    void setAmbientLight(const Color3& c) { m_ambient = c; }
    void chooseRandomLightDirection() {                                                                                // scene.cpp:453-470
        std::normal_distribution<float> n;
        Vector3 r{n(m_rng), -std::abs(n(m_rng)), -std::abs(n(m_rng))};
        r = r.normalized();
        const Matrix4& c = m_cameraPose;
        const Vector3 d{-(c.at(0, 0) * r.x + c.at(0, 1) * r.y + c.at(0, 2) * r.z), -(c.at(1, 0) * r.x + c.at(1, 1) * r.y + c.at(1, 2) * r.z),
                        -(c.at(2, 0) * r.x + c.at(2, 1) * r.y + c.at(2, 2) * r.z)};
        m_lightDirections = {d, Vector3{}, Vector3{}};
    }
    void setBackgroundColor(const Color4& c) { m_backgroundColor = c; }
    void setBackgroundPlanePose(const Matrix4& pose) { m_planePose = pose; }
    void setBackgroundPlaneSize(const Vector2& size) { m_planeSize = size; }
    void setManualExposure(float e) { m_manualExposure = e; }

    // everything RenderPass::render reads, as the C ABI takes it. `objs` owns the object array of the descriptor.
    slb_scene_desc describe(bool ssaoEnabled, const std::function<bool(const std::shared_ptr<Object>&)>& predicate,
                            std::vector<slb_object_desc>& objs) const {
        objs.clear();
        for (auto& obj : m_objects) {
            slb_object_desc o;
            std::memset(&o, 0, sizeof o);
            o.mesh = obj->mesh()->slbHandle();
            std::memcpy(o.pose, obj->pose().data(), 64);
            std::memcpy(o.pretransform, obj->mesh()->pretransform().data(), 64);
            o.class_index = obj->mesh()->classIndex();
            o.instance_index = obj->instanceIndex();
            o.metallic = obj->metallic(); o.roughness = obj->roughness();
            o.casts_shadows = obj->castsShadows() ? 1 : 0;
            o.visible = (!predicate || predicate(obj)) ? 1 : 0;       // evaluated on the host (render_pass.cpp:444,587)
            const Matrix4 I;
            std::memcpy(o.sticker_projection, I.data(), 64);
            objs.push_back(o);
        }
        slb_scene_desc d;
        std::memset(&d, 0, sizeof d);
        d.width = m_viewport.x; d.height = m_viewport.y;
        std::memcpy(d.projection, m_projection.data(), 64);
        const Matrix4 w2c = m_cameraPose.invertedRigid();
        std::memcpy(d.world_to_cam, w2c.data(), 64);
        for (int i = 0; i < 3; ++i) {
            d.light_directions[i][0] = m_lightDirections[i].x; d.light_directions[i][1] = m_lightDirections[i].y; d.light_directions[i][2] = m_lightDirections[i].z;
            d.light_colors[i][0] = m_lightColors[i].x; d.light_colors[i][1] = m_lightColors[i].y; d.light_colors[i][2] = m_lightColors[i].z;
        }
        d.ambient_light[0] = m_ambient.x; d.ambient_light[1] = m_ambient.y; d.ambient_light[2] = m_ambient.z;
        d.background_plane_size[0] = m_planeSize.x; d.background_plane_size[1] = m_planeSize.y;
        std::memcpy(d.background_plane_pose, m_planePose.data(), 64);
        d.manual_exposure = m_manualExposure;
        d.ssao_enabled = ssaoEnabled ? 1 : 0;
        d.objects = objs.data();
        d.n_objects = (int32_t)objs.size();
        return d;
    }
private:
    Context::Ptr m_ctx;
    ViewportSize m_viewport;
    Matrix4 m_cameraPose, m_projection, m_planePose;
    std::vector<std::shared_ptr<Object>> m_objects;
    std::array<Vector3, 3> m_lightDirections{};
    std::array<Color3, 3> m_lightColors{};
    Color3 m_ambient{0.0f, 0.0f, 0.0f};
    Color4 m_backgroundColor;
    Vector2 m_planeSize;
    float m_manualExposure = -1.0f;                                   // scene.h:198: < 0 = auto exposure
    std::mt19937 m_rng{0};
};

// ---- render pass (include/stillleben/render_pass.h:48-150) --------------------------------------------------------------
class RenderPass {
public:
    enum class Type { PBR, Phong, Flat };                             // stored, unused (the reference ignores it too)
    using DrawPredicate = std::function<bool(const std::shared_ptr<Object>&)>;

    // One attachment of the result: a dense device array [H][W][C] (cuda_interop.h:17-69 without map / unmap), plus a host
    // read-back in the layout Magnum's Image2D read-back has.
    template <class T, int C>
    struct Target {
        slb_ctx* ctx = nullptr; slb_result* res = nullptr; int target = 0, W = 0, H = 0;
        const T* devicePointer() const {
            void* ptrs[SLB_NUM_TARGETS]; size_t bpp[SLB_NUM_TARGETS];
            slb_result_ptrs(res, ptrs, bpp);
            return static_cast<const T*>(ptrs[target]);
        }
        std::vector<T> image() const {                               // synchronous host copy, row 0 = GL row 0
            std::vector<T> out((size_t)W * H * C);
            if (slb_result_read(ctx, res, target, 0, 1, out.data(), out.size() * sizeof(T)) != SLB_OK)
                throw Exception(std::string("RenderPass::Result read-back: ") + slb_last_error(ctx));
            return out;
        }
    };
    struct Result {
        Result(const Context::Ptr& c, int W, int H) : ctx(c) {
            c->check(slb_result_create(c->handle(), W, H, 1, SLB_TARGETS_ALL, nullptr, &res), "RenderPass::Result");
            auto bind = [&](auto& t, int id) { t.ctx = c->handle(); t.res = res; t.target = id; t.W = W; t.H = H; };
            bind(rgb, SLB_TARGET_RGB); bind(objectCoordinates, SLB_TARGET_COORD); bind(classIndex, SLB_TARGET_CLASS);
            bind(instanceIndex, SLB_TARGET_INSTANCE); bind(normals, SLB_TARGET_NORMAL); bind(vertexIndex, SLB_TARGET_VERTEX_INDEX);
            bind(barycentricCoeffs, SLB_TARGET_BARY); bind(camCoordinates, SLB_TARGET_CAM_COORD);
            width = W; height = H;
        }
        ~Result() { slb_result_destroy(ctx->handle(), res); }
        Result(const Result&) = delete;
        Context::Ptr ctx;
        slb_result* res = nullptr;
        int width = 0, height = 0;
        Target<uint8_t, 4> rgb;                                       // the reference's attachment names (render_pass.h:52-76)
        Target<float, 4> objectCoordinates;
        Target<uint16_t, 1> classIndex, instanceIndex;
        Target<float, 4> normals;
        Target<uint32_t, 4> vertexIndex;
        Target<float, 4> barycentricCoeffs, camCoordinates;
    };

    explicit RenderPass(Type type = Type::PBR, bool = true) : m_type(type) {}
    void setSSAOEnabled(bool on) { m_ssao = on; }
    bool ssaoEnabled() const { return m_ssao; }

    // RenderPass::render(Scene&, result, depthBufferResult, predicate): src/render_pass.cpp:303
    std::shared_ptr<Result> render(Scene& scene, const std::shared_ptr<Result>& preAllocatedResult = {}, Result* depthBufferResult = nullptr,
                                   const DrawPredicate& predicate = {}) {
        const Context::Ptr& ctx = scene.context();
        scene.loadVisual();
        std::shared_ptr<Result> result = preAllocatedResult ? preAllocatedResult : m_result;   // a second render overwrites the pass's own result
        const ViewportSize vp = scene.viewport();
        if (!result || result->width != vp.x || result->height != vp.y) {
            result = std::make_shared<Result>(ctx, vp.x, vp.y);
            if (!preAllocatedResult) m_result = result;
        }
        std::vector<slb_object_desc> objs;
        const slb_scene_desc desc = scene.describe(m_ssao, predicate, objs);
        ctx->check(slb_render_batch(ctx->handle(), &desc, 1, result->res, 0, depthBufferResult ? depthBufferResult->res : nullptr, nullptr),
                   "RenderPass::render");
        ctx->check(slb_ctx_synchronize(ctx->handle()), "RenderPass::render");   // the reference's render() returns finished textures
        return result;
    }
private:
    Type m_type;
    bool m_ssao = true;                                               // render_pass.h:150
    std::shared_ptr<Result> m_result;
};

}  // namespace sl
