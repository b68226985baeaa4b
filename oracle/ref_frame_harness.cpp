// Harness around the reference's OWN frustum-corner / shadow-matrix code (src/render_pass.cpp: computeFrustumCorners,
// computeShadowMapMatrix). oracle/build_ref.py extracts those two functions verbatim from the reference source into
// oracle/_ref/gen/render_pass_frame.inc (generated, never committed) and compiles them here against the reference's own
// GL-less Magnum (oracle/build_magnum.sh) — the rest of render_pass.cpp needs a GL context and stays out. The classes
// below give the extracted text exactly the accessors it calls on sl::Scene / sl::Object / sl::Mesh:
//   scene.camera().projectionMatrix() / cameraMatrix(), scene.objects(), obj->pose(), obj->mesh()->bbox()
// TEST INFRASTRUCTURE ONLY (tests/test_oracle_ref.py pins oracle/orc_render.cpp:frustum_corners / shadow_matrix on it).
#include <Corrade/Containers/StaticArray.h>
#include <Magnum/Magnum.h>
#include <Magnum/Math/Functions.h>
#include <Magnum/Math/Matrix3.h>
#include <Magnum/Math/Matrix4.h>
#include <Magnum/Math/Range.h>
#include <Magnum/Math/Vector3.h>
#include <Magnum/Math/Vector4.h>

#include <limits>
#include <memory>
#include <vector>

using namespace Magnum;

namespace sl {
struct Mesh {
    Matrix4 pretransform;
    Range3D raw;
    // src/mesh.cpp:1075-1081: the two stored corners through the pretransform, not re-sorted
    Range3D bbox() const { return Range3D{pretransform.transformPoint(raw.min()), pretransform.transformPoint(raw.max())}; }
};
struct Object {
    Matrix4 m_pose;
    std::shared_ptr<Mesh> m_mesh;
    Matrix4 pose() const { return m_pose; }
    const std::shared_ptr<Mesh>& mesh() const { return m_mesh; }
};
struct Camera {
    Matrix4 P, V;
    Matrix4 projectionMatrix() const { return P; }
    Matrix4 cameraMatrix() const { return V; }
};
struct Scene {
    Camera cam;
    std::vector<std::shared_ptr<Object>> objs;
    Camera& camera() { return cam; }
    const std::vector<std::shared_ptr<Object>>& objects() const { return objs; }
};
struct RenderPass { enum class Type { Flat, Phong, PBR }; };
}  // namespace sl

using namespace sl;
namespace {
#include "_ref/gen/render_pass_frame.inc"
}

extern "C" void ref_shadow_setup(const float* projection, const float* world_to_cam, int n_objects, const float* poses, const float* pretransforms,
                                 const float* bbox_min, const float* bbox_max, const float* light_dir, float* corners_out, float* shadow_out) {
    Scene scene;
    scene.cam.P = Matrix4::from(projection);        // column-major float[16], as Magnum stores it
    scene.cam.V = Matrix4::from(world_to_cam);
    for (int i = 0; i < n_objects; ++i) {
        auto mesh = std::make_shared<Mesh>();
        mesh->pretransform = Matrix4::from(pretransforms + 16 * i);
        mesh->raw = Range3D{Vector3::from(bbox_min + 3 * i), Vector3::from(bbox_max + 3 * i)};
        auto obj = std::make_shared<Object>();
        obj->m_pose = Matrix4::from(poses + 16 * i);
        obj->m_mesh = mesh;
        scene.objs.push_back(obj);
    }
    FrustumCorners corners = computeFrustumCorners(scene);
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) corners_out[3 * i + k] = corners[i][k];
    const Matrix4 sm = computeShadowMapMatrix(scene, corners, Vector3::from(light_dir));
    for (int k = 0; k < 16; ++k) shadow_out[k] = sm.data()[k];
}
