// Harness around the reference's OWN pure-math host functions of the render path's inputs, cut verbatim out of the reference
// sources at build time (oracle/build_ref.py: build_host -> oracle/_ref/gen/{scene_camera,mesh_pretransform,object_sticker}.inc,
// never committed) and compiled against the reference's GL-less Magnum:
//   src/scene.cpp   Scene::setCameraPose / setCameraLookAt / cameraPose / setCameraIntrinsics / setCameraProjection / setCameraFromFOV
//   src/mesh.cpp    Mesh::centerBBox / scaleToBBoxDiagonal / updatePretransform / setPretransform / bbox
//   src/object.cpp  Object::stickerViewProjection
// The classes below declare exactly the members those bodies touch. TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref.py pins the
// Python host mirror (stillleben_b200/sl.py, desc.py) on these functions.
#include <Corrade/Utility/Debug.h>
#include <Magnum/Magnum.h>
#include <Magnum/Math/Algorithms/Svd.h>
#include <Magnum/Math/Functions.h>
#include <Magnum/Math/Matrix3.h>
#include <Magnum/Math/Matrix4.h>
#include <Magnum/Math/Quaternion.h>
#include <Magnum/Math/Range.h>
#include <Magnum/Math/Vector3.h>

#include <cmath>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <tuple>

using namespace Magnum;

namespace sl {
class Scene {
public:
    void setCameraPose(const Magnum::Matrix4& pose);
    void setCameraLookAt(const Magnum::Vector3& position, const Magnum::Vector3& lookAt, const Magnum::Vector3& up);
    Magnum::Matrix4 cameraPose() const;
    void setCameraIntrinsics(float fx, float fy, float cx, float cy);
    void setCameraProjection(const Magnum::Matrix4& P);
    void setCameraFromFOV(Magnum::Rad fov);
    struct Camera {
        Vector2i vp; Matrix4 P;
        Vector2i viewport() const { return vp; }
        void setProjectionMatrix(const Matrix4& m) { P = m; }
    } camera;
    Camera* m_camera = &camera;
    struct CameraObject {
        Matrix4 T;
        void setTransformation(const Matrix4& m) { T = m; }
        Matrix4 absoluteTransformationMatrix() const { return T; }
    } m_cameraObject;
};
class Mesh {
public:
    enum class Scale { Exact, OrderOfMagnitude };
    void centerBBox();
    void scaleToBBoxDiagonal(float targetDiagonal, Scale mode);
    void updatePretransform();
    void setPretransform(const Magnum::Matrix4& m);
    Magnum::Range3D bbox() const;
    Range3D m_bbox;
    Matrix4 m_pretransformRigid, m_pretransform;
    float m_scale = 1.0f;
};
class Object {
public:
    Magnum::Matrix4 stickerViewProjection() const;
    std::shared_ptr<Mesh> m_mesh;
    Quaternion m_stickerRotation;
};

#include "_ref/gen/scene_camera.inc"
#include "_ref/gen/mesh_pretransform.inc"
#include "_ref/gen/object_sticker.inc"
}  // namespace sl

static void put(const Matrix4& m, float* out) { for (int k = 0; k < 16; ++k) out[k] = m.data()[k]; }   // column-major

extern "C" {
// mode 0: setCameraIntrinsics(a, b, c, d); mode 1: setCameraFromFOV(a radians)
void ref_projection(int W, int H, int mode, float a, float b, float c, float d, float* P_out) {
    sl::Scene s;
    s.camera.vp = {W, H};
    if (mode == 0) s.setCameraIntrinsics(a, b, c, d); else s.setCameraFromFOV(Rad{a});
    put(s.camera.P, P_out);
}
// returns 0, or 1 when setCameraPose throws ("Camera pose is not rigid")
int ref_look_at(const float* position, const float* look_at, const float* up, float* pose_out) {
    sl::Scene s;
    try { s.setCameraLookAt(Vector3::from(position), Vector3::from(look_at), Vector3::from(up)); } catch (const std::invalid_argument&) { return 1; }
    put(s.cameraPose(), pose_out);
    return 0;
}
int ref_set_camera_pose(const float* pose) {
    sl::Scene s;
    try { s.setCameraPose(Matrix4::from(pose)); } catch (const std::invalid_argument&) { return 1; }
    return 0;
}
// centerBBox (if center) then scaleToBBoxDiagonal (mode 1 exact, 2 order of magnitude, 0 skip) on a fresh mesh
void ref_mesh_normalise(const float* bbox_min, const float* bbox_max, int center, int mode, float target, float* pre_out, float* bbox_out) {
    sl::Mesh m;
    m.m_bbox = Range3D{Vector3::from(bbox_min), Vector3::from(bbox_max)};
    if (center) m.centerBBox();
    if (mode) m.scaleToBBoxDiagonal(target, mode == 1 ? sl::Mesh::Scale::Exact : sl::Mesh::Scale::OrderOfMagnitude);
    put(m.m_pretransform, pre_out);
    const Range3D b = m.bbox();
    for (int k = 0; k < 3; ++k) { bbox_out[k] = b.min()[k]; bbox_out[3 + k] = b.max()[k]; }
}
// setPretransform(m): returns 1 when it throws ("Scaling is not uniform")
int ref_mesh_set_pretransform(const float* m16, float* scale_out, float* rigid_out, float* pre_out) {
    sl::Mesh m;
    try { m.setPretransform(Matrix4::from(m16)); } catch (const std::invalid_argument&) { return 1; }
    *scale_out = m.m_scale;
    put(m.m_pretransformRigid, rigid_out);
    put(m.m_pretransform, pre_out);
    return 0;
}
void ref_sticker_projection(const float* bbox_min, const float* bbox_max, const float* pretransform, const float* quat_xyzw, float* out) {
    sl::Object o;
    o.m_mesh = std::make_shared<sl::Mesh>();
    o.m_mesh->m_bbox = Range3D{Vector3::from(bbox_min), Vector3::from(bbox_max)};
    o.m_mesh->m_pretransform = Matrix4::from(pretransform);
    o.m_stickerRotation = Quaternion{{quat_xyzw[0], quat_xyzw[1], quat_xyzw[2]}, quat_xyzw[3]};
    put(o.stickerViewProjection(), out);
}
}
