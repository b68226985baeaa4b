// Harness around the reference's OWN pure-math host functions of the render path's inputs, cut verbatim out of the reference
// sources at build time (oracle/build_ref.py: build_host -> oracle/_ref/gen/{scene_camera,mesh_pretransform,object_sticker}.inc,
// never committed) and compiled against the reference's GL-less Magnum:
//   src/scene.cpp   Scene::setCameraPose / setCameraLookAt / cameraPose / setCameraIntrinsics / setCameraProjection / setCameraFromFOV
//   src/mesh.cpp    Mesh::centerBBox / scaleToBBoxDiagonal / updatePretransform / setPretransform / bbox,
//                   Mesh::recomputeNormals, updateVertexPositionsAndColors, setVertexPositions (the vertex-edit path)
//   src/object.cpp  Object::stickerViewProjection
//   src/shaders/ssao_shader.cpp  the SSAO noise / kernel generator of the SSAOShader constructor (std::mt19937{0xdeadbeef}); the
//                   texture calls between the two loops run against a do-nothing GL::Texture2D
//   src/light_map.cpp  the sIBL (.ibl) reader: IBLSpec::load, LightSpec::load and, out of LightMap::load, the addLight lambda +
//                   the Sun / Light1 / Light2 group handling (the texture upload around them is GL and stays out)
// The classes below declare exactly the members those bodies touch. TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref.py pins the
// Python host mirror (stillleben_b200/sl.py, desc.py) on these functions.
#include <Corrade/Utility/Debug.h>
#include <Corrade/Utility/DebugStl.h>
#include <Magnum/Magnum.h>
#include <Magnum/Math/Algorithms/Svd.h>
#include <Magnum/Math/Functions.h>
#include <Magnum/Math/Matrix3.h>
#include <Magnum/Math/Matrix4.h>
#include <Magnum/Math/Quaternion.h>
#include <Magnum/Math/Range.h>
#include <Magnum/Math/Color.h>
#include <Magnum/Math/Vector3.h>
#include <Corrade/Containers/ArrayView.h>
#include <Corrade/Containers/GrowableArray.h>
#include <Corrade/Containers/Optional.h>
#include <Corrade/Utility/Configuration.h>
#include <Corrade/Utility/ConfigurationGroup.h>
#include <Corrade/Utility/String.h>
#include <Magnum/Math/ConfigurationValue.h>
#include <Magnum/Math/Constants.h>
#include <Magnum/Math/Vector2.h>
#include <Magnum/ImageView.h>
#include <Magnum/PixelFormat.h>
#include <Magnum/Sampler.h>
#include <Corrade/Containers/Array.h>
#include <random>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
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
// what Mesh::meshPoints() / meshNormals() / meshColors() / meshFaces() hand out: a view (copies alias the same storage)
template <class T> struct View {
    T* p; std::size_t n;
    std::size_t size() const { return n; }
    T& operator[](std::size_t i) const { return p[i]; }
};
class Mesh {
public:
    enum class Scale { Exact, OrderOfMagnitude };
    // vertex-edit path
    void recomputeNormals();
    void recompileMesh() {}                                   // GL buffer upload in the reference
    void updateVertexPositionsAndColors(const Corrade::Containers::ArrayView<int>& verticesIndex,
                                        const Corrade::Containers::ArrayView<Magnum::Vector3>& positionsUpdate,
                                        const Corrade::Containers::ArrayView<Magnum::Color4>& colorsUpdate);
    void setVertexPositions(const Corrade::Containers::ArrayView<Magnum::Vector3>& newVertices);
    View<Vector3> meshPoints() { return {points.data(), points.size()}; }
    View<Vector3> meshNormals() { return {normals.data(), normals.size()}; }
    View<Color4> meshColors() { return {colors.data(), colors.size()}; }
    View<UnsignedInt> meshFaces() { return {faces.data(), faces.size()}; }
    std::vector<Vector3> points, normals;
    std::vector<Color4> colors;
    std::vector<UnsignedInt> faces;
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
#include "_ref/gen/mesh_normals.inc"
#include "_ref/gen/mesh_vertex_edit.inc"
#include "_ref/gen/object_sticker.inc"
}  // namespace sl

namespace ssaoref {
using namespace Magnum;
namespace GL {
enum class TextureFormat { RGB32F };
struct Wrap2 { Wrap2(std::initializer_list<SamplerWrapping>) {} };
struct Texture2D {   // the calls the constructor makes on its noise texture: accepted; the uploaded texels are kept
    float texels[48] = {};
    Texture2D& setStorage(int, TextureFormat, const Vector2i&) { return *this; }
    Texture2D& setSubImage(int, const Vector2i&, const ImageView2D& image) {
        std::memcpy(texels, image.data().data(), sizeof texels);
        return *this;
    }
    Texture2D& setWrapping(const Wrap2&) { return *this; }
    Texture2D& setMinificationFilter(SamplerFilter, SamplerMipmap) { return *this; }
    Texture2D& setMagnificationFilter(SamplerFilter) { return *this; }
};
}  // namespace GL
struct Generator {
    GL::Texture2D m_noiseTexture;
    Vector3 m_ssaoKernel[64];
    Generator() {
#include "_ref/gen/ssao_tables.inc"
};
void tables(float* noise16x3, float* kernel64x3) {
    Generator g;
    std::memcpy(noise16x3, g.m_noiseTexture.texels, sizeof g.m_noiseTexture.texels);
    for (int i = 0; i < 64; ++i) for (int k = 0; k < 3; ++k) kernel64x3[3 * i + k] = g.m_ssaoKernel[i][k];
}
}  // namespace ssaoref

namespace iblref {
using namespace Corrade;
using namespace Magnum;
#include "_ref/gen/lightmap_specs.inc"
// REFfile / REFgamma / REFmulti of the [Reflection] group and the directional lights, as LightMap::load derives them.
// Returns -1 when the reference's load() would fail before touching GL, else the number of lights.
int lights(const char* path, char* ref_file, int ref_file_cap, float* gamma_multi, float* dirs, float* cols) {
    using namespace Utility;
    Containers::Array<Vector3> m_lightDirections;
    Containers::Array<Color3> m_lightColors;
    Configuration config{path, Configuration::Flag::ReadOnly};
    if (config.isEmpty()) return -1;
    auto reflectionGroup = config.group("Reflection");
    if (!reflectionGroup) return -1;
    auto refSpec = IBLSpec::load(*reflectionGroup, "REF");
    if (!refSpec) return -1;
    std::snprintf(ref_file, ref_file_cap, "%s", refSpec->file.c_str());
    gamma_multi[0] = refSpec->gamma; gamma_multi[1] = refSpec->multiplier;
#include "_ref/gen/lightmap_lights.inc"
    for (std::size_t i = 0; i < m_lightDirections.size(); ++i)
        for (int k = 0; k < 3; ++k) { dirs[3 * i + k] = m_lightDirections[i][k]; cols[3 * i + k] = m_lightColors[i][k]; }
    return (int)m_lightDirections.size();
}
}  // namespace iblref

static void put(const Matrix4& m, float* out) { for (int k = 0; k < 16; ++k) out[k] = m.data()[k]; }   // column-major

extern "C" {
void ref_ssao_tables(float* noise16x3, float* kernel64x3) { ssaoref::tables(noise16x3, kernel64x3); }
int ref_ibl_lights(const char* path, char* ref_file, int ref_file_cap, float* gamma_multi, float* dirs, float* cols) {
    return iblref::lights(path, ref_file, ref_file_cap, gamma_multi, dirs, cols);
}
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
// updateVertexPositionsAndColors (ids one-based; dpos / dcol may be NULL) or, with ids == NULL, setVertexPositions(dpos as the
// new positions). positions / colors are updated in place, normals_out receives the recomputed normals. Returns 1 on exception.
int ref_mesh_edit(float* positions, float* normals_out, float* colors, int n_verts, const unsigned* indices, int n_idx, const int* ids, int n_ids,
                  const float* dpos, const float* dcol) {
    sl::Mesh m;
    m.points.resize(n_verts); m.normals.resize(n_verts); m.colors.resize(n_verts);
    for (int i = 0; i < n_verts; ++i) { m.points[i] = Vector3::from(positions + 3 * i); m.normals[i] = Vector3::from(normals_out + 3 * i); m.colors[i] = Color4::from(colors + 4 * i); }
    m.faces.assign(indices, indices + n_idx);
    try {
        if (ids) {
            std::vector<int> idv(ids, ids + n_ids);
            std::vector<Vector3> dp; std::vector<Color4> dc;
            if (dpos) for (int i = 0; i < n_ids; ++i) dp.push_back(Vector3::from(dpos + 3 * i));
            if (dcol) for (int i = 0; i < n_ids; ++i) dc.push_back(Color4::from(dcol + 4 * i));
            m.updateVertexPositionsAndColors({idv.data(), idv.size()}, {dp.data(), dp.size()}, {dc.data(), dc.size()});
        } else {
            std::vector<Vector3> np;
            for (int i = 0; i < n_ids; ++i) np.push_back(Vector3::from(dpos + 3 * i));
            m.setVertexPositions({np.data(), np.size()});
        }
    } catch (const std::invalid_argument&) { return 1; }
    for (int i = 0; i < n_verts; ++i)
        for (int k = 0; k < 4; ++k) { if (k < 3) { positions[3 * i + k] = m.points[i][k]; normals_out[3 * i + k] = m.normals[i][k]; } colors[4 * i + k] = m.colors[i][k]; }
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
