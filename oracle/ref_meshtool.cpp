// ref_meshtool.cpp — driver around the REFERENCE'S OWN mesh front end (test infrastructure only).
//
// Links, unmodified and from where they lie under /root/reference:
//   src/mesh_tools/consolidate.cpp       sl::consolidateMesh(importer): the 68-byte vertex stream
//   src/mesh_tools/compute_tangents.cpp  tangents when the asset has none
//   contrib/{corrade,magnum,magnum-plugins}  CgltfImporter / StbImageImporter / Magnum Trade + MeshTools (GL-less build,
//                                        oracle/build_magnum.sh)
// and dumps what sl::Mesh::openFile (src/mesh.cpp:203-300) collects: consolidated vertices + indices, one record per
// scene object (index range, material id: object.cpp:110-119), materials resolved with the calls of
// RenderShader::setMaterial (src/shaders/render_shader.cpp:355-415), textures with their sampler state and images.
// The dump is read by oracle/ref_meshdump.py; tests compare stillleben_b200/gltf.py against it byte for byte.
//
//   meshtool <file.gltf|.glb> <out.bin>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

#include <Corrade/Containers/Array.h>
#include <Corrade/Containers/Optional.h>
#include <Corrade/Containers/Pointer.h>
#include <Corrade/Containers/StridedArrayView.h>
#include <Corrade/PluginManager/Manager.h>
#include <Corrade/Utility/Debug.h>
#include <Magnum/ImageView.h>
#include <Magnum/PixelFormat.h>
#include <Magnum/Math/Color.h>
#include <Magnum/Math/Matrix3.h>
#include <Magnum/Trade/AbstractImporter.h>
#include <Magnum/Trade/ImageData.h>
#include <Magnum/Trade/MaterialData.h>
#include <Magnum/Trade/MeshData.h>
#include <Magnum/Trade/MeshObjectData3D.h>
#include <Magnum/Trade/PbrMetallicRoughnessMaterialData.h>
#include <Magnum/Trade/SceneData.h>
#include <Magnum/Trade/TextureData.h>

#include <stillleben/mesh_tools/consolidate.h>

using namespace Magnum;
using namespace Corrade;

static int importPlugins() {
    CORRADE_PLUGIN_IMPORT(CgltfImporter)
    CORRADE_PLUGIN_IMPORT(StbImageImporter)
    CORRADE_PLUGIN_IMPORT(AnyImageImporter)
    return 1;
}
CORRADE_AUTOMATIC_INITIALIZER(importPlugins)

struct Out {
    FILE* f;
    void u32(uint32_t v) { fwrite(&v, 4, 1, f); }
    void i32(int32_t v) { fwrite(&v, 4, 1, f); }
    void f32(float v) { fwrite(&v, 4, 1, f); }
    void bytes(const void* p, size_t n) { fwrite(p, 1, n, f); }
};

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s <file.gltf|.glb> <out.bin>\n", argv[0]); return 2; }
    PluginManager::Manager<Trade::AbstractImporter> manager;
    Containers::Pointer<Trade::AbstractImporter> importer = manager.loadAndInstantiate("CgltfImporter");
    if (!importer || !importer->openFile(argv[1])) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }

    auto consolidated = sl::consolidateMesh(*importer);     // <- the reference's code
    if (!consolidated) { fprintf(stderr, "consolidateMesh failed\n"); return 1; }

    Out o{fopen(argv[2], "wb")};
    if (!o.f) return 1;
    o.u32(0x534c4d31u);   // "SLM1"
    const auto vdata = consolidated->data.vertexData();
    const auto idata = consolidated->data.indexData();
    o.u32((uint32_t)consolidated->vertexStride);
    o.u32((uint32_t)(vdata.size() / consolidated->vertexStride));
    o.u32((uint32_t)(idata.size() / 4));
    o.bytes(vdata.data(), vdata.size());
    o.bytes(idata.data(), idata.size());

    // scene objects in the order Object::populateParts / loadVisual walks them (object.cpp:86-92,110-133): the scene's
    // children depth first; a drawable for every mesh object whose consolidated sub-mesh exists
    std::vector<uint32_t> order;
    {
        const UnsignedInt n = importer->object3DCount();
        std::vector<Containers::Pointer<Trade::ObjectData3D>> objs(n);
        for (UnsignedInt i = 0; i < n; ++i) objs[i] = importer->object3D(i);
        struct Rec { static void go(std::vector<Containers::Pointer<Trade::ObjectData3D>>& objs, UnsignedInt id, std::vector<uint32_t>& out) {
            out.push_back(id);
            if (objs[id]) for (UnsignedInt c : objs[id]->children()) go(objs, c, out);
        } };
        if (importer->defaultScene() != -1) {
            auto scene = importer->scene(importer->defaultScene());
            for (UnsignedInt id : scene->children3D()) Rec::go(objs, id, order);
        } else if (n) order.push_back(0);
        std::vector<int32_t> recs;
        for (uint32_t id : order) {
            if (id >= consolidated->meshes.size() || !consolidated->meshes[id]) continue;
            if (!objs[id] || objs[id]->instanceType() != Trade::ObjectInstanceType3D::Mesh) continue;
            auto* mo = static_cast<const Trade::MeshObjectData3D*>(objs[id].get());
            recs.push_back((int32_t)consolidated->indexOffsets[id]);
            recs.push_back((int32_t)consolidated->meshes[id]->indexCount());
            recs.push_back(mo->material());
        }
        o.u32((uint32_t)(recs.size() / 3));
        o.bytes(recs.data(), recs.size() * 4);
    }

    // materials: the values RenderShader::setMaterial derives (render_shader.cpp:355-415)
    o.u32(importer->materialCount());
    for (UnsignedInt i = 0; i < importer->materialCount(); ++i) {
        auto data = importer->material(i);
        if (!data) { o.u32(0); continue; }
        o.u32(1);
        auto& material = data->as<Trade::PbrMetallicRoughnessMaterialData>();
        Float metallic = 0.04f, roughness = 0.5f;
        if (material.hasAttribute(Trade::MaterialAttribute::MetalnessTexture) | material.hasAttribute(Trade::MaterialAttribute::NoneRoughnessMetallicTexture)) metallic = 1.0f;
        if (material.hasAttribute(Trade::MaterialAttribute::RoughnessTexture) | material.hasAttribute(Trade::MaterialAttribute::NoneRoughnessMetallicTexture)) roughness = 1.0f;
        if (auto m = material.tryAttribute<Float>(Trade::MaterialAttribute::Metalness)) metallic = *m;
        if (auto r = material.tryAttribute<Float>(Trade::MaterialAttribute::Roughness)) roughness = *r;
        Color4 baseColor = material.baseColor();
        Color4 emissive = material.emissiveColor();
        for (int k = 0; k < 4; ++k) o.f32(baseColor.data()[k]);
        for (int k = 0; k < 4; ++k) o.f32(emissive.data()[k]);
        o.f32(metallic); o.f32(roughness);
        int32_t tex[5] = {-1, -1, -1, -1, -1};   // base, normal, metallic-roughness, emissive, occlusion
        if (auto t = data->tryAttribute<UnsignedInt>(Trade::MaterialAttribute::BaseColorTexture)) tex[0] = (int32_t)*t;
        else if (auto t2 = data->tryAttribute<UnsignedInt>(Trade::MaterialAttribute::DiffuseTexture)) tex[0] = (int32_t)*t2;
        if (auto t = data->tryAttribute<UnsignedInt>(Trade::MaterialAttribute::NormalTexture)) tex[1] = (int32_t)*t;
        if (material.hasNoneRoughnessMetallicTexture()) tex[2] = (int32_t)material.roughnessTexture();
        if (auto t = data->tryAttribute<UnsignedInt>(Trade::MaterialAttribute::EmissiveTexture)) tex[3] = (int32_t)*t;
        if (auto t = data->tryAttribute<UnsignedInt>(Trade::MaterialAttribute::OcclusionTexture)) { (void)t; tex[4] = (int32_t)material.occlusionTexture(); }
        o.bytes(tex, sizeof tex);
        Matrix3 tm = material.commonTextureMatrix();
        o.f32((tm - Matrix3{}).toVector().length());
    }

    // textures (mesh.cpp:633-663) and images
    o.u32(importer->textureCount());
    for (UnsignedInt i = 0; i < importer->textureCount(); ++i) {
        auto t = importer->texture(i);
        if (!t) { o.i32(-1); o.u32(0); o.u32(0); o.u32(0); o.u32(0); o.u32(0); continue; }
        o.i32((int32_t)t->image());
        o.u32((uint32_t)t->minificationFilter()); o.u32((uint32_t)t->mipmapFilter()); o.u32((uint32_t)t->magnificationFilter());
        o.u32((uint32_t)t->wrapping().x()); o.u32((uint32_t)t->wrapping().y());
    }
    o.u32(importer->image2DCount());
    for (UnsignedInt i = 0; i < importer->image2DCount(); ++i) {
        auto im = importer->image2D(i);
        int ch = 0;
        if (im && im->format() == PixelFormat::RGB8Unorm) ch = 3;
        else if (im && im->format() == PixelFormat::RGBA8Unorm) ch = 4;
        if (!ch) { o.u32(0); o.u32(0); o.u32(0); continue; }
        o.u32((uint32_t)im->size().x()); o.u32((uint32_t)im->size().y()); o.u32((uint32_t)ch);
        auto px = im->pixels();   // rows x cols x bytes, row 0 = first row of the imported data
        for (std::size_t r = 0; r < px.size()[0]; ++r)
            for (std::size_t c = 0; c < px.size()[1]; ++c) o.bytes(&px[r][c][0], ch);
    }
    fclose(o.f);
    return 0;
}
