// A C++ client of the reference's surface (include/stillleben_shim.hpp over the C ABI): replays the reference's own
// "vertex indices" test case (tests/basic.cpp:375-453) and the checks of its "render" case that do not need a mesh
// importer (tests/basic.cpp:108-261: instance ids in {0, 1}, coverage, class ids, depth range) on a cube built in
// code with the layout the reference's importer + consolidation produce for tests/cube.glb (24 vertices: four per
// face with the face normal, 12 triangles, 68-byte records, one-based vertex ids).
// Built by tests/cpp/Makefile, run on the GPU box by tests/test_gpu_cpp_shim.py. Exit code 0 = every check passed.
#include <stillleben_shim.hpp>

#include <algorithm>
#include <cstdio>
#include <set>

static int g_failed = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("CHECK FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_failed; } } while (0)

#pragma pack(push, 1)
struct Vertex68 { float position[3], uv[2], color[4], tangent[4]; uint32_t vertexIndex; float normal[3]; };
#pragma pack(pop)
static_assert(sizeof(Vertex68) == SLB_VERTEX_STRIDE, "consolidated vertex record (src/mesh_tools/consolidate.cpp:53-61)");

static sl::MeshData makeCube() {
    sl::MeshData d;
    const int axes[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 2, 0}, {1, 0, 2}, {2, 0, 1}, {2, 1, 0}};
    std::vector<Vertex68> v;
    for (int f = 0; f < 6; ++f) {
        const int a = axes[f][0], b = axes[f][1], n = axes[f][2];
        const float sign = (f & 1) ? -1.0f : 1.0f;   // axis pairs ordered so that a x b = +n for even f, -n for odd f
        float corners[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        const uint32_t base = (uint32_t)v.size();
        for (int k = 0; k < 4; ++k) {
            Vertex68 q{};
            q.position[a] = corners[k][0]; q.position[b] = corners[k][1]; q.position[n] = sign;
            q.normal[n] = sign;
            q.uv[0] = 0.5f * (corners[k][0] + 1); q.uv[1] = 0.5f * (corners[k][1] + 1);
            q.color[0] = q.color[1] = q.color[2] = q.color[3] = 1.0f;
            q.tangent[a] = 1.0f; q.tangent[3] = 1.0f;
            q.vertexIndex = (uint32_t)v.size() + 1;
            v.push_back(q);
        }
        const uint32_t quad[6] = {0, 1, 2, 0, 2, 3};
        for (uint32_t i : quad) d.indices.push_back(base + i);
    }
    d.vertices.assign(reinterpret_cast<uint8_t*>(v.data()), reinterpret_cast<uint8_t*>(v.data()) + v.size() * sizeof(Vertex68));
    d.submeshes.push_back({0u, (uint32_t)d.indices.size(), 0, 0u});
    slb_material m{};
    m.base_color[0] = 0.8f; m.base_color[1] = 0.3f; m.base_color[2] = 0.2f; m.base_color[3] = 1.0f;
    m.metallic = 0.04f; m.roughness = 0.5f;
    m.tex_base_color = m.tex_normal = m.tex_metallic_roughness = m.tex_emissive = m.tex_occlusion = -1;
    d.materials.push_back(m);
    d.bbox = {{-1, -1, -1}, {1, 1, 1}};
    return d;
}

int main() {
    auto context = sl::Context::CreateCUDA(0);
    if (!context) { std::printf("no CUDA context: %s\n", slb_last_error(nullptr)); return 2; }

    // ---- tests/basic.cpp:375-453 "vertex indices" -------------------------------------------------------------------
    {
        auto mesh = sl::Mesh::fromData(context, makeCube(), "cube");
        mesh->load();
        sl::Scene scene(context, sl::ViewportSize(640, 480));
        auto object = std::make_shared<sl::Object>();
        object->setMesh(mesh);
        scene.addObject(object);
        scene.setCameraLookAt({4.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f});
        scene.chooseRandomLightDirection();
        sl::RenderPass pass;
        auto ret = pass.render(scene);

        const std::vector<uint32_t> indices = ret->vertexIndex.image();        // RGBA32UI, .w = 0
        CHECK(indices[0] == 0 && indices[1] == 0 && indices[2] == 0);
        std::vector<bool> visible(25, false);
        uint32_t max = 0;
        for (size_t p = 0; p < indices.size() / 4; ++p)
            for (int i = 0; i < 3; ++i) {
                const uint32_t id = indices[4 * p + i];
                CHECK(id <= mesh->numVertices());
                if (id <= 24) visible[id] = true;
                max = std::max(max, id);
            }
        CHECK(max > 0);
        // 0 is always visible (background), and we should see 4 vertices: the +x face
        CHECK(std::count(visible.begin(), visible.end(), true) == 5);

        const std::vector<float> coeffs = ret->barycentricCoeffs.image();
        CHECK(coeffs.size() == indices.size());
        size_t covered = 0;
        for (size_t p = 0; p < indices.size() / 4; ++p) {
            if (indices[4 * p] == 0) continue;
            ++covered;
            CHECK(indices[4 * p] != indices[4 * p + 1]);
            CHECK(indices[4 * p] != indices[4 * p + 2]);
            CHECK(indices[4 * p + 1] != indices[4 * p + 2]);
            const float sum = coeffs[4 * p] + coeffs[4 * p + 1] + coeffs[4 * p + 2];
            CHECK(std::abs(sum - 1.0f) < 1e-5f);
        }
        // the +x face of a 2 m cube seen from 4 m with the default 58 degree lens: a square of ~2 / 3 * fx pixels
        const float fx = 640.0f / (2.0f * std::tan(0.5f * 58.0f * 3.14159265358979f / 180.0f));
        const float side = 2.0f / 3.0f * fx;
        CHECK(std::abs((float)covered - side * side) < 0.02f * side * side);
        std::printf("vertex indices: %zu covered pixels, max id %u\n", covered, max);
    }

    // ---- tests/basic.cpp:108-261 "render": ids, coverage, depth, a second render into the same pass ----------------
    {
        auto mesh = sl::Mesh::fromData(context, makeCube(), "cube");
        mesh->centerBBox();
        mesh->scaleToBBoxDiagonal(0.5f);
        CHECK(std::abs(mesh->bbox().size().length() - 0.5f) < 1e-6f);          // tests/basic.cpp:128-129
        mesh->setClassIndex(7);
        bool threw = false;
        try { mesh->setClassIndex(70000); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);
        sl::Scene scene(context, sl::ViewportSize(640, 480));
        auto object = std::make_shared<sl::Object>();
        object->setMesh(mesh);
        sl::Matrix4 pose = sl::Matrix4::translation({0.0f, 0.0f, 0.5f});       // 0.5 m in front of the camera (camera looks along +z)
        object->setPose(pose);
        scene.addObject(object);
        CHECK(object->instanceIndex() == 1);                                   // tests/basic.cpp:152
        scene.setLightDirections({sl::Vector3{0.3f, 0.2f, 1.0f}.normalized(), sl::Vector3{}, sl::Vector3{}});
        scene.setManualExposure(1.0f);
        sl::RenderPass pass;
        pass.setSSAOEnabled(false);
        auto ret = pass.render(scene);
        const auto inst = ret->instanceIndex.image();
        const auto cls = ret->classIndex.image();
        const auto coord = ret->objectCoordinates.image();
        const auto rgb = ret->rgb.image();
        size_t n_obj = 0;
        float zmin = 1e9f, zmax = -1e9f;
        std::set<uint16_t> ids;
        for (size_t p = 0; p < inst.size(); ++p) {
            ids.insert(inst[p]);
            CHECK((inst[p] == 1) == (cls[p] == 7));
            if (inst[p]) {
                ++n_obj;
                zmin = std::min(zmin, coord[4 * p + 3]); zmax = std::max(zmax, coord[4 * p + 3]);
                CHECK(rgb[4 * p + 3] == 255);
            } else CHECK(coord[4 * p + 3] == 3000.0f);                        // render_pass.cpp:316: invalid-coordinate clear value
        }
        CHECK(ids.size() == 2 && *ids.begin() == 0 && *ids.rbegin() == 1);     // tests/basic.cpp:236-247: instance ids in {0, 1}
        CHECK(n_obj > 150000 && n_obj < 260000);                                // a 0.289 m face at 0.356 m: ~468 x 468 pixels
        const float half = 0.25f / std::sqrt(3.0f);                            // half edge of the scaled cube
        CHECK(std::abs(zmin - (0.5f - half)) < 1e-3f && zmax <= 0.5f + half + 1e-3f);
        auto again = pass.render(scene);                                       // result == nullptr: the pass's own result is reused
        CHECK(again.get() == ret.get());
        auto pre = std::make_shared<sl::RenderPass::Result>(context, 640, 480);
        auto into = pass.render(scene, pre, nullptr, [](const std::shared_ptr<sl::Object>&) { return false; });   // predicate hides the object
        CHECK(into.get() == pre.get());
        const auto hidden = pre->instanceIndex.image();
        CHECK(std::all_of(hidden.begin(), hidden.end(), [](uint16_t v) { return v == 0; }));
        CHECK(pre->rgb.devicePointer() != nullptr && pre->rgb.devicePointer() != ret->rgb.devicePointer());
        std::printf("render: %zu object pixels, depth %.4f .. %.4f\n", n_obj, zmin, zmax);
    }
    std::printf(g_failed ? "FAILED: %d checks\n" : "ALL CHECKS PASSED\n", g_failed);
    return g_failed ? 1 : 0;
}
