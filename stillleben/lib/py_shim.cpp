// pybind11 module over the compiled C++ surface (include/stillleben_shim.hpp): the binding layer of the reference
// (python/src/bridge.cpp:23-42 and py_context / py_mesh / py_object / py_scene / py_render_pass.cpp) re-created over the
// shim's classes with the reference's Python names and argument meanings. Matrices cross the boundary as 4x4 row-major
// numpy arrays (m[row, col], as the reference's toTorch<Matrix4>), results as numpy arrays of the reference's shapes / dtypes.
// Built by stillleben/lib/Makefile into stillleben/lib/libstillleben_python<ext suffix>; tests/test_gpu_cpp_shim.py renders the
// same scene through this module and through the ctypes path and requires byte-identical targets.
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stillleben_shim.hpp>

namespace py = pybind11;

static std::shared_ptr<sl::Context> g_context;

static std::shared_ptr<sl::Context> context() {
    if (!g_context) throw std::logic_error("Call sl::init() first");          // py_context.cpp:69-75
    return g_context;
}
static sl::Matrix4 to_mat(const py::array_t<float, py::array::c_style | py::array::forcecast>& a) {
    if (a.ndim() != 2 || a.shape(0) != 4 || a.shape(1) != 4) throw std::invalid_argument("expected a 4x4 matrix");
    sl::Matrix4 m;
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m.at(r, c) = a.at(r, c);
    return m;
}
static py::array_t<float> from_mat(const sl::Matrix4& m) {
    py::array_t<float> a({4, 4});
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) a.mutable_at(r, c) = m.at(r, c);
    return a;
}
static sl::Vector3 to_vec3(const py::array_t<float, py::array::c_style | py::array::forcecast>& a) {
    if (a.size() != 3) throw std::invalid_argument("expected three values");
    return {a.data()[0], a.data()[1], a.data()[2]};
}
template <class T, int C>
static py::array_t<T> image(const sl::RenderPass::Target<T, C>& t, int channels_out) {
    const std::vector<T> host = t.image();
    py::array_t<T> a({t.H, t.W, channels_out});
    T* dst = a.mutable_data();
    for (size_t p = 0; p < (size_t)t.W * t.H; ++p)
        for (int c = 0; c < channels_out; ++c) dst[p * channels_out + c] = host[p * C + c];
    return a;
}

PYBIND11_MODULE(libstillleben_python, m) {
    m.doc() = "stillleben render path over libslb.so (pybind11 binding of include/stillleben_shim.hpp)";
    m.def("init", []() { if (!g_context) g_context = sl::Context::Create(); if (!g_context) throw std::runtime_error(slb_last_error(nullptr)); });
    m.def("init_cuda", [](unsigned int device, bool) { if (!g_context) g_context = sl::Context::CreateCUDA(device); if (!g_context) throw std::runtime_error(slb_last_error(nullptr)); },
          py::arg("device_index") = 0, py::arg("use_cuda") = true);
    m.def("_shutdown", []() { g_context.reset(); });

    py::class_<sl::Mesh, std::shared_ptr<sl::Mesh>>(m, "Mesh")
        // (the reference imports files with Assimp / Magnum; here the caller hands in the consolidated stream: 68-byte vertices,
        //  u32 indices, sub-meshes (index_offset, index_count, material), materials as (base_color[4], metallic, roughness))
        .def_static("from_data", [](py::array_t<uint8_t, py::array::c_style | py::array::forcecast> vertices,
                                    py::array_t<uint32_t, py::array::c_style | py::array::forcecast> indices,
                                    std::vector<std::tuple<uint32_t, uint32_t, int>> submeshes,
                                    std::vector<std::tuple<std::array<float, 4>, float, float>> materials,
                                    std::array<float, 3> bbox_min, std::array<float, 3> bbox_max) {
            sl::MeshData d;
            d.vertices.assign(vertices.data(), vertices.data() + vertices.size());
            d.indices.assign(indices.data(), indices.data() + indices.size());
            for (auto& s : submeshes) d.submeshes.push_back({std::get<0>(s), std::get<1>(s), std::get<2>(s), 0u});
            for (auto& mt : materials) {
                slb_material sm{};
                for (int k = 0; k < 4; ++k) sm.base_color[k] = std::get<0>(mt)[k];
                sm.metallic = std::get<1>(mt); sm.roughness = std::get<2>(mt);
                sm.tex_base_color = sm.tex_normal = sm.tex_metallic_roughness = sm.tex_emissive = sm.tex_occlusion = -1;
                d.materials.push_back(sm);
            }
            d.bbox = {{bbox_min[0], bbox_min[1], bbox_min[2]}, {bbox_max[0], bbox_max[1], bbox_max[2]}};
            return sl::Mesh::fromData(context(), std::move(d));
        })
        .def("center_bbox", &sl::Mesh::centerBBox)
        .def("scale_to_bbox_diagonal", [](sl::Mesh& mesh, float diag, const std::string& mode) {
            if (mode == "exact") mesh.scaleToBBoxDiagonal(diag, sl::Mesh::Scale::Exact);
            else if (mode == "order_of_magnitude") mesh.scaleToBBoxDiagonal(diag, sl::Mesh::Scale::OrderOfMagnitude);
            else throw std::invalid_argument("invalid value for mode argument");                  // py_mesh.cpp:398-407
        }, py::arg("target_diagonal"), py::arg("mode") = "exact")
        .def_property_readonly("pretransform", [](sl::Mesh& mesh) { return from_mat(mesh.pretransform()); })
        .def_property("class_index", &sl::Mesh::classIndex, &sl::Mesh::setClassIndex);

    py::class_<sl::Object, std::shared_ptr<sl::Object>>(m, "Object")
        .def(py::init([](const std::shared_ptr<sl::Mesh>& mesh) { auto o = std::make_shared<sl::Object>(); o->setMesh(mesh); return o; }))
        .def("pose", [](sl::Object& o) { return from_mat(o.pose()); })
        .def("set_pose", [](sl::Object& o, py::array_t<float, py::array::c_style | py::array::forcecast> p) { o.setPose(to_mat(p)); })
        .def_property("instance_index", &sl::Object::instanceIndex, &sl::Object::setInstanceIndex)
        .def_property("metallic", &sl::Object::metallic, &sl::Object::setMetallic)
        .def_property("roughness", &sl::Object::roughness, &sl::Object::setRoughness)
        .def_property("casts_shadows", &sl::Object::castsShadows, &sl::Object::setCastsShadows)
        .def_property_readonly("mesh", &sl::Object::mesh);

    py::class_<sl::Scene, std::shared_ptr<sl::Scene>>(m, "Scene")
        .def(py::init([](std::tuple<int, int> viewport) { return std::make_shared<sl::Scene>(context(), sl::ViewportSize(std::get<0>(viewport), std::get<1>(viewport))); }),
             py::arg("viewport_size"))
        .def("add_object", &sl::Scene::addObject)
        .def_property_readonly("objects", &sl::Scene::objects)
        .def("set_camera_pose", [](sl::Scene& s, py::array_t<float, py::array::c_style | py::array::forcecast> p) { s.setCameraPose(to_mat(p)); })
        .def("camera_pose", [](sl::Scene& s) { return from_mat(s.cameraPose()); })
        .def("set_camera_look_at", [](sl::Scene& s, py::array_t<float, py::array::c_style | py::array::forcecast> pos,
                                      py::array_t<float, py::array::c_style | py::array::forcecast> at, std::array<float, 3> up) {
            s.setCameraLookAt(to_vec3(pos), to_vec3(at), {up[0], up[1], up[2]});
        }, py::arg("position"), py::arg("look_at"), py::arg("up") = std::array<float, 3>{0.0f, 0.0f, 1.0f})
        .def("set_camera_intrinsics", &sl::Scene::setCameraIntrinsics)
        .def("set_camera_hfov", &sl::Scene::setCameraFromFOV)
        .def("projection_matrix", [](sl::Scene& s) { return from_mat(s.projectionMatrix()); })
        .def("choose_random_light_direction", &sl::Scene::chooseRandomLightDirection)
        .def_property("light_directions", [](sl::Scene& s) {
            py::array_t<float> a({3, 3});
            for (int i = 0; i < 3; ++i) { a.mutable_at(i, 0) = s.lightDirections()[i].x; a.mutable_at(i, 1) = s.lightDirections()[i].y; a.mutable_at(i, 2) = s.lightDirections()[i].z; }
            return a;
        }, [](sl::Scene& s, py::array_t<float, py::array::c_style | py::array::forcecast> d) {
            if (d.size() != 9) throw std::invalid_argument("expected a 3x3 array");
            std::array<sl::Vector3, 3> v;
            for (int i = 0; i < 3; ++i) v[i] = {d.data()[3 * i], d.data()[3 * i + 1], d.data()[3 * i + 2]};
            s.setLightDirections(v);
        })
        .def_property("manual_exposure", nullptr, &sl::Scene::setManualExposure);

    py::class_<sl::RenderPass::Result, std::shared_ptr<sl::RenderPass::Result>>(m, "RenderPassResult")
        // shapes / dtypes of python/src/py_render_pass.cpp:20-223
        .def("rgb", [](sl::RenderPass::Result& r) { return image(r.rgb, 4); })
        .def("class_index", [](sl::RenderPass::Result& r) { return image(r.classIndex, 1).attr("view")("int16"); })
        .def("instance_index", [](sl::RenderPass::Result& r) { return image(r.instanceIndex, 1).attr("view")("int16"); })
        .def("coordinates", [](sl::RenderPass::Result& r) { return image(r.objectCoordinates, 3); })
        .def("coordDepth", [](sl::RenderPass::Result& r) { return image(r.objectCoordinates, 4); })
        .def("normals", [](sl::RenderPass::Result& r) { return image(r.normals, 4); })
        .def("vertex_indices", [](sl::RenderPass::Result& r) { return image(r.vertexIndex, 3).attr("view")("int32"); })   // kInt, as the reference (torch has no uint32)
        .def("barycentric_coeffs", [](sl::RenderPass::Result& r) { return image(r.barycentricCoeffs, 3); })
        .def("cam_coordinates", [](sl::RenderPass::Result& r) { return image(r.camCoordinates, 4); });

    py::class_<sl::RenderPass>(m, "RenderPass")
        .def(py::init([](const std::string& shading) {
            if (shading != "pbr" && shading != "phong" && shading != "flat") throw std::invalid_argument("unknown shading type specified");   // py_render_pass.cpp:244
            return new sl::RenderPass(shading == "pbr" ? sl::RenderPass::Type::PBR : shading == "phong" ? sl::RenderPass::Type::Phong : sl::RenderPass::Type::Flat);
        }), py::arg("shading") = "pbr")
        .def_property("ssao_enabled", &sl::RenderPass::ssaoEnabled, &sl::RenderPass::setSSAOEnabled)
        .def("render", [](sl::RenderPass& pass, sl::Scene& scene, std::shared_ptr<sl::RenderPass::Result> result, std::shared_ptr<sl::RenderPass::Result> depth_peel,
                          sl::RenderPass::DrawPredicate predicate) {
            return pass.render(scene, result, depth_peel.get(), predicate);
        }, py::arg("scene"), py::arg("result") = nullptr, py::arg("depth_peel") = nullptr, py::arg("predicate") = nullptr);
}
