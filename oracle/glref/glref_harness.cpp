// glref — the reference's GLSL programs run by a REAL OpenGL implementation (Mesa 18.1 llvmpipe, the libGL that ships with
// Nsight Compute in this image) on an off-screen GLX context over a do-nothing Xlib (fake_x11.c). TEST INFRASTRUCTURE ONLY:
// tests/test_gl_ref.py compares the oracle (and through it the CUDA path) with the eight render targets this program writes.
//
// What is the reference's own code here (compiled / fed from the sources where they lie under /root/reference, nothing copied):
//   * src/shaders/render_shader.{glsl,vert,geom,frag}, shadow_shader.{vert,frag}, tone_map_shader.{vert,frag}: read from disk at run
//     time and handed to glShaderSource verbatim, behind the `#version 450` line Magnum's GL::Shader{GL::Version::GL450} emits;
//   * src/shaders/render_shader.cpp: the TextureInput / Uniform enums, the statements that build the #define header, and every setter
//     (setTransformations, setProjectionMatrix, setClassIndex / InstanceIndex, setLightMap, setManualLighting, both setMaterial,
//     setStickerProjection / Range, bind*Texture, setShadowMap) are cut out at build time (oracle/build_ref.py gl -> _ref/gen/*.inc)
//     and compiled below against stand-ins of Magnum's GL classes that forward to the GL entry points;
//   * Magnum's Math / Trade / Primitives (the reference's contrib/, GL-less build of oracle/build_magnum.sh).
// What is restated: the GL call sequence of RenderPass::render (src/render_pass.cpp:303-710 — state, attachments, clear values, draw
// order, pass order; each block below cites its lines), the mesh / texture upload of Mesh::loadVisual (src/mesh.cpp:624-745) and
// Magnum's MeshTools::compile attribute bindings for the plane / quad primitives.
//
// usage: glref <scene dump> <output file> [...]   (format: tests/glref_util.py)
#include <Corrade/Containers/Array.h>
#include <Corrade/Containers/ArrayView.h>
#include <Corrade/Containers/GrowableArray.h>
#include <Corrade/Containers/Optional.h>
#include <Corrade/Utility/Algorithms.h>
#include <Corrade/Utility/Debug.h>
#include <Corrade/Utility/FormatStl.h>
#include <Magnum/Magnum.h>
#include <Magnum/ImageView.h>
#include <Magnum/PixelFormat.h>
#include <Magnum/Sampler.h>
#include <Magnum/Math/Color.h>
#include <Magnum/Math/Matrix3.h>
#include <Magnum/Math/Matrix4.h>
#include <Magnum/Math/Range.h>
#include <Magnum/Math/Functions.h>
#include <Magnum/Primitives/Cube.h>
#include <Magnum/Primitives/Plane.h>
#include <Magnum/Primitives/Square.h>
#include <Magnum/Trade/MaterialData.h>
#include <Magnum/Trade/MeshData.h>
#include <Magnum/Trade/PbrMetallicRoughnessMaterialData.h>

#include <dlfcn.h>
#include <algorithm>
#include <array>
#include <chrono>
#include <random>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "fake_x11.h"
#include "mini_gl.h"

#define GLREF_DEFINE(ret, name, args) ret (*gl##name) args = nullptr;
GLREF_FUNCTIONS(GLREF_DEFINE)
#undef GLREF_DEFINE

static void fail(const std::string& what) { std::fprintf(stderr, "glref: %s\n", what.c_str()); std::exit(2); }
static void check_gl(const char* where) {
    const GLenum e = glGetError();
    if (e != GL_NO_ERROR) { char b[128]; std::snprintf(b, sizeof b, "GL error 0x%x at %s", e, where); fail(b); }
}

// ------------------------------------------------------------------------------------------------------------------------------
// off-screen context
// ------------------------------------------------------------------------------------------------------------------------------
static void create_context() {
    const char* path = std::getenv("GLREF_LIBGL");
    if (!path) fail("GLREF_LIBGL not set");
    void* gl = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!gl) fail(std::string("dlopen: ") + dlerror());
    typedef void* (*GetProc)(const char*);
    GetProc gpa = (GetProc)dlsym(gl, "glXGetProcAddress");
    typedef void** (*ChooseFB)(Display*, int, const int*, int*);
    typedef void* (*CreateAttribs)(Display*, void*, void*, int, const int*);
    typedef int (*MakeCurrent)(Display*, Drawable, void*);
    ChooseFB chooseFB = (ChooseFB)dlsym(gl, "glXChooseFBConfig");
    CreateAttribs createAttribs = (CreateAttribs)gpa("glXCreateContextAttribsARB");
    MakeCurrent makeCurrent = (MakeCurrent)dlsym(gl, "glXMakeCurrent");
    if (!gpa || !chooseFB || !createAttribs || !makeCurrent) fail("GLX entry points missing");
    Display* dpy = XOpenDisplay(nullptr);
    int n = 0;
    const int fb[] = {0x8011 /* GLX_RENDER_TYPE */, 0x1 /* RGBA_BIT */, 0x8010 /* DRAWABLE_TYPE */, 0x1 /* WINDOW_BIT */, 8 /* RED_SIZE */, 8,
                      9, 8, 10, 8, 12 /* DEPTH_SIZE */, 24, 5 /* DOUBLEBUFFER */, 0, 0};
    void** cfg = chooseFB(dpy, 0, fb, &n);
    if (!cfg || !n) fail("no GLX framebuffer configuration");
    // the version the reference asks Magnum for (GL::Version::GL450, core profile); llvmpipe 18.1 implements 3.3 + the extensions the
    // shaders need and is told to advertise 4.5 (MESA_GL_VERSION_OVERRIDE / MESA_GLSL_VERSION_OVERRIDE, set by the caller)
    const int attribs[] = {0x2091 /* MAJOR */, 4, 0x2092 /* MINOR */, 5, 0x9126 /* PROFILE_MASK */, 0x1 /* CORE */, 0};
    void* ctx = createAttribs(dpy, cfg[0], nullptr, 1, attribs);
    if (!ctx) fail("glXCreateContextAttribsARB failed");
    if (!makeCurrent(dpy, 0x500, ctx)) fail("glXMakeCurrent failed");
#define GLREF_LOAD(ret, name, args) gl##name = (ret (*) args)gpa("gl" #name); if (!gl##name) fail("missing gl" #name);
    GLREF_FUNCTIONS(GLREF_LOAD)
#undef GLREF_LOAD
}

// ------------------------------------------------------------------------------------------------------------------------------
// stand-ins of the Magnum GL classes the reference's shader front end touches
// ------------------------------------------------------------------------------------------------------------------------------
namespace Magnum { namespace GL {
struct TextureBase {
    GLuint id = 0; GLenum target = GL_TEXTURE_2D;
    void bind(Int unit) { glActiveTexture(GL_TEXTURE0 + (GLenum)unit); glBindTexture(target, id); }
};
struct Texture2D : TextureBase { Texture2D() { target = GL_TEXTURE_2D; } };
struct RectangleTexture : TextureBase { RectangleTexture() { target = GL_TEXTURE_RECTANGLE; } };
struct Texture2DArray : TextureBase { Texture2DArray() { target = GL_TEXTURE_2D_ARRAY; } };
struct CubeMapTexture : TextureBase { CubeMapTexture() { target = GL_TEXTURE_CUBE_MAP; } };
enum class CubeMapCoordinate : GLenum { PositiveX = 0x8515, NegativeX = 0x8516, PositiveY = 0x8517, NegativeY = 0x8518, PositiveZ = 0x8519, NegativeZ = 0x851A };

// AbstractShaderProgram::setUniform: the overloads the cut code calls (glUniform* on the bound program == Magnum's glProgramUniform*)
class AbstractShaderProgram {
public:
    GLuint id = 0;
    void use() { glUseProgram(id); }
    void setUniform(Int l, const Matrix4& m) { use(); glUniformMatrix4fv(l, 1, GL_FALSE, m.data()); }
    void setUniform(Int l, const Matrix3x3& m) { use(); glUniformMatrix3fv(l, 1, GL_FALSE, m.data()); }
    void setUniform(Int l, const Vector3& v) { use(); glUniform3fv(l, 1, v.data()); }
    void setUniform(Int l, const Vector4& v) { use(); glUniform4fv(l, 1, v.data()); }
    void setUniform(Int l, UnsignedInt v) { use(); glUniform1ui(l, v); }
    void setUniform(Int l, Float v) { use(); glUniform1f(l, v); }
    void setUniform(Int l, const Containers::Array<Vector3>& a) { use(); glUniform3fv(l, (GLsizei)a.size(), a[0].data()); }
    void setUniform(Int l, const Containers::Array<Color3>& a) { use(); glUniform3fv(l, (GLsizei)a.size(), a[0].data()); }
    void setUniform(Int l, const Containers::Array<Vector4>& a) { use(); glUniform4fv(l, (GLsizei)a.size(), a[0].data()); }
    void setUniform(Int l, const Containers::ArrayView<Matrix4>& a) { use(); glUniformMatrix4fv(l, (GLsizei)a.size(), GL_FALSE, a[0].data()); }
};
}  // namespace GL
namespace Shaders {   // attribute locations of Magnum's generic shader interface (contrib/magnum/src/Magnum/Shaders/GenericGL.h:284-379)
struct GenericGL3D {
    struct Position { enum : UnsignedInt { Location = 0 }; };
    struct TextureCoordinates { enum : UnsignedInt { Location = 1 }; };
    struct Color4 { enum : UnsignedInt { Location = 2 }; };
    struct Tangent4 { enum : UnsignedInt { Location = 3 }; };
    struct ObjectId { enum : UnsignedInt { Location = 4 }; };
    struct Normal { enum : UnsignedInt { Location = 5 }; };
};
}  // namespace Shaders
}  // namespace Magnum

using namespace Magnum;

namespace sl {
constexpr Magnum::UnsignedInt NumLights = 3;   // include/stillleben/common.h:17

class LightMap {   // the accessors RenderShader::setLightMap reads (include/stillleben/light_map.h)
public:
    GL::CubeMapTexture cube, irradiance, prefilter;
    GL::Texture2D lut;
    Containers::Array<Vector3> directions;
    Containers::Array<Color3> colors;
    GL::CubeMapTexture& cubeMap() { return cube; }
    GL::CubeMapTexture& irradianceMap() { return irradiance; }
    GL::CubeMapTexture& prefilterMap() { return prefilter; }
    GL::Texture2D& brdfLUT() { return lut; }
    Containers::ArrayView<const Vector3> lightDirections() const { return directions; }
    Containers::ArrayView<const Color3> lightColors() const { return colors; }
};

class MaterialOverride {   // src/shaders/render_shader.h:23-44
public:
    MaterialOverride& metallic(Float m) { _metallic = m; return *this; }
    MaterialOverride& roughness(Float r) { _roughness = r; return *this; }
    constexpr Float metallic() const { return _metallic; }
    constexpr Float roughness() const { return _roughness; }
private:
    Float _metallic = -1.0f, _roughness = -1.0f;
};

class RenderShader : public GL::AbstractShaderProgram {   // the declarations of src/shaders/render_shader.h:46-200 the cut bodies need
public:
    enum : UnsignedInt { ColorOutput = 0, ObjectCoordinatesOutput = 1, ClassIndexOutput = 2, InstanceIndexOutput = 3, NormalOutput = 4,
                         VertexIndexOutput = 5, BarycentricCoeffsOutput = 6, CamCoordinatesOutput = 7 };
    std::string buildHeader();
    RenderShader& bindDepthTexture(GL::RectangleTexture& texture);
    RenderShader& setTransformations(const Matrix4& meshToObject, const Matrix4& objectToWorld, const Matrix4& worldToCam);
    RenderShader& setProjectionMatrix(const Matrix4& projection);
    RenderShader& setClassIndex(unsigned int classIndex);
    RenderShader& setInstanceIndex(unsigned int instanceIndex);
    RenderShader& setLightMap(LightMap& lightMap);
    RenderShader& setManualLighting(const Containers::ArrayView<Vector3>& directions, const Containers::ArrayView<Color3>& colors, const Color3& ambientLight);
    RenderShader& setMaterial(const Trade::MaterialData& material, const Containers::ArrayView<Containers::Optional<GL::Texture2D>>& textures,
                              const MaterialOverride& materialOverride = {});
    RenderShader& setMaterial(const Trade::MaterialData& material, const Containers::ArrayView<GL::Texture2D*>& textures,
                              const MaterialOverride& materialOverride = {});
    RenderShader& setShadowMap(GL::Texture2DArray& shadowMaps, const Containers::ArrayView<Matrix4>& shadowMatrices);
    RenderShader& setStickerProjection(const Matrix4 proj);
    RenderShader& setStickerRange(const Range2D& range);
    RenderShader& bindStickerTexture(GL::RectangleTexture& texture);
};

#include "_ref/gen/render_shader_enums.inc"     // namespace { enum class TextureInput, enum class Uniform, eVal }

std::string RenderShader::buildHeader() {
#include "_ref/gen/render_shader_header.inc"    // std::string header = formatString(...); header += ...  (render_shader.cpp:95-214)
    return header;
}

#include "_ref/gen/render_shader_setters.inc"   // RenderShader::setTransformations ... setShadowMap, and the closing brace of namespace sl

// The SSAO noise texture + hemisphere kernel: the two loops of SSAOShader::SSAOShader (src/shaders/ssao_shader.cpp:72-112, cut out at
// build time); the texture calls between them land on this GL-backed stand-in and create the real 4x4 RGB32F noise texture.
namespace ssaogen {
namespace GL {
enum class TextureFormat { RGB32F };
struct Texture2D : Magnum::GL::Texture2D {
    Texture2D& setStorage(int, TextureFormat, const Vector2i& size) {
        glGenTextures(1, &id); glBindTexture(GL_TEXTURE_2D, id); glTexStorage2D(GL_TEXTURE_2D, 1, GL_RGB32F, size.x(), size.y()); return *this;
    }
    Texture2D& setSubImage(int, const Vector2i&, const ImageView2D& image) {
        glBindTexture(GL_TEXTURE_2D, id); glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
        glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, image.size().x(), image.size().y(), GL_RGB, GL_FLOAT, image.data().data()); return *this;
    }
    Texture2D& setWrapping(std::initializer_list<SamplerWrapping>) {
        glBindTexture(GL_TEXTURE_2D, id); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_REPEAT); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_REPEAT); return *this;
    }
    Texture2D& setMinificationFilter(SamplerFilter, SamplerMipmap) { glBindTexture(GL_TEXTURE_2D, id); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST); return *this; }
    Texture2D& setMagnificationFilter(SamplerFilter) { glBindTexture(GL_TEXTURE_2D, id); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST); return *this; }
};
}  // namespace GL
struct Tables {
    GL::Texture2D m_noiseTexture;
    Vector3 m_ssaoKernel[64];
    Tables() {
#include "_ref/gen/ssao_tables.inc"
};
}  // namespace ssaogen

// CUBE_MAP_SIDES (face + view matrix) and the inside-out cube of LightMap::load (src/light_map.cpp:178-253, cut out at build time)
namespace lightmapgen {
namespace GL { using Magnum::GL::CubeMapCoordinate; }
#include "_ref/gen/lightmap_cube.inc"
}  // namespace lightmapgen

// ------------------------------------------------------------------------------------------------------------------------------
// program construction (what Magnum's GL::Shader / AbstractShaderProgram do for RenderShader::RenderShader)
// ------------------------------------------------------------------------------------------------------------------------------
static std::string g_shader_dir;
// The shader files are compiled into this binary as string literals (oracle/build_ref.py writes _ref/gen/shader_blob.inc from
// /root/reference/src/shaders — what the reference's own build does with corrade-rc), so that oracle/_ref/glref also runs where the
// reference tree does not exist (the GPU box: bench.py --impl reference). GLREF_SHADER_DIR reads them from disk instead.
struct ShaderFile { const char* name; const char* text; };
static const ShaderFile kShaderFiles[] = {
#include "_ref/gen/shader_blob.inc"
};
static std::string read_file(const std::string& name) {
    if (g_shader_dir.empty()) {
        for (const ShaderFile& f : kShaderFiles) if (name == f.name) return f.text;
        fail("shader " + name + " is not compiled in");
    }
    std::ifstream f(g_shader_dir + "/" + name, std::ios::binary);
    if (!f) fail("cannot read " + g_shader_dir + "/" + name);
    std::stringstream ss; ss << f.rdbuf();
    return ss.str();
}
static GLuint compile_stage(GLenum type, const std::vector<std::string>& sources, const char* what) {
    const GLuint s = glCreateShader(type);
    std::vector<const GLchar*> ptrs; std::vector<GLint> lens;
    for (const std::string& src : sources) { ptrs.push_back(src.data()); lens.push_back((GLint)src.size()); }
    glShaderSource(s, (GLsizei)ptrs.size(), ptrs.data(), lens.data());
    glCompileShader(s);
    GLint ok = 0; glGetShaderiv(s, GL_COMPILE_STATUS, &ok);
    if (!ok) { char log[8192]; GLsizei n = 0; glGetShaderInfoLog(s, sizeof log, &n, log); fail(std::string(what) + " does not compile:\n" + log); }
    return s;
}
static GLuint link_program(const std::vector<GLuint>& stages, const char* what) {
    const GLuint p = glCreateProgram();
    for (GLuint s : stages) glAttachShader(p, s);
    glLinkProgram(p);
    GLint ok = 0; glGetProgramiv(p, GL_LINK_STATUS, &ok);
    if (!ok) { char log[8192]; GLsizei n = 0; glGetProgramInfoLog(p, sizeof log, &n, log); fail(std::string(what) + " does not link:\n" + log); }
    return p;
}
extern const std::string kVersion;
const std::string kVersion = "#version 450\n";   // GL::Shader{GL::Version::GL450, ...} (render_shader.cpp:81,92)

// ------------------------------------------------------------------------------------------------------------------------------
// scene dump
// ------------------------------------------------------------------------------------------------------------------------------
struct Reader {
    std::vector<char> buf; size_t at = 0;
    template <class T> T get() { T v; take(&v, sizeof v); return v; }
    void take(void* dst, size_t n) { if (at + n > buf.size()) fail("scene dump truncated"); std::memcpy(dst, buf.data() + at, n); at += n; }
    Matrix4 mat4() { Matrix4 m; take(m.data(), 64); return m; }   // column-major, as Magnum stores it
};
struct TexIn { int w, h, ch, wrap_s, wrap_t, min_f, mag_f, kind; std::vector<unsigned char> px; GLuint id = 0; std::vector<std::vector<unsigned char>> levels; };
struct MatIn { float base[4], emissive[4], metallic, roughness; int tex[5]; };
struct SubIn { int off, count, mat; };
struct MeshIn { int nv, ni; std::vector<char> verts; std::vector<UnsignedInt> idx; std::vector<SubIn> subs; std::vector<MatIn> mats; GLuint vao = 0; };
struct ObjIn { int mesh; Matrix4 pose, pre; int cls, inst; float metallic, roughness; int casts, visible, sticker; Matrix4 sticker_proj; float sticker_range[4]; };

static GLenum wrap_of(int w) { const GLenum t[4] = {GL_REPEAT, GL_CLAMP_TO_EDGE, GL_MIRRORED_REPEAT, GL_CLAMP_TO_BORDER}; return t[w & 3]; }
static GLenum filter_of(int f) {
    const GLenum t[6] = {GL_NEAREST, GL_LINEAR, GL_NEAREST_MIPMAP_NEAREST, GL_LINEAR_MIPMAP_NEAREST, GL_NEAREST_MIPMAP_LINEAR, GL_LINEAR_MIPMAP_LINEAR};
    return t[f % 6];
}

// Mesh::loadVisual texture set-up (src/mesh.cpp:634-665): filters and wrapping of the glTF sampler, maximum anisotropy, full mip chain
// allocated with setStorage, level 0 uploaded, the rest by glGenerateMipmap. Rectangle textures (stickers, background images, depth
// peel input: python/src/py_magnum.cpp / render_pass.cpp:411-419) are single-level.
static void upload_texture(TexIn& t) {
    glGenTextures(1, &t.id);
    glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
    const GLenum fmt = t.ch == 3 ? GL_RGB : GL_RGBA, ifmt = t.ch == 3 ? GL_RGB8 : GL_RGBA8;
    if (t.kind == 1) {
        glBindTexture(GL_TEXTURE_RECTANGLE, t.id);
        if (std::getenv("GLREF_FLOAT_TEXTURES")) {   // see below: the same texels, stored as float, take llvmpipe's float filtering path
            std::vector<float> fl((size_t)t.w * t.h * t.ch);
            for (size_t i = 0; i < fl.size(); ++i) fl[i] = (float)t.px[i] / 255.0f;
            glTexStorage2D(GL_TEXTURE_RECTANGLE, 1, t.ch == 3 ? GL_RGB32F : GL_RGBA32F, t.w, t.h);
            glTexSubImage2D(GL_TEXTURE_RECTANGLE, 0, 0, 0, t.w, t.h, fmt, GL_FLOAT, fl.data());
        } else {
            glTexStorage2D(GL_TEXTURE_RECTANGLE, 1, ifmt, t.w, t.h);
            glTexSubImage2D(GL_TEXTURE_RECTANGLE, 0, 0, 0, t.w, t.h, fmt, GL_UNSIGNED_BYTE, t.px.data());
        }
        // wrapping: GL's default for rectangle textures is clamp-to-edge (sl.Texture(tensor), py_magnum.cpp:147-151); Context::loadTexture
        // (context.cpp:596-598) switches to clamp-to-border with GL's default transparent-black border. REPEAT is not legal on this target.
        if (t.wrap_s == 3) { glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_BORDER); glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_BORDER); }
        glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MIN_FILTER, (GLint)filter_of(t.min_f > 1 ? 1 : t.min_f));
        glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MAG_FILTER, (GLint)filter_of(t.mag_f));
    } else {
        glBindTexture(GL_TEXTURE_2D, t.id);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, (GLint)filter_of(t.mag_f));
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, (GLint)filter_of(t.min_f));
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, (GLint)wrap_of(t.wrap_s));
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, (GLint)wrap_of(t.wrap_t));
        GLfloat aniso = 1.0f; glGetFloatv(GL_MAX_TEXTURE_MAX_ANISOTROPY, &aniso);
        if (glGetError() == GL_NO_ERROR && aniso > 1.0f && !std::getenv("GLREF_NO_ANISO")) glTexParameterf(GL_TEXTURE_2D, GL_TEXTURE_MAX_ANISOTROPY, aniso);
        const int levels = (int)Math::log2((UnsignedInt)std::max(t.w, t.h)) + 1;
        glTexStorage2D(GL_TEXTURE_2D, levels, ifmt, t.w, t.h);
        glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, t.w, t.h, fmt, GL_UNSIGNED_BYTE, t.px.data());
        glGenerateMipmap(GL_TEXTURE_2D);
        // the mip chain the GL implementation generated (8-bit), kept for the output file: tests compare it with the oracle's chain
        glPixelStorei(GL_PACK_ALIGNMENT, 1);
        t.levels.resize(levels);
        for (int l = 0; l < levels; ++l) {
            const int lw = std::max(1, t.w >> l), lh = std::max(1, t.h >> l);
            t.levels[l].resize((size_t)lw * lh * 4);
            glGetTexImage(GL_TEXTURE_2D, l, GL_RGBA, GL_UNSIGNED_BYTE, t.levels[l].data());
        }
        if (std::getenv("GLREF_FLOAT_TEXTURES")) {
            // llvmpipe filters 8-bit textures with 8-bit fixed-point weights (its AoS sampling path): an implementation shortcut, about
            // 4e-3 of a texel. The same texels (the implementation's own 8-bit mip chain, value / 255) in an RGBA32F texture take its
            // float path — same sampler state, same shader, no shortcut.
            GLuint f; glGenTextures(1, &f); glBindTexture(GL_TEXTURE_2D, f);
            glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, (GLint)filter_of(t.mag_f));
            glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, (GLint)filter_of(t.min_f));
            glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, (GLint)wrap_of(t.wrap_s));
            glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, (GLint)wrap_of(t.wrap_t));
            if (aniso > 1.0f && !std::getenv("GLREF_NO_ANISO")) glTexParameterf(GL_TEXTURE_2D, GL_TEXTURE_MAX_ANISOTROPY, aniso);
            glTexStorage2D(GL_TEXTURE_2D, levels, GL_RGBA32F, t.w, t.h);
            for (int l = 0; l < levels; ++l) {
                const int lw = std::max(1, t.w >> l), lh = std::max(1, t.h >> l);
                std::vector<float> fl((size_t)lw * lh * 4);
                for (size_t i = 0; i < fl.size(); ++i) fl[i] = (float)t.levels[l][i] / 255.0f;
                glTexSubImage2D(GL_TEXTURE_2D, l, 0, 0, lw, lh, GL_RGBA, GL_FLOAT, fl.data());
            }
            t.id = f;
        }
    }
    check_gl("texture upload");
}

// Mesh::loadVisual geometry (src/mesh.cpp:671-738): one vertex + one index buffer per mesh file, the 68-byte stream of
// consolidate.cpp:53-61 bound to Magnum's generic attribute locations; sub-meshes are index ranges of it
static void upload_mesh(MeshIn& m) {
    GLuint vbo, ibo;
    glGenVertexArrays(1, &m.vao); glBindVertexArray(m.vao);
    glGenBuffers(1, &vbo); glBindBuffer(GL_ARRAY_BUFFER, vbo);
    glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)m.verts.size(), m.verts.data(), GL_STATIC_DRAW);
    glGenBuffers(1, &ibo); glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, ibo);
    glBufferData(GL_ELEMENT_ARRAY_BUFFER, (GLsizeiptr)(m.idx.size() * 4), m.idx.data(), GL_STATIC_DRAW);
    const GLsizei stride = 68;
    glEnableVertexAttribArray(0); glVertexAttribPointer(0, 3, GL_FLOAT, GL_FALSE, stride, (const void*)0);     // position
    glEnableVertexAttribArray(1); glVertexAttribPointer(1, 2, GL_FLOAT, GL_FALSE, stride, (const void*)12);    // textureCoords
    glEnableVertexAttribArray(2); glVertexAttribPointer(2, 4, GL_FLOAT, GL_FALSE, stride, (const void*)20);    // color
    glEnableVertexAttribArray(3); glVertexAttribPointer(3, 4, GL_FLOAT, GL_FALSE, stride, (const void*)36);    // tangent
    glEnableVertexAttribArray(4); glVertexAttribIPointer(4, 1, GL_UNSIGNED_INT, stride, (const void*)52);       // vertexIndex (ObjectId)
    glEnableVertexAttribArray(5); glVertexAttribPointer(5, 3, GL_FLOAT, GL_FALSE, stride, (const void*)56);    // normal
    glBindVertexArray(0);
    check_gl("mesh upload");
}

// MeshTools::compile of a Magnum primitive (positions [+ normals + texture coordinates], triangle strip, not indexed): only the
// attributes the primitive has are enabled — the others read the GL default generic attribute value (0, 0, 0, 1)
struct PrimGL { GLuint vao = 0; GLsizei count = 0; };
static PrimGL upload_primitive(const Trade::MeshData& d, bool three_d) {
    PrimGL p; p.count = (GLsizei)d.vertexCount();
    if (d.primitive() != MeshPrimitive::TriangleStrip || d.isIndexed()) fail("primitive layout changed");
    GLuint vbo;
    glGenVertexArrays(1, &p.vao); glBindVertexArray(p.vao);
    std::vector<float> data;
    if (three_d) {
        auto pos = d.attribute<Vector3>(Trade::MeshAttribute::Position);
        auto nrm = d.attribute<Vector3>(Trade::MeshAttribute::Normal);
        auto uv = d.attribute<Vector2>(Trade::MeshAttribute::TextureCoordinates);
        for (UnsignedInt i = 0; i < d.vertexCount(); ++i)
            for (float v : {pos[i].x(), pos[i].y(), pos[i].z(), nrm[i].x(), nrm[i].y(), nrm[i].z(), uv[i].x(), uv[i].y()}) data.push_back(v);
        glGenBuffers(1, &vbo); glBindBuffer(GL_ARRAY_BUFFER, vbo);
        glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)(data.size() * 4), data.data(), GL_STATIC_DRAW);
        glEnableVertexAttribArray(0); glVertexAttribPointer(0, 3, GL_FLOAT, GL_FALSE, 32, (const void*)0);
        glEnableVertexAttribArray(5); glVertexAttribPointer(5, 3, GL_FLOAT, GL_FALSE, 32, (const void*)12);
        glEnableVertexAttribArray(1); glVertexAttribPointer(1, 2, GL_FLOAT, GL_FALSE, 32, (const void*)24);
    } else {
        auto pos = d.attribute<Vector2>(Trade::MeshAttribute::Position);
        for (UnsignedInt i = 0; i < d.vertexCount(); ++i) { data.push_back(pos[i].x()); data.push_back(pos[i].y()); }
        glGenBuffers(1, &vbo); glBindBuffer(GL_ARRAY_BUFFER, vbo);
        glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)(data.size() * 4), data.data(), GL_STATIC_DRAW);
        glEnableVertexAttribArray(0); glVertexAttribPointer(0, 2, GL_FLOAT, GL_FALSE, 8, (const void*)0);
    }
    glBindVertexArray(0);
    check_gl("primitive upload");
    return p;
}

// MeshTools::compile of an indexed triangle primitive: positions on attribute 0 (the only attribute the cube programs read)
static PrimGL upload_indexed(const Trade::MeshData& d) {
    PrimGL p;
    const Containers::Array<Vector3> pos = d.positions3DAsArray();
    const Containers::Array<UnsignedInt> idx = d.indicesAsArray();
    p.count = (GLsizei)idx.size();
    GLuint vbo, ibo;
    glGenVertexArrays(1, &p.vao); glBindVertexArray(p.vao);
    glGenBuffers(1, &vbo); glBindBuffer(GL_ARRAY_BUFFER, vbo); glBufferData(GL_ARRAY_BUFFER, (GLsizeiptr)(pos.size() * 12), pos.data(), GL_STATIC_DRAW);
    glGenBuffers(1, &ibo); glBindBuffer(GL_ELEMENT_ARRAY_BUFFER, ibo); glBufferData(GL_ELEMENT_ARRAY_BUFFER, (GLsizeiptr)(idx.size() * 4), idx.data(), GL_STATIC_DRAW);
    glEnableVertexAttribArray(0); glVertexAttribPointer(0, 3, GL_FLOAT, GL_FALSE, 12, (const void*)0);
    glBindVertexArray(0);
    check_gl("indexed primitive upload");
    return p;
}

// LightMap::load from the equirectangular image on (src/light_map.cpp:355-611): equirect -> cube (+ mips), irradiance convolution,
// GGX prefilter (5 levels), BRDF LUT, each a draw of the reference's cubemap_shader_* / brdf_shader programs. Sizes are the caller's
// (the reference's: 512 / 32 / 128 / 512); the sample counts live in the shader text.
struct LightMapIn { int w = 0, h = 0; std::vector<float> eq; int n_lights = 0; float dirs[9], cols[9]; int sizes[4]; };
static std::string read_file(const std::string& name);
static GLuint compile_stage(GLenum type, const std::vector<std::string>& sources, const char* what);
static GLuint link_program(const std::vector<GLuint>& stages, const char* what);
extern const std::string kVersion;
static void cube_params(GLuint id, int levels, int size) {
    glBindTexture(GL_TEXTURE_CUBE_MAP, id);
    glTexStorage2D(GL_TEXTURE_CUBE_MAP, levels, GL_RGBA32F, size, size);
    glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE); glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
    glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_WRAP_R, GL_CLAMP_TO_EDGE);
    glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR); glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
}
static void load_light_map(const LightMapIn& in, sl::LightMap& lm, const PrimGL& texturedPlane) {
    const std::string compat = read_file("compatibility.glsl"), common = read_file("common.glsl");
    GL::Texture2D equirect;                                   // :364-373 (the .hdr branch; loadTexture :166-174 is the same set-up)
    glGenTextures(1, &equirect.id); glBindTexture(GL_TEXTURE_2D, equirect.id);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
    { GLfloat aniso = 1.0f; glGetFloatv(GL_MAX_TEXTURE_MAX_ANISOTROPY, &aniso); if (glGetError() == GL_NO_ERROR && aniso > 1.0f && !std::getenv("GLREF_NO_ANISO")) glTexParameterf(GL_TEXTURE_2D, GL_TEXTURE_MAX_ANISOTROPY, aniso); }
    glTexStorage2D(GL_TEXTURE_2D, (int)Math::log2((UnsignedInt)in.w) + 1, GL_RGB32F, in.w, in.h);
    glPixelStorei(GL_UNPACK_ALIGNMENT, 1);
    glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, in.w, in.h, GL_RGB, GL_FLOAT, in.eq.data());
    glGenerateMipmap(GL_TEXTURE_2D);
    glEnable(GL_TEXTURE_CUBE_MAP_SEAMLESS);                   // :376
    const int envSize = in.sizes[0], irrSize = in.sizes[1], preSize = in.sizes[2], lutSize = in.sizes[3];
    GLuint fb, depth; glGenFramebuffers(1, &fb); glGenRenderbuffers(1, &depth);
    auto target = [&](int size) {                             // framebuffer.setViewport + depthBuffer.setStorage + attach (:380-384, :462-464, ...)
        glBindRenderbuffer(GL_RENDERBUFFER, depth); glRenderbufferStorage(GL_RENDERBUFFER, GL_DEPTH_COMPONENT24, size, size);
        glBindFramebuffer(GL_FRAMEBUFFER, fb); glFramebufferRenderbuffer(GL_FRAMEBUFFER, GL_DEPTH_ATTACHMENT, GL_RENDERBUFFER, depth);
        glViewport(0, 0, size, size);
    };
    const PrimGL cube = upload_indexed(lightmapgen::cubeFromInside());   // :388
    const Matrix4 perspective = Matrix4::perspectiveProjection(Deg{90.0f}, 1.0f, 0.1f, 10.0f);   // :389
    auto cube_program = [&](const char* frag) {               // CubeMapShader::CubeMapShader (cubemap_shader.cpp:41-62); uniforms: projection 0, view 1, roughness 2
        return link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, compat, read_file("cubemap_shader.vert")}, "cubemap_shader.vert"),
                             compile_stage(GL_FRAGMENT_SHADER, {kVersion, compat, read_file(frag)}, frag)}, frag);
    };
    auto draw_faces = [&](GLuint prog, GLuint dst, int level) {
        for (const auto& side : lightmapgen::CUBE_MAP_SIDES) {
            glBindFramebuffer(GL_FRAMEBUFFER, fb);
            glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, (GLenum)side.coordinate, dst, level);
            const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b);
            if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("cube framebuffer incomplete");
            glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
            glUseProgram(prog); glUniformMatrix4fv(1, 1, GL_FALSE, side.view.data());
            glBindVertexArray(cube.vao); glDrawElements(GL_TRIANGLES, cube.count, GL_UNSIGNED_INT, nullptr);
        }
    };
    // equirectangular -> cube map (:391-446)
    glGenTextures(1, &lm.cube.id); cube_params(lm.cube.id, (int)Math::log2((UnsignedInt)envSize) + 1, envSize);
    target(envSize);
    {
        const GLuint prog = cube_program("cubemap_shader_equirectangular.frag");
        glUseProgram(prog); glUniformMatrix4fv(0, 1, GL_FALSE, perspective.data());
        equirect.bind(0);
        draw_faces(prog, lm.cube.id, 0);
        glBindTexture(GL_TEXTURE_CUBE_MAP, lm.cube.id); glGenerateMipmap(GL_TEXTURE_CUBE_MAP);
    }
    // irradiance (:448-509)
    glGenTextures(1, &lm.irradiance.id); cube_params(lm.irradiance.id, (int)Math::log2((UnsignedInt)irrSize) + 1, irrSize);
    target(irrSize);
    {
        const GLuint prog = cube_program("cubemap_shader_irradiance.frag");
        glUseProgram(prog); glUniformMatrix4fv(0, 1, GL_FALSE, perspective.data());
        lm.cube.bind(0);
        draw_faces(prog, lm.irradiance.id, 0);
        glBindTexture(GL_TEXTURE_CUBE_MAP, lm.irradiance.id); glGenerateMipmap(GL_TEXTURE_CUBE_MAP);
    }
    // prefilter (:511-567): MAX_MIP_LEVELS 5, roughness = mip / 4
    glGenTextures(1, &lm.prefilter.id); cube_params(lm.prefilter.id, 5, preSize);
    {
        const GLuint prog = cube_program("cubemap_shader_prefilter.frag");
        glUseProgram(prog); glUniformMatrix4fv(0, 1, GL_FALSE, perspective.data());
        lm.cube.bind(0);
        for (int mip = 0; mip < 5; ++mip) {
            const int mipSize = (int)(preSize * Math::pow(0.5f, (float)mip));
            target(mipSize);
            glUseProgram(prog); glUniform1f(2, (float)mip / 4.0f);
            draw_faces(prog, lm.prefilter.id, mip);
        }
    }
    // BRDF LUT (:569-600): the textured plane primitive drawn with brdf_shader
    glGenTextures(1, &lm.lut.id); glBindTexture(GL_TEXTURE_2D, lm.lut.id);
    glTexStorage2D(GL_TEXTURE_2D, (int)Math::log2((UnsignedInt)lutSize) + 1, GL_RGBA32F, lutSize, lutSize);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    target(lutSize);
    {
        glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, lm.lut.id, 0);
        const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b);
        glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT);
        const GLuint prog = link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, compat, common, read_file("brdf_shader.vert")}, "brdf_shader.vert"),
                                          compile_stage(GL_FRAGMENT_SHADER, {kVersion, compat, read_file("brdf_shader.frag")}, "brdf_shader.frag")}, "brdf shader");
        glUseProgram(prog);
        glBindVertexArray(texturedPlane.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, texturedPlane.count);
        glBindTexture(GL_TEXTURE_2D, lm.lut.id); glGenerateMipmap(GL_TEXTURE_2D);
    }
    lm.directions = Containers::Array<Vector3>{(std::size_t)in.n_lights}; lm.colors = Containers::Array<Color3>{(std::size_t)in.n_lights};
    for (int i = 0; i < in.n_lights; ++i) { lm.directions[i] = Vector3{in.dirs[3 * i], in.dirs[3 * i + 1], in.dirs[3 * i + 2]}; lm.colors[i] = Color3{in.cols[3 * i], in.cols[3 * i + 1], in.cols[3 * i + 2]}; }
    glBindFramebuffer(GL_FRAMEBUFFER, 0);
    check_gl("light map");
}

static GLuint make_rect(GLenum ifmt, int W, int H, bool nearest) {
    GLuint t; glGenTextures(1, &t); glBindTexture(GL_TEXTURE_RECTANGLE, t);
    glTexStorage2D(GL_TEXTURE_RECTANGLE, 1, ifmt, W, H);
    if (nearest) { glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MAG_FILTER, GL_NEAREST); glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MIN_FILTER, GL_NEAREST); }
    return t;
}

static int render_scene(const char* dump_path, const char* out_path, bool first) {
    Reader r;
    {
        std::ifstream f(dump_path, std::ios::binary);
        if (!f) fail("cannot read the scene dump");
        r.buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    if (r.get<UnsignedInt>() != 0x46524C47u || r.get<UnsignedInt>() != 2u) fail("not a glref scene dump (version 2)");
    const int W = r.get<int>(), H = r.get<int>();
    const Matrix4 P = r.mat4(), V = r.mat4();
    Vector3 lightDir[3]; Color3 lightCol[3]; Color3 ambient;
    r.take(lightDir, 36); r.take(lightCol, 36); r.take(ambient.data(), 12);
    int shadowActive[3]; r.take(shadowActive, 12);
    Containers::Array<Matrix4> shadowMatrices{3};
    for (int i = 0; i < 3; ++i) shadowMatrices[i] = r.mat4();
    Vector2 planeSize; r.take(planeSize.data(), 8);
    const Matrix4 planePose = r.mat4();
    const int planeTex = r.get<int>();
    const int backgroundImage = r.get<int>();
    const float manualExposure = r.get<float>();
    const int ssao = r.get<int>();
    const int hasLightMap = r.get<int>();
    LightMapIn lmIn;
    if (hasLightMap) {
        lmIn.w = r.get<int>(); lmIn.h = r.get<int>();
        lmIn.eq.resize((size_t)lmIn.w * lmIn.h * 3); r.take(lmIn.eq.data(), lmIn.eq.size() * 4);
        lmIn.n_lights = r.get<int>(); r.take(lmIn.dirs, 36); r.take(lmIn.cols, 36); r.take(lmIn.sizes, 16);
    }
    const int hasPeel = r.get<int>();
    std::vector<float> peel((size_t)W * H * 4, 0.0f);   // m_zeroMinDepth (render_pass.cpp:411-419)
    if (hasPeel) r.take(peel.data(), peel.size() * 4);
    std::vector<TexIn> tex(r.get<int>());
    for (TexIn& t : tex) { r.take(&t.w, 32); t.px.resize((size_t)t.w * t.h * t.ch); r.take(t.px.data(), t.px.size()); }
    std::vector<MeshIn> meshes(r.get<int>());
    for (MeshIn& m : meshes) {
        m.nv = r.get<int>(); m.ni = r.get<int>(); const int ns = r.get<int>(), nm = r.get<int>();
        m.verts.resize((size_t)m.nv * 68); r.take(m.verts.data(), m.verts.size());
        m.idx.resize(m.ni); r.take(m.idx.data(), (size_t)m.ni * 4);
        m.subs.resize(ns); r.take(m.subs.data(), (size_t)ns * sizeof(SubIn));
        m.mats.resize(nm); r.take(m.mats.data(), (size_t)nm * sizeof(MatIn));
    }
    std::vector<ObjIn> objs(r.get<int>());
    for (ObjIn& o : objs) {
        o.mesh = r.get<int>(); o.pose = r.mat4(); o.pre = r.mat4(); o.cls = r.get<int>(); o.inst = r.get<int>();
        o.metallic = r.get<float>(); o.roughness = r.get<float>(); o.casts = r.get<int>(); o.visible = r.get<int>(); o.sticker = r.get<int>();
        o.sticker_proj = r.mat4(); r.take(o.sticker_range, 16);
    }

    if (first) create_context();
    if (first && std::getenv("GLREF_VERBOSE")) std::fprintf(stderr, "glref: %s | %s | GLSL %s\n", glGetString(GL_RENDERER), glGetString(GL_VERSION), glGetString(GL_SHADING_LANGUAGE_VERSION));
    glEnable(GL_TEXTURE_CUBE_MAP_SEAMLESS);   // Magnum's context set-up (contrib/magnum/src/Magnum/GL/Context.cpp: seamless cube maps on desktop GL)
    for (TexIn& t : tex) upload_texture(t);
    for (MeshIn& m : meshes) upload_mesh(m);
    const PrimGL plane = upload_primitive(Primitives::planeSolid(Primitives::PlaneFlag::TextureCoordinates), true);   // render_pass.cpp:268
    const PrimGL quad = upload_primitive(Primitives::squareSolid(), false);                                          // render_pass.cpp:266
    const PrimGL skyCube = upload_indexed(Primitives::cubeSolid());                                                  // render_pass.cpp:267
    sl::LightMap lightMap;
    if (hasLightMap) load_light_map(lmIn, lightMap, plane);
    if (hasLightMap && std::getenv("GLREF_IBL_LEVEL0")) {   // experiment knob: irradiance map and LUT read at level 0 only (what the oracle does)
        const std::string which = std::getenv("GLREF_IBL_LEVEL0");   // "1" = both, "irr" / "lut" = one of them
        if (which != "lut") { glBindTexture(GL_TEXTURE_CUBE_MAP, lightMap.irradiance.id); glTexParameteri(GL_TEXTURE_CUBE_MAP, GL_TEXTURE_MIN_FILTER, GL_LINEAR); }
        if (which != "irr") { glBindTexture(GL_TEXTURE_2D, lightMap.lut.id); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR); }
    }

    // programs ------------------------------------------------------------------------------------------------------------------
    sl::RenderShader render;
    {
        const std::string header = render.buildHeader(), bridge = read_file("render_shader.glsl");
        render.id = link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, header, bridge, read_file("render_shader.vert")}, "render_shader.vert"),
                                  compile_stage(GL_GEOMETRY_SHADER, {kVersion, header, bridge, read_file("render_shader.geom")}, "render_shader.geom"),
                                  compile_stage(GL_FRAGMENT_SHADER, {kVersion, header, bridge, read_file("render_shader.frag")}, "render_shader.frag")},
                                 "render shader");
    }
    // ShadowShader (shadow_shader.cpp:35-82): header = position attribute location + UNIFORM_TRANSFORMATION 0
    GLuint shadowProg;
    {
        const std::string header = "#define POSITION_ATTRIBUTE_LOCATION 0\n#define UNIFORM_TRANSFORMATION 0\n";
        shadowProg = link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, header, read_file("shadow_shader.vert")}, "shadow_shader.vert"),
                                   compile_stage(GL_FRAGMENT_SHADER, {kVersion, header, read_file("shadow_shader.frag")}, "shadow_shader.frag")}, "shadow shader");
    }
    // ToneMapShader (tone_map_shader.cpp:41-87): texture units Color 0, ObjectLuminance 1; uniform ManualExposure 0
    GLuint toneProg;
    {
        const std::string header = "#define POSITION_ATTRIBUTE_LOCATION 0\n#define COLOR_TEXTURE 0\n#define OBJECT_LUMINANCE_TEXTURE 1\n#define UNIFORM_MANUAL_EXPOSURE 0\n";
        toneProg = link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, header, read_file("tone_map_shader.vert")}, "tone_map_shader.vert"),
                                 compile_stage(GL_FRAGMENT_SHADER, {kVersion, header, read_file("tone_map_shader.frag")}, "tone_map_shader.frag")}, "tone map shader");
    }
    // BackgroundShader / BackgroundCubeShader / SSAOShader / SSAOApplyShader constructors: compatibility.glsl [+ common.glsl for the
    // vertex stage] + the stage source (background_shader.cpp:25-47, background_cube_shader.cpp, ssao_shader.cpp:41-55, ssao_apply_shader.cpp)
    const std::string compat = read_file("compatibility.glsl"), common = read_file("common.glsl");
    auto post_program = [&](const char* vert, const char* frag) {
        return link_program({compile_stage(GL_VERTEX_SHADER, {kVersion, compat, common, read_file(vert)}, vert),
                             compile_stage(GL_FRAGMENT_SHADER, {kVersion, compat, read_file(frag)}, frag)}, frag);
    };
    const GLuint bgProg = backgroundImage >= 0 ? post_program("background_shader.vert", "background_shader.frag") : 0;
    const GLuint bgCubeProg = hasLightMap ? post_program("background_cube_shader.vert", "background_cube_shader.frag") : 0;
    const GLuint ssaoProg = ssao ? post_program("ssao_shader.vert", "ssao_shader.frag") : 0;
    const GLuint ssaoApplyProg = ssao ? post_program("ssao_apply_shader.vert", "ssao_apply_shader.frag") : 0;
    check_gl("programs");

    // RenderPass::RenderPass (render_pass.cpp:271-292): 2048 x 2048 x NumLights depth array, linear filter, compare LEQUAL -----------------
    const int SR = 2048;
    GL::Texture2DArray shadowMaps;
    glGenTextures(1, &shadowMaps.id); glBindTexture(GL_TEXTURE_2D_ARRAY, shadowMaps.id);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE); glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_WRAP_R, GL_CLAMP_TO_EDGE);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_MAX_LEVEL, 0);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_COMPARE_FUNC, GL_LEQUAL);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_COMPARE_MODE, GL_COMPARE_REF_TO_TEXTURE);
    glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_MIN_FILTER, GL_LINEAR); glTexParameteri(GL_TEXTURE_2D_ARRAY, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    // GL::TextureFormat::DepthComponent = the unsized format; the reference's driver (NVIDIA) stores it as 24-bit fixed point. Mesa's
    // choice for the unsized format is also Z24; GLREF_SHADOW_FORMAT=24 asks for it by name.
    glTexImage3D(GL_TEXTURE_2D_ARRAY, 0, std::getenv("GLREF_SHADOW_SIZED") ? (GLint)GL_DEPTH_COMPONENT24 : (GLint)GL_DEPTH_COMPONENT, SR, SR, 3, 0, GL_DEPTH_COMPONENT, GL_FLOAT, nullptr);
    GLuint shadowFB[3];
    glGenFramebuffers(3, shadowFB);
    for (int i = 0; i < 3; ++i) {
        glBindFramebuffer(GL_FRAMEBUFFER, shadowFB[i]);
        glFramebufferTextureLayer(GL_FRAMEBUFFER, GL_DEPTH_ATTACHMENT, shadowMaps.id, 0, i);
        const GLenum none = GL_NONE; glDrawBuffers(1, &none);
        if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("shadow framebuffer incomplete");
    }
    check_gl("shadow maps");

    // RenderPass::render ----------------------------------------------------------------------------------------------------------
    glEnable(GL_DEPTH_TEST); glDisable(GL_CULL_FACE); glFrontFace(GL_CW); glDisable(GL_BLEND);   // render_pass.cpp:325-332
    // result textures and per-size buffers (render_pass.cpp:345-421)
    GLuint rgbTex = make_rect(GL_RGBA8, W, H, false), coordTex = make_rect(GL_RGBA32F, W, H, false), classTex = make_rect(GL_R16UI, W, H, true),
           instTex = make_rect(GL_R16UI, W, H, true), normalTex = make_rect(GL_RGBA32F, W, H, false), vidxTex = make_rect(GL_RGBA32UI, W, H, false),
           baryTex = make_rect(GL_RGBA32F, W, H, false), camTex = make_rect(GL_RGBA32F, W, H, false);
    const int levels = (int)Math::log2((UnsignedInt)std::max(W, H)) + 1;
    GLuint depthRB; glGenRenderbuffers(1, &depthRB); glBindRenderbuffer(GL_RENDERBUFFER, depthRB); glRenderbufferStorage(GL_RENDERBUFFER, GL_DEPTH_COMPONENT24, W, H);
    GL::Texture2D postprocessInput;
    glGenTextures(1, &postprocessInput.id); glBindTexture(GL_TEXTURE_2D, postprocessInput.id);
    glTexStorage2D(GL_TEXTURE_2D, levels, GL_RGBA32F, W, H);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAX_LEVEL, levels - 1);
    GL::Texture2D ssaoRGBInput, ssaoTexture;                  // render_pass.cpp:388-397
    if (ssao) {
        glGenTextures(1, &ssaoRGBInput.id); glBindTexture(GL_TEXTURE_2D, ssaoRGBInput.id);
        glTexStorage2D(GL_TEXTURE_2D, levels, GL_RGBA32F, W, H);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
        glGenTextures(1, &ssaoTexture.id); glBindTexture(GL_TEXTURE_2D, ssaoTexture.id);
        glTexStorage2D(GL_TEXTURE_2D, 1, GL_R32F, W, H);
        glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST); glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    }
    GL::RectangleTexture minDepth;
    glGenTextures(1, &minDepth.id); glBindTexture(GL_TEXTURE_RECTANGLE, minDepth.id);
    glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MAG_FILTER, GL_LINEAR);
    glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_WRAP_S, GL_CLAMP_TO_EDGE); glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_WRAP_T, GL_CLAMP_TO_EDGE);
    glTexStorage2D(GL_TEXTURE_RECTANGLE, 1, GL_RGBA32F, W, H);
    glTexSubImage2D(GL_TEXTURE_RECTANGLE, 0, 0, 0, W, H, GL_RGBA, GL_FLOAT, peel.data());
    check_gl("result textures");

    glFinish();
    const auto t_render0 = std::chrono::steady_clock::now();   // RenderPass::render proper starts here (assets and result buffers exist)
    // shadow pass (render_pass.cpp:423-460): front faces culled, every shadow caster's sub-meshes with shadowMatrix * meshToWorld
    glEnable(GL_CULL_FACE); glCullFace(GL_FRONT);
    for (int i = 0; i < 3; ++i) {
        if (lightCol[i] == Color3{0.0f} || lightDir[i] == Vector3{0.0f}) continue;
        if (!shadowActive[i]) fail("the dump marks a light inactive that render_pass.cpp:433 keeps");
        glBindFramebuffer(GL_FRAMEBUFFER, shadowFB[i]); glViewport(0, 0, SR, SR);
        glClear(GL_DEPTH_BUFFER_BIT);
        glUseProgram(shadowProg);
        for (const ObjIn& o : objs) {
            if (!o.visible || !o.casts) continue;
            const MeshIn& m = meshes[o.mesh];
            const Matrix4 t = shadowMatrices[i] * (o.pose * o.pre);
            glUniformMatrix4fv(0, 1, GL_FALSE, t.data());
            glBindVertexArray(m.vao);
            for (const SubIn& s : m.subs) glDrawElements(GL_TRIANGLES, s.count, GL_UNSIGNED_INT, (const void*)(size_t)(s.off * 4));
        }
    }
    glCullFace(GL_BACK); glDisable(GL_CULL_FACE);
    check_gl("shadow pass");

    // main framebuffer (render_pass.cpp:468-532)
    GLuint fb; glGenFramebuffers(1, &fb); glBindFramebuffer(GL_FRAMEBUFFER, fb); glViewport(0, 0, W, H);
    glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, ssao ? ssaoRGBInput.id : postprocessInput.id, 0);   // :469-473
    const GLuint rects[7] = {coordTex, classTex, instTex, normalTex, vidxTex, baryTex, camTex};
    for (int i = 0; i < 7; ++i) glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0 + 1 + i, GL_TEXTURE_RECTANGLE, rects[i], 0);
    glFramebufferRenderbuffer(GL_FRAMEBUFFER, GL_DEPTH_ATTACHMENT, GL_RENDERBUFFER, depthRB);
    GLenum bufs[8]; for (int i = 0; i < 8; ++i) bufs[i] = GL_COLOR_ATTACHMENT0 + i;   // output location i -> attachment i
    glDrawBuffers(8, bufs);
    if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("main framebuffer incomplete");
    glClear(GL_DEPTH_BUFFER_BIT);
    const GLfloat zero4[4] = {0, 0, 0, 0}, invalid[4] = {3000.0f, 3000.0f, 3000.0f, 3000.0f}, zero3[4] = {0, 0, 0, 1};   // 0x00000000_rgbf is a Color3: alpha 1
    const GLuint zeroui[4] = {0, 0, 0, 0};
    glClearBufferfv(GL_COLOR, 0, zero4); glClearBufferfv(GL_COLOR, 1, invalid); glClearBufferuiv(GL_COLOR, 2, zeroui); glClearBufferuiv(GL_COLOR, 3, zeroui);
    glClearBufferfv(GL_COLOR, 4, zero4); glClearBufferuiv(GL_COLOR, 5, zeroui); glClearBufferfv(GL_COLOR, 6, zero3); glClearBufferfv(GL_COLOR, 7, invalid);
    check_gl("clears");

    // lighting + per-frame uniforms (render_pass.cpp:534-543)
    if (hasLightMap) render.setLightMap(lightMap);
    else {
        Containers::Array<Vector3> dirs{3}; Containers::Array<Color3> cols{3};
        for (int i = 0; i < 3; ++i) { dirs[i] = lightDir[i]; cols[i] = lightCol[i]; }
        render.setManualLighting(dirs, cols, ambient);
    }
    render.bindDepthTexture(minDepth).setProjectionMatrix(P).setShadowMap(shadowMaps, shadowMatrices);
    render.use();

    // background plane (render_pass.cpp:545-582)
    if (planeSize.dot() > 0) {
        const Matrix4 scaledPoseInWorld = planePose * Matrix4::scaling({planeSize.x() / 2.0f, planeSize.y() / 2.0f, 1.0f});
        GL::Texture2D planeTexture; if (planeTex >= 0) planeTexture.id = tex[planeTex].id;
        Trade::MaterialData material = planeTex >= 0
            ? Trade::MaterialData{Trade::MaterialType::PbrMetallicRoughness, {{Trade::MaterialAttribute::BaseColor, Color4{1.0f}}, {Trade::MaterialAttribute::BaseColorTexture, 0u}}}
            : Trade::MaterialData{Trade::MaterialType::PbrMetallicRoughness, {{Trade::MaterialAttribute::BaseColor, Color4{0.0f, 0.8f, 0.0f, 1.0f}}}};
        auto textures = Containers::array<GL::Texture2D*>({planeTex >= 0 ? &planeTexture : nullptr});
        render.setClassIndex(0).setInstanceIndex(0).setStickerRange({}).setMaterial(material, textures, {})
            .setTransformations(Matrix4{Math::IdentityInit}, scaledPoseInWorld, V);
        render.use();
        glBindVertexArray(plane.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, plane.count);
    }
    // objects (render_pass.cpp:584-622)
    for (const ObjIn& o : objs) {
        if (!o.visible) continue;
        MeshIn& m = meshes[o.mesh];
        const Matrix4 objectToWorld = o.pose;
        const Matrix4 objectToCam = V * o.pose;
        const Matrix4 objectToCamInv = objectToCam.invertedRigid();
        const Matrix4 worldToCam = V;
        render.setClassIndex((unsigned)o.cls).setInstanceIndex((unsigned)o.inst).setStickerProjection(o.sticker_proj)
            .setStickerRange(Range2D::fromSize({o.sticker_range[0], o.sticker_range[1]}, {o.sticker_range[2], o.sticker_range[3]}));
        GL::RectangleTexture sticker;
        if (o.sticker >= 0) { sticker.id = tex[o.sticker].id; render.bindStickerTexture(sticker); }
        else { sticker.id = 0; render.bindStickerTexture(sticker); }   // a previous object's sticker must not leak into this harness' next draw
        auto materialOverride = sl::MaterialOverride{}.metallic(o.metallic).roughness(o.roughness);
        const Matrix4 meshToCam = V * (o.pose * o.pre);   // SceneGraph::Camera::draw: camera matrix * absolute transformation of the part
        const Matrix4 meshToObject = objectToCamInv * meshToCam;
        glBindVertexArray(m.vao);
        for (const SubIn& s : m.subs) {
            if (s.mat < 0 || s.mat >= (int)m.mats.size()) {   // Object::addPart: no material -> the context's default one (src/object.cpp:119-125, context.cpp:382-384)
                using namespace Math::Literals;
                Trade::MaterialData material{Trade::MaterialType::PbrMetallicRoughness, {{Trade::MaterialAttribute::BaseColor, 0x3bd267ff_srgbaf}}};
                Containers::Array<GL::Texture2D*> none;
                render.setMaterial(material, none, materialOverride).setTransformations(meshToObject, objectToWorld, worldToCam);
                render.use();
                glDrawElements(GL_TRIANGLES, s.count, GL_UNSIGNED_INT, (const void*)(size_t)(s.off * 4));
                continue;
            }
            const MatIn& mi = m.mats[s.mat];
            // the material as the importer hands it over: explicit factors + texture references (indices into this mesh's texture list)
            Containers::Array<Trade::MaterialAttributeData> attrs;
            arrayAppend(attrs, Corrade::InPlaceInit, Trade::MaterialAttribute::BaseColor, Color4{mi.base[0], mi.base[1], mi.base[2], mi.base[3]});
            arrayAppend(attrs, Corrade::InPlaceInit, Trade::MaterialAttribute::EmissiveColor, Color3{mi.emissive[0], mi.emissive[1], mi.emissive[2]});
            arrayAppend(attrs, Corrade::InPlaceInit, Trade::MaterialAttribute::Metalness, mi.metallic);
            arrayAppend(attrs, Corrade::InPlaceInit, Trade::MaterialAttribute::Roughness, mi.roughness);
            std::vector<GL::Texture2D> store(5);
            Containers::Array<GL::Texture2D*> textures{Corrade::ValueInit, 5};
            const Trade::MaterialAttribute names[5] = {Trade::MaterialAttribute::BaseColorTexture, Trade::MaterialAttribute::NormalTexture,
                                                       Trade::MaterialAttribute::NoneRoughnessMetallicTexture, Trade::MaterialAttribute::EmissiveTexture,
                                                       Trade::MaterialAttribute::OcclusionTexture};
            for (UnsignedInt k = 0; k < 5; ++k)
                if (mi.tex[k] >= 0) { store[k].id = tex[mi.tex[k]].id; textures[k] = &store[k]; arrayAppend(attrs, Corrade::InPlaceInit, names[k], k); }
            Trade::MaterialData material{Trade::MaterialType::PbrMetallicRoughness, std::move(attrs)};
            render.setMaterial(material, textures, materialOverride).setTransformations(meshToObject, objectToWorld, worldToCam);
            render.use();
            glDrawElements(GL_TRIANGLES, s.count, GL_UNSIGNED_INT, (const void*)(size_t)(s.off * 4));
        }
    }
    check_gl("main pass");
    glFrontFace(GL_CCW);                                                        // render_pass.cpp:630
    glBindTexture(GL_TEXTURE_2D, ssao ? ssaoRGBInput.id : postprocessInput.id); glGenerateMipmap(GL_TEXTURE_2D);   // :632-635

    // background image or sky box into colour attachment 0 only (render_pass.cpp:637-660)
    if (backgroundImage >= 0) {
        const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b);
        glUseProgram(bgProg);
        GL::RectangleTexture bg; bg.id = tex[backgroundImage].id; bg.bind(0);
        glBindVertexArray(quad.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, quad.count);
    } else if (hasLightMap) {
        const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b);
        glDepthFunc(GL_LEQUAL);
        glUseProgram(bgCubeProg);
        lightMap.cubeMap().bind(0);
        glUniformMatrix4fv(0, 1, GL_FALSE, V.data());       // setViewMatrix(cameraMatrix): location 0
        glUniformMatrix4fv(1, 1, GL_FALSE, P.data());       // setProjectionMatrix: location 1
        glBindVertexArray(skyCube.vao); glDrawElements(GL_TRIANGLES, skyCube.count, GL_UNSIGNED_INT, nullptr);
        glDepthFunc(GL_LESS);
    }
    check_gl("background");

    if (ssao) {   // render_pass.cpp:662-694
        ssaogen::Tables tables;                              // SSAOShader's constructor data: noise texture + kernel
        GLuint ssaoFB; glGenFramebuffers(1, &ssaoFB); glBindFramebuffer(GL_FRAMEBUFFER, ssaoFB); glViewport(0, 0, W, H);
        glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, ssaoTexture.id, 0);
        { const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b); }
        if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("SSAO framebuffer incomplete");
        const GLfloat one4[4] = {1, 1, 1, 1}; glClearBufferfv(GL_COLOR, 0, one4);
        glUseProgram(ssaoProg);
        glUniformMatrix4fv(65, 1, GL_FALSE, P.data());                                  // setProjection: m_projectionUniform{65} (ssao_shader.h:69)
        GL::RectangleTexture camR, nrmR; camR.id = camTex; nrmR.id = normalTex;
        camR.bind(0); nrmR.bind(1);                                                      // CoordinateLayer 0, NormalLayer 1
        tables.m_noiseTexture.bind(2);                                                   // bindNoise: NoiseLayer 2 + the kernel at m_samplesUniform{0}
        glUniform3fv(0, 64, tables.m_ssaoKernel[0].data());
        glBindVertexArray(quad.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, quad.count);

        GLuint applyFB; glGenFramebuffers(1, &applyFB); glBindFramebuffer(GL_FRAMEBUFFER, applyFB); glViewport(0, 0, W, H);
        glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_2D, postprocessInput.id, 0);
        { const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b); }
        if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("SSAO apply framebuffer incomplete");
        glClearBufferfv(GL_COLOR, 0, zero4);
        glUseProgram(ssaoApplyProg);
        ssaoTexture.bind(1); ssaoRGBInput.bind(0); camR.bind(2);                         // bindAO 1, bindColor 0, bindCoordinates 2 (ssao_apply_shader.cpp:24-29)
        glBindVertexArray(quad.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, quad.count);
        check_gl("SSAO");
    }

    // tone map (render_pass.cpp:696-710)
    GLuint postFB; glGenFramebuffers(1, &postFB); glBindFramebuffer(GL_FRAMEBUFFER, postFB); glViewport(0, 0, W, H);
    glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0, GL_TEXTURE_RECTANGLE, rgbTex, 0);
    { const GLenum b = GL_COLOR_ATTACHMENT0; glDrawBuffers(1, &b); }
    if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) fail("post-process framebuffer incomplete");
    glUseProgram(toneProg);
    postprocessInput.bind(0);                                                            // bindColor(m_postprocessInput)
    if (ssao) ssaoRGBInput.bind(1); else postprocessInput.bind(1);                       // bindObjectLuminance
    glUniform1f(0, manualExposure);
    glBindVertexArray(quad.vao); glDrawArrays(GL_TRIANGLE_STRIP, 0, quad.count);
    glFinish();
    check_gl("tone map");
    const double render_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_render0).count();
    if (std::getenv("GLREF_VERBOSE") || std::getenv("GLREF_TIMING")) std::fprintf(stderr, "glref: render_ms %.3f\n", render_ms);

    // the HDR buffer the tone map read
    std::vector<float> hdr((size_t)W * H * 4);
    glPixelStorei(GL_PACK_ALIGNMENT, 1);
    glBindTexture(GL_TEXTURE_2D, postprocessInput.id); glGetTexImage(GL_TEXTURE_2D, 0, GL_RGBA, GL_FLOAT, hdr.data());

    // read back (row 0 = window y 0, SURVEY 8 preamble: no flip)
    std::vector<unsigned char> rgb((size_t)W * H * 4);
    std::vector<float> coord((size_t)W * H * 4), normals(coord.size()), bary(coord.size()), cam(coord.size());
    std::vector<unsigned short> cls((size_t)W * H), inst((size_t)W * H);
    std::vector<UnsignedInt> vidx((size_t)W * H * 4);
    auto get = [&](GLuint t, GLenum f, GLenum ty, void* dst) { glBindTexture(GL_TEXTURE_RECTANGLE, t); glGetTexImage(GL_TEXTURE_RECTANGLE, 0, f, ty, dst); };
    get(rgbTex, GL_RGBA, GL_UNSIGNED_BYTE, rgb.data()); get(coordTex, GL_RGBA, GL_FLOAT, coord.data());
    get(classTex, GL_RED_INTEGER, GL_UNSIGNED_SHORT, cls.data()); get(instTex, GL_RED_INTEGER, GL_UNSIGNED_SHORT, inst.data());
    get(normalTex, GL_RGBA, GL_FLOAT, normals.data()); get(vidxTex, GL_RGBA_INTEGER, GL_UNSIGNED_INT, vidx.data());
    get(baryTex, GL_RGBA, GL_FLOAT, bary.data()); get(camTex, GL_RGBA, GL_FLOAT, cam.data());
    check_gl("read back");
    std::ofstream out(out_path, std::ios::binary);
    auto put = [&](const void* p, size_t n) { out.write((const char*)p, (std::streamsize)n); };
    put(rgb.data(), rgb.size()); put(coord.data(), coord.size() * 4); put(cls.data(), cls.size() * 2); put(inst.data(), inst.size() * 2);
    put(normals.data(), normals.size() * 4); put(vidx.data(), vidx.size() * 4); put(bary.data(), bary.size() * 4); put(cam.data(), cam.size() * 4);
    put(hdr.data(), hdr.size() * 4);
    if (hasLightMap) {   // the precomputed maps, in the layout of slb_lightmap_read: env level 0, irradiance, prefilter levels, LUT
        auto put_cube = [&](GLuint id, int level, int size) {
            std::vector<float> face((size_t)size * size * 4);
            glBindTexture(GL_TEXTURE_CUBE_MAP, id);
            for (int f = 0; f < 6; ++f) { glGetTexImage(GL_TEXTURE_CUBE_MAP_POSITIVE_X + f, level, GL_RGBA, GL_FLOAT, face.data()); put(face.data(), face.size() * 4); }
        };
        put_cube(lightMap.cube.id, 0, lmIn.sizes[0]);
        put_cube(lightMap.irradiance.id, 0, lmIn.sizes[1]);
        for (int mip = 0; mip < 5; ++mip) put_cube(lightMap.prefilter.id, mip, lmIn.sizes[2] >> mip);
        std::vector<float> lut((size_t)lmIn.sizes[3] * lmIn.sizes[3] * 4);
        glBindTexture(GL_TEXTURE_2D, lightMap.lut.id); glGetTexImage(GL_TEXTURE_2D, 0, GL_RGBA, GL_FLOAT, lut.data());
        put(lut.data(), lut.size() * 4);
    }
    if (std::getenv("GLREF_DUMP_SHADOW")) {   // the three depth layers of the shadow-map array as float (tests compare them with the oracle's d24 maps)
        std::vector<float> sm((size_t)SR * SR * 3);
        glBindTexture(GL_TEXTURE_2D_ARRAY, shadowMaps.id); glGetTexImage(GL_TEXTURE_2D_ARRAY, 0, GL_DEPTH_COMPONENT, GL_FLOAT, sm.data());
        put(sm.data(), sm.size() * 4);
    }
    if (const char* mp = std::getenv("GLREF_DUMP_MIPS")) {   // every 2-D texture's generated chain: i32 n_levels, then RGBA8 levels
        std::ofstream mo(mp, std::ios::binary);
        for (const TexIn& t : tex) {
            const int n = (int)t.levels.size(); mo.write((const char*)&n, 4);
            for (const auto& l : t.levels) mo.write((const char*)l.data(), (std::streamsize)l.size());
        }
    }
    if (!out) fail("cannot write the output");
    return 0;
}

// usage: glref <scene dump> <output file> [<scene dump> <output file> ...]   — one context, the scenes rendered one after the other
int main(int argc, char** argv) {
    if (argc < 3 || (argc - 1) % 2) fail("usage: glref <scene dump> <output file> [<scene dump> <output file> ...]");
    if (const char* sd = std::getenv("GLREF_SHADER_DIR")) g_shader_dir = sd;
    for (int i = 1; i + 1 < argc; i += 2) render_scene(argv[i], argv[i + 1], i == 1);
    return 0;
}
