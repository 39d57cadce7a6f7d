// Builtin functions of the scene language: same names, signatures, defaults and overload order as
// libSLRSceneGraph/API.cpp:99-1115 and Parser/BuiltinFunctions/*.cpp, implemented over the host
// classes of this repo. Model import (load3DModel) follows libSLRSceneGraph/node_constructor.cpp
// with the in-repo .assbin reader instead of assimp.
#include <cstdio>
#include "interp.h"
#include "../assets/assbin.h"
#include "../assets/images.h"
#include "../raycast.h"
#include <cmath>
#include <iostream>

namespace slr {
namespace lang {

namespace {

const Type R = Type::RealNumber;

Value realArg(const Args& a, const char* n) { return a.at(n); }
float f(const Args& a, const char* n) { return (float)a.at(n).d; }

FunctionRef fn(const std::vector<ArgInfo>& sig, const Function::Native& body) { return std::make_shared<Function>(sig, body); }
FunctionRef fnOver(const std::vector<std::vector<ArgInfo>>& sigs, const std::vector<Function::Native>& bodies) {
    return std::make_shared<Function>(sigs, bodies);
}
// XORShiftRNG (libSLR/RNGs/XORShiftRNG.cpp:21-37, RandomNumberGenerator.cpp:12-15). The seed is a SIGNED 32-bit integer in the
// reference's seeding loop, so `seed >> 30` is an arithmetic shift once an intermediate state has its top bit set.
struct XorShift128 {
    uint32_t s[4];
    explicit XorShift128(int32_t seed) {
        for (uint32_t i = 0; i < 4; ++i) {
            s[i] = 1812433253U * (uint32_t)(seed ^ (seed >> 30)) + i;
            seed = (int32_t)s[i];
        }
        for (int i = 0; i < 50; ++i) next();
    }
    uint32_t next() {
        const uint32_t t = s[0] ^ (s[0] << 11);
        s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
        return s[3] = (s[3] ^ (s[3] >> 19)) ^ (t ^ (t >> 8));
    }
    float float0cTo1o() {
        const uint32_t bits = (next() >> 9) | 0x3f800000u;
        float v;
        std::memcpy(&v, &bits, 4);
        return v - 1.0f;
    }
};

void def(Interpreter& in, const char* name, const FunctionRef& f) { in.defineGlobal(name, Value::Ref(Type::Function, f)); }

// Runs a nested "configuration" signature over a tuple (the reference's configFunc pattern).
Value withConfig(const ParameterList& params, const std::vector<ArgInfo>& sig, Interpreter& in, const Function::Native& body) {
    Function cfg(sig, body);
    return cfg.call(params, in);
}

bool strToSpectrumType(const std::string& s, SpectrumType* t) {
    if (s == "Reflectance") *t = SpectrumType::Reflectance;
    else if (s == "Illuminant") *t = SpectrumType::Illuminant;
    else if (s == "IndexOfRefraction") *t = SpectrumType::IndexOfRefraction;
    else return false;
    return true;
}
bool strToColorSpace(const std::string& s, ColorSpace* c) {
    if (s == "Rec709") *c = ColorSpace::sRGB;
    else if (s == "sRGB") *c = ColorSpace::sRGB_NonLinear;
    else if (s == "xyY") *c = ColorSpace::xyY;
    else if (s == "XYZ") *c = ColorSpace::XYZ;
    else return false;
    return true;
}

TextureMappingRef sharedTexCoordMapping() { static TextureMappingRef m = std::make_shared<TextureMapping>(); return m; }
TextureMappingRef sharedWorldPosMapping() {
    static TextureMappingRef m = [] { auto x = std::make_shared<TextureMapping>(); x->kind = SLRGPU_MAP_WORLD_POS; return x; }();
    return m;
}

Vec3 tupleToVec3(const Value& t, Interpreter& in, const char* what) {
    Args a;
    if (t.type != Type::Tuple || !mapParamsToArgs(t.tuple(), {{"x", R}, {"y", R}, {"z", R}}, &a)) in.fail(std::string("invalid ") + what);
    return Vec3((float)a["x"].d, (float)a["y"].d, (float)a["z"].d);
}

Value makeVertex(const Args& args, Interpreter& in) {
    Vertex v;
    v.position = tupleToVec3(args.at("position"), in, "vertex position");
    v.normal = tupleToVec3(args.at("normal"), in, "vertex normal");
    v.tangent = tupleToVec3(args.at("tangent"), in, "vertex tangent");
    Args tc;
    if (!mapParamsToArgs(args.at("texCoord").tuple(), {{"u", R}, {"v", R}}, &tc)) in.fail("invalid vertex texCoord");
    v.texCoord = Vec2((float)tc["u"].d, (float)tc["v"].d);
    return Value::Vtx(v);
}

const std::vector<ArgInfo> kVertexSig = {{"position", Type::Tuple}, {"normal", Type::Tuple}, {"tangent", Type::Tuple}, {"texCoord", Type::Tuple}};

// ---- model import ---------------------------------------------------------------------------
struct SurfaceAttributes { SurfaceMaterialRef material; Normal3DTextureRef normalMap; FloatTextureRef alphaMap; };

void makeTangent(float nx, float ny, float nz, float* s) {
    if (std::fabs(nx) > std::fabs(ny)) {
        float invLen = 1.0f / std::sqrt(nx * nx + nz * nz);
        s[0] = -nz * invLen; s[1] = 0.0f; s[2] = nx * invLen;
    } else {
        float invLen = 1.0f / std::sqrt(ny * ny + nz * nz);
        s[0] = 0.0f; s[1] = nz * invLen; s[2] = -ny * invLen;
    }
}

SurfaceAttributes defaultMaterial(const assbin::Material& m, const std::string& pathPrefix, Interpreter& in) {
    SurfaceAttributes out;
    std::string file;
    float color[3];
    SpectrumTextureRef diffuse;
    if (m.getString("$tex.file", assbin::kTexDiffuse, 0, &file)) {
        Image2DRef img = loadImageCached(pathPrefix + file, ImageStoreMode::AsIs, SpectrumType::Reflectance, in.rgbMode);
        diffuse = SpectrumTexture::imageTexture(sharedTexCoordMapping(), img);
    } else if (m.getColor("$clr.diffuse", color)) {
        diffuse = SpectrumTexture::constant(Spectrum::create(in.rgbMode, SpectrumType::Reflectance, ColorSpace::sRGB_NonLinear, color[0], color[1], color[2]));
    } else {
        diffuse = SpectrumTexture::constant(Spectrum::create(in.rgbMode, SpectrumType::Reflectance, ColorSpace::sRGB_NonLinear, 1.0f, 0.0f, 1.0f));
    }
    out.material = SurfaceMaterial::createMatte(diffuse, nullptr);
    if (m.getString("$tex.file", assbin::kTexDisplacement, 0, &file))
        out.normalMap = Normal3DTexture::imageTexture(sharedTexCoordMapping(),
                                                      loadImageCached(pathPrefix + file, ImageStoreMode::NormalTexture, SpectrumType::Reflectance, in.rgbMode));
    if (m.getString("$tex.file", assbin::kTexOpacity, 0, &file))
        out.alphaMap = FloatTexture::imageTexture(sharedTexCoordMapping(),
                                                  loadImageCached(pathPrefix + file, ImageStoreMode::AlphaTexture, SpectrumType::Reflectance, in.rgbMode));
    return out;
}

SurfaceAttributes userMaterial(const assbin::Material& m, const std::string& pathPrefix, const FunctionRef& proc, Interpreter& in) {
    std::string name;
    m.getString("?mat.name", 0, 0, &name);
    auto attrs = std::make_shared<ParameterList>();
    auto texList = [&](uint32_t semantic) {
        auto l = std::make_shared<ParameterList>();
        for (uint32_t i = 0; i < m.textureCount(semantic); ++i) {
            std::string file;
            if (m.getString("$tex.file", semantic, i, &file)) l->add("", Value::Str(pathPrefix + file));
        }
        return Value::Tuple(l);
    };
    attrs->add("diffuse textures", texList(assbin::kTexDiffuse));
    attrs->add("specular textures", texList(assbin::kTexSpecular));
    attrs->add("emissive textures", texList(assbin::kTexEmissive));
    attrs->add("height textures", texList(assbin::kTexHeight));
    attrs->add("normal textures", texList(assbin::kTexNormals));
    auto rgb = [](const float* c) {
        auto l = std::make_shared<ParameterList>();
        for (int i = 0; i < 3; ++i) l->add("", Value::Real(c[i]));      // Element(float) widens to RealNumber
        return Value::Tuple(l);
    };
    float color[3];
    if (m.getColor("$clr.diffuse", color)) attrs->add("diffuse color", rgb(color));
    if (m.getColor("$clr.specular", color)) attrs->add("specular color", rgb(color));
    if (m.getColor("$clr.emissive", color)) attrs->add("emissive color", rgb(color));
    ParameterList params;
    params.add("", Value::Str(name));
    params.add("", Value::Tuple(attrs));
    Value result = proc->call(params, in);
    if (result.type == Type::Tuple) {
        const ParameterList& t = result.tuple();
        if (!t.unnamed.empty() && t.unnamed[0].type == Type::SurfaceMaterial) {
            SurfaceAttributes out;
            out.material = t.unnamed[0].as<SurfaceMaterial>();
            if (t.unnamed.size() > 1 && t.unnamed[1].type == Type::NormalTexture) out.normalMap = t.unnamed[1].as<Normal3DTexture>();
            if (t.unnamed.size() > 2 && t.unnamed[2].type == Type::FloatTexture) out.alphaMap = t.unnamed[2].as<FloatTexture>();
            return out;
        }
    } else if (result.type == Type::SurfaceMaterial) {
        return SurfaceAttributes{result.as<SurfaceMaterial>(), nullptr, nullptr};
    }
    std::printf("User defined material function is invalid, fall back to the default function.\n");
    return defaultMaterial(m, pathPrefix, in);
}

InternalNodeRef constructNode(const assbin::Scene& sc, const assbin::Node& src, const std::vector<SurfaceAttributes>& mats) {
    if (src.meshes.empty() && src.children.empty()) return nullptr;
    auto node = std::make_shared<InternalNode>();
    node->name = src.name;
    // The reference feeds assimp's row-major element list to Matrix4x4's column-major array
    // constructor (node_constructor.cpp:45-52), i.e. it stores the transpose. Kept as is.
    Mat4 m;
    for (int c = 0; c < 4; ++c) m.c[c] = Vec4(src.transform[4 * c], src.transform[4 * c + 1], src.transform[4 * c + 2], src.transform[4 * c + 3]);
    node->setTransform(StaticTransform(m));
    for (uint32_t mi : src.meshes) {
        if (mi >= sc.meshes.size()) continue;
        const assbin::Mesh& mesh = sc.meshes[mi];
        if (mesh.primitiveTypes != assbin::kPrimitiveTriangle) { std::printf("ignored non triangle mesh.\n"); continue; }
        auto tm = std::make_shared<TriangleMeshNode>();
        const SurfaceAttributes& sa = mats[std::min<size_t>(mesh.materialIndex, mats.size() - 1)];
        const bool hasT = !mesh.tangents.empty(), hasUV = mesh.numUVComponents > 0 && !mesh.texCoords.empty();
        for (uint32_t v = 0; v < mesh.numVertices(); ++v) {
            Vertex o;
            o.position = Vec3(mesh.positions[3 * v], mesh.positions[3 * v + 1], mesh.positions[3 * v + 2]);
            o.normal = mesh.normals.empty() ? Vec3(0, 1, 0) : Vec3(mesh.normals[3 * v], mesh.normals[3 * v + 1], mesh.normals[3 * v + 2]);
            float t[3];
            if (hasT) { t[0] = mesh.tangents[3 * v]; t[1] = mesh.tangents[3 * v + 1]; t[2] = mesh.tangents[3 * v + 2]; }
            else makeTangent(o.normal.x, o.normal.y, o.normal.z, t);
            o.tangent = Vec3(t[0], t[1], t[2]);
            o.texCoord = hasUV ? Vec2(mesh.texCoords[3 * v], mesh.texCoords[3 * v + 1]) : Vec2(0, 0);
            float dotNT = dot(o.normal, o.tangent);
            if (std::fabs(dotNT) >= 0.01f) o.tangent = normalize(o.tangent - dotNT * o.normal);
            tm->addVertex(o);
        }
        std::vector<uint32_t> idx = mesh.indices;
        tm->addTriangles(sa.material, sa.normalMap, sa.alphaMap, std::move(idx));
        tm->name = mesh.name;
        node->addChildNode(tm);
    }
    for (const assbin::Node& c : src.children) {
        InternalNodeRef sub = constructNode(sc, c, mats);
        if (sub) node->addChildNode(sub);
    }
    return node;
}

}  // namespace

void registerBuiltins(Interpreter& in) {
    in.defineGlobal("root", Value::Ref(Type::Node, in.scene->rootNode()));

    def(in, "print", fn({{"value", Type::Any}}, [](const Args& a, Interpreter&) { std::cout << a.at("value").toString() << std::endl; return Value(); }));
    def(in, "addItem", fn({{"tuple", Type::Tuple}, {"key", Type::String, Value::Str("")}, {"item", Type::Any}}, [](const Args& a, Interpreter&) {
        a.at("tuple").as<ParameterList>()->add(a.at("key").s, a.at("item"));
        return a.at("tuple");
    }));
    def(in, "numElements", fn({{"tuple", Type::Tuple}}, [](const Args& a, Interpreter&) { return Value::Int((int32_t)a.at("tuple").tuple().numParams()); }));
    def(in, "Point", fn({{"x", R}, {"y", R}, {"z", R}}, [](const Args& a, Interpreter&) { return Value::Vec(Type::Point, Vec3(f(a, "x"), f(a, "y"), f(a, "z"))); }));
    def(in, "Vector", fn({{"x", R}, {"y", R}, {"z", R}}, [](const Args& a, Interpreter&) { return Value::Vec(Type::Vector, Vec3(f(a, "x"), f(a, "y"), f(a, "z"))); }));
    auto getter = [](int axis) {
        std::vector<Function::Native> fs;
        for (const char* n : {"point", "vector", "normal"})
            fs.push_back([axis, n](const Args& a, Interpreter&) { return Value::Real(a.at(n).v3[axis]); });
        return fnOver({{{"point", Type::Point}}, {{"vector", Type::Vector}}, {{"normal", Type::Normal}}}, fs);
    };
    def(in, "getX", getter(0)); def(in, "getY", getter(1)); def(in, "getZ", getter(2));
    {
        // the reference's generator is a static of its process (API.cpp:238-244) and a process reads one scene file: here
        // every interpreter (= every scene file read) starts the stream afresh, whatever the process read before
        auto rng = std::make_shared<XorShift128>(2112984105);
        def(in, "random", fn({}, [rng](const Args&, Interpreter&) { return Value::Real(rng->float0cTo1o()); }));
    }

    // ---- math (BuiltinFunctions/builtin_math.cpp)
    // the reference takes every argument as a `float` and calls the float overloads (builtin_math.cpp:15-84): sin(0.3) is
    // sinf(0.3f), not the double value -- the results feed transforms, so they are kept to the bit
    def(in, "min", fn({{"x0", R}, {"x1", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::min(f(a, "x0"), f(a, "x1"))); }));
    def(in, "max", fn({{"x0", R}, {"x1", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::max(f(a, "x0"), f(a, "x1"))); }));
    def(in, "clamp", fn({{"x", R}, {"min", R}, {"max", R}}, [](const Args& a, Interpreter&) {
        const float x = f(a, "x"), lo = f(a, "min"), hi = f(a, "max");
        return Value::Real(x < lo ? lo : (hi < x ? hi : x));           // std::clamp(x, min, max)
    }));
    def(in, "sqrt", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::sqrt(f(a, "x"))); }));
    def(in, "pow", fn({{"x", R}, {"e", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::pow(f(a, "x"), f(a, "e"))); }));
    def(in, "sin", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::sin(f(a, "x"))); }));
    def(in, "cos", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::cos(f(a, "x"))); }));
    def(in, "tan", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::tan(f(a, "x"))); }));
    def(in, "asin", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::asin(f(a, "x"))); }));
    def(in, "acos", fn({{"x", R}}, [](const Args& a, Interpreter&) { return Value::Real(std::acos(f(a, "x"))); }));
    def(in, "atan", fnOver({{{"x", R}}, {{"y", R}, {"x", R}}},
                           {[](const Args& a, Interpreter&) { return Value::Real(std::atan(f(a, "x"))); },
                            [](const Args& a, Interpreter&) { return Value::Real(std::atan2(f(a, "y"), f(a, "x"))); }}));
    def(in, "dot", fn({{"v0", Type::Vector}, {"v1", Type::Vector}}, [](const Args& a, Interpreter&) { return Value::Real(dot(a.at("v0").v3, a.at("v1").v3)); }));
    def(in, "cross", fn({{"v0", Type::Vector}, {"v1", Type::Vector}}, [](const Args& a, Interpreter&) { return Value::Vec(Type::Vector, cross(a.at("v0").v3, a.at("v1").v3)); }));
    def(in, "distance", fn({{"p0", Type::Point}, {"p1", Type::Point}}, [](const Args& a, Interpreter&) { return Value::Real((a.at("p1").v3 - a.at("p0").v3).length()); }));

    // ---- transforms (BuiltinFunctions/builtin_transform.cpp)
    def(in, "translate", fn({{"x", R}, {"y", R}, {"z", R}}, [](const Args& a, Interpreter&) { return Value::Matrix(translate(f(a, "x"), f(a, "y"), f(a, "z"))); }));
    def(in, "rotate", fn({{"angle", R}, {"axis", Type::Vector}}, [](const Args& a, Interpreter&) { return Value::Matrix(rotate(f(a, "angle"), a.at("axis").v3)); }));
    def(in, "rotateX", fn({{"angle", R}}, [](const Args& a, Interpreter&) { return Value::Matrix(rotate(f(a, "angle"), Vec3(1, 0, 0))); }));
    def(in, "rotateY", fn({{"angle", R}}, [](const Args& a, Interpreter&) { return Value::Matrix(rotate(f(a, "angle"), Vec3(0, 1, 0))); }));
    def(in, "rotateZ", fn({{"angle", R}}, [](const Args& a, Interpreter&) { return Value::Matrix(rotate(f(a, "angle"), Vec3(0, 0, 1))); }));
    def(in, "scale", fnOver({{{"s", R}}, {{"x", R}, {"y", R}, {"z", R}}},
                            {[](const Args& a, Interpreter&) { float s = f(a, "s"); return Value::Matrix(scale(s, s, s)); },
                             [](const Args& a, Interpreter&) { return Value::Matrix(scale(f(a, "x"), f(a, "y"), f(a, "z"))); }}));
    def(in, "lookAt", fn({{"eye", Type::Tuple}, {"target", Type::Tuple}, {"up", Type::Tuple}}, [](const Args& a, Interpreter& in) {
        Mat4 raw = lookAt(tupleToVec3(a.at("eye"), in, "eye"), tupleToVec3(a.at("target"), in, "target"), tupleToVec3(a.at("up"), in, "up"));
        return Value::Matrix(invert(raw) * rotate((float)M_PI, Vec3(0, 1, 0)));
    }));
    def(in, "AnimatedTransform", fn({{"tfStart", Type::Matrix}, {"tfEnd", Type::Matrix}, {"tBegin", R}, {"tEnd", R}}, [](const Args& a, Interpreter& in) -> Value {
        // builtin_transform.cpp:81-90: an AnimatedTransform between two key matrices; a Transform value carries it in
        // StaticTransform::anim (geom.h), with the begin key frame as its static part
        if (a.at("tfStart").m == a.at("tfEnd").m) return a.at("tfStart").convertTo(Type::Transform);
        auto tf = std::make_shared<StaticTransform>(a.at("tfStart").m);
        tf->anim = std::make_shared<AnimatedTransform>(*tf, StaticTransform(a.at("tfEnd").m), f(a, "tBegin"), f(a, "tEnd"));
        return Value::Ref(Type::Transform, tf);
    }));

    // ---- textures (BuiltinFunctions/builtin_texture.cpp)
    def(in, "Texture2DMapping", fn({{"type", Type::String, Value::Str("texcoord 2D")}, {"params", Type::Tuple, Value::Tuple(std::make_shared<ParameterList>())}},
                                   [](const Args& a, Interpreter& in) -> Value {
                                       if (a.at("type").s == "texcoord 2D") return Value::Ref(Type::Texture2DMapping, sharedTexCoordMapping());
                                       in.fail("Specified type is invalid.");
                                   }));
    def(in, "Texture3DMapping", fn({{"type", Type::String, Value::Str("texcoord 2D")}, {"params", Type::Tuple, Value::Tuple(std::make_shared<ParameterList>())}},
                                   [](const Args& a, Interpreter& in) -> Value {
                                       if (a.at("type").s == "texcoord 2D") return Value::Ref(Type::Texture3DMapping, sharedTexCoordMapping());
                                       if (a.at("type").s == "world pos") return Value::Ref(Type::Texture3DMapping, sharedWorldPosMapping());
                                       in.fail("Specified type is invalid.");
                                   }));
    const Value def2D = Value::Ref(Type::Texture2DMapping, sharedTexCoordMapping());
    const Value defWorld = Value::Ref(Type::Texture3DMapping, sharedWorldPosMapping());
    def(in, "SpectrumTexture", fnOver(
        {{{"spectrum", Type::Spectrum}}, {{"image", Type::Image2D}, {"mapping", Type::Texture2DMapping, def2D}}, {{"procedure", Type::String}, {"params", Type::Tuple}}},
        {[](const Args& a, Interpreter&) { return Value::Ref(Type::SpectrumTexture, SpectrumTexture::constant(a.at("spectrum").as<InputSpectrum>())); },
         [](const Args& a, Interpreter&) { return Value::Ref(Type::SpectrumTexture, SpectrumTexture::imageTexture(a.at("mapping").as<TextureMapping>(), a.at("image").as<Image2D>())); },
         [def2D, defWorld](const Args& a, Interpreter& in) -> Value {
             const std::string& proc = a.at("procedure").s;
             if (proc == "checker board")
                 return withConfig(a.at("params").tuple(), {{"c0", Type::Spectrum}, {"c1", Type::Spectrum}, {"mapping", Type::Texture2DMapping, def2D}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::SpectrumTexture, SpectrumTexture::checkerBoard(c.at("mapping").as<TextureMapping>(), c.at("c0").as<InputSpectrum>(), c.at("c1").as<InputSpectrum>()));
                                   });
             if (proc == "voronoi")
                 return withConfig(a.at("params").tuple(), {{"scale", R}, {"brightness", R, Value::Real(0.8f)}, {"mapping", Type::Texture3DMapping, defWorld}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::SpectrumTexture, SpectrumTexture::voronoi(c.at("mapping").as<TextureMapping>(), f(c, "scale"), f(c, "brightness")));
                                   });
             in.fail("Specified procedure is invalid.");
         }}));
    def(in, "NormalTexture", fnOver(
        {{{"image", Type::Image2D}, {"mapping", Type::Texture2DMapping, def2D}}, {{"procedure", Type::String}, {"params", Type::Tuple}}},
        {[](const Args& a, Interpreter&) { return Value::Ref(Type::NormalTexture, Normal3DTexture::imageTexture(a.at("mapping").as<TextureMapping>(), a.at("image").as<Image2D>())); },
         [def2D, defWorld](const Args& a, Interpreter& in) -> Value {
             const std::string& proc = a.at("procedure").s;
             if (proc == "checker board")
                 return withConfig(a.at("params").tuple(), {{"stepWidth", R, Value::Real(0.05)}, {"reverse", Type::Bool, Value::Bool(false)}, {"mapping", Type::Texture2DMapping, def2D}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::NormalTexture, Normal3DTexture::checkerBoard(c.at("mapping").as<TextureMapping>(), f(c, "stepWidth"), c.at("reverse").b));
                                   });
             if (proc == "voronoi")
                 return withConfig(a.at("params").tuple(), {{"scale", R}, {"thetaMax", R, Value::Real(M_PI / 6)}, {"mapping", Type::Texture3DMapping, defWorld}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::NormalTexture, Normal3DTexture::voronoi(c.at("mapping").as<TextureMapping>(), f(c, "scale"), f(c, "thetaMax")));
                                   });
             in.fail("Specified procedure is invalid.");
         }}));
    def(in, "FloatTexture", fnOver(
        {{{"value", R}}, {{"image", Type::Image2D}, {"mapping", Type::Texture2DMapping, def2D}}, {{"procedure", Type::String}, {"params", Type::Tuple}}},
        {[](const Args& a, Interpreter&) { return Value::Ref(Type::FloatTexture, FloatTexture::constant(f(a, "value"))); },
         [](const Args& a, Interpreter&) { return Value::Ref(Type::FloatTexture, FloatTexture::imageTexture(a.at("mapping").as<TextureMapping>(), a.at("image").as<Image2D>())); },
         [def2D, defWorld](const Args& a, Interpreter& in) -> Value {
             const std::string& proc = a.at("procedure").s;
             if (proc == "checker board")
                 return withConfig(a.at("params").tuple(), {{"c0", R}, {"c1", R}, {"mapping", Type::Texture2DMapping, def2D}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::FloatTexture, FloatTexture::checkerBoard(c.at("mapping").as<TextureMapping>(), f(c, "c0"), f(c, "c1")));
                                   });
             if (proc == "voronoi")
                 return withConfig(a.at("params").tuple(), {{"scale", R}, {"valueScale", R, Value::Real(1.0)}, {"flat", Type::Bool, Value::Bool(true)}, {"mapping", Type::Texture3DMapping, defWorld}}, in,
                                   [](const Args& c, Interpreter&) {
                                       return Value::Ref(Type::FloatTexture, FloatTexture::voronoi(c.at("mapping").as<TextureMapping>(), f(c, "scale"), f(c, "valueScale"), c.at("flat").b));
                                   });
             in.fail("Specified procedure is invalid.");
         }}));

    // ---- geometry
    def(in, "createVertex", fn(kVertexSig, makeVertex));
    def(in, "Spectrum", fnOver(
        {{{"type", Type::String}, {"value", R}},
         {{"type", Type::String, Value::Str("Reflectance")}, {"space", Type::String, Value::Str("sRGB")}, {"e0", R}, {"e1", R}, {"e2", R}},
         {{"type", Type::String, Value::Str("Reflectance")}, {"minWL", R}, {"maxWL", R}, {"values", Type::Tuple}},
         {{"type", Type::String, Value::Str("Reflectance")}, {"wls", Type::Tuple}, {"values", Type::Tuple}},
         {{"ID", Type::String}, {"idx", Type::Integer, Value::Int(0)}}},
        {[](const Args& a, Interpreter& in) -> Value {
             SpectrumType t;
             if (!strToSpectrumType(a.at("type").s, &t)) in.fail("Specified spectrum type is invalid.");
             float v = f(a, "value");
             return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, t, ColorSpace::sRGB, v, v, v));
         },
         [](const Args& a, Interpreter& in) -> Value {
             SpectrumType t; ColorSpace c;
             if (!strToSpectrumType(a.at("type").s, &t)) in.fail("Specified spectrum type is invalid.");
             if (!strToColorSpace(a.at("space").s, &c)) in.fail("Specified color space is invalid.");
             return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, t, c, f(a, "e0"), f(a, "e1"), f(a, "e2")));
         },
         [](const Args& a, Interpreter& in) -> Value {
             SpectrumType t;
             if (!strToSpectrumType(a.at("type").s, &t)) in.fail("Specified spectrum type is invalid.");
             const ParameterList& vl = a.at("values").tuple();
             // The reference resizes the vector to n and THEN push_backs the n values, so the first n
             // entries it hands to Spectrum::create are zeros (API.cpp:355-363). Reproduced faithfully.
             size_t n = vl.unnamed.size();
             std::vector<float> values(n, 0.0f);
             for (const Value& v : vl.unnamed) values.push_back((float)v.number());
             return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, t, f(a, "minWL"), f(a, "maxWL"), values.data(), (uint32_t)n));
         },
         [](const Args& a, Interpreter& in) -> Value {
             SpectrumType t;
             if (!strToSpectrumType(a.at("type").s, &t)) in.fail("Specified spectrum type is invalid.");
             const ParameterList& wl = a.at("wls").tuple();
             const ParameterList& vl = a.at("values").tuple();
             size_t n = wl.unnamed.size();
             if (n != vl.unnamed.size()) in.fail("The sizes of the wavelengths and the values are different.");
             std::vector<float> wls(n, 0.0f), values(n, 0.0f);      // same resize-then-push_back behaviour (API.cpp:380-392)
             for (size_t i = 0; i < n; ++i) { wls.push_back((float)wl.unnamed[i].number()); values.push_back((float)vl.unnamed[i].number()); }
             return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, t, wls.data(), values.data(), (uint32_t)n));
         },
         [](const Args& a, Interpreter& in) -> Value {
             const SpectralTables& T = SpectralTables::instance();
             const std::string& id = a.at("ID").s;
             int32_t idx = a.at("idx").i;
             if (id == "D65") {
                 if (idx != 0) in.fail("Specified index is out of range for this spectrum.");
                 const std::vector<float>& d = T.floats("illuminant/D65");
                 return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, SpectrumType::Illuminant, 300.0f, 830.0f, d.data(), (uint32_t)d.size()));
             }
             if (id == "ColorChecker") {
                 if (idx < 0 || idx >= 24) in.fail("Specified index is out of range for this spectrum.");
                 const std::vector<float>& d = T.floats("colorchecker/spectra");
                 return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, SpectrumType::Reflectance, 380.0f, 730.0f, d.data() + 36 * idx, 36));
             }
             if (!T.has("ior/" + id + "/meta")) in.fail("unrecognized spectrum ID.");
             if (idx < 0 || idx >= 2) in.fail("Specified index is out of range.");
             const std::vector<float>& meta = T.floats("ior/" + id + "/meta");
             const std::string key = "ior/" + id + (idx == 0 ? "/etas" : "/ks");
             if (!T.has(key)) in.fail("This IOR doesn't have the spectrum corresponding to the index specified.");
             const std::vector<float>& vals = T.floats(key);
             if (meta[0] == 0.0f)
                 return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, SpectrumType::IndexOfRefraction, meta[2], meta[3], vals.data(), (uint32_t)meta[1]));
             const std::vector<float>& lam = T.floats("ior/" + id + "/lambdas");
             return Value::Ref(Type::Spectrum, Spectrum::create(in.rgbMode, SpectrumType::IndexOfRefraction, lam.data(), vals.data(), (uint32_t)meta[1]));
         }}));
    def(in, "Image2D", fn({{"path", Type::String}, {"mode", Type::String, Value::Str("AsIs")}, {"type", Type::String, Value::Str("Reflectance")}},
                          [](const Args& a, Interpreter& in) -> Value {
                              ImageStoreMode mode;
                              const std::string& ms = a.at("mode").s;
                              if (ms == "AsIs") mode = ImageStoreMode::AsIs;
                              else if (ms == "Normal") mode = ImageStoreMode::NormalTexture;          // strToImageStoreMode, API.cpp:35-45
                              else if (ms == "Alpha") mode = ImageStoreMode::AlphaTexture;
                              else in.fail("Specified image store mode is invalid.");
                              SpectrumType t;
                              if (!strToSpectrumType(a.at("type").s, &t)) in.fail("Specified spectrum type is invalid.");
                              // note: the reference passes `path` as written (relative to the working directory), API.cpp:466
                              return Value::Ref(Type::Image2D, loadImageCached(a.at("path").s, mode, t, in.rgbMode));
                          }));

    def(in, "createSurfaceMaterial", fn({{"type", Type::String}, {"params", Type::Tuple}}, [](const Args& a, Interpreter& in) -> Value {
        const std::string& type = a.at("type").s;
        const ParameterList& p = a.at("params").tuple();
        const Type ST = Type::SpectrumTexture, FT = Type::FloatTexture, SM = Type::SurfaceMaterial;
        auto st = [](const Args& c, const char* n) { return c.at(n).as<SpectrumTexture>(); };
        auto ft = [](const Args& c, const char* n) { return c.at(n).as<FloatTexture>(); };
        auto sm = [](const Args& c, const char* n) { return c.at(n).as<SurfaceMaterial>(); };
        auto ret = [](const SurfaceMaterialRef& m) { return Value::Ref(Type::SurfaceMaterial, m); };
        if (type == "matte")
            return withConfig(p, {{"reflectance", ST}, {"sigma", FT, Value::Ref(FT, FloatTextureRef())}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createMatte(st(c, "reflectance"), ft(c, "sigma"))); });
        if (type == "metal")
            return withConfig(p, {{"coeffR", ST}, {"eta", ST}, {"k", ST}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createMetal(st(c, "coeffR"), st(c, "eta"), st(c, "k"))); });
        if (type == "glass")
            return withConfig(p, {{"coeff", ST}, {"etaExt", ST}, {"etaInt", ST}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createGlass(st(c, "coeff"), st(c, "etaExt"), st(c, "etaInt"))); });
        if (type == "Ward")
            return withConfig(p, {{"R", ST}, {"anisoX", FT}, {"anisoY", FT}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createModifiedWardDur(st(c, "R"), ft(c, "anisoX"), ft(c, "anisoY"))); });
        if (type == "Ashikhmin")
            return withConfig(p, {{"Rd", ST}, {"Rs", ST}, {"nx", FT}, {"ny", FT}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createAshikhminShirley(st(c, "Rd"), st(c, "Rs"), ft(c, "nx"), ft(c, "ny"))); });
        if (type == "microfacet metal")
            return withConfig(p, {{"eta", ST}, {"k", ST}, {"alpha_g", FT}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createMicrofacetMetal(st(c, "eta"), st(c, "k"), ft(c, "alpha_g"))); });
        if (type == "microfacet glass")
            return withConfig(p, {{"etaExt", ST}, {"etaInt", ST}, {"alpha_g", FT}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createMicrofacetGlass(st(c, "etaExt"), st(c, "etaInt"), ft(c, "alpha_g"))); });
        if (type == "inverse")
            return withConfig(p, {{"base", SM}}, in, [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createInverseMaterial(sm(c, "base"))); });
        if (type == "emitter")
            return withConfig(p, {{"scatter", SM}, {"emitter", Type::EmitterSurfaceProperty}}, in, [=](const Args& c, Interpreter&) {
                return ret(SurfaceMaterial::createEmitterSurfaceMaterial(sm(c, "scatter"), c.at("emitter").as<EmitterSurfaceProperty>()));
            });
        if (type == "mix")
            return withConfig(p, {{"mat0", SM}, {"mat1", SM}, {"factor", FT}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createMixedMaterial(sm(c, "mat0"), sm(c, "mat1"), ft(c, "factor"))); });
        if (type == "sum")
            return withConfig(p, {{"mat0", SM}, {"mat1", SM}}, in,
                              [=](const Args& c, Interpreter&) { return ret(SurfaceMaterial::createSummedMaterial(sm(c, "mat0"), sm(c, "mat1"))); });
        in.fail("Specified material type is invalid.");
    }));
    def(in, "createEmitterSurfaceProperty", fn({{"type", Type::String}, {"params", Type::Tuple}}, [](const Args& a, Interpreter& in) -> Value {
        if (a.at("type").s == "diffuse")
            return withConfig(a.at("params").tuple(), {{"emittance", Type::SpectrumTexture}}, in, [](const Args& c, Interpreter&) {
                return Value::Ref(Type::EmitterSurfaceProperty, SurfaceMaterial::createDiffuseEmitter(c.at("emittance").as<SpectrumTexture>()));
            });
        in.fail("Specified material type is invalid.");
    }));

    def(in, "createMesh", fn({{"vertices", Type::Tuple}, {"matGroups", Type::Tuple}}, [](const Args& a, Interpreter& in) -> Value {
        auto mesh = std::make_shared<TriangleMeshNode>();
        for (const Value& v : a.at("vertices").tuple().unnamed) {
            if (v.type == Type::Vertex) { mesh->addVertex(v.vtx); continue; }
            if (v.type != Type::Tuple) in.fail("createMesh: a vertex must be a Vertex or a tuple");
            Args va;
            if (!mapParamsToArgs(v.tuple(), kVertexSig, &va)) in.fail("createMesh: Parameters are invalid.");
            mesh->addVertex(makeVertex(va, in).vtx);
        }
        const std::vector<ArgInfo> sigGroup = {{"mat", Type::SurfaceMaterial},
                                               {"normal", Type::NormalTexture, Value::Ref(Type::NormalTexture, Normal3DTextureRef())},
                                               {"alpha", Type::FloatTexture, Value::Ref(Type::FloatTexture, FloatTextureRef())},
                                               {"triangles", Type::Tuple}};
        for (const Value& g : a.at("matGroups").tuple().unnamed) {
            Args ga;
            if (g.type != Type::Tuple || !mapParamsToArgs(g.tuple(), sigGroup, &ga)) in.fail("createMesh: material group parameters are invalid.");
            std::vector<uint32_t> idx;
            for (const Value& t : ga.at("triangles").tuple().unnamed) {
                Args ta;
                if (t.type != Type::Tuple || !mapParamsToArgs(t.tuple(), {{"v0", Type::Integer}, {"v1", Type::Integer}, {"v2", Type::Integer}}, &ta))
                    in.fail("createMesh: Parameters are invalid.");
                idx.push_back((uint32_t)ta["v0"].i); idx.push_back((uint32_t)ta["v1"].i); idx.push_back((uint32_t)ta["v2"].i);
            }
            mesh->addTriangles(ga.at("mat").as<SurfaceMaterial>(), ga.at("normal").as<Normal3DTexture>(), ga.at("alpha").as<FloatTexture>(), std::move(idx));
        }
        return Value::Ref(Type::Mesh, mesh);
    }));
    def(in, "createNode", fn({}, [](const Args&, Interpreter&) { return Value::Ref(Type::Node, std::make_shared<InternalNode>()); }));
    def(in, "copyNode", fn({{"src", Type::Node}}, [](const Args& a, Interpreter& in) -> Value {
        NodeRef copied;
        try { copied = a.at("src").as<InternalNode>()->copy(); } catch (const std::exception& e) { in.fail(e.what()); }
        return Value::Ref(Type::Node, std::static_pointer_cast<InternalNode>(copied));
    }));
    def(in, "createReferenceNode", fn({{"node", Type::Node}}, [](const Args& a, Interpreter&) {
        return Value::Ref(Type::ReferenceNode, std::make_shared<ReferenceNode>(a.at("node").as<InternalNode>()));
    }));
    def(in, "setTransform", fn({{"node", Type::Node}, {"transform", Type::Transform}}, [](const Args& a, Interpreter&) {
        a.at("node").as<InternalNode>()->setTransform(*a.at("transform").as<StaticTransform>());
        return Value();
    }));
    {
        std::vector<std::vector<ArgInfo>> sigs;
        std::vector<Function::Native> bodies;
        for (Type ct : {Type::Node, Type::ReferenceNode, Type::Mesh, Type::Camera}) {
            sigs.push_back({{"parent", Type::Node}, {"child", ct}});
            bodies.push_back([](const Args& a, Interpreter&) {
                a.at("parent").as<InternalNode>()->addChildNode(a.at("child").as<Node>());
                return Value();
            });
        }
        def(in, "addChild", fnOver(sigs, bodies));
    }
    // scanXZFromYPlus (API.cpp:926-983): a numX x numY grid of rays straight down onto `node` from 1.5 x the top of its bounds;
    // `callback(position, tangent, bitangent, normal)` runs for every hit (RTC3*.txt scatter grass instances with it). The
    // subtree is flattened and its SBVH -> QBVH built on the spot, the rays are cast on the host (host/raycast.cpp).
    def(in, "scanXZFromYPlus", fn({{"node", Type::Node}, {"numX", Type::Integer}, {"numY", Type::Integer}, {"randomness", R, Value::Real(0.0)},
                                   {"callback", Type::Function}}, [](const Args& a, Interpreter& in) -> Value {
        InternalNodeRef node = a.at("node").as<InternalNode>();
        const int numX = a.at("numX").i, numY = a.at("numY").i;
        const float randomness = f(a, "randomness");
        FunctionRef callback = a.at("callback").as<Function>();
        XorShift128 rng(50287412);
        FlatScene flat;
        GpuSceneBuilder b(flat);
        RenderingData data;
        node->resetFlattening();
        struct Reset { Node* n; ~Reset() { n->resetFlattening(); } } reset{node.get()};
        node->getRenderingData(b, nullptr, &data);
        if (data.objects.empty()) in.fail("scanXZFromYPlus: the node has no surfaces");
        const uint32_t top = b.createAggregate(std::move(data.objects));
        const BBox bounds = b.aggregates[top].sbvh.bounds;
        b.finalize(top);
        HostRayCaster caster(flat);
        for (int i = 0; i < numY; ++i) {
            for (int j = 0; j < numX; ++j) {
                const float purturbX = randomness * (rng.float0cTo1o() - 0.5f);
                const float purturbZ = randomness * (rng.float0cTo1o() - 0.5f);
                const Vec3 org(bounds.lo.x + (bounds.hi.x - bounds.lo.x) * (j + 0.5f + purturbX) / numX,
                               bounds.hi.y * 1.5f,
                               bounds.lo.z + (bounds.hi.z - bounds.lo.z) * (i + 0.5f + purturbZ) / numY);
                const Vec3 dir(0, -1, 0);
                HostHit hit;
                if (!caster.intersect(org, dir, 0.0f, INFINITY, &hit)) continue;
                HostSurfacePoint sp;
                caster.surfacePoint(hit, org, dir, &sp);
                ParameterList params;
                params.add("", Value::Vec(Type::Point, sp.p));
                params.add("", Value::Vec(Type::Vector, sp.sx));
                params.add("", Value::Vec(Type::Vector, sp.sy));
                params.add("", Value::Vec(Type::Normal, sp.sz));
                Value r = callback->call(params, in);
                if (r.isError()) in.fail(r.s);
            }
        }
        return Value();
    }));
    def(in, "load3DModel", fn({{"path", Type::String}, {"matProc", Type::Function, Value::Ref(Type::Function, FunctionRef())}}, [](const Args& a, Interpreter& in) -> Value {
        const std::string path = in.sceneDir + a.at("path").s;
        const std::string prefix = path.substr(0, path.find_last_of('/') + 1);
        FunctionRef proc = a.at("matProc").as<Function>();
        assbin::Scene sc;
        std::string err;
        if (!assbin::load(path, &sc, &err)) { std::printf("Failed to load %s.\n", path.c_str()); in.fail("Some errors occur during loading a 3D model: " + err); }
        std::printf("Reading: %s done.\n", path.c_str());
        std::vector<SurfaceAttributes> mats;
        for (const assbin::Material& m : sc.materials) mats.push_back(proc ? userMaterial(m, prefix, proc, in) : defaultMaterial(m, prefix, in));
        if (mats.empty()) mats.push_back(defaultMaterial(assbin::Material(), prefix, in));
        InternalNodeRef node = constructNode(sc, sc.root, mats);
        if (!node) in.fail("Some errors occur during loading a 3D model.");
        node->name = path;
        std::printf("Constructing: %s done.\n", path.c_str());
        return Value::Ref(Type::Node, node);
    }));
    def(in, "createPerspectiveCamera",
        fn({{"sensitivity", R, Value::Real(0.0)}, {"aspect", R, Value::Real(1.0)}, {"fovY", R, Value::Real(0.5235987756)},
            {"radius", R, Value::Real(0.0)}, {"imgDist", R, Value::Real(0.02)}, {"objDist", R, Value::Real(5.0)}},
           [](const Args& a, Interpreter&) {
               auto cam = std::make_shared<PerspectiveCamera>(f(a, "sensitivity"), f(a, "aspect"), f(a, "fovY"), f(a, "radius"), f(a, "imgDist"), f(a, "objDist"));
               return Value::Ref(Type::Camera, std::static_pointer_cast<Node>(std::make_shared<CameraNode>(cam)));
           }));
    def(in, "setRenderer", fn({{"method", Type::String}, {"config", Type::Tuple, Value::Tuple(std::make_shared<ParameterList>())}}, [](const Args& a, Interpreter& in) -> Value {
        const std::string& method = a.at("method").s;
        const ParameterList& cfg = a.at("config").tuple();
        RenderingContext* ctx = in.context;
        if (method == "PT" || method == "BPT") {
            // API.cpp setRenderer: "PT" -> PathTracingRenderer, "BPT" -> BidirectionalPathTracingRenderer; here their GPU twins
            return withConfig(cfg, {{"samples", Type::Integer, Value::Int(8)}}, in, [ctx, method](const Args& c, Interpreter&) {
                ctx->samples = (uint32_t)c.at("samples").i;
                ctx->rendererMethod = method;
                if (method == "BPT") ctx->renderer.reset(new GPUBidirectionalPathTracingRenderer(ctx->samples));
                else ctx->renderer.reset(new GPUPathTracingRenderer(ctx->samples));
                return Value();
            });
        }
        if (method == "debug") {
            // API.cpp:1037-1062: ("outputs": ("geometric normal", "shading normal", "shading tangent", "distance"))
            return withConfig(cfg, {{"outputs", Type::Tuple}}, in, [ctx, method](const Args& c, Interpreter&) {
                const ParameterList& outputs = c.at("outputs").tuple();
                bool flags[GPUDebugRenderer::NumChannels] = {false, false, false, false};
                for (size_t i = 0; i < outputs.unnamed.size(); ++i) {
                    const Value& e = outputs.unnamed[i];
                    if (e.type != Type::String) continue;
                    if (e.s == "geometric normal") flags[GPUDebugRenderer::GeometricNormal] = true;
                    else if (e.s == "shading normal") flags[GPUDebugRenderer::ShadingNormal] = true;
                    else if (e.s == "shading tangent") flags[GPUDebugRenderer::ShadingTangent] = true;
                    else if (e.s == "distance") flags[GPUDebugRenderer::Distance] = true;
                }
                ctx->rendererMethod = method;
                ctx->renderer.reset(new GPUDebugRenderer(flags));
                return Value();
            });
        }
        in.fail("Unknown method is specified.");
    }));
    def(in, "setRenderSettings",
        fn({{"width", Type::Integer, Value::Int(1024)}, {"height", Type::Integer, Value::Int(1024)}, {"timeStart", R, Value::Real(0.0)},
            {"timeEnd", R, Value::Real(0.0)}, {"brightness", R, Value::Real(1.0)}, {"rngSeed", Type::Integer, Value::Int(1509761209)}},
           [](const Args& a, Interpreter& in) {
               RenderingContext* c = in.context;
               c->width = a.at("width").i; c->height = a.at("height").i;
               c->timeStart = f(a, "timeStart"); c->timeEnd = f(a, "timeEnd");
               c->brightness = f(a, "brightness"); c->rngSeed = a.at("rngSeed").i;
               return Value();
           }));
    def(in, "setEnvironment", fn({{"path", Type::String}, {"scale", R, Value::Real(1.0)}}, [](const Args& a, Interpreter& in) {
        Image2DRef img = loadImageCached(in.sceneDir + a.at("path").s, ImageStoreMode::AsIs, SpectrumType::Illuminant, in.rgbMode);
        SpectrumTextureRef tex = SpectrumTexture::imageTexture(sharedTexCoordMapping(), img);
        in.scene->setEnvNode(std::make_shared<InfiniteSphereNode>(std::make_shared<IBLEmission>(tex, f(a, "scale"))));
        return Value();
    }));
}

}  // namespace lang
}  // namespace slr
