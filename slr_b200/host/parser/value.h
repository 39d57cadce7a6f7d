// Dynamic values of the SLR scene-description language.
// Same type lattice and implicit conversions as the reference's interpreter
// (libSLRSceneGraph/Parser/SceneParser.hpp:17-47, SceneParser.cpp:497-890), written as a plain
// tagged struct instead of type-erased shared_ptr<void> + per-type function tables.
#pragma once
#include "../scene.h"
#include "../shading.h"
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace slr {
namespace lang {

enum class Type : uint32_t {
    Bool = 0, Integer, RealNumber, String, Point, Vector, Normal, Matrix, Vertex, Transform, Spectrum, Image2D,
    Texture2DMapping, Texture3DMapping, SpectrumTexture, NormalTexture, FloatTexture, SurfaceMaterial,
    EmitterSurfaceProperty, Mesh, Camera, Node, ReferenceNode, Tuple, Function, Any, Void, Error
};
const char* typeName(Type t);

struct Value;
struct ParameterList {
    std::map<std::string, Value> named;
    std::vector<Value> unnamed;
    bool add(const std::string& key, const Value& v);
    size_t numParams() const { return named.size() + unnamed.size(); }
};
typedef std::shared_ptr<ParameterList> ParameterListRef;

class Function;
typedef std::shared_ptr<Function> FunctionRef;

struct Value {
    Type type = Type::Void;
    bool b = false;
    int32_t i = 0;
    double d = 0.0;
    std::string s;             // String / Error message
    Vec3 v3;                   // Point / Vector / Normal
    Mat4 m;                    // Matrix
    Vertex vtx;                // Vertex
    std::shared_ptr<void> ref; // every reference type

    Value() {}
    static Value Bool(bool v) { Value r; r.type = Type::Bool; r.b = v; return r; }
    static Value Int(int32_t v) { Value r; r.type = Type::Integer; r.i = v; return r; }
    static Value Real(double v) { Value r; r.type = Type::RealNumber; r.d = v; return r; }
    static Value Str(const std::string& v) { Value r; r.type = Type::String; r.s = v; return r; }
    static Value Error(const std::string& msg) { Value r; r.type = Type::Error; r.s = msg; return r; }
    static Value Vec(Type t, const Vec3& v) { Value r; r.type = t; r.v3 = v; return r; }
    static Value Matrix(const Mat4& mm) { Value r; r.type = Type::Matrix; r.m = mm; return r; }
    static Value Vtx(const Vertex& v) { Value r; r.type = Type::Vertex; r.vtx = v; return r; }
    template <typename T> static Value Ref(Type t, const std::shared_ptr<T>& p) { Value r; r.type = t; r.ref = p; return r; }
    static Value Tuple(const ParameterListRef& p) { return Ref(Type::Tuple, p); }

    template <typename T> std::shared_ptr<T> as() const { return std::static_pointer_cast<T>(ref); }
    const ParameterList& tuple() const { return *static_cast<const ParameterList*>(ref.get()); }
    bool isError() const { return type == Type::Error; }

    bool convertibleTo(Type t) const;
    Value convertTo(Type t) const;          // precondition: convertibleTo(t)
    double number() const { return type == Type::RealNumber ? d : type == Type::Integer ? (double)i : (double)b; }
    std::string toString() const;
};

struct ArgInfo {
    std::string name;
    Type expected = Type::Any;
    Value defaultValue;       // Void = required
    ArgInfo() {}
    ArgInfo(const std::string& n, Type t) : name(n), expected(t) {}
    ArgInfo(const std::string& n, Type t, const Value& d) : name(n), expected(t), defaultValue(d) {}
};
typedef std::map<std::string, Value> Args;

// Named parameters bind by key, unnamed ones to the first still-free argument whose type accepts
// them, defaults fill the rest (SceneParser.cpp:399-453).
bool mapParamsToArgs(const ParameterList& params, const std::vector<ArgInfo>& signature, Args* args);

struct Interpreter;
struct Statement;
typedef std::shared_ptr<Statement> StatementRef;

class Function {
public:
    typedef std::function<Value(const Args&, Interpreter&)> Native;
    std::vector<std::vector<ArgInfo>> signatures;
    std::vector<Native> natives;       // one per signature, or
    StatementRef body;                 // user-defined (single signature)
    Function() {}
    Function(const std::vector<ArgInfo>& sig, const Native& fn) : signatures{sig}, natives{fn} {}
    Function(const std::vector<std::vector<ArgInfo>>& sigs, const std::vector<Native>& fns) : signatures(sigs), natives(fns) {}
    Value call(const ParameterList& params, Interpreter& in) const;
};

}  // namespace lang
}  // namespace slr
