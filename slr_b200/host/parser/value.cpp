#include "interp.h"
#include <sstream>

namespace slr {
namespace lang {

const char* typeName(Type t) {
    static const char* names[] = {"Bool", "Integer", "RealNumber", "String", "Point", "Vector", "Normal", "Matrix", "Vertex",
                                  "Transform", "Spectrum", "Image2D", "Texture2DMapping", "Texture3DMapping", "SpectrumTexture",
                                  "NormalTexture", "FloatTexture", "SurfaceMaterial", "EmitterSurfaceProperty", "Mesh", "Camera",
                                  "Node", "ReferenceNode", "Tuple", "Function", "Any", "Void", "Error"};
    return names[(uint32_t)t];
}

bool ParameterList::add(const std::string& key, const Value& v) {
    if (key.empty()) { unnamed.push_back(v); return true; }
    if (named.count(key)) return false;
    named[key] = v;
    return true;
}

bool Value::convertibleTo(Type t) const {
    if (t == type) return true;
    switch (type) {
        case Type::Bool: return t == Type::Integer || t == Type::RealNumber;
        case Type::Integer: return t == Type::Bool || t == Type::RealNumber;
        case Type::Normal: return t == Type::Vector;
        case Type::Matrix: return t == Type::Transform;
        default: return false;
    }
}

Value Value::convertTo(Type t) const {
    if (t == type) return *this;
    switch (type) {
        case Type::Bool: return t == Type::Integer ? Int(b ? 1 : 0) : Real(b ? 1.0 : 0.0);
        case Type::Integer: return t == Type::Bool ? Bool(i != 0) : Real((double)i);
        case Type::Normal: return Vec(Type::Vector, v3);
        case Type::Matrix: return Ref(Type::Transform, std::make_shared<StaticTransform>(m));
        default: return Error("invalid conversion");
    }
}

std::string Value::toString() const {
    std::ostringstream o;
    switch (type) {
        case Type::Bool: o << b; break;
        case Type::Integer: o << i; break;
        case Type::RealNumber: o << d; break;
        case Type::String: o << '"' << s << '"'; break;
        case Type::Tuple: {
            const ParameterList& p = tuple();
            o << "(";
            bool first = true;
            for (const auto& kv : p.named) { o << (first ? "" : ", ") << '"' << kv.first << "\": " << kv.second.toString(); first = false; }
            for (const auto& v : p.unnamed) { o << (first ? "" : ", ") << v.toString(); first = false; }
            o << ")";
            break;
        }
        default: o << typeName(type); break;
    }
    return o.str();
}

bool mapParamsToArgs(const ParameterList& params, const std::vector<ArgInfo>& sig, Args* args) {
    args->clear();
    std::vector<bool> assigned(sig.size(), false);
    for (const auto& kv : params.named) {
        size_t idx = sig.size();
        for (size_t k = 0; k < sig.size(); ++k)
            if (sig[k].name == kv.first && (sig[k].expected == Type::Any || kv.second.convertibleTo(sig[k].expected))) { idx = k; break; }
        if (idx == sig.size()) { args->clear(); return false; }
        (*args)[sig[idx].name] = sig[idx].expected == Type::Any ? kv.second : kv.second.convertTo(sig[idx].expected);
        assigned[idx] = true;
    }
    for (const Value& v : params.unnamed) {
        size_t idx = sig.size();
        for (size_t k = 0; k < sig.size(); ++k)
            if (!assigned[k] && (sig[k].expected == Type::Any || v.convertibleTo(sig[k].expected))) { idx = k; break; }
        if (idx == sig.size()) { args->clear(); return false; }
        (*args)[sig[idx].name] = sig[idx].expected == Type::Any ? v : v.convertTo(sig[idx].expected);
        assigned[idx] = true;
    }
    for (size_t k = 0; k < sig.size(); ++k) {
        if (assigned[k]) continue;
        if (sig[k].defaultValue.type == Type::Void) { args->clear(); return false; }
        (*args)[sig[k].name] = sig[k].defaultValue;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// operators (SceneParser.cpp:551-890): the LEFT operand's type picks the rule
// ---------------------------------------------------------------------------------------------

static bool isNumeric(const Value& v) { return v.type == Type::Bool || v.type == Type::Integer || v.type == Type::RealNumber; }

Value opUnary(const std::string& op, const Value& v) {
    if (op == "+") {
        if (isNumeric(v) || v.type == Type::Point || v.type == Type::Vector || v.type == Type::Normal || v.type == Type::Matrix) return v;
        return Value::Error("Type does not have the unary + operator definition.");
    }
    if (op == "-") {
        switch (v.type) {
            case Type::Bool: return Value::Int(-(int)v.b);
            case Type::Integer: return Value::Int(-v.i);
            case Type::RealNumber: return Value::Real(-v.d);
            case Type::Point: case Type::Vector: case Type::Normal: return Value::Vec(v.type, -v.v3);
            case Type::Matrix: {
                Mat4 n;
                for (int c = 0; c < 4; ++c) n.c[c] = Vec4(-v.m.c[c].x, -v.m.c[c].y, -v.m.c[c].z, -v.m.c[c].w);
                return Value::Matrix(n);
            }
            default: return Value::Error("Type does not have the unary - operator definition.");
        }
    }
    if (op == "!") {
        if (isNumeric(v)) return Value::Bool(v.number() == 0.0);
        return Value::Error("Type does not have the ! operator definition.");
    }
    return Value::Error("unknown unary operator " + op);
}

static Value arith(const std::string& op, const Value& l, const Value& r) {
    // Real on the left => real arithmetic; Bool/Integer on the left => integer arithmetic when the
    // right side converts to Integer (Bool, Integer), else real.
    const bool leftReal = l.type == Type::RealNumber;
    if (!isNumeric(r)) return Value::Error(op + " operator does not support the right operand type.");
    const bool intOp = !leftReal && (r.type == Type::Integer || r.type == Type::Bool);
    if (intOp) {
        int32_t a = l.type == Type::Bool ? (int32_t)l.b : l.i;
        int32_t b = r.type == Type::Bool ? (int32_t)r.b : r.i;
        if (op == "+") return Value::Int(a + b);
        if (op == "-") return Value::Int(a - b);
        if (op == "*") return Value::Int(a * b);
        if (op == "/") { if (b == 0) return Value::Error("integer division by zero"); return Value::Int(a / b); }
        if (op == "%") { if (b == 0) return Value::Error("integer remainder by zero"); return Value::Int(a % b); }
        if (op == "<") return Value::Bool(a < b);
        if (op == ">") return Value::Bool(a > b);
        // Reference quirk kept (a scene file must mean here what it means there): Integer == converts the RIGHT operand to
        // Bool before comparing (SceneParser.cpp:699-706, `lVal == v1.asRaw<TypeMap::Bool>()`), so 3 == 3 is false and
        // 1 == 7 is true; <=, >= and != are derived from it (SceneParser.cpp:515-527).
        if (op == "==") return Value::Bool(a == (b != 0 ? 1 : 0));
    } else {
        if (op == "%") return Value::Error("% operator does not support the right operand type.");
        double a = l.number(), b = r.number();
        if (op == "+") return Value::Real(a + b);
        if (op == "-") return Value::Real(a - b);
        if (op == "*") return Value::Real(a * b);
        if (op == "/") return Value::Real(a / b);
        if (op == "<") return Value::Bool(a < b);
        if (op == ">") return Value::Bool(a > b);
        if (op == "==") return Value::Bool(a == b);
    }
    return Value::Error("unknown operator " + op);
}

Value opBinary(const std::string& op, const Value& l, const Value& r) {
    if (op == "&&" || op == "||") {
        if (!l.convertibleTo(Type::Bool)) return Value::Error("Left operand cannot be converted to a bool value.");
        if (!r.convertibleTo(Type::Bool)) return Value::Error("Right operand cannot be converted to a bool value.");
        bool a = l.convertTo(Type::Bool).b, b = r.convertTo(Type::Bool).b;
        return Value::Bool(op == "&&" ? (a && b) : (a || b));
    }
    if (op == "<=" || op == ">=" || op == "!=") {
        Value eq = opBinary("==", l, r);
        if (eq.isError()) return eq;
        if (op == "!=") return Value::Bool(!eq.b);
        Value rel = opBinary(op == "<=" ? "<" : ">", l, r);
        if (rel.isError()) return rel;
        return Value::Bool(eq.b || rel.b);
    }
    switch (l.type) {
        case Type::Bool: case Type::Integer: case Type::RealNumber:
            if (op == "==" && l.type == Type::Bool) {
                if (!r.convertibleTo(Type::Bool)) return Value::Error("== operator does not support the right operand type.");
                return Value::Bool(l.b == r.convertTo(Type::Bool).b);
            }
            return arith(op, l, r);
        case Type::String:
            if (r.type != Type::String) return Value::Error(op + " operator does not support the right operand type.");
            if (op == "+") return Value::Str(l.s + r.s);
            if (op == "==") return Value::Bool(l.s == r.s);
            break;
        case Type::Point: case Type::Vector: case Type::Normal:
            if (op == "==" && r.convertibleTo(l.type)) return Value::Bool(l.v3 == r.convertTo(l.type).v3);
            break;
        case Type::Matrix:
            if (op == "*") {
                if (isNumeric(r)) {
                    float s = (float)r.number();
                    Mat4 o;
                    for (int c = 0; c < 4; ++c) o.c[c] = Vec4(l.m.c[c].x * s, l.m.c[c].y * s, l.m.c[c].z * s, l.m.c[c].w * s);
                    return Value::Matrix(o);
                }
                if (r.type == Type::Vertex) {
                    Vertex v;
                    v.position = l.m.mulPoint(r.vtx.position);
                    v.normal = l.m.mulVector(r.vtx.normal);       // Matrix4x4 * Normal3D converts through Vector3D
                    v.tangent = l.m.mulVector(r.vtx.tangent);
                    v.texCoord = r.vtx.texCoord;
                    return Value::Vtx(v);
                }
                if (r.type == Type::Matrix) return Value::Matrix(l.m * r.m);
                return Value::Error("* operator does not support the right operand type.");
            }
            if (op == "==" && r.type == Type::Matrix) return Value::Bool(l.m == r.m);
            break;
        case Type::Spectrum:
            if (op == "*") {
                if (!isNumeric(r)) return Value::Error("* operator does not support the right operand type.");
                return Value::Ref(Type::Spectrum, l.as<InputSpectrum>()->createScaled((float)r.number()));
            }
            break;
        default: break;
    }
    return Value::Error(std::string("Left type (") + typeName(l.type) + ") does not have the " + op + " operator definition.");
}

}  // namespace lang
}  // namespace slr
