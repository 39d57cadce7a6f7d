// AST + interpreter state of the scene language (internal header).
#pragma once
#include "value.h"
#include "../renderer.h"

namespace slr {
namespace lang {

struct Expr;
typedef std::shared_ptr<Expr> ExprRef;

struct Param { ExprRef key, value; };

struct Expr {
    enum Kind { Literal, Variable, Tuple, Call, Index, Unary, Binary, Assign, IncDec } kind;
    Value literal;                 // Literal
    std::string name;              // Variable / Call / Assign / IncDec target
    std::string op;                // Unary / Binary / Assign / IncDec ("++*", "*++", ...)
    std::vector<Param> params;     // Tuple / Call
    ExprRef a, b;                  // operands
    int line = 0;
};

struct Statement {
    enum Kind { ExprStmt, Block, If, For, FuncDef, Return } kind;
    ExprRef expr, cond, post;                 // ExprStmt / If+For cond / For pre(expr), post
    std::vector<StatementRef> body;           // Block
    StatementRef thenStmt, elseStmt;          // If / For body (thenStmt) / FuncDef body
    std::string name;                         // FuncDef
    std::vector<std::pair<std::string, ExprRef>> argDefs;
    int line = 0;
};

struct RuntimeError { std::string message; int line; };

struct Interpreter {
    // frames of scopes: frame 0 is the global one; a function call pushes a frame
    std::vector<std::vector<std::map<std::string, Value>>> frames;
    Value returnValue;
    bool returnFlag = false;
    std::string sceneDir;          // absolute directory of the scene file, with trailing '/'
    Scene* scene = nullptr;
    RenderingContext* context = nullptr;
    bool rgbMode = false;

    Interpreter();
    bool lookup(const std::string& name, Value* out) const;
    Value* lookupInCurrentFrame(const std::string& name);
    void define(const std::string& name, const Value& v) { frames.back().back()[name] = v; }
    void defineGlobal(const std::string& name, const Value& v) { frames[0][0][name] = v; }

    Value eval(const ExprRef& e);
    void exec(const StatementRef& s);
    Value callFunction(const std::string& name, const ParameterList& params, int line);
    [[noreturn]] void fail(const std::string& msg, int line = 0) const { throw RuntimeError{msg, line}; }
};

std::vector<StatementRef> parseProgram(const std::string& source, const std::string& fileName);   // throws RuntimeError
void registerBuiltins(Interpreter& in);

// operators on values (Error values carry the message)
Value opUnary(const std::string& op, const Value& v);
Value opBinary(const std::string& op, const Value& l, const Value& r);

}  // namespace lang
}  // namespace slr
