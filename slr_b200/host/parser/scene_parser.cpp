// Lexer, parser and evaluator of the SLR scene language.
// Grammar and token rules restated from libSLRSceneGraph/Parser/SceneLexer.l:20-35 and
// SceneParser.yy:101-260 (operator precedence: = family < || < && < == != < relational < + - <
// * / % < prefix < postfix; tuples "(,)", "(x,)", "(a, b, ...)"; "key": value parameters).
#include "scene_parser.h"
#include "interp.h"
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <unistd.h>

namespace slr {
namespace lang {

// ---------------------------------------------------------------------------------------------
// lexer
// ---------------------------------------------------------------------------------------------
struct Token {
    enum Kind { End, Bool, Integer, Real, String, Id, Keyword, Op, Punct } kind = End;
    std::string text;
    bool b = false;
    int32_t i = 0;
    double d = 0;
    int line = 1;
};

static std::vector<Token> tokenize(const std::string& src) {
    std::vector<Token> out;
    size_t p = 0;
    int line = 1;
    auto err = [&](const std::string& m) { throw RuntimeError{m, line}; };
    while (p < src.size()) {
        char c = src[p];
        if (c == '\n') { ++line; ++p; continue; }
        if (std::isspace((unsigned char)c)) { ++p; continue; }
        if (c == '/' && p + 1 < src.size() && src[p + 1] == '/') { while (p < src.size() && src[p] != '\n') ++p; continue; }
        if (c == '/' && p + 1 < src.size() && src[p + 1] == '*') {
            p += 2;
            while (p + 1 < src.size() && !(src[p] == '*' && src[p + 1] == '/')) { if (src[p] == '\n') ++line; ++p; }
            if (p + 1 >= src.size()) err("unterminated comment");
            p += 2;
            continue;
        }
        Token t;
        t.line = line;
        if (c == '"') {
            size_t q = p + 1;
            std::string s;
            while (q < src.size() && src[q] != '"') {
                if (src[q] == '\n') err("Irregal literal: newline in string");
                if (src[q] == '\\' && q + 1 < src.size() && src[q + 1] == '"') { s += "\\\""; q += 2; continue; }
                s += src[q++];
            }
            if (q >= src.size()) err("unterminated string");
            t.kind = Token::String; t.text = s;
            p = q + 1;
            out.push_back(t);
            continue;
        }
        if (std::isdigit((unsigned char)c) || (c == '.' && p + 1 < src.size() && std::isdigit((unsigned char)src[p + 1]))) {
            size_t q = p;
            bool isReal = false;
            while (q < src.size() && std::isdigit((unsigned char)src[q])) ++q;
            if (q < src.size() && src[q] == '.') { isReal = true; ++q; while (q < src.size() && std::isdigit((unsigned char)src[q])) ++q; }
            if (q < src.size() && (src[q] == 'e' || src[q] == 'E')) {
                size_t r = q + 1;
                if (r < src.size() && (src[r] == '+' || src[r] == '-')) ++r;
                if (r < src.size() && std::isdigit((unsigned char)src[r])) {
                    while (r < src.size() && std::isdigit((unsigned char)src[r])) ++r;
                    isReal = true; q = r;
                }
            }
            t.text = src.substr(p, q - p);
            if (isReal) { t.kind = Token::Real; t.d = std::atof(t.text.c_str()); }
            else        { t.kind = Token::Integer; t.i = std::atoi(t.text.c_str()); }
            p = q;
            out.push_back(t);
            continue;
        }
        if (std::isalpha((unsigned char)c) || c == '_') {
            size_t q = p;
            while (q < src.size() && (std::isalnum((unsigned char)src[q]) || src[q] == '_')) ++q;
            t.text = src.substr(p, q - p);
            if (t.text == "True" || t.text == "true" || t.text == "False" || t.text == "false") {
                t.kind = Token::Bool; t.b = (t.text == "True" || t.text == "true");
            } else if (t.text == "if" || t.text == "else" || t.text == "for" || t.text == "function" || t.text == "return") {
                t.kind = Token::Keyword;
            } else {
                t.kind = Token::Id;
            }
            p = q;
            out.push_back(t);
            continue;
        }
        static const char* ops2[] = {"<=", ">=", "==", "!=", "&&", "||", "+=", "-=", "*=", "/=", "%=", "++", "--"};
        bool matched = false;
        for (const char* o : ops2)
            if (src.compare(p, 2, o) == 0) { t.kind = Token::Op; t.text = o; p += 2; matched = true; break; }
        if (!matched) {
            if (std::string("<>=+-*/%!").find(c) != std::string::npos) { t.kind = Token::Op; t.text = std::string(1, c); ++p; }
            else if (std::string(",:;(){}[]").find(c) != std::string::npos) { t.kind = Token::Punct; t.text = std::string(1, c); ++p; }
            else err(std::string("Irregal character: (") + c + ")");
        }
        out.push_back(t);
    }
    Token e; e.kind = Token::End; e.line = line;
    out.push_back(e);
    return out;
}

// ---------------------------------------------------------------------------------------------
// parser
// ---------------------------------------------------------------------------------------------
struct Parser {
    std::vector<Token> toks;
    size_t pos = 0;
    const Token& peek(size_t k = 0) const { return toks[std::min(pos + k, toks.size() - 1)]; }
    bool isPunct(const char* s, size_t k = 0) const { return peek(k).kind == Token::Punct && peek(k).text == s; }
    bool isOp(const char* s, size_t k = 0) const { return peek(k).kind == Token::Op && peek(k).text == s; }
    bool isKw(const char* s) const { return peek().kind == Token::Keyword && peek().text == s; }
    [[noreturn]] void fail(const std::string& m) const { throw RuntimeError{"syntax error: " + m + " near '" + peek().text + "'", peek().line}; }
    void expectPunct(const char* s) { if (!isPunct(s)) fail(std::string("expected '") + s + "'"); ++pos; }

    ExprRef mk(Expr::Kind k) { auto e = std::make_shared<Expr>(); e->kind = k; e->line = peek().line; return e; }

    std::vector<StatementRef> program() {
        std::vector<StatementRef> out;
        while (peek().kind != Token::End) out.push_back(statement());
        return out;
    }

    StatementRef statement() {
        auto s = std::make_shared<Statement>();
        s->line = peek().line;
        if (isPunct("{")) {
            ++pos;
            s->kind = Statement::Block;
            while (!isPunct("}")) { if (peek().kind == Token::End) fail("unterminated block"); s->body.push_back(statement()); }
            ++pos;
            return s;
        }
        if (isKw("if")) {
            ++pos; expectPunct("(");
            s->kind = Statement::If; s->cond = expression(); expectPunct(")");
            s->thenStmt = statement();
            if (isKw("else")) { ++pos; s->elseStmt = statement(); }
            return s;
        }
        if (isKw("for")) {
            ++pos; expectPunct("(");
            s->kind = Statement::For;
            s->expr = expression(); expectPunct(";");
            s->cond = expression(); expectPunct(";");
            s->post = expression(); expectPunct(")");
            s->thenStmt = statement();
            return s;
        }
        if (isKw("function")) {
            ++pos;
            if (peek().kind != Token::Id) fail("expected a function name");
            s->kind = Statement::FuncDef; s->name = toks[pos++].text;
            expectPunct("(");
            while (!isPunct(")")) {
                if (peek().kind != Token::Id) fail("expected an argument name");
                std::string an = toks[pos++].text;
                ExprRef def;
                if (isOp("=")) { ++pos; def = expression(); }
                s->argDefs.emplace_back(an, def);
                if (isPunct(",")) ++pos; else if (!isPunct(")")) fail("expected ',' or ')'");
            }
            ++pos;
            s->thenStmt = statement();
            return s;
        }
        if (isKw("return")) {
            ++pos;
            s->kind = Statement::Return;
            if (!isPunct(";")) s->expr = expression();
            expectPunct(";");
            return s;
        }
        s->kind = Statement::ExprStmt;
        s->expr = expression();
        expectPunct(";");
        return s;
    }

    // assignment (right associative, lowest precedence) only with an identifier on the left
    ExprRef expression() {
        if (peek().kind == Token::Id && peek(1).kind == Token::Op) {
            const std::string& o = peek(1).text;
            if (o == "=" || o == "+=" || o == "-=" || o == "*=" || o == "/=" || o == "%=") {
                auto e = mk(Expr::Assign);
                e->name = toks[pos].text; e->op = o;
                pos += 2;
                e->a = expression();
                return e;
            }
        }
        return binary(0);
    }

    static int precedence(const std::string& o) {
        if (o == "||") return 1;
        if (o == "&&") return 2;
        if (o == "==" || o == "!=") return 3;
        if (o == "<" || o == ">" || o == "<=" || o == ">=") return 4;
        if (o == "+" || o == "-") return 5;
        if (o == "*" || o == "/" || o == "%") return 6;
        return -1;
    }

    ExprRef binary(int minPrec) {
        ExprRef left = unary();
        while (peek().kind == Token::Op) {
            int pr = precedence(peek().text);
            if (pr < 0 || pr < minPrec) break;
            auto e = mk(Expr::Binary);
            e->op = toks[pos++].text;
            e->a = left;
            e->b = binary(pr + 1);
            left = e;
        }
        return left;
    }

    ExprRef unary() {
        if (isOp("+") || isOp("-") || isOp("!")) {
            auto e = mk(Expr::Unary);
            e->op = toks[pos++].text;
            e->a = postfix();
            return e;
        }
        if ((isOp("++") || isOp("--")) && peek(1).kind == Token::Id) {
            auto e = mk(Expr::IncDec);
            e->op = toks[pos].text + "*";
            e->name = toks[pos + 1].text;
            pos += 2;
            return e;
        }
        return postfix();
    }

    ExprRef postfix() {
        if (peek().kind == Token::Id && (isOp("++", 1) || isOp("--", 1))) {
            auto e = mk(Expr::IncDec);
            e->name = toks[pos].text;
            e->op = "*" + toks[pos + 1].text;
            pos += 2;
            return e;
        }
        ExprRef e = primary();
        while (isPunct("[")) {
            auto ix = mk(Expr::Index);
            ++pos;
            ix->a = e; ix->b = expression();
            expectPunct("]");
            e = ix;
        }
        return e;
    }

    Param parameter() {
        Param p;
        ExprRef first = expression();
        if (isPunct(":")) { ++pos; p.key = first; p.value = expression(); }
        else p.value = first;
        return p;
    }

    ExprRef primary() {
        const Token& t = peek();
        switch (t.kind) {
            case Token::Bool: { auto e = mk(Expr::Literal); e->literal = Value::Bool(t.b); ++pos; return e; }
            case Token::Integer: { auto e = mk(Expr::Literal); e->literal = Value::Int(t.i); ++pos; return e; }
            case Token::Real: { auto e = mk(Expr::Literal); e->literal = Value::Real(t.d); ++pos; return e; }
            case Token::String: { auto e = mk(Expr::Literal); e->literal = Value::Str(t.text); ++pos; return e; }
            case Token::Id: {
                if (isPunct("(", 1)) {
                    auto e = mk(Expr::Call);
                    e->name = t.text;
                    pos += 2;
                    while (!isPunct(")")) {
                        e->params.push_back(parameter());
                        if (isPunct(",")) ++pos;
                        else if (!isPunct(")")) fail("expected ',' or ')'");
                    }
                    ++pos;
                    return e;
                }
                auto e = mk(Expr::Variable);
                e->name = t.text;
                ++pos;
                return e;
            }
            case Token::Punct:
                if (t.text == "(") {
                    ++pos;
                    if (isPunct(",") && isPunct(")", 1)) { pos += 2; return mk(Expr::Tuple); }     // "(,)"
                    Param first = parameter();
                    if (isPunct(")") && !first.key) {                                             // "(expr)"
                        ++pos;
                        return first.value;
                    }
                    auto e = mk(Expr::Tuple);
                    e->params.push_back(first);
                    while (isPunct(",")) {
                        ++pos;
                        if (isPunct(")")) break;                                                  // "(x,)"
                        e->params.push_back(parameter());
                    }
                    expectPunct(")");
                    return e;
                }
                break;
            default: break;
        }
        fail("unexpected token");
    }
};

std::vector<StatementRef> parseProgram(const std::string& source, const std::string&) {
    Parser p;
    p.toks = tokenize(source);
    return p.program();
}

// ---------------------------------------------------------------------------------------------
// interpreter
// ---------------------------------------------------------------------------------------------
Interpreter::Interpreter() { frames.emplace_back(); frames.back().emplace_back(); }

bool Interpreter::lookup(const std::string& name, Value* out) const {
    for (size_t f = frames.size(); f-- > 0;)
        for (size_t s = frames[f].size(); s-- > 0;) {
            auto it = frames[f][s].find(name);
            if (it != frames[f][s].end()) { *out = it->second; return true; }
        }
    return false;
}

Value* Interpreter::lookupInCurrentFrame(const std::string& name) {
    auto& fr = frames.back();
    for (size_t s = fr.size(); s-- > 0;) {
        auto it = fr[s].find(name);
        if (it != fr[s].end()) return &it->second;
    }
    return nullptr;
}

Value Function::call(const ParameterList& params, Interpreter& in) const {
    Args args;
    for (size_t k = 0; k < signatures.size(); ++k) {
        if (!mapParamsToArgs(params, signatures[k], &args)) continue;
        in.frames.emplace_back();
        in.frames.back().emplace_back();
        for (const auto& kv : args) in.frames.back().back()[kv.first] = kv.second;
        Value ret;
        in.returnFlag = false;
        try {
            if (body) { in.exec(body); ret = in.returnValue; }
            else ret = natives[k](args, in);
        } catch (...) { in.frames.pop_back(); throw; }
        in.returnValue = Value();
        in.returnFlag = false;
        in.frames.pop_back();
        return ret;
    }
    return Value::Error("Parameters are invalid.");
}

Value Interpreter::callFunction(const std::string& name, const ParameterList& params, int line) {
    Value f;
    if (!lookup(name, &f) || f.type != Type::Function) fail("Function " + name + " is not defined.", line);
    Value r = f.as<Function>()->call(params, *this);
    if (r.isError()) fail(name + ": " + r.s, line);
    return r;
}

static ParameterListRef evalParams(Interpreter& in, const std::vector<Param>& ps, int line) {
    auto list = std::make_shared<ParameterList>();
    for (const Param& p : ps) {
        std::string key;
        if (p.key) {
            Value k = in.eval(p.key);
            if (k.type != Type::String) in.fail("Key expression must results in string type.", line);
            key = k.s;
        }
        list->add(key, in.eval(p.value));
    }
    return list;
}

Value Interpreter::eval(const ExprRef& e) {
    switch (e->kind) {
        case Expr::Literal: return e->literal;
        case Expr::Variable: {
            Value v;
            if (!lookup(e->name, &v)) fail("Undefined variable is used: " + e->name, e->line);
            return v;
        }
        case Expr::Tuple: return Value::Tuple(evalParams(*this, e->params, e->line));
        case Expr::Call: {
            ParameterListRef ps = evalParams(*this, e->params, e->line);
            return callFunction(e->name, *ps, e->line);
        }
        case Expr::Index: {
            Value t = eval(e->a), ix = eval(e->b);
            if (t.type != Type::Tuple) fail("Element access operator [] cannot be used to non tuple value.", e->line);
            const ParameterList& pl = t.tuple();
            if (ix.convertibleTo(Type::Integer)) {
                int32_t k = ix.convertTo(Type::Integer).i;
                if (k < 0 || (size_t)k >= pl.unnamed.size()) fail("Index value is out or range.", e->line);
                return pl.unnamed[k];
            }
            if (ix.type == Type::String) {
                auto it = pl.named.find(ix.s);
                if (it == pl.named.end()) fail("Index value is invalid.", e->line);
                return it->second;
            }
            fail("Index value must be integer or string compatible type.", e->line);
        }
        case Expr::Unary: {
            Value r = opUnary(e->op, eval(e->a));
            if (r.isError()) fail(r.s, e->line);
            return r;
        }
        case Expr::Binary: {
            Value l = eval(e->a), r = eval(e->b);
            Value out = opBinary(e->op, l, r);
            if (out.isError()) fail(out.s, e->line);
            return out;
        }
        case Expr::Assign: {
            Value rhs = eval(e->a);
            Value any;
            if (e->op != "=" && !lookup(e->name, &any)) fail("Undefined variable: " + e->name, e->line);
            Value* slot = lookupInCurrentFrame(e->name);
            if (!slot) { define(e->name, Value()); slot = lookupInCurrentFrame(e->name); if (e->op != "=") *slot = any; }
            if (e->op == "=") *slot = rhs;
            else {
                Value out = opBinary(e->op.substr(0, 1), *slot, rhs);
                if (out.isError()) fail(out.s, e->line);
                *slot = out;
            }
            return *slot;
        }
        case Expr::IncDec: {
            Value any;
            if (!lookup(e->name, &any)) fail("Undefined variable: " + e->name, e->line);
            Value* slot = lookupInCurrentFrame(e->name);
            if (!slot) { define(e->name, any); slot = lookupInCurrentFrame(e->name); }
            if (slot->type != Type::Integer && slot->type != Type::RealNumber) fail("++/-- need a numeric variable", e->line);
            Value old = *slot;
            const bool inc = e->op.find("++") != std::string::npos;
            if (slot->type == Type::Integer) slot->i += inc ? 1 : -1; else slot->d += inc ? 1.0 : -1.0;
            return e->op[0] == '*' ? old : *slot;
        }
    }
    fail("bad expression", e->line);
}

void Interpreter::exec(const StatementRef& s) {
    switch (s->kind) {
        case Statement::ExprStmt: eval(s->expr); break;
        case Statement::Block:
            frames.back().emplace_back();
            try {
                for (const StatementRef& c : s->body) { exec(c); if (returnFlag) break; }
            } catch (...) { frames.back().pop_back(); throw; }
            frames.back().pop_back();
            break;
        case Statement::If: {
            Value c = eval(s->cond);
            if (!c.convertibleTo(Type::Bool)) fail("Must provide a boolean value.", s->line);
            if (c.convertTo(Type::Bool).b) exec(s->thenStmt);
            else if (s->elseStmt) exec(s->elseStmt);
            break;
        }
        case Statement::For: {
            eval(s->expr);
            while (true) {
                Value c = eval(s->cond);
                if (!c.convertibleTo(Type::Bool)) fail("Must provide a boolean value.", s->line);
                if (!c.convertTo(Type::Bool).b) break;
                exec(s->thenStmt);
                if (returnFlag) break;
                eval(s->post);
            }
            break;
        }
        case Statement::FuncDef: {
            auto fn = std::make_shared<Function>();
            std::vector<ArgInfo> sig;
            for (const auto& ad : s->argDefs) {
                ArgInfo a(ad.first, Type::Any);
                if (ad.second) a.defaultValue = eval(ad.second);
                sig.push_back(a);
            }
            fn->signatures.push_back(sig);
            fn->body = s->thenStmt;
            define(s->name, Value::Ref(Type::Function, fn));
            break;
        }
        case Statement::Return:
            returnValue = s->expr ? eval(s->expr) : Value();
            returnFlag = true;
            break;
    }
}

}  // namespace lang

bool readScene(const std::string& filePath, Scene* scene, RenderingContext* context, std::string* error, bool rgbMode) {
    using namespace lang;
    std::ifstream f(filePath);
    if (!f) { if (error) *error = "cannot open " + filePath; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    Interpreter in;
    in.scene = scene;
    in.context = context;
    in.rgbMode = rgbMode;
    std::string path = filePath;
    for (char& c : path) if (c == '\\') c = '/';
    std::string prefix = path.substr(0, path.find_last_of('/') + 1);
    if (!path.empty() && path[0] == '/') in.sceneDir = prefix;
    else {
        char cwd[4096];
        std::string cur = getcwd(cwd, sizeof(cwd)) ? cwd : ".";
        in.sceneDir = cur + "/" + prefix;
    }
    try {
        registerBuiltins(in);
        std::vector<StatementRef> prog = parseProgram(ss.str(), filePath);
        for (const StatementRef& s : prog) in.exec(s);
    } catch (const RuntimeError& e) {
        if (error) *error = filePath + ":" + std::to_string(e.line) + ": " + e.message;
        return false;
    } catch (const std::exception& e) {
        if (error) *error = filePath + ": " + e.what();
        return false;
    }
    return true;
}

}  // namespace slr
