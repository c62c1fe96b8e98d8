// DEM/utils/Expression.hpp -- scalar expressions of the simulation time `t`, as the strings demo scripts hand to
// SetFamilyPrescribedLinVel / AngVel / Position and AddFamilyPrescribedAcc / AngAcc, e.g. "0.04 * sin(200 * t)",
// "-3.14 / 4", "(t > 1.0) ? 2.0 * sin(5.0 * deme::PI * (t - 1.0)) : 0".
// The reference pastes these strings into its integration kernel and compiles it at run time (jitify;
// src/kernel/DEMIntegrationKernels.cu:100-236, APIPrivate.cpp:1601-1722).  This core is compiled ahead of time, so the
// facade parses them into a small tree once and evaluates them on the host: constants are folded at set-up, and
// time-dependent ones are refreshed before every step through the (stream-ordered) family-table upload.
#pragma once
#include <cctype>
#include <cmath>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace deme {

/// A scalar expression of named variables.  The same grammar serves the prescriptions (one variable, t) and the region
/// strings of inspectors and family-change conditions (owner variables such as X, Y, Z); a leading "return" and a
/// trailing ';' -- the form such strings take in the reference, where they are pasted into a kernel -- are accepted.
class ScalarExpression {
  public:
    /// Parse; throws std::runtime_error with the offending position on a syntax error or an unknown name.
    ScalarExpression(const std::string& text, const std::vector<std::string>& variables) : m_text(text), m_vars(variables) {
        m_pos = 0;
        if (eat("return")) skip();
        m_root = ternary();
        while (eat(";")) {}
        skip();
        if (m_pos != m_text.size()) error("unexpected '" + std::string(1, m_text[m_pos]) + "'");
    }
    /// values[k] is the value of variables[k] of the constructor
    double Eval(const double* values) const { return eval(m_root, values); }
    bool IsConstant() const { return !m_uses_var; }
    const std::string& Text() const { return m_text; }

  private:
    enum Op { NUM, VAR, NEG, NOT, ADD, SUB, MUL, DIV, LT, GT, LE, GE, EQ, NE, AND, OR, SEL, F1, F2 };
    struct Node {
        Op op;
        double val;
        int a, b, c;
        int fn;
    };
    std::string m_text;
    std::vector<std::string> m_vars;
    size_t m_pos = 0;
    std::vector<Node> m_nodes;
    int m_root = -1;
    bool m_uses_var = false;

    [[noreturn]] void error(const std::string& what) const {
        std::string names;
        for (const std::string& v : m_vars) names += v + ", ";
        throw std::runtime_error("Expression \"" + m_text + "\": " + what + " at position " + std::to_string(m_pos) +
                                 ". Supported: numbers, " + names + "PI, + - * /, comparisons, && || !, ?:, and the functions sin cos "
                                 "tan asin acos atan atan2 sinh cosh tanh exp log log10 sqrt abs fabs pow erf erfc floor ceil "
                                 "fmin fmax min max.");
    }
    int add(Op op, int a = -1, int b = -1, int c = -1, double v = 0.0, int fn = 0) {
        m_nodes.push_back(Node{op, v, a, b, c, fn});
        return (int)m_nodes.size() - 1;
    }
    void skip() {
        while (m_pos < m_text.size() && isspace((unsigned char)m_text[m_pos])) m_pos++;
    }
    bool eat(const char* tok) {
        skip();
        const size_t n = strlen(tok);
        if (m_text.compare(m_pos, n, tok) == 0) {
            m_pos += n;
            return true;
        }
        return false;
    }
    int ternary() {
        const int c = logic_or();
        if (eat("?")) {
            const int a = ternary();
            if (!eat(":")) error("':' expected");
            const int b = ternary();
            return add(SEL, c, a, b);
        }
        return c;
    }
    int logic_or() {
        int a = logic_and();
        while (eat("||")) a = add(OR, a, logic_and());
        return a;
    }
    int logic_and() {
        int a = compare();
        while (eat("&&")) a = add(AND, a, compare());
        return a;
    }
    int compare() {
        int a = additive();
        for (;;) {
            if (eat("<=")) a = add(LE, a, additive());
            else if (eat(">=")) a = add(GE, a, additive());
            else if (eat("==")) a = add(EQ, a, additive());
            else if (eat("!=")) a = add(NE, a, additive());
            else if (eat("<")) a = add(LT, a, additive());
            else if (eat(">")) a = add(GT, a, additive());
            else return a;
        }
    }
    int additive() {
        int a = multiplicative();
        for (;;) {
            if (eat("+")) a = add(ADD, a, multiplicative());
            else if (eat("-")) a = add(SUB, a, multiplicative());
            else return a;
        }
    }
    int multiplicative() {
        int a = unary();
        for (;;) {
            if (eat("*")) a = add(MUL, a, unary());
            else if (eat("/")) a = add(DIV, a, unary());
            else return a;
        }
    }
    int unary() {
        if (eat("-")) return add(NEG, unary());
        if (eat("+")) return unary();
        skip();
        if (m_pos < m_text.size() && m_text[m_pos] == '!' && m_text.compare(m_pos, 2, "!=") != 0) {
            m_pos++;
            return add(NOT, unary());
        }
        return primary();
    }
    int primary() {
        skip();
        if (m_pos >= m_text.size()) error("unexpected end");
        const char ch = m_text[m_pos];
        if (isdigit((unsigned char)ch) || ch == '.') {
            size_t used = 0;
            double v = 0.0;
            try {
                v = std::stod(m_text.substr(m_pos), &used);
            } catch (...) {
                error("bad number");
            }
            m_pos += used;
            if (m_pos < m_text.size() && (m_text[m_pos] == 'f' || m_text[m_pos] == 'F')) m_pos++;  // 1.0f
            return add(NUM, -1, -1, -1, v);
        }
        if (ch == '(') {
            m_pos++;
            // a C cast such as (float) or (double) in front of a sub-expression is accepted and ignored
            const size_t save = m_pos;
            for (const char* ty : {"float", "double", "int"}) {
                if (eat(ty) && eat(")")) return unary();
                m_pos = save;
            }
            const int a = ternary();
            if (!eat(")")) error("')' expected");
            return a;
        }
        if (isalpha((unsigned char)ch) || ch == '_') {
            std::string name;
            while (m_pos < m_text.size() && (isalnum((unsigned char)m_text[m_pos]) || m_text[m_pos] == '_' || m_text[m_pos] == ':'))
                name.push_back(m_text[m_pos++]);
            for (const char* prefix : {"deme::", "std::"})
                if (name.rfind(prefix, 0) == 0) name = name.substr(strlen(prefix));
            for (size_t k = 0; k < m_vars.size(); k++)
                if (name == m_vars[k]) {
                    m_uses_var = true;
                    return add(VAR, -1, -1, -1, 0.0, (int)k);
                }
            if (name == "PI" || name == "M_PI") return add(NUM, -1, -1, -1, 3.14159265358979323846);
            static const char* f1[] = {"sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "exp", "log",
                                       "log10", "sqrt", "abs", "fabs", "erf", "erfc", "floor", "ceil", "sinf", "cosf",
                                       "sqrtf", "fabsf", "expf"};
            static const char* f2[] = {"pow", "atan2", "fmin", "fmax", "min", "max", "powf"};
            for (int k = 0; k < (int)(sizeof(f1) / sizeof(f1[0])); k++)
                if (name == f1[k]) {
                    if (!eat("(")) error("'(' expected after " + name);
                    const int a = ternary();
                    if (!eat(")")) error("')' expected");
                    return add(F1, a, -1, -1, 0.0, k);
                }
            for (int k = 0; k < (int)(sizeof(f2) / sizeof(f2[0])); k++)
                if (name == f2[k]) {
                    if (!eat("(")) error("'(' expected after " + name);
                    const int a = ternary();
                    if (!eat(",")) error("',' expected");
                    const int b = ternary();
                    if (!eat(")")) error("')' expected");
                    return add(F2, a, b, -1, 0.0, k);
                }
            error("unknown name '" + name + "'");
        }
        error("unexpected '" + std::string(1, ch) + "'");
    }
    double eval(int i, const double* t) const {
        const Node& n = m_nodes[i];
        switch (n.op) {
            case NUM: return n.val;
            case VAR: return t[n.fn];
            case NEG: return -eval(n.a, t);
            case NOT: return eval(n.a, t) == 0.0 ? 1.0 : 0.0;
            case ADD: return eval(n.a, t) + eval(n.b, t);
            case SUB: return eval(n.a, t) - eval(n.b, t);
            case MUL: return eval(n.a, t) * eval(n.b, t);
            case DIV: return eval(n.a, t) / eval(n.b, t);
            case LT: return eval(n.a, t) < eval(n.b, t);
            case GT: return eval(n.a, t) > eval(n.b, t);
            case LE: return eval(n.a, t) <= eval(n.b, t);
            case GE: return eval(n.a, t) >= eval(n.b, t);
            case EQ: return eval(n.a, t) == eval(n.b, t);
            case NE: return eval(n.a, t) != eval(n.b, t);
            case AND: return (eval(n.a, t) != 0.0) && (eval(n.b, t) != 0.0);
            case OR: return (eval(n.a, t) != 0.0) || (eval(n.b, t) != 0.0);
            case SEL: return eval(n.a, t) != 0.0 ? eval(n.b, t) : eval(n.c, t);
            case F1: {
                const double x = eval(n.a, t);
                switch (n.fn) {
                    case 0: case 19: return std::sin(x);
                    case 1: case 20: return std::cos(x);
                    case 2: return std::tan(x);
                    case 3: return std::asin(x);
                    case 4: return std::acos(x);
                    case 5: return std::atan(x);
                    case 6: return std::sinh(x);
                    case 7: return std::cosh(x);
                    case 8: return std::tanh(x);
                    case 9: case 23: return std::exp(x);
                    case 10: return std::log(x);
                    case 11: return std::log10(x);
                    case 12: case 21: return std::sqrt(x);
                    case 13: case 14: case 22: return std::fabs(x);
                    case 15: return std::erf(x);
                    case 16: return std::erfc(x);
                    case 17: return std::floor(x);
                    default: return std::ceil(x);
                }
            }
            case F2: {
                const double x = eval(n.a, t), y = eval(n.b, t);
                switch (n.fn) {
                    case 0: case 6: return std::pow(x, y);
                    case 1: return std::atan2(x, y);
                    case 2: case 4: return std::fmin(x, y);
                    default: return std::fmax(x, y);
                }
            }
        }
        return 0.0;
    }
};

/// Expressions of the simulation time alone (the prescription strings)
class TimeExpression : public ScalarExpression {
  public:
    explicit TimeExpression(const std::string& text) : ScalarExpression(text, {"t"}) {}
    double Eval(double t) const { return ScalarExpression::Eval(&t); }
};

}  // namespace deme
