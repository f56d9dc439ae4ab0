// fx_front.cpp -- pattern text -> syntax tree, plus literal (all / prefix / suffix) extraction.
//
// This is the host-side "compile the pattern" step of the B200 path.  It has to accept exactly
// the patterns Forgex accepts and give them exactly Forgex's meaning, including the corners
// that fall out of how the reference tokenises (reference: src/ast/syntax_tree_node_m.F90:133-215,
// src/ast/syntax_tree_graph_m.F90:205-1213, src/ast/character_array_m.F90:45-332,
// src/ast/syntax_tree_optimize_m.F90:42-291).  Grammar:
//     regex := term ('|' term)* ;  term := empty | suffixed+ ;  suffixed := primary [* + ? {m,n}]
#include <algorithm>
#include <cstring>

#include "fx_internal.hpp"

namespace fx {

// ---------------------------------------------------------------------------------------------
// byte-string helpers with Fortran CHARACTER semantics (blank padding on compare, trim, adjustl)
// ---------------------------------------------------------------------------------------------
typedef std::string bytes;

static bool blank_eq(const bytes& a, const bytes& b) {
    size_t n = a.size() > b.size() ? a.size() : b.size();
    for (size_t i = 0; i < n; i++) {
        char x = i < a.size() ? a[i] : ' ', y = i < b.size() ? b[i] : ' ';
        if (x != y) return false;
    }
    return true;
}
bool fortran_blank(const std::string& s) { return blank_eq(s, ""); }
static size_t trimmed_len(const bytes& s) {
    size_t n = s.size();
    while (n && s[n - 1] == ' ') n--;
    return n;
}
static bytes rtrim(const bytes& s) { return s.substr(0, trimmed_len(s)); }
static bytes ltrim_rtrim(const bytes& s) {  // trim(adjustl(s))
    size_t k = 0;
    while (k < s.size() && s[k] == ' ') k++;
    return rtrim(s.substr(k));
}
static bool is_char(const bytes& s, char c) { return blank_eq(s, bytes(1, c)); }

// Length in bytes of the character that starts at byte offset `pos` (0-based) under the
// reference's structural rule (utf8_m.f90:44-140, :195-246): lead byte class decides the length,
// the sequence must fit and consist of 10xxxxxx continuation bytes, otherwise it is one byte.
static size_t char_span(const bytes& s, size_t pos) {
    unsigned b = (unsigned char)s[pos];
    size_t n;
    if (b < 0x80) return 1;
    else if ((b >> 5) == 6) n = 2;
    else if ((b >> 4) == 14) n = 3;
    else if ((b >> 3) == 30) n = 4;
    else return 1;
    if (pos + n > s.size()) return 1;
    for (size_t k = 1; k < n; k++)
        if ((((unsigned char)s[pos + k]) >> 6) != 2) return 1;
    return n;
}

// Code point by plain bit concatenation chosen by the lead byte (utf8_m.f90:338-430).  The
// reference reads bytes that may lie beyond the character (blank padding of its 4-byte token
// buffer); `s` is passed with that padding where it matters.
static int code_point(const bytes& s) {
    if (s.size() > 4) return -1;
    unsigned b[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < s.size(); i++) b[i] = (unsigned char)s[i];
    if (s.empty()) return 0;
    if (b[0] < 0x80) return (int)b[0];
    if ((b[0] >> 3) == 30) return (int)((b[0] & 7) << 18 | (b[1] & 63) << 12 | (b[2] & 63) << 6 | (b[3] & 63));
    if ((b[0] >> 4) == 14) return (int)((b[0] & 15) << 12 | (b[1] & 63) << 6 | (b[2] & 63));
    if ((b[0] >> 5) == 6) return (int)((b[0] & 31) << 6 | (b[1] & 63));
    return 0;
}

static bytes encode_utf8(int cp) {  // utf8_m.f90:253-317
    bytes s;
    if (cp < 0x80) s += (char)cp;
    else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    else { s += (char)(0xF0 | ((cp >> 18) & 7)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    return s;
}

// ---------------------------------------------------------------------------------------------
// segment lists
// ---------------------------------------------------------------------------------------------
static const Range R_SENTINEL = {CP_SENTINEL, CP_SENTINEL};
static inline bool is_sentinel(const Range& r) { return r.lo == CP_SENTINEL && r.hi == CP_SENTINEL; }
static inline bool usable(const Range& r) { return r.lo != CP_SENTINEL && r.hi != CP_SENTINEL && r.lo <= r.hi; }

// sort by lower bound, cut at the first sentinel, fuse touching/overlapping ranges
// (segment_m.F90:450-506).  The exchange sort of the reference is not stable; ties only reorder
// ranges that are fused anyway.
static void normalise(std::vector<Range>& v) {
    if (v.empty()) return;
    std::stable_sort(v.begin(), v.end(), [](const Range& a, const Range& b) { return a.lo < b.lo; });
    size_t n = 1;
    while (n < v.size() && !is_sentinel(v[n])) n++;
    v.resize(n);
    std::vector<Range> out;
    out.push_back(v[0]);
    for (size_t i = 1; i < v.size(); i++) {
        if (out.back().hi >= v[i].lo - 1) out.back().hi = std::max(out.back().hi, v[i].hi);
        else out.push_back(v[i]);
    }
    v.swap(out);
}

// Complement as the reference computes it (segment_m.F90:199-253): gaps between the fused
// ranges; the FIRST gap starts at U+0020, later gaps start right after the previous range, the
// last one ends at U+10FFFF.  The reference sizes the result with a different origin (1 instead
// of 32), so when the first listed range starts at or below U+0020 but above U+0000 the array
// keeps one default-initialised element at the end; it is reproduced because the element count
// of a class decides whether the class is a "literal" (syntax_tree_optimize_m.F90:185-198).
static void complement(std::vector<Range>& v) {
    normalise(v);
    size_t slots = 0;
    int from = 1;
    for (auto& r : v) { if (from < r.lo) slots++; from = r.hi + 1; }
    if (from <= CP_MAX) slots++;
    std::vector<Range> out;
    from = 0x20;
    for (auto& r : v) {
        if (from < r.lo) out.push_back({from, r.lo - 1});
        from = r.hi + 1;
    }
    if (from <= CP_MAX) out.push_back({from, CP_MAX});
    while (out.size() < slots) out.push_back(R_SENTINEL);
    v.swap(out);
}

static std::vector<Range> set_word() { return {{97, 122}, {65, 90}, {48, 57}, {95, 95}}; }
static std::vector<Range> set_space() { return {{32, 32}, {9, 9}, {13, 13}, {10, 10}, {12, 12}, {0x3000, 0x3000}}; }
static std::vector<Range> set_digit() { return {{48, 57}}; }
static bool hex_digit_cp(int c) { return (c >= 48 && c <= 57) || (c >= 65 && c <= 70) || (c >= 97 && c <= 102); }

// `\xHH` / `\x{H..}` payload -> code point (segment_m.F90:349-404): at least two digits, hex only,
// must fit a 32-bit integer, must not exceed U+10FFFF.
static int parse_hex(const bytes& digits, int& cp) {
    if (blank_eq(digits, "") || digits.size() < 2) return ERR_HEX_DIGITS;
    unsigned long long v = 0;
    for (char ch : digits) {
        int d;
        if (ch == ' ') continue;
        if (ch >= '0' && ch <= '9') d = ch - '0';
        else if (ch >= 'a' && ch <= 'f') d = ch - 'a' + 10;
        else if (ch >= 'A' && ch <= 'F') d = ch - 'A' + 10;
        else return ERR_INVALID_HEX;
        v = v * 16 + (unsigned)d;
        if (v > 0xFFFFFFFFull) return ERR_INVALID_HEX;
    }
    long long sv = v > 0x7FFFFFFFull ? (long long)v - 0x100000000ll : (long long)v;
    if (sv < 0 || sv > CP_MAX) return ERR_UNICODE_EXCEED;
    cp = (int)sv;
    return OK;
}

// ---------------------------------------------------------------------------------------------
// lexer: one token of look-ahead; `text` is sticky exactly like the reference's token_char,
// which is only rewritten for literal characters, braces and the character after a backslash
// (syntax_tree_node_m.F90:183-207) -- `{..}` and `\x..` parsing observe the stale value.
// ---------------------------------------------------------------------------------------------
enum Tok { T_CHAR, T_BAR, T_LPAR, T_RPAR, T_BSLASH, T_QUES, T_STAR, T_PLUS, T_LBRK, T_RBRK, T_LBRACE, T_RBRACE,
           T_DOT, T_HYPHEN, T_CARET, T_DOLLAR, T_END };

struct Lexer {
    bytes src;
    size_t pos = 0;      // next unread byte
    bool exhausted = false;
    int tok = T_END;
    bytes text = bytes("\0   ", 4);  // 4-byte, blank padded

    static bytes pad4(const bytes& s) { bytes r = s.substr(0, std::min<size_t>(4, s.size())); r.resize(4, ' '); return r; }

    void next(bool in_class = false) {
        if (exhausted || pos >= src.size()) { tok = T_END; text = pad4(""); exhausted = true; return; }
        size_t n = char_span(src, pos);
        bytes c = pad4(src.substr(pos, n));
        size_t after = pos + n;
        bytes key = rtrim(c);
        if (in_class) {
            tok = is_char(key, ']') ? T_RBRK : is_char(key, '-') ? T_HYPHEN : is_char(key, '\\') ? T_BSLASH : T_CHAR;
            text = c;
        } else if (is_char(key, '|')) tok = T_BAR;
        else if (is_char(key, '(')) tok = T_LPAR;
        else if (is_char(key, ')')) tok = T_RPAR;
        else if (is_char(key, '*')) tok = T_STAR;
        else if (is_char(key, '+')) tok = T_PLUS;
        else if (is_char(key, '?')) tok = T_QUES;
        else if (is_char(key, '\\')) {
            tok = T_BSLASH;
            if (after < src.size()) {
                size_t m = char_span(src, after);
                text = pad4(src.substr(after, m));
                after += m;
            } else {  // pattern ends in a backslash: the escaped character is empty
                text = pad4("");
                exhausted = true;
            }
        } else if (is_char(key, '[')) tok = T_LBRK;
        else if (is_char(key, ']')) tok = T_RBRK;
        else if (is_char(key, '{')) { tok = T_LBRACE; text = c; }
        else if (is_char(key, '}')) { tok = T_RBRACE; text = c; }
        else if (is_char(key, '.')) tok = T_DOT;
        else if (is_char(key, '^')) tok = T_CARET;
        else if (is_char(key, '$')) tok = T_DOLLAR;
        else { tok = T_CHAR; text = c; }
        pos = after;
    }
};

// ---------------------------------------------------------------------------------------------
// parser
// ---------------------------------------------------------------------------------------------
struct Parser {
    Syntax& syn;
    Lexer lx;
    int depth = 0;  // '(' seen minus terms that ended on ')' (syntax_tree_graph_m.F90:274-276, :349-351)
    static const int NODE_LIMIT = 2048;

    explicit Parser(Syntax& s) : syn(s) {}

    struct Abort { int code; };

    bool ok() const { return syn.status == OK; }
    int fail(int code) { syn.status = code; return -1; }

    int add(Node n) {
        if ((int)syn.nodes.size() >= NODE_LIMIT) throw Abort{ERR_TREE_NODE_LIMIT};
        syn.nodes.push_back(n);
        return (int)syn.nodes.size() - 1;
    }
    int leaf(int lo, int hi) { Node n; n.op = N_CHAR; n.set.push_back({lo, hi}); return add(n); }
    int leaf_set(const std::vector<Range>& s) { Node n; n.op = N_CHAR; n.set = s; return add(n); }
    int inner(int op, int l, int r) { Node n; n.op = op; n.left = l; n.right = r; return add(n); }

    int parse_regex() {
        int left = parse_term();
        if (!ok()) return -1;
        while (lx.tok == T_BAR) {
            lx.next();
            int right = parse_term();
            if (!ok()) return -1;
            left = inner(N_UNION, left, right);
        }
        return left;
    }

    int parse_term() {
        int left;
        if (lx.tok == T_BAR || lx.tok == T_RPAR || lx.tok == T_END) {
            left = inner(N_EMPTY, -1, -1);
        } else {
            left = parse_suffixed();
            if (!ok()) return -1;
            while (lx.tok != T_BAR && lx.tok != T_RPAR && lx.tok != T_END) {
                int right = parse_suffixed();
                if (!ok()) return -1;
                left = inner(N_CONCAT, left, right);
            }
        }
        if (lx.tok == T_RPAR) depth--;
        return left;
    }

    int parse_suffixed() {
        int atom = parse_primary();
        if (!ok()) return -1;
        switch (lx.tok) {
            case T_STAR: atom = inner(N_CLOSURE, atom, -1); lx.next(); break;
            case T_PLUS: { int star = inner(N_CLOSURE, atom, -1); atom = inner(N_CONCAT, atom, star); lx.next(); break; }
            case T_QUES: { int e = inner(N_EMPTY, atom, -1); atom = inner(N_UNION, atom, e); lx.next(); break; }
            case T_LBRACE:
                atom = parse_counts(atom);
                if (!ok()) return -1;
                lx.next();
                break;
            default: break;
        }
        return atom;
    }

    int parse_primary() {
        int r = -1;
        switch (lx.tok) {
            case T_CHAR:
            case T_RBRACE: {  // a bare '}' is an ordinary character (syntax_tree_graph_m.F90:410-415)
                int cp = code_point(lx.text);
                r = leaf(cp, cp);
                lx.next();
                return r;
            }
            case T_LPAR:
                depth++;
                lx.next();
                r = parse_regex();
                if (!ok()) return -1;
                if (lx.tok != T_RPAR) return fail(ERR_PAREN_MISSING);
                lx.next();
                return r;
            case T_LBRK:
                r = parse_class();
                if (!ok()) return -1;
                if (lx.tok != T_RBRK) return fail(ERR_BRACKET_MISSING);
                lx.next();
                return r;
            case T_BSLASH:
                r = parse_escape();
                if (!ok()) return -1;
                lx.next();
                return r;
            case T_DOT: r = leaf(0x20, CP_MAX); lx.next(); return r;
            case T_CARET:
            case T_DOLLAR: r = line_anchor(); lx.next(); return r;
            case T_RBRK: return fail(ERR_BRACKET_UNEXPECTED);
            case T_RPAR: return fail(ERR_PAREN_UNEXPECTED);
            case T_LBRACE: return fail(ERR_INVALID_TIMES);
            case T_STAR: return fail(ERR_STAR_INCOMPLETE);
            case T_PLUS: return fail(ERR_PLUS_INCOMPLETE);
            case T_QUES: return fail(ERR_QUESTION_INCOMPLETE);
            default: return fail(ERR_SHOULD_NOT_HAPPEN);
        }
    }

    // `\n` outside a class: LF | CR LF, sharing the LF leaf (syntax_tree_graph_m.F90:559-577)
    int newline() {
        int cr = leaf(13, 13), lf = leaf(10, 10);
        int crlf = inner(N_CONCAT, cr, lf);
        return inner(N_UNION, lf, crlf);
    }
    // `^` and `$` are the same consuming atom: (LF | CR LF) | U+0000 (syntax_tree_graph_m.F90:581-605)
    int line_anchor() {
        int nl = newline();
        int nul = leaf(0, 0);
        return inner(N_UNION, nl, nul);
    }

    int parse_escape() {  // syntax_tree_graph_m.F90:611-725
        bytes k = rtrim(lx.text);
        std::vector<Range> set;
        if (is_char(k, 't')) return leaf(9, 9);
        if (is_char(k, 'n')) return newline();
        if (is_char(k, 'r')) return leaf(13, 13);
        if (is_char(k, 'd')) return leaf(48, 57);
        if (is_char(k, 'D')) { set = set_digit(); complement(set); return leaf_set(set); }
        if (is_char(k, 'w')) return leaf_set(set_word());
        if (is_char(k, 'W')) { set = set_word(); complement(set); return leaf_set(set); }
        if (is_char(k, 's')) return leaf_set(set_space());
        if (is_char(k, 'S')) { set = set_space(); complement(set); return leaf_set(set); }
        if (is_char(k, 'x')) {
            int cp = 0;
            int rc = hex_escape(cp);
            if (rc != OK) return fail(rc);
            return leaf(cp, cp);
        }
        if (blank_eq(k, "")) return fail(ERR_ESCAPE_MISSING);
        static const char punct[] = "[]{}()$\\|.?^*+-";
        for (const char* p = punct; *p; p++)
            if (is_char(k, *p)) { int cp = code_point(lx.text); return leaf(cp, cp); }
        return fail(ERR_ESCAPE_INVALID);
    }

    // after `\x`: two hex digits, or `{` digits `}` (syntax_tree_graph_m.F90:728-777).  The digits
    // are taken from the sticky token text, one byte per token.
    int hex_escape(int& cp) {
        lx.next();
        bool braced = lx.tok == T_LBRACE;
        if (braced) lx.next();
        bytes hex = lx.text.substr(0, 1);
        int count = 2;
        while (true) {
            if (!braced && count >= 3) break;
            lx.next();
            if (braced && lx.tok != T_RBRACE && lx.tok != T_CHAR) return ERR_BRACE_MISSING;
            if (lx.tok == T_RBRACE) break;
            hex += lx.text.substr(0, 1);
            count++;
        }
        return parse_hex(rtrim(hex), cp);
    }

    // ---- {m,n} -----------------------------------------------------------------------------
    static bool whole_field_integer(const bytes& s) {  // utility_m.f90:145-169 ('(1i19)' read)
        if (s.find(',') != bytes::npos || s.find(' ') != bytes::npos) return false;
        bytes f = s.substr(0, 19);
        size_t p = 0;
        if (p < f.size() && (f[p] == '+' || f[p] == '-')) p++;
        if (p >= f.size()) return f.empty();
        unsigned long long v = 0;
        for (; p < f.size(); p++) {
            if (f[p] < '0' || f[p] > '9') return false;
            v = v * 10 + (unsigned)(f[p] - '0');
            if (v > 4000000000000000000ull) return false;
        }
        return true;
    }
    // list-directed read of one default integer: <0 nothing to read, 0 ok (value may be left
    // untouched by a null item), >0 malformed (syntax_tree_graph_m.F90:863, :873)
    static int list_read_int(const bytes& s, int& v) {
        size_t p = 0;
        while (p < s.size() && s[p] == ' ') p++;
        if (p >= s.size()) return -1;
        if (s[p] == '/' || s[p] == ',') return 0;
        bool neg = false;
        if (s[p] == '+' || s[p] == '-') { neg = s[p] == '-'; p++; }
        if (p >= s.size() || s[p] < '0' || s[p] > '9') return 1;
        long long acc = 0;
        while (p < s.size() && s[p] >= '0' && s[p] <= '9') {
            acc = acc * 10 + (s[p] - '0');
            if (acc > 0x80000000ll) return 1;
            p++;
        }
        if (p < s.size() && s[p] != ' ' && s[p] != ',' && s[p] != '/') return 1;
        acc = neg ? -acc : acc;
        if (acc > 0x7FFFFFFFll || acc < -0x80000000ll) return 1;
        v = (int)acc;
        return 0;
    }

    int parse_counts(int operand) {  // syntax_tree_graph_m.F90:782-906
        bytes spec;
        lx.next();
        while (lx.tok != T_RBRACE) {
            spec += rtrim(lx.text);
            lx.next();
            if (lx.tok == T_END) return fail(ERR_BRACE_MISSING);
        }
        if (spec.empty()) return fail(ERR_INVALID_TIMES);
        if (spec.size() == 1 && spec[0] == ',') return fail(ERR_INVALID_TIMES);
        if (spec[0] == ',') spec = "0" + spec;
        if (whole_field_integer(spec)) spec = rtrim(spec) + "," + rtrim(spec);
        size_t commas = (size_t)std::count(spec.begin(), spec.end(), ',');
        if (commas > 1) return fail(ERR_INVALID_TIMES);
        size_t comma = spec.find(',');
        bytes lo_s = comma == bytes::npos ? bytes() : spec.substr(0, comma);
        bytes hi_s;
        size_t tl = trimmed_len(spec);
        size_t after = comma == bytes::npos ? 0 : comma + 1;
        if (after < tl) hi_s = spec.substr(after, tl - after);
        int lo = -9999, hi = -9999;
        int rc = list_read_int(lo_s, lo);
        if (rc > 0 || lo < 0) return fail(ERR_INVALID_TIMES);
        bool unbounded = blank_eq(rtrim(hi_s), "");
        if (!unbounded) {
            rc = list_read_int(hi_s, hi);
            if (rc > 0 || hi < 0) return fail(ERR_INVALID_TIMES);
        }
        int rmax = unbounded ? REPEAT_INF : hi;
        if (!(lo == 0 && rmax == 0)) {
            if (rmax != REPEAT_INF && lo > rmax) return fail(ERR_INVALID_TIMES);
        }
        Node n;
        n.op = N_REPEAT; n.left = operand; n.rmin = lo; n.rmax = rmax;
        return add(n);
    }

    // ---- [ ... ] -----------------------------------------------------------------------------
    struct Item {  // one element of a class after flag folding (character_array_m.F90:16-28)
        bytes ch;
        bool escaped = false, ranged = false, subtract = false;
        int width = 0;
    };

    int parse_class() {  // syntax_tree_graph_m.F90:448-556
        lx.next(true);
        bytes body;
        bool bs = false;
        while (lx.tok != T_RBRK) {
            if (lx.tok == T_END) return -1;  // caller reports the missing bracket
            body += lx.text.substr(0, char_span(lx.text, 0));
            bs = (lx.tok == T_BSLASH && !bs);
            lx.next(true);
            if (lx.tok == T_RBRK && bs) {  // `\]` stays inside the class
                body += lx.text.substr(0, char_span(lx.text, 0));
                lx.next(true);
            }
        }
        if (body.empty()) return fail(ERR_EMPTY_CLASS);
        bool negate = body[0] == '^';
        if (negate) body = body.substr(1);
        if (body.empty()) return fail(ERR_EMPTY_CLASS);
        std::vector<Range> set;
        int rc = class_members(body, set);
        if (rc != OK) return fail(rc);
        if (set.empty()) return fail(ERR_SHOULD_NOT_HAPPEN);
        if (negate) complement(set);
        return leaf_set(set);
    }

    static std::vector<Range> escape_members(const bytes& ch) {  // syntax_tree_graph_m.F90:1123-1213
        bytes k = rtrim(ch);
        std::vector<Range> v;
        if (is_char(k, 't')) v = {{9, 9}};
        else if (is_char(k, 'n')) v = {{10, 10}, {13, 13}};
        else if (is_char(k, 'r')) v = {{13, 13}};
        else if (is_char(k, 'd')) v = set_digit();
        else if (is_char(k, 'D')) { v = set_digit(); complement(v); }
        else if (is_char(k, 'w')) v = set_word();
        else if (is_char(k, 'W')) { v = set_word(); complement(v); }
        else if (is_char(k, 's')) v = set_space();
        else if (is_char(k, 'S')) { v = set_space(); complement(v); }
        else if (is_char(k, 'x')) v = {{-1, -1}};
        else if (is_char(k, '\\')) v = {{'\\', '\\'}};
        else if (is_char(k, '{')) v = {{'{', '{'}};
        else if (is_char(k, '}')) v = {{'}', '}'}};
        else if (is_char(k, '[')) v = {{'[', '['}};
        else if (is_char(k, ']')) v = {{']', ']'}};
        else v = {{-2, -2}};  // not escapable inside a class
        return v;
    }
    static int total_width(const std::vector<Range>& v) {
        int w = 0;
        for (auto& r : v) w += usable(r) ? r.hi - r.lo + 1 : -1;
        return w;
    }
    static int escape_width(const bytes& ch) {  // character_array_m.F90:145-222
        if (blank_eq(ch, "t") || blank_eq(ch, "r")) return 1;
        if (blank_eq(ch, "n")) return 2;
        if (blank_eq(ch, "d")) return 10;
        if (blank_eq(ch, "D")) { auto v = set_digit(); complement(v); return total_width(v); }
        if (blank_eq(ch, "w")) return total_width(set_word());
        if (blank_eq(ch, "W")) { auto v = set_word(); complement(v); return total_width(v); }
        if (blank_eq(ch, "s")) return 6;
        if (blank_eq(ch, "S")) { auto v = set_space(); complement(v); return total_width(v); }
        if (blank_eq(ch, "x") || blank_eq(ch, "\\") || blank_eq(ch, "{") || blank_eq(ch, "}") || blank_eq(ch, "[") ||
            blank_eq(ch, "]"))
            return 1;
        return -1;
    }

    // Body of a class -> segment list (syntax_tree_graph_m.F90:910-1118 with the three passes of
    // character_array_m.F90).  Returns a status code.
    static int class_members(const bytes& body, std::vector<Range>& out) {
        int pending = OK;
        if (body.size() >= 2 && body[0] == '-' && body[1] == '-') pending = ERR_MISPLACED_SUBTRACTION;  // noted, not returned yet (:940-945)
        // pass 1: split into characters
        std::vector<bytes> chars;
        for (size_t p = 0; p < body.size();) { size_t n = char_span(body, p); chars.push_back(body.substr(p, n)); p += n; }
        if (chars.empty()) return ERR_EMPTY_CLASS;
        // pass 2: fold backslashes and hyphens into flags (character_array_m.F90:75-140)
        size_t n = chars.size();
        std::vector<Item> items(n);
        size_t k = 0;
        bool zone = false;
        for (size_t i = 0; i < n; i++) {
            if (i > 0 && i + 1 < n) {
                bool twin = is_char(chars[i], '-') && is_char(chars[i + 1], '-');
                if (twin && !zone) {
                    for (size_t q = k; q < n; q++) items[q].subtract = true;
                    zone = true;
                    continue;
                }
                if (twin && zone) return ERR_MISPLACED_SUBTRACTION;
                if (is_char(chars[i - 1], '-') && is_char(chars[i], '-')) continue;
            }
            if (is_char(chars[i], '\\') && !items[k].escaped) items[k].escaped = true;
            else if (is_char(chars[i], '-') && i != 0) {
                if (k > 0) items[k - 1].ranged = true;  // with k == 0 the reference stores out of bounds; nothing is flagged
            } else { items[k].ch = chars[i]; k++; }
        }
        items.resize(k);
        if (pending == ERR_MISPLACED_SUBTRACTION) return pending;
        // pass 3: `\xHH` and `\x{...}` become one element holding the digits (character_array_m.F90:225-332)
        {
            std::vector<Item> folded(items.size());
            size_t siz = items.size(), j = 0, w = 0;
            bytes longhex;
            bool done = false;
            while (j < siz && !done) {
                if (blank_eq(items[j].ch, "x") && items[j].escaped) {
                    folded[w].ch = "x"; folded[w].escaped = true;
                    j++;
                    if (j >= siz) break;
                    w++;
                    if (j + 1 < siz) {
                        if (hex_digit_cp(code_point(items[j].ch)) && hex_digit_cp(code_point(items[j + 1].ch))) {
                            bytes two = rtrim(items[j].ch) + rtrim(items[j + 1].ch);
                            two.resize(2, ' ');
                            folded[w].ch = ltrim_rtrim(two);
                            folded[w].ranged = items[j + 1].ranged;
                            j += 2;
                            if (j >= siz) { done = true; break; }
                            w++;
                            continue;
                        } else if (blank_eq(items[j].ch, "{")) {
                            size_t i = j + 1;
                            while (true) {
                                if (i >= siz) return ERR_BRACE_MISSING;
                                bool close = blank_eq(items[i].ch, "}");
                                if (!close && !hex_digit_cp(code_point(items[i].ch))) return ERR_INVALID_HEX;
                                if (close) break;
                                longhex = ltrim_rtrim(longhex) + items[i].ch;
                                i++;
                            }
                            folded[w].ch = ltrim_rtrim(longhex);
                            folded[w].ranged = items[i].ranged;
                            j = i + 1;
                            if (j >= siz) { done = true; break; }
                            w++;
                            longhex.clear();
                            continue;
                        } else return ERR_INVALID_HEX;
                    } else return ERR_HEX_DIGITS;
                } else if (blank_eq(items[j].ch, "p")) {
                    return ERR_UNICODE_PROPERTY;  // any member `p`, escaped or not (character_array_m.F90:313-315)
                }
                folded[w] = items[j];
                j++;
                if (j >= siz) break;
                w++;
            }
            folded.resize(w + 1);
            items.swap(folded);
        }
        for (auto& it : items) it.width = it.escaped ? escape_width(it.ch) : 1;
        // capacity pass (:976-1024): the result array is sized here; a trailing hyphen after the
        // last of several elements leaves the array one slot short, so that class is rejected.
        long capacity = 0;
        for (size_t i = 0; i < items.size(); i++) {
            if (items[i].ranged && items[i].width != 1) return ERR_RANGE_WITH_ESCAPE;
            if (i > 0 && items[i - 1].ranged && items[i].width != 1) return ERR_RANGE_WITH_ESCAPE;
            if (items[i].subtract) return ERR_CLASS_SUBTRACTION;
            if (i > 0 && i + 1 == items.size() && items[i].ranged) {
                items[i].ranged = false;
                Item h; h.ch = "-"; h.subtract = items[i].subtract; h.width = 1;
                items.push_back(h);
                capacity += 1;
                break;
            }
            capacity += items[i].width;
        }
        if (capacity < 1) return ERR_SHOULD_NOT_HAPPEN;
        // member pass (:1030-1104).  A member is stored only while fewer than capacity-1 are held
        // (segment_m.F90:335).
        std::vector<Range> list;
        auto store = [&](const Range& r) -> bool {
            if (usable(r) && (long)list.size() <= capacity - 1 && (long)list.size() < capacity) { list.push_back(r); return true; }
            return false;
        };
        Range prev = R_SENTINEL, cur = R_SENTINEL;
        bool prev_ranged = false;
        for (size_t i = 0; i < items.size(); i++) {
            bytes ch = items[i].ch;
            bool esc = items[i].escaped;
            bool cur_ranged = items[i].ranged;
            if (i > 0) prev_ranged = items[i - 1].ranged;
            if (esc && blank_eq(ch, "x")) {
                i++;
                if (i >= items.size()) return ERR_SHOULD_NOT_HAPPEN;
                ch = items[i].ch;
                esc = items[i].escaped;
                int cp = 0;
                int rc = parse_hex(ch, cp);
                if (rc != OK) return rc;
                cur = {cp, cp};
            } else if (esc && blank_eq(ch, "p")) {
                return ERR_UNICODE_PROPERTY;
            } else {
                int cp = code_point(ch);
                cur = {cp, cp};
            }
            if (esc) {
                std::vector<Range> m = escape_members(ch);
                if (m[0].lo == -2) return ERR_ESCAPE_INVALID;
                if (m.size() > 1) {
                    for (auto& r : m) store(r);
                    prev = R_SENTINEL;
                    continue;
                }
                cur = m[0];
            }
            if (prev_ranged) {
                Range joined = {prev.lo, cur.hi};
                cur = usable(joined) ? joined : R_SENTINEL;
            }
            if (!cur_ranged) {
                if (!store(cur)) return ERR_INVALID_RANGE;
            }
            prev = cur;
        }
        if (list.empty()) return ERR_SHOULD_NOT_HAPPEN;
        out = list;
        return OK;
    }
};

void parse_pattern(const std::string& pattern, Syntax& out) {
    out = Syntax();
    Parser p(out);
    p.lx.src = pattern;
    try {
        p.lx.next();
        int root = p.parse_regex();
        if (out.status == OK) {
            if (p.depth > 0) out.status = ERR_PAREN_MISSING;
            else if (p.depth < 0) out.status = ERR_PAREN_UNEXPECTED;
        }
        out.root = out.status == OK ? root : -1;
    } catch (const Parser::Abort& a) {
        out.status = a.code;
        out.root = -1;
    }
}

// Entry-point preprocessing of the pattern text (forgex.F90:95, :182-190, :260).
std::string prepare_pattern(const std::string& pattern, bool match_mode) {
    if (!match_mode) return rtrim(pattern);
    bytes buf = pattern;
    size_t lead = 0;
    while (lead < pattern.size() && pattern[lead] == ' ') lead++;
    bool caret = lead < pattern.size() && pattern[lead] == '^';
    if (caret) buf = pattern.substr(1);  // byte 1 is dropped, whatever it is
    size_t tl = trimmed_len(pattern);
    bool dollar = tl > 0 && pattern[tl - 1] == '$';
    if (dollar) {
        size_t keep = tl - 1;  // an index into the ORIGINAL pattern
        if (keep < buf.size()) buf = buf.substr(0, keep);
    }
    return buf;
}

// ---------------------------------------------------------------------------------------------
// literal extraction (syntax_tree_optimize_m.F90:71-291)
// ---------------------------------------------------------------------------------------------
namespace {
struct Lit {
    bytes all, pre, suf;
    bool closure = false, cls = false;
};

bytes longer_of(const bytes& a, const bytes& b) { return trimmed_len(a) > trimmed_len(b) ? ltrim_rtrim(a) : ltrim_rtrim(b); }

bytes common_prefix(const bytes& a, const bytes& b) {  // character-wise, by a's character widths
    bytes r;
    size_t i = 0;
    while (i < a.size() && i < b.size()) {
        size_t na = char_span(a, i), nb = char_span(b, i);
        if (!blank_eq(a.substr(i, na), b.substr(i, nb))) break;
        r += a.substr(i, na);
        i += na;
    }
    return r;
}
bytes reversed_chars(const bytes& s) {
    bytes r;
    for (size_t i = 0; i < s.size();) { size_t n = char_span(s, i); r = s.substr(i, n) + r; i += n; }
    return r;
}
bytes common_suffix(const bytes& a, const bytes& b) { return reversed_chars(common_prefix(reversed_chars(a), reversed_chars(b))); }

// `acc` keeps its flags between calls when the caller reuses it, as the reference does.
void walk(const Syntax& syn, int idx, Lit& acc) {
    const Node& n = syn.nodes[(size_t)idx];
    acc.all.clear(); acc.pre.clear(); acc.suf.clear();
    Lit l, r;
    if (n.op == N_UNION || n.op == N_CONCAT) { walk(syn, n.left, l); walk(syn, n.right, r); }
    switch (n.op) {
        case N_UNION:
            acc.pre = common_prefix(l.pre, r.pre);
            acc.suf = common_suffix(l.suf, r.suf);
            acc.closure = true;
            break;
        case N_CONCAT: {
            acc.cls = l.cls || r.cls;
            acc.closure = l.closure || r.closure;
            // 16-way table of the reference (:106-177) folded into its distinct outcomes
            enum { BEST_PRE = 1, CAT_PRE = 2, L_PRE = 3 } pre_rule;
            enum { BEST_SUF = 1, CAT_SUF = 2, R_SUF = 3, NO_SUF = 4 } suf_rule;
            if (!l.cls && !r.cls) {
                if (!l.closure && !r.closure) { pre_rule = BEST_PRE; suf_rule = BEST_SUF; acc.all = l.all + r.all; }
                else if (!l.closure) { pre_rule = CAT_PRE; suf_rule = R_SUF; }
                else if (!r.closure) { pre_rule = L_PRE; suf_rule = CAT_SUF; }
                else { pre_rule = L_PRE; suf_rule = R_SUF; }
            } else if (!l.cls) {
                pre_rule = l.closure ? L_PRE : BEST_PRE; suf_rule = R_SUF;
            } else if (!r.cls) {
                pre_rule = L_PRE; suf_rule = r.closure ? R_SUF : BEST_SUF;
            } else {
                pre_rule = L_PRE; suf_rule = (!l.closure && r.closure) ? NO_SUF : R_SUF;
            }
            acc.pre = pre_rule == BEST_PRE ? longer_of(l.pre, l.all + r.pre) : pre_rule == CAT_PRE ? l.all + r.pre : l.pre;
            if (suf_rule == BEST_SUF) acc.suf = longer_of(r.suf, l.suf + r.all);
            else if (suf_rule == CAT_SUF) acc.suf = l.suf + r.all;
            else if (suf_rule == R_SUF) acc.suf = r.suf;
            break;
        }
        case N_CLOSURE: acc.closure = true; break;
        case N_CHAR:
            if (n.set.size() == 1 && usable(n.set[0]) && n.set[0].lo == n.set[0].hi) acc.all = acc.pre = acc.suf = encode_utf8(n.set[0].lo);
            else acc.cls = true;
            break;
        case N_REPEAT: {
            walk(syn, n.left, l);
            acc.cls = l.cls;
            for (int i = 0; i < n.rmin; i++) {
                walk(syn, n.left, l);
                acc.all += l.all; acc.pre += l.pre; acc.suf += l.suf;
                acc.cls = acc.cls || l.cls;
                if (l.closure) break;
            }
            acc.closure = (n.rmin != n.rmax) || l.closure;
            break;
        }
        default: acc.closure = true; break;
    }
}
}  // namespace

void extract_literals(const Syntax& syn, Literals& lit) {
    Lit acc;
    walk(syn, syn.root, acc);
    lit.all = acc.all; lit.prefix = acc.pre; lit.suffix = acc.suf;
}

const char* status_message(int code) {  // error_m.F90:41-211
    switch (code) {
        case OK: return "Given pattern is valid.";
        case ERR_GENERIC: return "ERROR: Pattern includes some syntax error.";
        case ERR_PAREN_MISSING: return "ERROR: Closing parenthesis is expected.";
        case ERR_PAREN_UNEXPECTED: return "ERROR: Unexpected closing parenthesis error.";
        case ERR_BRACKET_MISSING: return "ERROR: Closing square bracket is expected.";
        case ERR_BRACKET_UNEXPECTED: return "ERROR: Unexpected closing square bracket error.";
        case ERR_BRACE_MISSING: return "ERROR: Closing right curlybrace is expected.";
        case ERR_BRACE_UNEXPECTED: return "ERROR: Unexpected closing right curlybrace error.";
        case ERR_INVALID_TIMES: return "ERROR: Given quantifier range is invalid.";
        case ERR_ESCAPE_MISSING: return "ERROR: Pattern cannot end with a trailing unescaped backslash.";
        case ERR_ESCAPE_INVALID: return "ERROR: This token has no special meaning.";
        case ERR_EMPTY_CLASS: return "ERROR: Given class has no character.";
        case ERR_RANGE_WITH_ESCAPE: return "ERROR: Cannot create a range with shorthand escape sequence";
        case ERR_MISPLACED_SUBTRACTION: return "ERROR: Subtraction operator is misplaced in the given character class.";
        case ERR_INVALID_RANGE: return "ERROR: Given character range is invalid.";
        case ERR_CLASS_SUBTRACTION: return "ERROR: Character class subtraction hasn't implemented yet.";
        case ERR_STAR_INCOMPLETE: return "ERROR: Not quantifiable; star '*' operator is missing operand.";
        case ERR_PLUS_INCOMPLETE: return "ERROR: Not quantifiable; plus '+' operator is missing operand.";
        case ERR_QUESTION_INCOMPLETE: return "ERROR: Not quantifiable; question '?' operator is missing operand.";
        case ERR_INVALID_HEX: return "ERROR: Invalid characters detected. Ensure all characters are 0-9, A-F/a-f.";
        case ERR_HEX_DIGITS: return "ERROR: At least 2 hexadecimal digits are required (e.g., '0A' instead of 'A').";
        case ERR_UNICODE_EXCEED: return "ERROR: Given hex number exceeds the range of unicode codepoint.";
        case ERR_ALLOCATION: return "ERROR: Allocation is failed.";
        case ERR_TREE_NODE_LIMIT: return "forgex_b200: pattern needs more than 2048 syntax-tree nodes (Forgex aborts here).";
        case ERR_DFA_STATE_CAP: return "forgex_b200: eagerly built automaton exceeds the state cap.";
        case ERR_PREFILTER_UNSUPPORTED: return "forgex_b200: this pattern's literal prefilter is not result-neutral; not supported yet.";
        case ERR_BAD_ARGUMENT: return "forgex_b200: bad argument.";
        case ERR_NO_DEVICE: return "forgex_b200: no CUDA device / CUDA failure.";
        case ERR_WORK_BUDGET: return "forgex_b200: buffer search stopped; Forgex's candidate loop is super-linear on this text.";
        default: return "ERROR: Fatal error is happened.";
    }
}

}  // namespace fx
