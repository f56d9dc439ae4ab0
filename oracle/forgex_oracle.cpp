// forgex_oracle.cpp -- CPU restatement of Forgex's matching path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for the B200 kernels: a deliberately literal C++17
// restatement of the reference's own algorithm (Fortran, /root/reference/src), including its
// data-structure quirks, so that "bit-exact with Forgex" can be checked without a Fortran
// compiler (there is none in this image).  It is pinned against the reference's own
// known-answer tests (tests/golden/reference_*.json, transcribed by
// tools/transcribe_vectors.py from /root/reference/test/**).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (forgex_b200/) never links or calls it.
//
// Matching here is NOT a table DFA: as in the reference, every input character performs an
// NFA subset step + epsilon closure + linear search of the registered DFA states
// (src/automaton_m.F90:199-381).  What is deliberately not restated: the per-character deep
// copies of NFA nodes / segment arrays and the internal formatted reads in ichar_utf8 (they
// cost time but cannot change results), the DFA transition cache that is written but never
// read (src/automaton_m.F90:375-380), and the backward NFA transitions (never read by matching).
//
// Conventions: Fortran strings are std::string byte strings; all indices are 1-based as in the
// source; every function cites the reference lines it follows.

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace fxo {

typedef std::string fstr;

// ---------------------------------------------------------------------------------------------
// Fortran character semantics
// ---------------------------------------------------------------------------------------------
static inline fstr sub(const fstr& s, long a, long b) {  // s(a:b); zero length when b < a
    if (b < a) return fstr();
    if (a < 1) a = 1;
    if (b > (long)s.size()) b = (long)s.size();
    if (b < a) return fstr();
    return s.substr((size_t)(a - 1), (size_t)(b - a + 1));
}
static inline bool f_eq(const fstr& a, const fstr& b) {  // blank-padded comparison (==)
    size_t n = std::max(a.size(), b.size());
    for (size_t i = 0; i < n; i++) {
        unsigned char ca = i < a.size() ? (unsigned char)a[i] : ' ';
        unsigned char cb = i < b.size() ? (unsigned char)b[i] : ' ';
        if (ca != cb) return false;
    }
    return true;
}
static inline long f_len_trim(const fstr& s) {
    long n = (long)s.size();
    while (n > 0 && s[(size_t)n - 1] == ' ') n--;
    return n;
}
static inline fstr f_trim(const fstr& s) { return s.substr(0, (size_t)f_len_trim(s)); }
static inline fstr f_adjustl(const fstr& s) {
    size_t k = 0;
    while (k < s.size() && s[k] == ' ') k++;
    return s.substr(k) + fstr(k, ' ');
}
static inline long f_index(const fstr& s, const fstr& t) {  // index(s, t)
    if (t.size() > s.size()) return 0;
    if (t.empty()) return 1;
    size_t p = s.find(t);
    return p == fstr::npos ? 0 : (long)p + 1;
}
static inline long f_index_back(const fstr& s, const fstr& t) {  // index(s, t, back=.true.)
    if (t.size() > s.size()) return 0;
    if (t.empty()) return (long)s.size() + 1;
    size_t p = s.rfind(t);
    return p == fstr::npos ? 0 : (long)p + 1;
}

// ---------------------------------------------------------------------------------------------
// parameters (src/essential/parameters_m.f90)
// ---------------------------------------------------------------------------------------------
static const int TREE_NODE_HARD_LIMIT = 2048;    // :21-25
static const int INVALID_REPEAT_VAL = -9999;     // :29
static const int INFINITE_ = -9998;              // :30
static const int INVALID_CHAR_INDEX = -9999;     // :31
static const int UTF8_CODE_MAX = 1114111;        // :38
static const int UTF8_CODE_MIN = 32;             // :40
static const int UTF8_CODE_EMPTY = 0;            // :41
static const int UTF8_CODE_INVALID = -1;         // :42
static const int INVALID_INDEX = -9999;          // :79
static const int NFA_NULL_TRANSITION = -1;       // :91
static const int NFA_C_SIZE = 16;                // :106
static const int DFA_STATE_UNIT = 16;            // :122
static const int DFA_STATE_HARD_LIMIT = 1024 * 16 + 1;  // :126-130
static const int DFA_INVALID_INDEX = 0;          // :133
static const int ACCEPTED_EMPTY = -2;            // :154

// status codes (src/essential/error_m.F90:12-38)
enum {
    SYNTAX_VALID = 0, SYNTAX_ERR, SYNTAX_ERR_PARENTHESIS_MISSING, SYNTAX_ERR_PARENTHESIS_UNEXPECTED,
    SYNTAX_ERR_BRACKET_MISSING, SYNTAX_ERR_BRACKET_UNEXPECTED, SYNTAX_ERR_CURLYBRACE_MISSING,
    SYNTAX_ERR_CURLYBRACE_UNEXPECTED, SYNTAX_ERR_INVALID_TIMES, SYNTAX_ERR_ESCAPED_SYMBOL_MISSING,
    SYNTAX_ERR_ESCAPED_SYMBOL_INVALID, SYNTAX_ERR_EMPTY_CHARACTER_CLASS,
    SYNTAX_ERR_RANGE_WITH_ESCAPE_SEQUENCES, SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR,
    SYNTAX_ERR_INVALID_CHARACTER_RANGE, SYNTAX_ERR_CHAR_CLASS_SUBTRANCTION_NOT_IMPLEMENTED,
    SYNTAX_ERR_STAR_INCOMPLETE, SYNTAX_ERR_PLUS_INCOMPLETE, SYNTAX_ERR_QUESTION_INCOMPLETE,
    SYNTAX_ERR_INVALID_HEXADECIMAL, SYNTAX_ERR_HEX_DIGITS_NOT_ENOUGH, SYNTAX_ERR_UNICODE_EXCEED,
    SYNTAX_ERR_UNICODE_PROPERTY_NOT_IMPLEMENTED, SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN, ALLOCATION_ERR
};
// The reference aborts (`error stop`) in these situations; the oracle reports them instead.
static const int ERRSTOP_TREE_LIMIT = -1001;  // src/ast/syntax_tree_graph_m.F90:115-117
static const int ERRSTOP_DFA_LIMIT = -1002;   // src/lazy_dfa/lazy_dfa_graph_m.F90:90-92

struct ErrorStop { int code; };

static const char* error_message(int code) {  // src/essential/error_m.F90:127-211
    switch (code) {
        case SYNTAX_VALID: return "Given pattern is valid.";
        case SYNTAX_ERR: return "ERROR: Pattern includes some syntax error.";
        case SYNTAX_ERR_PARENTHESIS_MISSING: return "ERROR: Closing parenthesis is expected.";
        case SYNTAX_ERR_PARENTHESIS_UNEXPECTED: return "ERROR: Unexpected closing parenthesis error.";
        case SYNTAX_ERR_BRACKET_MISSING: return "ERROR: Closing square bracket is expected.";
        case SYNTAX_ERR_BRACKET_UNEXPECTED: return "ERROR: Unexpected closing square bracket error.";
        case SYNTAX_ERR_CURLYBRACE_MISSING: return "ERROR: Closing right curlybrace is expected.";
        case SYNTAX_ERR_CURLYBRACE_UNEXPECTED: return "ERROR: Unexpected closing right curlybrace error.";
        case SYNTAX_ERR_INVALID_TIMES: return "ERROR: Given quantifier range is invalid.";
        case SYNTAX_ERR_ESCAPED_SYMBOL_MISSING: return "ERROR: Pattern cannot end with a trailing unescaped backslash.";
        case SYNTAX_ERR_ESCAPED_SYMBOL_INVALID: return "ERROR: This token has no special meaning.";
        case SYNTAX_ERR_EMPTY_CHARACTER_CLASS: return "ERROR: Given class has no character.";
        case SYNTAX_ERR_RANGE_WITH_ESCAPE_SEQUENCES: return "ERROR: Cannot create a range with shorthand escape sequence";
        case SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR: return "ERROR: Subtraction operator is misplaced in the given character class.";
        case SYNTAX_ERR_INVALID_CHARACTER_RANGE: return "ERROR: Given character range is invalid.";
        case SYNTAX_ERR_CHAR_CLASS_SUBTRANCTION_NOT_IMPLEMENTED: return "ERROR: Character class subtraction hasn't implemented yet.";
        case SYNTAX_ERR_STAR_INCOMPLETE: return "ERROR: Not quantifiable; star '*' operator is missing operand.";
        case SYNTAX_ERR_PLUS_INCOMPLETE: return "ERROR: Not quantifiable; plus '+' operator is missing operand.";
        case SYNTAX_ERR_QUESTION_INCOMPLETE: return "ERROR: Not quantifiable; question '?' operator is missing operand.";
        case SYNTAX_ERR_INVALID_HEXADECIMAL: return "ERROR: Invalid characters detected. Ensure all characters are 0-9, A-F/a-f.";
        case SYNTAX_ERR_HEX_DIGITS_NOT_ENOUGH: return "ERROR: At least 2 hexadecimal digits are required (e.g., '0A' instead of 'A').";
        case SYNTAX_ERR_UNICODE_EXCEED: return "ERROR: Given hex number exceeds the range of unicode codepoint.";
        case ALLOCATION_ERR: return "ERROR: Allocation is failed.";
        default: return "ERROR: Fatal error is happened.";  // incl. UNICODE_PROPERTY (no case arm, :127-211)
    }
}

// ---------------------------------------------------------------------------------------------
// UTF-8 (src/essential/utf8_m.f90)
// ---------------------------------------------------------------------------------------------
static inline int ub(const fstr& s, long i) { return (unsigned char)s[(size_t)i - 1]; }

// :195-246
static bool is_valid_multiple_byte_character(const fstr& chara) {
    long siz = (long)chara.size();
    if (siz == 0) return false;  // never called with an empty string by the paths restated here
    int byte = ub(chara, 1);
    long expected;
    if ((byte >> 3) == 31) return false;
    else if ((byte >> 3) == 30) expected = 4;
    else if ((byte >> 4) == 14) expected = 3;
    else if ((byte >> 5) == 6) expected = 2;
    else if ((byte >> 7) == 0) expected = 1;
    else return false;
    if (expected != siz) return false;
    for (long i = 2; i <= expected; i++)
        if ((ub(chara, i) >> 6) != 2) return false;
    return true;
}

// :44-140  index of the last byte of the character starting at curr (curr itself if malformed)
static long idxutf8(const fstr& str, long curr) {
    long len = (long)str.size();
    if (curr > len) return INVALID_CHAR_INDEX;
    long tail = curr;
    for (long i = 0; i <= 3; i++) {
        if (curr + i > len) return curr;                      // :81-84
        int byte = ub(str, curr + i);
        if ((byte >> 6) == 2) continue;                       // :95 (also for i == 0)
        if (i == 0) {
            if ((byte >> 3) == 30) { tail = curr + 3; break; }
            if ((byte >> 4) == 14) { tail = curr + 2; break; }
            if ((byte >> 5) == 6) { tail = curr + 1; break; }
            if ((byte >> 7) == 0) { tail = curr; break; }
        } else {
            if ((byte >> 3) == 30 || (byte >> 4) == 14 || (byte >> 5) == 6 || (byte >> 7) == 0) {
                tail = curr + i - 1;
                break;
            }
        }
    }
    if (tail <= len) {
        if (!is_valid_multiple_byte_character(sub(str, curr, tail))) tail = curr;
    } else {
        tail = curr;
    }
    return tail;
}

// :146-163
static long next_idxutf8(const fstr& str, long curr) {
    long e = idxutf8(str, curr);
    return e != INVALID_CHAR_INDEX ? e + 1 : INVALID_CHAR_INDEX;
}

// :168-191
static void next_idxutf8_strict(const fstr& str, long curr, long& next, bool& is_valid) {
    is_valid = false;
    long ib = curr;
    long ie = idxutf8(str, ib);
    if (ie != INVALID_CHAR_INDEX) {
        is_valid = is_valid_multiple_byte_character(sub(str, ib, ie));
        next = ie + 1;
    } else {
        next = curr + 1;
        is_valid = false;
    }
}

// :253-317  (blank filler bytes + trim(adjustl()) in the source == the plain encoder, because
// no produced byte is 0x20 for code > 127)
static fstr char_utf8(int code) {
    fstr s;
    if (code > 127) {
        int b1 = (code >> 18) & 63, b2 = (code >> 12) & 63, b3 = (code >> 6) & 63, b4 = code & 63;
        if (code > 65535) {
            s += (char)(0xF0 | (b1 & 7)); s += (char)(0x80 | b2); s += (char)(0x80 | b3); s += (char)(0x80 | b4);
        } else if (code > 2047) {
            s += (char)(0xE0 | (b2 & 15)); s += (char)(0x80 | b3); s += (char)(0x80 | b4);
        } else {
            s += (char)(0xC0 | (b3 & 31)); s += (char)(0x80 | b4);
        }
    } else {
        s += (char)code;
    }
    return s;
}

// :338-430  plain bit concatenation selected by the lead byte; bytes past len(chara) are
// undefined in the source (read as 0 here -- only reachable with malformed UTF-8 in a pattern)
static int ichar_utf8(const fstr& chara) {
    if (chara.size() > 4) return -1;
    if (chara.empty()) return 0;
    int b[4] = {0, 0, 0, 0};
    for (size_t i = 0; i < chara.size(); i++) b[i] = (unsigned char)chara[i];
    if ((b[0] >> 7) == 0) return b[0];
    if ((b[0] >> 3) == 30) return ((((b[0] & 7) << 6 | (b[1] & 63)) << 6 | (b[2] & 63)) << 6) | (b[3] & 63);
    if ((b[0] >> 4) == 14) return (((b[0] & 15) << 6 | (b[1] & 63)) << 6) | (b[2] & 63);
    if ((b[0] >> 5) == 6) return ((b[0] & 31) << 6) | (b[1] & 63);
    return 0;
}

// :466-481
static long len_utf8(const fstr& str) {
    long i = 1, count = 0;
    while (i <= (long)str.size()) {
        long inext = idxutf8(str, i) + 1;
        count++;
        i = inext;
    }
    return count;
}

// :596-613
static fstr reverse_utf8(const fstr& str) {
    fstr ret;
    long i = 1;
    while (i != INVALID_CHAR_INDEX) {
        long ie = idxutf8(str, i);
        ret = sub(str, i, ie) + ret;
        i = next_idxutf8(str, i);
    }
    return ret;
}

// ---------------------------------------------------------------------------------------------
// segments (src/essential/segment_m.F90)
// ---------------------------------------------------------------------------------------------
struct Seg {
    int min, max;
    Seg() : min(UTF8_CODE_MAX + 2), max(UTF8_CODE_MAX + 2) {}  // :38-39
    Seg(int a, int b) : min(a), max(b) {}
    bool operator==(const Seg& o) const { return min == o.min && max == o.max; }
    bool operator!=(const Seg& o) const { return !(*this == o); }
    bool validate() const {  // :185-193
        Seg init;
        return min != init.min && max != init.max && min <= max;
    }
};
static const Seg SEG_INIT(UTF8_CODE_MAX + 2, UTF8_CODE_MAX + 2);
static const Seg SEG_ERROR(-2, -2);
static const Seg SEG_EPSILON(-1, -1);
static const Seg SEG_EMPTY(UTF8_CODE_EMPTY, UTF8_CODE_EMPTY);
static const Seg SEG_ANY(UTF8_CODE_MIN, UTF8_CODE_MAX);
static const Seg SEG_TAB(9, 9), SEG_LF(10, 10), SEG_FF(12, 12), SEG_CR(13, 13), SEG_SPACE(32, 32);
static const Seg SEG_UNDERSCORE(95, 95), SEG_DIGIT(48, 57), SEG_UPPERCASE(65, 90), SEG_LOWERCASE(97, 122);
static const Seg SEG_ZENKAKU_SPACE(12288, 12288);
static const Seg SEG_UPPER(UTF8_CODE_MAX + 1, UTF8_CODE_MAX + 1);
static const Seg SEG_WHOLE(0, UTF8_CODE_MAX);

static inline bool in_seg(int a, const Seg& s) { return s.min <= a && a <= s.max; }       // :100-107
static inline bool seg_in_seg(const Seg& a, const Seg& b) { return b.min <= a.min && a.max <= b.max; }  // :135-141
static bool in_hex(int a) { return in_seg(a, SEG_DIGIT) || in_seg(a, Seg(65, 70)) || in_seg(a, Seg(97, 102)); }  // :63

static void sort_segment_by_min(std::vector<Seg>& s) {  // :450-469 (exchange sort, kept literal)
    long n = (long)s.size();
    for (long i = 0; i < n - 1; i++)
        for (long j = i + 1; j < n; j++)
            if (s[i].min > s[j].min) std::swap(s[i], s[j]);
}

static void merge_segments(std::vector<Seg>& s) {  // :472-506
    long n = (long)s.size();
    if (n == 0) return;
    long m = 1;
    for (long i = 2; i <= n; i++) {
        if (s[i - 1] == SEG_INIT) break;
        m++;
    }
    n = m;
    if (n <= 1) { s.resize((size_t)n); return; }
    long j = 1;
    for (long i = 2; i <= n; i++) {
        if (s[j - 1].max >= s[i - 1].min - 1) {
            s[j - 1].max = std::max(s[j - 1].max, s[i - 1].max);
        } else {
            j++;
            s[j - 1] = s[i - 1];
        }
    }
    if (j <= n) s.resize((size_t)j);
}

static void invert_segment_list(std::vector<Seg>& list) {  // :199-253
    sort_segment_by_min(list);
    merge_segments(list);
    long n = (long)list.size();
    long count = 0;
    int current_min = UTF8_CODE_EMPTY + 1;     // sizing pass origin (:215)
    for (long i = 0; i < n; i++) {
        if (current_min < list[i].min) count++;
        current_min = list[i].max + 1;
    }
    if (current_min <= UTF8_CODE_MAX) count++;
    std::vector<Seg> nl((size_t)count);        // default elements are SEG_INIT
    count = 1;
    current_min = UTF8_CODE_MIN;               // fill pass origin (:234)
    for (long i = 0; i < n; i++) {
        if (current_min < list[i].min) {
            nl[count - 1].min = current_min;
            nl[count - 1].max = list[i].min - 1;
            count++;
        }
        current_min = list[i].max + 1;
    }
    if (current_min <= UTF8_CODE_MAX) {
        nl[count - 1].min = current_min;
        nl[count - 1].max = UTF8_CODE_MAX;
    }
    list = nl;
}

// :296-322
static Seg symbol_to_segment(const fstr& symbol) {
    if (f_eq(symbol, fstr(1, '\0'))) return SEG_EMPTY;
    if (f_eq(symbol, " ")) return SEG_SPACE;
    long e = idxutf8(symbol, 1);
    int code = ichar_utf8(sub(symbol, 1, e));
    return Seg(code, code);
}

// :327-344
static const int SEGMENT_REGISTERED = 0, SEGMENT_REJECTED = 1;
static void register_segment_to_list(std::vector<Seg>& list, const Seg& seg, long& k, int& ierr) {
    if (seg.validate() && k <= (long)list.size() - 1) {
        k++;
        list[(size_t)k - 1] = seg;
        ierr = SEGMENT_REGISTERED;
    } else {
        ierr = SEGMENT_REJECTED;
    }
}

// :349-404.  `read(str, '(zN)')`: blanks are skipped, anything that is not a hex digit is an
// I/O error; a value that does not fit the default integer is an I/O error too.
static void hex2seg(const fstr& str, Seg& seg, int& ierr) {
    seg = Seg(UTF8_CODE_INVALID, UTF8_CODE_INVALID);
    if (f_eq(str, "") || str.size() < 2) { ierr = SYNTAX_ERR_HEX_DIGITS_NOT_ENOUGH; return; }
    unsigned long long code = 0;
    bool ok = true;
    for (size_t i = 0; i < str.size() && ok; i++) {
        unsigned char c = (unsigned char)str[i];
        int d;
        if (c == ' ') continue;
        if (c >= '0' && c <= '9') d = c - '0';
        else if (c >= 'a' && c <= 'f') d = c - 'a' + 10;
        else if (c >= 'A' && c <= 'F') d = c - 'A' + 10;
        else { ok = false; break; }
        code = code * 16 + (unsigned)d;
        if (code > 0xFFFFFFFFull) ok = false;
    }
    if (!ok) { ierr = SYNTAX_ERR_INVALID_HEXADECIMAL; return; }
    long long scode = (code > 0x7FFFFFFFull) ? (long long)code - 0x100000000ll : (long long)code;
    if (!(scode >= 0 && scode <= UTF8_CODE_MAX)) { ierr = SYNTAX_ERR_UNICODE_EXCEED; return; }
    seg = Seg((int)scode, (int)scode);
    ierr = SYNTAX_VALID;
}

static int width_of_segment(const Seg& s) { return s.validate() ? s.max - s.min + 1 : -1; }  // :410-421
static int total_width_of_segment(const std::vector<Seg>& l) {
    int r = 0;
    for (auto& s : l) r += width_of_segment(s);
    return r;
}
static Seg join_two_segments(const Seg& a, const Seg& b) {  // :436-447
    Seg r(a.min, b.max);
    if (!r.validate()) r = SEG_INIT;
    return r;
}

// ---------------------------------------------------------------------------------------------
// priority queue (src/essential/priority_queue_m.f90) and disjoin (segment_disjoin_m.F90)
// ---------------------------------------------------------------------------------------------
struct PQueue {
    std::vector<Seg> heap;  // 1-based via heap[i-1]
    long number = 0;
    void enqueue(const Seg& seg) {  // :41-79
        number++;
        if ((long)heap.size() < number) heap.resize((size_t)number);
        heap[(size_t)number - 1] = seg;
        long n = number;
        while (n > 1) {
            long i = n / 2;
            Seg &a = heap[(size_t)n - 1], &b = heap[(size_t)i - 1];
            if (a.min < b.min || (a.min == b.min && a.max < b.max)) std::swap(a, b);
            n = i;
        }
    }
    void dequeue(Seg& res) {  // :82-114
        long n = number;
        res = heap[0];
        heap[0] = heap[(size_t)n - 1];
        number--;
        long i = 1;
        while (2 * i < n) {
            long j = 2 * i;
            if (j + 1 < n && heap[(size_t)j].min < heap[(size_t)j - 1].min) j++;
            if (heap[(size_t)j - 1].min < heap[(size_t)i - 1].min) std::swap(heap[(size_t)j - 1], heap[(size_t)i - 1]);
            i = j;
        }
    }
};

// segment_disjoin_m.F90:36-182
static void disjoin(std::vector<Seg>& list) {
    long siz = (long)list.size();
    if (siz <= 0) return;
    std::vector<Seg> old_list = list;
    PQueue pq;
    std::vector<Seg> buff((size_t)siz);
    for (long j = 0; j < siz; j++) pq.enqueue(old_list[(size_t)j]);
    for (long j = 0; j < siz; j++) pq.dequeue(buff[(size_t)j]);
    // index_list_from_segment_list :259-304 (sorted unique of min-1,min,min+1,max-1,max,max+1)
    std::vector<int> idx;
    for (auto& s : old_list) {
        idx.push_back(s.min - 1); idx.push_back(s.min); idx.push_back(s.min + 1);
        idx.push_back(s.max - 1); idx.push_back(s.max); idx.push_back(s.max + 1);
    }
    std::sort(idx.begin(), idx.end());
    idx.erase(std::unique(idx.begin(), idx.end()), idx.end());

    std::vector<Seg> out;  // `list(siz*2)`; growth beyond that is out of bounds in the source
    Seg nw = SEG_UPPER;
    auto reg = [&](Seg& n) {  // register_seg_list :189-202
        if (n.validate()) out.push_back(n);
        n = SEG_UPPER;
    };
    for (size_t m = 0; m < idx.size(); m++) {
        int i = idx[m];
        bool inside = false;
        for (auto& s : buff) inside = inside || in_seg(i, s);
        if (inside) { if (i < nw.min) nw.min = i; } else continue;
        bool flag = false;
        for (auto& s : buff) if (i + 1 == s.min) flag = true;
        if (flag) { nw.max = i; reg(nw); continue; }
        long count = 0;
        for (auto& s : buff) if (s.min == i) count++;
        if (count > 1) { nw.max = i; reg(nw); }
        count = 0;
        for (auto& s : buff) if (s.max == i) count++;
        if (count > 0) { nw.max = i; reg(nw); }
    }
    list = out;
}

// ---------------------------------------------------------------------------------------------
// tokens / AST (src/essential/enums_m.f90, src/ast/syntax_tree_node_m.F90, syntax_tree_graph_m.F90)
// ---------------------------------------------------------------------------------------------
enum { tk_char = 0, tk_union, tk_lpar, tk_rpar, tk_backslash, tk_question, tk_star, tk_plus, tk_lsbracket,
       tk_rsbracket, tk_lcurlybrace, tk_rcurlybrace, tk_dot, tk_hyphen, tk_caret, tk_dollar, tk_end };
enum { op_not_init = 0, op_char, op_concat, op_union, op_closure, op_repeat, op_empty };

static inline fstr char4(const fstr& s) {  // assignment to character(UTF8_CHAR_SIZE)
    fstr r = s.substr(0, std::min<size_t>(4, s.size()));
    r.resize(4, ' ');
    return r;
}

struct Tape {  // syntax_tree_node_m.F90:51-64
    fstr str;
    int current_token = 0;
    fstr token_char = char4(fstr(1, '\0'));  // EMPTY = char(0), blank padded
    long idx = 0;

    void get_token(bool class_flag = false) {  // :133-215
        long ib = idx;
        if (ib == INVALID_CHAR_INDEX || ib > (long)str.size()) {
            current_token = tk_end;
            token_char = char4("");
            return;
        }
        long ie = idxutf8(str, ib);
        fstr c = char4(sub(str, ib, ie));
        fstr tc = f_trim(c);
        if (class_flag) {
            if (f_eq(tc, "]")) current_token = tk_rsbracket;
            else if (f_eq(tc, "-")) current_token = tk_hyphen;
            else if (f_eq(tc, "\\")) current_token = tk_backslash;
            else current_token = tk_char;
            token_char = c;
        } else {
            if (f_eq(tc, "|")) current_token = tk_union;
            else if (f_eq(tc, "(")) current_token = tk_lpar;
            else if (f_eq(tc, ")")) current_token = tk_rpar;
            else if (f_eq(tc, "*")) current_token = tk_star;
            else if (f_eq(tc, "+")) current_token = tk_plus;
            else if (f_eq(tc, "?")) current_token = tk_question;
            else if (f_eq(tc, "\\")) {
                current_token = tk_backslash;
                ib = next_idxutf8(str, ie);
                ie = idxutf8(str, ib);
                token_char = char4(sub(str, ib, ie));  // empty when the backslash is the last byte
            } else if (f_eq(tc, "[")) current_token = tk_lsbracket;
            else if (f_eq(tc, "]")) current_token = tk_rsbracket;
            else if (f_eq(tc, "{")) { current_token = tk_lcurlybrace; token_char = c; }
            else if (f_eq(tc, "}")) { current_token = tk_rcurlybrace; token_char = c; }
            else if (f_eq(tc, ".")) current_token = tk_dot;
            else if (f_eq(tc, "^")) current_token = tk_caret;
            else if (f_eq(tc, "$")) current_token = tk_dollar;
            else { current_token = tk_char; token_char = c; }
        }
        idx = next_idxutf8(str, ib);
    }
};

struct TreeNode {  // syntax_tree_node_m.F90:35-49
    int op = op_not_init;
    std::vector<Seg> c;
    bool has_c = false;  // allocated(c)
    int left_i = INVALID_INDEX, right_i = INVALID_INDEX, parent_i = INVALID_INDEX, own_i = INVALID_INDEX;
    int min_repeat = 0, max_repeat = 0;
};

static TreeNode terminal_node() {  // :66-73 (own_i = INVALID_INDEX is what callers read)
    TreeNode t;
    t.left_i = 0; t.right_i = 0;
    t.min_repeat = INVALID_REPEAT_VAL; t.max_repeat = INVALID_REPEAT_VAL;
    return t;
}
static TreeNode make_atom(const Seg& s) {
    TreeNode n; n.op = op_char; n.c.assign(1, s); n.has_c = true; return n;
}
static TreeNode make_tree_node(int op) { TreeNode n; n.op = op; return n; }

// character_array_t (src/ast/character_array_m.F90:16-28)
struct CharArr {
    fstr c;
    bool is_escaped = false, is_hyphenated = false, is_subtract = false;
    int seg_size = 0;
};

struct Tree {  // syntax_tree_graph_m.F90:22-51
    std::vector<TreeNode> nodes;  // 1-based: nodes[i-1]
    int top = INVALID_INDEX;
    Tape tape;
    bool is_valid = true;
    int code = SYNTAX_VALID;
    int paren_balance = 0;

    TreeNode& at(int i) { return nodes[(size_t)i - 1]; }
    const TreeNode& at(int i) const { return nodes[(size_t)i - 1]; }
    TreeNode get_top() const { return at(top); }  // :193-199

    void register_node(TreeNode& node) {  // :141-158 with the limit of :101-129
        int t = top + 1;
        if (t > TREE_NODE_HARD_LIMIT) throw ErrorStop{ERRSTOP_TREE_LIMIT};
        if (t > (int)nodes.size()) nodes.resize((size_t)t);
        node.own_i = t;
        nodes[(size_t)t - 1] = node;
        top = t;
    }
    void register_connector(TreeNode& node, const TreeNode& left, const TreeNode& right) {  // :161-191
        register_node(node);
        int parent = at(top).own_i;
        at(parent).left_i = left.own_i;
        if (left.own_i != INVALID_INDEX) at(left.own_i).parent_i = parent;
        at(parent).right_i = right.own_i;
        if (right.own_i != INVALID_INDEX) at(right.own_i).parent_i = parent;
    }

    void build(const fstr& pattern) {  // :61-95
        nodes.clear();
        tape = Tape();
        tape.idx = 1;
        tape.str = pattern;
        top = 0;
        paren_balance = 0;
        is_valid = true;
        code = SYNTAX_VALID;
        tape.get_token();
        regex();
        if (!is_valid) return;
        if (paren_balance > 0) { is_valid = false; code = SYNTAX_ERR_PARENTHESIS_MISSING; }
        else if (paren_balance < 0) { is_valid = false; code = SYNTAX_ERR_PARENTHESIS_UNEXPECTED; }
        at(top).parent_i = 0;
    }

    void regex() {  // :205-238
        term();
        if (is_valid) {
            TreeNode left = get_top();
            while (tape.current_token == tk_union) {
                tape.get_token();
                term();
                if (!is_valid) break;
                TreeNode right = get_top();
                TreeNode node = make_tree_node(op_union);
                register_connector(node, left, right);
                left = get_top();
            }
        }
    }

    void term() {  // :241-278
        TreeNode term_ = terminal_node();
        int t = tape.current_token;
        if (t == tk_union || t == tk_rpar || t == tk_end) {
            TreeNode node = make_tree_node(op_empty);
            register_connector(node, term_, term_);
        } else {
            suffix_op();
            if (!is_valid) return;
            TreeNode left = get_top();
            while (tape.current_token != tk_union && tape.current_token != tk_rpar && tape.current_token != tk_end) {
                suffix_op();
                if (!is_valid) return;
                TreeNode right = get_top();
                TreeNode node = make_tree_node(op_concat);
                register_connector(node, left, right);
                left = get_top();
            }
        }
        if (tape.current_token == tk_rpar) paren_balance--;
    }

    void suffix_op() {  // :281-326
        TreeNode term_ = terminal_node();
        primary();
        if (!is_valid) return;
        TreeNode left = get_top();
        switch (tape.current_token) {
            case tk_star: {
                TreeNode node = make_tree_node(op_closure);
                register_connector(node, left, term_);
                tape.get_token();
                break;
            }
            case tk_plus: {
                TreeNode node = make_tree_node(op_closure);
                register_connector(node, left, term_);
                TreeNode right = get_top();
                node = make_tree_node(op_concat);
                register_connector(node, left, right);
                tape.get_token();
                break;
            }
            case tk_question: {
                TreeNode node = make_tree_node(op_empty);
                register_connector(node, left, term_);
                TreeNode right = get_top();
                node = make_tree_node(op_union);
                register_connector(node, left, right);
                tape.get_token();
                break;
            }
            case tk_lcurlybrace:
                times();
                if (!is_valid) return;
                tape.get_token();
                break;
            default: break;
        }
    }

    void fail(int c) { code = c; is_valid = false; }

    void primary() {  // :329-443
        TreeNode term_ = terminal_node();
        switch (tape.current_token) {
            case tk_char:
            case tk_rcurlybrace: {  // an unescaped closing brace is a literal (:410-415)
                int cp = ichar_utf8(tape.token_char);
                TreeNode node = make_atom(Seg(cp, cp));
                register_connector(node, term_, term_);
                tape.get_token();
                break;
            }
            case tk_lpar:
                paren_balance++;
                tape.get_token();
                regex();
                if (!is_valid) return;
                if (tape.current_token != tk_rpar) { fail(SYNTAX_ERR_PARENTHESIS_MISSING); return; }
                tape.get_token();
                break;
            case tk_lsbracket:
                char_class();
                if (!is_valid) return;
                if (tape.current_token != tk_rsbracket) { fail(SYNTAX_ERR_BRACKET_MISSING); return; }
                tape.get_token();
                break;
            case tk_backslash:
                shorthand();
                if (!is_valid) return;
                tape.get_token();
                break;
            case tk_dot: {
                TreeNode node = make_atom(SEG_ANY);
                register_connector(node, term_, term_);
                tape.get_token();
                break;
            }
            case tk_caret:
            case tk_dollar:
                caret_dollar();
                tape.get_token();
                break;
            case tk_rsbracket: fail(SYNTAX_ERR_BRACKET_UNEXPECTED); return;
            case tk_rpar: fail(SYNTAX_ERR_PARENTHESIS_UNEXPECTED); return;
            case tk_lcurlybrace: fail(SYNTAX_ERR_INVALID_TIMES); return;
            case tk_star: fail(SYNTAX_ERR_STAR_INCOMPLETE); return;
            case tk_plus: fail(SYNTAX_ERR_PLUS_INCOMPLETE); return;
            case tk_question: fail(SYNTAX_ERR_QUESTION_INCOMPLETE); return;
            default: fail(SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN); return;
        }
    }

    void char_class();   // :448-556
    void crlf();         // :559-577
    void caret_dollar(); // :581-605
    void shorthand();    // :611-725
    void hex2seg_tree(std::vector<Seg>& seglist);  // :728-777
    void times();        // :782-906
};

void Tree::crlf() {
    TreeNode t = terminal_node();
    TreeNode cr = make_atom(SEG_CR); register_connector(cr, t, t);
    TreeNode lf = make_atom(SEG_LF); register_connector(lf, t, t);
    TreeNode right = make_tree_node(op_concat); register_connector(right, cr, lf);
    TreeNode node = make_tree_node(op_union); register_connector(node, lf, right);
}

void Tree::caret_dollar() {
    TreeNode t = terminal_node();
    TreeNode cr = make_atom(SEG_CR); register_connector(cr, t, t);
    TreeNode lf = make_atom(SEG_LF); register_connector(lf, t, t);
    TreeNode node_r_r = make_tree_node(op_concat); register_connector(node_r_r, cr, lf);
    TreeNode node_r = make_tree_node(op_union); register_connector(node_r, lf, node_r_r);
    TreeNode empty_r = make_atom(SEG_EMPTY); register_connector(empty_r, t, t);
    TreeNode node = make_tree_node(op_union); register_connector(node, node_r, empty_r);
}

static std::vector<Seg> segs_w() { return {SEG_LOWERCASE, SEG_UPPERCASE, SEG_DIGIT, SEG_UNDERSCORE}; }
static std::vector<Seg> segs_s() { return {SEG_SPACE, SEG_TAB, SEG_CR, SEG_LF, SEG_FF, SEG_ZENKAKU_SPACE}; }

void Tree::shorthand() {
    TreeNode t = terminal_node();
    fstr tc = f_trim(tape.token_char);
    std::vector<Seg> seglist;
    auto atom = [&](const Seg& s) { TreeNode n = make_atom(s); register_connector(n, t, t); };
    if (f_eq(tc, "t")) { atom(SEG_TAB); return; }
    else if (f_eq(tc, "n")) { crlf(); return; }
    else if (f_eq(tc, "r")) { atom(SEG_CR); return; }
    else if (f_eq(tc, "d")) { atom(SEG_DIGIT); return; }
    else if (f_eq(tc, "D")) { seglist = {SEG_DIGIT}; invert_segment_list(seglist); }
    else if (f_eq(tc, "w")) { seglist = segs_w(); }
    else if (f_eq(tc, "W")) { seglist = segs_w(); invert_segment_list(seglist); }
    else if (f_eq(tc, "s")) { seglist = segs_s(); }
    else if (f_eq(tc, "S")) { seglist = segs_s(); invert_segment_list(seglist); }
    else if (f_eq(tc, "x")) { hex2seg_tree(seglist); if (!is_valid) return; }
    else if (f_eq(tc, "")) { fail(SYNTAX_ERR_ESCAPED_SYMBOL_MISSING); return; }
    else if (f_eq(tc, "[") || f_eq(tc, "]") || f_eq(tc, "{") || f_eq(tc, "}") || f_eq(tc, "(") || f_eq(tc, ")") ||
             f_eq(tc, "$") || f_eq(tc, "\\") || f_eq(tc, "|") || f_eq(tc, ".") || f_eq(tc, "?") || f_eq(tc, "^") ||
             f_eq(tc, "*") || f_eq(tc, "+") || f_eq(tc, "-")) {
        int cp = ichar_utf8(tape.token_char);
        atom(Seg(cp, cp));
        return;
    } else { fail(SYNTAX_ERR_ESCAPED_SYMBOL_INVALID); return; }
    TreeNode node;
    node.c = seglist; node.has_c = true; node.op = op_char;
    register_connector(node, t, t);
}

void Tree::hex2seg_tree(std::vector<Seg>& seglist) {
    fstr hex;
    tape.get_token();
    bool is_longer = tape.current_token == tk_lcurlybrace;
    bool is_two = !is_longer;
    if (is_longer) tape.get_token();
    hex = tape.token_char.substr(0, 1);
    int i = 2;
    while (true) {
        if (is_two && i >= 3) break;
        tape.get_token();
        if (is_longer && tape.current_token != tk_rcurlybrace && tape.current_token != tk_char) {
            is_valid = false; code = SYNTAX_ERR_CURLYBRACE_MISSING; return;
        }
        if (tape.current_token == tk_rcurlybrace) break;
        hex += tape.token_char.substr(0, 1);
        i++;
    }
    seglist.assign(1, Seg());
    hex2seg(f_trim(hex), seglist[0], code);
    if (code != SYNTAX_VALID) { is_valid = false; return; }
    is_valid = seg_in_seg(seglist[0], SEG_WHOLE);
    if (!is_valid) code = SYNTAX_ERR_UNICODE_EXCEED;
}

// src/essential/utility_m.f90:122-141
static void get_index_comma(const fstr& str, long& i, long& count) {
    i = 0; count = 0;
    fstr buf = str;
    while (true) {
        long j = f_index(buf, ",");
        if (i == 0) i = j;
        if (j == 0) break;
        buf[(size_t)j - 1] = '.';
        count++;
    }
}
// I-format read of the whole field: blanks ignored, optional sign, digits only.
static bool read_iformat(const fstr& s, long long& val) {
    fstr t;
    for (char ch : s) if (ch != ' ') t += ch;
    if (t.empty()) { val = 0; return true; }  // an all-blank field reads as zero
    size_t p = 0;
    bool neg = false;
    if (t[p] == '+' || t[p] == '-') { neg = t[p] == '-'; p++; }
    if (p >= t.size()) return false;
    long long v = 0;
    for (; p < t.size(); p++) {
        if (t[p] < '0' || t[p] > '9') return false;
        v = v * 10 + (t[p] - '0');
        if (v > 4000000000000000000ll) return false;
    }
    val = neg ? -v : v;
    return true;
}
// utility_m.f90:145-169
static bool is_integer(const fstr& chara) {
    long i = std::max(f_index(chara, ","), f_index(chara, " "));
    if (i != 0) return false;
    if (chara.size() > 19) {  // '(1i19)' reads only the first 19 columns
        long long v;
        return read_iformat(chara.substr(0, 19), v);
    }
    long long v;
    return read_iformat(chara, v);
}
// list-directed read of one default integer from an internal file (`read(c, fmt=*, iostat=ios) n`):
// ios < 0 (end of file) when there is no value; ios > 0 on a malformed or overflowing item.
// Returns ios sign; leaves val untouched when nothing is read.
static int read_listdirected_int(const fstr& s, int& val) {
    size_t p = 0;
    while (p < s.size() && s[p] == ' ') p++;
    if (p >= s.size()) return -1;
    if (s[p] == '/') return 0;        // slash terminates the list: nothing assigned
    if (s[p] == ',') return 0;        // null value
    size_t q = p;
    bool neg = false;
    if (s[q] == '+' || s[q] == '-') { neg = s[q] == '-'; q++; }
    if (q >= s.size() || s[q] < '0' || s[q] > '9') return 1;
    long long v = 0;
    while (q < s.size() && s[q] >= '0' && s[q] <= '9') {
        v = v * 10 + (s[q] - '0');
        if (v > 0x7FFFFFFFll + 1) return 1;
        q++;
    }
    if (q < s.size() && !(s[q] == ' ' || s[q] == ',' || s[q] == '/')) return 1;
    v = neg ? -v : v;
    if (v > 0x7FFFFFFFll || v < -0x80000000ll) return 1;
    val = (int)v;
    return 0;
}

void Tree::times() {
    TreeNode t = terminal_node();
    fstr buf;
    int arg[2] = {INVALID_REPEAT_VAL, INVALID_REPEAT_VAL};
    bool is_infinite = false;
    int min = INVALID_REPEAT_VAL, max = INVALID_REPEAT_VAL;
    tape.get_token();
    while (tape.current_token != tk_rcurlybrace) {
        buf += f_trim(tape.token_char);
        tape.get_token();
        if (tape.current_token == tk_end) { fail(SYNTAX_ERR_CURLYBRACE_MISSING); return; }
    }
    if (buf.size() == 0) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    else if (buf.size() == 1) {
        if (buf[0] == ',') { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    }
    if (buf[0] == ',') buf = "0" + buf;
    if (is_integer(buf)) buf = f_trim(buf) + "," + f_trim(buf);
    long i, num_comma;
    get_index_comma(buf, i, num_comma);
    if (num_comma > 1) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    fstr c1 = sub(buf, 1, i - 1), c2;
    if (i + 1 <= f_len_trim(buf)) c2 = sub(buf, i + 1, f_len_trim(buf));
    int ios = read_listdirected_int(c1, arg[0]);
    if (ios > 0 || arg[0] < 0) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    if (f_eq(f_trim(c2), "")) {
        is_infinite = true;
    } else {
        ios = read_listdirected_int(c2, arg[1]);
        if (ios > 0 || arg[1] < 0) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    }
    if (is_infinite) { min = arg[0]; max = INFINITE_; }
    else { min = arg[0]; max = arg[1]; }
    if (min == 0 && max == 0) {
    } else if (max != INFINITE_ && min > max) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    else if (max == INVALID_REPEAT_VAL && min > max) { fail(SYNTAX_ERR_INVALID_TIMES); return; }
    TreeNode node;
    node.op = op_repeat; node.min_repeat = min; node.max_repeat = max;
    TreeNode left = get_top();
    register_connector(node, left, t);
}

// ---- character classes ------------------------------------------------------------------
// character_array_m.F90:45-70
static bool character_string_to_array(const fstr& str, std::vector<CharArr>& array) {
    long siz = len_utf8(str);
    if (siz < 1) return false;
    array.assign((size_t)siz, CharArr());
    long ib = 0, ie = 0;
    for (long j = 1; j <= siz; j++) {
        ib = ie + 1;
        ie = idxutf8(str, ib);
        if (ib == INVALID_CHAR_INDEX || ie == INVALID_CHAR_INDEX) return true;
        array[(size_t)j - 1].c = sub(str, ib, ie);
    }
    return true;
}

// character_array_m.F90:75-140.  `temp(k-1)` with k == 1 is an out-of-bounds store in the source
// (a hyphen right after a leading backslash, e.g. `[\-a]`); it is dropped here.
static void parse_backslash_and_hyphen_in_char_array(std::vector<CharArr>& array, int& ierr) {
    long n = (long)array.size();
    if (n < 1) return;
    std::vector<CharArr> temp((size_t)n);
    long k = 1;
    bool zone = false;
    for (long i = 1; i <= n; i++) {
        if (1 < i && i < n) {
            bool hh = f_eq(array[(size_t)i - 1].c, "-") && f_eq(array[(size_t)i].c, "-");
            if (!zone) {
                if (hh) {
                    for (long q = k; q <= n; q++) temp[(size_t)q - 1].is_subtract = true;
                    zone = true;
                    continue;
                }
            } else {
                if (hh) { ierr = SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR; return; }
            }
            if (f_eq(array[(size_t)i - 2].c, "-") && f_eq(array[(size_t)i - 1].c, "-")) continue;
        }
        // temp(k) with k == n+1 can be read here in the source when every earlier element was a
        // plain character; it cannot happen because k <= i.
        if (f_eq(array[(size_t)i - 1].c, "\\") && !temp[(size_t)k - 1].is_escaped) {
            temp[(size_t)k - 1].is_escaped = true;
        } else if (f_eq(array[(size_t)i - 1].c, "-") && i != 1) {
            if (k - 1 >= 1) temp[(size_t)k - 2].is_hyphenated = true;
        } else {
            temp[(size_t)k - 1].c = array[(size_t)i - 1].c;
            k++;
        }
    }
    long siz = k - 1;
    array.assign(temp.begin(), temp.begin() + siz);
}

// character_array_m.F90:145-222
static void parse_segment_width_in_char_array(std::vector<CharArr>& array) {
    for (auto& a : array) {
        int n;
        if (a.is_escaped) {
            const fstr& c = a.c;
            if (f_eq(c, "t")) n = 1;
            else if (f_eq(c, "n")) n = 2;
            else if (f_eq(c, "r")) n = 1;
            else if (f_eq(c, "d")) n = width_of_segment(SEG_DIGIT);
            else if (f_eq(c, "D")) { std::vector<Seg> s = {SEG_DIGIT}; invert_segment_list(s); n = total_width_of_segment(s); }
            else if (f_eq(c, "w")) { n = total_width_of_segment(segs_w()); }
            else if (f_eq(c, "W")) { std::vector<Seg> s = segs_w(); invert_segment_list(s); n = total_width_of_segment(s); }
            else if (f_eq(c, "s")) n = 6;
            else if (f_eq(c, "S")) { std::vector<Seg> s = segs_s(); invert_segment_list(s); n = total_width_of_segment(s); }
            else if (f_eq(c, "x") || f_eq(c, "\\") || f_eq(c, "{") || f_eq(c, "}") || f_eq(c, "[") || f_eq(c, "]")) n = 1;
            else n = -1;
        } else {
            n = 1;
        }
        a.seg_size = n;
    }
}

// character_array_m.F90:225-332
static void parse_escape_sequence_with_argument(std::vector<CharArr>& ca, int& ierr) {
    ierr = SYNTAX_VALID;
    long siz = (long)ca.size();
    std::vector<CharArr> tmp((size_t)siz);
    fstr hex_long;
    long k = 1, j = 1;
    auto C = [&](long q) -> CharArr& { return ca[(size_t)q - 1]; };
    auto T = [&](long q) -> CharArr& { return tmp[(size_t)q - 1]; };
    while (j <= siz) {
        if (f_eq(C(j).c, "x") && C(j).is_escaped) {
            T(k).c = "x";
            T(k).is_escaped = true;
            j++;
            if (j > siz) break;
            k++;
            if (j + 1 <= siz) {
                if (in_hex(ichar_utf8(C(j).c)) && in_hex(ichar_utf8(C(j + 1).c))) {
                    fstr two = f_trim(C(j).c) + f_trim(C(j + 1).c);
                    two.resize(2, ' ');
                    T(k).c = f_trim(f_adjustl(two));
                    T(k).is_hyphenated = C(j + 1).is_hyphenated;
                    j += 2;
                    if (j > siz) break;
                    k++;
                    continue;
                } else if (f_eq(C(j).c, "{")) {
                    long i = j + 1;
                    while (true) {
                        if (i > siz) { ierr = SYNTAX_ERR_CURLYBRACE_MISSING; return; }
                        if (!f_eq(C(i).c, "}") && !in_hex(ichar_utf8(C(i).c))) { ierr = SYNTAX_ERR_INVALID_HEXADECIMAL; return; }
                        else if (f_eq(C(i).c, "}")) break;
                        hex_long = f_trim(f_adjustl(hex_long)) + C(i).c;
                        i++;
                    }
                    T(k).c = f_trim(f_adjustl(hex_long));
                    T(k).is_hyphenated = C(i).is_hyphenated;
                    j = i + 1;
                    if (j > siz) break;
                    k++;
                    hex_long.clear();
                    continue;
                } else { ierr = SYNTAX_ERR_INVALID_HEXADECIMAL; return; }
            } else { ierr = SYNTAX_ERR_HEX_DIGITS_NOT_ENOUGH; return; }
        } else if (f_eq(C(j).c, "p")) {
            ierr = SYNTAX_ERR_UNICODE_PROPERTY_NOT_IMPLEMENTED; return;
        }
        T(k) = C(j);
        j++;
        if (j > siz) break;
        k++;
    }
    ca.assign(tmp.begin(), tmp.begin() + k);
}

// syntax_tree_graph_m.F90:1123-1213
static std::vector<Seg> convert_escaped_character_into_segments(const fstr& chara) {
    fstr c = f_trim(chara);
    std::vector<Seg> l;
    auto one = [&](int cp) { l.assign(1, Seg(cp, cp)); };
    if (f_eq(c, "t")) l = {SEG_TAB};
    else if (f_eq(c, "n")) l = {SEG_LF, SEG_CR};
    else if (f_eq(c, "r")) l = {SEG_CR};
    else if (f_eq(c, "d")) l = {SEG_DIGIT};
    else if (f_eq(c, "D")) { l = {SEG_DIGIT}; invert_segment_list(l); }
    else if (f_eq(c, "w")) l = segs_w();
    else if (f_eq(c, "W")) { l = segs_w(); invert_segment_list(l); }
    else if (f_eq(c, "s")) l = segs_s();
    else if (f_eq(c, "S")) { l = segs_s(); invert_segment_list(l); }
    else if (f_eq(c, "x")) { l.assign(1, Seg()); int unused; hex2seg(chara, l[0], unused); }
    else if (f_eq(c, "p")) l = {SEG_ERROR};
    else if (f_eq(c, "\\")) one('\\');
    else if (f_eq(c, "{")) one('{');
    else if (f_eq(c, "}")) one('}');
    else if (f_eq(c, "[")) one('[');
    else if (f_eq(c, "]")) one(']');
    else l = {SEG_ERROR};
    return l;
}

// syntax_tree_graph_m.F90:910-1118
static void interpret_class_string(const fstr& str, std::vector<Seg>& seglist, bool& seglist_allocated,
                                   bool& is_valid, int& ierr) {
    is_valid = true;
    seglist_allocated = false;
    bool backslashed = false, prev_hyphenated = false, curr_hyphenated = false;
    Seg prev_seg, curr_seg;
    if (str.size() >= 2 && sub(str, 1, 2) == "--") {
        ierr = SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR;
        is_valid = false;  // no return here in the source (:940-945)
    }
    std::vector<CharArr> ca;
    if (!character_string_to_array(str, ca)) { ierr = SYNTAX_ERR_EMPTY_CHARACTER_CLASS; is_valid = false; return; }
    parse_backslash_and_hyphen_in_char_array(ca, ierr);
    if (ierr == SYNTAX_ERR_MISPLACED_SUBTRACTION_OPERATOR) { is_valid = false; return; }
    parse_escape_sequence_with_argument(ca, ierr);
    if (ierr != SYNTAX_VALID) { is_valid = false; return; }
    parse_segment_width_in_char_array(ca);

    long siz = 0;
    for (long i = 1; i <= (long)ca.size(); i++) {
        CharArr& e = ca[(size_t)i - 1];
        if (e.is_hyphenated && e.seg_size != 1) { ierr = SYNTAX_ERR_RANGE_WITH_ESCAPE_SEQUENCES; is_valid = false; return; }
        if (i > 1 && ca[(size_t)i - 2].is_hyphenated && e.seg_size != 1) {
            ierr = SYNTAX_ERR_RANGE_WITH_ESCAPE_SEQUENCES; is_valid = false; return;
        }
        if (e.is_subtract) { ierr = SYNTAX_ERR_CHAR_CLASS_SUBTRANCTION_NOT_IMPLEMENTED; is_valid = false; return; }
        if (i > 1 && i == (long)ca.size()) {
            if (e.is_hyphenated) {
                e.is_hyphenated = false;
                CharArr h; h.c = "-"; h.is_subtract = e.is_subtract; h.seg_size = 1;
                ca.push_back(h);
                siz = siz + 1;
                break;  // leaves before adding this element's own size (:1012-1016)
            }
        }
        siz += ca[(size_t)i - 1].seg_size;
    }
    if (siz < 1) { ierr = SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN; is_valid = false; return; }
    std::vector<Seg> list((size_t)siz);

    long j = 0;
    long i = 1;
    while (i <= (long)ca.size()) {
        fstr c = ca[(size_t)i - 1].c;
        backslashed = ca[(size_t)i - 1].is_escaped;
        curr_hyphenated = ca[(size_t)i - 1].is_hyphenated;
        if (i > 1) prev_hyphenated = ca[(size_t)i - 2].is_hyphenated;
        if (backslashed && f_eq(c, "x")) {
            i++;
            if (i > (long)ca.size()) { ierr = SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN; is_valid = false; return; }
            c = ca[(size_t)i - 1].c;
            backslashed = ca[(size_t)i - 1].is_escaped;
            hex2seg(c, curr_seg, ierr);
            if (ierr != SYNTAX_VALID) { is_valid = false; return; }
        } else if (backslashed && f_eq(c, "p")) {
            ierr = SYNTAX_ERR_UNICODE_PROPERTY_NOT_IMPLEMENTED; is_valid = false; return;
        } else {
            int cp = ichar_utf8(c);
            curr_seg = Seg(cp, cp);
        }
        if (backslashed) {
            std::vector<Seg> cache = convert_escaped_character_into_segments(c);
            if (cache[0] == SEG_ERROR) { ierr = SYNTAX_ERR_ESCAPED_SYMBOL_INVALID; is_valid = false; return; }
            if (cache.size() > 1) {
                for (auto& s : cache) register_segment_to_list(list, s, j, ierr);
                prev_seg = Seg();
                i++;
                continue;
            }
            curr_seg = cache[0];
        }
        if (prev_hyphenated) {
            curr_seg = join_two_segments(prev_seg, curr_seg);
            if (curr_seg == SEG_ERROR) { ierr = SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN; is_valid = false; return; }
        }
        if (!curr_hyphenated) {
            int jerr;
            register_segment_to_list(list, curr_seg, j, jerr);
            if (jerr == SEGMENT_REJECTED) { ierr = SYNTAX_ERR_INVALID_CHARACTER_RANGE; is_valid = false; return; }
        }
        prev_seg = curr_seg;
        i++;
    }
    if (j < 1) { ierr = SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN; is_valid = false; return; }
    seglist.assign(list.begin(), list.begin() + j);
    seglist_allocated = true;
}

void Tree::char_class() {
    TreeNode t = terminal_node();
    tape.get_token(true);
    fstr buf;
    bool backslashed = false;
    while (tape.current_token != tk_rsbracket) {
        if (tape.current_token == tk_end) return;
        long ie = idxutf8(tape.token_char, 1);
        buf += sub(tape.token_char, 1, ie);
        if (tape.current_token == tk_backslash && !backslashed) backslashed = true;
        else backslashed = false;
        tape.get_token(true);
        if (tape.current_token == tk_rsbracket && backslashed) {
            ie = idxutf8(tape.token_char, 1);
            buf += sub(tape.token_char, 1, ie);
            tape.get_token(true);
        }
    }
    if (buf.size() == 0) { fail(SYNTAX_ERR_EMPTY_CHARACTER_CLASS); return; }
    bool is_inverted = false;
    if (buf[0] == '^') { is_inverted = true; buf = buf.substr(1); }
    long siz = len_utf8(buf);
    if (siz < 1) { fail(SYNTAX_ERR_EMPTY_CHARACTER_CLASS); return; }
    std::vector<Seg> seglist;
    bool allocated_;
    interpret_class_string(buf, seglist, allocated_, is_valid, code);
    if (!is_valid) return;
    if (!allocated_) { fail(ALLOCATION_ERR); return; }
    if (seglist.size() < 1) { fail(SYNTAX_ERR_THIS_SHOULD_NOT_HAPPEN); return; }
    if (is_inverted) invert_segment_list(seglist);
    TreeNode node = make_tree_node(op_char);
    node.c = seglist; node.has_c = true;
    register_connector(node, t, t);
}

// ---------------------------------------------------------------------------------------------
// literal extraction (src/ast/syntax_tree_optimize_m.F90)
// ---------------------------------------------------------------------------------------------
struct Literal {  // :29-33
    fstr all, pref, suff, fact;
    bool flag_closure = false, flag_class = false;
};

static fstr best(const fstr& c1, const fstr& c2) {  // :229-241
    if (f_len_trim(c1) > f_len_trim(c2)) return f_trim(f_adjustl(c1));
    return f_trim(f_adjustl(c2));
}
static fstr same_part_of_prefix(const fstr& c1, const fstr& c2) {  // :244-273
    fstr res;
    long i = 1;
    while (true) {
        fstr p1 = sub(c1, i, idxutf8(c1, i));
        fstr p2 = sub(c2, i, idxutf8(c2, i));
        bool ret = next_idxutf8(c1, i) == INVALID_CHAR_INDEX || next_idxutf8(c2, i) == INVALID_CHAR_INDEX;
        if (ret) return res;
        if (f_eq(p1, p2)) res += p1; else break;
        i = next_idxutf8(c1, i);
    }
    return res;
}
static fstr same_part_of_suffix(const fstr& c1, const fstr& c2) {  // :276-291
    return reverse_utf8(same_part_of_prefix(reverse_utf8(c1), reverse_utf8(c2)));
}

// :71-226.  `lit`'s flags are intent(inout) and never reset, exactly as in the source.
static void best_factor(const Tree& tree, int idx, Literal& lit) {
    Literal lit_l, lit_r;
    const TreeNode& curr = tree.at(idx);
    lit.all.clear(); lit.pref.clear(); lit.suff.clear(); lit.fact.clear();
    if (curr.op == op_union || curr.op == op_concat) {
        best_factor(tree, curr.left_i, lit_l);
        best_factor(tree, curr.right_i, lit_r);
    }
    switch (curr.op) {
        case op_union:
            lit.pref = same_part_of_prefix(lit_l.pref, lit_r.pref);
            lit.suff = same_part_of_suffix(lit_l.suff, lit_r.suff);
            lit.flag_closure = true;
            break;
        case op_concat: {
            lit.flag_class = lit_l.flag_class || lit_r.flag_class;
            lit.flag_closure = lit_l.flag_closure || lit_r.flag_closure;
            bool Lc = lit_l.flag_class, Rc = lit_r.flag_class, Lk = lit_l.flag_closure, Rk = lit_r.flag_closure;
            if (!Lc && !Rc) {
                if (!Lk && !Rk) {          // N_class_N_closure
                    lit.all = lit_l.all + lit_r.all;
                    lit.pref = best(lit_l.pref, lit_l.all + lit_r.pref);
                    lit.suff = best(lit_r.suff, lit_l.suff + lit_r.all);
                } else if (!Lk && Rk) {    // N_class_R_closure
                    lit.pref = lit_l.all + lit_r.pref;
                    lit.suff = lit_r.suff;
                } else if (Lk && !Rk) {    // N_class_L_closure
                    lit.pref = lit_l.pref;
                    lit.suff = lit_l.suff + lit_r.all;
                } else {                   // N_class_LR_closure
                    lit.pref = lit_l.pref;
                    lit.suff = lit_r.suff;
                }
            } else if (!Lc && Rc) {
                if (!Lk) {                 // R_class_N_closure, R_class_R_closure
                    lit.pref = best(lit_l.pref, lit_l.all + lit_r.pref);
                    lit.suff = lit_r.suff;
                } else {                   // R_class_L_closure, R_class_LR_closure
                    lit.pref = lit_l.pref;
                    lit.suff = lit_r.suff;
                }
            } else if (Lc && !Rc) {
                if (!Rk) {                 // L_class_N_closure, L_class_L_closure
                    lit.pref = lit_l.pref;
                    lit.suff = best(lit_r.suff, lit_l.suff + lit_r.all);
                } else {                   // L_class_R_closure, L_class_LR_closure
                    lit.pref = lit_l.pref;
                    lit.suff = lit_r.suff;
                }
            } else {
                if (!Lk && Rk) {           // LR_class_R_closure: suffix is never assigned (:165-168)
                    lit.pref = lit_l.pref;
                } else {
                    lit.pref = lit_l.pref;
                    lit.suff = lit_r.suff;
                }
            }
            break;
        }
        case op_closure:
            lit.flag_closure = true;
            break;
        case op_char:
            if (curr.has_c) {
                if (curr.c.size() == 1) {
                    if (width_of_segment(curr.c[0]) == 1) {
                        lit.all = lit.pref = lit.suff = lit.fact = char_utf8(curr.c[0].min);
                    } else lit.flag_class = true;
                } else lit.flag_class = true;
            }
            break;
        case op_repeat: {
            best_factor(tree, curr.left_i, lit_l);
            lit.flag_class = lit_l.flag_class;
            for (int i = 1; i <= curr.min_repeat; i++) {
                best_factor(tree, curr.left_i, lit_l);
                lit.all += lit_l.all;
                lit.pref += lit_l.pref;
                lit.suff += lit_l.suff;
                lit.fact += lit_l.fact;
                lit.flag_class = lit.flag_class || lit_l.flag_class;
                if (lit_l.flag_closure) break;
            }
            lit.flag_closure = curr.min_repeat != curr.max_repeat;
            lit.flag_closure = lit.flag_closure || lit_l.flag_closure;
            break;
        }
        default:
            lit.flag_closure = true;
            break;
    }
}

static void extract_literal(const Tree& tree, fstr& all, fstr& prefix, fstr& suffix) {  // :42-67
    Literal lit;
    best_factor(tree, tree.top, lit);
    all = lit.all; prefix = lit.pref; suffix = lit.suff;
}

// ---------------------------------------------------------------------------------------------
// NFA (src/nfa/nfa_node_m.F90, nfa_graph_m.F90)
// ---------------------------------------------------------------------------------------------
struct NfaTra {  // nfa_node_m.F90:35-41
    std::vector<Seg> c;
    bool has_c = false;
    int c_top = 0;
    int dst = NFA_NULL_TRANSITION;
};
struct NfaNode {  // :43-57 (forward part only)
    std::vector<NfaTra> forward;  // 1-based: forward[j-1]; slot forward_top is the next free one
    int forward_top = 1;
    NfaTra& fw(int j) {
        if ((int)forward.size() < j) forward.resize((size_t)j);
        return forward[(size_t)j - 1];
    }
};

struct Nfa {
    std::vector<NfaNode> nodes;  // 1-based
    int nfa_top = 0;
    int entry = 0, exit_ = 0;
    std::vector<Seg> all_segments;
    std::vector<std::vector<int> > eps;  // derived view: epsilon successors per state (closure speed only)

    NfaNode& at(int i) {
        if ((int)nodes.size() < i) nodes.resize((size_t)i);
        return nodes[(size_t)i - 1];
    }
    int make_node() { nfa_top++; at(nfa_top); return nfa_top; }

    // nfa__add_transition :324-372 (forward half)
    void add_transition(int src, int dst, const Seg& c) {
        NfaNode& self = at(src);
        int j = NFA_NULL_TRANSITION;
        if (!self.forward.empty() && c != SEG_EPSILON) {
            for (int jj = 1; jj <= self.forward_top; jj++) {
                NfaTra& t = self.fw(jj);
                if (dst == t.dst && t.c_top < NFA_C_SIZE) j = jj;
            }
        }
        if (j == NFA_NULL_TRANSITION) j = self.forward_top;
        NfaTra& t = self.fw(j);
        if (!t.has_c) { t.c.assign(NFA_C_SIZE, Seg()); t.has_c = true; }
        t.c_top++;
        t.c[(size_t)t.c_top - 1] = c;
        t.dst = dst;
        if (j == self.forward_top) { self.forward_top++; self.fw(self.forward_top); }
    }

    void generate(const Tree& tree, int idx, int entry_, int exit__) {  // :166-267
        if (idx == INVALID_INDEX) return;
        const TreeNode& n = tree.at(idx);
        int entry_local = entry_;
        switch (n.op) {
            case op_char:
                for (size_t k = 0; k < n.c.size(); k++) add_transition(entry_, exit__, n.c[k]);
                break;
            case op_empty:
                add_transition(entry_, exit__, SEG_EPSILON);
                break;
            case op_union:
                generate(tree, n.left_i, entry_, exit__);
                generate(tree, n.right_i, entry_, exit__);
                break;
            case op_closure:
                generate_closure(tree, idx, entry_, exit__);
                break;
            case op_concat: {  // :270-290
                int node1 = make_node();
                generate(tree, n.left_i, entry_, node1);
                generate(tree, n.right_i, node1, exit__);
                break;
            }
            case op_repeat: {  // :215-262
                int min_repeat = n.min_repeat, max_repeat = n.max_repeat;
                int num_1st = min_repeat - 1;
                if (max_repeat == INFINITE_) num_1st++;
                for (int j = 1; j <= num_1st; j++) {
                    int node1 = make_node();
                    generate(tree, n.left_i, entry_local, node1);
                    entry_local = node1;
                }
                int num_2nd = (min_repeat == 0) ? max_repeat - 1 : max_repeat - min_repeat;
                for (int j = 1; j <= num_2nd; j++) {
                    int node2 = make_node();
                    generate(tree, n.left_i, entry_local, node2);
                    add_transition(node2, exit__, SEG_EPSILON);
                    entry_local = node2;
                }
                if (min_repeat == 0) add_transition(entry_, exit__, SEG_EPSILON);
                if (max_repeat == INFINITE_) generate_closure(tree, idx, entry_local, exit__);
                else generate(tree, n.left_i, entry_local, exit__);
                break;
            }
            default:
                throw ErrorStop{-1003};  // "This will not happen in 'generate_nfa'."
        }
    }
    void generate_closure(const Tree& tree, int idx, int entry_, int exit__) {  // :292-322
        int node1 = make_node();
        int node2 = make_node();
        add_transition(entry_, node1, SEG_EPSILON);
        generate(tree, tree.at(idx).left_i, node1, node2);
        add_transition(node2, node1, SEG_EPSILON);
        add_transition(node1, exit__, SEG_EPSILON);
    }

    void build(const Tree& tree) {  // build_nfa_graph :61-106
        nodes.clear();
        nfa_top = 0;
        entry = make_node();
        exit_ = make_node();
        generate(tree, tree.top, entry, exit_);
        for (int i = 1; i <= nfa_top; i++) {  // nfa__merge_segments_of_transition :667-692
            NfaNode& nd = at(i);
            if (nd.forward.empty()) continue;
            for (int j = 1; j <= nd.forward_top; j++) {
                NfaTra& t = nd.fw(j);
                if (t.has_c) {
                    sort_segment_by_min(t.c);
                    merge_segments(t.c);
                    t.c_top = (int)t.c.size();
                }
            }
        }
        disjoin_nfa();
        eps.assign((size_t)nfa_top + 1, std::vector<int>());
        for (int i = 1; i <= nfa_top; i++) {
            NfaNode& nd = at(i);
            if (nd.forward.empty()) continue;
            for (int j = 1; j <= nd.forward_top; j++) {
                NfaTra& t = nd.fw(j);
                if (!t.has_c) continue;
                bool any_eps = false;
                for (auto& s : t.c) any_eps = any_eps || s == SEG_EPSILON;
                if (any_eps && t.dst != NFA_NULL_TRANSITION) eps[(size_t)i].push_back(t.dst);
            }
        }
    }

    void disjoin_nfa() {  // :410-501
        PQueue q;
        for (int i = 1; i <= nfa_top; i++) {
            NfaNode& nd = at(i);
            for (int j = 1; j <= nd.forward_top - 1; j++) {
                NfaTra& t = nd.fw(j);
                if (t.dst != NFA_NULL_TRANSITION)
                    for (int k = 1; k <= t.c_top; k++)
                        if (t.c[(size_t)k - 1] != SEG_INIT) q.enqueue(t.c[(size_t)k - 1]);
            }
        }
        long num_f = q.number;
        std::vector<Seg> seg_list((size_t)num_f);
        long m = 0;
        for (long j = 1; j <= num_f; j++) {
            if (j == 1) { m++; q.dequeue(seg_list[0]); continue; }
            Seg cache;
            q.dequeue(cache);
            if (seg_list[(size_t)m - 1] != cache) { m++; seg_list[(size_t)m - 1] = cache; }
        }
        seg_list.resize((size_t)m);
        disjoin(seg_list);
        for (int i = 1; i <= nfa_top; i++) {
            NfaNode& nd = at(i);
            if (nd.forward.empty()) continue;
            for (int j = 1; j <= nd.forward_top; j++) disjoin_each_transition(nd.fw(j), seg_list);
        }
        all_segments = seg_list;
    }

    static void disjoin_each_transition(NfaTra& t, const std::vector<Seg>& seg_list) {  // :508-572
        if (!t.has_c) return;
        std::vector<Seg> tmp;
        for (int k = 1; k <= t.c_top; k++)
            for (auto& s : seg_list)
                if (seg_in_seg(s, t.c[(size_t)k - 1])) tmp.push_back(s);  // is_overlap_to_seg_list
        long n = (long)tmp.size();
        if ((long)t.c.size() < n) t.c.assign((size_t)n, Seg());
        for (long k = 0; k < n; k++) t.c[(size_t)k] = tmp[(size_t)k];
        long k = 0;  // update_c_top :558-572
        while (k + 1 <= (long)t.c.size()) {
            k++;
            if (t.c[(size_t)k - 1] == SEG_INIT) break;
        }
        t.c_top = (int)k;
    }
};

// ---------------------------------------------------------------------------------------------
// automaton: lazy subset construction (src/automaton_m.F90, src/lazy_dfa/*)
// ---------------------------------------------------------------------------------------------
typedef std::vector<unsigned char> StateSet;  // nfa_state_set_t%vec, 1-based via [i-1]

struct DfaNode {
    StateSet nfa_set;
    bool accepted = false;
};

struct Automaton {
    Nfa nfa;
    std::vector<DfaNode> dfa;  // 1-based: dfa[i-1]; dfa_top = dfa.size()+1
    int dfa_limit = DFA_STATE_UNIT;
    int initial_index = -1;
    long steps = 0;  // number of construct() calls (reported by the baseline harness)

    void preprocess(const Tree& tree) { nfa.build(tree); }  // :53-62

    void mark_eps(StateSet& set, int idx) {  // nfa_graph_m.F90:76-104 / automaton_m.F90:121-151
        set[(size_t)idx - 1] = 1;
        for (int d : nfa.eps[(size_t)idx])
            if (!set[(size_t)d - 1]) mark_eps(set, d);
    }
    void collect_eps(StateSet& set) {  // nfa_graph_m.F90:107-123
        for (int i = 1; i <= nfa.nfa_top; i++)
            if (set[(size_t)i - 1]) mark_eps(set, i);
    }
    int registered(const StateSet& set) const {  // lazy_dfa_graph_m.F90:122-143
        for (size_t i = 0; i < dfa.size(); i++)
            if (dfa[i].nfa_set == set) return (int)i + 1;
        return DFA_INVALID_INDEX;
    }
    int register_state(const StateSet& set) {  // automaton_m.F90:156-190
        int i = registered(set);
        if (i != DFA_INVALID_INDEX) return i;
        int dfa_top = (int)dfa.size() + 1;
        if (dfa_top >= dfa_limit) {  // lazy_dfa__reallocate :68-101
            int siz = dfa_limit;
            if (siz * 2 > DFA_STATE_HARD_LIMIT) throw ErrorStop{ERRSTOP_DFA_LIMIT};
            dfa_limit = siz * 2;
        }
        DfaNode n;
        n.nfa_set = set;
        n.accepted = set[(size_t)nfa.exit_ - 1] != 0;
        dfa.push_back(n);
        return dfa_top;
    }
    void init() {  // :66-100
        dfa.clear();
        dfa_limit = DFA_STATE_UNIT;
        StateSet s((size_t)nfa.nfa_top, 0);
        mark_eps(s, nfa.entry);
        initial_index = register_state(s);
    }
    // automaton__compute_reachable_state :199-267
    StateSet get_reachable(int curr_i, const fstr& symbol) {
        StateSet out((size_t)nfa.nfa_top, 0);
        const StateSet& cur = dfa[(size_t)curr_i - 1].nfa_set;
        Seg sym = symbol_to_segment(symbol);
        for (int i = 1; i <= nfa.nfa_top; i++) {
            if (!cur[(size_t)i - 1]) continue;
            NfaNode& nd = nfa.nodes[(size_t)i - 1];
            if (nd.forward.empty()) continue;
            for (int j = 1; j <= nd.forward_top && j <= (int)nd.forward.size(); j++) {
                NfaTra& t = nd.forward[(size_t)j - 1];
                if (t.dst == NFA_NULL_TRANSITION) continue;
                if (t.c_top < 1) continue;  // `do k = 1, c_top` runs zero times
                bool hit = false;           // `symbol .in. segs` over the whole array (:248-251)
                for (auto& s : t.c) if (seg_in_seg(sym, s)) { hit = true; break; }
                if (hit) out[(size_t)t.dst - 1] = 1;
            }
        }
        return out;
    }
    // automaton__construct_dfa :333-381
    int construct(int curr_i, const fstr& symbol) {
        steps++;
        StateSet set = get_reachable(curr_i, symbol);
        collect_eps(set);
        bool any = false;
        for (unsigned char b : set) any = any || b;
        if (!any) return DFA_INVALID_INDEX;
        int dst = registered(set);
        if (dst == DFA_INVALID_INDEX) dst = register_state(set);
        return dst;
    }
    bool accepted(int i) const { return dfa[(size_t)i - 1].accepted; }
};

// ---------------------------------------------------------------------------------------------
// drivers (src/api_internal_m.F90) and prefilter (src/essential/utility_m.f90:58-117)
// ---------------------------------------------------------------------------------------------
static const fstr REPLACEMENT = "\xEF\xBF\xBF";  // make_replacement_char, utf8_m.f90:433-438

static bool get_index_list_forward(const fstr& text, const fstr& prefix, const fstr& suffix,
                                   std::vector<long>& index_array) {  // returns allocated(index_array)
    long len_pre = (long)prefix.size();
    if (len_pre == 0) return false;
    index_array.assign(32, INVALID_CHAR_INDEX);
    long idx = f_index(text, prefix);
    long suf_idx = f_index_back(text, suffix);
    if (suf_idx == 0) suf_idx = INVALID_CHAR_INDEX;
    if (idx <= 0) return true;
    else if (suf_idx != INVALID_CHAR_INDEX) { if (idx <= suf_idx) index_array[0] = idx; }
    else index_array[0] = idx;
    long offset = idx + len_pre - 1;
    long i = 2;
    while (offset < (long)text.size()) {
        idx = f_index(text.substr((size_t)offset), prefix);
        if (idx <= 0) break;
        if ((long)index_array.size() < i) index_array.resize((size_t)i, INVALID_CHAR_INDEX);
        index_array[(size_t)i - 1] = idx + offset;
        i++;
        if (i > (long)index_array.size()) index_array.resize(index_array.size() * 2, INVALID_CHAR_INDEX);
        offset = offset + idx + len_pre - 1;
        if (suf_idx != INVALID_CHAR_INDEX && offset > suf_idx) break;
    }
    return true;
}

// api_internal_m.F90:31-167
static void do_matching_including(Automaton& am, const fstr& string, long& from, long& to,
                                  const fstr& prefix, const fstr& suffix) {
    fstr str = fstr(1, '\0') + string + fstr(1, '\0');
    long lstr = (long)str.size();
    from = 0; to = 0;
    bool do_brute_force = f_eq(prefix, "");
    long suf_idx = INVALID_CHAR_INDEX;
    int cur_i = am.initial_index;
    if (string.size() <= 1 && f_eq(string, "")) {
        if (am.accepted(cur_i)) { from = ACCEPTED_EMPTY; to = ACCEPTED_EMPTY; }
        return;
    }
    std::vector<long> index_list;
    if (!do_brute_force) {
        if (!get_index_list_forward(str, prefix, suffix, index_list)) return;
        if (index_list[0] == INVALID_CHAR_INDEX) do_brute_force = true;
    }
    long i, start;
    if (do_brute_force) {
        i = 1; start = i;
    } else {
        if (index_list[0] == 2) { start = 1; i = 0; }
        else { i = 1; start = index_list[0]; }
        if (!f_eq(suffix, "")) {
            suf_idx = f_index_back(string, suffix);
            if (suf_idx == 0) return;
        }
    }
    while (start < lstr) {
        long max_match = 0;
        long ci = start;
        cur_i = am.initial_index;
        if (suf_idx != INVALID_CHAR_INDEX && suf_idx < ci) break;
        while (cur_i != DFA_INVALID_INDEX) {
            if (am.accepted(cur_i) && ci != start) max_match = ci;
            if (ci > lstr) break;
            long next_ci; bool valid;
            next_idxutf8_strict(str, ci, next_ci, valid);
            int dst_i = valid ? am.construct(cur_i, sub(str, ci, next_ci - 1)) : am.construct(cur_i, REPLACEMENT);
            cur_i = dst_i;
            ci = next_ci;
        }
        if (max_match > 0) {
            from = start - 1;
            if (from == 0) from = 1;
            if (max_match >= lstr) to = (long)string.size();
            else to = max_match - 2;
            return;
        }
        if (do_brute_force) {
            bool valid; long nx;
            next_idxutf8_strict(str, start, nx, valid);
            start = nx;
            continue;
        }
        i++;
        if (i <= (long)index_list.size()) {
            start = index_list[(size_t)i - 1];
            if (start == INVALID_CHAR_INDEX) return;
        } else return;
    }
}

// api_internal_m.F90:171-303
static bool do_matching_exactly(Automaton& am, const fstr& string, const fstr& prefix, const fstr& suffix) {
    long len_pre = (long)prefix.size(), len_suf = (long)suffix.size(), n = (long)string.size();
    bool matches_pre = true, matches_post = true;
    if (n > 0 && len_pre > 0)
        if (f_eq(prefix, string) && len_pre == n) return true;
    if (len_pre > n || len_suf > n) return false;
    bool empty_pre = f_eq(prefix, ""), empty_post = f_eq(suffix, "");
    if (n > 0) {
        if (!empty_pre) matches_pre = f_eq(sub(string, 1, len_pre), prefix);
        if (!empty_post) matches_post = f_eq(sub(string, n - len_suf + 1, n), suffix);
    } else {
        matches_pre = len_pre == 0;
        matches_post = len_suf == 0;
    }
    bool runs_engine = (empty_pre || matches_pre) && (empty_post || matches_post);
    if (!runs_engine) return false;
    int cur_i = am.initial_index;
    if (n == 0) return am.accepted(cur_i);
    long max_match = 0, ci = 1;
    fstr str = fstr(1, '\0') + string + fstr(1, '\0');
    long lstr = (long)str.size();
    while (cur_i != DFA_INVALID_INDEX) {
        if (am.accepted(cur_i)) max_match = ci;
        if (ci > lstr) break;
        long next_ci; bool valid;
        next_idxutf8_strict(str, ci, next_ci, valid);
        int dst_i = valid ? am.construct(cur_i, sub(str, ci, next_ci - 1)) : am.construct(cur_i, REPLACEMENT);
        if (dst_i == DFA_INVALID_INDEX && ci == 1) {
            ci = 2;
            next_idxutf8_strict(str, ci, next_ci, valid);
            dst_i = valid ? am.construct(cur_i, sub(str, ci, next_ci - 1)) : am.construct(cur_i, REPLACEMENT);
        }
        cur_i = dst_i;
        ci = next_ci;
    }
    return max_match >= n + 2;
}

// ---------------------------------------------------------------------------------------------
// public API wrappers (src/forgex.F90)
// ---------------------------------------------------------------------------------------------
// A compiled pattern: everything the wrappers compute before they touch the text.  The reference
// rebuilds this on every call (src/forgex.F90:98, :139-140); keeping it is result-neutral because
// the lazy DFA is a deterministic function of the NFA.
struct Compiled {
    int mode = 0;  // 0: .in. / regex preprocessing (trim), 1: .match. preprocessing
    Tree tree;
    fstr all, prefix, suffix;
    Automaton am;
    bool has_automaton = false;
    int status = SYNTAX_VALID;
};

// src/essential/utility_m.f90:23-52
static bool is_there_caret_at_the_top(const fstr& pattern) {
    fstr buff = f_adjustl(pattern);
    if (buff.empty()) return false;
    return buff[0] == '^';
}
static bool is_there_dollar_at_the_end(const fstr& pattern) {
    fstr buff = f_trim(pattern);
    if (buff.empty()) return false;
    return buff[buff.size() - 1] == '$';
}

static void compile(Compiled& c, const fstr& pattern, int mode) {
    c.mode = mode;
    fstr buff;
    if (mode == 0) {
        buff = f_trim(pattern);                                   // forgex.F90:95, :260
    } else {                                                      // forgex.F90:182-190
        if (is_there_caret_at_the_top(pattern)) buff = sub(pattern, 2, (long)pattern.size());
        else buff = pattern;
        if (is_there_dollar_at_the_end(pattern)) buff = sub(buff, 1, f_len_trim(pattern) - 1);
    }
    c.tree.build(buff);
    c.status = c.tree.is_valid ? SYNTAX_VALID : c.tree.code;
    if (!c.tree.is_valid) return;
    extract_literal(c.tree, c.all, c.prefix, c.suffix);
}
static void ensure_automaton(Compiled& c) {
    if (c.has_automaton) return;
    c.am.preprocess(c.tree);
    c.am.init();
    c.has_automaton = true;
}

// operator__in (forgex.F90:74-160)
static bool api_in(Compiled& c, const fstr& str) {
    if (!c.tree.is_valid) return false;
    if (!f_eq(c.all, "")) {
        long from = f_index(str, c.all), to = INVALID_CHAR_INDEX;
        if (from > 0) to = from + (long)c.all.size() - 1;
        return from > 0 && to > 0;
    }
    ensure_automaton(c);
    long from, to;
    do_matching_including(c.am, str, from, to, c.prefix, c.suffix);
    if (from == ACCEPTED_EMPTY && to == ACCEPTED_EMPTY) return true;
    return from > 0 && to > 0;
}

// operator__match (forgex.F90:163-231)
static bool api_match(Compiled& c, const fstr& str) {
    if (!c.tree.is_valid) return false;
    if (!f_eq(c.all, "")) {
        if (str.size() == c.all.size()) return f_eq(str, c.all);
    }
    ensure_automaton(c);
    return do_matching_exactly(c.am, str, c.prefix, c.suffix);
}

// subroutine__regex (forgex.F90:235-347): res = text(from:to)
static void api_regex(Compiled& c, const fstr& text, long& from, long& to, long& length, int& status) {
    status = SYNTAX_VALID;
    if (!c.tree.is_valid) {
        length = 0; from = INVALID_CHAR_INDEX; to = INVALID_CHAR_INDEX; status = c.tree.code;
        return;
    }
    if (!f_eq(c.all, "")) {
        long from_l = f_index(text, c.all), to_l = INVALID_CHAR_INDEX;
        if (from_l > 0) to_l = from_l + (long)c.all.size() - 1;
        if (from_l > 0 && to_l > 0) { from = from_l; to = to_l; length = (long)c.all.size(); }
        else { from = 0; to = 0; length = 0; }
        return;
    }
    ensure_automaton(c);
    long from_l, to_l;
    do_matching_including(c.am, text, from_l, to_l, c.prefix, c.suffix);
    if (from_l == ACCEPTED_EMPTY && to_l == ACCEPTED_EMPTY) { from = 0; to = 0; length = 0; return; }
    if (from_l > 0 && to_l > 0) { length = to_l - from_l + 1; from = from_l; to = to_l; }
    else { length = 0; from = 0; to = 0; }
}

}  // namespace fxo

// =============================================================================================
// C interface for tests/ and bench.py (ctypes).  All functions return 0 / a result >= 0 on success
// and the negative ERRSTOP_* code where the reference would `error stop`.
// =============================================================================================
using namespace fxo;

#define GUARD_BEGIN try {
#define GUARD_END } catch (const ErrorStop& e) { return e.code; } catch (const std::bad_alloc&) { return -1999; }

extern "C" {

const char* fxo_error_message(int code) { return error_message(code); }

// is_valid_regex (forgex.F90:58-71); returns 1/0, status code in *status
int fxo_is_valid(const char* pattern, long plen, int* status) {
    GUARD_BEGIN {
        Tree t;
        t.build(f_trim(fstr(pattern, (size_t)plen)));
        if (status) *status = t.is_valid ? SYNTAX_VALID : t.code;
        return t.is_valid ? 1 : 0;
    } GUARD_END
}

// extract_literal on tree%build(pattern) with the pattern exactly as given (src/test_m.F90:103-146)
int fxo_literals(const char* pattern, long plen, char* all, long* all_len, char* prefix, long* prefix_len,
                 char* suffix, long* suffix_len, long cap) {
    GUARD_BEGIN {
        Tree t;
        t.build(fstr(pattern, (size_t)plen));
        if (!t.is_valid) return 1;
        fstr a, p, s;
        extract_literal(t, a, p, s);
        if ((long)a.size() > cap || (long)p.size() > cap || (long)s.size() > cap) return 2;
        memcpy(all, a.data(), a.size()); *all_len = (long)a.size();
        memcpy(prefix, p.data(), p.size()); *prefix_len = (long)p.size();
        memcpy(suffix, s.data(), s.size()); *suffix_len = (long)s.size();
        return 0;
    } GUARD_END
}

// one-shot calls: recompile per call exactly like the reference API does
int fxo_in(const char* pattern, long plen, const char* text, long tlen) {
    GUARD_BEGIN {
        Compiled c;
        compile(c, fstr(pattern, (size_t)plen), 0);
        return api_in(c, fstr(text, (size_t)tlen)) ? 1 : 0;
    } GUARD_END
}
int fxo_match(const char* pattern, long plen, const char* text, long tlen) {
    GUARD_BEGIN {
        Compiled c;
        compile(c, fstr(pattern, (size_t)plen), 1);
        return api_match(c, fstr(text, (size_t)tlen)) ? 1 : 0;
    } GUARD_END
}
int fxo_regex(const char* pattern, long plen, const char* text, long tlen, long* from, long* to, long* length,
              int* status) {
    GUARD_BEGIN {
        Compiled c;
        compile(c, fstr(pattern, (size_t)plen), 0);
        api_regex(c, fstr(text, (size_t)tlen), *from, *to, *length, *status);
        return 0;
    } GUARD_END
}

// compiled handles for batches and for the CPU baseline ("pattern compiled once per process")
void* fxo_compile(const char* pattern, long plen, int mode, int* status) {
    Compiled* c = new Compiled();
    try {
        compile(*c, fstr(pattern, (size_t)plen), mode);
        if (status) *status = c->status;
    } catch (const ErrorStop& e) {
        if (status) *status = e.code;
        delete c;
        return nullptr;
    }
    return c;
}
void fxo_free(void* h) { delete (Compiled*)h; }
long fxo_steps(void* h) { return ((Compiled*)h)->am.steps; }
long fxo_dfa_states(void* h) { return (long)((Compiled*)h)->am.dfa.size(); }

// op: 0 = .in., 1 = .match. (handle must have been compiled with the matching mode)
int fxo_bool_batch(void* h, int op, const char* buf, const long long* offsets, long n, unsigned char* out) {
    Compiled& c = *(Compiled*)h;
    GUARD_BEGIN {
        for (long i = 0; i < n; i++) {
            fstr s(buf + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
            out[i] = (op == 0 ? api_in(c, s) : api_match(c, s)) ? 1 : 0;
        }
        return 0;
    } GUARD_END
}
int fxo_bool_fixed(void* h, int op, const char* buf, long n, long stride, unsigned char* out) {
    Compiled& c = *(Compiled*)h;
    GUARD_BEGIN {
        for (long i = 0; i < n; i++) {
            fstr s(buf + i * stride, (size_t)stride);
            out[i] = (op == 0 ? api_in(c, s) : api_match(c, s)) ? 1 : 0;
        }
        return 0;
    } GUARD_END
}
int fxo_regex_batch(void* h, const char* buf, const long long* offsets, long n, long long* from, long long* to) {
    Compiled& c = *(Compiled*)h;
    GUARD_BEGIN {
        for (long i = 0; i < n; i++) {
            fstr s(buf + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
            long f, t, l; int st;
            api_regex(c, s, f, t, l, st);
            from[i] = f; to[i] = t;
        }
        return 0;
    } GUARD_END
}
// one buffer; 64-bit indices (the Fortran original is limited to default integers)
int fxo_regex_buffer(void* h, const char* buf, long long len, long long* from, long long* to) {
    Compiled& c = *(Compiled*)h;
    GUARD_BEGIN {
        long f, t, l; int st;
        api_regex(c, fstr(buf, (size_t)len), f, t, l, st);
        *from = f; *to = t;
        return 0;
    } GUARD_END
}

}  // extern "C"
