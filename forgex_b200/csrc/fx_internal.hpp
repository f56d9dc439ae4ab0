// fx_internal.hpp -- host-side pattern compiler of forgex_b200 (product code).
//
// Pipeline:  pattern bytes --(fx_front.cpp)--> syntax tree + literals
//                          --(fx_automata.cpp)--> Thompson NFA --> eager subset automata
//                          --> byte-level DFA over UTF-8 bytes (invalid-byte rule compiled in)
//                          --> flat device tables (fx_cabi.cu uploads them and launches kernels)
//
// The language recognised must equal Forgex's (reference: /root/reference/src/ast, src/nfa,
// src/automaton_m.F90); every rule cites the reference lines that define it.  Nothing here
// includes, links or calls oracle/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fx {

// Forgex status codes (reference: src/essential/error_m.F90:12-38), plus B200-side conditions.
enum Status : int {
    OK = 0,
    ERR_GENERIC = 1,
    ERR_PAREN_MISSING = 2,
    ERR_PAREN_UNEXPECTED = 3,
    ERR_BRACKET_MISSING = 4,
    ERR_BRACKET_UNEXPECTED = 5,
    ERR_BRACE_MISSING = 6,
    ERR_BRACE_UNEXPECTED = 7,
    ERR_INVALID_TIMES = 8,
    ERR_ESCAPE_MISSING = 9,
    ERR_ESCAPE_INVALID = 10,
    ERR_EMPTY_CLASS = 11,
    ERR_RANGE_WITH_ESCAPE = 12,
    ERR_MISPLACED_SUBTRACTION = 13,
    ERR_INVALID_RANGE = 14,
    ERR_CLASS_SUBTRACTION = 15,
    ERR_STAR_INCOMPLETE = 16,
    ERR_PLUS_INCOMPLETE = 17,
    ERR_QUESTION_INCOMPLETE = 18,
    ERR_INVALID_HEX = 19,
    ERR_HEX_DIGITS = 20,
    ERR_UNICODE_EXCEED = 21,
    ERR_UNICODE_PROPERTY = 22,
    ERR_SHOULD_NOT_HAPPEN = 23,
    ERR_ALLOCATION = 24,
    // conditions where the reference aborts with `error stop`, or that only exist on this side
    ERR_TREE_NODE_LIMIT = 101,     // > 2048 syntax-tree nodes (src/essential/parameters_m.f90:21-25)
    ERR_DFA_STATE_CAP = 102,       // eager automaton exceeds the state cap (reference cap: parameters_m.f90:126-130)
    ERR_PREFILTER_UNSUPPORTED = 103,  // literal prefilter of this pattern is not provably result-neutral
    ERR_BAD_ARGUMENT = 104,
    ERR_NO_DEVICE = 105,
    ERR_WORK_BUDGET = 106,
};

const char* status_message(int code);

struct Range {  // inclusive code-point range
    int lo, hi;
};

enum NodeOp { N_CHAR = 1, N_CONCAT, N_UNION, N_CLOSURE, N_REPEAT, N_EMPTY };

struct Node {
    int op = 0;
    std::vector<Range> set;  // N_CHAR: the segment list exactly as the reference stores it (may hold a sentinel)
    int left = -1, right = -1;
    int rmin = 0, rmax = 0;  // N_REPEAT (rmax == REPEAT_INF for `{m,}`)
};

static const int REPEAT_INF = -9998;
static const int CP_MAX = 0x10FFFF;
static const int CP_TOP = 0x1FFFFF;  // largest value a structurally valid 4-byte sequence decodes to
static const int CP_SENTINEL = CP_MAX + 2;  // default-initialised segment (segment_m.F90:38-46)

struct Syntax {
    std::vector<Node> nodes;
    int root = -1;
    int status = OK;
    bool valid() const { return status == OK; }
};

struct Literals {
    std::string all, prefix, suffix;
};

// fx_front.cpp
void parse_pattern(const std::string& pattern, Syntax& out);
void extract_literals(const Syntax& syn, Literals& lit);
std::string prepare_pattern(const std::string& pattern, bool match_mode);
bool fortran_blank(const std::string& s);  // s == '' under Fortran blank-padded comparison

// ---------------------------------------------------------------------------------------------
// automata
// ---------------------------------------------------------------------------------------------
struct Nfa {
    int n = 0;  // states are 1..n ; entry = 1, exit = 2
    int entry = 1, exit = 2;
    std::vector<std::vector<int> > eps;                             // eps[s] -> successors
    std::vector<std::vector<std::pair<Range, int> > > edges;        // edges[s] -> (range, dst)
    std::vector<int> cuts;  // alphabet: sorted code points where a class starts; last entry CP_TOP+1
};

// A deterministic automaton over code-point classes, produced by subset construction.
struct CpAutomaton {
    int nstates = 0;                  // state 0 = dead (never accepting, absorbs)
    int nclasses = 0;                 // classes of the code-point alphabet (class of cp: see class_of)
    std::vector<int> cuts;            // class c covers [cuts[c], cuts[c+1]-1]
    std::vector<int> delta;           // nstates x nclasses
    std::vector<uint8_t> accept;      // exit in set
    std::vector<uint8_t> end_accept;  // meaning depends on the mode, see build_* below
    int start = 0;                    // start state for text byte 0 (leading NUL already applied where the mode says so)
    int start_nul = 0;                // REGEX mode: state after the leading NUL from q0 (start position 1)
    int q0 = 0;                       // REGEX mode: state for starts >= 2
    int matched = -1;                 // absorbing "match found" state (IN mode), else -1
    bool q0_accepting = false;        // exit in closure(entry)  (blank-text rule, api_internal_m.F90:68-74)
    int class_of(int cp) const;
};

enum Mode { MODE_MATCH = 0, MODE_IN = 1, MODE_REGEX = 2, MODE_SPAN_FWD = 3 };

// Reverse automaton over the same code-point classes: determinised reversal of the NFA.  Walking the text
// backwards from the end of a match, state X = NFA states from which the exit is reachable through the symbols
// read so far; a position is a valid match start when X holds the entry (startok).
struct RevAutomaton {
    int nstates = 0;               // state 0 = dead
    int nclasses = 0;
    int start = 0;                 // reverse closure of {exit}
    std::vector<int> cuts;
    std::vector<uint16_t> delta;   // nstates x nclasses
    std::vector<uint8_t> startok;  // nstates
    // device form: delta16[r * nclasses + c] = next state | 0x8000 if the next state holds the NFA entry (<= 32767 states),
    // and a two-level class map of the code points below U+10000 so that the backward walk finds the class of a decoded
    // character with one or two shared-memory loads instead of a binary search over `cuts`:
    //   page[cp >> 6] < 0x80 : the class of all 64 code points of the block;  >= 0x80 : mixed[(page & 0x7F) * 64 + (cp & 63)]
    std::vector<uint16_t> delta16;
    std::vector<uint8_t> page;     // 1024 entries; empty when the automaton has more than 127 classes or mixed blocks
    std::vector<uint8_t> mixed;
};
int build_rev_automaton(const Nfa& nfa, int state_cap, RevAutomaton& out);

int build_nfa(const Syntax& syn, Nfa& nfa);
int build_cp_automaton(const Nfa& nfa, Mode mode, int state_cap, CpAutomaton& out);

// Flat byte-level tables, the thing the kernels walk.
//   next = table[(state << row_shift) + cls]   with cls = classmap[byte]
//   REGEX tables (flag_bits): word = state id (bits 0..13) | INTER (bit 14) | ACC (bit 15), <= 16383 states
//   MATCH / IN tables:        word = state id (16 bits), <= 65535 states; flags[] is read once, at the end
static const uint16_t W_ACC = 0x8000, W_INTER = 0x4000, W_STATE = 0x3FFF;
//   SPAN tables (span_words; the forward "ordered groups" automaton of the linear-time span path and of the long-buffer
//   state-map scan): word = state id (bits 0..11, <= 4095 states) | RA (bits 12..13) | INTER (bit 14) | ACC (bit 15).
//   RA != 0 marks a transition out of an in-sequence state on a byte that breaks the sequence: the pending bytes replay
//   as U+FFFF (api_internal_m.F90:129-133) and the last accepting boundary passed on the way lies RA-1 bytes before the
//   byte just read -- so a walker needs neither the start of the sequence nor a look at the flags:
//       if (w & W_RA) last = j + 1 - RA;   if (w & W_ACC) last = j + 1;        (j = index of the byte just read)
static const uint16_t W_SSTATE = 0x0FFF, W_RA = 0x3000;
static const int W_RA_SHIFT = 12;
// endinfo[state] of a SPAN table: what the end of the text does in this state.  bits 0..1 = RA of the pending bytes'
// replay (last = len + 1 - RA), bit 2 = the trailing NUL is consumed into an accept (last = len + 1)
enum EndInfo : uint8_t { EI_RA = 3, EI_NUL = 4 };
enum StateFlag : uint8_t {
    SF_ACC = 1,        // boundary state whose NFA set holds the exit
    SF_END = 2,        // "result is true if the text ends in this state" (MATCH / IN modes)
    SF_INTER = 4,      // inside a multi-byte sequence
    SF_MATCHED = 8,    // absorbing match state (IN mode)
    SF_FAILACC1 = 16,  // INTER only: replaying the pending bytes as U+FFFF passes an accepting state
    SF_FAILACC2 = 32,  //   after the 1st / 2nd / 3rd replayed byte (REGEX mode bookkeeping)
    SF_FAILACC3 = 64,
};

struct ByteTable {
    int nstates = 0;     // boundary states first (same ids as the CpAutomaton), then INTER states
    int nboundary = 0;
    int nclasses = 0;    // byte classes
    int row_shift = 0;   // row stride = 1 << row_shift entries
    uint8_t classmap[256];
    std::vector<uint16_t> table;   // nstates << row_shift
    std::vector<uint8_t> flags;    // nstates
    int start = 0, start_nul = 0, q0 = 0, matched = -1;
    bool q0_accepting = false;
    bool flag_bits = false;
    // direct form: 256 columns, no classmap lookup (filled when it fits, see fx_cabi.cu)
    std::vector<uint16_t> direct;  // nstates * 256
    // boolean tables only: result states are numbered last (state >= result_threshold <=> SF_END|SF_MATCHED),
    // and with <= 255 states a one-byte-per-entry 256-column table exists
    int result_threshold = 0;
    std::vector<uint8_t> direct8;  // nstates * 256, or empty
    // SPAN tables only (see W_RA above)
    bool span_words = false;
    std::vector<uint8_t> endinfo;  // nstates
};

int build_byte_table(const CpAutomaton& a, bool flag_bits, ByteTable& out, bool span_words = false);

// NFA simulation tables: the engine of patterns whose EAGER automaton passes the state cap.  Forgex builds its DFA
// lazily and only ever aborts on the number of states a text makes it VISIT (src/lazy_dfa/lazy_dfa_graph_m.F90:90-92);
// `.*a(a|b){500}c{20}` (test/test_api/test_case_005.f90:76-104) has an astronomically large DFA and a tiny visited one.
// Such patterns are matched on the device by the subset step itself, like the reference does per character
// (src/automaton_m.F90:199-381), on bit sets: trans[(s * nclasses + c) * words ..] = closure(move({s}, class c)).
struct NfaTables {
    int nstates = 0;     // bits 1..nstates are NFA states (bit 0 unused)
    int words = 0;       // 64-bit words per set
    int nclasses = 0;
    int exit = 2;
    std::vector<int> cuts;              // nclasses + 1
    std::vector<uint64_t> trans;        // (nstates + 1) x nclasses x words
    std::vector<uint64_t> q0;           // closure(entry)
    bool q0_accepting = false;
};
int build_nfa_tables(const Nfa& nfa, NfaTables& out);

// Whole compiled pattern (host side).
struct Program {
    int op = 0;
    int status = OK;
    std::string prepared;  // pattern after the entry point's own preprocessing
    Literals lit;
    bool literal_only = false;   // `all` is not blank: the reference never runs the automaton (.in./regex)
    bool prefix_active = false;  // a non-blank prefix exists (prefilter candidates, Q7b)
    int nfa_states = 0;
    CpAutomaton cp;
    ByteTable bt;
    // FX_OP_REGEX only: the linear-time span path (forward "ordered groups" automaton + reverse automaton);
    // absent (has_span == false) when a cap is exceeded or the pattern uses the prefix prefilter
    bool nfa_engine = false;     // the eager automaton passes the cap: matched by NFA simulation on the device (nfa_tables)
    NfaTables nfa_tables;
    bool has_span = false;          // ragged batches and the all-matches loop may use the span path (no prefix prefilter)
    bool has_span_tables = false;   // the span automata exist (also for a pattern with a prefix literal: the long-buffer
                                    // state-map scan uses them when the prefilter is provably result-neutral)
    CpAutomaton span_cp;
    ByteTable span_bt;
    RevAutomaton rev;
};

int compile_program(const std::string& pattern, int op, int state_cap, Program& out, bool want_span = true);

// an anchored code-point DFA explored by the host (the Fortran-side route, see compile_from_dfa in fx_automata.cpp)
struct DfaInput {
    const int32_t* cuts;      // ncls + 1
    const int32_t* delta;     // nstates x ncls, 0 = dead
    const uint8_t* accept;    // nstates
    int nstates, ncls, q0;
    std::string all, prefix, suffix;
};
int compile_from_dfa(const DfaInput& in, int op, int state_cap, Program& out, bool want_span = true);

}  // namespace fx
