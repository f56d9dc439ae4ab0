// fx_automata.cpp -- syntax tree -> NFA -> eagerly built automata -> byte-level DFA tables.
//
// Replaces, on the host, what the reference does lazily per input character
// (src/automaton_m.F90:199-381, src/lazy_dfa/*): every reachable subset state is built up
// front (with a stated cap), then the code-point automaton is turned into a DFA over raw bytes
// in which Forgex's text decoder (src/essential/utf8_m.f90:168-246, src/api_internal_m.F90:127-133:
// structural UTF-8, anything malformed = ONE byte presented as U+FFFF) is compiled into the
// transitions.  The result is the `.match.` / `.in.` / `regex` semantics of
// src/api_internal_m.F90:31-303 expressed as "walk a table, read a flag".
#include <algorithm>
#include <cstring>
#include <map>
#include <unordered_map>

#include "fx_internal.hpp"

namespace fx {

// ---------------------------------------------------------------------------------------------
// Thompson construction (reference: src/nfa/nfa_node_m.F90:166-322)
// ---------------------------------------------------------------------------------------------
namespace {

// The reference stores the outgoing transitions of a state in slots of at most 16 segments that
// share a destination (nfa_node_m.F90:324-372): a non-epsilon segment joins the LAST slot with
// the same destination that still has room -- even one created by an epsilon move -- otherwise
// it opens a new slot; an epsilon move always opens a new slot.  Each slot is later sorted and
// its touching segments fused (nfa_node_m.F90:667-692).  Because epsilon is the pseudo segment
// (-1,-1), it fuses with a segment that starts at U+0000; whether the epsilon move survives
// that depends on the alphabet split (see finish()).  The slots are kept only to reproduce this.
struct Slot {
    int dst;
    std::vector<Range> segs;
};
const int EPS = -1;
const int SLOT_CAP = 16;

struct Builder {
    const Syntax& syn;
    std::vector<std::vector<Slot> > out;  // out[s] : slots of state s (1-based states)
    int nstates = 0;

    explicit Builder(const Syntax& s) : syn(s) { out.resize(1); }

    int fresh() { out.emplace_back(); return ++nstates; }

    void arc(int src, int dst, Range seg) {
        std::vector<Slot>& slots = out[(size_t)src];
        int pick = -1;
        if (!(seg.lo == EPS && seg.hi == EPS))
            for (size_t j = 0; j < slots.size(); j++)
                if (slots[j].dst == dst && (int)slots[j].segs.size() < SLOT_CAP) pick = (int)j;
        if (pick < 0) { slots.push_back(Slot{dst, {}}); pick = (int)slots.size() - 1; }
        slots[(size_t)pick].segs.push_back(seg);
    }
    void eps(int src, int dst) { arc(src, dst, Range{EPS, EPS}); }

    void star(int operand, int from, int to) {  // nfa_node_m.F90:292-322
        int a = fresh(), b = fresh();
        eps(from, a);
        emit(operand, a, b);
        eps(b, a);
        eps(a, to);
    }

    void emit(int idx, int from, int to) {
        if (idx < 0) return;
        const Node& n = syn.nodes[(size_t)idx];
        switch (n.op) {
            case N_CHAR:
                for (const Range& r : n.set) arc(from, to, r);
                break;
            case N_EMPTY: eps(from, to); break;
            case N_UNION: emit(n.left, from, to); emit(n.right, from, to); break;
            case N_CLOSURE: star(n.left, from, to); break;
            case N_CONCAT: {
                int mid = fresh();
                emit(n.left, from, mid);
                emit(n.right, mid, to);
                break;
            }
            case N_REPEAT: {  // nfa_node_m.F90:215-262 (note the final unconditional copy)
                bool inf = n.rmax == REPEAT_INF;
                int cur = from;
                int mandatory = n.rmin - 1 + (inf ? 1 : 0);
                for (int j = 0; j < mandatory; j++) { int s = fresh(); emit(n.left, cur, s); cur = s; }
                int optional = n.rmin == 0 ? n.rmax - 1 : n.rmax - n.rmin;
                for (int j = 0; j < optional; j++) { int s = fresh(); emit(n.left, cur, s); eps(s, to); cur = s; }
                if (n.rmin == 0) eps(from, to);
                if (inf) star(n.left, cur, to);
                else emit(n.left, cur, to);
                break;
            }
            default: break;
        }
    }

    // sort + fuse one slot the way segment_m.F90:450-506 does (sentinel cuts the list)
    static void settle(std::vector<Range>& v) {
        std::stable_sort(v.begin(), v.end(), [](const Range& a, const Range& b) { return a.lo < b.lo; });
        size_t n = 1;
        while (n < v.size() && !(v[n].lo == CP_SENTINEL && v[n].hi == CP_SENTINEL)) n++;
        v.resize(std::min(n, v.size()));
        std::vector<Range> r;
        for (const Range& x : v) {
            if (!r.empty() && r.back().hi >= x.lo - 1) r.back().hi = std::max(r.back().hi, x.hi);
            else r.push_back(x);
        }
        v.swap(r);
    }

    void finish(Nfa& nfa) {
        for (auto& slots : out)
            for (auto& s : slots) settle(s.segs);
        // Does the reference's alphabet split keep -1 apart from 0?  (segment_disjoin_m.F90:98-163:
        // a piece ends at -1 when some segment starts at 0, when two distinct segments start at
        // -1, or when some segment ends at -1.)  Only then does a fused (-1..x) slot still act as
        // an epsilon move.
        bool starts_at_zero = false, ends_at_eps = false;
        std::vector<Range> starting_at_eps;
        for (auto& slots : out)
            for (auto& s : slots)
                for (const Range& r : s.segs) {
                    if (r.lo == CP_SENTINEL) continue;
                    if (r.lo == 0) starts_at_zero = true;
                    if (r.hi == EPS) ends_at_eps = true;
                    if (r.lo == EPS) {
                        bool seen = false;
                        for (const Range& q : starting_at_eps) seen = seen || (q.lo == r.lo && q.hi == r.hi);
                        if (!seen) starting_at_eps.push_back(r);
                    }
                }
        bool eps_splits = starts_at_zero || ends_at_eps || starting_at_eps.size() > 1;

        nfa.n = nstates;
        nfa.entry = 1;
        nfa.exit = 2;
        nfa.eps.assign((size_t)nstates + 1, {});
        nfa.edges.assign((size_t)nstates + 1, {});
        std::vector<int> cuts;
        cuts.push_back(0);
        cuts.push_back(CP_MAX + 1);
        cuts.push_back(CP_TOP + 1);  // 4-byte sequences decode up to 0x1FFFFF; nothing matches above U+10FFFF
        for (int s = 1; s <= nstates; s++)
            for (auto& slot : out[(size_t)s])
                for (Range r : slot.segs) {
                    if (r.lo == CP_SENTINEL) continue;
                    if (r.lo == EPS) {
                        if (r.hi == EPS || eps_splits) nfa.eps[(size_t)s].push_back(slot.dst);
                        if (r.hi == EPS) continue;
                        r.lo = 0;
                    }
                    if (r.hi > CP_MAX) r.hi = CP_MAX;
                    if (r.lo > r.hi) continue;
                    nfa.edges[(size_t)s].push_back({r, slot.dst});
                    cuts.push_back(r.lo);
                    cuts.push_back(r.hi + 1);
                }
        std::sort(cuts.begin(), cuts.end());
        cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
        nfa.cuts = cuts;
    }
};

}  // namespace

int build_nfa(const Syntax& syn, Nfa& nfa) {
    Builder b(syn);
    int entry = b.fresh(), exit = b.fresh();
    b.emit(syn.root, entry, exit);
    b.finish(nfa);
    return OK;
}

// ---------------------------------------------------------------------------------------------
// eager subset construction
// ---------------------------------------------------------------------------------------------
namespace {

typedef std::vector<uint64_t> Bits;

struct BitsHash {
    size_t operator()(const Bits& b) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (uint64_t w : b) { h ^= w + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

struct Subsets {
    const Nfa& nfa;
    size_t words;
    std::vector<std::vector<std::pair<int, int> > > by_class;  // per state: (class, dst) expanded from ranges
    Bits entry_closure;

    explicit Subsets(const Nfa& n) : nfa(n), words(((size_t)n.n + 64) / 64) {
        int ncls = (int)nfa.cuts.size() - 1;
        by_class.assign((size_t)nfa.n + 1, {});
        for (int s = 1; s <= nfa.n; s++)
            for (auto& e : nfa.edges[(size_t)s]) {
                int c0 = (int)(std::lower_bound(nfa.cuts.begin(), nfa.cuts.end(), e.first.lo) - nfa.cuts.begin());
                for (int c = c0; c < ncls && nfa.cuts[(size_t)c] <= e.first.hi; c++) by_class[(size_t)s].push_back({c, e.second});
            }
        entry_closure = Bits(words, 0);
        close_from(entry_closure, nfa.entry);
    }
    static bool has(const Bits& b, int s) { return (b[(size_t)s >> 6] >> (s & 63)) & 1; }
    static void put(Bits& b, int s) { b[(size_t)s >> 6] |= 1ull << (s & 63); }
    bool any(const Bits& b) const { for (uint64_t w : b) if (w) return true; return false; }

    void close_from(Bits& b, int s) const {  // epsilon closure, iterative
        std::vector<int> stack;
        if (!has(b, s)) { put(b, s); }
        stack.push_back(s);
        while (!stack.empty()) {
            int u = stack.back();
            stack.pop_back();
            for (int v : nfa.eps[(size_t)u])
                if (!has(b, v)) { put(b, v); stack.push_back(v); }
        }
    }
    // all successors of `from` (plus the entry closure when inject) on every class at once
    void step_all(const Bits& from, bool inject, std::vector<Bits>& next) const {
        int ncls = (int)nfa.cuts.size() - 1;
        next.assign((size_t)ncls, Bits(words, 0));
        for (int s = 1; s <= nfa.n; s++) {
            if (!(has(from, s) || (inject && has(entry_closure, s)))) continue;
            for (auto& cd : by_class[(size_t)s])
                if (!has(next[(size_t)cd.first], cd.second)) close_from(next[(size_t)cd.first], cd.second);
        }
    }
};

}  // namespace

int CpAutomaton::class_of(int cp) const {
    if (cp < 0 || cp > CP_TOP) return -1;
    return (int)(std::upper_bound(cuts.begin(), cuts.end(), cp) - cuts.begin()) - 1;
}

// Modes (all from the same NFA):
//  MATCH : anchored.  start = delta(q0, NUL) if alive else q0 (api_internal_m.F90:280-289);
//          end_accept[s] = exit in s  or  exit in delta(s, NUL)   (:261, :298)
//  IN    : search automaton: a new attempt is injected at every character boundary
//          (api_internal_m.F90:108-155), an accept after >= 1 symbol is absorbing ("matched");
//          start = state after the leading NUL; end_accept[s] = exit in delta_noinject(s, NUL)
//          (the trailing NUL is consumed by running attempts but is not a start, :108).
//          If the lone leading NUL is itself a match, the reference returns from that first
//          attempt (to = 0 unless it extends), so the automaton is then anchored at start 1.
//  REGEX : anchored, no absorption; accept[] per state; q0 and start_nul = delta(q0, NUL).
int build_cp_automaton(const Nfa& nfa, Mode mode, int state_cap, CpAutomaton& a) {
    Subsets ss(nfa);
    int ncls = (int)nfa.cuts.size() - 1;
    a = CpAutomaton();
    a.nclasses = ncls;
    a.cuts = nfa.cuts;
    int nul_class = a.class_of(0);
    a.q0_accepting = Subsets::has(ss.entry_closure, nfa.exit);

    std::unordered_map<Bits, int, BitsHash> index;
    std::vector<Bits> sets;
    std::vector<char> pinned;  // state whose own acceptance must not absorb (start states)
    auto intern = [&](const Bits& b) -> int {
        if (!ss.any(b) && mode != MODE_IN) return 0;
        auto it = index.find(b);
        if (it != index.end()) return it->second;
        int id = (int)sets.size();
        sets.push_back(b);
        index.emplace(b, id);
        return id;
    };
    sets.push_back(Bits(ss.words, 0));  // state 0: dead (MATCH/REGEX).  In IN mode the empty set is a live state, interned separately below.
    bool inject = false, absorb = false;
    Bits after_nul(ss.words, 0);
    {
        std::vector<Bits> nx;
        ss.step_all(Bits(ss.words, 0), true, nx);
        after_nul = nx[(size_t)nul_class];
    }
    std::vector<int> work;
    int matched = -1;
    if (mode == MODE_MATCH) {
        int q0 = intern(ss.entry_closure);
        int sn = intern(after_nul);
        a.start = sn != 0 ? sn : q0;
        a.q0 = q0;
        a.start_nul = sn;
    } else if (mode == MODE_REGEX) {
        a.q0 = intern(ss.entry_closure);
        a.start_nul = intern(after_nul);
        a.start = a.q0;
    } else {
        absorb = true;
        // sets[0] stays the unused dead row; a dedicated id marks "matched"
        matched = (int)sets.size();
        sets.push_back(Bits(ss.words, 0));  // placeholder row for the matched sink (never looked up by set)
        bool nul_is_match = Subsets::has(after_nul, nfa.exit);
        inject = !nul_is_match;
        // start state: its own acceptance does not count (the attempt must extend past the NUL)
        int id = (int)sets.size();
        sets.push_back(after_nul);
        if (!nul_is_match) index.emplace(after_nul, id);  // reusable only when it cannot be confused with a counted accept
        a.start = id;
        a.matched = matched;
    }
    // breadth-first closure over all classes
    std::vector<int> delta;
    auto grow = [&]() { delta.resize(sets.size() * (size_t)ncls, 0); };
    grow();
    std::vector<char> done;
    std::vector<Bits> nx;
    for (size_t cur = 0; cur < sets.size(); cur++) {
        if ((int)cur == 0 || (int)cur == matched) continue;
        ss.step_all(sets[cur], inject, nx);
        for (int c = 0; c < ncls; c++) {
            int dst;
            if (absorb && Subsets::has(nx[(size_t)c], nfa.exit)) dst = matched;
            else dst = intern(nx[(size_t)c]);
            if ((int)sets.size() > state_cap) return ERR_DFA_STATE_CAP;
            grow();
            delta[cur * (size_t)ncls + (size_t)c] = dst;
        }
    }
    if (matched >= 0)
        for (int c = 0; c < ncls; c++) delta[(size_t)matched * (size_t)ncls + (size_t)c] = matched;
    a.nstates = (int)sets.size();
    a.delta = delta;
    a.accept.assign((size_t)a.nstates, 0);
    a.end_accept.assign((size_t)a.nstates, 0);
    for (int s = 0; s < a.nstates; s++) {
        if (s == matched) { a.accept[(size_t)s] = 1; a.end_accept[(size_t)s] = 1; continue; }
        if (s == 0) continue;
        bool acc = Subsets::has(sets[(size_t)s], nfa.exit);
        a.accept[(size_t)s] = acc;
        if (mode == MODE_MATCH) {
            int t = delta[(size_t)s * (size_t)ncls + (size_t)nul_class];
            a.end_accept[(size_t)s] = acc || (t != 0 && Subsets::has(sets[(size_t)t], nfa.exit));
        } else if (mode == MODE_IN) {
            std::vector<Bits> t;
            ss.step_all(sets[(size_t)s], false, t);  // trailing NUL: consumed, but not a start
            a.end_accept[(size_t)s] = Subsets::has(t[(size_t)nul_class], nfa.exit);
            a.accept[(size_t)s] = 0;  // only the matched sink "accepts" in this mode
        } else {
            a.end_accept[(size_t)s] = acc;
        }
    }
    return OK;
}

// ---------------------------------------------------------------------------------------------
// spans in linear time: forward "ordered groups" automaton + reverse automaton
// ---------------------------------------------------------------------------------------------
// Forgex's search (api_internal_m.F90:108-155) tries the starts in order and returns at the first start whose
// anchored run accepts after >= 1 symbol, with that run's LAST accept as the end.  All runs can be advanced
// together if the automaton state remembers the ORDER of the runs: a state is a list of groups of NFA states, one
// group per start still in play, earliest start first.  Per symbol: a new group is injected for the start at this
// position (until something has matched), every group steps, an NFA state already held by an earlier group is
// dropped from later groups (same future, and the earlier start wins), empty groups vanish, and as soon as a group
// holds the exit every later group is cut off and injection stops ("matched").  Whenever the state holds the exit
// the kernel records the position: what it holds at the end is the last accept of the earliest start that ever
// accepted -- the end of Forgex's answer.  The start is then found by the reverse automaton walking backwards from
// that end: the leftmost position at which it holds the NFA entry.
namespace {

struct GroupState {
    bool matched = false;
    std::vector<Bits> groups;
    std::vector<uint64_t> key(size_t words) const {
        std::vector<uint64_t> k;
        k.push_back(matched ? 1 : 0);
        k.push_back(groups.size());
        for (const Bits& g : groups) k.insert(k.end(), g.begin(), g.end());
        (void)words;
        return k;
    }
};

}  // namespace

static int build_span_forward(const Nfa& nfa, int state_cap, CpAutomaton& a) {
    Subsets ss(nfa);
    const int ncls = (int)nfa.cuts.size() - 1;
    a = CpAutomaton();
    a.nclasses = ncls;
    a.cuts = nfa.cuts;
    const int nul_class = a.class_of(0);
    a.q0_accepting = Subsets::has(ss.entry_closure, nfa.exit);
    const size_t W = ss.words;

    // one step of a single group on every class at once
    auto step_group = [&](const Bits& g, std::vector<Bits>& out) { ss.step_all(g, false, out); };
    std::vector<Bits> injected;                    // closure(move(closure(entry), c)) per class
    step_group(ss.entry_closure, injected);

    std::map<std::vector<uint64_t>, int> index;
    std::vector<GroupState> states;
    std::vector<uint8_t> accept;
    auto intern = [&](const GroupState& st, bool acc) -> int {
        if (st.groups.empty() && st.matched) return 0;           // nothing can happen any more
        auto k = st.key(W);
        k.push_back(acc ? 1 : 0);
        auto it = index.find(k);
        if (it != index.end()) return it->second;
        int id = (int)states.size();
        states.push_back(st);
        accept.push_back(acc ? 1 : 0);
        index.emplace(k, id);
        return id;
    };
    // successor of `st` on class c; *acc tells whether the new state holds the exit
    auto successor = [&](const GroupState& st, const std::vector<std::vector<Bits> >& stepped, int c, bool inject,
                         bool* acc) -> GroupState {
        GroupState n;
        n.matched = st.matched;
        Bits seen(W, 0);
        auto add = [&](Bits g) {
            bool any = false;
            for (size_t w = 0; w < W; w++) { g[w] &= ~seen[w]; seen[w] |= g[w]; any = any || g[w]; }
            if (any) n.groups.push_back(g);
        };
        for (size_t gi = 0; gi < st.groups.size(); gi++) add(stepped[gi][(size_t)c]);
        if (inject && !st.matched) add(injected[(size_t)c]);
        *acc = false;
        for (size_t gi = 0; gi < n.groups.size(); gi++)
            if (Subsets::has(n.groups[gi], nfa.exit)) {
                n.groups.resize(gi + 1);                          // later starts can never win any more
                n.matched = true;
                *acc = true;
                break;
            }
        return n;
    };

    states.push_back(GroupState());                               // state 0: dead
    accept.push_back(0);
    GroupState empty;                                             // before any symbol: no run, nothing matched
    std::vector<int> delta;
    std::vector<uint8_t> end_accept;
    // the start state = after the leading NUL (start position 1 injected)
    {
        std::vector<std::vector<Bits> > none;
        bool acc;
        GroupState s0 = successor(empty, none, nul_class, true, &acc);
        a.start = intern(s0, acc);
        if (a.start == 0) {                                       // cannot happen (unmatched states are alive)
            a.start = intern(empty, false);
        }
    }
    for (size_t cur = 1; cur < states.size(); cur++) {
        GroupState st = states[cur];                              // copy: `states` grows below
        std::vector<std::vector<Bits> > stepped(st.groups.size());
        for (size_t gi = 0; gi < st.groups.size(); gi++) step_group(st.groups[gi], stepped[gi]);
        delta.resize(states.size() * (size_t)ncls, 0);
        end_accept.resize(states.size(), 0);
        for (int c = 0; c < ncls; c++) {
            bool acc;
            GroupState n = successor(st, stepped, c, true, &acc);
            int dst = intern(n, acc);
            if ((int)states.size() > state_cap) return ERR_DFA_STATE_CAP;
            delta.resize(states.size() * (size_t)ncls, 0);
            delta[cur * (size_t)ncls + (size_t)c] = dst;
        }
        bool eacc;
        successor(st, stepped, nul_class, false, &eacc);          // the trailing NUL is consumed but is not a start
        end_accept.resize(states.size(), 0);
        end_accept[cur] = eacc ? 1 : 0;
    }
    a.nstates = (int)states.size();
    delta.resize((size_t)a.nstates * (size_t)ncls, 0);
    end_accept.resize((size_t)a.nstates, 0);
    a.delta = delta;
    a.accept = accept;
    a.end_accept = end_accept;
    a.q0 = a.start;
    a.start_nul = a.start;
    a.matched = -1;
    return OK;
}

int build_rev_automaton(const Nfa& nfa, int state_cap, RevAutomaton& r) {
    const int ncls = (int)nfa.cuts.size() - 1;
    const size_t W = ((size_t)nfa.n + 64) / 64;
    // reversed adjacency
    std::vector<std::vector<int> > reps((size_t)nfa.n + 1);
    std::vector<std::vector<std::pair<int, int> > > redges((size_t)nfa.n + 1);   // dst -> (class, src)
    for (int s = 1; s <= nfa.n; s++) {
        for (int d : nfa.eps[(size_t)s]) reps[(size_t)d].push_back(s);
        for (auto& e : nfa.edges[(size_t)s]) {
            int c0 = (int)(std::lower_bound(nfa.cuts.begin(), nfa.cuts.end(), e.first.lo) - nfa.cuts.begin());
            for (int c = c0; c < ncls && nfa.cuts[(size_t)c] <= e.first.hi; c++) redges[(size_t)e.second].push_back({c, s});
        }
    }
    auto has = [](const Bits& b, int s) { return (b[(size_t)s >> 6] >> (s & 63)) & 1; };
    auto put = [](Bits& b, int s) { b[(size_t)s >> 6] |= 1ull << (s & 63); };
    auto close = [&](Bits& b) {
        std::vector<int> stack;
        for (int s = 1; s <= nfa.n; s++) if (has(b, s)) stack.push_back(s);
        while (!stack.empty()) {
            int u = stack.back();
            stack.pop_back();
            for (int v : reps[(size_t)u]) if (!has(b, v)) { put(b, v); stack.push_back(v); }
        }
    };
    r = RevAutomaton();
    r.nclasses = ncls;
    r.cuts = nfa.cuts;
    std::unordered_map<Bits, int, BitsHash> index;
    std::vector<Bits> sets;
    sets.push_back(Bits(W, 0));
    auto intern = [&](const Bits& b) -> int {
        bool any = false;
        for (uint64_t w : b) any = any || w;
        if (!any) return 0;
        auto it = index.find(b);
        if (it != index.end()) return it->second;
        int id = (int)sets.size();
        sets.push_back(b);
        index.emplace(b, id);
        return id;
    };
    Bits s0(W, 0);
    put(s0, nfa.exit);
    close(s0);
    r.start = intern(s0);
    std::vector<uint16_t> delta;
    for (size_t cur = 1; cur < sets.size(); cur++) {
        std::vector<Bits> next((size_t)ncls, Bits(W, 0));
        Bits x = sets[cur];
        for (int d = 1; d <= nfa.n; d++)
            if (has(x, d))
                for (auto& cs : redges[(size_t)d]) put(next[(size_t)cs.first], cs.second);
        delta.resize(sets.size() * (size_t)ncls, 0);
        for (int c = 0; c < ncls; c++) {
            close(next[(size_t)c]);
            int dst = intern(next[(size_t)c]);
            if ((int)sets.size() > state_cap) return ERR_DFA_STATE_CAP;
            delta.resize(sets.size() * (size_t)ncls, 0);
            delta[cur * (size_t)ncls + (size_t)c] = (uint16_t)dst;
        }
    }
    r.nstates = (int)sets.size();
    delta.resize((size_t)r.nstates * (size_t)ncls, 0);
    r.delta = delta;
    r.startok.assign((size_t)r.nstates, 0);
    for (int s = 1; s < r.nstates; s++) r.startok[(size_t)s] = has(sets[(size_t)s], nfa.entry) ? 1 : 0;
    if (r.nstates > 0x7FFF) return ERR_DFA_STATE_CAP;
    r.delta16.assign(r.delta.size(), 0);
    for (size_t i = 0; i < r.delta.size(); i++)
        r.delta16[i] = (uint16_t)(r.delta[i] | (r.startok[r.delta[i]] ? 0x8000 : 0));
    // two-level class map of the BMP
    r.page.clear();
    r.mixed.clear();
    if (ncls <= 127) {
        auto class_of = [&](int cp) { return (int)(std::upper_bound(r.cuts.begin(), r.cuts.end(), cp) - r.cuts.begin()) - 1; };
        std::vector<uint8_t> page(1024, 0), mixed;
        bool ok = true;
        for (int k = 0; k < 1024 && ok; k++) {
            const int lo = class_of(k << 6), hi = class_of((k << 6) + 63);
            if (lo == hi) { page[(size_t)k] = (uint8_t)lo; continue; }
            const size_t idx = mixed.size() / 64;
            if (idx > 127) { ok = false; break; }
            page[(size_t)k] = (uint8_t)(0x80 | idx);
            for (int x = 0; x < 64; x++) mixed.push_back((uint8_t)class_of((k << 6) + x));
        }
        if (ok) { r.page.swap(page); r.mixed.swap(mixed); }
    }
    return OK;
}

// ---------------------------------------------------------------------------------------------
// byte-level DFA
// ---------------------------------------------------------------------------------------------
// Boundary states are the code-point automaton's states.  A lead byte moves to an INTER state
// that remembers (origin state, sequence length, bytes still missing, which code-point window
// is still possible); the last continuation byte lands on the destination boundary state.  If a
// byte that is not 10xxxxxx arrives inside a sequence, the reference would have presented each
// pending byte as U+FFFF (api_internal_m.F90:129-133; utf8_m.f90:168-191 consumes one byte) and
// then handled the new byte from scratch -- so that entry is delta*(origin, FFFF^k) followed by
// the new byte's own boundary transition.  Sequences are decoded by bit concatenation only: no
// overlong / surrogate / range rejection (utf8_m.f90:383-429), code points above U+10FFFF match
// nothing.
namespace {

// Two INTER states are the same state when they agree on: how many bytes are still missing, where
// every completion lands, what the pending bytes replay to from here on (now, and after each further
// continuation byte), and the accept marks of the replay so far.  Neither the origin state nor the
// sequence length is part of the identity, so e.g. every search state of an ASCII-only pattern shares
// one small set of INTER states, and the overlong 3- and 4-byte forms share their tails.
struct InterKey {
    int missing;
    int failacc;
    int pending = 0;        // span tables: part of the identity when a replay accepts (RA counts back from the breaking byte)
    std::vector<int> tail;  // replay states chain[pending-1 .. total-2]
    std::vector<int> sig;   // (window-relative lo, dst) pairs, run-length form
    bool operator<(const InterKey& o) const {
        if (missing != o.missing) return missing < o.missing;
        if (failacc != o.failacc) return failacc < o.failacc;
        if (pending != o.pending) return pending < o.pending;
        if (tail != o.tail) return tail < o.tail;
        return sig < o.sig;
    }
};

struct ByteBuilder {
    const CpAutomaton& a;
    bool flag_bits;
    bool span_words = false;
    std::map<InterKey, int> inter_index;
    struct Inter { int origin, total, missing; long long lo; int chain[3]; int failacc; };  // window = [lo, lo + 64^missing)
    std::vector<Inter> inters;
    std::vector<std::vector<int> > rows;  // rows[state][byte]
    std::vector<uint8_t> flags;
    std::vector<int> remap;               // old state id -> new id (boolean tables)

    ByteBuilder(const CpAutomaton& au, bool fb) : a(au), flag_bits(fb) {}

    int cp_delta(int s, long long cp) const {
        return a.delta[(size_t)s * (size_t)a.nclasses + (size_t)a.class_of((int)cp)];
    }

    std::vector<int> signature(int s, long long lo, long long hi) const {
        std::vector<int> sig;
        long long p = lo;
        int last = -2;
        while (p <= hi) {
            int c = a.class_of((int)p);
            int d = a.delta[(size_t)s * (size_t)a.nclasses + (size_t)c];
            if (d != last) { sig.push_back((int)(p - lo)); sig.push_back(d); last = d; }
            p = a.cuts[(size_t)c + 1];
        }
        return sig;
    }

    // INTER state reached from `origin` after the lead byte and (total-1-missing) continuation
    // bytes, the code points still possible being [lo, lo + 64^missing).
    int inter_state(int origin, int total, int missing, long long lo) {
        long long span = 1ll << (6 * missing);
        int chain[3] = {-1, -1, -1};
        int st = origin;
        for (int i = 0; i < total - 1; i++) { st = cp_delta(st, 0xFFFF); chain[i] = st; }
        int pending = total - missing;
        InterKey k;
        k.missing = missing;
        k.failacc = 0;
        for (int i = 1; i <= pending; i++)
            if (a.accept[(size_t)chain[i - 1]]) k.failacc |= 1 << (i - 1);
        if (!flag_bits) k.failacc = 0;  // only the span kernels look at the replay marks
        if (span_words && k.failacc != 0) k.pending = pending;
        for (int i = pending - 1; i <= total - 2; i++) k.tail.push_back(chain[i]);
        k.sig = signature(origin, lo, lo + span - 1);
        if (k.sig.size() == 2 && k.sig[1] == 0 && k.failacc == 0) {
            // every completion is dead: the state is dead iff every possible replay is dead too
            bool all_dead = true;
            for (int t : k.tail) all_dead = all_dead && t == 0;
            if (all_dead) return 0;
        }
        auto it = inter_index.find(k);
        if (it != inter_index.end()) return it->second;
        int id = a.nstates + (int)inters.size();
        Inter in{origin, total, missing, lo, {chain[0], chain[1], chain[2]}, k.failacc};
        inters.push_back(in);
        inter_index.emplace(k, id);
        return id;
    }

    int build(ByteTable& out) {
        rows.assign((size_t)a.nstates, std::vector<int>(256, 0));
        for (int s = 0; s < a.nstates; s++) {
            std::vector<int>& row = rows[(size_t)s];
            for (int b = 0; b < 0x80; b++) row[(size_t)b] = cp_delta(s, b);
            int bad = cp_delta(s, 0xFFFF);  // stray continuation byte, or 11111xxx
            for (int b = 0x80; b < 0xC0; b++) row[(size_t)b] = bad;
            for (int b = 0xF8; b < 0x100; b++) row[(size_t)b] = bad;
            for (int b = 0xC0; b < 0xE0; b++) row[(size_t)b] = inter_state(s, 2, 1, (long long)(b & 31) << 6);
            for (int b = 0xE0; b < 0xF0; b++) row[(size_t)b] = inter_state(s, 3, 2, (long long)(b & 15) << 12);
            for (int b = 0xF0; b < 0xF8; b++) row[(size_t)b] = inter_state(s, 4, 3, (long long)(b & 7) << 18);
        }
        // INTER rows (the list grows while it is processed)
        for (size_t i = 0; i < inters.size(); i++) {
            Inter in = inters[i];
            std::vector<int> row(256, 0);
            int pending = in.total - in.missing;        // bytes swallowed so far
            int fallback = in.chain[pending - 1];       // origin after replaying them as U+FFFF
            for (int b = 0; b < 256; b++) {
                if ((b >> 6) == 2) {
                    long long lo = in.lo + ((long long)(b & 63) << (6 * (in.missing - 1)));
                    if (in.missing == 1) row[(size_t)b] = cp_delta(in.origin, lo);
                    else row[(size_t)b] = inter_state(in.origin, in.total, in.missing - 1, lo);
                } else {
                    row[(size_t)b] = rows[(size_t)fallback][(size_t)b];
                }
            }
            rows.push_back(row);
        }
        int total_states = a.nstates + (int)inters.size();
        if (total_states > (span_words ? (int)W_SSTATE : flag_bits ? (int)W_STATE : 0xFFFF)) return ERR_DFA_STATE_CAP;
        // flags
        flags.assign((size_t)total_states, 0);
        for (int s = 0; s < a.nstates; s++) {
            uint8_t f = 0;
            if (a.accept[(size_t)s]) f |= SF_ACC;
            if (a.end_accept[(size_t)s]) f |= SF_END;
            if (s == a.matched) f |= SF_MATCHED;
            flags[(size_t)s] = f;
        }
        for (size_t i = 0; i < inters.size(); i++) {
            const Inter& in = inters[i];
            int pending = in.total - in.missing;
            uint8_t f = SF_INTER;
            for (int k = 1; k <= pending; k++)
                if (in.failacc & (1 << (k - 1))) f |= (uint8_t)(SF_FAILACC1 << (k - 1));
            int st = in.chain[pending - 1];
            if (a.end_accept[(size_t)st]) f |= SF_END;   // text ends inside the sequence: pending bytes replay as U+FFFF
            if (st == a.matched) f |= SF_MATCHED;
            flags[(size_t)a.nstates + i] = f;
        }
        // Boolean tables: renumber so that every "result is true" state has an id >= result_threshold; the
        // kernels then decide with one compare instead of a flags[] lookup.  State 0 keeps its id.
        out.result_threshold = total_states;
        if (!flag_bits) {
            std::vector<int> perm((size_t)total_states, 0);
            int next = 1;
            for (int s = 1; s < total_states; s++)
                if (!(flags[(size_t)s] & (SF_END | SF_MATCHED))) perm[(size_t)s] = next++;
            out.result_threshold = next;
            for (int s = 1; s < total_states; s++)
                if (flags[(size_t)s] & (SF_END | SF_MATCHED)) perm[(size_t)s] = next++;
            std::vector<std::vector<int> > nrows((size_t)total_states);
            std::vector<uint8_t> nflags((size_t)total_states, 0);
            for (int s = 0; s < total_states; s++) {
                std::vector<int> r(256);
                for (int b = 0; b < 256; b++) r[(size_t)b] = perm[(size_t)rows[(size_t)s][(size_t)b]];
                nrows[(size_t)perm[(size_t)s]].swap(r);
                nflags[(size_t)perm[(size_t)s]] = flags[(size_t)s];
            }
            rows.swap(nrows);
            flags.swap(nflags);
            remap = perm;
        }
        // the words the kernels read: destination id + flag bits (flag-bit tables), + the replay mark RA of a
        // transition that breaks a multi-byte sequence (span tables)
        auto ra_of = [&](int s) -> int {       // RA of in-sequence state s: pending - (last replayed byte that accepts) + 1
            if (!span_words || s < a.nstates) return 0;
            const Inter& in = inters[(size_t)(s - a.nstates)];
            const int pending = in.total - in.missing;
            for (int k = pending; k >= 1; k--)
                if (in.failacc & (1 << (k - 1))) return pending - k + 1;
            return 0;
        };
        auto word = [&](int src, int b, int dst) -> uint16_t {
            uint16_t w = (uint16_t)dst;
            if (!flag_bits) return w;  // boolean kernels only read flags[] at the end of the text
            uint8_t f = flags[(size_t)dst];
            if (f & (SF_ACC | SF_MATCHED)) w |= W_ACC;
            if (f & SF_INTER) w |= W_INTER;
            if (span_words && (b >> 6) != 2) w |= (uint16_t)(ra_of(src) << W_RA_SHIFT);
            return w;
        };
        std::vector<std::vector<uint16_t> > words((size_t)total_states, std::vector<uint16_t>(256, 0));
        for (int s = 0; s < total_states; s++)
            for (int b = 0; b < 256; b++) words[(size_t)s][(size_t)b] = word(s, b, rows[(size_t)s][(size_t)b]);
        // byte classes: bytes whose columns (of words) agree in every state
        std::map<std::vector<uint16_t>, int> colid;
        out.nclasses = 0;
        for (int b = 0; b < 256; b++) {
            std::vector<uint16_t> col((size_t)total_states);
            for (int s = 0; s < total_states; s++) col[(size_t)s] = words[(size_t)s][(size_t)b];
            auto it = colid.find(col);
            if (it == colid.end()) { it = colid.emplace(col, out.nclasses++).first; }
            out.classmap[b] = (uint8_t)it->second;
        }
        out.row_shift = 0;
        while ((1 << out.row_shift) < out.nclasses) out.row_shift++;
        out.nstates = total_states;
        out.nboundary = a.nstates;
        out.table.assign((size_t)total_states << out.row_shift, 0);
        for (int s = 0; s < total_states; s++)
            for (int b = 0; b < 256; b++)
                out.table[((size_t)s << out.row_shift) + out.classmap[b]] = words[(size_t)s][(size_t)b];
        out.direct.assign((size_t)total_states * 256, 0);
        for (int s = 0; s < total_states; s++)
            for (int b = 0; b < 256; b++) out.direct[(size_t)s * 256 + (size_t)b] = words[(size_t)s][(size_t)b];
        out.span_words = span_words;
        out.endinfo.clear();
        if (span_words) {
            out.endinfo.assign((size_t)total_states, 0);
            for (int s = 0; s < total_states; s++) {
                uint8_t e = (uint8_t)ra_of(s);
                if (flags[(size_t)s] & SF_END) e |= EI_NUL;
                out.endinfo[(size_t)s] = e;
            }
        }
        out.direct8.clear();
        if (!flag_bits && total_states <= 255) {
            out.direct8.assign((size_t)total_states * 256, 0);
            for (int s = 0; s < total_states; s++)
                for (int b = 0; b < 256; b++) out.direct8[(size_t)s * 256 + (size_t)b] = (uint8_t)rows[(size_t)s][(size_t)b];
        }
        out.flags = flags;
        auto mapped = [&](int s) { return (s >= 0 && !remap.empty()) ? remap[(size_t)s] : s; };
        out.start = mapped(a.start); out.start_nul = mapped(a.start_nul); out.q0 = mapped(a.q0); out.matched = mapped(a.matched);
        out.q0_accepting = a.q0_accepting;
        out.flag_bits = flag_bits;
        return OK;
    }
};

}  // namespace

int build_byte_table(const CpAutomaton& a, bool flag_bits, ByteTable& out, bool span_words) {
    ByteBuilder bb(a, flag_bits);
    bb.span_words = span_words;
    return bb.build(out);
}

// ---------------------------------------------------------------------------------------------
// whole program
// ---------------------------------------------------------------------------------------------
int build_nfa_tables(const Nfa& nfa, NfaTables& t) {
    Subsets ss(nfa);
    const int ncls = (int)nfa.cuts.size() - 1;
    t = NfaTables();
    t.nstates = nfa.n;
    t.words = (int)ss.words;
    t.nclasses = ncls;
    t.exit = nfa.exit;
    t.cuts = nfa.cuts;
    t.q0 = ss.entry_closure;
    t.q0_accepting = Subsets::has(ss.entry_closure, nfa.exit);
    const size_t total = ((size_t)nfa.n + 1) * (size_t)ncls * ss.words;
    if (nfa.n > 8191 || total > (size_t)32 << 20) return ERR_DFA_STATE_CAP;        // 8191 NFA states / 256 MB of sets
    t.trans.assign(total, 0);
    for (int s = 1; s <= nfa.n; s++)
        for (auto& cd : ss.by_class[(size_t)s]) {
            Bits b(ss.words, 0);
            for (size_t w = 0; w < ss.words; w++) b[w] = t.trans[((size_t)s * (size_t)ncls + (size_t)cd.first) * ss.words + w];
            if (!Subsets::has(b, cd.second)) ss.close_from(b, cd.second);
            for (size_t w = 0; w < ss.words; w++) t.trans[((size_t)s * (size_t)ncls + (size_t)cd.first) * ss.words + w] = b[w];
        }
    return OK;
}

// NFA + literals -> every table the kernels walk (shared by the pattern route and the DFA route below)
static int compile_from_nfa(const Nfa& nfa, int op, int state_cap, Program& p, bool want_span) {
    p.literal_only = !fortran_blank(p.lit.all);
    p.prefix_active = !fortran_blank(p.lit.prefix);
    p.nfa_states = nfa.n;
    int rc = build_cp_automaton(nfa, (Mode)op, state_cap, p.cp);
    if (rc == ERR_DFA_STATE_CAP && want_span) {
        // the eager automaton is too large: the device simulates the NFA instead (want_span is false only for the
        // optional anchored twin of an `.in.` handle, which is simply absent then)
        if (build_nfa_tables(nfa, p.nfa_tables) == OK) { p.nfa_engine = true; p.cp = CpAutomaton(); return OK; }
    }
    if (rc != OK) { p.status = rc; return rc; }
    rc = build_byte_table(p.cp, op == MODE_REGEX, p.bt);
    if (rc != OK) { p.status = rc; return rc; }
    if (want_span && op == MODE_REGEX && !p.literal_only) {
        // linear-time span path; silently absent when a cap is exceeded (the anchored tables above still serve)
        if (build_span_forward(nfa, state_cap, p.span_cp) == OK && build_byte_table(p.span_cp, true, p.span_bt, true) == OK &&
            build_rev_automaton(nfa, 0xFFFF, p.rev) == OK) {
            p.has_span_tables = true;
            p.has_span = !p.prefix_active;      // with a prefix literal Forgex's candidates are the literal's occurrences
        }
    }
    return OK;
}

int compile_program(const std::string& pattern, int op, int state_cap, Program& p, bool want_span) {
    p = Program();
    p.op = op;
    p.prepared = prepare_pattern(pattern, op == MODE_MATCH);
    Syntax syn;
    parse_pattern(p.prepared, syn);
    p.status = syn.status;
    if (!syn.valid()) return p.status;
    extract_literals(syn, p.lit);
    Nfa nfa;
    build_nfa(syn, nfa);
    return compile_from_nfa(nfa, op, state_cap, p, want_span);
}

// The Fortran-side route: the host has already run Forgex's own front end and explored its automaton
// eagerly -- a breadth-first search that calls automaton%construct (src/automaton_m.F90:333) for every state and one
// representative symbol of every alphabet segment -- and hands over the resulting ANCHORED code-point DFA plus the
// literals extract_literal produced.  A DFA is an NFA: state s of the DFA becomes NFA state s + 2, accepting states get
// an epsilon move to the exit, and the same builders as above derive every mode's automaton from it (the search
// automaton of `.in.`, the ordered-groups and reverse automata of the span path are subset constructions over it).
//   cuts[ncls + 1]: ascending code points, class c = [cuts[c], cuts[c+1] - 1], cuts[0] = 0; delta[nstates x ncls] with
//   state 0 = dead; accept[nstates]; q0 = the state before any symbol.
int compile_from_dfa(const DfaInput& in, int op, int state_cap, Program& p, bool want_span) {
    p = Program();
    p.op = op;
    p.status = OK;
    p.lit.all = in.all; p.lit.prefix = in.prefix; p.lit.suffix = in.suffix;
    if (in.nstates < 1 || in.ncls < 1 || in.q0 <= 0 || in.q0 >= in.nstates || in.cuts[0] != 0) { p.status = ERR_BAD_ARGUMENT; return p.status; }
    for (int c = 0; c < in.ncls; c++) if (in.cuts[(size_t)c] >= in.cuts[(size_t)c + 1]) { p.status = ERR_BAD_ARGUMENT; return p.status; }
    Nfa nfa;
    nfa.n = in.nstates + 2;             // NFA states 1 (entry), 2 (exit), s + 2 for DFA state s >= 1
    nfa.entry = 1;
    nfa.exit = 2;
    nfa.eps.assign((size_t)nfa.n + 1, {});
    nfa.edges.assign((size_t)nfa.n + 1, {});
    nfa.eps[1].push_back(in.q0 + 2);
    std::vector<int> cuts(in.cuts, in.cuts + in.ncls + 1);
    if (cuts.back() > CP_MAX + 1) cuts.back() = CP_MAX + 1;     // nothing matches above U+10FFFF
    for (int s = 1; s < in.nstates; s++) {
        if (in.accept[(size_t)s]) nfa.eps[(size_t)s + 2].push_back(2);
        for (int c = 0; c < in.ncls; c++) {
            const int d = in.delta[(size_t)s * (size_t)in.ncls + (size_t)c];
            if (d < 0 || d >= in.nstates) { p.status = ERR_BAD_ARGUMENT; return p.status; }
            if (d == 0 || cuts[(size_t)c] > CP_MAX) continue;
            int hi = cuts[(size_t)c + 1] - 1;
            if (hi > CP_MAX) hi = CP_MAX;
            nfa.edges[(size_t)s + 2].push_back({Range{cuts[(size_t)c], hi}, d + 2});
        }
    }
    std::vector<int> all = cuts;
    all.push_back(0);
    all.push_back(CP_MAX + 1);
    all.push_back(CP_TOP + 1);
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    nfa.cuts = all;
    return compile_from_nfa(nfa, op, state_cap, p, want_span);
}

}  // namespace fx
