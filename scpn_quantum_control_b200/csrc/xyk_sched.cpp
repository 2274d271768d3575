// Schedule compiler: turns the reference's ordered gate list into cache-blocked tile passes.
//
// The reference applies, per Trotter layer, one diagonal phase and then every pair rotation
// exp(-i theta_ij (X_iX_j + Y_iY_j)) in lexicographic (i, j) order
// (reference: src/scpn_quantum_control/bridge/knm_hamiltonian.py:176-195 gives the term order,
// Qiskit's LieTrotter / SuzukiTrotter apply the terms in list order).  Two rotations commute
// iff they share no qubit, so any execution order that preserves the relative order of every
// pair of gates sharing a qubit is the same operator.  The compiler groups gates into
//   pass  = one read+write sweep over the state; 2^TB-amplitude tiles live in shared memory;
//   round = one shared-memory round trip; each thread holds 2^RB amplitudes in registers and
//           applies every enabled rotation among its RB register qubits;
// and chooses, for every pass, the bit permutation its store applies so that the next pass
// finds its tile qubits on the lowest physical index bits (contiguous tiles).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <thread>
#include <set>
#include <exception>
#include <stdexcept>

#include "xyk_internal.h"

namespace xyk {

void build_terms(int n, const double* K, const double* omega, std::vector<PairTerm>& pairs,
                 std::vector<ZTerm>& zterms, std::vector<double>& Ksym) {
    pairs.clear();
    zterms.clear();
    Ksym.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (i != j) Ksym[(size_t)i * n + j] = (K[(size_t)i * n + j] + K[(size_t)j * n + i]) / 2.0;
    for (int i = 0; i < n; ++i)
        if (std::fabs(omega[i]) > kSparsityEps) zterms.push_back({i, omega[i]});
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const double k = Ksym[(size_t)i * n + j];
            if (std::fabs(k) < kSparsityEps) continue;
            pairs.push_back({i, j, k});
        }
}

static void suzuki_fracs(int order, std::vector<double>& out) {
    if (order == 2) {
        out.push_back(1.0);
        return;
    }
    const double p = 1.0 / (4.0 - std::pow(4.0, 1.0 / (order - 1)));
    std::vector<double> inner;
    suzuki_fracs(order - 2, inner);
    const double w[5] = {p, p, 1.0 - 4.0 * p, p, p};
    for (double ww : w)
        for (double f : inner) out.push_back(ww * f);
}

std::vector<Segment> build_sequence(const std::vector<PairTerm>& pairs, double delta, int order,
                                    int reps, int* layers1_equiv) {
    std::vector<Segment> segs;
    const double tau = delta / reps;
    const int M = (int)pairs.size();
    int sweep = 0;
    if (order == 1) {
        for (int r = 0; r < reps; ++r) {
            Segment s;
            s.ztau = tau;
            for (const auto& p : pairs) s.gates.push_back({p.i, p.j, 2.0 * p.k * tau, sweep});
            ++sweep;
            segs.push_back(std::move(s));
        }
        if (layers1_equiv) *layers1_equiv = reps;
        return segs;
    }
    std::vector<double> fr;
    suzuki_fracs(order, fr);
    double pending_z = 0.0;
    for (int r = 0; r < reps; ++r)
        for (double f : fr) {
            const double t = f * tau;
            Segment s;
            if (M == 0) {  // diagonal only: the two half phases merge
                pending_z += t;
                continue;
            }
            s.ztau = pending_z + t / 2.0;
            for (int k = 0; k < M - 1; ++k)
                s.gates.push_back({pairs[k].i, pairs[k].j, 2.0 * pairs[k].k * t / 2.0, sweep});
            s.gates.push_back({pairs[M - 1].i, pairs[M - 1].j, 2.0 * pairs[M - 1].k * t, sweep});
            ++sweep;
            for (int k = M - 2; k >= 0; --k)
                s.gates.push_back({pairs[k].i, pairs[k].j, 2.0 * pairs[k].k * t / 2.0, sweep});
            ++sweep;
            pending_z = t / 2.0;
            segs.push_back(std::move(s));
        }
    if (pending_z != 0.0 || segs.empty()) {
        Segment s;
        s.ztau = pending_z;
        segs.push_back(std::move(s));
    }
    if (layers1_equiv) *layers1_equiv = 2 * reps * (int)fr.size();
    return segs;
}

// ---------------------------------------------------------------------------------------------
// dependency-closed executable set
// ---------------------------------------------------------------------------------------------
namespace {

struct Gate {
    int i, j, sweep;
    bool reverse;  // gate belongs to a descending sweep
};

// Gates of `g` (not yet done, index >= start) executable when only the qubits of `tmask` are
// on chip: scanning in sequence order, a gate runs iff both qubits are on chip and neither
// has been blocked by an earlier gate that could not run.
void closure(const std::vector<Gate>& g, const std::vector<char>& done, int start, uint64_t tmask,
             std::vector<int>& out) {
    out.clear();
    uint64_t blocked = 0;
    for (int idx = start; idx < (int)g.size(); ++idx) {
        if (done[idx]) continue;
        const uint64_t m = (1ull << g[idx].i) | (1ull << g[idx].j);
        if ((m & tmask) == m && !(m & blocked)) out.push_back(idx);
        else blocked |= m;
        if (__builtin_popcountll(tmask & ~blocked) < 2) break;
    }
}

// Same, restricted to what ONE register round can do: a single sweep (monotone order) and each
// pair at most once.
void round_closure(const std::vector<Gate>& g, const std::vector<int>& cand, const std::vector<char>& cdone,
                   uint64_t rmask, std::vector<int>& out) {
    out.clear();
    uint64_t blocked = 0;
    int sweep = -1;
    for (int c = 0; c < (int)cand.size(); ++c) {
        if (cdone[c]) continue;
        const Gate& q = g[cand[c]];
        const uint64_t m = (1ull << q.i) | (1ull << q.j);
        const bool fits = (m & rmask) == m && !(m & blocked) && (sweep < 0 || q.sweep == sweep);
        if (fits) {
            out.push_back(c);
            sweep = q.sweep;
        } else {
            blocked |= m;
        }
        if (__builtin_popcountll(rmask & ~blocked) < 2) break;
    }
}

int popc(uint64_t x) { return __builtin_popcountll(x); }


std::mutex g_memo_mutex;  // guards the process-wide tile-sequence and round-decomposition memos (held for a whole compile)

// The schedule compiler: four stages over shared state.
//   prepare()        gate lists per segment (sweep directions), merged diagonal phases, configuration signature
//   choose_tiles()   1. the tile (set of TB logical qubits) of every pass, segment by segment
//   search_rounds()  2. the register rounds of every pass
//   assign_layouts() 3. the order of the qubits on the physical bits before every pass
//   emit_passes()    4. PlanPass descriptors (thread / lane / warp assignment, coefficients), over-long passes cut in pieces
// Each stage can be exercised on its own through compile_plan_stages() (tests/test_schedule_cpu.py).
struct PlanCompiler {
    struct RawPass {
        uint64_t tmask;
        int seg;
        std::vector<int> gate_idx;  // indices into the segment's gate list
        bool has_phase;
    };
    struct RawRound {
        uint64_t rmask = 0;
        std::vector<int> cidx;  // indices into the pass's gate_idx
        bool reverse = false;
        bool split = false;     // split round: gates only between lmask and rmask & ~lmask
        uint64_t lmask = 0;
    };

    const int nq;
    const std::vector<Segment>& segs;
    const SchedConfig& cfg;
    int TB = 0, RB = 0;
    int RS = 0;                                // register qubits of a split round
    Plan plan;
    std::string cfg_sig;                       // everything of the configuration the memoised searches depend on
    std::vector<std::vector<Gate>> sg;         // prepare: gates per segment
    std::vector<double> zeff;                  // prepare: diagonal phase in front of every segment with gates
    std::vector<RawPass> raw;                  // stage 1: tiles and their gates
    int P = 0;                                 // number of passes
    std::vector<std::vector<RawRound>> rr;     // stage 2: rounds per pass
    std::vector<uint64_t> rlast, rfirst;       // stage 2: register qubits of the last / first round per pass
    int c_lane = 0;
    std::vector<std::vector<int>> layouts;     // stage 3: logical -> physical before pass p (P + 1 entries)
    std::vector<int> nshared, nfree;

    PlanCompiler(int nq_, const std::vector<Segment>& segs_, const SchedConfig& cfg_) : nq(nq_), segs(segs_), cfg(cfg_) {}

    int prev_index(int p) const {
        if (p > 0) return p - 1;
        return (cfg.cyclic && P > 1) ? P - 1 : -1;
    }

    void prepare() {
    if (nq < 1 || nq > kMaxQ) throw std::runtime_error("compile_plan: bad qubit count");
    TB = std::min(cfg.TB, nq);
    RB = std::min(cfg.RB, TB);
    RS = RB + 1;
    plan.nq = nq;
    plan.TB = TB;
    plan.RB = RB;
    {
        std::ostringstream os;
        os << nq << '/' << TB << '/' << RB << '/' << cfg.need_shared << '/' << cfg.lane_bits << '/' << cfg.fuse_io << cfg.split_rounds << cfg.cyclic
           << cfg.unb_first << cfg.unb_last << '/' << cfg.split_margin << '/' << cfg.split_min_left << '/' << cfg.beam_width << '/' << cfg.beam_cands << '/'
           << cfg.tile_groups << cfg.tile_groups_open << '/' << cfg.tile_beam << '/' << cfg.tile_seeds << '/' << cfg.next_tile_hint << '/' << cfg.prev_tile_hint << '/'
           << cfg.prev_rlast_hint << '/' << cfg.final_positions.size() << '/' << cfg.final_top.size() << '|';
        cfg_sig = os.str();
    }


    sg.assign(segs.size(), {});
    for (size_t s = 0; s < segs.size(); ++s) {
        int last_sweep = -1, sweep_no = -1;
        for (size_t k = 0; k < segs[s].gates.size(); ++k) {
            const GateOp& go = segs[s].gates[k];
            if (go.i < 0 || go.j >= nq || go.i >= go.j) throw std::runtime_error("compile_plan: bad gate");
            if (go.sweep != last_sweep) {
                last_sweep = go.sweep;
                ++sweep_no;
            }
            // a sweep is descending iff its first two gates are in descending lexicographic order;
            // decided below once the whole sweep is known
            sg[s].push_back({go.i, go.j, go.sweep, false});
        }
        // direction per sweep
        size_t a = 0;
        while (a < sg[s].size()) {
            size_t b = a;
            while (b < sg[s].size() && sg[s][b].sweep == sg[s][a].sweep) ++b;
            bool rev = false;
            if (b - a >= 2) {
                const Gate &x = sg[s][a], &y = sg[s][a + 1];
                rev = (x.i > y.i) || (x.i == y.i && x.j > y.j);
            }
            for (size_t k = a; k < b; ++k) sg[s][k].reverse = rev;
            a = b;
        }
    }

    // diagonal phases of gate-free segments commute with everything diagonal: merge them forward
    zeff.assign(segs.size(), 0.0);
    {
        double carry = 0.0;
        for (size_t s = 0; s < segs.size(); ++s) {
            if (segs[s].gates.empty()) carry += segs[s].ztau;
            else {
                zeff[s] = segs[s].ztau + carry;
                carry = 0.0;
            }
        }
        plan.trailing_ztau = carry;
        plan.has_trailing_phase = carry != 0.0;
    }
    }

    // ---- 1. choose the tile (set of TB logical qubits) of every pass, segment by segment ----
    void choose_tiles() {
    const uint64_t all_mask = nq >= 64 ? ~0ull : ((1ull << nq) - 1);
    // qubits of the program's first tile ranked by how late the first pass uses them: when the last tile
    // has to share qubits with the first one (cyclic programs), the late ones are free to become the
    // lane bits of the fused global I/O between the two passes
    std::vector<int> first_use(nq, 1 << 30);
    auto grow_tile = [&](const std::vector<Gate>& g, const std::vector<char>& done, int first, int seed,
                         uint64_t prev, uint64_t firstTile, int need_prev, int need_first) -> uint64_t {
        uint64_t T = (1ull << g[seed].i) | (1ull << g[seed].j);
        std::vector<int> ex;
        while (popc(T) < TB) {
            const int slots = TB - popc(T);
            const int def_prev = prev ? std::max(0, need_prev - popc(T & prev)) : 0;
            const int def_first = firstTile ? std::max(0, need_first - popc(T & firstTile)) : 0;
            uint64_t cands = all_mask & ~T;
            // when the remaining slots are just enough, only qubits that reduce a deficit qualify
            if (def_prev + def_first >= slots) {
                uint64_t c2 = 0;
                if (def_prev) c2 |= prev & ~T;
                if (def_first) c2 |= firstTile & ~T;
                // prefer qubits that reduce both deficits when both are outstanding
                if (def_prev && def_first && (prev & firstTile & ~T)) c2 = prev & firstTile & ~T;
                if (c2) cands = c2;
            }
            int best = -1, bq = -1, bshare = -1;
            for (int q = 0; q < nq; ++q) {
                if (!((cands >> q) & 1)) continue;
                closure(g, done, first, T | (1ull << q), ex);
                const int sc = (int)ex.size();
                const int share = (int)((prev >> q) & 1);
                const bool later = firstTile && bq >= 0 && ((firstTile >> q) & 1) && first_use[q] > first_use[bq];
                if (sc > best || (sc == best && share > bshare) || (sc == best && share == bshare && later)) {
                    best = sc;
                    bq = q;
                    bshare = share;
                }
            }
            if (bq < 0) break;
            T |= 1ull << bq;
        }
        return T;
    };

    uint64_t prev_tile = 0, first_tile = 0;
    const int need = std::min(cfg.need_shared, TB);
    // Tile sequence of a segment: breadth-first beam search over the pass count.  A state expands into one tile per
    // seed gate (the first cfg.tile_seeds gates that are ready, i.e. first in line for both their qubits), each grown
    // greedily from its seed; cfg.tile_beam states with the most gates done survive a depth.  Width 1 with one seed is
    // the plain greedy sequence.  Segments with the same gate pattern entered from the same tile reuse the result.
    struct TileState {
        std::vector<char> done;
        size_t ndone = 0;
        uint64_t prev = 0;
        std::vector<RawPass> passes;
    };
    struct TileMemo {
        std::vector<uint64_t> tiles;
    };
    // process-wide: the same gate pattern (another step size, another handle) reuses the searched tile sequence
    static std::map<std::string, TileMemo> tile_memo;
    if (tile_memo.size() > 4096) tile_memo.clear();
    size_t last_seg_with_gates = 0;
    for (size_t s = 0; s < segs.size(); ++s)
        if (!sg[s].empty()) last_seg_with_gates = s;
    for (size_t s = 0; s < segs.size(); ++s) {
        const std::vector<Gate>& g = sg[s];
        if (g.empty()) continue;
        auto make_pass = [&](const std::vector<char>& done, int first, uint64_t T, bool phase, RawPass& rp) {
            rp.tmask = T;
            rp.seg = (int)s;
            rp.has_phase = phase;
            closure(g, done, first, T, rp.gate_idx);
        };
        std::string key;
        {
            std::ostringstream os;
            os << cfg_sig << prev_tile << ':' << (s == last_seg_with_gates ? (cfg.cyclic ? first_tile : cfg.next_tile_hint) : 0) << ':';
            const int s0 = g[0].sweep;
            for (const Gate& q : g) os << q.i << ',' << q.j << ',' << (q.sweep - s0) << ';';
            key = os.str();
        }
        std::vector<RawPass> chosen;
        auto mit = tile_memo.find(key);
        if (mit != tile_memo.end()) {
            std::vector<char> done(g.size(), 0);
            int first = 0;
            bool phase = true;
            for (uint64_t T : mit->second.tiles) {
                while (done[first]) ++first;
                RawPass rp;
                make_pass(done, first, T, phase, rp);
                for (int k : rp.gate_idx) done[k] = 1;
                phase = false;
                chosen.push_back(std::move(rp));
            }
        } else if (cfg.tile_groups > 0 && nq > TB && (cfg.cyclic || cfg.tile_groups_open)) {
            // Cost-guided beam over tile sequences.  Candidate tiles: unions of TB / 3 groups of three consecutive qubit
            // indices (the natural blocks of a lexicographic sweep) that share >= `need` qubits with the previous tile,
            // plus the greedy tile.  A pass is priced with the measured model of the pass kernel:
            //   max(floor, 0.6 + 1.14 R + 0.05 G) ms, R ~ estimated rounds, G gates,
            //   floor = 5.7 ms (+1.2 ms when only three qubits are shared: 128-byte store chunks).
            const int ngroups = (nq + 2) / 3, per_tile = TB / 3;
            std::vector<uint64_t> gmask(ngroups, 0);
            for (int q = 0; q < nq; ++q) gmask[q / 3] |= 1ull << q;
            std::vector<uint64_t> cand_tiles;
            {
                std::vector<int> idx(per_tile);
                for (int k = 0; k < per_tile; ++k) idx[k] = k;
                if (ngroups >= per_tile)
                    while (true) {
                        uint64_t T = 0;
                        for (int k = 0; k < per_tile; ++k) T |= gmask[idx[k]];
                        cand_tiles.push_back(T);
                        int k = per_tile - 1;
                        while (k >= 0 && idx[k] == ngroups - per_tile + k) --k;
                        if (k < 0) break;
                        ++idx[k];
                        for (int m = k + 1; m < per_tile; ++m) idx[m] = idx[m - 1] + 1;
                    }
            }
            struct CState {
                std::vector<char> done;
                size_t ndone = 0;
                uint64_t prev = 0;
                double cost = 0.0;
                std::vector<uint64_t> tiles;
            };
            auto pass_cost = [&](const std::vector<int>& gate_idx, int shared, bool has_prev) {
                // rounds: nine gates per split round at best, plus one round per two index triples with internal gates
                // (a triangle never fits a split round)
                uint64_t tri = 0;
                for (int k : gate_idx)
                    if (g[k].i / 3 == g[k].j / 3) tri |= 1ull << (g[k].i / 3);
                const int R = ((int)gate_idx.size() + 8) / 9 + (popc(tri) + 1) / 2;
                const double body = 0.6 + 1.14 * R + 0.05 * (double)gate_idx.size();
                const double floor = 5.7 + ((has_prev && shared < 4) ? 1.2 : 0.0);
                return std::max(body, floor);
            };
            const int W = std::max(1, cfg.tile_groups);
            const double lambda = 0.24;  // ms of credit per executed gate when ranking partial sequences
            std::vector<CState> beam(1);
            beam[0].done.assign(g.size(), 0);
            beam[0].prev = prev_tile;
            CState best_full;
            bool have_full = false;
            std::vector<int> ex;
            for (int depth = 0; depth < 4 * (int)g.size() && !beam.empty(); ++depth) {
                std::vector<CState> next;
                for (const CState& st : beam) {
                    if (have_full && st.cost + 5.7 >= best_full.cost) continue;  // cannot beat the best complete sequence
                    int first = 0;
                    while (st.done[first]) ++first;
                    std::vector<uint64_t> tiles_here;
                    tiles_here.push_back(grow_tile(g, st.done, first, first, st.prev, 0, need, 0));
                    for (uint64_t T : cand_tiles)
                        if (!st.prev || popc(T & st.prev) >= std::min(need, popc(st.prev))) tiles_here.push_back(T);
                    for (size_t ti = 0; ti < tiles_here.size(); ++ti) {
                        const uint64_t T = tiles_here[ti];
                        if (ti > 0 && T == tiles_here[0]) continue;
                        closure(g, st.done, first, T, ex);
                        if (ex.empty()) continue;
                        CState ns;
                        ns.done = st.done;
                        for (int k : ex) ns.done[k] = 1;
                        ns.ndone = st.ndone + ex.size();
                        ns.prev = T;
                        ns.cost = st.cost + pass_cost(ex, st.prev ? popc(T & st.prev) : TB, st.prev != 0);
                        ns.tiles = st.tiles;
                        ns.tiles.push_back(T);
                        if (ns.ndone == g.size()) {
                            // closing the cycle: a last tile that does not share enough with the program's first tile costs a pass
                            // (open programs chained to a next one: its first tile, as far as it is local here)
                            const uint64_t ft = cfg.cyclic ? (first_tile ? first_tile : ns.tiles[0]) : cfg.next_tile_hint;
                            if (ft && s == last_seg_with_gates && popc(T & ft) < std::min(need, popc(ft))) ns.cost += 5.7;
                            if (!have_full || ns.cost < best_full.cost) {
                                best_full = ns;
                                have_full = true;
                            }
                        } else {
                            next.push_back(std::move(ns));
                        }
                    }
                }
                std::sort(next.begin(), next.end(), [&](const CState& a, const CState& b) {
                    return a.cost - lambda * (double)a.ndone < b.cost - lambda * (double)b.ndone;
                });
                beam.clear();
                for (CState& ns : next) {
                    bool dup = false;
                    for (const CState& o : beam)
                        if (o.prev == ns.prev && o.ndone == ns.ndone && o.done == ns.done) {
                            dup = true;
                            break;
                        }
                    if (!dup) beam.push_back(std::move(ns));
                    if ((int)beam.size() >= W) break;
                }
            }
            {
                // the plain greedy sequence priced with the same model: the beam's ranking can prune it away
                CState gs;
                gs.done.assign(g.size(), 0);
                gs.prev = prev_tile;
                int first = 0;
                while (gs.ndone < g.size()) {
                    while (gs.done[first]) ++first;
                    const uint64_t T = grow_tile(g, gs.done, first, first, gs.prev, 0, need, 0);
                    closure(g, gs.done, first, T, ex);
                    if (ex.empty()) break;
                    for (int k : ex) gs.done[k] = 1;
                    gs.ndone += ex.size();
                    gs.cost += pass_cost(ex, gs.prev ? popc(T & gs.prev) : TB, gs.prev != 0);
                    gs.prev = T;
                    gs.tiles.push_back(T);
                }
                if (gs.ndone == g.size()) {
                    const uint64_t ft = cfg.cyclic ? (first_tile ? first_tile : gs.tiles[0]) : cfg.next_tile_hint;
                    if (ft && s == last_seg_with_gates && popc(gs.prev & ft) < std::min(need, popc(ft))) gs.cost += 5.7;
                    if (!have_full || gs.cost < best_full.cost) {
                        best_full = gs;
                        have_full = true;
                    }
                }
            }
            if (!have_full) throw std::runtime_error("compile_plan: tile search found no complete sequence");
            {
                std::vector<char> done(g.size(), 0);
                int first = 0;
                bool phase = true;
                for (uint64_t T : best_full.tiles) {
                    while (done[first]) ++first;
                    RawPass rp;
                    make_pass(done, first, T, phase, rp);
                    for (int k : rp.gate_idx) done[k] = 1;
                    phase = false;
                    chosen.push_back(std::move(rp));
                }
            }
            TileMemo tm;
            tm.tiles = best_full.tiles;
            tile_memo[key] = std::move(tm);
        } else {
            std::vector<TileState> beam(1);
            beam[0].done.assign(g.size(), 0);
            beam[0].prev = prev_tile;
            const int W = std::max(1, cfg.tile_beam), NS = std::max(1, cfg.tile_seeds);
            bool finished = false;
            for (int depth = 0; !finished; ++depth) {
                if (depth > (int)g.size() + 1) throw std::runtime_error("compile_plan: tile search does not terminate");
                std::vector<TileState> next;
                for (const TileState& st : beam) {
                    int first = 0;
                    while (st.done[first]) ++first;
                    // seeds: gates first in line for both their qubits (the first undone gate is always one)
                    std::vector<int> seeds;
                    {
                        uint64_t seen = 0;
                        for (int k = first; k < (int)g.size() && (int)seeds.size() < NS; ++k) {
                            if (st.done[k]) continue;
                            const uint64_t m = (1ull << g[k].i) | (1ull << g[k].j);
                            if (!(m & seen)) seeds.push_back(k);
                            seen |= m;
                            if (seen == all_mask) break;
                        }
                    }
                    std::vector<uint64_t> tried;
                    for (int seed : seeds) {
                        const uint64_t T = grow_tile(g, st.done, first, seed, st.prev, 0, need, 0);
                        if (std::find(tried.begin(), tried.end(), T) != tried.end()) continue;
                        tried.push_back(T);
                        RawPass rp;
                        make_pass(st.done, first, T, depth == 0, rp);
                        if (rp.gate_idx.empty()) continue;
                        TileState ns = st;
                        for (int k : rp.gate_idx) ns.done[k] = 1;
                        ns.ndone += rp.gate_idx.size();
                        ns.prev = T;
                        ns.passes.push_back(std::move(rp));
                        next.push_back(std::move(ns));
                    }
                }
                if (next.empty()) throw std::runtime_error("compile_plan: no progress");
                // complete sequences: prefer one whose last tile already shares enough qubits with the program's first tile
                int pick = -1;
                for (int k = 0; k < (int)next.size(); ++k)
                    if (next[k].ndone == g.size()) {
                        const uint64_t ft = first_tile ? first_tile : next[k].passes[0].tmask;
                        const bool closes = !(cfg.cyclic && s == last_seg_with_gates) || popc(next[k].prev & ft) >= std::min(need, popc(ft));
                        if (pick < 0 || closes) pick = k;
                        if (closes) break;
                    }
                if (pick >= 0) {
                    chosen = std::move(next[pick].passes);
                    finished = true;
                    break;
                }
                std::stable_sort(next.begin(), next.end(), [](const TileState& a, const TileState& b) { return a.ndone > b.ndone; });
                beam.clear();
                for (TileState& ns : next) {
                    bool dup = false;
                    for (const TileState& o : beam)
                        if (o.prev == ns.prev && o.done == ns.done) dup = true;
                    if (!dup) beam.push_back(std::move(ns));
                    if ((int)beam.size() >= W) break;
                }
            }
            TileMemo tm;
            for (const RawPass& rp : chosen) tm.tiles.push_back(rp.tmask);
            tile_memo[key] = std::move(tm);
        }
        for (RawPass& rp : chosen) {
            if (!first_tile) first_tile = rp.tmask;
            prev_tile = rp.tmask;
            raw.push_back(std::move(rp));
        }
    }
    // cyclic closure: the last tile must share `need` qubits with the first one so that the
    // program can be replayed from its own end layout.  If it does not, peel gates off the last
    // pass(es) and re-plan them with the extra constraint.
    if (!raw.empty()) {
        const std::vector<Gate>& g0 = sg[raw[0].seg];
        for (size_t k = 0; k < raw[0].gate_idx.size(); ++k) {
            const Gate& q = g0[raw[0].gate_idx[k]];
            first_use[q.i] = std::min(first_use[q.i], (int)k);
            first_use[q.j] = std::min(first_use[q.j], (int)k);
        }
    }
    if (!cfg.cyclic) first_tile = cfg.next_tile_hint;  // open program: share with the next program's first tile
    if (first_tile && !raw.empty() && popc(raw.back().tmask & first_tile) < std::min(need, popc(first_tile)) && nq > TB) {
        const int s = raw.back().seg;
        const std::vector<Gate>& g = sg[s];
        // undo the last pass of segment s and re-plan the remaining gates of s
        std::vector<char> done(g.size(), 1);
        const bool had_phase = raw.back().has_phase;
        for (int k : raw.back().gate_idx) done[k] = 0;
        raw.pop_back();
        prev_tile = raw.empty() ? 0 : raw.back().tmask;
        size_t ndone = 0;
        for (char d : done) ndone += d;
        bool first_of_seg = had_phase;
        int first = 0;
        while (ndone < g.size()) {
            while (done[first]) ++first;
            uint64_t T = grow_tile(g, done, first, first, prev_tile, first_tile, need, std::min(need, popc(first_tile)));
            RawPass rp;
            rp.tmask = T;
            rp.seg = s;
            rp.has_phase = first_of_seg;
            closure(g, done, first, T, rp.gate_idx);
            if (rp.gate_idx.empty()) throw std::runtime_error("compile_plan: no progress (cyclic)");
            for (int k : rp.gate_idx) done[k] = 1;
            ndone += rp.gate_idx.size();
            first_of_seg = false;
            prev_tile = T;
            raw.push_back(std::move(rp));
        }
        if (raw.empty()) first_tile = 0;
    }
    }

    // ---- 2. register rounds of every pass (sets of RB qubits + the gates they execute) ----
    // The LAST round of a pass is chosen first (closure over the reversed sequence) and the FIRST
    // round next, both with a cap on how many qubits shared with the neighbouring tile they may
    // hold in registers: the qubits that stay out of both become the lowest physical bits of the
    // layout between the two passes, which lets the last round store and the next first round
    // load global memory directly with lanes on contiguous addresses.
    void search_rounds() {
    P = (int)raw.size();
    auto next_tile = [&](int p) -> uint64_t {
        if (P == 1 && cfg.cyclic) return 0;
        if (p + 1 < P) return raw[p + 1].tmask;
        return cfg.cyclic ? raw[0].tmask : cfg.next_tile_hint;
    };
    const bool can_split = cfg.split_rounds && (RS % 2 == 0) && TB - RS >= 3 && RS <= 8;
    rr.assign(P, {});
    rlast.assign(P, 0);
    rfirst.assign(P, 0);
    c_lane = std::min(cfg.lane_bits, TB);

    auto rev_round_closure = [&](const std::vector<Gate>& g, const std::vector<int>& cand,
                                 const std::vector<char>& cdone, uint64_t rmask, std::vector<int>& out) {
        out.clear();
        uint64_t blocked = 0;
        int sweep = -1;
        for (int c = (int)cand.size() - 1; c >= 0; --c) {
            if (cdone[c]) continue;
            const Gate& q = g[cand[c]];
            const uint64_t m = (1ull << q.i) | (1ull << q.j);
            const bool fits = (m & rmask) == m && !(m & blocked) && (sweep < 0 || q.sweep == sweep);
            if (fits) {
                out.push_back(c);
                sweep = q.sweep;
            } else {
                blocked |= m;
            }
            if (popc(rmask & ~blocked) < 2) break;
        }
    };
    // exhaustive search over RB-subsets of the tile; `limit_mask`/`limit` cap |R & limit_mask|;
    // `prefer` breaks ties towards subsets overlapping it (keeps the union with a neighbour small)
    // split-round closure: a gate runs iff it joins the left group with the right group
    auto split_closure = [&](const std::vector<Gate>& g, const std::vector<int>& cand, const std::vector<char>& cdone,
                             uint64_t lm, uint64_t rm, bool reverse_scan, std::vector<int>& out) {
        out.clear();
        uint64_t blocked = 0;
        int sweep = -1;
        const uint64_t both = lm | rm;
        const int nc = (int)cand.size();
        for (int cc = 0; cc < nc; ++cc) {
            const int c = reverse_scan ? nc - 1 - cc : cc;
            if (cdone[c]) continue;
            const Gate& q = g[cand[c]];
            const uint64_t bi = 1ull << q.i, bj = 1ull << q.j, m = bi | bj;
            const bool cross = ((bi & lm) && (bj & rm)) || ((bi & rm) && (bj & lm));
            const bool fits = cross && !(m & blocked) && (sweep < 0 || q.sweep == sweep);
            if (fits) {
                out.push_back(c);
                sweep = q.sweep;
            } else {
                blocked |= m;
            }
            if (popc(both & ~blocked) < 2) break;
        }
    };
    // (thread-local: the passes of a program are decomposed by several host threads, each with its own search state)
    static thread_local int cur_unb_kind;  // left-group size of the unbalanced split rounds the searches may use right now (0: none)
    cur_unb_kind = 0;
    auto search_split = [&](const RawPass& rp, const std::vector<Gate>& g, const std::vector<char>& cdone, bool reverse_scan,
                            uint64_t limit_mask, int limit, uint64_t prefer, RawRound& best) {
        std::vector<int> tq;
        for (int q = 0; q < nq; ++q)
            if ((rp.tmask >> q) & 1) tq.push_back(q);
        const int tn = (int)tq.size();
        best.rmask = 0;
        best.lmask = 0;
        best.cidx.clear();
        best.split = true;
        if (tn < RS) return;
        // left-group sizes: RS/2 | RS/2 always, plus the unbalanced bipartition of the kind being tried (2 | RS-2 or 1 | RS-1)
        const int Hunb = cfg.unb_last ? cur_unb_kind : 0;  // search_split serves the last rounds
        std::vector<int> idx(RS), ex;
        for (int k = 0; k < RS; ++k) idx[k] = k;
        int best_pref = -1;
        while (true) {
            uint64_t sm = 0;
            for (int k = 0; k < RS; ++k) sm |= 1ull << tq[idx[k]];
            if (popc(sm & limit_mask) <= limit) {
                for (int H : {RS / 2, Hunb}) {
                    if (H <= 0) continue;
                    std::vector<int> sub(H);
                    for (int k = 0; k < H; ++k) sub[k] = k;
                    while (true) {
                        // balanced split: left groups containing the smallest element only (the mirrored split runs the same gates)
                        if (!(2 * H == RS && sub[0] != 0)) {
                            uint64_t lm = 0;
                            for (int k = 0; k < H; ++k) lm |= 1ull << tq[idx[sub[k]]];
                            split_closure(g, rp.gate_idx, cdone, lm, sm & ~lm, reverse_scan, ex);
                            const int pref = popc(sm & prefer);
                            if (ex.size() > best.cidx.size() || (ex.size() == best.cidx.size() && !ex.empty() && pref > best_pref)) {
                                best.cidx = ex;
                                best.rmask = sm;
                                best.lmask = lm;
                                best_pref = pref;
                            }
                        }
                        int k = H - 1;
                        while (k >= 0 && sub[k] == RS - H + k) --k;
                        if (k < 0) break;
                        ++sub[k];
                        for (int m = k + 1; m < H; ++m) sub[m] = sub[m - 1] + 1;
                    }
                }
            }
            int k = RS - 1;
            while (k >= 0 && idx[k] == tn - RS + k) --k;
            if (k < 0) break;
            ++idx[k];
            for (int m = k + 1; m < RS; ++m) idx[m] = idx[m - 1] + 1;
        }
        if (!best.cidx.empty()) {
            best.reverse = g[rp.gate_idx[best.cidx[0]]].reverse;
            std::sort(best.cidx.begin(), best.cidx.end());
        }
    };
    auto search_complex = [&](const RawPass& rp, const std::vector<Gate>& g, const std::vector<char>& cdone,
                            bool reverse_scan, uint64_t limit_mask, int limit, uint64_t prefer, RawRound& best) {
        std::vector<int> tq;
        for (int q = 0; q < nq; ++q)
            if ((rp.tmask >> q) & 1) tq.push_back(q);
        const int tn = (int)tq.size();
        std::vector<int> idx(RB), ex;
        for (int k = 0; k < RB; ++k) idx[k] = k;
        best.rmask = 0;
        best.cidx.clear();
        int best_pref = -1;
        while (true) {
            uint64_t rm = 0;
            for (int k = 0; k < RB; ++k) rm |= 1ull << tq[idx[k]];
            if (popc(rm & limit_mask) <= limit) {
                if (reverse_scan) rev_round_closure(g, rp.gate_idx, cdone, rm, ex);
                else round_closure(g, rp.gate_idx, cdone, rm, ex);
                const int pref = popc(rm & prefer);
                if (ex.size() > best.cidx.size() || (ex.size() == best.cidx.size() && !ex.empty() && pref > best_pref)) {
                    best.cidx = ex;
                    best.rmask = rm;
                    best_pref = pref;
                }
            }
            int k = RB - 1;
            while (k >= 0 && idx[k] == tn - RB + k) --k;
            if (k < 0) break;
            ++idx[k];
            for (int m = k + 1; m < RB; ++m) idx[m] = idx[m - 1] + 1;
        }
        if (!best.cidx.empty()) {
            best.reverse = g[rp.gate_idx[best.cidx[0]]].reverse;
            std::sort(best.cidx.begin(), best.cidx.end());
        }
    };
    // the better of a complex round and (when allowed) a split round; ties go to the complex round
    // (half as many shared-memory instructions for the same bytes)
    static thread_local bool allow_split;
    allow_split = can_split;
    auto search_round = [&](const RawPass& rp, const std::vector<Gate>& g, const std::vector<char>& cdone,
                            bool reverse_scan, uint64_t limit_mask, int limit, uint64_t prefer, RawRound& best) {
        search_complex(rp, g, cdone, reverse_scan, limit_mask, limit, prefer, best);
        if (!allow_split) return;
        RawRound sp;
        search_split(rp, g, cdone, reverse_scan, limit_mask, limit, prefer, sp);
        if ((int)sp.cidx.size() >= (int)best.cidx.size() + cfg.split_margin) best = sp;
    };

    // top-K candidate rounds (distinct gate sets) for a beam search; score = 2 * gates - split (ties go to the complex round)
    struct Pool {
        size_t K;
        std::vector<RawRound> v;
        static int score(const RawRound& r) { return 2 * (int)r.cidx.size() - (r.split ? 1 : 0); }
        int floor_score() const { return v.size() < K ? 0 : score(v.back()); }
        void add(const RawRound& r) {
            if (r.cidx.empty()) return;
            const int sc = score(r);
            if (v.size() >= K && sc <= score(v.back())) return;
            std::vector<int> key(r.cidx);
            std::sort(key.begin(), key.end());
            for (auto& o : v) {
                std::vector<int> ok(o.cidx);
                std::sort(ok.begin(), ok.end());
                if (ok == key) {
                    if (sc > score(o)) o = r;
                    std::sort(v.begin(), v.end(), [](const RawRound& x, const RawRound& y) { return score(x) > score(y); });
                    return;
                }
            }
            v.push_back(r);
            std::sort(v.begin(), v.end(), [](const RawRound& x, const RawRound& y) { return score(x) > score(y); });
            if (v.size() > K) v.pop_back();
        }
    };
    auto collect_rounds = [&](const RawPass& rp, const std::vector<Gate>& g, const std::vector<char>& cd, bool with_split, bool with_unb, Pool& pool) {
        std::vector<int> tq;
        for (int q = 0; q < nq; ++q)
            if ((rp.tmask >> q) & 1) tq.push_back(q);
        const int tn = (int)tq.size();
        std::vector<int> ex;
        auto finish = [&](RawRound& r) {
            r.reverse = g[rp.gate_idx[r.cidx[0]]].reverse;
            std::sort(r.cidx.begin(), r.cidx.end());
        };
        // upper bound of what a register set can run: the undone gates it contains (prunes most closures once the
        // pool holds K good candidates; a pruned set could not have entered the pool, so the result is unchanged)
        std::vector<uint64_t> undone;
        for (size_t c = 0; c < rp.gate_idx.size(); ++c)
            if (!cd[c]) undone.push_back((1ull << g[rp.gate_idx[c]].i) | (1ull << g[rp.gate_idx[c]].j));
        auto contained = [&](uint64_t set) {
            int cnt = 0;
            for (uint64_t m : undone) cnt += (m & set) == m;
            return cnt;
        };
        {
            const int nr = std::min(RB, tn);
            std::vector<int> idx(nr);
            for (int k = 0; k < nr; ++k) idx[k] = k;
            while (true) {
                uint64_t rm = 0;
                for (int k = 0; k < nr; ++k) rm |= 1ull << tq[idx[k]];
                if (2 * contained(rm) > pool.floor_score()) round_closure(g, rp.gate_idx, cd, rm, ex);
                else ex.clear();
                if (2 * (int)ex.size() > pool.floor_score()) {
                    RawRound r;
                    r.rmask = rm;
                    r.cidx = ex;
                    finish(r);
                    pool.add(r);
                }
                int k = nr - 1;
                while (k >= 0 && idx[k] == tn - nr + k) --k;
                if (k < 0) break;
                ++idx[k];
                for (int m = k + 1; m < nr; ++m) idx[m] = idx[m - 1] + 1;
            }
        }
        if (!with_split || tn < RS) return;
        const int Hunb = with_unb ? cur_unb_kind : 0;
        std::vector<int> idx(RS);
        for (int k = 0; k < RS; ++k) idx[k] = k;
        while (true) {
            uint64_t sm = 0;
            for (int k = 0; k < RS; ++k) sm |= 1ull << tq[idx[k]];
            const int bound = 2 * std::min(contained(sm), (RS / 2) * (RS - RS / 2)) - 1;
            for (int H : {RS / 2, Hunb}) {
                if (H <= 0 || bound <= pool.floor_score()) continue;
                std::vector<int> sub(H);
                for (int k = 0; k < H; ++k) sub[k] = k;
                while (true) {
                    if (!(2 * H == RS && sub[0] != 0)) {
                        uint64_t lm = 0;
                        for (int k = 0; k < H; ++k) lm |= 1ull << tq[idx[sub[k]]];
                        split_closure(g, rp.gate_idx, cd, lm, sm & ~lm, false, ex);
                        if (2 * (int)ex.size() - 1 > pool.floor_score()) {
                            RawRound r;
                            r.rmask = sm;
                            r.lmask = lm;
                            r.split = true;
                            r.cidx = ex;
                            finish(r);
                            pool.add(r);
                        }
                    }
                    int k = H - 1;
                    while (k >= 0 && sub[k] == RS - H + k) --k;
                    if (k < 0) break;
                    ++sub[k];
                    for (int m = k + 1; m < H; ++m) sub[m] = sub[m - 1] + 1;
                }
            }
            int k = RS - 1;
            while (k >= 0 && idx[k] == tn - RS + k) --k;
            if (k < 0) break;
            ++idx[k];
            for (int m = k + 1; m < RS; ++m) idx[m] = idx[m - 1] + 1;
        }
    };

    std::vector<std::vector<char>> cdone(P);
    std::vector<RawRound> last_round(P), first_round(P);
    std::vector<char> single_round(P, 0);
    for (int p = 0; p < P; ++p) cdone[p].assign(raw[p].gate_idx.size(), 0);
    // identical passes (same tile, same gate pattern, same store constraint) recur in every sweep of a
    // high-order formula: their round decomposition is computed once
    struct RoundsMemo {
        RawRound last;
        std::vector<RawRound> front;
    };
    static std::map<std::string, RoundsMemo> memo;  // process-wide, like the tile memo (the caller holds its lock)
    if (memo.size() > 16384) memo.clear();
    auto memo_key = [&](int p) {
        const RawPass& rp = raw[p];
        const std::vector<Gate>& g = sg[rp.seg];
        std::ostringstream os;
        const uint64_t shared_next = rp.tmask & next_tile(p);
        os << cfg_sig << rp.tmask << ':' << shared_next << ':' << (rp.has_phase ? 1 : 0) << ':';
        const int s0 = rp.gate_idx.empty() ? 0 : g[rp.gate_idx[0]].sweep;
        for (int k : rp.gate_idx) os << g[k].i << ',' << g[k].j << ',' << (g[k].reverse ? 1 : 0) << ',' << (g[k].sweep - s0) << ';';
        return os.str();
    };
    // The passes of a program are independent here (a pass only looks at its own gates and at the neighbouring TILES):
    // those whose decomposition is not memoised yet -- one per distinct key -- are searched by a small pool of host threads,
    // the others are filled in from the memo afterwards.  The result does not depend on the thread count or timing: a memo hit
    // returns exactly what the search computes.
    std::vector<std::string> keys(P);
    for (int p = 0; p < P; ++p) keys[p] = memo_key(p);
    std::vector<int> todo;
    std::vector<char> computed(P, 0);
    {
        std::set<std::string> seen;
        for (int p = 0; p < P; ++p)
            if (memo.find(keys[p]) == memo.end() && seen.insert(keys[p]).second) {
                todo.push_back(p);
                computed[p] = 1;
            }
    }
    auto decompose_pass = [&](int p) {
        cur_unb_kind = 0;
        allow_split = can_split;
        const RawPass& rp = raw[p];
        const std::vector<Gate>& g = sg[rp.seg];
        // one decomposition per admissible kind of unbalanced split round (none / 2 | 4 / 1 | 5: a pass never mixes the
        // two, each lives in its own kernel instantiation); the fewest rounds win, ties go to the leaner kernel
        auto decompose = [&]() {
            cdone[p].assign(rp.gate_idx.size(), 0);
            rr[p].clear();
            const uint64_t shared_next = rp.tmask & next_tile(p);
            const int s = popc(shared_next);
            RawRound best;
            // fused store: the c_lane lowest output bits must carry shared qubits that this round keeps on lanes;
            // every other shared qubit may sit in registers (the next pass's bulk tile load does not care)
            if (cfg.fuse_io && s - c_lane >= 0) search_round(rp, g, cdone[p], true, shared_next, s - c_lane, 0, best);
            if (best.cidx.empty()) search_round(rp, g, cdone[p], true, 0, 99, 0, best);
            if (best.cidx.empty()) throw std::runtime_error("compile_plan: last-round search made no progress");
            if (best.split && best.cidx.size() == rp.gate_idx.size() && rp.has_phase) {
                // a single round that also carries the diagonal phase must hold whole complex amplitudes
                allow_split = false;
                best = RawRound();
                if (cfg.fuse_io && s - c_lane >= 0) search_round(rp, g, cdone[p], true, shared_next, s - c_lane, 0, best);
                if (best.cidx.empty()) search_round(rp, g, cdone[p], true, 0, 99, 0, best);
                allow_split = can_split;
            }
            for (int c : best.cidx) cdone[p][c] = 1;
            rlast[p] = best.rmask;
            last_round[p] = best;

            // (b) first and middle rounds: beam search for the fewest rounds (breadth first over the round count,
            // cfg.beam_width states kept per depth, cfg.beam_cands candidate rounds expanded per state)
            size_t ndone0 = best.cidx.size();
            if (ndone0 == rp.gate_idx.size()) {
                rr[p].clear();
                return;
            }
            struct State {
                std::vector<char> cd;
                std::vector<RawRound> rounds;
                size_t ndone;
                int nsplit;
            };
            std::vector<State> beam;
            beam.push_back(State{cdone[p], {}, ndone0, 0});
            const size_t total = rp.gate_idx.size();
            const int W = std::max(1, cfg.beam_width), KC = std::max(1, cfg.beam_cands);
            bool finished = false;
            State winner;
            for (int depth = 0; !finished; ++depth) {
                if (depth > 64) throw std::runtime_error("compile_plan: round search does not terminate");
                std::vector<State> next;
                for (const State& st : beam) {
                    Pool pool{(size_t)KC, {}};
                    const bool with_split = can_split && !(depth == 0 && rp.has_phase);  // the fused phase multiplies whole complex amplitudes
                    collect_rounds(rp, g, st.cd, with_split, depth > 0 || cfg.unb_first, pool);
                    if (pool.v.empty()) throw std::runtime_error("compile_plan: round search made no progress");
                    for (const RawRound& r : pool.v) {
                        State ns = st;
                        for (int c : r.cidx) ns.cd[c] = 1;
                        ns.ndone += r.cidx.size();
                        ns.nsplit += r.split ? 1 : 0;
                        ns.rounds.push_back(r);
                        next.push_back(std::move(ns));
                    }
                }
                std::stable_sort(next.begin(), next.end(), [](const State& x, const State& y) {
                    if (x.ndone != y.ndone) return x.ndone > y.ndone;
                    return x.nsplit < y.nsplit;
                });
                if (next.front().ndone == total) {
                    winner = next.front();
                    finished = true;
                    break;
                }
                beam.clear();
                for (State& ns : next) {
                    bool dup = false;
                    for (const State& o : beam)
                        if (o.cd == ns.cd) dup = true;
                    if (!dup) beam.push_back(std::move(ns));
                    if ((int)beam.size() >= W) break;
                }
            }
            rr[p] = winner.rounds;
            cdone[p] = winner.cd;
            rfirst[p] = rr[p][0].rmask;
            first_round[p] = rr[p][0];
        };
        std::vector<int> kinds{0};
        if (can_split && cfg.split_min_left <= 2 && RS / 2 > 2) kinds.push_back(2);
        if (can_split && cfg.split_min_left <= 1 && RS / 2 > 1) kinds.push_back(1);
        RoundsMemo best_dec;
        size_t best_n = 0;
        for (int kind : kinds) {
            cur_unb_kind = kind;
            decompose();
            const size_t nrd = rr[p].size() + 1;
            if (best_n == 0 || nrd < best_n) {
                best_n = nrd;
                best_dec = RoundsMemo{last_round[p], rr[p]};
            }
        }
        cur_unb_kind = 0;
        last_round[p] = best_dec.last;
        rr[p] = best_dec.front;
        rlast[p] = last_round[p].rmask;
        cdone[p].assign(rp.gate_idx.size(), 1);
        if (rr[p].empty()) {
            single_round[p] = 1;
            rfirst[p] = rlast[p];
        } else {
            rfirst[p] = rr[p][0].rmask;
            first_round[p] = rr[p][0];
        }
    };
    {
        unsigned nthreads = std::min<unsigned>({std::max(1u, std::thread::hardware_concurrency()), 8u, (unsigned)todo.size()});
        if (nthreads <= 1) {
            for (int p : todo) decompose_pass(p);
        } else {
            std::atomic<size_t> next_item{0};
            std::vector<std::exception_ptr> errors(nthreads);
            std::vector<std::thread> pool;
            for (unsigned t = 0; t < nthreads; ++t)
                pool.emplace_back([&, t]() {
                    try {
                        for (size_t k = next_item++; k < todo.size(); k = next_item++) decompose_pass(todo[k]);
                    } catch (...) {
                        errors[t] = std::current_exception();
                    }
                });
            for (auto& th : pool) th.join();
            for (auto& e : errors)
                if (e) std::rethrow_exception(e);
        }
    }
    for (int p : todo) memo[keys[p]] = RoundsMemo{last_round[p], rr[p]};
    for (int p = 0; p < P; ++p) {
        if (computed[p]) continue;
        const auto it = memo.find(keys[p]);
        if (it == memo.end()) throw std::runtime_error("compile_plan: round memo lost an entry");
        last_round[p] = it->second.last;
        rr[p] = it->second.front;
        rlast[p] = last_round[p].rmask;
        for (int c : last_round[p].cidx) cdone[p][c] = 1;
        for (auto& r : rr[p])
            for (int c : r.cidx) cdone[p][c] = 1;
        if (rr[p].empty()) {
            single_round[p] = 1;
            rfirst[p] = rlast[p];
        } else {
            rfirst[p] = rr[p][0].rmask;
            first_round[p] = rr[p][0];
        }
    }
    for (int p = 0; p < P; ++p) rr[p].push_back(last_round[p]);
    }

    // ---- 3. layouts: tile qubits of pass p sit on physical bits [0, TB) before pass p ----
    // order: qubits shared with the previous tile that neither the previous last round nor this
    // first round holds in registers (-> lanes of the fused global I/O), then the other shared
    // qubits, then the remaining tile qubits, then everything else; ascending inside each class.
    // `prev_layout` (positions of the qubits before the previous pass, empty if unknown): the three qubits that land on
    // output bits 0..2 of the previous pass are the low lane bits of its last round's fused store, and that round READS them
    // from shared-memory positions prev_layout[q]; three different bank classes there make those reads conflict free.
    void assign_layouts() {
    auto make_layout = [&](int p, std::vector<int>& perm, int& nshared, int& nfree, const std::vector<int>& prev_layout) {
        const uint64_t tile = raw[p].tmask;
        const int pm = prev_index(p);
        const bool hinted = pm < 0 && !cfg.cyclic && cfg.prev_tile_hint;
        const uint64_t prev = pm < 0 ? (hinted ? cfg.prev_tile_hint : 0) : raw[pm].tmask;
        const uint64_t rl_prev = pm < 0 ? (hinted ? cfg.prev_rlast_hint : 0) : rlast[pm];
        perm.assign(nq, -1);
        int pos = 0;
        // lowest: shared qubits the previous last round keeps on lanes (coalesced fused store; those that
        // this first round also keeps on lanes first), then the other shared qubits, then the rest of the
        // tile with the first round's THREAD qubits last: they land on the top tile positions, the second
        // set of conflict-free lane positions of the arrival layout
        const uint64_t shared = tile & prev;
        nshared = popc(shared);
        nfree = popc(shared & ~rl_prev);
        const uint64_t rest = tile & ~shared;
        const uint64_t classes[6] = {shared & ~rl_prev & ~rfirst[p], shared & ~rl_prev & rfirst[p], shared & rl_prev & rfirst[p],
                                     shared & rl_prev & ~rfirst[p], rest & rfirst[p], rest & ~rfirst[p]};
        uint64_t placed = 0;
        if (!prev_layout.empty() && pm >= 0 && c_lane == 3) {
            // candidates in preference order (class 0 first), pick the first triple of pairwise different bank classes
            std::vector<int> cand;
            for (int c = 0; c < 2; ++c)
                for (int q = 0; q < nq; ++q)
                    if ((classes[c] >> q) & 1) cand.push_back(q);
            const int nc = (int)cand.size();
            int best[3] = {-1, -1, -1};
            for (int a = 0; a < nc && best[0] < 0; ++a)
                for (int b = a + 1; b < nc && best[0] < 0; ++b)
                    for (int c = b + 1; c < nc && best[0] < 0; ++c) {
                        const int ca = lane_class(prev_layout[cand[a]]), cb = lane_class(prev_layout[cand[b]]), cc = lane_class(prev_layout[cand[c]]);
                        if (ca != cb && ca != cc && cb != cc) {
                            best[0] = cand[a];
                            best[1] = cand[b];
                            best[2] = cand[c];
                        }
                    }
            if (best[0] >= 0)
                for (int k = 0; k < 3; ++k) {
                    perm[best[k]] = pos++;
                    placed |= 1ull << best[k];
                }
        }
        for (uint64_t cls : classes)
            for (int q = 0; q < nq; ++q)
                if (((cls & ~placed) >> q) & 1) perm[q] = pos++;
        for (int q = 0; q < nq; ++q)
            if (!((tile >> q) & 1)) perm[q] = pos++;
    };
    layouts.assign(P + 1, {});
    nshared.assign(P + 1, 0);
    nfree.assign(P + 1, 0);
    {
        const std::vector<int> none;
        for (int p = 0; p < P; ++p) make_layout(p, layouts[p], nshared[p], nfree[p], p > 0 ? layouts[p - 1] : none);
        if (cfg.cyclic && P > 1)  // the first layout follows the last pass: second sweep with the wrap-around known
            for (int p = 0; p < P; ++p) make_layout(p, layouts[p], nshared[p], nfree[p], layouts[p > 0 ? p - 1 : P - 1]);
    }
    if (P > 0 && cfg.cyclic) {
        layouts[P] = layouts[0];
        nshared[P] = nshared[0];
        nfree[P] = nfree[0];
        plan.start_perm = layouts[0];
        plan.end_perm = layouts[0];
    } else if (P > 0 && !cfg.final_positions.empty()) {
        // open program chained to the next one: the caller dictates where every qubit lands
        layouts[P] = cfg.final_positions;
        int lowrun = 0;  // lowest output bits occupied by qubits of the last tile
        for (int b = 0; b < nq; ++b) {
            bool hit = false;
            for (int q = 0; q < nq; ++q)
                if (cfg.final_positions[q] == b && ((raw[P - 1].tmask >> q) & 1)) hit = true;
            if (!hit) break;
            ++lowrun;
        }
        nshared[P] = lowrun;
        nfree[P] = lowrun;
        plan.start_perm = layouts[0];
        plan.end_perm = layouts[P];
    } else if (P > 0) {
        // open program: the last pass keeps its tile qubits low (those its last round does not hold in
        // registers first, so that the store stays coalesced) and parks `final_top` on the highest bits
        uint64_t top = 0;
        for (int q : cfg.final_top) top |= 1ull << q;
        const uint64_t tile = raw[P - 1].tmask & ~top;
        std::vector<int>& perm = layouts[P];
        perm.assign(nq, -1);
        int pos = 0;
        const uint64_t classes[2] = {tile & ~rlast[P - 1], tile & rlast[P - 1]};
        for (uint64_t cls : classes)
            for (int q = 0; q < nq; ++q)
                if ((cls >> q) & 1) perm[q] = pos++;
        for (int q = 0; q < nq; ++q)
            if (!((tile >> q) & 1) && !((top >> q) & 1)) perm[q] = pos++;
        for (int q : cfg.final_top) perm[q] = pos++;
        nshared[P] = popc(tile);
        nfree[P] = popc(tile & ~rlast[P - 1]);
        plan.start_perm = layouts[0];
        plan.end_perm = layouts[P];
    } else {
        plan.start_perm.resize(nq);
        for (int q = 0; q < nq; ++q) plan.start_perm[q] = q;
        if (!cfg.cyclic && !cfg.final_top.empty()) {
            uint64_t top = 0;
            for (int q : cfg.final_top) top |= 1ull << q;
            int pos = 0;
            for (int q = 0; q < nq; ++q)
                if (!((top >> q) & 1)) plan.start_perm[q] = pos++;
            for (int q : cfg.final_top) plan.start_perm[q] = pos++;
        }
        plan.end_perm = plan.start_perm;
    }

    if (P > 0) {
        plan.first_tile_mask = raw[0].tmask;
        plan.last_tile_mask = raw[P - 1].tmask;
        plan.last_rlast_mask = rlast[P - 1];
    }
    }

    // ---- 4. emit passes ----
    void emit_passes() {
    for (int p = 0; p < P; ++p) {
        const RawPass& rp = raw[p];
        const std::vector<Gate>& g = sg[rp.seg];
        const std::vector<GateOp>& gop = segs[rp.seg].gates;
        PlanPass pp;
        std::memset(pp.tileq, -1, sizeof(pp.tileq));
        const std::vector<int>& lay = layouts[p];
        int qpos[64];
        std::vector<int> inv(nq);
        for (int q = 0; q < nq; ++q) {
            qpos[q] = lay[q];
            inv[lay[q]] = q;
            if (lay[q] < TB) pp.tileq[lay[q]] = q;
        }
        for (int b = 0; b < kMaxQ; ++b) pp.out_of_in[b] = b;
        for (int b = 0; b < nq; ++b) pp.out_of_in[b] = layouts[p + 1][inv[b]];
        pp.n_shared = (P == 1 && cfg.cyclic) ? TB : nshared[p + 1];
        pp.has_phase = rp.has_phase && zeff[rp.seg] != 0.0;
        pp.ztau = pp.has_phase ? zeff[rp.seg] : 0.0;
        pp.std_rot = false;
        pp.ngates = (int)rp.gate_idx.size();
        const int nrounds = (int)rr[p].size();
        // fused global I/O is possible when the lowest c_lane physical bits (before / after the
        // pass) carry qubits that the first / last round does not hold in registers
        bool fl = cfg.fuse_io && nq > TB, fs = cfg.fuse_io && nq > TB;
        // bulk tile load: one thread (non-register) qubit of round 0 on each conflict-free position pair {b, 9 + b}
        int load_lane[3] = {-1, -1, -1};
        for (int b = 0; b < c_lane && b < 3 && fl; ++b) {
            if (!((rfirst[p] >> inv[b]) & 1)) load_lane[b] = b;
            else if (9 + b < TB && !((rfirst[p] >> inv[9 + b]) & 1)) load_lane[b] = 9 + b;
            else fl = false;
        }
        {
            std::vector<int> inv_next(kMaxQ + 8, -1);
            for (int q = 0; q < nq; ++q) inv_next[layouts[p + 1][q]] = q;
            for (int b = 0; b < c_lane && fs; ++b) {
                const int q = inv_next[b];
                fs = q >= 0 && ((rp.tmask >> q) & 1) && !((rlast[p] >> q) & 1);
            }
        }
        if (nrounds == 1 && fl) fs = false;  // one round cannot serve both lane mappings
        pp.fuse_load = fl;
        pp.fuse_store = fs;

        // "warp qubits": the two tile qubits mapped to the warp-index bits of a round.  While
        // they stay the same (and out of the register sets) over consecutive rounds, each warp
        // exchanges data only with itself through shared memory and a warp-level barrier
        // suffices; the block-wide barrier is needed only where they change.
        const int NTB = TB - RB;                 // thread bits
        const int nwarpbits = NTB > 5 ? NTB - 5 : 0;
        std::vector<uint64_t> wsel(nrounds, 0);
        std::vector<uint64_t> forced(nrounds, 0);  // qubits that must be LANE bits in that round
        {
            if (fl)
                for (int b = 0; b < c_lane && b < 3; ++b) forced[0] |= 1ull << inv[load_lane[b]];
            if (fs) {
                // lanes of the last round follow ascending output position: the 5 lowest
                std::vector<int> cand;
                for (int q = 0; q < nq; ++q)
                    if (((rp.tmask >> q) & 1) && !((rr[p][nrounds - 1].rmask >> q) & 1)) cand.push_back(q);
                std::sort(cand.begin(), cand.end(), [&](int a, int b) { return layouts[p + 1][a] < layouts[p + 1][b]; });
                for (int k = 0; k < (int)cand.size() - nwarpbits; ++k) forced[nrounds - 1] |= 1ull << cand[k];
            }
            // a choice W of warp qubits is good for round r when the lanes it leaves (tile minus registers minus W) still
            // cover the three bank classes of the padded tile layout; rounds whose lane order is dictated by the fused
            // global load / store are exempt
            auto lanes_ok = [&](int r, uint64_t W) {
                if ((r == 0 && fl) || (r == nrounds - 1 && fs)) return true;
                int seen = 0;
                for (int q = 0; q < nq; ++q)
                    if (((rp.tmask & ~rr[p][r].rmask & ~W) >> q) & 1) seen |= 1 << lane_class(qpos[q]);
                return seen == 7;
            };
            auto pick_w = [&](uint64_t X, int ra, int rb, bool& good) -> uint64_t {
                std::vector<int> xs;
                for (int q = nq - 1; q >= 0; --q)
                    if ((X >> q) & 1) xs.push_back(q);  // highest-numbered first (deterministic)
                uint64_t fallback = 0;
                for (int q : xs)
                    if (popc(fallback) < nwarpbits) fallback |= 1ull << q;
                good = false;
                if (nwarpbits != 2) return fallback;
                for (size_t a = 0; a < xs.size(); ++a)
                    for (size_t b = a + 1; b < xs.size(); ++b) {
                        const uint64_t W = (1ull << xs[a]) | (1ull << xs[b]);
                        bool ok = true;
                        for (int r = ra; r <= rb && ok; ++r) ok = lanes_ok(r, W);
                        if (ok) {
                            good = true;
                            return W;
                        }
                    }
                return fallback;
            };
            int r0 = 0;
            while (r0 < nrounds && nwarpbits > 0) {
                uint64_t X = rp.tmask & ~rr[p][r0].rmask & ~forced[r0];
                bool good = false;
                uint64_t W = pick_w(X, r0, r0, good);
                int r1 = r0;
                while (r1 + 1 < nrounds) {
                    const uint64_t X2 = X & ~rr[p][r1 + 1].rmask & ~forced[r1 + 1];
                    if (popc(X2) < nwarpbits) break;
                    bool good2 = false;
                    const uint64_t W2 = pick_w(X2, r0, r1 + 1, good2);
                    if (good && !good2) break;  // extending the run would cost a bank conflict: take the block-wide barrier instead
                    X = X2;
                    W = W2;
                    good = good2;
                    ++r1;
                }
                for (int r = r0; r <= r1; ++r) wsel[r] = W;
                r0 = r1 + 1;
            }
        }

        for (int r = 0; r < nrounds; ++r) {
            const RawRound& R = rr[p][r];
            PlanRound pr;
            std::memset(&pr, 0, sizeof(pr));
            std::vector<int> rq;
            const int nreg = R.split ? RS : RB;
            pr.split = R.split ? 1 : 0;
            pr.nreg = nreg;
            if (R.split) {
                std::vector<int> lq, rgt;
                for (int q = 0; q < nq; ++q)
                    if ((R.rmask >> q) & 1) (((R.lmask >> q) & 1) ? lq : rgt).push_back(q);
                if (R.reverse) {
                    std::reverse(lq.begin(), lq.end());
                    std::reverse(rgt.begin(), rgt.end());
                }
                rq = lq;
                rq.insert(rq.end(), rgt.begin(), rgt.end());
                pr.nleft = (int)lq.size();
            } else {
                for (int q = 0; q < nq; ++q)
                    if ((R.rmask >> q) & 1) rq.push_back(q);
                if (R.reverse) std::reverse(rq.begin(), rq.end());
            }
            for (int k = 0; k < nreg; ++k) {
                pr.regq[k] = rq[k];
                pr.regpos[k] = qpos[rq[k]];
            }
            // lane positions (everything that is neither a register nor a warp qubit), warp positions
            std::vector<int> lanepos, warppos;
            for (int pos = 0; pos < TB; ++pos) {
                bool isreg = false;
                for (int k = 0; k < nreg; ++k) isreg |= (pr.regpos[k] == pos);
                if (isreg) continue;
                if ((wsel[r] >> inv[pos]) & 1) warppos.push_back(pos);
                else lanepos.push_back(pos);
            }
            const int nl = (int)lanepos.size();
            std::vector<int> order;
            if (r == 0 && fl) {
                // the three conflict-free positions of the arrival layout first, the other lanes ascending
                for (int b = 0; b < c_lane && b < 3; ++b) order.push_back(load_lane[b]);
                for (int lp : lanepos)
                    if (std::find(order.begin(), order.end(), lp) == order.end()) order.push_back(lp);
            } else if (r == nrounds - 1 && fs) {
                order = lanepos;  // ascending OUTPUT position: lanes 0.. write contiguous addresses
                std::sort(order.begin(), order.end(), [&](int a, int b) { return pp.out_of_in[a] < pp.out_of_in[b]; });
            } else {
                // three positions of three different bank classes first (bank-conflict free)
                int pick[3] = {-1, -1, -1};
                bool found = false;
                for (int a = 0; a < nl && !found; ++a)
                    for (int b = a + 1; b < nl && !found; ++b)
                        for (int c = b + 1; c < nl && !found; ++c) {
                            const int va = lane_class(lanepos[a]), vb = lane_class(lanepos[b]), vc = lane_class(lanepos[c]);
                            if (va != vb && va != vc && vb != vc) {
                                pick[0] = a; pick[1] = b; pick[2] = c;
                                found = true;
                            }
                        }
                std::vector<char> used(nl, 0);
                if (found)
                    for (int k = 0; k < 3; ++k) {
                        order.push_back(lanepos[pick[k]]);
                        used[pick[k]] = 1;
                    }
                for (int a = 0; a < nl; ++a)
                    if (!used[a]) order.push_back(lanepos[a]);
            }
            for (int wpos : warppos) order.push_back(wpos);
            const int nthr = (int)order.size();
            for (int b = 0; b < nthr; ++b) pr.thrpos[b] = order[b];
            // barrier kind after this round's exchange
            pr.cta_sync_after = 1;
            if (nwarpbits > 0 && r + 1 < nrounds && wsel[r + 1] == wsel[r]) pr.cta_sync_after = 0;
            for (int c : R.cidx) {
                const Gate& q = g[rp.gate_idx[c]];
                int a = -1, b = -1;
                for (int k = 0; k < nreg; ++k) {
                    if (pr.regq[k] == q.i || pr.regq[k] == q.j) {
                        if (a < 0) a = k;
                        else b = k;
                    }
                }
                int slot = 0;
                if (R.split) {
                    if (a < 0 || b < pr.nleft || a >= pr.nleft) throw std::runtime_error("compile_plan: split round gate inside one group");
                    slot = (RS - pr.nleft) * a + (b - pr.nleft);
                } else {
                    for (int x = 0; x < a; ++x) slot += RB - 1 - x;
                    slot += b - a - 1;
                }
                if ((pr.mask >> slot) & 1u) throw std::runtime_error("compile_plan: pair scheduled twice in one round");
                pr.mask |= 1u << slot;
                pr.phi[slot] = gop[rp.gate_idx[c]].phi;
            }
            pr.ngates = (int)R.cidx.size();
            pp.rounds.push_back(pr);
        }
        plan.total_gates += pp.ngates;
        plan.total_rounds += (int)pp.rounds.size();
        plan.passes.push_back(std::move(pp));
    }
    // split passes that carry more rounds than one descriptor can hold
    {
        std::vector<PlanPass> out;
        for (auto& pp : plan.passes) {
            if ((int)pp.rounds.size() <= cfg.max_rounds && pp.ngates <= cfg.max_pass_gates) {
                out.push_back(std::move(pp));
                continue;
            }
            // identity-layout continuation passes: same tile, no permutation until the last piece
            size_t off = 0;
            const size_t total = pp.rounds.size();
            while (off < total) {
                size_t cnt = 0;  // rounds of this piece: at most max_rounds, at most max_pass_gates gates (the kernel's coefficient list)
                for (int g = 0; cnt < std::min((size_t)cfg.max_rounds, total - off); ++cnt) {
                    g += pp.rounds[off + cnt].ngates;
                    if (g > cfg.max_pass_gates && cnt > 0) break;
                }
                PlanPass piece = pp;
                piece.rounds.assign(pp.rounds.begin() + off, pp.rounds.begin() + off + cnt);
                // the piece's store reads the whole tile out of shared memory: block-wide barrier after its last
                // round, whatever the barrier towards the next round of the unsplit pass was
                piece.rounds.back().cta_sync_after = 1;
                piece.has_phase = pp.has_phase && off == 0;
                piece.fuse_load = pp.fuse_load && off == 0;
                piece.fuse_store = pp.fuse_store && off + cnt == total;
                if (off + cnt < total) {
                    for (int b = 0; b < kMaxQ; ++b) piece.out_of_in[b] = b;
                    piece.n_shared = TB;
                }
                piece.ngates = 0;
                for (auto& r : piece.rounds) piece.ngates += r.ngates;
                out.push_back(std::move(piece));
                off += cnt;
            }
        }
        plan.passes.swap(out);
    }
    }
};

}  // namespace

// Intermediate results of the first `upto` stages (1: tiles, 2: + rounds, 3: + layouts) as JSON: the unit-test view of the
// schedule compiler (tests/test_schedule_stages_cpu.py).
std::string plan_stages_to_json(int nq, const std::vector<Segment>& segs, const SchedConfig& cfg, int upto) {
    std::lock_guard<std::mutex> memo_lock(g_memo_mutex);
    PlanCompiler pc(nq, segs, cfg);
    pc.prepare();
    pc.choose_tiles();
    if (upto >= 2) pc.search_rounds();
    if (upto >= 3) pc.assign_layouts();
    auto qubits = [](uint64_t m) {
        std::ostringstream os;
        os << '[';
        bool first = true;
        for (int q = 0; q < 64; ++q)
            if ((m >> q) & 1) {
                os << (first ? "" : ",") << q;
                first = false;
            }
        os << ']';
        return os.str();
    };
    std::ostringstream os;
    os << "{\"nq\":" << nq << ",\"TB\":" << pc.TB << ",\"RB\":" << pc.RB << ",\"RS\":" << pc.RS << ",\"need_shared\":" << std::min(cfg.need_shared, pc.TB)
       << ",\"stages\":" << std::min(std::max(upto, 1), 3) << ",\"passes\":[";
    for (size_t p = 0; p < pc.raw.size(); ++p) {
        const auto& rp = pc.raw[p];
        const auto& g = pc.sg[rp.seg];
        os << (p ? "," : "") << "{\"tile\":" << qubits(rp.tmask) << ",\"seg\":" << rp.seg << ",\"has_phase\":" << (rp.has_phase ? 1 : 0) << ",\"gates\":[";
        for (size_t k = 0; k < rp.gate_idx.size(); ++k)
            os << (k ? "," : "") << '[' << rp.gate_idx[k] << ',' << g[rp.gate_idx[k]].i << ',' << g[rp.gate_idx[k]].j << ',' << g[rp.gate_idx[k]].sweep << ']';
        os << ']';
        if (upto >= 2) {
            os << ",\"rounds\":[";
            for (size_t r = 0; r < pc.rr[p].size(); ++r) {
                const auto& R = pc.rr[p][r];
                os << (r ? "," : "") << "{\"reg\":" << qubits(R.rmask) << ",\"split\":" << (R.split ? 1 : 0) << ",\"left\":" << qubits(R.split ? R.lmask : 0)
                   << ",\"gates\":[";
                for (size_t k = 0; k < R.cidx.size(); ++k) os << (k ? "," : "") << rp.gate_idx[R.cidx[k]];
                os << "]}";
            }
            os << ']';
        }
        if (upto >= 3) {
            os << ",\"layout\":[";
            for (int q = 0; q < nq; ++q) os << (q ? "," : "") << pc.layouts[p][q];
            os << "],\"n_shared_in\":" << pc.nshared[p];
        }
        os << '}';
    }
    os << "],\"segments\":[";
    for (size_t s = 0; s < pc.sg.size(); ++s) os << (s ? "," : "") << pc.sg[s].size();
    os << ']';
    if (upto >= 3 && !pc.layouts.empty()) {
        os << ",\"final_layout\":[";
        for (size_t q = 0; q < pc.layouts.back().size(); ++q) os << (q ? "," : "") << pc.layouts.back()[q];
        os << ']';
    }
    os << '}';
    return os.str();
}

Plan compile_plan(int nq, const std::vector<Segment>& segs, const SchedConfig& cfg) {
    std::lock_guard<std::mutex> memo_lock(g_memo_mutex);
    PlanCompiler pc(nq, segs, cfg);
#ifdef XYK_EXPERIMENTS
    static const bool timing = std::getenv("XYK_SCHED_TIMING") != nullptr;   // development build: stage times on stderr
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t0 = now();
    pc.prepare();
    pc.choose_tiles();
    const auto t1 = now();
    pc.search_rounds();
    const auto t2 = now();
    pc.assign_layouts();
    pc.emit_passes();
    const auto t3 = now();
    if (timing)
        std::fprintf(stderr, "compile_plan nq=%d gates=%d passes=%zu: tiles %.1f ms, rounds %.1f ms, layouts+emit %.1f ms\n", nq, pc.plan.total_gates,
                     pc.plan.passes.size(), ms(t0, t1), ms(t1, t2), ms(t2, t3));
#else
    pc.prepare();
    pc.choose_tiles();
    pc.search_rounds();
    pc.assign_layouts();
    pc.emit_passes();
#endif
    return std::move(pc.plan);
}

std::vector<ShardOp> compile_sharded(int n, const std::vector<int>& global_qubits, const std::vector<Segment>& segs,
                                     const SchedConfig& cfg_in, std::vector<int>* global_final) {
    const int g = (int)global_qubits.size();
    const int nl = n - g;
    if (nl < 1) throw std::runtime_error("compile_sharded: no local qubits");
    std::vector<int> glob(global_qubits);
    // flattened sequence for next-use look-ahead
    struct FG { int i, j; };
    std::vector<FG> flat;
    for (size_t s = 0; s < segs.size(); ++s)
        for (size_t k = 0; k < segs[s].gates.size(); ++k) flat.push_back({segs[s].gates[k].i, segs[s].gates[k].j});
    std::vector<char> fdone(flat.size(), 0);
    size_t fpos = 0;  // first flat index of the current segment

    // ---- phase A: which gates run between which exchanges (no plans yet) ----
    struct Chunk {
        std::vector<int> loc2log, log2loc;
        Segment sub;                      // gates relabelled to local indices
        bool has_work = false;
        bool exchange_after = false;
        std::vector<int> outgoing, incoming;
    };
    std::vector<Chunk> chunks;
    bool restored_at_end = false;
    for (size_t s = 0; s < segs.size(); ++s) {
        const std::vector<GateOp>& gs = segs[s].gates;
        std::vector<char> done(gs.size(), 0);
        size_t ndone = 0;
        bool first = true;
        while (true) {
            uint64_t gmask = 0;
            for (int q : glob) gmask |= 1ull << q;
            Chunk ch;
            for (int q = 0; q < n; ++q)
                if (!((gmask >> q) & 1)) ch.loc2log.push_back(q);
            ch.log2loc.assign(n, -1);
            for (int l = 0; l < nl; ++l) ch.log2loc[ch.loc2log[l]] = l;
            ch.sub.ztau = first ? segs[s].ztau : 0.0;
            std::vector<int> taken;
            {
                uint64_t blocked = 0;
                for (size_t k = 0; k < gs.size(); ++k) {
                    if (done[k]) continue;
                    const uint64_t m = (1ull << gs[k].i) | (1ull << gs[k].j);
                    if (!(m & gmask) && !(m & blocked)) taken.push_back((int)k);
                    else blocked |= m;
                }
            }
            for (int k : taken) {
                GateOp go = gs[k];
                go.i = ch.log2loc[gs[k].i];
                go.j = ch.log2loc[gs[k].j];
                ch.sub.gates.push_back(go);
                done[k] = 1;
                fdone[fpos + k] = 1;
            }
            ndone += taken.size();
            const bool seg_finished = ndone == gs.size();
            ch.has_work = !ch.sub.gates.empty() || (first && ch.sub.ztau != 0.0);
            if (!seg_finished) {
                // the g local qubits whose next use lies furthest ahead leave
                std::vector<size_t> next_use(n, flat.size() + 1);
                for (size_t f = flat.size(); f-- > 0;)
                    if (!fdone[f]) {
                        next_use[flat[f].i] = f;
                        next_use[flat[f].j] = f;
                    }
                std::vector<int> cand(ch.loc2log);
                std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
                    if (next_use[a] != next_use[b]) return next_use[a] > next_use[b];
                    return a > b;
                });
                ch.outgoing.assign(cand.begin(), cand.begin() + g);
                std::sort(ch.outgoing.begin(), ch.outgoing.end());
                ch.incoming = glob;
                ch.exchange_after = true;
            }
            if (ch.has_work) first = false;
            // restore: after the very last gates the qubits that were sharded on entry go back to the rank bits, so
            // that the next call finds the same sharding and reuses this program (one more exchange per call)
            bool restored = false;
            if (seg_finished && cfg_in.restore_globals && s + 1 == segs.size() && glob != global_qubits) {
                bool feasible = true;
                for (int q : global_qubits)
                    for (int x : glob) feasible = feasible && (x != q);
                if (feasible) {
                    ch.outgoing = global_qubits;
                    ch.incoming = glob;
                    ch.exchange_after = true;
                    restored = true;
                    restored_at_end = true;
                }
            }
            if (ch.has_work || ch.exchange_after) chunks.push_back(ch);
            if (restored) glob = global_qubits;
            if (seg_finished) break;
            glob = ch.outgoing;  // rank bit k now carries outgoing[k]
        }
        fpos += gs.size();
    }
    if (global_final) *global_final = glob;

    // ---- phase B: compile the local programs back to front, so that the program before an exchange
    // can write (with peer stores) straight into the start layout of the program after it ----
    const int C = (int)chunks.size();
    std::vector<ShardOp> local_ops(C);
    std::vector<char> fused(C, 0);
    // how each program was chained to its successor (needed when it is recompiled in step (2))
    std::vector<uint64_t> nop_next_hint(C, 0);
    std::vector<std::vector<int>> nop_final_positions(C), nop_final_top(C);
    std::vector<uint64_t> nop_prev_hint(C, 0), nop_prev_rlast(C, 0);  // ... and to its predecessor
    // translate a mask over chunk a's local indices into chunk b's local indices (qubits local in both)
    auto translate = [&](uint64_t mask, const Chunk& a, const Chunk& b) -> uint64_t {
        uint64_t out = 0;
        for (int l = 0; l < nl; ++l)
            if ((mask >> l) & 1) {
                const int lb = b.log2loc[a.loc2log[l]];
                if (lb >= 0) out |= 1ull << lb;
            }
        return out;
    };
    auto base_cfg = [&]() {
        SchedConfig cfg = cfg_in;
        cfg.cyclic = false;
        cfg.final_top.clear();
        cfg.final_positions.clear();
        cfg.next_tile_hint = cfg.prev_tile_hint = cfg.prev_rlast_hint = 0;
        return cfg;
    };
    for (int c = C - 1; c >= 0; --c) {
        Chunk& ch = chunks[c];
        ShardOp& op = local_ops[c];
        op.loc2log = ch.loc2log;
        if (!ch.has_work) continue;
        const bool can_fuse = cfg_in.fuse_io && cfg_in.peer_exchange && ch.exchange_after && c + 1 < C &&
                              chunks[c + 1].has_work && !local_ops[c + 1].plan.passes.empty() && !ch.sub.gates.empty();
        std::vector<Segment> one{ch.sub};
        if (!can_fuse) {
            SchedConfig cfg = base_cfg();
            if (ch.exchange_after)
                for (int q : ch.outgoing) cfg.final_top.push_back(ch.log2loc[q]);
            nop_final_top[c] = cfg.final_top;
            op.plan = compile_plan(nl, one, cfg);
            continue;
        }
        Chunk& nx = chunks[c + 1];
        ShardOp& nop = local_ops[c + 1];
        // (1) this program with the sharing hint towards the next program's first tile (learn its last tile)
        SchedConfig cfg = base_cfg();
        cfg.next_tile_hint = translate(nop.plan.first_tile_mask, nx, ch);
        for (int q : ch.outgoing) cfg.final_top.push_back(ch.log2loc[q]);
        Plan probe = compile_plan(nl, one, cfg);
        if (probe.passes.empty()) {
            op.plan = std::move(probe);
            continue;
        }
        // (2) the next program again, now ordering its start layout around what this program leaves low
        {
            SchedConfig ncfg = base_cfg();
            ncfg.prev_tile_hint = translate(probe.last_tile_mask, ch, nx);
            ncfg.prev_rlast_hint = translate(probe.last_rlast_mask, ch, nx);
            nop_prev_hint[c + 1] = ncfg.prev_tile_hint;
            nop_prev_rlast[c + 1] = ncfg.prev_rlast_hint;
            // keep whatever chained the next program to ITS successor
            ncfg.next_tile_hint = nop_next_hint[c + 1];
            ncfg.final_positions = nop_final_positions[c + 1];
            ncfg.final_top = nop_final_top[c + 1];
            std::vector<Segment> nseg{nx.sub};
            nop.plan = compile_plan(nl, nseg, ncfg);
            if (nop.peer_last) {  // its own fused exchange data is unchanged (depends on the plan after it)
            }
        }
        // (3) this program for real: its last pass writes the next program's start layout
        cfg.final_top.clear();
        cfg.final_positions.assign(nl, -1);
        for (int l = 0; l < nl; ++l) {
            const int q = ch.loc2log[l];
            int k = -1;
            for (int x = 0; x < g; ++x)
                if (ch.outgoing[x] == q) k = x;
            if (k >= 0) cfg.final_positions[l] = nl + k;                        // leaves to rank bit k
            else cfg.final_positions[l] = nop.plan.start_perm[nx.log2loc[q]];   // stays: next program's start layout
        }
        nop_next_hint[c] = cfg.next_tile_hint;
        nop_final_positions[c] = cfg.final_positions;
        op.plan = compile_plan(nl, one, cfg);
        if (op.plan.passes.empty()) continue;
        fused[c] = 1;
        op.peer_last = true;
        op.outgoing_fused = ch.outgoing;
        op.incoming_fused = ch.incoming;
        op.rank_out.clear();
        for (int k = 0; k < g; ++k) op.rank_out.push_back(nop.plan.start_perm[nx.log2loc[ch.incoming[k]]]);
    }
    // ---- the restoring exchange wraps around: fuse it into the last program's last pass, which then stores (with
    // peer stores) straight into the START layout of the first program -- the next call begins without an exchange,
    // without a staging sweep and without a re-layout ----
    if (restored_at_end && cfg_in.fuse_io && cfg_in.peer_exchange && C >= 2) {
        const int L = C - 1;
        int F = -1;
        for (int c = 0; c < C && F < 0; ++c)
            if (chunks[c].has_work && !local_ops[c].plan.passes.empty()) F = c;
        if (F >= 0 && F != L && chunks[L].has_work && chunks[L].exchange_after && !fused[L] && !chunks[L].sub.gates.empty() &&
            !local_ops[L].plan.passes.empty()) {
            Chunk& chF = chunks[F];
            Chunk& chL = chunks[L];
            // (a) the first program again, its start layout ordered around what the last program leaves low
            SchedConfig fcfg = base_cfg();
            fcfg.prev_tile_hint = translate(local_ops[L].plan.last_tile_mask, chL, chF);
            fcfg.prev_rlast_hint = translate(local_ops[L].plan.last_rlast_mask, chL, chF);
            fcfg.next_tile_hint = nop_next_hint[F];
            fcfg.final_positions = nop_final_positions[F];
            fcfg.final_top = nop_final_top[F];
            std::vector<Segment> fseg{chF.sub};
            Plan planF = compile_plan(nl, fseg, fcfg);
            // (b) the last program again: its last pass writes that start layout, the restored qubits leave to the rank bits
            SchedConfig lcfg = base_cfg();
            lcfg.prev_tile_hint = nop_prev_hint[L];
            lcfg.prev_rlast_hint = nop_prev_rlast[L];
            lcfg.next_tile_hint = translate(planF.first_tile_mask, chF, chL);
            lcfg.final_positions.assign(nl, -1);
            bool ok = true;
            for (int l = 0; l < nl; ++l) {
                const int q = chL.loc2log[l];
                int k = -1;
                for (int x = 0; x < g; ++x)
                    if (chL.outgoing[x] == q) k = x;
                if (k >= 0) lcfg.final_positions[l] = nl + k;
                else if (chF.log2loc[q] >= 0) lcfg.final_positions[l] = planF.start_perm[chF.log2loc[q]];
                else ok = false;
            }
            for (int k = 0; k < g && ok; ++k) ok = chF.log2loc[chL.incoming[k]] >= 0;
            if (ok) {
                std::vector<Segment> lseg{chL.sub};
                Plan planL = compile_plan(nl, lseg, lcfg);
                // the program before the last one stores into the last program's start layout: it must not have moved
                if (!planL.passes.empty() && planL.start_perm == local_ops[L].plan.start_perm) {
                    local_ops[F].plan = std::move(planF);
                    ShardOp& op = local_ops[L];
                    op.plan = std::move(planL);
                    op.peer_last = true;
                    op.outgoing_fused = chL.outgoing;
                    op.incoming_fused = chL.incoming;
                    op.rank_out.clear();
                    for (int k = 0; k < g; ++k) op.rank_out.push_back(local_ops[F].plan.start_perm[chF.log2loc[chL.incoming[k]]]);
                    fused[L] = 1;
                }
            }
        }
    }
    std::vector<ShardOp> ops;
    for (int c = 0; c < C; ++c) {
        if (chunks[c].has_work) ops.push_back(std::move(local_ops[c]));
        if (chunks[c].exchange_after && !fused[c]) {
            ShardOp ex;
            ex.is_exchange = true;
            ex.outgoing = chunks[c].outgoing;
            ex.incoming = chunks[c].incoming;
            ops.push_back(ex);
        }
    }
    return ops;
}

std::string sharded_to_json(const std::vector<ShardOp>& ops) {
    std::ostringstream os;
    os << "{\"ops\":[";
    for (size_t k = 0; k < ops.size(); ++k) {
        const ShardOp& op = ops[k];
        os << (k ? "," : "");
        if (op.is_exchange) {
            os << "{\"kind\":\"exchange\",\"outgoing\":[";
            for (size_t i = 0; i < op.outgoing.size(); ++i) os << (i ? "," : "") << op.outgoing[i];
            os << "],\"incoming\":[";
            for (size_t i = 0; i < op.incoming.size(); ++i) os << (i ? "," : "") << op.incoming[i];
            os << "]}";
        } else {
            os << "{\"kind\":\"local\",\"peer_last\":" << (op.peer_last ? 1 : 0) << ",\"rank_out\":[";
            for (size_t i = 0; i < op.rank_out.size(); ++i) os << (i ? "," : "") << op.rank_out[i];
            os << "],\"outgoing\":[";
            for (size_t i = 0; i < op.outgoing_fused.size(); ++i) os << (i ? "," : "") << op.outgoing_fused[i];
            os << "],\"incoming\":[";
            for (size_t i = 0; i < op.incoming_fused.size(); ++i) os << (i ? "," : "") << op.incoming_fused[i];
            os << "],\"loc2log\":[";
            for (size_t i = 0; i < op.loc2log.size(); ++i) os << (i ? "," : "") << op.loc2log[i];
            os << "],\"end_perm\":[";
            for (size_t i = 0; i < op.plan.end_perm.size(); ++i) os << (i ? "," : "") << op.plan.end_perm[i];
            os << "],\"plan\":" << plan_to_json(op.plan) << "}";
        }
    }
    os << "]}";
    return os.str();
}

std::string plan_to_json(const Plan& plan) {
    std::ostringstream os;
    os.precision(17);
    os << "{\"nq\":" << plan.nq << ",\"TB\":" << plan.TB << ",\"RB\":" << plan.RB
       << ",\"total_gates\":" << plan.total_gates << ",\"total_rounds\":" << plan.total_rounds
       << ",\"layers1_equiv\":" << plan.layers1_equiv << ",\"trailing_ztau\":" << plan.trailing_ztau
       << ",\"start_perm\":[";
    for (size_t i = 0; i < plan.start_perm.size(); ++i) os << (i ? "," : "") << plan.start_perm[i];
    os << "],\"passes\":[";
    for (size_t p = 0; p < plan.passes.size(); ++p) {
        const PlanPass& pp = plan.passes[p];
        os << (p ? "," : "") << "{\"tileq\":[";
        for (int k = 0; k < plan.TB; ++k) os << (k ? "," : "") << pp.tileq[k];
        os << "],\"out_of_in\":[";
        for (int k = 0; k < plan.nq; ++k) os << (k ? "," : "") << pp.out_of_in[k];
        os << "],\"n_shared\":" << pp.n_shared << ",\"has_phase\":" << (pp.has_phase ? 1 : 0)
           << ",\"ztau\":" << pp.ztau << ",\"ngates\":" << pp.ngates << ",\"fuse_load\":" << (pp.fuse_load ? 1 : 0)
           << ",\"fuse_store\":" << (pp.fuse_store ? 1 : 0) << ",\"rounds\":[";
        for (size_t r = 0; r < pp.rounds.size(); ++r) {
            const PlanRound& pr = pp.rounds[r];
            const int nreg = pr.nreg ? pr.nreg : plan.RB;
            os << (r ? "," : "") << "{\"split\":" << pr.split << ",\"nleft\":" << pr.nleft << ",\"regq\":[";
            for (int k = 0; k < nreg; ++k) os << (k ? "," : "") << pr.regq[k];
            os << "],\"regpos\":[";
            for (int k = 0; k < nreg; ++k) os << (k ? "," : "") << pr.regpos[k];
            os << "],\"thrpos\":[";
            for (int k = 0; k < plan.TB - nreg; ++k) os << (k ? "," : "") << pr.thrpos[k];
            os << "],\"mask\":" << pr.mask << ",\"cta_sync_after\":" << pr.cta_sync_after << ",\"gates\":[";
            bool firstg = true;
            if (pr.split) {
                const int HL = pr.nleft, HR = nreg - pr.nleft;
                for (int a = 0; a < HL; ++a)
                    for (int b = 0; b < HR; ++b)
                        if ((pr.mask >> (HR * a + b)) & 1) {
                            os << (firstg ? "" : ",") << "[" << pr.regq[a] << "," << pr.regq[HL + b] << "," << pr.phi[HR * a + b] << "]";
                            firstg = false;
                        }
            } else {
                int slot = 0;
                for (int a = 0; a < plan.RB; ++a)
                    for (int b = a + 1; b < plan.RB; ++b, ++slot)
                        if ((pr.mask >> slot) & 1) {
                            os << (firstg ? "" : ",") << "[" << pr.regq[a] << "," << pr.regq[b] << ","
                               << pr.phi[slot] << "]";
                            firstg = false;
                        }
            }
            os << "]}";
        }
        os << "]}";
    }
    os << "]}";
    return os.str();
}

}  // namespace xyk
