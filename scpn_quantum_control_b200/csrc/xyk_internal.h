// Internal structures shared by the schedule compiler (host C++) and the CUDA engine.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace xyk {

constexpr int kMaxQ = 40;          // maximum number of local qubits handled by one handle
constexpr double kSparsityEps = 1e-15;  // reference: _constants.py COUPLING_SPARSITY_EPS

struct PairTerm {
    int i, j;
    double k;  // symmetrised coupling K_ij
};
struct ZTerm {
    int i;
    double w;
};

// One pair rotation of the flattened product-formula sequence.
struct GateOp {
    int i, j;    // logical qubits, i < j
    double phi;  // rotation angle 2 * K_ij * tau_effective
    int sweep;   // index of the monotone sweep this gate belongs to (a pair occurs once per sweep)
};

// A diagonal phase of "time" ztau followed by a run of pair gates.
struct Segment {
    double ztau;
    std::vector<GateOp> gates;
};

// Term list in the reference's order (bridge/knm_hamiltonian.py:172-195).
void build_terms(int n, const double* K, const double* omega, std::vector<PairTerm>& pairs,
                 std::vector<ZTerm>& zterms, std::vector<double>& Ksym);

// Flatten `reps` repetitions of the order-`order` formula over one interval `delta`.
// order 1: LieTrotter; 2: SuzukiTrotter(order=2); 4, 6: Suzuki recursion over the order-2 layer.
std::vector<Segment> build_sequence(const std::vector<PairTerm>& pairs, double delta, int order,
                                    int reps, int* layers1_equiv);

// Two kinds of register rounds:
//  * complex round: each thread holds 2^RB complex amplitudes (RB register qubits) and may apply any of
//    the RB(RB-1)/2 pairs among them;
//  * split round: a pair rotation only mixes Re(a) with Im(b) and Im(a) with Re(b).  When every gate of a
//    round joins a qubit of a "left" group with a qubit of a "right" group, the parity f of the right-group
//    bits labels two independent real half-states {Re a : f(a)=h} u {Im a : f(a)!=h}, h = 0, 1.  A thread
//    then holds ONE real component of 2^(RB+1) amplitudes in the same registers: RB+1 register qubits,
//    (RB+1)/2 on each side, 9 gates per round for RB = 5 where a complex round reaches only 6 on such a
//    bipartite gate set.  Thread bit 0 selects the half h.
struct PlanRound {
    int regq[8];      // logical qubit of register bit k (ascending for forward sweeps, descending for reverse);
                      // split rounds: left group on bits [0, nreg/2), right group on bits [nreg/2, nreg)
    int regpos[8];    // tile-local position of register bit k
    int thrpos[16];   // tile-local position of thread-qubit bit b (TB - nreg entries; split rounds: thread bit b + 1)
    uint32_t mask;    // enabled pair slots; complex: slot order = lexicographic (a, b), a < b over register bits;
                      // split: slot = (nreg/2) * a + b, a = index in the left group, b = index in the right group
    double phi[16];   // angle per slot (only enabled slots are meaningful)
    int ngates;
    int cta_sync_after;  // 1: block-wide barrier after this round's exchange, 0: the data stays inside each warp
    int split;        // 1: split round
    int nreg;         // register qubits of this round (RB, or RB + 1 for split rounds)
    int nleft;        // split rounds: size of the left group (nreg / 2, or down to SchedConfig::split_min_left)
};

struct PlanPass {
    int tileq[16];                 // logical qubit at tile-local position p (= physical position p on input)
    std::vector<PlanRound> rounds;
    int out_of_in[kMaxQ];          // physical input bit b -> physical output bit
    int n_shared;                  // tile qubits that land on output positions [0, n_shared) (contiguous chunk bits)
    bool has_phase;                // diagonal phase exp(+i ztau sum w_q (1-2b_q)) applied before the gates
    double ztau;
    bool std_rot;                  // some angle too large for the shear (lifting) form
    bool fuse_load;                // round 0 can read global memory directly (coalesced lane mapping exists)
    bool fuse_store;               // the last round can write global memory directly
    int ngates;
};

struct Plan {
    int nq = 0;  // local qubits
    int TB = 0, RB = 0;
    std::vector<int> start_perm;   // logical -> physical layout the program expects on entry
    std::vector<int> end_perm;     // layout on exit (== start_perm: programs are cyclic)
    std::vector<PlanPass> passes;
    double trailing_ztau = 0.0;    // diagonal phase after the last pass (0 if none)
    bool has_trailing_phase = false;
    int total_gates = 0, total_rounds = 0, layers1_equiv = 0;
    uint64_t first_tile_mask = 0, last_tile_mask = 0, last_rlast_mask = 0;  // over plan indices
};

struct SchedConfig {
    int TB = 12;           // tile bits (amplitudes per tile = 2^TB)
    int RB = 5;            // register bits (amplitudes per thread = 2^RB)
    int need_shared = 3;   // consecutive tiles must share this many qubits (contiguous output chunks)
    int max_rounds = 14;   // rounds a single pass descriptor can carry
    int max_pass_gates = 1 << 20;  // gates a single pass may carry (no limit in the current kernel)
    int lane_bits = 3;     // lowest physical bits that must map to lanes for fused global I/O (2^3 x 16 B = 128 B)
    bool fuse_io = true;   // plan first / last rounds so that they can touch global memory directly
    bool split_rounds = true;  // allow split rounds (RB + 1 register qubits over two real half-states)
    int tile_groups = 100;       // > 0: cost-guided beam (this width) over tiles made of index triples (cyclic programs)
    bool tile_groups_open = true;  // ... also for the open (non-cyclic) local programs of a sharded state
    int tile_beam = 1;         // tile search: sequences kept per depth (1 with tile_seeds = 1: greedy)
    int tile_seeds = 1;        // tile search: seed gates tried per state
    int beam_width = 6;        // round search: states kept per depth (1 = greedy)
    int beam_cands = 6;        // round search: candidate rounds expanded per state
    bool unb_first = true, unb_last = true;  // unbalanced split rounds allowed as the first / last round of a pass
    int split_min_left = 1;    // smallest left group of a split round: 3 = balanced 3 | 3 only, 2 adds 2 | 4, 1 adds 1 | 5
                               // (a pass uses one unbalanced kind only: each has its own kernel instantiation)
    int split_margin = 1;      // a split round must run at least this many more gates than the best complex round
                               // (it moves the same bytes with twice as many shared-memory instructions)
    bool cyclic = true;    // the program ends in its own start layout (it can be replayed back to back)
    bool peer_exchange = false;  // sharded: fuse exchanges into the preceding pass (needs mapped peer buffers)
    bool restore_globals = false;  // sharded: end every call with the sharding it started from (program reuse)
    std::vector<int> final_top;  // non-cyclic: qubits that must end on the highest physical bits, in this order
    // open programs chained across an exchange that is fused into the last pass's store:
    uint64_t next_tile_hint = 0;        // qubits (local indices) of the NEXT program's first tile that are local here
    uint64_t prev_tile_hint = 0;        // qubits (local indices) of the PREVIOUS program's last tile that are local here
    uint64_t prev_rlast_hint = 0;       // ... and those its last round holds in registers
    std::vector<int> final_positions;   // explicit output bit of every local index after the last pass
                                        // (>= nq: the qubit leaves to rank bit (position - nq))
};

// Compile the segments into tile passes over nq local qubits.  Throws std::runtime_error.
Plan compile_plan(int nq, const std::vector<Segment>& segs, const SchedConfig& cfg);

// Bank class of tile-local position p in the padded shared-memory tile layout (xyk_kernels.cuh,
// pad_slot): the 16-byte-column stride of position p is 1, 2 or 4 (mod 8) for p mod 3 = 0, 1, 2.
// Three lane bits on positions of three different classes are bank-conflict free.
inline int lane_class(int p) { return p % 3; }

std::string plan_to_json(const Plan& plan);

// Unit-test view of the schedule compiler: the intermediate results of its first `upto` stages (1: tile search, 2: + round
// search, 3: + layouts) as JSON.
std::string plan_stages_to_json(int nq, const std::vector<Segment>& segs, const SchedConfig& cfg, int upto);

// ---- sharded state (top g physical bits = rank): local programs separated by exchanges ----
struct ShardOp {
    bool is_exchange = false;
    // the exchange that follows this local program is fused into its last pass (peer stores):
    bool peer_last = false;
    std::vector<int> rank_out;  // peer_last: output bit (local) of the qubit currently on rank bit k
    std::vector<int> outgoing_fused, incoming_fused;
    // local program
    std::vector<int> loc2log;   // logical qubit of local index l (ascending)
    Plan plan;                  // over local indices
    // exchange: the g qubits on the top local physical bits swap with the g global qubits
    std::vector<int> outgoing;  // logical qubits leaving (to be placed on the top local bits, in this order)
    std::vector<int> incoming;  // logical qubits arriving (currently on rank bit k, in this order)
};

// Greedy schedule for a state whose `global_qubits` (logical ids, rank bit k -> global_qubits[k])
// are sharded across 2^g ranks: run every gate whose qubits are both local (dependency closure),
// then exchange the g global qubits with the g local qubits whose next use lies furthest ahead.
std::vector<ShardOp> compile_sharded(int n, const std::vector<int>& global_qubits, const std::vector<Segment>& segs,
                                     const SchedConfig& cfg, std::vector<int>* global_final);

std::string sharded_to_json(const std::vector<ShardOp>& ops);

}  // namespace xyk
