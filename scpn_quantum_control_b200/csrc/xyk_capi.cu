// C ABI (include/xyk.h) of the B200-native Kuramoto-XY statevector engine.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/xyk.h"
#include "xyk_internal.h"
#include "xyk_kernels.cuh"
#include "xyk_small.cuh"

using namespace xyk;

namespace {

constexpr int kTB = 12;  // complex128 tile: 2^12 amplitudes = 64 KiB of shared memory
constexpr int kRB = 5;   // 32 amplitudes (128 registers) per thread
using Desc = PassDesc<kTB, kRB>;
static_assert(sizeof(Desc) <= 4000, "pass descriptor must fit the classic 4 KB kernel-parameter space");

thread_local std::string g_global_error;

struct Program {
    Plan plan;
    std::vector<int> loc2log;       // logical qubit of plan index l (identity on a single GPU)
    std::vector<Desc> descs;        // one per pass
    std::vector<char> lift;         // shear form allowed for the pass
    std::vector<void*> dev_tabs;    // device tables owned by the program (fused-phase factors)
    Program() = default;
    Program(const Program&) = delete;
    Program& operator=(const Program&) = delete;
    ~Program() {
        for (void* p : dev_tabs) cudaFree(p);
    }
};

struct ShardStep {
    bool is_exchange = false;
    std::shared_ptr<Program> prog;
    std::vector<int> outgoing, incoming;
    bool peer_last = false;            // the last pass carries the exchange (peer stores)
    std::vector<int> rank_out;         // local output bit of the qubit arriving from rank bit k
};

}  // namespace

struct xyk_handle_s {
    int n = 0, nlocal = 0, world = 1, rank = 0, gbits = 0, dtype = 0, device = 0;
    cudaStream_t stream = nullptr;
    void* buf[2] = {nullptr, nullptr};
    size_t bytes = 0;
    int cur = 0;
    std::vector<int> perm;  // logical qubit -> physical bit (>= nlocal: sharded over ranks)
    bool have_problem = false, state_valid = false;
    std::vector<double> Ksym, omega;
    std::vector<PairTerm> pairs;
    std::vector<ZTerm> zterms;
    uint64_t problem_version = 0;
    int engine_opt = XYK_ENGINE_AUTO;
    int fuse_phase = 1;
    int tiled_obs = 1;
    int debug_mode = 0;  // timing experiments only (results are wrong): 1 = no gates, 2 = one empty round per pass
    int sm_count = 148;
    int pass_grid = 0;
    // scratch
    double2* d_tab = nullptr;     // 3 x 2^11 phase tables
    double zz_delta = 0.0;        // XXZ anisotropy (0: XY model)
    double* d_J = nullptr;        // n x n couplings K_ij * delta of the ZZ terms (logical indices)
    double2* d_zztab = nullptr;   // 2^12 low-bit factors of the fused ZZ diagonal
    double* d_partial = nullptr;  // reduction partials
    double* d_out = nullptr;      // reduced values
    double* h_out = nullptr;      // pinned staging
    std::map<std::tuple<uint64_t, int, int, double>, std::shared_ptr<Program>> programs;
    xyk_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> pass_events;  // per-launch timing of the tile pass kernel
    // sharded execution state machine
    std::vector<ShardStep> shard_steps;
    struct ShardCacheEntry {
        std::vector<ShardStep> steps;
        int gates = 0, rounds = 0, passes = 0;
    };
    std::map<std::string, ShardCacheEntry> shard_cache;  // compiled sharded programs by (problem, formula, sharding on entry)
    int64_t shard_compiles = 0;
    int shard_gates = 0, shard_rounds = 0, shard_passes = 0;  // totals of the sharded call in flight (stats at its end)
    size_t shard_pos = 0;
    bool exchange_pending = false;
    std::vector<int> ex_out, ex_in;
    int64_t exchanges = 0;
    int shard_phase = 0;               // progress inside a step whose last pass is a peer pass
    void* peer_buf[8][2] = {};         // state buffers of every rank (own pointers for rank == this rank)
    bool peer_open[8][2] = {};
    bool peers_ready = false;
    int peer_exchange_opt = 1;
    std::string err;
};

namespace {

constexpr int kTabStride = 2048;
constexpr int kRedBlocks = 1184;  // 8 x 148
constexpr int kPartialDoubles = kRedBlocks * (kMaxQ + 1) * 2;

int fail(xyk_handle h, int code, const std::string& msg) {
    if (h) h->err = msg;
    else g_global_error = msg;
    return code;
}

#define XYK_CUDA(h, expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            const bool oom__ = (e__ == cudaErrorMemoryAllocation);                             \
            return fail(h, oom__ ? XYK_ERR_NOMEM : XYK_ERR_CUDA,                               \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
        }                                                                                      \
    } while (0)

#define XYK_CHECK(h, cond, code, msg) \
    do {                              \
        if (!(cond)) return fail(h, code, msg); \
    } while (0)

size_t amp_bytes(int dtype) { return dtype == 0 ? 16 : 8; }

int grid_for(xyk_handle h, uint64_t items, int block) {
    const uint64_t want = (items + block - 1) / block;
    const uint64_t cap = (uint64_t)h->sm_count * 16;
    return (int)std::max<uint64_t>(1, std::min(want, cap));
}

int ensure_alt(xyk_handle h) {
    if (h->buf[1]) return XYK_OK;
    XYK_CUDA(h, cudaMalloc(&h->buf[1], h->bytes));
    return XYK_OK;
}

bool perm_is_identity(xyk_handle h) {
    for (int q = 0; q < h->n; ++q)
        if (h->perm[q] != q) return false;
    return true;
}

// |psi0> = prod_q Ry(theta_q)|0> in the current layout (theta_q = omega_q mod 2 pi when `angles` is null): three chunk
// tables built on the host
template <typename real>
int launch_init(xyk_handle h, const double* angles = nullptr) {
    const double two_pi = 2.0 * M_PI;
    std::vector<double> cq(h->n), sq(h->n);
    for (int q = 0; q < h->n; ++q) {
        double a = angles ? angles[q] : std::fmod(h->omega[q], two_pi);  // Python's % : result has the sign of the modulus
        if (!angles && a < 0.0) a += two_pi;
        cq[q] = std::cos(a / 2.0);
        sq[q] = std::sin(a / 2.0);
        if (angles && a == 0.0) sq[q] = 0.0;
    }
    std::vector<int> log_of_bit(h->n, -1);  // logical qubit on physical bit b
    for (int q = 0; q < h->n; ++q) log_of_bit[h->perm[q]] = q;
    const int nl = h->nlocal;
    using C = typename cplx<real>::type;
    if (nl < 4) {  // a handful of amplitudes: written from the host
        std::vector<C> psi0((size_t)1 << nl);
        for (size_t p = 0; p < psi0.size(); ++p) {
            double x = 1.0;
            for (int b = 0; b < h->n; ++b) {
                const int bit = b < nl ? (int)((p >> b) & 1) : ((h->rank >> (b - nl)) & 1);
                x *= bit ? sq[log_of_bit[b]] : cq[log_of_bit[b]];
            }
            psi0[p].x = (real)x;
            psi0[p].y = (real)0;
        }
        XYK_CUDA(h, cudaMemcpyAsync(h->buf[h->cur], psi0.data(), psi0.size() * sizeof(C), cudaMemcpyHostToDevice, h->stream));
        XYK_CUDA(h, cudaStreamSynchronize(h->stream));
        return XYK_OK;
    }
    int b1 = std::min(nl, 11), b2 = std::min(nl, 22);
    if (nl - b2 > 11) {
        b1 = (nl + 2) / 3;
        b2 = b1 + (nl - b1 + 1) / 2;
    }
    XYK_CHECK(h, b1 <= 11 && b2 - b1 <= 11 && nl - b2 <= 11, XYK_ERR_UNSUPPORTED, "product-state tables support at most 33 local qubits");
    double rank_factor = 1.0;
    for (int b = nl; b < h->n; ++b) rank_factor *= ((h->rank >> (b - nl)) & 1) ? sq[log_of_bit[b]] : cq[log_of_bit[b]];
    double* htab = reinterpret_cast<double*>(h->h_out);   // pinned staging: 3 x kTabStride doubles
    const int lo[3] = {0, b1, b2}, hi[3] = {b1, b2, nl};
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < (1 << (hi[c] - lo[c])); ++v) {
            double x = c == 2 ? rank_factor : 1.0;
            for (int b = lo[c]; b < hi[c]; ++b) x *= ((v >> (b - lo[c])) & 1) ? sq[log_of_bit[b]] : cq[log_of_bit[b]];
            htab[c * kTabStride + v] = x;
        }
    XYK_CUDA(h, cudaMemcpyAsync(h->d_tab, htab, sizeof(double) * 3 * kTabStride, cudaMemcpyHostToDevice, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));  // the staging buffer is reused by other calls
    init_product_kernel<real><<<grid_for(h, 1ull << nl, 256), 256, 0, h->stream>>>((C*)h->buf[h->cur], reinterpret_cast<const double*>(h->d_tab),
                                                                              kTabStride, nl, b1, b2);
    h->stats.kernel_launches++;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// diagonal phase exp(+i tau sum_q omega_q (1-2 b_q)) in the current layout (gate-by-gate path)
int apply_phase(xyk_handle h, double tau) {
    if (h->zterms.empty() || tau == 0.0) return XYK_OK;
    PhaseSpec S{};
    S.nbits = h->n;
    S.nlocal = h->nlocal;
    S.rank = h->rank;
    S.tau = tau;
    for (int b = 0; b < kMaxQ; ++b) S.w[b] = 0.0;
    for (const auto& z : h->zterms) S.w[h->perm[z.i]] = z.w;
    const int nl = h->nlocal;
    S.b1 = std::min(nl, 11);
    S.b2 = std::min(nl, 22);
    if (nl - S.b2 > 11) {  // spread three chunks evenly for very large shards
        S.b1 = (nl + 2) / 3;
        S.b2 = S.b1 + (nl - S.b1 + 1) / 2;
    }
    XYK_CHECK(h, S.b1 <= 11 && S.b2 - S.b1 <= 11 && nl - S.b2 <= 11, XYK_ERR_UNSUPPORTED,
              "phase tables support at most 33 local qubits");
    phase_table_kernel<<<dim3(8, 3), 256, 0, h->stream>>>(h->d_tab, S, kTabStride);
    const uint64_t dim = 1ull << nl;
    if (h->dtype == 0)
        phase_apply_kernel<double><<<grid_for(h, dim, 256), 256, 0, h->stream>>>(
            (double2*)h->buf[h->cur], h->d_tab, kTabStride, nl, S.b1, S.b2);
    else
        phase_apply_kernel<float><<<grid_for(h, dim, 256), 256, 0, h->stream>>>(
            (float2*)h->buf[h->cur], h->d_tab, kTabStride, nl, S.b1, S.b2);
    h->stats.kernel_launches += 2;
    h->stats.state_sweeps += 1;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// XXZ: exp(+i t sum_{i<j} K_ij delta Z_i Z_j) on the whole state in the current layout (one sweep)
int apply_zz_diag(xyk_handle h, double t) {
    if (h->zz_delta == 0.0 || t == 0.0 || h->pairs.empty()) return XYK_OK;
    ZZSpec S{};
    S.nbits = h->n;
    S.nlocal = h->nlocal;
    S.rank = h->rank;
    S.LB = std::min(h->nlocal, 12);
    S.t = t;
    for (int q = 0; q < h->n; ++q) S.log_of_bit[h->perm[q]] = q;
    zz_low_table_kernel<<<16, 256, 0, h->stream>>>(h->d_zztab, h->d_J, h->n, S);
    const uint64_t nblk = 1ull << (h->nlocal - S.LB);
    const int grid = (int)std::min<uint64_t>(nblk, (uint64_t)h->sm_count * 8);
    if (h->dtype == 0)
        zz_diag_kernel<double><<<grid, 256, 0, h->stream>>>((double2*)h->buf[h->cur], h->d_zztab, h->d_J, h->n, S);
    else
        zz_diag_kernel<float><<<grid, 256, 0, h->stream>>>((float2*)h->buf[h->cur], h->d_zztab, h->d_J, h->n, S);
    h->stats.kernel_launches += 2;
    h->stats.state_sweeps += 1;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// XXZ, gate-by-gate engine: the ZZ term of one pair right after its XX, YY terms (the reference's term order)
int apply_zz_simple(xyk_handle h, int i, int j, double theta) {
    const double c = std::cos(theta), s = std::sin(theta);
    if (h->dtype == 0)
        zz_gate_kernel<double><<<grid_for(h, 1ull << h->nlocal, 256), 256, 0, h->stream>>>((double2*)h->buf[h->cur], h->nlocal, h->perm[i], h->perm[j], c, s);
    else
        zz_gate_kernel<float><<<grid_for(h, 1ull << h->nlocal, 256), 256, 0, h->stream>>>((float2*)h->buf[h->cur], h->nlocal, h->perm[i], h->perm[j], c, s);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    return XYK_OK;
}

int apply_pair_simple(xyk_handle h, int i, int j, double phi) {
    const int pi = h->perm[i], pj = h->perm[j];
    XYK_CHECK(h, pi < h->nlocal && pj < h->nlocal, XYK_ERR_UNSUPPORTED,
              "gate-by-gate engine cannot touch a sharded (global) qubit");
    const double c = std::cos(phi), s = std::sin(phi);
    const uint64_t quarter = 1ull << (h->nlocal - 2);
    if (h->dtype == 0)
        pair_gate_kernel<double><<<grid_for(h, quarter, 256), 256, 0, h->stream>>>((double2*)h->buf[h->cur],
                                                                                 h->nlocal, pi, pj, c, s);
    else
        pair_gate_kernel<float><<<grid_for(h, quarter, 256), 256, 0, h->stream>>>((float2*)h->buf[h->cur],
                                                                                h->nlocal, pi, pj, c, s);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    return XYK_OK;
}

int run_segments_simple(xyk_handle h, const std::vector<Segment>& segs) {
    for (const auto& s : segs) {
        int rc = apply_phase(h, s.ztau);
        if (rc) return rc;
        for (const auto& g : s.gates) {
            rc = apply_pair_simple(h, g.i, g.j, g.phi);
            if (rc) return rc;
            if (h->zz_delta != 0.0) {  // g.phi = 2 K_ij tau: the ZZ term -K_ij delta Z_i Z_j evolves by exp(+i K_ij delta tau Z_i Z_j)
                rc = apply_zz_simple(h, g.i, g.j, 0.5 * g.phi * h->zz_delta);
                if (rc) return rc;
            }
        }
    }
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// permute the state so that logical qubit q sits on physical bit target[q]
int relayout(xyk_handle h, const std::vector<int>& target) {
    bool same = true;
    for (int q = 0; q < h->n; ++q) same &= (h->perm[q] == target[q]);
    if (same) return XYK_OK;
    PermSpec S{};
    S.nlocal = h->nlocal;
    for (int b = 0; b < kMaxQ; ++b) S.dst_of_src[b] = (uint8_t)b;
    for (int q = 0; q < h->n; ++q) {
        XYK_CHECK(h, (h->perm[q] < h->nlocal) == (target[q] < h->nlocal) &&
                         (h->perm[q] < h->nlocal || h->perm[q] == target[q]),
                  XYK_ERR_UNSUPPORTED, "relayout cannot move sharded qubits");
        if (h->perm[q] < h->nlocal) S.dst_of_src[h->perm[q]] = (uint8_t)target[q];
    }
    int rc = ensure_alt(h);
    if (rc) return rc;
    const uint64_t dim = 1ull << h->nlocal;
    if (h->dtype == 0)
        permute_kernel<double><<<grid_for(h, dim, 256), 256, 0, h->stream>>>(
            (const double2*)h->buf[h->cur], (double2*)h->buf[h->cur ^ 1], S);
    else
        permute_kernel<float><<<grid_for(h, dim, 256), 256, 0, h->stream>>>(
            (const float2*)h->buf[h->cur], (float2*)h->buf[h->cur ^ 1], S);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    XYK_CUDA(h, cudaGetLastError());
    h->cur ^= 1;
    h->perm = target;
    return XYK_OK;
}

void fill_desc(xyk_handle h, Program& prog, const std::vector<int>& loc2log, const PlanPass& pp,
               const std::vector<int>& perm_in, Desc& d, bool& lift_ok) {
    const Plan& plan = prog.plan;
    std::memset(&d, 0, sizeof(d));
    d.nq = plan.nq;
    d.nrounds = (int)pp.rounds.size();
    d.flags = (pp.fuse_load ? kPassFuseLoad : 0u) | (pp.fuse_store ? kPassFuseStore : 0u);
    // L2 prefetch of the CTA's next tile: issued when the last round starts (measured on B200: earlier
    // prefetches are partly evicted again and cost up to 35 % extra DRAM reads)
    d.flags |= kPassPfLate;
    for (int b = 0; b + kTB < plan.nq; ++b) d.hi_out[b] = (uint8_t)pp.out_of_in[kTB + b];
    // tile bits ordered by their output position (unfused store)
    int order[kTB];
    for (int k = 0; k < kTB; ++k) order[k] = k;
    std::sort(order, order + kTB, [&](int a, int b) { return pp.out_of_in[a] < pp.out_of_in[b]; });
    for (int m = 0; m < kTB; ++m) {
        d.st_lsrc[m] = (uint8_t)order[m];
        d.st_odst[m] = (uint8_t)pp.out_of_in[order[m]];
    }
    {
        const PlanRound& r0 = pp.rounds.front();
        for (int k = 0; k < (r0.split ? kRB + 1 : kRB); ++k) d.arr_reg[k] = arrival_stride_bytes(r0.regpos[k]);
    }
    const PlanRound& last = pp.rounds.back();
    const int last_nreg = last.split ? kRB + 1 : kRB;
    for (int b = 0; b < kTB - last_nreg; ++b) d.last_thr_out[b] = (uint8_t)pp.out_of_in[last.thrpos[b]];
    for (int k = 0; k < last_nreg; ++k) {
        d.last_reg_out[k] = (uint8_t)pp.out_of_in[last.regpos[k]];
        d.st_reg_idx[k] = pp.out_of_in[last.regpos[k]] < 32 ? (1u << pp.out_of_in[last.regpos[k]]) : 0u;  // local index bits only (< 2^32 amplitudes)
    }
    // fused diagonal phase: weights per INPUT physical bit, register table of round 0
    if (pp.has_phase && h->fuse_phase && !h->zterms.empty()) {
        d.flags |= kPassHasPhase;
        std::vector<double> wlog(h->n, 0.0), ph_w(kMaxQ, 0.0);  // ph_w: omega of the logical qubit on INPUT physical bit b
        for (const auto& z : h->zterms) wlog[z.i] = z.w;
        std::vector<char> is_local(h->n, 0);
        for (int l = 0; l < plan.nq; ++l) {
            ph_w[perm_in[l]] = wlog[loc2log[l]];
            is_local[loc2log[l]] = 1;
        }
        // sharded qubits: their bit is fixed by the rank (the caller guarantees h->perm is current for them)
        double pc = 0.0;
        for (int q = 0; q < h->n; ++q)
            if (!is_local[q]) pc += wlog[q] * (((h->rank >> (h->perm[q] - h->nlocal)) & 1) ? -1.0 : 1.0);
        const PlanRound& r0 = pp.rounds.front();   // a complex round (the schedule compiler puts the phase on one)
        for (int j = 0; j < (1 << kRB); ++j) {
            double ph = 0.0;
            for (int k = 0; k < kRB; ++k) ph += ph_w[r0.regpos[k]] * (((j >> k) & 1) ? -1.0 : 1.0);
            d.ph_reg[j] = make_double2(std::cos(pp.ztau * ph), std::sin(pp.ztau * ph));
        }
        // table factors of the thread bits (+ the constant rank part) and of the tile-id bits in two chunks
        const int nhi = plan.nq - kTB;
        d.ph_b1 = std::min(nhi, kPhTabBits);
        std::vector<double2> tab(kPhTabSize, make_double2(1.0, 0.0));
        for (int v = 0; v < (1 << (kTB - kRB)); ++v) {
            double ph = pc;
            for (int b = 0; b < kTB - kRB; ++b) ph += ph_w[r0.thrpos[b]] * (((v >> b) & 1) ? -1.0 : 1.0);
            tab[v] = make_double2(std::cos(pp.ztau * ph), std::sin(pp.ztau * ph));
        }
        for (int c = 0; c < 2; ++c) {
            const int lo = c == 0 ? 0 : d.ph_b1, hi = c == 0 ? d.ph_b1 : nhi;
            for (int v = 0; v < (1 << (hi - lo)); ++v) {
                double ph = 0.0;
                for (int b = lo; b < hi; ++b) ph += ph_w[kTB + b] * (((v >> (b - lo)) & 1) ? -1.0 : 1.0);
                tab[(c == 0 ? kPhTabLo : kPhTabHi) + v] = make_double2(std::cos(pp.ztau * ph), std::sin(pp.ztau * ph));
            }
        }
        void* dtab = nullptr;
        if (nhi - d.ph_b1 > kPhTabBits || cudaMalloc(&dtab, tab.size() * sizeof(double2)) != cudaSuccess ||
            // stream-ordered upload (a plain cudaMemcpy from pageable memory may return before its DMA has landed, and the
            // engine's stream does not wait for the legacy stream)
            cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess) {
            if (dtab) cudaFree(dtab);
            d.flags &= ~kPassHasPhase;  // falls back to the separate phase sweep (run_passes)
        } else {
            prog.dev_tabs.push_back(dtab);
            d.ph_tab = (const double2*)dtab;
        }
    }
    const std::vector<PlanRound>& rounds_used = pp.rounds;
    lift_ok = true;
    for (int r = 0; r < d.nrounds; ++r) {
        const PlanRound& pr = rounds_used[r];
        for (int s = 0; s < Desc::NSLOT; ++s)
            if ((pr.mask >> s) & 1u) {
                const double phi = std::remainder(pr.phi[s], 2.0 * M_PI);
                if (std::fabs(phi) > 1.5) lift_ok = false;
            }
    }
    for (int r = 0; r < d.nrounds; ++r) {
        const PlanRound& pr = rounds_used[r];
        Desc::Round& R = d.rounds[r];
        const int nreg = pr.split ? kRB + 1 : kRB;
        R.split = (uint16_t)(pr.split ? pr.nleft : 0);  // 0: complex round, else the left-group size of the split round
        for (int b = 0; b < kTB - nreg; ++b) R.thrpos[b] = (uint8_t)pr.thrpos[b];
        for (int k = 0; k < nreg; ++k) {
            R.regpos[k] = (uint8_t)pr.regpos[k];
            R.regsw[k] = (uint16_t)(16u * pad_slot(1u << pr.regpos[k]));
        }
        R.mask = pr.mask;
        R.cta_sync = (uint16_t)pr.cta_sync_after;
        for (int s = 0; s < Desc::NSLOT; ++s) {
            if (!((pr.mask >> s) & 1u)) continue;
            const double phi = std::remainder(pr.phi[s], 2.0 * M_PI);
            if (lift_ok) {
                R.ca[s] = -std::tan(phi / 2.0);
                R.cb[s] = std::sin(phi);
            } else {
                R.ca[s] = std::cos(phi);
                R.cb[s] = std::sin(phi);
            }
        }
    }
}

// tuning experiments (environment overrides of the schedule compiler's defaults): compiled only into the development
// build (-DXYK_EXPERIMENTS, `make experiments` -> tools/_build/libxyk_exp.so); the shipped library reads no environment
void apply_env(SchedConfig& cfg) {
#ifdef XYK_EXPERIMENTS
    if (const char* e = std::getenv("XYK_LANE_BITS")) cfg.lane_bits = std::atoi(e);
    if (const char* e = std::getenv("XYK_NEED_SHARED")) cfg.need_shared = std::atoi(e);
    if (const char* e = std::getenv("XYK_SPLIT")) cfg.split_rounds = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_TILE_GROUPS_OPEN")) cfg.tile_groups_open = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_TILE_GROUPS")) cfg.tile_groups = std::atoi(e);
    if (const char* e = std::getenv("XYK_TILE_BEAM")) cfg.tile_beam = std::atoi(e);
    if (const char* e = std::getenv("XYK_TILE_SEEDS")) cfg.tile_seeds = std::atoi(e);
    if (const char* e = std::getenv("XYK_BEAM")) cfg.beam_width = cfg.beam_cands = std::atoi(e);
    if (const char* e = std::getenv("XYK_RESTORE_GLOBALS")) cfg.restore_globals = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_FUSE_IO")) cfg.fuse_io = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_UNB_FIRST")) cfg.unb_first = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_UNB_LAST")) cfg.unb_last = std::atoi(e) != 0;
    if (const char* e = std::getenv("XYK_SPLIT_MIN_LEFT")) cfg.split_min_left = std::atoi(e);
    if (const char* e = std::getenv("XYK_SPLIT_MARGIN")) cfg.split_margin = std::atoi(e);
#else
    (void)cfg;
#endif
}

void finish_program(xyk_handle h, Program& prog) {
    prog.descs.resize(prog.plan.passes.size());
    prog.lift.resize(prog.plan.passes.size());
    std::vector<int> perm(prog.plan.start_perm);
    for (size_t p = 0; p < prog.plan.passes.size(); ++p) {
        bool lift = true;
        fill_desc(h, prog, prog.loc2log, prog.plan.passes[p], perm, prog.descs[p], lift);
        prog.lift[p] = lift ? 1 : 0;
        std::vector<int> np(perm);
        for (int q = 0; q < prog.plan.nq; ++q) np[q] = prog.plan.passes[p].out_of_in[perm[q]];
        perm.swap(np);
    }
}

std::shared_ptr<Program> get_program(xyk_handle h, double delta, int order, int reps) {
    const auto key = std::make_tuple(h->problem_version * 2 + (uint64_t)(h->fuse_phase ? 1 : 0), order, reps, delta);
    auto it = h->programs.find(key);
    if (it != h->programs.end()) return it->second;
    auto prog = std::make_shared<Program>();
    int l1 = 0;
    std::vector<Segment> segs = build_sequence(h->pairs, delta, order, reps, &l1);
    SchedConfig cfg;
    cfg.TB = kTB;
    cfg.RB = kRB;
    cfg.max_rounds = kMaxRounds;
    apply_env(cfg);
    prog->plan = compile_plan(h->nlocal, segs, cfg);
    prog->plan.layers1_equiv = l1;
    prog->loc2log.resize(h->nlocal);
    for (int l = 0; l < h->nlocal; ++l) prog->loc2log[l] = l;
    finish_program(h, *prog);
    if (h->programs.size() > 64) h->programs.clear();
    h->programs[key] = prog;
    return prog;
}

// the twelve instantiations of the pass kernel: shear-form rotations or not, peer (fused exchange) stores or not,
// and which unbalanced split round is compiled in: none (the lean kernel, ~30 % smaller), 2 | 4, or 1 | 5
using PassKernel = void (*)(const double2*, double2*, const Desc, uint64_t, const PeerPtrs);
// kPassGroups = 3: one CTA per SM with three tile groups in a token ring; 1: three free-running CTAs per SM
#ifndef XYK_PASS_GROUPS
#define XYK_PASS_GROUPS 3
#endif
constexpr int kPassGroups = XYK_PASS_GROUPS;
constexpr int kPassMinBlocks = kPassGroups == 1 ? 3 : 1;
template <int UNB>
PassKernel pass_kernel_unb(bool lift, bool peer) {
    if (lift) return peer ? tile_pass_kernel<kTB, kRB, true, kPassMinBlocks, true, UNB, kPassGroups> : tile_pass_kernel<kTB, kRB, true, kPassMinBlocks, false, UNB, kPassGroups>;
    return peer ? tile_pass_kernel<kTB, kRB, false, kPassMinBlocks, true, UNB, kPassGroups> : tile_pass_kernel<kTB, kRB, false, kPassMinBlocks, false, UNB, kPassGroups>;
}
// complex64 storage (8-byte amplitudes in global memory, FP64 on chip): single-device states only, so no peer variant
template <int UNB>
PassKernel pass_kernel_c64_unb(bool lift) {
    return lift ? tile_pass_kernel<kTB, kRB, true, kPassMinBlocks, false, UNB, kPassGroups, true> : tile_pass_kernel<kTB, kRB, false, kPassMinBlocks, false, UNB, kPassGroups, true>;
}
PassKernel pass_kernel_fn(bool lift, bool peer, int unb, bool c64 = false) {
    if (c64) return unb == 2 ? pass_kernel_c64_unb<2>(lift) : (unb == 1 ? pass_kernel_c64_unb<1>(lift) : pass_kernel_c64_unb<0>(lift));
    return unb == 2 ? pass_kernel_unb<2>(lift, peer) : (unb == 1 ? pass_kernel_unb<1>(lift, peer) : pass_kernel_unb<0>(lift, peer));
}
// 0: balanced rounds only; 1 / 2: the pass holds 1 | 5 / 2 | 4 split rounds (the schedule compiler never mixes the two)
int desc_unbalanced(const Desc& d) {
    int kind = 0;
#ifdef XYK_EXPERIMENTS
    static const char* force = std::getenv("XYK_FORCE_UNB");  // experiment: always run one of the larger kernels
    if (force) kind = std::atoi(force);
#endif
    for (int r = 0; r < d.nrounds; ++r)
        if (d.rounds[r].split == 1 || d.rounds[r].split == 2) kind = d.rounds[r].split;
    return kind;
}

int configure_pass_kernels(xyk_handle h) {
    const int smem = kPassGroups * (int)tile_bytes<kTB>();
    for (int k = 0; k < 12; ++k) XYK_CUDA(h, cudaFuncSetAttribute(pass_kernel_fn(k & 1, k & 2, k >> 2), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int k = 0; k < 6; ++k) XYK_CUDA(h, cudaFuncSetAttribute(pass_kernel_fn(k & 1, false, k >> 1, true), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    XYK_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pass_kernel_fn(true, false, 0), kPassGroups << (kTB - kRB), smem));
    if (occ < 1) occ = 1;
#if defined(XYK_TRACE) && defined(XYK_EXPERIMENTS)
    if (const char* e = std::getenv("XYK_PASS_OCC")) occ = std::max(1, std::min(occ, std::atoi(e)));  // development build only
    std::fprintf(stderr, "pass kernel occupancy: %d CTAs per SM\n", occ);
#endif
    h->pass_grid = occ * h->sm_count;
    return XYK_OK;
}

// prepare buffers / layout for a program (idempotent)
int begin_program(xyk_handle h, Program& prog) {
    int rc = ensure_alt(h);
    if (rc) return rc;
    if (!h->pass_grid) {
        rc = configure_pass_kernels(h);
        if (rc) return rc;
    }
    std::vector<int> target(h->perm);
    for (int l = 0; l < prog.plan.nq; ++l) target[prog.loc2log[l]] = prog.plan.start_perm[l];
    return relayout(h, target);
}

#ifdef XYK_TRACE
unsigned long long* g_trace_ptr = nullptr;
long g_trace_pass = -1;
size_t g_trace_words = 0;
#endif

// passes [p0, p1) of a program.  peer = true: the (single) pass scatters into the mapped buffers of
// every rank (the exchange fused into the store); the caller has fenced the ranks around it.
int run_passes(xyk_handle h, Program& prog, size_t p0, size_t p1, bool peer) {
    const Plan& plan = prog.plan;
    const uint64_t ntiles = 1ull << (h->nlocal - kTB);
    const int grid = (int)std::min<uint64_t>(ntiles, (uint64_t)h->pass_grid);  // few tiles: one per SM (group 0 of each CTA)
    const int smem = kPassGroups * (int)tile_bytes<kTB>();
    const int nthr = kPassGroups << (kTB - kRB);
    PeerPtrs peers{};
    if (peer)
        for (int r = 0; r < h->world; ++r) peers.p[r] = (double2*)h->peer_buf[r][h->cur ^ 1];
    int rc;
#ifdef XYK_EXPERIMENTS
    const char* stop_env = std::getenv("XYK_STOP_AFTER_PASS");  // debugging: run only the first k passes of a program
    const long stop_after = stop_env ? std::atol(stop_env) : -1;
#else
    constexpr long stop_after = -1;
#endif
    for (size_t p = p0; p < p1; ++p) {
        if (stop_after >= 0 && (long)p >= stop_after) break;
        const PlanPass& pp = plan.passes[p];
        if (pp.has_phase && !(prog.descs[p].flags & kPassHasPhase)) {
            rc = apply_phase(h, pp.ztau);
            if (rc) return rc;
        }
        if (pp.has_phase) {  // XXZ: all ZZ terms ride with the diagonal phase (grouped-diagonal splitting)
            rc = apply_zz_diag(h, pp.ztau);
            if (rc) return rc;
        }
        const double2* in = (const double2*)h->buf[h->cur];
        double2* out = (double2*)h->buf[h->cur ^ 1];
        if (h->stats.timing_enabled) {
            while (h->pass_events.size() < 2 * (p + 1)) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                h->pass_events.push_back(e);
            }
            cudaEventRecord(h->pass_events[2 * p], h->stream);
        }
#ifdef XYK_TRACE
        {   // development build only: the phase timeline of ONE pass (index g_trace_pass of the program) is recorded
            unsigned long long* tp = ((long)p == g_trace_pass) ? g_trace_ptr : nullptr;
            cudaMemcpyToSymbolAsync(xyk_trace_buf, &tp, sizeof(tp), 0, cudaMemcpyHostToDevice, h->stream);
        }
#endif
#ifdef XYK_EXPERIMENTS
        static const int ablate = std::getenv("XYK_ABLATE") ? std::atoi(std::getenv("XYK_ABLATE")) : 0;  // timing experiments only (WRONG results)
        if (ablate) {
            auto d = prog.descs[p];
            if (ablate & 1) for (int r = 0; r < d.nrounds; ++r) d.rounds[r].mask = 0;
            if (ablate & 2) d.nrounds = 1;
            if (ablate & 4) d.flags &= ~kPassHasPhase;
            if (ablate & 8) for (int r = 0; r < d.nrounds; ++r) d.rounds[r].cta_sync = 0;
            pass_kernel_fn(true, false, desc_unbalanced(d))<<<grid, nthr, smem, h->stream>>>(in, out, d, ntiles, peers);
        } else
#endif
        {
            pass_kernel_fn(prog.lift[p] != 0, peer, desc_unbalanced(prog.descs[p]), h->dtype != 0)<<<grid, nthr, smem, h->stream>>>(in, out, prog.descs[p], ntiles, peers);
        }
        if (h->stats.timing_enabled) cudaEventRecord(h->pass_events[2 * p + 1], h->stream);
        h->cur ^= 1;
        std::vector<int> np(h->perm);
        for (int l = 0; l < plan.nq; ++l) np[prog.loc2log[l]] = pp.out_of_in[h->perm[prog.loc2log[l]]];
        h->perm.swap(np);
        h->stats.kernel_launches++;
        h->stats.pass_launches++;
        h->stats.state_sweeps++;
    }
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

int end_program(xyk_handle h, Program& prog, size_t timed_passes) {
    const Plan& plan = prog.plan;
#ifdef XYK_EXPERIMENTS
    const bool stopped_early = std::getenv("XYK_STOP_AFTER_PASS") != nullptr;
#else
    constexpr bool stopped_early = false;
#endif
    if (plan.has_trailing_phase && !stopped_early) {
        int rc = apply_phase(h, plan.trailing_ztau);
        if (rc) return rc;
        rc = apply_zz_diag(h, plan.trailing_ztau);
        if (rc) return rc;
    }
    if (h->stats.timing_enabled) {
        cudaEventRecord(h->ev1, h->stream);
        cudaEventSynchronize(h->ev1);
        cudaEventElapsedTime(&h->stats.last_pass_ms, h->ev0, h->ev1);
        for (size_t p = 0; p < timed_passes; ++p) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->pass_events[2 * p], h->pass_events[2 * p + 1]);
            h->stats.pass_kernel_ms += ms;
            h->stats.pass_kernel_timed++;
#ifdef XYK_EXPERIMENTS
            static const bool dbg = std::getenv("XYK_DEBUG_PASSES") != nullptr;
#else
            constexpr bool dbg = false;
#endif
            if (dbg && p < plan.passes.size()) {
                const PlanPass& pp = plan.passes[p];
                std::fprintf(stderr, "pass %2zu: shared=%d fuse_load=%d fuse_store=%d rounds=%zu gates=%d lift=%d  %.3f ms  %.0f GB/s\n", p, pp.n_shared,
                             (int)pp.fuse_load, (int)pp.fuse_store, pp.rounds.size(), pp.ngates, (int)prog.lift[p], ms,
                             32.0 * (double)(1ull << h->nlocal) / (ms * 1e-3) / 1e9);
            }
        }
    }
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

int run_program_tiled(xyk_handle h, Program& prog) {
    int rc = begin_program(h, prog);
    if (rc) return rc;
    if (h->stats.timing_enabled) cudaEventRecord(h->ev0, h->stream);
    rc = run_passes(h, prog, 0, prog.plan.passes.size(), false);
    if (rc) return rc;
    rc = end_program(h, prog, prog.plan.passes.size());
    if (rc) return rc;
    h->stats.last_passes = (int)prog.plan.passes.size();
    h->stats.last_rounds = prog.plan.total_rounds;
    h->stats.last_gates = prog.plan.total_gates;
    h->stats.last_layers1 = prog.plan.layers1_equiv;
    return XYK_OK;
}

bool use_tiled(xyk_handle h) {
    if (h->dtype != 0 && h->world > 1) return false;  // complex64 states are single-device
    if (h->engine_opt == XYK_ENGINE_SIMPLE) return false;
    if (h->nlocal < kTB) return false;
    if (h->nlocal > 31) return false;  // the fused store of the pass kernel indexes amplitudes with 32-bit offsets (2 * 2^31 would wrap)
    if (h->engine_opt == XYK_ENGINE_TILED) return true;
    return h->nlocal >= 13;
}

template <typename real, int M>
void launch_obs_multi(xyk_handle h, const ObsSpec& S, int grid) {
    using C = typename cplx<real>::type;
    xy_expect_multi_kernel<real, M><<<grid, 256, 0, h->stream>>>((const C*)h->buf[h->cur], S, (double2*)h->d_partial);
}

// Rounds of one observable sweep: the tile-local positions in `want` (with their rows) are spread over register rounds of five;
// the three lowest thread bits of a round take one position of each pair {0,9}, {1,10}, {2,11} (conflict-free reads of the
// arrival layout), so a round never holds both positions of a pair in registers.
bool plan_obs_rounds(std::vector<std::pair<int, int>> want /* (tile-local position, row) */, ObsTileSpec& S) {
    S.nrounds = 0;
    auto has = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };
    auto partner = [](int p) { return p < 3 ? p + 9 : (p >= 9 ? p - 9 : -1); };
    while (!want.empty()) {
        if (S.nrounds == 3) return false;
        ObsTileSpec::Round& R = S.rounds[S.nrounds++];
        std::vector<int> regs, rows;
        for (auto it = want.begin(); it != want.end() && (int)regs.size() < 4;) {
            if (partner(it->first) >= 0 && has(regs, partner(it->first))) { ++it; continue; }
            regs.push_back(it->first);
            rows.push_back(it->second);
            it = want.erase(it);
        }
        for (int p = 3; (int)regs.size() < 4 && p < 9; ++p)   // fill with positions that are not measured here
            if (!has(regs, p)) {
                regs.push_back(p);
                rows.push_back(-1);
            }
        if ((int)regs.size() < 4) return false;
        std::vector<int> thr;
        for (int c = 0; c < 3; ++c) thr.push_back(!has(regs, c) ? c : c + 9);
        for (int p = 0; p < 12; ++p)
            if (!has(regs, p) && !has(thr, p)) thr.push_back(p);
        if (thr.size() != 8) return false;
        for (int k = 0; k < 4; ++k) {
            R.regpos[k] = (uint8_t)regs[k];
            R.row[k] = (int16_t)rows[k];
        }
        for (int b = 0; b < 8; ++b) R.thrpos[b] = (uint8_t)thr[b];
    }
    return true;
}

// <X_q>, <Y_q> for local qubits -> d_out[2q], d_out[2q+1]: twelve physical bits per sweep through shared-memory tiles
// (complex128, n_local >= 13): the contiguous low twelve bits first, then 12 - nb further bits per sweep next to a chunk of 2^nb
// contiguous amplitudes.  n_local = 30: three sweeps (the per-qubit loop of the reference reads the state thirty times).
int xy_expect_tiled_device(xyk_handle h) {
    const int nl = h->nlocal;
    std::vector<int> row_of_bit(nl, -1);
    for (int q = 0; q < h->n; ++q) {
        if (h->perm[q] < nl) row_of_bit[h->perm[q]] = q;
        else cudaMemsetAsync((double2*)h->d_partial + (size_t)q * kRedBlocks, 0, sizeof(double2) * kRedBlocks, h->stream);
    }
    const uint64_t ntiles = 1ull << (nl - 12);
    const int grid = (int)std::min<uint64_t>(ntiles, (uint64_t)3 * h->sm_count);
    const int smem = (int)arrival_bytes<12>();
    static bool attr_set = false;
    if (!attr_set) {
        XYK_CUDA(h, cudaFuncSetAttribute(xy_expect_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    const int rem = nl - 12;
    // chunk size of the strided sweeps: 2^nb amplitudes.  256-byte chunks (nb = 4, eight new bits per sweep) unless larger
    // chunks need no more sweeps; 128-byte chunks (nine bits per sweep) run slower than one more sweep (512 copies per tile)
    int nb = 4;
#ifdef XYK_EXPERIMENTS
    if (const char* e = std::getenv("XYK_OBS_NB")) nb = std::atoi(e);
    else
#endif
        for (int c = 8; c > 4; --c)
            if ((rem + (12 - c) - 1) / (12 - c) <= (rem + 7) / 8) { nb = c; break; }
    const int per = 12 - nb, nsweeps = 1 + (rem + per - 1) / per;
    int next_hi = 12;
    for (int sw = 0; sw < nsweeps; ++sw) {
        ObsTileSpec S{};
        S.nlocal = nl;
        S.stride = kRedBlocks;
        std::vector<std::pair<int, int>> want;
        if (sw == 0) {
            S.nb = 5;
            for (int p = 0; p < 12; ++p) {
                S.tile_pos[p] = (uint8_t)p;
                want.push_back({p, row_of_bit[p]});
            }
        } else {
            S.nb = nb;
            const int left = nl - next_hi, sweeps_left = nsweeps - sw;
            const int take = std::min(per, (left + sweeps_left - 1) / sweeps_left);   // even split over the remaining sweeps
            for (int p = 0; p < nb; ++p) S.tile_pos[p] = (uint8_t)p;
            std::vector<int> hi;
            for (int k = 0; k < take; ++k) hi.push_back(next_hi + k);
            for (int b = nb; (int)hi.size() < per; ++b)   // pad the tile with bits that are not measured in this sweep
                if (b < next_hi || b >= next_hi + take) hi.push_back(b);
            std::sort(hi.begin(), hi.end());
            for (int k = 0; k < per; ++k) {
                S.tile_pos[nb + k] = (uint8_t)hi[k];
                if (hi[k] >= next_hi && hi[k] < next_hi + take) want.push_back({nb + k, row_of_bit[hi[k]]});
            }
            next_hi += take;
        }
        XYK_CHECK(h, plan_obs_rounds(want, S), XYK_ERR_STATE, "observable sweep: no conflict-free round assignment");
        xy_expect_tile_kernel<<<grid, 128, smem, h->stream>>>((const double2*)h->buf[h->cur], S, (double2*)h->d_partial);
        h->stats.kernel_launches++;
        h->stats.state_sweeps++;
    }
    xy_finalize_kernel<<<h->n, 128, 0, h->stream>>>((const double2*)h->d_partial, grid, kRedBlocks, h->d_out);
    h->stats.kernel_launches++;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// <X_q>, <Y_q> for local qubits -> d_out[2q], d_out[2q+1] (already multiplied by 2).
// Five qubits per sweep (each thread holds the 32 amplitudes they span): ceil(n_local / 5) reads of the
// state instead of the n reads of the per-qubit loop (reference: pauli.rs:195-211 does n passes).
int xy_expect_device(xyk_handle h) {
    if (h->dtype == 0 && h->tiled_obs && h->nlocal >= 13) return xy_expect_tiled_device(h);
    const int grid_cap = kRedBlocks;
    std::vector<std::pair<int, int>> loc;  // (physical position, logical qubit)
    for (int q = 0; q < h->n; ++q) {
        if (h->perm[q] < h->nlocal) loc.push_back({h->perm[q], q});
        else cudaMemsetAsync((double2*)h->d_partial + (size_t)q * kRedBlocks, 0, sizeof(double2) * kRedBlocks, h->stream);
    }
    std::sort(loc.begin(), loc.end());
    const bool multi = h->tiled_obs && h->nlocal >= 10;
    // one grid size for every launch of this call (the finalize stage sums `grid` partials per row)
    const int Mmax = multi ? (int)std::min<size_t>(5, loc.size()) : 1;
    const int grid = std::max(1, std::min(grid_cap, grid_for(h, 1ull << (h->nlocal - Mmax), 256)));
    size_t i = 0;
    while (i < loc.size()) {
        const int M = multi ? (int)std::min<size_t>(5, loc.size() - i) : 1;
        ObsSpec S{};
        S.nlocal = h->nlocal;
        S.stride = kRedBlocks;
        for (int m = 0; m < M; ++m) {
            S.pos[m] = loc[i + m].first;
            S.qidx[m] = loc[i + m].second;
        }
        if (h->dtype == 0) {
            switch (M) {
                case 1: launch_obs_multi<double, 1>(h, S, grid); break;
                case 2: launch_obs_multi<double, 2>(h, S, grid); break;
                case 3: launch_obs_multi<double, 3>(h, S, grid); break;
                case 4: launch_obs_multi<double, 4>(h, S, grid); break;
                default: launch_obs_multi<double, 5>(h, S, grid); break;
            }
        } else {
            switch (M) {
                case 1: launch_obs_multi<float, 1>(h, S, grid); break;
                case 2: launch_obs_multi<float, 2>(h, S, grid); break;
                case 3: launch_obs_multi<float, 3>(h, S, grid); break;
                case 4: launch_obs_multi<float, 4>(h, S, grid); break;
                default: launch_obs_multi<float, 5>(h, S, grid); break;
            }
        }
        h->stats.kernel_launches++;
        i += M;
    }
    xy_finalize_kernel<<<h->n, 128, 0, h->stream>>>((const double2*)h->d_partial, grid, kRedBlocks, h->d_out);
    h->stats.kernel_launches++;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

// logical qubits currently sharded: rank bit k carries global_qubits[k]
std::vector<int> current_globals(xyk_handle h) {
    std::vector<int> g(h->gbits, -1);
    for (int q = 0; q < h->n; ++q)
        if (h->perm[q] >= h->nlocal) g[h->perm[q] - h->nlocal] = q;
    return g;
}

// place `outgoing[k]` on local physical bit nlocal - g + k (everything else keeps its relative order)
int stage_outgoing(xyk_handle h, const std::vector<int>& outgoing) {
    std::vector<int> target(h->perm);
    std::vector<char> is_out(h->n, 0);
    for (int q : outgoing) is_out[q] = 1;
    std::vector<std::pair<int, int>> locals;  // (current position, qubit)
    for (int q = 0; q < h->n; ++q)
        if (h->perm[q] < h->nlocal && !is_out[q]) locals.push_back({h->perm[q], q});
    std::sort(locals.begin(), locals.end());
    int pos = 0;
    for (auto& pq : locals) target[pq.second] = pos++;
    for (int q : outgoing) target[q] = pos++;
    return relayout(h, target);
}

int run_shard_steps(xyk_handle h) {
    while (h->shard_pos < h->shard_steps.size()) {
        ShardStep& st = h->shard_steps[h->shard_pos];
        if (!st.is_exchange) {
            Program& prog = *st.prog;
            const size_t P = prog.plan.passes.size();
            if (!st.peer_last) {
                int rc = run_program_tiled(h, prog);
                if (rc) return rc;
                h->shard_pos++;
                continue;
            }
            if (h->shard_phase == 0) {  // everything but the last pass, then fence the ranks
                int rc = begin_program(h, prog);
                if (rc) return rc;
                if (h->stats.timing_enabled) cudaEventRecord(h->ev0, h->stream);
                rc = run_passes(h, prog, 0, P - 1, false);
                if (rc) return rc;
                XYK_CUDA(h, cudaStreamSynchronize(h->stream));
                h->shard_phase = 1;
                return XYK_NEED_BARRIER;  // every rank is done reading the buffer its peers will overwrite
            }
            if (h->shard_phase == 1) {  // the pass that scatters into every rank's spare buffer
                const std::vector<int> in_glob = current_globals(h);
                int rc = run_passes(h, prog, P - 1, P, true);
                if (rc) return rc;
                for (int k = 0; k < h->gbits; ++k) h->perm[in_glob[k]] = st.rank_out[k];  // arrived from rank bit k
                XYK_CUDA(h, cudaStreamSynchronize(h->stream));
                h->exchanges++;
                h->shard_phase = 2;
                return XYK_NEED_BARRIER;  // every rank's stores have landed
            }
            if (h->stats.timing_enabled && P >= 1 && h->pass_events.size() >= 2 * P) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, h->pass_events[2 * (P - 1)], h->pass_events[2 * (P - 1) + 1]);
                h->stats.peer_pass_ms += ms;
                h->stats.peer_passes++;
            }
            int rc = end_program(h, prog, P);
            if (rc) return rc;
            h->shard_phase = 0;
            h->shard_pos++;
            continue;
        }
        for (int q : st.outgoing)
            XYK_CHECK(h, h->perm[q] < h->nlocal, XYK_ERR_STATE, "exchange: outgoing qubit is not local");
        int rc = stage_outgoing(h, st.outgoing);
        if (rc) return rc;
        rc = ensure_alt(h);
        if (rc) return rc;
        h->ex_out = st.outgoing;
        h->ex_in = current_globals(h);
        h->exchange_pending = true;
        XYK_CUDA(h, cudaStreamSynchronize(h->stream));  // the host's collective runs on another stream
        return XYK_NEED_EXCHANGE;
    }
    if (h->shard_passes > 0) {  // the per-program bookkeeping of run_program_tiled saw only the last local program
        h->stats.last_gates = h->shard_gates;
        h->stats.last_rounds = h->shard_rounds;
        h->stats.last_passes = h->shard_passes;
        h->shard_gates = h->shard_rounds = h->shard_passes = 0;
    }
    h->shard_steps.clear();
    h->shard_pos = 0;
    h->shard_phase = 0;
    return XYK_OK;
}

int build_shard_steps(xyk_handle h, const std::vector<Segment>& segs, int l1, double delta, int order, int reps) {
    SchedConfig cfg;
    cfg.TB = kTB;
    cfg.RB = kRB;
    cfg.max_rounds = kMaxRounds;
    cfg.peer_exchange = h->peers_ready && h->peer_exchange_opt;
    cfg.restore_globals = true;
    apply_env(cfg);
    // a sharded program depends on which logical qubit sits on which rank bit on entry; calls that start from
    // the same sharding (every call, when the program restores it) reuse the compiled steps
    std::string key;
    {
        std::ostringstream os;
        os.precision(17);
        os << h->problem_version << ':' << (h->fuse_phase ? 1 : 0) << ':' << (cfg.peer_exchange ? 1 : 0) << ':' << order << ':' << reps << ':' << delta;
        for (int q : current_globals(h)) os << ':' << q;
        key = os.str();
    }
    auto hit = h->shard_cache.find(key);
    if (hit != h->shard_cache.end()) {
        h->shard_steps = hit->second.steps;
        h->shard_pos = 0;
        h->stats.last_gates = h->shard_gates = hit->second.gates;
        h->stats.last_rounds = h->shard_rounds = hit->second.rounds;
        h->stats.last_passes = h->shard_passes = hit->second.passes;
        h->stats.last_layers1 = l1;
        return XYK_OK;
    }
    h->shard_compiles++;
    std::vector<ShardOp> ops;
    try {
        ops = compile_sharded(h->n, current_globals(h), segs, cfg, nullptr);
    } catch (const std::exception& e) {
        return fail(h, XYK_ERR_INVALID, e.what());
    }
    h->shard_steps.clear();
    h->shard_pos = 0;
    // perm of the sharded qubits evolves with the exchanges: descriptors need the rank-bit position of
    // every global qubit at the time the pass runs, so track it while building
    std::vector<int> saved_perm(h->perm);
    int gates = 0, rounds = 0, passes = 0;
    for (auto& op : ops) {
        ShardStep st;
        st.is_exchange = op.is_exchange;
        if (op.is_exchange) {
            st.outgoing = op.outgoing;
            st.incoming = op.incoming;
            for (int k = 0; k < h->gbits; ++k) {
                h->perm[op.outgoing[k]] = h->nlocal + k;                 // leaves to rank bit k
                h->perm[op.incoming[k]] = h->nlocal - h->gbits + k;      // arrives on the top local bits
            }
        } else {
            st.prog = std::make_shared<Program>();
            st.prog->plan = std::move(op.plan);
            st.prog->plan.layers1_equiv = 0;
            st.prog->loc2log = op.loc2log;
            finish_program(h, *st.prog);
            if (op.peer_last) {
                st.peer_last = true;
                st.rank_out = op.rank_out;
                uint64_t ob = 0;
                for (int k = 0; k < h->gbits; ++k)
                    if ((h->rank >> k) & 1) ob |= 1ull << op.rank_out[k];
                st.prog->descs.back().out_base = ob;
                // descriptor phase constants of later programs need the post-exchange rank bits
                for (int k = 0; k < h->gbits; ++k) {
                    h->perm[op.outgoing_fused[k]] = h->nlocal + k;
                    h->perm[op.incoming_fused[k]] = op.rank_out[k];
                }
            }
            gates += st.prog->plan.total_gates;
            rounds += st.prog->plan.total_rounds;
            passes += (int)st.prog->plan.passes.size();
        }
        h->shard_steps.push_back(std::move(st));
    }
    h->perm = saved_perm;
    h->stats.last_gates = h->shard_gates = gates;
    h->stats.last_rounds = h->shard_rounds = rounds;
    h->stats.last_passes = h->shard_passes = passes;
    h->stats.last_layers1 = l1;
    if (h->shard_cache.size() > 64) h->shard_cache.clear();
    xyk_handle_s::ShardCacheEntry& ce = h->shard_cache[key];
    ce.steps = h->shard_steps;
    ce.gates = gates;
    ce.rounds = rounds;
    ce.passes = passes;
    return XYK_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int xyk_version(void) { return 100; }

const char* xyk_last_error(xyk_handle h) { return h ? h->err.c_str() : g_global_error.c_str(); }

int xyk_device_count(int* count) {
    if (!count) return XYK_ERR_INVALID;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return fail(nullptr, XYK_ERR_CUDA, cudaGetErrorString(e));
    }
    return XYK_OK;
}

int xyk_device_memory(int device, size_t* free_bytes, size_t* total_bytes) {
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMemGetInfo(free_bytes, total_bytes);
    if (e != cudaSuccess) return fail(nullptr, XYK_ERR_CUDA, cudaGetErrorString(e));
    return XYK_OK;
}

int xyk_create(int n_qubits, int dtype, int device, int world_size, int rank, xyk_handle* out) {
    if (!out) return fail(nullptr, XYK_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > kMaxQ) return fail(nullptr, XYK_ERR_INVALID, "n_qubits out of range");
    if (dtype != 0 && dtype != 1) return fail(nullptr, XYK_ERR_INVALID, "dtype must be 0 (complex128) or 1 (complex64)");
    if (world_size < 1 || (world_size & (world_size - 1)) || rank < 0 || rank >= world_size)
        return fail(nullptr, XYK_ERR_INVALID, "world_size must be a power of two and 0 <= rank < world_size");
    int gbits = 0;
    while ((1 << gbits) < world_size) ++gbits;
    if (n_qubits - gbits < 1) return fail(nullptr, XYK_ERR_INVALID, "too few qubits for this world size");
    if (n_qubits - gbits > 34) return fail(nullptr, XYK_ERR_INVALID, "too many local qubits");
    std::unique_ptr<xyk_handle_s> h(new xyk_handle_s);
    h->n = n_qubits;
    h->gbits = gbits;
    h->nlocal = n_qubits - gbits;
    h->world = world_size;
    h->rank = rank;
    h->dtype = dtype;
    h->device = device;
    h->perm.resize(n_qubits);
    for (int q = 0; q < n_qubits; ++q) h->perm[q] = q;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, XYK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, XYK_ERR_CUDA, cudaGetErrorString(e));
    h->sm_count = prop.multiProcessorCount;
    h->bytes = ((size_t)1 << h->nlocal) * amp_bytes(dtype);
    e = cudaMalloc(&h->buf[0], h->bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, e == cudaErrorMemoryAllocation ? XYK_ERR_NOMEM : XYK_ERR_CUDA,
                    std::string("cudaMalloc(state): ") + cudaGetErrorString(e));
    }
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_tab, sizeof(double2) * 3 * kTabStride) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_J, sizeof(double) * kMaxQ * kMaxQ) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_zztab, sizeof(double2) * 4096) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_partial, sizeof(double) * kPartialDoubles) == cudaSuccess;
    ok = ok && cudaMalloc(&h->d_out, sizeof(double) * 4096) == cudaSuccess;
    ok = ok && cudaMallocHost(&h->h_out, sizeof(double) * 8192) == cudaSuccess;
    ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
    if (!ok) {
        std::string msg = std::string("scratch allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        xyk_destroy(h.release());
        return fail(nullptr, XYK_ERR_NOMEM, msg);
    }
    *out = h.release();
    return XYK_OK;
}

int xyk_destroy(xyk_handle h) {
    if (!h) return XYK_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int r = 0; r < 8; ++r)
        for (int w = 0; w < 2; ++w)
            if (h->peer_open[r][w]) cudaIpcCloseMemHandle(h->peer_buf[r][w]);
    for (int b = 0; b < 2; ++b)
        if (h->buf[b]) cudaFree(h->buf[b]);
    if (h->d_tab) cudaFree(h->d_tab);
    if (h->d_J) cudaFree(h->d_J);
    if (h->d_zztab) cudaFree(h->d_zztab);
    if (h->d_partial) cudaFree(h->d_partial);
    if (h->d_out) cudaFree(h->d_out);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t e : h->pass_events) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return XYK_OK;
}

int xyk_set_option(xyk_handle h, int option, int64_t value) {
    if (!h) return XYK_ERR_INVALID;
    switch (option) {
        case XYK_OPT_ENGINE:
            XYK_CHECK(h, value >= 0 && value <= 2, XYK_ERR_INVALID, "unknown engine");
            h->engine_opt = (int)value;
            return XYK_OK;
        case XYK_OPT_FUSE_PHASE:
            h->fuse_phase = value != 0;
            return XYK_OK;
        case XYK_OPT_TILED_OBS:
            h->tiled_obs = value != 0;
            return XYK_OK;
        case XYK_OPT_PEER_EXCHANGE:
            h->peer_exchange_opt = value != 0;
            return XYK_OK;
        case 100:
            h->debug_mode = (int)value;
            h->programs.clear();
    h->shard_cache.clear();
            return XYK_OK;
        default:
            return fail(h, XYK_ERR_INVALID, "unknown option");
    }
}

int xyk_set_problem(xyk_handle h, const double* K, const double* omega) {
    if (!h || !K || !omega) return XYK_ERR_INVALID;
    for (int i = 0; i < h->n * h->n; ++i) XYK_CHECK(h, std::isfinite(K[i]), XYK_ERR_INVALID, "K must be finite");
    for (int i = 0; i < h->n; ++i) XYK_CHECK(h, std::isfinite(omega[i]), XYK_ERR_INVALID, "omega must be finite");
    build_terms(h->n, K, omega, h->pairs, h->zterms, h->Ksym);
    h->omega.assign(omega, omega + h->n);
    h->have_problem = true;
    h->problem_version++;
    h->programs.clear();
    h->shard_cache.clear();
    h->zz_delta = 0.0;
    return XYK_OK;
}

int xyk_set_problem_xxz(xyk_handle h, const double* K, const double* omega, double delta) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, std::isfinite(delta), XYK_ERR_INVALID, "delta must be finite");
    XYK_CHECK(h, h->world == 1 || std::fabs(delta) <= kSparsityEps, XYK_ERR_UNSUPPORTED, "XXZ anisotropy on a sharded state is not supported");
    int rc = xyk_set_problem(h, K, omega);
    if (rc) return rc;
    if (std::fabs(delta) <= kSparsityEps) return XYK_OK;  // the reference drops the ZZ terms below the sparsity threshold
    XYK_CUDA(h, cudaSetDevice(h->device));
    std::vector<double> J(h->Ksym);
    for (double& v : J) v *= delta;
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));  // a ZZ sweep in flight still reads the previous table
    XYK_CUDA(h, cudaMemcpyAsync(h->d_J, J.data(), sizeof(double) * J.size(), cudaMemcpyHostToDevice, h->stream));   // stream-ordered
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));   // J is a local vector
    h->zz_delta = delta;
    return XYK_OK;
}

int xyk_term_counts(xyk_handle h, int* n_pairs, int* n_z) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->have_problem, XYK_ERR_STATE, "no problem set");
    if (n_pairs) *n_pairs = (int)h->pairs.size();
    if (n_z) *n_z = (int)h->zterms.size();
    return XYK_OK;
}

// product state of the problem's frequencies, written directly in `layout` (logical -> physical; null: canonical order)
static int init_product_in_layout(xyk_handle h, const std::vector<int>* layout) {
    XYK_CHECK(h, h->have_problem, XYK_ERR_STATE, "no problem set");
    XYK_CUDA(h, cudaSetDevice(h->device));
    // keep global bits where they are, local qubits in canonical order
    for (int q = 0; q < h->n; ++q) h->perm[q] = layout ? (*layout)[q] : q;
    int rc = h->dtype == 0 ? launch_init<double>(h) : launch_init<float>(h);
    if (rc) return rc;
    h->state_valid = true;
    return XYK_OK;
}

int xyk_init_product_state(xyk_handle h) {
    if (!h) return XYK_ERR_INVALID;
    return init_product_in_layout(h, nullptr);
}

int xyk_init_product_angles(xyk_handle h, const double* angles) {
    if (!h || !angles) return XYK_ERR_INVALID;
    for (int q = 0; q < h->n; ++q) XYK_CHECK(h, std::isfinite(angles[q]), XYK_ERR_INVALID, "angles must be finite");
    XYK_CUDA(h, cudaSetDevice(h->device));
    for (int q = 0; q < h->n; ++q) h->perm[q] = q;
    int rc = h->dtype == 0 ? launch_init<double>(h, angles) : launch_init<float>(h, angles);
    if (rc) return rc;
    h->state_valid = true;
    return XYK_OK;
}

int xyk_apply_layers(xyk_handle h, double delta, int order, int reps) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->have_problem, XYK_ERR_STATE, "no problem set");
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, order == 1 || order == 2 || order == 4 || order == 6, XYK_ERR_INVALID, "order must be 1, 2, 4 or 6");
    XYK_CHECK(h, reps >= 1, XYK_ERR_INVALID, "reps must be >= 1");
    XYK_CHECK(h, std::isfinite(delta), XYK_ERR_INVALID, "delta must be finite");
    XYK_CUDA(h, cudaSetDevice(h->device));
    if (h->world > 1) {
        XYK_CHECK(h, h->dtype == 0 && h->nlocal >= kTB && h->nlocal <= 31, XYK_ERR_UNSUPPORTED,
                  "sharded evolution needs complex128 and 12 to 31 local qubits");
        XYK_CHECK(h, !h->exchange_pending && h->shard_steps.empty(), XYK_ERR_STATE, "a sharded call is still in flight");
        int l1s = 0;
        std::vector<Segment> ssegs = build_sequence(h->pairs, delta, order, reps, &l1s);
        int rcs = build_shard_steps(h, ssegs, l1s, delta, order, reps);
        if (rcs) return rcs;
        return run_shard_steps(h);
    }
    if (use_tiled(h)) {
        std::shared_ptr<Program> prog;
        try {
            prog = get_program(h, delta, order, reps);
        } catch (const std::exception& e) {
            return fail(h, XYK_ERR_INVALID, e.what());
        }
        return run_program_tiled(h, *prog);
    }
    int l1 = 0;
    std::vector<Segment> segs = build_sequence(h->pairs, delta, order, reps, &l1);
    int gates = 0;
    for (auto& s : segs) gates += (int)s.gates.size();
    int rc = run_segments_simple(h, segs);
    h->stats.last_passes = 0;
    h->stats.last_rounds = 0;
    h->stats.last_gates = gates;
    h->stats.last_layers1 = l1;
    return rc;
}

int xyk_continue(xyk_handle h) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->world > 1, XYK_ERR_UNSUPPORTED, "not a sharded handle");
    XYK_CHECK(h, !h->exchange_pending, XYK_ERR_STATE, "an exchange is pending: call xyk_exchange_done first");
    XYK_CUDA(h, cudaSetDevice(h->device));
    return run_shard_steps(h);
}

int xyk_exchange_info(xyk_handle h, void** src, void** dst, size_t* bytes_per_block) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->exchange_pending, XYK_ERR_STATE, "no exchange pending");
    if (src) *src = h->buf[h->cur];
    if (dst) *dst = h->buf[h->cur ^ 1];
    if (bytes_per_block) *bytes_per_block = h->bytes / (size_t)h->world;
    return XYK_OK;
}

int xyk_exchange_done(xyk_handle h) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->exchange_pending, XYK_ERR_STATE, "no exchange pending");
    h->cur ^= 1;
    for (int k = 0; k < h->gbits; ++k) {
        h->perm[h->ex_out[k]] = h->nlocal + k;
        h->perm[h->ex_in[k]] = h->nlocal - h->gbits + k;
    }
    h->exchange_pending = false;
    h->exchanges++;
    h->stats.state_sweeps++;
    if (h->shard_pos < h->shard_steps.size() && h->shard_steps[h->shard_pos].is_exchange) h->shard_pos++;
    return XYK_OK;
}

int xyk_ipc_export(xyk_handle h, int which, unsigned char* handle64) {
    if (!h || !handle64 || which < 0 || which > 1) return XYK_ERR_INVALID;
    XYK_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_alt(h);
    if (rc) return rc;
    cudaIpcMemHandle_t mh;
    XYK_CUDA(h, cudaIpcGetMemHandle(&mh, h->buf[which]));
    static_assert(sizeof(mh) == 64, "CUDA IPC handle is 64 bytes");
    std::memcpy(handle64, &mh, 64);
    return XYK_OK;
}

int xyk_ipc_import(xyk_handle h, int peer_rank, int which, const unsigned char* handle64) {
    if (!h || !handle64 || which < 0 || which > 1 || peer_rank < 0 || peer_rank >= h->world || peer_rank >= 8)
        return XYK_ERR_INVALID;
    XYK_CUDA(h, cudaSetDevice(h->device));
    if (peer_rank == h->rank) {
        h->peer_buf[peer_rank][which] = h->buf[which];
    } else {
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, handle64, 64);
        void* p = nullptr;
        XYK_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->peer_buf[peer_rank][which] = p;
        h->peer_open[peer_rank][which] = true;
    }
    bool all = true;
    for (int r = 0; r < h->world; ++r)
        for (int w = 0; w < 2; ++w) all &= h->peer_buf[r][w] != nullptr;
    h->peers_ready = all;
    return XYK_OK;
}

int xyk_begin_exchange(xyk_handle h, const int* outgoing) {
    if (!h || !outgoing) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->world > 1, XYK_ERR_UNSUPPORTED, "not a sharded handle");
    XYK_CHECK(h, !h->exchange_pending && h->shard_steps.empty(), XYK_ERR_STATE, "a sharded call is still in flight");
    XYK_CUDA(h, cudaSetDevice(h->device));
    ShardStep st;
    st.is_exchange = true;
    st.outgoing.assign(outgoing, outgoing + h->gbits);
    h->shard_steps.push_back(st);
    h->shard_pos = 0;
    return run_shard_steps(h);
}

int xyk_xy_expectations(xyk_handle h, double* exp_x, double* exp_y) {
    if (!h || !exp_x || !exp_y) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CUDA(h, cudaSetDevice(h->device));
    int rc = xy_expect_device(h);
    if (rc) return rc;
    XYK_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * 2 * h->n, cudaMemcpyDeviceToHost, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < h->n; ++q) {
        exp_x[q] = h->h_out[2 * q];
        exp_y[q] = h->h_out[2 * q + 1];
    }
    return XYK_OK;
}

int xyk_xy_expectations_peer(xyk_handle h, double* exp_x, double* exp_y) {
    if (!h || !exp_x || !exp_y) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world > 1 && h->peers_ready, XYK_ERR_UNSUPPORTED, "peer reads need a sharded handle with mapped peer buffers");
    XYK_CHECK(h, h->dtype == 0, XYK_ERR_UNSUPPORTED, "sharded states are complex128");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const uint64_t dim = 1ull << h->nlocal;
    const int grid = std::min(kRedBlocks, grid_for(h, dim, 256));
    XYK_CUDA(h, cudaMemsetAsync(h->d_partial, 0, sizeof(double2) * (size_t)h->n * kRedBlocks, h->stream));
    for (int q = 0; q < h->n; ++q) {
        if (h->perm[q] < h->nlocal) continue;
        const int b = h->perm[q] - h->nlocal;
        const int upper = (h->rank >> b) & 1;   // both ranks of the pair work: the lower half of the indices here, the upper half there
        const double2* peer = (const double2*)h->peer_buf[h->rank ^ (1 << b)][h->cur];
        XYK_CHECK(h, peer != nullptr, XYK_ERR_STATE, "peer buffer not mapped");
        const uint64_t k0 = upper ? dim / 2 : 0, k1 = upper ? dim : dim / 2;
        peer_dot_kernel<<<grid, 256, 0, h->stream>>>((const double2*)h->buf[h->cur], peer, k0, k1, upper, (double2*)h->d_partial + (size_t)q * kRedBlocks);
        h->stats.kernel_launches++;
        h->stats.state_sweeps++;
    }
    xy_finalize_kernel<<<h->n, 128, 0, h->stream>>>((const double2*)h->d_partial, grid, kRedBlocks, h->d_out);
    h->stats.kernel_launches++;
    XYK_CUDA(h, cudaGetLastError());
    XYK_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * 2 * h->n, cudaMemcpyDeviceToHost, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < h->n; ++q) {
        exp_x[q] = h->h_out[2 * q];
        exp_y[q] = h->h_out[2 * q + 1];
    }
    return XYK_OK;
}

int xyk_order_parameter(xyk_handle h, double* R, double* psi) {
    if (!h || !R) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "order parameter of a sharded state is assembled by the host");
    std::vector<double> ex(h->n), ey(h->n);
    int rc = xyk_xy_expectations(h, ex.data(), ey.data());
    if (rc) return rc;
    double zr = 0.0, zi = 0.0;
    for (int q = 0; q < h->n; ++q) {
        zr += ex[q];
        zi += ey[q];
    }
    zr /= h->n;
    zi /= h->n;
    *R = std::sqrt(zr * zr + zi * zi);
    if (psi) *psi = std::atan2(zi, zr);
    return XYK_OK;
}

int xyk_z_expectations(xyk_handle h, double* exp_z) {
    if (!h || !exp_z) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const uint64_t dim = 1ull << h->nlocal;
    // power-of-two grid (z_expect_kernel): 2^T threads with nlocal - T <= 16 high bits left to the per-amplitude bookkeeping
    int grid = 1;
    while (grid * 2 <= kRedBlocks && (uint64_t)grid * 2 * 256 <= dim) grid *= 2;
    while (h->nlocal - (31 - __builtin_clz((unsigned)(grid * 256))) > 16) grid *= 2;   // very large shards: more blocks than kRedBlocks
    XYK_CHECK(h, (size_t)grid * (kMaxQ + 1) <= (size_t)kPartialDoubles, XYK_ERR_UNSUPPORTED, "z expectations: shard too large for the partial buffer");
    if (h->dtype == 0)
        z_expect_kernel<double><<<grid, 256, 0, h->stream>>>((const double2*)h->buf[h->cur], h->nlocal, h->d_partial);
    else
        z_expect_kernel<float><<<grid, 256, 0, h->stream>>>((const float2*)h->buf[h->cur], h->nlocal, h->d_partial);
    reduce_partials_kernel<<<kMaxQ + 1, 128, 0, h->stream>>>(h->d_partial, grid, kMaxQ + 1, h->d_out);
    h->stats.kernel_launches += 2;
    XYK_CUDA(h, cudaGetLastError());
    XYK_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * (kMaxQ + 1), cudaMemcpyDeviceToHost, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    const double tot = h->h_out[kMaxQ];
    for (int q = 0; q < h->n; ++q) {
        const int pos = h->perm[q];
        if (pos < h->nlocal) exp_z[q] = tot - 2.0 * h->h_out[pos];
        else exp_z[q] = ((h->rank >> (pos - h->nlocal)) & 1) ? -tot : tot;
    }
    return XYK_OK;
}

int xyk_norm2(xyk_handle h, double* norm2) {
    if (!h || !norm2) return XYK_ERR_INVALID;
    std::vector<double> z(h->n);
    int rc = xyk_z_expectations(h, z.data());
    if (rc) return rc;
    *norm2 = h->h_out[kMaxQ];
    return XYK_OK;
}

int xyk_correlation_matrix_xy(xyk_handle h, double* C) {
    if (!h || !C) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "single-GPU handles only");
    XYK_CHECK(h, h->n >= 2 || true, XYK_ERR_INVALID, "");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n;
    for (int i = 0; i < n * n; ++i) C[i] = 0.0;
    if (n < 2) return XYK_OK;
    const uint64_t quarter = 1ull << (h->nlocal - 2);
    const int grid = std::min(kRedBlocks, grid_for(h, quarter, 256));
    // batches of pairs: partials for pair index t at d_partial + t*kRedBlocks
    const int max_batch = kPartialDoubles / kRedBlocks;
    std::vector<std::pair<int, int>> pairs;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) pairs.push_back({i, j});
    for (size_t off = 0; off < pairs.size(); off += max_batch) {
        const int cnt = (int)std::min<size_t>(max_batch, pairs.size() - off);
        XYK_CHECK(h, cnt <= 4096, XYK_ERR_UNSUPPORTED, "too many pairs per batch");
        for (int t = 0; t < cnt; ++t) {
            const int pi = h->perm[pairs[off + t].first], pj = h->perm[pairs[off + t].second];
            if (h->dtype == 0)
                xxyy_expect_kernel<double><<<grid, 256, 0, h->stream>>>((const double2*)h->buf[h->cur], h->nlocal, pi, pj,
                                                                      h->d_partial + (size_t)t * kRedBlocks);
            else
                xxyy_expect_kernel<float><<<grid, 256, 0, h->stream>>>((const float2*)h->buf[h->cur], h->nlocal, pi, pj,
                                                                     h->d_partial + (size_t)t * kRedBlocks);
            h->stats.kernel_launches++;
        }
        reduce_rows_kernel<<<cnt, 128, 0, h->stream>>>(h->d_partial, grid, kRedBlocks, h->d_out);
        h->stats.kernel_launches++;
        XYK_CUDA(h, cudaGetLastError());
        XYK_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->stream));
        XYK_CUDA(h, cudaStreamSynchronize(h->stream));
        for (int t = 0; t < cnt; ++t) {
            const int i = pairs[off + t].first, j = pairs[off + t].second;
            C[i * n + j] = C[j * n + i] = 4.0 * h->h_out[t];
        }
    }
    return XYK_OK;
}

int xyk_energy(xyk_handle h, double* E) {
    if (!h || !E) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->have_problem, XYK_ERR_STATE, "no problem set");
    const int n = h->n;
    double e_zz = 0.0;   // XXZ: -sum K_ij delta <Z_i Z_j>, one sweep over the diagonal quadratic form
    if (h->zz_delta != 0.0 && !h->pairs.empty()) {
        XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "XXZ energy of a sharded state is not supported");
        XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
        XYK_CUDA(h, cudaSetDevice(h->device));
        ZZSpec S{};
        S.nbits = h->n;
        S.nlocal = h->nlocal;
        S.rank = h->rank;
        S.LB = std::min(h->nlocal, 12);
        S.t = 1.0;
        for (int q = 0; q < h->n; ++q) S.log_of_bit[h->perm[q]] = q;
        double* qtab = reinterpret_cast<double*>(h->d_zztab);   // 4096 doubles fit the 4096 double2 entries
        zz_low_qtable_kernel<<<16, 256, 0, h->stream>>>(qtab, h->d_J, h->n, S);
        const uint64_t nblk = 1ull << (h->nlocal - S.LB);
        const int grid = (int)std::min<uint64_t>(nblk, (uint64_t)kRedBlocks);
        if (h->dtype == 0)
            zz_energy_kernel<double><<<grid, 256, 0, h->stream>>>((const double2*)h->buf[h->cur], qtab, h->d_J, h->n, S, h->d_partial);
        else
            zz_energy_kernel<float><<<grid, 256, 0, h->stream>>>((const float2*)h->buf[h->cur], qtab, h->d_J, h->n, S, h->d_partial);
        reduce_partials_kernel<<<1, 128, 0, h->stream>>>(h->d_partial, grid, 1, h->d_out);
        h->stats.kernel_launches += 3;
        h->stats.state_sweeps += 1;
        XYK_CUDA(h, cudaGetLastError());
        XYK_CUDA(h, cudaMemcpyAsync(h->h_out, h->d_out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        XYK_CUDA(h, cudaStreamSynchronize(h->stream));
        e_zz = -h->h_out[0];   // d_J holds K_ij * delta: the term is -K_ij delta Z_i Z_j
    }
    std::vector<double> C((size_t)n * n), z(n);
    int rc = xyk_correlation_matrix_xy(h, C.data());
    if (rc) return rc;
    rc = xyk_z_expectations(h, z.data());
    if (rc) return rc;
    double e = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) e -= h->Ksym[(size_t)i * n + j] * C[(size_t)i * n + j];
    for (int i = 0; i < n; ++i) e -= h->omega[i] * z[i];
    *E = e + e_zz;
    return XYK_OK;
}

int xyk_apply_occupation_decay(xyk_handle h, double rate_zero, double rate_one, double scale) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "occupation decay of a sharded state is not supported");
    XYK_CHECK(h, std::isfinite(rate_zero) && std::isfinite(rate_one) && std::isfinite(scale), XYK_ERR_INVALID, "rates and scale must be finite");
    XYK_CUDA(h, cudaSetDevice(h->device));
    // psi_k *= scale * exp(-rate_zero * #zeros(k) - rate_one * #ones(k)): three host-built chunk tables (real factors) through
    // the table-apply kernel of the diagonal phase; the factor does not depend on the layout (it counts bits)
    const int nl = h->nlocal;
    int b1 = std::min(nl, 11), b2 = std::min(nl, 22);
    if (nl - b2 > 11) {
        b1 = (nl + 2) / 3;
        b2 = b1 + (nl - b1 + 1) / 2;
    }
    XYK_CHECK(h, b1 <= 11 && b2 - b1 <= 11 && nl - b2 <= 11, XYK_ERR_UNSUPPORTED, "decay tables support at most 33 local qubits");
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));   // the pinned staging buffer is shared with other calls
    std::vector<double2> tab((size_t)3 * kTabStride, make_double2(1.0, 0.0));
    const int lo[3] = {0, b1, b2}, hi[3] = {b1, b2, nl};
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < (1 << (hi[c] - lo[c])); ++v) {
            const int ones = __builtin_popcount((unsigned)v), zeros = (hi[c] - lo[c]) - ones;
            tab[(size_t)c * kTabStride + v] = make_double2((c == 0 ? scale : 1.0) * std::exp(-rate_zero * zeros - rate_one * ones), 0.0);
        }
    // stream-ordered upload: a plain cudaMemcpy from pageable memory may return before its DMA has landed, and the engine's
    // stream is non-blocking (it does not wait for the legacy stream) -- the kernel below would read the previous table
    XYK_CUDA(h, cudaMemcpyAsync(h->d_tab, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice, h->stream));
    const uint64_t dim = 1ull << nl;
    if (h->dtype == 0)
        phase_apply_kernel<double><<<grid_for(h, dim, 256), 256, 0, h->stream>>>((double2*)h->buf[h->cur], h->d_tab, kTabStride, nl, b1, b2);
    else
        phase_apply_kernel<float><<<grid_for(h, dim, 256), 256, 0, h->stream>>>((float2*)h->buf[h->cur], h->d_tab, kTabStride, nl, b1, b2);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    XYK_CUDA(h, cudaGetLastError());
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));   // d_tab is rewritten by the next phase / init call
    return XYK_OK;
}

int xyk_apply_jump(xyk_handle h, int qubit, int kind) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "jump operators on a sharded state are not supported");
    XYK_CHECK(h, qubit >= 0 && qubit < h->n && (kind == 0 || kind == 1), XYK_ERR_INVALID, "bad qubit or jump kind");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const uint64_t half = 1ull << (h->nlocal - 1);
    if (h->dtype == 0)
        jump_kernel<double><<<grid_for(h, half, 256), 256, 0, h->stream>>>((double2*)h->buf[h->cur], h->nlocal, h->perm[qubit], kind);
    else
        jump_kernel<float><<<grid_for(h, half, 256), 256, 0, h->stream>>>((float2*)h->buf[h->cur], h->nlocal, h->perm[qubit], kind);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

int xyk_apply_ry(xyk_handle h, const double* angles) {
    if (!h || !angles) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "single-qubit rotations of a sharded state are not supported");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const uint64_t half = 1ull << (h->nlocal - 1);
    for (int q = 0; q < h->n; ++q) {
        XYK_CHECK(h, std::isfinite(angles[q]), XYK_ERR_INVALID, "angles must be finite");
        if (angles[q] == 0.0) continue;
        const double c = std::cos(angles[q] / 2.0), s = std::sin(angles[q] / 2.0);
        if (h->dtype == 0)
            ry_gate_kernel<double><<<grid_for(h, half, 256), 256, 0, h->stream>>>((double2*)h->buf[h->cur], h->nlocal, h->perm[q], c, s);
        else
            ry_gate_kernel<float><<<grid_for(h, half, 256), 256, 0, h->stream>>>((float2*)h->buf[h->cur], h->nlocal, h->perm[q], c, s);
        h->stats.kernel_launches++;
        h->stats.state_sweeps++;
    }
    XYK_CUDA(h, cudaGetLastError());
    return XYK_OK;
}

int xyk_exact_step(xyk_handle h, double dt, double tol, int* terms_used) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->have_problem, XYK_ERR_STATE, "no problem set");
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "the Chebyshev step runs on a single-device state");
    XYK_CHECK(h, h->dtype == 0, XYK_ERR_UNSUPPORTED, "the Chebyshev step needs a complex128 state");
    XYK_CHECK(h, std::isfinite(dt) && std::isfinite(tol) && tol > 0.0, XYK_ERR_INVALID, "dt must be finite and tol positive");
    if (terms_used) *terms_used = 0;
    if (dt == 0.0) return XYK_OK;
    XYK_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_alt(h);
    if (rc) return rc;
    // spectral bound: ||K (XX + YY)|| = 2 |K|, ||K delta ZZ|| = |K delta|, ||w Z|| = |w|
    double E = 0.0;
    for (const auto& p : h->pairs) E += (2.0 + std::fabs(h->zz_delta)) * std::fabs(p.k);
    for (const auto& z : h->zterms) E += std::fabs(z.w);
    if (E == 0.0) return XYK_OK;  // H = 0
    E *= 1.0 + 1e-9;
    const double x = E * std::fabs(dt);
    // exp(-i H dt) = sum_k c_k T_k(H / E),  c_k = (2 - [k = 0]) (-i sgn dt)^k J_k(E |dt|)
    std::vector<double> J;
    for (int k = 0; k < 100000; ++k) {
        J.push_back(std::cyl_bessel_j((double)k, x));
        if (k > x + 2 && std::fabs(J[k]) < 0.1 * tol && std::fabs(J[k - 1]) < 0.1 * tol) break;
    }
    const int nterms = (int)J.size();
    auto coef = [&](int k) {
        const double mag = (k == 0 ? 1.0 : 2.0) * J[k];
        const int q = k & 3;  // (-i s)^k, s = sgn(dt)
        const double sg = dt > 0 ? 1.0 : -1.0;
        switch (q) {
            case 0: return make_double2(mag, 0.0);
            case 1: return make_double2(0.0, -sg * mag);
            case 2: return make_double2(-mag, 0.0);
            default: return make_double2(0.0, sg * mag);
        }
    };
    std::vector<HPair> hp;
    for (const auto& p : h->pairs) hp.push_back({1ull << h->perm[p.i], 1ull << h->perm[p.j], p.k, p.k * h->zz_delta});
    std::vector<double> wbit(h->n, 0.0);
    for (const auto& z : h->zterms) wbit[h->perm[z.i]] = z.w;
    HPair* d_pairs = nullptr;
    double* d_w = nullptr;
    double2* acc = nullptr;
    auto cleanup = [&]() {
        if (d_pairs) cudaFree(d_pairs);
        if (d_w) cudaFree(d_w);
        if (acc) cudaFree(acc);
    };
    cudaError_t e = cudaMalloc(&acc, h->bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, XYK_ERR_NOMEM, "xyk_exact_step: no device memory for the accumulator (three state vectors are needed)");
    }
    bool ok = cudaMalloc(&d_pairs, sizeof(HPair) * std::max<size_t>(1, hp.size())) == cudaSuccess &&
              cudaMalloc(&d_w, sizeof(double) * h->n) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_pairs, hp.data(), sizeof(HPair) * hp.size(), cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(d_w, wbit.data(), sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    if (!ok) {
        cleanup();
        return fail(h, XYK_ERR_CUDA, std::string("xyk_exact_step: ") + cudaGetErrorString(cudaGetLastError()));
    }
    const uint64_t dim = 1ull << h->nlocal;
    const int grid = grid_for(h, dim, 256);
    const size_t smem = sizeof(HPair) * hp.size() + sizeof(double) * h->n;
    double2* X = (double2*)h->buf[h->cur];       // phi_{k-1}, overwritten by phi_{k+1}
    double2* Y = (double2*)h->buf[h->cur ^ 1];   // phi_k
    const int np = (int)hp.size();
    // phi_1 = H phi_0 / E;  acc = c_0 phi_0 + c_1 phi_1
    cheb_step_kernel<<<grid, 256, smem, h->stream>>>(X, nullptr, Y, acc, d_pairs, np, d_w, h->n, dim, 1.0 / E, coef(1), coef(0), 1);
    h->stats.kernel_launches++;
    h->stats.state_sweeps++;
    for (int k = 1; k + 1 < nterms; ++k) {
        // phi_{k+1} = 2 H phi_k / E - phi_{k-1}  (written over phi_{k-1});  acc += c_{k+1} phi_{k+1}
        cheb_step_kernel<<<grid, 256, smem, h->stream>>>(Y, X, X, acc, d_pairs, np, d_w, h->n, dim, 2.0 / E, coef(k + 1), make_double2(0, 0), 0);
        std::swap(X, Y);
        h->stats.kernel_launches++;
        h->stats.state_sweeps++;
    }
    e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        cleanup();
        return fail(h, XYK_ERR_CUDA, std::string("xyk_exact_step: ") + cudaGetErrorString(e));
    }
    // the accumulator becomes the live state (layout unchanged)
    void* old_state = h->buf[h->cur];
    h->buf[h->cur] = acc;
    acc = reinterpret_cast<double2*>(old_state);
    cleanup();
    if (terms_used) *terms_used = nterms;
    return XYK_OK;
}

int xyk_run(xyk_handle h, const double* step_sizes, int m, int reps, int order, double* R_out) {
    if (!h || !R_out || (m > 0 && !step_sizes)) return XYK_ERR_INVALID;
    XYK_CHECK(h, m >= 0, XYK_ERR_INVALID, "m must be >= 0");
    XYK_CHECK(h, h->world == 1, XYK_ERR_UNSUPPORTED, "xyk_run drives a single-device state (sharded states are driven by the host)");
    // the state is prepared directly in the start layout of the first step's pass program: no re-layout sweep before it
    std::vector<int> start_layout;
    const bool formula_ok = (order == 1 || order == 2 || order == 4 || order == 6) && reps >= 1 && std::isfinite(m > 0 ? step_sizes[0] : 0.0);
    if (m > 0 && formula_ok && h->have_problem && use_tiled(h)) {   // (bad arguments are reported by xyk_apply_layers below)
        try {
            std::shared_ptr<Program> prog = get_program(h, step_sizes[0], order, reps);
            if (!prog->plan.passes.empty()) {
                start_layout.assign(h->n, 0);
                for (int l = 0; l < prog->plan.nq; ++l) start_layout[prog->loc2log[l]] = prog->plan.start_perm[l];
            }
        } catch (const std::exception& e) {
            return fail(h, XYK_ERR_INVALID, e.what());
        }
    }
    int rc = init_product_in_layout(h, start_layout.empty() ? nullptr : &start_layout);
    if (rc) return rc;
    // the whole trajectory is enqueued without a host round trip: R(t) stays on the device until the end
    double* d_R = nullptr;
    XYK_CUDA(h, cudaMalloc(&d_R, sizeof(double) * (size_t)(m + 1)));
    auto record = [&](int s) -> int {
        int r = xy_expect_device(h);
        if (r) return r;
        order_param_kernel<<<1, 32, 0, h->stream>>>(h->d_out, h->n, d_R + s);
        h->stats.kernel_launches++;
        return XYK_OK;
    };
    rc = record(0);
    for (int s = 0; s < m && rc == XYK_OK; ++s) {
        rc = xyk_apply_layers(h, step_sizes[s], order, reps);
        if (rc == XYK_OK) rc = record(s + 1);
    }
    if (rc == XYK_OK) {
        cudaError_t e = cudaMemcpyAsync(R_out, d_R, sizeof(double) * (size_t)(m + 1), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(h, XYK_ERR_CUDA, std::string("xyk_run: ") + cudaGetErrorString(e));
    } else {
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(d_R);
    return rc;
}

int xyk_get_state(xyk_handle h, void* dst, int is_device) {
    if (!h || !dst) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CUDA(h, cudaSetDevice(h->device));
    const void* src = h->buf[h->cur];
    if (h->world > 1) {
        int rcc = xyk_canonicalise(h);
        if (rcc) return rcc;
        src = h->buf[h->cur];
    } else if (!perm_is_identity(h)) {
        // canonical copy into the spare buffer; the live state keeps its layout
        PermSpec S{};
        S.nlocal = h->nlocal;
        for (int b = 0; b < kMaxQ; ++b) S.dst_of_src[b] = (uint8_t)b;
        for (int q = 0; q < h->n; ++q) {
            XYK_CHECK(h, (h->perm[q] < h->nlocal) == (q < h->nlocal), XYK_ERR_UNSUPPORTED,
                      "sharded state must be canonicalised first");
            if (h->perm[q] < h->nlocal) S.dst_of_src[h->perm[q]] = (uint8_t)q;
        }
        int rc = ensure_alt(h);
        if (rc) return rc;
        const uint64_t dim = 1ull << h->nlocal;
        if (h->dtype == 0)
            permute_kernel<double><<<grid_for(h, dim, 256), 256, 0, h->stream>>>((const double2*)h->buf[h->cur],
                                                                               (double2*)h->buf[h->cur ^ 1], S);
        else
            permute_kernel<float><<<grid_for(h, dim, 256), 256, 0, h->stream>>>((const float2*)h->buf[h->cur],
                                                                              (float2*)h->buf[h->cur ^ 1], S);
        h->stats.kernel_launches++;
        XYK_CUDA(h, cudaGetLastError());
        src = h->buf[h->cur ^ 1];
    }
    XYK_CUDA(h, cudaMemcpyAsync(dst, src, h->bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    return XYK_OK;
}

int xyk_set_state(xyk_handle h, const void* src, int is_device) {
    if (!h || !src) return XYK_ERR_INVALID;
    XYK_CUDA(h, cudaSetDevice(h->device));
    XYK_CUDA(h, cudaMemcpyAsync(h->buf[h->cur], src, h->bytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int q = 0; q < h->n; ++q) h->perm[q] = q;
    h->state_valid = true;
    return XYK_OK;
}

int xyk_state_buffer(xyk_handle h, void** ptr, size_t* bytes) {
    if (!h) return XYK_ERR_INVALID;
    if (ptr) *ptr = h->buf[h->cur];
    if (bytes) *bytes = h->bytes;
    return XYK_OK;
}

int xyk_get_layout(xyk_handle h, int* perm) {
    if (!h || !perm) return XYK_ERR_INVALID;
    for (int q = 0; q < h->n; ++q) perm[q] = h->perm[q];
    return XYK_OK;
}

int xyk_canonicalise(xyk_handle h) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CHECK(h, h->state_valid, XYK_ERR_STATE, "state not initialised");
    XYK_CUDA(h, cudaSetDevice(h->device));
    std::vector<int> target(h->perm);
    if (h->world == 1) {
        for (int q = 0; q < h->n; ++q) target[q] = q;
    } else {  // local qubits in ascending order on the local bits; sharded qubits stay where they are
        int pos = 0;
        for (int q = 0; q < h->n; ++q)
            if (h->perm[q] < h->nlocal) target[q] = pos++;
    }
    return relayout(h, target);
}

int xyk_synchronize(xyk_handle h) {
    if (!h) return XYK_ERR_INVALID;
    XYK_CUDA(h, cudaSetDevice(h->device));
    XYK_CUDA(h, cudaStreamSynchronize(h->stream));
    return XYK_OK;
}

int xyk_stream(xyk_handle h, void** stream) {
    if (!h || !stream) return XYK_ERR_INVALID;
    *stream = (void*)h->stream;
    return XYK_OK;
}

int xyk_get_stats(xyk_handle h, xyk_stats* out) {
    if (!h || !out) return XYK_ERR_INVALID;
    *out = h->stats;
    return XYK_OK;
}

int xyk_enable_timing(xyk_handle h, int on) {
    if (!h) return XYK_ERR_INVALID;
    h->stats.timing_enabled = on ? 1 : 0;
    return XYK_OK;
}

int64_t xyk_plan_describe(int n_local, int dtype, const double* K, const double* omega, int n, double delta,
                          int order, int reps, char* buf, int64_t cap) {
    if (n < 1 || n > kMaxQ || n_local < 1 || n_local > n || !K || !omega) return -1;
    try {
        std::vector<PairTerm> pairs;
        std::vector<ZTerm> z;
        std::vector<double> Ks;
        build_terms(n, K, omega, pairs, z, Ks);
        int l1 = 0;
        std::vector<Segment> segs = build_sequence(pairs, delta, order, reps, &l1);
        SchedConfig cfg;
        cfg.TB = kTB;
        cfg.RB = kRB;
        cfg.max_rounds = kMaxRounds;
            apply_env(cfg);
        std::string js;
        if (n_local < n) {  // sharded program: the g highest logical qubits start on the rank bits
            std::vector<int> glob;
            for (int q = n_local; q < n; ++q) glob.push_back(q);
            cfg.peer_exchange = (dtype & 2) != 0;  // test hook: bit 1 of `dtype` selects fused (peer-store) exchanges
            cfg.restore_globals = (dtype & 4) != 0;  //            bit 2: end with the sharding the program started from
            js = sharded_to_json(compile_sharded(n, glob, segs, cfg, nullptr));
        } else {
            Plan plan = compile_plan(n_local, segs, cfg);
            plan.layers1_equiv = l1;
            js = plan_to_json(plan);
        }
        if (buf && cap > (int64_t)js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
        return (int64_t)js.size() + 1;
    } catch (const std::exception& e) {
        g_global_error = e.what();
        return -1;
    }
}

int64_t xyk_plan_stages(const double* K, const double* omega, int n, double delta, int order, int reps, int upto_stage, char* buf,
                        int64_t cap) {
    if (n < 1 || n > kMaxQ || !K || !omega) return -1;
    try {
        std::vector<PairTerm> pairs;
        std::vector<ZTerm> z;
        std::vector<double> Ks;
        build_terms(n, K, omega, pairs, z, Ks);
        std::vector<Segment> segs = build_sequence(pairs, delta, order, reps, nullptr);
        SchedConfig cfg;
        cfg.TB = kTB;
        cfg.RB = kRB;
        cfg.max_rounds = kMaxRounds;
        apply_env(cfg);
        const std::string js = plan_stages_to_json(n, segs, cfg, upto_stage);
        if (buf && cap > (int64_t)js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
        return (int64_t)js.size() + 1;
    } catch (const std::exception& e) {
        g_global_error = e.what();
        return -1;
    }
}

int xyk_batch_run(int device, int n, int B, const double* K, const double* omega, const double* step_sizes, int m,
                  int reps, int order, double* R_out, void* final_states) {
    return small_batch_run(device, n, B, K, omega, step_sizes, m, reps, order, R_out, final_states, g_global_error);
}

}  // extern "C"

#ifdef XYK_TRACE
// development build only (tools/trace_pass.py): record the per-warp phase timeline of pass `pass_index`
extern "C" int xyk_debug_trace_begin(int pass_index, int k0, int nk) {
    g_trace_words = (size_t)148 * 3 * 4 * kTraceMaxE;
    if (!g_trace_ptr && cudaMalloc(&g_trace_ptr, g_trace_words * 8) != cudaSuccess) return XYK_ERR_CUDA;
    cudaMemset(g_trace_ptr, 0, g_trace_words * 8);
    cudaMemcpyToSymbol(xyk_trace_k0, &k0, sizeof(int));
    cudaMemcpyToSymbol(xyk_trace_nk, &nk, sizeof(int));
    g_trace_pass = pass_index;
    return cudaDeviceSynchronize() == cudaSuccess ? XYK_OK : XYK_ERR_CUDA;
}
extern "C" int64_t xyk_debug_trace_read(unsigned long long* dst, int64_t max_words) {
    if (!g_trace_ptr) return -1;
    cudaDeviceSynchronize();
    const size_t nw = std::min<size_t>((size_t)max_words, g_trace_words);
    if (cudaMemcpy(dst, g_trace_ptr, nw * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)nw;
}
#endif
