/* xyk.h -- C ABI of the B200-native Kuramoto-XY exact-statevector engine.
 *
 * Drop-in boundary for the reference's statevector route
 *   auto_solve -> QuantumKuramotoSolver(n, K, omega).run(t_max, dt, trotter_per_step)
 * (reference: src/scpn_quantum_control/phase/backend_selector.py:288-293,
 *  src/scpn_quantum_control/phase/xy_kuramoto.py:86-258).  Each entry point below names the
 * reference interface it replaces.  Plain pointers and sizes only: no torch / numpy types.
 *
 * Conventions
 *   - qubit q <-> bit q of the basis index (little endian), amplitudes interleaved (re, im).
 *   - dtype 0 = complex128 (double re, im), dtype 1 = complex64 (float re, im).
 *   - K is row-major double[n*n], omega is double[n]; the library symmetrises K as (K+K^T)/2,
 *     ignores its diagonal and drops |K_ij| < 1e-15, |omega_i| <= 1e-15 exactly like
 *     bridge/knm_hamiltonian.py:172-195.  All *validation* with the reference's error
 *     messages is done by the Python host layer before calling in.
 *   - every function returns 0 on success; a non-zero status has a message retrievable with
 *     xyk_last_error(handle) (or xyk_last_error(NULL) for failures of xyk_create).
 *   - a handle is not thread safe; all work is enqueued on a handle-owned CUDA stream and the
 *     calls that return host data synchronise that stream.
 *   - caller owns every host buffer for the duration of the call; the library owns device
 *     state inside the handle.
 */
#ifndef XYK_H_
#define XYK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xyk_handle_s *xyk_handle;

/* status codes */
#define XYK_OK 0
#define XYK_ERR_INVALID 1     /* bad argument                                   */
#define XYK_ERR_CUDA 2        /* CUDA runtime failure                           */
#define XYK_ERR_NOMEM 3       /* device allocation failed (maps to MemoryError) */
#define XYK_ERR_STATE 4       /* call sequence error (e.g. no problem set)      */
#define XYK_ERR_UNSUPPORTED 5 /* configuration not supported by this build      */
#define XYK_NEED_EXCHANGE 100 /* sharded run: host must perform the all-to-all  */
#define XYK_NEED_BARRIER 101  /* sharded run: host must fence all ranks (device sync + barrier), then xyk_continue */

/* engines (xyk_set_option(h, XYK_OPT_ENGINE, v)) */
#define XYK_ENGINE_AUTO 0   /* tiled for n_local >= 13, gate-by-gate below          */
#define XYK_ENGINE_SIMPLE 1 /* one kernel per gate, canonical layout                */
#define XYK_ENGINE_TILED 2  /* fused cache-blocked passes (shared-memory tiles)     */

#define XYK_OPT_ENGINE 1
#define XYK_OPT_FUSE_PHASE 2   /* 1 (default): Z phase folded into the first pass of a layer */
#define XYK_OPT_TILED_OBS 3    /* 1 (default): five qubits per observable sweep              */
#define XYK_OPT_PEER_EXCHANGE 4 /* 1 (default): fuse exchanges into the preceding pass (peer stores over NVLink)
                                  when the ranks' buffers have been mapped with xyk_ipc_export/import */

int xyk_version(void);
const char *xyk_last_error(xyk_handle h);
int xyk_device_count(int *count);
/* free / total bytes of device `device` */
int xyk_device_memory(int device, size_t *free_bytes, size_t *total_bytes);

/* Create an engine for n_qubits on CUDA device `device`.
 * world_size / rank: 1 / 0 for a single GPU.  For a sharded state (world_size = 2,4,8) this
 * handle owns the amplitudes whose top log2(world_size) *physical* index bits equal `rank`.
 * Replaces: QuantumKuramotoSolver.__init__ state ownership (xy_kuramoto.py:86-106) and the
 * dense budget gate's allocation (dense_budget.py:141-172; budget is checked by the host). */
int xyk_create(int n_qubits, int dtype, int device, int world_size, int rank, xyk_handle *out);
int xyk_destroy(xyk_handle h);
int xyk_set_option(xyk_handle h, int option, int64_t value);

/* Build the ordered, thresholded term list (gate order) from (K, omega).
 * Replaces: knm_to_hamiltonian / knm_to_xxz_hamiltonian(delta=0)
 * (bridge/knm_hamiltonian.py:141-217).  May be called again at any time to re-parameterise
 * the couplings without touching the state (control/realtime_feedback.py:98-137). */
int xyk_set_problem(xyk_handle h, const double *K, const double *omega);
/* XXZ anisotropy: H = -sum K_ij (XX + YY + delta ZZ) - sum omega_i Z_i.  Replaces: knm_to_xxz_hamiltonian(K, omega, delta)
 * (bridge/knm_hamiltonian.py:141-206; |delta| <= 1e-15 drops the ZZ terms like the reference).  The gate-by-gate engine
 * applies XX.YY then ZZ per pair in the reference's term order; the tiled engine groups all ZZ terms with the diagonal
 * Z phase of each (half) layer -- the same operator for orders >= 2 up to the splitting error of the formula, and
 * exact under the 6th-order composition used for evolution="reference".  Single-GPU handles. */
int xyk_set_problem_xxz(xyk_handle h, const double *K, const double *omega, double delta);

/* number of retained pair terms / Z terms of the current problem */
int xyk_term_counts(xyk_handle h, int *n_pairs, int *n_z);

/* psi0 = prod_i Ry(omega_i mod 2pi)|0>, generated on the device.
 * Replaces: xy_kuramoto.py:236-242 (Statevector.from_instruction of the Ry circuit). */
int xyk_init_product_state(xyk_handle h);

/* |psi> = prod_q Ry(angles[q])|0>, qubit 0 = least significant bit; no problem needs to be set.  angles = 0: |0...0>
 * (Statevector.from_label("0" * n), scripts/bench_s2_scaling_lite.py:262), angles = pi/2: |+>^n (the Floquet drive's
 * start, phase/floquet_kuramoto.py:131).  Works on sharded handles (every rank writes its shard). */
int xyk_init_product_angles(xyk_handle h, const double *angles);

/* Evolve one grid interval `delta` with `reps` repetitions of the product formula of `order`:
 *   order 1 = LieTrotter, 2 = SuzukiTrotter(order=2) (Qiskit term order), 4 / 6 = Suzuki
 *   compositions of the order-2 layer (the route to the exact per-step propagation the
 *   reference's run() numerically returns).
 * Replaces: sv.evolve(solver.evolve(step_dt, trotter_per_step))  (xy_kuramoto.py:132-154,246-247).
 * Sharded handles may return XYK_NEED_EXCHANGE: the host then performs the all-to-all
 * described by xyk_exchange_info(), calls xyk_exchange_done() and calls xyk_continue(). */
int xyk_apply_layers(xyk_handle h, double delta, int order, int reps);
int xyk_continue(xyk_handle h);
/* all-to-all description: send `src`, receive into `dst`; both hold 2^n_local amplitudes split
 * in world_size equal contiguous blocks (block s goes to / comes from rank s). */
int xyk_exchange_info(xyk_handle h, void **src, void **dst, size_t *bytes_per_block);
int xyk_exchange_done(xyk_handle h);
/* Request an exchange outside apply_layers (observables on sharded qubits, canonical layouts):
 * the g = log2(world_size) logical qubits outgoing[0..g) leave, the sharded ones arrive.
 * Returns XYK_NEED_EXCHANGE; the host then proceeds as above. */
int xyk_begin_exchange(xyk_handle h, const int *outgoing);
/* Map the two state buffers of every rank into this process (CUDA IPC, 64-byte handles exchanged by
 * the host).  Once all ranks' buffers are imported, an exchange that follows a pass is executed BY
 * that pass: its store scatters directly into the destination ranks' buffers over NVLink; the C ABI
 * then returns XYK_NEED_BARRIER before and after that pass instead of XYK_NEED_EXCHANGE. */
int xyk_ipc_export(xyk_handle h, int which, unsigned char *handle64);
int xyk_ipc_import(xyk_handle h, int peer_rank, int which, const unsigned char *handle64);

/* <X_q>, <Y_q> for every qubit (partial sums over this shard for sharded handles: the host
 * adds the per-rank arrays).  Qubits that are currently global on a sharded handle are
 * reported through need_exchange != 0 (see python host).
 * Replaces: scpn_quantum_engine.all_xy_expectations (scpn_quantum_engine/src/pauli.rs:180-214). */
int xyk_xy_expectations(xyk_handle h, double *exp_x, double *exp_y);
/* Sharded handles with mapped peer buffers: this rank's partial sums of <X_q>, <Y_q> for the SHARDED qubits (0 for local
 * qubits), computed by reading half of the partner shard through the peer mapping (NVLink; the two ranks of a pair split the index
 * range) -- no qubit exchange.  Every rank calls
 * it between two fences (all ranks idle on their state buffers); the sum over ranks is the expectation value. */
int xyk_xy_expectations_peer(xyk_handle h, double *exp_x, double *exp_y);

/* R = |sum_q (<X_q> + i<Y_q>)| / n, psi = arg.  Replaces: measure_order_parameter
 * (xy_kuramoto.py:156-185), state_order_param_sparse (pauli.rs:72-116),
 * order_param_from_statevector (sectors.rs:67-99).  Single-GPU handles only. */
int xyk_order_parameter(xyk_handle h, double *R, double *psi);
/* <Z_q> for every qubit.  Replaces: expectation_pauli_fast(.., pauli=2) (pauli.rs:122-173). */
int xyk_z_expectations(xyk_handle h, double *exp_z);
/* C_ij = <X_iX_j + Y_iY_j>, row-major double[n*n].  Replaces: correlation_matrix_xy
 * (scpn_quantum_engine/src/sectors.rs:105-145).  Single-GPU handles only. */
int xyk_correlation_matrix_xy(xyk_handle h, double *C);
/* <H> = -sum K_ij C_ij - sum omega_i <Z_i>.  Replaces: energy_expectation (xy_kuramoto.py:260-264). */
int xyk_energy(xyk_handle h, double *E);
/* sum_k |psi_k|^2 over this shard */
int xyk_norm2(xyk_handle h, double *norm2);

/* psi <- prod_q Ry_q(angles[q]) psi on the live state (Ry(t) = [[cos t/2, -sin t/2], [sin t/2, cos t/2]]); angles of
 * exactly 0 are skipped.  Single-GPU handles.  Replaces: the per-qubit `qc.ry(...)` correction layer that
 * RealtimeSyncFeedbackController applies with Statevector.evolve (control/realtime_feedback.py:209-227). */
int xyk_apply_ry(xyk_handle h, const double *angles);

/* psi <- exp(-i H dt) psi to `tol` (per amplitude, in the expansion coefficients) WITHOUT a product formula: Chebyshev
 * expansion of the propagator, one fused H.phi sweep per term (all pair terms summed in one kernel; XXZ anisotropy
 * included), Bessel coefficients J_k(E |dt|) on the host, E = 2 sum |K| (+ |delta| sum |K|) + sum |omega|.  Needs room for
 * three state vectors.  *terms_used (may be NULL) returns the number of Chebyshev terms.
 * Replaces: the exact sibling of the path, classical_exact_evolution's expm_multiply step
 * (hardware/classical.py:201-208, hardware/fast_classical.py:35-97) -- and gives the engine a second, independent
 * route to what Statevector.evolve returns under the reference's pinned Qiskit (Mode E).  Single-GPU, complex128. */
int xyk_exact_step(xyk_handle h, double dt, double tol, int *terms_used);

/* Quantum-jump (Monte Carlo wave function) building blocks on the live state; single-GPU handles.
 * Replaces: the non-Hermitian part of H_eff = H - (i/2) sum L^dagger L and the jump operators of mcwf_trajectory
 * (phase/tensor_jump.py:157-206).  For the reference's operators -- sqrt(gamma_amp) |1><0| and sqrt(gamma_deph / 2) Z per
 * qubit -- sum L^dagger L is diagonal in the computational basis and commutes with H (H conserves the number of ones), so
 * exp(-i H_eff dt) = exp(-i H dt) followed by a real factor per amplitude that depends on its bit counts only:
 *   xyk_apply_occupation_decay: psi_k *= scale * exp(-rate_zero * #zero bits(k) - rate_one * #one bits(k));
 *   xyk_apply_jump: kind 0: psi <- |1><0|_q psi (the reference's "-" matrix), kind 1: psi <- Z_q psi. */
int xyk_apply_occupation_decay(xyk_handle h, double rate_zero, double rate_one, double scale);
int xyk_apply_jump(xyk_handle h, int qubit, int kind);

/* The whole run() loop on the device: init, R[0], then for each step_sizes[s]: evolve, R[s+1].
 * Replaces: QuantumKuramotoSolver.run hot loop (xy_kuramoto.py:236-248). Single GPU only. */
int xyk_run(xyk_handle h, const double *step_sizes, int m, int reps, int order, double *R_out);

/* Copy the state out / in, in canonical (logical) index order.
 * is_device = 0: dst/src are host pointers; 1: device pointers on the handle's device.
 * dst/src hold 2^n_local amplitudes of the handle's dtype.  Sharded handles must be in the
 * canonical layout (use xyk_canonicalise first; may require exchanges).
 * Replaces: np.asarray(sv.data) (xy_kuramoto.py:167) / Statevector(data). */
int xyk_get_state(xyk_handle h, void *dst, int is_device);
int xyk_set_state(xyk_handle h, const void *src, int is_device);
/* raw access for zero-copy wrappers (torch tensors as buffers): current buffer, physical layout */
int xyk_state_buffer(xyk_handle h, void **ptr, size_t *bytes);
/* perm[q] = physical bit position of logical qubit q (length n_qubits) */
int xyk_get_layout(xyk_handle h, int *perm);
int xyk_canonicalise(xyk_handle h);
int xyk_synchronize(xyk_handle h);
/* the CUDA stream work is enqueued on (cudaStream_t as void*) */
int xyk_stream(xyk_handle h, void **stream);

/* Introspection for benchmarks / tests. */
typedef struct {
    int64_t kernel_launches; /* kernels launched by this handle since creation             */
    int64_t pass_launches;   /* of which fused tile passes                                 */
    int64_t state_sweeps;    /* read+write sweeps over the state issued by apply_layers    */
    int32_t last_passes;     /* tile passes of the last apply_layers call                  */
    int32_t last_rounds;     /* register rounds of the last apply_layers call              */
    int32_t last_gates;      /* pair gates of the last apply_layers call                   */
    int32_t last_layers1;    /* first-order-layer equivalents of the last apply_layers     */
    float last_pass_ms;      /* CUDA-event time of the whole last apply_layers call (if timing) */
    int32_t timing_enabled;
    double pass_kernel_ms;   /* sum of per-launch CUDA-event durations of the fused tile pass kernel */
    int64_t pass_kernel_timed; /* number of launches in that sum (timing enabled only)            */
    double peer_pass_ms;     /* of which: passes whose store is the fused exchange (peer stores)     */
    int64_t peer_passes;
} xyk_stats;
int xyk_get_stats(xyk_handle h, xyk_stats *out);
int xyk_enable_timing(xyk_handle h, int on);

/* Schedule compiler, host only (no device needed): writes a JSON description of the pass
 * program for `reps` repetitions of `order` on an n-qubit all-to-all (or given K) problem.
 * Returns the number of bytes required (including NUL); fills buf if cap is large enough. */
int64_t xyk_plan_describe(int n_local, int dtype, const double *K, const double *omega, int n,
                          double delta, int order, int reps, char *buf, int64_t cap);

/* Schedule compiler, host only: the intermediate results of its first `upto_stage` stages (1: tile search, 2: + round
 * search, 3: + layouts) of a single-device program as JSON -- the unit-test view of the compiler.  Same return
 * convention as xyk_plan_describe. */
int64_t xyk_plan_stages(const double *K, const double *omega, int n, double delta, int order, int reps,
                        int upto_stage, char *buf, int64_t cap);

/* Batched parameter sweep: B independent n-qubit (n <= 12) problems evolved concurrently,
 * one CTA per problem with the whole state in shared memory.
 *   K: double[B*n*n], omega: double[B*n], step_sizes: double[m] (shared grid),
 *   R_out: double[B*(m+1)], final_states: NULL or B*2^n amplitudes (host, complex128).
 * Replaces: a Python loop of QuantumKuramotoSolver(n, K_b, omega_b).run(...) (config 3). */
int xyk_batch_run(int device, int n, int B, const double *K, const double *omega,
                  const double *step_sizes, int m, int reps, int order, double *R_out,
                  void *final_states);

#ifdef __cplusplus
}
#endif
#endif /* XYK_H_ */
