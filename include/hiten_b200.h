/* hiten_b200.h -- C ABI of the B200-native HITEN propagation hot path.
 *
 * Drop-in boundary: these entry points are what a binding of the reference
 * (iamgadmarconi/hiten v0.5.4, pure Python + Numba) calls in place of its @njit kernels.  Each
 * function cites the reference routine it replaces (paths relative to src/hiten/).
 *
 * Conventions
 *   - plain C, no exceptions, no torch types; every function returns an int status:
 *       0 = HB_OK, negative = usage error (HB_ERR_*), positive = cudaError_t of the failed CUDA call;
 *   - all array pointers are DEVICE pointers owned by the caller (the library never frees or
 *     reallocates them); `stream` is a cudaStream_t passed as void*; calls are asynchronous;
 *   - batch state arrays are struct-of-arrays ("SoA"): component c of trajectory i lives at
 *     a[c * n + i], so warps read and write coalesced;
 *   - per-trajectory results: status[i] (HB_TRAJ_*), n_acc[i] / n_rej[i] accepted / rejected
 *     attempted steps ("RK step" of the headline metric = one attempted step);
 *   - the library keeps no mutable global state: the work-queue cursor and hit counter live in a
 *     caller-provided workspace of hb_workspace_bytes() bytes, so concurrent calls on different
 *     streams with different workspaces are safe (the reference calls backends from a thread pool,
 *     algorithms/poincare/synodic/engine.py:135).
 *
 * Arithmetic variants (hb_integ.arith)
 *   HB_ARITH_PARITY  separately rounded mul/add/div/sqrt in the reference's operation order
 *                    (Numba fastmath=False, algorithms/utils/config.py:1): reproduces the reference's
 *                    step sequence; this is the variant parity is claimed for.
 *   HB_ARITH_FAST    same algorithm and controller, FMA-contracted and algebraically simplified RHS.
 */
#ifndef HITEN_B200_H
#define HITEN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_BADARG (-1)
#define HB_ERR_UNSUPPORTED (-2)
#define HB_ERR_NODEVICE (-3)

/* per-trajectory status */
#define HB_TRAJ_OK 0
#define HB_TRAJ_HIT 1        /* terminal event found                                  */
#define HB_TRAJ_MAXSTEPS 2   /* attempt cap reached before tf                         */
#define HB_TRAJ_NONFINITE 3  /* state or step size became NaN/Inf                     */
#define HB_TRAJ_RECORD_OVERFLOW 4 /* hb_cr3bp_section2: more accepted steps than the scratch holds; rerun with hb_cr3bp_section */

/* integrator ids = RungeKutta._map keys (algorithms/integrators/rk.py:2950) */
#define HB_RK4 4
#define HB_RK6 6
#define HB_RK8 8
#define HB_RK45 45
#define HB_DOP853 853

#define HB_SYMPLECTIC 2   /* Tao extended-phase-space integrator (_ExtendedSymplectic), order in hb_cm_opts */

#define HB_ARITH_PARITY 0
#define HB_ARITH_FAST 1

/* CR3BP system + _DirectedSystem wrapper (algorithms/dynamics/rtbp.py:31, base.py:186-310). */
typedef struct {
    double mu;       /* mass parameter                                                         */
    int32_t fwd;     /* +1 forward, -1 backward (derivatives negated, base.py:296-305)         */
    int32_t flip_lo; /* fwd == -1: negate dy[flip_lo:flip_hi]; flip_lo < 0 -> all components   */
    int32_t flip_hi;
    int32_t _pad;
} hb_cr3bp;

/* Integrator settings (AdaptiveRK(order, max_step, rtol, atol), algorithms/dynamics/base.py:446-450;
 * min_step default 10*eps, rk.py:833-834). */
typedef struct {
    int32_t method;       /* HB_DOP853, HB_RK45, HB_RK4/6/8                                   */
    int32_t arith;        /* HB_ARITH_*                                                       */
    double rtol, atol;
    double max_step, min_step;
    int64_t max_attempts; /* safety cap per trajectory (the reference has none); <=0 -> 2^31-1 */
    int32_t n_fixed_steps;/* fixed-step methods: number of equal steps over [t0, tf]             */
    int32_t max_ctas;     /* 0: the persistent DOP853 launches use every SM.  > 0: at most this many CTAs (one CTA owns an
                             SM's register file), so that two small batches launched on two streams run side by side --
                             each on its share of the SMs with twice as many trajectories per lane -- instead of one after
                             the other with two launch tails (no counterpart in the reference; results do not depend on it) */
    const int32_t *order; /* NULL: the persistent DOP853 launches hand out trajectories 0, 1, 2, ...  Else a DEVICE array
                             holding a permutation of 0..n-1 (the caller's responsibility): trajectory order[q] is the q-th to
                             be handed out.  Outputs stay in the caller's indexing, so results do not depend on it; it is a
                             scheduling hint -- longest expected trajectories first shortens the launch's tail (SURVEY 8e:
                             "sorting by expected cost"; step counts spread 60..236 inside one manifold tube) */
} hb_integ;

/* Plane event g(t,y) = y[idx] - offset (algorithms/poincare/singlehit/backend.py:30-67) with the
 * direction / tolerance fields of EventConfig / EventOptions (algorithms/types/configs.py:195,
 * options.py:280). */
typedef struct {
    int32_t idx;
    int32_t direction;  /* 0 any, >0 increasing, <0 decreasing */
    double offset;
    double xtol, gtol;
} hb_event;

/* Synodic section detector settings = _SynodicDetectionBackend defaults
 * (algorithms/poincare/synodic/backend.py:458-659, algorithms/types/services/maps.py:753-774). */
typedef struct {
    int32_t idx;             /* section_axis component                                          */
    int32_t direction;       /* None -> 0, +1, -1                                               */
    double offset;           /* section_offset                                                  */
    int32_t proj_i, proj_j;  /* plane_coords -> state indices (dedup distance is measured here) */
    int32_t segment_refine;  /* sub-intervals per sample segment minus one                      */
    int32_t max_hits_per_traj; /* <=0: unlimited                                                */
    double tol_on_surface;
    double dedup_time_tol, dedup_point_tol;
} hb_section;

/* One section hit (record of _SectionHit, algorithms/poincare/core/types.py:45). 72 bytes. */
typedef struct {
    int64_t traj;     /* trajectory index in the batch                      */
    int64_t seq;      /* ordinal of the hit inside its trajectory            */
    double t;         /* integrator time (>= 0); caller applies the fwd sign */
    double state[6];
} hb_hit;

/* Bytes of caller-provided device workspace every batch call needs (>= 256). */
int64_t hb_workspace_bytes(void);

/* Library / device probe: fills sm_count and compute capability; HB_ERR_NODEVICE without a GPU. */
int hb_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor);

/* Batched 6-state propagation to tf, end states only.
 * Replaces the serial loop  for fraction in ...: _propagate_dynsys(..., steps=2)
 * (algorithms/types/services/manifold.py:381-409 -> algorithms/dynamics/base.py:346-458 ->
 *  algorithms/integrators/rk.py:2377-2549 / 1269-1399 / 533-588).
 * yf uses the reference's API semantics: the dense interpolant evaluated at tf on the last
 * accepted segment (what states[-1] of _propagate_dynsys is).
 * tf_per_traj may be NULL (then tf is used for all).  For fixed-step methods (HB_RK4/6/8,
 * rk.py:533-588) n_fixed_steps (or, if 0, integ->n_fixed_steps) is the number of equal steps over
 * [t0, tf] (t_vals = linspace(t0, tf, n_fixed_steps + 1)); HB_RK45 is _RK45 (rk.py:1269-1399).   */
int hb_cr3bp_propagate(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *y0_soa,
                       double t0, double tf, const double *tf_per_traj, int32_t n_fixed_steps,
                       double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status,
                       void *workspace, void *stream);

/* Same propagation with dense output on a shared ascending grid t_eval[m] (device pointer):
 * states_out[i][k][c] (row-major [n][m][6], the reference's states array per trajectory).
 * Replaces _integrate_dop853's dense-output pass (rk.py:2498-2543).                             */
int hb_cr3bp_dense(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *y0_soa,
                   const double *t_eval, int32_t m, double *states_out, int32_t *n_acc,
                   int32_t *n_rej, int32_t *status, void *workspace, void *stream);

/* Fused tube + section: propagates like hb_cr3bp_dense but, instead of storing the m samples, streams
 * them through the synodic detector (same semantics as hb_synodic_detect on the stored tube: hits are
 * the reference's linear interpolants between grid samples, times carry the sign of sys->fwd).  This is
 * Manifold.compute() followed by SynodicMap.compute() (system/manifold.py:226, system/maps/synodic.py:121)
 * without the 226 KB per trajectory of dense output -- the form BASELINE config 5 (1e6 trajectories) needs.
 * yf = the last grid sample (what states[-1] is).  Read the hit count with hb_read_hit_count.       */
int hb_cr3bp_section(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                     const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits, int64_t hit_capacity,
                     int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status,
                     void *workspace, void *stream);

/* The same hits as hb_cr3bp_section (bit for bit) computed as a PIPELINE of small kernels with a compact
 * intermediate (hb_section_scan.cu): the propagation kernel stores what the dense output of every accepted step
 * depends on (512 B per step) in `scratch`; data-parallel kernels then build the event component of each step's
 * interpolant, scan the grid samples, refine the segments that can hold a hit and order + de-duplicate the hits.
 * About 2x faster than the fused kernel (which is instruction-fetch bound) whenever the scratch fits:
 * hb_section2_scratch_bytes(n, steps_capacity) bytes for at most steps_capacity accepted steps per trajectory
 * (512 B per step + 4.1 KB per trajectory).  Trajectories that need more steps, or have more than 32 candidate
 * hits, get status = HB_TRAJ_RECORD_OVERFLOW and NO hits; their number is read with hb_read_record_overflow --
 * rerun those with hb_cr3bp_section.                                                                 */
#define HB_RECORDS_ALL 0
#define HB_RECORDS_NEAR_SECTION 1
int64_t hb_section2_scratch_bytes(int64_t n, int32_t steps_capacity);
int hb_cr3bp_section2(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                      const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits, int64_t hit_capacity,
                      int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status,
                      void *scratch, int64_t scratch_bytes, void *workspace, void *stream,
                      void *const *stage_events /* NULL, or 5 caller-owned cudaEvent_t recorded on `stream` before /
                      between / after the stages propagate+record, step scan, candidate emission, order+dedup
                      (measurement aid of bench.py; entries may be NULL; the library keeps no state) */,
                      int32_t records /* HB_RECORDS_ALL: every accepted step is recorded (steps_capacity = accepted steps
                      per trajectory; what hb_section2_filter needs).  HB_RECORDS_NEAR_SECTION: the propagation kernel
                      screens every accepted step (event component of its dense interpolant in fast arithmetic, quiet-step
                      bound with a widened margin) and records only the steps that can come near the section plane plus
                      their two neighbours -- about a tenth of the steps of a tube; steps_capacity then counts RECORDED
                      steps per trajectory.  Same hits bit for bit (every step left out is provably free of hits). */);

/* The same step -- same hits, counts and end states, bit for bit -- with the step records handed from the propagating
 * warps to scanning warps THROUGH SHARED MEMORY inside one persistent kernel (hb_section_stream.cu): nothing but
 * initial conditions, end states, the few step records that hold a noted segment and the hits touch HBM.  No per-
 * trajectory step capacity and no 512 B-per-step scratch: `scratch` holds the candidate / segment lists (4.2 KB per
 * trajectory) and a pool of `pool_records_per_traj` step records per trajectory on average (512 B each; a tube needs
 * ~3, ask for 8).  Trajectories with more than 32 noted segments or candidates, or that found the pool exhausted, get
 * status = HB_TRAJ_RECORD_OVERFLOW and NO hits (count: hb_read_record_overflow) -- rerun those with hb_cr3bp_section.
 * Trajectories that end in HB_TRAJ_MAXSTEPS / HB_TRAJ_NONFINITE report no hits.
 * Replaces the same reference routines as hb_cr3bp_section: services/manifold.py:381-440 (tube) +
 * poincare/synodic/backend.py:458-659, 382-455 (section).
 * stage_events: NULL, or 4 caller-owned cudaEvent_t recorded before the propagate+scan kernel, after it, after the
 * candidate emission and after order+dedup.                                                          */
int64_t hb_section3_scratch_bytes(int64_t n, int32_t pool_records_per_traj);
int hb_cr3bp_section3(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                      const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits, int64_t hit_capacity,
                      int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status,
                      void *scratch, int64_t scratch_bytes, void *workspace, void *stream,
                      void *const *stage_events);

/* Propagation with a terminal plane event (event always terminal, as in the reference):
 * replaces _integrate_dop853_until_event + _dop853_refine_in_step (rk.py:2680-2803, 2006-2102).
 * On a hit status[i] = HB_TRAJ_HIT, t_hit[i] / y_hit_soa hold the refined crossing; otherwise
 * t_hit[i] = tmax reached and y_hit = final state.                                              */
int hb_cr3bp_event(const hb_cr3bp *sys, const hb_integ *integ, const hb_event *ev, int64_t n,
                   const double *y0_soa, double t0, double tmax, const double *tmax_per_traj,
                   double *t_hit, double *y_hit_soa, int32_t *n_acc, int32_t *n_rej,
                   int32_t *status, void *workspace, void *stream);

/* Measured FP64 FMA throughput of this device: runs a register-resident DFMA chain for about
 * `millis` ms and returns flop/s (2 flop per FMA).  Used as the roofline denominator.           */
int hb_dfma_peak(double millis, double *flops_per_s, void *stream);

/* Batched 42-dimensional state + STM propagation: replaces _compute_stm / _var_equations /
 * _jacobian_crtbp (algorithms/dynamics/rtbp.py:258-340, 168-255, 77-165) as called by the manifold
 * service (algorithms/types/services/manifold.py:236-243), the corrector
 * (algorithms/corrector/operators.py:298-307) and orbit stability (services/orbits.py).
 * The initial condition is PHI0 = [I6 row-major, x0]; sys->flip_lo/hi select the derivative block the
 * backward wrapper negates: (36,42) = state only (what _compute_stm passes), <0 = everything.
 * phi_out[i][0..41] is the reference's flat PHI row at tf (Phi row-major in [0,36), state in [36,42)),
 * with the same dense-interpolant-at-tf semantics as hb_cr3bp_propagate.
 * integ->method: HB_DOP853 (the shipped default, the tuned kernel), HB_RK45 (_compute_stm(..., order=5): rk.py:1138-1399)
 * or HB_RK4 / HB_RK6 / HB_RK8 (method="fixed": rk.py:422-588; integ->n_fixed_steps steps over linspace(t0, tf), or one
 * step per interval of t_eval in the dense call).                                                  */
int hb_cr3bp_stm(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *x0_soa, double t0,
                 double tf, const double *tf_per_traj, double *phi_out, int32_t *n_acc, int32_t *n_rej,
                 int32_t *status, void *workspace, void *stream);

/* The same with dense output PHI[i][k][0..41] on an ascending grid: t_eval is one shared array of m
 * times, or (t_eval_per_traj != 0) one row of m times per trajectory (each orbit its own period). */
int hb_cr3bp_stm_dense(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *x0_soa,
                       const double *t_eval, int32_t m, int32_t t_eval_per_traj, double *phi_dense,
                       int32_t *n_acc, int32_t *n_rej, int32_t *status, void *workspace, void *stream);

/* ---- centre-manifold Poincare map (polynomial Hamiltonian) ---------------------------------------
 * Sparse real term table of the six partials dH/d(q1,q2,q3,p1,p2,p3): the reference's jac_H
 * (lists of packed complex128 coefficient arrays + clmo exponent tables, algorithms/polynomial/base.py:
 * 99-259) reduced to its non-zero terms in evaluation order.  `terms` is a DEVICE array of 16-byte
 * records {double coef; uint64 ex} with the exponents of (q1,q2,q3,p1,p2,p3) in bytes 0..5 of `ex` and
 * the homogeneous degree in bits 48..63 (summation is grouped by degree like _polynomial_evaluate). */
typedef struct {
    int32_t n_dof;      /* 3 */
    int32_t max_deg;    /* largest exponent in the table */
    int64_t ptr[7];     /* terms of partial p are [ptr[p], ptr[p+1]) */
    const void *terms;
} hb_polyham;

#define HB_MAX_TAO_SUBSTEPS 27
/* Options of _poincare_map (algorithms/poincare/centermanifold/backend.py:314-382). */
typedef struct {
    double dt;
    int32_t max_steps;
    int32_t method;   /* HB_RK4 / HB_RK6 / HB_RK8 ("fixed") or HB_SYMPLECTIC                       */
    int32_t order;    /* symplectic order 2, 4, 6, 8                                                */
    int32_t section;  /* section coordinate: 0 q2, 1 p2, 2 q3, 3 p3                                 */
    int32_t arith;    /* HB_ARITH_*                                                                 */
    int32_t n_sub;    /* filled by hb_cm_prepare: order-2 Tao kernels per dt and their parameters   */
    double sub_ts[HB_MAX_TAO_SUBSTEPS], sub_cos[HB_MAX_TAO_SUBSTEPS], sub_sin[HB_MAX_TAO_SUBSTEPS];
} hb_cm_opts;

/* Host-only: flattens the triple-jump recursion (algorithms/integrators/symplectic.py:509-560) into
 * sub_ts[] and evaluates omega = (c_omega*dt)^-order and cos/sin(2*omega*ts) with the host libm, exactly
 * as the reference does per step.  Needs no GPU.                                                  */
int hb_cm_prepare(hb_cm_opts *opts, double c_omega);

/* _poincare_map / _poincare_step / _detect_crossing: seeds[n][4] = (q2,p2,q3,p3) -> flags[n] (1 = section
 * crossing found within max_steps), out[n][4] the Hermite-interpolated crossing state, t_out[n].     */
int hb_cm_poincare_map(const hb_polyham *ham, const hb_cm_opts *opts, int64_t n, const double *seeds,
                       int32_t *flags, double *out, double *t_out, void *workspace, void *stream);

/* The same map with a kernel SPECIALISED at run time for this Hamiltonian (hb_cm_jit.cu): the gradient is
 * generated as straight-line code (powers in registers, coefficients as immediates, the reference's term order),
 * compiled for sm_100a with NVRTC and cached.  Same arguments, same results (bit-identical in the parity
 * variant); ~50x the throughput of the table-driven kernel.  Needs libnvrtc.so.12 and the CUDA driver at run time
 * (loaded lazily); returns HB_ERR_UNSUPPORTED if NVRTC is not available.                                  */
int hb_cm_poincare_map_jit(const hb_polyham *ham, const hb_cm_opts *opts, int64_t n, const double *seeds,
                           int32_t *flags, double *out, double *t_out, void *workspace, void *stream);

/* Host-only: generate + compile the specialised kernel for a term table in HOST memory (no GPU needed);
 * *cubin_bytes receives the cubin size, source_out (optional, source_cap bytes) the generated CUDA source or,
 * on failure, the compiler log.                                                                       */
int hb_cm_jit_compile_host(const void *terms_host, const int64_t *ptr, int32_t max_deg, const hb_cm_opts *opts,
                           int64_t *cubin_bytes, char *source_out, int64_t source_cap);

/* ---- _ExtendedSymplectic.integrate (algorithms/integrators/symplectic.py:877-1004; reached from
 * _propagate_dynsys(method="symplectic"), algorithms/dynamics/base.py:436-444): the Tao integrator of a polynomial
 * Hamiltonian system over a TIME GRID, for a batch of initial states that share the grid. ------------------------- */
typedef struct {
    int32_t order;   /* 2, 4, 6, 8                                                          */
    int32_t arith;   /* HB_ARITH_*                                                          */
    int32_t m;       /* grid points (>= 2, integrators/base.py:183)                         */
    int32_t n_sub;   /* order-2 Tao kernels per grid interval: from hb_tao_grid_prepare     */
} hb_symp_opts;

/* Host-only, needs no GPU.  t_vals_signed[m] is the grid the low-level routine sees (t_vals * fwd, symplectic.py:963).
 * For every grid interval: dt = t[i+1] - t[i], omega = (c_omega*dt)^-order (_get_tao_omega :38-60), the triple-jump
 * schedule of _recursive_update_poly (:543-560) and cos / sin(2*omega*ts) with the host libm, exactly as the reference
 * evaluates them per step.  tab[(m-1)][3][n_sub] = {ts[], cos[], sin[]}; tab == NULL only returns *n_sub.      */
int hb_tao_grid_prepare(const double *t_vals_signed, int32_t m, int32_t order, double c_omega, int32_t *n_sub,
                        double *tab, int64_t tab_capacity);

/* _integrate_symplectic (symplectic.py:564-653): y0[n][6] = [Q, P] -> traj[n][m][6] (row 0 = y0); the extended state
 * [Q,P,X,Y] is carried across the grid.  tao_tab is the DEVICE copy of hb_tao_grid_prepare's table.  Bit-identical to the
 * reference in the parity variant.                                                                              */
int hb_ham_symplectic_dense(const hb_polyham *ham, const hb_symp_opts *opts, int64_t n, const double *y0,
                            const double *tao_tab, double *traj, void *workspace, void *stream);

/* _integrate_symplectic_until_event (symplectic.py:657-782) with the plane event g = y[idx] - offset, refined by
 * bisection on the step's cubic Hermite interpolant (_hermite_refine_event_symplectic :282-367).  hit[i] = 1: t_hit[i]
 * (on the signed grid) / y_hit[i][6] are the refined event, n_rows[i] the trajectory rows before it; hit[i] = 0:
 * t_hit = t_vals[m-1], y_hit = last state, n_rows = m.  traj[n][m][6] may be NULL (rows past the event are untouched). */
int hb_ham_symplectic_event(const hb_polyham *ham, const hb_symp_opts *opts, const hb_event *ev, int64_t n,
                            const double *y0, const double *t_vals_signed, const double *tao_tab, double *traj,
                            int32_t *hit, double *t_hit, double *y_hit, int32_t *n_rows, void *workspace, void *stream);

/* The same two calls with a kernel SPECIALISED at run time for this Hamiltonian (hb_cm_jit.cu: the generated straight-line
 * gradient of hb_cm_poincare_map_jit inside the grid loop; NVRTC, cached per Hamiltonian and arithmetic).  ev == NULL:
 * hb_ham_symplectic_dense (hit / t_hit / y_hit / n_rows unused); otherwise hb_ham_symplectic_event.  Bit-identical to the
 * table-driven kernels in the parity variant; returns HB_ERR_UNSUPPORTED if NVRTC is not available.              */
int hb_ham_symplectic_jit(const hb_polyham *ham, const hb_symp_opts *opts, const hb_event *ev, int64_t n,
                          const double *y0, const double *t_vals_signed, const double *tao_tab, double *traj,
                          int32_t *hit, double *t_hit, double *y_hit, int32_t *n_rows, void *workspace, void *stream);
/* Host-only: generate + compile that kernel for a term table in HOST memory (no GPU needed); *cubin_bytes = cubin size. */
int hb_symp_jit_compile_host(const void *terms_host, const int64_t *ptr, int32_t arith, int64_t *cubin_bytes);

/* ---- _FixedStepRK.integrate on a polynomial Hamiltonian system: the `_ham` kernels of the RK classes
 * (algorithms/integrators/rk.py: _integrate_fixed_rk_ham :592-656, rk_embedded_step_ham_jit_kernel :216-270,
 * _integrate_fixed_rk_until_event_ham :722-757 + _hermite_refine_in_step :331-391).  method = HB_RK4 / HB_RK6 / HB_RK8;
 * t_vals[m] is a DEVICE array (one step per grid interval, h = t[i+1] - t[i]); the reference's `_ham` kernels take
 * system.rhs_params and never see a direction wrapper, so there is no direction argument.
 * dense: y0[n][6] -> traj[n][m][6] and (optional) derivs[n][m][6] = _Solution.derivatives.
 * event: as hb_ham_symplectic_event (hit / t_hit / y_hit / n_rows, traj optional).                              */
int hb_ham_rk_dense(const hb_polyham *ham, int32_t method, int32_t arith, int64_t n, const double *y0,
                    const double *t_vals, int32_t m, double *traj, double *derivs, void *workspace, void *stream);
int hb_ham_rk_event(const hb_polyham *ham, int32_t method, int32_t arith, const hb_event *ev, int64_t n,
                    const double *y0, const double *t_vals, int32_t m, double *traj, int32_t *hit, double *t_hit,
                    double *y_hit, int32_t *n_rows, void *workspace, void *stream);

/* ---- AdaptiveRK on a polynomial Hamiltonian system: the `_ham` kernels of _DOP853 / _RK45
 * (algorithms/integrators/rk.py: _integrate_dop853_ham :2553-2676 with _dop853_build_dense_cache_ham :1878,
 * _integrate_dop853_until_event_ham :2807-2868 with _dop853_refine_in_step_ham :2106, _integrate_rk45_ham :1403-1456,
 * _integrate_rk45_until_event_ham :1589-1633).  integ->method = HB_DOP853 or HB_RK45; same controller and first-step
 * rule as the CR3BP kernels; no direction argument (see hb_ham_rk_dense).
 * dense: y0[n][6], t_eval[m] (DEVICE, ascending) -> states[n][m][6] and (optional) derivs[n][m][6] = the vector field
 *        re-evaluated at every output sample (_Solution.derivatives); n_acc / n_rej / status per trajectory.
 * event: status[i] = HB_TRAJ_HIT with the refined t_hit[i] / y_hit[i][6], else t_hit = tmax reached, y_hit = last state. */
int hb_ham_adaptive_dense(const hb_polyham *ham, const hb_integ *integ, int64_t n, const double *y0,
                          const double *t_eval, int32_t m, double *states, double *derivs, int32_t *n_acc,
                          int32_t *n_rej, int32_t *status, void *workspace, void *stream);
int hb_ham_adaptive_event(const hb_polyham *ham, const hb_integ *integ, const hb_event *ev, int64_t n,
                          const double *y0, double t0, double tmax, double *t_hit, double *y_hit, int32_t *n_acc,
                          int32_t *n_rej, int32_t *status, void *workspace, void *stream);

/* Seed lifting for the centre-manifold map: replaces the per-seed Python Brent solves of
 * _CenterManifoldInterface.lift_plane_point / solve_missing_coord (algorithms/poincare/centermanifold/
 * interfaces.py:297-337, 212-268; solve_bracketed_brent algorithms/utils/rootfinding.py:92-190) that the seeding
 * strategies (centermanifold/strategies.py:67-520) call once per candidate.  `H` holds the Hamiltonian ITSELF
 * (hamsys.poly_H()) as polynomial 0 of a term table in hb_polyham format.  plane_pts[n][2] are points in the
 * section's plane coordinates -- (q2,p2) for sections q3/p3, (q3,p3) for sections q2/p2 --; states[n][4] receives
 * (q2,p2,q3,p3) with the section coordinate 0 and the missing coordinate solved from H = h0; ok[n] is 0 where the
 * reference returns None (no bracket, or Brent did not converge).  Bit-identical to the reference.        */
typedef struct {
    double h0;             /* energy level                                                  */
    double initial_guess;  /* 1e-3                                                          */
    double expand_factor;  /* 2.0                                                           */
    double xtol;           /* 1e-12                                                         */
    int32_t max_expand;    /* 40                                                            */
    int32_t symmetric;     /* also try the negative bracket                                 */
    int32_t section;       /* 0 q2, 1 p2, 2 q3, 3 p3                                        */
    int32_t max_iter;      /* Brent iterations, 200                                         */
} hb_cm_lift_opts;
int hb_cm_lift(const hb_polyham *H, const hb_cm_lift_opts *opts, int64_t n, const double *plane_pts,
               double *states, int32_t *ok, void *stream);

/* Connection search between two sets of section hits: replaces _ConnectionsBackend.run
 * (algorithms/connections/backends.py:425-540): radius pairing on the 2-D section plane (_radius_pairs_2d :100-171),
 * mutual-nearest filter (:468-489), segment refinement with each member's nearest same-set neighbour
 * (_nearest_neighbor_2d :174-233, _refine_pairs_on_section :323-423), Delta-V = |v_u - v_s| and the dv_tol /
 * bal_tol classification (:507-533).  pu[n_u][2] / ps[n_s][2] are the section points of the unstable (source) and
 * stable (target) manifolds, Xu / Xs their 6-states; all DEVICE arrays.  Accepted connections are appended to
 * out[0 .. capacity) in arbitrary order -- the reference's order is ascending (delta_v, index_u) --; n_out,
 * n_dropped (no room) and pairs_considered (the reference's metadata) are HOST outputs; the call synchronises
 * `stream`.  Indices, Delta-V, points and states are bit-identical to the reference.  Points must be finite.  */
typedef struct {
    int64_t index_u, index_s;   /* hit indices in pu / ps                                   */
    double delta_v;
    double point2d[2];          /* refined common point on the section plane                */
    double state_u[6], state_s[6];
    int64_t kind;               /* 0 ballistic (delta_v <= bal_tol), 1 impulsive            */
} hb_connection;
int64_t hb_connections_scratch_bytes(int64_t n_u, int64_t n_s);
int hb_connections(const double *pu, int64_t n_u, const double *ps, int64_t n_s, const double *Xu, const double *Xs,
                   double eps, double dv_tol, double bal_tol, hb_connection *out, int64_t capacity, int64_t *n_out,
                   int64_t *n_dropped, int64_t *pairs_considered, void *scratch, int64_t scratch_bytes, void *stream);

/* Manifold-tube initial conditions on the device: replaces the per-fraction host work of
 * _ManifoldDynamicsService._compute_manifold_section + _totime (algorithms/types/services/manifold.py:470-573) for a
 * whole tube.  phi_dense[n_samples][42] is the reference's PHI array of the orbit (row k = [Phi row-major, x] at
 * tt[k]), exactly what hb_cr3bp_stm_dense writes; tt[n_samples] its (signed) sample times; eigvec[6] the real part of
 * the chosen eigenvector; direction = +1 / -1 (manifold branch).  For every fraction f the sample index is the first
 * minimum of |f*period - |tt[k]||; for every (fraction k, displacement j) the initial condition
 *     x0W = x_k + (displacement_j / |MAN[0:3]|) * MAN,   MAN = direction * Phi_k @ eigvec,  tiny z / vz zeroed
 * goes to trajectory i = j*n_fractions + k of the SoA array x0w_soa[6][n_fractions*n_displacements] that the
 * propagation entry points read.  node_idx[n_fractions] (device, required) receives the sample indices.
 * All pointers are device pointers.  Bit-identical to the reference (incl. its BLAS accumulation order).      */
int hb_manifold_ics(const double *phi_dense, const double *tt, int32_t n_samples, double period,
                    const double *eigvec, int32_t direction, const double *fractions, int64_t n_fractions,
                    const double *displacements, int64_t n_displacements, double *x0w_soa, int32_t *node_idx,
                    void *stream);

/* Trajectory filters of Manifold.compute() on stored tubes: replaces the numpy safe-radius test of
 * _run_compute (algorithms/types/services/manifold.py:412-424) and _max_rel_energy_error
 * (algorithms/common/energy.py:27-76).  states is hb_cr3bp_dense's output [n][m][6]; out[i] = {min_k r1, min_k r2,
 * max_k |C_k - C_0| / |C_0|} (absolute drift when |C_0| <= 1e-14), NaN samples propagate into the minima like
 * np.min; keep[i] (optional) = 1 unless min r1 < safe_r1, min r2 < safe_r2 or the drift exceeds energy_tol.     */
typedef struct {
    double mu;
    double safe_r1, safe_r2;   /* safe_distance * body radius / (distance * 1e3), manifold.py:341-345 */
    double energy_tol;
} hb_tube_filter_opts;
int hb_tube_filter(const hb_tube_filter_opts *opts, int64_t n, const double *states, int32_t m, double *out,
                   int32_t *keep, void *stream);

/* Batched single-shooting differential correction of periodic orbits: n independent Newton problems
 * advanced in lock-step.  Replaces, per orbit, _NewtonBackend.run (algorithms/corrector/backends/newton.py:20-150)
 * with the residual / Jacobian of _SingleShootingOrbitOperators (algorithms/corrector/operators.py:319-452): the
 * residual is the state at the first crossing of the plane y[event_idx] = event_offset found by
 * _SingleHitBackend._cross (algorithms/poincare/singlehit/backend.py:164-282: window [0, pi] after a 1e-12 alignment
 * step, fallback window [pi/2 - 0.15, + pi]) restricted to res[] minus target[]; the Jacobian is the res x ctrl block
 * of the STM at the event time (minus _halo_quadratic_term, algorithms/types/services/orbits.py:917-946, when
 * halo_quadratic != 0) or central differences (finite_difference != 0, backends/base.py); the update is
 * _solve_delta_dense (cond > 1e8 -> ridge 1e-12) followed by the inf-norm cap max_delta and either the Armijo
 * back-tracking search (algorithms/corrector/stepping/armijo.py:60-170) or the plain step (stepping/plain.py).
 * Two controls and two residuals (every shipped family: halo, Lyapunov, vertical).  sys->fwd must be +1 and
 * integ->method HB_DOP853 (the shipped configs); integ supplies rtol / atol / arithmetic variant.
 * x0_soa / xc_soa are [6][n] SoA device arrays (initial guesses in, corrected initial states out); half_period[n] is
 * the event time of the corrected state (NaN unless converged); iterations[n], residual_norm[n] (inf-norm) and
 * status[n] (HB_CORR_*) mirror CorrectorOutput / the reference's exceptions.  rk_steps6 / rk_steps42 (optional HOST
 * outputs) receive the attempted 6-state and 42-state DOP853 steps of the whole call.  scratch: device block of
 * hb_correct_scratch_bytes(n) bytes.  The call synchronises `stream` (it reads the number of orbits still active
 * between launches -- 4 bytes -- and nothing else).
 * Event propagation is the bit-exact path and the 2x2 solve reproduces np.linalg.solve bit for bit, so the
 * finite-difference variant is bit-identical to the reference; the STM is tolerance-level (see hb_cr3bp_stm), so with
 * the analytic Jacobian corrected states agree with the reference at Newton-convergence level (<= 1e-10) and iteration counts to within
 * one (the reference's |R| < 1e-12 test sits on the 1e-12 noise floor of its own event solver).                 */
#define HB_CORR_CONVERGED 0
#define HB_CORR_MAX_ATTEMPTS 1   /* ConvergenceError: not converged after max_attempts                      */
#define HB_CORR_STEP_FAILED 2    /* ConvergenceError: step strategy found no productive step                */
#define HB_CORR_NO_EVENT 3       /* no plane crossing for the current iterate (the reference raises)        */
#define HB_CORR_SINGULAR 4       /* singular Jacobian / failed STM propagation                              */
typedef struct {
    int32_t ctrl[2], res[2];     /* OrbitCorrectionConfig.control_indices / residual_indices               */
    double target[2];
    int32_t event_idx;           /* _plane_crossing_factory coordinate                                     */
    int32_t halo_quadratic;      /* extra_jacobian = _halo_quadratic_term                                  */
    double event_offset;
    int32_t finite_difference;   /* NumericalConfig.finite_difference                                      */
    int32_t line_search;         /* NumericalConfig.line_search_enabled                                    */
    double tol, max_delta;       /* ConvergenceOptions (max_delta = inf: no cap)                           */
    double fd_step;              /* NumericalOptions                                                       */
    double alpha_reduction, min_alpha, armijo_c;
    int32_t max_attempts;
    int32_t _pad;
} hb_correct_opts;
int64_t hb_correct_scratch_bytes(int64_t n);
int hb_correct_orbits(const hb_cr3bp *sys, const hb_integ *integ, const hb_correct_opts *opts, int64_t n,
                      const double *x0_soa, double *xc_soa, double *half_period, int32_t *iterations,
                      double *residual_norm, int32_t *status, int64_t *rk_steps6, int64_t *rk_steps42,
                      void *scratch, int64_t scratch_bytes, void *workspace, void *stream);

/* The same filter quantities for the trajectories of the LAST hb_cr3bp_section2 call that used `scratch`, computed from
 * the step records it left there -- the section pipeline never stores the tube ("removing the need for
 * dense output when the caller only wants sections").  Every one of the m grid samples is rebuilt from its step's
 * interpolant and judged with the same arithmetic as hb_tube_filter: out / keep are bit-identical to hb_cr3bp_dense +
 * hb_tube_filter.  Pass the same n, t_eval, m, n_acc, status, scratch and scratch_bytes as to hb_cr3bp_section2 (and the
 * same sys / integ).  keep[i] = -1 (out = NaN) for trajectories without a complete record set (status != HB_TRAJ_OK,
 * e.g. HB_TRAJ_RECORD_OVERFLOW): judge those on a stored tube.                                                   */
int hb_section2_filter(const hb_cr3bp *sys, const hb_integ *integ, const hb_tube_filter_opts *opts, int64_t n,
                       const double *t_eval, int32_t m, const int32_t *n_acc, const int32_t *status,
                       const void *scratch, int64_t scratch_bytes, double *out, int32_t *keep, void *stream);

/* Synodic-section crossing detection on precomputed trajectories (linear branch, the one the
 * reference's defaults select): replaces _SynodicDetectionBackend.run / detect_on_trajectory /
 * _detect_with_segment_refine / _order_and_dedup_hits (algorithms/poincare/synodic/backend.py:
 * 823-887, 687-821, 458-659, 382-455).
 *   states  : samples of all trajectories concatenated, row-major [sum m_i][6] (the reference's
 *             per-trajectory `states` arrays back to back);
 *   times   : signed sample times, concatenated like states, or ONE shared array of m_uniform
 *             entries when times_shared != 0 (a manifold tube: forward * linspace(0, tf, steps));
 *   offsets : [n_traj + 1] sample offsets, or NULL when every trajectory has m_uniform samples.
 * Hits are appended to hits[0 .. hit_capacity) in arbitrary order ((traj, seq) gives the reference
 * order); hits_per_traj (optional, [n_traj]) receives the per-trajectory count.  The number of hits
 * and of hits dropped for lack of capacity is read back with hb_read_hit_count.                   */
int hb_synodic_detect(const hb_section *sec, int64_t n_traj, const double *states, const double *times,
                      const int64_t *offsets, int32_t m_uniform, int32_t times_shared, hb_hit *hits,
                      int64_t hit_capacity, int32_t *hits_per_traj, void *workspace, void *stream);

/* The same call for a request with interp_kind == "cubic" (backend.py:762): cubic Hermite g through the neighbouring
 * samples on the segment_refine + 1 sub-intervals, `newton_max_iter` Newton steps on the cubic clamped to the
 * sub-interval, cubic Hermite hit state (_detect_with_segment_refine :541-645; segment_refine == 0: _refine_hits_cubic
 * :274-379; poincare/utils.py _hermite_scalar / _hermite_der :54-148).  Where the reference's own `dt > 0.0` guards
 * fail (decreasing times: a backward tube) it computes the linear formulas, and so does this.  Bit-identical hits. */
int hb_synodic_detect_cubic(const hb_section *sec, int32_t newton_max_iter, int64_t n_traj, const double *states,
                            const double *times, const int64_t *offsets, int32_t m_uniform, int32_t times_shared,
                            hb_hit *hits, int64_t hit_capacity, int32_t *hits_per_traj, void *workspace, void *stream);

/* Hit / overflow counters of the last call that used `workspace` (synchronises `stream`). */
int hb_read_hit_count(const void *workspace, int64_t *n_hits, int64_t *n_overflow, void *stream);

/* Sharded runs (trajectories shard by index over the GPUs, one exchange at the end -- hit records and end
 * states to the gathering rank; the seam in the reference is its worker pool, algorithms/poincare/synodic/engine.py:92-139):
 * writes this shard's result into its slot of the gathering rank's receive buffer over NVLink peer memory, stream-ordered
 * behind the pipeline that produced it and WITHOUT a host round trip -- the hit count is read from `workspace` on the
 * device.  peer_slots[r] (r < world <= 16) = device pointer, mapped into this process, of this shard's slot in rank r's
 * buffer; slot layout in doubles: [8 header | 9 * hit_slots hit records | 6 * n_local end states (SoA)], 16-byte aligned,
 * hit_slots even.  Every rank's slot gets the header {hits, n_local, dropped hits, trajectories with
 * HB_TRAJ_RECORD_OVERFLOW, sendable (1 / 0), 0, 0, 0}; rank dst's also the payload -- unless hits were dropped, step
 * records overflowed or the hits exceed hit_slots (sendable = 0: the caller completes the shard and sends it with
 * host-sized copies; after the closing barrier all ranks see that flag).  For small shards (the copy-engine path of
 * hiten_b200/sharded.py overlaps better when a later persistent launch owns the SMs).                */
int hb_peer_put(void *const *peer_slots, int32_t world, int32_t dst, const hb_hit *hits, int64_t hit_slots,
                const double *yf_soa, int64_t n_local, const void *workspace, void *stream);

/* Number of trajectories of the last hb_cr3bp_section2 call that got HB_TRAJ_RECORD_OVERFLOW (no hits were
 * reported for them; rerun those with hb_cr3bp_section).  Synchronises `stream`.                   */
int hb_read_record_overflow(const void *workspace, int64_t *n_traj, void *stream);

/* Arithmetic self-test: for every pair (a[i], b[i]) evaluates the shared-reciprocal division and the
 * restated sqrt fast path used by the parity variant next to the compiler's div.rn / sqrt.rn, and
 * the controller's pow(b, a).  All pointers are device arrays of n doubles.                      */
int hb_selftest_arith(const double *a, const double *b, int64_t n, double *div_shared, double *div_ref,
                      double *sqrt_fast, double *sqrt_ref, double *pow_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HITEN_B200_H */
