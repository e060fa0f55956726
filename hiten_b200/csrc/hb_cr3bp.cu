// hb_cr3bp.cu -- batched 6-state CR3BP propagation with DOP853 on sm_100a.
//
// One trajectory per thread, all stage vectors in registers (13 x 6 doubles), tableau entries
// folded into the instruction stream at compile time (templates over hb_coeffs.h), per-thread
// adaptive step control, and a persistent-thread work queue: a lane whose trajectory reached tf
// immediately pulls the next index from an atomic cursor, so uneven step counts (2x inside one
// manifold tube) do not idle lanes.
//
// Reference routines restated here (paths relative to hiten/ in iamgadmarconi/hiten v0.5.4):
//   _crtbp_accel                      algorithms/dynamics/rtbp.py:31-74
//   _DirectedSystem wrapper           algorithms/dynamics/base.py:296-305
//   dop853_step_jit_kernel            algorithms/integrators/rk.py:1637-1708
//   _integrate_dop853 (+ dense)       algorithms/integrators/rk.py:2377-2549
//   _integrate_dop853_until_event     algorithms/integrators/rk.py:2680-2803
//   _dop853_build_dense_cache / _dop853_eval_dense / _dop853_refine_in_step   rk.py:1791-2102
//   controller helpers                algorithms/integrators/utils.py
#include "hb_cr3bp_common.cuh"

namespace {
using namespace hbc;

enum { MODE_FINAL = 0, MODE_DENSE = 1, MODE_EVENT = 2, MODE_RECORD = 3, MODE_RECORD_NEAR = 4 };
constexpr bool is_record(int mode) { return mode == MODE_RECORD || mode == MODE_RECORD_NEAR; }

// One accepted step's record in the thread's own (padded) shared-memory row, then ONE 512-byte bulk store.
HB_DEV void assemble_record(double *rec_row, double t, double t_new, const double (&y)[6], const double (&yh)[6],
                            const double (&k)[13][6], int step, bool last)
{
    double2 *q = (double2 *)rec_row;
    q[0] = make_double2(t, t_new);
#pragma unroll
    for (int d = 0; d < 6; d += 2) {
        q[1 + d / 2] = make_double2(y[d], y[d + 1]);
        q[4 + d / 2] = make_double2(yh[d], yh[d + 1]);
    }
#pragma unroll
    for (int j = 5; j < 13; ++j)
#pragma unroll
        for (int d = 0; d < 6; d += 2) q[7 + 3 * (j - 5) + d / 2] = make_double2(k[j][d], k[j][d + 1]);
    *(int2 *)(rec_row + HB_REC_META) = make_int2(step, last ? HB_REC_LAST : 0);
    rec_row[HB_REC_META + 1] = 0.0;
}
HB_DEV void store_record(const double *rec_row, double *dst)
{
#ifndef HB_REC_NOCOPY
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((unsigned)__cvta_generic_to_shared(rec_row)), "r"(HB_REC_DOUBLES * 8) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// The persistent-thread kernel.
// ---------------------------------------------------------------------------------------------
// CSEC (MODE_RECORD_NEAR only): the section component the screening evaluates, resolved at compile time -- a switch on
// the section index inside the step loop cost 1 % (an indirect branch into code far from the loop, every accepted step).
template <class AR, int MODE, int NEG, int CSEC = 0>
__global__ void __launch_bounds__(HB_BLOCK, HB_MINBLOCKS) k_dop853_6(const PropParams p)
{
    double y[6], yh[6], k[13][6];
    const Cr3bpRhs<AR, NEG> rhs{p};
    // MODE_RECORD: each thread assembles its step record in its own (padded) shared-memory row and hands it to the
    // copy engine as ONE 512-byte bulk store.  Sixteen per-lane 32-byte vector stores instead cost the LSU one tag
    // lookup per lane per store and capped the kernel at ~1.6 TB/s of record writes (measured).
    extern __shared__ __align__(128) unsigned char rec_smem[];
    double *rec_row = (double *)(rec_smem + (is_record(MODE) ? threadIdx.x * HB_REC_ROW_BYTES : 0));
    double t = 0.0, h = 0.0, err_prev = -1.0, tf = 0.0, g_prev = 0.0;
    long long idx = -1;
    long long attempts = 0;
    int nacc = 0, nrej = 0, cursor = 0;
    // MODE_RECORD_NEAR: records written so far; the row holds the previous step's record, not written yet (pending);
    // that step could come near the section plane (prev_near)
    int nrec = 0;
    bool pending = false, prev_near = false;
    bool have = false, exhausted = false;
    for (;;) {
        if (!have && !exhausted) {
            idx = hb_fetch_index(p.ws);
            if (idx < p.n) {
                if (p.order) idx = p.order[idx];          // scheduling hint: outputs stay in the caller's indexing
#pragma unroll
                for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
                if (p.h0) {
                    // the pre-pass left h0 and the accelerations of f(y0) in the output rows: a start is loads only
                    h = p.h0[idx];
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const bool ng = NEG == 1 || (NEG == 2 && ((p.negmask >> d) & 1u));
                        k[0][d] = ng ? hb_flip_sign(y[3 + d]) : y[3 + d];
                        k[0][3 + d] = p.h0[(long long)(3 + d) * p.n + idx];
                    }
                } else {
                    crtbp_rhs<AR, NEG>(y, p, k[0]);
                    h = initial_step<AR>(y, k[0], p);
                }
                t = p.t0;
                tf = p.tf_arr ? p.tf_arr[idx] : p.tf;
                err_prev = -1.0;
                nacc = 0; nrej = 0; cursor = 0; attempts = 0;
                nrec = 0; pending = false; prev_near = false;
                if (MODE == MODE_EVENT) g_prev = AR::sub(pick6(y, p.ev_idx), p.ev_off);
                have = true;
                if (!((t - tf) < 0.0)) {
                    // zero-length span: nothing to integrate
                    if (MODE != MODE_DENSE) {
#pragma unroll
                        for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                    } else {
                        for (int c = 0; c < p.m; ++c)
#pragma unroll
                            for (int d = 0; d < 6; ++d)
                                p.dense_out[((long long)idx * p.m + c) * 6 + d] = y[d];
                    }
                    if (MODE == MODE_EVENT) p.t_hit[idx] = t;
                    p.nacc[idx] = 0; p.nrej[idx] = 0; p.status[idx] = HB_TRAJ_OK;
                    if (MODE == MODE_RECORD_NEAR) p.nrec[idx] = 0;
                    have = false;
                }
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(0xffffffffu, !have && exhausted)) break;
        if (!have) continue;

        // ---- one attempted step (rk.py:2452-2484) ----
        h = hb_clamp_step(h, p.max_step, p.min_step);
        if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
        // (a rolled stage loop -- one copy of the vector field, switch on the stage index -- was measured: 10 % less
        // SASS but +20 % time from the extra moves and spills; the unrolled form stays)
        dop853_stages<AR>(y, k, h, yh, rhs);
        double n5 = 0.0, n3 = 0.0;
        dop853_err_sums<AR>(y, yh, k, h, p.rtol, p.atol, n5, n3);
        const double err = dop853_err_norm<AR>(n5, n3, h, 6.0);
        ++attempts;
        int fin = -1;  // >= 0: trajectory finished with this status

        if (err <= 1.0) {
            const double t_new = AR::add(t, h);
            ++nacc;
            const bool last = !((t_new - tf) < 0.0);
            if (MODE == MODE_EVENT) {
                const double g_new = AR::sub(pick6(yh, p.ev_idx), p.ev_off);
                if (hb_event_crossed(g_prev, g_new, p.ev_dir)) {
                    // _dop853_refine_in_step (rk.py:2079-2102): bisection on the dense interpolant
                    double F[7][6], ym[6];
                    dense_cache<AR>(y, yh, h, k, F, rhs);
                    double a = 0.0, b = 1.0, g_left = g_prev, xh = 1.0;
                    bool found = false;
                    for (int it = 0; it < 128; ++it) {
                        const double mid = AR::mul(0.5, AR::add(a, b));
                        dense_eval<AR>(y, F, mid, ym);
                        const double g_mid = AR::sub(pick6(ym, p.ev_idx), p.ev_off);
                        if (fabs(g_mid) <= p.gtol) { xh = mid; found = true; break; }
                        if (hb_crossed_direction(g_left, g_mid, p.ev_dir)) b = mid;
                        else { a = mid; g_left = g_mid; }
                        if (AR::mul(AR::sub(b, a), fabs(h)) <= p.xtol) break;
                    }
                    if (!found) xh = b;
                    dense_eval<AR>(y, F, xh, ym);
                    p.t_hit[idx] = AR::madd(xh, h, t);
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = ym[d];
                    fin = HB_TRAJ_HIT;
                } else {
                    g_prev = g_new;
                    if (last) {
                        p.t_hit[idx] = t_new;
#pragma unroll
                        for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = yh[d];
                        fin = HB_TRAJ_OK;
                    }
                }
            } else if (MODE == MODE_DENSE) {
                // searchsorted(ts, t_q, 'right') - 1 semantics (rk.py:2505-2509): this segment owns
                // every grid time t_q < t_new, and the last segment owns everything that is left.
                if (cursor < p.m && (last || p.t_eval[cursor] < t_new)) {
                    const double hseg = AR::sub(t_new, t);
                    double F[7][6], yo[6];
                    if (hseg != 0.0) dense_cache<AR>(y, yh, hseg, k, F, rhs);
                    while (cursor < p.m) {
                        const double tq = p.t_eval[cursor];
                        if (!(last || tq < t_new)) break;
                        if (hseg == 0.0) {
#pragma unroll
                            for (int d = 0; d < 6; ++d) yo[d] = y[d];
                        } else {
                            dense_eval<AR>(y, F, AR::div(AR::sub(tq, t), hseg), yo);
                        }
                        double *o = p.dense_out + ((long long)idx * p.m + cursor) * 6;
#pragma unroll
                        for (int d = 0; d < 6; ++d) o[d] = yo[d];
                        ++cursor;
                    }
                }
                if (last) fin = HB_TRAJ_OK;
            } else if (MODE == MODE_RECORD) {
                // Store what the dense output of this accepted step depends on (512 B, sixteen 256-bit stores); the
                // section-scan kernels rebuild the interpolant where they need it (hb_section_scan.cu).
                if (nacc <= p.rec_cap) {
                    double *r = p.rec + ((long long)idx * p.rec_cap + (nacc - 1)) * HB_REC_DOUBLES;
                    // the previous record of this thread must have left the row (a whole step ago: no wait in practice)
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    assemble_record(rec_row, t, t_new, y, yh, k, nacc - 1, last);
                    store_record(rec_row, r);
                }
            } else if (MODE == MODE_RECORD_NEAR) {
                // Sparse records: only the steps whose interpolant can come near the section plane (screened in fast
                // arithmetic with a widened margin, dop853_step_near_plane) and their two neighbours -- the scan needs
                // the sample before and after a noted segment -- are written.  The row always holds the last accepted
                // step, so the step BEFORE a near one can still be written when the near one shows up.
                const double hseg = AR::sub(t_new, t);
                bool near = true;
                if (hseg != 0.0 && fabs(hseg) * p.inv_grid_dt >= 2.0) {      // shorter steps may own no grid sample: keep
                    const Cr3bpRhs<ArFast, NEG> rf{p};
                    const double off = p.sink.sec.offset, tol = p.sink.sec.tol_on_surface;
                    near = dop853_step_near_plane<CSEC>(y, yh, hseg, k, rf, off, tol);
                }
                if (near && pending) {                        // the step before this one: write it first
                    if (nrec < p.rec_cap) store_record(rec_row, p.rec + ((long long)idx * p.rec_cap + nrec) * HB_REC_DOUBLES);
                    ++nrec;
                }
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the row is free again
                assemble_record(rec_row, t, t_new, y, yh, k, nacc - 1, last);
                if (near || prev_near) {
                    if (nrec < p.rec_cap) store_record(rec_row, p.rec + ((long long)idx * p.rec_cap + nrec) * HB_REC_DOUBLES);
                    ++nrec;
                    pending = false;
                } else {
                    pending = true;
                }
                prev_near = near;
            }
            if (is_record(MODE) || MODE == MODE_FINAL) {   // the dense interpolant at tf on the last segment
                if (last) {
                    const double hseg = AR::sub(t_new, t);
                    double yo[6];
                    if (hseg == 0.0) {
#pragma unroll
                        for (int d = 0; d < 6; ++d) yo[d] = y[d];
                    } else {
                        const double x = AR::div(AR::sub(tf, t), hseg);
                        if (x == 1.0) {
                            // every (1-x) factor is zero: y_old + (y_new - y_old)
#pragma unroll
                            for (int d = 0; d < 6; ++d) yo[d] = AR::add(AR::sub(yh[d], y[d]), y[d]);
                        } else {
                            double F[7][6];
                            dense_cache<AR>(y, yh, hseg, k, F, rhs);
                            dense_eval<AR>(y, F, x, yo);
                        }
                    }
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = yo[d];
                    fin = ((MODE == MODE_RECORD && nacc > p.rec_cap) || (MODE == MODE_RECORD_NEAR && nrec > p.rec_cap))
                              ? HB_TRAJ_RECORD_OVERFLOW : HB_TRAJ_OK;
                }
            }
            // advance
            t = t_new;
        } else {
            ++nrej;
        }
        {   // y <- y_new, k1 <- k13 (FSAL) as selects outside the branch: the compiler merged the two paths with ~75 register
            // moves per step when the copy sat inside it (measured: 1.2 % of the record kernel)
            const bool acc_ = err <= 1.0;
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = acc_ ? yh[d] : y[d]; k[0][d] = acc_ ? k[12][d] : k[0][d]; }
        }
        {   // The controller's factor AFTER the record / screening block, where the stage vectors are dead (measured: 1 % off
            // the record kernel).  One convergent pow: accepted and rejected lanes evaluate err**(-1/9) and err**(-1/8) in the
            // same instructions (the exponent is a per-lane operand); err_prev**alpha is carried from the step where
            // err_prev was the current error and shares that step's logarithm (in the parity build `err_prev` holds that
            // power, -1 before the first accepted step).
            const bool accepted = err <= 1.0;
            const double h_factor = hb_pi_factor_carried<AR>(err, err_prev, accepted, 8.0);   // err_prev: see there
            h = AR::mul(h, h_factor);
            if (!accepted) h = hb_clamp_step(h, p.max_step, p.min_step);
        }
        if (fin < 0) {
            if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
            else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            if (fin >= 0) {
                if (MODE != MODE_DENSE) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                }
                if (MODE == MODE_EVENT) p.t_hit[idx] = t;
            }
        }
        if (fin >= 0) {
            p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin;
            if (MODE == MODE_RECORD_NEAR) p.nrec[idx] = nrec;
            have = false;
        }
    }
    if (is_record(MODE)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
template <class AR, int MODE, int NEG, int CSEC = 0>
int launch_one(const PropParams &p, unsigned grid, cudaStream_t st)
{
    constexpr int smem = is_record(MODE) ? HB_BLOCK * HB_REC_ROW_BYTES : 0;
    // the opt-in is per device (and cheap): set on every launch, so a second device in the same process gets it too
    if (smem > 48 * 1024)
        HB_CUDA_TRY(cudaFuncSetAttribute(k_dop853_6<AR, MODE, NEG, CSEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_dop853_6<AR, MODE, NEG, CSEC><<<grid, HB_BLOCK, smem, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

template <class AR, int MODE, int NEG>
int launch_sec(const PropParams &p, unsigned grid, cudaStream_t st)
{
    if constexpr (MODE == MODE_RECORD_NEAR) {
        switch (p.sink.sec.idx) {
        case 0: return launch_one<AR, MODE, NEG, 0>(p, grid, st);
        case 1: return launch_one<AR, MODE, NEG, 1>(p, grid, st);
        case 2: return launch_one<AR, MODE, NEG, 2>(p, grid, st);
        case 3: return launch_one<AR, MODE, NEG, 3>(p, grid, st);
        case 4: return launch_one<AR, MODE, NEG, 4>(p, grid, st);
        case 5: return launch_one<AR, MODE, NEG, 5>(p, grid, st);
        default: return HB_ERR_BADARG;
        }
    } else {
        return launch_one<AR, MODE, NEG>(p, grid, st);
    }
}

template <class AR, int MODE>
int launch_neg(const PropParams &p, unsigned grid, cudaStream_t st)
{
    if (p.negmask == 0u) return launch_sec<AR, MODE, 0>(p, grid, st);
    if (p.negmask == 63u) return launch_sec<AR, MODE, 1>(p, grid, st);
    return launch_sec<AR, MODE, 2>(p, grid, st);
}

template <int MODE>
int launch(PropParams &p, int arith, cudaStream_t st)
{
    HB_CUDA_TRY(cudaMemsetAsync(p.ws, 0, sizeof(HbWorkspace), st));
    if (arith == HB_ARITH_PARITY && MODE != MODE_DENSE) {
        const int rc = first_steps_prepass<ArParity>(p, st);
        if (rc != HB_OK) return rc;
    }
    long long blocks_needed = (p.n + HB_BLOCK - 1) / HB_BLOCK;
    long long grid = (long long)HB_MINBLOCKS * sm_count();   // persistent: resident CTAs only
    if (p.max_ctas > 0 && p.max_ctas < grid) grid = p.max_ctas;   // the caller shares the device between concurrent batches
    if (blocks_needed < grid) grid = blocks_needed;
    if (grid < 1) grid = 1;
    if (arith == HB_ARITH_PARITY) return launch_neg<ArParity, MODE>(p, (unsigned)grid, st);
    return launch_neg<ArFast, MODE>(p, (unsigned)grid, st);
}

}  // namespace

extern "C" {

int64_t hb_workspace_bytes(void) { return (int64_t)sizeof(HbWorkspace); }

int hb_device_info(int32_t *sm, int32_t *major, int32_t *minor)
{
    int dev = 0, n = 0, ma = 0, mi = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HB_ERR_NODEVICE;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return HB_ERR_NODEVICE;
    cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
    if (sm) *sm = n;
    if (major) *major = ma;
    if (minor) *minor = mi;
    return HB_OK;
}

int hb_cr3bp_propagate(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *y0_soa,
                       double t0, double tf, const double *tf_per_traj, int32_t n_fixed_steps,
                       double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status,
                       void *workspace, void *stream)
{
    PropParams p{};
    int rc = fill_params(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (n < 0 || !workspace || (n > 0 && (!y0_soa || !yf_soa || !n_acc || !n_rej || !status))) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    p.n = n; p.y0 = y0_soa; p.t0 = t0; p.tf = tf; p.tf_arr = tf_per_traj;
    p.yf = yf_soa; p.nacc = n_acc; p.nrej = n_rej; p.status = status;
    p.ws = (HbWorkspace *)workspace;
    if (integ->method != HB_DOP853)
        return hb_rk_dispatch(p, integ->method, integ->arith, 0, n_fixed_steps > 0 ? n_fixed_steps : integ->n_fixed_steps,
                              (cudaStream_t)stream);
    return launch<MODE_FINAL>(p, integ->arith, (cudaStream_t)stream);
}

int hb_cr3bp_dense(const hb_cr3bp *sys, const hb_integ *integ, int64_t n, const double *y0_soa,
                   const double *t_eval, int32_t m, double *states_out, int32_t *n_acc,
                   int32_t *n_rej, int32_t *status, void *workspace, void *stream)
{
    PropParams p{};
    int rc = fill_params(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (n < 0 || m < 2 || !workspace || !t_eval ||
        (n > 0 && (!y0_soa || !states_out || !n_acc || !n_rej || !status)))
        return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    // t0 / tf are the ends of the grid (rk.py:2431-2432); fetch them from the device array
    double ends[2];
    cudaStream_t st = (cudaStream_t)stream;
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[0], t_eval, sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[1], t_eval + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    p.n = n; p.y0 = y0_soa; p.t0 = ends[0]; p.tf = ends[1]; p.tf_arr = nullptr;
    p.nacc = n_acc; p.nrej = n_rej; p.status = status;
    p.t_eval = t_eval; p.m = m; p.dense_out = states_out;
    p.ws = (HbWorkspace *)workspace;
    if (integ->method != HB_DOP853) return hb_rk_dispatch(p, integ->method, integ->arith, 1, 0, st);
    return launch<MODE_DENSE>(p, integ->arith, st);
}

int hb_cr3bp_event(const hb_cr3bp *sys, const hb_integ *integ, const hb_event *ev, int64_t n,
                   const double *y0_soa, double t0, double tmax, const double *tmax_per_traj,
                   double *t_hit, double *y_hit_soa, int32_t *n_acc, int32_t *n_rej,
                   int32_t *status, void *workspace, void *stream)
{
    PropParams p{};
    int rc = fill_params(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (!ev || ev->idx < 0 || ev->idx > 5) return HB_ERR_BADARG;
    if (n < 0 || !workspace || (n > 0 && (!y0_soa || !y_hit_soa || !t_hit || !n_acc || !n_rej || !status)))
        return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    p.n = n; p.y0 = y0_soa; p.t0 = t0; p.tf = tmax; p.tf_arr = tmax_per_traj;
    p.yf = y_hit_soa; p.t_hit = t_hit; p.nacc = n_acc; p.nrej = n_rej; p.status = status;
    p.ev_idx = ev->idx; p.ev_dir = ev->direction; p.ev_off = ev->offset; p.xtol = ev->xtol; p.gtol = ev->gtol;
    p.ws = (HbWorkspace *)workspace;
    if (integ->method != HB_DOP853)
        return hb_rk_dispatch(p, integ->method, integ->arith, 2, integ->n_fixed_steps, (cudaStream_t)stream);
    return launch<MODE_EVENT>(p, integ->arith, (cudaStream_t)stream);
}

// Kernel A of the two-kernel section path (hb_section_scan.cu): propagate and record every step's interpolant.
int hb_cr3bp_record_launch(const hb_cr3bp *sys, const hb_integ *integ, int32_t section_idx, int64_t n,
                           const double *y0_soa, double t0, double tf, double *rec, int32_t rec_cap, double *yf_soa,
                           int32_t *n_acc, int32_t *n_rej, int32_t *status, void *workspace, cudaStream_t st,
                           const hb_section *near_section, double inv_grid_dt, int32_t *n_rec)
{
    PropParams p{};
    int rc = fill_params(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (integ->method != HB_DOP853) return HB_ERR_UNSUPPORTED;
    p.n = n; p.y0 = y0_soa; p.t0 = t0; p.tf = tf; p.tf_arr = nullptr;
    p.yf = yf_soa; p.nacc = n_acc; p.nrej = n_rej; p.status = status;
    p.rec = rec; p.rec_cap = rec_cap; p.sink.sec.idx = section_idx;
    p.ws = (HbWorkspace *)workspace;
    if (near_section) {                                  // sparse records: only steps near this section plane
        p.sink.sec = *near_section;
        p.inv_grid_dt = inv_grid_dt;
        p.nrec = n_rec;
        return launch<MODE_RECORD_NEAR>(p, integ->arith, st);
    }
    return launch<MODE_RECORD>(p, integ->arith, st);
}

}  // extern "C"
