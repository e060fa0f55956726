// hb_cr3bp_section.cu -- fused manifold tube + synodic section: 6-state DOP853 with in-kernel section
// detection in the reference's semantics (hits are linear interpolants between dense samples on the
// dt grid; see hb_section.cuh), without ever storing the 226 KB/trajectory dense tube.
//
// Same one-trajectory-per-thread persistent kernel as hb_cr3bp.cu, plus three things that keep the
// ~4700 grid samples per trajectory from dominating:
//   1. only the EVENT COMPONENT of a sample is evaluated (1/6 of the dense polynomial); full states are
//      evaluated -- with identical arithmetic -- for segments that can hold a hit and for the last sample
//      of every step (it pairs with the first sample of the next step);
//   2. QUIET STEPS are skipped rigorously: on [0,1] the interpolant p(x) = y0 + x(F0 + (1-x)(F1 + x(F2 + ...)))
//      satisfies |p(x) - (y0 + x F0)| <= x(1-x) * sum_{i>=1}|F_i| <= S/4, so when both step ends are on the
//      same side of the plane by more than S/4 + tol_on_surface (+ rounding slack) no sample of the step can
//      change sign or be "on surface"; then only its first sample and its last two are evaluated;
//   3. the remaining NON-QUIET steps (those near the plane) are scanned WARP-COOPERATIVELY: the owner
//      lane broadcasts its 7 event-component coefficients and all 32 lanes evaluate 32 consecutive
//      grid samples at once, so one lane's 50-sample loop no longer stalls the other 31.
// Hits, hit order, de-duplication and end states are bit-identical to hb_cr3bp_dense + hb_synodic_detect
// (tests/test_gpu_synodic.py::test_fused_section_equals_two_kernel_chain).
//
// Reference: system/manifold.py:226 + system/maps/synodic.py:121, i.e. algorithms/integrators/rk.py:2377-2549
// followed by algorithms/poincare/synodic/backend.py:458-659.
#include "hb_cr3bp_common.cuh"

namespace {
using namespace hbc;

// event component of the dense interpolant at parameter xq (dense_eval of hb_dop853.cuh on one component)
template <class AR>
HB_DEV double g_component(const double (&Fe)[7], double ye_old, double hseg, double xq, double offset)
{
    double ge = ye_old;
    if (hseg != 0.0) {
        const double omx = AR::sub(1.0, xq);
        double v = 0.0;
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            v = AR::add(v, Fe[i]);
            v = AR::mul(v, ((6 - i) % 2 == 0) ? xq : omx);
        }
        ge = AR::add(v, ye_old);
    }
    return __dsub_rn(ge, offset);
}

template <class AR>
HB_DEV double x_param(double tq, double t, double hseg) { return (hseg == 0.0) ? 0.0 : AR::div(AR::sub(tq, t), hseg); }

template <class AR, int NEG>
__global__ void __launch_bounds__(HB_BLOCK, HB_MINBLOCKS) k_dop853_6_section(const PropParams p)
{
    const int lane = threadIdx.x & 31;
    double y[6], yh[6], k[13][6];
    const Cr3bpRhs<AR, NEG> rhs{p};
    double t = 0.0, h = 0.0, err_prev = -1.0, tf = 0.0;
    long long idx = -1, attempts = 0;
    int nacc = 0, nrej = 0, cursor = 0;
    bool have = false, exhausted = false;
    // detector context: last grid sample (state, event value), the event value before it, de-dup state
    double xs_prev[6], gs_prev = 0.0, gs_prev2 = 0.0;
    Dedup dd{0.0, 0.0, 0.0, 0};
    bool sec_alive = true;
    const double off = p.sink.sec.offset, tol_s = p.sink.sec.tol_on_surface;
    const int sidx = p.sink.sec.idx;

    for (;;) {
        if (!have && !exhausted) {
            idx = hb_fetch_index(p.ws);
            if (idx < p.n) {
#pragma unroll
                for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
                crtbp_rhs<AR, NEG>(y, p, k[0]);
                t = p.t0;
                tf = p.tf;
                h = p.h0 ? p.h0[idx] : initial_step<AR>(y, k[0], p);
                err_prev = -1.0;
                nacc = 0; nrej = 0; cursor = 0; attempts = 0;
                dd = Dedup{0.0, 0.0, 0.0, 0};
                sec_alive = true;
                have = true;
                if (!((t - tf) < 0.0)) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                    if (p.hits_per_traj) p.hits_per_traj[idx] = 0;
                    p.nacc[idx] = 0; p.nrej[idx] = 0; p.status[idx] = HB_TRAJ_OK;
                    have = false;
                }
            } else {
                exhausted = true;
            }
        }
        // Parity build: keep the CTA's warps in the same code region.  Its unrolled step + dense cache is ~75 KB
        // of SASS, far beyond the instruction cache, and warps that drift apart each stream it from L2 on their
        // own (ncu: no_instruction 8.6 stall cycles per issue); one CTA barrier per step cut 21.1 -> 15.0 ms.
        // The fast build (smaller code) measured slower with the barrier (10.3 -> 12.4 ms) and keeps warp scope.
        if constexpr (AR::parity) {
            if (__syncthreads_and(!have && exhausted)) break;
        } else {
            if (__all_sync(0xffffffffu, !have && exhausted)) break;
        }

        // ---- phase A: one attempted step per active lane (rk.py:2452-2484) ----
        double err = 0.0, t_new = t;
        bool accepted = false;
        if (have) {
            h = hb_clamp_step(h, p.max_step, p.min_step);
            if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
            dop853_stages<AR>(y, k, h, yh, rhs);
            double n5 = 0.0, n3 = 0.0;
            dop853_err_sums<AR>(y, yh, k, h, p.rtol, p.atol, n5, n3);
            err = dop853_err_norm<AR>(n5, n3, h, 6.0);
            ++attempts;
            accepted = err <= 1.0;
            t_new = AR::add(t, h);
        }
        const bool last = accepted && !((t_new - tf) < 0.0);

        // ---- phase B (per lane): dense coefficients of the accepted segment, quiet test ----
        double F[7][6], Fe[7], ye_old = 0.0, hseg = 0.0, xp_prev = 0.0;
        int cend = cursor;
        bool scan = false, prev_here = false;
        if (accepted && cursor < p.m && (last || p.t_eval[cursor] < t_new)) {
            hseg = AR::sub(t_new, t);
            if (hseg != 0.0) dense_cache<AR>(y, yh, hseg, k, F, rhs);
#pragma unroll
            for (int i = 0; i < 7; ++i) Fe[i] = pick(F[i], sidx);
            ye_old = pick(y, sidx);
            cend = p.m;
            if (!last) {   // samples owned by this segment: t_eval[c] < t_new  (searchsorted 'right' - 1, rk.py:2505)
                int c = cursor + (int)fmin(fmax((t_new - p.t_eval[cursor]) * p.inv_grid_dt, 0.0), (double)(p.m - cursor));
                while (c < p.m && p.t_eval[c] < t_new) ++c;
                while (c > cursor && !(p.t_eval[c - 1] < t_new)) --c;
                cend = c;
            }
            bool quiet = false;
            if (hseg != 0.0 && cend - cursor > 4) {
                const double g_old = __dsub_rn(ye_old, off);
                const double g_new = __dsub_rn(pick(yh, sidx), off);
                double S = 0.0;
#pragma unroll
                for (int i = 1; i < 7; ++i) S += fabs(Fe[i]);
                const double margin = 0.25 * S + tol_s + 1e-9 * (fabs(ye_old) + fabs(Fe[0]) + fabs(off)) + 1e-290;
                const bool same = (g_old > 0.0 && g_new > 0.0) || (g_old < 0.0 && g_new < 0.0);
                quiet = same && fmin(fabs(g_old), fabs(g_new)) > margin;
            }
            if (quiet) {
                // first sample of the step against the previous step's last sample
                const double tq = p.t_eval[cursor];
                const double xq = x_param<AR>(tq, t, hseg);
                const double g_now = g_component<AR>(Fe, ye_old, hseg, xq, off);
                if (cursor > 0 && sec_alive) {
                    const bool same = (gs_prev > 0.0 && g_now > 0.0) || (gs_prev < 0.0 && g_now < 0.0);
                    if (!same || fabs(gs_prev) < tol_s) {
                        double yo[6];
                        dense_eval<AR>(y, F, xq, yo);
                        sec_alive = process_segment(p.sink, dd, idx, 0, cursor > 1, gs_prev2,
                                                    __dmul_rn(p.tsign, p.t_eval[cursor - 1]), __dmul_rn(p.tsign, tq),
                                                    xs_prev, yo);
                    }
                }
                gs_prev2 = g_component<AR>(Fe, ye_old, hseg, x_param<AR>(p.t_eval[cend - 2], t, hseg), off);
                xp_prev = x_param<AR>(p.t_eval[cend - 1], t, hseg);
                gs_prev = g_component<AR>(Fe, ye_old, hseg, xp_prev, off);
                cursor = cend;
                prev_here = true;
            } else {
                scan = true;
            }
        }

        // ---- phase C (whole warp): cooperative scan of the non-quiet segments, one owner lane at a time ----
        unsigned req = __ballot_sync(0xffffffffu, scan);
        while (req) {
            const int L = __ffs(req) - 1;
            req &= req - 1;
            double bFe[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) bFe[i] = shfl_d(Fe[i], L);
            const double b_ye = shfl_d(ye_old, L), b_t = shfl_d(t, L), b_h = shfl_d(hseg, L);
            const int c0 = __shfl_sync(0xffffffffu, cursor, L), c1 = __shfl_sync(0xffffffffu, cend, L);
            double carry1 = shfl_d(gs_prev, L), carry2 = shfl_d(gs_prev2, L);
            for (int base = c0; base < c1; base += 32) {
                const int c = base + lane;
                const bool valid = c < c1;
                const double tq = p.t_eval[valid ? c : c1 - 1];
                const double g = g_component<AR>(bFe, b_ye, b_h, x_param<AR>(tq, b_t, b_h), off);
                double g_m1 = __shfl_up_sync(0xffffffffu, g, 1);
                double g_m2 = __shfl_up_sync(0xffffffffu, g, 2);
                if (lane == 0) { g_m1 = carry1; g_m2 = carry2; }
                if (lane == 1) g_m2 = carry1;
                const bool same = (g_m1 > 0.0 && g > 0.0) || (g_m1 < 0.0 && g < 0.0);
                unsigned fm = __ballot_sync(0xffffffffu, valid && c > 0 && (!same || fabs(g_m1) < tol_s));
                while (fm) {                         // rare: segments that can hold a hit, in grid order
                    const int i = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const double gk = shfl_d(g_m1, i), gk1 = shfl_d(g, i), gm2 = shfl_d(g_m2, i);
                    const int cs = base + i;         // segment (cs-1, cs)
                    const double t0s = __dmul_rn(p.tsign, p.t_eval[cs - 1]), t1s = __dmul_rn(p.tsign, p.t_eval[cs]);
                    auto eval = [&](int which, double (&out)[6]) {
                        const int cc = cs - 1 + which;
                        if (cc < c0) {               // the previous step's last sample
#pragma unroll
                            for (int d = 0; d < 6; ++d) out[d] = xs_prev[d];
                        } else if (hseg == 0.0) {
#pragma unroll
                            for (int d = 0; d < 6; ++d) out[d] = y[d];
                        } else {
                            dense_eval<AR>(y, F, x_param<AR>(p.t_eval[cc], t, hseg), out);
                        }
                    };
                    const bool owner = lane == L;
                    const bool al = process_segment_coop(p.sink, dd, idx, lane, owner && sec_alive, cs > 1, gm2, gk, gk1,
                                                         t0s, t1s, eval);
                    if (owner && sec_alive) sec_alive = al;
                }
                const int nvalid = min(32, c1 - base);
                const double last1 = shfl_d(g, nvalid - 1);
                const double last2 = shfl_d(g, nvalid >= 2 ? nvalid - 2 : 0);
                carry2 = (nvalid >= 2) ? last2 : carry1;
                carry1 = last1;
            }
            if (lane == L) {
                gs_prev = carry1;
                gs_prev2 = carry2;
                xp_prev = x_param<AR>(p.t_eval[c1 - 1], t, hseg);
                cursor = c1;
                prev_here = true;
            }
        }

        // ---- phase D (per lane): context for the next step, advance, controller ----
        if (prev_here) {
            if (hseg == 0.0) {
#pragma unroll
                for (int d = 0; d < 6; ++d) xs_prev[d] = y[d];
            } else {
                dense_eval<AR>(y, F, xp_prev, xs_prev);
            }
        }
        int fin = -1;
        const double h_factor = hb_pi_factor<AR>(err, err_prev, accepted, 8.0);   // both branches, one pow
        if (accepted) {
            ++nacc;
            if (last) {
#pragma unroll
                for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = xs_prev[d];
                fin = HB_TRAJ_OK;
            }
            t = t_new;
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = k[12][d]; }
            h = AR::mul(h, h_factor);
            err_prev = err;
        } else if (have) {
            ++nrej;
            h = AR::mul(h, h_factor);
            h = hb_clamp_step(h, p.max_step, p.min_step);
        }
        if (have && fin < 0) {
            if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
            else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
            if (fin >= 0) {
#pragma unroll
                for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
            }
        }
        if (fin >= 0) {
            p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin;
            if (p.hits_per_traj) p.hits_per_traj[idx] = dd.n;
            have = false;
        }
    }
}

template <class AR>
int launch_section(const PropParams &p_in, cudaStream_t st)
{
    PropParams p = p_in;
    HB_CUDA_TRY(cudaMemsetAsync(p.ws, 0, sizeof(HbWorkspace), st));
    if constexpr (AR::parity) {
        const int rc = first_steps_prepass<AR>(p, st);
        if (rc != HB_OK) return rc;
    }
    long long blocks_needed = (p.n + HB_BLOCK - 1) / HB_BLOCK;
    long long grid = (long long)HB_MINBLOCKS * sm_count();
    if (p.max_ctas > 0 && p.max_ctas < grid) grid = p.max_ctas;
    if (blocks_needed < grid) grid = blocks_needed;
    if (grid < 1) grid = 1;
    if (p.negmask == 0u) k_dop853_6_section<AR, 0><<<(unsigned)grid, HB_BLOCK, 0, st>>>(p);
    else if (p.negmask == 63u) k_dop853_6_section<AR, 1><<<(unsigned)grid, HB_BLOCK, 0, st>>>(p);
    else k_dop853_6_section<AR, 2><<<(unsigned)grid, HB_BLOCK, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

}  // namespace

extern "C" int hb_cr3bp_section(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                                const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits,
                                int64_t hit_capacity, int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc,
                                int32_t *n_rej, int32_t *status, void *workspace, void *stream)
{
    PropParams p{};
    int rc = fill_params(sys, integ, p);
    if (rc != HB_OK) return rc;
    if (!sec || sec->idx < 0 || sec->idx > 5 || sec->proj_i < 0 || sec->proj_i > 5 || sec->proj_j < 0 ||
        sec->proj_j > 5 || sec->segment_refine < 0 || hit_capacity < 0)
        return HB_ERR_BADARG;
    if (n < 0 || m < 2 || !workspace || !t_eval ||
        (n > 0 && (!y0_soa || !yf_soa || !n_acc || !n_rej || !status || (hit_capacity > 0 && !hits))))
        return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st)); return HB_OK; }
    double ends[2];
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[0], t_eval, sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[1], t_eval + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    p.n = n; p.y0 = y0_soa; p.t0 = ends[0]; p.tf = ends[1]; p.tf_arr = nullptr;
    p.yf = yf_soa; p.nacc = n_acc; p.nrej = n_rej; p.status = status;
    p.t_eval = t_eval; p.m = m;
    p.ws = (HbWorkspace *)workspace;
    p.sink.sec = *sec; p.sink.hits = hits; p.sink.capacity = hit_capacity; p.sink.ws = p.ws;
    p.hits_per_traj = hits_per_traj;
    p.tsign = sys->fwd < 0 ? -1.0 : 1.0;
    p.inv_grid_dt = (ends[1] > ends[0]) ? (double)(m - 1) / (ends[1] - ends[0]) : 0.0;
    return (integ->arith == HB_ARITH_PARITY) ? launch_section<ArParity>(p, st) : launch_section<ArFast>(p, st);
}
