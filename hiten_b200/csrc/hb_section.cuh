// hb_section.cuh -- the reference's synodic-section detector on ONE sample segment, plus hit de-duplication.
// Shared by the stand-alone detector (hb_synodic.cu, warp per trajectory, executed warp-uniformly) and the fused
// section mode of the DOP853 kernel (hb_cr3bp.cu, thread per trajectory).
// Reference: hiten/algorithms/poincare/synodic/backend.py  _detect_with_segment_refine :458-659 (linear branch),
// detect_on_trajectory :782-821 (segment_refine == 0), _order_and_dedup_hits :382-455.
#pragma once
#include "hb_common.cuh"

struct HitSink {
    hb_section sec;
    hb_hit *hits;
    long long capacity;
    HbWorkspace *ws;
};

HB_DEV double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

struct Dedup {
    double last_t, last_u, last_v;
    int n;
};

// _order_and_dedup_hits (backend.py:440-454); returns false when the per-trajectory cap is reached
HB_DEV bool push_hit(const HitSink &p, Dedup &dd, long long traj, double th, const double (&xh)[6], int lane)
{
    const int mh = p.sec.max_hits_per_traj;
    if (mh > 0 && dd.n >= mh) return false;
    const double u = (p.sec.proj_i == 0) ? xh[0] : (p.sec.proj_i == 1) ? xh[1] : (p.sec.proj_i == 2) ? xh[2]
                   : (p.sec.proj_i == 3) ? xh[3] : (p.sec.proj_i == 4) ? xh[4] : xh[5];
    const double v = (p.sec.proj_j == 0) ? xh[0] : (p.sec.proj_j == 1) ? xh[1] : (p.sec.proj_j == 2) ? xh[2]
                   : (p.sec.proj_j == 3) ? xh[3] : (p.sec.proj_j == 4) ? xh[4] : xh[5];
    if (dd.n > 0) {
        if (fabs(__dsub_rn(th, dd.last_t)) <= p.sec.dedup_time_tol) return true;
        const double du = __dsub_rn(u, dd.last_u), dv = __dsub_rn(v, dd.last_v);
        const double d2 = __dadd_rn(__dmul_rn(du, du), __dmul_rn(dv, dv));
        if (d2 <= __dmul_rn(p.sec.dedup_point_tol, p.sec.dedup_point_tol)) return true;
    }
    if (lane == 0) {
        const unsigned long long slot = atomicAdd(&p.ws->hit_count, 1ULL);
        if ((long long)slot < p.capacity) {
            hb_hit *h = p.hits + slot;
            h->traj = traj; h->seq = dd.n; h->t = th;
#pragma unroll
            for (int d = 0; d < 6; ++d) h->state[d] = xh[d];
        } else {
            atomicAdd(&p.ws->overflow, 1ULL);
        }
    }
    dd.last_t = th; dd.last_u = u; dd.last_v = v;
    dd.n++;
    return true;
}

HB_DEV double pick(const double (&x)[6], int i)
{
    double r = x[0];
#pragma unroll
    for (int d = 1; d < 6; ++d) r = (i == d) ? x[d] : r;
    return r;
}

// Full per-segment logic of _detect_with_segment_refine / the r == 0 path, executed warp-uniformly.
HB_DEV bool process_segment(const HitSink &p, Dedup &dd, long long traj, int lane, bool has_prev, double g_prev,
                            double t0, double t1, const double (&x0)[6], const double (&x1)[6])
{
    const int dir = p.sec.direction;
    const double gk = __dsub_rn(pick(x0, p.sec.idx), p.sec.offset);
    const double gk1 = __dsub_rn(pick(x1, p.sec.idx), p.sec.offset);
    bool accept_left = false;
    if (fabs(gk) < p.sec.tol_on_surface) {
        if (dir == 0) accept_left = true;
        else if (dir > 0) accept_left = (gk1 >= 0.0) || (has_prev && g_prev <= 0.0);
        else accept_left = (gk1 <= 0.0) || (has_prev && g_prev >= 0.0);
    }
    const int r = p.sec.segment_refine;
    double xh[6];
    if (r > 0) {
        if (accept_left && !push_hit(p, dd, traj, t0, x0, lane)) return false;
        const double step = __ddiv_rn(1.0, (double)(r + 1));
        for (int mm = 0; mm <= r; ++mm) {
            const double s_lo = __dmul_rn((double)mm, step), s_hi = __dmul_rn((double)(mm + 1), step);
            if (s_hi > 1.0 + 1e-15) break;
            if (accept_left && mm == 0) continue;
            const double g_lo = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_lo), gk), __dmul_rn(s_lo, gk1));
            const double g_hi = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_hi), gk), __dmul_rn(s_hi, gk1));
            bool crosses;
            if (dir == 0) crosses = (__dmul_rn(g_lo, g_hi) <= 0.0) && (g_lo != g_hi);
            else if (dir > 0) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
            else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
            if (!crosses) continue;
            double s_star;
            if (g_lo == g_hi) s_star = __dmul_rn(0.5, __dadd_rn(s_lo, s_hi));
            else {
                double al = __ddiv_rn(g_lo, __dsub_rn(g_lo, g_hi));
                al = fmin(1.0, fmax(0.0, al));
                s_star = __dadd_rn(s_lo, __dmul_rn(al, __dsub_rn(s_hi, s_lo)));
            }
            const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_star), t0), __dmul_rn(s_star, t1));
#pragma unroll
            for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(s_star, __dsub_rn(x1[d], x0[d])));
            if (!push_hit(p, dd, traj, th, xh, lane)) return false;
        }
    } else {
        if (accept_left) return push_hit(p, dd, traj, t0, x0, lane);
        bool crosses;
        if (dir == 0) crosses = (__dmul_rn(gk, gk1) <= 0.0) && (gk != gk1);
        else if (dir > 0) crosses = (gk < 0.0) && (gk1 >= 0.0);
        else crosses = (gk > 0.0) && (gk1 <= 0.0);
        if (!crosses) return true;
        double al = __ddiv_rn(gk, __dsub_rn(gk, gk1));
        al = fmin(1.0, fmax(0.0, al));
        const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), t0), __dmul_rn(al, t1));
#pragma unroll
        for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(al, __dsub_rn(x1[d], x0[d])));
        return push_hit(p, dd, traj, th, xh, lane);
    }
    return true;
}

// ---- the CUBIC branch of the detector (interp_kind == "cubic") --------------------------------------------------------
// Reference: backend.py _detect_with_segment_refine :541-553 (slopes of g from the neighbouring samples), :584-631 (Hermite g
// on the sub-intervals, Newton on the cubic clamped to the sub-interval), :627-645 (cubic Hermite hit state),
// _refine_hits_cubic :274-379 (segment_refine == 0: Newton clamped to [0, 1]); poincare/utils.py _hermite_scalar :54-98,
// _hermite_der :101-148.  _hermite_* are Numba functions, where `x ** 2` with a literal exponent is x * x; the weights of the
// hit state are plain Python floats, where `x ** 2` is libm pow(x, 2.0) (hb_pow_libm restates glibc's pow bit for bit).
// Every cubic formula is guarded by `dt > 0.0` in the reference, so a trajectory with decreasing times (a backward tube)
// gets the linear formulas -- reproduced.
HB_DEV double hb_sq_rn(double x) { return __dmul_rn(x, x); }
HB_DEV double hermite_scalar(double s, double y0, double y1, double dy0, double dy1, double dt)
{
    const double oms = __dsub_rn(1.0, s);
    const double h00 = __dmul_rn(__dadd_rn(1.0, __dmul_rn(2.0, s)), hb_sq_rn(oms));
    const double h10 = __dmul_rn(s, hb_sq_rn(oms));
    const double h01 = __dmul_rn(hb_sq_rn(s), __dsub_rn(3.0, __dmul_rn(2.0, s)));
    const double h11 = __dmul_rn(hb_sq_rn(s), __dsub_rn(s, 1.0));
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h00, y0), __dmul_rn(__dmul_rn(h10, dy0), dt)), __dmul_rn(h01, y1)),
                     __dmul_rn(__dmul_rn(h11, dy1), dt));
}
HB_DEV double hermite_der(double s, double y0, double y1, double dy0, double dy1, double dt)
{
    const double oms = __dsub_rn(1.0, s), sm1 = __dsub_rn(s, 1.0), two_s = __dmul_rn(2.0, s), six_s = __dmul_rn(6.0, s);
    const double dh00 = __dsub_rn(__dadd_rn(__dmul_rn(six_s, sm1), __dmul_rn(hb_sq_rn(oms), 2.0)),
                                  __dmul_rn(__dmul_rn(2.0, oms), __dadd_rn(1.0, two_s)));
    const double dh10 = __dadd_rn(hb_sq_rn(oms), __dmul_rn(s, __dmul_rn(2.0, sm1)));
    const double dh01 = __dsub_rn(__dmul_rn(six_s, oms), __dmul_rn(two_s, __dsub_rn(3.0, two_s)));
    const double dh11 = __dadd_rn(__dmul_rn(two_s, sm1), hb_sq_rn(s));
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dh00, y0), __dmul_rn(__dmul_rn(dh10, dy0), dt)), __dmul_rn(dh01, y1)),
                     __dmul_rn(__dmul_rn(dh11, dy1), dt));
}
// CPython float ** 2 (float_pow: shortcuts for a base of 1 and 0, libm pow otherwise)
HB_DEV double hb_py_sq(double x)
{
    if (x == 1.0) return 1.0;
    if (x == 0.0) return 0.0;
    return hb_pow_libm(x, 2.0);
}
// hit state at s on segment k (backend.py:627-645): cubic Hermite through samples k-1 .. k+2 when they exist
HB_DEV void cubic_hit_state(const double *T, const double *X, int m, int k, double s, double dt, double (&xh)[6])
{
    const double *x0 = X + (long long)k * 6, *x1 = x0 + 6;
    if (dt > 0.0 && k - 1 >= 0 && k + 2 < m) {
        const double *xm = x0 - 6, *xp = x0 + 12;
        const double dtm = __dsub_rn(T[k + 1], T[k - 1]), dtp = __dsub_rn(T[k + 2], T[k]);
        const double oms = __dsub_rn(1.0, s);
        const double h00 = __dmul_rn(__dadd_rn(1.0, __dmul_rn(2.0, s)), hb_py_sq(oms));
        const double h10 = __dmul_rn(s, hb_py_sq(oms));
        const double h01 = __dmul_rn(hb_py_sq(s), __dsub_rn(3.0, __dmul_rn(2.0, s)));
        const double h11 = __dmul_rn(hb_py_sq(s), __dsub_rn(s, 1.0));
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const double dx0 = __ddiv_rn(__dsub_rn(x1[d], xm[d]), dtm), dx1 = __ddiv_rn(__dsub_rn(xp[d], x0[d]), dtp);
            xh[d] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(h00, x0[d]), __dmul_rn(__dmul_rn(h10, dx0), dt)),
                                        __dmul_rn(h01, x1[d])), __dmul_rn(__dmul_rn(h11, dx1), dt));
        }
    } else {
#pragma unroll
        for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(s, __dsub_rn(x1[d], x0[d])));
    }
}

// Per-segment logic of the cubic request, executed warp-uniformly (every lane the same segment k of samples T / X).
HB_DEV bool process_segment_cubic(const HitSink &p, Dedup &dd, long long traj, int lane, int k, int m, const double *T,
                                  const double *X, int newton_max_iter)
{
    const int dir = p.sec.direction, ci = p.sec.idx;
    const double off = p.sec.offset;
    const double t0 = T[k], t1 = T[k + 1], dt = __dsub_rn(t1, t0);
    const double gk = __dsub_rn(X[(long long)k * 6 + ci], off), gk1 = __dsub_rn(X[(long long)(k + 1) * 6 + ci], off);
    const bool has_prev = k - 1 >= 0;
    const double g_prev = has_prev ? __dsub_rn(X[(long long)(k - 1) * 6 + ci], off) : 0.0;
    const bool cubic = dt > 0.0;
    double d0 = 0.0, d1 = 0.0;
    if (cubic) {
        d0 = has_prev ? __ddiv_rn(__dsub_rn(gk1, g_prev), __dsub_rn(t1, T[k - 1])) : __ddiv_rn(__dsub_rn(gk1, gk), dt);
        d1 = (k + 2 < m) ? __ddiv_rn(__dsub_rn(__dsub_rn(X[(long long)(k + 2) * 6 + ci], off), gk), __dsub_rn(T[k + 2], t0))
                         : __ddiv_rn(__dsub_rn(gk1, gk), dt);
    }
    bool accept_left = false;
    if (fabs(gk) < p.sec.tol_on_surface) {
        if (dir == 0) accept_left = true;
        else if (dir > 0) accept_left = (gk1 >= 0.0) || (has_prev && g_prev <= 0.0);
        else accept_left = (gk1 <= 0.0) || (has_prev && g_prev >= 0.0);
    }
    const int r = p.sec.segment_refine;
    double xh[6];
    if (r > 0) {
        if (accept_left) {
#pragma unroll
            for (int d = 0; d < 6; ++d) xh[d] = X[(long long)k * 6 + d];
            if (!push_hit(p, dd, traj, t0, xh, lane)) return false;
        }
        const double step = __ddiv_rn(1.0, (double)(r + 1));
        for (int mm = 0; mm <= r; ++mm) {
            const double s_lo = __dmul_rn((double)mm, step), s_hi = __dmul_rn((double)(mm + 1), step);
            if (s_hi > 1.0 + 1e-15) break;
            if (accept_left && mm == 0) continue;
            double g_lo, g_hi;
            if (cubic) {
                g_lo = hermite_scalar(s_lo, gk, gk1, d0, d1, dt);
                g_hi = hermite_scalar(s_hi, gk, gk1, d0, d1, dt);
            } else {
                g_lo = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_lo), gk), __dmul_rn(s_lo, gk1));
                g_hi = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_hi), gk), __dmul_rn(s_hi, gk1));
            }
            bool crosses;
            if (dir == 0) crosses = (__dmul_rn(g_lo, g_hi) <= 0.0) && (g_lo != g_hi);
            else if (dir > 0) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
            else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
            if (!crosses) continue;
            double s_star;
            if (g_lo == g_hi) s_star = __dmul_rn(0.5, __dadd_rn(s_lo, s_hi));
            else {
                double al = __ddiv_rn(g_lo, __dsub_rn(g_lo, g_hi));
                al = fmin(1.0, fmax(0.0, al));
                s_star = __dadd_rn(s_lo, __dmul_rn(al, __dsub_rn(s_hi, s_lo)));
            }
            if (cubic) {
                for (int it = 0; it < newton_max_iter; ++it) {
                    const double f = hermite_scalar(s_star, gk, gk1, d0, d1, dt);
                    const double df = hermite_der(s_star, gk, gk1, d0, d1, dt);
                    if (df == 0.0) break;
                    s_star = __dsub_rn(s_star, __ddiv_rn(f, df));
                    if (s_star < s_lo) { s_star = s_lo; break; }
                    if (s_star > s_hi) { s_star = s_hi; break; }
                }
            }
            const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_star), t0), __dmul_rn(s_star, t1));
            cubic_hit_state(T, X, m, k, s_star, dt, xh);
            if (!push_hit(p, dd, traj, th, xh, lane)) return false;
        }
    } else {
        if (accept_left) {
#pragma unroll
            for (int d = 0; d < 6; ++d) xh[d] = X[(long long)k * 6 + d];
            return push_hit(p, dd, traj, t0, xh, lane);
        }
        bool crosses;
        if (dir == 0) crosses = (__dmul_rn(gk, gk1) <= 0.0) && (gk != gk1);
        else if (dir > 0) crosses = (gk < 0.0) && (gk1 >= 0.0);
        else crosses = (gk > 0.0) && (gk1 <= 0.0);
        if (!crosses) return true;
        double s_star = __ddiv_rn(gk, __dsub_rn(gk, gk1));
        s_star = fmin(1.0, fmax(0.0, s_star));
        if (cubic) {
            for (int it = 0; it < newton_max_iter; ++it) {
                const double f = hermite_scalar(s_star, gk, gk1, d0, d1, dt);
                const double df = hermite_der(s_star, gk, gk1, d0, d1, dt);
                if (df == 0.0) break;
                s_star = __dsub_rn(s_star, __ddiv_rn(f, df));
                if (s_star < 0.0) { s_star = 0.0; break; }
                if (s_star > 1.0) { s_star = 1.0; break; }
            }
        }
        const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_star), t0), __dmul_rn(s_star, t1));
        cubic_hit_state(T, X, m, k, s_star, dt, xh);
        return push_hit(p, dd, traj, th, xh, lane);
    }
    return true;
}


// Warp-cooperative form of process_segment for the fused section kernel: every lane calls it with the SAME
// (broadcast) segment values; the r+1 sub-intervals of _detect_with_segment_refine are tested 32 at a time
// instead of in a 51-iteration scalar loop, and only the owner lane (is_owner) materialises the two end
// states -- lazily, through `eval`, and only when the segment actually produces a candidate -- and pushes
// hits.  Arithmetic per sub-interval is unchanged, candidates are pushed in the reference's order.
//   eval(which, out): which = 0 -> state at the left sample, 1 -> at the right sample (owner lane only).
template <class EVAL>
HB_DEV bool process_segment_coop(const HitSink &p, Dedup &dd, long long traj, int lane, bool is_owner, bool has_prev,
                                 double g_prev, double gk, double gk1, double t0, double t1, EVAL eval)
{
    const int dir = p.sec.direction;
    bool accept_left = false;
    if (fabs(gk) < p.sec.tol_on_surface) {
        if (dir == 0) accept_left = true;
        else if (dir > 0) accept_left = (gk1 >= 0.0) || (has_prev && g_prev <= 0.0);
        else accept_left = (gk1 <= 0.0) || (has_prev && g_prev >= 0.0);
    }
    const int r = p.sec.segment_refine;
    bool alive = true, have_states = false;
    double x0[6], x1[6], xh[6];
    if (r > 0) {
        const double step = __ddiv_rn(1.0, (double)(r + 1));
        // highest sub-interval index that passes the reference's `s_hi > 1.0 + 1e-15 -> break` test
        if (accept_left && is_owner) {
            eval(0, x0); eval(1, x1); have_states = true;
            alive = push_hit(p, dd, traj, t0, x0, 0);
        }
        for (int base = 0; base <= r; base += 32) {
            const int mm = base + lane;
            bool crosses = false;
            double g_lo = 0.0, g_hi = 0.0, s_lo = 0.0, s_hi = 0.0;
            bool stop = false;
            if (mm <= r) {
                s_lo = __dmul_rn((double)mm, step);
                s_hi = __dmul_rn((double)(mm + 1), step);
                stop = s_hi > 1.0 + 1e-15;
                if (!stop && !(accept_left && mm == 0)) {
                    g_lo = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_lo), gk), __dmul_rn(s_lo, gk1));
                    g_hi = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_hi), gk), __dmul_rn(s_hi, gk1));
                    if (dir == 0) crosses = (__dmul_rn(g_lo, g_hi) <= 0.0) && (g_lo != g_hi);
                    else if (dir > 0) crosses = (g_lo < 0.0) && (g_hi >= 0.0);
                    else crosses = (g_lo > 0.0) && (g_hi <= 0.0);
                }
            }
            // the reference BREAKS at the first sub-interval with s_hi > 1 + 1e-15: ignore everything after it
            const unsigned stopm = __ballot_sync(0xffffffffu, stop);
            unsigned cm = __ballot_sync(0xffffffffu, crosses);
            if (stopm) cm &= (1u << (__ffs(stopm) - 1)) - 1u;
            while (cm) {
                const int i = __ffs(cm) - 1;
                cm &= cm - 1;
                const double a_lo = shfl_d(g_lo, i), a_hi = shfl_d(g_hi, i);
                const double b_lo = shfl_d(s_lo, i), b_hi = shfl_d(s_hi, i);
                if (is_owner && alive) {
                    double s_star;
                    if (a_lo == a_hi) s_star = __dmul_rn(0.5, __dadd_rn(b_lo, b_hi));
                    else {
                        double al = __ddiv_rn(a_lo, __dsub_rn(a_lo, a_hi));
                        al = fmin(1.0, fmax(0.0, al));
                        s_star = __dadd_rn(b_lo, __dmul_rn(al, __dsub_rn(b_hi, b_lo)));
                    }
                    if (!have_states) { eval(0, x0); eval(1, x1); have_states = true; }
                    const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, s_star), t0), __dmul_rn(s_star, t1));
#pragma unroll
                    for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(s_star, __dsub_rn(x1[d], x0[d])));
                    alive = push_hit(p, dd, traj, th, xh, 0);
                }
            }
            if (stopm) break;
        }
    } else if (is_owner) {
        if (accept_left) {
            eval(0, x0);
            alive = push_hit(p, dd, traj, t0, x0, 0);
        } else {
            bool crosses;
            if (dir == 0) crosses = (__dmul_rn(gk, gk1) <= 0.0) && (gk != gk1);
            else if (dir > 0) crosses = (gk < 0.0) && (gk1 >= 0.0);
            else crosses = (gk > 0.0) && (gk1 <= 0.0);
            if (crosses) {
                double al = __ddiv_rn(gk, __dsub_rn(gk, gk1));
                al = fmin(1.0, fmax(0.0, al));
                eval(0, x0); eval(1, x1);
                const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), t0), __dmul_rn(al, t1));
#pragma unroll
                for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(al, __dsub_rn(x1[d], x0[d])));
                alive = push_hit(p, dd, traj, th, xh, 0);
            }
        }
    }
    return alive;
}
