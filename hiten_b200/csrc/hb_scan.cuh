// hb_scan.cuh -- pieces shared by the two section pipelines that scan step records:
//   hb_section_scan.cu    (hb_cr3bp_section2: records go through an HBM scratch, scan kernel one warp per trajectory)
//   hb_section_stream.cu  (hb_cr3bp_section3: records go through a shared-memory hand-off inside ONE kernel,
//                          producer warps propagate, consumer warps scan)
// Parameters of the scan stages, the event-component evaluation, grid lookups, the per-segment candidate logic of
// _detect_with_segment_refine (algorithms/poincare/synodic/backend.py:458-659) and the segment descriptors.
#pragma once
#include "hb_cr3bp_common.cuh"

namespace hbscan {
using namespace hbc;

constexpr int HB_CAND_CAP = 32;       // candidate hits per trajectory (before de-duplication)
constexpr int HB_CAND_DOUBLES = 8;    // key, t, state[6]
constexpr int HB_DESC_DOUBLES = 8;    // traj, cs, record of cs-1, record of cs, g(cs-1), g(cs), g(cs-2), pad

// per-trajectory part of the hb_cr3bp_section2 scratch besides the step records, in doubles: candidate and segment
// lists, four counters (candidates, segments, records, pad: 2 doubles), the compact segment index (HB_CAND_CAP ints)
constexpr long long HB_S2_FIXED_DOUBLES = HB_CAND_CAP * (HB_CAND_DOUBLES + HB_DESC_DOUBLES) + HB_CAND_CAP / 2 + 2;

struct ScanParams {
    PropParams prop;        // mu, 1-mu, sign mask (vector field of the extra stages)
    long long n;
    const double *rec;      // step records, HB_REC_DOUBLES doubles each: [n][rec_cap] (section2) or a pool (section3)
    int rec_cap;            // records per trajectory in the scratch (section2); INT_MAX: no per-trajectory limit
    const int *nacc;
    const int *nrec;        // sparse records (hb_cr3bp_section2, records = near): records per trajectory; else NULL
    int *status;
    const double *t_eval;
    int m;
    double tsign, inv_grid_dt;
    HitSink sink;
    int *hits_per_traj;
    int *cand_count;        // [n]
    double *cand;           // [n][HB_CAND_CAP][HB_CAND_DOUBLES]
    int *desc_count;        // [n]   segments that can hold a hit, found by k_step_scan
    double *desc;           // [n][HB_CAND_CAP][HB_DESC_DOUBLES]
    int *desc_total;        // [2]   entries at the front / at the back of the compact index below
    int *desc_index;        // [n * HB_CAND_CAP] positions in desc of all noted segments (k_compact_segments)
};

template <class AR>
HB_DEV double g_comp(const double *hdr, double xq, double offset)     // hdr = record header
{
    const double hseg = hdr[2];
    double ge = hdr[3];
    if (hseg != 0.0) {
        const double omx = AR::sub(1.0, xq);
        double v = 0.0;
#pragma unroll
        for (int i = 6; i >= 0; --i) {
            v = AR::add(v, hdr[4 + i]);
            v = AR::mul(v, ((6 - i) % 2 == 0) ? xq : omx);
        }
        ge = AR::add(v, hdr[3]);
    }
    return __dsub_rn(ge, offset);
}
template <class AR>
HB_DEV double xpar(double tq, double t, double hseg) { return (hseg == 0.0) ? 0.0 : AR::div(AR::sub(tq, t), hseg); }
// the same quotient with the reciprocal of hseg prepared once (AR::rcp): used where many samples share a step
template <class AR>
HB_DEV double xpar_by(double tq, double t, double hseg, double inv)
{
    return (hseg == 0.0) ? 0.0 : AR::div_by(AR::sub(tq, t), hseg, inv);
}

// y_old, y_new and the stage rows the dense output uses (k[1..4] do not enter it; k[0] = f(y_old) is recomputed)
template <class AR>
HB_DEV void load_record(const double *r, const PropParams &pp, double &t_old, double &t_new, double (&y)[6],
                        double (&yn)[6], double (&k)[13][6])
{
    double v[HB_REC_DOUBLES];
#pragma unroll
    for (int i = 0; i < HB_REC_DOUBLES; i += 4) hb_ld4(r + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    t_old = v[0]; t_new = v[1];
#pragma unroll
    for (int d = 0; d < 6; ++d) {
        y[d] = v[HB_REC_YOLD + d];
        yn[d] = v[HB_REC_YNEW + d];
#pragma unroll
        for (int j = 1; j < 5; ++j) k[j][d] = 0.0;
#pragma unroll
        for (int j = 5; j < 13; ++j) k[j][d] = v[HB_REC_K5 + 6 * (j - 5) + d];
    }
    crtbp_rhs<AR, 2>(y, pp, k[0]);
}

template <class AR>
__device__ __noinline__ void states_from_record(const double *r, const PropParams &pp, double tq0, double tq1,
                                                double (&out0)[6], double (&out1)[6])
{
    double t, t_new, y[6], yn[6], k[13][6], F[7][6];
    load_record<AR>(r, pp, t, t_new, y, yn, k);
    const double hseg = AR::sub(t_new, t);
    if (hseg == 0.0) {
#pragma unroll
        for (int d = 0; d < 6; ++d) { out0[d] = y[d]; out1[d] = y[d]; }
        return;
    }
    const Cr3bpRhs<AR, 2> rhs{pp};
    dense_cache<AR>(y, yn, hseg, k, F, rhs);
    dense_eval<AR>(y, F, xpar<AR>(tq0, t, hseg), out0);
    dense_eval<AR>(y, F, xpar<AR>(tq1, t, hseg), out1);
}

// first index c in [lo, m] with t_eval[c] >= tv  (guess from the uniform spacing, then fix up)
HB_DEV int first_at_or_after(const ScanParams &p, double tv, int lo)
{
    int c = (int)fmin(fmax((tv - p.t_eval[0]) * p.inv_grid_dt, (double)lo), (double)p.m);
    if (c > lo && c < p.m) {                                  // the guess is usually exact: settle it with two
        const double a = p.t_eval[c - 1], b = p.t_eval[c];    // independent loads instead of two dependent rounds
        if (a < tv && !(b < tv)) return c;
    }
    while (c < p.m && p.t_eval[c] < tv) ++c;
    while (c > lo && !(p.t_eval[c - 1] < tv)) --c;
    return c;
}

// The same lookup for the scan kernel, returning the grid values around the answer as well: t_eval[c], t_eval[c - 1]
// and t_eval[c - 2] are fetched in ONE round of independent loads (the grid does not stay in the little L1 left beside
// the staging rows, so every dependent lookup is a round trip to L2).  te0 = t_eval[0], loaded once per warp.
HB_DEV int first_at_or_after3(const ScanParams &p, double te0, double tv, double &tc, double &tm1, double &tm2)
{
    int c = (int)fmin(fmax((tv - te0) * p.inv_grid_dt, 0.0), (double)p.m);
    if (c > 0 && c < p.m) {
        const double a = p.t_eval[c - 1], b = p.t_eval[c], z = p.t_eval[max(c - 2, 0)];
        if (a < tv && !(b < tv)) {
            tc = b; tm1 = a; tm2 = z;
            return c;
        }
    }
    while (c < p.m && p.t_eval[c] < tv) ++c;
    while (c > 0 && !(p.t_eval[c - 1] < tv)) --c;
    tc = p.t_eval[min(c, p.m - 1)];
    tm1 = p.t_eval[max(c - 1, 0)];
    tm2 = p.t_eval[max(c - 2, 0)];
    return c;
}

// _detect_with_segment_refine on ONE segment (linear branch), emitting raw candidates in order
template <class EMIT>
HB_DEV void segment_candidates(const hb_section &sec, bool has_prev, double g_prev, double gk, double gk1, double t0,
                               double t1, const double (&x0)[6], const double (&x1)[6], EMIT emit)
{
    const int dir = sec.direction;
    bool accept_left = false;
    if (fabs(gk) < sec.tol_on_surface) {
        if (dir == 0) accept_left = true;
        else if (dir > 0) accept_left = (gk1 >= 0.0) || (has_prev && g_prev <= 0.0);
        else accept_left = (gk1 <= 0.0) || (has_prev && g_prev >= 0.0);
    }
    const int r = sec.segment_refine;
    double xh[6];
    int order = 0;
    if (r > 0) {
        if (accept_left) emit(order++, t0, x0);
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
            emit(order++, th, xh);
        }
    } else {
        if (accept_left) { emit(order++, t0, x0); return; }
        bool crosses;
        if (dir == 0) crosses = (__dmul_rn(gk, gk1) <= 0.0) && (gk != gk1);
        else if (dir > 0) crosses = (gk < 0.0) && (gk1 >= 0.0);
        else crosses = (gk > 0.0) && (gk1 <= 0.0);
        if (!crosses) return;
        double al = __ddiv_rn(gk, __dsub_rn(gk, gk1));
        al = fmin(1.0, fmax(0.0, al));
        const double th = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), t0), __dmul_rn(al, t1));
#pragma unroll
        for (int d = 0; d < 6; ++d) xh[d] = __dadd_rn(x0[d], __dmul_rn(al, __dsub_rn(x1[d], x0[d])));
        emit(order++, th, xh);
    }
}

// The scan stage only NOTES the segments that can hold a hit (8 numbers each, per-trajectory list); k_emit_candidates
// turns them into candidate hits.  Keeping the state reconstruction out of the scan keeps it small (no spills, no call).
// r0 / r1: positions (in records of HB_REC_DOUBLES doubles, from ScanParams::rec) of the step records that hold the
// segment's left / right sample -- `traj * rec_cap + step` in the HBM-scratch pipeline, a pool slot in the streamed one.
HB_DEV void store_segment(const ScanParams &p, long long traj, int slot, int cs, long long r0, long long r1, double gk,
                          double gk1, double gm2)
{
    if (slot >= HB_CAND_CAP) return;                          // counted, reported as overflow by k_order_dedup
    double *d = p.desc + (traj * HB_CAND_CAP + slot) * HB_DESC_DOUBLES;
    hb_st4(d, (double)traj, (double)cs, (double)r0, (double)r1);
    hb_st4(d + 4, gk, gk1, gm2, 0.0);
}

HB_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
HB_DEV void mbar_wait(unsigned mbar, unsigned parity)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}

// Can the segment between two neighbouring grid samples (event values gl, gr) produce a hit in process_segment()?
// A left node within the on-surface tolerance may; otherwise a directed section needs the segment itself to cross in
// that direction: the 51 sub-interval values are rounded points of the straight line from gl to gr, which rises or
// falls by |gr - gl| / 51 >= max(|gl|, |gr|) / 51 per sub-interval -- far above their rounding error -- so a segment
// running the other way has no sub-interval with (g_lo > 0, g_hi <= 0).  Segments dropped here are the upward
// crossings of a direction = -1 section: half of all flagged segments of a tube.
HB_DEV bool segment_may_hit(int dir, double gl, double gr, double tol)
{
    if (fabs(gl) < tol) return true;
    if (dir < 0) return gl > 0.0 && gr <= 0.0;
    if (dir > 0) return gl < 0.0 && gr >= 0.0;
    return !((gl > 0.0 && gr > 0.0) || (gl < 0.0 && gr < 0.0));
}


// B2 + B3 of either pipeline (defined in hb_section_scan.cu): compact index over the noted segments, candidate
// emission from the step records the descriptors point to, ordering + de-duplication.  `ev_emit_done` (optional) is
// recorded between B2 and B3.
int hb_scan_finish(const ScanParams &p, int arith, cudaStream_t st, cudaEvent_t ev_emit_done);

}  // namespace hbscan
