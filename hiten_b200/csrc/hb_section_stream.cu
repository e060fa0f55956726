// hb_section_stream.cu -- tube + synodic section with the step records handed over INSIDE the SM (hb_cr3bp_section3).
//
// hb_cr3bp_section2 (hb_section_scan.cu) writes what the dense output of every accepted step depends on (512 B) to an
// HBM scratch and reads it back in a second kernel: 44 GB written + 44 GB read per 1e6 trajectories against ~0.3 GB of
// algorithmic traffic, an 82 GB scratch, and a per-trajectory step capacity.  Here the same two bodies run in ONE
// persistent kernel as specialised warps of one CTA per SM:
//
//   PRODUCER warps (6): the DOP853 propagation loop of k_dop853_6 (hb_cr3bp.cu), one trajectory per thread, persistent
//     work queue.  After every attempted step a warp publishes a BATCH: the lanes whose step was accepted have written
//     their step record {t_old, t_new, y_old, y_new, k6..k13, trajectory, step number, flags} into their own row of
//     shared memory; an mbarrier (`full`) hands the batch to the consumer.
//   CONSUMER warps (2, three producers each, fixed round robin): lane j takes producer lane j's record and runs the body
//     of k_step_scan on it -- three extra DOP853 stages, the EVENT COMPONENT of the step's interpolant, the grid samples
//     the step owns, the quiet-step test, the warp-cooperative scan of the non-quiet steps -- with the per-trajectory
//     carries (first sample not owned yet, event values at the last two owned samples) kept per (producer, lane) in
//     shared memory instead of travelling along a warp.  A second mbarrier (`empty`) returns the rows.
//
// Only the records of steps that hold a NOTED segment (a few per trajectory) leave the SM: the consumer copies them
// into a record pool in HBM (warp-cooperative 512-byte copies, slots from a per-warp chunk of a global counter) and
// stores the segment descriptor; the rest of the pipeline -- k_compact_segments, k_emit_candidates, k_order_dedup -- is
// the one of hb_cr3bp_section2 (hb_scan_finish), reading records from the pool.  Same arithmetic per record, sample
// and segment as hb_cr3bp_section2, hence the same hits bit for bit (tests/test_gpu_synodic.py, tests/test_gpu_c5.py).
//
// Row retention: a segment that straddles two steps needs the record of the step that owns its left sample, which may
// be several accepted steps back (steps shorter than the grid spacing own no sample).  Every lane has TWO rows; the
// consumer publishes which of them holds that "owner" record (pin mask) and the producer writes the next record into
// the other one.  One batch per producer warp is in flight: the producer computes its next step while the consumer
// works on the batch, and waits for `empty` only before it writes rows again.
//
// Reference: algorithms/integrators/rk.py:2377-2549 (propagation), algorithms/poincare/synodic/backend.py:458-659
// (_detect_with_segment_refine), :382-455 (_order_and_dedup_hits).
#include <limits.h>

#include "hb_scan.cuh"

namespace {
using namespace hbc;
using namespace hbscan;

#ifndef HB_PC_NP
#define HB_PC_NP 4
#endif
#ifndef HB_PC_NC
#define HB_PC_NC 4
#endif
constexpr int PC_NP = HB_PC_NP;                  // producer warps per CTA
constexpr int PC_NC = HB_PC_NC;                  // consumer warps per CTA
constexpr int PC_PER_C = PC_NP / PC_NC;          // producers served by one consumer warp
constexpr int PC_THREADS = 32 * (PC_NP + PC_NC);
constexpr int PC_ROW = HB_REC_DOUBLES * 8 + 16;  // 528 B = 33 x 16 B: conflict-free 16-byte accesses across lanes
constexpr int PC_PLANE = 32 * PC_ROW;            // one row per lane
constexpr int PC_WARP_ROWS = 2 * PC_PLANE;       // two planes per producer warp
constexpr int PC_POOL_CHUNK = 128;               // record-pool slots a consumer warp takes per atomic
static_assert(PC_NP % PC_NC == 0, "every consumer warp serves the same number of producer warps");
static_assert((PC_ROW / 16) % 2 == 1, "odd number of 16-byte units per row");

enum { PC_FLAG_LAST = 1 };                       // record flag: last accepted step of the trajectory (t_new == tf)

// per producer warp: hand-off state
struct PcCtrl {
    unsigned long long full, empty;              // mbarriers (32 arrivals each)
    unsigned amask;                              // lanes with a record in this batch
    unsigned pin;                                // per lane: plane (0/1) that holds the owner record -- do not write
    unsigned exit;                               // producer warp is done
    unsigned pad[3];
};
// per (producer warp, lane): scan carries of the trajectory that lane is propagating (k_step_scan's warp-uniform
// carries, one set per trajectory)
struct PcCarry {
    double t_c;            // t_eval[c]
    double g1, g2;         // event function at samples c - 1, c - 2
    long long own_rec;     // pool slot of the record that owns sample c - 1 (-1: not copied yet)
    int c;                 // first grid sample not owned yet
    int ndesc;             // segments noted so far
    int pin;               // plane of the owner record
    int pad;
    long long pad2;
};
constexpr int PC_SMEM = PC_NP * PC_WARP_ROWS + PC_NP * (int)sizeof(PcCtrl) + PC_NP * 32 * (int)sizeof(PcCarry);
static_assert(PC_SMEM <= 227 * 1024, "shared memory budget of one CTA");

struct StreamParams {
    PropParams prop;
    ScanParams scan;
    double *pool;                   // [pool_cap][HB_REC_DOUBLES]
    long long pool_cap;
    unsigned long long *pool_count;
};

HB_DEV void mbar_arrive(unsigned mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------
// consumer: one batch of one producer warp
// ---------------------------------------------------------------------------------------------------------------
template <class AR, int C>
HB_DEV void consume_batch(const StreamParams &sp, const unsigned char *rows, PcCtrl *ctl, PcCarry *carry, int lane,
                          double te0, long long &pool_next, long long &pool_end)
{
    constexpr unsigned FULL = 0xffffffffu;
    const ScanParams &p = sp.scan;
    const unsigned amask = ctl->amask;
    const bool have_rec = (amask >> lane) & 1u;
    PcCarry cy = carry[lane];
    const double off = p.sink.sec.offset, tol_s = p.sink.sec.tol_on_surface;
    const int sdir = p.sink.sec.direction;
    const int plane = 1 - cy.pin;                                  // where the producer put this batch's record
    const unsigned char *rowb = rows + plane * PC_PLANE + lane * PC_ROW;
    const double *r = (const double *)rowb;
    double hdr[11];
#pragma unroll
    for (int i = 0; i < 11; ++i) hdr[i] = 0.0;
    int cend = cy.c, flags = 0;
    long long traj = -1;
    double t_c = cy.t_c, t_cm1 = 0.0, t_cm2 = 0.0;
    if (have_rec) {
        double v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            const double2 w = *(const double2 *)(r + i);
            v[i] = w.x; v[i + 1] = w.y;
        }
        traj = *(const long long *)(r + 62);
        const int2 meta = *(const int2 *)(r + 63);
        flags = meta.y;
        if (meta.x == 1) {                                         // first accepted step of a new trajectory
            cy.c = 0; cy.t_c = te0; cy.g1 = 0.0; cy.g2 = 0.0; cy.ndesc = 0; cy.own_rec = -1;
        }
        const double y[6] = {v[2], v[3], v[4], v[5], v[6], v[7]}, yn[6] = {v[8], v[9], v[10], v[11], v[12], v[13]};
        const double hseg = AR::sub(v[1], v[0]);
        hdr[0] = v[0]; hdr[1] = v[1]; hdr[2] = hseg; hdr[3] = y[C];
        if (hseg != 0.0) {
            const Cr3bpRhs<AR, 2> rhs{sp.prop};
            auto row = [&](int R, double (&kr)[6]) {
                const double2 *q = (const double2 *)(r + HB_REC_K5 + 6 * (R - 5));
                const double2 a = q[0], b = q[1], cc = q[2];
                kr[0] = a.x; kr[1] = a.y; kr[2] = b.x; kr[3] = b.y; kr[4] = cc.x; kr[5] = cc.y;
            };
            auto pick = [&](const double (&w)[6]) { return w[C]; };
            double f[7];
            dense_component<AR>(y, yn, hseg, row, pick, rhs, f);
#pragma unroll
            for (int i = 0; i < 7; ++i) hdr[4 + i] = f[i];
        }
        if (!(flags & PC_FLAG_LAST)) cend = first_at_or_after3(p, te0, v[1], t_c, t_cm1, t_cm2);
        else { cend = p.m; t_cm1 = p.t_eval[p.m - 1]; t_cm2 = p.t_eval[max(p.m - 2, 0)]; }
    }
    const int c0 = cy.c;
    const double t_c0 = cy.t_c;
    const int nown = (have_rec && c0 < cend) ? cend - c0 : 0;
    const bool owns = nown > 0;
    double g_first = 0.0, A = 0.0, B = 0.0;
    const double inv_h = (hdr[2] != 0.0) ? AR::rcp(hdr[2]) : 0.0;
    if (owns) {
        g_first = g_comp<AR>(hdr, xpar_by<AR>(t_c0, hdr[0], hdr[2], inv_h), off);
        A = g_first;
        if (nown >= 2) {
            A = g_comp<AR>(hdr, xpar_by<AR>(t_cm1, hdr[0], hdr[2], inv_h), off);
            B = (nown >= 3) ? g_comp<AR>(hdr, xpar_by<AR>(t_cm2, hdr[0], hdr[2], inv_h), off) : g_first;
        }
    }
    const double g_prev = cy.g1, gm2 = cy.g2;
    long long cur_rec = -1;                                        // pool slot of this batch's record of this lane
    int ndesc = cy.ndesc;

    // pool slots for the lanes in `need` (warp-uniform mask); -1 when the pool is exhausted
    auto alloc = [&](unsigned need) -> long long {
        const int cnt = __popc(need);
        if (pool_next + cnt > pool_end) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(sp.pool_count, (unsigned long long)PC_POOL_CHUNK);
            pool_next = (long long)__shfl_sync(FULL, base, 0);
            pool_end = pool_next + PC_POOL_CHUNK;
        }
        const long long mine = pool_next + __popc(need & ((1u << lane) - 1u));
        pool_next += cnt;
        return (mine < sp.pool_cap) ? mine : -1;
    };
    // the whole warp copies one 512-byte row (16 bytes per lane) into a pool slot
    auto copy_row = [&](int src_lane, int src_plane, long long slot) {
        if (slot < 0) return;
        const double2 w = *(const double2 *)(rows + src_plane * PC_PLANE + src_lane * PC_ROW + 16 * lane);
        double *dst = sp.pool + slot * HB_REC_DOUBLES + 2 * lane;
        asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst), "d"(w.x), "d"(w.y) : "memory");
    };

    bool scan = false;
    {   // the segment that ends at this step's first sample: its left sample belongs to the owner record
        const bool flag = owns && c0 > 0 && segment_may_hit(sdir, g_prev, g_first, tol_s);
        unsigned fm = __ballot_sync(FULL, flag);
        if (fm) {
            const unsigned need_own = __ballot_sync(FULL, flag && cy.own_rec < 0);
            if (need_own) {
                const long long s = alloc(need_own);
                if (flag && cy.own_rec < 0) cy.own_rec = (s < 0) ? -2 : s;
            }
            const long long s_cur = alloc(fm);
            if (flag) cur_rec = (s_cur < 0) ? -2 : s_cur;
            unsigned todo = fm;
            while (todo) {
                const int L = __ffs(todo) - 1;
                todo &= todo - 1;
                const int pinL = __shfl_sync(FULL, cy.pin, L);
                const long long ownL = __shfl_sync(FULL, cy.own_rec, L), curL = __shfl_sync(FULL, cur_rec, L);
                if ((need_own >> L) & 1u) copy_row(L, pinL, ownL);
                copy_row(L, 1 - pinL, curL);
            }
            if (flag) {
                if (cy.own_rec < 0 || cur_rec < 0) ndesc = HB_CAND_CAP + 1;          // pool exhausted: rerun this trajectory
                else {
                    store_segment(p, traj, ndesc, c0, cy.own_rec, cur_rec, g_prev, g_first, c0 > 1 ? gm2 : 0.0);
                    ++ndesc;
                }
            }
        }
    }
    if (owns) {
        scan = nown >= 2;
        // quiet step: no sample can be on the surface or change sign (|p(x) - (y0 + x F0)| <= sum_{i>=1}|F_i| / 4)
        if (scan && hdr[2] != 0.0) {
            const double g_old = __dsub_rn(hdr[3], off);
            const double g_new = g_comp<AR>(hdr, 1.0, off);
            double S = 0.0;
#pragma unroll
            for (int i = 1; i < 7; ++i) S += fabs(hdr[4 + i]);
            const double margin = 0.25 * S + tol_s + 1e-9 * (fabs(hdr[3]) + fabs(hdr[4]) + fabs(off)) + 1e-290;
            const bool same = (g_old > 0.0 && g_new > 0.0) || (g_old < 0.0 && g_new < 0.0);
            if (same && fmin(fabs(g_old), fabs(g_new)) > margin) scan = false;
        }
    }
    // cooperative scan of the non-quiet steps of this batch, one owner lane at a time (each lane = another trajectory)
    unsigned req = __ballot_sync(FULL, scan);
    while (req) {
        const int L = __ffs(req) - 1;
        req &= req - 1;
        double bh[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) bh[i] = shfl_d(hdr[i], L);
        const int b0 = __shfl_sync(FULL, c0, L), b1 = __shfl_sync(FULL, cend, L);
        double c1 = shfl_d(g_first, L), c2 = shfl_d(g_prev, L);
        const double binv = shfl_d(inv_h, L);
        const long long trajL = __shfl_sync(FULL, traj, L);
        const int planeL = 1 - __shfl_sync(FULL, cy.pin, L);
        long long recL = __shfl_sync(FULL, cur_rec, L);
        int nd = __shfl_sync(FULL, ndesc, L);
        double tq_next = p.t_eval[min(b0 + 1 + lane, b1 - 1)];
        for (int b = b0 + 1; b < b1; b += 32) {
            const int cs = b + lane;
            const bool valid = cs < b1;
            const double tq = tq_next;
            if (b + 32 < b1) tq_next = p.t_eval[min(cs + 32, b1 - 1)];
            const double g = g_comp<AR>(bh, xpar_by<AR>(tq, bh[0], bh[2], binv), off);
            double g_m1 = __shfl_up_sync(FULL, g, 1);
            double g_m2 = __shfl_up_sync(FULL, g, 2);
            if (lane == 0) { g_m1 = c1; g_m2 = c2; }
            if (lane == 1) g_m2 = c1;
            const bool flagged = valid && segment_may_hit(sdir, g_m1, g, tol_s);
            const unsigned fm = __ballot_sync(FULL, flagged);
            if (fm) {
                if (recL == -1) {                                  // first noted segment of this step: copy its record
                    recL = alloc(1u << L);
                    recL = __shfl_sync(FULL, recL, L);
                    if (recL < 0) recL = -2;
                    copy_row(L, planeL, recL);
                }
                if (recL < 0) nd = HB_CAND_CAP + 1;
                else {
                    if (flagged) store_segment(p, trajL, nd + __popc(fm & ((1u << lane) - 1u)), cs, recL, recL, g_m1, g, g_m2);
                    nd += __popc(fm);
                }
            }
            const int nvalid = min(32, b1 - b);
            const double l1 = shfl_d(g, nvalid - 1);
            const double l2 = shfl_d(g, nvalid >= 2 ? nvalid - 2 : 0);
            c2 = (nvalid >= 2) ? l2 : c1;
            c1 = l1;
        }
        if (lane == L) { ndesc = nd; cur_rec = recL; }
    }
    // carries of this lane's trajectory
    if (have_rec) {
        if (owns) {
            cy.g2 = (nown >= 2) ? B : cy.g1;
            cy.g1 = A;
            cy.own_rec = cur_rec;
            cy.pin = plane;                       // this record now owns the last sample: keep its row
        }
        cy.c = cend;
        cy.t_c = t_c;
        cy.ndesc = ndesc;
        // a trajectory that fails (max attempts, non-finite) never sends a last record: its list stays closed at 0
        if (flags & PC_FLAG_LAST) p.desc_count[traj] = ndesc;
    }
    carry[lane] = cy;
    const unsigned pin = __ballot_sync(FULL, cy.pin != 0);
    if (lane == 0) ctl->pin = pin;
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
template <class AR, int NEG, int C>
__global__ void __launch_bounds__(PC_THREADS, 1) k_tube_section_pc(const StreamParams sp)
{
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char pc_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    PcCtrl *ctl_all = (PcCtrl *)(pc_smem + PC_NP * PC_WARP_ROWS);
    PcCarry *carry_all = (PcCarry *)(pc_smem + PC_NP * PC_WARP_ROWS + PC_NP * sizeof(PcCtrl));
    if (threadIdx.x < PC_NP) {
        PcCtrl *c = ctl_all + threadIdx.x;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&c->full)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(smem_u32(&c->empty)) : "memory");
        c->amask = 0; c->pin = 0; c->exit = 0;
    }
    for (int i = threadIdx.x; i < PC_NP * 32; i += PC_THREADS) {
        PcCarry z{};
        z.own_rec = -1;
        carry_all[i] = z;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (wid >= PC_NP) {
        // ------------------------------------------------------------------ consumer warp
        const int cw = wid - PC_NP;
        const double te0 = sp.scan.t_eval[0];
        unsigned alive = (1u << PC_PER_C) - 1u, phase = 0;          // bit k: producer cw + k * PC_NC still running
        long long pool_next = 0, pool_end = 0;
        while (alive) {
            for (int k = 0; k < PC_PER_C; ++k) {
                if (!((alive >> k) & 1u)) continue;
                const int w = cw + k * PC_NC;
                PcCtrl *ctl = ctl_all + w;
                mbar_wait(smem_u32(&ctl->full), (phase >> k) & 1u);
                phase ^= 1u << k;
                if (ctl->exit) { alive &= ~(1u << k); continue; }
#ifndef HB_PC_NOCONSUME
                if (ctl->amask)
#else
                if (ctl->amask == 0xdeadbeefu)
#endif
                    consume_batch<AR, C>(sp, pc_smem + w * PC_WARP_ROWS, ctl, carry_all + w * 32, lane, te0, pool_next,
                                         pool_end);
                mbar_arrive(smem_u32(&ctl->empty));
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- producer warp
    const PropParams &p = sp.prop;
    PcCtrl *ctl = ctl_all + wid;
    unsigned char *rows = pc_smem + wid * PC_WARP_ROWS;
    const unsigned full_bar = smem_u32(&ctl->full), empty_bar = smem_u32(&ctl->empty);
    unsigned empty_phase = 1;                    // the first wait on a fresh barrier passes
    double y[6], yh[6], k[13][6];
    const Cr3bpRhs<AR, NEG> rhs{p};
    double t = 0.0, h = 0.0, err_prev = -1.0;
    const double tf = p.tf;
    long long idx = -1;
    long long attempts = 0;
    int nacc = 0, nrej = 0;
    bool have = false, exhausted = false;
    for (;;) {
        if (!have && !exhausted) {
            idx = hb_fetch_index(p.ws);
            if (idx < p.n) {
#pragma unroll
                for (int d = 0; d < 6; ++d) y[d] = p.y0[(long long)d * p.n + idx];
                crtbp_rhs<AR, NEG>(y, p, k[0]);
                t = p.t0;
                h = p.h0 ? p.h0[idx] : initial_step<AR>(y, k[0], p);
                err_prev = -1.0;
                nacc = 0; nrej = 0; attempts = 0;
                have = true;
                if (!((t - tf) < 0.0)) {         // zero-length span: nothing to integrate, no samples beyond the first
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                    p.nacc[idx] = 0; p.nrej[idx] = 0; p.status[idx] = HB_TRAJ_OK;
                    have = false;
                }
            } else {
                exhausted = true;
            }
        }
        if (__all_sync(FULL, !have && exhausted)) break;

        bool accepted = false, last = false;
        double t_new = 0.0;
        if (have) {
            // ---- one attempted step (rk.py:2452-2484) ----
            h = hb_clamp_step(h, p.max_step, p.min_step);
            if (AR::add(t, h) > tf) h = fabs(AR::sub(tf, t));
            dop853_stages<AR>(y, k, h, yh, rhs);
            double n5 = 0.0, n3 = 0.0;
            dop853_err_sums<AR>(y, yh, k, h, p.rtol, p.atol, n5, n3);
            const double err = dop853_err_norm<AR>(n5, n3, h, 6.0);
            ++attempts;
            const double h_factor = hb_pi_factor<AR>(err, err_prev, err <= 1.0, 8.0);
            accepted = err <= 1.0;
            if (accepted) {
                t_new = AR::add(t, h);
                ++nacc;
                last = !((t_new - tf) < 0.0);
            } else {
                ++nrej;
            }
            // (state update happens after the record has been written)
            if (!accepted) {
                h = AR::mul(h, h_factor);
                h = hb_clamp_step(h, p.max_step, p.min_step);
                int fin = -1;
                if (!(h == h) || !(err == err)) fin = HB_TRAJ_NONFINITE;
                else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
                if (fin >= 0) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                    p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin;
                    have = false;
                }
            } else {
                err_prev = err;
                h = AR::mul(h, h_factor);         // (applied to the NEW step below; h of this step is t_new - t)
            }
        }
        // ---- publish this iteration's batch ----
        const unsigned amask = __ballot_sync(FULL, accepted);
        mbar_wait(empty_bar, empty_phase);       // the consumer is done with the previous batch (rows + pin mask)
        empty_phase ^= 1u;
        if (accepted) {
            const unsigned pin = ctl->pin;
            double *rr = (double *)(rows + (1 - (int)((pin >> lane) & 1u)) * PC_PLANE + lane * PC_ROW);
            double2 *q = (double2 *)rr;
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
            *(long long *)(rr + 62) = idx;
            *(int2 *)(rr + 63) = make_int2(nacc, last ? PC_FLAG_LAST : 0);
        }
        if (lane == 0) ctl->amask = amask;
        mbar_arrive(full_bar);

        if (accepted) {
            int fin = -1;
            if (last) {                          // the dense interpolant at tf on the last segment
                const double hseg = AR::sub(t_new, t);
                double yo[6];
                if (hseg == 0.0) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) yo[d] = y[d];
                } else {
                    const double x = AR::div(AR::sub(tf, t), hseg);
                    if (x == 1.0) {
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
                fin = HB_TRAJ_OK;
            }
            t = t_new;
#pragma unroll
            for (int d = 0; d < 6; ++d) { y[d] = yh[d]; k[0][d] = k[12][d]; }
            if (fin < 0) {
                if (!(h == h)) fin = HB_TRAJ_NONFINITE;
                else if (attempts >= p.max_attempts) fin = HB_TRAJ_MAXSTEPS;
                if (fin >= 0) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) p.yf[(long long)d * p.n + idx] = y[d];
                }
            }
            if (fin >= 0) {
                p.nacc[idx] = nacc; p.nrej[idx] = nrej; p.status[idx] = fin;
                have = false;
            }
        }
    }
    // tell the consumer this warp is done
    mbar_wait(empty_bar, empty_phase);
    if (lane == 0) { ctl->amask = 0; ctl->exit = 1; }
    mbar_arrive(full_bar);
}

template <class AR, int NEG, int C>
int launch_pc3(const StreamParams &sp, unsigned grid, cudaStream_t st)
{
    HB_CUDA_TRY(cudaFuncSetAttribute(k_tube_section_pc<AR, NEG, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM));
    k_tube_section_pc<AR, NEG, C><<<grid, PC_THREADS, PC_SMEM, st>>>(sp);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
template <class AR, int NEG>
int launch_pc2(const StreamParams &sp, unsigned grid, cudaStream_t st)
{
    switch (sp.scan.sink.sec.idx) {
    case 0: return launch_pc3<AR, NEG, 0>(sp, grid, st);
    case 1: return launch_pc3<AR, NEG, 1>(sp, grid, st);
    case 2: return launch_pc3<AR, NEG, 2>(sp, grid, st);
    case 3: return launch_pc3<AR, NEG, 3>(sp, grid, st);
    case 4: return launch_pc3<AR, NEG, 4>(sp, grid, st);
    default: return launch_pc3<AR, NEG, 5>(sp, grid, st);
    }
}
template <class AR>
int launch_pc1(const StreamParams &sp, unsigned grid, cudaStream_t st)
{
    if (sp.prop.negmask == 0u) return launch_pc2<AR, 0>(sp, grid, st);
    if (sp.prop.negmask == 63u) return launch_pc2<AR, 1>(sp, grid, st);
    return launch_pc2<AR, 2>(sp, grid, st);
}

constexpr long long PC_FIXED_PER_TRAJ =
    (HB_CAND_CAP * (HB_CAND_DOUBLES + HB_DESC_DOUBLES) + HB_CAND_CAP / 2 + 1) * (long long)sizeof(double);

}  // namespace

extern "C" int64_t hb_section3_scratch_bytes(int64_t n, int32_t pool_records_per_traj)
{
    if (n < 0 || pool_records_per_traj < 1) return -1;
    const int64_t pool = (n * (int64_t)pool_records_per_traj + 148 * PC_NC * PC_POOL_CHUNK) * HB_REC_DOUBLES * 8;
    return n * PC_FIXED_PER_TRAJ + pool + 1024;
}

extern "C" int hb_cr3bp_section3(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                                 const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits,
                                 int64_t hit_capacity, int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc,
                                 int32_t *n_rej, int32_t *status, void *scratch, int64_t scratch_bytes, void *workspace,
                                 void *stream, void *const *stage_events)
{
    if (!sys || !integ || !sec) return HB_ERR_BADARG;
    if (integ->method != HB_DOP853) return HB_ERR_UNSUPPORTED;
    if (sec->idx < 0 || sec->idx > 5 || sec->proj_i < 0 || sec->proj_i > 5 || sec->proj_j < 0 || sec->proj_j > 5 ||
        sec->segment_refine < 0 || hit_capacity < 0)
        return HB_ERR_BADARG;
    if (n < 0 || m < 2 || !workspace || !t_eval ||
        (n > 0 && (!y0_soa || !yf_soa || !n_acc || !n_rej || !status || !scratch || (hit_capacity > 0 && !hits))))
        return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    auto mark = [&](int i) { if (stage_events && stage_events[i]) cudaEventRecord((cudaEvent_t)stage_events[i], st); };
    if (n == 0) { HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st)); return HB_OK; }
    if (n > 2147483647LL / HB_CAND_CAP) return HB_ERR_UNSUPPORTED;      // the compact segment index is 32-bit
    // scratch layout: cand | desc | cand_count, desc_count | desc_index | totals (256 B) | pool counter (256 B) | pool
    const long long pool_bytes = scratch_bytes - 1024 - n * PC_FIXED_PER_TRAJ;
    const long long pool_cap = pool_bytes / (HB_REC_DOUBLES * (long long)sizeof(double));
    if (pool_cap < 2 * PC_POOL_CHUNK) return HB_ERR_BADARG;
    double *cand = (double *)scratch;
    double *desc = cand + n * (long long)HB_CAND_CAP * HB_CAND_DOUBLES;
    int *cand_count = (int *)(desc + n * (long long)HB_CAND_CAP * HB_DESC_DOUBLES);    // n doubles = 2n ints
    int *desc_count = cand_count + n;
    int *desc_index = desc_count + n;                                   // n * HB_CAND_CAP ints
    // (the lists above end on a 4-byte boundary that depends on n: counters and pool start on the next 256-byte line)
    char *tail = (char *)(((uintptr_t)(desc_index + n * (long long)HB_CAND_CAP) + 255) & ~(uintptr_t)255);
    int *desc_total = (int *)tail;
    unsigned long long *pool_count = (unsigned long long *)(tail + 256);
    double *pool = (double *)(tail + 512);
    double ends[2];
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[0], t_eval, sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[1], t_eval + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    HB_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int) * 2 * (size_t)n, st));
    HB_CUDA_TRY(cudaMemsetAsync(desc_total, 0, 512, st));
    HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st));
    mark(0);
    StreamParams sp{};
    int rc = fill_params(sys, integ, sp.prop);
    if (rc != HB_OK) return rc;
    PropParams &pp = sp.prop;
    pp.n = n; pp.y0 = y0_soa; pp.t0 = ends[0]; pp.tf = ends[1]; pp.tf_arr = nullptr;
    pp.yf = yf_soa; pp.nacc = n_acc; pp.nrej = n_rej; pp.status = status;
    pp.ws = (HbWorkspace *)workspace;
    if (integ->arith == HB_ARITH_PARITY) {
        rc = first_steps_prepass<ArParity>(pp, st);
        if (rc != HB_OK) return rc;
    }
    ScanParams &p = sp.scan;
    p.prop = pp;
    p.n = n; p.rec = pool; p.rec_cap = INT_MAX; p.nacc = n_acc; p.status = status;
    p.t_eval = t_eval; p.m = m;
    p.tsign = sys->fwd < 0 ? -1.0 : 1.0;
    p.inv_grid_dt = (ends[1] > ends[0]) ? (double)(m - 1) / (ends[1] - ends[0]) : 0.0;
    p.sink.sec = *sec; p.sink.hits = hits; p.sink.capacity = hit_capacity; p.sink.ws = (HbWorkspace *)workspace;
    p.hits_per_traj = hits_per_traj;
    p.cand_count = cand_count; p.cand = cand; p.desc_count = desc_count; p.desc = desc; p.desc_total = desc_total;
    p.desc_index = desc_index;
    sp.pool = pool; sp.pool_cap = pool_cap; sp.pool_count = pool_count;
    long long grid = sm_count();                                       // persistent: one CTA per SM
    if (integ->max_ctas > 0 && integ->max_ctas < grid) grid = integ->max_ctas;
    const long long need = (n + 32 * PC_NP - 1) / (32 * PC_NP);
    if (need < grid) grid = need;
    rc = (integ->arith == HB_ARITH_PARITY) ? launch_pc1<ArParity>(sp, (unsigned)grid, st)
                                           : launch_pc1<ArFast>(sp, (unsigned)grid, st);
    if (rc != HB_OK) return rc;
    mark(1);
    rc = hb_scan_finish(p, integ->arith, st, stage_events ? (cudaEvent_t)stage_events[2] : nullptr);
    if (rc != HB_OK) return rc;
    mark(3);
    return HB_OK;
}
