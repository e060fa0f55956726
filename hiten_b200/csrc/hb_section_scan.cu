// hb_section_scan.cu -- tube + synodic section as a short pipeline with a compact intermediate (hb_cr3bp_section2).
//
//   A  k_dop853_6<.., MODE_RECORD, ..> (hb_cr3bp.cu): the plain propagation kernel, which additionally stores what
//      the dense output of every accepted step depends on (512 B: t_old, t_new, y_old, y_new, k6..k13) in a
//      caller-provided scratch [n][cap][64] -- ~45 KB per trajectory instead of the 226 KB of the 4713-sample tube.
//      It does NOT build the interpolant: that would double the loop body of a kernel that is instruction-fetch bound;
//   B1 k_step_scan: one warp per trajectory, a lane per step.  Runs the three extra DOP853 stages and builds the
//      EVENT COMPONENT of the step's interpolant in registers, finds the grid samples the step owns and proves
//      most steps quiet (interpolant bound, see hb_cr3bp_section.cu); the others are scanned by the whole warp,
//      32 grid samples per instruction, and the segments that can hold a hit are NOTED (8 numbers);
//   B2 k_compact_segments + k_emit_candidates: a compact index over the per-trajectory segment lists, then one
//      thread per noted segment rebuilds the two end states from the step records, runs the reference's
//      sub-interval logic and appends CANDIDATE hits {sample index, order, t, state};
//   B3 k_order_dedup: one thread per trajectory sorts its few candidates into the reference's order and applies
//      _order_and_dedup_hits (dedup against the previous kept hit, max_hits_per_traj), appending the hits.
//
// Why: the fused kernel (hb_cr3bp_section.cu) carries the scan state and 42 dense coefficients through the
// integration loop and is instruction-fetch bound (ncu); detection itself has no sequential dependence except
// the final de-duplication, so it is split off and made embarrassingly parallel.  Arithmetic per sample / segment
// is unchanged, so the hits are bit-identical to the fused kernel and to hb_cr3bp_dense + hb_synodic_detect
// (tests/test_gpu_synodic.py).  Trajectories with more accepted steps than `cap`, or more candidates than
// HB_CAND_CAP, get status HB_TRAJ_RECORD_OVERFLOW: rerun those with hb_cr3bp_section.
//
// Reference: algorithms/poincare/synodic/backend.py:458-659 (_detect_with_segment_refine), :382-455.
#include "hb_scan.cuh"
#include "hb_tubefilter.cuh"

extern "C" int hb_cr3bp_record_launch(const hb_cr3bp *sys, const hb_integ *integ, int32_t section_idx, int64_t n,
                                      const double *y0_soa, double t0, double tf, double *rec, int32_t rec_cap,
                                      double *yf_soa, int32_t *n_acc, int32_t *n_rej, int32_t *status, void *workspace,
                                      cudaStream_t st, const hb_section *near_section, double inv_grid_dt,
                                      int32_t *n_rec);

namespace {
using namespace hbc;
using namespace hbscan;

// One warp per trajectory, a lane per accepted step (chunks of 32 steps).  Each lane
//   1. rebuilds the EVENT COMPONENT of its step's dense interpolant from the stage record (three extra stages +
//      one column of the D rows: dense_component) -- the only HBM traffic of this kernel, 512 B per step;
//   2. finds the grid samples its step owns, t_old <= t_q < t_new (searchsorted 'right' - 1, rk.py:2505; the first
//      step also owns t_q = t0 and the last one everything that is left), and evaluates the event function at the
//      first / last / second-to-last of them; the segment that straddles two steps is tested with the neighbour's
//      values, which travel by shuffle (and, across chunks, in warp-uniform carries);
//   3. proves its step quiet where it can (interpolant bound, see hb_cr3bp_section.cu); the remaining steps are
//      scanned by the whole warp, 32 grid samples per instruction.
// Segments that can hold a hit are only NOTED here (8 numbers); k_emit_candidates refines them.
// The 32 records of a chunk are staged in shared memory by the copy engine: every lane issues ONE 512-byte bulk
// copy (cp.async.bulk, completion counted on a per-warp mbarrier) instead of 28 dependent vector loads, so the HBM
// latency is paid once per chunk and other warps compute meanwhile.  Rows are padded to 528 B (33 x 16 B): the
// lanes' 16-byte reads then hit distinct banks.  (Per-thread 16-byte cp.async copies instead of the bulk copies were
// measured: 18.0 ms instead of 11.1 ms per 1e6 trajectories.)
// L2 prefetch of the lane's next-chunk record (cp.async.bulk.prefetch.L2): measured 11.39 ms vs 11.18 ms without per
// 1e6 trajectories -- the kernel does not wait on DRAM latency (12 warps x 16 KB in flight per SM) -- so it is off
#ifndef HB_SCAN_PREFETCH
#define HB_SCAN_PREFETCH 0
#endif
// HB_SCAN_PIPELINED = 1 issues the next chunk's copies as soon as the rows of this chunk have been read: they fly
// during the scan phase, which works on the headers in registers. Measured on 1e6 trajectories: 11.27 -> 10.90 ms.
#ifndef HB_SCAN_PIPELINED
#define HB_SCAN_PIPELINED 1
#endif
// Staging geometry: a record carries 62 doubles (HB_REC_K12 + 6); rows of 62 x 8 = 496 B = 31 x 16 B need no padding
// (31 is odd: the 16-byte reads of a quarter warp fall in 8 different bank groups), so a warp stages 15.9 KB and 13
// one-warp CTAs fit an SM (13 x (15.9 + 1 KB reserved) = 220 KB); with 528-byte rows and 4-warp CTAs it was 12 warps.
#ifndef HB_SCAN_WARPS
#define HB_SCAN_WARPS 1
#endif
#ifndef HB_SCAN_MINBLOCKS
#define HB_SCAN_MINBLOCKS 13
#endif
// Sparse records carry their step number and flags in [62]: the whole 512-byte record is copied, into 528-byte rows (33
// x 16 B); there the scan is a small part of the step, so the lower occupancy does not matter.
template <bool SPARSE> struct ScanGeom {
    static constexpr int COPY = SPARSE ? HB_REC_DOUBLES * 8 : (HB_REC_K12 + 6) * 8;
    static constexpr int ROW = SPARSE ? HB_REC_DOUBLES * 8 + 16 : COPY;
    // records per chunk: a sparse trajectory has ~10 records, so 16-row chunks halve the staging memory (8.4 KB per
    // warp: the occupancy is then bounded by registers, 16 warps per SM instead of 12) at no cost in rounds
    static constexpr int CH = SPARSE ? 16 : 32;
    // ... and two such trajectories share a warp, one per half (GW lanes): the kernel is bound by per-warp latency with
    // a third of the lanes busy otherwise.  PACK x CH = 32 rows per warp in both modes.
    static constexpr int PACK = 32 / CH;
    static constexpr int SMEM = HB_SCAN_WARPS * (32 * ROW + 16);
    static_assert(COPY % 16 == 0 && (ROW / 16) % 2 == 1, "row: 16-byte multiple, odd number of 16-byte units");
};

// SPARSE: the scratch holds only the records of steps that can come near the section plane and their neighbours
// (MODE_RECORD_NEAR of hb_cr3bp.cu), p.nrec of them per trajectory, each tagged with its step number: lanes whose step
// does not follow the previous lane's step start an "island".  Every step that was left out is provably quiet and at
// least two grid spacings long (it owns samples, all of them away from the plane and on one side), so an island's first
// step -- itself such a quiet step, written only because its successor is near the plane, or the trajectory's first
// step -- has no segment to test against its predecessor; it finds its first sample with its own grid lookup.
template <class AR, int C, bool SPARSE>  // C = section component (compile time: the unused parts of the extra stages fall away)
__global__ void __launch_bounds__(32 * HB_SCAN_WARPS, HB_SCAN_MINBLOCKS) k_step_scan(const ScanParams p)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int HB_SCAN_COPY = ScanGeom<SPARSE>::COPY, HB_SCAN_ROW = ScanGeom<SPARSE>::ROW, CH = ScanGeom<SPARSE>::CH;
    constexpr int PACK = ScanGeom<SPARSE>::PACK, GW = 32 / PACK;         // trajectories per warp, lanes per trajectory
    constexpr unsigned GMASK = (GW == 32) ? FULL : ((1u << GW) - 1u);
    static_assert(GW == CH, "a group's lanes are the rows of its chunk");
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gl = lane & (GW - 1), gb = lane & ~(GW - 1);               // lane within its group, the group's first lane
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp_id * PACK >= p.n) return;                        // whole warp
    const long long traj_raw = warp_id * PACK + (lane / GW);
    const bool traj_ok = traj_raw < p.n;                      // (an odd batch leaves the last warp's second half idle)
    const long long traj = traj_ok ? traj_raw : p.n - 1;
    unsigned char *wbase = scan_smem + wid * (32 * HB_SCAN_ROW + 16);
    const double *row_ptr = (const double *)(wbase + lane * HB_SCAN_ROW);
    const unsigned mbar = smem_u32(wbase + 32 * HB_SCAN_ROW), row_u32 = smem_u32(row_ptr);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;
    // Pipelined: the first chunk is requested BEFORE the trajectory's step count is known (the count is a DRAM round
    // trip of its own): until then every row of the scratch is fair game -- rows past the count hold stale bytes that
    // no lane looks at.
    const int *count = SPARSE ? p.nrec : p.nacc;        // records of this trajectory in the scratch
    // (sparse records: a trajectory has ~10 of them, so the count is read first -- requesting a full chunk blindly
    // would read 16 KB per trajectory, three times what is there)
    int nacc = (HB_SCAN_PIPELINED && !SPARSE) ? p.rec_cap : (traj_ok ? min(count[traj], p.rec_cap) : 0);
    const double off = p.sink.sec.offset, tol_s = p.sink.sec.tol_on_surface;
    const int sdir = p.sink.sec.direction;
    int ndesc = 0;                         // segments noted so far (uniform within a group, like the carries below)
    int carry_c = 0;                       // first grid sample not owned yet
    int carry_step = 0;                    // record that owns sample carry_c - 1
    int carry_no = -1;                     // SPARSE: step number of the last record of the previous chunk
    double carry1 = 0.0, carry2 = 0.0;     // event function at samples carry_c - 1, carry_c - 2
    const double te0 = p.t_eval[0];
    double carry_t = te0;                  // t_eval[carry_c]
    // one 512-byte bulk copy per lane (the lane's own record of the chunk) or a plain arrival
    auto issue_chunk = [&](int b) {
        const int sb = b + gl;
        if (sb < nacc) {
            const double *src = p.rec + (traj * p.rec_cap + sb) * HB_REC_DOUBLES;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(HB_SCAN_COPY)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(row_u32), "l"(src), "r"(HB_SCAN_COPY), "r"(mbar) : "memory");
#if HB_SCAN_PREFETCH
            // the next chunk's record of this lane: start it towards L2 now, so that the blocking wait of the next
            // round pays an L2 hit instead of a DRAM round trip
            if (sb + CH < nacc)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + CH * HB_REC_DOUBLES),
                             "r"(HB_REC_DOUBLES * 8) : "memory");
#endif
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
        }
    };
#if HB_SCAN_PIPELINED
    if (!SPARSE) {
        issue_chunk(0);
        nacc = min(count[traj], p.rec_cap);
        if (nacc <= 0) mbar_wait(mbar, phase);                // nothing to scan: let the requested rows land before exit
    } else {
        issue_chunk(0);
    }
#endif
    int nacc_max = nacc;                                      // the warp runs as many rounds as its longest trajectory needs
    if (PACK > 1) nacc_max = max(nacc_max, __shfl_xor_sync(FULL, nacc_max, GW & 31));
    for (int base = 0; base < nacc_max; base += CH) {
        const int s = base + gl;
        const bool have_rec = s < nacc;
        double hdr[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) hdr[i] = 0.0;
        int cend = p.m;
        double t_c = 0.0, t_cm1 = 0.0, t_cm2 = 0.0;   // t_eval[cend], t_eval[cend - 1], t_eval[cend - 2]
        int step_no = s, own_c0 = 0;                  // SPARSE: the record's step number; first sample at or after t_old
        double own_t_c0 = te0;
#if !HB_SCAN_PIPELINED
        issue_chunk(base);
#endif
        mbar_wait(mbar, phase);
        phase ^= 1u;
        if (have_rec) {
            const double *r = row_ptr;
            double v[16];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const double2 w = *(const double2 *)(r + i);
                v[i] = w.x; v[i + 1] = w.y;
            }
            const double y[6] = {v[2], v[3], v[4], v[5], v[6], v[7]}, yn[6] = {v[8], v[9], v[10], v[11], v[12], v[13]};
            const double hseg = AR::sub(v[1], v[0]);
            hdr[0] = v[0]; hdr[1] = v[1]; hdr[2] = hseg; hdr[3] = y[C];
            if (hseg != 0.0) {
                const Cr3bpRhs<AR, 2> rhs{p.prop};
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
            bool is_last = s == nacc - 1;
            if (SPARSE) {
                const int2 meta = *(const int2 *)(r + HB_REC_META);
                step_no = meta.x;
                is_last = (meta.y & HB_REC_LAST) != 0;
            }
            if (!is_last) cend = first_at_or_after3(p, te0, v[1], t_c, t_cm1, t_cm2);
            else { t_cm1 = p.t_eval[p.m - 1]; t_cm2 = p.t_eval[max(p.m - 2, 0)]; }
            if (SPARSE && step_no > 0) {              // (only an island's first lane uses it; the lookup is cheap)
                double d1, d2;
                own_c0 = first_at_or_after3(p, te0, v[0], own_t_c0, d1, d2);
            }
        }
#if HB_SCAN_PIPELINED
        // the rows are not read again in this round (the scan below runs on the headers in registers): the next
        // chunk's copies fly while this chunk is scanned
        __syncwarp();
        if (base + CH < nacc_max) issue_chunk(base + CH);
#endif
        int c0 = __shfl_up_sync(FULL, cend, 1);               // t_new of a step is t_old of the next, bit for bit
        double t_c0 = __shfl_up_sync(FULL, t_c, 1);                       // t_eval[c0]: the previous lane's t_eval[cend]
        if (gl == 0) { c0 = carry_c; t_c0 = carry_t; }
        bool adjacent = true;                                 // this record's step follows the previous record's step
        if (SPARSE) {
            int prev_no = __shfl_up_sync(FULL, step_no, 1);
            if (gl == 0) prev_no = carry_no;
            adjacent = step_no == prev_no + 1;
            if (!adjacent) { c0 = own_c0; t_c0 = own_t_c0; }  // island start (or the trajectory's first step: c0 = 0)
        }
        const int nown = (have_rec && c0 < cend) ? cend - c0 : 0;
        const bool owns = nown > 0;
        // event function at the first (g_first), last (A) and second-to-last (B) owned sample
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
        // (masks and lane numbers below are relative to the lane's group)
        const unsigned own_mask = (__ballot_sync(FULL, owns) >> gb) & GMASK;
        const unsigned two_mask = (__ballot_sync(FULL, nown >= 2) >> gb) & GMASK;
        // nearest owner below this lane (q) and below that (q2): they hold samples c0 - 1 and (maybe) c0 - 2
        const unsigned below = own_mask & ((1u << gl) - 1u);
        const int q = below ? 31 - __clz(below) : -1;
        const unsigned below2 = (q >= 0) ? (own_mask & ((1u << q) - 1u)) : 0u;
        const int q2 = below2 ? 31 - __clz(below2) : -1;
        const double Aq = shfl_d(A, gb + (q >= 0 ? q : 0)), Bq = shfl_d(B, gb + (q >= 0 ? q : 0)),
                     Aq2 = shfl_d(A, gb + (q2 >= 0 ? q2 : 0));
        const double g_prev = (q >= 0) ? Aq : carry1;
        const double gm2 = (q >= 0) ? (((two_mask >> q) & 1u) ? Bq : (q2 >= 0 ? Aq2 : carry1)) : carry2;
        const int s_prev = (q >= 0) ? base + q : carry_step;
        bool scan = false;
        {                                                     // the segment that ends at this step's first sample
            const bool flag = owns && c0 > 0 && adjacent && segment_may_hit(sdir, g_prev, g_first, tol_s);
            const unsigned fm = (__ballot_sync(FULL, flag) >> gb) & GMASK;
            if (flag)
                store_segment(p, traj, ndesc + __popc(fm & ((1u << gl) - 1u)), c0, traj * p.rec_cap + s_prev,
                              traj * p.rec_cap + s, g_prev, g_first, c0 > 1 ? gm2 : 0.0);
            ndesc += __popc(fm);
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
        // cooperative scan of the non-quiet steps of this chunk, one owner lane at a time
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
            const int sL = base + (L & (GW - 1));
            const long long trajL = __shfl_sync(FULL, traj, L);      // (the whole warp scans: L may be in the other group)
            int ndL = __shfl_sync(FULL, ndesc, L);
            double tq_next = p.t_eval[min(b0 + 1 + lane, b1 - 1)];
            for (int b = b0 + 1; b < b1; b += 32) {
                const int cs = b + lane;
                const bool valid = cs < b1;
                const double tq = tq_next;                    // (out-of-range lanes read sample b1 - 1: unused)
                if (b + 32 < b1) tq_next = p.t_eval[min(cs + 32, b1 - 1)];   // in flight during this batch
                const double g = g_comp<AR>(bh, xpar_by<AR>(tq, bh[0], bh[2], binv), off);
                double g_m1 = __shfl_up_sync(FULL, g, 1);
                double g_m2 = __shfl_up_sync(FULL, g, 2);
                if (lane == 0) { g_m1 = c1; g_m2 = c2; }
                if (lane == 1) g_m2 = c1;
                const bool flagged = valid && segment_may_hit(sdir, g_m1, g, tol_s);
                const unsigned fm = __ballot_sync(FULL, flagged);
                if (flagged)
                    store_segment(p, trajL, ndL + __popc(fm & ((1u << lane) - 1u)), cs, trajL * p.rec_cap + sL,
                                  trajL * p.rec_cap + sL, g_m1, g, g_m2);
                ndL += __popc(fm);
                const int nvalid = min(32, b1 - b);
                const double l1 = shfl_d(g, nvalid - 1);
                const double l2 = shfl_d(g, nvalid >= 2 ? nvalid - 2 : 0);
                c2 = (nvalid >= 2) ? l2 : c1;
                c1 = l1;
            }
            if ((L & ~(GW - 1)) == gb) ndesc = ndL;            // the owner's group keeps its segment count
        }
        // carries for the next chunk (per group; the shuffles run for every lane)
        {
            const int qL = own_mask ? 31 - __clz(own_mask) : 0;
            const unsigned bl = own_mask & ((1u << qL) - 1u);
            const int qL2 = bl ? 31 - __clz(bl) : -1;
            const double a = shfl_d(A, gb + qL), b = shfl_d(B, gb + qL), a2 = shfl_d(A, gb + (qL2 >= 0 ? qL2 : 0));
            if (own_mask) {
                carry2 = ((two_mask >> qL) & 1u) ? b : (qL2 >= 0 ? a2 : carry1);
                carry1 = a;
                carry_step = base + qL;
            }
        }
        {
            const int cl = __shfl_sync(FULL, cend, gb + CH - 1);
            const double tl = shfl_d(t_c, gb + CH - 1);
            const int nl = __shfl_sync(FULL, step_no, gb + CH - 1);
            if (base + CH <= nacc) {                          // (a group whose records ran out keeps its carries: unused)
                carry_c = cl; carry_t = tl;
                if (SPARSE) carry_no = nl;
            }
        }
        __syncwarp();                                         // every lane is done with its row before the next copy
    }
    if (gl == 0 && traj_ok) p.desc_count[traj] = ndesc;
}

// Compact index of all noted segments (so that k_emit_candidates runs full warps): one thread per trajectory, two
// atomics per warp.  Segments whose two samples sit in ONE step (the common case) are indexed from the front, those
// that straddle two steps (two records to rebuild) from the back, so that almost every emit warp is of one kind.
HB_DEV int warp_exclusive_base(int mine, int lane, int *counter)
{
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    return __shfl_sync(0xffffffffu, base, 31) + incl - mine;
}

__global__ void __launch_bounds__(256) k_compact_segments(const ScanParams p)
{
    const long long traj = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    // a trajectory with more noted segments than its list holds -- or marked that way because the record pool of
    // hb_cr3bp_section3 ran dry, in which case the list is only partly written -- is reported as overflowed by
    // k_order_dedup and rerun by the caller: none of its segments is indexed
    const int cnt = (traj < p.n) ? p.desc_count[traj] : 0;
    const int nd = (cnt > HB_CAND_CAP) ? 0 : cnt;
    unsigned same = 0u;                                       // bit i: segment i has both samples in one step
    for (int i = 0; i < nd; ++i) {
        const double *d = p.desc + (traj * HB_CAND_CAP + i) * HB_DESC_DOUBLES;
        if (d[2] == d[3]) same |= 1u << i;
    }
    const int n_same = __popc(same), n_split = nd - n_same;
    int front = warp_exclusive_base(n_same, lane, &p.desc_total[0]);
    int back = warp_exclusive_base(n_split, lane, &p.desc_total[1]);
    const long long last = p.n * HB_CAND_CAP - 1;
    for (int i = 0; i < nd; ++i) {
        const int where = (int)(traj * HB_CAND_CAP + i);
        if ((same >> i) & 1u) p.desc_index[front++] = where;
        else p.desc_index[last - back++] = where;
    }
}

// One thread per noted segment (grid-stride over the compact index): rebuild the two end states from the step
// records and run the reference's sub-interval logic, appending candidate hits.
template <class AR>
__global__ void __launch_bounds__(128) k_emit_candidates(const ScanParams p)
{
    const int n_front = p.desc_total[0], total = n_front + p.desc_total[1];
    const long long last = p.n * HB_CAND_CAP - 1;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (long long)gridDim.x * blockDim.x) {
        const long long at = (it < n_front) ? it : last - (it - n_front);
        const double *d = p.desc + (long long)p.desc_index[at] * HB_DESC_DOUBLES;
        double d0, d1, d2, d3, gk, gk1, gm2, pad;
        hb_ld4(d, d0, d1, d2, d3);
        hb_ld4(d + 4, gk, gk1, gm2, pad);
        const long long traj = (long long)d0;
        const int cs = (int)d1;
        const long long r0 = (long long)d2, r1 = (long long)d3;
        double x0[6], x1[6];
        if (r0 == r1) {
            states_from_record<AR>(p.rec + r1 * HB_REC_DOUBLES, p.prop, p.t_eval[cs - 1], p.t_eval[cs], x0, x1);
        } else {
            double dummy[6];
            states_from_record<AR>(p.rec + r0 * HB_REC_DOUBLES, p.prop, p.t_eval[cs - 1], p.t_eval[cs - 1], x0, dummy);
            states_from_record<AR>(p.rec + r1 * HB_REC_DOUBLES, p.prop, p.t_eval[cs], p.t_eval[cs], x1, dummy);
        }
        auto emit = [&](int order, double th, const double (&xh)[6]) {
            const int k = atomicAdd(&p.cand_count[traj], 1);
            if (k < HB_CAND_CAP) {
                double *c = p.cand + (traj * HB_CAND_CAP + k) * HB_CAND_DOUBLES;
                // sort key: grid segment, then order inside it
                hb_st4(c, (double)cs * 4096.0 + (double)order, th, xh[0], xh[1]);
                hb_st4(c + 4, xh[2], xh[3], xh[4], xh[5]);
            }
        };
        segment_candidates(p.sink.sec, cs > 1, gm2, gk, gk1, __dmul_rn(p.tsign, p.t_eval[cs - 1]),
                           __dmul_rn(p.tsign, p.t_eval[cs]), x0, x1, emit);
    }
}

// _order_and_dedup_hits (backend.py:430-455) on the candidates of one trajectory
__global__ void __launch_bounds__(256) k_order_dedup(const ScanParams p)
{
    const long long traj = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (traj >= p.n) return;
    int k = p.cand_count[traj];
    const int n_records = p.nrec ? p.nrec[traj] : p.nacc[traj];
    if (n_records > p.rec_cap || k > HB_CAND_CAP || p.desc_count[traj] > HB_CAND_CAP) {
        // incomplete: report no hits for this trajectory; the caller reruns it with hb_cr3bp_section
        p.status[traj] = HB_TRAJ_RECORD_OVERFLOW;
        atomicAdd(&p.sink.ws->rec_overflow, 1ULL);
        if (p.hits_per_traj) p.hits_per_traj[traj] = 0;
        return;
    }
    const double *c = p.cand + traj * HB_CAND_CAP * HB_CAND_DOUBLES;
    unsigned long long used = 0ULL;
    Dedup dd{0.0, 0.0, 0.0, 0};
    for (int it = 0; it < k; ++it) {
        int best = -1;
        double bk = 0.0;
        for (int j = 0; j < k; ++j) {                        // k is tiny: selection by key
            if ((used >> j) & 1ULL) continue;
            const double key = c[j * HB_CAND_DOUBLES];
            if (best < 0 || key < bk) { best = j; bk = key; }
        }
        used |= 1ULL << best;
        const double *e = c + best * HB_CAND_DOUBLES;
        const double xh[6] = {e[2], e[3], e[4], e[5], e[6], e[7]};
        if (!push_hit(p.sink, dd, traj, e[1], xh, 0)) break;
    }
    if (p.hits_per_traj) p.hits_per_traj[traj] = dd.n;
}


// ---- trajectory filters on the step records (hb_section2_filter) -------------------------------------------
// Manifold.compute() drops a trajectory when it comes closer than the safe radii to a primary or when its Jacobi
// constant drifts (manifold.py:412-432) -- judged on the 4713 dense samples, which the section pipeline never
// stores.  This kernel evaluates exactly those samples from the step records that hb_cr3bp_section2 left in its
// scratch: one warp per trajectory; per chunk of 32 steps a lane builds the full interpolant of its step (three extra
// stages + 7 x 6 coefficients, dense_cache) and parks it in shared memory (51 doubles per step, odd stride:
// conflict-free), then the warp takes the chunk's steps one after the other: coefficients broadcast from shared
// memory into registers, the step's grid samples spread over the lanes 32 at a time -- six-component Horner
// evaluation, the reference's r1 / r2 / Jacobi expressions (hb_tubefilter.cuh) -- and reduces.  Same arithmetic per
// sample as hb_cr3bp_dense + hb_tube_filter: same bits.
constexpr int HB_FILT_WARPS = 4;
// 3 CTAs x 4 warps per SM (168 registers): measured 28.9 ms per 262144 trajectories vs 37.7 ms at 2 CTAs (255 registers)
// and 39.1 ms at 4 CTAs (128 registers: the sample loop spills)
#ifndef HB_FILT_MINBLOCKS
#define HB_FILT_MINBLOCKS 3
#endif
constexpr int HB_FILT_ROW = 51;       // F[7][6], y_old[6], t_old, hseg, 1/hseg
constexpr int HB_FILT_SMEM = HB_FILT_WARPS * (32 * HB_FILT_ROW * 8 + 32 * 4);

struct FilterParams {
    PropParams prop;
    long long n;
    const double *rec;
    int rec_cap;
    const int *nacc;
    const int *status;
    const double *t_eval;
    int m;
    double inv_grid_dt;
    hb_tube_filter_opts o;
    double *out;            // [n][3]
    int *keep;              // [n]: 1 keep, 0 discard, -1 not judged (trajectory has no complete record set)
};

template <class AR>
__global__ void __launch_bounds__(32 * HB_FILT_WARPS, HB_FILT_MINBLOCKS) k_record_filter(const FilterParams p)
{
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char filt_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long traj = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (traj >= p.n) return;                                  // whole warp
    double *rows = (double *)filt_smem + wid * (32 * HB_FILT_ROW);
    int *cends = (int *)((double *)filt_smem + HB_FILT_WARPS * (32 * HB_FILT_ROW)) + wid * 32;
    const int nacc = p.nacc[traj];
    if (p.status[traj] != HB_TRAJ_OK || nacc > p.rec_cap || nacc < 1) {
        if (lane == 0) {
            p.out[3 * traj] = p.out[3 * traj + 1] = p.out[3 * traj + 2] = CUDART_NAN;
            p.keep[traj] = -1;
        }
        return;
    }
    ScanParams sp{};                                           // first_at_or_after only reads the grid fields
    sp.t_eval = p.t_eval; sp.m = p.m; sp.inv_grid_dt = p.inv_grid_dt;
    const double mu1 = __dsub_rn(1.0, p.o.mu), mu2 = p.o.mu;
    const double *rec0 = p.rec + traj * (long long)p.rec_cap * HB_REC_DOUBLES;
    double s0[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) s0[d] = rec0[HB_REC_YOLD + d];   // sample 0 = y0
    const double C0 = jacobi_ref(s0, mu1, mu2), absC0 = fabs(C0);
    TubeFilterAcc acc;
    int carry_c = 0;
    for (int base = 0; base < nacc; base += 32) {
        const int s = base + lane;
        const bool have_rec = s < nacc;
        int cend = p.m;
        double *row = rows + lane * HB_FILT_ROW;
        if (have_rec) {
            double t, t_new, y[6], yn[6], k[13][6], F[7][6];
            load_record<AR>(rec0 + (long long)s * HB_REC_DOUBLES, p.prop, t, t_new, y, yn, k);
            const double hseg = AR::sub(t_new, t);
            if (hseg != 0.0) {
                const Cr3bpRhs<AR, 2> rhs{p.prop};
                dense_cache<AR>(y, yn, hseg, k, F, rhs);
            } else {
#pragma unroll
                for (int i = 0; i < 7; ++i)
#pragma unroll
                    for (int d = 0; d < 6; ++d) F[i][d] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < 7; ++i)
#pragma unroll
                for (int d = 0; d < 6; ++d) row[6 * i + d] = F[i][d];
#pragma unroll
            for (int d = 0; d < 6; ++d) row[42 + d] = y[d];
            row[48] = t;
            row[49] = hseg;
            row[50] = (hseg != 0.0) ? AR::rcp(hseg) : 0.0;
            if (s != nacc - 1) cend = first_at_or_after(sp, t_new, 0);
        }
        cends[lane] = cend;
        __syncwarp();
        // the chunk's steps one after the other (warp-uniform loop): the step's 48 coefficients are read once from
        // shared memory (broadcast) into registers, its samples go to the lanes 32 at a time
        const int nst = min(32, nacc - base);
        int c0 = carry_c;
        for (int sl = 0; sl < nst; ++sl) {
            const int c1 = cends[sl];
            if (c0 < c1) {
                const double *r = rows + sl * HB_FILT_ROW;
                double F[7][6], y[6];
#pragma unroll
                for (int i = 0; i < 7; ++i)
#pragma unroll
                    for (int d = 0; d < 6; ++d) F[i][d] = r[6 * i + d];
#pragma unroll
                for (int d = 0; d < 6; ++d) y[d] = r[42 + d];
                const double t_old = r[48], hseg = r[49], inv = r[50];
                for (int c = c0 + lane; c < c1; c += 32) {
                    double st[6];
                    if (hseg != 0.0) dense_eval<AR>(y, F, xpar_by<AR>(p.t_eval[c], t_old, hseg, inv), st);
                    else {
#pragma unroll
                        for (int d = 0; d < 6; ++d) st[d] = y[d];
                    }
                    acc.sample(st, c, p.o.mu, mu1, mu2, C0);
                }
                c0 = c1;
            }
        }
        carry_c = c0;
        __syncwarp();
    }
    acc.warp_reduce();
    if (lane == 0) acc.store(p.o, traj, absC0, p.out, p.keep);
}
}  // namespace

namespace {
template <class AR, int C>
int launch_scan_c(const ScanParams &p, unsigned grid, cudaStream_t st)
{
    // per-device attribute: set on every launch (cheap), so any device of the process gets the opt-in
    if (p.nrec) {
        HB_CUDA_TRY(cudaFuncSetAttribute(k_step_scan<AR, C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScanGeom<true>::SMEM));
        k_step_scan<AR, C, true><<<grid, 32 * HB_SCAN_WARPS, ScanGeom<true>::SMEM, st>>>(p);
        return HB_OK;
    }
    HB_CUDA_TRY(cudaFuncSetAttribute(k_step_scan<AR, C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScanGeom<false>::SMEM));
    k_step_scan<AR, C, false><<<grid, 32 * HB_SCAN_WARPS, ScanGeom<false>::SMEM, st>>>(p);
    return HB_OK;
}
template <class AR>
int launch_scan(const ScanParams &p, unsigned grid, cudaStream_t st)
{
    switch (p.sink.sec.idx) {
    case 0: return launch_scan_c<AR, 0>(p, grid, st);
    case 1: return launch_scan_c<AR, 1>(p, grid, st);
    case 2: return launch_scan_c<AR, 2>(p, grid, st);
    case 3: return launch_scan_c<AR, 3>(p, grid, st);
    case 4: return launch_scan_c<AR, 4>(p, grid, st);
    default: return launch_scan_c<AR, 5>(p, grid, st);
    }
}
}  // namespace

// B2 + B3 (shared with hb_section_stream.cu)
int hbscan::hb_scan_finish(const ScanParams &p, int arith, cudaStream_t st, cudaEvent_t ev_emit_done)
{
    k_compact_segments<<<(unsigned)((p.n + 255) / 256), 256, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    const unsigned tb = (unsigned)sm_count() * 8u;
    if (arith == HB_ARITH_PARITY) k_emit_candidates<ArParity><<<tb, 128, 0, st>>>(p);
    else k_emit_candidates<ArFast><<<tb, 128, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    if (ev_emit_done) HB_CUDA_TRY(cudaEventRecord(ev_emit_done, st));
    k_order_dedup<<<(unsigned)((p.n + 255) / 256), 256, 0, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}

extern "C" int64_t hb_section2_scratch_bytes(int64_t n, int32_t steps_capacity)
{
    if (n < 0 || steps_capacity < 1) return -1;
    steps_capacity = (steps_capacity + 31) / 32 * 32;
    return n * ((int64_t)steps_capacity * HB_REC_DOUBLES + HB_S2_FIXED_DOUBLES) * (int64_t)sizeof(double) + 256;
}

extern "C" int hb_cr3bp_section2(const hb_cr3bp *sys, const hb_integ *integ, const hb_section *sec, int64_t n,
                                 const double *y0_soa, const double *t_eval, int32_t m, hb_hit *hits,
                                 int64_t hit_capacity, int32_t *hits_per_traj, double *yf_soa, int32_t *n_acc,
                                 int32_t *n_rej, int32_t *status, void *scratch, int64_t scratch_bytes, void *workspace,
                                 void *stream, void *const *stage_events, int32_t records)
{
    if (!sys || !integ || !sec) return HB_ERR_BADARG;
    if (records != HB_RECORDS_ALL && records != HB_RECORDS_NEAR_SECTION) return HB_ERR_BADARG;
    // optional caller-owned CUDA events, recorded on `stream` around the four stages (no library state involved)
    auto mark = [&](int i, cudaStream_t s_) { if (stage_events && stage_events[i]) cudaEventRecord((cudaEvent_t)stage_events[i], s_); };
    if (integ->method != HB_DOP853) return HB_ERR_UNSUPPORTED;
    if (sec->idx < 0 || sec->idx > 5 || sec->proj_i < 0 || sec->proj_i > 5 || sec->proj_j < 0 || sec->proj_j > 5 ||
        sec->segment_refine < 0 || hit_capacity < 0)
        return HB_ERR_BADARG;
    if (n < 0 || m < 2 || !workspace || !t_eval ||
        (n > 0 && (!y0_soa || !yf_soa || !n_acc || !n_rej || !status || !scratch || (hit_capacity > 0 && !hits))))
        return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { HB_CUDA_TRY(cudaMemsetAsync(workspace, 0, sizeof(HbWorkspace), st)); return HB_OK; }
    const long long per_traj_fixed = HB_S2_FIXED_DOUBLES * (long long)sizeof(double);
    long long cap = ((scratch_bytes - 256) / n - per_traj_fixed) / (HB_REC_DOUBLES * (long long)sizeof(double));
    cap -= cap % 32;                                 // keeps every trajectory's records 32-step aligned
    if (cap < 32) return HB_ERR_BADARG;
    const int rec_cap = cap > 100000 ? 99968 : (int)cap;
    double *rec = (double *)scratch;
    double *cand = rec + n * (long long)rec_cap * HB_REC_DOUBLES;
    double *desc = cand + n * (long long)HB_CAND_CAP * HB_CAND_DOUBLES;
    int *cand_count = (int *)(desc + n * (long long)HB_CAND_CAP * HB_DESC_DOUBLES);    // 2n doubles = 4n ints
    int *desc_count = cand_count + n;
    int *rec_count = desc_count + n;                  // sparse records: records per trajectory
    int *desc_index = rec_count + 2 * n;              // n * HB_CAND_CAP ints
    int *desc_total = desc_index + n * (long long)HB_CAND_CAP;   // first of the 256 trailing bytes
    double ends[2];
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[0], t_eval, sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[1], t_eval + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    HB_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int) * 2 * (size_t)n, st));
    HB_CUDA_TRY(cudaMemsetAsync(desc_total, 0, 2 * sizeof(int), st));
    mark(0, st);
    const bool sparse = records == HB_RECORDS_NEAR_SECTION;
    const double inv_grid_dt = (ends[1] > ends[0]) ? (double)(m - 1) / (ends[1] - ends[0]) : 0.0;
    int rc = hb_cr3bp_record_launch(sys, integ, sec->idx, n, y0_soa, ends[0], ends[1], rec, rec_cap, yf_soa, n_acc, n_rej,
                                    status, workspace, st, sparse ? sec : nullptr, inv_grid_dt, sparse ? rec_count : nullptr);
    if (rc != HB_OK) return rc;
    mark(1, st);
    ScanParams p{};
    rc = fill_params(sys, integ, p.prop);
    if (rc != HB_OK) return rc;
    p.n = n; p.rec = rec; p.rec_cap = rec_cap; p.nacc = n_acc; p.status = status;
    p.nrec = sparse ? rec_count : nullptr;
    p.t_eval = t_eval; p.m = m;
    p.tsign = sys->fwd < 0 ? -1.0 : 1.0;
    p.inv_grid_dt = inv_grid_dt;
    p.sink.sec = *sec; p.sink.hits = hits; p.sink.capacity = hit_capacity; p.sink.ws = (HbWorkspace *)workspace;
    p.hits_per_traj = hits_per_traj;
    p.cand_count = cand_count; p.cand = cand; p.desc_count = desc_count; p.desc = desc; p.desc_total = desc_total; p.desc_index = desc_index;
    const int per_block = HB_SCAN_WARPS * (sparse ? ScanGeom<true>::PACK : 1);   // trajectories per block (warp or half-warp each)
    const long long b1 = (n + per_block - 1) / per_block;
    if (b1 > 2147483647LL) return HB_ERR_BADARG;
    rc = (integ->arith == HB_ARITH_PARITY) ? launch_scan<ArParity>(p, (unsigned)b1, st) : launch_scan<ArFast>(p, (unsigned)b1, st);
    if (rc != HB_OK) return rc;
    HB_CUDA_TRY(cudaGetLastError());
    mark(2, st);
    rc = hb_scan_finish(p, integ->arith, st, stage_events ? (cudaEvent_t)stage_events[3] : nullptr);
    if (rc != HB_OK) return rc;
    mark(4, st);
    return HB_OK;
}

// Trajectory filters of Manifold.compute() for the trajectories of the LAST hb_cr3bp_section2 call that used this
// scratch block (declared in include/hiten_b200.h).
extern "C" int hb_section2_filter(const hb_cr3bp *sys, const hb_integ *integ, const hb_tube_filter_opts *opts,
                                  int64_t n, const double *t_eval, int32_t m, const int32_t *n_acc,
                                  const int32_t *status, const void *scratch, int64_t scratch_bytes, double *out,
                                  int32_t *keep, void *stream)
{
    if (!sys || !integ || !opts || n < 0 || m < 2) return HB_ERR_BADARG;
    if (integ->method != HB_DOP853) return HB_ERR_UNSUPPORTED;
    if (n == 0) return HB_OK;
    if (!t_eval || !n_acc || !status || !scratch || !out || !keep) return HB_ERR_BADARG;
    const long long per_traj_fixed = HB_S2_FIXED_DOUBLES * (long long)sizeof(double);
    long long cap = ((scratch_bytes - 256) / n - per_traj_fixed) / (HB_REC_DOUBLES * (long long)sizeof(double));
    cap -= cap % 32;                                 // the same record capacity hb_cr3bp_section2 derived
    if (cap < 32) return HB_ERR_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    double ends[2];
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[0], t_eval, sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaMemcpyAsync(&ends[1], t_eval + (m - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    HB_CUDA_TRY(cudaStreamSynchronize(st));
    FilterParams p{};
    int rc = fill_params(sys, integ, p.prop);
    if (rc != HB_OK) return rc;
    p.n = n; p.rec = (const double *)scratch; p.rec_cap = cap > 100000 ? 99968 : (int)cap;
    p.nacc = n_acc; p.status = status; p.t_eval = t_eval; p.m = m;
    p.inv_grid_dt = (ends[1] > ends[0]) ? (double)(m - 1) / (ends[1] - ends[0]) : 0.0;
    p.o = *opts; p.out = out; p.keep = keep;
    const long long blocks = (n + HB_FILT_WARPS - 1) / HB_FILT_WARPS;
    if (blocks > 2147483647LL) return HB_ERR_BADARG;
    if (integ->arith == HB_ARITH_PARITY)
        HB_CUDA_TRY(cudaFuncSetAttribute(k_record_filter<ArParity>, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_FILT_SMEM));
    else
        HB_CUDA_TRY(cudaFuncSetAttribute(k_record_filter<ArFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_FILT_SMEM));
    if (integ->arith == HB_ARITH_PARITY) k_record_filter<ArParity><<<(unsigned)blocks, 32 * HB_FILT_WARPS, HB_FILT_SMEM, st>>>(p);
    else k_record_filter<ArFast><<<(unsigned)blocks, 32 * HB_FILT_WARPS, HB_FILT_SMEM, st>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
