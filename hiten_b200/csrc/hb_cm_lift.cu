// hb_cm_lift.cu -- centre-manifold seed lifting: plane point on a section -> 4-D state on the energy
// surface H = h0, one plane point per thread.
//
// Reference routines (paths relative to hiten/):
//   _CenterManifoldInterface.solve_missing_coord   algorithms/poincare/centermanifold/interfaces.py:212-268
//   _CenterManifoldInterface.lift_plane_point      interfaces.py:297-337   (build_constraint_dict / build_state :54-83)
//   solve_bracketed_brent                          algorithms/utils/rootfinding.py:92-190
//   _polynomial_evaluate / _poly_evaluate          algorithms/polynomial/operations.py:551-583, algebra.py:403-461
// The reference does one Python Brent solve per candidate seed (about 1 ms each); here every thread runs the same
// bracket expansion and the same Brent iteration on the same residual H(state) - h0 -- separately rounded IEEE
// operations in the reference's order, so roots are bit-identical.  The Hamiltonian is polynomial 0 of the term
// table (same record format as the gradient tables of hb_cm.cu); the table sits in shared memory (uniform reads),
// the per-thread power table x_v^e too, laid out [v][e][thread].
#include "hb_common.cuh"

namespace {

constexpr int LIFT_BLOCK = 128;

struct TermMeta {          // 16 bytes, as in hb_cm.cu
    double coef;
    unsigned long long ex; // exponents of (q1,q2,q3,p1,p2,p3) in bytes 0..5, degree in bits 48..63
};

struct LiftParams {
    hb_cm_lift_opts o;
    long long n;
    const double *pts;     // [n][2]
    double *states;        // [n][4] (q2,p2,q3,p3)
    int *ok;               // [n]
    const TermMeta *terms;
    int n_terms;
    int D;
};

struct Residual {
    const TermMeta *terms; // shared memory
    int n_terms, D1;
    double *pw;            // this thread's column of the shared power table
    double st[6];
    int idx;
    double h0;

    // power table of the five fixed coordinates (once per plane point)
    __device__ void prepare()
    {
#pragma unroll
        for (int v = 0; v < 6; ++v) set_powers(v, st[v]);
    }
    __device__ void set_powers(int v, double x)
    {
        double w = 1.0;
        pw[(v * D1) * LIFT_BLOCK] = 1.0;
        for (int e = 1; e < D1; ++e) {
            w = __dmul_rn(w, x);
            pw[(v * D1 + e) * LIFT_BLOCK] = w;
        }
    }
    // H(state with state[idx] = x) - h0: per-degree sums from 0.0 in term order, then summed by degree
    __device__ double operator()(double x)
    {
        set_powers(idx, x);
        double total = 0.0, acc = 0.0;
        int dcur = -1;
        for (int i = 0; i < n_terms; ++i) {
            const TermMeta tm = terms[i];
            const int d = (int)(tm.ex >> 48);
            if (d != dcur) {
                if (dcur >= 0) total = __dadd_rn(total, acc);
                acc = 0.0;
                dcur = d;
            }
            double t = 1.0;
#pragma unroll
            for (int v = 0; v < 6; ++v) {
                const int e = (int)((tm.ex >> (8 * v)) & 0xffu);
                t = __dmul_rn(t, pw[(v * D1 + e) * LIFT_BLOCK]);          // exponent 0 -> * 1.0 (exact)
            }
            acc = __dadd_rn(acc, __dmul_rn(tm.coef, t));
        }
        if (dcur >= 0) total = __dadd_rn(total, acc);
        return __dsub_rn(total, h0);
    }
};

// solve_bracketed_brent (rootfinding.py:92-190); returns false for None
__device__ bool brent(Residual &f, double a, double b, double xtol, int max_iter, double &root)
{
    double fa = f(a), fb = f(b);
    if (fa == 0.0) { root = a; return true; }
    if (fb == 0.0) { root = b; return true; }
    if (__dmul_rn(fa, fb) > 0.0) return false;
    double c = a, fc = fa, d = __dsub_rn(b, a), e = d;
    const double eps = 2.220446049250313e-16;
    double tol, m;
    for (int it = 0; it < max_iter; ++it) {
        if (fb == 0.0) { root = b; return true; }
        if (__dmul_rn(fb, fc) > 0.0) { c = a; fc = fa; d = __dsub_rn(b, a); e = d; }
        if (fabs(fc) < fabs(fb)) {
            a = b; b = c; c = a;
            fa = fb; fb = fc; fc = fa;
        }
        tol = __dadd_rn(__dmul_rn(__dmul_rn(2.0, eps), fabs(b)), __dmul_rn(0.5, xtol));
        m = __dmul_rn(0.5, __dsub_rn(c, b));
        if (fabs(m) <= tol) { root = b; return true; }
        if (fabs(e) >= tol && fabs(fa) > fabs(fb)) {
            const double s = __ddiv_rn(fb, fa);
            double p, q;
            if (a == c) {
                p = __dmul_rn(__dmul_rn(2.0, m), s);
                q = __dsub_rn(1.0, s);
            } else {
                const double q_ = __ddiv_rn(fa, fc), r = __ddiv_rn(fb, fc);
                p = __dmul_rn(s, __dsub_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, m), q_), __dsub_rn(q_, r)),
                                           __dmul_rn(__dsub_rn(b, a), __dsub_rn(r, 1.0))));
                q = __dmul_rn(__dmul_rn(__dsub_rn(q_, 1.0), __dsub_rn(r, 1.0)), __dsub_rn(s, 1.0));
            }
            if (p > 0.0) q = -q; else p = -p;
            const double lim1 = __dsub_rn(__dmul_rn(__dmul_rn(3.0, m), q), fabs(__dmul_rn(tol, q)));
            const double lim2 = fabs(__dmul_rn(e, q));
            if (__dmul_rn(2.0, p) < (lim2 < lim1 ? lim2 : lim1)) { e = d; d = __ddiv_rn(p, q); }
            else { d = m; e = m; }
        } else { d = m; e = m; }
        a = b; fa = fb;
        if (fabs(d) > tol) b = __dadd_rn(b, d);
        else b = __dadd_rn(b, (m > 0.0 ? tol : -tol));
        fb = f(b);
    }
    tol = __dadd_rn(__dmul_rn(__dmul_rn(2.0, eps), fabs(b)), __dmul_rn(0.5, xtol));
    m = __dmul_rn(0.5, __dsub_rn(c, b));
    if (fabs(m) <= tol || fb == 0.0) { root = b; return true; }
    return false;
}

// solve_missing_coord (interfaces.py:212-268)
__device__ bool solve_missing(Residual &f, const hb_cm_lift_opts &o, double &root)
{
    if (f(0.0) > 0.0) return false;
    double b = o.initial_guess, r_b = f(b);
    int n_expand = 0;
    while (r_b <= 0.0 && n_expand < o.max_expand) { b = __dmul_rn(b, o.expand_factor); r_b = f(b); ++n_expand; }
    if (r_b > 0.0) return brent(f, 0.0, b, o.xtol, o.max_iter, root);
    if (o.symmetric) {
        double a_neg = -o.initial_guess, r_a = f(a_neg);
        n_expand = 0;
        while (r_a <= 0.0 && n_expand < o.max_expand) { a_neg = __dmul_rn(a_neg, o.expand_factor); r_a = f(a_neg); ++n_expand; }
        if (r_a > 0.0) return brent(f, a_neg, 0.0, o.xtol, o.max_iter, root);
    }
    return false;
}

__global__ void __launch_bounds__(LIFT_BLOCK) k_cm_lift(const LiftParams p)
{
    extern __shared__ __align__(16) unsigned char lift_smem[];
    TermMeta *terms = (TermMeta *)lift_smem;
    double *pw_base = (double *)(lift_smem + (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15));
    for (int i = threadIdx.x; i < p.n_terms; i += blockDim.x) terms[i] = p.terms[i];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    // variable slots: q1 0, q2 1, q3 2, p1 3, p2 4, p3 5; sections 0 q2, 1 p2, 2 q3, 3 p3
    const int sec = p.o.section;
    const int pa = (sec >= 2) ? 1 : 2, pb = (sec >= 2) ? 4 : 5;          // plane coordinates (q2,p2) or (q3,p3)
    const int miss = (sec == 0) ? 4 : (sec == 1) ? 1 : (sec == 2) ? 5 : 2;  // q2->p2, p2->q2, q3->p3, p3->q3
    Residual f;
    f.terms = terms; f.n_terms = p.n_terms; f.D1 = p.D + 1; f.pw = pw_base + threadIdx.x; f.idx = miss; f.h0 = p.o.h0;
#pragma unroll
    for (int v = 0; v < 6; ++v) f.st[v] = 0.0;
    const double a = p.pts[2 * i], b = p.pts[2 * i + 1];
    f.st[pa] = a; f.st[pb] = b;
    f.prepare();
    double root = 0.0;
    const bool ok = solve_missing(f, p.o, root);
    double out[4] = {0.0, 0.0, 0.0, 0.0};                                  // (q2, p2, q3, p3)
    if (ok) {
        double full[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        full[pa] = a; full[pb] = b; full[miss] = root;
        out[0] = full[1]; out[1] = full[4]; out[2] = full[2]; out[3] = full[5];
    }
    p.ok[i] = ok ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) p.states[4 * i + c] = out[c];
}

}  // namespace

extern "C" int hb_cm_lift(const hb_polyham *H, const hb_cm_lift_opts *opts, int64_t n, const double *plane_pts,
                          double *states, int32_t *ok, void *stream)
{
    if (!H || !opts || n < 0) return HB_ERR_BADARG;
    if (H->n_dof != 3 || H->max_deg < 0 || H->max_deg > 30) return HB_ERR_UNSUPPORTED;
    if (opts->section < 0 || opts->section > 3 || opts->max_expand < 0 || opts->max_iter < 0) return HB_ERR_BADARG;
    if (n > 0 && (!plane_pts || !states || !ok || !H->terms)) return HB_ERR_BADARG;
    if (n == 0) return HB_OK;
    LiftParams p{};
    p.o = *opts; p.n = n; p.pts = plane_pts; p.states = states; p.ok = ok;
    p.terms = (const TermMeta *)H->terms;
    p.n_terms = (int)(H->ptr[1] - H->ptr[0]);
    p.terms += H->ptr[0];
    p.D = H->max_deg;
    const size_t smem = (((size_t)p.n_terms * sizeof(TermMeta) + 15) & ~(size_t)15) +
                        (size_t)6 * (p.D + 1) * LIFT_BLOCK * sizeof(double);
    if (smem > 227 * 1024) return HB_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        HB_CUDA_TRY(cudaFuncSetAttribute(k_cm_lift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cm_lift<<<(unsigned)((n + LIFT_BLOCK - 1) / LIFT_BLOCK), LIFT_BLOCK, smem, (cudaStream_t)stream>>>(p);
    HB_CUDA_TRY(cudaGetLastError());
    return HB_OK;
}
