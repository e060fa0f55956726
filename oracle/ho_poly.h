/* ho_poly.h -- polynomial Hamiltonian vector field + centre-manifold Poincare map of the oracle
 * (TEST INFRASTRUCTURE, see hiten_oracle.h). */
#ifndef HO_POLY_H
#define HO_POLY_H
#include "hiten_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Sparse real term table of the six partials dH/d(q1,q2,q3,p1,p2,p3) (the reference's jac_H), terms in
 * the reference's evaluation order: degree ascending, packed index ascending, zero coefficients skipped
 * (algorithms/polynomial/operations.py:551-583, algebra.py:403-461). */
struct ho_polyham {
    int n_dof;             /* 3 */
    int max_deg;           /* largest exponent that occurs */
    int64_t ptr[7];        /* CSR: terms of partial p are [ptr[p], ptr[p+1]) */
    const int32_t *deg;    /* [T] homogeneous degree of the term (summation is grouped by degree) */
    const double *coef;    /* [T] */
    const int32_t *exp;    /* [T][6] exponents of (q1,q2,q3,p1,p2,p3) */
};

double ho_poly_eval_partial(const ho_polyham *ham, int p, const double *point6);
void ho_polyham_rhs(const ho_polyham *ham, const double *y, double *dy);   /* dynamics/hamiltonian.py:35-90 */

void ho_tao_update(const ho_polyham *ham, double *q_ext, double dt, int order, double c_omega);

/* _ExtendedSymplectic.integrate (algorithms/integrators/symplectic.py:877-1004) on the SIGNED grid t_vals[m] the
 * low-level routines see (t_vals * fwd): _integrate_symplectic (:564-653) -> traj[m][6]; the extended state is carried
 * across the grid. */
int ho_symplectic_dense(const ho_polyham *ham, const double *y0, const double *t_vals, int m, int order,
                        double c_omega, double *traj);
/* _integrate_symplectic_until_event (:657-782) + _hermite_refine_event_symplectic (:282-367).  Returns 1 on a hit
 * (t_hit / y_hit refined, *n_rows = trajectory rows written before the event) or 0 (t_hit = t_vals[m-1], y_hit = last
 * row, *n_rows = m).  traj[m][6] may be NULL. */
int ho_symplectic_event(const ho_polyham *ham, const ho_event *ev, const double *y0, const double *t_vals, int m,
                        int order, double c_omega, double *t_hit, double *y_hit, double *traj, int *n_rows);

/* _poincare_map (algorithms/poincare/centermanifold/backend.py:314-382): seeds[n][4] = (q2,p2,q3,p3);
 * section: 0 q2, 1 p2, 2 q3, 3 p3; out[n][4], t_out[n], flags[n] (int64). */
int ho_cm_poincare_map(const ho_polyham *ham, const double *seeds, int64_t n, double dt, int order, int max_steps,
                       int use_symplectic, int section, double c_omega, int64_t *flags, double *out, double *t_out,
                       int n_threads);

/* Seed lifting: _CenterManifoldInterface.solve_missing_coord / lift_plane_point
 * (algorithms/poincare/centermanifold/interfaces.py:212-268, 297-337) with solve_bracketed_brent
 * (algorithms/utils/rootfinding.py:92-190).  `H` holds the Hamiltonian itself as polynomial 0 of the table.
 * section: 0 q2, 1 p2, 2 q3, 3 p3; pts[n][2] plane points; states[n][4] = (q2,p2,q3,p3); ok[n] (int64). */
int ho_cm_solve_missing(const ho_polyham *H, const double *fixed6, int solve_idx, double h0, double initial_guess,
                        double expand_factor, int max_expand, int symmetric, double xtol, double *root);
int ho_cm_lift(const ho_polyham *H, int section, const double *pts, int64_t n, double h0, double initial_guess,
               double expand_factor, int max_expand, int symmetric, double xtol, int64_t *ok, double *states);
#ifdef __cplusplus
}
#endif
#endif
