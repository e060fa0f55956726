/* ho_poly.h -- polynomial Hamiltonian vector field of the oracle (TEST INFRASTRUCTURE). */
#ifndef HO_POLY_H
#define HO_POLY_H
#include "hiten_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
void ho_polyham_rhs(const ho_polyham *ham, const double *y, double *dy);
#ifdef __cplusplus
}
#endif
#endif
