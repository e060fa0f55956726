#include "ho_poly.h"
void ho_polyham_rhs(const ho_polyham *ham, const double *y, double *dy) { (void)ham; (void)y; (void)dy; }
