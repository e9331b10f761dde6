/* rr_oracle_count.cpp -- the CPU oracle compiled with an operation-counting scalar type (TEST / MEASUREMENT
 * INFRASTRUCTURE, not a product path).
 *
 * rr_oracle.c is included unchanged with `real` = creal (rr_opcount.hpp): every + - * / the oracle executes bumps
 *   rr_ops_total   all floating-point operations of the dense formulation (what MJX's dense path does), and
 *   rr_ops_useful  those whose operands are not structural zeros (a product with a zero operand, a sum of two zeros)
 *                  -- the sparsity-exact operation count that bench.py's FP32 roofline divides by.
 * sqrt / sin / cos / pow go through double and are not counted (a few hundred per substep). */
#include "rr_opcount.hpp"
extern "C" {
long long rr_ops_total = 0, rr_ops_useful = 0;
}
#define RR_REAL creal
extern "C" {
#include "rr_oracle.c"
void rro_ops_reset(void) { rr_ops_total = rr_ops_useful = 0; }
long long rro_ops_total(void) { return rr_ops_total; }
long long rro_ops_useful(void) { return rr_ops_useful; }
}
