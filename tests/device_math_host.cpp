// Host build of csrc/mcs_math.cuh so its IEEE-exact arithmetic can be checked against libm without a GPU.
#include "../montecarloscattering.jl_b200/csrc/mcs_math.cuh"
extern "C" {
void t_sincos(const double* x, double* s, double* c, long n) { for (long i = 0; i < n; i++) mcs::sincos_bf(x[i], &s[i], &c[i]); }
void t_asin(const double* x, double* y, long n) { for (long i = 0; i < n; i++) y[i] = mcs::asin_bf(x[i]); }
}
