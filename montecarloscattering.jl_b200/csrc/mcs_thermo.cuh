// thermo_calcs (/root/reference/src/thermo_calcs.jl:31-355) evaluated on the device-resident tallies.
//
// One block per grid zone.  The zone's slab of d2N_pf, (T+2) x (M+2) doubles in the reference's [jth, k] order, lives in a
// global scratch array: the largest bin grid the ABI admits (201 x 201) does not fit shared memory, the usual one
// (~40 x 110) would, but the slab is touched four times in total and stays in L2.  HBM-bound, 16 B per PSD cell.
//
//   A  d2N = 1e-99 + therm_d2N_pf                              :43, :133-164 (binned per crossing by the transport kernel)
//   B  every CR cell of psd boosted by its bin centre          :178-207      (FP64 atomics: order-dependent rounding)
//   C  norm_fac, scaling, d2N_pop                              :209-226
//   D  the three normalisation cases and the sums              :242-352
#pragma once
#include <cuda_runtime.h>

namespace mcs {

struct ThermoParams {
    int ng, T, M;
    double c, m, n0, gam0, beta0, temperature_K;
    double psd_mom_min, bpd_mom, psd_cos_fine, delta_cos, psd_theta_min, bpd_th;
    const double *gsf, *ux;                        // [ng + 2], 1-based as in DevParams
    const double *cos_center, *pt_center, *zone_pop;
    const double *psd, *therm_pf;
    const unsigned long long* ncross;
    double* slab;                                   // [(T+2)(M+2) ng] scratch
    double *P_par, *P_perp, *e_dens, *pop;          // [ng]
};

// get_psd_bins.jl:16-39, 73-97 without the warning counters of the transport kernel's copies
__device__ __forceinline__ int thermo_bin_momentum(const ThermoParams& P, double pt) {
    int bin = pt < P.psd_mom_min ? 0 : (int)trunc(log10(pt / P.psd_mom_min) * P.bpd_mom) + 1;
    return min(bin, P.M);
}
__device__ __forceinline__ int thermo_bin_angle(const ThermoParams& P, double px, double pt) {
    if (pt == 0.0) return 0;
    const double p_cos = -px / pt;
    int bin;
    if (p_cos < P.psd_cos_fine) bin = P.T - (int)trunc((p_cos + 1) / P.delta_cos);
    else {
        const double th = acos(p_cos);
        bin = th < P.psd_theta_min ? 0 : (int)trunc(log10(th / P.psd_theta_min) * P.bpd_th) + 1;
    }
    return min(bin, P.T);
}

// fixed-shape tree over the block: the same summation order on every run
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* red /*[N * 256]*/) {
    const int t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < N; q++) red[q * 256 + t] = v[q];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s)
#pragma unroll
            for (int q = 0; q < N; q++) red[q * 256 + t] += red[q * 256 + t + s];
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < N; q++) v[q] = red[q * 256];
    __syncthreads();
}

__global__ void __launch_bounds__(256) thermo_kernel(ThermoParams P) {
    __shared__ double red[3 * 256];
    const int i = blockIdx.x + 1, t = threadIdx.x;
    const int T2 = P.T + 2, M2 = P.M + 2, n = T2 * M2;
    double* d2N = P.slab + (size_t)n * (size_t)(i - 1);
    const double* thp = P.therm_pf + (size_t)n * (size_t)(i - 1);
    const double* psd = P.psd + (size_t)n * (size_t)(i - 1);
    const double cl = P.c, mc = P.m * cl, E0 = P.m * (cl * cl);
    const double g = P.gsf[i], b = P.ux[i] / cl;
    const bool no_therm = P.ncross[i - 1] == 0ull;
    const double zpop = P.zone_pop[i - 1];

    for (int q = t; q < n; q += 256) d2N[q] = 1.0e-99 + thp[q];
    __syncthreads();
    for (int q = t; q < n; q += 256) {  // psd index q = k + M2 * jt
        const int k = q % M2, jt = q / M2;
        if (k > P.M || jt > P.T) continue;
        const double cell = psd[q];
        if (cell <= 1.0e-66) continue;
        const double pt = P.pt_center[k], px = pt * P.cos_center[jt];
        const double etot = hypot(pt * cl, E0);
        const double pxX = g * (px - b * etot / cl);
        const double ptX = sqrt(pt * pt - px * px + pxX * pxX);
        const int kX = thermo_bin_momentum(P, ptX), jX = thermo_bin_angle(P, pxX, ptX);
        atomicAdd(&d2N[jX + T2 * kX], cell);
    }
    __syncthreads();

    double s1[1] = {0.0};
    for (int q = t; q < n; q += 256) { const double d = d2N[q]; if (d > 1.0e-66) s1[0] += d; }
    block_sum(s1, red);
    double nf = s1[0];
    if (no_therm && nf > 0) nf += P.n0 / P.ux[i];
    if (nf > 0) nf = zpop / nf;
    double s2[2] = {0.0, 0.0};  // population after scaling, running maximum (as a sum slot is not needed: use max below)
    for (int q = t; q < n; q += 256) {
        double d = d2N[q];
        if (d > 1.0e-66) { d *= nf; d2N[q] = d; }
        if (d > 1.0e-66) s2[0] += d;
        s2[1] = fmax(s2[1], d);
    }
    // the maximum through the same tree (fmax is order-independent)
    red[t] = s2[0]; red[256 + t] = s2[1];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (t < s) { red[t] += red[t + s]; red[256 + t] = fmax(red[256 + t], red[256 + t + s]); }
        __syncthreads();
    }
    const double pop = red[0], dmax = red[256];
    __syncthreads();

    const double dens = P.gam0 * P.beta0 * P.n0 / sqrt(g * g - 1.0);
    double base_par = 0.0, base_perp = 0.0, base_en = 0.0, norm = 0.0;
    bool sum_cells = true;
    if (dmax < 1.0e-66 && no_therm) {
        const double pl = pow(dens, 5.0 / 3.0) * 1.380649e-16 * P.temperature_K;
        base_par = 1.0 / 3.0 * pl; base_perp = 2.0 / 3.0 * pl; base_en = 1.5 * pl;
        sum_cells = false;
    } else if (no_therm) {
        double pl = pow(dens, 5.0 / 3.0) * 1.380649e-16 * P.temperature_K;
        pl *= 1.0 - pop / zpop;
        base_par = 1.0 / 3.0 * pl; base_perp = 2.0 / 3.0 * pl; base_en = 1.5 * pl;
        norm = dens / zpop;
    } else {
        norm = dens / zpop;
    }
    double s3[3] = {0.0, 0.0, 0.0};
    if (sum_cells)
        for (int q = t; q < n; q += 256) {  // slab index q = jt + T2 * k
            const int jt = q % T2, k = q / T2;
            if (k > P.M || jt > P.T) continue;
            const double d = d2N[q];
            if (d < 1.0e-66) continue;
            const double pt = P.pt_center[k];
            const double gt = hypot(1.0, pt / mc);
            const double vel = pt * cl / (mc * gt);
            const double pf = 1.0 / 3.0 * pt * vel * norm, ef = (gt - 1.0) * E0;
            const double c2 = P.cos_center[jt] * P.cos_center[jt];
            s3[0] += d * pf * c2; s3[1] += d * pf * (1.0 - c2); s3[2] += ef * d * norm;
        }
    block_sum(s3, red);
    if (t == 0) {
        if (P.P_par) P.P_par[i - 1] = base_par + s3[0];
        if (P.P_perp) P.P_perp[i - 1] = base_perp + s3[1];
        if (P.e_dens) P.e_dens[i - 1] = base_en + s3[2];
        if (P.pop) P.pop[i - 1] = pop;
    }
}

}  // namespace mcs
