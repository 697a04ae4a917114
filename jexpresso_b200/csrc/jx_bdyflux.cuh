// jx_bdyflux.cuh -- boundary fluxes with the Monin-Obukhov wall model (second part):
//   build_custom_bcs_neumann!(::NSD_3D)                 src/kernel/boundaryconditions/BCs.jl:655-816  (bdy_fluxes, !bulk_fluxes, dry)
//   CM_MOST!, _surface_scales_dry, psi_m/psi_h, obukhov_length   src/kernel/physics/CM_MOST.jl:69-100, 144-152, 224-261
//   compute_surface_integral!, DSS_surface_integral!    src/kernel/boundaryconditions/surface_integral.jl:1-29
//   RHS .+= S_flux                                      rhs.jl:674-689 (before DSS_global_RHS! and the division by M)
// Two tiny kernels behind the element kernels: one thread per node of a "MOST" face evaluates the wall model and the
// quadrature-weighted flux; one thread per unique wall node sums its face contributions in the reference's order (face
// ascending, then i outer / j inner) and adds them to the right-hand side.  log / atan / pow are CUDA's (Julia's in the
// reference, libm's in the oracle: <= 2 ulp apart): this path is held to 1e-12 against the oracle, not to bit equality.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jx {

struct MostArgs {
    const double *u, *qe, *coords;   // state (after the Dirichlet projection) [neqs][npoin], reference state, coords [3][npoin]
    const int32_t *ip1, *ipsfc;      // per wall-face node: inside point connijk[e,i,j,ifirst_wall_node], surface point connijk[e,i,j,1]
    const double *normal;            // [nwn][3]
    const double *wJ;                // [nwn]: omega_i * omega_j * Jef
    double *sface;                   // [nwn][4]: S_face of the three momentum equations and the theta equation
    int64_t npoin;
    int nwn, lpert;
    double karman, z0_m, z0_h, cp, g, delta_hf, user_heatflux;
};

__device__ __forceinline__ double most_psi_m(double zeta) {
    if (zeta < 0) {
        const double x = pow(1 - 16.0 * zeta, 0.25);
        return 2 * log((1 + x) / 2) + log((1 + x * x) / 2) - 2 * atan(x) + 3.141592653589793 / 2;
    }
    return -5.0 * zeta;
}
__device__ __forceinline__ double most_psi_h(double zeta) {
    if (zeta < 0) {
        const double y = pow(1 - 16.0 * zeta, 0.5);
        return 2 * log((1 + y) / 2);
    }
    return -5.0 * zeta;
}
__device__ __forceinline__ double most_obukhov_length(double u_star, double T_ref, double Q_H, double cp, double karman, double g) {
    if (fabs(Q_H) < 1e-6) return 1e6;
    return -(u_star * u_star * u_star) * T_ref * cp / (karman * g * Q_H);
}

static __global__ void k_most_faces(const __grid_constant__ MostArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nwn) return;
    const int64_t N = a.npoin;
    const int64_t ip1 = a.ip1[t], isf = a.ipsfc[t];
    double rho, u_in, v_in, w_in, th_in, th_sfc;
    if (!a.lpert) {
        rho = a.u[ip1];
        u_in = a.u[ip1 + N * 1] / rho; v_in = a.u[ip1 + N * 2] / rho; w_in = a.u[ip1 + N * 3] / rho;
        th_in = a.u[ip1 + N * 4] / rho;
        th_sfc = a.u[isf + N * 4] / a.u[isf];
    } else {
        rho = a.u[ip1] + a.qe[ip1];
        u_in = (a.u[ip1 + N * 1] + a.qe[ip1 + N * 1]) / rho;
        v_in = (a.u[ip1 + N * 2] + a.qe[ip1 + N * 2]) / rho;
        w_in = (a.u[ip1 + N * 3] + a.qe[ip1 + N * 3]) / rho;
        th_in = (a.u[ip1 + N * 4] + a.qe[ip1 + N * 4]) / rho;
        th_sfc = (a.u[isf + N * 4] + a.qe[isf + N * 4]) / (a.u[isf] + a.qe[isf]);
    }
    const double nx = a.normal[3 * t], ny = a.normal[3 * t + 1], nz = a.normal[3 * t + 2];
    const double vproj = u_in * nx + v_in * ny + w_in * nz;
    u_in = u_in - vproj * nx; v_in = v_in - vproj * ny; w_in = w_in - vproj * nz;
    const double dx = a.coords[ip1] - a.coords[isf];
    const double dy = a.coords[N + ip1] - a.coords[N + isf];
    const double dz = a.coords[2 * N + ip1] - a.coords[2 * N + isf];
    const double z_ref = fabs(dx * nx + dy * ny + dz * nz);
    // CM_MOST! (dry): friction velocity and temperature scale by fixed-point iteration on the Obukhov length
    const double u_mag = sqrt(u_in * u_in + v_in * v_in + w_in * w_in);
    const double karman = a.karman, cp = a.cp, g = a.g, z0_m = a.z0_m, z0_h = a.z0_h;
    double u_star = karman * u_mag / log(z_ref / z0_m);
    double theta_star = karman * (th_in - th_sfc) / log(z_ref / z0_h);
    double Q_H = -rho * cp * u_star * theta_star;
    double L = most_obukhov_length(u_star, th_in, Q_H, cp, karman, g);
    for (int it = 0; it < 20; ++it) {
        const double zeta = z_ref / L, zeta0_m = z0_m / L, zeta0_h = z0_h / L;
        const double u_star_new = karman * u_mag / (log(z_ref / z0_m) - most_psi_m(zeta) + most_psi_m(zeta0_m));
        const double theta_star_new = karman * (th_in - th_sfc) / (log(z_ref / z0_h) - most_psi_h(zeta) + most_psi_h(zeta0_h));
        const double Q_H_new = -rho * cp * u_star_new * theta_star_new;
        const double L_new = most_obukhov_length(u_star_new, th_in, Q_H_new, cp, karman, g);
        const double err = fabs(L_new - L) / fmax(fabs(L), fabs(L_new));
        u_star = u_star_new; theta_star = theta_star_new; Q_H = Q_H_new; L = L_new;
        if (err < 1e-4) break;
    }
    const double tau_mag = rho * (u_star * u_star);
    const double F1 = -tau_mag * (u_in / (u_mag + 2.22e-16));
    const double F2 = -tau_mag * (v_in / (u_mag + 2.22e-16));
    const double F3 = -tau_mag * (w_in / (u_mag + 2.22e-16));
    const double wth = -u_star * theta_star;
    const double F4 = wth * (1.0 - a.delta_hf) + a.user_heatflux * a.delta_hf;
    const double wJ = a.wJ[t];
    a.sface[4 * t + 0] = wJ * F1;      // S_face starts from zero every evaluation (resetbdyfluxToZero!, rhs.jl:105-109)
    a.sface[4 * t + 1] = wJ * F2;
    a.sface[4 * t + 2] = wJ * F3;
    a.sface[4 * t + 3] = wJ * F4;
}

struct FluxAddArgs {
    double *rhs;                 // [neqs][npoin]: RHS before DSS_global_RHS! (deterministic mode) or the M^-1-scaled scatter target
    const double *Minv;          // non-null: the contributions are scaled like everything else the unordered mode scatters
    const double *sface;         // [nwn][4]
    const int32_t *node, *ptr, *hit;   // unique wall nodes, their hit ranges, wall-face-node index per hit (reference order)
    int64_t npoin;
    int nnode;
};

static __global__ void k_bdy_flux_add(const __grid_constant__ FluxAddArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nnode) return;
    const int64_t ip = a.node[t];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int h = a.ptr[t]; h < a.ptr[t + 1]; ++h) {
        const int w = a.hit[h];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = s[k] + a.sface[4 * w + k];     // DSS_surface_integral!: S_flux[ip, ieq] += S_face[...]
    }
    const double mi = a.Minv ? a.Minv[ip] : 1.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double *p = a.rhs + (size_t)(1 + k) * a.npoin + ip;
        *p = a.Minv ? *p + s[k] * mi : *p + s[k];                            // RHS .+= S_flux
    }
}

}  // namespace jx
