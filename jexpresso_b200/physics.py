"""Physical constants and the equation-set / scheme ids shared by host and device.

PhysicalConst restates src/kernel/physics/globalConstantsPhysics.jl:3-62 (only the
fields the RHS path reads); the perfect-gas law is constitutiveLaw.jl:22-24,
``P = C0 (ρθ)^γ`` with ``C0 = Rair^γ / pref^(γ-1)``.
"""
from __future__ import annotations

from dataclasses import dataclass

# equation-set ids understood by jx_set_problem (include/jexrhs.h)
EQ_EULER_THETA = 0      # CompEuler θ-form (problems/CompEuler/3d, problems/CompEuler/theta)
EQ_EULER_ENERGY = 1     # CompEuler total energy (problems/CompEuler/kelvinHelmholtzChan2022)
EQ_ADVDIFF = 2          # AdvDiff (problems/AdvDiff/kopriva, 3d_periodic)
EQ_SHALLOW_WATER = 3    # ShallowWater (problems/ShallowWater/SoliWaveIsland)
EQ_EULER_THETA_LES = 4  # CompEuler θ-form with the sponge / Coriolis / geostrophic source of problems/CompEuler/LESICP1

SCHEME_CK2N54 = 0       # CarpenterKennedy2N54
SCHEME_SSPRK54 = 1
SCHEME_SSPRK33 = 2

BC_SKIP = 0             # periodic* tags: no Dirichlet projection (BCs.jl:621-623)
BC_FREE_SLIP = 1        # user_bc_dirichlet! of CompEuler (free-slip projection)


@dataclass(frozen=True)
class PhysicalConst:
    Rair: float = 287.0
    cp: float = 1004.0
    cv: float = 718.0
    pref: float = 101200.0
    g: float = 9.80616
    # SGS closures (globalConstantsPhysics.jl:15-31; read by allocate_SGS, sgsStructs.jl:77-120)
    mu_mol: float = 1.8e-5
    kappa_mol: float = 2.4e-5
    Sc_t: float = 0.7
    Pr_t: float = 0.7
    Ri_crit: float = 0.25
    C_s: float = 0.21

    @property
    def gamma(self):
        return self.cp / self.cv

    @property
    def cpoverR(self):
        return self.cp / self.Rair

    @property
    def C0(self):
        return (self.Rair ** self.gamma) / self.pref ** (self.gamma - 1.0)

    def packed(self):
        """Packed array handed to jx_set_problem: [C0, γ, g, Rair, cp, cv, pref, γ-1]."""
        return [self.C0, self.gamma, self.g, self.Rair, self.cp, self.cv, self.pref, self.gamma - 1.0]


    def sgs_packed(self):
        """Constants handed to jx_set_sgs: [Pr_t, Sc_t, mu_mol, kappa_mol, Ri_crit, C_s]."""
        return [self.Pr_t, self.Sc_t, self.mu_mol, self.kappa_mol, self.Ri_crit, self.C_s]


VISC_AV, VISC_SMAG, VISC_VREM = 0, 1, 2     # inputs[:visc_model] = AV() | SMAG() | VREM()  (jx_set_sgs)


def swe_packed(g=9.81, h_wet=1.0e-3, cone_height=0.93, sigma_dry=25.0, cone_xc=12.5, cone_yc=0.0, cone_rc=3.6):
    """Packed constants of the ShallowWater functor (problems/ShallowWater/SoliWaveIsland/user_flux.jl:41-42,
    user_source.jl:1-5): phys[9] = cone height, [10] = dry-node relaxation rate, [11] = g, [12] = wet/dry film depth,
    [13],[14] = cone centre, [15] = cone radius."""
    ph = [0.0] * 16
    ph[9], ph[10], ph[11], ph[12], ph[13], ph[14], ph[15] = cone_height, sigma_dry, g, h_wet, cone_xc, cone_yc, cone_rc
    return ph


def advdiff_packed(u=0.5, v=1.0, w=0.0):
    """Packed constants of the AdvDiff functor: the constant wind of problems/AdvDiff/*/user_flux.jl in phys[8..10]."""
    ph = [0.0] * 16
    ph[8], ph[9], ph[10] = u, v, w
    return ph


def les_packed(zmax, lsponge=True, zsponge=0.0, f=1.0e-4, alpha=0.5, base=None):
    """Packed constants of the LESICP1 functor (problems/CompEuler/LESICP1/user_source.jl:30-103): PhysicalConst in [0..7],
    [8] = inputs[:lsponge], [9] = inputs[:zsponge], [10] = zmax of the mesh, [11] = Coriolis parameter, [12] = sponge alpha."""
    ph = list((base or PhysicalConst()).packed()) + [0.0] * 8
    ph[8], ph[9], ph[10], ph[11], ph[12] = (1.0 if lsponge else 0.0), float(zsponge), float(zmax), float(f), float(alpha)
    return ph
