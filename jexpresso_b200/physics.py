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
