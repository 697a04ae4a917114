// jx_inst_other.cu -- total-energy Euler (2D), AdvDiff (2D/3D), ShallowWater (2D), CompEuler theta + LES source (3D) instantiations
#include "jx_launch.cuh"

namespace jx {

#define JX_EN(NGL, VISC) make_node_set<2, NGL, EulerEnergy<2, false, false>, VISC>(JX_EQ_EULER_ENERGY, 0, 0)
#define JX_AD(NSD, NGL, VISC) make_node_set<NSD, NGL, AdvDiff<NSD, false, false>, VISC>(JX_EQ_ADVDIFF, 0, 0)
#define JX_SW(NGL, VISC) make_node_set<2, NGL, ShallowWater<2, false, false>, VISC>(JX_EQ_SHALLOW_WATER, 0, 0)
#define JX_LES(NGL, VISC) make_node_set<3, NGL, EulerThetaLES<true>, VISC>(JX_EQ_EULER_THETA_LES, 0, 0)
// VISC: 0 inviscid, 1 AV, 2 SGS closure (SMAG / VREM: jx_set_sgs; the Euler sets only)
#define JX_ROW(NGL) \
    JX_EN(NGL, 0), JX_EN(NGL, 1), JX_EN(NGL, 2), JX_AD(2, NGL, 0), JX_AD(2, NGL, 1), JX_AD(3, NGL, 0), \
    JX_AD(3, NGL, 1), JX_SW(NGL, 0), JX_SW(NGL, 1)

const KernelSet *lookup_other(int nsd, int ngl, int eq_id, int lvisc, int variant) {
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8), JX_LES(3, 0), JX_LES(3, 1), JX_LES(3, 2), JX_LES(5, 0), JX_LES(5, 1), JX_LES(5, 2)};
    for (const KernelSet &k : table)
        if (k.nsd == nsd && k.ngl == ngl && k.eq_id == eq_id && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
