// jx_inst_other.cu -- total-energy Euler (2D), AdvDiff (2D/3D), ShallowWater (2D), CompEuler theta + LES source (3D) instantiations
#include "jx_launch.cuh"

namespace jx {

#define JX_EN(NGL, VISC) make_node_set<2, NGL, EulerEnergy<2, false, false>, VISC>(JX_EQ_EULER_ENERGY, 0, 0)
#define JX_AD(NSD, NGL, VISC) make_node_set<NSD, NGL, AdvDiff<NSD, false, false>, VISC>(JX_EQ_ADVDIFF, 0, 0)
#define JX_SW(NGL, VISC) make_node_set<2, NGL, ShallowWater<2, false, false>, VISC>(JX_EQ_SHALLOW_WATER, 0, 0)
#define JX_LES(NGL, VISC) make_node_set<3, NGL, EulerThetaLES<true>, VISC>(JX_EQ_EULER_THETA_LES, 0, 0)
#define JX_ROW(NGL) \
    JX_EN(NGL, false), JX_EN(NGL, true), JX_AD(2, NGL, false), JX_AD(2, NGL, true), JX_AD(3, NGL, false), \
    JX_AD(3, NGL, true), JX_SW(NGL, false), JX_SW(NGL, true)

const KernelSet *lookup_other(int nsd, int ngl, int eq_id, int lvisc, int variant) {
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8), JX_LES(3, false), JX_LES(3, true), JX_LES(5, false), JX_LES(5, true)};
    for (const KernelSet &k : table)
        if (k.nsd == nsd && k.ngl == ngl && k.eq_id == eq_id && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
