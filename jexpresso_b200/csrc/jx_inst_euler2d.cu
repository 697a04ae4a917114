// jx_inst_euler2d.cu -- CompEuler θ-form, 2D: kernel instantiations (problems/CompEuler/theta hooks)
#include "jx_launch.cuh"

namespace jx {

#define JX_SET(NGL, PERT, POW, VISC) make_node_set<2, NGL, EulerTheta<2, PERT, POW>, VISC>(JX_EQ_EULER_THETA, PERT, POW)
// VISC: 0 inviscid, 1 AV, 2 SGS closure (SMAG / VREM: jx_set_sgs)
#define JX_ROW(NGL) \
    JX_SET(NGL, false, false, 0), JX_SET(NGL, false, false, 1), JX_SET(NGL, false, true, 0), \
    JX_SET(NGL, false, true, 1), JX_SET(NGL, true, false, 0), JX_SET(NGL, true, false, 1), \
    JX_SET(NGL, true, true, 0), JX_SET(NGL, true, true, 1), \
    JX_SET(NGL, false, false, 2), JX_SET(NGL, false, true, 2), JX_SET(NGL, true, false, 2), JX_SET(NGL, true, true, 2)

const KernelSet *lookup_euler_theta_2d(int ngl, int lpert, int jxpow, int lvisc, int variant) {
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8)};
    for (const KernelSet &k : table)
        if (k.ngl == ngl && k.lpert == lpert && k.jxpow == jxpow && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
