// jx_inst_euler2d.cu -- CompEuler θ-form, 2D: kernel instantiations (problems/CompEuler/theta hooks)
#include "jx_launch.cuh"

namespace jx {

#define JX_SET(NGL, PERT, POW, VISC) make_node_set<2, NGL, EulerTheta<2, PERT, POW>, VISC>(JX_EQ_EULER_THETA, PERT, POW)
#define JX_ROW(NGL) \
    JX_SET(NGL, false, false, false), JX_SET(NGL, false, false, true), JX_SET(NGL, false, true, false), \
    JX_SET(NGL, false, true, true), JX_SET(NGL, true, false, false), JX_SET(NGL, true, false, true), \
    JX_SET(NGL, true, true, false), JX_SET(NGL, true, true, true)

const KernelSet *lookup_euler_theta_2d(int ngl, int lpert, int jxpow, int lvisc, int variant) {
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8)};
    for (const KernelSet &k : table)
        if (k.ngl == ngl && k.lpert == lpert && k.jxpow == jxpow && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
