// jx_inst_euler3d.cu -- CompEuler θ-form, 3D: kernel instantiations (problems/CompEuler/3d hooks)
#include "jx_launch.cuh"

namespace jx {

// VISC: 0 inviscid, 1 AV, 2 SGS closure (SMAG / VREM: jx_set_sgs)
#define JX_SET(NGL, PERT, POW, VISC) make_node_set<3, NGL, EulerTheta<3, PERT, POW>, VISC>(JX_EQ_EULER_THETA, PERT, POW)
#define JX_ROW(NGL) \
    JX_SET(NGL, false, false, 0), JX_SET(NGL, false, false, 1), JX_SET(NGL, false, true, 0), \
    JX_SET(NGL, false, true, 1), JX_SET(NGL, true, false, 0), JX_SET(NGL, true, false, 1), \
    JX_SET(NGL, true, true, 0), JX_SET(NGL, true, true, 1), \
    JX_SET(NGL, false, false, 2), JX_SET(NGL, false, true, 2), JX_SET(NGL, true, false, 2), JX_SET(NGL, true, true, 2)

#define JX_TSET(NGL, ZW, PW, VAR, PERT, POW) make_team_set<NGL, EulerTheta<3, PERT, POW>, ZW, PW>(JX_EQ_EULER_THETA, PERT, POW, VAR)
#define JX_TROW(NGL, ZW, PW, VAR) \
    JX_TSET(NGL, ZW, PW, VAR, false, false), JX_TSET(NGL, ZW, PW, VAR, false, true), JX_TSET(NGL, ZW, PW, VAR, true, false), \
    JX_TSET(NGL, ZW, PW, VAR, true, true)

#define JX_TVQSET(PERT, POW) make_team_visc_quad_set<5, EulerTheta<3, PERT, POW>, 2, 2>(JX_EQ_EULER_THETA, PERT, POW, 13)
#define JX_TRISET(PERT, POW) make_tri_set<8, EulerTheta<3, PERT, POW>>(JX_EQ_EULER_THETA, PERT, POW, 12)

const KernelSet *lookup_euler_theta_3d(int ngl, int lpert, int jxpow, int lvisc, int variant) {
#ifdef JX_MIN_BUILD   // kernel experiments: nop 4, TOTAL, jx_pow only -- generic, team (9), tri (12), viscous team pass
    static const KernelSet table[] = {JX_SET(5, false, true, 0), JX_SET(5, false, true, 1), JX_SET(5, false, true, 2), JX_TSET(5, 2, 2, 9, false, true),
                                      JX_SET(8, false, true, 0), JX_TRISET(false, true), JX_TVQSET(false, true)};
#else
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8), JX_TROW(3, 3, 1, 8), JX_TROW(5, 2, 1, 8), JX_TROW(5, 2, 2, 9),
                                      JX_TRISET(false, false), JX_TRISET(false, true), JX_TRISET(true, false), JX_TRISET(true, true),
                                      JX_TVQSET(false, false), JX_TVQSET(false, true), JX_TVQSET(true, false), JX_TVQSET(true, true)};
#endif
    for (const KernelSet &k : table)
        if (k.ngl == ngl && k.lpert == lpert && k.jxpow == jxpow && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
