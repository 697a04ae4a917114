// jx_inst_euler3d.cu -- CompEuler θ-form, 3D: kernel instantiations (problems/CompEuler/3d hooks)
#include "jx_launch.cuh"

namespace jx {

#define JX_SET(NGL, PERT, POW, VISC) make_node_set<3, NGL, EulerTheta<3, PERT, POW>, VISC>(JX_EQ_EULER_THETA, PERT, POW)
#define JX_ROW(NGL) \
    JX_SET(NGL, false, false, false), JX_SET(NGL, false, false, true), JX_SET(NGL, false, true, false), \
    JX_SET(NGL, false, true, true), JX_SET(NGL, true, false, false), JX_SET(NGL, true, false, true), \
    JX_SET(NGL, true, true, false), JX_SET(NGL, true, true, true)

#define JX_PSET(NGL, PERT, POW, EXACT) make_pencil_set<NGL, EulerTheta<3, PERT, POW>, EXACT>(JX_EQ_EULER_THETA, PERT, POW)
#define JX_PROW(NGL) \
    JX_PSET(NGL, false, false, true), JX_PSET(NGL, false, true, true), JX_PSET(NGL, true, false, true), \
    JX_PSET(NGL, true, true, true), JX_PSET(NGL, false, false, false), JX_PSET(NGL, false, true, false), \
    JX_PSET(NGL, true, false, false), JX_PSET(NGL, true, true, false)

#define JX_WSET(NGL, PERT, POW, EXACT) make_wpencil_set<NGL, EulerTheta<3, PERT, POW>, EXACT>(JX_EQ_EULER_THETA, PERT, POW)
#define JX_WROW(NGL) \
    JX_WSET(NGL, false, false, true), JX_WSET(NGL, false, true, true), JX_WSET(NGL, true, false, true), \
    JX_WSET(NGL, true, true, true), JX_WSET(NGL, false, false, false), JX_WSET(NGL, false, true, false), \
    JX_WSET(NGL, true, false, false), JX_WSET(NGL, true, true, false)

#define JX_GSET(NGL, EPB, VAR, PERT, POW) make_gpencil_set<NGL, EulerTheta<3, PERT, POW>, EPB>(JX_EQ_EULER_THETA, PERT, POW, VAR)
#define JX_GROW(NGL, EPB, VAR) \
    JX_GSET(NGL, EPB, VAR, false, false), JX_GSET(NGL, EPB, VAR, false, true), JX_GSET(NGL, EPB, VAR, true, false), \
    JX_GSET(NGL, EPB, VAR, true, true)

#define JX_TSET(NGL, ZW, PW, VAR, PERT, POW) make_team_set<NGL, EulerTheta<3, PERT, POW>, ZW, PW>(JX_EQ_EULER_THETA, PERT, POW, VAR)
#define JX_TROW(NGL, ZW, PW, VAR) \
    JX_TSET(NGL, ZW, PW, VAR, false, false), JX_TSET(NGL, ZW, PW, VAR, false, true), JX_TSET(NGL, ZW, PW, VAR, true, false), \
    JX_TSET(NGL, ZW, PW, VAR, true, true)

#define JX_T2SET(NGL, VAR, PERT, POW) make_team2_set<NGL, EulerTheta<3, PERT, POW>, (VAR) == 11>(JX_EQ_EULER_THETA, PERT, POW, VAR)

const KernelSet *lookup_euler_theta_3d(int ngl, int lpert, int jxpow, int lvisc, int variant) {
#ifdef JX_MIN_BUILD   // kernel experiments: nop 4, TOTAL, jx_pow only -- generic, team (9) and team2 (10)
    static const KernelSet table[] = {JX_SET(5, false, true, false), JX_SET(5, false, true, true), JX_TSET(5, 2, 2, 9, false, true),
                                      JX_T2SET(5, 10, false, true), JX_T2SET(5, 11, false, true)};
#else
    static const KernelSet table[] = {JX_ROW(3), JX_ROW(5), JX_ROW(6), JX_ROW(8), JX_PROW(3), JX_PROW(5), JX_WROW(3), JX_WROW(5), JX_WROW(6),
                                      JX_GROW(3, 7, 5), JX_GROW(5, 5, 5), JX_GROW(5, 1, 6), JX_TROW(3, 3, 1, 8), JX_TROW(5, 2, 1, 8), JX_TROW(5, 2, 2, 9),
                                      JX_T2SET(5, 10, false, false), JX_T2SET(5, 10, false, true), JX_T2SET(5, 11, false, false), JX_T2SET(5, 11, false, true)};
#endif
    for (const KernelSet &k : table)
        if (k.ngl == ngl && k.lpert == lpert && k.jxpow == jxpow && k.lvisc == lvisc && k.variant == variant) return &k;
    return nullptr;
}

}  // namespace jx
