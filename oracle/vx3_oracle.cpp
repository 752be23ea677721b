// TEST INFRASTRUCTURE — CPU restatement ("oracle") of the Voxelyze3 device time step.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library,
// and only as the checker.  The product path (gpuvoxels_b200/csrc) never links or calls it.
//
// It restates, in plain scalar C++ on the flat VX3B_WorldDesc arrays, what the reference's
// device code does per step, double state with the reference's float casts kept where they
// occur.  Pinning: tests/test_oracle_vs_reference.py compares it with the UNMODIFIED reference
// CPU library (oracle/_ref/vx2_ref, built from /root/reference) on VXA inputs — Voxelyze2
// semantics directly and the VX3 stepping deltas (temperature phase offsets, cilia, angMom
// zeroing, double dt) through the harness — and against golden vectors in tests/golden/.
// VX3 voxel-voxel contact (A.4 below): the per-pair FORCE is pinned against the reference's own
// collision system (same formula, VX_Collision.cpp:42-56) on a two-body drop, a block landing on
// two pebbles and a cilia body pushing a target (tests/golden/*_vx2col.vxbin: contacts BETWEEN
// bodies, where both schemes watch the same pairs); the pair SELECTION inside one body (all
// surface pairs, depth-1 exclusion) has no executable reference here (the VX3 CUDA code cannot be built: boost +
// device-side cudaDeviceSynchronize; Voxelyze2 selects pairs differently): that part is PARITY UNPINNED.
//
// Reference map (all under /root/reference/Voxelyze3/src unless noted):
//   step sequencing          VX3/VX3_VoxelyzeKernel.cu:237-359 (doTimeStep)
//   temperature              VX3/VX3_VoxelyzeKernel.cu:219-235, 625-650; VX3/VX3_Voxel.h:90-105; VX3/VX3_Link.cu:80-83
//   link forces              VX3/VX3_Link.cu:90-218 (orientLink, updateForces), :220-260 (updateStrain),
//                            VX3/VX3_Link.h:97-191 (axis swizzles), Utils/VX3_Quat3D.h:347-476
//   voxel integration        VX3/VX3_Voxel.cu:162-277 (timeStep), :350-397 (force, moment), :399-426 (floorForce)
//   contact                  VX3/VX3_VoxelyzeKernel.cu:651-727, 833-843; VX3/VX3_Collision.cu:3-31
//   surface list             VX3/VX3_VoxelyzeKernel.cu:495-513; VX3/VX3_Voxel.cu:515-524
//   cilia                    VX3/VX3_VoxelyzeKernel.cu:846-859
//   dt                       VX3/VX3_VoxelyzeKernel.cu:184-217, 244-254 (VX3) / old/Voxelyze.cpp:286-311 (VX2)
//   COM / angle / closeness  VX3/VX3_VoxelyzeKernel.cu:314-334, 477-493, 530-563
//   world lifecycle          VX3/VX3_SimulationManager.cu:11-121 (CUDA_Simulation), :416-458 (collectResults)
//   formula VM               Utils/VX3_MathTree.h:50-193

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "vx3b200.h"

namespace {

struct V3 {
    double x = 0, y = 0, z = 0;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double f) { return {f * a.x, f * a.y, f * a.z}; }
inline V3 operator*(double f, V3 a) { return {f * a.x, f * a.y, f * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double len(V3 a) { return sqrt(len2(a)); }
inline V3 div(V3 a, double f) { // Vec3D::operator/ multiplies by the reciprocal (Utils/VX3_Vec3D.h:42)
    double inv = 1.0 / f;
    return {inv * a.x, inv * a.y, inv * a.z};
}
inline V3 normalized(V3 a) {
    double l = len(a);
    return l > 0 ? div(a, l) : a;
}

struct Q {
    double w = 1, x = 0, y = 0, z = 0;
};
inline Q qmul(Q a, Q f) { // Utils/VX3_Quat3D.h:217-222
    return {a.w * f.w - a.x * f.x - a.y * f.y - a.z * f.z, a.w * f.x + a.x * f.w + a.y * f.z - a.z * f.y,
            a.w * f.y - a.x * f.z + a.y * f.w + a.z * f.x, a.w * f.z + a.x * f.y - a.y * f.x + a.z * f.w};
}
inline Q conj(Q a) { return {a.w, -a.x, -a.y, -a.z}; }
inline V3 rotate(Q q, V3 f) { // RotateVec3D, Utils/VX3_Quat3D.h:439-448
    double tw = f.x * q.x + f.y * q.y + f.z * q.z;
    double tx = f.x * q.w - f.y * q.z + f.z * q.y;
    double ty = f.x * q.z + f.y * q.w - f.z * q.x;
    double tz = -f.x * q.y + f.y * q.x + f.z * q.w;
    return {q.w * tx + q.x * tw + q.y * tz - q.z * ty, q.w * ty - q.x * tz + q.y * tw + q.z * tx, q.w * tz + q.x * ty - q.y * tx + q.z * tw};
}
inline V3 rotate_inv(Q q, V3 f) { // RotateVec3DInv, Utils/VX3_Quat3D.h:465-474
    double tw = q.x * f.x + q.y * f.y + q.z * f.z;
    double tx = q.w * f.x - q.y * f.z + q.z * f.y;
    double ty = q.w * f.y + q.x * f.z - q.z * f.x;
    double tz = q.w * f.z - q.x * f.y + q.y * f.x;
    return {tw * q.x + tx * q.w + ty * q.z - tz * q.y, tw * q.y - tx * q.z + ty * q.w + tz * q.x, tw * q.z + tx * q.y - ty * q.x + tz * q.w};
}
inline V3 to_rotation_vector(Q q) { // Utils/VX3_Quat3D.h:347-359
    if (q.w >= 1.0 || q.w <= -1.0)
        return {};
    double sl = 1.0 - q.w * q.w;
    V3 v{q.x, q.y, q.z};
    if (sl < 2.4e-3)
        return (v * 2.0) * sqrt((2 - 2 * q.w) / sl);
    return div((v * 2.0) * acos(q.w), sqrt(sl));
}
inline Q from_rotation_vector(V3 in) { // Utils/VX3_Quat3D.h:363-382
    V3 th = div(in, 2);
    double m2 = len2(th), w, s;
    if (m2 * m2 < 5.328e-15) {
        w = 1.0 - 0.5 * m2;
        s = 1.0 - m2 / 6.0;
    } else {
        double m = sqrt(m2);
        w = cos(m);
        s = sin(m) / m;
    }
    return {w, th.x * s, th.y * s, th.z * s};
}
inline void from_angle_to_pos_x(Q &q, V3 from) { // Utils/VX3_Quat3D.h:384-437
    if (from.x == 0 && from.y == 0 && from.z == 0)
        return;
    const double SMALL = 1.732e-2, DISCARD = 1e-7, PI_ = 3.14159265358979;
    double yox = from.y / from.x, zox = from.z / from.x;
    if (yox < SMALL && yox > -SMALL && zox < SMALL && zox > -SMALL) {
        q.x = 0;
        q.y = 0.5 * zox;
        q.z = -0.5 * yox;
        q.w = 1 + 0.5 * (-q.y * q.y - q.z * q.z);
        return;
    }
    V3 n = from;
    double l = len(n);
    if (l > 0)
        n = div(n, l);
    double theta = acos(n.x);
    if (theta > PI_ - DISCARD) {
        q = {0, 0, 1, 0};
        return;
    }
    double axisMagInv = 1.0 / sqrt(n.z * n.z + n.y * n.y);
    double a = 0.5 * theta, s = sin(a);
    q.w = cos(a);
    q.x = 0;
    q.y = n.z * axisMagInv * s;
    q.z = -n.y * axisMagInv * s;
}

// axis swizzles (VX3_Link.h:97-191)
inline V3 to_axis_x(V3 v, int axis) { return axis == 1 ? V3{v.y, -v.x, v.z} : axis == 2 ? V3{v.z, v.y, -v.x} : v; }
inline Q to_axis_x(Q q, int axis) { return axis == 1 ? Q{q.w, q.y, -q.x, q.z} : axis == 2 ? Q{q.w, q.z, q.y, -q.x} : q; }
inline V3 to_axis_original(V3 v, int axis) { return axis == 1 ? V3{-v.y, v.x, v.z} : axis == 2 ? V3{-v.z, v.y, v.x} : v; }

double eval_formula(const VX3B_Token *tok, double x, double y, double z, double hit, double t, double angle, double closeness, int pairs, int nv) {
    // FIFO queue machine, Utils/VX3_MathTree.h:50-193; an empty program yields 0 here (UB in the reference)
    double values[VX3B_MAX_TOKENS + 1];
    values[0] = 0.0;
    int vc = 0, pc = 0;
    for (int i = 0; i < VX3B_MAX_TOKENS; i++) {
        double v = 0;
        int op = tok[i].op;
        if (op == VX3B_MT_END)
            return values[pc];
        if (op == VX3B_MT_CONST)
            v = tok[i].value;
        else if (op == VX3B_MT_E)
            v = 2.71828182845904523536;
        else if (op == VX3B_MT_PI)
            v = 3.14159265358979323846;
        else if (op == VX3B_MT_VAR) {
            double id = tok[i].value;
            v = id < 0.5 ? x : id < 1.5 ? y : id < 2.5 ? z : id < 3.5 ? hit : id < 4.5 ? t : id < 5.5 ? angle : id < 6.5 ? closeness : id < 7.5 ? pairs : nv;
        } else if (op == VX3B_MT_ADD || op == VX3B_MT_SUB || op == VX3B_MT_MUL || op == VX3B_MT_DIV || op == VX3B_MT_POW ||
                   op == VX3B_MT_GREATERTHAN || op == VX3B_MT_LESSTHAN || op == VX3B_MT_AND || op == VX3B_MT_OR) {
            double a = values[pc + 1], b = values[pc];
            pc += 2;
            v = op == VX3B_MT_ADD ? a + b : op == VX3B_MT_SUB ? a - b : op == VX3B_MT_MUL ? a * b : op == VX3B_MT_DIV ? a / b
                : op == VX3B_MT_POW ? pow(a, b) : op == VX3B_MT_GREATERTHAN ? (double)(a > b) : op == VX3B_MT_LESSTHAN ? (double)(a < b)
                : op == VX3B_MT_AND ? (double)(a && b) : (double)(a || b);
        } else {
            double a = values[pc];
            pc += 1;
            switch (op) {
            case VX3B_MT_SQRT: v = sqrt(a); break;
            case VX3B_MT_SIN: v = sin(a); break;
            case VX3B_MT_COS: v = cos(a); break;
            case VX3B_MT_TAN: v = tan(a); break;
            case VX3B_MT_ATAN: v = atan(a); break;
            case VX3B_MT_LOG: v = log(a); break;
            case VX3B_MT_INT: v = rint(a); break;
            case VX3B_MT_ABS: v = fabs(a); break;
            case VX3B_MT_NOT: v = !a; break;
            case VX3B_MT_NORMALCDF: v = 0.5 * erfc(-a * 0.70710678118654752440); break;
            default: return -1;
            }
        }
        if (pc > vc)
            return -1;
        values[vc++] = v;
    }
    return -1;
}

struct World {
    VX3B_WorldDesc d; // scalar fields + formulas; pointer members are re-pointed at the copies below
    std::vector<int16_t> lattice;
    std::vector<int32_t> vmat, lneg, lpos, lmat;
    std::vector<uint8_t> laxis, dof;
    std::vector<double> phase, baseCilia, shiftCilia, extT, extR;
    std::vector<float> extF, extM;
    std::vector<VX3B_VoxelMaterial> vm;
    std::vector<VX3B_LinkMaterial> lm;
    int nv = 0, nl = 0;

    // voxel state
    std::vector<V3> pos, linMom, angMom, contactForce, ciliaForce, initialPos;
    std::vector<Q> orient;
    std::vector<float> temp, previousDt;
    std::vector<uint8_t> staticFriction;
    std::vector<int32_t> vlinks; // [nv][6]
    std::vector<int32_t> surface;
    // signals and topology changes.  The reference's kernels race here; this restatement runs every "thread" of a
    // launch to completion in index order - the schedule of oracle/_ref/vx3_ref, ONE valid execution of the reference's code.
    std::vector<uint8_t> vremoved, vEnableAttach, lremoved, ldetached, matRemoved;
    std::vector<int32_t> lNewLink;
    std::vector<double> localSignal, localSignalDt, inactiveUntil, sigValue, sigActiveTime, paceMakerNext;
    bool surfaceChanged = true; // CUDA_Simulation sets it before the first step (VX3_SimulationManager.cu:22)
    bool initialPositionReinitialized = false;
    std::vector<int32_t> targets; // registerTargets (VX3_VoxelyzeKernel.cu:536-543)
    bool dynamic() const { return d.enable_signals || d.enable_attach || d.enable_detach || d.secondary_experiment; }
    // link state
    std::vector<V3> pos2, a1v, a2v, fNeg, fPos, mNeg, mPos;
    std::vector<uint8_t> smallAngle, velValid;
    std::vector<float> strain, maxStrain, strainOffset;
    std::vector<double> restLen;

    double t = 0, optimalDt = 0;
    float dt = 0;
    int64_t stepCount = 0;
    V3 com, finalCom, initialCom, comHist[2];
    int angleSampleTimes = 0, collisionCount = 0, numClosePairs = 0;
    double recentAngle = 0, targetCloseness = 0, fitness = 0;
    bool diverged = false, stopped = false;
    std::vector<std::pair<int32_t, int32_t>> lastPairs;

    const VX3B_VoxelMaterial &M(int v) const { return vm[vmat[v]]; }
    bool vx3() const { return !d.vx2_compat; }

    double base_size(int v) const { // VX3_Voxel.h:96-99 — (1 + tempe*alphaCTE) is a float expression
        return M(v).nomSize * (1 + temp[v] * M(v).alphaCTE);
    }
    double base_size_average(int v) const { // VX3_Voxel.h:102-105
        double b = base_size(v);
        return (b + b + b) / 3.0f;
    }
    float floor_penetration(int v) const { // VX3_Voxel.h:185-187
        return (float)(base_size_average(v) / 2 - M(v).nomSize / 2 - pos[v].z);
    }
    V3 original_position(int v) const {
        double s = M(v).nomSize;
        return {lattice[v * 3] * s, lattice[v * 3 + 1] * s, lattice[v * 3 + 2] * s};
    }

    void init(const VX3B_WorldDesc *in) {
        d = *in;
        nv = in->num_voxels;
        nl = in->num_links;
        lattice.assign(in->lattice, in->lattice + 3 * nv);
        vmat.assign(in->voxel_mat, in->voxel_mat + nv);
        auto cp = [&](const double *p, size_t n, std::vector<double> &dst) {
            if (p)
                dst.assign(p, p + n);
            else
                dst.assign(n, 0.0);
        };
        cp(in->phase_offset, nv, phase);
        cp(in->base_cilia, 3 * nv, baseCilia);
        cp(in->shift_cilia, 3 * nv, shiftCilia);
        cp(in->ext_translation, 3 * nv, extT);
        cp(in->ext_rotation, 3 * nv, extR);
        if (in->dof_fixed)
            dof.assign(in->dof_fixed, in->dof_fixed + nv);
        else
            dof.assign(nv, 0);
        if (in->ext_force)
            extF.assign(in->ext_force, in->ext_force + 3 * nv);
        else
            extF.assign(3 * nv, 0.0f);
        if (in->ext_moment)
            extM.assign(in->ext_moment, in->ext_moment + 3 * nv);
        else
            extM.assign(3 * nv, 0.0f);
        lneg.assign(in->link_neg, in->link_neg + nl);
        lpos.assign(in->link_pos, in->link_pos + nl);
        lmat.assign(in->link_mat, in->link_mat + nl);
        laxis.assign(in->link_axis, in->link_axis + nl);
        vm.assign(in->voxel_mats, in->voxel_mats + in->num_voxel_mats);
        lm.assign(in->link_mats, in->link_mats + in->num_link_mats);

        pos.resize(nv);
        linMom.assign(nv, V3{});
        angMom.assign(nv, V3{});
        contactForce.assign(nv, V3{});
        ciliaForce.assign(nv, V3{});
        orient.assign(nv, Q{});
        temp.assign(nv, in->initial_temp);
        previousDt.assign(nv, 0.0f);
        staticFriction.assign(nv, 1); // CVX_Voxel::reset sets FLOOR_STATIC_FRICTION (old/VX_Voxel.cpp:47-56)
        for (int v = 0; v < nv; v++)
            pos[v] = original_position(v);
        initialPos = pos;
        vlinks.assign((size_t)nv * 6, -1);
        for (int l = 0; l < nl; l++) { // direction 2a = positive side of the negative end, 2a+1 = negative side of the positive end
            vlinks[(size_t)lneg[l] * 6 + 2 * laxis[l]] = l;
            vlinks[(size_t)lpos[l] * 6 + 2 * laxis[l] + 1] = l;
        }
        for (int v = 0; v < nv; v++) { // surface = any of the 6 link slots empty (VX3_Voxel.cu:515-524)
            bool interior = true;
            for (int k = 0; k < 6; k++)
                if (vlinks[(size_t)v * 6 + k] < 0)
                    interior = false;
            if (!interior)
                surface.push_back(v);
        }
        pos2.assign(nl, V3{});
        a1v.assign(nl, V3{});
        a2v.assign(nl, V3{});
        fNeg.assign(nl, V3{});
        fPos.assign(nl, V3{});
        mNeg.assign(nl, V3{});
        mPos.assign(nl, V3{});
        smallAngle.assign(nl, 1);
        velValid.assign(nl, 0);
        strain.assign(nl, 0.0f);
        maxStrain.assign(nl, 0.0f);
        strainOffset.assign(nl, 0.0f);
        vremoved.assign(nv, 0);
        vEnableAttach.assign(nv, 1); // VX3_Voxel.h: enableAttach = true
        lremoved.assign(nl, 0);
        ldetached.assign(nl, 0);
        lNewLink.assign(nl, 0);
        matRemoved.assign(in->num_voxel_mats, 0);
        for (auto *a : {&localSignal, &localSignalDt, &inactiveUntil, &sigValue, &sigActiveTime, &paceMakerNext})
            a->assign(nv, 0.0);
        for (int v = 0; v < nv; v++)
            if (M(v).isTarget)
                targets.push_back(v);
        restLen.resize(nl);
        for (int l = 0; l < nl; l++)
            update_rest_length(l);
        update_com();
        initialCom = finalCom = com;
    }

    void update_rest_length(int l) { restLen[l] = 0.5 * (base_size(lneg[l]) + base_size(lpos[l])); } // VX3_Link.cu:80-83

    // VX3_Link::axialStiffness (VX3_Link.cu:277-286): a1 for nu == 0, else eHat * area / ((strain + 1) * restLength) with the link's
    // current values - at the first step, where the time step is fixed: strain 0, area size^2, rest length at the first temperature
    float axial_stiffness(int l) const {
        const VX3B_LinkMaterial &m = lm[lmat[l]];
        if (m.nu == 0.0f)
            return m.a1;
        float sz1 = (float)M(lneg[l]).nomSize, sz2 = (float)M(lpos[l]).nomSize;
        float area = 0.5f * (sz1 * sz1 + sz2 * sz2);
        return (float)(m.eHat * area / ((strain[l] + 1) * restLen[l]));
    }
    double recommended_dt() const {
        if (vx3()) { // VX3_VoxelyzeKernel.cu:184-217: double accumulation
            double maxF2 = 0.0f;
            for (int l = 0; l < nl; l++) {
                double m1 = M(lneg[l]).mass, m2 = M(lpos[l]).mass;
                double f2 = axial_stiffness(l) / (m1 < m2 ? m1 : m2);
                if (f2 > maxF2)
                    maxF2 = f2;
            }
            if (maxF2 <= 0.0f)
                for (int v = 0; v < nv; v++) {
                    double f2 = M(v).E * M(v).nomSize / M(v).mass;
                    if (f2 > maxF2)
                        maxF2 = f2;
                }
            if (maxF2 <= 0.0f)
                return 0.0f;
            return 1.0f / (6.283185f * sqrt(maxF2));
        }
        float maxF2 = 0.0f; // old/Voxelyze.cpp:286-311: float throughout
        for (int l = 0; l < nl; l++) {
            float m1 = M(lneg[l]).mass, m2 = M(lpos[l]).mass;
            float f2 = lm[lmat[l]].a1 / (m1 < m2 ? m1 : m2);
            if (f2 > maxF2)
                maxF2 = f2;
        }
        if (maxF2 <= 0.0f)
            for (int v = 0; v < nv; v++) {
                float f2 = (float)(M(v).E * M(v).nomSize / M(v).mass);
                if (f2 > maxF2)
                    maxF2 = f2;
            }
        if (maxF2 <= 0.0f)
            return 0.0f;
        return (float)(1.0f / (6.283185f * sqrtf(maxF2)));
    }

    void update_temperature() { // VX3_VoxelyzeKernel.cu:219-235, 625-650
        if (!d.vary_temp_enabled || !(d.temp_period > 0))
            return;
        for (int v = 0; v < nv; v++) {
            if (vremoved[v] || (vx3() && M(v).TurnOnThermalExpansionAfterThisManySeconds > t))
                continue;
            if (vx3() && M(v).fixed)
                continue;
            double T = d.temp_amplitude * sin(2 * 3.1415926f * (t / d.temp_period + phase[v]));
            if (!d.enable_expansion && T > 0)
                T = 0;
            temp[v] = (float)T;
        }
        for (int l = 0; l < nl; l++)
            update_rest_length(l);
    }

    // VX3_Material::stress (VX3_Material.cu:90-124).  transverseStrainSum: the link's currentTransverseStrainSum, which Voxelyze3
    // never refreshes after construction (updateTransverseInfo is commented out, VX3_Link.cu:147-150), i.e. 0.
    static float material_stress(const VX3B_LinkMaterial &m, float strain, float transverseStrainSum, bool forceLinear) {
        if (m.epsilonFail != -1.0f && strain > m.epsilonFail)
            return 0.0f; // isFailed(strain)
        if (strain <= m.strain_data[1] || m.linear || forceLinear) {
            if (m.nu == 0.0f)
                return m.E * strain;
            return m.eHat * ((1 - m.nu) * strain + m.nu * transverseStrainSum);
        }
        const int n = m.num_points;
        for (int i = 2; i < n; i++) {
            if (strain <= m.strain_data[i] || i == n - 1) {
                float Perc = (strain - m.strain_data[i - 1]) / (m.strain_data[i] - m.strain_data[i - 1]);
                float basicStress = m.stress_data[i - 1] + Perc * (m.stress_data[i] - m.stress_data[i - 1]);
                if (m.nu == 0.0f)
                    return basicStress;
                float modulus = (m.stress_data[i] - m.stress_data[i - 1]) / (m.strain_data[i] - m.strain_data[i - 1]);
                float modulusHat = modulus / ((1 - 2 * m.nu) * (1 + m.nu));
                float effectiveStrain = basicStress / modulus;
                float effectiveTransverseStrainSum = transverseStrainSum * (effectiveStrain / strain);
                return modulusHat * ((1 - m.nu) * effectiveStrain + m.nu * effectiveTransverseStrainSum);
            }
        }
        return 0.0f;
    }
    // VX3_Link::updateStrain (VX3_Link.cu:232-268): a non-linear material remembers its largest strain and backs off along E
    float update_strain(int l, const VX3B_LinkMaterial &m, float axialStrain) {
        const float tss = 0.0f;
        strain[l] = axialStrain;
        if (m.linear) {
            if (axialStrain > maxStrain[l])
                maxStrain[l] = axialStrain;
            return material_stress(m, axialStrain, tss, false);
        }
        float returnStress;
        if (axialStrain > maxStrain[l]) { // new territory on the stress/strain curve
            maxStrain[l] = axialStrain;
            returnStress = material_stress(m, axialStrain, tss, false);
            if (m.nu != 0.0f)
                strainOffset[l] = maxStrain[l] - material_stress(m, axialStrain, 0.0f, false) / (m.eHat * (1 - m.nu));
            else
                strainOffset[l] = maxStrain[l] - returnStress / m.E;
        } else { // backed off: linear with a strain offset according to the maximum plastic deformation
            float relativeStrain = axialStrain - strainOffset[l];
            if (m.nu != 0.0f)
                returnStress = material_stress(m, relativeStrain, tss, true);
            else
                returnStress = m.E * relativeStrain;
        }
        return returnStress;
    }

    void update_link(int l) { // VX3_Link.cu:135-218
        const VX3B_LinkMaterial &m = lm[lmat[l]];
        int vn = lneg[l], vp = lpos[l], axis = laxis[l];
        V3 oldPos2 = pos2[l], oldA1 = a1v[l], oldA2 = a2v[l];
        // orientLink (VX3_Link.cu:90-133)
        V3 d2 = to_axis_x(pos[vp] - pos[vn], axis);
        Q q1 = to_axis_x(orient[vn], axis), q2 = to_axis_x(orient[vp], axis);
        Q total = conj(q1);
        d2 = rotate(total, d2);
        q2 = qmul(total, q2);
        q1 = Q{};
        float smallTurn = (float)((fabs(d2.z) + fabs(d2.y)) / d2.x);
        float extendPerc = (float)(fabs(1 - d2.x / restLen[l]));
        const float HYST = 1.2f, BEND = 0.05f, EXT = 0.50f; // VoxCAD/src/VX2/include/types.h:57-59
        if (!smallAngle[l] && smallTurn < BEND && extendPerc < EXT) {
            smallAngle[l] = 1;
            velValid[l] = 0;
        } else if (smallAngle[l] && (smallTurn > HYST * BEND || extendPerc > HYST * EXT)) {
            smallAngle[l] = 0;
            velValid[l] = 0;
        }
        if (smallAngle[l]) {
            d2.x -= restLen[l];
        } else {
            from_angle_to_pos_x(q1, d2);
            total = qmul(q1, total);
            q2 = qmul(q1, q2);
            d2 = V3{len(d2) - restLen[l], 0, 0};
        }
        V3 th1 = to_rotation_vector(q1), th2 = to_rotation_vector(q2);
        pos2[l] = d2;
        a1v[l] = th1;
        a2v[l] = th2;
        V3 dPos2 = 0.5 * (d2 - oldPos2), dA1 = 0.5 * (th1 - oldA1), dA2 = 0.5 * (th2 - oldA2);

        float eps = (float)(d2.x / restLen[l]);
        float stress = update_strain(l, m, eps);
        if (m.epsilonFail != -1.0f && maxStrain[l] > m.epsilonFail) { // isFailed(): all four outputs zero
            fNeg[l] = fPos[l] = mNeg[l] = mPos[l] = V3{};
            return;
        }
        float b1 = m.b1, b2 = m.b2, b3 = m.b3, a2 = m.a2;
        float sz1 = (float)M(vn).nomSize, sz2 = (float)M(vp).nomSize;
        float area = 0.5f * (sz1 * sz1 + sz2 * sz2); // currentTransverseArea, nu == 0 (VX3_Link.cu:85-88)
        V3 FN{(double)(stress * area), b1 * d2.y - b2 * (th1.z + th2.z), b1 * d2.z + b2 * (th1.y + th2.y)};
        V3 FP = -FN;
        V3 MN{a2 * (th2.x - th1.x), -b2 * d2.z - b3 * (2 * th1.y + th2.y), b2 * d2.y - b3 * (2 * th1.z + th2.z)};
        V3 MP{a2 * (th1.x - th2.x), -b2 * d2.z - b3 * (th1.y + 2 * th2.y), b2 * d2.y - b3 * (th1.z + 2 * th2.z)};
        if (velValid[l]) { // local damping (VX3_Link.cu:172-188)
            float sqA1 = m.sqA1, sqA2xIp = m.sqA2xIp, sqB1 = m.sqB1, sqB2xFMp = m.sqB2xFMp, sqB3xIp = m.sqB3xIp;
            V3 posCalc{sqA1 * dPos2.x, sqB1 * dPos2.y - sqB2xFMp * (dA1.z + dA2.z), sqB1 * dPos2.z + sqB2xFMp * (dA1.y + dA2.y)};
            float dmN = 2 * M(vn).sqrtMass * M(vn).zetaInternal / previousDt[vn]; // dampingMultiplier, VX3_Voxel.h:206-208
            float dmP = 2 * M(vp).sqrtMass * M(vp).zetaInternal / previousDt[vp];
            FN = FN + dmN * posCalc;
            FP = FP - dmP * posCalc;
            MN = MN - (0.5 * dmN) * V3{-sqA2xIp * (dA2.x - dA1.x), sqB2xFMp * dPos2.z + sqB3xIp * (2 * dA1.y + dA2.y),
                                       -sqB2xFMp * dPos2.y + sqB3xIp * (2 * dA1.z + dA2.z)};
            MP = MP - (0.5 * dmP) * V3{sqA2xIp * (dA2.x - dA1.x), sqB2xFMp * dPos2.z + sqB3xIp * (dA1.y + 2 * dA2.y),
                                       -sqB2xFMp * dPos2.y + sqB3xIp * (dA1.z + 2 * dA2.z)};
        } else
            velValid[l] = 1;
        if (!smallAngle[l]) {
            FN = rotate_inv(q1, FN);
            MN = rotate_inv(q1, MN);
        }
        FP = rotate_inv(q2, FP);
        MP = rotate_inv(q2, MP);
        fNeg[l] = to_axis_original(FN, axis);
        fPos[l] = to_axis_original(FP, axis);
        mNeg[l] = to_axis_original(MN, axis);
        mPos[l] = to_axis_original(MP, axis);
        if (lNewLink[l]) { // a link made by attachment eases in: 1 % of its forces for SafetyGuard steps (VX3_Link.cu:206-213)
            fNeg[l] = fNeg[l] * 0.01;
            fPos[l] = fPos[l] * 0.01;
            mNeg[l] = mNeg[l] * 0.01;
            mPos[l] = mPos[l] * 0.01;
            lNewLink[l] -= 1;
        }
    }

    void update_contacts() { // VX3_VoxelyzeKernel.cu:682-831, 833-843; VX3_Collision.cu:3-31
        lastPairs.clear();
        int ns = (int)surface.size();
        for (int a = 0; a < ns; a++)
            for (int b = 0; b < a; b++) {
                int v1 = surface[a], v2 = surface[b];
                if (vremoved[v1] || vremoved[v2])
                    continue;
                const VX3B_VoxelMaterial &m1 = M(v1), &m2 = M(v2);
                if (m1.fixed && m2.fixed)
                    continue;
                V3 diff = pos[v1] - pos[v2];
                double R = (base_size_average(v1) + base_size_average(v2)) * 0.625;
                if (diff.x > R || diff.x < -R || diff.y > R || diff.y < -R || diff.z > R || diff.z < -R)
                    continue;
                if (len(diff) > R)
                    continue;
                bool linked = false; // is_neighbor(depth 1): directly linked voxels never collide
                for (int k = 0; k < 6; k++) {
                    int l = vlinks[(size_t)v1 * 6 + k];
                    if (l >= 0 && (lneg[l] == v2 || lpos[l] == v2))
                        linked = true;
                }
                if (linked)
                    continue;
                V3 force{};
                if (d.enable_collision) {
                    lastPairs.emplace_back(v1, v2);
                    float k1 = (float)(2 * m1.E * m1.nomSize), k2 = (float)(2 * m2.E * m2.nomSize); // penetrationStiffness
                    double stiff = 2.0f / (1.0f / k1 + 1.0f / k2);
                    double damp = 0.5f * (m1.zetaCollision * m1._2xSqMxExS + m2.zetaCollision * m2._2xSqMxExS);
                    V3 offset = pos[v2] - pos[v1];
                    double relDist = R - len(offset);
                    if (relDist > 0) {
                        V3 unit = normalized(offset);
                        double relVel = dot(linMom[v1] * m1.massInverse, unit) - dot(linMom[v2] * m2.massInverse, unit);
                        force = unit * (stiff * relDist + damp * relVel);
                    }
                    contactForce[v1] = contactForce[v1] + force;
                    contactForce[v2] = contactForce[v2] - force;
                    if ((m1.isTarget && !m2.isTarget) || (m2.isTarget && !m1.isTarget)) {
                        collisionCount++;
                        if (d.enable_signals) // the voxel that touched a target fires (VX3_VoxelyzeKernel.cu:706-716)
                            receive_signal(m1.isTarget ? v2 : v1, 100, t, true);
                    }
                }
                if (d.enable_attach && try_attach(v1, v2)) { // a pair that has just been linked does not push apart this step
                    contactForce[v1] = contactForce[v1] - force;
                    contactForce[v2] = contactForce[v2] + force;
                }
            }
        std::sort(lastPairs.begin(), lastPairs.end());
    }

    void update_cilia() { // VX3_VoxelyzeKernel.cu:846-859
        for (int v : surface) {
            if (vremoved[v] || M(v).Cilia == 0 || M(v).TurnOnCiliaAfterThisManySeconds > t)
                continue;
            V3 base{baseCilia[v * 3], baseCilia[v * 3 + 1], baseCilia[v * 3 + 2]}, shift{shiftCilia[v * 3], shiftCilia[v * 3 + 1], shiftCilia[v * 3 + 2]};
            ciliaForce[v] = rotate(orient[v], base + localSignal[v] * shift);
        }
    }

    V3 link_sum(int v, bool moments) const { // Σ over directions 0..5, local frame (VX3_Voxel.cu:352-357, 384-389)
        V3 s{};
        for (int k = 0; k < 6; k++) {
            int l = vlinks[(size_t)v * 6 + k];
            if (l < 0)
                continue;
            bool positiveEnd = (k % 2) == 1; // link->force(isNegative(dir)) -> forcePos for odd directions
            s = s + (moments ? (positiveEnd ? mPos[l] : mNeg[l]) : (positiveEnd ? fPos[l] : fNeg[l]));
        }
        return s;
    }

    void step_voxel(int v, double dtD) { // VX3_Voxel.cu:162-277
        const VX3B_VoxelMaterial &m = M(v);
        previousDt[v] = (float)dtD;
        if (dtD == 0.0f)
            return;
        uint8_t fixed = dof[v];
        V3 eT{extT[v * 3], extT[v * 3 + 1], extT[v * 3 + 2]}, eR{extR[v * 3], extR[v * 3 + 1], extR[v * 3 + 2]};
        Q eRQ = (eR.x != 0 || eR.y != 0 || eR.z != 0) ? from_rotation_vector(eR) : Q{};
        if (fixed == VX3B_DOF_ALL) {
            pos[v] = original_position(v) + eT;
            orient[v] = eRQ;
            linMom[v] = angMom[v] = V3{};
            return;
        }
        // force() (VX3_Voxel.cu:350-380)
        V3 F = rotate(orient[v], link_sum(v, false));
        F = F + V3{extF[v * 3], extF[v * 3 + 1], extF[v * 3 + 2]};
        F = F - (linMom[v] * m.massInverse) * (double)(m.zetaGlobal * m._2xSqMxExS);
        F.z += (double)(-m.mass * 9.80665f * m.gravMult);
        F = F - contactForce[v];
        contactForce[v] = V3{};
        F = F + ciliaForce[v] * m.Cilia;
        ciliaForce[v] = V3{};
        // force field (VX3_Voxel.cu:178-184)
        for (int k = 0; k < 3; k++)
            if (d.force_field[k][0].op != VX3B_MT_END) {
                double f = eval_formula(d.force_field[k], pos[v].x, pos[v].y, pos[v].z, collisionCount, t, recentAngle, targetCloseness,
                                        numClosePairs, nv);
                (k == 0 ? F.x : k == 1 ? F.y : F.z) += f;
            }
        V3 fric = F;
        bool floorOn = d.floor_enabled != 0;
        if (floorOn) { // floorForce (VX3_Voxel.cu:399-426)
            float pen = floor_penetration(v);
            if (pen >= 0) {
                V3 vel = linMom[v] * m.massInverse;
                float normalForce = (float)(2 * m.E * m.nomSize) * pen;
                F.z += normalForce - (double)(m.zetaCollision * m._2xSqMxExS) * vel.z;
                if (staticFriction[v]) {
                    float surfaceSq = (float)(F.x * F.x + F.y * F.y);
                    float frictionSq = (m.muStatic * normalForce) * (m.muStatic * normalForce);
                    if (surfaceSq > frictionSq)
                        staticFriction[v] = 0;
                } else {
                    F = F - (double)(m.muKinetic * normalForce) * normalized(V3{vel.x, vel.y, 0});
                }
            } else
                staticFriction[v] = 0;
        }
        fric = F - fric;
        linMom[v] = linMom[v] + F * dtD;
        V3 translate = linMom[v] * (dtD * m.massInverse);
        if (floorOn && floor_penetration(v) >= 0) {
            double work = fric.x * translate.x + fric.y * translate.y;
            double hKe = 0.5 * m.massInverse * (linMom[v].x * linMom[v].x + linMom[v].y * linMom[v].y);
            if (hKe + work <= 0)
                staticFriction[v] = 1;
            if (staticFriction[v]) {
                linMom[v].x = linMom[v].y = 0;
                translate.x = translate.y = 0;
            }
        } else
            staticFriction[v] = 0;
        pos[v] = pos[v] + translate;
        // moment() (VX3_Voxel.cu:382-397) and rotation
        V3 Mo = rotate(orient[v], link_sum(v, true));
        Mo = Mo + V3{extM[v * 3], extM[v * 3 + 1], extM[v * 3 + 2]};
        Mo = Mo - (angMom[v] * m.momentInertiaInverse) * (double)(m.zetaGlobal * m._2xSqIxExSxSxS);
        angMom[v] = angMom[v] + Mo * dtD;
        orient[v] = qmul(from_rotation_vector(angMom[v] * (dtD * m.momentInertiaInverse)), orient[v]);
        if (fixed) { // partial dof fixes (VX3_Voxel.cu:222-257)
            double s = m.nomSize;
            if (fixed & VX3B_DOF_TX) {
                pos[v].x = lattice[v * 3] * s + eT.x;
                linMom[v].x = 0;
            }
            if (fixed & VX3B_DOF_TY) {
                pos[v].y = lattice[v * 3 + 1] * s + eT.y;
                linMom[v].y = 0;
            }
            if (fixed & VX3B_DOF_TZ) {
                pos[v].z = lattice[v * 3 + 2] * s + eT.z;
                linMom[v].z = 0;
            }
            if (fixed & (VX3B_DOF_RX | VX3B_DOF_RY | VX3B_DOF_RZ)) {
                if ((fixed & (VX3B_DOF_RX | VX3B_DOF_RY | VX3B_DOF_RZ)) == (VX3B_DOF_RX | VX3B_DOF_RY | VX3B_DOF_RZ)) {
                    orient[v] = eRQ;
                    angMom[v] = V3{};
                } else {
                    V3 rv = to_rotation_vector(orient[v]);
                    if (fixed & VX3B_DOF_RX) {
                        rv.x = 0;
                        angMom[v].x = 0;
                    }
                    if (fixed & VX3B_DOF_RY) {
                        rv.y = 0;
                        angMom[v].y = 0;
                    }
                    if (fixed & VX3B_DOF_RZ) {
                        rv.z = 0;
                        angMom[v].z = 0;
                    }
                    orient[v] = from_rotation_vector(rv);
                }
            }
        }
        if (vx3() && floorOn && floor_penetration(v) >= 0 && staticFriction[v]) // VX3_Voxel.cu:258-264
            angMom[v] = V3{};
        if (d.enable_signals)
            step_signals(v);
    }

    void update_com() { // VX3_VoxelyzeKernel.cu:477-493
        double total = 0;
        V3 sum{};
        for (int v = 0; v < nv; v++) {
            if (!M(v).isMeasured)
                continue;
            double mass = M(v).mass;
            sum = sum + pos[v] * mass;
            total += mass;
        }
        com = total == 0 ? V3{} : div(sum, total);
    }

    void compute_target_closeness() { // VX3_VoxelyzeKernel.cu:545-563
        if (d.max_dist_in_voxel_lengths_to_count_as_pair == 0)
            return;
        double R = d.max_dist_in_voxel_lengths_to_count_as_pair * d.vox_size, ret = 0;
        numClosePairs = 0;
        for (size_t i = 0; i < targets.size(); i++)
            for (size_t j = i + 1; j < targets.size(); j++) {
                double dist = len(pos[targets[i]] - pos[targets[j]]);
                if (dist < R)
                    numClosePairs++;
                ret += 1 / dist;
            }
        targetCloseness = ret;
    }

    bool stop_condition_met() const { // VX3_VoxelyzeKernel.cu:162-174
        if (d.stop_condition[0].op == VX3B_MT_END)
            return false;
        return eval_formula(d.stop_condition, com.x, com.y, com.z, collisionCount, t, recentAngle, targetCloseness, numClosePairs, nv) > 0;
    }

    bool do_time_step() { // VX3_VoxelyzeKernel.cu:237-359
        update_temperature();
        stepCount++;
        if (!optimalDt) {
            optimalDt = recommended_dt();
            if (vx3() && optimalDt < 1e-10)
                optimalDt = 1e-10;
            dt = (float)(d.dt_frac * optimalDt);
        }
        if (dt == 0)
            return true;
        bool div = false;
        for (int l = 0; l < nl; l++) { // gpu_update_links (VX3_VoxelyzeKernel.cu:566-581)
            if (lremoved[l] || ldetached[l])
                continue;
            if (vx3() && M(lneg[l]).fixed && M(lpos[l]).fixed)
                continue;
            update_link(l);
            if (strain[l] > 100)
                div = true; // the reference samples one random link (VX3) / checks every link (VX2); we check every link
        }
        if (div)
            return false;
        if (dynamic() && surfaceChanged) { // VX3_VoxelyzeKernel.cu:282-286
            surfaceChanged = false;
            regenerate_surface();
        }
        if (d.enable_collision || d.enable_attach)
            update_contacts();
        if (d.enable_detach)
            update_detach();
        if (d.enable_cilia)
            update_cilia();
        double dtD = dt;
        for (int v = 0; v < nv; v++) { // gpu_update_voxels (VX3_VoxelyzeKernel.cu:582-624)
            if (vremoved[v])
                continue;
            if (vx3() && M(v).fixed)
                continue;
            step_voxel(v, dtD);
            { // "sticky status": every AttachCondition formula > 0 at the voxel's position
                vEnableAttach[v] = 0;
                bool any = false, all = true;
                for (int c = 0; c < 5 && all; c++) {
                    if (d.attach_condition[c][0].op == VX3B_MT_END)
                        continue; // (the reference evaluates an empty program to whatever its value stack last held: DESIGN.md section 8)
                    any = true;
                    all = eval_formula(d.attach_condition[c], pos[v].x, pos[v].y, pos[v].z, collisionCount, t, recentAngle, targetCloseness, numClosePairs,
                                       nv) > 0;
                }
                vEnableAttach[v] = any && all;
            }
        }
        if (vx3()) { // heading angle + closeness sampling (VX3_VoxelyzeKernel.cu:314-334)
            int cycle = (int)(d.temp_period / dt);
            if (cycle > 0 && stepCount % cycle == 0) {
                angleSampleTimes++;
                comHist[0] = comHist[1];
                comHist[1] = com;
                update_com();
                V3 A = comHist[0], B = comHist[1], C = com;
                auto eq = [](V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
                if (eq(B, C) || eq(A, B) || angleSampleTimes < 3)
                    recentAngle = 0;
                else
                    recentAngle = acos(dot(B - A, C - B) / (len(B - A) * len(C - B)));
                compute_target_closeness();
            }
        }
        if (d.secondary_experiment) { // VX3_VoxelyzeKernel.cu:336-350
            remove_voxels();
            if (!initialPositionReinitialized && d.reinitialize_initial_position_after < t) {
                initialPositionReinitialized = true;
                initialCom = com;
                initialPos = pos;
            }
        }
        t += dt;
        return true;
    }

    // regenerateSurfaceVoxels (VX3_VoxelyzeKernel.cu:495-513): a voxel is interior when all six link slots hold a link that is not
    // detached (VX3_Voxel.cu:515-524); removed voxels are not listed
    void regenerate_surface() {
        surface.clear();
        for (int v = 0; v < nv; v++) {
            bool interior = true;
            for (int k = 0; k < 6; k++) {
                int l = vlinks[(size_t)v * 6 + k];
                if (l < 0 || ldetached[l])
                    interior = false;
            }
            if (!interior && !vremoved[v])
                surface.push_back(v);
        }
    }
    // gpu_update_detach (VX3_VoxelyzeKernel.cu:946-967): a failed link lets go of both voxels
    void update_detach() {
        for (int l = 0; l < nl; l++) {
            if (lremoved[l] || ldetached[l])
                continue;
            const VX3B_LinkMaterial &m = lm[lmat[l]];
            if (m.epsilonFail != -1.0f && maxStrain[l] > m.epsilonFail) {
                ldetached[l] = 1;
                for (int k = 0; k < 6; k++) {
                    if (vlinks[(size_t)lneg[l] * 6 + k] == l)
                        vlinks[(size_t)lneg[l] * 6 + k] = -1;
                    if (vlinks[(size_t)lpos[l] * 6 + k] == l)
                        vlinks[(size_t)lpos[l] * 6 + k] = -1;
                }
                surfaceChanged = true;
            }
        }
    }
    // removeVoxels (VX3_VoxelyzeKernel.cu:365-399): a material whose time has come takes its voxels and their links out of the run
    void remove_voxels() {
        for (int i = 0; i < (int)vm.size(); i++) {
            const double after = vm[i].RemoveFromSimulationAfterThisManySeconds;
            if (matRemoved[i] || !(after > 0) || !(after < t))
                continue;
            for (int v = 0; v < nv; v++) {
                if (vmat[v] != i || vremoved[v])
                    continue;
                vremoved[v] = 1;
                for (int k = 0; k < 6; k++) {
                    const int l = vlinks[(size_t)v * 6 + k];
                    if (l < 0)
                        continue;
                    lremoved[l] = 1;
                    const int other = lneg[l] == v ? lpos[l] : lneg[l];
                    for (int q = 0; q < 6; q++)
                        if (vlinks[(size_t)other * 6 + q] == l) {
                            vlinks[(size_t)other * 6 + q] = -1;
                            break;
                        }
                    vlinks[(size_t)v * 6 + k] = -1;
                }
            }
            matRemoved[i] = 1;
            surfaceChanged = true;
        }
    }
    // is_neighbor (VX3_VoxelyzeKernel.cu:651-680): v2 reachable from v1 over at most `depth` links
    bool is_neighbor(int v1, int v2, int incoming, int depth) const {
        if (v1 == v2)
            return true;
        if (depth <= 0)
            return false;
        for (int k = 0; k < 6; k++) {
            const int l = vlinks[(size_t)v1 * 6 + k];
            if (l < 0 || l == incoming)
                continue;
            if (is_neighbor(lneg[l] == v1 ? lpos[l] : lneg[l], v2, l, depth - 1))
                return true;
        }
        return false;
    }
    // VX3_Voxel::receiveSignal (VX3_Voxel.cu:306-326)
    void receive_signal(int v, double value, double activeTime, bool force) {
        if (!force && inactiveUntil[v] > activeTime)
            return;
        if (value < 0.1)
            return;
        const VX3B_VoxelMaterial &m = M(v);
        inactiveUntil[v] = activeTime + m.inactivePeriod;
        localSignal[v] = value;
        sigValue[v] = value * m.signalValueDecay;
        if (sigValue[v] < 0.1)
            sigValue[v] = 0;
        sigActiveTime[v] = activeTime;
    }
    // propagateSignal, packMaker, localSignalDecay at the end of VX3_Voxel::timeStep (VX3_Voxel.cu:268-348)
    void step_signals(int v) {
        const VX3B_VoxelMaterial &m = M(v);
        if (!(sigActiveTime[v] > t) && !(sigValue[v] < 0.1)) {
            for (int k = 0; k < 6; k++) {
                const int l = vlinks[(size_t)v * 6 + k];
                if (l >= 0)
                    receive_signal(lneg[l] == v ? lpos[l] : lneg[l], sigValue[v], t + m.signalTimeDelay, false);
            }
            sigValue[v] = 0;
            sigActiveTime[v] = 0;
            inactiveUntil[v] = t + 2 * m.signalTimeDelay + m.inactivePeriod;
        }
        if (m.isPaceMaker && !(paceMakerNext[v] > t)) {
            receive_signal(v, 100, t, true);
            paceMakerNext[v] = t + m.PaceMakerPeriod;
        }
        if (!(localSignalDt[v] > t)) {
            if (localSignal[v] < 0.1)
                localSignal[v] = 0;
            else {
                localSignal[v] = localSignal[v] * 0.9;
                localSignalDt[v] = t + 0.01;
            }
        }
    }
    int find_link_material(int a, int b) const { // VX3_VoxelyzeKernel::combinedMaterial (VX3_VoxelyzeKernel.cu:515-528): either order
        for (int k = 0; k < (int)lm.size(); k++)
            if ((lm[k].mat1 == a && lm[k].mat2 == b) || (lm[k].mat1 == b && lm[k].mat2 == a))
                return k;
        return -1;
    }
    // the attachment half of handle_collision_attachment (VX3_VoxelyzeKernel.cu:718-831); returns true when a link was made
    bool try_attach(int v1, int v2) {
        if (!vEnableAttach[v1] || !vEnableAttach[v2])
            return false;
        const VX3B_VoxelMaterial &m1 = M(v1), &m2 = M(v2);
        if (m1.fixed || m2.fixed || vmat[v1] != vmat[v2] || !m1.sticky)
            return false;
        if (is_neighbor(v1, v2, -1, 5))
            return false;
        V3 e = pos[v1] - pos[v2];
        V3 ea = rotate_inv(orient[v1], V3{-e.x, -e.y, -e.z});
        V3 f{fabs(ea.x), fabs(ea.y), fabs(ea.z)};
        int axis, dir1, dir2;
        bool reverse = false;
        auto pick = [&](int a, double comp) {
            axis = a;
            if (comp < 0) {
                dir1 = 2 * a + 1; // *_NEG
                dir2 = 2 * a;
                reverse = true;
            } else {
                dir1 = 2 * a;
                dir2 = 2 * a + 1;
            }
        };
        if (f.x >= f.y && f.x >= f.z)
            pick(0, ea.x);
        else if (f.y >= f.x && f.y >= f.z)
            pick(1, ea.y);
        else
            pick(2, ea.z);
        if (vlinks[(size_t)v1 * 6 + dir1] >= 0 || vlinks[(size_t)v2 * 6 + dir2] >= 0)
            return false;
        const int mat = find_link_material(vmat[v1], vmat[v2]);
        if (mat < 0)
            return false; // (the host model lists the same-material link material of every sticky material)
        // VX3_Link(voxel1, dir1, voxel2, dir2): pVNeg = its second voxel, pVPos = its first (VX3_Link.cu:31-58)
        const int neg = reverse ? v2 : v1, posv = reverse ? v1 : v2;
        const int l = nl++;
        lneg.push_back(neg);
        lpos.push_back(posv);
        laxis.push_back((uint8_t)axis);
        lmat.push_back(mat);
        vlinks[(size_t)v1 * 6 + dir1] = l;
        vlinks[(size_t)v2 * 6 + dir2] = l;
        pos2.push_back(V3{});
        a1v.push_back(V3{});
        a2v.push_back(V3{});
        fNeg.push_back(V3{});
        fPos.push_back(V3{});
        mNeg.push_back(V3{});
        mPos.push_back(V3{});
        smallAngle.push_back(1);
        velValid.push_back(0);
        strain.push_back(0.0f);
        maxStrain.push_back(0.0f);
        strainOffset.push_back(0.0f);
        lremoved.push_back(0);
        ldetached.push_back(0);
        lNewLink.push_back(d.safety_guard);
        restLen.push_back(0.0);
        update_rest_length(l);
        surfaceChanged = true;
        return true;
    }

    int64_t run(int64_t maxSteps, bool useStop) { // CUDA_Simulation loop (VX3_SimulationManager.cu:62-69, 116-117)
        int64_t done = 0;
        for (; done < maxSteps; done++) {
            if (diverged)
                break;
            if (useStop && stop_condition_met()) {
                stopped = true;
                break;
            }
            if (!do_time_step()) {
                diverged = true;
                break;
            }
        }
        // (into finalCom: `com` is what the stop formula and the heading angle see, refreshed only on sampling steps - a run continued
        // by another call must see what one long run sees)
        const V3 sampled = com;
        update_com();
        finalCom = com;
        com = sampled;
        V3 off = finalCom - initialCom;
        fitness = eval_formula(d.fitness, off.x, off.y, off.z, collisionCount, t, recentAngle, targetCloseness, numClosePairs, nv);
        return done;
    }
};

} // namespace

struct vx3o_world {
    World w;
};

extern "C" {

int vx3o_create(const VX3B_WorldDesc *desc, vx3o_world **out) {
    if (!desc || !out)
        return VX3B_ERR_ARG;
    auto *o = new vx3o_world();
    o->w.init(desc);
    *out = o;
    return VX3B_OK;
}
void vx3o_free(vx3o_world *o) { delete o; }

int64_t vx3o_run(vx3o_world *o, int64_t max_steps, int use_stop_formula) { return o ? o->w.run(max_steps, use_stop_formula != 0) : -1; }

int vx3o_read_state(vx3o_world *o, double *pos3, double *orient4, double *lin3, double *ang3) {
    if (!o)
        return VX3B_ERR_ARG;
    World &w = o->w;
    for (int v = 0; v < w.nv; v++) {
        if (pos3) {
            pos3[v * 3] = w.pos[v].x;
            pos3[v * 3 + 1] = w.pos[v].y;
            pos3[v * 3 + 2] = w.pos[v].z;
        }
        if (orient4) {
            orient4[v * 4] = w.orient[v].w;
            orient4[v * 4 + 1] = w.orient[v].x;
            orient4[v * 4 + 2] = w.orient[v].y;
            orient4[v * 4 + 3] = w.orient[v].z;
        }
        if (lin3) {
            lin3[v * 3] = w.linMom[v].x;
            lin3[v * 3 + 1] = w.linMom[v].y;
            lin3[v * 3 + 2] = w.linMom[v].z;
        }
        if (ang3) {
            ang3[v * 3] = w.angMom[v].x;
            ang3[v * 3 + 1] = w.angMom[v].y;
            ang3[v * 3 + 2] = w.angMom[v].z;
        }
    }
    return VX3B_OK;
}

int vx3o_read_links(vx3o_world *o, double *pos2_3, double *a1v3, double *a2v3, int32_t *flags) {
    if (!o)
        return VX3B_ERR_ARG;
    World &w = o->w;
    for (int l = 0; l < w.nl; l++) {
        if (pos2_3) {
            pos2_3[l * 3] = w.pos2[l].x;
            pos2_3[l * 3 + 1] = w.pos2[l].y;
            pos2_3[l * 3 + 2] = w.pos2[l].z;
        }
        if (a1v3) {
            a1v3[l * 3] = w.a1v[l].x;
            a1v3[l * 3 + 1] = w.a1v[l].y;
            a1v3[l * 3 + 2] = w.a1v[l].z;
        }
        if (a2v3) {
            a2v3[l * 3] = w.a2v[l].x;
            a2v3[l * 3 + 1] = w.a2v[l].y;
            a2v3[l * 3 + 2] = w.a2v[l].z;
        }
        if (flags)
            flags[l] = (w.smallAngle[l] ? 1 : 0) | (w.velValid[l] ? 2 : 0);
    }
    return VX3B_OK;
}

// signals: [nv][6] = localSignal, localSignaldt, inactiveUntil, d_signal.value, d_signal.activeTime, packmakerNextPulse;
// vflags: [nv][2] = removed, enableAttach
int vx3o_read_signals(vx3o_world *o, double *sig6, int32_t *vflags2) {
    if (!o)
        return VX3B_ERR_ARG;
    World &w = o->w;
    for (int v = 0; v < w.nv; v++) {
        if (sig6) {
            double *q = sig6 + 6 * (size_t)v;
            q[0] = w.localSignal[v], q[1] = w.localSignalDt[v], q[2] = w.inactiveUntil[v], q[3] = w.sigValue[v], q[4] = w.sigActiveTime[v], q[5] = w.paceMakerNext[v];
        }
        if (vflags2) {
            vflags2[2 * v] = w.vremoved[v];
            vflags2[2 * v + 1] = w.vEnableAttach[v];
        }
    }
    return VX3B_OK;
}
// current link list (it grows by attachment): *n in = capacity in links, out = number of links; links6 = neg, pos, axis, removed,
// detached, isNewLink per link
int vx3o_read_topology(vx3o_world *o, int32_t *links6, int32_t *n) {
    if (!o || !n)
        return VX3B_ERR_ARG;
    World &w = o->w;
    const int cap = *n;
    *n = w.nl;
    if (links6)
        for (int l = 0; l < w.nl && l < cap; l++) {
            int32_t *q = links6 + 6 * (size_t)l;
            q[0] = w.lneg[l], q[1] = w.lpos[l], q[2] = w.laxis[l], q[3] = w.lremoved[l], q[4] = w.ldetached[l], q[5] = w.lNewLink[l];
        }
    return VX3B_OK;
}

int vx3o_result(vx3o_world *o, VX3B_WorldResult *r) {
    if (!o || !r)
        return VX3B_ERR_ARG;
    World &w = o->w;
    memset(r, 0, sizeof(*r));
    r->current_time = w.t;
    r->fitness_score = w.fitness;
    r->initial_com[0] = w.initialCom.x;
    r->initial_com[1] = w.initialCom.y;
    r->initial_com[2] = w.initialCom.z;
    r->current_com[0] = w.finalCom.x;
    r->current_com[1] = w.finalCom.y;
    r->current_com[2] = w.finalCom.z;
    r->vox_size = w.d.vox_size;
    r->recent_angle = w.recentAngle;
    r->target_closeness = w.targetCloseness;
    r->dt = w.dt;
    r->steps = w.stepCount;
    r->num_voxel = w.nv;
    r->num_close_pairs = w.numClosePairs;
    r->collision_count = w.collisionCount;
    r->diverged = w.diverged;
    r->stopped = w.stopped;
    for (int v = 0; v < w.nv; v++)
        if (w.M(v).isMeasured) { // collectResults (VX3_SimulationManager.cu:442-451)
            r->num_measured_voxel++;
            r->total_distance_of_all_voxels += len(w.pos[v] - w.initialPos[v]);
        }
    return VX3B_OK;
}

int vx3o_collision_pairs(vx3o_world *o, int32_t *pairs2, int32_t *n) {
    if (!o || !n)
        return VX3B_ERR_ARG;
    World &w = o->w;
    int cap = *n, cnt = (int)w.lastPairs.size();
    *n = cnt;
    if (pairs2)
        for (int i = 0; i < cnt && i < cap; i++) {
            pairs2[i * 2] = w.lastPairs[i].first;
            pairs2[i * 2 + 1] = w.lastPairs[i].second;
        }
    return VX3B_OK;
}

int vx3o_set_dt(vx3o_world *o, double dt) { // test hook: fixed step instead of DtFrac * recommendedTimeStep
    if (!o)
        return VX3B_ERR_ARG;
    o->w.optimalDt = dt / (o->w.d.dt_frac != 0 ? o->w.d.dt_frac : 1.0);
    o->w.dt = (float)dt;
    return VX3B_OK;
}

// CVX_External::setForce / setMoment between steps (VoxCAD/src/VX2/test/tVoxelyze.h:471-605, 1033-1160 change the load mid-run)
int vx3o_set_force(vx3o_world *o, int voxel, const double *force3, const double *moment3) {
    if (!o || voxel < 0 || voxel >= o->w.nv)
        return VX3B_ERR_ARG;
    for (int k = 0; k < 3; k++) {
        if (force3)
            o->w.extF[voxel * 3 + k] = (float)force3[k];
        if (moment3)
            o->w.extM[voxel * 3 + k] = (float)moment3[k];
    }
    return VX3B_OK;
}

// CVX_Material::setGlobalDamping between steps (tVoxelyze.h:1111-1158, kineticFriction)
int vx3o_set_global_damping(vx3o_world *o, int mat, double zeta) {
    if (!o || mat < 0 || mat >= (int)o->w.vm.size())
        return VX3B_ERR_ARG;
    o->w.vm[mat].zetaGlobal = (float)zeta;
    return VX3B_OK;
}

int vx3o_num_surface(vx3o_world *o) { return o ? (int)o->w.surface.size() : -1; }

} // extern "C"
