// Host-side construction of the device material tables from the C ABI's material records (shared by the static engine's upload and
// the dynamic-topology engine).  Reference: CVX_MaterialVoxel::updateDerived (VX3/VX3_MaterialVoxel.cu:49-71),
// CVX_MaterialLink::updateDerived (VX3/VX3_MaterialLink.cu:129-148) - the constants arrive precomputed in float, as the reference
// caches them; here they are only multiplied together where the reference multiplies two floats every step.
#pragma once
#include <algorithm>
#include <cstring>

#include "physics.cuh"

namespace vx3b {

// thermal: the world's temperature varies and some material expands (sizes are evaluated every step); varyTemp: it varies at all
inline DVoxMat make_dvoxmat(const VX3B_VoxelMaterial &m, bool thermal, bool varyTemp, float initialTemp) {
    DVoxMat o;
    memset(&o, 0, sizeof(o));
    o.E = m.E;
    o.alphaCTE = m.alphaCTE;
    o.muStatic = m.muStatic;
    o.muKinetic = m.muKinetic;
    o.mass = m.mass;
    o.massInverse = m.massInverse;
    o.momentInertiaInverse = m.momentInertiaInverse;
    o.sqrtMassZetaInt2 = 2 * m.sqrtMass * m.zetaInternal;
    o.globalDampTranslateC = m.zetaGlobal * m._2xSqMxExS;
    o.globalDampRotateC = m.zetaGlobal * m._2xSqIxExSxSxS;
    o.collisionDampTranslateC = m.zetaCollision * m._2xSqMxExS;
    o.penetrationStiffness = (float)(2 * m.E * m.nomSize);
    o.gravityForce = -m.mass * 9.80665f * m.gravMult;
    o.nomSizeF = (float)m.nomSize;
    o.nomSize = m.nomSize;
    o.cilia = m.Cilia;
    // sizes are constant unless the world is thermally actuated (then the kernels evaluate the expressions every step)
    o.sizeIsConst = thermal ? 0 : 1;
    const float T0 = varyTemp ? 0.0f : initialTemp; // varying but not thermal => every alphaCTE == 0
    o.baseSize0 = base_size_expr(m.nomSize, m.alphaCTE, T0);
    o.avgSize0 = base_size_average_expr(o.baseSize0);
    o.floorOffset0 = o.avgSize0 / 2 - m.nomSize / 2;
    o.fixed = m.fixed;
    o.isTarget = m.isTarget;
    o.isMeasured = m.isMeasured;
    return o;
}

inline DLinkMat make_dlinkmat(const VX3B_LinkMaterial &m, const VX3B_VoxelMaterial *voxelMats) {
    DLinkMat o;
    o.E = m.E;
    o.a2 = m.a2;
    o.b1 = m.b1;
    o.b2 = m.b2;
    o.b3 = m.b3;
    o.sqA1 = m.sqA1;
    o.sqA2xIp = m.sqA2xIp;
    o.sqB1 = m.sqB1;
    o.sqB2xFMp = m.sqB2xFMp;
    o.sqB3xIp = m.sqB3xIp;
    o.epsilonFail = m.epsilonFail;
    const float s1 = (float)voxelMats[m.mat1].nomSize, s2 = (float)voxelMats[m.mat2].nomSize;
    o.area = 0.5f * (s1 * s1 + s2 * s2); // currentTransverseArea for nu == 0 (VX3_Link.cu:85-88, VX3_Voxel.cu:496-499)
    o.matPair = (uint32_t)m.mat1 | ((uint32_t)m.mat2 << 16);
    return o;
}

inline DLinkModel make_dlinkmodel(const VX3B_LinkMaterial &m) {
    DLinkModel q;
    memset(&q, 0, sizeof(q));
    q.linear = m.linear;
    q.npts = std::max(2, std::min(m.num_points, VX3B_MAX_SS_POINTS));
    q.E = m.E;
    q.nu = m.nu;
    q.eHat = m.eHat;
    q.epsilonFail = m.epsilonFail;
    for (int i = 0; i < VX3B_MAX_SS_POINTS; i++) {
        q.strain[i] = m.strain_data[i];
        q.stress[i] = m.stress_data[i];
    }
    return q;
}
inline bool link_material_is_nonlinear(const VX3B_LinkMaterial &m) { return !m.linear || m.nu != 0.0f; }

} // namespace vx3b
