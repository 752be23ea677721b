// Host arithmetic shared by the engine's upload and the C ABI: the time step of a world.
#pragma once
#include <cmath>

#include "vx3b200.h"

namespace vx3b {

// temperature of voxel v when the time step is fixed, i.e. after the first updateTemperature() (VX3_VoxelyzeKernel.cu:237-254, 625-650)
inline float first_temperature(const VX3B_WorldDesc &d, int v) {
    const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
    if (!(d.vary_temp_enabled && d.temp_period > 0) || (!d.vx2_compat && m.fixed))
        return d.initial_temp;
    double T = d.temp_amplitude * sin(2 * 3.1415926f * (0.0 / d.temp_period + (d.phase_offset ? d.phase_offset[v] : 0.0)));
    if (!d.enable_expansion && T > 0)
        T = 0;
    return (float)T;
}
// VX3_Link::axialStiffness (VX3_Link.cu:277-286): a1 when Poisson's ratio is zero, else eHat * area / ((strain + 1) * restLength) with
// the link's values at the first step: strain 0, area = size^2 (VX3_Voxel.cu:493-513 at zero strain), rest length at the first temperature
inline float axial_stiffness(const VX3B_WorldDesc &d, int l) {
    const VX3B_LinkMaterial &m = d.link_mats[d.link_mat[l]];
    if (m.nu == 0.0f)
        return m.a1;
    const int a = d.link_neg[l], b = d.link_pos[l];
    const VX3B_VoxelMaterial &ma = d.voxel_mats[d.voxel_mat[a]], &mb = d.voxel_mats[d.voxel_mat[b]];
    const float sa = (float)ma.nomSize, sb = (float)mb.nomSize;
    const float area = 0.5f * (sa * sa + sb * sb);
    const double restLen = 0.5 * (ma.nomSize * (double)(1 + first_temperature(d, a) * ma.alphaCTE) + mb.nomSize * (double)(1 + first_temperature(d, b) * mb.alphaCTE));
    return (float)(m.eHat * area / ((0.0f + 1) * restLen));
}

// recommended time step exactly as the reference computes it (VX3: double, VX2: float)
inline double recommended_dt(const VX3B_WorldDesc &d) {
    const bool vx3 = !d.vx2_compat;
    if (vx3) { // VX3/VX3_VoxelyzeKernel.cu:184-217
        double maxF2 = 0.0f;
        for (int l = 0; l < d.num_links; l++) {
            double m1 = d.voxel_mats[d.voxel_mat[d.link_neg[l]]].mass, m2 = d.voxel_mats[d.voxel_mat[d.link_pos[l]]].mass;
            double f2 = axial_stiffness(d, l) / (m1 < m2 ? m1 : m2);
            if (f2 > maxF2)
                maxF2 = f2;
        }
        if (maxF2 <= 0.0f)
            for (int v = 0; v < d.num_voxels; v++) {
                const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
                double f2 = m.E * m.nomSize / m.mass;
                if (f2 > maxF2)
                    maxF2 = f2;
            }
        if (maxF2 <= 0.0f)
            return 0.0f;
        return 1.0f / (6.283185f * sqrt(maxF2));
    }
    float maxF2 = 0.0f; // old/Voxelyze.cpp:286-311
    for (int l = 0; l < d.num_links; l++) {
        float m1 = d.voxel_mats[d.voxel_mat[d.link_neg[l]]].mass, m2 = d.voxel_mats[d.voxel_mat[d.link_pos[l]]].mass;
        float f2 = d.link_mats[d.link_mat[l]].a1 / (m1 < m2 ? m1 : m2);
        if (f2 > maxF2)
            maxF2 = f2;
    }
    if (maxF2 <= 0.0f)
        for (int v = 0; v < d.num_voxels; v++) {
            const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
            float f2 = (float)(m.E * m.nomSize / m.mass);
            if (f2 > maxF2)
                maxF2 = f2;
        }
    if (maxF2 <= 0.0f)
        return 0.0f;
    return (float)(1.0f / (6.283185f * sqrtf(maxF2)));
}


// the step a world actually takes: VX3 dt = (float)(DtFrac * max(recommended, 1e-10)) (VX3_VoxelyzeKernel.cu:244-254);
// Voxelyze2 callers pass DtFrac * recommendedTimeStep() as a float
inline float world_time_step(const VX3B_WorldDesc &d) {
    double opt = recommended_dt(d);
    if (!d.vx2_compat && opt < 1e-10)
        opt = 1e-10;
    return (float)(d.dt_frac * opt);
}

} // namespace vx3b
