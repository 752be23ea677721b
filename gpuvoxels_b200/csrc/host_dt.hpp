// Host arithmetic shared by the engine's upload and the C ABI: the time step of a world.
#pragma once
#include <cmath>

#include "vx3b200.h"

namespace vx3b {

// recommended time step exactly as the reference computes it (VX3: double, VX2: float)
inline double recommended_dt(const VX3B_WorldDesc &d) {
    const bool vx3 = !d.vx2_compat;
    if (vx3) { // VX3/VX3_VoxelyzeKernel.cu:184-217
        double maxF2 = 0.0f;
        for (int l = 0; l < d.num_links; l++) {
            double m1 = d.voxel_mats[d.voxel_mat[d.link_neg[l]]].mass, m2 = d.voxel_mats[d.voxel_mat[d.link_pos[l]]].mass;
            double f2 = d.link_mats[d.link_mat[l]].a1 / (m1 < m2 ? m1 : m2);
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
