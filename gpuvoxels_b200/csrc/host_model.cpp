#include "host_model.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <sstream>

namespace vx3b {

// ------------------------------------------------------------------------------------------
// lattice / link construction
// ------------------------------------------------------------------------------------------

void HostModel::ensureExt() {
    size_t nv = voxelMat.size();
    dofFixed.resize(nv, 0);
    extForce.resize(nv * 3, 0.0f);
    extMoment.resize(nv * 3, 0.0f);
    extTranslation.resize(nv * 3, 0.0);
    extRotation.resize(nv * 3, 0.0);
}

int HostModel::setVoxel(int mat, int ix, int iy, int iz, std::string *err, bool isGhost) {
    if (mat < 0 || mat >= (int)mats.size()) {
        if (err)
            *err = "set_voxel: no such material index " + std::to_string(mat);
        return -1;
    }
    if (ix < -32768 || ix > 32767 || iy < -32768 || iy > 32767 || iz < -32768 || iz > 32767) {
        if (err)
            *err = "set_voxel: lattice index out of the short range the reference uses (VX3_Voxel.h:272)";
        return -1;
    }
    finalized = false;
    int existing = voxelAt(ix, iy, iz);
    if (existing >= 0) { // CVoxelyze::replaceVoxel: new material, links re-derived at finalize (link order kept)
        voxelMat[existing] = mat;
        return existing;
    }
    int v = numVoxels();
    lattice.insert(lattice.end(), {(int16_t)ix, (int16_t)iy, (int16_t)iz});
    voxelMat.push_back(mat);
    phaseOffset.push_back(0.0);
    baseCilia.insert(baseCilia.end(), 3, 0.0);
    shiftCilia.insert(shiftCilia.end(), 3, 0.0);
    occupied.set(ix, iy, iz, v);
    if (isGhost || !ghost.empty())
        ghost.resize((size_t)v + 1, 0);
    if (isGhost)
        ghost[v] = 1;
    ensureExt();
    // links to already-present neighbours, directions X_POS,X_NEG,Y_POS,Y_NEG,Z_POS,Z_NEG
    // (CVoxelyze::addVoxel, old/Voxelyze.cpp:439-461; end ordering CVX_Link ctor old/VX_Link.cpp:20-49)
    static const int off[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    for (int dir = 0; dir < 6; dir++) {
        int nx = ix + off[dir][0], ny = iy + off[dir][1], nz = iz + off[dir][2];
        int o = voxelAt(nx, ny, nz);
        if (o < 0)
            continue;
        if (isGhost && ghost[o])
            continue; // a link between two ghosts belongs to the neighbouring slab alone
        int axis = dir / 2;
        bool positive = (dir % 2) == 0; // neighbour lies on our positive side -> we are the negative end
        int neg = positive ? v : o, pos = positive ? o : v;
        // (v is new, so no link between the two exists yet)
        linkNeg.push_back(neg);
        linkPos.push_back(pos);
        linkAxis.push_back((uint8_t)axis);
        linkMat.push_back(-1); // resolved at finalize (materials may still change)
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// material derivation (float arithmetic mirrors the reference expression by expression)
// ------------------------------------------------------------------------------------------

static VX3B_VoxelMaterial derive_voxel_material(const MatInput &in, float nu, double nomSize, float gravMult) {
    VX3B_VoxelMaterial o;
    memset(&o, 0, sizeof(o));
    o.E = in.E;
    o.nu = nu;
    o.rho = in.rho;
    o.alphaCTE = in.alphaCTE;
    o.muStatic = in.muStatic;
    o.muKinetic = in.muKinetic;
    o.zetaInternal = in.zetaInternal;
    o.zetaGlobal = in.zetaGlobal;
    o.zetaCollision = in.zetaCollision;
    o.epsilonFail = in.epsilonFail;
    o.gravMult = gravMult;
    o.nomSize = nomSize;
    o.Cilia = in.Cilia;
    o.matid = in.matid;
    o.rgba = in.rgba;
    o.fixed = in.fixed;
    o.isTarget = in.isTarget;
    o.isMeasured = in.isMeasured;
    o.sticky = in.sticky;
    o.isPaceMaker = in.isPaceMaker;
    o.PaceMakerPeriod = in.PaceMakerPeriod;
    o.signalValueDecay = in.signalValueDecay;
    o.signalTimeDelay = in.signalTimeDelay;
    o.inactivePeriod = in.inactivePeriod;
    o.RemoveFromSimulationAfterThisManySeconds = in.removeAfter;
    o.TurnOnThermalExpansionAfterThisManySeconds = in.thermalAfter;
    o.TurnOnCiliaAfterThisManySeconds = in.ciliaAfter;
    o.linear = 1;
    // CVX_MaterialVoxel::updateDerived (old/VX_MaterialVoxel.cpp:58-80)
    double volume = nomSize * nomSize * nomSize;
    o.mass = (float)(volume * in.rho);
    o.momentInertia = (float)(o.mass * nomSize * nomSize / 6.0f);
    if (volume == 0 || o.mass == 0 || o.momentInertia == 0) {
        o.massInverse = o.sqrtMass = o.momentInertiaInverse = o._2xSqMxExS = o._2xSqIxExSxSxS = 0.0f;
        return o;
    }
    o.massInverse = 1.0f / o.mass;
    o.sqrtMass = sqrtf(o.mass);
    o.momentInertiaInverse = 1.0f / o.momentInertia;
    o._2xSqMxExS = (float)(2.0f * sqrt(o.mass * in.E * nomSize));
    o._2xSqIxExSxSxS = (float)(2.0f * sqrt(o.momentInertia * in.E * nomSize * nomSize * nomSize));
    return o;
}

// ---- stress/strain models: CVX_Material::setModelLinear / setModelBilinear / setModel (old/VX_Material.cpp:300-449) ----------
bool mat_set_model_linear(MatInput &m, float youngsModulus, float failureStress, std::string *err) {
    if (youngsModulus <= 0) {
        if (err)
            *err = "Young's modulus must be positive";
        return false;
    }
    if (failureStress != -1.0f && failureStress <= 0) {
        if (err)
            *err = "Failure stress must be positive";
        return false;
    }
    float tmpFailureStress = failureStress; // a dummy failure stress if none was provided
    if (tmpFailureStress == -1)
        tmpFailureStress = 1000000;
    const float tmpFailStrain = tmpFailureStress / youngsModulus;
    m.strainData = {0.0f, tmpFailStrain};
    m.stressData = {0.0f, tmpFailureStress};
    m.linear = true;
    m.E = youngsModulus;
    m.sigmaFail = failureStress;
    m.epsilonFail = (failureStress == -1) ? -1 : tmpFailStrain;
    return true;
}
bool mat_set_model_bilinear(MatInput &m, float youngsModulus, float plasticModulus, float yieldStress, float failureStress, std::string *err) {
    const char *why = nullptr;
    if (youngsModulus <= 0)
        why = "Young's modulus must be positive";
    else if (plasticModulus <= 0 || plasticModulus >= youngsModulus)
        why = "Plastic modulus must be positive but less than Young's modulus";
    else if (yieldStress <= 0)
        why = "Yield stress must be positive";
    else if (failureStress != -1.0f && failureStress <= yieldStress)
        why = "Failure stress must be positive and greater than the yield stress";
    if (why) {
        if (err)
            *err = why;
        return false;
    }
    const float yieldStrain = yieldStress / youngsModulus;
    float tmpFailureStress = failureStress;
    if (tmpFailureStress == -1)
        tmpFailureStress = 3 * yieldStress;
    const float tM = plasticModulus;
    const float tB = yieldStress - tM * yieldStrain; // y - m x = b
    const float tmpFailStrain = (tmpFailureStress - tB) / tM;
    m.strainData = {0.0f, yieldStrain, tmpFailStrain};
    m.stressData = {0.0f, yieldStress, tmpFailureStress};
    m.linear = false;
    m.E = youngsModulus;
    m.sigmaFail = failureStress;
    m.epsilonFail = failureStress == -1.0f ? -1.0f : tmpFailStrain;
    return true;
}
bool mat_set_model_data(MatInput &m, int n, const float *strain, const float *stress, std::string *err) {
    if (n > 0 && *strain == 0 && *stress == 0) { // a leading (0, 0) is implied
        strain++;
        stress++;
        n--;
    }
    const char *why = nullptr;
    if (n <= 0)
        why = "Not enough data points";
    else if (*strain <= 0 || *stress <= 0)
        why = "First stress and strain data points negative or zero";
    std::vector<float> sd{0.0f}, td{0.0f};
    float sweepStrain = 0.0f, sweepStress = 0.0f;
    for (int i = 0; i < n && !why; i++) {
        const float thisStrain = strain[i], thisStress = stress[i];
        if (thisStrain <= sweepStrain)
            why = "Out of order strain data";
        // (the reference compares with element 0 of both arrays, i.e. 0/0 = NaN: the slope test never fires - kept)
        else if (i > 0 && (thisStress - sweepStress) / (thisStrain - sweepStrain) > td[0] / sd[0])
            why = "Slope of stress/strain curve should never exceed that of the first line segment (youngs modulus)";
        sweepStrain = thisStrain;
        sweepStress = thisStress;
        sd.push_back(thisStrain);
        td.push_back(thisStress);
    }
    if (!why && (int)sd.size() > VX3B_MAX_SS_POINTS)
        why = "too many stress/strain data points (VX3B_MAX_SS_POINTS)";
    if (why) {
        if (err)
            *err = why;
        return false;
    }
    m.strainData = sd;
    m.stressData = td;
    m.E = td[1] / sd[1];
    m.sigmaFail = td.back();
    m.epsilonFail = sd.back();
    m.linear = (n == 1);
    return true;
}
// CVX_Material::modulus / strain (old/VX_Material.cpp:186-220)
static float model_modulus(const MatInput &m, float strain) {
    if (m.epsilonFail != -1.0f && strain > m.epsilonFail)
        return 0.0f;
    if (strain <= m.strainData[1] || m.linear)
        return m.E;
    const int n = (int)m.strainData.size();
    for (int i = 2; i < n; i++)
        if (strain <= m.strainData[i] || i == n - 1)
            return (m.stressData[i] - m.stressData[i - 1]) / (m.strainData[i] - m.strainData[i - 1]);
    return 0.0f;
}
static float model_strain_at(const MatInput &m, float stress) {
    if (stress <= m.stressData[1] || m.linear)
        return stress / m.E;
    const int n = (int)m.strainData.size();
    for (int i = 2; i < n; i++)
        if (stress <= m.stressData[i] || i == n - 1) {
            const float perc = (stress - m.stressData[i - 1]) / (m.stressData[i] - m.stressData[i - 1]);
            return m.strainData[i - 1] + perc * (m.strainData[i] - m.strainData[i - 1]);
        }
    return 0.0f;
}

// CVX_MaterialLink::updateAll + updateDerived (old/VX_MaterialLink.cpp:53-139 = VX3/VX3_MaterialLink.cu:53-148)
static bool derive_link_material(const VX3B_VoxelMaterial &m1, const MatInput &i1, const VX3B_VoxelMaterial &m2, const MatInput &i2, int a, int b,
                                 VX3B_LinkMaterial *out, std::string *err) {
    VX3B_LinkMaterial o;
    memset(&o, 0, sizeof(o));
    o.mat1 = a;
    o.mat2 = b;
    double nomSize = 0.5 * (m1.nomSize + m2.nomSize);
    float f1 = i1.sigmaFail, f2 = i2.sigmaFail, stressFail;
    if (f1 == -1.0f)
        stressFail = f2;
    else if (f2 == -1.0f)
        stressFail = f1;
    else
        stressFail = f1 < f2 ? f1 : f2;
    MatInput lm; // the link's own model
    if (i1.linear && i2.linear) {
        if (!mat_set_model_linear(lm, 2.0f * m1.E * m2.E / (m1.E + m2.E), stressFail, err))
            return false;
    } else { // at least one bilinear or data-based end: step up through the ascending strain points of both, springs in series
        std::vector<float> newStress{0.0f}, newStrain{0.0f};
        int it1 = 1, it2 = 1;
        while (it1 < (int)i1.strainData.size() && it2 < (int)i2.strainData.size()) {
            float strain = FLT_MAX;
            if (it1 < (int)i1.strainData.size())
                strain = i1.strainData[it1];
            if (it2 < (int)i2.strainData.size() && i2.strainData[it2] < strain)
                strain = i2.strainData[it2];
            if (strain == i1.strainData[it1])
                it1++;
            if (strain == i2.strainData[it2])
                it2++;
            const float modulus1 = model_modulus(i1, strain - FLT_EPSILON), modulus2 = model_modulus(i2, strain - FLT_EPSILON);
            const float thisModulus = 2.0f * modulus1 * modulus2 / (modulus1 + modulus2);
            const int last = (int)newStrain.size() - 1;
            const float s0 = newStrain[last], t0 = newStress[last];
            newStrain.push_back(strain);
            newStress.push_back(t0 + thisModulus * (strain - s0));
        }
        if (!mat_set_model_data(lm, (int)newStrain.size(), newStrain.data(), newStress.data(), err))
            return false;
        lm.sigmaFail = stressFail; // no failure specified before: possible in combinations of linear and bilinear materials
        lm.epsilonFail = stressFail == -1.0f ? -1.0f : model_strain_at(lm, stressFail);
    }
    o.E = lm.E;
    o.epsilonFail = lm.epsilonFail;
    o.linear = lm.linear ? 1 : 0;
    o.num_points = (int)lm.strainData.size();
    for (int k = 0; k < o.num_points; k++) {
        o.strain_data[k] = lm.strainData[k];
        o.stress_data[k] = lm.stressData[k];
    }
    // Poisson's ratio: chosen so that eHat ends up as the series spring of the two ends' eHat
    if (m1.nu == 0 && m2.nu == 0)
        o.nu = 0;
    else {
        const float eHat1 = m1.E / ((1 - 2 * m1.nu) * (1 + m1.nu)), eHat2 = m2.E / ((1 - 2 * m2.nu) * (1 + m2.nu));
        const float tmpEHat = 2 * eHat1 * eHat2 / (eHat1 + eHat2);
        const float tmpE = o.E;
        const float c2 = (tmpEHat - tmpE) / (2 * tmpEHat) + 0.0625; // (nu + 0.25)^2 = c2
        o.nu = (float)(sqrtf(c2) - 0.25); // float root (std::sqrt(float) in the reference), then the double subtraction
    }
    o.eHat = o.E / ((1 - 2 * o.nu) * (1 + o.nu));
    float L = (float)nomSize;
    o.a1 = o.E * L;
    o.a2 = o.E * L * L * L / (12.0f * (1 + o.nu));
    o.b1 = o.E * L;
    o.b2 = o.E * L * L / 2.0f;
    o.b3 = o.E * L * L * L / 6.0f;
    o.sqA1 = sqrtf(o.a1);
    o.sqA2xIp = sqrtf(o.a2 * L * L / 6.0f);
    o.sqB1 = sqrtf(o.b1);
    o.sqB2xFMp = sqrtf(o.b2 * L / 2.0f);
    o.sqB3xIp = sqrtf(o.b3 * L * L / 6.0f);
    *out = o;
    return true;
}

bool HostModel::finalize(std::string *err) {
    voxelMats.clear();
    for (auto &m : mats) {
        float nu = volumeEffects ? m.nuMemory : 0.0f; // CVX_Sim::EnableVolumeEffects, VX_Sim.cpp:445-458
        if (nu != 0.0f && vx2Compat) {
            // Voxelyze2 re-derives every link's transverse area and strain sum each step from neighbour strains the same sweep is
            // updating (CVX_Link::updateTransverseInfo under `omp parallel for`): order dependent, not restated.  Voxelyze3 compiles
            // that update out (VX3_Link.cu:147-150) and keeps the initial values; that form is supported.
            if (err)
                *err = "Poisson/volume effects with Voxelyze2 semantics (vx2_compat with nu != 0) are not supported (SURVEY.md §8f-4)";
            return false;
        }
        voxelMats.push_back(derive_voxel_material(m, nu, voxSize, gravMult));
        voxelMats.back().linear = m.linear ? 1 : 0;
    }
    linkMatPairs.clear();
    linkMats.clear();
    for (auto &pr : presetLinkMatPairs) { // a slab: the whole body's combined materials, in the whole body's order
        linkMatPairs.push_back(pr);
        VX3B_LinkMaterial lm;
        if (!derive_link_material(voxelMats[pr.first], mats[pr.first], voxelMats[pr.second], mats[pr.second], pr.first, pr.second, &lm, err))
            return false;
        linkMats.push_back(lm);
    }
    for (size_t l = 0; l < linkNeg.size(); l++) {
        // CVoxelyze::combinedMaterial (old/Voxelyze.cpp:614-628): first use creates, later uses find (either order)
        int a = voxelMat[linkNeg[l]], b = voxelMat[linkPos[l]];
        int found = -1;
        for (size_t k = 0; k < linkMatPairs.size(); k++)
            if ((linkMatPairs[k].first == a && linkMatPairs[k].second == b) || (linkMatPairs[k].first == b && linkMatPairs[k].second == a)) {
                found = (int)k;
                break;
            }
        if (found < 0) {
            found = (int)linkMatPairs.size();
            linkMatPairs.emplace_back(a, b);
            VX3B_LinkMaterial lm;
            if (!derive_link_material(voxelMats[a], mats[a], voxelMats[b], mats[b], a, b, &lm, err))
                return false;
            linkMats.push_back(lm);
        }
        linkMat[l] = found;
    }
    if (enableAttach) {
        // Attachment links two voxels of one sticky material (VX3_VoxelyzeKernel.cu:734-741) with the material
        // VX3_VoxelyzeKernel::combinedMaterial finds or makes on the device (:515-528); list those up front, after the used ones
        for (int a = 0; a < (int)voxelMats.size(); a++) {
            if (!voxelMats[a].sticky || voxelMats[a].fixed)
                continue;
            bool have = false;
            for (auto &pr : linkMatPairs)
                have = have || (pr.first == a && pr.second == a);
            if (have)
                continue;
            linkMatPairs.emplace_back(a, a);
            VX3B_LinkMaterial lm;
            if (!derive_link_material(voxelMats[a], mats[a], voxelMats[a], mats[a], a, a, &lm, err))
                return false;
            linkMats.push_back(lm);
        }
    }
    auto pad = [](std::vector<VX3B_Token> &t) {
        if (t.size() > VX3B_MAX_TOKENS)
            return false;
        VX3B_Token e;
        memset(&e, 0, sizeof(e));
        t.resize(VX3B_MAX_TOKENS, e);
        return true;
    };
    if (!pad(stopCondition) || !pad(fitness) || !pad(forceField[0]) || !pad(forceField[1]) || !pad(forceField[2]) || !pad(attachCondition[0]) ||
        !pad(attachCondition[1]) || !pad(attachCondition[2]) || !pad(attachCondition[3]) || !pad(attachCondition[4])) {
        if (err)
            *err = "formula longer than VX3B_MAX_TOKENS";
        return false;
    }
    ensureExt();
    finalized = true;
    return true;
}

void HostModel::fillDesc(VX3B_WorldDesc *d) const {
    memset(d, 0, sizeof(*d));
    d->num_voxels = numVoxels();
    d->num_links = (int)linkNeg.size();
    d->num_voxel_mats = (int)voxelMats.size();
    d->num_link_mats = (int)linkMats.size();
    d->vox_size = voxSize;
    d->lattice = lattice.data();
    d->voxel_mat = voxelMat.data();
    d->phase_offset = phaseOffset.data();
    d->base_cilia = baseCilia.data();
    d->shift_cilia = shiftCilia.data();
    d->dof_fixed = dofFixed.data();
    d->ext_force = extForce.data();
    d->ext_moment = extMoment.data();
    d->ext_translation = extTranslation.data();
    d->ext_rotation = extRotation.data();
    d->link_neg = linkNeg.data();
    d->link_pos = linkPos.data();
    d->link_axis = linkAxis.data();
    d->link_mat = linkMat.data();
    d->ghost = ghost.empty() ? nullptr : ghost.data();
    d->voxel_mats = voxelMats.data();
    d->link_mats = linkMats.data();
    d->dt_frac = dtFrac;
    d->floor_enabled = floorEnabled;
    d->vary_temp_enabled = varyTempEnabled;
    d->temp_amplitude = tempAmplitude;
    d->temp_period = tempPeriod;
    d->initial_temp = initialTemp;
    d->enable_collision = enableCollision;
    d->enable_cilia = enableCilia;
    d->enable_expansion = enableExpansion;
    d->vx2_compat = vx2Compat;
    d->max_dist_in_voxel_lengths_to_count_as_pair = maxDistPair;
    d->save_position_of_all_voxels = savePositions;
    d->record_step_size = recordStepSize;
    d->record_voxel = recordVoxel;
    d->record_link = recordLink;
    memcpy(d->stop_condition, stopCondition.data(), sizeof(d->stop_condition));
    memcpy(d->fitness, fitness.data(), sizeof(d->fitness));
    for (int k = 0; k < 3; k++)
        memcpy(d->force_field[k], forceField[k].data(), sizeof(d->force_field[k]));
    for (int k = 0; k < 5; k++)
        memcpy(d->attach_condition[k], attachCondition[k].data(), sizeof(d->attach_condition[k]));
    d->enable_signals = enableSignals;
    d->enable_attach = enableAttach;
    d->enable_detach = enableDetach;
    d->safety_guard = safetyGuard;
    d->secondary_experiment = secondaryExperiment;
    d->reinitialize_initial_position_after = reinitializeAfter;
}

// ------------------------------------------------------------------------------------------
// formulas
// ------------------------------------------------------------------------------------------

bool parse_math_tree(const XmlNode *root, std::vector<VX3B_Token> &out, std::string *err) {
    out.clear();
    if (!root)
        return true; // missing formula: empty program (the reference leaves token[0]=mtEND)
    static const std::pair<const char *, int> ops[] = {
        {"mtEND", VX3B_MT_END},       {"mtCONST", VX3B_MT_CONST},   {"mtE", VX3B_MT_E},
        {"mtPI", VX3B_MT_PI},         {"mtVAR", VX3B_MT_VAR},       {"mtADD", VX3B_MT_ADD},
        {"mtSUB", VX3B_MT_SUB},       {"mtMUL", VX3B_MT_MUL},       {"mtDIV", VX3B_MT_DIV},
        {"mtPOW", VX3B_MT_POW},       {"mtSQRT", VX3B_MT_SQRT},     {"mtSIN", VX3B_MT_SIN},
        {"mtCOS", VX3B_MT_COS},       {"mtTAN", VX3B_MT_TAN},       {"mtATAN", VX3B_MT_ATAN},
        {"mtLOG", VX3B_MT_LOG},       {"mtINT", VX3B_MT_INT},       {"mtABS", VX3B_MT_ABS},
        {"mtNOT", VX3B_MT_NOT},       {"mtGREATERTHAN", VX3B_MT_GREATERTHAN}, {"mtLESSTHAN", VX3B_MT_LESSTHAN},
        {"mtAND", VX3B_MT_AND},       {"mtOR", VX3B_MT_OR},         {"mtNORMALCDF", VX3B_MT_NORMALCDF}};
    static const char *vars[] = {"x", "y", "z", "hit", "t", "angle", "targetCloseness", "numClosePairs", "num_voxel"};
    // level-order walk, tokens collected then reversed, END last (VX3_SimulationManager.cu:157-186)
    std::vector<VX3B_Token> level;
    std::deque<const XmlNode *> frontier;
    frontier.push_back(root);
    while (!frontier.empty()) {
        const XmlNode *t = frontier.front();
        frontier.pop_front();
        for (auto &c : t->children) {
            VX3B_Token tok;
            memset(&tok, 0, sizeof(tok));
            int op = -1;
            for (auto &o : ops)
                if (c->name == o.first)
                    op = o.second;
            if (op < 0) {
                if (err)
                    *err = "formula: token operation not implemented: " + c->name;
                return false;
            }
            tok.op = op;
            std::string val = xml_trim(c->text);
            if (op == VX3B_MT_VAR) {
                int id = -1;
                for (int k = 0; k < 9; k++)
                    if (val == vars[k])
                        id = k;
                if (id < 0) {
                    if (err)
                        *err = "formula: no such variable: " + val;
                    return false;
                }
                tok.value = id;
            } else if (op == VX3B_MT_CONST) {
                char *end = nullptr;
                tok.value = strtod(val.c_str(), &end);
                if (end == val.c_str()) {
                    if (err)
                        *err = "formula: mtCONST with no number";
                    return false;
                }
            }
            level.push_back(tok);
            frontier.push_back(c.get());
        }
    }
    out.assign(level.rbegin(), level.rend());
    VX3B_Token e;
    memset(&e, 0, sizeof(e));
    out.push_back(e);
    if (out.size() > VX3B_MAX_TOKENS) {
        if (err)
            *err = "formula longer than VX3B_MAX_TOKENS tokens";
        return false;
    }
    return true;
}

double eval_math_tree(const VX3B_Token *tok, double x, double y, double z, double hit, double t, double angle, double closeness,
                      int numClosePairs, int numVoxel) {
    double values[VX3B_MAX_TOKENS + 1];
    values[0] = 0.0; // empty program evaluates to 0 (the reference reads an uninitialised slot here)
    int vc = 0, pc = 0;
    for (int i = 0; i < VX3B_MAX_TOKENS; i++) {
        double v = 0;
        switch (tok[i].op) {
        case VX3B_MT_END:
            return values[pc];
        case VX3B_MT_CONST:
            v = tok[i].value;
            break;
        case VX3B_MT_E:
            v = 2.71828182845904523536;
            break;
        case VX3B_MT_PI:
            v = 3.14159265358979323846;
            break;
        case VX3B_MT_VAR: {
            double id = tok[i].value;
            v = id < 0.5 ? x : id < 1.5 ? y : id < 2.5 ? z : id < 3.5 ? hit : id < 4.5 ? t : id < 5.5 ? angle : id < 6.5 ? closeness
                : id < 7.5 ? (double)numClosePairs : (double)numVoxel;
            break;
        }
        case VX3B_MT_SIN: v = sin(values[pc]); pc++; break;
        case VX3B_MT_COS: v = cos(values[pc]); pc++; break;
        case VX3B_MT_TAN: v = tan(values[pc]); pc++; break;
        case VX3B_MT_ATAN: v = atan(values[pc]); pc++; break;
        case VX3B_MT_LOG: v = log(values[pc]); pc++; break;
        case VX3B_MT_INT: v = rint(values[pc]); pc++; break;
        case VX3B_MT_NORMALCDF: v = 0.5 * erfc(-values[pc] * 0.70710678118654752440); pc++; break;
        case VX3B_MT_SQRT: v = sqrt(values[pc]); pc++; break;
        case VX3B_MT_ABS: v = fabs(values[pc]); pc++; break;
        case VX3B_MT_NOT: v = !values[pc]; pc++; break;
        case VX3B_MT_ADD: v = values[pc + 1] + values[pc]; pc += 2; break;
        case VX3B_MT_SUB: v = values[pc + 1] - values[pc]; pc += 2; break;
        case VX3B_MT_MUL: v = values[pc + 1] * values[pc]; pc += 2; break;
        case VX3B_MT_DIV: v = values[pc + 1] / values[pc]; pc += 2; break;
        case VX3B_MT_POW: v = pow(values[pc + 1], values[pc]); pc += 2; break;
        case VX3B_MT_GREATERTHAN: v = values[pc + 1] > values[pc]; pc += 2; break;
        case VX3B_MT_LESSTHAN: v = values[pc + 1] < values[pc]; pc += 2; break;
        case VX3B_MT_AND: v = values[pc + 1] && values[pc]; pc += 2; break;
        case VX3B_MT_OR: v = values[pc + 1] || values[pc]; pc += 2; break;
        default:
            return -1;
        }
        if (pc > vc)
            return -1; // token sequence error
        values[vc++] = v;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// VXA / VXD
// ------------------------------------------------------------------------------------------

bool merge_vxd(XmlNode &baseDoc, const XmlNode &vxdDoc, std::string *err) {
    const XmlNode *vxd = vxdDoc.child("VXD");
    if (!vxd)
        vxd = &vxdDoc; // ptree_merge: source.get_child("VXD", source)
    for (auto &c : vxd->children) {
        const std::string *rep = c->attr("replace");
        if (!rep || rep->empty())
            continue;
        xml_put_child(baseDoc, *rep, c->clone());
    }
    (void)err;
    return true;
}

namespace {

bool get_text(const XmlNode *parent, const char *tag, std::string *out) {
    if (!parent)
        return false;
    const XmlNode *c = parent->child(tag);
    if (!c)
        return false;
    std::string t = xml_trim(c->text);
    if (t.empty())
        return false; // CXML_Rip::FindLoadElement fails on an element without a text child
    *out = t;
    return true;
}
double get_double(const XmlNode *p, const char *tag, double dflt) {
    std::string t;
    return get_text(p, tag, &t) ? atof(t.c_str()) : dflt;
}
float get_float(const XmlNode *p, const char *tag, float dflt) {
    std::string t;
    return get_text(p, tag, &t) ? (float)atof(t.c_str()) : dflt;
}
int get_int(const XmlNode *p, const char *tag, int dflt) {
    std::string t;
    return get_text(p, tag, &t) ? atoi(t.c_str()) : dflt;
}
bool get_bool(const XmlNode *p, const char *tag, bool dflt) { return get_int(p, tag, dflt ? 1 : 0) != 0; }

// boost ptree get<bool>/get<int> on the merged tree (VX3_SimulationManager.cu:323-375)
int ptree_int(const XmlNode &vxa, const char *path, int dflt) {
    const XmlNode *n = vxa.find(path);
    if (!n)
        return dflt;
    std::string t = xml_trim(n->text);
    if (t.empty())
        return dflt;
    if (t == "true")
        return 1;
    if (t == "false")
        return 0;
    return atoi(t.c_str());
}
double ptree_double(const XmlNode &vxa, const char *path, double dflt) {
    const XmlNode *n = vxa.find(path);
    if (!n)
        return dflt;
    std::string t = xml_trim(n->text);
    return t.empty() ? dflt : atof(t.c_str());
}

std::vector<std::string> split_commas(const std::string &s) { // split(RawData, ',') as std::getline does
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ','))
        out.push_back(item);
    return out;
}

std::vector<const XmlNode *> layers_of(const XmlNode *parent) {
    std::vector<const XmlNode *> v;
    if (parent)
        for (auto &c : parent->children)
            if (c->name == "Layer")
                v.push_back(c.get());
    return v;
}

} // namespace

std::unique_ptr<HostModel> model_from_vxa(const std::string &vxaText, const char *vxdText, std::string *err, const SlabCut *cut) {
    std::string perr;
    auto doc = xml_parse(vxaText, &perr);
    if (!doc) {
        if (err)
            *err = "VXA: " + perr;
        return nullptr;
    }
    if (vxdText && *vxdText) {
        auto vxd = xml_parse(vxdText, &perr);
        if (!vxd) {
            if (err)
                *err = "VXD: " + perr;
            return nullptr;
        }
        merge_vxd(*doc, *vxd, err);
    }
    const XmlNode *vxa = doc->child("VXA");
    if (!vxa) {
        if (err)
            *err = "no <VXA> root element";
        return nullptr;
    }
    auto M = std::make_unique<HostModel>();
    M->mergedVxa = xml_write(*vxa, true);

    // ---- Simulator (CVX_Sim::ReadXML, VX_Sim.cpp:185-294) + VX3-only settings (VX3_SimulationManager.cu:323-375)
    const XmlNode *sim = vxa->child("Simulator");
    float bondZ = 0.1f, colZ = 1.0f, slowZ = 1.0f;
    if (sim) {
        if (const XmlNode *integ = sim->child("Integration"))
            M->dtFrac = get_double(integ, "DtFrac", 0.9);
        if (const XmlNode *damp = sim->child("Damping")) {
            bondZ = get_float(damp, "BondDampingZ", 0.1f);
            colZ = get_float(damp, "ColDampingZ", 1.0f);
            slowZ = get_float(damp, "SlowDampingZ", 1.0f);
        }
        if (const XmlNode *feat = sim->child("Features"))
            M->volumeEffects = get_bool(feat, "VolumeEffectsEnabled", false);
    }
    M->enableCollision = ptree_int(*vxa, "Simulator.AttachDetach.EnableCollision", 1) != 0;
    // signals and topology changes (VX3_SimulationManager.cu:325-371): stepped by the dynamic-topology kernel
    M->enableAttach = ptree_int(*vxa, "Simulator.AttachDetach.EnableAttach", 0) != 0;
    M->enableDetach = ptree_int(*vxa, "Simulator.AttachDetach.EnableDetach", 0) != 0;
    M->safetyGuard = ptree_int(*vxa, "Simulator.AttachDetach.SafetyGuard", 500);
    M->enableSignals = ptree_int(*vxa, "Simulator.EnableSignals", 0) != 0;
    M->secondaryExperiment = ptree_int(*vxa, "Simulator.SecondaryExperiment", 0) != 0;
    M->reinitializeAfter = ptree_double(*vxa, "Simulator.ReinitializeInitialPositionAfterThisManySeconds", 0.0);
    for (int c = 0; c < 5; c++)
        if (!parse_math_tree(vxa->find("Simulator.AttachDetach.AttachCondition.Condition_" + std::to_string(c)), M->attachCondition[c], err))
            return nullptr;
    M->enableCilia = ptree_int(*vxa, "Simulator.EnableCilia", 0) != 0;
    M->enableExpansion = ptree_int(*vxa, "Simulator.EnableExpansion", 0) != 0;
    M->savePositions = ptree_int(*vxa, "Simulator.SavePositionOfAllVoxels", 0) != 0;
    M->recordStepSize = ptree_int(*vxa, "Simulator.RecordHistory.RecordStepSize", 0);
    M->recordLink = ptree_int(*vxa, "Simulator.RecordHistory.RecordLink", 0);
    M->recordVoxel = ptree_int(*vxa, "Simulator.RecordHistory.RecordVoxel", 1);
    M->maxDistPair = ptree_double(*vxa, "Simulator.MaxDistInVoxelLengthsToCountAsPair", 0.0);
    if (!parse_math_tree(vxa->find("Simulator.StopCondition.StopConditionFormula"), M->stopCondition, err) ||
        !parse_math_tree(vxa->find("Simulator.FitnessFunction"), M->fitness, err) ||
        !parse_math_tree(vxa->find("Simulator.ForceField.x_forcefield"), M->forceField[0], err) ||
        !parse_math_tree(vxa->find("Simulator.ForceField.y_forcefield"), M->forceField[1], err) ||
        !parse_math_tree(vxa->find("Simulator.ForceField.z_forcefield"), M->forceField[2], err))
        return nullptr;

    // ---- Environment (CVX_Environment::ReadXML, old/VX_Environment.cpp:101-174; ctor defaults :26-35)
    bool gravEnabled = true, floorEnabled = true;
    double gravAcc = -9.81;
    M->tempEnabled = true;
    const XmlNode *env = vxa->child("Environment");
    if (env) {
        if (const XmlNode *bc = env->child("Boundary_Conditions")) {
            if (get_int(bc, "NumBCs", 0) != 0) {
                if (err)
                    *err = "Boundary_Conditions/FRegion are not read by this host model yet; use vx3b_model_set_fixed / set_force";
                return nullptr;
            }
        }
        if (const XmlNode *g = env->child("Gravity")) {
            gravEnabled = get_bool(g, "GravEnabled", false);
            gravAcc = get_double(g, "GravAcc", -9.81);
            floorEnabled = get_bool(g, "FloorEnabled", false);
        }
        if (const XmlNode *th = env->child("Thermal")) {
            M->tempEnabled = get_bool(th, "TempEnabled", false);
            M->tempBase = get_double(th, "TempBase", 25);
            std::string t;
            if (get_text(th, "TempAmplitude", &t))
                M->tempAmplitude = atof(t.c_str());
            else if (get_text(th, "TempAmp", &t))
                M->tempAmplitude = atof(t.c_str()) - M->tempBase;
            else
                M->tempAmplitude = 0;
            M->varyTempEnabled = get_bool(th, "VaryTempEnabled", false);
            M->tempPeriod = get_double(th, "TempPeriod", 0.1);
        }
    }
    (void)gravEnabled; // Import applies GravAcc whether or not GravEnabled is set (VX_Sim.cpp:67-68, :485)
    M->gravMult = (float)(-(float)gravAcc / 9.80665);
    M->floorEnabled = floorEnabled;
    // initial temperature: CVX_Sim::UpdateMatTemps at Import (VX_Sim.cpp:146-147, 466-483)
    if (M->tempEnabled) {
        if (M->varyTempEnabled)
            M->initialTemp = (M->tempPeriod == 0) ? (float)(0.0f - M->tempBase) : (float)((float)(M->tempBase + M->tempAmplitude * sin(0.0)) - M->tempBase);
        else
            M->initialTemp = (float)((float)(M->tempBase + M->tempAmplitude) - M->tempBase);
    }

    // ---- VXC (CVX_Object::ReadXML): lattice, palette, structure
    const XmlNode *vxc = vxa->child("VXC");
    if (!vxc)
        vxc = vxa->child("DMF");
    if (!vxc) {
        if (err)
            *err = "no <VXC> element";
        return nullptr;
    }
    M->voxSize = get_double(vxc->child("Lattice"), "Lattice_Dim", 0.001);
    if (M->voxSize <= 0)
        M->voxSize = FLT_MIN; // CVX_MaterialVoxel::setNominalSize

    std::vector<int> paletteToMat; // palette slot (1-based) -> model material
    if (const XmlNode *pal = vxc->child("Palette")) {
        for (auto &pm : pal->children) {
            if (pm->name != "Material")
                continue;
            MatInput m;
            m.matid = (int)paletteToMat.size() + 1; // CVX_Object::AddMat: matid = palette index (VX_Object.cpp:486)
            int matType = get_int(pm.get(), "MatType", 0);
            const XmlNode *mech = pm->child("Mechanical");
            if (matType != 0) {
                if (err)
                    *err = "palette material with MatType != SINGLE is not supported";
                return nullptr;
            }
            int matModel = get_int(mech, "MatModel", 0);
            double E = get_double(mech, "Elastic_Mod", 0), failStress = get_double(mech, "Fail_Stress", 0);
            // CVX_Sim::CopyMat (VX_Sim.cpp:368-420) -> CVX_Material::setModel* (old/VX_Material.cpp:300-449)
            std::string merr;
            bool ok = true;
            if (matModel == 0)
                ok = mat_set_model_linear(m, (float)E, -1.0f, &merr);
            else if (matModel == 1) // (a LINEAR_FAIL material without a positive Fail_Stress never fails)
                ok = mat_set_model_linear(m, (float)E, (float)failStress > 0 ? (float)failStress : -1.0f, &merr);
            else if (matModel == 2)
                ok = mat_set_model_bilinear(m, (float)E, (float)get_double(mech, "Plastic_Mod", 0), (float)get_double(mech, "Yield_Stress", 0), (float)failStress, &merr);
            else if (matModel == 3) { // <SSData><NumDataPts>n</NumDataPts><StrainData><Strain>..</Strain>..</StrainData><StressData><Stress>..
                std::vector<float> strain, stress;
                const XmlNode *ss = mech ? mech->child("SSData") : nullptr;
                const int npts = ss ? get_int(ss, "NumDataPts", 0) : 0;
                if (ss) {
                    if (const XmlNode *sd = ss->child("StrainData"))
                        for (auto &c : sd->children)
                            if (c->name == "Strain" && (int)strain.size() < npts)
                                strain.push_back((float)atof(c->text.c_str()));
                    if (const XmlNode *td = ss->child("StressData"))
                        for (auto &c : td->children)
                            if (c->name == "Stress" && (int)stress.size() < npts)
                                stress.push_back((float)atof(c->text.c_str()));
                }
                if ((int)strain.size() != npts || (int)stress.size() != npts || npts == 0) {
                    ok = false;
                    merr = "SSData needs NumDataPts Strain and Stress values";
                } else
                    ok = mat_set_model_data(m, npts, strain.data(), stress.data(), &merr);
            } else {
                ok = false;
                merr = "unknown MatModel " + std::to_string(matModel);
            }
            if (!ok) {
                // the reference keeps the material's previous (default) model when a setModel* call refuses its arguments
                // (CVX_Material::setModelLinear returns false); a world with such a material would run on E = 1 Pa - refuse instead
                if (err)
                    *err = "material " + std::to_string(m.matid) + ": " + merr;
                return nullptr;
            }
            float nu = (float)get_double(mech, "Poissons_Ratio", 0);
            if (nu < 0)
                nu = 0;
            if (nu >= 0.5)
                nu = 0.5 - FLT_EPSILON * 2;
            m.nuMemory = nu;
            m.rho = (float)get_double(mech, "Density", 0);
            if (m.rho <= 0)
                m.rho = FLT_MIN;
            m.alphaCTE = (float)get_double(mech, "CTE", 0);
            m.muStatic = (float)get_double(mech, "uStatic", 0);
            if (m.muStatic <= 0)
                m.muStatic = 0;
            m.muKinetic = (float)get_double(mech, "uDynamic", 0);
            if (m.muKinetic <= 0)
                m.muKinetic = 0;
            m.zetaInternal = bondZ <= 0 ? 0 : bondZ;
            m.zetaGlobal = slowZ <= 0 ? 0 : slowZ;
            m.zetaCollision = colZ <= 0 ? 0 : colZ;
            m.Cilia = get_double(mech, "Cilia", 0);
            m.fixed = get_int(mech, "Fixed", 0) != 0;
            m.isTarget = get_int(mech, "isTarget", 0) != 0;
            m.isMeasured = get_int(mech, "isMeasured", 1);
            m.sticky = get_int(mech, "Sticky", 0) != 0;
            m.isPaceMaker = get_int(mech, "isPaceMaker", 0) != 0;
            m.PaceMakerPeriod = get_double(mech, "PaceMakerPeriod", 0);
            m.signalValueDecay = get_double(mech, "signalValueDecay", 0.9);
            m.signalTimeDelay = get_double(mech, "signalTimeDelay", 0.03);
            m.inactivePeriod = get_double(mech, "inactivePeriod", 0.03);
            m.removeAfter = get_double(mech, "RemoveFromSimulationAfterThisManySeconds", 0.0);
            m.thermalAfter = get_double(mech, "TurnOnThermalExpansionAfterThisManySeconds", 0.0);
            m.ciliaAfter = get_double(mech, "TurnOnCiliaAfterThisManySeconds", 0.0);
            {
                const XmlNode *disp = pm->child("Display"); // CVXC_Material::ReadXML: Red/Green/Blue/Alpha in 0..1
                auto chan = [&](const char *name, double dflt) {
                    int v = (int)((float)get_double(disp, name, dflt) * 255);
                    return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                };
                m.rgba = chan("Red", 0.5) | (chan("Green", 0.5) << 8) | (chan("Blue", 0.5) << 16) | (chan("Alpha", 1.0) << 24);
            }
            paletteToMat.push_back(M->addMaterial(m));
        }
    }

    const XmlNode *st = vxc->child("Structure");
    if (!st) {
        if (err)
            *err = "no <Structure> element";
        return nullptr;
    }
    const std::string *comp = st->attr("Compression");
    if (!comp || *comp != "ASCII_READABLE") {
        if (err)
            *err = "Structure Compression must be ASCII_READABLE (got " + (comp ? *comp : std::string("none")) + ")";
        return nullptr;
    }
    int X = get_int(st, "X_Voxels", 1), Y = get_int(st, "Y_Voxels", 1), Z = get_int(st, "Z_Voxels", 1);
    auto dataLayers = layers_of(st->child("Data"));
    if ((int)dataLayers.size() < Z) {
        if (err)
            *err = "Structure/Data has fewer <Layer> elements than Z_Voxels";
        return nullptr;
    }
    std::vector<signed char> cell((size_t)X * Y * Z, 0); // (a byte per cell: a 400^3 body is 64 M cells)
    for (int z = 0; z < Z; z++) {
        const std::string &raw = dataLayers[z]->text; // CDATA content, untrimmed like tinyxml's CDATA text
        std::string t = raw.size() == (size_t)X * Y ? raw : xml_trim(raw);
        if (t.size() != (size_t)X * Y) {
            if (err)
                *err = "Voxel layer data not present or does not match expected size (layer " + std::to_string(z) + ")";
            return nullptr;
        }
        for (int k = 0; k < X * Y; k++)
            cell[(size_t)X * Y * z + k] = (signed char)(t[k] - 48);
    }
    // per-voxel CSV layers, indexed over present voxels only (VX_Object.cpp:1851-1959)
    std::vector<double> phase, baseC, shiftC;
    auto read_scalar_layers = [&](const char *tag, std::vector<double> &dst) {
        const XmlNode *n = st->child(tag);
        if (!n)
            return;
        auto ls = layers_of(n);
        for (int z = 0; z < Z && z < (int)ls.size(); z++) {
            auto arr = split_commas(ls[z]->text);
            for (int k = 0; k < X * Y; k++)
                if (cell[(size_t)X * Y * z + k] > 0)
                    dst.push_back((int)arr.size() > k ? atof(arr[k].c_str()) : 0.0);
        }
    };
    auto read_vec3_layers = [&](const char *tag, std::vector<double> &dst) {
        const XmlNode *n = st->child(tag);
        if (!n)
            return;
        auto ls = layers_of(n);
        for (int z = 0; z < Z && z < (int)ls.size(); z++) {
            auto arr = split_commas(ls[z]->text);
            for (int k = 0; k < X * Y; k++)
                if (cell[(size_t)X * Y * z + k] > 0) {
                    if ((int)arr.size() <= k * 3)
                        dst.insert(dst.end(), 3, 0.0);
                    else
                        for (int c = 0; c < 3; c++)
                            dst.push_back((int)arr.size() > 3 * k + c ? atof(arr[3 * k + c].c_str()) : 0.0);
                }
        }
    };
    read_scalar_layers("PhaseOffset", phase);
    read_vec3_layers("BaseCiliaForce", baseC);
    read_vec3_layers("ShiftCiliaForce", shiftC);

    // one slab of the body (include/vx3b200.h section 2b): cut the layers along the axis into nParts ranges holding (nearly) equal
    // numbers of voxels; this model gets range `part` and, as ghosts, the cells of the layer below / above it that a cut link joins
    // to one of its own.  The same rule on every rank, so neighbouring slabs agree on the cut and on the order of the exchange lists.
    const int dims[3] = {X, Y, Z};
    auto cell_at = [&](int x, int y, int z) -> int { return (x < 0 || y < 0 || z < 0 || x >= X || y >= Y || z >= Z) ? 0 : (int)cell[(size_t)X * Y * z + (size_t)X * y + x]; };
    int lo = 0, hi = 0;
    if (cut) {
        if (cut->axis < 0 || cut->axis > 2 || cut->nParts < 1 || cut->part < 0 || cut->part >= cut->nParts) {
            if (err)
                *err = "slab: axis must be 0..2 and 0 <= part < n_parts";
            return nullptr;
        }
        const int n = dims[cut->axis];
        std::vector<int64_t> perLayer((size_t)n, 0);
        int first = n, last = -1;
        for (int z = 0; z < Z; z++)
            for (int y = 0; y < Y; y++)
                for (int x = 0; x < X; x++)
                    if (cell[(size_t)X * Y * z + (size_t)X * y + x] > 0) {
                        const int c = cut->axis == 0 ? x : cut->axis == 1 ? y : z;
                        perLayer[c]++;
                        first = std::min(first, c), last = std::max(last, c);
                    }
        if (last < first || last - first + 1 < cut->nParts) {
            if (err)
                *err = "slab: cannot cut " + std::to_string(std::max(0, last - first + 1)) + " lattice layers into " + std::to_string(cut->nParts) + " slabs";
            return nullptr;
        }
        std::vector<int64_t> cum;
        int64_t run = 0;
        for (int c = first; c <= last; c++)
            cum.push_back(run += perLayer[c]);
        std::vector<int> bounds{first};
        for (int k = 1; k < cut->nParts; k++) { // first layer whose cumulative count reaches k/nParts of the body, then one past it
            const double want = (double)run * k / cut->nParts;
            int idx = (int)(std::lower_bound(cum.begin(), cum.end(), want, [](int64_t a, double b) { return (double)a < b; }) - cum.begin());
            int layer = idx + 1 + first;
            bounds.push_back(std::min(std::max(layer, bounds.back() + 1), last + 1));
        }
        bounds.push_back(last + 1);
        for (size_t k = 1; k < bounds.size(); k++)
            if (bounds[k] <= bounds[k - 1]) {
                if (err)
                    *err = "slab: a slab would be empty";
                return nullptr;
            }
        lo = bounds[cut->part], hi = bounds[cut->part + 1];
        M->slab.part = cut->part, M->slab.nParts = cut->nParts, M->slab.axis = cut->axis, M->slab.layerLo = lo, M->slab.layerHi = hi;
    }
    const int ax = cut ? cut->axis : 2;
    const int step[3] = {ax == 0 ? 1 : 0, ax == 1 ? 1 : 0, ax == 2 ? 1 : 0};

    // voxels in x-fastest, y, z order (VX_Sim.cpp:92-107)
    int64_t counter = 0; // occupied cells of the whole body so far = the voxel's index in the whole body
    for (int z = 0; z < Z; z++)
        for (int y = 0; y < Y; y++)
            for (int x = 0; x < X; x++) {
                int c = cell[(size_t)X * Y * z + (size_t)X * y + x];
                if (c <= 0)
                    continue;
                if (c > (int)paletteToMat.size()) {
                    if (err)
                        *err = "Structure refers to material " + std::to_string(c) + " which is not in the palette";
                    return nullptr;
                }
                bool isGhost = false;
                if (cut) {
                    const int co = ax == 0 ? x : ax == 1 ? y : z;
                    const bool own = co >= lo && co < hi;
                    if (!own) {
                        isGhost = (co == lo - 1 && cell_at(x + step[0], y + step[1], z + step[2]) > 0) ||
                                  (co == hi && cell_at(x - step[0], y - step[1], z - step[2]) > 0);
                        if (!isGhost) {
                            counter++;
                            continue;
                        }
                    }
                }
                int v = M->setVoxel(paletteToMat[c - 1], x, y, z, err, isGhost);
                if (v < 0)
                    return nullptr;
                if ((int64_t)phase.size() > counter)
                    M->phaseOffset[v] = phase[counter];
                if ((int64_t)baseC.size() > counter * 3 + 2)
                    for (int k = 0; k < 3; k++)
                        M->baseCilia[v * 3 + k] = baseC[counter * 3 + k];
                if ((int64_t)shiftC.size() > counter * 3 + 2)
                    for (int k = 0; k < 3; k++)
                        M->shiftCilia[v * 3 + k] = shiftC[counter * 3 + k];
                if (cut) {
                    const int co = ax == 0 ? x : ax == 1 ? y : z;
                    HostModel::Slab &S = M->slab;
                    S.globalIndex.push_back(counter);
                    if (isGhost)
                        (co < lo ? S.ghostLo : S.ghostHi).push_back(v);
                    else {
                        S.numOwn++;
                        // my boundary voxels that the lower / upper neighbour holds as ghosts = own ends of the cut links towards it
                        if (co == lo && cut->part > 0 && cell_at(x - step[0], y - step[1], z - step[2]) > 0)
                            S.sendLo.push_back(v);
                        if (co == hi - 1 && cut->part < cut->nParts - 1 && cell_at(x + step[0], y + step[1], z + step[2]) > 0)
                            S.sendHi.push_back(v);
                    }
                }
                counter++;
            }
    if (cut && M->ghost.size() < (size_t)M->numVoxels())
        M->ghost.resize((size_t)M->numVoxels(), 0);
    if (cut) {
        // whole-body figures every slab needs: voxel / link counts, and the material pairs the body's links join IN THE ORDER the whole
        // body first uses them (CVoxelyze::combinedMaterial, old/Voxelyze.cpp:614-628: a voxel links to its already-present -x, -y, -z
        // neighbours, which are the negative ends), so that link-material numbers and orientations equal the whole body's
        HostModel::Slab &S = M->slab;
        S.bodyVoxels = counter;
        for (int z = 0; z < Z; z++)
            for (int y = 0; y < Y; y++)
                for (int x = 0; x < X; x++) {
                    const int c = cell[(size_t)X * Y * z + (size_t)X * y + x];
                    if (c <= 0)
                        continue;
                    const int nb[3] = {cell_at(x - 1, y, z), cell_at(x, y - 1, z), cell_at(x, y, z - 1)};
                    for (int k = 0; k < 3; k++) {
                        if (nb[k] <= 0)
                            continue;
                        S.bodyLinks++;
                        const int a = paletteToMat[nb[k] - 1], b = paletteToMat[c - 1];
                        bool have = false;
                        for (auto &pr : M->presetLinkMatPairs)
                            have = have || (pr.first == a && pr.second == b) || (pr.first == b && pr.second == a);
                        if (!have)
                            M->presetLinkMatPairs.emplace_back(a, b);
                    }
                }
        if (M->presetLinkMatPairs.empty()) {
            if (err)
                *err = "slab: the body has no links";
            return nullptr;
        }
    }
    if (!M->finalize(err))
        return nullptr;
    if (cut) {
        // recommended time step as host_dt.hpp computes it on a whole-body description (VX3_VoxelyzeKernel.cu:184-217; old/Voxelyze.cpp:286-311)
        HostModel::Slab &S = M->slab;
        const bool vx3 = !M->vx2Compat;
        double maxF2d = 0.0f;
        float maxF2f = 0.0f;
        for (size_t k = 0; k < M->presetLinkMatPairs.size(); k++) {
            const VX3B_LinkMaterial &lm = M->linkMats[k];
            const int a = M->presetLinkMatPairs[k].first, b = M->presetLinkMatPairs[k].second;
            if (lm.nu != 0.0f) {
                if (err)
                    *err = "slab: a body with Poisson's ratio != 0 cannot be cut (its time step depends on every link's rest length)";
                return nullptr;
            }
            if (vx3) {
                const double m1 = M->voxelMats[a].mass, m2 = M->voxelMats[b].mass;
                maxF2d = std::max(maxF2d, (double)(lm.a1 / (m1 < m2 ? m1 : m2)));
            } else {
                const float m1 = M->voxelMats[a].mass, m2 = M->voxelMats[b].mass;
                maxF2f = std::max(maxF2f, lm.a1 / (m1 < m2 ? m1 : m2));
            }
        }
        double opt = vx3 ? 1.0f / (6.283185f * sqrt(maxF2d)) : (double)(float)(1.0f / (6.283185f * sqrtf(maxF2f)));
        if (vx3 && opt < 1e-10)
            opt = 1e-10;
        S.bodyDt = (double)(float)(M->dtFrac * opt);
    }
    return M;
}

} // namespace vx3b

// ------------------------------------------------------------------------------------------
// C ABI: host model
// ------------------------------------------------------------------------------------------

namespace vx3b {
thread_local std::string g_last_error;
void set_error(const std::string &e) { g_last_error = e; }
} // namespace vx3b

struct vx3b_model {
    std::unique_ptr<vx3b::HostModel> m;
};

extern "C" {

const char *vx3b_last_error(void) { return vx3b::g_last_error.c_str(); }

int vx3b_model_from_vxa(const char *vxa_xml, const char *vxd_xml, vx3b_model **out) {
    if (!vxa_xml || !out) {
        vx3b::set_error("vx3b_model_from_vxa: null argument");
        return VX3B_ERR_ARG;
    }
    std::string err;
    auto m = vx3b::model_from_vxa(vxa_xml, vxd_xml, &err);
    if (!m) {
        vx3b::set_error(err);
        return VX3B_ERR_PARSE;
    }
    *out = new vx3b_model{std::move(m)};
    return VX3B_OK;
}

static bool slurp(const char *path, std::string *out) {
    std::ifstream f(path, std::ios::binary);
    if (!f)
        return false;
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

int vx3b_model_from_vxa_slab(const char *vxa_xml, const char *vxd_xml, int axis, int part, int n_parts, vx3b_model **out) {
    if (!vxa_xml || !out) {
        vx3b::set_error("vx3b_model_from_vxa_slab: null argument");
        return VX3B_ERR_ARG;
    }
    std::string err;
    const vx3b::SlabCut cut{axis, part, n_parts};
    auto m = vx3b::model_from_vxa(vxa_xml, vxd_xml, &err, &cut);
    if (!m) {
        vx3b::set_error(err);
        return VX3B_ERR_PARSE;
    }
    *out = new vx3b_model{std::move(m)};
    return VX3B_OK;
}

int vx3b_model_slab_info(const vx3b_model *m, VX3B_SlabInfo *out) {
    if (!m || !out || m->m->slab.nParts == 0) {
        vx3b::set_error("vx3b_model_slab_info: not a slab model (vx3b_model_from_vxa_slab)");
        return VX3B_ERR_ARG;
    }
    const vx3b::HostModel::Slab &S = m->m->slab;
    memset(out, 0, sizeof(*out));
    out->part = S.part, out->n_parts = S.nParts, out->axis = S.axis, out->layer_lo = S.layerLo, out->layer_hi = S.layerHi, out->num_own = S.numOwn;
    out->n_send_lo = (int32_t)S.sendLo.size(), out->n_send_hi = (int32_t)S.sendHi.size();
    out->n_ghost_lo = (int32_t)S.ghostLo.size(), out->n_ghost_hi = (int32_t)S.ghostHi.size();
    out->body_voxels = S.bodyVoxels, out->body_links = S.bodyLinks, out->body_dt = S.bodyDt;
    out->send_lo = S.sendLo.data(), out->send_hi = S.sendHi.data(), out->ghost_lo = S.ghostLo.data(), out->ghost_hi = S.ghostHi.data();
    out->global_index = S.globalIndex.data();
    return VX3B_OK;
}

int vx3b_model_from_files(const char *vxa_path, const char *vxd_path, vx3b_model **out) {
    std::string a, d;
    if (!vxa_path || !slurp(vxa_path, &a)) {
        vx3b::set_error(std::string("cannot read ") + (vxa_path ? vxa_path : "(null)"));
        return VX3B_ERR_IO;
    }
    if (vxd_path && *vxd_path && !slurp(vxd_path, &d)) {
        vx3b::set_error(std::string("cannot read ") + vxd_path);
        return VX3B_ERR_IO;
    }
    return vx3b_model_from_vxa(a.c_str(), vxd_path && *vxd_path ? d.c_str() : nullptr, out);
}

int vx3b_model_new(double vox_size, vx3b_model **out) {
    if (!out || !(vox_size > 0)) {
        vx3b::set_error("vx3b_model_new: bad argument");
        return VX3B_ERR_ARG;
    }
    auto m = std::make_unique<vx3b::HostModel>();
    m->voxSize = vox_size;
    m->enableCollision = false; // CVoxelyze default (collisions off, old/Voxelyze.cpp:347)
    m->vx2Compat = true;        // programmatic models follow the Voxelyze2 library API
    *out = new vx3b_model{std::move(m)};
    return VX3B_OK;
}

int vx3b_model_add_material(vx3b_model *m, double youngs_modulus, double density) {
    if (!m || !(youngs_modulus > 0)) {
        vx3b::set_error("vx3b_model_add_material: Young's modulus must be positive");
        return VX3B_ERR_ARG;
    }
    vx3b::MatInput in; // CVX_Material::clear defaults (old/VX_Material.cpp:71-90)
    in.E = (float)youngs_modulus;
    in.rho = (float)density;
    if (in.rho <= 0)
        in.rho = FLT_MIN;
    in.matid = (int)m->m->mats.size() + 1;
    m->m->finalized = false;
    return m->m->addMaterial(in);
}

int vx3b_model_set_material_param(vx3b_model *m, int mat, const char *name, double value) {
    if (!m || !name || mat < 0 || mat >= (int)m->m->mats.size()) {
        vx3b::set_error("vx3b_model_set_material_param: bad material index");
        return VX3B_ERR_ARG;
    }
    vx3b::MatInput &M = m->m->mats[mat];
    std::string n = name;
    float f = (float)value;
    if (n == "youngs_modulus") { // CVX_Material::setModelLinear(E, current failure stress)
        std::string e;
        if (!vx3b::mat_set_model_linear(M, f, M.sigmaFail, &e)) {
            vx3b::set_error(e);
            return VX3B_ERR_ARG;
        }
    } else if (n == "density")
        M.rho = f <= 0 ? FLT_MIN : f;
    else if (n == "poissons_ratio") {
        if (f < 0)
            f = 0;
        if (f >= 0.5)
            f = 0.5 - FLT_EPSILON * 2;
        M.nuMemory = f;
        m->m->volumeEffects = m->m->volumeEffects || f != 0;
    } else if (n == "cte")
        M.alphaCTE = f;
    else if (n == "static_friction")
        M.muStatic = f <= 0 ? 0 : f;
    else if (n == "kinetic_friction")
        M.muKinetic = f <= 0 ? 0 : f;
    else if (n == "internal_damping")
        M.zetaInternal = f <= 0 ? 0 : f;
    else if (n == "global_damping")
        M.zetaGlobal = f <= 0 ? 0 : f;
    else if (n == "collision_damping")
        M.zetaCollision = f <= 0 ? 0 : f;
    else if (n == "fail_stress") {
        std::string e;
        if (!vx3b::mat_set_model_linear(M, M.E, f > 0 ? f : -1.0f, &e)) {
            vx3b::set_error(e);
            return VX3B_ERR_ARG;
        }
    } else if (n == "cilia")
        M.Cilia = value;
    else if (n == "fixed")
        M.fixed = value != 0;
    else if (n == "is_target")
        M.isTarget = value != 0;
    else if (n == "is_measured")
        M.isMeasured = (int)value;
    else if (n == "matid")
        M.matid = (int)value;
    else {
        vx3b::set_error("unknown material parameter: " + n);
        return VX3B_ERR_ARG;
    }
    m->m->finalized = false;
    return VX3B_OK;
}

// CVX_Material::setModelBilinear / setModel on a builder material (old/VX_Material.cpp:300-449)
int vx3b_model_set_material_bilinear(vx3b_model *m, int mat, double youngs_modulus, double plastic_modulus, double yield_stress, double failure_stress) {
    if (!m || mat < 0 || mat >= (int)m->m->mats.size()) {
        vx3b::set_error("vx3b_model_set_material_bilinear: bad material index");
        return VX3B_ERR_ARG;
    }
    std::string e;
    if (!vx3b::mat_set_model_bilinear(m->m->mats[mat], (float)youngs_modulus, (float)plastic_modulus, (float)yield_stress, (float)failure_stress, &e)) {
        vx3b::set_error(e);
        return VX3B_ERR_ARG;
    }
    m->m->finalized = false;
    return VX3B_OK;
}
int vx3b_model_set_material_data(vx3b_model *m, int mat, int num_points, const double *strain, const double *stress) {
    if (!m || mat < 0 || mat >= (int)m->m->mats.size() || num_points < 1 || !strain || !stress) {
        vx3b::set_error("vx3b_model_set_material_data: bad argument");
        return VX3B_ERR_ARG;
    }
    std::vector<float> sd(strain, strain + num_points), td(stress, stress + num_points);
    std::string e;
    if (!vx3b::mat_set_model_data(m->m->mats[mat], num_points, sd.data(), td.data(), &e)) {
        vx3b::set_error(e);
        return VX3B_ERR_ARG;
    }
    m->m->finalized = false;
    return VX3B_OK;
}

int vx3b_model_set_voxel(vx3b_model *m, int mat, int ix, int iy, int iz) {
    if (!m)
        return VX3B_ERR_ARG;
    std::string err;
    int v = m->m->setVoxel(mat, ix, iy, iz, &err);
    if (v < 0) {
        vx3b::set_error(err);
        return VX3B_ERR_ARG;
    }
    return v;
}

int vx3b_model_set_fixed(vx3b_model *m, int ix, int iy, int iz, int dof_mask, const double *translation3, const double *rotation3) {
    if (!m)
        return VX3B_ERR_ARG;
    int v = m->m->voxelAt(ix, iy, iz);
    if (v < 0) {
        vx3b::set_error("vx3b_model_set_fixed: no voxel at that lattice position");
        return VX3B_ERR_ARG;
    }
    m->m->ensureExt();
    m->m->dofFixed[v] = (uint8_t)(dof_mask & VX3B_DOF_ALL);
    for (int k = 0; k < 3; k++) { // CVX_External::setDisplacement: only fixed dofs keep a displacement
        m->m->extTranslation[v * 3 + k] = (translation3 && (dof_mask & (1 << k))) ? translation3[k] : 0.0;
        m->m->extRotation[v * 3 + k] = (rotation3 && (dof_mask & (8 << k))) ? rotation3[k] : 0.0;
    }
    m->m->anyExt = true;
    return VX3B_OK;
}

int vx3b_model_set_force(vx3b_model *m, int ix, int iy, int iz, const double *force3, const double *moment3) {
    if (!m)
        return VX3B_ERR_ARG;
    int v = m->m->voxelAt(ix, iy, iz);
    if (v < 0) {
        vx3b::set_error("vx3b_model_set_force: no voxel at that lattice position");
        return VX3B_ERR_ARG;
    }
    m->m->ensureExt();
    for (int k = 0; k < 3; k++) {
        if (force3)
            m->m->extForce[v * 3 + k] = (float)force3[k];
        if (moment3)
            m->m->extMoment[v * 3 + k] = (float)moment3[k];
    }
    m->m->anyExt = true;
    return VX3B_OK;
}

int vx3b_model_set_world_param(vx3b_model *m, const char *name, double value) {
    if (!m || !name)
        return VX3B_ERR_ARG;
    vx3b::HostModel &M = *m->m;
    std::string n = name;
    if (n == "gravity") // CVoxelyze::setGravity: multiples of g
        M.gravMult = (float)value;
    else if (n == "grav_acc") // CVX_Sim::SetGravityAccel
        M.gravMult = (float)(-(float)value / 9.80665);
    else if (n == "floor")
        M.floorEnabled = value != 0;
    else if (n == "dt_frac")
        M.dtFrac = value;
    else if (n == "vary_temp")
        M.varyTempEnabled = value != 0;
    else if (n == "temp_amplitude")
        M.tempAmplitude = value;
    else if (n == "temp_period")
        M.tempPeriod = value;
    else if (n == "initial_temp") // CVoxelyze::setAmbientTemperature
        M.initialTemp = (float)value;
    else if (n == "enable_collision")
        M.enableCollision = value != 0;
    else if (n == "enable_cilia")
        M.enableCilia = value != 0;
    else if (n == "enable_expansion")
        M.enableExpansion = value != 0;
    else if (n == "vx2_compat")
        M.vx2Compat = value != 0;
    else {
        vx3b::set_error("unknown world parameter: " + n);
        return VX3B_ERR_ARG;
    }
    return VX3B_OK;
}

int vx3b_model_finalize(vx3b_model *m) {
    if (!m)
        return VX3B_ERR_ARG;
    std::string err;
    if (!m->m->finalize(&err)) {
        vx3b::set_error(err);
        return VX3B_ERR_PARSE;
    }
    return VX3B_OK;
}

int vx3b_model_desc(const vx3b_model *m, VX3B_WorldDesc *out) {
    if (!m || !out)
        return VX3B_ERR_ARG;
    if (!m->m->finalized) {
        vx3b::set_error("vx3b_model_desc: model not finalized");
        return VX3B_ERR_STATE;
    }
    m->m->fillDesc(out);
    return VX3B_OK;
}

int vx3b_model_merged_vxa(const vx3b_model *m, char *buf, int64_t cap, int64_t *needed) {
    if (!m)
        return VX3B_ERR_ARG;
    const std::string &s = m->m->mergedVxa;
    if (needed)
        *needed = (int64_t)s.size() + 1;
    if (buf && cap > 0) {
        size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return VX3B_OK;
}

void vx3b_model_free(vx3b_model *m) { delete m; }

} // extern "C"
