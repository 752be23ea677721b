// Host model: VXA/VXD text or CVoxelyze-style calls -> flat VX3B_WorldDesc.
// Own implementation of what CVX_Sim::ReadVXA + Import (Voxelyze3/src/VXA/VX_Sim.cpp:59-183),
// CVoxelyze::addVoxel/addLink/combinedMaterial (Voxelyze3/src/old/Voxelyze.cpp:422-628) and the
// material classes' updateDerived (old/VX_MaterialVoxel.cpp:58-80, old/VX_MaterialLink.cpp:119-139)
// compute, without their pointer graph: voxels and links come out as index arrays in the
// reference's creation order, materials as constant tables.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "vx3b200.h"
#include "xml_min.hpp"

namespace vx3b {

struct MatInput { // CVX_Material state before derivation (floats, as the reference stores them)
    float E = 1.0f, nu = 0.0f, rho = 1.0f, alphaCTE = 0.0f, muStatic = 0.0f, muKinetic = 0.0f;
    float zetaInternal = 1.0f, zetaGlobal = 0.0f, zetaCollision = 0.0f;
    float sigmaFail = -1.0f, epsilonFail = -1.0f;
    // stress/strain model (CVX_Material::setModelLinear / setModelBilinear / setModel, old/VX_Material.cpp:300-449): the points
    // from (0, 0); linear materials carry the two points the reference gives them
    bool linear = true;
    std::vector<float> strainData{0.0f, 1000000.0f}, stressData{0.0f, 1000000.0f};
    uint32_t rgba = 0xffffffffu; // display colour, (int)(value*255) per channel (VX_Object.h:419-425)
    float nuMemory = 0.0f; // Poisson's ratio as given (muMemory, VX_Sim.cpp:85)
    double Cilia = 0;
    int matid = 0, fixed = 0, isTarget = 0, isMeasured = 1, sticky = 0;
    // signals and secondary-experiment switches (defaults of CVXC_Material::ReadXML, VXA/VX_Object.cpp:1420-1430)
    int isPaceMaker = 0;
    double PaceMakerPeriod = 0, signalValueDecay = 0.9, signalTimeDelay = 0.03, inactivePeriod = 0.03;
    double removeAfter = 0, thermalAfter = 0, ciliaAfter = 0;
};

bool mat_set_model_linear(MatInput &m, float youngsModulus, float failureStress, std::string *err);
bool mat_set_model_bilinear(MatInput &m, float youngsModulus, float plasticModulus, float yieldStress, float failureStress, std::string *err);
bool mat_set_model_data(MatInput &m, int n, const float *strain, const float *stress, std::string *err);

// lattice cell -> voxel index: open addressing over the packed (x, y, z) shorts.  (A std::map of tuples made a 4-million-voxel body
// take 33 s to build and ruled out the 8-million-voxel slabs of a 400^3 body.)
class LatticeIndex {
    std::vector<uint64_t> keys; // 0 = empty
    std::vector<int32_t> vals;
    size_t count = 0, mask = 0;
    static uint64_t pack(int x, int y, int z) { return (uint64_t)(uint16_t)x | ((uint64_t)(uint16_t)y << 16) | ((uint64_t)(uint16_t)z << 32) | (1ull << 48); }
    static size_t mix(uint64_t k) {
        k *= 0x9E3779B97F4A7C15ull;
        return (size_t)(k >> 17);
    }
    void grow() {
        std::vector<uint64_t> ok;
        std::vector<int32_t> ov;
        ok.swap(keys);
        ov.swap(vals);
        const size_t cap = ok.empty() ? 1024 : ok.size() * 2;
        keys.assign(cap, 0);
        vals.assign(cap, -1);
        mask = cap - 1;
        for (size_t i = 0; i < ok.size(); i++)
            if (ok[i]) {
                size_t at = mix(ok[i]) & mask;
                while (keys[at])
                    at = (at + 1) & mask;
                keys[at] = ok[i];
                vals[at] = ov[i];
            }
    }

  public:
    int find(int x, int y, int z) const {
        if (keys.empty())
            return -1;
        const uint64_t k = pack(x, y, z);
        for (size_t at = mix(k) & mask;; at = (at + 1) & mask) {
            if (keys[at] == k)
                return vals[at];
            if (!keys[at])
                return -1;
        }
    }
    void set(int x, int y, int z, int v) {
        if ((count + 1) * 2 > keys.size())
            grow();
        const uint64_t k = pack(x, y, z);
        size_t at = mix(k) & mask;
        while (keys[at] && keys[at] != k)
            at = (at + 1) & mask;
        if (!keys[at])
            count++;
        keys[at] = k;
        vals[at] = v;
    }
    void reserve(size_t n) {
        while (keys.size() < 2 * n + 2)
            grow();
    }
};

struct HostModel {
    // --- inputs -------------------------------------------------------------------------
    double voxSize = 0.001;
    std::vector<MatInput> mats;
    std::vector<int16_t> lattice;   // [nv][3]
    std::vector<int32_t> voxelMat;  // [nv]
    std::vector<double> phaseOffset, baseCilia, shiftCilia;
    std::vector<uint8_t> dofFixed;
    std::vector<float> extForce, extMoment;
    std::vector<double> extTranslation, extRotation;
    bool anyExt = false;
    LatticeIndex occupied;

    float gravMult = 0.0f;
    bool floorEnabled = false;
    double dtFrac = 0.9;
    bool tempEnabled = false, varyTempEnabled = false;
    double tempBase = 25, tempAmplitude = 0, tempPeriod = 0.1;
    float initialTemp = 0.0f;
    bool volumeEffects = false;
    bool enableCollision = true, enableCilia = false, enableExpansion = false, vx2Compat = false;
    bool savePositions = false;
    int recordStepSize = 0, recordVoxel = 1, recordLink = 0; // VXA.Simulator.RecordHistory (VX3_SimulationManager.cu:344-346)
    double maxDistPair = 0;
    std::vector<VX3B_Token> stopCondition, fitness, forceField[3], attachCondition[5];
    bool enableSignals = false, enableAttach = false, enableDetach = false, secondaryExperiment = false;
    int safetyGuard = 500;
    double reinitializeAfter = 0;

    // --- derived (finalize) ---------------------------------------------------------------
    std::vector<int32_t> linkNeg, linkPos, linkMat;
    std::vector<uint8_t> linkAxis;
    std::vector<std::pair<int, int>> linkMatPairs; // combinedMaterial list, in creation order
    std::vector<std::pair<int, int>> presetLinkMatPairs; // slab models: the whole body's list, created first
    std::vector<VX3B_VoxelMaterial> voxelMats;
    std::vector<VX3B_LinkMaterial> linkMats;
    bool finalized = false;
    std::string mergedVxa;

    // one body over several GPUs (include/vx3b200.h section 2b): filled when the model is ONE SLAB of a body (model_from_vxa with a cut)
    std::vector<uint8_t> ghost; // [nv] 1 = copy of a voxel the neighbouring slab owns; empty: not a slab
    struct Slab {
        int part = 0, nParts = 0, axis = 2, layerLo = 0, layerHi = 0, numOwn = 0; // own lattice layers [layerLo, layerHi) along axis
        int64_t bodyVoxels = 0, bodyLinks = 0;                                    // of the whole body
        double bodyDt = 0;                                                        // the time step the whole body takes
        std::vector<int32_t> sendLo, sendHi, ghostLo, ghostHi;                    // exchange lists (slab voxel indices, ascending)
        std::vector<int64_t> globalIndex;                                         // slab voxel -> voxel of the whole body (creation order)
    } slab;

    int numVoxels() const { return (int)voxelMat.size(); }
    int addMaterial(const MatInput &m) {
        mats.push_back(m);
        return (int)mats.size() - 1;
    }
    int setVoxel(int mat, int ix, int iy, int iz, std::string *err, bool isGhost = false);
    int voxelAt(int ix, int iy, int iz) const {
        return occupied.find(ix, iy, iz);
    }
    void ensureExt();
    bool finalize(std::string *err);
    void fillDesc(VX3B_WorldDesc *d) const;
};

// VXD ⊕ VXA merge (Utils/ctool.h:49-57): every child of <VXD> carrying replace="A.B.C" replaces that
// subtree of the base document.
bool merge_vxd(XmlNode &baseDoc, const XmlNode &vxdDoc, std::string *err);
// Reads a (merged) VXA document into a finalized model.  cut != nullptr: only slab `part` of `nParts` of the body along lattice axis
// `axis` - its own layers plus, as ghosts, the neighbouring slabs' voxels that a cut link joins to them - is instantiated.
struct SlabCut {
    int axis, part, nParts;
};
std::unique_ptr<HostModel> model_from_vxa(const std::string &vxaText, const char *vxdText, std::string *err, const SlabCut *cut = nullptr);
// Formula XML -> token array in reversed level order (VX3_SimulationManager.cu:157-275).
bool parse_math_tree(const XmlNode *formulaRoot, std::vector<VX3B_Token> &out, std::string *err);
// Host evaluation of a token array (Utils/VX3_MathTree.h:50-193); empty program -> 0.
double eval_math_tree(const VX3B_Token *tok, double x, double y, double z, double hit, double t, double angle, double closeness,
                      int numClosePairs, int numVoxel);

} // namespace vx3b
