// Engine-internal declarations shared by engine.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "physics.cuh"
#include "vx3b200.h"

namespace vx3b {

void set_error(const std::string &e);

// ---- 4-wide / 2-wide storage vectors per arithmetic type ---------------------------------------
struct alignas(16) D2 {
    double x, y;
};
struct alignas(32) D4 {
    double x, y, z, w;
};
template <class R> struct Store;
template <> struct Store<float> {
    using T4 = float4;
    using T2 = float2;
};
template <> struct Store<double> {
    using T4 = D4;
    using T2 = D2;
};
template <class R> VXD typename Store<R>::T4 mk4(R x, R y, R z, R w) {
    typename Store<R>::T4 v;
    v.x = x;
    v.y = y;
    v.z = z;
    v.w = w;
    return v;
}
template <class R> VXD typename Store<R>::T2 mk2(R x, R y) {
    typename Store<R>::T2 v;
    v.x = x;
    v.y = y;
    return v;
}
VXD uint32_t f2u(float f) { return __float_as_uint(f); }
VXD float u2f(uint32_t u) { return __uint_as_float(u); }

// ---- per-world device record ---------------------------------------------------------------------
enum {
    WF_FLOOR = 1, WF_VARY_TEMP = 2, WF_COLLISION = 4, WF_CILIA = 8, WF_EXPANSION = 16, WF_VX3 = 32, WF_FORCE_FIELD = 64,
    WF_STOP_FORMULA = 128, WF_RESIDENT = 256, // WF_RESIDENT: stepped by k_resident in this run; the streaming kernels skip it
    WF_THERMAL = 512,  // temperature varies AND some material has alphaCTE != 0 (otherwise temperature is unobservable)
    WF_CAN_FAIL = 1024, // some link material has a failure strain
    WF_FUSED = 2048,    // stepped by k_fused in this run (state ping-pongs between the main and the alternate buffers)
    WF_BRICK = 4096,    // stepped by k_brick in this run (state lives in the brick-ordered buffer sets, kernels_brick.cuh)
    WF_NONLIN = 8192    // some link material is bilinear / data based or has a non-zero Poisson's ratio: links use DLinkModel (k_links)
};
struct WorldD {
    // static
    int32_t voff, nv, loff, nl, soff, ns, vmatOff, lmatOff;
    int32_t nvm, nlm; // materials of this world
    int32_t residentClass;       // CTAs per SM of the world's k_resident size class (+16: spill variant); 0 = does not fit
    int32_t parity;              // dynamic: which buffer set holds the world's current state (k_fused ping-pong; 0 outside a run)
    uint32_t flags;
    int32_t cycle; // int(TempPeriod/dt): COM/angle sampling period in steps (0 = never)
    float dt;
    int32_t brickOk; // static: the world has brick tables (eligible for k_brick)
    double tempAmp, tempPeriod, voxSize, maxDistPair;
    // contact broad phase (k_resident): cell edge = largest watch distance of this world + skin
    double invCell, listR2; // 1/edge; edge^2 (a voxel's candidate list holds everything within the edge)
    double watchR2;         // (1 + 1e-12) * square of an UPPER BOUND of every pair's watch distance (exact when voxel sizes cannot change): quick reject
    float skinHalf2;        // (skin/2)^2: a surface voxel that moved further (relative to the common drift) since the last rebuild forces a new one
    int32_t listCap;        // entries per surface voxel's candidate list
    int32_t gridOff, gridLog2; // global-memory contact grid (kernels_contact.cuh): first bucket and log2(buckets); gridLog2 == 0: none
    // stop formula of the common shape `t - C` (every shipped experiment): compared directly instead of run through the formula machine
    int32_t stopIsTimeMinusC, gEpoch; // gEpoch (dynamic): epoch of the global contact grid's bucket heads (kernels_contact.cuh)
    double stopC;
    double tempK; // thermal worlds: 2 * 3.1415926f as the reference evaluates it (VX3_VoxelyzeKernel.cu:640), in double
    // dynamic
    double t;
    int64_t step;     // CurStepCount
    int64_t runSteps; // steps left in the current vx3b_batch_run
    int32_t active, diverged, stopped, divergedNow;
    int32_t angleSamples, numClosePairs, collisionCount, gRebuild; // gRebuild: the global contact lists are stale (set by k_cgrid_check)
    int32_t contactRebuilds, contactOverflows; // k_resident diagnostics: candidate-list rebuilds / rebuilds whose lists overflowed
    double recentAngle, closeness, fitness, totalDistance;
    double comAcc[4]; // Σ m·x, Σ m·y, Σ m·z, Σ m  (measured voxels)
    double distAcc;
    double com[3], initialCom[3], comHist[2][3];
    // centre of mass at the end of the last run (results, fitness).  Kept apart from `com`, which the reference refreshes only on the
    // sampling steps of a run (VX3_VoxelyzeKernel.cu:314-334): a run continued by another call must see what one long run sees
    double finalCom[3];
    int32_t gRefSlot, padG; // brick worlds with contact: brick-order slot of the reference voxel of the stale-list check (kernels_contact.cuh)
    int32_t numMeasured, gRebuildNext; // gRebuildNext: some surface voxel has moved skin/2 since the global contact lists were built: rebuild at the next step
    VX3B_Token stop[VX3B_MAX_TOKENS], fitnessF[VX3B_MAX_TOKENS], ff[3 * VX3B_MAX_TOKENS];
};

struct DExt { // compact table of voxels that carry boundary conditions
    uint32_t dof;
    float force[3], moment[3];
    double translation[3];
    double rotQ[4];
};

// the per-step read+write state: voxel state and link history (one set = one buffer of the k_fused ping-pong)
template <class R> struct StateSet {
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    D2 *pxy, *pzt;
    T4 *orient, *lm;
    T2 *am;
    T4 *lpos2, *la1, *la2;
};

// ---- all device arrays of a batch (struct of arrays; worlds concatenated) ------------------------------
template <class R> struct Arrays {
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    StateSet<R> alt; // second buffer set (allocated only when some world runs on the fused path)
    int nvTot, nlTot, nsTot, nWorlds;
    WorldD *worlds;
    // voxel state (read+write every step)
    D2 *pxy;  // pos.x, pos.y
    D2 *pzt;  // pos.z, {float temp | uint32 flags} packed in the second double's bits
    T4 *orient; // w,x,y,z
    T4 *lm;     // linMom.xyz, angMom.x
    T2 *am;     // angMom.y, angMom.z
    // voxel static
    int32_t *vmat;   // global index into vmats
    int32_t *vworld; // world of each voxel
    int32_t *vlinks; // [6][nvTot] global link index or -1 (direction-major)
    int32_t *vsurf;  // surface slot (global) or -1
    int32_t *vext;   // index into ext or -1
    short4 *lat;     // lattice ix,iy,iz (w unused)
    uint8_t *vghost; // 1 = ghost voxel of a slab (include/vx3b200.h section 2b), or nullptr
    double *phase;   // or nullptr
    D4 *baseCilia;   // xyz (w unused) or nullptr
    D2 *ipxy;        // initial positions (results: total_distance_of_all_voxels)
    double *ipz;
    DExt *ext;
    // surface
    int32_t *surf;   // [nsTot] global voxel index, creation order inside each world
    T4 *contactF;    // [nsTot] contact force accumulated for the voxel pass
    // links
    T4 *lpos2; // pos2.xyz, flags (bit pattern)
    T4 *la1;   // angle1v.xyz, maxStrain
    T4 *la2;   // angle2v.xyz, strain
    int2 *lends;     // global voxel index of (neg, pos)
    uint32_t *lmeta; // axis (bits 0-1) | global link-material index << 2
    int32_t *lworld;
    T4 *lfN, *lfP, *lmN, *lmP; // link forces/moments for the voxel pass (streaming path only)
    // materials
    DVoxMat *vmats;
    DLinkMat *lmats;
    DLinkModel *lmodels; // per link material, or nullptr when no world of the batch is WF_NONLIN
    float *lso;          // per link: strainOffset (VX3_Link.cu:232-268), or nullptr
    // pair recording (parity tests)
    int4 *pairBuf;
    int32_t *pairCount;
    int32_t pairCap;
};

template <class R> VXD StateSet<R> main_set(const Arrays<R> &A) { return StateSet<R>{A.pxy, A.pzt, A.orient, A.lm, A.am, A.lpos2, A.la1, A.la2}; }
// buffer set holding a world's current state
template <class R> VXD StateSet<R> cur_set(const Arrays<R> &A, int parity) { return parity ? A.alt : main_set(A); }
// what the contact kernels read (positions and linear momentum of SURFACE voxels): the world's current buffer set - except for a world
// on the brick path, whose state lives in brick order: its step kernel mirrors the surface voxels into the main arrays every step
template <class R> VXD StateSet<R> contact_set(const Arrays<R> &A, const WorldD &W) { return cur_set(A, (W.flags & WF_BRICK) ? 0 : W.parity); }

} // namespace vx3b
