// Precision-independent interface of the step engine; Engine<float> and Engine<double> live in separate
// translation units so the fp64 validation build can be compiled without FMA contraction (-fmad=false),
// matching the reference CPU library's (and the oracle's) unfused double arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vx3b200.h"

namespace vx3b {

class EngineBase {
  public:
    virtual ~EngineBase() {}
    virtual int upload(const VX3B_WorldDesc *worlds, int n) = 0;
    virtual int run(int64_t maxSteps, int useStop) = 0;
    virtual int results(VX3B_WorldResult *out) = 0;
    virtual int read_state(int world, double *pos, double *orient, double *lin, double *ang) = 0;
    virtual int read_initial_pos(int world, double *pos) = 0;
    virtual int read_links(int world, double *pos2, double *a1, double *a2, int32_t *flags) = 0;
    virtual int collision_pairs(int world, int32_t *pairs, int32_t *n) = 0;
    virtual int read_aux(int world, float *voxelTemp, float *linkStrain) = 0;
    virtual int set_force(int world, int voxel, const double *force, const double *moment) = 0;
    virtual int set_global_damping(int world, int mat, double zeta) = 0;
    virtual int run_async(int64_t steps) = 0;
    virtual int halo_set(int world, int listId, const int32_t *voxels, int n) = 0;
    virtual int halo_record_bytes() = 0;
    virtual int halo_export(int world, int listId, void *deviceBuffer) = 0;
    virtual int halo_import(int world, int listId, const void *deviceBuffer) = 0;
    virtual int halo_open(int world, int listId, VX3B_HaloEndpoint *out) = 0;
    virtual int halo_connect(int world, int listId, const VX3B_HaloEndpoint *peer, int sameProcess, const int32_t *peerSlots = nullptr) = 0;
    virtual int halo_slots(int world, int listId, int32_t *slots) = 0;
    virtual int run_slab(int64_t steps) = 0;
    virtual int halo_timeouts(int64_t *n, int64_t *waitNs) = 0;
    // dynamic topology: links in use, signal scalars + removed / attach flags per voxel, current link list
    virtual int num_links(int world, int32_t *n) = 0;
    virtual int read_signals(int world, double *sig6, int32_t *vflags2) = 0;
    virtual int read_topology(int world, int32_t *links6, int32_t *n) = 0;
    int device = 0, precision = 32;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    float lastMs = 0;
    int64_t lastLaunches = 0, lastVoxelSteps = 0;
    // options
    int path = 0;         // 0 auto (resident + brick/fused), 1 two-kernel streaming, 2 resident, 3 fused for all, 4 resident + two-kernel streaming,
                          // 5 brick for every eligible world, fused for the rest (3 and 5 must be set before upload)
    int requestedPath = -1; // the caller's `path` while a batch with non-linear / Poisson worlds forces it to 1 (restored at the next upload)
    int brickDim[3] = {8, 8, 4}; // lattice cells per brick of the fused and brick paths (the brick path halves them until a brick fits)
    int recordPairs = 0;  // keep the collision pair list of the last step (parity tests)
    int chunk = 64;       // steps per CUDA-graph replay on the streaming path
    double dtOverride = 0;
    int residentThreads = 0;    // threads per k_resident CTA; 0 = by size class (512 / CTAs per SM)
    double contactSkin = 0.15;  // skin of the contact candidate lists, in voxel sizes
    int contactListCap = 0;     // entries per candidate list; 0 = from the skin
    int extAll = 0;             // every voxel gets a boundary-condition slot at upload, so set_force() can load any voxel later
    int slabFused = 1;          // one body over several GPUs: exchange fused into the step kernel (1) or staged push / pull kernels (0)
    int contactGrid = 1;        // big worlds: uniform-grid broad phase + candidate lists (1) or the all-pairs scan (0)
};


EngineBase *make_engine_f32();
EngineBase *make_engine_f64();

} // namespace vx3b
