#pragma once
// Step engine: device memory, upload, run loop, read-back and the engine part of the C ABI.
// Replaces VX3_VoxelyzeKernel's host constructor (VX3/VX3_VoxelyzeKernel.cu:27-105),
// VX3_SimulationManager::startKernel / collectResults (VX3/VX3_SimulationManager.cu:384-458) and, through the
// kernels it launches, CUDA_Simulation + doTimeStep.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "engine.cuh"
#include "engine_base.hpp"
#include "host_dt.hpp"
#include "materials_dev.cuh"
#include "kernels_brick.cuh"
#include "kernels_contact.cuh"
#include "kernels_fused.cuh"
#include "kernels_resident.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

#define CUDA_TRY(expr)                                                                                                          \
    do {                                                                                                                        \
        cudaError_t _e = (expr);                                                                                                \
        if (_e != cudaSuccess) {                                                                                                \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                                                      \
            return VX3B_ERR_CUDA;                                                                                               \
        }                                                                                                                       \
    } while (0)

struct DeviceBuf { // owns one cudaMalloc allocation
    void *p = nullptr;
    size_t bytes = 0;
    ~DeviceBuf() { release(); }
    void release() {
        if (p)
            cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t n) {
        release();
        bytes = n;
        if (n == 0)
            return cudaSuccess;
        return cudaMalloc(&p, n);
    }
};

// Device memory of one context: a few large cudaMalloc chunks handed out in 256-byte aligned pieces and REUSED by the next upload
// (a generation after a generation has the same shape: steady-state uploads allocate nothing; cudaMalloc of ~40 arrays per
// upload cost 24 ms of a 100 ms upload - scripts/gpu_upload_probe.py).
struct DeviceArena {
    std::vector<std::unique_ptr<DeviceBuf>> chunks;
    size_t cur = 0, off = 0;
    void reset() { cur = off = 0; }
    void release() {
        chunks.clear();
        reset();
    }
    cudaError_t alloc(size_t bytes, void **out) {
        bytes = (bytes + 255) & ~(size_t)255;
        for (;;) {
            if (cur < chunks.size() && off + bytes <= chunks[cur]->bytes) {
                *out = (char *)chunks[cur]->p + off;
                off += bytes;
                return cudaSuccess;
            }
            if (cur + 1 < chunks.size()) {
                cur++;
                off = 0;
                continue;
            }
            auto c = std::make_unique<DeviceBuf>();
            cudaError_t e = c->alloc(std::max<size_t>(bytes, (size_t)64 << 20));
            if (e != cudaSuccess)
                return e;
            chunks.push_back(std::move(c));
            cur = chunks.size() - 1;
            off = 0;
        }
    }
};

struct WorldHost {
    int nv, nl, ns, voff, loff, soff, numMeasured;
    double recommendedDt;
    float dt;
};

static bool formula_empty(const VX3B_Token *t) { return t[0].op == VX3B_MT_END; }
// Upper bound of 0.625 * (size1 + size2) over every pair of voxels and every temperature the world can reach
// (VX3_VoxelyzeKernel.cu:660: watchDistance = (baseSizeAverage1 + baseSizeAverage2) * COLLISION_ENVELOPE_RADIUS).  The temperature is
// TempAmplitude * sin(...) in [-|A|, |A|], clamped to <= 0 unless EnableExpansion (VX3_VoxelyzeKernel.cu:638-647), or the initial one.
static double max_watch_distance(const VX3B_WorldDesc &d) {
    double maxSize = 0;
    const double amp = fabs(d.temp_amplitude), t0 = (double)d.initial_temp;
    const bool vary = d.vary_temp_enabled && d.temp_period > 0;
    for (int k = 0; k < d.num_voxel_mats; k++) {
        const double a = (double)d.voxel_mats[k].alphaCTE;
        double grow = std::max(0.0, t0 * a);
        if (vary)
            grow = std::max(grow, d.enable_expansion ? amp * fabs(a) : std::max(0.0, -amp * a));
        maxSize = std::max(maxSize, d.voxel_mats[k].nomSize * (1.0 + grow));
    }
    if (maxSize == 0)
        maxSize = d.vox_size;
    return 1.25 * maxSize * (1.0 + 1e-6);
}

template <class R> class Engine : public EngineBase {
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;

  public:
    Arrays<R> A;
    std::vector<WorldHost> H;
    DeviceArena arena;
    struct Staging { // host staging arrays of upload(), kept between uploads
        std::vector<D2> pxy, pzt, ipxy;
        std::vector<double> ipz, phase;
        std::vector<T4> orient, lm, lpos2;
        std::vector<T2> am;
        std::vector<int32_t> vmat, vworld, vlinks, vsurf, vext, surf, lworld;
        std::vector<short4> lat;
        std::vector<uint8_t> vghost;
        std::vector<D4> baseCilia;
        std::vector<DExt> ext;
        std::vector<int2> lends;
        std::vector<uint32_t> lmeta;
    } staging;
    // the persistent staging arrays are registered as pinned memory (H2D at link speed instead of through the driver's bounce
    // buffers: 7.5 -> ~25 GB/s); re-registered only when a vector had to grow
    struct Pinned {
        void *p = nullptr;
        size_t bytes = 0;
    };
    std::vector<Pinned> pinned;
    template <class V> void pin(size_t slot, V &v) {
        if (pinned.size() <= slot)
            pinned.resize(slot + 1);
        Pinned &r = pinned[slot];
        const size_t bytes = v.capacity() * sizeof(typename V::value_type);
        if (r.p == (void *)v.data() && r.bytes >= v.size() * sizeof(typename V::value_type))
            return;
        if (r.p)
            cudaHostUnregister(r.p);
        r = Pinned{};
        if (bytes >= (1u << 16) && cudaHostRegister((void *)v.data(), bytes, cudaHostRegisterDefault) == cudaSuccess) {
            r.p = (void *)v.data();
            r.bytes = bytes;
        }
        cudaGetLastError(); // registration is an optimisation: a failure is not an error
    }
    void unpin_all() {
        for (auto &r : pinned)
            if (r.p)
                cudaHostUnregister(r.p);
        pinned.clear();
        cudaGetLastError();
    }
    std::vector<ContactTile> hTiles;
    ContactTile *dTiles = nullptr;
    int *dActive = nullptr;
    int *hActive = nullptr; // pinned
    cudaGraphExec_t graphExec = nullptr; // the graph build_graph selected (owned by `graphs`)
    std::map<int, cudaGraphExec_t> graphs;
    int graphUseStop = -1, graphRecord = -1;
    bool anyCollision = false;
    std::vector<ResidentClass> classes; // k_resident size classes (kernels_resident.cuh)
    int numResident = 0;                // worlds that fit in shared memory (laid out first; streaming-only worlds follow)
    int vStream0 = 0, lStream0 = 0, tileStream0 = 0; // first voxel / link / contact tile of the streaming-only worlds
    int graphMixed = -1, graphFused = -1, graphGrid = -1;
    // fused tile path (kernels_fused.cuh): bricks of the streaming-only worlds (of every world when path 3 was chosen before upload)
    FusedArrays FA{};
    int fusedCapTab = 0, fusedCapOwn = 0, fusedBrick0 = 0; // fusedBrick0: first brick of the streaming-only worlds
    bool fusedAllWorlds = false;
    size_t fusedSmem = 0;
    // brick path (kernels_brick.cuh): brick-ordered copy of the eligible streaming-only worlds (of every eligible world for path 5)
    BrickArrays<R> BA{};
    BrickCaps brickCaps{};
    BrickLayout<R> brickLayout;
    size_t brickSmem = 0;
    bool brickLean = true;
    int brickDimsUsed[3] = {0, 0, 0};
    int brickGrid = 0, brickThreads = 256, nBrickWorlds = 0;
    bool brickDoesWorldStep = false;
    bool brickColl = false;     // some brick world has voxel-voxel contact (k_brick's COLL instantiation)
    ContactGrid CG{};           // global-memory contact broad phase of the big worlds (kernels_contact.cuh)
    bool gridCoversAll = false;
    int stepLaunches = 0;        // kernels one enqueue_step launches (counted while enqueueing)
    int sStream0 = 0;           // first surface slot of the streaming-only worlds
    size_t gridBuckets = 0;
    bool brickResident = false; // the truth lives in the brick-ordered sets (after vx3b_batch_run_async); API-order arrays are stale
    std::vector<int32_t> hVoxSlot; // global voxel -> brick-order slot (-1: not a brick world's voxel)
    std::vector<int32_t> linkSlot; // [world's loff + creation index] -> internal (axis-sorted) slot within the world
    struct HaloList {
        int world = -1, n = 0;
        int32_t *dVox = nullptr, *dSlots = nullptr;
    };
    HaloList halo[8];
    struct HaloLink { // device-side exchange: a ghost list owns `area`, a send list maps the neighbour's as `peer`
        void *area = nullptr, *peer = nullptr;
        bool peerIpc = false;
        HaloSync *sync = nullptr;
    };
    HaloLink haloLink[8];
    // exchange fused into k_brick: arrays of this slab the neighbours write into, the neighbours' arrays mapped here, flags
    std::vector<void *> peerExported;   // cudaMalloc'ed brick-set arrays (freed at the next upload)
    std::vector<void *> peerOpened;     // cudaIpcOpenMemHandle results
    BrickPeer<R> peer{};                // kernel parameter
    bool peerOn = false;                // every connected list is connected in fused form
    int peerLists = 0, stagedLists = 0; // connected / opened lists by kind
    std::map<int, cudaGraphExec_t> slabGraphs; // chunk length -> step + exchange graph
    bool slabGraphsStale = false;
    std::vector<int> brickWorldList;
    struct BrickHost {
        std::vector<BrickB> bricks;
        std::vector<int32_t> glist, borig, lorig;
        size_t recomputed = 0;
        std::vector<LinkDesc> ldesc;
        std::vector<uint32_t> bstat;
    };
    cudaEvent_t evBegin = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~Engine() override {
        release_graphs();
        if (hActive)
            cudaFreeHost(hActive);
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
        if (evBegin)
            cudaEventDestroy(evBegin);
        release_classes();
        if (stream)
            cudaStreamSynchronize(stream);
        release_halo_links();
        unpin_all();
        arena.release();
        if (ownStream && stream)
            cudaStreamDestroy(stream);
    }

    void release_classes() {
        for (auto &c : classes) {
            if (c.stream)
                cudaStreamDestroy(c.stream);
            if (c.done)
                cudaEventDestroy(c.done);
        }
        classes.clear();
    }

    template <class T> int dalloc(T **out, size_t count, const T *src = nullptr) {
        void *p = nullptr;
        CUDA_TRY(arena.alloc(std::max<size_t>(count, 1) * sizeof(T), &p));
        *out = (T *)p;
        if (src && count)
            CUDA_TRY(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, stream));
        else
            CUDA_TRY(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), stream));
        return VX3B_OK;
    }

    // per-world summary of upload()'s pre-pass: surface count, feature flags, shared-memory need -> k_resident size class
    struct Pre {
        int ns = 0, cls = 0, next = 0; // surface voxels, k_resident size class, voxels with boundary conditions
        bool full = false;             // needs k_resident's FULL instantiation (boundary conditions, force field, failure strain)
        uint32_t flags = 0;
        size_t bytes = 0;
        double invCell = 0, listR2 = 0;
        float skinHalf2 = 0;
        bool spill = false;
    };

    // ---- which brick a voxel belongs to (both brick builders).  Bricks are cut on one grid from the world's corner, or - when that needs
    // fewer of them - per connected body from each body's own corner: a scene of many bodies (BASELINE configs[3]: 8^3 cubes on a 9-cell
    // pitch) cut on ONE grid fills its bricks to 58 %; cut per body, every brick is full.
    struct BrickCut {
        std::vector<int32_t> body; // [nv] number of the voxel's connected body, in order of the bodies' first voxels
        int nBodies = 0;
        bool perBody = false;
        std::vector<int> mn;       // per body (or one entry for the world): lattice corner
        int64_t nb[3] = {1, 1, 1}; // bricks per axis of the largest box
        int d[3] = {1, 1, 1};
        int64_t cells() const { return (perBody ? (int64_t)nBodies : 1) * nb[0] * nb[1] * nb[2]; }
        int64_t key(short4 c, int v) const {
            const int64_t b = perBody ? body[v] : 0;
            const int *m = &mn[(size_t)b * 3];
            return ((b * nb[2] + (c.z - m[2]) / d[2]) * nb[1] + (c.y - m[1]) / d[1]) * nb[0] + (c.x - m[0]) / d[0];
        }
    };
    void brick_cut_bodies(const WorldHost &h, BrickCut &C) const { // union-find over the links; the root is a body's first voxel
        const auto &lends = staging.lends;
        std::vector<int32_t> &comp = C.body;
        comp.resize(h.nv);
        for (int v = 0; v < h.nv; v++)
            comp[v] = v;
        auto find = [&](int v) {
            while (comp[v] != v)
                v = comp[v] = comp[comp[v]];
            return v;
        };
        for (int l = 0; l < h.nl; l++) {
            const int a = find(lends[h.loff + l].x - h.voff), b = find(lends[h.loff + l].y - h.voff);
            if (a != b)
                comp[std::max(a, b)] = std::min(a, b);
        }
        std::vector<int32_t> rank(h.nv, -1);
        C.nBodies = 0;
        for (int v = 0; v < h.nv; v++) { // (roots come first: a root is its body's first voxel)
            const int r = find(v);
            if (rank[r] < 0)
                rank[r] = C.nBodies++;
            comp[v] = r;
        }
        for (int v = 0; v < h.nv; v++)
            comp[v] = rank[comp[v]];
    }
    // boxes of the bodies (or of the world) in bricks of dims[]; returns the number of brick cells of all boxes
    int64_t brick_cut_boxes(const WorldHost &h, const BrickCut &C, bool perBody, const int dims[3], std::vector<int> &mnOut, int64_t nb[3]) const {
        const auto &lat = staging.lat;
        const int nc = perBody ? C.nBodies : 1;
        mnOut.assign((size_t)nc * 3, 1 << 30);
        std::vector<int> mxs((size_t)nc * 3, -(1 << 30));
        for (int v = 0; v < h.nv; v++) {
            const int c = perBody ? C.body[v] : 0;
            for (int k = 0; k < 3; k++) {
                const int x = (&lat[h.voff + v].x)[k];
                mnOut[(size_t)c * 3 + k] = std::min(mnOut[(size_t)c * 3 + k], x);
                mxs[(size_t)c * 3 + k] = std::max(mxs[(size_t)c * 3 + k], x);
            }
        }
        nb[0] = nb[1] = nb[2] = 1;
        int64_t total = 0;
        for (int c = 0; c < nc; c++) {
            int64_t cells = 1;
            for (int k = 0; k < 3; k++) {
                const int64_t e = (mxs[(size_t)c * 3 + k] - mnOut[(size_t)c * 3 + k]) / dims[k] + 1;
                nb[k] = std::max(nb[k], e);
                cells *= e;
            }
            total += cells;
        }
        return total;
    }
    // the cut of world h for bricks of dims[]: per body when that saves a fifth of the bricks (and the key space stays small)
    void brick_cut_choose(const WorldHost &h, BrickCut &C, const int dims[3]) const {
        std::vector<int> mnW, mnB;
        int64_t nbW[3], nbB[3];
        const int64_t cellsW = brick_cut_boxes(h, C, false, dims, mnW, nbW);
        const int64_t cellsB = C.nBodies > 1 ? brick_cut_boxes(h, C, true, dims, mnB, nbB) : cellsW;
        C.perBody = C.nBodies > 1 && cellsB * 5 < cellsW * 4 && (int64_t)C.nBodies * nbB[0] * nbB[1] * nbB[2] <= (int64_t)8 * h.nv + 4096;
        C.mn = C.perBody ? mnB : mnW;
        for (int k = 0; k < 3; k++) {
            C.nb[k] = C.perBody ? nbB[k] : nbW[k];
            C.d[k] = dims[k];
        }
    }

    // ---- k_brick plan (host only; kernels_brick.cuh): which worlds take the brick path, their voxels and links in brick order, the
    // per-brick descriptor / gather tables, and the brick size that fits the shared memory.  Part of upload(); reads the staging arrays
    // upload() has filled, marks the chosen worlds in hW[].brickOk.
    int build_brick_plan(const VX3B_WorldDesc *worlds, int n, const std::vector<int> &perm, const std::vector<Pre> &pre, std::vector<WorldD> &hW,
                         const std::vector<DVoxMat> &vmats, size_t nvTot, int smemOptin, bool anyGhost, BrickHost &bh) {
        const auto &lat = staging.lat;
        const auto &lends = staging.lends;
        const auto &lmeta = staging.lmeta;
        const auto &vlinks = staging.vlinks;
        const auto &vmat = staging.vmat;
        const auto &vext = staging.vext;
        const auto &vghost = staging.vghost;
        {
            std::vector<int> bw; // brick worlds in layout order
            for (int oi = 0; oi < n; oi++) {
                const int w = perm[oi];
                const VX3B_WorldDesc &d = worlds[w];
                // (automatic path: worlds with voxel-voxel contact stay on k_fused - on the 265-body pile k_brick's COLL instantiation takes
                // 49.8 us per step against k_fused's 46.1 with one 8x8x8 brick per body: 3.6 bricks per CTA leave the ring nothing to overlap,
                // and every surface voxel costs three scattered stores into the API-order arrays; path = 5 asks for k_brick regardless)
                const bool wanted = path == 5 || (path == 0 && pre[w].cls == 0 && !(pre[w].flags & WF_COLLISION));
                if (!wanted || d.num_voxels == 0 || (pre[w].flags & WF_FORCE_FIELD) || (d.ghost && (pre[w].flags & WF_COLLISION)) || d.max_dist_in_voxel_lengths_to_count_as_pair != 0 ||
                    d.num_voxel_mats > BRICK_MAX_VMATS || d.num_link_mats > BRICK_MAX_LMATS)
                    continue;
                bw.push_back(w);
            }
            int dims[3] = {brickDim[0], brickDim[1], brickDim[2]};
            std::map<int, BrickCut> cuts; // per brick world: its connected bodies (found once), the cut for the current dims
            // one persistent CTA (512 threads, two-stage ring) may take the whole SM; small bricks leave room for two (256 threads each)
            const size_t limit = (size_t)smemOptin;
            while (!bw.empty()) {
                bh = BrickHost{};
                BrickCaps caps{0, 0, 0, 0, 1, 1};
                bool ok = true;
                std::vector<int32_t> slotOf, lslotOf, brickOf, local;
                for (int w : bw) {
                    const WorldHost &h = H[w];
                    const WorldD &W = hW[w];
                    caps.nvm = std::max(caps.nvm, W.nvm);
                    caps.nlm = std::max(caps.nlm, W.nlm);
                    BrickCut &cut = cuts[w];
                    if (cut.body.empty())
                        brick_cut_bodies(h, cut);
                    brick_cut_choose(h, cut, dims);
                    std::vector<std::pair<int64_t, int>> keyed(h.nv); // (brick key, voxel): creation order inside a brick = x fastest
                    const int64_t nCells = cut.cells();
                    auto key_of = [&](int v) { return cut.key(lat[h.voff + v], v); };
                    if (nCells <= (int64_t)4 * h.nv + 4096) {
                        // counting sort by brick: stable, so voxels keep their creation order (ascending index) inside a brick, which is
                        // what sorting the (key, index) pairs gives - at a tenth of the time on a million-voxel body
                        std::vector<int32_t> start((size_t)nCells + 1, 0);
                        for (int v = 0; v < h.nv; v++)
                            start[(size_t)key_of(v) + 1]++;
                        for (int64_t c = 0; c < nCells; c++)
                            start[(size_t)c + 1] += start[(size_t)c];
                        for (int v = 0; v < h.nv; v++) {
                            const int64_t k = key_of(v);
                            keyed[(size_t)start[(size_t)k]++] = {k, v};
                        }
                    } else { // a sparse world on a huge lattice: sort the pairs
                        for (int v = 0; v < h.nv; v++)
                            keyed[v] = {key_of(v), v};
                        std::sort(keyed.begin(), keyed.end());
                    }
                    slotOf.assign(h.nv, -1);
                    lslotOf.assign(h.nl, -1);
                    brickOf.assign(h.nv, -1);
                    local.assign(h.nv, -1);
                    const size_t firstBrick = bh.bricks.size();
                    bh.borig.reserve(bh.borig.size() + (size_t)h.nv + (size_t)h.nv / 16 + 64);
                    bh.bstat.reserve(bh.bstat.size() + (size_t)h.nv + (size_t)h.nv / 16 + 64);
                    bh.lorig.reserve(bh.lorig.size() + (size_t)h.nl);
                    bh.ldesc.reserve(bh.ldesc.size() + (size_t)h.nl + (size_t)h.nl / 4 + 64);
                    // pass 1: brick-order slots of voxels and owned links.  Brick extents in `keyed` first, then (host threads side by
                    // side over bricks) the number of links every brick owns, a serial prefix sum for the slot offsets, and the fill.
                    std::vector<size_t> bBeg; // index into keyed of every brick's first voxel (+ end sentinel)
                    for (size_t a = 0; a < keyed.size();) {
                        bBeg.push_back(a);
                        const int64_t key = keyed[a].first;
                        while (a < keyed.size() && keyed[a].first == key)
                            a++;
                    }
                    bBeg.push_back(keyed.size());
                    const size_t nNew = bBeg.size() - 1;
                    const int nT1 = nNew >= 256 ? std::max(1, std::min((int)std::thread::hardware_concurrency(), 16)) : 1;
                    auto over_bricks = [&](auto &&fn) { // fn(first brick, last brick) on nT1 threads
                        if (nT1 == 1) {
                            fn((size_t)0, nNew);
                            return;
                        }
                        std::vector<std::thread> pool;
                        for (int t = 1; t < nT1; t++)
                            pool.emplace_back([&, t] { fn(nNew * (size_t)t / nT1, nNew * (size_t)(t + 1) / nT1); });
                        fn((size_t)0, nNew / nT1);
                        for (auto &th : pool)
                            th.join();
                    };
                    std::vector<int32_t> nOwnedOf(nNew, 0);
                    over_bricks([&](size_t k0, size_t k1) {
                        for (size_t k = k0; k < k1; k++) {
                            int cnt = 0;
                            for (size_t q = bBeg[k]; q < bBeg[k + 1]; q++)
                                for (int axis = 0; axis < 3; axis++)
                                    cnt += vlinks[(size_t)(2 * axis + 1) * nvTot + h.voff + keyed[q].second] >= 0;
                            nOwnedOf[k] = cnt;
                        }
                    });
                    {
                        size_t vAt = bh.borig.size(), lAt = bh.lorig.size();
                        for (size_t k = 0; k < nNew; k++) {
                            vAt = (vAt + 3) & ~(size_t)3; // own runs start on a multiple of four slots: every array, down to the 4-byte
                            BrickB B;                     // static words, is then 16-byte aligned for the bulk copies
                            memset(&B, 0, sizeof(B));
                            B.world = w;
                            B.v0 = (int)vAt;
                            B.nOwn = (int)(bBeg[k + 1] - bBeg[k]);
                            B.l0 = (int)lAt;
                            B.nOwned = nOwnedOf[k];
                            bh.bricks.push_back(B);
                            vAt += (size_t)B.nOwn;
                            lAt += (size_t)B.nOwned;
                        }
                        bh.borig.resize(vAt, -1); // (padding slots keep -1 / 0)
                        bh.bstat.resize(vAt, 0);
                        bh.lorig.resize(lAt);
                    }
                    std::atomic<bool> anyExtVoxel{false};
                    over_bricks([&](size_t k0, size_t k1) {
                        for (size_t k = k0; k < k1; k++) {
                            const BrickB &B = bh.bricks[firstBrick + k];
                            const size_t a = bBeg[k], b = bBeg[k + 1];
                            size_t slot = (size_t)B.v0;
                            for (size_t q = a; q < b; q++, slot++) {
                                const int v = keyed[q].second, g = h.voff + v;
                                slotOf[v] = (int)slot;
                                brickOf[v] = (int)(firstBrick + k);
                                uint32_t mk = 0;
                                for (int dir = 0; dir < 6; dir++)
                                    if (vlinks[(size_t)dir * nvTot + g] >= 0)
                                        mk |= 1u << dir;
                                bh.borig[slot] = g;
                                bh.bstat[slot] = (uint32_t)(vmat[g] - W.vmatOff) | (mk << 8) | (vext[g] >= 0 ? 1u << 14 : 0u) | (vghost[g] ? 1u << 15 : 0u);
                                if (vext[g] >= 0)
                                    anyExtVoxel = true;
                            }
                            size_t ls = (size_t)B.l0;
                            for (int axis = 0; axis < 3; axis++)
                                for (size_t q = a; q < b; q++) {
                                    const int l = vlinks[(size_t)(2 * axis + 1) * nvTot + h.voff + keyed[q].second];
                                    if (l >= 0) {
                                        lslotOf[l - h.loff] = (int)ls;
                                        bh.lorig[ls++] = l;
                                    }
                                }
                        }
                    });
                    if (anyExtVoxel)
                        brickLean = false;
                    // pass 2: link descriptors (recomputed links first, then the owned run in slot order) and halo tables.  Bricks are
                    // independent here, so host threads build them side by side into per-brick pieces (each thread with its own
                    // voxel -> brick-local index scratch); the pieces are appended in brick order afterwards.
                    struct Piece {
                        std::vector<LinkDesc> desc;
                        std::vector<int32_t> glist; // halo voxel slots, then recomputed-link slots
                        int nHalo = 0, nHaloL = 0;
                    };
                    const size_t nb = bh.bricks.size() - firstBrick;
                    std::vector<Piece> pieces(nb);
                    std::vector<size_t> firstVox(nb + 1, 0); // index into keyed of every brick's first own voxel
                    for (size_t k = 0; k < nb; k++)
                        firstVox[k + 1] = firstVox[k] + (size_t)bh.bricks[firstBrick + k].nOwn;
                    auto build = [&](size_t k0, size_t k1, std::vector<int32_t> &loc) {
                        std::vector<int> halo; // world-local voxel index
                        for (size_t k = k0; k < k1; k++) {
                            const size_t bi = firstBrick + k, a = firstVox[k];
                            const BrickB &B = bh.bricks[bi];
                            Piece &P = pieces[k];
                            const int nOwnPad = (B.nOwn + 1) & ~1;
                            halo.clear();
                            for (int i = 0; i < B.nOwn; i++)
                                loc[keyed[a + i].second] = i;
                            auto local_of = [&](int v) {
                                if (loc[v] < 0) {
                                    loc[v] = nOwnPad + (int)halo.size();
                                    halo.push_back(v);
                                }
                                return loc[v];
                            };
                            auto mat_of = [&](int v) { return (uint32_t)(vmat[h.voff + v] - W.vmatOff); };
                            std::vector<int32_t> haloLinks;
                            P.desc.reserve((size_t)B.nOwned + (size_t)B.nOwned / 4 + 8);
                            for (int axis = 0; axis < 3; axis++)
                                for (int i = 0; i < B.nOwn; i++) {
                                    const int v = keyed[a + i].second;
                                    const int l = vlinks[(size_t)(2 * axis) * nvTot + h.voff + v]; // v is the negative end
                                    if (l < 0)
                                        continue;
                                    const int pos = lends[l].y - h.voff;
                                    if (brickOf[pos] == (int)bi)
                                        continue; // owned by this brick
                                    LinkDesc d;
                                    d.a = (uint16_t)i;
                                    d.b = (uint16_t)local_of(pos);
                                    d.lmat = (uint16_t)((lmeta[l] >> 2) - (uint32_t)W.lmatOff);
                                    d.flags = (uint16_t)(axis | BD_NEG_OWN | (mat_of(v) << 4) | (mat_of(pos) << 8));
                                    P.desc.push_back(d);
                                    haloLinks.push_back(lslotOf[l - h.loff]);
                                }
                            P.nHaloL = (int)haloLinks.size();
                            for (int axis = 0; axis < 3; axis++)
                                for (int i = 0; i < B.nOwn; i++) {
                                    const int v = keyed[a + i].second;
                                    const int l = vlinks[(size_t)(2 * axis + 1) * nvTot + h.voff + v]; // v is the positive end: owned
                                    if (l < 0)
                                        continue;
                                    const int neg = lends[l].x - h.voff;
                                    LinkDesc d;
                                    d.a = (uint16_t)local_of(neg);
                                    d.b = (uint16_t)i;
                                    d.lmat = (uint16_t)((lmeta[l] >> 2) - (uint32_t)W.lmatOff);
                                    d.flags = (uint16_t)(axis | BD_POS_OWN | (brickOf[neg] == (int)bi ? BD_NEG_OWN : 0) | (mat_of(neg) << 4) | (mat_of(v) << 8));
                                    P.desc.push_back(d);
                                }
                            P.nHalo = (int)halo.size();
                            P.glist.reserve(halo.size() + haloLinks.size());
                            for (int v : halo) {
                                P.glist.push_back(slotOf[v]);
                                loc[v] = -1;
                            }
                            for (int ls : haloLinks)
                                P.glist.push_back(ls);
                            for (int i = 0; i < B.nOwn; i++)
                                loc[keyed[a + i].second] = -1;
                        }
                    };
                    {
                        const int nT = nb >= 256 ? std::max(1, std::min((int)std::thread::hardware_concurrency(), 16)) : 1;
                        if (nT == 1)
                            build(0, nb, local);
                        else {
                            std::vector<std::thread> pool;
                            for (int t = 0; t < nT; t++)
                                pool.emplace_back([&, t] {
                                    std::vector<int32_t> loc((size_t)h.nv, -1);
                                    build(nb * (size_t)t / nT, nb * (size_t)(t + 1) / nT, loc);
                                });
                            for (auto &th : pool)
                                th.join();
                        }
                    }
                    for (size_t k = 0; k < nb; k++) {
                        BrickB &B = bh.bricks[firstBrick + k];
                        const Piece &P = pieces[k];
                        const int nOwnPad = (B.nOwn + 1) & ~1;
                        if (bh.ldesc.size() & 1)
                            bh.ldesc.push_back(LinkDesc{0, 0, 0, 0});
                        B.descBeg = (int)bh.ldesc.size();
                        bh.ldesc.insert(bh.ldesc.end(), P.desc.begin(), P.desc.end());
                        B.nHaloL = P.nHaloL;
                        bh.recomputed += (size_t)P.nHaloL;
                        B.gBeg = (int)bh.glist.size(); // a multiple of 4 (bulk copies move 16-byte units)
                        B.nHalo = P.nHalo;
                        bh.glist.insert(bh.glist.end(), P.glist.begin(), P.glist.end());
                        while (bh.glist.size() & 3)
                            bh.glist.push_back(0);
                        caps.own = std::max(caps.own, nOwnPad);
                        caps.halo = std::max(caps.halo, B.nHalo);
                        caps.owned = std::max(caps.owned, B.nOwned);
                        caps.haloL = std::max(caps.haloL, B.nHaloL);
                        if (nOwnPad + B.nHalo > 0xffff)
                            ok = false;
                    }
                }
                while (bh.borig.size() & 3) {
                    bh.borig.push_back(-1);
                    bh.bstat.push_back(0);
                }
                if (bh.ldesc.size() & 1)
                    bh.ldesc.push_back(LinkDesc{0, 0, 0, 0});
                brickLayout = BrickLayout<R>(caps);
                brickSmem = brickLayout.total;
                if (ok && brickSmem <= limit) {
                    brickCaps = caps;
                    for (int k = 0; k < 3; k++)
                        brickDimsUsed[k] = dims[k];
                    break;
                }
                int big = dims[1] >= dims[0] && dims[1] >= dims[2] ? 1 : dims[2] >= dims[0] ? 2 : 0; // ties: keep the x runs long
                if (dims[big] <= 1) { // cannot happen for lattice worlds; leave every world to k_fused
                    bw.clear();
                    bh = BrickHost{};
                    break;
                }
                dims[big] = (dims[big] + 1) / 2;
                brickLean = true;
            }
            nBrickWorlds = (int)bw.size();
            brickWorldList = bw;
            hVoxSlot.assign(anyGhost ? nvTot : 0, -1);
            if (anyGhost) {
                for (size_t sl = 0; sl < bh.borig.size(); sl++)
                    if (bh.borig[sl] >= 0)
                        hVoxSlot[bh.borig[sl]] = (int32_t)sl;
                std::vector<char> isBrick(n, 0);
                for (int w : bw)
                    isBrick[w] = 1;
                for (int w = 0; w < n; w++)
                    if (worlds[w].ghost && !isBrick[w]) {
                        set_error("vx3b_batch_upload: world " + std::to_string(w) + " has ghost voxels but is not stepped by the brick path (select option path = 5; no collision / force field / targets)");
                        return VX3B_ERR_ARG;
                    }
            }
            brickColl = false;
            for (int w : bw) {
                if (pre[w].flags & WF_COLLISION) { // voxel-voxel contact on the brick path: where the reference voxel of the stale-list check sits
                    brickColl = true;
                    const int g = hW[w].ns > 0 ? staging.surf[hW[w].soff + (hW[w].ns >> 1)] : -1;
                    hW[w].gRefSlot = 0;
                    for (size_t sl = 0; sl < bh.borig.size(); sl++)
                        if (bh.borig[sl] == g) {
                            hW[w].gRefSlot = (int32_t)sl;
                            break;
                        }
                }
                hW[w].brickOk = 1;
                if (pre[w].flags & (WF_VARY_TEMP | WF_THERMAL | WF_CILIA))
                    brickLean = false;
                for (int k = 0; k < hW[w].nvm; k++)
                    if (vmats[hW[w].vmatOff + k].fixed)
                        brickLean = false;
            }
        }
        return VX3B_OK;
    }

    // ---- k_fused tables (kernels_fused.cuh): bricks of every world the fused path may step, their voxel / link tables, the second
    // buffer set.  Part of upload(); reads the staging arrays upload() has filled.
    int build_fused_tables(int n, const std::vector<int> &perm, const std::vector<Pre> &pre, const std::vector<WorldD> &hW, size_t nvTot, size_t nlTot, int smemOptin) {
        int rc;
        const auto &lat = staging.lat;
        const auto &lends = staging.lends;
        const auto &lmeta = staging.lmeta;
        const auto &vlinks = staging.vlinks;
        {
            std::vector<BrickD> bricks;
            std::vector<int32_t> bvox;
            std::vector<uint8_t> mask;
            std::vector<BrickLink> blinks;
            std::vector<int32_t> local(nvTot, -1);
            fusedCapTab = fusedCapOwn = 0;
            fusedBrick0 = -1;
            const int BX = brickDim[0], BY = brickDim[1], BZ0 = brickDim[2];
            int fusedBZ = BZ0; // (for the debug line: the last world's)
            for (int oi = 0; oi < n; oi++) {
                int BZ = BZ0;
                const int w = perm[oi];
                const bool streamingOnly = pre[w].cls == 0;
                if (!streamingOnly && !fusedAllWorlds)
                    continue;
                if (hW[w].brickOk)
                    continue; // stepped by k_brick
                if (streamingOnly && fusedBrick0 < 0)
                    fusedBrick0 = (int)bricks.size();
                const WorldHost &h = H[w];
                if (h.nv == 0)
                    continue;
                // brick key of every voxel (BrickCut: one grid from the world's corner, or per connected body); voxels of a brick stay in
                // creation order (x runs are contiguous in memory)
                BrickCut cut;
                brick_cut_bodies(h, cut);
                int dims[3] = {BX, BY, BZ};
                if (cut.nBodies > 1 && BX * BY * BZ * 2 <= 512) {
                    // bodies that fit ONE brick of twice the height (the pile's 8^3 cubes in 8x8x8 bricks): no link is cut by a brick face, so
                    // nothing is recomputed and nothing gathered; 2.72e9 -> 2.93e9 voxel-steps/s on the pile
                    std::vector<int> mnTmp;
                    int64_t nbTmp[3];
                    const int d2[3] = {BX, BY, 2 * BZ};
                    if (brick_cut_boxes(h, cut, true, d2, mnTmp, nbTmp) == cut.nBodies && brick_cut_boxes(h, cut, true, dims, mnTmp, nbTmp) > cut.nBodies)
                        dims[2] = BZ = 2 * BZ;
                }
                fusedBZ = BZ;
                brick_cut_choose(h, cut, dims);
                std::vector<std::pair<int64_t, int>> keyed(h.nv);
                auto key_of = [&](int v) { return cut.key(lat[h.voff + v], v); };
                const int64_t nCells = cut.cells();
                if (nCells <= (int64_t)4 * h.nv + 4096) { // stable counting sort by brick (= sorting the (key, index) pairs, much cheaper)
                    std::vector<int32_t> start((size_t)nCells + 1, 0);
                    for (int v = 0; v < h.nv; v++)
                        start[(size_t)key_of(v) + 1]++;
                    for (int64_t c = 0; c < nCells; c++)
                        start[(size_t)c + 1] += start[(size_t)c];
                    for (int v = 0; v < h.nv; v++) {
                        const int64_t k = key_of(v);
                        keyed[(size_t)start[(size_t)k]++] = {k, v};
                    }
                } else {
                    for (int v = 0; v < h.nv; v++)
                        keyed[v] = {key_of(v), v};
                    std::sort(keyed.begin(), keyed.end());
                }
                std::vector<int64_t> brickOf(h.nv);
                for (auto &kv : keyed)
                    brickOf[kv.second] = kv.first;
                for (size_t a = 0; a < keyed.size();) {
                    size_t b = a;
                    while (b < keyed.size() && keyed[b].first == keyed[a].first)
                        b++;
                    BrickD B;
                    memset(&B, 0, sizeof(B));
                    B.world = w;
                    B.voxBeg = (int)bvox.size();
                    B.nOwn = (int)(b - a);
                    B.linkBeg = (int)blinks.size();
                    for (size_t q = a; q < b; q++) {
                        local[h.voff + keyed[q].second] = (int)(q - a);
                        bvox.push_back(h.voff + keyed[q].second);
                    }
                    mask.resize(bvox.size(), 0);
                    std::vector<int> halo;
                    auto local_of = [&](int g) { // brick-local table index of global voxel g (own or halo)
                        if (local[g] < 0) {
                            local[g] = B.nOwn + (int)halo.size();
                            halo.push_back(g);
                        }
                        return local[g];
                    };
                    for (size_t q = a; q < b; q++) {
                        const int g = h.voff + keyed[q].second;
                        uint8_t mk = 0;
                        for (int dir = 0; dir < 6; dir++) {
                            const int l = vlinks[(size_t)dir * nvTot + g];
                            if (l < 0)
                                continue;
                            mk |= (uint8_t)(1u << dir);
                            const int neg = lends[l].x, pos = lends[l].y;
                            const bool negOwn = brickOf[neg - h.voff] == keyed[a].first, posOwn = brickOf[pos - h.voff] == keyed[a].first;
                            // a link is listed once per brick: by its positive end when that end is here (owned), else by its negative end
                            if ((dir & 1) == 0 && posOwn)
                                continue;
                            BrickLink k;
                            memset(&k, 0, sizeof(k));
                            k.link = l;
                            k.a = (uint16_t)local_of(neg);
                            k.b = (uint16_t)local_of(pos);
                            k.flags = (uint16_t)((lmeta[l] & 3) | (posOwn ? BL_OWNED | BL_POS_OWN : 0) | (negOwn ? BL_NEG_OWN : 0));
                            k.lmat = lmeta[l] >> 2;
                            blinks.push_back(k);
                        }
                        mask[B.voxBeg + (q - a)] = mk;
                    }
                    for (int g : halo)
                        bvox.push_back(g);
                    mask.resize(bvox.size(), 0);
                    B.nHalo = (int)halo.size();
                    B.nLinks = (int)blinks.size() - B.linkBeg;
                    for (int q = 0; q < B.nOwn + B.nHalo; q++)
                        local[bvox[B.voxBeg + q]] = -1;
                    fusedCapTab = std::max(fusedCapTab, B.nOwn + B.nHalo);
                    fusedCapOwn = std::max(fusedCapOwn, B.nOwn);
                    bricks.push_back(B);
                    a = b;
                }
            }
            if (fusedBrick0 < 0)
                fusedBrick0 = (int)bricks.size();
            if (!bricks.empty()) {
                if (fusedCapTab > 0xffff) {
                    set_error("vx3b_batch_upload: brick too large for 16-bit local indices");
                    return VX3B_ERR_ARG;
                }
                BrickD *dB = nullptr;
                int32_t *dV = nullptr;
                BrickLink *dL = nullptr;
                uint8_t *dM = nullptr;
                if ((rc = dalloc(&dB, bricks.size(), bricks.data())) != VX3B_OK || (rc = dalloc(&dV, bvox.size(), bvox.data())) != VX3B_OK ||
                    (rc = dalloc(&dL, blinks.size(), blinks.data())) != VX3B_OK || (rc = dalloc(&dM, mask.size(), mask.data())) != VX3B_OK)
                    return rc;
                FA.nBricks = (int)bricks.size();
                FA.bricks = dB;
                FA.bvox = dV;
                FA.blinks = dL;
                FA.linkMask = dM;
                // second buffer set for the ping-pong (contents are produced by the first fused step)
                if ((rc = dalloc(&A.alt.pxy, nvTot)) != VX3B_OK || (rc = dalloc(&A.alt.pzt, nvTot)) != VX3B_OK || (rc = dalloc(&A.alt.orient, nvTot)) != VX3B_OK ||
                    (rc = dalloc(&A.alt.lm, nvTot)) != VX3B_OK || (rc = dalloc(&A.alt.am, nvTot)) != VX3B_OK || (rc = dalloc(&A.alt.lpos2, nlTot)) != VX3B_OK ||
                    (rc = dalloc(&A.alt.la1, nlTot)) != VX3B_OK || (rc = dalloc(&A.alt.la2, nlTot)) != VX3B_OK)
                    return rc;
                fusedSmem = FusedLayout<R>(fusedCapTab, fusedCapOwn).total;
                CUDA_TRY(cudaFuncSetAttribute(k_fused<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin)); // the limit is per function, not per context: never lower it
                if (getenv("VX3B_DEBUG"))
                    fprintf(stderr, "[vx3b] fused path: %zu bricks (%dx%dx%d), table <= %d voxels, own <= %d, %zu links incl. %.1f%% recomputed, %zu B shared memory\n",
                            bricks.size(), BX, BY, fusedBZ, fusedCapTab, fusedCapOwn, blinks.size(), 100.0 * (blinks.size() - (double)nlTot) / std::max<size_t>(nlTot, 1), fusedSmem);
            }
        }
        return VX3B_OK;
    }

    int upload(const VX3B_WorldDesc *worlds, int n) override {
        CUDA_TRY(cudaSetDevice(device));
        release_graphs();
        brickResident = false;
        for (auto &hl : halo)
            hl = HaloList{};
        if (stream)
            CUDA_TRY(cudaStreamSynchronize(stream)); // nothing of the previous batch is still in flight when its memory is reused
        release_halo_links();
        arena.reset();
        H.clear();
        hTiles.clear();
        release_classes();
        memset(&A, 0, sizeof(A));
        if (n <= 0 || !worlds) {
            set_error("vx3b_batch_upload: no worlds");
            return VX3B_ERR_ARG;
        }
        // ---- sizes
        size_t nvTot = 0, nlTot = 0, nvm = 0, nlm = 0;
        for (int w = 0; w < n; w++) {
            const VX3B_WorldDesc &d = worlds[w];
            if (d.num_voxels < 0 || d.num_links < 0 || (d.num_voxels > 0 && (!d.lattice || !d.voxel_mat || !d.voxel_mats)) ||
                (d.num_links > 0 && (!d.link_neg || !d.link_pos || !d.link_axis || !d.link_mat || !d.link_mats))) {
                set_error("vx3b_batch_upload: world " + std::to_string(w) + " has null arrays");
                return VX3B_ERR_ARG;
            }
            nvTot += d.num_voxels;
            nlTot += d.num_links;
            nvm += d.num_voxel_mats;
            nlm += d.num_link_mats;
        }
        if (nvTot > 0x7fffffffull || nlTot > 0x3fffffffull) {
            set_error("vx3b_batch_upload: batch too large for 32-bit indices");
            return VX3B_ERR_ARG;
        }
        const bool dbgT = getenv("VX3B_DEBUG") != nullptr;
        auto tPrev = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (!dbgT)
                return;
            auto now = std::chrono::steady_clock::now();
            fprintf(stderr, "[vx3b] upload %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tPrev).count());
            tPrev = now;
        };
        // ---- per-world summary: surface count, feature flags, shared-memory need -> k_resident size class ------------------
        std::vector<Pre> pre(n);
        std::vector<int> listCaps(n, 16);
        int dev = 0, smemPerSm = 0, smemOptin = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&smemPerSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        CUDA_TRY(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        {
            std::atomic<int> badWorld{-1};
            auto pre_pass = [&](int w, std::vector<uint8_t> &deg) {
                const VX3B_WorldDesc &d = worlds[w];
                Pre &p = pre[w];
                deg.assign(d.num_voxels, 0);
                // material indices index host vectors and device tables, and a link material's two voxel materials are packed into 16 bits
                // each: refuse what is out of range here instead of reading past the tables
                bool badMat = d.num_voxel_mats < 1 || d.num_voxel_mats > 65535 || d.num_link_mats < 0 || (d.num_links > 0 && d.num_link_mats < 1);
                for (int v = 0; v < d.num_voxels && !badMat; v++)
                    badMat = d.voxel_mat[v] < 0 || d.voxel_mat[v] >= d.num_voxel_mats;
                for (int k = 0; k < d.num_link_mats && !badMat; k++)
                    badMat = d.link_mats[k].mat1 < 0 || d.link_mats[k].mat1 >= d.num_voxel_mats || d.link_mats[k].mat2 < 0 || d.link_mats[k].mat2 >= d.num_voxel_mats;
                if (badMat) {
                    badWorld = w;
                    return;
                }
                for (int l = 0; l < d.num_links; l++) {
                    int a = d.link_neg[l], b = d.link_pos[l];
                    if (a < 0 || a >= d.num_voxels || b < 0 || b >= d.num_voxels || d.link_axis[l] > 2 || d.link_mat[l] < 0 || d.link_mat[l] >= d.num_link_mats) {
                        badWorld = w;
                        return;
                    }
                    deg[a]++;
                    deg[b]++;
                }
                for (int v = 0; v < d.num_voxels; v++) {
                    p.ns += deg[v] < 6;
                    bool bc = extAll || (d.dof_fixed && d.dof_fixed[v]);
                    for (int k = 0; k < 3 && !bc; k++)
                        bc = (d.ext_force && d.ext_force[v * 3 + k] != 0) || (d.ext_moment && d.ext_moment[v * 3 + k] != 0);
                    p.next += bc;
                }
                const bool vx3 = !d.vx2_compat;
                const bool hasFF = !formula_empty(d.force_field[0]) || !formula_empty(d.force_field[1]) || !formula_empty(d.force_field[2]);
                double maxAlpha = 0, maxSize = d.vox_size;
                bool canFail = false;
                for (int k = 0; k < d.num_voxel_mats; k++) {
                    maxAlpha = std::max(maxAlpha, (double)fabsf(d.voxel_mats[k].alphaCTE));
                    maxSize = std::max(maxSize, d.voxel_mats[k].nomSize);
                }
                for (int k = 0; k < d.num_link_mats; k++)
                    canFail = canFail || d.link_mats[k].epsilonFail != -1.0f;
                const bool vary = d.vary_temp_enabled && d.temp_period > 0;
                p.flags = (d.floor_enabled ? WF_FLOOR : 0) | (vary ? WF_VARY_TEMP : 0) | (d.enable_collision ? WF_COLLISION : 0) |
                          (d.enable_cilia ? WF_CILIA : 0) | (d.enable_expansion ? WF_EXPANSION : 0) | (vx3 ? WF_VX3 : 0) |
                          (hasFF ? WF_FORCE_FIELD : 0) | (!formula_empty(d.stop_condition) ? WF_STOP_FORMULA : 0) |
                          ((vary && maxAlpha != 0) ? WF_THERMAL : 0) | (canFail ? WF_CAN_FAIL : 0);
                p.full = p.next > 0 || hasFF || canFail;
                // contact broad-phase cell: >= the largest watch distance 0.625*(size1+size2), sizes at their largest temperature
                const double skin = contactSkin * d.vox_size;
                // list capacity: at rest the 12 face diagonals (1.414) enter once skin > 0.16, the 8 body diagonals (1.732) once skin > 0.48
                const int listCap = contactListCap > 0 ? ((contactListCap + 3) & ~3) : contactSkin < 0.16 ? 16 : contactSkin < 0.48 ? 24 : 36; // multiple of 4
                listCaps[w] = listCap;
                double cell = max_watch_distance(d) + skin;
                p.invCell = cell > 0 ? 1.0 / cell : 0.0;
                p.listR2 = cell * cell;
                p.skinHalf2 = (float)(0.25 * skin * skin * (1.0 - 1e-2)); // a little early rather than late (float compare)
                p.bytes = resident_bytes<R>(d.num_voxels, d.num_links, p.ns, d.num_voxel_mats, d.num_link_mats, p.flags, false, listCap);
                if (d.enable_collision && d.lattice && d.num_voxels > 0) {
                    // k_resident's list build recognises lattice neighbours on an index packed 10:10:10 (kernels_resident.cuh, B4)
                    int16_t lo[3] = {32767, 32767, 32767}, hi[3] = {-32768, -32768, -32768};
                    for (int v = 0; v < d.num_voxels; v++)
                        for (int k = 0; k < 3; k++) {
                            lo[k] = std::min(lo[k], d.lattice[3 * v + k]);
                            hi[k] = std::max(hi[k], d.lattice[3 * v + k]);
                        }
                    for (int k = 0; k < 3; k++)
                        if (lo[k] < 0 || hi[k] > 1023) // (the VXA lattice starts at 0; other worlds take the streaming path)
                            p.bytes = (size_t)-1;
                }
                p.cls = p.bytes == (size_t)-1 ? 0 : resident_class_of(p.bytes, (size_t)smemPerSm, (size_t)smemOptin);
                if (p.cls == 0 && p.bytes != (size_t)-1) { // too big: keep the per-link force arrays in global memory instead
                    p.bytes = resident_bytes<R>(d.num_voxels, d.num_links, p.ns, d.num_voxel_mats, d.num_link_mats, p.flags, true, listCap);
                    p.cls = resident_class_of(p.bytes, (size_t)smemPerSm, (size_t)smemOptin);
                    p.spill = p.cls > 0;
                    if (p.spill)
                        p.cls += 16;
                }
            };
            const int nThreads = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, n}));
            std::vector<std::thread> pool;
            auto work = [&](int t) {
                std::vector<uint8_t> deg;
                for (int w = t; w < n; w += nThreads)
                    pre_pass(w, deg);
            };
            for (int t = 1; t < nThreads; t++)
                pool.emplace_back(work, t);
            work(0);
            for (auto &th : pool)
                th.join();
            if (badWorld >= 0) {
                set_error("vx3b_batch_upload: a link of world " + std::to_string((int)badWorld) + " is malformed, or a voxel / link material index of it is out of range");
                return VX3B_ERR_ARG;
            }
        }
        // internal layout order: resident-capable worlds first (largest class = smallest worlds first does not matter; inside a
        // class the most expensive worlds are launched first), streaming-only worlds last so the streaming kernels cover a suffix
        std::vector<int> perm(n);
        for (int w = 0; w < n; w++)
            perm[w] = w;
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) {
            const bool ra = pre[a].cls > 0, rb = pre[b].cls > 0;
            if (ra != rb)
                return ra;
            if (pre[a].cls != pre[b].cls)
                return pre[a].cls < pre[b].cls;
            return worlds[a].num_links + worlds[a].num_voxels > worlds[b].num_links + worlds[b].num_voxels;
        });
        numResident = 0;
        vStream0 = lStream0 = tileStream0 = -1;
        H.resize(n);
        lap("pre-pass");
        // ---- host staging
        // The staging arrays persist in the context: a fresh 150 MB of std::vector per upload costs more in page faults than the
        // staging itself (scripts/gpu_upload_probe.py).  Every element is (re)written below, world by world.
        Staging &G = staging;
        std::vector<WorldD> hW(n);
        auto &pxy = G.pxy, &pzt = G.pzt, &ipxy = G.ipxy;
        auto &ipz = G.ipz, &phase = G.phase;
        auto &orient = G.orient, &lm = G.lm;
        auto &am = G.am;
        auto &vmat = G.vmat, &vworld = G.vworld, &vlinks = G.vlinks, &vsurf = G.vsurf, &vext = G.vext, &surf = G.surf;
        auto &lat = G.lat;
        auto &vghost = G.vghost;
        auto &baseCilia = G.baseCilia;
        auto &ext = G.ext;
        auto &lpos2 = G.lpos2;
        auto &lends = G.lends;
        auto &lmeta = G.lmeta;
        auto &lworld = G.lworld;
        linkSlot.resize(nlTot);
        pxy.resize(nvTot), pzt.resize(nvTot), ipxy.resize(nvTot), ipz.resize(nvTot), phase.resize(nvTot), orient.resize(nvTot), lm.resize(nvTot);
        am.resize(nvTot), vmat.resize(nvTot), vworld.resize(nvTot), vlinks.resize(6 * nvTot), vsurf.resize(nvTot), vext.resize(nvTot);
        lat.resize(nvTot), vghost.resize(nvTot), baseCilia.resize(nvTot), lpos2.resize(nlTot), lends.resize(nlTot), lmeta.resize(nlTot), lworld.resize(nlTot);
        {
            size_t k = 0;
            pin(k++, pxy), pin(k++, pzt), pin(k++, ipxy), pin(k++, ipz), pin(k++, phase), pin(k++, orient), pin(k++, lm), pin(k++, am), pin(k++, vmat);
            pin(k++, vworld), pin(k++, vlinks), pin(k++, vsurf), pin(k++, vext), pin(k++, lat), pin(k++, baseCilia), pin(k++, lpos2), pin(k++, lends);
            pin(k++, lmeta), pin(k++, lworld);
        }
        bool anyGhost = false;
        std::vector<DVoxMat> &vmats = hVmats; // kept: set_global_damping() edits one entry and copies it to the device again
        vmats.assign(nvm, DVoxMat{});
        hDampT.assign(nvm, 0.0f);
        hDampR.assign(nvm, 0.0f);
        std::vector<DLinkMat> lmats(nlm);
        // worlds with bilinear / data materials or a Poisson's ratio need the stress/strain model per link material and a strain offset
        // per link; they are stepped by the two-kernel streaming path, and so is the whole batch they are part of
        bool anyNonlin = false;
        for (int w = 0; w < n && !anyNonlin; w++)
            for (int k = 0; k < worlds[w].num_link_mats && !anyNonlin; k++)
                anyNonlin = !worlds[w].link_mats[k].linear || worlds[w].link_mats[k].nu != 0.0f;
        if (requestedPath >= 0)
            path = requestedPath; // (undo the override of an earlier batch)
        requestedPath = -1;
        if (anyNonlin && path != 1) {
            requestedPath = path;
            path = 1;
        }
        std::vector<DLinkModel> lmodels(anyNonlin ? nlm : 0);
        std::atomic<bool> anyPhaseA{false}, anyCiliaA{false}, anyGhostA{false};
        anyCollision = false;
        // offsets of every world in the concatenated arrays (layout order), then the worlds are staged in parallel
        struct Off {
            size_t v, l, vm, lm, s, e;
        };
        std::vector<Off> off(n);
        {
            size_t voff = 0, loff = 0, vmo = 0, lmo = 0, so = 0, eo = 0;
            int sStreamFirst = -1;
            for (int oi = 0; oi < n; oi++) {
                const int w = perm[oi];
                const VX3B_WorldDesc &d = worlds[w];
                if (pre[w].cls > 0)
                    numResident++;
                else if (vStream0 < 0) {
                    vStream0 = (int)voff;
                    lStream0 = (int)loff;
                    tileStream0 = (int)hTiles.size();
                }
                off[w] = Off{voff, loff, vmo, lmo, so, eo};
                if (pre[w].cls == 0 && sStreamFirst < 0)
                    sStreamFirst = (int)so;
                if (d.enable_collision) {
                    anyCollision = true;
                    for (int f = 0; f < pre[w].ns; f += 128)
                        hTiles.push_back(ContactTile{w, f});
                }
                voff += d.num_voxels;
                loff += d.num_links;
                vmo += d.num_voxel_mats;
                lmo += d.num_link_mats;
                so += pre[w].ns;
                eo += pre[w].next;
            }
            surf.resize(so);
            ext.resize(eo);
            sStream0 = sStreamFirst < 0 ? (int)so : sStreamFirst;
        }
        auto stage_world = [&](int w) {
            const VX3B_WorldDesc &d = worlds[w];
            WorldD &W = hW[w];
            memset(&W, 0, sizeof(W));
            const size_t voff = off[w].v, loff = off[w].l, vmo = off[w].vm, lmo = off[w].lm;
            size_t surfAt = off[w].s, extAt = off[w].e;
            for (int k = 0; k < 6; k++)
                std::fill_n(vlinks.begin() + (size_t)k * nvTot + voff, d.num_voxels, -1);
            std::fill_n(vsurf.begin() + voff, d.num_voxels, -1);
            std::fill_n(vext.begin() + voff, d.num_voxels, -1);
            std::fill_n(vghost.begin() + voff, d.num_voxels, (uint8_t)0);
            WorldHost h;
            h.nv = d.num_voxels;
            h.nl = d.num_links;
            h.voff = (int)voff;
            h.loff = (int)loff;
            h.soff = (int)surfAt;
            const bool vx3 = !d.vx2_compat;
            // dt: VX3_VoxelyzeKernel.cu:244-254 / caller-side DtFrac*recommendedTimeStep for Voxelyze2
            double opt = recommended_dt(d);
            if (vx3 && opt < 1e-10)
                opt = 1e-10;
            h.recommendedDt = opt;
            h.dt = (float)(d.dt_frac * opt);
            if (dtOverride > 0)
                h.dt = (float)dtOverride;
            W.voff = h.voff;
            W.nv = h.nv;
            W.loff = h.loff;
            W.nl = h.nl;
            W.soff = h.soff;
            W.vmatOff = (int)vmo;
            W.lmatOff = (int)lmo;
            W.dt = h.dt;
            W.tempAmp = d.temp_amplitude;
            W.tempPeriod = d.temp_period;
            W.voxSize = d.vox_size;
            W.maxDistPair = d.max_dist_in_voxel_lengths_to_count_as_pair;
            W.cycle = (h.dt > 0 && vx3) ? (int)(d.temp_period / h.dt) : 0;
            W.flags = pre[w].flags;
            W.nvm = d.num_voxel_mats;
            W.nlm = d.num_link_mats;
            W.invCell = pre[w].invCell;
            W.listR2 = pre[w].listR2;
            W.skinHalf2 = pre[w].skinHalf2;
            W.listCap = listCaps[w];
            W.watchR2 = 0; // set below once the materials' constant sizes are known
            W.residentClass = pre[w].cls;
            W.gRebuild = 1; // global contact lists (kernels_contact.cuh): built at the first step, in epoch 1
            W.gEpoch = 1;
            W.gRebuildNext = 0;
            memcpy(W.stop, d.stop_condition, sizeof(W.stop));
            memcpy(W.fitnessF, d.fitness, sizeof(W.fitnessF));
            memcpy(W.ff, d.force_field, sizeof(W.ff));
            // materials
            for (int k = 0; k < d.num_voxel_mats; k++) {
                const VX3B_VoxelMaterial &m = d.voxel_mats[k];
                vmats[vmo + k] = make_dvoxmat(m, (pre[w].flags & WF_THERMAL) != 0, (pre[w].flags & WF_VARY_TEMP) != 0, d.initial_temp);
                hDampT[vmo + k] = m._2xSqMxExS;
                hDampR[vmo + k] = m._2xSqIxExSxSxS;
            }
            { // quick reject of contact candidates: nothing beyond the largest watch distance matters
                double Rw = max_watch_distance(d);
                if (!(pre[w].flags & WF_THERMAL)) { // sizes cannot change: the exact largest value
                    double maxAvg = 0;
                    for (int k = 0; k < d.num_voxel_mats; k++)
                        maxAvg = std::max(maxAvg, vmats[vmo + k].avgSize0);
                    Rw = (maxAvg + maxAvg) * 0.625;
                }
                W.watchR2 = Rw * Rw * (1.0 + 1e-12);
            }
            {
                const VX3B_Token *t = d.stop_condition; // <mtSUB><mtVAR>t</mtVAR><mtCONST>C</mtCONST></mtSUB> = CONST C, VAR t, SUB, END
                const bool tMinusC = t[0].op == VX3B_MT_CONST && t[1].op == VX3B_MT_VAR && t[1].value > 3.5 && t[1].value < 4.5 && t[2].op == VX3B_MT_SUB &&
                                     t[3].op == VX3B_MT_END;
                W.stopIsTimeMinusC = tMinusC ? 1 : 0;
                W.stopC = tMinusC ? t[0].value : 0.0;
                W.tempK = (double)(2 * 3.1415926f);
            }
            for (int k = 0; k < d.num_link_mats; k++) {
                const VX3B_LinkMaterial &m = d.link_mats[k];
                lmats[lmo + k] = make_dlinkmat(m, d.voxel_mats);
                if (anyNonlin) {
                    lmodels[lmo + k] = make_dlinkmodel(m);
                    if (link_material_is_nonlinear(m))
                        W.flags |= WF_NONLIN;
                }
            }
            // voxels
            const bool vary = (W.flags & WF_VARY_TEMP) != 0;
            // A world big enough to matter on its own (one 100^3 body = 1e6 voxels, 3e6 links) is staged by several host threads:
            // every loop below runs over chunks, and the three running cursors (boundary-condition slots, per-axis link slots,
            // surface slots) start each chunk at the prefix sum of the chunks before it, so the layout is the serial one.
            const int nChunks = d.num_voxels >= 200000 ? std::max(1, std::min((int)std::thread::hardware_concurrency(), 16)) : 1;
            auto chunked = [&](int count, auto &&fn) { // fn(chunk, begin, end)
                if (nChunks == 1) {
                    fn(0, 0, count);
                    return;
                }
                std::vector<std::thread> pool;
                for (int c = 1; c < nChunks; c++)
                    pool.emplace_back([&, c] { fn(c, (int)((int64_t)count * c / nChunks), (int)((int64_t)count * (c + 1) / nChunks)); });
                fn(0, 0, (int)((int64_t)count / nChunks));
                for (auto &th : pool)
                    th.join();
            };
            auto has_ext = [&](int v) {
                if (extAll || (d.dof_fixed && d.dof_fixed[v]))
                    return true;
                for (int k = 0; k < 3; k++)
                    if ((d.ext_force && d.ext_force[v * 3 + k] != 0) || (d.ext_moment && d.ext_moment[v * 3 + k] != 0))
                        return true;
                return false;
            };
            std::vector<size_t> extStart(nChunks + 1, extAt);
            if (nChunks > 1) {
                std::vector<size_t> cnt(nChunks, 0);
                chunked(d.num_voxels, [&](int c, int v0, int v1) {
                    size_t k = 0;
                    for (int v = v0; v < v1; v++)
                        k += has_ext(v);
                    cnt[c] = k;
                });
                for (int c = 0; c < nChunks; c++)
                    extStart[c + 1] = extStart[c] + cnt[c];
            }
            chunked(d.num_voxels, [&](int chunk, int vBeg, int vEnd) {
            size_t extAt = extStart[chunk];
            // (flags collected locally: a seq_cst store per voxel to one shared cache line from 16 threads was most of this loop's time)
            bool anyPhaseL = false, anyCiliaL = false, anyGhostL = false;
            for (int v = vBeg; v < vEnd; v++) {
                size_t g = voff + v;
                const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
                short ix = d.lattice[v * 3], iy = d.lattice[v * 3 + 1], iz = d.lattice[v * 3 + 2];
                double px = ix * m.nomSize, py = iy * m.nomSize, pz = iz * m.nomSize;
                float temp = d.initial_temp;
                double ph = d.phase_offset ? d.phase_offset[v] : 0.0;
                if (ph != 0)
                    anyPhaseL = true;
                if (vary && !(vx3 && m.fixed)) { // first step's gpu_update_temperature at t = 0
                    double T = d.temp_amplitude * sin(2 * 3.1415926f * (0.0 / d.temp_period + ph));
                    if (!d.enable_expansion && T > 0)
                        T = 0;
                    temp = (float)T;
                }
                phase[g] = ph;
                pxy[g] = D2{px, py};
                uint32_t tb;
                memcpy(&tb, &temp, 4);
                unsigned long long bits = ((unsigned long long)VF_STATIC_FRICTION << 32) | tb; // voxels start in static friction
                double packed;
                memcpy(&packed, &bits, 8);
                pzt[g] = D2{pz, packed};
                ipxy[g] = D2{px, py};
                ipz[g] = pz;
                orient[g].x = R(1);
                orient[g].y = orient[g].z = orient[g].w = R(0);
                lm[g].x = lm[g].y = lm[g].z = lm[g].w = R(0);
                am[g].x = am[g].y = R(0);
                vmat[g] = (int)vmo + d.voxel_mat[v];
                vworld[g] = w;
                lat[g] = make_short4(ix, iy, iz, 0);
                if (d.ghost && d.ghost[v]) {
                    vghost[g] = 1;
                    anyGhostL = true;
                }
                D4 bc{0, 0, 0, 0};
                if (d.base_cilia) {
                    bc.x = d.base_cilia[v * 3];
                    bc.y = d.base_cilia[v * 3 + 1];
                    bc.z = d.base_cilia[v * 3 + 2];
                    if (d.enable_cilia && (bc.x != 0 || bc.y != 0 || bc.z != 0))
                        anyCiliaL = true;
                }
                baseCilia[g] = bc;
                uint8_t dof = d.dof_fixed ? d.dof_fixed[v] : 0;
                bool hasF = false;
                for (int k = 0; k < 3; k++)
                    hasF = hasF || (d.ext_force && d.ext_force[v * 3 + k] != 0) || (d.ext_moment && d.ext_moment[v * 3 + k] != 0);
                if (dof || hasF || extAll) {
                    DExt e;
                    memset(&e, 0, sizeof(e));
                    e.dof = dof;
                    for (int k = 0; k < 3; k++) {
                        e.force[k] = d.ext_force ? d.ext_force[v * 3 + k] : 0.0f;
                        e.moment[k] = d.ext_moment ? d.ext_moment[v * 3 + k] : 0.0f;
                        e.translation[k] = d.ext_translation ? d.ext_translation[v * 3 + k] : 0.0;
                    }
                    double rx = d.ext_rotation ? d.ext_rotation[v * 3] : 0, ry = d.ext_rotation ? d.ext_rotation[v * 3 + 1] : 0,
                           rz = d.ext_rotation ? d.ext_rotation[v * 3 + 2] : 0;
                    e.rotQ[0] = 1;
                    if (rx != 0 || ry != 0 || rz != 0) { // CVX_External::rotationChanged -> Quat3D(rotation vector)
                        double hx = rx / 2, hy = ry / 2, hz = rz / 2, m2 = hx * hx + hy * hy + hz * hz, qw, s;
                        if (m2 * m2 < 5.328e-15) {
                            qw = 1.0 - 0.5 * m2;
                            s = 1.0 - m2 / 6.0;
                        } else {
                            double mm = sqrt(m2);
                            qw = cos(mm);
                            s = sin(mm) / mm;
                        }
                        e.rotQ[0] = qw;
                        e.rotQ[1] = hx * s;
                        e.rotQ[2] = hy * s;
                        e.rotQ[3] = hz * s;
                    }
                    vext[g] = (int)extAt;
                    ext[extAt++] = e;
                }
            }
            if (anyPhaseL)
                anyPhaseA = true;
            if (anyCiliaL)
                anyCiliaA = true;
            if (anyGhostL)
                anyGhostA = true;
            });
            // links.  Internal order = by axis (x-links, y-links, z-links, each in creation order): a warp of the link pass then
            // works on one axis, and the axis-dependent swizzles stop diverging.  linkSlot maps creation order -> internal slot for
            // the read-backs; nothing else depends on link order (force sums go by direction slot).
            std::vector<int> axisCnt((size_t)nChunks * 3, 0), axisStart((size_t)nChunks * 3, 0);
            chunked(d.num_links, [&](int c, int l0, int l1) {
                int k[3] = {0, 0, 0};
                for (int l = l0; l < l1; l++)
                    k[d.link_axis[l]]++;
                for (int ax = 0; ax < 3; ax++)
                    axisCnt[(size_t)c * 3 + ax] = k[ax];
            });
            {
                int at = 0;
                for (int ax = 0; ax < 3; ax++)
                    for (int c = 0; c < nChunks; c++) {
                        axisStart[(size_t)c * 3 + ax] = at;
                        at += axisCnt[(size_t)c * 3 + ax];
                    }
            }
            chunked(d.num_links, [&](int chunk, int lBeg, int lEnd) {
            int axisAt[3] = {axisStart[(size_t)chunk * 3], axisStart[(size_t)chunk * 3 + 1], axisStart[(size_t)chunk * 3 + 2]};
            for (int l = lBeg; l < lEnd; l++) {
                int a = d.link_neg[l], b = d.link_pos[l], ax = d.link_axis[l];
                const size_t g = loff + (size_t)axisAt[ax]++;
                linkSlot[loff + l] = (int32_t)(g - loff);
                lends[g] = make_int2((int)voff + a, (int)voff + b);
                lmeta[g] = (uint32_t)ax | ((uint32_t)(lmo + d.link_mat[l]) << 2);
                lworld[g] = w;
                vlinks[(size_t)(2 * ax) * nvTot + voff + a] = (int)g;     // positive-side slot of the negative end
                vlinks[(size_t)(2 * ax + 1) * nvTot + voff + b] = (int)g; // negative-side slot of the positive end
                T4 p2;
                p2.x = p2.y = p2.z = R(0);
                uint32_t fl = LF_SMALL_ANGLE; // CVX_Link::reset: smallAngle = true, LOCAL_VELOCITY_VALID = false
                if constexpr (sizeof(R) == 4) {
                    float f;
                    memcpy(&f, &fl, 4);
                    p2.w = f;
                } else {
                    long long b64 = fl;
                    double dd;
                    memcpy(&dd, &b64, 8);
                    p2.w = dd;
                }
                lpos2[g] = p2;
            }
            });
            // surface list: voxels with an empty link slot, creation order (VX3_VoxelyzeKernel.cu:495-513)
            auto is_surface = [&](int v) {
                for (int k = 0; k < 6; k++)
                    if (vlinks[(size_t)k * nvTot + voff + v] < 0)
                        return true;
                return false;
            };
            std::vector<size_t> surfStart(nChunks + 1, surfAt);
            if (nChunks > 1) {
                std::vector<size_t> cnt(nChunks, 0);
                chunked(d.num_voxels, [&](int c, int v0, int v1) {
                    size_t k = 0;
                    for (int v = v0; v < v1; v++)
                        k += is_surface(v);
                    cnt[c] = k;
                });
                for (int c = 0; c < nChunks; c++)
                    surfStart[c + 1] = surfStart[c] + cnt[c];
            }
            chunked(d.num_voxels, [&](int chunk, int v0, int v1) {
                size_t at = surfStart[chunk];
                for (int v = v0; v < v1; v++)
                    if (is_surface(v)) {
                        vsurf[voff + v] = (int)at;
                        surf[at++] = (int)(voff + v);
                    }
                if (chunk == nChunks - 1)
                    surfAt = at;
            });
            h.ns = (int)surfAt - h.soff;
            W.ns = h.ns;
            h.numMeasured = 0;
            for (int v = 0; v < d.num_voxels; v++)
                h.numMeasured += d.voxel_mats[d.voxel_mat[v]].isMeasured != 0;
            H[w] = h;
        };
        {
            const int nThreads = std::max(1, std::min({(int)std::thread::hardware_concurrency(), 16, n}));
            std::vector<std::thread> pool;
            for (int t = 1; t < nThreads; t++)
                pool.emplace_back([&, t] {
                    for (int oi = t; oi < n; oi += nThreads) // layout order is by size: interleaving balances the threads
                        stage_world(perm[oi]);
                });
            for (int oi = 0; oi < n; oi += nThreads)
                stage_world(perm[oi]);
            for (auto &th : pool)
                th.join();
        }
        const bool anyPhase = anyPhaseA, anyCilia = anyCiliaA;
        anyGhost = anyGhostA;
        // global-memory contact grid (kernels_contact.cuh) for the collision worlds the streaming kernels may step: the worlds too
        // big for k_resident, or every world when a streaming path was chosen for all of them
        const bool gridAll = path == 1 || path == 3 || path == 5;
        const int sCover0 = gridAll ? 0 : sStream0;
        {
            size_t buckets = 0;
            for (int oi = 0; oi < n; oi++) {
                const int w = perm[oi];
                if (!contactGrid || !worlds[w].enable_collision || H[w].ns < 2 || !(gridAll || pre[w].cls == 0))
                    continue;
                int lg = 6;
                while ((1 << lg) < 2 * H[w].ns && lg < 22)
                    lg++;
                hW[w].gridOff = (int)buckets;
                hW[w].gridLog2 = lg;
                buckets += (size_t)1 << lg;
            }
            gridBuckets = buckets;
            gridCoversAll = true; // every collision world the streaming kernels may step has a grid: k_contact is not needed
            for (int oi = 0; oi < n; oi++) {
                const int w = perm[oi];
                if (worlds[w].enable_collision && H[w].ns >= 2 && (gridAll || pre[w].cls == 0) && hW[w].gridLog2 == 0)
                    gridCoversAll = false;
            }
        }

        lap("host staging");
        // ---- brick path plan (host only; kernels_brick.cuh): which worlds, brick tables, shared-memory fit
        BrickHost bh;
        BA = BrickArrays<R>{};
        brickLean = true;
        if (int rcPlan = build_brick_plan(worlds, n, perm, pre, hW, vmats, nvTot, smemOptin, anyGhost, bh))
            return rcPlan;
        lap("brick plan");
        // ---- device
        if (!stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            ownStream = true;
        }
        if (!ev0) {
            CUDA_TRY(cudaEventCreate(&ev0));
            CUDA_TRY(cudaEventCreate(&ev1));
        }
        A.nvTot = (int)nvTot;
        A.nlTot = (int)nlTot;
        A.nsTot = (int)surf.size();
        A.nWorlds = n;
        int rc;
#define UP(field, vec)                                                                                                          \
    if ((rc = dalloc(&A.field, (vec).size(), (vec).data())) != VX3B_OK)                                                         \
        return rc;
        UP(worlds, hW);
        UP(pxy, pxy);
        UP(pzt, pzt);
        UP(orient, orient);
        UP(lm, lm);
        UP(am, am);
        UP(vmat, vmat);
        UP(vworld, vworld);
        UP(vlinks, vlinks);
        UP(vsurf, vsurf);
        UP(vext, vext);
        UP(lat, lat);
        if (anyGhost) {
            UP(vghost, vghost);
        }
        UP(ipxy, ipxy);
        UP(ipz, ipz);
        UP(ext, ext);
        UP(surf, surf);
        UP(lpos2, lpos2);
        if ((rc = dalloc(&A.la1, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        if ((rc = dalloc(&A.la2, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        UP(lends, lends);
        UP(lmeta, lmeta);
        UP(lworld, lworld);
        if ((rc = dalloc(&A.lfN, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        if ((rc = dalloc(&A.lfP, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        if ((rc = dalloc(&A.lmN, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        if ((rc = dalloc(&A.lmP, nlTot)) != VX3B_OK) // zero-filled on the device
            return rc;
        UP(vmats, vmats);
        UP(lmats, lmats);
        if (anyNonlin) {
            UP(lmodels, lmodels);
            if ((rc = dalloc(&A.lso, nlTot)) != VX3B_OK) // zero-filled on the device
                return rc;
        }
        if (anyPhase) {
            UP(phase, phase);
        }
        if (anyCilia) {
            UP(baseCilia, baseCilia);
        }
#undef UP
        if ((rc = dalloc(&A.contactF, surf.size())) != VX3B_OK)
            return rc;
        if ((rc = dalloc(&dTiles, hTiles.size(), hTiles.data())) != VX3B_OK)
            return rc;
        CG = ContactGrid{};
        if (gridBuckets > 0) {
            const size_t cover = surf.size() - (size_t)sCover0; // arrays are indexed by absolute surface slot: pointers are pre-offset
            unsigned long long *dHead = nullptr; // zero-filled: epoch 0, which no world ever has
            int32_t *dList = nullptr, *dCount = nullptr;
            int2 *dLink = nullptr;
            float4 *dRef = nullptr;
            short4 *dSlat = nullptr;
            std::vector<short4> slat(cover); // static: lattice index of every covered surface voxel, w = its material is Fixed
            for (size_t k = 0; k < cover; k++) {
                const int g = surf[(size_t)sCover0 + k];
                slat[k] = lat[g];
                slat[k].w = vmats[vmat[g]].fixed ? 1 : 0;
            }
            if ((rc = dalloc(&dHead, gridBuckets)) != VX3B_OK || (rc = dalloc(&dLink, cover)) != VX3B_OK || (rc = dalloc(&dRef, cover)) != VX3B_OK ||
                (rc = dalloc(&dList, cover * CGRID_CAP)) != VX3B_OK || (rc = dalloc(&dCount, cover)) != VX3B_OK || (rc = dalloc(&dSlat, cover, slat.data())) != VX3B_OK)
                return rc;
            CUDA_TRY(cudaStreamSynchronize(stream)); // (the staging vector dies here)
            CG.s0 = sCover0;
            CG.head = dHead;
            CG.link = dLink - sCover0;
            CG.ref = dRef - sCover0;
            CG.slat = dSlat - sCover0;
            CG.list = dList - (size_t)sCover0 * CGRID_CAP;
            CG.count = dCount - sCover0;
            CG.cap = contactListCap > 0 ? std::min((int)CGRID_CAP, contactListCap) : (int)CGRID_CAP;
            CG.nBuckets = (int)std::min<size_t>(gridBuckets, 0x7fffffff);
        }
        if ((rc = dalloc(&dActive, 1)) != VX3B_OK)
            return rc;
        A.pairCap = 1 << 20;
        if ((rc = dalloc(&A.pairBuf, recordPairs ? (size_t)A.pairCap : 1)) != VX3B_OK)
            return rc;
        if ((rc = dalloc(&A.pairCount, 1)) != VX3B_OK)
            return rc;
        if (!hActive)
            CUDA_TRY(cudaMallocHost(&hActive, sizeof(int)));
        lap("device alloc + H2D enqueue");
        // initial centre of mass (CUDA_Simulation: updateCurrentCenterOfMass + InitializeCenterOfMass)
        if (nvTot) {
            k_reduce_com<R><<<(unsigned)((nvTot + 255) / 256), 256, 0, stream>>>(A, 0);
            k_world_finish<R><<<(n + 127) / 128, 128, 0, stream>>>(A, 0);
        }
        CUDA_TRY(cudaGetLastError());
        // ---- fused tile path: cut the streaming-only worlds (all worlds for path 3) into bricks --------------------------------
        FA = FusedArrays{};
        fusedAllWorlds = path == 3 || path == 5;
        if ((rc = build_fused_tables(n, perm, pre, hW, nvTot, nlTot, smemOptin)) != VX3B_OK)
            return rc;
        lap("fused tables");
        // ---- brick path: tables and the two brick-ordered buffer sets
        if (!bh.bricks.empty()) {
            BrickB *dB = nullptr;
            int32_t *dG = nullptr, *dVO = nullptr, *dLO = nullptr;
            LinkDesc *dD = nullptr;
            uint32_t *dS = nullptr;
            if ((rc = dalloc(&dB, bh.bricks.size(), bh.bricks.data())) != VX3B_OK || (rc = dalloc(&dG, bh.glist.size(), bh.glist.data())) != VX3B_OK ||
                (rc = dalloc(&dVO, bh.borig.size(), bh.borig.data())) != VX3B_OK ||
                (rc = dalloc(&dLO, bh.lorig.size(), bh.lorig.data())) != VX3B_OK || (rc = dalloc(&dD, bh.ldesc.size(), bh.ldesc.data())) != VX3B_OK ||
                (rc = dalloc(&dS, bh.bstat.size(), bh.bstat.data())) != VX3B_OK)
                return rc;
            BA.nBricks = (int)bh.bricks.size();
            BA.bricks = dB;
            BA.glist = dG;
            BA.borig = dVO;
            BA.lorig = dLO;
            BA.ldesc = dD;
            BA.bstat = dS;
            if (brickColl) { // brick-order slot -> global surface slot
                std::vector<int32_t> bsurf(bh.borig.size(), -1);
                for (size_t sl = 0; sl < bsurf.size(); sl++)
                    if (bh.borig[sl] >= 0)
                        bsurf[sl] = staging.vsurf[bh.borig[sl]];
                int32_t *dBS = nullptr;
                if ((rc = dalloc(&dBS, bsurf.size(), bsurf.data())) != VX3B_OK)
                    return rc;
                CUDA_TRY(cudaStreamSynchronize(stream)); // (the staging vector dies here)
                BA.bsurf = dBS;
            }
            int32_t *dWL = nullptr;
            int *dTicket = nullptr;
            if ((rc = dalloc(&dWL, brickWorldList.size(), brickWorldList.data())) != VX3B_OK || (rc = dalloc(&dTicket, 1)) != VX3B_OK)
                return rc;
            BA.worldList = dWL;
            BA.nWorlds = (int)brickWorldList.size();
            BA.ticket = dTicket;
            BA.nvSlots = (int)bh.borig.size();
            BA.nlSlots = (int)bh.lorig.size();
            // a slab (ghost voxels): the arrays a neighbouring slab's kernel writes into get allocations of their own, which CUDA IPC can export
            const bool exportable = std::any_of(vghost.begin(), vghost.end(), [](uint8_t g) { return g != 0; });
            auto own_alloc = [&](auto **out, size_t count) -> int {
                if (!exportable)
                    return dalloc(out, count);
                void *q = nullptr;
                const size_t bytes = std::max<size_t>(count, 1) * sizeof(**out);
                CUDA_TRY(cudaMalloc(&q, bytes));
                peerExported.push_back(q);
                CUDA_TRY(cudaMemsetAsync(q, 0, bytes, stream));
                *out = (std::remove_reference_t<decltype(**out)> *)q;
                return VX3B_OK;
            };
            for (int k = 0; k < 2; k++) {
                BrickSet<R> &S = BA.set[k];
                if ((rc = own_alloc(&S.pxy, BA.nvSlots)) != VX3B_OK || (rc = own_alloc(&S.pzt, BA.nvSlots)) != VX3B_OK || (rc = own_alloc(&S.orient, BA.nvSlots)) != VX3B_OK ||
                    (rc = dalloc(&S.lm, BA.nvSlots)) != VX3B_OK || (rc = dalloc(&S.am, BA.nvSlots)) != VX3B_OK || (rc = dalloc(&S.lpos2, BA.nlSlots)) != VX3B_OK ||
                    (rc = dalloc(&S.la1, BA.nlSlots)) != VX3B_OK || (rc = dalloc(&S.la2, BA.nlSlots)) != VX3B_OK)
                    return rc;
            }
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin)); // per function, shared by every context
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
            CUDA_TRY(cudaFuncSetAttribute(k_brick<R, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemOptin));
            // persistent grid: as many CTAs as fit at once, each stepping consecutive bricks in chunks
            int numSm = 0;
            CUDA_TRY(cudaDeviceGetAttribute(&numSm, cudaDevAttrMultiProcessorCount, dev));
            const int perSm = std::max(1, std::min(2, (int)((size_t)smemPerSm / (brickSmem + 1024))));
            brickGrid = std::min(BA.nBricks, perSm * numSm);
            BA.chunk = std::max(1, std::min((int)BRICK_CHUNK, (BA.nBricks + brickGrid - 1) / brickGrid));
            brickThreads = perSm == 1 ? BRICK_MAXT : 256;
            if (getenv("VX3B_DEBUG"))
                fprintf(stderr, "[vx3b] brick path: %zu bricks (%dx%dx%d), own <= %d, halo <= %d, owned links <= %d, recomputed links <= %d (%.1f%%), %zu B shared memory, grid %d x %d threads x chunk %d, %s\n",
                        bh.bricks.size(), brickDimsUsed[0], brickDimsUsed[1], brickDimsUsed[2], brickCaps.own, brickCaps.halo, brickCaps.owned, brickCaps.haloL,
                        100.0 * (double)bh.recomputed / std::max<size_t>(bh.lorig.size(), 1), brickSmem, brickGrid, brickThreads, BA.chunk, brickLean ? "lean" : "full");
        }
        lap("brick tables");
        // ---- k_resident size classes: one launch + stream per class
        if (vStream0 < 0) {
            vStream0 = (int)nvTot;
            lStream0 = (int)nlTot;
            tileStream0 = (int)hTiles.size();
        }
        static const int ks[] = {1 + 16, 2 + 16, 3 + 16, 1, 2, 3, 4, 6}; // most expensive worlds are launched first
        for (int kcf = 0; kcf < 2 * (int)(sizeof(ks) / sizeof(ks[0])); kcf++) {
            const int kc = ks[kcf >> 1];
            ResidentClass c;
            const int k = kc & 15;
            c.ctasPerSm = k;
            c.spill = kc >= 16;
            c.full = (kcf & 1) != 0;
            for (int oi = 0; oi < n; oi++)
                if (pre[perm[oi]].cls == kc && pre[perm[oi]].full == c.full) {
                    c.worlds.push_back(perm[oi]);
                    c.smem = std::max(c.smem, pre[perm[oi]].bytes);
                }
            if (c.worlds.empty())
                continue;
            // registers cap the SM at 512 threads of this kernel (__launch_bounds__(512, 1) -> 128 registers/thread)
            const int maxT = c.full ? res_maxt<R, true>() : res_maxt<R, false>();
            c.threads = residentThreads > 0 ? std::max(32, std::min(maxT, (residentThreads + 31) / 32 * 32)) : std::max(128, (maxT / k) / 32 * 32);
            if ((rc = dalloc(&c.dList, c.worlds.size(), c.worlds.data())) != VX3B_OK)
                return rc;
            CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
            if (getenv("VX3B_DEBUG"))
                fprintf(stderr, "[vx3b] resident class: %d CTA/SM%s%s, %zu worlds, %d threads, %zu B shared memory\n", k, c.spill ? " (spill)" : "",
                        c.full ? " (full)" : "", c.worlds.size(), c.threads, c.smem);
            classes.push_back(std::move(c));
        }
        if (getenv("VX3B_DEBUG"))
            fprintf(stderr, "[vx3b] %d worlds, %d resident, %d streaming-only\n", n, numResident, n - numResident);
        if (!evBegin)
            CUDA_TRY(cudaEventCreateWithFlags(&evBegin, cudaEventDisableTiming));
        lap("classes");
        CUDA_TRY(cudaStreamSynchronize(stream)); // host staging vectors die here
        lap("final sync");
        return VX3B_OK;
    }

    // ---- streaming path: CUDA graph of `chunk` steps -----------------------------------------------------
    // mixed: only the streaming-only worlds (a suffix of every array) are stepped here; the others run in k_resident
    // fused: one k_fused launch per step instead of k_links + k_voxels
    void enqueue_step(int useStop, bool mixed, bool fused) {
        const int l0 = mixed ? lStream0 : 0, v0 = mixed ? vStream0 : 0, t0 = mixed ? tileStream0 : 0;
        const unsigned lb = (A.nlTot - l0 + 255) / 256, vb = (A.nvTot - v0 + 255) / 256, wb = (A.nWorlds + 127) / 128;
        int launches = 0;
        if (!fused && A.nlTot > l0) {
            k_links<R><<<lb, 256, 0, stream>>>(A, l0);
            launches++;
        }
        ContactGrid G{};
        int nsCover = 0;
        if (anyCollision && (int)hTiles.size() > t0) {
            const bool grid = contactGrid && CG.nBuckets > 0;
            if (grid) { // broad phase (insert: only on the steps a world rebuilds) + lists and narrow phase in one kernel
                G = CG;
                G.s0 = std::max(CG.s0, mixed ? sStream0 : 0);
                nsCover = A.nsTot - G.s0;
                if (nsCover > 0) {
                    k_cgrid_insert<R><<<(unsigned)((nsCover + 255) / 256), 256, 0, stream>>>(A, G);
                    k_cgrid_narrow<R><<<(unsigned)((nsCover + 127) / 128), 128, 0, stream>>>(A, G, recordPairs);
                    launches += 2;
                } else
                    G = ContactGrid{};
            }
            if (!grid || !gridCoversAll) { // the all-pairs scan for the worlds without a grid
                k_contact<R><<<(unsigned)hTiles.size() - t0, 128, 0, stream>>>(A, dTiles + t0, recordPairs, grid ? 1 : 0);
                launches++;
            }
        }
        if (fused) {
            const int b0 = mixed ? fusedBrick0 : 0;
            FusedArrays F = FA;
            F.bricks += b0;
            // every world this step touches is a brick world: k_brick's last CTA does the end-of-step bookkeeping itself
            brickDoesWorldStep = BA.nBricks > 0 && FA.nBricks <= b0 && nBrickWorlds == (mixed ? (int)H.size() - numResident : (int)H.size());
            if (FA.nBricks > b0) {
                k_fused<R><<<(unsigned)(FA.nBricks - b0), 256, fusedSmem, stream>>>(A, F, fusedCapTab, fusedCapOwn, G);
                launches++;
            }
            if (BA.nBricks > 0) {
                const int ws = brickDoesWorldStep ? 1 : 0;
                if (peerOn && ws) { // one body over several GPUs: the exchange with the neighbouring slabs is part of the kernel
                    if (brickLean)
                        k_brick<R, true, true><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive, peer);
                    else
                        k_brick<R, false, true><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive, peer);
                } else if (brickColl) { // voxel-voxel contact: contact force in, surface voxels mirrored out, stale-list check of the grid G
                    if (brickLean)
                        k_brick<R, true, false, true><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive, BrickPeer<R>{}, G);
                    else
                        k_brick<R, false, false, true><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive, BrickPeer<R>{}, G);
                } else if (brickLean)
                    k_brick<R, true><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive);
                else
                    k_brick<R, false><<<(unsigned)brickGrid, brickThreads, brickSmem, stream>>>(A, BA, brickLayout, ws, useStop, dActive);
                launches++;
            }
            if (brickDoesWorldStep) {
                stepLaunches = launches;
                return;
            }
        } else if (A.nvTot > v0) {
            k_voxels<R><<<vb, 256, 0, stream>>>(A, v0);
            launches++;
            if (nsCover > 0) { // stale candidate lists?  (the fused path asks inside its voxel pass)
                k_cgrid_check<R><<<(unsigned)((nsCover + 255) / 256), 256, 0, stream>>>(A, G);
                launches++;
            }
        }
        k_world_step<R><<<wb, 128, 0, stream>>>(A, useStop, dActive);
        stepLaunches = launches + 1;
    }
    // one instantiated graph per chunk length (a run of 200 steps replays the 64-step graph three times and an 8-step one once;
    // re-capturing on every change of length costs milliseconds)
    int build_graph(int steps, int useStop, bool mixed, bool fused) {
        if (graphUseStop != useStop || graphRecord != recordPairs || graphMixed != (int)mixed || graphFused != (int)fused || graphGrid != contactGrid)
            release_graphs();
        auto it = graphs.find(steps);
        if (it != graphs.end()) {
            graphExec = it->second;
            return VX3B_OK;
        }
        cudaGraph_t g;
        CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        for (int s = 0; s < steps; s++) {
            if (s == steps - 1)
                k_zero_int<<<1, 1, 0, stream>>>(dActive);
            enqueue_step(useStop, mixed, fused);
        }
        CUDA_TRY(cudaStreamEndCapture(stream, &g));
        CUDA_TRY(cudaGraphInstantiate(&graphExec, g, 0));
        cudaGraphDestroy(g);
        graphs[steps] = graphExec;
        graphUseStop = useStop;
        graphRecord = recordPairs;
        graphMixed = (int)mixed;
        graphFused = (int)fused;
        graphGrid = contactGrid;
        return VX3B_OK;
    }
    void release_graphs() {
        for (auto &kv : graphs)
            cudaGraphExecDestroy(kv.second);
        graphs.clear();
        graphExec = nullptr;
    }

    int run_streaming(int64_t maxSteps, int useStop, bool mixed, bool fused) {
        int64_t done = 0;
        while (done < maxSteps) {
            int steps = (int)std::min<int64_t>(chunk, maxSteps - done);
            int rc = build_graph(steps, useStop, mixed, fused);
            if (rc != VX3B_OK)
                return rc;
            CUDA_TRY(cudaGraphLaunch(graphExec, stream));
            lastLaunches += (int64_t)steps * stepLaunches + 1;
            done += steps;
            CUDA_TRY(cudaMemcpyAsync(hActive, dActive, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (*hActive == 0)
                break;
        }
        return VX3B_OK;
    }

    // every size class on its own stream, all of them ordered after k_world_begin and before the final reductions
    int run_resident(int64_t maxSteps, int useStop) {
        CUDA_TRY(cudaEventRecord(evBegin, stream));
        // (classes are launched as built: the most expensive worlds first.  Reversed, or with the several-per-SM classes first, the run
        // time of a 1,024-world shard moves by +-5 % either way depending on the shard - 85.1 / 92.7 / 87.8 / 90.1 ms as built, 85.6 /
        // 88.5 / 92.4 / 88.0 reversed on four shards - : how the block scheduler packs CTAs of different shared-memory size decides
        // the tail, not the order)
        for (auto &c : classes) {
            CUDA_TRY(cudaStreamWaitEvent(c.stream, evBegin, 0));
            int rc = resident_launch<R>(A, c, maxSteps, useStop, recordPairs);
            if (rc != VX3B_OK)
                return rc;
            CUDA_TRY(cudaEventRecord(c.done, c.stream));
            lastLaunches += 1;
        }
        return VX3B_OK;
    }
    int join_resident() {
        for (auto &c : classes)
            CUDA_TRY(cudaStreamWaitEvent(stream, c.done, 0));
        return VX3B_OK;
    }

    int run(int64_t maxSteps, int useStop) override {
        CUDA_TRY(cudaSetDevice(device));
        if (H.empty()) {
            set_error("vx3b_batch_run: nothing uploaded");
            return VX3B_ERR_STATE;
        }
        lastLaunches = 0;
        std::vector<int64_t> before(H.size());
        int rc = read_steps(before);
        if (rc != VX3B_OK)
            return rc;
        const unsigned wb = (A.nWorlds + 127) / 128;
        CUDA_TRY(cudaEventRecord(ev0, stream));
        // path 0: k_resident for the worlds that fit in shared memory, k_fused (concurrently, on the main stream) for the rest;
        // path 1: the two-kernel streaming path (k_links + k_voxels) for every world; path 2: k_resident for every world (error
        // if one does not fit); path 3: k_fused for every world (must be chosen before upload); path 4: as 0 but with the
        // two-kernel streaming path instead of k_fused
        if (path == 2 && numResident != (int)H.size()) {
            set_error("vx3b_batch_run: resident path requested but " + std::to_string(H.size() - numResident) + " world(s) do not fit in shared memory");
            return VX3B_ERR_ARG;
        }
        if ((path == 3 || path == 5) && !fusedAllWorlds) {
            set_error("vx3b_batch_run: the fused / brick path for every world must be selected (option path = 3 or 5) before vx3b_batch_upload");
            return VX3B_ERR_ARG;
        }
        if (recordPairs) // the pair list is that of the last step of a fixed-length run (W.runSteps == 1)
            CUDA_TRY(cudaMemsetAsync(A.pairCount, 0, sizeof(int), stream));
        const bool resident = (path == 0 || path == 2 || path == 4) && numResident > 0;
        const bool streaming = path == 1 || path == 3 || path == 5 || numResident != (int)H.size();
        const bool fused = streaming && (path == 0 || path == 3 || path == 5) && (FA.nBricks > 0 || BA.nBricks > 0);
        k_world_begin<R><<<wb, 128, 0, stream>>>(A, (long long)maxSteps, useStop, resident ? 1 : 0, fused ? 1 : 0);
        lastLaunches++;
        const unsigned nbb = (unsigned)((std::max(BA.nvSlots, BA.nlSlots) + 255) / 256);
        if (brickResident && !(fused && BA.nBricks > 0) && (rc = ensure_api_state()) != VX3B_OK)
            return rc;
        if (fused && BA.nBricks > 0 && !brickResident) { // API-order state -> brick order
            k_brick_pack<R><<<nbb, 256, 0, stream>>>(A, BA);
            lastLaunches++;
        }
        if (resident && (rc = run_resident(maxSteps, useStop)) != VX3B_OK)
            return rc;
        if (streaming && (rc = run_streaming(maxSteps, useStop, resident, fused)) != VX3B_OK)
            return rc;
        if (resident && (rc = join_resident()) != VX3B_OK)
            return rc;
        if (fused) { // bring every world's state back to the main buffers
            const unsigned nb = (unsigned)((std::max(A.nvTot, A.nlTot) + 255) / 256);
            if (FA.nBricks > 0) {
                k_unparity<R><<<nb, 256, 0, stream>>>(A, 0, 0);
                lastLaunches++;
            }
            if (BA.nBricks > 0) {
                k_brick_unpack<R><<<nbb, 256, 0, stream>>>(A, BA);
                lastLaunches++;
            }
            k_clear_parity<R><<<wb, 128, 0, stream>>>(A);
            lastLaunches++;
            brickResident = false;
        }
        if (A.nvTot) {
            k_reduce_com<R><<<(unsigned)((A.nvTot + 255) / 256), 256, 0, stream>>>(A, 1);
            k_world_finish<R><<<wb, 128, 0, stream>>>(A, 1);
            lastLaunches += 2;
        }
        CUDA_TRY(cudaEventRecord(ev1, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventElapsedTime(&lastMs, ev0, ev1));
        std::vector<int64_t> after(H.size());
        rc = read_steps(after);
        if (rc != VX3B_OK)
            return rc;
        lastVoxelSteps = 0;
        for (size_t w = 0; w < H.size(); w++)
            lastVoxelSteps += (after[w] - before[w]) * (int64_t)H[w].nv;
        return VX3B_OK;
    }

    // ---- one body over several GPUs: asynchronous stepping with the state kept in brick order, halo lists ----------------------
    // Brings the API-order arrays (and the end-of-run reductions) up to date after asynchronous runs; a no-op otherwise.
    int ensure_api_state() {
        if (!brickResident)
            return VX3B_OK;
        CUDA_TRY(cudaSetDevice(device));
        const unsigned wb = (A.nWorlds + 127) / 128, nbb = (unsigned)((std::max(BA.nvSlots, BA.nlSlots) + 255) / 256);
        k_brick_unpack<R><<<nbb, 256, 0, stream>>>(A, BA);
        k_clear_parity<R><<<wb, 128, 0, stream>>>(A);
        k_reduce_com<R><<<(unsigned)((A.nvTot + 255) / 256), 256, 0, stream>>>(A, 1);
        k_world_finish<R><<<wb, 128, 0, stream>>>(A, 1);
        CUDA_TRY(cudaGetLastError());
        brickResident = false;
        return VX3B_OK;
    }
    int run_async(int64_t steps) override {
        CUDA_TRY(cudaSetDevice(device));
        if (H.empty() || BA.nBricks == 0 || nBrickWorlds != (int)H.size() || !fusedAllWorlds || FA.nBricks > 0) {
            set_error("vx3b_batch_run_async: every world must be on the brick path (option path = 5 before upload; no collision, force field or targets)");
            return VX3B_ERR_STATE;
        }
        const unsigned wb = (A.nWorlds + 127) / 128, nbb = (unsigned)((std::max(BA.nvSlots, BA.nlSlots) + 255) / 256);
        k_world_begin<R><<<wb, 128, 0, stream>>>(A, (long long)steps, 0, 0, 1);
        if (!brickResident) {
            k_brick_pack<R><<<nbb, 256, 0, stream>>>(A, BA);
            brickResident = true;
        }
        int64_t done = 0;
        while (done < steps) {
            const int c = (int)std::min<int64_t>(chunk, steps - done);
            int rc = build_graph(c, 0, false, true);
            if (rc != VX3B_OK)
                return rc;
            CUDA_TRY(cudaGraphLaunch(graphExec, stream));
            done += c;
        }
        return VX3B_OK;
    }
    int halo_record_bytes() override { return (int)sizeof(HaloRec<R>); }
    int halo_set(int world, int listId, const int32_t *voxels, int n) override {
        CUDA_TRY(cudaSetDevice(device));
        if (!check_world(world) || listId < 0 || listId >= 8 || n < 0 || (n > 0 && !voxels)) {
            set_error("vx3b_batch_halo_set: bad world, list id or list");
            return VX3B_ERR_ARG;
        }
        const WorldHost &h = H[world];
        std::vector<int32_t> vox(n), slots(n);
        for (int i = 0; i < n; i++) {
            if (voxels[i] < 0 || voxels[i] >= h.nv) {
                set_error("vx3b_batch_halo_set: voxel index out of range");
                return VX3B_ERR_ARG;
            }
            vox[i] = h.voff + voxels[i];
            slots[i] = hVoxSlot.empty() ? -1 : hVoxSlot[vox[i]];
            if (slots[i] < 0 && BA.nBricks > 0 && !hVoxSlot.empty()) {
                set_error("vx3b_batch_halo_set: voxel is not on the brick path");
                return VX3B_ERR_ARG;
            }
        }
        if (hVoxSlot.empty() && BA.nBricks > 0) { // no ghost anywhere in the batch: build the slot map now
            std::vector<int32_t> borig(BA.nvSlots);
            CUDA_TRY(cudaMemcpy(borig.data(), BA.borig, borig.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
            hVoxSlot.assign(A.nvTot, -1);
            for (size_t sl = 0; sl < borig.size(); sl++)
                if (borig[sl] >= 0)
                    hVoxSlot[borig[sl]] = (int32_t)sl;
            for (int i = 0; i < n; i++)
                slots[i] = hVoxSlot[vox[i]];
        }
        HaloList &L = halo[listId];
        L.world = world;
        L.n = n;
        int rc;
        if ((rc = dalloc(&L.dVox, n, vox.data())) != VX3B_OK || (rc = dalloc(&L.dSlots, n, slots.data())) != VX3B_OK)
            return rc;
        CUDA_TRY(cudaStreamSynchronize(stream)); // the staging vectors die here
        return VX3B_OK;
    }
    int halo_transfer(int world, int listId, void *buf, bool exporting) {
        CUDA_TRY(cudaSetDevice(device));
        if (listId < 0 || listId >= 8 || halo[listId].world != world || !buf) {
            set_error("vx3b_batch_halo_export/import: list not registered for this world");
            return VX3B_ERR_ARG;
        }
        const HaloList &L = halo[listId];
        if (L.n == 0)
            return VX3B_OK;
        const unsigned nb = (unsigned)((L.n + 127) / 128);
        if (exporting)
            k_halo_export<R><<<nb, 128, 0, stream>>>(A, BA, world, L.dVox, L.dSlots, L.n, brickResident ? 1 : 0, (HaloRec<R> *)buf);
        else
            k_halo_import<R><<<nb, 128, 0, stream>>>(A, BA, world, L.dVox, L.dSlots, L.n, brickResident ? 1 : 0, (const HaloRec<R> *)buf);
        CUDA_TRY(cudaGetLastError());
        return VX3B_OK;
    }
    int halo_export(int world, int listId, void *buf) override { return halo_transfer(world, listId, buf, true); }
    int halo_import(int world, int listId, const void *buf) override { return halo_transfer(world, listId, const_cast<void *>(buf), false); }

    // ---- device-side exchange (include/vx3b200.h section 2b; kernels: k_halo_push / k_halo_pull) -----------------------------------
    // Receive area of a ghost list, in this device's memory: [flag, padded to 128 B][buffer 0: n records][buffer 1: n records].
    // Plain cudaMalloc (not the arena): cudaIpcGetMemHandle needs an allocation of its own.
    static size_t halo_area_bytes(int n) { return 128 + 2 * (size_t)n * sizeof(HaloRec<R>); }
    void release_halo_links() {
        for (void *q : peerOpened)
            cudaIpcCloseMemHandle(q);
        peerOpened.clear();
        if (peer.sync)
            cudaFree(peer.sync);
        if (peer.sendLo)
            cudaFree((void *)peer.sendLo);
        if (peer.sendHi)
            cudaFree((void *)peer.sendHi);
        peer = BrickPeer<R>{};
        peerOn = false;
        peerLists = stagedLists = 0;
        for (void *q : peerExported)
            cudaFree(q);
        peerExported.clear();
        for (auto &k : haloLink) {
            if (k.peer && k.peerIpc)
                cudaIpcCloseMemHandle(k.peer);
            if (k.area)
                cudaFree(k.area);
            if (k.sync)
                cudaFree(k.sync);
            k = HaloLink{};
        }
        for (auto &kv : slabGraphs)
            cudaGraphExecDestroy(kv.second);
        slabGraphs.clear();
    }
    int halo_sync_alloc(HaloLink &K) {
        if (!K.sync) {
            CUDA_TRY(cudaMalloc(&K.sync, sizeof(HaloSync)));
            CUDA_TRY(cudaMemset(K.sync, 0, sizeof(HaloSync)));
        }
        return VX3B_OK;
    }
    int halo_open(int world, int listId, VX3B_HaloEndpoint *out) override {
        CUDA_TRY(cudaSetDevice(device));
        if (listId < 0 || listId >= 8 || halo[listId].world != world || halo[listId].n <= 0) {
            set_error("vx3b_batch_halo_open: list not registered for this world (vx3b_batch_halo_set), or empty");
            return VX3B_ERR_ARG;
        }
        HaloLink &K = haloLink[listId];
        if (K.peer) {
            set_error("vx3b_batch_halo_open: this list is already connected as a send list");
            return VX3B_ERR_STATE;
        }
        int rc = halo_sync_alloc(K);
        if (rc != VX3B_OK)
            return rc;
        if (!K.area) {
            CUDA_TRY(cudaMalloc(&K.area, halo_area_bytes(halo[listId].n)));
            CUDA_TRY(cudaMemset(K.area, 0, halo_area_bytes(halo[listId].n)));
        }
        memset(out, 0, sizeof(*out));
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, K.area) == cudaSuccess) // (fails e.g. under a profiler's restrictions: same-process use still works)
            memcpy(out->ipc_handle, &h, sizeof(h) < sizeof(out->ipc_handle) ? sizeof(h) : sizeof(out->ipc_handle));
        else
            cudaGetLastError();
        out->device_ptr = (uint64_t)(uintptr_t)K.area;
        out->device = device;
        out->n = halo[listId].n;
        out->record_bytes = (int32_t)sizeof(HaloRec<R>);
        if (!peerExported.empty() && BA.nBricks > 0) { // for the exchange fused into the step kernel
            void *arr[6] = {BA.set[0].pxy, BA.set[0].pzt, BA.set[0].orient, BA.set[1].pxy, BA.set[1].pzt, BA.set[1].orient};
            for (int k = 0; k < 6; k++) {
                out->state_ptr[k] = (uint64_t)(uintptr_t)arr[k];
                if (cudaIpcGetMemHandle(&h, arr[k]) == cudaSuccess)
                    memcpy(out->state_ipc_handle[k], &h, sizeof(h) < 64 ? sizeof(h) : 64);
                else
                    cudaGetLastError();
            }
        }
        slabGraphsStale = true;
        return VX3B_OK;
    }
    int halo_slots(int world, int listId, int32_t *slots) override {
        CUDA_TRY(cudaSetDevice(device));
        if (listId < 0 || listId >= 8 || halo[listId].world != world || halo[listId].n <= 0 || BA.nBricks == 0) {
            set_error("vx3b_batch_halo_slots: list not registered for this world, empty, or not on the brick path");
            return VX3B_ERR_ARG;
        }
        CUDA_TRY(cudaMemcpy(slots, halo[listId].dSlots, (size_t)halo[listId].n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        return VX3B_OK;
    }
    // fused form: map the neighbour's six state arrays, note where it keeps my boundary voxels, mark them in the static words
    int connect_fused(int listId, const VX3B_HaloEndpoint *ep, int sameProcess, const int32_t *peerSlots, void *peerArea) {
        const HaloList &L = halo[listId];
        const bool lower = listId == 0; // list 0 feeds the lower neighbour, any other id the upper one (gpuvoxels_b200/slab.py: SEND_LO, SEND_HI)
        void *arr[6];
        for (int k = 0; k < 6; k++) {
            if (!ep->state_ptr[k]) {
                set_error("vx3b_batch_halo_connect_fused: the neighbour's endpoint does not export its state arrays (not a slab, or not on the brick path)");
                return VX3B_ERR_ARG;
            }
            if (sameProcess)
                arr[k] = (void *)(uintptr_t)ep->state_ptr[k];
            else {
                cudaIpcMemHandle_t h;
                memcpy(&h, ep->state_ipc_handle[k], sizeof(h) < 64 ? sizeof(h) : 64);
                CUDA_TRY(cudaIpcOpenMemHandle(&arr[k], h, cudaIpcMemLazyEnablePeerAccess));
                peerOpened.push_back(arr[k]);
            }
        }
        if (!peer.sync) {
            CUDA_TRY(cudaMalloc(&peer.sync, sizeof(BrickPeerSync)));
            CUDA_TRY(cudaMemset(peer.sync, 0, sizeof(BrickPeerSync)));
        }
        int32_t *&table = const_cast<int32_t *&>(lower ? peer.sendLo : peer.sendHi);
        if (!table)
            CUDA_TRY(cudaMalloc(&table, (size_t)BA.nvSlots * sizeof(int32_t)));
        int32_t *dRemote = nullptr;
        CUDA_TRY(cudaMalloc(&dRemote, (size_t)L.n * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(dRemote, peerSlots, (size_t)L.n * sizeof(int32_t), cudaMemcpyHostToDevice));
        k_brick_mark<<<(unsigned)((L.n + 255) / 256), 256, 0, stream>>>(const_cast<uint32_t *>(BA.bstat), L.dSlots, L.n, lower ? BS_SEND_LO : BS_SEND_HI, table, dRemote);
        CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(dRemote);
        using T4 = typename Store<R>::T4;
        if (lower) {
            peer.loPxy0 = (D2 *)arr[0], peer.loPzt0 = (D2 *)arr[1], peer.loOri0 = (T4 *)arr[2];
            peer.loPxy1 = (D2 *)arr[3], peer.loPzt1 = (D2 *)arr[4], peer.loOri1 = (T4 *)arr[5];
            peer.loFlag = (unsigned long long *)peerArea;
        } else {
            peer.hiPxy0 = (D2 *)arr[0], peer.hiPzt0 = (D2 *)arr[1], peer.hiOri0 = (T4 *)arr[2];
            peer.hiPxy1 = (D2 *)arr[3], peer.hiPzt1 = (D2 *)arr[4], peer.hiOri1 = (T4 *)arr[5];
            peer.hiFlag = (unsigned long long *)peerArea;
        }
        peerLists++;
        return VX3B_OK;
    }
    int halo_connect(int world, int listId, const VX3B_HaloEndpoint *peer, int sameProcess, const int32_t *peerSlots) override {
        CUDA_TRY(cudaSetDevice(device));
        if (listId < 0 || listId >= 8 || halo[listId].world != world || halo[listId].n <= 0) {
            set_error("vx3b_batch_halo_connect: list not registered for this world (vx3b_batch_halo_set), or empty");
            return VX3B_ERR_ARG;
        }
        if (peer->n != halo[listId].n || peer->record_bytes != (int32_t)sizeof(HaloRec<R>)) {
            set_error("vx3b_batch_halo_connect: the neighbour's ghost list has " + std::to_string(peer->n) + " records of " + std::to_string(peer->record_bytes) +
                      " bytes, this send list " + std::to_string(halo[listId].n) + " of " + std::to_string(sizeof(HaloRec<R>)));
            return VX3B_ERR_ARG;
        }
        HaloLink &K = haloLink[listId];
        if (K.area) {
            set_error("vx3b_batch_halo_connect: this list is already open as a ghost list");
            return VX3B_ERR_STATE;
        }
        int rc = halo_sync_alloc(K);
        if (rc != VX3B_OK)
            return rc;
        if (K.peer && K.peerIpc)
            cudaIpcCloseMemHandle(K.peer);
        K.peer = nullptr;
        if (sameProcess) {
            K.peer = (void *)(uintptr_t)peer->device_ptr;
            K.peerIpc = false;
            if (peer->device != device) { // another device of this process: needs peer access (NVLink)
                cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error(std::string("vx3b_batch_halo_connect: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                    return VX3B_ERR_CUDA;
                }
                cudaGetLastError();
            }
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, peer->ipc_handle, sizeof(h) < sizeof(peer->ipc_handle) ? sizeof(h) : sizeof(peer->ipc_handle));
            CUDA_TRY(cudaIpcOpenMemHandle(&K.peer, h, cudaIpcMemLazyEnablePeerAccess));
            K.peerIpc = true;
        }
        slabGraphsStale = true;
        if (peerSlots)
            return connect_fused(listId, peer, sameProcess, peerSlots, K.peer);
        stagedLists++;
        return VX3B_OK;
    }
    // one launch for every connected send list, then one for every opened ghost list (two sides per launch: blockIdx.y)
    void enqueue_exchange() {
        for (int pull = 0; pull < 2; pull++) {
            HaloSide<R> side[8];
            int world = -1, ns = 0, nmax = 0;
            for (int id = 0; id < 8; id++) {
                const HaloLink &K = haloLink[id];
                const HaloList &L = halo[id];
                char *base = (char *)(pull ? K.area : K.peer);
                if (!base || L.n <= 0)
                    continue;
                side[ns++] = HaloSide<R>{L.dSlots, L.n, (HaloRec<R> *)(base + 128), (HaloRec<R> *)(base + 128) + L.n, (unsigned long long *)base, K.sync};
                world = L.world;
                nmax = std::max(nmax, L.n);
            }
            for (int k = 0; k < ns; k += 2) {
                const dim3 grid((unsigned)((nmax + HALO_THREADS - 1) / HALO_THREADS), k + 1 < ns ? 2 : 1);
                if (pull)
                    k_halo_pull<R><<<grid, HALO_THREADS, 0, stream>>>(A, BA, world, side[k], side[k + 1 < ns ? k + 1 : k]);
                else
                    k_halo_push<R><<<grid, HALO_THREADS, 0, stream>>>(A, BA, world, side[k], side[k + 1 < ns ? k + 1 : k]);
            }
        }
    }
    int run_slab(int64_t steps) override {
        CUDA_TRY(cudaSetDevice(device));
        if (H.empty() || BA.nBricks == 0 || nBrickWorlds != (int)H.size() || !fusedAllWorlds || FA.nBricks > 0) {
            set_error("vx3b_batch_run_slab: every world must be on the brick path (option path = 5 before upload; no collision, force field or targets)");
            return VX3B_ERR_STATE;
        }
        for (int id = 0, w = -1; id < 8; id++)
            if ((haloLink[id].area || haloLink[id].peer) && halo[id].n > 0) {
                if (w >= 0 && halo[id].world != w) {
                    set_error("vx3b_batch_run_slab: the connected halo lists must belong to one world");
                    return VX3B_ERR_STATE;
                }
                w = halo[id].world;
            }
        const unsigned wb = (A.nWorlds + 127) / 128, nbb = (unsigned)((std::max(BA.nvSlots, BA.nlSlots) + 255) / 256);
        k_world_begin<R><<<wb, 128, 0, stream>>>(A, (long long)steps, 0, 0, 1);
        if (!brickResident) {
            k_brick_pack<R><<<nbb, 256, 0, stream>>>(A, BA);
            brickResident = true;
        }
        if (slabGraphsStale) {
            for (auto &kv : slabGraphs)
                cudaGraphExecDestroy(kv.second);
            slabGraphs.clear();
            slabGraphsStale = false;
            // fused form when every connected send list is: the opened ghost lists are then fed by the neighbours' kernels
            if (peerLists > 0 && stagedLists > 0) {
                set_error("vx3b_batch_run_slab: connect every send list the same way (vx3b_batch_halo_connect or vx3b_batch_halo_connect_fused)");
                return VX3B_ERR_STATE;
            }
            const bool fused = peerLists > 0 && slabFused;
            if (fused && !peerOn) {
                for (int id = 0; id < 8; id++) {
                    const HaloLink &K = haloLink[id];
                    const HaloList &L = halo[id];
                    if (!K.area || L.n <= 0)
                        continue;
                    // the ghost list that sits below my own voxels is fed by the lower neighbour (gpuvoxels_b200/slab.py: GHOST_LO = 2, GHOST_HI = 3)
                    (id == 2 ? peer.myLo : peer.myHi) = (const unsigned long long *)K.area;
                    k_brick_mark<<<(unsigned)((L.n + 255) / 256), 256, 0, stream>>>(const_cast<uint32_t *>(BA.bstat), L.dSlots, L.n, BS_REMOTE_FED, nullptr, nullptr);
                }
                CUDA_TRY(cudaGetLastError());
            }
            peerOn = fused;
        }
        int64_t done = 0;
        while (done < steps) {
            const int c = (int)std::min<int64_t>(chunk, steps - done);
            auto it = slabGraphs.find(c);
            if (it == slabGraphs.end()) {
                cudaGraph_t g;
                cudaGraphExec_t ge;
                CUDA_TRY(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                for (int k = 0; k < c; k++) {
                    enqueue_step(0, false, true);
                    if (!peerOn)
                        enqueue_exchange();
                }
                CUDA_TRY(cudaStreamEndCapture(stream, &g));
                CUDA_TRY(cudaGraphInstantiate(&ge, g, 0));
                cudaGraphDestroy(g);
                it = slabGraphs.emplace(c, ge).first;
            }
            CUDA_TRY(cudaGraphLaunch(it->second, stream));
            done += c;
        }
        return VX3B_OK;
    }
    int halo_timeouts(int64_t *n, int64_t *waitNs) override {
        CUDA_TRY(cudaSetDevice(device));
        *n = 0;
        if (waitNs)
            *waitNs = 0;
        if (peer.sync) {
            BrickPeerSync ps;
            CUDA_TRY(cudaMemcpy(&ps, peer.sync, sizeof(ps), cudaMemcpyDeviceToHost));
            *n += ps.timeouts;
            if (waitNs)
                *waitNs += (int64_t)ps.waitNs;
        }
        for (auto &K : haloLink)
            if (K.sync) {
                HaloSync hs;
                CUDA_TRY(cudaMemcpy(&hs, K.sync, sizeof(hs), cudaMemcpyDeviceToHost));
                *n += hs.timeouts;
                if (waitNs)
                    *waitNs += (int64_t)hs.waitNs;
            }
        return VX3B_OK;
    }

    std::vector<WorldD> resultScratch;
    int read_steps(std::vector<int64_t> &out) { // only the step counters: a strided copy of 8 bytes per world, not 5.5 KB
        CUDA_TRY(cudaMemcpy2DAsync(out.data(), sizeof(int64_t), (const char *)A.worlds + offsetof(WorldD, step), sizeof(WorldD), sizeof(int64_t),
                                   H.size(), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        return VX3B_OK;
    }

    int results(VX3B_WorldResult *out) override {
        CUDA_TRY(cudaSetDevice(device));
        {
            int rcs = ensure_api_state();
            if (rcs != VX3B_OK)
                return rcs;
        }
        if (!out || H.empty()) {
            set_error("vx3b_batch_results: nothing uploaded");
            return VX3B_ERR_STATE;
        }
        std::vector<WorldD> &hW = resultScratch; // kept between calls: a fresh 5.6 MB vector per generation is 1,400 page faults
        hW.resize(H.size());
        CUDA_TRY(cudaMemcpyAsync(hW.data(), A.worlds, hW.size() * sizeof(WorldD), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (size_t w = 0; w < H.size(); w++) {
            const WorldD &W = hW[w];
            VX3B_WorldResult &r = out[w];
            memset(&r, 0, sizeof(r));
            r.current_time = W.t;
            r.fitness_score = W.fitness;
            for (int k = 0; k < 3; k++) {
                r.initial_com[k] = W.initialCom[k];
                r.current_com[k] = W.finalCom[k];
            }
            r.total_distance_of_all_voxels = W.totalDistance;
            r.vox_size = W.voxSize;
            r.recent_angle = W.recentAngle;
            r.target_closeness = W.closeness;
            r.dt = W.dt;
            r.steps = W.step;
            r.num_voxel = W.nv;
            r.num_close_pairs = W.numClosePairs;
            r.collision_count = W.collisionCount;
            r.diverged = W.diverged;
            r.stopped = W.stopped;
            r.contact_rebuilds = W.contactRebuilds;
            r.contact_overflows = W.contactOverflows;
            r.num_measured_voxel = H[w].numMeasured;
        }
        return VX3B_OK;
    }

    bool check_world(int world) {
        if (world < 0 || world >= (int)H.size()) {
            set_error("world index out of range");
            return false;
        }
        return true;
    }

    int read_state(int world, double *pos, double *orient, double *lin, double *ang) override {
        CUDA_TRY(cudaSetDevice(device));
        {
            int rcs = ensure_api_state();
            if (rcs != VX3B_OK)
                return rcs;
        }
        if (!check_world(world))
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        std::vector<D2> pxy(h.nv), pzt(h.nv);
        std::vector<T4> o(h.nv), lm(h.nv);
        std::vector<T2> am(h.nv);
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(pxy.data(), A.pxy + h.voff, h.nv * sizeof(D2), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(pzt.data(), A.pzt + h.voff, h.nv * sizeof(D2), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(o.data(), A.orient + h.voff, h.nv * sizeof(T4), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(lm.data(), A.lm + h.voff, h.nv * sizeof(T4), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(am.data(), A.am + h.voff, h.nv * sizeof(T2), cudaMemcpyDeviceToHost));
        for (int v = 0; v < h.nv; v++) {
            if (pos) {
                pos[v * 3] = pxy[v].x;
                pos[v * 3 + 1] = pxy[v].y;
                pos[v * 3 + 2] = pzt[v].x;
            }
            if (orient) {
                orient[v * 4] = o[v].x;
                orient[v * 4 + 1] = o[v].y;
                orient[v * 4 + 2] = o[v].z;
                orient[v * 4 + 3] = o[v].w;
            }
            if (lin) {
                lin[v * 3] = lm[v].x;
                lin[v * 3 + 1] = lm[v].y;
                lin[v * 3 + 2] = lm[v].z;
            }
            if (ang) {
                ang[v * 3] = lm[v].w;
                ang[v * 3 + 1] = am[v].x;
                ang[v * 3 + 2] = am[v].y;
            }
        }
        return VX3B_OK;
    }

    // CVX_Material::setGlobalDamping on a live world (the reference's kineticFriction test lowers it after the voxel has settled,
    // VoxCAD/src/VX2/test/tVoxelyze.h:1111-1158): globalDampingTranslateC / RotateC = zeta * 2*sqrt(m*E*s) / 2*sqrt(I*E*s^3)
    // (VX_MaterialVoxel.h).  Every kernel reads the material table from global memory at launch, so the next run sees it.
    std::vector<DVoxMat> hVmats;
    std::vector<float> hDampT, hDampR;
    int set_global_damping(int world, int mat, double zeta) override {
        CUDA_TRY(cudaSetDevice(device));
        if (!check_world(world))
            return VX3B_ERR_ARG;
        std::vector<WorldD> one(1);
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(one.data(), A.worlds + world, sizeof(WorldD), cudaMemcpyDeviceToHost));
        if (mat < 0 || mat >= one[0].nvm) {
            set_error("vx3b_batch_set_global_damping: no such material");
            return VX3B_ERR_ARG;
        }
        const size_t k = (size_t)one[0].vmatOff + (size_t)mat;
        hVmats[k].globalDampTranslateC = (float)zeta * hDampT[k];
        hVmats[k].globalDampRotateC = (float)zeta * hDampR[k];
        CUDA_TRY(cudaMemcpy(A.vmats + k, &hVmats[k], sizeof(DVoxMat), cudaMemcpyHostToDevice));
        return VX3B_OK;
    }

    // CVX_External::setForce / setMoment on a live world (the reference's friction, impulse and large-deformation tests change
    // the load between steps, VoxCAD/src/VX2/test/tVoxelyze.h:471-605,1033-1160).  The voxel needs a boundary-condition slot:
    // it has one if it was fixed or loaded at upload, or if option ext_all was set before the upload.
    int set_force(int world, int voxel, const double *force, const double *moment) override {
        CUDA_TRY(cudaSetDevice(device));
        if (!check_world(world))
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        if (voxel < 0 || voxel >= h.nv) {
            set_error("vx3b_batch_set_force: no such voxel");
            return VX3B_ERR_ARG;
        }
        const int ei = staging.vext[h.voff + voxel];
        if (ei < 0) {
            set_error("vx3b_batch_set_force: the voxel has no boundary-condition slot (set option ext_all = 1 before vx3b_batch_upload)");
            return VX3B_ERR_STATE;
        }
        DExt &e = staging.ext[ei];
        for (int k = 0; k < 3; k++) {
            if (force)
                e.force[k] = (float)force[k];
            if (moment)
                e.moment[k] = (float)moment[k];
        }
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(A.ext + ei, &e, sizeof(DExt), cudaMemcpyHostToDevice));
        return VX3B_OK;
    }

    int read_aux(int world, float *voxelTemp, float *linkStrain) override {
        CUDA_TRY(cudaSetDevice(device));
        {
            int rcs = ensure_api_state();
            if (rcs != VX3B_OK)
                return rcs;
        }
        if (!check_world(world))
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        if (voxelTemp && h.nv) { // temperature = low word of pzt's second double
            std::vector<D2> b(h.nv);
            CUDA_TRY(cudaMemcpyAsync(b.data(), A.pzt + h.voff, sizeof(D2) * h.nv, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            for (int v = 0; v < h.nv; v++) {
                unsigned long long bits;
                memcpy(&bits, &b[v].y, 8);
                const uint32_t lo = (uint32_t)(bits & 0xffffffffu);
                memcpy(&voxelTemp[v], &lo, 4);
            }
        }
        if (linkStrain && h.nl) { // strain of the last step = la2.w
            std::vector<T4> a(h.nl);
            CUDA_TRY(cudaMemcpyAsync(a.data(), A.la2 + h.loff, sizeof(T4) * h.nl, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            for (int l = 0; l < h.nl; l++)
                linkStrain[l] = (float)a[linkSlot[h.loff + l]].w;
        }
        return VX3B_OK;
    }

    int read_initial_pos(int world, double *pos) override {
        CUDA_TRY(cudaSetDevice(device));
        if (!check_world(world) || !pos)
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        std::vector<D2> pxy(h.nv);
        std::vector<double> pz(h.nv);
        CUDA_TRY(cudaMemcpy(pxy.data(), A.ipxy + h.voff, h.nv * sizeof(D2), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(pz.data(), A.ipz + h.voff, h.nv * sizeof(double), cudaMemcpyDeviceToHost));
        for (int v = 0; v < h.nv; v++) {
            pos[v * 3] = pxy[v].x;
            pos[v * 3 + 1] = pxy[v].y;
            pos[v * 3 + 2] = pz[v];
        }
        return VX3B_OK;
    }

    int read_links(int world, double *pos2, double *a1, double *a2, int32_t *flags) override {
        CUDA_TRY(cudaSetDevice(device));
        {
            int rcs = ensure_api_state();
            if (rcs != VX3B_OK)
                return rcs;
        }
        if (!check_world(world))
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        std::vector<T4> p(h.nl), x1(h.nl), x2(h.nl);
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(p.data(), A.lpos2 + h.loff, h.nl * sizeof(T4), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(x1.data(), A.la1 + h.loff, h.nl * sizeof(T4), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(x2.data(), A.la2 + h.loff, h.nl * sizeof(T4), cudaMemcpyDeviceToHost));
        for (int l = 0; l < h.nl; l++) {
            const int i = linkSlot[h.loff + l]; // creation order out, internal (axis-sorted) order in
            if (pos2) {
                pos2[l * 3] = p[i].x;
                pos2[l * 3 + 1] = p[i].y;
                pos2[l * 3 + 2] = p[i].z;
            }
            if (a1) {
                a1[l * 3] = x1[i].x;
                a1[l * 3 + 1] = x1[i].y;
                a1[l * 3 + 2] = x1[i].z;
            }
            if (a2) {
                a2[l * 3] = x2[i].x;
                a2[l * 3 + 1] = x2[i].y;
                a2[l * 3 + 2] = x2[i].z;
            }
            if (flags) {
                uint32_t f;
                if constexpr (sizeof(R) == 4)
                    memcpy(&f, &p[i].w, 4);
                else {
                    long long b;
                    memcpy(&b, &p[i].w, 8);
                    f = (uint32_t)b;
                }
                flags[l] = (int32_t)f;
            }
        }
        return VX3B_OK;
    }

    // the static engine's share of the dynamic-topology read-backs: nothing changes, so they report the uploaded topology and the
    // reference's initial signal state
    int num_links(int world, int32_t *n) override {
        if (!check_world(world) || !n)
            return VX3B_ERR_ARG;
        *n = H[world].nl;
        return VX3B_OK;
    }
    int read_signals(int world, double *sig6, int32_t *vflags2) override {
        if (!check_world(world))
            return VX3B_ERR_ARG;
        if (sig6)
            std::fill_n(sig6, (size_t)H[world].nv * 6, 0.0);
        if (vflags2)
            for (int v = 0; v < H[world].nv; v++)
                vflags2[2 * v] = 0, vflags2[2 * v + 1] = 1;
        return VX3B_OK;
    }
    int read_topology(int world, int32_t *links6, int32_t *n) override {
        if (!check_world(world) || !n)
            return VX3B_ERR_ARG;
        const WorldHost &h = H[world];
        const int cap = *n;
        *n = h.nl;
        if (links6)
            for (int l = 0; l < h.nl && l < cap; l++) {
                const size_t g = (size_t)h.loff + linkSlot[(size_t)h.loff + l];
                int32_t *q = links6 + 6 * (size_t)l;
                q[0] = staging.lends[g].x - h.voff, q[1] = staging.lends[g].y - h.voff, q[2] = (int32_t)(staging.lmeta[g] & 3u), q[3] = q[4] = q[5] = 0;
            }
        return VX3B_OK;
    }

    int collision_pairs(int world, int32_t *pairs, int32_t *n) override {
        CUDA_TRY(cudaSetDevice(device));
        {
            int rcs = ensure_api_state();
            if (rcs != VX3B_OK)
                return rcs;
        }
        if (!check_world(world) || !n)
            return VX3B_ERR_ARG;
        if (!recordPairs) {
            set_error("vx3b_batch_collision_pairs: set option record_pairs=1 before vx3b_batch_upload");
            return VX3B_ERR_STATE;
        }
        int count = 0;
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(&count, A.pairCount, sizeof(int), cudaMemcpyDeviceToHost));
        if (count > A.pairCap) {
            set_error("vx3b_batch_collision_pairs: pair buffer overflow");
            return VX3B_ERR_STATE;
        }
        std::vector<int4> buf(count);
        if (count)
            CUDA_TRY(cudaMemcpy(buf.data(), A.pairBuf, count * sizeof(int4), cudaMemcpyDeviceToHost));
        std::vector<std::pair<int, int>> mine;
        for (auto &p : buf)
            if (p.x == world)
                mine.emplace_back(p.y, p.z);
        std::sort(mine.begin(), mine.end());
        int cap = *n;
        *n = (int)mine.size();
        if (pairs)
            for (int i = 0; i < (int)mine.size() && i < cap; i++) {
                pairs[i * 2] = mine[i].first;
                pairs[i * 2 + 1] = mine[i].second;
            }
        return VX3B_OK;
    }
};


} // namespace vx3b
