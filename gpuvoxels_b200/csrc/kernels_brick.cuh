// Brick path: the fused link + voxel step for large bodies (BASELINE configs[2]: 100^3) on a BRICK-ORDERED copy of the state.
//
// k_fused (kernels_fused.cuh) showed where the time goes on a 1e6-voxel body: not HBM (25 % of peak) but instruction issue and
// exposed load latency - per-link index tables, gathered 16-byte loads with two dependent global loads per link, generic
// per-voxel feature tests, 3-way bank conflicts.  This path changes the data layout and the schedule instead of the arithmetic:
//   * at the start of a run the voxels and links of every eligible world are copied ("packed") into brick order: the own
//     voxels of a brick (BX x BY x BZ lattice cells, x fastest) and the links a brick owns (those whose POSITIVE end is an own
//     voxel; x-, y-, z-links in that order) are contiguous, 16-byte aligned runs of the state arrays;
//   * PERSISTENT CTAs, each stepping a contiguous chunk of bricks through a TWO-STAGE shared-memory ring: while the CTA
//     computes brick k out of one stage, the other stage fills with brick k+1 - its own runs, link history and link
//     descriptors by 1-D bulk copies (cp.async.bulk, the TMA engine's linear mode: UBLKCP in SASS) that complete on the
//     stage's mbarrier, its halo (voxels of neighbouring bricks that share a link with it, and the history of the links it
//     recomputes) by 16-byte cp.async gathers whose index list was itself bulk-copied one brick earlier.  Nothing in the two
//     compute phases waits on a global load;
//   * link forces travel from the link threads to the voxel threads through shared memory in a [voxel][37] layout (odd
//     stride: conflict-free on both sides, every access an immediate offset from one address); new state goes straight from
//     registers to HBM, coalesced, into the other buffer set (ping-pong by W.parity, as k_fused);
//   * the world's material tables are staged in shared memory (once per CTA while the world does not change) with the
//     per-material damping multiplier 2*sqrt(m)*zeta/dt and the per-link-material rest lengths already formed;
//   * the link axis is a template parameter (one switch per link instead of seven swizzle selects); the LEAN instantiation
//     (no boundary conditions, cilia, thermal actuation or Fixed materials in any brick world) compiles the feature tests away.
// Links cut by a brick face are recomputed by the brick that does not own them, from the same inputs, so no force crosses
// CTAs (same rule as k_fused).  Worlds with voxel-voxel contact, a force field or target materials stay on k_fused: their
// per-step helpers read the API-order arrays.
//
// Replaces gpu_update_links + gpu_update_voxels (+ gpu_update_temperature, gpu_update_cilia_force)
// (VX3/VX3_VoxelyzeKernel.cu:566-650, 846-859) for large worlds.  Same deliberate difference as k_fused on divergence.
#pragma once
#include "engine.cuh"
#include "kernels_contact.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

struct BrickB {
    int32_t world;
    int32_t v0, nOwn, nHalo;    // own voxels: brick-order slots [v0, v0+nOwn), v0 a multiple of 4
    int32_t l0, nOwned, nHaloL; // owned links: brick-order slots [l0, l0+nOwned); nHaloL recomputed links
    int32_t descBeg;            // even: descriptors of the recomputed links, then of the owned links
    int32_t gBeg;               // multiple of 4: gather list = nHalo voxel slots, then nHaloL link slots (padded to a multiple of 4)
    int32_t pad0, pad1, pad2;
};
// One link as seen from one brick.  a/b: table index of the negative / positive end (own voxel i -> i, halo voxel j -> nOwnPad + j)
struct alignas(8) LinkDesc {
    uint16_t a, b, lmat, flags; // lmat: world-local link material; flags: bits 0-1 axis, BD_NEG_OWN, BD_POS_OWN, bits 4-7 / 8-11 world-local voxel material of a / b
};
enum { BD_NEG_OWN = 4, BD_POS_OWN = 8 };
#ifndef BRICK_MAXT
#define BRICK_MAXT 512 // threads of a k_brick CTA (one CTA per SM); -DBRICK_MAXT=768 for experiments: 85 registers per thread
#endif
enum { BRICK_MAX_VMATS = 16, BRICK_MAX_LMATS = 48, BRICK_CHUNK = 64, BRICK_FSTRIDE = 37 };

// One brick-ordered buffer set.  Unlike the API-order StateSet the angular momentum has its own 16-byte array (am = angMom.xyz, 0;
// lm = linMom.xyz, 0), so the translation half and the rotation half of a voxel update write disjoint arrays, and the VX3 "static
// friction zeroes angMom" rule travels as a flag bit (VF_ZERO_ANGMOM in pzt) applied by whoever loads angMom next.
template <class R> struct BrickSet {
    using T4 = typename Store<R>::T4;
    D2 *pxy, *pzt;
    T4 *orient, *lm, *am;
    T4 *lpos2, *la1, *la2;
};
template <class R> struct BrickArrays {
    int nBricks, chunk;      // chunk: consecutive bricks per CTA and round (<= BRICK_CHUNK)
    const BrickB *bricks;
    const LinkDesc *ldesc;
    const int32_t *glist;    // gather lists (brick-order voxel slots, then link slots)
    const uint32_t *bstat;   // per brick-order voxel slot: world-local material | link mask << 8 | has boundary conditions << 14 | ghost << 15
    const int32_t *borig;    // brick-order voxel slot -> global (API-order) voxel index, -1 for alignment padding
    const int32_t *bsurf;    // brick-order voxel slot -> global surface slot or -1 (allocated when a brick world has voxel-voxel contact)
    const int32_t *lorig;    // brick-order link slot -> global link index
    BrickSet<R> set[2];
    int nvSlots, nlSlots;
    const int32_t *worldList; // the brick worlds (for the end-of-step bookkeeping folded into k_brick)
    int nWorlds;
    int *ticket;              // CTAs that have finished the current launch
};
// One body over several GPUs, exchange FUSED into the step kernel (include/vx3b200.h section 2b): the voxel pass stores the new
// position / temperature / orientation of a boundary voxel not only into this slab's arrays but also - over NVLink, into the
// neighbouring slab's IPC-mapped arrays - into the slot that slab keeps for it as a ghost; the last CTA of the launch raises a
// sequence flag in the neighbour's memory, and the neighbour's next launch starts by waiting for it.  No pack / unpack kernels,
// no staging buffers: one launch per time step, as on one GPU.  Safe with two buffer sets: a slab can start step s + 1 (which
// overwrites, in the neighbour, the set that neighbour READ in step s) only after it has seen the neighbour's flag of step s.
enum { BS_SEND_LO = 1u << 16, BS_SEND_HI = 1u << 17, BS_REMOTE_FED = 1u << 18 }; // bits of the per-slot static word (BrickArrays::bstat)
struct BrickPeerSync {
    unsigned long long seq;    // launches completed = the sequence number the neighbours' flags must have reached before the next one
    unsigned long long waitNs; // time CTA 0 spent waiting for the neighbours (diagnostic)
    unsigned int timeouts;     // launches that gave up waiting (10 s): the run is invalid
    unsigned int pad;
};
template <class R> struct BrickPeer { // (no arrays: a dynamically indexed kernel parameter lives in local memory)
    using T4 = typename Store<R>::T4;
    D2 *loPxy0, *loPxy1, *loPzt0, *loPzt1; // the lower neighbour's buffer sets 0 / 1 (mapped), NULL without such a neighbour
    T4 *loOri0, *loOri1;
    D2 *hiPxy0, *hiPxy1, *hiPzt0, *hiPzt1;
    T4 *hiOri0, *hiOri1;
    const int32_t *sendLo, *sendHi;        // [my brick-order slot] -> the neighbour's slot of its ghost copy (read where BS_SEND_* is set)
    unsigned long long *loFlag, *hiFlag;   // in the neighbours' memory: raised to seq + 1 when this launch's stores are visible
    const unsigned long long *myLo, *myHi; // in my memory: raised by the lower / upper neighbour
    BrickPeerSync *sync;
};

struct BrickCaps {
    int own, halo, owned, haloL, nvm, nlm; // own even
    int tab() const { return own + halo; }
    int links() const { return (owned + haloL + 1) & ~1; }
    int glist() const { return (halo + haloL + 3) & ~3; }
};

template <class R> struct BrickLayout { // byte offsets into the CTA's dynamic shared memory; computed once on the host
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    uint32_t bar, hdr, glist0, glistBytes, F, vmat, lmat, damp, len, stage0, stageBytes; // stage s / list slot k at base + s * bytes
                                                                                           // (no arrays: a dynamically indexed kernel parameter lives in local memory)
    uint32_t pxy, pzt, ori, lm, am, h0, h1, h2, desc, stat; // relative to a stage
    uint32_t total;
    static uint32_t up(size_t x) { return (uint32_t)((x + 127) & ~(size_t)127); }
    BrickLayout() {}
    explicit BrickLayout(const BrickCaps &c) {
        uint32_t o = 0;
        auto take = [&](size_t bytes) {
            uint32_t at = o;
            o = up(o + bytes);
            return at;
        };
        bar = take(16);
        hdr = take(sizeof(BrickB) * BRICK_CHUNK);
        glist0 = take(sizeof(int32_t) * c.glist());
        glistBytes = take(sizeof(int32_t) * c.glist()) - glist0;
        F = take(sizeof(R) * BRICK_FSTRIDE * c.own);
        vmat = take(sizeof(DVoxMat) * c.nvm);
        lmat = take(sizeof(DLinkMat) * c.nlm);
        damp = take(sizeof(float) * c.nvm);
        len = take(sizeof(LinkLen<R>) * c.nlm);
        const uint32_t base = o;
        o = 0;
        pxy = take(sizeof(D2) * c.tab());
        pzt = take(sizeof(D2) * c.tab());
        ori = take(sizeof(T4) * c.tab());
        lm = take(sizeof(T4) * c.own);
        am = take(sizeof(T4) * c.own);
        h0 = take(sizeof(T4) * (c.owned + c.haloL));
        h1 = take(sizeof(T4) * (c.owned + c.haloL));
        h2 = take(sizeof(T4) * (c.owned + c.haloL));
        desc = take(sizeof(LinkDesc) * c.links());
        stat = take(sizeof(uint32_t) * ((c.own + 3) & ~3));
        stageBytes = o;
        stage0 = base;
        total = base + 2 * stageBytes;
    }
};

// ---- mbarrier / bulk-copy / cp.async PTX (the guide's TMA + mbarrier producer/consumer pattern in its 1-D form) ------------
VXD uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
VXD void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
VXD void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
VXD void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
VXD void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
VXD void mbar_wait(uint32_t bar, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{\n"
                     ".reg .pred p;\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                     "selp.u32 %0, 1, 0, p;\n"
                     "}"
                     : "=r"(done)
                     : "r"(bar), "r"(phase)
                     : "memory");
    } while (!done);
}
VXD unsigned long long halo_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
template <int BYTES> VXD void gather16(uint32_t dst, const void *src) { // BYTES = 16 or 32
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    if constexpr (BYTES == 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"((const char *)src + 16) : "memory");
}
VXD void gather_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <class R> VXD uint32_t hist_flags(typename Store<R>::T4 s0) {
    if constexpr (sizeof(R) == 4)
        return f2u(s0.w);
    else
        return (uint32_t)__double_as_longlong(s0.w);
}
template <class R> VXD R hist_flags_pack(uint32_t f) {
    if constexpr (sizeof(R) == 4)
        return u2f(f);
    else
        return __longlong_as_double((long long)f);
}

// One link of phase 1 with the axis as a compile-time constant (the caller switches once per link).
template <class R, bool LEAN, int AXIS>
VXD void brick_link(const DLinkMat &lmat, const LinkLen<R> &len0, const DVoxMat *sVm, const float *sDamp, int ma, int mb, double voxSize, D2 axy, D2 azt, D2 bxy,
                    D2 bzt, typename Store<R>::T4 oa, typename Store<R>::T4 ob, LinkHist<R> &h, LinkForces<R> &f) {
    V3<double> dpos{bxy.x - axy.x, bxy.y - axy.y, bzt.x - azt.x};
    Qt<R> qn{oa.x, oa.y, oa.z, oa.w}, qp{ob.x, ob.y, ob.z, ob.w};
    LinkLen<R> len = len0;
    if constexpr (!LEAN) {
        const float ta = u2f((uint32_t)((unsigned long long)__double_as_longlong(azt.y) & 0xffffffffu));
        const float tb = u2f((uint32_t)((unsigned long long)__double_as_longlong(bzt.y) & 0xffffffffu));
        len = link_len<R>(voxSize, 0.5 * (base_size(sVm[ma], ta) + base_size(sVm[mb], tb)));
    }
    link_update<R, AXIS>(lmat, AXIS, dpos, len, qn, qp, sDamp[ma], sDamp[mb], h, f);
}

// k_world_step's job for the brick worlds, by one thread; kept out of line so that the formula evaluator's stack frame and
// registers stay out of the step kernel proper
template <class R> __device__ __noinline__ void brick_world_step(WorldD *worlds, const int32_t *worldList, int nWorlds, int useStop, int *activeCount) {
    const Arrays<R> *none = nullptr; // brick worlds count no target pairs: the arrays are never looked at
    for (int i = 0; i < nWorlds; i++) {
        WorldD &W = worlds[worldList[i]];
        if (!W.active || !(W.flags & WF_BRICK))
            continue;
        W.parity ^= 1;
        cgrid_end_of_step(W);
        // fields other CTAs of the still-running kernel wrote - divergedNow with plain stores, comAcc with atomics - are fetched with
        // atomics and stored back, so that the plain loads of the bookkeeping code cannot be served from a stale L1 line of this SM
        // (a thread always sees its own stores)
        if (atomicOr(&W.divergedNow, 0))
            W.divergedNow = 1;
        for (int k = 0; k < 4; k++) {
            const double fresh = atomicAdd(&W.comAcc[k], 0.0);
            W.comAcc[k] = fresh;
        }
        world_end_of_step<R, false>(*none, W, useStop);
        if (W.active && activeCount)
            atomicAdd(activeCount, 1);
    }
}

// COLL: some brick world has voxel-voxel contact.  The contact kernels (kernels_contact.cuh, k_contact) run before this one and read
// SURFACE voxels from the main (API-order) arrays: the voxel pass takes a surface voxel's contact force from A.contactF, mirrors its new
// position and linear momentum into the main arrays, and does the stale-list check of the global contact grid G.
template <class R, bool LEAN, bool PEER = false, bool COLL = false>
__global__ void __launch_bounds__(BRICK_MAXT, 1) k_brick(Arrays<R> A, BrickArrays<R> BA, const BrickLayout<R> L, int worldStep, int useStop, int *activeCount, const BrickPeer<R> P = BrickPeer<R>{},
                                                          const ContactGrid G = ContactGrid{}) {
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    extern __shared__ __align__(128) unsigned char smem_brick[];
    unsigned char *smem = smem_brick;
    const int tid = threadIdx.x, nt = blockDim.x;
    const BrickB *sHdr = (const BrickB *)(smem + L.hdr);
    R *sF = (R *)(smem + L.F);
    DVoxMat *sVm = (DVoxMat *)(smem + L.vmat);
    DLinkMat *sLmat = (DLinkMat *)(smem + L.lmat);
    float *sDamp = (float *)(smem + L.damp);
    LinkLen<R> *sLen = (LinkLen<R> *)(smem + L.len);
    const uint32_t bar0 = smem_u32(smem + L.bar);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    if constexpr (PEER) { // the neighbours' stores of the previous step into my ghost slots must have landed
        if (tid == 0) {
            const unsigned long long want = P.sync->seq, t0 = halo_now_ns();
            bool ok = true;
            // (after one timeout the run is invalid anyway: later launches do not wait again, so a dead neighbour costs 10 s, not 10 s per step)
            while (P.sync->timeouts == 0 &&
                   ((P.myLo && *(const volatile unsigned long long *)P.myLo < want) || (P.myHi && *(const volatile unsigned long long *)P.myHi < want))) {
                __nanosleep(64);
                if (halo_now_ns() - t0 > 10000000000ull) { // 10 s: a neighbour is gone; never hang the device
                    ok = false;
                    break;
                }
            }
            __threadfence_system();
            if (blockIdx.x == 0) {
                P.sync->waitNs += halo_now_ns() - t0;
                if (!ok)
                    P.sync->timeouts++;
            }
        }
        __syncthreads();
        asm volatile("fence.proxy.async;" ::: "memory"); // the bulk copies below read what another device's ordinary stores wrote
    }
    uint32_t phaseBits = 0; // bit s: parity the next wait on stage s's mbarrier expects
    int curWorld = -1;
    // per-world scalars (valid while curWorld is)
    uint32_t wflags = 0;
    bool vx3 = false, active = false, sampling = false;
    float dt = 0;
    double voxSize = 0;
    BrickSet<R> src = BA.set[0], dst = BA.set[1];
    WorldD *Wp = nullptr;
    [[maybe_unused]] bool gridCheck = false; // COLL: this world's candidate lists live in the global contact grid G
    [[maybe_unused]] double refX = 0, refY = 0, refZ = 0; // ... and the position of its reference voxel at the start of the step

    // Fill stage `s` with brick B: bulk copies of its runs (+ the gather list of the brick after it, `next`, into its list slot)
    // on the stage's mbarrier by thread 0, halo gathers by everybody.  idx: the brick's gather list (shared or global memory).
    // ft / fnt: this thread's rank among, and the number of, the threads that take part.
    // what: 3 = everything, 1 = the bulk copies only (one thread: ft == 0), 2 = the halo gathers only
    auto fill = [&](const BrickB &B, int s, const int32_t *idx, bool idxShared, const BrickB *next, int nextSlot, int ft, int fnt, int what) {
        const uint32_t idxS = idxShared ? smem_u32(idx) : 0u;
        auto index = [&](int i) -> int {
            if (idxShared) {
                int g;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(g) : "r"(idxS + 4u * (uint32_t)i));
                return g;
            }
            return __ldg(idx + i);
        };
        unsigned char *st = smem + L.stage0 + (uint32_t)s * L.stageBytes;
        const int nOwnPad = (B.nOwn + 1) & ~1;
        if (ft == 0 && (what & 1)) {
            const uint32_t bar = bar0 + 8 * s;
            const uint32_t bV16 = (uint32_t)nOwnPad * 16u, bV4 = (uint32_t)nOwnPad * (uint32_t)sizeof(T4);
            const uint32_t bL = (uint32_t)B.nOwned * (uint32_t)sizeof(T4), bD = (uint32_t)((B.nHaloL + B.nOwned + 1) & ~1) * (uint32_t)sizeof(LinkDesc);
            const uint32_t bG = next ? (uint32_t)((next->nHalo + next->nHaloL + 3) & ~3) * 4u : 0u;
            const uint32_t bS = (uint32_t)((B.nOwn + 3) & ~3) * 4u; // per-voxel static words (material, link mask, flags)
            mbar_expect_tx(bar, 2 * bV16 + 3 * bV4 + 3 * bL + bD + bG + bS);
            if (nOwnPad) {
                bulk_g2s(smem_u32(st + L.pxy), src.pxy + B.v0, bV16, bar);
                bulk_g2s(smem_u32(st + L.pzt), src.pzt + B.v0, bV16, bar);
                bulk_g2s(smem_u32(st + L.ori), src.orient + B.v0, bV4, bar);
                bulk_g2s(smem_u32(st + L.lm), src.lm + B.v0, bV4, bar);
                bulk_g2s(smem_u32(st + L.am), src.am + B.v0, bV4, bar);
                bulk_g2s(smem_u32(st + L.stat), BA.bstat + B.v0, bS, bar);
            }
            if (B.nOwned) { // owned history lands behind the recomputed links' (gathered) history: one array indexed by link number
                const uint32_t off = (uint32_t)B.nHaloL * (uint32_t)sizeof(T4);
                bulk_g2s(smem_u32(st + L.h0) + off, src.lpos2 + B.l0, bL, bar);
                bulk_g2s(smem_u32(st + L.h1) + off, src.la1 + B.l0, bL, bar);
                bulk_g2s(smem_u32(st + L.h2) + off, src.la2 + B.l0, bL, bar);
            }
            if (bD)
                bulk_g2s(smem_u32(st + L.desc), BA.ldesc + B.descBeg, bD, bar);
            if (bG)
                bulk_g2s(smem_u32(smem + L.glist0 + (uint32_t)nextSlot * L.glistBytes), BA.glist + next->gBeg, bG, bar);
        }
        if (!(what & 2))
            return;
        const uint32_t dPxy = smem_u32(st + L.pxy) + (uint32_t)nOwnPad * 16u, dPzt = smem_u32(st + L.pzt) + (uint32_t)nOwnPad * 16u;
        const uint32_t dOri = smem_u32(st + L.ori) + (uint32_t)nOwnPad * (uint32_t)sizeof(T4);
#pragma unroll 1
        for (int i = ft; i < B.nHalo; i += fnt) {
            const int g = index(i);
            gather16<16>(dPxy + 16u * i, src.pxy + g);
            gather16<16>(dPzt + 16u * i, src.pzt + g);
            gather16<sizeof(T4)>(dOri + (uint32_t)sizeof(T4) * i, src.orient + g);
        }
        const uint32_t dH0 = smem_u32(st + L.h0), dH1 = smem_u32(st + L.h1), dH2 = smem_u32(st + L.h2);
#pragma unroll 1
        for (int i = ft; i < B.nHaloL; i += fnt) {
            const int g = index(B.nHalo + i);
            gather16<sizeof(T4)>(dH0 + (uint32_t)sizeof(T4) * i, src.lpos2 + g);
            gather16<sizeof(T4)>(dH1 + (uint32_t)sizeof(T4) * i, src.la1 + g);
            gather16<sizeof(T4)>(dH2 + (uint32_t)sizeof(T4) * i, src.la2 + g);
        }
    };

    for (int base = blockIdx.x * BA.chunk; base < BA.nBricks; base += gridDim.x * BA.chunk) {
        const int cnt = min(BA.chunk, BA.nBricks - base);
        __syncthreads(); // previous round's readers of the header array are done (also orders the mbarrier init)
        {
            const uint32_t *g = (const uint32_t *)(BA.bricks + base);
            uint32_t *d = (uint32_t *)(smem + L.hdr);
#pragma unroll 1
            for (int i = tid; i < cnt * (int)(sizeof(BrickB) / 4); i += nt)
                d[i] = g[i];
        }
        __syncthreads();
        bool filled = false; // brick k's stage is already in flight (prefetched during brick k-1)
        int s = 0;           // stage of brick k
#pragma unroll 1
        for (int k = 0; k < cnt; k++) {
            const BrickB B = sHdr[k];
            if (B.world != curWorld) { // (re)stage the world: scalars in registers, material tables in shared memory
                curWorld = B.world;
                Wp = A.worlds + B.world;
                const WorldD &W = *Wp;
                wflags = W.flags;
                vx3 = wflags & WF_VX3;
                active = W.active && (wflags & WF_BRICK);
                dt = W.dt;
                voxSize = W.voxSize;
                sampling = vx3 && W.cycle > 0 && ((W.step + 1) % W.cycle == 0);
                const bool par = W.parity != 0; // field-wise selects: indexing BA.set[] dynamically would put it in local memory
                const BrickSet<R> &s0 = BA.set[0], &s1 = BA.set[1];
                src.pxy = par ? s1.pxy : s0.pxy, src.pzt = par ? s1.pzt : s0.pzt, src.orient = par ? s1.orient : s0.orient;
                src.lm = par ? s1.lm : s0.lm, src.am = par ? s1.am : s0.am;
                src.lpos2 = par ? s1.lpos2 : s0.lpos2, src.la1 = par ? s1.la1 : s0.la1, src.la2 = par ? s1.la2 : s0.la2;
                dst.pxy = par ? s0.pxy : s1.pxy, dst.pzt = par ? s0.pzt : s1.pzt, dst.orient = par ? s0.orient : s1.orient;
                dst.lm = par ? s0.lm : s1.lm, dst.am = par ? s0.am : s1.am;
                dst.lpos2 = par ? s0.lpos2 : s1.lpos2, dst.la1 = par ? s0.la1 : s1.la1, dst.la2 = par ? s0.la2 : s1.la2;
                if constexpr (COLL) {
                    gridCheck = G.nBuckets > 0 && (wflags & WF_COLLISION) && W.gridLog2 > 0 && W.soff >= G.s0;
                    if (gridCheck) { // from the source set, which no thread writes during this launch
                        const D2 ra = src.pxy[W.gRefSlot], rb = src.pzt[W.gRefSlot];
                        refX = ra.x, refY = ra.y, refZ = rb.x;
                    }
                }
                if (active) {
                    __syncthreads(); // nobody still reads the previous world's tables
                    const uint32_t *gv = (const uint32_t *)(A.vmats + W.vmatOff), *gl = (const uint32_t *)(A.lmats + W.lmatOff);
                    uint32_t *sv = (uint32_t *)sVm, *sl = (uint32_t *)sLmat;
                    const int nwv = W.nvm * (int)(sizeof(DVoxMat) / 4), nwl = W.nlm * (int)(sizeof(DLinkMat) / 4);
#pragma unroll 1
                    for (int i = tid; i < nwv; i += nt)
                        sv[i] = gv[i];
#pragma unroll 1
                    for (int i = tid; i < nwl; i += nt)
                        sl[i] = gl[i];
                    if (tid < W.nvm) { // dampingMultiplier numerator / previousDt (VX3/VX3_Voxel.h:206-208); Fixed voxels never set previousDt
                        const DVoxMat &m = A.vmats[W.vmatOff + tid];
                        sDamp[tid] = m.sqrtMassZetaInt2 / ((vx3 && m.fixed) ? 0.0f : dt);
                    }
                    if (tid >= 32 && tid - 32 < W.nlm) { // rest lengths per link material at sizes the temperature cannot change
                        const uint32_t mp = A.lmats[W.lmatOff + tid - 32].matPair;
                        sLen[tid - 32] = link_len<R>(W.voxSize, 0.5 * (A.vmats[W.vmatOff + (mp & 0xffff)].baseSize0 + A.vmats[W.vmatOff + (mp >> 16)].baseSize0));
                    }
                    __syncthreads();
                }
            }
            if (!active) {
                filled = false;
                continue;
            }
            const bool haveNext = k + 1 < cnt && sHdr[k + 1].world == curWorld;
            if (!filled) { // first brick of a round / of a world: nothing was prefetched, indices come from global memory
                fill(B, s, BA.glist + B.gBeg, false, haveNext ? &sHdr[k + 1] : nullptr, (k + 1) & 1, tid, nt, 3);
                gather_wait_all();
                __syncthreads();
            }
            mbar_wait(bar0 + 8 * s, (phaseBits >> s) & 1);
            phaseBits ^= 1u << s;
            filled = haveNext; // brick k+1 is prefetched at the end of phase 1

            unsigned char *st = smem + L.stage0 + (uint32_t)s * L.stageBytes;
            const D2 *sPxy = (const D2 *)(st + L.pxy), *sPzt = (const D2 *)(st + L.pzt);
            const T4 *sOri = (const T4 *)(st + L.ori), *sLm = (const T4 *)(st + L.lm);
            const T4 *sAm = (const T4 *)(st + L.am);
            const T4 *sH0 = (const T4 *)(st + L.h0), *sH1 = (const T4 *)(st + L.h1), *sH2 = (const T4 *)(st + L.h2);
            const LinkDesc *sDesc = (const LinkDesc *)(st + L.desc);
            const int nOwn = B.nOwn, nLinks = B.nHaloL + B.nOwned;
            // phase 2 roles: with at least two threads per own voxel (and no per-voxel extras: LEAN) the lower half of the CTA
            // integrates translations, the upper half rotations; otherwise one thread does both halves of its voxel
            const int half = nt >> 1;
            const bool split = LEAN && nOwn <= half;
            const bool rotRole = split && tid >= half;
            const int myVox = rotRole ? tid - half : tid;
            const uint32_t *sStat = (const uint32_t *)(st + L.stat);

            // The bulk copies of brick k+1 (the bytes: its own runs, link history, descriptors, static words, and brick k+2's gather list)
            // start NOW, a whole brick ahead - one thread, a dozen instructions; the other stage's last readers passed the barrier that
            // ended brick k-1.  Its halo gathers (hundreds of instructions) still wait for the end of the link pass, below.
            if (haveNext && tid == nt - 128) {
                const bool haveNext2 = k + 2 < cnt && sHdr[k + 2].world == curWorld;
                fill(sHdr[k + 1], s ^ 1, nullptr, true, haveNext2 ? &sHdr[k + 2] : nullptr, k & 1, 0, 128, 1);
            }
            // ---- phase 1: links (recomputed ones first, then the owned run) ---------------------------------------------------
            bool diverged = false;
#pragma unroll 1 // (unroll 2 - two links of a thread in flight - was measured: 235 us per step against 156, instruction-cache bound)
            for (int r = tid; r < nLinks; r += nt) {
                const LinkDesc kd = sDesc[r];
                const int axis = kd.flags & 3, ia = kd.a, ib = kd.b;
                const int o = r - B.nHaloL; // >= 0: owned link, slot l0 + o
                const T4 s0 = sH0[r], s1 = sH1[r], s2 = sH2[r];
                const int ma = (kd.flags >> 4) & 15, mb = (kd.flags >> 8) & 15;
                LinkForces<R> f;
                f.fNeg = f.fPos = f.mNeg = f.mPos = V3<R>{R(0), R(0), R(0)};
                f.strain = 0;
                bool skip = false;
                if constexpr (!LEAN)
                    skip = vx3 && sVm[ma].fixed && sVm[mb].fixed; // gpu_update_links skips these
                if (!skip) {
                    const D2 axy = sPxy[ia], azt = sPzt[ia], bxy = sPxy[ib], bzt = sPzt[ib];
                    const T4 oa = sOri[ia], ob = sOri[ib];
                    LinkHist<R> h;
                    h.pos2 = {s0.x, s0.y, s0.z};
                    h.a1v = {s1.x, s1.y, s1.z};
                    h.a2v = {s2.x, s2.y, s2.z};
                    h.flags = hist_flags<R>(s0);
                    h.maxStrain = (float)s1.w;
                    const DLinkMat &lmat = sLmat[kd.lmat];
                    const LinkLen<R> &len = sLen[kd.lmat];
                    if (axis == 0)
                        brick_link<R, LEAN, 0>(lmat, len, sVm, sDamp, ma, mb, voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
                    else if (axis == 1)
                        brick_link<R, LEAN, 1>(lmat, len, sVm, sDamp, ma, mb, voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
                    else
                        brick_link<R, LEAN, 2>(lmat, len, sVm, sDamp, ma, mb, voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
                    if (o >= 0) {
                        dst.lpos2[B.l0 + o] = mk4<R>(h.pos2.x, h.pos2.y, h.pos2.z, hist_flags_pack<R>(h.flags));
                        dst.la1[B.l0 + o] = mk4<R>(h.a1v.x, h.a1v.y, h.a1v.z, (R)h.maxStrain);
                        dst.la2[B.l0 + o] = mk4<R>(h.a2v.x, h.a2v.y, h.a2v.z, (R)f.strain);
                        if (f.strain > 100.0f)
                            diverged = true;
                    }
                } else if (o >= 0) { // carry the untouched history over to the other buffer
                    dst.lpos2[B.l0 + o] = s0;
                    dst.la1[B.l0 + o] = s1;
                    dst.la2[B.l0 + o] = s2;
                }
                if (kd.flags & BD_NEG_OWN) { // the negative end takes the link in its positive-direction slot 2*axis
                    R *p = sF + ia * BRICK_FSTRIDE + 12 * axis;
                    p[0] = f.fNeg.x, p[1] = f.fNeg.y, p[2] = f.fNeg.z;
                    p[3] = f.mNeg.x, p[4] = f.mNeg.y, p[5] = f.mNeg.z;
                }
                if (kd.flags & BD_POS_OWN) { // the positive end in its negative-direction slot 2*axis+1
                    R *p = sF + ib * BRICK_FSTRIDE + 12 * axis + 6;
                    p[0] = f.fPos.x, p[1] = f.fPos.y, p[2] = f.fPos.z;
                    p[3] = f.mPos.x, p[4] = f.mPos.y, p[5] = f.mPos.z;
                }
            }
            if (diverged)
                Wp->divergedNow = 1;
            // Prefetch brick k+1 into the other stage (its last readers passed the barrier that ended brick k-1).  The last four warps
            // do it, after their links: they hold the highest link numbers, so they are the ones with an iteration to spare.
            // (Issuing it before the links instead costs 20 % - measured, profiles/r01m: it takes issue slots while all 16 warps
            // still have link work.  Round 2 re-measured it with the gathers spread over ALL threads at the start of the link pass: 166 us
            // per step against 156.)
            if (haveNext && tid >= nt - 128) {
                const bool haveNext2 = k + 2 < cnt && sHdr[k + 2].world == curWorld;
                fill(sHdr[k + 1], s ^ 1, (const int32_t *)(smem + L.glist0 + (uint32_t)((k + 1) & 1) * L.glistBytes), true, haveNext2 ? &sHdr[k + 2] : nullptr, k & 1, tid - (nt - 128), 128, 2);
            }
            __syncthreads();

            // ---- phase 2: own voxels -----------------------------------------------------------------------------------------------
            WorldStepParams wp;
            wp.dt = dt;
            wp.floorEnabled = (wflags & WF_FLOOR) != 0;
            wp.vx3 = vx3;
            wp.forceField = nullptr;
#pragma unroll 1
            for (int i = myVox; i < nOwn; i += (split ? nOwn : nt)) {
                const uint32_t stw = sStat[i];
                const DVoxMat &m = sVm[stw & 255];
                const D2 pzt = sPzt[i];
                const T4 ori = sOri[i];
                const unsigned long long packed = (unsigned long long)__double_as_longlong(pzt.y);
                VoxState<R> v;
                v.temp = u2f((uint32_t)(packed & 0xffffffffu));
                v.flags = (uint32_t)(packed >> 32);
                v.orient = {ori.x, ori.y, ori.z, ori.w};
                const uint32_t mk = (stw >> 8) & 63u;
                const bool allSix = __all_sync(__activemask(), mk == 63u); // interior warp: sum in the reference's direction order, no tests
                const R *p = sF + i * BRICK_FSTRIDE;
                const int vs = B.v0 + i;
                auto sum_slots = [&](int off) {
                    V3<R> a{R(0), R(0), R(0)};
                    if (allSix) {
                        a = V3<R>{p[off], p[off + 1], p[off + 2]};
#pragma unroll
                        for (int q = 1; q < 6; q++)
                            a = a + V3<R>{p[6 * q + off], p[6 * q + off + 1], p[6 * q + off + 2]};
                    } else {
#pragma unroll
                        for (int q = 0; q < 6; q++)
                            if (mk & (1u << q))
                                a = a + V3<R>{p[6 * q + off], p[6 * q + off + 1], p[6 * q + off + 2]};
                    }
                    return a;
                };
                const bool ghost = stw & (1u << 15); // a voxel another GPU's slab owns: carried through, refreshed by the halo import
                if constexpr (PEER)
                    if (stw & BS_REMOTE_FED)
                        continue; // a ghost the owning slab's kernel writes straight into my other buffer set
                // PEER: where the neighbours keep this voxel as a ghost, in the buffer set they write this step (the same parity as mine)
                [[maybe_unused]] const bool dpar1 = PEER && dst.pxy == BA.set[1].pxy;
                if (split) { // LEAN only
                    if (ghost) {
                        if (!rotRole) {
                            dst.pxy[vs] = sPxy[i];
                            dst.pzt[vs] = pzt;
                            dst.lm[vs] = sLm[i];
                        } else {
                            dst.orient[vs] = ori;
                            dst.am[vs] = sAm[i];
                        }
                        continue;
                    }
                    if (!rotRole) {
                        const D2 pxy = sPxy[i];
                        const T4 lm = sLm[i];
                        v.pos = {pxy.x, pxy.y, pzt.x};
                        v.linMom = {lm.x, lm.y, lm.z};
                        const V3<R> zero{R(0), R(0), R(0)};
                        V3<R> contact = zero;
                        [[maybe_unused]] int ss = -1;
                        if constexpr (COLL)
                            if (wflags & WF_COLLISION) {
                                ss = BA.bsurf[vs];
                                if (ss >= 0) {
                                    const auto c = A.contactF[ss];
                                    contact = {c.x, c.y, c.z};
                                }
                            }
                        const bool sf = voxel_translate<R, !COLL>(m, wp, nullptr, v, sum_slots(0), contact, zero); // (PLAIN ignores the contact force)
                        const uint32_t fl = (sf ? VF_STATIC_FRICTION : 0u) | (voxel_zero_angmom(m, wp, sf, v.temp, v.pos.z) ? VF_ZERO_ANGMOM : 0u);
                        dst.pxy[vs] = D2{v.pos.x, v.pos.y};
                        const unsigned long long b2 = ((unsigned long long)fl << 32) | (unsigned long long)f2u(v.temp);
                        dst.pzt[vs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                        dst.lm[vs] = mk4<R>(v.linMom.x, v.linMom.y, v.linMom.z, R(0));
                        if constexpr (COLL)
                            if (ss >= 0) { // what the contact kernels of the next step read, and: are the candidate lists stale?
                                const int g = A.surf[ss];
                                A.pxy[g] = D2{v.pos.x, v.pos.y};
                                A.pzt[g] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                                A.lm[g] = mk4<R>(v.linMom.x, v.linMom.y, v.linMom.z, R(0));
                                if (gridCheck)
                                    cgrid_check_one(*Wp, G, ss, v.pos.x, v.pos.y, v.pos.z, refX, refY, refZ);
                            }
                        if constexpr (PEER) {
                            if (stw & BS_SEND_LO) {
                                const int rs = P.sendLo[vs];
                                (dpar1 ? P.loPxy1 : P.loPxy0)[rs] = D2{v.pos.x, v.pos.y};
                                (dpar1 ? P.loPzt1 : P.loPzt0)[rs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                            }
                            if (stw & BS_SEND_HI) {
                                const int rs = P.sendHi[vs];
                                (dpar1 ? P.hiPxy1 : P.hiPxy0)[rs] = D2{v.pos.x, v.pos.y};
                                (dpar1 ? P.hiPzt1 : P.hiPzt0)[rs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                            }
                        }
                        if (sampling) {
                            const double mass = m.mass;
                            world_accumulate(Wp->comAcc, B.world, m.isMeasured != 0 && !ghost, v.pos.x * mass, v.pos.y * mass, v.pos.z * mass, mass);
                        }
                    } else {
                        const T4 am = sAm[i];
                        v.angMom = (v.flags & VF_ZERO_ANGMOM) ? V3<R>{R(0), R(0), R(0)} : V3<R>{am.x, am.y, am.z};
                        voxel_rotate<R, true>(m, wp, nullptr, v, sum_slots(3));
                        dst.orient[vs] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                        dst.am[vs] = mk4<R>(v.angMom.x, v.angMom.y, v.angMom.z, R(0));
                        if constexpr (PEER) {
                            if (stw & BS_SEND_LO)
                                (dpar1 ? P.loOri1 : P.loOri0)[P.sendLo[vs]] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                            if (stw & BS_SEND_HI)
                                (dpar1 ? P.hiOri1 : P.hiOri0)[P.sendHi[vs]] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                        }
                    }
                    continue;
                }
                const D2 pxy = sPxy[i];
                const T4 lm = sLm[i], am = sAm[i];
                v.pos = {pxy.x, pxy.y, pzt.x};
                v.linMom = {lm.x, lm.y, lm.z};
                v.angMom = (v.flags & VF_ZERO_ANGMOM) ? V3<R>{R(0), R(0), R(0)} : V3<R>{am.x, am.y, am.z};
                v.flags &= VF_STATIC_FRICTION;
                bool fixedMat = ghost;
                [[maybe_unused]] int ssG = -1;
                if constexpr (!LEAN)
                    fixedMat = ghost || (vx3 && m.fixed);
                if (!fixedMat) {
                    const V3<R> F = sum_slots(0), M = sum_slots(3);
                    V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
                    if constexpr (COLL)
                        if (wflags & WF_COLLISION) {
                            ssG = BA.bsurf[vs];
                            if (ssG >= 0) {
                                const auto c = A.contactF[ssG];
                                contact = {c.x, c.y, c.z};
                            }
                        }
                    VoxExt ext;
                    const VoxExt *pe = nullptr;
                    V3<double> nominal{0, 0, 0};
                    if constexpr (!LEAN) {
                        const int g = BA.borig[B.v0 + i];
                        if ((wflags & WF_CILIA) && m.cilia != 0 && A.baseCilia && A.vsurf[g] >= 0) {
                            D4 b = A.baseCilia[g];
                            cilia = qrot(v.orient, V3<R>{(R)b.x, (R)b.y, (R)b.z});
                        }
                        if (stw & (1u << 14)) {
                            const DExt &x = A.ext[A.vext[g]];
                            const short4 lat = A.lat[g];
                            nominal = {lat.x * m.nomSize, lat.y * m.nomSize, lat.z * m.nomSize};
                            ext.dof = x.dof;
                            ext.force = {x.force[0], x.force[1], x.force[2]};
                            ext.moment = {x.moment[0], x.moment[1], x.moment[2]};
                            ext.translation = {x.translation[0], x.translation[1], x.translation[2]};
                            ext.rotationQ = {x.rotQ[0], x.rotQ[1], x.rotQ[2], x.rotQ[3]};
                            pe = &ext;
                        }
                    }
                    voxel_update<R, LEAN && !COLL>(m, wp, pe, nominal, v, F, M, contact, cilia);
                    if constexpr (!LEAN) {
                        const WorldD &W = *Wp;
                        if ((wflags & WF_VARY_TEMP) && W.tempPeriod > 0) { // temperature the next step will see (start-of-step time t+dt)
                            const int g = BA.borig[B.v0 + i];
                            v.temp = voxel_temperature(W.tempAmp, W.tempPeriod, A.phase ? A.phase[g] : 0.0, W.t + (double)dt, (wflags & WF_EXPANSION) != 0);
                        }
                    }
                }
                dst.pxy[vs] = D2{v.pos.x, v.pos.y};
                const unsigned long long b2 = ((unsigned long long)v.flags << 32) | (unsigned long long)f2u(v.temp);
                dst.pzt[vs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                dst.orient[vs] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                dst.lm[vs] = mk4<R>(v.linMom.x, v.linMom.y, v.linMom.z, R(0));
                dst.am[vs] = mk4<R>(v.angMom.x, v.angMom.y, v.angMom.z, R(0));
                if constexpr (COLL)
                    if (ssG >= 0) {
                        const int g = A.surf[ssG];
                        A.pxy[g] = D2{v.pos.x, v.pos.y};
                        A.pzt[g] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                        A.lm[g] = mk4<R>(v.linMom.x, v.linMom.y, v.linMom.z, R(0));
                        if (gridCheck)
                            cgrid_check_one(*Wp, G, ssG, v.pos.x, v.pos.y, v.pos.z, refX, refY, refZ);
                    }
                if constexpr (PEER) {
                    if (stw & BS_SEND_LO) {
                        const int rs = P.sendLo[vs];
                        (dpar1 ? P.loPxy1 : P.loPxy0)[rs] = D2{v.pos.x, v.pos.y};
                        (dpar1 ? P.loPzt1 : P.loPzt0)[rs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                        (dpar1 ? P.loOri1 : P.loOri0)[rs] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                    }
                    if (stw & BS_SEND_HI) {
                        const int rs = P.sendHi[vs];
                        (dpar1 ? P.hiPxy1 : P.hiPxy0)[rs] = D2{v.pos.x, v.pos.y};
                        (dpar1 ? P.hiPzt1 : P.hiPzt0)[rs] = D2{v.pos.z, __longlong_as_double((long long)b2)};
                        (dpar1 ? P.hiOri1 : P.hiOri0)[rs] = mk4<R>(v.orient.w, v.orient.x, v.orient.y, v.orient.z);
                    }
                }
                if (sampling) {
                    const double mass = m.mass;
                    world_accumulate(Wp->comAcc, B.world, m.isMeasured != 0 && !ghost, v.pos.x * mass, v.pos.y * mass, v.pos.z * mass, mass);
                }
            }
            gather_wait_all(); // this thread's gathers for brick k+1 have landed ...
            __syncthreads();   // ... and everybody's have; stage s and the force slots are free again
            s ^= 1;
        }
    }
    // End-of-step bookkeeping of the brick worlds (k_world_step's job) by the last CTA to finish, when the host asked for it:
    // one launch per time step instead of two.
    if (worldStep) {
        __syncthreads();
        if (tid == 0) {
            if constexpr (PEER)
                __threadfence_system(); // this CTA's stores into the neighbours' memory are visible there before the flag can be
            else
                __threadfence();
            if (atomicAdd(BA.ticket, 1) == (int)gridDim.x - 1) {
                *BA.ticket = 0;
                __threadfence();
                brick_world_step<R>(A.worlds, BA.worldList, BA.nWorlds, useStop, activeCount);
                if constexpr (PEER) {
                    const unsigned long long next = P.sync->seq + 1;
                    P.sync->seq = next;
                    __threadfence_system();
                    if (P.loFlag)
                        *(volatile unsigned long long *)P.loFlag = next;
                    if (P.hiFlag)
                        *(volatile unsigned long long *)P.hiFlag = next;
                }
            }
        }
    }
}

// Start of a run: the API-order state of every brick world -> buffer set 0 in brick order (W.parity is 0 outside a run).
template <class R> __global__ void k_brick_pack(Arrays<R> A, BrickArrays<R> BA) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < BA.nvSlots) {
        const int g = BA.borig[i];
        if (g >= 0 && (A.worlds[A.vworld[g]].flags & WF_BRICK)) {
            const auto lm = A.lm[g];
            const auto am = A.am[g];
            BA.set[0].pxy[i] = A.pxy[g];
            BA.set[0].pzt[i] = A.pzt[g];
            BA.set[0].orient[i] = A.orient[g];
            BA.set[0].lm[i] = mk4<R>(lm.x, lm.y, lm.z, R(0));
            BA.set[0].am[i] = mk4<R>(lm.w, am.x, am.y, R(0));
        }
    }
    if (i < BA.nlSlots) {
        const int g = BA.lorig[i];
        if (g >= 0 && (A.worlds[A.lworld[g]].flags & WF_BRICK)) {
            BA.set[0].lpos2[i] = A.lpos2[g];
            BA.set[0].la1[i] = A.la1[g];
            BA.set[0].la2[i] = A.la2[g];
        }
    }
}
// End of a run: current buffer set -> API-order arrays (read-back, reductions and every other kernel look only there).
template <class R> __global__ void k_brick_unpack(Arrays<R> A, BrickArrays<R> BA) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < BA.nvSlots) {
        const int g = BA.borig[i];
        if (g >= 0) {
            const WorldD &W = A.worlds[A.vworld[g]];
            if (W.flags & WF_BRICK) {
                const BrickSet<R> &s = BA.set[W.parity];
                D2 pzt = s.pzt[i];
                unsigned long long packed = (unsigned long long)__double_as_longlong(pzt.y);
                auto lm = s.lm[i];
                auto am = s.am[i];
                if ((packed >> 32) & VF_ZERO_ANGMOM) { // the pending VX3 zeroing (see BrickSet)
                    am.x = am.y = am.z = R(0);
                    packed &= ~((unsigned long long)VF_ZERO_ANGMOM << 32);
                    pzt.y = __longlong_as_double((long long)packed);
                }
                A.pxy[g] = s.pxy[i];
                A.pzt[g] = pzt;
                A.orient[g] = s.orient[i];
                A.lm[g] = mk4<R>(lm.x, lm.y, lm.z, am.x);
                A.am[g] = mk2<R>(am.y, am.z);
            }
        }
    }
    if (i < BA.nlSlots) {
        const int g = BA.lorig[i];
        if (g >= 0) {
            const WorldD &W = A.worlds[A.lworld[g]];
            if (W.flags & WF_BRICK) {
                const BrickSet<R> &s = BA.set[W.parity];
                A.lpos2[g] = s.lpos2[i];
                A.la1[g] = s.la1[i];
                A.la2[g] = s.la2[i];
            }
        }
    }
}

// ---- halo exchange of a body cut into slabs (include/vx3b200.h section 2b) -------------------------------------------------
// One record per listed voxel: what the link pass of the neighbouring slab reads from it.
template <class R> struct HaloRec {
    double x, y, z;
    unsigned long long packed; // temperature | flags, as stored in pzt
    typename Store<R>::T4 orient;
};
// resident: the state lives in the brick-ordered sets (between asynchronous runs), else in the API-order arrays
template <class R> __global__ void k_halo_export(Arrays<R> A, BrickArrays<R> BA, int world, const int32_t *vox, const int32_t *slots, int n, int resident, HaloRec<R> *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    D2 a, b;
    typename Store<R>::T4 o;
    if (resident) {
        const BrickSet<R> &s = BA.set[A.worlds[world].parity];
        a = s.pxy[slots[i]], b = s.pzt[slots[i]], o = s.orient[slots[i]];
    } else
        a = A.pxy[vox[i]], b = A.pzt[vox[i]], o = A.orient[vox[i]];
    HaloRec<R> r;
    r.x = a.x, r.y = a.y, r.z = b.x;
    r.packed = (unsigned long long)__double_as_longlong(b.y);
    r.orient = o;
    out[i] = r;
}
template <class R> __global__ void k_halo_import(Arrays<R> A, BrickArrays<R> BA, int world, const int32_t *vox, const int32_t *slots, int n, int resident, const HaloRec<R> *in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const HaloRec<R> r = in[i];
    const D2 a{r.x, r.y}, b{r.z, __longlong_as_double((long long)r.packed)};
    if (resident) {
        const BrickSet<R> &s = BA.set[A.worlds[world].parity];
        s.pxy[slots[i]] = a, s.pzt[slots[i]] = b, s.orient[slots[i]] = r.orient;
    } else
        A.pxy[vox[i]] = a, A.pzt[vox[i]] = b, A.orient[vox[i]] = r.orient;
}

// ---- device-side exchange: no host in the loop ----------------------------------------------------------------------------
// The neighbouring slab's receive area (two record buffers + a sequence flag, in ITS memory) is mapped into this process (CUDA IPC
// over NVLink, or a plain pointer when both slabs live in one process).  k_halo_push writes this step's records straight into the
// buffer of the step's parity and then raises the flag to the step's sequence number; k_halo_pull on the other side waits for that
// number and unpacks.  Two buffers suffice: a sender can be at most one step ahead of the slab it feeds (its step s + 2 needs that
// slab's step s + 1, which is enqueued after that slab's pull of step s).
struct HaloSync {
    unsigned long long seq;    // exchanges completed on this side of the link
    unsigned int ticket;       // CTAs of the running kernel that have finished
    unsigned int timeouts;     // pulls that gave up waiting (the run is invalid; reported by the host)
    unsigned long long waitNs; // time the first CTA of a pull spent waiting for the neighbour's flag, summed over the steps (diagnostic)
};
// set-up of the fused exchange: flag bits into the static words of the listed slots, and (send lists) my slot -> the neighbour's slot
static __global__ void k_brick_mark(uint32_t *bstat, const int32_t *slots, int n, uint32_t bit, int32_t *sendSlot, const int32_t *remoteSlots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    atomicOr(&bstat[slots[i]], bit);
    if (sendSlot)
        sendSlot[slots[i]] = remoteSlots[i];
}

// One side (lower / upper neighbour) of an exchange kernel; blockIdx.y selects it, so a slab's two pushes are ONE launch and so are its pulls.
template <class R> struct HaloSide {
    const int32_t *slots;       // brick-order slot of every listed voxel
    int n;
    HaloRec<R> *buf0, *buf1;    // push: the neighbour's two record buffers (mapped); pull: my own
    unsigned long long *flag;   // push: the neighbour's flag; pull: mine
    HaloSync *sync;
};
constexpr int HALO_THREADS = 128;
// Records are staged in shared memory and leave as consecutive 16-byte pieces, so a warp's store covers 512 contiguous bytes of the
// neighbour's buffer: stores of single 48-byte records at a 48-byte stride went over NVLink as 16-byte packets and cost 270 us per
// time step on the 160,000-voxel faces of a 400^3 body's slabs.
template <class R> __global__ void __launch_bounds__(HALO_THREADS) k_halo_push(Arrays<R> A, BrickArrays<R> BA, int world, HaloSide<R> lo, HaloSide<R> hi) {
    constexpr int P = (int)(sizeof(HaloRec<R>) / 16);
    static_assert(sizeof(HaloRec<R>) % 16 == 0, "halo record is copied in 16-byte pieces");
    __shared__ uint4 stage[HALO_THREADS * P];
    const HaloSide<R> &S = blockIdx.y ? hi : lo;
    const unsigned long long next = S.sync->seq + 1;
    const int base = blockIdx.x * HALO_THREADS, i = base + threadIdx.x;
    if (i < S.n) {
        const BrickSet<R> &s = BA.set[A.worlds[world].parity];
        const D2 a = s.pxy[S.slots[i]], b = s.pzt[S.slots[i]];
        HaloRec<R> r;
        r.x = a.x, r.y = a.y, r.z = b.x;
        r.packed = (unsigned long long)__double_as_longlong(b.y);
        r.orient = s.orient[S.slots[i]];
        const uint4 *r4 = (const uint4 *)&r;
#pragma unroll
        for (int k = 0; k < P; k++)
            stage[threadIdx.x * P + k] = r4[k];
    }
    __syncthreads();
    const int cnt = min(HALO_THREADS, S.n - base);
    if (cnt > 0) {
        uint4 *dst = (uint4 *)(((next & 1) ? S.buf1 : S.buf0) + base);
        for (int j = threadIdx.x; j < cnt * P; j += HALO_THREADS)
            dst[j] = stage[j];
    }
    // ONE system-scope fence per CTA, by the thread that then takes the ticket (fences are cumulative over the barrier): a fence by
    // each of the 128 threads made a slab's push cost tens of microseconds once the stores went over NVLink
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system(); // the records are visible to the peer before the flag can be
        if (atomicAdd(&S.sync->ticket, 1u) == gridDim.x - 1) { // the last CTA of this side to finish
            S.sync->ticket = 0;
            S.sync->seq = next;
            __threadfence_system();
            *(volatile unsigned long long *)S.flag = next;
        }
    }
}
template <class R> __global__ void __launch_bounds__(HALO_THREADS) k_halo_pull(Arrays<R> A, BrickArrays<R> BA, int world, HaloSide<R> lo, HaloSide<R> hi) {
    constexpr int P = (int)(sizeof(HaloRec<R>) / 16);
    __shared__ uint4 stage[HALO_THREADS * P];
    __shared__ int ok;
    const HaloSide<R> &S = blockIdx.y ? hi : lo;
    const unsigned long long want = S.sync->seq + 1;
    const int base = blockIdx.x * HALO_THREADS, i = base + threadIdx.x;
    const int cnt = min(HALO_THREADS, S.n - base);
    if (threadIdx.x == 0) {
        ok = 1;
        if (cnt > 0) {
            const unsigned long long t0 = halo_now_ns();
            while (S.sync->timeouts == 0 && *(const volatile unsigned long long *)S.flag < want) { // (one timeout ends the waiting for good)
                __nanosleep(100);
                if (halo_now_ns() - t0 > 10000000000ull) { // 10 s: the neighbour is gone; never hang the device
                    ok = 0;
                    break;
                }
            }
            if (blockIdx.x == 0)
                S.sync->waitNs += halo_now_ns() - t0;
            __threadfence_system();
        }
    }
    __syncthreads();
    const int good = ok;
    if (good && cnt > 0) {
        const uint4 *src = (const uint4 *)(((want & 1) ? S.buf1 : S.buf0) + base); // (cache-volatile loads: the lines were written by another device)
        for (int j = threadIdx.x; j < cnt * P; j += HALO_THREADS)
            stage[j] = __ldcv(src + j);
    }
    __syncthreads();
    if (good && i < S.n) {
        HaloRec<R> r;
        uint4 *r4 = (uint4 *)&r;
#pragma unroll
        for (int k = 0; k < P; k++)
            r4[k] = stage[threadIdx.x * P + k];
        const BrickSet<R> &s = BA.set[A.worlds[world].parity];
        s.pxy[S.slots[i]] = D2{r.x, r.y}, s.pzt[S.slots[i]] = D2{r.z, __longlong_as_double((long long)r.packed)}, s.orient[S.slots[i]] = r.orient;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!good)
            atomicAdd(&S.sync->timeouts, 1u);
        if (atomicAdd(&S.sync->ticket, 1u) == gridDim.x - 1) {
            S.sync->ticket = 0;
            S.sync->seq = want;
        }
    }
}

} // namespace vx3b
