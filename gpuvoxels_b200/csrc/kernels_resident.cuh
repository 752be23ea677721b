// Resident path: one CTA per world, the whole world state lives in shared memory for the entire run
// (all steps in ONE launch), so a generation of small xenobot-scale worlds touches HBM only at the
// start and at the end.  Worlds are independent, so there is no inter-CTA communication at all.
//
// Replaces, for worlds that fit, the reference's one-thread-per-world CUDA_Simulation driver with its
// 3-6 child launches + device-wide syncs per step (VX3/VX3_SimulationManager.cu:11-121,
// VX3/VX3_VoxelyzeKernel.cu:237-359).
//
// Contact (reference: all Ns^2/2 surface pairs every step, VX3_VoxelyzeKernel.cu:833-843) is a uniform-grid
// broad phase kept in shared memory plus per-voxel candidate lists with a skin (Verlet lists):
//   rebuild (only when some surface voxel has moved more than skin/2, relative to the world's common drift, since the last
//     rebuild): surface voxels are counting-sorted by hashed cell (edge = largest watch distance + skin); each surface voxel
//     scans the 27 cells around it and keeps every non-linked voxel within (watch distance + skin) in its list;
//   every step: each surface voxel tests only its list with the reference's exact double-precision predicate.
// Any pair the reference would watch at the current step is in the lists (triangle inequality), so the watched-pair SET
// is exactly the reference's; lists are filled in (cell, index) order, which keeps the force sums deterministic.
#pragma once
#include "engine.cuh"
#include "kernels_stream.cuh"

namespace vx3b {


// shared-memory layout for capacities (capV, capL, capS); every array is 16-byte aligned
template <class R> struct ResidentLayout {
    size_t world, vmats, lmats;
    size_t pos, orient, linMom, angMom, temp, vflags, vmat, vlinks, lat, phase, bsa, cilia, vext, vsurfslot;
    size_t lpos2, la1, la2, lflags, lmaxs, lends, lmeta, lfN, lfP, lmN, lmP;
    size_t surf, contactF, skey, sorted, cnt, rec4, rec2, nbr, nbrCnt, ref;
    size_t total;
    int H; // hash buckets (power of two)
    __host__ __device__ static size_t up(size_t x) { return (x + 15) & ~(size_t)15; }
    // spill: the four per-link force/moment arrays stay in global memory (A.lfN ...; L2-resident) so that worlds whose
    // full state exceeds one SM's shared memory can still run in k_resident
    __host__ __device__ ResidentLayout(int capV, int capL, int capS, int capVM, int capLM, uint32_t wflags, bool spill, int listCap) {
        const bool collide = (wflags & WF_COLLISION) && capS > 1;
        size_t o = 0;
        auto take = [&](size_t bytes) {
            size_t at = o;
            o = up(o + bytes);
            return at;
        };
        world = take(offsetof(WorldD, fitnessF)); // the kernel needs the record up to and including the stop-condition program
        vmats = take(sizeof(DVoxMat) * capVM);
        lmats = take(sizeof(DLinkMat) * capLM);
        pos = take(sizeof(double) * 3 * capV);
        orient = take(sizeof(R) * 4 * capV);
        linMom = take(sizeof(R) * 3 * capV);
        angMom = take(sizeof(R) * 3 * capV);
        temp = take((wflags & WF_THERMAL) ? sizeof(float) * capV : 0); // temperature is unobservable otherwise
        vflags = take(capV);
        vmat = take(capV);
        vlinks = take(sizeof(uint16_t) * 6 * capV);
        lat = take(sizeof(short4) * capV);
        // thermal worlds: per voxel the phase offset (validation build) or sin / cos of K * phase (float build: voxel_temperature_split),
        // and the cached baseSizeAverage at the voxel's current temperature
        phase = take((wflags & WF_THERMAL) ? sizeof(double) * (sizeof(R) == 4 ? 2 : 1) * capV : 0);
        bsa = take((wflags & WF_THERMAL) ? sizeof(double) * capV : 0);
        cilia = take((wflags & WF_CILIA) ? sizeof(float) * 3 * capV : 0);
        vext = take(sizeof(int32_t) * capV);
        vsurfslot = take(sizeof(uint16_t) * capV);
        lpos2 = take(sizeof(R) * 3 * capL);
        la1 = take(sizeof(R) * 3 * capL);
        la2 = take(sizeof(R) * 3 * capL);
        lflags = take(capL);
        lmaxs = take((wflags & WF_CAN_FAIL) ? sizeof(float) * capL : 0);
        lends = take(sizeof(uint16_t) * 2 * capL);
        lmeta = take(sizeof(uint16_t) * capL);
        surf = take(sizeof(uint16_t) * capS);
        H = 64;
        while (H < 2 * capS && H < 8192)
            H *= 2;
        // the four per-link force arrays and the grid scratch of the contact rebuild share one region: the scratch is only
        // alive between the start of a step and its link pass, when the forces of the previous step are dead
        const size_t forceBytes = spill ? 0 : 4 * up(sizeof(R) * 3 * capL);
        // scratch = cell keys, sorted order, bucket counters, and the cell-sorted records the list build streams through
        // (float4 position relative to the world's reference voxel | packed lattice index, and {cell key, surface index})
        const size_t scratchBytes = collide ? up(sizeof(uint32_t) * capS) + up(sizeof(uint16_t) * capS) + up(sizeof(int32_t) * H) +
                                                  up(sizeof(float4) * capS) + up(sizeof(uint2) * capS)
                                            : 0;
        const size_t shared = take(forceBytes > scratchBytes ? forceBytes : scratchBytes);
        lfN = shared;
        lfP = lfN + up(sizeof(R) * 3 * capL);
        lmN = lfP + up(sizeof(R) * 3 * capL);
        lmP = lmN + up(sizeof(R) * 3 * capL);
        skey = shared;
        sorted = skey + up(sizeof(uint32_t) * capS);
        cnt = sorted + up(sizeof(uint16_t) * capS);
        rec4 = cnt + up(sizeof(int32_t) * H);
        rec2 = rec4 + up(sizeof(float4) * capS);
        if (collide) {
            contactF = take(sizeof(R) * 3 * capS);
            nbr = take(sizeof(uint16_t) * listCap * capS);
            nbrCnt = take(capS);
            ref = take(sizeof(float) * 3 * capS);
        } else
            contactF = nbr = nbrCnt = ref = o;
        total = o;
    }
};

struct ResidentShared { // block-wide scalars
    double sinA, cosA; // float build, thermal worlds: sin / cos of K * (t + dt) / period of the current step (thread 0, before the link pass)
    int32_t diverged, running, rebuild, overflow;
    int32_t narrowNext, padN[3]; // next unclaimed surface voxel of the narrow phase (reset during the voxel pass)
    // target-hit counter of even / odd steps, each in its own 16-byte granule: the compiler fetches neighbouring scalars with one
    // vector load, and the clear of one slot must not overlap the (vector) read of anything else
    int32_t hits, padA[3];
    int32_t hits1, padB[3];
    int32_t warpSums[32];
    double com[4];
};

// Threads of one SM's k_resident CTAs together (register budget: 65536 / threads per thread).  The lean float instantiation needs 104
// registers when left alone; measured on the 1,024-world generation (profiles/r02i): 512 threads 5.59e9 voxel-steps/s, 608 5.75e9,
// 768 (80 registers, 88 bytes of spills) 6.08e9, 896 5.95e9, 1,024 (64 registers) 5.66e9.  The full and the fp64 instantiations
// want 128 and more and stay at 512.
#ifndef RES_MAXT
#define RES_MAXT 512
#endif
#ifndef RES_MAXT_LEAN
#define RES_MAXT_LEAN 768
#endif
template <class R, bool FULL> constexpr int res_maxt() { return (sizeof(R) == 4 && !FULL) ? RES_MAXT_LEAN : RES_MAXT; }
// -DRES_PROFILE: per-phase clock64() accounting of k_resident (warp-cycles, summed over all warps and steps; scripts/gpu_resident_phases.py).
// 0 rebuild (with its barriers), 1 link pass, 2 narrow phase, 3 wait at the barrier after them, 4 voxel pass, 5 bookkeeping + barrier(s)
// at the end of the step, 8 warp-steps counted, 9 rebuilds (warp-rebuilds)
// -DRES_MARKERS: pmevent instructions (PMTRIG in SASS, free at run time) at the phase boundaries, so that the per-instruction
// counters of an `ncu --import-source` capture can be summed per phase by address range (scripts/ncu_phases.py)
#ifdef RES_MARKERS
#define RM(n) asm volatile("pmevent " #n ";");
#else
#define RM(n)
#endif
#ifdef RES_PROFILE
static __device__ unsigned long long g_resProf[16];
#define RP_DECL long long rp[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, rpT = clock64(), rpT2 = 0, rpSteps = 0, rpRebuilds = 0;
// sub-phases of the rebuild (slots 6, 7, 12, 13, 14: keys + counts, scan, scatter + bucket sort, records, lists), timed separately
#define RP2_START rpT2 = clock64();
#define RP2(i)                                                                                                                  \
    {                                                                                                                           \
        const long long rpN = clock64();                                                                                        \
        rp[i] += rpN - rpT2;                                                                                                    \
        rpT2 = rpN;                                                                                                             \
    }
#define RP(i)                                                                                                                   \
    {                                                                                                                           \
        const long long rpN = clock64();                                                                                        \
        rp[i] += rpN - rpT;                                                                                                     \
        rpT = rpN;                                                                                                              \
    }
#else
#define RP_DECL
#define RP(i)
#define RP2_START
#define RP2(i)
#endif
VXD uint32_t cell_hash(uint32_t key, int log2H) { return (key * 2654435761u) >> (32 - log2H); }

// worldList: the worlds of one size class (every CTA lays its shared memory out for its own world's counts; the class only
// fixes the dynamic shared-memory size, i.e. how many CTAs share an SM).
// FULL = false: no world of the launch has boundary conditions, a force field or a failure strain - their code (the boundary-condition
// branches of the voxel update, three formula evaluations per voxel, the failure tests of the link) is compiled away, which keeps
// the generation kernel's hot loops small (instruction cache) and its register budget for the arithmetic.
template <class R, bool SPILL, bool FULL>
__global__ void __launch_bounds__((res_maxt<R, FULL>()), 1) k_resident(Arrays<R> A, const int32_t *worldList, long long maxSteps, int useStop, int recordPairs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ ResidentShared sh;
    const int w = worldList[blockIdx.x];
    WorldD &Wg = A.worlds[w];
    if (!Wg.active)
        return;
    const ResidentLayout<R> L(Wg.nv, max(Wg.nl, 1), max(Wg.ns, 1), max(Wg.nvm, 1), max(Wg.nlm, 1), Wg.flags, SPILL, Wg.listCap);
    WorldD &W = *(WorldD *)(smem + L.world);
    DVoxMat *vmats = (DVoxMat *)(smem + L.vmats);
    DLinkMat *lmats = (DLinkMat *)(smem + L.lmats);
    double *sPos = (double *)(smem + L.pos);
    R *sOri = (R *)(smem + L.orient);
    R *sLin = (R *)(smem + L.linMom);
    R *sAng = (R *)(smem + L.angMom);
    float *sTemp = (float *)(smem + L.temp);
    uint8_t *sVflags = smem + L.vflags;
    uint8_t *sVmat = smem + L.vmat;
    uint16_t *sVlinks = (uint16_t *)(smem + L.vlinks);
    short4 *sLat = (short4 *)(smem + L.lat);
    double *sPhase = (double *)(smem + L.phase);
    double *sBsa = (double *)(smem + L.bsa);
    float *sCilia = (float *)(smem + L.cilia);
    int32_t *sVext = (int32_t *)(smem + L.vext);
    uint16_t *sVsurf = (uint16_t *)(smem + L.vsurfslot);
    R *sP2 = (R *)(smem + L.lpos2);
    R *sA1 = (R *)(smem + L.la1);
    R *sA2 = (R *)(smem + L.la2);
    uint8_t *sLflags = smem + L.lflags;
    float *sLmax = (float *)(smem + L.lmaxs);
    uint16_t *sLends = (uint16_t *)(smem + L.lends);
    uint16_t *sLmeta = (uint16_t *)(smem + L.lmeta);
    // per-link forces/moments for the voxel pass: shared memory (stride 3) or, spilled, the global arrays (stride 4)
    constexpr int FS = SPILL ? 4 : 3;
    R *sFN = SPILL ? (R *)(A.lfN + Wg.loff) : (R *)(smem + L.lfN);
    R *sFP = SPILL ? (R *)(A.lfP + Wg.loff) : (R *)(smem + L.lfP);
    R *sMN = SPILL ? (R *)(A.lmN + Wg.loff) : (R *)(smem + L.lmN);
    R *sMP = SPILL ? (R *)(A.lmP + Wg.loff) : (R *)(smem + L.lmP);
    uint16_t *sSurf = (uint16_t *)(smem + L.surf);
    R *sCF = (R *)(smem + L.contactF);
    uint32_t *sKey = (uint32_t *)(smem + L.skey);
    uint16_t *sSorted = (uint16_t *)(smem + L.sorted);
    int32_t *sCnt = (int32_t *)(smem + L.cnt);
    float4 *sRec4 = (float4 *)(smem + L.rec4);
    uint2 *sRec2 = (uint2 *)(smem + L.rec2);
    uint16_t *sNbr = (uint16_t *)(smem + L.nbr);
    uint8_t *sNbrCnt = smem + L.nbrCnt;
    float *sRef = (float *)(smem + L.ref);

    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- load the world into shared memory --------------------------------------------------------------
    {
        const uint32_t *src = (const uint32_t *)&Wg;
        uint32_t *dst = (uint32_t *)&W;
        for (int i = tid; i < (int)(offsetof(WorldD, fitnessF) / 4); i += nt)
            dst[i] = src[i];
    }
    __syncthreads();
    const int nv = W.nv, nl = W.nl, ns = W.ns, voff = W.voff, loff = W.loff, soff = W.soff;
    const uint32_t wflags = W.flags;
    const bool vx3 = wflags & WF_VX3, floorOn = wflags & WF_FLOOR, collide = (wflags & WF_COLLISION) && ns > 1;
    const bool ciliaOn = (wflags & WF_CILIA) && A.baseCilia, varyTemp = (wflags & WF_THERMAL) && W.tempPeriod > 0;
    const bool canFail = FULL && (wflags & WF_CAN_FAIL), thermal = wflags & WF_THERMAL;
    const float dt = W.dt;
    const float invDt = 1.0f / dt;
    const double voxSize = W.voxSize, tempAmp = W.tempAmp, tempPeriod = W.tempPeriod, invCell = W.invCell, listR2 = W.listR2;
    const float skinHalf2 = W.skinHalf2;
    const double watchR2 = W.watchR2; // > 0: upper bound of every pair's squared watch distance (uniform constant sizes)
    const int RES_K = W.listCap; // capacity of one surface voxel's candidate list (overflow -> that step scans the grid directly)
    const int cycle = W.cycle;
    const int H = L.H, log2H = 31 - __clz(H);
    for (int i = tid; i < W.nvm * (int)(sizeof(DVoxMat) / 4); i += nt)
        ((uint32_t *)vmats)[i] = ((const uint32_t *)(A.vmats + W.vmatOff))[i];
    for (int i = tid; i < W.nlm * (int)(sizeof(DLinkMat) / 4); i += nt)
        ((uint32_t *)lmats)[i] = ((const uint32_t *)(A.lmats + W.lmatOff))[i];
    for (int v = tid; v < nv; v += nt) {
        const int g = voff + v;
        float temp;
        uint32_t fl;
        V3<double> p = load_pos(A, g, &temp, &fl);
        sPos[3 * v] = p.x;
        sPos[3 * v + 1] = p.y;
        sPos[3 * v + 2] = p.z;
        auto o = A.orient[g];
        auto lm = A.lm[g];
        auto am = A.am[g];
        sOri[4 * v] = o.x;
        sOri[4 * v + 1] = o.y;
        sOri[4 * v + 2] = o.z;
        sOri[4 * v + 3] = o.w;
        sLin[3 * v] = lm.x;
        sLin[3 * v + 1] = lm.y;
        sLin[3 * v + 2] = lm.z;
        sAng[3 * v] = lm.w;
        sAng[3 * v + 1] = am.x;
        sAng[3 * v + 2] = am.y;
        if (thermal)
            sTemp[v] = temp;
        sVflags[v] = (uint8_t)fl;
        sVmat[v] = (uint8_t)(A.vmat[g] - W.vmatOff);
        for (int k = 0; k < 6; k++) {
            int l = A.vlinks[(size_t)k * A.nvTot + g];
            sVlinks[6 * v + k] = l < 0 ? (uint16_t)0xffff : (uint16_t)(l - loff);
        }
        sLat[v] = A.lat[g];
        if (thermal) {
            const double ph = (varyTemp && A.phase) ? A.phase[g] : 0.0;
            if constexpr (sizeof(R) == 4) {
                double sn, cs;
                sincos(W.tempK * ph, &sn, &cs);
                sPhase[2 * v] = sn;
                sPhase[2 * v + 1] = cs;
            } else
                sPhase[v] = ph;
            sBsa[v] = base_size_average(A.vmats[A.vmat[g]], temp);
        }
        if (ciliaOn) {
            D4 b = A.baseCilia[g];
            sCilia[3 * v] = (float)b.x;
            sCilia[3 * v + 1] = (float)b.y;
            sCilia[3 * v + 2] = (float)b.z;
        }
        sVext[v] = A.vext[g];
        int slot = A.vsurf[g];
        sVsurf[v] = slot < 0 ? (uint16_t)0xffff : (uint16_t)(slot - soff);
    }
    for (int l = tid; l < nl; l += nt) {
        const int g = loff + l;
        auto s0 = A.lpos2[g], s1 = A.la1[g], s2 = A.la2[g];
        sP2[3 * l] = s0.x;
        sP2[3 * l + 1] = s0.y;
        sP2[3 * l + 2] = s0.z;
        sA1[3 * l] = s1.x;
        sA1[3 * l + 1] = s1.y;
        sA1[3 * l + 2] = s1.z;
        sA2[3 * l] = s2.x;
        sA2[3 * l + 1] = s2.y;
        sA2[3 * l + 2] = s2.z;
        uint32_t fl;
        if constexpr (sizeof(R) == 4)
            fl = f2u(s0.w);
        else
            fl = (uint32_t)__double_as_longlong(s0.w);
        sLflags[l] = (uint8_t)fl;
        if (canFail)
            sLmax[l] = (float)s1.w;
        int2 e = A.lends[g];
        sLends[2 * l] = (uint16_t)(e.x - voff);
        sLends[2 * l + 1] = (uint16_t)(e.y - voff);
        uint32_t meta = A.lmeta[g];
        sLmeta[l] = (uint16_t)((meta & 3) | (((meta >> 2) - W.lmatOff) << 2));
        for (int k = 0; k < 3; k++)
            sFN[FS * l + k] = sFP[FS * l + k] = sMN[FS * l + k] = sMP[FS * l + k] = R(0);
    }
    for (int s = tid; s < ns; s += nt)
        sSurf[s] = (uint16_t)(A.surf[soff + s] - voff);
    if (collide) {
        for (int s = tid; s < ns; s += nt)
            sCF[3 * s] = sCF[3 * s + 1] = sCF[3 * s + 2] = R(0);
        for (int h = tid; h < H; h += nt)
            sCnt[h] = 0;
    }
    if (tid == 0) {
        sh.diverged = 0;
        sh.hits = sh.hits1 = 0;
        sh.running = 1;
        sh.rebuild = 1;
        sh.overflow = 0;
        sh.narrowNext = 0;
        sh.com[0] = sh.com[1] = sh.com[2] = sh.com[3] = 0;
    }
    __syncthreads();

    // world scalars that change during the run live in registers of every thread (all threads compute them identically)
    double t = W.t;
    long long step = W.step;
    long long left = maxSteps;
    int collisionCount = W.collisionCount;
    bool running = true;
    bool rebuild = true; // contact candidate lists must be rebuilt this step (uniform over the block)
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    RP_DECL

    while (running) {
        // ---- contact step 1 (only when the candidate lists are stale): grid + lists from the positions of the step start.
        // The grid scratch (keys, sorted order, bucket counters) aliases the per-link force arrays, which are dead here.
        bool overflow = false;
        int hits = 0;
        float driftX = 0, driftY = 0, driftZ = 0;
        // one surface voxel against one candidate with the reference's exact predicate (both members of a pair evaluate it in
        // the same orientation, voxel1 = the later surface voxel, so the two forces are bit-identical and no atomics are needed)
        auto narrow = [&](bool direct) {
            const bool rec = recordPairs && left == 1; // parity tests read the pair set of the run's last step
            // Surface voxels are handed out 32 at a time from a block-wide counter: in the link pass the warps past nl % nt have
            // one link fewer than the others, and a list walk costs about as much as a link, so whoever is free first takes
            // more of them (a static i = tid + k*nt split made the slowest warp do 2 links AND a walk).  The result of a
            // surface voxel does not depend on which thread computes it.
            while (true) {
                int base = 0;
                if (lane == 0)
                    base = atomicAdd(&sh.narrowNext, 32);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= ns)
                    break;
                const int i = base + lane;
                if (i >= ns)
                    continue;
                const int v1 = sSurf[i];
                const DVoxMat &m1 = vmats[sVmat[v1]];
                const V3<double> p1{sPos[3 * v1], sPos[3 * v1 + 1], sPos[3 * v1 + 2]};
                const V3<R> l1{sLin[3 * v1], sLin[3 * v1 + 1], sLin[3 * v1 + 2]};
                const double avg1 = thermal ? sBsa[v1] : m1.avgSize0;
                V3<R> acc{R(0), R(0), R(0)};
                // candidate = VOXEL index (the surface list ascends in voxel index, so "the later surface voxel" = the larger index)
                auto pair = [&](int v2) {
                    const V3<double> p2{sPos[3 * v2], sPos[3 * v2 + 1], sPos[3 * v2 + 2]};
                    { // nothing beyond the world's largest possible watch distance matters: reject on the squared distance alone
                        const double ex = p1.x - p2.x, ey = p1.y - p2.y, ez = p1.z - p2.z;
                        if (ex * ex + ey * ey + ez * ez > watchR2)
                            return;
                    }
                    const DVoxMat &m2 = vmats[sVmat[v2]];
                    const V3<R> l2{sLin[3 * v2], sLin[3 * v2 + 1], sLin[3 * v2 + 2]};
                    const double avg2 = thermal ? sBsa[v2] : m2.avgSize0;
                    V3<R> f;
                    if (v1 > v2) {
                        if (contact_pair_sized<R>(m1, avg1, p1, l1, m2, avg2, p2, l2, f)) {
                            acc = acc + f;
                            if ((m1.isTarget != 0) != (m2.isTarget != 0))
                                hits++;
                            if (rec) {
                                int slot = atomicAdd(A.pairCount, 1);
                                if (slot < A.pairCap)
                                    A.pairBuf[slot] = make_int4(w, v1, v2, 0);
                            }
                        }
                    } else if (contact_pair_sized<R>(m2, avg2, p2, l2, m1, avg1, p1, l1, f))
                        acc = acc - f;
                };
                if (!direct) {
                    const int cnt = sNbrCnt[i];
                    for (int k = 0; k < cnt; k++)
                        pair(sNbr[i * RES_K + k]);
                } else { // some list overflowed at this rebuild: scan the grid directly (and rebuild again next step)
                    const short4 c1 = sLat[v1];
                    const uint32_t key = sKey[i];
                    const int cx = key & 1023, cy = (key >> 10) & 1023, cz = key >> 20;
                    for (int n = 0; n < 27; n++) {
                        const int dx = n % 3 - 1, dy = (n / 3) % 3 - 1, dz = n / 9 - 1;
                        const uint32_t k2 = (uint32_t)((cx + dx) & 1023) | ((uint32_t)((cy + dy) & 1023) << 10) | ((uint32_t)((cz + dz) & 1023) << 20);
                        const uint32_t h = cell_hash(k2, log2H);
                        const int beg = h ? sCnt[h - 1] : 0, end = sCnt[h];
                        for (int p = beg; p < end; p++) {
                            const int j = sSorted[p];
                            if (j == i || sKey[j] != k2)
                                continue;
                            const int v2 = sSurf[j];
                            const short4 c2 = sLat[v2];
                            if (abs(c1.x - c2.x) + abs(c1.y - c2.y) + abs(c1.z - c2.z) == 1 || (m1.fixed && vmats[sVmat[v2]].fixed))
                                continue;
                            pair(v2);
                        }
                    }
                }
                sCF[3 * i] = acc.x;
                sCF[3 * i + 1] = acc.y;
                sCF[3 * i + 2] = acc.z;
            }
        };
        RM(1)
        if (collide && rebuild) {
            RP2_START
            for (int h = tid; h < H; h += nt)
                sCnt[h] = 0;
            if (tid == 0)
                sh.overflow = 0;
            __syncthreads();
            // Only now has every thread read sh.rebuild (at the end of the previous step, with no barrier after that read): clearing
            // it before this barrier let a slow warp see 0 where the others saw 1 and skip the barriers of this block
            // (compute-sanitizer racecheck; it showed up as an illegal access once the link pass got faster).
            if (tid == 0)
                sh.rebuild = 0; // the voxel pass may set it again
            for (int i = tid; i < ns; i += nt) {
                const int v = sSurf[i];
                const int cx = (int)floor(sPos[3 * v] * invCell), cy = (int)floor(sPos[3 * v + 1] * invCell), cz = (int)floor(sPos[3 * v + 2] * invCell);
                const uint32_t key = (uint32_t)(cx & 1023) | ((uint32_t)(cy & 1023) << 10) | ((uint32_t)(cz & 1023) << 20);
                sKey[i] = key;
                atomicAdd(&sCnt[cell_hash(key, log2H)], 1);
            }
            __syncthreads();
            RP2(6)
            // ---- B1: exclusive scan of the bucket counts -------------------------------------------------------------
            const int per = (H + nt - 1) / nt, b0 = tid * per;
            int sum = 0;
            for (int k = 0; k < per; k++)
                if (b0 + k < H)
                    sum += sCnt[b0 + k];
            int inc = sum;
            for (int o = 1; o < 32; o <<= 1) {
                int n = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o)
                    inc += n;
            }
            if (lane == 31)
                sh.warpSums[wid] = inc;
            __syncthreads();
            if (tid < 32) {
                int ws = tid < nw ? sh.warpSums[tid] : 0, wi = ws;
                for (int o = 1; o < 32; o <<= 1) {
                    int n = __shfl_up_sync(0xffffffffu, wi, o);
                    if (tid >= o)
                        wi += n;
                }
                sh.warpSums[tid] = wi - ws; // exclusive
            }
            __syncthreads();
            int offset = sh.warpSums[wid] + inc - sum;
            for (int k = 0; k < per; k++)
                if (b0 + k < H) {
                    int c = sCnt[b0 + k];
                    sCnt[b0 + k] = offset;
                    offset += c;
                }
            __syncthreads();
            RP2(7)
            // ---- B2: scatter (after it sCnt[h] = END of bucket h; its start is sCnt[h-1]) --------------------------------
            for (int i = tid; i < ns; i += nt)
                sSorted[atomicAdd(&sCnt[cell_hash(sKey[i], log2H)], 1)] = (uint16_t)i;
            __syncthreads();
            // ---- B3: order every bucket by surface index (the atomics above are not ordered) ---------------------------------
            for (int h = tid; h < H; h += nt) {
                const int beg = h ? sCnt[h - 1] : 0, end = sCnt[h];
                for (int p = beg + 1; p < end; p++) {
                    uint16_t x = sSorted[p];
                    int q = p - 1;
                    while (q >= beg && sSorted[q] > x) {
                        sSorted[q + 1] = sSorted[q];
                        q--;
                    }
                    sSorted[q + 1] = x;
                }
            }
            __syncthreads();
            RP2(12)
            // ---- B3b: cell-sorted records.  The list build below tests ~100 voxels per surface voxel; chasing sorted -> key ->
            // surface -> lattice -> 3 doubles for each of them was a chain of five dependent loads.  One record per sorted slot:
            // {cell key, surface index} (8 B) and {position relative to the world's first surface voxel as float3, lattice index
            // packed 10:10:10 | fixed-material bit} (16 B).  The float distance test below is CONSERVATIVE (radius padded by far
            // more than the rounding of a relative position), so a list is a superset of the exact one; every listed candidate
            // still goes through the reference's exact double-precision predicate in the narrow phase.
            const int vOrg = sSurf[0];
            const double ox = sPos[3 * vOrg], oy = sPos[3 * vOrg + 1], oz = sPos[3 * vOrg + 2];
            for (int p = tid; p < ns; p += nt) {
                const int j = sSorted[p];
                const int v = sSurf[j];
                const short4 c = sLat[v];
                const uint32_t pk = (uint32_t)(c.x & 1023) | ((uint32_t)(c.y & 1023) << 10) | ((uint32_t)(c.z & 1023) << 20) |
                                    (vmats[sVmat[v]].fixed != 0 ? 0x40000000u : 0u);
                sRec4[p] = make_float4((float)(sPos[3 * v] - ox), (float)(sPos[3 * v + 1] - oy), (float)(sPos[3 * v + 2] - oz), __uint_as_float(pk));
                sRec2[p] = make_uint2(sKey[j], (uint32_t)j | ((uint32_t)v << 16)); // surface index | voxel index << 16
            }
            __syncthreads();
            RP2(13)
            // ---- B4: candidate lists, one thread per surface voxel.  A flat loop that either tests the next voxel of the current
            // cell or moves on to the next of the 27 cells, so the lanes of a warp stay busy although their cells hold different
            // numbers of voxels (a nested cell/voxel loop costs the warp the SUM over cells of the longest bucket).
            const float listR2f = (float)(listR2 * 1.002); // padded: relative float positions are good to ~1e-7 of the world's extent
            for (int i = tid; i < ns; i += nt) {
                const int v1 = sSurf[i];
                const double x1d = sPos[3 * v1], y1d = sPos[3 * v1 + 1], z1d = sPos[3 * v1 + 2];
                const float x1 = (float)(x1d - ox), y1 = (float)(y1d - oy), z1 = (float)(z1d - oz);
                const short4 c1 = sLat[v1];
                const int pk1 = (int)((uint32_t)(c1.x & 1023) | ((uint32_t)(c1.y & 1023) << 10) | ((uint32_t)(c1.z & 1023) << 20));
                const uint32_t fixed1 = vmats[sVmat[v1]].fixed != 0 ? 0x40000000u : 0u;
                const uint32_t key = sKey[i];
                const int cx = key & 1023, cy = (key >> 10) & 1023, cz = key >> 20;
                int n = -1, p = 0, end = 0, count = 0;
                uint32_t k2 = 0;
                while (true) {
                    if (p < end) {
                        const uint2 kj = sRec2[p];
                        if (kj.x == k2 && (int)(kj.y & 0xffffu) != i) {
                            const float4 r = sRec4[p];
                            const float dx = r.x - x1, dy = r.y - y1, dz = r.z - z1;
                            const uint32_t pk2 = __float_as_uint(r.w);
                            // directly linked voxels (lattice neighbours; static topology) never collide: is_neighbor(depth 1).
                            // The packed difference is +-1, +-2^10 or +-2^20 exactly for lattice neighbours (lattice indices
                            // in 0..1023: the upload pre-pass keeps other worlds off this kernel); nor do two `fixed` voxels collide
                            const int dl = abs(pk1 - (int)(pk2 & 0x3fffffffu));
                            if (dx * dx + dy * dy + dz * dz <= listR2f && dl != 1 && dl != 1024 && dl != 1048576 && !(fixed1 & pk2)) {
                                if (count < RES_K)
                                    sNbr[i * RES_K + count] = (uint16_t)(kj.y >> 16); // the candidate's voxel index
                                count++;
                            }
                        }
                        p++;
                    } else {
                        if (++n == 27)
                            break;
                        k2 = (uint32_t)((cx + n % 3 - 1) & 1023) | ((uint32_t)((cy + (n / 3) % 3 - 1) & 1023) << 10) |
                             ((uint32_t)((cz + n / 9 - 1) & 1023) << 20);
                        const uint32_t h = cell_hash(k2, log2H);
                        p = h ? sCnt[h - 1] : 0;
                        end = sCnt[h];
                    }
                }
                sNbrCnt[i] = (uint8_t)min(count, RES_K);
                if (count > RES_K)
                    sh.overflow = 1;
                sRef[3 * i] = (float)x1d;
                sRef[3 * i + 1] = (float)y1d;
                sRef[3 * i + 2] = (float)z1d;
            }
            __syncthreads();
            RP2(14)
            overflow = sh.overflow != 0;
            if (tid == 0) {
                W.contactRebuilds++;
                W.contactOverflows += overflow;
                if (overflow)
                    sh.rebuild = 1;
            }
            if (overflow) { // the grid scratch dies with the link pass below, so the direct scan has to come first
                narrow(true);
                __syncthreads();
            }
#ifdef RES_PROFILE
            rpRebuilds++;
#endif
        }
        RP(0)
        RM(2)
        if constexpr (sizeof(R) == 4) {
            // the sine of this step's temperature update (t + dt), taken once: read by the voxel pass, after the barrier below
            if (varyTemp && tid == 0) {
                double sn, cs;
                sincos(W.tempK * ((t + (double)dt) / tempPeriod), &sn, &cs);
                sh.sinA = sn;
                sh.cosA = cs;
            }
        }
        // ---- phase A: link pass, and the contact narrow phase over the candidate lists (independent of each other) ------------
        for (int l = tid; l < nl; l += nt) {
            const int a = sLends[2 * l], b = sLends[2 * l + 1];
            const uint32_t meta = sLmeta[l];
            const int axis = meta & 3;
            const DVoxMat &mn = vmats[sVmat[a]], &mp = vmats[sVmat[b]];
            if (vx3 && mn.fixed && mp.fixed) { // gpu_update_links skips these; their force slots must still be defined
                for (int k = 0; k < 3; k++)
                    sFN[FS * l + k] = sFP[FS * l + k] = sMN[FS * l + k] = sMP[FS * l + k] = R(0);
                continue;
            }
            const DLinkMat &lmat = lmats[meta >> 2];
            V3<double> dpos{sPos[3 * b] - sPos[3 * a], sPos[3 * b + 1] - sPos[3 * a + 1], sPos[3 * b + 2] - sPos[3 * a + 2]};
            Qt<R> qn{sOri[4 * a], sOri[4 * a + 1], sOri[4 * a + 2], sOri[4 * a + 3]};
            Qt<R> qp{sOri[4 * b], sOri[4 * b + 1], sOri[4 * b + 2], sOri[4 * b + 3]};
            double restLen = 0.5 * (base_size(mn, thermal ? sTemp[a] : 0.0f) + base_size(mp, thermal ? sTemp[b] : 0.0f));
            LinkHist<R> h;
            h.pos2 = {sP2[3 * l], sP2[3 * l + 1], sP2[3 * l + 2]};
            h.a1v = {sA1[3 * l], sA1[3 * l + 1], sA1[3 * l + 2]};
            h.a2v = {sA2[3 * l], sA2[3 * l + 1], sA2[3 * l + 2]};
            h.flags = sLflags[l];
            h.maxStrain = canFail ? sLmax[l] : -1e30f;
            // dampingMultiplier 2*sqrt(m)*zeta / previousDt (VX3_Voxel.h:206-208; Fixed voxels never set previousDt): the validation
            // build divides as the reference does, the float build multiplies by 1/dt (an IEEE float division is ~10 instructions)
            float dmN, dmP;
            if constexpr (sizeof(R) == 8) {
                dmN = mn.sqrtMassZetaInt2 / ((vx3 && mn.fixed) ? 0.0f : dt);
                dmP = mp.sqrtMassZetaInt2 / ((vx3 && mp.fixed) ? 0.0f : dt);
            } else {
                dmN = mn.sqrtMassZetaInt2 * ((vx3 && mn.fixed) ? __int_as_float(0x7f800000) : invDt);
                dmP = mp.sqrtMassZetaInt2 * ((vx3 && mp.fixed) ? __int_as_float(0x7f800000) : invDt);
            }
            LinkForces<R> f;
            // links are stored by axis (engine upload), so a warp is on one axis except at the two boundaries: the axis can be a
            // template parameter (one switch instead of seven swizzle selects per link)
            const LinkLen<R> len = link_len<R>(voxSize, restLen);
            if (axis == 0)
                link_update<R, 0>(lmat, 0, dpos, len, qn, qp, dmN, dmP, h, f);
            else if (axis == 1)
                link_update<R, 1>(lmat, 1, dpos, len, qn, qp, dmN, dmP, h, f);
            else
                link_update<R, 2>(lmat, 2, dpos, len, qn, qp, dmN, dmP, h, f);
            sP2[3 * l] = h.pos2.x;
            sP2[3 * l + 1] = h.pos2.y;
            sP2[3 * l + 2] = h.pos2.z;
            sA1[3 * l] = h.a1v.x;
            sA1[3 * l + 1] = h.a1v.y;
            sA1[3 * l + 2] = h.a1v.z;
            sA2[3 * l] = h.a2v.x;
            sA2[3 * l + 1] = h.a2v.y;
            sA2[3 * l + 2] = h.a2v.z;
            sLflags[l] = (uint8_t)h.flags;
            if (canFail)
                sLmax[l] = h.maxStrain;
            sFN[FS * l] = f.fNeg.x;
            sFN[FS * l + 1] = f.fNeg.y;
            sFN[FS * l + 2] = f.fNeg.z;
            sFP[FS * l] = f.fPos.x;
            sFP[FS * l + 1] = f.fPos.y;
            sFP[FS * l + 2] = f.fPos.z;
            sMN[FS * l] = f.mNeg.x;
            sMN[FS * l + 1] = f.mNeg.y;
            sMN[FS * l + 2] = f.mNeg.z;
            sMP[FS * l] = f.mPos.x;
            sMP[FS * l + 1] = f.mPos.y;
            sMP[FS * l + 2] = f.mPos.z;
            if (f.strain > 100.0f)
                sh.diverged = 1;
        }
        RP(1)
        RM(3)
        if (collide && !overflow) {
            // common displacement since the rebuild (that of one reference voxel): the rebuild test in the voxel pass removes it,
            // because only RELATIVE motion of two voxels can bring an unlisted pair inside the watch distance
            const int s0 = ns >> 1, v0 = sSurf[s0];
            driftX = (float)sPos[3 * v0] - sRef[3 * s0];
            driftY = (float)sPos[3 * v0 + 1] - sRef[3 * s0 + 1];
            driftZ = (float)sPos[3 * v0 + 2] - sRef[3 * s0 + 2];
            narrow(false);
        }
        // the hit counter alternates between two slots by step parity: the slot of this step is read after the barrier below and
        // cleared during the NEXT step (after that step's barrier), so no clear can overtake a read or an add
        int32_t *hitSlot = (step & 1) ? &sh.hits1 : &sh.hits, *otherSlot = (step & 1) ? &sh.hits : &sh.hits1;
        if (hits)
            atomicAdd(hitSlot, hits);
        RP(2)
        RM(4)
        __syncthreads();
        RP(3)
        const bool divergedNow = sh.diverged != 0;
        collisionCount += *hitSlot;
        if (tid == 0) {
            *otherSlot = 0;
            sh.narrowNext = 0; // every warp has left the narrow phase (barrier above); the next one starts after this step's last barrier
        }
        const bool sampling = vx3 && cycle > 0 && ((step + 1) % cycle == 0);
        // ---- phase F: voxel pass (+ reset of the grid counters and, on ordinary steps, the world bookkeeping) ---------------
        RM(5)
        if (!divergedNow) {
            double cx = 0, cy = 0, cz = 0, cm = 0;
            for (int v = tid; v < nv; v += nt) {
                const DVoxMat &m = vmats[sVmat[v]];
                VoxState<R> s;
                s.pos = {sPos[3 * v], sPos[3 * v + 1], sPos[3 * v + 2]};
                if (!(vx3 && m.fixed)) {
                    s.orient = {sOri[4 * v], sOri[4 * v + 1], sOri[4 * v + 2], sOri[4 * v + 3]};
                    s.linMom = {sLin[3 * v], sLin[3 * v + 1], sLin[3 * v + 2]};
                    s.angMom = {sAng[3 * v], sAng[3 * v + 1], sAng[3 * v + 2]};
                    s.temp = thermal ? sTemp[v] : 0.0f;
                    if (thermal) { // floor offset from the cached average size (valid for s.temp): no double division per floor test
                        s.haveFloorOffset = true;
                        s.floorOffset = sBsa[v] / 2 - m.nomSize / 2;
                    }
                    s.flags = sVflags[v];
                    V3<R> F{R(0), R(0), R(0)}, M{R(0), R(0), R(0)};
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const int l = sVlinks[6 * v + k];
                        if (l == 0xffff)
                            continue;
                        const R *f = (k & 1) ? sFP : sFN, *mo = (k & 1) ? sMP : sMN;
                        F = F + V3<R>{f[FS * l], f[FS * l + 1], f[FS * l + 2]};
                        M = M + V3<R>{mo[FS * l], mo[FS * l + 1], mo[FS * l + 2]};
                    }
                    V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
                    const int slot = sVsurf[v];
                    if (slot != 0xffff) {
                        if (collide)
                            contact = {sCF[3 * slot], sCF[3 * slot + 1], sCF[3 * slot + 2]};
                        if (ciliaOn && m.cilia != 0)
                            cilia = qrot(s.orient, V3<R>{(R)sCilia[3 * v], (R)sCilia[3 * v + 1], (R)sCilia[3 * v + 2]});
                    }
                    WorldStepParams wp;
                    wp.dt = dt;
                    wp.floorEnabled = floorOn;
                    wp.vx3 = vx3;
                    wp.forceField = (FULL && (wflags & WF_FORCE_FIELD)) ? Wg.ff : nullptr; // (not part of the shared-memory copy)
                    wp.fv = FormulaVars{0, 0, 0, (double)collisionCount, t, W.recentAngle, W.closeness, W.numClosePairs, nv};
                    if constexpr (FULL) {
                        VoxExt ext;
                        const VoxExt *pe = nullptr;
                        const int ei = sVext[v];
                        const short4 lat = sLat[v];
                        V3<double> nominal{lat.x * m.nomSize, lat.y * m.nomSize, lat.z * m.nomSize};
                        if (ei >= 0) {
                            const DExt &x = A.ext[ei];
                            ext.dof = x.dof;
                            ext.force = {x.force[0], x.force[1], x.force[2]};
                            ext.moment = {x.moment[0], x.moment[1], x.moment[2]};
                            ext.translation = {x.translation[0], x.translation[1], x.translation[2]};
                            ext.rotationQ = {x.rotQ[0], x.rotQ[1], x.rotQ[2], x.rotQ[3]};
                            pe = &ext;
                        }
                        voxel_update<R>(m, wp, pe, nominal, s, F, M, contact, cilia);
                    } else
                        voxel_update<R>(m, wp, nullptr, V3<double>{0, 0, 0}, s, F, M, contact, cilia);
                    if (varyTemp && m.alphaCTE != 0.0f) { // temperature is only observable through alphaCTE
                        if constexpr (sizeof(R) == 4)
                            s.temp = voxel_temperature_split(tempAmp, sh.sinA, sh.cosA, sPhase[2 * v], sPhase[2 * v + 1], (wflags & WF_EXPANSION) != 0);
                        else
                            s.temp = voxel_temperature(tempAmp, tempPeriod, sPhase[v], t + (double)dt, (wflags & WF_EXPANSION) != 0);
                        sBsa[v] = base_size_average_expr(base_size_expr(m.nomSize, m.alphaCTE, s.temp)); // for the next step's floor and contact tests
                    }
                    sPos[3 * v] = s.pos.x;
                    sPos[3 * v + 1] = s.pos.y;
                    sPos[3 * v + 2] = s.pos.z;
                    sOri[4 * v] = s.orient.w;
                    sOri[4 * v + 1] = s.orient.x;
                    sOri[4 * v + 2] = s.orient.y;
                    sOri[4 * v + 3] = s.orient.z;
                    sLin[3 * v] = s.linMom.x;
                    sLin[3 * v + 1] = s.linMom.y;
                    sLin[3 * v + 2] = s.linMom.z;
                    sAng[3 * v] = s.angMom.x;
                    sAng[3 * v + 1] = s.angMom.y;
                    sAng[3 * v + 2] = s.angMom.z;
                    if (thermal)
                        sTemp[v] = s.temp;
                    sVflags[v] = (uint8_t)s.flags;
                    if (collide && slot != 0xffff) { // moved more than skin/2 since the candidate lists were built?
                        const float ddx = (float)s.pos.x - sRef[3 * slot] - driftX, ddy = (float)s.pos.y - sRef[3 * slot + 1] - driftY,
                                    ddz = (float)s.pos.z - sRef[3 * slot + 2] - driftZ;
                        if (ddx * ddx + ddy * ddy + ddz * ddz > skinHalf2)
                            sh.rebuild = 1;
                    }
                }
                if (sampling && m.isMeasured) {
                    double mass = m.mass;
                    cx += s.pos.x * mass;
                    cy += s.pos.y * mass;
                    cz += s.pos.z * mass;
                    cm += mass;
                }
            }
            if (sampling) { // warp-shuffle reduction, one shared atomic per warp
                for (int o = 16; o > 0; o >>= 1) {
                    cx += __shfl_xor_sync(0xffffffffu, cx, o);
                    cy += __shfl_xor_sync(0xffffffffu, cy, o);
                    cz += __shfl_xor_sync(0xffffffffu, cz, o);
                    cm += __shfl_xor_sync(0xffffffffu, cm, o);
                }
                if ((tid & 31) == 0) {
                    atomicAdd(&sh.com[0], cx);
                    atomicAdd(&sh.com[1], cy);
                    atomicAdd(&sh.com[2], cz);
                    atomicAdd(&sh.com[3], cm);
                }
            }
        }
        RP(4)
        RM(6)
        // ---- end-of-step bookkeeping on the shared-memory copy of the world record.  On ordinary steps the last thread does
        // it right after its own voxels (nothing it needs depends on the other threads); on COM-sampling steps and on
        // divergence it waits for the voxel pass.
        const bool late = sampling || divergedNow;
        if (!late && tid == nt - 1) {
            W.collisionCount = collisionCount;
            W.divergedNow = 0;
            world_end_of_step(A, W, useStop);
            sh.running = W.active;
        }
        __syncthreads();
        if (late) {
            if (tid == 0) {
                W.collisionCount = collisionCount;
                W.divergedNow = divergedNow;
                if (sampling && !divergedNow)
                    for (int k = 0; k < 4; k++) {
                        W.comAcc[k] = sh.com[k];
                        sh.com[k] = 0;
                    }
                if (sampling && !divergedNow && W.maxDistPair != 0) { // closeness needs current positions in HBM
                    for (int v = 0; v < nv; v++) {
                        float tv;
                        load_pos(A, voff + v, &tv, nullptr);
                        store_pos(A, voff + v, V3<double>{sPos[3 * v], sPos[3 * v + 1], sPos[3 * v + 2]}, thermal ? sTemp[v] : tv, sVflags[v]);
                    }
                }
                world_end_of_step(A, W, useStop);
                sh.running = W.active;
            }
            __syncthreads();
        }
        running = sh.running != 0;
        rebuild = sh.rebuild != 0;
        RP(5)
        RM(7)
#ifdef RES_PROFILE
        rpSteps++;
#endif
        step++;
        left--;
        if (!divergedNow)
            t += dt;
    }

#ifdef RES_PROFILE
    if (lane == 0) {
        for (int i = 0; i < 16; i++)
            if (i < 8 || i > 11)
                atomicAdd(&g_resProf[i], (unsigned long long)rp[i]);
        atomicAdd(&g_resProf[8], (unsigned long long)rpSteps);
        atomicAdd(&g_resProf[9], (unsigned long long)rpRebuilds);
        atomicAdd(&g_resProf[10], (unsigned long long)(rpSteps * ((nl + nt - 1) / nt)));  // link rounds
        atomicAdd(&g_resProf[11], (unsigned long long)(rpSteps * ((nv + nt - 1) / nt)));  // voxel rounds
    }
#endif
    // ---- write the world back ------------------------------------------------------------------------------------
    __syncthreads();
    {
        // dynamic part of the world record (everything from `t` up to the formula programs)
        const size_t from = offsetof(WorldD, t) / 4, to = offsetof(WorldD, stop) / 4;
        const uint32_t *src = (const uint32_t *)&W;
        uint32_t *dst = (uint32_t *)&Wg;
        for (int i = (int)from + tid; i < (int)to; i += nt)
            dst[i] = src[i];
    }
    for (int v = tid; v < nv; v += nt) {
        const int g = voff + v;
        float tv;
        load_pos(A, g, &tv, nullptr); // a world without thermal actuation never changes (or stages) its temperatures
        store_pos(A, g, V3<double>{sPos[3 * v], sPos[3 * v + 1], sPos[3 * v + 2]}, thermal ? sTemp[v] : tv, sVflags[v]);
        A.orient[g] = mk4<R>(sOri[4 * v], sOri[4 * v + 1], sOri[4 * v + 2], sOri[4 * v + 3]);
        A.lm[g] = mk4<R>(sLin[3 * v], sLin[3 * v + 1], sLin[3 * v + 2], sAng[3 * v]);
        A.am[g] = mk2<R>(sAng[3 * v + 1], sAng[3 * v + 2]);
    }
    for (int l = tid; l < nl; l += nt) {
        const int g = loff + l;
        R fl;
        if constexpr (sizeof(R) == 4)
            fl = u2f((uint32_t)sLflags[l]);
        else
            fl = __longlong_as_double((long long)sLflags[l]);
        A.lpos2[g] = mk4<R>(sP2[3 * l], sP2[3 * l + 1], sP2[3 * l + 2], fl);
        A.la1[g] = mk4<R>(sA1[3 * l], sA1[3 * l + 1], sA1[3 * l + 2], canFail ? (R)sLmax[l] : R(0));
        A.la2[g] = mk4<R>(sA2[3 * l], sA2[3 * l + 1], sA2[3 * l + 2], R(0));
        if constexpr (!SPILL) {
            A.lfN[g] = mk4<R>(sFN[3 * l], sFN[3 * l + 1], sFN[3 * l + 2], R(0));
            A.lfP[g] = mk4<R>(sFP[3 * l], sFP[3 * l + 1], sFP[3 * l + 2], R(0));
            A.lmN[g] = mk4<R>(sMN[3 * l], sMN[3 * l + 1], sMN[3 * l + 2], R(0));
            A.lmP[g] = mk4<R>(sMP[3 * l], sMP[3 * l + 1], sMP[3 * l + 2], R(0));
        }
    }
}

// ---- host side: size classes ---------------------------------------------------------------------------------------
// A world needs resident_bytes() of dynamic shared memory.  Worlds are binned by how many of them fit on one SM at the
// same time (6, 4, 3, 2 or 1 CTAs per SM); each class is one k_resident launch on its own stream so the block scheduler
// packs small and large worlds onto the same SMs.  Worlds that do not fit even alone are stepped by the streaming kernels.
struct ResidentClass {
    int ctasPerSm = 1, threads = 512;
    bool spill = false; // per-link force arrays in global memory (worlds too big for shared memory otherwise)
    bool full = true;   // some world has boundary conditions, a force field or a failure strain (k_resident<.., FULL>)
    size_t smem = 0; // dynamic shared memory of the launch = largest member
    std::vector<int32_t> worlds;
    int32_t *dList = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
};

template <class R> size_t resident_bytes(int nv, int nl, int ns, int nvm, int nlm, uint32_t wflags, bool spill, int listCap) {
    if (nv <= 0 || nv >= 0xffff || nl >= 0xffff || ns >= 0xffff || nvm > 255 || nlm >= (1 << 14))
        return (size_t)-1;
    return ResidentLayout<R>(nv, std::max(nl, 1), std::max(ns, 1), std::max(nvm, 1), std::max(nlm, 1), wflags, spill, listCap).total;
}

// per-SM shared memory budget: 228 KB per SM, 1 KB reserved per resident CTA, plus this kernel's static shared memory
inline int resident_class_of(size_t bytes, size_t smemPerSm, size_t maxOptin) {
    static const int ks[] = {6, 4, 3, 2, 1};
    for (int k : ks) {
        size_t lim = smemPerSm / k - 1024 - sizeof(ResidentShared) - 64;
        if (k == 1)
            lim = std::min(lim, maxOptin - sizeof(ResidentShared) - 64);
        if (bytes <= lim)
            return k;
    }
    return 0;
}

template <class R>
int resident_launch(const Arrays<R> &A, const ResidentClass &c, int64_t maxSteps, int useStop, int recordPairs) {
    if (c.worlds.empty())
        return VX3B_OK;
    auto kernel = c.spill ? (c.full ? k_resident<R, true, true> : k_resident<R, true, false>) : (c.full ? k_resident<R, false, true> : k_resident<R, false, false>);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e == cudaSuccess) {
        kernel<<<(unsigned)c.worlds.size(), c.threads, c.smem, c.stream>>>(A, c.dList, (long long)maxSteps, useStop, recordPairs);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        set_error(std::string("k_resident launch: ") + cudaGetErrorString(e));
        return VX3B_ERR_CUDA;
    }
    return VX3B_OK;
}

} // namespace vx3b
