// Dynamic-topology path: worlds with signals, attachment, detachment, voxel removal or the "secondary experiment"
// material switches.  One CTA per world, every step of a run inside ONE launch, state in global memory (L2-resident: such worlds are
// a few thousand voxels), phases separated by block barriers.
//
// Reference kernels replaced: gpu_update_links, gpu_update_attach -> handle_collision_attachment, gpu_update_detach,
// gpu_update_cilia_force, gpu_update_voxels with propagateSignal / packMaker / localSignalDecay, regenerateSurfaceVoxels, removeVoxels
// (VX3/VX3_VoxelyzeKernel.cu:237-399, 495-513, 566-968; VX3/VX3_Voxel.cu:279-348; VX3/VX3_Collision.cu:3-31).
//
// The reference's kernels race on signals (a sender writes its neighbours' fields while their own threads read and reset them) and
// on link creation; the outcome depends on the thread schedule.  This kernel computes, in parallel, the outcome of ONE valid
// schedule - every "thread" of a launch run to completion in index order, the schedule of the reference's step code built for the
// host - so that it can be checked against it:
//   * contact forces are summed per voxel in ascending partner order; target hits raise signals by an idempotent flag;
//   * attachment candidates are found in parallel against the topology at the start of the phase (every test a new link can only
//     turn from pass to fail) and the few survivors are linked by one thread in pair order, re-testing what an earlier new link may
//     have changed;
//   * the signal phase finds every voxel's state "at its turn" by fixed-point sweeps over the lower-indexed senders (a DAG: the
//     sweeps reach the serial result exactly), then applies the turn and the higher-indexed senders.
#pragma once
#include "engine.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

struct P3 {
    double x, y, z;
};

struct DynMatX { // per voxel material: what the dynamic features read beyond DVoxMat
    int32_t isPaceMaker, sticky, localIndex, pad;
    double paceMakerPeriod, signalValueDecay, signalTimeDelay, inactivePeriod, removeAfter, thermalAfter, ciliaAfter;
};

enum { DYN_ERR_LINK_POOL = 1, DYN_ERR_CANDIDATES = 2 };

struct DynWorld {
    WorldD W; // voff / loff / soff: offsets into DynArrays (loff counts link CAPACITY); nl = links in use; ns = listed surface voxels
    int32_t lcap, surfaceChanged, safetyGuard, enableSignals, enableAttach, enableDetach, secondary, reinitDone;
    int32_t nCand, candOff, candCap, error, nTargets, toff, attachLinks, detachedLinks;
    int32_t listsValid, listOverflow, listRebuilds, padList; // contact candidate lists (Verlet lists) of the world
    double reinitAfter;
    VX3B_Token attach[5 * VX3B_MAX_TOKENS];
};

template <class R> struct DynCand { // a pair that would attach on the topology at the start of the contact phase
    int32_t a, b, v1, v2, dir1, dir2, axis, reverse, mat, pad;
    R fx, fy, fz;
};

template <class R> struct DynArrays {
    using T4 = typename Store<R>::T4;
    DynWorld *worlds;
    int nWorlds;
    // voxels
    P3 *pos, *ipos;
    T4 *orient; // w,x,y,z
    V3<R> *linMom, *angMom, *contactF;
    float *temp;
    uint32_t *vflags; // VF_STATIC_FRICTION
    int32_t *vmat;    // index into vmats / matx
    short4 *lat;
    int32_t *vlinks; // [v][6] link index inside the world, or -1
    int32_t *vext;   // index into ext or -1
    double *phase;   // or nullptr
    P3 *baseCilia, *shiftCilia; // or nullptr
    double *sig;  // [v][6] localSignal, localSignalDt, inactiveUntil, signal value, signal activeTime, pace maker's next pulse
    double *sigT; // [v][4] scratch: inactiveUntil, localSignal, signal value, signal activeTime of the voxel at its turn
    uint8_t *vremoved, *vattach, *vhit, *vsurfFlag;
    int32_t *surf;      // surface list per world (capacity nv)
    int32_t *clist, *ccount; // per surface slot: candidate list (DYN_LIST_CAP surface slots, ascending) and its length
    P3 *cref;           // per surface slot: position when the lists were built
    int32_t *targets;   // voxels of target materials per world
    double *rowScratch; // [v] per-row partial sums of the target closeness
    // links (capacity lcap per world)
    int2 *lends;     // voxel index inside the world of (neg, pos)
    uint32_t *lmeta; // axis | index into lmats << 2
    V3<R> *lpos2, *la1, *la2, *lfN, *lfP, *lmN, *lmP;
    uint32_t *lflags;
    float *lmaxStrain, *lstrain, *lso;
    uint8_t *lremoved, *ldetached;
    int32_t *lnew;
    // tables
    DVoxMat *vmats;
    DynMatX *matx;
    uint8_t *matRemoved;
    DLinkMat *lmats;
    DLinkModel *lmodels;
    DExt *ext;
    DynCand<R> *cand;
    int4 *pairBuf;
    int32_t *pairCount;
    int32_t pairCap;
};

struct SigState {
    double iu, ls, sv, sa; // inactiveUntil, localSignal, signal value, signal activeTime
};
// VX3_Voxel::receiveSignal (VX3/VX3_Voxel.cu:306-326)
VXD void receive_signal(SigState &s, const DynMatX &m, double value, double activeTime, bool force) {
    if (!force && s.iu > activeTime)
        return;
    if (value < 0.1)
        return;
    s.iu = activeTime + m.inactivePeriod;
    s.ls = value;
    s.sv = value * m.signalValueDecay;
    if (s.sv < 0.1)
        s.sv = 0;
    s.sa = activeTime;
}
VXD bool signal_fires(const SigState &s, double t) { return !(s.sa > t) && !(s.sv < 0.1); } // propagateSignal's gate (VX3_Voxel.cu:279-304)

// is_neighbor(voxel1, voxel2, nullptr, 5) (VX3/VX3_VoxelyzeKernel.cu:651-680): v2 reachable from v1 over at most five links
__device__ inline bool dyn_is_neighbor5(const int32_t *vlinks, const int2 *lends, int v1, int v2) {
    if (v1 == v2)
        return true;
    int node[5], inc[5], k[5];
    int d = 0;
    node[0] = v1, inc[0] = -1, k[0] = 0;
    while (d >= 0) {
        if (k[d] == 6) {
            d--;
            continue;
        }
        const int l = vlinks[(size_t)node[d] * 6 + k[d]++];
        if (l < 0 || l == inc[d])
            continue;
        const int2 e = lends[l];
        const int o = e.x == node[d] ? e.y : e.x;
        if (o == v2)
            return true;
        if (d + 1 < 5) {
            d++;
            node[d] = o, inc[d] = l, k[d] = 0;
        }
    }
    return false;
}

// Σ over the block in a fixed order (deterministic run to run): shuffle tree inside a warp, then the warps in order.  Every thread
// of the block must call it; the result is valid in thread 0.
VXD double dyn_block_sum(double x, double *sh) {
    for (int o = 16; o > 0; o >>= 1)
        x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
        sh[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = 0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)((blockDim.x + 31) >> 5); k++)
            s += sh[k];
    return s;
}

enum { DYN_THREADS = 512, DYN_TILE = 128, DYN_LIST_CAP = 40 };

template <class R> struct DynSurfRec {
    double x, y, z, avg;
    R lx, ly, lz;
    int32_t mat, v;
    int32_t attachOk;
};

template <class R> __global__ void __launch_bounds__(DYN_THREADS, 1) k_dynamic(DynArrays<R> D, long long maxSteps, int useStop, int recordPairsOpt) {
    __shared__ DynSurfRec<R> tile[DYN_TILE];
    __shared__ double red[DYN_THREADS / 32];
    __shared__ int warpCnt[DYN_THREADS / 32];
    __shared__ int sBase, sActive, sChanged, sAnyFire;

    DynWorld &DW = D.worlds[blockIdx.x];
    WorldD &W = DW.W;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nv = W.nv;
    const size_t voff = (size_t)W.voff, loff = (size_t)W.loff;
    const uint32_t fl = W.flags;
    const bool vx3 = fl & WF_VX3, collision = fl & WF_COLLISION, doAttach = DW.enableAttach != 0, doSignals = DW.enableSignals != 0;
    const float dtf = W.dt;
    // per-world views
    P3 *pos = D.pos + voff;
    typename Store<R>::T4 *orient = D.orient + voff;
    V3<R> *linMom = D.linMom + voff, *angMom = D.angMom + voff, *contactF = D.contactF + voff;
    float *temp = D.temp + voff;
    uint32_t *vflags = D.vflags + voff;
    const int32_t *vmat = D.vmat + voff;
    int32_t *vlinks = D.vlinks + voff * 6;
    double *sig = D.sig + voff * 6, *sigT = D.sigT + voff * 4;
    uint8_t *vremoved = D.vremoved + voff, *vattach = D.vattach + voff, *vhit = D.vhit + voff, *vsurfFlag = D.vsurfFlag + voff;
    int32_t *surf = D.surf + voff;
    int2 *lends = D.lends + loff;
    uint32_t *lmeta = D.lmeta + loff;
    V3<R> *lfN = D.lfN + loff, *lfP = D.lfP + loff, *lmN = D.lmN + loff, *lmP = D.lmP + loff;
    uint8_t *lremoved = D.lremoved + loff, *ldetached = D.ldetached + loff;
    int32_t *lnew = D.lnew + loff;
    DynCand<R> *cand = D.cand + DW.candOff;

    auto stepping = [&](int v) { // voxels that run VX3_Voxel::timeStep to its end, i.e. take their turn in the signal phase
        if (vremoved[v])
            return false;
        if (vx3 && D.vmats[vmat[v]].fixed)
            return false;
        const int ei = D.vext[voff + v];
        return !(ei >= 0 && D.ext[ei].dof == VX3B_DOF_ALL);
    };
    auto other_end = [&](int l, int v) {
        const int2 e = lends[l];
        return e.x == v ? e.y : e.x;
    };

    if (tid == 0) { // k_world_begin's job
        W.runSteps = maxSteps;
        W.active = (!W.diverged && !W.stopped && maxSteps > 0 && nv > 0 && !DW.error) ? 1 : 0;
        if (W.active && useStop && (fl & WF_STOP_FORMULA) && stop_condition_met(W)) {
            W.stopped = 1;
            W.active = 0;
        }
        sActive = W.active;
    }
    __syncthreads();

    while (sActive) {
        const double t = W.t;
        const int recordPairs = recordPairsOpt && W.runSteps == 1;
        // ---- temperature (VX3_VoxelyzeKernel.cu:219-235, 625-650) ---------------------------------------------------------------
        if ((fl & WF_VARY_TEMP) && W.tempPeriod > 0)
            for (int v = tid; v < nv; v += nt) {
                if (vremoved[v])
                    continue;
                const int mi = vmat[v];
                if (vx3 && (D.matx[mi].thermalAfter > t || D.vmats[mi].fixed))
                    continue;
                temp[v] = voxel_temperature(W.tempAmp, W.tempPeriod, D.phase ? D.phase[voff + v] : 0.0, t, (fl & WF_EXPANSION) != 0);
            }
        if (tid == 0) {
            W.step++;
            sChanged = 0;
        }
        __syncthreads();
        // ---- links (VX3_VoxelyzeKernel.cu:566-581, VX3_Link.cu:135-218) -----------------------------------------------------------
        const int nl = W.nl;
        for (int l = tid; l < nl; l += nt) {
            if (lremoved[l] || ldetached[l])
                continue;
            const uint32_t meta = lmeta[l];
            const int axis = meta & 3;
            const DLinkMat lm = D.lmats[meta >> 2];
            const int2 e = lends[l];
            const DVoxMat &mn = D.vmats[vmat[e.x]];
            const DVoxMat &mp = D.vmats[vmat[e.y]];
            if (vx3 && mn.fixed && mp.fixed)
                continue;
            const P3 pn = pos[e.x], pp = pos[e.y];
            const auto on = orient[e.x], op = orient[e.y];
            Qt<R> qn{on.x, on.y, on.z, on.w}, qp{op.x, op.y, op.z, op.w};
            const double restLen = 0.5 * (base_size(mn, temp[e.x]) + base_size(mp, temp[e.y]));
            LinkHist<R> h;
            h.pos2 = D.lpos2[loff + l];
            h.a1v = D.la1[loff + l];
            h.a2v = D.la2[loff + l];
            h.flags = D.lflags[loff + l];
            h.maxStrain = D.lmaxStrain[loff + l];
            const float dmN = mn.sqrtMassZetaInt2 / ((vx3 && mn.fixed) ? 0.0f : dtf);
            const float dmP = mp.sqrtMassZetaInt2 / ((vx3 && mp.fixed) ? 0.0f : dtf);
            LinkForces<R> f;
            const V3<double> dpos{pp.x - pn.x, pp.y - pn.y, pp.z - pn.z};
            if (fl & WF_NONLIN) {
                float so = D.lso[loff + l];
                link_update<R>(lm, axis, dpos, W.voxSize, qn, qp, restLen, dmN, dmP, h, f, &D.lmodels[meta >> 2], &so);
                D.lso[loff + l] = so;
            } else
                link_update<R>(lm, axis, dpos, W.voxSize, qn, qp, restLen, dmN, dmP, h, f);
            if (lnew[l]) { // a link made by attachment eases in: 1 % of its forces for SafetyGuard steps (VX3_Link.cu:206-213)
                const R k = R(0.01);
                f.fNeg = f.fNeg * k, f.fPos = f.fPos * k, f.mNeg = f.mNeg * k, f.mPos = f.mPos * k;
                lnew[l] -= 1;
            }
            D.lpos2[loff + l] = h.pos2;
            D.la1[loff + l] = h.a1v;
            D.la2[loff + l] = h.a2v;
            D.lflags[loff + l] = h.flags;
            D.lmaxStrain[loff + l] = h.maxStrain;
            D.lstrain[loff + l] = f.strain;
            lfN[l] = f.fNeg, lfP[l] = f.fPos, lmN[l] = f.mNeg, lmP[l] = f.mPos;
            if (f.strain > 100.0f)
                W.divergedNow = 1;
        }
        __syncthreads();
        if (W.divergedNow) { // doTimeStep returned false: links were updated, voxels were not, time does not advance
            __syncthreads();
            if (tid == 0) {
                W.diverged = 1;
                W.active = 0;
                sActive = 0;
            }
            __syncthreads();
            break;
        }
        // ---- surface list (regenerateSurfaceVoxels, VX3_VoxelyzeKernel.cu:282-286, 495-513), in voxel order ------------------------
        if (DW.surfaceChanged) {
            if (tid == 0)
                sBase = 0;
            __syncthreads();
            for (int v0 = 0; v0 < nv; v0 += nt) {
                const int v = v0 + tid;
                bool isSurf = false;
                if (v < nv) {
                    bool interior = true;
                    for (int k = 0; k < 6; k++) {
                        const int l = vlinks[(size_t)v * 6 + k];
                        if (l < 0 || ldetached[l])
                            interior = false;
                    }
                    isSurf = !interior && !vremoved[v];
                    vsurfFlag[v] = isSurf ? 1 : 0;
                }
                const unsigned b = __ballot_sync(0xffffffffu, isSurf);
                if (lane == 0)
                    warpCnt[warp] = __popc(b);
                __syncthreads();
                int at = sBase;
                for (int k = 0; k < warp; k++)
                    at += warpCnt[k];
                if (isSurf)
                    surf[at + __popc(b & ((1u << lane) - 1u))] = v;
                __syncthreads();
                if (tid == 0) {
                    int s = 0;
                    for (int k = 0; k < (nt >> 5); k++)
                        s += warpCnt[k];
                    sBase += s;
                }
                __syncthreads();
            }
            if (tid == 0) {
                W.ns = sBase;
                DW.surfaceChanged = 0;
                DW.listsValid = 0; // the candidate lists index the surface list
            }
            __syncthreads();
        }
        // ---- contact, target hits, attachment candidates (VX3_VoxelyzeKernel.cu:682-843, VX3_Collision.cu:3-31) -------------------
        // The reference tests all ns^2/2 surface pairs every step.  Here every surface voxel keeps a candidate list (Verlet list):
        // the surface voxels within the world's largest watch distance + a skin at the last all-pairs scan, in ascending order.  The
        // lists stay valid until some surface voxel has moved more than skin / 2 (or the surface list itself changed); every step
        // each voxel applies the reference's exact predicate to its candidates only, in the all-pairs order - the same pair set, the
        // same sums.  A list that overflows sends the world back to the all-pairs scan for that step.
        const int ns = W.ns;
        if ((collision || doAttach) && ns > 0) {
            auto load = [&](int slot) {
                DynSurfRec<R> r;
                const int v = surf[slot];
                const int mi = vmat[v];
                const DVoxMat &m = D.vmats[mi];
                const P3 p = pos[v];
                const V3<R> lmv = linMom[v];
                r.x = p.x, r.y = p.y, r.z = p.z;
                r.avg = base_size_average(m, temp[v]);
                r.lx = lmv.x, r.ly = lmv.y, r.lz = lmv.z;
                r.mat = mi;
                r.v = vremoved[v] ? -1 : v;
                r.attachOk = (vattach[v] && !m.fixed && D.matx[mi].sticky) ? 1 : 0;
                return r;
            };
            // one pair, seen from surface voxel i (record `me`, linked voxels nb[6]) towards surface voxel j (record `o`)
            auto pair = [&](int i, const DynSurfRec<R> &me, const int *nb, int j, const DynSurfRec<R> &o, V3<R> &acc, int &hits) {
                if (o.v < 0)
                    return;
                const DVoxMat &m1 = D.vmats[me.mat];
                const DVoxMat &m2 = D.vmats[o.mat];
                if (m1.fixed && m2.fixed)
                    return;
                // the reference evaluates a pair once, voxel1 = the later surface voxel, and gives voxel2 the negative; both members
                // evaluate it in that orientation, so the two forces are the same bits
                const bool hi = i > j;
                V3<R> f;
                bool watched;
                if (hi)
                    watched = contact_pair_sized<R>(m1, me.avg, {me.x, me.y, me.z}, {me.lx, me.ly, me.lz}, m2, o.avg, {o.x, o.y, o.z}, {o.lx, o.ly, o.lz}, f);
                else
                    watched = contact_pair_sized<R>(m2, o.avg, {o.x, o.y, o.z}, {o.lx, o.ly, o.lz}, m1, me.avg, {me.x, me.y, me.z}, {me.lx, me.ly, me.lz}, f);
                if (!watched)
                    return;
                if (o.v == nb[0] || o.v == nb[1] || o.v == nb[2] || o.v == nb[3] || o.v == nb[4] || o.v == nb[5])
                    return; // is_neighbor(depth 1): directly linked voxels never collide
                if (collision) {
                    acc = hi ? acc + f : acc - f;
                    if (hi) {
                        if ((m1.isTarget != 0) != (m2.isTarget != 0)) {
                            hits++;
                            if (doSignals) // the voxel that touched a target fires (VX3_VoxelyzeKernel.cu:706-716): forced, the same values from every hit
                                vhit[m1.isTarget ? o.v : me.v] = 1;
                        }
                        if (recordPairs) {
                            const int slot = atomicAdd(D.pairCount, 1);
                            if (slot < D.pairCap)
                                D.pairBuf[slot] = make_int4(blockIdx.x, me.v, o.v, 0);
                        }
                    }
                } else
                    f = V3<R>{R(0), R(0), R(0)};
                if (doAttach && hi && me.attachOk && o.attachOk && me.mat == o.mat) {
                    // the attachment half of handle_collision_attachment (VX3_VoxelyzeKernel.cu:718-831) on the topology at the start of
                    // this phase; a link made earlier in the phase can only turn these tests from pass to fail
                    const int v1 = me.v, v2 = o.v;
                    const auto q = orient[v1];
                    const Qt<double> qd{(double)q.x, (double)q.y, (double)q.z, (double)q.w};
                    const V3<double> ea = qrot_inv(qd, V3<double>{-(me.x - o.x), -(me.y - o.y), -(me.z - o.z)});
                    const double fx = fabs(ea.x), fy = fabs(ea.y), fz = fabs(ea.z);
                    int axis;
                    double comp;
                    if (fx >= fy && fx >= fz)
                        axis = 0, comp = ea.x;
                    else if (fy >= fx && fy >= fz)
                        axis = 1, comp = ea.y;
                    else
                        axis = 2, comp = ea.z;
                    const int reverse = comp < 0 ? 1 : 0;
                    const int dir1 = reverse ? 2 * axis + 1 : 2 * axis, dir2 = reverse ? 2 * axis : 2 * axis + 1;
                    if (vlinks[(size_t)v1 * 6 + dir1] >= 0 || vlinks[(size_t)v2 * 6 + dir2] >= 0)
                        return;
                    const int ml = D.matx[me.mat].localIndex;
                    int mat = -1;
                    for (int q2 = 0; q2 < W.nlm && mat < 0; q2++) { // combinedMaterial (VX3_VoxelyzeKernel.cu:515-528)
                        const uint32_t mp2 = D.lmats[W.lmatOff + q2].matPair;
                        if ((int)(mp2 & 0xffff) == ml && (int)(mp2 >> 16) == ml)
                            mat = W.lmatOff + q2;
                    }
                    if (mat < 0)
                        return;
                    if (dyn_is_neighbor5(vlinks, lends, v1, v2)) // (last: the walk over up to five links is the expensive test)
                        return;
                    const int at = atomicAdd(&DW.nCand, 1);
                    if (at < DW.candCap)
                        cand[at] = DynCand<R>{i, j, v1, v2, dir1, dir2, axis, reverse, mat, 0, f.x, f.y, f.z};
                }
            };
            auto neighbours = [&](const DynSurfRec<R> &me, int *nb) {
                for (int k = 0; k < 6; k++) {
                    const int l = me.v >= 0 ? vlinks[(size_t)me.v * 6 + k] : -1;
                    nb[k] = l >= 0 ? other_end(l, me.v) : -1;
                }
            };
            int32_t *clist = D.clist + voff * DYN_LIST_CAP;
            int32_t *ccount = D.ccount + voff;
            P3 *cref = D.cref + voff;
            // are the lists still good?  (a voxel that moved more than skin / 2 since they were built may have met one it does not list)
            const int validBefore = DW.listsValid; // (read by every thread before anyone may clear it)
            __syncthreads();
            if (validBefore) {
                bool moved = false;
                for (int i = tid; i < ns; i += nt) {
                    const P3 p = pos[surf[i]], q = cref[i];
                    const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
                    moved = moved || (float)(dx * dx + dy * dy + dz * dz) > W.skinHalf2;
                }
                if (moved)
                    DW.listsValid = 0;
            }
            __syncthreads();
            if (!DW.listsValid) { // all-pairs scan through shared-memory tiles: lists of everything within watch distance + skin
                __syncthreads();
                if (tid == 0)
                    DW.listOverflow = 0;
                for (int i0 = 0; i0 < ns; i0 += nt) {
                    const int i = i0 + tid;
                    const bool have = i < ns;
                    P3 me{0, 0, 0};
                    bool meOk = false;
                    if (have) {
                        const int v = surf[i];
                        me = pos[v];
                        meOk = !vremoved[v];
                        cref[i] = me;
                    }
                    int cnt = 0;
                    for (int base = 0; base < ns; base += DYN_TILE) {
                        __syncthreads();
                        if (tid < DYN_TILE && base + tid < ns) {
                            const int v = surf[base + tid];
                            const P3 p = pos[v];
                            tile[tid].x = p.x, tile[tid].y = p.y, tile[tid].z = p.z;
                            tile[tid].v = vremoved[v] ? -1 : v;
                        }
                        __syncthreads();
                        if (!have || !meOk)
                            continue;
                        const int lim = min((int)DYN_TILE, ns - base);
                        for (int k = 0; k < lim; k++) {
                            const double dx = me.x - tile[k].x, dy = me.y - tile[k].y, dz = me.z - tile[k].z;
                            if (dx * dx + dy * dy + dz * dz > W.listR2 || base + k == i || tile[k].v < 0)
                                continue;
                            if (cnt < DYN_LIST_CAP)
                                clist[(size_t)i * DYN_LIST_CAP + cnt] = base + k;
                            cnt++;
                        }
                    }
                    if (have) {
                        ccount[i] = cnt;
                        if (cnt > DYN_LIST_CAP)
                            DW.listOverflow = 1;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    DW.listsValid = DW.listOverflow ? 0 : 1;
                    DW.listRebuilds++;
                }
                __syncthreads();
            }
            if (DW.listsValid) { // narrow phase over the candidate lists
                int hits = 0;
                for (int i = tid; i < ns; i += nt) {
                    const DynSurfRec<R> me = load(i);
                    if (me.v < 0)
                        continue;
                    int nb[6];
                    neighbours(me, nb);
                    V3<R> acc{R(0), R(0), R(0)};
                    const int cnt = ccount[i];
                    for (int c = 0; c < cnt; c++) {
                        const int j = clist[(size_t)i * DYN_LIST_CAP + c];
                        pair(i, me, nb, j, load(j), acc, hits);
                    }
                    contactF[me.v] = acc;
                }
                for (int o = 16; o > 0; o >>= 1)
                    hits += __shfl_xor_sync(0xffffffffu, hits, o);
                if (lane == 0 && hits)
                    atomicAdd(&W.collisionCount, hits);
            } else // a list overflowed: the reference's all-pairs scan for this step
                for (int i0 = 0; i0 < ns; i0 += nt) {
                    const int i = i0 + tid;
                    const bool have = i < ns;
                    DynSurfRec<R> me;
                    int nb[6];
                    if (have) {
                        me = load(i);
                        neighbours(me, nb);
                    }
                    V3<R> acc{R(0), R(0), R(0)};
                    int hits = 0;
                    for (int base = 0; base < ns; base += DYN_TILE) {
                        __syncthreads();
                        if (tid < DYN_TILE && base + tid < ns)
                            tile[tid] = load(base + tid);
                        __syncthreads();
                        if (!have || me.v < 0)
                            continue;
                        const int lim = min((int)DYN_TILE, ns - base);
                        for (int k = 0; k < lim; k++)
                            if (base + k != i)
                                pair(i, me, nb, base + k, tile[k], acc, hits);
                    }
                    if (have && me.v >= 0)
                        contactF[me.v] = acc;
                    for (int o = 16; o > 0; o >>= 1)
                        hits += __shfl_xor_sync(0xffffffffu, hits, o);
                    if (lane == 0 && hits)
                        atomicAdd(&W.collisionCount, hits);
                }
            __syncthreads();
            // ---- the survivors are linked by one thread, in pair order (a ascending, then b) -----------------------------------
            if (tid == 0 && DW.nCand > 0) {
                const int n = DW.nCand;
                if (n > DW.candCap)
                    DW.error = DYN_ERR_CANDIDATES;
                else {
                    int made = 0;
                    for (int done = 0; done < n; done++) {
                        int best = -1;
                        for (int c = 0; c < n; c++) {
                            if (cand[c].a < 0)
                                continue;
                            if (best < 0 || cand[c].a < cand[best].a || (cand[c].a == cand[best].a && cand[c].b < cand[best].b))
                                best = c;
                        }
                        DynCand<R> c = cand[best];
                        cand[best].a = -1;
                        if (vlinks[(size_t)c.v1 * 6 + c.dir1] >= 0 || vlinks[(size_t)c.v2 * 6 + c.dir2] >= 0)
                            continue;
                        if (made > 0 && dyn_is_neighbor5(vlinks, lends, c.v1, c.v2))
                            continue;
                        if (W.nl >= DW.lcap) {
                            DW.error = DYN_ERR_LINK_POOL;
                            break;
                        }
                        // VX3_Link(voxel1, dir1, voxel2, dir2): pVNeg = its second voxel for a positive direction (VX3_Link.cu:31-58)
                        const int l = W.nl++;
                        const size_t L = loff + l;
                        lends[l] = make_int2(c.reverse ? c.v2 : c.v1, c.reverse ? c.v1 : c.v2);
                        lmeta[l] = (uint32_t)c.axis | ((uint32_t)c.mat << 2);
                        vlinks[(size_t)c.v1 * 6 + c.dir1] = l;
                        vlinks[(size_t)c.v2 * 6 + c.dir2] = l;
                        const V3<R> z{R(0), R(0), R(0)};
                        D.lpos2[L] = z, D.la1[L] = z, D.la2[L] = z;
                        lfN[l] = z, lfP[l] = z, lmN[l] = z, lmP[l] = z;
                        D.lflags[L] = LF_SMALL_ANGLE;
                        D.lmaxStrain[L] = 0.0f, D.lstrain[L] = 0.0f;
                        if (D.lso)
                            D.lso[L] = 0.0f;
                        lremoved[l] = 0, ldetached[l] = 0;
                        lnew[l] = DW.safetyGuard;
                        DW.surfaceChanged = 1;
                        DW.attachLinks++;
                        made++;
                        // a pair that has just been linked does not push apart this step
                        const V3<R> f{c.fx, c.fy, c.fz};
                        contactF[c.v1] = contactF[c.v1] - f;
                        contactF[c.v2] = contactF[c.v2] + f;
                    }
                }
                DW.nCand = 0;
            }
            if (doSignals)
                for (int v = tid; v < nv; v += nt)
                    if (vhit[v]) {
                        vhit[v] = 0;
                        double *q = sig + (size_t)v * 6;
                        SigState s{q[2], q[0], q[3], q[4]};
                        receive_signal(s, D.matx[vmat[v]], 100.0, t, true);
                        q[2] = s.iu, q[0] = s.ls, q[3] = s.sv, q[4] = s.sa;
                    }
            __syncthreads();
        }
        // ---- detachment (gpu_update_detach, VX3_VoxelyzeKernel.cu:946-967): a failed link lets go of both voxels --------------------
        if (DW.enableDetach) {
            const int nl2 = W.nl;
            for (int l = tid; l < nl2; l += nt) {
                if (lremoved[l] || ldetached[l])
                    continue;
                const float ef = D.lmats[lmeta[l] >> 2].epsilonFail;
                if (ef != -1.0f && D.lmaxStrain[loff + l] > ef) {
                    ldetached[l] = 1;
                    const int2 e = lends[l];
                    for (int k = 0; k < 6; k++) {
                        if (vlinks[(size_t)e.x * 6 + k] == l)
                            vlinks[(size_t)e.x * 6 + k] = -1;
                        if (vlinks[(size_t)e.y * 6 + k] == l)
                            vlinks[(size_t)e.y * 6 + k] = -1;
                    }
                    DW.surfaceChanged = 1;
                    atomicAdd(&DW.detachedLinks, 1);
                }
            }
            __syncthreads();
        }
        // ---- voxels (gpu_update_cilia_force + gpu_update_voxels, VX3_VoxelyzeKernel.cu:582-624, 846-859; VX3_Voxel.cu:162-277) ------
        const bool sampling = vx3 && W.cycle > 0 && (W.step % W.cycle == 0);
        double cx = 0, cy = 0, cz = 0, cm = 0;
        WorldStepParams wp;
        wp.dt = dtf;
        wp.floorEnabled = (fl & WF_FLOOR) != 0;
        wp.vx3 = vx3;
        wp.forceField = (fl & WF_FORCE_FIELD) ? W.ff : nullptr;
        wp.fv = FormulaVars{0, 0, 0, (double)W.collisionCount, t, W.recentAngle, W.closeness, W.numClosePairs, nv};
        for (int v = tid; v < nv; v += nt) {
            const int mi = vmat[v];
            const DVoxMat &m = D.vmats[mi];
            P3 p = pos[v];
            if (!vremoved[v] && !(vx3 && m.fixed)) {
                VoxState<R> s;
                s.pos = {p.x, p.y, p.z};
                s.temp = temp[v];
                s.flags = vflags[v];
                const auto o = orient[v];
                s.orient = {o.x, o.y, o.z, o.w};
                s.linMom = linMom[v];
                s.angMom = angMom[v];
                V3<R> F{R(0), R(0), R(0)}, M{R(0), R(0), R(0)};
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const int l = vlinks[(size_t)v * 6 + k];
                    if (l < 0)
                        continue;
                    F = F + ((k & 1) ? lfP[l] : lfN[l]);
                    M = M + ((k & 1) ? lmP[l] : lmN[l]);
                }
                V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
                if (collision || doAttach) {
                    contact = contactF[v];
                    contactF[v] = V3<R>{R(0), R(0), R(0)};
                }
                if ((fl & WF_CILIA) && vsurfFlag[v] && m.cilia != 0 && D.baseCilia && !(D.matx[mi].ciliaAfter > t)) {
                    const P3 b = D.baseCilia[voff + v];
                    V3<double> c{b.x, b.y, b.z};
                    if (D.shiftCilia) {
                        const P3 sh = D.shiftCilia[voff + v];
                        const double lsv = sig[(size_t)v * 6];
                        c = {b.x + lsv * sh.x, b.y + lsv * sh.y, b.z + lsv * sh.z};
                    }
                    cilia = qrot(s.orient, V3<R>{(R)c.x, (R)c.y, (R)c.z});
                }
                VoxExt ext;
                const VoxExt *pe = nullptr;
                const int ei = D.vext[voff + v];
                const short4 lat = D.lat[voff + v];
                const V3<double> nominal{lat.x * m.nomSize, lat.y * m.nomSize, lat.z * m.nomSize};
                if (ei >= 0) {
                    const DExt &x = D.ext[ei];
                    ext.dof = x.dof;
                    ext.force = {x.force[0], x.force[1], x.force[2]};
                    ext.moment = {x.moment[0], x.moment[1], x.moment[2]};
                    ext.translation = {x.translation[0], x.translation[1], x.translation[2]};
                    ext.rotationQ = {x.rotQ[0], x.rotQ[1], x.rotQ[2], x.rotQ[3]};
                    pe = &ext;
                }
                voxel_update<R>(m, wp, pe, nominal, s, F, M, contact, cilia);
                p = P3{s.pos.x, s.pos.y, s.pos.z};
                pos[v] = p;
                vflags[v] = s.flags;
                orient[v] = mk4<R>(s.orient.w, s.orient.x, s.orient.y, s.orient.z);
                linMom[v] = s.linMom;
                angMom[v] = s.angMom;
                if (doAttach) { // "sticky status": every AttachCondition formula > 0 at the voxel's new position (VX3_VoxelyzeKernel.cu:607-622)
                    bool any = false, all = true;
                    FormulaVars fv = wp.fv;
                    fv.x = p.x, fv.y = p.y, fv.z = p.z;
                    for (int c = 0; c < 5 && all; c++) {
                        const VX3B_Token *prog = DW.attach + c * VX3B_MAX_TOKENS;
                        if (prog[0].op == VX3B_MT_END)
                            continue;
                        any = true;
                        all = eval_formula(prog, fv) > 0;
                    }
                    vattach[v] = (any && all) ? 1 : 0;
                }
            }
            if (sampling && m.isMeasured) {
                const double mass = m.mass;
                cx += p.x * mass, cy += p.y * mass, cz += p.z * mass, cm += mass;
            }
        }
        __syncthreads();
        // ---- signals: propagateSignal, packMaker, localSignalDecay of every voxel, as if taken in voxel order (VX3_Voxel.cu:268-348) ---
        if (doSignals) {
            if (tid == 0)
                sAnyFire = 0;
            __syncthreads();
            for (int v = tid; v < nv; v += nt) { // sweep 0: the state at its turn of a voxel no lower-indexed neighbour sends to
                const double *q = sig + (size_t)v * 6;
                double *T = sigT + (size_t)v * 4;
                T[0] = q[2], T[1] = q[0], T[2] = q[3], T[3] = q[4];
                if (stepping(v) && signal_fires(SigState{q[2], q[0], q[3], q[4]}, t))
                    sAnyFire = 1;
            }
            __syncthreads();
            // neighbours of a voxel in ascending order (the order in which their turns come)
            auto sorted_neighbours = [&](int v, int *out) {
                int n = 0;
                for (int k = 0; k < 6; k++) {
                    const int l = vlinks[(size_t)v * 6 + k];
                    if (l < 0)
                        continue;
                    const int o = other_end(l, v);
                    int at = n++;
                    while (at > 0 && out[at - 1] > o) {
                        out[at] = out[at - 1];
                        at--;
                    }
                    out[at] = o;
                }
                return n;
            };
            if (sAnyFire) {
                for (int sweep = 0; sweep < nv + 1; sweep++) {
                    __syncthreads();
                    if (tid == 0)
                        sChanged = 0;
                    __syncthreads();
                    for (int v = tid; v < nv; v += nt) {
                        if (!stepping(v))
                            continue;
                        const double *q = sig + (size_t)v * 6;
                        SigState s{q[2], q[0], q[3], q[4]};
                        int nbv[6];
                        const int n = sorted_neighbours(v, nbv);
                        const DynMatX &mv = D.matx[vmat[v]];
                        for (int k = 0; k < n && nbv[k] < v; k++) {
                            const int u = nbv[k];
                            if (!stepping(u))
                                continue;
                            const double *Tu = sigT + (size_t)u * 4;
                            const SigState su{Tu[0], Tu[1], Tu[2], Tu[3]};
                            if (signal_fires(su, t))
                                receive_signal(s, mv, su.sv, t + D.matx[vmat[u]].signalTimeDelay, false);
                        }
                        double *T = sigT + (size_t)v * 4;
                        if (T[0] != s.iu || T[1] != s.ls || T[2] != s.sv || T[3] != s.sa) {
                            T[0] = s.iu, T[1] = s.ls, T[2] = s.sv, T[3] = s.sa;
                            sChanged = 1;
                        }
                    }
                    __syncthreads();
                    if (!sChanged)
                        break;
                }
            }
            __syncthreads();
            for (int v = tid; v < nv; v += nt) {
                if (vremoved[v])
                    continue;
                double *q = sig + (size_t)v * 6;
                const DynMatX &mv = D.matx[vmat[v]];
                const bool mine = stepping(v);
                const double *T = sigT + (size_t)v * 4;
                SigState s{T[0], T[1], T[2], T[3]};
                double lsd = q[1], pm = q[5];
                if (mine) {
                    if (signal_fires(s, t)) { // it has sent to its neighbours
                        s.sv = 0;
                        s.sa = 0;
                        s.iu = t + 2 * mv.signalTimeDelay + mv.inactivePeriod;
                    }
                    if (mv.isPaceMaker && !(pm > t)) {
                        receive_signal(s, mv, 100.0, t, true);
                        pm = t + mv.paceMakerPeriod;
                    }
                    if (!(lsd > t)) {
                        if (s.ls < 0.1)
                            s.ls = 0;
                        else {
                            s.ls = s.ls * 0.9;
                            lsd = t + 0.01;
                        }
                    }
                }
                if (sAnyFire) {
                    int nbv[6];
                    const int n = sorted_neighbours(v, nbv);
                    for (int k = 0; k < n; k++) {
                        const int u = nbv[k];
                        if ((mine && u < v) || !stepping(u)) // (lower-indexed senders are part of the state at its turn)
                            continue;
                        const double *Tu = sigT + (size_t)u * 4;
                        const SigState su{Tu[0], Tu[1], Tu[2], Tu[3]};
                        if (signal_fires(su, t))
                            receive_signal(s, mv, su.sv, t + D.matx[vmat[u]].signalTimeDelay, false);
                    }
                }
                q[0] = s.ls, q[1] = lsd, q[2] = s.iu, q[3] = s.sv, q[4] = s.sa, q[5] = pm;
            }
            __syncthreads();
        }
        // ---- the serial tail of doTimeStep (VX3_VoxelyzeKernel.cu:314-359) ----------------------------------------------------------
        if (sampling) {
            const double sx = dyn_block_sum(cx, red), sy = dyn_block_sum(cy, red), sz = dyn_block_sum(cz, red), sm = dyn_block_sum(cm, red);
            const int ntg = DW.nTargets;
            const int32_t *tg = D.targets + DW.toff;
            const bool closeness = W.maxDistPair != 0;
            if (closeness) { // computeTargetCloseness (VX3_VoxelyzeKernel.cu:545-563): row sums in parallel, rows added in order
                const double Rr = W.maxDistPair * W.voxSize;
                int pairs = 0;
                for (int i = tid; i < ntg; i += nt) {
                    const P3 a = pos[tg[i]];
                    double row = 0;
                    for (int j = i + 1; j < ntg; j++) {
                        const P3 b = pos[tg[j]];
                        const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
                        const double d = sqrt(dx * dx + dy * dy + dz * dz);
                        if (d < Rr)
                            pairs++;
                        row += 1 / d;
                    }
                    D.rowScratch[voff + i] = row;
                }
                for (int o = 16; o > 0; o >>= 1)
                    pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
                __syncthreads();
                if (lane == 0)
                    warpCnt[warp] = pairs;
                __syncthreads();
            }
            if (tid == 0) {
                W.comAcc[0] = sx, W.comAcc[1] = sy, W.comAcc[2] = sz, W.comAcc[3] = sm;
                W.angleSamples++;
                const double Ax = W.comHist[1][0], Ay = W.comHist[1][1], Az = W.comHist[1][2];
                const double Bx = W.com[0], By = W.com[1], Bz = W.com[2];
                for (int k = 0; k < 3; k++) {
                    W.comHist[0][k] = W.comHist[1][k];
                    W.comHist[1][k] = W.com[k];
                }
                finish_com(W);
                const double Cx = W.com[0], Cy = W.com[1], Cz = W.com[2];
                const bool BC = (Bx == Cx && By == Cy && Bz == Cz), AB = (Ax == Bx && Ay == By && Az == Bz);
                if (BC || AB || W.angleSamples < 3)
                    W.recentAngle = 0;
                else {
                    const double ux = Bx - Ax, uy = By - Ay, uz = Bz - Az, wx = Cx - Bx, wy = Cy - By, wz = Cz - Bz;
                    W.recentAngle = acos((ux * wx + uy * wy + uz * wz) / (sqrt(ux * ux + uy * uy + uz * uz) * sqrt(wx * wx + wy * wy + wz * wz)));
                }
                if (closeness) {
                    double ret = 0;
                    int pairs = 0;
                    for (int i = 0; i < ntg; i++)
                        ret += D.rowScratch[voff + i];
                    for (int k = 0; k < (nt >> 5); k++)
                        pairs += warpCnt[k];
                    W.closeness = ret;
                    W.numClosePairs = pairs;
                }
            }
            __syncthreads();
        }
        if (DW.secondary) { // removeVoxels + the re-initialised initial position (VX3_VoxelyzeKernel.cu:336-350, 365-399)
            for (int i = 0; i < W.nvm; i++) {
                const int mi = W.vmatOff + i;
                const double after = D.matx[mi].removeAfter;
                if (D.matRemoved[mi] || !(after > 0) || !(after < t))
                    continue; // (the same for every thread)
                for (int v = tid; v < nv; v += nt) {
                    if (vmat[v] != mi || vremoved[v])
                        continue;
                    vremoved[v] = 1;
                    for (int k = 0; k < 6; k++) {
                        const int l = vlinks[(size_t)v * 6 + k];
                        if (l < 0)
                            continue;
                        lremoved[l] = 1;
                        const int o = other_end(l, v);
                        for (int q = 0; q < 6; q++)
                            if (vlinks[(size_t)o * 6 + q] == l)
                                vlinks[(size_t)o * 6 + q] = -1;
                        vlinks[(size_t)v * 6 + k] = -1;
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    D.matRemoved[mi] = 1;
                    DW.surfaceChanged = 1;
                }
                __syncthreads();
            }
            if (!DW.reinitDone && DW.reinitAfter < t) {
                for (int v = tid; v < nv; v += nt)
                    D.ipos[voff + v] = pos[v];
                __syncthreads();
                if (tid == 0) {
                    DW.reinitDone = 1;
                    for (int k = 0; k < 3; k++)
                        W.initialCom[k] = W.com[k];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            W.t += W.dt;
            W.runSteps--;
            if (DW.error)
                W.active = 0;
            else if (W.runSteps <= 0)
                W.active = 0;
            else if (useStop && (fl & WF_STOP_FORMULA) && stop_condition_met(W)) {
                W.stopped = 1;
                W.active = 0;
            }
            sActive = W.active;
        }
        __syncthreads();
    }
    // ---- end of the run: centre of mass, distance travelled, fitness (k_reduce_com + k_world_finish) -------------------------------
    {
        double cx = 0, cy = 0, cz = 0, cm = 0, dist = 0;
        for (int v = tid; v < nv; v += nt) {
            const DVoxMat &m = D.vmats[vmat[v]];
            if (!m.isMeasured)
                continue;
            const P3 p = pos[v], q = D.ipos[voff + v];
            const double mass = m.mass;
            cx += p.x * mass, cy += p.y * mass, cz += p.z * mass, cm += mass;
            const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
            dist += sqrt(dx * dx + dy * dy + dz * dz);
        }
        const double sx = dyn_block_sum(cx, red), sy = dyn_block_sum(cy, red), sz = dyn_block_sum(cz, red), sm = dyn_block_sum(cm, red),
                     sd = dyn_block_sum(dist, red);
        if (tid == 0) {
            const double inv = sm == 0 ? 0.0 : 1.0 / sm;
            W.finalCom[0] = inv * sx, W.finalCom[1] = inv * sy, W.finalCom[2] = inv * sz;
            W.totalDistance = sd;
            W.fitness = eval_formula(W.fitnessF, world_vars(W, W.finalCom[0] - W.initialCom[0], W.finalCom[1] - W.initialCom[1], W.finalCom[2] - W.initialCom[2]));
        }
    }
}

} // namespace vx3b
