// Streaming path: state lives in HBM, one link pass + one voxel pass per step over all worlds of the
// batch at once (flat indexing across worlds).  This is the generic path (any world size, any mix);
// the resident and fused-tile kernels specialise it.
//
// Reference kernels replaced: gpu_update_links, gpu_update_attach (contact part), gpu_update_cilia_force,
// gpu_update_voxels, gpu_update_temperature (VX3/VX3_VoxelyzeKernel.cu:566-650, 833-859) and the serial
// bookkeeping in doTimeStep / CUDA_Simulation (VX3_VoxelyzeKernel.cu:237-359, VX3_SimulationManager.cu:62-117).
#pragma once
#include "engine.cuh"

namespace vx3b {

template <class R> VXD V3<double> load_pos(const Arrays<R> &A, int v, float *temp, uint32_t *flags) {
    D2 a = A.pxy[v], b = A.pzt[v];
    unsigned long long bits = (unsigned long long)__double_as_longlong(b.y);
    if (temp)
        *temp = u2f((uint32_t)(bits & 0xffffffffu));
    if (flags)
        *flags = (uint32_t)(bits >> 32);
    return {a.x, a.y, b.x};
}
template <class R> VXD void store_pos(const Arrays<R> &A, int v, V3<double> p, float temp, uint32_t flags) {
    A.pxy[v] = D2{p.x, p.y};
    unsigned long long bits = ((unsigned long long)flags << 32) | (unsigned long long)f2u(temp);
    A.pzt[v] = D2{p.z, __longlong_as_double((long long)bits)};
}

// Adds (x,y,z,m) into a world's accumulator with one atomic per warp when the whole warp is in the same world.
VXD void world_accumulate(double *acc4, int world, bool contributes, double x, double y, double z, double m) {
    unsigned active = __activemask();
    unsigned same = __match_any_sync(active, world);
    if (same == active && active == 0xffffffffu) {
        if (!contributes)
            x = y = z = m = 0;
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
            z += __shfl_xor_sync(0xffffffffu, z, o);
            m += __shfl_xor_sync(0xffffffffu, m, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(acc4 + 0, x);
            atomicAdd(acc4 + 1, y);
            atomicAdd(acc4 + 2, z);
            atomicAdd(acc4 + 3, m);
        }
    } else if (contributes) {
        atomicAdd(acc4 + 0, x);
        atomicAdd(acc4 + 1, y);
        atomicAdd(acc4 + 2, z);
        atomicAdd(acc4 + 3, m);
    }
}

// ---- link pass --------------------------------------------------------------------------------------
template <class R> __global__ void __launch_bounds__(256) k_links(Arrays<R> A, int first) {
    int l = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= A.nlTot)
        return;
    const int w = A.lworld[l];
    WorldD &W = A.worlds[w];
    if (!W.active || (W.flags & WF_RESIDENT))
        return;
    const uint32_t meta = A.lmeta[l];
    const int axis = meta & 3;
    const DLinkMat lm = A.lmats[meta >> 2];
    const int2 e = A.lends[l];
    const DVoxMat &mn = A.vmats[A.vmat[e.x]];
    const DVoxMat &mp = A.vmats[A.vmat[e.y]];
    const bool vx3 = W.flags & WF_VX3;
    if (vx3 && mn.fixed && mp.fixed)
        return;
    float tn, tp;
    V3<double> pn = load_pos(A, e.x, &tn, nullptr), pp = load_pos(A, e.y, &tp, nullptr);
    auto on = A.orient[e.x], op = A.orient[e.y];
    Qt<R> qn{on.x, on.y, on.z, on.w}, qp{op.x, op.y, op.z, op.w}; // stored w,x,y,z in .x .y .z .w
    double restLen = 0.5 * (base_size(mn, tn) + base_size(mp, tp));
    LinkHist<R> h;
    auto s0 = A.lpos2[l], s1 = A.la1[l], s2 = A.la2[l];
    h.pos2 = {s0.x, s0.y, s0.z};
    h.a1v = {s1.x, s1.y, s1.z};
    h.a2v = {s2.x, s2.y, s2.z};
    if constexpr (sizeof(R) == 4)
        h.flags = f2u(s0.w);
    else
        h.flags = (uint32_t)__double_as_longlong(s0.w);
    h.maxStrain = (float)s1.w;
    // dampingMultiplier = 2*sqrtMass*zetaInternal/previousDt (VX3_Voxel.h:206-208); previousDt == dt for every voxel that steps
    float dmN = mn.sqrtMassZetaInt2 / ((vx3 && mn.fixed) ? 0.0f : W.dt);
    float dmP = mp.sqrtMassZetaInt2 / ((vx3 && mp.fixed) ? 0.0f : W.dt);
    LinkForces<R> f;
    if (W.flags & WF_NONLIN) { // bilinear / data materials, Poisson's ratio: the material's stress/strain model + the link's strain offset
        float so = A.lso[l];
        link_update<R>(lm, axis, V3<double>{pp.x - pn.x, pp.y - pn.y, pp.z - pn.z}, W.voxSize, qn, qp, restLen, dmN, dmP, h, f, &A.lmodels[meta >> 2], &so);
        A.lso[l] = so;
    } else
        link_update<R>(lm, axis, V3<double>{pp.x - pn.x, pp.y - pn.y, pp.z - pn.z}, W.voxSize, qn, qp, restLen, dmN, dmP, h, f);
    R fl;
    if constexpr (sizeof(R) == 4)
        fl = u2f(h.flags);
    else
        fl = __longlong_as_double((long long)h.flags);
    A.lpos2[l] = mk4<R>(h.pos2.x, h.pos2.y, h.pos2.z, fl);
    A.la1[l] = mk4<R>(h.a1v.x, h.a1v.y, h.a1v.z, (R)h.maxStrain);
    A.la2[l] = mk4<R>(h.a2v.x, h.a2v.y, h.a2v.z, (R)f.strain);
    A.lfN[l] = mk4<R>(f.fNeg.x, f.fNeg.y, f.fNeg.z, R(0));
    A.lfP[l] = mk4<R>(f.fPos.x, f.fPos.y, f.fPos.z, R(0));
    A.lmN[l] = mk4<R>(f.mNeg.x, f.mNeg.y, f.mNeg.z, R(0));
    A.lmP[l] = mk4<R>(f.mPos.x, f.mPos.y, f.mPos.z, R(0));
    if (f.strain > 100.0f)
        W.divergedNow = 1; // every link is checked (the reference samples one random link per step)
}

// ---- contact: all pairs of surface voxels inside a world, tiled through shared memory --------------------
struct ContactTile { // work item of k_contact: one block = 128 surface voxels of one world
    int32_t world, first;
};
template <class R> struct SurfRec {
    double x, y, z;
    R lx, ly, lz;
    float temp;
    int32_t mat, v;
    short4 lat;
};
// gridMode: the grid broad phase (kernels_contact.cuh) runs before this kernel; only worlds it does not cover (no grid) are left to
// the all-pairs scan
template <class R> __global__ void __launch_bounds__(128) k_contact(Arrays<R> A, const ContactTile *tiles, int recordPairs, int gridMode) {
    __shared__ SurfRec<R> tile[128];
    const ContactTile ct = tiles[blockIdx.x];
    WorldD &W = A.worlds[ct.world];
    if (!W.active || !(W.flags & WF_COLLISION) || (W.flags & WF_RESIDENT))
        return;
    if (gridMode && W.gridLog2 > 0)
        return;
    const int ns = W.ns, soff = W.soff;
    recordPairs = recordPairs && W.runSteps == 1; // parity tests read the pair set of the run's last step
    const int i = ct.first + threadIdx.x; // local surface index
    SurfRec<R> me;
    const bool have = i < ns;
    const StateSet<R> cur = contact_set(A, W);
    auto load = [&](int slot) {
        SurfRec<R> r;
        int v = A.surf[soff + slot];
        const D2 pa = cur.pxy[v], pb = cur.pzt[v];
        const float temp = u2f((uint32_t)((unsigned long long)__double_as_longlong(pb.y) & 0xffffffffu));
        const V3<double> p{pa.x, pa.y, pb.x};
        auto lm = cur.lm[v];
        r.x = p.x;
        r.y = p.y;
        r.z = p.z;
        r.lx = lm.x;
        r.ly = lm.y;
        r.lz = lm.z;
        r.temp = temp;
        r.mat = A.vmat[v];
        r.v = v;
        r.lat = A.lat[v];
        return r;
    };
    if (have)
        me = load(i);
    V3<R> acc{R(0), R(0), R(0)};
    int hits = 0;
    for (int base = 0; base < ns; base += 128) {
        __syncthreads();
        if (base + threadIdx.x < ns)
            tile[threadIdx.x] = load(base + threadIdx.x);
        __syncthreads();
        if (!have)
            continue;
        const int lim = min(128, ns - base);
        const DVoxMat &m1 = A.vmats[me.mat];
        for (int k = 0; k < lim; k++) {
            const int j = base + k;
            if (j == i)
                continue;
            const SurfRec<R> &o = tile[k];
            const DVoxMat &m2 = A.vmats[o.mat];
            if (m1.fixed && m2.fixed)
                continue;
            // directly linked voxels (lattice neighbours; static topology) never collide: is_neighbor(depth 1)
            int man = abs(me.lat.x - o.lat.x) + abs(me.lat.y - o.lat.y) + abs(me.lat.z - o.lat.z);
            if (man == 1)
                continue;
            V3<R> f;
            // The reference evaluates the pair once with voxel1 = the later surface voxel (first > second) and gives
            // voxel2 the negative; evaluating it from both sides in that same orientation keeps both forces identical.
            bool watched;
            if (i > j) {
                watched = contact_pair<R>(m1, me.temp, {me.x, me.y, me.z}, {me.lx, me.ly, me.lz}, m2, o.temp, {o.x, o.y, o.z}, {o.lx, o.ly, o.lz}, f);
                if (watched) {
                    acc = acc + f;
                    if ((m1.isTarget != 0) != (m2.isTarget != 0))
                        hits++;
                    if (recordPairs) {
                        int slot = atomicAdd(A.pairCount, 1);
                        if (slot < A.pairCap)
                            A.pairBuf[slot] = make_int4(ct.world, me.v - W.voff, o.v - W.voff, 0);
                    }
                }
            } else {
                watched = contact_pair<R>(m2, o.temp, {o.x, o.y, o.z}, {o.lx, o.ly, o.lz}, m1, me.temp, {me.x, me.y, me.z}, {me.lx, me.ly, me.lz}, f);
                if (watched)
                    acc = acc - f;
            }
        }
    }
    if (have)
        A.contactF[soff + i] = mk4<R>(acc.x, acc.y, acc.z, R(0));
    // collisionCount (target vs non-target watched pairs), one atomic per warp
    for (int o = 16; o > 0; o >>= 1)
        hits += __shfl_xor_sync(0xffffffffu, hits, o);
    if ((threadIdx.x & 31) == 0 && hits)
        atomicAdd(&W.collisionCount, hits);
}

// ---- voxel pass -------------------------------------------------------------------------------------
template <class R> __global__ void __launch_bounds__(256) k_voxels(Arrays<R> A, int first) {
    int v = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.nvTot)
        return;
    const int w = A.vworld[v];
    WorldD &W = A.worlds[w];
    if (!W.active || W.divergedNow || (W.flags & WF_RESIDENT))
        return;
    const DVoxMat &m = A.vmats[A.vmat[v]];
    const bool vx3 = W.flags & WF_VX3;
    const bool sampling = vx3 && W.cycle > 0 && ((W.step + 1) % W.cycle == 0);
    VoxState<R> s;
    s.pos = load_pos(A, v, &s.temp, &s.flags);
    if (!(vx3 && m.fixed)) {
        auto o = A.orient[v];
        auto lm = A.lm[v];
        auto am = A.am[v];
        s.orient = {o.x, o.y, o.z, o.w};
        s.linMom = {lm.x, lm.y, lm.z};
        s.angMom = {lm.w, am.x, am.y};
        V3<R> F{R(0), R(0), R(0)}, M{R(0), R(0), R(0)};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            int l = A.vlinks[(size_t)k * A.nvTot + v];
            if (l < 0)
                continue;
            auto f = (k & 1) ? A.lfP[l] : A.lfN[l];
            auto mo = (k & 1) ? A.lmP[l] : A.lmN[l];
            F = F + V3<R>{f.x, f.y, f.z};
            M = M + V3<R>{mo.x, mo.y, mo.z};
        }
        V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
        const int slot = A.vsurf[v];
        if (slot >= 0) {
            if (W.flags & WF_COLLISION) {
                auto c = A.contactF[slot];
                contact = {c.x, c.y, c.z};
            }
            if ((W.flags & WF_CILIA) && m.cilia != 0 && A.baseCilia) {
                D4 b = A.baseCilia[v];
                cilia = qrot(s.orient, V3<R>{(R)b.x, (R)b.y, (R)b.z});
            }
        }
        WorldStepParams wp;
        wp.dt = W.dt;
        wp.floorEnabled = (W.flags & WF_FLOOR) != 0;
        wp.vx3 = vx3;
        wp.forceField = (W.flags & WF_FORCE_FIELD) ? W.ff : nullptr;
        wp.fv = FormulaVars{0, 0, 0, (double)W.collisionCount, W.t, W.recentAngle, W.closeness, W.numClosePairs, W.nv};
        VoxExt ext;
        const VoxExt *pe = nullptr;
        const int ei = A.vext[v];
        short4 lat = A.lat[v];
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
        if ((W.flags & WF_VARY_TEMP) && W.tempPeriod > 0) // temperature the next step will see (start-of-step time t+dt)
            s.temp = voxel_temperature(W.tempAmp, W.tempPeriod, A.phase ? A.phase[v] : 0.0, W.t + (double)W.dt, (W.flags & WF_EXPANSION) != 0);
        store_pos(A, v, s.pos, s.temp, s.flags);
        A.orient[v] = mk4<R>(s.orient.w, s.orient.x, s.orient.y, s.orient.z);
        A.lm[v] = mk4<R>(s.linMom.x, s.linMom.y, s.linMom.z, s.angMom.x);
        A.am[v] = mk2<R>(s.angMom.y, s.angMom.z);
    }
    if (sampling) { // centre of mass after this step (VX3_VoxelyzeKernel.cu:314-321, 477-493)
        double mass = m.mass;
        world_accumulate(W.comAcc, w, m.isMeasured != 0, s.pos.x * mass, s.pos.y * mass, s.pos.z * mass, mass);
    }
}

// ---- centre of mass / distance reduction (initialisation and end of run) ------------------------------------
template <class R> __global__ void __launch_bounds__(256) k_reduce_com(Arrays<R> A, int withDistance) {
    __shared__ double part[8][5];
    const int first = blockIdx.x * blockDim.x, v = first + threadIdx.x;
    const bool in = v < A.nvTot;
    const int w = A.vworld[in ? v : A.nvTot - 1];
    WorldD &W = A.worlds[w];
    double x = 0, y = 0, z = 0, mass = 0, dist = 0;
    bool measured = false;
    if (in) {
        const DVoxMat &m = A.vmats[A.vmat[v]];
        measured = m.isMeasured != 0 && !(A.vghost && A.vghost[v]); // a ghost is measured by the slab that owns it
        V3<double> p = load_pos(A, v, nullptr, nullptr);
        mass = m.mass;
        x = p.x * mass, y = p.y * mass, z = p.z * mass;
        if (withDistance && measured) {
            D2 a = A.ipxy[v];
            double dz = p.z - A.ipz[v], dx = p.x - a.x, dy = p.y - a.y;
            dist = sqrt(dx * dx + dy * dy + dz * dz);
        }
    }
    // a block inside one world (worlds are contiguous): shuffle + shared-memory reduction, five atomics per block
    if (A.vworld[first] == A.vworld[min(first + (int)blockDim.x, A.nvTot) - 1]) {
        if (!measured)
            x = y = z = mass = 0;
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
            z += __shfl_xor_sync(0xffffffffu, z, o);
            mass += __shfl_xor_sync(0xffffffffu, mass, o);
            dist += __shfl_xor_sync(0xffffffffu, dist, o);
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
            part[warp][0] = x, part[warp][1] = y, part[warp][2] = z, part[warp][3] = mass, part[warp][4] = dist;
        }
        __syncthreads();
        if (threadIdx.x < 5) {
            double s = 0;
            for (int k = 0; k < (int)(blockDim.x >> 5); k++)
                s += part[k][threadIdx.x];
            if (threadIdx.x < 4)
                atomicAdd(W.comAcc + threadIdx.x, s);
            else if (withDistance)
                atomicAdd(&W.distAcc, s);
        }
        return;
    }
    if (!in)
        return;
    world_accumulate(W.comAcc, w, measured, x, y, z, mass);
    if (withDistance && measured)
        atomicAdd(&W.distAcc, dist);
}

VXD void finish_com(WorldD &W) {
    if (W.comAcc[3] == 0) {
        W.com[0] = W.com[1] = W.com[2] = 0;
    } else {
        double inv = 1.0 / W.comAcc[3];
        W.com[0] = inv * W.comAcc[0];
        W.com[1] = inv * W.comAcc[1];
        W.com[2] = inv * W.comAcc[2];
    }
    W.comAcc[0] = W.comAcc[1] = W.comAcc[2] = W.comAcc[3] = 0;
}

VXD FormulaVars world_vars(const WorldD &W, double x, double y, double z) {
    return FormulaVars{x, y, z, (double)W.collisionCount, W.t, W.recentAngle, W.closeness, W.numClosePairs, W.nv};
}
// StopConditionMet (VX3_VoxelyzeKernel.cu:160-166): formula > 0.  `t - C` is evaluated directly (the same subtraction the formula
// machine would do, Utils/VX3_MathTree.h:123-126) - every world evaluates this every step.
VXD bool stop_condition_met(const WorldD &W) {
    if (W.stopIsTimeMinusC)
        return W.t - W.stopC > 0;
    return eval_formula(W.stop, world_vars(W, W.com[0], W.com[1], W.com[2])) > 0;
}

// mode 0: after upload (initial COM).  mode 1: end of a run (final COM, fitness, total distance).
template <class R> __global__ void k_world_finish(Arrays<R> A, int mode) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.nWorlds)
        return;
    WorldD &W = A.worlds[w];
    if (mode == 0) {
        finish_com(W);
        for (int k = 0; k < 3; k++)
            W.initialCom[k] = W.finalCom[k] = W.com[k];
    } else {
        const double inv = W.comAcc[3] == 0 ? 0.0 : 1.0 / W.comAcc[3];
        for (int k = 0; k < 3; k++)
            W.finalCom[k] = inv * W.comAcc[k];
        W.comAcc[0] = W.comAcc[1] = W.comAcc[2] = W.comAcc[3] = 0;
        W.totalDistance = W.distAcc;
        W.distAcc = 0;
        W.fitness = eval_formula(W.fitnessF, world_vars(W, W.finalCom[0] - W.initialCom[0], W.finalCom[1] - W.initialCom[1], W.finalCom[2] - W.initialCom[2]));
    }
}

// Before the first step of a run: arm the run counter and test the stop condition (CUDA_Simulation checks it before every step).
template <class R> __global__ void k_world_begin(Arrays<R> A, long long maxSteps, int useStop, int residentMode, int fusedMode) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.nWorlds)
        return;
    WorldD &W = A.worlds[w];
    // residentMode 1: worlds that fit run in k_resident; fusedMode 1: the others run in k_brick when they have brick tables,
    // else in k_fused (fusedMode 0: in k_links + k_voxels)
    uint32_t fl = W.flags & ~(uint32_t)(WF_RESIDENT | WF_FUSED | WF_BRICK);
    if (residentMode && W.residentClass > 0)
        fl |= WF_RESIDENT;
    else if (fusedMode)
        fl |= W.brickOk ? WF_BRICK : WF_FUSED;
    W.flags = fl;
    W.runSteps = maxSteps;
    W.active = (!W.diverged && !W.stopped && maxSteps > 0 && W.nv > 0) ? 1 : 0;
    if (W.active && useStop && (W.flags & WF_STOP_FORMULA) && stop_condition_met(W)) {
        W.stopped = 1;
        W.active = 0;
    }
}

// target closeness: Σ 1/d over target pairs, close-pair count (VX3_VoxelyzeKernel.cu:545-563); serial, runs every `cycle` steps
template <class R> __device__ void compute_target_closeness(const Arrays<R> &A, WorldD &W) {
    if (W.maxDistPair == 0)
        return;
    double Rr = W.maxDistPair * W.voxSize, ret = 0;
    int pairs = 0;
    const StateSet<R> cur = cur_set(A, W.parity);
    auto load_pos = [&](const Arrays<R> &, int v, float *, uint32_t *) {
        const D2 a = cur.pxy[v], b = cur.pzt[v];
        return V3<double>{a.x, a.y, b.x};
    };
    for (int i = 0; i < W.nv; i++) {
        if (!A.vmats[A.vmat[W.voff + i]].isTarget)
            continue;
        V3<double> pi = load_pos(A, W.voff + i, nullptr, nullptr);
        for (int j = i + 1; j < W.nv; j++) {
            if (!A.vmats[A.vmat[W.voff + j]].isTarget)
                continue;
            V3<double> pj = load_pos(A, W.voff + j, nullptr, nullptr);
            double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
            double d = sqrt(dx * dx + dy * dy + dz * dz);
            if (d < Rr)
                pairs++;
            ret += 1 / d;
        }
    }
    W.numClosePairs = pairs;
    W.closeness = ret;
}

// End-of-step bookkeeping of one world (the serial tail of doTimeStep + the loop head of CUDA_Simulation).
// TARGETS = false: the caller knows the world counts no target pairs (maxDistPair == 0) and passes no arrays
template <class R, bool TARGETS = true> __device__ void world_end_of_step(const Arrays<R> &A, WorldD &W, int useStop) {
    W.step++;
    if (W.divergedNow) { // doTimeStep returned false: links were updated, voxels were not, time does not advance
        W.diverged = 1;
        W.active = 0;
        return;
    }
    if ((W.flags & WF_VX3) && W.cycle > 0 && (W.step % W.cycle == 0)) {
        W.angleSamples++;
        double Ax = W.comHist[1][0], Ay = W.comHist[1][1], Az = W.comHist[1][2]; // A = old B
        double Bx = W.com[0], By = W.com[1], Bz = W.com[2];                         // B = COM before this update
        for (int k = 0; k < 3; k++) {
            W.comHist[0][k] = W.comHist[1][k];
            W.comHist[1][k] = W.com[k];
        }
        finish_com(W);
        double Cx = W.com[0], Cy = W.com[1], Cz = W.com[2];
        bool BC = (Bx == Cx && By == Cy && Bz == Cz), AB = (Ax == Bx && Ay == By && Az == Bz);
        if (BC || AB || W.angleSamples < 3)
            W.recentAngle = 0;
        else {
            double ux = Bx - Ax, uy = By - Ay, uz = Bz - Az, vx = Cx - Bx, vy = Cy - By, vz = Cz - Bz;
            W.recentAngle = acos((ux * vx + uy * vy + uz * vz) / (sqrt(ux * ux + uy * uy + uz * uz) * sqrt(vx * vx + vy * vy + vz * vz)));
        }
        if constexpr (TARGETS)
            compute_target_closeness(A, W);
    }
    W.t += W.dt;
    W.runSteps--;
    if (W.runSteps <= 0)
        W.active = 0;
    else if (useStop && (W.flags & WF_STOP_FORMULA) && stop_condition_met(W)) {
        W.stopped = 1;
        W.active = 0;
    }
}

// Global contact grid (kernels_contact.cuh), end of a step, one thread per world (k_world_step, or the last CTA of k_fused): lists found stale are rebuilt at the next step, in a
// new epoch - every bucket head of the old one reads as empty, so nothing has to be cleared.  The flags are read with atomics:
// they were written by other CTAs of a kernel that may still be running on this SM's L1.
VXD void cgrid_end_of_step(WorldD &W) {
    if (!(W.flags & WF_COLLISION) || W.gridLog2 <= 0)
        return;
    const int stale = atomicExch(&W.gRebuildNext, 0);
    W.gRebuild = stale;
    if (stale)
        W.gEpoch++;
}
// k_world_step's work for one streaming world
template <class R> VXD void stream_world_step(const Arrays<R> &A, WorldD &W, int useStop, int *activeCount) {
    if (!W.active || (W.flags & WF_RESIDENT))
        return;
    if (W.flags & (WF_FUSED | WF_BRICK))
        W.parity ^= 1; // k_fused / k_brick has just written the other buffer set
    cgrid_end_of_step(W);
    world_end_of_step(A, W, useStop);
    if (W.active && activeCount)
        atomicAdd(activeCount, 1);
}
template <class R> __global__ void k_world_step(Arrays<R> A, int useStop, int *activeCount) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= A.nWorlds)
        return;
    stream_world_step<R>(A, A.worlds[w], useStop, activeCount);
}

static __global__ void k_zero_int(int *p) { *p = 0; }

} // namespace vx3b
