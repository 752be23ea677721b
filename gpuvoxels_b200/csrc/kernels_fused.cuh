// Fused tile path: ONE launch per time step does the link pass and the voxel pass of every streaming world; link forces go
// from the link threads to the voxel threads through shared memory instead of HBM, and each voxel's / link's state is read
// once and written once per step.  This is the path for bodies too large for k_resident (BASELINE configs[2]: 100^3).
//
// The lattice of a world is cut into bricks of BX x BY x BZ cells; one CTA steps one brick:
//   phase 0  stage pos + orient (+ temperature, material) of the brick's own voxels and of its halo (voxels of neighbouring
//            bricks that share a link with it) in shared memory;
//   phase 1  every link with an end in the brick: those created with one of the brick's voxels (its -x/-y/-z links, i.e. the
//            brick holds their POSITIVE end) are "owned" - their history is updated and written back; links whose positive end
//            lives in a neighbouring brick are recomputed here from the same inputs only for the force on the brick's voxel
//            (the owner computes the identical numbers, so no force has to travel between CTAs).  Forces land in one of the six
//            direction slots of the voxel they act on;
//   phase 2  each own voxel sums its slots in the reference's direction order, integrates and writes its new state.
// State and link history are double-buffered (step parity picks source and destination): a neighbouring CTA may still need a
// voxel's or a link's OLD value after its owner has produced the new one.
//
// Replaces gpu_update_links + gpu_update_voxels (+ gpu_update_temperature, gpu_update_cilia_force) per step
// (VX3/VX3_VoxelyzeKernel.cu:566-650, 846-859) for large worlds; k_contact and k_world_step run around it as on the
// streaming path.  One deliberate difference: when a link diverges (strain > 100) the voxels of that step have already been
// integrated by the time the world is stopped (the reference returns from doTimeStep before its voxel pass).
#pragma once
#include "engine.cuh"
#include "kernels_contact.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

struct BrickD {
    int32_t world, voxBeg, nOwn, nHalo, linkBeg, nLinks, pad0, pad1;
};
struct BrickLink { // one link as seen from one brick
    int32_t link;  // global link index (history arrays)
    uint16_t a, b; // brick-local table index of the negative / positive end
    uint16_t flags; // bits 0-1 axis, bit 2: owned (history is written), bit 3: negative end is an own voxel, bit 4: positive end is
    uint16_t pad;
    uint32_t lmat;  // global link-material index
};
enum { BL_OWNED = 4, BL_NEG_OWN = 8, BL_POS_OWN = 16 };

struct FusedArrays {
    int nBricks;
    const BrickD *bricks;
    const int32_t *bvox;      // per brick: global voxel index of own voxels, then of halo voxels
    const BrickLink *blinks;
    const uint8_t *linkMask;  // per own voxel (same indexing as bvox' own part, compacted per brick): which of the 6 slots hold a link
};

template <class R> struct FusedLayout {
    size_t pos, ori, temp, mat, F, M, total;
    __host__ __device__ static size_t up(size_t x) { return (x + 15) & ~(size_t)15; }
    __host__ __device__ FusedLayout(int capTab, int capOwn) {
        size_t o = 0;
        auto take = [&](size_t bytes) {
            size_t at = o;
            o = up(o + bytes);
            return at;
        };
        pos = take(sizeof(double) * 3 * capTab);
        ori = take(sizeof(R) * 4 * capTab);
        temp = take(sizeof(float) * capTab);
        mat = take(sizeof(int32_t) * capTab);
        F = take(sizeof(R) * 18 * capOwn);
        M = take(sizeof(R) * 18 * capOwn);
        total = o;
    }
};

// src = buffer set holding the world's state at the start of the step (W.parity), dst = the other one
template <class R>
__device__ __forceinline__ void fused_brick(const Arrays<R> &A, const FusedArrays &FA, int capTab, int capOwn, const ContactGrid &G, const BrickD &B, WorldD &W, unsigned char *smem) {
    const StateSet<R> src = cur_set(A, W.parity), dst = cur_set(A, W.parity ^ 1);
    const FusedLayout<R> L(capTab, capOwn);
    double *tPos = (double *)(smem + L.pos);
    R *tOri = (R *)(smem + L.ori);
    float *tTemp = (float *)(smem + L.temp);
    int32_t *tMat = (int32_t *)(smem + L.mat);
    R *sF = (R *)(smem + L.F);
    R *sM = (R *)(smem + L.M);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nTab = B.nOwn + B.nHalo;
    const int32_t *bvox = FA.bvox + B.voxBeg;
    const uint32_t wflags = W.flags;
    const bool vx3 = wflags & WF_VX3;
    const float dt = W.dt;

    // ---- phase 0: stage the end states the links need ------------------------------------------------------------------------
    for (int i = tid; i < nTab; i += nt) {
        const int g = bvox[i];
        const D2 pa = src.pxy[g], pb = src.pzt[g];
        const float temp = u2f((uint32_t)((unsigned long long)__double_as_longlong(pb.y) & 0xffffffffu));
        tPos[3 * i] = pa.x;
        tPos[3 * i + 1] = pa.y;
        tPos[3 * i + 2] = pb.x;
        auto o = src.orient[g];
        tOri[4 * i] = o.x;
        tOri[4 * i + 1] = o.y;
        tOri[4 * i + 2] = o.z;
        tOri[4 * i + 3] = o.w;
        tTemp[i] = temp;
        tMat[i] = A.vmat[g];
    }
    __syncthreads();

    // ---- phase 1: links ------------------------------------------------------------------------------------------------------------
    const BrickLink *bl = FA.blinks + B.linkBeg;
    bool diverged = false;
    for (int r = tid; r < B.nLinks; r += nt) {
        const BrickLink k = bl[r];
        const int axis = k.flags & 3;
        const DVoxMat &mn = A.vmats[tMat[k.a]], &mp = A.vmats[tMat[k.b]];
        const bool skip = vx3 && mn.fixed && mp.fixed; // gpu_update_links skips these; their slots must still be defined
        LinkForces<R> f;
        f.fNeg = f.fPos = f.mNeg = f.mPos = V3<R>{R(0), R(0), R(0)};
        if (!skip) {
            const DLinkMat lm = A.lmats[k.lmat];
            V3<double> dpos{tPos[3 * k.b] - tPos[3 * k.a], tPos[3 * k.b + 1] - tPos[3 * k.a + 1], tPos[3 * k.b + 2] - tPos[3 * k.a + 2]};
            Qt<R> qn{tOri[4 * k.a], tOri[4 * k.a + 1], tOri[4 * k.a + 2], tOri[4 * k.a + 3]};
            Qt<R> qp{tOri[4 * k.b], tOri[4 * k.b + 1], tOri[4 * k.b + 2], tOri[4 * k.b + 3]};
            double restLen = 0.5 * (base_size(mn, tTemp[k.a]) + base_size(mp, tTemp[k.b]));
            LinkHist<R> h;
            auto s0 = src.lpos2[k.link], s1 = src.la1[k.link], s2 = src.la2[k.link];
            h.pos2 = {s0.x, s0.y, s0.z};
            h.a1v = {s1.x, s1.y, s1.z};
            h.a2v = {s2.x, s2.y, s2.z};
            if constexpr (sizeof(R) == 4)
                h.flags = f2u(s0.w);
            else
                h.flags = (uint32_t)__double_as_longlong(s0.w);
            h.maxStrain = (float)s1.w;
            float dmN = mn.sqrtMassZetaInt2 / ((vx3 && mn.fixed) ? 0.0f : dt);
            float dmP = mp.sqrtMassZetaInt2 / ((vx3 && mp.fixed) ? 0.0f : dt);
            link_update<R>(lm, axis, dpos, W.voxSize, qn, qp, restLen, dmN, dmP, h, f);
            if (k.flags & BL_OWNED) {
                R fl;
                if constexpr (sizeof(R) == 4)
                    fl = u2f(h.flags);
                else
                    fl = __longlong_as_double((long long)h.flags);
                dst.lpos2[k.link] = mk4<R>(h.pos2.x, h.pos2.y, h.pos2.z, fl);
                dst.la1[k.link] = mk4<R>(h.a1v.x, h.a1v.y, h.a1v.z, (R)h.maxStrain);
                dst.la2[k.link] = mk4<R>(h.a2v.x, h.a2v.y, h.a2v.z, (R)f.strain);
                if (f.strain > 100.0f)
                    diverged = true;
            }
        } else if (k.flags & BL_OWNED) { // carry the untouched history over to the other buffer
            dst.lpos2[k.link] = src.lpos2[k.link];
            dst.la1[k.link] = src.la1[k.link];
            dst.la2[k.link] = src.la2[k.link];
        }
        if (k.flags & BL_NEG_OWN) { // the negative end takes the link in its positive-direction slot 2*axis
            R *pf = sF + 18 * k.a + 6 * axis, *pm = sM + 18 * k.a + 6 * axis;
            pf[0] = f.fNeg.x, pf[1] = f.fNeg.y, pf[2] = f.fNeg.z;
            pm[0] = f.mNeg.x, pm[1] = f.mNeg.y, pm[2] = f.mNeg.z;
        }
        if (k.flags & BL_POS_OWN) { // the positive end in its negative-direction slot 2*axis+1
            R *pf = sF + 18 * k.b + 6 * axis + 3, *pm = sM + 18 * k.b + 6 * axis + 3;
            pf[0] = f.fPos.x, pf[1] = f.fPos.y, pf[2] = f.fPos.z;
            pm[0] = f.mPos.x, pm[1] = f.mPos.y, pm[2] = f.mPos.z;
        }
    }
    if (diverged)
        W.divergedNow = 1;
    __syncthreads();

    // ---- phase 2: own voxels ---------------------------------------------------------------------------------------------------------
    const bool sampling = vx3 && W.cycle > 0 && ((W.step + 1) % W.cycle == 0);
    const uint8_t *mask = FA.linkMask + B.voxBeg;
    const bool gridCheck = G.nBuckets > 0 && (wflags & WF_COLLISION) && W.gridLog2 > 0 && W.soff >= G.s0;
    double refX = 0, refY = 0, refZ = 0;
    if (gridCheck) {
        const int vm = A.surf[W.soff + (W.ns >> 1)];
        const D2 am = src.pxy[vm], bm = src.pzt[vm];
        refX = am.x, refY = am.y, refZ = bm.x;
    }
    for (int i = tid; i < B.nOwn; i += nt) {
        const int v = bvox[i];
        const DVoxMat &m = A.vmats[tMat[i]];
        VoxState<R> s;
        uint32_t bits;
        {
            D2 b = src.pzt[v];
            bits = (uint32_t)((unsigned long long)__double_as_longlong(b.y) >> 32);
        }
        s.pos = {tPos[3 * i], tPos[3 * i + 1], tPos[3 * i + 2]};
        s.temp = tTemp[i];
        s.flags = bits;
        s.orient = {tOri[4 * i], tOri[4 * i + 1], tOri[4 * i + 2], tOri[4 * i + 3]};
        auto lm = src.lm[v];
        auto am = src.am[v];
        s.linMom = {lm.x, lm.y, lm.z};
        s.angMom = {lm.w, am.x, am.y};
        if (!(vx3 && m.fixed)) {
            V3<R> F{R(0), R(0), R(0)}, M{R(0), R(0), R(0)};
            const uint32_t mk = mask[i];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (!(mk & (1u << k)))
                    continue;
                const R *pf = sF + 18 * i + 3 * k, *pm = sM + 18 * i + 3 * k;
                F = F + V3<R>{pf[0], pf[1], pf[2]};
                M = M + V3<R>{pm[0], pm[1], pm[2]};
            }
            V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
            const int slot = A.vsurf[v];
            if (slot >= 0) {
                if (wflags & WF_COLLISION) {
                    auto c = A.contactF[slot];
                    contact = {c.x, c.y, c.z};
                }
                if ((wflags & WF_CILIA) && m.cilia != 0 && A.baseCilia) {
                    D4 b = A.baseCilia[v];
                    cilia = qrot(s.orient, V3<R>{(R)b.x, (R)b.y, (R)b.z});
                }
            }
            WorldStepParams wp;
            wp.dt = dt;
            wp.floorEnabled = (wflags & WF_FLOOR) != 0;
            wp.vx3 = vx3;
            wp.forceField = (wflags & WF_FORCE_FIELD) ? W.ff : nullptr;
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
            if ((wflags & WF_VARY_TEMP) && W.tempPeriod > 0) // temperature the next step will see (start-of-step time t+dt)
                s.temp = voxel_temperature(W.tempAmp, W.tempPeriod, A.phase ? A.phase[v] : 0.0, W.t + (double)dt, (wflags & WF_EXPANSION) != 0);
        }
        {
            dst.pxy[v] = D2{s.pos.x, s.pos.y};
            unsigned long long b2 = ((unsigned long long)s.flags << 32) | (unsigned long long)f2u(s.temp);
            dst.pzt[v] = D2{s.pos.z, __longlong_as_double((long long)b2)};
            dst.orient[v] = mk4<R>(s.orient.w, s.orient.x, s.orient.y, s.orient.z);
            dst.lm[v] = mk4<R>(s.linMom.x, s.linMom.y, s.linMom.z, s.angMom.x);
            dst.am[v] = mk2<R>(s.angMom.y, s.angMom.z);
        }
        if (gridCheck) { // k_cgrid_check's job: stale candidate lists?  (reference voxel from the source set, which nobody writes)
            const int slot = A.vsurf[v];
            if (slot >= 0)
                cgrid_check_one(W, G, slot, s.pos.x, s.pos.y, s.pos.z, refX, refY, refZ);
        }
        if (sampling) {
            double mass = m.mass;
            world_accumulate(W.comAcc, B.world, m.isMeasured != 0, s.pos.x * mass, s.pos.y * mass, s.pos.z * mass, mass);
        }
    }
}

// G: the global contact grid (G.nBuckets == 0: none): the voxel pass also does k_cgrid_check's job.
// (Folding k_world_step into the last CTA to finish, as k_brick does, was measured and dropped: with ~900 one-brick CTAs the
// fence + ticket of every CTA and the serial tail cost 7 us per step more than the separate launch does.)
template <class R> __global__ void __launch_bounds__(256, 2) k_fused(Arrays<R> A, FusedArrays FA, int capTab, int capOwn, ContactGrid G) {
    extern __shared__ __align__(16) unsigned char smem[];
    const BrickD B = FA.bricks[blockIdx.x];
    WorldD &W = A.worlds[B.world];
    if (W.active && (W.flags & WF_FUSED))
        fused_brick<R>(A, FA, capTab, capOwn, G, B, W, smem);
}

// End of a run: worlds whose current state sits in the alternate buffers are copied back, so that outside vx3b_batch_run the
// main buffers always hold the truth (read-back, the resident and streaming kernels never look at the alternate set).
template <class R> __global__ void k_unparity(Arrays<R> A, int v0, int l0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = v0 + i, l = l0 + i;
    if (v < A.nvTot && A.worlds[A.vworld[v]].parity && (A.worlds[A.vworld[v]].flags & WF_FUSED)) {
        A.pxy[v] = A.alt.pxy[v];
        A.pzt[v] = A.alt.pzt[v];
        A.orient[v] = A.alt.orient[v];
        A.lm[v] = A.alt.lm[v];
        A.am[v] = A.alt.am[v];
    }
    if (l < A.nlTot && A.worlds[A.lworld[l]].parity && (A.worlds[A.lworld[l]].flags & WF_FUSED)) {
        A.lpos2[l] = A.alt.lpos2[l];
        A.la1[l] = A.alt.la1[l];
        A.la2[l] = A.alt.la2[l];
    }
}
template <class R> __global__ void k_clear_parity(Arrays<R> A) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < A.nWorlds)
        A.worlds[w].parity = 0;
}

} // namespace vx3b
