// Brick path: the fused link + voxel step for large bodies (BASELINE configs[2]: 100^3) on a BRICK-ORDERED copy of the state.
//
// k_fused (kernels_fused.cuh) showed where the time goes on a 1e6-voxel body: not HBM (25 % of peak) but instruction issue and
// exposed load latency - per-link index tables, gathered 16-byte loads with two dependent global loads per link, generic
// per-voxel feature tests, 3-way bank conflicts.  This path changes the data layout instead of the arithmetic:
//   * at the start of a run the voxels and links of every eligible world are copied ("packed") into brick order: the own
//     voxels of a brick (BX x BY x BZ lattice cells, x fastest) and the links a brick owns (those whose POSITIVE end is an own
//     voxel; x-, y-, z-links in that order) are contiguous, 16-byte aligned runs of the state arrays;
//   * one CTA steps one brick.  One thread arms an mbarrier and issues eight 1-D bulk copies (cp.async.bulk, the TMA engine's
//     linear mode: UBLKCP in SASS) that bring the brick's voxel state, link history and link descriptors into shared memory
//     while the other threads gather the halo (voxels of neighbouring bricks that share a link with this one) and prefetch
//     the history of the recomputed halo links; nothing in the two compute phases waits on a global load;
//   * link forces travel from the link threads to the voxel threads through shared memory in a [slot][component][voxel]
//     layout (conflict-free on both sides); new state goes straight from registers to HBM, coalesced, into the other buffer
//     set (ping-pong by W.parity, as k_fused);
//   * the world's material tables are staged in shared memory once per CTA with the per-material damping multiplier
//     2*sqrt(m)*zeta/dt already divided;
//   * LEAN instantiation (no boundary conditions, no cilia, no thermal actuation in any brick world of the launch) compiles
//     the feature tests away.
// Links cut by a brick face are recomputed by the brick that does not own them, from the same inputs, so no force crosses
// CTAs (same rule as k_fused).  Worlds with voxel-voxel contact, a force field or target materials stay on k_fused: their
// per-step helpers read the API-order arrays.
//
// Replaces gpu_update_links + gpu_update_voxels (+ gpu_update_temperature, gpu_update_cilia_force)
// (VX3/VX3_VoxelyzeKernel.cu:566-650, 846-859) for large worlds.  Same deliberate difference as k_fused on divergence.
#pragma once
#include "engine.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

struct BrickB {
    int32_t world;
    int32_t v0, nOwn, nHalo;    // own voxels: brick-order slots [v0, v0+nOwn), v0 even; halo: haloVox[haloBeg .. +nHalo)
    int32_t l0, nOwned, nHaloL; // owned links: brick-order slots [l0, l0+nOwned); recomputed links: haloLink[haloLinkBeg .. +nHaloL)
    int32_t haloBeg, descBeg, haloLinkBeg; // descBeg even: descriptors of the halo links, then of the owned links
    int32_t pad0, pad1;
};
// One link as seen from one brick.  a/b: table index of the negative / positive end (own voxel i -> i, halo voxel j -> nOwnPad + j)
struct alignas(8) LinkDesc {
    uint16_t a, b, lmat, flags; // lmat: world-local link material; flags: bits 0-1 axis, BD_NEG_OWN, BD_POS_OWN, bits 4-7 / 8-11 world-local voxel material of a / b
};
enum { BD_NEG_OWN = 4, BD_POS_OWN = 8 };
enum { BRICK_MAX_VMATS = 16, BRICK_MAX_LMATS = 48 };

template <class R> struct BrickArrays {
    int nBricks;
    const BrickB *bricks;
    const int32_t *haloVox;  // brick-order voxel slot of each halo voxel
    const LinkDesc *ldesc;
    const int32_t *haloLink; // brick-order link slot of each recomputed link
    const uint32_t *bstat;   // per brick-order voxel slot: world-local material | link mask << 8 | has boundary conditions << 14
    const int32_t *borig;    // brick-order voxel slot -> global (API-order) voxel index, -1 for alignment padding
    const int32_t *lorig;    // brick-order link slot -> global link index
    StateSet<R> set[2];
    int nvSlots, nlSlots;
};
struct BrickCaps {
    int own, halo, owned, links, nvm, nlm; // own and links are even
};

template <class R> struct BrickLayout { // byte offsets into the CTA's dynamic shared memory; computed once on the host
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    uint32_t bar, pxy, pzt, ori, lm, am, h0, h1, h2, desc, F, vmat, lmat, damp, len, total;
    int capOwn;
    static uint32_t up(size_t x) { return (uint32_t)((x + 127) & ~(size_t)127); }
    BrickLayout() {}
    explicit BrickLayout(const BrickCaps &c) {
        uint32_t o = 0;
        auto take = [&](size_t bytes) {
            uint32_t at = o;
            o = up(o + bytes);
            return at;
        };
        const size_t tab = (size_t)c.own + c.halo;
        capOwn = c.own;
        bar = take(8);
        pxy = take(sizeof(D2) * tab);
        pzt = take(sizeof(D2) * tab);
        ori = take(sizeof(T4) * tab);
        lm = take(sizeof(T4) * c.own);
        am = take(sizeof(T2) * c.own);
        h0 = take(sizeof(T4) * c.owned);
        h1 = take(sizeof(T4) * c.owned);
        h2 = take(sizeof(T4) * c.owned);
        desc = take(sizeof(LinkDesc) * c.links);
        F = take(sizeof(R) * 36 * c.own);
        vmat = take(sizeof(DVoxMat) * c.nvm);
        lmat = take(sizeof(DLinkMat) * c.nlm);
        damp = take(sizeof(float) * c.nvm);
        len = take(sizeof(LinkLen<R>) * c.nlm);
        total = o;
    }
};

// ---- mbarrier / bulk-copy PTX (sm_90+; the guide's TMA + mbarrier pattern in its 1-D form) --------------------------------
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

template <class R, bool LEAN> __global__ void __launch_bounds__(256, 2) k_brick(Arrays<R> A, BrickArrays<R> BA, const BrickLayout<R> L) {
    using T4 = typename Store<R>::T4;
    using T2 = typename Store<R>::T2;
    extern __shared__ __align__(128) unsigned char smem_brick[];
    unsigned char *smem = smem_brick;
    const BrickB B = BA.bricks[blockIdx.x];
    WorldD &W = A.worlds[B.world];
    if (!W.active || !(W.flags & WF_BRICK))
        return;
    const int par = W.parity;
    const StateSet<R> src = BA.set[par], dst = BA.set[par ^ 1];
    D2 *sPxy = (D2 *)(smem + L.pxy), *sPzt = (D2 *)(smem + L.pzt);
    T4 *sOri = (T4 *)(smem + L.ori), *sLm = (T4 *)(smem + L.lm);
    T2 *sAm = (T2 *)(smem + L.am);
    T4 *sH0 = (T4 *)(smem + L.h0), *sH1 = (T4 *)(smem + L.h1), *sH2 = (T4 *)(smem + L.h2);
    const LinkDesc *sDesc = (const LinkDesc *)(smem + L.desc);
    R *sF = (R *)(smem + L.F);
    DVoxMat *sVm = (DVoxMat *)(smem + L.vmat);
    DLinkMat *sLmat = (DLinkMat *)(smem + L.lmat);
    float *sDamp = (float *)(smem + L.damp);
    LinkLen<R> *sLen = (LinkLen<R> *)(smem + L.len);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nOwn = B.nOwn, nOwnPad = (nOwn + 1) & ~1, nLinks = B.nHaloL + B.nOwned, nLinksPad = (nLinks + 1) & ~1;
    const int capOwn = L.capOwn;
    const uint32_t wflags = W.flags;
    const bool vx3 = wflags & WF_VX3;
    const float dt = W.dt;

    // ---- phase 0: bulk copies of the brick's own runs (one thread), halo gather + material staging (everybody) -------------
    const uint32_t bar = smem_u32(smem + L.bar);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const uint32_t bV16 = (uint32_t)nOwnPad * 16u, bV4 = (uint32_t)nOwnPad * (uint32_t)sizeof(T4), bV2 = (uint32_t)nOwnPad * (uint32_t)sizeof(T2);
        const uint32_t bL = (uint32_t)B.nOwned * (uint32_t)sizeof(T4), bD = (uint32_t)nLinksPad * (uint32_t)sizeof(LinkDesc);
        mbar_expect_tx(bar, 2 * bV16 + 2 * bV4 + bV2 + 3 * bL + bD);
        if (nOwnPad) {
            bulk_g2s(smem_u32(sPxy), src.pxy + B.v0, bV16, bar);
            bulk_g2s(smem_u32(sPzt), src.pzt + B.v0, bV16, bar);
            bulk_g2s(smem_u32(sOri), src.orient + B.v0, bV4, bar);
            bulk_g2s(smem_u32(sLm), src.lm + B.v0, bV4, bar);
            bulk_g2s(smem_u32(sAm), src.am + B.v0, bV2, bar);
        }
        if (B.nOwned) {
            bulk_g2s(smem_u32(sH0), src.lpos2 + B.l0, bL, bar);
            bulk_g2s(smem_u32(sH1), src.la1 + B.l0, bL, bar);
            bulk_g2s(smem_u32(sH2), src.la2 + B.l0, bL, bar);
        }
        if (nLinksPad)
            bulk_g2s(smem_u32(sDesc), BA.ldesc + B.descBeg, bD, bar);
    }
    // history of this thread's first recomputed link and the static word of its first voxel: in flight across the barrier
    T4 pre0, pre1, pre2;
    pre0 = pre1 = pre2 = mk4<R>(R(0), R(0), R(0), R(0));
    if (tid < B.nHaloL) {
        const int hl = BA.haloLink[B.haloLinkBeg + tid];
        pre0 = src.lpos2[hl];
        pre1 = src.la1[hl];
        pre2 = src.la2[hl];
    }
    uint32_t stat0 = tid < nOwn ? BA.bstat[B.v0 + tid] : 0u;
    for (int i = tid; i < B.nHalo; i += nt) {
        const int g = BA.haloVox[B.haloBeg + i];
        sPxy[nOwnPad + i] = src.pxy[g];
        sPzt[nOwnPad + i] = src.pzt[g];
        sOri[nOwnPad + i] = src.orient[g];
    }
    {
        const uint32_t *gv = (const uint32_t *)(A.vmats + W.vmatOff), *gl = (const uint32_t *)(A.lmats + W.lmatOff);
        uint32_t *sv = (uint32_t *)sVm, *sl = (uint32_t *)sLmat;
        const int nwv = W.nvm * (int)(sizeof(DVoxMat) / 4), nwl = W.nlm * (int)(sizeof(DLinkMat) / 4);
        for (int i = tid; i < nwv; i += nt)
            sv[i] = gv[i];
        for (int i = tid; i < nwl; i += nt)
            sl[i] = gl[i];
        if (tid < W.nvm) { // dampingMultiplier numerator / previousDt (VX3/VX3_Voxel.h:206-208); Fixed voxels never set previousDt
            const DVoxMat &m = A.vmats[W.vmatOff + tid];
            sDamp[tid] = m.sqrtMassZetaInt2 / ((vx3 && m.fixed) ? 0.0f : dt);
        }
        if (tid >= 32 && tid - 32 < W.nlm) { // rest lengths per link material at the sizes of a world whose temperature cannot change them
            const uint32_t mp = A.lmats[W.lmatOff + tid - 32].matPair;
            sLen[tid - 32] = link_len<R>(W.voxSize, 0.5 * (A.vmats[W.vmatOff + (mp & 0xffff)].baseSize0 + A.vmats[W.vmatOff + (mp >> 16)].baseSize0));
        }
    }
    __syncthreads();
    mbar_wait(bar, 0);

    // ---- phase 1: links (recomputed halo links first, then the owned run) -----------------------------------------------------
    bool diverged = false;
    for (int r = tid; r < nLinks; r += nt) {
        const LinkDesc k = sDesc[r];
        const int axis = k.flags & 3, ia = k.a, ib = k.b;
        const int o = r - B.nHaloL; // >= 0: owned link, slot l0 + o
        T4 s0, s1, s2;
        if (o >= 0) {
            s0 = sH0[o];
            s1 = sH1[o];
            s2 = sH2[o];
        } else if (r == tid) {
            s0 = pre0;
            s1 = pre1;
            s2 = pre2;
        } else {
            const int hl = BA.haloLink[B.haloLinkBeg + r];
            s0 = src.lpos2[hl];
            s1 = src.la1[hl];
            s2 = src.la2[hl];
        }
        const int ma = (k.flags >> 4) & 15, mb = (k.flags >> 8) & 15;
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
            const DLinkMat &lmat = sLmat[k.lmat];
            const LinkLen<R> &len = sLen[k.lmat];
            if (axis == 0)
                brick_link<R, LEAN, 0>(lmat, len, sVm, sDamp, ma, mb, W.voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
            else if (axis == 1)
                brick_link<R, LEAN, 1>(lmat, len, sVm, sDamp, ma, mb, W.voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
            else
                brick_link<R, LEAN, 2>(lmat, len, sVm, sDamp, ma, mb, W.voxSize, axy, azt, bxy, bzt, oa, ob, h, f);
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
        if (k.flags & BD_NEG_OWN) { // the negative end takes the link in its positive-direction slot 2*axis
            R *p = sF + (12 * axis) * capOwn + ia;
            p[0] = f.fNeg.x, p[capOwn] = f.fNeg.y, p[2 * capOwn] = f.fNeg.z;
            p[3 * capOwn] = f.mNeg.x, p[4 * capOwn] = f.mNeg.y, p[5 * capOwn] = f.mNeg.z;
        }
        if (k.flags & BD_POS_OWN) { // the positive end in its negative-direction slot 2*axis+1
            R *p = sF + (12 * axis + 6) * capOwn + ib;
            p[0] = f.fPos.x, p[capOwn] = f.fPos.y, p[2 * capOwn] = f.fPos.z;
            p[3 * capOwn] = f.mPos.x, p[4 * capOwn] = f.mPos.y, p[5 * capOwn] = f.mPos.z;
        }
    }
    if (diverged)
        W.divergedNow = 1;
    __syncthreads();

    // ---- phase 2: own voxels ---------------------------------------------------------------------------------------------------------
    const bool sampling = vx3 && W.cycle > 0 && ((W.step + 1) % W.cycle == 0);
    for (int i = tid; i < nOwn; i += nt) {
        const uint32_t st = (i == tid) ? stat0 : BA.bstat[B.v0 + i];
        const DVoxMat &m = sVm[st & 255];
        const D2 pxy = sPxy[i], pzt = sPzt[i];
        const T4 ori = sOri[i], lm = sLm[i];
        const T2 am = sAm[i];
        const unsigned long long packed = (unsigned long long)__double_as_longlong(pzt.y);
        VoxState<R> s;
        s.pos = {pxy.x, pxy.y, pzt.x};
        s.temp = u2f((uint32_t)(packed & 0xffffffffu));
        s.flags = (uint32_t)(packed >> 32);
        s.orient = {ori.x, ori.y, ori.z, ori.w};
        s.linMom = {lm.x, lm.y, lm.z};
        s.angMom = {lm.w, am.x, am.y};
        bool fixedMat = false;
        if constexpr (!LEAN)
            fixedMat = vx3 && m.fixed;
        if (!fixedMat) {
            V3<R> F{R(0), R(0), R(0)}, M{R(0), R(0), R(0)};
            const uint32_t mk = st >> 8;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (!(mk & (1u << k)))
                    continue;
                const R *p = sF + (6 * k) * capOwn + i;
                F = F + V3<R>{p[0], p[capOwn], p[2 * capOwn]};
                M = M + V3<R>{p[3 * capOwn], p[4 * capOwn], p[5 * capOwn]};
            }
            V3<R> contact{R(0), R(0), R(0)}, cilia{R(0), R(0), R(0)};
            WorldStepParams wp;
            wp.dt = dt;
            wp.floorEnabled = (wflags & WF_FLOOR) != 0;
            wp.vx3 = vx3;
            wp.forceField = nullptr;
            VoxExt ext;
            const VoxExt *pe = nullptr;
            V3<double> nominal{0, 0, 0};
            if constexpr (!LEAN) {
                const int g = BA.borig[B.v0 + i];
                if ((wflags & WF_CILIA) && m.cilia != 0 && A.baseCilia && A.vsurf[g] >= 0) {
                    D4 b = A.baseCilia[g];
                    cilia = qrot(s.orient, V3<R>{(R)b.x, (R)b.y, (R)b.z});
                }
                if (st & (1u << 14)) {
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
            voxel_update<R>(m, wp, pe, nominal, s, F, M, contact, cilia);
            if constexpr (!LEAN) {
                if ((wflags & WF_VARY_TEMP) && W.tempPeriod > 0) { // temperature the next step will see (start-of-step time t+dt)
                    const int g = BA.borig[B.v0 + i];
                    s.temp = voxel_temperature(W.tempAmp, W.tempPeriod, A.phase ? A.phase[g] : 0.0, W.t + (double)dt, (wflags & WF_EXPANSION) != 0);
                }
            }
        }
        const int v = B.v0 + i;
        dst.pxy[v] = D2{s.pos.x, s.pos.y};
        const unsigned long long b2 = ((unsigned long long)s.flags << 32) | (unsigned long long)f2u(s.temp);
        dst.pzt[v] = D2{s.pos.z, __longlong_as_double((long long)b2)};
        dst.orient[v] = mk4<R>(s.orient.w, s.orient.x, s.orient.y, s.orient.z);
        dst.lm[v] = mk4<R>(s.linMom.x, s.linMom.y, s.linMom.z, s.angMom.x);
        dst.am[v] = mk2<R>(s.angMom.y, s.angMom.z);
        if (sampling) {
            const double mass = m.mass;
            world_accumulate(W.comAcc, B.world, m.isMeasured != 0, s.pos.x * mass, s.pos.y * mass, s.pos.z * mass, mass);
        }
    }
}

// Start of a run: the API-order state of every brick world -> buffer set 0 in brick order (W.parity is 0 outside a run).
template <class R> __global__ void k_brick_pack(Arrays<R> A, BrickArrays<R> BA) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < BA.nvSlots) {
        const int g = BA.borig[i];
        if (g >= 0 && (A.worlds[A.vworld[g]].flags & WF_BRICK)) {
            BA.set[0].pxy[i] = A.pxy[g];
            BA.set[0].pzt[i] = A.pzt[g];
            BA.set[0].orient[i] = A.orient[g];
            BA.set[0].lm[i] = A.lm[g];
            BA.set[0].am[i] = A.am[g];
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
                const StateSet<R> &s = BA.set[W.parity];
                A.pxy[g] = s.pxy[i];
                A.pzt[g] = s.pzt[i];
                A.orient[g] = s.orient[i];
                A.lm[g] = s.lm[i];
                A.am[g] = s.am[i];
            }
        }
    }
    if (i < BA.nlSlots) {
        const int g = BA.lorig[i];
        if (g >= 0) {
            const WorldD &W = A.worlds[A.lworld[g]];
            if (W.flags & WF_BRICK) {
                const StateSet<R> &s = BA.set[W.parity];
                A.lpos2[g] = s.lpos2[i];
                A.la1[g] = s.la1[i];
                A.la2[g] = s.la2[i];
            }
        }
    }
}

} // namespace vx3b
