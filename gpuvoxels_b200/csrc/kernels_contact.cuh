// Voxel-voxel contact for worlds too big for k_resident (BASELINE configs[3]: many bodies, self-collision on): a uniform-grid
// broad phase in global memory with per-voxel candidate lists (Verlet lists with a skin), then the narrow phase over the lists.
//
// The reference tests all Ns^2/2 surface pairs every step (gpu_update_attach, VX3/VX3_VoxelyzeKernel.cu:833-843, 682-727);
// k_contact (kernels_stream.cuh) restates that with shared-memory tiles and stays as the fallback and cross-check.  Here:
//   check   every step, one thread per surface voxel: has it moved more than skin/2 (relative to the world's common drift) since
//           its lists were built?  -> the world rebuilds this step;
//   rebuild (only then) hashed cell buckets as linked lists (atomicExch push), then each surface voxel walks the 27 cells around
//           it and lists every surface voxel within (largest watch distance + skin) that it may collide with (not a lattice
//           neighbour, not both Fixed), sorted ascending so that force sums keep the all-pairs order;
//   narrow  every step: the reference's exact double-precision predicate on the listed candidates only; both members of a pair
//           evaluate it in the same orientation (voxel1 = the later surface voxel), so the forces are equal and opposite without
//           atomics and - candidate order being the all-pairs order - BIT-IDENTICAL to k_contact's.
// Any pair the reference would watch at the current step is within the list radius of the rebuild positions (triangle
// inequality on the skin), so the watched-pair set is exactly the reference's.  A list that overflows sends its world back to
// k_contact for that step.  Same scheme as k_resident's shared-memory version, sized for 1e5+ surface voxels.
#pragma once
#include "engine.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

enum { CGRID_CAP = 32 }; // candidates per surface voxel (12 face + 8 body diagonals of a cubic lattice at rest, and slack)

struct ContactGrid {
    int s0;              // first surface slot of the streaming-only worlds (the kernels cover [s0, nsTot))
    int32_t *head;       // [sum of buckets] first surface slot (world-local) of every bucket or -1
    int32_t *next;       // [nsTot] chain
    uint32_t *key;       // [nsTot] packed cell of the voxel at the last rebuild
    float4 *ref;         // [nsTot] position at the last rebuild
    int32_t *list;       // [nsTot][CGRID_CAP] world-local surface slots, ascending
    int32_t *count;      // [nsTot]
    const int2 *buckets; // work items of the clear kernel: (world, first bucket) per 1024 buckets
    int nBucketTiles;
};

VXD uint32_t cgrid_key(double x, double y, double z, double invCell) {
    const int cx = (int)floor(x * invCell), cy = (int)floor(y * invCell), cz = (int)floor(z * invCell);
    return (uint32_t)(cx & 1023) | ((uint32_t)(cy & 1023) << 10) | ((uint32_t)(cz & 1023) << 20);
}
VXD uint32_t cgrid_hash(uint32_t key, int log2H) { return (key * 2654435761u) >> (32 - log2H); }

template <class R> VXD bool cgrid_world(const Arrays<R> &A, int s, int &w, int &v, int &local) {
    v = A.surf[s];
    w = A.vworld[v];
    const WorldD &W = A.worlds[w];
    local = s - W.soff;
    return W.active && (W.flags & WF_COLLISION) && !(W.flags & WF_RESIDENT) && W.gridLog2 > 0;
}

// every step: stale lists?  (the first surface voxel's thread also opens the step: hit counting happens in k_cgrid_narrow)
template <class R> __global__ void __launch_bounds__(256) k_cgrid_check(Arrays<R> A, ContactGrid G) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.nsTot)
        return;
    int w, v, i;
    if (!cgrid_world(A, s, w, v, i))
        return;
    WorldD &W = A.worlds[w];
    if (W.gRebuild)
        return;
    const StateSet<R> cur = cur_set(A, W.parity);
    const int sm = W.soff + (W.ns >> 1), vm = A.surf[sm]; // common drift = that of one reference voxel
    const D2 a = cur.pxy[v], b = cur.pzt[v], am = cur.pxy[vm], bm = cur.pzt[vm];
    const float4 r = G.ref[s], rm = G.ref[sm];
    const float dx = ((float)a.x - r.x) - ((float)am.x - rm.x), dy = ((float)a.y - r.y) - ((float)am.y - rm.y), dz = ((float)b.x - r.z) - ((float)bm.x - rm.z);
    if (dx * dx + dy * dy + dz * dz > W.skinHalf2)
        W.gRebuild = 1; // benign race: everybody writes 1
}
template <class R> __global__ void __launch_bounds__(256) k_cgrid_clear(Arrays<R> A, ContactGrid G) {
    const int2 t = G.buckets[blockIdx.x];
    const WorldD &W = A.worlds[t.x];
    if (!W.active || !W.gRebuild || (W.flags & WF_RESIDENT) || W.gridLog2 == 0)
        return;
    const int H = 1 << W.gridLog2;
    for (int k = threadIdx.x; k < 1024; k += blockDim.x)
        if (t.y + k < H)
            G.head[W.gridOff + t.y + k] = -1;
}
template <class R> __global__ void __launch_bounds__(256) k_cgrid_insert(Arrays<R> A, ContactGrid G) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.nsTot)
        return;
    int w, v, i;
    if (!cgrid_world(A, s, w, v, i))
        return;
    WorldD &W = A.worlds[w];
    if (!W.gRebuild)
        return;
    const StateSet<R> cur = cur_set(A, W.parity);
    const D2 a = cur.pxy[v], b = cur.pzt[v];
    const uint32_t key = cgrid_key(a.x, a.y, b.x, W.invCell);
    G.key[s] = key;
    G.ref[s] = make_float4((float)a.x, (float)a.y, (float)b.x, 0.f);
    G.next[s] = atomicExch(&G.head[W.gridOff + cgrid_hash(key, W.gridLog2)], i);
    if (i == 0) {
        W.contactRebuilds++;
        W.gOverflow = 0;
    }
}
template <class R> __global__ void __launch_bounds__(128) k_cgrid_lists(Arrays<R> A, ContactGrid G) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.nsTot)
        return;
    int w, v, i;
    if (!cgrid_world(A, s, w, v, i))
        return;
    WorldD &W = A.worlds[w];
    if (!W.gRebuild)
        return;
    const StateSet<R> cur = cur_set(A, W.parity);
    const D2 a = cur.pxy[v], b = cur.pzt[v];
    const short4 c1 = A.lat[v];
    const bool fixed1 = A.vmats[A.vmat[v]].fixed != 0;
    const uint32_t key = G.key[s];
    const int cx = key & 1023, cy = (key >> 10) & 1023, cz = key >> 20;
    const double listR2 = W.listR2;
    int32_t *list = G.list + (size_t)s * CGRID_CAP;
    int n = 0;
    bool overflow = false;
    for (int c = 0; c < 27; c++) {
        const int dx = c % 3 - 1, dy = (c / 3) % 3 - 1, dz = c / 9 - 1;
        const uint32_t k2 = (uint32_t)((cx + dx) & 1023) | ((uint32_t)((cy + dy) & 1023) << 10) | ((uint32_t)((cz + dz) & 1023) << 20);
        for (int j = G.head[W.gridOff + cgrid_hash(k2, W.gridLog2)]; j >= 0; j = G.next[W.soff + j]) {
            if (j == i || G.key[W.soff + j] != k2)
                continue;
            const int v2 = A.surf[W.soff + j];
            const short4 c2 = A.lat[v2];
            if (abs(c1.x - c2.x) + abs(c1.y - c2.y) + abs(c1.z - c2.z) == 1 || (fixed1 && A.vmats[A.vmat[v2]].fixed))
                continue; // lattice neighbours are linked and never collide (is_neighbor depth 1); two Fixed voxels are skipped
            const D2 a2 = cur.pxy[v2], b2 = cur.pzt[v2];
            const double ex = a.x - a2.x, ey = a.y - a2.y, ez = b.x - b2.x;
            if (ex * ex + ey * ey + ez * ez > listR2)
                continue;
            if (n == CGRID_CAP) {
                overflow = true;
                continue;
            }
            int p = n++; // insertion sort: ascending surface slot = the order in which the all-pairs scan meets them
            while (p > 0 && list[p - 1] > j) {
                list[p] = list[p - 1];
                p--;
            }
            list[p] = j;
        }
    }
    G.count[s] = n;
    if (overflow) {
        W.gOverflow = 1; // this world takes the all-pairs kernel this step and rebuilds again at the next
        W.contactOverflows++;
    }
}
// every step: forces from the listed candidates.  Worlds whose lists overflowed are left to k_contact (launched after this).
template <class R> __global__ void __launch_bounds__(128) k_cgrid_narrow(Arrays<R> A, ContactGrid G, int recordPairs) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    int hits = 0, w = -1, v = 0, i = 0;
    bool mine = s < A.nsTot && cgrid_world(A, s, w, v, i);
    if (mine && A.worlds[w].gOverflow)
        mine = false;
    if (mine) {
        WorldD &W = A.worlds[w];
        const bool rec = recordPairs && W.runSteps == 1;
        const StateSet<R> cur = cur_set(A, W.parity);
        auto load = [&](int vv, V3<double> &p, V3<R> &l, float &temp) {
            const D2 pa = cur.pxy[vv], pb = cur.pzt[vv];
            temp = u2f((uint32_t)((unsigned long long)__double_as_longlong(pb.y) & 0xffffffffu));
            p = {pa.x, pa.y, pb.x};
            const auto lm = cur.lm[vv];
            l = {lm.x, lm.y, lm.z};
        };
        V3<double> p1;
        V3<R> l1;
        float t1;
        load(v, p1, l1, t1);
        const DVoxMat &m1 = A.vmats[A.vmat[v]];
        const double watchR2 = W.watchR2;
        V3<R> acc{R(0), R(0), R(0)};
        const int n = G.count[s];
        const int32_t *list = G.list + (size_t)s * CGRID_CAP;
        for (int k = 0; k < n; k++) {
            const int j = list[k], v2 = A.surf[W.soff + j];
            V3<double> p2;
            V3<R> l2;
            float t2;
            load(v2, p2, l2, t2);
            if (watchR2 > 0) { // constant voxel sizes: reject on the squared distance alone
                const double ex = p1.x - p2.x, ey = p1.y - p2.y, ez = p1.z - p2.z;
                if (ex * ex + ey * ey + ez * ez > watchR2)
                    continue;
            }
            const DVoxMat &m2 = A.vmats[A.vmat[v2]];
            V3<R> f;
            if (i > j) {
                if (contact_pair<R>(m1, t1, p1, l1, m2, t2, p2, l2, f)) {
                    acc = acc + f;
                    if ((m1.isTarget != 0) != (m2.isTarget != 0))
                        hits++;
                    if (rec) {
                        int slot = atomicAdd(A.pairCount, 1);
                        if (slot < A.pairCap)
                            A.pairBuf[slot] = make_int4(w, v - W.voff, v2 - W.voff, 0);
                    }
                }
            } else if (contact_pair<R>(m2, t2, p2, l2, m1, t1, p1, l1, f))
                acc = acc - f;
        }
        A.contactF[s] = mk4<R>(acc.x, acc.y, acc.z, R(0));
        if (i == 0)
            W.gRebuild = 0; // lists are fresh (or still valid); k_cgrid_check of the next step decides again
        if (hits)
            atomicAdd(&W.collisionCount, hits);
    }
}

} // namespace vx3b
