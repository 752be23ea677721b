// Voxel-voxel contact for worlds too big for k_resident (BASELINE configs[3]: many bodies, self-collision on): a uniform-grid
// broad phase in global memory with per-voxel candidate lists (Verlet lists with a skin), then the narrow phase over the lists.
//
// The reference tests all Ns^2/2 surface pairs every step (gpu_update_attach, VX3/VX3_VoxelyzeKernel.cu:833-843, 682-727);
// k_contact (kernels_stream.cuh) restates that with shared-memory tiles and stays as the fallback and cross-check.  Here:
//   check   every step, inside the voxel pass (k_fused) or right after it (k_cgrid_check, two-kernel path): has a surface voxel moved
//           more than skin/2 (relative to the world's common drift) since its lists were built?  -> the world rebuilds at the next step;
//   insert  (rebuild steps only) hashed cell buckets as linked lists (atomicExch push); bucket heads carry the epoch of the rebuild,
//           so a new rebuild needs no clearing pass;
//   narrow  (k_cgrid_narrow) on rebuild steps each surface voxel first walks the 27 cells around it and lists every surface voxel
//           within (largest watch distance + skin) that it may collide with (not a lattice neighbour, not both Fixed), sorted
//           ascending so that force sums keep the all-pairs order; every step: the reference's exact double-precision predicate on
//           the listed candidates only; both members of a pair evaluate it in the same orientation (voxel1 = the later surface
//           voxel), so the forces are equal and opposite without atomics and - candidate order being the all-pairs order -
//           BIT-IDENTICAL to k_contact's.
// Any pair the reference would watch at the current step is within the list radius of the rebuild positions (triangle
// inequality on the skin), so the watched-pair set is exactly the reference's.  A voxel whose list overflows takes its candidates
// from the grid directly, in the same order.  Same scheme as k_resident's shared-memory version, sized for 1e5+ surface voxels.
// Per time step: insert (exits at once when no world rebuilds), narrow, then the step kernel and k_world_step - 4 launches with
// k_fused (8 in round 1).
#pragma once
#include "engine.cuh"
#include "kernels_stream.cuh"

namespace vx3b {

enum { CGRID_CAP = 32 }; // candidates per surface voxel (12 face + 8 body diagonals of a cubic lattice at rest, and slack)

struct ContactGrid {
    int s0;              // first surface slot of the streaming-only worlds (the kernels cover [s0, nsTot))
    unsigned long long *head; // [sum of buckets] epoch << 32 | first surface slot (world-local) of the bucket; another epoch than the world's = empty
    int2 *link;          // [nsTot] {next slot of the bucket's chain or -1, packed cell of the voxel at the last rebuild}
    float4 *ref;         // [nsTot] position at the last rebuild
    const short4 *slat;  // [nsTot] static: lattice index of the surface voxel, w = its material is Fixed.  A chain element is judged from
                         // link / ref / slat, three INDEPENDENT loads by surface slot (the list build used to chase five dependent ones
                         // per candidate: next -> key -> surface voxel -> lattice -> position; 80 us per rebuild on the 78,440-surface-voxel pile)
    int32_t *list;       // [nsTot][CGRID_CAP] world-local surface slots, ascending
    int32_t *count;      // [nsTot]; cap + 1 = the list overflowed: the voxel walks the grid itself
    int cap;             // entries per list in use (<= CGRID_CAP)
    int nBuckets;        // sum over the worlds (0 = no grid)
};

VXD uint32_t cgrid_key(double x, double y, double z, double invCell) {
    const int cx = (int)floor(x * invCell), cy = (int)floor(y * invCell), cz = (int)floor(z * invCell);
    return (uint32_t)(cx & 1023) | ((uint32_t)(cy & 1023) << 10) | ((uint32_t)(cz & 1023) << 20);
}
VXD uint32_t cgrid_hash(uint32_t key, int log2H) { return (key * 2654435761u) >> (32 - log2H); }
VXD bool cgrid_covers(const WorldD &W) { return W.active && (W.flags & WF_COLLISION) && !(W.flags & WF_RESIDENT) && W.gridLog2 > 0; }

template <class R> VXD bool cgrid_world(const Arrays<R> &A, int s, int &w, int &v, int &local) {
    v = A.surf[s];
    w = A.vworld[v];
    const WorldD &W = A.worlds[w];
    local = s - W.soff;
    return cgrid_covers(W);
}

// Stale lists?  One surface voxel: has it moved more than skin/2, relative to the drift `c` of the world's reference voxel, since
// the lists were built?  Any c common to the voxels of one check serves (|d_i - d_j| <= |d_i - c| + |d_j - c|); the callers take
// the reference voxel's position from the buffer set that no thread of the running kernel writes.  The answer is for the NEXT step
// (W.gRebuildNext; the world's end-of-step bookkeeping turns it into gRebuild and opens a new epoch of the bucket heads).
VXD void cgrid_check_one(WorldD &W, const ContactGrid &G, int s, double x, double y, double z, double mx, double my, double mz) {
    const int sm = W.soff + (W.ns >> 1);
    const float4 r = G.ref[s], rm = G.ref[sm];
    const float dx = ((float)x - r.x) - ((float)mx - rm.x), dy = ((float)y - r.y) - ((float)my - rm.y), dz = ((float)z - r.z) - ((float)mz - rm.z);
    if (dx * dx + dy * dy + dz * dz > W.skinHalf2)
        W.gRebuildNext = 1; // benign race: everybody writes 1
}
// the streaming (two-kernel) path: after k_voxels, before k_world_step
template <class R> __global__ void __launch_bounds__(256) k_cgrid_check(Arrays<R> A, ContactGrid G) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.nsTot)
        return;
    int w, v, i;
    if (!cgrid_world(A, s, w, v, i))
        return;
    WorldD &W = A.worlds[w];
    const StateSet<R> cur = cur_set(A, W.parity);
    const int vm = A.surf[W.soff + (W.ns >> 1)];
    const D2 a = cur.pxy[v], b = cur.pzt[v], am = cur.pxy[vm], bm = cur.pzt[vm];
    cgrid_check_one(W, G, s, a.x, a.y, b.x, am.x, am.y, bm.x);
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
    const StateSet<R> cur = contact_set(A, W);
    const D2 a = cur.pxy[v], b = cur.pzt[v];
    const uint32_t key = cgrid_key(a.x, a.y, b.x, W.invCell);
    G.ref[s] = make_float4((float)a.x, (float)a.y, (float)b.x, 0.f);
    const unsigned long long epoch = (unsigned long long)(uint32_t)W.gEpoch << 32;
    const unsigned long long old = atomicExch(&G.head[W.gridOff + cgrid_hash(key, W.gridLog2)], epoch | (unsigned long long)(uint32_t)i);
    G.link[s] = make_int2((old >> 32) == (epoch >> 32) ? (int)(uint32_t)old : -1, (int)key);
    if (i == 0)
        W.contactRebuilds++;
}
// first slot of a bucket's chain in the world's current epoch, or -1
VXD int cgrid_first(const WorldD &W, const ContactGrid &G, uint32_t k2) {
    const unsigned long long h = G.head[W.gridOff + cgrid_hash(k2, W.gridLog2)];
    return (uint32_t)(h >> 32) == (uint32_t)W.gEpoch ? (int)(uint32_t)h : -1;
}
// Every step, one thread per surface voxel: (only when the world's lists are stale) walk the 27 cells around the voxel and list,
// ascending, every surface voxel within (largest watch distance + skin) it may collide with; then the forces from the listed
// candidates.  A voxel whose list overflowed walks the grid itself every step, taking the candidates in ascending order too (the
// cells are those of the last rebuild: whatever is within the watch distance now was within the list radius then).
template <class R> __global__ void __launch_bounds__(128) k_cgrid_narrow(Arrays<R> A, ContactGrid G, int recordPairs) {
    const int s = G.s0 + blockIdx.x * blockDim.x + threadIdx.x;
    int hits = 0, w = -1, v = 0, i = 0;
    if (s >= A.nsTot || !cgrid_world(A, s, w, v, i))
        return;
    WorldD &W = A.worlds[w];
    if (!W.gRebuild && G.count[s] == 0) { // most surface voxels of a scene touch nothing: no candidates, no force, nothing else to load
        A.contactF[s] = mk4<R>(R(0), R(0), R(0), R(0));
        return;
    }
    const StateSet<R> cur = contact_set(A, W);
    const short4 c1 = G.slat[s];
    const DVoxMat &m1 = A.vmats[A.vmat[v]];
    const uint32_t key = (uint32_t)G.link[s].y;
    const int cx = key & 1023, cy = (key >> 10) & 1023, cz = key >> 20;
    int32_t *list = G.list + (size_t)s * CGRID_CAP;
    auto cell_key = [&](int c) {
        const int dx = c % 3 - 1, dy = (c / 3) % 3 - 1, dz = c / 9 - 1;
        return (uint32_t)((cx + dx) & 1023) | ((uint32_t)((cy + dy) & 1023) << 10) | ((uint32_t)((cz + dz) & 1023) << 20);
    };
    auto may_collide = [&](int j) { // not itself, not a lattice neighbour (linked: is_neighbor depth 1), not both Fixed
        if (j == i)
            return false;
        const short4 c2 = G.slat[W.soff + j];
        return !(abs(c1.x - c2.x) + abs(c1.y - c2.y) + abs(c1.z - c2.z) == 1 || (c1.w && c2.w));
    };
    if (W.gRebuild) {
        // positions as float, as stored at this rebuild, against the list radius padded by 0.1 %: a superset of the exact list (every
        // listed candidate still goes through the reference's double-precision predicate below, in ascending order)
        const float listR2 = (float)W.listR2 * 1.002f;
        const float4 r1 = G.ref[s];
        int n = 0;
        bool overflow = false;
        int first[27]; // the 27 bucket heads first: independent loads, instead of one more hop at the start of every cell's chain
#pragma unroll
        for (int c = 0; c < 27; c++)
            first[c] = cgrid_first(W, G, cell_key(c));
#pragma unroll
        for (int c = 0; c < 27; c++) {
            if (overflow)
                break;
            const uint32_t k2 = cell_key(c);
            for (int j = first[c]; j >= 0;) {
                const int2 lk = G.link[W.soff + j];
                const float4 r2 = G.ref[W.soff + j];
                const int jj = j;
                j = lk.x;
                if ((uint32_t)lk.y != k2 || !may_collide(jj))
                    continue;
                const float ex = r1.x - r2.x, ey = r1.y - r2.y, ez = r1.z - r2.z;
                if (ex * ex + ey * ey + ez * ez > listR2)
                    continue;
                if (n == G.cap) {
                    overflow = true;
                    break;
                }
                int p = n++; // insertion sort: ascending surface slot = the order in which the all-pairs scan meets them
                while (p > 0 && list[p - 1] > jj) {
                    list[p] = list[p - 1];
                    p--;
                }
                list[p] = jj;
            }
        }
        G.count[s] = overflow ? G.cap + 1 : n;
        if (overflow)
            atomicAdd(&W.contactOverflows, 1);
    }
    const bool rec = recordPairs && W.runSteps == 1;
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
    const double watchR2 = W.watchR2;
    V3<R> acc{R(0), R(0), R(0)};
    auto pair = [&](int j) {
        const int v2 = A.surf[W.soff + j];
        V3<double> p2;
        V3<R> l2;
        float t2;
        load(v2, p2, l2, t2);
        if (watchR2 > 0) { // constant voxel sizes: reject on the squared distance alone
            const double ex = p1.x - p2.x, ey = p1.y - p2.y, ez = p1.z - p2.z;
            if (ex * ex + ey * ey + ez * ez > watchR2)
                return;
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
    };
    const int n = G.count[s];
    if (n <= G.cap) {
        for (int k = 0; k < n; k++)
            pair(list[k]);
    } else { // overflowed list: the candidates straight from the grid, smallest slot first (selection: rare, and exact)
        for (int last = -1;;) {
            int best = 0x7fffffff;
            for (int c = 0; c < 27; c++) {
                const uint32_t k2 = cell_key(c);
                for (int j = cgrid_first(W, G, k2); j >= 0;) {
                    const int2 lk = G.link[W.soff + j];
                    if (j > last && j < best && (uint32_t)lk.y == k2 && may_collide(j))
                        best = j;
                    j = lk.x;
                }
            }
            if (best == 0x7fffffff)
                break;
            pair(best);
            last = best;
        }
    }
    A.contactF[s] = mk4<R>(acc.x, acc.y, acc.z, R(0));
    if (hits)
        atomicAdd(&W.collisionCount, hits);
}

} // namespace vx3b
