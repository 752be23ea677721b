// Host side of the dynamic-topology path (kernels_dynamic.cuh) and the multiplexer that lets one context hold both kinds of world:
// worlds with signals / attachment / detachment / removal / material switches go to DynEngine, every other world to Engine<R>, and
// the C ABI's world indices are mapped to the right side.  A batch without dynamic worlds passes straight through.
// Reference: VX3_SimulationManager::readVXD / startKernel / collectResults (VX3/VX3_SimulationManager.cu:277-458).
#pragma once
#include "engine_impl.cuh"
#include "kernels_dynamic.cuh"

namespace vx3b {

// which engine steps a world (VX3_SimulationManager.cu:325-371: the switches; VX3_VoxelyzeKernel.cu:632-634, 852-854: the material
// timers, which the static kernels do not look at)
inline bool world_is_dynamic(const VX3B_WorldDesc &d) {
    if (d.enable_signals || d.enable_attach || d.enable_detach || d.secondary_experiment)
        return true;
    for (int k = 0; k < d.num_voxel_mats; k++) {
        const VX3B_VoxelMaterial &m = d.voxel_mats[k];
        if (m.TurnOnThermalExpansionAfterThisManySeconds > 0 || m.TurnOnCiliaAfterThisManySeconds > 0)
            return true;
    }
    return false;
}

template <class R> class DynEngine {
    using T4 = typename Store<R>::T4;

  public:
    struct HostWorld {
        int nv, nl0, lcap, voff, loff, numMeasured;
    };
    DynArrays<R> D{};
    std::vector<HostWorld> H;
    std::vector<std::unique_ptr<DeviceBuf>> bufs; // (DeviceBuf owns its allocation and must not be copied by a growing vector)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int device = 0;
    float lastMs = 0;
    int64_t lastVoxelSteps = 0;
    std::vector<int64_t> stepsBefore;

    ~DynEngine() {
        release();
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
        if (stream)
            cudaStreamDestroy(stream);
    }
    void release() {
        if (stream)
            cudaStreamSynchronize(stream);
        bufs.clear();
        H.clear();
        memset(&D, 0, sizeof(D));
    }
    template <class T> int put(T **out, const std::vector<T> &src, size_t minCount = 1) {
        bufs.emplace_back(new DeviceBuf());
        DeviceBuf &b = *bufs.back();
        const size_t n = std::max(src.size(), minCount);
        CUDA_TRY(b.alloc(n * sizeof(T)));
        CUDA_TRY(cudaMemset(b.p, 0, n * sizeof(T)));
        if (!src.empty())
            CUDA_TRY(cudaMemcpy(b.p, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
        *out = (T *)b.p;
        return VX3B_OK;
    }
    template <class T> int zeros(T **out, size_t count) { return put(out, std::vector<T>(), std::max<size_t>(count, 1)); }

    int upload(const std::vector<const VX3B_WorldDesc *> &worlds, double dtOverride, int recordPairs) {
        CUDA_TRY(cudaSetDevice(device));
        release();
        if (!stream)
            CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        if (!ev0) {
            CUDA_TRY(cudaEventCreate(&ev0));
            CUDA_TRY(cudaEventCreate(&ev1));
        }
        const int n = (int)worlds.size();
        size_t nvTot = 0, lcapTot = 0, nvm = 0, nlm = 0, candTot = 0, next = 0;
        for (int w = 0; w < n; w++) {
            const VX3B_WorldDesc &d = *worlds[w];
            if (d.num_voxels <= 0 || !d.lattice || !d.voxel_mat || !d.voxel_mats || (d.num_links > 0 && (!d.link_neg || !d.link_pos || !d.link_axis || !d.link_mat || !d.link_mats))) {
                set_error("vx3b_batch_upload: a dynamic-topology world has no voxels or null arrays");
                return VX3B_ERR_ARG;
            }
            if (d.ghost) {
                set_error("vx3b_batch_upload: slabs (ghost voxels) cannot be combined with signals / attachment / detachment / removal");
                return VX3B_ERR_ARG;
            }
            for (int l = 0; l < d.num_links; l++)
                if (d.link_neg[l] < 0 || d.link_neg[l] >= d.num_voxels || d.link_pos[l] < 0 || d.link_pos[l] >= d.num_voxels || d.link_axis[l] > 2 ||
                    d.link_mat[l] < 0 || d.link_mat[l] >= d.num_link_mats) {
                    set_error("vx3b_batch_upload: a dynamic-topology world has a link with an end, axis or material out of range");
                    return VX3B_ERR_ARG;
                }
            for (int v = 0; v < d.num_voxels; v++)
                if (d.voxel_mat[v] < 0 || d.voxel_mat[v] >= d.num_voxel_mats) {
                    set_error("vx3b_batch_upload: a dynamic-topology world has a voxel material out of range");
                    return VX3B_ERR_ARG;
                }
            nvTot += d.num_voxels;
            // links only grow by attachment; a detached link keeps its index, so the pool is the initial links + every free slot pair
            lcapTot += (size_t)d.num_links + (d.enable_attach ? 3 * (size_t)d.num_voxels + 64 : 0);
            nvm += d.num_voxel_mats;
            nlm += d.num_link_mats;
            candTot += d.enable_attach ? (size_t)std::max(1024, d.num_voxels) : 1;
        }
        std::vector<DynWorld> hW(n);
        std::vector<P3> pos(nvTot), baseCilia(nvTot), shiftCilia(nvTot);
        std::vector<T4> orient(nvTot);
        std::vector<float> temp(nvTot);
        std::vector<uint32_t> vflags(nvTot, VF_STATIC_FRICTION); // CVX_Voxel::reset sets FLOOR_STATIC_FRICTION (old/VX_Voxel.cpp:47-56)
        std::vector<int32_t> vmat(nvTot), vlinks(nvTot * 6, -1), vext(nvTot, -1), targets;
        std::vector<short4> lat(nvTot);
        std::vector<double> phase(nvTot, 0.0);
        std::vector<uint8_t> vattach(nvTot, 1); // VX3_Voxel.h: enableAttach = true
        std::vector<int2> lends(lcapTot);
        std::vector<uint32_t> lmeta(lcapTot, 0), lflags(lcapTot, LF_SMALL_ANGLE);
        std::vector<DVoxMat> vmats(nvm);
        std::vector<DynMatX> matx(nvm);
        std::vector<DLinkMat> lmats(nlm);
        std::vector<DLinkModel> lmodels(nlm);
        std::vector<DExt> ext;
        bool anyPhase = false, anyCilia = false, anyShift = false, anyNonlin = false;
        size_t voff = 0, loff = 0, vmo = 0, lmo = 0, candOff = 0;
        H.resize(n);
        for (int w = 0; w < n; w++) {
            const VX3B_WorldDesc &d = *worlds[w];
            DynWorld &DW = hW[w];
            memset(&DW, 0, sizeof(DW));
            WorldD &W = DW.W;
            const bool vx3 = !d.vx2_compat;
            const bool vary = d.vary_temp_enabled && d.temp_period > 0;
            double maxAlpha = 0;
            bool canFail = false, nonlin = false;
            for (int k = 0; k < d.num_voxel_mats; k++)
                maxAlpha = std::max(maxAlpha, (double)fabsf(d.voxel_mats[k].alphaCTE));
            for (int k = 0; k < d.num_link_mats; k++) {
                canFail = canFail || d.link_mats[k].epsilonFail != -1.0f;
                nonlin = nonlin || link_material_is_nonlinear(d.link_mats[k]);
            }
            anyNonlin = anyNonlin || nonlin;
            const bool hasFF = !formula_empty(d.force_field[0]) || !formula_empty(d.force_field[1]) || !formula_empty(d.force_field[2]);
            const bool thermal = vary && maxAlpha != 0;
            W.flags = (d.floor_enabled ? WF_FLOOR : 0) | (vary ? WF_VARY_TEMP : 0) | (d.enable_collision ? WF_COLLISION : 0) | (d.enable_cilia ? WF_CILIA : 0) |
                      (d.enable_expansion ? WF_EXPANSION : 0) | (vx3 ? WF_VX3 : 0) | (hasFF ? WF_FORCE_FIELD : 0) |
                      (!formula_empty(d.stop_condition) ? WF_STOP_FORMULA : 0) | (thermal ? WF_THERMAL : 0) | (canFail ? WF_CAN_FAIL : 0) | (nonlin ? WF_NONLIN : 0);
            double opt = recommended_dt(d); // VX3_VoxelyzeKernel.cu:244-254
            if (vx3 && opt < 1e-10)
                opt = 1e-10;
            float dt = (float)(d.dt_frac * opt);
            if (dtOverride > 0)
                dt = (float)dtOverride;
            const int lcap = d.num_links + (d.enable_attach ? 3 * d.num_voxels + 64 : 0);
            W.voff = (int)voff, W.nv = d.num_voxels, W.loff = (int)loff, W.nl = d.num_links, W.soff = (int)voff, W.ns = 0;
            W.vmatOff = (int)vmo, W.lmatOff = (int)lmo, W.nvm = d.num_voxel_mats, W.nlm = d.num_link_mats;
            W.dt = dt;
            W.cycle = (dt > 0 && vx3) ? (int)(d.temp_period / dt) : 0;
            W.tempAmp = d.temp_amplitude, W.tempPeriod = d.temp_period, W.voxSize = d.vox_size;
            W.maxDistPair = d.max_dist_in_voxel_lengths_to_count_as_pair;
            memcpy(W.stop, d.stop_condition, sizeof(W.stop));
            memcpy(W.fitnessF, d.fitness, sizeof(W.fitnessF));
            memcpy(W.ff, d.force_field, sizeof(W.ff));
            memcpy(DW.attach, d.attach_condition, sizeof(DW.attach));
            DW.lcap = lcap;
            DW.surfaceChanged = 1; // CUDA_Simulation sets isSurfaceChanged before the first step (VX3_SimulationManager.cu:22)
            DW.safetyGuard = d.safety_guard;
            DW.enableSignals = d.enable_signals, DW.enableAttach = d.enable_attach, DW.enableDetach = d.enable_detach, DW.secondary = d.secondary_experiment;
            DW.reinitAfter = d.reinitialize_initial_position_after;
            { // contact candidate lists: everything within the largest watch distance + a skin of 0.4 voxels
                const double skin = 0.4 * d.vox_size, cell = max_watch_distance(d) + skin;
                W.listR2 = cell * cell;
                W.skinHalf2 = (float)(0.25 * skin * skin * (1.0 - 1e-2)); // a little early rather than late (float compare)
            }
            DW.candOff = (int)candOff;
            DW.candCap = d.enable_attach ? std::max(1024, d.num_voxels) : 1;
            DW.toff = (int)targets.size();
            for (int k = 0; k < d.num_voxel_mats; k++) {
                const VX3B_VoxelMaterial &m = d.voxel_mats[k];
                vmats[vmo + k] = make_dvoxmat(m, thermal, vary, d.initial_temp);
                DynMatX &x = matx[vmo + k];
                memset(&x, 0, sizeof(x));
                x.isPaceMaker = m.isPaceMaker, x.sticky = m.sticky, x.localIndex = k;
                x.paceMakerPeriod = m.PaceMakerPeriod, x.signalValueDecay = m.signalValueDecay, x.signalTimeDelay = m.signalTimeDelay;
                x.inactivePeriod = m.inactivePeriod, x.removeAfter = m.RemoveFromSimulationAfterThisManySeconds;
                x.thermalAfter = m.TurnOnThermalExpansionAfterThisManySeconds, x.ciliaAfter = m.TurnOnCiliaAfterThisManySeconds;
            }
            for (int k = 0; k < d.num_link_mats; k++) {
                lmats[lmo + k] = make_dlinkmat(d.link_mats[k], d.voxel_mats);
                lmodels[lmo + k] = make_dlinkmodel(d.link_mats[k]);
            }
            int measured = 0;
            for (int v = 0; v < d.num_voxels; v++) {
                const size_t g = voff + v;
                const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
                const int16_t *L = d.lattice + 3 * (size_t)v;
                pos[g] = P3{L[0] * m.nomSize, L[1] * m.nomSize, L[2] * m.nomSize};
                lat[g] = make_short4(L[0], L[1], L[2], 0);
                orient[g].x = R(1), orient[g].y = R(0), orient[g].z = R(0), orient[g].w = R(0); // w,x,y,z (mk4 is device code)
                temp[g] = d.initial_temp;
                vmat[g] = (int)vmo + d.voxel_mat[v];
                measured += m.isMeasured != 0;
                if (m.isTarget)
                    targets.push_back(v);
                if (d.phase_offset) {
                    phase[g] = d.phase_offset[v];
                    anyPhase = anyPhase || phase[g] != 0;
                }
                if (d.base_cilia) {
                    baseCilia[g] = P3{d.base_cilia[v * 3], d.base_cilia[v * 3 + 1], d.base_cilia[v * 3 + 2]};
                    anyCilia = true;
                }
                if (d.shift_cilia) {
                    shiftCilia[g] = P3{d.shift_cilia[v * 3], d.shift_cilia[v * 3 + 1], d.shift_cilia[v * 3 + 2]};
                    anyShift = true;
                }
                const uint8_t dof = d.dof_fixed ? d.dof_fixed[v] : 0;
                bool hasF = false;
                for (int k = 0; k < 3; k++)
                    hasF = hasF || (d.ext_force && d.ext_force[v * 3 + k] != 0) || (d.ext_moment && d.ext_moment[v * 3 + k] != 0);
                if (dof || hasF) {
                    DExt e;
                    memset(&e, 0, sizeof(e));
                    e.dof = dof;
                    for (int k = 0; k < 3; k++) {
                        e.force[k] = d.ext_force ? d.ext_force[v * 3 + k] : 0.0f;
                        e.moment[k] = d.ext_moment ? d.ext_moment[v * 3 + k] : 0.0f;
                        e.translation[k] = d.ext_translation ? d.ext_translation[v * 3 + k] : 0.0;
                    }
                    const double rx = d.ext_rotation ? d.ext_rotation[v * 3] : 0, ry = d.ext_rotation ? d.ext_rotation[v * 3 + 1] : 0,
                                 rz = d.ext_rotation ? d.ext_rotation[v * 3 + 2] : 0;
                    e.rotQ[0] = 1;
                    if (rx != 0 || ry != 0 || rz != 0) { // CVX_External::rotationChanged -> Quat3D(rotation vector)
                        const double hx = rx / 2, hy = ry / 2, hz = rz / 2, m2 = hx * hx + hy * hy + hz * hz;
                        double qw, s;
                        if (m2 * m2 < 5.328e-15) {
                            qw = 1.0 - 0.5 * m2;
                            s = 1.0 - m2 / 6.0;
                        } else {
                            const double mm = sqrt(m2);
                            qw = cos(mm);
                            s = sin(mm) / mm;
                        }
                        e.rotQ[0] = qw, e.rotQ[1] = hx * s, e.rotQ[2] = hy * s, e.rotQ[3] = hz * s;
                    }
                    vext[g] = (int)ext.size();
                    ext.push_back(e);
                }
            }
            DW.nTargets = (int)targets.size() - DW.toff;
            for (int l = 0; l < d.num_links; l++) { // creation order kept: a new link takes the next index, as in the reference's d_v_links
                const size_t g = loff + l;
                const int a = d.link_neg[l], b = d.link_pos[l], ax = d.link_axis[l];
                lends[g] = make_int2(a, b);
                lmeta[g] = (uint32_t)ax | ((uint32_t)(lmo + d.link_mat[l]) << 2);
                vlinks[(voff + a) * 6 + 2 * ax] = l;     // positive-side slot of the negative end
                vlinks[(voff + b) * 6 + 2 * ax + 1] = l; // negative-side slot of the positive end
            }
            H[w] = HostWorld{d.num_voxels, d.num_links, lcap, (int)voff, (int)loff, measured};
            voff += d.num_voxels;
            loff += lcap;
            vmo += d.num_voxel_mats;
            lmo += d.num_link_mats;
            candOff += DW.candCap;
            next += 0;
        }
        (void)next;
        // initial centre of mass (VX3_VoxelyzeKernel.cu:477-493), in voxel order as the reference sums it
        for (int w = 0; w < n; w++) {
            WorldD &W = hW[w].W;
            double sx = 0, sy = 0, sz = 0, sm = 0;
            for (int v = 0; v < W.nv; v++) {
                const DVoxMat &m = vmats[vmat[W.voff + v]];
                if (!m.isMeasured)
                    continue;
                const double mass = m.mass;
                sx += pos[W.voff + v].x * mass, sy += pos[W.voff + v].y * mass, sz += pos[W.voff + v].z * mass, sm += mass;
            }
            if (sm != 0) {
                const double inv = 1.0 / sm; // Vec3D::operator/ multiplies by the reciprocal (Utils/VX3_Vec3D.h:42)
                W.com[0] = inv * sx, W.com[1] = inv * sy, W.com[2] = inv * sz;
            }
            for (int k = 0; k < 3; k++)
                W.initialCom[k] = W.finalCom[k] = W.com[k];
            W.numMeasured = H[w].numMeasured;
        }
        D.nWorlds = n;
        int rc;
#define PUT(field, vec)                                                                                                         \
    if ((rc = put(&D.field, vec)) != VX3B_OK)                                                                                   \
        return rc;
#define ZEROS(field, count)                                                                                                     \
    if ((rc = zeros(&D.field, count)) != VX3B_OK)                                                                               \
        return rc;
        PUT(worlds, hW);
        PUT(pos, pos);
        PUT(ipos, pos);
        PUT(orient, orient);
        ZEROS(linMom, nvTot);
        ZEROS(angMom, nvTot);
        ZEROS(contactF, nvTot);
        PUT(temp, temp);
        PUT(vflags, vflags);
        PUT(vmat, vmat);
        PUT(lat, lat);
        PUT(vlinks, vlinks);
        PUT(vext, vext);
        if (anyPhase)
            PUT(phase, phase);
        if (anyCilia)
            PUT(baseCilia, baseCilia);
        if (anyShift)
            PUT(shiftCilia, shiftCilia);
        ZEROS(sig, nvTot * 6);
        ZEROS(sigT, nvTot * 4);
        ZEROS(vremoved, nvTot);
        PUT(vattach, vattach);
        ZEROS(vhit, nvTot);
        ZEROS(vsurfFlag, nvTot);
        ZEROS(surf, nvTot);
        ZEROS(clist, nvTot * DYN_LIST_CAP);
        ZEROS(ccount, nvTot);
        ZEROS(cref, nvTot);
        PUT(targets, targets);
        ZEROS(rowScratch, nvTot);
        PUT(lends, lends);
        PUT(lmeta, lmeta);
        ZEROS(lpos2, lcapTot);
        ZEROS(la1, lcapTot);
        ZEROS(la2, lcapTot);
        ZEROS(lfN, lcapTot);
        ZEROS(lfP, lcapTot);
        ZEROS(lmN, lcapTot);
        ZEROS(lmP, lcapTot);
        PUT(lflags, lflags);
        ZEROS(lmaxStrain, lcapTot);
        ZEROS(lstrain, lcapTot);
        if (anyNonlin)
            ZEROS(lso, lcapTot);
        ZEROS(lremoved, lcapTot);
        ZEROS(ldetached, lcapTot);
        ZEROS(lnew, lcapTot);
        PUT(vmats, vmats);
        PUT(matx, matx);
        ZEROS(matRemoved, nvm);
        PUT(lmats, lmats);
        PUT(lmodels, lmodels);
        PUT(ext, ext);
        ZEROS(cand, candTot);
        (void)recordPairs;
        D.pairCap = 1 << 18; // (the option may be set after the upload)
        ZEROS(pairBuf, (size_t)D.pairCap);
        ZEROS(pairCount, 1);
#undef PUT
#undef ZEROS
        return VX3B_OK;
    }

    int read_worlds(std::vector<DynWorld> &out) {
        out.resize(H.size());
        CUDA_TRY(cudaMemcpy(out.data(), D.worlds, H.size() * sizeof(DynWorld), cudaMemcpyDeviceToHost));
        return VX3B_OK;
    }
    int read_steps(std::vector<int64_t> &out) {
        out.resize(H.size());
        CUDA_TRY(cudaMemcpy2D(out.data(), sizeof(int64_t), (const char *)D.worlds + offsetof(DynWorld, W) + offsetof(WorldD, step), sizeof(DynWorld),
                              sizeof(int64_t), H.size(), cudaMemcpyDeviceToHost));
        return VX3B_OK;
    }
    // enqueue the whole run (one launch, one CTA per world) on this engine's own stream; run_wait() joins it
    int run_launch(int64_t maxSteps, int useStop, int recordPairs) {
        CUDA_TRY(cudaSetDevice(device));
        int rc = read_steps(stepsBefore);
        if (rc != VX3B_OK)
            return rc;
        if (recordPairs)
            CUDA_TRY(cudaMemsetAsync(D.pairCount, 0, sizeof(int), stream));
        CUDA_TRY(cudaEventRecord(ev0, stream));
        k_dynamic<R><<<(unsigned)H.size(), DYN_THREADS, 0, stream>>>(D, (long long)maxSteps, useStop, recordPairs && D.pairCap > 1 ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(ev1, stream));
        return VX3B_OK;
    }
    int run_wait() {
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventElapsedTime(&lastMs, ev0, ev1));
        std::vector<DynWorld> w;
        int rc = read_worlds(w);
        if (rc != VX3B_OK)
            return rc;
        lastVoxelSteps = 0;
        for (size_t k = 0; k < H.size(); k++) {
            lastVoxelSteps += (w[k].W.step - stepsBefore[k]) * (int64_t)H[k].nv;
            if (w[k].error) {
                set_error(w[k].error == DYN_ERR_LINK_POOL ? "dynamic-topology world: the link pool is exhausted (attachment / detachment cycles)"
                                                          : "dynamic-topology world: more attachment candidates in one step than the candidate buffer holds");
                return VX3B_ERR_STATE;
            }
        }
        return VX3B_OK;
    }

    int result(int k, VX3B_WorldResult &r) {
        DynWorld dw;
        CUDA_TRY(cudaMemcpy(&dw, D.worlds + k, sizeof(DynWorld), cudaMemcpyDeviceToHost));
        const WorldD &W = dw.W;
        memset(&r, 0, sizeof(r));
        r.current_time = W.t;
        r.fitness_score = W.fitness;
        for (int i = 0; i < 3; i++) {
            r.initial_com[i] = W.initialCom[i];
            r.current_com[i] = W.finalCom[i];
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
        r.num_measured_voxel = H[k].numMeasured;
        return VX3B_OK;
    }
    template <class T> int fetch(std::vector<T> &out, const T *src, size_t count) {
        out.resize(count);
        if (count)
            CUDA_TRY(cudaMemcpy(out.data(), src, count * sizeof(T), cudaMemcpyDeviceToHost));
        return VX3B_OK;
    }
    int current_links(int k, int *nl) {
        int32_t v;
        CUDA_TRY(cudaMemcpy(&v, (const char *)(D.worlds + k) + offsetof(DynWorld, W) + offsetof(WorldD, nl), sizeof(int32_t), cudaMemcpyDeviceToHost));
        *nl = v;
        return VX3B_OK;
    }
    int read_state(int k, double *pos, double *orient, double *lin, double *ang) {
        const HostWorld &h = H[k];
        std::vector<P3> p;
        std::vector<T4> o;
        std::vector<V3<R>> l, a;
        int rc;
        if ((rc = fetch(p, D.pos + h.voff, h.nv)) || (rc = fetch(o, D.orient + h.voff, h.nv)) || (rc = fetch(l, D.linMom + h.voff, h.nv)) ||
            (rc = fetch(a, D.angMom + h.voff, h.nv)))
            return rc;
        for (int v = 0; v < h.nv; v++) {
            if (pos)
                pos[v * 3] = p[v].x, pos[v * 3 + 1] = p[v].y, pos[v * 3 + 2] = p[v].z;
            if (orient)
                orient[v * 4] = o[v].x, orient[v * 4 + 1] = o[v].y, orient[v * 4 + 2] = o[v].z, orient[v * 4 + 3] = o[v].w;
            if (lin)
                lin[v * 3] = l[v].x, lin[v * 3 + 1] = l[v].y, lin[v * 3 + 2] = l[v].z;
            if (ang)
                ang[v * 3] = a[v].x, ang[v * 3 + 1] = a[v].y, ang[v * 3 + 2] = a[v].z;
        }
        return VX3B_OK;
    }
    int read_initial_pos(int k, double *pos) {
        std::vector<P3> p;
        int rc = fetch(p, D.ipos + H[k].voff, H[k].nv);
        if (rc)
            return rc;
        for (int v = 0; v < H[k].nv; v++)
            pos[v * 3] = p[v].x, pos[v * 3 + 1] = p[v].y, pos[v * 3 + 2] = p[v].z;
        return VX3B_OK;
    }
    int read_links(int k, double *pos2, double *a1, double *a2, int32_t *flags) {
        int nl, rc;
        if ((rc = current_links(k, &nl)))
            return rc;
        const size_t lo = H[k].loff;
        std::vector<V3<R>> p, x, y;
        std::vector<uint32_t> f;
        if ((rc = fetch(p, D.lpos2 + lo, nl)) || (rc = fetch(x, D.la1 + lo, nl)) || (rc = fetch(y, D.la2 + lo, nl)) || (rc = fetch(f, D.lflags + lo, nl)))
            return rc;
        for (int l = 0; l < nl; l++) {
            if (pos2)
                pos2[l * 3] = p[l].x, pos2[l * 3 + 1] = p[l].y, pos2[l * 3 + 2] = p[l].z;
            if (a1)
                a1[l * 3] = x[l].x, a1[l * 3 + 1] = x[l].y, a1[l * 3 + 2] = x[l].z;
            if (a2)
                a2[l * 3] = y[l].x, a2[l * 3 + 1] = y[l].y, a2[l * 3 + 2] = y[l].z;
            if (flags)
                flags[l] = (int32_t)(f[l] & 3u);
        }
        return VX3B_OK;
    }
    int read_aux(int k, float *voxelTemp, float *linkStrain) {
        int nl, rc;
        if ((rc = current_links(k, &nl)))
            return rc;
        if (voxelTemp)
            CUDA_TRY(cudaMemcpy(voxelTemp, D.temp + H[k].voff, H[k].nv * sizeof(float), cudaMemcpyDeviceToHost));
        if (linkStrain && nl)
            CUDA_TRY(cudaMemcpy(linkStrain, D.lstrain + H[k].loff, nl * sizeof(float), cudaMemcpyDeviceToHost));
        return VX3B_OK;
    }
    int read_signals(int k, double *sig6, int32_t *vflags2) {
        const HostWorld &h = H[k];
        if (sig6)
            CUDA_TRY(cudaMemcpy(sig6, D.sig + (size_t)h.voff * 6, (size_t)h.nv * 6 * sizeof(double), cudaMemcpyDeviceToHost));
        if (vflags2) {
            std::vector<uint8_t> r, a;
            int rc;
            if ((rc = fetch(r, D.vremoved + h.voff, h.nv)) || (rc = fetch(a, D.vattach + h.voff, h.nv)))
                return rc;
            for (int v = 0; v < h.nv; v++)
                vflags2[2 * v] = r[v], vflags2[2 * v + 1] = a[v];
        }
        return VX3B_OK;
    }
    int read_topology(int k, int32_t *links6, int32_t *n) {
        int nl, rc;
        if ((rc = current_links(k, &nl)))
            return rc;
        const int cap = *n;
        *n = nl;
        if (!links6)
            return VX3B_OK;
        const int m = std::min(cap, nl);
        const size_t lo = H[k].loff;
        std::vector<int2> e;
        std::vector<uint32_t> meta;
        std::vector<uint8_t> rem, det;
        std::vector<int32_t> nw;
        if ((rc = fetch(e, D.lends + lo, m)) || (rc = fetch(meta, D.lmeta + lo, m)) || (rc = fetch(rem, D.lremoved + lo, m)) ||
            (rc = fetch(det, D.ldetached + lo, m)) || (rc = fetch(nw, D.lnew + lo, m)))
            return rc;
        for (int l = 0; l < m; l++) {
            int32_t *q = links6 + 6 * (size_t)l;
            q[0] = e[l].x, q[1] = e[l].y, q[2] = (int32_t)(meta[l] & 3u), q[3] = rem[l], q[4] = det[l], q[5] = nw[l];
        }
        return VX3B_OK;
    }
    int collision_pairs(int k, int32_t *pairs, int32_t *n) {
        int cnt = 0;
        CUDA_TRY(cudaMemcpy(&cnt, D.pairCount, sizeof(int), cudaMemcpyDeviceToHost));
        cnt = std::min(cnt, D.pairCap);
        std::vector<int4> buf;
        int rc = fetch(buf, D.pairBuf, (size_t)cnt);
        if (rc)
            return rc;
        std::vector<std::pair<int, int>> mine;
        for (auto &p : buf)
            if (p.x == k)
                mine.emplace_back(p.y, p.z);
        std::sort(mine.begin(), mine.end());
        const int cap = *n;
        *n = (int)mine.size();
        if (pairs)
            for (int i = 0; i < (int)mine.size() && i < cap; i++)
                pairs[2 * i] = mine[i].first, pairs[2 * i + 1] = mine[i].second;
        return VX3B_OK;
    }
};

// One context = the static engine (every kernel path of engine_impl.cuh) + the dynamic-topology engine; world indices of the C ABI
// are those of the uploaded batch.
template <class R> class EngineMux : public EngineBase {
  public:
    Engine<R> S;
    DynEngine<R> Dy;
    struct Where {
        bool dyn;
        int idx;
    };
    std::vector<Where> where; // empty: no dynamic world in the batch, indices pass straight through
    std::vector<VX3B_WorldDesc> staticDescs;
    int nStatic = 0, nDyn = 0;

    void push_options() {
        S.device = device, S.precision = precision;
        if (stream && S.stream != stream) { // a caller-provided stream (vx3b_ctx_set_stream)
            if (S.ownStream && S.stream)
                cudaStreamDestroy(S.stream);
            S.stream = stream;
            S.ownStream = false;
        }
        if (S.requestedPath < 0)
            S.path = path;
        else
            S.requestedPath = path;
        for (int k = 0; k < 3; k++)
            S.brickDim[k] = brickDim[k];
        S.recordPairs = recordPairs, S.chunk = chunk, S.dtOverride = dtOverride, S.residentThreads = residentThreads, S.contactSkin = contactSkin;
        S.contactListCap = contactListCap, S.extAll = extAll, S.contactGrid = contactGrid, S.slabFused = slabFused;
        Dy.device = device;
    }
    bool locate(int world, Where &w) {
        const int n = where.empty() ? nStatic : (int)where.size();
        if (world < 0 || world >= n) {
            set_error("world index out of range");
            return false;
        }
        w = where.empty() ? Where{false, world} : where[world];
        return true;
    }
    int no_dynamic(const char *what) {
        set_error(std::string(what) + " is not available for a world stepped by the dynamic-topology kernel (signals / attachment / detachment / removal)");
        return VX3B_ERR_ARG;
    }

    int upload(const VX3B_WorldDesc *worlds, int n) override {
        push_options();
        where.clear();
        staticDescs.clear();
        nStatic = nDyn = 0;
        if (n <= 0 || !worlds)
            return S.upload(worlds, n); // (reports the error)
        bool any = false;
        for (int w = 0; w < n && !any; w++)
            any = world_is_dynamic(worlds[w]);
        if (!any) {
            nStatic = n;
            Dy.release();
            return S.upload(worlds, n);
        }
        std::vector<const VX3B_WorldDesc *> dyn;
        where.resize(n);
        for (int w = 0; w < n; w++) {
            if (world_is_dynamic(worlds[w])) {
                if (worlds[w].vx2_compat) {
                    set_error("vx3b_batch_upload: signals / attachment / detachment / removal are Voxelyze3 features (vx2_compat must be 0)");
                    return VX3B_ERR_ARG;
                }
                where[w] = Where{true, (int)dyn.size()};
                dyn.push_back(&worlds[w]);
            } else {
                where[w] = Where{false, (int)staticDescs.size()};
                staticDescs.push_back(worlds[w]);
            }
        }
        nStatic = (int)staticDescs.size(), nDyn = (int)dyn.size();
        int rc = Dy.upload(dyn, dtOverride, recordPairs);
        if (rc != VX3B_OK)
            return rc;
        if (nStatic)
            return S.upload(staticDescs.data(), nStatic);
        return VX3B_OK;
    }
    int run(int64_t maxSteps, int useStop) override {
        push_options();
        if (!nDyn) {
            const int rc = S.run(maxSteps, useStop);
            lastMs = S.lastMs, lastLaunches = S.lastLaunches, lastVoxelSteps = S.lastVoxelSteps;
            return rc;
        }
        int rc = Dy.run_launch(maxSteps, useStop, recordPairs); // concurrently with the static worlds' kernels
        if (rc != VX3B_OK)
            return rc;
        int rcS = VX3B_OK;
        if (nStatic)
            rcS = S.run(maxSteps, useStop);
        rc = Dy.run_wait();
        if (rcS != VX3B_OK)
            return rcS;
        if (rc != VX3B_OK)
            return rc;
        lastMs = std::max(nStatic ? S.lastMs : 0.0f, Dy.lastMs);
        lastLaunches = (nStatic ? S.lastLaunches : 0) + 1;
        lastVoxelSteps = (nStatic ? S.lastVoxelSteps : 0) + Dy.lastVoxelSteps;
        return VX3B_OK;
    }
    int results(VX3B_WorldResult *out) override {
        if (where.empty())
            return S.results(out);
        std::vector<VX3B_WorldResult> rs((size_t)nStatic);
        if (nStatic) {
            const int rc = S.results(rs.data());
            if (rc != VX3B_OK)
                return rc;
        }
        for (size_t w = 0; w < where.size(); w++) {
            if (where[w].dyn) {
                const int rc = Dy.result(where[w].idx, out[w]);
                if (rc != VX3B_OK)
                    return rc;
            } else
                out[w] = rs[where[w].idx];
        }
        return VX3B_OK;
    }
    int read_state(int world, double *pos, double *orient, double *lin, double *ang) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.read_state(w.idx, pos, orient, lin, ang) : S.read_state(w.idx, pos, orient, lin, ang);
    }
    int read_initial_pos(int world, double *pos) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.read_initial_pos(w.idx, pos) : S.read_initial_pos(w.idx, pos);
    }
    int read_links(int world, double *pos2, double *a1, double *a2, int32_t *flags) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.read_links(w.idx, pos2, a1, a2, flags) : S.read_links(w.idx, pos2, a1, a2, flags);
    }
    int collision_pairs(int world, int32_t *pairs, int32_t *n) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.collision_pairs(w.idx, pairs, n) : S.collision_pairs(w.idx, pairs, n);
    }
    int read_aux(int world, float *voxelTemp, float *linkStrain) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.read_aux(w.idx, voxelTemp, linkStrain) : S.read_aux(w.idx, voxelTemp, linkStrain);
    }
    int set_force(int world, int voxel, const double *force, const double *moment) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_set_force") : S.set_force(w.idx, voxel, force, moment);
    }
    int set_global_damping(int world, int mat, double zeta) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_set_global_damping") : S.set_global_damping(w.idx, mat, zeta);
    }
    int num_links(int world, int32_t *n) override {
        Where w;
        if (!locate(world, w) || !n)
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.current_links(w.idx, n) : S.num_links(w.idx, n);
    }
    int read_signals(int world, double *sig6, int32_t *vflags2) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? Dy.read_signals(w.idx, sig6, vflags2) : S.read_signals(w.idx, sig6, vflags2);
    }
    int read_topology(int world, int32_t *links6, int32_t *n) override {
        Where w;
        if (!locate(world, w) || !n)
            return VX3B_ERR_ARG;
        if (w.dyn)
            return Dy.read_topology(w.idx, links6, n);
        return S.read_topology(w.idx, links6, n);
    }
    int run_async(int64_t steps) override {
        if (nDyn)
            return no_dynamic("vx3b_batch_run_async");
        push_options();
        return S.run_async(steps);
    }
    int halo_set(int world, int listId, const int32_t *voxels, int n) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_set") : S.halo_set(w.idx, listId, voxels, n);
    }
    int halo_record_bytes() override { return S.halo_record_bytes(); }
    int halo_export(int world, int listId, void *deviceBuffer) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_export") : S.halo_export(w.idx, listId, deviceBuffer);
    }
    int halo_import(int world, int listId, const void *deviceBuffer) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_import") : S.halo_import(w.idx, listId, deviceBuffer);
    }
    int halo_open(int world, int listId, VX3B_HaloEndpoint *out) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_open") : S.halo_open(w.idx, listId, out);
    }
    int halo_connect(int world, int listId, const VX3B_HaloEndpoint *peer, int sameProcess, const int32_t *peerSlots) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_connect") : S.halo_connect(w.idx, listId, peer, sameProcess, peerSlots);
    }
    int halo_slots(int world, int listId, int32_t *slots) override {
        Where w;
        if (!locate(world, w))
            return VX3B_ERR_ARG;
        return w.dyn ? no_dynamic("vx3b_batch_halo_slots") : S.halo_slots(w.idx, listId, slots);
    }
    int run_slab(int64_t steps) override {
        if (nDyn)
            return no_dynamic("vx3b_batch_run_slab");
        push_options();
        return S.run_slab(steps);
    }
    int halo_timeouts(int64_t *n, int64_t *waitNs) override { return S.halo_timeouts(n, waitNs); }
};

} // namespace vx3b
