// TEST INFRASTRUCTURE — not part of the product path.
//
// Harness around the UNMODIFIED reference CPU implementation (Voxelyze2 `CVoxelyze`
// + the Voxelyze3 host VXA reader). It is compiled by oracle/Makefile from the
// sources where they lie under /root/reference (nothing is copied into this repo)
// into oracle/_ref/vx2_ref.  It is used
//   * by tests/ to pin oracle/vx3_oracle.cpp and the CUDA engine against the real
//     reference on the same VXA text, and to regenerate tests/golden/*.vxbin;
//   * by bench.py as the `cpu_baseline` / `--impl reference` arm (kind "reference").
//
// What it drives (reference file:line):
//   CXML_Rip::LoadFile                         Voxelyze3/src/old/XML_Rip.h:49
//   CVX_Sim::ReadVXA / Import                  Voxelyze3/src/VXA/VX_Sim.cpp:155-183, 59-153
//   CVoxelyze::doTimeStep / recommendedTimeStep Voxelyze3/src/old/Voxelyze.cpp:251-311
// With --vx3 it additionally applies, from OUTSIDE the reference classes (all of
// their members are public), the stepping deltas that Voxelyze3's device code adds
// on top of Voxelyze2 (SURVEY.md §7 "two references that disagree"):
//   * dt computed in double then truncated to float   VX3_VoxelyzeKernel.cu:184-217,237-254
//   * per-voxel phase-offset temperature, clamped <=0  VX3_VoxelyzeKernel.cu:625-650
//   * cilia force on surface voxels                    VX3_VoxelyzeKernel.cu:846-859
//   * angMom zeroed under floor static friction        VX3_Voxel.cu:258-264
// VX3 voxel-voxel contact is NOT emulated here (Voxelyze2's own contact model
// differs); that part of the oracle is "parity unpinned" by the reference.

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>
#ifdef USE_OMP
#include <omp.h>
#endif

#include "VX_Environment.h"
#include "VX_MaterialLink.h"
#include "VX_MaterialVoxel.h"
#include "VX_Link.h"
#include "VX_Voxel.h"
#include "Voxelyze.h"
#include "VX_Object.h"
#include "VX_Sim.h"
#include "XML_Rip.h"

namespace {

// --- tiny tagged-array container ("vxbin"): "name dtype rows cols\n" + raw bytes ---
struct Dump {
    FILE *f = nullptr;
    bool open(const std::string &p) {
        f = fopen(p.c_str(), "wb");
        return f != nullptr;
    }
    void put(const char *name, const char *dtype, size_t rows, size_t cols, const void *data, size_t elt) {
        fprintf(f, "%s %s %zu %zu\n", name, dtype, rows, cols);
        if (rows * cols)
            fwrite(data, elt, rows * cols, f);
    }
    void f64(const char *name, const std::vector<double> &v, size_t cols) {
        put(name, "f64", cols ? v.size() / cols : 0, cols, v.data(), 8);
    }
    void i32(const char *name, const std::vector<int32_t> &v, size_t cols) {
        put(name, "i32", cols ? v.size() / cols : 0, cols, v.data(), 4);
    }
    void close() {
        if (f)
            fclose(f);
        f = nullptr;
    }
};

struct Args {
    std::string vxa, dump, dump_model;
    long steps = 0;
    bool vx3 = false, enable_expansion = false, enable_cilia = false, no_temp = false, timing = false, vx2_collisions = false;
    int threads = 1;
    double dt_override = 0;
    int repeat = 1;
};

template <class T, class... A> void add(std::vector<T> &v, A... a) { (v.push_back((T)a), ...); }

int index_of(const std::vector<CVX_Voxel *> &list, const CVX_Voxel *p) {
    for (size_t i = 0; i < list.size(); i++)
        if (list[i] == p)
            return (int)i;
    return -1;
}

double vx3_recommended_dt(CVoxelyze &Vx) {
    // VX3_VoxelyzeKernel.cu:184-217 — same scan as Voxelyze2 but accumulated in double.
    double MaxFreq2 = 0.0f;
    for (CVX_Link *pL : Vx.linksList) {
        double m1 = pL->pVNeg->mat->mass(), m2 = pL->pVPos->mat->mass();
        double f2 = pL->axialStiffness() / (m1 < m2 ? m1 : m2);
        if (f2 > MaxFreq2)
            MaxFreq2 = f2;
    }
    if (MaxFreq2 <= 0.0f)
        for (CVX_Voxel *pV : Vx.voxelsList) {
            double f2 = pV->mat->youngsModulus() * pV->mat->nomSize / pV->mat->mass();
            if (f2 > MaxFreq2)
                MaxFreq2 = f2;
        }
    if (MaxFreq2 <= 0.0f)
        return 0.0f;
    return 1.0f / (6.283185f * sqrt(MaxFreq2));
}

void dump_state(Dump &d, CVoxelyze &Vx, double t, double dt, double optimalDt, long steps) {
    const auto &V = Vx.voxelsList;
    const auto &L = Vx.linksList;
    std::vector<double> pos, orient, linMom, angMom, temp;
    std::vector<int32_t> flags;
    for (CVX_Voxel *v : V) {
        add(pos, v->pos.x, v->pos.y, v->pos.z);
        add(orient, v->orient.w, v->orient.x, v->orient.y, v->orient.z);
        add(linMom, v->linMom.x, v->linMom.y, v->linMom.z);
        add(angMom, v->angMom.x, v->angMom.y, v->angMom.z);
        temp.push_back(v->temp);
        flags.push_back((int32_t)v->boolStates);
    }
    std::vector<double> pos2, a1v, a2v, fN, fP, mN, mP, lscal;
    std::vector<int32_t> lflags;
    for (CVX_Link *l : L) {
        add(pos2, l->pos2.x, l->pos2.y, l->pos2.z);
        add(a1v, l->angle1v.x, l->angle1v.y, l->angle1v.z);
        add(a2v, l->angle2v.x, l->angle2v.y, l->angle2v.z);
        add(fN, l->forceNeg.x, l->forceNeg.y, l->forceNeg.z);
        add(fP, l->forcePos.x, l->forcePos.y, l->forcePos.z);
        add(mN, l->momentNeg.x, l->momentNeg.y, l->momentNeg.z);
        add(mP, l->momentPos.x, l->momentPos.y, l->momentPos.z);
        add(lscal, (double)l->strain, (double)l->maxStrain, l->currentRestLength);
        lflags.push_back((l->smallAngle ? 1 : 0) | ((l->boolStates & 1) ? 2 : 0));
    }
    std::vector<double> sc = {t, dt, optimalDt, (double)steps};
    d.f64("scalars", sc, 4);
    d.f64("pos", pos, 3);
    d.f64("orient", orient, 4);
    d.f64("linMom", linMom, 3);
    d.f64("angMom", angMom, 3);
    d.f64("temp", temp, 1);
    d.i32("vflags", flags, 1);
    d.f64("link_pos2", pos2, 3);
    d.f64("link_angle1v", a1v, 3);
    d.f64("link_angle2v", a2v, 3);
    d.f64("link_forceNeg", fN, 3);
    d.f64("link_forcePos", fP, 3);
    d.f64("link_momentNeg", mN, 3);
    d.f64("link_momentPos", mP, 3);
    d.f64("link_scalars", lscal, 3);
    d.i32("link_flags", lflags, 1);
}

void dump_model(Dump &d, CVX_Sim &Sim, CVX_Environment &Env) {
    CVoxelyze &Vx = Sim.Vx;
    std::vector<int32_t> vox; // ix iy iz matIndex matid
    std::vector<double> vextra; // phaseOffset baseCilia(3) shiftCilia(3)
    std::vector<int32_t> ext; // has_ext dofFixed
    std::vector<double> extv; // force(3) moment(3) translation(3) rotation(3)
    for (CVX_Voxel *v : Vx.voxelsList) {
        int mi = -1;
        for (size_t i = 0; i < Vx.voxelMats.size(); i++)
            if (Vx.voxelMats[i] == v->mat)
                mi = (int)i;
        add(vox, v->ix, v->iy, v->iz, mi, v->matid);
        add(vextra, v->phaseOffset, v->baseCiliaForce.x, v->baseCiliaForce.y, v->baseCiliaForce.z, v->shiftCiliaForce.x,
                                     v->shiftCiliaForce.y, v->shiftCiliaForce.z);
        if (v->ext) {
            add(ext, 1, (int32_t)v->ext->dofFixed);
            add(extv, v->ext->extForce.x, v->ext->extForce.y, v->ext->extForce.z, v->ext->extMoment.x, v->ext->extMoment.y,
                                     v->ext->extMoment.z, v->ext->extTranslation.x, v->ext->extTranslation.y, v->ext->extTranslation.z,
                                     v->ext->extRotation.x, v->ext->extRotation.y, v->ext->extRotation.z);
        } else {
            add(ext, 0, 0);
            extv.insert(extv.end(), 12, 0.0);
        }
    }
    std::vector<int32_t> links; // neg pos axis
    std::vector<double> lmat;   // E nu a1 a2 b1 b2 b3 sqA1 sqA2xIp sqB1 sqB2xFMp sqB3xIp epsilonFail restLength
    for (CVX_Link *l : Vx.linksList) {
        add(links, index_of(Vx.voxelsList, l->pVNeg), index_of(Vx.voxelsList, l->pVPos), (int32_t)l->axis);
        CVX_MaterialLink *m = l->mat;
        add(lmat, m->E, m->nu, m->_a1, m->_a2, m->_b1, m->_b2, m->_b3, m->_sqA1, m->_sqA2xIp, m->_sqB1, m->_sqB2xFMp,
                                 m->_sqB3xIp, m->epsilonFail, l->currentRestLength);
    }
    std::vector<double> mats; // per voxel material
    for (CVX_MaterialVoxel *m : Vx.voxelMats) {
        add(mats, m->E, m->nu, m->rho, m->alphaCTE, m->muStatic, m->muKinetic, m->zetaInternal, m->zetaGlobal, m->zetaCollision,
                                 m->nomSize, m->gravMult, m->_mass, m->_massInverse, m->_sqrtMass, m->_momentInertia, m->_momentInertiaInverse,
                                 m->_2xSqMxExS, m->_2xSqIxExSxSxS, m->Cilia, (double)m->fixed, (double)m->isTarget, (double)m->isMeasured,
                                 (double)m->matid, m->epsilonFail);
    }
    std::vector<double> env = {Vx.voxSize,
                               Sim.DtFrac,
                               (double)Vx.recommendedTimeStep(),
                               vx3_recommended_dt(Vx),
                               (double)Vx.gravity(),
                               (double)Vx.isFloorEnabled(),
                               (double)Vx.isCollisionsEnabled(),
                               (double)Env.IsTempEnabled(),
                               (double)Env.IsTempVaryEnabled(),
                               Env.GetTempBase(),
                               Env.GetTempAmplitude(),
                               Env.GetTempPeriod()};
    d.f64("env", env, env.size());
    d.i32("voxels", vox, 5);
    d.f64("voxel_extra", vextra, 7);
    d.i32("ext_flags", ext, 2);
    d.f64("ext_values", extv, 12);
    d.i32("links", links, 3);
    d.f64("link_mats", lmat, 14);
    d.f64("voxel_mats", mats, 24);
}

} // namespace

int main(int argc, char **argv) {
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (k == "--vxa")
            a.vxa = next();
        else if (k == "--steps")
            a.steps = atol(next());
        else if (k == "--dump")
            a.dump = next();
        else if (k == "--dump-model")
            a.dump_model = next();
        else if (k == "--vx3")
            a.vx3 = true;
        else if (k == "--enable-expansion")
            a.enable_expansion = atoi(next()) != 0;
        else if (k == "--enable-cilia")
            a.enable_cilia = atoi(next()) != 0;
        else if (k == "--vx2-collisions")
            a.vx2_collisions = atoi(next()) != 0;
        else if (k == "--no-temp")
            a.no_temp = true;
        else if (k == "--time")
            a.timing = true;
        else if (k == "--threads")
            a.threads = atoi(next());
        else if (k == "--dt")
            a.dt_override = atof(next());
        else if (k == "--repeat")
            a.repeat = atoi(next());
        else {
            fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 2;
        }
    }
    if (a.vxa.empty()) {
        fprintf(stderr, "usage: vx2_ref --vxa FILE --steps N [--vx3] [--enable-cilia 1] [--enable-expansion 1] [--no-temp] [--dt X]\n"
                        "               [--threads T] [--dump OUT] [--dump-model OUT] [--time]\n");
        return 2;
    }
#ifdef USE_OMP
    omp_set_num_threads(a.threads > 0 ? a.threads : 1);
#endif

    double total_loop_s = 0;
    long long total_voxel_steps = 0;
    int nv = 0, nl = 0;
    for (int rep = 0; rep < a.repeat; rep++) {
        // heap-allocated and intentionally leaked: the reference's destructors crash on a world that never stepped
        CVX_Environment &Env = *new CVX_Environment;
        CVX_Sim &Sim = *new CVX_Sim;
        CVX_Object &Obj = *new CVX_Object;
        Env.pObj = &Obj;
        Sim.pEnv = &Env;
        CXML_Rip XML;
        std::string msg;
        if (!XML.LoadFile(a.vxa, &msg)) {
            fprintf(stderr, "cannot load %s: %s\n", a.vxa.c_str(), msg.c_str());
            return 1;
        }
        Sim.ReadVXA(&XML, &msg);
        Sim.Import(NULL, NULL, &msg);
        CVoxelyze &Vx = Sim.Vx;
        if (a.vx2_collisions) { // Voxelyze2's own watch-list contact (Voxelyze.cpp:670-750).  Off and on again: enableCollisions
            Vx.enableCollisions(false); // returns early when the world-level flag already has the requested value
            Vx.enableCollisions(true);  // (Voxelyze.cpp:612-622), which would leave the per-voxel flags of Import()'s voxels unset
        }
        nv = (int)Vx.voxelsList.size();
        nl = (int)Vx.linksList.size();

        if (!a.dump_model.empty() && rep == 0) {
            Dump d;
            if (!d.open(a.dump_model))
                return 1;
            dump_model(d, Sim, Env);
            d.close();
        }

        double optimalDt = a.vx3 ? vx3_recommended_dt(Vx) : (double)Vx.recommendedTimeStep();
        if (a.vx3 && optimalDt < 1e-10)
            optimalDt = 1e-10;
        float dt = (float)(Sim.DtFrac * optimalDt);
        if (a.dt_override > 0)
            dt = (float)a.dt_override;

        const double A = Env.GetTempAmplitude(), P = Env.GetTempPeriod();
        const bool vary = Env.IsTempVaryEnabled() && P > 0 && !a.no_temp;
        double t = 0;

        auto t0 = std::chrono::steady_clock::now();
        for (long s = 0; s < a.steps; s++) {
            if (a.vx3) {
                if (vary) {
                    for (CVX_Voxel *v : Vx.voxelsList) {
                        if (v->mat->fixed)
                            continue;
                        double T = A * sin(2 * 3.1415926f * (t / P + v->phaseOffset));
                        if (!a.enable_expansion && T > 0)
                            T = 0;
                        v->setTemperature(T);
                    }
                }
                if (a.enable_cilia) {
                    for (CVX_Voxel *v : Vx.voxelsList) {
                        if (!v->isSurface() || v->mat->Cilia == 0)
                            continue;
                        Vec3D<double> c = v->orient.RotateVec3D(v->baseCiliaForce) * v->mat->Cilia;
                        v->external()->setForce((float)c.x, (float)c.y, (float)c.z);
                    }
                }
            }
            if (!Vx.doTimeStep(dt)) {
                fprintf(stderr, "diverged at step %ld\n", s);
                break;
            }
            if (a.vx3) {
                for (CVX_Voxel *v : Vx.voxelsList)
                    if (v->isFloorEnabled() && v->floorPenetration() >= 0 && v->isFloorStaticFriction())
                        v->angMom = Vec3D<double>(0, 0, 0);
            }
            t += dt;
        }
        if (getenv("VX2_REF_DEBUG"))
            fprintf(stderr, "collisions world=%d list=%zu voxel0=%d watchlist0=%zu\n", (int)Vx.isCollisionsEnabled(), Vx.collisionList()->size(),
                    (int)Vx.voxelsList[0]->isCollisionsEnabled(), Vx.voxelsList[0]->colWatch ? Vx.voxelsList[0]->colWatch->size() : (size_t)999);
        auto t1 = std::chrono::steady_clock::now();
        total_loop_s += std::chrono::duration<double>(t1 - t0).count();
        total_voxel_steps += (long long)nv * a.steps;

        if (!a.dump.empty() && rep == 0) {
            Dump d;
            if (!d.open(a.dump))
                return 1;
            dump_state(d, Vx, t, dt, optimalDt, a.steps);
            d.close();
        }
    }
    fflush(nullptr);
    if (a.timing)
        printf("{\"voxels\": %d, \"links\": %d, \"steps\": %ld, \"repeat\": %d, \"threads\": %d, \"loop_s\": %.6f, \"voxel_steps_per_s\": %.6e}\n", nv, nl,
               a.steps, a.repeat, a.threads, total_loop_s, total_loop_s > 0 ? total_voxel_steps / total_loop_s : 0.0);
    fflush(nullptr);
    _exit(0); // skip static/global teardown of the reference classes
}
