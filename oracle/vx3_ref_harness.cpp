// TEST INFRASTRUCTURE — not part of the product path.
//
// Harness around the reference's own VOXELYZE3 step code, run on the CPU.  oracle/Makefile compiles the reference's VX3
// "device" sources (Voxelyze3/src/VX3/VX3_{Voxel,Link,Collision,Material,MaterialVoxel,MaterialLink,External,VoxelyzeKernel}.cu,
// plain scalar C++ behind __device__) where they lie under /root/reference with g++ against oracle/vx3_shim/ (host stand-ins
// for cuda_runtime.h / curand / the three boost helpers) into oracle/_ref/vx3_ref.  Nothing is copied into this repo.  Two
// textual rewrites happen in the compiler's input pipe (oracle/Makefile, `sed`), neither touches arithmetic:
//   * `kernel<<<grid, block>>>(args)`  ->  `vx3shim::launcher(grid, block, kernel)(args)`: every thread of the grid runs to
//     completion one after the other (x index outermost, ascending; y inside it).  Where the reference's kernels race
//     (contact force sums, signals, link creation: DESIGN.md §8) a serial schedule is ONE valid execution of them, and it is the
//     one oracle/vx3_oracle.cpp restates;
//   * one recording call after `VX3_Collision collision(voxel1, voxel2);` (VX3_VoxelyzeKernel.cu:683) so the set of watched
//     pairs of a step can be read back (the reference keeps no such list).
//
// What this file does itself (the reference does it in VX3_SimulationManager.cu, which needs boost::property_tree):
//   * VXA -> CVX_Sim through the reference's own reader (CXML_Rip, CVX_Sim::ReadVXA / Import: VXA/VX_Sim.cpp:59-183), then
//     the reference's constructor VX3_VoxelyzeKernel(CVX_Sim*) (VX3_VoxelyzeKernel.cu:25-114);
//   * the VX3-only settings readVXD() takes from the XML (VX3_SimulationManager.cu:277-352), read here with the reference's
//     tinyxml; formulas through the same breadth-first token order as ParseMathTree (:159-275);
//   * the driver loop of CUDA_Simulation (VX3_SimulationManager.cu:11-121): syncVectors, saveInitialPosition, registerTargets,
//     initial centre of mass, then StopConditionMet / doTimeStep up to --steps times, final centre of mass and fitness.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <set>
#include <stack>
#include <string>
#include <vector>
#include <unistd.h>

#include "VX3_VoxelyzeKernel.cuh"
#include "VX_Object.h"
#include "VX_Sim.h"
#include "XML_Rip.h"
#include "tinyxml.h"

// ---- the recording hook the Makefile's sed inserts into handle_collision_attachment -----------------------------------------
static std::vector<std::pair<const void *, const void *>> g_pairs;
static bool g_record = false;
void vx3shim_pair_hook(const void *v1, const void *v2) {
    if (g_record)
        g_pairs.emplace_back(v1, v2);
}

namespace {

struct Dump {
    FILE *f = nullptr;
    void put(const std::string &name, const char *dtype, size_t rows, size_t cols, const void *data, size_t elt) {
        fprintf(f, "%s %s %zu %zu\n", name.c_str(), dtype, rows, cols);
        if (rows * cols)
            fwrite(data, elt, rows * cols, f);
    }
    void f64(const std::string &name, const std::vector<double> &v, size_t cols) { put(name, "f64", cols ? v.size() / cols : 0, cols, v.data(), 8); }
    void i32(const std::string &name, const std::vector<int32_t> &v, size_t cols) { put(name, "i32", cols ? v.size() / cols : 0, cols, v.data(), 4); }
};
template <class T, class... A> void add(std::vector<T> &v, A... a) { (v.push_back((T)a), ...); }

TiXmlElement *find(TiXmlElement *root, const char *path) { // "A.B.C" below root
    TiXmlElement *e = root;
    std::string p = path;
    size_t i = 0;
    while (e && i <= p.size()) {
        size_t j = p.find('.', i);
        if (j == std::string::npos)
            j = p.size();
        e = e->FirstChildElement(p.substr(i, j - i).c_str());
        i = j + 1;
    }
    return e;
}
std::string text_of(TiXmlElement *e) {
    const char *t = e ? e->GetText() : nullptr;
    std::string s = t ? t : "";
    while (!s.empty() && isspace((unsigned char)s.back()))
        s.pop_back();
    return s;
}
double get_d(TiXmlElement *root, const char *path, double def) {
    TiXmlElement *e = find(root, path);
    return e && e->GetText() ? atof(e->GetText()) : def;
}

// VX3_SimulationManager::ParseMathTree (VX3_SimulationManager.cu:159-275): breadth-first over the formula's elements, every
// element pushed on a stack behind an mtEND; the program is the stack popped, i.e. the reversed level order
void parse_math_tree(VX3_MathTreeToken *out, TiXmlElement *root, const char *path) {
    TiXmlElement *node = find(root, path);
    if (!node)
        return;
    std::queue<TiXmlElement *> frontier;
    std::stack<std::pair<std::string, std::string>> tokens;
    tokens.push({"mtEND", ""});
    frontier.push(node);
    while (!frontier.empty()) {
        TiXmlElement *t = frontier.front();
        frontier.pop();
        for (TiXmlElement *c = t->FirstChildElement(); c; c = c->NextSiblingElement()) {
            tokens.push({c->Value(), text_of(c)});
            frontier.push(c);
        }
    }
    static const char *names[] = {"mtEND", "mtCONST", "mtE", "mtPI", "mtVAR", "mtADD", "mtSUB", "mtMUL", "mtDIV", "mtPOW", "mtSQRT", "mtSIN", "mtCOS",
                                  "mtTAN", "mtATAN", "mtLOG", "mtINT", "mtABS", "mtNOT", "mtGREATERTHAN", "mtLESSTHAN", "mtAND", "mtOR", "mtNORMALCDF"};
    static const char *vars[] = {"x", "y", "z", "hit", "t", "angle", "targetCloseness", "numClosePairs", "num_voxel"};
    int i = 0;
    while (!tokens.empty()) {
        auto tok = tokens.top();
        tokens.pop();
        int op = -1;
        for (int k = 0; k < 24; k++)
            if (tok.first == names[k])
                op = k;
        if (op < 0) {
            fprintf(stderr, "vx3_ref: unknown formula token %s\n", tok.first.c_str());
            exit(3);
        }
        out[i].op = (VX3_MathTreeOperator)op;
        if (op == mtVAR) {
            out[i].value = -1;
            for (int k = 0; k < 9; k++)
                if (tok.second == vars[k])
                    out[i].value = k;
        } else if (op == mtCONST)
            out[i].value = std::stod(tok.second);
        i++;
    }
}

struct Args {
    std::string vxa, dump;
    long steps = 0;
    bool use_stop = false, pairs = false, timing = false;
    std::vector<long> dump_at;
    int repeat = 1;
};

int vindex(VX3_VoxelyzeKernel *k, const void *p) { return (int)((const VX3_Voxel *)p - k->d_voxels); }

void dump_state(Dump &d, VX3_VoxelyzeKernel *k, const std::string &sfx, float dt, long steps, bool diverged) {
    std::vector<double> pos, orient, linMom, angMom, temp, sig;
    std::vector<int32_t> flags;
    for (int i = 0; i < k->num_d_voxels; i++) {
        VX3_Voxel *v = &k->d_voxels[i];
        add(pos, v->pos.x, v->pos.y, v->pos.z);
        add(orient, v->orient.w, v->orient.x, v->orient.y, v->orient.z);
        add(linMom, v->linMom.x, v->linMom.y, v->linMom.z);
        add(angMom, v->angMom.x, v->angMom.y, v->angMom.z);
        temp.push_back(v->tempe);
        add(flags, (int32_t)v->boolStates, (int32_t)v->removed, (int32_t)v->enableAttach);
        add(sig, v->localSignal, v->localSignaldt, v->inactiveUntil, v->d_signal.value, v->d_signal.activeTime, v->packmakerNextPulse);
    }
    std::vector<int32_t> links; // neg, pos, axis, removed, isDetached, isNewLink
    std::vector<double> lstate; // pos2(3) angle1v(3) angle2v(3) strain maxStrain restLength
    for (unsigned i = 0; i < k->d_v_links.size(); i++) {
        VX3_Link *l = k->d_v_links[i];
        add(links, vindex(k, l->pVNeg), vindex(k, l->pVPos), (int32_t)l->axis, (int32_t)l->removed, (int32_t)l->isDetached, (int32_t)l->isNewLink);
        add(lstate, l->pos2.x, l->pos2.y, l->pos2.z, l->angle1v.x, l->angle1v.y, l->angle1v.z, l->angle2v.x, l->angle2v.y, l->angle2v.z, (double)l->strain,
            (double)l->maxStrain, l->currentRestLength);
    }
    std::vector<int32_t> surf;
    for (int i = 0; i < k->num_d_surface_voxels; i++)
        surf.push_back(vindex(k, k->d_surface_voxels[i]));
    std::vector<double> sc = {k->currentTime, (double)dt, k->OptimalDt, (double)steps, (double)k->collisionCount, k->fitness_score, k->currentCenterOfMass.x,
                              k->currentCenterOfMass.y, k->currentCenterOfMass.z, k->initialCenterOfMass.x, k->initialCenterOfMass.y, k->initialCenterOfMass.z,
                              k->recentAngle, k->targetCloseness, (double)k->numClosePairs, (double)k->angleSampleTimes, (double)diverged, (double)k->CurStepCount};
    d.f64("scalars" + sfx, sc, sc.size());
    d.f64("pos" + sfx, pos, 3);
    d.f64("orient" + sfx, orient, 4);
    d.f64("linMom" + sfx, linMom, 3);
    d.f64("angMom" + sfx, angMom, 3);
    d.f64("temp" + sfx, temp, 1);
    d.i32("vflags" + sfx, flags, 3);
    d.f64("signals" + sfx, sig, 6);
    d.i32("links" + sfx, links, 6);
    d.f64("link_state" + sfx, lstate, 12);
    d.i32("surface" + sfx, surf, 1);
    std::vector<int32_t> pairs;
    std::set<std::pair<int, int>> seen;
    for (auto &p : g_pairs)
        seen.insert({vindex(k, p.first), vindex(k, p.second)});
    for (auto &p : seen)
        add(pairs, p.first, p.second);
    d.i32("pairs" + sfx, pairs, 2);
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
        else if (k == "--use-stop")
            a.use_stop = true;
        else if (k == "--time")
            a.timing = true;
        else if (k == "--repeat")
            a.repeat = atoi(next());
        else if (k == "--dump-at") { // comma separated step counts: state arrays get the suffix _N
            std::string s = next();
            size_t p = 0;
            while (p < s.size()) {
                a.dump_at.push_back(atol(s.c_str() + p));
                p = s.find(',', p);
                if (p == std::string::npos)
                    break;
                p++;
            }
        } else {
            fprintf(stderr, "unknown argument %s\n", k.c_str());
            return 2;
        }
    }
    if (a.vxa.empty()) {
        fprintf(stderr, "usage: vx3_ref --vxa MERGED.vxa --steps N [--use-stop] [--dump OUT] [--dump-at N1,N2,...] [--time] [--repeat R]\n");
        return 2;
    }
    Dump d;
    if (!a.dump.empty() && !(d.f = fopen(a.dump.c_str(), "wb")))
        return 1;
    double loop_s = 0;
    long long voxel_steps = 0;
    int nv = 0, nl = 0;
    for (int rep = 0; rep < a.repeat; rep++) {
        // heap-allocated and leaked on purpose (the reference's destructors free through a background "memory cleaner")
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
        VX3_VoxelyzeKernel &h = *new VX3_VoxelyzeKernel(&Sim);

        TiXmlDocument doc;
        if (!doc.LoadFile(a.vxa.c_str())) {
            fprintf(stderr, "tinyxml cannot load %s\n", a.vxa.c_str());
            return 1;
        }
        TiXmlElement *root = doc.RootElement(); // <VXA>
        parse_math_tree(h.StopConditionFormula, root, "Simulator.StopCondition.StopConditionFormula");
        h.EnableCollision = get_d(root, "Simulator.AttachDetach.EnableCollision", 1) != 0;
        h.enableAttach = get_d(root, "Simulator.AttachDetach.EnableAttach", 0) != 0;
        h.enableDetach = get_d(root, "Simulator.AttachDetach.EnableDetach", 0) != 0;
        h.watchDistance = get_d(root, "Simulator.AttachDetach.watchDistance", 1.0);
        h.boundingRadius = get_d(root, "Simulator.AttachDetach.boundingRadius", 0.75);
        h.SafetyGuard = (int)get_d(root, "Simulator.AttachDetach.SafetyGuard", 500);
        for (int c = 0; c < 5; c++)
            parse_math_tree(h.AttachCondition[c], root, ("Simulator.AttachDetach.AttachCondition.Condition_" + std::to_string(c)).c_str());
        h.RecordStepSize = 0; // no history stream from this harness
        parse_math_tree(h.fitness_function, root, "Simulator.FitnessFunction");
        parse_math_tree(h.force_field.token_x_forcefield, root, "Simulator.ForceField.x_forcefield");
        parse_math_tree(h.force_field.token_y_forcefield, root, "Simulator.ForceField.y_forcefield");
        parse_math_tree(h.force_field.token_z_forcefield, root, "Simulator.ForceField.z_forcefield");
        h.SavePositionOfAllVoxels = (int)get_d(root, "Simulator.SavePositionOfAllVoxels", 0);
        h.MaxDistInVoxelLengthsToCountAsPair = get_d(root, "Simulator.MaxDistInVoxelLengthsToCountAsPair", 0);
        h.EnableCilia = (int)get_d(root, "Simulator.EnableCilia", 0);
        h.EnableSignals = (int)get_d(root, "Simulator.EnableSignals", 0);
        h.SecondaryExperiment = (int)get_d(root, "Simulator.SecondaryExperiment", 0);
        h.ReinitializeInitialPositionAfterThisManySeconds = get_d(root, "Simulator.ReinitializeInitialPositionAfterThisManySeconds", 0.0);
        h.EnableExpansion = (int)get_d(root, "Simulator.EnableExpansion", 0);

        // the reference copies the object to the device bit by bit and runs on the copy (VX3_SimulationManager.cu:348)
        VX3_VoxelyzeKernel *k = (VX3_VoxelyzeKernel *)malloc(sizeof(VX3_VoxelyzeKernel));
        memcpy((void *)k, (void *)&h, sizeof(VX3_VoxelyzeKernel));
        nv = k->num_d_voxels;
        nl = k->num_d_links;

        // CUDA_Simulation (VX3_SimulationManager.cu:11-121)
        k->syncVectors();
        k->saveInitialPosition();
        k->isSurfaceChanged = true;
        k->registerTargets();
        k->updateCurrentCenterOfMass();
        k->InitializeCenterOfMass();
        bool diverged = false;
        long done = 0;
        size_t nextDump = 0;
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (long j = 0; j < a.steps; j++) {
            if (a.use_stop && k->StopConditionMet())
                break;
            g_record = d.f && rep == 0; // the pair set of the most recent step (cleared below), whichever step turns out to be the last
            g_pairs.clear();
            if (!k->doTimeStep()) {
                diverged = true;
                break;
            }
            done = j + 1;
            if (rep == 0 && d.f && nextDump < a.dump_at.size() && done == a.dump_at[nextDump]) {
                dump_state(d, k, "_" + std::to_string(done), (float)(k->DtFrac * k->OptimalDt), done, false);
                nextDump++;
            }
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        loop_s += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        voxel_steps += (long long)nv * done;
        k->updateCurrentCenterOfMass();
        k->computeFitness();
        if (rep == 0 && d.f)
            dump_state(d, k, "", (float)(k->DtFrac * k->OptimalDt), done, diverged);
    }
    if (d.f)
        fclose(d.f);
    if (a.timing)
        printf("{\"voxels\": %d, \"links\": %d, \"steps\": %ld, \"repeat\": %d, \"threads\": 1, \"loop_s\": %.6f, \"voxel_steps_per_s\": %.6e}\n", nv, nl, a.steps,
               a.repeat, loop_s, loop_s > 0 ? voxel_steps / loop_s : 0.0);
    fflush(nullptr);
    _exit(0);
}
