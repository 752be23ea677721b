// Job level of the C ABI: a directory of base.vxa + *.vxd -> report XML.
// Replaces vx3_node_worker's main (Voxelyze3/src/Executables/vx3_node_worker.cu:32-142) and
// VX3_SimulationManager::start / readVXD / collectResults / sortResults
// (Voxelyze3/src/VX3/VX3_SimulationManager.cu:138-155, 277-380, 416-460), with one host thread per device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <vector>

#include "host_model.hpp"
#include "vx3b200.h"

namespace vx3b {
void set_error(const std::string &e);

struct JobWorld {
    std::string file;
    vx3b_model *model = nullptr;
    VX3B_WorldDesc desc;
    VX3B_WorldResult res;
    std::vector<double> pos, initPos;
};

static bool read_file(const std::string &p, std::string *out) {
    std::ifstream f(p, std::ios::binary);
    if (!f)
        return false;
    std::stringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

static std::string fmt_double(double v) { // boost ptree's stream translator: max_digits10 significant digits
    char b[64];
    snprintf(b, sizeof(b), "%.17g", v);
    return b;
}

static bool ends_with_ci(const std::string &s, const std::string &ext) { // ctool::u_with_ext: case-insensitive extension match
    if (s.size() < ext.size())
        return false;
    for (size_t i = 0; i < ext.size(); i++)
        if (tolower((unsigned char)s[s.size() - ext.size() + i]) != tolower((unsigned char)ext[i]))
            return false;
    return true;
}

// fitness descending, NaN last (VX3_SimulationResult.h:26-33)
static bool better(const JobWorld *a, const JobWorld *b) {
    if (std::isnan(b->res.fitness_score))
        return !std::isnan(a->res.fitness_score);
    if (std::isnan(a->res.fitness_score))
        return false;
    return a->res.fitness_score > b->res.fitness_score;
}

} // namespace vx3b

// Longest-processing-time-first: worlds in descending cost, each to the currently lightest shard (ties: lowest shard index,
// then lowest world index, so every rank of a multi-process job computes the same assignment).
extern "C" int vx3b_shard_worlds(const int64_t *cost, int n, int num_shards, int32_t *shard_of) {
    using namespace vx3b;
    if (n < 0 || num_shards < 1 || (n > 0 && (!cost || !shard_of))) {
        set_error("vx3b_shard_worlds: bad argument");
        return VX3B_ERR_ARG;
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++)
        order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<int64_t> load(num_shards, 0);
    for (int i : order) {
        int best = 0;
        for (int s = 1; s < num_shards; s++)
            if (load[s] < load[best])
                best = s;
        shard_of[i] = best;
        load[best] += cost[i] > 0 ? cost[i] : 1;
    }
    return VX3B_OK;
}

namespace vx3b {
// every *.vxd of a directory, sorted by name (what Voxelyze3.cpp:101-105 lists into the task file)
static std::vector<std::string> list_vxd(const std::string &dir) {
    std::vector<std::string> files;
    if (DIR *d = opendir(dir.c_str())) {
        while (dirent *e = readdir(d))
            if (ends_with_ci(e->d_name, ".vxd"))
                files.push_back(e->d_name);
        closedir(d);
    }
    std::sort(files.begin(), files.end());
    return files;
}
static std::vector<std::string> to_list(const char *const *files, int n) {
    std::vector<std::string> v;
    for (int i = 0; i < n; i++)
        v.push_back(files[i] ? files[i] : "");
    return v;
}

static int job_run(const std::string &basePath, const std::string &dir, const std::vector<std::string> &files, const char *report_path, int num_devices,
                   int precision) {
    std::string base;
    if (!read_file(basePath, &base)) {
        set_error("vx3b_job_run: cannot read " + basePath);
        return VX3B_ERR_IO;
    }
    if (files.empty()) {
        set_error("vx3b_job_run: no .vxd files for " + dir);
        return VX3B_ERR_IO;
    }
    // A file that cannot be read or parsed is reported and left out, like the reference, which prints the reader's message
    // and carries on with the next file (VX3_SimulationManager.cu:305-309); the job only fails when nothing is left.
    std::vector<JobWorld> worlds;
    std::string firstError;
    for (size_t i = 0; i < files.size(); i++) {
        std::string vxd;
        JobWorld w;
        w.file = files[i];
        int rc = read_file(dir + "/" + files[i], &vxd) ? vx3b_model_from_vxa(base.c_str(), vxd.c_str(), &w.model) : VX3B_ERR_IO;
        if (rc != VX3B_OK) {
            const std::string msg = files[i] + ": " + (rc == VX3B_ERR_IO ? std::string("cannot read file") : std::string(vx3b_last_error()));
            fprintf(stderr, "ERROR: %s (skipped)\n", msg.c_str());
            if (firstError.empty())
                firstError = msg;
            continue;
        }
        vx3b_model_desc(w.model, &w.desc);
        worlds.push_back(std::move(w));
    }
    if (worlds.empty()) {
        set_error("vx3b_job_run: no usable .vxd (" + firstError + ")");
        return VX3B_ERR_PARSE;
    }
    // split over devices, one host thread and one context per device.  The reference deals files round-robin
    // (vx3_node_worker.cu:88-93); worlds differ 10x in cost, so we deal them longest-processing-time-first instead.
    std::vector<std::vector<int>> sub(num_devices);
    {
        std::vector<int64_t> cost(worlds.size());
        std::vector<int32_t> shard(worlds.size());
        for (size_t i = 0; i < worlds.size(); i++)
            cost[i] = (int64_t)worlds[i].desc.num_voxels + 2 * (int64_t)worlds[i].desc.num_links;
        vx3b_shard_worlds(cost.data(), (int)worlds.size(), num_devices, shard.data());
        for (size_t i = 0; i < worlds.size(); i++)
            sub[shard[i]].push_back((int)i);
    }
    std::vector<int> status(num_devices, VX3B_OK);
    std::vector<std::string> errors(num_devices);
    std::vector<std::thread> threads;
    for (int dev = 0; dev < num_devices; dev++) {
        if (sub[dev].empty())
            continue;
        threads.emplace_back([&, dev]() {
            vx3b_ctx *ctx = nullptr;
            auto fail = [&](int rc) {
                status[dev] = rc;
                errors[dev] = vx3b_last_error();
                if (ctx)
                    vx3b_ctx_destroy(ctx);
            };
            // a generation of many worlds, none of which wants its voxel positions back: the pipelined runner (generation.cpp) steps the
            // first sub-batches while the later ones are still being staged and copied
            bool wantsPositions = false;
            for (int i : sub[dev])
                wantsPositions = wantsPositions || worlds[i].desc.save_position_of_all_voxels;
            if (sub[dev].size() >= 64 && !wantsPositions) {
                vx3b_gen *gen = nullptr;
                int rc = vx3b_gen_create(dev, precision, 6, &gen);
                if (rc != VX3B_OK)
                    return fail(rc);
                std::vector<const vx3b_model *> ms;
                for (int i : sub[dev])
                    ms.push_back(worlds[i].model);
                std::vector<VX3B_WorldResult> res(ms.size());
                rc = vx3b_gen_run(gen, ms.data(), (int)ms.size(), 1000000, 1, res.data());
                if (rc != VX3B_OK) {
                    status[dev] = rc;
                    errors[dev] = vx3b_last_error();
                } else
                    for (size_t k = 0; k < sub[dev].size(); k++)
                        worlds[sub[dev][k]].res = res[k];
                vx3b_gen_destroy(gen);
                return;
            }
            int rc = vx3b_ctx_create(dev, precision, &ctx);
            if (rc != VX3B_OK)
                return fail(rc);
            std::vector<VX3B_WorldDesc> descs;
            for (int i : sub[dev])
                descs.push_back(worlds[i].desc);
            if ((rc = vx3b_batch_upload(ctx, descs.data(), (int)descs.size())) != VX3B_OK)
                return fail(rc);
            if ((rc = vx3b_batch_run(ctx, 1000000, 1)) != VX3B_OK) // "Maximum Steps 1000000" (VX3_SimulationManager.cu:62)
                return fail(rc);
            std::vector<VX3B_WorldResult> res(descs.size());
            if ((rc = vx3b_batch_results(ctx, res.data())) != VX3B_OK)
                return fail(rc);
            for (size_t k = 0; k < sub[dev].size(); k++) {
                JobWorld &w = worlds[sub[dev][k]];
                w.res = res[k];
                if (w.desc.save_position_of_all_voxels) {
                    w.pos.resize((size_t)w.desc.num_voxels * 3);
                    w.initPos.resize((size_t)w.desc.num_voxels * 3);
                    vx3b_batch_read_state(ctx, (int)k, w.pos.data(), nullptr, nullptr, nullptr);
                    vx3b_batch_read_initial_pos(ctx, (int)k, w.initPos.data());
                }
            }
            vx3b_ctx_destroy(ctx);
        });
    }
    for (auto &t : threads)
        t.join();
    int rcAll = VX3B_OK;
    for (int dev = 0; dev < num_devices; dev++)
        if (status[dev] != VX3B_OK) {
            rcAll = status[dev];
            set_error("device " + std::to_string(dev) + ": " + errors[dev]);
        }
    if (rcAll == VX3B_OK) {
        std::vector<JobWorld *> order;
        for (auto &w : worlds)
            order.push_back(&w);
        std::stable_sort(order.begin(), order.end(), better);
        // report layout: vx3_node_worker.cu:98-141
        std::string inputName = dir;
        while (inputName.size() > 1 && inputName.back() == '/')
            inputName.pop_back();
        size_t slash = inputName.find_last_of('/');
        if (slash != std::string::npos)
            inputName = inputName.substr(slash + 1);
        XmlNode report;
        report.name = "report";
        auto add = [](XmlNode &p, const std::string &name, const std::string &text) -> XmlNode & {
            auto n = std::make_unique<XmlNode>();
            n->name = name;
            n->text = text;
            p.children.push_back(std::move(n));
            return *p.children.back();
        };
        add(report, "inputdir", inputName);
        XmlNode &best = add(report, "bestfit", "");
        add(best, "filename", order[0]->file);
        add(best, "fitness_score", fmt_double(order[0]->res.fitness_score));
        XmlNode &detail = add(report, "detail", "");
        for (JobWorld *w : order) {
            std::string name = w->file.substr(0, w->file.find('.'));
            XmlNode &n = add(detail, name, "");
            const VX3B_WorldResult &r = w->res;
            add(n, "currentTime", fmt_double(r.current_time));
            add(n, "fitness_score", fmt_double(r.fitness_score));
            add(n, "num_voxel", std::to_string(r.num_voxel));
            add(n, "num_measured_voxel", std::to_string(r.num_measured_voxel));
            add(n, "voxSize", fmt_double(r.vox_size));
            add(n, "numClosePairs", std::to_string(r.num_close_pairs));
            XmlNode &ic = add(n, "initialCenterOfMass", "");
            add(ic, "x", fmt_double(r.initial_com[0]));
            add(ic, "y", fmt_double(r.initial_com[1]));
            add(ic, "z", fmt_double(r.initial_com[2]));
            XmlNode &cc = add(n, "currentCenterOfMass", "");
            add(cc, "x", fmt_double(r.current_com[0]));
            add(cc, "y", fmt_double(r.current_com[1]));
            add(cc, "z", fmt_double(r.current_com[2]));
            add(n, "total_distance_of_all_voxels", fmt_double(r.total_distance_of_all_voxels));
            if (w->desc.save_position_of_all_voxels) {
                std::string ip, p, mats;
                for (int v = 0; v < w->desc.num_voxels; v++) { // std::to_string(double): "%f"
                    ip += std::to_string(w->initPos[v * 3]) + "," + std::to_string(w->initPos[v * 3 + 1]) + "," + std::to_string(w->initPos[v * 3 + 2]) + ";";
                    p += std::to_string(w->pos[v * 3]) + "," + std::to_string(w->pos[v * 3 + 1]) + "," + std::to_string(w->pos[v * 3 + 2]) + ";";
                    mats += std::to_string(w->desc.voxel_mats[w->desc.voxel_mat[v]].matid) + ";";
                }
                add(n, "init_pos", ip);
                add(n, "pos", p);
                add(n, "mats", mats);
            }
        }
        std::ofstream out(report_path, std::ios::binary);
        if (!out) {
            set_error(std::string("vx3b_job_run: cannot write ") + report_path);
            rcAll = VX3B_ERR_IO;
        } else
            out << xml_write(report, true);
    }
    for (auto &w : worlds)
        if (w.model)
            vx3b_model_free(w.model);
    return rcAll;
}
} // namespace vx3b

extern "C" int vx3b_job_run(const char *input_dir, const char *report_path, int num_devices, int precision) {
    using namespace vx3b;
    if (!input_dir || !report_path || num_devices < 1) {
        set_error("vx3b_job_run: bad argument");
        return VX3B_ERR_ARG;
    }
    const std::string dir = input_dir;
    return job_run(dir + "/base.vxa", dir, list_vxd(dir), report_path, num_devices, precision);
}
extern "C" int vx3b_job_run_files(const char *base_vxa_path, const char *input_dir, const char *const *vxd_files, int num_files, const char *report_path,
                                  int num_devices, int precision) {
    using namespace vx3b;
    if (!base_vxa_path || !input_dir || !report_path || num_devices < 1 || num_files < 0 || (num_files > 0 && !vxd_files)) {
        set_error("vx3b_job_run_files: bad argument");
        return VX3B_ERR_ARG;
    }
    return job_run(base_vxa_path, input_dir, to_list(vxd_files, num_files), report_path, num_devices, precision);
}

// ---- .history stream -----------------------------------------------------------------------------------------
// Replaces the history printf()s of CUDA_Simulation (VX3/VX3_SimulationManager.cu:40-114): same text, produced on the host from
// read-backs every real_stepsize steps instead of from device-side printf, one world at a time (the reference interleaves the
// output of concurrently running worlds at random).
namespace vx3b {

struct HistoryWorld {
    const VX3B_WorldDesc *d;
    std::vector<int> surface;            // voxels with a free face, creation order (VX3_VoxelyzeKernel.cu:495-513)
    std::vector<int> vlinks;             // [nv][6] link per direction or -1
    std::vector<double> pos, orient;
    std::vector<float> temp, strain;
};

// VX3_Voxel::cornerOffset (VX3/VX3_Voxel.cu:147-159) for the all-negative / all-positive corner, returned as float like the reference
static void corner_offset(const HistoryWorld &H, int v, bool positive, float out[3]) {
    const VX3B_WorldDesc &d = *H.d;
    const VX3B_VoxelMaterial &m = d.voxel_mats[d.voxel_mat[v]];
    const double base = m.nomSize * (double)(1 + H.temp[v] * m.alphaCTE); // baseSize(), VX3_Voxel.h:90-92
    for (int i = 0; i < 3; i++) {
        const int l = H.vlinks[(size_t)v * 6 + 2 * i + (positive ? 0 : 1)];
        double s = positive ? 1.0 : -1.0;
        if (l >= 0) {
            const VX3B_LinkMaterial &lm = d.link_mats[d.link_mat[l]];
            const bool failed = lm.epsilonFail != -1.0f && H.strain[l] > lm.epsilonFail;
            if (!failed) {
                // VX3_Link::axialStrain(positiveEnd), strainRatio = E(pos end) / E(neg end) (VX3_Link.cu:72-74)
                const float ratio = d.voxel_mats[d.voxel_mat[d.link_pos[l]]].E / d.voxel_mats[d.voxel_mat[d.link_neg[l]]].E;
                const float ax = positive ? 2.0f * H.strain[l] * ratio / (1.0f + ratio) : 2.0f * H.strain[l] / (1.0f + ratio);
                s = (1 + ax) * (positive ? 1 : -1);
            }
        }
        out[i] = (float)(0.5 * base * s);
    }
}

static void write_frame(FILE *f, const HistoryWorld &H, long long j, double t) {
    const VX3B_WorldDesc &d = *H.d;
    const double vs = 1 / 0.001;
    if (d.record_voxel) {
        fprintf(f, "<<<Step%lld Time:%f>>>", j, t);
        for (int v : H.surface) {
            const double *p = &H.pos[3 * (size_t)v], *q = &H.orient[4 * (size_t)v];
            fprintf(f, "%.1f,%.1f,%.1f,", p[0] * vs, p[1] * vs, p[2] * vs);
            const double angle = 2.0 * acos(q[0] > 1 ? 1 : q[0]) * 57.29577951308232; // Quat3D::AngleDegrees
            fprintf(f, "%.1f,%.2f,%.2f,%.2f,", angle, q[1], q[2], q[3]);
            float nnn[3], ppp[3];
            corner_offset(H, v, false, nnn);
            corner_offset(H, v, true, ppp);
            fprintf(f, "%.1f,%.1f,%.1f,%.1f,%.1f,%.1f,", nnn[0] * vs, nnn[1] * vs, nnn[2] * vs, ppp[0] * vs, ppp[1] * vs, ppp[2] * vs);
            fprintf(f, "%d,", d.voxel_mats[d.voxel_mat[v]].matid);
            fprintf(f, "%.1f,", 0.0); // localSignal: signals are not simulated (EnableSignals = 0)
            fprintf(f, ";");
        }
        fprintf(f, "<<<>>>");
    }
    if (d.record_link) {
        fprintf(f, "|[[[%lld]]]", j);
        for (int l = 0; l < d.num_links; l++) {
            const double *a = &H.pos[3 * (size_t)d.link_pos[l]], *b = &H.pos[3 * (size_t)d.link_neg[l]];
            fprintf(f, "%.4f,%.4f,%.4f,", a[0], a[1], a[2]);
            fprintf(f, "%.4f,%.4f,%.4f,", b[0], b[1], b[2]);
            fprintf(f, ";");
        }
        fprintf(f, "[[[]]]");
    }
    fprintf(f, "\n");
}

} // namespace vx3b

namespace vx3b {
static int job_history(const std::string &basePath, const std::string &dir, const std::vector<std::string> &files, const char *out_path, int precision) {
    std::string base;
    if (!read_file(basePath, &base)) {
        set_error("vx3b_job_history: cannot read " + basePath);
        return VX3B_ERR_IO;
    }
    FILE *f = out_path ? fopen(out_path, "w") : stdout;
    if (!f) {
        set_error(std::string("vx3b_job_history: cannot write ") + out_path);
        return VX3B_ERR_IO;
    }
    int rc = VX3B_OK;
    vx3b_ctx *ctx = nullptr;
    bool any = false; // does any world record? (no GPU work, no context otherwise)
    for (size_t i = 0; i < files.size() && !any; i++) {
        std::string vxd;
        vx3b_model *model = nullptr;
        if (read_file(dir + "/" + files[i], &vxd) && vx3b_model_from_vxa(base.c_str(), vxd.c_str(), &model) == VX3B_OK) {
            VX3B_WorldDesc d;
            vx3b_model_desc(model, &d);
            any = d.record_step_size != 0;
            vx3b_model_free(model);
        }
    }
    if (!any) {
        if (out_path)
            fclose(f);
        return VX3B_OK;
    }
    if ((rc = vx3b_ctx_create(0, precision, &ctx)) != VX3B_OK) {
        if (out_path)
            fclose(f);
        return rc;
    }
    vx3b_ctx_set_option(ctx, "path", 1); // the generic link + voxel kernels: the only ones that keep every link's last strain
    for (size_t i = 0; i < files.size() && rc == VX3B_OK; i++) {
        std::string vxd;
        vx3b_model *model = nullptr;
        if (!read_file(dir + "/" + files[i], &vxd) || vx3b_model_from_vxa(base.c_str(), vxd.c_str(), &model) != VX3B_OK) {
            fprintf(stderr, "%s: %s\n", files[i].c_str(), vx3b_last_error()); // the reference prints and carries on
            continue;
        }
        VX3B_WorldDesc d;
        vx3b_model_desc(model, &d);
        double dt = 0;
        vx3b_world_time_step(&d, &dt);
        if (d.record_step_size && d.num_voxels > 0 && (rc = vx3b_batch_upload(ctx, &d, 1)) == VX3B_OK) {
            HistoryWorld H;
            H.d = &d;
            H.vlinks.assign((size_t)d.num_voxels * 6, -1);
            for (int l = 0; l < d.num_links; l++) {
                H.vlinks[(size_t)d.link_neg[l] * 6 + 2 * d.link_axis[l]] = l;
                H.vlinks[(size_t)d.link_pos[l] * 6 + 2 * d.link_axis[l] + 1] = l;
            }
            for (int v = 0; v < d.num_voxels; v++)
                for (int k = 0; k < 6; k++)
                    if (H.vlinks[(size_t)v * 6 + k] < 0) {
                        H.surface.push_back(v);
                        break;
                    }
            H.pos.resize(3 * (size_t)d.num_voxels);
            H.orient.resize(4 * (size_t)d.num_voxels);
            H.temp.resize(d.num_voxels);
            H.strain.resize(std::max(d.num_links, 1));
            // header (VX3_SimulationManager.cu:40-58)
            fprintf(f, "\n{{{setting}}}<rescale>0.001</rescale>\n");
            for (int k = 0; k < d.num_voxel_mats; k++) {
                const uint32_t c = d.voxel_mats[k].rgba;
                fprintf(f, "{{{setting}}}<matcolor><id>%d</id><r>%.2f</r><g>%.2f</g><b>%.2f</b><a>%.2f</a></matcolor>\n", d.voxel_mats[k].matid,
                        (c & 255) / 255., ((c >> 8) & 255) / 255., ((c >> 16) & 255) / 255., ((c >> 24) & 255) / 255.);
            }
            fprintf(f, "\n{{{setting}}}<voxel_size>%f</voxel_size>\n", d.vox_size);
            const double recommended = d.dt_frac != 0 ? dt / d.dt_frac : 0; // dt = float(DtFrac * recommendedTimeStep)
            const int real_stepsize = (int)(d.record_step_size / (10000 * recommended * d.dt_frac)) + 1;
            fprintf(f, "real_stepsize: %d ; recommendedTimeStep %f; d_v3->DtFrac %f . \n", real_stepsize, recommended, d.dt_frac);
            // frames after step j+1 for j = 0, real_stepsize, 2*real_stepsize, ... while the stop condition allows (cap 1e6 steps)
            long long done = 0;
            for (long long j = 0; j < 1000000 && rc == VX3B_OK; j += real_stepsize) {
                const long long target = j + 1;
                if ((rc = vx3b_batch_run(ctx, target - done, 1)) != VX3B_OK)
                    break;
                VX3B_WorldResult r;
                if ((rc = vx3b_batch_results(ctx, &r)) != VX3B_OK)
                    break;
                if (r.steps < target || r.diverged) // stopped by its formula (or diverged) before reaching this frame
                    break;
                done = r.steps;
                if ((rc = vx3b_batch_read_state(ctx, 0, H.pos.data(), H.orient.data(), nullptr, nullptr)) != VX3B_OK ||
                    (rc = vx3b_batch_read_aux(ctx, 0, H.temp.data(), H.strain.data())) != VX3B_OK)
                    break;
                write_frame(f, H, j, r.current_time);
            }
        }
        vx3b_model_free(model);
    }
    vx3b_ctx_destroy(ctx);
    if (out_path)
        fclose(f);
    else
        fflush(f);
    return rc;
}
} // namespace vx3b

extern "C" int vx3b_job_history(const char *input_dir, const char *out_path, int precision) {
    using namespace vx3b;
    if (!input_dir) {
        set_error("vx3b_job_history: bad argument");
        return VX3B_ERR_ARG;
    }
    const std::string dir = input_dir;
    return job_history(dir + "/base.vxa", dir, list_vxd(dir), out_path, precision);
}
extern "C" int vx3b_job_history_files(const char *base_vxa_path, const char *input_dir, const char *const *vxd_files, int num_files, const char *out_path,
                                      int precision) {
    using namespace vx3b;
    if (!base_vxa_path || !input_dir || num_files < 0 || (num_files > 0 && !vxd_files)) {
        set_error("vx3b_job_history_files: bad argument");
        return VX3B_ERR_ARG;
    }
    return job_history(base_vxa_path, input_dir, to_list(vxd_files, num_files), out_path, precision);
}

// The task file the front-end hands to the worker (Voxelyze3.cpp:101-111; read back at vx3_node_worker.cu:64-75):
// <vxa>path of the base VXA</vxa> <input_dir>directory of the VXD files</input_dir> <vxd><f>file</f>...</vxd>
extern "C" int vx3b_job_run_task(const char *vxt_path, const char *report_path, int num_devices, int precision, int stream_history) {
    using namespace vx3b;
    if (!vxt_path || !report_path || num_devices < 1) {
        set_error("vx3b_job_run_task: bad argument");
        return VX3B_ERR_ARG;
    }
    std::string text, err;
    if (!read_file(vxt_path, &text)) {
        set_error(std::string("vx3b_job_run_task: cannot read ") + vxt_path);
        return VX3B_ERR_IO;
    }
    auto root = xml_parse(text, &err); // decodes the predefined entities boost's write_xml escapes paths with
    if (!root) {
        set_error(std::string("vx3b_job_run_task: ") + vxt_path + ": " + err);
        return VX3B_ERR_PARSE;
    }
    const XmlNode *vxa = root->child("vxa"), *dirNode = root->child("input_dir"), *vxd = root->child("vxd");
    if (!dirNode || xml_trim(dirNode->text).empty()) {
        set_error(std::string("vx3b_job_run_task: no input_dir in ") + vxt_path);
        return VX3B_ERR_PARSE;
    }
    const std::string dir = xml_trim(dirNode->text);
    const std::string base = vxa && !xml_trim(vxa->text).empty() ? xml_trim(vxa->text) : dir + "/base.vxa";
    std::vector<std::string> files;
    if (vxd)
        for (auto &c : vxd->children)
            files.push_back(xml_trim(c->text));
    if (stream_history) {
        int rc = job_history(base, dir, files, nullptr, precision);
        if (rc != VX3B_OK)
            return rc;
    }
    return job_run(base, dir, files, report_path, num_devices, precision);
}
