// A generation of independent worlds on ONE device, pipelined (C ABI section 3b).
// Replaces what VX3_SimulationManager::start does for one device - readVXD, startKernel, collectResults one after the other
// (Voxelyze3/src/VX3/VX3_SimulationManager.cu:138-155, 277-458) - with a software pipeline: the worlds are cut (in descending order of cost) into `parts`
// sub-batches, each with its own context, stream and host thread; the sub-batches are uploaded one after the other (each upload
// uses all host cores), and a sub-batch starts stepping the moment it is resident, so the GPU works while the rest of the
// generation is still being staged and copied.  The step kernels of the sub-batches share the SMs (k_resident: one CTA per world).
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vx3b200.h"

namespace vx3b {
void set_error(const std::string &e);
}

struct vx3b_gen {
    int device = 0, precision = 32;
    std::vector<vx3b_ctx *> ctx;
    int64_t lastVoxelSteps = 0;
};

extern "C" {

int vx3b_gen_create(int device, int precision, int parts, vx3b_gen **out) {
    if (!out || parts < 1 || parts > 64) {
        vx3b::set_error("vx3b_gen_create: parts must be 1..64");
        return VX3B_ERR_ARG;
    }
    auto g = new vx3b_gen;
    g->device = device, g->precision = precision;
    for (int k = 0; k < parts; k++) {
        vx3b_ctx *c = nullptr;
        int rc = vx3b_ctx_create(device, precision, &c);
        if (rc != VX3B_OK) {
            for (auto *p : g->ctx)
                vx3b_ctx_destroy(p);
            delete g;
            return rc;
        }
        g->ctx.push_back(c);
    }
    *out = g;
    return VX3B_OK;
}

void vx3b_gen_destroy(vx3b_gen *g) {
    if (!g)
        return;
    for (auto *p : g->ctx)
        vx3b_ctx_destroy(p);
    delete g;
}

int vx3b_gen_set_option(vx3b_gen *g, const char *name, double value) {
    if (!g)
        return VX3B_ERR_ARG;
    for (auto *p : g->ctx) {
        int rc = vx3b_ctx_set_option(p, name, value);
        if (rc != VX3B_OK)
            return rc;
    }
    return VX3B_OK;
}

int64_t vx3b_gen_total_voxel_steps(vx3b_gen *g) { return g ? g->lastVoxelSteps : 0; }

int vx3b_gen_run(vx3b_gen *g, const vx3b_model *const *models, int n, int64_t max_steps, int use_stop_formula, VX3B_WorldResult *results) {
    if (!g || !models || n < 1 || !results) {
        vx3b::set_error("vx3b_gen_run: null argument or no worlds");
        return VX3B_ERR_ARG;
    }
    const int parts = std::min((int)g->ctx.size(), n);
    // Sub-batch k = the k-th run of the worlds in descending order of cost.  The step kernel's CTAs (one per world) are handed to the SMs
    // in launch order, so across the sub-batches the biggest worlds still go first and the small ones fill the tail, as in one batch;
    // dealing the worlds round-robin (every sub-batch a mix of sizes) was measured and lost 10 %: big worlds of the later sub-batches
    // then start late and stick out at the end.
    std::vector<int64_t> cost(n);
    for (int i = 0; i < n; i++) {
        VX3B_WorldDesc d;
        int rc = vx3b_model_desc(models[i], &d);
        if (rc != VX3B_OK)
            return rc;
        cost[i] = (int64_t)d.num_voxels + 2 * (int64_t)d.num_links;
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; i++)
        order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
    std::vector<std::vector<int>> sub(parts);
    for (int q = 0; q < n; q++)
        sub[(int64_t)q * parts / n].push_back(order[q]);
    std::mutex mu;
    std::condition_variable cv;
    int turn = 0; // whose upload it is
    std::vector<int> status(parts, VX3B_OK);
    std::vector<std::string> errors(parts);
    std::vector<int64_t> vsteps(parts, 0);
    auto work = [&](int k) {
        std::vector<const vx3b_model *> mine;
        for (int i : sub[k])
            mine.push_back(models[i]);
        {
            std::unique_lock<std::mutex> lock(mu);
            cv.wait(lock, [&] { return turn == k; });
        }
        int r = mine.empty() ? VX3B_OK : vx3b_batch_upload_models(g->ctx[k], mine.data(), (int)mine.size());
        {
            std::lock_guard<std::mutex> lock(mu);
            turn = k + 1; // the next sub-batch is staged while this one steps
        }
        cv.notify_all();
        if (r == VX3B_OK && !mine.empty())
            r = vx3b_batch_run(g->ctx[k], max_steps, use_stop_formula);
        std::vector<VX3B_WorldResult> res(mine.size());
        if (r == VX3B_OK && !mine.empty())
            r = vx3b_batch_results(g->ctx[k], res.data());
        if (r != VX3B_OK) {
            status[k] = r;
            errors[k] = vx3b_last_error();
            return;
        }
        for (size_t q = 0; q < sub[k].size(); q++)
            results[sub[k][q]] = res[q];
        vsteps[k] = mine.empty() ? 0 : vx3b_ctx_total_voxel_steps(g->ctx[k]);
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < parts; k++)
        threads.emplace_back(work, k);
    work(0);
    for (auto &t : threads)
        t.join();
    g->lastVoxelSteps = 0;
    for (int k = 0; k < parts; k++) {
        if (status[k] != VX3B_OK) {
            vx3b::set_error("sub-batch " + std::to_string(k) + ": " + errors[k]);
            return status[k];
        }
        g->lastVoxelSteps += vsteps[k];
    }
    return VX3B_OK;
}

} // extern "C"
