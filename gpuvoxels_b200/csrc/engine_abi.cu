// Engine part of the C ABI (include/vx3b200.h, section 2).
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "engine_base.hpp"
#include "host_dt.hpp"

namespace vx3b {
void set_error(const std::string &e);
#define CUDA_TRY(expr)                                                                                                          \
    do {                                                                                                                        \
        cudaError_t _e = (expr);                                                                                                \
        if (_e != cudaSuccess) {                                                                                                \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                                                      \
            return VX3B_ERR_CUDA;                                                                                               \
        }                                                                                                                       \
    } while (0)
} // namespace vx3b

struct vx3b_ctx {
    std::unique_ptr<vx3b::EngineBase> e;
};

extern "C" {

const char *vx3b_version(void) { return "vx3b200 0.1 (sm_100a)"; }

int vx3b_ctx_create(int device, int precision, vx3b_ctx **out) {
    using namespace vx3b;
    if (!out || (precision != VX3B_FP32 && precision != VX3B_FP64)) {
        set_error("vx3b_ctx_create: precision must be VX3B_FP32 or VX3B_FP64");
        return VX3B_ERR_ARG;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("vx3b_ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); this engine has no CPU fallback");
        return VX3B_ERR_CUDA;
    }
    if (device < 0 || device >= count) {
        set_error("vx3b_ctx_create: device index out of range");
        return VX3B_ERR_ARG;
    }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("vx3b_ctx_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
        return VX3B_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    auto *c = new vx3b_ctx();
    c->e.reset(precision == VX3B_FP32 ? make_engine_f32() : make_engine_f64());
    c->e->device = device;
    c->e->precision = precision;
    *out = c;
    return VX3B_OK;
}

void vx3b_ctx_destroy(vx3b_ctx *ctx) { delete ctx; }

int vx3b_ctx_set_option(vx3b_ctx *ctx, const char *name, double value) {
    if (!ctx || !name)
        return VX3B_ERR_ARG;
    std::string n = name;
    if (n == "path")
        ctx->e->path = (int)value;
    else if (n == "record_pairs")
        ctx->e->recordPairs = value != 0;
    else if (n == "chunk")
        ctx->e->chunk = std::max(1, (int)value);
    else if (n == "dt_override")
        ctx->e->dtOverride = value;
    else if (n == "resident_threads")
        ctx->e->residentThreads = (int)value;
    else if (n == "brick_x" || n == "brick_y" || n == "brick_z")
        ctx->e->brickDim[n[6] - 'x'] = std::max(1, std::min(64, (int)value));
    else if (n == "contact_list_cap")
        ctx->e->contactListCap = std::max(0, std::min(255, (int)value));
    else if (n == "contact_grid")
        ctx->e->contactGrid = value != 0;
    else if (n == "slab_fused")
        ctx->e->slabFused = value != 0;
    else if (n == "ext_all")
        ctx->e->extAll = value != 0;
    else if (n == "contact_skin")
        ctx->e->contactSkin = value < 0 ? 0 : value;
    else {
        vx3b::set_error("unknown option: " + n);
        return VX3B_ERR_ARG;
    }
    return VX3B_OK;
}

int vx3b_ctx_set_stream(vx3b_ctx *ctx, uint64_t s) {
    if (!ctx)
        return VX3B_ERR_ARG;
    if (ctx->e->ownStream && ctx->e->stream)
        cudaStreamDestroy(ctx->e->stream);
    ctx->e->stream = (cudaStream_t)s;
    ctx->e->ownStream = false;
    return VX3B_OK;
}

int vx3b_batch_upload(vx3b_ctx *ctx, const VX3B_WorldDesc *worlds, int n) { return ctx ? ctx->e->upload(worlds, n) : VX3B_ERR_ARG; }
int vx3b_batch_upload_models(vx3b_ctx *ctx, const vx3b_model *const *models, int n) {
    if (!ctx || n < 0 || (n > 0 && !models))
        return VX3B_ERR_ARG;
    std::vector<VX3B_WorldDesc> descs((size_t)n);
    for (int i = 0; i < n; i++) {
        int rc = models[i] ? vx3b_model_desc(models[i], &descs[i]) : VX3B_ERR_ARG;
        if (rc != VX3B_OK)
            return rc;
    }
    return ctx->e->upload(descs.data(), n);
}
int vx3b_batch_run(vx3b_ctx *ctx, int64_t max_steps, int use_stop) { return ctx ? ctx->e->run(max_steps, use_stop) : VX3B_ERR_ARG; }
int vx3b_world_time_step(const VX3B_WorldDesc *world, double *dt) {
    if (!world || !dt)
        return VX3B_ERR_ARG;
    *dt = (double)vx3b::world_time_step(*world);
    return VX3B_OK;
}
int vx3b_batch_set_force(vx3b_ctx *ctx, int world, int voxel, const double *f, const double *m) { return ctx ? ctx->e->set_force(world, voxel, f, m) : VX3B_ERR_ARG; }
int vx3b_batch_set_global_damping(vx3b_ctx *ctx, int world, int mat, double zeta) { return ctx ? ctx->e->set_global_damping(world, mat, zeta) : VX3B_ERR_ARG; }
int vx3b_batch_read_aux(vx3b_ctx *ctx, int world, float *t, float *s) { return ctx ? ctx->e->read_aux(world, t, s) : VX3B_ERR_ARG; }
int vx3b_batch_run_async(vx3b_ctx *ctx, int64_t steps) { return ctx ? ctx->e->run_async(steps) : VX3B_ERR_ARG; }
int vx3b_batch_halo_set(vx3b_ctx *ctx, int world, int list_id, const int32_t *voxels, int n) {
    return ctx ? ctx->e->halo_set(world, list_id, voxels, n) : VX3B_ERR_ARG;
}
int vx3b_batch_halo_record_bytes(vx3b_ctx *ctx) { return ctx ? ctx->e->halo_record_bytes() : VX3B_ERR_ARG; }
int vx3b_batch_halo_export(vx3b_ctx *ctx, int world, int list_id, void *buf) { return ctx ? ctx->e->halo_export(world, list_id, buf) : VX3B_ERR_ARG; }
int vx3b_batch_halo_import(vx3b_ctx *ctx, int world, int list_id, const void *buf) { return ctx ? ctx->e->halo_import(world, list_id, buf) : VX3B_ERR_ARG; }
int vx3b_batch_halo_open(vx3b_ctx *ctx, int world, int list_id, VX3B_HaloEndpoint *out) { return ctx && out ? ctx->e->halo_open(world, list_id, out) : VX3B_ERR_ARG; }
int vx3b_batch_halo_connect(vx3b_ctx *ctx, int world, int list_id, const VX3B_HaloEndpoint *peer, int same_process) {
    return ctx && peer ? ctx->e->halo_connect(world, list_id, peer, same_process) : VX3B_ERR_ARG;
}
int vx3b_batch_halo_slots(vx3b_ctx *ctx, int world, int list_id, int32_t *slots) { return ctx && slots ? ctx->e->halo_slots(world, list_id, slots) : VX3B_ERR_ARG; }
int vx3b_batch_halo_connect_fused(vx3b_ctx *ctx, int world, int list_id, const VX3B_HaloEndpoint *peer, const int32_t *peer_slots, int same_process) {
    return ctx && peer && peer_slots ? ctx->e->halo_connect(world, list_id, peer, same_process, peer_slots) : VX3B_ERR_ARG;
}
int vx3b_batch_run_slab(vx3b_ctx *ctx, int64_t steps) { return ctx ? ctx->e->run_slab(steps) : VX3B_ERR_ARG; }
int vx3b_batch_halo_timeouts(vx3b_ctx *ctx, int64_t *n) { return ctx && n ? ctx->e->halo_timeouts(n, nullptr) : VX3B_ERR_ARG; }
int vx3b_batch_halo_wait_ns(vx3b_ctx *ctx, int64_t *ns) {
    int64_t n;
    return ctx && ns ? ctx->e->halo_timeouts(&n, ns) : VX3B_ERR_ARG;
}
int vx3b_batch_results(vx3b_ctx *ctx, VX3B_WorldResult *out) { return ctx ? ctx->e->results(out) : VX3B_ERR_ARG; }
int vx3b_batch_read_state(vx3b_ctx *ctx, int world, double *p, double *o, double *l, double *a) {
    return ctx ? ctx->e->read_state(world, p, o, l, a) : VX3B_ERR_ARG;
}
int vx3b_batch_read_initial_pos(vx3b_ctx *ctx, int world, double *p) { return ctx ? ctx->e->read_initial_pos(world, p) : VX3B_ERR_ARG; }
int vx3b_batch_read_links(vx3b_ctx *ctx, int world, double *p2, double *a1, double *a2, int32_t *f) {
    return ctx ? ctx->e->read_links(world, p2, a1, a2, f) : VX3B_ERR_ARG;
}
int vx3b_batch_collision_pairs(vx3b_ctx *ctx, int world, int32_t *pairs, int32_t *n) {
    return ctx ? ctx->e->collision_pairs(world, pairs, n) : VX3B_ERR_ARG;
}
int vx3b_batch_num_links(vx3b_ctx *ctx, int world, int32_t *n) { return ctx ? ctx->e->num_links(world, n) : VX3B_ERR_ARG; }
int vx3b_batch_read_signals(vx3b_ctx *ctx, int world, double *sig6, int32_t *vflags2) { return ctx ? ctx->e->read_signals(world, sig6, vflags2) : VX3B_ERR_ARG; }
int vx3b_batch_read_topology(vx3b_ctx *ctx, int world, int32_t *links6, int32_t *n) { return ctx ? ctx->e->read_topology(world, links6, n) : VX3B_ERR_ARG; }
int vx3b_ctx_timing(vx3b_ctx *ctx, float *ms, int64_t *launches) {
    if (!ctx)
        return VX3B_ERR_ARG;
    if (ms)
        *ms = ctx->e->lastMs;
    if (launches)
        *launches = ctx->e->lastLaunches;
    return VX3B_OK;
}
int64_t vx3b_ctx_total_voxel_steps(vx3b_ctx *ctx) { return ctx ? ctx->e->lastVoxelSteps : 0; }

} // extern "C"
