/* vx3b200.h — C ABI of the B200-native Voxelyze step engine (libvx3b200.so).
 *
 * This is the drop-in boundary for the hot path of liusida/gpuVoxels:
 * everything the reference does between "host model built from VXA/VXD" and "results
 * copied back" — i.e. what VX3_SimulationManager::readVXD/startKernel/collectResults
 * (Voxelyze3/src/VX3/VX3_SimulationManager.cu:277-458) hand to and take from the
 * CUDA_Simulation kernel (VX3_SimulationManager.cu:11-121).
 *
 * Conventions: extern "C", plain structs, int status (0 = ok, negative = error; text via
 * vx3b_last_error()), caller owns host buffers, the library owns device memory, one
 * context per device, one host thread per context.  No torch / C++ types cross this line.
 * There is NO CPU fallback: every compute entry point fails with VX3B_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Three layers are exported:
 *   1. host model   (vx3b_model_*)  VXA/VXD text or CVoxelyze-style calls  -> VX3B_WorldDesc
 *   2. step engine  (vx3b_ctx_*, vx3b_batch_*)  VX3B_WorldDesc[] -> stepped state/results
 *   3. job          (vx3b_job_run)  directory of base.vxa + *.vxd -> report XML
 */
#ifndef VX3B200_H
#define VX3B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VX3B_OK 0
#define VX3B_ERR_ARG (-1)
#define VX3B_ERR_CUDA (-2)
#define VX3B_ERR_PARSE (-3)
#define VX3B_ERR_IO (-4)
#define VX3B_ERR_STATE (-5)

/* precision of the device state and arithmetic */
#define VX3B_FP32 32 /* production: float4 SoA, displacement-based positions          */
#define VX3B_FP64 64 /* validation build: double state with the reference's float casts */

/* link axis / direction enums: VoxCAD/src/VX2/include/types.h:6-56 */
#define VX3B_AXIS_X 0
#define VX3B_AXIS_Y 1
#define VX3B_AXIS_Z 2
/* dof bits (CVX_External::dofFixed, types.h): X_TRANSLATE=1<<0 .. Z_ROTATE=1<<5 */
#define VX3B_DOF_TX 1
#define VX3B_DOF_TY 2
#define VX3B_DOF_TZ 4
#define VX3B_DOF_RX 8
#define VX3B_DOF_RY 16
#define VX3B_DOF_RZ 32
#define VX3B_DOF_ALL 63

/* Formula VM opcodes — same numbering as VX3_MathTreeOperator (Utils/VX3_MathTree.h:4-29). */
enum {
    VX3B_MT_END = 0, VX3B_MT_CONST, VX3B_MT_E, VX3B_MT_PI, VX3B_MT_VAR, VX3B_MT_ADD, VX3B_MT_SUB, VX3B_MT_MUL,
    VX3B_MT_DIV, VX3B_MT_POW, VX3B_MT_SQRT, VX3B_MT_SIN, VX3B_MT_COS, VX3B_MT_TAN, VX3B_MT_ATAN, VX3B_MT_LOG,
    VX3B_MT_INT, VX3B_MT_ABS, VX3B_MT_NOT, VX3B_MT_GREATERTHAN, VX3B_MT_LESSTHAN, VX3B_MT_AND, VX3B_MT_OR,
    VX3B_MT_NORMALCDF
};
#define VX3B_MAX_TOKENS 64 /* the reference allows 1024 (VX3_VoxelyzeKernel.cuh:113-119); no shipped formula exceeds 40 */

typedef struct VX3B_Token {
    int32_t op;   /* VX3B_MT_* */
    int32_t pad;
    double value; /* constant, or variable id 0..8 = x,y,z,hit,t,angle,targetCloseness,numClosePairs,num_voxel */
} VX3B_Token;

/* One palette material with the derived constants exactly as the reference caches them
 * (float storage!): CVX_Material (Voxelyze3/src/old/VX_Material.h), CVX_MaterialVoxel::updateDerived
 * (VX3_MaterialVoxel.cu:49-71). */
typedef struct VX3B_VoxelMaterial {
    float E, nu, rho, alphaCTE, muStatic, muKinetic, zetaInternal, zetaGlobal, zetaCollision;
    float epsilonFail;           /* -1 = never fails */
    float gravMult;              /* multiples of g; gravity force = -mass*9.80665f*gravMult */
    float mass, massInverse, sqrtMass, momentInertia, momentInertiaInverse;
    float _2xSqMxExS, _2xSqIxExSxSxS;
    double nomSize;
    double Cilia;                /* scale of the cilia force */
    int32_t matid;               /* palette ID reported in <mats> */
    int32_t fixed, isTarget, isMeasured, sticky, linear; /* linear = 0: bilinear or data model (the link materials carry the curves) */
    uint32_t rgba;               /* display colour r | g << 8 | b << 16 | a << 24, each (int)(VXA value * 255) (VX_Object.h:419-425): .history only */
    /* signals (VX3_Voxel.cu:279-348; defaults of VXA/VX_Object.cpp:1420-1424) and the "secondary experiment" switches
     * (VX3_VoxelyzeKernel.cu:365-399, 632-634, 852-854); times in seconds, 0 = off */
    int32_t isPaceMaker, padSignal;
    double PaceMakerPeriod, signalValueDecay, signalTimeDelay, inactivePeriod;
    double RemoveFromSimulationAfterThisManySeconds, TurnOnThermalExpansionAfterThisManySeconds, TurnOnCiliaAfterThisManySeconds;
} VX3B_VoxelMaterial;

/* Combined material of a voxel pair: CVX_MaterialLink::updateAll/updateDerived
 * (VX3_MaterialLink.cu:53-148). */
#define VX3B_MAX_SS_POINTS 16 /* points of a stress/strain model, the leading (0, 0) included */
typedef struct VX3B_LinkMaterial {
    float E, nu, eHat;
    float a1, a2, b1, b2, b3;
    float sqA1, sqA2xIp, sqB1, sqB2xFMp, sqB3xIp;
    float epsilonFail;           /* -1 = never fails */
    int32_t mat1, mat2;          /* indices into voxel_mats */
    /* stress/strain model (CVX_Material::strainData / stressData, VX3_Material.cu:90-156): linear = 1: the single line E;
     * else num_points points from (0, 0), piecewise linear, extrapolated along the last segment; a link backs off along E from
     * its largest strain (VX3_Link.cu:232-268) */
    int32_t linear, num_points;
    float strain_data[VX3B_MAX_SS_POINTS], stress_data[VX3B_MAX_SS_POINTS];
} VX3B_LinkMaterial;

/* One world (= one VXA⊕VXD), flattened.  Voxels are in the reference's creation order
 * (x fastest, then y, then z over occupied lattice cells — VX_Sim.cpp:92-107), links in the
 * order CVoxelyze::addVoxel creates them (Voxelyze.cpp:439-461): that order defines the
 * surface list and therefore the collision pair enumeration. */
typedef struct VX3B_WorldDesc {
    int32_t num_voxels, num_links, num_voxel_mats, num_link_mats;
    double vox_size;

    const int16_t *lattice;      /* [nv][3] ix,iy,iz */
    const int32_t *voxel_mat;    /* [nv] index into voxel_mats */
    const double *phase_offset;  /* [nv] or NULL */
    const double *base_cilia;    /* [nv][3] or NULL */
    const double *shift_cilia;   /* [nv][3] or NULL (needs signals; accepted, unused while EnableSignals=0) */
    const uint8_t *dof_fixed;    /* [nv] VX3B_DOF_* mask or NULL */
    const float *ext_force;      /* [nv][3] or NULL (float, as CVX_External stores it) */
    const float *ext_moment;     /* [nv][3] or NULL */
    const double *ext_translation; /* [nv][3] or NULL prescribed displacement of fixed dofs */
    const double *ext_rotation;  /* [nv][3] or NULL rotation vector of fixed rotational dofs */

    const int32_t *link_neg;     /* [nl] voxel index of the negative end */
    const int32_t *link_pos;     /* [nl] voxel index of the positive end */
    const uint8_t *link_axis;    /* [nl] VX3B_AXIS_* */
    const int32_t *link_mat;     /* [nl] index into link_mats */

    const VX3B_VoxelMaterial *voxel_mats;
    const VX3B_LinkMaterial *link_mats;

    /* world scalars (VX3_VoxelyzeKernel.cuh:50-161, VX3_SimulationManager.cu:323-375) */
    double dt_frac;
    int32_t floor_enabled;
    int32_t vary_temp_enabled;
    double temp_amplitude, temp_period;
    int32_t enable_collision, enable_cilia, enable_expansion;
    int32_t vx2_compat;          /* 1: Voxelyze2 semantics (float dt, no angMom zeroing, no Fixed-material skip) */
    double max_dist_in_voxel_lengths_to_count_as_pair;
    int32_t save_position_of_all_voxels;
    float initial_temp;          /* temperature offset every voxel starts with (CVX_Sim::UpdateMatTemps at Import) */

    VX3B_Token stop_condition[VX3B_MAX_TOKENS]; /* empty (op[0]==END) = never stops by formula */
    VX3B_Token fitness[VX3B_MAX_TOKENS];
    VX3B_Token force_field[3][VX3B_MAX_TOKENS]; /* x,y,z; empty = no field */

    /* One body over several GPUs (section 2b): [nv] or NULL; 1 = ghost, a copy of a voxel another GPU's slab owns.  Ghosts are
     * never integrated; their state arrives through vx3b_batch_halo_import.  Links between two ghosts must not be listed. */
    const uint8_t *ghost;

    /* VXA.Simulator.RecordHistory (VX3_SimulationManager.cu:344-346): 0 = no .history stream */
    int32_t record_step_size, record_voxel, record_link;

    /* Topology-changing features and signals (VX3_SimulationManager.cu:325-371).  Worlds that use any of them are
     * stepped by the engine's dynamic-topology kernel, in the serial thread order of oracle/_ref/vx3_ref. */
    int32_t enable_signals;      /* VXA.Simulator.EnableSignals: target hits and pace makers raise signals that travel along links */
    int32_t enable_attach;       /* AttachDetach.EnableAttach: sticky voxels of one material that touch grow a link */
    int32_t enable_detach;       /* AttachDetach.EnableDetach: failed links let go */
    int32_t safety_guard;        /* AttachDetach.SafetyGuard: steps a new link's forces are scaled by 0.01 (VX3_Link.cu:206-213) */
    int32_t secondary_experiment; /* VXA.Simulator.SecondaryExperiment: removal of materials / re-initialised initial position */
    int32_t pad_dyn;
    double reinitialize_initial_position_after; /* seconds; with secondary_experiment */
    VX3B_Token attach_condition[5][VX3B_MAX_TOKENS]; /* all five must be > 0 at a voxel's position for it to attach; empty = CONST 0... see host */
} VX3B_WorldDesc;

/* Numeric fields of VX3_SimulationResult (VX3_SimulationResult.h:4-24). */
typedef struct VX3B_WorldResult {
    double current_time;
    double fitness_score;
    double initial_com[3];
    double current_com[3];
    double total_distance_of_all_voxels;
    double vox_size;
    double recent_angle, target_closeness;
    double dt;                   /* the (float-truncated) step actually used */
    int64_t steps;               /* steps taken (CurStepCount) */
    int32_t num_voxel, num_measured_voxel, num_close_pairs, collision_count;
    int32_t diverged, stopped;   /* diverged: some link strain > 100; stopped: stop formula > 0 */
    int32_t contact_rebuilds, contact_overflows; /* diagnostics of the resident contact broad phase (0 on the streaming path) */
} VX3B_WorldResult;

typedef struct vx3b_ctx vx3b_ctx;
typedef struct vx3b_model vx3b_model;

const char *vx3b_last_error(void);
const char *vx3b_version(void);

/* ---------------------------------------------------------------- 1. host model ------
 * Replaces CVX_Sim::ReadVXA + Import (Voxelyze3/src/VXA/VX_Sim.cpp:155-183, 59-153) and the
 * VXD⊕VXA merge (Voxelyze3/src/Utils/ctool.h:49-57, VX3_SimulationManager.cu:277-380). */
int vx3b_model_from_vxa(const char *vxa_xml, const char *vxd_xml /* may be NULL */, vx3b_model **out);
int vx3b_model_from_files(const char *vxa_path, const char *vxd_path /* may be NULL */, vx3b_model **out);
/* CVoxelyze-style programmatic construction (VoxCAD/src/VX2/include/Voxelyze.h:43-134) */
int vx3b_model_new(double vox_size, vx3b_model **out);
int vx3b_model_add_material(vx3b_model *m, double youngs_modulus, double density); /* returns material index >= 0 */
int vx3b_model_set_material_param(vx3b_model *m, int mat, const char *name, double value);
/* CVX_Material::setModelBilinear (failure_stress -1 = none) and setModel (points from or after (0, 0), ascending) on a builder
 * material (old/VX_Material.cpp:300-449); links of such voxels step with the non-linear stress model (VX3_Link.cu:232-268). */
int vx3b_model_set_material_bilinear(vx3b_model *m, int mat, double youngs_modulus, double plastic_modulus, double yield_stress, double failure_stress);
int vx3b_model_set_material_data(vx3b_model *m, int mat, int num_points, const double *strain, const double *stress);
int vx3b_model_set_voxel(vx3b_model *m, int mat, int ix, int iy, int iz);
int vx3b_model_set_fixed(vx3b_model *m, int ix, int iy, int iz, int dof_mask, const double *translation3, const double *rotation3);
int vx3b_model_set_force(vx3b_model *m, int ix, int iy, int iz, const double *force3, const double *moment3);
int vx3b_model_set_world_param(vx3b_model *m, const char *name, double value); /* gravity, floor, dt_frac, ... */
int vx3b_model_finalize(vx3b_model *m); /* builds links + combined materials + derived constants */
int vx3b_model_desc(const vx3b_model *m, VX3B_WorldDesc *out); /* pointers stay owned by the model */
int vx3b_model_merged_vxa(const vx3b_model *m, char *buf, int64_t cap, int64_t *needed); /* merged VXA text (for the reference harness) */
void vx3b_model_free(vx3b_model *m);

/* ONE SLAB of a body (section 2b; no reference counterpart - a reference world never spans devices).  Parses the same VXA (+ VXD)
 * but keeps only the voxels whose lattice coordinate along `axis` (0 x, 1 y, 2 z) falls into range `part` of `n_parts` ranges
 * holding (nearly) equal numbers of voxels, plus - flagged in VX3B_WorldDesc.ghost - the voxels of the two neighbouring slabs that
 * a cut link joins to one of its own.  Every process builds only its own slab: nothing of the size of the whole body but one
 * byte per lattice cell.  Voxels keep the creation order of the whole body (VX_Sim.cpp:92-107). */
typedef struct VX3B_SlabInfo {
    int32_t part, n_parts, axis;
    int32_t layer_lo, layer_hi;   /* own lattice layers [layer_lo, layer_hi) along axis */
    int32_t num_own;              /* voxels of this model that are not ghosts */
    int32_t n_send_lo, n_send_hi, n_ghost_lo, n_ghost_hi;
    int64_t body_voxels, body_links; /* of the whole body */
    double body_dt;               /* the time step the WHOLE body takes (context option "dt_override" of every slab) */
    /* exchange lists, voxel indices of THIS model, ascending = in the whole body's creation order, so a sender's list and the
     * receiver's ghost list name the same voxels in the same order; owned by the model */
    const int32_t *send_lo, *send_hi, *ghost_lo, *ghost_hi;
    const int64_t *global_index;  /* [num_voxels] voxel of this model -> voxel of the whole body */
} VX3B_SlabInfo;
int vx3b_model_from_vxa_slab(const char *vxa_xml, const char *vxd_xml /* may be NULL */, int axis, int part, int n_parts, vx3b_model **out);
int vx3b_model_slab_info(const vx3b_model *m, VX3B_SlabInfo *out);

/* ---------------------------------------------------------------- 2. step engine -----
 * Replaces VX3_VoxelyzeKernel (host ctor, VX3_VoxelyzeKernel.cu:27-105) + CUDA_Simulation
 * + doTimeStep and all gpu_update_* kernels (VX3_VoxelyzeKernel.cu:237-359, 566-650, 833-859). */
int vx3b_ctx_create(int device, int precision, vx3b_ctx **out);
void vx3b_ctx_destroy(vx3b_ctx *ctx);
/* 0 = automatic; 1 = generic streaming kernels (link pass + voxel pass); 2 = resident kernel
 * (one CTA per world, state in shared memory); 3 = fused tile kernel (big bodies). */
int vx3b_ctx_set_option(vx3b_ctx *ctx, const char *name, double value);
int vx3b_batch_upload(vx3b_ctx *ctx, const VX3B_WorldDesc *worlds, int n);
/* The same for n finalized models (saves a caller in another language from marshalling n 5-KB descriptors; this is what
 * VX3_SimulationManager::readVXD + startKernel do with their per-file CVX_Sim objects, VX3_SimulationManager.cu:277-414). */
int vx3b_batch_upload_models(vx3b_ctx *ctx, const vx3b_model *const *models, int n);
/* Advance every world by up to max_steps more steps.  use_stop_formula: test the world's stop
 * condition before every step as CUDA_Simulation does (VX3_SimulationManager.cu:62-69).
 * Worlds that diverge (any link strain > 100) stop, as VX3_VoxelyzeKernel.cu:271-280. */
int vx3b_batch_run(vx3b_ctx *ctx, int64_t max_steps, int use_stop_formula);
int vx3b_batch_results(vx3b_ctx *ctx, VX3B_WorldResult *out /* [n] */);
/* state of one world in creation order, always as double; any pointer may be NULL */
int vx3b_batch_read_state(vx3b_ctx *ctx, int world, double *pos3, double *orient4 /* w,x,y,z */, double *lin_mom3, double *ang_mom3);
int vx3b_batch_read_initial_pos(vx3b_ctx *ctx, int world, double *pos3);
int vx3b_batch_read_links(vx3b_ctx *ctx, int world, double *pos2_3, double *angle1v3, double *angle2v3, int32_t *flags /* bit0 smallAngle bit1 localVelocityValid */);
/* collision pair set found in the LAST executed step of `world`: pairs (i,j), i>j, voxel indices in
 * creation order, sorted ascending by (i,j).  *n in: capacity (pairs), out: count. */
int vx3b_batch_collision_pairs(vx3b_ctx *ctx, int world, int32_t *pairs2, int32_t *n);
/* Dynamic topology.  Links in use in `world` (it grows by attachment; vx3b_batch_read_links and the link_strain of
 * vx3b_batch_read_aux write that many entries).  Signal state per voxel: sig6 [nv][6] = localSignal, localSignaldt, inactiveUntil,
 * d_signal.value, d_signal.activeTime, packmakerNextPulse (VX3_Voxel.h); vflags2 [nv][2] = removed, enableAttach; either may be NULL.
 * Current link list: *n in = capacity in links, out = links in use; links6 [n][6] = negative end, positive end, axis, removed,
 * detached, isNewLink (steps left of the SafetyGuard easing); links6 may be NULL to ask for the count. */
int vx3b_batch_num_links(vx3b_ctx *ctx, int world, int32_t *n);
int vx3b_batch_read_signals(vx3b_ctx *ctx, int world, double *sig6, int32_t *vflags2);
int vx3b_batch_read_topology(vx3b_ctx *ctx, int world, int32_t *links6, int32_t *n);
/* what the .history writer needs beyond read_state: voxel temperatures [nv] and link strains [nl]; either may be NULL */
int vx3b_batch_read_aux(vx3b_ctx *ctx, int world, float *voxel_temp, float *link_strain);
/* Change the external force / moment of one voxel of an uploaded world between runs (CVX_External::setForce / setMoment,
 * VoxCAD/src/VX2/include/VX_External.h; the reference's own tests do this between doTimeStep calls, tVoxelyze.h:471-605,
 * 1033-1160).  Either pointer may be NULL (unchanged).  The voxel must have had a boundary condition at upload, or the
 * context option "ext_all" = 1 must have been set before vx3b_batch_upload. */
int vx3b_batch_set_force(vx3b_ctx *ctx, int world, int voxel, const double *force3, const double *moment3);
/* CVX_Material::setGlobalDamping of one voxel material of an uploaded world between runs (the reference's kineticFriction test,
 * tVoxelyze.h:1111-1158, lowers it once the voxel has settled). */
int vx3b_batch_set_global_damping(vx3b_ctx *ctx, int world, int mat, double zeta);
/* device time (CUDA events) of the last vx3b_batch_run, and how many kernels it launched */
int vx3b_ctx_timing(vx3b_ctx *ctx, float *device_ms, int64_t *kernel_launches);
int64_t vx3b_ctx_total_voxel_steps(vx3b_ctx *ctx); /* sum over worlds of nv*steps in the last run */
/* torch interop: run on an existing stream (cudaStream_t as integer); 0 = own stream */
int vx3b_ctx_set_stream(vx3b_ctx *ctx, uint64_t cuda_stream);

/* ---------------------------------------------------------------- 2b. one body over several GPUs ----
 * No reference counterpart: a reference world never spans devices.  The body is cut into slabs; each
 * slab is uploaded as one world holding its own voxels plus one ghost layer per cut face (VX3B_WorldDesc.ghost) and every link
 * with at least one own end - a cut link is stepped on both sides from identical inputs, so its history stays identical without
 * a force return trip.  Per time step: vx3b_batch_run_async(ctx, 1), export the boundary voxels the neighbours hold as ghosts,
 * exchange the buffers (NCCL send/recv on the same stream), import into the own ghosts.
 * State stays in the engine's internal order between asynchronous runs; every read-back call synchronises it first. */
/* host-only: the time step a world takes when uploaded (VX3: (float)(DtFrac * recommendedTimeStep), VX3_VoxelyzeKernel.cu:184-254).
 * Slabs of one body must all step with the WHOLE body's value: pass it to each slab context as option "dt_override". */
int vx3b_world_time_step(const VX3B_WorldDesc *world, double *dt);
int vx3b_batch_run_async(vx3b_ctx *ctx, int64_t steps); /* enqueue `steps` steps on the context's stream: no host sync, no stop formula */
/* register voxel list `list_id` (0..7) of `world`: creation-order voxel indices, copied */
int vx3b_batch_halo_set(vx3b_ctx *ctx, int world, int list_id, const int32_t *voxels, int n);
int vx3b_batch_halo_record_bytes(vx3b_ctx *ctx); /* bytes per voxel record: pos 3 x f64, temperature | flags, orientation 4 x f32/f64 */
/* pack / unpack the current records of a registered list into / from a DEVICE buffer of n * record_bytes, on the context's stream */
int vx3b_batch_halo_export(vx3b_ctx *ctx, int world, int list_id, void *device_buffer);
int vx3b_batch_halo_import(vx3b_ctx *ctx, int world, int list_id, const void *device_buffer);
/* Device-side exchange: no host, no NCCL call in the time-step loop.  A slab's ghost list gets a RECEIVE AREA in its own device
 * memory (two record buffers + a sequence flag); the neighbour that owns those voxels maps the area (CUDA IPC across processes, a
 * plain pointer inside one process; stores travel over NVLink) and its k_halo_push writes each step's records straight into it and
 * raises the flag; k_halo_pull on the receiving side waits for the flag and refreshes the ghosts.  vx3b_batch_run_slab enqueues,
 * per time step, the step kernel, the pushes and the pulls (captured in one CUDA graph per 64 steps). */
typedef struct VX3B_HaloEndpoint {
    uint8_t ipc_handle[64];  /* cudaIpcMemHandle_t of the receive area */
    uint64_t device_ptr;     /* its address in the creating process (used when same_process != 0) */
    int32_t device, n;       /* CUDA device of the area; records per buffer */
    int32_t record_bytes, pad;
    /* for the exchange FUSED into the step kernel (below): the brick-ordered position / orientation arrays of both buffer sets
     * (set 0: pos.xy, pos.z|temperature|flags, orientation; then set 1), each an allocation of its own; 0 = not available */
    uint8_t state_ipc_handle[6][64];
    uint64_t state_ptr[6];
} VX3B_HaloEndpoint;
/* create (or return) the receive area of ghost list `ghost_list_id` */
int vx3b_batch_halo_open(vx3b_ctx *ctx, int world, int ghost_list_id, VX3B_HaloEndpoint *out);
/* connect send list `send_list_id` to the neighbour's receive area (its record count and record size must match) */
int vx3b_batch_halo_connect(vx3b_ctx *ctx, int world, int send_list_id, const VX3B_HaloEndpoint *peer, int same_process);
/* The exchange FUSED into the step kernel - the default when every connected list was connected with vx3b_batch_halo_connect_fused:
 * the step kernel itself stores the new state of a boundary voxel into the neighbour's (mapped) brick-ordered arrays, at the slot
 * the neighbour keeps for its ghost copy, raises the neighbour's flag when the launch is complete, and starts by waiting for its own
 * flags.  ONE launch per time step, nothing staged.  vx3b_batch_halo_slots: the internal slots of a registered list ([n] int32), which
 * the neighbour passes as peer_slots (same order as its send list). */
int vx3b_batch_halo_slots(vx3b_ctx *ctx, int world, int list_id, int32_t *slots);
int vx3b_batch_halo_connect_fused(vx3b_ctx *ctx, int world, int send_list_id, const VX3B_HaloEndpoint *peer, const int32_t *peer_slots, int same_process);
/* `steps` time steps, each followed by the push of every connected send list and the pull of every opened ghost list; asynchronous
 * like vx3b_batch_run_async.  Every slab of the body must be given the same number of steps. */
int vx3b_batch_run_slab(vx3b_ctx *ctx, int64_t steps);
/* after a synchronisation: how many pulls gave up waiting for their neighbour (10 s each); non-zero = the run is invalid */
int vx3b_batch_halo_timeouts(vx3b_ctx *ctx, int64_t *n);
/* diagnostic: nanoseconds the pulls have spent waiting for their neighbours' flags since the lists were connected (load imbalance) */
int vx3b_batch_halo_wait_ns(vx3b_ctx *ctx, int64_t *ns);

/* ---------------------------------------------------------------- 3. job -------------
 * Replaces vx3_node_worker main + VX3_SimulationManager::start (vx3_node_worker.cu:32-142,
 * VX3_SimulationManager.cu:138-155): every *.vxd in input_dir over base.vxa, worlds sharded over
 * `num_devices` GPUs, report XML in the reference's layout written to report_path. */
int vx3b_job_run(const char *input_dir, const char *report_path, int num_devices, int precision);
/* The same with the base VXA and the VXD file names (relative to input_dir) given explicitly: what the worker reads from its
 * task file (vx3_node_worker.cu:64-75) and hands to VX3_SimulationManager(sub_batches, base, input_dir, nDevices).  A VXD that
 * cannot be read or parsed is reported on stderr and left out of the report; the call fails only if none is usable. */
int vx3b_job_run_files(const char *base_vxa_path, const char *input_dir, const char *const *vxd_files, int num_files, const char *report_path,
                       int num_devices, int precision);
/* The worker's whole job: read the task file <vxa>, <input_dir>, <vxd><f>...</f></vxd> the front-end wrote (Voxelyze3.cpp:101-111),
 * stream the .history of recording worlds to stdout first when stream_history != 0, then run and write the report. */
int vx3b_job_run_task(const char *vxt_path, const char *report_path, int num_devices, int precision, int stream_history);
/* The .history stream of `Voxelyze3 -i <dir>` without -o (README.md:84-100, VX3_SimulationManager.cu:40-114): for every *.vxd
 * over base.vxa whose RecordStepSize is non-zero, the settings header, then every real_stepsize steps one frame of surface-voxel
 * records (and link records if RecordLink), in the reference's text format, to out_path (NULL = stdout).  One world at a time. */
int vx3b_job_history(const char *input_dir, const char *out_path, int precision);
int vx3b_job_history_files(const char *base_vxa_path, const char *input_dir, const char *const *vxd_files, int num_files, const char *out_path,
                           int precision);
/* ---------------------------------------------------------------- 3b. a generation on one device, pipelined ----
 * What VX3_SimulationManager::start does for ONE device - readVXD, startKernel, collectResults, one after the other
 * (VX3_SimulationManager.cu:138-155, 277-458) - as a software pipeline: the worlds are dealt (by cost) to `parts` sub-batches, each with
 * its own context, stream and host thread; the sub-batches are uploaded one after the other and each starts stepping the moment it is
 * resident, so the GPU works while the rest of the generation is still being staged and copied.  results [n] come back in input order.
 * Contexts (and their device memory) are kept between calls: a generation after a generation allocates nothing. */
typedef struct vx3b_gen vx3b_gen;
int vx3b_gen_create(int device, int precision, int parts, vx3b_gen **out);
int vx3b_gen_set_option(vx3b_gen *gen, const char *name, double value); /* vx3b_ctx_set_option on every sub-batch */
int vx3b_gen_run(vx3b_gen *gen, const vx3b_model *const *models, int n, int64_t max_steps, int use_stop_formula, VX3B_WorldResult *results);
int64_t vx3b_gen_total_voxel_steps(vx3b_gen *gen); /* of the last vx3b_gen_run */
void vx3b_gen_destroy(vx3b_gen *gen);

/* Host-only: assign n worlds of the given cost to num_shards devices / ranks, longest-processing-time-first (the reference
 * deals files round-robin, vx3_node_worker.cu:88-93).  Deterministic, so every rank of a multi-process job agrees. */
int vx3b_shard_worlds(const int64_t *cost, int n, int num_shards, int32_t *shard_of);

#ifdef __cplusplus
}
#endif
#endif /* VX3B200_H */
