"""gpuvoxels_b200 — B200-native Voxelyze step engine behind the gpuVoxels interfaces.

Python here is only a binding over the C ABI in include/vx3b200.h (libvx3b200.so, built from
gpuvoxels_b200/csrc by __graft_entry__.build()).  Names mirror the reference:

  Model               ~ CVX_Sim::ReadVXA/Import + CVoxelyze builder calls   (Voxelyze3/src/VXA/VX_Sim.cpp, old/Voxelyze.cpp)
  Context             ~ VX3_VoxelyzeKernel[] on one device + CUDA_Simulation (VX3/VX3_SimulationManager.cu:11-121, 384-458)
  SimulationManager   ~ VX3_SimulationManager                              (VX3/VX3_SimulationManager.cuh:15-46)
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import FP32, FP64, WorldDesc, WorldResult  # noqa: F401

__all__ = ["Model", "ArrayModel", "Context", "SimulationManager", "FP32", "FP64", "VX3Error", "lib"]


class VX3Error(RuntimeError):
    pass


def lib():
    return _abi.load()


def _check(rc, what):
    if rc < 0:
        raise VX3Error(f"{what}: {lib().vx3b_last_error().decode(errors='replace')} (status {rc})")
    return rc


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Model:
    """Host model of one world.  Build from VXA(+VXD) text or with CVoxelyze-style calls."""

    def __init__(self, handle):
        self._h = handle
        self._desc = None

    # -- construction ---------------------------------------------------------------------
    @classmethod
    def from_vxa(cls, vxa_text, vxd_text=None):
        h = C.c_void_p()
        a = vxa_text.encode() if isinstance(vxa_text, str) else vxa_text
        d = None if vxd_text is None else (vxd_text.encode() if isinstance(vxd_text, str) else vxd_text)
        _check(lib().vx3b_model_from_vxa(a, d, C.byref(h)), "vx3b_model_from_vxa")
        return cls(h)

    @classmethod
    def from_files(cls, vxa_path, vxd_path=None):
        h = C.c_void_p()
        _check(lib().vx3b_model_from_files(os.fsencode(vxa_path), None if vxd_path is None else os.fsencode(vxd_path), C.byref(h)),
               "vx3b_model_from_files")
        return cls(h)

    @classmethod
    def from_vxa_slab(cls, vxa_text, vxd_text=None, axis=2, part=0, n_parts=1):
        """ONE SLAB of the body (vx3b_model_from_vxa_slab): own voxels of range `part` of `n_parts` along `axis` + ghosts"""
        h = C.c_void_p()
        _check(lib().vx3b_model_from_vxa_slab(vxa_text.encode() if isinstance(vxa_text, str) else vxa_text,
                                              None if vxd_text is None else vxd_text.encode() if isinstance(vxd_text, str) else vxd_text,
                                              int(axis), int(part), int(n_parts), C.byref(h)), "vx3b_model_from_vxa_slab")
        return cls(h)

    def slab_info(self):
        """exchange lists and whole-body figures of a slab model, as numpy arrays (copies)"""
        si = _abi.SlabInfo()
        _check(lib().vx3b_model_slab_info(self._h, C.byref(si)), "vx3b_model_slab_info")
        out = {k: getattr(si, k) for k in ("part", "n_parts", "axis", "layer_lo", "layer_hi", "num_own", "body_voxels", "body_links", "body_dt")}
        for name in ("send_lo", "send_hi", "ghost_lo", "ghost_hi"):
            n = getattr(si, "n_" + name)
            out[name] = np.ctypeslib.as_array(getattr(si, name), shape=(n,)).copy() if n else np.zeros(0, dtype=np.int32)
        nv = self.num_voxels
        out["global_index"] = np.ctypeslib.as_array(si.global_index, shape=(nv,)).copy() if nv else np.zeros(0, dtype=np.int64)
        return out

    @classmethod
    def new(cls, vox_size):
        h = C.c_void_p()
        _check(lib().vx3b_model_new(float(vox_size), C.byref(h)), "vx3b_model_new")
        return cls(h)

    def add_material(self, youngs_modulus, density, **params):
        self._desc = None
        m = _check(lib().vx3b_model_add_material(self._h, float(youngs_modulus), float(density)), "vx3b_model_add_material")
        for k, v in params.items():
            self.set_material_param(m, k, v)
        return m

    def set_material_param(self, mat, name, value):
        self._desc = None
        _check(lib().vx3b_model_set_material_param(self._h, mat, name.encode(), float(value)), "vx3b_model_set_material_param")

    def set_material_bilinear(self, mat, youngs_modulus, plastic_modulus, yield_stress, failure_stress=-1.0):
        """CVX_Material::setModelBilinear"""
        self._desc = None
        _check(lib().vx3b_model_set_material_bilinear(self._h, mat, float(youngs_modulus), float(plastic_modulus), float(yield_stress), float(failure_stress)),
               "vx3b_model_set_material_bilinear")

    def set_material_data(self, mat, strain, stress):
        """CVX_Material::setModel: stress/strain points, ascending, from or after (0, 0)"""
        self._desc = None
        a, b = np.ascontiguousarray(strain, dtype=np.float64), np.ascontiguousarray(stress, dtype=np.float64)
        _check(lib().vx3b_model_set_material_data(self._h, mat, len(a), _dp(a), _dp(b)), "vx3b_model_set_material_data")

    def set_voxel(self, mat, ix, iy, iz):
        self._desc = None
        return _check(lib().vx3b_model_set_voxel(self._h, mat, ix, iy, iz), "vx3b_model_set_voxel")

    def set_fixed(self, ix, iy, iz, dof_mask=_abi.DOF_ALL, translation=None, rotation=None):
        self._desc = None
        t = None if translation is None else np.ascontiguousarray(translation, dtype=np.float64)
        r = None if rotation is None else np.ascontiguousarray(rotation, dtype=np.float64)
        _check(lib().vx3b_model_set_fixed(self._h, ix, iy, iz, dof_mask, _dp(t), _dp(r)), "vx3b_model_set_fixed")

    def set_force(self, ix, iy, iz, force=None, moment=None):
        self._desc = None
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
        m = None if moment is None else np.ascontiguousarray(moment, dtype=np.float64)
        _check(lib().vx3b_model_set_force(self._h, ix, iy, iz, _dp(f), _dp(m)), "vx3b_model_set_force")

    def set_world_param(self, name, value):
        self._desc = None
        _check(lib().vx3b_model_set_world_param(self._h, name.encode(), float(value)), "vx3b_model_set_world_param")

    def finalize(self):
        _check(lib().vx3b_model_finalize(self._h), "vx3b_model_finalize")
        self._desc = None
        return self

    # -- access ---------------------------------------------------------------------------
    @property
    def desc(self):
        if self._desc is None:
            d = WorldDesc()
            rc = lib().vx3b_model_desc(self._h, C.byref(d))
            if rc == _abi_err_state():
                self.finalize()
                rc = lib().vx3b_model_desc(self._h, C.byref(d))
            _check(rc, "vx3b_model_desc")
            self._desc = d
        return self._desc

    @property
    def num_voxels(self):
        return self.desc.num_voxels

    @property
    def num_links(self):
        return self.desc.num_links

    def array(self, name):
        """numpy copy of one of the flat arrays of the world description."""
        d = self.desc
        nv, nl = d.num_voxels, d.num_links
        shapes = {
            "lattice": (nv, 3), "voxel_mat": (nv,), "phase_offset": (nv,), "base_cilia": (nv, 3), "shift_cilia": (nv, 3),
            "dof_fixed": (nv,), "ext_force": (nv, 3), "ext_moment": (nv, 3), "ext_translation": (nv, 3), "ext_rotation": (nv, 3),
            "link_neg": (nl,), "link_pos": (nl,), "link_axis": (nl,), "link_mat": (nl,), "ghost": (nv,),
        }
        ptr = getattr(d, name)
        shape = shapes[name]
        n = int(np.prod(shape))
        if n == 0 or not ptr:
            return np.zeros(shape, dtype=np.ctypeslib.as_array(ptr, (1,)).dtype if ptr else np.float64)
        return np.ctypeslib.as_array(ptr, (n,)).reshape(shape).copy()

    def voxel_materials(self):
        d = self.desc
        return [d.voxel_mats[i] for i in range(d.num_voxel_mats)]

    def link_materials(self):
        d = self.desc
        return [d.link_mats[i] for i in range(d.num_link_mats)]

    def merged_vxa(self):
        need = C.c_int64()
        lib().vx3b_model_merged_vxa(self._h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        lib().vx3b_model_merged_vxa(self._h, buf, need.value, None)
        return buf.value.decode()

    def __del__(self):
        try:
            if self._h:
                lib().vx3b_model_free(self._h)
                self._h = None
        except Exception:
            pass


def _abi_err_state():
    return -5


class ArrayModel:
    """A world description assembled from numpy arrays (a slab of a bigger Model, gpuvoxels_b200.slab): same `.desc`,
    `.num_voxels`, `.num_links`, `.array()` surface as Model.  Material tables and world scalars are shared with `parent`."""

    _PTR = {"lattice": C.c_int16, "voxel_mat": C.c_int32, "phase_offset": C.c_double, "base_cilia": C.c_double, "shift_cilia": C.c_double,
            "dof_fixed": C.c_uint8, "ext_force": C.c_float, "ext_moment": C.c_float, "ext_translation": C.c_double, "ext_rotation": C.c_double,
            "link_neg": C.c_int32, "link_pos": C.c_int32, "link_axis": C.c_uint8, "link_mat": C.c_int32, "ghost": C.c_uint8}

    def __init__(self, parent, arrays):
        self.parent = parent
        d = WorldDesc()
        C.memmove(C.byref(d), C.byref(parent.desc), C.sizeof(WorldDesc))  # scalars, formulas, material tables
        self._arrays = {}
        for name, ctype in self._PTR.items():
            a = arrays.get(name)
            if a is None:
                setattr(d, name, C.POINTER(ctype)())
                continue
            a = np.ascontiguousarray(a, dtype=np.dtype(ctype))
            self._arrays[name] = a
            setattr(d, name, a.ctypes.data_as(C.POINTER(ctype)))
        d.num_voxels = len(self._arrays["voxel_mat"])
        d.num_links = len(self._arrays["link_neg"])
        self._desc = d

    @property
    def desc(self):
        return self._desc

    @property
    def num_voxels(self):
        return self._desc.num_voxels

    @property
    def num_links(self):
        return self._desc.num_links

    def array(self, name):
        return self._arrays[name].copy()


class Context:
    """One device context: a batch of worlds resident in HBM, stepped by the CUDA engine."""

    def __init__(self, device=0, precision=FP32):
        self._h = C.c_void_p()
        self._n = 0
        self._models = []
        _check(lib().vx3b_ctx_create(int(device), int(precision), C.byref(self._h)), "vx3b_ctx_create")
        self.precision = precision

    def set_option(self, name, value):
        _check(lib().vx3b_ctx_set_option(self._h, name.encode(), float(value)), "vx3b_ctx_set_option")

    def set_stream(self, stream):
        _check(lib().vx3b_ctx_set_stream(self._h, int(stream)), "vx3b_ctx_set_stream")

    def upload(self, models):
        models = list(models)
        if all(isinstance(m, Model) for m in models):  # model handles straight through: no marshalling of n descriptors here
            for m in models:
                m.desc  # (finalizes a model that was not yet)
            handles = (C.c_void_p * len(models))(*[m._h for m in models])
            _check(lib().vx3b_batch_upload_models(self._h, handles, len(models)), "vx3b_batch_upload_models")
        else:  # views of models (slabs): descriptors assembled on this side
            arr = (WorldDesc * len(models))(*[m.desc for m in models])
            _check(lib().vx3b_batch_upload(self._h, arr, len(models)), "vx3b_batch_upload")
        self._n = len(models)
        self._models = models  # keep host arrays alive (the library copies, but descs point into them)
        self._nv = [m.num_voxels for m in models]
        self._nl = [m.num_links for m in models]

    def run(self, max_steps, use_stop_formula=False):
        _check(lib().vx3b_batch_run(self._h, int(max_steps), 1 if use_stop_formula else 0), "vx3b_batch_run")

    # -- one body over several GPUs (include/vx3b200.h section 2b) ---------------------------------
    def run_async(self, steps):
        _check(lib().vx3b_batch_run_async(self._h, int(steps)), "vx3b_batch_run_async")

    def halo_set(self, world, list_id, voxels):
        v = np.ascontiguousarray(voxels, dtype=np.int32)
        _check(lib().vx3b_batch_halo_set(self._h, int(world), int(list_id), v.ctypes.data_as(C.POINTER(C.c_int32)), len(v)), "vx3b_batch_halo_set")

    def halo_record_bytes(self):
        return _check(lib().vx3b_batch_halo_record_bytes(self._h), "vx3b_batch_halo_record_bytes")

    def halo_export(self, world, list_id, device_ptr):
        _check(lib().vx3b_batch_halo_export(self._h, int(world), int(list_id), C.c_void_p(int(device_ptr))), "vx3b_batch_halo_export")

    def halo_import(self, world, list_id, device_ptr):
        _check(lib().vx3b_batch_halo_import(self._h, int(world), int(list_id), C.c_void_p(int(device_ptr))), "vx3b_batch_halo_import")

    def halo_open(self, world, list_id):
        """receive area of a ghost list (device-side exchange); returns the endpoint to hand to the neighbour"""
        ep = _abi.HaloEndpoint()
        _check(lib().vx3b_batch_halo_open(self._h, world, list_id, C.byref(ep)), "vx3b_batch_halo_open")
        return ep

    def halo_connect(self, world, list_id, endpoint, same_process=False, peer_slots=None):
        """peer_slots (the neighbour's halo_slots of its ghost list): the exchange fused into the step kernel; None: staged push / pull kernels"""
        if isinstance(endpoint, (bytes, bytearray)):
            endpoint = _abi.HaloEndpoint.from_buffer_copy(endpoint)
        if peer_slots is None:
            _check(lib().vx3b_batch_halo_connect(self._h, world, list_id, C.byref(endpoint), 1 if same_process else 0), "vx3b_batch_halo_connect")
        else:
            ps = np.ascontiguousarray(peer_slots, dtype=np.int32)
            _check(lib().vx3b_batch_halo_connect_fused(self._h, world, list_id, C.byref(endpoint), ps.ctypes.data_as(C.POINTER(C.c_int32)), 1 if same_process else 0),
                   "vx3b_batch_halo_connect_fused")

    def halo_slots(self, world, list_id, n):
        out = np.empty(n, dtype=np.int32)
        _check(lib().vx3b_batch_halo_slots(self._h, world, list_id, out.ctypes.data_as(C.POINTER(C.c_int32))), "vx3b_batch_halo_slots")
        return out

    def run_slab(self, steps):
        _check(lib().vx3b_batch_run_slab(self._h, int(steps)), "vx3b_batch_run_slab")

    def halo_timeouts(self):
        n = C.c_int64()
        _check(lib().vx3b_batch_halo_timeouts(self._h, C.byref(n)), "vx3b_batch_halo_timeouts")
        return n.value

    def halo_wait_ns(self):
        n = C.c_int64()
        _check(lib().vx3b_batch_halo_wait_ns(self._h, C.byref(n)), "vx3b_batch_halo_wait_ns")
        return n.value

    def results(self):
        out = (WorldResult * self._n)()
        _check(lib().vx3b_batch_results(self._h, out), "vx3b_batch_results")
        return list(out)

    def read_state(self, world=0):
        nv = self._nv[world]
        pos, orient = np.empty((nv, 3)), np.empty((nv, 4))
        lin, ang = np.empty((nv, 3)), np.empty((nv, 3))
        _check(lib().vx3b_batch_read_state(self._h, world, _dp(pos), _dp(orient), _dp(lin), _dp(ang)), "vx3b_batch_read_state")
        return {"pos": pos, "orient": orient, "linMom": lin, "angMom": ang}

    def read_initial_pos(self, world=0):
        pos = np.empty((self._nv[world], 3))
        _check(lib().vx3b_batch_read_initial_pos(self._h, world, _dp(pos)), "vx3b_batch_read_initial_pos")
        return pos

    def num_links(self, world=0):
        """links in use (a world with attachment grows new ones)"""
        n = C.c_int32()
        _check(lib().vx3b_batch_num_links(self._h, world, C.byref(n)), "vx3b_batch_num_links")
        return n.value

    def read_signals(self, world=0):
        """(signals [nv][6], flags [nv][2] = removed, enableAttach): include/vx3b200.h vx3b_batch_read_signals"""
        nv = self._nv[world]
        sig, fl = np.empty((nv, 6)), np.empty((nv, 2), dtype=np.int32)
        _check(lib().vx3b_batch_read_signals(self._h, world, _dp(sig), fl.ctypes.data_as(C.POINTER(C.c_int32))), "vx3b_batch_read_signals")
        return sig, fl

    def read_topology(self, world=0):
        """current link list [nl][6] = neg, pos, axis, removed, detached, isNewLink"""
        n = C.c_int32(self.num_links(world))
        out = np.empty((n.value, 6), dtype=np.int32)
        _check(lib().vx3b_batch_read_topology(self._h, world, out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)), "vx3b_batch_read_topology")
        return out[: n.value]

    def read_links(self, world=0):
        nl = self.num_links(world)
        p2, a1, a2 = np.empty((nl, 3)), np.empty((nl, 3)), np.empty((nl, 3))
        fl = np.empty((nl,), dtype=np.int32)
        _check(lib().vx3b_batch_read_links(self._h, world, _dp(p2), _dp(a1), _dp(a2), fl.ctypes.data_as(C.POINTER(C.c_int32))),
               "vx3b_batch_read_links")
        return {"pos2": p2, "angle1v": a1, "angle2v": a2, "flags": fl}

    def set_force(self, world, voxel, force=None, moment=None):
        """CVX_External::setForce / setMoment between runs; needs a boundary-condition slot (option ext_all=1 before upload)."""
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
        m = None if moment is None else np.ascontiguousarray(moment, dtype=np.float64)
        _check(lib().vx3b_batch_set_force(self._h, world, int(voxel), None if f is None else _dp(f), None if m is None else _dp(m)),
               "vx3b_batch_set_force")

    def set_global_damping(self, world, mat, zeta):
        """CVX_Material::setGlobalDamping of voxel material `mat` of world `world` between runs."""
        _check(lib().vx3b_batch_set_global_damping(self._h, int(world), int(mat), float(zeta)), "vx3b_batch_set_global_damping")

    def collision_pairs(self, world=0, capacity=1 << 20):
        buf = np.empty((capacity, 2), dtype=np.int32)
        n = C.c_int32(capacity)
        _check(lib().vx3b_batch_collision_pairs(self._h, world, buf.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)),
               "vx3b_batch_collision_pairs")
        if n.value > capacity:
            return self.collision_pairs(world, n.value)
        return buf[: n.value].copy()

    def timing(self):
        ms, k = C.c_float(), C.c_int64()
        _check(lib().vx3b_ctx_timing(self._h, C.byref(ms), C.byref(k)), "vx3b_ctx_timing")
        return ms.value, k.value

    def total_voxel_steps(self):
        return lib().vx3b_ctx_total_voxel_steps(self._h)

    def close(self):
        if self._h:
            lib().vx3b_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Generation:
    """A generation of worlds on one device, pipelined over `parts` sub-batches (vx3b_gen_*): sub-batch k + 1 is staged and copied
    while sub-batch k already steps.  run() returns the results in the order of `models`."""

    def __init__(self, device=0, precision=FP32, parts=4):
        self._h = C.c_void_p()
        _check(lib().vx3b_gen_create(int(device), int(precision), int(parts), C.byref(self._h)), "vx3b_gen_create")

    def set_option(self, name, value):
        _check(lib().vx3b_gen_set_option(self._h, name.encode(), float(value)), "vx3b_gen_set_option")

    def run(self, models, max_steps, use_stop_formula=False):
        for m in models:
            m.desc  # (finalizes a model that was not yet)
        handles = (C.c_void_p * len(models))(*[m._h for m in models])
        out = (WorldResult * len(models))()
        _check(lib().vx3b_gen_run(self._h, handles, len(models), int(max_steps), 1 if use_stop_formula else 0, out), "vx3b_gen_run")
        return list(out)

    def total_voxel_steps(self):
        return lib().vx3b_gen_total_voxel_steps(self._h)

    def close(self):
        if self._h:
            lib().vx3b_gen_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SimulationManager:
    """Mirror of VX3_SimulationManager(sub_batches, base, input_dir, num_devices): start() is blocking and
    leaves h_results sorted by fitness descending, NaN last (VX3_SimulationResult.h:26-33)."""

    def __init__(self, sub_batches, base, input_dir, num_of_devices, precision=FP32):
        self.sub_batches = [list(b) for b in sub_batches]
        self.base = base
        self.input_dir = input_dir
        self.num_of_devices = num_of_devices
        self.precision = precision
        self.h_results = []

    def start(self):
        import math
        import threading
        per_device = [[] for _ in self.sub_batches]
        errors = []

        def run_device(device_index, files):  # one host thread + one context per device, like csrc/job.cpp
            try:
                models = [Model.from_files(self.base, os.path.join(self.input_dir, f)) for f in files]
                ctx = Context(device_index, self.precision)
                ctx.upload(models)
                ctx.run(1000000, use_stop_formula=True)
                for k, (f, m, r) in enumerate(zip(files, models, ctx.results())):
                    pos = ctx.read_state(k)["pos"] if m.desc.save_position_of_all_voxels else None
                    per_device[device_index].append({"vxa_filename": os.path.basename(f), "result": r, "model": m, "pos": pos})
                ctx.close()
            except Exception as e:  # noqa: BLE001 - re-raised on the caller's thread below
                errors.append(e)

        threads = [threading.Thread(target=run_device, args=(d, files)) for d, files in enumerate(self.sub_batches) if files]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        self.h_results = [e for dev in per_device for e in dev]
        self.h_results.sort(key=lambda e: (math.isnan(e["result"].fitness_score), -e["result"].fitness_score
                                           if not math.isnan(e["result"].fitness_score) else 0.0))
        return self.h_results
