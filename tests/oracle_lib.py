"""Test-side access to the checker: oracle/libvx3oracle.so (our CPU restatement) and
oracle/_ref/vx2_ref (the unmodified reference CPU path + harness).  TEST INFRASTRUCTURE ONLY —
nothing under gpuvoxels_b200/ imports this module."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

from gpuvoxels_b200._abi import WorldDesc, WorldResult

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libvx3oracle.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "vx2_ref")
_DP = C.POINTER(C.c_double)

_lib = None


def oracle():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        lib = C.CDLL(ORACLE_SO)
        lib.vx3o_create.argtypes = [C.POINTER(WorldDesc), C.POINTER(C.c_void_p)]
        lib.vx3o_free.argtypes = [C.c_void_p]
        lib.vx3o_free.restype = None
        lib.vx3o_run.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        lib.vx3o_run.restype = C.c_int64
        lib.vx3o_read_state.argtypes = [C.c_void_p, _DP, _DP, _DP, _DP]
        lib.vx3o_read_links.argtypes = [C.c_void_p, _DP, _DP, _DP, C.POINTER(C.c_int32)]
        lib.vx3o_result.argtypes = [C.c_void_p, C.POINTER(WorldResult)]
        lib.vx3o_collision_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.vx3o_set_dt.argtypes = [C.c_void_p, C.c_double]
        lib.vx3o_num_surface.argtypes = [C.c_void_p]
        lib.vx3o_set_force.argtypes = [C.c_void_p, C.c_int, _DP, _DP]
        lib.vx3o_set_global_damping.argtypes = [C.c_void_p, C.c_int, C.c_double]
        lib.vx3o_read_signals.argtypes = [C.c_void_p, _DP, C.POINTER(C.c_int32)]
        lib.vx3o_read_topology.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(_DP)


class OracleWorld:
    def __init__(self, model):
        self.model = model  # keeps the host arrays alive
        self.nv, self.nl = model.num_voxels, model.num_links
        self._h = C.c_void_p()
        rc = oracle().vx3o_create(C.byref(model.desc), C.byref(self._h))
        assert rc == 0

    def set_dt(self, dt):
        oracle().vx3o_set_dt(self._h, float(dt))

    def set_force(self, voxel, force=None, moment=None):
        f = None if force is None else np.ascontiguousarray(force, dtype=np.float64)
        m = None if moment is None else np.ascontiguousarray(moment, dtype=np.float64)
        assert oracle().vx3o_set_force(self._h, int(voxel), None if f is None else _dp(f), None if m is None else _dp(m)) == 0

    def set_global_damping(self, mat, zeta):
        assert oracle().vx3o_set_global_damping(self._h, int(mat), float(zeta)) == 0

    def run(self, steps, use_stop_formula=False):
        return oracle().vx3o_run(self._h, int(steps), 1 if use_stop_formula else 0)

    def state(self):
        pos, orient = np.empty((self.nv, 3)), np.empty((self.nv, 4))
        lin, ang = np.empty((self.nv, 3)), np.empty((self.nv, 3))
        oracle().vx3o_read_state(self._h, _dp(pos), _dp(orient), _dp(lin), _dp(ang))
        return {"pos": pos, "orient": orient, "linMom": lin, "angMom": ang}

    def links(self):
        p2, a1, a2 = np.empty((self.nl, 3)), np.empty((self.nl, 3)), np.empty((self.nl, 3))
        fl = np.empty((self.nl,), dtype=np.int32)
        oracle().vx3o_read_links(self._h, _dp(p2), _dp(a1), _dp(a2), fl.ctypes.data_as(C.POINTER(C.c_int32)))
        return {"pos2": p2, "angle1v": a1, "angle2v": a2, "flags": fl}

    def signals(self):
        """(signals [nv][6]: localSignal, localSignaldt, inactiveUntil, d_signal.value, d_signal.activeTime, packmakerNextPulse;
        flags [nv][2]: removed, enableAttach)"""
        sig, fl = np.empty((self.nv, 6)), np.empty((self.nv, 2), dtype=np.int32)
        oracle().vx3o_read_signals(self._h, _dp(sig), fl.ctypes.data_as(C.POINTER(C.c_int32)))
        return sig, fl

    def topology(self):
        """current links [nl][6]: neg, pos, axis, removed, detached, isNewLink (attachment appends links)"""
        n = C.c_int32(0)
        oracle().vx3o_read_topology(self._h, None, C.byref(n))
        buf = np.empty((max(n.value, 1), 6), dtype=np.int32)
        n2 = C.c_int32(n.value)
        oracle().vx3o_read_topology(self._h, buf.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n2))
        return buf[: n.value].copy()

    def time(self):
        return self.result().current_time

    def result(self):
        r = WorldResult()
        oracle().vx3o_result(self._h, C.byref(r))
        return r

    def collision_pairs(self):
        n = C.c_int32(0)
        oracle().vx3o_collision_pairs(self._h, None, C.byref(n))
        buf = np.empty((max(n.value, 1), 2), dtype=np.int32)
        n2 = C.c_int32(n.value)
        oracle().vx3o_collision_pairs(self._h, buf.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n2))
        return buf[: n.value].copy()

    def num_surface(self):
        return oracle().vx3o_num_surface(self._h)

    def __del__(self):
        try:
            if self._h:
                oracle().vx3o_free(self._h)
                self._h = None
        except Exception:
            pass


# ---- vxbin container written by oracle/ref_harness.cpp -------------------------------------
def read_vxbin(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    i = 0
    while i < len(data):
        j = data.index(b"\n", i)
        name, dtype, rows, cols = data[i:j].decode().split()
        rows, cols = int(rows), int(cols)
        dt = {"f64": np.float64, "i32": np.int32}[dtype]
        n = rows * cols * np.dtype(dt).itemsize
        out[name] = np.frombuffer(data[j + 1: j + 1 + n], dtype=dt).reshape(rows, cols).copy()
        i = j + 1 + n
    return out


def write_vxbin(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.ndim == 1:
                a = a.reshape(-1, 1)
            dtype = {"float64": "f64", "int32": "i32"}[str(a.dtype)]
            f.write(f"{name} {dtype} {a.shape[0]} {a.shape[1]}\n".encode())
            f.write(a.tobytes())


def have_reference():
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(vxa_text, steps, vx3=False, enable_cilia=False, enable_expansion=False, no_temp=False, dt=None, threads=1,
                  want_model=False, timing=False, repeat=1, vx2_collisions=False):
    """Run the reference CPU path on `vxa_text`; returns (state dict, model dict or None, timing dict or None)."""
    with tempfile.TemporaryDirectory() as td:
        vxa = os.path.join(td, "in.vxa")
        with open(vxa, "w") as f:
            f.write(vxa_text)
        cmd = [REF_BIN, "--vxa", vxa, "--steps", str(steps), "--dump", os.path.join(td, "s.vxbin"), "--threads", str(threads),
               "--repeat", str(repeat)]
        if vx3:
            cmd.append("--vx3")
        if enable_cilia:
            cmd += ["--enable-cilia", "1"]
        if enable_expansion:
            cmd += ["--enable-expansion", "1"]
        if no_temp:
            cmd.append("--no-temp")
        if vx2_collisions:
            cmd += ["--vx2-collisions", "1"]
        if dt is not None:
            cmd += ["--dt", repr(float(dt))]
        if want_model:
            cmd += ["--dump-model", os.path.join(td, "m.vxbin")]
        if timing:
            cmd.append("--time")
        out = subprocess.run(cmd, check=True, capture_output=True, text=True)
        state = read_vxbin(os.path.join(td, "s.vxbin"))
        model = read_vxbin(os.path.join(td, "m.vxbin")) if want_model else None
        tim = json.loads(out.stdout.strip().splitlines()[-1]) if timing else None
    return state, model, tim


REF3_BIN = os.path.join(ROOT, "oracle", "_ref", "vx3_ref")


def have_reference_vx3():
    return os.path.exists(REF3_BIN) and os.access(REF3_BIN, os.X_OK)


def run_reference_vx3(merged_vxa_text, steps, use_stop=False, dump_at=(), timing=False, repeat=1):
    """Run the reference's own Voxelyze3 step code (host build, oracle/_ref/vx3_ref: serial thread schedule) on a MERGED VXA text
    (base VXA with the VXD applied).  Returns (dict of arrays; checkpoints carry the suffix _N, the final state none), timing."""
    with tempfile.TemporaryDirectory() as td:
        vxa = os.path.join(td, "in.vxa")
        with open(vxa, "w") as f:
            f.write(merged_vxa_text)
        cmd = [REF3_BIN, "--vxa", vxa, "--steps", str(steps), "--dump", os.path.join(td, "s.vxbin"), "--repeat", str(repeat)]
        if use_stop:
            cmd.append("--use-stop")
        if dump_at:
            cmd += ["--dump-at", ",".join(str(int(x)) for x in dump_at)]
        if timing:
            cmd.append("--time")
        out = subprocess.run(cmd, check=True, capture_output=True, text=True)
        state = read_vxbin(os.path.join(td, "s.vxbin"))
        tim = json.loads(out.stdout.strip().splitlines()[-1]) if timing else None
    return state, tim


def rel_err(a, b):
    """max |a-b| / max(|b|) over the array — the 'relative' of the north-star tolerances."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)
