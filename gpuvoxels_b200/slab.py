"""One body over several GPUs: 1-D slab decomposition with halo exchange (BASELINE configs[4]).

The reference has no counterpart - a world never spans devices (VX3_SimulationManager.cu:138-155 deals whole files to GPUs).
The lattice is regular and links only join 6-neighbours, so a cut perpendicular to one axis severs links of that axis only:

  * rank r owns the voxels whose lattice coordinate along the axis falls in its range;
  * it also holds, as GHOSTS, the voxels of the two neighbouring slabs that a cut link joins to one of its own voxels, and every
    link with at least one own end.  A cut link therefore exists on both sides and is stepped on both from identical inputs
    (link history is a pure function of both ends' states), so the two copies stay identical and no force ever travels back;
  * after every time step each rank sends the new position / temperature / orientation of its boundary voxels to the neighbour
    that holds them as ghosts (28-48 bytes per voxel and face) - the only data-path exchange, neighbour to neighbour.

`native_part` asks the C++ host model for ONE slab (vx3b_model_from_vxa_slab: every rank parses the text and builds only its own
voxels, ghosts and exchange lists - what a 400^3 body needs); `split_body` is the same cut in numpy on a whole-body model and
stays as its cross-check (tests/test_slab.py).  `SlabRunner` drives one slab through the C ABI, with either of two exchanges:

  * device-side (`connect` + `run`): the neighbour's receive area is mapped into this process (CUDA IPC over NVLink; a plain
    pointer inside one process), k_halo_push stores the boundary records straight into it and raises a flag, k_halo_pull waits
    for the flag - no host, no NCCL call in the loop, 64 time steps per CUDA graph (vx3b_batch_run_slab);
  * host-driven (`step_begin` / `step_end` around a caller-supplied exchange of packed buffers: NCCL send/recv in bench.py
    --slab-exchange nccl, plain device copies in the tests) - the round-1 path, kept as the reference point.
"""
import numpy as np

from . import ArrayModel

_VOXEL_ARRAYS = ("lattice", "voxel_mat", "phase_offset", "base_cilia", "shift_cilia", "dof_fixed", "ext_force", "ext_moment",
                 "ext_translation", "ext_rotation")


class SlabPart:
    """What rank `rank` of `n_parts` needs: its sub-world and the four exchange lists (indices into the sub-world, each in
    ascending order of the body's voxel numbering, so that a sender's list and the receiver's ghost list line up)."""

    def __init__(self, rank, n_parts, model, global_index, own, send_lo, send_hi, ghost_lo, ghost_hi):
        self.rank, self.n_parts, self.model = rank, n_parts, model
        self.global_index = global_index  # sub-world voxel -> voxel of the whole body
        self.own = own                    # bool per sub-world voxel
        self.send_lo, self.send_hi, self.ghost_lo, self.ghost_hi = send_lo, send_hi, ghost_lo, ghost_hi

    @property
    def num_own(self):
        return int(self.own.sum())


def native_part(vxa_text, vxd_text, rank, n_parts, axis=2):
    """Rank `rank`'s slab straight from the VXA (+ VXD) text, built by the C++ host model; .body_dt is the whole body's time step"""
    from . import Model
    m = Model.from_vxa_slab(vxa_text, vxd_text, axis, rank, n_parts)
    si = m.slab_info()
    own = m.array("ghost") == 0
    p = SlabPart(rank, n_parts, m, si["global_index"], own, si["send_lo"], si["send_hi"], si["ghost_lo"], si["ghost_hi"])
    p.body_dt, p.body_voxels, p.body_links = si["body_dt"], si["body_voxels"], si["body_links"]
    return p


def slab_bounds(coords, n_parts):
    """Cut positions along the axis: n_parts ranges of lattice layers holding (nearly) equal numbers of voxels."""
    lo, hi = int(coords.min()), int(coords.max())
    counts = np.bincount(coords - lo, minlength=hi - lo + 1)
    cum = np.cumsum(counts)
    bounds = [lo]
    for k in range(1, n_parts):
        layer = int(np.searchsorted(cum, cum[-1] * k / n_parts, side="left")) + 1 + lo
        bounds.append(min(max(layer, bounds[-1] + 1), hi + 1))
    bounds.append(hi + 1)
    return bounds


def split_body(model, n_parts, axis=2):
    """Cuts `model` into n_parts slabs along lattice axis `axis`; returns one SlabPart per rank (every rank can compute all of
    them, or only its own: the result depends on the model alone)."""
    lat = model.array("lattice")
    nv = len(lat)
    neg, pos = model.array("link_neg"), model.array("link_pos")
    ax, lmat = model.array("link_axis"), model.array("link_mat")
    per_voxel = {}
    for name in _VOXEL_ARRAYS:
        if getattr(model.desc, name):
            per_voxel[name] = model.array(name)
    coord = lat[:, axis].astype(np.int64)
    bounds = slab_bounds(coord, n_parts)
    if len(set(bounds)) != len(bounds):
        raise ValueError(f"cannot cut {bounds[-1] - bounds[0]} lattice layers into {n_parts} slabs")
    part_of = np.searchsorted(np.array(bounds[1:]), coord, side="right")
    cut = part_of[neg] != part_of[pos]  # links severed by a cut (always links of `axis`)
    if np.any(ax[cut] != axis) or np.any(np.abs(part_of[neg[cut]] - part_of[pos[cut]]) != 1):
        raise ValueError("a link joins voxels that are not lattice neighbours across the cut")
    parts = []
    for r in range(n_parts):
        own_full = part_of == r
        ghost_full = np.zeros(nv, dtype=bool)
        ghost_full[pos[cut & (part_of[neg] == r)]] = True  # upper neighbours of my top layer
        ghost_full[neg[cut & (part_of[pos] == r)]] = True  # lower neighbours of my bottom layer
        keep_v = own_full | ghost_full
        gidx = np.nonzero(keep_v)[0]                       # creation order is preserved
        remap = np.full(nv, -1, dtype=np.int64)
        remap[gidx] = np.arange(len(gidx))
        keep_l = (own_full[neg] | own_full[pos]) & keep_v[neg] & keep_v[pos]  # at least one own end; no ghost-ghost link
        arrays = {name: a[gidx] for name, a in per_voxel.items()}
        arrays["ghost"] = ghost_full[gidx].astype(np.uint8)
        arrays["link_neg"] = remap[neg[keep_l]].astype(np.int32)
        arrays["link_pos"] = remap[pos[keep_l]].astype(np.int32)
        arrays["link_axis"] = ax[keep_l]
        arrays["link_mat"] = lmat[keep_l]
        sub_part = part_of[gidx]
        # my boundary voxels that the lower / upper neighbour holds as ghosts = own ends of the cut links towards it
        send_lo_full = np.zeros(nv, dtype=bool)
        send_hi_full = np.zeros(nv, dtype=bool)
        send_lo_full[pos[cut & (part_of[pos] == r)]] = True
        send_hi_full[neg[cut & (part_of[neg] == r)]] = True
        parts.append(SlabPart(r, n_parts, ArrayModel(model, arrays), gidx, own_full[gidx],
                              send_lo=np.nonzero(send_lo_full[gidx])[0].astype(np.int32), send_hi=np.nonzero(send_hi_full[gidx])[0].astype(np.int32),
                              ghost_lo=np.nonzero(sub_part == r - 1)[0].astype(np.int32), ghost_hi=np.nonzero(sub_part == r + 1)[0].astype(np.int32)))
    for a, b in zip(parts[:-1], parts[1:]):  # a sender's list and the receiver's ghost list name the same voxels in the same order
        assert np.array_equal(a.global_index[a.send_hi], b.global_index[b.ghost_lo])
        assert np.array_equal(b.global_index[b.send_lo], a.global_index[a.ghost_hi])
    return parts


SEND_LO, SEND_HI, GHOST_LO, GHOST_HI = 0, 1, 2, 3  # halo list ids of a slab context


class SlabRunner:
    """One slab on one device.  `exchange(send_lo, send_hi, recv_lo, recv_hi)` moves the packed boundary records (torch uint8
    tensors on this device; a buffer is None where the slab has no neighbour) - see bench.py for the NCCL version."""

    def __init__(self, part, device, precision, dt, options=None, stream=None):
        import torch
        from . import Context
        self.part, self.torch = part, torch
        # the engine, the pack / unpack kernels and the exchange must share ONE stream; torch's legacy default stream has
        # handle 0, which the C ABI reads as "use your own", so a real stream object is required
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        self.ctx = Context(device, precision)
        self.ctx.set_option("path", 5)          # every world on the brick path: the only one that knows ghosts
        self.ctx.set_option("dt_override", dt)  # the whole body's time step, not this slab's
        for k, v in (options or {}).items():
            self.ctx.set_option(k, v)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ctx.upload([part.model])
        rec = self.ctx.halo_record_bytes()
        self.buf = {}
        for lid, idx in ((SEND_LO, part.send_lo), (SEND_HI, part.send_hi), (GHOST_LO, part.ghost_lo), (GHOST_HI, part.ghost_hi)):
            self.ctx.halo_set(0, lid, idx)
            self.buf[lid] = torch.empty(len(idx) * rec, dtype=torch.uint8, device=f"cuda:{device}") if len(idx) else None
        self.halo_bytes_per_step = sum(b.numel() for b in self.buf.values() if b is not None)

    def step_begin(self):
        """one time step + packing of the two boundary layers (all enqueued on the current stream)"""
        self.ctx.run_async(1)
        for lid in (SEND_LO, SEND_HI):
            if self.buf[lid] is not None:
                self.ctx.halo_export(0, lid, self.buf[lid].data_ptr())

    def step_end(self):
        """after the exchange filled the receive buffers: refresh the ghosts"""
        for lid in (GHOST_LO, GHOST_HI):
            if self.buf[lid] is not None:
                self.ctx.halo_import(0, lid, self.buf[lid].data_ptr())

    # -- device-side exchange -----------------------------------------------------------------------------------------------
    def open_endpoints(self):
        """receive sides of my two ghost lists -> {"lo": (endpoint bytes, slots) | None, "hi": ...} to hand to the neighbours"""
        out = {}
        for name, lid, idx in (("lo", GHOST_LO, self.part.ghost_lo), ("hi", GHOST_HI, self.part.ghost_hi)):
            out[name] = (bytes(self.ctx.halo_open(0, lid)), self.ctx.halo_slots(0, lid, len(idx))) if len(idx) else None
        return out

    def connect(self, lower_hi_endpoint, upper_lo_endpoint, same_process=False, fused=True):
        """my send lists feed the lower neighbour's "hi" ghosts and the upper neighbour's "lo" ghosts; fused: the step kernel stores
        straight into the neighbour's state arrays (one launch per time step), else staged push / pull kernels"""
        for lid, idx, ep in ((SEND_LO, self.part.send_lo, lower_hi_endpoint), (SEND_HI, self.part.send_hi, upper_lo_endpoint)):
            if len(idx):
                self.ctx.halo_connect(0, lid, ep[0], same_process, ep[1] if fused else None)

    def run(self, steps):
        """`steps` time steps with the device-side exchange after each (asynchronous; every slab must be given the same count)"""
        self.ctx.run_slab(steps)

    def own_state(self):
        s = self.ctx.read_state(0)
        return {k: v[self.part.own] for k, v in s.items()}, self.part.global_index[self.part.own]

    def close(self):
        self.ctx.close()


def body_time_step(model):
    """The time step the whole body takes, which every slab must use (the stiffest link may sit in another slab): host
    arithmetic of vx3b_world_time_step on the whole body's description."""
    import ctypes as C
    from . import _check, lib
    dt = C.c_double()
    _check(lib().vx3b_world_time_step(C.byref(model.desc), C.byref(dt)), "vx3b_world_time_step")
    return dt.value
