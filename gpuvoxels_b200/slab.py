"""One body over several GPUs: 1-D slab decomposition with halo exchange (BASELINE configs[4]).

The reference has no counterpart - a world never spans devices (VX3_SimulationManager.cu:138-155 deals whole files to GPUs).
The lattice is regular and links only join 6-neighbours, so a cut perpendicular to one axis severs links of that axis only:

  * rank r owns the voxels whose lattice coordinate along the axis falls in its range;
  * it also holds, as GHOSTS, the voxels of the two neighbouring slabs that a cut link joins to one of its own voxels, and every
    link with at least one own end.  A cut link therefore exists on both sides and is stepped on both from identical inputs
    (link history is a pure function of both ends' states), so the two copies stay identical and no force ever travels back;
  * after every time step each rank sends the new position / temperature / orientation of its boundary voxels to the neighbour
    that holds them as ghosts (28-48 bytes per voxel and face) - the only data-path exchange, neighbour to neighbour.

`split_body` is pure numpy on the host model; `SlabRunner` drives one slab through the C ABI (vx3b_batch_run_async +
vx3b_batch_halo_export / _import) with the exchange done by a caller-supplied function: NCCL send/recv over torch.distributed
in bench.py, plain device copies when several slabs share one GPU (tests).
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
