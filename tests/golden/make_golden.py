"""Regenerates tests/golden/*.vxbin by running the UNMODIFIED reference CPU library (oracle/_ref/vx2_ref, built from
/root/reference by oracle/Makefile) on VXA text.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the reference's model dump (voxels, links, material constants) and its state after N steps.
`mode` vx2 = plain CVoxelyze::doTimeStep; vx3 = with the Voxelyze3 stepping deltas applied by the harness
(oracle/ref_harness.cpp --vx3: double dt, phase-offset temperature, cilia, angMom zeroing)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gpuvoxels_b200 as gv  # noqa: E402
import oracle_lib as O  # noqa: E402
from gpuvoxels_b200 import worlds as W  # noqa: E402

CASES = {
    # name: (vxa text, vxd text, steps, mode, cilia)
    "example_vx2_1000": (open(os.path.join(ROOT, "tests", "data", "example.vxa")).read(), None, 1000, "vx2", False),
    "example_vx3_1000": (open(os.path.join(ROOT, "tests", "data", "example.vxa")).read(), None, 1000, "vx3", False),
    "small_vx3_2000": (W.small_block_vxa(), None, 2000, "vx3", False),
    "xeno3_vx3_cilia_800": (W.xenobot_base_vxa(), W.xenobot_vxd(3), 800, "vx3", True),
    "cube8_vx2_400": (W.cube_vxa(8), None, 400, "vx2", False),
}


def main():
    assert O.have_reference(), "build oracle/_ref first: make -C oracle ref"
    for name, (vxa, vxd, steps, mode, cilia) in CASES.items():
        m = gv.Model.from_vxa(vxa, vxd)
        state, model, _ = O.run_reference(m.merged_vxa(), steps, vx3=(mode == "vx3"), enable_cilia=cilia, want_model=True)
        out = {}
        for k in ("scalars", "pos", "orient", "linMom", "angMom", "link_pos2", "link_angle1v", "link_angle2v", "link_flags"):
            out["state_" + k] = state[k]
        for k, v in model.items():
            out["model_" + k] = v
        O.write_vxbin(os.path.join(HERE, name + ".vxbin"), out)
        print(name, "voxels", len(state["pos"]), "links", len(state["link_pos2"]), "dt", state["scalars"][0, 1])


TWO_BODY_STEPS = (40, 60, 100, 300, 1000)


def two_body_model():
    m = gv.Model.from_vxa(W.two_body_drop_vxa())
    m.set_world_param("vx2_compat", 1)
    m.set_world_param("vary_temp", 0)
    return m


def two_body_drop():
    """The reference's OWN collision system (--vx2-collisions 1: CVoxelyze::enableCollisions, Voxelyze.cpp:612-750) on the
    two-voxel drop: pins the per-pair contact force the VX3 step shares with it (VX_Collision.cpp:42-56)."""
    m = two_body_model()
    out = {}
    for steps in TWO_BODY_STEPS:
        state, _, _ = O.run_reference(m.merged_vxa(), steps, vx2_collisions=True)
        out[f"pos_{steps}"] = state["pos"]
        out[f"linMom_{steps}"] = state["linMom"]
    O.write_vxbin(os.path.join(HERE, "two_body_drop_vx2col.vxbin"), out)
    print("two_body_drop_vx2col", {k: v.shape for k, v in out.items()})


PEBBLE_STEPS = (200, 400, 1000, 2000)


def pebbles_model():
    m = gv.Model.from_vxa(W.block_on_pebbles_vxa())
    m.set_world_param("vx2_compat", 1)
    m.set_world_param("vary_temp", 0)
    return m


def block_on_pebbles():
    """The reference's own collision system on a multi-voxel body landing on two pebbles (several watched pairs per voxel,
    friction, tipping): every contact is inter-body, where VX3 and Voxelyze2 select the same pairs."""
    m = pebbles_model()
    out = {}
    for steps in PEBBLE_STEPS:
        state, _, _ = O.run_reference(m.merged_vxa(), steps, vx2_collisions=True)
        out[f"pos_{steps}"] = state["pos"]
        out[f"orient_{steps}"] = state["orient"]
    O.write_vxbin(os.path.join(HERE, "block_on_pebbles_vx2col.vxbin"), out)
    print("block_on_pebbles_vx2col", {k: v.shape for k, v in out.items()})


PUSHER_STEPS = (200, 1000, 2000, 4000)


def pusher_model():
    m = gv.Model.from_vxa(W.xenobot_base_vxa(), W.cilia_pusher_vxd())
    m.set_world_param("enable_cilia", 1)
    return m


def cilia_pusher():
    """VX3 semantics end to end against the reference: harness --vx3 (double dt, cilia, angMom zeroing) + the reference's own
    collision system, on a cilia body pushing a target pebble 1.4 m along the floor."""
    m = pusher_model()
    out = {}
    for steps in PUSHER_STEPS:
        state, _, _ = O.run_reference(m.merged_vxa(), steps, vx3=True, enable_cilia=True, vx2_collisions=True)
        out[f"pos_{steps}"] = state["pos"]
        out[f"orient_{steps}"] = state["orient"]
    O.write_vxbin(os.path.join(HERE, "cilia_pusher_vx3_vx2col.vxbin"), out)
    print("cilia_pusher_vx3_vx2col", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "two_body":
        two_body_drop()  # (the other fixtures are left as they are)
    elif len(sys.argv) > 1 and sys.argv[1] == "pebbles":
        block_on_pebbles()
    elif len(sys.argv) > 1 and sys.argv[1] == "pusher":
        cilia_pusher()
    else:
        main()
        two_body_drop()
        block_on_pebbles()
        cilia_pusher()
