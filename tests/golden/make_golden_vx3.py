"""Regenerates tests/golden/vx3ref_*.vxbin by running the reference's OWN Voxelyze3 step code, built for the host
(oracle/_ref/vx3_ref: the VX3 .cu sources compiled where they lie under /root/reference against oracle/vx3_shim, serial thread
schedule; oracle/Makefile target ref3), on merged VXA text.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_vx3.py [case ...]

Unlike make_golden.py (Voxelyze2's CVoxelyze + stepping deltas applied by the harness) these are the reference's VX3 semantics
end to end: its all-pairs contact with the depth-1 neighbour exclusion INSIDE a body, target hits, cilia, force field, the
formula machine, the centre-of-mass sampling and the fitness."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gpuvoxels_b200 as gv  # noqa: E402
import oracle_lib as O  # noqa: E402
from gpuvoxels_b200 import worlds as W  # noqa: E402

REF_INPUTS = os.path.join(ROOT, "tests", "data", "ref_inputs")


def _read(*parts):
    with open(os.path.join(REF_INPUTS, *parts)) as f:
        return f.read()


CIRCLE_PUSH_BOTS = [f"bot_{i:04d}" for i in range(11)]

# name: (vxa text, vxd text or None, steps, checkpoints (state also kept at these step counts), run to the stop condition?)
CASES = {
    "vx3ref_xeno7_contact_500": (W.xenobot_base_vxa(), W.xenobot_vxd(7), 500, (100, 300), False),
    "vx3ref_xeno15_contact_stop1s": (W.xenobot_base_vxa(stop_seconds=1.0), W.xenobot_vxd(15), 1000000, (300,), True),
    "vx3ref_pile_k3c4_600": (W.contact_pile_vxa(k=3, cube=4), None, 600, (260,), False),
    "vx3ref_forcefield_800": (W.force_field_vxa(), None, 800, (1, 100), False),
    "vx3ref_example_1000": (open(os.path.join(ROOT, "tests", "data", "example.vxa")).read(), None, 1000, (), False),
    # bilinear + data-model materials bent past their yield points and swinging back; the same with Poisson's ratio 0.3
    "vx3ref_plastic_beam_6000": (W.plastic_beam_vxa(), None, 6000, (1000, 3000), False),
    "vx3ref_plastic_beam_poisson_6000": (W.plastic_beam_vxa(volume_effects=True, nu=0.3), None, 6000, (1000, 3000), False),
    # signals, attachment, detachment, removal: the reference's own c.signaling / h.target_only inputs and two synthetic worlds
    "vx3ref_signaling_data1_3000": (_read("c.signaling_data1", "base.vxa"), _read("c.signaling_data1", "rob.vxd"), 3000, (1000, 2000), False),
    "vx3ref_signaling_data2_3000": (_read("c.signaling_data2", "base.vxa"), _read("c.signaling_data2", "rob.vxd"), 3000, (1000, 2000), False),
    "vx3ref_target_only_300": (_read("h.target_only_data1", "base.vxa"), _read("h.target_only_data1", "bot_0000.vxd"), 300, (), False),
    "vx3ref_sticky_pebbles_3000": (W.sticky_pebbles_vxa(), None, 3000, (200, 600, 1500), False),
    "vx3ref_breaking_chain_3000": (W.breaking_chain_vxa(), None, 3000, (300, 1000), False),
    "vx3ref_memleaking_bot0000_1000": (_read("d.memleaking_data1", "base.vxa"), _read("d.memleaking_data1", "bot_0000.vxd"), 1000, (300,), False),
}
for _b in CIRCLE_PUSH_BOTS:
    # the reference's own e.circle_push inputs: 1,000 steps for state parity, and the whole run (stop at t = 5 s) for the report
    CASES[f"vx3ref_circle_push_{_b}_1000"] = (_read("e.circle_push_data901", "base.vxa"), _read("e.circle_push_data901", _b + ".vxd"), 1000, (300,), False)
    CASES[f"vx3ref_circle_push_{_b}_full"] = (_read("e.circle_push_data901", "base.vxa"), _read("e.circle_push_data901", _b + ".vxd"), 1000000, (), True)

LINK_STATE_CASES = ("vx3ref_xeno7_contact_500", "vx3ref_example_1000", "vx3ref_forcefield_800", "vx3ref_plastic_beam_6000", "vx3ref_plastic_beam_poisson_6000")  # link history kept for these only (size)
KEEP = ("scalars", "pos", "orient", "linMom", "angMom", "pairs")
DYNAMIC_CASES = ("vx3ref_signaling_data1_3000", "vx3ref_signaling_data2_3000", "vx3ref_target_only_300", "vx3ref_sticky_pebbles_3000", "vx3ref_breaking_chain_3000")
KEEP_DYNAMIC = ("signals", "links", "vflags", "surface")  # signal state per voxel, the (grown) link list, removed / enableAttach flags
SCALARS = ("currentTime", "dt", "OptimalDt", "steps", "collisionCount", "fitness_score", "com_x", "com_y", "com_z", "initial_com_x", "initial_com_y",
           "initial_com_z", "recentAngle", "targetCloseness", "numClosePairs", "angleSampleTimes", "diverged", "CurStepCount")


def merged(vxa, vxd):
    return gv.Model.from_vxa(vxa, vxd).merged_vxa() if vxd else vxa


def generate(name):
    vxa, vxd, steps, checkpoints, use_stop = CASES[name]
    state, _ = O.run_reference_vx3(merged(vxa, vxd), steps, use_stop=use_stop, dump_at=checkpoints)
    out = {}
    full = name.endswith("_full")
    for k, v in state.items():
        base = k.split("_")[0]
        if full:  # whole runs are chaotic: only the world-level results are a fixture
            if k == "scalars":
                out[k] = v
        elif base in KEEP or (k.startswith("link_state") and name in LINK_STATE_CASES) or (name in DYNAMIC_CASES and base in KEEP_DYNAMIC):
            if name == "vx3ref_target_only_300" and base in ("linMom", "angMom", "pairs", "links"):
                continue  # 2,964 voxels: keep the fixture small (positions, orientations, signals, flags)
            out[k] = v
    O.write_vxbin(os.path.join(HERE, name + ".vxbin"), out)
    sc = state["scalars"][0]
    return f"{name}: voxels {len(state['pos'])} steps {int(sc[3])} hits {int(sc[4])} fitness {sc[5]:.6g} pairs {len(state['pairs'])}"


CUBE100_STEPS = 20


def cube100_sample_indices(n=100):
    """voxel indices (x fastest, then y, then z: the reference's creation order) of a corner column, the centre column and the
    diagonal of the bottom layer"""
    idx = lambda x, y, z: (z * n + y) * n + x  # noqa: E731
    return np.array([idx(0, 0, z) for z in range(n)] + [idx(n // 2, n // 2, z) for z in range(n)] + [idx(i, i, 0) for i in range(n)], dtype=np.int32)


def cube100():
    """BASELINE configs[2] at full size: the 100^3 body, 20 steps of the reference's VX3 code; a voxel subsample is the fixture."""
    state, _ = O.run_reference_vx3(W.cube_vxa(100), CUBE100_STEPS)
    pick = cube100_sample_indices()
    out = {"scalars": state["scalars"], "sample": pick.reshape(-1, 1)}
    for k in ("pos", "orient", "linMom", "angMom"):
        out[k] = state[k][pick]
    lat = np.stack(np.meshgrid(np.arange(100), np.arange(100), np.arange(100), indexing="ij"), -1)  # [z][y][x] -> (z, y, x)
    nominal = lat[..., ::-1].reshape(-1, 3) * 0.01
    disp = state["pos"] - nominal
    out["disp_max_abs"] = np.abs(disp).max(axis=0).reshape(1, 3)
    out["disp_z_layer_mean"] = disp[:, 2].reshape(100, -1).mean(axis=1).reshape(-1, 1)  # mean vertical displacement per z layer
    O.write_vxbin(os.path.join(HERE, "vx3ref_cube100_20.vxbin"), out)
    return f"vx3ref_cube100_20: voxels {len(state['pos'])} max |disp| {out['disp_max_abs']}"


def main(argv):
    assert O.have_reference_vx3(), "build oracle/_ref/vx3_ref first: make -C oracle ref3"
    names = argv or (list(CASES) + ["cube100"])
    with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as pool:  # each case is its own vx3_ref process
        for line in pool.map(lambda n: cube100() if n == "cube100" else generate(n), names):
            print(line, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
