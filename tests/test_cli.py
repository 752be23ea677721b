"""Voxelyze3 / vx3_node_worker front-ends: the reference's command-line contract (Voxelyze3/src/Executables/Voxelyze3.cpp:25-131,
vx3_node_worker.cu:32-142).  Usage errors return 1; a run returns 0 and writes the report even-handedly with -o/-f."""
import os
import subprocess
import tempfile
import xml.etree.ElementTree as ET

import pytest

from gpuvoxels_b200 import worlds as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FRONT = os.path.join(ROOT, "bin", "Voxelyze3")
WORKER = os.path.join(ROOT, "bin", "vx3_node_worker")


def _have_bins():
    return os.path.exists(FRONT) and os.path.exists(WORKER)


def _generation(td, n=3, stop=0.05):
    d = os.path.join(td, "data")
    os.makedirs(d)
    open(os.path.join(d, "base.vxa"), "w").write(W.xenobot_base_vxa(stop_seconds=stop))
    for i in range(n):
        open(os.path.join(d, f"bot_{i:04d}.VXD" if i == 1 else f"bot_{i:04d}.vxd"), "w").write(W.xenobot_vxd(40 + i))  # extension is case-insensitive
    return d


@pytest.mark.skipif(not _have_bins(), reason="bin/ not built (python __graft_entry__.py)")
def test_usage_errors_return_one():
    with tempfile.TemporaryDirectory() as td:
        d = _generation(td)
        run = lambda *a: subprocess.run([FRONT, *a], cwd=td, capture_output=True, text=True)  # noqa: E731
        assert run().returncode == 1  # no -i
        r = run("-i", os.path.join(td, "missing"))
        assert r.returncode == 1 and "input directory not found" in r.stdout
        os.remove(os.path.join(d, "base.vxa"))
        r = run("-i", d)
        assert r.returncode == 1 and "No base.vxa" in r.stdout
        open(os.path.join(d, "base.vxa"), "w").write(W.xenobot_base_vxa())
        open(os.path.join(td, "report.xml"), "w").write("old")
        r = run("-i", d, "-o", "report.xml")
        assert r.returncode == 1 and "output file exists" in r.stdout  # refuses to overwrite without -f
        r = run("-i", d, "-w", os.path.join(td, "no_such_worker"))
        assert r.returncode == 1 and "worker" in r.stdout
        w = subprocess.run([WORKER], capture_output=True, text=True)
        assert w.returncode == 1


@pytest.mark.gpu
@pytest.mark.skipif(not _have_bins(), reason="bin/ not built (python __graft_entry__.py)")
def test_generation_through_the_front_end():
    with tempfile.TemporaryDirectory() as td:
        d = _generation(td)
        r = subprocess.run([FRONT, "-i", d, "-o", "report.xml", "-l", "-f"], cwd=td, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert os.path.isdir(os.path.join(td, "workspace", "locally"))
        root = ET.parse(os.path.join(td, "report.xml")).getroot()
        detail = list(root.find("detail"))
        assert sorted(x.tag for x in detail) == ["bot_0000", "bot_0001", "bot_0002"]
        fits = [float(x.find("fitness_score").text) for x in detail]
        assert fits == sorted(fits, reverse=True)
        assert root.find("bestfit/fitness_score").text == detail[0].find("fitness_score").text
        for x in detail:
            assert float(x.find("currentTime").text) > 0.05


@pytest.mark.gpu
@pytest.mark.skipif(not _have_bins(), reason="bin/ not built (python __graft_entry__.py)")
def test_worker_honours_the_task_file():
    """vx3_node_worker reads <vxa>, <input_dir> and the <vxd><f> list from the .vxt (vx3_node_worker.cu:64-75): a task that lists
    a subset simulates only that subset, over the base VXA it names; XML entities in paths are decoded (the front-end escapes
    them); a VXD that cannot be parsed is reported and left out instead of failing the generation."""
    with tempfile.TemporaryDirectory() as td:
        d = os.path.join(td, "gen & <1>")
        os.makedirs(d)
        base = os.path.join(td, "elsewhere.vxa")  # NOT <input_dir>/base.vxa
        open(base, "w").write(W.xenobot_base_vxa(stop_seconds=0.03))
        for i in range(4):
            open(os.path.join(d, f"bot_{i:04d}.vxd"), "w").write(W.xenobot_vxd(50 + i))
        open(os.path.join(d, "broken.vxd"), "w").write("<VXD><Structure replace=")
        esc = lambda s: s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;")  # noqa: E731
        vxt = os.path.join(td, "task.vxt")
        open(vxt, "w").write(f'<?xml version="1.0" encoding="utf-8"?>\n<vxa>{esc(base)}</vxa><input_dir>{esc(d)}</input_dir>'
                             "<vxd><f>bot_0001.vxd</f><f>broken.vxd</f><f>bot_0003.vxd</f></vxd>")
        vxr = os.path.join(td, "out.vxr")
        r = subprocess.run([WORKER, "-i", vxt, "-o", vxr], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "broken.vxd" in r.stderr
        root = ET.parse(vxr).getroot()
        assert root.find("inputdir").text == "gen & <1>"
        assert sorted(x.tag for x in root.find("detail")) == ["bot_0001", "bot_0003"]
        # and through the front-end: the whole directory, the broken file skipped
        open(os.path.join(d, "base.vxa"), "w").write(W.xenobot_base_vxa(stop_seconds=0.03))
        r = subprocess.run([FRONT, "-i", d, "-o", "report.xml", "-l", "-f"], cwd=td, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert sorted(x.tag for x in ET.parse(os.path.join(td, "report.xml")).getroot().find("detail")) == [f"bot_{i:04d}" for i in range(4)]
