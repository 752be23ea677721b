"""Synthetic VXA / VXD text generators for the BASELINE.json configurations.

Everything is emitted as real VXA/VXD XML so the same text feeds our host model and the reference's own
parser (oracle/_ref).  Fixed seeds; numpy only.
"""
import numpy as np

_MAT_DEFAULTS = dict(name="mat", E=5e6, density=1e6, nu=0.35, cte=0.0, u_static=1.0, u_dynamic=0.5, cilia=0.0,
                     is_target=0, fixed=0, is_measured=1, mat_model=0, fail_stress=0.0, plastic_mod=0.0, yield_stress=0.0, ss_data=None)


def material(**kw):
    m = dict(_MAT_DEFAULTS)
    m.update(kw)
    return m


def _mat_xml(i, m):
    extra = ""
    if m["cilia"]:
        extra += f"<Cilia>{m['cilia']:g}</Cilia>"
    if m["is_target"]:
        extra += "<isTarget>1</isTarget>"
    if m["fixed"]:
        extra += "<Fixed>1</Fixed>"
    for tag in ("Sticky", "isPaceMaker", "PaceMakerPeriod", "signalValueDecay", "signalTimeDelay", "inactivePeriod", "RemoveFromSimulationAfterThisManySeconds",
                "TurnOnThermalExpansionAfterThisManySeconds", "TurnOnCiliaAfterThisManySeconds"):
        if tag in m and m[tag]:
            extra += f"<{tag}>{m[tag]!r}</{tag}>"
    if not m["is_measured"]:
        extra += "<isMeasured>0</isMeasured>"
    ss = ""
    if m["ss_data"]:  # MatModel 3: stress/strain points (CVXC_Material::ReadXML, VXA/VX_Object.cpp:1395-1417)
        pts = list(m["ss_data"])
        ss = (f"<SSData><NumDataPts>{len(pts)}</NumDataPts><StrainData>" + "".join(f"<Strain>{float(a)!r}</Strain>" for a, _ in pts) +
              "</StrainData><StressData>" + "".join(f"<Stress>{float(b)!r}</Stress>" for _, b in pts) + "</StressData></SSData>")
    return (f'<Material ID="{i}"><MatType>0</MatType><Name>{m["name"]}</Name>'
            f"<Display><Red>0.5</Red><Green>0.5</Green><Blue>0.5</Blue><Alpha>1</Alpha></Display>"
            f"<Mechanical>{extra}<MatModel>{m['mat_model']}</MatModel>{ss}<Elastic_Mod>{m['E']!r}</Elastic_Mod>"
            f"<Plastic_Mod>{m['plastic_mod']!r}</Plastic_Mod><Yield_Stress>{m['yield_stress']!r}</Yield_Stress><FailModel>0</FailModel>"
            f"<Fail_Stress>{m['fail_stress']!r}</Fail_Stress><Fail_Strain>0</Fail_Strain><Density>{m['density']!r}</Density>"
            f"<Poissons_Ratio>{m['nu']!r}</Poissons_Ratio><CTE>{m['cte']!r}</CTE><MaterialTempPhase>0</MaterialTempPhase>"
            f"<uStatic>{m['u_static']!r}</uStatic><uDynamic>{m['u_dynamic']!r}</uDynamic></Mechanical></Material>")


def structure_xml(data, phase=None, base_cilia=None, replace=None):
    """data: int array [Z][Y][X] of palette ids (0 = empty).  phase: float [Z][Y][X].  base_cilia: float [Z][Y][X][3]."""
    data = np.asarray(data)
    Z, Y, X = data.shape
    assert data.max() <= 74, "ASCII_READABLE stores id+48 in one char"
    attr = f' replace="{replace}"' if replace else ""
    out = [f'<Structure{attr} Compression="ASCII_READABLE"><X_Voxels>{X}</X_Voxels><Y_Voxels>{Y}</Y_Voxels><Z_Voxels>{Z}</Z_Voxels><Data>']
    for z in range(Z):
        out.append("<Layer><![CDATA[" + "".join(chr(48 + int(c)) for c in data[z].reshape(-1)) + "]]></Layer>")
    out.append("</Data>")
    if phase is not None:
        out.append("<PhaseOffset>")
        for z in range(Z):
            out.append("<Layer><![CDATA[" + ", ".join(repr(float(p)) for p in np.asarray(phase)[z].reshape(-1)) + "]]></Layer>")
        out.append("</PhaseOffset>")
    if base_cilia is not None:
        out.append("<BaseCiliaForce>")
        for z in range(Z):
            out.append("<Layer><![CDATA[" + ", ".join(repr(float(p)) for p in np.asarray(base_cilia)[z].reshape(-1)) + "]]></Layer>")
        out.append("</BaseCiliaForce>")
    out.append("</Structure>")
    return "".join(out)


def formula_xml(tag, body):
    return f"<{tag}>{body}</{tag}>"


STOP_AT = lambda seconds: f"<mtSUB><mtVAR>t</mtVAR><mtCONST>{seconds!r}</mtCONST></mtSUB>"  # noqa: E731
FITNESS_XY_DISTANCE = ("<mtSQRT><mtADD><mtMUL><mtVAR>x</mtVAR><mtVAR>x</mtVAR></mtMUL>"
                       "<mtMUL><mtVAR>y</mtVAR><mtVAR>y</mtVAR></mtMUL></mtADD></mtSQRT>")


def vxa_document(materials, structure, lattice_dim=0.01, dt_frac=0.9, bond_damping=1.0, col_damping=0.8, slow_damping=0.01,
                 grav_acc=-9.81, floor=True, temp_enabled=True, temp_amplitude=0.0, temp_base=25.0, vary_temp=False,
                 temp_period=0.1, enable_collision=None, enable_cilia=False, enable_expansion=False, stop_formula=None,
                 fitness_formula=None, force_field=None, save_positions=False, volume_effects=False, enable_signals=False, enable_attach=False,
                 enable_detach=False, attach_conditions=None, safety_guard=None, secondary_experiment=False, reinitialize_after=None):
    sim = []
    if enable_signals:
        sim.append("<EnableSignals>1</EnableSignals>")
    if secondary_experiment:
        sim.append("<SecondaryExperiment>1</SecondaryExperiment>")
    if reinitialize_after is not None:
        sim.append(f"<ReinitializeInitialPositionAfterThisManySeconds>{reinitialize_after!r}</ReinitializeInitialPositionAfterThisManySeconds>")
    if enable_cilia:
        sim.append("<EnableCilia>1</EnableCilia>")
    if enable_expansion:
        sim.append("<EnableExpansion>1</EnableExpansion>")
    if save_positions:
        sim.append("<SavePositionOfAllVoxels>1</SavePositionOfAllVoxels>")
    if enable_collision is not None or enable_attach or enable_detach:
        ad = f"<EnableCollision>{int(enable_collision if enable_collision is not None else 1)}</EnableCollision>"
        ad += f"<EnableAttach>{int(enable_attach)}</EnableAttach><EnableDetach>{int(enable_detach)}</EnableDetach>"
        if safety_guard is not None:
            ad += f"<SafetyGuard>{int(safety_guard)}</SafetyGuard>"
        if attach_conditions:
            ad += "<AttachCondition>" + "".join(f"<Condition_{i}>{c}</Condition_{i}>" for i, c in enumerate(attach_conditions)) + "</AttachCondition>"
        sim.append(f"<AttachDetach>{ad}</AttachDetach>")
    if fitness_formula:
        sim.append(formula_xml("FitnessFunction", fitness_formula))
    if force_field:
        sim.append("<ForceField>" + "".join(formula_xml(f"{ax}_forcefield", f) for ax, f in zip("xyz", force_field) if f) + "</ForceField>")
    sim.append(f"<Integration><Integrator>0</Integrator><DtFrac>{dt_frac!r}</DtFrac></Integration>")
    sim.append(f"<Damping><BondDampingZ>{bond_damping!r}</BondDampingZ><ColDampingZ>{col_damping!r}</ColDampingZ>"
               f"<SlowDampingZ>{slow_damping!r}</SlowDampingZ></Damping>")
    sim.append("<Collisions><SelfColEnabled>1</SelfColEnabled><ColSystem>3</ColSystem><CollisionHorizon>3</CollisionHorizon></Collisions>")
    sim.append(f"<Features><VolumeEffectsEnabled>{int(volume_effects)}</VolumeEffectsEnabled></Features>")
    stop = formula_xml("StopConditionFormula", stop_formula) if stop_formula else ""
    sim.append(f"<StopCondition><StopConditionType>0</StopConditionType><StopConditionValue>99</StopConditionValue>{stop}</StopCondition>")
    env = (f"<Environment><Boundary_Conditions><NumBCs>0</NumBCs></Boundary_Conditions>"
           f"<Gravity><GravEnabled>1</GravEnabled><GravAcc>{grav_acc!r}</GravAcc><FloorEnabled>{int(floor)}</FloorEnabled></Gravity>"
           f"<Thermal><TempEnabled>{int(temp_enabled)}</TempEnabled><TempAmplitude>{temp_amplitude!r}</TempAmplitude>"
           f"<TempBase>{temp_base!r}</TempBase><VaryTempEnabled>{int(vary_temp)}</VaryTempEnabled><TempPeriod>{temp_period!r}</TempPeriod>"
           f"</Thermal></Environment>")
    palette = "".join(_mat_xml(i + 1, m) for i, m in enumerate(materials))
    vxc = (f'<VXC Version="0.94"><Lattice><Lattice_Dim>{lattice_dim!r}</Lattice_Dim></Lattice>'
           f"<Voxel><Vox_Name>BOX</Vox_Name></Voxel><Palette>{palette}</Palette>{structure}</VXC>")
    return ('<?xml version="1.0" encoding="ISO-8859-1"?>\n<VXA Version="1.1"><Simulator>' + "".join(sim) + "</Simulator>" + env + vxc + "</VXA>\n")


# ---------------------------------------------------------------------------------------------------------------
def small_block_vxa(enable_collision=False, seed=3):
    """4x4x3 block, two materials (one thermally actuated with phase offsets), floor + gravity.  The smoke-test world."""
    rng = np.random.default_rng(seed)
    data = np.ones((3, 4, 4), dtype=int)
    data[:, 1:3, 1:3] = 2
    data[0, 0, 0] = 0
    data[2, 3, :2] = 0
    phase = rng.uniform(0, 1, size=data.shape)
    mats = [material(name="soft", E=5e6, cte=0.0), material(name="active", E=1e7, cte=0.01)]
    return vxa_document(mats, structure_xml(data, phase=phase), dt_frac=0.5, temp_amplitude=14.4714, vary_temp=True, temp_period=0.2,
                        enable_collision=enable_collision)


def cube_vxa(n, steps_hint=None, enable_collision=False):
    """C3: n^3 single-material soft body under gravity on the floor (doc/misc/example.vxa's damping), no self-collision."""
    data = np.ones((n, n, n), dtype=int)
    return vxa_document([material(name="body", E=5e6, density=1e6)], structure_xml(data), dt_frac=0.9, bond_damping=1.0, col_damping=0.8,
                        slow_damping=0.01, grav_acc=-9.81, floor=True, enable_collision=enable_collision)


XENOBOT_DIMS = (7, 14, 14)  # Z, Y, X — e.circle_push uses 12x12x7; two more rows/columns keep the pebbles clear of the body


def xenobot_base_vxa(stop_seconds=5.0, enable_collision=True):
    """C2 base.vxa modelled on Experiments/xenobot/e.circle_push/base.vxa: cilia-driven body (material 1, Cilia=8) and
    target pebbles (material 2), GravAcc=-1, DtFrac=1, TempPeriod 0.1, stop at t = stop_seconds."""
    mats = [material(name="Cilia", E=5e6, density=5e5, cilia=8.0), material(name="TargetPebble", E=5e6, density=5e5, is_target=1)]
    fitness = ("<mtADD><mtADD><mtABS><mtVAR>angle</mtVAR></mtABS><mtLOG><mtADD><mtVAR>hit</mtVAR><mtCONST>1</mtCONST></mtADD></mtLOG></mtADD>"
               "<mtLOG><mtADD><mtMUL><mtVAR>x</mtVAR><mtVAR>x</mtVAR></mtMUL><mtMUL><mtVAR>y</mtVAR><mtVAR>y</mtVAR></mtMUL></mtADD></mtLOG></mtADD>")
    placeholder = structure_xml(np.array([[[1, 1], [1, 1]], [[0, 0], [0, 0]]]))
    return vxa_document(mats, placeholder, dt_frac=1.0, grav_acc=-1.0, floor=True, temp_amplitude=7.72173, vary_temp=True, temp_period=0.1,
                        enable_collision=enable_collision, enable_cilia=True, stop_formula=STOP_AT(stop_seconds), fitness_formula=fitness)


def _largest_component(mask):
    Z, Y, X = mask.shape
    seen = np.zeros_like(mask, dtype=bool)
    best = None
    for start in zip(*np.nonzero(mask)):
        if seen[start]:
            continue
        comp, stack = [], [start]
        seen[start] = True
        while stack:
            z, y, x = stack.pop()
            comp.append((z, y, x))
            for dz, dy, dx in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):
                q = (z + dz, y + dy, x + dx)
                if 0 <= q[0] < Z and 0 <= q[1] < Y and 0 <= q[2] < X and mask[q] and not seen[q]:
                    seen[q] = True
                    stack.append(q)
        if best is None or len(comp) > len(best):
            best = comp
    out = np.zeros_like(mask, dtype=bool)
    for c in best or []:
        out[c] = True
    return out


def xenobot_vxd(index, pebbles=True):
    """C2 robot: largest 6-connected component of rand < p on the 12x12x7 body region of the lattice, p ~ U(0.35, 0.75),
    numpy default_rng(1000 + index); BaseCiliaForce ~ U(-1,1)^3 (z component 0, as evolved bots push in-plane); optionally
    two target pebbles in the far corner as in data1/bot_0017.vxd."""
    rng = np.random.default_rng(1000 + index)
    Z, Y, X = XENOBOT_DIMS
    p = rng.uniform(0.35, 0.75)
    mask = np.zeros((Z, Y, X), dtype=bool)
    mask[:, :12, :12] = rng.uniform(size=(Z, 12, 12)) < p
    body = _largest_component(mask)
    # drop the body onto the floor: shift down so its lowest layer is z = 0
    zs = np.nonzero(body.any(axis=(1, 2)))[0]
    body = np.roll(body, -int(zs.min()), axis=0) if len(zs) else body
    data = body.astype(int)
    if pebbles:
        data[0, Y - 1, X - 2:] = 2
        data[1, Y - 1, X - 2:] = 2
    cilia = np.zeros((Z, Y, X, 3))
    cilia[..., :2] = rng.uniform(-1, 1, size=(Z, Y, X, 2))
    cilia[data != 1] = 0
    return "<VXD>" + structure_xml(data, base_cilia=cilia, replace="VXA.VXC.Structure") + "</VXD>"


def xenobot_cost(index):
    """Estimated cost (voxels + 2 x links) of the body xenobot_vxd(index) builds, from its fill probability alone (the first draw of
    the world's generator): what a driver needs to deal a generation to devices by cost before any world is built.  (A job that has
    the models in hand uses their real counts: csrc/job.cpp.)"""
    p = np.random.default_rng(1000 + index).uniform(0.35, 0.75)
    cells = XENOBOT_DIMS[0] * 12 * 12
    return int(cells * p + 2 * 3 * cells * p * p)


def two_body_drop_vxa():
    """One voxel resting on the floor and a second one dropped onto it from two cells up: the smallest scene with a watched
    contact pair (the reference's TEST(CVoxelyze, collisions), VoxCAD/src/VX2/test/tVoxelyze.h:1167-1197, as a VXA)."""
    data = np.zeros((3, 1, 1), dtype=int)
    data[0, 0, 0], data[2, 0, 0] = 1, 2
    mats = [material(name="lower", E=1e6, density=1e6), material(name="upper", E=1e6, density=1e6)]
    return vxa_document(mats, structure_xml(data), dt_frac=0.9, grav_acc=-9.81, floor=True, enable_collision=True)


def block_on_pebbles_vxa():
    """A stiff 3x3x2 block dropped from two cells up onto two single-voxel pebbles lying on the floor: it lands on them, tips,
    pushes them aside and ends on the floor.  Every contact is between different bodies, so VX3's all-pairs scheme and
    Voxelyze2's watch lists (which differ only in which pairs INSIDE a body they exclude) watch the same pairs."""
    data = np.zeros((5, 3, 3), dtype=int)
    data[0, 0, 0] = 2
    data[0, 2, 1] = 2
    data[2:4, :, :] = 1
    mats = [material(name="block", E=3e7, density=1e6), material(name="pebble", E=3e7, density=1e6)]
    return vxa_document(mats, structure_xml(data), dt_frac=0.9, grav_acc=-9.81, floor=True, enable_collision=True)


def cilia_pusher_vxd():
    """VXD for xenobot_base_vxa(): a solid 3x3x2 cilia body (BaseCiliaForce +x) with one target pebble two cells ahead of it.
    The body walks into the pebble and pushes it along the floor: cilia, contact with a target (hit counter), floor friction -
    the ingredients of a xenobot run - while every contact stays inter-body (a solid block has no unlinked pair near 1.25 cells)."""
    Z, Y, X = 3, 5, 9
    data = np.zeros((Z, Y, X), dtype=int)
    data[0:2, 1:4, 0:3] = 1
    data[0, 2, 5] = 2
    cilia = np.zeros((Z, Y, X, 3))
    cilia[..., 0] = 1.0
    cilia[data != 1] = 0
    return "<VXD>" + structure_xml(data, base_cilia=cilia, replace="VXA.VXC.Structure") + "</VXD>"


def contact_pile_vxa(k=3, cube=4, seed=7):
    """C4 synthetic: k x k grid of cube^3 bodies dropped from staggered heights so they pile up and collide."""
    rng = np.random.default_rng(seed)
    gap = cube + 1
    heights = rng.integers(0, 3, size=(k, k))
    Z = cube * 2 + 3 + cube
    data = np.zeros((Z, k * gap, k * gap), dtype=int)
    for a in range(k):
        for b in range(k):
            z0 = int(heights[a, b])
            data[z0:z0 + cube, a * gap:a * gap + cube, b * gap:b * gap + cube] = 1
    # a second layer of bodies above, offset by half a cube so they land on the gaps
    for a in range(k - 1):
        for b in range(k - 1):
            z0 = cube + 3
            data[z0:z0 + cube, a * gap + cube // 2 + 1:a * gap + cube // 2 + 1 + cube, b * gap + cube // 2 + 1:b * gap + cube // 2 + 1 + cube] = 2
    mats = [material(name="lower", E=5e6, density=1e6), material(name="upper", E=1e7, density=1e6)]
    return vxa_document(mats, structure_xml(data), dt_frac=0.9, grav_acc=-9.81, floor=True, enable_collision=True)


# ---- formula XML helpers (Utils/VX3_MathTree.h operators) --------------------------------------------------------------
def mt_const(v):
    return f"<mtCONST>{float(v)!r}</mtCONST>"


def mt_var(name):
    return f"<mtVAR>{name}</mtVAR>"


def mt(op, *args):
    """mt("ADD", a, b) -> <mtADD>a b</mtADD>; mt("E") / mt("PI") are leaves"""
    return f"<mt{op}>" + "".join(args) + f"</mt{op}>"


def force_field_formulas():
    """Three force-field formulas (N per voxel) that between them use EVERY operator of the reference's formula machine
    (Utils/VX3_MathTree.h:50-193) on position- and time-dependent arguments."""
    x, y, z, t = mt_var("x"), mt_var("y"), mt_var("z"), mt_var("t")
    fx = mt("MUL", mt("MUL", mt("POW", mt("ADD", mt("ABS", mt("SUB", x, mt_const(0.01))), mt_const(1)), mt_const(0.5)), mt("GREATERTHAN", z, mt_const(0.004))),
            mt_const(14.0))
    fy = mt("SUB", mt("MUL", mt("NORMALCDF", mt("SUB", mt("MUL", t, mt_const(20)), mt_const(1))), mt_const(9.0)),
            mt("MUL", mt("LESSTHAN", y, mt_const(0.01)), mt_const(2.0)))
    wave = mt("ADD", mt("SIN", mt("MUL", t, mt_const(50))), mt("COS", mt("DIV", x, mt_const(0.01))))
    logic = mt("OR", mt("AND", mt("GREATERTHAN", x, mt_const(0.005)), mt("NOT", mt("LESSTHAN", y, mt_const(0.005)))), mt("INT", mt("MUL", t, mt_const(3))))
    misc = mt("ADD", mt("ATAN", mt("TAN", mt("MUL", z, mt_const(10)))), mt("MUL", mt("SQRT", mt("ADD", mt("LOG", mt("ADD", mt("E"), t)), mt("PI"))), mt_const(0.1)))
    fz = mt("ADD", mt("MUL", mt_const(0.2), wave), mt("ADD", mt("MUL", mt_const(0.4), logic), misc))
    return fx, fy, fz


def force_field_vxa(enable_collision=False):
    """A 3x3x2 two-material block under gravity on the floor, thermally actuated, inside a force field (VX3_Voxel.cu:179-184)."""
    rng = np.random.default_rng(11)
    data = np.ones((2, 3, 3), dtype=int)
    data[:, 1, 1] = 2
    phase = rng.uniform(0, 1, size=data.shape)
    mats = [material(name="soft", E=5e6, cte=0.0), material(name="active", E=1e7, cte=0.01)]
    return vxa_document(mats, structure_xml(data, phase=phase), dt_frac=0.5, temp_amplitude=14.4714, vary_temp=True, temp_period=0.2,
                        enable_collision=enable_collision, force_field=force_field_formulas(),
                        fitness_formula=mt("ADD", mt("MUL", mt_var("x"), mt_const(100)), mt("POW", mt_const(2), mt_var("z"))))


def merge_vxd(vxa_text, vxd_text):
    """VXD applied to the base VXA the way Utils/ctool.h:49-57 (ptree_merge / put_child) does it: every child of <VXD> carrying a
    `replace="VXA.a.b.c"` attribute replaces the first node on that path (missing levels are created).  Pure Python (xml.etree):
    the reference arm of bench.py prepares its input with this, without loading the CUDA library."""
    import xml.etree.ElementTree as ET
    if not vxd_text:
        return vxa_text
    root = ET.fromstring(vxa_text.encode("latin-1", "replace") if isinstance(vxa_text, str) else vxa_text)
    vxd = ET.fromstring(vxd_text.encode("latin-1", "replace") if isinstance(vxd_text, str) else vxd_text)
    for child in list(vxd):
        path = child.attrib.get("replace", "")
        if not path:
            continue
        parts = path.split(".")
        assert parts[0] == root.tag, path
        node = root
        for name in parts[1:-1]:
            nxt = node.find(name)
            if nxt is None:
                nxt = ET.SubElement(node, name)
            node = nxt
        new = ET.fromstring(ET.tostring(child))
        new.tag = parts[-1]
        new.tail = None
        old = node.find(parts[-1])
        if old is None:
            node.append(new)
        else:
            i = list(node).index(old)
            node.remove(old)
            node.insert(i, new)
    return '<?xml version="1.0" encoding="ISO-8859-1"?>\n' + ET.tostring(root, encoding="unicode")



def plastic_beam_vxa(volume_effects=False, nu=0.0, DENS=2e4):
    """A cantilever of three materials - bilinear (yields at strain 0.004), data model (three segments) and a linear tip weight -
    held by a block of `Fixed` material and bent by gravity hard enough to take the bilinear and data links to strains of 0.05,
    past their yield points but short of failure; it then swings back along E (the strain-offset path of VX3_Link.cu:232-268).
    volume_effects + nu: Poisson's ratio as Voxelyze3 applies it (eHat, frozen transverse terms)."""
    data = np.zeros((4, 2, 6), dtype=int)  # [z][y][x]
    data[3, :, 0] = 4          # the fixed anchor block (top layer, x = 0)
    data[3, :, 1:3] = 1        # bilinear
    data[3, :, 3:5] = 2        # data model
    data[3, :, 5] = 3          # linear tip weight
    data[2, :, 5] = 3
    mats = [material(name="bilinear", E=1e6, density=DENS, nu=nu, mat_model=2, plastic_mod=2.5e5, yield_stress=4e3, fail_stress=5e4),
            material(name="data", E=1e6, density=DENS, nu=nu, mat_model=3, ss_data=[(0.004, 4e3), (0.02, 1.2e4), (0.08, 2.0e4)]),
            material(name="tip", E=5e6, density=2 * DENS, nu=nu),
            material(name="anchor", E=5e7, density=1e6, nu=nu, fixed=1)]
    return vxa_document(mats, structure_xml(data), lattice_dim=0.01, dt_frac=0.5, bond_damping=0.3, slow_damping=0.002, grav_acc=-9.81, floor=False,
                        enable_collision=False, volume_effects=volume_effects)



def sticky_pebbles_vxa(safety_guard=50):
    """Attachment (VX3_VoxelyzeKernel.cu:718-831): a 2x2x1 raft of sticky voxels resting on the floor and two sticky pebbles of the same
    material dropped onto it from two cells up, one of them off-centre.  On contact they grow links to the raft (and, once they lean
    together, possibly to each other); a new link eases in over SafetyGuard steps.  A pace-maker voxel in the raft fires signals that
    travel along the links, old and new, and modulate the cilia push of the pebbles (ShiftCiliaForce)."""
    data = np.zeros((4, 3, 3), dtype=int)
    data[0, 0:2, 0:2] = 1
    data[0, 0, 0] = 2      # pace maker (different material: its links to the raft use the 1-2 link material)
    data[3, 0, 0] = 1      # pebble A above the pace maker
    data[3, 1, 1] = 1      # pebble B above the opposite corner
    mats = [material(name="sticky", E=2e6, density=1e6, Sticky=1, cilia=0.2, signalValueDecay=0.8, signalTimeDelay=0.004, inactivePeriod=0.01),
            material(name="pacer", E=2e6, density=1e6, isPaceMaker=1, PaceMakerPeriod=0.05, signalValueDecay=0.8, signalTimeDelay=0.004, inactivePeriod=0.01)]
    Z, Y, X = data.shape
    cilia = np.zeros((Z, Y, X, 3))
    cilia[3, 0, 0] = (0.3, 0.1, 0.0)
    shift = np.zeros((Z, Y, X, 3))
    shift[..., 0] = 0.002
    shift[data == 0] = 0
    st = structure_xml(data, base_cilia=cilia).replace("</Structure>", "<ShiftCiliaForce>" + "".join(
        "<Layer><![CDATA[" + ", ".join(repr(float(p)) for p in shift[z].reshape(-1)) + "]]></Layer>" for z in range(Z)) + "</ShiftCiliaForce></Structure>")
    return vxa_document(mats, st, dt_frac=0.5, grav_acc=-9.81, floor=True, enable_collision=True, enable_cilia=True, enable_signals=True, enable_attach=True,
                        attach_conditions=[mt_const(1)], safety_guard=safety_guard)


def breaking_chain_vxa():
    """Detachment (VX3_VoxelyzeKernel.cu:946-967) and removal (:365-399): a vertical chain hanging from a `Fixed` voxel whose middle
    material fails at 2 % strain under the weight below it - the failed link lets go, the lower part drops to the floor - and whose
    bottom material is taken out of the simulation after 0.03 s (SecondaryExperiment); the initial position is re-initialised at 0.02 s."""
    data = np.zeros((8, 1, 1), dtype=int)
    data[7, 0, 0] = 4
    data[5:7, 0, 0] = 1
    data[3:5, 0, 0] = 2
    data[1:3, 0, 0] = 3
    mats = [material(name="strong", E=1e7, density=1e6), material(name="weak", E=2e5, density=1e6, mat_model=1, fail_stress=3.5e3),
            material(name="vanishing", E=1e7, density=1.5e6, RemoveFromSimulationAfterThisManySeconds=0.03), material(name="anchor", E=1e7, density=1e6, fixed=1)]
    return vxa_document(mats, structure_xml(data), dt_frac=0.5, bond_damping=0.5, grav_acc=-9.81, floor=True, enable_collision=True, enable_detach=True,
                        secondary_experiment=True, reinitialize_after=0.02, fitness_formula=mt_var("z"))
