"""Known-answer scenarios replayed from the reference's own Voxelyze2 unit tests
(/root/reference/VoxCAD/src/VX2/test/tVoxelyze.h), expressed with our CVoxelyze-style builder.
Each scenario returns (model, steps, check) where check(state_dict) asserts the reference's expected value.
The same scenarios run on the CPU oracle (-m "not gpu") and through the C ABI on the GPU (-m gpu)."""
import numpy as np

import gpuvoxels_b200 as gv

TX, TY, TZ, RX, RY, RZ = 1, 2, 4, 8, 16, 32
OFF = {"X_POS": (1, 0, 0), "X_NEG": (-1, 0, 0), "Y_POS": (0, 1, 0), "Y_NEG": (0, -1, 0), "Z_POS": (0, 0, 1), "Z_NEG": (0, 0, -1)}


def rotvec(q, robust=False):
    """Quat3D::ToRotationVector (VoxCAD/src/VX2/include/Quat3D.h).  robust=True: the same rotation vector formed from the
    vector part (2*atan2(|v|, w)), for states that were stored in float, where w rounds to exactly 1 for tiny angles."""
    w, v = q[0], np.array(q[1:])
    if robust:
        n = np.linalg.norm(v)
        return v * (2.0 * np.arctan2(n, w) / n) if n > 0 else np.zeros(3)
    if w >= 1.0 or w <= -1.0:
        return np.zeros(3)
    sl = 1.0 - w * w
    if sl < 2.4e-3:
        return v * 2.0 * np.sqrt((2 - 2 * w) / sl)
    return v * 2.0 * np.arccos(w) / np.sqrt(sl)


def _base(sz=0.001, E=1e6, rho=1e3, internal=1.0, glob=0.2):
    m = gv.Model.new(sz)
    mat = m.add_material(E, rho, internal_damping=internal, global_damping=glob)
    m.set_world_param("dt_frac", 1.0)  # tests step with ts = recommendedTimeStep()
    return m, mat


def two_vox(direction, force, moment, free_dof, expected, what, steps=1000):
    """test2Vox (tVoxelyze.h:66-113): voxel 1 fixed, voxel 2 loaded; all dofs of voxel 2 fixed except free_dof."""
    m, mat = _base()
    off = OFF[direction]
    m.set_voxel(mat, 0, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_voxel(mat, *off)
    m.set_fixed(*off, 63 & ~free_dof)
    m.set_force(*off, force=force, moment=moment)
    m.finalize()
    idx = 1

    def check(state, tol=1e-5):
        if what in (TX, TY, TZ):
            k = {TX: 0, TY: 1, TZ: 2}[what]
            data = state["pos"][idx][k] - off[k] * 0.001
        else:
            k = {RX: 0, RY: 1, RZ: 2}[what]
            data = rotvec(state["orient"][idx], robust=tol > 1e-5)[k]
        if expected == 0:
            assert abs(data) < 1e-12, data
        else:
            assert abs(data - expected) < abs(expected) * max(tol, 1e-5), (data, expected)
    return m, steps, check


def single_bond_cases():
    """subset of TEST(CVoxelyze, singleBondFixedFree), tVoxelyze.h:142-199"""
    F, M = 1e-3, 1e-9
    c = []
    for d in ("X_POS", "X_NEG"):
        c.append((f"{d}-FPX", two_vox(d, (F, 0, 0), None, TX, 1e-6, TX)))
        c.append((f"{d}-FPY", two_vox(d, (0, F, 0), None, TY, 1e-6, TY)))
    c.append(("X_POS-FNZ", two_vox("X_POS", (0, 0, -F), None, TZ, -1e-6, TZ)))
    c.append(("Y_POS-FPX", two_vox("Y_POS", (F, 0, 0), None, TX, 1e-6, TX)))
    c.append(("Z_NEG-FPZ", two_vox("Z_NEG", (0, 0, F), None, TZ, 1e-6, TZ)))
    c.append(("X_POS-MPX", two_vox("X_POS", None, (M, 0, 0), RX, 1.2e-5, RX)))
    c.append(("X_POS-MNY", two_vox("X_POS", None, (0, -M, 0), RY, -3e-6, RY)))
    c.append(("Y_POS-MPY", two_vox("Y_POS", None, (0, M, 0), RY, 1.2e-5, RY)))
    c.append(("Z_POS-MPX", two_vox("Z_POS", None, (M, 0, 0), RX, 3e-6, RX)))
    c.append(("X_POS-MPY-TZ", two_vox("X_POS", None, (0, M, 0), TZ | RY, -6e-9, TZ)))
    c.append(("X_POS-FPY-RZ", two_vox("X_POS", (0, F, 0), None, TY | RZ, 6e-3, RZ)))
    return c


def internal_damping():
    """TEST(CVoxelyze, internalDamping), tVoxelyze.h:261-285: expects 1e-9/6 (EXPECT_FLOAT_EQ)"""
    m, mat = _base(internal=1.0, glob=0.0)
    m.set_voxel(mat, 0, 0, 0)
    m.set_force(0, 0, 0, force=(1e-6, 1e-6, 1e-6))
    for off in OFF.values():
        m.set_voxel(mat, *off)
        m.set_fixed(*off, 63)
    m.finalize()

    def check(state, tol=4e-7):
        assert abs(state["pos"][0][1] - 1e-9 / 6) < max(tol, 4e-7) * 1e-9 / 6, state["pos"][0]
    return m, 100, check


def global_damping_cantilever():
    """second half of TEST(CVoxelyze, globalDamping), tVoxelyze.h:316-334: expects 4e-9 (EXPECT_FLOAT_EQ)"""
    m, mat = _base(internal=0.0, glob=0.25)
    m.set_voxel(mat, 0, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_voxel(mat, 1, 0, 0)
    m.set_force(1, 0, 0, force=(1e-6, 1e-6, 1e-6))
    m.finalize()

    def check(state, tol=4e-7):
        assert abs(state["pos"][1][1] - 4e-9) < max(tol, 4e-7) * 4e-9, state["pos"][1]
    return m, 300, check


def combined_damping():
    """TEST(CVoxelyze, combinedDamping) first pass, tVoxelyze.h:336-363: voxel(3,0,0).z == 1.742e-8 +- 1e-10"""
    m, mat = _base(internal=1.0, glob=0.05)
    idx = None
    for i in range(4):
        for j in range(3):
            for k in range(3):
                v = m.set_voxel(mat, i, j, k)
                if i == 0:
                    m.set_fixed(i, j, k, 63)
                elif i == 3:
                    m.set_force(i, j, k, force=(0, 0, 1e-6))
                if (i, j, k) == (3, 0, 0):
                    idx = v
    m.finalize()

    def check(state, tol=0):
        assert abs(state["pos"][idx][2] - 1.742e-8) < 1e-10 + tol * 1.742e-8, state["pos"][idx]
    return m, 1000, check


def large_deformation_damping():
    """TEST(CVoxelyze, largeDeformationDamping), tVoxelyze.h:442-469: z == 9.5587e-4 +- 1e-7 after 200 steps"""
    m, mat = _base(internal=1.0, glob=0.2)
    m.set_voxel(mat, 0, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_voxel(mat, 1, 0, 0)
    m.set_force(1, 0, 0, force=(-0.2, 0.0, 0.2))
    m.finalize()

    def check(state, tol=0):
        assert abs(state["pos"][1][2] - 9.5587e-4) < 1e-7 + tol * 9.5587e-4, state["pos"][1]
    return m, 200, check


def multi_simple2():
    """TEST(CVoxelyze, multiSimple2), tVoxelyze.h:642-680: x(7) - 0.007 == 3.5035e-6 +- 1e-10 after 10000 steps"""
    m = gv.Model.new(0.001)
    m1 = m.add_material(1e6, 1e3, internal_damping=1.0, global_damping=0.01)
    m2 = m.add_material(1e9, 1e3, internal_damping=1.0, global_damping=0.01)
    m.set_world_param("dt_frac", 1.0)
    for i in range(8):
        m.set_voxel(m1 if (i // 2) % 2 == 0 else m2, i, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_force(7, 0, 0, force=(1e-3, 0, 0))
    m.finalize()

    def check(state, tol=0):
        assert abs((state["pos"][7][0] - 0.007) - 3.5035e-6) < 1e-10 + tol * 3.5035e-6, state["pos"][7]
    return m, 10000, check


def double_bond_cantilever(axis, sign):
    """TEST(CVoxelyze, doubleBondCantilever), tVoxelyze.h:520-566: tip deflection 1e-8 (axial) / 1.6e-7 (FL^3/3EI) within 1e-4"""
    m, mat = _base(internal=1.0, glob=0.1)
    for i in range(3):
        m.set_voxel(mat, i, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    i = axis
    fixed = (TX if i != 0 else 0) | (TY if i != 1 else 0) | (TZ if i != 2 else 0) | RX | (RY if i != 2 else 0) | (RZ if i != 1 else 0)
    m.set_fixed(2, 0, 0, fixed)
    f = [0.0, 0.0, 0.0]
    f[i] = sign * 5e-6
    m.set_force(2, 0, 0, force=f)
    m.finalize()
    expected = sign * (1e-8 if i == 0 else 1.6e-7)

    def check(state, tol=0):
        data = (state["pos"][2] - np.array([0.002, 0, 0]))[i]
        assert abs(data - expected) < abs(expected) * (1e-4 + tol), (data, expected)
    return m, 3000, check


def temperature(delta, expected):
    """TEST(CVoxelyze, temperature), tVoxelyze.h:990-1032: bimorph tip z = 2.55e-5 (+5 K) / -2.591e-5 (-5 K), +-1e-8"""
    m = gv.Model.new(0.001)
    m1 = m.add_material(1e6, 1e3, internal_damping=1.0, global_damping=0.15, cte=0.01)
    m2 = m.add_material(1e7, 1e3, internal_damping=1.0, global_damping=0.15)
    m.set_world_param("dt_frac", 1.0)
    idx = None
    for i in range(3):
        v = m.set_voxel(m1, i, 0, 0)
        if i == 2:
            idx = v
        m.set_voxel(m2, i, 0, 1)
    m.set_fixed(0, 0, 0, 63)
    m.set_fixed(0, 0, 1, 63)
    m.set_world_param("initial_temp", delta)
    m.finalize()

    def check(state, tol=0):
        assert abs(state["pos"][idx][2] - expected) < 1e-8 + tol * abs(expected), state["pos"][idx]
    return m, 500, check


def all_kats():
    k = [("singleBond-" + n, c) for n, c in single_bond_cases()]
    k += [("internalDamping", internal_damping()), ("globalDampingCantilever", global_damping_cantilever()),
          ("combinedDamping", combined_damping()), ("largeDeformationDamping", large_deformation_damping()),
          ("multiSimple2", multi_simple2()), ("temperature+5", temperature(5.0, 2.55e-5)), ("temperature-5", temperature(-5.0, -2.591e-5))]
    for axis in range(3):
        for sign in (1.0, -1.0):
            k.append((f"doubleBondCantilever-{axis}{'+' if sign > 0 else '-'}", double_bond_cantilever(axis, sign)))
    return k + more_kats()


# ---------------------------------------------------------------------------------------------------------------------
# More of the reference's suite.  Plain cases first (same (model, steps, check) form as above) ...
def scale(i):
    """TEST(CVoxelyze, scale), tVoxelyze.h:380-407: force 10^-i on every axis -> y = 4*10^-(i+3) within 0.1 % after 240 steps
    (the reference carries the state from one force level to the next; the equilibrium it checks does not depend on that)."""
    m, mat = _base(internal=1.0, glob=0.2)
    m.set_voxel(mat, 0, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_voxel(mat, 1, 0, 0)
    force = float(np.float32(1.0) / np.float32(10.0) ** np.float32(i))
    result = float(np.float32(4.0) / np.float32(10.0) ** np.float32(i + 3))
    m.set_force(1, 0, 0, force=(force, force, force))
    m.finalize()

    def check(state, tol=0):
        assert abs(state["pos"][1][1] - result) < result * (1e-3 + tol), (state["pos"][1], result)
    return m, 240, check


def multi_simple():
    """TEST(CVoxelyze, multiSimple), tVoxelyze.h:607-640: x(3) - 0.003 == 1.5015e-6 (EXPECT_FLOAT_EQ) after 800 steps"""
    m = gv.Model.new(0.001)
    m1 = m.add_material(1e6, 1e3, internal_damping=0.01, global_damping=0.03)
    m2 = m.add_material(1e9, 1e3, internal_damping=0.01, global_damping=0.03)
    m.set_world_param("dt_frac", 1.0)
    for i in range(2):
        m.set_voxel(m1, 2 * i, 0, 0)
        m.set_voxel(m2, 2 * i + 1, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    m.set_force(3, 0, 0, force=(1e-3, 0, 0))
    m.finalize()

    def check(state, tol=4e-7):
        data = state["pos"][3][0] - 0.003
        assert abs(data - 1.5015e-6) < 1.5015e-6 * max(tol, 4e-7), data
    return m, 800, check


def collisions():
    """TEST(CVoxelyze, collisions), tVoxelyze.h:1167-1197: a voxel dropped from two cells up onto a fixed voxel is held above it by
    the contact force: z > 0.001 after 150 steps (an inequality, so it also holds for VX3's all-pairs contact, which is what runs)"""
    m = gv.Model.new(0.001)
    mat = m.add_material(1e6, 1e6, global_damping=0.0)
    m.set_world_param("dt_frac", 1.0)
    m.set_world_param("floor", 1)
    m.set_world_param("gravity", 1.0)
    m.set_world_param("enable_collision", 1)
    m.set_voxel(mat, 0, 0, 0)
    m.set_fixed(0, 0, 0, 63)
    top = m.set_voxel(mat, 0, 0, 2)
    m.finalize()

    def check(state, tol=0):
        assert state["pos"][top][2] > 0.001, state["pos"][top]
    return m, 150, check


def more_kats():
    return [(f"scale-1e-{i}", scale(i)) for i in range(3, 8)] + [("multiSimple", multi_simple()), ("collisions", collisions())]


# ... then the scripted ones: the reference test reads the state every step or changes the load between steps.
# A script is (build() -> model, drive(world, tol)); `world` offers run(n), state(), links(), time(), set_force(voxel, force, moment),
# set_global_damping(mat, zeta).
def calc_period(timestep, data):
    """calcPeriod, tVoxelyze.h:42-61: period from the zero crossings of a trace"""
    zc = []
    for i in range(1, len(data)):
        if (data[i - 1] <= 0 and data[i] > 0) or (data[i - 1] >= 0 and data[i] < 0):
            zc.append(timestep * ((i - 1) + data[i - 1] / (data[i - 1] - data[i])))
    if len(zc) < 2:
        return -1.0
    return float(np.float32(np.mean(np.diff(zc)) * 2))


def frequency():
    """TEST(CVoxelyze, freqency), tVoxelyze.h:409-443: undamped axial oscillation, period 2*pi/sqrt(1e9) within 0.1 %"""
    def build():
        m = gv.Model.new(0.001)
        mat = m.add_material(1e6, 1e3, internal_damping=0.0)
        m.set_world_param("dt_frac", 0.1)  # ts = recommendedTimeStep()/10
        m.set_voxel(mat, 0, 0, 0)
        m.set_fixed(0, 0, 0, 63)
        m.set_voxel(mat, 1, 0, 0)
        m.set_fixed(1, 0, 0, 63 & ~TX)
        m.set_force(1, 0, 0, force=(1e-3, 0, 0))
        m.finalize()
        return m

    def drive(world, tol=0):
        data = []
        for _ in range(1000):
            world.run(1)
            data.append(world.state()["pos"][1][0] - 0.001001)
        ts = world.time() / 1000
        expected = 2 * 3.1415926 / np.sqrt(1e9)
        assert abs(calc_period(ts, data) - expected) < expected * (1e-3 + tol), (calc_period(ts, data), expected)
    return build, drive


def large_deformation():
    """TEST(CVoxelyze, largeDeformation), tVoxelyze.h:474-518: 100 rising loads of 100 steps each; the link is in the small-angle
    regime at load 0 and out of it at load 20; the second difference of the tip height stays below 5000 (no jump at the switch)"""
    def build():
        m, mat = _base(internal=1.0, glob=0.2)
        m.set_voxel(mat, 0, 0, 0)
        m.set_fixed(0, 0, 0, 63)
        m.set_voxel(mat, 1, 0, 0)
        m.finalize()
        return m

    def drive(world, tol=0):
        last = last_d = max_d2 = 0.0
        ts = None
        for i in range(100):
            world.set_force(1, force=(0.0, 0.0, float(np.float32(i) * np.float32(2e-3))))
            world.run(100)
            if ts is None:
                ts = float(np.float32(world.time() / 100))
            small = bool(world.links()["flags"][0] & 1)
            if i == 0:
                assert small
            if i == 20:
                assert not small
            val = float(np.float32(world.state()["pos"][1][2]))
            if i > 0:
                d = (val - last) / ts
                if i > 1:
                    max_d2 = max(max_d2, abs((d - last_d) / ts))
                last_d = d
            last = val
            assert max_d2 < 5000, (i, max_d2)
    return build, drive


def impulse():
    """TEST(CVoxelyze, impulse), tVoxelyze.h:568-605: a 100 N one-step kick on a 4x2 strip must die out: |z| < 1e-5 after 1000 steps"""
    idx = {}

    def build():
        m, mat = _base(internal=1.0, glob=0.05)
        for i in range(4):
            for j in range(2):
                v = m.set_voxel(mat, i, j, 0)
                if i == 0:
                    m.set_fixed(i, j, 0, 63)
                if (i, j) == (3, 0):
                    idx["v"] = v
        m.finalize()
        return m

    def drive(world, tol=0):
        v = idx["v"]
        world.run(10)
        world.set_force(v, force=(0, 0, 100))
        world.run(1)
        world.set_force(v, force=(0, 0, 0))
        world.run(989)
        z = world.state()["pos"][v][2]
        assert abs(float(np.float32(z))) < 1e-5, z
    return build, drive


def static_friction(gravity, mu_s, load, slides):
    """TEST(CVoxelyze, staticFriction), tVoxelyze.h:1033-1109: a voxel resting on the floor does not move AT ALL (x == 0.0) under a
    horizontal force below mu_s * N and does move above it; repeated at 2 g and with mu_s = 2.  (resetTime() between the
    reference's phases restores the initial state, so every phase is a fresh world here.)"""
    def build():
        m = gv.Model.new(0.001)
        mat = m.add_material(1e6, 1e3, static_friction=mu_s, kinetic_friction=0.1, global_damping=1.0)
        m.set_world_param("dt_frac", 1.0)
        m.set_world_param("floor", 1)
        m.set_world_param("gravity", gravity)
        m.set_voxel(mat, 0, 0, 0)
        m.finalize()
        return m

    def drive(world, tol=0):
        normal = float(np.float32(1e3 * 0.001 * 0.001 * 0.001 * 9.80665))
        world.run(50)  # let it settle down
        world.set_force(0, force=(float(np.float32(load) * np.float32(normal)), 0.0, 0.0))
        world.run(10)
        x = world.state()["pos"][0][0]
        assert (x != 0.0) if slides else (x == 0.0), x
    return build, drive


def kinetic_friction():
    """TEST(CVoxelyze, kineticFriction), tVoxelyze.h:1111-1158: a voxel settled on the floor, global damping then lowered to 1e-4,
    pushed with 2 N (mu_k = 0.1): after 10 steps the work of (F - mu_k N) equals the kinetic energy, ASSERT_NEAR(..., 2e-16)"""
    def build():
        m = gv.Model.new(0.001)
        mat = m.add_material(1e6, 1e3, static_friction=1.0, kinetic_friction=0.1, global_damping=1.0)
        m.set_world_param("dt_frac", 1.0)
        m.set_world_param("floor", 1)
        m.set_world_param("gravity", 1.0)
        m.set_voxel(mat, 0, 0, 0)
        m.finalize()
        return m

    def drive(world, tol=0):
        mass = 1e3 * 0.001 * 0.001 * 0.001
        normal = float(np.float32(mass * 9.80665))
        world.run(50)
        ts = float(np.float32(world.time() / 50))
        world.set_global_damping(0, 1e-4)
        h_force = float(np.float32(2.0) * np.float32(normal))
        world.set_force(0, force=(h_force, 0.0, 0.0))
        last = vel = energy = 0.0
        for step in range(10):
            world.run(1)
            cur = world.state()["pos"][0][0]
            vel = (cur - last) / ts
            factor = 0.5 if step == 9 else 1.0  # only half the energy of the last step: the discrete energy sum
            energy += float(np.float32(factor)) * (h_force - float(np.float32(0.1)) * normal) * (cur - last)
            last = cur
        assert abs(energy - 0.5 * mass * vel * vel) < 2e-16 + tol * abs(energy), (energy, 0.5 * mass * vel * vel)
        world.set_force(0, force=(0.0, 0.0, 0.0))
        world.run(200)  # (the reference only logs the slide to a stop)
        assert np.isfinite(world.state()["pos"]).all()
    return build, drive


def deformable_material():
    """TEST(CVoxelyze, deformableMaterial), tVoxelyze.h:845-886: a 5x3x3 bar of a BILINEAR material (E 1e6, plastic modulus 5e5, yield
    stress 1e5), clamped at x = 0, pulled with 0.2 N per voxel at x = 4 for 400 steps - past the yield point - then released for
    250 steps: the permanent set of voxel (4,1,1) is 4e-4 +- 1e-7 m (strain offset of VX_Link.cpp updateStrain)."""
    def build():
        m = gv.Model.new(0.001)
        mat = m.add_material(1e6, 1e3, internal_damping=1.0, global_damping=0.2)
        m.set_material_bilinear(mat, 1e6, 5e5, 1e5)
        m.set_world_param("dt_frac", 1.0)
        for i in range(5):
            for j in range(3):
                for k in range(3):
                    m.set_voxel(mat, i, j, k)
                    if i == 0:
                        m.set_fixed(i, j, k, 63)
                    if i == 4:
                        m.set_force(i, j, k, force=(0.2, 0.0, 0.0))
        m.finalize()
        return m

    def drive(world, tol=0):
        idx = lambda i, j, k: (i * 3 + j) * 3 + k  # noqa: E731
        world.run(400)
        stretched = world.state()["pos"][idx(4, 1, 1)][0] - 0.004
        assert stretched > 6e-4  # beyond the elastic limit: yield strain 0.1 over four links = 4e-4
        for j in range(3):
            for k in range(3):
                world.set_force(idx(4, j, k), force=(0.0, 0.0, 0.0))
        world.run(250)
        x = float(np.float32(world.state()["pos"][idx(4, 1, 1)][0] - 0.004))
        assert abs(x - 4e-4) < 1e-7 + tol * 4e-4, x
    return build, drive


def scripted_kats():
    k = [("freqency", frequency()), ("largeDeformation", large_deformation()), ("impulse", impulse()), ("kineticFriction", kinetic_friction()),
         ("deformableMaterial", deformable_material())]
    for g, mu, load, slides in ((1.0, 1.0, 0.9, False), (1.0, 1.0, 1.1, True), (2.0, 1.0, 1.9, False), (2.0, 1.0, 2.1, True),
                                (1.0, 2.0, 1.9, False), (1.0, 2.0, 2.1, True)):
        k.append((f"staticFriction-g{g:g}-mu{mu:g}-{load:g}N", static_friction(g, mu, load, slides)))
    return k
