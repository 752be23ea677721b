// Device-side physics of the Voxelyze step, storage-agnostic: every function works on values in
// registers so the streaming kernels (HBM-resident state), the resident kernel (state in shared
// memory) and the fused tile kernel all share one implementation.
//
// R = float  : production arithmetic.  Positions stay double in storage; the link pass subtracts
//              the two end positions and the nominal lattice offset in double and only then drops
//              to float, so strain resolution does not degrade as a body travels.
// R = double : validation arithmetic; reproduces the reference's float casts at the same places
//              (SURVEY.md §7 "mixed precision in the reference").
//
// Equations follow SURVEY.md Appendix A; reference locations are cited per function
// (paths under /root/reference/Voxelyze3/src).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vx3b200.h"

namespace vx3b {

#define VXD __device__ __forceinline__

template <class R> struct V3 {
    R x, y, z;
};
template <class R> struct Qt {
    R w, x, y, z;
};

template <class R> VXD V3<R> mk3(R x, R y, R z) { return V3<R>{x, y, z}; }
template <class R> VXD V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> VXD V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> VXD V3<R> operator-(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R> VXD V3<R> operator*(V3<R> a, R f) { return {f * a.x, f * a.y, f * a.z}; }
template <class R> VXD V3<R> operator*(R f, V3<R> a) { return {f * a.x, f * a.y, f * a.z}; }
template <class R> VXD R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> VXD R len2(V3<R> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

// float build: approximate sqrt / division / reciprocal (MUFU + one multiply, <= 2 ulp).  The IEEE forms cost a range check and,
// whenever an operand is zero or tiny - exactly what a body at rest feeds them - a ~100-instruction slow path (profiles/r01g).
VXD float vsqrt(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
VXD double vsqrt(double x) { return sqrt(x); }
VXD float vdiv(float a, float b) { return __fdividef(a, b); }
VXD double vdiv(double a, double b) { return a / b; }
VXD float vrcp(float a) {
    float r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
VXD double vrcp(double a) { return 1.0 / a; }
// float build: asin / acos without the libm slow paths (Cephes asinf core polynomial, |error| < 1e-7 on |x| <= 0.5, plus the
// half-angle reductions).  A crushed region of a big body keeps its links in the large-angle branch, where every warp that holds
// one such link pays for these (profiles/r01k: +30 M warp instructions per step on the 100^3 body once it has hit the floor).
VXD float asin_poly(float x) { // |x| <= 0.5
    const float z = x * x;
    const float p = (((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f;
    return x + x * z * p;
}
VXD float asin_fast(float x) { // 0 <= x <= 1
    const bool big = x > 0.5f;
    const float y = big ? vsqrt(0.5f * (1.0f - x)) : x;
    const float a = asin_poly(y);
    return big ? 1.5707963268f - 2.0f * a : a;
}
VXD float acos_fast(float x) { // -1 <= x <= 1
    const float ax = fabsf(x);
    const bool big = ax > 0.5f;
    const float y = big ? vsqrt(0.5f * (1.0f - ax)) : x;
    const float a = asin_poly(y);
    return big ? (x > 0.0f ? 2.0f * a : 3.14159265359f - 2.0f * a) : 1.5707963268f - a;
}
VXD float vabs(float x) { return fabsf(x); }
VXD double vabs(double x) { return fabs(x); }
VXD float vacos(float x) { return acosf(x); }
VXD double vacos(double x) { return acos(x); }
VXD float vasin(float x) { return asinf(x); }
VXD double vasin(double x) { return asin(x); }
VXD void vsincos(float a, float *s, float *c) { sincosf(a, s, c); }
VXD void vsincos(double a, double *s, double *c) { sincos(a, s, c); }

template <class R> VXD Qt<R> qmul(Qt<R> a, Qt<R> f) { // Utils/VX3_Quat3D.h:217-222
    return {a.w * f.w - a.x * f.x - a.y * f.y - a.z * f.z, a.w * f.x + a.x * f.w + a.y * f.z - a.z * f.y,
            a.w * f.y - a.x * f.z + a.y * f.w + a.z * f.x, a.w * f.z + a.x * f.y - a.y * f.x + a.z * f.w};
}
template <class R> VXD Qt<R> qconj(Qt<R> a) { return {a.w, -a.x, -a.y, -a.z}; }
template <class R> VXD V3<R> qrot(Qt<R> q, V3<R> f) { // RotateVec3D, Utils/VX3_Quat3D.h:439-448
    R tw = f.x * q.x + f.y * q.y + f.z * q.z;
    R tx = f.x * q.w - f.y * q.z + f.z * q.y;
    R ty = f.x * q.z + f.y * q.w - f.z * q.x;
    R tz = -f.x * q.y + f.y * q.x + f.z * q.w;
    return {q.w * tx + q.x * tw + q.y * tz - q.z * ty, q.w * ty - q.x * tz + q.y * tw + q.z * tx, q.w * tz + q.x * ty - q.y * tx + q.z * tw};
}
template <class R> VXD V3<R> qrot_inv(Qt<R> q, V3<R> f) { // RotateVec3DInv, Utils/VX3_Quat3D.h:465-474
    R tw = q.x * f.x + q.y * f.y + q.z * f.z;
    R tx = q.w * f.x - q.y * f.z + q.z * f.y;
    R ty = q.w * f.y + q.x * f.z - q.z * f.x;
    R tz = q.w * f.z - q.x * f.y + q.y * f.x;
    return {tw * q.x + tx * q.w + ty * q.z - tz * q.y, tw * q.y - tx * q.z + ty * q.w + tz * q.x, tw * q.z + tx * q.y - ty * q.x + tz * q.w};
}

// Quaternion -> rotation vector (Utils/VX3_Quat3D.h:347-359).
VXD V3<double> rotvec(Qt<double> q) {
    if (q.w >= 1.0 || q.w <= -1.0)
        return {0, 0, 0};
    double sl = 1.0 - q.w * q.w;
    V3<double> v2{2.0 * q.x, 2.0 * q.y, 2.0 * q.z};
    if (sl < 2.4e-3)
        return v2 * sqrt((2 - 2 * q.w) / sl);
    double inv = 1.0 / sqrt(sl);
    return (v2 * acos(q.w)) * inv;
}
// float: 1 - w*w cancels catastrophically near identity, so use |v| directly (equal for unit q):
// rotvec = 2 * v * asin(|v|)/|v|, series near zero.
VXD V3<float> rotvec(Qt<float> q) {
    float s2 = q.x * q.x + q.y * q.y + q.z * q.z;
    float k;
    if (s2 < 2.4e-3f)
        k = 2.0f + s2 * (0.33333334f + s2 * 0.15f); // 2*asin(s)/s = 2 + s^2/3 + 3 s^4/20 + ...
    else {
        float s = vsqrt(s2);
        // asin of the vector part where it is well conditioned, acos(w) (the reference's branch) elsewhere
        const float half = (q.w > 0.70710678f) ? asin_fast(fminf(s, 1.0f)) : acos_fast(fminf(fmaxf(q.w, -1.0f), 1.0f));
        k = vdiv(2.0f * half, s);
    }
    return {k * q.x, k * q.y, k * q.z};
}

// Rotation vector -> quaternion (Utils/VX3_Quat3D.h:363-382).
VXD Qt<double> quat_from_rotvec(V3<double> in) {
    V3<double> th{0.5 * in.x, 0.5 * in.y, 0.5 * in.z};
    double m2 = len2(th), w, s;
    if (m2 * m2 < 5.328e-15) {
        w = 1.0 - 0.5 * m2;
        s = 1.0 - m2 / 6.0;
    } else {
        double m = sqrt(m2), sn, cs;
        sincos(m, &sn, &cs);
        w = cs;
        s = sn / m;
    }
    return {w, th.x * s, th.y * s, th.z * s};
}
VXD Qt<float> quat_from_rotvec(V3<float> in) {
    V3<float> th{0.5f * in.x, 0.5f * in.y, 0.5f * in.z};
    float m2 = len2(th), w, s;
    if (m2 < 1.0e-3f) { // two more series terms than the reference's double branch: exact to float precision here
        w = 1.0f - m2 * (0.5f - m2 * 0.041666668f);
        s = 1.0f - m2 * (0.16666667f - m2 * 0.0083333338f);
    } else {
        float m = vsqrt(m2), sn, cs;
        __sincosf(m, &sn, &cs); // |m| is a fraction of a radian here: the fast path's absolute error (2^-21) is below float resolution of w
        w = cs;
        s = vdiv(sn, m);
    }
    return {w, th.x * s, th.y * s, th.z * s};
}

// Quaternion that rotates `from` onto +X (Utils/VX3_Quat3D.h:384-437); q is left untouched for a zero vector.
template <class R> VXD void quat_to_pos_x(Qt<R> &q, V3<R> from) {
    if (from.x == R(0) && from.y == R(0) && from.z == R(0))
        return;
    const R SMALL = R(1.732e-2);
    R yox = vdiv(from.y, from.x), zox = vdiv(from.z, from.x);
    if (yox < SMALL && yox > -SMALL && zox < SMALL && zox > -SMALL) {
        q.x = R(0);
        q.y = R(0.5) * zox;
        q.z = R(-0.5) * yox;
        q.w = R(1) + R(0.5) * (-q.y * q.y - q.z * q.z);
        return;
    }
    if constexpr (sizeof(R) == 4) {
        // float build, without trigonometry: the unit quaternion taking n to +X is (1 + n.x, n x X) normalised
        // = (cos(t/2), sin(t/2) * (0, n.z, -n.y) / |(n.y, n.z)|), the same rotation the reference builds from acos / sincos
        const R inv = rsqrtf(len2(from));
        const V3<R> n{inv * from.x, inv * from.y, inv * from.z};
        const R c1 = R(1) + n.x;
        if (c1 < R(1e-7)) { // pointing along -X: the reference's 180 degree case
            q = {R(0), R(0), R(1), R(0)};
            return;
        }
        const R k = rsqrtf(R(2) * c1);
        q = {c1 * k, R(0), n.z * k, -n.y * k};
    } else {
        R l = vsqrt(len2(from));
        V3<R> n = from;
        if (l > R(0)) {
            R inv = vrcp(l);
            n = {inv * from.x, inv * from.y, inv * from.z};
        }
        R theta = vacos(n.x);
        if (theta > R(3.14159265358979) - R(1e-7)) {
            q = {R(0), R(0), R(1), R(0)};
            return;
        }
        R axisMagInv = vrcp(vsqrt(n.z * n.z + n.y * n.y));
        R sn, cs;
        vsincos(R(0.5) * theta, &sn, &cs);
        q.w = cs;
        q.x = R(0);
        q.y = n.z * axisMagInv * sn;
        q.z = -n.y * axisMagInv * sn;
    }
}

// axis swizzles (VX3/VX3_Link.h:97-191)
template <class R> VXD V3<R> to_axis_x(V3<R> v, int axis) {
    return axis == 1 ? V3<R>{v.y, -v.x, v.z} : axis == 2 ? V3<R>{v.z, v.y, -v.x} : v;
}
template <class R> VXD Qt<R> to_axis_x(Qt<R> q, int axis) {
    return axis == 1 ? Qt<R>{q.w, q.y, -q.x, q.z} : axis == 2 ? Qt<R>{q.w, q.z, q.y, -q.x} : q;
}
template <class R> VXD V3<R> to_axis_original(V3<R> v, int axis) {
    return axis == 1 ? V3<R>{-v.y, v.x, v.z} : axis == 2 ? V3<R>{-v.z, v.y, v.x} : v;
}

// ------------------------------------------------------------------------------------------
// device-side constant tables
// ------------------------------------------------------------------------------------------
struct DVoxMat { // per voxel material, pre-multiplied where the reference multiplies two floats every step
    float E, alphaCTE, muStatic, muKinetic;
    float mass, massInverse, momentInertiaInverse, sqrtMassZetaInt2; // 2*sqrtMass*zetaInternal (numerator of dampingMultiplier)
    float globalDampTranslateC, globalDampRotateC, collisionDampTranslateC, penetrationStiffness;
    float gravityForce; // -mass*9.80665f*gravMult
    float nomSizeF;     // (float)nomSize
    double nomSize;
    double cilia;
    // sizes when the voxel temperature cannot change them (world without thermal actuation, or alphaCTE == 0): the values the
    // expressions below produce, evaluated once on the host (they contain a double division the step would otherwise repeat)
    double baseSize0, avgSize0, floorOffset0; // base_size, base_size_average, avg/2 - nomSize/2
    int32_t fixed, isTarget, isMeasured, sizeIsConst;
};
struct DLinkMat {
    float E, a2, b1, b2, b3, sqA1, sqA2xIp, sqB1, sqB2xFMp, sqB3xIp, epsilonFail, area; // area: currentTransverseArea (nu == 0)
    uint32_t matPair; // world-local voxel materials the link material combines: mat1 | mat2 << 16
};

// Stress/strain model of a link material for worlds with bilinear / data materials or a non-zero Poisson's ratio (the streaming
// kernels; every other world's links are the single line E of DLinkMat): CVX_Material::strainData / stressData, nu, eHat.
struct DLinkModel {
    int32_t linear, npts;
    float E, nu, eHat, epsilonFail;
    float strain[VX3B_MAX_SS_POINTS], stress[VX3B_MAX_SS_POINTS];
};
// VX3_Material::stress (VX3/VX3_Material.cu:90-124).  transverseStrainSum is the link's currentTransverseStrainSum, which Voxelyze3
// never refreshes after construction (updateTransverseInfo() is commented out, VX3/VX3_Link.cu:147-150): 0.
VXD float material_stress(const DLinkModel &m, float strain, float transverseStrainSum, bool forceLinear) {
    if (m.epsilonFail != -1.0f && strain > m.epsilonFail)
        return 0.0f;
    if (strain <= m.strain[1] || m.linear || forceLinear) {
        if (m.nu == 0.0f)
            return m.E * strain;
        return m.eHat * ((1 - m.nu) * strain + m.nu * transverseStrainSum);
    }
    const int n = m.npts;
    for (int i = 2; i < n; i++) {
        if (strain <= m.strain[i] || i == n - 1) {
            const float Perc = (strain - m.strain[i - 1]) / (m.strain[i] - m.strain[i - 1]);
            const float basicStress = m.stress[i - 1] + Perc * (m.stress[i] - m.stress[i - 1]);
            if (m.nu == 0.0f)
                return basicStress;
            const float modulus = (m.stress[i] - m.stress[i - 1]) / (m.strain[i] - m.strain[i - 1]);
            const float modulusHat = modulus / ((1 - 2 * m.nu) * (1 + m.nu));
            const float effectiveStrain = basicStress / modulus;
            const float effectiveTransverseStrainSum = transverseStrainSum * (effectiveStrain / strain);
            return modulusHat * ((1 - m.nu) * effectiveStrain + m.nu * effectiveTransverseStrainSum);
        }
    }
    return 0.0f;
}
// VX3_Link::updateStrain (VX3/VX3_Link.cu:232-268): a non-linear material remembers its largest strain and backs off along E
VXD float update_strain(const DLinkModel &m, float axialStrain, float &maxStrain, float &strainOffset) {
    const float tss = 0.0f;
    if (m.linear) {
        if (axialStrain > maxStrain)
            maxStrain = axialStrain;
        return material_stress(m, axialStrain, tss, false);
    }
    float returnStress;
    if (axialStrain > maxStrain) { // new territory on the stress/strain curve
        maxStrain = axialStrain;
        returnStress = material_stress(m, axialStrain, tss, false);
        if (m.nu != 0.0f)
            strainOffset = maxStrain - material_stress(m, axialStrain, 0.0f, false) / (m.eHat * (1 - m.nu));
        else
            strainOffset = maxStrain - returnStress / m.E;
    } else { // backed off a non-linear material: linear with a strain offset according to the maximum plastic deformation
        const float relativeStrain = axialStrain - strainOffset;
        returnStress = m.nu != 0.0f ? material_stress(m, relativeStrain, tss, true) : m.E * relativeStrain;
    }
    return returnStress;
}

// link history flags
enum { LF_SMALL_ANGLE = 1, LF_VEL_VALID = 2 };
// voxel flags
enum { VF_STATIC_FRICTION = 1, VF_ZERO_ANGMOM = 2 }; // VF_ZERO_ANGMOM: brick-ordered sets only (k_brick applies the VX3 zeroing lazily)

template <class R> struct LinkHist {
    V3<R> pos2, a1v, a2v;
    uint32_t flags;
    float maxStrain;
};
template <class R> struct LinkForces {
    V3<R> fNeg, fPos, mNeg, mPos;
    float strain;
};

// base size along one axis = nomSize * (1 + temp*alphaCTE), the bracket being a float expression (VX3/VX3_Voxel.h:96-99)
__host__ __device__ inline double base_size_expr(double nomSize, float alphaCTE, float temp) { return nomSize * (double)(1 + temp * alphaCTE); }
__host__ __device__ inline double base_size_average_expr(double b) { return (b + b + b) / 3.0f; } // VX3/VX3_Voxel.h:102-105
VXD double base_size(const DVoxMat &m, float temp) { return m.sizeIsConst ? m.baseSize0 : base_size_expr(m.nomSize, m.alphaCTE, temp); }
VXD double base_size_average(const DVoxMat &m, float temp) {
    return m.sizeIsConst ? m.avgSize0 : base_size_average_expr(base_size_expr(m.nomSize, m.alphaCTE, temp));
}

// One link: orientLink + updateForces (VX3/VX3_Link.cu:90-218, updateStrain :220-231 linear / nu == 0 branch).
//   dpos    = p₊ − p₋ (double, global frame); qn/qp = end orientations; restLen from both ends' temperatures
//   dmN/dmP = dampingMultiplier of the negative/positive end voxel (2*sqrtMass*zetaInternal/previousDt)
// lengths of one link in the arithmetic type: the nominal lattice spacing stays double (it is subtracted from a double position
// difference), the rest are what the float formulation needs; constant per link material while sizes do not change
template <class R> struct LinkLen {
    double nomLen;
    R rest, nomMinusRest, invRest; // (R)restLen, (R)(nomLen - restLen), 1/rest (float build only)
};
template <class R> VXD LinkLen<R> link_len(double nomLen, double restLen) {
    LinkLen<R> l;
    l.nomLen = nomLen;
    l.rest = (R)restLen;
    l.nomMinusRest = (R)(nomLen - restLen);
    l.invRest = vrcp(l.rest);
    return l;
}
// AXIS >= 0: the link axis is a compile-time constant (the caller dispatches once per link; the seven swizzles fold away);
// AXIS == -1: run-time axis
// model / strainOffset: the link material's stress/strain model and the link's strain offset, for worlds that need them (nullptr
// otherwise: the single line E of `m`)
template <class R, int AXIS = -1>
VXD void link_update(const DLinkMat &m, int axisRt, V3<double> dpos, const LinkLen<R> len, Qt<R> qn, Qt<R> qp, float dmN, float dmP, LinkHist<R> &h,
                     LinkForces<R> &out, const DLinkModel *model = nullptr, float *strainOffset = nullptr) {
    const int axis = AXIS >= 0 ? AXIS : axisRt;
    const double nomLen = len.nomLen;
    const V3<R> oldPos2 = h.pos2, oldA1 = h.a1v, oldA2 = h.a2v;
    Qt<R> q1 = to_axis_x(qn, axis), q2 = to_axis_x(qp, axis);
    Qt<R> total = qconj(q1);
    q2 = qmul(total, q2);
    V3<R> d; // end-to-end vector in the negative end's frame, link axis on +X
    const R rest = len.rest;
    R dxMinusRest; // d.x - restLen, formed without cancellation in float
    if constexpr (sizeof(R) == 8) {
        d = qrot(total, to_axis_x(V3<R>{dpos.x, dpos.y, dpos.z}, axis));
        dxMinusRest = d.x - rest;
    } else {
        // d = Rᵀ(e + L·x̂) with e = dpos − L·x̂ small: rotate e in float, add L·Rᵀx̂ analytically
        V3<double> dd = to_axis_x(dpos, axis);
        V3<R> e{(R)(dd.x - nomLen), (R)dd.y, (R)dd.z};
        V3<R> er = qrot(total, e);
        R L = (R)nomLen;
        // Rᵀx̂ for unit q1=(w,x,y,z): (1-2(y²+z²), 2(xy - wz), 2(xz + wy))
        R cx = R(-2) * (q1.y * q1.y + q1.z * q1.z); // = Rᵀx̂.x - 1
        R cy = R(2) * (q1.x * q1.y - q1.w * q1.z);
        R cz = R(2) * (q1.x * q1.z + q1.w * q1.y);
        d = {er.x + L * cx + L, er.y + L * cy, er.z + L * cz};
        dxMinusRest = er.x + L * cx + len.nomMinusRest;
    }
    q1 = {R(1), R(0), R(0), R(0)};
    float smallTurn = (float)vdiv(vabs(d.z) + vabs(d.y), d.x);
    float extendPerc; // |1 - d.x/restLen|
    if constexpr (sizeof(R) == 8)
        extendPerc = (float)vabs(R(1) - d.x / rest);
    else
        extendPerc = vabs(dxMinusRest * len.invRest);
    const float HYST = 1.2f, BEND = 0.05f, EXT = 0.50f; // VoxCAD/src/VX2/include/types.h:57-59
    bool small = h.flags & LF_SMALL_ANGLE;
    bool valid = h.flags & LF_VEL_VALID;
    if (!small && smallTurn < BEND && extendPerc < EXT) {
        small = true;
        valid = false;
    } else if (small && (smallTurn > HYST * BEND || extendPerc > HYST * EXT)) {
        small = false;
        valid = false;
    }
    V3<R> th1{R(0), R(0), R(0)};
    if (small) {
        d.x = dxMinusRest;
    } else {
        quat_to_pos_x(q1, d);
        total = qmul(q1, total);
        q2 = qmul(q1, q2);
        d = {vsqrt(len2(d)) - rest, R(0), R(0)};
        th1 = rotvec(q1);
    }
    V3<R> th2 = rotvec(q2);
    h.pos2 = d;
    h.a1v = th1;
    h.a2v = th2;

    float eps;
    if constexpr (sizeof(R) == 8)
        eps = (float)(d.x / rest);
    else
        eps = d.x * len.invRest;
    out.strain = eps;
    const bool canFail = m.epsilonFail != -1.0f;
    float stress;
    if (model)
        stress = update_strain(*model, eps, h.maxStrain, *strainOffset);
    else {
        if (eps > h.maxStrain)
            h.maxStrain = eps;
        stress = (canFail && eps > m.epsilonFail) ? 0.0f : m.E * eps;
    }
    if (canFail && h.maxStrain > m.epsilonFail) { // isFailed(): outputs zero, flags keep their pre-damping value
        out.fNeg = out.fPos = out.mNeg = out.mPos = V3<R>{R(0), R(0), R(0)};
        h.flags = (small ? LF_SMALL_ANGLE : 0) | (valid ? LF_VEL_VALID : 0);
        return;
    }
    const R b1 = m.b1, b2 = m.b2, b3 = m.b3, a2 = m.a2;
    V3<R> FN{(R)(stress * m.area), b1 * d.y - b2 * (th1.z + th2.z), b1 * d.z + b2 * (th1.y + th2.y)};
    V3<R> FP = -FN;
    V3<R> MN{a2 * (th2.x - th1.x), -b2 * d.z - b3 * (R(2) * th1.y + th2.y), b2 * d.y - b3 * (R(2) * th1.z + th2.z)};
    V3<R> MP{a2 * (th1.x - th2.x), -b2 * d.z - b3 * (th1.y + R(2) * th2.y), b2 * d.y - b3 * (th1.z + R(2) * th2.z)};
    if (valid) { // local damping from half-deltas of the link's own history (VX3/VX3_Link.cu:172-188)
        V3<R> dP = R(0.5) * (d - oldPos2), dA1 = R(0.5) * (th1 - oldA1), dA2 = R(0.5) * (th2 - oldA2);
        const R sqA1 = m.sqA1, sqA2xIp = m.sqA2xIp, sqB1 = m.sqB1, sqB2xFMp = m.sqB2xFMp, sqB3xIp = m.sqB3xIp;
        V3<R> pc{sqA1 * dP.x, sqB1 * dP.y - sqB2xFMp * (dA1.z + dA2.z), sqB1 * dP.z + sqB2xFMp * (dA1.y + dA2.y)};
        FN = FN + (R)dmN * pc;
        FP = FP - (R)dmP * pc;
        R tw = sqA2xIp * (dA2.x - dA1.x);
        MN = MN - (R(0.5) * (R)dmN) * V3<R>{-tw, sqB2xFMp * dP.z + sqB3xIp * (R(2) * dA1.y + dA2.y), -sqB2xFMp * dP.y + sqB3xIp * (R(2) * dA1.z + dA2.z)};
        MP = MP - (R(0.5) * (R)dmP) * V3<R>{tw, sqB2xFMp * dP.z + sqB3xIp * (dA1.y + R(2) * dA2.y), -sqB2xFMp * dP.y + sqB3xIp * (dA1.z + R(2) * dA2.z)};
    } else
        valid = true;
    if (!small) {
        FN = qrot_inv(q1, FN);
        MN = qrot_inv(q1, MN);
    }
    FP = qrot_inv(q2, FP);
    MP = qrot_inv(q2, MP);
    out.fNeg = to_axis_original(FN, axis);
    out.fPos = to_axis_original(FP, axis);
    out.mNeg = to_axis_original(MN, axis);
    out.mPos = to_axis_original(MP, axis);
    h.flags = (small ? LF_SMALL_ANGLE : 0) | (valid ? LF_VEL_VALID : 0);
}
template <class R>
VXD void link_update(const DLinkMat &m, int axis, V3<double> dpos, double nomLen, Qt<R> qn, Qt<R> qp, double restLen, float dmN, float dmP,
                     LinkHist<R> &h, LinkForces<R> &out, const DLinkModel *model = nullptr, float *strainOffset = nullptr) {
    link_update<R, -1>(m, axis, dpos, link_len<R>(nomLen, restLen), qn, qp, dmN, dmP, h, out, model, strainOffset);
}

// ------------------------------------------------------------------------------------------
// formula VM (Utils/VX3_MathTree.h:50-193): FIFO queue machine over a reversed level-order program
// ------------------------------------------------------------------------------------------
struct FormulaVars {
    double x, y, z, hit, t, angle, closeness;
    int numClosePairs, numVoxel;
};
__device__ inline double eval_formula(const VX3B_Token *tok, const FormulaVars &v) {
    double values[VX3B_MAX_TOKENS + 1];
    values[0] = 0.0;
    int vc = 0, pc = 0;
    for (int i = 0; i < VX3B_MAX_TOKENS; i++) {
        const int op = tok[i].op;
        double r = 0;
        if (op == VX3B_MT_END)
            return values[pc];
        if (op == VX3B_MT_CONST)
            r = tok[i].value;
        else if (op == VX3B_MT_E)
            r = 2.71828182845904523536;
        else if (op == VX3B_MT_PI)
            r = 3.14159265358979323846;
        else if (op == VX3B_MT_VAR) {
            double id = tok[i].value;
            r = id < 0.5 ? v.x : id < 1.5 ? v.y : id < 2.5 ? v.z : id < 3.5 ? v.hit : id < 4.5 ? v.t : id < 5.5 ? v.angle
                : id < 6.5 ? v.closeness : id < 7.5 ? (double)v.numClosePairs : (double)v.numVoxel;
        } else if (op == VX3B_MT_ADD || op == VX3B_MT_SUB || op == VX3B_MT_MUL || op == VX3B_MT_DIV || op == VX3B_MT_POW ||
                   op == VX3B_MT_GREATERTHAN || op == VX3B_MT_LESSTHAN || op == VX3B_MT_AND || op == VX3B_MT_OR) {
            double a = values[pc + 1], b = values[pc];
            pc += 2;
            r = op == VX3B_MT_ADD ? a + b : op == VX3B_MT_SUB ? a - b : op == VX3B_MT_MUL ? a * b : op == VX3B_MT_DIV ? a / b
                : op == VX3B_MT_POW ? pow(a, b) : op == VX3B_MT_GREATERTHAN ? (double)(a > b) : op == VX3B_MT_LESSTHAN ? (double)(a < b)
                : op == VX3B_MT_AND ? (double)(a != 0 && b != 0) : (double)(a != 0 || b != 0);
        } else {
            double a = values[pc];
            pc += 1;
            switch (op) {
            case VX3B_MT_SQRT: r = sqrt(a); break;
            case VX3B_MT_SIN: r = sin(a); break;
            case VX3B_MT_COS: r = cos(a); break;
            case VX3B_MT_TAN: r = tan(a); break;
            case VX3B_MT_ATAN: r = atan(a); break;
            case VX3B_MT_LOG: r = log(a); break;
            case VX3B_MT_INT: r = rint(a); break;
            case VX3B_MT_ABS: r = fabs(a); break;
            case VX3B_MT_NOT: r = (double)(a == 0); break;
            case VX3B_MT_NORMALCDF: r = normcdf(a); break;
            default: return -1;
            }
        }
        if (pc > vc)
            return -1;
        values[vc++] = r;
    }
    return -1;
}

// ------------------------------------------------------------------------------------------
// voxel integration (VX3/VX3_Voxel.cu:162-277, force :350-380, moment :382-397, floorForce :399-426)
// ------------------------------------------------------------------------------------------
struct WorldStepParams { // what one voxel needs to know about its world for one step
    float dt;
    int floorEnabled, vx3;
    const VX3B_Token *forceField; // [3][VX3B_MAX_TOKENS] or nullptr
    FormulaVars fv;               // hit/t/angle/closeness/pairs/nv for the force field
};
template <class R> struct VoxState {
    V3<double> pos;
    Qt<R> orient;
    V3<R> linMom, angMom;
    uint32_t flags;
    float temp;
    // baseSizeAverage()/2 - nominalSize()/2 at `temp` when the caller keeps it per voxel (k_resident caches baseSizeAverage of
    // thermally actuated voxels: it costs a double division, and the floor and every contact candidate need it)
    bool haveFloorOffset = false;
    double floorOffset;
};
struct VoxExt { // boundary conditions of one voxel (CVX_External)
    uint32_t dof;
    V3<float> force, moment;
    V3<double> translation;
    Qt<double> rotationQ;
};

VXD float floor_penetration(const DVoxMat &m, float temp, double z) { // VX3/VX3_Voxel.h:185-187
    if (m.sizeIsConst)
        return (float)(m.floorOffset0 - z);
    return (float)(base_size_average(m, temp) / 2 - m.nomSize / 2 - z);
}
template <class R> VXD float floor_penetration(const DVoxMat &m, const VoxState<R> &s, double z) {
    if (s.haveFloorOffset) // given by the caller (same expression, evaluated once per temperature change)
        return (float)(s.floorOffset - z);
    return floor_penetration(m, s.temp, z);
}

// Fsum/Msum: sums of the ≤6 link forces/moments in the voxel's local frame; contact/cilia in the global frame.
// PLAIN: the caller guarantees no boundary conditions, force field, contact or cilia force (compiled away).
// The update is the composition of a translation half (reads orient; writes pos, linMom; returns the friction mode) and a
// rotation half (writes angMom, orient) that depend on each other only through VX3's "static friction zeroes angMom" rule, which
// the caller applies afterwards - so a kernel may give the two halves of one voxel to two threads (k_brick).
template <class R, bool PLAIN = false>
VXD bool voxel_translate(const DVoxMat &m, const WorldStepParams &wp, const VoxExt *ext, VoxState<R> &s, V3<R> Fsum, V3<R> contact, V3<R> cilia) {
    const R dt = (R)wp.dt;
    if constexpr (PLAIN)
        ext = nullptr;
    V3<R> F = qrot(s.orient, Fsum);
    if (ext)
        F = F + V3<R>{(R)ext->force.x, (R)ext->force.y, (R)ext->force.z};
    F = F - (s.linMom * (R)m.massInverse) * (R)m.globalDampTranslateC;
    F.z += (R)m.gravityForce;
    if constexpr (!PLAIN) {
        F = F - contact;
        F = F + cilia * (R)m.cilia;
    }
    if (!PLAIN && wp.forceField) {
        FormulaVars fv = wp.fv;
        fv.x = s.pos.x;
        fv.y = s.pos.y;
        fv.z = s.pos.z;
        if (wp.forceField[0].op != VX3B_MT_END)
            F.x += (R)eval_formula(wp.forceField, fv);
        if (wp.forceField[VX3B_MAX_TOKENS].op != VX3B_MT_END)
            F.y += (R)eval_formula(wp.forceField + VX3B_MAX_TOKENS, fv);
        if (wp.forceField[2 * VX3B_MAX_TOKENS].op != VX3B_MT_END)
            F.z += (R)eval_formula(wp.forceField + 2 * VX3B_MAX_TOKENS, fv);
    }
    V3<R> fric = F;
    bool staticFric = s.flags & VF_STATIC_FRICTION;
    bool onFloor = false; // floor enabled and penetrating at the start of the step
    if (wp.floorEnabled) {
        float pen = floor_penetration(m, s, s.pos.z);
        onFloor = pen >= 0;
        if (pen >= 0) {
            V3<R> vel = s.linMom * (R)m.massInverse;
            float normalForce = m.penetrationStiffness * pen;
            F.z += (R)normalForce - (R)m.collisionDampTranslateC * vel.z;
            if (staticFric) {
                float surfaceSq = (float)(F.x * F.x + F.y * F.y);
                float frictionSq = (m.muStatic * normalForce) * (m.muStatic * normalForce);
                if (surfaceSq > frictionSq)
                    staticFric = false;
            } else {
                R hl = vsqrt(vel.x * vel.x + vel.y * vel.y);
                if (hl > R(0)) {
                    R k = (R)(m.muKinetic * normalForce) * vrcp(hl);
                    F.x -= k * vel.x;
                    F.y -= k * vel.y;
                }
            }
        } else
            staticFric = false;
    }
    fric = F - fric;
    s.linMom = s.linMom + F * dt;
    V3<R> tr = s.linMom * (dt * (R)m.massInverse);
    if (onFloor) {
        R work = fric.x * tr.x + fric.y * tr.y;
        R hKe = (R(0.5) * (R)m.massInverse) * (s.linMom.x * s.linMom.x + s.linMom.y * s.linMom.y);
        if (hKe + work <= R(0))
            staticFric = true;
        if (staticFric) {
            s.linMom.x = s.linMom.y = R(0);
            tr.x = tr.y = R(0);
        }
    } else
        staticFric = false;
    s.pos = {s.pos.x + (double)tr.x, s.pos.y + (double)tr.y, s.pos.z + (double)tr.z};
    return staticFric;
}
template <class R, bool PLAIN = false> VXD void voxel_rotate(const DVoxMat &m, const WorldStepParams &wp, const VoxExt *ext, VoxState<R> &s, V3<R> Msum) {
    const R dt = (R)wp.dt;
    if constexpr (PLAIN)
        ext = nullptr;
    V3<R> Mo = qrot(s.orient, Msum);
    if (ext)
        Mo = Mo + V3<R>{(R)ext->moment.x, (R)ext->moment.y, (R)ext->moment.z};
    Mo = Mo - (s.angMom * (R)m.momentInertiaInverse) * (R)m.globalDampRotateC;
    s.angMom = s.angMom + Mo * dt;
    s.orient = qmul(quat_from_rotvec(s.angMom * (dt * (R)m.momentInertiaInverse)), s.orient);
    if constexpr (sizeof(R) == 4) { // float build keeps the quaternion unit length; the reference (double) never renormalises
        R n2 = s.orient.w * s.orient.w + s.orient.x * s.orient.x + s.orient.y * s.orient.y + s.orient.z * s.orient.z;
        R inv = rsqrtf(n2);
        s.orient = {s.orient.w * inv, s.orient.x * inv, s.orient.y * inv, s.orient.z * inv};
    }
}
// VX3 only: a voxel resting on the floor in static friction has its angular momentum cleared (VX3/VX3_Voxel.cu:258-264)
VXD bool voxel_zero_angmom(const DVoxMat &m, const WorldStepParams &wp, bool staticFric, float temp, double newZ) {
    return wp.vx3 && wp.floorEnabled && staticFric && floor_penetration(m, temp, newZ) >= 0;
}
template <class R> VXD bool voxel_zero_angmom(const DVoxMat &m, const WorldStepParams &wp, bool staticFric, const VoxState<R> &s) {
    return wp.vx3 && wp.floorEnabled && staticFric && floor_penetration(m, s, s.pos.z) >= 0;
}
template <class R, bool PLAIN = false>
VXD void voxel_update(const DVoxMat &m, const WorldStepParams &wp, const VoxExt *ext, V3<double> nominalPos, VoxState<R> &s, V3<R> Fsum, V3<R> Msum,
                      V3<R> contact, V3<R> cilia) {
    if constexpr (PLAIN)
        ext = nullptr;
    if (ext && ext->dof == VX3B_DOF_ALL) {
        s.pos = nominalPos + ext->translation;
        s.orient = {(R)ext->rotationQ.w, (R)ext->rotationQ.x, (R)ext->rotationQ.y, (R)ext->rotationQ.z};
        s.linMom = s.angMom = V3<R>{R(0), R(0), R(0)};
        return;
    }
    const bool staticFric = voxel_translate<R, PLAIN>(m, wp, ext, s, Fsum, contact, cilia);
    voxel_rotate<R, PLAIN>(m, wp, ext, s, Msum);
    if (ext && ext->dof) { // partial dof fixes (VX3/VX3_Voxel.cu:222-257)
        const uint32_t f = ext->dof;
        if (f & VX3B_DOF_TX) {
            s.pos.x = nominalPos.x + ext->translation.x;
            s.linMom.x = R(0);
        }
        if (f & VX3B_DOF_TY) {
            s.pos.y = nominalPos.y + ext->translation.y;
            s.linMom.y = R(0);
        }
        if (f & VX3B_DOF_TZ) {
            s.pos.z = nominalPos.z + ext->translation.z;
            s.linMom.z = R(0);
        }
        const uint32_t rotAll = VX3B_DOF_RX | VX3B_DOF_RY | VX3B_DOF_RZ;
        if (f & rotAll) {
            if ((f & rotAll) == rotAll) {
                s.orient = {(R)ext->rotationQ.w, (R)ext->rotationQ.x, (R)ext->rotationQ.y, (R)ext->rotationQ.z};
                s.angMom = V3<R>{R(0), R(0), R(0)};
            } else {
                V3<R> rv = rotvec(s.orient);
                if (f & VX3B_DOF_RX) {
                    rv.x = R(0);
                    s.angMom.x = R(0);
                }
                if (f & VX3B_DOF_RY) {
                    rv.y = R(0);
                    s.angMom.y = R(0);
                }
                if (f & VX3B_DOF_RZ) {
                    rv.z = R(0);
                    s.angMom.z = R(0);
                }
                s.orient = quat_from_rotvec(rv);
            }
        }
    }
    if (voxel_zero_angmom(m, wp, staticFric, s))
        s.angMom = V3<R>{R(0), R(0), R(0)};
    s.flags = staticFric ? VF_STATIC_FRICTION : 0;
}

// Temperature of a voxel at time t (VX3/VX3_VoxelyzeKernel.cu:638-647); the sine is taken in double as the reference does.
VXD float voxel_temperature(double amplitude, double period, double phase, double t, int enableExpansion) {
    double T = amplitude * sin(2 * 3.1415926f * (t / period + phase));
    if (!enableExpansion && T > 0)
        T = 0;
    return (float)T;
}

// The same temperature from per-voxel sin / cos of K*phase and per-step sin / cos of K*t/period (K = 2 * 3.1415926f): the float
// build's form in k_resident, where one thread takes the sine of the step and every voxel pays four multiply-adds instead of a
// double-precision sin() of its own.  Agrees with the expression above to ~1e-14 of the amplitude before the cast to float.
VXD float voxel_temperature_split(double amplitude, double sinA, double cosA, double sinB, double cosB, int enableExpansion) {
    double T = amplitude * (sinA * cosB + cosA * sinB);
    if (!enableExpansion && T > 0)
        T = 0;
    return (float)T;
}

// Contact between two surface voxels (VX3/VX3_VoxelyzeKernel.cu:682-727, VX3/VX3_Collision.cu:3-31).
// Returns true when (1,2) is a watched pair (inside the envelope on every axis and in distance); `force` is the force
// the reference adds to voxel 1's contactForce (voxel 2 gets the negative).  The caller has already excluded linked
// voxels and pairs whose materials are both `fixed`.
// avg1 / avg2: baseSizeAverage() of the two voxels at their current temperatures
template <class R>
VXD bool contact_pair_sized(const DVoxMat &m1, double avg1, V3<double> p1, V3<R> linMom1, const DVoxMat &m2, double avg2, V3<double> p2, V3<R> linMom2,
                            V3<R> &force) {
    double Rw = (avg1 + avg2) * 0.625;
    double dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
    force = V3<R>{R(0), R(0), R(0)};
    double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > Rw * Rw * (1.0 + 1e-12)) // certainly outside: the reference's box test and Length() > watchDistance both reject it
        return false;
    if (dx > Rw || dx < -Rw || dy > Rw || dy < -Rw || dz > Rw || dz < -Rw)
        return false;
    double dist = sqrt(d2);
    if (dist > Rw)
        return false;
    double relDist = Rw - dist;
    if (relDist > 0) {
        float stiff = 2.0f / (1.0f / m1.penetrationStiffness + 1.0f / m2.penetrationStiffness);
        float damp = 0.5f * (m1.collisionDampTranslateC + m2.collisionDampTranslateC);
        double inv = dist > 0 ? 1.0 / dist : 0.0;
        V3<R> unit{(R)(-dx * inv), (R)(-dy * inv), (R)(-dz * inv)}; // from voxel 1 towards voxel 2
        R relVel = dot(linMom1 * (R)m1.massInverse, unit) - dot(linMom2 * (R)m2.massInverse, unit);
        force = unit * ((R)stiff * (R)relDist + (R)damp * relVel);
    }
    return true;
}
template <class R>
VXD bool contact_pair(const DVoxMat &m1, float temp1, V3<double> p1, V3<R> linMom1, const DVoxMat &m2, float temp2, V3<double> p2,
                      V3<R> linMom2, V3<R> &force) {
    return contact_pair_sized<R>(m1, base_size_average(m1, temp1), p1, linMom1, m2, base_size_average(m2, temp2), p2, linMom2, force);
}

} // namespace vx3b
