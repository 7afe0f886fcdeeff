// pe_sphere_dev.cuh -- device helpers shared by the sphere kernels (pe_sphere.cu) and the batched cloud aggregation
// (pe_aggregate.cu): the candidate box of getSphereCrsFromXyz (pdb_eda/cutils.pyx:238-243), the separable squared
// distance tables of orthogonal cells and the in-sphere enumeration in the reference's order.
#pragma once
#include "pe_common.cuh"

namespace pe {

constexpr int kDMax = 64;        // widest tabulated box edge (2R+2); wider boxes use the generic path
constexpr int kSphereWarps = 4;  // warps (= atoms) per CTA

struct AxisTab {
    double sq[kDMax];  // fl((coord - atom)^2) per index along this axis
    int off[kDMax];    // wrapped element offset contribution, or kInvalidOff
};

struct AtomBox {
    int lo[3];
    int dim[3];
};

// Box of getSphereCrsFromXyz (pdb_eda/cutils.pyx:238-243) and the distance threshold.
__device__ __forceinline__ void atom_box(const pe_geom &g, double ax, double ay, double az, float radius, AtomBox &b,
                                         double &thr) {
    const double r = (double)radius;
    int c0, r0, s0, rc, rr, rs;
    xyz2crs(g, ax, ay, az, c0, r0, s0);
    xyz2crs(g, __dadd_rn(g.origin[0], r), __dadd_rn(g.origin[1], r), __dadd_rn(g.origin[2], r), rc, rr, rs);
    b.lo[0] = c0 - rc - 1;
    b.lo[1] = r0 - rr - 1;
    b.lo[2] = s0 - rs - 1;
    b.dim[0] = max(2 * rc + 2, 0);
    b.dim[1] = max(2 * rr + 2, 0);
    b.dim[2] = max(2 * rs + 2, 0);
    thr = sphere_threshold(r);
}

// Axis term of the separable squared distance for crs axis `axis`, index k (orthogonal cells only).
__device__ __forceinline__ double axis_sq(const pe_geom &g, int axis, int k, double ax, double ay, double az) {
    const int i = g.map2crs[axis];  // xyz axis carried by this crs axis
    const double coord = __dadd_rn(__dmul_rn((double)k, g.grid_length[i]), g.origin[i]);
    const double d = __dsub_rn(coord, sel3(ax, ay, az, i));
    return __dmul_rn(d, d);
}

__device__ __forceinline__ int axis_off(const pe_geom &g, int axis, int k) {
    const int w = wrap_index(k, g.ncrs[axis], g.crs_interval[axis]);
    if (w < 0) return kInvalidOff;
    return axis == 0 ? w : (axis == 1 ? w * g.ncrs[0] : w * g.ncrs[0] * g.ncrs[1]);
}

__device__ __forceinline__ void fill_tables(const pe_geom &g, const AtomBox &b, double ax, double ay, double az,
                                            AxisTab *tab /* [2]: row axis, section axis */, int lane) {
    for (int k = lane; k < b.dim[1]; k += 32) {
        tab[0].sq[k] = axis_sq(g, 1, b.lo[1] + k, ax, ay, az);
        tab[0].off[k] = axis_off(g, 1, b.lo[1] + k);
    }
    for (int k = lane; k < b.dim[2]; k += 32) {
        tab[1].sq[k] = axis_sq(g, 2, b.lo[2] + k, ax, ay, az);
        tab[1].off[k] = axis_off(g, 2, b.lo[2] + k);
    }
    __syncwarp();
}

// Calls f(ic, ir, is, value_is_valid, rho) for every in-sphere voxel of the box, lanes in parallel.
template <class F>
__device__ __forceinline__ void for_each_inside(const pe_geom &g, const float *__restrict__ rho, const AtomBox &b,
                                                double ax, double ay, double az, double T, AxisTab *tab, int lane, F f) {
    const int D0 = b.dim[0], D1 = b.dim[1], D2 = b.dim[2];
    const bool tabulated = g.orthogonal && D1 <= kDMax && D2 <= kDMax;
    if (tabulated) {
        fill_tables(g, b, ax, ay, az, tab, lane);
        const bool caseb = g.map2xyz[2] == 0;
        const int inner = caseb ? 2 : g.map2xyz[2];
        const int outer = 3 - inner;
        const AxisTab &ti = tab[inner - 1];
        const AxisTab &to = tab[outer - 1];
        const int Di = sel3(b.dim[0], b.dim[1], b.dim[2], inner), Do = sel3(b.dim[0], b.dim[1], b.dim[2], outer);
        int d0p = 1;
        while (d0p < D0 && d0p < 32) d0p <<= 1;
        const int rpi = 32 / d0p, lrow = lane / d0p, lc = lane % d0p;
        for (int cbase = 0; cbase < D0; cbase += 32) {
            const int ic = cbase + lc;
            if (ic >= D0) continue;
            const int c = b.lo[0] + ic;
            const double sqc = axis_sq(g, 0, c, ax, ay, az);
            const int offc = axis_off(g, 0, c);
            for (int ko = lrow; ko < Do; ko += rpi) {
                const double sqo = to.sq[ko];
                const int offo = to.off[ko];
                const double P = __dadd_rn(sqc, sqo);
                for (int ki = 0; ki < Di; ++ki) {
                    const double sqi = ti.sq[ki];
                    const int offi = ti.off[ki];
                    const double d2 = caseb ? __dadd_rn(__dadd_rn(sqo, sqi), sqc) : __dadd_rn(P, sqi);
                    if (!(d2 <= T)) continue;
                    const bool ok = (offc | offo | offi) >= 0;
                    const float v = ok ? __ldg(rho + (offc + offo + offi)) : 0.f;
                    const int ir = inner == 1 ? ki : ko, is = inner == 2 ? ki : ko;
                    f(ic, ir, is, ok, v);
                }
            }
        }
        __syncwarp();
    } else {
        const int64_t vol = (int64_t)D0 * D1 * D2;
        for (int64_t m = lane; m < vol; m += 32) {
            const int ic = (int)(m % D0);
            const int64_t t = m / D0;
            const int ir = (int)(t % D1), is = (int)(t / D1);
            const int c = b.lo[0] + ic, r = b.lo[1] + ir, s = b.lo[2] + is;
            double vx, vy, vz;
            crs2xyz(g, c, r, s, vx, vy, vz);
            if (!(dist2(ax, ay, az, vx, vy, vz) <= T)) continue;
            const int oc = axis_off(g, 0, c), orr = axis_off(g, 1, r), os = axis_off(g, 2, s);
            const bool ok = (oc | orr | os) >= 0;
            const float v = ok ? __ldg(rho + (oc + orr + os)) : 0.f;
            f(ic, ir, is, ok, v);
        }
        __syncwarp();
    }
}

// Density predicate of getSphereCrsFromXyz (pdb_eda/cutils.pyx:245); all operands are exact float32 values.
__device__ __forceinline__ bool passes(float v, float cut) { return (0.f < cut && cut < v) || (v < cut && cut < 0.f) || cut == 0.f; }

}  // namespace pe
