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
                // membership of the whole line first, then the gathers four at a time: with the load behind the distance test of
                // its own iteration a lane paid one memory round trip per in-sphere voxel (24 % of the count pass's stall samples)
                for (int kb = 0; kb < Di; kb += 32) {
                    uint32_t inside = 0u;
                    const int ke = min(Di - kb, 32);
                    for (int ki = 0; ki < ke; ++ki) {
                        const double sqi = ti.sq[kb + ki];
                        const double d2 = caseb ? __dadd_rn(__dadd_rn(sqo, sqi), sqc) : __dadd_rn(P, sqi);
                        inside |= (d2 <= T ? 1u : 0u) << ki;
                    }
                    while (inside) {
                        int kk[4];
                        float vv[4];
                        bool okk[4], has[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            has[u] = inside != 0u;
                            kk[u] = kb + __ffs((int)inside) - 1;
                            inside &= inside - 1u;
                            vv[u] = 0.f;
                            okk[u] = false;
                            if (has[u]) {
                                const int offi = ti.off[kk[u]];
                                okk[u] = (offc | offo | offi) >= 0;
                                if (okk[u]) vv[u] = __ldg(rho + (offc + offo + offi));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (!has[u]) continue;
                            const int ir = inner == 1 ? kk[u] : ko, is = inner == 2 ? kk[u] : ko;
                            f(ic, ir, is, okk[u], vv[u]);
                        }
                    }
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

// float -> double widening.  An integer-pipe bit-twiddling version was tried when F2F.F64.F32 showed up as the top
// stall of the first capture (profiles/r01a_first_path.md); once the gather loops kept four independent loads in
// flight the plain conversion (one issue slot, its latency hidden) became the faster one again: 257 -> 231 us on
// the C2 region pass.
__device__ __forceinline__ double widen(float v) { return (double)v; }

// Effective cutoffs: a class whose cutoff is 0 is switched off by an unreachable threshold.
__device__ __forceinline__ float eff_pos(float cp) { return cp > 0.f ? cp : __int_as_float(0x7f800000); }
__device__ __forceinline__ float eff_neg(float cn) { return cn < 0.f ? cn : __int_as_float(0xff800000); }

struct SphereAcc {
    int n_all = 0, n_pos = 0, n_neg = 0, bad = 0;
    double s_all = 0.0, s_pos = 0.0, s_neg = 0.0;
    // v must be 0 when the voxel is not taken; cp / cn are the effective cutoffs (cp > 0 > cn).
    __device__ __forceinline__ void add(bool take, float v, float cp, float cn) {
        const double d = widen(v);
        s_all += d;
        if (take) ++n_all;
        if (v > cp) {
            ++n_pos;
            s_pos += d;
        }
        if (v < cn) {
            ++n_neg;
            s_neg += d;
        }
    }
    // the same with the negative class switched off (cn == -inf): a third less work per voxel
    __device__ __forceinline__ void add_pos(bool take, float v, float cp) {
        const double d = widen(v);
        s_all += d;
        if (take) ++n_all;
        if (v > cp) {
            ++n_pos;
            s_pos += d;
        }
    }
};

// ------------------------------------------------------------------------------------------------ exact chords
// MODE 0: z is carried by the row axis, 1: by the section axis, 2: by the column axis.
// Squared distance of column k of a box row exactly as the reference adds it: fl(fl(X2 + Y2) + Z2).
template <int MODE>
__device__ __forceinline__ bool row_pred(const double *sqc, int k, double A, double B, double T) {
    // MODE 0/1: A = square of the non-z axis among (row, section), B = square of the z axis;
    // MODE 2  : A = fl(row square + section square), the column carries z.
    const double d2 = (MODE == 2) ? __dadd_rn(A, sqc[k]) : __dadd_rn(__dadd_rn(sqc[k], A), B);
    return d2 <= T;
}

// In-sphere columns [kl, kh] of one box row, exactly.  Along the columns of a box row the squared distance falls to the
// column nearest the atom (km) and rises again (every rounding step is monotone), so the in-sphere columns are one
// interval around that column.  The interval is GUESSED in float32 from the chord of the sphere along the row
// (half-width sqrt(T - A - B) in columns around the atom's fractional column xc) and then VERIFIED with the exact
// float64 predicate: inside at kl and kh, outside at kl - 1 and kh + 1 -- four independent tests that, by unimodality,
// prove the guess.  A guess that fails (an end within ~1e-5 columns of a grid point, or a row that only just misses the
// sphere) falls back to two binary searches with the same exact predicate.  Rows that miss the sphere leave after one
// exact test without touching the table: every rounding step is monotone and the column term is >= 0, so
// d2 >= fl(A + B) for every column of the row.  Returns false when no column of the row is inside.
// 1 / sqrt(x) for the chord guess only: the bare MUFU.RSQ (rsqrtf() wraps it in denormal scaling, six more instructions
// per box row; a denormal remainder just yields a guess that the exact tests reject)
__device__ __forceinline__ float rsqrt_guess(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int MODE>
__device__ __forceinline__ bool row_chord(const double *sqc, int nC, int km, float xc, float inv_gl, double A, double B, double T,
                                          int &kl, int &kh) {
    const double AB = (MODE == 2) ? A : __dadd_rn(A, B);
    if (!(AB <= T)) return false;  // exact: the row misses the sphere
    const float rem = (float)__dsub_rn(T, AB);
    const float h = rem > 0.f ? rem * rsqrt_guess(rem) * inv_gl : 0.f;
    const float xl = xc - h, xr = xc + h;
    const float fl_ = ceilf(xl), fr_ = floorf(xr);
    // The float32 guess is within 2e-5 columns of the real chord ends (|xc|, h <= 64 columns; rsqrt.approx 2^-22 relative), and the
    // reference's rounded float64 predicate moves an end by less than 1e-8 columns once the half chord exceeds 1e-6 columns.  So
    // when both guessed ends are more than kChordEps away from every grid point the interval is PROVEN without evaluating the
    // predicate (the exact tests below only run for the ~0.1 % of ends that fall within kChordEps of a column).
    constexpr float kChordEps = 2.5e-4f;
    if (h > 0.01f && fl_ - xl > kChordEps && xl - (fl_ - 1.f) > kChordEps && xr - fr_ > kChordEps && (fr_ + 1.f) - xr > kChordEps) {
        kl = max((int)fl_, 0);
        kh = min((int)fr_, nC - 1);
        return kl <= kh;
    }
    kl = min(max((int)fl_, 0), km);
    kh = max(min((int)fr_, nC - 1), km);
    const bool in_l = row_pred<MODE>(sqc, kl, A, B, T), in_h = row_pred<MODE>(sqc, kh, A, B, T);
    const bool out_l = kl == 0 || !row_pred<MODE>(sqc, kl - 1, A, B, T);
    const bool out_h = kh == nC - 1 || !row_pred<MODE>(sqc, kh + 1, A, B, T);
    if (in_l && in_h && out_l && out_h) return true;
    if (!row_pred<MODE>(sqc, km, A, B, T)) return false;  // the row misses the sphere
    int lo = 0, hi = km;  // smallest k in [0, km] inside
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (row_pred<MODE>(sqc, mid, A, B, T))
            hi = mid;
        else
            lo = mid + 1;
    }
    kl = lo;
    lo = km;
    hi = nC - 1;  // largest k in [km, nC) inside
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (row_pred<MODE>(sqc, mid, A, B, T))
            lo = mid;
        else
            hi = mid - 1;
    }
    kh = lo;
    return true;
}

}  // namespace pe
