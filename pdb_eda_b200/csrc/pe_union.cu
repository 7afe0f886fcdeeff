// pe_union.cu -- set-union sphere sums, one WARP per group of atoms (the grouped form of pe_sphere_sums).
//
// Replaces getSphereCrsFromXyzList (pdb_eda/cutils.pyx:250-271) + the sums of calculateRegionDensity / calculateRegionDiscrepancy
// (pdb_eda/densityAnalysis.py:1037-1068, :1160-1211) for every residue (group) of a structure at once.
//
// Round 1's kernel (sphere_union_kernel, pe_sphere.cu) gave a group to a CTA of four warps and separated table building,
// membership and gather by block barriers; its phase split on C2 was 16 % prologue, 37 % membership, 41 % gather, 7 %
// epilogue (profiles/r02_union_phases.md) with the warps of a block waiting on each other at every barrier.  Here a warp owns
// a group from start to finish -- boxes, bounding box, per-atom tables, membership bitmap, compaction, gather, reduction all
// stay inside the warp (shuffles and __syncwarp only), groups are handed out through one atomic counter, and the four warps
// of a CTA only share the SM:
//   * tile = 32 columns x 32 rows x 32 sections of the group's bounding box (a residue at the default 3.5 A radius on a 0.5 A
//     grid is one tile), one 32-bit bitmap word per (row, section): 4 KB per warp;
//   * atoms are taken one after the other: the warp tabulates the atom's squares along the three axes (one entry per lane),
//     then every lane takes box rows (row, section) of that atom, finds the row's in-sphere columns exactly (row_chord) and ORs
//     the run into its word -- no atomics: within an atom every word belongs to one lane;
//   * gather: the bitmap rows are taken 32 at a time (one per lane); a row's bits are almost always ONE run, which is written
//     to the warp's list of map offsets with a counted loop (runs with holes fall back to bit scanning); the list is then
//     walked densely with 8 independent loads in flight per lane, each voxel of the union read exactly once.
// Exactness, summation order (fixed) and outputs are those of the round-1 kernel; skewed cells and groups whose tile would
// need more than 32 columns per bitmap word are handled by splitting the bounding box into more tiles (columns included).
#include <limits.h>
#include "pe_sphere_dev.cuh"

#ifndef PE_UNION_SWEEP
#define PE_UNION_SWEEP 0
#endif
#ifndef PE_UNION_PHASE_CYCLES
#define PE_UNION_PHASE_CYCLES 0
#endif

namespace pe {

// Diagnostic (compiled in with -DPE_UNION_PHASE_CYCLES=1 only): warp cycles by phase, summed over all warps since the last
// read through pe_sphere_union_cycles(): 0 boxes + bounding box, 1 tile offsets + per-atom tables, 2 membership, 3 gather,
// 4 reduction + output.
__device__ unsigned long long g_union_warp_cycles[5];
#if PE_UNION_PHASE_CYCLES
#define PHASE_MARK(k)                              \
    do {                                           \
        const long long now__ = clock64();         \
        t_phase[k] += now__ - t_mark;              \
        t_mark = now__;                            \
    } while (0)
#else
#define PHASE_MARK(k) do { } while (0)
#endif

constexpr int kUW = 4;            // warps per CTA
constexpr int kUT = 32;           // tile edge
#ifndef PE_UNION_CTAS
#define PE_UNION_CTAS 7
#endif
constexpr int kUCtas = PE_UNION_CTAS;                                   // resident CTAs per SM the kernel is shaped for
constexpr int kUList = kUCtas >= 9 ? 256 : (kUCtas >= 8 ? 448 : 704);    // list entries per warp
#ifndef PE_UNION_LOADS
#define PE_UNION_LOADS 8
#endif
constexpr int kULoads = PE_UNION_LOADS;  // loads in flight per lane (4: 165 us, 8: 144 us, 12 / 16: 149-153 us on C2's region pass)

struct WarpUnion {
    uint32_t bits[kUT * kUT];     // [row][section]
    int list[kUList];
    double sqC[kUT], sqR[kUT], sqS[kUT];
    int offC[kUT], offR[kUT], offS[kUT];
};

// Membership of one atom in the warp's tile, exactly, without square roots: a lane takes one box row of the atom and one
// direction along the section axis, starting at the section nearest the atom (isc) and walking away from it.  Along that walk
// the squared distance of every column only grows (each rounding step of fl(fl(X2 + Y2) + Z2) is monotone in the section
// term), and inside a (row, section) line the in-sphere columns are one interval around the column nearest the atom (km), so
// the interval [kl, kh] can only SHRINK from one section to the next: the lane keeps it and re-tests just its two ends with
// the reference's float64 predicate (row_pred), stepping an end inwards while it is outside.  About two tests per line
// instead of a chord guess with four (round 1) -- and no floating-point guess at all.
template <int MODE>
__device__ __forceinline__ void v2_mark_atom(WarpUnion &w, int lane, int cl, int nC, int rl, int nR, int sl, int nS, double T, int kmin,
                                             int smin) {
    const double *sqc = w.sqC;
    int km = kmin;  // the column nearest the atom: the box's centre column or, after rounding, one of its neighbours
    if (km > 0 && sqc[km - 1] < sqc[km]) --km;
    else if (km + 1 < nC && sqc[km + 1] < sqc[km]) ++km;
    int isc = smin;  // likewise the section nearest the atom
    if (isc > 0 && w.sqS[isc - 1] < w.sqS[isc]) --isc;
    else if (isc + 1 < nS && w.sqS[isc + 1] < w.sqS[isc]) ++isc;
    for (int item = lane; item < 2 * nR; item += 32) {
        const int ir = item >> 1, dir = item & 1;          // dir 0: isc, isc + 1, ...; dir 1: isc - 1, isc - 2, ...
        const double sr = w.sqR[ir];
        int kl = 0, kh = nC - 1;
        uint32_t *words = w.bits + (rl + ir) * kUT + sl;
        for (int is = dir ? isc - 1 : isc; dir ? is >= 0 : is < nS; is += dir ? -1 : 1) {
            const double ss = w.sqS[is];
            const double A = (MODE == 0) ? ss : ((MODE == 1) ? sr : __dadd_rn(sr, ss));
            const double B = (MODE == 0) ? sr : ss;
            if (!row_pred<MODE>(sqc, km, A, B, T)) break;  // the line misses the sphere, and so do all lines further out
            while (!row_pred<MODE>(sqc, kl, A, B, T)) ++kl;  // stops at km at the latest
            while (!row_pred<MODE>(sqc, kh, A, B, T)) --kh;
            const int count = kh - kl + 1;
            const uint32_t run = count >= 32 ? ~0u : ((1u << count) - 1u);
            words[is] |= run << (cl + kl);  // within an atom this word belongs to this lane alone
        }
    }
}

// The same with round 1's chord guess (row_chord): every (row, section) line on its own -- more instructions per line than
// the sweep, but the lines are independent (no chain from one section to the next).  Selected with -DPE_UNION_SWEEP=0.
template <int MODE>
__device__ __forceinline__ void v2_mark_atom_chord(WarpUnion &w, int lane, int cl, int nC, int rl, int nR, int sl, int nS, double T, float xc,
                                                   float inv_gl, int kmin) {
    const double *sqc = w.sqC;
    int km = kmin;
    if (km > 0 && sqc[km - 1] < sqc[km]) --km;
    else if (km + 1 < nC && sqc[km + 1] < sqc[km]) ++km;
    const int shift = nS > 1 ? 32 - __clz(nS - 1) : 0;
    const int total = nR << shift;
#pragma unroll 2
    for (int li = lane; li < total; li += 32) {
        const int is = li & ((1 << shift) - 1), ir = li >> shift;
        if (is >= nS) continue;
        const double sr = w.sqR[ir], ss = w.sqS[is];
        const double A = (MODE == 0) ? ss : ((MODE == 1) ? sr : __dadd_rn(sr, ss));
        const double B = (MODE == 0) ? sr : ss;
        int kl, kh;
        if (!row_chord<MODE>(sqc, nC, km, xc, inv_gl, A, B, T, kl, kh)) continue;  // the row misses the sphere
        const int count = kh - kl + 1;
        const uint32_t run = count >= 32 ? ~0u : ((1u << count) - 1u);
        w.bits[(rl + ir) * kUT + (sl + is)] |= run << (cl + kl);  // within an atom this word belongs to this lane alone
    }
}

// Sums of one step of the gather: kULoads values per lane; the float64 additions form a small tree (fixed shape:
// deterministic) so that no accumulator carries a chain of eight dependent additions.
template <bool HASNEG, int N>
__device__ __forceinline__ void step_sum(const float *v, SphereAcc &acc, float cp, float cn) {
    double d[N], p[N], q[N];
#pragma unroll
    for (int u = 0; u < N; ++u) {
        d[u] = widen(v[u]);
        const bool pos = v[u] > cp;
        p[u] = pos ? d[u] : 0.0;
        acc.n_pos += pos ? 1 : 0;
        if (HASNEG) {
            const bool neg = v[u] < cn;
            q[u] = neg ? d[u] : 0.0;
            acc.n_neg += neg ? 1 : 0;
        }
    }
#pragma unroll
    for (int o = 1; o < N; o <<= 1) {
#pragma unroll
        for (int u = 0; u + o < N; u += 2 * o) {
            d[u] += d[u + o];
            p[u] += p[u + o];
            if (HASNEG) q[u] += q[u + o];
        }
    }
    acc.s_all += d[0];
    acc.s_pos += p[0];
    if (HASNEG) acc.s_neg += q[0];
}

// Gather of one tile: every voxel of the union is read once.  The bitmap rows are taken one tile row (= up to 32 sections, one
// per lane) at a time; a warp prefix sum of the popcounts places every row's voxels in a list of map offsets (a row's bits are
// almost always ONE run, written with a counted loop; rows with holes are bit-scanned), and the list is walked densely, 8
// (or, for short lists, 4) independent loads per lane and step.
// Measured and dropped (profiles/r02_union_phases.md): keeping a batch's loads in flight across the compaction of the next one,
// in registers (146.3 us) or as cp.async copies into the list itself (147.9 us) against 144.4 us for this plain form; 6 / 7 / 8 / 9
// CTAs per SM 165 / 144 / 146-148 (spills at 64 registers) / 174 us.  No single pipe limits the kernel (issue slots 64 % busy; ALU
// 38 %, LSU 32 %, XU 20 %, FP64 12 %): it is the length of each warp's dependent instruction chains at 28 warps per SM.
template <bool CHECKED, bool HASNEG>
__device__ __forceinline__ void v2_gather(WarpUnion &w, const float *__restrict__ rho, SphereAcc &acc, float cp, float cn, int lane,
                                          int tR, int tS) {
    int *list = w.list;
    const int oc0 = w.offC[0];
    for (int r0 = 0; r0 < tR; ++r0) {  // one tile row = tS (<= 32) words = one batch
        uint32_t wd = lane < tS ? w.bits[r0 * kUT + lane] : 0u;
        if (!__any_sync(kFull, wd != 0u)) continue;
        if (wd != 0u) w.bits[r0 * kUT + lane] = 0u;  // leave the bitmap clear
        int orr = 0, osum = 0;
        if (lane < tS) {
            const int o1 = w.offR[r0], o2 = w.offS[lane];
            orr = o1 | o2;
            osum = (int)((unsigned)o1 + (unsigned)o2);
        }
        const int c = __popc(wd);
        bool pending = c > 0;
        while (true) {  // one round unless the batch overflows the list
            const int cc = pending ? c : 0;
            const int excl = warp_excl_scan(cc, lane);
            const bool fits = excl + cc <= kUList;
            const int total = __reduce_max_sync(kFull, fits ? excl + cc : 0);
            if (pending && fits) {
                int *dst = list + excl;
                if (CHECKED) {
                    uint32_t x = wd;
                    while (x) {
                        const int bcol = __ffs((int)x) - 1;
                        x &= x - 1u;
                        const int oc = w.offC[bcol];
                        *dst++ = ((orr | oc) < 0) ? -1 : (int)((unsigned)osum + (unsigned)oc);
                    }
                } else {
                    const int first = __ffs((int)wd) - 1;
                    const int rowbase = (int)((unsigned)osum + (unsigned)oc0);
                    if (((wd >> first) + 1u) & (wd >> first)) {  // holes: more than one run
                        uint32_t x = wd;
                        while (x) {
                            const int bcol = __ffs((int)x) - 1;
                            x &= x - 1u;
                            *dst++ = rowbase + bcol;
                        }
                    } else {  // one run of c columns starting at `first`
                        const int v0 = rowbase + first;
#pragma unroll 4
                        for (int j = 0; j < c; ++j) dst[j] = v0 + j;
                    }
                }
                pending = false;
            }
            __syncwarp();
            if (lane == 0) acc.n_all += total;
            for (int i0 = 0; i0 < total; i0 += 32 * kULoads) {
                const bool half = total - i0 <= 32 * (kULoads / 2);  // warp-uniform: short lists issue half the loads
                float v[kULoads];
#pragma unroll
                for (int u = 0; u < kULoads; ++u) {
                    v[u] = 0.f;
                    if (u >= kULoads / 2 && half) continue;
                    const int idx = i0 + 32 * u + lane;
                    const int e = idx < total ? list[idx] : -2;
                    if (e >= 0) v[u] = __ldg(rho + e);
                    if (CHECKED) acc.bad |= (e == -1) ? 1 : 0;
                }
                if (half)
                    step_sum<HASNEG, kULoads / 2>(v, acc, cp, cn);
                else
                    step_sum<HASNEG, kULoads>(v, acc, cp, cn);
            }
            __syncwarp();  // the list is rewritten by the next round / batch
            if (!__any_sync(kFull, pending)) break;
        }
    }
}

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(kFull, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(kUW * 32, kUCtas)
    sphere_union_warp_kernel(const __grid_constant__ pe_geom g, const float *__restrict__ rho, int n_groups,
                             const int32_t *__restrict__ group_start, const double *__restrict__ xyz, const float *__restrict__ radius,
                             int32_t *box, double *thr, float cp, float cn, int *__restrict__ counter,
                             double *__restrict__ out /* n_groups x PE_SPHERE_NOUT */) {
    __shared__ WarpUnion wsh[kUW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpUnion &w = wsh[warp];
    cp = eff_pos(cp);
    cn = eff_neg(cn);
    const bool has_neg = cn > __int_as_float(0xff800000);
    const int icx = g.map2crs[0];  // xyz axis carried by the columns
    const float inv_gl = (float)(1.0 / g.grid_length[icx]);
    for (int i = lane; i < kUT * kUT; i += 32) w.bits[i] = 0u;  // cleared once: the gather leaves it clear
    __syncwarp();
#if PE_UNION_PHASE_CYCLES
    long long t_mark = clock64();
    long long t_phase[5] = {0, 0, 0, 0, 0};
#endif
    for (;;) {
        int grp = 0;
        if (lane == 0) grp = atomicAdd(counter, 1);
        grp = __shfl_sync(kFull, grp, 0);
        if (grp >= n_groups) break;
        const int a0 = group_start[grp], a1 = group_start[grp + 1];
        // boxes and thresholds of the group's atoms (lanes over atoms), bounding box by warp reductions
        int ulo0 = INT_MAX, ulo1 = INT_MAX, ulo2 = INT_MAX, uhi0 = INT_MIN, uhi1 = INT_MIN, uhi2 = INT_MIN;
        double candidates = 0.0;
        for (int a = a0 + lane; a < a1; a += 32) {
            AtomBox bb;
            double tt;
            atom_box(g, xyz[3 * a], xyz[3 * a + 1], xyz[3 * a + 2], radius[a], bb, tt);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                box[6 * a + k] = bb.lo[k];
                box[6 * a + 3 + k] = bb.dim[k];
            }
            thr[a] = tt;
            candidates += (double)bb.dim[0] * (double)bb.dim[1] * (double)bb.dim[2];
            if (bb.dim[0] > 0 && bb.dim[1] > 0 && bb.dim[2] > 0) {
                ulo0 = min(ulo0, bb.lo[0]);
                ulo1 = min(ulo1, bb.lo[1]);
                ulo2 = min(ulo2, bb.lo[2]);
                uhi0 = max(uhi0, bb.lo[0] + bb.dim[0]);
                uhi1 = max(uhi1, bb.lo[1] + bb.dim[1]);
                uhi2 = max(uhi2, bb.lo[2] + bb.dim[2]);
            }
        }
        __syncwarp();  // boxes visible to the whole warp
        candidates = warp_sum(candidates);
        ulo0 = warp_min(ulo0);
        ulo1 = warp_min(ulo1);
        ulo2 = warp_min(ulo2);
        uhi0 = warp_max(uhi0);
        uhi1 = warp_max(uhi1);
        uhi2 = warp_max(uhi2);
        SphereAcc acc;
        PHASE_MARK(0);
        if (uhi0 > ulo0) {
            for (int ts0 = ulo2; ts0 < uhi2; ts0 += kUT)
                for (int tr0 = ulo1; tr0 < uhi1; tr0 += kUT)
                    for (int tc0 = ulo0; tc0 < uhi0; tc0 += kUT) {
                        const int tC = min(kUT, uhi0 - tc0), tR = min(kUT, uhi1 - tr0), tS = min(kUT, uhi2 - ts0);
                        // wrapped element offsets of the tile's indices; is every index stored and are the columns adjacent?
                        int invalid = 0;
                        {
                            const int oc = lane < tC ? axis_off(g, 0, tc0 + lane) : 0;
                            const int orw = lane < tR ? axis_off(g, 1, tr0 + lane) : 0;
                            const int os = lane < tS ? axis_off(g, 2, ts0 + lane) : 0;
                            w.offC[lane] = oc;
                            w.offR[lane] = orw;
                            w.offS[lane] = os;
                            invalid = (oc | orw | os) < 0 ? 1 : 0;
                            const int prev = __shfl_up_sync(kFull, oc, 1);
                            if (lane > 0 && lane < tC && oc != prev + 1) invalid = 1;
                        }
                        const bool checked = __any_sync(kFull, invalid != 0);
                        // membership, atom by atom
                        for (int a = a0; a < a1; ++a) {
                            const int32_t *bx = box + 6 * a;
                            const int b0 = bx[0], b1 = bx[1], b2 = bx[2], d0 = bx[3], d1 = bx[4], d2 = bx[5];
                            if (d0 <= 0 || d1 <= 0 || d2 <= 0) continue;
                            const int cl = max(b0, tc0), ch = min(b0 + d0, tc0 + tC);
                            const int rl = max(b1, tr0), rh = min(b1 + d1, tr0 + tR);
                            const int sl = max(b2, ts0), sh = min(b2 + d2, ts0 + tS);
                            if (cl >= ch || rl >= rh || sl >= sh) continue;
                            const double ax = xyz[3 * a], ay = xyz[3 * a + 1], az = xyz[3 * a + 2];
                            const double T = thr[a];
                            if (g.orthogonal) {
                                const int nC = ch - cl, nR = rh - rl, nS = sh - sl;
                                if (lane < nC) w.sqC[lane] = axis_sq(g, 0, cl + lane, ax, ay, az);
                                if (lane < nR) w.sqR[lane] = axis_sq(g, 1, rl + lane, ax, ay, az);
                                if (lane < nS) w.sqS[lane] = axis_sq(g, 2, sl + lane, ax, ay, az);
                                __syncwarp();
                                PHASE_MARK(1);
                                // centre column of the box (range(c - R - 1, c + R + 1): c = lo + dim / 2), clamped into the tile part
                                const int kmin = min(max(b0 + d0 / 2 - cl, 0), nC - 1);
#if PE_UNION_SWEEP
                                const int smin = min(max(b2 + d2 / 2 - sl, 0), nS - 1);
                                v2_mark_atom<MODE>(w, lane, cl - tc0, nC, rl - tr0, nR, sl - ts0, nS, T, kmin, smin);
#else
                                const float xc = (float)((sel3(ax, ay, az, icx) - g.origin[icx]) / g.grid_length[icx] - (double)cl);
                                v2_mark_atom_chord<MODE>(w, lane, cl - tc0, nC, rl - tr0, nR, sl - ts0, nS, T, xc, inv_gl, kmin);
#endif
                                __syncwarp();  // tables are rewritten by the next atom; its rows may share words with this one
                                PHASE_MARK(2);
                            } else {
                                const int nc = ch - cl, n1 = rh - rl;
                                const int vol = nc * n1 * (sh - sl);
                                for (int m = lane; m < vol; m += 32) {
                                    const int ic = m % nc, t = m / nc;
                                    const int c = cl + ic, r = rl + t % n1, s = sl + t / n1;
                                    double vx, vy, vz;
                                    crs2xyz(g, c, r, s, vx, vy, vz);
                                    if (!(dist2(ax, ay, az, vx, vy, vz) <= T)) continue;
                                    atomicOr(&w.bits[(r - tr0) * kUT + (s - ts0)], 1u << (c - tc0));
                                }
                                __syncwarp();
                            }
                        }
                        // gather every voxel of the union once
                        PHASE_MARK(1);
                        if (checked) {
                            if (has_neg)
                                v2_gather<true, true>(w, rho, acc, cp, cn, lane, tR, tS);
                            else
                                v2_gather<true, false>(w, rho, acc, cp, cn, lane, tR, tS);
                        } else {
                            if (has_neg)
                                v2_gather<false, true>(w, rho, acc, cp, cn, lane, tR, tS);
                            else
                                v2_gather<false, false>(w, rho, acc, cp, cn, lane, tR, tS);
                        }
                        __syncwarp();
                        PHASE_MARK(3);
                    }
        }
        const int n_all = warp_sum(acc.n_all), n_pos = warp_sum(acc.n_pos), n_neg = warp_sum(acc.n_neg);
        const int bad = warp_sum(acc.bad);
        const double s_all = warp_sum(acc.s_all), s_pos = warp_sum(acc.s_pos), s_neg = warp_sum(acc.s_neg);
        if (lane == 0) {
            double *o = out + (int64_t)grp * PE_SPHERE_NOUT;
            o[0] = (double)n_all;
            o[1] = s_all;
            o[2] = (double)n_pos;
            o[3] = s_pos;
            o[4] = (double)n_neg;
            o[5] = s_neg;
            o[6] = bad ? 0.0 : 1.0;
            o[7] = candidates;
        }
        PHASE_MARK(4);
    }
#if PE_UNION_PHASE_CYCLES
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) atomicAdd(g_union_warp_cycles + k, (unsigned long long)t_phase[k]);
    }
#endif
}

int launch_union_warp(const pe_geom *g, const float *d_rho, int n_groups, const int32_t *d_group_start, const double *d_xyz,
                      const float *d_radius, int32_t *box, double *thr, float cut_pos, float cut_neg, int *d_counter, double *d_out,
                      cudaStream_t st) {
    const int mode = g->map2xyz[2] == 1 ? 0 : (g->map2xyz[2] == 2 ? 1 : 2);  // crs axis that carries z
    PE_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int), st));
    const int warps_needed = n_groups;
    int grid = (warps_needed + kUW - 1) / kUW;
    const int max_grid = sm_count() * kUCtas;
    if (grid > max_grid) grid = max_grid;
    if (mode == 0)
        PE_LAUNCH("sphere_union_kernel", st, sphere_union_warp_kernel<0><<<grid, kUW * 32, 0, st>>>(
            *g, d_rho, n_groups, d_group_start, d_xyz, d_radius, box, thr, cut_pos, cut_neg, d_counter, d_out));
    else if (mode == 1)
        PE_LAUNCH("sphere_union_kernel", st, sphere_union_warp_kernel<1><<<grid, kUW * 32, 0, st>>>(
            *g, d_rho, n_groups, d_group_start, d_xyz, d_radius, box, thr, cut_pos, cut_neg, d_counter, d_out));
    else
        PE_LAUNCH("sphere_union_kernel", st, sphere_union_warp_kernel<2><<<grid, kUW * 32, 0, st>>>(
            *g, d_rho, n_groups, d_group_start, d_xyz, d_radius, box, thr, cut_pos, cut_neg, d_counter, d_out));
    PE_LAUNCH_CHECK();
    return PE_OK;
}

}  // namespace pe

extern "C" int pe_sphere_union_warp_cycles(unsigned long long *out5) {
    PE_CHECK_ARG(out5 != nullptr, "pe_sphere_union_warp_cycles: null pointer");
    PE_CUDA(cudaMemcpyFromSymbol(out5, pe::g_union_warp_cycles, sizeof(unsigned long long) * 5));
    unsigned long long zero[5] = {0, 0, 0, 0, 0};
    PE_CUDA(cudaMemcpyToSymbol(pe::g_union_warp_cycles, zero, sizeof(zero)));
    return PE_OK;
}
