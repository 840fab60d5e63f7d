/*
 * oracle/gfn_oracle.c - CPU restatement of the reference g-function pair path.
 *
 * TEST INFRASTRUCTURE.  This file is the *checker*, not the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product path (libgfn.so -> sm_100a kernels) never links, imports or calls anything here and
 * has no CPU fallback.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the reference's own
 * compiled module (oracle/_ref, built by oracle/build_ref.py from the unmodified
 * /root/reference/newanalysis/gfunction/gfunction.pyx) in tests/test_oracle.py, and against the
 * committed fixtures tests/golden/ that were generated from that module.  The reference itself
 * ships no tests or golden vectors for this path.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared  (no -march, no -ffast-math: the
 * reference binary contains no fused multiply-adds, and neither may this one).
 *
 * What is restated (reference file:line, all in newanalysis/gfunction/gfunction.pyx):
 *   gfn_oracle_calc_histo   <- cdef void calcHisto               :110-165
 *   gfn_oracle_add_histo    <- RDF._addHisto                     :466-480
 *   gfn_oracle_calc_histo_ts<- def calcHistoTS                   :168-192   (next-row N3)
 *   gfn_oracle_calc_histo_shells <- cdef void calcHistoVoronoi / calcHistoVoronoiNonSelf :570-664, :667-761 (next-row N2)
 *   gfn_oracle_add_histo_shells  <- RDF_voronoi._addHisto       :1023-1047
 *   gfn_oracle_calc_histo_ts_voro<- def calcHistoTSVoro          :195-224
 *   gfn_oracle_com_by_residue / gfn_oracle_dip_by_residue <- newanalysis/helpers/helpers.pyx:71-98, :101-123 (next-row N1)
 */
#include <math.h>
#include <stddef.h>

/* Per-frame pair loop, gfunction.pyx:110-165.
 *
 * dest has the reference layout [mode k = 0..6][core i][bin] flattened, 7*ncore*histo_n doubles.
 * Returns 0, or 1 when an in-range pair divides by a zero norm (the reference raises
 * ZeroDivisionError out of the prange there, Cython cdivision=False; the histogram is then left
 * partially updated - here: all pairs before the offending one, in i-major/j-ascending order of
 * the row that hit it; other rows complete).
 *
 * Quirk Q3 (upper edge inclusive): pos may equal histo_n; the reference then writes one element
 * past the row, i.e. into bin 0 of the next row / next mode.  That is reproduced as long as the
 * flat index stays inside dest; a write past the very end of dest (undefined behaviour in the
 * reference) is dropped and counted in *oob (may be NULL).
 */
int gfn_oracle_calc_histo(const double *core_xyz, const double *surr_xyz,
                          const double *core_dip, const double *surr_dip,
                          double *dest, const double *box_dim,
                          int ncore, int nsurround,
                          float histo_min, float histo_max, int histo_n,
                          double invdx, int mode_sel, long long *oob)
{
    const long long ix1 = (long long)ncore * histo_n;   /* :114  (reference uses int) */
    const long long total = 7 * ix1;
    const int need_c = (mode_sel & 2) || (mode_sel & 4) || (mode_sel & 16) || (mode_sel & 32);  /* :121 */
    const int need_s = (mode_sel & 2) || (mode_sel & 8) || (mode_sel & 16) || (mode_sel & 64);  /* :135 */
    int zerodiv = 0;
    long long n_oob = 0;
    int i;

#define PUT(K, V) do { long long ix_ = (long long)(K) * ix1 + ix2 + pos; \
                       if (ix_ < total) dest[ix_] += (V); else lost++; } while (0)

#pragma omp parallel for schedule(static) reduction(|:zerodiv) reduction(+:n_oob)
    for (i = 0; i < ncore; i++) {                                   /* :119 */
        const long long ix2 = (long long)i * histo_n;               /* :120 */
        double cpx = 0, cpy = 0, cpz = 0, clen = 0;
        double spx = 0, spy = 0, spz = 0, slen = 0;
        long long lost = 0;
        int j, j_begin, stop = 0;
        if (need_c) {                                               /* :121-125 */
            cpx = core_dip[i * 3]; cpy = core_dip[i * 3 + 1]; cpz = core_dip[i * 3 + 2];
            clen = sqrt(cpx * cpx + cpy * cpy + cpz * cpz);
        }
        j_begin = (ncore == nsurround) ? i : 0;                     /* :126-129 */
        for (j = j_begin; j < nsurround && !stop; j++) {            /* :130 */
            const double off_diag = (ncore == nsurround && i != j) ? 2.0 : 1.0;   /* :131-134 */
            double dx, dy, dz, dist, incr;
            int pos;
            if (need_s) {                                           /* :135-139 */
                spx = surr_dip[j * 3]; spy = surr_dip[j * 3 + 1]; spz = surr_dip[j * 3 + 2];
                slen = sqrt(spx * spx + spy * spy + spz * spz);
            }
            dx = surr_xyz[j * 3]     - core_xyz[i * 3];             /* :141-143 */
            dy = surr_xyz[j * 3 + 1] - core_xyz[i * 3 + 1];
            dz = surr_xyz[j * 3 + 2] - core_xyz[i * 3 + 2];
            /* :145-147  single-shift minimum image; ((0<d)-(d<0)) is an int in {-1,0,1} */
            if (fabs(dx) > box_dim[3]) dx = dx - ((0.0 < dx) - (dx < 0.0)) * box_dim[0];
            if (fabs(dy) > box_dim[4]) dy = dy - ((0.0 < dy) - (dy < 0.0)) * box_dim[1];
            if (fabs(dz) > box_dim[5]) dz = dz - ((0.0 < dz) - (dz < 0.0)) * box_dim[2];
            dist = sqrt(dx * dx + dy * dy + dz * dz);               /* :148 */
            if (dist < histo_min || dist > histo_max || floor(dist * invdx) == 0) continue;   /* :149 */
            pos = (int)floor((dist - histo_min) * invdx);           /* :150 */

            if (mode_sel & 1) PUT(0, 1.0 * off_diag);               /* :152 */
            if ((mode_sel & 2) || (mode_sel & 16)) {                /* :154-157 */
                double den = clen * slen;
                if (den == 0) { zerodiv = 1; stop = 1; continue; }
                incr = (cpx * spx + cpy * spy + cpz * spz) / den;
                if (mode_sel & 2)  PUT(1, incr * off_diag);
                if (mode_sel & 16) PUT(4, (1.5 * incr * incr - 0.5) * off_diag);
            }
            if ((mode_sel & 4) || (mode_sel & 32)) {                /* :158-161 */
                double den = clen * dist;
                if (den == 0) { zerodiv = 1; stop = 1; continue; }
                incr = (cpx * dx + cpy * dy + cpz * dz) / den;
                if (mode_sel & 4)  PUT(2, incr * off_diag);
                if (mode_sel & 32) PUT(5, (1.5 * incr * incr - 0.5) * off_diag);
            }
            if ((mode_sel & 8) || (mode_sel & 64)) {                /* :162-165 */
                double den = slen * dist;
                if (den == 0) { zerodiv = 1; stop = 1; continue; }
                incr = (spx * dx + spy * dy + spz * dz) / den;
                if (mode_sel & 8)  PUT(3, incr * off_diag);
                if (mode_sel & 64) PUT(6, (1.5 * incr * incr - 0.5) * off_diag);
            }
        }
        n_oob += lost;
    }
#undef PUT
    if (oob) *oob += n_oob;
    return zerodiv;
}

/* RDF._addHisto, gfunction.pyx:466-480: histogram_out[k][b] += sum_{i=0..n1-1} histogram[k][i][b],
 * rows summed in ascending i (deterministic, thread-count independent).  Accumulates into out. */
void gfn_oracle_add_histo(double *out, const double *histogram, int n1, int histo_n)
{
    const long long ix2 = (long long)n1 * histo_n;
    int b;
#pragma omp parallel for schedule(static)
    for (b = 0; b < histo_n; b++) {                                 /* :476 */
        int j, k;
        for (j = 0; j < n1; j++) {                                  /* :477 */
            const long long ix = (long long)histo_n * j;
            for (k = 0; k < 7; k++)                                 /* :479-480 */
                out[(long long)k * histo_n + b] += histogram[k * ix2 + ix + b];
        }
    }
}

/* calcHistoTS, gfunction.pyx:168-192: g000 time series, cubic box, result[t][i][pos] += 1.
 * result_t is the [ncore][nbins] slab of frame t inside the C-contiguous result array; `room` doubles are addressable
 * from result_t on (to the end of the whole array).  The reference indexes its memoryview without a bounds check
 * (boundscheck(False), :167), so pos >= nbins (Q3) lands at flat offset i*nbins + pos: bin 0 of the next core row, or of
 * the next frame's first row for the last core site.  Reproduced while it stays inside `room`; beyond it (where the
 * reference writes past its array) the pair is dropped and counted in the return value. */
long long gfn_oracle_calc_histo_ts_room(const double *core_xyz, const double *surr_xyz, double *result_t,
                                        int ncore, int nsurr, int nbins, double boxl,
                                        float histo_min, float histo_max, double invdx, long long room);

long long gfn_oracle_calc_histo_ts(const double *core_xyz, const double *surr_xyz, double *result_t,
                                   int ncore, int nsurr, int nbins, double boxl,
                                   float histo_min, float histo_max, double invdx)
{
    return gfn_oracle_calc_histo_ts_room(core_xyz, surr_xyz, result_t, ncore, nsurr, nbins, boxl, histo_min, histo_max,
                                         invdx, (long long)ncore * nbins);
}

long long gfn_oracle_calc_histo_ts_room(const double *core_xyz, const double *surr_xyz, double *result_t,
                                        int ncore, int nsurr, int nbins, double boxl,
                                        float histo_min, float histo_max, double invdx, long long room)
{
    const double boxl2 = boxl / 2.0;                                /* :176 */
    long long dropped = 0;
    int i;
#pragma omp parallel for schedule(static) reduction(+:dropped)
    for (i = 0; i < ncore; i++) {
        int j;
        for (j = 0; j < nsurr; j++) {
            double dx = surr_xyz[j * 3]     - core_xyz[i * 3];
            double dy = surr_xyz[j * 3 + 1] - core_xyz[i * 3 + 1];
            double dz = surr_xyz[j * 3 + 2] - core_xyz[i * 3 + 2];
            double dist;
            int pos;
            if (fabs(dx) > boxl2) dx = dx - ((0.0 < dx) - (dx < 0.0)) * boxl;
            if (fabs(dy) > boxl2) dy = dy - ((0.0 < dy) - (dy < 0.0)) * boxl;
            if (fabs(dz) > boxl2) dz = dz - ((0.0 < dz) - (dz < 0.0)) * boxl;
            dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < histo_min || dist > histo_max || floor(dist * invdx) == 0) continue;
            pos = (int)floor((dist - histo_min) * invdx);
            /* rows are thread-private except for the Q3 spill into the next row's first bins: atomic adds keep the
             * (integer) counts exact whatever the interleaving */
            if ((long long)i * nbins + pos < room) {
#pragma omp atomic
                result_t[(long long)i * nbins + pos] += 1.0;
            } else dropped++;
        }
    }
    return dropped;
}

/* Reduced-memory variant used for big parity/bench cases (cfg3+ row blocks): same per-pair
 * arithmetic as gfn_oracle_calc_histo, but every thread owns a private 7 x histo_n histogram and
 * rows [row_begin,row_end) are accumulated straight into it, so no 7*n1*histo_n array is needed.
 * g000 is exact (integers); weighted sums differ from the reference only in summation order.
 * Overflow (pos >= histo_n) is dropped and counted (returned via *oob), never spilled.
 * Returns 1 on zero-norm division like the function above. */
int gfn_oracle_calc_histo_rows(const double *core_xyz, const double *surr_xyz,
                               const double *core_dip, const double *surr_dip,
                               double *out7, const double *box_dim,
                               int ncore, int nsurround, int row_begin, int row_end,
                               float histo_min, float histo_max, int histo_n,
                               double invdx, int mode_sel, long long *oob)
{
    const int need_c = (mode_sel & 2) || (mode_sel & 4) || (mode_sel & 16) || (mode_sel & 32);
    const int need_s = (mode_sel & 2) || (mode_sel & 8) || (mode_sel & 16) || (mode_sel & 64);
    int zerodiv = 0;
    long long n_oob = 0;
#pragma omp parallel reduction(|:zerodiv) reduction(+:n_oob)
    {
        double local[7 * 4096];
        double *h = local;
        int b, i;
        if (histo_n > 4096) { zerodiv = 2; }
        else {
        for (b = 0; b < 7 * histo_n; b++) h[b] = 0.0;
#pragma omp for schedule(dynamic, 16)
        for (i = row_begin; i < row_end; i++) {
            double cpx = 0, cpy = 0, cpz = 0, clen = 0, spx = 0, spy = 0, spz = 0, slen = 0;
            int j, j_begin;
            if (need_c) {
                cpx = core_dip[i * 3]; cpy = core_dip[i * 3 + 1]; cpz = core_dip[i * 3 + 2];
                clen = sqrt(cpx * cpx + cpy * cpy + cpz * cpz);
            }
            j_begin = (ncore == nsurround) ? i : 0;
            for (j = j_begin; j < nsurround; j++) {
                const double off_diag = (ncore == nsurround && i != j) ? 2.0 : 1.0;
                double dx, dy, dz, dist, incr, den;
                int pos;
                dx = surr_xyz[j * 3]     - core_xyz[i * 3];
                dy = surr_xyz[j * 3 + 1] - core_xyz[i * 3 + 1];
                dz = surr_xyz[j * 3 + 2] - core_xyz[i * 3 + 2];
                if (fabs(dx) > box_dim[3]) dx = dx - ((0.0 < dx) - (dx < 0.0)) * box_dim[0];
                if (fabs(dy) > box_dim[4]) dy = dy - ((0.0 < dy) - (dy < 0.0)) * box_dim[1];
                if (fabs(dz) > box_dim[5]) dz = dz - ((0.0 < dz) - (dz < 0.0)) * box_dim[2];
                dist = sqrt(dx * dx + dy * dy + dz * dz);
                if (dist < histo_min || dist > histo_max || floor(dist * invdx) == 0) continue;
                pos = (int)floor((dist - histo_min) * invdx);
                if (pos >= histo_n) { n_oob++; continue; }
                if (need_s) {
                    spx = surr_dip[j * 3]; spy = surr_dip[j * 3 + 1]; spz = surr_dip[j * 3 + 2];
                    slen = sqrt(spx * spx + spy * spy + spz * spz);
                }
                if (mode_sel & 1) h[pos] += 1.0 * off_diag;
                if ((mode_sel & 2) || (mode_sel & 16)) {
                    den = clen * slen;
                    if (den == 0) { zerodiv = 1; continue; }
                    incr = (cpx * spx + cpy * spy + cpz * spz) / den;
                    if (mode_sel & 2)  h[histo_n + pos] += incr * off_diag;
                    if (mode_sel & 16) h[4 * histo_n + pos] += (1.5 * incr * incr - 0.5) * off_diag;
                }
                if ((mode_sel & 4) || (mode_sel & 32)) {
                    den = clen * dist;
                    if (den == 0) { zerodiv = 1; continue; }
                    incr = (cpx * dx + cpy * dy + cpz * dz) / den;
                    if (mode_sel & 4)  h[2 * histo_n + pos] += incr * off_diag;
                    if (mode_sel & 32) h[5 * histo_n + pos] += (1.5 * incr * incr - 0.5) * off_diag;
                }
                if ((mode_sel & 8) || (mode_sel & 64)) {
                    den = slen * dist;
                    if (den == 0) { zerodiv = 1; continue; }
                    incr = (spx * dx + spy * dy + spz * dz) / den;
                    if (mode_sel & 8)  h[3 * histo_n + pos] += incr * off_diag;
                    if (mode_sel & 64) h[6 * histo_n + pos] += (1.5 * incr * incr - 0.5) * off_diag;
                }
            }
        }
#pragma omp critical
        for (b = 0; b < 7 * histo_n; b++) out7[b] += h[b];
        }
    }
    if (oob) *oob += n_oob;
    return zerodiv;
}


/* comByResidue, newanalysis/helpers/helpers.pyx:71-98 (velcomByResidue :41-68 is the same arithmetic on velocities):
 * com[i][k] = (sum_j coor[rfa[i]+j][k] * masses[rfa[i]+j]) / (sum_j masses[rfa[i]+j]), j ascending. */
void gfn_oracle_com_by_residue(const double *coor, const double *masses, int nres, const int *apr, const int *rfa, double *com)
{
    int i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < nres; i++) {
        double tot = 0.0, s[3] = {0.0, 0.0, 0.0};
        int j, k;
        for (j = 0; j < apr[i]; j++) {
            const int a = rfa[i] + j;
            for (k = 0; k < 3; k++) s[k] += coor[3 * (long long)a + k] * masses[a];
            tot += masses[a];
        }
        for (k = 0; k < 3; k++) com[3 * (long long)i + k] = s[k] / tot;
    }
}

/* dipByResidue, helpers.pyx:101-123: dip[i][k] = sum_j (coor[a][k] - com[i][k]) * charges[a]. */
void gfn_oracle_dip_by_residue(const double *coor, const double *charges, int nres, const int *apr, const int *rfa,
                               const double *com, double *dip)
{
    int i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < nres; i++) {
        double s[3] = {0.0, 0.0, 0.0};
        int j, k;
        for (j = 0; j < apr[i]; j++) {
            const int a = rfa[i] + j;
            for (k = 0; k < 3; k++) s[k] += (coor[3 * (long long)a + k] - com[3 * (long long)i + k]) * charges[a];
        }
        for (k = 0; k < 3; k++) dip[3 * (long long)i + k] = s[k];
    }
}


/* ---- N2: shell-resolved pair loop ------------------------------------------------------------------------
 * calcHistoVoronoi (gfunction.pyx:570-664) and calcHistoVoronoiNonSelf (:667-761, exclude_self != 0).
 * Same pair arithmetic as calcHisto; every in-range pair also looks up the Delaunay shell of its two
 * molecules, shell = delaunay[(i / apr_core) * nsurround + j / apr_surround] - 1 with values outside
 * [0, nshells) sent to the last shell (:601-605; the row stride IS nsurround in the reference), and the
 * row layout becomes [bin][shell]: flat index k*nshells*ncore*histo_n + i*nshells*histo_n + pos*nshells + shell
 * (:574, :589, :648-651).  delaunay_elems bounds the lookup (the reference does not check); an index past
 * it returns 2 (nothing written).  Returns 1 on a zero-norm division like gfn_oracle_calc_histo. */
int gfn_oracle_calc_histo_shells(const double *core_xyz, const double *surr_xyz,
                                 const double *core_dip, const double *surr_dip,
                                 double *dest, const double *box_dim,
                                 int ncore, int nsurround,
                                 float histo_min, float histo_max, int histo_n,
                                 double invdx, int mode_sel, int nm_core, int nm_surround, int nshells,
                                 const int *delaunay, long long delaunay_elems, int exclude_self, long long *oob)
{
    const long long row = (long long)nshells * histo_n;
    const long long ix1 = row * ncore;                                  /* :574 */
    const long long total = 7 * ix1;
    const int apr_core = ncore / nm_core, apr_surround = nsurround / nm_surround;   /* :582-583 */
    const int need_c = (mode_sel & 2) || (mode_sel & 4) || (mode_sel & 16) || (mode_sel & 32);
    const int need_s = (mode_sel & 2) || (mode_sel & 8) || (mode_sel & 16) || (mode_sel & 64);
    int status = 0;
    long long n_oob = 0;
    int i;
    if (apr_core < 1 || apr_surround < 1) return 2;
    if ((long long)((ncore - 1) / apr_core) * nsurround + (nsurround - 1) / apr_surround >= delaunay_elems) return 2;

#pragma omp parallel for schedule(static) reduction(|:status) reduction(+:n_oob)
    for (i = 0; i < ncore; i++) {                                       /* :588 */
        const long long ix2 = (long long)i * row;                       /* :590 */
        const int mol_i = i / apr_core;                                 /* :591 */
        double cpx = 0, cpy = 0, cpz = 0, clen = 0, spx = 0, spy = 0, spz = 0, slen = 0;
        long long lost = 0;
        int j, stop = 0;
        if (need_c) {                                                   /* :593-597 */
            cpx = core_dip[i * 3]; cpy = core_dip[i * 3 + 1]; cpz = core_dip[i * 3 + 2];
            clen = sqrt(cpx * cpx + cpy * cpy + cpz * cpz);
        }
        for (j = (ncore == nsurround) ? i : 0; j < nsurround && !stop; j++) {   /* :598-602 */
            const int mol_j = j / apr_surround;                         /* :601 */
            const double w = (ncore == nsurround && i != j) ? 2.0 : 1.0;        /* :607-610 */
            double dx, dy, dz, dist, c;
            long long at;
            int shell, pos, k;
            if (exclude_self && mol_i == mol_j) continue;               /* :700 */
            shell = delaunay[(long long)mol_i * nsurround + mol_j] - 1; /* :602-603 */
            if (shell < 0 || shell >= nshells) shell = nshells - 1;     /* :604-605 */
            if (need_s) {                                               /* :611-615 */
                spx = surr_dip[j * 3]; spy = surr_dip[j * 3 + 1]; spz = surr_dip[j * 3 + 2];
                slen = sqrt(spx * spx + spy * spy + spz * spz);
            }
            dx = surr_xyz[j * 3]     - core_xyz[i * 3];                 /* :617-619 */
            dy = surr_xyz[j * 3 + 1] - core_xyz[i * 3 + 1];
            dz = surr_xyz[j * 3 + 2] - core_xyz[i * 3 + 2];
            if (fabs(dx) > box_dim[3]) dx = dx - ((0.0 < dx) - (dx < 0.0)) * box_dim[0];   /* :621-623 */
            if (fabs(dy) > box_dim[4]) dy = dy - ((0.0 < dy) - (dy < 0.0)) * box_dim[1];
            if (fabs(dz) > box_dim[5]) dz = dz - ((0.0 < dz) - (dz < 0.0)) * box_dim[2];
            dist = sqrt(dx * dx + dy * dy + dz * dz);                   /* :624 */
            if (dist < histo_min || dist > histo_max || floor(dist * invdx) == 0) continue;   /* :625 */
            pos = (int)floor((dist - histo_min) * invdx);               /* :626 */
            at = ix2 + (long long)pos * nshells + shell;                /* :648-649 */
            for (k = 0; k < 3 && !stop; k++) {
                /* k = 0: core.surr dipoles (:653-656), 1: core dipole.distance (:657-660), 2: surr dipole.distance (:661-664) */
                const int lin = k == 0 ? 2 : (k == 1 ? 4 : 8), quad = lin << 3;
                double den, num;
                if (k == 0 && (mode_sel & 1)) {                         /* :651 */
                    if (at < total) dest[at] += 1.0 * w; else lost++;
                }
                if (!(mode_sel & lin) && !(mode_sel & quad)) continue;
                if (k == 0)      { den = clen * slen; num = cpx * spx + cpy * spy + cpz * spz; }
                else if (k == 1) { den = clen * dist; num = cpx * dx + cpy * dy + cpz * dz; }
                else             { den = slen * dist; num = spx * dx + spy * dy + spz * dz; }
                if (den == 0) { status = 1; stop = 1; break; }
                c = num / den;
                if (mode_sel & lin)  { long long a = (long long)(k + 1) * ix1 + at; if (a < total) dest[a] += c * w; else lost++; }
                if (mode_sel & quad) { long long a = (long long)(k + 4) * ix1 + at; if (a < total) dest[a] += (1.5 * c * c - 0.5) * w; else lost++; }
            }
        }
        n_oob += lost;
    }
    if (oob) *oob += n_oob;
    return status;
}

/* RDF_voronoi._addHisto, gfunction.pyx:1023-1047: out[k][bin][shell] += sum_i histogram[k][i][bin][shell],
 * rows in ascending i.  Accumulates into out. */
void gfn_oracle_add_histo_shells(double *out, const double *histogram, int n1, int histo_n, int nshells)
{
    const long long row = (long long)nshells * histo_n, ix2 = row * n1;
    long long e;
#pragma omp parallel for schedule(static)
    for (e = 0; e < row; e++) {                                         /* :1043-1044: (bin, shell) pairs */
        int j, k;
        for (j = 0; j < n1; j++)                                        /* :1045 */
            for (k = 0; k < 7; k++)                                     /* :1047 */
                out[k * row + e] += histogram[k * ix2 + row * j + e];
    }
}

/* calcHistoTSVoro, gfunction.pyx:195-224: result[t][i][shell-1][pos] += 1 with shell = min(ds[t][i][j], maxshell);
 * result_t is the [ncore][maxshell][nbins] slab of frame t, ds_t the [ncore][nsurr] char matrix of frame t.
 * A negative first index wraps once (Cython wraparound=True, boundscheck=False); pos >= nbins lands where
 * the flat index points (next shell / next core row) and is dropped past the end of the slab.
 * Returns the number of dropped pairs. */
long long gfn_oracle_calc_histo_ts_voro(const double *core_xyz, const double *surr_xyz, double *result_t,
                                        const signed char *ds_t, int maxshell,
                                        int ncore, int nsurr, int nbins, double boxl,
                                        float histo_min, float histo_max, double invdx)
{
    const double boxl2 = boxl / 2.0;                                    /* :203 */
    const long long total = (long long)ncore * maxshell * nbins;
    long long dropped = 0;
    int i;
#pragma omp parallel for schedule(static) reduction(+:dropped)
    for (i = 0; i < ncore; i++) {
        int j;
        for (j = 0; j < nsurr; j++) {
            double dx = surr_xyz[j * 3]     - core_xyz[i * 3];
            double dy = surr_xyz[j * 3 + 1] - core_xyz[i * 3 + 1];
            double dz = surr_xyz[j * 3 + 2] - core_xyz[i * 3 + 2];
            double dist;
            long long at;
            int pos, shell;
            if (fabs(dx) > boxl2) dx = dx - ((0.0 < dx) - (dx < 0.0)) * boxl;
            if (fabs(dy) > boxl2) dy = dy - ((0.0 < dy) - (dy < 0.0)) * boxl;
            if (fabs(dz) > boxl2) dz = dz - ((0.0 < dz) - (dz < 0.0)) * boxl;
            dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (dist < histo_min || dist > histo_max || floor(dist * invdx) == 0) continue;
            pos = (int)floor((dist - histo_min) * invdx);
            shell = ds_t[(long long)i * nsurr + j];                     /* :219 */
            if (shell > maxshell) shell = maxshell;                     /* :220-221 */
            shell -= 1;
            if (shell < 0) shell += maxshell;                           /* wraparound of index shell-1 */
            if (shell < 0) { dropped++; continue; }
            at = ((long long)i * maxshell + shell) * nbins + pos;       /* :224 */
            if (at < total) result_t[at] += 1.0; else dropped++;
        }
    }
    return dropped;
}
