/* pce_oracle.c -- CPU oracle for the microstate path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library; the
 * product (pcetk_b200/) never does and has no CPU fallback.
 *
 * Part 1 restates, in plain C and in this file's own words, the reference algorithms of
 *   extensions/csource/EnergyModel.c, StateVector.c and MCModelDefault.c (line ranges cited at every
 *   function).  It is pinned two ways (tests/test_oracle.py): against the reference's golden vectors in
 *   the tests/golden npz files, and -- in the build container -- bit for bit against oracle/_ref, the
 *   reference's own sources compiled unmodified.
 * Part 2 is the executable specification of the GPU Monte Carlo chain: the same Metropolis semantics
 *   driven by the counter-based Philox4x32-10 stream and the cached-field energy update the CUDA kernels
 *   use, operation for operation, so that per-chain counts can be compared exactly.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PCE_R    0.001987165392   /* EnergyModel.h:31 */
#define PCE_LN10 2.302585092994   /* EnergyModel.h:32 (truncated on purpose) */
#define PCE_TOO_SMALL (-500.0)    /* MCModelDefault.h:39 */

typedef struct {
    int           nsites, ninst;
    double        temperature;
    const int    *first, *last;      /* [nsites] global instance range per site */
    const int    *protons;           /* [ninst] */
    const double *gintr, *gmodel;    /* [ninst] */
    const double *w;                 /* [ninst*ninst] symmetrised, full square */
} OModel;

/* ------------------------------------------------------------------------------------------------
 * Part 1: restatement of the reference
 * ---------------------------------------------------------------------------------------------- */

/* (W_ij + W_ji)/2 -- EnergyModel_SymmetrizeInteractions, EnergyModel.c:85-87 (pCore
 * SymmetricMatrix_CopyFromReal2DArray; pinned by twosites.out Wmax = 5.954225). */
void oracle_symmetrize (int n, const double *wraw, double *wsym) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double v = 0.5 * (wraw[(size_t) i * n + j] + wraw[(size_t) j * n + i]);
            wsym[(size_t) i * n + j] = v;
            wsym[(size_t) j * n + i] = v;
        }
}

/* max |W_ij - W_ji| and the verdict -- EnergyModel_CheckInteractionsSymmetric, EnergyModel.c:78-80 */
int oracle_check_symmetric (int n, const double *wraw, double tolerance, double *max_deviation) {
    double maxd = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) {
            double d = fabs (wraw[(size_t) i * n + j] - wraw[(size_t) j * n + i]);
            if (d > maxd) maxd = d;
        }
    *max_deviation = maxd;
    return maxd <= tolerance;
}

static double mu_of (double temperature, double pH) {
    return -PCE_R * temperature * PCE_LN10 * pH;          /* EnergyModel.c:223, same evaluation order */
}

/* EnergyModel_CalculateMicrostateEnergy, EnergyModel.c:204-224: site-ordered sums, j < i pairs */
static double energy_folded (const OModel *m, const int *state, double pH) {
    double gintr = 0.0, w = 0.0;
    int    nprotons = 0;
    for (int i = 0; i < m->nsites; i++) {
        gintr    += m->gintr[state[i]];
        nprotons += m->protons[state[i]];
        const double *row = m->w + (size_t) state[i] * m->ninst;
        for (int j = 0; j < i; j++) w += row[state[j]];
    }
    return gintr - nprotons * mu_of (m->temperature, pH) + w;
}

/* EnergyModel_CalculateMicrostateEnergyUnfolded, EnergyModel.c:232-245 */
static double energy_unfolded (const OModel *m, const int *state, double pH) {
    double gmodel = 0.0;
    int    nprotons = 0;
    for (int i = 0; i < m->nsites; i++) {
        gmodel   += m->gmodel[state[i]];
        nprotons += m->protons[state[i]];
    }
    return gmodel - nprotons * mu_of (m->temperature, pH);
}

static OModel make_model (int nsites, int ninst, double temperature, const int *first, const int *last,
                          const int *protons, const double *gintr, const double *gmodel, const double *w) {
    OModel m;
    m.nsites = nsites; m.ninst = ninst; m.temperature = temperature;
    m.first = first; m.last = last; m.protons = protons; m.gintr = gintr; m.gmodel = gmodel; m.w = w;
    return m;
}

double oracle_energy (int nsites, int ninst, double temperature, const int *first, const int *last,
                      const int *protons, const double *gintr, const double *gmodel, const double *w,
                      const int *state, double pH, int unfolded) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    return unfolded ? energy_unfolded (&m, state, pH) : energy_folded (&m, state, pH);
}

/* odometer, site 0 fastest -- StateVector_Increment, StateVector.c:370-384 */
static int increment (const OModel *m, int *state) {
    for (int i = 0; i < m->nsites; i++) {
        if (state[i] < m->last[i]) { state[i]++; return 1; }
        state[i] = m->first[i];
    }
    return 0;
}

/* EnergyModel_CalculateProbabilitiesAnalytically[Unfolded], EnergyModel.c:342-371, made of
 * EnergyModel_CalculateZ (:252-281: energies, running minimum, shift, scale, exp, serial sum) and
 * EnergyModel_CalculateProbabilitiesFromZ (:318-337: second odometer pass, then scale by 1/Z).
 * Returns 0, or 1 if the Boltzmann-factor array cannot be allocated.  out_z / out_gmin may be NULL. */
int oracle_analytic (int nsites, int ninst, double temperature, const int *first, const int *last,
                     const int *protons, const double *gintr, const double *gmodel, const double *w,
                     double pH, double gzero, int unfolded, long long nstates,
                     double *out_p, double *out_z, double *out_gmin) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    double *b = (double *) malloc (sizeof (double) * (size_t) nstates);
    int    *state = (int *) malloc (sizeof (int) * (size_t) nsites);
    if (b == NULL || state == NULL) { free (b); free (state); return 1; }
    for (int i = 0; i < nsites; i++) state[i] = first[i];
    double gmin = (unfolded ? energy_unfolded (&m, state, pH) : energy_folded (&m, state, pH)) - gzero;
    for (long long s = 0; s < nstates; s++) {
        double g = (unfolded ? energy_unfolded (&m, state, pH) : energy_folded (&m, state, pH)) - gzero;
        if (g < gmin) gmin = g;
        b[s] = g;
        increment (&m, state);
    }
    const double scale = -1.0 / (PCE_R * temperature);
    double z = 0.0;
    for (long long s = 0; s < nstates; s++) b[s] += -gmin;
    for (long long s = 0; s < nstates; s++) b[s] *= scale;
    for (long long s = 0; s < nstates; s++) b[s] = exp (b[s]);
    for (long long s = 0; s < nstates; s++) z += b[s];
    if (out_p != NULL) {
        for (int k = 0; k < ninst; k++) out_p[k] = 0.0;
        for (int i = 0; i < nsites; i++) state[i] = first[i];
        for (long long s = 0; s < nstates; s++) {
            for (int i = 0; i < nsites; i++) out_p[state[i]] += b[s];
            increment (&m, state);
        }
        const double inv = 1.0 / z;
        for (int k = 0; k < ninst; k++) out_p[k] *= inv;
    }
    if (out_z != NULL) *out_z = z;
    if (out_gmin != NULL) *out_gmin = gmin;
    free (b); free (state);
    return 0;
}

/* The same quantities with COMPENSATED accumulation, for state counts where the reference's own naive serial double sum
 * (EnergyModel.c:275-279 for Z, :330-333 for the per-instance sums) carries more rounding error than the bar the GPU kernel
 * is held to (1e-12 relative; 2^26 terms accumulate ~1e-12 in a serial double sum).  Same algorithm -- every state's energy
 * by the reference's site-ordered sum (EnergyModel.c:204-224), the Gmin shift of :262-275, one exp per state -- but
 * sums are kept in long double (64-bit mantissa) and the state range is split over `nthreads` slices, each started by
 * decoding its first state index (site 0 the fastest digit, StateVector.c:370-384).  Two passes (minimum, then sums), no
 * nstates-sized array.  Returns 0. */
typedef struct {
    const OModel *m;
    long long s0, s1;
    int pass, unfolded;
    double pH, gzero, gmin, local_min;
    long double *acc;        /* [ninst + 1]: per-instance sums, then Z */
} CompSlice;

static void *compensated_slice (void *arg) {
    CompSlice *c = (CompSlice *) arg;
    const OModel *m = c->m;
    int *state = (int *) malloc (sizeof (int) * (size_t) m->nsites);
    long long x = c->s0;
    for (int i = 0; i < m->nsites; i++) { const int n = m->last[i] - m->first[i] + 1; state[i] = m->first[i] + (int) (x % n); x /= n; }
    const double scale = -1.0 / (PCE_R * m->temperature);
    for (long long s = c->s0; s < c->s1; s++) {
        const double g = (c->unfolded ? energy_unfolded (m, state, c->pH) : energy_folded (m, state, c->pH)) - c->gzero;
        if (c->pass == 0) { if (s == c->s0 || g < c->local_min) c->local_min = g; }
        else {
            const double b = exp ((g - c->gmin) * scale);          /* (:268-273: shift, scale, exp) */
            c->acc[m->ninst] += (long double) b;
            for (int i = 0; i < m->nsites; i++) c->acc[state[i]] += (long double) b;
        }
        increment (m, state);
    }
    free (state);
    return NULL;
}

int oracle_analytic_compensated (int nsites, int ninst, double temperature, const int *first, const int *last,
                                 const int *protons, const double *gintr, const double *gmodel, const double *w,
                                 double pH, double gzero, int unfolded, long long nstates, int nthreads,
                                 double *out_p, double *out_z, double *out_gmin) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if ((long long) nthreads > nstates) nthreads = (int) nstates;
    CompSlice   *slices = (CompSlice *) calloc ((size_t) nthreads, sizeof (CompSlice));
    pthread_t   *threads = (pthread_t *) calloc ((size_t) nthreads, sizeof (pthread_t));
    long double *sums = (long double *) calloc ((size_t) nthreads * (size_t) (ninst + 1), sizeof (long double));
    if (slices == NULL || threads == NULL || sums == NULL) { free (slices); free (threads); free (sums); return 1; }
    double gmin = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        for (int t = 0; t < nthreads; t++) {
            CompSlice *c = &slices[t];
            c->m = &m; c->s0 = nstates * t / nthreads; c->s1 = nstates * (t + 1) / nthreads; c->pass = pass; c->unfolded = unfolded;
            c->pH = pH; c->gzero = gzero; c->gmin = gmin; c->acc = sums + (size_t) t * (size_t) (ninst + 1);
            if (pthread_create (&threads[t], NULL, compensated_slice, c) != 0) { compensated_slice (c); threads[t] = 0; }
        }
        for (int t = 0; t < nthreads; t++) if (threads[t]) pthread_join (threads[t], NULL);
        if (pass == 0) { gmin = slices[0].local_min; for (int t = 1; t < nthreads; t++) if (slices[t].local_min < gmin) gmin = slices[t].local_min; }
    }
    long double z = 0.0L;
    for (int t = 0; t < nthreads; t++) z += sums[(size_t) t * (size_t) (ninst + 1) + ninst];
    if (out_p != NULL)
        for (int k = 0; k < ninst; k++) {
            long double v = 0.0L;
            for (int t = 0; t < nthreads; t++) v += sums[(size_t) t * (size_t) (ninst + 1) + k];
            out_p[k] = (double) (v / z);
        }
    if (out_z != NULL) *out_z = (double) z;
    if (out_gmin != NULL) *out_gmin = gmin;
    free (slices); free (threads); free (sums);
    return 0;
}

/* energies of the first n states in odometer order (what twosites.py:24-29 prints) */
void oracle_enumerate_energies (int nsites, int ninst, double temperature, const int *first, const int *last,
                                const int *protons, const double *gintr, const double *gmodel, const double *w,
                                double pH, long long n, double *out) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    int *state = (int *) malloc (sizeof (int) * (size_t) nsites);
    for (int i = 0; i < nsites; i++) state[i] = first[i];
    for (long long s = 0; s < n; s++) { out[s] = energy_folded (&m, state, pH); increment (&m, state); }
    free (state);
}

/* MCModelDefault_FindMaxInteraction + _FindPairs, MCModelDefault.c:231-288.  Pair order: site i
 * ascending, then j < i ascending; stored as (a = site i, b = site j).  Returns the number found;
 * fills up to max_pairs entries. */
int oracle_find_pairs (int nsites, int ninst, const int *first, const int *last, const double *w, double limit,
                       int max_pairs, int *pair_a, int *pair_b, double *pair_wmax) {
    int nfound = 0;
    for (int i = 0; i < nsites; i++)
        for (int j = 0; j < i; j++) {
            double wmax = 0.0;
            for (int a = first[i]; a <= last[i]; a++)
                for (int b = first[j]; b <= last[j]; b++) {
                    double v = fabs (w[(size_t) a * ninst + b]);
                    if (v > wmax) wmax = v;
                }
            if (wmax >= limit) {
                if (nfound < max_pairs) { pair_a[nfound] = i; pair_b[nfound] = j; pair_wmax[nfound] = wmax; }
                nfound++;
            }
        }
    return nfound;
}

/* MT19937, the generator the reference requests from pCore (MCModelDefault.c:28) */
typedef struct { uint32_t mt[624]; int mti; } MT;
static void mt_seed (MT *g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; i++) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t) i;
    g->mti = 624;
}
static uint32_t mt_next (MT *g) {
    if (g->mti >= 624) {
        for (int k = 0; k < 624; k++) {
            uint32_t y = (g->mt[k] & 0x80000000u) | (g->mt[(k + 1) % 624] & 0x7fffffffu);
            g->mt[k] = g->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->mti = 0;
    }
    uint32_t y = g->mt[g->mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}
static double mt_real (MT *g) { return (double) mt_next (g) * (1.0 / 4294967296.0); }

/* MCModelDefault_Metropolis, MCModelDefault.c:71-85 */
static int metropolis_mt (double x, MT *g) {
    if (x < 0.0) return 1;
    if (-x < PCE_TOO_SMALL) return 0;
    return mt_real (g) < exp (-x);
}

typedef struct { const OModel *m; int *state; const int *pair_a, *pair_b; int npairs; MT *g; } MCRef;

/* MCModelDefault_Move, MCModelDefault.c:90-122 */
static int move_mt (MCRef *c, double pH, double g, double *gnew) {
    const OModel *m = c->m;
    int site = (int) (mt_next (c->g) % (uint32_t) m->nsites);
    int inst;
    do { inst = (int) (mt_next (c->g) % (uint32_t) (m->last[site] - m->first[site] + 1)) + m->first[site]; } while (inst == c->state[site]);
    int    old = c->state[site];
    double dgintr = m->gintr[inst] - m->gintr[old];
    int    dprot  = m->protons[inst] - m->protons[old];
    double w = 0.0;
    for (int j = 0; j < m->nsites; j++)
        w += m->w[(size_t) inst * m->ninst + c->state[j]] - m->w[(size_t) old * m->ninst + c->state[j]];
    double dg = dgintr - dprot * mu_of (m->temperature, pH) + w;
    double beta = 1.0 / (PCE_R * m->temperature);
    if (metropolis_mt (dg * beta, c->g)) { c->state[site] = inst; *gnew = g + dg; return 1; }
    return 0;
}

/* MCModelDefault_DoubleMove, MCModelDefault.c:127-172 */
static int double_move_mt (MCRef *c, double pH, double g, double *gnew) {
    const OModel *m = c->m;
    int p  = (int) (mt_next (c->g) % (uint32_t) c->npairs);
    int sa = c->pair_a[p], sb = c->pair_b[p];
    int ia, ib;
    do { ia = (int) (mt_next (c->g) % (uint32_t) (m->last[sa] - m->first[sa] + 1)) + m->first[sa]; } while (ia == c->state[sa]);
    do { ib = (int) (mt_next (c->g) % (uint32_t) (m->last[sb] - m->first[sb] + 1)) + m->first[sb]; } while (ib == c->state[sb]);
    int oa = c->state[sa], ob = c->state[sb];
    const size_t n = (size_t) m->ninst;
    double dgintr = m->gintr[ia] - m->gintr[oa] + m->gintr[ib] - m->gintr[ob];
    int    dprot  = m->protons[ia] - m->protons[oa] + m->protons[ib] - m->protons[ob];
    double w = m->w[ia * n + ib] - m->w[oa * n + ob];
    for (int j = 0; j < m->nsites; j++) {
        if (j == sa || j == sb) continue;
        int aj = c->state[j];
        w += m->w[ia * n + aj] - m->w[oa * n + aj] + m->w[ib * n + aj] - m->w[ob * n + aj];
    }
    double dg = dgintr - dprot * mu_of (m->temperature, pH) + w;
    double beta = 1.0 / (PCE_R * m->temperature);
    if (metropolis_mt (dg * beta, c->g)) { c->state[sa] = ia; c->state[sb] = ib; *gnew = g + dg; return 1; }
    return 0;
}

/* MCModelDefault_MCScan, MCModelDefault.c:178-209 */
static double scan_mt (MCRef *c, double pH, long long *counters) {
    const OModel *m = c->m;
    double g = energy_folded (m, c->state, pH), gnew = 0.0;
    int nmoves = m->nsites + c->npairs;
    for (int k = 0; k < nmoves; k++) {
        int select = (int) (mt_next (c->g) % (uint32_t) nmoves);
        if (select < m->nsites) { counters[0]++; if (move_mt (c, pH, g, &gnew)) { counters[1]++; g = gnew; } }
        else                    { counters[2]++; if (double_move_mt (c, pH, g, &gnew)) { counters[3]++; g = gnew; } }
    }
    return g;
}

/* MCModelDefault_Equilibration + _Production, MCModelDefault.c:298-327 (StateVector_Randomize,
 * StateVector.c:163-170; MCModelDefault_UpdateProbabilities, MCModelDefault.c:215-226), one MT19937
 * stream seeded once.  Runs npH pH points back to back on the continuing stream, exactly like a
 * TitrationCurves sweep (TitrationCurves.py:119-121).  energies (may be NULL): [npH*nprod]. */
int oracle_mc_mt (int nsites, int ninst, double temperature, const int *first, const int *last,
                  const int *protons, const double *gintr, const double *gmodel, const double *w,
                  int npairs, const int *pair_a, const int *pair_b,
                  int nequil, int nprod, uint32_t seed, int nph, const double *ph,
                  double *out_p, double *energies, long long *counters) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    MT g; mt_seed (&g, seed);
    int *state = (int *) malloc (sizeof (int) * (size_t) nsites);
    for (int i = 0; i < nsites; i++) state[i] = first[i];
    MCRef c = { &m, state, pair_a, pair_b, npairs, &g };
    long long scratch[4];
    for (int q = 0; q < nph; q++) {
        double *p = out_p + (size_t) q * ninst;
        long long *cnt = counters ? counters + 4 * q : scratch;
        for (int i = 0; i < nsites; i++) state[i] = (int) (mt_next (&g) % (uint32_t) (last[i] - first[i] + 1)) + first[i];
        cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
        for (int s = 0; s < nequil; s++) scan_mt (&c, ph[q], cnt);
        cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
        for (int k = 0; k < ninst; k++) p[k] = 0.0;
        for (int s = 0; s < nprod; s++) {
            double e = scan_mt (&c, ph[q], cnt);
            if (energies) energies[(size_t) q * nprod + s] = e;
            for (int i = 0; i < nsites; i++) p[state[i]] += 1.0;
        }
        double scale = 1.0 / nprod;
        for (int k = 0; k < ninst; k++) p[k] *= scale;
    }
    free (state);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Part 2: executable specification of the GPU chain (Philox stream + cached fields)
 * ---------------------------------------------------------------------------------------------- */

/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Constants from the paper. */
void oracle_philox4x32_10 (const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t mulhi32 (uint32_t a, uint32_t b, uint32_t *lo) {
    uint64_t p = (uint64_t) a * b;
    if (lo) *lo = (uint32_t) p;
    return (uint32_t) (p >> 32);
}

/* One chain of the GPU sampler, operation for operation (see DESIGN.md "MC chain specification"):
 *   stream   : trial t of chain c at pH index q uses Philox block ctr = {lo(t>>1), hi(t>>1), c, q},
 *              key = {lo(seed), hi(seed)}; words (0,1) for even t, (2,3) for odd t.
 *              Initial state: ctr = {i>>2, 0xFFFFFFFF, c, q}, word i&3 -> first_i + mulhi(word, n_i).
 *   proposal : idx = mulhi(w0, nsites+npairs), remainder bits drive the instance draws; idx < nsites is
 *              a single-site move of site idx (statistically identical to MCModelDefault.c:189 +
 *              :97: a uniform type draw followed by an independent uniform site draw); the new
 *              instance is uniform over the *other* instances (k in [0,n-1), skip current) instead of
 *              the reference's rejection loop (MCModelDefault.c:99-101) -- same distribution.
 *   energy   : everything in units of RT.  wb = beta*W, gtb[k] = beta*(Gintr[k] - protons[k]*mu(pH)).  Only energy
 *              DIFFERENCES between instances of one site ever enter a decision, so every chain caches the fused
 *              fields RELATIVE TO THE FIRST INSTANCE OF THE SITE: for instance k of site s with first instance f,
 *              g[r(k)] = (gtb[k] - gtb[f]) + sum_j wr[a_j][r(k)],  wr[a][r(k)] = wb[a][k] - wb[a][f]  (one rounding),
 *              r(k) = k - s - 1 (the first instances are not stored: their relative field is 0) -- ninst - nsites
 *              values per chain instead of ninst.  With G(k) = 0 for a first instance and g[r(k)] otherwise, a single
 *              move has dG/RT = G(new) - G(old); a double move adds the second site and the pair cross term
 *              wb[ia][ib] - wb[ia][ob] - wb[oa][ib] + wb[oa][ob].  An accepted move old -> new adds the difference row
 *              E[r] = wr[new][r] - wr[old][r] (one rounding; stored once per instance pair lo < hi of a site, the
 *              opposite direction is its exact negation) to every g[r].  Equal in exact arithmetic to
 *              MCModelDefault.c:103-117 and :146-164 (same-site W is zero by construction).
 *   accept   : MCModelDefault_Metropolis (MCModelDefault.c:71-85) with u = w1 * 2^-32.
 *   counting : residence times, equal to adding 1 to the active instance of every site after every
 *              production scan (MCModelDefault.c:215-226, :306-309).
 * Outputs: counts[ninst] (production scans during which each instance was active), counters[4]
 * (moves, moves accepted, flips, flips accepted over production), esum = sum over production scans of
 * the tracked energy, final_state[nsites], final_energy. */
int oracle_mc_philox_chain (int nsites, int ninst, double temperature, const int *first, const int *last,
                            const int *protons, const double *gintr, const double *gmodel, const double *w,
                            int npairs, const int *pair_a, const int *pair_b,
                            int nequil, int nprod, uint64_t seed, uint32_t chain, uint32_t ph_index, double pH,
                            long long *counts, long long *counters, double *esum, int *final_state, double *final_energy) {
    OModel m = make_model (nsites, ninst, temperature, first, last, protons, gintr, gmodel, w);
    const uint32_t key[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
    const size_t n = (size_t) ninst;
    int       *state = (int *) malloc (sizeof (int) * (size_t) nsites);
    long long *since = (long long *) calloc ((size_t) nsites, sizeof (long long));
    double    *wb    = (double *) malloc (sizeof (double) * n * n);
    uint32_t   word[4];
    const double mu   = mu_of (temperature, pH);
    const double beta = 1.0 / (PCE_R * temperature);
    for (size_t e = 0; e < n * n; e++) wb[e] = beta * w[e];
    /* relative fields: site of every instance, r(k), and wr[a][r] = wb[a][k] - wb[a][first of k's site] */
    const size_t nr = n - (size_t) nsites;
    int    *site_of = (int *) malloc (sizeof (int) * n);
    double *g       = (double *) malloc (sizeof (double) * (nr > 0 ? nr : 1));
    double *wr      = (double *) malloc (sizeof (double) * n * (nr > 0 ? nr : 1));
    for (int i = 0; i < nsites; i++) for (int k = first[i]; k <= last[i]; k++) site_of[k] = i;
    for (size_t a = 0; a < n; a++)
        for (size_t k = 0; k < n; k++) {
            const int sk = site_of[k], f = first[sk];
            if ((int) k != f) wr[a * nr + (k - (size_t) sk - 1)] = wb[a * n + k] - wb[a * n + (size_t) f];
        }
#define RELG(k_) ((k_) == first[site_of[(k_)]] ? 0.0 : g[(size_t) (k_) - (size_t) site_of[(k_)] - 1])
    /* initial state */
    for (int i = 0; i < nsites; i++) {
        if ((i & 3) == 0 || i == 0) {
            uint32_t ctr[4] = { (uint32_t) (i >> 2), 0xFFFFFFFFu, chain, ph_index };
            oracle_philox4x32_10 (ctr, key, word);
        }
        state[i] = first[i] + (int) mulhi32 (word[i & 3], (uint32_t) (last[i] - first[i] + 1), NULL);
    }
    for (size_t k = 0; k < n; k++) {
        const int sk = site_of[k], f = first[sk];
        if ((int) k == f) continue;
        const size_t r = k - (size_t) sk - 1;
        const double gk = beta * (gintr[k] - (double) protons[k] * mu), gf = beta * (gintr[f] - (double) protons[f] * mu);
        double acc = gk - gf;
        for (int j = 0; j < nsites; j++) acc += wr[(size_t) state[j] * nr + r];
        g[r] = acc;
    }
    double grt = beta * energy_folded (&m, state, pH);
    const uint32_t nmoves = (uint32_t) (nsites + npairs);
    for (size_t k = 0; k < n; k++) counts[k] = 0;
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    double es = 0.0;
    uint64_t t = 0;
    for (long long scan = 0; scan < (long long) nequil + nprod; scan++) {
        const long long tp = scan < nequil ? 0 : scan - nequil;
        const int production = scan >= nequil;
        for (uint32_t mv = 0; mv < nmoves; mv++, t++) {
            if ((t & 1) == 0) {
                uint64_t b = t >> 1;
                uint32_t ctr[4] = { (uint32_t) b, (uint32_t) (b >> 32), chain, ph_index };
                oracle_philox4x32_10 (ctr, key, word);
            }
            const uint32_t w0 = word[2 * (t & 1)], w1 = word[2 * (t & 1) + 1];
            uint32_t lo;
            const uint32_t idx = mulhi32 (w0, nmoves, &lo);
            int sa, sb = -1, na, nb = 0, oa, ob = -1, ia, ib = -1, valid = 1;
            if (idx < (uint32_t) nsites) { sa = (int) idx; }
            else { sa = pair_a[idx - (uint32_t) nsites]; sb = pair_b[idx - (uint32_t) nsites]; }
            oa = state[sa]; na = last[sa] - first[sa] + 1;
            if (na < 2) valid = 0;
            uint32_t lo2 = 0;
            uint32_t ka = na >= 2 ? mulhi32 (lo, (uint32_t) (na - 1), &lo2) : 0;
            ia = first[sa] + (int) ka; if (ia >= oa) ia++;
            if (sb >= 0) {
                ob = state[sb]; nb = last[sb] - first[sb] + 1;
                if (nb < 2) valid = 0;
                uint32_t kb = nb >= 2 ? mulhi32 (lo2, (uint32_t) (nb - 1), NULL) : 0;
                ib = first[sb] + (int) kb; if (ib >= ob) ib++;
            }
            if (production) counters[sb >= 0 ? 2 : 0]++;
            if (!valid) continue;
            double x = RELG (ia) - RELG (oa);
            if (sb >= 0) {
                double cross = wb[(size_t) ia * n + ib] - wb[(size_t) ia * n + ob] - wb[(size_t) oa * n + ib] + wb[(size_t) oa * n + ob];
                x = (x + (RELG (ib) - RELG (ob))) + cross;
            }
            int accept;
            if (x < 0.0) accept = 1;
            else if (-x < PCE_TOO_SMALL) accept = 0;
            else accept = ((double) w1 * (1.0 / 4294967296.0)) < exp (-x);
            if (!accept) continue;
            if (production) counters[sb >= 0 ? 3 : 1]++;
            counts[oa] += tp - since[sa]; since[sa] = tp; state[sa] = ia;
            if (sb >= 0) { counts[ob] += tp - since[sb]; since[sb] = tp; state[sb] = ib; }
            for (size_t r = 0; r < nr; r++) {
                /* E = wr[hi] - wr[lo] of the instance pair, added with the sign of the direction */
                const int hia = ia > oa ? ia : oa, loa = ia > oa ? oa : ia;
                const double ea = wr[(size_t) hia * nr + r] - wr[(size_t) loa * nr + r];
                double v = ia > oa ? g[r] + ea : g[r] - ea;
                if (sb >= 0) {
                    const int hib = ib > ob ? ib : ob, lob = ib > ob ? ob : ib;
                    const double eb = wr[(size_t) hib * nr + r] - wr[(size_t) lob * nr + r];
                    v = ib > ob ? v + eb : v - eb;
                }
                g[r] = v;
            }
            grt = grt + x;
        }
        if (production) es += grt;
    }
    for (int i = 0; i < nsites; i++) counts[state[i]] += (long long) nprod - since[i];
    if (esum) *esum = es / beta;
    if (final_state) for (int i = 0; i < nsites; i++) final_state[i] = state[i];
    if (final_energy) *final_energy = grt / beta;
#undef RELG
    free (state); free (since); free (g); free (wr); free (site_of); free (wb);
    return 0;
}
