// pce_enum_plan.h -- host-side planning of kernel 2 (factorised exhaustive enumeration).
//
// Plain C++ on plain arrays (no CUDA types): the library (pce_analytic.cu) and the CTA-program emulator the CPU tests
// build (tests/emu/enum_emu.cpp) compile the same planner, so the roles, tables and scales the GPU kernel works with
// are the ones the CPU tests check against the oracle.
//
// Replaces the loop structure of EnergyModel_CalculateZ / _CalculateProbabilitiesFromZ (extensions/csource/EnergyModel.c
// :252-281, :318-337): instead of one energy evaluation per microstate and pH value, sites take ROLES and the Boltzmann
// factor of a state is a product of table entries (see pce_enum_core.cuh):
//
//   M   <= 8 sites, <= 128 configurations: one configuration per thread, fixed for the whole kernel (split Lo x Hi)
//   I   <= 4 sites, <= 16 configurations: the innermost loop, one DFMA per state, proton-count groups known in advance
//   J   one binary site (fastest kernel variant only): both of its instances are handled by the same thread with two sets
//       of register factors, so every tile-table entry fetched from shared memory feeds two DFMAs
//   O1  <= 3 sites, <= 8 configurations: unrolled loop around it ("tile" = one configuration of O1 x O2)
//   O2  <= 16 configurations, sorted by proton count: rolled loop
//   F   all remaining sites: a configuration of them is a "chunk"; chunks are dealt out to persistent CTAs
//
// One enumeration yields the proton-count bins of Z and of every instance of the M and F sites; the sites that were
// looped (I, O1, O2) get their bins from a second enumeration with the roles rotated.
//
// Bins are stored with a per-bin scale exp(sigma(n)): sigma is the convex envelope, over the pH values of the call,
// of tangent lines through (approximate) minimum-energy states -- so every pH value of a grid of ANY width finds its
// relevant states near 1 in double precision (round 1 needed one enumeration per ~9 pH units of a 32-proton model).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace pce_enum {

constexpr int kBlock    = 128;   // threads per CTA = max configurations of the M group
constexpr int kMarginalBlock = 256;   // threads per block of the marginal program
constexpr int kMaxM     = 8;
constexpr int kMaxI     = 4;
constexpr int kIStates  = 16;    // max configurations of the I group (tile table columns)
constexpr int kXStates  = 32;    // I x J configurations (per-thread register factors)
constexpr int kMaxO1    = 3;
constexpr int kO1States = 8;
constexpr int kMaxO2    = 7;
constexpr int kO2States = 16;
constexpr int kMaxTiles = kO1States * kO2States;
constexpr int kMaxO     = kMaxO1 + kMaxO2;
constexpr int kMaxF     = 96;
constexpr int kMaxRadix = 8;     // instances per site in the M / I / O roles
constexpr int kMaxInner = (kMaxM + kMaxI + 1) * kMaxRadix;   // instances of the M, I and J sites
constexpr int kMaxOinst = kMaxO * kMaxRadix;
constexpr double kR = 0.001987165392, kLn10 = 2.302585092994;      // EnergyModel.h:31-32

struct ModelView {            // host arrays of the model (W symmetric, row-major, row stride ninst)
    int nsites, ninst;
    const int32_t *first, *last, *protons;
    const double  *g;         // intrinsic energies (Gintr, or Gmodel for the unfolded state)
    const double  *w;
    double temperature;
    int    unfolded;
};

inline double mu_of (double temperature, double ph) { return -kR * temperature * kLn10 * ph; }

// ---------------------------------------------------------------------------------------------------
// roles of one enumeration
// ---------------------------------------------------------------------------------------------------
struct RunPlan {
    std::vector<int> M, I, J, O1, O2, F;   // site ids per role (J: none or one)
    int nLo = 0;
    std::vector<uint8_t> coverM, coverF;   // this run provides the bins of the site's instances
    int  mode = 0;                         // 0 generic | 1: I = 4 and O1 = 3 canonical sites (binary, one proton apart) | 2: + a canonical J site
};

struct Plan {
    std::vector<RunPlan> runs;
    int    Nb = 0;                 // absolute proton-count bins 0 .. ptotal
    int    base = 0;               // sum over sites of the smallest proton count
    double ph_star = 7.0, gref = 0.0;
    double nstates = 1.0;
    std::vector<double> sigma;     // [Nb] log-scale of every bin
};

inline int radix (const ModelView &m, int s) { return m.last[s] - m.first[s] + 1; }

inline int min_protons (const ModelView &m, int s) {
    int best = m.protons[m.first[s]];
    for (int k = m.first[s]; k <= m.last[s]; k++) best = std::min (best, (int) m.protons[k]);
    return best;
}
inline int max_protons (const ModelView &m, int s) {
    int best = m.protons[m.first[s]];
    for (int k = m.first[s]; k <= m.last[s]; k++) best = std::max (best, (int) m.protons[k]);
    return best;
}
inline int max_protons_total (const ModelView &m) {
    int total = 0;
    for (int i = 0; i < m.nsites; i++) total += std::max (0, max_protons (m, i));
    return total;
}
inline bool canonical (const ModelView &m, int s) {
    return radix (m, s) == 2 && std::abs (m.protons[m.first[s]] - m.protons[m.first[s] + 1]) == 1;
}

// instances of a site in ascending proton order (stable): digit d of the site in any role means order[d]
inline void sorted_instances (const ModelView &m, int s, int *out) {
    const int n = radix (m, s);
    for (int d = 0; d < n; d++) out[d] = m.first[s] + d;
    std::stable_sort (out, out + n, [&] (int a, int b) { return m.protons[a] < m.protons[b]; });
}

// A low-energy state at the given pH by coordinate descent (started from `a`, or from the per-site minima of the
// tilted intrinsic energies when a is empty).  Returns its energy at that pH; only rough optimality matters.
// `field` caches sum_j W[k][a_j] for every instance k (the site's own instances contribute nothing: same-site entries are
// ignored throughout); it is built on the first call and updated per changed site, so a warm-started descent for the next
// pH value of a grid costs O(ninst) per sweep instead of O(ninst * nsites).
inline double descend (const ModelView &m, double ph, std::vector<int> &a, std::vector<double> *field_cache = nullptr) {
    const double mu = mu_of (m.temperature, ph);
    const size_t n = (size_t) m.ninst;
    auto tilted = [&] (int k) { return m.g[k] - m.protons[k] * mu; };
    std::vector<double> local;
    std::vector<double> &field = field_cache ? *field_cache : local;
    bool fresh = false;
    if ((int) a.size () != m.nsites) {
        a.assign (m.nsites, 0);
        for (int i = 0; i < m.nsites; i++) {
            int best = m.first[i];
            for (int k = m.first[i]; k <= m.last[i]; k++) if (tilted (k) < tilted (best)) best = k;
            a[i] = best;
        }
        fresh = true;
    }
    std::vector<int> site_of;
    if (!m.unfolded) {
        site_of.resize (n);
        for (int i = 0; i < m.nsites; i++) for (int k = m.first[i]; k <= m.last[i]; k++) site_of[(size_t) k] = i;
        if (fresh || field.size () != n) {
            field.assign (n, 0.0);
            for (size_t k = 0; k < n; k++) {
                double f = 0.0;
                for (int j = 0; j < m.nsites; j++) if (j != site_of[k]) f += m.w[k * n + a[j]];
                field[k] = f;
            }
        }
        for (int sweep = 0; sweep < 20; sweep++) {
            bool changed = false;
            for (int i = 0; i < m.nsites; i++) {
                int    best = a[i];
                double ebest = tilted (a[i]) + field[(size_t) a[i]];
                for (int k = m.first[i]; k <= m.last[i]; k++) {
                    const double e = tilted (k) + field[(size_t) k];
                    if (e < ebest) { ebest = e; best = k; }
                }
                if (best != a[i]) {
                    const double *w_new = m.w + (size_t) best * n, *w_old = m.w + (size_t) a[i] * n;       // W is symmetric: rows serve as columns
                    for (size_t k = 0; k < n; k++) if (site_of[k] != i) field[k] += w_new[k] - w_old[k];
                    a[i] = best; changed = true;
                }
            }
            if (!changed) break;
        }
    }
    double e = 0.0, pairs = 0.0;
    for (int i = 0; i < m.nsites; i++) {
        e += tilted (a[i]);
        if (!m.unfolded) pairs += field[(size_t) a[i]];
    }
    return e + 0.5 * pairs;
}

inline int protons_of (const ModelView &m, const std::vector<int> &a) {
    int n = 0;
    for (int i = 0; i < m.nsites; i++) n += m.protons[a[i]];
    return n;
}

// ---------------------------------------------------------------------------------------------------
// make_plan: reference pH and energy, bin scales, and the role assignment of every run
// ---------------------------------------------------------------------------------------------------
// grid_slots: CTAs the enumeration kernel keeps resident over all shards (sizes the O2 role); max_mode caps the kernel variant
// (test hook: 0 forces the generic accumulation path, 1 the variant without the J site)
inline int make_plan (const ModelView &m, const double *ph, int nph, Plan *plan, std::string *error, long long grid_slots = 0, int max_mode = 2) {
    auto bad = [&] (const char *text) { if (error) *error = text; return 1; };
    if (nph < 1) return bad ("empty pH list");
    for (int k = 0; k < m.ninst; k++) if (m.protons[k] < 0) return bad ("negative proton count");
    for (int s = 0; s < m.nsites; s++) if (radix (m, s) < 1) return bad ("site without instances");
    double lo = ph[0], hi = ph[0];
    for (int q = 1; q < nph; q++) { lo = std::min (lo, ph[q]); hi = std::max (hi, ph[q]); }
    plan->ph_star = 0.5 * (lo + hi);
    const int ptotal = max_protons_total (m);
    plan->Nb = ptotal + 1;
    // the bin scales absorb the pH dependence of whole bins; within a partial sum the protons of the M, I and O1 sites
    // (at most ~16) still see the half-width of the grid
    if (0.5 * (hi - lo) * kLn10 * std::min (ptotal, 16) > 600.0) return bad ("pH range too wide for one enumeration pass; split the pH grid");
    plan->base = 0;
    for (int s = 0; s < m.nsites; s++) plan->base += min_protons (m, s);
    plan->nstates = 1.0;
    for (int i = 0; i < m.nsites; i++) plan->nstates *= radix (m, i);

    // reference energy at pH*, and the bin scales: for every pH value q of the call (ascending) a low-energy state s_q
    // with n_q protons and energy e_q at pH*.  E_s(pH*) >= e_q + (mu_q - mu*) (n_s - n_q) (- the descent's error) for
    // every state s, because s_q minimises E(q) = E(pH*) - n (mu_q - mu*): the maximum over q of these tangent lines
    // is a convex lower bound of the minimum energy at each proton count, tight where it matters (at the pH values asked for).
    std::vector<int> a;
    plan->gref = descend (m, plan->ph_star, a);
    {
        std::vector<double> sorted (ph, ph + nph);
        std::sort (sorted.begin (), sorted.end ());
        const double mu_star = mu_of (m.temperature, plan->ph_star);
        const double beta = 1.0 / (kR * m.temperature);
        std::vector<double> tn, te, ts;
        std::vector<int> state;
        std::vector<double> field;
        for (int q = 0; q < nph; q++) {
            if (q > 0 && sorted[q] == sorted[q - 1]) continue;
            const double eq = descend (m, sorted[q], state, &field);                 // energy at pH q, warm-started from the previous pH
            const int    nq = protons_of (m, state);
            const double delta = mu_of (m.temperature, sorted[q]) - mu_star;
            tn.push_back ((double) nq); te.push_back (eq + nq * delta - plan->gref); ts.push_back (delta);     // E(pH*) = E(q) + n (mu_q - mu*)
        }
        plan->sigma.assign ((size_t) plan->Nb, 0.0);
        for (int n = 0; n < plan->Nb; n++) {
            double best = -1.0e300;
            for (size_t k = 0; k < tn.size (); k++) best = std::max (best, te[k] + ts[k] * ((double) n - tn[k]));
            plan->sigma[(size_t) n] = beta * best;
        }
    }

    if (m.nsites > kMaxF + kMaxM + kMaxI + 1 + kMaxO) return bad ("too many sites for analytic enumeration");
    std::vector<uint8_t> covered (m.nsites, 0);
    int ncovered = 0;
    plan->runs.clear ();
    while (ncovered < m.nsites || plan->runs.empty ()) {
        RunPlan run;
        std::vector<uint8_t> used (m.nsites, 0);
        auto small = [&] (int s) { return radix (m, s) <= kMaxRadix; };
        // M: uncovered sites first, then any, while the product of radices fits in a CTA
        long long pm = 1;
        for (int pass = 0; pass < 2; pass++)
            for (int s = 0; s < m.nsites && (int) run.M.size () < kMaxM; s++) {
                if (used[s] || !small (s) || (pass == 0 && covered[s])) continue;
                if (pm * radix (m, s) <= kBlock) { run.M.push_back (s); used[s] = 1; pm *= radix (m, s); }
            }
        {   // Lo / Hi split: balance the two table sizes (P_hi <= 64)
            long long plo = 1;
            run.nLo = 0;
            for (size_t i = 0; i < run.M.size (); i++) {
                if (plo * plo >= pm) break;
                plo *= radix (m, run.M[i]); run.nLo = (int) i + 1;
            }
            while (run.nLo < (int) run.M.size () && pm / plo > 64) { plo *= radix (m, run.M[run.nLo]); run.nLo++; }
        }
        // I and O1: canonical sites (binary, one proton apart) so that the compile-time bin structure of the fast kernel
        // applies; already covered sites preferred (their bins are not needed from this run).  Otherwise any small sites.
        auto pick = [&] (std::vector<int> &role, int max_sites, long long max_states, bool only_canonical) {
            long long p = 1;
            for (int s : role) p *= radix (m, s);
            for (int pass = 0; pass < 2; pass++)
                for (int s = 0; s < m.nsites && (int) role.size () < max_sites; s++) {
                    if (used[s] || !small (s) || (only_canonical && !canonical (m, s)) || (pass == 0 && !covered[s])) continue;
                    if (p * radix (m, s) <= max_states) { role.push_back (s); used[s] = 1; p *= radix (m, s); }
                }
            return p;
        };
        int ncanon = 0;
        for (int s = 0; s < m.nsites; s++) if (!used[s] && canonical (m, s)) ncanon++;
        run.mode = ncanon >= kMaxI + 1 + kMaxO1 ? 2 : ncanon >= kMaxI + kMaxO1 ? 1 : 0;
        run.mode = std::min (run.mode, max_mode);
        pick (run.I, kMaxI, kIStates, run.mode > 0);
        if (run.mode == 2) pick (run.J, 1, 2, true);
        pick (run.O1, kMaxO1, kO1States, run.mode > 0);
        if (run.mode == 0) { pick (run.I, kMaxI, kIStates, false); pick (run.O1, kMaxO1, kO1States, false); }
        // O2: further sites looped inside a chunk.  A chunk costs a fixed overhead (~ one O2 configuration's worth of work) and the
        // chunks are dealt out to grid_slots CTAs, so the prefix of candidate sites with the least
        // ceil (chunks / slots) * (1 + P_O2) is taken; without a grid size, as many sites as fit.
        {
            double outer = 1.0;
            for (int s = 0; s < m.nsites; s++) if (!used[s]) outer *= radix (m, s);
            std::vector<int> candidates;
            long long p = 1;
            // (after the first run only covered sites qualify: an uncovered site looped here would cost a whole further run)
            for (int pass = 0; pass < (ncovered > 0 ? 1 : 2); pass++)
                for (int s = 0; s < m.nsites && (int) candidates.size () < kMaxO2; s++) {
                    if (used[s] || !small (s) || (pass == 0 && !covered[s])) continue;
                    if (std::find (candidates.begin (), candidates.end (), s) != candidates.end ()) continue;
                    if (p * radix (m, s) > kO2States) continue;
                    candidates.push_back (s); p *= radix (m, s);
                }
            size_t best = candidates.size ();
            if (grid_slots > 0) {
                double best_cost = std::ceil (outer / (double) grid_slots) * 2.0;
                best = 0;
                p = 1;
                for (size_t k = 0; k < candidates.size (); k++) {
                    p *= radix (m, candidates[k]);
                    const double cost = std::ceil (outer / (double) p / (double) grid_slots) * (1.0 + (double) p);
                    if (cost <= best_cost) { best_cost = cost; best = k + 1; }
                }
            }
            for (size_t k = 0; k < best; k++) { run.O2.push_back (candidates[k]); used[candidates[k]] = 1; }
        }
        for (int s = 0; s < m.nsites; s++) if (!used[s]) run.F.push_back (s);
        if ((int) run.F.size () > kMaxF) return bad ("too many sites for analytic enumeration");
        run.coverM.assign (run.M.size (), 0);
        run.coverF.assign (run.F.size (), 0);
        int newly = 0;
        for (size_t i = 0; i < run.M.size (); i++) if (!covered[run.M[i]]) { run.coverM[i] = 1; covered[run.M[i]] = 1; newly++; }
        for (size_t j = 0; j < run.F.size (); j++) if (!covered[run.F[j]]) { run.coverF[j] = 1; covered[run.F[j]] = 1; newly++; }
        ncovered += newly;
        plan->runs.push_back (run);
        if (newly == 0 && ncovered < m.nsites) return bad ("cannot cover all sites (a site has more than 8 instances and too many states)");
        if (plan->runs.size () > 16) return bad ("too many enumeration passes");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Tables of one run, built on the host (integers, a few KB) and uploaded; the numeric tables are built on the
// device by the static kernel (pce_enum_core.cuh)
// ---------------------------------------------------------------------------------------------------
struct RunTables {
    // sizes
    int nM = 0, nLo = 0, nI = 0, nJ = 0, nO1 = 0, nO2 = 0, nF = 0, nS = 0;      // nS = nM + nI + nJ inner sites
    int P_lo = 1, P_hi = 1, P_M = 1, P_I = 1, P_J = 1, P_O1 = 1, P_O2 = 1, P_O = 1;
    int nInner = 0, nMinst = 0, nOinst = 0;       // instances in M + I + J sites (in this order), in M sites, in O1 + O2 sites (O1 first)
    int maxrelM = 0, maxrelI = 0, maxrelO1 = 0, maxrelO2 = 0, maxrelF = 0;
    int NbC = 1, NbCp = 1, NbP = 1, NbAll = 1;    // rows of: chunk bins relative to thread protons | absolute in thread protons | persistent | all relative
    int nU = 1, nV = 1, cbar = 0;
    int mode = 0;
    long long nchunks = 1;
    unsigned flushI = 1;                          // bit i: sorted I configuration i is the last of its proton count
    // integer tables (device copies are made of exactly these arrays, in this order: see upload in pce_analytic.cu)
    std::vector<int16_t> innerInst, innerSite, innerOff;        // [nInner] global instance, inner site index; [nS + 1] start of each inner site
    std::vector<int16_t> oInst, oOff;                           // [nOinst] global instance of every O instance; [nO + 1]
    std::vector<int16_t> idxM;        // [kBlock][kMaxM] compact inner index of every M site's instance in thread tid's configuration
    std::vector<int16_t> idxLo, idxHi, idxI;      // [P_lo][kMaxM], [P_hi][kMaxM], [kXStates][kMaxI + 1] compact inner indices of configuration c = j * 16 + i
                                                  // (i: I configurations sorted by proton count; last entry: the J instance)
    std::vector<int16_t> idxJ;                    // [2] compact inner index of the J instances
    std::vector<int16_t> mlo, mhi;                // [kBlock] Lo / Hi configuration of thread tid
    std::vector<uint8_t> nrelM, nrelI, nrelO1, nrelO2;          // protons above the role's minimum: [kBlock], [kXStates] (I + J), [kO1States], [kO2States]
    std::vector<uint8_t> bO;                      // [P_O][kMaxO] compact O-instance index of every O site in tile t = t2 * P_O1 + t1
    std::vector<int16_t> fInst;                   // [nF][kMaxF-digit <= 255]: flattened with fOff: instance (global) of digit d of F site f
    std::vector<int32_t> fOff, fRad;              // [nF + 1], [nF]
    std::vector<uint8_t> fRel;                    // protons above the site's minimum, same layout as fInst
    std::vector<long long> fStride;               // [nF] digit stride within the chunk index
    std::vector<int32_t> fShift;                  // [nF] log2 of the stride when stride and radix are powers of two, else -1
    std::vector<int16_t> rowM;                    // [sum radM] global instance of (M site, digit), flattened: marginal rows
    std::vector<int32_t> rowMoff;                 // [nM + 1]
    std::vector<uint8_t> digitM;                  // [kBlock][kMaxM] digit of every M site in thread tid's configuration
    std::vector<uint8_t> coverM, coverF;          // [nM], [nF] this run provides the bins of the site
    std::vector<double>  D, sigmaU;               // [nU][nV] bin-scale corrections; [nU] tile-level log-scale
};

inline int build_tables (const ModelView &m, const Plan &plan, const RunPlan &run, RunTables *t, std::string *error) {
    auto bad = [&] (const char *text) { if (error) *error = text; return 1; };
    RunTables &T = *t;
    T = RunTables ();
    T.nM = (int) run.M.size (); T.nLo = run.nLo; T.nI = (int) run.I.size (); T.nJ = (int) run.J.size ();
    T.nO1 = (int) run.O1.size (); T.nO2 = (int) run.O2.size ();
    T.nF = (int) run.F.size (); T.nS = T.nM + T.nI + T.nJ;
    T.mode = run.mode;
    if (T.mode > 0 && (T.nI != kMaxI || T.nO1 != kMaxO1)) T.mode = 0;
    if (T.mode == 2 && T.nJ != 1) T.mode = 1;
    if (T.mode < 2 && T.nJ != 0) return bad ("internal: J site without its kernel variant");
    T.coverM = run.coverM; T.coverF = run.coverF;
    if (T.coverM.empty ()) T.coverM.push_back (0);
    if (T.coverF.empty ()) T.coverF.push_back (0);
    int order[64];
    // inner instances: M sites then I sites, each in ascending proton order
    std::vector<int> radM, radI, radO1, radO2;
    auto add_inner = [&] (int s, std::vector<int> &rads) {
        const int n = radix (m, s);
        sorted_instances (m, s, order);
        T.innerOff.push_back ((int16_t) T.innerInst.size ());
        for (int d = 0; d < n; d++) { T.innerInst.push_back ((int16_t) order[d]); T.innerSite.push_back ((int16_t) (T.innerOff.size () - 1)); }
        rads.push_back (n);
    };
    for (int s : run.M) add_inner (s, radM);
    T.nMinst = (int) T.innerInst.size ();
    for (int s : run.I) add_inner (s, radI);
    std::vector<int> radJ;
    for (int s : run.J) add_inner (s, radJ);
    T.innerOff.push_back ((int16_t) T.innerInst.size ());
    T.nInner = (int) T.innerInst.size ();
    if (T.nInner > kMaxInner) return bad ("too many instances in the thread-level sites");
    auto add_outer = [&] (int s, std::vector<int> &rads) {
        const int n = radix (m, s);
        sorted_instances (m, s, order);
        T.oOff.push_back ((int16_t) T.oInst.size ());
        for (int d = 0; d < n; d++) T.oInst.push_back ((int16_t) order[d]);
        rads.push_back (n);
    };
    for (int s : run.O1) add_outer (s, radO1);
    for (int s : run.O2) add_outer (s, radO2);
    T.oOff.push_back ((int16_t) T.oInst.size ());
    T.nOinst = (int) T.oInst.size ();
    auto rel = [&] (int inst, int site) { return m.protons[inst] - min_protons (m, site); };

    for (int i = 0; i < T.nM; i++) { if (i < T.nLo) T.P_lo *= radM[i]; else T.P_hi *= radM[i]; }
    T.P_M = T.P_lo * T.P_hi;
    for (int r : radI) T.P_I *= r;
    for (int r : radJ) T.P_J *= r;
    for (int r : radO1) T.P_O1 *= r;
    for (int r : radO2) T.P_O2 *= r;
    T.P_O = T.P_O1 * T.P_O2;
    if (T.P_M > kBlock || T.P_hi > 64 || T.P_I > kIStates || T.P_J > 2 || T.P_O1 > kO1States || T.P_O2 > kO2States) return bad ("internal: role sizes out of range");

    // M configurations: thread tid = m_lo + P_lo * m_hi, site 0 the fastest digit
    T.idxM.assign ((size_t) kBlock * kMaxM, 0); T.digitM.assign ((size_t) kBlock * kMaxM, 0);
    T.mlo.assign (kBlock, 0); T.mhi.assign (kBlock, 0); T.nrelM.assign (kBlock, 0);
    T.idxLo.assign ((size_t) std::max (T.P_lo, 1) * kMaxM, 0); T.idxHi.assign ((size_t) std::max (T.P_hi, 1) * kMaxM, 0);
    for (int tid = 0; tid < T.P_M; tid++) {
        int x = tid, np = 0;
        for (int i = 0; i < T.nM; i++) {
            const int d = x % radM[i];
            x /= radM[i];
            const int j = T.innerOff[i] + d;
            T.idxM[(size_t) tid * kMaxM + i] = (int16_t) j; T.digitM[(size_t) tid * kMaxM + i] = (uint8_t) d;
            np += rel (T.innerInst[j], run.M[i]);
        }
        T.mlo[tid] = (int16_t) (tid % T.P_lo); T.mhi[tid] = (int16_t) (tid / T.P_lo); T.nrelM[tid] = (uint8_t) np;
        T.maxrelM = std::max (T.maxrelM, np);
    }
    for (int x = 0; x < T.P_lo; x++) { int y = x; for (int i = 0; i < T.nLo; i++) { T.idxLo[(size_t) x * kMaxM + i] = (int16_t) (T.innerOff[i] + y % radM[i]); y /= radM[i]; } }
    for (int x = 0; x < T.P_hi; x++) { int y = x; for (int i = T.nLo; i < T.nM; i++) { T.idxHi[(size_t) x * kMaxM + (i - T.nLo)] = (int16_t) (T.innerOff[i] + y % radM[i]); y /= radM[i]; } }
    // I configurations sorted by proton count (stable); canonical: groups of C(4, k) configurations
    {
        std::vector<std::pair<int, int>> cfg;
        for (int c = 0; c < T.P_I; c++) {
            int x = c, np = 0;
            for (int i = 0; i < T.nI; i++) { np += rel (T.innerInst[T.innerOff[T.nM + i] + x % radI[i]], run.I[i]); x /= radI[i]; }
            cfg.push_back (std::make_pair (np, c));
        }
        std::stable_sort (cfg.begin (), cfg.end ());
        T.idxI.assign ((size_t) kXStates * (kMaxI + 1), 0); T.nrelI.assign (kXStates, 0);
        T.idxJ.assign (2, 0);
        T.flushI = 0;
        for (int j = 0; j < T.P_J; j++) {
            const int jinst = T.nJ ? T.innerOff[T.nM + T.nI] + j : 0;
            const int jrel = T.nJ ? rel (T.innerInst[jinst], run.J[0]) : 0;
            if (T.nJ) T.idxJ[j] = (int16_t) jinst;
            for (int i = 0; i < T.P_I; i++) {
                int x = cfg[i].second;
                int16_t *row = &T.idxI[(size_t) (j * kIStates + i) * (kMaxI + 1)];
                for (int k = 0; k < T.nI; k++) { row[k] = (int16_t) (T.innerOff[T.nM + k] + x % radI[k]); x /= radI[k]; }
                row[kMaxI] = (int16_t) jinst;
                T.nrelI[j * kIStates + i] = (uint8_t) (cfg[i].first + jrel);
                T.maxrelI = std::max (T.maxrelI, cfg[i].first + jrel);
                if (j == 0 && (i + 1 == T.P_I || cfg[i + 1].first != cfg[i].first)) T.flushI |= 1u << i;
            }
        }
        if (T.mode == 2 && (T.nrelI[kIStates] != 1 || T.nrelI[0] != 0)) return bad ("internal: canonical J site out of order");
    }
    // O1 configurations in natural order (canonical: protons above the minimum = popcount of t1); O2 sorted by proton count
    T.nrelO1.assign (kO1States, 0); T.nrelO2.assign (kO2States, 0);
    std::vector<int> o2cfg ((size_t) T.P_O2, 0);
    {
        std::vector<std::pair<int, int>> cfg;
        for (int c = 0; c < T.P_O2; c++) {
            int x = c, np = 0;
            for (int i = 0; i < T.nO2; i++) { np += rel (T.oInst[T.oOff[T.nO1 + i] + x % radO2[i]], run.O2[i]); x /= radO2[i]; }
            cfg.push_back (std::make_pair (np, c));
        }
        std::stable_sort (cfg.begin (), cfg.end ());
        for (int c = 0; c < T.P_O2; c++) { o2cfg[c] = cfg[c].second; T.nrelO2[c] = (uint8_t) cfg[c].first; T.maxrelO2 = std::max (T.maxrelO2, cfg[c].first); }
    }
    T.bO.assign ((size_t) T.P_O * kMaxO, 0);
    for (int t2 = 0; t2 < T.P_O2; t2++)
        for (int t1 = 0; t1 < T.P_O1; t1++) {
            uint8_t *row = &T.bO[(size_t) (t2 * T.P_O1 + t1) * kMaxO];
            int x = t1, np = 0;
            for (int i = 0; i < T.nO1; i++) { const int j = T.oOff[i] + x % radO1[i]; x /= radO1[i]; row[i] = (uint8_t) j; np += rel (T.oInst[j], run.O1[i]); }
            T.nrelO1[t1] = (uint8_t) np;
            T.maxrelO1 = std::max (T.maxrelO1, np);
            x = o2cfg[t2];
            for (int i = 0; i < T.nO2; i++) { row[T.nO1 + i] = (uint8_t) (T.oOff[T.nO1 + i] + x % radO2[i]); x /= radO2[i]; }
        }
    if (T.mode > 0) for (int t1 = 0; t1 < T.P_O1; t1++) if (T.nrelO1[t1] != __builtin_popcount ((unsigned) t1)) return bad ("internal: canonical O1 sites out of order");
    // F sites: digit tables, strides (site 0 of the role is the fastest digit of the chunk index)
    {
        long long stride = 1;
        double    count = 1.0;
        for (int f = 0; f < T.nF; f++) {
            const int s = run.F[f], n = radix (m, s);
            if (n > 64) return bad ("a site has more than 64 instances");
            sorted_instances (m, s, order);
            T.fOff.push_back ((int32_t) T.fInst.size ()); T.fRad.push_back (n); T.fStride.push_back (stride);
            {
                int shift = -1;
                if ((n & (n - 1)) == 0 && (stride & (stride - 1)) == 0) { shift = 0; while ((1LL << shift) < stride) shift++; }
                T.fShift.push_back (shift);
            }
            int mx = 0;
            for (int d = 0; d < n; d++) { T.fInst.push_back ((int16_t) order[d]); T.fRel.push_back ((uint8_t) rel (order[d], s)); mx = std::max (mx, rel (order[d], s)); }
            T.maxrelF += mx;
            count *= n;
            if (count > 4.0e18) return bad ("too many chunks");
            stride *= n;
        }
        T.fOff.push_back ((int32_t) T.fInst.size ());
        if (T.fShift.empty ()) T.fShift.push_back (-1);
        T.nchunks = stride;
    }
    // marginal rows of the M sites
    for (int i = 0; i < T.nM; i++) {
        T.rowMoff.push_back ((int32_t) T.rowM.size ());
        for (int d = 0; d < radM[i]; d++) T.rowM.push_back (T.innerInst[T.innerOff[i] + d]);
    }
    T.rowMoff.push_back ((int32_t) T.rowM.size ());
    // bins
    T.NbC  = T.maxrelO1 + T.maxrelO2 + T.maxrelI + 1;
    T.NbCp = T.NbC + T.maxrelM;
    T.NbP  = T.maxrelF + T.NbC;
    T.NbAll = T.maxrelF + T.NbCp;
    if (plan.base + T.NbAll > plan.Nb) return bad ("internal: bin count mismatch");
    // bin scales: the tile constant carries sigma (base + u + cbar), u = protons of the chunk and the O2 configuration above
    // their minimum; D[u][v] corrects a flushed partial sum with v = protons of O1 + I + M configuration to sigma (base + u + v)
    T.nU = T.maxrelF + T.maxrelO2 + 1;
    T.nV = T.maxrelO1 + T.maxrelI + T.maxrelM + 1;
    T.cbar = (T.nV - 1) / 2;
    T.D.assign ((size_t) T.nU * T.nV, 1.0); T.sigmaU.assign ((size_t) T.nU, 0.0);
    for (int u = 0; u < T.nU; u++) {
        const double su = plan.sigma[(size_t) std::min (plan.Nb - 1, plan.base + u + T.cbar)];
        T.sigmaU[(size_t) u] = su;
        for (int v = 0; v < T.nV; v++) T.D[(size_t) u * T.nV + v] = std::exp (plan.sigma[(size_t) std::min (plan.Nb - 1, plan.base + u + v)] - su);
    }
    return 0;
}

}  // namespace pce_enum
