// enum_emu.cpp -- CPU emulator of kernel 2's CTA programs.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles pcetk_b200/csrc/pce_enum_plan.h and pce_enum_core.cuh for the host and runs every phase of the static, the
// enumeration and the marginal programs for tid = 0..255 in a loop where the CUDA kernels have a __syncthreads().  The
// planner, the tables, the index arithmetic and the floating-point expressions are therefore the GPU's own; only the
// scheduling differs.  tests/test_enum_emulator.py compares the resulting bins (finished by the library's
// pce_analytic_finish) with the oracle on CPU.  Built by tests/emu/build.py into tests/emu/_build/ (git-ignored).
#define PCE_ENUM_EMULATE 1
#include "../../pcetk_b200/csrc/pce_enum_core.cuh"

#include <cstdio>

using namespace pce_enum;

namespace {

struct HostLocate {
    template <typename T> const T *operator() (const std::vector<T> &v) const { return v.data (); }
};

double butterfly_sum (double *v) {         // the order of warp_sum (__shfl_xor 16, 8, 4, 2, 1), lane 0's result
    for (int o = 16; o > 0; o >>= 1) for (int l = 0; l < 32; l++) if ((l & o) == 0) { const double s = v[l] + v[l ^ o]; v[l] = s; v[l ^ o] = s; }
    return v[0];
}

template <int MODE>
void run_enum (const EnumParams &p, int grid) {
    const EnumLayout L (p);
    std::vector<unsigned char> smem (L.total + 64);
    std::vector<EnumThread> regs (kBlock);
    for (int cta = 0; cta < grid; cta++) {
        EnumCta<MODE> c (p, smem.data ());
        for (int tid = 0; tid < kBlock; tid++) c.init (tid, regs[tid]);
        for (long long chunk = cta; chunk < p.nchunks_local; chunk += grid) {
            for (int tid = 0; tid < kBlock; tid++) c.phase0 (tid, p.chunk_begin + chunk);
            for (int tid = 0; tid < kBlock; tid++) c.phase1 (tid);
            for (int tid = 0; tid < kBlock; tid++) c.phase2 (tid);
            for (int tid = 0; tid < kBlock; tid++) c.phase3 (tid);
            for (int tid = 0; tid < kBlock; tid++) c.phase4 (tid, regs[tid]);
            for (int tid = 0; tid < kBlock; tid++) c.accumulate (tid, regs[tid]);
            for (int tid = 0; tid < kBlock; tid++) c.fold (tid);
            if (p.needF)
                for (int row = 0; row < p.NbCp; row++) {
                    double v[32];
                    for (int lane = 0; lane < 32; lane++) v[lane] = c.row_partial (row, lane);
                    c.store_row (row, chunk, butterfly_sum (v));
                }
        }
        for (int tid = 0; tid < kBlock; tid++) c.finish (tid, cta);
    }
}

}  // namespace

// bins: [(1 + ninst)][Nb] (scaled by exp (sigma[n])), sigma: [Nb].  grid = emulated CTAs of the enumeration program.
extern "C" int emu_analytic_bins (int nsites, int ninst, const int32_t *first, const int32_t *last, const int32_t *protons, const double *g,
                                  const double *w, double temperature, int unfolded, const double *ph, int nph, int shard, int nshards,
                                  int grid, int grid_slots, int max_mode, double *bins, double *sigma, int *info, char *error, int error_length) {
    ModelView m = { nsites, ninst, first, last, protons, g, w, temperature, unfolded };
    Plan plan;
    std::string message;
    auto fail = [&] (int code) { snprintf (error, (size_t) error_length, "%s", message.c_str ()); return code; };
    if (make_plan (m, ph, nph, &plan, &message, grid_slots, max_mode)) return fail (1);
    const int Nb = plan.Nb;
    for (int n = 0; n < Nb; n++) sigma[n] = plan.sigma[(size_t) n];
    for (size_t e = 0; e < (size_t) (1 + ninst) * Nb; e++) bins[e] = 0.0;
    std::vector<double> gt ((size_t) ninst);
    const double mu_star = mu_of (temperature, plan.ph_star);
    for (int a = 0; a < ninst; a++) gt[(size_t) a] = g[a] - (double) protons[a] * mu_star;
    info[0] = (int) plan.runs.size (); info[1] = Nb; info[2] = 0; info[3] = 0;
    for (size_t r = 0; r < plan.runs.size (); r++) {
        RunTables T;
        if (build_tables (m, plan, plan.runs[r], &T, &message)) return fail (2);
        EnumParams p;
        MarginalParams mp;
        std::memset (&p, 0, sizeof (p)); std::memset (&mp, 0, sizeof (mp));
        std::vector<double> numeric (kNumericDoubles, 0.0);
        bind_run (T, p, mp, HostLocate (), numeric.data ());
        p.ninst = ninst; p.ldw = ninst; p.unfolded = unfolded; p.gt = gt.data (); p.w = w;
        p.beta = 1.0 / (kR * temperature); p.gref = plan.gref;
        const long long c0 = T.nchunks * shard / nshards, c1 = T.nchunks * (shard + 1) / nshards;
        p.chunk_begin = c0; p.nchunks_local = c1 - c0; p.ldF = p.nchunks_local;
        info[2] |= T.mode << (2 * r); info[3] = (int) T.P_O2;
        if (p.nchunks_local <= 0) continue;
        const int g_ctas = (int) std::min<long long> (grid, p.nchunks_local);
        std::vector<double> Tbuf ((size_t) g_ctas * T.NbP * kBlock, 0.0), Ftab ((size_t) T.NbAll * p.ldF, 0.0), Mtab ((size_t) T.NbP * kBlock, 0.0);
        p.T = Tbuf.data (); p.FtabT = Ftab.data ();
        {   // static program
            const StaticLayout SL (p);
            std::vector<unsigned char> smem (SL.total + 64);
            StaticCta s (p, smem.data ());
            for (int tid = 0; tid < kBlock; tid++) s.phase0a (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase0b (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase1 (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase2 (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase3 (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase3b (tid);
            for (int tid = 0; tid < kBlock; tid++) s.phase4 (tid);
        }
        if (T.mode == 2) run_enum<2> (p, g_ctas); else if (T.mode == 1) run_enum<1> (p, g_ctas); else run_enum<0> (p, g_ctas);
        for (size_t e = 0; e < Mtab.size (); e++) { double v = 0.0; for (int c = 0; c < g_ctas; c++) v += Tbuf[(size_t) c * Mtab.size () + e]; Mtab[e] = v; }
        mp.Nb = Nb; mp.base = plan.base; mp.write_z = r == 0 ? 1 : 0; mp.grid = g_ctas;
        mp.chunk_begin = c0; mp.nchunks_local = p.nchunks_local; mp.ldF = p.ldF; mp.Mtab = Mtab.data (); mp.FtabT = Ftab.data (); mp.bins = bins;
        const int rows = marginal_rows (T);
        for (int b = 0; b < rows; b++) {
            int kind, site, digit, row;
            if (!marginal_target (mp, b, &kind, &site, &digit, &row)) continue;
            const long long count = kind == 2 ? mp.nchunks_local : mp.P_M;
            for (int n_rel = 0; n_rel < T.NbAll; n_rel++) {
                double part[kMarginalBlock / 32];
                for (int warp = 0; warp < kMarginalBlock / 32; warp++) {
                    double v[32];
                    for (int lane = 0; lane < 32; lane++) {
                        double acc = 0.0;
                        for (long long e = warp * 32 + lane; e < count; e += kMarginalBlock) acc += marginal_term (mp, kind, site, digit, n_rel, e);
                        v[lane] = acc;
                    }
                    part[warp] = butterfly_sum (v);
                }
                double total = 0.0;
                for (int warp = 0; warp < kMarginalBlock / 32; warp++) total += part[warp];
                bins[(size_t) row * Nb + plan.base + n_rel] += total;
            }
        }
    }
    return 0;
}
