// pce_analytic.cu -- kernel 2: exhaustive enumeration of the partition function (launch side).
//
// Replaces EnergyModel_CalculateZ / _CalculateProbabilitiesFromZ / _CalculateProbabilitiesAnalytically
// (extensions/csource/EnergyModel.c:252-281, 318-337, 342-371).  The reference recomputes every
// microstate energy from scratch (O(nsites^2/2) gathers per state, its own TODO at :264-267), stores
// nstates Boltzmann factors, and repeats all of it for every pH value.  Here:
//
//   (1) G_s(pH) = A_s + n_s * R T ln10 * pH, so states are BINNED BY PROTON COUNT n_s and one
//       enumeration at a reference pH* serves the whole pH grid:
//           Z(pH)    = sum_n  Zbin[n]    * 10^(-n (pH - pH*)),
//           p_a(pH)  = sum_n  Sbin[a][n] * 10^(-n (pH - pH*)) / Z(pH);
//       every bin carries its own scale exp(sigma(n)) (pce_enum_plan.h), so one pass serves a pH grid of any width;
//   (2) sites take roles and the Boltzmann factor of a state is a product of table entries: one DFMA per state on
//       register operands (pce_enum_core.cuh holds the CTA programs; this file launches them);
//   (3) bins of the instances of the M and F sites fall out of per-thread and per-chunk partial sums; the sites that
//       were looped get theirs from a second enumeration with the roles rotated.
//
// Every partial result is a plain sum with a common reference energy and common bin scales, so shards of the chunk range --
// on one GPU or many -- combine by addition (one all-reduce).  No floating-point atomics: results are reproducible.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pce_common.cuh"
#include "pce_enum_core.cuh"

using pce::fail;
using pce_enum::EnumParams;
using pce_enum::MarginalParams;
using pce_enum::kBlock;

namespace {

__global__ void __launch_bounds__ (kBlock) k_enum_static (const EnumParams p) {
    extern __shared__ __align__ (16) unsigned char smem[];
    const pce_enum::StaticCta cta (p, smem);
    const int tid = threadIdx.x;
    cta.phase0a (tid); __syncthreads ();
    cta.phase0b (tid); __syncthreads ();
    cta.phase1 (tid); __syncthreads ();
    cta.phase2 (tid); __syncthreads ();
    cta.phase3 (tid); __syncthreads ();
    cta.phase3b (tid); __syncthreads ();
    cta.phase4 (tid);
}

// Persistent CTAs of 128 threads: CTA b takes chunks b, b + gridDim.x, ... of this shard's range.  MODE 2 (two I x J register
// sets per thread) runs three CTAs per SM at up to 168 registers, the others four at 128.
template <int MODE>
__global__ void __launch_bounds__ (kBlock, MODE == 2 ? 3 : 4) k_enum (const EnumParams p) {
    extern __shared__ __align__ (16) unsigned char smem[];
    const pce_enum::EnumCta<MODE> cta (p, smem);
    pce_enum::EnumThread r;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cta.init (tid, r);
    __syncthreads ();
    for (long long c = blockIdx.x; c < p.nchunks_local; c += gridDim.x) {
        cta.phase0 (tid, p.chunk_begin + c); __syncthreads ();
        cta.phase1 (tid); __syncthreads ();
        cta.phase2 (tid); __syncthreads ();
        cta.phase3 (tid); __syncthreads ();
        cta.phase4 (tid, r); __syncthreads ();
        cta.accumulate (tid, r); __syncthreads ();
        cta.fold (tid);
        if (p.needF)
            for (int row = warp; row < p.NbCp; row += kBlock / 32) {
                const double total = pce::warp_sum (cta.row_partial (row, lane));
                if (lane == 0) cta.store_row (row, c, total);
            }
        // the next chunk's phase0 writes only the F digits (read after its barrier); the chunk bins are zeroed in phase1
    }
    cta.finish (tid, (int) blockIdx.x);
}

// Mtab[e] = sum over the CTAs of T[cta][e]: blockIdx.y sums a segment of CTAs into partial[segment][e]; the block that
// finishes last for a column group (a counter per blockIdx.x, reset by that block) adds the segments in segment order,
// so the result does not depend on which block that was.
__global__ void __launch_bounds__ (128) k_enum_reduce (const double *__restrict__ T, int ctas, int width, int per_segment, double *__restrict__ partial,
                                                       unsigned *__restrict__ counters, double *__restrict__ Mtab) {
    __shared__ int s_last;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * per_segment, c1 = min (ctas, c0 + per_segment);
    if (e < width) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int c = c0;
        for (; c + 3 < c1; c += 4) {
            a0 += T[(size_t) c * width + e]; a1 += T[(size_t) (c + 1) * width + e];
            a2 += T[(size_t) (c + 2) * width + e]; a3 += T[(size_t) (c + 3) * width + e];
        }
        for (; c < c1; c++) a0 += T[(size_t) c * width + e];
        partial[(size_t) blockIdx.y * width + e] = (a0 + a1) + (a2 + a3);
    }
    __threadfence ();
    __syncthreads ();
    if (threadIdx.x == 0) {
        const unsigned ticket = atomicAdd (&counters[blockIdx.x], 1u);
        s_last = ticket == gridDim.y - 1 ? 1 : 0;
        if (s_last) counters[blockIdx.x] = 0u;
    }
    __syncthreads ();
    if (s_last && e < width) {
        __threadfence ();
        double v = 0.0;
        for (unsigned s = 0; s < gridDim.y; s++) v += __ldcg (partial + (size_t) s * width + e);
        Mtab[e] = v;
    }
}

// blockIdx.x: Z | (M site, digit) | (F site, digit); blockIdx.y: bin.  Elements are spread over the threads in a fixed
// pattern and combined in a fixed order.
__global__ void __launch_bounds__ (pce_enum::kMarginalBlock) k_enum_marginals (const MarginalParams p) {
    constexpr int kThreads = pce_enum::kMarginalBlock;
    __shared__ double s_part[kThreads / 32];
    int kind, site, digit, row;
    if (!pce_enum::marginal_target (p, (int) blockIdx.x, &kind, &site, &digit, &row)) return;
    const int tid = threadIdx.x, n_rel = (int) blockIdx.y;
    const long long count = kind == 2 ? p.nchunks_local : p.P_M;
    double acc = 0.0;
    for (long long e = tid; e < count; e += kThreads) acc += pce_enum::marginal_term (p, kind, site, digit, n_rel, e);
    acc = pce::warp_sum (acc);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads ();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < kThreads / 32; w++) total += s_part[w];
        p.bins[(size_t) row * p.Nb + p.base + n_rel] += total;
    }
}

__global__ void k_tilt (int ninst, const double *__restrict__ g, const int32_t *__restrict__ protons, double mu_star, double *__restrict__ gt) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < ninst) gt[a] = g[a] - (double) protons[a] * mu_star;
}

// ---------------------------------------------------------------------------------------------------
// host side: plan cache, table upload, launches
// ---------------------------------------------------------------------------------------------------
pce_enum::ModelView view_of (const pce_model *m, int unfolded) {
    pce_enum::ModelView v;
    v.nsites = m->nsites; v.ninst = m->ninst;
    v.first = m->first.data (); v.last = m->last.data (); v.protons = m->protons.data ();
    v.g = unfolded ? m->gmodel.data () : m->gintr.data ();
    v.w = const_cast<pce_model *> (m)->host_wsym ();
    v.temperature = m->temperature; v.unfolded = unfolded;
    return v;
}

// all tables of one run packed into one byte blob (16-byte aligned pieces); offsets are keyed by the vector's address
struct Blob {
    std::vector<unsigned char> bytes;
    std::vector<std::pair<const void *, size_t>> where;
    template <typename T> void add (const std::vector<T> &v) {
        size_t o = (bytes.size () + 15) & ~(size_t) 15;
        bytes.resize (o + v.size () * sizeof (T) + 16, 0);
        if (!v.empty ()) std::memcpy (bytes.data () + o, v.data (), v.size () * sizeof (T));
        where.push_back (std::make_pair ((const void *) &v, o));
    }
};
struct BlobLocate {
    const Blob &blob;
    const unsigned char *base;
    template <typename T> const T *operator() (const std::vector<T> &v) const {
        for (const auto &entry : blob.where) if (entry.first == (const void *) &v) return (const T *) (base + entry.second);
        return nullptr;
    }
};
void pack_tables (const pce_enum::RunTables &T, Blob *blob) {
    blob->add (T.innerInst); blob->add (T.innerSite); blob->add (T.innerOff); blob->add (T.oInst); blob->add (T.idxM); blob->add (T.idxLo);
    blob->add (T.idxHi); blob->add (T.idxI); blob->add (T.idxJ); blob->add (T.mlo); blob->add (T.mhi); blob->add (T.fInst); blob->add (T.nrelM); blob->add (T.nrelI);
    blob->add (T.nrelO1); blob->add (T.nrelO2); blob->add (T.bO); blob->add (T.fRel); blob->add (T.fOff); blob->add (T.fRad); blob->add (T.fStride); blob->add (T.fShift);
    blob->add (T.D); blob->add (T.sigmaU); blob->add (T.digitM); blob->add (T.rowM); blob->add (T.rowMoff); blob->add (T.coverM); blob->add (T.coverF);
}

}  // namespace

// the plan of the last call, kept with the model: a titration curve asks for bins and then for their finish with the same
// arguments, and repeated curves of an unchanged model skip the planning and the table upload altogether
struct pce_enum_cache {
    int generation = -1, unfolded = -1;
    long long grid_slots = -1;
    int  max_mode = 2;
    bool on_device = false, static_done = false;      // static_done: the numeric tables of every run are on the device
    std::vector<double> ph;
    pce_enum::Plan plan;
    std::vector<pce_enum::RunTables> tables;
    std::vector<Blob> blobs;
    std::vector<size_t> blob_offset;
    size_t blob_total = 0;
};

void pce_enum_cache_free (pce_enum_cache *cache) { delete cache; }

namespace {

bool same_list (const std::vector<double> &a, const double *b, int n) {
    if ((int) a.size () != n) return false;
    for (int k = 0; k < n; k++) if (a[(size_t) k] != b[k]) return false;
    return true;
}

// plan (cached when the device copy of the model is current; the CPU-only finish of the tests plans afresh)
int get_plan (pce_model *m, const double *ph, int nph, int unfolded, long long grid_slots, bool with_tables, pce_enum_cache **out) {
    // test hooks: PCE_ENUM_GENERIC=1 forces the generic accumulation path, PCE_ENUM_MODE=0|1|2 caps the kernel variant
    int max_mode = std::getenv ("PCE_ENUM_GENERIC") != nullptr ? 0 : 2;
    if (const char *cap = std::getenv ("PCE_ENUM_MODE")) max_mode = std::max (0, std::min (2, std::atoi (cap)));
    pce_enum_cache *cache = m->enum_cache;
    const bool current = cache != nullptr && !m->dirty && cache->generation == m->generation && cache->unfolded == unfolded && same_list (cache->ph, ph, nph) &&
                         cache->max_mode == max_mode;
    if (current && (!with_tables || (cache->grid_slots == grid_slots && !cache->tables.empty ()))) { *out = cache; return PCE_OK; }
    if (cache == nullptr) { cache = new pce_enum_cache (); m->enum_cache = cache; m->enum_cache_free = pce_enum_cache_free; }
    const pce_enum::ModelView view = view_of (m, unfolded);
    std::string message;
    cache->generation = -1;
    cache->tables.clear (); cache->blobs.clear (); cache->blob_offset.clear (); cache->on_device = false; cache->static_done = false;
    if (pce_enum::make_plan (view, ph, nph, &cache->plan, &message, grid_slots, max_mode)) return fail (PCE_ERR_LIMIT, message);
    if (with_tables) {
        cache->tables.resize (cache->plan.runs.size ());
        cache->blobs.resize (cache->plan.runs.size ());
        size_t total = 0;
        for (size_t r = 0; r < cache->plan.runs.size (); r++) {
            if (pce_enum::build_tables (view, cache->plan, cache->plan.runs[r], &cache->tables[r], &message)) return fail (PCE_ERR_LIMIT, message);
            pack_tables (cache->tables[r], &cache->blobs[r]);
            cache->blob_offset.push_back (total);
            total += (cache->blobs[r].bytes.size () + 255) & ~(size_t) 255;
        }
        cache->blob_total = total;
    }
    cache->generation = m->dirty ? -1 : m->generation;
    cache->unfolded = unfolded; cache->grid_slots = grid_slots; cache->max_mode = max_mode;
    cache->ph.assign (ph, ph + nph);
    *out = cache;
    return PCE_OK;
}

int resident_ctas (void (*kernel) (const EnumParams), size_t smem, int *out) {
    int value = 0;
    PCE_CUDA (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    PCE_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&value, kernel, kBlock, smem));
    *out = value;
    return PCE_OK;
}

// One enumeration per planned run (roles rotated so that every site is covered), accumulated into d_bins.
int run_plan (pce_model *m, pce_enum_cache *cache, int unfolded, int shard, int nshards, int sms, double *d_bins) {
    const pce_enum::Plan &plan = cache->plan;
    const size_t nbins_total = (size_t) (1 + m->ninst) * plan.Nb;
    PCE_CUDA (cudaMemsetAsync (d_bins, 0, nbins_total * sizeof (double), m->stream));
    // scratch lives in the model (grow-only): repeated curves do not pay cudaMalloc / cudaFree
    pce::DeviceBuffer<double> &d_gt = m->s_gt, &d_T = m->s_T, &d_Mtab = m->s_Mtab, &d_Ftab = m->s_Ftab, &d_numeric = m->s_xs, &d_partial = m->s_seg;
    pce::DeviceBuffer<int32_t> &d_blob = m->s_ms;
    PCE_CUDA (d_gt.resize (m->ninst));
    if (d_numeric.count < plan.runs.size () * pce_enum::kNumericDoubles) cache->static_done = false;
    PCE_CUDA (d_numeric.resize (plan.runs.size () * pce_enum::kNumericDoubles));
    if (m->s_cnt.count < 1024) {
        PCE_CUDA (m->s_cnt.resize (1024));
        PCE_CUDA (cudaMemsetAsync (m->s_cnt.ptr, 0, 1024 * sizeof (unsigned), m->stream));
    }
    if (!cache->on_device) {
        if (d_blob.count * sizeof (int32_t) < cache->blob_total) cache->on_device = false;
        PCE_CUDA (d_blob.resize ((cache->blob_total + 3) / 4));
        for (size_t r = 0; r < cache->blobs.size (); r++)
            PCE_CUDA (cudaMemcpyAsync ((unsigned char *) d_blob.ptr + cache->blob_offset[r], cache->blobs[r].bytes.data (), cache->blobs[r].bytes.size (),
                                       cudaMemcpyHostToDevice, m->stream));
        cache->on_device = cache->generation >= 0;
    }
    const double mu_star = pce::mu_of (m->temperature, plan.ph_star);
    k_tilt<<<(m->ninst + 127) / 128, 128, 0, m->stream>>> (m->ninst, unfolded ? m->d_gmodel.ptr : m->d_gintr.ptr, m->d_protons.ptr, mu_star, d_gt.ptr);
    pce::count_launch ();
    size_t limit = 0;
    {
        int value = 0;
        PCE_CUDA (cudaDeviceGetAttribute (&value, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
        limit = (size_t) value;
    }
    for (size_t r = 0; r < plan.runs.size (); r++) {
        const pce_enum::RunTables &T = cache->tables[r];
        EnumParams p;
        MarginalParams mp;
        std::memset (&p, 0, sizeof (p)); std::memset (&mp, 0, sizeof (mp));
        const BlobLocate locate = { cache->blobs[r], (const unsigned char *) d_blob.ptr + cache->blob_offset[r] };
        pce_enum::bind_run (T, p, mp, locate, d_numeric.ptr + r * pce_enum::kNumericDoubles);
        p.ninst = m->ninst; p.ldw = m->ldw; p.unfolded = unfolded; p.gt = d_gt.ptr; p.w = m->d_w.ptr;
        p.beta = 1.0 / (PCE_R * m->temperature); p.gref = plan.gref;
        // this shard's slice of the chunk range
        const long long c0 = (long long) ((__int128) T.nchunks * shard / nshards), c1 = (long long) ((__int128) T.nchunks * (shard + 1) / nshards);
        p.chunk_begin = c0; p.nchunks_local = c1 - c0; p.ldF = p.nchunks_local;
        if (p.nchunks_local <= 0) continue;
        if (p.nchunks_local > (1LL << 28)) return fail (PCE_ERR_LIMIT, "too many chunks in one shard of the enumeration");
        const pce_enum::EnumLayout layout (p);
        const pce_enum::StaticLayout static_layout (p);
        if (layout.total + 1024 > limit || static_layout.total + 1024 > limit) return fail (PCE_ERR_LIMIT, "too many proton-count bins for shared memory");
        void (*kernel) (const EnumParams) = T.mode == 2 ? k_enum<2> : T.mode == 1 ? k_enum<1> : k_enum<0>;
        int per_sm = 0;
        int status = resident_ctas (kernel, layout.total, &per_sm);
        if (status != PCE_OK) return status;
        if (per_sm < 1) return fail (PCE_ERR_LIMIT, "the enumeration kernel does not fit on an SM");
        const int ctas = (int) std::min<long long> ((long long) sms * per_sm, p.nchunks_local);
        PCE_CUDA (d_T.resize ((size_t) ctas * T.NbP * kBlock));
        PCE_CUDA (d_Mtab.resize ((size_t) T.NbP * kBlock));
        PCE_CUDA (d_Ftab.resize ((size_t) T.NbAll * (size_t) p.ldF));
        p.T = d_T.ptr; p.FtabT = d_Ftab.ptr;
        if (p.needF) PCE_CUDA (cudaMemsetAsync (d_Ftab.ptr, 0, (size_t) T.NbAll * (size_t) p.ldF * sizeof (double), m->stream));
        if (!cache->static_done) {      // the numeric tables depend on W and the roles only: kept while the plan is
            PCE_CUDA (cudaFuncSetAttribute (k_enum_static, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) static_layout.total));
            k_enum_static<<<1, kBlock, static_layout.total, m->stream>>> (p);
            PCE_CUDA (cudaGetLastError ());
            pce::count_launch ();
        }
        kernel<<<(unsigned) ctas, kBlock, layout.total, m->stream>>> (p);
        PCE_CUDA (cudaGetLastError ());
        const int width = T.NbP * kBlock, columns = (width + 127) / 128;
        const int per_segment = 16, segments = (ctas + per_segment - 1) / per_segment;
        if (columns > 1024) return fail (PCE_ERR_LIMIT, "too many proton-count bins for the reduction");
        PCE_CUDA (d_partial.resize ((size_t) segments * width));
        k_enum_reduce<<<dim3 ((unsigned) columns, (unsigned) segments), 128, 0, m->stream>>> (d_T.ptr, ctas, width, per_segment, d_partial.ptr, m->s_cnt.ptr, d_Mtab.ptr);
        PCE_CUDA (cudaGetLastError ());
        mp.Nb = plan.Nb; mp.base = plan.base; mp.write_z = r == 0 ? 1 : 0; mp.grid = ctas;
        mp.chunk_begin = c0; mp.nchunks_local = p.nchunks_local; mp.ldF = p.ldF;
        mp.Mtab = d_Mtab.ptr; mp.FtabT = d_Ftab.ptr; mp.bins = d_bins;
        k_enum_marginals<<<dim3 ((unsigned) pce_enum::marginal_rows (T), (unsigned) T.NbAll), pce_enum::kMarginalBlock, 0, m->stream>>> (mp);
        PCE_CUDA (cudaGetLastError ());
        pce::count_launch (3);
    }
    cache->static_done = cache->generation >= 0;
    cudaError_t err = cudaStreamSynchronize (m->stream);
    if (err != cudaSuccess) return fail (PCE_ERR_CUDA, std::string ("analytic enumeration: ") + cudaGetErrorString (err));
    return PCE_OK;
}

int radix (const pce_model *m, int s) { return m->last[s] - m->first[s] + 1; }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" long long pce_analytic_bins_size (const pce_model *m) {
    if (m == nullptr) return -1;
    return (long long) (1 + m->ninst) * (pce_enum::max_protons_total (view_of (m, 0)) + 1);
}

static int check_analytic_preconditions (pce_model *m, int unfolded, double max_states, double *nstates) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    int status = m->sites_ok ();
    if (status != PCE_OK) return status;
    if (!unfolded && !m->symmetrized) return fail (PCE_ERR_STATE, "First calculate electrostatic energies.");
    double n = 1.0;
    for (int i = 0; i < m->nsites; i++) n *= (double) radix (m, i);
    const double engine_limit = 17592186044416.0;   // 2^44
    if (max_states <= 0.0 || max_states > engine_limit) max_states = engine_limit;
    if (n > max_states) {
        char text[128];
        snprintf (text, sizeof (text), "Maximum number of states for analytic treatment (%.0f) exceeded.", max_states);
        return fail (PCE_ERR_LIMIT, text);
    }
    if (nstates) *nstates = n;
    return PCE_OK;
}

extern "C" int pce_analytic_bins (pce_model *m, const double *ph, int nph, int unfolded, int shard, int nshards, double max_states, double *d_bins) {
    int status = check_analytic_preconditions (m, unfolded, max_states, nullptr);
    if (status != PCE_OK) return status;
    if (nph < 1 || shard < 0 || nshards < 1 || shard >= nshards) return fail (PCE_ERR_VALUE, "invalid pH list or shard");
    status = m->upload ();
    if (status != PCE_OK) return status;
    int sms = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, m->device));
    pce_enum_cache *cache = nullptr;
    status = get_plan (m, ph, nph, unfolded, (long long) sms * 3 * nshards, true, &cache);      // three resident CTAs per SM on every shard
    if (status != PCE_OK) return status;
    return run_plan (m, cache, unfolded, shard, nshards, sms, d_bins);
}

// Probabilities (and ln Z) at every pH value from the scaled bins: bin n stands for
//   sum over the states with n protons of exp (-beta (E(pH*) - gref)) * exp (sigma[n]),
// and moving from pH* to pH multiplies it by exp (-beta n (mu* - mu)).  Extended precision, logarithmic weights.
extern "C" int pce_analytic_finish (pce_model *m, const double *bins, const double *ph, int nph, int unfolded, double gzero,
                                    double *out_p, double *out_lnz) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (nph < 1) return fail (PCE_ERR_VALUE, "invalid pH list");
    pce_enum_cache *cache = nullptr;
    int status = get_plan (m, ph, nph, unfolded, 0, false, &cache);
    if (status != PCE_OK) return status;
    const pce_enum::Plan &plan = cache->plan;
    const int Nb = plan.Nb;
    const long double beta = 1.0L / ((long double) PCE_R * m->temperature);
    const long double mu_star = pce::mu_of (m->temperature, plan.ph_star);
    // Logarithmic weights with ~32 digits (pairs of doubles: the range is that of the bin scales, e^(+-700) and beyond, and the weights
    // must come out to 1e-16 relative); the weighted sums themselves are sums of at most ~100 positive doubles <= 1 and stay in double
    // precision.  (x87 long double arithmetic and expl took 75 us per 41-point curve here; this takes 10.)
    struct Pair { double hi, lo; };
    auto two_sum = [] (double a, double b) -> Pair { const double s = a + b, bb = s - a; return { s, (a - (s - bb)) + (b - bb) }; };
    std::vector<double> logb (Nb, 0.0), weight (Nb), whi (Nb), wlo (Nb);
    for (int n = 0; n < Nb; n++) if (bins[n] > 0.0) logb[n] = std::log (bins[n]);          // (only picks the shift: any value near the maximum does)
    std::vector<double> wn ((size_t) Nb * nph);          // normalised weights [bin][pH]: pH innermost, so the sums below vectorise
    for (int q = 0; q < nph; q++) {
        const long double c = beta * (mu_star - (long double) pce::mu_of (m->temperature, ph[q]));      // log-weight of one more proton
        const double c_hi = (double) c, c_lo = (double) (c - (long double) c_hi);
        double shift = -1.0e300;
        for (int n = 0; n < Nb; n++) {
            // logw = -sigma[n] - n c
            const double p_hi = (double) n * c_hi, p_lo = std::fma ((double) n, c_hi, -p_hi) + (double) n * c_lo;
            Pair w = two_sum (-plan.sigma[(size_t) n], -p_hi);
            w.lo -= p_lo;
            whi[n] = w.hi; wlo[n] = w.lo;
            if (bins[n] > 0.0) shift = std::max (shift, logb[n] + w.hi);
        }
        double z = 0.0;
        for (int n = 0; n < Nb; n++) {
            double value = 0.0;
            if (bins[n] > 0.0) {
                Pair a = two_sum (whi[n], -shift);
                a.lo += wlo[n];
                const double hi = a.hi + a.lo, lo = a.lo - (hi - a.hi);
                value = hi < -745.0 ? 0.0 : hi > 690.0 ? 1.0e300 : std::exp (hi) * (1.0 + lo);       // bins[n] * weight[n] <= ~1
            }
            weight[n] = value;
            z += bins[n] * value;
        }
        if (!(z > 0.0)) return fail (PCE_ERR_STATE, "partition function underflowed; split the pH grid");
        const double inv = 1.0 / z;
        for (int n = 0; n < Nb; n++) wn[(size_t) n * nph + q] = weight[n] * inv;
        if (out_lnz) out_lnz[q] = (double) (logl ((long double) z) + (long double) shift - beta * ((long double) plan.gref - gzero));
    }
    std::vector<double> acc (nph);
    for (int a = 0; a < m->ninst; a++) {
        const double *row = bins + (size_t) (1 + a) * Nb;
        std::fill (acc.begin (), acc.end (), 0.0);
        double *__restrict__ sum = acc.data ();
        for (int n = 0; n < Nb; n++) {
            const double r = row[n];
            const double *__restrict__ w = &wn[(size_t) n * nph];
            for (int q = 0; q < nph; q++) sum[q] += r * w[q];
        }
        if (out_p) for (int q = 0; q < nph; q++) out_p[(size_t) q * m->ninst + a] = acc[q];
        m->prob[a] = acc[nph - 1];
    }
    return PCE_OK;
}

// log-scales of the proton-count bins that pce_analytic_bins would use for this pH list (tests: bins from another source,
// e.g. the CPU emulator of the kernel, can be finished by pce_analytic_finish)
extern "C" int pce_analytic_bin_scales (pce_model *m, const double *ph, int nph, int unfolded, double *sigma) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (nph < 1) return fail (PCE_ERR_VALUE, "invalid pH list");
    pce_enum_cache *cache = nullptr;
    int status = get_plan (m, ph, nph, unfolded, 0, false, &cache);
    if (status != PCE_OK) return status;
    for (int n = 0; n < cache->plan.Nb; n++) sigma[n] = cache->plan.sigma[(size_t) n];
    return PCE_OK;
}

extern "C" int pce_analytic (pce_model *m, const double *ph, int nph, int unfolded, double gzero, double max_states, double *out_p, double *out_lnz) {
    int status = check_analytic_preconditions (m, unfolded, max_states, nullptr);
    if (status != PCE_OK) return status;
    PCE_CUDA (cudaSetDevice (m->device));
    const long long n = pce_analytic_bins_size (m);
    pce::DeviceBuffer<double> d_bins;
    PCE_CUDA (d_bins.resize ((size_t) n));
    status = pce_analytic_bins (m, ph, nph, unfolded, 0, 1, max_states, d_bins.ptr);
    std::vector<double> bins ((size_t) n);
    if (status == PCE_OK) {
        cudaError_t err = cudaMemcpy (bins.data (), d_bins.ptr, (size_t) n * sizeof (double), cudaMemcpyDeviceToHost);
        if (err != cudaSuccess) status = fail (PCE_ERR_CUDA, std::string ("analytic readback: ") + cudaGetErrorString (err));
    }
    d_bins.release ();
    if (status != PCE_OK) return status;
    return pce_analytic_finish (m, bins.data (), ph, nph, unfolded, gzero, out_p, out_lnz);
}

extern "C" int pce_analytic_gmin (pce_model *m, double ph, int unfolded, double max_states, double *gmin) {
    // Exact minimum over all states, evaluated with kernel 1 on batches of enumerated states.
    // (Needed only to reproduce the reference's Gmin-relative Z, EnergyModel.c:262-275.)
    double nstates = 0.0;
    int status = check_analytic_preconditions (m, unfolded, max_states, &nstates);
    if (status != PCE_OK) return status;
    if (nstates > 268435456.0) return fail (PCE_ERR_LIMIT, "pce_analytic_gmin supports at most 2^28 states");
    const long long total = (long long) nstates, batch = 1 << 20;
    std::vector<int32_t> states;
    std::vector<double>  out;
    double best = DBL_MAX;
    std::vector<int> digit (m->nsites, 0);
    for (long long s0 = 0; s0 < total; s0 += batch) {
        const long long nb = std::min (batch, total - s0);
        states.resize ((size_t) nb * m->nsites);
        out.resize ((size_t) nb);
        for (long long s = 0; s < nb; s++) {
            for (int i = 0; i < m->nsites; i++) states[(size_t) s * m->nsites + i] = m->first[i] + digit[i];
            for (int i = 0; i < m->nsites; i++) {     // StateVector_Increment, StateVector.c:370-384
                if (digit[i] + 1 < radix (m, i)) { digit[i]++; break; }
                digit[i] = 0;
            }
        }
        status = pce_energy (m, states.data (), nb, &ph, 1, unfolded, out.data ());
        if (status != PCE_OK) return status;
        for (long long s = 0; s < nb; s++) best = std::min (best, out[(size_t) s]);
    }
    *gmin = best;
    return PCE_OK;
}
