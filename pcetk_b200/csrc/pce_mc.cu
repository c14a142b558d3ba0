// pce_mc.cu -- kernel 3: Metropolis Monte Carlo over the (pH x chain) grid.
//
// Replaces extensions/csource/MCModelDefault.c (Move :90-122, DoubleMove :127-172, MCScan :178-209,
// UpdateProbabilities :215-226, FindPairs :231-288, Production/Equilibration :298-327).
//
// Algorithm (specified operation for operation by oracle_mc_philox_chain in oracle/pce_oracle.c):
//   * every chain keeps cached fields RELATIVE TO THE FIRST INSTANCE OF EACH SITE, g[r(k)] = h[k] - h[first], with
//     h[k] = Gintr'[k] + sum_j W[k][a_j] (only differences between instances of one site ever enter a decision):
//     ninst - nsites values per chain; a proposal costs O(1) (dG = G(new) - G(old)), an accepted move one coalesced
//     add of a precomputed difference row E = wr[hi] - wr[lo] -- instead of the reference's 2*nsites strided
//     gathers per trial (MCModelDefault.c:109-111);
//   * random numbers are Philox4x32-10 blocks addressed by (seed, pH index, chain, trial), two trials
//     per block, so any chain can be replayed in isolation and nothing has to be checkpointed;
//   * occupancies are residence times (scan of last change per site) instead of nsites increments per scan.
//
// Two kernels:
//   k_mc_thread  one chain per THREAD; difference rows, cross-term table and per-chain state in shared
//                memory.  For models whose tables fit on chip next to >= 8 warps of chains (lysozyme,
//                defensin, ...).  Accepted moves are applied warp-cooperatively so that the row add
//                never runs with 1/32 lanes.
//   k_mc_warp    one chain per WARP, fields in shared memory; the lanes evaluate 32 consecutive trials
//                speculatively in parallel, the first accepted one ends the round.  Difference rows
//                streamed from L1/L2 with 128-bit loads.  For everything else (ifp20; the 1000-site
//                synthetic, whose 60 MB of difference rows are L2-resident).
#include <algorithm>
#include <type_traits>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pce_common.cuh"

using pce::fail;

struct pce_mc {
    pce_model *model = nullptr;
    double     limit = 2.0;
    int        nequil = 500, nprod = 20000, nchains = 1, kernel = 0;
    uint64_t   seed = 1;
    std::vector<int32_t> pair_a, pair_b;
    std::vector<double>  pair_wmax;
    pce::DeviceBuffer<uint32_t> d_pair_packed;
    pce::DeviceBuffer<uint32_t> d_pair_x;
    pce::DeviceBuffer<double>   d_cross;
    bool cross_valid = false, cross_overflow = false;   // overflow: the table exceeds 2^24 entries -> 255-register variant only
    int  bound_generation = -1;
    pce::DeviceBuffer<double>  d_ph, d_esum;
    pce::DeviceBuffer<unsigned long long> d_counts;
    pce::DeviceBuffer<long long> d_counters;
    // tables of the relative-field formulation, built on the device from the uploaded beta-scaled W
    pce::DeviceBuffer<double>  d_wr;                 // wr[a][r(k)] = wb[a][k] - wb[a][first of k's site]   [ninst][ldr]
    pce::DeviceBuffer<double>  d_dw;                 // difference rows E = wr[hi] - wr[lo] of every instance pair of a site [nrows][ldd]
    pce::DeviceBuffer<int32_t> d_dbase;              // first difference row of each site
    pce::DeviceBuffer<int32_t> d_rk, d_rf;           // instance / first instance of the site of relative slot r
    int  nrel = 0, ldr = 0, nrows = 0;
    int  rel_generation = -1;                         // model generation d_wr, d_rk, d_rf were built from
    int  dw_generation = -1, dw_ldd = 0;              // model generation / row stride / slot layout the difference rows were built for
    bool dw_abs = false, dw_valid = false;
    size_t ncross = 0;
    bool pairs_dirty = true;
};

namespace {

struct McParams {
    int nsites, ninst, npairs, ldw;
    const int32_t *first, *nsi, *protons;
    const uint32_t *pair_packed;            // site a | site b << 16
    const uint32_t *pair_x;                 // cross-table base of the pair | ceil (4 * bound of |cross term|) << 24
    const double   *cross;                  // precomputed pair cross terms (k_mc_warp: nullptr = gather the four W entries)
    const double  *gintr, *w, *wb, *ph;      // wb = beta * W (units of RT)
    int       nph, ph_index_offset, nchains, nequil, nprod;
    long long chain_offset;
    uint32_t  k0, k1;
    double    temperature;
    unsigned long long *counts;     // [nph*ninst]
    long long          *counters;   // [nph*4]
    double             *esum;       // [nph]
    // optional per-chain outputs (parity tests)
    long long *chain_counts;        // [nph*nchains*ninst]
    int32_t   *chain_state;         // [nph*nchains*nsites]
    double    *chain_energy, *chain_esum;   // [nph*nchains]
    long long *chain_counters;      // [nph*nchains*4]
    double    *trajectory;          // [nprod] energies of chain 0 of pH 0
    long long *log_rows;            // [(nequil/f + nprod/f) * 5] scan, moves, accepted, flips, accepted per block of f scans (chain 0, pH 0)
    int        log_frequency;
    // relative fields (see the file header): nrel = ninst - nsites slots per chain
    int            nrel, ldr, nrows, ncross;
    const double  *wr;              // [ninst][ldr] wr[a][r(k)] = wb[a][k] - wb[a][first of k's site]
    const int32_t *rk, *rf;         // [nrel] instance of slot r / first instance of its site
    const int32_t *site_of;         // [ninst] site of every instance
    const double  *dw;              // difference rows E = wr[hi] - wr[lo] [nrows][ldd] (nullptr: stream the two wr rows themselves)
    const int32_t *dbase;           // [nsites] first difference row of the site
    int            ldd;             // row stride of dw (warp kernel: the padded field length of the launch)
    // chain-per-thread kernel: the shared-memory layout, computed on the host (ThreadLayout) so that the kernel can
    // re-read an offset from the constant bank instead of keeping it in a register or recomputing it
    uint32_t       tl_e, tl_cross, tl_h, tl_eng, tl_counts, tl_nact, tl_since, tl_move, tl_erow, tl_state;
    int            tl_hstride, tl_estride, tl_kslots;
    int            ph_minor;        // linearisation of the (pH, chain) grid: 1 = g = chain * nph + q (consecutive threads take consecutive pH
                                    // values, so every warp and block carries the same mix of cheap and expensive pH values), 0 = g = q * nchains + chain
    // chain-per-warp kernel: the shared-memory layout (WarpLayout), likewise
    uint32_t       wl_shared, wl_per_warp, wl_since, wl_state, wl_ring, wl_since16;
    int            wl_ldh;
    uint32_t       nmoves;          // nsites + npairs
    unsigned long long ntrials;     // (nequil + nprod) * nmoves (< 2^32, checked at launch)
    uint32_t       nmoves_magic;    // floor (2^32 / nmoves) + 1 when a round can cross several scan ends (2 <= nmoves <= 64), else 0
};

// shared-memory carve-up of k_mc_thread; identical on host (sizing) and device (pointers)
struct ThreadLayout {
    size_t off_e, off_cross, off_h, off_eng, off_counts, off_nact, off_since, off_move, off_erow, off_state, total;
    int    hstride, estride, kslots;
    __host__ __device__ ThreadLayout (int nsites, int ninst, int npairs, int nrel, int nrows, int ncross, int block, int nchains, int nph, int since_bytes, bool ph_minor,
                                      bool energy) {
        hstride = (nrel > 0 ? nrel : 1) | 1;       // odd: lanes of a warp fall on distinct bank pairs
        estride = nrel > 0 ? nrel : 1;
        // occupancy counters per pH value a block can meet: a block covers `block` consecutive (pH, chain) pairs of the
        // linear grid -- pH-minor (g = chain * nph + q): min (block, nph) values; chain-major (g = q * nchains + chain):
        // at most (block - 1) / nchains + 2
        kslots = ph_minor ? block : (block - 1) / (nchains > 0 ? nchains : 1) + 2;
        if (kslots > nph) kslots = nph;
        size_t o = 0;
        off_e = o;      o += (size_t) (nrows > 0 ? nrows : 1) * estride * 8;      // difference rows E
        off_cross = o;  o += (size_t) (ncross > 0 ? ncross : 1) * 8;              // pair cross terms
        off_h = o;      o += (size_t) block * hstride * 8;                        // relative fields, one row per chain
        // tracked energy and its sum over production scans, per chain -- only when the caller asked for energy sums (the
        // reference's quiet path returns occupancies only, MCModelDefault.pyx:60-78): without them a lysozyme chain takes 247
        // instead of 263 bytes and 896 instead of 768 chains fit
        off_eng = o;    o += energy ? (size_t) block * 16 : 0;
        // occupancy counters of the NON-FIRST instances only (u32: block * nprod < 2^32 is checked on the host); the first instance of
        // a site gets (chains of the slot) * nprod - (its siblings' counts) when the block flushes
        // (+ one slot that the idle threads of the last block count into: they run a chain like everybody else, so the trial loop
        // carries no "active" flag)
        off_counts = o; o += (size_t) (kslots + 1) * (nrel > 0 ? nrel : 1) * 4;
        off_nact = o;   o += (size_t) kslots * 4;
        // residence stamps (u16 when nprod < 65536) and states are site-major; rows of block + 4 bytes (stamps: words) so that lanes
        // working on different sites fall into different banks (a row of `block` bytes is a multiple of 128: 4-way conflicts)
        off_since = o;  o += (size_t) nsites * (block * since_bytes + 4);
        o = (o + 15) & ~(size_t) 15;
        // one 16-byte entry per move (sites, then pairs), FOUR interleaved copies: lane l reads copy l & 3, so the random entries of a
        // quarter-warp's 128-bit loads fall into at most two instead of up to eight entries per bank group
        off_move = o;   o += (size_t) (nsites + npairs) * 64;
        off_erow = o;   o += (size_t) (nrows > 0 ? nrows : 1) * 8;      // per (current, drawn) instance of a site: its difference row and direction
        off_state = o;  o += (size_t) nsites * (block + 4);
        total = (o + 15) & ~(size_t) 15;
    }
};

__device__ __forceinline__ double ph_to_mu (double temperature, double ph) {
    // -R*T*ln10*pH, left to right (EnergyModel.c:223 / MCModelDefault.c:114)
    return __dmul_rn (__dmul_rn (__dmul_rn (-PCE_R, temperature), PCE_LN10), ph);
}

// MCModelDefault_Metropolis (MCModelDefault.c:71-85) on x = dG/RT: accept if x < 0; reject if x > 500; else accept
// iff u < exp(-x), u = w1 * 2^-32.  An fp32 exp screens: lanes it cannot decide with a 1e-3 relative margin (and the
// corner cases) take the fp64 path, so the decision always equals the fp64 rule.
__device__ __noinline__ bool metropolis_exact (double x, uint32_t w1) {
    if (x < 0.0) return true;
    if (-x < PCE_TOO_SMALL) return false;
    return ((double) w1 * (1.0 / 4294967296.0)) < exp (-x);
}

// (called by whole warps.  ex2.approx.ftz flushes results below 2^-126 to zero: there u < exp(-x) can only hold for u = 0, which
// lands in the exact path through `uf <= 0`, and so does x > 500, where the exact rule rejects.)
__device__ __forceinline__ bool metropolis (double x, uint32_t w1) {
    const float uf = __uint2float_rn (w1) * 2.3283064365386963e-10f;
    float ef;
    asm ("ex2.approx.ftz.f32 %0, %1;" : "=f"(ef) : "f"((float) x * -1.4426950408889634f));
    bool accept = (x < 0.0) || (uf < ef * 0.999f);
    const bool unsure = !accept && (uf <= ef * 1.001f);
    if (__any_sync (0xffffffffu, unsure)) {
        if (unsure) accept = metropolis_exact (x, w1);
    }
    return accept;
}

__device__ __forceinline__ double lds_f64 (uint32_t address) {
    double v;
    asm volatile ("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(address) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64 (uint32_t address, double v) {
    asm volatile ("st.shared.f64 [%0], %1;" : : "r"(address), "d"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8 (uint32_t address) {
    uint32_t v;
    asm volatile ("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(address) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8 (uint32_t address, uint32_t v) {
    asm volatile ("st.shared.u8 [%0], %1;" : : "r"(address), "r"(v) : "memory");
}
template <typename T> __device__ __forceinline__ uint32_t lds_stamp (uint32_t address) {
    uint32_t v;
    if (sizeof (T) == 2) asm volatile ("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(address) : "memory");
    else asm volatile ("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(address) : "memory");
    return v;
}
template <typename T> __device__ __forceinline__ void sts_stamp (uint32_t address, uint32_t v) {
    if (sizeof (T) == 2) asm volatile ("st.shared.u16 [%0], %1;" : : "r"(address), "r"(v) : "memory");
    else asm volatile ("st.shared.u32 [%0], %1;" : : "r"(address), "r"(v) : "memory");
}
__device__ __forceinline__ double flip_sign (double v, uint32_t mask) {
    return __hiloint2double (__double2hiint (v) ^ (int) mask, __double2loint (v));
}

// ---------------------------------------------------------------------------------------------------
// k_mc_thread.  The (pH, chain) grid is linearised pH-minor (g = chain * nph + q: consecutive threads take consecutive pH
// values, so every warp and block carries the same mix of cheap and expensive pH values; chain-major when the per-block
// occupancy counters of all pH values would not fit) and cut into blocks of blockDim.x threads; every thread carries its own
// pH.  Energies are in units of RT, the intrinsic term is fused into the cached fields and the fields are relative to the
// first instance of each site, so a single-move proposal is at most two shared-memory loads and one subtraction.  Trials are
// processed two per Philox block (no prefetch: the registers were worth more).  The model lives in shared memory as the
// difference rows E (one per instance pair of a site) and the pair cross-term table: W itself is never read by the trial loop.
// ---------------------------------------------------------------------------------------------------
// A trial starts with ONE 128-bit load of its move-table entry (both sites, their instance counts and field bases, the pair's
// cross-term base, the row-table base); states are LOCAL instance indices; the difference row of an accepted move comes from a
// row table indexed by (current, draw); every per-thread array is addressed with explicit 32-bit ld / st.shared.
// Register budgets: every SM sub-partition holds 16384 registers and a quarter of the block's warps, so a block of
// 512 / 640 / 768 / 896 / 1024 threads may use 128 / 96 / 80 / 72 / 64 registers per thread (__launch_bounds__ would round
// the block up to the next multiple of 256 threads and allow fewer).  Uncapped, the short-field build takes 72 registers and
// 272 instructions in the trial loop (6.0 warp-instructions per trial); under a smaller cap ptxas rematerialises loop invariants
// inside the loop (310 at 64), so the launch code picks the uncapped build whenever its registers fit the block (896 threads for
// the fixtures) and the capped ones otherwise.  The loop state is on a diet: no Philox prefetch, a 32-bit scan counter, the
// tracked energy (ENERGY builds) in shared memory, the layout passed in the kernel parameters.
template <bool PER_CHAIN, typename SinceT, int MAX_REGS, int SHORT_FIELDS, bool ENERGY>      // SHORT_FIELDS: 0 nrel > 32, 1 nrel <= 32, 2 nrel <= 16
__global__ void __maxnreg__ (MAX_REGS) k_mc_thread (const McParams p) {
    static_assert (ENERGY || !PER_CHAIN, "per-chain outputs include the energy");
    extern __shared__ __align__ (16) unsigned char smem[];
    const int B = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp_base = tid & ~31;
    double             *s_e      = (double *) (smem + p.tl_e);
    double             *s_cross  = (double *) (smem + p.tl_cross);
    double             *s_h      = (double *) (smem + p.tl_h);
    double             *my_eng   = (double *) (smem + p.tl_eng) + threadIdx.x;      // [0]: tracked energy G/RT, [B]: its sum over production scans (ENERGY)
    uint32_t           *s_counts = (uint32_t *) (smem + p.tl_counts);   // [kslots][nrel]: non-first instances
    uint32_t           *s_nact   = (uint32_t *) (smem + p.tl_nact);     // [kslots]: chains of this block per pH slot
    SinceT             *s_since  = (SinceT *) (smem + p.tl_since);
    // one entry per move, fetched with ONE 128-bit load per trial: x = site a | site b << 16 (single move: b = a),
    // y = (n_a - 1) | g_a << 16, z = (n_b - 1) | g_b << 16, w = base of the pair in the cross-term table (15 bits) | base of
    // site a in the row table << 15 (16 bits) | invalid << 31 (a site with one instance).  g = 8 * (first - site): local index l >= 1 of the site owns the field slot at byte g + 8 l - 8.
    uint4              *s_move   = (uint4 *) (smem + p.tl_move);             // entry i of copy c: s_move[4 * i + c]
    // difference row of every single-site move: entry 2 * (first row of the site) + current * (n - 1) + draw -> the row (its shared-memory
    // address in the short-field build) | 1 << 31 when the move runs against the row's direction (new < old)
    uint32_t           *s_rowtab = (uint32_t *) (smem + p.tl_erow);
    uint8_t            *s_state  = smem + p.tl_state;
    struct { int hstride, estride, kslots; } L = { p.tl_hstride, p.tl_estride, p.tl_kslots };

    const int nsites = p.nsites, ninst = p.ninst, nrel = p.nrel;
    const long long total = (long long) p.nph * p.nchains;
    const long long g0    = (long long) blockIdx.x * B;             // first (pH, chain) pair of this block
    const long long gidx  = g0 + tid;
    const bool active = gidx < total;
    const long long gclamped = active ? gidx : total - 1;
    const int q      = (int) (p.ph_minor ? gclamped % p.nph : gclamped / p.nchains);      // pH point of this thread
    const int local  = (int) (p.ph_minor ? gclamped / p.nph : gclamped % p.nchains);      // chain within this launch
    // first pH value of the block's counter slots (pH-minor: the values wrap around, all of them when the block is large)
    const int q_base = p.ph_minor ? (B >= p.nph ? 0 : (int) (g0 % p.nph)) : (int) (g0 / p.nchains);
    const uint32_t chain = (uint32_t) (p.chain_offset + local);
    const uint32_t phidx = (uint32_t) (p.ph_index_offset + q);
    const int my_slot = q - q_base + (q < q_base ? p.nph : 0);
    uint32_t *my_counts = s_counts + (size_t) (active ? my_slot : L.kslots) * (nrel > 0 ? nrel : 1);      // indexed by the relative slot r(k) = k - site - 1
    // per-thread arrays as 32-bit shared-memory addresses (one register each instead of a 64-bit pointer; explicit ld/st.shared, so
    // that no generic-to-shared window arithmetic is left in the trial loop)
    const uint32_t s_base = (uint32_t) __cvta_generic_to_shared (smem);
    const uint32_t o_state = s_base + p.tl_state + (uint32_t) tid, o_since = s_base + p.tl_since + (uint32_t) tid * (uint32_t) sizeof (SinceT);
    const uint32_t Bs = (uint32_t) B + 4u, Bt = (uint32_t) B * (uint32_t) sizeof (SinceT) + 4u;       // padded row lengths in bytes (ThreadLayout)
#define my_state(site_) ((int) lds_u8 (o_state + (uint32_t) (site_) * Bs))
#define set_state(site_, v_) sts_u8 (o_state + (uint32_t) (site_) * Bs, (uint32_t) (v_))
#define my_since(site_) lds_stamp<SinceT> (o_since + (uint32_t) (site_) * Bt)
#define set_since(site_, v_) sts_stamp<SinceT> (o_since + (uint32_t) (site_) * Bt, (uint32_t) (v_))
    long long *my_chain_counts = PER_CHAIN ? p.chain_counts + ((size_t) q * p.nchains + local) * ninst : nullptr;

    // stage the model: difference rows, cross terms, packed site and pair tables
    for (int e = tid; e < p.nrows * nrel; e += B) s_e[(e / nrel) * L.estride + (e % nrel)] = p.dw[(size_t) (e / nrel) * p.ldd + (e % nrel)];
    for (int e = tid; e < p.ncross; e += B) s_cross[e] = p.cross[e];
    for (int k = tid; k < (L.kslots + 1) * (nrel > 0 ? nrel : 1); k += B) s_counts[k] = 0u;
    for (int k = tid; k < L.kslots; k += B) s_nact[k] = 0u;
    for (int i = tid; i < nsites; i += B) {
        const int n = p.nsi[i], base = p.dbase[i];
        for (int o = 0; o < n; o++)
            for (int k = 0; k < n - 1; k++) {
                const int l = k + (k >= o ? 1 : 0), mx = l > o ? l : o, mn = l > o ? o : l;
                const uint32_t row = (uint32_t) (base + ((mx * (mx - 1)) >> 1) + mn);
                s_rowtab[2 * base + o * (n - 1) + k] = (SHORT_FIELDS ? (uint32_t) __cvta_generic_to_shared (s_e) + row * (uint32_t) (L.estride * 8) : row) | (l < o ? 0x80000000u : 0u);
            }
    }
    for (int i = tid; i < nsites + p.npairs; i += B) {
        const uint32_t pr = i < nsites ? (uint32_t) i | ((uint32_t) i << 16) : p.pair_packed[i - nsites];
        const int a = (int) (pr & 0xffffu), b = (int) (pr >> 16), na = p.nsi[a], nb = p.nsi[b];
        uint4 entry;
        entry.x = pr;
        entry.y = (uint32_t) (na - 1) | ((uint32_t) (p.first[a] - a) << 19);
        entry.z = (uint32_t) (nb - 1) | ((uint32_t) (p.first[b] - b) << 19);
        entry.w = (i < nsites ? 0u : p.pair_x[i - nsites] & 0x7fffu) | ((uint32_t) (2 * p.dbase[a]) << 15) | ((na < 2 || nb < 2) ? 0x80000000u : 0u);
        s_move[4 * i] = entry; s_move[4 * i + 1] = entry; s_move[4 * i + 2] = entry; s_move[4 * i + 3] = entry;
    }

    // initial state: Philox block (i>>2, 0xFFFFFFFF, chain, pH index), word i&3 -> first_i + mulhi(word, n_i); the state array holds
    // LOCAL instance indices (0 .. n_i - 1), so that the trial loop never subtracts the site's first instance
    {
        pce::Philox4 r = { 0, 0, 0, 0 };
        for (int i = 0; i < nsites; i++) {
            if ((i & 3) == 0) r = pce::philox4x32_10 ((uint32_t) (i >> 2), 0xFFFFFFFFu, chain, phidx, p.k0, p.k1);
            const uint32_t word = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
            set_state (i, __umulhi (word, (uint32_t) p.nsi[i]));
            set_since (i, 0u);
        }
    }
    __syncthreads ();
    if (!PER_CHAIN && active) atomicAdd (&s_nact[my_slot], 1u);

    const double mu   = ph_to_mu (p.temperature, p.ph[q]);
    const double beta = 1.0 / (PCE_R * p.temperature);

    // relative fields, cooperatively: for each chain c of this warp, lanes over the slots r;
    // g[r] = (gtb[k] - gtb[f]) + sum_j wr[a_j][r] with j sequential (spec order).  mu differs per chain: broadcast it.
    for (int c = 0; c < 32; c++) {
        const double mu_c = __shfl_sync (0xffffffffu, mu, c);
        double *hc = s_h + (size_t) (warp_base + c) * L.hstride;
        const uint8_t *sc = s_state + warp_base + c;
        for (int r = lane; r < nrel; r += 32) {
            const int k = p.rk[r], f = p.rf[r];
            const double gk = __dmul_rn (beta, __dsub_rn (p.gintr[k], __dmul_rn ((double) p.protons[k], mu_c)));
            const double gf = __dmul_rn (beta, __dsub_rn (p.gintr[f], __dmul_rn ((double) p.protons[f], mu_c)));
            double acc = __dsub_rn (gk, gf);
            for (int j = 0; j < nsites; j++) acc = __dadd_rn (acc, p.wr[(size_t) (p.first[j] + sc[(size_t) j * Bs]) * p.ldr + r]);
            hc[r] = acc;
        }
    }
    __syncwarp ();
    uint32_t o_g = s_base + p.tl_h + (uint32_t) (tid * L.hstride) * 8u - 8u;      // slot -1: local index 1 of a site with (first - site) = 0 reads slot 0
    asm volatile ("" : "+r"(o_g));            // opaque: kept in a register instead of being rebuilt from the parameters at every use
#define fld(base_, l_) lds_f64 ((base_) + (uint32_t) (l_) * 8u)      // field of local index l >= 1; base = o_g + g of the site

    // initial energy in the reference's order (EnergyModel.c:204-224), then to units of RT
    {
        double gi = 0.0, ws = 0.0;
        int    np = 0;
        for (int i = 0; i < nsites; i++) {
            const int ai = p.first[i] + my_state (i);
            gi = __dadd_rn (gi, p.gintr[ai]);
            np += p.protons[ai];
            const double *row = p.w + (size_t) ai * p.ldw;
            for (int j = 0; j < i; j++) ws = __dadd_rn (ws, row[p.first[j] + my_state (j)]);
        }
        if (ENERGY) {
            my_eng[0] = __dmul_rn (beta, __dadd_rn (__dsub_rn (gi, __dmul_rn ((double) np, mu)), ws));
            my_eng[B] = 0.0;
        }
    }

    const uint32_t nmoves = p.nmoves;
    const uint32_t nscans = (uint32_t) p.nequil + (uint32_t) p.nprod;
    uint32_t c_macc = 0, c_flips = 0, c_facc = 0;
    uint32_t lg_m = 0, lg_ma = 0, lg_f = 0, lg_fa = 0;
    uint32_t scan = 0, mv = 0;            // production <=> scan >= nequil; production scan index tp = scan - nequil

    // one trial: proposal from (w0), Metropolis with (w1), cooperative application of accepted moves
    const bool has_pairs = p.npairs > 0;
    const uint4 *my_move = s_move + (lane & 3);
    const uint32_t h_addr = (uint32_t) __cvta_generic_to_shared (s_h), e_addr = (uint32_t) __cvta_generic_to_shared (s_e);
    auto trial = [&] (const uint32_t w0, const uint32_t w1) {
        const unsigned long long prod = (unsigned long long) w0 * nmoves;
        const uint32_t idx = (uint32_t) (prod >> 32), lo = (uint32_t) prod;
        const bool     dbl = idx >= (uint32_t) nsites;
        const uint4    move = my_move[4u * idx];
        const int sa = (int) (move.x & 0xffffu), sb = (int) (move.x >> 16);
        // local indices throughout: l* = proposed, o* = current instance of the site; the first instance of a site is the zero
        // of its relative fields
        const uint32_t nam1 = move.y & 0xffffu;
        uint32_t a_fld = o_g + (move.y >> 16);
        const int oa = my_state (sa);
        const bool valid = (int) move.w >= 0;
        const unsigned long long proda = (unsigned long long) lo * nam1;
        const int ka = (int) (proda >> 32);
        int la = ka + (ka >= oa ? 1 : 0);
        int lb = 0, ob = 0;
        uint32_t b_fld = 0u;
        double x;
#define PCE_PIN(v_) asm volatile ("" : "+r"(v_))       /* evaluated once, here: ptxas otherwise re-derives it under every predicate that uses it */
        if (has_pairs) {
            // evaluated for every lane (a warp almost always holds a double move; no divergent branch); for single
            // moves lb == ob and the cross term is zero, so the value equals the single-move formula bit for bit
            const uint32_t nbm1 = move.z & 0xffffu;
            b_fld = o_g + (move.z >> 16);
            ob = my_state (sb);
            const int kb = (int) __umulhi ((uint32_t) proda, nbm1);
            lb = kb + (kb >= ob ? 1 : 0);
            if (!dbl || !valid) lb = ob;
            if (!valid) la = oa;
            PCE_PIN (la); PCE_PIN (lb); PCE_PIN (a_fld); PCE_PIN (b_fld);
            double cross = 0.0;
            if (dbl && valid) cross = s_cross[(move.w & 0x7fffu) + ((uint32_t) oa * nam1 + (uint32_t) ka) * (nbm1 * (nbm1 + 1u)) + (uint32_t) ob * nbm1 + (uint32_t) kb];
            const double gan = la ? fld (a_fld, la) : 0.0, gao = oa ? fld (a_fld, oa) : 0.0;
            const double gbn = lb ? fld (b_fld, lb) : 0.0, gbo = ob ? fld (b_fld, ob) : 0.0;
            x = __dadd_rn (__dadd_rn (__dsub_rn (gan, gao), __dsub_rn (gbn, gbo)), cross);
            c_flips += dbl ? 1u : 0u;                    // (the counters restart when the production scans begin)
        } else {
            if (!valid) la = oa;
            PCE_PIN (la); PCE_PIN (a_fld);
            const double gan = la ? fld (a_fld, la) : 0.0, gao = oa ? fld (a_fld, oa) : 0.0;
            x = __dsub_rn (gan, gao);
        }
        const bool accept = metropolis (x, w1) && valid;       // (every lane takes part in the vote inside)
        if (PER_CHAIN && p.log_rows != nullptr) { if (dbl) { lg_f++; lg_fa += accept; } else { lg_m++; lg_ma += accept; } }

        // apply accepted moves: two accepting chains per round, one per half-warp; each lane updates slots
        // r = sub, sub+16, sub+32, ... of its chain's field vector: g[r] += (+-)E_a[r] (+ (+-)E_b[r])
        unsigned pending = __ballot_sync (0xffffffffu, accept);
        if (pending) {
            if (SHORT_FIELDS) {
                // nrel <= 32 (compile-time variant, chosen at launch) -- short field vectors (the fixtures: lysozyme 23 slots,
                // defensin 15): every lane owns ONE slot, so a round is straight-line code -- one chain per round on all 32
                // lanes, or two chains (one per half-warp) when the vector fits 16 lanes.  The issue slots, not the lanes,
                // bound this kernel: 1.9 accepting chains per warp-step x ~20 instructions instead of 1.2 rounds x ~60 of
                // the strided loop below.  The accepting lane hands out the shared-memory ADDRESS of its difference row with
                // the direction in the sign bit (and bit 30 = a second site follows: double moves are accepted 3e-5 of
                // the time, so their second row takes a separate, warp-uniform pass).
                uint32_t pa = 0u, pb = 0u;
                if (accept) {
                    pa = s_rowtab[((move.w >> 15) & 0xffffu) + (uint32_t) oa * nam1 + (uint32_t) ka];
                    if (dbl) {
                        pb = s_rowtab[((s_move[4 * sb].w >> 15) & 0xffffu) + (uint32_t) ob * (move.z & 0xffffu) + (uint32_t) (lb - (lb > ob ? 1 : 0))];
                        pa |= 0x40000000u;
                    }
                }
                const bool     second   = __any_sync (0xffffffffu, accept && dbl);
                constexpr bool pairwise = SHORT_FIELDS == 2;
                const int      sub   = pairwise ? (lane & 15) : lane;
                const bool     mine  = sub < nrel;
                const uint32_t o     = (uint32_t) sub * 8u;
                const uint32_t hs8   = (uint32_t) L.hstride * 8u;
                uint32_t hbase = h_addr + (uint32_t) warp_base * hs8 + o;
                asm volatile ("" : "+r"(hbase));
                do {
                    // highest accepting lane first (one FLO instead of BREV + FLO; the chains are independent, any order does)
                    int src = 31 - __clz ((int) pending);
                    pending ^= 1u << src;
                    if (pairwise) {
                        const int s1 = 31 - __clz ((int) pending);      // -1 when no second chain is waiting
                        if (pending) pending ^= 1u << s1;
                        if (lane >> 4) src = s1;
                    }
                    const uint32_t y = __shfl_sync (0xffffffffu, pa, src);       // (the source lane is taken modulo 32: -1 reads lane 31, unused)
                    const bool     work = pairwise ? (mine && src >= 0) : mine;
                    const uint32_t hslot = hbase + (uint32_t) src * hs8;
                    if (work) sts_f64 (hslot, __dadd_rn (lds_f64 (hslot), flip_sign (lds_f64 ((y & 0x3fffffffu) + o), y & 0x80000000u)));
                    if (second) {        // warp-uniform and rare
                        const uint32_t yb = __shfl_sync (0xffffffffu, pb, src);
                        if (work && (y & 0x40000000u)) sts_f64 (hslot, __dadd_rn (lds_f64 (hslot), flip_sign (lds_f64 ((yb & 0x3fffffffu) + o), yb & 0x80000000u)));
                    }
                } while (pending);
            } else {
            // difference row and direction of each changed site: row | reverse << 14, site b in the upper half, double << 30
            uint32_t packed = 0u;
            if (accept) {
                const uint32_t ta = s_rowtab[((move.w >> 15) & 0xffffu) + (uint32_t) oa * nam1 + (uint32_t) ka];
                packed = (ta & 0x3fffu) | (ta >> 31 << 14);
                if (dbl) {
                    const uint32_t tb = s_rowtab[((s_move[4 * sb].w >> 15) & 0xffffu) + (uint32_t) ob * (move.z & 0xffffu) + (uint32_t) (lb - (lb > ob ? 1 : 0))];
                    packed |= ((tb & 0x3fffu) << 15) | (tb >> 31 << 29) | (1u << 30);
                }
            }
            // longer field vectors: two chains per round (16 lanes each) up to 48 slots, else one chain per round (32 lanes)
            const int  lpc  = nrel > 48 ? 32 : 16;
            const int  half = lpc == 16 ? lane >> 4 : 0, sub = lane & (lpc - 1);
            do {
                const int s0 = __ffs (pending) - 1;
                pending &= pending - 1;
                int s1 = -1;
                if (lpc == 16 && pending) { s1 = __ffs (pending) - 1; pending &= pending - 1; }
                const int  src  = half ? s1 : s0;
                const bool work = src >= 0;
                const uint32_t y = __shfl_sync (0xffffffffu, packed, src & 31);
                if (work) {
                    const bool two_sites = (y >> 30) & 1u;
                    const uint32_t fa_ = (y >> 14) & 1u ? 0x80000000u : 0u, fb_ = (y >> 29) & 1u ? 0x80000000u : 0u;
                    const uint32_t hrow = h_addr + (uint32_t) ((warp_base + src) * L.hstride) * 8u;
                    const uint32_t ra = e_addr + (uint32_t) ((y & 0x3fffu) * L.estride) * 8u, rb = e_addr + (uint32_t) (((y >> 15) & 0x3fffu) * L.estride) * 8u;
#pragma unroll 1
                    for (int r = sub; r < nrel; r += 2 * lpc) {
                        const uint32_t o0 = (uint32_t) r * 8u, o1 = o0 + (uint32_t) lpc * 8u;
                        const bool p1 = r + lpc < nrel;
                        double h0 = lds_f64 (hrow + o0), a0 = lds_f64 (ra + o0), h1 = 0.0, a1 = 0.0;
                        if (p1) { h1 = lds_f64 (hrow + o1); a1 = lds_f64 (ra + o1); }
                        h0 = __dadd_rn (h0, flip_sign (a0, fa_)); h1 = __dadd_rn (h1, flip_sign (a1, fa_));
                        if (two_sites) {
                            h0 = __dadd_rn (h0, flip_sign (lds_f64 (rb + o0), fb_));
                            if (p1) h1 = __dadd_rn (h1, flip_sign (lds_f64 (rb + o1), fb_));
                        }
                        sts_f64 (hrow + o0, h0);
                        if (p1) sts_f64 (hrow + o1, h1);
                    }
                }
            } while (pending);
            }
            __syncwarp ();
            if (accept) {
                const bool production = scan >= (uint32_t) p.nequil;
                const uint32_t tp = production ? scan - (uint32_t) p.nequil : 0u;
                if (ENERGY) my_eng[0] = __dadd_rn (my_eng[0], x);
                const uint32_t since_a = my_since (sa);
                if (PER_CHAIN) { if (active) my_chain_counts[(int) (move.y >> 19) + sa + oa] += (long long) (tp - since_a); }       // (idle threads run the last chain again)
                else if (tp != since_a && oa) atomicAdd (&my_counts[(int) (move.y >> 19) + oa - 1], tp - since_a);
                set_since (sa, tp);
                set_state (sa, la);
                if (dbl) {
                    const uint32_t since_b = my_since (sb);
                    if (PER_CHAIN) { if (active) my_chain_counts[(int) (move.z >> 19) + sb + ob] += (long long) (tp - since_b); }
                    else if (tp != since_b && ob) atomicAdd (&my_counts[(int) (move.z >> 19) + ob - 1], tp - since_b);
                    set_since (sb, tp);
                    set_state (sb, lb);
                    c_facc++;
                } else c_macc++;
            }
        }
        // scan bookkeeping (MCModelDefault.c:298-312): after the last trial of a production scan the energy is recorded
        if (__builtin_expect (++mv == nmoves, 0)) {       // (warp-uniform)
            __syncwarp ();        // keeps this block a branch target: predicated, its dozen instructions would issue after every trial
            mv = 0;
            const bool production = scan >= (uint32_t) p.nequil;
            if (ENERGY && production) {
                const double grt = my_eng[0];
                my_eng[B] += grt;
                if (PER_CHAIN && p.trajectory != nullptr && q == 0 && local == 0 && active) p.trajectory[scan - (uint32_t) p.nequil] = grt / beta;
            }
            if (PER_CHAIN && p.log_rows != nullptr) {        // the per-block counter table of MCModelDefault.pyx:85-151
                const long long done = (long long) (production ? scan - (uint32_t) p.nequil : scan) + 1;
                if (done % p.log_frequency == 0) {
                    if (q == 0 && local == 0 && active) {
                        long long *row = p.log_rows + 5 * ((production ? p.nequil / p.log_frequency : 0) + done / p.log_frequency - 1);
                        row[0] = done; row[1] = lg_m; row[2] = lg_ma; row[3] = lg_f; row[4] = lg_fa;
                    }
                    lg_m = lg_ma = lg_f = lg_fa = 0;
                }
                if (!production && scan + 1 == (uint32_t) p.nequil) lg_m = lg_ma = lg_f = lg_fa = 0;
            }
            scan++;
            if (scan == (uint32_t) p.nequil) c_macc = c_flips = c_facc = 0u;      // the move counters cover the production scans
        }
    };

    for (unsigned long long b = 0; scan < nscans; b++) {
        const pce::Philox4 r = pce::philox4x32_10 ((uint32_t) b, (uint32_t) (b >> 32), chain, phidx, p.k0, p.k1);
#pragma unroll 1
        for (int half = 0; half < 2; half++) {          // one copy of the trial body in the instruction cache
            if (half && scan >= nscans) break;           // odd number of trials: the last block serves one
            trial (half ? r.z : r.x, half ? r.w : r.y);
        }
    }

    // flush residence times of the final state
    const uint32_t c_moves = (uint32_t) ((unsigned long long) p.nprod * nmoves - c_flips);
    if (active) {
        for (int i = 0; i < nsites; i++) {
            const int      a = my_state (i), gi = (int) (s_move[4 * i].y >> 19);       // local index, first - site
            const uint32_t d = (uint32_t) p.nprod - my_since (i);
            if (PER_CHAIN) my_chain_counts[gi + i + a] += (long long) d;
            else if (d && a) atomicAdd (&my_counts[gi + a - 1], d);
        }
        if (PER_CHAIN) {
            const size_t c = (size_t) q * p.nchains + local;
            for (int i = 0; i < nsites; i++) p.chain_state[c * nsites + i] = (int) (s_move[4 * i].y >> 19) + i + my_state (i);
            p.chain_energy[c] = my_eng[0] / beta;
            p.chain_esum[c]   = my_eng[B] / beta;
            p.chain_counters[c * 4 + 0] = c_moves; p.chain_counters[c * 4 + 1] = c_macc;
            p.chain_counters[c * 4 + 2] = c_flips; p.chain_counters[c * 4 + 3] = c_facc;
        }
    }
    if (!PER_CHAIN) {
        __syncthreads ();
        if (p.counts != nullptr) {
            // non-first instances: the block's counters; first instance of a site: all production scans of the slot's chains minus its siblings
            const int nr = nrel > 0 ? nrel : 1;
            for (int k = tid; k < L.kslots * nrel; k += B) {
                int qq = q_base + k / nrel;
                if (p.ph_minor && qq >= p.nph) qq -= p.nph;
                if (qq < p.nph && s_counts[k]) atomicAdd (&p.counts[(size_t) qq * ninst + p.rk[k % nrel]], (unsigned long long) s_counts[k]);
            }
            for (int k = tid; k < L.kslots * nsites; k += B) {
                const int slot = k / nsites, i = k - slot * nsites;
                int qq = q_base + slot;
                if (p.ph_minor && qq >= p.nph) qq -= p.nph;
                if (qq >= p.nph || s_nact[slot] == 0u) continue;
                const int gi = (int) (s_move[4 * i].y >> 19), ni = (int) (s_move[4 * i].y & 0xffffu) + 1;
                unsigned long long first_count = (unsigned long long) s_nact[slot] * (unsigned long long) p.nprod;
                for (int d = 1; d < ni; d++) first_count -= (unsigned long long) s_counts[(size_t) slot * nr + (gi + d - 1)];
                if (first_count) atomicAdd (&p.counts[(size_t) qq * ninst + gi + i], first_count);
            }
        }
        // counters and energy sums: one atomic per thread (pH values may differ within a warp)
        if (active) {
            if (ENERGY && p.esum != nullptr) atomicAdd (&p.esum[q], my_eng[B] / beta);
            if (p.counters != nullptr) {
                atomicAdd ((unsigned long long *) &p.counters[q * 4 + 0], (unsigned long long) c_moves);
                atomicAdd ((unsigned long long *) &p.counters[q * 4 + 1], (unsigned long long) c_macc);
                atomicAdd ((unsigned long long *) &p.counters[q * 4 + 2], (unsigned long long) c_flips);
                atomicAdd ((unsigned long long *) &p.counters[q * 4 + 3], (unsigned long long) c_facc);
            }
        }
    }
}
#undef my_state
#undef set_state
#undef my_since
#undef set_since
#undef fld
#undef PCE_PIN

// ---------------------------------------------------------------------------------------------------
// Tables of the relative-field formulation and the application of an accepted move in the warp kernel.
//   wr[a][r(k)] = wb[a][k] - wb[a][first of k's site]     (k_build_wr; rows of it initialise the fields)
//   E[row][r]   = wr[hi][r] - wr[lo][r]                   (k_build_dw; one row per instance pair lo < hi of a site)
// An accepted move old -> new adds (+-)E of its instance pair to every field slot (the opposite direction is the exact
// negation): one row of L2 traffic per changed site.  Loads are issued in fixed batches of U unpredicated 128-bit loads per
// lane over a field vector padded to whole batches, so that no serial remainder loop is left.
// ---------------------------------------------------------------------------------------------------
__global__ void k_build_wr (const double *__restrict__ wb, int ldw, int nrel, int ldr, const int32_t *__restrict__ rk, const int32_t *__restrict__ rf,
                            double *__restrict__ wr) {
    const double *row = wb + (size_t) blockIdx.x * ldw;
    for (int r = threadIdx.x; r < ldr; r += blockDim.x) wr[(size_t) blockIdx.x * ldr + r] = r < nrel ? __dsub_rn (row[rk[r]], row[rf[r]]) : 0.0;
}

__global__ void k_build_dw (const double *__restrict__ wr, int nrel, int ldr, int ldd, const int32_t *__restrict__ row_lo,
                            const int32_t *__restrict__ row_hi, double *__restrict__ dw, int abs, int ninst,
                            const int32_t *__restrict__ site_of, const int32_t *__restrict__ first) {
    const int e = blockIdx.x;
    const double *hi = wr + (size_t) row_hi[e] * ldr, *lo = wr + (size_t) row_lo[e] * ldr;
    if (!abs) {     // compact slots
        for (int r = threadIdx.x; r < ldd; r += blockDim.x) dw[(size_t) e * ldd + r] = r < nrel ? __dsub_rn (hi[r], lo[r]) : 0.0;     // zero padding
    } else {        // instance-indexed slots: zero at the first instance of every site
        for (int k = threadIdx.x; k < ldd; k += blockDim.x) {
            double v = 0.0;
            if (k < ninst) {
                const int site = site_of[k];
                if (k != first[site]) v = __dsub_rn (hi[k - site - 1], lo[k - site - 1]);
            }
            dw[(size_t) e * ldd + k] = v;
        }
    }
}

// g += (+-)E1 (+ (+-)E2): n2 (the padded field length in double2) is a multiple of 32 * U, so every batch is U
// unpredicated 128-bit loads per lane and row, all issued before the first use
template <int U>
__device__ __forceinline__ void apply_diff_rows (double2 *h2, int n2, int lane, const double2 *r1, uint32_t f1, const double2 *r2, uint32_t f2, bool two) {
    if (U == 1) {        // field vectors of up to 64 slots (lysozyme, defensin; their own build): one 128-bit load per lane, no batches
        if (lane < n2) {
            const double2 a = __ldg (r1 + lane);
            double2 v = h2[lane];
            v.x = __dadd_rn (v.x, flip_sign (a.x, f1));
            v.y = __dadd_rn (v.y, flip_sign (a.y, f1));
            if (two) {
                const double2 b = __ldg (r2 + lane);
                v.x = __dadd_rn (v.x, flip_sign (b.x, f2));
                v.y = __dadd_rn (v.y, flip_sign (b.y, f2));
            }
            h2[lane] = v;
        }
        return;
    }
    if (U < 16) {       // small fields: one copy of each loop, the direction as a sign flip per element
        if (!two) {
            for (int k0 = lane; k0 < n2; k0 += 32 * U) {
                double2 a[U];
#pragma unroll
                for (int u = 0; u < U; u++) a[u] = __ldg (r1 + k0 + 32 * u);
#pragma unroll
                for (int u = 0; u < U; u++) {
                    double2 v = h2[k0 + 32 * u];
                    v.x = __dadd_rn (v.x, flip_sign (a[u].x, f1));
                    v.y = __dadd_rn (v.y, flip_sign (a[u].y, f1));
                    h2[k0 + 32 * u] = v;
                }
            }
        } else {
            for (int k0 = lane; k0 < n2; k0 += 32 * U) {
                double2 a[U], b[U];
#pragma unroll
                for (int u = 0; u < U; u++) { a[u] = __ldg (r1 + k0 + 32 * u); b[u] = __ldg (r2 + k0 + 32 * u); }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    double2 v = h2[k0 + 32 * u];
                    v.x = __dadd_rn (__dadd_rn (v.x, flip_sign (a[u].x, f1)), flip_sign (b[u].x, f2));
                    v.y = __dadd_rn (__dadd_rn (v.y, flip_sign (a[u].y, f1)), flip_sign (b[u].y, f2));
                    h2[k0 + 32 * u] = v;
                }
            }
        }
        return;
    }
    // large fields (64 elements per lane and row): the directions are warp-uniform, so there is one loop per sign pattern with the
    // negation folded into the add (v - a == v + (-a) bit for bit) instead of a sign flip per element
    auto one = [&] (auto neg1) {
        for (int k0 = lane; k0 < n2; k0 += 32 * U) {
            double2 a[U];
#pragma unroll
            for (int u = 0; u < U; u++) a[u] = __ldg (r1 + k0 + 32 * u);
#pragma unroll
            for (int u = 0; u < U; u++) {
                double2 v = h2[k0 + 32 * u];
                v.x = decltype (neg1)::value ? __dsub_rn (v.x, a[u].x) : __dadd_rn (v.x, a[u].x);
                v.y = decltype (neg1)::value ? __dsub_rn (v.y, a[u].y) : __dadd_rn (v.y, a[u].y);
                h2[k0 + 32 * u] = v;
            }
        }
    };
    auto both = [&] (auto neg1, auto neg2) {
        for (int k0 = lane; k0 < n2; k0 += 32 * U) {
            double2 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; u++) { a[u] = __ldg (r1 + k0 + 32 * u); b[u] = __ldg (r2 + k0 + 32 * u); }
#pragma unroll
            for (int u = 0; u < U; u++) {
                double2 v = h2[k0 + 32 * u];
                v.x = decltype (neg1)::value ? __dsub_rn (v.x, a[u].x) : __dadd_rn (v.x, a[u].x);
                v.y = decltype (neg1)::value ? __dsub_rn (v.y, a[u].y) : __dadd_rn (v.y, a[u].y);
                v.x = decltype (neg2)::value ? __dsub_rn (v.x, b[u].x) : __dadd_rn (v.x, b[u].x);
                v.y = decltype (neg2)::value ? __dsub_rn (v.y, b[u].y) : __dadd_rn (v.y, b[u].y);
                h2[k0 + 32 * u] = v;
            }
        }
    };
    using T = std::true_type;
    using F = std::false_type;
    if (!two) { if (f1) one (T ()); else one (F ()); }
    else if (f1) { if (f2) both (T (), T ()); else both (T (), F ()); }
    else { if (f2) both (F (), T ()); else both (F (), F ()); }
}

// the same from the two wr rows of each instance pair (models whose difference rows exceed the memory budget); nr2 = ldr / 2
template <int U>
__device__ __forceinline__ void apply_wr_rows (double2 *h2, int nr2, int lane, const double2 *ahi, const double2 *alo, uint32_t f1,
                                               const double2 *bhi, const double2 *blo, uint32_t f2, bool two) {
    if (!two) {
#pragma unroll (U)
        for (int k = lane; k < nr2; k += 32) {
            const double2 a = __ldg (ahi + k), b = __ldg (alo + k);
            double2 v = h2[k];
            v.x = __dadd_rn (v.x, flip_sign (__dsub_rn (a.x, b.x), f1));
            v.y = __dadd_rn (v.y, flip_sign (__dsub_rn (a.y, b.y), f1));
            h2[k] = v;
        }
    } else {
#pragma unroll (U / 2 > 0 ? U / 2 : 1)
        for (int k = lane; k < nr2; k += 32) {
            const double2 a = __ldg (ahi + k), b = __ldg (alo + k), c = __ldg (bhi + k), d = __ldg (blo + k);
            double2 v = h2[k];
            v.x = __dadd_rn (__dadd_rn (v.x, flip_sign (__dsub_rn (a.x, b.x), f1)), flip_sign (__dsub_rn (c.x, d.x), f2));
            v.y = __dadd_rn (__dadd_rn (v.y, flip_sign (__dsub_rn (a.y, b.y), f1)), flip_sign (__dsub_rn (c.y, d.y), f2));
            h2[k] = v;
        }
    }
}

template <int U, bool ROWS_ONLY>
__device__ __forceinline__ void apply_move (const McParams &p, const uint32_t *s_site, double *h, int ldh, int lane,
                                            int sa, int sb, int ia, int oa, int ib, int ob, bool dbl) {
    double2 *h2 = (double2 *) h;
    // instance pair (lo < hi, local indices) and direction of each changed site
    const int fa = (int) (s_site[sa] & 0xffffu), la = ia - fa, pa = oa - fa;
    const int mxa = la > pa ? la : pa, mna = la > pa ? pa : la;
    const uint32_t f1 = la < pa ? 0x80000000u : 0u;
    int fb = 0, mxb = 1, mnb = 0;
    uint32_t f2 = 0u;
    if (dbl) {
        fb = (int) (s_site[sb] & 0xffffu);
        const int lb = ib - fb, pb = ob - fb;
        mxb = lb > pb ? lb : pb; mnb = lb > pb ? pb : lb;
        f2 = lb < pb ? 0x80000000u : 0u;
    }
    if (ROWS_ONLY || p.dw != nullptr) {
        // (first difference row of a site: second half of the CTA's site table; the table of rows holds < 2^32 doubles)
        const uint32_t *s_dbase = s_site + p.nsites;
        const double2 *r1 = (const double2 *) (p.dw + (s_dbase[sa] + (uint32_t) (((mxa * (mxa - 1)) >> 1) + mna)) * (uint32_t) p.ldd), *r2 = r1;
        if (dbl) r2 = (const double2 *) (p.dw + (s_dbase[sb] + (uint32_t) (((mxb * (mxb - 1)) >> 1) + mnb)) * (uint32_t) p.ldd);
        if (U == 4 && ldh == 192) apply_diff_rows<3> (h2, 96, lane, r1, f1, r2, f2, dbl);       // (warp-uniform)
        else apply_diff_rows<U> (h2, ldh >> 1, lane, r1, f1, r2, f2, dbl);
    } else {
        apply_wr_rows<(U < 8 ? U : 8)> (h2, p.ldr >> 1, lane, (const double2 *) (p.wr + (size_t) (fa + mxa) * p.ldr), (const double2 *) (p.wr + (size_t) (fa + mna) * p.ldr), f1,
                                       (const double2 *) (p.wr + (size_t) (fb + mxb) * p.ldr), (const double2 *) (p.wr + (size_t) (fb + mnb) * p.ldr), f2, dbl);
    }
}

// ---------------------------------------------------------------------------------------------------
// k_mc_warp: one chain per warp, trials evaluated SPECULATIVELY IN PARALLEL BY THE LANES.
//
// A trial's proposal and Metropolis test depend on the chain's state only; as long as no move is accepted the state
// does not change, so the next 32*TPL trials can all be evaluated at once against the current state, one (or TPL) per
// lane.  The first accepted trial (ballot + find-first-set) is applied; the trials before it were rightly rejected,
// the ones after it are evaluated again in the next round.  The sequence of decisions is exactly that of one-at-a-time
// evaluation (the chain specification), at 1/32 .. 1/64 of the instructions when acceptance is low, and the four
// pair cross-term gathers of up to 64 trials are in flight together instead of one trial's.
//
// The trials wait in a per-warp shared-memory ring (16 bytes each: drawn instances | w1 | proposal target | pair entry), refilled
// 64 trials at a time: lane l generates Philox block b + l (two trials) and decodes everything that does not depend on the
// chain's state -- target site or pair (its L2 latency is paid long before the trial is evaluated) and the drawn instances.
// A round never crosses the end of a scan, so the per-scan bookkeeping stays once per scan.
// Accepted moves are applied by all lanes with 128-bit loads of the beta-scaled W rows from L2.
// Shared memory per warp: g[ldh] doubles (relative fields) | since[nsites] u32 | state[nsites] u16 | ring[RING] uint4
// ---------------------------------------------------------------------------------------------------
struct WarpLayout {
    size_t off_site, off_h, off_since, off_state, off_ring, per_warp, shared_bytes;
    int    ldh, ring;
    // abs: the relative fields are stored at the instance's own index with a zero at every first instance (small models:
    // no index arithmetic in the proposal); else compactly, without the first instances (large models: ninst - nsites slots)
    // since_bytes: width of the residence stamps (2 when nprod < 65536: a 1000-site chain then takes 21.9 instead of 23.9 KB and ten
    // instead of nine fit an SM)
    __host__ __device__ WarpLayout (int nsites, int ninst, int tpl, int upd, bool abs, int since_bytes) {
        const int nfield = abs ? ninst : (ninst > nsites ? ninst - nsites : 1);
        // whole batches of the field update (apply_diff_rows: unpredicated loads); small fields of 129 .. 192 slots (ifp20: 171) are
        // updated in one batch of THREE loads per lane instead of four
        const int gran = (abs && upd == 4 && nfield > 128 && nfield <= 192) ? 192 : 64 * upd;
        ldh  = (nfield + gran - 1) / gran * gran;
        ring = tpl <= 2 ? 128 : 256;
        shared_bytes = ((size_t) nsites * 8 + 15) & ~(size_t) 15;     // packed site table and first difference row per site, one copy per CTA
        off_site = 0;
        size_t o = 0;
        off_h = o;     o += (size_t) ldh * 8;
        off_since = o; o += ((size_t) nsites * since_bytes + 3) & ~(size_t) 3;
        off_state = o; o += (size_t) nsites * 2;
        o = (o + 15) & ~(size_t) 15;
        off_ring = o;  o += (size_t) ring * 16;      // one 16-byte entry per trial
        per_warp = (o + 15) & ~(size_t) 15;
    }
};

// Register budgets (16384 registers per SM sub-partition, a quarter of the resident warps each): 1 CTA of 8 warps 255,
// 2 CTAs 128, 3 CTAs 80, 4 CTAs 64.  Measured on ifp20: 3.4e10 / 3.8e10 / 3.7e10 moves/s for 2 / 3 / 4 CTAs -- at 64
// registers ptxas moves the warp-uniform bookkeeping off the uniform datapath (12 % more instructions), so three CTAs
// are the default for small fields (PCE_MC_CTAS overrides).
// TINY: field vectors of up to 64 slots get a build of their own without the batched row loops -- the round loop of the default
// titration curve is latency-bound down to its instruction-cache footprint (59.4 ms against 61.8 with one more loop compiled in)
template <bool PER_CHAIN, int MIN_CTAS, int TPL, bool ENERGY, bool TINY>
__global__ void __launch_bounds__ (MIN_CTAS == 1 ? 320 : 256, MIN_CTAS) k_mc_warp (const McParams p) {
    static_assert (ENERGY || !PER_CHAIN, "per-chain outputs include the energy");
    static_assert (!TINY || MIN_CTAS != 1, "the large-field build has no tiny variant");
    extern __shared__ __align__ (16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = blockDim.x >> 5;
    constexpr int RING = TPL <= 2 ? 128 : 256;
    constexpr int UPD  = MIN_CTAS == 1 ? 16 : TINY ? 1 : 4;      // 128-bit loads per lane, row and batch when a move is applied
    constexpr bool FALLBACK = MIN_CTAS == 1;          // only the 255-register variant carries the four-gather cross-term path
    constexpr bool ABS = MIN_CTAS != 1;               // small fields: instance-indexed slots (see WarpLayout)
    // The shared-memory layout (WarpLayout) comes from the host in the kernel parameters and every array is addressed by a
    // 32-bit offset: nothing of it has to stay live across the round loop or be recomputed inside it.
    uint32_t o_warp = p.wl_shared + (uint32_t) wib * p.wl_per_warp;          // this warp's slice; fields first
    asm volatile ("" : "+r"(o_warp));       // kept in a register (ptxas otherwise rebuilds it from the thread index inside the round loop)
#define s_site(i_) (*(uint32_t *) (smem + 4u * (uint32_t) (i_)))                     /* first | ninst << 16 per site (CTA-wide) */
#define h(i_)      (*(double *)   (smem + o_warp + 8u * (uint32_t) (i_)))
    // residence stamps: 16-bit in the large-field build when nprod < 65536 (warp-uniform choice; one lane touches them per accepted move)
    auto since = [&] (int i) -> uint32_t {
        return (!ABS && p.wl_since16) ? (uint32_t) *(const uint16_t *) (smem + o_warp + p.wl_since + 2u * (uint32_t) i) : *(const uint32_t *) (smem + o_warp + p.wl_since + 4u * (uint32_t) i);
    };
    auto set_since = [&] (int i, uint32_t v) {
        if (!ABS && p.wl_since16) *(uint16_t *) (smem + o_warp + p.wl_since + 2u * (uint32_t) i) = (uint16_t) v;
        else *(uint32_t *) (smem + o_warp + p.wl_since + 4u * (uint32_t) i) = v;
    };
#define state(i_)  (*(uint16_t *) (smem + o_warp + p.wl_state + 2u * (uint32_t) (i_)))
#define ring(i_)   (*(uint4 *)    (smem + o_warp + p.wl_ring + 16u * (uint32_t) (i_)))  /* per trial: drawn instances | w1 | target | cross-table base and bound */

    const int nsites = p.nsites, ninst = p.ninst, ldw = p.ldw;
    for (int i = threadIdx.x; i < nsites; i += blockDim.x) { s_site (i) = (uint32_t) p.first[i] | ((uint32_t) p.nsi[i] << 16); s_site (nsites + i) = (uint32_t) p.dbase[i]; }
    __syncthreads ();
    const long long gidx = (long long) blockIdx.x * warps + wib;       // chain-major linear (pH, chain) index
    if (gidx >= (long long) p.nph * p.nchains) return;              // whole warp leaves together (no later block barrier)
    // chain-major on purpose: the warps of a CTA work at the same pH value and therefore on the same few difference rows
    // (L1 hits; the pH-minor order the chain-per-thread kernel uses for load balance costs ifp20 30 % here), and CTAs of
    // 8 chains are fine-grained enough for the block scheduler to balance cheap and expensive pH values
    const int q     = (int) (gidx / p.nchains);
    const int local = (int) (gidx % p.nchains);
    const uint32_t chain = (uint32_t) (p.chain_offset + local);
    const uint32_t phidx = (uint32_t) (p.ph_index_offset + q);

    for (int i = lane; i < nsites; i += 32) {
        pce::Philox4 r = pce::philox4x32_10 ((uint32_t) (i >> 2), 0xFFFFFFFFu, chain, phidx, p.k0, p.k1);
        const uint32_t word = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
        state (i) = (uint16_t) (p.first[i] + (int) __umulhi (word, (uint32_t) p.nsi[i]));
        set_since (i, 0u);
    }
    __syncwarp ();
    const double mu   = ph_to_mu (p.temperature, p.ph[q]);
    const double beta = 1.0 / (PCE_R * p.temperature);
    auto gtb = [&] (int k) -> double {       // intrinsic term of instance k in units of RT
        return __dmul_rn (beta, __dsub_rn (__ldg (p.gintr + k), __dmul_rn ((double) __ldg (p.protons + k), mu)));
    };
    // relative fields g[r] = (gtb[k] - gtb[f]) + sum_j wr[a_j][r], j sequential (spec order)
    if (ABS) {
        // instance-indexed slots: lanes over the instances, zero at the first instance of every site
        for (int k = lane; k < p.wl_ldh; k += 32) {
            double acc = 0.0;
            if (k < ninst) {
                const int site = __ldg (p.site_of + k), f = __ldg (p.first + site);
                if (k != f) {
                    const double *col = p.wr + (k - site - 1);
                    acc = __dsub_rn (gtb (k), gtb (f));
                    for (int j = 0; j < nsites; j++) acc = __dadd_rn (acc, __ldg (col + (size_t) state (j) * p.ldr));
                }
            }
            h (k) = acc;
        }
    } else {
        // compact slots: a chunk of 256 consecutive slots per pass (each lane owns 4 double2), every wr row of the state
        // streamed once per chunk with 128-bit loads
        constexpr int CH = 4;
        const int n2 = p.wl_ldh >> 1, nr2 = p.ldr >> 1, nrel = p.nrel;
        for (int c0 = 0; c0 < n2; c0 += 32 * CH) {
            double2 acc[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int r = 2 * (c0 + c * 32 + lane);
                acc[c].x = r < nrel ? __dsub_rn (gtb (__ldg (p.rk + r)), gtb (__ldg (p.rf + r))) : 0.0;
                acc[c].y = r + 1 < nrel ? __dsub_rn (gtb (__ldg (p.rk + r + 1)), gtb (__ldg (p.rf + r + 1))) : 0.0;
            }
#pragma unroll 2
            for (int j = 0; j < nsites; j++) {
                const double2 *row = (const double2 *) (p.wr + (size_t) state (j) * p.ldr);
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const int k2 = c0 + c * 32 + lane;
                    if (k2 < nr2) {
                        const double2 v = __ldg (row + k2);
                        acc[c].x = __dadd_rn (acc[c].x, v.x);
                        acc[c].y = __dadd_rn (acc[c].y, v.y);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const int k2 = c0 + c * 32 + lane;
                if (k2 < n2) ((double2 *) &h (0))[k2] = acc[c];
            }
        }
    }
    __syncwarp ();

    // initial energy in units of RT: sum_i gtb[a_i] + sum_{i > j} wb[a_i][a_j], lanes over i and the partial sums combined, so
    // the value may differ from the serial reference-order sum (EnergyModel.c:204-224) in the last bits.  It only seeds
    // the tracked energy (informational output).
    double grt = 0.0;
    if (ENERGY) {
        double gi = 0.0, ws = 0.0;
        for (int i0 = 0; i0 < nsites; i0 += 32) {          // trip counts are warp-uniform (keeps the whole kernel on the uniform datapath)
            const int  i  = i0 + lane;
            const bool in = i < nsites;
            const int  ai = in ? state (i) : 0;
            if (in) gi += gtb (ai);
            const double *row = p.wb + (size_t) ai * ldw;
            const int jend = i0 + 32 < nsites ? i0 + 32 : nsites;
            for (int j = 0; j < jend; j++) {
                const double v = __ldg (row + state (j));
                if (in && j < i) ws += v;
            }
        }
        grt = pce::warp_sum (gi) + pce::warp_sum (ws);
    }

    const uint32_t nmoves = p.nmoves;
    const uint32_t ntrials = (uint32_t) p.ntrials;       // 32-bit trial counters: the launch refuses longer chains
    uint32_t c_macc = 0, c_flips = 0, c_facc = 0;
    double   esum = 0.0;
    long long scan = 0;
    uint32_t  mv = 0;
    bool      production = p.nequil == 0;
    const bool per_scan = PER_CHAIN && (p.log_rows != nullptr || p.trajectory != nullptr);
    // the chain's energy is followed only in the ENERGY build (energy sums, per-chain outputs): the reference's quiet path returns
    // occupancies only (MCModelDefault.pyx:60-78)
    constexpr bool track = ENERGY;
    unsigned long long *counts = PER_CHAIN ? (unsigned long long *) (p.chain_counts + ((size_t) q * p.nchains + local) * ninst)
                                           : p.counts + (size_t) q * ninst;
    uint32_t lg_m = 0, lg_ma = 0, lg_f = 0, lg_fa = 0;

    // The state-independent part of a trial is decoded ONCE, when its random words are generated (a trial is evaluated in every
    // round until it is consumed, 2-3 times on average): the proposal target (site | site << 16 for a single move, b = a, or the
    // packed pair), the drawn instances before they skip the current one (first + draw, 16 bits per site; 0xffffffff: a site with a
    // single instance, nothing to propose) and the pair's entry (cross-table base | quantised bound).
    auto decode = [&] (const uint32_t w0, uint32_t &pr, uint32_t &draws, uint32_t &px) {
        const unsigned long long prod = (unsigned long long) w0 * nmoves;
        const uint32_t idx = (uint32_t) (prod >> 32), lo = (uint32_t) prod;
        const bool dbl = idx >= (uint32_t) nsites;
        pr = dbl ? __ldg (p.pair_packed + (idx - nsites)) : (idx | (idx << 16));
        px = dbl ? __ldg (p.pair_x + (idx - nsites)) : 0u;
        const uint32_t a_info = s_site (pr & 0xffffu), b_info = s_site (pr >> 16);
        const uint32_t na = a_info >> 16, nb = b_info >> 16;
        const unsigned long long proda = (unsigned long long) lo * (na - 1u);
        const uint32_t ka = (uint32_t) (proda >> 32), kb = __umulhi ((uint32_t) proda, nb - 1u);
        draws = (na >= 2u && nb >= 2u) ? (((a_info & 0xffffu) + ka) | (((b_info & 0xffffu) + kb) << 16)) : 0xffffffffu;
    };

    const bool use_cross = p.cross != nullptr;
    uint32_t tcur = 0, tfill = 0;        // next trial to decide; trials [tcur, tfill) are in the ring
    auto refill = [&] () {               // 64 more trials: lane l generates Philox block tfill / 2 + l and decodes its two trials
        const uint32_t bl = (tfill >> 1) + (uint32_t) lane;
        const pce::Philox4 r = pce::philox4x32_10 (bl, 0u, chain, phidx, p.k0, p.k1);
        const uint32_t slot = ((uint32_t) tfill & (uint32_t) (RING - 1)) + 2u * (uint32_t) lane;
        uint32_t pr0, pr1, dr0, dr1, px0, px1;
        decode (r.x, pr0, dr0, px0);
        decode (r.z, pr1, dr1, px1);
        ring (slot)      = make_uint4 (dr0, r.y, pr0, px0);
        ring (slot + 1u) = make_uint4 (dr1, r.w, pr1, px1);
        tfill += 64u;
    };
    // Large fields, one trial per lane: the ring is topped up in the shadow of the round's cross-term gathers (below) instead of
    // at the start of a round, where the L2 round trip of the pair table would be exposed; the ring then holds the current
    // round, the next one and the 64 new trials (32 + 32 + 64 = RING).
    constexpr bool EARLY_REFILL = !ABS && TPL == 1;
    if (EARLY_REFILL) { refill (); __syncwarp (); }
    while (tcur < ntrials) {
        if (!EARLY_REFILL && tfill - tcur < (uint32_t) (32 * TPL)) {
            do refill (); while (tfill - tcur < (uint32_t) (32 * TPL));
            __syncwarp ();       // (otherwise the barrier that ends a round has already ordered the ring)
        }

        // ---- evaluate 32*TPL trials against the current state ----
        // A round may run across scan ends (a scan of a small model is shorter than a round: lysozyme 24 trials); only the
        // logging / trajectory mode, which reports per scan, keeps its rounds within one scan.
        const uint32_t left_total = ntrials - tcur;
        uint32_t n_act = left_total < (uint32_t) (32 * TPL) ? left_total : (uint32_t) (32 * TPL);
        if (per_scan && nmoves - mv < n_act) n_act = nmoves - mv;
        uint32_t e_pr[TPL], e_a[TPL], e_b[TPL], e_w1[TPL];     // sa | sb << 16,  ia | oa << 16,  ib | ob << 16
        double   e_x[TPL], e_c[TPL][4];
        bool     e_dbl[TPL], e_need[TPL];
        unsigned accm[TPL], dblm[TPL];
        // phase 1: decode every trial of the round, form the field part of the energy change, and issue the pair
        // cross-term gathers (all in flight together) -- except for double moves that are CERTAINLY rejected whatever
        // their cross term is: x >= x0 - B with B the pair's bound, so u > exp (-(x0 - B)) (fp32, directed rounding,
        // 1e-3 margin) decides without touching W.  Most double moves end here: 4 L2 sectors saved each.
#pragma unroll
        for (int u = 0; u < TPL; u++) {
            const uint32_t off  = (uint32_t) (u * 32 + lane);
            const uint32_t slot = ((uint32_t) tcur + off) & (uint32_t) (RING - 1);
            const uint4 entry = ring (slot);
            const uint32_t draws = entry.x, pr = entry.z, px = entry.w;
            e_w1[u] = entry.y;
            const float pbound = 0.25f * (float) (px >> 24);                     // the pair's bound, quantised upwards (0: none)
            const int  sa = (int) (pr & 0xffffu), sb = (int) (pr >> 16);
            const bool dbl = sa != sb;                                           // (a pair joins two different sites)
            const bool valid = draws != 0xffffffffu;
            const int  oa = state (sa);
            int ia = (int) (draws & 0xffffu);
            if (ia >= oa) ia++;
            if (!valid) ia = oa;
            int ib = 0, ob = 0;
            e_c[u][0] = e_c[u][1] = e_c[u][2] = e_c[u][3] = 0.0;
            e_need[u] = false;
            double x = 0.0;
            if (dbl) {
                ob = state (sb);
                ib = (int) (draws >> 16);
                if (ib >= ob) ib++;
                if (!valid) ib = ob;
                // relative fields: 0 for a first instance (stored as such in the instance-indexed layout)
                uint32_t ia_info = 0u, ib_info = 0u;
                if (!ABS) { ia_info = s_site (sa); ib_info = s_site (sb); }
                const int fa = (int) (ia_info & 0xffffu), fb = (int) (ib_info & 0xffffu);
                const double gan = ABS ? h (ia) : ia != fa ? h (ia - sa - 1) : 0.0, gao = ABS ? h (oa) : oa != fa ? h (oa - sa - 1) : 0.0;
                const double gbn = ABS ? h (ib) : ib != fb ? h (ib - sb - 1) : 0.0, gbo = ABS ? h (ob) : ob != fb ? h (ob - sb - 1) : 0.0;
                x = __dadd_rn (__dsub_rn (gan, gao), __dsub_rn (gbn, gbo));
                const float xlow = __fsub_rd (__double2float_rd (x), pbound);            // <= x0 - B <= x
                const float ulow = __uint2float_rd (e_w1[u]) * 2.3283064365386963e-10f;  // <= u
                // (small fields: the cross-term table is cache-resident, the gather is cheaper than the test)
                const bool certain = !ABS && (px >> 24) != 0u && xlow > 0.0f && ulow > __expf (-xlow) * 1.001f;
                if (valid && !certain) {
                    if (!FALLBACK || use_cross) {
                        // one gather: the cross term of (old a, draw a, old b, draw b), precomputed with the same operations
                        if (ABS) { ia_info = s_site (sa); ib_info = s_site (sb); }
                        const int f_a = (int) (ia_info & 0xffffu), f_b = (int) (ib_info & 0xffffu), na = (int) (ia_info >> 16), nb = (int) (ib_info >> 16);
                        const int la = oa - f_a, lb = ob - f_b, ka = (int) (draws & 0xffffu) - f_a, kb = (int) (draws >> 16) - f_b;
                        const uint32_t e = (px & 0xffffffu) + (uint32_t) ((la * (na - 1) + ka) * (nb * (nb - 1)) + lb * (nb - 1) + kb);
                        e_c[u][0] = __ldg (p.cross + e);
                    } else {
                        const double *ra = p.wb + (size_t) ia * ldw, *ro = p.wb + (size_t) oa * ldw;
                        e_c[u][0] = __ldg (ra + ib); e_c[u][1] = __ldg (ra + ob); e_c[u][2] = __ldg (ro + ib); e_c[u][3] = __ldg (ro + ob);
                    }
                    e_need[u] = true;
                }
            } else {
                const int fa = ABS ? 0 : (int) (s_site (sa) & 0xffffu);
                const double gan = ABS ? h (ia) : ia != fa ? h (ia - sa - 1) : 0.0, gao = ABS ? h (oa) : oa != fa ? h (oa - sa - 1) : 0.0;
                x = __dsub_rn (gan, gao);
                e_need[u] = valid;
            }
            e_dbl[u] = dbl; e_x[u] = x;
            dblm[u] = __ballot_sync (0xffffffffu, dbl);
            e_pr[u] = pr; e_a[u] = (uint32_t) ia | ((uint32_t) oa << 16); e_b[u] = (uint32_t) ib | ((uint32_t) ob << 16);
        }
        if (EARLY_REFILL && tfill - tcur < 64u) refill ();       // (writes slots [tfill, tfill + 64): none of this round's)
        // phase 2: cross terms and Metropolis tests
#pragma unroll
        for (int u = 0; u < TPL; u++) {
            double x = e_x[u];
            if (e_dbl[u]) x = __dadd_rn (x, (!FALLBACK || use_cross) ? e_c[u][0] : __dadd_rn (__dsub_rn (__dsub_rn (e_c[u][0], e_c[u][1]), e_c[u][2]), e_c[u][3]));
            const bool accept = metropolis (x, e_w1[u]) && e_need[u];
            accm[u] = __ballot_sync (0xffffffffu, accept);
            e_x[u] = x;
        }
        // every lane has evaluated a trial; in the last round of a chain (and in the per-scan mode, whose rounds stop at scan ends)
        // only the first n_act of them exist -- warp-uniform and rare, so the lanes carry no "active" flag through the phases above
        if (n_act < (uint32_t) (32 * TPL)) {
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const int c = (int) n_act - u * 32;
                accm[u] &= c >= 32 ? 0xffffffffu : c <= 0 ? 0u : ((1u << c) - 1u);
            }
        }

        // ---- the first accepted trial ends the round ----
        int ustar = -1, lstar = 0;
#pragma unroll
        for (int u = TPL - 1; u >= 0; u--) if (accm[u]) { ustar = u; lstar = __ffs (accm[u]) - 1; }
        const uint32_t consumed = ustar >= 0 ? (uint32_t) (ustar * 32 + lstar + 1) : n_act;
        // scan ends inside the consumed trials: k_before of them lie before the last consumed trial (they see the energy as
        // it was), k_total counts the one right after it as well (it sees the energy after an accepted move)
        uint32_t k_before = 0, k_total = 0;
        if (mv + consumed >= nmoves) {
            const uint32_t past = mv + consumed;
            // small models: multiply-high by the reciprocal, exact as long as past * nmoves < 2^32 (not in the large-field build, whose
            // warp-uniform bookkeeping stays on the uniform datapath)
            if (ABS && p.nmoves_magic) { k_before = __umulhi (past - 1u, p.nmoves_magic); k_total = __umulhi (past, p.nmoves_magic); }
            else if (nmoves > (uint32_t) (32 * TPL)) { k_before = past - 1u >= nmoves ? 1u : 0u; k_total = 1u; }      // at most one scan end
            else if (ABS) { k_before = past - 1u; k_total = past; }        // the reciprocal covers 2 <= nmoves <= 64: this is a model of one site
            else { k_before = (past - 1u) / nmoves; k_total = past / nmoves; }
        }
        uint32_t ndbl = 0, ndbl_prod = 0;
        if (production) {
            if (TPL == 1) ndbl = (uint32_t) __popc (dblm[0] & (0xffffffffu >> (32u - consumed)));        // 1 <= consumed <= 32
            else {
#pragma unroll
                for (int u = 0; u < TPL; u++) {
                    const int c = (int) consumed - u * 32;
                    const unsigned m = c >= 32 ? 0xffffffffu : c <= 0 ? 0u : ((1u << c) - 1u);
                    ndbl += (uint32_t) __popc (dblm[u] & m);
                }
            }
            ndbl_prod = ndbl;
        } else {
            // the round may run into the production phase: j_prod = first consumed trial that belongs to it
            const long long wait = (long long) (p.nequil - scan) * (long long) nmoves - (long long) mv;
            const uint32_t j_prod = wait > (long long) (32 * TPL) ? (uint32_t) (32 * TPL) : (uint32_t) wait;
#pragma unroll
            for (int u = 0; u < TPL; u++) {
                const int c = (int) consumed - u * 32, c0 = (int) j_prod - u * 32;
                const unsigned m = c >= 32 ? 0xffffffffu : c <= 0 ? 0u : ((1u << c) - 1u);
                const unsigned m0 = c0 >= 32 ? 0xffffffffu : c0 <= 0 ? 0u : ((1u << c0) - 1u);
                ndbl += (uint32_t) __popc (dblm[u] & m);
                ndbl_prod += (uint32_t) __popc (dblm[u] & m & ~m0);
            }
        }
        c_flips += ndbl_prod;
        if (PER_CHAIN && p.log_rows != nullptr) { lg_f += ndbl; lg_m += consumed - ndbl; }
        // energy of the production scans that ended before the last consumed trial
        for (uint32_t k = 0; track && k < k_before; k++) {
            if (scan + (long long) k >= (long long) p.nequil) {
                esum += grt;
                if (PER_CHAIN && p.trajectory != nullptr && q == 0 && local == 0 && lane == 0) p.trajectory[scan + k - p.nequil] = grt / beta;
            }
        }

        if (ustar >= 0) {
            uint32_t s_pr = 0, s_a = 0, s_b = 0;
            double   s_x = 0.0;
#pragma unroll
            for (int u = 0; u < TPL; u++) if (u == ustar) { s_pr = e_pr[u]; s_a = e_a[u]; s_b = e_b[u]; s_x = e_x[u]; }
            bool dbl = false;
#pragma unroll
            for (int u = 0; u < TPL; u++) if (u == ustar) dbl = (dblm[u] >> lstar) & 1u;
            s_pr = __shfl_sync (0xffffffffu, s_pr, lstar);
            s_a  = __shfl_sync (0xffffffffu, s_a, lstar);
            if (dbl) s_b = __shfl_sync (0xffffffffu, s_b, lstar);       // (warp-uniform: the winner's flag comes from a ballot)
            if (track) grt = __dadd_rn (grt, __shfl_sync (0xffffffffu, s_x, lstar));
            const int sa = (int) (s_pr & 0xffffu), sb = (int) (s_pr >> 16);
            const int ia = (int) (s_a & 0xffffu), oa = (int) (s_a >> 16), ib = (int) (s_b & 0xffffu), ob = (int) (s_b >> 16);
            if (PER_CHAIN && p.log_rows != nullptr) { if (dbl) lg_fa++; else lg_ma++; }
            apply_move<UPD, ABS> (p, (const uint32_t *) smem, &h (0), p.wl_ldh, lane, sa, sb, ia, oa, ib, ob, dbl);
            // the accepted trial lies in scan `scan + k_before`: production there decides the stamp and the counters
            const long long scan_acc = scan + (long long) k_before;
            const bool      prod_acc = scan_acc >= (long long) p.nequil;
            const uint32_t  tp_acc = prod_acc ? (uint32_t) (scan_acc - p.nequil) : 0u;
            if (lane == 0) {
                const uint32_t since_a = since (sa);
                if (tp_acc != since_a) atomicAdd (&counts[oa], (unsigned long long) (tp_acc - since_a));
                set_since (sa, tp_acc);
                state (sa) = (uint16_t) ia;
                if (dbl) {
                    const uint32_t since_b = since (sb);
                    if (tp_acc != since_b) atomicAdd (&counts[ob], (unsigned long long) (tp_acc - since_b));
                    set_since (sb, tp_acc);
                    state (sb) = (uint16_t) ib;
                }
            }
            if (prod_acc) { if (dbl) c_facc++; else c_macc++; }
        }
        __syncwarp ();       // state, stamps and fields of this round are visible to every lane before the next one

        // the scan (if any) that ends right after the last consumed trial
        if (track && k_total > k_before && scan + (long long) k_before >= (long long) p.nequil) {
            esum += grt;
            if (PER_CHAIN && p.trajectory != nullptr && q == 0 && local == 0 && lane == 0) p.trajectory[scan + k_before - p.nequil] = grt / beta;
        }
        mv = mv + consumed - k_total * nmoves;
        tcur += consumed;
        if (k_total) {     // scan bookkeeping (MCModelDefault.c:298-312)
            if (PER_CHAIN && p.log_rows != nullptr) {       // per_scan mode: k_total == 1
                const long long done = (production ? scan - p.nequil : scan) + 1;
                if (done % p.log_frequency == 0) {
                    if (q == 0 && local == 0 && lane == 0) {
                        long long *row = p.log_rows + 5 * ((production ? p.nequil / p.log_frequency : 0) + done / p.log_frequency - 1);
                        row[0] = done; row[1] = lg_m; row[2] = lg_ma; row[3] = lg_f; row[4] = lg_fa;
                    }
                    lg_m = lg_ma = lg_f = lg_fa = 0;
                }
                if (!production && scan + 1 == p.nequil) lg_m = lg_ma = lg_f = lg_fa = 0;
            }
            scan += (long long) k_total;
            production = scan >= p.nequil;
        }
    }

    __syncwarp ();
    for (int i = lane; i < nsites; i += 32) {
        const uint32_t d = (uint32_t) p.nprod - since (i);
        if (d) atomicAdd (&counts[state (i)], (unsigned long long) d);
    }
    const uint32_t c_moves = (uint32_t) ((unsigned long long) p.nprod * nmoves - c_flips);
    if (PER_CHAIN) {
        const size_t c = (size_t) q * p.nchains + local;
        for (int i = lane; i < nsites; i += 32) p.chain_state[c * nsites + i] = state (i);
        if (lane == 0) {
            p.chain_energy[c] = grt / beta;
            p.chain_esum[c]   = esum / beta;
            p.chain_counters[c * 4 + 0] = c_moves; p.chain_counters[c * 4 + 1] = c_macc;
            p.chain_counters[c * 4 + 2] = c_flips; p.chain_counters[c * 4 + 3] = c_facc;
        }
    } else if (lane == 0) {
        if (ENERGY && p.esum != nullptr) atomicAdd (&p.esum[q], esum / beta);
        if (p.counters != nullptr) {
            atomicAdd ((unsigned long long *) &p.counters[q * 4 + 0], (unsigned long long) c_moves);
            atomicAdd ((unsigned long long *) &p.counters[q * 4 + 1], (unsigned long long) c_macc);
            atomicAdd ((unsigned long long *) &p.counters[q * 4 + 2], (unsigned long long) c_flips);
            atomicAdd ((unsigned long long *) &p.counters[q * 4 + 3], (unsigned long long) c_facc);
        }
    }
}
#undef s_site
#undef h
#undef state
#undef ring

// ---------------------------------------------------------------------------------------------------
// launch planning
// ---------------------------------------------------------------------------------------------------
// upd: 128-bit loads per lane and batch of the field update; ctas_per_sm: the register budget (resident CTAs of 8 warps);
// resident: CTAs of `block` threads that actually share an SM (more, smaller CTAs when a small ensemble is spread evenly)
struct Plan { int kernel, block; size_t smem; int ctas_per_sm; int tpl; int upd; int ph_minor; int resident; };

int smem_limit (int device, size_t *limit) {
    int value = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&value, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *limit = (size_t) value;
    return PCE_OK;
}

// sizes of the relative-field tables of a model: slots per chain, difference rows, cross-table entries
struct TableSizes { int nrel, nrows; size_t ncross; };
TableSizes table_sizes (const pce_mc *mc) {
    const pce_model *m = mc->model;
    TableSizes t = { m->ninst - m->nsites, 0, 0 };
    for (int i = 0; i < m->nsites; i++) { const int n = m->last[i] - m->first[i] + 1; t.nrows += n * (n - 1) / 2; }
    for (size_t k = 0; k < mc->pair_a.size (); k++) {
        const int na = m->last[mc->pair_a[k]] - m->first[mc->pair_a[k]] + 1, nb = m->last[mc->pair_b[k]] - m->first[mc->pair_b[k]] + 1;
        t.ncross += (size_t) na * (na - 1) * nb * (nb - 1);
    }
    return t;
}

// Cost model of one launch, in units of the time one warp of the chain-per-thread kernel needs when it has an SM to itself.
// Measured on B200 (profiles/README.md, round 2, `scripts/r02/planner_sweep.py`: lysozyme and defensin, 41 pH values x
// 1 .. 2772 chains): a warp-step of the chain-per-thread kernel is latency-bound at 1470 cycles alone and slows by 1.8 % per
// co-resident warp (24 warps: x 1.43); the chain-per-warp kernel (speculative lanes) finishes a chain in 0.19 of that time and
// a full grid of resident warps in 0.138, so it wins below ~8 waves of its own grid (~28 000 chains) and loses above.
struct CostModel {
    static constexpr double thread_per_warp = 0.018;     // slow-down of a chain-per-thread warp-step per co-resident warp
    static constexpr double warp_floor = 0.19, warp_per_wave = 0.138, warp_fixed = 0.03;
};

// chain-per-thread plan: the difference rows and the cross-term table must fit in shared memory next to the chains.  Every
// block size (multiple of 32, up to 768 threads -- 896 for short field vectors) that fits is priced with the cost model for
// the `total` = nph x nchains chains of this launch: a small ensemble is spread over all SMs in small blocks (few warps per SM,
// each running at its latency floor), a large one takes the block that puts the most warps on an SM.
bool plan_thread (const pce_mc *mc, size_t limit, int nph, int sms, bool energy, Plan *plan, double *cost_out) {
    const pce_model *m = mc->model;
    const TableSizes t = table_sizes (mc);
    if (m->ninst > 256 || t.nrows > 16383 || t.ncross >= (1u << 15)) return false;       // field widths of the packed move table (the tables must fit in shared memory anyway)
    const char *forced = std::getenv ("PCE_MC_BLOCK");      // tuning aid: force the block size
    const int forced_block = forced ? std::atoi (forced) : 0;
    // (round 1, lysozyme, moves/s: 640 threads 7.6e10, 768 threads 8.5e10; round 2: 768 threads 1.06e11, 896 threads -- the quiet build
    // of the short-field kernel, see below -- 1.12e11, and 1.45e11 after the trial front end was rewritten)
    // pH-minor linearisation (load balance: acceptance, hence cost, varies 10x over a titration curve) unless the per-block
    // occupancy counters of all pH values would take more than 48 KB of shared memory
    const bool ph_minor = (size_t) std::min (nph, 768) * m->ninst * 4 <= 48 * 1024 && std::getenv ("PCE_MC_CHAIN_MAJOR") == nullptr;
    // Short field vectors (nrel <= 32, SHORT_FIELDS build): the uncapped build takes 72 registers, which fits 896 threads
    // (7 warps per sub-partition x 72 x 32 = 16128 registers) -- taken when the chains also fit in shared memory (lysozyme:
    // 247 B per chain without energy tracking; with it, 263 B stop at 832 threads, whose uneven 7/7/6/6 warps do not pay).
    int most = 768;
    if (t.nrel <= 32) {
        cudaFuncAttributes attributes;
        const bool narrow = mc->nprod < 65536;
        const cudaError_t err = narrow ? (energy ? cudaFuncGetAttributes (&attributes, k_mc_thread<false, uint16_t, 128, 1, true>)
                                                 : cudaFuncGetAttributes (&attributes, k_mc_thread<false, uint16_t, 128, 1, false>))
                                       : (energy ? cudaFuncGetAttributes (&attributes, k_mc_thread<false, uint32_t, 128, 1, true>)
                                                 : cudaFuncGetAttributes (&attributes, k_mc_thread<false, uint32_t, 128, 1, false>));
        if (err == cudaSuccess && ((attributes.numRegs + 7) & ~7) * 32 * 7 <= 16384) most = 896;
        else if (err != cudaSuccess) (void) cudaGetLastError ();
    }
    const double total = (double) nph * (double) mc->nchains;
    double best_cost = 0.0;
    int    best_warps = 0;
    for (int block = 32; block <= (forced_block ? 1024 : most); block += 32) {
        if (forced_block && block != forced_block) continue;
        if (!forced_block && block > 768 && block % 128 != 0) continue;       // beyond 768 only whole warps per sub-partition
        ThreadLayout L (m->nsites, m->ninst, (int) mc->pair_a.size (), t.nrel, t.nrows, (int) t.ncross, block, mc->nchains, nph, mc->nprod < 65536 ? 2 : 4, ph_minor, energy);
        if (L.total > limit) continue;
        if ((double) block * (double) mc->nprod >= 4294967295.0) continue;      // per-block occupancy counters are 32-bit
        // resident CTAs: shared memory (1 KB reserved per CTA), threads, registers (88 allocated up to 640 threads, then 80 / 72)
        const int regs = block <= 640 ? 88 : block <= 768 ? 80 : 72;
        const int ctas = (int) std::max<size_t> (1, std::min<size_t> (std::min<size_t> ((228 * 1024) / (L.total + 1024), (size_t) (2048 / block)),
                                                                      (size_t) (65536 / (regs * block))));
        const int    warps_full = ctas * (block / 32);
        const double capacity = (double) sms * ctas * block;
        const double waves = std::ceil (total / capacity);
        const double rest = total - (waves - 1.0) * capacity;                                  // chains of the last wave
        const double warps_last = std::min ((double) warps_full, std::ceil (std::ceil (rest / block) / sms) * (block / 32));
        const double cost = (waves - 1.0) * (1.0 + CostModel::thread_per_warp * warps_full) + (1.0 + CostModel::thread_per_warp * warps_last);
        // ties go to the larger block (measured: 768 threads 64.4 ms against 2 x 384 threads 69.0 ms on a full lysozyme grid)
        if (best_warps == 0 || cost <= best_cost + 1e-9) {
            best_cost = cost; best_warps = warps_full;
            plan->kernel = 1; plan->block = block; plan->smem = L.total; plan->ctas_per_sm = ctas; plan->ph_minor = ph_minor ? 1 : 0; plan->resident = ctas;
        }
    }
    if (cost_out) *cost_out = best_cost;
    // auto mode: with fewer than 8 resident warps the chain-per-warp kernel does better; a forced choice runs with what fits
    return best_warps >= (mc->kernel == 1 ? 1 : 8);
}

// trials evaluated per lane and round by the chain-per-warp kernel: 1 (default; measured best on every model, the kernel
// being bound by the L1TEX/LSU wavefront rate, not by the latency of a round) or 2 (PCE_MC_TPL=2, tuning aid)
int warp_tpl () {
    const char *forced = std::getenv ("PCE_MC_TPL");
    return forced && std::atoi (forced) == 2 ? 2 : 1;
}

bool plan_warp (const pce_mc *mc, size_t limit, int nph, int sms, Plan *plan, double *cost_out) {
    const pce_model *m = mc->model;
    if (m->ninst > 65535 || m->nsites > 65535) return false;
    const int tpl = warp_tpl ();
    // small fields: 4-deep update batches and two or three resident CTAs of 8 warps (128 / 80 registers); large fields
    // (one CTA fits): 255 registers and 16-deep batches
    for (int pass = 0; pass < 3; pass++) {
        // 128-bit loads per lane and batch of the field update: 16 for large fields, 4 for small ones, 1 when the whole vector
        // (instance-indexed: ninst slots) fits one load per lane
        const int upd = pass == 1 ? 16 : m->ninst <= 64 ? 1 : 4;
        WarpLayout L (m->nsites, m->ninst, tpl, upd, pass != 1, (pass == 1 && mc->nprod < 65536) ? 2 : 4);
        if (L.per_warp + L.shared_bytes > limit) continue;
        // large fields (one resident CTA): as many warps as shared memory holds, up to 10 (204 registers each); else 8 per CTA
        int warps = (int) std::min<size_t> ((limit - L.shared_bytes) / L.per_warp, pass == 1 ? 10 : 8);
        if (const char *cap = std::getenv ("PCE_MC_WARPS")) warps = std::max (1, std::min (warps, std::atoi (cap)));      // tuning aid
        const size_t smem = L.shared_bytes + (size_t) warps * L.per_warp;
        const int ctas = (int) std::min<size_t> ((228 * 1024) / (smem + 1024), (size_t) (2048 / (warps * 32)));
        const char *forced = std::getenv ("PCE_MC_CTAS");       // tuning aid: resident CTAs per SM of the small-field variant (2..4)
        const int  most = forced ? std::max (2, std::min (4, std::atoi (forced))) : 3;
        if (pass == 0 && (ctas < 2 || mc->cross_overflow)) continue;
        if (pass == 2 && mc->cross_overflow) continue;      // only the one-CTA 255-register variant gathers W itself
        plan->kernel = 2; plan->block = warps * 32; plan->smem = smem;
        plan->ctas_per_sm = pass == 0 ? std::min (ctas, most) : 1; plan->tpl = tpl; plan->upd = upd;
        plan->resident = plan->ctas_per_sm;
        // An ensemble of less than one wave: the SMs with the most warps set the time (328 CTAs of 8 warps on 148 SMs
        // leave 32 SMs with 24 warps and 116 with 16), so cut the CTAs down to the size that levels the warps per SM
        const long long total = (long long) nph * mc->nchains, per_sm = (long long) plan->ctas_per_sm * warps;
        if (pass != 1 && total < (long long) sms * per_sm && std::getenv ("PCE_MC_WARPS") == nullptr) {
            int best_wpc = warps;
            long long best_load = ((total + warps - 1) / warps + sms - 1) / sms * warps;
            for (int wpc = warps / 2; wpc >= 1; wpc /= 2) {
                const long long load = ((total + wpc - 1) / wpc + sms - 1) / sms * wpc;
                if (load < best_load) { best_load = load; best_wpc = wpc; }
            }
            if (best_wpc != warps) {
                plan->block = best_wpc * 32;
                plan->smem = L.shared_bytes + (size_t) best_wpc * L.per_warp;
                plan->resident = (int) std::min<long long> (32, per_sm / best_wpc);
            }
        }
        if (cost_out) {
            const double waves = (double) nph * (double) mc->nchains / ((double) sms * plan->ctas_per_sm * warps);
            *cost_out = std::max (CostModel::warp_floor, CostModel::warp_fixed + CostModel::warp_per_wave * waves);
        }
        return true;
    }
    return false;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
static int find_pairs (pce_mc *mc) {
    // MCModelDefault_FindPairs + _FindMaxInteraction, MCModelDefault.c:231-288: sites i > j are a pair
    // when max |W| over their instances reaches the limit; order (i ascending, j < i ascending).
    pce_model *m = mc->model;
    const size_t n = (size_t) m->ninst;
    const double *ws = m->host_wsym ();
    mc->pair_a.clear (); mc->pair_b.clear (); mc->pair_wmax.clear ();
    for (int i = 0; i < m->nsites; i++)
        for (int j = 0; j < i; j++) {
            double wmax = 0.0;
            for (int a = m->first[i]; a <= m->last[i]; a++)
                for (int b = m->first[j]; b <= m->last[j]; b++) {
                    const double v = std::fabs (ws[(size_t) a * n + b]);
                    if (v > wmax) wmax = v;
                }
            if (wmax >= mc->limit) { mc->pair_a.push_back (i); mc->pair_b.push_back (j); mc->pair_wmax.push_back (wmax); }
        }
    mc->pairs_dirty = true;
    return PCE_OK;
}

extern "C" int pce_mc_create (pce_model *model, double limit, int nequil, int nprod, long long seed, int nchains, pce_mc **out) {
    if (out == nullptr) return fail (PCE_ERR_VALUE, "null output handle");
    *out = nullptr;
    if (model == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (!model->symmetrized) return fail (PCE_ERR_STATE, "First calculate electrostatic energies.");
    int status = model->sites_ok ();
    if (status != PCE_OK) return status;
    if (nequil < 0 || nprod < 1 || nchains < 1) return fail (PCE_ERR_VALUE, "invalid scan or chain counts");
    pce_mc *mc = new (std::nothrow) pce_mc ();
    if (mc == nullptr) return fail (PCE_ERR_ALLOC, "Cannot allocate Monte Carlo model.");
    mc->model = model; mc->limit = limit; mc->nequil = nequil; mc->nprod = nprod; mc->nchains = nchains;
    if (seed <= 0) {   // MCModelDefault.c:32-34: time() when no seed is given
        seed = (long long) std::chrono::duration_cast<std::chrono::nanoseconds> (std::chrono::system_clock::now ().time_since_epoch ()).count ();
    }
    mc->seed = (uint64_t) seed;
    find_pairs (mc);
    *out = mc;
    return PCE_OK;
}

extern "C" void pce_mc_destroy (pce_mc *mc) {
    if (mc == nullptr) return;
    mc->d_pair_packed.release (); mc->d_pair_x.release (); mc->d_cross.release (); mc->d_ph.release (); mc->d_esum.release ();
    mc->d_counts.release (); mc->d_counters.release (); mc->d_dw.release (); mc->d_dbase.release (); mc->d_wr.release (); mc->d_rk.release (); mc->d_rf.release ();
    delete mc;
}

extern "C" int pce_mc_npairs (const pce_mc *mc) { return mc ? (int) mc->pair_a.size () : -1; }

extern "C" int pce_mc_get_pair (const pce_mc *mc, int index, int *site_a, int *site_b, double *wmax) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (index < 0 || index >= (int) mc->pair_a.size ()) return fail (PCE_ERR_INDEX, "Index (" + std::to_string (index) + ") out of range.");
    *site_a = mc->pair_a[index]; *site_b = mc->pair_b[index]; *wmax = mc->pair_wmax[index];
    return PCE_OK;
}

extern "C" int pce_mc_set_scans (pce_mc *mc, int nequil, int nprod) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (nequil < 0 || nprod < 1) return fail (PCE_ERR_VALUE, "invalid scan counts");
    mc->nequil = nequil; mc->nprod = nprod;
    return PCE_OK;
}

extern "C" int pce_mc_set_chains (pce_mc *mc, int nchains) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (nchains < 1) return fail (PCE_ERR_VALUE, "invalid chain count");
    mc->nchains = nchains;
    return PCE_OK;
}

extern "C" int pce_mc_set_kernel (pce_mc *mc, int kernel) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (kernel < 0 || kernel > 2) return fail (PCE_ERR_VALUE, "kernel must be 0 (auto), 1 (thread) or 2 (warp)");
    mc->kernel = kernel;
    return PCE_OK;
}

extern "C" int pce_mc_get_info (const pce_mc *mc, int *kernel, int *nchains, int *nequil, int *nprod, long long *seed) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (kernel) *kernel = mc->kernel;
    if (nchains) *nchains = mc->nchains;
    if (nequil) *nequil = mc->nequil;
    if (nprod) *nprod = mc->nprod;
    if (seed) *seed = (long long) mc->seed;
    return PCE_OK;
}

// energy: the launch tracks chain energies (per-chain outputs, or the caller asked for energy sums)
static int make_mc_plan (pce_mc *mc, int nph, bool energy, Plan *plan) {
    size_t limit = 0;
    int status = smem_limit (mc->model->device, &limit);
    if (status != PCE_OK) return status;
    int sms = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, mc->model->device));
    bool ok = false;
    if (mc->kernel == 1) ok = plan_thread (mc, limit, nph, sms, energy, plan, nullptr);
    else if (mc->kernel == 2) ok = plan_warp (mc, limit, nph, sms, plan, nullptr);
    else {
        // auto: price both kernels for this ensemble (CostModel) where both can run the model: the chain-per-thread kernel needs
        // the tables in shared memory and wins on large ensembles, the chain-per-warp kernel finishes small ones 5x sooner
        Plan thread_plan = *plan, warp_plan = *plan;
        double thread_cost = 0.0, warp_cost = 0.0;
        const bool thread_ok = plan_thread (mc, limit, nph, sms, energy, &thread_plan, &thread_cost);
        const bool warp_ok = plan_warp (mc, limit, nph, sms, &warp_plan, &warp_cost);
        if (thread_ok && (!warp_ok || thread_cost <= warp_cost)) { *plan = thread_plan; ok = true; }
        else if (warp_ok) { *plan = warp_plan; ok = true; }
    }
    if (!ok) return fail (PCE_ERR_LIMIT, "model too large for the selected Monte Carlo kernel");
    return PCE_OK;
}

extern "C" int pce_mc_plan (pce_mc *mc, int nph, int *kernel, int *block, int *ctas_per_sm, int *chains_per_wave, long long *smem_bytes) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    PCE_CUDA (cudaSetDevice (mc->model->device));
    Plan plan = { 0, 0, 0, 1, 1, 4, 1, 1 };
    int status = make_mc_plan (mc, nph, false, &plan);       // the quiet path (occupancies only)
    if (status != PCE_OK) return status;
    int sms = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, mc->model->device));
    if (kernel) *kernel = plan.kernel;
    if (block) *block = plan.block;
    if (ctas_per_sm) *ctas_per_sm = plan.ctas_per_sm;
    if (chains_per_wave) *chains_per_wave = sms * plan.resident * (plan.kernel == 1 ? plan.block : plan.block / 32);
    if (smem_bytes) *smem_bytes = (long long) plan.smem;
    return PCE_OK;
}

// Tables of the relative-field formulation, built on the device from the uploaded beta-scaled W whenever the model was
// uploaded again: wr (ninst rows of nrel slots) and the difference rows E[site][(lo, hi)] = wr[hi] - wr[lo] of every
// instance pair of every site, with row stride ldd.  The difference rows are skipped (the warp kernel then streams the two
// wr rows) when they would exceed PCE_DW_BUDGET_MB (default 100 MB: they should stay L2-resident).
static int build_relative_tables (pce_mc *mc, int ldd, bool abs) {
    pce_model *m = mc->model;
    if (mc->rel_generation != m->generation) {
        const int nrel = m->ninst - m->nsites;
        std::vector<int32_t> rk ((size_t) std::max (nrel, 1), 0), rf ((size_t) std::max (nrel, 1), 0);
        for (int i = 0; i < m->nsites; i++)
            for (int k = m->first[i] + 1; k <= m->last[i]; k++) { rk[(size_t) (k - i - 1)] = k; rf[(size_t) (k - i - 1)] = m->first[i]; }
        mc->nrel = nrel;
        mc->ldr  = (std::max (nrel, 1) + 3) & ~3;
        PCE_CUDA (mc->d_rk.resize (rk.size ())); PCE_CUDA (mc->d_rf.resize (rf.size ()));
        PCE_CUDA (mc->d_wr.resize ((size_t) m->ninst * mc->ldr));
        PCE_CUDA (cudaMemcpyAsync (mc->d_rk.ptr, rk.data (), rk.size () * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
        PCE_CUDA (cudaMemcpyAsync (mc->d_rf.ptr, rf.data (), rf.size () * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
        k_build_wr<<<(unsigned) m->ninst, 256, 0, m->stream>>> (m->d_wb.ptr, m->ldw, nrel, mc->ldr, mc->d_rk.ptr, mc->d_rf.ptr, mc->d_wr.ptr);
        pce::count_launch ();
        PCE_CUDA (cudaGetLastError ());
        PCE_CUDA (cudaStreamSynchronize (m->stream));        // the staging vectors die with this scope
        mc->rel_generation = m->generation;
        mc->dw_generation = -1;
    }
    if (mc->dw_generation == m->generation && mc->dw_ldd == ldd && mc->dw_abs == abs) return PCE_OK;
    mc->dw_generation = m->generation;
    mc->dw_ldd = ldd;
    mc->dw_abs = abs;
    mc->dw_valid = false;
    std::vector<int32_t> base ((size_t) m->nsites), lo, hi;
    for (int i = 0; i < m->nsites; i++) {
        base[i] = (int32_t) lo.size ();
        const int n = m->last[i] - m->first[i] + 1;
        for (int mx = 1; mx < n; mx++)
            for (int mn = 0; mn < mx; mn++) { lo.push_back (m->first[i] + mn); hi.push_back (m->first[i] + mx); }
    }
    PCE_CUDA (mc->d_dbase.resize ((size_t) m->nsites));
    PCE_CUDA (cudaMemcpyAsync (mc->d_dbase.ptr, base.data (), (size_t) m->nsites * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
    PCE_CUDA (cudaStreamSynchronize (m->stream));
    const char *budget_env = std::getenv ("PCE_DW_BUDGET_MB");
    const double budget = (budget_env ? std::atof (budget_env) : 100.0) * 1024.0 * 1024.0;
    const size_t nrows = lo.size ();
    mc->nrows = (int) nrows;
    if (nrows == 0) {        // every site has a single instance: nothing can move, an empty table is a valid one
        PCE_CUDA (mc->d_dw.resize (1));
        mc->dw_valid = true;
        return PCE_OK;
    }
    if ((double) nrows * ldd * sizeof (double) > budget) return PCE_OK;
    pce::DeviceBuffer<int32_t> d_lo, d_hi;
    PCE_CUDA (d_lo.resize (nrows)); PCE_CUDA (d_hi.resize (nrows));
    PCE_CUDA (mc->d_dw.resize (nrows * (size_t) ldd));
    PCE_CUDA (cudaMemcpyAsync (d_lo.ptr, lo.data (), nrows * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
    PCE_CUDA (cudaMemcpyAsync (d_hi.ptr, hi.data (), nrows * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
    k_build_dw<<<(unsigned) nrows, 256, 0, m->stream>>> (mc->d_wr.ptr, mc->nrel, mc->ldr, ldd, d_lo.ptr, d_hi.ptr, mc->d_dw.ptr, abs ? 1 : 0, m->ninst,
                                                         m->d_site_of_inst.ptr, m->d_first.ptr);
    pce::count_launch ();
    cudaError_t err = cudaGetLastError ();
    if (err == cudaSuccess) err = cudaStreamSynchronize (m->stream);     // the staging vectors and index buffers die with this scope
    d_lo.release (); d_hi.release ();
    if (err != cudaSuccess) return fail (PCE_ERR_CUDA, std::string ("difference rows: ") + cudaGetErrorString (err));
    mc->dw_valid = true;
    return PCE_OK;
}

static int launch_mc (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset, bool per_chain, McParams &p) {
    pce_model *m = mc->model;
    if (nph <= 0) return PCE_OK;
    if ((long long) mc->nequil + mc->nprod > 4294967295LL) return fail (PCE_ERR_LIMIT, "too many scans");
    // the kernels count the trials of a chain in 32 bits
    if (((double) mc->nequil + (double) mc->nprod) * (double) (m->nsites + (int) mc->pair_a.size ()) >= 4294967295.0 - 256.0)
        return fail (PCE_ERR_LIMIT, "too many trials per chain ((nequil + nprod) x (nsites + npairs) must stay below 2^32)");
    int status = m->upload ();
    if (status != PCE_OK) return status;
    // the W matrix may have changed since the pairs were found (ScaleInteractions etc.): keep the reference's
    // behaviour (pairs are fixed at Initialize, MCModelDefault.pyx:37-54) and only upload them once
    const int npairs = (int) mc->pair_a.size ();
    if (mc->pairs_dirty) {
        PCE_CUDA (mc->d_pair_packed.resize (npairs > 0 ? npairs : 1));
        std::vector<uint32_t> packed ((size_t) (npairs > 0 ? npairs : 1), 0u);
        for (int k = 0; k < npairs; k++) packed[k] = (uint32_t) mc->pair_a[k] | ((uint32_t) mc->pair_b[k] << 16);
        PCE_CUDA (cudaMemcpyAsync (mc->d_pair_packed.ptr, packed.data (), packed.size () * sizeof (uint32_t), cudaMemcpyHostToDevice, m->stream));
        PCE_CUDA (cudaStreamSynchronize (m->stream));
        mc->pairs_dirty = false;
    }
    if (mc->bound_generation != m->generation) {
        // Per pair: every cross term wb[ia][ib] - wb[ia][ob] - wb[oa][ib] + wb[oa][ob], indexed by (old a, draw a, old b,
        // draw b) and evaluated with the kernels' operation order on the same beta-scaled values (one gather instead of
        // four in k_mc_warp), and an upper bound of its magnitude, quantised upwards to 1/4 RT: the warp kernel rejects a
        // double move without reading anything when the bound already decides it.
        const double beta = 1.0 / (PCE_R * m->temperature);
        const size_t n = (size_t) m->ninst;
        std::vector<uint32_t> extra ((size_t) (npairs > 0 ? npairs : 1), 0u);
        std::vector<double>   table;
        bool fits = true;
        for (int k = 0; k < npairs; k++) {
            const int sa = mc->pair_a[k], sb = mc->pair_b[k];
            const int fa = m->first[sa], fb = m->first[sb], na = m->last[sa] - fa + 1, nb = m->last[sb] - fb + 1;
            const size_t base = table.size ();
            double worst = 0.0;
            for (int la = 0; la < na; la++)
                for (int ka = 0; ka < na - 1; ka++)
                    for (int lb = 0; lb < nb; lb++)
                        for (int kb = 0; kb < nb - 1; kb++) {
                            const int oa = fa + la, ob = fb + lb, ia = fa + ka + (fa + ka >= oa ? 1 : 0), ib = fb + kb + (fb + kb >= ob ? 1 : 0);
                            // volatile: every product and difference is rounded on its own (no host-side contraction)
                            volatile double c0 = beta * m->wsym_at (ia, ib), c1 = beta * m->wsym_at (ia, ob);
                            volatile double c2 = beta * m->wsym_at (oa, ib), c3 = beta * m->wsym_at (oa, ob);
                            volatile double t1 = c0 - c1, t2 = t1 - c2;
                            const double cross = t2 + c3;
                            table.push_back (cross);
                            worst = std::max (worst, std::fabs (cross));
                        }
            const double q = std::ceil (worst * (1.0 + 1e-6) * 4.0 + 1e-9);       // >= 1; 0 is reserved for "no shortcut"
            if (base >= (1u << 24)) fits = false;
            extra[k] = (uint32_t) (base & 0xffffffu) | ((q <= 255.0 ? (uint32_t) q : 0u) << 24);
        }
        const char *table_env = std::getenv ("PCE_CROSS_TABLE");          // 0: gather the four W entries instead (test hook)
        mc->cross_valid = fits && npairs > 0 && !(table_env && std::atoi (table_env) == 0);
        mc->cross_overflow = npairs > 0 && !mc->cross_valid;
        PCE_CUDA (mc->d_pair_x.resize (extra.size ()));
        PCE_CUDA (cudaMemcpyAsync (mc->d_pair_x.ptr, extra.data (), extra.size () * sizeof (uint32_t), cudaMemcpyHostToDevice, m->stream));
        if (mc->cross_valid) {
            PCE_CUDA (mc->d_cross.resize (table.size ()));
            mc->ncross = table.size ();
            PCE_CUDA (cudaMemcpyAsync (mc->d_cross.ptr, table.data (), table.size () * sizeof (double), cudaMemcpyHostToDevice, m->stream));
        }
        PCE_CUDA (cudaStreamSynchronize (m->stream));
        mc->bound_generation = m->generation;
    }
    PCE_CUDA (mc->d_ph.resize (nph));
    PCE_CUDA (cudaMemcpyAsync (mc->d_ph.ptr, ph, nph * sizeof (double), cudaMemcpyHostToDevice, m->stream));

    const bool energy = per_chain || p.esum != nullptr;
    Plan plan = { 0, 0, 0, 1, 1, 4, 1, 1 };
    status = make_mc_plan (mc, nph, energy, &plan);
    if (status != PCE_OK) return status;

    p.nsites = m->nsites; p.ninst = m->ninst; p.npairs = npairs; p.ldw = m->ldw;
    p.first = m->d_first.ptr; p.nsi = m->d_nsite_inst.ptr; p.protons = m->d_protons.ptr;
    p.pair_packed = mc->d_pair_packed.ptr; p.pair_x = mc->d_pair_x.ptr; p.cross = mc->cross_valid ? mc->d_cross.ptr : nullptr;
    p.gintr = m->d_gintr.ptr; p.w = m->d_w.ptr; p.wb = m->d_wb.ptr; p.ph = mc->d_ph.ptr;
    p.nph = nph; p.ph_index_offset = ph_index_offset; p.nchains = mc->nchains; p.nequil = mc->nequil; p.nprod = mc->nprod;
    p.chain_offset = chain_offset;
    p.k0 = (uint32_t) mc->seed; p.k1 = (uint32_t) (mc->seed >> 32);
    p.temperature = m->temperature;

    {
        // row stride of the difference rows: the padded field length of a chain-per-warp launch, else the slot count
        const int nrel = std::max (m->ninst - m->nsites, 1);
        const bool abs = plan.kernel == 2 && plan.upd != 16;         // small-field variants of the warp kernel: instance-indexed slots
        const int ldd = plan.kernel == 2 ? WarpLayout (m->nsites, m->ninst, plan.tpl, plan.upd, abs, (!abs && mc->nprod < 65536) ? 2 : 4).ldh : ((nrel + 3) & ~3);
        status = build_relative_tables (mc, ldd, abs);
        if (status != PCE_OK) return status;
        if (plan.kernel == 1 && (!mc->dw_valid || (npairs > 0 && !mc->cross_valid)))
            return fail (PCE_ERR_LIMIT, "the chain-per-thread kernel needs the difference rows and the cross-term table");
        if (abs && !mc->dw_valid) return fail (PCE_ERR_LIMIT, "the small-field chain-per-warp kernel needs the difference rows");
        p.nrel = mc->nrel; p.ldr = mc->ldr; p.nrows = mc->nrows; p.ncross = (int) mc->ncross;
        p.wr = mc->d_wr.ptr; p.rk = mc->d_rk.ptr; p.rf = mc->d_rf.ptr; p.site_of = m->d_site_of_inst.ptr;
        p.dw = mc->dw_valid ? mc->d_dw.ptr : nullptr;
        p.dbase = mc->d_dbase.ptr;
        p.ldd = ldd;
    }
    long long blocks;
    if (plan.kernel == 1) blocks = ((long long) nph * mc->nchains + plan.block - 1) / plan.block;     // chain-major linear grid
    else blocks = ((long long) nph * mc->nchains + plan.block / 32 - 1) / (plan.block / 32);
    if (blocks > 2147483647LL) return fail (PCE_ERR_LIMIT, "too many chains in one launch");
    p.nmoves = (uint32_t) (m->nsites + npairs);
    p.ntrials = (unsigned long long) ((long long) mc->nequil + mc->nprod) * p.nmoves;
    p.nmoves_magic = (p.nmoves >= 2u && p.nmoves <= 64u) ? (uint32_t) (4294967296ull / p.nmoves) + 1u : 0u;
    if (plan.kernel == 2) {
        const bool since16 = plan.upd == 16 && mc->nprod < 65536;        // (large-field build only)
        const WarpLayout WL (m->nsites, m->ninst, plan.tpl, plan.upd, plan.upd != 16, since16 ? 2 : 4);
        p.wl_since16 = since16 ? 1u : 0u;
        p.wl_shared = (uint32_t) WL.shared_bytes; p.wl_per_warp = (uint32_t) WL.per_warp; p.wl_since = (uint32_t) WL.off_since;
        p.wl_state = (uint32_t) WL.off_state; p.wl_ring = (uint32_t) WL.off_ring; p.wl_ldh = WL.ldh;
    }
    if (plan.kernel == 1) {
        const ThreadLayout TL (m->nsites, m->ninst, npairs, p.nrel, p.nrows, p.ncross, plan.block, mc->nchains, nph, mc->nprod < 65536 ? 2 : 4, plan.ph_minor != 0, energy);
        p.ph_minor = plan.ph_minor;
        p.tl_e = (uint32_t) TL.off_e; p.tl_cross = (uint32_t) TL.off_cross; p.tl_h = (uint32_t) TL.off_h; p.tl_eng = (uint32_t) TL.off_eng; p.tl_counts = (uint32_t) TL.off_counts; p.tl_nact = (uint32_t) TL.off_nact;
        p.tl_since = (uint32_t) TL.off_since; p.tl_move = (uint32_t) TL.off_move;
        p.tl_erow = (uint32_t) TL.off_erow; p.tl_state = (uint32_t) TL.off_state;
        p.tl_hstride = TL.hstride; p.tl_estride = TL.estride; p.tl_kslots = TL.kslots;
        if (TL.total != plan.smem) return fail (PCE_ERR_STATE, "shared-memory layout of the chain-per-thread kernel changed between planning and launch");

    }
    if (plan.kernel == 1) {
        void (*kernel) (const McParams);
        const int short_fields = p.nrel <= 16 ? 2 : p.nrel <= 32 ? 1 : 0;      // one field slot per lane (half-warp) when an accepted move is applied
        // the first build (fewest instructions first) whose registers fit: a quarter of the block's warps share the 16384
        // registers of an SM sub-partition, allocated in units of 8 per thread
        auto fits = [&] (void (*candidate) (const McParams)) -> bool {
            cudaFuncAttributes attributes;
            if (cudaFuncGetAttributes (&attributes, candidate) != cudaSuccess) return false;
            const int allocated = (attributes.numRegs + 7) & ~7, warps_per_partition = (plan.block / 32 + 3) / 4;
            return warps_per_partition * allocated * 32 <= 16384;
        };
#define PCE_THREAD_VARIANT(S_, R_, F_) (per_chain ? k_mc_thread<true, S_, R_, F_, true> : energy ? k_mc_thread<false, S_, R_, F_, true> : k_mc_thread<false, S_, R_, F_, false>)
#define PCE_THREAD_KERNEL(S_, R_) (short_fields == 2 ? PCE_THREAD_VARIANT (S_, R_, 2) : short_fields == 1 ? PCE_THREAD_VARIANT (S_, R_, 1) : PCE_THREAD_VARIANT (S_, R_, 0))
#define PCE_THREAD_PICK(S_) (fits (PCE_THREAD_KERNEL (S_, 128)) ? PCE_THREAD_KERNEL (S_, 128) : fits (PCE_THREAD_KERNEL (S_, 80)) ? PCE_THREAD_KERNEL (S_, 80) \
                             : fits (PCE_THREAD_KERNEL (S_, 72)) ? PCE_THREAD_KERNEL (S_, 72) : PCE_THREAD_KERNEL (S_, 64))
        if (mc->nprod < 65536) kernel = PCE_THREAD_PICK (uint16_t);
        else kernel = PCE_THREAD_PICK (uint32_t);
#undef PCE_THREAD_PICK
#undef PCE_THREAD_KERNEL
#undef PCE_THREAD_VARIANT
        PCE_CUDA (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) plan.smem));
        kernel<<<(unsigned) blocks, plan.block, plan.smem, m->stream>>> (p);
    } else {
        // one resident CTA (large fields): up to 255 registers; two: 128
        void (*kernel) (const McParams);
#define PCE_WARP_VARIANT(C_, TPL_, T_) (per_chain ? k_mc_warp<true, C_, TPL_, true, T_> : energy ? k_mc_warp<false, C_, TPL_, true, T_> : k_mc_warp<false, C_, TPL_, false, T_>)
#define PCE_WARP_SMALL(C_, TPL_) (plan.upd == 1 ? PCE_WARP_VARIANT (C_, TPL_, true) : PCE_WARP_VARIANT (C_, TPL_, false))
#define PCE_WARP_KERNEL(TPL_) (plan.upd == 16 ? PCE_WARP_VARIANT (1, TPL_, false) : plan.ctas_per_sm >= 4 ? PCE_WARP_SMALL (4, TPL_) \
                              : plan.ctas_per_sm == 3 ? PCE_WARP_SMALL (3, TPL_) : PCE_WARP_SMALL (2, TPL_))
        if (plan.tpl == 2) kernel = PCE_WARP_KERNEL (2);
        else kernel = PCE_WARP_KERNEL (1);
#undef PCE_WARP_KERNEL
#undef PCE_WARP_SMALL
#undef PCE_WARP_VARIANT
        PCE_CUDA (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) plan.smem));
        // shared-memory carveout for exactly the planned number of resident CTAs (the rest of the 228 KB stays L1 for the
        // difference rows and the cross-term table); without __launch_bounds__ the driver has no occupancy hint of its own
        const int carveout = (int) std::min<size_t> (100, ((size_t) plan.resident * (plan.smem + 1024) * 100) / (228 * 1024) + 1);
        PCE_CUDA (cudaFuncSetAttribute (kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
        kernel<<<(unsigned) blocks, plan.block, plan.smem, m->stream>>> (p);
    }
    pce::count_launch ();
    PCE_CUDA (cudaGetLastError ());
    return PCE_OK;
}

extern "C" int pce_mc_run_device (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset,
                                  long long *d_counts, long long *d_counters, double *d_esum) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    pce_model *m = mc->model;
    PCE_CUDA (cudaSetDevice (m->device));
    if (d_counts) PCE_CUDA (cudaMemsetAsync (d_counts, 0, (size_t) nph * m->ninst * sizeof (long long), m->stream));
    if (d_counters) PCE_CUDA (cudaMemsetAsync (d_counters, 0, (size_t) nph * 4 * sizeof (long long), m->stream));
    if (d_esum) PCE_CUDA (cudaMemsetAsync (d_esum, 0, (size_t) nph * sizeof (double), m->stream));
    McParams p;
    std::memset (&p, 0, sizeof (p));
    p.counts = (unsigned long long *) d_counts; p.counters = d_counters; p.esum = d_esum;
    return launch_mc (mc, ph, nph, ph_index_offset, chain_offset, false, p);
}

extern "C" int pce_mc_run (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset,
                           long long *counts, long long *counters, double *esum) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    pce_model *m = mc->model;
    if (nph <= 0) return PCE_OK;
    PCE_CUDA (cudaSetDevice (m->device));
    PCE_CUDA (mc->d_counts.resize ((size_t) nph * m->ninst));
    PCE_CUDA (mc->d_counters.resize ((size_t) nph * 4));
    PCE_CUDA (mc->d_esum.resize (nph));
    // energy sums only when the caller takes them (the quiet path of the reference returns occupancies only)
    int status = pce_mc_run_device (mc, ph, nph, ph_index_offset, chain_offset, (long long *) mc->d_counts.ptr, mc->d_counters.ptr, esum ? mc->d_esum.ptr : nullptr);
    if (status != PCE_OK) return status;
    std::vector<long long> host_counts ((size_t) nph * m->ninst);
    PCE_CUDA (cudaMemcpyAsync (host_counts.data (), mc->d_counts.ptr, host_counts.size () * sizeof (long long), cudaMemcpyDeviceToHost, m->stream));
    if (counters) PCE_CUDA (cudaMemcpyAsync (counters, mc->d_counters.ptr, (size_t) nph * 4 * sizeof (long long), cudaMemcpyDeviceToHost, m->stream));
    if (esum) PCE_CUDA (cudaMemcpyAsync (esum, mc->d_esum.ptr, (size_t) nph * sizeof (double), cudaMemcpyDeviceToHost, m->stream));
    PCE_CUDA (cudaStreamSynchronize (m->stream));
    if (counts) std::memcpy (counts, host_counts.data (), host_counts.size () * sizeof (long long));
    // probabilities of the last pH point into the model (MCModelDefault.c:310-311: counts / nprod)
    const double scale = 1.0 / ((double) mc->nchains * (double) mc->nprod);
    for (int k = 0; k < m->ninst; k++) m->prob[k] = (double) host_counts[(size_t) (nph - 1) * m->ninst + k] * scale;
    return PCE_OK;
}

extern "C" int pce_mc_run_chains (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset,
                                  long long *chain_counts, int32_t *chain_state, double *chain_energy, double *chain_esum, long long *chain_counters) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    pce_model *m = mc->model;
    if (nph <= 0) return PCE_OK;
    PCE_CUDA (cudaSetDevice (m->device));
    const size_t nc = (size_t) nph * mc->nchains;
    pce::DeviceBuffer<long long> d_counts, d_counters;
    pce::DeviceBuffer<int32_t>   d_state;
    pce::DeviceBuffer<double>    d_energy, d_esum;
    PCE_CUDA (d_counts.resize (nc * m->ninst)); PCE_CUDA (d_counters.resize (nc * 4));
    PCE_CUDA (d_state.resize (nc * m->nsites)); PCE_CUDA (d_energy.resize (nc)); PCE_CUDA (d_esum.resize (nc));
    PCE_CUDA (cudaMemsetAsync (d_counts.ptr, 0, nc * m->ninst * sizeof (long long), m->stream));
    McParams p;
    std::memset (&p, 0, sizeof (p));
    p.chain_counts = d_counts.ptr; p.chain_state = d_state.ptr; p.chain_energy = d_energy.ptr; p.chain_esum = d_esum.ptr;
    p.chain_counters = d_counters.ptr;
    int status = launch_mc (mc, ph, nph, ph_index_offset, chain_offset, true, p);
    if (status == PCE_OK) {
        cudaError_t err = cudaStreamSynchronize (m->stream);
        if (err == cudaSuccess && chain_counts) err = cudaMemcpy (chain_counts, d_counts.ptr, nc * m->ninst * sizeof (long long), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && chain_state) err = cudaMemcpy (chain_state, d_state.ptr, nc * m->nsites * sizeof (int32_t), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && chain_energy) err = cudaMemcpy (chain_energy, d_energy.ptr, nc * sizeof (double), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && chain_esum) err = cudaMemcpy (chain_esum, d_esum.ptr, nc * sizeof (double), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && chain_counters) err = cudaMemcpy (chain_counters, d_counters.ptr, nc * 4 * sizeof (long long), cudaMemcpyDeviceToHost);
        if (err != cudaSuccess) status = fail (PCE_ERR_CUDA, std::string ("mc per-chain readback: ") + cudaGetErrorString (err));
    }
    d_counts.release (); d_counters.release (); d_state.release (); d_energy.release (); d_esum.release ();
    return status;
}

extern "C" int pce_mc_run_trajectory (pce_mc *mc, double ph, int ph_index_offset, double *energies, long long *counts, long long *counters) {
    return pce_mc_run_logged (mc, ph, ph_index_offset, 0, energies, counts, counters, nullptr);
}

extern "C" int pce_mc_run_logged (pce_mc *mc, double ph, int ph_index_offset, int log_frequency, double *energies, long long *counts, long long *counters,
                                  long long *log_rows) {
    if (mc == nullptr) return fail (PCE_ERR_VALUE, "null sampler");
    if (log_rows != nullptr && log_frequency < 1) return fail (PCE_ERR_VALUE, "logFrequency must be positive");
    pce_model *m = mc->model;
    PCE_CUDA (cudaSetDevice (m->device));
    const int saved_chains = mc->nchains;
    mc->nchains = 1;                                   // the reference's trajectory is one chain (MCModelDefault.pyx:79-159)
    pce::DeviceBuffer<long long> d_counts, d_counters, d_log;
    pce::DeviceBuffer<int32_t>   d_state;
    pce::DeviceBuffer<double>    d_energy, d_esum, d_traj;
    const size_t nrows = log_rows ? (size_t) (mc->nequil / log_frequency + mc->nprod / log_frequency) : 0;
    int status = PCE_OK;
    do {
        cudaError_t err = d_counts.resize (m->ninst);
        if (err == cudaSuccess) err = d_log.resize (nrows * 5 + 1);
        if (err == cudaSuccess) err = cudaMemsetAsync (d_log.ptr, 0, (nrows * 5 + 1) * sizeof (long long), m->stream);
        if (err == cudaSuccess) err = d_counters.resize (4);
        if (err == cudaSuccess) err = d_state.resize (m->nsites);
        if (err == cudaSuccess) err = d_energy.resize (1);
        if (err == cudaSuccess) err = d_esum.resize (1);
        if (err == cudaSuccess) err = d_traj.resize (mc->nprod);
        if (err == cudaSuccess) err = cudaMemsetAsync (d_counts.ptr, 0, m->ninst * sizeof (long long), m->stream);
        if (err != cudaSuccess) { status = fail (PCE_ERR_CUDA, std::string ("trajectory buffers: ") + cudaGetErrorString (err)); break; }
        McParams p;
        std::memset (&p, 0, sizeof (p));
        p.chain_counts = d_counts.ptr; p.chain_state = d_state.ptr; p.chain_energy = d_energy.ptr; p.chain_esum = d_esum.ptr;
        p.chain_counters = d_counters.ptr; p.trajectory = d_traj.ptr;
        if (log_rows) { p.log_rows = d_log.ptr; p.log_frequency = log_frequency; }
        status = launch_mc (mc, &ph, 1, ph_index_offset, 0, true, p);
        if (status != PCE_OK) break;
        err = cudaStreamSynchronize (m->stream);
        if (err == cudaSuccess && energies) err = cudaMemcpy (energies, d_traj.ptr, mc->nprod * sizeof (double), cudaMemcpyDeviceToHost);
        std::vector<long long> host_counts (m->ninst);
        if (err == cudaSuccess) err = cudaMemcpy (host_counts.data (), d_counts.ptr, m->ninst * sizeof (long long), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && counters) err = cudaMemcpy (counters, d_counters.ptr, 4 * sizeof (long long), cudaMemcpyDeviceToHost);
        if (err == cudaSuccess && log_rows && nrows) err = cudaMemcpy (log_rows, d_log.ptr, nrows * 5 * sizeof (long long), cudaMemcpyDeviceToHost);
        if (err != cudaSuccess) { status = fail (PCE_ERR_CUDA, std::string ("trajectory readback: ") + cudaGetErrorString (err)); break; }
        if (counts) std::memcpy (counts, host_counts.data (), m->ninst * sizeof (long long));
        for (int k = 0; k < m->ninst; k++) m->prob[k] = (double) host_counts[k] / (double) mc->nprod;
    } while (0);
    mc->nchains = saved_chains;
    d_counts.release (); d_counters.release (); d_state.release (); d_energy.release (); d_esum.release (); d_traj.release (); d_log.release ();
    return status;
}
