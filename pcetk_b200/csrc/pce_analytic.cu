// pce_analytic.cu -- kernel 2: exhaustive enumeration of the partition function.
//
// Replaces EnergyModel_CalculateZ / _CalculateProbabilitiesFromZ / _CalculateProbabilitiesAnalytically
// (extensions/csource/EnergyModel.c:252-281, 318-337, 342-371).  The reference recomputes every
// microstate energy from scratch (O(nsites^2/2) gathers per state, its own TODO at :264-267), stores
// nstates Boltzmann factors, and repeats all of it for every pH value.  This kernel does neither:
//
//   (1) G_s(pH) = A_s + n_s * R T ln10 * pH, so states are BINNED BY PROTON COUNT n_s and one
//       enumeration at a reference pH* serves the whole pH grid:
//           Z(pH)    = sum_n  Zbin[n]    * 10^(-n (pH - pH*)),
//           p_a(pH)  = sum_n  Sbin[a][n] * 10^(-n (pH - pH*)) / Z(pH);
//   (2) sites are split into roles -- M (one configuration per thread, fixed for the kernel), I (a few
//       configurations looped in registers), O (looped by the CTA, "tiles"), F (fixed per CTA, "chunk") --
//       and the energy of a state is a sum of terms that each depend on one role group plus static
//       cross terms.  The Boltzmann factor is therefore a PRODUCT of table entries:
//           b = cOut * eLo[m_lo] * eHi[m_hi] * xMM[m] * eI[i] * xMI[m][i]
//       with a handful of exp() per tile instead of one per state;
//   (3) per-instance sums need no work in the inner loop for sites whose instance is fixed per thread (M)
//       or per CTA (F): they fall out of the per-thread bins at the end.  The sites that were looped (I, O)
//       are covered by a second enumeration with the roles rotated.
//
// Every partial result is a plain sum with a common reference energy, so shards of the chunk range -- on
// one GPU or many -- combine by addition (one all-reduce).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "pce_common.cuh"

using pce::fail;

namespace {

constexpr int kBlock     = 256;   // threads per CTA = max configurations of the M group
constexpr int kMaxM      = 8;
constexpr int kMaxI      = 4;
constexpr int kIStates   = 16;    // max configurations of the I group (per-thread register table)
constexpr int kMaxOuter  = 96;    // max sites in O + F
constexpr int kMaxMI     = 320;   // max instances in M + I sites
constexpr long long kMaxChunks = 16384;

constexpr int kTB = 8;            // tiles (outer configurations) processed per synchronisation round

struct EnumParams {
    int nsites, ninst, ldw, unfolded;
    const double  *gt;          // tilted intrinsic energies G'[a] = G[a] - protons[a]*mu(pH*)
    const double  *w;
    const int32_t *protons, *first;
    int   nM, nLo, nI, nO, nF;  // group sizes; M = Lo (first nLo) + Hi
    int   P_lo, P_hi, P_M, P_I;
    long long P_O;              // tiles per chunk
    long long chunk_begin, nchunks;
    int   Nb, NbRel;            // proton-count bins: absolute (output) / per thread, relative to the thread's M protons
    int   wstage;               // W sub-blocks (M/I x O, O x O) staged in shared memory per chunk
    double beta, gref;
    int16_t siteM[kMaxM], radM[kMaxM];
    int16_t siteI[kMaxI], radI[kMaxI];
    int16_t siteOut[kMaxOuter], radOut[kMaxOuter];   // O sites first (fastest digit first), then F sites
    int16_t firstOut[kMaxOuter];                     // first instance of each outer site
    unsigned strideO[kMaxOuter];                     // digit stride of each O site within the tile index
    int16_t permI[kIStates], binI[kIStates];         // I configurations sorted by proton count; protons of each
    unsigned flushI;                                 // bit i set: sorted configuration i is the last of its proton count
    int      nMI, pow2O, nOinst, canon;              // instances in M+I sites; every O radix a power of two; instances in O sites; canonical I group (-1: no)
    uint8_t  shiftO[kMaxOuter];                      // digit shift of each O site when pow2O
    double *T;                  // [nchunks][NbRel][kBlock] per-thread bins
    // per-thread statics, computed once by k_enum_static: xs[(1 + kIStates) * kBlock] = xMM then xMI[i], coalesced;
    // ms[3 * kBlock] = m_lo, m_hi, protons of the M configuration; shift = minMM + minMI
    double  *xs;
    int32_t *ms;
    double  *shift;
};

__device__ __forceinline__ double block_min (double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v = fmin (v, __shfl_xor_sync (0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads ();
    double r = scratch[0];
    for (int k = 1; k < kBlock / 32; k++) r = fmin (r, scratch[k]);
    __syncthreads ();
    return r;
}

// Static per-thread factors: they depend on the thread's M configuration only, so they are computed once per
// enumeration (one CTA) instead of once per chunk.
__global__ void __launch_bounds__ (kBlock) k_enum_static (const EnumParams p) {
    __shared__ double s_scratch[kBlock / 32];
    const int tid = threadIdx.x, ldw = p.ldw;
    const bool folded = !p.unfolded, mine = tid < p.P_M;
    int    aM[kMaxM];
    int    m_lo = 0, m_hi = 0, nMp = 0;
    double eMM = 0.0;
    {
        int x = mine ? tid : 0, stride_lo = 1, stride_hi = 1;
        for (int i = 0; i < kMaxM; i++) {
            aM[i] = 0;
            if (i < p.nM) {
                const int d = x % p.radM[i];
                x /= p.radM[i];
                aM[i] = p.first[p.siteM[i]] + d;
                nMp += p.protons[aM[i]];
                if (i < p.nLo) { m_lo += d * stride_lo; stride_lo *= p.radM[i]; } else { m_hi += d * stride_hi; stride_hi *= p.radM[i]; }
            }
        }
        if (folded)
            for (int i = 0; i < p.nM; i++)
                for (int j = 0; j < i; j++) eMM += p.w[(size_t) aM[i] * ldw + aM[j]];
    }
    // static cross terms E_MI(m,i) + E_II(i)
    double xMI[kIStates];
    for (int i = 0; i < kIStates; i++) {
        double e = 0.0;
        if (i < p.P_I && folded) {
            int aI[kMaxI];
            int x = p.permI[i];                                  // sorted position -> I configuration
            for (int k = 0; k < kMaxI; k++) { aI[k] = 0; if (k < p.nI) { aI[k] = p.first[p.siteI[k]] + x % p.radI[k]; x /= p.radI[k]; } }
            for (int k = 0; k < p.nI; k++) {
                const double *row = p.w + (size_t) aI[k] * ldw;
                for (int l = 0; l < k; l++) e += row[aI[l]];
                for (int l = 0; l < p.nM; l++) e += row[aM[l]];
            }
        }
        xMI[i] = e;
    }
    // normalise by the minima over all (m) and (m,i): every factor <= 1; the minima go into the tile constant
    double vmin = mine ? eMM : DBL_MAX;
    const double minMM = block_min (vmin, s_scratch);
    vmin = DBL_MAX;
    for (int i = 0; i < kIStates; i++) if (mine && i < p.P_I) vmin = fmin (vmin, xMI[i]);
    const double minMI = block_min (vmin, s_scratch);
    p.xs[tid] = mine ? exp (-p.beta * (eMM - minMM)) : 0.0;
    for (int i = 0; i < kIStates; i++) p.xs[(1 + i) * kBlock + tid] = (mine && i < p.P_I) ? exp (-p.beta * (xMI[i] - minMI)) : 0.0;
    p.ms[tid] = m_lo; p.ms[kBlock + tid] = m_hi; p.ms[2 * kBlock + tid] = nMp;
    if (tid == 0) *p.shift = minMM + minMI;
}

constexpr int kMaxO = 32;         // max sites looped by the CTA (O); larger problems put more sites into F

// Dynamic shared-memory carve-up of k_enum (same arithmetic on host and device).
struct EnumLayout {
    size_t off_bins, off_gF, off_wMO, off_wOO, off_eInst, off_smin, off_hF, off_eLo, off_eHi, off_eI, off_cOut, off_cExp, off_inst, off_aF, off_nOut,
           off_idx, off_pair, off_instO, total;
    int    nMI, nIdx, nSites, nOinst, nPairs;
    __host__ __device__ EnumLayout (int Nb, int nMI_, int P_lo, int P_hi, int P_I, int nF, int nM, int nI, int nO, int nOinst_, int wstage) {
        nMI = nMI_; nIdx = P_lo + P_hi + P_I; nSites = nM + nI; nOinst = nOinst_; nPairs = nO * (nO - 1) / 2;
        size_t o = 0;
        off_bins = o; o += (size_t) Nb * kBlock * 8;
        off_gF = o;   o += (size_t) nMI * 8;
        off_wMO = o;  o += wstage ? (size_t) nMI * nOinst_ * 8 : 0;
        off_wOO = o;  o += wstage ? (size_t) nOinst_ * nOinst_ * 8 : 0;
        off_eInst = o; o += (size_t) kTB * nMI * 8;
        off_smin = o; o += (size_t) kTB * (nSites > 0 ? nSites : 1) * 8;
        off_hF = o;   o += (size_t) (nOinst > 0 ? nOinst : 1) * 8;
        off_eLo = o;  o += (size_t) kTB * P_lo * 8;
        off_eHi = o;  o += (size_t) kTB * P_hi * 8;
        o = (o + 15) & ~(size_t) 15;
        off_eI = o;   o += (size_t) kTB * 16 * 8;            // padded to 16 entries, 16-byte aligned (128-bit loads)
        off_cOut = o; o += (size_t) kTB * 8;
        off_cExp = o; o += (size_t) kTB * 8;
        off_inst = o; o += (size_t) nMI * 4;
        off_aF = o;   o += (size_t) (nF > 0 ? nF : 1) * 4;
        off_nOut = o; o += (size_t) kTB * 4;
        off_instO = o; o += (size_t) (nOinst > 0 ? nOinst : 1) * 4;
        off_idx = o;  o += (size_t) nIdx * kMaxM * 2;      // per table entry: compact-list index of each of its sites
        off_pair = o; o += (size_t) (nPairs > 0 ? nPairs : 1) * 2;
        total = (o + 15) & ~(size_t) 15;
    }
};

__host__ __device__ constexpr unsigned canon_flush_mask (int n) {
    // I configurations of n binary one-proton sites sorted by proton count: bins of C(n,k) entries
    return n == 4 ? 0xC411u : n == 3 ? 0xC9u : n == 2 ? 0xDu : n == 1 ? 0x3u : 0x1u;
}

// One CTA per chunk (configuration of the F sites).  See the file header for the algorithm.  Tiles (O
// configurations) are processed kTB per round: fields and tables for kTB tiles are built between two barriers,
// then every thread accumulates its kTB x P_I states.  CANON >= 0: the I group is CANON binary sites whose two
// instances differ by one proton, so the proton-count structure of the I configurations is known at compile time.
template <int CANON>
__global__ void __launch_bounds__ (kBlock, 3) k_enum (const EnumParams p) {
    static_assert (kTB == kBlock / 32, "one warp per tile in phase (b)");
    extern __shared__ __align__ (16) unsigned char smem[];
    const EnumLayout L (p.NbRel, p.nMI, p.P_lo, p.P_hi, p.P_I, p.nF, p.nM, p.nI, p.nO, p.nOinst, p.wstage);
    double   *bins   = (double *) (smem + L.off_bins);       // [NbRel][kBlock]
    double   *s_gF   = (double *) (smem + L.off_gF);         // F-site fields on the M/I instances (per chunk)
    double   *s_wMO  = (double *) (smem + L.off_wMO);        // [nMI][nOinst] W between M/I instances and O instances
    double   *s_wOO  = (double *) (smem + L.off_wOO);        // [nOinst][nOinst]
    double   *s_eInst = (double *) (smem + L.off_eInst);     // [kTB][nMI] exp(-beta (gOut - site minimum)) per instance
    double   *s_smin = (double *) (smem + L.off_smin);       // [kTB][nM+nI] per-site minimum of gOut
    double   *s_hF   = (double *) (smem + L.off_hF);         // F-site fields + intrinsic energy on the O instances
    double   *s_eLo  = (double *) (smem + L.off_eLo);        // [kTB][P_lo]
    double   *s_eHi  = (double *) (smem + L.off_eHi);        // [kTB][P_hi]
    double   *s_eI   = (double *) (smem + L.off_eI);         // [kTB][16]
    double   *s_cOut = (double *) (smem + L.off_cOut);
    double   *s_cExp = (double *) (smem + L.off_cExp);
    int32_t  *s_inst = (int32_t *) (smem + L.off_inst);      // compact list of the M and I instances
    int32_t  *s_aF   = (int32_t *) (smem + L.off_aF);        // active instance of every F site
    int32_t  *s_nOut = (int32_t *) (smem + L.off_nOut);
    int32_t  *s_instO = (int32_t *) (smem + L.off_instO);    // compact list of the O instances (global index | protons << 16)
    uint16_t *s_idx  = (uint16_t *) (smem + L.off_idx);      // [nIdx][kMaxM]
    uint16_t *s_pair = (uint16_t *) (smem + L.off_pair);     // O-O pairs: o | o2 << 8
    __shared__ int32_t s_off[kMaxM + kMaxI + 1];             // start of each M / I site in the compact list
    __shared__ int32_t s_offO[kMaxO + 1];
    __shared__ uint32_t s_oFirst[kMaxO], s_oShift[kMaxO], s_oRad[kMaxO], s_oStride[kMaxO];   // O-site tables (lane-indexed in phase b)
    __shared__ double  s_AF;
    __shared__ int     s_nF;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldw = p.ldw, nMI = p.nMI, nS = p.nM + p.nI;
    const bool folded = !p.unfolded;
    const long long chunk = p.chunk_begin + blockIdx.x;

    // ---- setup -----------------------------------------------------------------------------------------
    if (tid == 0) {
        int o = 0;
        for (int i = 0; i < p.nM; i++) { s_off[i] = o; for (int k = 0; k < p.radM[i]; k++) s_inst[o++] = p.first[p.siteM[i]] + k; }
        for (int i = 0; i < p.nI; i++) { s_off[p.nM + i] = o; for (int k = 0; k < p.radI[i]; k++) s_inst[o++] = p.first[p.siteI[i]] + k; }
        s_off[nS] = o;
        long long c = chunk;                                 // F digits of this chunk
        for (int j = 0; j < p.nF; j++) { s_aF[j] = p.firstOut[p.nO + j] + (int) (c % p.radOut[p.nO + j]); c /= p.radOut[p.nO + j]; }
        o = 0;
        int np = 0;
        for (int j = 0; j < p.nO; j++) {
            s_oFirst[j] = (uint32_t) p.firstOut[j]; s_oShift[j] = p.shiftO[j]; s_oRad[j] = (uint32_t) p.radOut[j]; s_oStride[j] = p.strideO[j];
            s_offO[j] = o;
            for (int k = 0; k < p.radOut[j]; k++) { const int a = p.firstOut[j] + k; s_instO[o++] = a | (p.protons[a] << 16); }
            for (int j2 = 0; j2 < j; j2++) s_pair[np++] = (uint16_t) (j | (j2 << 8));
        }
        s_offO[p.nO] = o;
    }
    for (int e = tid; e < p.NbRel * kBlock; e += kBlock) bins[e] = 0.0;
    __syncthreads ();
    // index tables: which compact-list entries make up table entry x (no divisions in the tile loop)
    for (int x = tid; x < L.nIdx; x += kBlock) {
        uint16_t *row = s_idx + x * kMaxM;
        if (x < p.P_lo) {
            int y = x;
            for (int i = 0; i < p.nLo; i++) { row[i] = (uint16_t) (s_off[i] + y % p.radM[i]); y /= p.radM[i]; }
        } else if (x < p.P_lo + p.P_hi) {
            int y = x - p.P_lo;
            for (int i = p.nLo; i < p.nM; i++) { row[i - p.nLo] = (uint16_t) (s_off[i] + y % p.radM[i]); y /= p.radM[i]; }
        } else {
            int y = p.permI[x - p.P_lo - p.P_hi];                // sorted position -> I configuration
            for (int i = 0; i < p.nI; i++) { row[i] = (uint16_t) (s_off[p.nM + i] + y % p.radI[i]); y /= p.radI[i]; }
        }
    }
    // this thread's statics
    const bool mine = tid < p.P_M;
    const double eMM = p.xs[tid];
    double xMI[kIStates];
#pragma unroll
    for (int i = 0; i < kIStates; i++) xMI[i] = p.xs[(1 + i) * kBlock + tid];
    const int m_lo = p.ms[tid], m_hi = p.ms[kBlock + tid];
    const double static_shift = *p.shift;

    // W sub-blocks used by the tile loop, staged once per CTA (they do not depend on the chunk): with shared memory
    // full there is no L1 left, so every global W read in the loop would be an L2 round trip
    const int nOi = L.nOinst;
    if (p.wstage && folded) {
        for (int e = tid; e < nMI * nOi; e += kBlock) s_wMO[e] = p.w[(size_t) s_inst[e / nOi] * ldw + (s_instO[e % nOi] & 0xffff)];
        for (int e = tid; e < nOi * nOi; e += kBlock) s_wOO[e] = p.w[(size_t) (s_instO[e / nOi] & 0xffff) * ldw + (s_instO[e % nOi] & 0xffff)];
    }
    // per chunk: fields of the F sites on the M/I instances and on the O instances, F-F energy, F protons
    for (int j = tid; j < nMI + L.nOinst; j += kBlock) {
        const int a = j < nMI ? s_inst[j] : (s_instO[j - nMI] & 0xffff);
        double acc = p.gt[a];
        if (folded) for (int f = 0; f < p.nF; f++) acc += p.w[(size_t) a * ldw + s_aF[f]];
        if (j < nMI) s_gF[j] = acc; else s_hF[j - nMI] = acc;
    }
    if (warp == kBlock / 32 - 1) {
        double e = 0.0;
        int    np = 0;
        for (int f = lane; f < p.nF; f += 32) {
            const int a = s_aF[f];
            e += p.gt[a];
            np += p.protons[a];
            if (folded) for (int f2 = 0; f2 < f; f2++) e += p.w[(size_t) a * ldw + s_aF[f2]];
        }
        e = pce::warp_sum (e);
        for (int o = 16; o > 0; o >>= 1) np += __shfl_xor_sync (0xffffffffu, np, o);
        if (lane == 0) { s_AF = static_shift + e; s_nF = np; }
    }
    __syncthreads ();

    // O digit of a tile: shift/mask when every O radix is a power of two, else divide
    auto digit = [&] (unsigned x, int o) -> int {      // o uniform across the warp: kernel-parameter tables
        return p.pow2O ? (int) ((x >> p.shiftO[o]) & (unsigned) (p.radOut[o] - 1)) : (int) ((x / p.strideO[o]) % (unsigned) p.radOut[o]);
    };
    auto digit_s = [&] (unsigned x, int o) -> int {    // o differs per lane: shared-memory tables (no constant-bank replays)
        return p.pow2O ? (int) ((x >> s_oShift[o]) & (s_oRad[o] - 1u)) : (int) ((x / s_oStride[o]) % s_oRad[o]);
    };
    const int n_lo = kTB * p.P_lo, n_hi = kTB * p.P_hi, n_i = kTB * p.P_I;
    const int n_terms = p.nO + L.nPairs;

    // ---- rounds of kTB tiles ---------------------------------------------------------------------------
    for (long long tile0 = 0; tile0 < p.P_O; tile0 += kTB) {
        // (a) O-site fields on the instances of every M / I site, the per-site minimum, and the normalised Boltzmann
        // factor of every instance: one item per (tile, site).  Table entries are then products of per-site factors.
        for (int e = tid; e < kTB * nS; e += kBlock) {
            const int tb = e / nS, s = e - tb * nS;
            const unsigned x = (unsigned) (tile0 + tb);
            const int j0 = s_off[s], nj = s_off[s + 1] - j0;
            double g[8];                                       // a site has at most 8 instances here (checked on the host)
            double vmin = DBL_MAX;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c < nj) {
                    double acc = s_gF[j0 + c];
                    if (folded) {
                        if (p.wstage) {
                            const double *row = s_wMO + (j0 + c) * nOi;
                            for (int o = 0; o < p.nO; o++) acc += row[s_offO[o] + digit (x, o)];
                        } else {
                            const double *row = p.w + (size_t) s_inst[j0 + c] * ldw;
                            for (int o = 0; o < p.nO; o++) acc += row[p.firstOut[o] + digit (x, o)];
                        }
                    }
                    g[c] = acc;
                    vmin = fmin (vmin, acc);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; c++) if (c < nj) s_eInst[tb * nMI + j0 + c] = exp (-p.beta * (g[c] - vmin));
            s_smin[tb * nS + s] = vmin;
        }
        // (b) O energy (intrinsic + O-F via hF, O-O pairs) and O protons: warp tb takes tile tb, lanes over terms
        {
            const int tb = warp;
            const unsigned x = (unsigned) (tile0 + tb);
            double e = 0.0;
            int    np = 0;
            if (tile0 + tb < p.P_O) {
                for (int t = lane; t < n_terms; t += 32) {
                    if (t < p.nO) {
                        const int k = s_offO[t] + digit_s (x, t);
                        e += s_hF[k];
                        np += s_instO[k] >> 16;
                    } else if (folded) {
                        const int pr = s_pair[t - p.nO], o = pr & 0xff, o2 = pr >> 8;
                        if (p.wstage) e += s_wOO[(s_offO[o] + digit_s (x, o)) * nOi + s_offO[o2] + digit_s (x, o2)];
                        else e += p.w[(size_t) ((int) s_oFirst[o] + digit_s (x, o)) * ldw + (int) s_oFirst[o2] + digit_s (x, o2)];
                    }
                }
            }
            e = pce::warp_sum (e);
            for (int o = 16; o > 0; o >>= 1) np += __shfl_xor_sync (0xffffffffu, np, o);
            if (lane == 0) { s_cOut[tb] = s_AF + e; s_nOut[tb] = s_nF + np; }
        }
        __syncthreads ();

        // (c) tables: every entry is the product of the per-site factors of its configuration (each <= 1; the per-site
        // minima go into the tile constant).  Items are grouped by table (Lo | Hi | I | tile constant).
        for (int e = tid; e < n_lo + n_hi + n_i + kTB; e += kBlock) {
            if (e < n_lo) {
                const int tb = e / p.P_lo, x = e - tb * p.P_lo;
                const double *f = s_eInst + tb * nMI;
                const uint16_t *row = s_idx + x * kMaxM;
                double v = 1.0;
                for (int i = 0; i < p.nLo; i++) v *= f[row[i]];
                s_eLo[e] = v;
            } else if (e < n_lo + n_hi) {
                const int g = e - n_lo, tb = g / p.P_hi, x = g - tb * p.P_hi;
                const double *f = s_eInst + tb * nMI;
                const uint16_t *row = s_idx + (p.P_lo + x) * kMaxM;
                double v = 1.0;
                for (int i = 0; i < p.nM - p.nLo; i++) v *= f[row[i]];
                s_eHi[g] = v;
            } else if (e < n_lo + n_hi + n_i) {
                const int g = e - n_lo - n_hi, tb = g / p.P_I, x = g - tb * p.P_I;
                const double *f = s_eInst + tb * nMI;
                const uint16_t *row = s_idx + (p.P_lo + p.P_hi + x) * kMaxM;
                double v = 1.0;
                for (int i = 0; i < p.nI; i++) v *= f[row[i]];
                s_eI[tb * 16 + x] = v;
            } else {
                const int tb = e - n_lo - n_hi - n_i;
                double sum = s_cOut[tb];
                for (int i = 0; i < nS; i++) sum += s_smin[tb * nS + i];
                s_cExp[tb] = (tile0 + tb < p.P_O) ? exp (-p.beta * (sum - p.gref)) : 0.0;
            }
        }
        __syncthreads ();

        // (f) accumulate this thread's kTB x P_I states into its proton-count bins.  I configurations are sorted by
        // proton count, so states of one bin are summed in registers and written with one read-modify-write.
        if (mine) {
#pragma unroll 2
            for (int tb = 0; tb < kTB; tb++) {
                const double c_tile = s_cExp[tb] * s_eLo[tb * p.P_lo + m_lo] * s_eHi[tb * p.P_hi + m_hi] * eMM;
                double *mybins = bins + (s_nOut[tb] + p.binI[0]) * kBlock + tid;
                const double *eI = s_eI + tb * 16;
                if (CANON >= 0) {
                    constexpr unsigned mask = canon_flush_mask (CANON >= 0 ? CANON : 0);
                    constexpr int kN = 1 << (CANON >= 0 ? CANON : 0);
                    double ev[kN];
                    if (kN >= 2) {
#pragma unroll
                        for (int i = 0; i < kN / 2; i++) { const double2 t = ((const double2 *) eI)[i]; ev[2 * i] = t.x; ev[2 * i + 1] = t.y; }
                    } else {
                        ev[0] = eI[0];
                    }
                    double acc = 0.0;
                    int    k = 0;
#pragma unroll
                    for (int i = 0; i < kN; i++) {
                        acc += ev[i] * xMI[i];
                        if ((mask >> i) & 1u) { mybins[k * kBlock] += c_tile * acc; acc = 0.0; k++; }
                    }
                } else {
                    double acc = 0.0;
#pragma unroll
                    for (int i = 0; i < kIStates; i++) {
                        if (i < p.P_I) {
                            acc += eI[i] * xMI[i];
                            if ((p.flushI >> i) & 1u) { mybins[(p.binI[i] - p.binI[0]) * kBlock] += c_tile * acc; acc = 0.0; }
                        }
                    }
                }
            }
        }
        __syncthreads ();
    }

    // ---- flush -------------------------------------------------------------------------------------------
    double *out = p.T + (size_t) blockIdx.x * p.NbRel * kBlock;
    for (int e = tid; e < p.NbRel * kBlock; e += kBlock) out[e] = bins[e];
}

// out[e] = sum over c < count of in[c * stride + e] for the slice of c owned by blockIdx.y (deterministic order);
// used twice: chunks -> segments -> one table
__global__ void k_reduce_rows (const double *__restrict__ in, long long count, int width, int per_segment, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= width) return;
    const long long c0 = (long long) blockIdx.y * per_segment;
    const long long c1 = c0 + per_segment < count ? c0 + per_segment : count;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    long long c = c0;
    for (; c + 3 < c1; c += 4) {
        a0 += in[(size_t) c * width + e];
        a1 += in[(size_t) (c + 1) * width + e];
        a2 += in[(size_t) (c + 2) * width + e];
        a3 += in[(size_t) (c + 3) * width + e];
    }
    for (; c < c1; c++) a0 += in[(size_t) c * width + e];
    out[(size_t) blockIdx.y * width + e] = (a0 + a1) + (a2 + a3);
}

// Ftab[chunk][n] = sum over threads of T[chunk][n - nM(tid)][tid] (per-thread bins are relative to the protons of the
// thread's M configuration); one warp per (chunk, n), lanes over threads in a fixed order
__global__ void k_reduce_threads (const double *__restrict__ T, long long nchunks, int NbRel, int Nb, const int32_t *__restrict__ nMp,
                                  double *__restrict__ Ftab) {
    const long long wid = ((long long) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nchunks * Nb) return;
    const long long chunk = wid / Nb;
    const int n = (int) (wid - chunk * Nb);
    const double *base = T + (size_t) chunk * NbRel * kBlock;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; k++) {
        const int t = lane + 32 * k, rel = n - nMp[t];
        if (rel >= 0 && rel < NbRel) acc += base[(size_t) rel * kBlock + t];
    }
    acc = pce::warp_sum (acc);
    if (lane == 0) Ftab[wid] = acc;
}

// Marginals.  rows: 0 = Z, 1 + a = instance a.  One block per row to fill; threads over bins.
//   M site instance a: sum over tid with digit(tid) = a of Mtab[n][tid]
//   F site instance a: sum over chunks with digit(chunk) = a of Ftab[chunk][n]
struct MarginalParams {
    int Nb, NbRel, P_M, nM, nF, nO, write_z;
    const int32_t *nMp;      // protons of each thread's M configuration
    long long chunk_begin, nchunks;
    int16_t siteM[kMaxM], radM[kMaxM], coverM[kMaxM];
    int16_t siteF[kMaxOuter], radF[kMaxOuter], coverF[kMaxOuter];
    const int32_t *first;
    const double  *Mtab, *Ftab;
    double        *bins;    // [(1+ninst)][Nb], accumulated (+=)
};

__global__ void __launch_bounds__ (kBlock) k_marginals (const MarginalParams p) {
    // blockIdx.x enumerates: 0 -> Z; then (M site i, digit d) for every M site; then (F site j, digit d).
    // Each block sums, for every bin n, the source entries whose digit matches: entries are spread over the 256
    // threads (entry e -> thread e % 256) and combined in a fixed order, so the result is deterministic.
    __shared__ double s_part[kBlock / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = blockIdx.x;
    int kind = 0, row = 0;              // 0 = Z (all threads of Mtab), 1 = M site, 2 = F site
    unsigned stride = 1, rad = 1, digit = 0;
    if (b == 0) {
        if (!p.write_z) return;
    } else {
        b -= 1;
        bool found = false;
        unsigned st = 1;
        for (int i = 0; i < p.nM && !found; i++) {
            if (b < p.radM[i]) {
                if (!p.coverM[i]) return;
                kind = 1; row = 1 + p.first[p.siteM[i]] + b; stride = st; rad = (unsigned) p.radM[i]; digit = (unsigned) b; found = true;
            } else { b -= p.radM[i]; st *= (unsigned) p.radM[i]; }
        }
        st = 1;
        for (int j = 0; j < p.nF && !found; j++) {
            if (b < p.radF[j]) {
                if (!p.coverF[j]) return;
                kind = 2; row = 1 + p.first[p.siteF[j]] + b; stride = st; rad = (unsigned) p.radF[j]; digit = (unsigned) b; found = true;
            } else { b -= p.radF[j]; st *= (unsigned) p.radF[j]; }
        }
        if (!found) return;
    }
    const long long count = kind == 2 ? p.nchunks : p.P_M;
    const unsigned  base  = kind == 2 ? (unsigned) p.chunk_begin : 0u;
    // one bin per block (blockIdx.y); entries are spread over the 256 threads (entry e -> thread e % 256, up to 64 per
    // thread: 16384 chunks / 256 threads) and combined in a fixed order
    const int n = blockIdx.y;
    double acc = 0.0;
    for (int k = 0; k < 64; k++) {
        const long long e = tid + (long long) kBlock * k;
        if (e < count && (kind == 0 || ((base + (unsigned) e) / stride) % rad == digit)) {
            if (kind == 2) acc += p.Ftab[(size_t) e * p.Nb + n];
            else {
                const int rel = n - p.nMp[e];
                if (rel >= 0 && rel < p.NbRel) acc += p.Mtab[(size_t) rel * kBlock + e];
            }
        }
    }
    acc = pce::warp_sum (acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads ();
    if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < kBlock / 32; w++) total += s_part[w];
        p.bins[(size_t) row * p.Nb + n] += total;
    }
}

__global__ void k_tilt (int ninst, const double *__restrict__ g, const int32_t *__restrict__ protons, double mu_star, double *__restrict__ gt) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < ninst) gt[a] = g[a] - (double) protons[a] * mu_star;
}

// ---------------------------------------------------------------------------------------------------
// host-side planning
// ---------------------------------------------------------------------------------------------------
struct RunPlan {
    std::vector<int> M, I, O, F;       // site ids
    int nLo = 0;
    std::vector<uint8_t> coverM, coverF;   // whether this run provides the marginals of the site
};

struct Plan {
    std::vector<RunPlan> runs;
    int    Nb = 0;
    double ph_star = 7.0, gref = 0.0;
    double nstates = 1.0;
};

int radix (const pce_model *m, int s) { return m->last[s] - m->first[s] + 1; }

// reference energy: a local minimum of the energy at pH* found by coordinate descent from the per-site
// minima of the tilted intrinsic energies.  Only its rough value matters (it centres the exponent range).
double reference_energy (const pce_model *m, int unfolded, double ph_star) {
    const double mu = pce::mu_of (m->temperature, ph_star);
    const size_t n = (size_t) m->ninst;
    const std::vector<double> &g = unfolded ? m->gmodel : m->gintr;
    std::vector<int> a (m->nsites);
    auto tilted = [&] (int k) { return g[k] - m->protons[k] * mu; };
    for (int i = 0; i < m->nsites; i++) {
        int best = m->first[i];
        for (int k = m->first[i]; k <= m->last[i]; k++) if (tilted (k) < tilted (best)) best = k;
        a[i] = best;
    }
    if (!unfolded) {
        for (int sweep = 0; sweep < 20; sweep++) {
            bool changed = false;
            for (int i = 0; i < m->nsites; i++) {
                int best = a[i];
                double ebest = 0.0;
                bool have = false;
                for (int k = m->first[i]; k <= m->last[i]; k++) {
                    double e = tilted (k);
                    for (int j = 0; j < m->nsites; j++) if (j != i) e += m->wsym[(size_t) k * n + a[j]];
                    if (!have || e < ebest) { have = true; ebest = e; best = k; }
                }
                if (best != a[i]) { a[i] = best; changed = true; }
            }
            if (!changed) break;
        }
    }
    double e = 0.0;
    for (int i = 0; i < m->nsites; i++) {
        e += tilted (a[i]);
        if (!unfolded) for (int j = 0; j < i; j++) e += m->wsym[(size_t) a[i] * n + a[j]];
    }
    return e;
}

int max_protons_total (const pce_model *m) {
    int total = 0;
    for (int i = 0; i < m->nsites; i++) {
        int best = 0;
        for (int k = m->first[i]; k <= m->last[i]; k++) best = std::max (best, (int) m->protons[k]);
        total += best;
    }
    return total;
}

int make_plan (const pce_model *m, const double *ph, int nph, int unfolded, Plan *plan) {
    if (m->nsites > kMaxOuter) return fail (PCE_ERR_LIMIT, "too many sites for analytic enumeration");
    for (int k = 0; k < m->ninst; k++) if (m->protons[k] < 0) return fail (PCE_ERR_VALUE, "negative proton count");
    double lo = ph[0], hi = ph[0];
    for (int q = 1; q < nph; q++) { lo = std::min (lo, ph[q]); hi = std::max (hi, ph[q]); }
    plan->ph_star = 0.5 * (lo + hi);
    const int ptotal = max_protons_total (m);
    plan->Nb = ptotal + 1;
    if (0.5 * (hi - lo) * PCE_LN10 * ptotal > 600.0)
        return fail (PCE_ERR_LIMIT, "pH range too wide for one enumeration pass; split the pH grid");
    plan->gref = reference_energy (m, unfolded, plan->ph_star);
    plan->nstates = 1.0;
    for (int i = 0; i < m->nsites; i++) plan->nstates *= radix (m, i);

    std::vector<uint8_t> covered (m->nsites, 0);
    int ncovered = 0;
    plan->runs.clear ();
    while (ncovered < m->nsites || plan->runs.empty ()) {
        RunPlan run;
        std::vector<uint8_t> used (m->nsites, 0);
        // M: uncovered sites first, then any, while the product of radices fits in a CTA
        long long pm = 1;
        for (int pass = 0; pass < 2; pass++)
            for (int s = 0; s < m->nsites && (int) run.M.size () < kMaxM; s++) {
                if (used[s] || (pass == 0 && covered[s])) continue;
                if (pm * radix (m, s) <= kBlock) { run.M.push_back (s); used[s] = 1; pm *= radix (m, s); }
            }
        // Lo / Hi split: balance the two table sizes
        {
            long long plo = 1;
            run.nLo = 0;
            for (size_t i = 0; i < run.M.size (); i++) {
                if (plo * plo >= pm) break;
                plo *= radix (m, run.M[i]); run.nLo = (int) i + 1;
            }
        }
        // I: canonical sites (binary, instances one proton apart) so that the kernel's compile-time bin structure applies;
        // covered ones preferred (their marginals are not needed from this run).  Fallback: any sites (generic kernel).
        long long pi = 1;
        auto canonical = [&] (int s) { return radix (m, s) == 2 && std::abs (m->protons[m->first[s]] - m->protons[m->first[s] + 1]) == 1; };
        for (int pass = 0; pass < 2; pass++)
            for (int s = 0; s < m->nsites && (int) run.I.size () < kMaxI; s++) {
                if (used[s] || !canonical (s) || (pass == 0 && !covered[s])) continue;
                if (pi * 2 <= kIStates) { run.I.push_back (s); used[s] = 1; pi *= 2; }
            }
        if (run.I.empty ())
            for (int pass = 0; pass < 2; pass++)
                for (int s = 0; s < m->nsites && (int) run.I.size () < kMaxI; s++) {
                    if (used[s] || (pass == 0 && !covered[s])) continue;
                    if (pi * radix (m, s) <= kIStates) { run.I.push_back (s); used[s] = 1; pi *= radix (m, s); }
                }
        // outer sites: covered ones become the low (O) digits, uncovered ones the high (F) digits
        std::vector<int> outer;
        for (int s = 0; s < m->nsites; s++) if (!used[s] && covered[s]) outer.push_back (s);
        for (int s = 0; s < m->nsites; s++) if (!used[s] && !covered[s]) outer.push_back (s);
        // F = as many high digits as keep the chunk count within kMaxChunks
        long long pf = 1;
        int nF = 0;
        for (int j = (int) outer.size () - 1; j >= 0; j--) {
            if (pf * radix (m, outer[j]) > kMaxChunks) break;
            pf *= radix (m, outer[j]); nF++;
        }
        if ((int) outer.size () - nF > kMaxO) return fail (PCE_ERR_LIMIT, "too many sites for analytic enumeration");
        const int nO = (int) outer.size () - nF;
        run.O.assign (outer.begin (), outer.begin () + nO);
        run.F.assign (outer.begin () + nO, outer.end ());
        run.coverM.assign (run.M.size (), 0);
        run.coverF.assign (run.F.size (), 0);
        int newly = 0;
        for (size_t i = 0; i < run.M.size (); i++) if (!covered[run.M[i]]) { run.coverM[i] = 1; covered[run.M[i]] = 1; newly++; }
        for (size_t j = 0; j < run.F.size (); j++) if (!covered[run.F[j]]) { run.coverF[j] = 1; covered[run.F[j]] = 1; newly++; }
        ncovered += newly;
        plan->runs.push_back (run);
        if (newly == 0 && ncovered < m->nsites) return fail (PCE_ERR_LIMIT, "cannot cover all sites (too many states)");
        if (plan->runs.size () > 16) return fail (PCE_ERR_LIMIT, "too many enumeration passes");
    }
    return PCE_OK;
}

// One enumeration per planned run (roles rotated so that every site is covered), accumulated into d_bins.
int run_plan (pce_model *m, const Plan &plan, int unfolded, int shard, int nshards, double *d_bins) {
    const size_t nbins_total = (size_t) (1 + m->ninst) * plan.Nb;
    PCE_CUDA (cudaMemsetAsync (d_bins, 0, nbins_total * sizeof (double), m->stream));
    // scratch lives in the model (grow-only): repeated curves do not pay cudaMalloc/cudaFree of ~1 GB
    pce::DeviceBuffer<double> &d_gt = m->s_gt, &d_T = m->s_T, &d_Mtab = m->s_Mtab, &d_Ftab = m->s_Ftab, &d_seg = m->s_seg, &d_xs = m->s_xs;
    pce::DeviceBuffer<int32_t> &d_ms = m->s_ms;
    PCE_CUDA (d_gt.resize (m->ninst));
    const double mu_star = pce::mu_of (m->temperature, plan.ph_star);
    k_tilt<<<(m->ninst + 127) / 128, 128, 0, m->stream>>> (m->ninst, unfolded ? m->d_gmodel.ptr : m->d_gintr.ptr, m->d_protons.ptr, mu_star, d_gt.ptr);
    pce::count_launch ();
    int status = PCE_OK;
    for (size_t r = 0; r < plan.runs.size () && status == PCE_OK; r++) {
        const RunPlan &run = plan.runs[r];
        EnumParams p;
        std::memset (&p, 0, sizeof (p));
        p.nsites = m->nsites; p.ninst = m->ninst; p.ldw = m->ldw; p.unfolded = unfolded;
        p.gt = d_gt.ptr; p.w = m->d_w.ptr; p.protons = m->d_protons.ptr; p.first = m->d_first.ptr;
        p.nM = (int) run.M.size (); p.nLo = run.nLo; p.nI = (int) run.I.size (); p.nO = (int) run.O.size (); p.nF = (int) run.F.size ();
        p.P_lo = p.P_hi = p.P_I = 1;
        int nMI = 0;
        for (int i = 0; i < p.nM; i++) {
            p.siteM[i] = (int16_t) run.M[i]; p.radM[i] = (int16_t) radix (m, run.M[i]); nMI += p.radM[i];
            if (i < p.nLo) p.P_lo *= p.radM[i]; else p.P_hi *= p.radM[i];
        }
        p.P_M = p.P_lo * p.P_hi;
        if (p.P_hi > 64 || p.P_M > kBlock) return fail (PCE_ERR_LIMIT, "internal: thread-level site split out of range");
        for (int i = 0; i < p.nI; i++) { p.siteI[i] = (int16_t) run.I[i]; p.radI[i] = (int16_t) radix (m, run.I[i]); p.P_I *= p.radI[i]; nMI += p.radI[i]; }
        if (nMI > kMaxMI) return fail (PCE_ERR_LIMIT, "too many instances in the thread-level sites");
        p.P_O = 1;
        long long pf = 1;
        for (int j = 0; j < p.nO; j++) { p.siteOut[j] = (int16_t) run.O[j]; p.radOut[j] = (int16_t) radix (m, run.O[j]); p.P_O *= p.radOut[j]; }
        for (int j = 0; j < p.nF; j++) { p.siteOut[p.nO + j] = (int16_t) run.F[j]; p.radOut[p.nO + j] = (int16_t) radix (m, run.F[j]); pf *= p.radOut[p.nO + j]; }
        {
            unsigned stride = 1, shift = 0;
            p.pow2O = 1;
            for (int j = 0; j < p.nO; j++) {
                p.strideO[j] = stride; stride *= (unsigned) p.radOut[j];
                const int r = p.radOut[j];
                if (r & (r - 1)) p.pow2O = 0;
                p.shiftO[j] = (uint8_t) shift;
                int bits = 0;
                while ((1 << bits) < r) bits++;
                shift += (unsigned) bits;
            }
            p.nMI = nMI;
            p.nOinst = 0;
            for (int j = 0; j < p.nO; j++) p.nOinst += p.radOut[j];
            if (p.nO > kMaxO) return fail (PCE_ERR_LIMIT, "too many sites for analytic enumeration");
            p.canon = p.nI;
            for (int i = 0; i < p.nI; i++) {
                const int f = m->first[p.siteI[i]];
                if (p.radI[i] != 2 || std::abs (m->protons[f] - m->protons[f + 1]) != 1) p.canon = -1;
            }
            for (int j = 0; j < p.nO + p.nF; j++) p.firstOut[j] = (int16_t) m->first[p.siteOut[j]];
            // I configurations sorted by proton count (stable)
            std::vector<std::pair<int, int>> order;
            for (int c = 0; c < p.P_I; c++) {
                int x = c, np = 0;
                for (int i = 0; i < p.nI; i++) { np += m->protons[m->first[p.siteI[i]] + x % p.radI[i]]; x /= p.radI[i]; }
                order.push_back (std::make_pair (np, c));
            }
            std::stable_sort (order.begin (), order.end ());
            p.flushI = 0;
            for (int i = 0; i < p.P_I; i++) {
                p.permI[i] = (int16_t) order[i].second; p.binI[i] = (int16_t) order[i].first;
                if (i + 1 == p.P_I || order[i + 1].first != order[i].first) p.flushI |= 1u << i;
            }
        }
        if (p.P_O > 2147483647LL) return fail (PCE_ERR_LIMIT, "too many tiles per chunk");
        p.Nb = plan.Nb; p.beta = 1.0 / (PCE_R * m->temperature); p.gref = plan.gref;
        {
            int maxM = 0;
            for (int i = 0; i < p.nM; i++) {
                int best = 0;
                for (int k = m->first[p.siteM[i]]; k <= m->last[p.siteM[i]]; k++) best = std::max (best, (int) m->protons[k]);
                maxM += best;
                if (p.radM[i] > 8) return fail (PCE_ERR_LIMIT, "sites with more than 8 instances are not supported by the enumeration kernel");
            }
            for (int i = 0; i < p.nI; i++) if (p.radI[i] > 8) return fail (PCE_ERR_LIMIT, "sites with more than 8 instances are not supported by the enumeration kernel");
            p.NbRel = plan.Nb - maxM;
            p.wstage = ((size_t) nMI * p.nOinst + (size_t) p.nOinst * p.nOinst) * 8 <= 16 * 1024 ? 1 : 0;
        }
        // this shard's slice of the chunk range
        const long long c0 = pf * shard / nshards, c1 = pf * (shard + 1) / nshards;
        p.chunk_begin = c0; p.nchunks = c1 - c0;
        if (p.nchunks <= 0) continue;
        const EnumLayout layout (p.NbRel, nMI, p.P_lo, p.P_hi, p.P_I, p.nF, p.nM, p.nI, p.nO, p.nOinst, p.wstage);
        const size_t smem = layout.total;
        size_t limit = 0;
        {
            int value = 0;
            PCE_CUDA (cudaDeviceGetAttribute (&value, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
            limit = (size_t) value;
        }
        if (smem + 2 * 1024 > limit) return fail (PCE_ERR_LIMIT, "too many proton-count bins for shared memory");
        const int width = p.NbRel * kBlock;
        const int per_segment = 64;
        const int nseg = (int) ((p.nchunks + per_segment - 1) / per_segment);
        PCE_CUDA (d_T.resize ((size_t) p.nchunks * width));
        PCE_CUDA (d_seg.resize ((size_t) nseg * width));
        PCE_CUDA (d_Mtab.resize ((size_t) width));
        PCE_CUDA (d_Ftab.resize ((size_t) p.nchunks * plan.Nb));
        PCE_CUDA (d_xs.resize ((size_t) (1 + kIStates) * kBlock + 1));
        PCE_CUDA (d_ms.resize ((size_t) 3 * kBlock));
        p.T = d_T.ptr; p.xs = d_xs.ptr; p.ms = d_ms.ptr; p.shift = d_xs.ptr + (size_t) (1 + kIStates) * kBlock;
        k_enum_static<<<1, kBlock, 0, m->stream>>> (p);
        PCE_CUDA (cudaGetLastError ());
        {
            void (*kernel) (const EnumParams) = k_enum<-1>;
            switch (p.canon) {
                case 0: kernel = k_enum<0>; break;
                case 1: kernel = k_enum<1>; break;
                case 2: kernel = k_enum<2>; break;
                case 3: kernel = k_enum<3>; break;
                case 4: kernel = k_enum<4>; break;
                default: break;
            }
            PCE_CUDA (cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
            kernel<<<(unsigned) p.nchunks, kBlock, smem, m->stream>>> (p);
        }
        PCE_CUDA (cudaGetLastError ());
        k_reduce_rows<<<dim3 ((width + 127) / 128, nseg), 128, 0, m->stream>>> (d_T.ptr, p.nchunks, width, per_segment, d_seg.ptr);
        PCE_CUDA (cudaGetLastError ());
        k_reduce_rows<<<dim3 ((width + 127) / 128, 1), 128, 0, m->stream>>> (d_seg.ptr, nseg, width, nseg, d_Mtab.ptr);
        PCE_CUDA (cudaGetLastError ());
        {
            const long long warps = p.nchunks * plan.Nb;
            const long long blocks = (warps * 32 + 255) / 256;
            k_reduce_threads<<<(unsigned) blocks, 256, 0, m->stream>>> (d_T.ptr, p.nchunks, p.NbRel, plan.Nb, d_ms.ptr + 2 * kBlock, d_Ftab.ptr);
            PCE_CUDA (cudaGetLastError ());
        }
        MarginalParams mp;
        std::memset (&mp, 0, sizeof (mp));
        mp.Nb = plan.Nb; mp.NbRel = p.NbRel; mp.nMp = d_ms.ptr + 2 * kBlock; mp.P_M = p.P_M; mp.nM = p.nM; mp.nF = p.nF; mp.nO = p.nO; mp.write_z = r == 0 ? 1 : 0;
        mp.chunk_begin = c0; mp.nchunks = p.nchunks;
        int rows = 1;
        for (int i = 0; i < p.nM; i++) { mp.siteM[i] = p.siteM[i]; mp.radM[i] = p.radM[i]; mp.coverM[i] = run.coverM[i]; rows += p.radM[i]; }
        for (int j = 0; j < p.nF; j++) { mp.siteF[j] = p.siteOut[p.nO + j]; mp.radF[j] = p.radOut[p.nO + j]; mp.coverF[j] = run.coverF[j]; rows += mp.radF[j]; }
        mp.first = m->d_first.ptr; mp.Mtab = d_Mtab.ptr; mp.Ftab = d_Ftab.ptr; mp.bins = d_bins;
        k_marginals<<<dim3 (rows, plan.Nb), kBlock, 0, m->stream>>> (mp);
        PCE_CUDA (cudaGetLastError ());
        pce::count_launch (6);
    }
    cudaError_t err = cudaStreamSynchronize (m->stream);
    if (err != cudaSuccess) return fail (PCE_ERR_CUDA, std::string ("analytic enumeration: ") + cudaGetErrorString (err));
    return status;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" long long pce_analytic_bins_size (const pce_model *m) {
    if (m == nullptr) return -1;
    return (long long) (1 + m->ninst) * (max_protons_total (m) + 1);
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
    Plan plan;
    status = make_plan (m, ph, nph, unfolded, &plan);
    if (status != PCE_OK) return status;
    return run_plan (m, plan, unfolded, shard, nshards, d_bins);
}

extern "C" int pce_analytic_finish (pce_model *m, const double *bins, const double *ph, int nph, int unfolded, double gzero,
                                    double *out_p, double *out_lnz) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    Plan plan;
    int status = make_plan (m, ph, nph, unfolded, &plan);
    if (status != PCE_OK) return status;
    const int Nb = plan.Nb;
    const long double beta = 1.0L / ((long double) PCE_R * m->temperature);
    const long double mu_star = pce::mu_of (m->temperature, plan.ph_star);
    std::vector<long double> weight (Nb);
    for (int q = 0; q < nph; q++) {
        // weight of bin n relative to pH*: exp(-beta * n * (mu* - mu)), shifted by the largest log-term of Z
        const long double mu = pce::mu_of (m->temperature, ph[q]);
        long double shift = -1.0e300L;
        for (int n = 0; n < Nb; n++) {
            if (bins[n] > 0.0) {
                const long double t = logl ((long double) bins[n]) - beta * n * (mu_star - mu);
                if (t > shift) shift = t;
            }
        }
        long double z = 0.0L;
        for (int n = 0; n < Nb; n++) {
            weight[n] = expl (-beta * n * (mu_star - mu) - shift);
            z += (long double) bins[n] * weight[n];
        }
        if (!(z > 0.0L)) return fail (PCE_ERR_STATE, "partition function underflowed; split the pH grid");
        for (int a = 0; a < m->ninst; a++) {
            const double *row = bins + (size_t) (1 + a) * Nb;
            long double s = 0.0L;
            for (int n = 0; n < Nb; n++) s += (long double) row[n] * weight[n];
            const double value = (double) (s / z);
            if (out_p) out_p[(size_t) q * m->ninst + a] = value;
            if (q == nph - 1) m->prob[a] = value;
        }
        if (out_lnz) out_lnz[q] = (double) (logl (z) + shift - beta * ((long double) plan.gref - gzero));
    }
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
