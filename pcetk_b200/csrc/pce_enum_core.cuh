// pce_enum_core.cuh -- kernel 2: the CTA programs of the factorised exhaustive enumeration.
//
// Replaces EnergyModel_CalculateZ / _CalculateProbabilitiesFromZ (extensions/csource/EnergyModel.c:252-281, 318-337).
// Roles (M, I, O1, O2, F) and bin scales are planned on the host (pce_enum_plan.h).  With
//
//   E(state) = E_F + sum_o h_F[b_o] + E_OO + sum_inner-sites ( f_F[a] + sum_o W[a][b_o] ) + E_MM + E_MI + E_II
//
// the Boltzmann factor of a state is a PRODUCT of factors that each depend on few roles, every one normalised to <= 1
// by a minimum that moves into the tile constant:
//
//   b = Qc[t][i] * xs[m][i] * P1[m][t1] * ( A[m] * P2lo[t2][m_lo] * P2hi[t2][m_hi] )
//
//   xs[m][i]   = exp(-beta (E_MM + E_MI + E_II - min))                        per enumeration, in REGISTERS (16 doubles)
//   P1[m][t1]  = exp(-beta sum_{M sites, O1 sites} (W - wmin))                per enumeration, in registers (8 doubles)
//   P2lo/hi    = the same for the O2 sites, factorised over the Lo / Hi halves of the M configuration (shared memory)
//   A[m]       = product over the M sites of exp(-beta (f_F[a] - site minimum)) per chunk
//   Qc[t][i]   = Q[t][i] * EIF[i] * cE[t]: I-site factors times the tile constant, which carries every minimum, the
//                reference energy and the bin scale exp(sigma)                  per chunk, shared memory, broadcast reads
//
// so the inner loop is ONE DFMA per state on register operands plus one broadcast 128-bit shared-memory load per two
// states; proton-count bins of the I x O1 configurations are compile-time register accumulators (canonical sites), flushed
// to shared memory when the O2 proton count changes.  There is no per-tile table building and no barrier inside a chunk's
// accumulation (round 1: three barriers per 8 tiles, fp64 pipe 17 % busy).
//
// Every function below is written against a thread index and plain pointers: the CUDA kernels (pce_analytic.cu) call them
// with __syncthreads() between the phases, the CPU emulator of the tests (tests/emu/enum_emu.cpp) runs each phase for
// tid = 0..255 in a loop.  No function here synchronises.
#pragma once

#include <cfloat>
#include <cmath>

#include "pce_enum_plan.h"

#if defined(__CUDACC__) && !defined(PCE_ENUM_EMULATE)
#define PCE_DEV __device__ __forceinline__
#define PCE_HD __host__ __device__ __forceinline__
#define PCE_UNROLL _Pragma ("unroll")
#else
#define PCE_DEV inline
#define PCE_HD inline
#define PCE_UNROLL
#endif

namespace pce_enum {

struct EnumParams {
    // model
    int ninst, ldw, unfolded;
    const double *gt, *w;        // tilted intrinsic energies G[a] - protons[a] mu(pH*); W [ninst][ldw]
    double beta, gref;
    // sizes (RunTables)
    int nM, nLo, nI, nJ, nO1, nO2, nO, nF, nS;
    int P_lo, P_hi, P_M, P_I, P_J, P_O1, P_O2, P_O;
    int nInner, nMinst, nOinst;
    int NbC, NbCp, NbP, NbAll, nU, nV, mode, needF;
    unsigned flushI;
    long long chunk_begin, nchunks_local, ldF;
    // integer tables
    const int16_t *innerInst, *innerSite, *innerOff, *oInst, *idxM, *idxLo, *idxHi, *idxI, *idxJ, *mlo, *mhi, *fInst;
    const uint8_t *nrelM, *nrelI, *nrelO1, *nrelO2, *bO, *fRel;
    const int32_t *fOff, *fRad, *fShift;
    const long long *fStride;
    const double *D, *sigmaU;
    // numeric tables written by the static program
    double *xs;                  // [kXStates][kBlock] static factor of (thread's M configuration, I x J configuration c = j * 16 + i)
    double *P1;                  // [kO1States][kBlock]
    double *P2lo, *P2hi;         // [P_O2][P_lo], [P_O2][P_hi]
    double *Q;                   // [P_O][kIStates]
    double *QJ;                  // [P_O][2] the same for the two instances of the J site
    double *EO;                  // [P_O] O-O energy of the tile plus the minima taken out of the inner-O factors
    double *shiftS;              // [1] minimum taken out of xs
    // outputs of the enumeration program
    double *T;                   // [grid][NbP][kBlock] per-thread bins of every CTA
    double *FtabT;               // [NbAll][ldF] per-chunk bins (zeroed before the launch)
};

PCE_HD size_t align16 (size_t o) { return (o + 15) & ~(size_t) 15; }

// ---------------------------------------------------------------------------------------------------
// static program: one CTA per enumeration builds the numeric tables
// ---------------------------------------------------------------------------------------------------
struct StaticLayout {
    size_t off_wmin, off_vo1, off_vo2, off_vmi, off_eii, off_red, off_en, off_inst, off_sw, total;
    int    staged;               // W sub-blocks (inner x inner, inner x O, O x O) staged in shared memory
    PCE_HD StaticLayout (const EnumParams &p) {
        size_t o = 0;
        off_wmin = o; o += (size_t) (p.nS > 0 ? p.nS : 1) * (p.nOinst > 0 ? p.nOinst : 1) * 8;
        off_vo1 = o;  o += (size_t) (p.nInner > 0 ? p.nInner : 1) * kO1States * 8;
        off_vo2 = o;  o += (size_t) (p.nInner > 0 ? p.nInner : 1) * kO2States * 8;
        off_vmi = o;  o += (size_t) (p.nMinst > 0 ? p.nMinst : 1) * kXStates * 8;
        off_eii = o;  o += (size_t) kXStates * 8;
        off_red = o;  o += (size_t) kBlock * 8;
        off_en = o;   o += (size_t) kXStates * kBlock * 8;
        off_inst = o; o += align16 ((size_t) (p.nInner + p.nOinst + 2) * 4);
        off_sw = o;
        const size_t nw = (size_t) p.nInner * p.nInner + (size_t) p.nInner * p.nOinst + (size_t) p.nOinst * p.nOinst;
        staged = nw * 8 <= 96 * 1024 ? 1 : 0;
        if (staged) o += nw * 8;
        total = align16 (o);
    }
};

// One CTA per enumeration.  Every lookup of the phases below goes to shared memory: the instance lists and (when they fit)
// the W sub-blocks between the inner (M, I) and the O instances are staged first -- a chain of three dependent global
// loads per W entry made the first version of this program take 70 us.
struct StaticCta {
    const EnumParams &p;
    double *wmin, *vo1, *vo2, *vmi, *eii, *red, *en, *sw;
    int32_t *inner, *oinst;
    int staged;
    PCE_DEV StaticCta (const EnumParams &params, unsigned char *smem) : p (params) {
        const StaticLayout L (params);
        wmin = (double *) (smem + L.off_wmin); vo1 = (double *) (smem + L.off_vo1); vo2 = (double *) (smem + L.off_vo2);
        vmi = (double *) (smem + L.off_vmi); eii = (double *) (smem + L.off_eii); red = (double *) (smem + L.off_red);
        en = (double *) (smem + L.off_en); inner = (int32_t *) (smem + L.off_inst); oinst = inner + params.nInner;
        sw = (double *) (smem + L.off_sw); staged = L.staged;
    }
    PCE_DEV double W (int a, int b) const { return p.w[(size_t) a * p.ldw + b]; }
    // W between compact indices: inner j x inner j2, inner j x O instance o, O x O
    PCE_DEV double Wii (int j, int j2) const { return staged ? sw[j * p.nInner + j2] : W (inner[j], inner[j2]); }
    PCE_DEV double Wio (int j, int o) const { return staged ? sw[p.nInner * p.nInner + j * p.nOinst + o] : W (inner[j], oinst[o]); }
    PCE_DEV double Woo (int o, int o2) const { return staged ? sw[p.nInner * (p.nInner + p.nOinst) + o * p.nOinst + o2] : W (oinst[o], oinst[o2]); }

    // S0a: instance lists; S0b: W sub-blocks
    PCE_DEV void phase0a (int tid) const {
        for (int j = tid; j < p.nInner; j += kBlock) inner[j] = p.innerInst[j];
        for (int o = tid; o < p.nOinst; o += kBlock) oinst[o] = p.oInst[o];
    }
    PCE_DEV void phase0b (int tid) const {
        if (!staged || p.unfolded) return;
        const int n1 = p.nInner * p.nInner, n2 = n1 + p.nInner * p.nOinst, n3 = n2 + p.nOinst * p.nOinst;
        for (int e = tid; e < n3; e += kBlock) {
            int a, b;
            if (e < n1) { a = inner[e / p.nInner]; b = inner[e % p.nInner]; }
            else if (e < n2) { a = inner[(e - n1) / p.nOinst]; b = oinst[(e - n1) % p.nOinst]; }
            else { a = oinst[(e - n2) / p.nOinst]; b = oinst[(e - n2) % p.nOinst]; }
            sw[e] = W (a, b);
        }
    }
    // S1: wmin[s][o] = min over the instances j of inner site s of W[j][O instance o]
    PCE_DEV void phase1 (int tid) const {
        for (int e = tid; e < p.nS * p.nOinst; e += kBlock) {
            const int s = e / p.nOinst, o = e - s * p.nOinst;
            double v = 0.0;
            if (!p.unfolded) {
                v = DBL_MAX;
                for (int k = p.innerOff[s]; k < p.innerOff[s + 1]; k++) v = fmin (v, Wio (k, o));
            }
            wmin[e] = v;
        }
    }
    // S2: per-instance partial sums
    PCE_DEV void phase2 (int tid) const {
        const bool folded = !p.unfolded;
        for (int e = tid; e < p.nInner * kO1States; e += kBlock) {          // VO1[j][t1]
            const int j = e / kO1States, t1 = e - j * kO1States, s = p.innerSite[j];
            double v = 0.0;
            if (folded && t1 < p.P_O1)
                for (int o = 0; o < p.nO1; o++) { const int b = p.bO[(size_t) t1 * kMaxO + o]; v += Wio (j, b) - wmin[s * p.nOinst + b]; }
            vo1[e] = v;
        }
        for (int e = tid; e < p.nInner * kO2States; e += kBlock) {          // VO2[j][t2]
            const int j = e / kO2States, t2 = e - j * kO2States, s = p.innerSite[j];
            double v = 0.0;
            if (folded && t2 < p.P_O2)
                for (int o = 0; o < p.nO2; o++) { const int b = p.bO[(size_t) t2 * p.P_O1 * kMaxO + p.nO1 + o]; v += Wio (j, b) - wmin[s * p.nOinst + b]; }
            vo2[e] = v;
        }
        const int nIJ = p.nI + p.nJ;                                        // sites of an I x J configuration: idxI[c][0 .. nI) and, with a J site, idxI[c][kMaxI]
        for (int e = tid; e < p.nMinst * kXStates; e += kBlock) {           // Vmi[j][c] = sum over the I and J sites of W[j][instance of configuration c]
            const int j = e / kXStates, c = e - j * kXStates;
            double v = 0.0;
            if (folded && (c & (kIStates - 1)) < p.P_I && c / kIStates < p.P_J) {
                for (int l = 0; l < p.nI; l++) v += Wii (j, p.idxI[c * (kMaxI + 1) + l]);
                if (p.nJ) v += Wii (j, p.idxI[c * (kMaxI + 1) + kMaxI]);
            }
            vmi[e] = v;
        }
        for (int c = tid; c < kXStates; c += kBlock) {
            double v = 0.0;
            if (folded && (c & (kIStates - 1)) < p.P_I && c / kIStates < p.P_J) {
                int site[kMaxI + 1];
                for (int l = 0; l < p.nI; l++) site[l] = p.idxI[c * (kMaxI + 1) + l];
                if (p.nJ) site[p.nI] = p.idxI[c * (kMaxI + 1) + kMaxI];
                for (int l = 0; l < nIJ; l++) for (int l2 = 0; l2 < l; l2++) v += Wii (site[l], site[l2]);
            }
            eii[c] = v;
        }
    }
    // S3: static energies of thread tid, en[c][tid] = E_MM + E_M(IJ)(c) + E_(IJ)(IJ)(c), and the thread's minimum
    PCE_DEV void phase3 (int tid) const {
        int idx[kMaxM];
        for (int k = 0; k < kMaxM; k++) idx[k] = p.idxM[tid * kMaxM + k];
        double emm = 0.0, vmin = DBL_MAX;
        if (!p.unfolded && tid < p.P_M)
            for (int k = 0; k < p.nM; k++) for (int k2 = 0; k2 < k; k2++) emm += Wii (idx[k], idx[k2]);
        for (int c = 0; c < kXStates; c++) {
            double v = emm + eii[c];
            if (tid < p.P_M && (c & (kIStates - 1)) < p.P_I && c / kIStates < p.P_J) {
                for (int k = 0; k < p.nM; k++) v += vmi[idx[k] * kXStates + c];
                vmin = fmin (vmin, v);
            }
            en[c * kBlock + tid] = v;
        }
        red[tid] = vmin;
    }
    PCE_DEV void phase3b (int tid) const {        // serial minimum (256 values; once per enumeration)
        if (tid == 0) {
            double v = red[0];
            for (int k = 1; k < kBlock; k++) v = fmin (v, red[k]);
            *p.shiftS = v;
            red[0] = v;
        }
    }
    // S4: the tables
    PCE_DEV void phase4 (int tid) const {
        const double shift = red[0];
        int idx[kMaxM];
        for (int k = 0; k < kMaxM; k++) idx[k] = p.idxM[tid * kMaxM + k];
        for (int c = 0; c < kXStates; c++)
            p.xs[c * kBlock + tid] = (tid < p.P_M && (c & (kIStates - 1)) < p.P_I && c / kIStates < p.P_J) ? exp (-p.beta * (en[c * kBlock + tid] - shift)) : 0.0;
        for (int t1 = 0; t1 < kO1States; t1++) {
            double v = 0.0;
            if (tid < p.P_M && t1 < p.P_O1) {
                double s = 0.0;
                for (int k = 0; k < p.nM; k++) s += vo1[idx[k] * kO1States + t1];
                v = exp (-p.beta * s);
            }
            p.P1[t1 * kBlock + tid] = v;
        }
        for (int e2 = tid; e2 < p.P_O2 * p.P_lo; e2 += kBlock) {
            const int t2 = e2 / p.P_lo, x = e2 - t2 * p.P_lo;
            double s = 0.0;
            for (int k = 0; k < p.nLo; k++) s += vo2[p.idxLo[x * kMaxM + k] * kO2States + t2];
            p.P2lo[e2] = exp (-p.beta * s);
        }
        for (int e2 = tid; e2 < p.P_O2 * p.P_hi; e2 += kBlock) {
            const int t2 = e2 / p.P_hi, x = e2 - t2 * p.P_hi;
            double s = 0.0;
            for (int k = 0; k < p.nM - p.nLo; k++) s += vo2[p.idxHi[x * kMaxM + k] * kO2States + t2];
            p.P2hi[e2] = exp (-p.beta * s);
        }
        for (int e2 = tid; e2 < p.P_O * kIStates; e2 += kBlock) {
            const int t = e2 / kIStates, i = e2 - t * kIStates, t2 = t / p.P_O1, t1 = t - t2 * p.P_O1;
            double v = 0.0;
            if (i < p.P_I) {
                double s = 0.0;
                for (int l = 0; l < p.nI; l++) { const int j = p.idxI[i * (kMaxI + 1) + l]; s += vo1[j * kO1States + t1] + vo2[j * kO2States + t2]; }
                v = exp (-p.beta * s);
            }
            p.Q[e2] = v;
        }
        for (int e2 = tid; e2 < p.P_O * 2; e2 += kBlock) {                 // QJ[t][j]: the J instance against the tile's O configuration
            const int t = e2 >> 1, j = e2 & 1, t2 = t / p.P_O1, t1 = t - t2 * p.P_O1;
            double v = j == 0 ? 1.0 : 0.0;
            if (p.nJ) { const int a = p.idxJ[j]; v = exp (-p.beta * (vo1[a * kO1States + t1] + vo2[a * kO2States + t2])); }
            p.QJ[e2] = v;
        }
        for (int t = tid; t < p.P_O; t += kBlock) {
            const uint8_t *row = p.bO + (size_t) t * kMaxO;
            double v = 0.0;
            for (int o = 0; o < p.nO; o++) {
                double wsum = 0.0;
                for (int s = 0; s < p.nS; s++) wsum += wmin[s * p.nOinst + row[o]];       // the minima taken out of the inner-O factors
                v += wsum;
                if (!p.unfolded) for (int o2 = 0; o2 < o; o2++) v += Woo (row[o], row[o2]);
            }
            p.EO[t] = v;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// enumeration program: persistent CTA, chunks blockIdx.x, blockIdx.x + gridDim.x, ...
// ---------------------------------------------------------------------------------------------------
struct EnumLayout {
    size_t off_pb, off_cb, off_qc, off_qj, off_p2lo, off_p2hi, off_ce, off_ff, off_ef, off_minf, off_eif, off_elo, off_ehi, off_d, off_su, off_eo,
           off_af, off_rf, off_misc, off_bo, off_nrel, off_idx, off_inner, total;
    PCE_HD EnumLayout (const EnumParams &p) {
        size_t o = 0;
        off_pb = o;   o += (size_t) p.NbP * kBlock * 8;
        off_cb = o;   o += (size_t) p.NbC * kBlock * 8;
        off_qc = o;   o += (size_t) p.P_O * kIStates * 8;
        off_qj = o;   o += (size_t) p.P_O * 2 * 8;
        off_p2lo = o; o += (size_t) p.P_O2 * p.P_lo * 8;
        off_p2hi = o; o += (size_t) p.P_O2 * p.P_hi * 8;
        off_ce = o;   o += (size_t) p.P_O * 8;
        off_ff = o;   o += (size_t) (p.nInner + p.nOinst + p.nF + 1) * 8;
        off_ef = o;   o += (size_t) (p.nInner + 1) * 8;
        off_minf = o; o += (size_t) (p.nS + 1) * 8;
        off_eif = o;  o += (size_t) (kIStates + 2) * 8;        // EIF[16] | EJF[2]
        off_elo = o;  o += (size_t) p.P_lo * 8;
        off_ehi = o;  o += (size_t) p.P_hi * 8;
        off_d = o;    o += (size_t) p.nU * p.nV * 8;
        off_su = o;   o += (size_t) p.nU * 8;
        off_eo = o;   o += (size_t) p.P_O * 8;
        off_af = o;   o += (size_t) (p.nF + 1) * 4;
        off_rf = o;   o += (size_t) (p.nF + 1) * 4;
        off_misc = o; o += 16;
        off_bo = o;   o += align16 ((size_t) p.P_O * kMaxO);
        off_nrel = o; o += (size_t) (kIStates + kO1States + kO2States + kBlock);        // nrelI | nrelO1 | nrelO2 | nrelM
        o = align16 (o);
        off_idx = o;  o += (size_t) (kIStates * (kMaxI + 1) + (p.P_lo + p.P_hi) * kMaxM) * 2;      // idxI (j = 0 rows) | idxLo | idxHi
        o = align16 (o);
        off_inner = o; o += (size_t) (2 * p.nInner + p.nS + 1 + p.nOinst + 4) * 2;           // innerInst | innerSite | innerOff | oInst
        total = align16 (o);
    }
};

struct EnumThread {              // per-thread registers
    double xs[kXStates];         // static factors of the thread's M configuration, c = j * 16 + i (zero beyond P_I / P_J)
    double p1[kO1States];
    double A;                    // chunk factor of the M configuration
    int    mlo, mhi, nrelM;
};

// MODE 0: run-time proton groups, bins in shared memory | 1: canonical I (4 sites) and O1 (3 sites): register accumulators |
// 2: + a canonical J site whose two instances the thread handles with two sets of register factors
template <int MODE>
struct EnumCta {
    const EnumParams &p;
    double *PB, *CB, *Qc, *QJc, *P2lo, *P2hi, *cE, *fF, *EF, *minF, *EIF, *EJF, *ELo, *EHi, *D, *sU, *EO;
    int32_t *aF, *rF, *misc;     // misc[0] = protons of the chunk above the F minimum
    uint8_t *bO, *nrelI, *nrelO1, *nrelO2, *nrelM;
    int16_t *idxI, *idxLo, *idxHi, *innerInst, *innerSite, *innerOff, *oInst;

    PCE_DEV EnumCta (const EnumParams &params, unsigned char *smem) : p (params) {
        const EnumLayout L (params);
        PB = (double *) (smem + L.off_pb); CB = (double *) (smem + L.off_cb); Qc = (double *) (smem + L.off_qc); QJc = (double *) (smem + L.off_qj);
        P2lo = (double *) (smem + L.off_p2lo); P2hi = (double *) (smem + L.off_p2hi); cE = (double *) (smem + L.off_ce);
        fF = (double *) (smem + L.off_ff); EF = (double *) (smem + L.off_ef); minF = (double *) (smem + L.off_minf);
        EIF = (double *) (smem + L.off_eif); EJF = EIF + kIStates; ELo = (double *) (smem + L.off_elo); EHi = (double *) (smem + L.off_ehi);
        D = (double *) (smem + L.off_d); sU = (double *) (smem + L.off_su); EO = (double *) (smem + L.off_eo);
        aF = (int32_t *) (smem + L.off_af); rF = (int32_t *) (smem + L.off_rf); misc = (int32_t *) (smem + L.off_misc);
        bO = smem + L.off_bo; nrelI = smem + L.off_nrel; nrelO1 = nrelI + kIStates; nrelO2 = nrelO1 + kO1States; nrelM = nrelO2 + kO2States;
        idxI = (int16_t *) (smem + L.off_idx); idxLo = idxI + kIStates * (kMaxI + 1); idxHi = idxLo + params.P_lo * kMaxM;
        innerInst = (int16_t *) (smem + L.off_inner); innerSite = innerInst + params.nInner; innerOff = innerSite + params.nInner;
        oInst = innerOff + params.nS + 1;
    }

    // once per CTA: stage the tables, load the thread's registers, zero the bins
    PCE_DEV void init (int tid, EnumThread &r) const {
        for (int e = tid; e < p.NbP * kBlock; e += kBlock) PB[e] = 0.0;
        for (int e = tid; e < p.P_O2 * p.P_lo; e += kBlock) P2lo[e] = p.P2lo[e];
        for (int e = tid; e < p.P_O2 * p.P_hi; e += kBlock) P2hi[e] = p.P2hi[e];
        for (int e = tid; e < p.nU * p.nV; e += kBlock) D[e] = p.D[e];
        for (int e = tid; e < p.nU; e += kBlock) sU[e] = p.sigmaU[e];
        for (int e = tid; e < p.P_O; e += kBlock) EO[e] = p.EO[e];
        for (int e = tid; e < p.P_O * kMaxO; e += kBlock) bO[e] = p.bO[e];
        for (int e = tid; e < kIStates; e += kBlock) nrelI[e] = p.nrelI[e];
        for (int e = tid; e < kO1States; e += kBlock) nrelO1[e] = p.nrelO1[e];
        for (int e = tid; e < kO2States; e += kBlock) nrelO2[e] = p.nrelO2[e];
        for (int e = tid; e < kBlock; e += kBlock) nrelM[e] = p.nrelM[e];
        for (int e = tid; e < kIStates * (kMaxI + 1); e += kBlock) idxI[e] = p.idxI[e];
        for (int e = tid; e < p.P_lo * kMaxM; e += kBlock) idxLo[e] = p.idxLo[e];
        for (int e = tid; e < p.P_hi * kMaxM; e += kBlock) idxHi[e] = p.idxHi[e];
        for (int e = tid; e < p.nInner; e += kBlock) { innerInst[e] = p.innerInst[e]; innerSite[e] = p.innerSite[e]; }
        for (int e = tid; e <= p.nS; e += kBlock) innerOff[e] = p.innerOff[e];
        for (int e = tid; e < p.nOinst; e += kBlock) oInst[e] = p.oInst[e];
        PCE_UNROLL
        for (int c = 0; c < (MODE == 2 ? kXStates : kIStates); c++) r.xs[c] = p.xs[c * kBlock + tid];
        PCE_UNROLL
        for (int t1 = 0; t1 < kO1States; t1++) r.p1[t1] = p.P1[t1 * kBlock + tid];
        r.mlo = p.mlo[tid]; r.mhi = p.mhi[tid]; r.nrelM = p.nrelM[tid];
        r.A = 0.0;
    }

    // P0: digits of the chunk
    PCE_DEV void phase0 (int tid, long long chunk) const {
        if (tid < p.nF) {
            const int shift = p.fShift[tid];
            const int d = shift >= 0 ? (int) (((unsigned long long) chunk >> shift) & (unsigned long long) (p.fRad[tid] - 1))
                                     : (int) ((unsigned long long) (chunk / p.fStride[tid]) % (unsigned long long) p.fRad[tid]);
            aF[tid] = p.fInst[p.fOff[tid] + d];
            rF[tid] = p.fRel[p.fOff[tid] + d];
        }
    }
    // P1: fields of the F configuration on the inner and O instances, F-F energy terms; own column of the chunk bins zeroed
    PCE_DEV void phase1 (int tid) const {
        const bool folded = !p.unfolded;
        const int n1 = p.nInner, n2 = p.nInner + p.nOinst, n3 = n2 + p.nF;
        for (int j = tid; j < n3; j += kBlock) {
            const int a = j < n1 ? innerInst[j] : j < n2 ? oInst[j - n1] : aF[j - n2];
            double acc = p.gt[a];
            if (folded) {
                const int nf = j < n2 ? p.nF : j - n2;          // F-F pairs once: f2 < f
                const double *row = p.w + (size_t) a * p.ldw;
                int f = 0;
                for (; f + 4 <= nf; f += 4) {                   // four independent gathers in flight (L2 latency once per batch), added in order
                    double v[4];
                    PCE_UNROLL
                    for (int u = 0; u < 4; u++) v[u] = row[aF[f + u]];
                    PCE_UNROLL
                    for (int u = 0; u < 4; u++) acc += v[u];
                }
                for (; f < nf; f++) acc += row[aF[f]];
            }
            fF[j] = acc;
        }
        for (int row = 0; row < p.NbC; row++) CB[row * kBlock + tid] = 0.0;
        if (tid == kBlock - 1) {
            int np = 0;
            for (int f = 0; f < p.nF; f++) np += rF[f];
            misc[0] = np;
        }
    }
    // P2: per-site minima of the inner fields and the normalised factor of every inner instance
    PCE_DEV void phase2 (int tid) const {
        if (tid < p.nInner) {
            const int s = innerSite[tid];
            double vmin = DBL_MAX;
            for (int k = innerOff[s]; k < innerOff[s + 1]; k++) vmin = fmin (vmin, fF[k]);
            EF[tid] = exp (-p.beta * (fF[tid] - vmin));
            if (tid == innerOff[s]) minF[s] = vmin;
        }
    }
    // P3: tile constants (with the bin scale of the tile's chunk + O2 protons) and the chunk factors of the I / Lo / Hi configurations
    PCE_DEV void phase3 (int tid) const {
        const int n_items = p.P_O + kIStates + p.P_lo + p.P_hi;
        if (tid < 2) EJF[tid] = p.nJ ? EF[p.idxJ[tid]] : (tid == 0 ? 1.0 : 0.0);
        for (int e = tid; e < n_items; e += kBlock) {
            if (e < p.P_O) {
                const int t = e;
                double sum = EO[t] + *p.shiftS - p.gref;
                for (int f = 0; f < p.nF; f++) sum += fF[p.nInner + p.nOinst + f];
                for (int s = 0; s < p.nS; s++) sum += minF[s];
                const uint8_t *row = bO + t * kMaxO;
                for (int o = 0; o < p.nO; o++) sum += fF[p.nInner + row[o]];
                const int u = misc[0] + nrelO2[t / p.P_O1];
                cE[t] = exp (-p.beta * sum + sU[u]);
            } else if (e < p.P_O + kIStates) {
                const int i = e - p.P_O;
                double v = 0.0;
                if (i < p.P_I) { v = 1.0; for (int l = 0; l < p.nI; l++) v *= EF[idxI[i * (kMaxI + 1) + l]]; }
                EIF[i] = v;
            } else if (e < p.P_O + kIStates + p.P_lo) {
                const int x = e - p.P_O - kIStates;
                double v = 1.0;
                for (int k = 0; k < p.nLo; k++) v *= EF[idxLo[x * kMaxM + k]];
                ELo[x] = v;
            } else {
                const int x = e - p.P_O - kIStates - p.P_lo;
                double v = 1.0;
                for (int k = 0; k < p.nM - p.nLo; k++) v *= EF[idxHi[x * kMaxM + k]];
                EHi[x] = v;
            }
        }
    }
    // P4: the tile table of this chunk and the thread's chunk factor
    PCE_DEV void phase4 (int tid, EnumThread &r) const {
        // (the Q entries come from L2: four loads of the thread are issued before the first store, two latencies instead of eight)
        const int total = p.P_O * kIStates;
        for (int e0 = tid; e0 < total; e0 += 4 * kBlock) {
            double qv[4];
            PCE_UNROLL
            for (int k = 0; k < 4; k++) { const int e = e0 + k * kBlock; qv[k] = e < total ? p.Q[e] : 0.0; }
            PCE_UNROLL
            for (int k = 0; k < 4; k++) { const int e = e0 + k * kBlock; if (e < total) Qc[e] = qv[k] * (cE[e / kIStates] * EIF[e % kIStates]); }
        }
        for (int e = tid; e < p.P_O * 2; e += kBlock) QJc[e] = p.QJ[e] * EJF[e & 1];
        r.A = ELo[r.mlo] * EHi[r.mhi];
    }

    // main: the thread's P_O x P_I x P_J states of this chunk
    PCE_DEV void accumulate (int tid, const EnumThread &r) const {
        const int nrelF = misc[0];
        if (tid >= p.P_M) return;
        if constexpr (MODE > 0) {
            // I = 4 and O1 = 3 canonical sites (and the J site): protons above the minimum are popcounts, so the bins of the
            // I x J x O1 configurations are register accumulators indexed at compile time; flushed when the O2 proton count
            // changes (O2 is sorted by it).  Sorted I configurations: proton groups of 1, 4, 6, 4, 1.
            constexpr int NACC = MODE == 2 ? 9 : 8;
            double acc[NACC];
            PCE_UNROLL
            for (int j = 0; j < NACC; j++) acc[j] = 0.0;
            int g_prev = nrelO2[0];
            for (int t2 = 0; t2 < p.P_O2; t2++) {
                const int g = nrelO2[t2];
                if (g != g_prev) { flush_canon<NACC> (tid, r, acc, nrelF, g_prev); g_prev = g; }
                const double B = r.A * P2lo[t2 * p.P_lo + r.mlo] * P2hi[t2 * p.P_hi + r.mhi];
                const double *q = Qc + (size_t) t2 * (kO1States * kIStates);
                const double *qj = QJc + (size_t) t2 * (kO1States * 2);
                PCE_UNROLL
                for (int t1 = 0; t1 < kO1States; t1++) {
                    const double *qt = q + t1 * kIStates;
                    const double pb = r.p1[t1] * B;
                    constexpr int pc[8] = { 0, 1, 1, 2, 1, 2, 2, 3 };
                    PCE_UNROLL
                    for (int j = 0; j < (MODE == 2 ? 2 : 1); j++) {
                        const double *x = r.xs + j * kIStates;
                        const double g0 = qt[0] * x[0];
                        const double g1 = qt[1] * x[1] + qt[2] * x[2] + qt[3] * x[3] + qt[4] * x[4];
                        const double g2 = qt[5] * x[5] + qt[6] * x[6] + qt[7] * x[7] + qt[8] * x[8] + qt[9] * x[9] + qt[10] * x[10];
                        const double g3 = qt[11] * x[11] + qt[12] * x[12] + qt[13] * x[13] + qt[14] * x[14];
                        const double g4 = qt[15] * x[15];
                        const double pbj = MODE == 2 ? pb * qj[t1 * 2 + j] : pb;
                        acc[pc[t1] + j + 0] += pbj * g0; acc[pc[t1] + j + 1] += pbj * g1; acc[pc[t1] + j + 2] += pbj * g2;
                        acc[pc[t1] + j + 3] += pbj * g3; acc[pc[t1] + j + 4] += pbj * g4;
                    }
                }
            }
            flush_canon<NACC> (tid, r, acc, nrelF, g_prev);
        } else {
            for (int t2 = 0; t2 < p.P_O2; t2++) {
                const int g = nrelO2[t2];
                const double B = r.A * P2lo[t2 * p.P_lo + r.mlo] * P2hi[t2 * p.P_hi + r.mhi];
                PCE_UNROLL
                for (int t1 = 0; t1 < kO1States; t1++) {
                    if (t1 < p.P_O1) {
                        const double *qt = Qc + (size_t) (t2 * p.P_O1 + t1) * kIStates;
                        const double pb = r.p1[t1] * B;
                        const int v0 = nrelO1[t1];
                        const double *drow = D + (size_t) (nrelF + g) * p.nV + v0 + r.nrelM;
                        double *cb = CB + (size_t) (g + v0) * kBlock + tid;
                        double a = 0.0;
                        PCE_UNROLL
                        for (int i = 0; i < kIStates; i++) {
                            if (i < p.P_I) {
                                a += qt[i] * r.xs[i];
                                if ((p.flushI >> i) & 1u) { const int n = nrelI[i]; cb[n * kBlock] += pb * a * drow[n]; a = 0.0; }
                            }
                        }
                    }
                }
            }
        }
    }
    template <int NACC>
    PCE_DEV void flush_canon (int tid, const EnumThread &r, double *acc, int nrelF, int g) const {
        const double *drow = D + (size_t) (nrelF + g) * p.nV + r.nrelM;
        double *cb = CB + (size_t) g * kBlock + tid;
        PCE_UNROLL
        for (int j = 0; j < NACC; j++) { cb[j * kBlock] += acc[j] * drow[j]; acc[j] = 0.0; }
    }

    // P5a: the chunk bins go into the thread's persistent bins (own column only)
    PCE_DEV void fold (int tid) const {
        if (tid >= p.P_M) return;
        const int nrelF = misc[0];
        for (int jc = 0; jc < p.NbC; jc++) PB[(size_t) (nrelF + jc) * kBlock + tid] += CB[(size_t) jc * kBlock + tid];
    }
    // P5b: sums of the chunk bins over the threads -> FtabT (one warp per row; lanes combine in a fixed butterfly order).
    // The chunk bins are relative to each thread's own M protons; row r of FtabT is absolute in them:
    // partial (r, lane) = sum over k of CB[r - nrelM(t)][t], t = lane + 32 k
    PCE_DEV double row_partial (int row, int lane) const {
        double v = 0.0;
        PCE_UNROLL
        for (int k = 0; k < kBlock / 32; k++) {
            const int t = lane + 32 * k, rel = row - (int) nrelM[t];
            if (rel >= 0 && rel < p.NbC) v += CB[(size_t) rel * kBlock + t];
        }
        return v;
    }
    PCE_DEV void store_row (int row, long long chunk_local, double total) const { p.FtabT[(size_t) (misc[0] + row) * p.ldF + chunk_local] = total; }

    // end of the CTA: persistent bins to global memory
    PCE_DEV void finish (int tid, int cta) const {
        double *out = p.T + (size_t) cta * p.NbP * kBlock;
        for (int e = tid; e < p.NbP * kBlock; e += kBlock) out[e] = PB[e];
    }
};

// ---------------------------------------------------------------------------------------------------
// post-processing: bins of Z and of the instances of the M and F sites from the per-thread and per-chunk bins
// ---------------------------------------------------------------------------------------------------
struct MarginalParams {
    int Nb, base, NbP, NbAll, P_M, nM, nF, write_z, grid;
    long long chunk_begin, nchunks_local, ldF;
    const uint8_t *nrelM, *digitM;           // [kBlock], [kBlock][kMaxM]
    const int16_t *rowM, *fInst;             // marginal rows
    const int32_t *rowMoff, *fOff, *fRad;
    const long long *fStride;
    const int32_t *fShift;                   // >= 0: stride and radix of the F site are powers of two (digit = (chunk >> shift) & (radix - 1))
    const uint8_t *coverM, *coverF;
    const double *Mtab, *FtabT;
    double *bins;                            // [(1 + ninst)][Nb], accumulated
};

// which source a marginal block sums: kind 0 = Z, 1 = M site digit, 2 = F site digit; returns false for "nothing to do"
PCE_DEV bool marginal_target (const MarginalParams &p, int b, int *kind, int *site, int *digit, int *row) {
    if (b == 0) { *kind = 0; *site = 0; *digit = 0; *row = 0; return p.write_z != 0; }
    b -= 1;
    for (int i = 0; i < p.nM; i++) {
        const int n = p.rowMoff[i + 1] - p.rowMoff[i];
        if (b < n) { *kind = 1; *site = i; *digit = b; *row = 1 + p.rowM[p.rowMoff[i] + b]; return p.coverM[i] != 0; }
        b -= n;
    }
    for (int f = 0; f < p.nF; f++) {
        const int n = p.fRad[f];
        if (b < n) { *kind = 2; *site = f; *digit = b; *row = 1 + p.fInst[p.fOff[f] + b]; return p.coverF[f] != 0; }
        b -= n;
    }
    return false;
}

// contribution of element e (thread of the M table, or chunk) of bin n_rel to the block's sum
PCE_DEV double marginal_term (const MarginalParams &p, int kind, int site, int digit, int n_rel, long long e) {
    if (kind == 2) {
        const unsigned long long chunk = (unsigned long long) (p.chunk_begin + e);
        const int shift = p.fShift[site];
        const int d = shift >= 0 ? (int) ((chunk >> shift) & (unsigned long long) (p.fRad[site] - 1))
                                 : (int) ((chunk / (unsigned long long) p.fStride[site]) % (unsigned long long) p.fRad[site]);
        if (d != digit) return 0.0;
        return p.FtabT[(size_t) n_rel * p.ldF + e];
    }
    if (kind == 1 && p.digitM[e * kMaxM + site] != digit) return 0.0;
    const int rel = n_rel - (int) p.nrelM[e];
    return (rel >= 0 && rel < p.NbP) ? p.Mtab[(size_t) rel * kBlock + e] : 0.0;
}

// ---------------------------------------------------------------------------------------------------
// host helpers shared by the library and the emulator: sizes and table pointers of a run into the kernel parameters
// ---------------------------------------------------------------------------------------------------
constexpr size_t kNumericDoubles = (size_t) kXStates * kBlock + (size_t) kO1States * kBlock + (size_t) kO2States * kBlock + (size_t) kO2States * 64 +
                                   (size_t) kMaxTiles * kIStates + (size_t) kMaxTiles * 2 + kMaxTiles + 8;

// `locate (vector)` returns the address of the (host or device) copy of one of the RunTables vectors;
// `numeric` is a buffer of kNumericDoubles doubles for the tables the static program writes
template <typename Locate>
inline void bind_run (const RunTables &T, EnumParams &p, MarginalParams &mp, Locate &&locate, double *numeric) {
    p.nM = T.nM; p.nLo = T.nLo; p.nI = T.nI; p.nJ = T.nJ; p.nO1 = T.nO1; p.nO2 = T.nO2; p.nO = T.nO1 + T.nO2; p.nF = T.nF; p.nS = T.nS;
    p.P_lo = T.P_lo; p.P_hi = T.P_hi; p.P_M = T.P_M; p.P_I = T.P_I; p.P_J = T.P_J; p.P_O1 = T.P_O1; p.P_O2 = T.P_O2; p.P_O = T.P_O;
    p.nInner = T.nInner; p.nMinst = T.nMinst; p.nOinst = T.nOinst;
    p.NbC = T.NbC; p.NbCp = T.NbCp; p.NbP = T.NbP; p.NbAll = T.NbAll; p.nU = T.nU; p.nV = T.nV; p.mode = T.mode;
    p.flushI = T.flushI;
    p.innerInst = locate (T.innerInst); p.innerSite = locate (T.innerSite); p.innerOff = locate (T.innerOff); p.oInst = locate (T.oInst);
    p.idxM = locate (T.idxM); p.idxLo = locate (T.idxLo); p.idxHi = locate (T.idxHi); p.idxI = locate (T.idxI); p.idxJ = locate (T.idxJ);
    p.mlo = locate (T.mlo); p.mhi = locate (T.mhi); p.fInst = locate (T.fInst);
    p.nrelM = locate (T.nrelM); p.nrelI = locate (T.nrelI); p.nrelO1 = locate (T.nrelO1); p.nrelO2 = locate (T.nrelO2);
    p.bO = locate (T.bO); p.fRel = locate (T.fRel); p.fOff = locate (T.fOff); p.fRad = locate (T.fRad); p.fStride = locate (T.fStride); p.fShift = locate (T.fShift);
    p.D = locate (T.D); p.sigmaU = locate (T.sigmaU);
    double *o = numeric;
    p.xs = o; o += (size_t) kXStates * kBlock;
    p.P1 = o; o += (size_t) kO1States * kBlock;
    p.P2lo = o; o += (size_t) kO2States * kBlock;
    p.P2hi = o; o += (size_t) kO2States * 64;
    p.Q = o; o += (size_t) kMaxTiles * kIStates;
    p.QJ = o; o += (size_t) kMaxTiles * 2;
    p.EO = o; o += kMaxTiles;
    p.shiftS = o;
    p.needF = 0;
    for (int f = 0; f < T.nF; f++) if (T.coverF[(size_t) f]) p.needF = 1;
    mp.NbP = T.NbP; mp.NbAll = T.NbAll; mp.P_M = T.P_M; mp.nM = T.nM; mp.nF = T.nF;
    mp.nrelM = p.nrelM; mp.digitM = locate (T.digitM); mp.rowM = locate (T.rowM); mp.rowMoff = locate (T.rowMoff);
    mp.fInst = p.fInst; mp.fOff = p.fOff; mp.fRad = p.fRad; mp.fStride = p.fStride; mp.fShift = p.fShift;
    mp.coverM = locate (T.coverM); mp.coverF = locate (T.coverF);
}

// number of blockIdx.x values of the marginal program: Z, every digit of every M site, every digit of every F site
inline int marginal_rows (const RunTables &T) {
    int rows = 1 + (int) T.rowM.size ();
    for (int f = 0; f < T.nF; f++) rows += T.fRad[(size_t) f];
    return rows;
}

}  // namespace pce_enum
