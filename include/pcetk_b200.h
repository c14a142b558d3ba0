/* pcetk_b200.h -- C ABI of the B200 microstate engine.
 *
 * Drop-in boundary for the hot path of efliks/pcetk: what the reference's Cython layer
 * (extensions/cython/ContinuumElectrostatics.{StateVector,EnergyModel,MCModelDefault}.pyx) binds from
 *   extensions/cinclude/StateVector.h:53-84, EnergyModel.h:64-101, MCModelDefault.h:58-71
 * is reachable from here through opaque handles, plain pointers and sizes (no torch types), so that
 * ctypes / Cython / cgo / JNI can bind it.  Data functions (setters, getters, StateVector_*) map one
 * to one.  The compute functions map at the granularity a GPU launch has: the reference's inner-loop
 * steps have no entry point of their own --
 *   MCModelDefault_Metropolis / _Move / _DoubleMove / _MCScan / _UpdateProbabilities (MCModelDefault.h:63-67)
 *     run inside pce_mc_run* (one launch = _Equilibration + _Production of every chain);
 *   EnergyModel_CalculateZ / _CalculateProbabilitiesFromZ (EnergyModel.h:81-82) run inside pce_analytic*.
 * Each declaration names the reference interface it replaces.  See INTEGRATION.md for the binding a
 * reference maintainer would add (including the logging / trajectory branch of MCModelDefault.pyx:79-159).
 *
 * Conventions
 *   - every function returns a pce_status (0 = ok); pce_last_error() gives the text for the calling thread
 *     (the reference threads a `Status *` out-parameter instead, Status.h in pCore);
 *   - instance indices are GLOBAL instance indices unless a parameter is called `local`;
 *   - a model handle is safe for one call at a time; results are returned in caller-owned buffers and
 *     are also stored in the model's probability array, as the reference does (EnergyModel.h:50);
 *   - pointers named d_* are DEVICE pointers on the model's device, all others are HOST pointers;
 *   - kernels are launched on the stream set with pce_set_stream (default: the legacy default stream);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with PCE_ERR_CUDA.
 */
#ifndef PCETK_B200_H
#define PCETK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pce_model pce_model; /* replaces EnergyModel + its private StateVector (EnergyModel.h:39-60) */
typedef struct pce_mc    pce_mc;    /* replaces MCModelDefault (MCModelDefault.h:41-54) */
typedef struct pce_state pce_state; /* replaces StateVector (StateVector.h:26-50); its pair list lives in pce_mc */

typedef enum {
    PCE_OK            = 0,
    PCE_ERR_ALLOC     = 1, /* Status_MemoryAllocationFailure */
    PCE_ERR_INDEX     = 2, /* Status_IndexOutOfRange */
    PCE_ERR_VALUE     = 3, /* Status_ValueError */
    PCE_ERR_SHAPE     = 4, /* Status_ArrayNonConformableSizes */
    PCE_ERR_STATE     = 5, /* call order: sites not set, interactions not symmetrised, ... */
    PCE_ERR_CUDA      = 6, /* CUDA runtime error or no device */
    PCE_ERR_LIMIT     = 7  /* too many states / instances for the requested kernel */
} pce_status;

/* ---- library / device -------------------------------------------------------------------------- */
const char *pce_last_error (void);
const char *pce_version (void);
int pce_device_count (int *count);
int pce_device_name (int device, char *buffer, int length);
int pce_set_stream (pce_model *model, void *cuda_stream);       /* cudaStream_t; NULL = default stream */
int pce_synchronize (pce_model *model);
/* launch counter: kernels launched by this library since the last reset (bench.py's gpu_launches) */
long long pce_launch_count (int reset);
/* measurement hook (no reference counterpart): streaming read bandwidth in GB/s of a zero-filled buffer of `bytes`
 * read `passes` times with 128-bit loads -- L2 bandwidth for buffers well below the L2 size, HBM for large ones */
int pce_probe_read_bandwidth (int device, long long bytes, int passes, double *gb_per_s);
/* on-chip peaks measured on this device (denominators of bench.py's rooflines): kind 0 = DFMA rate in 1e9 DFMA/s,
 * kind 1 = conflict-free shared-memory read bandwidth in GB/s, kind 2 = issue rate in 1e9 warp-instructions/s */
int pce_probe_on_chip (int device, int kind, double *rate);

/* ---- model: allocation (EnergyModel_Allocate / _Deallocate, EnergyModel.h:64-65) ---------------- */
int  pce_model_create (int nsites, int ninstances, double temperature, int device, pce_model **out);
void pce_model_destroy (pce_model *model);
int  pce_model_nsites (const pce_model *model);
int  pce_model_ninstances (const pce_model *model);
double pce_model_temperature (const pce_model *model);

/* StateVector_SetSite (StateVector.h:68) for the model's site table; ranges must be contiguous and
 * ascending across sites (EnergyModel.c:217-221 relies on it). */
int pce_model_set_site (pce_model *model, int site, int first, int last);
int pce_model_set_sites (pce_model *model, const int32_t *first, const int32_t *last);
int pce_model_get_site (const pce_model *model, int site, int *first, int *last);
/* number of microstates as a double (exact below 2^53); EnergyModel.pyx:86-110 computes the capped int */
int pce_model_nstates (const pce_model *model, double *nstates);

/* EnergyModel_Set* / _Get* (EnergyModel.h:87-99) */
int pce_model_set_gmodel (pce_model *model, int instance, double value);
int pce_model_set_gintr (pce_model *model, int instance, double value);
int pce_model_set_protons (pce_model *model, int instance, int value);
int pce_model_set_probability (pce_model *model, int instance, double value);
int pce_model_set_interaction (pce_model *model, int instance_a, int instance_b, double value);
int pce_model_get_gmodel (const pce_model *model, int instance, double *value);
int pce_model_get_gintr (const pce_model *model, int instance, double *value);
int pce_model_get_protons (const pce_model *model, int instance, int *value);
int pce_model_get_probability (const pce_model *model, int instance, double *value);
int pce_model_get_interaction (const pce_model *model, int instance_a, int instance_b, double *value);
int pce_model_get_interaction_symmetric (const pce_model *model, int instance_a, int instance_b, double *value); /* EnergyModel_GetInterSymmetric */
int pce_model_get_deviation (const pce_model *model, int instance_a, int instance_b, double *value);             /* EnergyModel_GetDeviation */
/* bulk forms of the setters / getters (any pointer may be NULL to skip it); interactions is
 * ninstances x ninstances row-major, raw (EnergyModel.h:46) */
int pce_model_load (pce_model *model, const double *gintr, const double *gmodel, const int32_t *protons, const double *interactions);
int pce_model_get_probabilities (const pce_model *model, double *probabilities);
int pce_model_set_probabilities (pce_model *model, const double *probabilities);
int pce_model_get_symmetric_matrix (const pce_model *model, double *w /* ninstances^2 */);

/* EnergyModel_CheckInteractionsSymmetric / _SymmetrizeInteractions / _ResetInteractions /
 * _ScaleInteractions (EnergyModel.h:68-71) */
int pce_model_check_symmetric (const pce_model *model, double tolerance, double *max_deviation, int *is_symmetric);
int pce_model_symmetrize (pce_model *model);      /* the averaging runs on the device at the next upload; host readers of the symmetric
                                                    * matrix see the same values (formed on demand); it changes on this call only */
int pce_model_reset_interactions (pce_model *model);
int pce_model_scale_interactions (pce_model *model, double scale);
/* EnergyModel_StateVectorFromProbabilities (EnergyModel.h:72), without the reference's off-by-one */
int pce_model_state_from_probabilities (const pce_model *model, int32_t *state_global);

/* copy the model to its device (done implicitly by compute calls when the host copy changed).  Device tables derived from the
 * model (sampler tables, enumeration plan) are rebuilt only when the uploaded content differs from the previous upload. */
int pce_model_upload (pce_model *model);

/* ---- state vector (StateVector.h:53-84; host object, bound by ContinuumElectrostatics.StateVector.pyx) ----------------
 * One active instance per site, as GLOBAL instance indices; `local` = index within the site.  Getters write -1 and return
 * PCE_ERR_INDEX for a site out of range (the reference returns -1 and sets Status_IndexOutOfRange, StateVector.c:236-292);
 * setters return PCE_ERR_VALUE for an instance outside the site ("Invalid value (%d).", Status_ValueError).
 * pce_state_active() is the int32 row pce_energy / pce_mc_run_chains exchange: no conversion between the two. */
int  pce_state_create (int nsites, pce_state **out);                                   /* StateVector_Allocate :53 */
int  pce_state_create_for_model (const pce_model *model, pce_state **out);             /* Allocate + SetSite per site, StateVector.pyx:80-103 */
void pce_state_destroy (pce_state *state);                                             /* StateVector_Deallocate :57 */
int  pce_state_clone (const pce_state *state, pce_state **out);                        /* StateVector_Clone :60 */
int  pce_state_copy_to (const pce_state *state, pce_state *other);                     /* StateVector_CopyTo :61 (lengths are checked here) */
int  pce_state_nsites (const pce_state *state);
int  pce_state_nsubstate (const pce_state *state);
int  pce_state_reset (pce_state *state);                                               /* StateVector_Reset :64 */
int  pce_state_reset_substate (pce_state *state);                                      /* StateVector_ResetSubstate :65 */
int  pce_state_reset_to_maximum (pce_state *state);                                    /* StateVector_ResetToMaximum :66 */
int  pce_state_randomize (pce_state *state, unsigned long long seed);                  /* StateVector_Randomize :67 (Philox words instead of MT19937) */
int  pce_state_set_site (pce_state *state, int site, int first, int last);             /* StateVector_SetSite :70 */
int  pce_state_get_site (const pce_state *state, int site, int *first, int *last);
int  pce_state_is_substate (const pce_state *state, int site, int *flag);              /* StateVector_IsSubstate :73 */
int  pce_state_get_item (const pce_state *state, int site, int *local);                /* StateVector_GetItem :74 */
int  pce_state_set_item (pce_state *state, int site, int local);                       /* StateVector_SetItem :75 */
int  pce_state_get_actual_item (const pce_state *state, int site, int *global);        /* StateVector_GetActualItem :76 */
int  pce_state_set_actual_item (pce_state *state, int site, int global);               /* StateVector_SetActualItem :77 */
int  pce_state_allocate_substate (pce_state *state, int nssites);                      /* StateVector_AllocateSubstate :54 (a second call fails) */
int  pce_state_get_substate_item (const pce_state *state, int index, int *site);       /* StateVector_GetSubstateItem :78 */
int  pce_state_set_substate_item (pce_state *state, int selected_site, int index);     /* StateVector_SetSubstateItem :79 */
int  pce_state_increment (pce_state *state, int *more);                                /* StateVector_Increment :82: *more = 0 and the all-first state after the last state */
int  pce_state_increment_substate (pce_state *state, int *more);                       /* StateVector_IncrementSubstate :83 */
/* bulk forms (no reference counterpart) */
const int32_t *pce_state_active (const pce_state *state);                              /* the row of global indices, valid until the vector is destroyed */
int  pce_state_get_active (const pce_state *state, int32_t *out);
int  pce_state_set_active (pce_state *state, const int32_t *in);                       /* validated against the site ranges */
int  pce_state_set_index (pce_state *state, long long index);                          /* mixed-radix decode: the state `index` increments after the all-first state */
int  pce_state_get_index (const pce_state *state, long long *index);
/* every state of the substate in IncrementSubstate order (the loop of Substate.py:99-108) as rows of nsites global indices:
 * one batch for pce_energy.  states == NULL: only *count.  Leaves the substate at its first state. */
int  pce_state_enumerate_substate (pce_state *state, long long capacity, int32_t *states, long long *count);

/* ---- kernel 1: microstate energies --------------------------------------------------------------
 * EnergyModel_CalculateMicrostateEnergy[Unfolded] (EnergyModel.h:75-76), batched: `states` holds
 * nstates rows of nsites GLOBAL instance indices; out[s * nph + q] = G(state s, pH q).
 * Bit-identical to the reference's arithmetic (same summation order, no FMA contraction). */
int pce_energy (pce_model *model, const int32_t *states, long long nstates, const double *ph, int nph, int unfolded, double *out);
int pce_energy_device (pce_model *model, const int32_t *d_states, long long nstates, const double *d_ph, int nph, int unfolded, double *d_out);

/* ---- kernel 2: exhaustive enumeration ------------------------------------------------------------
 * EnergyModel_CalculateProbabilitiesAnalytically[Unfolded], _CalculateZfolded/_Zunfolded
 * (EnergyModel.h:79-86) for a whole pH grid in one call.  States are binned by proton count, so one
 * enumeration serves every pH value (kernel 2).
 *
 * pce_analytic_bins_size: number of doubles in a bins buffer for this model.
 * pce_analytic_bins     : enumerate shard `shard` of `nshards` (equal slices of the outer state index)
 *                         into d_bins (device, zeroed by the call).  Bins of different shards ADD, which
 *                         is the only multi-GPU exchange the path needs (one all-reduce(sum)).  Bin n of
 *                         row r is stored scaled by exp(sigma[n]) (pce_analytic_bin_scales): the scales
 *                         depend on the model and the pH list only, so every shard uses the same ones.
 * pce_analytic_bin_scales: (host) sigma[0 .. bins_size / (1 + ninst)) for this pH list.
 * pce_analytic_finish   : (host) turn summed bins into probabilities for each pH: out_p[q*ninst + k];
 *                         out_lnz[q] = ln sum_s exp(-(G_s(pH_q) - gzero)/RT) (may be NULL); also stores
 *                         the last pH's probabilities in the model.
 * pce_analytic          : the three steps on one device with host buffers.
 * max_states: refuse above this many states (reference: 2^26, EnergyModel.pyx:11,156); <= 0 = engine limit. */
long long pce_analytic_bins_size (const pce_model *model);
int pce_analytic_bins (pce_model *model, const double *ph, int nph, int unfolded, int shard, int nshards, double max_states, double *d_bins);
int pce_analytic_finish (pce_model *model, const double *bins /* host */, const double *ph, int nph, int unfolded, double gzero,
                         double *out_p, double *out_lnz);
int pce_analytic_bin_scales (pce_model *model, const double *ph, int nph, int unfolded, double *sigma);
int pce_analytic (pce_model *model, const double *ph, int nph, int unfolded, double gzero, double max_states, double *out_p, double *out_lnz);
/* exact minimum microstate energy at one pH (the shift the reference applies, EnergyModel.c:262-275):
 * with it, Z as returned by EnergyModel_CalculateZfolded is exp(lnz + Gmin/RT). */
int pce_analytic_gmin (pce_model *model, double ph, int unfolded, double max_states, double *gmin);

/* ---- kernel 3: Metropolis Monte Carlo -----------------------------------------------------------
 * MCModelDefault_Allocate / _Deallocate / _LinkToEnergyModel / _FindPairs (MCModelDefault.h:58-68).
 * seed <= 0 seeds from the clock like the reference (MCModelDefault.c:32-34).  nchains independent
 * chains are run per pH value (the reference runs one). */
int  pce_mc_create (pce_model *model, double limit, int nequil, int nprod, long long seed, int nchains, pce_mc **out);
void pce_mc_destroy (pce_mc *mc);
int  pce_mc_npairs (const pce_mc *mc);
int  pce_mc_get_pair (const pce_mc *mc, int index, int *site_a, int *site_b, double *wmax);   /* StateVector_GetPair */
int  pce_mc_set_scans (pce_mc *mc, int nequil, int nprod);
int  pce_mc_set_chains (pce_mc *mc, int nchains);
int  pce_mc_set_kernel (pce_mc *mc, int kernel);   /* 0 = auto, 1 = chain per thread, 2 = chain per warp */
int  pce_mc_get_info (const pce_mc *mc, int *kernel, int *nchains, int *nequil, int *nprod, long long *seed);
/* launch geometry the sampler would use for nph pH values (any output may be NULL): chains_per_wave is the number of
 * chains resident on the whole device at once -- size nph*nchains as a multiple of it to run in full waves */
int  pce_mc_plan (pce_mc *mc, int nph, int *kernel, int *block, int *ctas_per_sm, int *chains_per_wave, long long *smem_bytes);

/* MCModelDefault_Equilibration + _Production (MCModelDefault.h:70-71) for nph pH values x nchains
 * chains in one launch.  Chains [chain_offset, chain_offset + nchains) of a global ensemble are run
 * (ranks of a multi-GPU job pass disjoint offsets; streams depend only on (seed, pH index offset +
 * q, chain)).  Outputs (any may be NULL):
 *   counts  [nph * ninst]  production scans during which each instance was active, summed over chains
 *   counters[nph * 4]      moves, moves accepted, double moves, double moves accepted (production)
 *   esum    [nph]          sum over chains and production scans of the microstate energy
 * pce_mc_run also stores counts/(nchains*nprod) of the LAST pH value in the model's probabilities.
 * A chain counts its trials in 32 bits: (nequil + nprod) * (nsites + npairs) >= 2^32 returns PCE_ERR_LIMIT. */
int pce_mc_run (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset,
                long long *counts, long long *counters, double *esum);
int pce_mc_run_device (pce_mc *mc, const double *ph /* host */, int nph, int ph_index_offset, long long chain_offset,
                       long long *d_counts, long long *d_counters, double *d_esum);
/* per-chain outputs for parity tests against the chain specification (oracle_mc_philox_chain):
 * chain_counts[nph*nchains*ninst], chain_state[nph*nchains*nsites], chain_energy[nph*nchains] */
int pce_mc_run_chains (pce_mc *mc, const double *ph, int nph, int ph_index_offset, long long chain_offset,
                       long long *chain_counts, int32_t *chain_state, double *chain_energy, double *chain_esum, long long *chain_counters);
/* MCModelDefault.pyx:79-159 trajectory mode: energies[nprod] of chain 0 after every production scan; the chain draws
 * the random stream (seed, pH index ph_index_offset, chain 0) */
int pce_mc_run_trajectory (pce_mc *mc, double ph, int ph_index_offset, double *energies, long long *counts, long long *counters);
/* the same with the per-block counter table of the logging branch: log_rows[(nequil/f + nprod/f) * 5] holds, per block of
 * f = log_frequency scans (equilibration blocks first), {scans done in the phase, moves, moves accepted, double moves,
 * double moves accepted} -- the rows MCModelDefault.pyx:85-151 prints */
int pce_mc_run_logged (pce_mc *mc, double ph, int ph_index_offset, int log_frequency, double *energies, long long *counts, long long *counters,
                       long long *log_rows);

/* ---- test hooks ----------------------------------------------------------------------------------- */
/* Philox4x32-10 blocks computed ON THE DEVICE: out[4*i..] for counters ctr[4*i..], one key */
int pce_test_philox (int device, const uint32_t *ctr, int nblocks, const uint32_t key[2], uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* PCETK_B200_H */
