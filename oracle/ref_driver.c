/* C-ABI facade over the reference's own, unmodified C sources (compiled from /root/reference where they
 * lie; see oracle/Makefile).  Test/benchmark infrastructure only: it plays the role of the Cython layer
 * (ContinuumElectrostatics.EnergyModel.pyx:72-110, ...MCModelDefault.pyx:27-78) for a ctypes caller. */
#include <stdlib.h>
#include <string.h>
#include "EnergyModel.h"
#include "MCModelDefault.h"

typedef struct {
    EnergyModel *em;
    StateVector *vector;   /* caller-visible vector, like the Python StateVector object */
} RefModel;

void *ref_model_create (int nsites, int ninstances, double temperature, const int *first, const int *last,
                        const double *gintr, const double *gmodel, const int *protons, const double *wraw,
                        long long nstates_cap) {
    Status status = Status_Continue;
    RefModel *self = (RefModel *) calloc (1, sizeof (RefModel));
    int i, j;
    long long nstates = 1;
    if (self == NULL) return NULL;
    self->em = EnergyModel_Allocate (nsites, ninstances, &status);
    if (status != Status_Continue || self->em == NULL) { free (self); return NULL; }
    self->em->temperature = temperature;
    self->em->ninstances  = ninstances;
    for (i = 0; i < nsites; i++) {
        StateVector_SetSite (self->em->vector, i, first[i], last[i], &status);
        if (nstates <= nstates_cap) nstates *= (last[i] - first[i] + 1);   /* EnergyModel.pyx:108-109 */
    }
    self->em->nstates = (Integer) (nstates > 2147483647LL ? 2147483647LL : nstates);
    for (i = 0; i < ninstances; i++) {
        EnergyModel_SetGintr   (self->em, i, gintr[i]);
        EnergyModel_SetGmodel  (self->em, i, gmodel[i]);
        EnergyModel_SetProtons (self->em, i, protons[i]);
        for (j = 0; j < ninstances; j++) EnergyModel_SetInteraction (self->em, i, j, wraw[(size_t) i * ninstances + j]);
    }
    EnergyModel_SymmetrizeInteractions (self->em, &status);
    self->vector = StateVector_Clone (self->em->vector, &status);
    return self;
}

void ref_model_destroy (void *handle) {
    RefModel *self = (RefModel *) handle;
    if (self == NULL) return;
    StateVector_Deallocate (self->vector);
    EnergyModel_Deallocate (self->em);
    free (self);
}

int ref_check_symmetric (void *handle, double tolerance, double *max_deviation) {
    RefModel *self = (RefModel *) handle;
    return (int) EnergyModel_CheckInteractionsSymmetric (self->em, tolerance, max_deviation);
}

double ref_get_w (void *handle, int a, int b) {
    RefModel *self = (RefModel *) handle;
    return EnergyModel_GetInterSymmetric (self->em, a, b);
}

/* state = global instance index per site */
double ref_energy (void *handle, const int *state, double pH, int unfolded) {
    RefModel *self = (RefModel *) handle;
    Status status = Status_Continue;
    int i;
    for (i = 0; i < self->vector->nsites; i++) StateVector_SetActualItem (self->vector, i, state[i], &status);
    if (status != Status_Continue) return 0.0 / 0.0;
    return unfolded ? EnergyModel_CalculateMicrostateEnergyUnfolded (self->em, self->vector, pH)
                    : EnergyModel_CalculateMicrostateEnergy         (self->em, self->vector, pH);
}

/* energies of the first n states in odometer order (StateVector_Increment), twosites.py:24-29 */
int ref_enumerate_energies (void *handle, double pH, long long n, double *out) {
    RefModel *self = (RefModel *) handle;
    long long k;
    StateVector_Reset (self->vector);
    for (k = 0; k < n; k++) {
        out[k] = EnergyModel_CalculateMicrostateEnergy (self->em, self->vector, pH);
        if (!StateVector_Increment (self->vector)) { k++; break; }
    }
    return (int) k;
}

int ref_analytic (void *handle, double pH, int unfolded, double *out_p) {
    RefModel *self = (RefModel *) handle;
    Status status = Status_Continue;
    if (unfolded) EnergyModel_CalculateProbabilitiesAnalyticallyUnfolded (self->em, pH, &status);
    else          EnergyModel_CalculateProbabilitiesAnalytically         (self->em, pH, &status);
    if (status != Status_Continue) return (int) status;
    memcpy (out_p, self->em->probabilities->data, sizeof (double) * (size_t) self->em->ninstances);
    return 0;
}

double ref_z (void *handle, double pH, double Gzero, int unfolded) {
    RefModel *self = (RefModel *) handle;
    Status status = Status_Continue;
    return unfolded ? EnergyModel_CalculateZunfolded (self->em, pH, Gzero, &status)
                    : EnergyModel_CalculateZfolded   (self->em, pH, Gzero, &status);
}

/* --- Monte Carlo -------------------------------------------------------------------------------- */
void *ref_mc_create (void *handle, double limit, int nequil, int nprod, int seed) {
    RefModel *self = (RefModel *) handle;
    Status status = Status_Continue;
    Integer npairs;
    MCModelDefault *mc = MCModelDefault_Allocate (limit, nequil, nprod, seed, &status);
    if (mc == NULL) return NULL;
    MCModelDefault_LinkToEnergyModel (mc, self->em, &status);
    npairs = MCModelDefault_FindPairs (mc, -1, &status);               /* MCModelDefault.pyx:47-51 */
    if (npairs > 0) MCModelDefault_FindPairs (mc, npairs, &status);
    return mc;
}

void ref_mc_destroy (void *mc) { if (mc != NULL) MCModelDefault_Deallocate ((MCModelDefault *) mc); }

int ref_mc_npairs (void *handle) { return ((MCModelDefault *) handle)->vector->npairs; }

int ref_mc_get_pair (void *handle, int index, int *site_a, int *site_b, double *wmax) {
    Status status = Status_Continue;
    StateVector_GetPair (((MCModelDefault *) handle)->vector, index, site_a, site_b, wmax, &status);
    return (int) status;
}

/* quiet path of CalculateOwnerProbabilities (MCModelDefault.pyx:68-78) */
int ref_mc_run (void *handle, double pH, double *out_p) {
    MCModelDefault *mc = (MCModelDefault *) handle;
    MCModelDefault_Equilibration (mc, pH);
    MCModelDefault_Production    (mc, pH);
    memcpy (out_p, mc->energyModel->probabilities->data, sizeof (double) * (size_t) mc->energyModel->ninstances);
    return 0;
}

/* logging/trajectory path (MCModelDefault.pyx:79-159): per-scan energies and total counters */
int ref_mc_run_trajectory (void *handle, double pH, double *out_p, double *energies, long long *counters) {
    MCModelDefault *mc = (MCModelDefault *) handle;
    Integer moves = 0, movesAcc = 0, flips = 0, flipsAcc = 0, i;
    Integer nmoves = mc->vector->nsites + mc->vector->npairs;
    StateVector_Randomize (mc->vector, mc->generator);
    for (i = 0; i < mc->nequil; i++) MCModelDefault_MCScan (mc, pH, nmoves, &moves, &movesAcc, &flips, &flipsAcc);
    moves = movesAcc = flips = flipsAcc = 0;
    Real1DArray_Set (mc->energyModel->probabilities, 0.0);
    for (i = 0; i < mc->nprod; i++) {
        Real G = MCModelDefault_MCScan (mc, pH, nmoves, &moves, &movesAcc, &flips, &flipsAcc);
        if (energies != NULL) energies[i] = G;
        MCModelDefault_UpdateProbabilities (mc);
    }
    Real1DArray_Scale (mc->energyModel->probabilities, 1.0 / mc->nprod);
    memcpy (out_p, mc->energyModel->probabilities->data, sizeof (double) * (size_t) mc->energyModel->ninstances);
    if (counters != NULL) { counters[0] = moves; counters[1] = movesAcc; counters[2] = flips; counters[3] = flipsAcc; }
    return 0;
}

/* ---- the caller-visible StateVector, function by function (StateVector.h:53-84): what the Cython class
 * ContinuumElectrostatics.StateVector.pyx forwards to.  Every wrapper returns the Status the reference set
 * (0 = Status_Continue) and hands values back through out-parameters. ---- */
int ref_vector_nsites (void *handle) { return ((RefModel *) handle)->vector->nsites; }
void ref_vector_reset (void *handle) { StateVector_Reset (((RefModel *) handle)->vector); }
void ref_vector_reset_to_maximum (void *handle) { StateVector_ResetToMaximum (((RefModel *) handle)->vector); }
void ref_vector_reset_substate (void *handle) { StateVector_ResetSubstate (((RefModel *) handle)->vector); }
int ref_vector_increment (void *handle) { return StateVector_Increment (((RefModel *) handle)->vector) ? 1 : 0; }
int ref_vector_increment_substate (void *handle) { return StateVector_IncrementSubstate (((RefModel *) handle)->vector) ? 1 : 0; }
int ref_vector_get_item (void *handle, int site, int *value) {
    Status status = Status_Continue;
    *value = StateVector_GetItem (((RefModel *) handle)->vector, site, &status);
    return (int) status;
}
int ref_vector_set_item (void *handle, int site, int local) {
    Status status = Status_Continue;
    StateVector_SetItem (((RefModel *) handle)->vector, site, local, &status);
    return (int) status;
}
int ref_vector_get_actual_item (void *handle, int site, int *value) {
    Status status = Status_Continue;
    *value = StateVector_GetActualItem (((RefModel *) handle)->vector, site, &status);
    return (int) status;
}
int ref_vector_set_actual_item (void *handle, int site, int global) {
    Status status = Status_Continue;
    StateVector_SetActualItem (((RefModel *) handle)->vector, site, global, &status);
    return (int) status;
}
int ref_vector_is_substate (void *handle, int site, int *flag) {
    Status status = Status_Continue;
    *flag = StateVector_IsSubstate (((RefModel *) handle)->vector, site, &status) ? 1 : 0;
    return (int) status;
}
int ref_vector_allocate_substate (void *handle, int nssites) {
    Status status = Status_Continue;
    StateVector_AllocateSubstate (((RefModel *) handle)->vector, nssites, &status);
    return (int) status;
}
int ref_vector_set_substate_item (void *handle, int selected_site, int index) {
    Status status = Status_Continue;
    StateVector_SetSubstateItem (((RefModel *) handle)->vector, selected_site, index, &status);
    return (int) status;
}
int ref_vector_get_substate_item (void *handle, int index, int *site) {
    Status status = Status_Continue;
    *site = StateVector_GetSubstateItem (((RefModel *) handle)->vector, index, &status);
    return (int) status;
}
void ref_vector_state (void *handle, int *out) {
    StateVector *vector = ((RefModel *) handle)->vector;
    int i;
    for (i = 0; i < vector->nsites; i++) out[i] = vector->sites[i].indexActive;
}
