// pce_state.cu -- the state vector: one active instance per titratable site.  Host code (no kernel): the state is a
// few integers that reach the GPU as int32 rows of global instance indices (pce_energy, pce_mc_run_chains), so it lives
// on the host as three parallel int32 arrays -- pce_state_active() IS such a row, no conversion on the way to a kernel.
//
// Replaces extensions/csource/StateVector.c behind the extern list StateVector.h:53-84 (the Cython class
// ContinuumElectrostatics.StateVector.pyx binds exactly these).  The reference keeps an array of structs
// {isSubstate, indexSite, indexActive, indexFirst, indexLast} and an array of POINTERS into it for the substate
// (StateVector.h:26-50); here the substate is a list of site indices.  The pair list the reference parks in the
// vector (StateVector.h:38-50, filled by MCModelDefault_FindPairs) belongs to pce_mc in this library.
#include <limits>

#include "pce_common.cuh"

using pce::fail;

struct pce_state {
    std::vector<int32_t> first, last, active;      // [nsites] global instance indices
    std::vector<uint8_t> in_substate;              // [nsites]
    std::vector<int32_t> substate;                 // [nssites] site indices, -1 = slot not set yet
    bool substate_allocated = false;
};

namespace {

int index_error (int index) { return fail (PCE_ERR_INDEX, "Index (" + std::to_string (index) + ") out of range."); }
int value_error (int value) { return fail (PCE_ERR_VALUE, "Invalid value (" + std::to_string (value) + ")."); }

inline bool site_ok (const pce_state *s, int site) { return site >= 0 && site < (int) s->active.size (); }

// one odometer step over the listed sites (all sites when `sites` is null): the first site that is not at its last
// instance advances, every site before it wraps to its first instance (StateVector.c:370-384 / :389-404)
bool odometer_step (pce_state *s, const int32_t *sites, int count) {
    for (int k = 0; k < count; k++) {
        const int i = sites ? sites[k] : k;
        if (s->active[i] < s->last[i]) { s->active[i]++; return true; }
        s->active[i] = s->first[i];
    }
    return false;
}

}  // namespace

// ---- allocation, cloning (StateVector_Allocate / _Deallocate / _Clone / _CopyTo, StateVector.h:53-61) ----------------
extern "C" int pce_state_create (int nsites, pce_state **out) {
    if (out == nullptr) return fail (PCE_ERR_VALUE, "null output handle");
    *out = nullptr;
    if (nsites < 0) return fail (PCE_ERR_VALUE, "invalid number of sites");
    pce_state *s = new (std::nothrow) pce_state ();
    if (s == nullptr) return fail (PCE_ERR_ALLOC, "Cannot allocate state vector.");
    try {
        s->first.assign (nsites, 0); s->last.assign (nsites, 0); s->active.assign (nsites, 0); s->in_substate.assign (nsites, 0);
    } catch (const std::bad_alloc &) {
        delete s;
        return fail (PCE_ERR_ALLOC, "Cannot allocate state vector.");
    }
    *out = s;
    return PCE_OK;
}

extern "C" int pce_state_create_for_model (const pce_model *m, pce_state **out) {
    // what the constructor of the Cython class does: Allocate, then SetSite with every site's range (StateVector.pyx:80-103)
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    int status = m->sites_ok ();
    if (status != PCE_OK) return status;
    status = pce_state_create (m->nsites, out);
    if (status != PCE_OK) return status;
    for (int i = 0; i < m->nsites; i++) { (*out)->first[i] = m->first[i]; (*out)->last[i] = m->last[i]; (*out)->active[i] = m->first[i]; }
    return PCE_OK;
}

extern "C" void pce_state_destroy (pce_state *s) { delete s; }

extern "C" int pce_state_copy_to (const pce_state *s, pce_state *other) {
    // sites only, like the reference's memcpy of the TitrSite array (the substate list is not copied, StateVector.c:114-116);
    // the reference does not check that the two vectors have the same length
    if (s == nullptr || other == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (s->active.size () != other->active.size ()) return fail (PCE_ERR_SHAPE, "state vectors of different lengths");
    other->first = s->first; other->last = s->last; other->active = s->active; other->in_substate = s->in_substate;
    return PCE_OK;
}

extern "C" int pce_state_clone (const pce_state *s, pce_state **out) {
    if (out == nullptr) return fail (PCE_ERR_VALUE, "null output handle");
    *out = nullptr;
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    int status = pce_state_create ((int) s->active.size (), out);
    if (status != PCE_OK) return status;
    return pce_state_copy_to (s, *out);
}

extern "C" int pce_state_nsites (const pce_state *s) { return s ? (int) s->active.size () : -1; }
extern "C" int pce_state_nsubstate (const pce_state *s) { return s ? (int) s->substate.size () : -1; }

// ---- all items at once (StateVector.h:63-67) ---------------------------------------------------------------------------
extern "C" int pce_state_reset (pce_state *s) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    s->active = s->first;
    return PCE_OK;
}

extern "C" int pce_state_reset_to_maximum (pce_state *s) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    s->active = s->last;
    return PCE_OK;
}

extern "C" int pce_state_reset_substate (pce_state *s) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    for (int32_t i : s->substate)
        if (i >= 0) s->active[i] = s->first[i];
    return PCE_OK;
}

extern "C" int pce_state_randomize (pce_state *s, unsigned long long seed) {
    // StateVector_Randomize (StateVector.c:163-170): instance = word % ninstances + first, one random word per site.  The
    // words come from Philox4x32-10 blocks (counter = site / 4, key = seed) instead of the reference's Mersenne Twister --
    // the same generator the Monte Carlo kernels use for their initial states.
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    pce::Philox4 block = { 0, 0, 0, 0 };
    for (int i = 0; i < (int) s->active.size (); i++) {
        if ((i & 3) == 0) block = pce::philox4x32_10 ((uint32_t) (i >> 2), 0u, 0u, 0u, (uint32_t) seed, (uint32_t) (seed >> 32));
        const uint32_t word = (i & 3) == 0 ? block.x : (i & 3) == 1 ? block.y : (i & 3) == 2 ? block.z : block.w;
        s->active[i] = s->first[i] + (int32_t) (word % (uint32_t) (s->last[i] - s->first[i] + 1));
    }
    return PCE_OK;
}

// ---- items (StateVector.h:69-79) ---------------------------------------------------------------------------------------
extern "C" int pce_state_set_site (pce_state *s, int site, int first, int last) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (last < first) return fail (PCE_ERR_VALUE, "invalid instance range");      // (the reference stores any range)
    s->first[site] = first; s->last[site] = last; s->active[site] = first; s->in_substate[site] = 0;
    return PCE_OK;
}

extern "C" int pce_state_get_site (const pce_state *s, int site, int *first, int *last) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (first) *first = s->first[site];
    if (last) *last = s->last[site];
    return PCE_OK;
}

extern "C" int pce_state_is_substate (const pce_state *s, int site, int *flag) {
    if (flag) *flag = 0;
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (flag) *flag = s->in_substate[site];
    return PCE_OK;
}

extern "C" int pce_state_get_item (const pce_state *s, int site, int *local) {
    if (local) *local = -1;                                // the reference returns -1 on error (StateVector.c:254-257)
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (local) *local = s->active[site] - s->first[site];
    return PCE_OK;
}

extern "C" int pce_state_set_item (pce_state *s, int site, int local) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    const long long global = (long long) local + s->first[site];
    if (global < s->first[site] || global > s->last[site]) return value_error (local);
    s->active[site] = (int32_t) global;
    return PCE_OK;
}

extern "C" int pce_state_get_actual_item (const pce_state *s, int site, int *global) {
    if (global) *global = -1;
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (global) *global = s->active[site];
    return PCE_OK;
}

extern "C" int pce_state_set_actual_item (pce_state *s, int site, int global) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (!site_ok (s, site)) return index_error (site);
    if (global < s->first[site] || global > s->last[site]) return value_error (global);
    s->active[site] = global;
    return PCE_OK;
}

// ---- substate (StateVector_AllocateSubstate :54, _GetSubstateItem / _SetSubstateItem :78-79) ---------------------------
extern "C" int pce_state_allocate_substate (pce_state *s, int nssites) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    // a second allocation is refused, like the reference (StateVector.c:44-47)
    if (s->substate_allocated || nssites < 0) return fail (PCE_ERR_ALLOC, "Memory allocation failure.");
    s->substate.assign (nssites, -1);
    s->substate_allocated = true;
    return PCE_OK;
}

extern "C" int pce_state_get_substate_item (const pce_state *s, int index, int *site) {
    if (site) *site = -1;
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (index < 0 || index >= (int) s->substate.size ()) return index_error (index);
    if (site) *site = s->substate[index];
    return PCE_OK;
}

extern "C" int pce_state_set_substate_item (pce_state *s, int selected_site, int index) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (index < 0 || index >= (int) s->substate.size ()) return index_error (index);
    if (!site_ok (s, selected_site)) return value_error (selected_site);
    s->in_substate[selected_site] = 1;
    s->substate[index] = selected_site;
    return PCE_OK;
}

// ---- incrementation (StateVector.h:82-83) ------------------------------------------------------------------------------
extern "C" int pce_state_increment (pce_state *s, int *more) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    const bool advanced = odometer_step (s, nullptr, (int) s->active.size ());
    if (more) *more = advanced ? 1 : 0;
    return PCE_OK;
}

extern "C" int pce_state_increment_substate (pce_state *s, int *more) {
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    for (int32_t i : s->substate)      // (the reference would dereference the NULL pointer of an unset slot)
        if (i < 0) return fail (PCE_ERR_STATE, "substate site has not been set");
    const bool advanced = odometer_step (s, s->substate.data (), (int) s->substate.size ());
    if (more) *more = advanced ? 1 : 0;
    return PCE_OK;
}

// ---- bulk forms (no reference counterpart): rows for the kernels --------------------------------------------------------
extern "C" const int32_t *pce_state_active (const pce_state *s) { return s ? s->active.data () : nullptr; }

extern "C" int pce_state_get_active (const pce_state *s, int32_t *out) {
    if (s == nullptr || out == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    std::copy (s->active.begin (), s->active.end (), out);
    return PCE_OK;
}

extern "C" int pce_state_set_active (pce_state *s, const int32_t *in) {
    if (s == nullptr || in == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    for (size_t i = 0; i < s->active.size (); i++)
        if (in[i] < s->first[i] || in[i] > s->last[i]) return value_error (in[i]);
    std::copy (in, in + s->active.size (), s->active.begin ());
    return PCE_OK;
}

extern "C" int pce_state_set_index (pce_state *s, long long index) {
    // mixed-radix decode, site 0 the fastest digit: the state reached by `index` increments from the all-first state
    if (s == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    if (index < 0) return fail (PCE_ERR_VALUE, "negative state index");
    for (size_t i = 0; i < s->active.size (); i++) {
        const long long radix = (long long) s->last[i] - s->first[i] + 1;
        s->active[i] = s->first[i] + (int32_t) (index % radix);
        index /= radix;
    }
    if (index != 0) { s->active = s->first; return fail (PCE_ERR_INDEX, "state index beyond the number of states"); }
    return PCE_OK;
}

extern "C" int pce_state_get_index (const pce_state *s, long long *index) {
    if (s == nullptr || index == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    long long value = 0, weight = 1;
    bool overflow = false;
    for (size_t i = 0; i < s->active.size (); i++) {
        const long long radix = (long long) s->last[i] - s->first[i] + 1, digit = s->active[i] - s->first[i];
        if (digit != 0) {
            if (overflow || digit > (std::numeric_limits<long long>::max () - value) / weight) return fail (PCE_ERR_LIMIT, "state index exceeds 63 bits");
            value += digit * weight;
        }
        if (weight > std::numeric_limits<long long>::max () / radix) overflow = true; else weight *= radix;
    }
    *index = value;
    return PCE_OK;
}

extern "C" int pce_state_enumerate_substate (pce_state *s, long long capacity, int32_t *states, long long *count) {
    // every state of the substate in IncrementSubstate order (the loop of Substate.py:99-108), one row of global indices
    // each: the batch kernel 1 evaluates in one launch.  With states == NULL only the count is returned.
    if (s == nullptr || count == nullptr) return fail (PCE_ERR_VALUE, "null state vector");
    for (int32_t i : s->substate)
        if (i < 0) return fail (PCE_ERR_STATE, "substate site has not been set");
    // the loop itself decides when the substate is exhausted (exactly the reference's do-while, whatever the site list)
    pce_state_reset_substate (s);
    const size_t n = s->active.size ();
    long long rows = 0;
    bool more = true;
    while (more) {
        if (states != nullptr) {
            if (rows >= capacity) { pce_state_reset_substate (s); return fail (PCE_ERR_SHAPE, "substate buffer too small"); }
            std::copy (s->active.begin (), s->active.end (), states + (size_t) rows * n);
        }
        rows++;
        more = odometer_step (s, s->substate.data (), (int) s->substate.size ());
    }
    *count = rows;
    return PCE_OK;        // the last step wrapped: the substate is back at its first state
}
