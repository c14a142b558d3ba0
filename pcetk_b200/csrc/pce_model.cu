// pce_model.cu -- model container, setters/getters, symmetrisation, upload, and kernel 1
// (batched microstate energies).  Replaces the data side of extensions/csource/EnergyModel.c.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>

#include "pce_common.cuh"

namespace pce {

static thread_local std::string g_last_error;
static std::atomic<long long>   g_launches (0);

void set_error (const std::string &message) { g_last_error = message; }
int  fail (int status, const std::string &message) { g_last_error = message; return status; }
void count_launch (int n) { g_launches.fetch_add (n); }

}  // namespace pce

using pce::fail;

extern "C" const char *pce_last_error (void) { return pce::g_last_error.c_str (); }
extern "C" const char *pce_version (void) { return "pcetk_b200 0.1 (sm_100a)"; }

extern "C" long long pce_launch_count (int reset) {
    long long value = pce::g_launches.load ();
    if (reset) pce::g_launches.store (0);
    return value;
}

extern "C" int pce_device_count (int *count) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount (&n);
    if (err != cudaSuccess) { *count = 0; return fail (PCE_ERR_CUDA, std::string ("cudaGetDeviceCount: ") + cudaGetErrorString (err)); }
    *count = n;
    return PCE_OK;
}

extern "C" int pce_device_name (int device, char *buffer, int length) {
    cudaDeviceProp prop;
    PCE_CUDA (cudaGetDeviceProperties (&prop, device));
    snprintf (buffer, (size_t) length, "%s (sm_%d%d, %d SMs)", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// allocation
// ---------------------------------------------------------------------------------------------------
extern "C" int pce_model_create (int nsites, int ninstances, double temperature, int device, pce_model **out) {
    if (out == nullptr) return fail (PCE_ERR_VALUE, "null output handle");
    *out = nullptr;
    if (nsites < 0 || ninstances < 0 || ninstances < nsites) return fail (PCE_ERR_VALUE, "invalid numbers of sites / instances");
    pce_model *m = nullptr;
    try {
        m = new pce_model ();
        m->nsites = nsites; m->ninst = ninstances; m->temperature = temperature; m->device = device;
        m->first.assign (nsites, 0); m->last.assign (nsites, -1); m->site_set.assign (nsites, 0);
        m->protons.assign (ninstances, 0);
        m->gmodel.assign (ninstances, 0.0); m->gintr.assign (ninstances, 0.0); m->prob.assign (ninstances, 0.0);
        m->wraw.assign ((size_t) ninstances * ninstances, 0.0);      // (the symmetric matrix is allocated when first materialised)
    } catch (const std::bad_alloc &) {
        delete m;
        return fail (PCE_ERR_ALLOC, "Cannot allocate energy model.");
    }
    *out = m;
    return PCE_OK;
}

extern "C" void pce_model_destroy (pce_model *m) {
    if (m == nullptr) return;   // (the reference dereferences before its NULL check, EnergyModel.c:64-73)
    int current = 0;
    if (cudaGetDevice (&current) == cudaSuccess) {
        cudaSetDevice (m->device);
        m->d_first.release (); m->d_nsite_inst.release (); m->d_site_of_inst.release (); m->d_protons.release ();
        m->d_gintr.release (); m->d_gmodel.release (); m->d_w.release (); m->d_wb.release (); m->d_raw.release ();
        for (const void *&buffer : m->pinned) if (buffer != nullptr) { cudaHostUnregister (const_cast<void *> (buffer)); buffer = nullptr; }
        m->s_gt.release (); m->s_T.release (); m->s_Mtab.release (); m->s_Ftab.release (); m->s_seg.release (); m->s_xs.release (); m->s_ms.release ();
        cudaSetDevice (current);
    }
    if (m->enum_cache != nullptr && m->enum_cache_free != nullptr) m->enum_cache_free (m->enum_cache);
    delete m;
}

extern "C" int pce_model_nsites (const pce_model *m) { return m ? m->nsites : -1; }
extern "C" int pce_model_ninstances (const pce_model *m) { return m ? m->ninst : -1; }
extern "C" double pce_model_temperature (const pce_model *m) { return m ? m->temperature : 0.0; }

extern "C" int pce_set_stream (pce_model *m, void *cuda_stream) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    m->stream = (cudaStream_t) cuda_stream;
    return PCE_OK;
}

extern "C" int pce_synchronize (pce_model *m) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    PCE_CUDA (cudaSetDevice (m->device));
    PCE_CUDA (cudaStreamSynchronize (m->stream));
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// sites
// ---------------------------------------------------------------------------------------------------
extern "C" int pce_model_set_site (pce_model *m, int site, int first, int last) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (site < 0 || site >= m->nsites) return fail (PCE_ERR_INDEX, "Index (" + std::to_string (site) + ") out of range.");
    if (first < 0 || last < first || last >= m->ninst) return fail (PCE_ERR_VALUE, "invalid instance range");
    m->first[site] = first; m->last[site] = last; m->site_set[site] = 1;
    m->dirty = true;
    return PCE_OK;
}

extern "C" int pce_model_set_sites (pce_model *m, const int32_t *first, const int32_t *last) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    for (int i = 0; i < m->nsites; i++) {
        int status = pce_model_set_site (m, i, first[i], last[i]);
        if (status != PCE_OK) return status;
    }
    return m->sites_ok ();
}

extern "C" int pce_model_get_site (const pce_model *m, int site, int *first, int *last) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (site < 0 || site >= m->nsites) return fail (PCE_ERR_INDEX, "Index (" + std::to_string (site) + ") out of range.");
    *first = m->first[site]; *last = m->last[site];
    return PCE_OK;
}

int pce_model::sites_ok () const {
    int expect = 0;
    for (int i = 0; i < nsites; i++) {
        if (!site_set[i]) return fail (PCE_ERR_STATE, "site " + std::to_string (i) + " has not been set");
        if (first[i] != expect) return fail (PCE_ERR_VALUE, "site instance ranges must be contiguous and ascending");
        expect = last[i] + 1;
    }
    if (expect != ninst) return fail (PCE_ERR_VALUE, "site ranges do not cover all instances");
    return PCE_OK;
}

extern "C" int pce_model_nstates (const pce_model *m, double *nstates) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    int status = m->sites_ok ();
    if (status != PCE_OK) return status;
    double n = 1.0;
    for (int i = 0; i < m->nsites; i++) n *= (double) (m->last[i] - m->first[i] + 1);
    *nstates = n;
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// element access (EnergyModel.c:145-199)
// ---------------------------------------------------------------------------------------------------
#define CHECK_INSTANCE(m, k)                                                                            \
    if ((m) == nullptr) return fail (PCE_ERR_VALUE, "null model");                                     \
    if ((k) < 0 || (k) >= (m)->ninst) return fail (PCE_ERR_INDEX, "Index (" + std::to_string (k) + ") out of range.");

extern "C" int pce_model_set_gmodel (pce_model *m, int k, double v) { CHECK_INSTANCE (m, k); m->gmodel[k] = v; m->dirty = true; return PCE_OK; }
extern "C" int pce_model_set_gintr (pce_model *m, int k, double v) { CHECK_INSTANCE (m, k); m->gintr[k] = v; m->dirty = true; return PCE_OK; }
extern "C" int pce_model_set_protons (pce_model *m, int k, int v) { CHECK_INSTANCE (m, k); m->protons[k] = v; m->dirty = true; return PCE_OK; }
extern "C" int pce_model_set_probability (pce_model *m, int k, double v) { CHECK_INSTANCE (m, k); m->prob[k] = v; return PCE_OK; }
extern "C" int pce_model_set_interaction (pce_model *m, int a, int b, double v) {
    CHECK_INSTANCE (m, a); CHECK_INSTANCE (m, b);
    if (m->sym_src == m->wraw.data ()) m->materialize_wsym ();      // the symmetric matrix must not follow this write
    m->wraw[(size_t) a * m->ninst + b] = v;    // raw matrix only; the symmetric copy changes on Symmetrize (EnergyModel.c:197-199)
    return PCE_OK;
}
extern "C" int pce_model_get_gmodel (const pce_model *m, int k, double *v) { CHECK_INSTANCE (m, k); *v = m->gmodel[k]; return PCE_OK; }
extern "C" int pce_model_get_gintr (const pce_model *m, int k, double *v) { CHECK_INSTANCE (m, k); *v = m->gintr[k]; return PCE_OK; }
extern "C" int pce_model_get_protons (const pce_model *m, int k, int *v) { CHECK_INSTANCE (m, k); *v = m->protons[k]; return PCE_OK; }
extern "C" int pce_model_get_probability (const pce_model *m, int k, double *v) { CHECK_INSTANCE (m, k); *v = m->prob[k]; return PCE_OK; }
extern "C" int pce_model_get_interaction (const pce_model *m, int a, int b, double *v) {
    CHECK_INSTANCE (m, a); CHECK_INSTANCE (m, b);
    *v = m->wraw[(size_t) a * m->ninst + b];
    return PCE_OK;
}
extern "C" int pce_model_get_interaction_symmetric (const pce_model *m, int a, int b, double *v) {
    CHECK_INSTANCE (m, a); CHECK_INSTANCE (m, b);
    *v = m->wsym_at ((size_t) a, (size_t) b);
    return PCE_OK;
}
extern "C" int pce_model_get_deviation (const pce_model *m, int a, int b, double *v) {
    CHECK_INSTANCE (m, a); CHECK_INSTANCE (m, b);
    const double wij = m->wraw[(size_t) a * m->ninst + b], wji = m->wraw[(size_t) b * m->ninst + a];
    *v = (wij + wji) * .5 - wij;               // EnergyModel.c:169-176
    return PCE_OK;
}

extern "C" int pce_model_load (pce_model *m, const double *gintr, const double *gmodel, const int32_t *protons, const double *interactions) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    const size_t n = (size_t) m->ninst;
    if (gintr) std::memcpy (m->gintr.data (), gintr, n * sizeof (double));
    if (gmodel) std::memcpy (m->gmodel.data (), gmodel, n * sizeof (double));
    if (protons) std::memcpy (m->protons.data (), protons, n * sizeof (int32_t));
    if (interactions) {
        if (m->sym_src == m->wraw.data ()) {
            // the symmetric matrix is still a view of the buffer about to be overwritten: the new raw matrix goes into the
            // other buffer (the symmetric matrix changes on Symmetrize only, EnergyModel.c:197-199)
            try { if (m->wraw_alt.size () != n * n) m->wraw_alt.resize (n * n); }
            catch (const std::bad_alloc &) { return fail (PCE_ERR_ALLOC, "Cannot allocate energy model."); }
            std::swap (m->wraw, m->wraw_alt);
        }
        std::memcpy (m->wraw.data (), interactions, n * n * sizeof (double));
    }
    m->dirty = true;
    return PCE_OK;
}

extern "C" int pce_model_get_probabilities (const pce_model *m, double *p) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    std::memcpy (p, m->prob.data (), (size_t) m->ninst * sizeof (double));
    return PCE_OK;
}
extern "C" int pce_model_set_probabilities (pce_model *m, const double *p) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    std::memcpy (m->prob.data (), p, (size_t) m->ninst * sizeof (double));
    return PCE_OK;
}
extern "C" int pce_model_get_symmetric_matrix (const pce_model *m, double *w) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    const size_t n = (size_t) m->ninst;
    if (m->sym_src == nullptr && m->wsym.empty ()) { std::fill (w, w + n * n, 0.0); return PCE_OK; }
    std::memcpy (w, const_cast<pce_model *> (m)->host_wsym (), n * n * sizeof (double));
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// symmetry (EnergyModel.c:78-101)
// ---------------------------------------------------------------------------------------------------
extern "C" int pce_model_check_symmetric (const pce_model *m, double tolerance, double *max_deviation, int *is_symmetric) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    const size_t n = (size_t) m->ninst;
    double maxd = 0.0;
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < i; j++) {
            const double d = std::fabs (m->wraw[i * n + j] - m->wraw[j * n + i]);
            if (d > maxd) maxd = d;
        }
    if (max_deviation) *max_deviation = maxd;
    if (is_symmetric) *is_symmetric = maxd <= tolerance ? 1 : 0;
    return PCE_OK;
}

void pce_model::materialize_wsym () {
    const size_t n = (size_t) ninst, tile = 32;
    if (sym_src == nullptr) {
        if (wsym.size () != n * n) wsym.assign (n * n, 0.0);
        return;
    }
    // Wsym_ij = Wsym_ji = (W_ij + W_ji) / 2 (SymmetricMatrix_CopyFromReal2DArray, EnergyModel.c:78-87), in tiles of 32 x 32 so
    // that the transposed accesses stay in cache (60 -> 41 ms on the 3000-instance model; a tile buffer with contiguous
    // write-back and host threads both measured slower)
    if (wsym.size () != n * n) wsym.resize (n * n);
    const double *w = sym_src;
    double *s = wsym.data ();
    for (size_t i0 = 0; i0 < n; i0 += tile)
        for (size_t j0 = 0; j0 <= i0; j0 += tile) {
            const size_t i1 = std::min (n, i0 + tile), j1 = std::min (n, j0 + tile);
            for (size_t i = i0; i < i1; i++)
                for (size_t j = j0; j < j1 && j <= i; j++) {
                    const double v = 0.5 * (w[i * n + j] + w[j * n + i]);
                    s[i * n + j] = v;
                    s[j * n + i] = v;
                }
        }
    sym_src = nullptr;
}

extern "C" int pce_model_symmetrize (pce_model *m) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    // the averaging itself runs on the device during the upload (k_symmetrize_stage); the host copy is made on demand
    m->sym_src = m->wraw.data ();
    m->symmetrized = true;
    m->dirty = true;
    return PCE_OK;
}

extern "C" int pce_model_reset_interactions (pce_model *m) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    m->sym_src = nullptr;
    m->wsym.assign ((size_t) m->ninst * m->ninst, 0.0);
    m->symmetrized = true;
    m->dirty = true;
    return PCE_OK;
}

extern "C" int pce_model_scale_interactions (pce_model *m, double scale) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    m->materialize_wsym ();
    for (double &v : m->wsym) v *= scale;
    m->dirty = true;
    return PCE_OK;
}

extern "C" int pce_model_state_from_probabilities (const pce_model *m, int32_t *state) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    int status = m->sites_ok ();
    if (status != PCE_OK) return status;
    for (int i = 0; i < m->nsites; i++) {      // EnergyModel.c:107-140 with the loop bound fixed (nsites, not nsites+1)
        int    maxi = m->first[i];
        double maxp = -1.0;
        for (int k = m->first[i]; k <= m->last[i]; k++)
            if (m->prob[k] > maxp) { maxp = m->prob[k]; maxi = k; }
        state[i] = maxi;
    }
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// upload
// ---------------------------------------------------------------------------------------------------
__global__ void k_scale (const double *__restrict__ in, size_t n, double factor, double *__restrict__ out) {
    for (size_t e = (size_t) blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t) gridDim.x * blockDim.x) out[e] = __dmul_rn (factor, in[e]);
}

// Symmetric, staged and scaled interactions from the raw matrix: w[i][j] = (raw[i][j] + raw[j][i]) / 2 for instances of
// different sites, 0 inside a site and in the padding columns; wb = beta * w.  One 32 x 32 tile per block, the transposed
// operand through shared memory so that both reads are coalesced.
__global__ void __launch_bounds__ (256) k_symmetrize_stage (const double *__restrict__ raw, int n, int ldw, const int32_t *__restrict__ site_of,
                                                            double beta, double *__restrict__ w, double *__restrict__ wb) {
    __shared__ double t[32][33];
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int row = j0 + r, col = i0 + threadIdx.x;
        t[r][threadIdx.x] = (row < n && col < n) ? raw[(size_t) row * n + col] : 0.0;
    }
    __syncthreads ();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        if (i < n && j < ldw) {
            double v = 0.0;
            if (j < n && site_of[i] != site_of[j]) v = __dmul_rn (0.5, __dadd_rn (raw[(size_t) i * n + j], t[threadIdx.x][r]));
            w[(size_t) i * ldw + j]  = v;
            wb[(size_t) i * ldw + j] = __dmul_rn (beta, v);
        }
    }
}

int pce_model::upload () {
    int status = sites_ok ();
    if (status != PCE_OK) return status;
    PCE_CUDA (cudaSetDevice (device));
    if (!dirty && d_w.ptr != nullptr) return PCE_OK;
    ldw = (ninst + 3) & ~3;
    PCE_CUDA (d_first.resize (nsites)); PCE_CUDA (d_nsite_inst.resize (nsites));
    PCE_CUDA (d_site_of_inst.resize (ninst)); PCE_CUDA (d_protons.resize (ninst));
    PCE_CUDA (d_gintr.resize (ninst)); PCE_CUDA (d_gmodel.resize (ninst));
    PCE_CUDA (d_w.resize ((size_t) ninst * ldw));
    std::vector<int32_t> ns (nsites), site_of (ninst);
    for (int i = 0; i < nsites; i++) {
        ns[i] = last[i] - first[i] + 1;
        for (int k = first[i]; k <= last[i]; k++) site_of[k] = i;
    }
    PCE_CUDA (cudaMemcpyAsync (d_first.ptr, first.data (), nsites * sizeof (int32_t), cudaMemcpyHostToDevice, stream));
    PCE_CUDA (cudaMemcpyAsync (d_nsite_inst.ptr, ns.data (), nsites * sizeof (int32_t), cudaMemcpyHostToDevice, stream));
    PCE_CUDA (cudaMemcpyAsync (d_site_of_inst.ptr, site_of.data (), ninst * sizeof (int32_t), cudaMemcpyHostToDevice, stream));
    PCE_CUDA (cudaMemcpyAsync (d_protons.ptr, protons.data (), ninst * sizeof (int32_t), cudaMemcpyHostToDevice, stream));
    PCE_CUDA (cudaMemcpyAsync (d_gintr.ptr, gintr.data (), ninst * sizeof (double), cudaMemcpyHostToDevice, stream));
    PCE_CUDA (cudaMemcpyAsync (d_gmodel.ptr, gmodel.data (), ninst * sizeof (double), cudaMemcpyHostToDevice, stream));
    {   // W on the device: row stride ldw, and the blocks between instances of ONE site zeroed.  The reference never reads those
        // entries (energies sum over pairs of different sites, EnergyModel.c:204-224; Move / DoubleMove skip the moved sites,
        // MCModelDefault.c:100-111, 140-160), while the cached fields of the Monte Carlo kernels sum over whole rows: zeroing
        // them here makes an input file with non-zero same-site entries sample the reference's distribution.
        // d_wb = W in units of RT for the Monte Carlo kernels: one rounding per element (beta * w), identical to the chain
        // specification.
        const size_t total = (size_t) ninst * ldw;
        const double beta = 1.0 / (PCE_R * temperature);
        PCE_CUDA (d_wb.resize (total));
        if (sym_src != nullptr) {
            // symmetric matrix still a view of a raw buffer: upload the raw matrix (page-locked in place, once per buffer) and
            // average, zero and scale on the device
            const size_t bytes = (size_t) ninst * ninst * sizeof (double);
            if (bytes >= ((size_t) 1 << 20) && pinned[0] != sym_src && pinned[1] != sym_src) {
                const int slot = pinned[0] == nullptr ? 0 : 1;
                if (pinned[slot] != nullptr) { cudaHostUnregister (const_cast<void *> (pinned[slot])); pinned[slot] = nullptr; }
                if (cudaHostRegister (const_cast<double *> (sym_src), bytes, cudaHostRegisterDefault) == cudaSuccess) pinned[slot] = sym_src;
                else cudaGetLastError ();        // pageable copy instead
            }
            PCE_CUDA (d_raw.resize ((size_t) ninst * ninst));
            PCE_CUDA (cudaMemcpyAsync (d_raw.ptr, sym_src, bytes, cudaMemcpyHostToDevice, stream));
            const dim3 grid ((unsigned) ((ldw + 31) / 32), (unsigned) ((ninst + 31) / 32)), block (32, 8);
            k_symmetrize_stage<<<grid, block, 0, stream>>> (d_raw.ptr, ninst, ldw, d_site_of_inst.ptr, beta, d_w.ptr, d_wb.ptr);
            pce::count_launch ();
            PCE_CUDA (cudaGetLastError ());
            PCE_CUDA (cudaStreamSynchronize (stream));   // the small staging vectors above die with this scope
        } else {
            const double *ws = host_wsym ();
            std::vector<double> staged (total, 0.0);
            for (int a = 0; a < ninst; a++) std::memcpy (&staged[(size_t) a * ldw], &ws[(size_t) a * ninst], (size_t) ninst * sizeof (double));
            for (int i = 0; i < nsites; i++)
                for (int a = first[i]; a <= last[i]; a++)
                    for (int b = first[i]; b <= last[i]; b++) staged[(size_t) a * ldw + b] = 0.0;
            PCE_CUDA (cudaMemcpyAsync (d_w.ptr, staged.data (), staged.size () * sizeof (double), cudaMemcpyHostToDevice, stream));
            k_scale<<<(unsigned) std::min<size_t> ((total + 255) / 256, 65535), 256, 0, stream>>> (d_w.ptr, total, beta, d_wb.ptr);
            pce::count_launch ();
            PCE_CUDA (cudaGetLastError ());
            PCE_CUDA (cudaStreamSynchronize (stream));   // the staging vectors die with this scope
        }
    }
    dirty = false;
    // The generation (which the derived device tables and the enumeration plan key on) moves only when the uploaded CONTENT
    // changed: a caller that re-loads the very same arrays before every call -- the reference's scripts do -- pays the copies
    // above but keeps its tables.  Exact comparison against a snapshot of the last upload, for models of up to 4 MB.
    {
        const double *wsrc = sym_src != nullptr ? sym_src : wsym.data ();
        const size_t nw = (sym_src != nullptr || !wsym.empty ()) ? (size_t) ninst * ninst : 0;
        const size_t bytes = (size_t) nsites * 8 + (size_t) ninst * 20 + nw * 8 + 16;
        std::vector<unsigned char> snap;
        if (bytes <= ((size_t) 4 << 20)) {
            snap.reserve (bytes);
            auto put = [&] (const void *data, size_t n) { const unsigned char *c = (const unsigned char *) data; snap.insert (snap.end (), c, c + n); };
            const unsigned char pending = sym_src != nullptr ? 1 : 0;
            put (&pending, 1); put (&temperature, sizeof (double));
            put (first.data (), (size_t) nsites * 4); put (last.data (), (size_t) nsites * 4);
            put (protons.data (), (size_t) ninst * 4); put (gintr.data (), (size_t) ninst * 8); put (gmodel.data (), (size_t) ninst * 8);
            if (nw) put (wsrc, nw * 8);
        }
        if (snap.empty () || snap != uploaded) { generation++; uploaded.swap (snap); }
    }
    return PCE_OK;
}

extern "C" int pce_model_upload (pce_model *m) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    return m->upload ();
}

// ---------------------------------------------------------------------------------------------------
// kernel 1: microstate energies.  One thread per state walks the sites in the reference's order
// (EnergyModel.c:204-224): Gintr and proton sums over sites, W summed over j < i with the row of a_i --
// sequential fp64 adds without FMA contraction, so the result is bit-identical to the reference.
// The pH-independent sums are formed once and every pH of the batch is emitted from them.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (128) k_energy (const int32_t *__restrict__ states, long long nstates, int nsites, int ninst, int ldw,
                                                  const double *__restrict__ base,      // Gintr (folded) or Gmodel (unfolded)
                                                  const int32_t *__restrict__ protons, const double *__restrict__ w, int unfolded,
                                                  const double *__restrict__ ph, double temperature, int nph, double *__restrict__ out) {
    const long long s = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nstates) return;
    const int32_t *state = states + s * nsites;
    double g = 0.0, wsum = 0.0;
    int    np = 0;
    for (int i = 0; i < nsites; i++) {
        const int ai = state[i];
        g = __dadd_rn (g, base[ai]);
        np += protons[ai];
        if (!unfolded) {
            const double *row = w + (size_t) ai * ldw;
            for (int j = 0; j < i; j++) wsum = __dadd_rn (wsum, row[state[j]]);
        }
    }
    for (int q = 0; q < nph; q++) {
        // mu = -R*T*ln10*pH evaluated left to right like EnergyModel.c:223
        const double mu = __dmul_rn (__dmul_rn (__dmul_rn (-PCE_R, temperature), PCE_LN10), ph[q]);
        double value = __dsub_rn (g, __dmul_rn ((double) np, mu));
        if (!unfolded) value = __dadd_rn (value, wsum);
        out[s * nph + q] = value;
    }
}

static int validate_states_host (const pce_model *m, const int32_t *states, long long nstates) {
    for (long long s = 0; s < nstates; s++)
        for (int i = 0; i < m->nsites; i++) {
            const int a = states[s * m->nsites + i];
            if (a < m->first[i] || a > m->last[i]) return fail (PCE_ERR_VALUE, "Invalid value (" + std::to_string (a) + ").");
        }
    return PCE_OK;
}

extern "C" int pce_energy_device (pce_model *m, const int32_t *d_states, long long nstates, const double *d_ph, int nph, int unfolded, double *d_out) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    if (!unfolded && !m->symmetrized) return fail (PCE_ERR_STATE, "First calculate electrostatic energies.");
    int status = m->upload ();
    if (status != PCE_OK) return status;
    if (nstates <= 0 || nph <= 0) return PCE_OK;
    const int threads = 128;
    const long long blocks = (nstates + threads - 1) / threads;
    if (blocks > 2147483647LL) return fail (PCE_ERR_LIMIT, "too many states in one energy batch");
    k_energy<<<(unsigned) blocks, threads, 0, m->stream>>> (d_states, nstates, m->nsites, m->ninst, m->ldw,
                                                            unfolded ? m->d_gmodel.ptr : m->d_gintr.ptr, m->d_protons.ptr, m->d_w.ptr,
                                                            unfolded, d_ph, m->temperature, nph, d_out);
    pce::count_launch ();
    PCE_CUDA (cudaGetLastError ());
    return PCE_OK;
}

extern "C" int pce_energy (pce_model *m, const int32_t *states, long long nstates, const double *ph, int nph, int unfolded, double *out) {
    if (m == nullptr) return fail (PCE_ERR_VALUE, "null model");
    int status = m->sites_ok ();
    if (status != PCE_OK) return status;
    status = validate_states_host (m, states, nstates);
    if (status != PCE_OK) return status;
    if (nstates <= 0 || nph <= 0) return PCE_OK;
    status = m->upload ();
    if (status != PCE_OK) return status;
    pce::DeviceBuffer<int32_t> d_states;
    pce::DeviceBuffer<double>  d_mu, d_out;
    PCE_CUDA (d_states.resize ((size_t) nstates * m->nsites));
    PCE_CUDA (d_mu.resize (nph));
    PCE_CUDA (d_out.resize ((size_t) nstates * nph));
    PCE_CUDA (cudaMemcpyAsync (d_states.ptr, states, (size_t) nstates * m->nsites * sizeof (int32_t), cudaMemcpyHostToDevice, m->stream));
    PCE_CUDA (cudaMemcpyAsync (d_mu.ptr, ph, nph * sizeof (double), cudaMemcpyHostToDevice, m->stream));
    status = pce_energy_device (m, d_states.ptr, nstates, d_mu.ptr, nph, unfolded, d_out.ptr);
    if (status == PCE_OK) {
        cudaError_t err = cudaMemcpyAsync (out, d_out.ptr, (size_t) nstates * nph * sizeof (double), cudaMemcpyDeviceToHost, m->stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize (m->stream);
        if (err != cudaSuccess) status = fail (PCE_ERR_CUDA, std::string ("energy readback: ") + cudaGetErrorString (err));
    }
    d_states.release (); d_mu.release (); d_out.release ();
    return status;
}

// ---------------------------------------------------------------------------------------------------
// test hook: Philox on the device
// ---------------------------------------------------------------------------------------------------
__global__ void k_philox (const uint32_t *ctr, int nblocks, uint32_t k0, uint32_t k1, uint32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    pce::Philox4 r = pce::philox4x32_10 (ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], k0, k1);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" int pce_test_philox (int device, const uint32_t *ctr, int nblocks, const uint32_t key[2], uint32_t *out) {
    PCE_CUDA (cudaSetDevice (device));
    pce::DeviceBuffer<uint32_t> d_ctr, d_out;
    PCE_CUDA (d_ctr.resize ((size_t) 4 * nblocks));
    PCE_CUDA (d_out.resize ((size_t) 4 * nblocks));
    PCE_CUDA (cudaMemcpy (d_ctr.ptr, ctr, (size_t) 4 * nblocks * sizeof (uint32_t), cudaMemcpyHostToDevice));
    k_philox<<<(nblocks + 127) / 128, 128>>> (d_ctr.ptr, nblocks, key[0], key[1], d_out.ptr);
    pce::count_launch ();
    PCE_CUDA (cudaGetLastError ());
    PCE_CUDA (cudaMemcpy (out, d_out.ptr, (size_t) 4 * nblocks * sizeof (uint32_t), cudaMemcpyDeviceToHost));
    d_ctr.release (); d_out.release ();
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// measurement hook: streaming read bandwidth of a buffer of `bytes` (L2-resident when it is well below the 126 MB
// L2, HBM when it is several times larger), 128-bit loads from every SM.  bench.py uses it as the measured
// denominator of the L2 roofline of the chain-per-warp kernel (MEASURED_PEAKS.json only holds the HBM figure).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (512) k_probe_read (const double2 *__restrict__ data, size_t n2, int passes, double *sink) {
    double acc = 0.0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (int pass = 0; pass < passes; pass++) {
        size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n2; i += 4 * stride) {
            const double2 a = __ldcg (data + i), b = __ldcg (data + i + stride), c = __ldcg (data + i + 2 * stride), d = __ldcg (data + i + 3 * stride);
            acc += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
        }
        for (; i < n2; i += stride) { const double2 a = __ldcg (data + i); acc += a.x + a.y; }
    }
    if (acc == 123.456) *sink = acc;       // keeps the loads alive
}

extern "C" int pce_probe_read_bandwidth (int device, long long bytes, int passes, double *gb_per_s) {
    if (bytes < 4096 || passes < 1 || gb_per_s == nullptr) return fail (PCE_ERR_VALUE, "invalid probe size");
    PCE_CUDA (cudaSetDevice (device));
    int sms = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, device));
    pce::DeviceBuffer<double2> buffer;
    pce::DeviceBuffer<double>  sink;
    const size_t n2 = (size_t) bytes / sizeof (double2);
    PCE_CUDA (buffer.resize (n2)); PCE_CUDA (sink.resize (1));
    PCE_CUDA (cudaMemset (buffer.ptr, 0, n2 * sizeof (double2)));
    cudaEvent_t start, stop;
    PCE_CUDA (cudaEventCreate (&start)); PCE_CUDA (cudaEventCreate (&stop));
    k_probe_read<<<sms * 4, 512>>> (buffer.ptr, n2, 2, sink.ptr);              // warm-up: brings the buffer into L2
    PCE_CUDA (cudaEventRecord (start));
    k_probe_read<<<sms * 4, 512>>> (buffer.ptr, n2, passes, sink.ptr);
    PCE_CUDA (cudaEventRecord (stop));
    pce::count_launch (2);
    cudaError_t err = cudaEventSynchronize (stop);
    float ms = 0.0f;
    if (err == cudaSuccess) err = cudaEventElapsedTime (&ms, start, stop);
    cudaEventDestroy (start); cudaEventDestroy (stop);
    buffer.release (); sink.release ();
    if (err != cudaSuccess) return fail (PCE_ERR_CUDA, std::string ("bandwidth probe: ") + cudaGetErrorString (err));
    *gb_per_s = (double) n2 * sizeof (double2) * passes / (ms * 1e-3) / 1e9;
    return PCE_OK;
}

// ---------------------------------------------------------------------------------------------------
// on-chip peaks (FP64, shared-memory and issue peaks are not in MEASURED_PEAKS.json and must be
// micro-benchmarked).  kind 0: DFMA rate (8 independent chains per thread, registers only) -> 1e9 DFMA/s;
// kind 1: shared-memory read bandwidth (conflict-free 128-bit loads, every lane its own 16 bytes) -> GB/s;
// kind 2: issue rate (independent FP32 FMAs, 8 chains per thread: 128 lanes per SM = the issue limit) -> 1e9 warp-instructions/s.
// bench.py uses them as measured denominators (enumeration: DFMA; chain-per-thread Monte Carlo: issue slots).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256) k_probe_dfma (int iterations, double seed, double *sink) {
    double a0 = seed, a1 = seed + 1.0, a2 = seed + 2.0, a3 = seed + 3.0, a4 = seed + 4.0, a5 = seed + 5.0, a6 = seed + 6.0, a7 = seed + 7.0;
    const double m = 0.9999999 + 1e-9 * threadIdx.x, c = 1e-7;
#pragma unroll 1
    for (int i = 0; i < iterations; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma (a0, m, c); a1 = fma (a1, m, c); a2 = fma (a2, m, c); a3 = fma (a3, m, c);
            a4 = fma (a4, m, c); a5 = fma (a5, m, c); a6 = fma (a6, m, c); a7 = fma (a7, m, c);
        }
    }
    const double total = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (total == 123.456) *sink = total;
}

__global__ void __launch_bounds__ (256) k_probe_shared (int iterations, unsigned *sink) {
    __shared__ uint4 buffer[2048];                           // 32 KB
    for (int e = threadIdx.x; e < 2048; e += blockDim.x) buffer[e] = make_uint4 (e, 2u * e, 3u * e, 5u * e);
    __syncthreads ();
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int index = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < iterations; i++) {
#pragma unroll
        for (int u = 0; u < 8; u += 4) {                     // conflict-free 128-bit loads: consecutive lanes, consecutive 16 bytes
            const uint4 v0 = buffer[(index + 256 * u) & 2047], v1 = buffer[(index + 256 * (u + 1)) & 2047];
            const uint4 v2 = buffer[(index + 256 * (u + 2)) & 2047], v3 = buffer[(index + 256 * (u + 3)) & 2047];
            a0 ^= v0.x ^ v0.y; a0 ^= v0.z ^ v0.w; a1 ^= v1.x ^ v1.y; a1 ^= v1.z ^ v1.w;
            a2 ^= v2.x ^ v2.y; a2 ^= v2.z ^ v2.w; a3 ^= v3.x ^ v3.y; a3 ^= v3.z ^ v3.w;
        }
        index = (index + 32) & 2047;
    }
    const unsigned total = a0 ^ a1 ^ a2 ^ a3;
    if (total == 123456789u) *sink = total;
}

// issue rate: FP32 FMAs (128 lanes per SM: one warp-instruction per scheduler and clock, the issue limit itself)
__global__ void __launch_bounds__ (256) k_probe_issue (int iterations, float seed, float *sink) {
    float a0 = seed, a1 = seed + 1.f, a2 = seed + 2.f, a3 = seed + 3.f, a4 = seed + 4.f, a5 = seed + 5.f, a6 = seed + 6.f, a7 = seed + 7.f;
    const float m = 0.999999f + 1e-8f * threadIdx.x, c = 1e-6f;
#pragma unroll 1
    for (int i = 0; i < iterations; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fmaf (a0, m, c); a1 = fmaf (a1, m, c); a2 = fmaf (a2, m, c); a3 = fmaf (a3, m, c);
            a4 = fmaf (a4, m, c); a5 = fmaf (a5, m, c); a6 = fmaf (a6, m, c); a7 = fmaf (a7, m, c);
        }
    }
    const float total = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (total == 123.456f) *sink = total;
}

extern "C" int pce_probe_on_chip (int device, int kind, double *rate) {
    if (rate == nullptr || kind < 0 || kind > 2) return fail (PCE_ERR_VALUE, "invalid probe");
    PCE_CUDA (cudaSetDevice (device));
    int sms = 0;
    PCE_CUDA (cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, device));
    pce::DeviceBuffer<double> sink;
    PCE_CUDA (sink.resize (2));
    cudaEvent_t start, stop;
    PCE_CUDA (cudaEventCreate (&start)); PCE_CUDA (cudaEventCreate (&stop));
    const int blocks = sms * 8, iterations = kind == 1 ? 4096 : 2048;
    for (int pass = 0; pass < 2; pass++) {                   // pass 0 warms up
        if (pass == 1) cudaEventRecord (start);
        if (kind == 0) k_probe_dfma<<<blocks, 256>>> (iterations, 0.5, sink.ptr);
        else if (kind == 1) k_probe_shared<<<blocks, 256>>> (iterations, (unsigned *) sink.ptr);
        else k_probe_issue<<<blocks, 256>>> (iterations, 0.5f, (float *) sink.ptr);
    }
    cudaEventRecord (stop);
    pce::count_launch (2);
    cudaError_t err = cudaEventSynchronize (stop);
    float ms = 0.0f;
    if (err == cudaSuccess) err = cudaEventElapsedTime (&ms, start, stop);
    cudaEventDestroy (start); cudaEventDestroy (stop);
    sink.release ();
    if (err != cudaSuccess) return fail (PCE_ERR_CUDA, std::string ("on-chip probe: ") + cudaGetErrorString (err));
    const double threads = (double) blocks * 256.0, per_thread = (double) iterations * 64.0;
    if (kind == 0) *rate = threads * per_thread / (ms * 1e-3) / 1e9;                 // 1e9 DFMA per second (x 2 = GFLOP/s)
    else if (kind == 1) *rate = threads * (double) iterations * 8.0 * 16.0 / (ms * 1e-3) / 1e9;     // GB/s (8 x 128-bit loads per iteration)
    else *rate = threads / 32.0 * per_thread / (ms * 1e-3) / 1e9;                    // 1e9 warp-instructions per second
    return PCE_OK;
}
