// pce_common.cuh -- shared host/device definitions of the B200 microstate engine.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pcetk_b200.h"

// Reference constants, EnergyModel.h:31-32 / MCModelDefault.h:39
#define PCE_R 0.001987165392
#define PCE_LN10 2.302585092994
#define PCE_TOO_SMALL (-500.0)

namespace pce {

void set_error (const std::string &message);
int  fail (int status, const std::string &message);
void count_launch (int n = 1);

#define PCE_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t err__ = (call);                                                                      \
        if (err__ != cudaSuccess) {                                                                      \
            return pce::fail (PCE_ERR_CUDA, std::string (#call) + ": " + cudaGetErrorString (err__));    \
        }                                                                                                \
    } while (0)

// mu(pH) = -R*T*ln10*pH with the reference's evaluation order (EnergyModel.c:223): ((-R*T)*ln10)*pH
inline double mu_of (double temperature, double ph) { return -PCE_R * temperature * PCE_LN10 * ph; }

template <typename T>
struct DeviceBuffer {
    T     *ptr   = nullptr;
    size_t count = 0;
    DeviceBuffer () = default;
    DeviceBuffer (const DeviceBuffer &) = delete;
    DeviceBuffer &operator= (const DeviceBuffer &) = delete;
    ~DeviceBuffer () { release (); }           // error returns between resize and release do not leak
    cudaError_t resize (size_t n) {
        if (n <= count && ptr != nullptr) return cudaSuccess;
        release ();
        cudaError_t err = cudaMalloc ((void **) &ptr, (n > 0 ? n : 1) * sizeof (T));
        if (err == cudaSuccess) count = n; else ptr = nullptr;
        return err;
    }
    void release () {
        if (ptr != nullptr) cudaFree (ptr);
        ptr = nullptr; count = 0;
    }
};

}  // namespace pce

struct pce_enum_cache;     // plan and tables of the last enumeration (pce_analytic.cu)

// The model: host copy (authoritative for setters/getters) + device mirror.
struct pce_model {
    int    nsites = 0, ninst = 0, device = 0;
    double temperature = 300.0;
    std::vector<int32_t> first, last;          // [nsites]
    std::vector<uint8_t> site_set;             // [nsites]
    std::vector<int32_t> protons;              // [ninst]
    std::vector<double>  gmodel, gintr, prob;  // [ninst]
    std::vector<double>  wraw;                 // [ninst*ninst] EnergyModel.h:46 "interactions"
    std::vector<double>  wraw_alt;             // second raw buffer: a bulk load while `sym_src` views wraw lands here and the two swap
    std::vector<double>  wsym;                 // [ninst*ninst] full square of EnergyModel.h:48 "symmetricmatrix"; allocated when first materialised
    // SymmetrizeInteractions does not touch 2 x ninst^2 doubles on the host: it records which raw buffer the symmetric matrix
    // is the average of (`sym_src`), the device forms it during the upload, and the host copy is made when a host reader
    // asks for the whole matrix (materialize_wsym).  Single elements are averaged on the fly (wsym_at).
    const double *sym_src = nullptr;
    const void   *pinned[2] = { nullptr, nullptr };     // raw buffers registered with the driver for the upload
    bool   symmetrized = false;
    bool   dirty       = true;                 // host copy newer than device copy
    int    generation  = 0;                    // bumped by every upload of NEW content (derived device tables key on it)
    std::vector<unsigned char> uploaded;       // snapshot of the last uploaded content (small models), see upload ()
    cudaStream_t stream = nullptr;

    // device mirror
    int ldw = 0;                               // row stride of d_w in doubles (multiple of 4)
    pce::DeviceBuffer<int32_t> d_first, d_nsite_inst, d_site_of_inst, d_protons;
    pce::DeviceBuffer<double>  d_gintr, d_gmodel, d_w, d_wb;   // d_wb = W / (R T), for the Monte Carlo kernels
    pce::DeviceBuffer<double>  d_raw;                          // raw interactions as uploaded (input of the device-side symmetrisation)

    // scratch of the enumeration kernel (grow-only)
    pce::DeviceBuffer<double>  s_gt, s_T, s_Mtab, s_Ftab, s_seg, s_xs;
    pce::DeviceBuffer<int32_t> s_ms;
    pce::DeviceBuffer<unsigned> s_cnt;         // block counters of the reduction (zero between launches)

    pce_enum_cache *enum_cache = nullptr;
    void (*enum_cache_free) (pce_enum_cache *) = nullptr;

    int  sites_ok () const;                    // PCE_OK when every site is set, contiguous and ascending
    int  upload ();
    void materialize_wsym ();                  // host copy of the symmetric matrix, if it is still a view of a raw buffer
    const double *host_wsym () { materialize_wsym (); return wsym.data (); }
    double wsym_at (size_t a, size_t b) const {
        const size_t n = (size_t) ninst;
        if (sym_src != nullptr) return 0.5 * (sym_src[a * n + b] + sym_src[b * n + a]);      // the same rounding as the full pass
        return wsym.empty () ? 0.0 : wsym[a * n + b];
    }
};

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace pce {

struct Philox4 { uint32_t x, y, z, w; };

// Philox4x32-10 (Salmon et al., SC'11); bit-identical to oracle_philox4x32_10.  The two 32x32->64 products per
// round compile to one IMAD.WIDE.U32 each.
__host__ __device__ __forceinline__ Philox4 philox4x32_10 (uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long) 0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long) 0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t) p1; c3 = (uint32_t) p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 out = { c0, c1, c2, c3 };
    return out;
}

__device__ __forceinline__ double warp_sum (double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync (0xffffffffu, v, o);
    return v;
}

}  // namespace pce
#endif
