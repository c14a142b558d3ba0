#ifndef _SHIM_RNG
#define _SHIM_RNG
#include "Real.h"
#include "Cardinal.h"
/* MT19937 (Matsumoto & Nishimura 1998), the generator type the reference asks pCore for
 * (MCModelDefault.c:28).  NextReal = u32 / 2^32 in [0,1) is an assumption about pCore; the stream is
 * not pinned by any fixture (the reference seeds from time(), MCModelDefault.c:32-34). */
typedef enum { RandomNumberGeneratorType_MersenneTwister = 0 } RandomNumberGeneratorType;
typedef struct {
    unsigned int mt[624];
    int          mti;
} RandomNumberGenerator;
extern RandomNumberGenerator *RandomNumberGenerator_Allocate   (const RandomNumberGeneratorType type);
extern void                   RandomNumberGenerator_Deallocate (RandomNumberGenerator **self);
extern void                   RandomNumberGenerator_SetSeed    (RandomNumberGenerator *self, const Cardinal seed);
extern Cardinal               RandomNumberGenerator_NextCardinal (const RandomNumberGenerator *self);
extern Real                   RandomNumberGenerator_NextReal     (const RandomNumberGenerator *self);
#endif
