/* pCore shim implementations (test infrastructure only; see shim/Real.h). */
#include <math.h>
#include <stdlib.h>
#include "Real1DArray.h"
#include "Integer1DArray.h"
#include "Real2DArray.h"
#include "SymmetricMatrix.h"
#include "RandomNumberGenerator.h"
#include "Memory.h"

Real1DArray *Real1DArray_Allocate (const Integer length, Status *status) {
    Real1DArray *self = (Real1DArray *) malloc (sizeof (Real1DArray));
    if (self != NULL) {
        self->length = length;
        self->data   = (Real *) calloc ((size_t) (length > 0 ? length : 1), sizeof (Real));
        if (self->data == NULL) { free (self); self = NULL; }
    }
    if (self == NULL) Status_Set (status, Status_MemoryAllocationFailure);
    return self;
}
void Real1DArray_Deallocate (Real1DArray **self) {
    if (self != NULL && *self != NULL) { free ((*self)->data); free (*self); *self = NULL; }
}
void Real1DArray_Set (Real1DArray *self, const Real value) { Integer i; for (i = 0; i < self->length; i++) self->data[i] = value; }
void Real1DArray_Scale (Real1DArray *self, const Real value) { Integer i; for (i = 0; i < self->length; i++) self->data[i] *= value; }
void Real1DArray_AddScalar (Real1DArray *self, const Real value) { Integer i; for (i = 0; i < self->length; i++) self->data[i] += value; }
void Real1DArray_Exp (Real1DArray *self) { Integer i; for (i = 0; i < self->length; i++) self->data[i] = exp (self->data[i]); }
Real Real1DArray_Sum (const Real1DArray *self) { Integer i; Real sum = 0.0; for (i = 0; i < self->length; i++) sum += self->data[i]; return sum; }

Integer1DArray *Integer1DArray_Allocate (const Integer length, Status *status) {
    Integer1DArray *self = (Integer1DArray *) malloc (sizeof (Integer1DArray));
    if (self != NULL) {
        self->length = length;
        self->data   = (Integer *) calloc ((size_t) (length > 0 ? length : 1), sizeof (Integer));
        if (self->data == NULL) { free (self); self = NULL; }
    }
    if (self == NULL) Status_Set (status, Status_MemoryAllocationFailure);
    return self;
}
void Integer1DArray_Deallocate (Integer1DArray **self) {
    if (self != NULL && *self != NULL) { free ((*self)->data); free (*self); *self = NULL; }
}

Real2DArray *Real2DArray_Allocate (const Integer length0, const Integer length1, Status *status) {
    Real2DArray *self = (Real2DArray *) malloc (sizeof (Real2DArray));
    if (self != NULL) {
        size_t n = (size_t) length0 * (size_t) length1;
        self->length0 = length0;
        self->length1 = length1;
        self->data    = (Real *) calloc (n > 0 ? n : 1, sizeof (Real));
        if (self->data == NULL) { free (self); self = NULL; }
    }
    if (self == NULL) Status_Set (status, Status_MemoryAllocationFailure);
    return self;
}
void Real2DArray_Deallocate (Real2DArray **self) {
    if (self != NULL && *self != NULL) { free ((*self)->data); free (*self); *self = NULL; }
}
Boolean Real2DArray_IsSymmetric (const Real2DArray *self, const Real *tolerance, Real *deviation) {
    Integer i, j;
    Real    maxd = 0.0;
    if (self->length0 != self->length1) return False;
    for (i = 0; i < self->length0; i++) {
        for (j = 0; j < i; j++) {
            Real d = fabs (Real2DArray_Item (self, i, j) - Real2DArray_Item (self, j, i));
            if (d > maxd) maxd = d;
        }
    }
    if (deviation != NULL) *deviation = maxd;
    return (maxd <= *tolerance) ? True : False;
}

SymmetricMatrix *SymmetricMatrix_Allocate (const Integer dimension) {
    SymmetricMatrix *self = (SymmetricMatrix *) malloc (sizeof (SymmetricMatrix));
    if (self != NULL) {
        self->dimension = dimension;
        self->size      = (dimension * (dimension + 1)) / 2;
        self->data      = (Real *) calloc ((size_t) (self->size > 0 ? self->size : 1), sizeof (Real));
        if (self->data == NULL) { free (self); self = NULL; }
    }
    return self;
}
void SymmetricMatrix_Deallocate (SymmetricMatrix **self) {
    if (self != NULL && *self != NULL) { free ((*self)->data); free (*self); *self = NULL; }
}
/* (W_ij + W_ji)/2, pinned by twosites.out: Wmax 5.954225 = mean of -5.97112 and -5.93733 */
void SymmetricMatrix_CopyFromReal2DArray (SymmetricMatrix *self, const Real2DArray *other, Status *status) {
    Integer i, j;
    if (other->length0 != self->dimension || other->length1 != self->dimension) {
        Status_Set (status, Status_ArrayNonConformableSizes);
        return;
    }
    for (i = 0; i < self->dimension; i++) {
        for (j = 0; j <= i; j++) {
            self->data[(i * (i + 1)) / 2 + j] = 0.5 * (Real2DArray_Item (other, i, j) + Real2DArray_Item (other, j, i));
        }
    }
}
void SymmetricMatrix_Set (SymmetricMatrix *self, const Real value) { Integer i; for (i = 0; i < self->size; i++) self->data[i] = value; }
void SymmetricMatrix_Scale (SymmetricMatrix *self, const Real value) { Integer i; for (i = 0; i < self->size; i++) self->data[i] *= value; }

/* MT19937 */
RandomNumberGenerator *RandomNumberGenerator_Allocate (const RandomNumberGeneratorType type) {
    RandomNumberGenerator *self = (RandomNumberGenerator *) malloc (sizeof (RandomNumberGenerator));
    (void) type;
    if (self != NULL) RandomNumberGenerator_SetSeed (self, 5489u);
    return self;
}
void RandomNumberGenerator_Deallocate (RandomNumberGenerator **self) {
    if (self != NULL && *self != NULL) { free (*self); *self = NULL; }
}
void RandomNumberGenerator_SetSeed (RandomNumberGenerator *self, const Cardinal seed) {
    int i;
    self->mt[0] = seed;
    for (i = 1; i < 624; i++) {
        self->mt[i] = 1812433253u * (self->mt[i - 1] ^ (self->mt[i - 1] >> 30)) + (unsigned int) i;
    }
    self->mti = 624;
}
Cardinal RandomNumberGenerator_NextCardinal (const RandomNumberGenerator *cself) {
    RandomNumberGenerator *self = (RandomNumberGenerator *) cself;
    unsigned int y;
    if (self->mti >= 624) {
        int k;
        for (k = 0; k < 624; k++) {
            y = (self->mt[k] & 0x80000000u) | (self->mt[(k + 1) % 624] & 0x7fffffffu);
            self->mt[k] = self->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        self->mti = 0;
    }
    y  = self->mt[self->mti++];
    y ^= (y >> 11);
    y ^= (y << 7)  & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
Real RandomNumberGenerator_NextReal (const RandomNumberGenerator *self) {
    return (Real) RandomNumberGenerator_NextCardinal (self) * (1.0 / 4294967296.0);
}
