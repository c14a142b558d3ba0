#ifndef _SHIM_SYMMETRICMATRIX
#define _SHIM_SYMMETRICMATRIX
#include "Real.h"
#include "Integer.h"
#include "Status.h"
#include "Real2DArray.h"
/* packed lower triangle, data[i(i+1)/2 + j], i >= j (EnergyModel.h:35-37 indexes it that way) */
typedef struct { Integer dimension, size; Real *data; } SymmetricMatrix;
extern SymmetricMatrix *SymmetricMatrix_Allocate            (const Integer dimension);
extern void             SymmetricMatrix_Deallocate          (SymmetricMatrix **self);
extern void             SymmetricMatrix_CopyFromReal2DArray (SymmetricMatrix *self, const Real2DArray *other, Status *status);
extern void             SymmetricMatrix_Set                 (SymmetricMatrix *self, const Real value);
extern void             SymmetricMatrix_Scale               (SymmetricMatrix *self, const Real value);
#endif
