#ifndef _SHIM_REAL1DARRAY
#define _SHIM_REAL1DARRAY
#include "Real.h"
#include "Integer.h"
#include "Status.h"
typedef struct { Integer length; Real *data; } Real1DArray;
#define Real1DArray_Item(self, i)        ((self)->data[(i)])
#define Real1DArray_ItemPointer(self, i) (&((self)->data[(i)]))
#define Real1DArray_Data(self)           ((self)->data)
extern Real1DArray *Real1DArray_Allocate   (const Integer length, Status *status);
extern void         Real1DArray_Deallocate (Real1DArray **self);
extern void         Real1DArray_Set        (Real1DArray *self, const Real value);
extern void         Real1DArray_Scale      (Real1DArray *self, const Real value);
extern void         Real1DArray_AddScalar  (Real1DArray *self, const Real value);
extern void         Real1DArray_Exp        (Real1DArray *self);
extern Real         Real1DArray_Sum        (const Real1DArray *self);
#endif
