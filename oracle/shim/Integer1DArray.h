#ifndef _SHIM_INTEGER1DARRAY
#define _SHIM_INTEGER1DARRAY
#include "Integer.h"
#include "Status.h"
typedef struct { Integer length; Integer *data; } Integer1DArray;
#define Integer1DArray_Item(self, i) ((self)->data[(i)])
extern Integer1DArray *Integer1DArray_Allocate   (const Integer length, Status *status);
extern void            Integer1DArray_Deallocate (Integer1DArray **self);
#endif
