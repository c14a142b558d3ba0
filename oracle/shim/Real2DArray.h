#ifndef _SHIM_REAL2DARRAY
#define _SHIM_REAL2DARRAY
#include "Real.h"
#include "Integer.h"
#include "Boolean.h"
#include "Status.h"
typedef struct { Integer length0, length1; Real *data; } Real2DArray;
#define Real2DArray_Item(self, i, j) ((self)->data[(i) * (self)->length1 + (j)])
extern Real2DArray *Real2DArray_Allocate    (const Integer length0, const Integer length1, Status *status);
extern void         Real2DArray_Deallocate  (Real2DArray **self);
/* True iff max |a_ij - a_ji| <= tolerance; pinned by the fixture logs: ifp20 reports a maximum
 * deviation of 0.0533 (= max |W_ij - W_ji|) as NOT symmetric at 0.05, twosites' 0.0406 as symmetric. */
extern Boolean      Real2DArray_IsSymmetric (const Real2DArray *self, const Real *tolerance, Real *deviation);
#endif
