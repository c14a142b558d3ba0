#ifndef _SHIM_STATUS
#define _SHIM_STATUS
typedef enum {
    Status_Continue = 0,
    Status_ArrayNonConformableSizes,
    Status_IndexOutOfRange,
    Status_MemoryAllocationFailure,
    Status_ValueError
} Status;
#define Status_Set(status, value) { if ((status) != NULL) *(status) = (value); }
#endif
