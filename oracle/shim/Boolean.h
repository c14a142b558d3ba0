#ifndef _SHIM_BOOLEAN
#define _SHIM_BOOLEAN
typedef enum { False = 0, True = 1 } Boolean;
#endif
