#ifndef _SHIM_INTEGER
#define _SHIM_INTEGER
typedef int Integer;
#endif
