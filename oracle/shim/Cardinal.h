#ifndef _SHIM_CARDINAL
#define _SHIM_CARDINAL
typedef unsigned int Cardinal;
#endif
