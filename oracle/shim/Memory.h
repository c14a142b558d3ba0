#ifndef _SHIM_MEMORY
#define _SHIM_MEMORY
#include <stdlib.h>
typedef size_t CSize;
#define MEMORY_ALLOCATE(object, type)              { object = (type *) malloc (sizeof (type)); }
#define MEMORY_ALLOCATEARRAY(object, length, type) { object = (type *) calloc ((CSize) (length), sizeof (type)); }
#define MEMORY_DEALLOCATE(object)                  { free (object); object = NULL; }
#endif
