/* pCore shim (test infrastructure): the reference's C sources include pDynamo pCore headers that are
 * not vendored under /root/reference.  These few headers restate only the types and calls the three
 * hot-path files use (call sites: EnergyModel.c:36-44,79,86,93,100,261-279,323,336;
 * MCModelDefault.c:28-36,80,97,100,302,311; StateVector.c:17,28,38,168). */
#ifndef _SHIM_REAL
#define _SHIM_REAL
typedef double Real;
#endif
