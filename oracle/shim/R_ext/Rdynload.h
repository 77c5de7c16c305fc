/* TEST INFRASTRUCTURE ONLY (oracle): stand-in for <R_ext/Rdynload.h>, enough
 * for the reference's model trampolines
 * (/root/reference/inst/include/dde/dde.c:10-78). */
#ifndef DDE_ORACLE_SHIM_RDYNLOAD_H
#define DDE_ORACLE_SHIM_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
DL_FUNC R_GetCCallable(const char *package, const char *name);
#endif
