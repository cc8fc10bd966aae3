/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Stand-in for the reference's missing `../helper2.h` (included by
 * /root/reference/smatcher.h:31).  The reference tree does not ship it; the only
 * things ac/ac.c, wu/wu.c, sh/sh.c and sbom/sbom.c need from it are MIN(), MAX() and fail()
 * (wu/wu.c:131, sbom/sbom.c:190, ac/ac.c:46).  Resolved through `-I oracle/ref_shim/inc`
 * ("inc/../helper2.h") so that the reference sources are compiled where they
 * lie, unmodified and uncopied. */
#ifndef ORACLE_REF_SHIM_HELPER2_H
#define ORACLE_REF_SHIM_HELPER2_H
#include <stdio.h>
#include <stdlib.h>
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
static inline void fail(const char *msg) {
	fputs(msg, stderr);
	exit(1);
}
#endif
