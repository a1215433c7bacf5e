/* A plain-C consumer of include/gpxb200.h, linked against gpx_b200/libgpxb200.so (tests/test_abi_cpu.py builds and
 * runs it).  It shows that the boundary is a C ABI -- no C++ or torch types in the signatures -- and exercises the
 * error contract without a GPU: context creation on a machine with no device must fail with a positive CUDA error
 * code and leave a message for gpxb_last_error(); with a GPU it creates a context, checks that a null operand is
 * reported as an argument error (< 0) instead of crashing, and destroys the context. */
#include <stdio.h>
#include <string.h>

#include "gpxb200.h"

int main(void) {
  char msg[256];
  gpxb_ctx* ctx = NULL;
  int rc;
  printf("version %d\n", gpxb_version());
  rc = gpxb_ctx_create(0, &ctx);
  if (rc != 0) {
    msg[0] = '\0';
    gpxb_last_error(msg, sizeof msg);
    printf("ctx_create rc %d msg [%s]\n", rc, msg);
    /* argument errors are negative and never touch the device */
    rc = gpxb_ctx_create(0, NULL);
    printf("ctx_create(NULL) rc %d\n", rc);
    return (rc < 0 && strlen(msg) > 0) ? 0 : 1;
  }
  printf("ctx ok launches %llu\n", gpxb_ctx_launches(ctx));
  /* a call with a null operand is an argument error (< 0), not a crash */
  rc = gpxb_gram(ctx, GPXB_SE, NULL, 4, NULL, 4, 2, 1.0, 0.0, NULL, 4, 0, NULL);
  printf("gram(NULL) rc %d\n", rc);
  if (rc >= 0) return 2;
  return gpxb_ctx_destroy(ctx);
}
