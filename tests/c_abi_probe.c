/* A plain C99 caller of the drop-in boundary: include/smpl_b200.h must compile as C (no C++-isms, no torch types)
 * and libsmpl_b200.so must link from C.  No compute call is made (this runs on boxes without a GPU).
 * Built and run by tests/test_capi_cpu.py::test_header_is_plain_c_and_links_from_c. */
#include "smpl_b200.h"

#include <stdio.h>
#include <string.h>

int main(void) {
  SmplB200Model* model = 0;
  SmplB200ModelDesc desc;
  SmplB200ForwardOpts opts;
  int st;
  memset(&desc, 0, sizeof desc);
  memset(&opts, 0, sizeof opts);
  opts.struct_size = (uint32_t)sizeof opts;
  printf("version %u\n", (unsigned)smplb200_version());
  st = smplb200_model_create(0, &model);               /* NULL descriptor: invalid argument, never a crash */
  printf("create(NULL) %d %s\n", st, smplb200_strerror(st));
  desc.struct_size = 4;                                 /* wrong struct size: rejected before anything is read */
  st = smplb200_model_create(&desc, &model) == 1 ? st : 99;
  smplb200_model_destroy(0);                            /* no-op */
  return (st == 1 && model == 0 && smplb200_workspace_bytes(0, 16, SMPLB200_PREC_AUTO) == 0) ? 0 : 1;
}
