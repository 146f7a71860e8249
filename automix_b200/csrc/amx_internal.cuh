// amx_internal.cuh -- host-side objects behind the opaque handles of include/amx.h.
#pragma once

#include "amx_common.cuh"

struct amx_proposal {
  amx_fam_hdr hdr;   // host copy (shapes, offsets)
  void *blob_dev;    // [amx_fam_hdr | doubles] on the device
  int blob_bytes;
};

namespace amx {
int upload_blob(const amx_fam_hdr &h, const double *data, void **dev, int *bytes);
}
